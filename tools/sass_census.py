#!/usr/bin/env python
"""Per-kernel census of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UBLKCP = TMA, HMMA = legacy mma.sync), from the compiled objects in
transfer_em_b200/csrc (no GPU needed):   python tools/sass_census.py > profiles/sass_census_r1.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "ATOMG", "RED")
print("%-34s %s" % ("kernel (object)", "  ".join("%7s" % o for o in OPS)))
for obj in sorted(glob.glob(os.path.join(ROOT, "transfer_em_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = re.sub(r"\(.*", "", name)
            counts[cur] = collections.Counter()
            continue
        if cur:
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                for o in OPS:
                    if op.startswith(o):
                        counts[cur][o] += 1
    for k, c in counts.items():
        if any(c[o] for o in ("UTCHMMA", "UTMALDG", "UBLKCP", "HMMA", "LDTM")):
            print("%-34s %s" % ((k + " (" + os.path.basename(obj)[:-2] + ")")[:34], "  ".join("%7d" % c[o] for o in OPS)))
