#!/usr/bin/env python
"""Top SASS instructions of an .ncu-rep by executed count and by stall samples, with opcode histogram."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; body = rows[2:]
iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot_i = sum(int(r[iI]) for r in body); tot_s = sum(int(r[iSm]) for r in body)
print("total warp-instructions %d, samples %d, SASS lines %d" % (tot_i, tot_s, len(body)))
ops = collections.Counter(); ops_s = collections.Counter()
for r in body:
    op = r[iS].split()[0] if not r[iS].strip().startswith('@') else r[iS].split()[1]
    op = op.split('.')[0]
    ops[op] += int(r[iI]); ops_s[op] += int(r[iSm])
print("by opcode (instr share | sample share):")
for op, n in ops.most_common(18):
    print("  %-10s %5.1f%% | %5.1f%%" % (op, 100.0 * n / tot_i, 100.0 * ops_s[op] / max(tot_s, 1)))
print("top lines by samples:")
for r in sorted(body, key=lambda r: -int(r[iSm]))[:topn]:
    print("  %6s smp %9s exec  %s" % (r[iSm], r[iI], r[iS].strip()[:90]))
