#!/usr/bin/env python
"""Copies the artefacts written by tools/make_profiles.sh (gpurun_out/prof) into profiles/ under the per-round names
and derives profiles/ncu_traffic.json (DRAM bytes per launch of each captured kernel, read by bench.py)."""
import json, os, re, shutil, subprocess, sys
SRC, DST, R = "gpurun_out/prof", "profiles", "r2"
os.makedirs(DST, exist_ok=True)
cp = lambda a, b: (shutil.copy(os.path.join(SRC, a), os.path.join(DST, b)) if os.path.exists(os.path.join(SRC, a)) else print("missing", a))
cp("bench_n1.json", f"bench_{R}_n1.json"); cp("tags.txt", f"tags_{R}.txt"); cp("launches.csv", f"launches_{R}.csv")
cp("config4_layers.txt", f"config4_layers_{R}.txt"); cp("config3_layers.txt", f"config3_layers_{R}.txt")
cp("bench_reference.json", f"bench_{R}_reference.json"); cp("bench_reference_config5.json", f"bench_{R}_reference_config5.json")
cp("bench_config5.json", f"bench_{R}_config5.json"); cp("bench_config4.json", f"bench_{R}_config4.json"); cp("config4_tags.txt", f"config4_tags_{R}.txt")
if os.path.exists(os.path.join(SRC, "launches.csv")):
    out = subprocess.run([sys.executable, "tools/summarize_launches.py", os.path.join(SRC, "launches.csv")], capture_output=True, text=True).stdout
    open(os.path.join(DST, f"launches_{R}_summary.md"), "w").write(out)
traffic = {}
for f in sorted(os.listdir(SRC)):
    m = re.match(r"ncu_(.*)\.txt$", f)
    if not m:
        continue
    shutil.copy(os.path.join(SRC, f), os.path.join(DST, f"ncu_{R}_{m.group(1)}.txt"))
    txt = open(os.path.join(SRC, f)).read()
    kn = re.search(r"== (?:void )?(?:<unnamed>::)?(\w+)", txt)
    def val(key):
        mm = re.search(key + r"\s+([0-9.,]+) (\w+)", txt)
        if not mm:
            return 0.0
        return float(mm.group(1).replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(mm.group(2), 1)
    dur = re.search(r"gpu__time_duration.sum\s+([0-9.,]+) (\w+)", txt)
    if kn:
        parts = m.group(1).split("_")
        traffic.setdefault(kn.group(1), []).append({"capture": m.group(1), "tag": parts[0] + "." + parts[1], "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                                    "duration": f"{dur.group(1)} {dur.group(2)}" if dur else None})
# bench.py reads {kernel: {"dram_bytes_per_launch": mean over the captured launches}}
for v in traffic.values():
    for c in v:
        c["capture"] = f"ncu_{R}_" + c["capture"]
js = {k: {"dram_bytes_per_launch": sum(c["dram_bytes_per_launch"] for c in v) / len(v), "captures": v} for k, v in traffic.items()}
js["commit"] = subprocess.run(["git", "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()     # commit the captures were taken at (run make_profiles on a clean tree)
json.dump(js, open(os.path.join(DST, "ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in js.items() if isinstance(v, dict)}, indent=1))
