#!/bin/bash
# Measured artefacts of the wide (wf <= 2) kernels, BASELINE config 4 (run on a B200 box from the repository root:
#   gpurun --timeout 1200 -- 'bash tools/make_profiles_config4.sh'); writes gpurun_out/prof4/, copied to profiles/ by hand.
set -u
O=gpurun_out/prof4
mkdir -p $O
# 1. config 4, one GPU's share: full train step with per-(layer, op) and per-kernel CUDA-event timing
TEM_BENCH_TAGS=1 python bench.py --wf 1 --dim 110 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-inference > $O/bench_config4.json 2> $O/config4_tags.txt
# 2. width sweep point wf = 2
TEM_BENCH_TAGS=1 python bench.py --wf 2 --batch 2 --steps 3 --warmup 3 --no-cpu-baseline --no-inference > $O/bench_wf2.json 2> $O/wf2_tags.txt
# 3. isolated layers of config 4 (batch 4, n = 74 geometry): 3x3x3 and 4x4x4 stride-2 / transposed, all three passes
python tools/op_bench.py B=4 w1.fwd w1.dgrad w1.wgrad w2.fwd w2.dgrad w2.wgrad w3.fwd w3.dgrad w4.fwd w4.dgrad w5.fwd w6.fwd w6.dgrad w7.fwd w7.dgrad w7.wgrad \
    w8.fwd w8.dgrad w9.fwd w9.dgrad w10.fwd w10.dgrad w10.wgrad > $O/config4_layers.txt 2>&1
# 4. ablations (TEM_S2_DBG bits: 1 no epilogue memory traffic / conversion, 2 no weight loads, 4 no input loads, 8 no g loads (wgrad),
#    16 no LeakyReLU' operand loads, 32 no stores): which stage bounds each wide kernel
for d in 0 1 2 4 7 16 32; do echo "== TEM_S2_DBG=$d"; TEM_S2_DBG=$d python tools/op_bench.py B=4 w2.dgrad w9.fwd w2.fwd w9.dgrad w1.fwd w10.dgrad; done > $O/ablation_wide.txt 2>&1
for d in 0 1 4 8 12; do echo "== TEM_S2_DBG=$d"; TEM_S2_DBG=$d python tools/op_bench.py B=4 w1.wgrad w7.wgrad w10.wgrad; done >> $O/ablation_wide.txt 2>&1
echo "== TEM_S2_NO_SWIZZLE=1 (8-channel plane tiles, SWIZZLE_NONE) vs default (128B-swizzled rows), MMA only (TEM_S2_DBG=7)" >> $O/ablation_wide.txt
TEM_S2_NO_SWIZZLE=1 TEM_S2_DBG=7 python tools/op_bench.py B=4 w2.dgrad w9.fwd >> $O/ablation_wide.txt 2>&1
# 5. one full ncu capture per new kernel
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o $O/ncu_$1 python tools/op_bench.py ${4:-} $2 > $O/ncu_$1.log 2>&1; }
cap w2_dgrad_conv_upw w2.dgrad conv_upw B=4
cap w9_fwd_conv_upw w9.fwd conv_upw B=4
cap w2_fwd_conv_downw w2.fwd conv_downw B=4
cap w1_wgrad_tcw w1.wgrad wgrad_tcw B=4
for r in $O/ncu_*.ncu-rep; do b=$(basename $r .ncu-rep); { python tools/ncu_summary.py $r; python tools/ncu_hot.py $r 20; } > $O/$b.txt 2>&1; done
python - <<'PY' > $O/ncu_traffic_config4.json
import csv, glob, io, json, os, subprocess
out = {}
for rep in sorted(glob.glob('gpurun_out/prof4/ncu_*.ncu-rep')):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    def get(name):
        i = h.index(name); x = float(v[i]); unit = u[i]
        return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
    out[os.path.basename(rep)[4:-8]] = {'dram_read_bytes': get('dram__bytes_read.sum'), 'dram_write_bytes': get('dram__bytes_write.sum'),
                                        'duration_us': float(v[h.index('gpu__time_duration.sum')])}
print(json.dumps(out, indent=1))
PY
rm -f $O/ncu_w9*.ncu-rep $O/ncu_w1*.ncu-rep
ls -la $O
