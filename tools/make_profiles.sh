#!/bin/bash
# Regenerates the measured artefacts kept under profiles/ (run on a B200 box from the repository root, e.g.
#   gpurun --timeout 1500 -- 'bash tools/make_profiles.sh'
# writes into gpurun_out/; tools/collect_profiles.py then turns the .ncu-rep files into the text summaries of profiles/).
set -u
O=gpurun_out/prof
mkdir -p $O
# 1. headline bench (default flags: config 3 train step + config 5 tiled inference of 1024^3) with per-(layer, op) and per-kernel
#    CUDA-event timing; config 5 and config 4 as their own lines
TEM_BENCH_TAGS=1 python bench.py > $O/bench_n1.json 2> $O/tags.txt
python bench.py --config 5 > $O/bench_config5.json 2> $O/bench_config5.err
TEM_BENCH_TAGS=1 python bench.py --config 4 --no-cpu-baseline > $O/bench_config4.json 2> $O/config4_tags.txt
# 2. launch list of the same step (one ncu pass, durations only; cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference > $O/ncu_launch.log 2>&1
# 3. one full capture per top kernel at the bench shapes (isolated op, tools/op_bench.py)
cap() {  # name op kernel-regex [extra op_bench args]
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o $O/ncu_$1 python tools/op_bench.py ${4:-} $2 > $O/ncu_$1.log 2>&1
}
cap g1_fwd_conv3_tc3 g1.fwd conv3_tc3
cap g10_dgrad_conv3_tc3 g10.dgrad conv3_tc3
cap g1_wgrad_tc g1.wgrad wgrad_tc_kernel
cap g7_wgrad_tc g7.wgrad wgrad_tc_kernel
cap g2_dgrad_conv_up g2.dgrad conv_up_tc
cap g2_fwd_conv_down g2.fwd conv_down_tc
cap g2_wgrad_tc_s2 g2.wgrad wgrad_tc_s2
cap g0_wgrad_c1tc g0.wgrad wgrad_c1tc
cap g0_fwd_conv_c1tc g0.fwd conv_c1tc
cap g11_dgrad_conv_c1tc g11.dgrad conv_c1tc
cap g1_dgrad_conv3_tc3 g1.dgrad conv3_tc3
cap g6_wgrad_tc_s2 g6.wgrad wgrad_tc_s2
cap g11_fwd_conv3_tc3 g11.fwd conv3_tc3
cap w1_fwd_conv3_tcw w1.fwd conv3_tcw B=4
cap w7_fwd_conv3_tcw w7.fwd conv3_tcw B=4
# the fused discriminator tail only exists inside a model pass: capture it from a short train-step run
timeout 300 ncu --set full --clock-control none --import-source on -k regex:disc_tail_bwd -s 8 -c 1 -f -o $O/ncu_dtail_bwd_disc_tail python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference > $O/ncu_dtail_bwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:disc_tail_fwd -s 8 -c 1 -f -o $O/ncu_dtail_fwd_disc_tail python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference > $O/ncu_dtail_fwd.log 2>&1
# 4. BASELINE config 4 layers (wf = 1: 64/128/256 channels), forward and data gradient, batch 4
python tools/op_bench.py B=4 w1.fwd w1.dgrad w3.fwd w3.dgrad w5.fwd w7.fwd w7.dgrad w8.fwd w8.dgrad w10.fwd w10.dgrad > $O/config4_layers.txt 2>&1
# 5. per-op timing of the config 3 layers (isolated, batch 8)
python tools/op_bench.py g0.fwd g0.wgrad g1.fwd g1.dgrad g1.wgrad g2.fwd g2.dgrad g2.wgrad g3.fwd g3.dgrad g3.wgrad g4.fwd g4.dgrad g4.wgrad \
    g5.fwd g5.wgrad g6.fwd g6.dgrad g6.wgrad g7.fwd g7.dgrad g7.wgrad g8.fwd g8.dgrad g8.wgrad g9.fwd g9.dgrad g9.wgrad g10.fwd g10.dgrad g10.wgrad \
    g11.fwd g11.wgrad d0.wgrad d1.fwd d1.dgrad d1.wgrad d4.fwd d4.dgrad d4.wgrad d6.fwd > $O/config3_layers.txt 2>&1
# 6. the CPU arm (train step and per-tile inference loop)
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --impl reference --config 5 --steps 3 > $O/bench_reference_config5.json 2>> $O/bench_reference.err
ls -la $O
# 8. text summaries of the captures (key metrics + opcode histogram + hottest SASS lines); only the four headline
#    reports travel back as .ncu-rep (gpurun merges at most 64 MiB)
for r in $O/ncu_*.ncu-rep; do
  b=$(basename $r .ncu-rep)
  { python tools/ncu_summary.py $r; python tools/ncu_hot.py $r 20; } > $O/$b.txt 2>&1
done
mkdir -p $O/rep
for k in g1_fwd_conv3_tc3 g7_wgrad_tc; do mv $O/ncu_$k.ncu-rep $O/rep/ 2>/dev/null; done
rm -f $O/ncu_*.ncu-rep
