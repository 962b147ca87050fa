set -u
O=gpurun_out/prof_mid
mkdir -p $O
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o $O/ncu_$1 python tools/op_bench.py ${4:-} $2 > $O/ncu_$1.log 2>&1; }
python tools/op_bench.py g1.fwd g0.fwd g7.wgrad g2.dgrad > $O/plain.txt 2>&1 || exit 1
cap g1_fwd_conv3_tc3 g1.fwd conv3_tc3
cap g0_fwd_conv_c1 g0.fwd conv_c1in
cap g7_wgrad_tc g7.wgrad wgrad_tc_kernel
cap g2_dgrad_conv_up g2.dgrad conv_up_tc
for r in $O/ncu_*.ncu-rep; do b=$(basename $r .ncu-rep); { python tools/ncu_summary.py $r; python tools/ncu_hot.py $r 25; } > $O/$b.txt 2>&1; done
rm -f $O/*.ncu-rep
cat $O/plain.txt | grep -v Warn
