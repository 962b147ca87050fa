O=gpurun_out/prof2; mkdir -p $O
python tools/op_bench.py B=4 w2.fwd w2.dgrad w4.fwd w4.dgrad w6.fwd w6.dgrad w9.fwd w9.dgrad w1.wgrad w2.wgrad w7.wgrad w10.wgrad > $O/config4_s2_layers.txt 2>&1
cat $O/config4_s2_layers.txt
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o $O/ncu_$1 python tools/op_bench.py ${4:-} $2 > $O/ncu_$1.log 2>&1; }
cap w2_dgrad_conv_upw w2.dgrad conv_upw B=4
cap w2_fwd_conv_downw w2.fwd conv_downw B=4
cap w9_fwd_conv_upw w9.fwd conv_upw B=4
for r in $O/ncu_*.ncu-rep; do b=$(basename $r .ncu-rep); { python tools/ncu_summary.py $r; python tools/ncu_hot.py $r 25; } > $O/$b.txt 2>&1; done
rm -f $O/ncu_w9*.ncu-rep
ls -la $O
