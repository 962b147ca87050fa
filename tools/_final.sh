O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
python tools/op_bench.py B=4 w1.wgrad w2.wgrad w3.wgrad w4.wgrad w5.wgrad w6.wgrad w7.wgrad w8.wgrad w9.wgrad w10.wgrad > $O/config4_wgrad_layers.txt 2>&1; cat $O/config4_wgrad_layers.txt
TEM_BENCH_TAGS=1 timeout 300 python bench.py --wf 1 --dim 110 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-inference > $O/bench_config4.json 2> $O/config4_tags.txt; grep KERNEL $O/config4_tags.txt | head -7
TEM_BENCH_TAGS=1 timeout 300 python bench.py --wf 2 --batch 2 --steps 3 --warmup 3 --no-cpu-baseline --no-inference > $O/bench_wf2.json 2> $O/wf2_tags.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; head -c 300 $O/bench_n1.json
