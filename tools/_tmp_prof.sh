O=gpurun_out/r1b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
TEM_BENCH_TAGS=1 timeout 300 python bench.py --wf 1 --dim 110 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-inference > $O/bench_config4.json 2> $O/bench_config4.err; grep KERNEL $O/bench_config4.err
timeout 300 python bench.py --wf 2 --batch 2 --steps 3 --warmup 3 --no-cpu-baseline --no-inference > $O/bench_wf2.json 2> $O/bench_wf2.err; tail -c 300 $O/bench_wf2.json
TEM_BENCH_TAGS=1 timeout 600 python bench.py > $O/bench_n1.json 2> $O/tags.txt; head -c 400 $O/bench_n1.json
