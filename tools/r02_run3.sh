#!/bin/bash
# round 2, GPU call 3: parity suite, bench line, relight timings, 3 MP slab timing (8-GPU proxy) with / without early producer
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r02c_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err
echo "bench rc=$?" >> $O/r02c_bench_n1.err
for f in lrgb rgb; do python tools/relight_once.py 360 $f fast; done > $O/r02c_relight_once.txt 2>&1
for s in 0 1; do PTM_STACK_STABLE=$s python bench.py --height 500 --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-extra; done > $O/r02c_slab3mp.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fit_lrgb_u8_tma|quantise" -s 40 -c 4 -o $O/r02c_prof_slab3mp \
    python bench.py --height 500 --steps 10 --warmup 20 --no-e2e --no-cpu-baseline --no-extra > $O/r02c_ncu_slab.log 2>&1
tail -12 $O/r02c_pytest_gpu.log; cat $O/r02c_relight_once.txt; tail -c 600 $O/r02c_bench_n1.err; cat $O/r02c_slab3mp.txt | cut -c1-400
