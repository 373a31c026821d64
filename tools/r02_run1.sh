#!/bin/bash
# round 2, GPU call 1: parity tests, bench line, relight timings + one ncu capture of each tolerance kernel
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r02_smi.txt
nproc > $O/r02_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
echo "bench rc=$?" >> $O/r02_bench_n1.err
for f in lrgb rgb; do for m in fast exact; do python tools/relight_once.py 360 $f $m; done; done > $O/r02_relight_once.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_lrgb_fast -s 3 -c 1 -o $O/r02_prof_relight_lrgb_fast \
    python tools/relight_once.py 60 lrgb fast > $O/r02_ncu_relight_lrgb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_rgb_fast -s 3 -c 1 -o $O/r02_prof_relight_rgb_fast \
    python tools/relight_once.py 60 rgb fast > $O/r02_ncu_relight_rgb.log 2>&1
tail -5 $O/r02_pytest_gpu.log; cat $O/r02_relight_once.txt; tail -c 1500 $O/r02_bench_n1.err
