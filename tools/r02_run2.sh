#!/bin/bash
# round 2, GPU call 2: full parity suite, bench line, relight timings + ncu, host-copy ceilings
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r02b_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err
echo "bench rc=$?" >> $O/r02b_bench_n1.err
for f in lrgb rgb; do python tools/relight_once.py 360 $f fast; done > $O/r02b_relight_once.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_lrgb_fast -s 3 -c 1 -o $O/r02b_prof_relight_lrgb_fast \
    python tools/relight_once.py 60 lrgb fast > $O/r02b_ncu_relight_lrgb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_rgb_fast -s 3 -c 1 -o $O/r02b_prof_relight_rgb_fast \
    python tools/relight_once.py 60 rgb fast > $O/r02b_ncu_relight_rgb.log 2>&1
timeout 300 python tools/ubench_h2d.py 2 > $O/r02b_ubench_h2d.txt 2>&1
tail -25 $O/r02b_pytest_gpu.log; cat $O/r02b_relight_once.txt; tail -c 1200 $O/r02b_bench_n1.err; cat $O/r02b_ubench_h2d.txt
