#!/bin/bash
# round 2, GPU call 8: full GPU test suite, N=1 bench line, ncu launch list + full capture of K1/K2 of the shipped kernels
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02z_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02z_bench_n1.json 2> $O/r02z_bench_n1.err; echo "rc=$?" >> $O/r02z_bench_n1.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --no-dropin > $O/r02z_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02z_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --no-dropin > $O/r02z_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fit_lrgb_u8_tma|quantise" -s 8 -c 2 -f -o $O/r02z_prof_k1k2 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --no-dropin > $O/r02z_ncu_k1k2.log 2>&1
tail -5 $O/r02z_pytest_gpu.log; cut -c1-600 $O/r02z_bench_n1.json; tail -2 $O/r02z_bench_n1.err
