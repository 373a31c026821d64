#!/bin/bash
# round 2, GPU call 5: relight ceilings (tuning build), whole parity suite, bench line with the drop-in breakdown
O=gpurun_out
mkdir -p $O
for dbg in 0 1 2 3 4; do PTM_B200_LIB=$PWD/rti_b200/lib_tuning/libptm_b200.so PTM_RELIGHT_DEBUG=$dbg PTM_RELIGHT_VARIANT=0 python tools/relight_once.py 360 lrgb fast | sed "s/^/debug $dbg: /"; done > $O/r02e_relight_debug.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $O/r02e_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/r02e_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02e_bench_n1.json 2> $O/r02e_bench_n1.err
echo "bench rc=$?" >> $O/r02e_bench_n1.err
cat $O/r02e_relight_debug.txt; tail -6 $O/r02e_pytest_gpu.log; tail -c 800 $O/r02e_bench_n1.err; python -c "
import json; d=json.load(open('$O/r02e_bench_n1.json')); print(json.dumps(d.get('e2e_dropin'))); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
