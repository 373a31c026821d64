#!/bin/bash
# round 2, final check of the committed tree on one B200: GPU tests, smoke(), the driver's bench command
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02fin_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02fin_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02fin_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02fin_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02fin_bench_n1.json 2> $O/r02fin_bench_n1.err; echo "rc=$?" >> $O/r02fin_bench_n1.err
tail -3 $O/r02fin_pytest_gpu.log; tail -2 $O/r02fin_smoke.log; cut -c1-300 $O/r02fin_bench_n1.json; tail -1 $O/r02fin_bench_n1.err
