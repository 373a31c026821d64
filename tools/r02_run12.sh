#!/bin/bash
# round 2, GPU call 12: drop-in sequence (speculative scaling, short slabs in the colour loop): tests + bench leg
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_ptmlib_py.py tests/test_gpu_cli.py -m gpu -x -q > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02p_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-extra > $O/r02p_bench_n1.json 2> $O/r02p_bench_n1.err; echo "rc=$?" >> $O/r02p_bench_n1.err
tail -5 $O/r02p_pytest.log; python -c "
import json; d=json.load(open('$O/r02p_bench_n1.json')); print(json.dumps(d['e2e_dropin'], indent=1)); print(d['e2e']['ms_per_step'])"; tail -2 $O/r02p_bench_n1.err
