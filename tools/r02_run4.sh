#!/bin/bash
# round 2, GPU call 4: relight variants, write ceiling, new tests, drop-in timing
O=gpurun_out
mkdir -p $O
python tools/ubench_write.py > $O/r02d_ubench_write.txt 2>&1
for v in 0 1 2 3; do for f in lrgb rgb; do PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 360 $f fast | sed "s/^/variant $v: /"; done; done > $O/r02d_relight_variants.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_relight_fast.py tests/test_gpu_cli.py tests/test_ptmlib_py.py tests/test_gpu_extras.py -m gpu -q > $O/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r02d_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_lrgb_fast -s 3 -c 1 -o $O/r02d_prof_relight_lrgb_fast \
    python tools/relight_once.py 60 lrgb fast > $O/r02d_ncu_relight_lrgb.log 2>&1
cat $O/r02d_ubench_write.txt $O/r02d_relight_variants.txt; tail -8 $O/r02d_pytest.log
