#!/bin/bash
# round 2, GPU call 10: tensor-core RGB relight kernel (parity, chunk shapes, FFMA kernel beside it, ncu capture)
# NOTE: the PTM_RELIGHT_MMA_SHAPE numbers of this run are those of the tuning build (legend in profiles/r02_relight_rgb_mma.txt)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_relight_fast.py -m gpu -x -q -s > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02j_pytest.log
{
for sh in 0 1 2 3 4 5 6 7; do PTM_RELIGHT_MMA_SHAPE=$sh python tools/relight_once.py 360 rgb fast | sed "s/^/mma shape $sh: /"; done
PTM_RELIGHT_RGB_ENGINE=ffma python tools/relight_once.py 360 rgb fast | sed "s/^/ffma: /"
python tools/relight_once.py 360 lrgb fast
} > $O/r02j_relight.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_rgb_mma -c 1 -f -o $O/r02j_prof_relight_rgb_mma \
    python tools/relight_once.py 360 rgb fast > $O/r02j_ncu.log 2>&1
tail -25 $O/r02j_pytest.log; cat $O/r02j_relight.txt
