#!/bin/bash
# round 2, GPU call 7: block-flush relight kernel (variants 9 = 4 px, 13 = 8 px) + ceilings, parity
O=gpurun_out
mkdir -p $O
T=$PWD/rti_b200/lib_tuning/libptm_b200.so
{
for v in 5 9 13; do PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 360 lrgb fast | sed "s/^/variant $v: /"; done
for v in 9 13; do for dbg in 2 5; do PTM_B200_LIB=$T PTM_RELIGHT_DEBUG=$dbg PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 360 lrgb fast | sed "s/^/variant $v debug $dbg: /"; done; done
for v in 9 13; do PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 60 lrgb fast | sed "s/^/variant $v: /"; done
} > $O/r02g_relight.txt 2>&1
for v in 9 13; do echo "== PTM_RELIGHT_VARIANT=$v"; PTM_RELIGHT_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_relight_fast.py -m gpu -q 2>&1 | tail -3; done > $O/r02g_pytest_variants.log 2>&1
cat $O/r02g_relight.txt $O/r02g_pytest_variants.log
