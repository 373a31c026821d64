#!/bin/bash
# round 2, GPU call 6: relight shapes (4 / 8 pixels per thread, narrow / wide stores) + their ceilings, parity per variant
O=gpurun_out
mkdir -p $O
T=$PWD/rti_b200/lib_tuning/libptm_b200.so
{
for v in 0 1 4 5 6 7; do PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 360 lrgb fast | sed "s/^/variant $v: /"; done
for v in 0 4; do for dbg in 1 2 4 5; do PTM_B200_LIB=$T PTM_RELIGHT_DEBUG=$dbg PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 360 lrgb fast | sed "s/^/variant $v debug $dbg: /"; done; done
} > $O/r02f_relight.txt 2>&1
for v in 0 3 4 5 7; do echo "== PTM_RELIGHT_VARIANT=$v"; PTM_RELIGHT_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_relight_fast.py -m gpu -q -k "not full_size" 2>&1 | tail -3; done > $O/r02f_pytest_variants.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sixteen_bit" > $O/r02f_pytest_16.log 2>&1
cat $O/r02f_relight.txt $O/r02f_pytest_variants.log; tail -3 $O/r02f_pytest_16.log
