#!/bin/bash
# round 2, GPU call 19: LRGB tolerance relight with 512-byte-aligned store order (variant 13) against the shipped variant 5
O=gpurun_out
mkdir -p $O
{
for v in 5 13 5 13; do PTM_RELIGHT_VARIANT=$v python tools/relight_once.py 360 lrgb fast | sed "s/^/variant $v: /"; done
PTM_RELIGHT_VARIANT=13 timeout 600 python -m pytest tests/test_gpu_relight_fast.py -m gpu -q -k "lrgb or full_size_sweep" 2>&1 | tail -2
} > $O/r02aa_relight_align.txt 2>&1
cat $O/r02aa_relight_align.txt
