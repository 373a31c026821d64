#!/bin/bash
# round 2, GPU call 11: tensor-core RGB relight kernel shapes (timing only), then ncu of the chosen one
# NOTE: the PTM_RELIGHT_MMA_SHAPE numbers of this run are those of the tuning build (legend in profiles/r02_relight_rgb_mma.txt)
O=gpurun_out
mkdir -p $O
{
for sh in ${SHAPES:-0 7 8 9 10 11 12 13 14}; do PTM_RELIGHT_MMA_SHAPE=$sh python tools/relight_once.py 360 rgb fast | sed "s/^/mma shape $sh: /"; done
} > $O/r02k_relight.txt 2>&1
cat $O/r02k_relight.txt
if [ -n "$NCU_SHAPE" ]; then
PTM_RELIGHT_MMA_SHAPE=$NCU_SHAPE timeout 600 ncu --set full --clock-control none --import-source on -k regex:relight_rgb_mma -c 1 -f -o $O/r02k_prof_relight_rgb_mma \
    python tools/relight_once.py 360 rgb fast > $O/r02k_ncu.log 2>&1
fi
