#!/bin/bash
# round 2, GPU call 18: K2 tail size sweep beyond 96 MB, then the full final N=1 evidence run
O=gpurun_out
mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
{
for t in 96 128 192 96; do
  for h in 4000 1000; do
    st=$((200000 / h)); echo "== tail $t MB, $h rows"; PTM_K2_TAIL_MB=$t timeout 300 python bench.py --height $h --steps $st --warmup 5 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['frac'])"
  done
done
} > $O/r02z_tail_sweep.txt 2>&1
cat $O/r02z_tail_sweep.txt
bash tools/r02_run8.sh
