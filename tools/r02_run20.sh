#!/bin/bash
# round 2, GPU call 20: L2 policy of K1's coefficient stores (0 none, 1 evict_first, 2 evict_last = shipped) with the scratch float plane
O=gpurun_out
mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
{
for hnt in 2 0 1 2; do
  for h in 4000 1000 500; do
    st=$((200000 / h)); echo "== store hint $hnt, $h rows"; PTM_K1_STORE_HINT=$hnt timeout 300 python bench.py --height $h --steps $st --warmup 5 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['frac'])"
  done
done
} > $O/r02ab_store_hint.txt 2>&1
cat $O/r02ab_store_hint.txt
