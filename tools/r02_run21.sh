#!/bin/bash
# round 2, GPU call 21 (experiment): persisting-L2 window over the end of the scratch float plane (PTM_L2_PERSIST_MB), with and
# without the evict_last hint on K1's coefficient stores
O=gpurun_out
mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
{
for cfg in "0 2" "64 2" "96 2" "96 0" "64 0" "0 2"; do
  set -- $cfg; pm=$1; hnt=$2
  for h in 4000 1000 500; do
    st=$((200000 / h)); echo "== persist $pm MB, store hint $hnt, $h rows"; PTM_L2_PERSIST_MB=$pm PTM_K2_TAIL_MB=${pm/#0/96} PTM_K1_STORE_HINT=$hnt timeout 300 python bench.py --height $h --steps $st --warmup 5 $B 2>$O/r02ad_err.txt | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['frac'])"; grep persisting $O/r02ad_err.txt | head -1
  done
done
} > $O/r02ad_persist.txt 2>&1
cat $O/r02ad_persist.txt
