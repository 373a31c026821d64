#!/bin/bash
# round 2, GPU call 13: stack_stable promise re-measured with the PDL chain intact (no event records in the timed loop); smoke()
O=gpurun_out
mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
{
for ss in 0 1 0 1; do
  echo "== 3 MP slab, PTM_STACK_STABLE=$ss"; PTM_STACK_STABLE=$ss timeout 300 python bench.py --height 500 --steps 200 --warmup 20 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
done
for ss in 0 1; do
  echo "== 24 MP, PTM_STACK_STABLE=$ss"; PTM_STACK_STABLE=$ss timeout 300 python bench.py --steps 50 --warmup 5 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
done
} > $O/r02s_stack_stable.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02s_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02s_smoke.log
cat $O/r02s_stack_stable.txt; tail -3 $O/r02s_smoke.log
