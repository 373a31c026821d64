#!/bin/bash
# round 2, GPU call 9: 512-pixel-tile ring kernel (parity, 3 MP slab and 24 MP timings against the 1024-pixel tiles)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mma.py -m gpu -x -q > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02i_pytest.log
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
{
for t in 1024 512; do for v in 0 2 4; do
  echo "== 3 MP slab, tile $t, variant $v"; PTM_TMA_TILE=$t PTM_TMA_VARIANT=$v timeout 300 python bench.py --height 500 --steps 200 --warmup 20 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
done; done
for t in 1024 512; do
  echo "== 24 MP, tile $t"; PTM_TMA_TILE=$t timeout 300 python bench.py --steps 20 --warmup 3 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
  echo "== 6 MP, tile $t"; PTM_TMA_TILE=$t timeout 300 python bench.py --height 1000 --steps 100 --warmup 10 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
  echo "== 12 MP, tile $t"; PTM_TMA_TILE=$t timeout 300 python bench.py --height 2000 --steps 50 --warmup 5 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
done
} > $O/r02i_tiles.txt 2>&1
tail -4 $O/r02i_pytest.log; cat $O/r02i_tiles.txt
