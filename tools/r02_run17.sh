#!/bin/bash
# round 2, GPU call 17: tail-first consuming K2 inside the fused encode: parity test, HBM traffic (range replay), step time
# NOTE: run17 drove the scratch plane with PTM_K2_CONSUME=0|1 of the intermediate build; today: PTM_KEEP_COEFFS=1|0
O=gpurun_out
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "consum or lrgb_encode or fused or interleave" > $O/r02w_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02w_pytest.log
{
for cfg in "0 64" "1 32" "1 64" "1 96"; do
  set -- $cfg; k=$1; t=$2
  for rows in 4000; do
    PTM_K2_CONSUME=$k PTM_K2_TAIL_MB=$t timeout 600 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r02w_range.csv \
      python tools/range_traffic.py $rows 4 > $O/r02w_range.log 2>&1
    echo "== consume $k tail $t MB rows $rows"; grep -v "^==" $O/r02w_range.csv | tail -3 | cut -d, -f11,13
  done
  for h in 4000 4000 1000 500; do
    st=$((200000 / h)); echo "== consume $k tail $t MB, $h rows"; PTM_K2_CONSUME=$k PTM_K2_TAIL_MB=$t timeout 300 python bench.py --height $h --steps $st --warmup 5 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
  done
done
} > $O/r02w_tail.txt 2>&1
tail -3 $O/r02w_pytest.log; cat $O/r02w_tail.txt
