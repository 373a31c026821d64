#!/bin/bash
# round 2, GPU call 16: K2 walking the float plane backwards with L2 discards inside the FUSED encode (PTM_K2_CONSUME=1):
# NOTE: PTM_K2_CONSUME (whole plane backwards) existed only in the intermediate build this script ran against; the shipped form is
# ptmb_set_coeffs_scratch (tail first, bench.py default; PTM_KEEP_COEFFS=1 keeps the float plane) -- see tools/r02_run17.sh
# HBM traffic of whole encodes (range replay) and step time, 24 MP and 6 MP
O=gpurun_out
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
for k in 0 1; do
  for rows in 1000 4000; do
    PTM_K2_CONSUME=$k timeout 600 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r02v_range_${rows}_k$k.csv \
      python tools/range_traffic.py $rows 4 > $O/r02v_range_${rows}_k$k.log 2>&1
    echo "== consume $k rows $rows"; grep -v "^==" $O/r02v_range_${rows}_k$k.csv | tail -3 | cut -d, -f11,13
  done
  for i in 1 2; do
  echo "== consume $k, 24 MP step"; PTM_K2_CONSUME=$k timeout 300 python bench.py --steps 50 --warmup 5 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
  done
  echo "== consume $k, 6 MP step"; PTM_K2_CONSUME=$k timeout 300 python bench.py --height 1000 --steps 100 --warmup 10 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['k1_ms'], d['roofline']['frac'])"
done > $O/r02v_consume.txt 2>&1
cat $O/r02v_consume.txt
