#!/bin/bash
# round 2, GPU call 15: HBM traffic of whole encodes, L2 untouched between K1 and K2 (ncu range replay)
O=gpurun_out
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for rows in 500 1000 4000; do
  timeout 300 python tools/range_traffic.py $rows 4 > $O/r02u_plain_$rows.log 2>&1 &&
  timeout 600 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file $O/r02u_range_$rows.csv \
      python tools/range_traffic.py $rows 4 > $O/r02u_range_$rows.log 2>&1
  echo "== rows $rows"; tail -1 $O/r02u_range_$rows.log; grep -v "^==" $O/r02u_range_$rows.csv | tail -4
done
