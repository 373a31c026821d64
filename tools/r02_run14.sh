#!/bin/bash
# round 2, GPU call 14: DRAM bytes of K1 / K2 with the caches left as the previous kernel left them (single-pass metric
# collection, --cache-control none): what K2 really reads from HBM inside the encode loop, at 24 MP and on a 3 MP slab;
# then three plain repetitions of the headline run (spread)
O=gpurun_out
mkdir -p $O
B="--no-e2e --no-cpu-baseline --no-extra --no-dropin"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:"fit_lrgb_u8_tma|quantise" -s 20 -c 6 --csv --log-file $O/r02t_warm_24mp.csv \
    python bench.py --steps 10 --warmup 5 $B > $O/r02t_warm_24mp.log 2>&1
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:"fit_lrgb_u8_tma|quantise" -s 20 -c 6 --csv --log-file $O/r02t_warm_3mp.csv \
    python bench.py --height 500 --steps 10 --warmup 5 $B > $O/r02t_warm_3mp.log 2>&1
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 3 $B | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks'])"; done > $O/r02t_repeat.txt 2>&1
grep -v "^==" $O/r02t_warm_24mp.csv | cut -d, -f5,13,15 | tail -26; grep -v "^==" $O/r02t_warm_3mp.csv | cut -d, -f5,13,15 | tail -26; cat $O/r02t_repeat.txt
