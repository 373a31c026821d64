#!/bin/bash
# Round-end evidence run on one B200: bench line, ncu launch list of the same command, full ncu captures of
# K1 / K2 (headline path) and of the tensor-core kernel (RGB format), engine / format / relight benches.
# Everything lands in gpurun_out/; tools/ncu_summary.py turns the .ncu-rep files into profiles/*.txt here.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/final_bench_n1.json 2> $O/final_bench_n1.err
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/final_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/final_ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/final_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fit_lrgb_u8_tma|quantise" -s 8 -c 2 -o $O/final_prof_k1k2 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/final_ncu_k1k2.log 2>&1
ENGINES=mma python tools/bench_engines.py 6000 4000 2 > $O/final_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fit_mma -s 12 -c 1 -o $O/final_prof_mma_rgb \
    env ENGINES=mma python tools/bench_engines.py 6000 4000 2 > $O/final_ncu_mma.log 2>&1
python tools/bench_engines.py > $O/final_engines.txt 2> $O/final_engines.err
for e in ffma mma; do echo "engine $e"; PTM_FIT_ENGINE=$e python tools/bench_formats.py 50; done > $O/final_formats.txt 2> $O/final_formats.err
python tools/bench_relight.py > $O/final_relight.txt 2> $O/final_relight.err
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/final_smi.txt
