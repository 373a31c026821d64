#!/bin/bash
# round 2: N-GPU torchrun bench line only (the driver's launch), after the scratch float plane became the default
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > $O/r02y_bench_n$N.json 2> $O/r02y_bench_n$N.err
echo "rc=$?" >> $O/r02y_bench_n$N.err
cut -c1-700 $O/r02y_bench_n$N.json; tail -2 $O/r02y_bench_n$N.err
