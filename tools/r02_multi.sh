#!/bin/bash
# round 2, multi-GPU call: N = $1 GPUs.  2-GPU pytest (in-process multi + torchrun-style), bench both drivers, reference arm.
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02m_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r02m_pytest_multi_n$N.log 2>&1
echo "pytest rc=$?" >> $O/r02m_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > $O/r02m_bench_torchrun_n$N.json 2> $O/r02m_bench_torchrun_n$N.err
echo "torchrun rc=$?" >> $O/r02m_bench_torchrun_n$N.err
timeout 900 python bench.py --gpus $N --driver c --steps 20 --warmup 3 > $O/r02m_bench_c_n$N.json 2> $O/r02m_bench_c_n$N.err
echo "driver c rc=$?" >> $O/r02m_bench_c_n$N.err
timeout 600 python tools/ubench_h2d_multi.py $N 1.0 $2 > $O/r02m_ubench_h2d_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/r02m_bench_ref_n$N.json 2> $O/r02m_bench_ref_n$N.err
tail -8 $O/r02m_pytest_multi_n$N.log
for f in torchrun c ref; do echo "== $f"; cut -c1-1500 $O/r02m_bench_${f}_n$N.json; tail -3 $O/r02m_bench_${f}_n$N.err; done
cat $O/r02m_ubench_h2d_n$N.txt | grep aggregate
if [ "$N" = "2" ]; then
  for v in 0 1 2 3; do PTM_RELIGHT_RGB_VARIANT=$v python tools/relight_once.py 360 rgb fast | sed "s/^/rgb variant $v: /"; done > $O/r02m_relight_rgb.txt 2>&1
  python tools/relight_once.py 360 lrgb fast >> $O/r02m_relight_rgb.txt 2>&1
  cat $O/r02m_relight_rgb.txt
fi
