#!/bin/bash
# 2 GPUs: bucketed allreduce + second-stream backward together (cfg2, cfg3), and the 2-GPU DataParallel test
mkdir -p gpurun_out
for wl in cfg2 cfg3; do
  ( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/n4_bench_${wl}.json 2> gpurun_out/n4_bench_${wl}.err
done
for f in gpurun_out/n4_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; tail -2 ${f%.json}.err; done
