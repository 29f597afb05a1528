#!/bin/bash
# 8 GPUs: weak-scaling training benches (cfg2, cfg3 = the BASELINE 8-GPU configuration) and patch-sharded inference (cfg5)
mkdir -p gpurun_out
for wl in cfg2 cfg3; do
  ( timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/n8_bench_${wl}.json 2> gpurun_out/n8_bench_${wl}.err
done
for f in gpurun_out/n8_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
