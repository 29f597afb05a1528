#!/bin/bash
# 2 GPUs: DataParallel replicas through the engine, data-parallel training benches, patch-sharded inference
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_options.py -q -s -k "data_parallel" -p no:cacheprovider ) > gpurun_out/g_dp_test.log 2>&1
for wl in cfg2 cfg3 cfg5; do
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/g_bench_${wl}_n2.json 2> gpurun_out/g_bench_${wl}_n2.err
done
( NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/g_nccl_info.log 2>&1
tail -3 gpurun_out/g_dp_test.log
for f in gpurun_out/g_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
grep -E "NVLS|NVLink|via P2P|Connected all" gpurun_out/g_nccl_info.log | head -5
