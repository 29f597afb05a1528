#!/bin/bash
# memcheck of the round's new non-TMA kernels (small test shapes), then the default bench line once more (traffic field from the final ncu summary)
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "stem or maxpool or act_bwd or fused_stats or gn_bwd" ) > gpurun_out/final2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/final2_memcheck.log | head -10
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/final2_bench_cfg2.json 2> gpurun_out/final2_bench_cfg2.err
grep '^{' gpurun_out/final2_bench_cfg2.json | head -c 400; echo
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/final2_bench_ref.json 2> gpurun_out/final2_bench_ref.err
grep '^{' gpurun_out/final2_bench_ref.json | head -c 600; echo
