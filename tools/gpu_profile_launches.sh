#!/bin/bash
# ncu launch list (device time per launch, cold-cache & serialised: compare SHARES) of one training step.
mkdir -p gpurun_out
W=3
python - > gpurun_out/launch_count.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
import pytorch3dunet_b200 as P, bench
m = P.get_model(bench.CFG).cuda()
x = torch.rand(2, 1, 128, 128, 128, device="cuda"); t = (torch.rand_like(x) > 0.5).float()
o, l = m(x, return_logits=True); P.losses.bce_dice_loss(l, t).backward(); torch.cuda.synchronize()
print(sum(P.last_launch_counts()))
PY
cat gpurun_out/launch_count.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"
wc -l gpurun_out/launches_r01.csv
