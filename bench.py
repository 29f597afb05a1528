"""Headline benchmark: UNet3D f_maps=32 depth=4 training step (forward + BCEDiceLoss + backward) on 2x1x128^3
patches (BASELINE.json configs[1]) -> patches/sec, whole job over N GPUs (data-parallel replicas + one NCCL
gradient allreduce per step).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # the b200 engine
    python bench.py --impl reference ...                            # the reference's own torch-CPU path (oracle port)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=32, num_levels=4)
BATCH, SIZE = 2, 128
WORKLOAD = "UNet3D f_maps=32 depth=4, batch 2x1x128^3, fwd+BCEDiceLoss+bwd (BASELINE cfg 2)"
# algorithmic work per patch
GFLOP_PER_PATCH_TRAIN = 2841.5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused-loss", type=int, default=int(os.environ.get("B200UNET_FUSED_LOSS", "1")),
                    help="1: BCEDiceLoss through the engine's two-pass kernels (csrc/loss_ops.cu) instead of eager torch ops")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "samples": len(self.rows), "reasons": sorted(reasons)}


def cpu_reference_steps(size, batch, steps, warmup, threads=None):
    """The reference's own path = torch CPU ops; timed through the oracle port (the reference package itself cannot
    travel to the GPU box).  Returns seconds per step (median)."""
    import torch
    from oracle import unet3d_oracle as O
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    sd = {k: v.requires_grad_(True) for k, v in O.random_state_dict(CFG, seed=0).items()}
    x = torch.rand(batch, 1, size, size, size)
    t = (torch.rand(batch, 1, size, size, size) > 0.5).float()
    times = []
    for i in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        _, logits = O.forward(sd, CFG, x)
        O.bce_dice_loss(logits, t).backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    return times[len(times) // 2], torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: batch 1 of the same patch size; K steps of ~7 s each on 8 cores
    steps = min(args.steps, 3)
    warm = min(args.warmup, 1)
    sec, threads = cpu_reference_steps(args.size, 1, steps, warm)
    val = 1.0 / sec
    line = {"metric": "UNet3D patches/sec (1x128^3 bf16) train step", "value": val, "unit": "patches/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "sample": f"batch 1 of the {args.size}^3 patch per step (the reference's torch-CPU ops via the oracle port)"},
            "cpu_baseline": {"value": val, "unit": "patches/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} steps x 1 patch {args.size}^3, fp32, torch CPU ({threads} threads)"},
            "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


ALLOC_SETTLE_STEPS = 8


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pytorch3dunet_b200 as P
    from pytorch3dunet_b200 import engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)  # identical replicas on every rank (the reference's DataParallel broadcasts rank-0 weights)
    model = P.get_model(CFG).to(dev)
    params = [p for p in model.parameters()]
    torch.manual_seed(1000 + rank)  # rank-dependent synthetic patches (weak scaling: per-GPU batch fixed)
    B, S = args.batch, args.size
    x_host = torch.rand(B, 1, S, S, S).pin_memory()
    t_host = (torch.rand(B, 1, S, S, S) > 0.5).float().pin_memory()
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)
    from pytorch3dunet_b200.parallel import GradAllReducer
    reducer = GradAllReducer(params, world)

    def step(x, t):
        for p in params:
            p.grad = None
        out, logits = model(x, return_logits=True)
        loss = P.losses.bce_dice_loss(logits, t, fused=bool(args.fused_loss))
        loss.backward()
        reducer()  # one gradient allreduce per step over NVLink (replaces DataParallel's reduce-to-GPU-0, trainer.py:203-204)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # setup, not warm-up: let torch's caching allocator reach its steady state (it keeps cudaMalloc-ing new segments -- a device
    # synchronisation each -- for the first ~8 identical steps while its best-fit split pattern converges; tools/alloc_probe.py)
    for _ in range(ALLOC_SETTLE_STEPS):
        step(x_dev, t_dev)
    for _ in range(max(args.warmup, 3)):
        step(x_dev, t_dev)
    barrier()
    fwd_l, bwd_l = P.last_launch_counts()

    # ---- timed region 1: inputs resident in HBM; per-launch CUDA events on the conv kernels (roofline leg) ----
    E.TIMING = []
    with ClockSampler(local) as clk:
        ms = timed(lambda: step(x_dev, t_dev), args.steps)
    timing, E.TIMING = E.TIMING, None
    patches = world * B * args.steps
    value = patches / (ms / 1e3)

    # ---- timed region 2: end to end through the public nn.Module API with HOST buffers ----
    def e2e_step():
        x = x_host.to(dev, non_blocking=True)
        t = t_host.to(dev, non_blocking=True)
        return step(x, t).item()  # device -> host read of the loss, as trainer.py:241 does every iteration

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = patches / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    by = {}
    for tag, fl, a, b in timing:
        d = by.setdefault(tag, [0.0, 0.0, 0])
        d[0] += fl
        d[1] += a.elapsed_time(b)
        d[2] += 1
    tc = {k: v for k, v in by.items() if k.endswith("_tc")}
    tc_flops = sum(v[0] for v in tc.values())
    tc_ms = sum(v[1] for v in tc.values())
    achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    traffic = None
    try:  # dram bytes of one launch of the largest-share kernel, from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_r01_full_summary.json")))["wgrad_halo_kernel"]
        traffic = (float(prof["dram__bytes_read.sum"]["value"]) + float(prof["dram__bytes_write.sum"]["value"])) * 1e6
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "tcgen05 conv kernels: conv3_halo_kernel / conv3_igemm_kernel (fprop+dgrad), wgrad_halo_kernel / conv3_wgrad_igemm_kernel",
                "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_note": "dram read+write bytes of ONE wgrad_halo_kernel<32> launch (32->32 @ 2x128^3; algorithmic 537 MB: x + dz), profiles/ncu_r01_full_summary.md",
                "peak_source": pk_src + " sustained bf16",
                "share_of_step": tc_ms / ms if ms else None,
                "per_kernel": {k: {"tflops": v[0] / (v[1] / 1e3) / 1e12 if v[1] else None, "ms_per_step": v[1] / args.steps, "launches_per_step": v[2] / args.steps}
                               for k, v in by.items()}}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sec, threads = cpu_reference_steps(S, 1, 2, 1)
        cpu_baseline = {"value": 1.0 / sec, "unit": "patches/s", "cores": threads, "kind": "port",
                        "sample": f"2 steps x 1 patch {S}^3 fwd+BCEDice+bwd, fp32 torch CPU ({threads} threads) via oracle/unet3d_oracle.py"}

    line = {"metric": "UNet3D patches/sec (1x128^3 bf16) train step", "value": value, "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD if (S == SIZE and B == BATCH) else f"UNet3D f32 d4 batch {B}x1x{S}^3", "per_gpu_batch": B,
                       "parallelism": f"dp{world}", "l2": "per-step working set (~1 GB of bf16 activations per patch) >> 126 MB L2; no explicit flush",
                       "setup_steps_before_warmup": ALLOC_SETTLE_STEPS,
                       "loss": "fused b200 kernels" if args.fused_loss else "torch ops"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "patches/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": x_host.numel() * 4 + t_host.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": (fwd_l + bwd_l + (3 if args.fused_loss else 0)) * args.steps, "clocks": clk.summary(),
            "tflops_effective": value * GFLOP_PER_PATCH_TRAIN / 1e3}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
