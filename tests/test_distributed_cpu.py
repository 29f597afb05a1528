"""CPU, world_size 2 on gloo: the host-side logic of the data-parallel path (gradient allreduce, patch sharding).
The engine itself needs a GPU; here the replicas' gradients are synthetic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pytorch3dunet_b200 as P
    from pytorch3dunet_b200.parallel import GradAllReducer, shard_indices
    torch.manual_seed(0)  # identical replicas on every rank
    model = P.get_model(dict(name="UNet3D", in_channels=1, out_channels=1, f_maps=8, num_levels=2))
    params = list(model.parameters())
    g = torch.Generator().manual_seed(100 + rank)  # rank-dependent "gradients"
    local = []
    for p in params:
        p.grad = torch.randn(p.shape, generator=g)
        local.append(p.grad.clone())
    GradAllReducer(params)()
    # expected: mean over ranks, recomputed locally from the known seeds
    ok = True
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for p in params:
        exp = sum(torch.randn(p.shape, generator=gg) for gg in gens) / world
        ok = ok and torch.allclose(p.grad, exp, atol=1e-6)
    # replicas stay bit-identical after an identical optimizer step
    opt = torch.optim.SGD(params, lr=0.1)
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    shards = shard_indices(147, rank, world)  # BASELINE cfg 5: 147 patches
    q.put((rank, ok, same, len(shards), shards[:3]))
    dist.destroy_process_group()


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pytorch3dunet_b200 as P
    from pytorch3dunet_b200.optim import BucketedAllReduce, FlatParameters
    torch.manual_seed(0)
    model = P.get_model(dict(name="ResidualUNet3D", in_channels=1, out_channels=1, f_maps=16, num_levels=3))
    flat = FlatParameters(model)
    red = BucketedAllReduce(flat, world, n_buckets=3, min_bucket_bytes=1024)
    ok = True
    launched_early = 0
    for step in range(2):   # two steps: the bucket bookkeeping must reset
        g = torch.Generator().manual_seed(100 * step + rank)
        flat.grad.copy_(torch.randn(flat.numel, generator=g))
        # the engine reports finished parameter gradients in reverse parameter order; a bucket's allreduce starts as soon as it is full
        for k in reversed(flat.names):
            flat.written(k)
        launched_early += sum(red.launched)
        red.finish()
        exp = sum(torch.randn(flat.numel, generator=torch.Generator().manual_seed(100 * step + r)) for r in range(world))
        # padding elements between parameters are reduced too (contiguous bucket ranges): compare everything
        ok = ok and torch.allclose(flat.grad, exp, atol=1e-5)
    for k, p in model.named_parameters():
        ok = ok and p.grad.data_ptr() == flat.grad_views[k].data_ptr()
    q.put((rank, ok, launched_early, len(red.buckets)))
    dist.destroy_process_group()


def test_bucketed_allreduce_over_flat_gradients_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "bucketed allreduce result differs from the sum over ranks"
    assert all(r[3] == 3 and r[2] == 6 for r in res), res   # every bucket was launched by the notifications, before finish()


def test_gradient_allreduce_and_patch_sharding_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "allreduced gradients differ from the mean over ranks"
    assert all(r[2] for r in res), "replicas diverged"
    assert [r[3] for r in res] == [74, 73]
    assert res[0][4] == [0, 2, 4] and res[1][4] == [1, 3, 5]
