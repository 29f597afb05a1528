"""Device-resident sliding-window inference: the B200 counterpart of the reference's patch loader + StandardPredictor loop
(SURVEY.md section 8(f) rows f-1 and f-2).

Reference behaviour reproduced (file:line in the reference checkout):
  SliceBuilder patch grid, z outermost / x innermost      pytorch3dunet/datasets/utils.py:192-287
  halo: mirror-pad the raw volume, slice halo-extended     datasets/utils.py:518-546, datasets/hdf5.py:16-20,154-190
  per batch: input.to(device), model(input), crop halo,    pytorch3dunet/unet3d/predictor.py:148-193
  .cpu().numpy(), prediction_array[index] = pred            (later patches overwrite earlier ones)

What the reference does per PATCH on the host (numpy slice of the padded volume, pageable H2D copy, synchronous D2H copy, numpy
slice assignment) happens here per VOLUME and on the device:

  host volume --(Z-slabs: memcpy into a pinned ring, cudaMemcpyAsync on a COPY stream, one event per slab)--> (C,Z,Y,X) fp32 in HBM
  for every patch of this rank (patch i -> rank i mod world):   [compute stream; waits only for the slabs the patch touches]
      b200_patch_gather_f32   reflect-padded, halo-extended patch straight out of the resident volume (no padded copy)
      model(patch)            the engine, forward only (nothing taped under no_grad)
      b200_patch_scatter_f32  halo crop + write of the voxels this patch is the LAST writer of, into the resident output volume
  output volume --(Z-slabs finalised row by row of the patch grid, async D2H into pinned memory on the copy stream)--> host array

"Last writer wins" is evaluated analytically (`owner_tables`): the last patch covering a voxel is, per axis, the highest patch
index whose [start, stop) contains the coordinate, so every voxel is written exactly once, patches may run in any order and
on any GPU, and the per-rank output volumes are disjoint: the multi-GPU merge is one sum-reduce (NCCL) to rank 0.
"""
from __future__ import annotations

import threading

import numpy as np
import torch

from . import patches as PT
from ._lib import lib


def owner_tables(size, origins, patch):
    """owner[c] = index (into `origins`) of the LAST patch along this axis whose [o, o+patch) contains coordinate c; -1 if none."""
    own = np.full(int(size), -1, dtype=np.int32)
    for k, o in enumerate(origins):  # ascending k: later patches overwrite
        own[o:o + patch] = k
    return own


class SlabFeed:
    """Hand-over of uploaded Z-slabs from the staging thread to the compute loop: (z_hi, cuda event) pairs in ascending z."""

    def __init__(self):
        self.cv = threading.Condition()
        self.items, self.closed, self.error, self.used, self.thread = [], False, None, 0, None

    def push(self, z_hi, ev):
        with self.cv:
            self.items.append((z_hi, ev))
            self.cv.notify_all()

    def fail(self, e):
        with self.cv:
            self.error = e
            self.cv.notify_all()

    def close(self):
        with self.cv:
            self.closed = True
            self.cv.notify_all()

    def wait_until(self, need, stream):
        """make `stream` wait for every slab below plane `need` (blocks the host until those slabs have been ENQUEUED)"""
        while True:
            with self.cv:
                covered = self.items[self.used - 1][0] if self.used else 0
                if covered >= need:
                    return
                while self.used >= len(self.items) and not self.closed and self.error is None:
                    self.cv.wait()
                if self.error is not None:
                    raise self.error
                if self.used >= len(self.items):
                    return  # closed: everything there is has been waited for
                ev = self.items[self.used][1]
                self.used += 1
            stream.wait_event(ev)


class PatchPlan:
    """Patch grid of a (Z,Y,X) volume in the reference's order, with the per-axis last-writer tables."""

    def __init__(self, spatial, patch, stride, halo=(0, 0, 0)):
        self.spatial = tuple(int(v) for v in spatial)
        self.patch = tuple(int(v) for v in patch)
        self.stride = tuple(int(v) for v in stride)
        self.halo = tuple(int(v) for v in halo)
        self.origins = [PT.patch_origins(self.spatial[a], self.patch[a], self.stride[a]) for a in range(3)]
        self.grid = tuple(len(o) for o in self.origins)
        self.owners = [owner_tables(self.spatial[a], self.origins[a], self.patch[a]) for a in range(3)]
        for a in range(3):
            if self.halo[a] >= self.spatial[a]:
                raise ValueError("halo must be smaller than the volume (single reflection, np.pad(mode='reflect'))")

    def __len__(self):
        return self.grid[0] * self.grid[1] * self.grid[2]

    def item(self, i):
        """-> ((iz,iy,ix), (z0,y0,x0)) of patch i in the reference's enumeration order"""
        ny, nx = self.grid[1], self.grid[2]
        iz, r = divmod(i, ny * nx)
        iy, ix = divmod(r, nx)
        return (iz, iy, ix), (self.origins[0][iz], self.origins[1][iy], self.origins[2][ix])

    def slices(self):
        return PT.build_slices(self.spatial, self.patch, self.stride)

    def finalised_z(self, iz):
        """[z_lo, z_hi): output planes whose owner along z is patch-grid row iz (complete once every patch of rows <= iz ran)."""
        idx = np.nonzero(self.owners[0] == iz)[0]
        return (int(idx[0]), int(idx[-1]) + 1) if idx.size else (0, 0)


class VolumePredictor:
    """predict(volume) -> (C_out, Z, Y, X) float32 probabilities, identical in patch order / indices / overwrite semantics to the
    reference's StandardPredictor on an in-memory volume (`prediction_array`, predictor.py:113-193)."""

    def __init__(self, model, patch_shape, stride_shape, halo_shape=(0, 0, 0), device=None, world=1, rank=0, slab_planes=32,
                 process_group=None):
        self.model = model
        self.patch_shape, self.stride_shape, self.halo_shape = tuple(patch_shape), tuple(stride_shape), tuple(halo_shape)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("VolumePredictor runs on a CUDA device (the b200 engine has no CPU path)")
        self.world, self.rank, self.pg = int(world), int(rank), process_group
        self.slab_planes = int(slab_planes)
        self.L = lib()
        self.copy_stream = torch.cuda.Stream(self.device)
        self._ring = None
        self._host_out = None
        self.stats = {}

    # ------------------------------------------------------------------ staging (row f-1)
    def _pinned_ring(self, slab_elems, slots=3):
        if self._ring is None or self._ring[0].numel() < slab_elems or len(self._ring) < slots:
            self._ring = [torch.empty(slab_elems, dtype=torch.float32).pin_memory() for _ in range(slots)]
            self._ring_free = [None] * slots
        return self._ring

    def upload(self, vol, background=True):
        """host (C,Z,Y,X) fp32 -> device, Z-slab by Z-slab on the copy stream: host memcpy into a pinned ring slot, then
        cudaMemcpyAsync, double/triple buffered.  With `background` the staging loop runs on a helper thread so that the caller can
        start enqueueing patches as soon as the first slabs are on their way.  Returns (device volume, SlabFeed)."""
        C, Z, Y, X = vol.shape
        dvol = torch.empty((C, Z, Y, X), dtype=torch.float32, device=self.device)
        # the block may be a recycled one with work still queued on the compute stream; it is written on the copy stream
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        dvol.record_stream(self.copy_stream)
        src = torch.from_numpy(vol) if isinstance(vol, np.ndarray) else vol
        pinned_src = src.is_pinned()
        sp = max(1, min(self.slab_planes, Z))
        ring = None if pinned_src else self._pinned_ring(C * sp * Y * X)
        feed = SlabFeed()

        def run():
            try:
                with torch.cuda.device(self.device):
                    for k, z0 in enumerate(range(0, Z, sp)):
                        z1 = min(Z, z0 + sp)
                        with torch.cuda.stream(self.copy_stream):
                            if pinned_src:
                                dvol[:, z0:z1].copy_(src[:, z0:z1], non_blocking=True)
                            else:
                                slot = k % len(ring)
                                if self._ring_free[slot] is not None:
                                    self._ring_free[slot].synchronize()  # the slot's previous H2D copy must have drained
                                stage = ring[slot][: C * (z1 - z0) * Y * X].view(C, z1 - z0, Y, X)
                                stage.copy_(src[:, z0:z1])  # host memcpy into pinned memory (overlaps the previous slab's DMA)
                                dvol[:, z0:z1].copy_(stage, non_blocking=True)
                                ev_free = torch.cuda.Event()
                                ev_free.record(self.copy_stream)
                                self._ring_free[slot] = ev_free
                            ev = torch.cuda.Event()
                            ev.record(self.copy_stream)
                        feed.push(z1, ev)
            except BaseException as e:  # surfaced by the consumer
                feed.fail(e)
            finally:
                feed.close()

        if background and not pinned_src:
            feed.thread = threading.Thread(target=run, daemon=True)
            feed.thread.start()
        else:
            run()
        return dvol, feed

    # ------------------------------------------------------------------ the loop
    @torch.no_grad()
    def predict_device(self, dvol, upload_events=None, out=None, on_row_done=None):
        """Sliding window over a device-resident (C,Z,Y,X) volume -> device-resident (C_out,Z,Y,X) output (this rank's voxels; the
        rest stays zero).  `on_row_done(iz, z_lo, z_hi, event)` is called when the output planes [z_lo, z_hi) are final."""
        L = self.L
        C, Z, Y, X = dvol.shape
        plan = PatchPlan((Z, Y, X), self.patch_shape, self.stride_shape, self.halo_shape)
        hz, hy, hx = plan.halo
        pz, py, px = (plan.patch[a] + 2 * plan.halo[a] for a in range(3))
        own = [torch.from_numpy(o).to(self.device) for o in plan.owners]
        compute = torch.cuda.current_stream(self.device)
        st = compute.cuda_stream
        was_training = self.model.training
        self.model.eval()  # predictor.py:152
        n_done = 0
        try:
            for i in range(len(plan)):
                (iz, iy, ix), (z0, y0, x0) = plan.item(i)
                mine = (i % self.world) == self.rank
                if mine:
                    if upload_events is not None:
                        need = min(Z, z0 + plan.patch[0] + hz)  # deepest input plane this patch reads (reflection stays inside)
                        if z0 - hz < 0:
                            need = max(need, min(Z, hz - z0 + 1))
                        if z0 + plan.patch[0] + hz > Z:
                            need = Z
                        upload_events.wait_until(need, compute)
                    patch_t = torch.empty((1, C, pz, py, px), dtype=torch.float32, device=self.device)
                    L.call("b200_patch_gather_f32", dvol.data_ptr(), C, Z, Y, X, z0 - hz, y0 - hy, x0 - hx, pz, py, px, patch_t.data_ptr(), st)
                    pred = self.model(patch_t)
                    if pred.dtype != torch.float32 or not pred.is_contiguous():
                        pred = pred.float().contiguous()
                    if out is None:
                        out = torch.zeros((pred.shape[1], Z, Y, X), dtype=torch.float32, device=self.device)
                    L.call("b200_patch_scatter_f32", pred.data_ptr(), pred.shape[1], pz, py, px, hz, hy, hx, out.data_ptr(), Z, Y, X,
                           z0, y0, x0, iz, iy, ix, own[0].data_ptr(), own[1].data_ptr(), own[2].data_ptr(), st)
                    n_done += 1
                last_of_row = (iy == plan.grid[1] - 1) and (ix == plan.grid[2] - 1)
                if last_of_row and on_row_done is not None and out is not None:
                    ev = torch.cuda.Event()
                    ev.record(compute)
                    on_row_done(iz, *plan.finalised_z(iz), ev)
        finally:
            self.model.train(was_training)
        self.stats = dict(patches_total=len(plan), patches_this_rank=n_done, grid=plan.grid)
        return out

    def _pinned_out(self, shape):
        if self._host_out is None or tuple(self._host_out.shape) != tuple(shape):
            self._host_out = torch.empty(tuple(shape), dtype=torch.float32).pin_memory()
        return self._host_out

    def predict(self, volume, out_host=None):
        """End to end: host volume (Z,Y,X) or (C,Z,Y,X) fp32 -> host (C_out,Z,Y,X) fp32 (rank 0 holds the merged result when
        world > 1; the other ranks return their partial volume).  The returned array aliases a pinned buffer that the next call
        reuses -- copy it if it has to outlive the next predict()."""
        vol = np.asarray(volume, dtype=np.float32) if not torch.is_tensor(volume) else volume
        if vol.ndim == 3:
            vol = vol[None]
        with torch.cuda.device(self.device):
            dvol, events = self.upload(vol)
            if self.world > 1:
                out = self.predict_device(dvol, events)
                torch.distributed.reduce(out, dst=0, op=torch.distributed.ReduceOp.SUM, group=self.pg)  # disjoint shards: merge = sum
                host = self._pinned_out(out.shape) if out_host is None else out_host
                if self.rank == 0:
                    host.copy_(out, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                return host.numpy() if out_host is None else host
            # single GPU: stream finished rows of the patch grid back while the next rows compute
            holder = {}

            def row_done(iz, z_lo, z_hi, ev):
                if z_hi <= z_lo:
                    return
                if "host" not in holder:
                    o = holder["out"]
                    holder["host"] = self._pinned_out(o.shape) if out_host is None else out_host
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(ev)
                    holder["host"][:, z_lo:z_hi].copy_(holder["out"][:, z_lo:z_hi], non_blocking=True)

            C_out = getattr(self.model, "spec", {}).get("out_channels") if hasattr(self.model, "spec") else None
            if C_out is None:
                C_out = self.model(torch.zeros((1, vol.shape[0]) + tuple(p + 2 * h for p, h in zip(self.patch_shape, self.halo_shape)),
                                               device=self.device)).shape[1]
            holder["out"] = torch.zeros((C_out,) + tuple(vol.shape[1:]), dtype=torch.float32, device=self.device)
            self.predict_device(dvol, events, out=holder["out"], on_row_done=row_done)
            self.copy_stream.synchronize()
            torch.cuda.current_stream(self.device).synchronize()
            host = holder["host"]
            return host.numpy() if out_host is None else host
