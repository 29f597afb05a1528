"""CPU: the analytic last-writer-wins rule of the device write-back (pytorch3dunet_b200.pipeline.owner_tables / PatchPlan) against the
host restatement of the reference's predictor loop (patches.assemble_last_writer_wins: later patches overwrite earlier ones)."""
import numpy as np
import pytest


@pytest.mark.parametrize("shape,patch,stride,halo", [
    ((70, 90, 80), (32, 40, 32), (24, 25, 32), (0, 0, 0)),
    ((70, 90, 80), (32, 40, 32), (24, 25, 32), (4, 6, 8)),
    ((64, 64, 64), (32, 32, 32), (32, 32, 32), (0, 0, 0)),       # no overlap
    ((33, 65, 65), (16, 32, 32), (7, 11, 13), (2, 2, 2)),        # ragged: the flush-with-the-end last patch overlaps by more than patch - stride
    ((40, 40, 40), (40, 40, 40), (20, 20, 20), (3, 3, 3)),       # a single patch
])
def test_owner_rule_equals_sequential_overwrite(shape, patch, stride, halo):
    from pytorch3dunet_b200 import patches as PT
    from pytorch3dunet_b200.pipeline import PatchPlan
    plan = PatchPlan(shape, patch, stride, halo)
    idx = PT.build_slices(shape, patch, stride)
    assert len(plan) == len(idx)
    rng = np.random.default_rng(0)
    # every patch predicts distinct values (with its halo attached, as the model would return them)
    full = tuple(p + 2 * h for p, h in zip(patch, halo))
    preds = [rng.standard_normal((2,) + full).astype(np.float32) for _ in idx]
    want = PT.assemble_last_writer_wins(preds, idx, (2,) + tuple(shape), halo)
    # device rule, emulated: patch i writes voxel v iff owner_z[z]==iz and owner_y[y]==iy and owner_x[x]==ix -- in ANY patch order
    got = np.full((2,) + tuple(shape), np.nan, dtype=np.float32)
    writes = np.zeros(shape, dtype=np.int32)
    for i in reversed(range(len(plan))):
        (iz, iy, ix), (z0, y0, x0) = plan.item(i)
        assert (z0, y0, x0) == tuple(s.start for s in idx[i])
        mz = plan.owners[0][z0:z0 + patch[0]] == iz
        my = plan.owners[1][y0:y0 + patch[1]] == iy
        mx = plan.owners[2][x0:x0 + patch[2]] == ix
        m = mz[:, None, None] & my[None, :, None] & mx[None, None, :]
        core = preds[i][:, halo[0]:halo[0] + patch[0], halo[1]:halo[1] + patch[1], halo[2]:halo[2] + patch[2]]
        region = got[:, z0:z0 + patch[0], y0:y0 + patch[1], x0:x0 + patch[2]]
        region[:, m] = core[:, m]
        writes[z0:z0 + patch[0], y0:y0 + patch[1], x0:x0 + patch[2]] += m
    assert np.all(writes == 1)            # every voxel is written exactly once
    assert np.array_equal(got, want)      # and holds what sequential overwriting leaves behind
    # rows of the patch grid finalise disjoint z-ranges that tile the volume
    zr = [plan.finalised_z(k) for k in range(plan.grid[0])]
    assert zr[0][0] == 0 and zr[-1][1] == shape[0] and all(a[1] == b[0] for a, b in zip(zr, zr[1:]))


def test_slab_feed_orders_waits():
    import threading
    from pytorch3dunet_b200.pipeline import SlabFeed

    class FakeStream:
        def __init__(self):
            self.waited = []

        def wait_event(self, ev):
            self.waited.append(ev)

    feed, st = SlabFeed(), FakeStream()
    t = threading.Thread(target=lambda: [feed.push(z, f"ev{z}") for z in (32, 64, 96, 128)] + [feed.close()])
    t.start()
    feed.wait_until(40, st)
    assert st.waited == ["ev32", "ev64"]
    feed.wait_until(64, st)
    assert st.waited == ["ev32", "ev64"]
    feed.wait_until(128, st)
    t.join()
    assert st.waited == ["ev32", "ev64", "ev96", "ev128"]
    feed.wait_until(10 ** 9, st)  # closed: returns
