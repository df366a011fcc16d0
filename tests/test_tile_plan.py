"""CPU: the tiled path's dead-work elimination (sr100.engine.plan_tiles / hr_extent) against the ORACLE's stitch.

The oracle (oracle/tiling.py, pinned to the reference's own rebuild_from_patches_Step outputs) tells, for every pixel
of the final 4H x 4W image, which tile and which patch-local coordinate the reference's overwrite order leaves
there.  The plan is valid iff every such pixel comes from a tile the plan runs and lies at least 7 pixels (the HR
stage's receptive-field radius: 2 blocks x (1+2) + the 3x3 tail) inside the HR extent the plan computes for it."""
import numpy as np
import pytest

from oracle import tiling as ot


def _ownership(h, w, patch=96, step=64, scale=4):
    ch, cw = ot.canvas_size(h, w, patch, step)
    canvas = np.zeros((ch, cw, 3))
    _, counts = ot.extract_patches_step(canvas, (patch, patch), step)
    n = counts[0] * counts[1]
    P = patch * scale
    patches = np.zeros((n, P, P, 3))
    patches[..., 0] = np.arange(n)[:, None, None] + 1          # tile id + 1 (0 = never written)
    patches[..., 1] = np.arange(P)[None, :, None]              # local y
    patches[..., 2] = np.arange(P)[None, None, :]              # local x
    full = ot.rebuild_from_patches_step((ch, cw), patches, (patch, patch), counts, scale, step)
    return full[:scale * h, :scale * w], counts


@pytest.mark.parametrize("hw", [(512, 512), (288, 288), (256, 256), (280, 280), (344, 228), (20, 20), (33, 50),
                                (64, 64), (65, 63), (96, 200), (339, 510), (130, 1)])
def test_plan_covers_every_final_pixel(hw):
    from sr100.engine import plan_tiles
    h, w = hw
    own, full_counts = _ownership(h, w)
    (gh, gw), (lh, lw), ext = plan_tiles(h, w)
    assert lh <= full_counts[0] and lw <= full_counts[1] and len(ext) == lh * lw
    tid = own[..., 0].astype(int) - 1
    assert tid.min() >= 0                                       # every final pixel is written by some tile
    wi, hi = tid // full_counts[0], tid % full_counts[0]        # reference order n = wi*cnt_h + hi
    assert hi.max() < lh and wi.max() < lw                      # ... by a tile the plan runs
    eh = np.array([e[0] for e in ext]).reshape(lw, lh)[wi, hi]  # plan order n' = wi*lh + hi
    ew = np.array([e[1] for e in ext]).reshape(lw, lh)[wi, hi]
    assert np.all(own[..., 1] + 7 <= eh - 1) and np.all(own[..., 2] + 7 <= ew - 1)
    for e in ext:
        assert e[0] % 4 == 0 and e[1] % 4 == 0 and 4 <= e[0] <= 384 and 4 <= e[1] <= 384
    # the gather's virtual canvas yields exactly the live counts
    assert ot.extract_patches_step(np.zeros((gh, gw, 3)), (96, 96), 64)[1] == (lh, lw)


def test_plan_full_canvas_is_the_reference_grid():
    from sr100.engine import plan_tiles
    for h, w in [(512, 512), (344, 228), (339, 510), (1080, 1920)]:
        (gh, gw), counts, ext = plan_tiles(h, w, full_canvas=True)
        assert (gh, gw) == ot.canvas_size(h, w) and ext is None
        assert counts == ot.extract_patches_step(np.zeros((gh, gw, 3)), (96, 96), 64)[1]
    # SURVEY 8d tiling geometry
    assert plan_tiles(1080, 1920, full_canvas=True)[1] == (18, 31)
    assert plan_tiles(1080, 1920)[1] == (17, 30)
    assert sum(np.prod(plan_tiles(h, w)[1]) for h, w in [(512, 512), (288, 288), (256, 256), (280, 280), (344, 228)]) == 154


def _rasterise(rows, out_h, out_w, P):
    """What the fused tail epilogue does with a stitch-tile table: patch n writes (n + 1, local y, local x) into
    every pixel it owns that falls inside the image."""
    img = np.zeros((out_h, out_w, 3))
    cover = np.zeros((out_h, out_w), dtype=int)
    for n, t in enumerate(rows):
        ys = np.arange(t["oy0"], t["oy1"])
        xs = np.arange(t["ox0"], t["ox1"])
        Y, X = t["y0"] + ys, t["x0"] + xs
        ys, Y = ys[(Y >= 0) & (Y < out_h)], Y[(Y >= 0) & (Y < out_h)]
        xs, X = xs[(X >= 0) & (X < out_w)], X[(X >= 0) & (X < out_w)]
        img[np.ix_(Y, X)] = np.stack(np.broadcast_arrays(float(n + 1), ys[:, None], xs[None, :]), axis=-1)
        cover[np.ix_(Y, X)] += 1
    return img, cover


@pytest.mark.parametrize("hw,patch", [((512, 512), 96), ((344, 228), 96), ((20, 20), 96), ((65, 63), 96),
                                      ((339, 510), 96), ((130, 1), 96), ((100, 30), 32), ((70, 45), 32), ((40, 100), 32)])
def test_stitch_tile_table_is_the_reference_overwrite_order(hw, patch):
    """Engine.stitch_tiles_host (the ownership rectangles the fused tail-conv stitch writes) against the oracle's
    rebuild_from_patches_step (pinned to the reference's own function): same tile and same patch-local coordinate in
    every pixel of the uncropped canvas (full grid) and of the final image (live tiles only), every pixel written at
    most once."""
    from sr100.engine import Engine, plan_tiles
    h, w = hw
    scale, step, P = 4, 64, patch * 4
    ch, cw = ot.canvas_size(h, w, patch, step)
    _, counts = ot.extract_patches_step(np.zeros((ch, cw, 3)), (patch, patch), step)
    n = counts[0] * counts[1]
    patches = np.zeros((n, P, P, 3))
    patches[..., 0] = np.arange(n)[:, None, None] + 1
    patches[..., 1] = np.arange(P)[None, :, None]
    patches[..., 2] = np.arange(P)[None, None, :]
    want = ot.rebuild_from_patches_step((ch, cw), patches, (patch, patch), counts, scale, step)
    rows, offs, nbytes = Engine.stitch_tiles_host([(scale * ch, scale * cw, counts, 0)], patch, step, scale)
    got, cover = _rasterise(rows, scale * ch, scale * cw, P)
    assert cover.max() <= 1 and np.array_equal(got, want)
    # live tiles only, cropped to the final image
    (gh, gw), (lh, lw), ext = plan_tiles(h, w, patch, step, scale)
    rows, _, _ = Engine.stitch_tiles_host([(scale * h, scale * w, (lh, lw), 0)], patch, step, scale)
    got, cover = _rasterise(rows, scale * h, scale * w, P)
    final = want[:scale * h, :scale * w]
    tid = final[..., 0].astype(int) - 1
    live_id = np.where(tid >= 0, (tid // counts[0]) * lh + tid % counts[0] + 1, 0)     # renumber into the live grid
    assert cover.max() <= 1
    assert np.array_equal(got[..., 0], live_id) and np.array_equal(got[..., 1:] * (live_id > 0)[..., None],
                                                                   final[..., 1:] * (live_id > 0)[..., None])
    # a column strip of a sharded image: tiles [lo, hi) shifted by -x0
    lo, hi = lh * (lw // 2), lh * lw
    if hi > lo:
        from sr100 import ops
        x0, x1 = ops.shard_strip((lh, lw), (patch, patch), step, scale, scale * w, lo, hi)
        rows, _, _ = Engine.stitch_tiles_host([(scale * h, x1 - x0, (lh, lw), -x0)], patch, step, scale)
        got, cover = _rasterise(rows[lo:hi], scale * h, x1 - x0, P)
        sel = (live_id[:, x0:x1] > lo) & (live_id[:, x0:x1] <= hi)
        assert np.array_equal(got[..., 0] > 0, sel)
        assert np.array_equal(got[..., 0][sel] + lo, live_id[:, x0:x1][sel])
