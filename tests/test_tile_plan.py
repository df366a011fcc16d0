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
