"""-m gpu: patch gather / stitch kernels and the upscaleStepPatch pipeline.  Bit-exact bar: the outputs of the
reference's own functions (tests/golden/tiling_ref.npz) and the numpy oracle on seeded random shapes."""
import os

import numpy as np
import pytest
import torch

from oracle import tiling as ot

pytestmark = pytest.mark.gpu


def test_gather_and_stitch_match_reference_golden(golden_dir):
    import img_utils
    z = np.load(golden_dir + "/tiling_ref.npz")
    for ci in range(5):
        ch, cw, p, st, sc, cnt_h, cnt_w = [int(v) for v in z["c%d_meta" % ci]]
        crng = np.random.default_rng(1000 + ci)
        canvas = crng.integers(0, 256, size=(ch, cw, 3)).astype(np.float64)
        patches, counts = img_utils.extract_patches_Step(canvas, (p, p), st)
        assert patches.dtype == np.float64 and counts == (cnt_h, cnt_w)
        assert np.array_equal(patches.sum(axis=(1, 2, 3)), z["c%d_patches_sum" % ci])
        assert np.array_equal(patches[0].astype(np.uint8), z["c%d_patches_first" % ci])
        assert np.array_equal(patches[-1].astype(np.uint8), z["c%d_patches_last" % ci])
        up = crng.integers(-20, 281, size=(patches.shape[0], p * sc, p * sc, 3)).astype(np.float32)
        rebuilt = img_utils.rebuild_from_patches_Step(canvas, up, (p, p), counts, sc, st)
        assert rebuilt.dtype == np.float64 and rebuilt.shape == (ch * sc, cw * sc, 3)
        assert np.array_equal(rebuilt.astype(np.int16), z["c%d_rebuilt" % ci])


@pytest.mark.parametrize("seed", range(6))
def test_gather_stitch_random_shapes_vs_oracle(seed):
    import img_utils
    rng = np.random.default_rng(seed)
    p = int(rng.integers(5, 20))
    st = int(rng.integers(2, p + 1))
    sc = int(rng.choice([1, 2, 4]))
    ch, cw = int(rng.integers(p + 1, 70)), int(rng.integers(p + 1, 70))
    canvas = rng.integers(0, 256, size=(ch, cw, 3)).astype(np.float64)
    got, counts = img_utils.extract_patches_Step(canvas, (p, p), st)
    want, wcounts = ot.extract_patches_step(canvas, (p, p), st)
    assert counts == wcounts and np.array_equal(got, want)
    up = rng.random((got.shape[0], p * sc, p * sc, 3)).astype(np.float32)
    if p * sc > 16:   # the 8-px crop needs patches wider than 16 HR px to be meaningful
        r = img_utils.rebuild_from_patches_Step(canvas, up, (p, p), counts, sc, st)
        w = ot.rebuild_from_patches_step((ch, cw), up, (p, p), counts, sc, st)
        assert np.array_equal(r, w)


def test_extract_raises_like_reference():
    import img_utils
    with pytest.raises(ValueError, match="Height of the patch"):
        img_utils.extract_patches_Step(np.zeros((8, 30, 3)), (12, 12), 8)
    with pytest.raises(ValueError, match="Width of the patch"):
        img_utils.extract_patches_Step(np.zeros((30, 8, 3)), (12, 12), 8)


def test_fused_u8_gather_equals_canvas_path():
    from sr100 import ops
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, size=(339, 510, 3)).astype(np.uint8)        # DIV2K-shaped LR image (config 3)
    ch, cw = ops.canvas_size(339, 510, 96, 64)
    assert (ch, cw) == (448, 640)
    got, counts = ops.patch_gather_u8(torch.from_numpy(img).cuda(), (ch, cw), (96, 96), 64, divisor=255.0)
    assert counts == (6, 9) and got.shape == (54, 96, 96, 3)
    want, _ = ot.extract_patches_step(ot.make_canvas(img, 96, 64), (96, 96), 64)
    want = want.astype(np.float32) / 255.
    assert np.array_equal(got.cpu().numpy(), want)                          # same fp32 division, bit-exact


def test_identity_network_property_at_full_size():
    """Config 5 shape (1080x1920 -> 558 tiles): gather -> nearest-neighbour x4 per tile -> stitch reproduces the
    nearest-neighbour x4 of the canvas exactly; uncovered pixels stay 0."""
    from sr100 import ops
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, size=(1080, 1920, 3)).astype(np.uint8)
    ch, cw = ops.canvas_size(1080, 1920, 96, 64)
    assert (ch, cw) == (1216, 2048)
    patches, counts = ops.patch_gather_u8(torch.from_numpy(img).cuda(), (ch, cw), (96, 96), 64, divisor=1.0)
    assert counts == (18, 31)
    up = patches.repeat_interleave(4, dim=1).repeat_interleave(4, dim=2).contiguous()
    _, out = ops.patch_stitch(up, counts, (96, 96), 64, 4, (ch, cw), mul=1.0, want_f32=False, want_u8=True)
    out = out.cpu().numpy()
    want = np.repeat(np.repeat(img, 4, axis=0), 4, axis=1)
    assert np.array_equal(out[:4320, :7680], want)
    # last patch row/col ends at 64*17+96 = 1184 < 1216: the canvas tail is never written
    assert out[1184 * 4:, :, :].max() == 0 and out[:, (64 * 30 + 96) * 4:, :].max() == 0


def test_upscale_step_patch_end_to_end(tmp_path):
    """models.DifvdsrDouble.upscaleStepPatch on a file vs the oracle pipeline with the oracle network."""
    from PIL import Image
    import models
    from oracle import model as om
    rng = np.random.default_rng(21)
    from scipy.ndimage import uniform_filter
    img = uniform_filter(rng.integers(0, 256, size=(40, 30, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)
    path = str(tmp_path / "img.bmp")
    Image.fromarray(img).save(path)
    weights = om.init_weights(99, bias_scale=0.02)
    # random-init outputs are ~0.05; scale the tail so the uint8 output is not all zeros
    w, b = weights["conv2d_85"]
    weights["conv2d_85"] = (w * 8.0, b + 0.3)
    wfile = str(tmp_path / "weights.npz")
    np.savez(wfile, **{k + "/kernel:0": v[0] for k, v in weights.items()},
             **{k + "/bias:0": v[1] for k, v in weights.items()})
    os.environ["SR100_WEIGHTS"] = wfile
    try:
        m = models.DifvdsrDouble(1)
        m.upscaleStepPatch(path, scalemulti=4, patch_size=32, suffix="scaled", verbose=False)
        full = m.upscaleStepPatch(path, return_image=True, scalemulti=4, patch_size=32, verbose=False)
    finally:
        del os.environ["SR100_WEIGHTS"]
    out_path = str(tmp_path / "img_scaled(1x).bmp")
    assert os.path.exists(out_path)
    got = np.asarray(Image.open(out_path))
    want_full, want = ot.upscale_step_patch(img, lambda x: om.forward_numpy(weights, x), 32, 64, 4)
    assert got.shape == want.shape == (160, 120, 3)
    assert full.shape == want_full.shape
    assert got.max() > 30                                       # a real image, not zeros
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1                                          # truncation may flip one LSB near integers
    assert (d > 0).mean() < 0.2
    assert np.abs(full.astype(int) - want_full.astype(int)).max() <= 1


@pytest.mark.parametrize("shape,patch,step,divisor", [((5, 70, 45), 96, 64, 255.0), ((3, 339, 510), 96, 64, 255.0),
                                                      ((2, 50, 61), 32, 16, 1.0), ((2, 40, 40), 17, 8, 255.0),
                                                      ((2, 50, 61), 32, 16, 100.0), ((3, 33, 35), 32, 16, 255.0)])
def test_batched_gather_equals_per_image_gather_and_oracle(shape, patch, step, divisor):
    """One launch for a batch of same-shaped images (BASELINE config 3) == the per-image gathers, which are the
    reference's extract_patches_Step on the zero-padded canvas (bit-exact; 17-px patches take the scalar kernel)."""
    import torch
    from oracle import tiling as ot
    from sr100 import ops
    rng = np.random.default_rng(shape[1])
    imgs = rng.integers(0, 256, size=shape + (3,)).astype(np.uint8)
    ch, cw = ops.canvas_size(shape[1], shape[2], patch, step)
    dev = torch.from_numpy(imgs).cuda()
    got, counts = ops.patch_gather_u8_batched(dev, (ch, cw), (patch, patch), step, divisor=divisor)
    n = counts[0] * counts[1]
    assert got.shape == (shape[0] * n, patch, patch, 3)
    for m in range(shape[0]):
        one, c1 = ops.patch_gather_u8(dev[m], (ch, cw), (patch, patch), step, divisor=divisor)
        assert c1 == counts and torch.equal(one, got[m * n:(m + 1) * n])
        canvas = np.zeros((ch, cw, 3), dtype=np.float64)
        canvas[:shape[1], :shape[2]] = imgs[m]
        want, wc = ot.extract_patches_step(canvas, (patch, patch), step)
        assert tuple(wc) == tuple(counts)
        assert np.array_equal(got[m * n:(m + 1) * n].cpu().numpy(), (want.astype(np.float32) / np.float32(divisor)))


def test_gather_divides_every_byte_value_like_numpy():
    """The /255 of models.py:336 for all 256 byte values (the kernel's arithmetic path: one product + two FMAs) and
    for an image view that starts at an odd address (the word loads + funnel shift), bit for bit against numpy's
    correctly rounded float32 division."""
    import torch
    from sr100 import ops
    base = torch.arange(3 * 40 * 44 + 7, dtype=torch.int64).remainder(256).to(torch.uint8).cuda()
    for off in (0, 1, 2, 3, 5):
        img = base[off:off + 3 * 40 * 44].view(40, 44, 3)
        ch, cw = ops.canvas_size(40, 44, 32, 16)
        got, counts = ops.patch_gather_u8(img, (ch, cw), (32, 32), 16, divisor=255.0)
        host = img.cpu().numpy()
        canvas = np.zeros((ch, cw, 3), dtype=np.uint8)
        canvas[:40, :44] = host
        n = 0
        g = got.cpu().numpy()
        assert set(np.unique(host)) == set(range(256))
        for wi in range(counts[1]):
            for hi in range(counts[0]):
                want = canvas[hi * 16:hi * 16 + 32, wi * 16:wi * 16 + 32].astype(np.float32) / np.float32(255)
                assert np.array_equal(g[n], want), (off, n)
                n += 1
