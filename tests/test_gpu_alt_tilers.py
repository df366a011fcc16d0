"""-m gpu: the alternative tilers through the C ABI against oracle/alt_tilers.py + oracle/pil_resample.py
(bit-exact vs the installed Pillow): the x4 patch shrink and the averaging stitch are integer / fixed-order arithmetic
and must match bit for bit; the end-to-end methods go through the bf16 conv stack and are compared as uint8 images."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,p,step,stretch", [(40, 36, 32, 4, True), (24, 20, 16, 1, False), (132, 136, 128, 4, True),
                                                (16, 12, 8, 4, True), (9, 11, 4, 1, False)])
def test_patch_down4_is_scipy_imresize_bicubic(H, W, p, step, stretch):
    from oracle import pil_resample as pr
    from sr100 import alt_tilers as at
    rng = np.random.default_rng(H * 7 + p)
    img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    img[: H // 2, : W // 2] //= 4                        # patches with a narrow value range: bytescale matters
    img[H // 2:, W // 2:] = 77                           # flat patches: cmax == cmin
    got = at.patch_down4(torch.from_numpy(img).cuda(), p, step, stretch).cpu().numpy()
    cnt_h, cnt_w = (H - p) // step + 1, (W - p) // step + 1
    assert got.shape == (cnt_h * cnt_w, p // 4, p // 4, 3)
    q = p // 4
    for a in range(cnt_h):
        for b in range(cnt_w):
            patch = img[a * step:a * step + p, b * step:b * step + p]
            want = pr.imresize_bicubic(patch.astype(np.float64) if stretch else patch, (q, q))
            assert np.array_equal(got[a * cnt_w + b], want.astype(np.float32) / np.float32(255.0)), (a, b)


@pytest.mark.parametrize("H,W,P,step,pad,rows", [(24, 20, 16, 4, 4, 99), (24, 20, 16, 4, 4, 1), (14, 17, 8, 1, 0, 3),
                                                 (12, 12, 12, 4, 4, 99), (24, 20, 16, 1, 0, 99), (40, 33, 8, 1, 0, 7)])
def test_patch_average_matches_reference_loops_bit_for_bit(H, W, P, step, pad, rows):
    from itertools import product
    from oracle import alt_tilers as oat
    from sr100 import alt_tilers as at
    rng = np.random.default_rng(H + W + P)
    cnt_h, cnt_w = (H - P) // step + 1, (W - P) // step + 1
    patches = (rng.random((cnt_h * cnt_w, P, P, 3)) * 1.2 - 0.1).astype(np.float32)      # some values clip
    pd = torch.from_numpy(patches).cuda()
    chunks = [(a0, min(cnt_h, a0 + rows), pd[a0 * cnt_w:min(cnt_h, a0 + rows) * cnt_w].contiguous())
              for a0 in range(0, cnt_h, rows)]
    u8, f64 = at.patch_average(chunks, P, step, pad, cnt_h, cnt_w, (H, W), mul=255.0, want_f64=True,
                               sklearn_count=(pad == 0))
    scaled = patches.astype(np.float32) * 255.
    if pad:
        want = oat.reconstruct_from_patches_2dlocal((H - P + 1, W - P + 1), (P, P), scaled, (H, W, 3), step)
    else:
        want = np.zeros((H, W, 3))
        for pch, (i, j) in zip(scaled, product(range(H - P + 1), range(W - P + 1))):
            want[i:i + P, j:j + P] += pch
        for i in range(H):
            for j in range(W):
                want[i, j] /= float(min(i + 1, P, H - i) * min(j + 1, P, W - j))
    assert np.array_equal(f64.cpu().numpy(), want)                   # same float64 additions in the same order
    assert np.array_equal(u8.cpu().numpy(), np.clip(want, 0, 255).astype('uint8'))


@pytest.fixture(scope="module")
def sr(tmp_path_factory):
    import os
    import models
    from oracle import model as om
    weights = om.init_weights(1234, bias_scale=0.01)
    w, b = weights["conv2d_85"]
    weights["conv2d_85"] = (w * 8.0, b + 0.3)
    m = models.DifvdsrDouble(1)
    model = m.create_model(4, 4)
    model.engine.set_weights_dict(weights)
    wfile = str(tmp_path_factory.mktemp("w") / "w.h5")
    model.save_weights(wfile)
    os.environ["SR100_WEIGHTS"] = wfile
    m._loaded_from = None
    yield m, weights
    os.environ.pop("SR100_WEIGHTS", None)


def _smooth(rng, h, w):
    from scipy.ndimage import uniform_filter
    return uniform_filter(rng.integers(0, 256, size=(h, w, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)


def test_upscale_patch_end_to_end(sr, tmp_path):
    from PIL import Image
    from oracle import alt_tilers as oat
    from oracle import model as om
    m, weights = sr
    rng = np.random.default_rng(2)
    img = _smooth(rng, 21, 18)                                       # not a multiple of 4: padded to 24 x 20
    path = str(tmp_path / "e.png")
    Image.fromarray(img).save(path)
    got = m.upscalePatch(path, return_image=True, patch_size=16, verbose=False)
    want = oat.upscale_patch(img, lambda x: om.forward_numpy(weights, x), patch_size=16)
    assert got.shape == want.shape == (21, 18, 3) and got.dtype == np.uint8
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 2 and (d > 0).mean() < 0.25, (d.max(), (d > 0).mean())
    m.upscalePatch(path, patch_size=16, verbose=False, suffix="enh")
    saved = np.asarray(Image.open(str(tmp_path / "e_enh(1x).png")))
    assert np.array_equal(saved, got)
    with pytest.raises(ValueError):
        m.upscalePatch(path, return_image=True, patch_size=32, verbose=False)      # patch larger than the image
    with pytest.raises(UnboundLocalError):
        m.upscalePatch(path, mode="fast", verbose=False)                           # the reference's own failure


def test_upscale_patch_mode_end_to_end(sr, tmp_path):
    from PIL import Image
    from oracle import alt_tilers as oat
    from oracle import model as om
    m, weights = sr
    rng = np.random.default_rng(3)
    img = _smooth(rng, 6, 5)
    path = str(tmp_path / "u.png")
    Image.fromarray(img).save(path)
    got = m.upscale(path, return_image=True, patch_size=16, mode="patch", verbose=False)
    big, want = oat.upscale_patch_mode(img, lambda x: om.forward_numpy(weights, x), patch_size=16)
    assert got.shape == want.shape == (24, 20, 3)
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "u_Ascaled(1x).png"))), big)   # the side file (:657)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 2 and (d > 0).mean() < 0.25, (d.max(), (d > 0).mean())


def test_img_utils_reconstructions_match_reference_goldens(golden_dir):
    """img_utils.reconstruct_from_patches_2dlocal / combine_patches / reconstruct_from_patches_2dloc on the device vs
    the outputs of the reference's own functions (tests/golden/alt_tilers_ref.npz, oracle/refgen_alt.py): float64,
    bit for bit, NaN where the reference has NaN."""
    import os
    import img_utils
    g = np.load(os.path.join(golden_dir, "alt_tilers_ref.npz"))
    img = g["img"]
    full = img_utils.make_patchesOrig(img.astype(np.float64), 1, 8)
    got = img_utils.reconstruct_from_patches_2dlocal(full, g["cnn"], img.shape, step=4)
    assert np.array_equal(got, g["rec_local"], equal_nan=True)
    for tag in ("b", "c"):
        h, w, p, n = g["%s_meta" % tag]
        dense = np.zeros(((h - p + 1) * (w - p + 1), p, p, 3), dtype=np.uint8)
        got = img_utils.reconstruct_from_patches_2dlocal(dense, g["%s_cnn" % tag], (h, w, 3), step=4)
        assert np.array_equal(got, g["%s_rec" % tag], equal_nan=True)
    vals = np.random.default_rng(77).normal(120, 60, size=tuple(g["dense_shape"])).astype(np.float32)
    assert np.array_equal(img_utils.combine_patches(vals, img.shape, 1), g["combine"])
    assert np.array_equal(img_utils.reconstruct_from_patches_2dloc(vals, img.shape), g["rec_loc"])
    dv = np.random.default_rng(78).normal(120, 60, size=(49, 8, 8, 3)).astype(np.float32)
    assert np.array_equal(img_utils.combine_patches(dv, tuple(g["small_shape"]), 1), g["small_combine"])
    with pytest.raises(ValueError):
        img_utils.combine_patches(vals.astype(np.float64) + 1e-9, img.shape, 1)     # not float32-representable
