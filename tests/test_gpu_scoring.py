"""-m gpu: scoring kernels (crop + RGB->Y + PSNR + SSIM) and the PSNR.py / scorpath.py mirrors vs the oracle and the
reference's own PSNR.py outputs.  Tolerances from SURVEY 8(d): PSNR +-0.01 dB, SSIM +-1e-4 (we assert far tighter)."""
import numpy as np
import pytest
import torch

from oracle import scoring as osc

pytestmark = pytest.mark.gpu


def _pair(rng, h, w, sigma=3.0):
    a = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    b = np.clip(a.astype(np.float64) + rng.normal(0, sigma, size=a.shape), 0, 255).astype(np.uint8)
    return a, b


@pytest.mark.parametrize("shape", [(64, 64), (57, 91), (27, 27), (512, 512), (228, 344)])
def test_score_pair_matches_oracle(shape):
    import scorpath
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a, b = _pair(rng, *shape)
    psnr, ssim_rgb, ssim_y = scorpath.score_pair(a, b, 10)
    w_psnr, w_rgb, w_y = osc.score_pair(a, b, 10)
    assert abs(psnr - w_psnr) < 1e-7
    assert abs(ssim_rgb - w_rgb) < 1e-9
    assert abs(ssim_y - w_y) < 1e-9


def test_score_identical_and_errors():
    from sr100 import ops
    from sr100 import _lib as L
    a = np.random.default_rng(0).integers(0, 256, size=(40, 40, 3)).astype(np.uint8)
    r = ops.score_pair(torch.from_numpy(a).cuda(), torch.from_numpy(a).cuda(), 10)
    assert r["ssim_y"] == pytest.approx(1.0, abs=1e-12) and r["psnr_y"] == float("inf")
    with pytest.raises(L.SrError):
        ops.score_pair(torch.from_numpy(a[:24, :24].copy()).cuda(), torch.from_numpy(a[:24, :24].copy()).cuda(), 10)
    with pytest.raises(ValueError):
        ops.score_pair(torch.from_numpy(a).cuda(), torch.from_numpy(a[:30].copy()).cuda(), 10)


def test_batch_scale_checksum():
    """DIV2K-shaped pair (config 3 output size 1356x2040): blocks partition the image exactly once:
    sum of squared Y error from the kernel == numpy, n_pix/n_win closed form."""
    from sr100 import ops
    rng = np.random.default_rng(8)
    a, b = _pair(rng, 1356, 2040)
    r = ops.score_pair(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 10)
    ya, yb = osc.rgb2ycbcr_y(a[10:-10, 10:-10]), osc.rgb2ycbcr_y(b[10:-10, 10:-10])
    assert r["n_pix"] == 1336 * 2020 and r["n_win"] == 1330 * 2014
    assert r["sum_sq_y"] == pytest.approx(float(np.sum((ya - yb) ** 2)), rel=1e-10)


def test_rgb2y_and_psnr_mirrors_match_reference_golden(golden_dir):
    import PSNR
    import scorpath
    z = np.load(golden_dir + "/psnr_ref.npz")
    a, b, ya, yb = z["a"], z["b"], z["ya"], z["yb"]
    assert np.abs(scorpath.setimgrgb2ycbcr(a) - ya).max() < 1e-10
    assert PSNR.psnrNITRE(yb, ya, 0) == pytest.approx(float(z["psnrNITRE_y"]), abs=1e-9)
    assert PSNR.psnrNITRE(yb, ya, 4) == pytest.approx(float(z["psnrNITRE_y_shave4"]), abs=1e-9)
    assert PSNR.PSNRTorch(yb, ya, 0) == pytest.approx(float(z["PSNRTorch_y"]), abs=1e-9)
    assert PSNR.PSNRTorch(ya, ya, 0) == 100
    assert PSNR.psnrVDSR(yb, ya, 2) == pytest.approx(float(z["psnrVDSR_y_2"]), abs=1e-9)
    assert PSNR.psnrSVLAB(a, b) == pytest.approx(float(z["psnrSVLAB_u8"]), abs=1e-9)
    # uint8 inputs: the reference's own-dtype (wrap-around) differences, PSNR.py:14,28-29
    assert PSNR.psnrVDSR(b, a, 2) == pytest.approx(float(z["psnrVDSR_u8_2"]), abs=1e-9)
    assert PSNR.PSNRTorch(b, a, 0) == pytest.approx(float(z["PSNRTorch_u8"]), abs=1e-9)
    assert PSNR.psnrNITRE(b, a, 0) == pytest.approx(float(z["psnrNITRE_u8"]), abs=1e-9)
    assert np.array_equal(PSNR.im2double(a)[:2, :3], z["im2double_a"])
    assert np.array_equal(PSNR.im2doubleZ(a)[:2, :3], z["im2doubleZ_a"])
    import models
    assert models.psnr2(a.astype(np.float64), b.astype(np.float64)) == pytest.approx(
        osc.psnr_torch(a, b, 0), abs=1e-9)


def test_cv2_colour_mirrors_match_reference_golden(golden_dir):
    import scorpath
    z = np.load(golden_dir + "/cv2_colour_ref.npz")
    assert np.array_equal(scorpath.rgb2ycbcrCV(z["a"]), z["ycbcr"])
    assert np.array_equal(scorpath.ycbcr2rgb(z["ycbcr"].copy()), z["rgb_back"])
    assert scorpath.crop_border(np.zeros((30, 40, 3)), 10).shape == (10, 20, 3)


def test_scorpath_main_on_directory(tmp_path, capsys):
    from PIL import Image
    import scorpath
    rng = np.random.default_rng(2)
    want = []
    for name, shape in (("bird_GT", (72, 72)), ("woman_GT", (57, 86))):
        a, b = _pair(rng, *shape)
        Image.fromarray(a).save(str(tmp_path / (name + ".bmp")))
        Image.fromarray(b).save(str(tmp_path / (name + "_scaled(1x).bmp")))
        want.append(osc.score_pair(a, b, 10))
    mp, ms, my = scorpath.main(str(tmp_path) + "/")
    assert mp == pytest.approx(np.mean([w[0] for w in want]), abs=1e-7)
    assert ms == pytest.approx(np.mean([w[1] for w in want]), abs=1e-9)
    assert my == pytest.approx(np.mean([w[2] for w in want]), abs=1e-9)
    assert "SCOR MEAN psnr" in capsys.readouterr().out


def test_scoring_is_bit_reproducible_and_handles_odd_sizes():
    """All cross-block sums of the scoring kernel are 64-bit integer atomics (exact integer window arithmetic,
    2^-40 fixed-point SSIM values): repeated runs give bit-identical results whatever order the blocks finish in,
    and sizes that are not multiples of the 32-column band or the 8-row step agree with the oracle like the others."""
    from oracle import scoring as osc
    from sr100 import ops
    rng = np.random.default_rng(11)
    for h, w, crop in ((61, 83, 10), (27, 27, 10), (200, 33, 0), (95, 310, 3)):
        a = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
        ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        runs = [ops.score_pair(ad, bd, crop=crop) for _ in range(3)]
        assert runs[0] == runs[1] == runs[2]
        wp, wrgb, wy = osc.score_pair(a, b, crop)
        assert abs(runs[0]["psnr_y"] - wp) < 1e-7
        assert abs(runs[0]["ssim_y"] - wy) < 1e-9 and abs(runs[0]["ssim_rgb"] - wrgb) < 1e-9
    big_a = torch.randint(0, 256, (1356, 2040, 3), dtype=torch.uint8, device="cuda")
    big_b = torch.randint(0, 256, (1356, 2040, 3), dtype=torch.uint8, device="cuda")
    r = [ops.score_pair(big_a, big_b) for _ in range(4)]
    assert all(x == r[0] for x in r)


def test_scoring_band_and_chunk_geometry():
    """The scoring kernel walks 32-column bands in chunks of rows (8 pixel rows per step): shapes whose last band has
    pixels but no window (cropped width 33..38), the smallest legal image (7 x 7), tall narrow images (many chunks,
    one band), images shorter than one step, the extreme pixel values (largest 64-bit Y moments) and base pointers
    that are not 4-byte aligned (the pixel bytes come from aligned word loads + funnel shift)."""
    from oracle import scoring as osc
    from sr100 import ops
    rng = np.random.default_rng(21)
    cases = [(7, 7, 0), (8, 39, 0), (40, 33, 0), (40, 38, 0), (41, 71, 0), (900, 20, 0), (1500, 41, 3), (13, 300, 2),
             (70, 64, 0), (64, 70, 0)]
    for h, w, crop in cases:
        a = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-30, 31, size=a.shape), 0, 255).astype(np.uint8)
        r = ops.score_pair(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), crop=crop)
        wp, wrgb, wy = osc.score_pair(a, b, crop)
        assert abs(r["psnr_y"] - wp) < 1e-7, (h, w, crop)
        assert abs(r["ssim_y"] - wy) < 1e-9 and abs(r["ssim_rgb"] - wrgb) < 1e-9, (h, w, crop)
        assert r["n_pix"] == (h - 2 * crop) * (w - 2 * crop) and r["n_win"] == (h - 2 * crop - 6) * (w - 2 * crop - 6)
    # extremes: white vs black, white vs white-ish noise, black vs black-ish noise
    h, w = 50, 45
    white, black = np.full((h, w, 3), 255, np.uint8), np.zeros((h, w, 3), np.uint8)
    near_white = (255 - rng.integers(0, 3, size=(h, w, 3))).astype(np.uint8)
    for a, b in ((white, black), (white, near_white), (black, 255 - near_white)):
        r = ops.score_pair(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), crop=0)
        wp, wrgb, wy = osc.score_pair(a, b, 0)
        assert abs(r["psnr_y"] - wp) < 1e-7
        assert abs(r["ssim_y"] - wy) < 1e-9 and abs(r["ssim_rgb"] - wrgb) < 1e-9
    # misaligned base pointers: the images start 1, 2 and 3 bytes into their allocations
    h, w = 45, 50
    a = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-9, 10, size=a.shape), 0, 255).astype(np.uint8)
    want = ops.score_pair(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), crop=4)
    for oa, ob in ((1, 0), (2, 3), (3, 1)):
        fa = torch.zeros(h * w * 3 + 8, dtype=torch.uint8, device="cuda")
        fb = torch.zeros(h * w * 3 + 8, dtype=torch.uint8, device="cuda")
        va, vb = fa[oa:oa + h * w * 3].view(h, w, 3), fb[ob:ob + h * w * 3].view(h, w, 3)
        va.copy_(torch.from_numpy(a))
        vb.copy_(torch.from_numpy(b))
        assert ops.score_pair(va, vb, crop=4) == want


def test_batched_scoring_equals_per_pair_scoring():
    """sr_score_batch_u8: any mix of shapes in one launch per 32 pairs (40 pairs here: two launches), one read-back;
    every result is bit-identical to the pair scored on its own (the sums are integers, so the different chunking of a
    batched launch cannot show).  scorpath.main scores a directory through this entry point
    (test_scorpath_main_on_directory)."""
    from sr100 import ops
    rng = np.random.default_rng(33)
    shapes = [(27, 27), (64, 48), (90, 33), (45, 200), (130, 71)] * 8
    pairs = []
    for h, w in shapes:
        a = rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-15, 16, size=a.shape), 0, 255).astype(np.uint8)
        pairs.append((torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()))
    got = ops.score_pairs(pairs, crop=10)
    want = [ops.score_pair(a, b, crop=10) for a, b in pairs]
    assert got == want
    assert ops.score_pairs([], crop=10) == []
    with pytest.raises(ValueError):
        ops.score_pairs([(pairs[0][0], pairs[1][0])], crop=10)
    from sr100 import _lib as L
    with pytest.raises(L.SrError):
        ops.score_pairs([pairs[0], (pairs[0][0][:20].contiguous(), pairs[0][1][:20].contiguous())], crop=10)
