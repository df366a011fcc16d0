"""CPU: the checker side of the alternative tilers - oracle/pil_resample.py is pinned bit for bit to
the Pillow of this image, the product's coefficient table (sr100.alt_tilers.pil_bicubic_coeffs) equals the oracle's,
and the restated tilers behave (identity-like predict reproduces a smooth image closely)."""
import numpy as np
import pytest
from PIL import Image

from oracle import alt_tilers as oat
from oracle import pil_resample as pr


@pytest.mark.parametrize("h,w,oh,ow", [(32, 32, 8, 8), (128, 128, 32, 32), (16, 16, 4, 4), (10, 9, 40, 36),
                                       (17, 23, 68, 92), (64, 48, 16, 12), (8, 8, 2, 2), (5, 7, 5, 28)])
def test_resize_restatement_is_bit_exact_vs_pillow(h, w, oh, ow):
    rng = np.random.default_rng(h * 100 + w)
    a = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(pr.resize_bicubic_u8(a, oh, ow), want)


def test_bytescale_semantics():
    a = np.array([[10.0, 20.0], [30.0, 50.0]])
    assert np.array_equal(pr.bytescale(a), np.array([[0, 64], [128, 255]], dtype=np.uint8))     # (v-10)*6.375 + .5
    assert np.array_equal(pr.bytescale(np.full((2, 2), 7.0)), np.zeros((2, 2), dtype=np.uint8))  # flat: cscale = 1
    u = np.array([[1, 2]], dtype=np.uint8)
    assert pr.bytescale(u) is u                                                                   # uint8 passes through


def test_product_coefficients_equal_oracle():
    from sr100 import alt_tilers as at
    for n_in, n_out in [(32, 8), (128, 32), (16, 4), (96, 24), (8, 2)]:
        b0, k0 = pr.precompute_coeffs(n_in, n_out)
        b1, k1 = at.pil_bicubic_coeffs(n_in, n_out)
        assert np.array_equal(b0, b1) and np.array_equal(k0, k1)
        assert k0.shape[1] == 17 and abs(int(k0[n_out // 2].sum()) - (1 << 22)) <= 8


def _nearest_x4(x):
    return np.repeat(np.repeat(x, 4, axis=1), 4, axis=2)


def test_restated_tilers_with_a_x4_replicating_predict():
    rng = np.random.default_rng(1)
    img = np.zeros((21, 18, 3), dtype=np.uint8)
    img[...] = np.linspace(0, 255, 18).astype(np.uint8)[None, :, None]
    img[3:9, 2:7] = 255                                             # full range inside most patches
    out = oat.upscale_patch(img, _nearest_x4, patch_size=16)
    assert out.shape == img.shape and out.dtype == np.uint8
    big, out4 = oat.upscale_patch_mode(img[:6, :5], _nearest_x4, patch_size=8)
    assert big.shape == (24, 20, 3) and out4.shape == (24, 20, 3)
    assert np.abs(out4.astype(int) - big.astype(int)).mean() < 40    # shrink + replicate + average ~ a blur of `big`
    with pytest.raises(ValueError):
        oat.upscale_patch(img[:8, :8], _nearest_x4, patch_size=16)   # patch larger than the image


# ---- goldens produced by running the reference's own img_utils helpers (oracle/refgen_alt.py) -----------------------
@pytest.fixture(scope="module")
def gold(golden_dir):
    import os
    return np.load(os.path.join(golden_dir, "alt_tilers_ref.npz"))


def test_oracle_reconstruct_local_matches_reference(gold):
    img = gold["img"]
    want = gold["rec_local"]
    got = oat.reconstruct_from_patches_2dlocal((13, 21), (8, 8), gold["cnn"], img.shape, 4)
    assert np.isnan(want).any()                     # p = 8, pad = 4: interior patches contribute nothing
    assert np.array_equal(got, want, equal_nan=True)
    for tag in ("b", "c"):
        h, w, p, n = gold["%s_meta" % tag]
        got = oat.reconstruct_from_patches_2dlocal((h - p + 1, w - p + 1), (p, p), gold["%s_cnn" % tag], (h, w, 3), 4)
        assert np.array_equal(got, gold["%s_rec" % tag], equal_nan=True)


def test_extraction_mirrors_match_reference(gold):
    import img_utils
    img = gold["img"]
    full = img_utils.make_patchesOrig(img.astype(np.float64), 1, 8)
    assert tuple(gold["orig_shape"]) == full.shape and full.dtype == np.uint8
    assert np.array_equal(full.reshape(full.shape[0], -1).sum(1), gold["orig_sum"])
    sel = img_utils.extract_patches_2dlocal(img.astype(np.float64), full, (8, 8), step=4)
    assert tuple(gold["sel_shape"]) == sel.shape and sel.dtype == np.float64
    assert np.array_equal(sel.reshape(sel.shape[0], -1).sum(1), gold["sel_sum"])
    dense = img_utils.make_patches(img, 1, 8)
    assert tuple(gold["dense_shape"]) == dense.shape
    sub = img_utils.subimage_build_patch_global(img, 6, 8, 0)
    assert tuple(gold["sub_shape"]) == sub.shape
    assert np.array_equal(sub.reshape(sub.shape[0], -1).sum(1), gold["sub_sum"])
    with pytest.raises(ValueError):
        img_utils.extract_patches_2dlocal(img[:4], full, (8, 8), step=4)
    # scipy.misc.imresize restatement: uint8 goes straight to Pillow, float input is contrast-stretched first
    a = img[:16, :12]
    assert np.array_equal(img_utils.imresize(a, (4, 3), interp='bicubic'), pr.imresize_bicubic(a, (4, 3)))
    f = a.astype(np.float64) / 4
    assert np.array_equal(img_utils.imresize(f, (4, 3), interp='bicubic'), pr.imresize_bicubic(f, (4, 3)))
    assert img_utils.imresize(a, 50).shape == (8, 6, 3) and img_utils.imresize(a, 0.25).shape == (4, 3, 3)
