"""Generate golden fixtures by running the REFERENCE's own numpy code (read-only /root/reference)
under stub modules.  Runs only in the authoring container (the GPU box has no /root/reference);
the outputs are committed under tests/golden/.  TEST INFRASTRUCTURE -- never imported by the product.

    python oracle/refgen.py

What can run verbatim: img_utils.extract_patches_Step /
rebuild_from_patches_Step (img_utils.py:601-724), PSNR.py, the vendored sklearn patch functions of
imgpatch.py (source lines 24-338 exec'd; extract_patches needs arr[tuple(slices)] on numpy 2.x),
and cv2 for scorpath.rgb2ycbcrCV / ycbcr2rgb.  TensorFlow/Keras/skimage are not installable here, so
the float conv path and SSIM/rgb2ycbcr are NOT pinned by this script ("parity unpinned").
"""
import io
import contextlib
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _install_stubs():
    if not hasattr(np, "float"):
        np.float = float  # PSNR.im2double uses np.float (numpy < 1.24)
    misc = types.ModuleType("scipy.misc")
    saved = []

    def imsave(path, arr):
        saved.append((path, np.asarray(arr).shape))

    def imread(path, mode="RGB"):
        from PIL import Image
        return np.asarray(Image.open(path).convert(mode))

    def imresize(*a, **k):
        raise NotImplementedError("scipy.misc.imresize is not available")

    misc.imsave, misc.imread, misc.imresize, misc._saved = imsave, imread, imresize, saved
    import scipy
    scipy.misc = misc
    sys.modules["scipy.misc"] = misc
    filt = types.ModuleType("scipy.ndimage.filters")
    from scipy.ndimage import gaussian_filter
    filt.gaussian_filter = gaussian_filter
    sys.modules["scipy.ndimage.filters"] = filt
    for name in ("skimage", "skimage.util", "skimage.util.shape"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.util.shape"].view_as_windows = lambda *a, **k: None
    keras = types.ModuleType("keras")
    backend = types.ModuleType("keras.backend")
    backend.image_dim_ordering = lambda: "tf"
    keras.backend = backend
    sys.modules["keras"] = keras
    sys.modules["keras.backend"] = backend
    import sklearn.feature_extraction.image as ski
    if not hasattr(ski, "check_array"):
        from sklearn.utils import check_array
        ski.check_array = check_array
    if not hasattr(ski, "extract_patches"):
        ski.extract_patches = ski._extract_patches  # renamed in sklearn >= 0.24
    return misc


def load_img_utils():
    _install_stubs()
    os.environ["HOME"] = tempfile.mkdtemp(prefix="refhome_")  # import creates dataset dirs there
    sys.path.insert(0, REF)
    import img_utils  # noqa: E402  (the reference module, verbatim)
    return img_utils


def load_psnr():
    _install_stubs()
    sys.path.insert(0, REF)
    import PSNR  # noqa: E402
    return PSNR


def load_imgpatch_functions():
    """exec imgpatch.py lines 24-338 (skips the import-time os.listdir and the trailing script)."""
    import numbers
    from itertools import product
    from numpy.lib.stride_tricks import as_strided
    from sklearn.utils import check_array, check_random_state
    src = open(os.path.join(REF, "imgpatch.py")).read().splitlines()
    body = "\n".join(src[23:338])
    body = body.replace("arr[slices].strides", "arr[tuple(slices)].strides")  # numpy 2.x indexing
    ns = dict(np=np, numbers=numbers, product=product, as_strided=as_strided, check_array=check_array,
              check_random_state=check_random_state)
    exec(compile(body, "imgpatch.py[24:338]", "exec"), ns)
    return ns


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20181018)
    iu = load_img_utils()

    # ---- tiling: extract_patches_Step / rebuild_from_patches_Step on small canvases -------------
    tiling = {}
    cases = [  # (canvas_h, canvas_w, patch, step, scale)
        (24, 24, 12, 8, 4),     # 2x2 grid
        (32, 40, 12, 8, 4),     # 3x4 grid, non-square
        (16, 48, 12, 8, 2),     # single row of patches
        (13, 13, 12, 8, 4),     # single patch (dim - p = 1)
        (40, 24, 16, 8, 1),     # scale 1
    ]
    for ci, (ch, cw, p, st, sc) in enumerate(cases):
        crng = np.random.default_rng(1000 + ci)  # tests regenerate canvas/up from this seed
        canvas = crng.integers(0, 256, size=(ch, cw, 3)).astype(np.float64)
        patches, (cnt_h, cnt_w) = quiet(iu.extract_patches_Step, canvas, (p, p), st)
        up = crng.integers(-20, 281, size=(patches.shape[0], p * sc, p * sc, 3)).astype(np.float32)
        rebuilt = quiet(iu.rebuild_from_patches_Step, canvas, up, (p, p), (cnt_h, cnt_w), sc, st)
        tiling["c%d_meta" % ci] = np.array([ch, cw, p, st, sc, cnt_h, cnt_w], dtype=np.int64)
        tiling["c%d_canvas" % ci] = canvas.astype(np.uint8)
        tiling["c%d_patches_sum" % ci] = patches.sum(axis=(1, 2, 3))
        tiling["c%d_patches_first" % ci] = patches[0].astype(np.uint8)
        tiling["c%d_patches_last" % ci] = patches[-1].astype(np.uint8)
        tiling["c%d_rebuilt" % ci] = rebuilt.astype(np.int16)  # exact: up holds small integers
    # patch counts and canvas geometry of the reference arithmetic at the BASELINE shapes
    geo = []
    for (h, w) in [(512, 512), (288, 288), (256, 256), (280, 280), (344, 228), (339, 510), (1080, 1920),
                   (128, 128), (100, 37), (64, 64), (1, 1), (32, 160)]:
        ch, cw = h + 96, w + 96
        if cw % 64 != 0 or ch % 64 != 0:
            cw = int((cw / 64) + 1) * 64
            ch = int((ch / 64) + 1) * 64
        cnt_h = len([x for x in range(ch - 96) if x == 0 or x % 64 == 0])
        cnt_w = len([x for x in range(cw - 96) if x == 0 or x % 64 == 0])
        geo.append([h, w, ch, cw, cnt_h, cnt_w])
    tiling["geometry_96_64"] = np.array(geo, dtype=np.int64)
    # ValueError when the patch is larger than the image (img_utils.py:605-611)
    try:
        quiet(iu.extract_patches_Step, np.zeros((8, 30, 3)), (12, 12), 8)
        tiling["raises_h"] = np.array([0])
    except ValueError:
        tiling["raises_h"] = np.array([1])
    np.savez_compressed(os.path.join(OUT, "tiling_ref.npz"), **tiling)

    # ---- training data pipeline: _index_generator / image_generator (img_utils.py:290-398) ----------------
    gen = {}
    for gi, (N, bs, shuffle, seed) in enumerate([(10, 4, True, 3), (7, 7, True, 11), (5, 2, False, None), (3, 8, True, 0)]):
        g = iu._index_generator(N, bs, shuffle, seed)
        idx, cur, cbs = [], [], []
        for _ in range(9):
            a_, c_, b_ = next(g)
            idx.append(np.pad(np.asarray(a_), (0, bs - len(a_)), constant_values=-1))
            cur.append(c_)
            cbs.append(b_)
        gen["g%d_args" % gi] = np.array([N, bs, int(shuffle), -1 if seed is None else seed])
        gen["g%d_idx" % gi] = np.array(idx)
        gen["g%d_cur" % gi] = np.array(cur)
        gen["g%d_bs" % gi] = np.array(cbs)
    # image_generator on a temp dataset of 16x16 PNG pairs (the only shape the shipped defaults accept:
    # scale_factor=1 -> image_shape (16,16,3) for X and y, img_utils.py:302-309)
    from PIL import Image
    d = tempfile.mkdtemp(prefix="refdata_") + "/"
    os.makedirs(d + "X")
    os.makedirs(d + "y")
    drng = np.random.default_rng(77)
    for k in range(5):
        Image.fromarray(drng.integers(0, 256, size=(16, 16, 3)).astype(np.uint8)).save(d + "X/im%d.png" % k)
        Image.fromarray(drng.integers(0, 256, size=(16, 16, 3)).astype(np.uint8)).save(d + "y/im%d.png" % k)
    ig = quiet(lambda: iu.image_generator(d, scale_factor=1, shuffle=True, batch_size=2, seed=5))
    bx, by = quiet(next, ig)
    bx2, by2 = quiet(next, ig)
    gen["ig_bx"], gen["ig_by"], gen["ig_bx2"], gen["ig_by2"] = bx, by, bx2, by2
    np.savez_compressed(os.path.join(OUT, "generator_ref.npz"), **gen)

    # ---- PSNR.py ---------------------------------------------------------------------------------
    P = load_psnr()
    a = rng.integers(0, 256, size=(40, 52, 3)).astype(np.uint8)
    noise = rng.normal(0, 3, size=a.shape)
    b = np.clip(a.astype(np.float64) + noise, 0, 255).astype(np.uint8)
    ya = 16 + (65.481 * a[..., 0] + 128.553 * a[..., 1] + 24.966 * a[..., 2]) / 255.0
    yb = 16 + (65.481 * b[..., 0] + 128.553 * b[..., 1] + 24.966 * b[..., 2]) / 255.0
    psnr = dict(a=a, b=b, ya=ya, yb=yb)
    psnr["psnrNITRE_y"] = np.array(P.psnrNITRE(yb, ya, 0))
    psnr["psnrNITRE_y_shave4"] = np.array(P.psnrNITRE(yb, ya, 4))
    psnr["PSNRTorch_y"] = np.array(P.PSNRTorch(yb, ya, 0))
    psnr["PSNRTorch_same"] = np.array(P.PSNRTorch(ya, ya, 0))
    psnr["psnrVDSR_y_2"] = np.array(P.psnrVDSR(yb, ya, 2))
    psnr["psnrSVLAB_u8"] = np.array(P.psnrSVLAB(a, b))
    # integer inputs: the reference subtracts in the arrays' own dtype (uint8 wrap-around), and PSNRTorch also squares
    # in it (`imdff ** 2` with an int exponent stays uint8) -- the mirror reproduces both
    psnr["psnrVDSR_u8_2"] = np.array(P.psnrVDSR(b, a, 2))
    psnr["PSNRTorch_u8"] = np.array(P.PSNRTorch(b, a, 0))
    psnr["psnrNITRE_u8"] = np.array(P.psnrNITRE(b, a, 0))
    psnr["im2double_a"] = P.im2double(a)[:2, :3]
    psnr["im2doubleZ_a"] = P.im2doubleZ(a)[:2, :3]
    np.savez_compressed(os.path.join(OUT, "psnr_ref.npz"), **psnr)

    # ---- imgpatch.py (vendored sklearn) ------------------------------------------------------------
    ns = load_imgpatch_functions()
    one = np.arange(16).reshape(4, 4)
    dense = ns["extract_patches_2d"](one, (2, 2))
    ip = dict(doc_patches=dense)
    img = rng.integers(0, 256, size=(20, 28, 3)).astype(np.float64)
    d2 = ns["extract_patches_2d"](img, (8, 8))
    sel = quiet(ns["extract_patches_2dlocal"], img, d2, (8, 8), step=4)
    rec = ns["reconstruct_from_patches_2dlocal"](d2, sel, img.shape, step=4)
    avg = ns["reconstruct_from_patches_2d"](d2[:, :, :, :], img.shape, step=1)
    ip.update(img=img.astype(np.uint8), dense_shape=np.array(d2.shape), dense_sum=d2.sum(axis=(1, 2, 3)),
              sel_shape=np.array(sel.shape), sel_sum=sel.sum(axis=(1, 2, 3)), rec=rec.astype(np.float32),
              avg_step1=avg.astype(np.float32))
    avg4 = ns["reconstruct_from_patches_2d"](d2, img.shape, step=4)
    ip["avg_step4"] = avg4.astype(np.float32)
    ip["n_patches"] = np.array([ns["_compute_n_patches"](20, 28, 8, 8, None),
                                ns["_compute_n_patches"](20, 28, 8, 8, 10),
                                ns["_compute_n_patches"](20, 28, 8, 8, 0.5)])
    np.savez_compressed(os.path.join(OUT, "imgpatch_ref.npz"), **ip)

    # ---- cv2 colour path (scorpath.rgb2ycbcrCV / ycbcr2rgb, scorpath.py:48-62) ---------------------
    import cv2
    im_rgb = a.astype(np.float32)
    ycrcb = cv2.cvtColor(im_rgb, cv2.COLOR_RGB2YCR_CB)
    ycbcr = ycrcb[:, :, (0, 2, 1)].astype(np.float32)
    ycbcr[:, :, 0] = (ycbcr[:, :, 0] * (235 - 16) + 16) / 255.0
    ycbcr[:, :, 1:] = (ycbcr[:, :, 1:] * (240 - 16) + 16) / 255.0
    back = ycbcr.copy()
    back[:, :, 0] = (back[:, :, 0] * 255.0 - 16) / (235 - 16)
    back[:, :, 1:] = (back[:, :, 1:] * 255.0 - 16) / (240 - 16)
    rgb_back = cv2.cvtColor(back[:, :, (0, 2, 1)].astype(np.float32), cv2.COLOR_YCR_CB2RGB)
    np.savez_compressed(os.path.join(OUT, "cv2_colour_ref.npz"), a=a, ycbcr=ycbcr, rgb_back=rgb_back)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
