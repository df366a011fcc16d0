"""CPU tests: the oracle against the reference's own outputs (tests/golden, produced by oracle/refgen.py
running the reference code verbatim) and against closed-form properties."""
import numpy as np
import pytest

from oracle import model as om
from oracle import scoring as osc
from oracle import shuffle as osh
from oracle import tiling as ot


def test_tiling_matches_reference_golden(golden_dir):
    z = np.load(golden_dir + "/tiling_ref.npz")
    for ci in range(5):
        ch, cw, p, st, sc, cnt_h, cnt_w = [int(v) for v in z["c%d_meta" % ci]]
        crng = np.random.default_rng(1000 + ci)
        canvas = crng.integers(0, 256, size=(ch, cw, 3)).astype(np.float64)
        assert np.array_equal(canvas.astype(np.uint8), z["c%d_canvas" % ci])
        patches, counts = ot.extract_patches_step(canvas, (p, p), st)
        assert counts == (cnt_h, cnt_w)
        assert np.array_equal(patches.sum(axis=(1, 2, 3)), z["c%d_patches_sum" % ci])
        assert np.array_equal(patches[0].astype(np.uint8), z["c%d_patches_first" % ci])
        assert np.array_equal(patches[-1].astype(np.uint8), z["c%d_patches_last" % ci])
        up = crng.integers(-20, 281, size=(patches.shape[0], p * sc, p * sc, 3)).astype(np.float32)
        rebuilt = ot.rebuild_from_patches_step((ch, cw), up, (p, p), counts, sc, st)
        assert np.array_equal(rebuilt.astype(np.int16), z["c%d_rebuilt" % ci])
    assert int(z["raises_h"][0]) == 1
    with pytest.raises(ValueError):
        ot.extract_patches_step(np.zeros((8, 30, 3)), (12, 12), 8)


def test_canvas_geometry_matches_reference(golden_dir):
    geo = np.load(golden_dir + "/tiling_ref.npz")["geometry_96_64"]
    for h, w, ch, cw, cnt_h, cnt_w in geo.tolist():
        assert ot.canvas_size(h, w, 96, 64) == (ch, cw)
        _, counts = ot.extract_patches_step(np.zeros((ch, cw, 3)), (96, 96), 64)
        assert counts == (cnt_h, cnt_w)
    # Set5 shapes give 81/25/25/25/30 tiles, DIV2K-shape 54, 1080p 558
    tiles = {(int(r[0]), int(r[1])): int(r[4] * r[5]) for r in geo}
    assert tiles[(512, 512)] == 81 and tiles[(288, 288)] == 25 and tiles[(344, 228)] == 30
    assert tiles[(339, 510)] == 54 and tiles[(1080, 1920)] == 558


def test_identity_network_stitch_property():
    """an 'identity' network (nearest-neighbour x4 of each patch) stitches back to the
    nearest-neighbour x4 of the image -> defines patch order and ownership."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(70, 45, 3)).astype(np.uint8)
    full, out = ot.upscale_step_patch(img, lambda x: np.repeat(np.repeat(x, 4, axis=1), 4, axis=2), 24, 16, 4)
    want = np.repeat(np.repeat(img, 4, axis=0), 4, axis=1)
    # float32 (v/255)*255 may land just below v -> truncation can lose 1; compare with that tolerance
    assert out.shape == want.shape
    assert np.abs(out.astype(int) - want.astype(int)).max() <= 1


def test_psnr_matches_reference_golden(golden_dir):
    z = np.load(golden_dir + "/psnr_ref.npz")
    ya, yb, a, b = z["ya"], z["yb"], z["a"], z["b"]
    assert osc.psnr_nitre(yb, ya, 0) == pytest.approx(float(z["psnrNITRE_y"]), abs=1e-10)
    assert osc.psnr_nitre(yb, ya, 4) == pytest.approx(float(z["psnrNITRE_y_shave4"]), abs=1e-10)
    assert osc.psnr_torch(yb, ya, 0) == pytest.approx(float(z["PSNRTorch_y"]), abs=1e-10)
    assert osc.psnr_torch(ya, ya, 0) == float(z["PSNRTorch_same"]) == 100
    assert osc.psnr_vdsr(yb, ya, 2) == pytest.approx(float(z["psnrVDSR_y_2"]), abs=1e-10)
    assert osc.psnr_svlab(a, b) == pytest.approx(float(z["psnrSVLAB_u8"]), abs=1e-10)
    # PSNRTorch == psnrNITRE on 0..255 inputs
    assert float(z["PSNRTorch_y"]) == pytest.approx(float(z["psnrNITRE_y"]), abs=1e-9)
    # the Y formula used by the fixture equals the skimage restatement
    assert np.allclose(osc.rgb2ycbcr_y(a), ya, atol=1e-10)


def test_imgpatch_docstring_known_answer(golden_dir):
    """imgpatch.py:193-213 (sklearn docstring): extract_patches_2d(arange(16).reshape(4,4), (2,2))."""
    z = np.load(golden_dir + "/imgpatch_ref.npz")
    p = z["doc_patches"]
    assert p.shape == (9, 2, 2)
    assert p[0].tolist() == [[0, 1], [4, 5]] and p[1].tolist() == [[1, 2], [5, 6]] and p[8].tolist() == [[10, 11], [14, 15]]


def test_ssim_basic_properties():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, size=(40, 50, 3)).astype(np.uint8)
    assert osc.ssim(a, a, 255.0, multichannel=True) == pytest.approx(1.0, abs=1e-12)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    s = osc.ssim(a, b, 255.0, multichannel=True)
    assert 0.5 < s < 1.0
    assert osc.ssim(a, b, 255.0, multichannel=True) == pytest.approx(osc.ssim(b, a, 255.0, multichannel=True), abs=1e-12)


def test_shuffle_index_maps():
    """The three orderings of SURVEY 8a-6, derived here by replaying the reference tensor programs."""
    r, C, H, W = 2, 3, 3, 4
    x = np.arange(1 * H * W * C * r * r, dtype=np.float64).reshape(1, H, W, C * r * r)
    a = osh.phase_shift_subpixel(x, r)
    b = osh.depth_to_scale_tf(x, r, C)
    c = osh.depth_to_scale_th(x.transpose(0, 3, 1, 2), r, C).transpose(0, 2, 3, 1)
    d = osh.depth_to_space_tf(x, r)
    for Y in range(H * r):
        for X in range(W * r):
            for ch in range(C):
                src = x[0, Y // r, X // r]
                assert a[0, Y, X, ch] == src[ch * r * r + (X % r) * r + (Y % r)]
                assert b[0, Y, X, ch] == src[ch * r * r + (X % r) * r + (Y % r)]
                assert c[0, Y, X, ch] == src[ch * r * r + (Y % r) * r + (X % r)]
                assert d[0, Y, X, ch] == src[((Y % r) * r + (X % r)) * C + ch]


def test_model_structure_counts():
    specs = om.layer_specs()
    assert len(specs) == 86
    assert om.n_params() == 21838211
    assert [s[1] for s in specs[1:5]] == [3, 5, 5, 3]     # creation order inside a 5/3 block
    assert specs[-1] == ("conv2d_85", 3, 128, 3)
    flop_lr = 2 * 3 * 128 + 16 * (2 * 9 + 2 * 25) * 128 * 128 * 2 + 6 * 18 * 128 * 128 * 2
    flop_hr = 16 * (2 * 68 * 128 * 128 * 2 + 9 * 128 * 3 * 2)
    assert flop_lr + flop_hr == 110605056


def test_bilinear_legacy_semantics():
    import torch
    x = torch.arange(12, dtype=torch.float32).reshape(1, 1, 3, 4)
    y = om.bilinear_x4_tf1(x)[0, 0]
    assert y.shape == (12, 16)
    assert float(y[0, 0]) == 0.0 and float(y[0, 1]) == 0.25 and float(y[0, 4]) == 1.0
    assert float(y[0, 13]) == float(y[0, 12]) == 3.0          # last 3 columns replicate the edge
    assert float(y[4, 0]) == 4.0 and float(y[1, 0]) == 1.0    # row step 4 per LR row
    assert float(y[11, 15]) == 11.0


def test_oracle_forward_small_runs_and_is_deterministic():
    w = om.init_weights(7, bias_scale=0.01)
    x = np.random.default_rng(0).random((1, 8, 8, 3)).astype(np.float32)
    y1 = om.forward_numpy(w, x)
    y2 = om.forward_numpy(w, x)
    assert y1.shape == (1, 32, 32, 3) and np.array_equal(y1, y2) and y1.min() >= 0
    y64 = om.forward_numpy(w, x, dtype=__import__("torch").float64)
    assert np.abs(y1 - y64).max() < 1e-5


def test_keras_adam_first_step():
    import torch
    p = torch.nn.Parameter(torch.tensor([1.0, -2.0]))
    opt = om.KerasAdam([p], lr=1e-4)
    opt.step([torch.tensor([0.5, -0.25])])
    # t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1) -> step ~= lr*sign(g) (eps=1e-7 slightly less)
    assert torch.allclose(p.detach(), torch.tensor([1.0 - 1e-4, -2.0 + 1e-4]), atol=1e-8)


def test_oracle_conv_is_keras_conv2d_by_definition():
    """The torch conv inside oracle/model.py against the definition Keras documents for Conv2D(padding='same',
    strides 1): out[n,y,x,co] = b[co] + sum_{ky,kx,ci} in[n, y+ky-p, x+kx-p, ci] * kernel[ky,kx,ci,co] (cross-
    correlation, HWIO kernel, zero padding p = (k-1)/2) -- written out as plain numpy loops, no library conv."""
    import torch
    from oracle import model as om
    rng = np.random.default_rng(3)
    for k, cin, cout in ((1, 3, 5), (3, 4, 6), (5, 2, 3)):
        x = rng.standard_normal((2, 6, 7, cin))
        w = rng.standard_normal((k, k, cin, cout))
        b = rng.standard_normal(cout)
        p = (k - 1) // 2
        xp = np.zeros((2, 6 + 2 * p, 7 + 2 * p, cin))
        xp[:, p:p + 6, p:p + 7] = x
        want = np.zeros((2, 6, 7, cout))
        for ky in range(k):
            for kx in range(k):
                want += np.einsum("nyxc,co->nyxo", xp[:, ky:ky + 6, kx:kx + 7], w[ky, kx])
        want += b
        # the same call pattern as DifvdsrDoubleOracle.conv: HWIO -> OIHW, F.conv2d, padding (k-1)//2
        got = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2),
                                         torch.from_numpy(w).permute(3, 2, 0, 1).contiguous(), torch.from_numpy(b),
                                         padding=p).permute(0, 2, 3, 1).numpy()
        assert np.abs(got - want).max() < 1e-12
    # and through the oracle class itself on the first (1x1) layer: relu(x @ W + b)
    weights = om.init_weights(5, bias_scale=0.1)
    m = om.DifvdsrDoubleOracle(weights, dtype=torch.float64)
    x = rng.random((1, 4, 5, 3))
    y = m.conv("level1", torch.from_numpy(x).permute(0, 3, 1, 2), relu=True).permute(0, 2, 3, 1).detach().numpy()
    w0, b0 = weights["level1"]
    assert np.abs(y - np.maximum(x @ w0[0, 0].astype(np.float64) + b0, 0)).max() < 1e-12


def test_bf16_operand_emulation_is_the_same_graph():
    """oracle/emu_bf16.py (the engine's rounding points laid over the oracle graph, used by the sharp gradient test
    on the GPU) must be the SAME function as the fp32 oracle up to bf16 rounding: loss within 1e-4 relative, every
    layer's gradient within 3e-2 relative L2 -- and not identical (the rounding points are really there)."""
    import torch
    from oracle import emu_bf16 as emu
    from oracle import model as om
    w = om.init_weights(1234, bias_scale=0.01)
    rng = np.random.default_rng(7)
    x = rng.random((1, 8, 10, 3)).astype(np.float32)
    y = rng.random((1, 32, 40, 3)).astype(np.float32)
    loss, g = emu.gradients(w, x, y)
    m = om.DifvdsrDoubleOracle(w)
    want_loss = om.mse_loss(m(torch.from_numpy(x)), torch.from_numpy(y))
    grads = torch.autograd.grad(want_loss, list(m.parameters()))
    by = {}
    for (pn, _), q in zip(m.named_parameters(), grads):
        kind, name = pn.split(".")
        by.setdefault(name, {})[kind] = q.numpy()
    assert abs(loss - float(want_loss.detach())) <= 1e-4 * float(want_loss.detach())
    worst = 0.0
    for name in m.names:
        gw = np.transpose(by[name]["w"], (2, 3, 1, 0)).astype(np.float64)
        rel = np.linalg.norm(g[name][0] - gw) / np.linalg.norm(gw)
        relb = np.linalg.norm(g[name][1] - by[name]["b"]) / max(np.linalg.norm(by[name]["b"]), 1e-30)
        worst = max(worst, rel, relb)
        assert rel <= 3e-2 and relb <= 3e-2, (name, rel, relb)
    assert worst > 1e-4


# ---------------------------------------------------------------------------------------------------------------
# Cross-checks of the float-path restatements against INDEPENDENT implementations that are installed here (scipy,
# torch.optim, numpy).  Keras 2 / TF 1 / skimage themselves are not installable offline, so what stays
# unpinned after these is only the library's documented *definition* (which coordinates TF1's legacy bilinear samples,
# where Keras puts epsilon), not the arithmetic that follows from it.

def test_bilinear_restatement_equals_scipy_linear_interpolation():
    """tf.image.resize_bilinear (TF1 default: align_corners=False, no half-pixel centres) samples the source at
    dst * (in / out) and clamps at the far edge: scipy.ndimage.map_coordinates(order=1, mode='nearest') evaluated
    at exactly those coordinates is an independent implementation of the same interpolation."""
    import torch
    from scipy.ndimage import map_coordinates
    rng = np.random.default_rng(3)
    for h, w in ((5, 7), (12, 9), (1, 4)):
        x = rng.random((h, w))
        got = om.bilinear_x4_tf1(torch.from_numpy(x).reshape(1, 1, h, w))[0, 0].numpy()
        yy, xx = np.meshgrid(np.arange(4 * h) * 0.25, np.arange(4 * w) * 0.25, indexing="ij")
        want = map_coordinates(x, [yy, xx], order=1, mode="nearest")
        assert got.shape == want.shape == (4 * h, 4 * w)
        assert np.abs(got - want).max() <= 1e-12


def test_keras_adam_restatement_equals_torch_adam_up_to_epsilon_placement():
    """Keras 2: p -= lr * sqrt(1 - b2^t) / (1 - b1^t) * m / (sqrt(v) + eps).  torch.optim.Adam divides sqrt(v) by
    sqrt(1 - b2^t) BEFORE adding its eps, i.e. it is the same update with eps_torch = eps_keras / sqrt(1 - b2^t):
    feeding torch that per-step epsilon must reproduce the restatement step for step."""
    import math
    import torch
    rng = np.random.default_rng(5)
    p0 = rng.standard_normal(257)
    pk = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    pt = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-7
    keras = om.KerasAdam([pk], lr=lr, beta_1=b1, beta_2=b2, epsilon=eps)
    ref = torch.optim.Adam([pt], lr=lr, betas=(b1, b2), eps=eps)
    for t in range(1, 8):
        g = torch.from_numpy(rng.standard_normal(257) * 10.0 ** rng.integers(-6, 1))
        keras.step([g])
        ref.param_groups[0]["eps"] = eps / math.sqrt(1.0 - b2 ** t)
        pt.grad = g.clone()
        ref.step()
        assert float((pk.detach() - pt.detach()).abs().max()) <= 1e-12, t


def test_ssim_restatement_equals_window_by_window_definition():
    """compare_ssim(win_size=7, uniform window, use_sample_covariance=True, data_range=255): for every 7x7 window
    the sample means / variances / covariance straight from numpy (np.var / np.cov with ddof=1), the SSIM formula of
    Wang et al. with K1 = 0.01, K2 = 0.03, and the mean over all windows that lie inside the image (= the map
    cropped by 3 px) -- against the filtered-moments form the oracle (like skimage) evaluates."""
    from numpy.lib.stride_tricks import sliding_window_view
    rng = np.random.default_rng(9)
    x = rng.integers(0, 256, size=(19, 23)).astype(np.float64)
    y = np.clip(x + rng.normal(0, 12, size=x.shape), 0, 255)
    wx = sliding_window_view(x, (7, 7)).reshape(-1, 49)
    wy = sliding_window_view(y, (7, 7)).reshape(-1, 49)
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    vals = []
    for a, b in zip(wx, wy):
        ux, uy = a.mean(), b.mean()
        vx, vy = a.var(ddof=1), b.var(ddof=1)
        vxy = np.cov(a, b, ddof=1)[0, 1]
        vals.append((2 * ux * uy + c1) * (2 * vxy + c2) / ((ux * ux + uy * uy + c1) * (vx + vy + c2)))
    assert abs(osc.ssim(x, y, 255.0) - float(np.mean(vals))) <= 1e-10
    rgb_a = rng.integers(0, 256, size=(15, 16, 3)).astype(np.float64)
    rgb_b = np.clip(rgb_a + rng.normal(0, 5, size=rgb_a.shape), 0, 255)
    per = [osc.ssim(rgb_a[..., c], rgb_b[..., c], 255.0) for c in range(3)]
    assert abs(osc.ssim(rgb_a, rgb_b, 255.0, multichannel=True) - float(np.mean(per))) <= 1e-15


def test_mse_and_glorot_follow_their_keras_definitions():
    """loss='mse' = mean over ALL elements (keras.losses.mean_squared_error averaged over the batch);
    glorot_uniform = U(+-sqrt(6 / (fan_in + fan_out))) with fan = kh * kw * channels (keras.initializers)."""
    import torch
    rng = np.random.default_rng(1)
    a, b = rng.random((2, 8, 8, 3)), rng.random((2, 8, 8, 3))
    assert abs(float(om.mse_loss(torch.from_numpy(a), torch.from_numpy(b))) - float(((a - b) ** 2).mean())) <= 1e-15
    w = om.init_weights(3)
    for name, k, cin, cout in om.layer_specs():
        kern, bias = w[name]
        limit = np.sqrt(6.0 / (k * k * cin + k * k * cout))
        assert kern.shape == (k, k, cin, cout) and np.abs(kern).max() <= limit and not bias.any()
        if kern.size > 10000:
            assert np.abs(kern).max() >= 0.99 * limit and abs(kern.mean()) <= 0.02 * limit
