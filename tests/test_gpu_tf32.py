"""-m gpu: the tf32 option of the conv stack (sr_conv_desc.precision = 1, Engine(precision='tf32')).

The reference's Conv2D is fp32 (Keras/TF, models.py:1177-1199); BASELINE.json's north_star asks for max-abs <= 1e-4
against it on [0,1] outputs for the tf32 path.  Kernel-level: against the CPU oracle's conv on the SAME tf32-rounded
operands (fp32 accumulation on both sides -> only summation-order noise); model-level: against the fp32 oracle graph
with identical random-init weights."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32_round(a):
    """numpy restatement of cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits."""
    i = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((i + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def run_conv_tf32(lib, xs, ws, bias, relu=0, alpha=1.0, beta=0.0, res=None, cout=128, pair=0, comp=None):
    from sr100 import _lib as L
    dev = "cuda"
    NB, H, W, _ = xs[0].shape
    keep = []
    d = L.ConvDesc()
    d.nsrc = len(xs)
    for s, (x, w) in enumerate(zip(xs, ws)):
        k = w.shape[0]
        xd = torch.from_numpy(x).to(dev).contiguous()
        wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).to(dev)
        pk = torch.empty(lib.sr_packed_weight_bytes_tf32(k, cout), dtype=torch.uint8, device=dev)
        L.check(lib.sr_pack_conv_weights_tf32(L.ptr(wd), k, cout, L.ptr(pk), L.stream_ptr()))
        d.in_[s], d.wpacked[s], d.ksize[s] = xd.data_ptr(), pk.data_ptr(), k
        keep += [xd, wd, pk]
    d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, cout
    bd = torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32)).to(dev) if bias is not None else None
    d.bias = bd.data_ptr() if bd is not None else None
    d.alpha, d.beta, d.relu = alpha, beta, relu
    rd = torch.from_numpy(res).to(dev) if res is not None else None
    d.res_f32 = rd.data_ptr() if rd is not None else None
    of = torch.full((NB, H, W, cout), float("nan"), device=dev)
    ot = torch.full((NB, H, W, cout), float("nan"), device=dev) if cout == 128 else None
    d.out_f32 = of.data_ptr()
    d.out_tf32 = ot.data_ptr() if ot is not None else None
    d.a_mode, d.nacc, d.pair, d.precision = 0, 2, pair, 1
    if comp:
        d.comp_h, d.comp_w = comp
    plan = C.c_void_p()
    L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
    L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
    torch.cuda.synchronize()
    lib.sr_conv_plan_destroy(plan)
    return of.cpu().numpy(), (ot.cpu().numpy() if ot is not None else None)


def oracle_conv_tf32(xs, ws, bias, relu=0, alpha=1.0, beta=0.0, res=None):
    import torch.nn.functional as F
    acc = None
    for x, w in zip(xs, ws):
        k = w.shape[0]
        y = F.conv2d(torch.from_numpy(x).double().permute(0, 3, 1, 2),
                     torch.from_numpy(tf32_round(w)).double().permute(3, 2, 0, 1), padding=(k - 1) // 2)
        acc = y if acc is None else acc + y
    if bias is not None:
        acc = acc + torch.from_numpy(np.asarray(bias, dtype=np.float64)).view(1, -1, 1, 1)
    out = alpha * acc.permute(0, 2, 3, 1)
    if res is not None:
        out = out + beta * torch.from_numpy(res).double()
    if relu:
        out = out.clamp_min(0)
    return out.numpy()


CASES = [
    # name, ks, NB, H, W, cout, relu, alpha, beta, residual
    ("k1", (1,), 2, 20, 96, 128, 0, 1.0, 0.0, False),
    ("k3_relu", (3,), 2, 20, 96, 128, 1, 1.0, 0.0, False),
    ("k5", (5,), 2, 20, 96, 128, 0, 1.0, 0.0, False),
    ("dual_53_block_end", (5, 3), 2, 20, 96, 128, 0, 0.1, 0.9, True),
    ("light_block_end", (3,), 3, 20, 96, 128, 0, 0.1, 1.0, True),
    ("tail_cout3", (3,), 2, 20, 96, 3, 1, 1.0, 0.0, False),
    ("ragged_50x33", (3,), 3, 33, 50, 128, 0, 1.0, 0.0, False),
    ("wide_384", (5,), 1, 12, 384, 128, 0, 1.0, 0.0, False),
    ("single_pixel", (5,), 1, 1, 1, 128, 0, 1.0, 0.0, False),
    ("one_col", (5,), 2, 70, 1, 128, 0, 1.0, 0.0, False),
    ("lr_tile_96", (5,), 2, 96, 96, 128, 1, 1.0, 0.0, False),
]


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tf32_conv_matches_oracle(lib, case, pair):
    name, ks, NB, H, W, cout, relu, alpha, beta, with_res = case
    rng = np.random.default_rng(__import__("zlib").crc32(name.encode()))
    xs = [tf32_round(rng.standard_normal((NB, H, W, 128)).astype(np.float32) * 0.5) for _ in ks]
    ws = [rng.standard_normal((k, k, 128, cout)).astype(np.float32) / np.sqrt(k * k * 128) for k in ks]
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    res = rng.standard_normal((NB, H, W, cout)).astype(np.float32) if with_res else None
    got, got_t = run_conv_tf32(lib, xs, ws, bias, relu, alpha, beta, res, cout, pair)
    want = oracle_conv_tf32(xs, ws, bias, relu, alpha, beta, res)
    assert not np.isnan(got).any(), "some output pixels were never written"
    # identical tf32 operands, exact products, fp32 accumulation in TMEM vs the float64 oracle sum: only the
    # accumulator's rounding (K up to 4352 terms, partial sums of magnitude ~1: a few fp32 ulps of 1.0) is left
    assert np.abs(got - want).max() <= 6e-5, float(np.abs(got - want).max())
    if got_t is not None:
        assert np.array_equal(got_t.view(np.uint32), tf32_round(got).view(np.uint32))    # bit-exact rounding


def test_tf32_compute_extent_leaves_the_rest_untouched(lib):
    rng = np.random.default_rng(5)
    x = tf32_round(rng.standard_normal((2, 40, 56, 128)).astype(np.float32))
    w = rng.standard_normal((3, 3, 128, 128)).astype(np.float32) / 34.0
    got, _ = run_conv_tf32(lib, [x], [w], None, comp=(21, 37))
    want = oracle_conv_tf32([x], [w], None)
    assert np.abs(got[:, :21, :37] - want[:, :21, :37]).max() <= 6e-5
    assert np.isnan(got[:, 21:]).all() and np.isnan(got[:, :, 37:]).all()


def test_tf32_rejects_bf16_tensors(lib):
    from sr100 import _lib as L
    d = L.ConvDesc()
    buf = torch.zeros(1, 8, 8, 128, device="cuda")
    pk = torch.zeros(lib.sr_packed_weight_bytes_tf32(3, 128), dtype=torch.uint8, device="cuda")
    d.nsrc, d.NB, d.H, d.W, d.cin, d.cout, d.precision = 1, 1, 8, 8, 128, 128, 1
    d.in_[0], d.wpacked[0], d.ksize[0] = buf.data_ptr(), pk.data_ptr(), 3
    d.out_bf16 = buf.data_ptr()
    plan = C.c_void_p()
    assert lib.sr_conv_plan_create(C.byref(d), C.byref(plan)) == -2
    d.out_bf16, d.precision, d.out_tf32 = None, 0, buf.data_ptr()
    assert lib.sr_conv_plan_create(C.byref(d), C.byref(plan)) == -1


def test_round_tf32_kernel_bit_exact(lib):
    from sr100 import _lib as L
    rng = np.random.default_rng(2)
    a = (rng.standard_normal(1 << 16) * np.exp(rng.uniform(-20, 20, 1 << 16))).astype(np.float32)
    a[:8] = [0.0, -0.0, 1.0, -1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 + 2.0 ** -20, -(1.0 + 2.0 ** -11), 3.4e38 / 4]
    x = torch.from_numpy(a).cuda()
    y = torch.empty_like(x)
    L.check(lib.sr_round_tf32(L.ptr(x), x.numel(), L.ptr(y), L.stream_ptr()))
    assert np.array_equal(y.cpu().numpy().view(np.uint32), tf32_round(a).view(np.uint32))


def _smooth_images(rng, n, h, w):
    from scipy.ndimage import uniform_filter
    img = rng.integers(0, 256, size=(n, h + 4, w + 4, 3)).astype(np.float32)
    img = uniform_filter(img, size=(1, 5, 5, 1))[:, 2:-2, 2:-2]
    return (img / 255.0).astype(np.float32)


@pytest.mark.parametrize("shape", [(2, 24, 24), (1, 48, 40), (1, 96, 96)])
def test_tf32_model_within_1e4_of_fp32_graph(shape):
    """The north_star bound for the tf32 path: max-abs <= 1e-4 on the model's outputs against the fp32 graph with
    identical random-init (glorot_uniform) weights, and closer than the bf16 engine."""
    from oracle import model as om
    from sr100.engine import Engine
    weights = om.init_weights(1234, bias_scale=0.01)
    rng = np.random.default_rng(shape[1])
    x = _smooth_images(rng, *shape)
    want = om.forward_numpy(weights, x.astype(np.float64), dtype=torch.float64)
    xd = torch.from_numpy(x).cuda()
    got = Engine(weights, precision="tf32").forward_device(xd).cpu().numpy()
    ref16 = Engine(weights).forward_device(xd).cpu().numpy()
    err, err16 = float(np.abs(got - want).max()), float(np.abs(ref16 - want).max())
    assert got.shape == want.shape and got.min() >= 0
    assert err <= 1e-4, err
    assert err < err16, (err, err16)


def test_tf32_model_on_order_one_outputs():
    """Same comparison with the tail lifted so the outputs span [0.25, 0.75] instead of ~0.05: the absolute error
    scales with the output, 5e-4 holds (tf32 operand rounding is 2^-12 relative)."""
    from oracle import model as om
    from sr100.engine import Engine
    weights = om.init_weights(1234, bias_scale=0.01)
    w, b = weights["conv2d_85"]
    weights["conv2d_85"] = (w * 8.0, b + 0.3)
    x = _smooth_images(np.random.default_rng(0), 1, 32, 32)
    want = om.forward_numpy(weights, x.astype(np.float64), dtype=torch.float64)
    got = Engine(weights, precision="tf32").forward_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert want.max() > 0.5
    assert np.abs(got - want).max() <= 5e-4, float(np.abs(got - want).max())


def test_tf32_tiled_upscale_and_dead_work_elimination():
    """upscaleStepPatch's device path in tf32 mode: uint8 image within one level of the oracle's, and the
    dead-work-eliminated run bit-identical to the literal full tiling inside the final image."""
    import models
    from oracle import model as om
    from oracle import tiling as ot
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(3)
    img = uniform_filter(rng.integers(0, 256, size=(70, 45, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)
    weights = om.init_weights(1234, bias_scale=0.01)
    w, b = weights["conv2d_85"]
    weights["conv2d_85"] = (w * 8.0, b + 0.3)
    m = models.DifvdsrDouble(1)
    m.precision = "tf32"
    model = m.create_model(96, 96)
    assert model.engine.tf32
    model.engine.set_weights_dict(weights)
    got = m.upscale_arrays([img], patch_size=96)[0]
    _, want = ot.upscale_step_patch(img, lambda x: om.forward_numpy(weights, x), 96, 64, 4)
    assert got.shape == (280, 180, 3)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    assert (got != want).mean() < 2e-2          # astype(uint8) truncates: an error of 0.05 levels flips ~1 % of the pixels
    dev = torch.from_numpy(img).cuda()
    a = model.engine.upscale_images_device([dev], patch=96, step=64)[0].cpu().numpy()
    full = model.engine.upscale_images_device([dev], patch=96, step=64, full_canvas=True)[0].cpu().numpy()
    assert np.array_equal(a, full[:280, :180])


def test_training_needs_the_bf16_engine():
    from sr100.engine import Engine
    from sr100.train import Trainer
    with pytest.raises(NotImplementedError):
        Trainer(Engine(precision="tf32"))
