"""-m gpu: the reference's two older graphs, Difvdsr4 (models.py:992-1142) and Difvdsr (models.py:1274-1357), run by
sr100.planenet on the tensor-core conv kernel (two 128-channel planes per tensor) against the CPU restatement
oracle/other_models.py (torch fp32).  Tolerance: BASELINE.json north_star's bf16 bound, max-abs 2e-2 on outputs of
O(1) magnitude; the LeakyReLU epilogue and the x2 bilinear kernel are also checked alone (bit-exact / fp32-exact)."""
import ctypes as C

import numpy as np
import pytest
import torch

from gpu_util import bf16_round, oracle_conv

pytestmark = pytest.mark.gpu


def test_leaky_relu_epilogue_matches_oracle(lib):
    from sr100 import _lib as L
    rng = np.random.default_rng(2)
    NB, H, W = 2, 9, 20
    x = bf16_round(rng.normal(0, 1, size=(NB, H, W, 128)))
    w = (rng.standard_normal((3, 3, 128, 128)) / np.sqrt(9 * 128)).astype(np.float32)
    bias = rng.normal(0, 0.1, size=128).astype(np.float32)
    for slope in (0.2, 0.001):
        xd = torch.from_numpy(x).cuda().to(torch.bfloat16)
        wd = torch.from_numpy(w).cuda()
        pk = torch.empty(lib.sr_packed_weight_bytes(3, 128), dtype=torch.uint8, device="cuda")
        L.check(lib.sr_pack_conv_weights(L.ptr(wd), 3, 128, 0, L.ptr(pk), L.stream_ptr()))
        bd = torch.from_numpy(bias).cuda()
        of = torch.empty(NB, H, W, 128, device="cuda")
        d = L.ConvDesc()
        d.nsrc = 1
        d.in_[0], d.wpacked[0], d.ksize[0] = xd.data_ptr(), pk.data_ptr(), 3
        d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, 128
        d.bias, d.alpha, d.beta, d.relu, d.leaky_slope = bd.data_ptr(), 1.0, 0.0, 2, slope
        d.out_f32 = of.data_ptr()
        d.a_mode, d.nacc, d.pair = 0, 2, 1
        plan = C.c_void_p()
        L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
        L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
        torch.cuda.synchronize()
        lib.sr_conv_plan_destroy(plan)
        lin = oracle_conv([x], [w], bias)
        want = np.where(lin >= 0, lin, slope * lin)
        got = of.cpu().numpy()
        # the epilogue stages the accumulator as bf16 before bias/activation: half-ulp = 2^-9 relative
        assert np.abs(got - want).max() <= 2.0 ** -8 * np.abs(lin).max() + 1e-6
        neg = lin < -0.05
        assert neg.any() and np.all(got[neg] < 0) and np.allclose(got[neg] / lin[neg], slope, rtol=0.1)


@pytest.mark.parametrize("shape", [(1, 1, 1, 8), (2, 5, 7, 128), (1, 3, 4, 16)])
def test_bilinear2_matches_tf1_legacy(lib, shape):
    from oracle import other_models as om
    from sr100 import _lib as L
    rng = np.random.default_rng(shape[2])
    x = rng.normal(0, 1, size=shape).astype(np.float32)
    NB, H, W, Cc = shape
    xd = torch.from_numpy(x).cuda()
    of = torch.empty(NB, 2 * H, 2 * W, Cc, device="cuda")
    ob = torch.empty(NB, 2 * H, 2 * W, Cc, device="cuda", dtype=torch.bfloat16)
    L.check(lib.sr_bilinear2_fwd(L.ptr(xd), 0, NB, H, W, Cc, L.ptr(ob), L.ptr(of), L.stream_ptr()))
    want = om.bilinear_tf1(torch.from_numpy(x).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).numpy()
    assert np.array_equal(of.cpu().numpy(), want)                 # t in {0, 0.5}: same three fp32 lerps
    assert np.array_equal(ob.float().cpu().numpy(), bf16_round(want))


@pytest.mark.parametrize("shape", [(2, 5, 7, 128), (1, 1, 3, 128), (1, 9, 4, 256)])
def test_bilinear2_adjoint_matches_autograd(lib, shape):
    """sr_bilinear2_bwd (training of Difvdsr4: the adjoint of Lambda(resize2bil), models.py:932-940, 1034, 1041)
    against torch autograd through the oracle's TF1-legacy bilinear."""
    from oracle import other_models as om
    from sr100 import _lib as L
    rng = np.random.default_rng(2)
    NB, H, W, Cc = shape
    x = torch.from_numpy(rng.standard_normal((NB, Cc, H, W))).requires_grad_(True)
    g = rng.standard_normal((NB, 2 * H, 2 * W, Cc)).astype(np.float32)
    y = om.bilinear_tf1(x, 2)
    (want,) = torch.autograd.grad(y, x, torch.from_numpy(g).double().permute(0, 3, 1, 2))
    gd = torch.from_numpy(g).cuda()
    gin = torch.full((NB, H, W, Cc), float("nan"), device="cuda")
    L.check(lib.sr_bilinear2_bwd(L.ptr(gd), NB, H, W, Cc, L.ptr(gin), L.stream_ptr()))
    assert np.abs(gin.cpu().numpy() - want.permute(0, 2, 3, 1).numpy()).max() <= 1e-5


def _check(arch, specs_fn, fwd, shape, gain, tol=2e-2):
    from oracle import other_models as om
    from sr100.planenet import PlaneNet
    specs = specs_fn()
    weights = om.init_weights(specs, seed=7, bias_scale=0.02, gain=gain)
    tail = specs[-1][0]
    w, b = weights[tail]
    weights[tail] = (w * 3.0, b + 0.3)                            # lift the ReLU'd output off zero
    rng = np.random.default_rng(5)
    x = rng.random(shape).astype(np.float32)
    eng = PlaneNet(arch, weights)
    got = eng.forward_device(torch.from_numpy(x).cuda()).cpu().numpy()
    want = fwd(weights, x)
    s = eng.scale
    assert got.shape == want.shape == (shape[0], s * shape[1], s * shape[2], 3)
    assert want.max() > 0.2 and (want > 0).mean() > 0.3           # a live output, not a dead ReLU
    err = np.abs(got - want).max()
    assert err <= tol, "max-abs error %.4g" % err
    # replay (CUDA graph after the first eager run) gives the same bits
    again = eng.forward_device(torch.from_numpy(x).cuda()).cpu().numpy()
    third = eng.forward_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, again) and np.array_equal(got, third)
    return eng, weights


def test_difvdsr4_forward_matches_oracle():
    from oracle import other_models as om
    _check("difvdsr4", om.difvdsr4_specs, om.forward_difvdsr4, (2, 10, 12, 3), gain=1.0)


def test_difvdsr4_single_image_odd_shape():
    from oracle import other_models as om
    _check("difvdsr4", om.difvdsr4_specs, om.forward_difvdsr4, (1, 7, 5, 3), gain=1.0)


def test_difvdsr_forward_matches_oracle():
    from oracle import other_models as om
    _check("difvdsr", om.difvdsr_specs, om.forward_difvdsr, (2, 14, 18, 3), gain=1.0)


def test_difvdsr_skips_the_zero_half_of_its_second_plane(monkeypatch):
    """192 channels = a full plane + a plane with 64 real channels: sr_conv_desc.cin_valid = 64 drops the two all-zero
    K chunks of that plane from every launch (K = 192 instead of 256).  Zeros add nothing to an fp32 accumulator, so
    the outputs equal the padded launches bit for bit -- single image (single-CTA kernel) and a batch (CTA pairs)."""
    from oracle import other_models as om
    from sr100.planenet import PlaneNet
    weights = om.init_weights(om.difvdsr_specs(), seed=11, bias_scale=0.02, gain=1.0)
    rng = np.random.default_rng(6)
    for shape in ((1, 20, 12, 3), (3, 14, 18, 3)):
        x = torch.from_numpy(rng.random(shape).astype(np.float32)).cuda()
        eng = PlaneNet("difvdsr", weights)
        assert eng.skip_zero_k
        got = eng.forward_device(x)
        flops_skip = eng.net(*shape[:3]).conv_flops
        monkeypatch.setenv("SR100_SKIP_ZERO_K", "0")
        ref = PlaneNet("difvdsr", weights)
        monkeypatch.delenv("SR100_SKIP_ZERO_K")
        assert not ref.skip_zero_k
        want = ref.forward_device(x)
        assert torch.equal(got, want)
        assert flops_skip < 0.80 * ref.net(*shape[:3]).conv_flops      # 192/256 of the K work


def test_model_constructors_predict_weights_and_tiling(tmp_path, monkeypatch):
    """models.Difvdsr4 / models.Difvdsr behind the reference API: create_model, predict, get/set_weights, .h5 round
    trip (Difvdsr always loads its file, models.py:1322), upscaleStepPatch through the generic tiler."""
    import models
    from oracle import other_models as om
    from oracle import tiling as ot
    from PIL import Image
    rng = np.random.default_rng(3)
    # Difvdsr4
    m4 = models.Difvdsr4(1)
    model = m4.create_model(8, 8)
    assert model.output_shape == (None, 32, 32, 3) and len(model.layers) == 66
    assert model.count_params() == sum(k * k * ci * co + co for _, k, ci, co in om.difvdsr4_specs())
    w4 = om.init_weights(om.difvdsr4_specs(), seed=11, bias_scale=0.02)
    t = om.difvdsr4_specs()[-1][0]
    w4[t] = (w4[t][0] * 3.0, w4[t][1] + 0.3)
    model.engine.set_weights_dict(w4)
    x = rng.random((1, 8, 8, 3)).astype(np.float32)
    y = model.predict(x)
    assert np.abs(y - om.forward_difvdsr4(w4, x)).max() <= 2e-2
    wfile = str(tmp_path / "w4.h5")
    model.save_weights(wfile)
    m4b = models.Difvdsr4(1)
    monkeypatch.setenv("SR100_WEIGHTS", wfile)
    model_b = m4b.create_model(8, 8, load_weights=True)
    assert np.array_equal(model_b.predict(x), y)
    assert np.isfinite(model.train_on_batch(x, rng.random((1, 32, 32, 3)).astype(np.float32)))   # fit's step (planetrain)
    # tiled CLI path (upscaleStepPatch) through the generic gather / predict / stitch, against the oracle's tiling
    img = rng.integers(0, 256, size=(40, 50, 3)).astype(np.uint8)
    path = str(tmp_path / "im.png")
    Image.fromarray(img).save(path)
    canvas = m4b.upscaleStepPatch(path, return_image=True, patch_size=32, scalemulti=4, verbose=False)
    _, want = ot.upscale_step_patch(img, lambda p: om.forward_difvdsr4(w4, p), 32, 64, 4)
    assert canvas.shape[0] >= 160 and canvas.shape[1] >= 200
    d = np.abs(canvas[:160, :200].astype(int) - want.astype(int))
    assert d.max() <= 6 and (d > 1).mean() < 0.02                  # uint8 after x255 truncation of a 2e-2 float path
    # Difvdsr: cannot be built without its weight file, exactly like the reference
    monkeypatch.setenv("SR100_WEIGHTS", str(tmp_path / "missing.h5"))
    with pytest.raises(OSError):
        models.Difvdsr(1).create_model(8, 8)
    from sr100 import h5lite
    wd = om.init_weights(om.difvdsr_specs(), seed=12, bias_scale=0.02)
    t = om.difvdsr_specs()[-1][0]
    wd[t] = (wd[t][0] * 3.0, wd[t][1] + 0.3)
    dfile = str(tmp_path / "wd.h5")
    h5lite.save_keras_weights(dfile, wd, order=[s[0] for s in om.difvdsr_specs()])
    monkeypatch.setenv("SR100_WEIGHTS", dfile)
    md = models.Difvdsr(1)
    modeld = md.create_model(12, 10)
    assert modeld.output_shape == (None, 10, 12, 3) and len(modeld.layers) == 130   # width-major, models.py:121
    xd = rng.random((2, 10, 12, 3)).astype(np.float32)
    assert np.abs(modeld.predict(xd) - om.forward_difvdsr(wd, xd)).max() <= 2e-2


def test_evaluate_mirrors_reference_validation_loop(tmp_path, monkeypatch, capsys):
    """models._evaluate (models.py:1519-1622) on a same-size model (Difvdsr): the network is fed the byte-scaled 0..255
    image, like the reference; PSNR printed per image, prediction written to val_predict/.  A x4 model fails in psnr()
    with the reference's own assertion."""
    import os
    import models
    from PIL import Image
    from oracle import other_models as om
    from oracle import pil_resample as pr
    from sr100 import h5lite
    rng = np.random.default_rng(21)
    val = str(tmp_path / "val") + "/"
    imgs = {}
    for sub, names in (("set5", ["a.png", "b.png"]), ("set14", ["c.png"])):
        os.makedirs(val + sub)
        for n in names:
            im = rng.integers(0, 256, size=(10, 12, 3), dtype=np.uint8)
            Image.fromarray(im).save(val + sub + "/" + n)
            imgs[n] = im
    wd = om.init_weights(om.difvdsr_specs(), seed=12, bias_scale=0.02)
    t = om.difvdsr_specs()[-1][0]
    wd[t] = (wd[t][0] * 0.01, wd[t][1] + 0.3)
    dfile = str(tmp_path / "wd.h5")
    h5lite.save_keras_weights(dfile, wd, order=[s[0] for s in om.difvdsr_specs()])
    monkeypatch.setenv("SR100_WEIGHTS", dfile)
    monkeypatch.chdir(tmp_path)
    md = models.Difvdsr(1)
    md.evaluate(val)
    out = capsys.readouterr().out
    assert out.count("Validated image") == 3 and "Average PRNS value of validation images" in out
    for n, im in imgs.items():
        x = pr.imresize_bicubic(pr.imresize_bicubic(im.astype(np.float32) / 255., (10, 12)), (10, 12))
        want = om.forward_difvdsr(wd, x[None].astype(np.float32))[0]
        psnr = models.psnr(im.astype(np.float32) / 255., np.clip(want, 0, 255) / 255)
        line = [l for l in out.splitlines() if "Validated image : %s" % n in l][0]
        assert abs(float(line.split("PSNR value :")[1]) - psnr) < 0.05
        got = np.asarray(Image.open(str(tmp_path / "val_predict" / ("Image ScaleGen_%s_generated.png" % n[:-4]))))
        assert got.shape == (10, 12, 3)
    m4 = models.Difvdsr4(1)
    monkeypatch.setenv("SR100_WEIGHTS", str(tmp_path / "none.h5"))
    with pytest.raises(AssertionError):
        m4.create_model(8, 8)            # weights not loaded: plain random init
        m4.create_model = lambda *a, **k: m4.model
        m4.evaluate(val)                 # x4 output vs same-size target: psnr() asserts equal shapes


# ------------------------------------------------------------------------------------------------------------------
# training of the two older graphs (sr100.planetrain; reference: fit of Difvdsr4 / Difvdsr, models.py:1079-1080,
# 1332-1333, on graphs compiled with loss='mse', Adam(1e-4, 0.9))

def _train_case(arch, shape, seed=21):
    from oracle import other_models as om
    specs = (om.difvdsr4_specs if arch == "difvdsr4" else om.difvdsr_specs)()
    weights = om.init_weights(specs, seed=seed, bias_scale=0.02, gain=1.0)
    tail = specs[-1][0]
    w, b = weights[tail]
    weights[tail] = (w * 3.0, b + 0.3)                            # a live ReLU'd output
    s = 4 if arch == "difvdsr4" else 1
    rng = np.random.default_rng(seed + 1)
    x = rng.random(shape + (3,)).astype(np.float32)
    y = rng.random((shape[0], s * shape[1], s * shape[2], 3)).astype(np.float32)
    return specs, weights, x, y


@pytest.mark.parametrize("arch,shape", [("difvdsr4", (2, 6, 10)), ("difvdsr", (2, 12, 14)), ("difvdsr4", (1, 5, 4))])
def test_plane_trainer_gradients_match_oracle_autograd(arch, shape):
    """Forward with saved activations + the whole backward of both graphs (two-plane input-gradient convs with ReLU /
    LeakyReLU masks and residual algebra, four 128 x 128 filter-gradient launches per layer, bias sums, bilinear x2
    adjoint, K = 27 tail, 1x1 head) against torch autograd on the fp64 oracle graph: the loss within 1 %, every
    layer's kernel gradient within 10 % relative L2 / cosine >= 0.995 (bf16 operands through up to 130 layers), the
    frozen Difvdsr head without gradient, and nothing in the zero padding of the plane blocks."""
    from oracle import other_models as om
    from sr100.planenet import PlaneNet
    from sr100.planetrain import PlaneTrainer
    specs, weights, x, y = _train_case(arch, shape)
    net = PlaneNet(arch, weights)
    tr = PlaneTrainer(net)
    g = tr.graph(*shape)
    tr._load(g, x, y)
    tr.forward_backward_device(g)
    torch.cuda.synchronize()
    loss = float(g.loss_sum.item()) / g.n_local
    want_loss, want, pred = om.loss_and_grads(arch, weights, x, y)
    assert np.abs(g.out.cpu().numpy() - pred).max() <= 2e-2
    assert abs(loss - want_loss) <= 1e-2 * want_loss, (loss, want_loss)
    got = tr.grads_dict()
    worst = (0.0, None)
    for name, k, cin, cout in specs:
        gw, gb = got[name]
        ww, wb = want[name]
        assert gw.shape == ww.shape and gb.shape == wb.shape
        if arch == "difvdsr" and name == "level1":
            assert not gw.any() and not gb.any()                  # trainable=False (models.py:1304)
            continue
        nw = float(np.linalg.norm(ww))
        assert nw > 0, name
        rel = float(np.linalg.norm(gw - ww)) / nw
        cos = float((gw * ww).sum() / (np.linalg.norm(gw) * nw))
        relb = float(np.linalg.norm(gb - wb)) / max(float(np.linalg.norm(wb)), 1e-30)
        worst = max(worst, (rel, name))
        assert rel <= 0.1 and cos >= 0.995 and relb <= 0.1, (name, rel, cos, relb)
    # the assembled gradients are everything: the padded rows / columns of the 192-channel blocks hold zeros
    real = sum(gw.size + gb.size for gw, gb in got.values())
    assert int((tr.grads != 0).sum()) <= real
    print("worst layer: rel L2 %.4f (%s)" % worst)


@pytest.mark.parametrize("arch,shape", [("difvdsr4", (2, 6, 6)), ("difvdsr", (2, 10, 10))])
def test_plane_trainer_step_follows_the_oracle_and_reduces_the_loss(arch, shape):
    """One Adam step lands where Keras-Adam on the oracle gradients lands (the first step moves every weight with a
    non-zero gradient by ~lr), the padding of the plane blocks stays zero, the Keras-shaped facade trains
    (models.<class>.create_model().train_on_batch) and the loss goes down."""
    import models
    from oracle import other_models as om
    from sr100.planenet import PlaneNet
    from sr100.planetrain import PlaneTrainer
    specs, weights, x, y = _train_case(arch, shape, seed=33)
    net = PlaneNet(arch, weights)
    tr = PlaneTrainer(net, lr=1e-4)
    l0 = tr.train_on_batch(x, y)
    torch.cuda.synchronize()
    want_loss, want, _ = om.loss_and_grads(arch, weights, x, y)
    assert abs(l0 - want_loss) <= 1e-2 * want_loss
    after = net.get_weights_dict()
    checked = 0
    for name, k, cin, cout in specs:
        if arch == "difvdsr" and name == "level1":
            assert np.array_equal(after[name][0], weights[name][0])
            continue
        gw = want[name][0]
        big = np.abs(gw) > 0.05 * np.abs(gw).max()                # clear of sign flips from bf16 noise
        step = after[name][0] - weights[name][0]
        assert np.all(np.sign(step[big]) == -np.sign(gw[big])), name
        # t = 1: m = (1-b1) g, v = (1-b2) g^2  =>  step = -lr * g / (|g| + eps / sqrt(1-b2))
        expect = -1e-4 * gw[big] / (np.abs(gw[big]) + 1e-7 / np.sqrt(1.0 - 0.999))
        assert np.abs(step[big] - expect).max() <= 0.15 * np.abs(expect).max() + 1e-7, name
        assert np.abs(step[big]).max() <= 1e-4 * (1 + 1e-3)
        checked += int(big.sum())
    assert checked > 1000
    # round trip through the Keras-shaped weights keeps the arena's padding at zero
    arena = net.param_arena.clone()
    net.set_weights_dict(after)
    assert torch.equal(arena, net.param_arena)
    cls = models.Difvdsr4 if arch == "difvdsr4" else models.Difvdsr
    m = cls(1)
    m._engine = PlaneNet(arch, weights)
    m.force_load = False
    model = m.create_model(shape[1], shape[2])
    # (the reference's 1e-4 overshoots on this 2-image batch with the lifted tail: sign-like first Adam steps)
    model.compile(optimizer=models._Adam(1e-5, 0.9), loss='mse', metrics=['accuracy'])
    losses = [model.train_on_batch(x, y) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], losses
