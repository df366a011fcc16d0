"""-m gpu: whole DifvdsrDouble forward (engine through the Keras-like Model facade) against the CPU oracle with
identical random-init weights.  Tolerance from BASELINE.json north_star: max-abs <= 2e-2 on [0,1] outputs for the
bf16 path and <= 0.01 dB PSNR."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _smooth_images(rng, n, h, w):
    from scipy.ndimage import uniform_filter
    img = rng.integers(0, 256, size=(n, h + 4, w + 4, 3)).astype(np.float32)
    img = uniform_filter(img, size=(1, 5, 5, 1))[:, 2:-2, 2:-2]
    return (img / 255.0).astype(np.float32)


@pytest.fixture(scope="module")
def weights():
    from oracle import model as om
    return om.init_weights(1234, bias_scale=0.01)


@pytest.fixture(scope="module")
def dmodel(weights):
    import models
    m = models.DifvdsrDouble(1)
    model = m.create_model(24, 24)
    model.engine.set_weights_dict(weights)
    return m, model


def _psnr(a, b):
    return -10 * np.log10(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))


@pytest.mark.parametrize("shape", [(2, 24, 24), (1, 48, 40), (1, 96, 96), (1, 128, 128)])   # last: BASELINE configs[0]
def test_forward_matches_oracle(dmodel, weights, shape):
    from oracle import model as om
    _, model = dmodel
    rng = np.random.default_rng(shape[1])
    x = _smooth_images(rng, *shape)
    got = model.predict(x, batch_size=1)
    want = om.forward_numpy(weights, x)
    assert got.shape == want.shape == (shape[0], 4 * shape[1], 4 * shape[2], 3)
    assert got.dtype == np.float32 and got.min() >= 0
    err = np.abs(got - want).max()
    assert err <= 2e-2, err                      # north_star bf16 tolerance
    assert err <= 2e-3, err                      # what the fp32-stream design actually achieves (outputs ~0.05)
    # PSNR against a synthetic HR target must agree within 0.01 dB
    target = np.clip(want + rng.normal(0, 0.02, size=want.shape), 0, 1)
    assert abs(_psnr(got, target) - _psnr(want, target)) <= 0.01


def test_zero_bias_default_and_stream_variants(weights):
    """All three residual-stream modes stay inside the tolerance; bf16 stream is the loosest."""
    from oracle import model as om
    from sr100.engine import Engine
    rng = np.random.default_rng(11)
    x = _smooth_images(rng, 1, 32, 32)
    want = om.forward_numpy(weights, x)
    errs = {}
    for stream in ("bf16", "lr32", "fp32"):
        eng = Engine(weights, stream=stream)
        got = eng.forward_device(torch.from_numpy(x).cuda()).cpu().numpy()
        errs[stream] = float(np.abs(got - want).max())
        assert errs[stream] <= 2e-2
    assert errs["fp32"] <= errs["bf16"] + 1e-6


def test_a_modes_agree(weights):
    from sr100.engine import Engine
    rng = np.random.default_rng(12)
    x = torch.from_numpy(_smooth_images(rng, 2, 24, 24)).cuda()
    y0 = Engine(weights, a_mode=0, nacc=4, pair=0).forward_device(x).cpu().numpy()
    y1 = Engine(weights, a_mode=1, nacc=4, pair=0).forward_device(x).cpu().numpy()
    y2 = Engine(weights, a_mode=0, nacc=2, pair=0).forward_device(x).cpu().numpy()
    y3 = Engine(weights, a_mode=0, nacc=2, pair=1).forward_device(x).cpu().numpy()
    assert np.abs(y0 - y1).max() < 1e-5 and np.abs(y0 - y2).max() < 1e-5 and np.abs(y0 - y3).max() < 1e-5


def test_sub_batching_is_invisible(weights):
    from sr100.engine import Engine
    rng = np.random.default_rng(13)
    x = torch.from_numpy(_smooth_images(rng, 5, 16, 16)).cuda()
    a = Engine(weights, max_pixels=16 * 16 * 2).forward_device(x).cpu().numpy()     # 2 + 2 + 1
    b = Engine(weights, max_pixels=16 * 16 * 64).forward_device(x).cpu().numpy()
    assert np.array_equal(a, b)


def models_mod():
    import models
    return models


def test_weights_roundtrip_and_npz(dmodel, weights, tmp_path):
    _, model = dmodel
    ws = model.get_weights()
    assert len(ws) == 172 and ws[0].shape == (1, 1, 3, 128) and ws[-2].shape == (3, 3, 128, 3)
    assert sum(w.size for w in ws) == 21838211 == model.count_params()
    assert np.array_equal(ws[2], weights["conv2d_1"][0])
    p = model.save_weights(str(tmp_path / "w.h5"))
    with open(p, "rb") as f:
        assert f.read(8) == b"\x89HDF\r\n\x1a\n"            # a real HDF5 file (sr100.h5lite), Keras 2 layout
    model.save_weights(str(tmp_path / "w.npz"))
    mod3 = models_mod().DifvdsrDouble(1).create_model(24, 24)
    mod3.load_weights(str(tmp_path / "w.npz"))
    assert all(np.array_equal(a, b) for a, b in zip(mod3.get_weights(), ws))
    import models
    m2 = models.DifvdsrDouble(1)
    mod2 = m2.create_model(24, 24)
    mod2.load_weights(str(tmp_path / "w.h5"))
    assert all(np.array_equal(a, b) for a, b in zip(mod2.get_weights(), ws))
    with pytest.raises(OSError):
        mod2.load_weights(str(tmp_path / "missing.h5"))
    with pytest.raises(ValueError):
        model.set_weights(ws[:-1])
    # get_weights() follows Keras' model.layers order (depth-sorted): level1, then a3, c5, b5, d3 of the first block
    assert [l.name for l in model.layers[:5]] == ["level1", "conv2d_1", "conv2d_3", "conv2d_2", "conv2d_4"]
    assert np.array_equal(ws[4], weights["conv2d_3"][0]) and np.array_equal(ws[6], weights["conv2d_2"][0])


def test_load_weights_is_positional_like_keras(dmodel, weights, tmp_path):
    """Keras' load_weights (by_name=False, what models.py:1218 calls) zips the file's weighted layers in
    `layer_names` order with model.layers and never looks at names:
    (1) a file whose auto-names are offset (second create_model without clear_session: conv2d_86...) loads;
    (2) a file written in CREATION order (conv2d_1, conv2d_2, conv2d_3, ...) loads without a shape error and with the
        two 5x5 kernels of every 5/3 block swapped -- exactly what real Keras would do with such a file;
    (3) by_name=True matches group names instead."""
    from sr100 import h5lite
    from sr100 import keras_graph as kg
    from sr100.engine import layer_specs
    _, model = dmodel
    model.engine.set_weights_dict(weights)
    # (1) offset names, Keras order
    layers = kg.difvdsr_double_layers()
    ren = {}
    for n, has_w in layers:
        ren[n] = "conv2d_%d" % (int(n.split("_")[1]) + 85) if n.startswith("conv2d_") else n
    p1 = str(tmp_path / "offset.h5")
    h5lite.save_keras_weights(p1, {ren[n]: weights[n] for n, w in layers if w}, layers=[(ren[n], w) for n, w in layers])
    m2 = models_mod().DifvdsrDouble(1).create_model(24, 24)
    m2.load_weights(p1)
    got = m2.engine.get_weights_dict()
    assert all(np.array_equal(got[n][0], weights[n][0]) and np.array_equal(got[n][1], weights[n][1]) for n in weights)
    with pytest.raises(KeyError):
        m2.load_weights(p1, by_name=True)
    # (2) creation-order file: positional load swaps conv2d_2 <-> conv2d_3 (both 5x5x128x128) in every 5/3 block
    p2 = str(tmp_path / "creation_order.h5")
    h5lite.save_keras_weights(p2, weights, order=[s[0] for s in layer_specs()])
    m2.load_weights(p2)
    got = m2.engine.get_weights_dict()
    assert np.array_equal(got["conv2d_2"][0], weights["conv2d_3"][0])
    assert np.array_equal(got["conv2d_3"][0], weights["conv2d_2"][0])
    assert np.array_equal(got["conv2d_1"][0], weights["conv2d_1"][0])
    assert np.array_equal(got["conv2d_65"][0], weights["conv2d_65"][0])       # light blocks: creation == Keras order
    # (3) by name the same file is read back as written
    m2.load_weights(p2, by_name=True)
    got = m2.engine.get_weights_dict()
    assert all(np.array_equal(got[n][0], weights[n][0]) for n in weights)


def test_predict_rejects_bad_shapes(dmodel):
    _, model = dmodel
    with pytest.raises(ValueError):
        model.predict(np.zeros((24, 24, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        model.predict(np.zeros((1, 24, 24, 4), dtype=np.float32))


def test_cuda_graph_replay_matches_eager(weights):
    """The captured launch sequence reproduces the eager one bit for bit, also after the weights change in place."""
    from sr100.engine import Engine
    rng = np.random.default_rng(13)
    x = torch.from_numpy(_smooth_images(rng, 2, 24, 24)).cuda()
    eager = Engine(weights, use_graphs=False)
    graphed = Engine(weights, use_graphs=True)
    want = eager.forward_device(x).cpu().numpy()
    outs = [graphed.forward_device(x).cpu().numpy() for _ in range(3)]   # eager, capture, replay
    assert graphed.graph_ready()
    for o in outs:
        assert np.array_equal(o, want)
    w2 = {k: (v[0] * 0.5, v[1] + 0.01) for k, v in weights.items()}
    eager.set_weights_dict(w2)
    graphed.set_weights_dict(w2)
    assert np.array_equal(graphed.forward_device(x).cpu().numpy(), eager.forward_device(x).cpu().numpy())
