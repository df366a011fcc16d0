"""-m gpu: sub-pixel layers (keras_subpixel.Subpixel, advanced.SubPixelUpscaling / SubpixelConv2D / depth_to_scale_*)
against numpy replays of the reference tensor programs (oracle/shuffle.py).  Shuffles are bit-exact."""
import numpy as np
import pytest

from oracle import shuffle as osh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("r,C,H,W", [(2, 3, 5, 7), (4, 3, 6, 4), (3, 1, 4, 4), (4, 8, 9, 3)])
def test_depth_to_space_orderings_bit_exact(r, C, H, W):
    import advanced
    rng = np.random.default_rng(r * 100 + C)
    x = rng.standard_normal((2, H, W, C * r * r)).astype(np.float32)
    d = advanced.SubpixelConv2D((None, H, W, C * r * r), scale=r)
    assert np.array_equal(d(x), osh.depth_to_space_tf(x, r))
    assert d.compute_output_shape((None, H, W, C * r * r)) == (None, H * r, W * r, C)
    th = advanced.depth_to_scale_th(np.ascontiguousarray(x.transpose(0, 3, 1, 2)), r, C)
    assert np.array_equal(th, osh.depth_to_scale_th(x.transpose(0, 3, 1, 2), r, C))
    if C in (1, 3):
        up = advanced.SubPixelUpscaling(r, C)
        assert np.array_equal(up(x), osh.depth_to_scale_tf(x, r, C))
        assert up.get_output_shape_for((2, H, W, C * r * r)) == (2, H * r, W * r, C)
    from sr100 import ops
    import torch
    assert np.array_equal(ops.depth_to_space(torch.from_numpy(x).cuda(), r, 0).cpu().numpy(),
                          osh.phase_shift_subpixel(x, r))


@pytest.mark.parametrize("padding,act", [("valid", None), ("same", "relu")])
def test_subpixel_layer_fused_conv_shuffle(padding, act):
    """keras_subpixel.py:124-172 smoke-test contract (Subpixel(3,(3,3),2) doubles the resolution) + values."""
    from keras_subpixel import Subpixel
    rng = np.random.default_rng(5)
    x = rng.random((2, 10, 12, 3)).astype(np.float32)
    layer = Subpixel(3, (3, 3), 2, padding=padding, activation=act, seed=3)
    y = layer(x)
    k, b = layer.get_weights()
    assert k.shape == (3, 3, 3, 12) and b.shape == (12,)
    b2 = rng.uniform(-0.1, 0.1, size=12).astype(np.float32)
    layer.set_weights([k, b2])
    y = layer(x)
    conv = osh.conv2d_nhwc(x, k, b2, same=(padding == "same"), relu=(act == "relu"))
    want = osh.phase_shift_subpixel(conv, 2)
    assert y.shape == want.shape == layer.compute_output_shape((2, 10, 12, 3))
    assert np.abs(y - want).max() < 1e-5
    # fused epilogue store == unfused conv followed by the shuffle kernel, bit for bit
    import torch
    from sr100 import ops
    xd, kd, bd = (torch.from_numpy(v).cuda() for v in (x, k, b2))
    unf = ops.depth_to_space(ops.conv2d_direct(xd, kd, bd, same=(padding == "same"), relu=(act == "relu")), 2, 0)
    assert np.array_equal(unf.cpu().numpy(), y)
    cfg = layer.get_config()
    assert cfg["r"] == 2 and cfg["filters"] == 3 and "rank" not in cfg and "dilation_rate" not in cfg
