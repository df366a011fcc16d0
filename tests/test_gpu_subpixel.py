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


@pytest.mark.parametrize("k,r,C,order", [(3, 4, 3, 0), (5, 2, 8, 1), (3, 2, 32, 2), (1, 4, 8, 0),
                                         (3, 3, 3, 0), (3, 2, 5, 2), (5, 3, 11, 1)])   # r*C % 4 != 0: scalar stores
def test_tensor_core_conv_with_fused_shuffle(k, r, C, order):
    """128-channel bf16 input: the tcgen05 conv stores straight to the depth-to-space position.  Bit-identical to the
    same kernel's unfused output followed by the shuffle kernel, and equal to the oracle conv within fp32 rounding."""
    import torch
    from sr100 import ops
    from gpu_util import bf16_round, run_tc_conv, oracle_conv
    from sr100 import _lib as L
    rng = np.random.default_rng(k * 10 + r)
    NB, H, W, cout = 2, 9, 20, r * r * C
    x = bf16_round(rng.normal(0, 0.5, size=(NB, H, W, 128)))
    w = (rng.normal(0, 1, size=(k, k, 128, cout)) / np.sqrt(k * k * 128)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, size=cout).astype(np.float32)
    xd = torch.from_numpy(x).cuda().to(torch.bfloat16)
    got = ops.conv2d_tc_shuffle(xd, torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), r, order, relu=True)
    assert got.shape == (NB, H * r, W * r, C)
    # unfused: the same tensor-core kernel with the weights zero-padded to 128 outputs, then the shuffle kernel
    w128 = np.zeros((k, k, 128, 128), dtype=np.float32); w128[..., :cout] = w
    b128 = np.zeros(128, dtype=np.float32); b128[:cout] = b
    unf, _ = run_tc_conv(L.require_device(), [x], [w128], b128, relu=1, nacc=2, pair=1)
    # the unfused fp32 output went through the same bf16 staging, so the fused store must match it exactly
    shuf = ops.depth_to_space(torch.from_numpy(np.ascontiguousarray(unf[..., :cout])).cuda(), r, order).cpu().numpy()
    assert np.array_equal(got.cpu().numpy(), shuf)
    want = osh_shuffle(oracle_conv([x], [w], b, relu=1), r, order, C)
    assert np.abs(got.cpu().numpy() - want).max() <= 2.0 ** -8 * max(1.0, np.abs(want).max())


def osh_shuffle(x, r, order, C):
    if order == 0:
        return osh.phase_shift_subpixel(x, r)
    if order == 1:
        return osh.depth_to_scale_th(np.ascontiguousarray(x.transpose(0, 3, 1, 2)), r, C).transpose(0, 2, 3, 1)
    return osh.depth_to_space_tf(x, r)
