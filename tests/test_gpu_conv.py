"""-m gpu: the tcgen05 conv, head, and bilinear kernels against the CPU oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from gpu_util import bf16_round, oracle_conv, run_tc_conv

pytestmark = pytest.mark.gpu


def _rand(rng, shape, scale=0.5):
    return bf16_round(rng.standard_normal(shape).astype(np.float32) * scale)


CASES = [
    # name, ks, NB, H, W, cout, relu, alpha, beta, res
    ("k1", (1,), 2, 20, 96, 128, 0, 1.0, 0.0, None),
    ("k3_relu", (3,), 2, 20, 96, 128, 1, 1.0, 0.0, None),
    ("k5", (5,), 2, 20, 96, 128, 0, 1.0, 0.0, None),
    ("dual_53_block_end", (5, 3), 2, 20, 96, 128, 0, 0.1, 0.9, "f32"),
    ("light_block_end_bf16res", (3,), 2, 20, 96, 128, 0, 0.1, 1.0, "bf16"),
    ("tail_cout3", (3,), 2, 20, 96, 3, 1, 1.0, 0.0, None),
    ("ragged_50x33", (3,), 3, 33, 50, 128, 0, 1.0, 0.0, None),
    ("wide_384", (5,), 1, 12, 384, 128, 0, 1.0, 0.0, None),
    ("single_pixel", (5,), 1, 1, 1, 128, 0, 1.0, 0.0, None),
    ("one_row", (3,), 1, 1, 200, 128, 1, 1.0, 0.0, None),
    ("one_col", (5,), 2, 70, 1, 128, 0, 1.0, 0.0, None),
    ("train_patch_48", (5,), 4, 48, 48, 128, 0, 1.0, 0.0, None),
]


@pytest.mark.parametrize("a_mode,nacc,pair", [(0, 4, 0), (1, 4, 0), (0, 2, 0), (0, 4, 1), (0, 2, 1)])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_matches_oracle(lib, case, a_mode, nacc, pair):
    name, ks, NB, H, W, cout, relu, alpha, beta, res_kind = case
    rng = np.random.default_rng(__import__("zlib").crc32(name.encode()))
    xs = [_rand(rng, (NB, H, W, 128)) for _ in ks]
    ws = [rng.standard_normal((k, k, 128, cout)).astype(np.float32) / np.sqrt(k * k * 128) for k in ks]
    bias = rng.standard_normal(cout).astype(np.float32) * 0.1
    res = None
    if res_kind:
        res = rng.standard_normal((NB, H, W, cout)).astype(np.float32)
        if res_kind == "bf16":
            res = bf16_round(res)
    got32, got16 = run_tc_conv(lib, xs, ws, bias, relu, alpha, beta, res, res_kind or "f32", cout, a_mode, nacc, pair)
    want = oracle_conv(xs, ws, bias, relu, alpha, beta, res)
    assert not np.isnan(got32).any(), "some output pixels were never written"
    # identical operands, fp32 accumulation on both sides.  For cout == 128 the kernel stages the raw accumulator
    # in bf16 before bias / alpha / residual, so the fp32 output may differ from the oracle by |alpha| * half a
    # bf16 ulp of the accumulator (plus summation-order noise); the 3-channel tail stays in fp32 throughout.
    raw = np.abs(oracle_conv(xs, ws, None, 0, 1.0, 0.0, None))
    half_ulp = np.where(raw > 0, 2.0 ** (np.floor(np.log2(np.maximum(raw, 1e-30))) - 8), 0.0)
    tol = 2e-4 + (abs(alpha) * half_ulp * 1.01 if cout == 128 else 0.0)
    assert (np.abs(got32 - want) <= tol).all(), float(np.abs(got32 - want).max())
    assert np.abs(got16 - want).max() < 2e-2 * max(1.0, np.abs(want).max())


def test_conv_linearity_at_full_tile_size(lib):
    """Size-independent property at the BASELINE tile size (a batch of 96x96 LR tiles, 5x5 conv, 3-channel tail
    variant so the output stays fp32): conv(a + b) == conv(a) + conv(b).  a, b are multiples of 1/8 so a + b is
    exactly representable in bf16."""
    rng = np.random.default_rng(5)
    a = rng.integers(-16, 17, size=(8, 96, 96, 128)).astype(np.float32) / 8
    b = rng.integers(-16, 17, size=(8, 96, 96, 128)).astype(np.float32) / 8
    assert np.array_equal(bf16_round(a + b), a + b)
    w = rng.standard_normal((5, 5, 128, 3)).astype(np.float32) / np.sqrt(3200)
    ya, _ = run_tc_conv(lib, [a], [w], None, cout=3)
    yb, _ = run_tc_conv(lib, [b], [w], None, cout=3)
    yab, _ = run_tc_conv(lib, [a + b], [w], None, cout=3)
    assert np.abs((ya + yb) - yab).max() < 1e-4
    # and the 128-channel kernel agrees with the tail kernel on a shared output channel
    w128 = np.zeros((5, 5, 128, 128), dtype=np.float32)
    w128[..., :3] = w
    y128, _ = run_tc_conv(lib, [a], [w128], None)
    assert np.abs(y128[..., :3] - ya).max() <= 2e-4 + np.abs(ya).max() * 2.0 ** -8


def test_head1x1_matches_oracle(lib):
    from sr100 import _lib as L
    rng = np.random.default_rng(1)
    x = rng.random((3, 17, 23, 3)).astype(np.float32)
    w = rng.uniform(-0.2, 0.2, size=(1, 1, 3, 128)).astype(np.float32)
    b = rng.uniform(-0.05, 0.05, size=128).astype(np.float32)
    xd, wd, bd = (torch.from_numpy(v).cuda() for v in (x, w, b))
    o16 = torch.empty(3, 17, 23, 128, device="cuda", dtype=torch.bfloat16)
    o32 = torch.empty(3, 17, 23, 128, device="cuda")
    L.check(lib.sr_head1x1_fwd(L.ptr(xd), L.ptr(wd), L.ptr(bd), 3 * 17 * 23, L.ptr(o16), L.ptr(o32), L.stream_ptr()))
    want = np.maximum(x.reshape(-1, 3).astype(np.float64) @ w.reshape(3, 128).astype(np.float64) + b, 0).reshape(3, 17, 23, 128)
    assert np.abs(o32.cpu().numpy() - want).max() < 1e-6
    assert np.array_equal(o16.cpu().float().numpy(), bf16_round(o32.cpu().numpy()))


def test_bilinear4_bit_exact_vs_oracle(lib):
    from oracle import model as om
    from sr100 import ops
    rng = np.random.default_rng(2)
    for shape in [(2, 5, 7, 16), (1, 1, 1, 8), (1, 24, 24, 128), (1, 3, 96, 8)]:
        x = rng.standard_normal(shape).astype(np.float32)
        got = ops.bilinear4(torch.from_numpy(x).cuda()).cpu().numpy()
        want = om.bilinear_x4_tf1(torch.from_numpy(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
        assert got.shape == want.shape
        assert np.array_equal(got, want)          # fp32 lerp with the oracle's operation order: bit-exact
        xb = bf16_round(x)
        gotb = ops.bilinear4(torch.from_numpy(xb).cuda().to(torch.bfloat16), out_dtype=torch.bfloat16).float().cpu().numpy()
        wantb = bf16_round(om.bilinear_x4_tf1(torch.from_numpy(xb).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy())
        assert np.array_equal(gotb, wantb)


def test_bilinear4_adjoint_vs_oracle_autograd(lib):
    from oracle import model as om
    from sr100 import ops
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((2, 6, 5, 8)).astype(np.float32)).requires_grad_(True)
    y = om.bilinear_x4_tf1(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
    y.backward(g)
    got = ops.bilinear4_bwd(g.cuda().contiguous()).cpu().numpy()
    assert np.abs(got - x.grad.numpy()).max() < 1e-5
