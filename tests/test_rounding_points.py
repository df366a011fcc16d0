"""CPU: which of the engine's rounding points matter (oracle/emu_bf16.py against the fp64 oracle graph).

The 128-wide bf16 epilogues round the fp32 accumulator to bf16 before bias / alpha / residual, also when the
destination is the fp32 residual stream (csrc/conv_tc.cu, epilogue_staged_acc).  Measured here: carrying the
accumulator unrounded through the epilogue does not move the forward error -- the bf16 rounding of the conv OPERANDS
dominates -- so the staging tile stays bf16 (half the shared memory and store traffic of an fp32 tile)."""
import numpy as np
import torch


def test_epilogue_accumulator_rounding_is_not_what_bounds_the_forward_error():
    from scipy.ndimage import uniform_filter
    from oracle import emu_bf16 as emu
    from oracle import model as om
    w = om.init_weights(1234, bias_scale=0.01)
    rng = np.random.default_rng(0)
    x = (uniform_filter(rng.integers(0, 256, size=(1, 20, 20, 3)).astype(np.float32), size=(1, 5, 5, 1)) / 255.0)
    x = x.astype(np.float32)
    ref = om.forward_numpy(w, x, dtype=torch.float64)
    err = {}
    for acc_round in (True, False):
        with torch.no_grad():
            y = emu.DifvdsrDoubleBf16Emu(w, acc_round=acc_round)(torch.from_numpy(x)).numpy()
        d = np.abs(y - ref)
        err[acc_round] = (float(d.max()), float(np.sqrt((d ** 2).mean())))
    print("forward error vs fp64 graph (max-abs, rms): rounded accumulator %s, unrounded %s" % (err[True], err[False]))
    assert err[True][0] <= 2e-3                                   # the bf16 engine's tolerance
    assert abs(err[True][1] - err[False][1]) <= 0.1 * err[False][1]      # rms error: same within 10 %
    assert err[True][0] <= 1.25 * err[False][0]
