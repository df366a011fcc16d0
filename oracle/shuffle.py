"""numpy emulation of the reference's three sub-pixel shuffles (TEST INFRASTRUCTURE, oracle/__init__.py).
Each function replays the reference's tensor program step by step (reshape / transpose / split / concat),
so the index map is derived from the reference code, not assumed.  PARITY UNPINNED (TF / Theano absent)."""
import itertools

import numpy as np


def phase_shift_subpixel(I, r):
    """keras_subpixel.Subpixel._phase_shift (keras_subpixel.py:64-84), NHWC."""
    bsize, a, b, c = I.shape
    X = I.reshape(bsize, a, b, c // (r * r), r, r)
    X = X.transpose(0, 1, 2, 5, 4, 3)                                  # bsize, a, b, r, r, c/(r*r)
    X = np.concatenate([X[:, i] for i in range(a)], axis=2)            # bsize, b, a*r, r, c/(r*r)
    X = np.concatenate([X[:, i] for i in range(b)], axis=2)            # bsize, a*r, b*r, c/(r*r)
    return X


def _phase_shift_tf(I, r):
    """advanced.depth_to_scale_tf._phase_shift (advanced.py:111-122), one colour group, NHWC with r*r channels."""
    bsize, a, b, c = I.shape
    X = I.reshape(bsize, a, b, r, r)
    X = X.transpose(0, 1, 2, 4, 3)
    X = np.split(X, a, axis=1)                                         # a x [bsize, 1, b, r, r]
    X = np.concatenate([x.squeeze(axis=1) for x in X], axis=2)         # bsize, b, a*r, r
    X = np.split(X, b, axis=1)                                         # b x [bsize, 1, a*r, r]
    X = np.concatenate([x.squeeze(axis=1) for x in X], axis=2)         # bsize, a*r, b*r
    return X.reshape(bsize, a * r, b * r, 1)


def depth_to_scale_tf(x, r, channels):
    """advanced.py:104-129: 3 colour groups (hard-coded split into 3, advanced.py:125)."""
    if channels > 1:
        groups = np.split(x, 3, axis=3)
        return np.concatenate([_phase_shift_tf(g, r) for g in groups], axis=3)
    return _phase_shift_tf(x, r)


def depth_to_scale_th(x_nchw, r, channels):
    """advanced.py:87-100: out[:, :, y::r, x::r] += input[:, r*y + x :: r*r]."""
    b, k, row, col = x_nchw.shape
    out = np.zeros((b, channels, row * r, col * r), dtype=x_nchw.dtype)
    for y, x in itertools.product(range(r), repeat=2):
        out[:, :, y::r, x::r] += x_nchw[:, r * y + x::r * r, :, :]
    return out


def depth_to_space_tf(x, r):
    """tf.depth_to_space (NHWC, DCR): out[b, h*r+i, w*r+j, c] = in[b, h, w, (i*r + j)*C + c]."""
    b, h, w, c = x.shape
    C = c // (r * r)
    X = x.reshape(b, h, w, r, r, C).transpose(0, 1, 3, 2, 4, 5)
    return X.reshape(b, h * r, w * r, C)


def conv2d_nhwc(x, w_hwio, bias=None, same=True, relu=False):
    """Plain fp64 cross-correlation (Keras Conv2D semantics) for small shapes."""
    n, h, wd, cin = x.shape
    k = w_hwio.shape[0]
    cout = w_hwio.shape[3]
    p = (k - 1) // 2 if same else 0
    xp = np.pad(x.astype(np.float64), ((0, 0), (p, p), (p, p), (0, 0)))
    oh, ow = (h, wd) if same else (h - k + 1, wd - k + 1)
    out = np.zeros((n, oh, ow, cout))
    for ky in range(k):
        for kx in range(k):
            out += np.tensordot(xp[:, ky:ky + oh, kx:kx + ow, :], w_hwio[ky, kx].astype(np.float64), axes=([3], [0]))
    if bias is not None:
        out += bias
    if relu:
        out = np.maximum(out, 0)
    return out
