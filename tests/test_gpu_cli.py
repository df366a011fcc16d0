"""-m gpu: the `main_dirpath.py <imgpath>` CLI drop-in (main_dirpath.py:1-55) on a temporary directory."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_writes_x4_images(tmp_path):
    from PIL import Image
    from oracle import model as om
    rng = np.random.default_rng(0)
    d = tmp_path / "imgs"
    d.mkdir()
    Image.fromarray(rng.integers(0, 256, size=(33, 50, 3)).astype(np.uint8)).save(str(d / "a.png"))
    Image.fromarray(rng.integers(0, 256, size=(20, 20, 3)).astype(np.uint8)).save(str(d / "b.bmp"))
    weights = om.init_weights(5)
    wfile = str(tmp_path / "w.npz")
    np.savez(wfile, **{k + "/kernel:0": v[0] for k, v in weights.items()},
             **{k + "/bias:0": v[1] for k, v in weights.items()})
    env = dict(os.environ, SR100_WEIGHTS=wfile)
    cli = os.path.join(ROOT, "image-enhance-keras_b200", "main_dirpath.py")
    p = subprocess.run([sys.executable, cli, str(d) + "/", "--suffix", "scaled"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert Image.open(str(d / "a_scaled(1x).png")).size == (200, 132)
    assert Image.open(str(d / "b_scaled(1x).bmp")).size == (80, 80)
    # bad --model is rejected by the same assert as the reference (main_dirpath.py:27)
    p = subprocess.run([sys.executable, cli, str(d) + "/", "--model", "sr"], env=env, capture_output=True, text=True)
    assert p.returncode != 0 and "Model type must be" in p.stderr


def test_upscale_fast_mode_whole_image(tmp_path):
    """models.upscale(mode='fast') (models.py:606-852): whole image through the network, `_A` side file, uint8 output
    = clip(predict * 255) truncated, against the oracle forward."""
    from PIL import Image
    import models
    from oracle import model as om
    rng = np.random.default_rng(3)
    from scipy.ndimage import uniform_filter
    img = uniform_filter(rng.integers(0, 256, size=(36, 52, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)
    p = str(tmp_path / "pic.png")
    Image.fromarray(img).save(p)
    weights = om.init_weights(21, bias_scale=0.01)
    w, b = weights["conv2d_85"]
    weights["conv2d_85"] = (w * 8.0, b + 0.3)
    wfile = str(tmp_path / "w.npz")
    np.savez(wfile, **{k + "/kernel:0": v[0] for k, v in weights.items()},
             **{k + "/bias:0": v[1] for k, v in weights.items()})
    os.environ["SR100_WEIGHTS"] = wfile
    try:
        m = models.DifvdsrDouble(1)
        got = m.upscale(p, return_image=True, mode="fast", verbose=False)
        m.upscale(p, mode="fast", verbose=False)
    finally:
        del os.environ["SR100_WEIGHTS"]
    want = np.clip(om.forward_numpy(weights, img[None].astype(np.float32) / 255.)[0] * 255., 0, 255).astype(np.uint8)
    assert got.shape == want.shape == (144, 208, 3)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1          # truncation flips near integers
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "pic_Ascaled(1x).png"))), img)
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "pic_scaled(1x).png"))), got)
    # mode='patch' (dense patches of the x4-bicubic image): tests/test_gpu_alt_tilers.py
