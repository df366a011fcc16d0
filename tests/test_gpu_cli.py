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
