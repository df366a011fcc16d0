"""-m gpu: dataset preparation on the device (sr100.dataprep, img_utils.transform_images) against the oracle and the
reference function's own outputs.  Byte / integer / float64 work: bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import dataprep as odp
from oracle.refgen_dataprep import KEEP, synthetic_image

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,out", [((100, 130), (256, 256)), ((300, 500), (256, 256)), ((64, 64), (16, 16)),
                                       ((16, 16), (32, 32)), ((37, 51), (80, 23)), ((40, 40), (40, 17)),
                                       ((40, 40), (90, 40)), ((33, 20), (33, 20))])
@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
def test_resize_kernel_equals_pillow_restatement(shape, out, interp):
    from sr100 import dataprep
    rng = np.random.default_rng(shape[0] * 7 + out[1])
    imgs = rng.integers(0, 256, size=(3,) + shape + (3,)).astype(np.uint8)
    got = dataprep.resize_u8(torch.from_numpy(imgs).cuda(), out[0], out[1], interp).cpu().numpy()
    for n in range(3):
        assert np.array_equal(got[n], odp.resize_u8(imgs[n], out[0], out[1], interp))
    one = dataprep.resize_u8(torch.from_numpy(imgs[1]).cuda(), out[0], out[1], interp).cpu().numpy()
    assert np.array_equal(one, got[1])


@pytest.mark.parametrize("shape", [(256, 256), (5, 7), (3, 3), (2, 9), (1, 1)])
def test_sharpen_kernel_equals_pillow_restatement(shape):
    from sr100 import dataprep
    rng = np.random.default_rng(shape[1])
    imgs = np.stack([rng.integers(0, 256, size=shape + (3,)), rng.integers(0, 2, size=shape + (3,)) * 255]).astype(np.uint8)
    got = dataprep.sharpen_u8(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for n in range(2):
        assert np.array_equal(got[n], odp.sharpen_u8(imgs[n]))


@pytest.mark.parametrize("patch", [16, 32, 64])
def test_patch_samples_bytescale_and_gaussian_bit_exact(patch):
    from oracle.pil_resample import bytescale
    from sr100 import dataprep
    rng = np.random.default_rng(patch)
    img = synthetic_image(patch, 256, 256)
    img[:patch, :patch] = 77                                   # a constant sample: cscale == 0 -> 1
    pos = np.array([(0, 0), (256 - patch, 256 - patch), (16, 48), (100, 3)] +
                   [tuple(rng.integers(0, 256 - patch, size=2)) for _ in range(6)], dtype=np.int32)
    y, g = dataprep.patch_samples(torch.from_numpy(img).cuda(), pos, patch)
    y, g = y.cpu().numpy(), g.cpu().numpy()
    for i, (r, c) in enumerate(pos):
        ip = img[r:r + patch, c:c + patch].astype(np.float64)
        assert np.array_equal(y[i], bytescale(ip))
        assert np.array_equal(g[i], bytescale(odp.gaussian_filter_f64(ip, 0.5)))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_transform_image_device_equals_reference_output(golden_dir, tag):
    from sr100 import dataprep
    z = np.load(golden_dir + "/dataprep_ref.npz")
    seed, h, w, sf, tu = [int(v) for v in z[tag + "_meta"]]
    y, x = dataprep.transform_image_device(torch.from_numpy(synthetic_image(seed, h, w)).cuda(), sf, bool(tu))
    y, x = y.cpu().numpy(), x.cpu().numpy()
    assert list(y.shape) + list(x.shape) == list(z[tag + "_shapes"])
    assert np.array_equal(y[KEEP], z[tag + "_y_keep"]) and np.array_equal(x[KEEP], z[tag + "_x_keep"])
    assert hashlib.sha256(y.tobytes()).digest() == z[tag + "_y_sha"].tobytes()
    assert hashlib.sha256(x.tobytes()).digest() == z[tag + "_x_sha"].tobytes()


def test_transform_images_writes_the_reference_files(tmp_path, capsys):
    from PIL import Image
    import img_utils
    src, dst = str(tmp_path / "in") + "/", str(tmp_path / "out") + "/"
    os.makedirs(src)
    imgs = {"a.png": synthetic_image(3, 120, 90), "b.bmp": synthetic_image(4, 300, 280)}
    for name, im in imgs.items():
        Image.fromarray(im).save(src + name)
    img_utils.transform_images(src, dst, scaling_factor=2, max_nb_images=-1, true_upscale=False)
    out = capsys.readouterr().out
    assert "Transforming 2 images." in out and "Images transformed." in out
    assert len(os.listdir(dst + "X/")) == 512 and len(os.listdir(dst + "y/")) == 512
    for index, name in enumerate(os.listdir(src), start=1):
        ys, xs = odp.transform_image(imgs[name], 2, False)
        for i in (0, 57, 255):
            assert np.array_equal(np.asarray(Image.open(dst + "y/%d_%d.png" % (index, i + 1))), ys[i])
            assert np.array_equal(np.asarray(Image.open(dst + "X/%d_%d.png" % (index, i + 1))), xs[i])
    with pytest.raises(ValueError):
        from sr100 import dataprep
        dataprep.transform_image_device(torch.zeros(64, 64, 3, dtype=torch.uint8, device="cuda"), scaling_factor=8)
