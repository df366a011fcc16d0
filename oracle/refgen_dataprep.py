"""Golden fixture for the dataset-preparation path: the REFERENCE'S OWN img_utils.transform_images
(img_utils.py:44-123, read-only /root/reference) run verbatim on one synthetic image, with scipy.misc (removed from
scipy 1.3) stubbed by PIL-backed functions written from scipy 1.2's pilutil.py (toimage = bytescale + fromarray,
imresize = toimage + Image.resize, imfilter = Image.filter(ImageFilter.SHARPEN), imsave = toimage + save) and the
real scipy.ndimage.gaussian_filter.  Runs only in the authoring container; output: tests/golden/dataprep_ref.npz.
TEST INFRASTRUCTURE.

    python oracle/refgen_dataprep.py
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refgen  # noqa: E402
from oracle.pil_resample import bytescale  # noqa: E402

KEEP = [0, 1, 13, 14, 100, 195, 196, 255]       # patches stored in full (the rest are covered by the digests)


def synthetic_image(seed, h, w):
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(seed)
    return uniform_filter(rng.integers(0, 256, size=(h, w, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)


def main():
    from PIL import Image, ImageFilter
    iu = refgen.load_img_utils()
    import scipy.misc as misc
    saved = {}

    def toimage(arr):
        return Image.fromarray(bytescale(np.asarray(arr)))

    def imresize(arr, size, interp='bilinear', mode=None):
        func = {'nearest': 0, 'lanczos': 1, 'bilinear': 2, 'bicubic': 3, 'cubic': 3}
        return np.asarray(toimage(arr).resize((size[1], size[0]), resample=func[interp]))

    def imfilter(arr, ftype):
        assert ftype == 'sharpen'
        return np.asarray(toimage(arr).filter(ImageFilter.SHARPEN))

    def imsave(name, arr):
        saved[name] = np.asarray(toimage(arr))

    misc.imfilter = imfilter
    iu.imresize, iu.imsave = imresize, imsave          # names bound by `from scipy.misc import ...` (img_utils.py:5)
    out = {}
    for tag, seed, h, w, sf, true_up in (("a", 11, 200, 310, 2, False), ("b", 12, 256, 256, 4, True)):
        src = tempfile.mkdtemp(prefix="dp_in_") + "/"
        dst = tempfile.mkdtemp(prefix="dp_out_") + "/"
        img = synthetic_image(seed, h, w)
        Image.fromarray(img).save(src + "img.png")
        saved.clear()
        refgen.quiet(iu.transform_images, src, dst, scaling_factor=sf, max_nb_images=-1, true_upscale=true_up)
        n = 256
        ys = np.stack([saved[dst + "/y/" + "1_%d.png" % (i + 1)] for i in range(n)])
        xs = np.stack([saved[dst + "/X/" + "1_%d.png" % (i + 1)] for i in range(n)])
        out[tag + "_meta"] = np.array([seed, h, w, sf, int(true_up)])
        out[tag + "_y_keep"], out[tag + "_x_keep"] = ys[KEEP], xs[KEEP]
        out[tag + "_y_sha"] = np.frombuffer(hashlib.sha256(ys.tobytes()).digest(), dtype=np.uint8)
        out[tag + "_x_sha"] = np.frombuffer(hashlib.sha256(xs.tobytes()).digest(), dtype=np.uint8)
        out[tag + "_shapes"] = np.array(list(ys.shape) + list(xs.shape))
    out["keep"] = np.array(KEEP)
    path = os.path.join(refgen.OUT, "dataprep_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
