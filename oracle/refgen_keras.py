"""Golden vectors of the FLOAT path from the real reference stack -- the recipe for closing "parity unpinned".

TEST INFRASTRUCTURE.  This script cannot run in the authoring container or on the GPU box: it needs the 2018
stack the reference was written for -- none of it is in the offline wheelhouse:

    python 3.6        tensorflow 1.12 - 1.15 (CPU is enough)     keras 2.2.4       h5py 2.x
    numpy < 1.20      scipy 1.1 - 1.2 (scipy.misc.imread/imresize) + Pillow < 7      scikit-image 0.14 - 0.17
    scikit-learn < 0.24

On such a machine, from a checkout of this repository next to a checkout of the reference:

    CUDA_VISIBLE_DEVICES=-1 python oracle/refgen_keras.py /path/to/image-enhance-keras

It imports the reference's OWN models.py (DifvdsrDouble.create_model, models.py:1159-1222), sets the seed-1234
glorot weights of oracle/model.init_weights (biases U(+-0.01)) by layer name, and writes

    tests/golden/keras_ref.npz          model.predict outputs, mse loss and per-layer gradients of one batch,
                                        weights after one Adam step (model.train_on_batch), tf.image.resize_bilinear
                                        x4 of a seeded tensor, skimage rgb2ycbcr / compare_ssim of seeded image pairs,
                                        model.layers names (the depth-sorted order) and get_weights() shapes
    tests/golden/keras_block53_ref.h5   model.save_weights() of a 4-channel two-branch 5/3 block built by the
                                        reference's own _residual_block_light53 (a real libhdf5/Keras-written file for
                                        sr100.h5lite's reader, layer order and layout)

    (keras_ref.npz["h5lite_writer_loaded_by_keras"]: the reverse check -- a file written by sr100.h5lite is loaded by
     Keras' own load_weights through libhdf5 and compared)

tests/test_keras_golden.py consumes both files when they exist (and is skipped, saying so, while they do not):
the oracle restatement (oracle/model.py, oracle/scoring.py, sr100.keras_graph) is then pinned to the real stack
and DESIGN.md's "parity unpinned" rows can be struck.  Everything is seeded; the inputs are stored next to the
outputs so the consumer does not depend on numpy's generator being identical across versions.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def smooth(rng, shape):
    """Seeded, not-white-noise test images in [0,1] (5x5 box blur of uniform noise), float32 NHWC."""
    from scipy.ndimage import uniform_filter
    return uniform_filter(rng.random_sample(shape).astype(np.float32), size=(1, 5, 5, 1)).astype(np.float32)


def main(ref_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, ref_dir)
    os.chdir(ref_dir)                      # img_utils / models use relative paths (weights_Double/...)
    import keras
    from keras import backend as K
    from keras.layers import Input, Conv2D
    from keras.models import Model
    import tensorflow as tf
    import models as ref_models           # the reference's models.py, unmodified
    from oracle import model as om

    out = {"versions": np.array([keras.__version__, tf.__version__, np.__version__])}
    weights = om.init_weights(1234, bias_scale=0.01)
    rng = np.random.RandomState(20181019)

    # ---------------------------------------------------------------- forward: model.predict (models.py:342)
    for tag, shape in (("a", (1, 24, 24, 3)), ("b", (2, 16, 20, 3))):
        K.clear_session()                 # fresh auto-names: level1, conv2d_1 .. conv2d_85
        m = ref_models.DifvdsrDouble(1)
        model = m.create_model(shape[1], shape[2], load_weights=False)
        for name, (w, b) in weights.items():
            model.get_layer(name).set_weights([w, b])
        x = smooth(rng, shape)
        out["x_" + tag] = x
        out["predict_" + tag] = model.predict(x, batch_size=1).astype(np.float32)
        if tag == "a":
            out["layer_names"] = np.array([l.name for l in model.layers])
            out["weighted_layer_names"] = np.array([l.name for l in model.layers if l.weights])
            out["get_weights_shapes"] = np.array([str(w.shape) for w in model.get_weights()])
        if tag == "b":
            # -------------------------------------------------------- loss + gradients (compile(mse), models.py:1213)
            y = smooth(rng, (shape[0], 4 * shape[1], 4 * shape[2], 3))
            out["y_b"] = y
            y_ph = K.placeholder(shape=(None, None, None, 3))
            loss = K.mean(K.square(model.output - y_ph))
            tw = model.trainable_weights
            f = K.function([model.input, y_ph], [loss] + K.gradients(loss, tw))
            res = f([x, y])
            out["loss_b"] = np.array(res[0])
            for wv, g in zip(tw, res[1:]):
                out["grad/" + wv.name] = np.asarray(g, dtype=np.float32)      # e.g. grad/conv2d_3/kernel:0
            # -------------------------------------------------------- one Adam step (Adam(1e-4, 0.9), models.py:1212)
            l1 = model.train_on_batch(x, y)
            out["train_loss_b"] = np.array(l1[0] if isinstance(l1, (list, tuple)) else l1)
            for name in ("level1", "conv2d_1", "conv2d_2", "conv2d_3", "conv2d_66", "conv2d_85"):
                w1, b1 = model.get_layer(name).get_weights()
                out["adam1/" + name + "/kernel"] = w1
                out["adam1/" + name + "/bias"] = b1

    # ---------------------------------------------------------------- tf.image.resize_bilinear x4 (models.py:1392-1399)
    K.clear_session()
    t = rng.random_sample((2, 7, 9, 5)).astype(np.float32)
    out["bilinear_in"] = t
    out["bilinear_x4"] = K.get_session().run(tf.image.resize_bilinear(tf.constant(t), [28, 36]))

    # ---------------------------------------------------------------- scoring (scorpath.py:190-191, 226, 228)
    from skimage.color import rgb2ycbcr
    from skimage.measure import compare_ssim
    a = (smooth(rng, (1, 60, 76, 3))[0] * 255).astype(np.uint8)
    b = np.clip(a.astype(np.float64) + rng.normal(0, 3, size=a.shape), 0, 255).astype(np.uint8)
    out["score_a"], out["score_b"] = a, b
    ya, yb = rgb2ycbcr(a)[:, :, 0], rgb2ycbcr(b)[:, :, 0]
    out["rgb2ycbcr_y_a"] = ya
    out["ssim_y"] = np.array(compare_ssim(ya, yb, data_range=255))
    out["ssim_rgb"] = np.array(compare_ssim(a, b, data_range=255, multichannel=True))

    # ---------------------------------------------------------------- a real Keras/libhdf5-written weight file
    K.clear_session()
    m = ref_models.DifvdsrDouble(1)
    inp = Input(shape=(8, 8, 3))
    h = Conv2D(4, (1, 1), activation="relu", padding="same", name="level1")(inp)
    h = m._residual_block_light53(h, 4, train=True)      # the reference's own block builder (models.py:1248-1270)
    h = m._residual_block_light(h, 4, train=True)        # models.py:1231-1245
    small = Model(inp, Conv2D(3, (3, 3), padding="same", activation="relu")(h))
    r2 = np.random.RandomState(7)
    for l in small.layers:
        if l.weights:
            l.set_weights([r2.normal(size=w.shape).astype(np.float32) for w in l.get_weights()])
    small.save_weights(os.path.join(OUT, "keras_block53_ref.h5"))
    out["block53_layer_names"] = np.array([l.name for l in small.layers])
    for l in small.layers:
        for wv, val in zip(l.weights, l.get_weights()):
            out["block53/" + wv.name] = val
    xs = smooth(rng, (1, 8, 8, 3))
    out["block53_x"], out["block53_predict"] = xs, small.predict(xs)
    # ... and the other direction: a file written by sr100.h5lite (pure Python, no libhdf5) opened by h5py / loaded by
    # Keras' own load_weights into the same model -- the pin for the WRITER
    sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))
    from sr100 import h5lite
    ours = os.path.join(OUT, "h5lite_written_block53.h5")
    wd = {l.name: tuple(l.get_weights()) for l in small.layers if l.weights}
    h5lite.save_keras_weights(ours, wd, layers=[(l.name, bool(l.weights)) for l in small.layers])
    before = [w.copy() for w in small.get_weights()]
    for l in small.layers:
        if l.weights:
            l.set_weights([np.zeros_like(w) for w in l.get_weights()])
    small.load_weights(ours)                               # Keras + libhdf5 reading our bytes
    ok = all(np.array_equal(a, b) for a, b in zip(before, small.get_weights()))
    out["h5lite_writer_loaded_by_keras"] = np.array(ok)
    assert ok, "Keras did not read back what sr100.h5lite wrote"

    np.savez_compressed(os.path.join(OUT, "keras_ref.npz"), **out)
    print("wrote", os.path.join(OUT, "keras_ref.npz"), "and keras_block53_ref.h5 (keras %s, tf %s)"
          % (keras.__version__, tf.__version__))


if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    main(os.path.abspath(sys.argv[1]))
