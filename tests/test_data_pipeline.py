"""CPU: the training data pipeline (img_utils._index_generator / image_generator, reference img_utils.py:290-398)
against golden outputs of the REFERENCE functions themselves (oracle/refgen.py -> tests/golden/generator_ref.npz),
plus the Keras-callback records used by fit (ModelCheckpoint filename template models.py:141-142, 1155;
HistoryCheckpoint advanced.py:9-46)."""
import os

import numpy as np


def test_index_generator_matches_reference(golden_dir):
    import img_utils
    z = np.load(os.path.join(golden_dir, "generator_ref.npz"))
    for gi in range(4):
        N, bs, shuffle, seed = (int(v) for v in z["g%d_args" % gi])
        g = img_utils._index_generator(N, bs, bool(shuffle), None if seed < 0 else seed)
        if seed < 0:
            np.random.seed(0)          # unshuffled: no randomness involved
        for step in range(9):
            idx, cur, cbs = next(g)
            want = z["g%d_idx" % gi][step]
            want = want[want >= 0]
            assert list(idx) == list(want) and cur == z["g%d_cur" % gi][step] and cbs == z["g%d_bs" % gi][step]


def test_image_generator_matches_reference(golden_dir, tmp_path):
    """Same files, same seed -> the same (batch_x, batch_y) the reference yields (float64 NHWC in [0,1])."""
    import img_utils
    from PIL import Image
    z = np.load(os.path.join(golden_dir, "generator_ref.npz"))
    d = str(tmp_path) + "/"
    os.makedirs(d + "X")
    os.makedirs(d + "y")
    drng = np.random.default_rng(77)            # the generator of the files in oracle/refgen.py
    for k in range(5):
        Image.fromarray(drng.integers(0, 256, size=(16, 16, 3)).astype(np.uint8)).save(d + "X/im%d.png" % k)
        Image.fromarray(drng.integers(0, 256, size=(16, 16, 3)).astype(np.uint8)).save(d + "y/im%d.png" % k)
    ig = img_utils.image_generator(d, scale_factor=1, shuffle=True, batch_size=2, seed=5)
    bx, by = next(ig)
    bx2, by2 = next(ig)
    for got, key in ((bx, "ig_bx"), (by, "ig_by"), (bx2, "ig_bx2"), (by2, "ig_by2")):
        assert got.dtype == z[key].dtype == np.float64 and np.array_equal(got, z[key])


def test_x4_pairs_are_accepted(tmp_path):
    """The reference generator as shipped only yields equal-size X / y (so its x4 model cannot train on it,
    SURVEY 0); ours takes the shapes from the files, so (h,w) / (4h,4w) pairs feed the x4 model."""
    import img_utils
    from PIL import Image
    d = str(tmp_path) + "/"
    os.makedirs(d + "X")
    os.makedirs(d + "y")
    rng = np.random.default_rng(1)
    for k in range(3):
        Image.fromarray(rng.integers(0, 256, size=(8, 8, 3)).astype(np.uint8)).save(d + "X/%d.png" % k)
        Image.fromarray(rng.integers(0, 256, size=(32, 32, 3)).astype(np.uint8)).save(d + "y/%d.png" % k)
    bx, by = next(img_utils.image_generator(d, scale_factor=1, batch_size=3, shuffle=False))
    assert bx.shape == (3, 8, 8, 3) and by.shape == (3, 32, 32, 3) and 0.0 <= bx.min() and by.max() <= 1.0


def test_checkpoint_and_history_callbacks(tmp_path):
    import advanced
    import models

    class FakeModel(object):
        saved = []

        def save_weights(self, path, overwrite=True):
            self.saved.append(path)

    os.chdir(str(tmp_path))
    cb = models._ModelCheckpoint("weights_Double/weights025-{epoch:02d}-{val_acc:.2f}.h5", monitor='val_PSNRLoss',
                                 save_best_only=False, mode='max', save_weights_only=True, period=1)
    fm = FakeModel()
    cb.set_model(fm)
    cb.on_epoch_end(16, {"val_acc": 0.9312, "loss": 0.1})
    assert fm.saved == ["weights_Double/weights025-17-0.93.h5"]       # the shipped file name (models.py:1217)
    hc = advanced.HistoryCheckpoint(str(tmp_path / "hist.txt"))
    hc.on_train_begin({})
    hc.on_epoch_end(0, {"loss": 0.5, "val_loss": 0.6})
    hc.on_epoch_end(1, {"loss": 0.4, "val_loss": 0.5})
    txt = open(str(tmp_path / "hist.txt")).read()
    assert "loss" in txt and "0.4" in txt
