"""sr100.h5lite (HDF5 without h5py) - the weight-file row of SURVEY.md 8(f): `model.load_weights(...h5)`
(models.py:1217-1218) and `ModelCheckpoint(... .h5, save_weights_only=True)` (models.py:141-142).

The reader is pinned against a file written by libhdf5 (tests/golden/matlab73_testdouble.h5 = scipy's
`testhdf5_7.4_GLNX86.mat`, MATLAB 7.3: superblock v0 behind a 512-byte user block, symbol-table group, object
header v1, layout message v2, attribute v1); the writer's datatype / dataspace / attribute messages are compared
byte for byte with the ones libhdf5 put into that file."""
import os

import numpy as np
import pytest

from sr100 import h5lite as h


@pytest.fixture(scope="module")
def golden(golden_dir):
    return os.path.join(golden_dir, "matlab73_testdouble.h5")


def test_reads_libhdf5_written_file(golden):
    f = h.open_file(golden)
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.is_dataset and not f.is_dataset
    a = d.read()
    assert a.dtype == np.dtype("<f8") and a.shape == (9, 1)
    assert np.array_equal(a.ravel(), np.arange(9) * (np.pi / 4))
    assert d.attrs == {"MATLAB_class": b"double"}
    with pytest.raises(KeyError):
        f["nope"]


def test_writer_messages_match_libhdf5_bytes(golden):
    raw = open(golden, "rb").read()
    # dataset object header of /testdouble at file offset 1488: messages laid out by libhdf5 1.8
    assert h._dtype_msg(np.float64) == raw[1528:1548]
    assert h._space_msg((9, 1)) == raw[1560:1584]
    assert h._attr_msg("MATLAB_class", b"double") == raw[1648:1696 - 2]      # libhdf5 pads the message to 8
    # group structures: heap prefix / B-tree node / symbol node signatures and field layout
    data = h.write_file(None, {"testdouble": np.arange(9).reshape(9, 1) * (np.pi / 4)},
                        {"": {}, "testdouble": {}})
    assert data[:8] == raw[512:520] and data[8:20] == raw[520:532]          # superblock v0 prefix incl. K values
    g = h.open_file(data)
    assert np.array_equal(g["testdouble"].read(), h.open_file(golden)["testdouble"].read())
    snod = data.index(b"SNOD")
    gs = raw.index(b"SNOD")
    assert data[snod:snod + 8] == raw[gs:gs + 8]                             # version 1, one symbol
    assert data[snod + 8:snod + 16] == raw[gs + 8:gs + 16]                   # name offset 8 in the local heap
    tree, gt = data.index(b"TREE"), raw.index(b"TREE")
    assert data[tree:tree + 32] == raw[gt:gt + 32]                           # level 0, 1 entry, no siblings, key 0
    assert data[tree + 40:tree + 48] == raw[gt + 40:gt + 48]                 # key 1 = heap offset of the last name
    assert raw.index(b"\x01\x00\x02\x00\x01\x00\x00\x00", gt) - gt == 544     # libhdf5: node sized for 2K = 32 children
    assert data[tree + 544:tree + 552] == b"\x01\x00\x01\x00\x01\x00\x00\x00"  # ours: the group's object header follows


def _weights(rng, cin=8):
    w = {"level1": (rng.normal(size=(1, 1, 3, cin)).astype(np.float32), rng.normal(size=cin).astype(np.float32))}
    for i in range(1, 86):
        k = 3 if i % 2 else 5
        w["conv2d_%d" % i] = (rng.normal(size=(k, k, cin, cin)).astype(np.float32),
                              rng.normal(size=cin).astype(np.float32))
    return w


def test_keras_layout_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    w = _weights(rng)
    p = str(tmp_path / "weights025-17-0.93.h5")
    h.save_keras_weights(p, w)
    f = h.open_file(p)
    assert f.attrs["backend"] == b"tensorflow"
    assert [n.decode() for n in f.attrs["layer_names"]] == list(w)            # creation order, not sorted
    assert f.keys() == sorted(w)                                              # 86 links: 11 symbol nodes
    assert [n.decode() for n in f["conv2d_10"].attrs["weight_names"]] == ["conv2d_10/kernel:0", "conv2d_10/bias:0"]
    assert f["conv2d_10/conv2d_10"].keys() == ["bias:0", "kernel:0"]
    r = h.load_keras_weights(p)
    assert list(r) == list(w)
    for n in w:
        assert r[n][0].dtype == np.float32
        assert np.array_equal(r[n][0], w[n][0]) and np.array_equal(r[n][1], w[n][1])
    with pytest.raises(KeyError):
        h.load_keras_weights(p, ["level1", "conv2d_86"])


def test_full_model_file_and_chunked_deflate(tmp_path):
    rng = np.random.default_rng(1)
    k, b = rng.normal(size=(3, 3, 16, 8)).astype(np.float32), rng.normal(size=8).astype(np.float32)
    tree = {"model_weights": {"conv2d_1": {"conv2d_1": {"kernel:0": k, "bias:0": b}}, "activation_1": {}},
            "optimizer_weights": {"iterations:0": np.int64(7)}}
    attrs = {"model_weights": {"layer_names": [b"conv2d_1", b"activation_1"]},
             "model_weights/conv2d_1": {"weight_names": [b"conv2d_1/kernel:0", b"conv2d_1/bias:0"]},
             "model_weights/activation_1": {"weight_names": np.zeros((0,), dtype="S1")}}
    for kw in ({}, {"chunk_rows": 2, "deflate": 4}, {"chunk_rows": 1}):
        p = str(tmp_path / "m.h5")
        h.write_file(p, tree, attrs, **kw)
        r = h.load_keras_weights(p)
        assert list(r) == ["conv2d_1"] and np.array_equal(r["conv2d_1"][0], k) and np.array_equal(r["conv2d_1"][1], b)
        assert h.open_file(p)["optimizer_weights/iterations:0"].read() == 7
    big = rng.integers(-5, 5, size=(37, 5, 3)).astype(np.int16)
    h.write_file(str(tmp_path / "c.h5"), {"a": big}, chunk_rows=8, deflate=9)
    assert np.array_equal(h.open_file(str(tmp_path / "c.h5"))["a"].read(), big)     # ragged last chunk


def test_rejects_non_hdf5(tmp_path):
    p = str(tmp_path / "x.h5")
    with open(p, "wb") as f:
        f.write(b"not an hdf5 file" * 100)
    with pytest.raises(OSError):
        h.open_file(p)


def test_keras_model_layers_order_of_difvdsr_double():
    """Keras sorts `model.layers` by depth, then by traversal order from the output (sr100.keras_graph): in every 5/3
    block the weighted layers come a3, c5, b5, d3 -- conv2d_1, conv2d_3, conv2d_2, conv2d_4 -- not in creation order;
    light blocks and the chain around them are in creation order."""
    from sr100 import keras_graph as kg
    layers = kg.difvdsr_double_layers()
    assert len(layers) == 1 + 1 + 18 * 10 + 6 * 5 + 1 + 1 and layers[0] == ("input_1", False)
    assert [n for n, _ in layers[1:12]] == ["level1", "conv2d_1", "conv2d_3", "activation_1", "activation_2", "conv2d_2",
                                            "conv2d_4", "add_1", "lambda_2", "lambda_1", "add_2"]
    w = kg.difvdsr_double_weighted_order()
    want = ["level1"]
    n = 0
    for _ in range(16):
        want += ["conv2d_%d" % (n + i) for i in (1, 3, 2, 4)]
        n += 4
    want += ["conv2d_%d" % i for i in range(65, 77)]
    n = 76
    for _ in range(2):
        want += ["conv2d_%d" % (n + i) for i in (1, 3, 2, 4)]
        n += 4
    want.append("conv2d_85")
    assert w == want and sorted(w, key=lambda s: (s != "level1", int(s.split("_")[1]) if "_" in s else 0)) == \
        ["level1"] + ["conv2d_%d" % i for i in range(1, 86)]


def test_save_weights_layout_is_keras_model_layers_order(tmp_path):
    """The writer emits `layer_names` in model.layers order with weightless groups (empty weight_names), and the
    positional reader returns the weighted layers in exactly the order Keras' load_weights zips them."""
    from sr100 import keras_graph as kg
    rng = np.random.default_rng(3)
    w = _weights(rng)
    p = str(tmp_path / "w.h5")
    h.save_keras_weights(p, w, layers=kg.difvdsr_double_layers())
    f = h.open_file(p)
    names = [n.decode() for n in f.attrs["layer_names"]]
    assert names == [n for n, _ in kg.difvdsr_double_layers()]
    assert len(f["activation_1"].attrs["weight_names"]) == 0 and f["add_3"].keys() == []
    pos = h.load_keras_weights_positional(p)
    assert [n for n, _ in pos] == kg.difvdsr_double_weighted_order()
    assert [n for n, _ in pos][1:5] == ["conv2d_1", "conv2d_3", "conv2d_2", "conv2d_4"]
    for n, arrs in pos:
        assert np.array_equal(arrs[0], w[n][0]) and np.array_equal(arrs[1], w[n][1])
    assert set(h.load_keras_weights(p)) == set(w)            # by name: weightless groups are skipped
