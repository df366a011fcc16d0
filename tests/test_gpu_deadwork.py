"""-m gpu: the tiled inference path with dead-work elimination (live tiles only, 272x272 HR stage) produces
bit-identical final images to the reference's literal scheme (every tile of the padded canvas, full 384x384 HR
stage, uncropped canvas then [0:4H, 0:4W], models.py:382-412), and matches the CPU oracle's tiling around the same
network."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sr_model():
    import models
    from oracle import model as om
    m = models.DifvdsrDouble(1)
    model = m.create_model(96, 96)
    model.engine.set_weights_dict(om.init_weights(77, bias_scale=0.05))   # biases: zero-input tiles are not zero
    return m


@pytest.mark.parametrize("shapes", [[(33, 50)], [(20, 20), (130, 70)], [(256, 256), (65, 200)], [(344, 228)]])
def test_cropped_path_is_bit_identical_to_full_tiling(sr_model, shapes):
    rng = np.random.default_rng(sum(h * w for h, w in shapes))
    imgs = [rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in shapes]
    fast = sr_model.upscale_arrays(imgs)
    canv = sr_model.upscale_arrays(imgs, return_canvas=True)
    for im, f, c in zip(imgs, fast, canv):
        h, w = im.shape[:2]
        assert f.shape == (4 * h, 4 * w, 3) and f.dtype == np.uint8
        assert np.array_equal(f, c[:4 * h, :4 * w])


def test_cropped_float_outputs_identical_on_owned_region(sr_model):
    """Same check before quantisation: the fp32 patch outputs agree exactly wherever the stitch reads them."""
    from sr100 import ops
    from sr100.engine import plan_tiles
    eng = sr_model.model.engine
    rng = np.random.default_rng(4)
    img = torch.from_numpy(rng.integers(0, 256, size=(150, 90, 3)).astype(np.uint8)).cuda()
    (gh, gw), counts, ext = plan_tiles(150, 90)
    p, got = ops.patch_gather_u8(img, (gh, gw), (96, 96), 64, divisor=255.0)
    full = eng.forward_device(p).clone()
    crop = eng.forward_device(p, extents=ext).clone()
    a, _ = ops.patch_stitch(full, counts, (96, 96), 64, 4, (150, 90), mul=255.0)
    b, _ = ops.patch_stitch(crop, counts, (96, 96), 64, 4, (150, 90), mul=255.0)
    assert torch.equal(a, b)
    for n, (eh, ew) in enumerate(ext):          # and the whole computed corner except the 7-px contaminated rim
        assert torch.equal(full[n, :eh - 7, :ew - 7], crop[n, :eh - 7, :ew - 7])


def test_upscale_step_patch_file_and_canvas(sr_model, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, size=(70, 101, 3)).astype(np.uint8)
    path = str(tmp_path / "x.png")
    Image.fromarray(img).save(path)
    wfile = str(tmp_path / "w.npz")
    sr_model.model.save_weights(wfile)
    import os
    os.environ["SR100_WEIGHTS"] = wfile
    try:
        sr_model._loaded_from = None
        canvas = sr_model.upscaleStepPatch(path, return_image=True, patch_size=96, verbose=False)
        sr_model.upscaleStepPatch(path, patch_size=96, verbose=False)
    finally:
        del os.environ["SR100_WEIGHTS"]
    out = np.asarray(Image.open(str(tmp_path / "x_scaled(1x).png")))
    assert canvas.shape == (4 * 192, 4 * 256, 3)                  # 70+96 -> 192, 101+96 -> 256 (models.py:248-256)
    assert out.shape == (280, 404, 3) and np.array_equal(out, canvas[:280, :404])


def test_large_image_config5_shape_and_identity(sr_model):
    """BASELINE.json configs[4] at full size (1080x1920 -> 4320x7680, 558 tiles in the reference grid, 510 run):
    the fast path equals the literal tiling bit for bit on every pixel of the final image."""
    rng = np.random.default_rng(9)
    from scipy.ndimage import uniform_filter
    img = uniform_filter(rng.integers(0, 256, size=(1080, 1920, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)
    fast = sr_model.upscale_arrays([img])[0]
    assert fast.shape == (4320, 7680, 3)
    canvas = sr_model.upscale_arrays([img], return_canvas=True)[0]
    assert canvas.shape == (4 * 1216, 4 * 2048, 3)
    assert np.array_equal(fast, canvas[:4320, :7680])


def test_edge_shapes_and_empty(sr_model):
    """Empty batch, 1x1 image, exact multiples of the 64-px step, very thin images: shapes follow 4H x 4W and the
    fast path still equals the literal tiling."""
    assert sr_model.upscale_arrays([]) == []
    rng = np.random.default_rng(21)
    shapes = [(1, 1), (64, 64), (32, 160), (1, 300), (97, 3)]
    imgs = [rng.integers(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in shapes]
    fast = sr_model.upscale_arrays(imgs)
    canv = sr_model.upscale_arrays(imgs, return_canvas=True)
    for im, f, c in zip(imgs, fast, canv):
        h, w = im.shape[:2]
        assert f.shape == (4 * h, 4 * w, 3)
        assert np.array_equal(f, c[:4 * h, :4 * w])
    with pytest.raises((ValueError, RuntimeError)):
        sr_model.upscale_arrays([np.zeros((8, 8), dtype=np.uint8)])       # not H x W x 3


def test_batch_of_equal_shapes_config3_equals_one_by_one(sr_model):
    """BASELINE.json configs[2] shape (339x510 LR images, 54 tiles each; 12 of the 64 here): a batch of same-shaped
    images goes through ONE gather launch and shared sub-batches of the conv stack -- every image must come out
    bit-identical to running it alone, in a mixed list too (runs of equal shapes are batched, the rest are not)."""
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(33)
    def im(h, w):
        return uniform_filter(rng.integers(0, 256, size=(h, w, 3)).astype(np.float32), size=(5, 5, 1)).astype(np.uint8)
    batch = [im(339, 510) for _ in range(12)]
    together = sr_model.upscale_arrays(batch)
    assert all(t.shape == (1356, 2040, 3) for t in together)
    for i in (0, 5, 11):
        assert np.array_equal(sr_model.upscale_arrays([batch[i]])[0], together[i])
    mixed = [batch[0], batch[1], im(100, 80), batch[2], im(100, 80), im(100, 80)]
    out = sr_model.upscale_arrays(mixed)
    assert np.array_equal(out[0], together[0]) and np.array_equal(out[1], together[1]) and np.array_equal(out[3], together[2])
    assert np.array_equal(out[2], sr_model.upscale_arrays([mixed[2]])[0])
    assert np.array_equal(out[5], sr_model.upscale_arrays([mixed[5]])[0])


@pytest.mark.parametrize("shape,world", [((200, 330), 3), ((97, 610), 4), ((339, 510), 8), ((130, 70), 2)])
def test_tile_sharded_image_is_bit_identical(shape, world):
    """BASELINE config 5's sharding run as a LOGICAL split on one GPU: every logical rank runs its contiguous range of
    the column-major live-tile index and stitches the pixels its tiles own into a uint8 column strip
    (sr_patch_stitch_range); the strips OR-ed together on rank 0 equal the single-rank image bit for bit."""
    from oracle import model as om
    from sr100.engine import Engine
    rng = np.random.default_rng(shape[0] + world)
    img = torch.from_numpy(rng.integers(0, 256, size=shape + (3,)).astype(np.uint8)).cuda()
    w = om.init_weights(1234, bias_scale=0.01)
    k, b = w["conv2d_85"]
    w["conv2d_85"] = (k * 8.0, b + 0.3)
    eng = Engine(w)
    want = eng.upscale_images_device([img])[0]
    strips = {}

    def collect(rank):
        def cb(send):
            strips[rank] = send.clone()
            return [strips[r] for r in range(world)] if rank == 0 else None
        return cb

    got = None
    for rank in reversed(range(world)):
        got = eng.upscale_image_sharded(img, world=world, rank=rank, group_gather=collect(rank))
    assert got is not None and got.shape == want.shape
    assert torch.equal(got, want)
    # what crosses the link: owned uint8 pixels, not fp32 patches
    _, counts, _, shards = eng.shard_plan(shape[0], shape[1], world)
    sent = sum(4 * shape[0] * max(s[3] - s[2] for s in shards) * 3 for s in shards[1:])
    patches = sum((s[1] - s[0]) * 384 * 384 * 3 * 4 for s in shards[1:])
    assert sent < patches
