"""CPU: argument handling of the `main_dirpath.py` drop-in (flags, defaults, validation messages and the per-file call of
the reference, main_dirpath.py:6-53) with the model stubbed out -- no device needed."""
import importlib.util
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def cli():
    spec = importlib.util.spec_from_file_location("sr_cli_under_test",
                                                  os.path.join(ROOT, "image-enhance-keras_b200", "main_dirpath.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_defaults_match_the_reference_schema(cli):
    cfg = cli.parse(["some/dir/"])
    assert cfg == dict(path="some/dir/", suffix="scaled", model_type="didbl", mode="fast", scale_factor=1, save=False,
                       patch_size=8)
    cfg = cli.parse(["d/", "--model", "DIDBL", "--scale", "4", "--mode", "PATCH", "--save_intermediate", "Yes",
                     "--suffix", "x", "--patch_size", "64"])
    assert cfg == dict(path="d/", suffix="x", model_type="didbl", mode="patch", scale_factor=4, save=True, patch_size=64)


@pytest.mark.parametrize("argv, message", [
    (["d/", "--model", "sr"], "Model type must be"),
    (["d/", "--mode", "slow"], "Mode of operation must be"),
    (["d/", "--patch_size", "0"], "Patch size must be a positive integer"),
])
def test_validation_messages(cli, argv, message):
    with pytest.raises(AssertionError, match=message):
        cli.parse(argv)


def test_str_to_bool(cli):
    assert [cli.strToBool(v) for v in ("True", "yes", "T", "1", "False", "no", "0", "")] == [True] * 4 + [False] * 4


def test_every_directory_entry_goes_through_upscale_step_patch(cli, tmp_path, monkeypatch):
    for name in ("a.png", "b.bmp"):
        (tmp_path / name).write_bytes(b"")
    calls = []

    class FakeModel:
        def __init__(self, scale_factor):
            calls.append(("init", scale_factor))

        def upscaleStepPatch(self, path, **kw):
            calls.append((path, kw))

    monkeypatch.setitem(sys.modules, "models", types.SimpleNamespace(DifvdsrDouble=FakeModel))
    d = str(tmp_path) + "/"
    cli.main([d, "--scale", "2", "--suffix", "up", "--save_intermediate", "true"])
    assert calls[0] == ("init", 2)
    want_kw = dict(save_intermediate=True, scalemulti=4, patch_size=96, suffix="up")
    assert sorted(calls[1:]) == sorted([(d + "a.png", want_kw), (d + "b.bmp", want_kw)])


def test_learn_script_builds_the_model_and_fits_180_epochs(monkeypatch):
    """learn.py:11-22: DifvdsrDouble(1).create_model(); fit(nb_epochs=180)."""
    spec = importlib.util.spec_from_file_location("sr_learn_under_test",
                                                  os.path.join(ROOT, "image-enhance-keras_b200", "learn.py"))
    learn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(learn)
    calls = []

    class FakeModel:
        def __init__(self, scale):
            calls.append(("init", scale))

        def create_model(self):
            calls.append(("create_model",))

        def fit(self, **kw):
            calls.append(("fit", kw))

    monkeypatch.setitem(sys.modules, "models", types.SimpleNamespace(DifvdsrDouble=FakeModel))
    learn.main([])
    assert calls == [("init", 1), ("create_model",), ("fit", dict(nb_epochs=180))]
    del calls[:]
    learn.main(["3"])
    assert calls[-1] == ("fit", dict(nb_epochs=3))
