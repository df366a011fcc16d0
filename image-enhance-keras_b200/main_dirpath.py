"""Drop-in for the reference CLI (main_dirpath.py:1-55):

    python main_dirpath.py <imgpath/> [--model didbl] [--scale 1] [--mode fast|patch] [--save_intermediate False]
                                      [--suffix scaled] [--patch_size 8]

Same flags, defaults, validation messages and behaviour: every entry of the directory (`imgpath + name`, so the
trailing slash matters exactly as it does there, main_dirpath.py:50-51) goes through
`DifvdsrDouble.upscaleStepPatch(scalemulti=4, patch_size=96)` (main_dirpath.py:53).  As in the reference `--mode` and
`--patch_size` are validated and then unused, and 'didbl' is the only model name its assert lets through
(main_dirpath.py:27; the branches for the other names below it are unreachable).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

# flag -> (argparse keywords); the schema of main_dirpath.py:6-16
_CLI = (
    ("imgpath", dict(type=str, help="directory of input images (include the trailing slash)")),
    ("--model", dict(type=str, default="didbl", help="network name; only 'didbl' (DifvdsrDouble) is accepted")),
    ("--scale", dict(default=1, help="value passed to the model constructor (it ends up in the output file name)")),
    ("--mode", dict(type=str, default="fast", help="'fast' or 'patch'; checked, not used")),
    ("--save_intermediate", dict(dest="save", type=str, default="False", help="also write the intermediate image")),
    ("--suffix", dict(type=str, default="scaled", help="suffix of the written files")),
    ("--patch_size", dict(type=int, default=8, help="positive integer; checked, not used (tiles are 96 px)")),
)
_TRUE_WORDS = ("true", "yes", "t", "1")


def strToBool(v):
    return v.lower() in _TRUE_WORDS


def parse(argv=None):
    """Parsed and validated settings as a dict; raises AssertionError with the reference's messages
    (main_dirpath.py:27, 31, 37)."""
    ap = argparse.ArgumentParser(description="x4 super-resolution of every image in a directory (sr100 engine)")
    for flag, kw in _CLI:
        ap.add_argument(flag, **kw)
    ns = ap.parse_args(argv)
    model_type, mode = str(ns.model).lower(), str(ns.mode).lower()
    assert model_type in ["didbl"], 'Model type must be either "sr", "esr", "dsr", "ddsr" or "rnsr"'
    assert mode in ["fast", "patch"], 'Mode of operation must be either "fast" or "patch"'
    scale_factor, save = int(ns.scale), strToBool(ns.save)
    patch_size = int(ns.patch_size)
    assert patch_size > 0, "Patch size must be a positive integer"
    return dict(path=ns.imgpath, suffix=ns.suffix, model_type=model_type, mode=mode, scale_factor=scale_factor,
                save=save, patch_size=patch_size)


def main(argv=None):
    cfg = parse(argv)
    import models
    net = models.DifvdsrDouble(cfg["scale_factor"])
    for name in os.listdir(cfg["path"]):
        net.upscaleStepPatch(cfg["path"] + name, save_intermediate=cfg["save"], scalemulti=4, patch_size=96,
                             suffix=cfg["suffix"])


if __name__ == "__main__":
    main()
