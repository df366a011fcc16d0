"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/sr100.h
declares, and its host-only entry points agree with the reference arithmetic.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sr100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libsr100.so does not export %s" % n


def test_ctypes_signatures_cover_header(lib):
    from sr100 import _lib as L
    assert sorted(L.SIGNATURES) == declared_symbols()


def test_version_and_error_string(lib):
    assert lib.sr_version() == 100
    from sr100 import _lib as L
    rc = lib.sr_conv_plan_create(None, None)
    assert rc == -1
    assert b"null" in lib.sr_last_error_string()
    d = L.ConvDesc()
    d.nsrc, d.cin, d.cout = 3, 128, 128
    plan = C.c_void_p()
    assert lib.sr_conv_plan_create(C.byref(d), C.byref(plan)) == -1
    d.nsrc, d.cin = 1, 64
    assert lib.sr_conv_plan_create(C.byref(d), C.byref(plan)) == -2      # SR_ERR_UNSUPPORTED
    assert b"cin" in lib.sr_last_error_string()


def test_struct_layouts_match_the_library(lib):
    from sr100 import _lib as L
    structs = (L.ConvDesc, L.ConvPlanInfo, L.PackItem, L.WgradDesc, L.WgradPlanInfo, L.ScoreResult, L.ModelConfig,
               L.ForwardDesc, L.TrainDesc, L.ModelRunInfo, L.StitchTile)
    for which, st in enumerate(structs):
        assert lib.sr_abi_struct_size(which) == C.sizeof(st), st.__name__
    assert lib.sr_abi_struct_size(99) == 0
    # the binding stub printed in INTEGRATION.md declares the same sr_conv_desc
    import re
    src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = src[src.index("class sr_conv_desc"):src.index("stream = ctypes.c_void_p")]
    names = re.findall(r'\("(\w+)",', block)
    assert names == [f[0] for f in L.ConvDesc._fields_]


def test_patch_count_and_canvas_match_reference(lib, golden_dir):
    geo = np.load(golden_dir + "/tiling_ref.npz")["geometry_96_64"]
    for h, w, ch, cw, cnt_h, cnt_w in geo.tolist():
        a, b = C.c_int(), C.c_int()
        assert lib.sr_canvas_size(h, w, 96, 64, C.byref(a), C.byref(b)) == 0
        assert (a.value, b.value) == (ch, cw)
        assert lib.sr_patch_count(ch, 96, 64) == cnt_h
        assert lib.sr_patch_count(cw, 96, 64) == cnt_w
    # brute force against the reference's range()/modulo loops (img_utils.py:622,629)
    for dim in range(1, 80):
        for p in (1, 5, 12):
            for st in (1, 4, 8):
                want = len([x for x in range(dim - p) if x == 0 or x % st == 0])
                assert lib.sr_patch_count(dim, p, st) == want


def test_packed_weight_bytes(lib):
    assert lib.sr_packed_weight_bytes(3, 128) == 9 * 128 * 128 * 2
    assert lib.sr_packed_weight_bytes(5, 128) == 25 * 128 * 128 * 2
    assert lib.sr_packed_weight_bytes(3, 3) == 9 * 128 * 16 * 2


def test_model_layer_table_is_the_keras_creation_order(lib):
    """sr_model_layer (host only): 'level1', conv2d_1..conv2d_85 as models.py:1177-1199 creates them, laid out in the
    flat arena exactly like the engine's parameter slices (kernel HWIO, then bias)."""
    from sr100.engine import layer_specs
    specs = layer_specs()
    assert lib.sr_model_num_layers() == len(specs) == 86
    name = C.create_string_buffer(16)
    k, cin, cout, wo, bo = C.c_int(), C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
    off = 0
    for i, (n, kk, ci, co) in enumerate(specs):
        assert lib.sr_model_layer(i, name, C.byref(k), C.byref(cin), C.byref(cout), C.byref(wo), C.byref(bo)) == 0
        assert (name.value.decode(), k.value, cin.value, cout.value) == (n, kk, ci, co)
        assert wo.value == off and bo.value == off + kk * kk * ci * co
        off = bo.value + co
    assert lib.sr_model_param_count() == off == 21838211
    assert lib.sr_model_layer(86, name, None, None, None, None, None) == -1
    cfg = __import__("sr100._lib", fromlist=["ModelConfig"]).ModelConfig()
    lib.sr_model_default_config(C.byref(cfg))
    assert (cfg.precision, cfg.stream_lr_fp32, cfg.nacc, cfg.pair, cfg.use_graphs) == (0, 1, 2, 1, 1)


def test_null_pointer_errors_do_not_touch_the_device(lib):
    assert lib.sr_head1x1_fwd(None, None, None, 10, None, None, None) == -1
    assert lib.sr_patch_stitch(None, 1, 1, 96, 96, 64, 4, 192, 192, 1.0, None, None, None) == -1
    assert lib.sr_score_pair_u8(None, None, 64, 64, 10, None, None) == -1
    assert lib.sr_adam_step(None, None, None, None, 4, 1e-4, 0.9, 0.999, 1e-7, 1, 1.0, None) == -1


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under image-enhance-keras_b200/ may reference it."""
    pkg = os.path.join(ROOT, "image-enhance-keras_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), os.path.join(d, f)


def test_engine_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sr100 import _lib as L
    from sr100.engine import Engine
    with pytest.raises(L.SrError):
        Engine()
    import models
    with pytest.raises(L.SrError):
        models.DifvdsrDouble(1).create_model(32, 32)
