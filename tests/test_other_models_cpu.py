"""CPU: the restated Difvdsr4 / Difvdsr graphs (oracle/other_models.py) and their product-side layer tables agree on
names, shapes and parameter counts with the reference source (models.py:992-1142, 1274-1357)."""
import numpy as np

from oracle import other_models as om


def test_layer_tables_match_between_oracle_and_product():
    from sr100 import planenet as pn
    assert pn.difvdsr4_specs() == om.difvdsr4_specs()
    assert pn.difvdsr_specs() == om.difvdsr_specs()
    s4 = om.difvdsr4_specs()
    # level1 (1x1, 3->256) + 2 convs x (6 + 20 + 6) blocks + tail (models.py:1024-1047)
    assert len(s4) == 66 and s4[0] == ("level1", 1, 3, 256) and s4[-1] == ("conv2d_65", 3, 256, 3)
    assert sum(k * k * ci * co + co for _, k, ci, co in s4) == 1024 + 64 * (9 * 256 * 256 + 256) + 9 * 256 * 3 + 3
    sd = om.difvdsr_specs()
    # level1 (3x3, 3->192) + 4 convs x 32 blocks + tail (models.py:1304-1308)
    assert len(sd) == 130 and sd[0] == ("level1", 3, 3, 192) and sd[-1] == ("conv2d_129", 3, 192, 3)


def test_oracle_graphs_run_and_scale():
    rng = np.random.default_rng(0)
    x = rng.random((1, 6, 5, 3)).astype(np.float32)
    w4 = om.init_weights(om.difvdsr4_specs(), seed=1, bias_scale=0.05)
    y4 = om.forward_difvdsr4(w4, x)
    assert y4.shape == (1, 24, 20, 3) and np.isfinite(y4).all() and (y4 >= 0).all()
    wd = om.init_weights(om.difvdsr_specs(), seed=1, bias_scale=0.05)
    yd = om.forward_difvdsr(wd, x)
    assert yd.shape == (1, 6, 5, 3) and np.isfinite(yd).all() and (yd >= 0).all()


def test_bilinear_tf1_x2_known_values():
    import torch
    x = torch.tensor([[0.0, 2.0, 6.0]]).view(1, 1, 1, 3)
    y = om.bilinear_tf1(x, 2)
    assert y.shape == (1, 1, 2, 6)
    assert y[0, 0, 0].tolist() == y[0, 0, 1].tolist() == [0.0, 1.0, 2.0, 4.0, 6.0, 6.0]        # legacy sampling: src = dst/2, the last sample replicates the edge


def test_constructors_import_without_gpu():
    import models
    m4, md = models.Difvdsr4(2), models.Difvdsr(1)
    assert m4.weight_path.startswith("weights_Difvdsr2scale/") and md.weight_path.startswith("weights_Difvdsr/")
    assert m4.scale_factor == 2 and md.model_name == "Image ScaleGen"
