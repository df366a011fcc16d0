"""Launch a few representative kernels once each (after a warm-up) so that `ncu --set full -k regex:...` can capture
them: wgrad (k5 / k3 at the training shapes), the staged sub-pixel shuffle, the batched patch gather, the fused
conv + shuffle.  Dev tool for profiles/; not a benchmark.

    ncu --set full --clock-control none -k regex:'wgrad_tc_kernel|depth_to_space_tiled|patch_gather_u8_rows' \
        -o gpurun_out/r01_ncu_misc python tools/ncu_targets.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    import torch
    from sr100 import _lib as L
    from sr100 import ops
    lib = L.require_device()
    dev = "cuda"
    st = L.stream_ptr
    # wgrad at the training shapes
    ws = torch.empty(lib.sr_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)
    for k, NB, H in ((5, 32, 192), (3, 32, 192), (5, 256, 48)):
        x = (torch.randn(NB, H, H, 128, device=dev) * 0.5).to(torch.bfloat16)
        g = (torch.randn(NB, H, H, 128, device=dev) * 0.5).to(torch.bfloat16)
        dw = torch.zeros(k, k, 128, 128, device=dev)
        d = L.WgradDesc()
        d.x_bf16, d.g_bf16, d.NB, d.H, d.W, d.ksize = x.data_ptr(), g.data_ptr(), NB, H, H, k
        d.scale, d.accumulate, d.dw_hwio = 1.0, 0, dw.data_ptr()
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        plan = C.c_void_p()
        L.check(lib.sr_wgrad_plan_create(C.byref(d), C.byref(plan)))
        for _ in range(2):
            L.check(lib.sr_wgrad_plan_run(plan, st()))
        torch.cuda.synchronize()
        lib.sr_wgrad_plan_destroy(plan)
        del x, g
    # sub-pixel shuffle alone, config-3 size
    x = torch.randn(64, 339, 510, 48, device=dev)
    out = torch.empty(64, 1356, 2040, 3, device=dev)
    for _ in range(2):
        L.check(lib.sr_depth_to_space(L.ptr(x), 64, 339, 510, 3, 4, 0, L.ptr(out), st()))
    torch.cuda.synchronize()
    del x, out
    # batched gather, config-3 size
    batch = torch.randint(0, 256, (64, 339, 510, 3), dtype=torch.uint8, device=dev)
    ch, cw = ops.canvas_size(339, 510)
    po = torch.empty(64 * 54, 96, 96, 3, device=dev)
    for _ in range(2):
        ops.patch_gather_u8_batched(batch, (ch, cw), (96, 96), 64, out=po)
    torch.cuda.synchronize()
    # fused conv + shuffle
    xb = (torch.randn(296, 96, 96, 128, device=dev) * 0.5).to(torch.bfloat16)
    w = torch.randn(3, 3, 128, 48, device=dev) / 34.0
    b = torch.zeros(48, device=dev)
    for _ in range(2):
        ops.conv2d_tc_shuffle(xb, w, b, 4, 0, relu=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
