"""Sub-pixel shuffle measurements (SURVEY.md 8d "pixel shuffle ... for the A/B vs fused"):

  * sr_depth_to_space alone (the unfused reference form: keras_subpixel.py:64-84, advanced.py:87-129,195-196):
    algorithmic bytes = read + write of the tensor once, against MEASURED_PEAKS.json hbm_gbs;
  * the tensor-core conv with the shuffle fused into its epilogue (sr_conv_desc.shuffle_r), next to the same conv
    storing the unshuffled fp32 tensor followed by the shuffle pass -- what fusing saves.

    python tools/probe_shuffle.py [--out gpurun_out/probe_shuffle.jsonl]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe_shuffle.jsonl"))
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    import torch
    from sr100 import _lib as L
    lib = L.require_device()
    peak = 6551.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    dev = "cuda"
    recs = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed(fn, flush_l2=True):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.iters):
            if flush_l2:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.iters

    def emit(**rec):
        recs.append(rec)
        print(json.dumps(rec), flush=True)

    # ---- the shuffle alone
    for name, NB, H, W, Cc, r in (("config3_lr_64x339x510_r4_C3", 64, 339, 510, 3, 4),
                                  ("tiles_186x96x96_r4_C3", 186, 96, 96, 3, 4),
                                  ("tiles_64x96x96_r2_C32", 64, 96, 96, 32, 2),
                                  ("one_image_339x510_r4_C3", 1, 339, 510, 3, 4)):
        x = torch.randn(NB, H, W, Cc * r * r, device=dev)
        out = torch.empty(NB, H * r, W * r, Cc, device=dev)
        nbytes = 2 * x.numel() * 4
        for order in (0, 1, 2):
            ms = timed(lambda: L.check(lib.sr_depth_to_space(L.ptr(x), NB, H, W, Cc, r, order, L.ptr(out), L.stream_ptr())))
            emit(kernel="depth_to_space_tiled_kernel", case=name, order=order, algorithmic_bytes=nbytes, ms=round(ms, 4),
                 gbs=round(nbytes / ms / 1e6, 1), frac_of_hbm_peak=round(nbytes / ms / 1e6 / peak, 3), peak_gbs=peak)
        del x, out

    # ---- fused conv + shuffle vs conv -> fp32 tensor -> shuffle pass
    for name, NB, H, W, k, r, Cc in (("subpixel_k3_128to48_r4", 296, 96, 96, 3, 4, 3),
                                     ("subpixel_k5_128to128_r2", 296, 96, 96, 5, 2, 32)):
        cout = Cc * r * r
        x = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16)
        w = torch.randn(k, k, 128, cout, device=dev) / (k * k * 128) ** 0.5
        b = torch.randn(cout, device=dev) * 0.1
        packed = torch.empty(lib.sr_packed_weight_bytes(k, cout), dtype=torch.uint8, device=dev)
        L.check(lib.sr_pack_conv_weights(L.ptr(w), k, cout, 0, L.ptr(packed), L.stream_ptr()))
        out = torch.empty(NB, H * r, W * r, Cc, device=dev)

        def plan(shuffle, dst, w_packed, co):
            d = L.ConvDesc()
            d.nsrc = 1
            d.in_[0], d.wpacked[0], d.ksize[0] = x.data_ptr(), w_packed.data_ptr(), k
            d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, co
            d.bias = b128.data_ptr() if co == 128 else b.data_ptr()
            d.alpha, d.beta, d.relu = 1.0, 0.0, 1
            d.out_f32 = dst.data_ptr()
            d.a_mode, d.nacc, d.pair = 0, 2, 1
            if shuffle:
                d.shuffle_r, d.shuffle_order = r, 0
            h = C.c_void_p()
            L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(h)))
            return h

        b128 = torch.zeros(128, device=dev)
        b128[:cout] = b
        fused = plan(True, out, packed, cout)
        ms_f = timed(lambda: L.check(lib.sr_conv_plan_run(fused, L.stream_ptr())), flush_l2=False)
        flops = 2.0 * NB * H * W * k * k * 128 * cout
        emit(kernel="conv_tc_pair_kernel<2,2,4> (fused shuffle)", case=name, ms=round(ms_f, 4),
             algorithmic_tflops=round(flops / ms_f / 1e9, 1), executed_tflops=round(flops * 128 / cout / ms_f / 1e9, 1),
             out_bytes=out.numel() * 4, out_gbs=round(out.numel() * 4 / ms_f / 1e6, 1))
        # unfused: the dense 128-wide launch (zero-padded weights) storing fp32 [N,H,W,128], then the shuffle of its
        # first cout channels (for cout == 128 exactly the reference's two steps)
        w128 = torch.zeros(k, k, 128, 128, device=dev)
        w128[..., :cout] = w
        packed128 = torch.empty(lib.sr_packed_weight_bytes(k, 128), dtype=torch.uint8, device=dev)
        L.check(lib.sr_pack_conv_weights(L.ptr(w128), k, 128, 0, L.ptr(packed128), L.stream_ptr()))
        dense = torch.empty(NB, H, W, 128, device=dev)
        unf = plan(False, dense, packed128, 128)
        ms_c = timed(lambda: L.check(lib.sr_conv_plan_run(unf, L.stream_ptr())), flush_l2=False)
        mid = torch.randn(NB, H, W, cout, device=dev)
        ms_s = timed(lambda: L.check(lib.sr_depth_to_space(L.ptr(mid), NB, H, W, Cc, r, 0, L.ptr(out), L.stream_ptr())))
        emit(kernel="unfused: conv (fp32 [N,H,W,128]) + depth_to_space", case=name, conv_ms=round(ms_c, 4),
             shuffle_ms=round(ms_s, 4), total_ms=round(ms_c + ms_s, 4), fused_ms=round(ms_f, 4),
             fused_speedup=round((ms_c + ms_s) / ms_f, 3),
             hbm_bytes_avoided=2 * mid.numel() * 4)
        lib.sr_conv_plan_destroy(fused)
        lib.sr_conv_plan_destroy(unf)
        del x, out, dense, mid
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for r_ in recs:
            f.write(json.dumps(r_) + "\n")


if __name__ == "__main__":
    main()
