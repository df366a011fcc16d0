"""Achieved HBM bandwidth of the bandwidth-bound kernels of the hot path (algorithmic bytes =
compulsory reads + writes at the boundary dtypes), against MEASURED_PEAKS.json hbm_gbs.

    python tools/probe_bw.py [--out gpurun_out/probe_bw.jsonl]

Sizes follow the configs: LR stage of 186 tiles of 96x96 (config 2), one 339x510 image (config 3), a
1356x2040 output pair for the scoring kernel, the 21.84 M-parameter arena for Adam.  Every timed buffer set is
larger than the 126 MB L2 or rotated so that the inputs come from HBM."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe_bw.jsonl"))
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import torch
    from sr100 import _lib as L
    from sr100 import ops
    lib = L.require_device()
    peak = 6551.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    dev = "cuda"
    st = L.stream_ptr
    recs = []
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)      # 256 MB, only ever read

    def timed(name, nbytes, fn, note=""):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.iters):
            # evict L2 between timed launches: a 256 MB write, then a 256 MB READ of another buffer, so that the
            # cache holds clean lines when the timed kernel starts (after the write alone ~100 MB of dirty lines
            # were written back INSIDE the timed region: 10-15 us charged to kernels that take 20-100 us)
            flush.fill_(1)
            flush_rd.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / a.iters
        gbs = nbytes / ms / 1e6
        rec = dict(kernel=name, algorithmic_bytes=int(nbytes), ms=round(ms, 4), gbs=round(gbs, 1),
                   frac_of_hbm_peak=round(gbs / peak, 3), peak_gbs=peak, note=note)
        recs.append(rec)
        print(json.dumps(rec), flush=True)

    # ---- reference points on this box, same method: what plain torch kernels reach at the sizes probed below
    for mb in (128, 1024):
        n = mb * 1024 * 1024 // 4
        a_, b_ = torch.empty(n, device=dev), torch.empty(n, device=dev)
        timed("reference: torch copy (read + write), %d MB each way" % mb, 2 * n * 4, lambda: b_.copy_(a_))
        timed("reference: torch fill (write only), %d MB" % mb, n * 4, lambda: b_.fill_(1.0))
        timed("reference: torch sum (read only), %d MB" % mb, n * 4, lambda: a_.sum())
        del a_, b_
    # ---- bilinear x4 (fp32 LR stream in, bf16 HR out): 512 + 16*256 B per LR pixel
    NB, H, W = 186, 96, 96
    x32 = torch.randn(NB, H, W, 128, device=dev)
    hr16 = torch.empty(NB, 4 * H, 4 * W, 128, device=dev, dtype=torch.bfloat16)
    timed("bilinear4_fwd_kernel<f32->bf16>", NB * H * W * (512 + 16 * 256),
          lambda: L.check(lib.sr_bilinear4_fwd(L.ptr(x32), 0, NB, H, W, 128, L.ptr(hr16), None, st())),
          "186 tiles of 96x96x128")
    del hr16
    # ---- bilinear adjoint (fp32 HR grad in, fp32 LR grad out), 32 tiles
    NB2 = 32
    g32 = torch.randn(NB2, 4 * H, 4 * W, 128, device=dev)
    gin = torch.empty(NB2, H, W, 128, device=dev)
    timed("bilinear4_bwd_kernel", NB2 * H * W * (16 * 512 + 512),
          lambda: L.check(lib.sr_bilinear4_bwd(L.ptr(g32), NB2, H, W, 128, L.ptr(gin), st())), "32 tiles")
    del g32, gin
    # ---- head 1x1: 12 B in, 256 + 512 B out per pixel
    xin = torch.rand(NB, H, W, 3, device=dev)
    s16 = torch.empty(NB, H, W, 128, device=dev, dtype=torch.bfloat16)
    w0, b0 = torch.randn(3, 128, device=dev), torch.randn(128, device=dev)
    timed("head1x1_kernel", NB * H * W * (12 + 256 + 512),
          lambda: L.check(lib.sr_head1x1_fwd(L.ptr(xin), L.ptr(w0), L.ptr(b0), NB * H * W, L.ptr(s16), L.ptr(x32), st())))
    # ---- patch gather: 64 images of 339x510 (config 3) -> 54 tiles each
    imgs = [torch.randint(0, 256, (339, 510, 3), dtype=torch.uint8, device=dev) for _ in range(8)]
    ch, cw = ops.canvas_size(339, 510)

    def gather_all():
        for im in imgs:
            ops.patch_gather_u8(im, (ch, cw), (96, 96), 64)
    timed("patch_gather_kernel<u8> x8 images", 8 * (339 * 510 * 3 + 54 * 96 * 96 * 3 * 4), gather_all,
          "8 launches of 339x510 -> 54 tiles; includes the output allocation")
    batch = torch.randint(0, 256, (64, 339, 510, 3), dtype=torch.uint8, device=dev)
    bout64 = torch.empty(64 * 54, 96, 96, 3, device=dev)
    timed("patch_gather_u8_rows_kernel (64 x 339x510, one launch)", 64 * (339 * 510 * 3 + 54 * 96 * 96 * 3 * 4),
          lambda: ops.patch_gather_u8_batched(batch, (ch, cw), (96, 96), 64, out=bout64),
          "config 3: 64 images -> 3456 tiles in one launch, preallocated output")
    del batch, bout64
    # ---- stitch + quantise: 54 tiles of 384x384x3 fp32 -> uint8 canvas (owned pixels only are compulsory)
    outp = torch.rand(54, 384, 384, 3, device=dev)
    cnt = (ops.patch_count(ch, 96, 64), ops.patch_count(cw, 96, 64))
    timed("patch_stitch_kernel", 16 * ch * cw * 3 * (4 + 1),
          lambda: ops.patch_stitch(outp, cnt, (96, 96), 64, 4, (ch, cw), mul=255.0, want_f32=False, want_u8=True),
          "one 339x510 image (448x640 canvas x4); 15 B per output pixel")
    # ---- the same two kernels at config-5 size (one 1080x1920 image, 558 tiles): enough bytes to leave launch latency
    big = torch.randint(0, 256, (1080, 1920, 3), dtype=torch.uint8, device=dev)
    bch, bcw = ops.canvas_size(1080, 1920)
    timed("patch_gather_u8_rows_kernel (1080x1920)", 1080 * 1920 * 3 + 558 * 96 * 96 * 3 * 4,
          lambda: ops.patch_gather_u8(big, (bch, bcw), (96, 96), 64), "558 tiles; includes the output allocation")
    bout = torch.rand(558, 384, 384, 3, device=dev)
    bcnt = (ops.patch_count(bch, 96, 64), ops.patch_count(bcw, 96, 64))
    timed("patch_stitch_vec4_kernel (1080x1920 -> 4320x7680)", 16 * 1080 * 1920 * 3 * (4 + 1),
          lambda: ops.patch_stitch(bout, bcnt, (96, 96), 64, 4, (1080, 1920), mul=255.0, want_f32=False, want_u8=True),
          "stitched straight into the final 4H x 4W image; 15 B per output pixel")
    del bout
    # ---- scoring: two uint8 1356x2040 RGB images
    a8 = torch.randint(0, 256, (1356, 2040, 3), dtype=torch.uint8, device=dev)
    b8 = torch.randint(0, 256, (1356, 2040, 3), dtype=torch.uint8, device=dev)
    res = torch.zeros(128, dtype=torch.uint8, device=dev)
    timed("score_pair_kernel", 2 * 1356 * 2040 * 3,
          lambda: L.check(lib.sr_score_pair_u8(L.ptr(a8), L.ptr(b8), 1356, 2040, 10, L.ptr(res), st())),
          "one 1356x2040 pair (Y-PSNR + Y-SSIM + RGB-SSIM fused)")
    # ---- Adam over the whole parameter arena: 4 reads + 3 writes fp32
    n = 21838211
    pp, gg, mm, vv = (torch.randn(n, device=dev) for _ in range(4))
    vv.abs_()
    timed("adam_kernel", n * 28,
          lambda: L.check(lib.sr_adam_step(L.ptr(pp), L.ptr(gg), L.ptr(mm), L.ptr(vv), n, 1e-4, 0.9, 0.999, 1e-7, 3, 1.0, st())))
    # ---- training elementwise
    g16 = torch.randn(NB * H * W, 128, device=dev).to(torch.bfloat16)
    acc = torch.zeros(128, device=dev)
    timed("colsum_bf16_kernel", NB * H * W * 256,
          lambda: L.check(lib.sr_colsum_bf16(L.ptr(g16), NB * H * W, 1.0, L.ptr(acc), st())))
    npix = 32 * 384 * 384
    pred, tgt = torch.rand(npix, 3, device=dev), torch.rand(npix, 3, device=dev)
    g128 = torch.empty(npix, 128, device=dev, dtype=torch.bfloat16)
    ls = torch.zeros(1, device=dev, dtype=torch.float64)
    timed("mse_tail_grad_kernel", npix * (24 + 256),
          lambda: L.check(lib.sr_mse_tail_grad(L.ptr(pred), L.ptr(tgt), npix, 3, npix * 3, L.ptr(g128), L.ptr(ls), st())))
    g128b = torch.empty(npix, 128, device=dev, dtype=torch.bfloat16)
    db3 = torch.zeros(3, device=dev)
    timed("mse_tail_grad_col_kernel", npix * (24 + 256),
          lambda: L.check(lib.sr_mse_tail_grad_col(L.ptr(pred), L.ptr(tgt), 32, 384, 384, npix * 3, L.ptr(g128b), L.ptr(ls),
                                                   L.ptr(db3), st())), "im2col of the loss gradient + bias gradient")
    del g128, g128b, pred, tgt
    # ---- bilinear x2 (Difvdsr4)
    x2 = torch.randn(64, 96, 96, 128, device=dev)
    o2 = torch.empty(64, 192, 192, 128, device=dev, dtype=torch.bfloat16)
    timed("bilinear2_fwd_kernel<f32->bf16>", 64 * 96 * 96 * (512 + 4 * 256),
          lambda: L.check(lib.sr_bilinear2_fwd(L.ptr(x2), 0, 64, 96, 96, 128, L.ptr(o2), None, st())))
    del x2, o2
    # ---- minibatch assembly from the HBM-resident dataset: 256 pairs of 48x48 / 192x192 uint8 -> float32
    n_ds = 4096
    ds_y = torch.randint(0, 256, (n_ds, 192, 192, 3), dtype=torch.uint8, device=dev)
    idx = torch.randperm(n_ds, device=dev)[:256].to(torch.int64)
    by = torch.empty(256, 192, 192, 3, device=dev)
    timed("batch_gather_u8_kernel", 256 * 192 * 192 * 3 * (1 + 4),
          lambda: L.check(lib.sr_batch_gather_u8(L.ptr(ds_y), 192 * 192 * 3, n_ds, L.ptr(idx), 256, 255.0, L.ptr(by), st())),
          "256 HR targets of 192x192: 1 B read + 4 B written per element")
    del ds_y, by
    # ---- alternative tilers: x4 bicubic shrink of every 4th 128-px patch of a 512x512 image, averaging stitch
    from sr100 import alt_tilers as at
    im = torch.randint(0, 256, (512, 512, 3), dtype=torch.uint8, device=dev)
    cnt_p = (512 - 128) // 4 + 1
    bnd, kk = at.pil_bicubic_coeffs(128, 32)
    co = (torch.from_numpy(bnd).to(dev), torch.from_numpy(kk).to(dev))
    timed("patch_down4_kernel", cnt_p * cnt_p * (128 * 128 * 3 + 32 * 32 * 3 * 4),
          lambda: at.patch_down4(im, 128, 4, True, None, None, co),
          "9409 patches of 128x128 (bytescale + Pillow fixed-point bicubic): integer compute, patch reads hit L2")
    rows = 48
    pv = torch.rand(rows * cnt_p, 128, 128, 3, device=dev)
    accd = torch.zeros(512, 512, 3, device=dev, dtype=torch.float64)
    cntd = torch.zeros(512, 512, device=dev, dtype=torch.int32)
    timed("patch_average_accumulate_kernel", rows * cnt_p * 128 * 128 * 3 * 4,
          lambda: L.check(lib.sr_patch_average_accumulate(L.ptr(pv), 128, 4, 4, cnt_p, cnt_p, 0, rows, cnt_p - 1, cnt_p - 1,
                                                          255.0, 512, 512, L.ptr(accd), L.ptr(cntd), st())),
          "48 grid rows x 97 patches of 128x128x3 fp32 (0.9 GB) read once, float64 sums in patch order")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
