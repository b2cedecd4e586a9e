#!/usr/bin/env python
"""Secondary measurements (not the headline bench line):
  * companions: fused softmax / layernorm fw+bw achieved HBM GB/s vs the measured copy peak
    (SURVEY.md 8d shapes; algorithmic bytes: softmax fw 8 B/elem, bw 12; layernorm fw 8, bw 12)
  * sweep: flash-attention forward TFLOP/s over seq 512..8192, head_dim 64/128, causal / full
    (BASELINE.json config #3), plus fwd+bwd at config #4 geometry with and without masks
Writes one JSON document to stdout (and to --out).  Device-resident tensors, CUDA events, L2
flushed between timed launches (companions) or tensors larger than L2 (attention)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402

c_void_p = ctypes.c_void_p


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def time_call(fn, lib, reps=10, flush=True):
    t = dev.Timer()
    best, tot = 1e30, 0.0
    for _ in range(2):
        fn()
    for _ in range(reps):
        if flush:
            lib.fa_flush_l2()
        t.start()
        fn()
        ms = t.stop()
        best = min(best, ms)
        tot += ms
    return tot / reps, best


def dev_rand(shape, seed, lo=-1.0, hi=1.0):
    """uniform fp32 tensor on device: one 64 MiB random block tiled over the tensor (bandwidth does not depend on the
    values, and gigabytes of host RNG would dominate the run)"""
    out = dev.DeviceArray(shape, "f32")
    lib = fb._lib.load("flashattention_kernel")
    n = out.size
    step = 1 << 24
    h = np.random.default_rng(seed).uniform(lo, hi, min(step, n)).astype(np.float32)
    for s in range(0, n, step):
        c = min(step, n - s)
        lib.fa_h2d(c_void_p(out.ptr + 4 * s), h.ctypes.data_as(c_void_p), c * 4)
    return out


def companions(P, quiet=False):
    sm = fb._lib.load("softmax_kernel")
    ln = fb._lib.load("layernorm_kernel")
    fl = fb._lib.load("flashattention_kernel")
    res = []
    for (B, H, F, T) in [(8, 16, 512, 512), (8, 16, 1024, 1024), (32, 16, 2048, 2048)]:
        x = dev_rand((B, H, F, T), 3)
        y = dev_rand((B, H, F, T), 4, 0.0, 1.0)
        valid = np.random.default_rng(5).integers(T // 2, T + 1, B)
        mask = dev.DeviceArray.from_numpy(np.where(np.arange(T)[None, :] < valid[:, None], 0.0, -1e8).astype(np.float32))
        n = B * H * F * T
        for name, fn, bpe in [
            ("softmax_fw_masked", lambda: sm.fa_attn_softmax_dev(x.ptr, mask.ptr, B, H, F, T, 0, None), 8),
            ("softmax_fw_future", lambda: sm.fa_attn_softmax_dev(x.ptr, None, B, H, F, T, 1, None), 8),
            ("softmax_bw", lambda: sm.fa_attn_softmax_bw_dev(x.ptr, y.ptr, B * H * F, T, None), 12),
        ]:
            avg, best = time_call(fn, fl)
            gbs = n * bpe / (avg * 1e-3) / 1e9
            res.append(dict(op=name, shape=[B, H, F, T], ms=avg, ms_best=best, gbs=gbs, frac_of_hbm_peak=gbs / P["hbm"],
                            bytes_per_elem=bpe))
        x.free(), y.free()
    for (rows, h) in [(1024, 32), (8192, 1024), (32768, 4096)]:
        x = dev_rand((rows, h), 6)
        dy = dev_rand((rows, h), 7)
        yb = dev.DeviceArray((rows, h), "f32")
        dx = dev.DeviceArray((rows, h), "f32")
        g = dev_rand((h,), 8)
        b = dev_rand((h,), 9)
        var, mean = dev.DeviceArray((rows,), "f32"), dev.DeviceArray((rows,), "f32")
        dg, db = dev.DeviceArray((h,), "f32"), dev.DeviceArray((h,), "f32")
        n = rows * h
        fw = lambda: ln.fa_layernorm_dev(yb.ptr, var.ptr, mean.ptr, x.ptr, g.ptr, b.ptr, rows, h, None)  # noqa: E731
        bw = lambda: ln.fa_layernorm_bw_dev(dg.ptr, db.ptr, dx.ptr, dy.ptr, x.ptr, g.ptr, b.ptr, var.ptr, mean.ptr,  # noqa: E731
                                            rows, h, None)
        for name, fn, bpe in [("layernorm_fw", fw, 8), ("layernorm_bw", bw, 12)]:
            avg, best = time_call(fn, fl)
            gbs = n * bpe / (avg * 1e-3) / 1e9
            res.append(dict(op=name, shape=[rows, h], ms=avg, ms_best=best, gbs=gbs, frac_of_hbm_peak=gbs / P["hbm"],
                            bytes_per_elem=bpe))
    return res


def attn_case(P, B, H, N, d, causal, kv=None, bwd=False, reps=8, quiet=False):
    fl = fb._lib.load("flashattention_kernel")
    rng = np.random.default_rng(0)
    base = dev.to_bf16_bits(rng.standard_normal((1, H, N, d)).astype(np.float32))

    def mk():
        a = dev.DeviceArray((B, H, N, d), "bf16")
        for b in range(B):
            fl.fa_h2d(c_void_p(a.ptr + b * H * N * d * 2), base.ctypes.data_as(c_void_p), base.nbytes)
        return a

    Q, K, V = mk(), mk(), mk()
    kv_len = None
    dkv = None
    if kv:
        kv_len = np.random.default_rng(2).integers(N // 2, N + 1, B).astype(np.int32)
        dkv = dev.DeviceArray.from_numpy(kv_len)
    out = (dev.DeviceArray((B, H, N, d), "bf16"), dev.DeviceArray((B, H, N), "f32"), dev.DeviceArray((B, H, N), "f32"))
    f_fwd = dev.attn_flops(B, H, N, d, causal, kv_len, False)
    flush = (4 * B * H * N * d * 2) < (256 << 20)
    avg, best = time_call(lambda: dev.flash_fwd(Q, K, V, causal=causal, kv_len=dkv, out=out), fl, reps, flush)
    r = dict(B=B, H=H, N=N, d=d, causal=bool(causal), padding=bool(kv), fwd_ms=avg, fwd_tflops=f_fwd / (avg * 1e-3) / 1e12)
    r["fwd_frac_sustained"] = r["fwd_tflops"] / P["tf_sustained"]
    r["fwd_frac_burst"] = r["fwd_tflops"] / P["tf_burst"]
    r["fwd_frac_datasheet"] = r["fwd_tflops"] / 2250.0
    if bwd:
        dO = mk()
        grads = tuple(dev.DeviceArray((B, H, N, d), "bf16") for _ in range(3))
        f_bwd = dev.attn_flops(B, H, N, d, causal, kv_len, True)
        avg, best = time_call(lambda: dev.flash_bwd(Q, K, V, out[0], dO, out[1], out[2], causal=causal, kv_len=dkv,
                                                     out=grads), fl, reps, flush)
        r.update(bwd_ms=avg, bwd_tflops=f_bwd / (avg * 1e-3) / 1e12)
        r["bwd_frac_sustained"] = r["bwd_tflops"] / P["tf_sustained"]
        r["bwd_frac_burst"] = r["bwd_tflops"] / P["tf_burst"]
        r["bwd_frac_datasheet"] = r["bwd_tflops"] / 2250.0
    if not quiet:
        print("case", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if "frac" not in k},
              file=sys.stderr, flush=True)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="companions,sweep,cfg4")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    P = peaks()
    doc = {"peaks": P}
    what = args.what.split(",")
    if "companions" in what:
        doc["companions"] = companions(P)
    if "sweep" in what:  # BASELINE.json config #3: fwd sweep, batch 8, 16 heads
        doc["fwd_sweep"] = [attn_case(P, 8, 16, N, d, c) for d in (128, 64) for c in (False, True)
                            for N in (512, 1024, 2048, 4096, 8192)]
    if "seqsweep" in what:   # the BASELINE metric itself: fwd+bwd TFLOP/s vs sequence length (32 heads, batch 8)
        doc["fwdbwd_vs_seq"] = []
        for d in (128, 64):
            for c in (False, True):
                for N in (512, 1024, 2048, 4096, 8192):
                    r = attn_case(P, 8, 32, N, d, c, bwd=True, reps=6)
                    tot = (r["fwd_ms"] + r["bwd_ms"]) * 1e-3
                    fl = r["fwd_tflops"] * r["fwd_ms"] * 1e-3 + r["bwd_tflops"] * r["bwd_ms"] * 1e-3
                    r["fwdbwd_tflops"] = fl / tot
                    r["fwdbwd_frac_sustained"] = r["fwdbwd_tflops"] / P["tf_sustained"]
                    r["fwdbwd_frac_datasheet"] = r["fwdbwd_tflops"] / 2250.0
                    doc["fwdbwd_vs_seq"].append(r)
    if "cfg4" in what:   # config #4 geometry: fwd+bwd, 32 heads, N=4096, d=128, +- causal, +- padding
        doc["cfg4"] = [attn_case(P, 8, 32, 4096, 128, c, kv, bwd=True) for c in (False, True) for kv in (False, True)]
    txt = json.dumps(doc, indent=1)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(txt)


if __name__ == "__main__":
    main()
