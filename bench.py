#!/usr/bin/env python
"""Benchmark of the fused-attention hot path (BASELINE.json metric: attention fwd+bwd TFLOP/s).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU arm (rank 0 only)

Workload (config.workload): cfg4 of SURVEY.md 8(d) -- flash-attention fwd+bwd training step,
B=8 per GPU, 32 heads, seq 4096, head_dim 128, bf16, key-padding masks (kv_len[b] in [N/2, N]).
A "step" = one forward + one backward over the batch.  Each GPU owns its own batch slice
(batch x head sharding, no collective on the data path) -> weak scaling; `value` = algorithmic
FLOPs of all ranks / max-over-ranks device time.

Timing: CUDA events on the launching (default) stream, W warm-up steps, K timed steps bracketed
by barrier + device sync; inputs (8 x 268 MB) exceed the 126 MB L2, so no explicit flush.
`e2e` = the same step through the reference-facing legacy C ABI (host fp32 buffers in pinned
memory, H2D/D2H inside the timed region).  `cpu_baseline` = the reference's composed numba CPU
path (oracle port) on a bounded sample, rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, H, N, d, causal)
    "cfg4": dict(B=8, H=32, N=4096, d=128, causal=False, padding=True,
                 desc="flash-attention fwd+bwd training step, B=8/GPU H=32 N=4096 d=128 bf16, key padding kv_len in [N/2,N]"),
    "cfg4_causal": dict(B=8, H=32, N=4096, d=128, causal=True, padding=True,
                        desc="cfg4 with a causal mask on top of key padding"),
    # BASELINE config #5: batch 64 x 32 heads x seq 8192 x d 128, sharded by batch over the GPUs (strong scaling):
    # B here is the GLOBAL batch; each rank takes B / world_size of it
    "cfg5": dict(B=64, H=32, N=8192, d=128, causal=False, padding=False, strong=True,
                 desc="batch x head-sharded attention fwd+bwd, global B=64 H=32 N=8192 d=128 bf16 (config #5)"),
    "cfg5_causal": dict(B=64, H=32, N=8192, d=128, causal=True, padding=False, strong=True,
                        desc="config #5 with a causal mask"),
    "small": dict(B=2, H=4, N=1024, d=128, causal=False, padding=True, desc="smoke-sized workload"),
}
CPU_SAMPLE = dict(B=2, H=32, N=512, d=128)  # bounded sample of the same op for the CPU arms


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_kernel_shares.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_traffic_bytes_per_launch"][kernel])
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)")
    return dict(tflops=1590.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Number of samples seen so far (call at the start of the timed region)."""
        return len(self.rows)

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[first:]:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off (if the cpuset allows), so the pinned host buffers
    of the end-to-end leg are first-touched on that node and the DMA does not cross the socket interconnect.
    Returns a short description for the JSON line (None when nothing was changed)."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True, timeout=20).strip().lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed or allowed == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, allowed)
        return f"node{node}:{len(allowed)}cpus"
    except Exception:
        return None


def _agg(dist):
    from flashattn_b200.sharding import Aggregator
    return Aggregator(dist, "cuda")


def barrier(dist):
    _agg(dist).barrier()


def reduce_max(dist, x):
    return _agg(dist).max(x)


def reduce_sum(dist, x):
    return _agg(dist).sum(x)


def cpu_reference_run(steps, warmup, cores=None):
    """The reference's composed attention on the host cores (oracle port), bounded sample."""
    import numba
    from oracle import numba_composed as NC
    if cores:
        numba.set_num_threads(min(cores, numba.config.NUMBA_NUM_THREADS))
    ncores = numba.get_num_threads()
    c = CPU_SAMPLE
    rng = np.random.default_rng(0)
    Q, K, V, dO = (rng.standard_normal((c["B"], c["H"], c["N"], c["d"])).astype(np.float32) for _ in range(4))
    for _ in range(max(1, warmup)):  # includes the numba JIT
        NC.attention_fwd_bwd(Q, K, V, dO)
    t0 = time.perf_counter()
    for _ in range(steps):
        NC.attention_fwd_bwd(Q, K, V, dO)
    dt = (time.perf_counter() - t0) / steps
    flops = 14.0 * c["B"] * c["H"] * c["N"] * c["N"] * c["d"]
    return dict(value=flops / dt / 1e12, unit="TFLOP/s", cores=int(ncores), kind="port",
                sample=f"composed attention fwd+bwd fp32, B={c['B']} H={c['H']} N={c['N']} d={c['d']}, "
                       f"{steps} runs, {dt * 1e3:.0f} ms each"), dt


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    cb, dt = cpu_reference_run(steps, min(args.warmup, 1) or 1)
    w = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "attention fwd+bwd TFLOP/s", "value": cb["value"], "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample": cb["sample"],
                   "note": "reference's numba CPU composed attention (oracle port: oracle/numba_composed.py); "
                           "the reference's CUDA flash kernel does not compile (SURVEY.md 2.3)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    args.warmup = max(args.warmup, 3)

    rank, world, local, dist = dist_setup(args.gpus)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    import flashattn_b200 as fb
    from flashattn_b200 import device as dev
    lib = fb._lib.load("flashattention_kernel")
    if lib.fa_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    fb._lib.check(lib, lib.fa_set_device(local))
    peaks = load_peaks()

    w = WORKLOADS[args.workload]
    B, H, N, d, causal = w["B"], w["H"], w["N"], w["d"], w["causal"]
    if w.get("strong"):             # strong scaling: the global batch is split across the ranks
        from flashattn_b200.sharding import shard_batch
        b0, b1 = shard_batch(B, world, rank)
        B = max(1, b1 - b0)
    rng = np.random.default_rng(1000 + rank)
    kv_len = rng.integers(N // 2, N + 1, B).astype(np.int32) if w["padding"] else None

    # ---- synthetic device-resident inputs (bf16), generated per (b) slice to bound host memory
    def dev_randn(seed):
        out = dev.DeviceArray((B, H, N, d), "bf16")
        r = np.random.default_rng(seed)
        per = H * N * d
        for b in range(B):
            chunk = dev.to_bf16_bits(r.standard_normal(per, dtype=np.float32))
            fb._lib.check(lib, lib.fa_h2d(ctypes.c_void_p(out.ptr + b * per * 2), chunk.ctypes.data_as(ctypes.c_void_p),
                                          per * 2))
        return out

    Q, K, V, dO = (dev_randn(10 * rank + i) for i in range(4))
    dkv = dev.DeviceArray.from_numpy(kv_len) if kv_len is not None else None
    O = dev.DeviceArray((B, H, N, d), "bf16")
    m, l = dev.DeviceArray((B, H, N), "f32"), dev.DeviceArray((B, H, N), "f32")
    grads = tuple(dev.DeviceArray((B, H, N, d), "bf16") for _ in range(3))

    flops_f = dev.attn_flops(B, H, N, d, causal, kv_len, backward=False)
    flops_b = dev.attn_flops(B, H, N, d, causal, kv_len, backward=True)
    flops_nominal = 14.0 * B * H * N * N * d * (0.5 if causal else 1.0)

    def step():
        dev.flash_fwd(Q, K, V, causal=causal, kv_len=dkv, out=(O, m, l))
        dev.flash_bwd(Q, K, V, O, dO, m, l, causal=causal, kv_len=dkv, out=grads)

    sampler = ClockSampler(local)
    sampler.start()           # started before the warm-up: nvidia-smi needs ~1 s to emit its first line
    for _ in range(args.warmup):
        step()
    dev.sync()

    # ---- timed region: K steps, events around every fwd and bwd
    ev = [[lib.fa_event_create() for _ in range(3)] for _ in range(args.steps)]
    launches0 = lib.fa_launch_count()
    barrier(dist)
    dev.sync()
    mark = sampler.mark()
    for i in range(args.steps):
        lib.fa_event_record(ev[i][0], None)
        dev.flash_fwd(Q, K, V, causal=causal, kv_len=dkv, out=(O, m, l))
        lib.fa_event_record(ev[i][1], None)
        dev.flash_bwd(Q, K, V, O, dO, m, l, causal=causal, kv_len=dkv, out=grads)
        lib.fa_event_record(ev[i][2], None)
    dev.sync()
    barrier(dist)
    launches = int(lib.fa_launch_count() - launches0)
    probe = 0
    t_probe = time.perf_counter()
    while sampler.mark() - mark < 4 and time.perf_counter() - t_probe < 3.0:
        step()          # short timed regions: keep the same load running (untimed) until nvidia-smi has sampled it
        dev.sync()
        probe += 1
    clocks = sampler.stop(mark)
    clocks["untimed_probe_steps"] = probe
    total_ms = lib.fa_event_elapsed_ms(ev[0][0], ev[-1][2])
    fwd_ms = float(np.mean([lib.fa_event_elapsed_ms(e[0], e[1]) for e in ev]))
    bwd_ms = float(np.mean([lib.fa_event_elapsed_ms(e[1], e[2]) for e in ev]))
    total_ms = reduce_max(dist, total_ms)
    all_flops = reduce_sum(dist, (flops_f + flops_b) * args.steps)
    ms_per_step = total_ms / args.steps
    value = all_flops / (total_ms * 1e-3) / 1e12

    # ---- end to end through the legacy (reference-facing) host-pointer ABI
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(fb, lib, B, H, N, d, causal, kv_len, flops_f + flops_b, min(args.steps, 3), dist)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline, _ = cpu_reference_run(3, 1)

    if rank == 0:
        fwd_tf = flops_f / (fwd_ms * 1e-3) / 1e12
        bwd_tf = flops_b / (bwd_ms * 1e-3) / 1e12
        dom = "bwd" if bwd_ms >= fwd_ms else "fwd"
        dom_tf = bwd_tf if dom == "bwd" else fwd_tf
        line = {
            "metric": "attention fwd+bwd TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if w.get("strong") else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": w["desc"], "per_gpu": {"B": B, "H": H, "N": N, "d": d, "causal": causal},
                       "parallelism": f"batch x head shards over {world} GPU(s), no collective",
                       "flops": "effective (key padding excluded): fwd 4*H*d*sum_b N*kv_len[b], bwd 10*...",
                       "nominal_tflops": flops_nominal * world / (ms_per_step * 1e-3) / 1e12,
                       "l2": "inputs larger than L2 (8 tensors x 268 MB per GPU); no flush needed"},
            "roofline": {"bound": "tensor", "kernel": f"sm100::{dom}_kernel (tcgen05/TMEM), timed with its pre/post passes",
                         "achieved": dom_tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": dom_tf / peaks["tflops"],
                         "traffic": ncu_traffic(f"{dom}_kernel") if args.workload == "cfg4" else None,
                         "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01_kernel_shares.json)",
                         "peak_source": peaks["source"]},
            "kernels": {"fwd_ms": fwd_ms, "fwd_tflops": fwd_tf, "fwd_frac_measured_sustained": fwd_tf / peaks["tflops"],
                        "fwd_frac_measured_burst": fwd_tf / peaks["burst"], "fwd_frac_datasheet_2250": fwd_tf / 2250.0,
                        "bwd_ms": bwd_ms, "bwd_tflops": bwd_tf, "bwd_frac_measured_sustained": bwd_tf / peaks["tflops"],
                        "bwd_frac_measured_burst": bwd_tf / peaks["burst"], "bwd_frac_datasheet_2250": bwd_tf / 2250.0},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
            "numa_binding_rank0": numa,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_e2e(fb, lib, B, H, N, d, causal, kv_len, flops_step, steps, dist):
    """Same step through launch_flashattention_{forward,backward}_masked (host fp32 buffers, pinned)."""
    # 8 pinned fp32 tensors per rank (4.3 GB at cfg4): with several ranks on one host, keep the total pinned
    # footprint under ~40 % of the available host memory by shrinking this rank's e2e batch if necessary.
    from flashattn_b200 import device as dev
    world = int(os.environ.get("WORLD_SIZE", "1"))
    try:
        avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        per_b = 8 * H * N * d * 4 + 2 * H * N * 4
        B = max(1, min(B, int(0.4 * avail / world / per_b)))
    except Exception:
        pass
    if kv_len is not None:
        kv_len = kv_len[:B]
    flops_step = dev.attn_flops(B, H, N, d, causal, kv_len, backward=False) + \
        dev.attn_flops(B, H, N, d, causal, kv_len, backward=True)
    n = B * H * N * d
    r = B * H * N

    def pinned(count):
        if os.environ.get("BENCH_E2E_PAGEABLE") == "1":     # what a numpy-backed minitorch tensor hands over
            return None, np.empty(count, dtype=np.float32)
        p = lib.fa_malloc_host(count * 4)
        if not p:
            raise MemoryError("pinned allocation failed")
        return p, np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(count,))

    bufs = {k: pinned(n) for k in ("Q", "K", "V", "O", "dO", "dQ", "dK", "dV")}
    stats = {k: pinned(r) for k in ("l", "m")}
    rng = np.random.default_rng(5)
    for k in ("Q", "K", "V", "dO"):
        a = bufs[k][1]
        step_ = 1 << 24
        for s in range(0, n, step_):
            a[s:s + step_] = rng.standard_normal(min(step_, n - s), dtype=np.float32)
    mask = None
    mptr = None
    if kv_len is not None:
        mask = np.where(np.arange(N)[None, :] < kv_len[:, None], 0.0, -1e8).astype(np.float32)
        mptr = mask.ctypes.data_as(ctypes.c_void_p)
    lib.fa_set_mode(fb._lib.FA_MODE_BF16)
    A = {k: v[1] for k, v in bufs.items()}
    S = {k: v[1] for k, v in stats.items()}

    def step():
        lib.launch_flashattention_forward_masked(A["Q"], A["K"], A["V"], A["O"], S["l"], S["m"], mptr, int(causal),
                                                 B, H, N, d)
        fb._lib.check(lib)
        lib.launch_flashattention_backward_masked(A["Q"], A["K"], A["V"], A["O"], A["dQ"], A["dK"], A["dV"], A["dO"],
                                                  S["l"], S["m"], mptr, int(causal), B, H, N, d)
        fb._lib.check(lib)
        return float(A["dQ"][0])  # the step's result is read on the host

    step()
    barrier(dist)
    lib.fa_sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    lib.fa_sync()
    barrier(dist)
    dt = reduce_max(dist, time.perf_counter() - t0)
    total = reduce_sum(dist, flops_step * steps)
    lib.fa_set_mode(fb._lib.FA_MODE_FP32)
    for p, _ in list(bufs.values()) + list(stats.values()):
        if p:
            lib.fa_free_host(p)
    h2d = (3 * n + 5 * n + 2 * r) * 4 + (mask.nbytes * 2 if mask is not None else 0)   # fwd: Q,K,V; bwd: Q,K,V,O,dO,m,l
    d2h = (n + 2 * r + 3 * n) * 4
    return {"value": total / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "ms_per_step": dt / steps * 1e3, "batch_per_gpu": B,
            "path": "launch_flashattention_forward_masked + launch_flashattention_backward_masked (legacy C ABI, "
                    "fp32 %s host buffers, FA_MODE_BF16), wall clock incl. H2D/D2H" % (
                        "pageable" if os.environ.get("BENCH_E2E_PAGEABLE") == "1" else "pinned")}


if __name__ == "__main__":
    main()
