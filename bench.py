#!/usr/bin/env python
"""Benchmark of the encode+decode transform path (BASELINE.json metric: images/s at 512^2, patch 14).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the whole path over one batch of synthetic images:
  RGB -> IPT -> truncated 2-D DCT -> token grid -> score / sort / pack -> PatchNorm -> LFQ
  -> inverse PatchNorm -> un-patchify -> truncated IDCT -> IPT -> RGB
(BASELINE config 2: B=256 RGB fp32 512x512, patch 14, max 32x32 tiles, beta=0, max_seq_len 3072,
LFQ 14 codebooks x 14 bits).  Multi-GPU (torchrun, one rank per GPU): the batch is sharded by image,
weak scaling, no data-path collective; the PatchNorm statistic fit (with its all-reduce) runs once
before the timed region.

Prints ONE JSON line (rank 0).  `--impl reference` times the oracle port of the reference's CPU
algorithm on the host cores instead (bounded sample per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encode+decode images/sec at 512^2 patch14"
UNIT = "images/s"
WORKLOAD = dict(workload="config2: synthetic 512x512 RGB fp32, patch 14, max_patch 32x32, beta=0, "
                         "max_seq_len 3072, PatchNorm frozen, LFQ 14x14bit, decode to RGB",
                image_size=512, patch_size=14)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--chunk", type=int, default=32, help="images per chunk of the host streaming round trip")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dct-impl", default="tc", choices=["tc", "tc_plain", "fp32"],
                    help="tc = tcgen05 split-precision GEMMs (default), fp32 = exact FFMA GEMMs")
    return ap.parse_args()


# ----------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """Runs the oracle pipeline on a chunk of images in one process."""
    seed, n, size = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import dcta_oracle as O
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3, size, size), dtype=np.float32)
    fe = O.FeatureExtractor(3, 14, 0.0, 32, 32, 3072)
    pn = O.PatchNorm(32, 32, 14, 3)
    pn.frozen = True      # identity-like tables: same arithmetic per token as fitted ones
    lfq = O.LFQ(codebook_size=2 ** 14, num_codebooks=14)
    t0 = time.perf_counter()
    rec, codes = O.run_pipeline(x, fe, pn, lfq)
    return time.perf_counter() - t0, float(rec[0].sum())


def cpu_reference_throughput(n_images: int, size: int, procs: int):
    """images/s of the oracle port over `procs` worker processes (one image chunk each)."""
    import multiprocessing as mp
    procs = max(1, min(procs, n_images))
    chunks = [n_images // procs + (1 if i < n_images % procs else 0) for i in range(procs)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(0, 1, 64)] * procs)            # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i + 1, c, size) for i, c in enumerate(chunks)])
        dt = time.perf_counter() - t0
    return n_images / dt, dt, procs


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = a.cpu_sample or 16 * cores
    # bounded: about (warmup + steps) * per_step images in total
    times = []
    for i in range(a.warmup + a.steps):
        ips, dt, procs = cpu_reference_throughput(per_step, a.size, cores)
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = per_step * len(times) / total
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(WORKLOAD, global_batch=per_step),
                cpu_baseline=dict(value=value, unit=UNIT, cores=procs, kind="port",
                                  sample=f"{per_step} images of 512x512 per step over {procs} worker processes "
                                         "(oracle/dcta_oracle.py run_pipeline)"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            t0 = time.time()          # nvidia-smi needs a moment to start: wait for its first line
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def mark(self):
        """Number of samples taken so far (samples after this belong to the timed region)."""
        try:
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop(self, skip: int = 0):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for i, ln in enumerate(open(self.path)):
                if i < skip:
                    continue
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    import dct_autoencoder_b200 as D
    from dct_autoencoder_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    B, S = a.batch, a.size
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072, dct_impl=a.dct_impl)
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(dev).eval()
    pipe = D.TransformPipeline(fe, pn, lfq)

    # PatchNorm statistics fitted on a different batch (all-reduced over the ranks when world > 1)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    fit_x = torch.rand(min(B, 64), 3, S, S, device=dev, generator=g)
    t0 = time.perf_counter()
    pipe.fit_norm(fit_x)
    torch.cuda.synchronize()
    fit_ms = 1e3 * (time.perf_counter() - t0)
    del fit_x

    g.manual_seed(rank)
    x = torch.rand(B, 3, S, S, device=dev, generator=g)        # 805 MB at B=256: larger than the 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- device-resident throughput (value)
    def step():
        rec, codes = pipe.roundtrip(x)
        return rec, codes

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    skip = sampler.mark() if sampler else 0
    l0 = _lib.launch_count
    ms = timed(step, a.steps, 0)
    launches = _lib.launch_count - l0
    if sampler and ms < 400:       # keep the GPU under the same load until a few samples exist
        t_end = time.time() + 0.5
        while time.time() < t_end:
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop(skip) if sampler else None
    value = world * B * a.steps / (ms / 1e3)
    mode = ("fused: PatchNorm + LFQ code words formed in the forward DCT epilogue, decode straight from code words "
            "(bit-identical to staged)") if pipe.fusable() else "staged"

    # ---- the same job through the drop-in modules one by one (every intermediate materialised)
    staged_steps = max(2, min(a.steps, 5))
    ms_staged = timed(lambda: pipe.roundtrip_staged(x), staged_steps, 1)
    staged_value = world * B * staged_steps / (ms_staged / 1e3)

    # ---- end to end through the public API with HOST buffers (pinned in, pinned out)
    # pinned buffers live on the NUMA node of the allocating thread: allocate next to the GPU (no-op on one node)
    old_affinity = os.sched_getaffinity(0)
    numa_bound = D.util.bind_to_gpu_numa(dev)
    hx = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
    hx.copy_(x)
    h_rec = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
    h_codes = torch.empty((B, 3072, 14), dtype=torch.int64).pin_memory()

    def step_e2e():
        # public host-to-host call: pinned host images in, pinned host images + codes out;
        # H2D, kernels and D2H of successive 32-image chunks overlap on three streams
        pipe.roundtrip_host(hx, h_rec, h_codes, chunk=a.chunk)

    e2e_steps = max(2, min(a.steps, 8))
    ms_e2e = timed(step_e2e, e2e_steps, 2)
    e2e_value = world * B * e2e_steps / (ms_e2e / 1e3)
    if numa_bound:
        os.sched_setaffinity(0, old_affinity)

    # ---- per-stage device times (CUDA events around each public call) for the roofline object
    stages = stage_times(torch, D, pipe, x, dev, reps=3)
    K = min(S // 14, 32) * 14
    flops_fwd = 3 * 2 * S * K * (S + K) * B          # C*2*H*K*(W+K) per image
    dct_ms = stages["dct_fwd"]
    achieved = flops_fwd / (dct_ms / 1e3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback"
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    if fe.dct_impl == "tc" and D.util.fold_ok(S, S, K, K):
        # fold_gemm_kernel: 4 launches per step (forward passes 1-2, inverse passes 1-2).  After the fold the
        # kernel is bounded by HBM (its operands and outputs stream once), not by the tensor pipe.
        # Algorithmic bytes per launch: fp16 hi/lo quadrants in, next operand / fp32 tiles out.
        Kq, Sq = K // 2, S // 2
        xq = 3 * 4 * Sq * Sq * 4            # image quadrants, hi + lo fp16            (= 3*S*S*4)
        pt = 3 * 2 * K * Sq * 4             # forward intermediate P^T [a][plane][kw][h/2], hi + lo
        yt = 3 * K * K * 4                  # fp32 token grid / coefficient quadrants hi + lo (same size)
        qt = 3 * 4 * Sq * Kq * 4            # inverse intermediate Q^T, hi + lo
        zq = 3 * 4 * Sq * Sq * 4            # fp32 quadrant transforms
        alg = [xq + pt, pt + yt, yt + qt, qt + zq]
        fwd_key = "dct_fwd"
        if "dct_fwd_codes" in stages:
            # the timed (fused) step runs forward pass 2 as fold_codes_kernel: P^T in, 14 code words (int32) + the
            # maximum per token out instead of the fp32 token grid
            n_tok = 3 * (K // 14) * (K // 14)
            alg[1] = pt + n_tok * (14 * 4 + 4)
            fwd_key = "dct_fwd_codes"
        gemm_ms = stages[fwd_key] + stages["dct_inv"]
        achieved_gbs = sum(alg) * B / (gemm_ms / 1e3) / 1e9
        tens = (flops_fwd / (stages[fwd_key] / 1e3) / 1e12)
        roofline = dict(bound="hbm", kernel="fold_gemm_kernel x3 + fold_codes_kernel (the 4 DCT GEMM launches of the timed step: "
                        "forward pass 1, forward pass 2 -> code words, inverse passes 1-2; "
                        "cta_group::2 tcgen05 fp16x3 split precision, folded basis resident in shared memory)",
                        achieved=achieved_gbs, peak=hbm, unit="GB/s", frac=achieved_gbs / hbm,
                        # dram__bytes_read + write per launch, ncu --set full of this workload (profiles/r02v_kernels_ncu_full_b256.txt:
                        # 1470 MB forward pass 1, 821 MB fold_codes_kernel, 1272 / 1459 MB inverse passes)
                        traffic=(1.256e9 if (B == 256 and S == 512 and "dct_fwd_codes" in stages) else None),
                        launches_per_step=4, avg_launch_ms=gemm_ms / 4,
                        algorithmic_bytes_per_launch=sum(alg) * B / 4,
                        peak_source=("measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback"),
                        tensor=dict(achieved=tens, peak=peak_tf, unit="TFLOP/s", frac=tens / peak_tf,
                                    note="forward passes: SURVEY 8(d) flops 3*2*H*K*(W+K) per image (plain basis GEMMs); the "
                                         "folded kernel executes 1/2 of them, each as 3 tensor MMAs, so frac <= 2/3"),
                        note="achieved = algorithmic operand + output bytes of the 4 launches / their CUDA-event time; "
                             "traffic: see profiles/ (ncu dram bytes per launch)")
    else:
        if fe.dct_impl in ("tc", "tc_plain"):
            kname = "gemm_split_kernel (forward DCT: 2 launches, tcgen05 fp16x3 split precision)"
            note = ("algorithmic flops 3*2*H*K*(W+K) per image; the kernel executes 3x that many tensor flops "
                    "(hi*hi + hi*lo + lo*hi), so frac <= 1/3 by construction")
        else:
            kname = "sgemm_tile_kernel (forward DCT: 2 launches, exact-fp32 FFMA)"
            note = "algorithmic flops 3*2*H*K*(W+K) per image"
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak_tf, unit="TFLOP/s",
                        frac=achieved / peak_tf, traffic=None, peak_source=peak_src, note=note)
    staged_bytes = 26004480 * B
    pipeline_hbm = dict(bound="hbm", mode="staged API (roundtrip_staged)",
                        achieved=staged_bytes * staged_steps / (ms_staged / 1e3) / 1e9, peak=hbm, unit="GB/s",
                        note="whole step, staged-API algorithmic bytes 26,004,480 B/img (SURVEY 8d)")
    pipeline_hbm["frac"] = pipeline_hbm["achieved"] / hbm
    fused_bytes = 6736896 * B * world
    pipeline_hbm["fused"] = dict(achieved=fused_bytes * a.steps / (ms / 1e3) / 1e9 / world, peak=hbm, unit="GB/s",
                                 frac=fused_bytes * a.steps / (ms / 1e3) / 1e9 / world / hbm,
                                 note="the timed step (fused mode) against the fully-fused lower bound of SURVEY 8(d): "
                                      "6,736,896 B/img (image in, codes + metadata out, image out)")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(WORKLOAD, global_batch=B * world, per_gpu_batch=B, parallelism=f"image-sharded x{world}",
                            dct_impl=a.dct_impl, mode=mode,
                            l2="inputs (805 MB/GPU at B=256) larger than L2, no flush needed",
                            patchnorm_fit_ms=fit_ms),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=hx.numel() * 4 * world,
                         d2h_bytes_per_step=(h_rec.numel() * 4 + h_codes.numel() * 8) * world,
                         ms_per_step=ms_e2e / e2e_steps),
                gpu_launches=launches, clocks=clocks, roofline=roofline, pipeline_hbm=pipeline_hbm,
                staged=dict(value=staged_value, unit=UNIT, ms_per_step=ms_staged / staged_steps,
                            note="same job through the drop-in modules one by one (roundtrip_staged)"),
                stages_ms=stages)

    if rank == 0 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_sample or 160 * cores          # ~10 s of CPU work on all cores
        ips, dt, procs = cpu_reference_throughput(n, S, cores)
        line["cpu_baseline"] = dict(value=ips, unit=UNIT, cores=procs, kind="port",
                                    sample=f"{n} images of {S}x{S} over {procs} worker processes, {dt:.1f} s "
                                           "(oracle/dcta_oracle.py run_pipeline)")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def stage_times(torch, D, pipe, x, dev, reps=3):
    """Median device time (ms) of each stage of one step, CUDA events on the launching stream."""
    fe, pn, lfq = pipe.extractor, pipe.norm, pipe.quantizer
    out = {}

    def t(name, fn):
        res = None
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[name] = statistics.median(ts)
        return res

    from dct_autoencoder_b200 import _lib
    U = D.util
    H, W = x.shape[-2:]
    p = fe.patch_size
    _, _, th, tw = fe._geometry(H, W)
    KH, KW = th * p, tw * p
    fold = fe.dct_impl == "tc" and U.fold_ok(H, W, KH, KW)
    tc = fe.dct_impl in ("tc", "tc_plain") and not fold
    if fold:
        hi, lo, dc = t("rgb_to_ipt_fold", lambda: U.rgb_to_ipt_fold(x))
        tiles = t("dct_fwd", lambda: U.dct2_fwd_fold(hi, lo, dc, KH, KW, tile_p=p, channels=3))
        if (pipe.fusable() and lfq.num_codebooks == p and lfq.codebook_dim == p
                and _lib.load().dcta_fold_codes_supported(H, W, KH, KW, p)):
            # what the fused step runs: pass 1 + pass 2 straight to code words (no token grid)
            t("dct_fwd_codes", lambda: U.dct2_fwd_fold_codes(hi, lo, dc, KH, KW, p, 3, pn))
        del hi, lo
    elif tc:
        hi, lo, dc = t("rgb_to_ipt_split", lambda: U.rgb_to_ipt_split(x))
        tiles = t("dct_fwd", lambda: U.dct2_fwd_tc(hi, lo, dc, KH, KW, tile_p=p, channels=3))
        del hi, lo
    else:
        ipt = t("rgb_to_ipt", lambda: U.rgb_to_ipt(x))
        tiles = t("dct_fwd", lambda: U.dct2_truncated(ipt, KH, KW, tile_p=p, channels=3))
        del ipt
    t("score_sort", lambda: fe._sorted_order(tiles))
    del tiles
    batch = t("encode_total(process_batch)", lambda: fe.process_batch(x))
    normed = t("patchnorm_fwd", lambda: pn(batch))
    q = t("lfq", lambda: lfq(normed, mask=~batch.key_pad_mask))[0]
    b2 = batch.shallow_copy()
    b2.patches = q
    inv = t("patchnorm_inv", lambda: pn.inverse_norm(b2))
    b2.patches = inv
    n = len(b2.patch_sizes)
    slot_map, _ = t("slot_map", lambda: fe._slot_map(b2, th, tw))
    st = _lib.stream_ptr(dev)
    ipt2 = None
    if fold:
        y_hi = torch.empty((2, 2, n * 3, KH // 2, U._round8(KW // 2)), dtype=torch.float16, device=dev)
        y_lo = torch.empty_like(y_hi)
        dc2 = torch.empty(n * 3, dtype=torch.float32, device=dev)
        t("unpatchify", lambda: _lib.call("dcta_unpatchify_fold", _lib.ptr(inv), _lib.ptr(slot_map), None, n, 3, th, tw,
                                          p, KH, KW, H, W, _lib.ptr(y_hi), _lib.ptr(y_lo), _lib.ptr(dc2), st))
        zq = t("dct_inv", lambda: U.dct2_inv_fold(y_hi, y_lo, KH, KW, H, W))
        t("unfold_ipt_to_rgb", lambda: U.unfold_ipt_to_rgb(zq, dc2, H, W))
        del zq
    elif tc:
        y_hi = torch.empty((n, 3, KH, U._round8(KW)), dtype=torch.float16, device=dev)
        y_lo = torch.empty_like(y_hi)
        dc2 = torch.empty(n * 3, dtype=torch.float32, device=dev)
        t("unpatchify", lambda: _lib.call("dcta_unpatchify_split", _lib.ptr(inv), _lib.ptr(slot_map), None, n, 3, th, tw,
                                          p, KH, KW, U._round8(KW), H, W, _lib.ptr(y_hi), _lib.ptr(y_lo), _lib.ptr(dc2), st))
        ipt2 = t("dct_inv", lambda: U.dct2_inv_tc(y_hi, y_lo, dc2, KW, H, W))
    else:
        planes = torch.empty((n, 3, KH, KW), dtype=torch.float32, device=dev)
        t("unpatchify", lambda: _lib.call("dcta_unpatchify", _lib.ptr(inv), _lib.ptr(slot_map), None, n, 3, th, tw, p,
                                          KH, KW, _lib.ptr(planes), st))
        ipt2 = t("dct_inv", lambda: U.idct2_truncated(planes, H, W))
    if ipt2 is not None:
        t("ipt_to_rgb", lambda: U.ipt_to_rgb(ipt2))
    if pipe.fusable():
        del ipt2, inv, q, normed
        fb, fcodes = t("fused:encode_codes(total)", lambda: pipe.encode_codes(x))
        t("fused:decode_codes(total)", lambda: pipe.decode_codes(fb, fcodes))
    return out


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))
