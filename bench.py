#!/usr/bin/env python
"""Benchmark of the encode+decode transform path (BASELINE.json metric: images/s at 512^2, patch 14).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
                    [--config 2|3a|3b|4|5] [--global-batch G]

`--config` selects the BASELINE.json configuration (default 2 = the headline metric; the driver runs that one):
  2s       config 2 with the SECONDARY quantiser of SURVEY 8(d): LFQ(dim=196, codebook_size=8192, num_codebooks=16), i.e.
           the conf/patch14-l.json quantiser with its 196 -> 208 -> 196 projections (weights manual_seed(2))
  3a / 3b  synthetic 1024^2, batch 128, max_seq_len 1024: pure top-k cap / beta = 0.004 variable k + row packing
  4        VectorQuantize nearest code, codebook 8192 x 256, 512 x 3072 tokens (metric: tokens/s)
  5        4096 images of 512^2 in total, sharded G / N per GPU (strong scaling), one PatchNorm statistic-fit step
           with its all-reduce, then the config-2 pipeline

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
    ap.add_argument("--config", default="2", choices=["2", "2s", "3a", "3b", "4", "5"])
    ap.add_argument("--global-batch", type=int, default=4096, help="config 5: images in total over all GPUs")
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
class Ctx:
    """One rank: device, process group, timing helpers."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import dct_autoencoder_b200 as D
        from dct_autoencoder_b200 import _lib
        self.torch, self.dist, self.D, self.lib = torch, dist, D, _lib
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            # create the NCCL communicator NOW (first collective = hundreds of ms), outside everything timed
            t = torch.zeros(1, device=self.dev)
            dist.all_reduce(t)
            torch.cuda.synchronize()
        _lib.load()
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            self.peaks = {}
        self.hbm = float(self.peaks.get("hbm_gbs", 6650.0))
        self.peak_tf = float(self.peaks.get("bf16_tflops_sustained", 1400.0))
        self.peak_src = "measured (MEASURED_PEAKS.json)" if self.peaks else "fallback (B200_PROFILING.md)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """ms for `steps` calls: barrier + synchronize on both sides, CUDA events, MAX over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def finish(self, line):
        if self.rank == 0:
            print(json.dumps(line), flush=True)
        if self.world > 1:
            self.dist.destroy_process_group()
        return 0


def traffic_from_profiles(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed summary of the
    `ncu --set full` capture of this workload (profiles/traffic.json, written by profiles/summarize.py traffic)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(key)
        return (float(e["bytes_per_launch"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def make_pipe(ctx, max_seq_len=3072, beta=0.0, dct_impl="tc"):
    D = ctx.D
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, beta, 32, 32, max_seq_len, dct_impl=dct_impl)
    pn = D.PatchNorm(32, 32, 14, 3).to(ctx.dev)
    lfq = D.LFQ(codebook_size=2 ** 14, num_codebooks=14).to(ctx.dev).eval()
    return D.TransformPipeline(fe, pn, lfq)


def fit_norm_timed(ctx, pipe, fit_x, steps=10):
    """The PatchNorm statistic fit (main.py:115-149: `steps` update steps, default 10 in main.py:297), with its two
    all-reduces per step when world > 1.  Returns a dict: first (cold) step, warm step time (CUDA events, max over
    ranks), the two all-reduces alone, and whether every rank ended with identical tables."""
    torch, dist = ctx.torch, ctx.dist
    t0 = time.perf_counter()
    pipe.fit_norm(fit_x)
    torch.cuda.synchronize()
    first_ms = 1e3 * (time.perf_counter() - t0)
    ms = ctx.timed(lambda: pipe.fit_norm(fit_x), steps, 1)
    out = dict(fit_first_step_ms=first_ms, fit_step_ms=ms / steps, fit_steps_timed=steps,
               fit_images_per_rank=int(fit_x.shape[0]))
    if ctx.world > 1:
        n_pos, z = 3 * 32 * 32, 196
        packed = torch.zeros(n_pos + n_pos * z, device=ctx.dev)
        abs_dev = torch.zeros(n_pos * z, device=ctx.dev)

        def two_allreduces():
            dist.all_reduce(packed)
            dist.all_reduce(abs_dev)
        out["allreduce_pair_ms"] = ctx.timed(two_allreduces, 20, 3) / 20
        out["allreduce_bytes"] = int(packed.numel() * 4 + abs_dev.numel() * 4)
        same = True
        for t in (pipe.norm.n.data, pipe.norm.median.data, pipe.norm.b.data):
            hi, lo = t.clone(), t.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            same = same and bool(torch.equal(hi, lo))
        out["stats_identical_across_ranks"] = same
    return out


def link_bandwidth(ctx, nbytes=256 << 20, reps=4):
    """Pinned-memory copy rates of this box (GB/s): host->device alone, device->host alone, and per direction with
    both running -- the denominator of `e2e` (boxes of the pool differ by 3x here, the kernels do not)."""
    torch = ctx.torch
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(ctx.dev), torch.cuda.Stream(ctx.dev)

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def rate(fns):
        for f in fns:
            f()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        return nbytes * reps / (time.perf_counter() - t0) / 1e9

    out = dict(h2d_GBps=rate([h2d]), d2h_GBps=rate([d2h]), both_GBps_per_direction=rate([h2d, d2h]))
    return {k: round(v, 2) for k, v in out.items()}


def e2e_host(ctx, pipe, x, steps, chunk):
    """End to end through the public host API (TransformPipeline.roundtrip_host): pinned host buffers in, pinned host
    buffers out, copies inside the timed region.  Returns (fp32-contract dict, compact-mode dict)."""
    torch, D = ctx.torch, ctx.D
    B = x.shape[0]
    old_affinity = os.sched_getaffinity(0)
    numa_bound = D.util.bind_to_gpu_numa(ctx.dev)      # pinned buffers live on the allocating thread's NUMA node
    link = link_bandwidth(ctx)
    hx = torch.empty(tuple(x.shape), dtype=torch.float32).pin_memory()
    hx.copy_(x)
    h_rec = torch.empty(tuple(x.shape), dtype=torch.float32).pin_memory()
    s, c = pipe.extractor.max_seq_len, pipe.quantizer.num_codebooks
    h_codes = torch.empty((B, s, c), dtype=torch.int64).pin_memory()
    marks = []

    def host_step():
        pipe.roundtrip_host(hx, h_rec, h_codes, chunk=chunk)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(ev)
    ms = ctx.timed(host_step, steps, 3)
    torch.cuda.synchronize()
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - steps - 1, len(marks) - 1)]
    fp32 = dict(value=ctx.world * B * steps / (ms / 1e3), unit=UNIT, h2d_bytes_per_step=hx.numel() * 4 * ctx.world,
                d2h_bytes_per_step=(h_rec.numel() * 4 + h_codes.numel() * 8) * ctx.world, ms_per_step=ms / steps,
                ms_per_step_min=min(per_step), ms_per_step_max=max(per_step),
                contract="the reference's types: fp32 images in, fp32 images + int64 codes out")
    # the floor the link of THIS box sets for those bytes (both directions busy at once), and how close the step is
    floor_ms = max(hx.numel() * 4, h_rec.numel() * 4 + h_codes.numel() * 8) / (link["both_GBps_per_direction"] * 1e6)
    fp32.update(link=link, link_floor_ms=floor_ms, frac_of_link_floor=floor_ms / (ms / steps))
    # compact: the same computation with 8-bit pixels on both sides of the link and the codes as wire records
    hx8 = (hx * 255).round().to(torch.uint8).pin_memory()
    h_rec8 = torch.empty(tuple(x.shape), dtype=torch.uint8).pin_memory()
    rec_bytes = D.dct_patches.wire_record_bytes(c, pipe.quantizer.codebook_dim)
    h_wire = torch.empty((B, s, rec_bytes), dtype=torch.uint8).pin_memory()
    h_counts = torch.empty(B, dtype=torch.int32).pin_memory()
    ms8 = ctx.timed(lambda: pipe.roundtrip_host(hx8, h_rec8, h_wire, chunk=chunk, compact=True, out_counts=h_counts),
                    steps, 2)
    compact = dict(value=ctx.world * B * steps / (ms8 / 1e3), unit=UNIT, h2d_bytes_per_step=hx8.numel() * ctx.world,
                   d2h_bytes_per_step=(h_rec8.numel() + h_wire.numel() + h_counts.numel() * 4) * ctx.world,
                   ms_per_step=ms8 / steps,
                   contract="uint8 pixels in (read as x/255, exact), uint8 pixels out (save_image quantisation), codes as "
                            "27-byte wire records (u16 c|h|w + 14x14 code bits): same codes, 4x fewer PCIe bytes")
    if numa_bound:
        os.sched_setaffinity(0, old_affinity)
    return fp32, compact


def run_ours(a):
    ctx = Ctx()
    if a.config in ("3a", "3b"):
        return run_config3(ctx, a)
    if a.config == "2s":
        return run_config2s(ctx, a)
    if a.config == "4":
        return run_config4(ctx, a)
    if a.config == "5":
        return run_config5(ctx, a)
    torch, D, _lib = ctx.torch, ctx.D, ctx.lib
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B, S = a.batch, a.size
    pipe = make_pipe(ctx, dct_impl=a.dct_impl)
    fe, pn = pipe.extractor, pipe.norm

    # PatchNorm statistics fitted on a different batch (all-reduced over the ranks when world > 1)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    fit_x = torch.rand(min(B, 64), 3, S, S, device=dev, generator=g)
    fit = fit_norm_timed(ctx, pipe, fit_x)
    del fit_x

    g.manual_seed(rank)
    x = torch.rand(B, 3, S, S, device=dev, generator=g)        # 805 MB at B=256: larger than the 126 MB L2

    # ---- device-resident throughput (value)
    def step():
        return pipe.roundtrip(x)

    sampler = ClockSampler(ctx.local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    skip = sampler.mark() if sampler else 0
    l0 = _lib.launch_count
    ms = ctx.timed(step, a.steps, 0)
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

    # ---- the same step replayed from a CUDA graph (one driver call per step, no Python between the launches)
    graphed = None
    try:
        gr = pipe.graphed(x)
        ms_g = ctx.timed(lambda: gr(), a.steps, 2)
        graphed = dict(value=world * B * a.steps / (ms_g / 1e3), unit=UNIT, ms_per_step=ms_g / a.steps,
                       launches_per_replay=gr.launches,
                       note="TransformPipeline.graphed(x): the same launches captured once in a CUDA graph")
        del gr
    except Exception as e:          # capture is an optimisation of the host side, never a requirement
        graphed = dict(error=str(e)[:200])

    # ---- the same job through the drop-in modules one by one (every intermediate materialised)
    staged_steps = max(2, min(a.steps, 5))
    ms_staged = ctx.timed(lambda: pipe.roundtrip_staged(x), staged_steps, 1)
    staged_value = world * B * staged_steps / (ms_staged / 1e3)

    # ---- end to end through the public API with HOST buffers (pinned in, pinned out)
    e2e_steps = max(2, min(a.steps, 8))
    e2e, e2e_compact = e2e_host(ctx, pipe, x, e2e_steps, a.chunk)

    # ---- per-stage device times (CUDA events around each public call) for the roofline object
    stages = stage_times(torch, D, pipe, x, dev, reps=3)
    K = min(S // 14, 32) * 14
    flops_fwd = 3 * 2 * S * K * (S + K) * B          # C*2*H*K*(W+K) per image
    dct_ms = stages["dct_fwd"]
    achieved = flops_fwd / (dct_ms / 1e3) / 1e12
    peak_tf, peak_src, hbm = ctx.peak_tf, ctx.peak_src, ctx.hbm
    launch_times = step_launch_times(ctx, step, reps=max(3, min(a.steps, 10)))
    names = [n for n, _ in launch_times]
    if (fe.dct_impl == "tc" and D.util.fold_ok(S, S, K, K) and names.count("fold_gemm") == 3 and names.count("fold_codes") == 1
            and pipe.fusable()):
        # The DCT GEMM launches of the TIMED step, each timed live by the library's own CUDA events on the launching
        # stream (dcta_profile_begin / _end): forward pass 1, forward pass 2 -> code words (fold_codes_kernel),
        # inverse pass 1 with the operand generated from code bits, inverse pass 2.
        # Algorithmic bytes per launch: fp16 hi/lo operands in, next operand / code words / fp32 out.
        Kq, Sq = K // 2, S // 2
        xq = 3 * 4 * Sq * Sq * 4            # image quadrants, hi + lo fp16            (= 3*S*S*4)
        pt = 3 * 2 * K * Sq * 4             # forward intermediate P^T [a][plane][kw][h/2], hi + lo
        qt = 3 * 4 * Sq * Kq * 4            # inverse intermediate Q^T, hi + lo
        zq = 3 * 4 * Sq * Sq * 4            # fp32 quadrant transforms
        n_tok = 3 * (K // 14) * (K // 14)
        bv = 3 * 4 * ((Kq + 31) // 32) * ((Kq + 31) // 32) * 256      # sign/valid words read by inverse pass 1
        gi = [i for i, n in enumerate(names) if n == "fold_gemm"]
        ci = names.index("fold_codes")
        # executed tensor flops per image (every product is three fp16 MMAs; fold_codes runs 256 basis rows of which Kq are valid)
        f_fwd1 = 2 * Sq * Sq * Kq * 12 * 3
        f_codes = 2 * 256 * ((Kq + 255) // 256) * K * Sq * 6 * 3
        f_inv1 = 2 * Kq * (((Kq + 31) // 32) * 32) * Sq * 12 * 3
        f_inv2 = 2 * Sq * Kq * Sq * 12 * 3
        fam = [("fold_gemm_kernel<0> forward pass 1", launch_times[gi[0]][1], xq + pt, f_fwd1),
               ("fold_codes_kernel forward pass 2 -> LFQ code words", launch_times[ci][1], pt + n_tok * (14 * 4 + 4), f_codes),
               ("fold_gemm_kernel<0, GEN> inverse pass 1, operand generated from code bits", launch_times[gi[1]][1], bv + qt, f_inv1),
               ("fold_gemm_kernel<1> inverse pass 2", launch_times[gi[2]][1], qt + zq, f_inv2)]
        gemm_ms = sum(f[1] for f in fam)
        alg_total = sum(f[2] for f in fam) * B
        achieved_gbs = alg_total / (gemm_ms / 1e3) / 1e9
        tens = flops_fwd / ((fam[0][1] + fam[1][1]) / 1e3) / 1e12
        traffic, traffic_src = (traffic_from_profiles("dct_gemm_launches_b256_512") if (B == 256 and S == 512) else (None, None))
        roofline = dict(bound="hbm", kernel="fold_gemm_kernel x3 + fold_codes_kernel (the 4 DCT GEMM launches of the timed step; "
                        "cta_group::2 tcgen05 fp16x3 split precision, folded basis resident in shared memory)",
                        achieved=achieved_gbs, peak=hbm, unit="GB/s", frac=achieved_gbs / hbm,
                        traffic=traffic, traffic_source=traffic_src,
                        launches_per_step=4, avg_launch_ms=gemm_ms / 4,
                        algorithmic_bytes_per_launch=alg_total / 4,
                        share_of_step=gemm_ms / sum(t for _, t in launch_times),
                        launches=[dict(kernel=k, ms=t, algorithmic_bytes=ab * B, GBps=ab * B / (t / 1e3) / 1e9,
                                       frac=ab * B / (t / 1e3) / 1e9 / hbm, executed_tflops=fl * B / (t / 1e3) / 1e12,
                                       tensor_frac_executed=fl * B / (t / 1e3) / 1e12 / peak_tf) for k, t, ab, fl in fam],
                        launches_note="frac: algorithmic bytes against the measured HBM peak; tensor_frac_executed: the MMAs the "
                                      "kernel issues (three fp16 products per multiply, padding rows included) against the "
                                      "measured dense bf16 peak -- the launches are bound by both at once",
                        peak_source=peak_src, timing="CUDA events recorded by libdcta after each launch of the timed step "
                                                     "(dcta_profile_begin/_end), mean of several steps",
                        tensor=dict(achieved=tens, peak=peak_tf, unit="TFLOP/s", frac=tens / peak_tf,
                                    note="forward passes: SURVEY 8(d) flops 3*2*H*K*(W+K) per image (plain basis GEMMs); the "
                                         "folded kernel executes 1/2 of them, each as 3 tensor MMAs, so frac <= 2/3"),
                        note="achieved = algorithmic operand + output bytes of the 4 launches / their device time "
                             "(the kernels' OWN hi/lo operands, not SURVEY 8d's fused bound: that is pipeline_hbm.fused); "
                             "inverse pass 1 no longer reads coefficient planes, so it is bound by the tensor pipe and its "
                             "epilogue, not by HBM")
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
                            patchnorm_fit=fit),
                e2e=e2e, e2e_compact=e2e_compact,
                gpu_launches=launches, clocks=clocks, roofline=roofline, pipeline_hbm=pipeline_hbm,
                graphed=graphed,
                staged=dict(value=staged_value, unit=UNIT, ms_per_step=ms_staged / staged_steps,
                            note="same job through the drop-in modules one by one (roundtrip_staged)"),
                stages_ms=stages, launch_times_ms=[[n, round(t, 5)] for n, t in launch_times],
                launch_times_note="per-launch device times of the same step with every launch on ONE stream (differences of "
                                  "consecutive CUDA events): in the timed step sort_tokens and pack_codes_grid run on a side "
                                  "stream beside the decode (TransformPipeline.overlap_pack), so ms_per_step is below their sum")
    if world > 1:
        line["e2e"]["note"] = ("all ranks share the host's memory and PCIe root complex: the host-to-host rate saturates there, "
                               "not on a collective (see e2e_compact for the same job with 4x fewer link bytes)")

    if rank == 0 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_sample or 160 * cores          # ~10 s of CPU work on all cores
        ips, dt, procs = cpu_reference_throughput(n, S, cores)
        line["cpu_baseline"] = dict(value=ips, unit=UNIT, cores=procs, kind="port",
                                    sample=f"{n} images of {S}x{S} over {procs} worker processes, {dt:.1f} s "
                                           "(oracle/dcta_oracle.py run_pipeline)")
    return ctx.finish(line)


def run_config2s(ctx, a):
    """Config 2 with the secondary quantiser: LFQ(dim=196, codebook_size=8192, num_codebooks=16) -- the
    conf/patch14-l.json quantiser: project_in 196 -> 208, 16 codebooks x 13 bits, project_out 208 -> 196 (lfq.py:60-62,
    164, 212).  The projections run on the split-precision tcgen05 GEMM (linear.py); every stage is a libdcta kernel."""
    torch, D = ctx.torch, ctx.D
    dev = ctx.dev
    B, S = a.batch, 512
    fe = D.DCTAutoencoderFeatureExtractor(3, 14, 0.0, 32, 32, 3072, dct_impl=a.dct_impl)
    pn = D.PatchNorm(32, 32, 14, 3).to(dev)
    torch.manual_seed(2)
    lfq = D.LFQ(dim=196, codebook_size=8192, num_codebooks=16).to(dev).eval()
    pipe = D.TransformPipeline(fe, pn, lfq)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + ctx.rank)
    fit = fit_norm_timed(ctx, pipe, torch.rand(min(B, 64), 3, S, S, device=dev, generator=g))
    g.manual_seed(ctx.rank)
    x = torch.rand(B, 3, S, S, device=dev, generator=g)
    l0 = ctx.lib.launch_count
    ms = ctx.timed(lambda: pipe.roundtrip(x), a.steps, a.warmup)
    launches = ctx.lib.launch_count - l0
    times = step_launch_times(ctx, lambda: pipe.roundtrip(x), reps=3)
    value = ctx.world * B * a.steps / (ms / 1e3)
    staged_bytes = (26004480 + 2 * 3072 * 208 * 4 * 2) * B          # + project_in / project_out activations written and read
    line = dict(metric="encode+decode images/sec at 512^2 patch14, LFQ 16 x 13 bit with 196->208->196 projections (config 2, secondary quantiser)",
                value=value, unit=UNIT, n_gpus=ctx.world, steps=a.steps, warmup=a.warmup, ms_per_step=ms / a.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload="config2 secondary: synthetic 512x512 RGB fp32, patch 14, max_patch 32x32, beta=0, max_seq_len 3072, "
                                     "PatchNorm frozen, LFQ(dim=196, codebook_size=8192, num_codebooks=16) with projections, decode to RGB",
                            image_size=S, patch_size=14, global_batch=B * ctx.world, per_gpu_batch=B, fusable=pipe.fusable(),
                            l2="inputs (805 MB/GPU at B=256) larger than L2", patchnorm_fit=fit),
                gpu_launches=launches,
                pipeline_hbm=dict(bound="hbm", achieved=staged_bytes * a.steps / (ms / 1e3) / 1e9, peak=ctx.hbm, unit="GB/s",
                                  note="staged-API algorithmic bytes (SURVEY 8d) + the projection activations"),
                launch_times_ms=[[n, round(t, 5)] for n, t in times])
    line["pipeline_hbm"]["frac"] = line["pipeline_hbm"]["achieved"] / ctx.hbm
    return ctx.finish(line)


def run_config3(ctx, a):
    """BASELINE config 3: synthetic 1024^2, batch 128, max_seq_len 1024 (73 x 73 tiles -> 3072 in-bounds candidates).
    3a: beta = 0, pure top-k cap, one image per row.  3b: random.seed(42), beta = 0.004: variable k drawn on the host in
    the reference's RNG order, several images per row, padding, multi-image decode."""
    import random
    torch, D = ctx.torch, ctx.D
    B, S = (a.batch if a.batch != 256 else 128), 1024
    beta = 0.004 if a.config == "3b" else 0.0
    pipe = make_pipe(ctx, max_seq_len=1024, beta=beta, dct_impl=a.dct_impl)
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(1000 + ctx.rank)
    random.seed(7)
    fit = fit_norm_timed(ctx, pipe, torch.rand(16, 3, S, S, device=ctx.dev, generator=g), steps=3)
    g.manual_seed(ctx.rank)
    x = torch.rand(B, 3, S, S, device=ctx.dev, generator=g)       # 1.6 GB: larger than L2
    random.seed(42)
    ks = [pipe.extractor._choose_k(3072) for _ in range(B)]        # the reference's draw order, one per image
    batch, codes = pipe.encode_codes(x, ks)
    rows, tokens = int(codes.shape[0]), int((~batch.key_pad_mask).sum())
    l0 = ctx.lib.launch_count
    ms = ctx.timed(lambda: pipe.roundtrip(x, ks), a.steps, a.warmup)
    launches = ctx.lib.launch_count - l0
    value = ctx.world * B * a.steps / (ms / 1e3)
    line = dict(metric=f"encode+decode images/sec at 1024^2 patch14 (config {a.config})", value=value, unit=UNIT,
                n_gpus=ctx.world, steps=a.steps, warmup=a.warmup, ms_per_step=ms / a.steps, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"config{a.config}: synthetic 1024x1024 RGB fp32, patch 14, max_patch 32x32, beta={beta}, "
                                     "max_seq_len 1024, PatchNorm frozen, LFQ 14x14bit, decode to RGB",
                            image_size=S, patch_size=14, global_batch=B * ctx.world, per_gpu_batch=B,
                            rows=rows, tokens=tokens, tokens_per_image_min=min(ks), tokens_per_image_max=max(ks),
                            fusable=pipe.fusable(), l2="inputs (1.6 GB/GPU) larger than L2", patchnorm_fit=fit),
                gpu_launches=launches,
                pipeline_hbm=dict(bound="hbm", achieved=(2 * 3 * S * S * 4) * B * a.steps / (ms / 1e3) / 1e9, peak=ctx.hbm,
                                  unit="GB/s", note="image in + image out only (the fully-fused bound of this config)"))
    line["pipeline_hbm"]["frac"] = line["pipeline_hbm"]["achieved"] / ctx.hbm
    line["launch_times_ms"] = [[n, round(t, 5)] for n, t in step_launch_times(ctx, lambda: pipe.roundtrip(x, ks), reps=3)]
    return ctx.finish(line)


def run_config4(ctx, a):
    """BASELINE config 4: VectorQuantize(dim=256, codebook_size=8192) eval forward on randn(512, 3072, 256), codebook
    normal_() seed 3, mask = ones.  The token batch is processed in slices of 64 x 3072 tokens."""
    torch, D = ctx.torch, ctx.D
    dev = ctx.dev
    vq = D.VectorQuantize(dim=256, codebook_size=8192).to(dev).eval()
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    vq._codebook.embed.copy_(torch.randn(1, 8192, 256, device=dev, generator=g))
    g.manual_seed(ctx.rank)
    n_slices, sl = 8, 64
    x = torch.randn(sl, 3072, 256, device=dev, generator=g)       # one slice (201 MB > L2), reused for the 8 slices
    mask = torch.ones(sl, 3072, dtype=torch.bool, device=dev)
    T = n_slices * sl * 3072

    def step():
        for _ in range(n_slices):
            q, ind, _ = vq(x, mask=mask)
        return ind
    l0 = ctx.lib.launch_count
    ms = ctx.timed(step, a.steps, a.warmup)
    launches = ctx.lib.launch_count - l0
    flops = 2.0 * T * 8192 * 256
    tf = flops * a.steps / (ms / 1e3) / 1e12
    line = dict(metric="VectorQuantize nearest-code tokens/sec, codebook 8192 x 256 (config 4)",
                value=ctx.world * T * a.steps / (ms / 1e3), unit="tokens/s", n_gpus=ctx.world, steps=a.steps,
                warmup=a.warmup, ms_per_step=ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload="config4: VectorQuantize eval forward, randn(512, 3072, 256) x codebook 8192 x 256, "
                                     "nearest code + gather, indices int64", tokens=T, impl=vq.vq_impl),
                gpu_launches=launches,
                roofline=dict(bound="tensor", kernel="vq nearest-code distance GEMM (tcgen05) + argmin epilogue",
                              achieved=tf, peak=ctx.peak_tf, unit="TFLOP/s", frac=tf / ctx.peak_tf, traffic=None,
                              peak_source=ctx.peak_src,
                              note="algorithmic flops 2*T*8192*256 (SURVEY 8d); the whole forward incl. operand "
                                   "preparation, |e|^2 and the gather"))
    line["launch_times_ms"] = [[n, round(t, 5)] for n, t in step_launch_times(ctx, step, reps=2)]
    return ctx.finish(line)


def run_config5(ctx, a):
    """BASELINE config 5: G = 4096 images of 512^2 in total, sharded G / N per GPU (strong scaling); one PatchNorm
    statistic-fit step with its all-reduce on the first 64 images of each shard, then the config-2 pipeline over the
    whole shard in batches of 256."""
    torch, D = ctx.torch, ctx.D
    G, S, Bc = a.global_batch, a.size, 256
    assert G % ctx.world == 0
    per = G // ctx.world
    pipe = make_pipe(ctx, dct_impl=a.dct_impl)
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(ctx.rank)                                        # seed = rank
    n_buf = min(per, 1024)                                         # distinct images resident per GPU (3.2 GB per 256)
    x = torch.rand(n_buf, 3, S, S, device=ctx.dev, generator=g)
    fit = fit_norm_timed(ctx, pipe, x[:64], steps=10)
    chunks = [x[(i % n_buf):(i % n_buf) + min(Bc, per - i)] for i in range(0, per, Bc)]

    def step():
        for c in chunks:
            pipe.roundtrip(c)
    l0 = ctx.lib.launch_count
    ms = ctx.timed(step, a.steps, a.warmup)
    launches = ctx.lib.launch_count - l0
    line = dict(metric=METRIC, value=G * a.steps / (ms / 1e3), unit=UNIT, n_gpus=ctx.world, steps=a.steps,
                warmup=a.warmup, ms_per_step=ms / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(WORKLOAD, workload="config5: " + WORKLOAD["workload"][len("config2: "):] +
                            f"; {G} images in total, {per} per GPU in batches of {Bc}, statistic fit all-reduced over the ranks",
                            global_batch=G, per_gpu_batch=per, parallelism=f"image-sharded x{ctx.world}",
                            patchnorm_fit=fit, l2="805 MB per 256-image batch, larger than L2"),
                gpu_launches=launches)
    return ctx.finish(line)


def step_launch_times(ctx, step, reps=5):
    """[(launch group, mean device ms)] of one step, from the library's own per-launch CUDA events
    (dcta_profile_begin / dcta_profile_end: an event on the launching stream after every launch group)."""
    _lib = ctx.lib
    step()
    ctx.torch.cuda.synchronize()
    runs = []
    for _ in range(reps):
        with _lib.profile(ctx.dev) as prof:
            step()
        runs.append(prof.groups)
    n = max(set(len(r) for r in runs), key=[len(r) for r in runs].count)      # steps with the usual launch sequence
    runs = [r for r in runs if len(r) == n]
    return [(runs[0][i][0], sum(r[i][1] for r in runs) / len(runs)) for i in range(n)]


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
