#!/usr/bin/env python
"""bench.py -- genes/s of the full deseq2() hot path (dispersion + IRLS + Wald) on B200, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic counts: method-of-moments start values,
initial mu (lin_reg_mu or IRLS), genewise dispersion MLE, dispersion trend + prior, MAP dispersions, the
log-fold-change IRLS and the Wald test -- the Inference calls `DeseqDataSet.deseq2()` +
`DeseqStats.run_wald_test()` make (reference dds.py:516-562, ds.py:303-360).  Size factors are computed once
outside the timed region for both arms (median of ratios is "next" scope).

Workload = BASELINE.json configs[1]: 20 000 genes x 200 samples, one 2-level factor, per GPU (weak scaling:
every rank owns a 20 000-gene shard, the trend/prior step all-gathers the per-gene vectors over NCCL).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (counts already in HBM), timed per step with
CUDA events on the library's stream, L2 flushed between steps, max over ranks.  `e2e` = the same metric through the
reference-facing plugin calls with HOST buffers (H2D/D2H inside the timed region).  `roofline` describes the
dominant kernel, `cpu_baseline` the oracle port (joblib over genes, like the reference) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genes/sec full deseq2() fit (disp+IRLS+Wald)"
UNIT = "genes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--samples", type=int, default=200)
    ap.add_argument("--design", default="two_level", choices=["two_level", "factorial", "continuous"])
    ap.add_argument("--cpu-sample-genes", type=int, default=20000,
                    help="genes of the workload the CPU baseline is timed on (20 000 x 200 is ~15 s on the GPU box)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per gene override (0 = auto)")
    return ap.parse_args()


def workload(args, rank):
    from pydeseq2_b200.pipeline import median_of_ratios
    from pydeseq2_b200.synth import make_counts

    counts, X, _ = make_counts(args.samples, args.genes, args.design, seed=rank)
    G_in = counts.shape[1]
    counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])  # dds.py:729-731 (throughput counts input G)
    # size factors are per sample and global over genes: every rank derives them from the rank-0 shard's generator
    ref_counts = counts if rank == 0 else make_counts(args.samples, args.genes, args.design, seed=0)[0]
    _, sf = median_of_ratios(ref_counts[:, ~(ref_counts == 0).all(0)])
    return counts, X, sf, G_in


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(kernel)
    return None


def cpu_fit(counts, X, sf, n_cpus):
    """The reference CPU path (oracle port: numpy/scipy per gene, joblib/loky over genes)."""
    from oracle import nbglm  # the one place bench.py may execute oracle/: the CPU baseline
    from pydeseq2_b200.pipeline import fit_host

    inf = nbglm.OracleInference(n_cpus=n_cpus)
    t0 = time.perf_counter()
    fit_host(counts, X, inf, size_factors=sf)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    counts, X, sf, _ = workload(args, 0)
    n = min(args.cpu_sample_genes, 8000, counts.shape[1])  # bounded: K + W passes must end within minutes
    sample = np.ascontiguousarray(counts[:, :n])
    cores = os.cpu_count() or 1
    cpu_fit(sample[:, :256], X, sf, cores)  # spawn the loky pool outside the timed region
    for _ in range(args.warmup):
        cpu_fit(sample, X, sf, cores)
    times = [cpu_fit(sample, X, sf, cores) for _ in range(args.steps)]
    dt = float(np.mean(times))
    val = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config(args, note=f"reference CPU path: oracle port of DefaultInference (numpy/scipy per gene, joblib/loky, "
                                        f"{cores} processes); each step = the first {n} genes of the workload"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first {n} of {args.genes} genes x {args.samples} samples, {args.steps} timed passes"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config(args, note=None):
    p = {"two_level": 2, "factorial": 3, "continuous": 3}[args.design]
    c = {"workload": f"{args.genes} genes x {args.samples} samples per GPU, {args.design} design (p={p}); BASELINE.json configs[1]"
                     if (args.genes, args.samples, args.design) == (20000, 200, "two_level")
                     else f"{args.genes} genes x {args.samples} samples per GPU, {args.design} design (p={p})",
         "genes_per_gpu": args.genes, "samples": args.samples, "p": p,
         "parallelism": f"gene shards x{args.gpus}, one NCCL all-gather of per-gene vectors before the trend fit",
         "size_factors": "median of ratios, precomputed outside the timed region (both arms)",
         "l2": "flushed between timed steps (256 MiB device memset)"}
    if note:
        c["note"] = note
    return c


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from pydeseq2_b200 import _lib
    from pydeseq2_b200.inference import B200Inference
    from pydeseq2_b200.pipeline import ResidentFit, fit_host
    from pydeseq2_b200.sharding import NcclComm

    counts, X, sf, G_in = workload(args, rank)
    N, G = counts.shape
    inf = B200Inference(device=local, lanes_per_gene=args.lanes)
    ctx = inf._ops.ctx
    comm = None
    if world > 1:
        import torch

        uid = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(NcclComm.make_unique_id(ctx)), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = NcclComm(ctx, _all_sizes(dist, G, world), rank, uid.cpu().numpy().tobytes())

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch

            dist.barrier()
            torch.cuda.synchronize()

    rf = ResidentFit(ctx, X, sf, comm=comm)
    rf.upload(counts)
    flush = ctx.malloc(256 << 20)

    def flush_l2():
        ctx.check(ctx.lib.pdq_memset(ctx.h, _lib.c_dptr(flush), 1, 256 << 20))
        ctx.sync()

    # ---------------------------------------------------------------- value: device-resident
    clk = ClockSampler(local).__enter__()  # sampled from the warm-up until the end of the e2e loop (steps last only ms)
    for _ in range(max(args.warmup, 3)):
        rf.run()
    barrier()
    launches0 = ctx.launches()
    step_ms = []
    if True:
        for _ in range(args.steps):
            flush_l2()
            barrier()
            ctx.record(0)
            rf.run()
            ctx.record(1)
            ctx.sync()
            step_ms.append(ctx.elapsed_ms(0, 1))
    barrier()
    launches = ctx.launches() - launches0
    total_ms = float(np.sum(step_ms))
    if dist is not None:
        import torch

        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = G_in * world / (ms_per_step * 1e-3)

    # ---------------------------------------------------------------- per-kernel timings -> roofline
    rf.run(profile=True)
    rf.run(profile=True)
    stages = dict(rf.stage_ms)
    rf.with_cooks = True  # Cook's distances: reported for information, NOT part of the timed step
    rf.run(profile=True)
    stages["cooks_untimed"] = rf.stage_ms.get("cooks")
    rf.with_cooks = False
    from pydeseq2_b200.pipeline import fit_shrink_prior_var

    res_sh = rf.run()
    k_sh = X.shape[1] - 1
    ps_sh = float(min(np.sqrt(fit_shrink_prior_var(np.asarray(res_sh["lfc"])[:, k_sh], np.asarray(res_sh["se"]))), 1.0))
    rf.lfc_shrink(res_sh, k_sh, prior_scale=ps_sh)  # first call allocates its buffers
    t0 = time.perf_counter()
    rf.lfc_shrink(res_sh, k_sh, prior_scale=ps_sh)  # apeGLM shrinkage: for information, NOT in the timed step
    stages["lfc_shrink_untimed"] = (time.perf_counter() - t0) * 1e3
    rf.device_size_factors()  # first call allocates its scratch
    ctx.sync()
    t0 = time.perf_counter()
    rf.device_size_factors()  # median of ratios on the device: reported for information, NOT part of the timed step
    stages["size_factors_dev_untimed"] = (time.perf_counter() - t0) * 1e3
    alg_bytes = {"mom_dispersions": 8, "lin_reg_mu": 16, "irls_init": 24, "alpha_mle_genewise": 16, "alpha_mle_map": 16,
                 "irls_lfc": 24, "wald_test": 8}  # bytes per (gene, sample): SURVEY.md §8(d)
    kern = {k: v for k, v in stages.items() if k in alg_bytes}
    top = max(kern, key=kern.get)
    peak, peak_src = peaks()
    achieved = alg_bytes[top] * N * G / (kern[top] * 1e-3) / 1e9
    kname = {"alpha_mle_genewise": "k_alpha_mle", "alpha_mle_map": "k_alpha_mle", "irls_lfc": "k_irls", "irls_init": "k_irls",
             "lin_reg_mu": "k_lin_reg_mu", "wald_test": "k_wald", "mom_dispersions": "k_mom_from_counts"}[top]
    roofline = {"bound": "hbm", "kernel": f"{kname} ({top})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic_from_profiles(kname), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[top] * N * G, "kernel_ms": kern[top],
                "note": "FP64-pipe bound (lgamma/digamma/exp/log per gene-sample-iteration), see DESIGN.md §5"}
    # second roofline, the one that binds: FP64 arithmetic.  Peak = DFMA throughput measured now on this device; achieved = the
    # kernel's FP64 flop count (2*DFMA + DMUL + DADD thread instructions, from the committed ncu capture of the same workload,
    # scaled to this run's gene-sample count) over its CUDA-event duration.
    prof = {}
    if os.path.exists(os.path.join(ROOT, "profiles", "traffic.json")):
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    flop_ref = prof.get("fp64_flop_per_launch", {}).get(kname)
    fp64_peak = ctx.fp64_peak_tflops()
    fp64 = {"peak_tflops": fp64_peak, "peak_source": "measured in this run (pdq_fp64_peak_tflops: DFMA chains on every SM)"}
    if flop_ref:
        flop = flop_ref * (N * G) / (19800 * 200)
        fp64.update({"achieved_tflops": flop / (kern[top] * 1e-3) / 1e12, "flop_per_launch": flop,
                     "frac": flop / (kern[top] * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None,
                     "pipe_active_pct_ncu": prof.get("fp64_pipe_active_pct", {}).get(kname)})
    roofline["fp64"] = fp64

    # ---------------------------------------------------------------- e2e: plugin calls with host buffers
    c_host = ctx.pinned_empty(counts.shape, np.int64)
    c_host[:] = counts
    n_host = ctx.pinned_empty(counts.shape, np.float64)  # layers["normed_counts"] of the orchestrator (dds.py:700-708)
    np.divide(counts, sf[:, None], out=n_host)
    n_means = n_host.mean(0)  # var["_normed_means"], also a product of fit_size_factors (dds.py:708)
    for _ in range(2):
        fit_host(c_host, X, inf, size_factors=sf, comm=comm, normed_counts=n_host, normed_means=n_means)
    barrier()
    ops = inf._ops
    h0, d0 = ops.h2d_bytes, ops.d2h_bytes
    e2e_t = []
    e2e_T = {}
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        fit_host(c_host, X, inf, size_factors=sf, comm=comm, timings=e2e_T, normed_counts=n_host, normed_means=n_means)
        ctx.sync()
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t))
    time.sleep(0.25)  # let nvidia-smi emit at least one more sample
    clk.__exit__()
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": G_in * world / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": (ops.h2d_bytes - h0) // args.steps, "d2h_bytes_per_step": (ops.d2h_bytes - d0) // args.steps}

    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = min(args.cpu_sample_genes, G)
        cores = os.cpu_count() or 1
        sample = np.ascontiguousarray(counts[:, :n])
        cpu_fit(sample[:, :256], X, sf, cores)  # pool start-up outside the timed region
        dt = cpu_fit(sample, X, sf, cores)
        cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n} of {G_in} genes x {N} samples, one pass after pool warm-up ({dt:.1f} s)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config(args), "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
                "stages_ms": {k: round(v, 4) for k, v in stages.items()},
                "e2e_calls_ms": {k: round(v * 1e3 / args.steps, 3) for k, v in e2e_T.items()}, "device": ctx.info()["name"]}
        print(json.dumps(line), flush=True)
    rf.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _all_sizes(dist, G, world):
    import torch

    t = torch.tensor([G], device="cuda", dtype=torch.int64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]


if __name__ == "__main__":
    main()
