#!/usr/bin/env python
"""Timing probe of the host-buffer wald_test call under the residency cache (tuning aid)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pydeseq2_b200.inference import B200Inference  # noqa: E402
from pydeseq2_b200.pipeline import median_of_ratios  # noqa: E402
from pydeseq2_b200.synth import make_counts  # noqa: E402

counts, X, _ = make_counts(200, 20000, "two_level", seed=0)
counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
sf = median_of_ratios(counts)[1]
inf = B200Inference(device=0)
ctx = inf._ops.ctx
G = counts.shape[1]
disp = np.full(G, 0.1)
ridge = np.diag(np.repeat(1e-6, 2))
con = np.array([0.0, 1.0])


def t(label, fn, n=5):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label:60s} " + " ".join(f"{x:7.3f}" for x in ts), flush=True)


cp = ctx.pinned_empty(counts.shape, np.int64)
cp[:] = counts
beta, mu, hat, conv = inf.irls(cp, sf, X, disp, 0.5, 1e-8)
t("irls (counts pinned, hit)", lambda: inf.irls(cp, sf, X, disp, 0.5, 1e-8))
t("wald: mu = pinned OUTPUT block of irls (GPU-written)", lambda: inf.wald_test(X, disp, beta, mu, ridge, con, 0.0, None))
mu_pg = np.array(mu)
t("wald: mu = pageable copy", lambda: inf.wald_test(X, disp, beta, mu_pg, ridge, con, 0.0, None))
mu_pc = ctx.pinned_empty(mu.shape, np.float64)
mu_pc[:] = mu
t("wald: mu = pinned block written by the CPU", lambda: inf.wald_test(X, disp, beta, mu_pc, ridge, con, 0.0, None))


def seq():
    b, m, h, c = inf.irls(cp, sf, X, disp, 0.5, 1e-8)
    t0 = time.perf_counter()
    inf.wald_test(X, disp, b, m, ridge, con, 0.0, None)
    return (time.perf_counter() - t0) * 1e3


print("wald right after irls (fresh output block each time):", [round(seq(), 3) for _ in range(5)])
t0 = time.perf_counter(); s = float(mu.sum()); print("numpy sum of the GPU-written pinned block ms", (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); s = float(mu_pg.sum()); print("numpy sum of the pageable copy ms", (time.perf_counter() - t0) * 1e3)
print(ctx.residency_stats())
