#!/usr/bin/env python
"""Per-stage device timings of the resident pass for one workload (tuning aid; bench.py is the contract).
    PDQ_LIB=pydeseq2_b200/libvariant.so python scripts/stage_bench.py --genes 60000 --samples 500 --design factorial"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pydeseq2_b200 import _lib  # noqa: E402
from pydeseq2_b200.inference import B200Inference  # noqa: E402
from pydeseq2_b200.pipeline import ResidentFit, median_of_ratios  # noqa: E402
from pydeseq2_b200.synth import make_counts  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=20000)
ap.add_argument("--samples", type=int, default=200)
ap.add_argument("--design", default="two_level")
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
counts, X, _ = make_counts(a.samples, a.genes, a.design, seed=0)
counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
sf = median_of_ratios(counts)[1]
inf = B200Inference(device=0, lanes_per_gene=a.lanes)
ctx = inf._ops.ctx
rf = ResidentFit(ctx, X, sf)
rf.upload(counts)
flush = ctx.malloc(256 << 20)
for _ in range(3):
    rf.run()
ms = []
for _ in range(a.steps):
    ctx.check(ctx.lib.pdq_memset(ctx.h, _lib.c_dptr(flush), 1, 256 << 20))
    ctx.sync()
    rf.run(events=(0, 1))
    ms.append(ctx.elapsed_ms(0, 1))
rf.run(profile=True)
rf.run(profile=True)
print(json.dumps({"lib": os.environ.get("PDQ_LIB", "default"), "workload": f"{a.genes}x{a.samples} {a.design} lanes={a.lanes}",
                  "ms_per_step": round(float(np.mean(ms)), 4), "min": round(float(np.min(ms)), 4),
                  "stages_ms": {k: round(v, 4) for k, v in rf.stage_ms.items()}}))
