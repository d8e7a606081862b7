import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pydeseq2_b200.inference import B200Inference
from pydeseq2_b200.pipeline import median_of_ratios
from pydeseq2_b200.synth import make_counts
counts, X, _ = make_counts(200, 20000, "two_level", 0)
counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
sf = median_of_ratios(counts)[1]
N, G = counts.shape
rng = np.random.default_rng(0)
disp = np.exp(rng.normal(-1.5, 1.0, G))
inf = B200Inference(device=0)
b1 = inf.irls(counts, sf, X, disp, 0.5, 1e-8)
b1b = inf.irls(counts, sf, X, disp, 0.5, 1e-8)
print("repeat identical:", [bool(np.array_equal(a, b)) for a, b in zip(b1, b1b)])
perm = rng.permutation(G)
b2 = inf.irls(np.ascontiguousarray(counts[:, perm]), sf, X, disp[perm], 0.5, 1e-8)
bad = np.flatnonzero((b2[0] != b1[0][perm]).any(1))
print("mismatching genes", len(bad), "of", G)
if len(bad):
    g = bad[:10]
    rel = np.abs(b2[0][bad] - b1[0][perm][bad]) / np.abs(b1[0][perm][bad])
    print("max rel", rel.max(), "median", np.median(rel))
    print("positions (perm order)", g, "orig index", perm[g])
    print("tile of perm pos", g // 4, "tile of orig", perm[g] // 4)
    # iterations? compare counts small
    print("min count of bad genes", counts[:, perm[g]].min(0), "max", counts[:, perm[g]].max(0))
    print("disp", disp[perm[g]])
os.environ["PDQ_RESIDENCY"] = "0"
inf2 = B200Inference(device=0)
c1 = inf2.irls(counts, sf, X, disp, 0.5, 1e-8)
c2 = inf2.irls(np.ascontiguousarray(counts[:, perm]), sf, X, disp[perm], 0.5, 1e-8)
print("no residency: mismatching genes", int((c2[0] != c1[0][perm]).any(1).sum()), " vs resident run equal:", bool(np.array_equal(c1[0], b1[0])))
