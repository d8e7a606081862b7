"""GPU suite: the same edge cases / rarely taken branches on the CUDA path (C ABI)."""
import pytest

import edge_cases as ec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def inf():
    from pydeseq2_b200.inference import B200Inference

    return B200Inference(device=0)


def _force(inf, flag):
    def f(on):
        inf._ops.ctx.check(inf._ops.lib.pdq_set_debug_flags(inf._ops.ctx.h, flag if on else 0))

    return f


def test_zero_genes(inf):
    ec.check_zero_genes(inf)


def test_no_replicates(inf):
    ec.check_no_replicates_raises(inf)


def test_empty(inf):
    ec.check_empty_gene_set(inf)


def test_dtypes_layouts(inf):
    ec.check_input_dtypes_and_layouts(inf)


def test_wald_nan(inf):
    ec.check_nan_propagation_in_wald(inf)


def test_irls_optimizer_branch(inf):
    ec.check_irls_bounded_optimizer(inf, _force(inf, 1))


@pytest.mark.parametrize("name", ec.GRID_BETA)
def test_irls_grid_fallback(inf, name):
    ec.check_irls_grid_fallback(inf, _force(inf, 1 | 8), name)  # PDQ_DEBUG_FORCE_IRLS_OPTIMIZER | PDQ_DEBUG_FAIL_IRLS_OPTIMIZER


def test_alpha_grid_fallback(inf):
    ec.check_alpha_grid(inf, _force(inf, 2))


def test_design_too_large_is_reported(inf):
    import numpy as np

    from pydeseq2_b200._lib import B200Error

    X = np.ones((20, 17))
    with pytest.raises(B200Error, match="p <= 16"):
        inf.lin_reg_mu(np.ones((20, 3), dtype=np.int64), np.ones(20), X, 0.5)


def test_size_factors(inf):
    ec.check_size_factors(inf)


def test_cooks(inf):
    ec.check_cooks(inf)


def test_shrink_grid_fallback(inf):
    ec.check_shrink_grid(inf, _force(inf, 4))


def test_shrink_arguments(inf):
    ec.check_shrink_arguments(inf)


@pytest.mark.parametrize("p,N", [(1, 9), (2, 13), (3, 17), (4, 21), (5, 24), (6, 27), (7, 31), (8, 35), (9, 91), (12, 121), (16, 161)])
def test_every_design_width(inf, p, N):
    from oracle import nbglm

    ec.check_design_width(inf, nbglm.OracleInference(n_cpus=4), p, N)


def test_content_addressed_residency():
    """Host-buffer entry points key device copies of (N, G) buffers on a checksum of the full content: fresh host copies of the
    same content skip the upload (also for an OUTPUT handed back: device-side and host-side checksums must agree bit for bit),
    any change of the content does not, and results never depend on the cache."""
    import numpy as np

    from pydeseq2_b200.inference import B200Inference
    from pydeseq2_b200.synth import make_counts

    counts, X, _ = make_counts(64, 4096, "factorial", seed=3)  # 2 MB per (N, G) array: above the 1 MB threshold
    sf = np.exp(np.random.default_rng(0).normal(0, 0.1, 64))
    inf = B200Inference(device=0)
    ctx = inf._ops.ctx
    ctx.residency_clear()
    s0 = ctx.residency_stats()
    mu1 = inf.lin_reg_mu(counts, sf, X, 0.5)                      # counts: miss
    mu2 = inf.lin_reg_mu(np.array(counts), sf, X, 0.5)            # fresh copy, same content: hit
    s1 = ctx.residency_stats()
    assert s1["hits"] - s0["hits"] == 1 and s1["misses"] - s0["misses"] == 1
    np.testing.assert_array_equal(mu1, mu2)
    disp = np.full(counts.shape[1], 0.1)
    a1, c1 = inf.alpha_mle(np.array(counts), X, np.array(mu1), disp, 1e-8, 64.0)  # counts hit; mu_hat: OUTPUT of the call above
    s2 = ctx.residency_stats()
    assert s2["hits"] - s1["hits"] == 2, "an output handed back as a fresh copy must be found by its device-side checksum"
    changed = np.array(counts)
    changed[17, 123] += 1
    mu3 = inf.lin_reg_mu(changed, sf, X, 0.5)
    s3 = ctx.residency_stats()
    assert s3["hits"] == s2["hits"] and s3["misses"] - s2["misses"] == 1
    assert not np.array_equal(mu3[:, 123], mu1[:, 123]) and np.array_equal(np.delete(mu3, 123, 1), np.delete(mu1, 123, 1))
    swapped = np.array(counts)                                     # same multiset of words, other positions: must miss
    swapped[[0, 1]] = swapped[[1, 0]]
    inf.lin_reg_mu(swapped, sf, X, 0.5)
    assert ctx.residency_stats()["hits"] == s3["hits"]
    off = B200Inference(device=0)
    off._ops.ctx.check(off._ops.lib.pdq_residency_clear(off._ops.ctx.h))
    os_env = __import__("os").environ
    os_env["PDQ_RESIDENCY"] = "0"
    try:
        plain = B200Inference(device=0)
        a2, c2 = plain.alpha_mle(counts, X, mu1, disp, 1e-8, 64.0)
        assert plain._ops.ctx.residency_stats()["hits"] == 0 and plain._ops.ctx.residency_stats()["misses"] == 0
    finally:
        del os_env["PDQ_RESIDENCY"]
    np.testing.assert_array_equal(a1, a2)
    np.testing.assert_array_equal(c1, c2)


def test_many_samples_design_read_from_global_memory(inf):
    """N = 8 000 (p = 3): the design pack (384 KB) exceeds any shared-memory stage; the kernels read it from global memory.
    Every plugin method of the hot path against the oracle (reference: utils.py:273-438 has no sample limit)."""
    import os

    import numpy as np

    from oracle import nbglm
    from pydeseq2_b200.pipeline import median_of_ratios
    from pydeseq2_b200.synth import make_counts

    counts, X, _ = make_counts(8000, 64, "continuous", seed=21)
    counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
    sf = median_of_ratios(counts)[1]
    ora = nbglm.OracleInference(n_cpus=os.cpu_count())
    disp = np.full(counts.shape[1], 0.2)
    b, m, h, cv = inf.irls(counts, sf, X, disp, 0.5, 1e-8)
    rb, rm, rh, rcv = ora.irls(counts, sf, X, disp, 0.5, 1e-8)
    ok = (rcv == 1) & (cv == 1)
    assert ok.mean() > 0.9
    np.testing.assert_allclose(b[ok], rb[ok], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(m[:, ok], rm[:, ok], rtol=1e-6)
    np.testing.assert_allclose(h[:, ok], rh[:, ok], rtol=1e-6, atol=1e-12)
    a, ac = inf.alpha_mle(counts, X, np.ascontiguousarray(rm), disp, 1e-8, 8000.0)
    ra, rac = ora.alpha_mle(counts, X, np.ascontiguousarray(rm), disp, 1e-8, 8000.0)
    both = (ac == 1) & (rac == 1) & (ra > 1e-5)
    np.testing.assert_allclose(a[both], ra[both], rtol=1e-4)
    ridge = np.diag(np.repeat(1e-6, 3))
    got = inf.wald_test(X, disp, rb, rm, ridge, np.array([0.0, 0.0, 1.0]), 0.0, None)
    want = ora.wald_test(X, disp, rb, rm, ridge, np.array([0.0, 0.0, 1.0]), 0.0, None)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=1e-8, atol=1e-300)
    np.testing.assert_allclose(inf.lin_reg_mu(counts, sf, X, 0.5), ora.lin_reg_mu(counts, sf, X, 0.5), rtol=1e-9)


def test_csv_ingestion_into_pinned_memory(inf, tmp_path):
    """pydeseq2_b200.io: a genes-x-samples CSV parsed by native host threads into a page-locked (samples, genes) matrix that the
    plugin calls consume as it is."""
    import numpy as np
    import pandas as pd

    from pydeseq2_b200.io import read_counts_csv
    from pydeseq2_b200.synth import make_counts

    counts, X, _ = make_counts(24, 300, "two_level", seed=5)
    p = tmp_path / "counts.csv"
    pd.DataFrame(counts.T, index=[f"g{i}" for i in range(300)], columns=[f"s{j}" for j in range(24)]).to_csv(p)
    tab = read_counts_csv(p, ctx=inf._ops.ctx)
    np.testing.assert_array_equal(tab.counts, counts)
    assert tab.samples[0] == "s0" and tab.genes[-1] == "g299"
    sf = np.ones(24)
    np.testing.assert_array_equal(inf.lin_reg_mu(tab.counts, sf, X, 0.5), inf.lin_reg_mu(counts, sf, X, 0.5))


def test_trend_prior_grid_wide_equals_cluster_version(inf):
    """Vectors of >= 64 k genes run the trend + prior fit as one cooperative launch over all SMs (global-memory reductions, grid
    barrier); it must reproduce the one-cluster kernel (same algorithm, other reduction tree) -- also inside a replayed CUDA graph."""
    import os

    import numpy as np

    from pydeseq2_b200.pipeline import ResidentFit, median_of_ratios
    from pydeseq2_b200.synth import make_counts

    rng = np.random.default_rng(5)
    n = 150_001
    means = np.exp(rng.normal(4.0, 2.0, n))
    gw = np.clip((4.0 / means + 0.1) * np.exp(rng.normal(0, 0.6, n)), 1e-8, 60.0)
    gw[rng.choice(n, 300, replace=False)] = 1e-8     # flat-tail genes (excluded from the prior by the 100 * min_disp rule)
    means[rng.choice(n, 50, replace=False)] = np.nan  # padding of short shards
    out = {}
    for mode in ("0", "1"):
        os.environ["PDQ_TREND_GRID"] = mode
        try:
            out[mode] = inf.trend_and_prior(means, gw, 1e-8, 60.0, 60, 3)
        finally:
            del os.environ["PDQ_TREND_GRID"]
    a, b = out["0"], out["1"]
    assert a is not None and b is not None
    np.testing.assert_allclose(b[0], a[0], rtol=1e-10)          # coefficients
    np.testing.assert_allclose(b[1], a[1], rtol=1e-10, equal_nan=True)
    assert b[2] == a[2] and b[3] == a[3] and b[4] == a[4]        # exact medians: squared log residual, prior variance; rounds
    # inside the resident pass (70 000 genes > 64 k: grid version by default), eager then captured + replayed
    counts, X, _ = make_counts(24, 70_000, "two_level", seed=2)
    counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
    rf = ResidentFit(inf._ops.ctx, X, median_of_ratios(counts[:, :20000])[1])
    rf.upload(counts)
    r1, r2, r3 = rf.run(), rf.run(), rf.run()
    np.testing.assert_array_equal(r1["dispersions"], r3["dispersions"])
    np.testing.assert_array_equal(r1["trend"].coeffs, r3["trend"].coeffs)
    os.environ["PDQ_TREND_GRID"] = "0"
    try:
        rf._drop_graph()
        rf._eager_key = None
        rc = rf.run()
    finally:
        del os.environ["PDQ_TREND_GRID"]
    np.testing.assert_allclose(r3["trend"].coeffs, rc["trend"].coeffs, rtol=1e-10)
    np.testing.assert_allclose(r3["dispersions"], rc["dispersions"], rtol=1e-8)
    rf.close()
