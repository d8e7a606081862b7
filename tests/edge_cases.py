"""Edge cases the reference's tests cover at the orchestrator level (tests/test_edge_cases.py), restated at the plugin
boundary, plus the rarely taken branches (optimiser branch of irls, grid fallback of alpha_mle).  Shared by the CPU
suite (host emulator) and the GPU suite."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import nbglm
from parity import assert_close
from pydeseq2_b200.pipeline import fit_host, median_of_ratios
from pydeseq2_b200.synth import make_counts


def check_zero_genes(inf):
    """All-zero genes come back as NaN and do not disturb the other genes (reference tests/test_edge_cases.py:10-52)."""
    counts, X, _ = make_counts(40, 60, "two_level", seed=21)
    counts = counts[:, ~(counts == 0).all(0)]
    _, sf = median_of_ratios(counts)
    G = counts.shape[1]
    zero = np.zeros(G, dtype=bool)
    zero[np.random.default_rng(42).choice(G, G // 3, replace=False)] = True
    with_zeros = counts.copy()
    with_zeros[:, zero] = 0
    a = fit_host(with_zeros, X, inf, size_factors=sf)
    b = fit_host(np.ascontiguousarray(counts[:, ~zero]), X, inf, size_factors=sf)
    for name in ("lfc", "dispersions", "pvalue", "stat", "se"):
        va, vb = getattr(a, name), getattr(b, name)
        assert np.isnan(va[zero]).all(), name
        np.testing.assert_allclose(va[~zero], vb, rtol=1e-9, atol=1e-12, err_msg=name)


def check_no_replicates_raises(inf):
    """N == p: fit_rough_dispersions raises the reference's ValueError (utils.py:839-844)."""
    X = np.eye(3)
    with pytest.raises(ValueError, match="no replicates"):
        inf.fit_rough_dispersions(np.ones((3, 5)), X)


def check_empty_gene_set(inf):
    X = np.array([[1.0, 0], [1, 0], [1, 1], [1, 1]])
    sf = np.ones(4)
    c = np.zeros((4, 0), dtype=np.int64)
    b, m, h, cv = inf.irls(c, sf, X, np.zeros(0), 0.5, 1e-8)
    assert b.shape == (0, 2) and m.shape == (4, 0) and h.shape == (4, 0) and cv.shape == (0,)
    a, cv = inf.alpha_mle(c, X, np.zeros((4, 0)), np.zeros(0), 1e-8, 10.0)
    assert a.shape == (0,)
    assert inf.lin_reg_mu(c, sf, X, 0.5).shape == (4, 0)


def check_input_dtypes_and_layouts(inf):
    """int32 / float counts, pandas objects, Fortran-ordered and sliced inputs give the same answer."""
    import pandas as pd

    g = load_golden("calls_factorial_n30")
    c, X, sf, disp = g["counts"], g["X"], g["sf"], g["mom"]
    ref = inf.irls(c, sf, X, disp, 0.5, 1e-8)
    for variant in (c.astype(np.int32), c.astype(np.float64), np.asfortranarray(c), pd.DataFrame(c)):
        got = inf.irls(variant, pd.Series(sf), pd.DataFrame(X), pd.Series(disp), 0.5, 1e-8)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a, b)
    with pytest.raises(ValueError):
        inf.irls(c, sf[:-1], X, disp, 0.5, 1e-8)
    with pytest.raises(AssertionError):
        inf.irls(c, sf, X, disp, 0.5, 1e-8, optimizer="Nelder-Mead")  # utils.py:343
    with pytest.raises(KeyError):
        inf.wald_test(X, disp, g["lfc_beta"], g["lfc_mu"], g["ridge"], g["contrast"], 0.0, "sideways")


def check_nan_propagation_in_wald(inf):
    """All-zero genes reach wald_test as NaN rows (ds.py:320-347): NaN in, NaN out, other genes untouched."""
    g = load_golden("calls_two_level_n24")
    X, disp, lfc, mu = g["X"], g["disp"].copy(), g["lfc_beta"].copy(), g["lfc_mu"].copy()
    disp[3] = np.nan
    lfc[3] = np.nan
    mu[:, 3] = np.nan
    pv, st, se = inf.wald_test(X, disp, lfc, mu, g["ridge"], g["contrast"], 0.0, None)
    assert np.isnan(pv[3]) and np.isnan(st[3]) and np.isnan(se[3])
    keep = np.arange(len(disp)) != 3
    assert_close(st[keep], g["wald_two_sided_stat"][keep], 1e-9, "stat with a NaN gene present")


def check_irls_bounded_optimizer(inf, force):
    """The optimiser branch (utils.py:374-413).  `force(True)` sends every gene through it: the bounded minimiser of the
    convex NB objective must agree with scipy's L-BFGS-B on the genes where the reference reports success and the
    clamp max(mu, min_mu) is inactive (the objective is smooth there)."""
    g = load_golden("calls_factorial_n30")
    c, X, sf, disp = g["counts"], g["X"], g["sf"], g["disp"]
    plain = inf.irls(c, sf, X, disp, 0.5, 1e-8)
    force(True)
    try:
        b, m, h, cv = inf.irls(c, sf, X, disp, 0.5, 1e-8)
    finally:
        force(False)
    assert inf.last_irls_fallbacks == c.shape[1]
    unclamped = (plain[1] >= 0.5).all(0)
    ok = unclamped & (cv == 1)
    assert ok.sum() >= 0.6 * c.shape[1]
    # IRLS stops on a deviance-ratio test well before the optimum (beta moves by ~1e-4..1e-3 when it is tightened,
    # SURVEY.md App. B), so the two branches agree only to that level ...
    assert_close(b[ok], plain[0][ok], 5e-3, "optimiser branch vs IRLS", atol=1e-4)

    # ... while against the reference's OWN optimiser branch (maxiter=1 forces it, utils.py:374) the penalised objective
    # must be at least as low (scipy's L-BFGS-B stops at factr 1e7) and the coefficients close
    def f(beta, i):
        mu = np.maximum(sf * np.exp(X @ beta), 0.5)
        return nbglm.nb_nll(c[:, i], mu, disp[i]) + 0.5 * 1e-6 * (beta**2).sum()

    for i in np.flatnonzero(ok)[:12]:
        rb, _, _, rc = nbglm.irls_gene(c[:, i], sf, X, disp[i], 0.5, 1e-8, maxiter=1)
        assert f(b[i], i) <= f(rb, i) + 1e-9 * abs(f(rb, i)), i
        if rc:
            np.testing.assert_allclose(b[i], rb, rtol=2e-3, atol=2e-4)
    # active bounds: every coefficient stays inside the box, and the objective is no worse than scipy's bounded answer
    b2, _, _, cv2 = inf.irls(c[:, :8], sf, X, disp[:8], 0.5, 1e-8, min_beta=-0.7, max_beta=0.7)
    assert np.all(np.abs(b2) <= 0.7 + 1e-12)
    for i in range(8):
        rb, _, _, rc = nbglm.irls_gene(c[:, i], sf, X, disp[i], 0.5, 1e-8, -0.7, 0.7)
        assert f(b2[i], i) <= f(rb, i) + 1e-7 * abs(f(rb, i)), i


GRID_BETA = ["two_level_n24", "large_counts_n12"]


def check_irls_grid_fallback(inf, force, name):
    """utils.py:402-409: the optimiser reports failure on a two-column design -> `grid_fit_beta` (grid_search.py:145-221).
    `force(True)` sends every gene through the optimiser branch AND makes it report failure; the fixtures hold the REAL
    reference's grid answer (oracle/make_golden.py: main_grid_beta).  The coefficients are grid nodes: equal to the last bit,
    unless two nodes tie within rounding; mu / hat follow from them like the tail of irls_solver; converged = False."""
    g = load_golden("grid_beta_" + name)
    c, X, sf, disp, want = g["counts"], g["X"], g["sf"], g["disp"], g["beta"]
    force(True)
    try:
        b, m, h, cv = inf.irls(c, sf, X, disp, 0.5, 1e-8)
    finally:
        force(False)
    assert inf.last_irls_fallbacks == c.shape[1]
    assert (np.asarray(cv) == 0).all()
    np.testing.assert_allclose(b, want, rtol=0, atol=1e-12)
    for i in range(c.shape[1]):  # tail of irls_solver at the grid's coefficients (utils.py:423-438)
        mu = sf * np.exp(X @ want[i])
        muc = np.maximum(mu, 0.5)
        W = muc / (1.0 + muc * disp[i])
        Hinv = np.linalg.inv((X.T * W) @ X + 1e-6 * np.eye(2))
        hat = W * np.einsum("ij,jk,ik->i", X, Hinv, X)
        np.testing.assert_allclose(np.asarray(m)[:, i], mu, rtol=1e-10)
        np.testing.assert_allclose(np.asarray(h)[:, i], hat, rtol=1e-8, atol=1e-14)


def check_alpha_grid(inf, force):
    """Grid fallback (grid_search.py:54-142): forced for every gene, compared with the reference's grid search."""
    g = load_golden("calls_two_level_n24")
    c, X, mu, mom = g["counts"], g["X"], g["mu_hat"], g["mom"]
    N = c.shape[0]
    force(True)
    try:
        a, cv = inf.alpha_mle(c, X, mu, mom, 1e-8, float(max(10, N)))
    finally:
        force(False)
    assert (cv == 0).all()
    want = np.array([np.exp(nbglm.grid_fit_alpha(c[:, i], X, mu[:, i], mom[i], 1e-8, float(max(10, N)))) for i in range(c.shape[1])])
    # identical grid points; ties between neighbouring points may break differently (fine grid step = 0.42 %)
    close = np.isclose(a, want, rtol=1e-9)
    assert close.mean() >= 0.9
    assert np.all(np.abs(np.log(a) - np.log(want)) <= 2 * (np.log(max(10, N)) - np.log(1e-8)) / 99 * 2 / 99 + 1e-12)


def check_shrink_grid(inf, force):
    """apeGLM grid fallback (grid_search.py:224-318), forced for every gene of a two-column design: same grid nodes as the
    reference, so the result is the oracle's bit for bit unless two nodes tie to rounding."""
    for name, n in (("shrink_two_level_n24", 10), ("shrink_few_samples_n4", 6), ("shrinktape_large_counts", 5)):
        g = load_golden(name)
        c = np.ascontiguousarray(g["counts"][:, :n])
        args = (float(g["prior_no_shrink_scale"]), float(g["prior_scale"]), "L-BFGS-B", int(g["shrink_index"]))
        force(True)
        try:
            lfcs, ih, conv = inf.lfc_shrink_nbinom_glm(g["X"], c, g["size"][:n], g["offset"], *args)
        finally:
            force(False)
        assert (conv == 0).all() and inf.last_shrink_grid == n
        for i in range(n):
            b, h, cv = nbglm.nbinom_glm_gene(g["X"], c[:, i], g["size"][i], g["offset"], *args, force_grid=True)
            assert not cv
            np.testing.assert_allclose(lfcs[i], b, rtol=0, atol=2.1 * 60.0 / 59.0 * 2.0 / 59.0)  # at most one fine-grid node away
            if np.array_equal(lfcs[i], b):
                np.testing.assert_allclose(ih[i], h, rtol=1e-9)
        assert np.mean([np.array_equal(lfcs[i], nbglm.nbinom_glm_gene(g["X"], c[:, i], g["size"][i], g["offset"], *args,
                                                                      force_grid=True)[0]) for i in range(n)]) >= 0.8


def check_shrink_arguments(inf):
    g = load_golden("shrink_two_level_n24")
    a = (g["X"], g["counts"], g["size"], g["offset"], 15.0, 1.0)
    with pytest.raises(NotImplementedError):
        inf.lfc_shrink_nbinom_glm(*a, "Newton-CG", 1)
    with pytest.raises(IndexError):
        inf.lfc_shrink_nbinom_glm(*a, "L-BFGS-B", 2)
    with pytest.raises(ValueError):
        inf.lfc_shrink_nbinom_glm(g["X"], g["counts"], g["size"][:-1], g["offset"], 15.0, 1.0, "L-BFGS-B", 1)
    lfcs, ih, conv = inf.lfc_shrink_nbinom_glm(g["X"], g["counts"][:, :0], g["size"][:0], g["offset"], 15.0, 1.0, "L-BFGS-B", 1)
    assert lfcs.shape == (0, 2) and ih.shape == (0, 2, 2) and conv.shape == (0,)
    # genes are independent and the call is deterministic: a permutation of the genes permutes the results bit for bit
    perm = np.random.default_rng(0).permutation(g["counts"].shape[1])
    r1 = inf.lfc_shrink_nbinom_glm(*a, "L-BFGS-B", 1)
    r2 = inf.lfc_shrink_nbinom_glm(g["X"], np.ascontiguousarray(g["counts"][:, perm]), g["size"][perm], g["offset"], 15.0, 1.0,
                                   "L-BFGS-B", 1)
    for u, v in zip(r1, r2):
        np.testing.assert_array_equal(u[perm], v)


def check_size_factors(inf):
    """Median-of-ratios on the device: same values as the reference (golden `final_size_factors` of the reference's
    shipped datasets and seeded inputs), and -- the north star's criterion -- bit-identical RANKS across samples."""
    from oracle import nbglm  # the checker: restatement of preprocessing.py:5-102, pinned against the reference's goldens

    for name in ("tape_single_factor", "tape_continuous", "tape_wide"):
        t = load_golden(name)
        normed, sf = inf.size_factors(t["counts"])
        np.testing.assert_allclose(sf, t["final_size_factors"], rtol=1e-12)
        np.testing.assert_array_equal(np.argsort(np.argsort(sf)), np.argsort(np.argsort(t["final_size_factors"])))
    for N, G, seed in ((200, 3000, 0), (37, 501, 1), (8, 64, 2)):
        counts, _, _ = make_counts(N, G, "two_level", seed)
        want_normed, want = nbglm.deseq2_norm(counts)
        normed, sf, logmeans = inf.size_factors(counts, return_logmeans=True)
        np.testing.assert_allclose(sf, want, rtol=1e-12)
        np.testing.assert_array_equal(np.argsort(np.argsort(sf)), np.argsort(np.argsort(want)))
        np.testing.assert_allclose(normed, want_normed, rtol=1e-12)
        with np.errstate(divide="ignore"):  # deseq2_norm_fit's first output (preprocessing.py:31-59), -inf for genes holding a zero
            want_lm = np.log(counts).mean(0)
        np.testing.assert_array_equal(np.isinf(logmeans), np.isinf(want_lm))
        np.testing.assert_allclose(logmeans[~np.isinf(want_lm)], want_lm[~np.isinf(want_lm)], rtol=1e-12)
    # every gene holds a zero -> ValueError, like dds.fit_size_factors' fallback trigger
    bad = np.array([[0, 3, 5], [2, 0, 1], [4, 1, 0]], dtype=np.int64)
    with pytest.raises(ValueError, match="at least one zero"):
        inf.size_factors(bad)


def check_cooks(inf):
    """Cook's distances, trimmed-moments dispersions and the two per-gene decisions against the REAL orchestrator's layers
    (tape goldens) and, on seeded inputs with planted outliers, against the oracle."""
    for name in ("tape_single_factor", "tape_multi_factor", "tape_continuous", "tape_wide"):
        t = load_golden(name)
        nz = t["final_non_zero"] == 1
        counts = np.ascontiguousarray(t["counts"][:, nz])
        ck, disp, outl, repl = inf.calculate_cooks(counts, t["final_size_factors"], t["design"], t["final_mu_LFC"], t["final_hat"])
        assert_close(disp, t["final_robust_disp"], 1e-10, f"{name}: robust dispersions")
        assert_close(ck, t["final_cooks"][:, nz], 1e-9, f"{name}: cooks")
        np.testing.assert_array_equal(outl, t["final_cooks_outlier"][nz] == 1)
        np.testing.assert_array_equal(repl, t["final_replaced"][nz] == 1)
        ck2, disp2, outl2, repl2 = inf.calculate_cooks(counts, t["final_size_factors"], t["design"], t["final_mu_LFC"],
                                                       t["final_hat"], return_matrix=False)
        assert ck2 is None
        np.testing.assert_array_equal(disp2, disp)
    t = load_golden("tape_multi_factor_outliers")  # a design level with ONE replicate: that sample sits in no cell
    nz = t["final_non_zero"] == 1
    normed = t["counts"][:, nz] / t["final_size_factors"][:, None]
    _, disp, _, _ = inf.calculate_cooks(np.ascontiguousarray(t["counts"][:, nz]), t["final_size_factors"], t["design"],
                                        np.ones_like(normed), np.full_like(normed, 0.1))
    assert_close(disp, t["final_robust_disp"], 1e-10, "outlier tape: robust dispersions")
    # seeded inputs with planted outliers, several designs (cells of 100, of 25, and none at all)
    rng = np.random.default_rng(3)
    for N, G, kind in ((40, 300, "two_level"), (60, 200, "factorial"), (30, 100, "continuous"), (9, 64, "five")):
        counts, X, _ = make_counts(N, G, kind, seed=N)
        counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
        G = counts.shape[1]
        hit = rng.choice(G, G // 10, replace=False)
        counts[rng.integers(0, N, len(hit)), hit] *= 40
        counts[0, hit[:3]] += 5000
        normed, sf = median_of_ratios(counts)
        disp = np.clip(nbglm.fit_moments_dispersions(normed, sf), 1e-3, 10.0)
        ref_inf = nbglm.OracleInference(n_cpus=1)
        beta, mu, hat, _ = ref_inf.irls(counts, sf, X, disp, 0.5, 1e-8)
        mu, hat = np.ascontiguousarray(mu), np.ascontiguousarray(hat)
        want_ck, want_disp = nbglm.calculate_cooks(counts, normed, X, mu, hat)
        ck, rd, outl, repl = inf.calculate_cooks(counts, sf, X, mu, hat)
        assert_close(rd, want_disp, 1e-10, f"{kind}: robust dispersions")
        assert_close(ck, want_ck, 1e-9, f"{kind}: cooks")
        np.testing.assert_array_equal(outl, nbglm.cooks_outlier(counts, want_ck, X))
        from scipy.stats import f

        np.testing.assert_array_equal(repl, (want_ck > f.ppf(0.99, X.shape[1], N - X.shape[1])).any(0))
        if kind != "continuous" and N >= 30:
            assert outl.sum() >= 1


# --------------------------------------------------------------------------- random designs of every template width
def _design(rng, N, p):
    cols = [np.ones(N)]
    for j in range(1, p):
        if j % 3 == 0:
            cols.append(rng.normal(0, 1, N))                       # continuous covariate
        else:
            cols.append((rng.permutation(N) % 2).astype(float))    # balanced two-level factor
    X = np.stack(cols, axis=1)
    assert np.linalg.matrix_rank(X) == p
    return X


def _counts(rng, X, G):
    N, p = X.shape
    beta = np.concatenate([rng.normal(4, 1.5, (G, 1)) * np.log(2), rng.normal(0, 0.4, (G, p - 1))], axis=1)
    sf = np.exp(rng.normal(0, 0.2, N))
    mu = sf[:, None] * np.exp(X @ beta.T)
    alpha = 4 / np.exp(beta[:, 0]) + 0.1
    r = 1 / alpha
    y = rng.negative_binomial(r[None, :], r[None, :] / (r[None, :] + mu)).astype(np.int64)
    return np.ascontiguousarray(y[:, ~(y == 0).all(0)])



def check_design_width(emu, ora, p, N):
    """Every plugin method on a seeded random design of width p (see tests/test_emu_random_designs.py)."""
    rng = np.random.default_rng(100 + p)
    X = _design(rng, N, p)
    c = _counts(rng, X, 48)
    G = c.shape[1]
    normed, sf = median_of_ratios(c)
    max_disp = float(max(10, N))
    # method of moments
    np.testing.assert_allclose(emu.fit_rough_dispersions(normed, X), ora.fit_rough_dispersions(normed, X), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(emu.fit_moments_dispersions(normed, sf), ora.fit_moments_dispersions(normed, sf), rtol=1e-10)
    mom = np.clip(np.minimum(ora.fit_rough_dispersions(normed, X), ora.fit_moments_dispersions(normed, sf)), 1e-8, max_disp)
    # lin_reg_mu / irls
    np.testing.assert_allclose(emu.lin_reg_mu(c, sf, X, 0.5), ora.lin_reg_mu(c, sf, X, 0.5), rtol=1e-9)
    b, m, h, cv = emu.irls(c, sf, X, mom, 0.5, 1e-8)
    rb, rm, rh, rcv = ora.irls(c, sf, X, mom, 0.5, 1e-8)
    # IRLS oscillated to maxiter: the reference's L-BFGS-B branch (utils.py:374-403).  The emulator reports it per gene; the CUDA
    # path only counts such genes, so there they are recognised by their (small) distance from the reference's loose optimum
    st = getattr(emu._ops, "last_status", None)
    via_optimizer = (st != 0) if st is not None else ~np.isclose(b, rb, rtol=1e-6, atol=1e-9).all(axis=1)
    ok = (rcv == 1) & (cv == 1) & ~via_optimizer
    assert ok.mean() > 0.9
    # optimiser-branch genes: both sides minimise the same convex objective, the reference only to L-BFGS-B's loose tolerance
    opt = via_optimizer & (rcv == 1) & (cv == 1)
    np.testing.assert_allclose(b[opt], rb[opt], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(b[ok], rb[ok], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(m[:, ok], rm[:, ok], rtol=1e-6)
    np.testing.assert_allclose(h[:, ok], rh[:, ok], rtol=1e-6, atol=1e-12)
    # dispersions (genewise and MAP)
    mu_hat = np.ascontiguousarray(rm)
    a, ac = emu.alpha_mle(c, X, mu_hat, mom, 1e-8, max_disp)
    ra, rac = ora.alpha_mle(c, X, mu_hat, mom, 1e-8, max_disp)
    interior = (rac == 1) & (ac == 1) & (ra > 1e-5) & (ra < 0.99 * max_disp)
    assert interior.mean() > 0.5
    assert np.mean(np.isclose(a[interior], ra[interior], rtol=1e-4)) >= 0.97   # flat optima at small N stop path-dependently
    trend = np.clip(ra, 1e-8, max_disp) * np.exp(rng.normal(0, 0.3, G))
    a2, ac2 = emu.alpha_mle(c, X, mu_hat, trend, 1e-8, max_disp, prior_disp_var=0.5, cr_reg=True, prior_reg=True)
    ra2, rac2 = ora.alpha_mle(c, X, mu_hat, trend, 1e-8, max_disp, prior_disp_var=0.5, cr_reg=True, prior_reg=True)
    both = (rac2 == 1) & (ac2 == 1)
    np.testing.assert_allclose(a2[both], ra2[both], rtol=1e-4)
    # Wald on the oracle's fit, contrast on the last coefficient
    disp = np.clip(ra2, 1e-8, max_disp)
    contrast = np.zeros(p)
    contrast[-1] = 1.0
    ridge = np.diag(np.repeat(1e-6, p))
    for alt, null in ((None, 0.0), ("greaterAbs", 0.2), ("less", 0.1)):
        pv, st, se = emu.wald_test(X, disp, rb, np.ascontiguousarray(rm), ridge, contrast, null, alt)
        rp, rs, rse = ora.wald_test(X, disp, rb, np.ascontiguousarray(rm), ridge, contrast, null, alt)
        np.testing.assert_allclose(se, rse, rtol=1e-9)
        np.testing.assert_allclose(st, rs, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(pv, rp, rtol=1e-8, atol=1e-300)
    # apeGLM shrinkage of the last coefficient: same optimiser path as scipy's L-BFGS-B for every width
    lf, ih, scv = emu.lfc_shrink_nbinom_glm(X, c, 1.0 / disp, np.log(sf), 15, 0.4, "L-BFGS-B", p - 1)
    rl, rih, rscv = ora.lfc_shrink_nbinom_glm(X, c, 1.0 / disp, np.log(sf), 15, 0.4, "L-BFGS-B", p - 1)
    np.testing.assert_array_equal(scv, rscv)
    close = np.isclose(lf, rl, rtol=1e-6, atol=1e-9).all(axis=1)
    assert close.mean() >= 0.97, (p, np.abs(lf - rl).max())   # a stop-test flip would show as a ~1e-3 difference
    np.testing.assert_allclose(ih[close], rih[close], rtol=1e-5)
