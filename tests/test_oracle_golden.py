"""CPU suite: the oracle restatement (oracle/nbglm.py) against vectors produced by the REAL
reference (oracle/make_golden.py).  This is what pins the oracle (task spec ③)."""
import numpy as np
import pytest

from oracle import nbglm
from conftest import load_golden, tape_calls

CALLS = ["calls_two_level_n24", "calls_factorial_n30", "calls_continuous_n40", "calls_two_level_n200",
         "calls_large_counts_n12", "calls_five_columns_n36", "calls_intercept_n10", "calls_few_samples_n4"]
TAPES = ["tape_single_factor", "tape_multi_factor", "tape_continuous", "tape_wide", "tape_multi_factor_outliers"]
# the oracle runs the same scipy/numpy wheels as the reference: expect agreement to rounding
RTOL = 1e-9


@pytest.mark.parametrize("name", CALLS)
def test_oracle_calls(name):
    g = load_golden(name)
    inf = nbglm.OracleInference(n_cpus=1)
    c, X, sf, N = g["counts"], g["X"], g["sf"], g["counts"].shape[0]
    max_disp = max(10.0, N)
    np.testing.assert_allclose(nbglm.deseq2_norm(c)[1], sf, rtol=1e-12)
    np.testing.assert_allclose(inf.fit_rough_dispersions(g["normed"], X), g["rough"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(inf.fit_moments_dispersions(g["normed"], sf), g["moments"], rtol=1e-10)
    np.testing.assert_allclose(inf.lin_reg_mu(c, sf, X, 0.5), g["lin_mu"], rtol=1e-9)
    b, m, h, cv = inf.irls(c, sf, X, g["mom"], 0.5, 1e-8)
    np.testing.assert_allclose(b, g["irls0_beta"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(m, g["irls0_mu"], rtol=RTOL)
    np.testing.assert_allclose(h, g["irls0_hat"], rtol=1e-8)
    np.testing.assert_array_equal(cv, g["irls0_conv"])
    a, cv = inf.alpha_mle(c, X, g["mu_hat"], g["mom"], 1e-8, max_disp)
    np.testing.assert_allclose(a, g["gw_alpha"], rtol=RTOL)
    np.testing.assert_array_equal(cv, g["gw_conv"])
    a, cv = inf.alpha_mle(c, X, g["mu_hat"], g["trend"], 1e-8, max_disp, prior_disp_var=float(g["prior_var"]),
                          cr_reg=True, prior_reg=True)
    np.testing.assert_allclose(a, g["map_alpha"], rtol=RTOL)
    np.testing.assert_array_equal(cv, g["map_conv"])
    b, m, h, cv = inf.irls(c, sf, X, g["disp"], 0.5, 1e-8)
    np.testing.assert_allclose(b, g["lfc_beta"], rtol=RTOL, atol=1e-12)
    for alt, null in ((None, 0.0), ("greater", 0.3), ("less", 0.3), ("greaterAbs", 0.3), ("lessAbs", 0.3)):
        tag = alt or "two_sided"
        pv, st, se = inf.wald_test(X, g["disp"], g["lfc_beta"], g["lfc_mu"], g["ridge"], g["contrast"], null, alt)
        np.testing.assert_allclose(se, g[f"wald_{tag}_se"], rtol=1e-10)
        np.testing.assert_allclose(st, g[f"wald_{tag}_stat"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(pv, g[f"wald_{tag}_p"], rtol=1e-9, atol=1e-300)
    co, pred, ok = inf.dispersion_trend_gamma_glm(1.0 / g["normed"].mean(0), np.clip(g["gw_alpha"], 1e-8, max_disp))
    np.testing.assert_allclose(co, g["trend_coeffs"], rtol=1e-7)


@pytest.mark.parametrize("name", TAPES)
def test_oracle_replays_reference_tape(name):
    """Replay every Inference call the real orchestrator made on the reference's shipped datasets."""
    t = load_golden(name)
    inf = nbglm.OracleInference(n_cpus=1)
    seen = set()
    for meth, args, kw, outs in tape_calls(t):
        seen.add(meth)
        res = getattr(inf, meth)(*args, **kw)
        res = res if isinstance(res, tuple) else (res,)
        for got, want in zip(res, outs):
            np.testing.assert_allclose(np.asarray(got, dtype=float), want, rtol=1e-7, atol=1e-12, equal_nan=True,
                                       err_msg=f"{name}:{meth}")
    assert {"irls", "alpha_mle", "wald_test", "fit_rough_dispersions", "fit_moments_dispersions"} <= seen


@pytest.mark.parametrize("name", ["two_level_n24", "large_counts_n12"])
def test_oracle_grid_fit_beta(name):
    """The restated `grid_fit_beta` (grid_search.py:145-221) lands on the real reference's grid node."""
    g = load_golden("grid_beta_" + name)
    for i in range(g["counts"].shape[1]):
        got = nbglm.grid_fit_beta(g["counts"][:, i], g["sf"], g["X"], g["disp"][i])
        np.testing.assert_allclose(got, g["beta"][i], rtol=0, atol=1e-12)


def test_nb_nll_is_a_normalised_pmf():
    """Same property the reference's only hot-path unit test checks (tests/test_utils.py:11-33)."""
    for mu, alpha in ((3.0, 0.5), (40.0, 0.05), (0.7, 2.0)):
        y = np.arange(0, 4000)
        logp = -np.array([nbglm.nb_nll(np.array([k]), np.array([mu]), alpha) for k in y])
        pmf = np.exp(logp)
        assert abs(pmf.sum() - 1) < 1e-6
        assert abs((pmf * y).sum() - mu) < 1e-4 * mu
        assert abs((pmf * (y - mu) ** 2).sum() - (mu + alpha * mu**2)) < 1e-3 * (mu + alpha * mu**2)


@pytest.mark.parametrize("name", TAPES)
def test_oracle_cooks(name):
    """Cook's distances / robust dispersions / Cook's p-value filter against the real orchestrator's layers."""
    t = load_golden(name)
    nz = t["final_non_zero"] == 1
    counts = t["counts"][:, nz]
    X, sf = t["design"], t["final_size_factors"]
    normed = counts / sf[:, None]
    np.testing.assert_allclose(nbglm.robust_method_of_moments_disp(normed, X), t["final_robust_disp"], rtol=1e-12)
    if name == "tape_multi_factor_outliers":
        return  # after a refit the stored mu/hat belong to the refitted genes; the pre-refit cooks layer is kept as is
    cooks, disp = nbglm.calculate_cooks(counts, normed, X, t["final_mu_LFC"], t["final_hat"])
    np.testing.assert_allclose(cooks, t["final_cooks"][:, nz], rtol=1e-10)
    np.testing.assert_array_equal(nbglm.cooks_outlier(counts, cooks, X), t["final_cooks_outlier"][nz] == 1)


# --------------------------------------------------------------------------- apeGLM shrinkage
from parity import SHRINK  # noqa: E402


@pytest.mark.parametrize("name", SHRINK)
def test_oracle_lfc_shrink(name):
    g = load_golden(name)
    k = int(g["shrink_index"])
    lfcs, ih, conv = nbglm.OracleInference(n_cpus=1).lfc_shrink_nbinom_glm(
        g["X"], g["counts"], g["size"], g["offset"], float(g["prior_no_shrink_scale"]), float(g["prior_scale"]), "L-BFGS-B", k)
    np.testing.assert_allclose(lfcs, g["lfcs"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(ih, g["inv_hessians"], rtol=RTOL)
    np.testing.assert_array_equal(conv, g["converged"])
    if not np.isnan(g["prior_var"]):  # the prior scale the reference derived from the MLE LFCs and their SEs (ds.py:551-585)
        col = g["mle_lfc"][:, k] if g["mle_lfc"].ndim == 2 else g["mle_lfc"]
        assert nbglm.fit_shrink_prior_var(col, g["mle_se"]) == pytest.approx(float(g["prior_var"]), rel=1e-9)


@pytest.mark.parametrize("name", ["shrink_two_level_n24", "shrink_five_columns_n36", "shrink_large_counts_n12", "shrinktape_continuous"])
def test_restated_lbfgsb_walks_scipys_path(name):
    """oracle/lbfgsb_restated.py (the specification of the CUDA optimiser) against scipy's L-BFGS-B on the reference's
    objective: same number of iterations and evaluations, same end point."""
    from scipy.optimize import minimize

    from oracle.lbfgsb_restated import minimize_lbfgsb_unbounded

    g = load_golden(name)
    X, p = g["X"], g["X"].shape[1]
    for i in range(g["counts"].shape[1]):
        args = (X, g["counts"][:, i], g["size"][i], g["offset"], float(g["prior_no_shrink_scale"]), float(g["prior_scale"]),
                int(g["shrink_index"]))
        cn = np.maximum(nbglm.nbinom_fn(np.zeros(p), *args), 1)
        fun = lambda b: nbglm.nbinom_fn(b, *args) / cn  # noqa: E731
        jac = lambda b: nbglm.nbinom_grad(b, *args) / cn  # noqa: E731
        x0 = np.ones(p) * 0.1 * (-1) ** np.arange(p)
        ref = minimize(fun, x0, jac=jac, method="L-BFGS-B", options={"ftol": 1e-8, "gtol": 1e-8})
        x, ok, nit, nfev = minimize_lbfgsb_unbounded(fun, jac, x0)
        assert (nit, nfev, ok) == (ref.nit, ref.nfev, ref.success)
        np.testing.assert_allclose(x, ref.x, rtol=1e-8, atol=1e-12)
