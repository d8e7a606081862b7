"""Shared parity checks: the same assertions run against the host emulator (CPU suite) and against the
CUDA library through the C ABI (`-m gpu`).  Tolerances follow BASELINE.json's north star: LFCs, dispersions
and Wald statistics within 1e-4 relative of the reference CPU backend (the per-call checks below are much
tighter where the arithmetic allows it)."""
import numpy as np

from conftest import tape_calls

ALTS = ((None, 0.0), ("greater", 0.3), ("less", 0.3), ("greaterAbs", 0.3), ("lessAbs", 0.3))

# per-call isolation tolerances (identical inputs into reference and device code)
TOL_BETA = 1e-6     # same start, same update, same stopping rule -> differences are rounding only
TOL_MU = 1e-6
TOL_HAT = 1e-6
TOL_ALPHA = 2e-5    # L-BFGS-B itself stops within ~3e-6 of the optimum; bar is 1e-4
TOL_WALD = 1e-9


def rel_err(got, want, floor=1e-12):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    both_nan = np.isnan(got) & np.isnan(want)
    err = np.abs(got - want) / np.maximum(np.abs(want), floor)
    err[both_nan] = 0.0
    return err


def assert_close(got, want, rtol, what, atol=0.0, mask=None):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    bad = ~(np.isclose(got, want, rtol=rtol, atol=atol, equal_nan=True))
    if mask is not None:
        bad &= mask
    assert not bad.any(), (
        f"{what}: {bad.sum()} / {bad.size} entries off by more than rtol={rtol}; worst rel err "
        f"{np.nanmax(rel_err(got, want)[bad]):.3e} at {np.argwhere(bad)[:5].tolist()}"
    )


def check_calls(inf, g, tol_alpha=TOL_ALPHA):
    """`g` is a calls_*.npz golden (outputs of the REAL reference); `inf` a B200Inference."""
    c, X, sf, N = g["counts"], g["X"], g["sf"], g["counts"].shape[0]
    max_disp = max(10.0, N)
    assert_close(inf.fit_rough_dispersions(g["normed"], X), g["rough"], 1e-8, "rough", atol=1e-12)
    assert_close(inf.fit_moments_dispersions(g["normed"], sf), g["moments"], 1e-9, "moments", atol=1e-12)
    assert_close(inf.lin_reg_mu(c, sf, X, 0.5), g["lin_mu"], 1e-9, "lin_reg_mu")

    for tag, disp in (("irls0", g["mom"]), ("lfc", g["disp"])):
        b, m, h, cv = inf.irls(c, sf, X, disp, 0.5, 1e-8)
        # genes the reference itself flags as not converged (IRLS hit maxiter, then L-BFGS-B gave up on the kinked
        # clamped objective, utils.py:374-403) carry an optimiser-path-dependent value: flags must agree, values need not
        ok = g[f"{tag}_conv"] == 1.0
        assert_close(b[ok], g[f"{tag}_beta"][ok], TOL_BETA, f"{tag} beta", atol=1e-9)
        assert_close(m[:, ok], g[f"{tag}_mu"][:, ok], TOL_MU, f"{tag} mu", atol=1e-12)
        assert_close(h[:, ok], g[f"{tag}_hat"][:, ok], TOL_HAT, f"{tag} hat", atol=1e-12)
        np.testing.assert_array_equal(cv, g[f"{tag}_conv"])
        assert np.isfinite(b).all()
        assert cv.dtype == np.float64

    lo, hi = 1e-8, max_disp
    for tag, kw, ahat in (("gw", {}, g["mom"]),
                          ("map", dict(prior_disp_var=float(g["prior_var"]), cr_reg=True, prior_reg=True), g["trend"])):
        a, cv = inf.alpha_mle(c, X, g["mu_hat"], ahat, lo, hi, **kw)
        want = g[f"{tag}_alpha"]
        ok = g[f"{tag}_conv"] == 1.0
        # at tiny dispersions the reference's own gradient is rounding noise (1/alpha^2 cancellation): compare
        # those genes after the caller's clip only (dds.py:792-794)
        interior = ok & (want > 1e-5) & (want < hi * (1 - 1e-9))
        assert_close(a[interior], want[interior], tol_alpha, f"{tag} alpha (interior optimum)")
        edge = ok & ~interior
        assert np.all(np.clip(a[edge], lo, hi) < 1e-4) or not edge.any() or np.allclose(
            np.clip(a[edge], lo, hi), np.clip(want[edge], lo, hi), rtol=1e-3), f"{tag} alpha (bound cases)"
        assert (cv[ok] == 1.0).all()

    for alt, null in ALTS:
        tag = alt or "two_sided"
        pv, st, se = inf.wald_test(X, g["disp"], g["lfc_beta"], g["lfc_mu"], g["ridge"], g["contrast"], null, alt)
        assert_close(se, g[f"wald_{tag}_se"], TOL_WALD, f"wald {tag} se")
        assert_close(st, g[f"wald_{tag}_stat"], TOL_WALD, f"wald {tag} stat", atol=1e-14)
        assert_close(pv, g[f"wald_{tag}_p"], 1e-8, f"wald {tag} p", atol=1e-300)


def check_tape(inf, t, name):
    """Replay every Inference call the reference orchestrator made (tape_*.npz)."""
    for meth, args, kw, outs in tape_calls(t):
        res = getattr(inf, meth)(*args, **kw)
        res = res if isinstance(res, tuple) else (res,)
        for k, (got, want) in enumerate(zip(res, outs)):
            got = np.asarray(got, dtype=float)
            if got.ndim == 0:
                got, want = got.reshape(1), np.ravel(want)
            if meth == "alpha_mle" and k == 0:
                conv = outs[1] == 1.0
                assert_close(got[conv], want[conv], 1e-4, f"{name}:{meth} alpha")
            elif meth == "alpha_mle":
                assert (got[outs[1] == 1.0] == 1.0).all()
            elif meth == "dispersion_trend_gamma_glm":
                # the device fit converges to the minimiser; scipy's L-BFGS-B (the reference) stops within ~1e-5 of it
                assert_close(got, want, 1e-4, f"{name}:{meth}[{k}]")
            elif meth == "wald_test" and k == 0:
                assert_close(got, want, 1e-7, f"{name}:{meth} p", atol=1e-300)
            else:
                assert_close(got, want, 1e-6, f"{name}:{meth}[{k}]", atol=1e-10)


SHRINK = ["shrink_two_level_n24", "shrink_factorial_n30", "shrink_factorial_n30_idx1", "shrink_continuous_n40",
          "shrink_two_level_n200", "shrink_two_level_n200_noadapt", "shrink_large_counts_n12", "shrink_five_columns_n36",
          "shrink_few_samples_n4", "shrinktape_single_factor", "shrinktape_single_factor_noadapt", "shrinktape_multi_factor",
          "shrinktape_continuous", "shrinktape_large_counts"]
# apeGLM: the backend walks the reference's optimiser path, so it lands on the reference's (loosely converged) iterate:
# measured <= 5e-10 on every fixture; the tolerance leaves room for a different summation order on the device
TOL_SHRINK = 1e-7


def check_shrink(inf, g, tol=TOL_SHRINK):
    """`lfc_shrink_nbinom_glm` against the real reference's outputs (oracle/make_golden.py `gen_shrink_*`)."""
    k = int(g["shrink_index"])
    lfcs, ih, conv = inf.lfc_shrink_nbinom_glm(g["X"], g["counts"], g["size"], g["offset"], float(g["prior_no_shrink_scale"]),
                                               float(g["prior_scale"]), "L-BFGS-B", k)
    assert lfcs.shape == g["lfcs"].shape and ih.shape == g["inv_hessians"].shape
    np.testing.assert_array_equal(conv, g["converged"])
    assert_close(lfcs, g["lfcs"], tol, "shrunk coefficients", atol=1e-10)
    assert_close(ih, g["inv_hessians"], tol, "inverse Hessians", atol=1e-14)
    if "r_log2FoldChange" in g:  # the reference's own test: within 2 % of R's apeglm (tests/test_pydeseq2.py:256-296 ...)
        got = lfcs[:, k] / np.log(2)
        assert np.nanmax(np.abs(g["r_log2FoldChange"] - got) / np.abs(g["r_log2FoldChange"])) < 0.02
        se = np.sqrt(np.abs(ih[:, k, k])) / np.log(2)
        np.testing.assert_allclose(se, g["final_lfcSE"], rtol=tol)
