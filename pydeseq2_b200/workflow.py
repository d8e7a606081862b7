"""``DeseqDataSet.deseq2()`` + ``DeseqStats.summary()`` end to end on bare arrays, for any ``Inference`` backend.

`pipeline.fit_host` covers the hot path (dispersions, LFCs, Wald).  This module adds the steps the reference's orchestrator
runs around it, in the reference's order, so that a table computed here can be compared with ``DeseqStats.results_df``:

* Cook's distances and the outlier refit -- ``calculate_cooks`` (dds.py:986-1040), ``_replace_outliers`` (dds.py:1301-1358),
  ``_refit_without_outliers`` (dds.py:1360-1458), ``cooks_outlier`` (dds.py:1066-1110)          [SURVEY.md §8 f-1]
* p-value post-processing -- ``_cooks_filtering`` (ds.py:544-549), ``_independent_filtering`` (ds.py:486-528) with its lowess
  smoother (utils.py:1379-1442), ``_p_value_adjustment`` (ds.py:530-542)                         [SURVEY.md §8 f-4]
* the results table of ``summary()`` (ds.py:278-285), optionally with apeGLM-shrunk LFCs (ds.py:363-443).

Everything per gene and heavy goes through the backend (plugin calls + the two extra methods ``calculate_cooks`` /
``lfc_shrink_nbinom_glm``); what stays here is G-length bookkeeping on a handful of genes, exactly the part the reference keeps in
its orchestrator.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import ceil, floor

import numpy as np

from .pipeline import LN2, FitResult, fit_host, fit_shrink_prior_var, lin_mu_branch


def n_or_more_replicates(X: np.ndarray, min_replicates: int) -> np.ndarray:
    """Per sample: does its design row occur at least ``min_replicates`` times (utils.py:888-911)."""
    _, inv, cnt = np.unique(np.asarray(X), axis=0, return_inverse=True, return_counts=True)
    return cnt[np.ravel(inv)] >= min_replicates


def trimmed_mean_rows(x: np.ndarray, trim: float) -> np.ndarray:
    """Mean over axis 0 after dropping floor(n * trim) values at each end (utils.py:567-599, ``axis=0`` form)."""
    n = x.shape[0]
    k = floor(n * trim)
    return np.sort(x, axis=0)[k:n - k].mean(axis=0)


def lowess(x: np.ndarray, y: np.ndarray, frac: float = 2.0 / 3.0, iters: int = 3) -> np.ndarray:
    """Robust locally weighted linear regression as the reference runs it (utils.py:1379-1442): tricube weights with the
    bandwidth at the sorted distance of index ceil(frac * n), ``iters`` bisquare re-weightings; each local line comes from
    the minimum-norm solution of its 2 x 2 normal equations."""
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    n = len(x)
    r = int(ceil(frac * n))
    dist = np.abs(x[:, None] - x[None, :])
    h = np.maximum(np.sort(dist, axis=0)[r], 1e-12)           # h[i]: distance from x[i] to its r-th neighbour
    w = np.clip(np.abs(np.nan_to_num((x[:, None] - x[None, :]) / h)), 0.0, 1.0)
    w = (1 - w**3) ** 3                                        # w[:, i]: weights of all points around x[i]
    est, delta = np.zeros(n), np.ones(n)
    for _ in range(iters):
        for i in range(n):
            wi = delta * w[:, i]
            A = np.array([[wi.sum(), (wi * x).sum()], [(wi * x).sum(), (wi * x * x).sum()]])
            b = np.array([(wi * y).sum(), (wi * y * x).sum()])
            c0, c1 = np.linalg.lstsq(A, b, rcond=None)[0]
            est[i] = c0 + c1 * x[i]
        res = y - est
        s = np.median(np.abs(res))
        delta = (np.abs(res) > 0).astype(float) if s == 0 else np.clip(res / (6.0 * s), -1, 1)
        delta = (1 - delta**2) ** 2
    return est


def bh_adjust(p: np.ndarray) -> np.ndarray:
    """Benjamini-Hochberg adjusted p-values (what ``scipy.stats.false_discovery_control(method="bh")`` returns)."""
    p = np.asarray(p, dtype=float)
    m = p.size
    order = np.argsort(p)
    adj = p[order] * m / np.arange(1, m + 1)
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    out = np.empty(m)
    out[order] = np.clip(adj, 0.0, 1.0)
    return out


def independent_filtering(base_mean: np.ndarray, pvalues: np.ndarray, alpha: float = 0.05) -> np.ndarray:
    """Adjusted p-values with the mean-count filter chosen to maximise rejections (ds.py:486-528)."""
    lower = float(np.mean(base_mean == 0))
    upper = 0.95 if lower < 0.95 else 1.0
    theta = np.linspace(lower, upper, 50)
    cutoffs = np.quantile(base_mean, theta)
    G = len(pvalues)
    padj = np.full((G, len(theta)), np.nan)
    tested = ~np.isnan(pvalues)
    for i, cut in enumerate(cutoffs):
        use = (base_mean >= cut) & tested
        if use.any():
            padj[use, i] = bh_adjust(pvalues[use])
    num_rej = (padj < alpha).sum(0).astype(int)
    fit = lowess(theta, num_rej, frac=1 / 5)
    j = 0
    if num_rej.max() > 10:
        resid = num_rej[num_rej > 0] - fit[num_rej > 0]
        thresh = fit.max() - np.sqrt(np.mean(resid**2))
        above = np.where(num_rej > thresh)[0]
        if above.size:
            j = int(above[0])
    return padj[:, j]


@dataclass
class Deseq2Results:
    """Columns of ``DeseqStats.results_df`` (ds.py:278-285) plus what ``DeseqDataSet.var`` holds after ``deseq2()``."""
    base_mean: np.ndarray
    log2_fold_change: np.ndarray
    lfc_se: np.ndarray           # log2 scale
    stat: np.ndarray
    pvalue: np.ndarray
    padj: np.ndarray
    lfc: np.ndarray              # (G, p) natural log
    dispersions: np.ndarray
    genewise_dispersions: np.ndarray
    fitted_dispersions: np.ndarray
    size_factors: np.ndarray
    non_zero: np.ndarray
    replaced: np.ndarray
    refitted: np.ndarray
    cooks_outlier: np.ndarray
    cooks: np.ndarray | None
    fit: FitResult = field(repr=False, default=None)
    shrink_prior_scale: float | None = None


def deseq2_results(counts, X, inference, contrast=None, size_factors=None, refit_cooks=True, min_replicates=7, cooks_filter=True,
                   independent_filter=True, alpha=0.05, lfc_null=0.0, alt_hypothesis=None, fit_type="parametric", min_mu=0.5,
                   min_disp=1e-8, max_disp=10.0, beta_tol=1e-8, shrink_coeff=None, shrink_adapt=True, keep_cooks=False,
                   comm=None) -> Deseq2Results:
    """The reference's default analysis -- ``dds.deseq2()`` then ``DeseqStats(dds, contrast).summary()`` (and ``lfc_shrink`` when
    ``shrink_coeff`` is given) -- through ``inference``.  ``counts`` (N, G) non-negative integers, ``X`` (N, p) expanded design,
    ``contrast`` (p,) numeric contrast vector (default: last coefficient).

    ``comm`` (``sharding.TorchDistComm`` / ``NcclComm``): ``counts`` is this rank's gene shard and ``size_factors`` must be given.
    Everything per gene -- fits, Cook's distances, the outlier refit -- stays local; the three steps that look at ALL genes are
    exchanged: the dispersion trend / prior (inside ``fit_host``), the multiple-testing step (base means and p-values are
    all-gathered, adjusted identically on every rank, and the local slice is kept) and the apeGLM prior scale."""
    from scipy.stats import f as f_dist

    counts = np.ascontiguousarray(counts, dtype=np.int64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    N, G = counts.shape
    p = X.shape[1]
    if contrast is None:
        contrast = np.zeros(p)
        contrast[-1] = 1.0
    contrast = np.asarray(contrast, dtype=float)
    if lfc_null < 0 and alt_hypothesis in ("greaterAbs", "lessAbs"):
        raise ValueError(f"The alternative hypothesis being {alt_hypothesis}, please provide a positive lfc_null value (got {lfc_null}).")
    if comm is not None and size_factors is None:
        raise ValueError("gene shards: size factors are per sample and global over genes -- pass them in")
    fit = fit_host(counts, X, inference, contrast=contrast, size_factors=size_factors, min_mu=min_mu, min_disp=min_disp, max_disp=max_disp,
                   beta_tol=beta_tol, fit_type=fit_type, lfc_null=lfc_null, alt_hypothesis=alt_hypothesis, comm=comm)
    max_disp = max(max_disp, N)
    sf, nz = fit.size_factors, fit.non_zero
    lfc, disp = fit.lfc.copy(), fit.dispersions.copy()
    genewise, fitted = _full(fit.genewise, nz, G), _full(fit.trend.fitted, nz, G)
    base_mean = np.zeros(G)
    base_mean[nz] = fit.normed_means                     # all-zero genes have mean 0
    pv, stat, se = fit.pvalue.copy(), fit.stat.copy(), fit.se.copy()

    # ---- Cook's distances (dds.py:986-1040) on the non-zero genes
    c_nz = counts if nz.all() else np.ascontiguousarray(counts[:, nz])
    ck = inference.calculate_cooks(c_nz, sf, X, fit.mu_lfc, fit.hat)[0]
    cooks = np.full((N, G), np.nan)
    cooks[:, nz] = ck
    cutoff = f_dist.ppf(0.99, p, N - p)
    above = cooks > cutoff

    # ---- outlier replacement and refit (dds.py:1042-1064, 1301-1458)
    replaced, refitted, new_zero = np.zeros(G, bool), np.zeros(G, bool), np.zeros(G, bool)
    replace_cooks = None
    if refit_cooks:
        replaceable = n_or_more_replicates(X, min_replicates)
        if replaceable.any():
            replaced = above.any(axis=0)
        if replaced.any():
            sub = counts[:, replaced].copy()
            base = trimmed_mean_rows(sub / sf[:, None], 0.2)
            repl = (base[None, :] * sf[:, None]).astype(int)   # truncation towards zero, like DataFrame.astype(int)
            mask = replaceable[:, None] & above[:, replaced]
            sub[mask] = repl[mask]
            zero_now = (sub == 0).all(axis=0)
            ridx = np.flatnonzero(replaced)
            new_zero[ridx[zero_now]] = True
            refitted[ridx[~zero_now]] = True
            base_mean[new_zero] = 0.0
            lfc[new_zero] = 0.0
            if refitted.any():
                r = _refit_subset(np.ascontiguousarray(sub[:, ~zero_now]), X, sf, inference, fit, min_mu, min_disp, max_disp, beta_tol)
                base_mean[refitted], lfc[refitted] = r["normed_means"], r["lfc"]
                genewise[refitted], fitted[refitted], disp[refitted] = r["genewise"], r["fitted"], r["disp"]
                replace_cooks = cooks.copy()
                replace_cooks[np.ix_(replaceable, refitted)] = 0.0
                # Wald statistics of the refitted genes (per-gene independent: same as re-running run_wald_test on every gene)
                mu_r = np.exp(X @ r["lfc"].T) * sf[:, None]
                pv[refitted], stat[refitted], se[refitted] = inference.wald_test(
                    X, r["disp"], r["lfc"], mu_r, np.diag(np.repeat(1e-6, p)), contrast, LN2 * lfc_null, alt_hypothesis)
            if new_zero.any():
                # run_wald_test (ds.py:355-360) on genes whose replacement made them all-zero
                se[new_zero], stat[new_zero], pv[new_zero] = 0.0, 0.0, 1.0

    # ---- which genes lose their p-value (dds.py:1066-1110)
    use_for_max = n_or_more_replicates(X, 3)
    src = replace_cooks if (refit_cooks and refitted.any() and replace_cooks is not None) else cooks
    outlier = (src[use_for_max] > cutoff).any(axis=0)
    if outlier.any():
        pos = cooks[:, outlier].argmax(0)
        top = counts[:, outlier][pos, np.arange(len(pos))]
        outlier[outlier] = (counts[:, outlier] > top).sum(0) < 3
    if cooks_filter:
        pv[outlier] = np.nan

    padj = _adjust(base_mean, pv, independent_filter, alpha, comm)

    res = Deseq2Results(base_mean, lfc @ contrast / LN2, se / LN2, stat, pv, padj, lfc, disp, genewise, fitted, sf, nz, replaced, refitted,
                        outlier, cooks if keep_cooks else None, fit)
    if shrink_coeff is not None:   # DeseqStats.lfc_shrink (ds.py:363-443): LFC column and SE replaced, p-values untouched
        k = int(shrink_coeff)
        scale = 1.0
        if shrink_adapt:
            lk, sk = lfc[:, k], se
            if comm is not None:
                g = comm.allgather_table({"l": lk, "s": sk})
                lk, sk = g["l"], g["s"]
            scale = float(np.minimum(np.sqrt(fit_shrink_prior_var(lk, sk)), 1))
        sh, ih, _ = inference.lfc_shrink_nbinom_glm(X, c_nz, 1.0 / disp[nz], np.log(sf), 15, scale, "L-BFGS-B", k)
        lfc[nz, k] = np.asarray(sh)[:, k]
        se = se.copy()
        se[nz] = np.sqrt(np.abs(np.asarray(ih)[:, k, k]))
        res.lfc, res.log2_fold_change, res.lfc_se, res.shrink_prior_scale = lfc, lfc[:, k] / LN2, se / LN2, scale
    return res


def _adjust(base_mean, pv, independent_filter, alpha, comm=None):
    """Multiple testing (ds.py:486-542): global over the genes of all shards."""
    bm_all, pv_all = base_mean, pv
    if comm is not None:
        g = comm.allgather_table({"bm": base_mean, "pv": pv})
        bm_all, pv_all = g["bm"], g["pv"]
    if independent_filter:
        padj = independent_filtering(bm_all, pv_all, alpha)
    else:
        padj = np.full(len(pv_all), np.nan)
        ok = ~np.isnan(pv_all)
        padj[ok] = bh_adjust(pv_all[ok])
    if comm is not None:
        padj = padj[comm.local_slice()]
    return padj


def deseq2_results_resident(rf, contrast=None, refit_cooks=True, min_replicates=7, cooks_filter=True, independent_filter=True,
                            alpha=0.05, lfc_null=0.0, alt_hypothesis=None, fit_type="parametric", non_zero=None,
                            adjust=True) -> Deseq2Results:
    """:func:`deseq2_results` on a :class:`pipeline.ResidentFit` whose counts are already in HBM (``rf.upload(counts)``, all-zero
    genes dropped by the caller): the hot path, Cook's distances and their per-gene decisions run in ONE resident pass; only
    per-gene vectors come back.  For the few genes whose outlier counts get replaced (dds.py:1301-1358) the mu / hat columns are
    gathered on the device, the replacement is decided on the host (R columns), and the refit (dds.py:1360-1458) runs resident
    again on the compact matrix.  Multiple testing is G-length host work as in :func:`deseq2_results`.  ``non_zero``: boolean mask of
    the uploaded genes within the caller's full gene list (all-zero genes are not fitted but take part in the independent
    filtering with base mean 0, ds.py:486-528); the returned tables then have the full length."""
    from scipy.stats import f as f_dist

    X, N, p, G = rf.X, rf.N, rf.p, rf.G
    sf = rf.sf
    if contrast is None:
        contrast = np.zeros(p)
        contrast[-1] = 1.0
    contrast = np.asarray(contrast, dtype=float)
    if lfc_null < 0 and alt_hypothesis in ("greaterAbs", "lessAbs"):
        raise ValueError(f"The alternative hypothesis being {alt_hypothesis}, please provide a positive lfc_null value (got {lfc_null}).")
    was = rf.with_cooks
    rf.with_cooks = True
    try:
        r = rf.run(contrast=contrast, lfc_null=lfc_null, alt_hypothesis=alt_hypothesis, fit_type=fit_type)
    finally:
        rf.with_cooks = was
    base_mean = np.array(r["normed_means"])
    lfc, disp = np.array(r["lfc"]), np.array(r["dispersions"])
    genewise, fitted = np.array(r["genewise"]), np.array(r["fitted"])
    pv, stat, se = np.array(r["pvalue"]), np.array(r["stat"]), np.array(r["se"])
    outlier = np.array(r["cooks_outlier"], dtype=bool)
    cutoff = f_dist.ppf(0.99, p, N - p)
    replaced, refitted, new_zero = np.zeros(G, bool), np.zeros(G, bool), np.zeros(G, bool)
    if refit_cooks:
        replaceable = n_or_more_replicates(X, min_replicates)
        if replaceable.any():
            replaced = np.array(r["cooks_replaced"], dtype=bool)
        if replaced.any():
            ridx = np.flatnonzero(replaced)
            counts = rf._h_counts
            sub = np.ascontiguousarray(counts[:, ridx])
            mu, hat = rf.gather_columns("mu", ridx), rf.gather_columns("hat", ridx)
            # Cook's distances of the replaced genes (dds.py:1022-1036) from the resident mu / hat and the device's robust dispersions
            a = np.asarray(r["robust_dispersions"])[ridx]
            V = mu + a[None, :] * mu**2
            ck = (sub - mu) ** 2 / V / p * (hat / (1 - hat) ** 2)
            above = ck > cutoff
            base = trimmed_mean_rows(sub / sf[:, None], 0.2)
            repl = (base[None, :] * sf[:, None]).astype(int)   # truncation towards zero, like DataFrame.astype(int)
            mask = replaceable[:, None] & above
            new = sub.copy()
            new[mask] = repl[mask]
            zero_now = (new == 0).all(axis=0)
            new_zero[ridx[zero_now]] = True
            refitted[ridx[~zero_now]] = True
            base_mean[new_zero] = 0.0
            lfc[new_zero] = 0.0
            if refitted.any():
                q = rf.refit_subset(np.ascontiguousarray(new[:, ~zero_now]), r["trend"], r["prior_var"], contrast, lfc_null, alt_hypothesis)
                base_mean[refitted], lfc[refitted] = q["normed_means"], q["lfc"]
                genewise[refitted], fitted[refitted], disp[refitted] = q["genewise"], q["fitted"], q["disp"]
                pv[refitted], stat[refitted], se[refitted] = q["pvalue"], q["stat"], q["se"]
                # which refitted genes still lose their p-value (dds.py:1066-1110 on `replace_cooks`: distances of the replaceable
                # samples are zeroed for refitted genes): samples in cells of 3..min_replicates-1 replicates can still flag them
                use_for_max = n_or_more_replicates(X, 3)
                keep = ~zero_now
                ck_r = ck[:, keep].copy()
                ck_r[replaceable] = 0.0
                out_r = (ck_r[use_for_max] > cutoff).any(axis=0)
                if out_r.any():
                    pos = ck[:, keep][:, out_r].argmax(0)
                    top = sub[:, keep][:, out_r][pos, np.arange(len(pos))]
                    out_r[out_r] = (sub[:, keep][:, out_r] > top).sum(0) < 3
                outlier[ridx[keep]] = out_r
            if new_zero.any():
                se[new_zero], stat[new_zero], pv[new_zero] = 0.0, 0.0, 1.0   # ds.py:355-360
    if cooks_filter:
        pv[outlier] = np.nan
    nz = np.ones(G, bool)
    if non_zero is not None:
        nz = np.asarray(non_zero, dtype=bool)
        assert int(nz.sum()) == G, "non_zero must select exactly the uploaded genes"

        def full(v, fill):
            out = np.full((len(nz),) + v.shape[1:], fill, dtype=v.dtype if v.dtype != bool else bool)
            out[nz] = v
            return out

        base_mean, lfc = full(base_mean, 0.0), full(lfc, np.nan)
        disp, genewise, fitted = full(disp, np.nan), full(genewise, np.nan), full(fitted, np.nan)
        pv, stat, se = full(pv, np.nan), full(stat, np.nan), full(se, np.nan)
        replaced, refitted, outlier = full(replaced, False), full(refitted, False), full(outlier, False)
    # adjust=False stops where deseq2() + run_wald_test() + the Cook's filter stop (no independent filtering / BH: `padj` is NaN)
    padj = _adjust(base_mean, pv, independent_filter, alpha, None) if adjust else np.full(len(pv), np.nan)
    return Deseq2Results(base_mean, lfc @ contrast / LN2, se / LN2, stat, pv, padj, lfc, disp, genewise, fitted, sf, nz, replaced, refitted,
                         outlier, None, None)


def _full(v, nz, G):
    out = np.full(G, np.nan)
    out[nz] = v
    return out


def _refit_subset(sub, X, sf, inference, fit: FitResult, min_mu, min_disp, max_disp, beta_tol):
    """``_refit_without_outliers`` (dds.py:1393-1458) on the replaced counts: genewise dispersions from scratch, the SAME trend
    function and prior as the main fit (neither is re-estimated), MAP dispersions, LFCs."""
    normed = sub / sf[:, None]
    mom = np.clip(np.minimum(inference.fit_rough_dispersions(normed, X), inference.fit_moments_dispersions(normed, sf)), min_disp, max_disp)
    if lin_mu_branch(X):
        mu_hat = inference.lin_reg_mu(sub, sf, X, min_mu)
    else:
        mu_hat = inference.irls(sub, sf, X, mom, min_mu, beta_tol)[1]
    mu_hat = np.ascontiguousarray(mu_hat)
    gw = np.clip(inference.alpha_mle(sub, X, mu_hat, mom, min_disp, max_disp)[0], min_disp, max_disp)
    means = normed.mean(0)
    co = fit.trend.coeffs
    fitted = (co[0] + co[1] / means) if fit.trend.kind == "parametric" else np.full(len(means), co[0])
    mp = inference.alpha_mle(sub, X, mu_hat, fitted, min_disp, max_disp, prior_disp_var=fit.prior_var, cr_reg=True, prior_reg=True)[0]
    disp = np.clip(mp, min_disp, max_disp)
    keep_gw = np.log(gw) > np.log(fitted) + 2 * np.sqrt(fit.squared_logres)
    disp[keep_gw] = gw[keep_gw]
    lfc = np.asarray(inference.irls(sub, sf, X, disp, min_mu, beta_tol)[0])
    return {"normed_means": means, "genewise": gw, "fitted": fitted, "disp": disp, "lfc": lfc}
