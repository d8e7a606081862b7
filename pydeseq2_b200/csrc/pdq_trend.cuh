// pdq_trend.cuh -- parametric dispersion trend  alpha ~ a0 + a1 / mean  as a gamma GLM with identity link
// (reference: default_inference.py:200-230 for one fit, dds.py:1199-1275 for the outer loop that drops
// genes far from the curve and refits until the coefficients stop moving).
//
// The reference minimises  L(c) = mean(t/m + log m),  m = c0 + c1 x,  with scipy L-BFGS-B from (1, 1) under
// c >= 1e-12.  L-BFGS-B stops within ~4e-6 (relative) of the minimiser (measured); this code
// converges to the minimiser itself: Fisher scoring (expected Hessian, always positive definite) while the
// observed Hessian is indefinite, Newton steps once it is positive definite, backtracking on L, projection
// onto the bounds.  The whole fit -- all iterations of all outer rounds -- runs inside ONE kernel launch of
// one cooperative block; per iteration the G-length sums are block-reduced (warp shuffles + shared memory).
// Compiled for the host emulator as well (one "thread").
#pragma once

#include "pdq_math.cuh"

namespace pdq {

struct TrendOut {
    double c0, c1;
    double status;   // 0 = ok, 1 = failed (not converged or a coefficient <= 1e-10) -> caller falls back to the mean trend
    double n_outer;  // outer rounds performed (dds.py:1236 loop)
    double n_used;   // genes left in the fit after outlier filtering
    double n_iter;   // inner iterations summed over rounds
    double loss;
    double last_converged;  // 1 when the most recent single fit converged (the reference's `res.success`)
};

// Reducer concept: tid(), nthreads(), sum(double) -> block-wide total visible to every thread, sync().
struct SerialReducer {
    PDQ_HD int tid() const { return 0; }
    PDQ_HD int nthreads() const { return 1; }
    PDQ_HD double sum(double v) { return v; }
    PDQ_HD void sync() {}
};

struct TrendSums {
    double L, g0, g1, h00, h01, h11, f00, f01, f11, n;
};

// x[i]: covariate (1/mean) or the mean itself when `x_is_mean`; t[i]: genewise dispersion (clipped to [lo, hi]).
template <class R>
PDQ_HD TrendSums trend_sums(R& red, const double* x, const double* t, const unsigned char* keep, size_t n, bool x_is_mean,
                            double lo, double hi, double c0, double c1) {
    double L = 0, g0 = 0, g1 = 0, h00 = 0, h01 = 0, h11 = 0, f00 = 0, f01 = 0, f11 = 0, cnt = 0;
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        if (!keep[i]) continue;
        const double xv = x_is_mean ? 1.0 / x[i] : x[i];
        double tv = t[i];
        tv = (tv < lo) ? lo : ((tv > hi) ? hi : tv);
        if (!(tv == tv)) continue;  // np.nanmean skips NaN targets
        const double m = fma(c1, xv, c0);
        const double im = 1.0 / m;
        const double r = tv * im;
        L += r + log(m);
        const double gi = -(r - 1.0) * im;          // d/dm (t/m + log m) = (1 - t/m)/m
        g0 += gi;
        g1 += gi * xv;
        const double hi2 = (2.0 * r - 1.0) * im * im;  // d2/dm2
        h00 += hi2;
        h01 += hi2 * xv;
        h11 += hi2 * xv * xv;
        const double fi = im * im;                  // expected (Fisher) curvature
        f00 += fi;
        f01 += fi * xv;
        f11 += fi * xv * xv;
        cnt += 1.0;
    }
    TrendSums s;
    s.L = red.sum(L); s.g0 = red.sum(g0); s.g1 = red.sum(g1);
    s.h00 = red.sum(h00); s.h01 = red.sum(h01); s.h11 = red.sum(h11);
    s.f00 = red.sum(f00); s.f01 = red.sum(f01); s.f11 = red.sum(f11);
    s.n = red.sum(cnt);
    return s;
}

template <class R>
PDQ_HD double trend_loss(R& red, const double* x, const double* t, const unsigned char* keep, size_t n, bool x_is_mean,
                         double lo, double hi, double c0, double c1) {
    double L = 0;
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        if (!keep[i]) continue;
        const double xv = x_is_mean ? 1.0 / x[i] : x[i];
        double tv = t[i];
        tv = (tv < lo) ? lo : ((tv > hi) ? hi : tv);
        if (!(tv == tv)) continue;
        const double m = fma(c1, xv, c0);
        L += tv / m + log(m);
    }
    return red.sum(L);
}

// one GLM fit from (1, 1); returns true when converged
template <class R>
PDQ_HD bool trend_fit_once(R& red, const double* x, const double* t, const unsigned char* keep, size_t n, bool x_is_mean,
                           double lo, double hi, double& c0, double& c1, double& loss, int& iters) {
    const double kLB = 1e-12;  // bounds=[(1e-12, inf)] (default_inference.py:224)
    c0 = 1.0;
    c1 = 1.0;
    bool ok = false;
    for (int it = 0; it < 200; ++it) {
        ++iters;
        const TrendSums s = trend_sums(red, x, t, keep, n, x_is_mean, lo, hi, c0, c1);
        loss = s.L / s.n;
        if (!(s.L == s.L) || s.n < 2.0) return false;
        // variables held at the lower bound with the gradient pushing outward are fixed
        const bool fix0 = (c0 <= kLB && s.g0 > 0.0), fix1 = (c1 <= kLB && s.g1 > 0.0);
        double d0 = 0.0, d1 = 0.0;
        {
            double a = s.h00, b = s.h01, d = s.h11;
            const bool pd = (a > 0.0) && (a * d - b * b > 1e-12 * a * d);
            if (!pd) { a = s.f00; b = s.f01; d = s.f11; }  // Fisher scoring: always positive definite
            if (fix0 && fix1) {
                ok = true;
                break;
            } else if (fix0) {
                d1 = -s.g1 / d;
            } else if (fix1) {
                d0 = -s.g0 / a;
            } else {
                const double det = a * d - b * b;
                d0 = -(d * s.g0 - b * s.g1) / det;
                d1 = -(a * s.g1 - b * s.g0) / det;
            }
        }
        // Close to the minimiser (tiny Newton decrement) the summed loss can no longer resolve progress in FP64, but
        // the Newton step itself still can: take it without a line search (quadratic convergence, Hessian PD here).
        double n0 = c0, n1 = c1, Ln = s.L;
        const double dec = -(s.g0 * d0 + s.g1 * d1);
        if (dec <= 1e-9 * fabs(s.L)) {
            n0 = fmax(c0 + d0, kLB);
            n1 = fmax(c1 + d1, kLB);
        } else {
            // backtracking on the loss, projecting onto c >= 1e-12
            double step = 1.0;
            bool acc = false;
            for (int ls = 0; ls < 40; ++ls) {
                n0 = fmax(fma(step, d0, c0), kLB);
                n1 = fmax(fma(step, d1, c1), kLB);
                Ln = trend_loss(red, x, t, keep, n, x_is_mean, lo, hi, n0, n1);
                if (Ln <= s.L + 1e-4 * (s.g0 * (n0 - c0) + s.g1 * (n1 - c1))) {
                    acc = true;
                    break;
                }
                step *= 0.5;
            }
            if (!acc) return false;
        }
        const double rel = fmax(fabs(n0 - c0) / fmax(fabs(n0), 1e-300), fabs(n1 - c1) / fmax(fabs(n1), 1e-300));
        c0 = n0;
        c1 = n1;
        loss = Ln / s.n;
        if (rel < 1e-12 || (dec <= 1e-9 * fabs(s.L) && rel < 1e-9 && it > 60)) {
            ok = true;
            break;
        }
    }
    return ok;
}

// Full outer loop of dds.py:1199-1275.  `keep` (n bytes, scratch) holds the genes still in the fit.
template <class R>
PDQ_HD void trend_fit_outer(R& red, const double* x, const double* t, unsigned char* keep, size_t n, bool x_is_mean,
                            double lo, double hi, bool outer, TrendOut* out) {
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        const double xv = x_is_mean ? 1.0 / x[i] : x[i];
        keep[i] = (xv == xv) && (fabs(xv) <= 1.7976931348623157e308);  // drop inf / NaN covariates (dds.py:1225-1232)
    }
    red.sync();
    double o0 = 0.1, o1 = 0.1, c0 = 1.0, c1 = 1.0, loss = 0.0;
    int rounds = 0, iters = 0;
    bool failed = false, last_conv = false;
    for (;;) {
        const double l0 = log(fabs(c0 / o0)), l1 = log(fabs(c1 / o1));
        if (!(c0 > 1e-10 && c1 > 1e-10) || !(l0 * l0 + l1 * l1 >= 1e-6)) break;  // dds.py:1236-1238
        o0 = c0;
        o1 = c1;
        const bool conv = trend_fit_once(red, x, t, keep, n, x_is_mean, lo, hi, c0, c1, loss, iters);
        ++rounds;
        last_conv = conv;
        if (!conv || c0 <= 1e-10 || c1 <= 1e-10) {  // dds.py:1243-1252 -> mean trend
            failed = true;
            break;
        }
        if (!outer) break;
        // drop genes far from the curve before refitting (dds.py:1255-1265); uses the UNclipped-by-us genewise values
        for (size_t i = red.tid(); i < n; i += red.nthreads()) {
            if (!keep[i]) continue;
            const double xv = x_is_mean ? 1.0 / x[i] : x[i];
            double tv = t[i];
            tv = (tv < lo) ? lo : ((tv > hi) ? hi : tv);
            const double ratio = tv / fma(c1, xv, c0);
            if (ratio < 1e-4 || ratio >= 15.0) keep[i] = 0;
        }
        red.sync();
    }
    double used = 0.0;
    for (size_t i = red.tid(); i < n; i += red.nthreads()) used += keep[i] ? 1.0 : 0.0;
    used = red.sum(used);
    if (red.tid() == 0) {
        out->c0 = c0;
        out->c1 = c1;
        out->status = failed ? 1.0 : 0.0;
        out->n_outer = (double)rounds;
        out->n_used = used;
        out->n_iter = (double)iters;
        out->loss = loss;
        out->last_converged = last_conv ? 1.0 : 0.0;
    }
}

}  // namespace pdq
