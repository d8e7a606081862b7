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

#include <string.h>

#include "pdq_fast.cuh"
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
    // dispersion prior (dds.py:840-884), filled by trend_prior(): squared scaled MAD of the log residuals of the genes
    // with genewise >= 100 * min_disp, the prior variance max(sq - trigamma((N-p)/2), 0.25), and the genes counted
    double squared_logres, prior_var, n_above;
    double pad[5];
};

// Reducer concept: tid(), nthreads(), sum(double) -> block-wide total visible to every thread, sync().
struct SerialReducer {
    PDQ_HD int tid() const { return 0; }
    PDQ_HD int nthreads() const { return 1; }
    PDQ_HD double sum(double v) { return v; }
    PDQ_HD void sum_many(double*, int) {}  // in-place totals of k values, visible to every thread
    PDQ_HD void sync() {}
    // thread-private in the emulator; block-shared scratch on the device
    PDQ_HD int local_tid() const { return 0; }
    PDQ_HD int local_nthreads() const { return 1; }
    PDQ_HD void local_sync() {}
    PDQ_HD double max_one(double v) { return v; }
    PDQ_HD void hist_begin() {}
    PDQ_HD unsigned* hist_merge(unsigned* hist) { return hist; }
    PDQ_HD void find_bin(const unsigned* hist, unsigned*, size_t k, int& d, size_t& cum) {
        cum = 0;
        for (d = 0; d < 256; ++d) {
            const size_t c = hist[d];
            if (cum + c > k) break;
            cum += c;
        }
    }
};

struct TrendSums {
    double L, g0, g1, h00, h01, h11, f00, f01, f11, n;
};

// xs[i]: covariate; ts[i]: target, NaN when the gene is not (or no longer) part of the fit.  Both are prepared once per
// launch by trend_prepare (clip, 1/mean, validity), so that a pass is: two loads, one reciprocal, one log, ten sums.
template <class R>
PDQ_HD TrendSums trend_sums(R& red, const double* xs, const double* ts, size_t n, double c0, double c1) {
    double L = 0, g0 = 0, g1 = 0, h00 = 0, h01 = 0, h11 = 0, f00 = 0, f01 = 0, f11 = 0, cnt = 0;
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        const double tv = ts[i];
        if (!(tv == tv)) continue;  // dropped gene / np.nanmean skips NaN targets
        const double xv = xs[i];
        const double m = fma(c1, xv, c0);
        const double im = fast_rcp(m);
        const double r = tv * im;
        L += r + fast_log(m);
        const double gi = -(r - 1.0) * im;             // d/dm (t/m + log m) = (1 - t/m)/m
        g0 += gi;
        g1 += gi * xv;
        const double hi2 = (2.0 * r - 1.0) * im * im;  // d2/dm2
        h00 += hi2;
        h01 += hi2 * xv;
        h11 += hi2 * xv * xv;
        const double fi = im * im;                     // expected (Fisher) curvature
        f00 += fi;
        f01 += fi * xv;
        f11 += fi * xv * xv;
        cnt += 1.0;
    }
    double v[10] = {L, g0, g1, h00, h01, h11, f00, f01, f11, cnt};
    red.sum_many(v, 10);  // one packed reduction round for all ten sums
    TrendSums s;
    s.L = v[0]; s.g0 = v[1]; s.g1 = v[2];
    s.h00 = v[3]; s.h01 = v[4]; s.h11 = v[5];
    s.f00 = v[6]; s.f01 = v[7]; s.f11 = v[8];
    s.n = v[9];
    return s;
}

// One GLM fit; returns true when converged.  Starts from the incoming (c0, c1): (1, 1) for the first round like the
// reference (default_inference.py:221); later rounds of the outer loop warm-start from the previous optimum -- the
// minimiser does not depend on the start, only the iteration count does.  Every trial point costs ONE pass: the sums at
// the candidate serve both the Armijo test and, once accepted, the next Newton direction.
template <class R>
PDQ_HD bool trend_fit_once(R& red, const double* xs, const double* ts, size_t n, double& c0, double& c1, double& loss,
                           int& iters) {
    const double kLB = 1e-12;  // bounds=[(1e-12, inf)] (default_inference.py:224)
    TrendSums s = trend_sums(red, xs, ts, n, c0, c1);
    ++iters;
    for (int it = 0; it < 200; ++it) {
        loss = s.L / s.n;
        if (!(s.L == s.L) || s.n < 2.0) return false;
        // variables held at the lower bound with the gradient pushing outward are fixed
        const bool fix0 = (c0 <= kLB && s.g0 > 0.0), fix1 = (c1 <= kLB && s.g1 > 0.0);
        double d0 = 0.0, d1 = 0.0;
        {
            double a = s.h00, b = s.h01, d = s.h11;
            const bool pd = (a > 0.0) && (a * d - b * b > 1e-12 * a * d);
            if (!pd) { a = s.f00; b = s.f01; d = s.f11; }  // Fisher scoring: always positive definite
            if (fix0 && fix1) return true;
            if (fix0) {
                d1 = -s.g1 / d;
            } else if (fix1) {
                d0 = -s.g0 / a;
            } else {
                const double det = a * d - b * b;
                d0 = -(d * s.g0 - b * s.g1) / det;
                d1 = -(a * s.g1 - b * s.g0) / det;
            }
        }
        const double dec = -(s.g0 * d0 + s.g1 * d1);           // Newton decrement (>= 0 for a descent direction)
        const bool endgame = dec <= 1e-9 * fabs(s.L);           // the summed loss no longer resolves progress in FP64
        double step = 1.0, n0 = c0, n1 = c1;
        TrendSums sn = s;
        bool acc = false;
        for (int ls = 0; ls < 40; ++ls) {
            n0 = fmax(fma(step, d0, c0), kLB);
            n1 = fmax(fma(step, d1, c1), kLB);
            sn = trend_sums(red, xs, ts, n, n0, n1);
            ++iters;
            if (endgame || sn.L <= s.L + 1e-4 * (s.g0 * (n0 - c0) + s.g1 * (n1 - c1))) {
                acc = true;
                break;
            }
            step *= 0.5;
        }
        if (!acc || !(sn.L == sn.L)) return false;
        const double rel = fmax(fabs(n0 - c0) / fmax(fabs(n0), 1e-300), fabs(n1 - c1) / fmax(fabs(n1), 1e-300));
        c0 = n0;
        c1 = n1;
        s = sn;
        loss = s.L / s.n;
        if (rel < 1e-12 || (endgame && rel < 1e-9 && it > 60)) return true;
    }
    return false;
}

// validity, clipping and the covariate are resolved once: xs / ts are n doubles of scratch each
template <class R>
PDQ_HD void trend_prepare(R& red, const double* x, const double* t, size_t n, bool x_is_mean, double lo, double hi,
                          double* xs, double* ts) {
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        const double xv = x_is_mean ? 1.0 / x[i] : x[i];
        double tv = t[i];
        tv = (tv < lo) ? lo : ((tv > hi) ? hi : tv);
        const bool ok = (xv == xv) && (fabs(xv) <= 1.7976931348623157e308);  // drop inf / NaN covariates (dds.py:1225-1232)
        xs[i] = xv;
        ts[i] = ok ? tv : (0.0 / 0.0);
    }
    red.sync();
}

// Full outer loop of dds.py:1199-1275 on the prepared vectors; ts[i] is set to NaN when gene i is dropped.
template <class R>
PDQ_HD TrendOut trend_fit_outer(R& red, const double* xs, double* ts, size_t n, bool outer) {
    double o0 = 0.1, o1 = 0.1, c0 = 1.0, c1 = 1.0, loss = 0.0;
    int rounds = 0, iters = 0;
    bool failed = false, last_conv = false;
    for (;;) {
        const double l0 = log(fabs(c0 / o0)), l1 = log(fabs(c1 / o1));
        if (!(c0 > 1e-10 && c1 > 1e-10) || !(l0 * l0 + l1 * l1 >= 1e-6)) break;  // dds.py:1236-1238
        o0 = c0;
        o1 = c1;
        const bool conv = trend_fit_once(red, xs, ts, n, c0, c1, loss, iters);
        ++rounds;
        last_conv = conv;
        if (!conv || c0 <= 1e-10 || c1 <= 1e-10) {  // dds.py:1243-1252 -> mean trend
            failed = true;
            break;
        }
        if (!outer) break;
        // drop genes far from the curve before refitting (dds.py:1255-1265)
        for (size_t i = red.tid(); i < n; i += red.nthreads()) {
            const double tv = ts[i];
            if (!(tv == tv)) continue;
            const double ratio = tv / fma(c1, xs[i], c0);
            if (ratio < 1e-4 || ratio >= 15.0) ts[i] = 0.0 / 0.0;
        }
        red.sync();
    }
    double used = 0.0;
    for (size_t i = red.tid(); i < n; i += red.nthreads()) used += (ts[i] == ts[i]) ? 1.0 : 0.0;
    red.sum_many(&used, 1);
    TrendOut o;  // identical in every thread (all decisions were taken on block/cluster-wide sums)
    o.c0 = c0;
    o.c1 = c1;
    o.status = failed ? 1.0 : 0.0;
    o.n_outer = (double)rounds;
    o.n_used = used;
    o.n_iter = (double)iters;
    o.loss = loss;
    o.last_converged = last_conv ? 1.0 : 0.0;
    o.squared_logres = o.prior_var = o.n_above = 0.0 / 0.0;
    for (int i = 0; i < 5; ++i) o.pad[i] = 0.0;
    return o;
}

// ---------------------------------------------------------------------------------------------------------
// Dispersion prior (dds.py:840-884): residuals r = log(genewise) - log(fitted) over the genes with
// genewise >= 100 * min_disp, squared scaled MAD (utils.py:1210-1227), prior_var = max(sq - trigamma((N-p)/2), 0.25).
// The two medians are exact order statistics found by an MSB-first radix select (8-bit digits, 256-bin histogram
// in block-shared memory) over an order-preserving integer image of the doubles; every block of the cluster runs
// the selection redundantly on the full vector (no cross-block traffic, identical results everywhere).
// ---------------------------------------------------------------------------------------------------------
PDQ_HD uint64_t f64_key(double v) {
    uint64_t u;
#if defined(__CUDA_ARCH__)
    u = (uint64_t)__double_as_longlong(v);
#else
    memcpy(&u, &v, 8);
#endif
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
PDQ_HD double f64_from_key(uint64_t k) {
    const uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}

// k-th smallest (0-based) of { |res[i] - center| or res[i] } over the non-NaN entries.  The blocks of the cluster
// histogram disjoint slices of the vector and merge their 256-bin histograms (reducer hist_* hooks); `hist` holds
// 256 local bins + 256 merged bins + 2 result slots.
template <class R>
PDQ_HD double select_kth(R& red, const double* res, size_t n, bool absdev, double center, unsigned* hist, size_t k) {
    uint64_t prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = red.local_tid(); b < 512; b += red.local_nthreads()) hist[b] = 0;
        red.hist_begin();
        for (size_t i = red.tid(); i < n; i += red.nthreads()) {
            const double r = res[i];
            if (!(r == r)) continue;
            const uint64_t key = f64_key(absdev ? fabs(r - center) : r);
            if ((key & mask) == prefix) {
#if defined(__CUDA_ARCH__)
                atomicAdd(&hist[(key >> shift) & 0xff], 1u);
#else
                ++hist[(key >> shift) & 0xff];
#endif
            }
        }
        unsigned* tot = red.hist_merge(hist);  // totals over the cluster, identical in every block
        size_t cum = 0;
        int d = 0;
        red.find_bin(tot, hist + 512, k, d, cum);  // first bin whose cumulative count exceeds k, and the count before it
        k -= cum;
        prefix |= (uint64_t)d << shift;
        mask |= (uint64_t)0xff << shift;
        red.local_sync();
    }
    return f64_from_key(prefix);
}

template <class R>
PDQ_HD double median_of(R& red, const double* res, size_t n, size_t cnt, bool absdev, double center, unsigned* hist) {
    if (cnt == 0) return 0.0 / 0.0;  // np.median([]) = nan
    const size_t k = cnt / 2;
    const double hi = select_kth(red, res, n, absdev, center, hist, k);  // upper median
    if (cnt & 1) return hi;
    // even count: the lower median is the largest value below `hi`, unless `hi` is duplicated down to rank k-1
    double less = 0.0, mx = -1.7976931348623157e308;
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        const double r = res[i];
        if (!(r == r)) continue;
        const double v = absdev ? fabs(r - center) : r;
        if (v < hi) {
            less += 1.0;
            mx = v > mx ? v : mx;
        }
    }
    red.sum_many(&less, 1);
    mx = red.max_one(mx);
    const double lo = ((size_t)less >= k) ? mx : hi;
    return 0.5 * (lo + hi);
}

// `res` : n doubles of scratch; `hist`: 514 unsigned in block-shared memory (see select_kth); `trigamma_c` = polygamma(1, (N-p)/2)
template <class R>
PDQ_HD void trend_prior(R& red, const double* means, const double* t, size_t n, double lo, double hi, double min_disp,
                        double trigamma_c, double* res, unsigned* hist, TrendOut& out) {
    const double c0 = out.c0, c1 = out.c1;
    double cnt = 0.0;
    // every thread fills (and later only ever reads) its own stride of the scratch vector: no barrier needed
    for (size_t i = red.tid(); i < n; i += red.nthreads()) {
        double tv = t[i];
        tv = (tv < lo) ? lo : ((tv > hi) ? hi : tv);
        const double fit = c0 + c1 / means[i];
        const bool use = (tv >= 100.0 * min_disp) && (means[i] == means[i]);
        res[i] = use ? (log(tv) - log(fit)) : (0.0 / 0.0);
        cnt += use ? 1.0 : 0.0;
    }
    red.sum_many(&cnt, 1);
    const size_t m = (size_t)cnt;
    const double med = median_of(red, res, n, m, false, 0.0, hist);
    const double mad = median_of(red, res, n, m, true, med, hist) * 1.4826022185056018;  // / (sqrt(2) erfinv(1/2))
    const double sq = mad * mad;
    out.squared_logres = sq;
    const double pv = sq - trigamma_c;
    out.prior_var = (pv > 0.25 || pv != pv) ? pv : 0.25;  // np.maximum propagates NaN
    out.n_above = cnt;
}

}  // namespace pdq
