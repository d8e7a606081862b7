// pdq_shrink.cuh -- apeGLM log-fold-change shrinkage, one lane-group per gene.
//
// Reference: utils.py:990-1145 (`nbinomGLM`), :1148-1207 (`nbinomFn`), grid_search.py:224-318 (`grid_fit_shrink_beta`),
// fan-out default_inference.py:232-264, caller ds.py:363-443.
//
// Per gene the reference minimises   f(b) = prior(b) - sum_i [ y_i x_i'b - (y_i + s) logaddexp(x_i'b + o_i, log s) ]
// (s = 1/dispersion, o = log size factor; normal prior on every coefficient, Cauchy-type prior log1p((b_k/scale)^2) on the
// shrunk one), scaled by max(f(0), 1), with scipy's L-BFGS-B at ftol = gtol = 1e-8 and keeps the iterate the optimiser stops
// at.  That iterate sits 1e-3..1e-2 (relative) away from the true optimum, so agreeing with the reference to 1e-4 means
// walking the same path: `shrink_gene` is the unconstrained L-BFGS-B iteration itself (steepest descent / two-loop BFGS
// direction with m = 10, More-Thuente dcsrch/dcstep line search, the same update-skipping, restart and stopping rules;
// specification and scipy cross-check: oracle/lbfgsb_restated.py).  On the golden vectors it takes the same number of
// iterations and evaluations as scipy for every gene and ends within 1e-9 of the reference's coefficients.
//
// One objective+gradient evaluation is one sweep over the gene's samples by its T lanes; the optimiser logic between two
// sweeps is replicated on every lane of the gene (all lanes hold bit-identical sums after the butterfly), and the sweep
// loop runs in lock-step over the warp (`grp.any`) because the reductions shuffle across the full warp.
#pragma once

#include "pdq_gene.cuh"

namespace pdq {

struct ShrinkParams {
    double inv_var0;  // 1 / prior_no_shrink_scale^2
    double scale2;    // prior_scale^2
    int k;            // shrink_index
};

constexpr int kShrinkOk = 0;
constexpr int kShrinkNeedsGrid = 1;  // `not converged and num_vars == 2` (utils.py:1125-1141)
constexpr int kShrinkMem = 10;       // scipy maxcor
constexpr int kShrinkMaxLs = 20;     // scipy maxls
constexpr int kShrinkMaxIter = 15000;  // scipy maxiter
constexpr int kShrinkMaxEval = 4000;   // evaluation budget of this implementation (scipy: maxfun = 15000)

// unscaled f and (optionally) gradient, utils.py:1076-1089 and :1191-1207
template <int P, bool GRAD>
PDQ_HD void shrink_eval(const Group& grp, const DesignS& d, const ShrinkParams& prm, const int64_t* y, int64_t ld, double size,
                        double lsize, const double (&beta)[P], double& f, double (&g)[P]) {
    double nll = 0.0;
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) g[j] = 0.0;
    for (int n = grp.si; n < d.N; n += grp.T) {
        double x[P];
        load_x<P>(d, n, x);
        const double yv = (double)y[n * ld];
        double xb = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) xb = fma(x[j], beta[j], xb);
        const double e = xb + d.lsf[n * d.RS];
        const double dd = e - lsize;
        const double t = exp(-fabs(dd));                    // in (0, 1]
        const double lae = fmax(e, lsize) + log1p(t);       // logaddexp(e, log s)
        const double ys = yv + size;
        nll += yv * xb - ys * lae;
        if (GRAD) {
            const double q = ((dd >= 0.0) ? 1.0 : t) / (1.0 + t);  // 1 / (1 + s exp(-e)) without overflow
            const double gi = ys * q - yv;
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) g[j] = fma(gi, x[j], g[j]);
        }
    }
    nll = grp.sum(nll);
    if (GRAD) group_sum_vec<P>(grp, g);
    double prior = 0.0;
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) {
        if (j == prm.k) {
            prior += log1p(beta[j] * beta[j] / prm.scale2);
            if (GRAD) g[j] += 2.0 * beta[j] / (prm.scale2 + beta[j] * beta[j]);
        } else {
            prior += 0.5 * beta[j] * beta[j] * prm.inv_var0;
            if (GRAD) g[j] += beta[j] * prm.inv_var0;
        }
    }
    f = prior - nll;
}

// inverse of the reference's "Hessian" at beta (utils.py:1091-1108, 1143).  The reference adds the prior curvatures h_j to
// every entry of COLUMN j (a numpy broadcasting slip: `+ np.diag(h)` with h already a matrix), i.e. it inverts
// M = X'FX + 1 h' -- reproduced here through Sherman-Morrison on the Cholesky factor of A = X'FX:
//   M^-1 = A^-1 - (A^-1 1)(A^-1 h)' / (1 + h' A^-1 1).
template <int P>
PDQ_HD void shrink_inv_hessian(const Group& grp, const DesignS& d, const ShrinkParams& prm, const int64_t* y, int64_t ld, double size,
                               double lsize, const double (&beta)[P], double* ih_out, bool write) {
    Sym<P> A;
    sym_zero<P>(A);
    for (int n = grp.si; n < d.N; n += grp.T) {
        double x[P];
        load_x<P>(d, n, x);
        const double yv = (double)y[n * ld];
        double xb = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) xb = fma(x[j], beta[j], xb);
        const double dd = xb + d.lsf[n * d.RS] - lsize;
        const double t = exp(-fabs(dd));
        const double w = (yv + size) * (t / ((1.0 + t) * (1.0 + t)));  // (y+s) s e / (s+e)^2 = (y+s) q (1-q)
        sym_rank1<P>(A, w, x);
    }
    group_sum_sym<P>(grp, A);
    Sym<P> Ai;
    chol<P>(A);
    chol_inverse<P>(A, Ai);
    double h[P], u[P], v[P];
    const double bk = beta[prm.k], den = prm.scale2 + bk * bk;
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) h[j] = (j == prm.k) ? 2.0 * (prm.scale2 - bk * bk) / (den * den) : prm.inv_var0;
    double hu = 0.0;
PDQ_UNROLL_P
    for (int i = 0; i < P; ++i) {
        double su = 0.0, sv = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) {
            const double aij = Ai.a[j <= i ? tri(i, j) : tri(j, i)];
            su += aij;
            sv = fma(aij, h[j], sv);
        }
        u[i] = su;
        v[i] = sv;
    }
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) hu = fma(h[j], u[j], hu);
    const double inv = 1.0 / (1.0 + hu);
    if (write && grp.si == 0) {
PDQ_UNROLL_P
        for (int i = 0; i < P; ++i)
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) ih_out[i * P + j] = Ai.a[j <= i ? tri(i, j) : tri(j, i)] - u[i] * v[j] * inv;
    }
}

// ---- MINPACK-2 dcstep / dcsrch (line search of L-BFGS-B), state kept per gene ---------------------------------------
struct LineSearch {
    double finit, ginit, gtest, width, width1;
    double stx, fx, gx, sty, fy, gy, stmin, stmax;
    bool brackt;
    int stage;
};

PDQ_HD double dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double stp, double fp, double dp,
                     bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / fabs(dx));
    double stpf;
    if (fp > fx) {  // higher value: the minimum is bracketed
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
        const double stpc = stx + r * (stp - stx);
        const double stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (fabs(stpc - stx) < fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {  // derivatives of opposite sign: bracketed
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        double gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
        const double stpc = stp + r * (stx - stp);
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
        brackt = true;
    } else if (fabs(dp) < fabs(dx)) {  // lower value, same sign, the derivative shrinks
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = fmax(fabs(theta), fmax(fabs(dx), fabs(dp)));
        double gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
        double stpc;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        const double stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (fabs(stpc - stp) < fabs(stpq - stp)) ? stpc : stpq;
            if (stp > stx) stpf = fmin(stp + 0.66 * (sty - stp), stpf);
            else stpf = fmax(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (fabs(stpc - stp) > fabs(stpq - stp)) ? stpc : stpq;
            stpf = fmin(stpmax, stpf);
            stpf = fmax(stpmin, stpf);
        }
    } else {  // lower value, same sign, the derivative does not shrink
        if (brackt) {
            const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            const double s = fmax(fabs(theta), fmax(fabs(dy), fabs(dp)));
            double gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
            stpf = stp + r * (sty - stp);
        } else if (stp > stx) {
            stpf = stpmax;
        } else {
            stpf = stpmin;
        }
    }
    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    return stpf;
}

constexpr double kLsFtol = 1e-3, kLsGtol = 0.9, kLsXtol = 0.1, kLsStpMax = 1e10;

PDQ_HD void dcsrch_start(LineSearch& s, double stp, double f, double g) {
    s.brackt = false;
    s.stage = 1;
    s.finit = f;
    s.ginit = g;
    s.gtest = kLsFtol * g;
    s.width = kLsStpMax;
    s.width1 = s.width / 0.5;
    s.stx = 0.0; s.fx = f; s.gx = g;
    s.sty = 0.0; s.fy = f; s.gy = g;
    s.stmin = 0.0;
    s.stmax = stp + 4.0 * stp;
}

// returns true when another evaluation is wanted (task 'FG'; stp updated), false on CONVERGENCE / WARNING
PDQ_HD bool dcsrch_step(LineSearch& s, double& stp, double f, double g) {
    const double ftest = s.finit + stp * s.gtest;
    if (s.stage == 1 && f <= ftest && g >= 0.0) s.stage = 2;
    bool stop = false;
    if (s.brackt && (stp <= s.stmin || stp >= s.stmax)) stop = true;
    if (s.brackt && s.stmax - s.stmin <= kLsXtol * s.stmax) stop = true;
    if (stp == kLsStpMax && f <= ftest && g <= s.gtest) stop = true;
    if (stp == 0.0 && (f > ftest || g >= s.gtest)) stop = true;
    if (f <= ftest && fabs(g) <= kLsGtol * (-s.ginit)) stop = true;
    if (stop) return false;
    if (s.stage == 1 && f <= s.fx && f > ftest) {
        const double gt = s.gtest;
        const double fm = f - stp * gt, gm = g - gt;
        double fxm = s.fx - s.stx * gt, fym = s.fy - s.sty * gt, gxm = s.gx - gt, gym = s.gy - gt;
        stp = dcstep(s.stx, fxm, gxm, s.sty, fym, gym, stp, fm, gm, s.brackt, s.stmin, s.stmax);
        s.fx = fxm + s.stx * gt;
        s.fy = fym + s.sty * gt;
        s.gx = gxm + gt;
        s.gy = gym + gt;
    } else {
        stp = dcstep(s.stx, s.fx, s.gx, s.sty, s.fy, s.gy, stp, f, g, s.brackt, s.stmin, s.stmax);
    }
    if (s.brackt) {
        if (fabs(s.sty - s.stx) >= 0.66 * s.width1) stp = s.stx + 0.5 * (s.sty - s.stx);
        s.width1 = s.width;
        s.width = fabs(s.sty - s.stx);
        s.stmin = fmin(s.stx, s.sty);
        s.stmax = fmax(s.stx, s.sty);
    } else {
        s.stmin = stp + 1.1 * (stp - s.stx);
        s.stmax = stp + 4.0 * (stp - s.stx);
    }
    stp = fmin(fmax(stp, 0.0), kLsStpMax);
    if ((s.brackt && (stp <= s.stmin || stp >= s.stmax)) || (s.brackt && s.stmax - s.stmin <= kLsXtol * s.stmax)) stp = s.stx;
    return true;
}

// ---- the optimiser ---------------------------------------------------------------------------------------------------
template <int P>
struct Lbfgs {
    double S[kShrinkMem][P], Y[kShrinkMem][P], sy[kShrinkMem];
    int col, head;  // number of stored pairs, slot of the oldest
    double theta;
};

// z = x + (L-BFGS step) ; with an empty memory z = x - g / theta  (Cauchy point of L-BFGS-B without bounds)
template <int P>
PDQ_HD void lbfgs_point(const Lbfgs<P>& m, const double (&x)[P], const double (&g)[P], double (&z)[P]) {
    double q[P];
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) q[j] = -g[j];
    if (m.col == 0) {
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) z[j] = x[j] + (1.0 / m.theta) * q[j];
        return;
    }
    double al[kShrinkMem];
    for (int i = m.col - 1; i >= 0; --i) {  // newest -> oldest
        const int s = (m.head + i) % kShrinkMem;
        double sq = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) sq = fma(m.S[s][j], q[j], sq);
        const double a = sq / m.sy[s];
        al[i] = a;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) q[j] = q[j] - a * m.Y[s][j];
    }
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) q[j] = q[j] / m.theta;
    for (int i = 0; i < m.col; ++i) {  // oldest -> newest
        const int s = (m.head + i) % kShrinkMem;
        double yr = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) yr = fma(m.Y[s][j], q[j], yr);
        const double b = yr / m.sy[s];
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) q[j] = q[j] + m.S[s][j] * (al[i] - b);
    }
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) z[j] = x[j] + q[j];
}

template <int P>
PDQ_HD void shrink_gene(const Group& grp, const DesignS& d, const ShrinkParams& prm, const int64_t* y, int64_t ld, double size,
                        double* beta_out, double* ih_out, double* conv_out, int* status_out, bool valid, bool force_grid) {
    const double lsize = log(size);
    const double ftol = 1e-8, gtol = 1e-8;  // utils.py:1116-1119
    double x[P], g[P], f;
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) x[j] = 0.0;
    shrink_eval<P, false>(grp, d, prm, y, ld, size, lsize, x, f, g);
    const double cnst = (f < 1.0) ? 1.0 : f;  // np.maximum(scale_cnst, 1)
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) x[j] = (j & 1) ? -0.1 : 0.1;  // beta_init, utils.py:1051
    shrink_eval<P, true>(grp, d, prm, y, ld, size, lsize, x, f, g);
    f = f / cnst;
PDQ_UNROLL_P
    for (int j = 0; j < P; ++j) g[j] = g[j] / cnst;

    Lbfgs<P> mem;
    mem.col = 0;
    mem.head = 0;
    mem.theta = 1.0;
    LineSearch ls;
    double t[P], r[P], dir[P], z[P], xt[P];
    double fold = f, gdold = 0.0, stp = 1.0;
    int nit = 0, iback = 0;
    bool active = valid, ok = false;

    // prepares the next line search from (x, f, g); returns false on ABNORMAL_TERMINATION_IN_LNSRCH
    auto begin_iteration = [&]() -> bool {
        for (;;) {
            lbfgs_point<P>(mem, x, g, z);
            double dtd = 0.0, gd = 0.0;
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) {
                dir[j] = z[j] - x[j];
                dtd = fma(dir[j], dir[j], dtd);
            }
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) gd = fma(g[j], dir[j], gd);
            stp = (nit == 0) ? fmin(1.0 / sqrt(dtd), kLsStpMax) : 1.0;
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) { t[j] = x[j]; r[j] = g[j]; }
            fold = f;
            gdold = gd;
            iback = 0;
            if (!(gd < 0.0)) {  // not a descent direction: drop the memory and retry, or give up
                if (mem.col == 0) return false;
                mem.col = 0;
                mem.head = 0;
                mem.theta = 1.0;
                continue;
            }
            dcsrch_start(ls, stp, f, gd);
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) xt[j] = (stp == 1.0) ? z[j] : fma(stp, dir[j], t[j]);
            return true;
        }
    };

    {
        double gn = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) gn = fmax(gn, fabs(g[j]));
        if (active && gn <= gtol) { ok = true; active = false; }
        if (active && !begin_iteration()) active = false;
    }
    for (int ev = 0; ev < kShrinkMaxEval; ++ev) {
        if (!grp.any(active)) break;
        double ft, gt[P];
        shrink_eval<P, true>(grp, d, prm, y, ld, size, lsize, xt, ft, gt);
        if (!active) continue;
        f = ft / cnst;
        double gd = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) {
            g[j] = gt[j] / cnst;
            x[j] = xt[j];
        }
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) gd = fma(g[j], dir[j], gd);
        if (dcsrch_step(ls, stp, f, gd)) {  // another trial point
            if (++iback >= kShrinkMaxLs) {  // line search failed: restore, restart from steepest descent or give up
PDQ_UNROLL_P
                for (int j = 0; j < P; ++j) { x[j] = t[j]; g[j] = r[j]; }
                f = fold;
                if (mem.col == 0) { active = false; continue; }
                mem.col = 0;
                mem.head = 0;
                mem.theta = 1.0;
                if (!begin_iteration()) active = false;
                continue;
            }
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) xt[j] = (stp == 1.0) ? z[j] : fma(stp, dir[j], t[j]);
            continue;
        }
        // ---- new iterate
        ++nit;
        double gn = 0.0;
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) gn = fmax(gn, fabs(g[j]));
        if (gn <= gtol || (fold - f) <= ftol * fmax(fmax(fabs(fold), fabs(f)), 1.0)) {
            ok = true;
            active = false;
            continue;
        }
        if (nit >= kShrinkMaxIter) { active = false; continue; }
        // memory update (skipped when the curvature s'y is not safely positive)
        double rr = 0.0, yv[P];
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) {
            yv[j] = g[j] - r[j];
            rr = fma(yv[j], yv[j], rr);
        }
        const double dr = (stp == 1.0) ? (gd - gdold) : (gd - gdold) * stp;
        const double ddum = (stp == 1.0) ? -gdold : -gdold * stp;
        if (dr > 2.220446049250313e-16 * ddum) {
            int slot;
            if (mem.col < kShrinkMem) {
                slot = (mem.head + mem.col) % kShrinkMem;
                ++mem.col;
            } else {
                slot = mem.head;
                mem.head = (mem.head + 1) % kShrinkMem;
            }
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) {
                mem.S[slot][j] = (stp == 1.0) ? dir[j] : stp * dir[j];
                mem.Y[slot][j] = yv[j];
            }
            double sy = 0.0;
PDQ_UNROLL_P
            for (int j = 0; j < P; ++j) sy = fma(mem.Y[slot][j], mem.S[slot][j], sy);
            mem.sy[slot] = sy;
            mem.theta = rr / dr;
        }
        if (!begin_iteration()) active = false;
    }
    // `active` still set: evaluation budget exhausted -> not converged
    if (force_grid) ok = false;
    shrink_inv_hessian<P>(grp, d, prm, y, ld, size, lsize, x, ih_out, valid);
    if (valid && grp.si == 0) {
PDQ_UNROLL_P
        for (int j = 0; j < P; ++j) beta_out[j] = x[j];
        *conv_out = ok ? 1.0 : 0.0;
        *status_out = (!ok && P == 2) ? kShrinkNeedsGrid : kShrinkOk;
    }
}

// grid fallback for two-column designs (grid_search.py:224-318): 60 x 60 nodes on [-30, 30]^2, then 60 x 60 on one coarse cell
// either side of the best node; np.argmin's first minimum in row-major order.  The reference calls nbinomFn with its default
// shrink_index = 1 here whatever the caller asked for.
PDQ_HD void shrink_grid_gene(const Group& grp, const DesignS& d, ShrinkParams prm, const int64_t* y, int64_t ld, double size,
                             double* beta_out, double* ih_out, bool valid) {
    constexpr int K = 60;
    const double lsize = log(size);
    const ShrinkParams asked = prm;
    prm.k = 1;
    double b[2] = {0.0, 0.0}, g[2], f;
    shrink_eval<2, false>(grp, d, asked, y, ld, size, lsize, b, f, g);
    const double cnst = (f < 1.0) ? 1.0 : f;
    double lo0 = -30.0, hi0 = 30.0, lo1 = -30.0, hi1 = 30.0;
    double best0 = lo0, best1 = lo1;
    for (int pass = 0; pass < 2; ++pass) {
        double best = 0.0;
        int bi = 0, bj = 0;
        for (int i = 0; i < K; ++i) {
            b[0] = linspace_at(lo0, hi0, K, i);
            for (int j = 0; j < K; ++j) {
                b[1] = linspace_at(lo1, hi1, K, j);
                shrink_eval<2, false>(grp, d, prm, y, ld, size, lsize, b, f, g);
                f = f / cnst;
                if ((i == 0 && j == 0) || f < best) { best = f; bi = i; bj = j; }
            }
        }
        best0 = linspace_at(lo0, hi0, K, bi);
        best1 = linspace_at(lo1, hi1, K, bj);
        if (pass == 0) {
            const double delta = linspace_at(-30.0, 30.0, K, 1) - (-30.0);
            lo0 = best0 - delta; hi0 = best0 + delta;
            lo1 = best1 - delta; hi1 = best1 + delta;
        }
    }
    b[0] = best0;
    b[1] = best1;
    shrink_inv_hessian<2>(grp, d, asked, y, ld, size, lsize, b, ih_out, valid);
    if (valid && grp.si == 0) {
        beta_out[0] = b[0];
        beta_out[1] = b[1];
    }
}

}  // namespace pdq
