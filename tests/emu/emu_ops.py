"""ctypes front-end of the host emulator (tests/emu/pdq_emu.cpp) with the `_CudaOps` interface, so the
CPU suite can push the *device* algorithms through `B200Inference`'s marshalling.  Test infrastructure."""
import ctypes as C

import numpy as np

from . import build as _build

f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


class EmuOps:
    def __init__(self, force_optimizer=False, force_grid=False, force_shrink_grid=False, lanes=1):
        self.lib = C.CDLL(_build.build())
        self.lib.emu_lgamma.restype = C.c_double
        self.lib.emu_lgamma.argtypes = [C.c_double]
        self.lib.emu_digamma.restype = C.c_double
        self.lib.emu_digamma.argtypes = [C.c_double]
        self.force_optimizer = int(force_optimizer)
        self.force_grid = int(force_grid)
        self.force_shrink_grid = int(force_shrink_grid)
        self.lanes = int(lanes)  # 2..32: one std::thread per lane of a gene, the device's exchange patterns (pdq_emu.cpp)
        self.last_status = None

    def empty(self, shape):
        return np.empty(shape, dtype=np.float64)

    def __getattribute__(self, name):
        # the lane count is a global of the emulator library: select this object's before any of its calls
        if name in ("lin_reg_mu", "irls", "alpha_mle", "wald_test", "rough", "moments", "mom_from_counts", "cooks", "lfc_shrink"):
            lib = object.__getattribute__(self, "lib")
            assert lib.emu_set_lanes(object.__getattribute__(self, "lanes")) == 0
        return object.__getattribute__(self, name)

    def lin_reg_mu(self, counts, ld, N, G, sf, X, p, min_mu, mu):
        rc = self.lib.emu_lin_reg_mu(_p(counts, i64p), C.c_int64(ld), N, G, _p(sf, f64p), _p(X, f64p), p, C.c_double(min_mu),
                                     _p(mu, f64p))
        assert rc == 0

    def irls(self, counts, ld, N, G, sf, X, p, disp, min_mu, beta_tol, min_beta, max_beta, maxiter, beta, mu, hat, conv):
        status = np.zeros(G, dtype=np.int32)
        rc = self.lib.emu_irls(_p(counts, i64p), C.c_int64(ld), N, G, _p(sf, f64p), _p(X, f64p), p, _p(disp, f64p),
                               C.c_double(min_mu), C.c_double(beta_tol), C.c_double(min_beta), C.c_double(max_beta),
                               maxiter, _p(beta, f64p), _p(mu, f64p), _p(hat, f64p), _p(conv, f64p), _p(status, i32p),
                               self.force_optimizer)
        assert rc == 0
        self.last_status = status
        return int((status != 0).sum())

    def alpha_mle(self, counts, ld, N, G, X, p, mu, ld_mu, alpha_hat, min_disp, max_disp, prior_var, cr_reg, prior_reg,
                  alpha, conv):
        status = np.zeros(G, dtype=np.int32)
        rc = self.lib.emu_alpha_mle(_p(counts, i64p), C.c_int64(ld), N, G, _p(X, f64p), p, _p(mu, f64p), C.c_int64(ld_mu),
                                    _p(alpha_hat, f64p), C.c_double(min_disp), C.c_double(max_disp), C.c_double(prior_var),
                                    cr_reg, prior_reg, _p(alpha, f64p), _p(conv, f64p), _p(status, i32p), self.force_grid)
        assert rc == 0
        self.last_status = status

    def wald_test(self, X, N, p, disp, lfc, mu, ld_mu, G, ridge, contrast, lfc_null, alt, pv, stat, se):
        rc = self.lib.emu_wald_test(_p(X, f64p), N, p, _p(disp, f64p), _p(lfc, f64p), _p(mu, f64p), C.c_int64(ld_mu), G,
                                    _p(ridge, f64p), _p(contrast, f64p), C.c_double(lfc_null), alt, _p(pv, f64p),
                                    _p(stat, f64p), _p(se, f64p))
        assert rc == 0

    def rough(self, normed, ld, N, G, X, p, out):
        assert self.lib.emu_rough(_p(normed, f64p), C.c_int64(ld), N, G, _p(X, f64p), p, _p(out, f64p)) == 0

    def moments(self, normed, ld, N, G, sf, out, all_zero):
        assert self.lib.emu_moments(_p(normed, f64p), C.c_int64(ld), N, G, _p(sf, f64p), _p(out, f64p), _p(all_zero, f64p)) == 0

    def mom_from_counts(self, counts, sf, X, min_disp, max_disp):
        counts, sf, X = np.ascontiguousarray(counts), np.ascontiguousarray(sf), np.ascontiguousarray(X)
        N, G = counts.shape
        a = np.empty(G)
        m = np.empty(G)
        mu = np.empty((N, G))
        rc = self.lib.emu_mom_from_counts(_p(counts, i64p), C.c_int64(G), N, G, _p(sf, f64p), _p(X, f64p), X.shape[1],
                                          C.c_double(min_disp), C.c_double(max_disp), _p(a, f64p), _p(m, f64p), C.c_double(0.5),
                                          _p(mu, f64p))
        assert rc == 0
        return a, m, mu

    def trend_glm(self, cov, targets):
        out = np.zeros(16)
        rc = self.lib.emu_trend_fit(_p(cov, f64p), _p(targets, f64p), C.c_size_t(len(cov)), 0, C.c_double(-np.inf),
                                    C.c_double(np.inf), 0, C.c_double(0.0), C.c_double(0.0), 0, _p(out, f64p))
        assert rc == 0
        return out[:2].copy(), out[0] + out[1] * cov, bool(out[7])

    def trend_outer(self, means, gw, lo, hi, min_disp=1e-8, trigamma_c=0.0, with_prior=True):
        out = np.zeros(16)
        assert self.lib.emu_trend_fit(_p(means, f64p), _p(gw, f64p), C.c_size_t(len(gw)), 1, C.c_double(lo), C.c_double(hi), 1,
                                      C.c_double(min_disp), C.c_double(trigamma_c), int(with_prior), _p(out, f64p)) == 0
        return out

    def lfc_shrink(self, X, counts, ld, N, G, p, size, offset, prior_no_shrink_scale, prior_scale, shrink_index, lfcs, inv_hessians,
                   conv):
        status = np.zeros(G, dtype=np.int32)
        rc = self.lib.emu_lfc_shrink(_p(X, f64p), _p(counts, i64p), C.c_int64(ld), N, G, p, _p(size, f64p), _p(offset, f64p),
                                     C.c_double(prior_no_shrink_scale), C.c_double(prior_scale), shrink_index, _p(lfcs, f64p),
                                     _p(inv_hessians, f64p), _p(conv, f64p), _p(status, i32p), self.force_shrink_grid)
        assert rc == 0
        self.last_status = status
        return int((status != 0).sum())

    def size_factors(self, counts, ld, N, G, sf, logmeans=None):
        assert self.lib.emu_size_factors(_p(counts, i64p), C.c_int64(ld), N, G, _p(sf, f64p)) == 0
        if logmeans is not None:  # per-gene mean log count: the emulator entry returns the size factors only
            with np.errstate(divide="ignore"):
                logmeans[:] = np.log(np.asarray(counts)[:, :G]).mean(0)

    def cooks(self, counts, ld, N, G, sf, X, p, mu, hat, ld2, cutoff, cooks, disp, outlier, replaced):
        assert self.lib.emu_cooks(_p(counts, i64p), C.c_int64(ld), N, G, _p(sf, f64p), _p(X, f64p), p, _p(mu, f64p), _p(hat, f64p),
                                  C.c_int64(ld2), C.c_double(cutoff), _p(cooks, f64p) if cooks is not None else None,
                                  _p(disp, f64p), _p(outlier, f64p), _p(replaced, f64p)) == 0
