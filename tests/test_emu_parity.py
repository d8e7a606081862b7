"""CPU suite: the DEVICE algorithms (pdq_gene.cuh compiled for the host, one lane per gene) pushed through
B200Inference's marshalling, against golden vectors from the real reference."""
import numpy as np
import pytest

from conftest import load_golden
from parity import SHRINK, check_calls, check_shrink, check_tape
from pydeseq2_b200.inference import B200Inference
from emu.emu_ops import EmuOps

CALLS = ["calls_two_level_n24", "calls_factorial_n30", "calls_continuous_n40", "calls_two_level_n200",
         "calls_large_counts_n12", "calls_five_columns_n36", "calls_intercept_n10", "calls_few_samples_n4"]
TAPES = ["tape_single_factor", "tape_multi_factor", "tape_continuous", "tape_wide", "tape_multi_factor_outliers"]


@pytest.fixture(scope="module")
def inf():
    return B200Inference(_ops=EmuOps())


@pytest.mark.parametrize("name", CALLS)
def test_emu_calls(inf, name):
    # counts of 1e5-1e7: the objective itself is only resolved to ~1e-5 in float64 (cancellation in nb_nll), so the
    # reference's L-BFGS-B and the root search agree to the north-star 1e-4 there, not to 2e-5
    check_calls(inf, load_golden(name), **({"tol_alpha": 1e-4} if "large_counts" in name else {}))


@pytest.mark.parametrize("name", TAPES)
def test_emu_tape(inf, name):
    check_tape(inf, load_golden(name), name)


def test_special_functions():
    from scipy.special import gammaln, polygamma

    ops = EmuOps()
    xs = np.concatenate([np.geomspace(1e-4, 9.99, 400), np.geomspace(10, 1e9, 400), np.arange(1, 40) + 0.0])
    lg = np.array([ops.lib.emu_lgamma(float(x)) for x in xs])
    dg = np.array([ops.lib.emu_digamma(float(x)) for x in xs])
    np.testing.assert_allclose(lg, gammaln(xs), rtol=2e-14, atol=3e-14)
    np.testing.assert_allclose(dg, polygamma(0, xs), rtol=2e-14, atol=3e-14)


def test_fast_log_exp_accuracy():
    """The kernels' own log/exp (constant-bank polynomials, pdq_fast.cuh) against libm: <= 2 ulp."""
    import ctypes as C

    ops = EmuOps()
    for f in (ops.lib.emu_fast_log, ops.lib.emu_fast_exp):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2.0, 20000), [1.0, 2.0, 0.5, 1e-300, 1e300]])
    got = np.array([ops.lib.emu_fast_log(float(x)) for x in xs])
    want = np.log(xs)
    ulp = np.spacing(np.abs(want)) + 1e-320
    assert np.max(np.abs(got - want) / np.maximum(ulp, np.spacing(1e-16))) <= 2.0 or np.allclose(got, want, rtol=4e-16, atol=3e-16)
    np.testing.assert_allclose(got, want, rtol=5e-16, atol=3e-16)
    es = np.concatenate([rng.uniform(-690, 690, 20000), rng.uniform(-1, 1, 20000), [0.0, 1.0, -1.0, 709.0, -745.0]])
    got = np.array([ops.lib.emu_fast_exp(float(x)) for x in es])
    np.testing.assert_allclose(got, np.exp(es), rtol=5e-16)
    assert np.isnan(ops.lib.emu_fast_exp(float("nan"))) and np.isnan(ops.lib.emu_fast_log(float("nan")))
    assert ops.lib.emu_fast_log(0.0) == -np.inf and np.isnan(ops.lib.emu_fast_log(-1.0))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_device_trend_and_prior_match_host_glue(seed):
    """The on-device trend (outer loop included) + MAD prior (radix-select medians) against the host glue that
    reproduces dds.py:1199-1275 / 840-884 with the reference's scipy fit."""
    from scipy.special import polygamma

    from oracle import nbglm
    from pydeseq2_b200.pipeline import fit_prior_var, fit_trend

    class Ref:
        dispersion_trend_gamma_glm = staticmethod(nbglm.dispersion_trend_gamma_glm)

    ops = EmuOps()
    rng = np.random.default_rng(seed)
    G, N, p = [2001, 6000][seed % 2], 60, 3
    means = np.exp(rng.normal(4, 2, G) * np.log(2))
    gw = (4 / means + 0.1) * np.exp(rng.normal(0, [0.3, 0.8][seed // 2], G))
    gw[rng.integers(0, G, 40)] = 1e-8          # collapsed estimates are excluded from the prior (dds.py:868-875)
    gw[rng.integers(0, G, 10)] *= 100.0        # far above the curve: dropped by the outer loop
    out = ops.trend_outer(means, gw, 1e-8, float(N), 1e-8, float(polygamma(1, (N - p) / 2)))
    gwc = np.clip(gw, 1e-8, N)
    tr = fit_trend(Ref(), means, gwc, 1e-8)
    sq, pv = fit_prior_var(gwc, tr.fitted, N, p, 1e-8)
    assert out[2] == 0.0 and int(out[3]) == tr.n_iter
    np.testing.assert_allclose(out[:2], tr.coeffs, rtol=1e-4)   # L-BFGS-B's own slack
    np.testing.assert_allclose(out[8], sq, rtol=2e-4)
    np.testing.assert_allclose(out[9], pv, rtol=2e-4)
    assert int(out[10]) == int((gwc >= 1e-6).sum())
    # with the SAME coefficients the prior is exact: feed the device coefficients to the host formula
    fitted = out[0] + out[1] / means
    sq2, pv2 = fit_prior_var(gwc, fitted, N, p, 1e-8)
    np.testing.assert_allclose(out[8], sq2, rtol=1e-12)
    np.testing.assert_allclose(out[9], pv2, rtol=1e-12)


@pytest.mark.parametrize("name", ["calls_two_level_n24", "calls_two_level_n200", "calls_factorial_n30", "calls_few_samples_n4"])
def test_fused_mom_kernel(name):
    """k_mom_from_counts (resident path): min/clip of the two moment estimators, normalised means and -- same
    projection -- the lin_reg_mu start values, against the reference's separate calls."""
    g = load_golden(name)
    c, X, sf = g["counts"], g["X"], g["sf"]
    N = c.shape[0]
    a, m, mu = EmuOps().mom_from_counts(c, sf, X, 1e-8, float(max(10, N)))
    np.testing.assert_allclose(a, g["mom"], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(m, g["normed"].mean(0), rtol=1e-12)
    np.testing.assert_allclose(mu, g["lin_mu"], rtol=1e-9)


@pytest.mark.parametrize("name", SHRINK)
def test_emu_lfc_shrink(inf, name):
    check_shrink(inf, load_golden(name))


# ---- the same algorithms with T > 1 cooperating lanes per gene (one thread each; tests/emu/pdq_emu.cpp): the strided sample walks,
# ---- the split table builds and the replicated control flow of the kernels, with the device's own reduction order
@pytest.mark.parametrize("lanes,name", [(2, "calls_two_level_n24"), (8, "calls_two_level_n24"), (4, "calls_factorial_n30"),
                                        (8, "calls_continuous_n40"), (4, "calls_five_columns_n36"), (4, "calls_few_samples_n4"),
                                        (8, "calls_large_counts_n12"), (2, "calls_intercept_n10")])
def test_emu_calls_multi_lane(lanes, name):
    check_calls(B200Inference(_ops=EmuOps(lanes=lanes)), load_golden(name), **({"tol_alpha": 1e-4} if "large_counts" in name else {}))


@pytest.mark.parametrize("lanes,name", [(4, "shrink_two_level_n24"), (8, "shrink_five_columns_n36"), (2, "shrink_few_samples_n4"),
                                        (8, "shrinktape_continuous"), (16, "shrink_two_level_n200")])
def test_emu_lfc_shrink_multi_lane(lanes, name):
    check_shrink(B200Inference(_ops=EmuOps(lanes=lanes)), load_golden(name))


@pytest.mark.parametrize("lanes,name", [(4, "tape_multi_factor_outliers"), (8, "tape_wide")])
def test_emu_tape_multi_lane(lanes, name):
    check_tape(B200Inference(_ops=EmuOps(lanes=lanes)), load_golden(name), name)


def test_lane_count_does_not_change_results():
    """Only the summation tree differs between lane-group widths (cf. the GPU suite's test of the same name)."""
    g = load_golden("calls_two_level_n200")
    ref_full = B200Inference(_ops=EmuOps(lanes=1)).irls(g["counts"], g["sf"], g["X"], g["mom"], 0.5, 1e-8)
    for lanes in (2, 16, 32):
        got = B200Inference(_ops=EmuOps(lanes=lanes)).irls(g["counts"][:, :24], g["sf"], g["X"], g["mom"][:24], 0.5, 1e-8)
        ref = tuple(np.ascontiguousarray(r[..., :24]) if r.shape[-1] == g["counts"].shape[1] else r[:24] for r in ref_full)
        for a, b in zip(got, ref):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-13)
