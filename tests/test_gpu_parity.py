"""GPU suite (`-m gpu`): parity of the CUDA path, called through the C ABI, against (a) golden vectors from the
real reference and (b) the oracle on seeded inputs."""
import numpy as np
import pytest

from conftest import load_golden
from parity import SHRINK, assert_close, check_calls, check_shrink, check_tape

pytestmark = pytest.mark.gpu

CALLS = ["calls_two_level_n24", "calls_factorial_n30", "calls_continuous_n40", "calls_two_level_n200",
         "calls_large_counts_n12", "calls_five_columns_n36", "calls_intercept_n10", "calls_few_samples_n4"]
TAPES = ["tape_single_factor", "tape_multi_factor", "tape_continuous", "tape_wide", "tape_multi_factor_outliers"]


@pytest.fixture(scope="module")
def inf():
    from pydeseq2_b200.inference import B200Inference

    return B200Inference(device=0)


@pytest.mark.parametrize("name", CALLS)
def test_gpu_calls_vs_reference_golden(inf, name):
    # counts of 1e5-1e7: the objective itself is only resolved to ~1e-5 in float64 (cancellation in nb_nll), so the
    # reference's L-BFGS-B and the root search agree to the north-star 1e-4 there, not to 2e-5
    check_calls(inf, load_golden(name), **({"tol_alpha": 1e-4} if "large_counts" in name else {}))


@pytest.mark.parametrize("name", TAPES)
def test_gpu_replays_reference_tape(inf, name):
    check_tape(inf, load_golden(name), name)


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
def test_lane_group_width_does_not_change_results(lanes):
    """T lanes per gene only changes the summation tree: results must agree to rounding."""
    from pydeseq2_b200.inference import B200Inference

    g = load_golden("calls_continuous_n40")
    inf = B200Inference(device=0, lanes_per_gene=lanes)
    check_calls(inf, g)


def test_ragged_gene_counts_and_strided_input(inf):
    """G not a multiple of the genes per block, odd N (padded pack), row pitch > G."""
    g = load_golden("calls_factorial_n30")
    c, X, sf = g["counts"], g["X"], g["sf"]
    full = inf.irls(c, sf, X, g["mom"], 0.5, 1e-8)
    for G in (1, 3, 17, 33):
        sub = inf.irls(c[:, :G], sf, X, g["mom"][:G], 0.5, 1e-8)  # non-contiguous view -> ld = 40 > G
        np.testing.assert_array_equal(sub[0], full[0][:G])
        np.testing.assert_array_equal(sub[1], full[1][:, :G])
        np.testing.assert_array_equal(sub[2], full[2][:, :G])
        np.testing.assert_array_equal(sub[3], full[3][:G])
    # odd N
    N = 29
    b1, m1, h1, c1 = inf.irls(c[:N], sf[:N], X[:N], g["mom"], 0.5, 1e-8)
    from oracle import nbglm

    rb, rm, rh, rc = nbglm.OracleInference(n_cpus=1).irls(c[:N], sf[:N], X[:N], g["mom"], 0.5, 1e-8)
    assert_close(b1, rb, 1e-6, "odd-N beta", atol=1e-9)
    assert_close(h1, rh, 1e-6, "odd-N hat", atol=1e-12)


@pytest.mark.parametrize("name", SHRINK)
def test_gpu_lfc_shrink_vs_reference_golden(inf, name):
    check_shrink(inf, load_golden(name))
