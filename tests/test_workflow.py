"""CPU suite: the orchestration around the plugin calls (`pydeseq2_b200/workflow.py`: outlier refit, Cook's and independent
filtering, BH) against the final tables of the REAL orchestrator -- first with the oracle as backend (isolates the host logic:
agreement to rounding), then with the device algorithms through the host emulator (north-star tolerance)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import nbglm
from parity import E2E, TAPES_E2E, check_e2e
from pydeseq2_b200 import workflow as wf


class OracleBackend(nbglm.OracleInference):
    """The oracle plus the two methods the workflow needs beyond the ABC."""

    def calculate_cooks(self, counts, size_factors, design_matrix, mu, hat, return_matrix=True):
        return nbglm.calculate_cooks(counts, counts / size_factors[:, None], design_matrix, mu, hat)


@pytest.fixture(scope="module")
def emu():
    from emu.emu_ops import EmuOps
    from pydeseq2_b200.inference import B200Inference

    return B200Inference(_ops=EmuOps())


@pytest.mark.parametrize("name", TAPES_E2E + E2E)
def test_workflow_host_logic_with_oracle_backend(name):
    check_e2e(OracleBackend(n_cpus=4), load_golden(name), 1e-8, name)


@pytest.mark.parametrize("name", TAPES_E2E + E2E)
def test_workflow_with_device_algorithms(emu, name):
    check_e2e(emu, load_golden(name), 1e-4, name, max_frac=0.005)


def test_lowess_and_bh_against_reference_formulas():
    from scipy.stats import false_discovery_control

    rng = np.random.default_rng(0)
    p = rng.uniform(size=200) ** 3
    np.testing.assert_allclose(wf.bh_adjust(p), false_discovery_control(p, method="bh"), rtol=1e-14)
    x = np.linspace(0.0, 0.95, 50)
    y = np.round(40 * np.exp(-((x - 0.3) ** 2) / 0.05) + rng.normal(0, 2, 50))
    est = wf.lowess(x, y, frac=1 / 5)
    # property checks: local lines reproduce a straight line exactly, and the smoother is translation-equivariant
    np.testing.assert_allclose(wf.lowess(x, 3 * x + 1, frac=1 / 5), 3 * x + 1, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(wf.lowess(x, y + 7, frac=1 / 5), est + 7, rtol=1e-9, atol=1e-9)


def test_replicates_and_trimmed_mean():
    X = np.array([[1, 0]] * 7 + [[1, 1]] * 3, dtype=float)
    np.testing.assert_array_equal(wf.n_or_more_replicates(X, 7), [True] * 7 + [False] * 3)
    np.testing.assert_array_equal(wf.n_or_more_replicates(X, 3), [True] * 10)
    x = np.arange(20.0).reshape(10, 2)
    np.testing.assert_allclose(wf.trimmed_mean_rows(x, 0.2), x[2:8].mean(0))


def test_shrunk_results_table(emu):
    g = load_golden("e2e_two_level_n24")
    r0 = wf.deseq2_results(g["counts"], g["design"], emu, g["contrast"])
    r1 = wf.deseq2_results(g["counts"], g["design"], emu, g["contrast"], shrink_coeff=1)
    np.testing.assert_array_equal(r0.pvalue, r1.pvalue)      # shrinkage leaves the tests untouched (ds.py:363-367)
    np.testing.assert_array_equal(r0.padj, r1.padj)
    assert 0 < r1.shrink_prior_scale <= 1
    nz = r1.non_zero
    a0, a1 = np.abs(r0.log2_fold_change[nz]), np.abs(r1.log2_fold_change[nz])
    assert (a1 <= a0 + 1e-6).mean() > 0.9 and np.median(a1) < 0.7 * np.median(a0)   # towards zero
    ref = wf.deseq2_results(g["counts"], g["design"], OracleBackend(n_cpus=4), g["contrast"], shrink_coeff=1)
    close = np.isclose(r1.log2_fold_change, ref.log2_fold_change, rtol=1e-4, atol=1e-8, equal_nan=True)
    assert close.mean() > 0.99
