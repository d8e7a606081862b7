"""CPU suite: the host glue of pydeseq2_b200.pipeline (size factors, trend loop, prior, outlier rule, Wald
input) reproduces what the REAL orchestrator (dds.py / ds.py) produced on the reference's shipped datasets
(tape_*.npz `final_*` fields), when driven with the oracle backend and with the emulated device backend."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import nbglm
from pydeseq2_b200.inference import B200Inference
from pydeseq2_b200.pipeline import fit_host
from emu.emu_ops import EmuOps

TAPES = ["tape_single_factor", "tape_multi_factor", "tape_continuous", "tape_wide"]


@pytest.mark.parametrize("name", TAPES)
def test_glue_matches_reference_orchestrator(name):
    t = load_golden(name)
    r = fit_host(t["counts"], t["design"], nbglm.OracleInference(n_cpus=1), contrast=t["contrast"])
    np.testing.assert_allclose(r.size_factors, t["final_size_factors"], rtol=1e-12)
    np.testing.assert_allclose(r.lfc, t["final_LFC"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(r.dispersions, t["final_dispersions"], rtol=1e-8)
    np.testing.assert_allclose(r.genewise, t["final_genewise"][r.non_zero], rtol=1e-8)
    np.testing.assert_allclose(r.pvalue, t["final_pvalues"], rtol=1e-7)
    np.testing.assert_allclose(r.stat, t["final_stat"], rtol=1e-8)
    np.testing.assert_allclose(r.se, t["final_se"], rtol=1e-8)


@pytest.mark.parametrize("name", TAPES)
def test_chained_device_algorithms_match_reference_and_R(name):
    """Chained pipeline with the (emulated) device numerics: north-star tolerance 1e-4 vs the reference, and the
    reference's own 2-4 % tolerance vs the stored R DESeq2 results (tests/test_pydeseq2.py:932-942)."""
    t = load_golden(name)
    r = fit_host(t["counts"], t["design"], B200Inference(_ops=EmuOps()), contrast=t["contrast"])
    np.testing.assert_allclose(r.lfc, t["final_LFC"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(r.dispersions, t["final_dispersions"], rtol=1e-4)
    np.testing.assert_allclose(r.stat, t["final_stat"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(r.pvalue, t["final_pvalues"], rtol=1e-3)
    c = t["contrast"]
    log2fc = (r.lfc @ c) / np.log(2)
    tol = 0.04 if name != "tape_single_factor" else 0.02
    np.testing.assert_allclose(log2fc, t["r_log2FoldChange"], rtol=tol)
    np.testing.assert_allclose(r.pvalue, t["r_pvalue"], rtol=tol)
