"""CPU suite: the DEVICE algorithms (pdq_gene.cuh compiled for the host, one lane per gene) pushed through
B200Inference's marshalling, against golden vectors from the real reference."""
import numpy as np
import pytest

from conftest import load_golden
from parity import check_calls, check_tape
from pydeseq2_b200.inference import B200Inference
from emu.emu_ops import EmuOps

CALLS = ["calls_two_level_n24", "calls_factorial_n30", "calls_continuous_n40", "calls_two_level_n200"]
TAPES = ["tape_single_factor", "tape_multi_factor", "tape_continuous", "tape_wide"]


@pytest.fixture(scope="module")
def inf():
    return B200Inference(_ops=EmuOps())


@pytest.mark.parametrize("name", CALLS)
def test_emu_calls(inf, name):
    check_calls(inf, load_golden(name))


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
