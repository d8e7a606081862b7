"""CPU suite: the device algorithms (host emulator) against the pinned oracle on seeded RANDOM designs -- every template width
p = 1..16, mixed categorical / continuous columns, odd sample counts -- for each plugin method.  The golden fixtures cover
p in {1, 2, 3, 4, 5}; this sweep makes sure no width-specific code path (register Cholesky, packed indices, the L-BFGS memory of the
shrinkage kernel) is left untested."""
import pytest

import edge_cases as ec

from emu.emu_ops import EmuOps
from oracle import nbglm
from pydeseq2_b200.inference import B200Inference


@pytest.fixture(scope="module")
def backends():
    return B200Inference(_ops=EmuOps()), nbglm.OracleInference(n_cpus=4)


# p = 9..16: the wide-design path (same source, loops over the design columns not unrolled, matrices in local memory)
WIDTHS = [(1, 9), (2, 13), (3, 17), (4, 21), (5, 24), (6, 27), (7, 31), (8, 35), (9, 91), (12, 121), (16, 161)]


@pytest.mark.parametrize("p,N", WIDTHS)
def test_every_design_width(backends, p, N):
    ec.check_design_width(backends[0], backends[1], p, N)
