"""CPU suite: edge cases and rarely taken branches of the DEVICE algorithms through the host emulator."""
import pytest

import edge_cases as ec
from emu.emu_ops import EmuOps
from pydeseq2_b200.inference import B200Inference


@pytest.fixture()
def backend():
    ops = EmuOps()
    return B200Inference(_ops=ops), ops


def test_zero_genes(backend):
    ec.check_zero_genes(backend[0])


def test_no_replicates(backend):
    ec.check_no_replicates_raises(backend[0])


def test_empty(backend):
    ec.check_empty_gene_set(backend[0])


def test_dtypes_layouts(backend):
    ec.check_input_dtypes_and_layouts(backend[0])


def test_wald_nan(backend):
    ec.check_nan_propagation_in_wald(backend[0])


def test_irls_optimizer_branch(backend):
    inf, ops = backend
    ec.check_irls_bounded_optimizer(inf, lambda on: setattr(ops, "force_optimizer", int(on)))


def test_alpha_grid_fallback(backend):
    inf, ops = backend
    ec.check_alpha_grid(inf, lambda on: setattr(ops, "force_grid", int(on)))


def test_size_factors(backend):
    ec.check_size_factors(backend[0])


def test_cooks(backend):
    ec.check_cooks(backend[0])


def test_shrink_grid_fallback(backend):
    inf, ops = backend
    ec.check_shrink_grid(inf, lambda on: setattr(ops, "force_shrink_grid", int(on)))


def test_shrink_arguments(backend):
    ec.check_shrink_arguments(backend[0])
