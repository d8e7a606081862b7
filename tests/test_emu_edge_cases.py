"""CPU suite: edge cases and rarely taken branches of the DEVICE algorithms through the host emulator."""
import pytest

import edge_cases as ec
from emu.emu_ops import EmuOps
from pydeseq2_b200.inference import B200Inference


@pytest.fixture(params=[1, 4], ids=["one-lane", "four-lanes"])
def backend(request):
    ops = EmuOps(lanes=request.param)
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


@pytest.mark.parametrize("name", ec.GRID_BETA)
def test_irls_grid_fallback(backend, name):
    inf, ops = backend
    ec.check_irls_grid_fallback(inf, lambda on: setattr(ops, "force_optimizer", 3 if on else 0), name)


def test_alpha_grid_fallback(backend):
    inf, ops = backend
    ec.check_alpha_grid(inf, lambda on: setattr(ops, "force_grid", int(on)))


def test_size_factors(backend):
    ec.check_size_factors(backend[0])


def test_cooks(backend):
    ec.check_cooks(backend[0])


def test_shrink_grid_fallback(backend):
    inf, ops = backend
    if ops.lanes > 1:
        pytest.skip("14 400 objective sweeps per gene: the grid has no lane-specific logic beyond the sums tested elsewhere")
    ec.check_shrink_grid(inf, lambda on: setattr(ops, "force_shrink_grid", int(on)))


def test_shrink_arguments(backend):
    ec.check_shrink_arguments(backend[0])


def test_call_trace():
    import numpy as np

    from conftest import load_golden
    from pydeseq2_b200.workflow import deseq2_results

    inf = B200Inference(_ops=EmuOps(), trace=True)
    g = load_golden("e2e_edge_few_samples_and_outlier")
    deseq2_results(g["counts"], g["design"], inf, g["contrast"], shrink_coeff=1)
    seen = [t["method"] for t in inf.trace]
    # the call order of deseq2() + summary() + lfc_shrink, refit of the replaced genes included
    assert seen[:4] == ["fit_rough_dispersions", "fit_moments_dispersions", "lin_reg_mu", "alpha_mle"]
    assert seen.count("alpha_mle") == 4 and seen.count("irls") == 2 and "calculate_cooks" in seen and seen[-1] == "lfc_shrink_nbinom_glm"
    assert all(t["ms"] >= 0 for t in inf.trace)
    assert B200Inference(_ops=EmuOps()).trace is None
