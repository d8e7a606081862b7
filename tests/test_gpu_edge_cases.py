"""GPU suite: the same edge cases / rarely taken branches on the CUDA path (C ABI)."""
import pytest

import edge_cases as ec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def inf():
    from pydeseq2_b200.inference import B200Inference

    return B200Inference(device=0)


def _force(inf, flag):
    def f(on):
        inf._ops.ctx.check(inf._ops.lib.pdq_set_debug_flags(inf._ops.ctx.h, flag if on else 0))

    return f


def test_zero_genes(inf):
    ec.check_zero_genes(inf)


def test_no_replicates(inf):
    ec.check_no_replicates_raises(inf)


def test_empty(inf):
    ec.check_empty_gene_set(inf)


def test_dtypes_layouts(inf):
    ec.check_input_dtypes_and_layouts(inf)


def test_wald_nan(inf):
    ec.check_nan_propagation_in_wald(inf)


def test_irls_optimizer_branch(inf):
    ec.check_irls_bounded_optimizer(inf, _force(inf, 1))


def test_alpha_grid_fallback(inf):
    ec.check_alpha_grid(inf, _force(inf, 2))


def test_design_too_large_is_reported(inf):
    import numpy as np

    from pydeseq2_b200._lib import B200Error

    X = np.ones((10, 9))
    with pytest.raises(B200Error, match="p <= 8"):
        inf.lin_reg_mu(np.ones((10, 3), dtype=np.int64), np.ones(10), X, 0.5)


def test_size_factors(inf):
    ec.check_size_factors(inf)


def test_cooks(inf):
    ec.check_cooks(inf)


def test_shrink_grid_fallback(inf):
    ec.check_shrink_grid(inf, _force(inf, 4))


def test_shrink_arguments(inf):
    ec.check_shrink_arguments(inf)


@pytest.mark.parametrize("p,N", [(1, 9), (2, 13), (3, 17), (4, 21), (5, 24), (6, 27), (7, 31), (8, 35)])
def test_every_design_width(inf, p, N):
    from oracle import nbglm

    ec.check_design_width(inf, nbglm.OracleInference(n_cpus=4), p, N)
