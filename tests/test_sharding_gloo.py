"""CPU suite: the multi-rank host logic (gene shards + the trend all-gather) with world_size 2 over gloo.
Each rank fits its shard with the oracle backend; together they must reproduce the single-process fit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import nbglm
    from pydeseq2_b200.pipeline import fit_host, median_of_ratios
    from pydeseq2_b200.sharding import TorchDistComm, shard_bounds, shard_sizes
    from pydeseq2_b200.synth import make_counts

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    counts, X, _ = make_counts(24, 61, "two_level", seed=5)   # 61 genes: ragged shards (31 + 30)
    counts = counts[:, ~(counts == 0).all(0)]
    _, sf = median_of_ratios(counts)
    G = counts.shape[1]
    lo, hi = shard_bounds(G, world, rank)
    comm = TorchDistComm(shard_sizes(G, world))
    r = fit_host(counts[:, lo:hi], X, nbglm.OracleInference(n_cpus=1), size_factors=sf, comm=comm)
    # end-of-call exchange (SURVEY.md §8 e): one packed all-gather, every rank ends up with the full tables
    full = comm.allgather_table({"lfc": r.lfc, "disp": r.dispersions, "pv": r.pvalue})
    q.put((rank, lo, hi, r.lfc, r.dispersions, r.pvalue, r.trend.coeffs, r.prior_var, full))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gene_shards_reproduce_single_process_fit():
    import torch.multiprocessing as mp

    from oracle import nbglm
    from pydeseq2_b200.pipeline import fit_host, median_of_ratios
    from pydeseq2_b200.synth import make_counts

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    counts, X, _ = make_counts(24, 61, "two_level", seed=5)
    counts = counts[:, ~(counts == 0).all(0)]
    _, sf = median_of_ratios(counts)
    full = fit_host(counts, X, nbglm.OracleInference(n_cpus=1), size_factors=sf)
    lfc = np.concatenate([g[3] for g in got])
    disp = np.concatenate([g[4] for g in got])
    pv = np.concatenate([g[5] for g in got])
    np.testing.assert_allclose(lfc, full.lfc, rtol=1e-12)
    np.testing.assert_allclose(disp, full.dispersions, rtol=1e-12)
    np.testing.assert_allclose(pv, full.pvalue, rtol=1e-12)
    for g in got:
        np.testing.assert_allclose(g[6], full.trend.coeffs, rtol=1e-12)
        assert g[7] == pytest.approx(full.prior_var, rel=1e-12)
        # the gathered tables are identical on both ranks and equal the concatenation of the shards
        np.testing.assert_array_equal(g[8]["lfc"], lfc)
        np.testing.assert_array_equal(g[8]["disp"], disp)
        np.testing.assert_array_equal(g[8]["pv"], pv)


def test_shard_bounds_cover_and_partition():
    from pydeseq2_b200.sharding import shard_bounds, shard_sizes

    for G in (1, 7, 8, 61, 20000, 1_000_003):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(G, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == G
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert sum(shard_sizes(G, world)) == G


def _worker_workflow(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from pydeseq2_b200.pipeline import median_of_ratios
    from pydeseq2_b200.sharding import TorchDistComm, shard_bounds, shard_sizes
    from pydeseq2_b200.workflow import deseq2_results
    from test_workflow import OracleBackend

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_two_level_n24.npz"))
    counts, X, contrast = g["counts"][:, :151], g["design"], g["contrast"]     # 151 genes: ragged shards
    _, sf = median_of_ratios(counts)
    lo, hi = shard_bounds(counts.shape[1], world, rank)
    comm = TorchDistComm(shard_sizes(counts.shape[1], world))
    r = deseq2_results(counts[:, lo:hi], X, OracleBackend(n_cpus=1), contrast, size_factors=sf, comm=comm, shrink_coeff=1)
    q.put((rank, r.pvalue, r.padj, r.log2_fold_change, r.replaced, r.cooks_outlier, r.shrink_prior_scale))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_workflow_matches_single_process():
    """deseq2() + summary() + lfc_shrink over two gene shards: the refit stays local, the multiple-testing step and the apeGLM prior
    are exchanged -- the concatenated tables must equal the single-process ones."""
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from pydeseq2_b200.pipeline import median_of_ratios
    from pydeseq2_b200.workflow import deseq2_results
    from test_workflow import OracleBackend

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_workflow, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_two_level_n24.npz"))
    counts = g["counts"][:, :151]
    _, sf = median_of_ratios(counts)
    full = deseq2_results(counts, g["design"], OracleBackend(n_cpus=4), g["contrast"], size_factors=sf, shrink_coeff=1)
    assert full.replaced.sum() > 0
    np.testing.assert_allclose(np.concatenate([x[1] for x in got]), full.pvalue, rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(np.concatenate([x[2] for x in got]), full.padj, rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(np.concatenate([x[3] for x in got]), full.log2_fold_change, rtol=1e-8, atol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(np.concatenate([x[4] for x in got]), full.replaced)
    np.testing.assert_array_equal(np.concatenate([x[5] for x in got]), full.cooks_outlier)
    for x in got:
        assert x[6] == pytest.approx(full.shrink_prior_scale, rel=1e-10)
