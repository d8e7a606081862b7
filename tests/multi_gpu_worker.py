"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), gene shards over NCCL through the C ABI.

Every rank fits its contiguous gene shard with `ResidentFit(comm=NcclComm)`; the pass ends with the exchange of the result
tables -- grouped NCCL all-gathers or the peer-memory push kernel (PDQ_MG_EXCHANGE = nccl | peer).  Checks (rank 0 holds the single-GPU fit of the WHOLE matrix as the reference):
  * trend coefficients / prior variance of the sharded pass == single-GPU pass (same kernel on the gathered vectors; the NaN
    pads of short shards only change the summation order),
  * every per-gene table gathered on EVERY rank == the single-GPU tables,
  * the same with the CUDA graph replay of the pass (ragged shards included).
Prints one JSON line on rank 0; exit code 0 = all checks passed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pydeseq2_b200 import _lib
    from pydeseq2_b200.inference import B200Inference
    from pydeseq2_b200.pipeline import ResidentFit, median_of_ratios
    from pydeseq2_b200.sharding import NcclComm, shard_bounds, shard_sizes
    from pydeseq2_b200.synth import make_counts

    N, G, kind = (int(os.environ.get("PDQ_MG_N", 60)), int(os.environ.get("PDQ_MG_G", 3001)), os.environ.get("PDQ_MG_DESIGN", "factorial"))
    counts, X, _ = make_counts(N, G, kind, seed=11)
    counts = np.ascontiguousarray(counts[:, ~(counts == 0).all(0)])
    G = counts.shape[1]
    sf = median_of_ratios(counts)[1]
    inf = B200Inference(device=local)
    ctx = inf._ops.ctx
    uid = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(NcclComm.make_unique_id(ctx)), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    sizes = shard_sizes(G, world)  # ragged whenever world does not divide G
    comm = NcclComm(ctx, sizes, rank, uid.cpu().numpy().tobytes())
    lo, hi = shard_bounds(G, world, rank)
    rf = ResidentFit(ctx, X, sf, comm=comm, with_cooks=True)
    rf.exchange = os.environ.get("PDQ_MG_EXCHANGE", "nccl")  # "peer": must map the peers' windows or fail
    rf.upload(counts[:, lo:hi])
    errs = {}
    ok = True
    ref = None
    if rank == 0:
        inf1 = B200Inference(device=local)
        rf1 = ResidentFit(inf1._ops.ctx, X, sf, with_cooks=True)
        rf1.upload(counts)
        ref = rf1.run()
    dist.barrier()
    for tag in ("eager", "eager2", "graph", "graph2"):  # third pass captures, fourth replays
        full = rf.gather_results(rf.run())
        # every rank must hold identical tables: compare a digest across ranks
        digest = np.concatenate([np.nan_to_num(np.ravel(full[k]), nan=-1.0) for k in ("dispersions", "lfc", "pvalue", "stat", "se", "genewise")])
        t = torch.from_numpy(np.array([digest.sum(), np.abs(digest).sum(), float(len(digest))])).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        same = all(bool(torch.equal(o, outs[0])) for o in outs)
        ok = ok and same
        if rank == 0:
            e = {"ranks_identical": same, "trend": float(np.max(np.abs(full["trend"].coeffs / ref["trend"].coeffs - 1))),
                 "prior_var": abs(full["prior_var"] / ref["prior_var"] - 1)}
            for k in ("genewise", "dispersions", "lfc", "pvalue", "stat", "se", "normed_means", "map"):
                a, b = np.asarray(full[k], float), np.asarray(ref[k], float)
                assert a.shape == b.shape, (k, a.shape, b.shape)
                with np.errstate(invalid="ignore", divide="ignore"):
                    e[k] = float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
            for k in ("cooks_outlier", "cooks_replaced"):
                e[k + "_equal"] = bool(np.array_equal(full[k], ref[k]))
            e["flags_equal"] = bool(np.array_equal(full["lfc_converged"], ref["lfc_converged"]))
            errs[tag] = e
            # the sharded trend sums the same numbers in another order: 1e-9 on the coefficients, which the MAP prior and the
            # IRLS stopping rule pass on attenuated; p-values of ~1e-300 amplify relative differences
            ok = ok and e["trend"] < 1e-9 and e["prior_var"] < 1e-9 and e["genewise"] < 1e-12 and e["dispersions"] < 1e-6 and \
                e["lfc"] < 1e-6 and e["stat"] < 1e-6 and e["se"] < 1e-6 and e["cooks_outlier_equal"] and e["flags_equal"]
    # a pass without the table exchange (peer flavour: ends with the bare barrier) returns the same shard results
    last = rf.run()
    rf.gather = False
    for _ in range(3):
        again = rf.run()
        ok = ok and all(np.array_equal(np.asarray(again[k]), np.asarray(last[k]), equal_nan=True) for k in ("dispersions", "lfc", "pvalue"))
    rf.gather = True
    used_peer = rf._win is not None
    ok = ok and (used_peer == (rf.exchange == "peer"))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "sizes": sizes, "N": N, "G": G, "design": kind, "exchange": "peer" if used_peer else "nccl", "ok": bool(flag.item()),
                          "errors": errs}), flush=True)
    rf.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
