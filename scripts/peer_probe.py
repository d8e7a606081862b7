"""Diagnostic (torchrun, >= 2 GPUs): can the ranks map each other's device memory (CUDA IPC windows of `sharding.PeerWindow`)?
Prints the error of every rank, the device topology flags, and -- when the window opens -- the time of a push of a few sizes."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pydeseq2_b200 import _lib
    from pydeseq2_b200.inference import B200Inference
    from pydeseq2_b200.sharding import NcclComm, PeerUnavailable

    inf = B200Inference(device=local)
    ctx = inf._ops.ctx
    uid = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(NcclComm.make_unique_id(ctx)), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = NcclComm(ctx, [1000] * world, rank, uid.cpu().numpy().tobytes())
    rec = {"rank": rank, "visible": os.environ.get("CUDA_VISIBLE_DEVICES"), "n_dev": torch.cuda.device_count(),
           "can_access_peer": [bool(torch.cuda.can_device_access_peer(local, d)) for d in range(torch.cuda.device_count()) if d != local]}
    try:
        n = 1 << 21
        win = comm.open_window(world * n * 8)
        rec["window"] = "ok"
        src = ctx.malloc(n * 8)
        h = ctx.pinned_empty((n,))
        h[:] = rank + 1
        ctx.h2d(src, h)
        ctx.sync()
        times = {}
        for count in (0, 20000, 340000, n):
            for _ in range(3):
                win.push([(src, 0)] if count else [], count)
            ctx.sync()
            dist.barrier()
            ctx.record(0)
            for _ in range(20):
                win.push([(src, 0)] if count else [], count)
            ctx.record(1)
            ctx.sync()
            times[count] = round(ctx.elapsed_ms(0, 1) / 20 * 1e3, 2)
        win.check()
        out = ctx.pinned_empty((world * n,))
        ctx.d2h(out, win.data)
        ctx.sync()
        rec["push_us"] = times
        rec["data_ok"] = bool(all(np.all(out[r * n:(r + 1) * n] == r + 1) for r in range(world)))
        win.close()
    except (PeerUnavailable, RuntimeError) as e:
        rec["window"] = f"{type(e).__name__}: {e}"
    print(json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
