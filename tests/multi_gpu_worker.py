"""Worker of tests/test_multi_gpu_nccl.py: run under torchrun, one rank per GPU.  A mesh is split into
contiguous column slabs (sharding.slab); every rank runs the full BGC + DMS + MACROS step on its slab
and all-reduces the 64-double inventory over the ctx's NCCL communicator.  Rank 0 also computes the
WHOLE mesh alone on its GPU: the all-reduced vector must equal the single-GPU one up to the order of
the additions, and the counts exactly.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import parity  # noqa: E402


def step(host, ctx, bgc, dms, mac):
    ctx.inventory_reset()
    host.BGC_SourceSink(ctx, bgc, True, True)
    host.BGC_SurfaceFluxes(ctx, bgc)
    host.DMS_SourceSink(ctx, dms, True)
    host.DMS_SurfaceFluxes(ctx, dms)
    host.MACROS_SourceSink(ctx, mac, True)
    ctx.carbonate_join()


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = parity.pkg
    host = pkg.host
    parms = host.Parms()
    nL, mesh = 60, int(os.environ.get("BGC_TEST_MESH_COLUMNS", "6001"))
    c0, nC = pkg.sharding.slab(rank, world, mesh, even=bool(int(os.environ.get("BGC_TEST_EVEN", "1"))))
    dev = "cuda:%d" % local

    ctx = host.Context(nL, nC, device=local, parms=parms)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init_rank(world, rank, uid[0])
    ctx.inventory_enable(True)
    ctx.set_deferred_join(True)
    bgc, dms, mac = (host.DeviceBgcColumns(nL, nC, device=dev), host.DeviceDmsColumns(nL, nC, device=dev),
                     host.DeviceMacrosColumns(nL, nC, device=dev))
    pkg.synth_fill_device(parms, bgc, dms, mac, c0, ragged=True)
    step(host, ctx, bgc, dms, mac)               # cold
    step(host, ctx, bgc, dms, mac)               # warm: the step that is compared
    reduced = ctx.inventory_allreduce()
    local_vec = ctx.inventory_get()
    # the bench's pattern: the compute part replayed from a CUDA graph, the all-reduce issued after it
    ctx.graph_capture_begin()
    step(host, ctx, bgc, dms, mac)
    g = ctx.graph_capture_end()
    for _ in range(3):
        ctx.graph_launch(g)
        ctx.inventory_allreduce_begin()
    replayed = ctx.inventory_allreduce_end()

    t = torch.from_numpy(local_vec).to(dev)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    gathered = np.sum(np.stack([p.cpu().numpy() for p in parts]), axis=0)
    out = None
    if rank == 0:
        one = host.Context(nL, mesh, device=local, parms=parms)
        one.inventory_enable(True)
        b1, d1, m1 = (host.DeviceBgcColumns(nL, mesh, device=dev), host.DeviceDmsColumns(nL, mesh, device=dev),
                      host.DeviceMacrosColumns(nL, mesh, device=dev))
        pkg.synth_fill_device(parms, b1, d1, m1, 0, ragged=True)
        step(host, one, b1, d1, m1)
        step(host, one, b1, d1, m1)
        single = one.inventory_allreduce()      # no communicator: the identity
        one.close()

        def rel(a, b):
            # [0..51]: the 52 tendency sums.  [52..59] are the conservation residuals Jint_*: round-off
            # noise around zero by construction, compared in absolute terms below.
            m = np.abs(b[:52]) > 0
            return float(np.max(np.abs(a[:52] - b[:52])[m] / np.abs(b[:52])[m]))
        out = {"world": world, "mesh_columns": mesh, "slab": [c0, nC],
               "rel_vs_single_gpu": rel(reduced, single), "rel_vs_gathered": rel(reduced, gathered),
               "rel_graph_replay_vs_eager": rel(replayed, reduced) if np.any(reduced[:60]) else 0.0,
               "jint_abs_diff": float(np.max(np.abs(reduced[52:60] - single[52:60]))),
               "jint_scale": float(np.max(np.abs(single[:30]))),
               "counts": [float(reduced[60]), float(reduced[61])], "counts_single": [float(single[60]), float(single[61])],
               "nonzero_sums": int(np.count_nonzero(single[:52])),
               "zero_pattern_equal": bool(np.array_equal(reduced[:52] == 0, single[:52] == 0))}
    ctx.graph_destroy(g)
    ctx.synchronize()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
