"""World-size-2 test of the N > 1 host logic on CPU (gloo): contiguous column
slabs see bit-identical synthetic inputs to the unsharded mesh, per-slab results
concatenate to the unsharded result, and the inventory all-reduce is the sum of
the per-slab inventories.  The numerics run through the CPU oracle here (no GPU
in this container); the same slab/column0 plumbing drives bench.py on GPUs."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _inventory(cols):
    m = cols.active_mask()
    dz = np.where(m, cols.cell_thickness, 0.0)
    inv = np.zeros(64)
    inv[:30] = np.einsum("kcn,kc->n", cols.BGC_tendencies, dz)
    inv[60] = m.sum()
    inv[61] = (cols.number_of_active_levels[:cols.nColumns] > 0).sum()
    return inv


def _worker(rank, world, port, ntotal, nL, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = parity.pkg
    o = parity.oracle()
    po = o.Parms()
    first, n = pkg.sharding.slab(rank, world, ntotal)
    cols, _, _ = parity.make_bgc(nL, n, po, ragged=True, column0=first)
    o.BGC_SourceSink(po, cols, True)
    inv = torch.from_numpy(_inventory(cols))
    dist.all_reduce(inv, op=dist.ReduceOp.SUM)
    np.savez(os.path.join(tmp, "r%d.npz" % rank), tend=cols.BGC_tendencies, tr=cols.BGC_tracers,
             kmax=cols.number_of_active_levels, inv=inv.numpy(), first=first)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slabs_reproduce_the_unsharded_mesh(tmp_path):
    world, ntotal, nL = 2, 101, 20
    mp.spawn(_worker, args=(world, _free_port(), ntotal, nL, str(tmp_path)), nprocs=world, join=True)
    o = parity.oracle()
    po = o.Parms()
    full, _, _ = parity.make_bgc(nL, ntotal, po, ragged=True)
    o.BGC_SourceSink(po, full, True)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    assert [int(p["first"]) for p in parts] == [0, 51]
    assert np.array_equal(np.concatenate([p["tr"] for p in parts], axis=1), full.BGC_tracers)
    assert np.array_equal(np.concatenate([p["kmax"] for p in parts]), full.number_of_active_levels)
    assert np.array_equal(np.concatenate([p["tend"] for p in parts], axis=1), full.BGC_tendencies)
    want = _inventory(full)
    for p in parts:
        np.testing.assert_allclose(p["inv"][:30], want[:30], rtol=1e-12, atol=1e-18)
        assert p["inv"][60] == want[60] and p["inv"][61] == want[61]


def test_slab_plans():
    sh = parity.pkg.sharding
    for n in (0, 1, 7, 101, 235160):
        for w in (1, 2, 3, 8):
            for even in (False, True):
                plan = sh.slabs(w, n, even)
                assert sum(c for _, c in plan) == n
                pos = 0
                for first, c in plan:          # contiguous, in rank order, no overlap
                    assert first == pos or c == 0
                    pos += c
                if even:                       # every non-empty slab but the last one is even
                    nonempty = [c for _, c in plan if c]
                    assert all(c % 2 == 0 for c in nonempty[:-1])
    assert sh.slabs(8, 235160, even=True)[0] == (0, 29396)
    assert sh.slabs(8, 235160)[0] == (0, 29395)
