"""world_size-2 / world_size-4 gloo worker (CPU): host-side logic of the slab decomposition -- partition rule, unique-id
broadcast plumbing, merge of per-rank canonical pair lists. No CUDA."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402

from parm_b200 import sharded, workloads as W  # noqa: E402
from parity_util import cpu_system  # noqa: E402


def main():
    sharded.init_distributed("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world in (2, 4)
    # 1. the 128-byte id of rank 0 reaches every rank unchanged
    uid = bytes(range(128)) if rank == 0 else None
    got = sharded.broadcast_bytes(uid, 0)
    assert got == bytes(range(128))
    # 2. partition: every atom owned exactly once, inside its slab, also for unwrapped coordinates
    w = W.random_system(900, 3, W.KIND_LJATTRACTREPULSE, seed=21, ntypes=2)
    gid, x, v, m = sharded.partition_workload(w, rank, world)
    Lx = w["L"][0]
    wx = x[:, 0] - Lx * np.floor(x[:, 0] / Lx)
    assert np.all(wx >= rank * Lx / world - 1e-12) and np.all(wx < (rank + 1) * Lx / world + 1e-12)
    parts = [None] * world
    dist.all_gather_object(parts, gid)
    allg = np.concatenate(parts)
    assert len(allg) == 900 and len(np.unique(allg)) == 900
    # 3. merge of per-rank pair lists == the whole list (each pair is emitted by the owner of `first`)
    c = cpu_system("port", w, collection=False)
    a, b = c.pairs()
    mine = np.isin(a, gid)
    dist.all_gather_object(parts, (a[mine], b[mine]))
    ma, mb = sharded.merge_pairs(parts)
    assert np.array_equal(ma, a) and np.array_equal(mb, b)
    # 4. slab generator: disjoint global ids, right box, right density
    s = sharded.lj_lattice_slab(6, 5, 4, rank, world)
    dist.all_gather_object(parts, s["gid"])
    allg = np.concatenate(parts)
    assert len(np.unique(allg)) == s["n_global"] == 6 * world * 5 * 4
    assert abs(s["n_global"] / np.prod(s["L"]) - 1.1939) < 1e-12
    assert np.all(sharded.slab_of(s["x"][:, 0], s["L"][0], world) == rank)
    dist.barrier()
    if rank == 0:
        print("GLOO_WORKER_PASSED")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
