#!/usr/bin/env python
"""Times the pair kernel (and the whole step) of BASELINE config 3 on the equilibrated state under environment switches
that libparm_b200 reads per launch / per rebuild.   python tools/pair_variants.py [--side 100] [--equil 600] 'A=1 B=2' ...
Each positional argument is one variant: space-separated NAME=VALUE pairs ('' = defaults)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import capi, sim, workloads as W  # noqa: E402
from parm_b200.capi import C  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=100)
    ap.add_argument("--equil", type=int, default=600)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("variants", nargs="*", default=[""])
    a = ap.parse_args()
    w = W.lj_lattice((a.side,) * 3, seed=3003)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    bench.equilibrate(collec, a.equil)
    f_ref = None
    import numpy as np
    for var in a.variants:
        env = dict(kv.split("=", 1) for kv in var.split()) if var.strip() else {}
        os.environ.update(env)
        bench.align_to_rebuild(collec)  # per-rebuild switches take effect here
        collec.timestep(6)
        capi.call("parm_sync", atoms._h)
        capi.call("parm_profile_enable", atoms._h, 1)
        collec.timestep(a.steps)
        pms = (C.c_double * 4)()
        pcnt = (C.c_uint64 * 4)()
        capi.call("parm_profile_read", atoms._h, pms, pcnt)
        capi.call("parm_profile_enable", atoms._h, 0)
        mean_n, mx = nl.stats()
        out = dict(env=env, n_atoms=atoms.n, steps=a.steps, mean_full_neighbors=mean_n, max_row=mx, tile=nl.tile_stats(),
                   k1_ms=pms[0] / max(pcnt[0], 1), force_ms=pms[1] / max(pcnt[1], 1), k3_ms=pms[2] / max(pcnt[2], 1),
                   rebuild_ms_each=pms[3] / max(pcnt[3], 1), rebuilds=int(pcnt[3]), T=float(collec.temp()))
        print(json.dumps(out), flush=True)
        for k in env:
            os.environ.pop(k, None)
    atoms.close()


if __name__ == "__main__":
    main()
