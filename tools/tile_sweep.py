#!/usr/bin/env python
"""Times the N = 1e6 LJ bench workload with the gather pair kernel and with the cell-tile pair kernel in every
(team, entries per lane, chunk size) variant.   python tools/tile_sweep.py [--steps 100] [--side 100]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import capi, sim, workloads as W  # noqa: E402
from parm_b200.capi import C  # noqa: E402


def run(w, steps, env):
    for k in [k for k in os.environ if k.startswith("PARM_B200_TILE")]:
        del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(20)
    capi.call("parm_sync", atoms._h)
    r0 = collec.stats()["rebuilds"]
    t0 = time.perf_counter()
    collec.timestep(steps)
    capi.call("parm_sync", atoms._h)
    dt = time.perf_counter() - t0
    nreb = collec.stats()["rebuilds"] - r0
    capi.call("parm_profile_enable", atoms._h, 1)
    collec.timestep(50)
    pms = (C.c_double * 4)()
    pcnt = (C.c_uint64 * 4)()
    capi.call("parm_profile_read", atoms._h, pms, pcnt)
    mean_n, mx = nl.stats()
    out = dict(env=env, tile=nl.tile_stats(), n_atoms=atoms.n, steps=steps, atom_steps_per_s=atoms.n * steps / dt,
               ms_per_step=dt / steps * 1e3, rebuilds=nreb, mean_full_neighbors=mean_n,
               k1_ms=pms[0] / max(pcnt[0], 1), force_ms=pms[1] / max(pcnt[1], 1), k3_ms=pms[2] / max(pcnt[2], 1),
               rebuild_ms_each=pms[3] / max(pcnt[3], 1), energy=collec.energy())
    print(json.dumps(out), flush=True)
    atoms.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--side", type=int, default=100)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    w = W.config3(a.side)
    run(w, a.steps, {"PARM_B200_TILE": 0})
    combos = [(4, 8, 128), (4, 4, 128), (2, 8, 128), (4, 8, 96), (4, 8, 160), (4, 8, 192)]
    if a.quick:
        combos = combos[:2]
    for team, v, ch in combos:
        run(w, a.steps, {"PARM_B200_TILE": 1, "PARM_B200_TILE_TEAM": team, "PARM_B200_TILE_V": v, "PARM_B200_TILE_CH": ch})


if __name__ == "__main__":
    main()
