#!/usr/bin/env python
"""Step latency of small systems: BASELINE config 1 (N = 1000, LennardJonesCutPair) and LJ lattices of a few sizes.
   python tools/small_n_probe.py [--steps 2000]     (environment switches are read by the library at start-up)"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import capi, sim, workloads as W  # noqa: E402


def run(name, w, steps):
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(50)
    capi.call("parm_sync", atoms._h)
    r0 = collec.stats()["rebuilds"]
    t0 = time.perf_counter()
    collec.timestep(steps)
    capi.call("parm_sync", atoms._h)
    dt = time.perf_counter() - t0
    mean_n, mx = nl.stats()
    print(json.dumps(dict(case=name, env={k: v for k, v in os.environ.items() if k.startswith("PARM_B200_")}, n_atoms=atoms.n,
                          steps=steps, us_per_step=dt / steps * 1e6, atom_steps_per_s=atoms.n * steps / dt,
                          rebuilds=collec.stats()["rebuilds"] - r0, mean_full_neighbors=mean_n, tile=nl.tile_stats(),
                          energy=collec.energy())), flush=True)
    atoms.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--sides", type=int, nargs="*", default=[16, 24, 32])
    a = ap.parse_args()
    run("config 1: LJatoms-like N=1000", W.config1(), a.steps)
    for s in a.sides:
        run("LJ lattice %d^3" % s, W.lj_lattice((s, s, s), seed=3003), max(200, a.steps // 4))


if __name__ == "__main__":
    main()
