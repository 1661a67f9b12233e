#!/usr/bin/env python
"""Times every single-GPU BASELINE.json configuration (not only the bench line) for a few hundred steps.
   python tools/config_timings.py [--steps 200]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import capi, sim, workloads as W  # noqa: E402
from parm_b200.capi import C  # noqa: E402


def run(name, w, steps):
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(20)
    capi.call("parm_sync", atoms._h)
    r0 = collec.stats()["rebuilds"]
    t0 = time.perf_counter()
    collec.timestep(steps)
    capi.call("parm_sync", atoms._h)
    dt = time.perf_counter() - t0
    capi.call("parm_profile_enable", atoms._h, 1)
    collec.timestep(50)
    pms = (C.c_double * 4)()
    pcnt = (C.c_uint64 * 4)()
    capi.call("parm_profile_read", atoms._h, pms, pcnt)
    mean_n, mx = nl.stats()
    out = dict(config=name, n_atoms=atoms.n, ndim=w["ndim"], steps=steps, atom_steps_per_s=atoms.n * steps / dt,
               ms_per_step=dt / steps * 1e3, rebuilds=collec.stats()["rebuilds"] - r0 - int(pcnt[3]), mean_full_neighbors=mean_n,
               k1_ms=pms[0] / max(pcnt[0], 1), force_ms=pms[1] / max(pcnt[1], 1), k3_ms=pms[2] / max(pcnt[2], 1),
               rebuild_ms_each=pms[3] / max(pcnt[3], 1), energy=collec.energy(), temp=collec.temp())
    print(json.dumps(out), flush=True)
    atoms.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    run("1: LJatoms-like LennardJonesCutPair N=1000", W.config1(), a.steps * 5)
    run("2: 2-D bidisperse harmonic RepulsionPair N=100k", W.config2(), a.steps * 2)
    run("3: 3-D LJAttractRepulsePair N=1M", W.config3(), a.steps)
    w4 = W.config4()
    w4["seed"] = 4004
    run("4: 3-D binary LJRepulsePair (WCA-like) N=4M, CollectionSol", w4, a.steps)


if __name__ == "__main__":
    main()
