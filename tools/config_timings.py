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


def functor_lattice(kind, variant, side):
    """1e6-atom 3-D lattice at the LJ bench density with the parameters of W.functor_system for one functor."""
    import numpy as np
    base = W.lj_lattice((side, side, side), seed=3003, T=0.3)
    f = W.functor_system(kind, variant, ndim=3, n=base["x"].shape[0], seed=kind)
    for k in ("params", "types", "eps_table", "sig_table"):
        base[k] = f[k]
    base["params"][:, 1] = np.where(base["params"][:, 1] > 1.1, 1.0, 0.9)  # keep sigma_ij below the lattice spacing
    if kind == W.KIND_EISMCLACHLAN:
        base["params"][:, 1] = 0.55
    base.update(kind=kind, variant=variant, dt=0.001)
    return base


def widening(steps):
    """every NListed functor under CollectionVerlet and every extra integrator on the LJ workload,
    1e6 atoms each, same timing as the config table."""
    side = 100
    for kind, variant in [(W.KIND_LJATTRACTCUT, ""), (W.KIND_LJATTRACTFIXEDREPULSE, ""), (W.KIND_EISMCLACHLAN, ""),
                          (W.KIND_LJISH, ""), (W.KIND_LJATTRACTREPULSESIGS, ""), (W.KIND_REPULSIONDRAG, ""),
                          (W.KIND_LOISOHERN, ""), (W.KIND_LOISLIN, ""), (W.KIND_REPULSION, "II")]:
        run("functor kind %d%s, CollectionVerlet, N=1M" % (kind, variant), functor_lattice(kind, variant, side), steps)
    for integ, (name, params) in sorted(W.INTEGRATOR_CASES.items()):
        w = W.config3(side)
        if integ == 4:
            params = (0.05,)
        w.update(integrator=integ, integ_params=params, seed=1)
        run("integrator %s, LJAttractRepulsePair, N=1M" % name, w, steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--widening", action="store_true", help="time the SURVEY 8(f) functors and integrators instead")
    a = ap.parse_args()
    if a.widening:
        return widening(a.steps)
    run("1: LJatoms-like LennardJonesCutPair N=1000", W.config1(), a.steps * 5)
    run("2: 2-D bidisperse harmonic RepulsionPair N=100k", W.config2(), a.steps * 2)
    run("3: 3-D LJAttractRepulsePair N=1M", W.config3(), a.steps)
    w4 = W.config4()
    w4["seed"] = 4004
    run("4: 3-D binary LJRepulsePair (WCA-like) N=4M, CollectionSol", w4, a.steps)


if __name__ == "__main__":
    main()
