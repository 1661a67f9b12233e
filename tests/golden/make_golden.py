"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(oracle/_ref/libparm_ref{2,3}d.so, built from /root/reference/src by oracle/Makefile).

    python tests/golden/make_golden.py

The reference carries no golden vectors of its own, so these files ARE the
pin: inputs + the reference's outputs (pair lists in reference order, per-atom forces, energy,
virial, stress, short trajectories, Langevin steps with injected Gaussians).
Every array is stored exactly (float64 / uint32 in .npz); tests compare bit-for-bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import cpu  # noqa: E402
from parm_b200 import workloads as W  # noqa: E402
from parity_util import cpu_system  # noqa: E402

INPUT_KEYS = ("ndim", "L", "x", "v", "m", "kind", "params", "types", "eps_table", "skin", "dt", "integrator")
OPTIONAL_KEYS = ("damping", "T", "variant", "sig_table", "integ_params")


def record(w, steps=40, noise=None):
    s = cpu_system("ref", w, collection=True)
    out = {k: np.asarray(w[k]) for k in INPUT_KEYS}
    for k in OPTIONAL_KEYS:
        if k in w:
            out[k] = np.asarray(w[k])
    a, b = s.pairs()
    out["pairs_first"], out["pairs_last"] = a, b
    f, p = s.forces_and_pressure()
    out["forces"], out["virial"] = f, np.asarray(p)
    out["energy"] = np.asarray(s.inter_energy())
    out["stress"] = s.inter_stress()
    out["contacts"] = np.asarray(s.inter_contacts(), dtype=np.uint64)
    s.set_forces(True)
    if noise is not None:
        s.inject_noise(noise)
        out["noise"] = noise
    s.timestep(steps)
    x, v, acc, ff = s.get_atoms()
    out.update(steps=np.asarray(steps), x_end=x, v_end=v, a_end=acc, f_end=ff, which_end=np.asarray(s.which()),
               E_end=np.asarray(s.energy()), K_end=np.asarray(s.kinetic_energy()), P_end=np.asarray(s.pressure()),
               T_end=np.asarray(s.temp()), scalars_end=s.get_scalars(), L_end=s.get_box())
    if int(w.get("integrator", 0)) == 11:
        out["nlcg_end"] = s.nlcg_get()
    a, b = s.pairs()
    out["pairs_first_end"], out["pairs_last_end"] = a, b
    return out


def main():
    cpu.build(("ref",))
    cases = {}
    for ndim in (2, 3):
        for kind in range(4):
            w = W.random_system(96 if ndim == 3 else 80, ndim, kind, seed=100 + 10 * ndim + kind, ntypes=3, frozen=2, T=0.5)
            cases["random_%dd_kind%d" % (ndim, kind)] = record(w)
    cases["hertzian12"] = record(W.hertzian12(), steps=200)
    w = W.lj_lattice((6, 6, 6), seed=7)
    cases["lj_lattice216"] = record(w, steps=60)
    # CollectionSol (config 4 functor) with injected Gaussians
    w = W.config4(shape=(5, 5, 5), seed=44)
    rng = np.random.default_rng(4)
    steps = 25
    z = rng.standard_normal((steps, w["x"].shape[0], 2, 3))
    cases["sol_wca125"] = record(w, steps=steps, noise=z)
    w = W.random_system(70, 2, 1, seed=61, ntypes=1, frozen=3, T=0.3)
    w.update(integrator=W.SOL, damping=0.7, T=0.3)
    nm = int((w["m"] > 0).sum())
    z = np.random.default_rng(5).standard_normal((steps, nm, 2, 2))
    cases["sol_harmonic2d"] = record(w, steps=steps, noise=z)
    # the remaining NListed functors of sim.i:621-643
    for k, (kind, variant) in enumerate(W.FUNCTOR_CASES):
        ndim = 2 if k % 4 == 3 else 3
        w = W.functor_system(kind, variant, ndim=ndim, n=90 if ndim == 3 else 70, seed=500 + k)
        cases["functor_k%d%s_%dd" % (kind, variant, ndim)] = record(w, steps=30)
    # the other fixed-box integrators
    for integ, (name, _) in W.INTEGRATOR_CASES.items():
        ndim = 2 if integ % 3 == 0 else 3
        w = W.integrator_system(integ, ndim=ndim, n=100 if ndim == 3 else 80, seed=700 + integ)
        z = None
        if integ == 3:
            z = np.random.default_rng(integ).standard_normal((30, int((w["m"] > 0).sum()), ndim))
        cases["integ_%s_%dd" % (name, ndim)] = record(w, steps=30, noise=z)
    # CollectionNLCG, set up like pyparm/packmin.py
    for ndim in (2, 3):
        cases["nlcg_packer_%dd" % ndim] = record(W.packer_system(ndim=ndim, n=120, seed=30 + ndim), steps=40)
    for name, d in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "pairs", len(d["pairs_first"]), "E", float(d["energy"]), "which_end", int(d["which_end"]))


if __name__ == "__main__":
    main()
