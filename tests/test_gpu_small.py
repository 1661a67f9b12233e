"""Small systems (BASELINE config 1, LJatoms.cpp N = 1000): CollectionVerlet::timestep(n) runs as one persistent
cooperative kernel (csrc/small.cu). Same bars as everywhere: trajectories against the CPU oracle, identical rebuild steps;
and the path must really be the one that ran (a handful of launches for hundreds of steps)."""
import os

import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu


def _run(w, steps, persist):
    from parm_b200 import sim
    old = os.environ.get("PARM_B200_SMALL_PERSIST")
    os.environ["PARM_B200_SMALL_PERSIST"] = "1" if persist else "0"
    try:
        box, atoms, inter, nl, collec = sim.from_workload(w)
        collec.set_forces(True)
        l0 = collec.stats()["launches"]
        collec.timestep(steps)
        out = dict(x=atoms.peek("x").copy(), v=atoms.peek("v").copy(), f=atoms.peek("f").copy(), E=collec.energy(),
                   which=nl.which(), launches=collec.stats()["launches"] - l0, pairs=nl.pairs())
        atoms.close()
        return out
    finally:
        if old is None:
            os.environ.pop("PARM_B200_SMALL_PERSIST", None)
        else:
            os.environ["PARM_B200_SMALL_PERSIST"] = old


def test_config1_persistent_kernel_vs_oracle(oracle_built):
    """LJatoms.cpp's system (LennardJonesCutPair, N = 1000, skin 1.0): 400 steps, no rebuild in between."""
    w = W.config1()
    g = _run(w, 400, True)
    assert g["launches"] <= 8, "the persistent kernel did not run (%d launches for 400 steps)" % g["launches"]
    c = cpu_system("port", w, injected=True)
    c.set_forces(True)
    c.timestep(400)
    cx, cv, _, cf = c.get_atoms()
    assert rel_err_vec(g["x"] - w["x"], cx - w["x"]) < 1e-9
    assert rel_err_vec(g["v"], cv) < 1e-9
    assert rel_err_vec(g["f"], cf) < 1e-8
    assert rel_err(g["E"], c.energy()) < 1e-10
    assert g["which"] == c.which()


def test_persistent_kernel_across_rebuilds_vs_oracle_and_per_step_path(oracle_built):
    """Hot LJ liquid, 2744 atoms: the drift rule fires every few steps; the kernel must stop after exactly the step that
    triggered (collection.cpp:468), the host rebuilds, the next launch carries on. Against the oracle and against the
    per-step kernels (same K1 / K3 expressions: positions agree to rounding of the pair arithmetic)."""
    w = W.lj_lattice((14, 14, 14), seed=4321)
    steps = 60
    g = _run(w, steps, True)
    p = _run(w, steps, False)
    c = cpu_system("port", w, injected=True)
    c.set_forces(True)
    c.timestep(steps)
    assert g["which"] == c.which() == p["which"] and g["which"] > 4
    assert g["launches"] < p["launches"] // 2
    cx = c.get_atoms()[0]
    assert rel_err_vec(g["x"] - w["x"], cx - w["x"]) < 1e-9
    assert rel_err_vec(g["x"] - w["x"], p["x"] - w["x"]) < 1e-10
    assert rel_err(g["E"], c.energy()) < 1e-10
    ca, cb = c.pairs()
    assert np.array_equal(g["pairs"][0], ca) and np.array_equal(g["pairs"][1], cb)


def test_persistent_kernel_frozen_atoms_and_single_steps(oracle_built):
    """Frozen atoms (m = inf: v = 0, a = 0, position untouched, collection.cpp:447,460) and a call pattern that mixes
    timestep(1) (general path) with timestep(n)."""
    from parm_b200 import sim
    w = W.lj_lattice((10, 10, 10), seed=77)
    w["m"] = w["m"].copy()
    w["m"][::7] = np.inf
    box, atoms, inter, nl, collec = sim.from_workload(w)
    c = cpu_system("port", w, injected=True)
    collec.set_forces(True)
    c.set_forces(True)
    for k in (1, 17, 1, 2, 30):
        collec.timestep(k)
        c.timestep(k)
    cx, cv, ca_, cf = c.get_atoms()
    assert rel_err_vec(atoms.peek("x") - w["x"], cx - w["x"]) < 1e-9
    assert np.array_equal(atoms.peek("x")[::7], w["x"][::7])
    assert np.all(atoms.peek("v")[::7] == 0.0) and np.all(atoms.peek("a")[::7] == 0.0)
    assert rel_err_vec(atoms.peek("v"), cv) < 1e-9
    assert nl.which() == c.which()
