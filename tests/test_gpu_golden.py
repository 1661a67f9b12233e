"""GPU vs the committed golden fixtures (outputs of the unmodified reference, tests/golden/)."""
import glob
import os

import numpy as np
import pytest

from parity_util import rel_err, rel_err_vec
from test_oracle_pin import load

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gpu_reproduces_reference_fixture(path):
    from parm_b200 import sim
    w, d = load(path)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    a, b = nl.pairs()
    assert np.array_equal(a, d["pairs_first"]) and np.array_equal(b, d["pairs_last"])  # bit-exact pair set
    atoms.reset_forces()
    p = inter.set_forces_get_pressure(box)
    assert rel_err_vec(atoms.peek("f"), d["forces"]) < 1e-10
    assert rel_err(p, d["virial"]) < 1e-10
    assert rel_err(inter.energy(box), d["energy"]) < 1e-10
    assert rel_err(inter.stress(box), d["stress"]) < 1e-10
    if "contacts" in d:
        assert (inter.contacts(box), inter.overlaps(box)) == tuple(int(q) for q in d["contacts"])
    collec.set_forces(True)
    if "noise" in d:
        collec.inject_noise(d["noise"])
    collec.timestep(int(d["steps"]))
    assert nl.which() == int(d["which_end"])
    assert rel_err_vec(atoms.peek("x") - w["x"], d["x_end"] - w["x"]) < 1e-9
    assert rel_err_vec(atoms.peek("v"), d["v_end"]) < 1e-9
    assert rel_err_vec(atoms.peek("a"), d["a_end"]) < 1e-8
    assert rel_err(collec.energy(), d["E_end"]) < 1e-10
    assert rel_err(collec.kinetic_energy(), d["K_end"]) < 1e-10
    assert rel_err(collec.pressure(), d["P_end"]) < 1e-9
    assert rel_err(collec.temp(), d["T_end"]) < 1e-10
    if "L_end" in d:
        assert rel_err(box.box_shape(), d["L_end"]) < 1e-10
    a, b = nl.pairs()
    assert np.array_equal(a, d["pairs_first_end"]) and np.array_equal(b, d["pairs_last_end"])
