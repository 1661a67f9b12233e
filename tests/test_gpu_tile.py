"""The cell-tile pair kernel (csrc/tile.cu, csrc/force_tile.cuh): positions staged per chunk in shared memory,
16-bit tile-local rows, OriginBox::diff (box.hpp:103) applied once per staged atom. Same bars as the gather
kernel: forces / energy / virial / stress within 1e-10 of the CPU oracle, identical rebuild steps."""
import os

import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import cpu_system, rel_err, rel_err_vec

pytestmark = pytest.mark.gpu
TOL = 1e-10


class _env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update({k: str(v) for k, v in self.kv.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _check_against_oracle(w, sim, steps=0, expect_wide=None):
    box, atoms, inter, nl, collec = sim.from_workload(w)
    active, chunks, max_tile, wide = nl.tile_stats()
    assert active and chunks > 0 and max_tile > 0, "the cell-tile kernel is not in use: %r" % ((active, chunks, max_tile),)
    if expect_wide is not None:
        assert (wide > 0) == expect_wide, "wide chunks: %d of %d" % (wide, chunks)
    c = cpu_system("port", w, injected=True)
    a, b = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(a, ca) and np.array_equal(b, cb)
    atoms.reset_forces()
    gp = inter.set_forces_get_pressure(box)
    cf, cp = c.forces_and_pressure()
    assert rel_err_vec(atoms.peek("f"), cf) < TOL
    assert rel_err(gp, cp) < TOL
    assert rel_err(inter.energy(box), c.inter_energy()) < TOL
    assert rel_err(inter.pressure(box), c.inter_pressure()) < TOL
    assert rel_err(inter.stress(box), c.inter_stress()) < TOL
    if steps:
        collec.set_forces(True)
        c.set_forces(True)
        collec.timestep(steps)
        c.timestep(steps)
        assert nl.which() == c.which() and nl.which() > 1
        assert nl.tile_stats()[0]
        assert rel_err_vec(atoms.peek("x") - w["x"], c.get_atoms()[0] - w["x"]) < 1e-9
        assert rel_err_vec(atoms.peek("f"), c.get_atoms()[3]) < 1e-8
        assert rel_err(collec.energy(), c.energy()) < TOL
        a, b = nl.pairs()
        ca, cb = c.pairs()
        assert np.array_equal(a, ca) and np.array_equal(b, cb)
    return nl


@pytest.mark.parametrize("team,v,ch", [(4, 8, 128), (4, 4, 64), (8, 4, 112), (2, 8, 96)])
def test_tile_narrow_chunks_vs_oracle(oracle_built, team, v, ch):
    """Box of 12 cells per axis: chunks cover a few cells of a column, no per-pair minimum image; every
    (team, entries per lane, chunk size) variant of the kernel."""
    from parm_b200 import sim
    with _env(PARM_B200_TILE_TEAM=team, PARM_B200_TILE_V=v, PARM_B200_TILE_CH=ch):
        w = W.lj_lattice((36, 36, 36), seed=811 + team)
        _check_against_oracle(w, sim, steps=14 if team == 4 and v == 8 else 0, expect_wide=False)


def test_tile_wide_chunks_small_box(oracle_built):
    """Five cells per axis: a chunk spans its whole column, the tile is wider than half the box and the kernel
    keeps the per-pair minimum image on top of the staged one."""
    from parm_b200 import sim
    w = W.lj_lattice((15, 15, 15), seed=812)
    # (systems this small run the 16-lane gather kernel / the persistent kernel by default: switch both off)
    with _env(PARM_B200_SMALL_TEAM=0, PARM_B200_SMALL_PERSIST=0):
        _check_against_oracle(w, sim, steps=20, expect_wide=True)


@pytest.mark.parametrize("small_paths", [0, 1])
def test_power_of_two_box_row_pads(oracle_built, small_paths):
    """L = 16 exactly: 1e100 / L is exact, so the per-pair minimum image folds the far-away sentinel of the row pads to
    distance ZERO -- pads must be recognised by their index (tile kernel with wide chunks, persistent small-system
    kernel), not by their distance. small_paths = 1: the default kernels for a system of this size."""
    from parm_b200 import sim
    w = W.lj_lattice((16, 16, 16), rho=1.0, seed=816)
    assert np.all(w["L"] == 16.0)
    env = {} if small_paths else dict(PARM_B200_SMALL_TEAM=0, PARM_B200_SMALL_PERSIST=0)
    with _env(**env):
        nl = _check_against_oracle(w, sim, steps=20, expect_wide=True)


def test_tile_unwrapped_coordinates(oracle_built):
    """ParM keeps unwrapped positions: atoms carry arbitrary multiples of L. The staged coordinate
    min_image(x - origin) must pick the right image of each of them."""
    from parm_b200 import sim
    w = W.lj_lattice((30, 30, 30), seed=813)
    rng = np.random.default_rng(5)
    w["x"] = w["x"] + rng.integers(-3, 4, w["x"].shape) * w["L"][None, :]
    _check_against_oracle(w, sim, steps=10, expect_wide=False)


@pytest.mark.parametrize("kind", [W.KIND_LJCUT, W.KIND_LJREPULSE])
def test_tile_other_lj_functors(oracle_built, kind):
    """LennardJonesCutPair and LJRepulsePair share the kernel (LJRepulsePair rows are short: force the tile path)."""
    from parm_b200 import sim
    with _env(PARM_B200_TILE_MIN_NEIGHBORS=0):
        w = W.lj_lattice((30, 30, 30), seed=814 + kind, kind=kind)
        _check_against_oracle(w, sim, steps=10)


def test_tile_equals_gather_kernel_1m():
    """BASELINE configs[2] size: both kernels on the same list agree to rounding (different summation order only)."""
    from parm_b200 import sim
    w = W.config3()
    out = []
    for tile in (1, 0):
        with _env(PARM_B200_TILE=tile):
            box, atoms, inter, nl, collec = sim.from_workload(w)
            assert nl.tile_stats()[0] == bool(tile)
            if tile:
                assert nl.tile_stats()[3] == 0  # no wide chunk at this size
            collec.set_forces(True)
            out.append((atoms.peek("f").copy(), collec.potential_energy(), collec.virial()))
            del box, atoms, inter, nl, collec
    assert rel_err_vec(out[0][0], out[1][0]) < 1e-12
    assert rel_err(out[0][1], out[1][1]) < 1e-12
    assert rel_err(out[0][2], out[1][2]) < 1e-12


def test_tile_with_ignored_pairs(oracle_built):
    """NeighborList::ignore compacts the rows before they are localised."""
    from parm_b200 import sim
    w = W.lj_lattice((24, 24, 24), seed=815)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    a, b = nl.pairs()
    sel = np.arange(0, len(a), 97)
    nl.ignore(a[sel], b[sel])
    nl.update_list(False)
    assert nl.tile_stats()[0]
    c = cpu_system("port", w, injected=True)
    c.ignore(a[sel], b[sel])
    c.update_list(True)
    ga, gb = nl.pairs()
    ca, cb = c.pairs()
    assert np.array_equal(ga, ca) and np.array_equal(gb, cb)
    collec.set_forces(True)
    c.set_forces(True)
    assert rel_err_vec(atoms.peek("f"), c.get_atoms()[3]) < TOL
    assert rel_err(collec.potential_energy(), c.potential_energy()) < TOL
