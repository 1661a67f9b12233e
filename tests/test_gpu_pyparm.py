"""the `pyparm` package (d2/d3 with the SWIG names of sim.i:616-670), trajectory output from
asynchronous frame downloads (pyparm/xyzfile.py:7-76, LJatoms.cpp:130-158) and Grid (trackers.hpp:227-309)."""
import io
import math

import numpy as np
import pytest

from parm_b200 import workloads as W

pytestmark = pytest.mark.gpu


class _PairedCollection:
    """Body of pyparm/tests.py:218-281 (PairedCollectionTest.setUp / resetPositions / reset), written against the
    module object `sim3` exactly as the reference writes it."""
    phi, N = 0.3, 12

    def __init__(self, sim3, seed, collection_type, collection_args, pair_type, atom_type, atom_args):
        np.random.seed(seed)
        sigmas = [1.] * (self.N // 2) + [1.4] * (self.N - self.N // 2)
        masses = [s ** 3 for s in sigmas]
        radii = np.asarray(sigmas) / 2.
        V = np.sum(radii ** 3) * 4 / 3 * np.pi / self.phi
        self.L = float(V ** (1. / 3.))
        self.box = sim3.OriginBox(self.L)
        self.atoms = sim3.AtomVec(masses)
        self.interaction = pair_type(self.box, self.atoms, 0.4)
        for args in zip(self.atoms, *atom_args(self.N, sigmas)):
            self.interaction.add(atom_type(*args))
        self.masses = masses
        self.resetPositions()
        args = [self.box, self.atoms] + list(collection_args) + [[self.interaction], [self.interaction.neighbor_list()], []]
        self.collec = collection_type(*args)

    def resetPositions(self):
        np.random.seed(131)
        for a, m in zip(self.atoms, self.masses):
            a.x = np.random.uniform(0., self.L, size=(3,))
            a.v = np.random.normal(size=(3,)) / m
            a.f = np.random.normal(size=(3,))
        self.interaction.neighbor_list().update_list(True)

    def reset(self):
        self.resetPositions()
        self.collec.scale_velocities_to_temp(1.0)
        for _ in range(1000):
            self.collec.timestep()
            self.collec.scale_velocities_to_temp(1.0)


def test_pyparm_d3_random_hertzian_verlet():
    """pyparm/tests.py:284-317 (RandomHertzianVerletTest.testEnergy) through `from pyparm import d3 as sim3`."""
    from pyparm import d3 as sim3
    t = _PairedCollection(sim3, 131, sim3.CollectionVerlet, [0.01], sim3.Repulsion, sim3.EpsSigExpAtom,
                          lambda N, sig: ([1.2] * N, sig, [2.0] * N))
    t.reset()
    collec = t.collec
    EKUTs = []
    lastE = collec.energy()
    for _ in range(300):
        for _ in range(10):
            collec.timestep()
            assert np.allclose(collec.energy(), lastE, rtol=1e-2)
            lastE = collec.energy()
        EKUTs.append((collec.energy(), collec.kinetic_energy(), collec.potential_energy(), collec.temp()))
    E, K, U, T = np.asarray(EKUTs).T
    assert np.allclose(np.mean(T), 1.0, rtol=2e-1)
    assert np.allclose(np.std(E) / np.mean(E), 0.0, atol=1e-2)


def test_pyparm_module_names():
    """The NListed instantiations of sim.i:621-643 and the classes the reference's scripts construct."""
    from pyparm import d2, d3
    names = ["OriginBox", "AtomVec", "NeighborList", "LJRepulse", "LJAttractCut", "LJAttractICut", "LJAttractIICut", "LJIICut",
             "LJAttractRepulse", "LJAttractFixedRepulse", "EisMclachlan", "LJish", "LJAttractRepulseSigs", "Repulsion",
             "RepulsionII", "HertzianDrag", "LoisOhern", "LoisLin", "LoisLinMin", "LoisOhernMin", "CollectionVerlet",
             "CollectionSol", "CollectionNLCG", "CollectionNoseHoover", "EpsSigAtom", "EpsSigExpAtom", "IEpsSigCutAtom",
             "RsqTracker", "ISFTracker", "EnergyTracker", "Grid", "Vec"]
    for m in (d2, d3):
        for n in names:
            assert hasattr(m, n), (m.__name__, n)
    assert d2.NDIM == 2 and d3.NDIM == 3
    assert d2.OriginBox(3.0).box_shape().shape == (2,) and d3.AtomVec(5, 1.0).ndim == 3


def test_xyz_writer_async_frames_match_state():
    """Frames started before further timestep() calls hold the state of the moment they were started."""
    from parm_b200 import sim
    from pyparm.xyzfile import XYZwriter, NPZwriter
    w = W.lj_lattice((10, 10, 10), seed=5)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(30)
    x0, v0 = atoms.peek("x").copy(), atoms.peek("v").copy()
    f = io.StringIO()
    wr = XYZwriter(f, usevels=True)
    wr.begin_frame(atoms, time=0.12)
    collec.timestep(40)           # overlaps the copy of the frame
    wr.finish_frame()
    lines = f.getvalue().splitlines()
    assert int(lines[0]) == atoms.n and lines[1].startswith("time=0.12")
    data = np.array([[float(t) for t in ln.split()[1:]] for ln in lines[2:]])
    assert data.shape == (atoms.n, 6)
    assert np.abs(data[:, :3] - x0).max() <= 5.1e-5 and np.abs(data[:, 3:] - v0).max() <= 5.1e-5
    assert np.abs(atoms.peek("x") - x0).max() > 1e-3   # the run did move on meanwhile
    import os
    import tempfile
    p = os.path.join(tempfile.mkdtemp(), "traj.npz")
    nz = NPZwriter(p)
    for k in range(3):
        nz.begin_frame(atoms, time=float(k))
        collec.timestep(5)
    xk = atoms.peek("x")
    nz.close()
    d = np.load(p)
    assert d["x"].shape == (3, atoms.n, 3) and d["v"].shape == (3, atoms.n, 3)
    assert not np.array_equal(d["x"][2], xk) and np.array_equal(d["time"], [0.0, 1.0, 2.0])


@pytest.mark.parametrize("ndim", [3, 2])
def test_grid_locs_match_reference_rule(ndim):
    """Grid::get_loc on the device (trackers.cpp:192-219) against its restatement with math.remainder, unwrapped
    coordinates included; all_pairs() covers every pair closer than a cell width."""
    from parm_b200 import sim
    rng = np.random.default_rng(7)
    n = 500
    L = np.array([7.0, 8.5, 6.25][:ndim])
    x = rng.uniform(-2 * L, 3 * L, (n, ndim))
    x[0] = 0.0
    x[1] = L          # == widths -> 0
    x[2] = L / 2
    box = sim.OriginBox(L)
    atoms = sim.AtomVec(n, 1.0, ndim)
    for i, a in enumerate(atoms):
        a.x = x[i]
    g = sim.Grid(box, atoms, [4, 5, 3][:ndim])
    g.make_grid()

    def ref_loc(v):
        k = []
        for d in range(ndim):
            r = math.remainder(v[d] - L[d] / 2.0, L[d]) + L[d] / 2.0
            q = int(math.floor(r * g.widths[d] / L[d]))
            k.append(0 if q == g.widths[d] else q)
        return (k[2] * g.widths[1] + k[1]) * g.widths[0] + k[0] if ndim == 3 else k[1] * g.widths[0] + k[0]

    want = np.array([ref_loc(v) for v in x])
    assert np.array_equal(g.locs, want)
    assert sorted(sum(g.gridlocs, [])) == list(range(n))
    assert g.get_loc(x[5]) == want[5]
    pairs = set(g.all_pairs())
    cw = (L / np.array(g.widths)).min()
    d = box.diff(np.repeat(x, n, axis=0), np.tile(x, (n, 1)), atoms).reshape(n, n, ndim)
    close = np.argwhere(np.sqrt((d * d).sum(-1)) < cw * 0.999)
    for i, j in close:
        if j < i:
            assert (int(i), int(j)) in pairs
    assert set(g.all_pairs(7)) == {j for (i, j) in pairs if i == 7} | {i for (i, j) in pairs if j == 7}


def test_cpp_facade_grid_and_async_frames():
    """examples/facade_grid_snapshot.cpp: Grid and AtomVec::snapshot_begin / snapshot_wait through the drop-in headers."""
    import os
    import subprocess
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "bin", "facade_grid_snapshot3d")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
