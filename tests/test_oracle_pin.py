"""CPU-only: pins the oracles.

1. oracle/parm_oracle.c (the plain-C restatement) must reproduce, bit for bit, every golden
   fixture in tests/golden/ -- those were produced by the UNMODIFIED reference sources compiled
   into oracle/_ref (tests/golden/make_golden.py).
2. Where the compiled reference is present (this container, and the GPU box via the prebuilt .so)
   it is re-run and must reproduce the fixtures too, and agree with the port on fresh inputs.
3. The cell-list pair finder used for large N equals NeighborList::update_list(true).
4. The reference's own hot-path test (pyparm/tests.py:284-317) holds on the oracle.
"""
import glob
import math
import os

import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import backends, cpu_system

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    d = dict(np.load(path))
    w = {k: d[k] for k in ("L", "x", "v", "m", "params", "types", "eps_table")}
    for k in ("ndim", "kind", "integrator"):
        w[k] = int(d[k])
    for k in ("skin", "dt"):
        w[k] = float(d[k])
    for k in ("damping", "T"):
        if k in d:
            w[k] = float(d[k])
    if "variant" in d:
        w["variant"] = str(d["variant"])
    if "sig_table" in d:
        w["sig_table"] = d["sig_table"]
    if "integ_params" in d:
        w["integ_params"] = tuple(float(q) for q in d["integ_params"])
    return w, d


def replay(backend, w, d):
    s = cpu_system(backend, w, collection=True)
    out = {}
    out["pairs_first"], out["pairs_last"] = s.pairs()
    f, p = s.forces_and_pressure()
    out["forces"], out["virial"] = f, np.asarray(p)
    out["energy"] = np.asarray(s.inter_energy())
    out["stress"] = s.inter_stress()
    out["contacts"] = np.asarray(s.inter_contacts(), dtype=np.uint64)
    s.set_forces(True)
    if "noise" in d:
        s.inject_noise(d["noise"])
    s.timestep(int(d["steps"]))
    x, v, a, ff = s.get_atoms()
    out.update(x_end=x, v_end=v, a_end=a, f_end=ff, which_end=np.asarray(s.which()), E_end=np.asarray(s.energy()),
               K_end=np.asarray(s.kinetic_energy()), P_end=np.asarray(s.pressure()), T_end=np.asarray(s.temp()),
               scalars_end=s.get_scalars(), L_end=s.get_box())
    if int(w.get("integrator", 0)) == 11:
        out["nlcg_end"] = s.nlcg_get()
    out["pairs_first_end"], out["pairs_last_end"] = s.pairs()
    return out


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_bit_exact(oracle_built, path):
    assert GOLDEN, "golden fixtures missing"
    w, d = load(path)
    for be in backends(oracle_built):
        out = replay(be, w, d)
        for k, v in out.items():
            assert np.array_equal(np.asarray(v), d[k]), "%s: %s differs from the reference fixture" % (be, k)


@pytest.mark.parametrize("ndim,kind", [(3, 2), (2, 1), (3, 0), (3, 3), (2, 2)])
def test_cell_list_equals_reference_rebuild(oracle_built, ndim, kind):
    """InjectedNeighborList / port cell list == NeighborList::update_list(true) (trackers.cpp:55-69)."""
    w = W.random_system(2500, ndim, kind, seed=3 + kind, ntypes=2)
    for be in backends(oracle_built):
        a = cpu_system(be, w, injected=False, collection=False).pairs()
        b = cpu_system(be, w, injected=True, collection=False).pairs()
        assert len(a[0]) > 1000
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_port_equals_reference_fresh_inputs(oracle_built):
    if "ref" not in backends(oracle_built):
        pytest.skip("compiled reference not present")
    w = W.lj_lattice((7, 7, 7), seed=123)
    r, p = cpu_system("ref", w), cpu_system("port", w)
    for s in (r, p):
        s.set_forces(True)
        s.timestep(120)
    for a, b in zip(r.get_atoms(), p.get_atoms()):
        assert np.array_equal(a, b)
    assert r.which() == p.which() and r.energy() == p.energy() and r.pressure() == p.pressure()


@pytest.mark.parametrize("kind,variant", W.FUNCTOR_CASES, ids=["k%d%s" % c for c in W.FUNCTOR_CASES])
def test_port_equals_reference_all_functors(oracle_built, kind, variant):
    """the C restatement of every NListed functor is bit-equal to the compiled reference, with
    discrete species and with per-atom continuous parameters, in 2-D and 3-D."""
    if "ref" not in backends(oracle_built):
        pytest.skip("compiled reference not present")
    for ndim, cont in ((3, False), (2, True)):
        w = W.functor_system(kind, variant, ndim=ndim, n=150, seed=900 + kind, continuous=cont)
        r, p = cpu_system("ref", w), cpu_system("port", w)
        assert len(r.pairs()[0]) > 200
        for s in (r, p):
            s.set_forces(True)
        fr, pr = r.forces_and_pressure()
        fp, pp = p.forces_and_pressure()
        assert np.abs(fr).max() > 0 and np.array_equal(fr, fp) and pr == pp
        assert r.inter_energy() == p.inter_energy() and r.inter_energy() != 0
        assert np.array_equal(r.inter_stress(), p.inter_stress())
        assert r.inter_contacts() == p.inter_contacts() and r.inter_contacts()[0] > 0
        for s in (r, p):
            s.set_forces(True)
            s.timestep(50)
        for a, b in zip(r.get_atoms(), p.get_atoms()):
            assert np.array_equal(a, b)
        assert r.energy() == p.energy()


@pytest.mark.parametrize("integ", sorted(W.INTEGRATOR_CASES), ids=[W.INTEGRATOR_CASES[k][0] for k in sorted(W.INTEGRATOR_CASES)])
def test_port_equals_reference_all_integrators(oracle_built, integ):
    """the C restatement of every extra integrator is bit-equal to the compiled reference."""
    if "ref" not in backends(oracle_built):
        pytest.skip("compiled reference not present")
    for ndim in (3, 2):
        w = W.integrator_system(integ, ndim=ndim, n=160, seed=40 + integ)
        outs = []
        for be in ("ref", "port"):
            s = cpu_system(be, w)
            s.set_forces(True)
            if integ == 3:
                s.inject_noise(np.random.default_rng(1).standard_normal((40, int((w["m"] > 0).sum()), ndim)))
            s.timestep(40)
            outs.append((s.get_atoms(), s.energy(), s.which(), s.get_scalars()))
        for a, b in zip(outs[0][0], outs[1][0]):
            assert np.array_equal(a, b)
        assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2]
        if integ == 5:
            assert np.array_equal(outs[0][3], outs[1][3]) and outs[0][3][0] != 0


@pytest.mark.parametrize("ndim", [3, 2])
def test_trackers_port_equals_reference(oracle_built, ndim):
    """RsqTracker / ISFTracker / EnergyTracker of the C restatement against the reference's own
    constraints.cpp (compiled in place), bit for bit -- including reset(), set_U0(box) and frozen atoms."""
    if "ref" not in backends(oracle_built):
        pytest.skip("compiled reference not present")
    w = W.random_system(130, ndim, 2, seed=15, ntypes=2, frozen=2, T=0.5)
    outs = []
    for be in ("ref", "port"):
        s = cpu_system(be, w)
        s.set_forces(True)
        r = s.add_rsq_tracker([1, 3, 10], True)
        i = s.add_isf_tracker([0.7, 6.3], [2, 5], False)
        e = s.add_energy_tracker(3)
        s.timestep(37)
        o = [s.tracker_counts(r), s.rsq_read(r, 0), s.rsq_read(r, 2), s.tracker_counts(i), s.isf_read(i, 1), s.energy_tracker_read(e)]
        s.energy_tracker_set_U0(e)
        s.tracker_reset(r)
        s.tracker_update(i)
        s.timestep(11)
        o += [s.rsq_read(r, 1), s.isf_read(i, 0), s.energy_tracker_read(e), s.tracker_counts(r)]
        outs.append(o)

    def eq(a, b):
        if isinstance(a, (tuple, list)):
            return all(eq(x, y) for x, y in zip(a, b))
        return np.array_equal(np.asarray(a), np.asarray(b))
    assert outs[0][0] == [40, 13, 4] and outs[0][5][0] > 0
    for a, b in zip(*outs):
        assert eq(a, b)


def chain_ignores(n, rng, extra=40):
    """Bonded-neighbour style exclusions (i,i+1), (i,i+2) plus a few random pairs, duplicates and both orders."""
    a = np.concatenate([np.arange(n - 1), np.arange(n - 2), rng.integers(0, n, extra), np.arange(5)])
    b = np.concatenate([np.arange(1, n), np.arange(2, n), rng.integers(0, n, extra), np.arange(1, 6)])
    keep = a != b
    flip = rng.random(a.size) < 0.5
    a, b = np.where(flip, b, a)[keep], np.where(flip, a, b)[keep]
    return a.astype(np.uint32), b.astype(np.uint32)


def test_ignore_port_equals_reference(oracle_built):
    """NeighborList::ignore (trackers.hpp:190-193, trackers.cpp:64): excluded pairs never enter the list, the
    next update_list(false) rebuilds; O(N^2) reference loop == cell-list finder == C port."""
    if "ref" not in backends(oracle_built):
        pytest.skip("compiled reference not present")
    rng = np.random.default_rng(5)
    w = W.random_system(900, 3, 2, seed=77, ntypes=2)
    a, b = chain_ignores(900, rng)
    lists = []
    for be, inj in (("ref", False), ("ref", True), ("port", False), ("port", True)):
        s = cpu_system(be, w, injected=inj)
        n0 = len(s.pairs()[0])
        w0 = s.which()
        s.ignore(a, b)
        assert s.update_list(False) and s.which() == w0 + 1  # ignorechanged forces the rebuild
        pa, pb = s.pairs()
        assert 0 < len(pa) < n0
        key = set(zip(np.maximum(a, b).tolist(), np.minimum(a, b).tolist()))
        assert s.ignore_size() == len(key)
        assert not (set(zip(pa.tolist(), pb.tolist())) & key)
        s.set_forces(True)
        s.timestep(40)
        lists.append((pa, pb, s.get_atoms(), s.energy()))
    for other in lists[1:]:
        assert np.array_equal(lists[0][0], other[0]) and np.array_equal(lists[0][1], other[1])
        for x, y in zip(lists[0][2], other[2]):
            assert np.array_equal(x, y)
        assert lists[0][3] == other[3]


def test_box_diff_is_ieee_remainder(oracle_built):
    rng = np.random.default_rng(0)
    L = np.array([3.0, 4.5, 7.25])
    s = cpu_system("port", dict(L=L, x=np.zeros((2, 3)), v=np.zeros((2, 3)), m=np.ones(2), kind=0, skin=0.1,
                                params=np.ones((2, 3))), collection=False)
    for _ in range(200):
        r1, r2 = rng.uniform(-30, 30, 3), rng.uniform(-30, 30, 3)
        assert np.array_equal(s.box_diff(r1, r2), np.array([math.remainder(a, b) for a, b in zip(r1 - r2, L)]))


def test_reference_hertzian_verlet_test(oracle_built):
    """pyparm/tests.py:284-317 RandomHertzianVerletTest.testEnergy, on the oracle: warm up with a
    thermostat for 1000 steps, then 10^4 NVE steps: |dE|/E < 1e-2 per step, <T> = 1 +- 10 %, std(E)/mean(E) < 1e-2."""
    w = W.hertzian12()
    be = backends(oracle_built)[-1]
    s = cpu_system(be, w, collection=False)
    s.make_collection(0, w["dt"])
    # constructor with interactions/trackers given -> initialize(): set_forces(true)
    s.set_forces(True)
    s.scale_velocities_to_temp(1.0)
    for _ in range(1000):
        s.timestep(1)
        s.scale_velocities_to_temp(1.0)
    lastE = s.energy()
    Es, Ts = [], []
    for _ in range(1000):
        for _ in range(10):
            s.timestep(1)
            E = s.energy()
            assert abs(E - lastE) <= 1e-2 * abs(lastE)
            lastE = E
        Es.append(E)
        Ts.append(s.temp())
    assert abs(np.mean(Ts) - 1.0) < 0.1
    assert abs(np.std(Es) / np.mean(Es)) < 1e-2


@pytest.mark.parametrize("integ", [0, 1], ids=["verlet", "sol"])
def test_port_equals_reference_frozen_mass_rules(oracle_built, integ):
    """Frozen atoms under both hot-path integrators: m == 0, m < 0 and m == inf. CollectionVerlet tests
    `m <= 0 || isinf(m)` in both loops (collection.cpp:445,459); CollectionSol tests `m <= 0` in its first loop
    (plus isinf) but `m == 0` later (collection.cpp:278 vs :304,313) -- the C restatement must reproduce that asymmetry bit
    for bit, because the GPU kernels are checked against it."""
    if "ref" not in backends(oracle_built):
        pytest.skip("compiled reference not present")
    w = W.random_system(180, 3, W.KIND_LJREPULSE, seed=4242, ntypes=1, frozen=6, T=0.3)
    rng = np.random.default_rng(5)
    free = np.nonzero(w["m"] > 0)[0]
    w["m"][rng.choice(free, 3, replace=False)] = np.inf
    if integ == 0:  # a negative mass would enter CollectionSol's sqrt(T/m): Verlet only
        w["m"][rng.choice(np.nonzero(np.isfinite(w["m"]) & (w["m"] > 0))[0], 3, replace=False)] = -1.0
    w["v"] = w["v"].copy()
    w["v"][~np.isfinite(w["v"])] = 0.0
    if integ == 1:
        w.update(integrator=1, damping=1.0, T=0.3)
    outs = []
    for be in ("ref", "port"):
        s = cpu_system(be, w)
        s.set_forces(True)
        if integ == 1:
            # the first loop skips `m <= 0 or isinf(m)` before drawing (collection.cpp:278-281): those atoms draw no noise
            nmob = int(((w["m"] > 0) & np.isfinite(w["m"])).sum())
            s.inject_noise(np.random.default_rng(9).standard_normal((60, nmob, 2, 3)))
        s.timestep(60)
        outs.append((s.get_atoms(), s.which(), s.kinetic_energy()))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b, equal_nan=True)
    assert outs[0][1] == outs[1][1] and (outs[0][2] == outs[1][2] or (np.isnan(outs[0][2]) and np.isnan(outs[1][2])))
    x_end = outs[0][0][0]
    frozen = (w["m"] <= 0) | np.isinf(w["m"])
    if integ == 0:
        assert np.array_equal(x_end[frozen], w["x"][frozen])  # frozen atoms never move under Verlet
