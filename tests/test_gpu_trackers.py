"""RsqTracker, ISFTracker and EnergyTracker with their accumulators on the device, registered with a
Collection (updated at the end of every step without a host round trip), against the oracle -- the reference's own
constraints.cpp compiled in place when present, else the C restatement."""
import numpy as np
import pytest

from parm_b200 import workloads as W
from parity_util import backends, cpu_system, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ndim,integ", [(3, 0), (2, 0), (3, 1), (3, 5)], ids=["verlet3d", "verlet2d", "sol3d", "nosehoover3d"])
def test_trackers_match_oracle(oracle_built, ndim, integ):
    from parm_b200 import sim
    if ndim == 3:
        w = W.lj_lattice((9, 9, 9), seed=31 + integ, T=1.2)
        w["m"] = np.random.default_rng(1).uniform(0.5, 2.0, w["m"].shape)  # mass-weighted centre of mass
    else:
        w = W.config2(nx=30, ny=36, seed=33)
        w["v"] = w["v"] * 15.0
    steps = 90
    if integ == 1:
        w.update(integrator=1, damping=0.5, T=1.0)
    elif integ == 5:
        w.update(integrator=5, integ_params=(20.0, 1.0))
    be = backends(oracle_built)[-1]
    box, atoms, inter, nl, collec = sim.from_workload(w)
    s = cpu_system(be, w, injected=True)
    collec.set_forces(True)
    s.set_forces(True)
    ns, ks, isf_ns = [1, 4, 25], [0.9, 6.1], [3, 10]
    rsq = sim.RsqTracker(atoms, ns, True)
    rsq0 = sim.RsqTracker(atoms, [2], False)
    isf = sim.ISFTracker(atoms, ks, isf_ns, False)
    et = sim.EnergyTracker(atoms, [inter], 4)
    for t in (rsq, rsq0, isf, et):
        collec.add_tracker(t)
    h = [s.add_rsq_tracker(ns, True), s.add_rsq_tracker([2], False), s.add_isf_tracker(ks, isf_ns, False), s.add_energy_tracker(4)]
    if integ == 1:
        z = np.random.default_rng(2).standard_normal((steps + 20, int((w["m"] > 0).sum()), 2, ndim))
        collec.inject_noise(z)
        s.inject_noise(z)
    collec.timestep(steps)
    s.timestep(steps)
    assert nl.which() == s.which() and nl.which() > 1

    def check():
        assert [int(c) for c in rsq.counts()] == s.tracker_counts(h[0])
        for k in range(len(ns)):
            if s.tracker_counts(h[0])[k] == 0:  # no update of this lag since reset(): the means are 0/0 on both sides
                assert np.isnan(rsq.xyz2()[k]).all()
                continue
            a, b, c = s.rsq_read(h[0], k)
            assert a.max() > 0
            assert rel_err(rsq.xyz2()[k], a) < 1e-8 and rel_err(rsq.xyz4()[k], b) < 1e-8 and rel_err(rsq.r4()[k], c) < 1e-8
        assert rel_err(rsq0.xyz2()[0], s.rsq_read(h[1], 0)[0]) < 1e-8
        assert rel_err(rsq.r2()[1], s.rsq_read(h[0], 1)[0].sum(axis=1)) < 1e-8
        assert [int(c) for c in isf.counts()] == s.tracker_counts(h[2])
        for k in range(len(isf_ns)):
            ref = s.isf_read(h[2], k)
            assert np.abs(isf.ISFxyz()[k] - ref).max() < 1e-8
            assert np.abs(isf.ISFs()[k] - ref.mean(axis=2)).max() < 1e-8
        e = s.energy_tracker_read(h[3])
        assert et.n() == int(e[0]) > 0
        got = [et.E(), et.U(), et.K(), et.E_squared_mean(), et.U_squared_mean(), et.K_squared_mean(), et.get_U0()]
        assert rel_err(got[:3], e[1:4]) < 1e-9 and rel_err(got[3:6], e[4:7]) < 1e-9 and rel_err(got[6], e[7]) < 1e-9 or e[7] == 0
        assert et.E_std() >= 0

    check()
    # set_U0(box), reset(), explicit update() calls, then more steps one at a time
    et.set_U0()
    s.energy_tracker_set_U0(h[3])
    rsq.reset()
    s.tracker_reset(h[0])
    isf.update(box)
    s.tracker_update(h[2])
    for _ in range(20):
        collec.timestep()
    s.timestep(20)
    check()


def test_tracker_argument_errors():
    from parm_b200 import sim
    atoms = sim.AtomVec(np.ones(10), ndim=3)
    with pytest.raises(Exception):
        sim.RsqTracker(atoms, [0, 3])  # t % 0
