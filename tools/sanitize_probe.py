"""Small end-to-end run that touches every kernel family; meant to be run under compute-sanitizer."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from parm_b200 import sim, workloads as W
def go(w, steps=12, trackers=False, ignore=False):
    box, atoms, inter, nl, collec = sim.from_workload(w)
    if ignore:
        n = atoms.n
        nl.ignore(np.arange(n - 1), np.arange(1, n))
    collec.set_forces(True)
    if trackers:
        for t in (sim.RsqTracker(atoms, [1, 3], True), sim.ISFTracker(atoms, [1.0], [2], False), sim.EnergyTracker(atoms, [inter], 2)):
            collec.add_tracker(t)
    collec.timestep(steps)
    e = collec.energy()
    inter.stress(box); inter.contacts(box); nl.pairs()
    atoms.close()
    return e
print("lj", go(W.lj_lattice((7, 7, 7), seed=1), trackers=True))
print("lj2species", go(dict(W.lj_lattice((8, 8, 8), seed=2), types=(np.arange(512) % 2).astype(np.uint32), eps_table=np.array([[1.0, 1.5], [1.5, 0.5]])), ignore=True))
print("2d", go(W.config2(nx=20, ny=24)))
for kind, variant in W.FUNCTOR_CASES:
    print("functor", kind, variant, go(W.functor_system(kind, variant, ndim=3, n=300, seed=kind), steps=6))
    print("functor-cont", kind, variant, go(W.functor_system(kind, variant, ndim=2, n=200, seed=kind, continuous=True), steps=4))
for integ in sorted(W.INTEGRATOR_CASES):
    print("integ", integ, go(W.integrator_system(integ, ndim=3, n=300, seed=integ), steps=8))
w = W.packer_system(ndim=3, n=200, seed=1)
print("nlcg", go(w, steps=6))
w = W.config4(shape=(6, 6, 6)); print("sol", go(w, steps=10))
