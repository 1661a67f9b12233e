"""Builds the namespace of pyparm.d2 / pyparm.d3: parm_b200.sim with NDIM fixed."""
import numpy as np

from parm_b200 import sim


class Vec:
    """Vec / Vector2d / Vector3d of the SWIG module: a float64 array of NDIM entries with the few methods scripts use."""


def populate(ns, ndim):
    def vec(*xs):
        if len(xs) == 1 and np.ndim(xs[0]) == 1:
            xs = tuple(xs[0])
        if len(xs) not in (0, ndim):
            raise ValueError("Vec of %d dimensions needs %d components" % (ndim, ndim))
        return np.array(xs if xs else (0.0,) * ndim, dtype=np.float64)

    class VecType:
        """Callable like the SWIG Vec constructor; Vec.Zero() as in Eigen."""
        def __call__(self, *xs):
            return vec(*xs)

        @staticmethod
        def Zero():
            return np.zeros(ndim)

    class OriginBox(sim.OriginBox):
        def __init__(self, L):
            sim.OriginBox.__init__(self, L, ndim)

    class AtomVec(sim.AtomVec):
        def __init__(self, N_or_masses, mass=None, device=0):
            sim.AtomVec.__init__(self, N_or_masses, mass, ndim, device)

    ns.update(NDIM=ndim, Vec=VecType(), OriginBox=OriginBox, AtomVec=AtomVec)
    # everything else keeps its reference name (sim.i:616-670 and the classes of box/trackers/interaction/collection.hpp)
    for name in dir(sim):
        if name.startswith("_") or name in ns:
            continue
        obj = getattr(sim, name)
        if isinstance(obj, type) or callable(obj):
            ns[name] = obj
