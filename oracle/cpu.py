"""TEST INFRASTRUCTURE ONLY -- ctypes front-end for the two CPU oracles.

* backend "ref":  oracle/_ref/libparm_ref{2,3}d.so -- the UNMODIFIED reference
  sources compiled against oracle/shim (see oracle/Makefile, oracle/ref_harness.cpp).
* backend "port": oracle/libparm_oracle.so -- the plain-C restatement
  (oracle/parm_oracle.c), pinned bit-for-bit against "ref" by tests/test_oracle_pin.py.

Nothing under parm_b200/ may import this module; only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() do.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KIND_LJREPULSE, KIND_REPULSION, KIND_LJATTRACTREPULSE, KIND_LJCUT = 0, 1, 2, 3
(KIND_LJATTRACTCUT, KIND_LJATTRACTFIXEDREPULSE, KIND_EISMCLACHLAN, KIND_LJISH, KIND_LJATTRACTREPULSESIGS,
 KIND_REPULSIONDRAG, KIND_LOISOHERN, KIND_LOISLIN, KIND_LOISOHERNMIN, KIND_LOISLINMIN) = range(4, 14)
VERLET, SOL = 0, 1

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def build(which=("port", "ref")):
    """Build the oracle libraries (idempotent; 'ref' needs /root/reference)."""
    targets = [t for t in which]
    subprocess.check_call(["make", "-s", "-f", os.path.join(HERE, "Makefile")] + targets)


def have(backend, ndim=3):
    return os.path.exists(_libpath(backend, ndim))


def _libpath(backend, ndim):
    if backend == "ref":
        return os.path.join(HERE, "_ref", "libparm_ref%dd.so" % ndim)
    return os.path.join(HERE, "libparm_oracle.so")


_libs = {}


def _load(backend, ndim):
    key = (backend, ndim if backend == "ref" else 0)
    if key in _libs:
        return _libs[key]
    path = _libpath(backend, ndim)
    if not os.path.exists(path):
        raise FileNotFoundError("oracle library missing: %s (run make -f oracle/Makefile)" % path)
    lib = C.CDLL(path)
    p = "ref_" if backend == "ref" else "port_"
    vp = C.c_void_p

    def sig(name, res, args):
        f = getattr(lib, p + name)
        f.restype = res
        f.argtypes = args
        return f

    api = {}
    if backend == "ref":
        api["sys_create"] = sig("sys_create", vp, [C.c_uint32, _dp, _dp, _dp, _dp])
        api["ndim"] = sig("ndim", C.c_int, [])
        api["inject_noise"] = sig("inject_noise", None, [_dp, C.c_size_t])
        api["probe_draw_order"] = sig("probe_draw_order", None, [C.POINTER(C.c_int)])
        api["noise_consumed"] = sig("noise_consumed", C.c_size_t, [])
    else:
        api["sys_create"] = sig("sys_create", vp, [C.c_int, C.c_uint32, _dp, _dp, _dp, _dp])
        api["inject_noise"] = sig("inject_noise", None, [vp, _dp, C.c_size_t])
        api["get_sol_constants"] = sig("get_sol_constants", None, [vp, _dp])
    api["sys_destroy"] = sig("sys_destroy", None, [vp])
    api["add_interaction"] = sig("add_interaction", C.c_int,
                                 [vp, C.c_int, C.c_double, _dp, _u32p, _dp, C.c_int, _u8p, C.c_int, C.c_int])
    api["add_interaction_ex"] = sig("add_interaction_ex", C.c_int,
                                    [vp, C.c_int, C.c_double, _dp, C.c_int, _u32p, _dp, _dp, C.c_int, _u8p, C.c_int, C.c_int])
    api["inter_contacts"] = sig("inter_contacts", C.c_int, [vp, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)])
    api["make_collection"] = sig("make_collection", C.c_int, [vp, C.c_int, C.c_double, C.c_double, C.c_double])
    api["make_collection_ex"] = sig("make_collection_ex", C.c_int, [vp, C.c_int, _dp, C.c_int])
    api["get_scalars"] = sig("get_scalars", None, [vp, _dp])
    api["nlcg_set"] = sig("nlcg_set", C.c_int, [vp, C.c_int, C.c_double])
    api["nlcg_get"] = sig("nlcg_get", C.c_int, [vp, _dp])
    api["nlcg_set_forces"] = sig("nlcg_set_forces", C.c_int, [vp, C.c_int, C.c_int])
    api["nlcg_reset"] = sig("nlcg_reset", C.c_int, [vp])
    api["nlcg_descend"] = sig("nlcg_descend", C.c_int, [vp])
    api["nlcg_reduce"] = sig("nlcg_reduce", C.c_double, [vp, C.c_int])
    api["get_box"] = sig("get_box", None, [vp, _dp])
    api["set_box"] = sig("set_box", None, [vp, _dp])
    ullp = C.POINTER(C.c_ulonglong)
    api["add_rsq_tracker"] = sig("add_rsq_tracker", C.c_int, [vp, ullp, C.c_int, C.c_int])
    api["add_isf_tracker"] = sig("add_isf_tracker", C.c_int, [vp, _dp, C.c_int, ullp, C.c_int, C.c_int])
    api["add_energy_tracker"] = sig("add_energy_tracker", C.c_int, [vp, C.c_uint])
    api["tracker_update"] = sig("tracker_update", None, [vp, C.c_int])
    api["tracker_reset"] = sig("tracker_reset", None, [vp, C.c_int])
    api["tracker_counts"] = sig("tracker_counts", None, [vp, C.c_int, ullp, C.c_int])
    api["rsq_read"] = sig("rsq_read", None, [vp, C.c_int, C.c_int, _dp, _dp, _dp])
    api["isf_read"] = sig("isf_read", None, [vp, C.c_int, C.c_int, _dp])
    api["energy_tracker_read"] = sig("energy_tracker_read", None, [vp, C.c_int, _dp])
    api["energy_tracker_set_U0"] = sig("energy_tracker_set_U0", None, [vp, C.c_int, C.c_int, C.c_double])
    api["update_list"] = sig("update_list", C.c_int, [vp, C.c_int, C.c_int])
    api["which"] = sig("which", C.c_uint32, [vp, C.c_int])
    api["ignore"] = sig("ignore", None, [vp, C.c_int, _u32p, _u32p, C.c_uint64])
    api["ignore_size"] = sig("ignore_size", C.c_uint32, [vp, C.c_int])
    api["numpairs"] = sig("numpairs", C.c_uint32, [vp, C.c_int])
    api["get_pairs"] = sig("get_pairs", None, [vp, C.c_int, _u32p, _u32p])
    api["set_atoms"] = sig("set_atoms", None, [vp, _dp, _dp, _dp, _dp])
    api["get_atoms"] = sig("get_atoms", None, [vp, _dp, _dp, _dp, _dp])
    api["box_diff"] = sig("box_diff", None, [vp, _dp, _dp, _dp])
    api["box_V"] = sig("box_V", C.c_double, [vp])
    api["reset_forces"] = sig("reset_forces", None, [vp])
    api["inter_set_forces"] = sig("inter_set_forces", None, [vp, C.c_int])
    api["inter_set_forces_get_pressure"] = sig("inter_set_forces_get_pressure", C.c_double, [vp, C.c_int])
    api["inter_energy"] = sig("inter_energy", C.c_double, [vp, C.c_int])
    api["inter_pressure"] = sig("inter_pressure", C.c_double, [vp, C.c_int])
    api["inter_stress"] = sig("inter_stress", None, [vp, C.c_int, _dp])
    api["timestep"] = sig("timestep", None, [vp, C.c_int])
    api["set_forces"] = sig("set_forces", None, [vp, C.c_int])
    for n in ("energy", "potential_energy", "kinetic_energy", "pressure", "virial", "degrees_of_freedom", "mass"):
        api[n] = sig(n, C.c_double, [vp])
    api["temp"] = sig("temp", C.c_double, [vp, C.c_int])
    api["reset_com_velocity"] = sig("reset_com_velocity", None, [vp])
    api["scale_velocities"] = sig("scale_velocities", None, [vp, C.c_double])
    api["scale_velocities_to_temp"] = sig("scale_velocities_to_temp", None, [vp, C.c_double, C.c_int])
    api["scale_velocities_to_energy"] = sig("scale_velocities_to_energy", None, [vp, C.c_double])
    api["com_velocity"] = sig("com_velocity", None, [vp, _dp])
    api["momentum"] = sig("momentum", None, [vp, _dp])
    api["atoms_kinetic_energy"] = sig("atoms_kinetic_energy", C.c_double, [vp, _dp])
    api["add_velocity"] = sig("add_velocity", None, [vp, _dp])
    _libs[key] = api
    return api


class CpuSystem:
    """One OriginBox + AtomVec + interactions + (optional) Collection on the CPU."""

    def __init__(self, backend, L, x, v=None, m=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.n, self.ndim = x.shape
        self.backend = backend
        self.api = _load(backend, self.ndim)
        self.L = np.ascontiguousarray(np.broadcast_to(np.asarray(L, dtype=np.float64), (self.ndim,)))
        v = np.zeros_like(x) if v is None else np.ascontiguousarray(v, dtype=np.float64)
        m = np.ones(self.n) if m is None else np.ascontiguousarray(np.broadcast_to(m, (self.n,)), dtype=np.float64)
        if backend == "ref":
            assert self.api["ndim"]() == self.ndim
            self.h = self.api["sys_create"](self.n, _d(self.L), _d(x), _d(v), _d(m))
        else:
            self.h = self.api["sys_create"](self.ndim, self.n, _d(self.L), _d(x), _d(v), _d(m))
        self.m = m
        self._keep = []

    def close(self):
        if self.h:
            if self.backend == "ref":
                self.api["inject_noise"](None, 0)
            self.api["sys_destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_interaction(self, kind, skin, params, types=None, eps_table=None, member=None, injected=False, share_nl=-1,
                        sig_table=None):
        """params: n x nper (nper <= 5), layout per kind as in include/parm_b200.h."""
        params = np.ascontiguousarray(params, dtype=np.float64)
        params = params.reshape(self.n, params.size // self.n if self.n else 3)
        t = None if types is None else np.ascontiguousarray(types, dtype=np.uint32)
        e = None if eps_table is None else np.ascontiguousarray(eps_table, dtype=np.float64)
        sg = None if sig_table is None else np.ascontiguousarray(sig_table, dtype=np.float64)
        nt = 0 if e is None else e.shape[0]
        mem = None if member is None else np.ascontiguousarray(member, dtype=np.uint8)
        r = self.api["add_interaction_ex"](
            self.h, kind, float(skin), _d(params), params.shape[1],
            None if t is None else t.ctypes.data_as(_u32p), _d(e), _d(sg), nt,
            None if mem is None else mem.ctypes.data_as(_u8p), int(injected), share_nl)
        if r < 0:
            raise RuntimeError("oracle add_interaction failed: %d" % r)
        return r

    def make_collection(self, integrator, dt, damping=0.0, T=0.0, params=None):
        """integrator: PARM_INTEG_* (include/parm_b200.h); params: the constructor arguments after dt."""
        if params is None:
            params = (damping, T)
        p = np.ascontiguousarray((dt,) + tuple(params), dtype=np.float64)
        r = self.api["make_collection_ex"](self.h, integrator, _d(p), p.size)
        if r:
            raise ValueError("oracle make_collection failed: %d" % r)

    # ---- CollectionNLCG (selectors: include/parm_b200.h PARM_NLCG_*) ----
    def nlcg_set(self, which, value):
        if self.api["nlcg_set"](self.h, which, float(value)):
            raise ValueError("not a CollectionNLCG / unknown parameter")

    def nlcg_get(self):
        out = np.zeros(16)
        if self.api["nlcg_get"](self.h, _d(out)):
            raise ValueError("not a CollectionNLCG")
        return out

    def nlcg_set_forces(self, constraints_and_a=True, setV=True):
        self.api["nlcg_set_forces"](self.h, int(constraints_and_a), int(setV))

    def nlcg_reset(self):
        self.api["nlcg_reset"](self.h)

    def nlcg_descend(self):
        self.api["nlcg_descend"](self.h)

    def nlcg_reduce(self, what):
        return self.api["nlcg_reduce"](self.h, what)

    def get_box(self):
        out = np.zeros(self.ndim)
        self.api["get_box"](self.h, _d(out))
        return out

    def set_box(self, L):
        L = np.ascontiguousarray(np.broadcast_to(np.asarray(L, dtype=np.float64), (self.ndim,)))
        self.api["set_box"](self.h, _d(L))

    # ---- statistics trackers (constraints.hpp:260-414); added to the collection like add_tracker() ----
    def add_rsq_tracker(self, ns, usecom=True):
        ns = np.ascontiguousarray(ns, dtype=np.uint64)
        self._stat_meta = getattr(self, "_stat_meta", {})
        t = self.api["add_rsq_tracker"](self.h, ns.ctypes.data_as(C.POINTER(C.c_ulonglong)), ns.size, int(usecom))
        self._stat_meta[t] = (len(ns), 0)
        return t

    def add_isf_tracker(self, ks, ns, usecom=False):
        ns = np.ascontiguousarray(ns, dtype=np.uint64)
        ks = np.ascontiguousarray(ks, dtype=np.float64)
        self._stat_meta = getattr(self, "_stat_meta", {})
        t = self.api["add_isf_tracker"](self.h, _d(ks), ks.size, ns.ctypes.data_as(C.POINTER(C.c_ulonglong)), ns.size, int(usecom))
        self._stat_meta[t] = (len(ns), len(ks))
        return t

    def add_energy_tracker(self, n_skip=1):
        return self.api["add_energy_tracker"](self.h, int(n_skip))

    def tracker_update(self, t):
        self.api["tracker_update"](self.h, t)

    def tracker_reset(self, t):
        self.api["tracker_reset"](self.h, t)

    def tracker_counts(self, t):
        n = self._stat_meta[t][0]
        out = (C.c_ulonglong * max(n, 1))()
        self.api["tracker_counts"](self.h, t, out, n)
        return [int(out[k]) for k in range(n)]

    def rsq_read(self, t, single):
        """(xyz2, xyz4, r4) means of lag `single`."""
        a, b, c = np.zeros((self.n, self.ndim)), np.zeros((self.n, self.ndim)), np.zeros(self.n)
        self.api["rsq_read"](self.h, t, single, _d(a), _d(b), _d(c))
        return a, b, c

    def isf_read(self, t, single):
        """ISFxyz of lag `single` as a complex array (nks, n, ndim)."""
        nks = self._stat_meta[t][1]
        out = np.zeros((nks, self.n, self.ndim, 2))
        self.api["isf_read"](self.h, t, single, _d(out))
        return out[..., 0] + 1j * out[..., 1]

    def energy_tracker_read(self, t):
        """n(), E(), U(), K(), E_squared_mean(), U_squared_mean(), K_squared_mean(), get_U0()."""
        out = np.zeros(8)
        self.api["energy_tracker_read"](self.h, t, _d(out))
        return out

    def energy_tracker_set_U0(self, t, U0=None):
        self.api["energy_tracker_set_U0"](self.h, t, int(U0 is None), 0.0 if U0 is None else float(U0))

    def get_scalars(self):
        """(xi, lns) of CollectionNoseHoover."""
        out = np.zeros(2)
        self.api["get_scalars"](self.h, _d(out))
        return out

    def update_list(self, force=True, nl=0):
        return bool(self.api["update_list"](self.h, nl, int(force)))

    def which(self, nl=0):
        return self.api["which"](self.h, nl)

    def ignore(self, a, b, nl=0):
        """NeighborList::ignore for equal-length arrays of AtomVec indices (trackers.hpp:190-193)."""
        a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.uint32)
        b = np.ascontiguousarray(np.atleast_1d(b), dtype=np.uint32)
        self.api["ignore"](self.h, nl, a.ctypes.data_as(_u32p), b.ctypes.data_as(_u32p), a.size)

    def ignore_size(self, nl=0):
        return self.api["ignore_size"](self.h, nl)

    def pairs(self, nl=0):
        """Pairs in the reference's own order: (first, last) = (later atom i, earlier atom j<i)."""
        n = self.api["numpairs"](self.h, nl)
        a = np.empty(n, np.uint32)
        b = np.empty(n, np.uint32)
        if n:
            self.api["get_pairs"](self.h, nl, a.ctypes.data_as(_u32p), b.ctypes.data_as(_u32p))
        return a, b

    def set_atoms(self, x=None, v=None, a=None, f=None):
        arrs = [None if q is None else np.ascontiguousarray(q, dtype=np.float64) for q in (x, v, a, f)]
        self.api["set_atoms"](self.h, *[_d(q) for q in arrs])

    def get_atoms(self):
        out = [np.empty((self.n, self.ndim)) for _ in range(4)]
        self.api["get_atoms"](self.h, *[_d(q) for q in out])
        return tuple(out)

    def box_diff(self, r1, r2):
        r1 = np.ascontiguousarray(r1, dtype=np.float64)
        r2 = np.ascontiguousarray(r2, dtype=np.float64)
        out = np.empty(self.ndim)
        self.api["box_diff"](self.h, _d(r1), _d(r2), _d(out))
        return out

    def forces(self, k=0):
        """reset_forces + Interaction::set_forces -> per-atom f."""
        self.api["reset_forces"](self.h)
        self.api["inter_set_forces"](self.h, k)
        return self.get_atoms()[3]

    def forces_and_pressure(self, k=0):
        self.api["reset_forces"](self.h)
        p = self.api["inter_set_forces_get_pressure"](self.h, k)
        return self.get_atoms()[3], p

    def inter_energy(self, k=0):
        return self.api["inter_energy"](self.h, k)

    def inter_pressure(self, k=0):
        return self.api["inter_pressure"](self.h, k)

    def inter_stress(self, k=0):
        out = np.empty((self.ndim, self.ndim))
        self.api["inter_stress"](self.h, k, _d(out))
        return out

    def inter_contacts(self, k=0):
        """(contacts, overlaps) of NListed<A,P> (interaction.hpp:2126-2151)."""
        c, o = C.c_ulonglong(0), C.c_ulonglong(0)
        if self.api["inter_contacts"](self.h, k, C.byref(c), C.byref(o)):
            raise RuntimeError("oracle inter_contacts failed")
        return c.value, o.value

    def timestep(self, n=1):
        self.api["timestep"](self.h, n)

    def set_forces(self, constraints_and_a=True):
        self.api["set_forces"](self.h, int(constraints_and_a))

    def inject_noise(self, z):
        """z: (steps, n_mobile, 2, ndim) standard normals: [.,.,0,:] -> x1, [.,.,1,:] -> x2 of
        BivariateGauss::gen_vecs (vecrand.cpp:73-85)."""
        z = np.ascontiguousarray(z, dtype=np.float64)
        if z.shape[-1] != self.ndim:  # the reference's draw order is undone along the last axis
            raise ValueError("inject_noise: last axis must be ndim = %d, got shape %r" % (self.ndim, z.shape))
        if self.backend == "ref":
            order = (C.c_int * self.ndim)()
            self.api["probe_draw_order"](order)
            order = list(order)  # order[c] = draw index that lands in component c
            draws = np.empty_like(z)
            for c, k in enumerate(order):
                draws[..., k] = z[..., c]
            z = np.ascontiguousarray(draws)
            self._keep.append(z)
            self.api["inject_noise"](_d(z), z.size)
        else:
            self._keep.append(z)
            self.api["inject_noise"](self.h, _d(z), z.size)

    def __getattr__(self, name):
        if name in ("energy", "potential_energy", "kinetic_energy", "pressure", "virial", "degrees_of_freedom", "mass"):
            return lambda: self.api[name](self.h)
        raise AttributeError(name)

    def temp(self, minuscomv=True):
        return self.api["temp"](self.h, int(minuscomv))

    def reset_com_velocity(self):
        self.api["reset_com_velocity"](self.h)

    def scale_velocities(self, s):
        self.api["scale_velocities"](self.h, s)

    def scale_velocities_to_temp(self, T, minuscomv=True):
        self.api["scale_velocities_to_temp"](self.h, T, int(minuscomv))

    def scale_velocities_to_energy(self, E):
        self.api["scale_velocities_to_energy"](self.h, E)

    def com_velocity(self):
        out = np.empty(self.ndim)
        self.api["com_velocity"](self.h, _d(out))
        return out

    def momentum(self):
        out = np.empty(self.ndim)
        self.api["momentum"](self.h, _d(out))
        return out

    def atoms_kinetic_energy(self, v0):
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        return self.api["atoms_kinetic_energy"](self.h, _d(v0))

    def add_velocity(self, dv):
        dv = np.ascontiguousarray(dv, dtype=np.float64)
        self.api["add_velocity"](self.h, _d(dv))
