"""Host-side mirror of ParM's public classes over the parm_b200 C ABI.

Same class names, argument meaning and error behaviour as the reference's SWIG
modules pyparm.d2 / pyparm.d3 (src/sim.i:616-670) for the hot path:

  OriginBox            box.hpp:97-158
  AtomVec / Atom       box.hpp:234-249, 441-479   (+ AtomGroup reductions box.cpp:239-260, 401-431)
  NeighborList         trackers.hpp:157-214
  LJRepulse            NListed<EpsSigAtom, LJRepulsePair>          sim.i:621
  Repulsion            NListed<EpsSigExpAtom, RepulsionPair>       sim.i:636
  LJAttractRepulse     NListed<IEpsSigCutAtom, LJAttractRepulsePair>  sim.i:629
  LJCut                NListed<EpsSigCutAtom, LennardJonesCutPair> (LJatoms.cpp:37-38; no SWIG name)
  CollectionVerlet     collection.hpp:360-374
  CollectionSol        collection.hpp:205-263

All numerical work happens in libparm_b200.so on the GPU; this module only keeps the
host mirror of `Atom` (AoS, identical layout to the reference's struct) coherent with the
device copy using coarse dirty flags (SURVEY 8b "Mutation model").
"""
import numpy as np

from . import capi
from .capi import C, call, dp, u32p, u8p, u64p


def _dptr(a):
    return None if a is None else a.ctypes.data_as(dp)


def atom_dtype(ndim):
    """struct Atom {Vec x, v, a, f; flt m;} (box.hpp:234-249): 104 bytes in 3-D, 72 in 2-D."""
    return np.dtype([("x", np.float64, (ndim,)), ("v", np.float64, (ndim,)), ("a", np.float64, (ndim,)),
                     ("f", np.float64, (ndim,)), ("m", np.float64)])


class OriginBox:
    """Rectilinear periodic box (box.hpp:97-158)."""

    def __init__(self, L, ndim=None):
        L = np.atleast_1d(np.asarray(L, dtype=np.float64))
        if L.size == 1:
            L = np.full(3 if ndim is None else ndim, float(L[0]))
        self.boxsize = np.ascontiguousarray(L)
        self.ndim = self.boxsize.size
        if self.ndim not in (2, 3):
            raise ValueError("OriginBox: NDIM must be 2 or 3")
        self._ctxs = []

    def box_shape(self):
        return self.boxsize.copy()

    def V(self):  # box.hpp:122,127
        b = self.boxsize
        return float(b[0] * b[1] * b[2]) if self.ndim == 3 else float(b[0] * b[1])

    def L(self):  # box.hpp:123,128
        b = self.boxsize
        return float((b[0] + b[1] + b[2]) / 3.0) if self.ndim == 3 else float((b[0] + b[1]) / 2.0)

    def _attach(self, atoms):
        if atoms not in self._ctxs:
            self._ctxs.append(atoms)
            call("parm_set_box", atoms._h, _dptr(self.boxsize))

    def resize_to(self, newsize):  # box.cpp:21-25 (does not move atoms)
        self.boxsize = np.ascontiguousarray(np.broadcast_to(np.asarray(newsize, dtype=np.float64), (self.ndim,)))
        for a in self._ctxs:
            call("parm_set_box", a._h, _dptr(self.boxsize))
        return self.V()

    def resize(self, factor):  # box.cpp:3-6
        return self.resize_to(self.boxsize * factor)

    def resize_to_V(self, newV):  # box.cpp:34-38
        return self.resize((newV / self.V()) ** (1.0 / self.ndim))

    def resize_to_L(self, newL):  # box.cpp:46-50
        return self.resize(newL / self.V() ** (1.0 / self.ndim))

    def _pull(self, atoms):
        """The device changed the box (CollectionNLCG::stepx): refresh this object and the other contexts."""
        L = np.zeros(self.ndim)
        call("parm_get_box", atoms._h, _dptr(L))
        self.boxsize = L
        for a in self._ctxs:
            if a is not atoms:
                call("parm_set_box", a._h, _dptr(self.boxsize))

    def diff(self, r1, r2, atoms=None):
        """OriginBox::diff evaluated ON THE DEVICE (box.hpp:103). r1, r2: (..., ndim)."""
        if atoms is None:
            if not self._ctxs:
                raise capi.ParmInvalid("OriginBox.diff needs an AtomVec context")
            atoms = self._ctxs[0]
        r1 = np.ascontiguousarray(r1, dtype=np.float64)
        r2 = np.ascontiguousarray(r2, dtype=np.float64)
        out = np.empty_like(r1)
        call("parm_box_diff", atoms._h, r1.size // self.ndim, _dptr(r1), _dptr(r2), _dptr(out))
        return out


class Atom:
    """Proxy for one `Atom&` of an AtomVec (reads/writes go to the host mirror)."""
    __slots__ = ("_av", "_i")

    def __init__(self, av, i):
        object.__setattr__(self, "_av", av)
        object.__setattr__(self, "_i", i)

    def __getattr__(self, k):
        if k in ("x", "v", "a", "f", "m"):
            return self._av._host()[k][self._i]
        raise AttributeError(k)

    def __setattr__(self, k, val):
        if k not in ("x", "v", "a", "f", "m"):
            raise AttributeError(k)
        self._av._host()[k][self._i] = val

    def n(self):  # iterating an AtomVec yields objects that also serve as AtomID (sim.i AtomVec.__iter__ -> get_id)
        return self._i


class AtomID:
    def __init__(self, av, n):
        self.atoms, self._n = av, n

    def n(self):
        return self._n


class AtomVec:
    """AtomVec(N, mass) / AtomVec(masses) (box.hpp:441-479). Owns the device context."""

    def __init__(self, N_or_masses, mass=None, ndim=3, device=0):
        if mass is None:
            masses = np.asarray(N_or_masses, dtype=np.float64).ravel()
        else:
            masses = np.full(int(N_or_masses), float(mass))
        self.ndim = ndim
        self.n = masses.size
        self.atoms = np.zeros(self.n, dtype=atom_dtype(ndim))  # zero(): x=v=a=f=0 (box.hpp:445-452)
        self.atoms["m"] = masses
        h = C.c_void_p()
        call("parm_ctx_create", ndim, self.n, device, C.byref(h))
        self._h = h
        self._registered = False
        if self.n:
            try:
                call("parm_host_register", self.atoms.ctypes.data, self.atoms.nbytes)
                self._registered = True
            except capi.ParmError:
                pass
        self._host_dirty = True   # host mirror holds data the device has not seen
        self._dev_newer = False   # device holds data the host mirror has not seen

    # -- coherence ---------------------------------------------------------------
    def _field_ptrs(self):
        base = self.atoms.ctypes.data
        vec = self.ndim * 8
        return [C.cast(base + k * vec, dp) for k in range(5)]

    def _host(self):
        """Host mirror, current, and assumed modified by the caller (a mutable view escapes)."""
        self.sync_to_host()
        self._host_dirty = True
        return self.atoms

    def sync_to_host(self):
        if self._dev_newer and self.n:
            p = self._field_ptrs()
            call("parm_download_atoms", self._h, capi.ALL, p[0], p[1], p[2], p[3], p[4], self.atoms.itemsize, self.atoms.itemsize)
        self._dev_newer = False

    def sync_to_device(self):
        if self._host_dirty and self.n:
            p = self._field_ptrs()
            call("parm_upload_atoms", self._h, capi.ALL, p[0], p[1], p[2], p[3], p[4], self.atoms.itemsize, self.atoms.itemsize)
        self._host_dirty = False

    def _device_op(self, modifies=True):
        self.sync_to_device()
        if modifies:
            self._dev_newer = True

    # -- element access ----------------------------------------------------------------
    def __len__(self):
        return self.n

    def size(self):
        return self.n

    def __getitem__(self, i):
        if not -self.n <= i < self.n:
            raise IndexError(i)
        return Atom(self, i % self.n)

    def __iter__(self):
        return (Atom(self, i) for i in range(self.n))

    def get_id(self, n):
        return AtomID(self, n)

    # -- asynchronous trajectory frames (parm_snapshot_begin / _wait; pyparm/xyzfile.py, LJatoms.cpp:130-158) -----------
    def snapshot_begin(self, velocities=True):
        """Start copying the current x (and v) to the host on a second stream; timestep() calls made before
        snapshot_wait() overlap the copy."""
        self._device_op(modifies=False)
        call("parm_snapshot_begin", self._h, capi.X | (capi.V if velocities else 0))
        self._snap_v = bool(velocities)

    def snapshot_wait(self):
        """(x, v) of the frame started by snapshot_begin(), arrays of shape (n, ndim); v is None without velocities."""
        x = np.empty((self.n, self.ndim))
        v = np.empty((self.n, self.ndim)) if self._snap_v else None
        call("parm_snapshot_wait", self._h, _dptr(x), _dptr(v) if v is not None else None)
        return x, v

    x = property(lambda self: self._host()["x"])
    v = property(lambda self: self._host()["v"])
    a = property(lambda self: self._host()["a"])
    f = property(lambda self: self._host()["f"])
    m = property(lambda self: self._host()["m"])

    def peek(self, field):
        """Read-only copy of a field without marking the host mirror modified."""
        self.sync_to_host()
        return self.atoms[field].copy()

    # -- AtomGroup reductions, computed on the device (box.cpp:239-260, 401-431) ----------
    def _reduce(self, what, v0=None, nout=1):
        self._device_op(modifies=False)
        out = np.zeros(4)
        z = None if v0 is None else np.ascontiguousarray(v0, dtype=np.float64)
        call("parm_reduce", self._h, what, _dptr(z), _dptr(out))
        return float(out[0]) if nout == 1 else out[:nout].copy()

    def mass(self):
        return self._reduce(capi.RED_MASS)

    def momentum(self):
        return self._reduce(capi.RED_MOMENTUM, nout=self.ndim)

    def com(self):
        return self._reduce(capi.RED_COM, nout=self.ndim)

    def com_force(self):
        return self._reduce(capi.RED_COMFORCE, nout=self.ndim)

    def com_velocity(self):
        return self.momentum() / self.mass()

    def kinetic_energy(self, originvelocity=None):
        return self._reduce(capi.RED_KE, originvelocity)

    def add_velocity(self, dv):
        self._device_op()
        dv = np.ascontiguousarray(dv, dtype=np.float64)
        call("parm_add_velocity", self._h, _dptr(dv))

    def reset_com_velocity(self):
        self.add_velocity(-self.com_velocity())

    def reset_forces(self):
        self._device_op()
        call("parm_reset_forces", self._h)

    def close(self):
        if self._h:
            if self._registered:
                try:
                    call("parm_host_unregister", self.atoms.ctypes.data)
                except Exception:
                    pass
            capi.lib().parm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NeighborList:
    """NeighborList(box, atoms, skin) (trackers.hpp:157-214)."""

    def __init__(self, box, atoms, skin):
        self.box, self.atoms, self.skin = box, atoms, float(skin)
        box._attach(atoms)
        h = C.c_void_p()
        call("parm_nlist_create", atoms._h, self.skin, C.byref(h))
        self._h = h
        self._diam = np.full(atoms.n, -1.0)
        self._diam_dirty = False

    def add(self, atom_id, diameter):
        i = atom_id.n() if hasattr(atom_id, "n") else int(atom_id)
        if self._diam[i] >= 0:  # SubGroup::add (box.hpp:495-501)
            raise ValueError("Cannot add AtomID to SubGroup: it already exists.")
        self._diam[i] = diameter
        self._diam_dirty = True

    def _flush(self):
        if self._diam_dirty:
            call("parm_nlist_set_diameters", self._h, _dptr(self._diam))
            self._diam_dirty = False

    def _set_diameters_from_inter(self, diam):
        self._diam = np.array(diam, dtype=np.float64)
        self._diam_dirty = False

    def update_list(self, force=True):
        self._flush()
        self.atoms._device_op()  # a rebuild re-orders the device arrays but not their AtomVec-indexed content
        r = C.c_int(0)
        call("parm_nlist_update", self._h, int(force), C.byref(r))
        return bool(r.value)

    def update(self, box=None):
        return self.update_list(False)

    def which(self):
        u = C.c_uint32(0)
        call("parm_nlist_which", self._h, C.byref(u))
        return u.value

    def numpairs(self):
        u = C.c_uint64(0)
        call("parm_nlist_numpairs", self._h, C.byref(u))
        return u.value

    def size(self):
        return int((self._diam >= 0).sum())

    def ignore(self, a, b):
        """ignore(AtomID a, AtomID b) (trackers.hpp:190-193); a and b may also be equal-length index arrays."""
        ia = np.atleast_1d(np.asarray([q.n() if hasattr(q, "n") else q for q in np.atleast_1d(a)], dtype=np.uint32))
        ib = np.atleast_1d(np.asarray([q.n() if hasattr(q, "n") else q for q in np.atleast_1d(b)], dtype=np.uint32))
        if ia.shape != ib.shape:
            raise ValueError("ignore: a and b must have the same length")
        call("parm_nlist_ignore", self._h, ia.ctypes.data_as(u32p), ib.ctypes.data_as(u32p), ia.size)

    def ignore_size(self):
        u = C.c_uint64(0)
        call("parm_nlist_ignore_size", self._h, C.byref(u))
        return u.value

    def pairs(self):
        """curpairs as two uint32 arrays (first, last) in the reference's order (trackers.cpp:59-68)."""
        n = self.numpairs()
        a = np.empty(n, np.uint32)
        b = np.empty(n, np.uint32)
        if n:
            call("parm_nlist_download_pairs", self._h, a.ctypes.data_as(u32p), b.ctypes.data_as(u32p), n)
        return a, b

    def __iter__(self):
        a, b = self.pairs()
        return iter(zip(a.tolist(), b.tolist()))

    def stats(self):
        mean = C.c_double(0)
        mx = C.c_uint32(0)
        call("parm_nlist_stats", self._h, C.byref(mean), C.byref(mx))
        return mean.value, mx.value

    def tile_stats(self):
        """(active, chunks, max_tile_atoms, wide_chunks) of the cell-tile pair kernel after the last rebuild."""
        act = C.c_int(0)
        nch, mt, wide = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        call("parm_nlist_tile_stats", self._h, C.byref(act), C.byref(nch), C.byref(mt), C.byref(wide))
        return bool(act.value), nch.value, mt.value, wide.value


class Grid:
    """Grid (trackers.hpp:227-309): cells no narrower than any atom; the atoms of the same and of the neighbouring cells
    are an atom's candidate neighbours. get_loc (trackers.cpp:192-219) runs on the device for all atoms at once
    (parm_grid_locs); the per-cell lists are assembled here from the 4-byte cell indices."""

    def __init__(self, box, atoms, width_or_minwidth=1, goalwidth=None):
        self.box, self.atoms = box, atoms
        nd = atoms.ndim
        self.minwidth, self.goalwidth = -1.0, -1.0
        if goalwidth is not None:  # Grid(box, atoms, minwidth, goalwidth)
            self.minwidth, self.goalwidth = float(width_or_minwidth), float(goalwidth)
            self.widths = [1] * nd
            self.optimize_widths()
        elif np.ndim(width_or_minwidth) == 0:
            self.widths = [int(width_or_minwidth)] * nd
        else:
            self.widths = [int(w) for w in width_or_minwidth]
        self.gridlocs = []

    def optimize_widths(self):  # trackers.cpp:168-190
        if self.minwidth <= 0:
            return
        nd = self.atoms.ndim
        wpa = (self.box.V() * self.goalwidth / self.atoms.size()) ** (1.0 / nd)
        wpa = max(wpa, self.minwidth)
        self.widths = [int(np.floor(b / wpa)) for b in self.box.box_shape()]
        if min(self.widths) < 3:
            self.widths = [1] * nd

    def numcells(self, i=None):
        return int(np.prod(self.widths)) if i is None else self.widths[i]

    def make_grid(self):
        self.box._attach(self.atoms)
        self.atoms._device_op(modifies=False)
        w = np.ascontiguousarray(self.widths + [1] * (3 - len(self.widths)), dtype=np.uint32)
        loc = np.empty(self.atoms.n, np.uint32)
        call("parm_grid_locs", self.atoms._h, w.ctypes.data_as(capi.u32p), loc.ctypes.data_as(capi.u32p))
        self.locs = loc
        order = np.argsort(loc, kind="stable")
        bounds = np.searchsorted(loc[order], np.arange(self.numcells() + 1))
        self.gridlocs = [order[bounds[c]:bounds[c + 1]].tolist() for c in range(self.numcells())]  # sets of AtomID, by index

    def get_loc(self, v, bsize=None):
        bsize = self.box.box_shape() if bsize is None else np.asarray(bsize, dtype=np.float64)
        v = np.asarray(v, dtype=np.float64)
        import math
        r = np.array([math.remainder(v[d] - bsize[d] / 2.0, bsize[d]) for d in range(len(bsize))]) + bsize / 2.0
        k = [int(np.floor(r[d] * self.widths[d] / bsize[d])) for d in range(len(bsize))]
        k = [0 if k[d] == self.widths[d] else k[d] for d in range(len(bsize))]
        return (k[2] * self.widths[1] + k[1]) * self.widths[0] + k[0] if len(bsize) == 3 else k[1] * self.widths[0] + k[0]

    def neighbors(self, i):  # trackers.cpp:125-166 (the cell itself included; duplicates when an axis has < 3 cells)
        w = self.widths
        if len(w) == 2:
            x, y = i % w[0], i // w[0]
            return [((y + dy) % w[1]) * w[0] + (x + dx) % w[0] for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        x, y, z = i % w[0], (i // w[0]) % w[1], i // (w[0] * w[1])
        return [(((z + dz) % w[2]) * w[1] + (y + dy) % w[1]) * w[0] + (x + dx) % w[0]
                for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]

    def all_pairs(self, a=None):
        """all_pairs(): every (i, j), j < i ... of neighbouring cells, each once; all_pairs(n): the candidates of atom n."""
        if a is not None:
            n = a.n() if hasattr(a, "n") else int(a)
            out = []
            for c in dict.fromkeys(self.neighbors(int(self.locs[n]))):
                out.extend(j for j in self.gridlocs[c] if j != n)
            return out
        pairs = set()
        for c, members in enumerate(self.gridlocs):
            if not members:
                continue
            for c2 in dict.fromkeys(self.neighbors(c)):
                for i in members:
                    for j in self.gridlocs[c2]:
                        if j < i:
                            pairs.add((i, j))
        return sorted(pairs)


# ---- per-atom parameter structs; .p follows the table in include/parm_b200.h (parm_inter_set_params_ex) ----
class _PairAtom:
    type = 0

    def _set(self, a, *p):
        self.id = a
        self.p = tuple(p) + (0.0,) * (5 - len(p))


class EpsSigAtom(_PairAtom):  # interaction.hpp:857-865
    def __init__(self, a, epsilon, sigma):
        self._set(a, epsilon, sigma)


class EpsSigExpAtom(_PairAtom):  # interaction.hpp:1454-1465
    def __init__(self, a, eps, sigma, exponent):
        self._set(a, eps, sigma, exponent)


class EpsSigCutAtom(_PairAtom):  # interaction.hpp:897-905
    def __init__(self, a, epsilon, sigma, cut):
        self._set(a, epsilon, sigma, cut)


class IEpsSigCutAtom(_PairAtom):  # interaction.hpp:989-1018
    def __init__(self, a, epsilons, indx, sigma, cut):
        self._set(a, 0.0, sigma, cut)
        self.type = int(indx)
        self.epsilons = list(epsilons)


class IEpsISigCutAtom(_PairAtom):  # interaction.hpp:911-960
    def __init__(self, a, epsilons, sigmas, indx, cut):
        assert len(epsilons) == len(sigmas)
        self._set(a, 0.0, 0.0, cut)
        self.type = int(indx)
        self.epsilons, self.sigmas = list(epsilons), list(sigmas)


class IEpsISigExpAtom(_PairAtom):  # interaction.hpp:1477-1524
    def __init__(self, a, epsilons, sigmas, indx, exponent=2.5):
        assert len(epsilons) == len(sigmas)
        self._set(a, 0.0, 0.0, exponent)
        self.type = int(indx)
        self.epsilons, self.sigmas = list(epsilons), list(sigmas)


class IEpsRepsSigCutAtom(_PairAtom):  # interaction.hpp:1307-1341
    def __init__(self, a, epsilons, repeps, sigma, indx, cut):
        self._set(a, 0.0, sigma, cut, repeps)
        self.type = int(indx)
        self.epsilons = list(epsilons)


class IEpsRepsSigExpCutAtom(_PairAtom):  # interaction.hpp:1051-1093
    def __init__(self, a, epsilons, repeps, sigma, n, indx, cut):
        self._set(a, 0.0, sigma, cut, repeps, n)
        self.type = int(indx)
        self.epsilons = list(epsilons)


class EisMclachlanAtom(_PairAtom):  # interaction.hpp:1415-1423
    def __init__(self, a, dist, sigmai):
        self._set(a, sigmai, dist)


class EpsEpsSigSigCutAtom(_PairAtom):  # interaction.hpp:1149-1169
    def __init__(self, a, eps_r, eps_a, sigma_r, sigma_a, cut):
        self._set(a, eps_r, sigma_r, cut, eps_a, sigma_a)


class EpsSigExpDragAtom(_PairAtom):  # interaction.hpp:1598-1611
    def __init__(self, a, eps, sigma, gamma, exponent=2.5):
        self._set(a, eps, sigma, exponent, gamma)


class LoisOhernAtom(_PairAtom):  # interaction.hpp:1679-1691
    def __init__(self, a, eps, sigma, C, l):
        self._set(a, eps, sigma, C, l)


class LoisLinAtom(_PairAtom):  # interaction.hpp:1764-1780
    def __init__(self, a, eps, sigma, depth, width):
        self._set(a, eps, sigma, (depth / width) if width > 0 else 0.0, width)


def _max_sizes(kind, p, types, sig_table):
    """A::max_size() per atom (what NListed::add hands to NeighborList::add, interaction.hpp:1906-1910)."""
    sigma = p[:, 1] if sig_table is None else sig_table.max(axis=1)[types]
    if kind in (capi.PAIR_LJREPULSE, capi.PAIR_REPULSION, capi.PAIR_REPULSIONDRAG):
        return sigma
    if kind == capi.PAIR_EISMCLACHLAN:
        return p[:, 1]
    if kind == capi.PAIR_LJATTRACTREPULSESIGS:
        return p[:, 1] + p[:, 4] * (p[:, 2] - 1)
    if kind in (capi.PAIR_LOISOHERN, capi.PAIR_LOISOHERNMIN):
        return p[:, 1] * (1 + p[:, 2] + p[:, 3])
    if kind in (capi.PAIR_LOISLIN, capi.PAIR_LOISLINMIN):
        return p[:, 1] * (1 + p[:, 3])
    return sigma * p[:, 2]


class _NListed:
    kind = None

    def __init__(self, *args):
        # NListed(box, atoms, skin) or NListed(atoms, neighbors)  (interaction.hpp:1903-1914)
        if len(args) == 3:
            box, atoms, skin = args
            self.neighbors = NeighborList(box, atoms, skin)
        elif len(args) == 2:
            atoms, self.neighbors = args
        else:
            raise TypeError("NListed(box, atoms, skin) or NListed(atoms, neighbors)")
        self.atoms = atoms
        h = C.c_void_p()
        call("parm_inter_create", atoms._h, self.neighbors._h, self.kind, C.byref(h))
        self._h = h
        n = atoms.n
        self._params = np.zeros((n, 5))
        self._types = np.zeros(n, np.uint32)
        self._member = np.zeros(n, np.uint8)
        self._eps_rows, self._sig_rows = {}, {}
        self._eps_table = self._sig_table = None
        self._dirty = False

    def add(self, atm):
        i = atm.id.n() if hasattr(atm.id, "n") else int(atm.id)
        if self._member[i]:
            raise ValueError("Cannot add AtomID to SubGroup: it already exists.")
        self._params[i] = atm.p
        self._types[i] = atm.type
        self._member[i] = 1
        if hasattr(atm, "epsilons"):
            self._eps_rows[atm.type] = atm.epsilons
        if hasattr(atm, "sigmas"):
            self._sig_rows[atm.type] = atm.sigmas
        self._dirty = True

    def add_many(self, params, types=None, eps_table=None, member=None, sig_table=None):
        """Bulk form of add(): params (n, nper <= 5) as in include/parm_b200.h."""
        n = self.atoms.n
        params = np.asarray(params, dtype=np.float64)
        params = params.reshape(n, params.size // n if n else 5)
        self._params = np.zeros((n, 5))
        self._params[:, :params.shape[1]] = params
        self._types = np.zeros(n, np.uint32) if types is None else np.ascontiguousarray(types, dtype=np.uint32)
        self._member = np.ones(n, np.uint8) if member is None else np.ascontiguousarray(member, dtype=np.uint8)
        self._eps_table = None if eps_table is None else np.ascontiguousarray(eps_table, dtype=np.float64)
        self._sig_table = None if sig_table is None else np.ascontiguousarray(sig_table, dtype=np.float64)
        self._dirty = True

    @staticmethod
    def _table(tab, rows):
        if tab is None and rows:
            nt = max(max(len(r) for r in rows.values()), max(rows) + 1)
            tab = np.zeros((nt, nt))
            for t, row in rows.items():
                tab[t, :len(row)] = row
        return tab

    def _flush(self):
        if not self._dirty:
            return
        tab = self._table(self._eps_table, self._eps_rows)
        stab = self._table(self._sig_table, self._sig_rows)
        nt = 0 if tab is None else tab.shape[0]
        diam = np.where(self._member > 0, _max_sizes(self.kind, self._params, self._types, stab), -1.0)
        shared = self.neighbors._diam
        # NListed::add forwards max_size() to the (possibly shared) NeighborList
        newdiam = np.where(self._member > 0, diam, shared)
        call("parm_inter_set_params_ex", self._h, _dptr(self._params), 5, self._types.ctypes.data_as(u32p), _dptr(tab),
             _dptr(stab), nt, self._member.ctypes.data_as(u8p), 0)
        call("parm_nlist_set_diameters", self.neighbors._h, _dptr(np.ascontiguousarray(newdiam)))
        self.neighbors._set_diameters_from_inter(newdiam)
        self._dirty = False

    def neighbor_list(self):
        self._flush()
        return self.neighbors

    def _ready(self, modifies):
        self._flush()
        self.neighbors._flush()
        self.atoms._device_op(modifies)

    def energy(self, box=None):
        self._ready(False)
        e = C.c_double(0)
        call("parm_inter_energy", self._h, C.byref(e))
        return e.value

    def pressure(self, box=None):
        self._ready(False)
        e = C.c_double(0)
        call("parm_inter_pressure", self._h, C.byref(e))
        return e.value

    def stress(self, box=None):
        self._ready(False)
        D = self.atoms.ndim
        out = np.zeros((D, D))
        call("parm_inter_stress", self._h, _dptr(out))
        return out

    def set_forces(self, box=None):
        self._ready(True)
        call("parm_inter_set_forces", self._h, 0, None)

    def set_forces_get_pressure(self, box=None):
        self._ready(True)
        out = np.zeros(1)
        call("parm_inter_set_forces", self._h, capi.WANT_VIRIAL, _dptr(out))
        return float(out[0])

    def set_forces_get_stress(self, box=None):
        self._ready(True)
        D = self.atoms.ndim
        out = np.zeros((D, D))
        call("parm_inter_set_forces", self._h, capi.WANT_STRESS, _dptr(out))
        return out

    def contacts(self, box=None):
        self._ready(False)
        a, b = C.c_uint64(0), C.c_uint64(0)
        call("parm_inter_contacts", self._h, C.byref(a), C.byref(b))
        return a.value

    def overlaps(self, box=None):
        self._ready(False)
        a, b = C.c_uint64(0), C.c_uint64(0)
        call("parm_inter_contacts", self._h, C.byref(a), C.byref(b))
        return b.value


# the NListed instantiations of sim.i:621-643, under the same names
class LJRepulse(_NListed):  # NListed<EpsSigAtom, LJRepulsePair>
    kind = capi.PAIR_LJREPULSE


LJRepulsive = LJRepulse  # planned rename, src/namereplacements.txt:181


class Repulsion(_NListed):  # NListed<EpsSigExpAtom, RepulsionPair>
    kind = capi.PAIR_REPULSION


class RepulsionII(_NListed):  # NListed<IEpsISigExpAtom, RepulsionPair>
    kind = capi.PAIR_REPULSION


class LJAttractRepulse(_NListed):  # NListed<IEpsSigCutAtom, LJAttractRepulsePair>
    kind = capi.PAIR_LJATTRACTREPULSE


class LJCut(_NListed):  # NListed<EpsSigCutAtom, LennardJonesCutPair>
    kind = capi.PAIR_LJCUT


class LJIICut(_NListed):  # NListed<IEpsISigCutAtom, LennardJonesCutPair>
    kind = capi.PAIR_LJCUT


class LJAttractCut(_NListed):  # NListed<EpsSigCutAtom, LJAttractCutPair>
    kind = capi.PAIR_LJATTRACTCUT


class LJAttractICut(_NListed):  # NListed<IEpsSigCutAtom, LJAttractCutPair>
    kind = capi.PAIR_LJATTRACTCUT


class LJAttractIICut(_NListed):  # NListed<IEpsISigCutAtom, LJAttractCutPair>
    kind = capi.PAIR_LJATTRACTCUT


class LJAttractFixedRepulse(_NListed):  # NListed<IEpsRepsSigCutAtom, LJAttractFixedRepulsePair>
    kind = capi.PAIR_LJATTRACTFIXEDREPULSE


class EisMclachlan(_NListed):  # NListed<EisMclachlanAtom, EisMclachlanPair>
    kind = capi.PAIR_EISMCLACHLAN


class LJish(_NListed):  # NListed<IEpsRepsSigExpCutAtom, LJishPair>
    kind = capi.PAIR_LJISH


class LJAttractRepulseSigs(_NListed):  # NListed<EpsEpsSigSigCutAtom, LJAttractRepulseSigsPair>
    kind = capi.PAIR_LJATTRACTREPULSESIGS


class HertzianDrag(_NListed):  # NListed<EpsSigExpDragAtom, RepulsionDragPair>
    kind = capi.PAIR_REPULSIONDRAG


class LoisOhern(_NListed):  # NListed<LoisOhernAtom, LoisOhernPair>
    kind = capi.PAIR_LOISOHERN


class LoisLin(_NListed):  # NListed<LoisLinAtom, LoisLinPair>
    kind = capi.PAIR_LOISLIN


class LoisOhernMin(_NListed):  # NListed<LoisOhernAtom, LoisOhernPairMinCLs>
    kind = capi.PAIR_LOISOHERNMIN


class LoisLinMin(_NListed):  # NListed<LoisLinAtom, LoisLinPairMin>
    kind = capi.PAIR_LOISLINMIN


PAIR_CLASSES = {
    capi.PAIR_LJREPULSE: LJRepulse, capi.PAIR_REPULSION: Repulsion, capi.PAIR_LJATTRACTREPULSE: LJAttractRepulse,
    capi.PAIR_LJCUT: LJCut, capi.PAIR_LJATTRACTCUT: LJAttractCut, capi.PAIR_LJATTRACTFIXEDREPULSE: LJAttractFixedRepulse,
    capi.PAIR_EISMCLACHLAN: EisMclachlan, capi.PAIR_LJISH: LJish, capi.PAIR_LJATTRACTREPULSESIGS: LJAttractRepulseSigs,
    capi.PAIR_REPULSIONDRAG: HertzianDrag, capi.PAIR_LOISOHERN: LoisOhern, capi.PAIR_LOISLIN: LoisLin,
    capi.PAIR_LOISOHERNMIN: LoisOhernMin, capi.PAIR_LOISLINMIN: LoisLinMin}


class _StatTracker:
    """Statistics trackers with their accumulators in device memory (csrc/trackers.cu)."""

    def update(self, box=None):
        self.atoms._device_op(False)
        call("parm_tracker_update", self._h)

    def reset(self):
        self.atoms._device_op(False)
        call("parm_tracker_reset", self._h)

    def counts(self):
        out = (C.c_uint64 * max(self._nlags, 1))()
        call("parm_tracker_counts", self._h, out, self._nlags)
        return [float(out[k]) for k in range(self._nlags)]


class RsqTracker(_StatTracker):  # constraints.hpp:342-368
    def __init__(self, atoms, ns, usecom=True):
        self.atoms = atoms
        ns = np.ascontiguousarray(ns, dtype=np.uint64)
        self._nlags = ns.size
        atoms._device_op(False)
        h = C.c_void_p()
        call("parm_rsq_create", atoms._h, ns.ctypes.data_as(capi.u64p), ns.size, int(usecom), C.byref(h))
        self._h = h

    def _read(self, k):
        n, D = self.atoms.n, self.atoms.ndim
        a, b, c = np.zeros((n, D)), np.zeros((n, D)), np.zeros(n)
        call("parm_rsq_read", self._h, k, _dptr(a), _dptr(b), _dptr(c))
        return a, b, c

    def xyz2(self):
        return [self._read(k)[0] for k in range(self._nlags)]

    def xyz4(self):
        return [self._read(k)[1] for k in range(self._nlags)]

    def r4(self):
        return [self._read(k)[2] for k in range(self._nlags)]

    def r2(self):  # constraints.cpp:520-535: row sums of xyz2
        out = []
        for a in self.xyz2():
            out.append(a[:, 0] + (a[:, 1] + a[:, 2]) if a.shape[1] == 3 else a[:, 0] + a[:, 1])
        return out


class ISFTracker(_StatTracker):  # constraints.hpp:393-414
    def __init__(self, atoms, ks, ns, usecom=False):
        self.atoms = atoms
        ns = np.ascontiguousarray(ns, dtype=np.uint64)
        ks = np.ascontiguousarray(ks, dtype=np.float64)
        self._nlags, self._nks = ns.size, ks.size
        atoms._device_op(False)
        h = C.c_void_p()
        call("parm_isf_create", atoms._h, _dptr(ks), ks.size, ns.ctypes.data_as(capi.u64p), ns.size, int(usecom), C.byref(h))
        self._h = h

    def ISFxyz(self):
        """per lag: complex array (nks, n, ndim)."""
        out = []
        for k in range(self._nlags):
            buf = np.zeros((self._nks, self.atoms.n, self.atoms.ndim, 2))
            call("parm_isf_read", self._h, k, _dptr(buf))
            out.append(buf[..., 0] + 1j * buf[..., 1])
        return out

    def ISFs(self):  # constraints.cpp:621-635: mean over the axes
        return [a.sum(axis=2) / a.shape[2] for a in self.ISFxyz()]


class EnergyTracker(_StatTracker):  # constraints.hpp:260-316
    _nlags = 0

    def __init__(self, atoms, interactions, n_skip=1):
        self.atoms = atoms
        self.interactions = list(interactions)
        for i in self.interactions:
            i._flush()
        atoms._device_op(False)
        arr = (C.c_void_p * max(len(self.interactions), 1))(*[i._h for i in self.interactions])
        h = C.c_void_p()
        call("parm_energy_tracker_create", atoms._h, arr, len(self.interactions), int(n_skip), C.byref(h))
        self._h = h

    def update(self, box=None):
        for i in self.interactions:
            i._ready(False)
        call("parm_tracker_update", self._h)

    def _sums(self):
        out = np.zeros(8)
        call("parm_energy_tracker_read", self._h, _dptr(out))
        return out

    def set_U0(self, U0=None):
        """set_U0(flt) or, with a Box / no argument, set_U0(Box&): the current potential energy."""
        from_box = U0 is None or isinstance(U0, OriginBox)
        for i in self.interactions:
            i._ready(False)
        call("parm_energy_tracker_set_u0", self._h, int(from_box), 0.0 if from_box else float(U0))

    def get_U0(self):
        return float(self._sums()[7])

    def n(self):
        return int(self._sums()[0])

    def E(self):
        s = self._sums()
        return s[1] / s[0]

    def U(self):
        s = self._sums()
        return s[2] / s[0]

    def K(self):
        s = self._sums()
        return s[3] / s[0]

    def E_squared_mean(self):
        s = self._sums()
        return s[4] / s[0]

    def U_squared_mean(self):
        s = self._sums()
        return s[5] / s[0]

    def K_squared_mean(self):
        s = self._sums()
        return s[6] / s[0]

    def E_std(self):
        s = self._sums()
        return np.sqrt(s[4] / s[0] - s[1] * s[1] / s[0] / s[0])

    def U_std(self):
        s = self._sums()
        return np.sqrt(s[5] / s[0] - s[2] * s[2] / s[0] / s[0])

    def K_std(self):
        s = self._sums()
        return np.sqrt(s[6] / s[0] - s[3] * s[3] / s[0] / s[0])


class Collection:
    """Collection base (collection.hpp:23-130, collection.cpp:3-208)."""

    def __init__(self, box, atoms, interactions=(), trackers=(), constraints=()):
        if constraints:
            raise capi.ParmUnsupported("Constraints are outside the hot-path scope (DESIGN.md)")
        self.box, self.atoms = box, atoms
        box._attach(atoms)
        self.interactions, self.trackers, self.stat_trackers = [], [], []
        for t in trackers:
            self._push_tracker(t)
        for i in interactions:
            self._push_interaction(i)

    def _push_interaction(self, inter):
        if not isinstance(inter, _NListed):
            raise capi.ParmUnsupported("only NListed interactions run on the device")
        inter._flush()
        self.interactions.append(inter)

    def _push_tracker(self, t):
        if isinstance(t, _StatTracker):
            self.stat_trackers.append(t)
            return
        if not isinstance(t, NeighborList):
            raise capi.ParmUnsupported("only NeighborList, RsqTracker, ISFTracker and EnergyTracker run on the device")
        self.trackers.append(t)

    def _ready(self, modifies=True):
        for i in self.interactions:
            i._flush()
        for t in self.trackers:
            t._flush()
        self.atoms._device_op(modifies)

    def add_interaction(self, inter):
        self._push_interaction(inter)
        self._ready()
        call("parm_integ_add_interaction", self._h, inter._h)

    def add_tracker(self, t):
        self._push_tracker(t)
        self._ready()
        if isinstance(t, _StatTracker):
            call("parm_integ_add_stat_tracker", self._h, t._h)
        else:
            call("parm_integ_add_tracker", self._h, t._h)

    def add(self, obj):
        return self.add_tracker(obj) if isinstance(obj, (NeighborList, _StatTracker)) else self.add_interaction(obj)

    def initialize(self):
        self._ready()
        call("parm_integ_initialize", self._h)

    def set_forces(self, constraints_and_a=True):
        self._ready()
        call("parm_integ_set_forces", self._h, int(constraints_and_a))

    def timestep(self, nsteps=1):
        self._ready()
        call("parm_integ_timestep", self._h, int(nsteps))

    def update_trackers(self):
        self._ready()
        call("parm_integ_update_trackers", self._h)

    def potential_energy(self):
        self._ready(False)
        e = C.c_double(0)
        call("parm_integ_potential_energy", self._h, C.byref(e))
        return e.value

    def kinetic_energy(self):
        return self.atoms.kinetic_energy()

    def energy(self):
        return self.potential_energy() + self.kinetic_energy()

    def virial(self):
        self._ready(False)
        e = C.c_double(0)
        call("parm_integ_virial", self._h, C.byref(e))
        return e.value

    def pressure(self):  # collection.cpp:85-96
        return (2.0 * self.kinetic_energy() + self.virial()) / self.box.V() / float(self.atoms.ndim)

    def degrees_of_freedom(self):
        return self.atoms._reduce(capi.RED_NDOF)

    def com_velocity(self):
        return self.atoms.com_velocity()

    def temp(self, minuscomv=True):  # collection.cpp:135-142
        v = self.atoms.com_velocity() if minuscomv else None
        ndof = int(self.degrees_of_freedom())
        if minuscomv:
            ndof -= self.atoms.ndim
        return self.atoms.kinetic_energy(v) * 2 / ndof

    def reset_com_velocity(self):
        self.atoms.reset_com_velocity()

    def scale_velocities(self, scaleby):
        self.atoms._device_op()
        call("parm_scale_velocities", self.atoms._h, float(scaleby))

    def scale_velocities_to_temp(self, T, minuscomv=True):  # collection.cpp:31-35
        t = self.temp(minuscomv)
        self.scale_velocities(np.sqrt(T / t))

    def scale_velocities_to_energy(self, E):  # collection.cpp:37-43
        E0 = self.energy()
        k0 = self.kinetic_energy()
        goalkinetic = k0 + (E - E0)
        self.scale_velocities(np.sqrt(goalkinetic / k0))

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        call("parm_integ_stats", self._h, C.byref(a), C.byref(b), C.byref(c))
        return dict(steps=a.value, rebuilds=b.value, launches=c.value)


class CollectionVerlet(Collection):
    def __init__(self, box, atoms, dt, interactions=(), trackers=(), constraints=()):
        Collection.__init__(self, box, atoms, interactions, trackers, constraints)
        h = C.c_void_p()
        call("parm_verlet_create", atoms._h, float(dt), C.byref(h))
        self._h = h
        self.dt = float(dt)
        _construct(self)

    def set_dt(self, dt):
        self.dt = float(dt)
        call("parm_integ_set_dt", self._h, self.dt)


class CollectionSol(Collection):
    def __init__(self, box, atoms, dt, damping, desired_temperature, interactions=(), trackers=(), constraints=(), seed=0):
        Collection.__init__(self, box, atoms, interactions, trackers, constraints)
        h = C.c_void_p()
        call("parm_sol_create", atoms._h, float(dt), float(damping), float(desired_temperature), int(seed), C.byref(h))
        self._h = h
        self.dt = float(dt)
        _construct(self)

    def set_dt(self, dt):
        self.dt = float(dt)
        call("parm_integ_set_dt", self._h, self.dt)

    def change_temperature(self, damping, desired_temperature):
        call("parm_integ_set_temperature", self._h, float(damping), float(desired_temperature))

    def inject_noise(self, z):
        """Exact-parity hook: z (steps, n_mobile, 2, ndim) standard normals (see parm_integ_inject_noise)."""
        self._ready()
        if z is None:
            call("parm_integ_inject_noise", self._h, None, 0)
            return
        z = np.ascontiguousarray(z, dtype=np.float64)
        call("parm_integ_inject_noise", self._h, _dptr(z), z.size)

    def sol_constants(self):
        out = np.zeros(6)
        call("parm_integ_get_sol_constants", self._h, _dptr(out))
        return out


class _GenericCollection(Collection):
    """The other fixed-box integrators (include/parm_b200.h PARM_INTEG_*)."""
    integ_type = None

    def _make(self, box, atoms, params, interactions, trackers, constraints, seed=0):
        Collection.__init__(self, box, atoms, interactions, trackers, constraints)
        h = C.c_void_p()
        p = np.ascontiguousarray(params, dtype=np.float64)
        call("parm_integ_create", atoms._h, self.integ_type, _dptr(p), p.size, int(seed), C.byref(h))
        self._h = h
        self.dt = float(params[0])
        _construct(self)

    def set_dt(self, dt):
        self.dt = float(dt)
        call("parm_integ_set_dt", self._h, self.dt)

    def _scalars(self):
        out = np.zeros(2)
        call("parm_integ_get_scalars", self._h, _dptr(out))
        return out


class CollectionDamped(_GenericCollection):  # collection.hpp:273-295
    integ_type = capi.INTEG_DAMPED

    def __init__(self, box, atoms, dt, damping, interactions=(), trackers=(), constraints=()):
        self._make(box, atoms, (dt, damping), interactions, trackers, constraints)

    def change_damping(self, damp):
        call("parm_integ_set_param", self._h, 3, float(damp))


class CollectionSolHT(_GenericCollection):  # collection.hpp:328-353
    integ_type = capi.INTEG_SOLHT
    inject_noise = CollectionSol.inject_noise  # z: (steps, n_mobile, ndim) standard normals

    def __init__(self, box, atoms, dt, damping, desired_temperature, interactions=(), trackers=(), constraints=(), seed=0):
        self._make(box, atoms, (dt, damping, desired_temperature), interactions, trackers, constraints, seed)

    def change_temperature(self, newdt, damp, desired_temperature):
        self.set_dt(newdt)
        call("parm_integ_set_param", self._h, 3, float(damp))
        call("parm_integ_set_param", self._h, 2, float(desired_temperature))


class CollectionOverdamped(_GenericCollection):  # collection.hpp:376-393
    integ_type = capi.INTEG_OVERDAMPED

    def __init__(self, box, atoms, dt, gamma=1.0, interactions=(), trackers=(), constraints=()):
        self._make(box, atoms, (dt, gamma), interactions, trackers, constraints)


class CollectionNoseHoover(_GenericCollection):  # collection.hpp:567-598
    integ_type = capi.INTEG_NOSEHOOVER

    def __init__(self, box, atoms, dt, Q, T, interactions=(), trackers=(), constraints=()):
        self._make(box, atoms, (dt, Q, T), interactions, trackers, constraints)
        self.Q, self.T = float(Q), float(T)

    def set_Q(self, Q):
        self.Q = float(Q)
        call("parm_integ_set_param", self._h, 1, self.Q)

    def reset_bath(self):
        call("parm_integ_reset_bath", self._h)

    def get_xi(self):
        return float(self._scalars()[0])

    def get_lns(self):
        return float(self._scalars()[1])

    def hamiltonian(self):  # collection.cpp:1244-1249
        xi, lns = self._scalars()
        return (self.kinetic_energy() + self.potential_energy() + (xi * xi * self.Q / 2) +
                (self.degrees_of_freedom() * lns * self.T))


class CollectionGaussianT(_GenericCollection):  # collection.hpp:600-621
    integ_type = capi.INTEG_GAUSSIANT

    def __init__(self, box, atoms, dt, interactions=(), trackers=(), constraints=()):
        self._make(box, atoms, (dt,), interactions, trackers, constraints)


class CollectionGear3A(_GenericCollection):  # collection.hpp:623-637
    integ_type = capi.INTEG_GEAR3A

    def __init__(self, box, atoms, dt, interactions=(), trackers=(), constraints=()):
        self._make(box, atoms, (dt,), interactions, trackers, constraints)


class _GearN(_GenericCollection):
    def __init__(self, box, atoms, dt, ncorrectionsteps=1, interactions=(), trackers=(), constraints=()):
        self._make(box, atoms, (dt, ncorrectionsteps), interactions, trackers, constraints)


class CollectionGear4A(_GearN):  # collection.hpp:639-671
    integ_type = capi.INTEG_GEAR4A


class CollectionGear5A(_GearN):  # collection.hpp:673-709
    integ_type = capi.INTEG_GEAR5A


class CollectionGear6A(_GearN):  # collection.hpp:711-755
    integ_type = capi.INTEG_GEAR6A


class CollectionNLCG(Collection):
    """CollectionNLCG (collection.hpp:400-474): the packer's conjugate-gradient minimiser of U + P0 V."""

    def __init__(self, box, atoms, dt, P0, interactions=(), trackers=(), constraints=(), kappa=10.0, kmax=1000, secmax=40,
                 seceps=1e-20):
        Collection.__init__(self, box, atoms, interactions, trackers, constraints)
        h = C.c_void_p()
        call("parm_nlcg_create", atoms._h, float(dt), float(P0), float(kappa), float(kmax), int(secmax), float(seceps), C.byref(h))
        self._h = h
        self.dt = float(dt)
        _construct(self)

    def _state(self):
        out = np.zeros(16)
        call("parm_nlcg_get", self._h, _dptr(out))
        return out

    def _set(self, which, value):
        self._ready()
        call("parm_nlcg_set", self._h, which, float(value))

    def _reduce(self, what):
        self._ready(False)
        out = C.c_double(0)
        call("parm_nlcg_reduce", self._h, what, C.byref(out))
        return out.value

    def timestep(self, nsteps=1):
        Collection.timestep(self, nsteps)
        self.box._pull(self.atoms)

    def set_forces(self, constraints_and_a=True, setV=True):
        self._ready()
        call("parm_nlcg_set_forces", self._h, int(constraints_and_a), int(setV))

    def reset(self):
        self._ready()
        call("parm_nlcg_reset", self._h)

    def descend(self):
        self._ready()
        call("parm_nlcg_descend", self._h)
        self.box._pull(self.atoms)

    def set_dt(self, dt):
        self.dt = float(dt)
        self._set(0, dt)

    def set_pressure_goal(self, P):
        self._set(1, P)

    def get_pressure_goal(self):
        return float(self._state()[1])

    P0 = property(get_pressure_goal)

    def set_kappa(self, k):
        self._set(2, k)

    def set_max_alpha(self, a):
        self._set(3, a)

    def set_max_alpha_fraction(self, a):
        self._set(4, a)

    def set_max_dx(self, d):
        self._set(5, d)

    def set_max_step(self, m):
        self._set(6, m)

    def fdotf(self):
        return self._reduce(0)

    def fdota(self):
        return self._reduce(1)

    def fdotv(self):
        return self._reduce(2)

    def vdotv(self):
        return self._reduce(3)

    def kinetic_energy(self):
        return self._reduce(4)

    def pressure(self):
        return self._reduce(5)

    def hamiltonian(self):
        return self._reduce(6)


INTEGRATOR_CLASSES = {
    capi.INTEG_VERLET: CollectionVerlet, capi.INTEG_SOL: CollectionSol, capi.INTEG_DAMPED: CollectionDamped,
    capi.INTEG_SOLHT: CollectionSolHT, capi.INTEG_OVERDAMPED: CollectionOverdamped,
    capi.INTEG_NOSEHOOVER: CollectionNoseHoover, capi.INTEG_GAUSSIANT: CollectionGaussianT,
    capi.INTEG_GEAR3A: CollectionGear3A, capi.INTEG_GEAR4A: CollectionGear4A, capi.INTEG_GEAR5A: CollectionGear5A,
    capi.INTEG_GEAR6A: CollectionGear6A, capi.INTEG_NLCG: CollectionNLCG}


def _construct(collec):
    """Collection constructor tail: the vectors were stored as passed (no per-add update_trackers),
    then Collection::initialize() runs (collection.cpp:3-19)."""
    collec._ready()
    lib = capi.lib()
    for t in collec.trackers:
        capi.check(lib.parm_integ_register_tracker(collec._h, t._h))
    for t in collec.stat_trackers:
        capi.check(lib.parm_integ_register_stat_tracker(collec._h, t._h))
    for i in collec.interactions:
        capi.check(lib.parm_integ_register_interaction(collec._h, i._h))
    call("parm_integ_initialize", collec._h)


def from_workload(w, device=0, collection=True):
    """Build (box, atoms, interaction, neighbor list, collection) from a parm_b200.workloads dict,
    following LJatoms.cpp:30-83: NListed(box, atoms, skin); add atoms; update_list(true);
    Collection(box, atoms, dt); add_tracker(nl); add_interaction(I)."""
    ndim = w["ndim"]
    box = OriginBox(w["L"], ndim)
    atoms = AtomVec(w["m"], ndim=ndim, device=device)
    atoms.atoms["x"] = w["x"]
    atoms.atoms["v"] = w["v"]
    inter = PAIR_CLASSES[w["kind"]](box, atoms, w["skin"])
    from . import workloads
    eps_table, sig_table = workloads.tables(w)
    inter.add_many(w["params"], w.get("types"), eps_table, w.get("member"), sig_table)
    nl = inter.neighbor_list()
    nl.update_list(True)
    collec = None
    if collection:
        integ = int(w.get("integrator", 0))
        if integ == 0:
            collec = CollectionVerlet(box, atoms, w["dt"])
        elif integ == 1:
            collec = CollectionSol(box, atoms, w["dt"], w["damping"], w["T"], seed=w.get("seed", 0))
        elif integ == capi.INTEG_NLCG:  # integ_params: (P0, kappa, kmax, secmax, seceps)
            P0, kappa, kmax, secmax, seceps = w["integ_params"]
            collec = CollectionNLCG(box, atoms, w["dt"], P0, kappa=kappa, kmax=kmax, secmax=int(secmax), seceps=seceps)
        else:  # w["integ_params"]: the constructor arguments after dt
            collec = INTEGRATOR_CLASSES[integ](box, atoms, w["dt"], *w.get("integ_params", ()))
        collec.add_tracker(nl)
        collec.add_interaction(inter)
    return box, atoms, inter, nl, collec
