"""Slab decomposition across the GPUs of one box: one process per GPU (torchrun), NCCL over NVLink.

torch.distributed is plumbing only (rendezvous, broadcasting the NCCL unique id, the max-over-ranks
of the timings); the halo exchange, migration and reductions are NCCL calls issued by
libparm_b200.so itself on the context's stream (parm_b200/csrc/shard.cu), so a C++ caller gets the
same path through include/parm_b200.h.

Each rank owns the atoms whose wrapped x lies in [rank*Lx/G, (rank+1)*Lx/G). The reference-facing
classes of parm_b200.sim (NeighborList, LJAttractRepulse, ..., CollectionVerlet, CollectionSol) are
used unchanged on top of ShardedAtomVec; every rank makes the same calls in the same order.
"""
import json
import os
import time

import numpy as np

from . import capi, sim, workloads
from .capi import C, call, dp, u32p


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


# ---- host-side partition logic (pure numpy; covered by the gloo CPU tests) ----------------------
def slab_of(x0, Lx, world):
    """Rank owning wrapped coordinate x0 (same rule as the device binning: floor((x mod L)/Ls))."""
    w = x0 - Lx * np.floor(x0 / Lx)
    r = np.floor(w / (Lx / world)).astype(np.int64)
    return np.clip(r, 0, world - 1)


def partition_workload(w, rank, world):
    """This rank's share of a whole-system workload dict: (gid, x, v, m) of the atoms in its slab."""
    owner = slab_of(w["x"][:, 0], w["L"][0], world)
    sel = np.nonzero(owner == rank)[0]
    return sel.astype(np.uint32), w["x"][sel], w["v"][sel], w["m"][sel]


def lj_lattice_slab(side_x_per_rank, side_y, side_z, rank, world, rho=1.1939, T=1.44, jitter=0.05, seed=3003):
    """Config 3/5 state point, generated slab by slab (no rank ever holds the whole system):
    rank r owns lattice planes ix in [r*sx, (r+1)*sx). Same formulas as workloads.lj_lattice."""
    sx = side_x_per_rank
    a = rho ** (-1.0 / 3.0)
    rng = np.random.default_rng(seed + 7919 * rank)
    ix, iy, iz = np.meshgrid(np.arange(rank * sx, (rank + 1) * sx), np.arange(side_y), np.arange(side_z), indexing="ij")
    sites = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    gid = ((ix.ravel() * side_y + iy.ravel()) * side_z + iz.ravel()).astype(np.uint32)
    n = len(gid)
    x = sites * a + rng.uniform(-jitter, jitter, (n, 3)) * a + 0.5 * a
    v = rng.standard_normal((n, 3)) * np.sqrt(T)
    L = np.array([sx * world, side_y, side_z], dtype=np.float64) * a
    return dict(L=L, gid=gid, x=x, v=v, m=np.ones(n), n_global=sx * world * side_y * side_z)


def merge_pairs(parts):
    """Union of the per-rank canonical pair lists (each pair is emitted by the rank owning `first`)."""
    a = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.uint32)
    b = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.uint32)
    o = np.lexsort((b, a))
    return a[o], b[o]


# ---- distributed plumbing ----------------------------------------------------------------------
def init_distributed(backend=None):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            import torch
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=int(os.environ.get("RANK", "0")),
                                world_size=int(os.environ.get("WORLD_SIZE", "1")))
    return dist


def broadcast_bytes(payload, src=0):
    """Rank `src`'s bytes on every rank (used for the 128-byte NCCL unique id)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


class ShardedAtomVec:
    """The local part of a global AtomVec. Presents what parm_b200.sim's classes need (`_h`, `n`, `ndim`,
    `_device_op`, reductions); there is no whole-system host mirror -- use get_local()/set_local()."""

    def __init__(self, n_global, ndim, rank, world, device, cap_slots, unique_id):
        self.n = int(n_global)
        self.ndim = ndim
        self.rank, self.world = rank, world
        h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, 128)
        call("parm_ctx_create_sharded", ndim, self.n, int(cap_slots), device, rank, world, C.cast(idbuf, C.c_void_p), C.byref(h))
        self._h = h

    # sim.* compatibility
    def _device_op(self, modifies=True):
        pass

    def sync_to_host(self):
        pass

    def set_local(self, gid, x, v=None, a=None, f=None, m=None):
        gid = np.ascontiguousarray(gid, np.uint32)
        n = len(gid)
        arrs = [None if q is None else np.ascontiguousarray(q, np.float64) for q in (x, v, a, f)]
        m = np.ones(n) if m is None else np.ascontiguousarray(m, np.float64)
        call("parm_shard_set_atoms", self._h, n, gid.ctypes.data_as(u32p), _d(arrs[0]), _d(arrs[1]), _d(arrs[2]), _d(arrs[3]), _d(m))

    def info(self):
        out = (C.c_uint32 * 6)()
        call("parm_shard_info", self._h, out)
        return dict(zip(("n_local", "ghosts_down", "ghosts_up", "send_down", "send_up", "slots"), list(out)))

    def rebuild_stats(self):
        out = (C.c_uint64 * 2)()
        call("parm_shard_rebuild_stats", self._h, out)
        return {"one_sort": int(out[0]), "two_sorts": int(out[1])}

    def get_local(self):
        n = C.c_uint32(0)
        call("parm_shard_get_atoms", self._h, 0, C.byref(n), None, None, None, None, None, None)
        n = n.value
        gid = np.empty(n, np.uint32)
        x, v, a, f = (np.empty((n, self.ndim)) for _ in range(4))
        m = np.empty(n)
        call("parm_shard_get_atoms", self._h, n, None, gid.ctypes.data_as(u32p), _d(x), _d(v), _d(a), _d(f), _d(m))
        return dict(gid=gid, x=x, v=v, a=a, f=f, m=m)

    def pinned_buffers(self, cap):
        """Page-locked host arrays (gid, x, v, a, f, m) for `cap` atoms: get_local_into()/put_local_from() then move
        the local state at PCIe speed, like the pinned AoS mirror of the single-GPU AtomVec."""
        bufs = dict(gid=np.zeros(cap, np.uint32), m=np.zeros(cap))
        for k in "xvaf":
            bufs[k] = np.zeros((cap, self.ndim))
        for q in bufs.values():
            call("parm_host_register", C.c_void_p(q.ctypes.data), q.nbytes)
        self._pinned = getattr(self, "_pinned", []) + [bufs]
        return bufs

    def get_local_into(self, bufs):
        """parm_shard_get_atoms into caller-owned (pinned) arrays; returns the number of local atoms."""
        n = C.c_uint32(0)
        call("parm_shard_get_atoms", self._h, 0, C.byref(n), None, None, None, None, None, None)
        call("parm_shard_get_atoms", self._h, len(bufs["gid"]), None, bufs["gid"].ctypes.data_as(u32p), _d(bufs["x"]),
             _d(bufs["v"]), _d(bufs["a"]), _d(bufs["f"]), _d(bufs["m"]))
        return n.value

    def put_local_from(self, bufs, n):
        call("parm_shard_put_atoms", self._h, n, _d(bufs["x"]), _d(bufs["v"]), _d(bufs["a"]), _d(bufs["f"]))

    _reduce = sim.AtomVec._reduce
    mass = sim.AtomVec.mass
    momentum = sim.AtomVec.momentum
    com = sim.AtomVec.com
    com_force = sim.AtomVec.com_force
    com_velocity = sim.AtomVec.com_velocity
    kinetic_energy = sim.AtomVec.kinetic_energy
    add_velocity = sim.AtomVec.add_velocity
    reset_com_velocity = sim.AtomVec.reset_com_velocity
    reset_forces = sim.AtomVec.reset_forces

    def close(self):
        if self._h:
            for bufs in getattr(self, "_pinned", []):
                for q in bufs.values():
                    capi.lib().parm_host_unregister(C.c_void_p(q.ctypes.data))
            self._pinned = []
            capi.lib().parm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_atoms(n_global, ndim, cap_slots, rank=None, world=None, device=None):
    """Collective: creates the NCCL communicator (unique id from rank 0) and this rank's context."""
    dist = init_distributed()
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    device = int(os.environ.get("LOCAL_RANK", rank)) if device is None else device
    uid = None
    if rank == 0:
        buf = C.create_string_buffer(128)
        call("parm_nccl_unique_id", C.cast(buf, C.c_void_p))
        uid = buf.raw
    uid = broadcast_bytes(uid, 0)
    return ShardedAtomVec(n_global, ndim, rank, world, device, cap_slots, uid)


def build_system(L, n_global, gid, x, v, m, kind, params, types, eps_table, skin, dt, integrator=0, damping=0.0, T=0.0,
                 seed=0, cap_factor=1.6):
    """Collective twin of sim.from_workload for a slab-decomposed system. params/types are GLOBAL arrays."""
    ndim = x.shape[1]
    dist = init_distributed()
    world = dist.get_world_size()
    cap = int(cap_factor * max(len(gid), n_global // world)) + 4096
    atoms = make_atoms(n_global, ndim, cap)
    box = sim.OriginBox(L, ndim)
    box._attach(atoms)
    atoms.set_local(gid, x, v, None, None, m)
    inter = sim.PAIR_CLASSES[kind](box, atoms, skin)
    # the epsilon table belongs to the functors whose atom struct carries `epsilons` (workloads.tables)
    needs_table = kind in (capi.PAIR_LJATTRACTREPULSE, capi.PAIR_LJATTRACTFIXEDREPULSE, capi.PAIR_LJISH)
    inter.add_many(params, types, eps_table if needs_table else None)
    nl = inter.neighbor_list()
    nl.update_list(True)
    if integrator == 0:
        collec = sim.CollectionVerlet(box, atoms, dt)
    else:
        collec = sim.CollectionSol(box, atoms, dt, damping, T, seed=seed)
    collec.add_tracker(nl)
    collec.add_interaction(inter)
    return box, atoms, inter, nl, collec


# ---- bench.py --gpus N (N > 1) --------------------------------------------------------------------
def _bench_system(side_x_per_rank, side_y, side_z, rank, world, seed=3003):
    s = lj_lattice_slab(side_x_per_rank, side_y, side_z, rank, world, seed=seed)
    n_global = s["n_global"]
    params = np.tile(np.array([1.0, 1.0, 2.5]), (n_global, 1))
    box, atoms, inter, nl, collec = build_system(
        s["L"], n_global, s["gid"], s["x"], s["v"], s["m"], capi.PAIR_LJATTRACTREPULSE, params,
        np.zeros(n_global, np.uint32), np.ones((1, 1)), 0.3, 0.004)
    collec.reset_com_velocity()
    collec.scale_velocities_to_temp(1.44)
    collec.set_forces(True)
    return n_global, box, atoms, inter, nl, collec


def _timed_window(collec, atoms, stream, K):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events on the context's stream, max over ranks."""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = capi.lib().parm_b200_launch_count()
    r0 = collec.stats()["rebuilds"]
    dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    collec.timestep(K)
    e1.record(stream)
    call("parm_sync", atoms._h)
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), capi.lib().parm_b200_launch_count() - l0, collec.stats()["rebuilds"] - r0


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(local, sysfs="/sys"):
    """Keep this rank's host threads -- and with them the page-locked staging they first-touch -- on the NUMA node its
    GPU hangs off (the e2e leg of an N > 1 bench line moves ~200 MB per rank and step through host memory). Best
    effort: whatever goes wrong leaves the affinity as it was; returns what was done for the bench line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")) as fh:
            node = int(fh.read())
        if node < 0:
            return {"gpu": bdf, "numa_node": node, "bound": False, "reason": "no NUMA information for the device"}
        with open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)) as fh:
            cpus = _parse_cpulist(fh.read())
        use = cpus & os.sched_getaffinity(0)
        if len(use) < 16:  # (too few to share between the ranks of the node and their NCCL proxy threads: stay unbound)
            return {"gpu": bdf, "numa_node": node, "bound": False,
                    "reason": "only %d of the node's cpus are in this process's cpu set" % len(use)}
        os.sched_setaffinity(0, use)
        return {"gpu": bdf, "numa_node": node, "bound": True, "cpus": len(use)}
    except Exception as exc:  # noqa: BLE001 -- reporting only
        return {"bound": False, "reason": ("%s: %s" % (type(exc).__name__, exc))[:120]}


def bench_main(args, rank, world, local, metric, unit, config, peak, hooks=None):
    """hooks (supplied by bench.py, which alone may drive the CPU oracle): "parity" -> dict printed as parity_check,
    "cpu_baseline" -> dict (rank 0 only), "equilibrate"(collec, steps) -> dict."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    host_binding = bind_to_gpu_numa_node(local) if os.environ.get("PARM_B200_BENCH_BIND", "1") != "0" else {"bound": False, "reason": "PARM_B200_BENCH_BIND=0"}
    init_distributed("nccl")
    from bench import ClockSampler
    hooks = hooks or {}
    parity = hooks["parity"](rank, world) if "parity" in hooks else None
    # slabs stacked along x: --side-z lattice planes per GPU along the slab axis, --side x --side across
    n_global, box, atoms, inter, nl, collec = _bench_system(args.side_z, args.side, args.side, rank, world)
    st = C.c_void_p()
    call("parm_get_stream", atoms._h, C.byref(st))
    stream = torch.cuda.ExternalStream(st.value, device=local)
    K, W = args.steps, max(args.warmup, 3)
    equil = hooks["equilibrate"](collec, args.equil) if "equilibrate" in hooks else None
    collec.timestep(W)
    if equil is not None and "align" in hooks:  # the timed window starts from a freshly built list (every rank takes the same steps)
        equil["alignment_steps"] = hooks["align"](collec)
        equil["timed_region_starts"] = "on the first step after a neighbour-list rebuild"
    call("parm_sync", atoms._h)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms, launches, rebuilds = _timed_window(collec, atoms, stream, K)
    clocks = sampler.stop() if rank == 0 else None
    value = n_global * K / (ms * 1e-3)

    # per-kernel-class timing (same K steps again), force-kernel roofline on this rank's share
    call("parm_profile_enable", atoms._h, 1)
    collec.timestep(K)
    pms = (C.c_double * 4)()
    pcnt = (C.c_uint64 * 4)()
    call("parm_profile_read", atoms._h, pms, pcnt)
    call("parm_profile_enable", atoms._h, 0)
    mean_n, _ = nl.stats()
    info = atoms.info()
    force_ms = pms[1] / max(pcnt[1], 1)
    hbm, hbm_src = peak
    bytes_force = (16 * 3 + 16 + 4 * mean_n) * info["n_local"]
    achieved = bytes_force / (force_ms * 1e-3) / 1e9
    halo_bytes = 32 * (info["send_down"] + info["send_up"])   # double4 positions sent per step by this rank

    # steady state: a window long enough that the rebuild count is not quantised
    Ks = max(args.steady_steps, K)
    ms_s, _, rb_s = _timed_window(collec, atoms, stream, Ks)
    steady = {"steps": Ks, "ms_per_step": ms_s / Ks, "value": n_global * Ks / (ms_s * 1e-3), "unit": unit,
              "rebuilds": int(rb_s), "steps_per_rebuild": Ks / max(rb_s, 1), "T_end": float(collec.temp())}

    # end to end with host buffers: every step the full local state is downloaded into page-locked host arrays
    # and x, v, a, f are written back from that host copy before the next step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Ke = max(3, min(args.e2e_steps, K))
    bufs = atoms.pinned_buffers(int(1.05 * info["n_local"]) + 4096)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    nbytes_up = nbytes_dn = 0
    for _ in range(Ke):
        nloc = atoms.get_local_into(bufs)
        nbytes_dn = nloc * (4 * 8 * 3 + 8 + 4)
        nbytes_up = nloc * (4 * 8 * 3)
        atoms.put_local_from(bufs, nloc)
        call("parm_integ_timestep", collec._h, 1)
    e1.record(stream)
    call("parm_sync", atoms._h)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms_e = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device="cuda")
    dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e = {"value": n_global * Ke / (float(ms_e.item()) * 1e-3), "unit": unit, "h2d_bytes_per_step": int(nbytes_up),
           "d2h_bytes_per_step": int(nbytes_dn), "steps": Ke,
           "note": "per rank, pinned host arrays: full local state (x,v,a,f,m,id) device->host and x,v,a,f "
                   "host->device every step"}
    tile = nl.tile_stats()
    try:  # reporting only: how this rank's list rebuilds migrated their atoms (csrc/shard.cu)
        rebuild_paths = atoms.rebuild_stats()
    except Exception as exc:  # never lose the bench line over it
        rebuild_paths = {"error": str(exc)}
    del collec, inter, nl
    atoms.close()

    # BASELINE configs[4] (N = 16e6 = 200 x 200 x 400 sites on 8 GPUs, strong-scaling point of the north star): run
    # after the weak-scaling window whenever the job has 8 ranks, so that the driver's own 8-GPU run carries it
    config5 = None
    if world == 8 and not args.no_config5:
        n5, box5, atoms5, inter5, nl5, collec5 = _bench_system(50, 200, 200, rank, world, seed=5005)
        call("parm_get_stream", atoms5._h, C.byref(st))
        stream5 = torch.cuda.ExternalStream(st.value, device=local)
        eq5 = hooks["equilibrate"](collec5, args.equil) if "equilibrate" in hooks else None
        collec5.timestep(W)
        if eq5 is not None and "align" in hooks:
            eq5["alignment_steps"] = hooks["align"](collec5)
        call("parm_sync", atoms5._h)
        K5 = max(K, 100)
        ms5, _, rb5 = _timed_window(collec5, atoms5, stream5, K5)
        config5 = {"workload": "BASELINE configs[4]: 3D LJ N=16e6 (200x200x400 sites, 2e6 atoms per GPU), slab8, same "
                               "state point and equilibration as the weak-scaling line",
                   "n_atoms": int(n5), "value": n5 * K5 / (ms5 * 1e-3), "unit": unit, "ms_per_step": ms5 / K5, "steps": K5,
                   "rebuilds": int(rb5), "equilibration": eq5, "tile": list(nl5.tile_stats())}
        del collec5, inter5, nl5
        atoms5.close()

    cpu = None
    if rank == 0 and "cpu_baseline" in hooks and not args.no_cpu:
        cpu = hooks["cpu_baseline"]()
    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "rebuilds_in_timed_region": int(rebuilds),
            "roofline": {"bound": "hbm",
                         "kernel": ("k_force_tile<LJAttractRepulse> (rank 0 share; cell tiles staged in shared memory)"
                                    if tile[0] else "k_force<LJAttractRepulse> (rank 0 share)"),
                         "tile": dict(zip(("active", "chunks", "max_tile_atoms", "wide_chunks"), tile)),
                         "achieved": achieved,
                         "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": hbm_src,
                         "algorithmic_bytes_per_launch": bytes_force, "mean_full_neighbors": mean_n, "kernel_ms": force_ms,
                         "step_share": {"integrate1_drift_ms": pms[0] / max(pcnt[0], 1), "force_ms": force_ms,
                                        "integrate2_ms": pms[2] / max(pcnt[2], 1),
                                        "rebuild_ms_each": pms[3] / max(pcnt[3], 1), "rebuilds": int(pcnt[3]), "steps": K},
                         "rank0_slots": info, "halo_bytes_sent_per_step_rank0": int(halo_bytes),
                         "rank0_rebuild_paths": rebuild_paths},
            "steady_state": steady, "equilibration": equil, "parity_check": parity, "config5_16M": config5,
            "cpu_baseline": cpu, "rank0_host_binding": host_binding,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
