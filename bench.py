#!/usr/bin/env python
"""Headline benchmark: atom-steps/s of the MD hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" is one CollectionVerlet::timestep() (collection.cpp:442-469) over the whole system:
K1 integrate + skin-drift reduction -> pair forces -> K3 integrate -> (when the drift rule triggered) neighbour rebuild.
Rebuilds are inside the timed region (amortised), setup is not. The timed state is the equilibrated T = 1.44
configuration (--equil untimed steps with velocity rescaling from the jittered lattice), not the lattice itself.

N = 1 workload: BASELINE.json configs[2] -- 3-D LJ, LJAttractRepulsePair cut 2.5 sigma, N = 1e6
(the largest single-GPU configuration the metric is quoted on; SURVEY 8d cfg 3: rho = 1.1939,
T = 1.44, skin 0.3, dt 0.004).  N > 1: the same per-GPU slab stacked along x, the slab axis (weak scaling),
i.e. 1e6 atoms per GPU, slab-decomposed with NCCL halo exchange (config 5 is N=16e6 on 8 GPUs
= 2e6 per GPU; the 8-GPU run also times it and prints it as config5_16M).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "atom-steps/sec (3D LJ, N=1M–16M) at 1/2/4/8 B200 + % HBM roofline"
UNIT = "atom-steps/s"
SAMPLE_SIDE = 32  # CPU reference sample: 32^3 = 32768 atoms of the same state point


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


EQUIL_EVERY = 100


def equilibrate(collec, steps, T=1.44):
    """Untimed: melt the jittered simple-cubic start and hold the state point (collection.cpp:31-35 rescaling every
    EQUIL_EVERY steps), so that the timed region sees the disordered T = 1.44 configuration the config names."""
    done = 0
    while done < steps:
        k = min(EQUIL_EVERY, steps - done)
        collec.timestep(k)
        collec.scale_velocities_to_temp(T)
        done += k
    return {"steps": steps, "rescale_every": EQUIL_EVERY, "T_target": T, "T_end": float(collec.temp())}


def align_to_rebuild(collec, max_steps=12):
    """Untimed: single steps until one of them ends with a neighbour-list rebuild, so that the timed window always
    starts from a freshly built list (like the first step of any run). A 20-step window then holds the same number of
    rebuilds in every run and every round; the unaligned long-window figure is `steady_state`."""
    r0 = collec.stats()["rebuilds"]
    for k in range(max_steps):
        collec.timestep(1)
        if collec.stats()["rebuilds"] > r0:
            return k + 1
    return max_steps


def kernel_source_sha16():
    """Hash of the pair-kernel sources of the library that is running (ties profiles/force_kernel_traffic.json to it)."""
    import hashlib
    h = hashlib.sha256()
    for f in ("force_tile.cuh", "force_kernel.cuh", "tile.cu", "force.cu", "pairs.cuh"):
        with open(os.path.join(ROOT, "parm_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def kernel_traffic(tile_active):
    """DRAM bytes per launch of the pair kernel from the committed ncu capture -- only when that capture was taken
    from these very kernel sources (source_sha16), else None (the number would be stale)."""
    tp = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    try:
        tj = json.load(open(tp))
        ent = tj if tile_active else tj.get("gather_kernel", {})
        if ent.get("source_sha16") != kernel_source_sha16():
            return None, "profiles/force_kernel_traffic.json is from other kernel sources (sha %s, running %s)" % (
                ent.get("source_sha16"), kernel_source_sha16())
        return ent.get("dram_bytes_per_launch"), ent.get("kernel")
    except Exception as exc:
        return None, "no capture (%s)" % exc


def build_cpu_sample(backend_pref=("ref", "port")):
    """The reference's own CPU implementation on a bounded sample of the same workload."""
    from oracle import cpu
    from parm_b200 import workloads
    w = workloads.lj_lattice((SAMPLE_SIDE,) * 3, seed=3003)
    kind = None
    for be in backend_pref:
        if cpu.have(be, 3):
            kind = be
            break
    if kind is None:
        cpu.build(("port",))
        kind = "port"
    s = cpu.CpuSystem(kind, w["L"], w["x"], w["v"], w["m"])
    # the reference's O(N^2) update_list would take ~1 min per rebuild at this N:
    # pairs come from the harness cell list (proved identical to update_list at small N)
    s.add_interaction(w["kind"], w["skin"], w["params"], w["types"], w["eps_table"], injected=True)
    s.update_list(True)
    s.make_collection(0, w["dt"])
    s.set_forces(True)
    return s, w, kind


def time_cpu(steps, warmup):
    s, w, kind = build_cpu_sample()
    n = w["x"].shape[0]
    s.timestep(warmup)
    t0 = time.perf_counter()
    s.timestep(steps)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt, n, kind


def cpu_sample_desc(n, steps, kind):
    return ("%d-atom cubic sample (32^3 sites) of the same LJ state point (rho=1.1939, T=1.44, cut 2.5, skin 0.3, "
            "dt 0.004), %d CollectionVerlet steps, 1 thread; pair lists from the harness cell list "
            "(the reference's own O(N^2) update_list needs ~1 min per rebuild at this N and is excluded); %s"
            % (n, steps, "unmodified reference sources compiled into oracle/_ref" if kind == "ref"
               else "plain-C restatement oracle/parm_oracle.c"))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    val, dt, n, kind = time_cpu(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, args.gpus), n_atoms_timed=int(n),
                       timed_sample="the CPU arm times a %d-atom sample of this workload (see cpu_baseline.sample)" % n),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference" if kind == "ref" else "port",
                         "sample": cpu_sample_desc(n, steps, kind)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    n = args.side * args.side * args.side_z * world
    return {"workload": "3D LJ (NListed<IEpsSigCutAtom,LJAttractRepulsePair>, cut 2.5 sigma), CollectionVerlet NVE, "
                        "simple-cubic start %dx%dx%d per GPU melted and held at T=1.44 for %d untimed steps before the "
                        "timed region, rho=1.1939, skin=0.3, dt=0.004" %
                        (args.side, args.side, args.side_z, args.equil),
            "n_atoms": n, "atoms_per_gpu": n // world, "parallelism": "slab%d" % world if world > 1 else "single",
            "l2": "working set (positions+velocities+forces+neighbour list ~0.7 GB per 1e6 atoms) is larger than the "
                  "126 MB L2, no explicit flush"}


def run_ours(args):
    import torch
    from parm_b200 import capi, sim, workloads
    from parm_b200.capi import C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from parm_b200 import sharded

        def parity_hook(rank, world):
            # the oracle as the checker, before the timed region (tests/mgpu_check.py "lj3d_tile": slab-decomposed
            # run vs the CPU restatement on rank 0); the numbers travel in the bench line as "parity_check"
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import mgpu_check
            w = workloads.lj_lattice((14 * world, 28, 28), seed=17)
            w["name"] = "lj3d_tile %dx28x28" % (14 * world)
            return mgpu_check.parity_metrics(w, 25)

        def cpu_hook():
            val, dt, ncpu, kind = time_cpu(100, 3)
            return {"value": val, "unit": UNIT, "cores": 1, "kind": "reference" if kind == "ref" else "port",
                    "sample": cpu_sample_desc(ncpu, 100, kind)}

        hooks = {"equilibrate": equilibrate, "cpu_baseline": cpu_hook, "align": align_to_rebuild}
        if not args.no_parity:
            hooks["parity"] = parity_hook
        return sharded.bench_main(args, rank, world, local, METRIC, UNIT, workload_config(args, world), peaks(), hooks)
    torch.cuda.set_device(local)
    w = workloads.lj_lattice((args.side, args.side, args.side_z), seed=3003)
    n = w["x"].shape[0]
    box, atoms, inter, nl, collec = sim.from_workload(w, device=local)
    collec.set_forces(True)
    st = C.c_void_p()
    capi.call("parm_get_stream", atoms._h, C.byref(st))
    stream = torch.cuda.ExternalStream(st.value, device=local)
    K, W = args.steps, max(args.warmup, 3)

    equil = equilibrate(collec, args.equil)
    collec.timestep(W)
    equil["alignment_steps"] = align_to_rebuild(collec)
    equil["timed_region_starts"] = "on the first step after a neighbour-list rebuild"
    capi.call("parm_sync", atoms._h)
    # ---- timed region: K steps, device clock, rebuilds included
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = capi.lib().parm_b200_launch_count()
    r0 = collec.stats()["rebuilds"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    collec.timestep(K)
    e1.record(stream)
    capi.call("parm_sync", atoms._h)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = capi.lib().parm_b200_launch_count() - l0
    rebuilds = collec.stats()["rebuilds"] - r0
    clocks = sampler.stop()
    value = n * K / (ms * 1e-3)

    # ---- same K steps again with per-kernel-class CUDA events (roofline of the dominant kernel)
    capi.call("parm_profile_enable", atoms._h, 1)
    collec.timestep(K)
    pms = (C.c_double * 4)()
    pcnt = (C.c_uint64 * 4)()
    capi.call("parm_profile_read", atoms._h, pms, pcnt)
    capi.call("parm_profile_enable", atoms._h, 0)
    mean_n, max_n = nl.stats()
    force_ms = pms[1] / max(pcnt[1], 1)
    hbm, hbm_src = peaks()
    bytes_force = (16 * 3 + 16 + 4 * mean_n) * n   # K2 = 16D + 16 + 4n per atom
    bytes_step = (104 * 3 + 24 + 4 * mean_n) * n   # whole step
    achieved = bytes_force / (force_ms * 1e-3) / 1e9
    traffic, traffic_src = kernel_traffic(nl.tile_stats()[0])
    roofline = {
        "bound": "hbm",
        "kernel": ("k_force_tile<LJAttractRepulse> (pair force, full neighbour rows, positions staged per cell tile in "
                   "shared memory)" if nl.tile_stats()[0] else "k_force<LJAttractRepulse> (pair force, full neighbour rows)"),
        "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
        "traffic_source": traffic_src, "kernel_source_sha16": kernel_source_sha16(),
        "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_force, "mean_full_neighbors": mean_n,
        "kernel_ms": force_ms,
        "fp64_gflops_est": 30.0 * mean_n * n / (force_ms * 1e-3) / 1e9,  # ~30 flop/pair, SURVEY 8d
        "tile": dict(zip(("active", "chunks", "max_tile_atoms", "wide_chunks"), nl.tile_stats())),
        "step_share": {"integrate1_drift_ms": pms[0] / max(pcnt[0], 1), "force_ms": force_ms,
                       "integrate2_ms": pms[2] / max(pcnt[2], 1),
                       "rebuild_ms_each": pms[3] / max(pcnt[3], 1), "rebuilds": int(pcnt[3]), "steps": K},
        "whole_step_gbs": bytes_step * K / (ms * 1e-3) / 1e9,
    }

    # ---- steady state: a window long enough that the rebuild count is not quantised (the driver's K may hold 3 or 4)
    Ks = max(args.steady_steps, K)
    r0 = collec.stats()["rebuilds"]
    torch.cuda.synchronize()
    e0.record(stream)
    collec.timestep(Ks)
    e1.record(stream)
    capi.call("parm_sync", atoms._h)
    torch.cuda.synchronize()
    ms_s = e0.elapsed_time(e1)
    rb_s = collec.stats()["rebuilds"] - r0
    steady = {"steps": Ks, "ms_per_step": ms_s / Ks, "value": n * Ks / (ms_s * 1e-3), "unit": UNIT, "rebuilds": int(rb_s),
              "steps_per_rebuild": Ks / max(rb_s, 1), "T_end": float(collec.temp())}

    # ---- secondary bound: fp64 pipe. Peak from this library's own DFMA-chain probe (csrc/probe.cu), measured now
    # on this GPU; flops per listed pair as counted by SURVEY 8d (~30, FMA = 2).
    pf, pc = C.c_double(), C.c_double()
    try:
        capi.call("parm_b200_probe_peaks", local, C.byref(pf), C.byref(pc))
        roofline["fp64"] = {"achieved": roofline["fp64_gflops_est"] / 1e3, "peak": pf.value / 1e3, "unit": "TFLOP/s",
                            "frac": roofline["fp64_gflops_est"] / pf.value, "flops_per_pair": 30.0,
                            "peak_source": "measured now: 8 independent DFMA chains per thread (parm_b200_probe_peaks)"}
        roofline["own_copy_kernel_gbs"] = pc.value  # 1 GiB -> 1 GiB double4 streaming copy, read + write bytes
    except capi.ParmError as exc:  # the probe is reporting only: never lose the bench line over it
        roofline["fp64"] = {"error": str(exc)}

    # ---- end to end through the C ABI with HOST buffers: every step uploads the full Atom array from the
    # pinned AoS mirror, runs timestep(), and downloads the full Atom array (what a caller of the reference
    # API sees: host Atom structs current after every timestep()).
    Ke = max(3, min(args.e2e_steps, K))
    p = atoms._field_ptrs()
    isz = atoms.atoms.itemsize
    atoms.sync_to_host()
    capi.call("parm_sync", atoms._h)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(Ke):
        capi.call("parm_upload_atoms", atoms._h, capi.ALL, p[0], p[1], p[2], p[3], p[4], isz, isz)
        capi.call("parm_integ_timestep", collec._h, 1)
        capi.call("parm_download_atoms", atoms._h, capi.ALL, p[0], p[1], p[2], p[3], p[4], isz, isz)
    e1.record(stream)
    capi.call("parm_sync", atoms._h)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms_e2e = max(e0.elapsed_time(e1), wall * 1e3)
    e2e = {"value": n * Ke / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n * isz),
           "d2h_bytes_per_step": int(n * isz), "steps": Ke,
           "note": "full struct Atom array (x,v,a,f,m) copied both ways every step from pinned host memory"}

    cpu = None
    if not args.no_cpu:
        cs, cw = 100, 3   # ~15 s of single-core work
        val, dt, ncpu, kind = time_cpu(cs, cw)
        cpu = {"value": val, "unit": UNIT, "cores": 1, "kind": "reference" if kind == "ref" else "port",
               "sample": cpu_sample_desc(ncpu, cs, kind)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, 1), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches), "rebuilds_in_timed_region": int(rebuilds), "roofline": roofline,
        "steady_state": steady, "equilibration": equil, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--side", type=int, default=100, help="lattice sites along x and y")
    ap.add_argument("--side-z", type=int, default=100, help="lattice sites along z PER GPU")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--equil", type=int, default=600, help="untimed equilibration steps (velocity rescaling to T=1.44)")
    ap.add_argument("--steady-steps", type=int, default=300, help="length of the steady_state window")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the pre-run parity_check against the oracle")
    ap.add_argument("--no-config5", action="store_true", help="N = 8: skip the BASELINE config 5 (16e6 atoms) window")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
