#!/usr/bin/env python
"""CPU baseline table of BASELINE.md section 3: the reference's own code (oracle/_ref, unmodified sources, 1 thread)
on the BASELINE potentials at N in {1000, 4096, 10648, 32768}:
  (i)  steady-state atom-steps/s of Collection*::timestep() with the pair list kept fresh by the harness cell list
       (InjectedNeighborList: same pair set as the reference's update_list, proved at small N), and
  (ii) seconds per NeighborList::update_list(true) -- the reference's own O(N^2) rebuild -- separately.
   python tests/cpu_baseline_table.py [--max-n 32768] [--rebuild-max-n 10648] [--pot lj,wca,harm2d] [--seconds 4]
The CPU figures are a reported baseline, not an optimisation target. Lives under tests/ because it drives oracle/
(test infrastructure; nothing outside tests/, smoke() and bench.py's cpu_baseline leg touches it)."""
import argparse
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402
from parm_b200 import workloads as W  # noqa: E402

SIDES3 = {1000: (10, 10, 10), 4096: (16, 16, 16), 10648: (22, 22, 22), 32768: (32, 32, 32)}
SIDES2 = {1000: (25, 40), 4096: (64, 64), 10648: (88, 121), 32768: (128, 256)}


def workload(pot, n):
    if pot == "lj":
        return W.lj_lattice(SIDES3[n], seed=3003)          # config 3/5 state point
    if pot == "wca":
        return W.config4(SIDES3[n], seed=4004)             # config 4 (CollectionSol)
    return W.config2(*SIDES2[n], seed=2002)                # config 2 (2-D)


def system(w, backend, injected):
    s = cpu.CpuSystem(backend, w["L"], w["x"], w["v"], w["m"])
    eps_table, sig_table = W.tables(w)  # only the functors whose atom struct carries `epsilons` take a table
    s.add_interaction(w["kind"], w["skin"], w["params"], w.get("types"), eps_table, injected=injected, sig_table=sig_table)
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=int, default=32768)
    ap.add_argument("--rebuild-max-n", type=int, default=10648)
    ap.add_argument("--pot", default="lj,wca,harm2d")
    ap.add_argument("--seconds", type=float, default=4.0, help="steady-state timing window per case")
    a = ap.parse_args()
    backend = "ref" if cpu.have("ref", 3) and cpu.have("ref", 2) else "port"
    if backend == "port":
        cpu.build(("port",))
    host = {"cpu": platform.processor() or platform.machine(), "nproc": os.cpu_count(), "threads_used": 1,
            "backend": "unmodified reference sources (oracle/_ref)" if backend == "ref" else "C restatement (oracle/parm_oracle.c)"}
    try:
        with open("/proc/cpuinfo") as fh:
            host["cpu"] = [l.split(":", 1)[1].strip() for l in fh if l.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps({"host": host}), flush=True)
    for pot in a.pot.split(","):
        for n in sorted(SIDES3):
            if n > a.max_n:
                continue
            w = workload(pot, n)
            s = system(w, backend, injected=True)
            s.update_list(True)
            pairs0 = len(s.pairs()[0])  # at the initial positions, compared with update_list(true) below
            s.make_collection(w.get("integrator", 0), w["dt"], w.get("damping", 0.0), w.get("T", 0.0))
            s.set_forces(True)
            s.timestep(3)
            k, t = 0, 0.0
            chunk = max(1, int(2e5 / n))
            r0 = s.which()
            t0 = time.perf_counter()
            while t < a.seconds:
                s.timestep(chunk)
                k += chunk
                t = time.perf_counter() - t0
            out = {"potential": pot, "workload": w.get("name", "lj_lattice"), "n_atoms": n, "ndim": w["ndim"], "steps": k,
                   "atom_steps_per_s": n * k / t, "ms_per_step": t / k * 1e3, "rebuilds_in_window": s.which() - r0,
                   "pairs": pairs0}
            s.close()
            if n <= a.rebuild_max_n:
                s2 = system(w, backend, injected=False)
                t0 = time.perf_counter()
                s2.update_list(True)   # NeighborList::update_list(true), trackers.cpp:55-69: all (i, j<i) pairs
                out["update_list_s"] = time.perf_counter() - t0
                assert len(s2.pairs()[0]) == out["pairs"], "harness cell list and update_list(true) disagree"
                out["ns_per_pair_check"] = out["update_list_s"] / (n * (n - 1) / 2) * 1e9
                s2.close()
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
