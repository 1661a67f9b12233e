"""Shared helpers for the parity tests."""
import numpy as np

from oracle.cpu import CpuSystem


def backends(cpu):
    """Oracle backends available on this machine: the C port always, the compiled reference if present."""
    out = ["port"]
    if cpu.have("ref", 3) and cpu.have("ref", 2):
        out.append("ref")
    return out


def cpu_system(backend, w, injected=False, collection=True):
    """Oracle twin of parm_b200.sim.from_workload (same call order as LJatoms.cpp:30-83)."""
    s = CpuSystem(backend, w["L"], w["x"], w["v"], w["m"])
    from parm_b200.workloads import tables
    eps_table, sig_table = tables(w)
    s.add_interaction(w["kind"], w["skin"], w["params"], w.get("types"), eps_table, w.get("member"),
                      injected=injected, sig_table=sig_table)
    s.update_list(True)
    if collection:
        integ = int(w.get("integrator", 0))
        if integ == 0:
            s.make_collection(0, w["dt"])
        elif integ == 1:
            s.make_collection(1, w["dt"], w["damping"], w["T"])
        else:
            s.make_collection(integ, w["dt"], params=tuple(w.get("integ_params", ())))
    return s


def rel_err_vec(a, b):
    """max |a-b| relative to the largest vector norm in b (per-component floors are meaningless)."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.sqrt((b * b).sum(-1))) if b.size else 1.0
    if scale == 0:
        scale = 1.0
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (s if s > 0 else 1.0))
