import sys, json, time
sys.path.insert(0, '/root/repo')
import numpy as np
from parm_b200 import capi, sim, workloads as W
from parm_b200.capi import C
def run(name, w, steps=100):
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True); collec.timestep(20)
    capi.call("parm_profile_enable", atoms._h, 1)
    collec.timestep(steps)
    pms = (C.c_double * 4)(); pcnt = (C.c_uint64 * 4)()
    capi.call("parm_profile_read", atoms._h, pms, pcnt)
    print(name, "force_ms %.4f" % (pms[1]/max(pcnt[1],1)), "nbr %.1f" % nl.stats()[0], flush=True)
    atoms.close()
w = W.config3(100); run("LJ one species (constants)", w)
w = W.config3(100); w["params"][::2, 2] = 2.5000001; run("LJ two species (table)", w)
w = W.config3(100); w["params"][:, 2] = 2.5 + 1e-7*np.random.default_rng(0).random(w["params"].shape[0]); run("LJ continuous (per-pair)", w)
w4 = W.config4(shape=(100,100,100)); w4["integrator"]=0; run("WCA binary (table)", w4)
w4 = W.config4(shape=(100,100,100)); w4["integrator"]=0; w4["params"][:,1]=1.2; run("WCA one species", w4)
