import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from parm_b200 import sim, workloads as W
from parity_util import cpu_system, rel_err, rel_err_vec
w = W.random_system(700, 3, 2, seed=33, ntypes=3, frozen=4)
box, atoms, inter, nl, _ = sim.from_workload(w, collection=False)
ga, gb = nl.pairs()
c = cpu_system('port', w, collection=False)
ca, cb = c.pairs()
print('pairs', len(ga), len(ca), np.array_equal(ga,ca) and np.array_equal(gb,cb))
atoms.reset_forces(); gp = inter.set_forces_get_pressure(box); gf = atoms.peek('f')
cf, cp = c.forces_and_pressure()
print('f', rel_err_vec(gf,cf), 'p', rel_err(gp,cp), 'E', rel_err(inter.energy(), c.inter_energy()))
print(nl.stats())
