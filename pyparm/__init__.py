"""pyparm: the reference's Python package name over the B200 library.

The reference builds `pyparm.d2` / `pyparm.d3` from src/sim.i with SWIG (sim.i:616-670 names the
NListed instantiations); SWIG, Eigen and Boost are absent here, so the two modules are thin
NDIM-bound views of parm_b200.sim (ctypes over the C ABI in include/parm_b200.h). Code written
against the reference -- `from pyparm import d3 as sim3`, pyparm/tests.py:3-4 -- finds the same
class names, constructor signatures and methods. `pyparm.xyzfile` writes trajectories from
asynchronous frame downloads (parm_snapshot_begin / _wait).
"""
from . import d2, d3  # noqa: F401
