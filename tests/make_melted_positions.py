#!/usr/bin/env python
"""Positions of the config-3 LJ system after some CollectionVerlet steps on the CPU oracle (test infrastructure), for
tools/bank_model.py --positions:   python tests/make_melted_positions.py --side 36 --steps 300 --out /tmp/melted36.npy"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402
from parm_b200 import workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=36)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--out", required=True)
a = ap.parse_args()
w = W.lj_lattice((a.side,) * 3, seed=3003)
cpu.build(("port",))
s = cpu.CpuSystem("port", w["L"], w["x"], w["v"], w["m"])
s.add_interaction(w["kind"], w["skin"], w["params"], w["types"], w["eps_table"], injected=True)
s.update_list(True)
s.make_collection(0, w["dt"])
s.set_forces(True)
s.timestep(a.steps)
np.save(a.out, s.get_atoms()[0])
