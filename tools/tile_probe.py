#!/usr/bin/env python
"""Small driver for ncu captures of the cell-tile pair kernel: N = side^3 LJ lattice, a few steps.
   python tools/tile_probe.py [--side 100] [--steps 8]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import capi, sim, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=100)
    ap.add_argument("--steps", type=int, default=8)
    a = ap.parse_args()
    w = W.config3(a.side)
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(a.steps)
    capi.call("parm_sync", atoms._h)
    print("tile", nl.tile_stats(), "rebuilds", nl.which(), "E", collec.energy())


if __name__ == "__main__":
    main()
