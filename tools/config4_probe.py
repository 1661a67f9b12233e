#!/usr/bin/env python
"""Small driver for ncu captures of BASELINE config 4 (binary WCA-like LJRepulsePair, N = 4e6, CollectionSol).
   python tools/config4_probe.py [--steps 8]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import capi, sim, workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    a = ap.parse_args()
    w = W.config4()
    box, atoms, inter, nl, collec = sim.from_workload(w)
    collec.set_forces(True)
    collec.timestep(a.steps)
    capi.call("parm_sync", atoms._h)
    print("config4 N", atoms.n, "rebuilds", nl.which(), "mean n", nl.stats())


if __name__ == "__main__":
    main()
