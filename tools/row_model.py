#!/usr/bin/env python
"""CPU model of what the cell-tile pair kernel walks (no GPU): from the positions of a config-3-like LJ system it
rebuilds, in numpy, the slot order, the column chunks and the full neighbour rows the way nlist.cu / tile.cu lay them
out and reports the numbers the design notes argue with (DESIGN.md sections 4 and 9):
  * row lengths and the entries a warp (8 consecutive atoms of a chunk, 4 lanes each) really walks when its rows are
    padded to 32 entries of the warp's longest row (round 1 / first half of round 2) or to 8 (partial last pass);
  * the share of listed pairs that lie beyond the cut-off (they are in the list because of the skin);
  * the share of listed pairs whose two atoms sit in the SAME chunk (what Newton's third law inside a chunk could save).
   python tools/row_model.py [--side 36] [--positions melted.npy] [--chunk 120]"""
import argparse
import json
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=36)
    ap.add_argument("--positions", default=None, help=".npy from tests/make_melted_positions.py (liquid state)")
    ap.add_argument("--chunk", type=int, default=120)
    a = ap.parse_args()
    w = W.lj_lattice((a.side,) * 3, seed=3003)
    L = w["L"]
    if a.positions:
        w["x"] = np.load(a.positions)
    x = np.mod(w["x"], L)
    n = len(x)
    rc, rl = 2.5, 2.5 + w["skin"]
    nc = np.floor(L / rl).astype(int)
    cell3 = np.minimum((x / L * nc).astype(int), nc - 1)
    cid = (cell3[:, 0] * nc[1] + cell3[:, 1]) * nc[2] + cell3[:, 2]
    order = np.argsort(cid, kind="stable")            # slot -> atom
    xs = x[order]
    col = (cell3[order, 0] * nc[1] + cell3[order, 1])  # column of every slot
    # chunks: every column's atoms spread evenly over ceil(count / chunk) chunks (tile.cu: k_tile_chunks)
    chunk_of = np.empty(n, np.int64)
    first_in_chunk = np.empty(n, np.int64)
    nchunks = 0
    col_start = np.searchsorted(col, np.arange(nc[0] * nc[1] + 1))
    for q in range(nc[0] * nc[1]):
        a0, a1 = col_start[q], col_start[q + 1]
        if a1 == a0:
            continue
        k = -(-(a1 - a0) // a.chunk)
        size = -(-(a1 - a0) // k)
        idx = np.arange(a1 - a0)
        chunk_of[a0:a1] = nchunks + idx // size
        first_in_chunk[a0:a1] = a0 + (idx // size) * size
        nchunks += k
    tree = cKDTree(xs, boxsize=L)
    pairs = tree.query_pairs(rl, output_type="ndarray")
    d = xs[pairs[:, 0]] - xs[pairs[:, 1]]
    d -= L * np.rint(d / L)
    r = np.sqrt((d * d).sum(1))
    cnt = np.bincount(np.concatenate([pairs[:, 0], pairs[:, 1]]), minlength=n)
    # warps: 8 consecutive atoms of a chunk (64 teams per block: atoms a0 + 8 w .. a0 + 8 w + 7 of each round of 64)
    rank = np.arange(n) - first_in_chunk
    warp_key = chunk_of * 1000 + rank // 8
    _, inv = np.unique(warp_key, return_inverse=True)
    longest = np.zeros(inv.max() + 1, np.int64)
    np.maximum.at(longest, inv, cnt)
    lanes = np.bincount(inv)                           # atoms in the warp (the last warp of a chunk is partial)
    def walked(gran):
        return ((-(-longest // gran)) * gran * 8).sum() / n  # entries per atom the warps execute (idle teams included)
    same_chunk = chunk_of[pairs[:, 0]] == chunk_of[pairs[:, 1]]
    # lead: deal the atoms of a chunk to the teams in order of row length (a per-chunk permutation table), so that the 8
    # rows a warp walks in lock step are as equal as they can be
    srt = np.lexsort((-cnt, chunk_of))                 # by chunk, longest rows first
    rank_s = np.empty(n, np.int64)
    rank_s[srt] = np.arange(n) - np.searchsorted(chunk_of[srt], chunk_of[srt])
    _, inv_s = np.unique(chunk_of * 1000 + rank_s // 8, return_inverse=True)
    longest_s = np.zeros(inv_s.max() + 1, np.int64)
    np.maximum.at(longest_s, inv_s, cnt)
    walked_sorted8 = ((-(-longest_s // 8)) * 8 * 8).sum() / n
    out = {
        "atoms": int(n), "cells": nc.tolist(), "chunks": int(nchunks), "atoms_per_chunk": n / nchunks,
        "state": "positions from " + a.positions if a.positions else "jittered lattice (t = 0)",
        "mean_full_neighbours": float(cnt.mean()), "row_length_std": float(cnt.std()), "row_length_max": int(cnt.max()),
        "mean_longest_row_of_a_warp": float(longest.mean()),
        "entries_walked_per_atom_pad32": walked(32), "entries_walked_per_atom_pad8": walked(8),
        "passes_per_warp_and_atom_pad32": float((-(-longest // 32)).mean()),
        "pad8_over_pad32": walked(8) / walked(32),
        "entries_walked_per_atom_pad8_rows_sorted_by_length_inside_the_chunk": walked_sorted8,
        "partial_warps_share": float((lanes < 8).mean()),
        "listed_pairs_beyond_cutoff": float((r > rc).mean()),
        "listed_pairs_inside_one_chunk": float(same_chunk.mean()),
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
