#!/usr/bin/env python
"""CPU model of the shared-memory wavefronts of the cell-tile pair kernel (force_tile.cuh): rebuilds, in numpy, the
slot order, the chunk tiles and the 16-bit rows of a config-3-like LJ system exactly as nlist.cu / tile.cu lay them
out, replays the LDS.128 (x,y) and LDS.64 (z) reads of every warp step of the TEAM=4, V=8 kernel and counts wavefronts
(a phase needs as many wavefronts as the most loaded bank group has DISTINCT addresses). Compares
  current : entry k of a row is read by lane k & 3 at step (k & 31) >> 2 (k_tile_localize)
  banked  : entries ordered so that lane tl of team q reads class 4*((q + step) & 3) + tl of (tile index mod 16),
            surplus entries of a class fill the holes of the others (lead 2).
   python tools/bank_model.py [--side 46] [--positions melted.npy]
No GPU needed; the model is checked against ncu: 6.4-7.9 wavefronts per LDS.128 and 3.5-4.3 per LDS.64 measured."""
import argparse
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parm_b200 import workloads as W  # noqa: E402

CH = 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=46)
    ap.add_argument("--max-chunks", type=int, default=400)
    ap.add_argument("--positions", default=None, help=".npy with positions of the same system after some steps (liquid, like "
                    "the bench after warm-up); tests/make_melted_positions.py writes one with the CPU oracle")
    a = ap.parse_args()
    w = W.lj_lattice((a.side,) * 3, seed=3003)
    L = w["L"]
    if a.positions:
        w["x"] = np.load(a.positions)
        assert w["x"].shape == (a.side ** 3, 3)
    x = np.mod(w["x"], L)
    n = len(x)
    rl = 2.5 + w["skin"]
    nc = np.floor(L / rl).astype(int)
    cell3 = np.minimum((x / L * nc).astype(int), nc - 1)
    cid = (cell3[:, 0] * nc[1] + cell3[:, 1]) * nc[2] + cell3[:, 2]
    order = np.argsort(cid, kind="stable")          # slot -> atom (stable counting sort by cell id)
    slot_of = np.empty(n, np.int64)
    slot_of[order] = np.arange(n)
    cs = cell3[order]                                # cell of every slot
    cid_s = cid[order]
    ncell = int(np.prod(nc))
    cell_start = np.searchsorted(cid_s, np.arange(ncell + 1))
    tree = cKDTree(x[order], boxsize=L)
    pairs = tree.query_pairs(rl, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]])
    j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    # build order of a row: stencil column (dx, dy), then relative z cell, then slot
    d = cs[j] - cs[i]
    d = (d + 1) % nc - 1                              # wrapped cell offsets in {-1, 0, 1}
    assert np.all(np.abs(d) <= 1)
    col = (d[:, 0] + 1) * 3 + (d[:, 1] + 1)
    key = np.lexsort((j, d[:, 2], col, i))
    i, j, col = i[key], j[key], col[key]
    row_start = np.searchsorted(i, np.arange(n + 1))
    print("atoms %d, cells %s, mean full neighbours %.1f" % (n, nc, len(i) / n))

    ncol = nc[0] * nc[1]
    col_start = cell_start[::nc[2]]
    tot = {"cur": np.zeros(2), "bank": np.zeros(2), "steps": 0, "ideal128": 0.0, "ideal64": 0.0}
    nch = 0
    rng = np.random.default_rng(1)
    cols_to_do = rng.permutation(ncol)
    for q in cols_to_do:
        cx, cy = q // nc[1], q % nc[1]
        for a0 in range(col_start[q], col_start[q + 1], CH):
            if nch >= a.max_chunks:
                break
            nch += 1
            b0 = min(a0 + CH, col_start[q + 1])
            zlo, zhi = cs[a0, 2], cs[b0 - 1, 2]
            # tile: 9 stencil columns x (one or two) z pieces, tile.cu:k_tile_chunks
            if zhi - zlo + 3 > nc[2]:
                pieces = [(0, nc[2] - 1)]
            elif zlo - 1 < 0:
                pieces = [(zlo - 1 + nc[2], nc[2] - 1), (0, zhi + 1)]
            elif zhi + 1 >= nc[2]:
                pieces = [(zlo - 1, nc[2] - 1), (0, zhi + 1 - nc[2])]
            else:
                pieces = [(zlo - 1, zhi + 1)]
            seg = []                                   # (column 0..8, first slot, end slot, tile offset)
            off = 0
            for xx in (cx - 1, cx, cx + 1):
                for yy in (cy - 1, cy, cy + 1):
                    cbase = ((xx % nc[0]) * nc[1] + (yy % nc[1])) * nc[2]
                    for (za, zb) in pieces:
                        jb, je = cell_start[cbase + za], cell_start[cbase + zb + 1]
                        seg.append((jb, je, off))
                        off += je - jb
                    if len(pieces) == 1:
                        seg.append((0, 0, off))
            ntile = off
            seg = np.array(seg).reshape(9, 2, 3)

            def local(jj, cc):
                s0 = seg[cc, 0]
                s1 = seg[cc, 1]
                in0 = (jj >= s0[:, 0]) & (jj < s0[:, 1])
                return np.where(in0, s0[:, 2] + jj - s0[:, 0], s1[:, 2] + jj - s1[:, 0])

            na = b0 - a0
            cnts = row_start[a0 + 1:b0 + 1] - row_start[a0:b0]
            kmax = int((cnts.max() + 31) // 32 * 32)
            rows = np.full((CH, kmax), ntile, np.int64)      # natural order, sentinel padded
            for t in range(na):
                r0, r1 = row_start[a0 + t], row_start[a0 + t + 1]
                rows[t, :r1 - r0] = local(j[r0:r1], col[r0:r1])
            assert rows[:na].max() <= ntile and rows.min() >= 0
            mypad = np.zeros(CH, np.int64)
            mypad[:na] = (cnts + 31) // 32 * 32

            def banked(row, cnt, q4):
                """bank-aware order of one row: returns the row in NATURAL-READ order, i.e. out[k] is what lane k & 3 reads at
                step (k >> 2) -- so that the same replay code serves both layouts."""
                G = int((cnt + 31) // 32 * 32) // 4          # steps
                out = np.full(kmax, ntile, np.int64)
                free = np.ones(4 * G, bool)
                ent = row[:cnt]
                cls = ent % 16
                spill = []
                nextk = np.zeros(16, np.int64)
                for e in ent:
                    c = int(e % 16)
                    tl, h = c & 3, c >> 2
                    g = 4 * nextk[c] + ((h - q4) & 3)
                    nextk[c] += 1
                    if g < G:
                        out[4 * g + tl] = e
                        free[4 * g + tl] = False
                    else:
                        spill.append(e)
                # surplus entries: first the holes of the sibling class (same lane, same 16-byte bank group: class ^ 8), so that
                # the LDS.128 phase stays conflict free and only the z read can collide; then any hole
                rest = []
                for e in spill:
                    c = int(e % 16) ^ 8
                    tl, h = c & 3, c >> 2
                    g = 4 * nextk[c] + ((h - q4) & 3)
                    if g < G and free[4 * g + tl]:
                        nextk[c] += 1
                        out[4 * g + tl] = e
                        free[4 * g + tl] = False
                    else:
                        rest.append(e)
                holes = np.nonzero(free)[0]
                for e, hpos in zip(rest, holes):
                    out[hpos] = e
                # sentinel slots keep ntile + (their class) so that they never conflict: model them as broadcast-free
                return out, G

            # replay: warps of 8 consecutive atoms; lanes (team, tl); step g reads natural index 4 g + tl
            for layout in ("cur", "bank"):
                for w0 in range(0, CH, 8):
                    if w0 >= na:
                        break
                    teams = range(w0, min(w0 + 8, CH))
                    if layout == "cur":
                        rr = rows[w0:w0 + 8]
                    else:
                        rr = np.stack([banked(rows[t], int(cnts[t]) if t < na else 0, t & 3)[0] for t in teams])
                    gmax = int(mypad[w0:w0 + 8].max()) // 4
                    for g in range(gmax):
                        idx = rr[:, 4 * g:4 * g + 4].copy()                  # [team, lane]
                        active = (4 * g < mypad[w0:w0 + 8])[:, None] & np.ones((1, 4), bool)
                        if layout == "bank":                                  # sentinel of the slot's own class: no conflict
                            sent = idx == ntile
                            tq = (np.arange(w0, w0 + 8) & 3)[:, None]
                            wantc = 4 * ((tq + g) & 3) + np.arange(4)[None, :]
                            idx = np.where(sent, (1 << 20) + wantc, idx)       # a sentinel slot in the bank of the class the lane wants
                        w128 = 0
                        for ph in range(4):                                   # quarter-warp phases: teams 2 ph, 2 ph + 1
                            v = idx[2 * ph:2 * ph + 2][active[2 * ph:2 * ph + 2]]
                            if len(v):
                                u = np.unique(v)
                                w128 += np.bincount(u % 8, minlength=8).max()
                        w64 = 0
                        for ph in range(2):                                   # half-warp phases: teams 4 ph .. 4 ph + 3
                            v = idx[4 * ph:4 * ph + 4][active[4 * ph:4 * ph + 4]]
                            if len(v):
                                u = np.unique(v)
                                w64 += np.bincount(u % 16, minlength=16).max()
                        tot[layout] += (w128, w64)
                        if layout == "cur":
                            tot["steps"] += 1
        if nch >= a.max_chunks:
            break
    s = tot["steps"]
    print("chunks %d, warp steps %d" % (nch, s))
    for layout in ("cur", "bank"):
        print("%-5s LDS.128 %.2f wavefronts per warp step, LDS.64 %.2f, together %.2f"
              % (layout, tot[layout][0] / s, tot[layout][1] / s, tot[layout].sum() / s))


if __name__ == "__main__":
    main()
