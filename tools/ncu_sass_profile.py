#!/usr/bin/env python
"""Reads an `ncu --page source --csv` export (SASS view) and prints where the warp instructions, the stall samples
and the shared-memory wavefronts of one kernel go: by opcode, by execution-count group (= loop nest level) and
by region between barriers.   python tools/ncu_sass_profile.py gpurun_out/<name>_source.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    col = {c: hdr.index(c) for c in ("Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared",
                                     "L1 Wavefronts Shared Ideal")}

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0

    ins = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        ins.append((num(r[col["Instructions Executed"]]), num(r[col["# Samples"]]), r[col["Source"]].strip(),
                    num(r[col["L1 Wavefronts Shared"]]), num(r[col["L1 Wavefronts Shared Ideal"]])))
    itot = sum(i[0] for i in ins) or 1
    stot = sum(i[1] for i in ins) or 1
    print("warp instructions %d, stall samples %d, shared wavefronts %d (ideal %d)"
          % (itot, stot, sum(i[3] for i in ins), sum(i[4] for i in ins)))
    byop, sop = collections.Counter(), collections.Counter()
    for n, s, t, _, _ in ins:
        tt = t.split()
        if not tt:
            continue
        op = tt[1] if tt[0].startswith("@") and len(tt) > 1 else tt[0]
        byop[op] += n
        sop[op] += s
    print("-- by opcode")
    for op, n in byop.most_common(18):
        print("  %-22s %12d %5.1f %% of instructions, %5.1f %% of samples" % (op, n, 100.0 * n / itot, 100.0 * sop[op] / stot))
    grp = collections.defaultdict(lambda: [0, 0, 0])
    for n, s, _, _, _ in ins:
        grp[n][0] += 1
        grp[n][1] += n
        grp[n][2] += s
    print("-- by execution count (loop level)")
    for n, (c, t, s) in sorted(grp.items(), key=lambda kv: -kv[1][1])[:10]:
        print("  executed %10d times: %4d instructions = %5.1f %% of instructions, %5.1f %% of samples"
              % (n, c, 100.0 * t / itot, 100.0 * s / stot))
    bars = [i for i, x in enumerate(ins) if x[2].startswith("BAR")]
    print("-- by region between barriers (SASS order)")
    edges = [0] + [b + 1 for b in bars] + [len(ins)]
    for a, b in zip(edges[:-1], edges[1:]):
        if b > a:
            print("  SASS lines %4d..%4d: %5.1f %% of instructions, %5.1f %% of samples"
                  % (a, b - 1, 100.0 * sum(x[0] for x in ins[a:b]) / itot, 100.0 * sum(x[1] for x in ins[a:b]) / stot))
    print("-- shared-memory instructions with the most wavefronts")
    for n, s, t, w, wi in sorted(ins, key=lambda x: -x[3])[:8]:
        if w:
            print("  %-40s executed %9d  wavefronts %9d  ideal %9d  (%.2fx)" % (t[:40], n, w, wi, w / max(wi, 1)))


if __name__ == "__main__":
    main(sys.argv[1])
