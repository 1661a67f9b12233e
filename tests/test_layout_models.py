"""CPU-only models of two layout arguments the CUDA code relies on (no GPU, no library call): they restate, in numpy,
what the kernels assume, so that a change of either layout that breaks the assumption fails here first.

1. Pair kernel, partial last pass (csrc/force_tile.cuh: tile_rows / csrc/tile.cu: k_tile_localize*): entry k of a row
   is stored at (k / 32) * 32 + (k % 4) * 8 + (k % 32) / 4, i.e. lane tl of the atom's 4-lane team reads, as element e
   of its 8-entry vector of pass p, the row entry 32 p + 4 e + tl. When the LONGEST row of a warp has at most 4 m entries
   left in a pass, every real entry of every team sits in elements 0 .. m-1: the pass may stop after m elements.
2. Slab rebuild with one sort (csrc/shard.cu): sorting [owned slots in their old order | arrivals] stably by the new
   cell id, with the leavers binned into the halo layers (numbered last), leaves the owned atoms in exactly the order of
   the two-sort path (stable sort of the owned set -> drop the leavers -> append arrivals -> stable sort again)."""
import numpy as np


def stored_position(k):
    return (k // 32) * 32 + (k % 4) * 8 + (k % 32) // 4


def test_lane_vector_layout_is_a_permutation_of_each_pass():
    k = np.arange(160)
    pos = stored_position(k)
    assert sorted(pos.tolist()) == k.tolist()
    assert np.all(pos // 32 == k // 32)  # an entry never leaves its pass


def test_partial_last_pass_holds_every_real_entry():
    rng = np.random.default_rng(0)
    sentinel = 0xFFFF
    for _ in range(300):
        lengths = rng.integers(0, 161, size=8)  # the 8 rows a warp walks in lock step
        longest = int(lengths.max())
        rows = []
        for n in lengths:
            npad = -(-int(n) // 32) * 32
            r = np.full(max(npad, 32), sentinel, np.int64)
            r[stored_position(np.arange(n))] = np.arange(n)  # real entries carry their own index
            rows.append(r)
        for k0 in range(0, longest, 32):
            rem = longest - k0
            m = 8 if rem > 24 else 6 if rem > 16 else 4 if rem > 8 else 2  # the kernel's choice (steps of two)
            for r, n in zip(rows, lengths):
                if k0 >= len(r):
                    continue  # the team ran out of passes: it walks sentinel words
                for tl in range(4):
                    vec = r[k0 + tl * 8:k0 + tl * 8 + 8]
                    assert np.all(vec[m:] == sentinel), (lengths, k0, m, tl, vec)
                    real = vec[:m][vec[:m] != sentinel]
                    assert np.all(real == k0 + 4 * np.nonzero(vec[:m] != sentinel)[0] + tl)


def _stable_sort_by(cell, order):
    return order[np.argsort(cell[order], kind="stable")]


def test_one_sort_rebuild_leaves_the_two_sort_order():
    rng = np.random.default_rng(1)
    ncell_interior, halo_up, halo_dn = 40, 40, 41  # halo layers are numbered behind the interior cells
    for _ in range(200):
        n_old = int(rng.integers(20, 200))
        cell_old = rng.integers(0, ncell_interior, n_old)
        leaver = rng.random(n_old) < 0.1
        cell_old = np.where(leaver, rng.choice([halo_up, halo_dn], n_old), cell_old)
        r = int(rng.integers(0, 12))
        cell_arr = rng.integers(0, ncell_interior, r)
        cell = np.concatenate([cell_old, cell_arr])
        ids = np.arange(n_old + r)  # slot n_old + q holds arrival q (from up first, then from down: any fixed order)
        # two sorts: sort the owned set, keep what stayed, append the arrivals, sort again
        first = _stable_sort_by(cell, ids[:n_old])
        keep = first[cell[first] < ncell_interior]
        two = _stable_sort_by(cell, np.concatenate([keep, ids[n_old:]]))
        # one sort over everything; the leavers land behind the owned set and are dropped
        one = _stable_sort_by(cell, ids)
        n_new = int((cell < ncell_interior).sum())
        assert np.array_equal(one[:n_new], two)
        assert np.all(cell[one[n_new:]] >= ncell_interior)
        # ... leavers upwards first, then downwards: where the two-sort path's phase 1 left them
        assert np.all(np.diff(cell[one[n_new:]]) >= 0)
