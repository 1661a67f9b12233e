// Grid (trackers.hpp:227-309) and asynchronous trajectory frames through the C++ drop-in headers.
// Prints "OK" when (1) the device binning equals Grid::get_loc evaluated on the host for every atom, (2) all_pairs()
// holds every pair closer than a cell width, (3) a frame started before further timestep() calls holds the state of the
// moment it was started.
#include <cstdio>
#include <cstdlib>
#include "collection.hpp"

int main() {
    const uint N = 1000;
    const flt L = 12.04;
    sptr<OriginBox> box(new OriginBox(L));
    sptr<AtomVec> atomptr(new AtomVec(N, 1.0));
    AtomVec &atoms = *atomptr;
    seed(7);
    uint side = 10;
    for (uint i = 0; i < N; i++) {
        Vec x;
        uint k[3] = {i % side, (i / side) % side, i / (side * side)};
        for (uint d = 0; d < NDIM; d++) x[d] = (k[d] + 0.5) * L / side + 0.05 * rand_vec()[d] - 3 * L * (i % 3);
        atoms[i].x = x;
        atoms[i].v = rand_vec();
    }
    Grid grid(box, atomptr, 4);
    grid.make_grid();
    for (uint i = 0; i < N; i++)
        if (grid.locs[i] != grid.get_loc(atoms[i].x, box->box_shape())) { printf("FAIL loc %u\n", i); return 1; }
    vector<IDPair> ps = grid.all_pairs();
    std::set<std::pair<uint, uint> > have;
    for (size_t k = 0; k < ps.size(); k++) have.insert(std::make_pair(ps[k].first().n(), ps[k].last().n()));
    const flt cw = L / 4;
    for (uint i = 0; i < N; i++)
        for (uint j = 0; j < i; j++)
            if (box->diff(atoms[i].x, atoms[j].x).norm() < 0.999 * cw && !have.count(std::make_pair(i, j))) { printf("FAIL pair %u %u\n", i, j); return 1; }
    // a short LJ run with a frame taken in flight
    sptr<NeighborList> nl(new NeighborList(box, atomptr, 0.4));
    sptr<NListed<EpsSigCutAtom, LennardJonesCutPair> > lj(new NListed<EpsSigCutAtom, LennardJonesCutPair>(atomptr, nl));
    for (uint i = 0; i < N; i++) lj->add(EpsSigCutAtom(atoms.get_id(i), 1.0, 1.0, 2.5));
    CollectionVerlet collec(box, atomptr, 1e-3);
    collec.add_interaction(lj);
    collec.add_tracker(nl);
    for (int s = 0; s < 20; s++) collec.timestep();
    vector<Vec> x0(N), v0(N), xs(N), vs(N);
    for (uint i = 0; i < N; i++) { x0[i] = atoms.read(i).x; v0[i] = atoms.read(i).v; }
    atoms.snapshot_begin();
    for (int s = 0; s < 50; s++) collec.timestep();
    atoms.snapshot_wait(xs.data(), vs.data());
    flt moved = 0;
    for (uint i = 0; i < N; i++) {
        if ((xs[i] - x0[i]).norm() != 0 || (vs[i] - v0[i]).norm() != 0) { printf("FAIL frame %u\n", i); return 1; }
        moved = std::max(moved, (atoms.read(i).x - x0[i]).norm());
    }
    if (!(moved > 1e-4)) { printf("FAIL the run did not move on\n"); return 1; }
    printf("OK pairs %zu moved %g\n", ps.size(), (double)moved);
    return 0;
}
