// parm_b200 drop-in for the parts of ParM's src/constraints.hpp that run on the device: the Constraint
// interface (constraints.hpp:15-22; concrete constraints are out of scope) and the statistics
// trackers RsqTracker, ISFTracker, EnergyTracker (constraints.hpp:260-414), whose accumulators live in device
// memory (csrc/trackers.cu): add_tracker()ing them to a Collection costs no host synchronisation per step.
// Eigen's N x NDIM tables are returned as parm_b200::Table (rows(), cols(), operator()(i, j), row(i)).
#include "interaction.hpp"
#ifndef PARM_B200_CONSTRAINTS_H
#define PARM_B200_CONSTRAINTS_H
class Constraint {
   public:
    virtual void apply_positions(Box &box) = 0;
    virtual void apply_velocities(Box &box) = 0;
    virtual void apply_forces(Box &box) = 0;
    virtual int constrained_dof() = 0;
    virtual ~Constraint() {}
};

template <class T, unsigned int N>
struct barray {  // boost::array stand-in (vecrand.hpp:53)
    T elems[N];
    barray() { for (unsigned int i = 0; i < N; i++) elems[i] = T(); }
    T &operator[](unsigned int i) { return elems[i]; }
    const T &operator[](unsigned int i) const { return elems[i]; }
    unsigned int size() const { return N; }
};

namespace parm_b200 {
// stand-in for Eigen::Matrix<flt, Eigen::Dynamic, NDIM>, row-major
class Table {
    vector<flt> d;
    uint r;

   public:
    Table() : r(0) {}
    Table(uint rows, uint) : d((size_t)rows * NDIM, 0.0), r(rows) {}
    uint rows() const { return r; }
    uint cols() const { return NDIM; }
    flt &operator()(uint i, uint j) { return d[(size_t)i * NDIM + j]; }
    const flt &operator()(uint i, uint j) const { return d[(size_t)i * NDIM + j]; }
    Vec row(uint i) const { Vec v; for (uint j = 0; j < NDIM; j++) v[j] = (*this)(i, j); return v; }
    flt *data() { return d.data(); }
};

// a statistics tracker whose update is a device kernel; Collection hands its handle to the integrator
class DeviceTracker : public StateTracker {
   protected:
    sptr<AtomVec> av;
    parm_tracker *trk;
    uint nlags;

   public:
    DeviceTracker(sptr<AtomGroup> atoms) : av(boost::dynamic_pointer_cast<AtomVec>(atoms)), trk(NULL), nlags(0) {
        if (!av) throw std::runtime_error("parm_b200: statistics trackers need an AtomVec");
    }
    ~DeviceTracker() { parm_tracker_destroy(trk); }
    parm_tracker *tracker_handle() { return trk; }
    void update(Box &) {
        av->device(false);
        check(parm_tracker_update(trk));
    }
    void reset() {
        av->device(false);
        check(parm_tracker_reset(trk));
    }
    vector<flt> counts() {
        vector<uint64_t> c(nlags ? nlags : 1);
        check(parm_tracker_counts(trk, c.data(), (int)nlags));
        vector<flt> out(nlags);
        for (uint k = 0; k < nlags; k++) out[k] = (flt)c[k];
        return out;
    }
};
}  // namespace parm_b200

class RsqTracker : public parm_b200::DeviceTracker {  // constraints.hpp:342-368
    void read(uint k, parm_b200::Table *a, parm_b200::Table *b, vector<flt> *c) {
        parm_b200::check(parm_rsq_read(trk, (int)k, a ? a->data() : NULL, b ? b->data() : NULL, c ? c->data() : NULL));
    }

   public:
    RsqTracker(sptr<AtomGroup> atoms, vector<unsigned long> ns, bool usecom = true) : DeviceTracker(atoms) {
        vector<uint64_t> v(ns.begin(), ns.end());
        nlags = (uint)v.size();
        parm_b200::check(parm_rsq_create(av->device(false), v.data(), (int)v.size(), usecom ? 1 : 0, &trk));
    }
    vector<parm_b200::Table> xyz2() {
        vector<parm_b200::Table> out;
        for (uint k = 0; k < nlags; k++) {
            parm_b200::Table t(av->size(), NDIM);
            read(k, &t, NULL, NULL);
            out.push_back(t);
        }
        return out;
    }
    vector<parm_b200::Table> xyz4() {
        vector<parm_b200::Table> out;
        for (uint k = 0; k < nlags; k++) {
            parm_b200::Table t(av->size(), NDIM);
            read(k, NULL, &t, NULL);
            out.push_back(t);
        }
        return out;
    }
    vector<vector<flt> > r2() {  // constraints.cpp:520-535
        vector<parm_b200::Table> x = xyz2();
        vector<vector<flt> > out;
        for (uint k = 0; k < nlags; k++) {
            vector<flt> v(av->size());
            for (uint i = 0; i < av->size(); i++) v[i] = x[k].row(i).sum();
            out.push_back(v);
        }
        return out;
    }
    vector<vector<flt> > r4() {
        vector<vector<flt> > out;
        for (uint k = 0; k < nlags; k++) {
            vector<flt> v(av->size());
            read(k, NULL, NULL, &v);
            out.push_back(v);
        }
        return out;
    }
};

class ISFTracker : public parm_b200::DeviceTracker {  // constraints.hpp:393-414
    uint nks;

   public:
    ISFTracker(sptr<AtomGroup> atoms, vector<flt> ks, vector<unsigned long> ns, bool usecom = false)
        : DeviceTracker(atoms), nks((uint)ks.size()) {
        vector<uint64_t> v(ns.begin(), ns.end());
        nlags = (uint)v.size();
        parm_b200::check(parm_isf_create(av->device(false), ks.data(), (int)ks.size(), v.data(), (int)v.size(), usecom ? 1 : 0, &trk));
    }
    // [lag][k][atom][axis]
    vector<vector<vector<barray<cmplx, NDIM> > > > ISFxyz() {
        vector<vector<vector<barray<cmplx, NDIM> > > > out(nlags);
        const uint n = av->size();
        vector<flt> buf((size_t)nks * n * NDIM * 2 + 1);
        for (uint l = 0; l < nlags; l++) {
            parm_b200::check(parm_isf_read(trk, (int)l, buf.data()));
            out[l].assign(nks, vector<barray<cmplx, NDIM> >(n));
            for (uint k = 0; k < nks; k++)
                for (uint i = 0; i < n; i++)
                    for (uint j = 0; j < NDIM; j++) {
                        const size_t q = ((((size_t)k * n + i) * NDIM) + j) * 2;
                        out[l][k][i][j] = cmplx(buf[q], buf[q + 1]);
                    }
        }
        return out;
    }
    // [lag][k][atom]: mean over the axes (constraints.cpp:621-635)
    vector<vector<vector<cmplx> > > ISFs() {
        vector<vector<vector<barray<cmplx, NDIM> > > > x = ISFxyz();
        vector<vector<vector<cmplx> > > out(nlags);
        for (uint l = 0; l < nlags; l++) {
            out[l].assign(nks, vector<cmplx>(av->size(), cmplx(0, 0)));
            for (uint k = 0; k < nks; k++)
                for (uint i = 0; i < av->size(); i++) {
                    for (uint j = 0; j < NDIM; j++) out[l][k][i] += x[l][k][i][j];
                    out[l][k][i] /= NDIM;
                }
        }
        return out;
    }
};

class EnergyTracker : public parm_b200::DeviceTracker {  // constraints.hpp:260-316
    vector<sptr<Interaction> > interactions;
    void sums(flt *s8) {
        av->device(false);
        parm_b200::check(parm_energy_tracker_read(trk, s8));
    }
    flt mean(int q) { flt s[8]; sums(s); return s[q] / s[0]; }
    flt stdev(int q1, int q2) { flt s[8]; sums(s); return sqrt(s[q2] / s[0] - s[q1] * s[q1] / s[0] / s[0]); }

   public:
    EnergyTracker(sptr<AtomGroup> atoms, vector<sptr<Interaction> > inters, uint n_skip = 1)
        : DeviceTracker(atoms), interactions(inters) {
        vector<parm_inter *> hs;
        for (size_t k = 0; k < inters.size(); k++) {
            parm_b200::DeviceInteraction *d = dynamic_cast<parm_b200::DeviceInteraction *>(inters[k].get());
            if (!d) throw std::runtime_error("parm_b200: EnergyTracker needs NListed interactions");
            hs.push_back(d->device_handle());
        }
        parm_b200::check(parm_energy_tracker_create(av->device(false), hs.empty() ? NULL : hs.data(), (int)hs.size(), n_skip, &trk));
    }
    void set_U0(flt newU0) { parm_b200::check(parm_energy_tracker_set_u0(trk, 0, newU0)); }
    void set_U0(Box &) {
        av->device(false);
        parm_b200::check(parm_energy_tracker_set_u0(trk, 1, 0.0));
    }
    flt get_U0() { flt s[8]; sums(s); return s[7]; }
    flt E() { return mean(1); }
    flt U() { return mean(2); }
    flt K() { return mean(3); }
    flt E_std() { return stdev(1, 4); }
    flt U_std() { return stdev(2, 5); }
    flt K_std() { return stdev(3, 6); }
    flt E_squared_mean() { return mean(4); }
    flt U_squared_mean() { return mean(5); }
    flt K_squared_mean() { return mean(6); }
    uint n() { flt s[8]; sums(s); return (uint)s[0]; }
};
#endif
