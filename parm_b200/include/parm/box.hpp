// parm_b200 drop-in for ParM's src/box.hpp: OriginBox, Atom, AtomID, AtomGroup, AtomVec.
// Same public names and semantics as the reference (box.hpp:48-158, 228-479); `AtomVec` owns a
// device context (include/parm_b200.h) and keeps the reference's AoS `Atom[N]` as a page-locked
// host mirror so that `Atom&` / `Atom*` handed to user code stay valid. Coherence is lazy, with
// two coarse flags (SURVEY 8b "Mutation model"):
//   - any non-const element access (operator[], get, get_id, iteration) brings the mirror up to
//     date and marks it "possibly modified": it is uploaded before the next device operation;
//   - timestep(), set_forces(), velocity scaling ... mark the device copy as newer.
// Group reductions (kinetic_energy, momentum, ...) run on the device and never force a download.
#include <stdexcept>
#include <cstdlib>
#include "vecrand.hpp"

#ifndef PARM_B200_BOX_H
#define PARM_B200_BOX_H

#include <string>
#include <vector>

#include "parm_b200.h"

#define sptr boost::shared_ptr
typedef const unsigned int cuint;

namespace parm_b200 {
// C status -> the reference's exception types (SURVEY 8b "Error convention")
inline void check(int rc) {
    if (rc == PARM_OK) return;
    std::string msg = parm_b200_last_error();
    if (rc == PARM_ERR_INVALID) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
}  // namespace parm_b200

class AtomGroup;
class AtomVec;

class Box {
   public:
    virtual Vec diff(Vec r1, Vec r2) = 0;
    virtual flt V() = 0;
    virtual ~Box() {}
};

// vec_mod, box.hpp:69-78 (host-side convenience for single vectors; the hot path evaluates the
// same IEEE remainder inside the kernels)
inline Vec vec_mod(Vec r1, Vec r2) {
    Vec o;
    for (uint i = 0; i < NDIM; i++) o[i] = remainder(r1[i], r2[i]);
    return o;
}

class OriginBox : public Box {
   protected:
    Vec boxsize;
    std::vector<parm_ctx *> ctxs;  // device contexts using this box
    void push() {
        for (size_t k = 0; k < ctxs.size(); k++) parm_b200::check(parm_set_box(ctxs[k], boxsize.data()));
    }

   public:
    OriginBox(Vec size) : boxsize(size) {}
#ifdef VEC3D
    OriginBox(flt L) : boxsize(L, L, L) {}
    flt V() { return boxsize[0] * boxsize[1] * boxsize[2]; }
    flt L() { return (boxsize[0] + boxsize[1] + boxsize[2]) / 3.0; }
#else
    OriginBox(flt L) : boxsize(L, L) {}
    flt V() { return boxsize[0] * boxsize[1]; }
    flt L() { return (boxsize[0] + boxsize[1]) / 2.0; }
#endif
    Vec diff(Vec r1, Vec r2) { return vec_mod((r1 - r2), boxsize); }
    Vec box_shape() { return boxsize; }
    Vec rand_loc() {  // box.hpp:145-151
        Vec v = rand_vec_boxed();
        for (uint i = 0; i < NDIM; i++) v[i] *= boxsize[i];
        return diff(v, Vec::Zero());
    }
    // resize without moving atoms (box.cpp:3-7, 21-25, 34-39, 46-50)
    flt resize_to(Vec newsize) { boxsize = newsize; push(); return V(); }
    flt resize(flt factor) { boxsize *= factor; push(); return V(); }
    flt resize_to_V(flt newV) { return resize(pow(newV / V(), OVERNDIM)); }
    flt resize_to_L(flt newL) { return resize(newL / L()); }
    // used by AtomVec-bound classes: this box now also describes that device context
    void attach(parm_ctx *c) {
        for (size_t k = 0; k < ctxs.size(); k++)
            if (ctxs[k] == c) return;
        ctxs.push_back(c);
        parm_b200::check(parm_set_box(c, boxsize.data()));
    }
    // the device changed the box of context c (CollectionNLCG::stepx): refresh boxsize and the other contexts
    void pull(parm_ctx *c) {
        flt L[3] = {0, 0, 0};
        parm_b200::check(parm_get_box(c, L));
        for (uint i = 0; i < NDIM; i++) boxsize[i] = L[i];
        for (size_t k = 0; k < ctxs.size(); k++)
            if (ctxs[k] != c) parm_b200::check(parm_set_box(ctxs[k], boxsize.data()));
    }
    void detach(parm_ctx *c) {
        for (size_t k = 0; k < ctxs.size(); k++)
            if (ctxs[k] == c) { ctxs.erase(ctxs.begin() + k); return; }
    }
};

struct Atom {  // box.hpp:234-249, identical layout
    Vec x;
    Vec v;
    Vec a;
    Vec f;
    flt m;
};

class AtomRef {
   private:
    Atom *ptr;

   public:
    inline AtomRef() : ptr(NULL) {}
    inline AtomRef(Atom *a) : ptr(a) {}
    inline Atom &operator*() const { return *ptr; }
    inline Atom *operator->() const { return ptr; }
    inline bool operator==(const AtomRef &other) const { return other.ptr == ptr; }
    inline bool operator==(const Atom *other) const { return other == ptr; }
    inline bool operator!=(const AtomRef &other) const { return other.ptr != ptr; }
    inline bool operator<(const AtomRef &other) const { return ptr < other.ptr; }
    inline bool operator<=(const AtomRef &other) const { return ptr <= other.ptr; }
    inline bool operator>=(const AtomRef &other) const { return ptr >= other.ptr; }
    inline bool operator>(const AtomRef &other) const { return ptr > other.ptr; }
    inline bool is_null() { return ptr == NULL; }
};

class AtomID : public AtomRef {
   private:
    uint num;

   public:
    inline AtomID() : AtomRef(), num(UINT_MAX) {}
    inline AtomID(Atom *a, uint n) : AtomRef(a), num(n) {}
    inline uint n() const { return num; }
};

class IDPair {
   private:
    AtomID id1, id2;

   public:
    IDPair() : id1(), id2() {}
    IDPair(AtomID a, AtomID b) : id1(a), id2(b) {}
    inline AtomID first() const { return id1; }
    inline AtomID last() const { return id2; }
};

class AtomIter {
   private:
    uint i;
    AtomGroup &g;

   public:
    AtomIter(AtomGroup &g, uint i) : i(i), g(g) {}
    bool operator!=(const AtomIter &other) const { return i != other.i; }
    Atom &operator*() const;
    inline const AtomIter &operator++() { ++i; return *this; }
};

// AtomGroup interface (box.hpp:336-433); the reductions are implemented by AtomVec on the device.
class AtomGroup {
   public:
    virtual AtomVec &vec() = 0;
    virtual Atom &operator[](cuint n) = 0;
    virtual Atom &operator[](cuint n) const = 0;
    virtual Atom &get(cuint n) { return ((*this)[n]); }
    virtual AtomID get_id(cuint n) = 0;
    virtual uint size() const = 0;
    virtual AtomIter begin() { return AtomIter(*this, 0); }
    virtual AtomIter end() { return AtomIter(*this, (uint)size()); }
    virtual Vec com() const = 0;
    virtual Vec com_force() const = 0;
    virtual Vec com_velocity() const = 0;
    virtual flt mass() const = 0;
    virtual flt kinetic_energy(const Vec originvelocity = Vec::Zero()) const = 0;
    virtual Vec momentum() const = 0;
    virtual void add_velocity(Vec v) = 0;
    void reset_com_velocity() { add_velocity(-com_velocity()); }
    virtual void randomize_velocities(flt T) = 0;
    virtual void reset_forces() = 0;
    virtual ~AtomGroup() {}
};
inline Atom &AtomIter::operator*() const { return g[i]; }

class AtomVec : public virtual AtomGroup {
   private:
    Atom *atoms;
    uint sz;
    parm_ctx *ctx;
    bool registered;
    mutable bool host_dirty;  // mirror may hold changes the device has not seen
    mutable bool dev_newer;   // device holds results the mirror has not seen
    // STORED REFERENCES (difference from the reference, where Atom& / AtomID are live views of the only copy): the host
    // array is a MIRROR. operator[] / get_id() synchronise at the moment of the call; an `Atom&` or `AtomID` kept across a
    // device operation (timestep(), set_forces(), update_list(), a reduction after one of those) goes stale: reads
    // through it see the state at the time of the call, writes through it are never uploaded. Re-fetch with operator[]
    // after device work, or bracket the access with sync_to_host() / touch_host(). PARM_B200_CHECK_VIEWS=1 (environment)
    // turns on a checksum of the mirror that makes device() throw std::logic_error when it finds such a lost write.
    bool check_views;
    mutable unsigned long long mirror_sum;
    unsigned long long checksum() const { // FNV-1a over the mirror (debug aid only)
        unsigned long long h = 1469598103934665603ull;
        const unsigned char *p = reinterpret_cast<const unsigned char *>(atoms);
        for (size_t i = 0, n = sizeof(Atom) * (size_t)sz; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
    void init(int device) {
        const char *cv = getenv("PARM_B200_CHECK_VIEWS");
        check_views = cv && atoi(cv) != 0;
        mirror_sum = 0;
        atoms = new Atom[sz ? sz : 1];
        for (uint i = 0; i < sz; i++) {
            atoms[i].x = Vec::Zero();
            atoms[i].v = Vec::Zero();
            atoms[i].f = Vec::Zero();
            atoms[i].a = Vec::Zero();
        }
        ctx = NULL;
        parm_b200::check(parm_ctx_create(NDIM, sz, device, &ctx));
        registered = sz && parm_host_register(atoms, sizeof(Atom) * (size_t)sz) == PARM_OK;
        host_dirty = true;
        dev_newer = false;
    }
    AtomVec &operator=(const AtomVec &);

   public:
    AtomVec(vector<double> masses, int device = 0) : sz((uint)masses.size()) {
        init(device);
        for (uint i = 0; i < sz; i++) atoms[i].m = masses[i];
    }
    AtomVec(uint N, flt mass, int device = 0) : sz(N) {
        init(device);
        for (uint i = 0; i < sz; i++) atoms[i].m = mass;
    }
    AtomVec(AtomVec &other) : sz(other.size()) {
        init(0);
        other.sync_to_host();
        for (uint i = 0; i < sz; i++) atoms[i] = other.atoms[i];
    }
    ~AtomVec() {
        if (registered) parm_host_unregister(atoms);
        parm_ctx_destroy(ctx);
        delete[] atoms;
    }

    // ---- coherence (explicit escape hatches are public) ----
    void sync_to_host() const {
        if (dev_newer && sz)
            parm_b200::check(parm_download_atoms(ctx, PARM_ALL, atoms[0].x.data(), atoms[0].v.data(), atoms[0].a.data(),
                                                 atoms[0].f.data(), &atoms[0].m, sizeof(Atom), sizeof(Atom)));
        if (dev_newer && check_views) mirror_sum = checksum();
        dev_newer = false;
    }
    void sync_to_device() const {
        if (!host_dirty && !dev_newer && check_views && sz && checksum() != mirror_sum)
            throw std::logic_error("AtomVec: the host mirror changed without operator[] / get_id() / touch_host(): an Atom& or "
                                   "AtomID kept across a device operation was written through (parm/box.hpp, STORED REFERENCES)");
        if (host_dirty && sz) {
            parm_b200::check(parm_upload_atoms(ctx, PARM_ALL, atoms[0].x.data(), atoms[0].v.data(), atoms[0].a.data(),
                                               atoms[0].f.data(), &atoms[0].m, sizeof(Atom), sizeof(Atom)));
            if (check_views) mirror_sum = checksum();
        }
        host_dirty = false;
    }
    // called by every facade class before it launches device work
    parm_ctx *device(bool modifies = true) const {
        sync_to_device();
        if (modifies) dev_newer = true;
        return ctx;
    }
    parm_ctx *context() const { return ctx; }
    void touch_host() const {
        sync_to_host();
        host_dirty = true;
    }

    // Asynchronous trajectory frame (the writefile() of LJatoms.cpp:130-158 without stalling the run): snapshot_begin()
    // starts the device->host copy of x (and v) on a second stream, timestep() calls made before snapshot_wait()
    // overlap it; snapshot_wait() fills x[i], v[i] (AtomVec order; v may be NULL).
    void snapshot_begin(bool velocities = true) { parm_b200::check(parm_snapshot_begin(device(false), PARM_X | (velocities ? PARM_V : 0u))); }
    void snapshot_wait(Vec *x, Vec *v = NULL) {
        vector<double> bx((size_t)sz * NDIM + 1), bv((size_t)sz * NDIM + 1);
        parm_b200::check(parm_snapshot_wait(ctx, x ? bx.data() : NULL, v ? bv.data() : NULL));
        for (uint i = 0; i < sz; i++)
            for (uint d = 0; d < NDIM; d++) {
                if (x) x[i][d] = bx[(size_t)i * NDIM + d];
                if (v) v[i][d] = bv[(size_t)i * NDIM + d];
            }
    }
    // read-only access: refreshes the mirror if the device is ahead, but does NOT mark it modified, so that a read
    // followed by a timestep() does not upload the whole array again (operator[] has to assume a write)
    const Atom &read(cuint n) const { sync_to_host(); return atoms[n]; }
    AtomVec &vec() { return *this; }
    inline Atom &operator[](cuint n) { touch_host(); return atoms[n]; }
    inline Atom &operator[](cuint n) const { touch_host(); return atoms[n]; }
    inline AtomID get_id(cuint n) {
        if (n > sz) return AtomID();  // sic: box.hpp:473
        touch_host();                 // the AtomID carries a raw Atom* the caller may read or write through
        return AtomID(atoms + n, n);
    }
    inline uint size() const { return sz; }

    // ---- AtomGroup reductions on the device (box.cpp:239-260, 401-431) ----
    Vec reduce_vec(int what) const {
        flt out[4] = {0, 0, 0, 0};
        parm_b200::check(parm_reduce(device(false), what, NULL, out));
        Vec v;
        for (uint i = 0; i < NDIM; i++) v[i] = out[i];
        return v;
    }
    Vec com() const { return reduce_vec(PARM_RED_COM); }
    Vec com_force() const { return reduce_vec(PARM_RED_COMFORCE); }
    Vec momentum() const { return reduce_vec(PARM_RED_MOMENTUM); }
    Vec com_velocity() const { return momentum() / mass(); }
    flt mass() const {
        flt out[4];
        parm_b200::check(parm_reduce(device(false), PARM_RED_MASS, NULL, out));
        return out[0];
    }
    flt kinetic_energy(const Vec originvelocity = Vec::Zero()) const {
        flt out[4];
        parm_b200::check(parm_reduce(device(false), PARM_RED_KE, originvelocity.data(), out));
        return out[0];
    }
    flt mobile_dof() const {
        flt out[4];
        parm_b200::check(parm_reduce(device(false), PARM_RED_NDOF, NULL, out));
        return out[0];
    }
    void add_velocity(Vec v) { parm_b200::check(parm_add_velocity(device(), v.data())); }
    void reset_forces() { parm_b200::check(parm_reset_forces(device())); }
    void randomize_velocities(flt T) {  // box.cpp:419-425 (host RNG, then uploaded lazily)
        touch_host();
        for (uint i = 0; i < sz; i++) {
            Atom &a = atoms[i];
            if (a.m == 0 or isinf(a.m)) continue;
            a.v = rand_vec() * sqrt(T / a.m);
        }
    }
};

#endif
