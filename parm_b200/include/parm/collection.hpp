// parm_b200 drop-in for ParM's src/collection.hpp -- Collection, CollectionVerlet, CollectionSol
// (collection.hpp:23-130, 205-263, 360-374; collection.cpp:3-208, 210-322, 442-469) and the other fixed-box
// integrators that only need set_forces(): CollectionDamped, SolHT, Overdamped, NoseHoover, GaussianT,
// Gear3A-6A (collection.hpp:273-353, 376-393, 567-755).
// timestep() enqueues K1 -> force kernel(s) -> K3+drift on the device and returns; nothing is
// copied to the host until user code touches an Atom or asks for a scalar.
#ifndef PARM_B200_COLLECTION_H
#define PARM_B200_COLLECTION_H

#include <cstdio>
#include <set>

#include "constraints.hpp"

class Collection {
   protected:
    sptr<Box> box;
    sptr<AtomGroup> atoms;
    vector<sptr<Interaction> > interactions;
    vector<sptr<StateTracker> > trackers;
    vector<sptr<Constraint> > constraints;
    AtomVec *av;
    parm_integ *integ;

    static parm_inter *dev(sptr<Interaction> &i) {
        parm_b200::DeviceInteraction *d = dynamic_cast<parm_b200::DeviceInteraction *>(i.get());
        if (!d) throw std::runtime_error("parm_b200: only NListed interactions of the in-scope pair types run on the device (no CPU fallback)");
        return d->device_handle();
    }
    static parm_tracker *stat(sptr<StateTracker> &t) {  // RsqTracker / ISFTracker / EnergyTracker, else NULL
        parm_b200::DeviceTracker *d = dynamic_cast<parm_b200::DeviceTracker *>(t.get());
        return d ? d->tracker_handle() : NULL;
    }
    static parm_nlist *dev(sptr<StateTracker> &t) {
        NeighborList *n = dynamic_cast<NeighborList *>(t.get());
        if (!n) throw std::runtime_error("parm_b200: only NeighborList, RsqTracker, ISFTracker and EnergyTracker run on the device");
        return n->handle();
    }
    parm_ctx *ready(bool modifies = true) {
        for (size_t k = 0; k < interactions.size(); k++) dev(interactions[k]);
        for (size_t k = 0; k < trackers.size(); k++)
            if (!stat(trackers[k])) dev(trackers[k]);
        return av->device(modifies);
    }
    void bind() {  // common constructor tail: resolve the AtomVec, attach the box
        av = dynamic_cast<AtomVec *>(atoms.get());
        if (!av) throw std::runtime_error("parm_b200: Collection needs an AtomVec");
        if (!constraints.empty()) throw std::runtime_error("parm_b200: constraints are outside the hot-path scope (DESIGN.md)");
        OriginBox *ob = dynamic_cast<OriginBox *>(box.get());
        if (!ob) throw std::runtime_error("parm_b200: Collection needs an OriginBox");
        ob->attach(av->context());
    }
    void register_all(bool should_initialize) {  // collection.cpp:3-11
        for (size_t k = 0; k < trackers.size(); k++) {
            if (stat(trackers[k])) parm_b200::check(parm_integ_register_stat_tracker(integ, stat(trackers[k])));
            else parm_b200::check(parm_integ_register_tracker(integ, dev(trackers[k])));
        }
        for (size_t k = 0; k < interactions.size(); k++) parm_b200::check(parm_integ_register_interaction(integ, dev(interactions[k])));
        if (should_initialize) initialize();
    }
    void update_trackers() {
        ready();
        parm_b200::check(parm_integ_update_trackers(integ));
    }
    virtual flt set_forces_get_pressure(bool = true) {
        throw std::runtime_error("parm_b200: Collection::set_forces_get_pressure is used only by the NPT/NLCG collections (out of scope)");
    }

   public:
    Collection(sptr<Box> box, sptr<AtomGroup> atoms, vector<sptr<Interaction> > is = vector<sptr<Interaction> >(),
               vector<sptr<StateTracker> > ts = vector<sptr<StateTracker> >(),
               vector<sptr<Constraint> > cs = vector<sptr<Constraint> >())
        : box(box), atoms(atoms), interactions(is), trackers(ts), constraints(cs), av(NULL), integ(NULL) {
        bind();
    }
    virtual ~Collection() { parm_integ_destroy(integ); }

    virtual void initialize() {  // collection.cpp:13-19
        ready();
        parm_b200::check(parm_integ_initialize(integ));
    }
    virtual void set_forces(bool constraints_and_a = true) {
        ready();
        parm_b200::check(parm_integ_set_forces(integ, constraints_and_a ? 1 : 0));
    }
    virtual void timestep() = 0;
    //! nsteps x timestep() in one call (one host round trip per rebuild decision only)
    void timesteps(int nsteps) {
        ready();
        parm_b200::check(parm_integ_timestep(integ, nsteps));
    }

    flt degrees_of_freedom() { ready(false); return av->mobile_dof(); }  // collection.cpp:116-133 (no constraints)
    flt potential_energy() {
        ready(false);
        flt e = 0;
        parm_b200::check(parm_integ_potential_energy(integ, &e));
        return e;
    }
    flt energy() { return potential_energy() + kinetic_energy(); }
    virtual flt temp(bool minuscomv = true) {  // collection.cpp:135-142
        Vec v = Vec::Zero();
        if (minuscomv) v = com_velocity();
        int ndof = (int)degrees_of_freedom();
        if (minuscomv) ndof -= NDIM;
        return atoms->kinetic_energy(v) * 2 / ndof;
    }
    virtual flt kinetic_energy() { return atoms->kinetic_energy(); }
    virtual flt virial() {
        ready(false);
        flt w = 0;
        parm_b200::check(parm_integ_virial(integ, &w));
        return w;
    }
    virtual flt pressure() {  // collection.cpp:85-96
        flt V = box->V();
        flt E = 2.0 * kinetic_energy();
        E += virial();
        return E / V / flt(NDIM);
    }
    sptr<Box> get_box() { return box; }
    inline Vec com() { return atoms->com(); }
    inline Vec com_velocity() { return atoms->com_velocity(); }
    void reset_com_velocity() { atoms->reset_com_velocity(); }
    void scale_velocities(flt scaleby) { parm_b200::check(parm_scale_velocities(av->device(), scaleby)); }
    void scale_velocities_to_temp(flt T, bool minuscomv = true) {  // collection.cpp:31-35
        flt t = temp(minuscomv);
        scale_velocities(sqrt(T / t));
    }
    void scale_velocities_to_energy(flt E) {  // collection.cpp:37-43
        flt E0 = energy();
        flt k0 = kinetic_energy();
        flt goalkinetic = k0 + (E - E0);
        scale_velocities(sqrt(goalkinetic / k0));
    }
    virtual void add_interaction(sptr<Interaction> inter) {  // collection.hpp:113-116
        parm_inter *h = dev(inter);
        interactions.push_back(inter);
        ready();
        parm_b200::check(parm_integ_add_interaction(integ, h));
    }
    virtual void add_tracker(sptr<StateTracker> track) {  // collection.hpp:117-120
        if (parm_tracker *t = stat(track)) {
            trackers.push_back(track);
            ready();
            parm_b200::check(parm_integ_add_stat_tracker(integ, t));
            return;
        }
        parm_nlist *h = dev(track);
        trackers.push_back(track);
        ready();
        parm_b200::check(parm_integ_add_tracker(integ, h));
    }
    virtual void add_constraint(sptr<Constraint>) {
        throw std::runtime_error("parm_b200: constraints are outside the hot-path scope (DESIGN.md)");
    }
    void add(sptr<Interaction> a) { add_interaction(a); }
    void add(sptr<StateTracker> a) { add_tracker(a); }
    vector<sptr<Interaction> > get_interactions() { return interactions; }
};

class CollectionVerlet : public Collection {
   protected:
    flt dt;

   public:
    CollectionVerlet(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt,
                     vector<sptr<Interaction> > interactions = vector<sptr<Interaction> >(),
                     vector<sptr<StateTracker> > trackers = vector<sptr<StateTracker> >(),
                     vector<sptr<Constraint> > constraints = vector<sptr<Constraint> >())
        : Collection(box, atoms, interactions, trackers, constraints), dt(dt) {
        parm_b200::check(parm_verlet_create(av->context(), dt, &integ));
        register_all(true);
    }
    void timestep() {
        ready();
        parm_b200::check(parm_integ_timestep(integ, 1));
    }
    void set_dt(flt newdt) {
        dt = newdt;
        parm_b200::check(parm_integ_set_dt(integ, dt));
    }
};

class CollectionSol : public Collection {
   protected:
    flt dt, damping, desT;

   public:
    CollectionSol(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, const flt damping, const flt desired_temperature,
                  vector<sptr<Interaction> > interactions = vector<sptr<Interaction> >(),
                  vector<sptr<StateTracker> > trackers = vector<sptr<StateTracker> >(),
                  vector<sptr<Constraint> > constraints = vector<sptr<Constraint> >())
        : Collection(box, atoms, interactions, trackers, constraints), dt(dt), damping(damping), desT(desired_temperature) {
        // dt <= 0 -> std::invalid_argument, like collection.cpp:222-224
        parm_b200::check(parm_sol_create(av->context(), dt, damping, desired_temperature, (uint64_t)parm_b200::randengine()(), &integ));
        register_all(true);
    }
    void change_temperature(const flt damp, const flt desired_temperature) {
        damping = damp;
        desT = desired_temperature;
        parm_b200::check(parm_integ_set_temperature(integ, damp, desired_temperature));
    }
    void set_dt(const flt newdt) {
        dt = newdt;
        parm_b200::check(parm_integ_set_dt(integ, dt));
    }
    void timestep() {
        ready();
        parm_b200::check(parm_integ_timestep(integ, 1));
    }
};

namespace parm_b200 {
// Shared body of the integrators created through parm_integ_create (include/parm_b200.h PARM_INTEG_*).
class DeviceCollection : public Collection {
   protected:
    flt dt;
    void make(int type, const flt *params, int nparams, uint64_t seed = 0) {
        check(parm_integ_create(av->context(), type, params, nparams, seed, &integ));
        register_all(true);
    }
    void scalars(flt *out2) {
        ready(false);
        check(parm_integ_get_scalars(integ, out2));
    }

   public:
    DeviceCollection(sptr<Box> box, sptr<AtomGroup> atoms, flt dt, vector<sptr<Interaction> > is, vector<sptr<StateTracker> > ts,
                     vector<sptr<Constraint> > cs)
        : Collection(box, atoms, is, ts, cs), dt(dt) {}
    void timestep() {
        ready();
        check(parm_integ_timestep(integ, 1));
    }
    void set_dt(flt newdt) {
        dt = newdt;
        check(parm_integ_set_dt(integ, dt));
    }
};
}  // namespace parm_b200

#define PARM_B200_VECS                                                                    \
    vector<sptr<Interaction> > interactions = vector<sptr<Interaction> >(),              \
                               vector<sptr<StateTracker> > trackers = vector<sptr<StateTracker> >(), \
                               vector<sptr<Constraint> > constraints = vector<sptr<Constraint> >()

class CollectionDamped : public parm_b200::DeviceCollection {  // collection.hpp:273-295, collection.cpp:324-381
   public:
    CollectionDamped(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, const flt damping, PARM_B200_VECS)
        : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {
        flt p[2] = {dt, damping};
        make(PARM_INTEG_DAMPED, p, 2);
    }
    void change_damping(const flt damp) { parm_b200::check(parm_integ_set_param(integ, 3, damp)); }
};

class CollectionSolHT : public parm_b200::DeviceCollection {  // collection.hpp:328-353, collection.cpp:383-440
   public:
    CollectionSolHT(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, const flt damping, const flt desired_temperature,
                    PARM_B200_VECS)
        : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {
        flt p[3] = {dt, damping, desired_temperature};
        make(PARM_INTEG_SOLHT, p, 3, (uint64_t)parm_b200::randengine()());
    }
    void change_temperature(const flt newdt, const flt damp, const flt desired_temperature) {
        set_dt(newdt);
        parm_b200::check(parm_integ_set_param(integ, 3, damp));
        parm_b200::check(parm_integ_set_param(integ, 2, desired_temperature));
    }
};

class CollectionOverdamped : public parm_b200::DeviceCollection {  // collection.hpp:376-393, collection.cpp:471-492
   public:
    CollectionOverdamped(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, const flt gamma = 1.0, PARM_B200_VECS)
        : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {
        flt p[2] = {dt, gamma};
        make(PARM_INTEG_OVERDAMPED, p, 2);
    }
};

class CollectionNoseHoover : public parm_b200::DeviceCollection {  // collection.hpp:567-598, collection.cpp:1170-1249
   protected:
    flt Q, T;

   public:
    CollectionNoseHoover(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, const flt Q, const flt T, PARM_B200_VECS)
        : DeviceCollection(box, atoms, dt, interactions, trackers, constraints), Q(Q), T(T) {
        flt p[3] = {dt, Q, T};
        make(PARM_INTEG_NOSEHOOVER, p, 3);
    }
    void set_Q(flt newQ) {
        Q = newQ;
        parm_b200::check(parm_integ_set_param(integ, 1, Q));
    }
    void reset_bath() { parm_b200::check(parm_integ_reset_bath(integ)); }
    flt get_xi() { flt s[2]; scalars(s); return s[0]; }
    flt get_lns() { flt s[2]; scalars(s); return s[1]; }
    flt hamiltonian() {
        flt s[2];
        scalars(s);
        flt H = (kinetic_energy() + potential_energy() + (s[0] * s[0] * Q / 2) + (degrees_of_freedom() * s[1] * T));
        assert(not isnan(H));
        return H;
    }
};

class CollectionGaussianT : public parm_b200::DeviceCollection {  // collection.hpp:600-621, collection.cpp:1251-1299
   public:
    CollectionGaussianT(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, PARM_B200_VECS)
        : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {
        make(PARM_INTEG_GAUSSIANT, &dt, 1);
    }
};

class CollectionGear3A : public parm_b200::DeviceCollection {  // collection.hpp:623-637, collection.cpp:1301-1322
   public:
    CollectionGear3A(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, PARM_B200_VECS)
        : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {
        make(PARM_INTEG_GEAR3A, &dt, 1);
    }
};

class CollectionNLCG : public Collection {  // collection.hpp:400-474, collection.cpp:494-854
   protected:
    sptr<OriginBox> obox;
    flt state(int k) {
        flt o[16];
        parm_b200::check(parm_nlcg_get(integ, o));
        return o[k];
    }
    void set(int which, flt v) {
        ready();
        parm_b200::check(parm_nlcg_set(integ, which, v));
    }
    flt reduce(int what) {
        ready(false);
        flt r = 0;
        parm_b200::check(parm_nlcg_reduce(integ, what, &r));
        return r;
    }

   public:
    CollectionNLCG(sptr<OriginBox> box, sptr<AtomGroup> atoms, const flt dt, const flt P0,
                   vector<sptr<Interaction> > interactions = vector<sptr<Interaction> >(),
                   vector<sptr<StateTracker> > trackers = vector<sptr<StateTracker> >(),
                   vector<sptr<Constraint> > constraints = vector<sptr<Constraint> >(), const flt kappa = 10.0,
                   const flt kmax = 1000, const uint secmax = 40, const flt seceps = 1e-20)
        : Collection(boost::static_pointer_cast<Box>(box), atoms, interactions, trackers, constraints), obox(box) {
        parm_b200::check(parm_nlcg_create(av->context(), dt, P0, kappa, kmax, secmax, seceps, &integ));
        register_all(true);
    }
    flt kinetic_energy() { return reduce(PARM_NLCG_KINETIC); }  // Note: masses are ignored
    flt pressure() { return reduce(PARM_NLCG_PRESSURE); }
    flt hamiltonian() { return reduce(PARM_NLCG_HAMILTONIAN); }
    flt fdota() { return reduce(PARM_NLCG_FDOTA); }
    flt fdotf() { return reduce(PARM_NLCG_FDOTF); }
    flt fdotv() { return reduce(PARM_NLCG_FDOTV); }
    flt vdotv() { return reduce(PARM_NLCG_VDOTV); }
    void set_forces(bool constraints_and_a = true) { set_forces(constraints_and_a, true); }
    void set_forces(bool constraints_and_a, bool setV) {
        ready();
        parm_b200::check(parm_nlcg_set_forces(integ, constraints_and_a ? 1 : 0, setV ? 1 : 0));
    }
    void timestep() {
        ready();
        parm_b200::check(parm_integ_timestep(integ, 1));
        obox->pull(av->context());
    }
    void descend() {
        ready();
        parm_b200::check(parm_nlcg_descend(integ));
        obox->pull(av->context());
    }
    void reset() {
        ready();
        parm_b200::check(parm_nlcg_reset(integ));
    }
    void set_dt(flt newdt) { set(PARM_NLCG_DT, newdt); }
    void set_pressure_goal(flt P) { set(PARM_NLCG_P0, P); }
    flt get_pressure_goal() { return state(1); }
    void set_kappa(flt k) { set(PARM_NLCG_KAPPA, k); }
    void set_max_alpha(flt a) { set(PARM_NLCG_ALPHAMAX, a); }
    void set_max_alpha_fraction(flt a) { set(PARM_NLCG_AFRAC, a); }
    void set_max_dx(flt d) { set(PARM_NLCG_DXMAX, d); }
    void set_max_step(flt m) { set(PARM_NLCG_STEPMAX, m); }
    // the reference's public tracking members, as accessors
    flt get_alpha() { return state(8); }
    flt get_beta() { return state(9); }
    flt get_betaused() { return state(10); }
    flt get_dxsum() { return state(11); }
    flt get_alphavmax() { return state(12); }
    uint get_sec() { return (uint)state(13); }
};

#define PARM_B200_GEAR(NAME, TYPE)                                                                               \
    class NAME : public parm_b200::DeviceCollection {                                                            \
       public:                                                                                                   \
        NAME(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, uint ncorrectionsteps, PARM_B200_VECS)          \
            : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {                            \
            flt p[2] = {dt, (flt)ncorrectionsteps};                                                              \
            make(TYPE, p, 2);                                                                                    \
        }                                                                                                        \
        NAME(sptr<Box> box, sptr<AtomGroup> atoms, const flt dt, PARM_B200_VECS)                                 \
            : DeviceCollection(box, atoms, dt, interactions, trackers, constraints) {                            \
            flt p[2] = {dt, 1.0};                                                                                \
            make(TYPE, p, 2);                                                                                    \
        }                                                                                                        \
    }
PARM_B200_GEAR(CollectionGear4A, PARM_INTEG_GEAR4A);  // collection.hpp:639-671, collection.cpp:1324-1364
PARM_B200_GEAR(CollectionGear5A, PARM_INTEG_GEAR5A);  // collection.hpp:673-709, collection.cpp:1366-1407
PARM_B200_GEAR(CollectionGear6A, PARM_INTEG_GEAR6A);  // collection.hpp:711-755, collection.cpp:1409-1496
#undef PARM_B200_GEAR
#undef PARM_B200_VECS

#endif
