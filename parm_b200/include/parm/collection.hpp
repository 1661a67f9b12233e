// parm_b200 drop-in for ParM's src/collection.hpp -- Collection, CollectionVerlet, CollectionSol
// (collection.hpp:23-130, 205-263, 360-374; collection.cpp:3-208, 210-322, 442-469).
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
    static parm_nlist *dev(sptr<StateTracker> &t) {
        NeighborList *n = dynamic_cast<NeighborList *>(t.get());
        if (!n) throw std::runtime_error("parm_b200: only NeighborList trackers run on the device (statistics trackers are out of scope)");
        return n->handle();
    }
    parm_ctx *ready(bool modifies = true) {
        for (size_t k = 0; k < interactions.size(); k++) dev(interactions[k]);
        for (size_t k = 0; k < trackers.size(); k++) dev(trackers[k]);
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
        for (size_t k = 0; k < trackers.size(); k++) parm_b200::check(parm_integ_register_tracker(integ, dev(trackers[k])));
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

#endif
