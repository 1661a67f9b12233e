// parm_b200: only the Constraint interface of ParM's src/constraints.hpp (constraints.hpp:15-22);
// concrete constraints and the statistics trackers are outside the hot-path scope.
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
#endif
