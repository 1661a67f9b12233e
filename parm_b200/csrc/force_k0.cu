// Pair-force kernel instantiations (force_kernel.cuh), split by functor for parallel compilation.
#include "force_kernel.cuh"

PARM_INSTANTIATE_FORCE_KIND(PARM_PAIR_LJREPULSE)
PARM_INSTANTIATE_FORCE_KIND(PARM_PAIR_REPULSION)
PARM_INSTANTIATE_FORCE_KIND(PARM_PAIR_LJATTRACTREPULSE)
