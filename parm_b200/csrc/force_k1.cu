// Pair-force kernel instantiations (force_kernel.cuh), split by functor for parallel compilation.
#include "force_kernel.cuh"

PARM_INSTANTIATE_FORCE_KIND(PARM_PAIR_LJCUT)
PARM_INSTANTIATE_FORCE_KIND(PARM_PAIR_LJATTRACTCUT)
PARM_INSTANTIATE_FORCE_KIND(PARM_PAIR_LJATTRACTFIXEDREPULSE)
