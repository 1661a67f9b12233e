// Cell-tile pair kernel instantiations (force_tile.cuh), part 1
#include "force_tile.cuh"
PARM_INSTANTIATE_FORCE_TILE_KIND(PARM_PAIR_LJATTRACTREPULSE)
PARM_INSTANTIATE_FORCE_TILE_KIND(PARM_PAIR_LJCUT)
