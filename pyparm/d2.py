"""pyparm.d2: the two-dimensional build of the reference's SWIG module (make pyparm/_sim2d.so, VEC2D)."""
from ._bind import populate as _populate

_populate(globals(), 2)
