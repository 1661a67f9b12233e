"""pyparm.d3: the three-dimensional build of the reference's SWIG module (make pyparm/_sim3d.so, VEC3D)."""
from ._bind import populate as _populate

_populate(globals(), 3)
