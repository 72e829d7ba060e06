"""pycgtool_b200: PyCGTOOL's per-frame hot path on B200 GPUs.

Drop-in for the three reference methods on the path --
``Mapping.apply``, ``BondSet.apply`` and ``BondSet.boltzmann_invert`` -- with
the reference's constructors and result attributes, backed by hand-written
CUDA kernels for sm_100a in ``libpycgtool_b200.so`` (C ABI in
``include/pycgtool_b200.h``).  There is no CPU fallback: without the built
library and a CUDA device the three methods raise.
"""

from .frame import Frame
from .mapping import BeadMap, Mapping, VirtualMap
from .bondset import Bond, BondSet
from .functionalforms import FunctionalForm, get_functional_forms
from .pipeline import map_and_measure

__version__ = "0.1.0"
__all__ = ["Frame", "Mapping", "BeadMap", "VirtualMap", "BondSet", "Bond", "FunctionalForm",
           "get_functional_forms", "map_and_measure"]
