"""Functional forms used to turn bond statistics into force constants.

API-compatible with ``pycgtool/functionalforms.py``: forms are discovered
through ``FunctionalForm.__subclasses__()`` (so users can add their own), take
injected ``mean_func`` / ``variance_func`` callables, and expose ``eqm``,
``fconst`` and the GROMACS type ids.  The built-in forms additionally carry
``device_form``: the id the Boltzmann epilogue kernel (``cgb_boltzmann``)
evaluates on the GPU, so for them no Python arithmetic runs on the hot path.
User-defined forms get ``mean_func`` / ``variance_func`` that return the
GPU-computed statistics (see ``bondset.Bond.boltzmann_invert``).
"""

import abc
import enum
import math

import numpy as np

from . import _lib

GAS_CONSTANT = 8.314  # J / (mol K); the reference's value (functionalforms.py:102)


def get_functional_forms():
    """Enum of every known functional form, keyed by class name."""
    return enum.Enum("FunctionalForms", {cls.__name__: cls for cls in FunctionalForm.__subclasses__()})


class FunctionalForm(metaclass=abc.ABCMeta):
    """Converts the spread of a measured coordinate into a force constant."""

    #: id evaluated by the GPU epilogue; ``None`` = evaluate ``fconst`` in Python
    device_form = None

    def __init__(self, mean_func=np.nanmean, variance_func=np.nanvar):
        self._mean_func = mean_func
        self._variance_func = variance_func

    def eqm(self, values, temp):  # pylint: disable=unused-argument
        """Equilibrium value: the injected mean of the measurements."""
        return self._mean_func(values)

    @abc.abstractmethod
    def fconst(self, values, temp):
        """Force constant from the measurements at temperature ``temp``."""
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def gromacs_type_ids(self):
        """GROMACS potential type ids when used as (length, angle, dihedral)."""
        raise NotImplementedError

    @classmethod
    def gromacs_type_id_by_natoms(cls, natoms):
        type_id = cls.gromacs_type_ids[natoms - 2]
        if type_id is None:
            raise TypeError(f"The functional form {cls.__name__} does not have a defined GROMACS "
                            f"potential type when used with {natoms} atoms.")
        return type_id


def _rt(temp):
    return GAS_CONSTANT * temp / 1000.0


class Harmonic(FunctionalForm):
    """k = RT / var."""

    gromacs_type_ids = (1, 1, 1)
    device_form = _lib.FORM_IDS["Harmonic"]

    def fconst(self, values, temp):
        return _rt(temp) / self._variance_func(values)


class CosHarmonic(FunctionalForm):
    """Cosine-based angle potential: k = RT / (sin^2(mean) var)."""

    gromacs_type_ids = (None, 2, None)
    device_form = _lib.FORM_IDS["CosHarmonic"]

    def fconst(self, values, temp):
        spread = self._variance_func(values)
        return _rt(temp) / (math.sin(self.eqm(values, temp)) ** 2 * spread)


class MartiniDefaultLength(FunctionalForm):
    """Fixed MARTINI bond force constant."""

    gromacs_type_ids = (1, None, None)
    device_form = _lib.FORM_IDS["MartiniDefaultLength"]

    def fconst(self, values, temp):
        return 1250.0


class MartiniDefaultAngle(FunctionalForm):
    """Fixed MARTINI angle force constant."""

    gromacs_type_ids = (None, 2, None)
    device_form = _lib.FORM_IDS["MartiniDefaultAngle"]

    def fconst(self, values, temp):
        return 25.0


class MartiniDefaultDihedral(FunctionalForm):
    """Fixed MARTINI dihedral force constant."""

    gromacs_type_ids = (None, None, 1)
    device_form = _lib.FORM_IDS["MartiniDefaultDihedral"]

    def fconst(self, values, temp):
        return 50.0
