"""Oracle for the statistics and Boltzmann inversion.  TEST INFRASTRUCTURE ONLY.

Restates ``pycgtool/util.py:24-53`` (circular mean / variance),
``pycgtool/functionalforms.py:92-148`` (force-constant formulas) and
``pycgtool/bondset.py:62-91`` (``Bond.boltzmann_invert``).  Arithmetic is kept
in the reference's dtypes: float32 values with NumPy's float32 pairwise
reductions, then float64 scalar maths for the force constant.
"""

import math

import numpy as np


def circular_mean(values):
    """``util.circular_mean`` (util.py:24-37): atan2(<sin>, <cos>), NaN-skipping."""
    values = np.asarray(values)
    vec = np.array([np.cos(values), np.sin(values)]).T
    x_av, y_av = np.nanmean(vec, axis=0)
    return np.arctan2(y_av, x_av)


def circular_variance(values):
    """``util.circular_variance`` (util.py:40-53): <wrap(v - mean)^2>."""
    values = np.asarray(values)
    two_pi = 2 * np.pi
    diff = values - circular_mean(values)
    diff -= np.rint(diff / two_pi) * two_pi
    return np.nanmean(np.square(diff))


def mean_and_variance(values, natoms):
    """Statistic selection of ``bondset.py:172-179``: lengths arithmetic, angles
    *and* dihedrals circular."""
    if natoms > 2:
        return circular_mean(values), circular_variance(values)
    return np.nanmean(values), np.nanvar(values)


RT_PER_K = 8.314 / 1000.0

# form name -> gromacs type ids for (length, angle, dihedral); functionalforms.py:98,115,127,136,145
TYPE_IDS = {
    "Harmonic": (1, 1, 1),
    "CosHarmonic": (None, 2, None),
    "MartiniDefaultLength": (1, None, None),
    "MartiniDefaultAngle": (None, 2, None),
    "MartiniDefaultDihedral": (None, None, 1),
}


def default_forms(default_fc=False, length_form=None, angle_form=None, dihedral_form=None):
    """``BondSet._get_default_func_forms`` (bondset.py:108-147) -> {natoms: form name}."""
    if default_fc:
        forms = ["MartiniDefaultLength", "MartiniDefaultAngle", "MartiniDefaultDihedral"]
    else:
        forms = ["Harmonic", "CosHarmonic", "Harmonic"]
    for k, override in enumerate((length_form, angle_form, dihedral_form)):
        if override is not None:
            forms[k] = override
    return {2: forms[0], 3: forms[1], 4: forms[2]}


def force_constant(form, mean, var, temp):
    """functionalforms.py:101-104,117-121,129-148.  var == 0 -> inf (bondset.py:89-91)."""
    rt = 8.314 * temp / 1000.0
    if form == "Harmonic":
        denom = float(var)
    elif form == "CosHarmonic":
        denom = math.sin(mean) ** 2 * float(var)
    elif form == "MartiniDefaultLength":
        return 1250.0
    elif form == "MartiniDefaultAngle":
        return 25.0
    elif form == "MartiniDefaultDihedral":
        return 50.0
    else:
        raise KeyError(form)
    if denom == 0:
        return math.inf
    return rt / denom


def boltzmann_invert(values, natoms, form, temp=310.0):
    """``Bond.boltzmann_invert`` (bondset.py:62-91) -> (eqm, fconst).

    Dihedral eqm is the phase shift, not the mean: ``eqm -= pi`` then wrapped to
    [-pi, pi) (bondset.py:81-84).  Empty ``values`` raises ValueError.
    """
    values = np.asarray(values)
    if len(values) == 0:
        raise ValueError("No bonds were measured")
    mean, var = mean_and_variance(values, natoms)
    eqm = float(mean)        # numpy-1.x promotion: float32 scalar (op) Python float -> float64
    if natoms == 4:
        eqm = eqm - np.pi
        eqm = ((eqm + np.pi) % (2 * np.pi)) - np.pi
    return float(eqm), force_constant(form, float(mean), var, temp)
