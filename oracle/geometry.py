"""Oracle for bonded-term geometry.  TEST INFRASTRUCTURE ONLY.

The arithmetic of ``BondSet.apply`` lives in the reference's un-vendored
dependency **mdtraj 1.9.7** (``poetry.lock:391-392``), reached at
``pycgtool/bondset.py:501-505,530`` as ``mdtraj.compute_distances / angles /
dihedrals(frame._trajectory, indices)`` with the default ``periodic=True``.
This module restates mdtraj's published algorithm (``mdtraj/geometry/
distance.py``, ``angle.py``, ``dihedral.py`` and their kernels) in float32
NumPy:

* no unit cell -> plain differences;
* unit cell with all angles ~ 90 deg (``np.allclose``) -> orthorhombic minimum
  image ``r -= L * round(r * (1 / L))`` with the box diagonal (the C kernel's form);
* otherwise triclinic: reduce the box vectors, subtract rounded multiples of
  c, b, a, then take the shortest of the 27 neighbouring images.

Anchors: the reference's own outputs of these calls, ``tests/data/ALLA_*.dat``
(1001 frames x 18 terms) and ``tests/test_bondset.py:64-85,158-182``.
"""

import numpy as np

F32 = np.float32


def _reduced_box(box):
    """mdtraj ``_reduce_box_vectors``: (F,3,3) rows a,b,c -> reduced a,b,c."""
    a = box[:, 0].astype(F32).copy()
    b = box[:, 1].astype(F32).copy()
    c = box[:, 2].astype(F32).copy()
    c -= b * np.round(c[:, 1] / b[:, 1])[:, None]
    c -= a * np.round(c[:, 0] / a[:, 0])[:, None]
    b -= a * np.round(b[:, 0] / a[:, 0])[:, None]
    return a, b, c


def is_orthorhombic(box):
    """``np.allclose(traj.unitcell_angles, 90)`` on angles derived from the vectors."""
    v = box.astype(np.float64)
    a, b, c = v[:, 0], v[:, 1], v[:, 2]

    def ang(p, q):
        cosv = (p * q).sum(axis=1) / np.sqrt((p * p).sum(axis=1) * (q * q).sum(axis=1))
        return np.degrees(np.arccos(cosv))

    angles = np.stack([ang(b, c), ang(a, c), ang(a, b)], axis=1).astype(F32)
    return bool(np.allclose(angles, 90))


def displacement(xyz, i, j, box=None):
    """Minimum-image vector ``x[j] - x[i]`` for index arrays i, j -> (F, n, 3) float32."""
    r = (xyz[:, j] - xyz[:, i]).astype(F32)
    if box is None:
        return r
    if is_orthorhombic(box):
        # mdtraj's C kernel multiplies by the float32 inverse box length instead of dividing
        # (distancekernel.h: r12 -= round(r12 * inv_box_size) * box_size).
        length = np.stack([box[:, 0, 0], box[:, 1, 1], box[:, 2, 2]], axis=1).astype(F32)[:, None, :]
        inverse = (F32(1.0) / length).astype(F32)
        return (r - length * np.round((r * inverse).astype(F32))).astype(F32)
    a, b, c = _reduced_box(box)
    a, b, c = a[:, None, :], b[:, None, :], c[:, None, :]
    r = r - c * np.round(r[..., 2:3] / c[..., 2:3])
    r = r - b * np.round(r[..., 1:2] / b[..., 1:2])
    r = r - a * np.round(r[..., 0:1] / a[..., 0:1])
    r = r.astype(F32)
    best = r.copy()
    best_d2 = (r * r).sum(axis=-1)
    for ii in (-1, 0, 1):
        for jj in (-1, 0, 1):
            for kk in (-1, 0, 1):
                cand = (r + a * F32(ii) + b * F32(jj) + c * F32(kk)).astype(F32)
                d2 = (cand * cand).sum(axis=-1)
                take = d2 < best_d2
                best[take] = cand[take]
                best_d2 = np.where(take, d2, best_d2)
    return best


def distances(xyz, pairs, box=None):
    """``mdtraj.compute_distances`` -> (F, n) float32 (nm)."""
    pairs = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    r = displacement(xyz, pairs[:, 0], pairs[:, 1], box)
    return np.sqrt((r * r).sum(axis=-1)).astype(F32)


def angles(xyz, triplets, box=None):
    """``mdtraj.compute_angles`` -> (F, n) float32 radians in [0, pi]."""
    t = np.asarray(triplets, dtype=np.int32).reshape(-1, 3)
    u = displacement(xyz, t[:, 1], t[:, 0], box)
    v = displacement(xyz, t[:, 1], t[:, 2], box)
    dot = (u * v).sum(axis=-1)
    norm = np.sqrt((u * u).sum(axis=-1) * (v * v).sum(axis=-1))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.arccos(np.clip(dot / norm, -1, 1)).astype(F32)


def dihedrals(xyz, quartets, box=None):
    """``mdtraj.compute_dihedrals`` -> (F, n) float32 radians in (-pi, pi]."""
    q = np.asarray(quartets, dtype=np.int32).reshape(-1, 4)
    b1 = displacement(xyz, q[:, 0], q[:, 1], box)
    b2 = displacement(xyz, q[:, 1], q[:, 2], box)
    b3 = displacement(xyz, q[:, 2], q[:, 3], box)
    c1 = np.cross(b2, b3)
    c2 = np.cross(b1, b2)
    p1 = (b1 * c1).sum(axis=-1) * np.sqrt((b2 * b2).sum(axis=-1))
    p2 = (c1 * c2).sum(axis=-1)
    return np.arctan2(p1, p2).astype(F32)


MEASURE = {2: distances, 3: angles, 4: dihedrals}
