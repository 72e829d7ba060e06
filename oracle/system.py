"""Oracle data container -- the attribute surface ``pycgtool/frame.py`` exposes.

TEST INFRASTRUCTURE ONLY.  A ``System`` is the oracle's stand-in for the
reference ``Frame`` (``pycgtool/frame.py:66-272``): residues in topology order,
each with an ordered list of atom names and the global index of its first atom
(mdtraj topologies keep a residue's atoms contiguous), plus the trajectory
arrays the hot path reads (``frame.py:138-159``).
"""

import dataclasses
import typing

import numpy as np

from . import fileio


@dataclasses.dataclass
class Res:
    name: str
    atoms: typing.List[str]
    first: int  # global index of atoms[0]

    def index_of(self, atom_name):
        """mdtraj ``Residue.atom(name)``: first atom with that name, else KeyError."""
        for k, nm in enumerate(self.atoms):
            if nm == atom_name:
                return self.first + k
        raise KeyError(atom_name)


@dataclasses.dataclass
class System:
    residues: typing.List[Res]
    xyz: np.ndarray                         # (F, N, 3) float32, nm
    box_vectors: typing.Optional[np.ndarray] = None   # (F, 3, 3) float32 rows a,b,c, or None
    time: typing.Optional[np.ndarray] = None

    @property
    def natoms(self):
        return self.xyz.shape[1]

    @property
    def n_frames(self):
        return self.xyz.shape[0]

    @property
    def unitcell_lengths(self):
        """|a|,|b|,|c| per frame, float32, or None (``frame.py:147-153``)."""
        if self.box_vectors is None:
            return None
        v = self.box_vectors.astype(np.float64)
        return np.sqrt((v * v).sum(axis=2)).astype(np.float32)

    @property
    def unitcell_angles(self):
        """alpha(b,c), beta(a,c), gamma(a,b) in degrees, float32, or None."""
        if self.box_vectors is None:
            return None
        v = self.box_vectors.astype(np.float64)
        a, b, c = v[:, 0], v[:, 1], v[:, 2]

        def ang(p, q):
            cosv = (p * q).sum(axis=1) / np.sqrt((p * p).sum(axis=1) * (q * q).sum(axis=1))
            return np.degrees(np.arccos(cosv))

        return np.stack([ang(b, c), ang(a, c), ang(a, b)], axis=1).astype(np.float32)


def from_residue_lists(residues, xyz, box_vectors=None, time=None):
    """Build a System from ``[(resname, [atomnames])]`` as the readers return it."""
    out, first = [], 0
    for name, atoms in residues:
        out.append(Res(name, list(atoms), first))
        first += len(atoms)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    if first != xyz.shape[1]:
        raise ValueError("Number of atoms does not match between topology and trajectory files")
    if time is None:
        time = np.arange(xyz.shape[0], dtype=np.float32)
    return System(out, xyz, box_vectors, time)


def load(topology_file, trajectory_file=None):
    """``Frame(topology_file, trajectory_file)`` (``frame.py:69-110``) for .gro/.xtc."""
    residues, xyz, box = fileio.read_gro(topology_file)
    time = None
    if trajectory_file is not None:
        xyz, time, box = fileio.read_xtc(trajectory_file)
        if not box.any():
            box = None
    return from_residue_lists(residues, xyz, box, time)
