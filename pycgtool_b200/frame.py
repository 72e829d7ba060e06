"""``Frame``: the data contract of the hot path.

Mirrors the attribute surface of the reference ``Frame``
(``pycgtool/frame.py:66-272``) -- ``residues``, ``atoms``, ``natoms``,
``n_frames``, ``unitcell_lengths``, ``unitcell_angles``, ``time``,
``_trajectory.xyz`` and per-atom ``coords`` views -- on top of the
array-backed containers in :mod:`pycgtool_b200.topology`.  A frame can also be
created straight from arrays (``Frame.from_arrays``), host or device resident,
which is how synthetic and pre-loaded trajectories enter the accelerated path.
"""

import logging
import pathlib

import numpy as np

from . import formats
from .topology import Topology, Trajectory

logger = logging.getLogger(__name__)


class UnsupportedFormatException(Exception):
    """Raised when a topology / trajectory format cannot be parsed."""

    def __init__(self, msg=None):
        super().__init__(msg or "Topology/Trajectory format not supported by this reader")


class NonMatchingSystemError(ValueError):
    """Raised when topology and trajectory do not describe the same system."""

    def __init__(self, msg=None):
        super().__init__(msg or "Number of atoms does not match between topology and trajectory files")


class Frame:
    """Topology plus trajectory, loaded fully into memory (``frame.py:69-117``)."""

    def __init__(self, topology_file=None, trajectory_file=None, frame_start=0, frame_end=None,
                 device_decode=False):
        """``device_decode=True`` expands an .xtc trajectory on the GPU (only the compressed file
        crosses PCIe) and leaves the coordinates device-resident for ``Mapping.apply``."""
        if topology_file is None:
            self._topology = Topology()   # an empty frame to be filled in (CG output)
            self._topology.owner = self
            return
        try:
            traj = formats.load(topology_file)
            self._topology = traj.topology
            if trajectory_file is not None:
                try:
                    traj = formats.load(trajectory_file, top=self._topology,
                                        on_device=device_decode and frame_start == 0)
                except ValueError as exc:
                    raise NonMatchingSystemError from exc
                # frame.py:119-136: frame_end only applies together with a non-zero frame_start
                if frame_start != 0:
                    traj = traj[frame_start:frame_end] if frame_end is not None else traj[frame_start:]
        except OSError as exc:
            if "format is not supported" in str(exc):
                raise UnsupportedFormatException from exc
            raise
        self._attach(traj)

    @classmethod
    def from_arrays(cls, topology, xyz=None, unitcell_vectors=None, time=None, xyz_device=None):
        """Frame over existing coordinates (numpy ``xyz`` and/or a CUDA tensor ``xyz_device``)."""
        frame = cls()
        frame._topology = topology
        frame._attach(Trajectory(xyz, topology, time, unitcell_vectors, xyz_device=xyz_device))
        return frame

    def _attach(self, traj):
        traj.topology = self._topology
        self._trajectory = traj
        self._topology.owner = self
        self.unitcell_lengths = traj.unitcell_lengths
        self.unitcell_angles = traj.unitcell_angles
        self.time = traj.time

    # ------------------------------------------------------------------ topology views
    @property
    def residues(self):
        return self._topology.residues

    def residue(self, *args, **kwargs):
        return self._topology.residue(*args, **kwargs)

    @property
    def atoms(self):
        return self._topology.atoms

    def atom(self, *args, **kwargs):
        return self._topology.atom(*args, **kwargs)

    @property
    def natoms(self):
        return self._topology.n_atoms

    @property
    def n_frames(self):
        return self._trajectory.n_frames

    # ------------------------------------------------------------------ incremental construction
    def add_residue(self, name, chain=None, **kwargs):
        if hasattr(self, "_trajectory"):
            raise TypeError("Cannot edit residues if a trajectory has been loaded")
        return self._topology.add_residue(name, chain, **kwargs)

    def add_atom(self, name, element, residue):
        if hasattr(self, "_trajectory"):
            raise TypeError("Cannot edit atoms if a trajectory has been loaded")
        return self._topology.add_atom(name, element, residue)

    def build_trajectory(self):
        """Assemble a trajectory from per-atom ``coords`` (``frame.py:246-272``)."""
        top = self._topology
        columns = [top._pending_coords[i] for i in range(top.n_atoms)]
        n_frames = len(self.time)
        if columns:
            xyz = np.array(columns, dtype=np.float32).swapaxes(0, 1)
        else:
            xyz = np.empty((n_frames, 0, 3), dtype=np.float32)
        vectors = getattr(self, "_unitcell_vectors", None)
        new = Trajectory(xyz, top, self.time, vectors)
        top._pending_coords = {}
        if hasattr(self, "_trajectory"):
            self._trajectory += new
        else:
            self._trajectory = new

    # ------------------------------------------------------------------ output
    def save(self, filename, frame_number=None, **kwargs):
        """Write the trajectory (.gro: one frame; .xtc: all frames or the selected one)."""
        from . import util
        filename = pathlib.Path(filename)
        util.backup_file(filename)
        if filename.suffix.lower() == ".gro":
            formats.write_gro(filename, self._trajectory, 0 if frame_number is None else frame_number)
        elif filename.suffix.lower() == ".xtc":
            traj = self._trajectory if frame_number is None else self._trajectory.slice(frame_number)
            formats.write_xtc(filename, traj, **kwargs)
        else:
            raise UnsupportedFormatException(f"cannot write '{filename.suffix}' files")
