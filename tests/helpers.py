"""Shared test helpers: bridges between the product containers and the oracle."""

import numpy as np

from oracle import system as osystem


def oracle_system(topology, xyz, box_vectors=None, time=None):
    """Oracle ``System`` with the same residues / atoms as a product ``Topology``."""
    top = topology
    residues = []
    for r in range(top.n_residues):
        a, b = int(top.res_first[r]), int(top.res_first[r + 1])
        residues.append((top.res_names[top.res_name_id[r]], [top.atom_names[i] for i in top.atom_name_id[a:b]]))
    return osystem.from_residue_lists(residues, np.asarray(xyz, dtype=np.float32), box_vectors, time)


def ulp_diff(a, b):
    """Largest difference in units in the last place between two float32 arrays."""
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return int(np.abs(a - b).max()) if a.size else 0
