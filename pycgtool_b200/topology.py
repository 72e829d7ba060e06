"""Array-backed topology and trajectory containers.

The reference leans on ``mdtraj.Topology`` / ``mdtraj.Trajectory``
(``pycgtool/frame.py:117,143-159,206-229,264-272``), one Python object per
atom.  That does not scale to the 20M-atom systems this package targets, so
the topology is a handful of integer arrays (residue name id, first atom of
each residue, atom name id) with thin ``Residue`` / ``Atom`` views created on
demand; the attribute surface the hot path and the reference tests use
(``.name``, ``.index``, ``.atoms``, ``.atom(name)``, ``.coords``) is kept.

A ``Trajectory`` holds coordinates either on the host (numpy) or on the GPU
(torch tensor) and converts lazily, so mapped frames stay resident in HBM for
the measurement stage and are only copied back when somebody asks.
"""

import numpy as np


class Atom:
    """View of one atom (or CG bead)."""

    __slots__ = ("_top", "index")

    def __init__(self, top, index):
        self._top = top
        self.index = int(index)

    @property
    def name(self):
        return self._top.atom_names[self._top.atom_name_id[self.index]]

    @property
    def residue(self):
        return self._top.residue(int(np.searchsorted(self._top.res_first, self.index, side="right") - 1))

    @property
    def coords(self):
        """(F, 3) view into the owner's coordinates (``frame.py:143-144``)."""
        if self.index in self._top._pending_coords:
            return self._top._pending_coords[self.index]
        owner = self._top.owner
        if owner is None or not hasattr(owner, "_trajectory"):
            raise AttributeError("atom has no coordinates")
        return owner._trajectory.xyz[:, self.index]

    @coords.setter
    def coords(self, value):
        self._top._pending_coords[self.index] = value

    def __repr__(self):
        return f"<Atom {self.index} {self.name}>"

    def __eq__(self, other):       # views are created on demand: identity is (topology, index)
        return isinstance(other, Atom) and other._top is self._top and other.index == self.index

    def __hash__(self):
        return hash((id(self._top), "atom", self.index))


class Residue:
    """View of one residue: a contiguous run of atoms."""

    __slots__ = ("_top", "index")

    def __init__(self, top, index):
        self._top = top
        self.index = int(index)

    @property
    def name(self):
        return self._top.res_names[self._top.res_name_id[self.index]]

    @property
    def n_atoms(self):
        return int(self._top.res_first[self.index + 1] - self._top.res_first[self.index])

    def __eq__(self, other):       # views are created on demand: identity is (topology, index)
        return isinstance(other, Residue) and other._top is self._top and other.index == self.index

    def __hash__(self):
        return hash((id(self._top), "residue", self.index))

    @property
    def atoms(self):
        first = int(self._top.res_first[self.index])
        return (Atom(self._top, i) for i in range(first, first + self.n_atoms))

    def atom(self, key):
        """Atom by position in the residue or by name (first match; KeyError if absent)."""
        first = int(self._top.res_first[self.index])
        if isinstance(key, str):
            top = self._top
            try:
                want = top.atom_names.index(key)
            except ValueError:
                raise KeyError(key)
            ids = top.atom_name_id[first:first + self.n_atoms]
            hits = np.nonzero(ids == want)[0]
            if len(hits) == 0:
                raise KeyError(key)
            return Atom(top, first + int(hits[0]))
        if key < 0 or key >= self.n_atoms:
            raise IndexError(key)
        return Atom(self._top, first + key)

    def __repr__(self):
        return f"<Residue {self.index} {self.name}>"


class Topology:
    """Residues as contiguous atom runs, names interned in small tables."""

    def __init__(self):
        self.res_names = []                             # table of distinct residue names
        self.atom_names = []                            # table of distinct atom names
        self.res_name_id = np.zeros(0, dtype=np.int32)  # (R,)
        self.res_first = np.zeros(1, dtype=np.int64)    # (R+1,) first atom of each residue
        self.atom_name_id = np.zeros(0, dtype=np.int32) # (N,)
        # numbering as read from a file (mdtraj's residue.resSeq / atom.serial); None for topologies built in
        # memory, which mdtraj numbers from zero when it writes them
        self.res_seq = None
        self.atom_serial = None
        self._owner = None        # weak reference to the Frame whose coordinates the atom views read
        self._pending_coords = {}
        self._grow_res, self._grow_first, self._grow_atoms = None, None, None

    # ------------------------------------------------------------------ bulk constructors
    @classmethod
    def from_arrays(cls, res_names, res_name_id, res_first, atom_names, atom_name_id):
        top = cls()
        top.res_names = list(res_names)
        top.atom_names = list(atom_names)
        top.res_name_id = np.ascontiguousarray(res_name_id, dtype=np.int32)
        top.res_first = np.ascontiguousarray(res_first, dtype=np.int64)
        top.atom_name_id = np.ascontiguousarray(atom_name_id, dtype=np.int32)
        if len(top.res_first) != len(top.res_name_id) + 1 or top.res_first[-1] != len(top.atom_name_id):
            raise ValueError("inconsistent topology arrays")
        return top

    @property
    def owner(self):
        return None if self._owner is None else self._owner()

    @owner.setter
    def owner(self, frame):
        import weakref
        self._owner = None if frame is None else weakref.ref(frame)

    @classmethod
    def from_residue_lists(cls, residues):
        """``[(resname, [atomname, ...]), ...]`` -> Topology."""
        rtab, atab = {}, {}
        rid, first, aid = [], [0], []
        for rname, atoms in residues:
            rid.append(rtab.setdefault(rname, len(rtab)))
            for a in atoms:
                aid.append(atab.setdefault(a, len(atab)))
            first.append(len(aid))
        return cls.from_arrays(list(rtab), rid, first, list(atab), aid)

    # ------------------------------------------------------------------ incremental API (frame.py:194-229)
    def _flush(self):
        if self._grow_res is None:
            return
        self.res_name_id = np.array(self._grow_res, dtype=np.int32)
        self.res_first = np.array(self._grow_first + [len(self._grow_atoms)], dtype=np.int64)
        self.atom_name_id = np.array(self._grow_atoms, dtype=np.int32)

    def _start_growth(self):
        if self._grow_res is None:
            self._grow_res = list(self.res_name_id)
            self._grow_first = list(self.res_first[:-1])
            self._grow_atoms = list(self.atom_name_id)

    def add_chain(self):
        return 0

    def chain(self, index):
        if index != 0:
            raise IndexError(index)
        return 0

    def add_residue(self, name, chain=None, **_):
        self._start_growth()
        if name not in self.res_names:
            self.res_names.append(name)
        self._grow_res.append(self.res_names.index(name))
        self._grow_first.append(len(self._grow_atoms))
        self._flush()
        return Residue(self, len(self._grow_res) - 1)

    def add_atom(self, name, element, residue):
        self._start_growth()
        if residue.index != len(self._grow_res) - 1:
            raise ValueError("atoms can only be appended to the last residue")
        if name not in self.atom_names:
            self.atom_names.append(name)
        self._grow_atoms.append(self.atom_names.index(name))
        self._flush()
        return Atom(self, len(self._grow_atoms) - 1)

    # ------------------------------------------------------------------ views
    @property
    def n_atoms(self):
        return len(self.atom_name_id)

    @property
    def n_residues(self):
        return len(self.res_name_id)

    @property
    def residues(self):
        return (Residue(self, i) for i in range(self.n_residues))

    def residue(self, index):
        if index < 0 or index >= self.n_residues:
            raise IndexError(index)
        return Residue(self, index)

    @property
    def atoms(self):
        return (Atom(self, i) for i in range(self.n_atoms))

    def atom(self, index):
        if index < 0 or index >= self.n_atoms:
            raise IndexError(index)
        return Atom(self, index)

    # ------------------------------------------------------------------ vectorised name lookups
    def residue_templates(self):
        """Group residues by (name, atom-name sequence).

        Returns ``(template_of_residue int32 (R,), templates)`` where ``templates`` is a
        list of ``(res_name_id, atom_name_id_row)``.  Name-based lookups are resolved
        once per template and broadcast to its residues, which is what turns the
        reference's O(residues x names) Python scans into array operations.
        """
        cached = getattr(self, "_templates", None)
        if cached is not None and cached[0] == (self.n_residues, self.n_atoms):
            return cached[1], cached[2]
        lens = np.diff(self.res_first)
        tid = np.zeros(self.n_residues, dtype=np.int32)
        templates = []
        if self.n_residues:
            keys = self.res_name_id.astype(np.int64) * (int(lens.max()) + 1) + lens
            for key in np.unique(keys):
                members = np.nonzero(keys == key)[0]
                natoms = int(lens[members[0]])
                firsts = self.res_first[members]
                block = self.atom_name_id[firsts[:, None] + np.arange(natoms)[None, :]] if natoms else \
                    np.zeros((len(members), 0), dtype=np.int32)
                rows, inverse = np.unique(block, axis=0, return_inverse=True)
                inverse = inverse.reshape(-1)
                tid[members] = len(templates) + inverse
                name_id = int(self.res_name_id[members[0]])
                templates.extend((name_id, row) for row in rows)
        self._templates = ((self.n_residues, self.n_atoms), tid, templates)
        return tid, templates

    def atom_index_by_name(self, atom_name, residues):
        """Global index of the first atom called ``atom_name`` in each of ``residues``
        (array of residue indices); -1 where the residue has no such atom."""
        residues = np.asarray(residues, dtype=np.int64)
        tid, templates = self.residue_templates()
        try:
            want = self.atom_names.index(atom_name)
        except ValueError:
            return np.full(len(residues), -1, dtype=np.int64)
        local = np.full(len(templates), -1, dtype=np.int64)
        for t, (_, row) in enumerate(templates):
            hits = np.nonzero(row == want)[0]
            if len(hits):
                local[t] = hits[0]
        off = local[tid[residues]]
        return np.where(off >= 0, self.res_first[residues] + off, -1)

    def __eq__(self, other):
        if not isinstance(other, Topology):
            return NotImplemented
        return (self.n_atoms == other.n_atoms and self.n_residues == other.n_residues
                and np.array_equal(self.res_first, other.res_first)
                and [self.res_names[i] for i in self.res_name_id] == [other.res_names[i] for i in other.res_name_id]
                and [self.atom_names[i] for i in self.atom_name_id] == [other.atom_names[i] for i in other.atom_name_id])

    __hash__ = None


def lengths_and_angles(vectors):
    """Unit-cell lengths (nm) and angles (degrees) from (F,3,3) row vectors, float32."""
    v = np.asarray(vectors, dtype=np.float64)
    a, b, c = v[:, 0], v[:, 1], v[:, 2]
    la, lb, lc = (np.sqrt((x * x).sum(axis=1)) for x in (a, b, c))

    def ang(p, q, lp, lq):
        return np.degrees(np.arccos((p * q).sum(axis=1) / (lp * lq)))

    lengths = np.stack([la, lb, lc], axis=1).astype(np.float32)
    angles = np.stack([ang(b, c, lb, lc), ang(a, c, la, lc), ang(a, b, la, lb)], axis=1).astype(np.float32)
    return lengths, angles


class Trajectory:
    """Coordinates + unit cell + time for one topology; host and/or device resident."""

    def __init__(self, xyz=None, topology=None, time=None, unitcell_vectors=None, xyz_device=None):
        if xyz is None and xyz_device is None:
            raise ValueError("Trajectory needs coordinates")
        self._xyz = None if xyz is None else np.ascontiguousarray(xyz, dtype=np.float32)
        self._xyz_dev = xyz_device
        self.topology = topology
        shape = self._xyz.shape if self._xyz is not None else tuple(xyz_device.shape)
        if len(shape) != 3 or shape[2] != 3:
            raise ValueError("xyz must have shape (n_frames, n_atoms, 3)")
        if topology is not None and topology.n_atoms != shape[1]:
            raise ValueError("Number of atoms does not match between topology and trajectory files")
        self._shape = shape
        self.time = np.arange(shape[0], dtype=np.float32) if time is None else np.asarray(time)
        self.unitcell_vectors = unitcell_vectors

    # ---- coordinates
    @property
    def n_frames(self):
        return self._shape[0]

    @property
    def n_atoms(self):
        return self._shape[1]

    def __len__(self):
        return self.n_frames

    @property
    def xyz(self):
        """Host copy, (F, N, 3) float32 (device -> host on first use)."""
        if self._xyz is None:
            self._xyz = self._xyz_dev.cpu().numpy()
        return self._xyz

    @property
    def on_device(self):
        return self._xyz_dev is not None

    def device_xyz(self, device=None):
        """Device copy as a torch tensor (host -> device on first use)."""
        if self._xyz_dev is None:
            from . import device as dev
            self._xyz_dev = dev.to_device(self._xyz, device)
        return self._xyz_dev

    def global_box_flags(self, device=None):
        """``(every box length of every frame is non-zero, a unit cell is present, mdtraj's orthorhombic test)``
        decided over the WHOLE trajectory: with the frames sharded across GPUs the reference's decisions
        (``mapping.py:406-408``; mdtraj's ``np.allclose(angles, 90)``) hold for all ranks together, so the three
        flags are combined in one collective.  Evaluated once per trajectory and handed on to the CG trajectory
        mapped from it (same unit cell)."""
        flags = getattr(self, "_global_flags", None)
        if flags is None:
            from . import dist
            lengths = self.unitcell_lengths
            flags = dist.all_true_many([lengths is not None and bool(np.all(lengths)), self._vectors is not None,
                                        bool(self.orthorhombic)], device=device)
            self._global_flags = flags
        return flags

    def share_unitcell(self, other):
        """Take over the unit cell of ``other`` (same frames) with everything already derived from it -- lengths and
        angles, the orthorhombic test, the device copies, the decisions taken over all ranks -- instead of deriving
        them again: the CG trajectory mapped from a trajectory has its box."""
        if other.n_frames != self.n_frames:
            raise ValueError("unit cells can only be shared between trajectories of the same frames")
        for name in ("_vectors", "_vectors_dev", "_lengths_dev", "_ortho", "_global_flags", "unitcell_lengths",
                     "unitcell_angles"):
            setattr(self, name, getattr(other, name, None))

    def drop_device(self):
        """Forget the device copies (the host copy stays): the next use uploads again."""
        if self._xyz is not None:
            self._xyz_dev = None
        self._vectors_dev = None

    def device_box_lengths(self, device):
        """(F, 3) float32 box lengths on the device (uploaded once), or None without a unit cell."""
        if self.unitcell_lengths is None:
            return None
        cached = getattr(self, "_lengths_dev", None)
        if cached is None or cached.device != device:
            from . import device as dev
            cached = dev.to_device(self.unitcell_lengths, device, dtype=np.float32)
            self._lengths_dev = cached
        return cached

    def device_box_vectors(self, device):
        """(F, 9) float32 box vectors on the device (uploaded once), or None without a unit cell."""
        if self._vectors is None:
            return None
        cached = getattr(self, "_vectors_dev", None)
        if cached is None or cached.device != device:
            from . import device as dev
            cached = dev.to_device(self._vectors.reshape(self.n_frames, 9), device, dtype=np.float32)
            self._vectors_dev = cached
        return cached

    @property
    def orthorhombic(self):
        """mdtraj's test for the orthorhombic minimum-image fast path: every unit-cell angle of
        every frame is close to 90 degrees (``np.allclose``); evaluated once."""
        if self._vectors is None:
            return True
        if getattr(self, "_ortho", None) is None:
            self._ortho = bool(np.allclose(self.unitcell_angles, 90))
        return self._ortho

    # ---- unit cell
    @property
    def unitcell_vectors(self):
        return self._vectors

    @unitcell_vectors.setter
    def unitcell_vectors(self, vectors):
        if vectors is None or not np.any(vectors):
            self._vectors = None
            self._vectors_dev = None
            self._lengths_dev = None
            self._ortho = None
            self._global_flags = None
            self.unitcell_lengths = None
            self.unitcell_angles = None
            return
        vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        if vectors.shape != (self.n_frames, 3, 3):
            raise TypeError("unitcell_vectors must have shape (n_frames, 3, 3)")
        self._vectors = vectors
        self._vectors_dev = None
        self._lengths_dev = None
        self._ortho = None
        self._global_flags = None
        self.unitcell_lengths, self.unitcell_angles = lengths_and_angles(vectors)

    def slice(self, key):
        """Sub-trajectory of the selected frame(s) (host side)."""
        if isinstance(key, (int, np.integer)):
            key = slice(int(key), int(key) + 1)
        vec = None if self._vectors is None else self._vectors[key]
        return Trajectory(self.xyz[key], self.topology, self.time[key], vec)

    def __getitem__(self, key):
        return self.slice(key)

    def __iadd__(self, other):
        """Append frames (``frame.py:268-269``)."""
        if other.n_atoms != self.n_atoms:
            raise ValueError("Number of atoms does not match")
        xyz = np.concatenate([self.xyz, other.xyz], axis=0)
        time = np.concatenate([self.time, other.time])
        vec = None
        if self._vectors is not None and other._vectors is not None:
            vec = np.concatenate([self._vectors, other._vectors], axis=0)
        return Trajectory(xyz, self.topology, time, vec)
