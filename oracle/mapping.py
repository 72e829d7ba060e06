"""Oracle for atom->bead mapping.  TEST INFRASTRUCTURE ONLY.

Restates ``pycgtool/mapping.py``: the ``.map`` parse and weight rules
(``:26-127,135-357``), the residue/bead matching of ``_cg_frame_setup`` and
``apply`` (``:359-443``) and ``calc_coords_weight`` (``:446-463``).  The loop
structure of ``apply`` (Python over residues x beads, NumPy over frames) is kept
on purpose: this is also the CPU baseline that ``bench.py`` times.
"""

import copy
import dataclasses
import typing

import numpy as np

from . import elements, fileio, system


@dataclasses.dataclass
class Bead:
    name: str
    num: int
    type: str
    atoms: typing.List[str]
    charge: float = 0.0
    mass: float = 0.0
    virtual: bool = False
    weights_by_center: dict = dataclasses.field(default_factory=dict)
    weights: typing.Optional[np.ndarray] = None

    def __post_init__(self):
        n = len(self.atoms)
        # mapping.py:55-58 -- geom and first are float32
        self.weights_by_center["geom"] = np.array([1.0 / n] * n, dtype=np.float32)
        self.weights_by_center["first"] = np.array([1.0] + [0.0] * (n - 1), dtype=np.float32)
        self.weights = self.weights_by_center["geom"]


def _beads_from_section(rows):
    """``Mapping._mol_map_from_section`` (mapping.py:250-286)."""
    beads, manual_charges = [], False
    for num, row in enumerate(rows):
        virtual = False
        if row[0].startswith("@"):
            if row[0] != "@v":
                raise ValueError(f'Invalid line prefix "{row[0]}" in mapping file')
            virtual = True
            row = row[1:]
        name, typ, first, *atoms = row
        charge = 0
        try:
            charge = float(first)      # optional charge column (mapping.py:271-275)
            manual_charges = True
        except ValueError:
            atoms.insert(0, first)
        if not atoms:
            raise ValueError(f"Bead {name} specification contains no atoms")
        beads.append(Bead(name, num, typ, atoms, charge=charge, virtual=virtual))
    return beads, manual_charges


def _guess_masses(beads):
    """``BeadMap.guess_atom_masses`` + ``Mapping._guess_atom_masses`` (mapping.py:77-106,334-357)."""
    for bead in beads:
        if bead.virtual or bead.mass:
            continue
        masses = np.zeros(len(bead.atoms), dtype=np.float32)
        for i, atom in enumerate(bead.atoms):
            try:
                masses[i] = elements.mass_by_symbol(atom[:2])
            except KeyError:
                try:
                    masses[i] = elements.mass_by_symbol(atom[0])
                except KeyError:
                    raise RuntimeError(f"Mass of atom {atom} could not be automatically assigned, "
                                       "map_center=mass is not available.")
        bead.mass = sum(masses)
        if not np.all(masses):
            raise RuntimeError("Some atom masses could not be automatically assigned, "
                               "map_center=mass is not available")
        masses /= bead.mass
        bead.weights_by_center["mass"] = masses
    for bead in beads:
        if bead.virtual:
            m = np.array([b.mass for b in beads if b.name in bead.atoms], dtype=np.float32)
            bead.weights_by_center["mass"] = m / sum(m)


def _apply_itp(beads, itp_mol, manual_charges):
    """``Mapping._itp_section_into_mol_map`` (mapping.py:213-248): float64 mass weights."""
    table = {toks[4]: (float(toks[6]), float(toks[7])) for toks in itp_mol["atoms"]}
    for bead in beads:
        if bead.virtual:
            continue
        m = np.array([table[a][1] for a in bead.atoms])
        bead.mass = sum(m)
        m /= bead.mass
        bead.weights_by_center["mass"] = m
        if not manual_charges:
            for a in bead.atoms:
                bead.charge += table[a][0]
    for bead in beads:
        if bead.virtual:
            m = np.array([b.mass for b in beads if b.name in bead.atoms])
            bead.weights_by_center["mass"] = m / sum(m)
            if not manual_charges:
                bead.charge = sum(b.charge for b in beads if b.name in bead.atoms)


def parse_map(path, map_center="geom", virtual_map_center="geom", itp_path=None):
    """``Mapping.__init__`` (mapping.py:135-169) -> ``{molecule: [Bead]}``."""
    mols, manual = {}, {}
    for mol, rows in fileio.read_cfg(path).items():
        mols[mol], manual[mol] = _beads_from_section(rows)
    masses_set = False
    if itp_path is not None:
        itp = fileio.read_itp(itp_path)
        for mol, beads in mols.items():
            if mol in itp:
                _apply_itp(beads, itp[mol], manual[mol])
        masses_set = True
    if (map_center == "mass" or virtual_map_center == "mass") and not masses_set:
        for beads in mols.values():
            _guess_masses(beads)
    for beads in mols.values():            # mapping.py:200-211
        for bead in beads:
            bead.weights = bead.weights_by_center[virtual_map_center if bead.virtual else map_center]
    # mapping.py:288-317 -- an alias keyed by mdtraj's renamed residue, atoms renamed too
    for mol in list(mols):
        if mol not in fileio.RESIDUE_RENAMES:
            continue
        new = fileio.rename_residue(mol)
        alias = copy.deepcopy(mols[mol])
        for bead in alias:
            bead.atoms = [fileio.rename_atom(new, a) for a in bead.atoms]
        mols[new] = alias
    return mols


def calc_coords_weight(ref_coords, coords, weights, box=None):
    """``calc_coords_weight`` (mapping.py:446-463), same NumPy expression order:
    subtract ref, minimum image with box *lengths* (IEEE divide, round-half-even,
    un-fused multiply/subtract), weight, sum over atoms in order, add ref."""
    vectors = coords - ref_coords
    if box is not None:
        vectors -= box * np.rint(vectors / box)
    weights_t = weights[np.newaxis, np.newaxis].T
    result = np.sum(weights_t * vectors, axis=0)
    result += ref_coords
    return result


def mapped_residues(mols, sysm):
    """AA residues whose name has a mapping, in topology order (mapping.py:370-389,411)."""
    return [res for res in sysm.residues if res.name in mols]


def assignment(mols, sysm):
    """Integer assignment tables implied by the name lookups of ``apply``
    (mapping.py:411-438): CSR over output beads in output order.

    Returns ``(bead_ptr int64 (B+1,), index int64 (nnz,), is_virtual bool (B,))``;
    for real beads ``index`` holds global atom indices, for virtual beads it
    holds global CG bead indices.
    """
    ptr, idx, virt = [0], [], []
    cg_first = 0
    for res in mapped_residues(mols, sysm):
        beads = mols[res.name]
        names = [b.name for b in beads]
        for bead in beads:
            if bead.virtual:
                idx.extend(cg_first + names.index(a) for a in bead.atoms)
            else:
                idx.extend(res.index_of(a) for a in bead.atoms)
            ptr.append(len(idx))
            virt.append(bead.virtual)
        cg_first += len(beads)
    return np.array(ptr, dtype=np.int64), np.array(idx, dtype=np.int64), np.array(virt, dtype=bool)


def apply(mols, sysm):
    """``Mapping.apply`` (mapping.py:391-443) -> CG ``System``.

    Box handling (mapping.py:406-408): lengths only, disabled for the whole
    call when the frame has no box or any length of any frame is zero.
    """
    box = sysm.unitcell_lengths
    if box is None or not np.all(box):
        box = None
    xyz = sysm.xyz
    cg_residues, columns = [], []
    for res in mapped_residues(mols, sysm):
        beads = mols[res.name]
        cg_residues.append((res.name, [b.name for b in beads]))
        names = [b.name for b in beads]
        local = [None] * len(beads)
        for k, bead in enumerate(beads):         # real beads (mapping.py:415-429)
            if bead.virtual:
                continue
            ref = xyz[:, res.index_of(bead.atoms[0])]
            if len(bead.atoms) == 1:
                local[k] = ref
            else:
                coords = np.array([xyz[:, res.index_of(a)] for a in bead.atoms], dtype=np.float32)
                local[k] = calc_coords_weight(ref, coords, bead.weights, box)
        for k, bead in enumerate(beads):         # virtual beads (mapping.py:431-438)
            if bead.virtual:
                coords = np.array([local[names.index(a)] for a in bead.atoms], dtype=np.float32)
                local[k] = calc_coords_weight(coords[0], coords, bead.weights, box)
        columns.extend(local)
    n_frames = xyz.shape[0]
    if columns:                                  # frame.py:246-259
        cg_xyz = np.array(columns, dtype=np.float32).swapaxes(0, 1)
    else:
        cg_xyz = np.empty((n_frames, 0, 3), dtype=np.float32)
    return system.from_residue_lists(cg_residues, np.ascontiguousarray(cg_xyz), sysm.box_vectors, sysm.time)
