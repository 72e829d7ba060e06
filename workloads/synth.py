"""Synthetic systems of the shapes named in BASELINE.json.

Used by the benchmark and the parity tests; nothing here is on the hot path.
A system is a topology (array-backed, so 20M atoms are cheap), the text of a
``.map`` / ``.bnd`` pair describing it, and a coordinate generator.  Molecule
centres are uniform in an orthorhombic box, atoms are ``centre + N(0, sigma)``
and are then *wrapped into the box*, so molecules straddle the periodic
boundaries and the unwrap of the mapping is exercised; the box jitters by +-1 %
per frame.  Chain molecules (config S3) are random walks instead, so angles and
dihedrals are non-degenerate and some dihedrals straddle +-pi.
"""

import dataclasses
import pathlib
import tempfile

import numpy as np

from pycgtool_b200.topology import Topology

MARTINI_12 = ["NC3", "PO4", "GL1", "GL2", "C1A", "C2A", "C3A", "C4A", "C1B", "C2B", "C3B", "C4B"]
MARTINI_8 = ["NC3", "PO4", "GL1", "GL2", "C1A", "C2A", "C1B", "C2B"]
LIPID_BONDS_12 = [("NC3", "PO4"), ("PO4", "GL1"), ("GL1", "GL2"), ("GL1", "C1A"), ("C1A", "C2A"), ("C2A", "C3A"),
                  ("C3A", "C4A"), ("GL2", "C1B"), ("C1B", "C2B"), ("C2B", "C3B"), ("C3B", "C4B")]
LIPID_BONDS_8 = [("NC3", "PO4"), ("PO4", "GL1"), ("GL1", "GL2"), ("GL2", "C1A"), ("C1A", "C2A"), ("C2A", "C1B"),
                 ("C1B", "C2B")]

#: molecule type -> (bead names, atoms per bead, bonds)
MOLECULES = {
    "DOPC": (MARTINI_12, [12, 12, 12, 12, 11, 11, 11, 11, 12, 12, 11, 11], LIPID_BONDS_12),   # 138 atoms
    "POPC": (MARTINI_12, [12, 12] + [11] * 10, LIPID_BONDS_12),                               # 134 atoms
    "DPPE": (MARTINI_12, [11, 11, 11, 11, 10, 10, 10, 10, 11, 10, 10, 10], LIPID_BONDS_12),   # 125 atoms
    "CHOL": (MARTINI_8, [10, 10, 9, 9, 9, 9, 9, 9], LIPID_BONDS_8),                           # 74 atoms
    "W4": (["W"], [12], []),                                                                  # 4 waters -> 1 bead
}


@dataclasses.dataclass
class SyntheticSystem:
    name: str
    topology: Topology
    map_text: str
    bnd_text: str
    box: np.ndarray           # (3,) nominal orthorhombic box lengths, nm
    chain: bool = False       # coordinates are random-walk chains (S3)
    sigma: float = 0.25
    seed: int = 0

    @property
    def n_atoms(self):
        return self.topology.n_atoms

    def write_files(self, directory=None):
        """Write the .map / .bnd pair; returns their paths."""
        directory = pathlib.Path(directory or tempfile.mkdtemp(prefix="cgb_synth_"))
        directory.mkdir(parents=True, exist_ok=True)
        map_path, bnd_path = directory / f"{self.name}.map", directory / f"{self.name}.bnd"
        map_path.write_text(self.map_text)
        bnd_path.write_text(self.bnd_text)
        return map_path, bnd_path

    # ------------------------------------------------------------------ coordinates
    def box_vectors(self, n_frames, first_frame=0):
        """(F, 3, 3) float32 orthorhombic box with +-1 % jitter, deterministic per frame index."""
        rng = np.random.default_rng(self.seed + 7919)
        jitter = rng.uniform(0.99, 1.01, size=(first_frame + n_frames, 3))[first_frame:]
        vec = np.zeros((n_frames, 3, 3), dtype=np.float32)
        for d in range(3):
            vec[:, d, d] = (self.box[d] * jitter[:, d]).astype(np.float32)
        return vec

    def coordinates_host(self, n_frames, first_frame=0):
        """(F, N, 3) float32 on the host (small systems / tests)."""
        top = self.topology
        res_len = np.diff(top.res_first)
        box = self.box_vectors(n_frames, first_frame)
        out = np.empty((n_frames, top.n_atoms, 3), dtype=np.float32)
        for k in range(n_frames):
            rng = np.random.default_rng((self.seed, first_frame + k))
            length = np.array([box[k, 0, 0], box[k, 1, 1], box[k, 2, 2]], dtype=np.float64)
            centres = rng.uniform(0, 1, size=(top.n_residues, 3)) * length
            if self.chain:
                steps = rng.normal(size=(top.n_atoms, 3))
                steps /= np.linalg.norm(steps, axis=1, keepdims=True)
                steps *= rng.normal(0.12, 0.01, size=(top.n_atoms, 1))
                walk = np.cumsum(steps, axis=0)
                walk -= np.repeat(walk[top.res_first[:-1]], res_len, axis=0)
                pos = np.repeat(centres, res_len, axis=0) + walk
            else:
                pos = np.repeat(centres, res_len, axis=0) + rng.normal(0, self.sigma, size=(top.n_atoms, 3))
            out[k] = np.mod(pos, length).astype(np.float32)
        return out

    def coordinates_device(self, n_frames, device, first_frame=0, chunk=None):
        """(F, N, 3) float32 CUDA tensor, generated on the device in frame chunks."""
        import torch
        top = self.topology
        if chunk is None:
            # ~256 MB of coordinates per chunk, a power of two that depends on the system only
            chunk = 1
            while chunk < 16384 and 2 * chunk * top.n_atoms * 12 <= (256 << 20):
                chunk *= 2
        gen = torch.Generator(device=device)
        res_len = torch.from_numpy(np.diff(top.res_first)).to(device)
        res_first = torch.from_numpy(top.res_first[:-1].copy()).to(device)
        out = torch.empty((n_frames, top.n_atoms, 3), dtype=torch.float32, device=device)
        res_of_atom = torch.repeat_interleave(torch.arange(top.n_residues, device=device), res_len)
        # Frames are generated in absolute chunks [c*chunk, (c+1)*chunk) seeded by c, so that any
        # frame shard (first_frame, n_frames) sees exactly the coordinates of the whole trajectory.
        last = first_frame + n_frames
        for c in range(first_frame // chunk, (last + chunk - 1) // chunk):
            a0 = c * chunk                                   # absolute first frame of the chunk
            lo, hi = max(a0, first_frame), min(a0 + chunk, last)
            gen.manual_seed(self.seed * 1000003 + c)
            n = chunk
            rng = np.random.default_rng(self.seed + 7919)
            jitter = rng.uniform(0.99, 1.01, size=(a0 + chunk, 3))[a0:]
            clen = torch.from_numpy((self.box[None, :] * jitter).astype(np.float32)).to(device)   # (chunk, 3)
            centres = torch.rand((n, top.n_residues, 3), generator=gen, device=device) * clen[:, None, :]
            if self.chain:
                steps = torch.randn((n, top.n_atoms, 3), generator=gen, device=device)
                steps = steps / steps.norm(dim=2, keepdim=True)
                steps = steps * (0.12 + 0.01 * torch.randn((n, top.n_atoms, 1), generator=gen, device=device))
                walk = torch.cumsum(steps, dim=1)
                walk = walk - walk[:, res_first][:, res_of_atom]
                pos = centres[:, res_of_atom] + walk
            else:
                pos = centres[:, res_of_atom] + self.sigma * torch.randn(
                    (n, top.n_atoms, 3), generator=gen, device=device)
            pos = torch.remainder(pos, clen[:, None, :])
            out[lo - first_frame:hi - first_frame] = pos[lo - a0:hi - a0]
            del centres, pos
        return out


def _build(name, composition, box, seed, with_bonds=True, chain=False, sigma=0.25):
    """``composition``: list of (molecule type, count), laid out in that order."""
    res_names, atom_tab = [], {}
    rid, first, aid = [], [0], []
    map_lines, bnd_lines = [], []
    templates = {}
    for mol, _ in composition:
        if mol in templates:
            continue
        beads, sizes, bonds = MOLECULES[mol]
        names, k = [], 0
        map_lines.append(f"[ {mol} ]")
        for bead, size in zip(beads, sizes):
            atoms = [f"A{k + i}" for i in range(size)]
            k += size
            names.extend(atoms)
            map_lines.append(f"{bead} P4 " + " ".join(atoms))
        map_lines.append("")
        if with_bonds and bonds:
            bnd_lines.append(f"[ {mol} ]")
            bnd_lines.extend(f"{a} {b}" for a, b in bonds)
            bnd_lines.append("")
        ids = np.array([atom_tab.setdefault(a, len(atom_tab)) for a in names], dtype=np.int32)
        res_names.append(mol)
        templates[mol] = (len(res_names) - 1, ids)
    blocks_rid, blocks_aid, lens = [], [], []
    for mol, count in composition:
        name_id, ids = templates[mol]
        blocks_rid.append(np.full(count, name_id, dtype=np.int32))
        blocks_aid.append(np.tile(ids, count))
        lens.append(np.full(count, len(ids), dtype=np.int64))
    rid = np.concatenate(blocks_rid)
    aid = np.concatenate(blocks_aid)
    first = np.concatenate([[0], np.cumsum(np.concatenate(lens))])
    top = Topology.from_arrays(res_names, rid, first, list(atom_tab), aid)
    return SyntheticSystem(name, top, "\n".join(map_lines) + "\n", "\n".join(bnd_lines) + "\n",
                           np.array(box, dtype=np.float64), chain=chain, sigma=sigma, seed=seed)


def membrane(n_lipids_each=576, n_water=0, seed=1234, name="S2"):
    """S2 / S4: DOPC + POPC bilayer patch, optionally solvated with W4 residues."""
    comp = [("DOPC", n_lipids_each), ("POPC", n_lipids_each)]
    if n_water:
        comp.append(("W4", n_water))
    n_atoms = n_lipids_each * (138 + 134) + 12 * n_water
    edge = max(4.0, (n_atoms / 100.0) ** (1.0 / 3.0))     # ~100 atoms / nm^3
    return _build(name, comp, (edge, edge, edge), seed)


def s2():
    """Config 2: 1,152-lipid bilayer, MARTINI mapping (N = 156,672, B = 13,824)."""
    return membrane(576, 0, seed=1234, name="S2")


def s4():
    """Config 4: 5M-atom solvated membrane (N = 4,999,872, B = 417,424)."""
    return membrane(576, 403600, seed=3456, name="S4")


def s5(scale=1.0):
    """Config 5: 20M-atom multi-lipid membrane (N = 19,999,992, B = 1,710,442 at scale 1)."""
    n_lip = max(1, int(round(9216 * scale)))
    n_w = max(1, int(round(1304938 * scale)))
    comp = [("DOPC", n_lip), ("POPC", n_lip), ("DPPE", n_lip), ("CHOL", n_lip), ("W4", n_w)]
    n_atoms = n_lip * (138 + 134 + 125 + 74) + 12 * n_w
    edge = (n_atoms / 100.0) ** (1.0 / 3.0)
    return _build("S5", comp, (edge, edge, edge), seed=4567)


def small_molecule(n_beads=24, atoms_per_bead=4, seed=2345):
    """Config 3: one chain molecule, 96 atoms -> 24 beads; 23 bonds + 22 angles + 21 dihedrals."""
    beads = [f"B{i}" for i in range(n_beads)]
    MOLECULES["MOL"] = (beads, [atoms_per_bead] * n_beads, [])
    sys = _build("S3", [("MOL", 1)], (6.0, 6.0, 6.0), seed, with_bonds=False, chain=True)
    lines = ["[ MOL ]"]
    lines += [f"{beads[i]} {beads[i + 1]}" for i in range(n_beads - 1)]
    lines += [f"{beads[i]} {beads[i + 1]} {beads[i + 2]}" for i in range(n_beads - 2)]
    lines += [f"{beads[i]} {beads[i + 1]} {beads[i + 2]} {beads[i + 3]}" for i in range(n_beads - 3)]
    sys.bnd_text = "\n".join(lines) + "\n"
    return sys


class Options:
    """Duck-typed options object accepted by ``Mapping`` and ``BondSet`` (as the reference's
    argparse namespace / test ``DummyOptions``)."""
    map_center = "geom"
    virtual_map_center = "geom"
    constr_threshold = 100000
    generate_angles = True
    generate_dihedrals = False
