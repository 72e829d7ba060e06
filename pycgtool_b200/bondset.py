"""Bonded-term measurement and Boltzmann inversion with the reference's ``BondSet`` API.

``BondSet(filename, options)`` parses a ``.bnd`` file and generates angles /
dihedrals from the bond graph exactly as ``pycgtool/bondset.py:149-234`` does.
``BondSet.apply(frame)`` resolves the bead names of every term (with ``+``/``-``
neighbour-residue prefixes, ``bondset.py:509-528``) into one integer column
table per topology, then a single fused GPU pass (``cgb_measure``) produces the
values matrix and the float64 statistics accumulators -- the work the
reference spreads over ``mdtraj.compute_*`` plus ~15 NumPy passes
(``bondset.py:530``, ``util.py:24-53``).  ``boltzmann_invert`` reduces the
accumulators across GPUs when the trajectory is sharded by frame, finishes the
circular variance of dihedrals and evaluates the functional forms in the
``cgb_boltzmann`` epilogue kernel.
"""

import itertools
import logging
import math

import numpy as np

from . import _lib, device, dist
from .functionalforms import get_functional_forms
from .mapping import VirtualMap
from .parsers import CFG
from .util import circular_mean, circular_variance, extend_graph_chain, file_write_lines, transpose_and_sample

logger = logging.getLogger(__name__)

#: default fixed-bin histogram layout: bins, and [lo, hi] per term kind (nm / radians)
HIST_BINS = 128
HIST_RANGES = {2: (0.0, 2.0), 3: (0.0, math.pi), 4: (-math.pi, math.pi)}


class Bond:
    """One bonded term (length, angle or dihedral by number of atoms) and its results."""

    def __init__(self, atoms, atom_numbers=None, func_form=None):
        self.atoms = atoms
        self.atom_numbers = atom_numbers
        self.eqm = None
        self.fconst = None
        self._func_form = func_form
        self.gromacs_type_id = func_form.gromacs_type_id_by_natoms(len(atoms))
        self._values = []
        self._source = None      # (BondSet state, bond id) when measured on the GPU
        self.histogram = None    # (counts uint64 (nbins,), edges float64 (nbins+1,)) after boltzmann_invert

    def __iter__(self):
        return iter(self.atoms)

    # ---- values: frame-major float32, fetched from the GPU on first use (bondset.py:531)
    @property
    def values(self):
        if self._values is None:
            state, bond_id = self._source
            self._values = state.fetch_values(bond_id)
        return self._values

    @values.setter
    def values(self, new):
        self._values = new
        self._source = None

    def boltzmann_invert(self, temp=310):
        """Equilibrium value and force constant from the measured values (``bondset.py:62-91``)."""
        if self._source is not None:
            state, bond_id = self._source
            if state.n_values(bond_id) == 0:
                raise ValueError("No bonds were measured between beads {0}".format(self.atoms))
            state.invert(temp)
            state.assign(self, bond_id, temp)
            return
        if len(self.values) == 0:
            raise ValueError("No bonds were measured between beads {0}".format(self.atoms))
        mean, var = _stats_of_values(np.asarray(self.values, dtype=np.float32), len(self.atoms))
        _finish_bond(self, mean, var, temp)

    def __repr__(self):
        try:
            return "<Bond containing atoms {0} with r_0 {1:.3f} and force constant {2:.3e}>".format(
                ", ".join(self.atoms), self.eqm, self.fconst)
        except (AttributeError, TypeError):
            return "<Bond containing atoms {0}>".format(", ".join(self.atoms))


def _finish_bond(bond, mean, var, temp, device_result=None):
    """Set ``eqm`` / ``fconst`` from the statistics.

    Built-in forms take the epilogue kernel's numbers; a user-defined form is
    evaluated in Python with mean / variance hooks that return the GPU statistics
    (the reference's own injection point, ``functionalforms.py:30-41``).
    """
    form = bond._func_form
    if device_result is not None and form.device_form is not None:
        bond.eqm, bond.fconst = device_result
        return
    saved = form._mean_func, form._variance_func
    form._mean_func, form._variance_func = (lambda _values: mean), (lambda _values: var)
    try:
        eqm = form.eqm(None, temp)
        if len(bond.atoms) == 4:
            eqm -= math.pi
            eqm = ((eqm + math.pi) % (2 * math.pi)) - math.pi
        bond.eqm = eqm
        try:
            bond.fconst = form.fconst(None, temp)
        except (FloatingPointError, ZeroDivisionError):
            bond.fconst = math.inf   # variance is zero: a single distinct value (bondset.py:89-91)
    finally:
        form._mean_func, form._variance_func = saved


def _stats_of_values(values, natoms):
    """Mean / variance of a host array through the GPU accumulators (``cgb_value_stats``)."""
    import torch
    lib = _lib.load()
    dev = device.default_device()
    vals = device.to_device(values, dev)
    row = torch.zeros(_lib.NPART, dtype=torch.float64, device=dev)
    finite = values[np.isfinite(values)]
    shift = float(finite[0]) if len(finite) else 0.0
    stream = device.stream_handle()
    _lib.check(lib.cgb_value_stats(device.dptr(vals), len(values), natoms, shift, device.dptr(row), stream),
               "cgb_value_stats")
    if natoms == 4:
        zero = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.cgb_circular_pass2(device.dptr(vals), len(values), 1, device.dptr(zero), 1,
                                          device.dptr(row), stream), "cgb_circular_pass2")
    out = torch.empty(4, dtype=torch.float64, device=dev)
    nat = torch.tensor([natoms], dtype=torch.int32, device=dev)
    form = torch.tensor([_lib.FORM_STATS_ONLY], dtype=torch.int32, device=dev)
    sh = torch.tensor([shift], dtype=torch.float32, device=dev)
    _lib.check(lib.cgb_boltzmann(device.dptr(row), device.dptr(sh), device.dptr(nat), device.dptr(form),
                                 310.0, 1, device.dptr(out), stream), "cgb_boltzmann")
    res = out.cpu().numpy()
    return float(res[2]), float(res[3])


class MeasurePlan:
    """Integer column table for one (bond set, topology) pair."""

    def __init__(self):
        self.n_beads = 0
        self.n_bonds = 0
        self.counts = (0, 0, 0)       # columns per kind (lengths, angles, dihedrals)
        self.col_beads = None         # int32 (M, 4)
        self.col_bond = None          # int32 (M,)
        self.col_slot = None          # int32 (M,) compact Bond number inside each tile of _lib.MEASURE_TILE columns of a kind
        self.max_slots = 0
        self.tile_span = None         # int32 (T, 2): (first bead, beads) of the range each tile gathers from
        self.max_span = 0
        self.bond_cols = []           # per bond: int64 column ids in instance order
        self.bond_indices = []        # per bond: int32 (n_inst, natoms) -- what mdtraj would receive
        self.natoms = None            # int32 (n_bonds,)
        self._dev = {}


class _MeasureState:
    """Device buffers of one ``BondSet.apply`` call."""

    def __init__(self, bondset, plan, n_frames, values, partials, hist, shift, hist_lo, hist_hi, dev):
        self.bondset, self.plan, self.n_frames = bondset, plan, n_frames
        self.values, self.partials, self.hist, self.shift = values, partials, hist, shift
        self.hist_lo, self.hist_hi = hist_lo, hist_hi
        self.dev = dev
        self.reduced = False
        self.result = None
        self.result_temp = None
        self.hist_host = None

    def n_values(self, bond_id):
        return self.n_frames * len(self.plan.bond_cols[bond_id])

    def fetch_values(self, bond_id):
        import torch
        cols = self.plan.bond_cols[bond_id]
        if len(cols) == 0 or self.n_frames == 0:
            return np.zeros(0, dtype=np.float32)
        # values are stored as one (F, n_kind) plane per kind (see include/pycgtool_b200.h)
        n2, n3, n4 = self.plan.counts
        kind = int(self.plan.natoms[bond_id])
        col0, n_k = {2: (0, n2), 3: (n2, n3), 4: (n2 + n3, n4)}[kind]
        plane = self.values[self.n_frames * col0:self.n_frames * (col0 + n_k)].view(self.n_frames, n_k)
        idx = torch.from_numpy(cols - col0).to(self.dev)
        return plane.index_select(1, idx).reshape(-1).cpu().numpy()

    def invert(self, temp):
        """Cross-GPU reduction, dihedral pass 2 and the epilogue kernel (idempotent per ``temp``)."""
        import torch
        if self.result is not None and self.result_temp == temp:
            return
        lib = _lib.load()
        plan = self.plan
        stream = device.stream_handle()
        d = BondSet._device_plan(plan, self.dev)
        n2, n3, n4 = plan.counts
        if not self.reduced:
            # frames are sharded across GPUs: sum the per-bond moments (and histograms) once
            dist.allreduce_sum(self.partials)
            if self.hist is not None:
                dist.allreduce_sum(self.hist)
            if n4:
                # dihedrals: the circular mean is final now, so finish <wrap(v - mean)^2> (util.py:49-53)
                plane4 = self.values[self.n_frames * (n2 + n3):]
                _lib.check(lib.cgb_circular_pass2(device.dptr(plane4), self.n_frames, n4,
                                                  device.dptr(d["col_bond"][n2 + n3:]), n4,
                                                  device.dptr(self.partials), stream), "cgb_circular_pass2")
                if dist.is_distributed():
                    wrapped = self.partials[:, 5].contiguous()
                    dist.allreduce_sum(wrapped)
                    self.partials[:, 5] = wrapped
            self.reduced = True
        out = torch.empty((plan.n_bonds, 4), dtype=torch.float64, device=self.dev)
        _lib.check(lib.cgb_boltzmann(device.dptr(self.partials), device.dptr(self.shift), device.dptr(d["natoms"]),
                                     device.dptr(d["forms"]), float(temp), plan.n_bonds, device.dptr(out), stream),
                   "cgb_boltzmann")
        self.result = out.cpu().numpy()
        self.result_temp = temp
        if self.hist is not None and self.hist_host is None:
            self.hist_host = self.hist.cpu().numpy().view(np.uint64)

    def assign(self, bond, bond_id, temp):
        eqm, fconst, mean, var = (float(x) for x in self.result[bond_id])
        if math.isnan(mean):
            return          # nothing finite was measured: leave eqm / fconst as None
        _finish_bond(bond, mean, var, temp, device_result=(eqm, fconst))
        if self.hist_host is not None:
            edges = np.linspace(float(self.hist_lo[bond_id]), float(self.hist_hi[bond_id]),
                                self.hist_host.shape[1] + 1)
            bond.histogram = (self.hist_host[bond_id].copy(), edges)


class BondSet:
    """Dictionary ``molecule name -> [Bond]`` plus the GPU-backed ``apply`` / ``boltzmann_invert``."""

    def _get_default_func_forms(self, options):
        """Forms by number of atoms (``bondset.py:108-147``)."""
        if getattr(options, "default_fc", False):
            names = ["MartiniDefaultLength", "MartiniDefaultAngle", "MartiniDefaultDihedral"]
        else:
            names = ["Harmonic", "CosHarmonic", "Harmonic"]
        known = get_functional_forms()
        for slot, attr in enumerate(("length_form", "angle_form", "dihedral_form")):
            override = getattr(options, attr, None)
            if override is not None:
                names[slot] = override
        return [None, None] + [known[name].value for name in names]

    def __init__(self, filename, options):
        self._molecules = {}
        self._fconst_constr_threshold = options.constr_threshold
        self._temperature = getattr(options, "temperature", 310.0)
        self._functional_forms = self._get_default_func_forms(options)
        self.histogram_bins = getattr(options, "histogram_bins", HIST_BINS)
        self.histogram_ranges = dict(HIST_RANGES)
        self.staged = True            # stream bead tiles through shared memory (False: gather from global)
        self._plans = {}
        self._state = None

        def make_bond(atomlist):
            circular = len(atomlist) > 2
            form = self._functional_forms[len(atomlist)](
                circular_mean if circular else np.nanmean,
                circular_variance if circular else np.nanvar)
            return Bond(atoms=atomlist, func_form=form)

        with CFG(filename) as cfg:
            for mol_name, rows in cfg.items():
                mol_bonds = self._molecules[mol_name] = []
                for atomlist in rows:
                    if len(set(atomlist)) != len(atomlist):
                        raise ValueError("Defined bond '{0}' contains duplicate atoms".format(atomlist))
                    mol_bonds.append(make_bond(atomlist))
                if not any(len(bond.atoms) > 2 for bond in mol_bonds):
                    # no angle or dihedral listed: derive them from the bond graph (bondset.py:199-222)
                    angles, dihedrals = self._create_angles(mol_bonds)
                    if options.generate_angles:
                        mol_bonds.extend(make_bond(atomlist) for atomlist in angles)
                    if options.generate_dihedrals:
                        mol_bonds.extend(make_bond(atomlist) for atomlist in dihedrals)

    @staticmethod
    def _create_angles(mol_bonds):
        pairs = [bond.atoms for bond in mol_bonds if len(bond.atoms) == 2]
        angles = extend_graph_chain(pairs, pairs)
        return angles, extend_graph_chain(angles, pairs)

    # ------------------------------------------------------------------ selectors (bondset.py:236-328)
    def get_bonds(self, mol, natoms, select=lambda x: True):
        if natoms == -1:
            return [bond for bond in self._molecules[mol] if select(bond)]
        return [bond for bond in self._molecules[mol] if len(bond.atoms) == natoms and select(bond)]

    def get_bond_lengths(self, mol, with_constr=False):
        bonds = [bond for bond in self._molecules[mol] if len(bond.atoms) == 2]
        if with_constr:
            return bonds
        return [bond for bond in bonds if bond.fconst < self._fconst_constr_threshold]

    def get_bond_length_constraints(self, mol):
        return [bond for bond in self._molecules[mol]
                if len(bond.atoms) == 2 and bond.fconst >= self._fconst_constr_threshold]

    def get_bond_angles(self, mol, exclude_triangle=True):
        angles = [bond for bond in self._molecules[mol] if len(bond.atoms) == 3]
        if exclude_triangle:
            edges = {frozenset(bond.atoms) for bond in self.get_bond_lengths(mol, with_constr=True)}

            def closed(atoms):
                return all(frozenset((atoms[j - 1], atoms[j])) in edges for j in range(3))

            angles = [angle for angle in angles if not closed(angle.atoms)]
        return angles

    def get_bond_dihedrals(self, mol):
        return [bond for bond in self._molecules[mol] if len(bond.atoms) == 4]

    def get_virtual_beads(self, mapping):
        return [bead for bead in mapping if isinstance(bead, VirtualMap)]

    def _populate_atom_numbers(self, mapping):
        """Bead numbers of every bond from the mapping's bead order (``bondset.py:330-360``)."""
        for mol in self._molecules:
            if mol not in mapping:
                continue
            order = [bead.name for bead in mapping[mol]]
            for bond in self._molecules[mol]:
                try:
                    bond.atom_numbers = [order.index(atom.lstrip("+-")) for atom in bond.atoms]
                except ValueError as exc:
                    missing = [atom for atom in bond.atoms if atom.lstrip("+-") not in order]
                    exc.args = ("Bead(s) {0} do(es) not exist in residue {1}".format(missing, mol),)
                    raise

    # ------------------------------------------------------------------ text output (bondset.py:362-494)
    def write_itp(self, filename, mapping):
        self._populate_atom_numbers(mapping)
        file_write_lines(filename, self.itp_text(mapping))

    def itp_text(self, mapping):
        """Lines of a GROMACS .itp with beads and bonded parameters; same layout as the reference."""

        def section(bonds, header, print_fconst=True, multiplicity=None, rad2deg=False):
            lines = ["\n[ {0:s} ]".format(header)] if bonds else []
            for bond in bonds:
                eqm = math.degrees(bond.eqm) if rad2deg else bond.eqm
                line = " ".join("{0:4d}".format(num + 1) for num in bond.atom_numbers)
                line += " {0:4d} {1:12.5f}".format(bond.gromacs_type_id, eqm)
                if print_fconst:
                    line += " {0:12.5f}".format(bond.fconst)
                if multiplicity is not None:
                    line += " {0:4d}".format(multiplicity)
                lines.append(line)
            return lines

        out = [
            ";",
            "; Topology prepared automatically using PyCGTOOL",
            "; James Graham <J.A.Graham@soton.ac.uk> 2016",
            "; University of Southampton",
            "; https://github.com/jag1g13/pycgtool",
            ";",
        ]
        for mol in self._molecules:
            if mol not in mapping:
                logger.warning("Molecule '%s' present in bonding file, but not in mapping. "
                               "Parameters have not been calculated.", mol)
                continue
            if any(bond.fconst is None for bond in self._molecules[mol]):
                logger.warning("Molecule '%s' has no measured bond values. "
                               "Parameters have not been calculated.", mol)
                continue
            out.append("\n[ moleculetype ]")
            out.append("{0:4s} {1:4d}".format(mol, 1))
            out.append("\n[ atoms ]")
            for i, bead in enumerate(mapping[mol], start=1):
                line = "{0:4d} {1:4s} {2:4d} {3:4s} {4:4s} {5:4d} {6:8.3f}".format(
                    i, bead.type, 1, mol, bead.name, i, bead.charge)
                if isinstance(bead, VirtualMap):
                    line += " {0:8.3f}".format(bead.mass)
                out.append(line)
            virtual_beads = self.get_virtual_beads(mapping[mol])
            if virtual_beads:
                out.append("\n[ virtual_sitesn ]")
                exclusions = ["\n[ exclusions ]"]
                for vbead in virtual_beads:
                    ids = sorted(bead.num + 1 for bead in mapping[mol] if bead.name in vbead.atoms)
                    ids_text = " ".join(map(str, ids))
                    out.append("{0:^6d} {1:^6d} {2}".format(vbead.num + 1, vbead.gromacs_type_id, ids_text))
                    exclusions.append("{} ".format(vbead.num + 1) + ids_text)
                out.extend(exclusions)
            out.extend(section(self.get_bond_lengths(mol), "bonds"))
            out.extend(section(self.get_bond_angles(mol), "angles", rad2deg=True))
            out.extend(section(self.get_bond_dihedrals(mol), "dihedrals", multiplicity=1, rad2deg=True))
            out.extend(section(self.get_bond_length_constraints(mol), "constraints", print_fconst=False))
        return out

    # ------------------------------------------------------------------ plan (bondset.py:507-528)
    def compile_plan(self, topology):
        """Resolve every bond's bead names to bead indices for all instances.

        One instance per residue whose name equals the molecule; a name prefixed
        ``-`` / ``+`` is looked up in the previous / next residue whatever its name;
        instances whose neighbour does not exist (chain ends) are skipped.  A name
        missing from its residue raises ``KeyError`` like ``Residue.atom`` does.
        Columns are ordered by kind, then by the position of their first bead, so
        adjacent GPU lanes gather adjacent beads.
        """
        top = topology
        plan = MeasurePlan()
        plan.n_beads = top.n_atoms
        bonds = list(itertools.chain(*self._molecules.values()))
        plan.n_bonds = len(bonds)
        plan.natoms = np.array([len(b.atoms) for b in bonds], dtype=np.int32)
        n_res = top.n_residues
        res_name = {nm: i for i, nm in enumerate(top.res_names)}
        bond_id = 0
        rows_by_kind = {2: [], 3: [], 4: []}
        for mol_name, mol_bonds in self._molecules.items():
            home = np.nonzero(top.res_name_id == res_name[mol_name])[0] if mol_name in res_name else \
                np.zeros(0, dtype=np.int64)
            for bond in mol_bonds:
                k = len(bond.atoms)
                ok = np.ones(len(home), dtype=bool)
                cols = []
                for name in bond.atoms:
                    shift = {"-": -1, "+": 1}.get(name[0], 0)
                    target = home + shift
                    ok &= (target >= 0) & (target < n_res)
                    cols.append((name.lstrip("-+"), np.clip(target, 0, max(n_res - 1, 0))))
                table = np.zeros((int(ok.sum()), k), dtype=np.int64)
                for j, (bare, target) in enumerate(cols):
                    found = top.atom_index_by_name(bare, target[ok])
                    if np.any(found < 0):
                        raise KeyError(bare)
                    table[:, j] = found
                plan.bond_indices.append(table.astype(np.int32))
                rows_by_kind[k].append((bond_id, table))
                bond_id += 1
        col_beads, col_bond = [], []
        plan.bond_cols = [None] * plan.n_bonds
        offset = 0
        counts = []
        for k in (2, 3, 4):
            items = rows_by_kind[k]
            n_k = sum(len(t) for _, t in items)
            counts.append(n_k)
            if n_k == 0:
                for b, _ in items:
                    plan.bond_cols[b] = np.zeros(0, dtype=np.int64)
                continue
            tables = np.concatenate([t for _, t in items], axis=0)
            owners = np.concatenate([np.full(len(t), b, dtype=np.int32) for b, t in items])
            order = np.argsort(tables.min(axis=1), kind="stable")
            position = np.empty(n_k, dtype=np.int64)
            position[order] = np.arange(n_k)
            padded = np.concatenate([tables, np.repeat(tables[:, -1:], 4 - k, axis=1)], axis=1)
            col_beads.append(padded[order])
            col_bond.append(owners[order])
            start = 0
            for b, t in items:
                plan.bond_cols[b] = offset + position[start:start + len(t)]
                start += len(t)
            offset += n_k
        plan.counts = tuple(counts)
        plan.col_beads = (np.concatenate(col_beads, axis=0) if col_beads else np.zeros((0, 4))).astype(np.int32)
        plan.col_bond = (np.concatenate(col_bond) if col_bond else np.zeros(0)).astype(np.int32)
        plan.col_slot, plan.max_slots = self._tile_slots(plan.col_bond, plan.counts)
        plan.tile_span, plan.max_span = self._tile_spans(plan.col_beads, plan.counts)
        return plan

    @staticmethod
    def _tile_spans(col_beads, counts):
        """Contiguous bead range covering each tile of ``CGB_MEASURE_TILE`` columns (staged kernels)."""
        tile = _lib.MEASURE_TILE
        rows, start = [], 0
        for n_k in counts:
            if n_k:
                block = col_beads[start:start + n_k]
                lo, hi = block.min(axis=1), block.max(axis=1)
                cuts = np.arange(0, n_k, tile)
                first = np.minimum.reduceat(lo, cuts)
                last = np.maximum.reduceat(hi, cuts)
                rows.append(np.stack([first, last - first + 1], axis=1))
            start += n_k
        if not rows:
            return np.zeros((0, 2), dtype=np.int32), 0
        spans = np.concatenate(rows, axis=0).astype(np.int32)
        return spans, int(spans[:, 1].max())

    @staticmethod
    def _tile_slots(col_bond, counts):
        """Number the distinct Bonds inside every tile of ``CGB_MEASURE_TILE`` consecutive columns of a
        kind (see ``include/pycgtool_b200.h``): the measurement CTAs keep per-Bond sums and histograms
        in shared memory indexed by this compact slot."""
        tile = _lib.MEASURE_TILE
        slots = np.zeros(len(col_bond), dtype=np.int32)
        max_slots, start = 0, 0
        for n_k in counts:
            bonds = col_bond[start:start + n_k].astype(np.int64)
            if n_k:
                width = min(n_k, tile)
                tile_id = np.arange(n_k) // width
                # rank of each (tile, bond) pair among the pairs of its tile
                key = tile_id * (int(bonds.max()) + 1) + bonds
                uniq, inverse = np.unique(key, return_inverse=True)
                first_of_tile = np.searchsorted(uniq // (int(bonds.max()) + 1), np.arange(tile_id[-1] + 1))
                local = inverse - first_of_tile[tile_id]
                slots[start:start + n_k] = local
                max_slots = max(max_slots, int(local.max()) + 1)
            start += n_k
        return slots, max_slots

    def _plan_for(self, topology):
        key = id(topology)
        cached = self._plans.get(key)
        if cached is None or cached[0] is not topology or cached[1].n_beads != topology.n_atoms:
            cached = (topology, self.compile_plan(topology))
            self._plans = {key: cached}
        return cached[1]

    def _form_ids(self):
        ids = []
        for bond in itertools.chain(*self._molecules.values()):
            form = bond._func_form.device_form
            ids.append(_lib.FORM_STATS_ONLY if form is None else form)
        return np.array(ids, dtype=np.int32)

    @staticmethod
    def _device_plan(plan, dev):
        key = str(dev)
        if key not in plan._dev:
            plan._dev[key] = {
                "col_beads": device.to_device(plan.col_beads, dev),
                "col_bond": device.to_device(plan.col_bond, dev),
                "col_slot": device.to_device(plan.col_slot, dev),
                "tile_span": device.to_device(plan.tile_span, dev),
                "natoms": device.to_device(plan.natoms, dev),
            }
        return plan._dev[key]

    # ------------------------------------------------------------------ apply (bondset.py:496-531)
    def apply(self, frame, histogram=True):
        """Measure every bonded term in every frame of ``frame`` on the GPU.

        Afterwards each ``bond.values`` is the frame-major float32 array the
        reference stores (fetched lazily), and the statistics accumulators needed by
        ``boltzmann_invert`` are already on the device.
        """
        import os
        import time
        import torch
        device.require_cuda()
        lib = _lib.load()
        trace = os.environ.get("CGB_TRACE")
        marks = [("start", time.perf_counter())]

        def mark(label):
            if trace:
                torch.cuda.synchronize()
                marks.append((label, time.perf_counter()))

        self._state = None          # release the previous call's buffers before allocating new ones
        plan = self._plan_for(frame._topology)
        mark("plan")
        traj = frame._trajectory
        xyz = traj.device_xyz()
        dev = xyz.device
        n_frames = int(xyz.shape[0])
        n2, n3, n4 = plan.counts
        n_cols = n2 + n3 + n4
        d = self._device_plan(plan, dev)
        if "forms" not in d:
            d["forms"] = device.to_device(self._form_ids(), dev)

        # mdtraj: orthorhombic fast path only if every angle of every frame is ~90 degrees
        box_dev, ortho = traj.device_box_vectors(dev), int(traj.orthorhombic)
        mark("box")

        values = torch.empty(n_frames * n_cols, dtype=torch.float32, device=dev)   # one plane per kind
        partials = torch.zeros((plan.n_bonds, _lib.NPART), dtype=torch.float64, device=dev)
        hist = hist_lo = hist_hi = lo = hi = None
        if histogram and plan.n_bonds:
            nbins = int(self.histogram_bins)
            hist = torch.zeros((plan.n_bonds, nbins), dtype=torch.int64, device=dev)
            key = ("hist", tuple(sorted(self.histogram_ranges.items())))
            if key not in d:
                lo = np.array([self.histogram_ranges[int(k)][0] for k in plan.natoms], dtype=np.float32)
                hi = np.array([self.histogram_ranges[int(k)][1] for k in plan.natoms], dtype=np.float32)
                d[key] = (lo, hi, device.to_device(lo, dev), device.to_device(hi, dev))
            lo, hi, hist_lo, hist_hi = d[key]
        stream = device.stream_handle()
        mark("alloc")

        def measure(nf, vals, parts, shift, hst):
            _lib.check(lib.cgb_measure(device.dptr(xyz), device.dptr(box_dev), ortho, nf, plan.n_beads,
                                       device.dptr(d["col_beads"]), device.dptr(d["col_bond"]),
                                       device.dptr(d["col_slot"]), plan.max_slots,
                                       device.dptr(d["tile_span"]) if self.staged else None, plan.max_span,
                                       n2, n3, n4,
                                       plan.n_bonds, device.dptr(shift), device.dptr(vals), device.dptr(parts),
                                       device.dptr(hst), device.dptr(hist_lo), device.dptr(hist_hi),
                                       int(self.histogram_bins), stream), "cgb_measure")

        # Shift for the one-pass variance: each bond's first value in the first frame of rank 0, so
        # that identical samples give exactly zero variance and all shards share the same shift.
        shift = torch.zeros(plan.n_bonds, dtype=torch.float32, device=dev)
        if n_cols and n_frames:
            scratch = torch.zeros_like(partials)
            measure(1, values, scratch, None, None)
            if "first_col" not in d:
                first_col = np.array([cols[0] if len(cols) else 0 for cols in plan.bond_cols], dtype=np.int64)
                d["first_col"] = device.to_device(first_col, dev)
                d["has_col"] = device.to_device(np.array([len(cols) > 0 for cols in plan.bond_cols]), dev)
            has = d["has_col"]
            # the one-frame pre-pass wrote planes of a 1-frame trajectory: plane offset 1*c0 == column index
            picked = values[:n_cols].index_select(0, d["first_col"])
            shift = torch.where(has & torch.isfinite(picked), picked, torch.zeros_like(picked)).contiguous()
        dist.broadcast(shift, src=0)
        mark("shift")
        if n_cols and n_frames:
            measure(n_frames, values, partials, shift, hist)
        mark("measure")

        self._state = _MeasureState(self, plan, n_frames, values, partials, hist, shift,
                                    None if hist is None else lo, None if hist is None else hi, dev)
        for bond_id, bond in enumerate(itertools.chain(*self._molecules.values())):
            bond._values = None
            bond._source = (self._state, bond_id)
            bond.eqm = bond.fconst = None
        if trace:
            print("BondSet.apply trace:", ", ".join(
                f"{label} {1e3 * (t - marks[i][1]):.2f} ms" for i, (label, t) in enumerate(marks[1:])))

    def boltzmann_invert(self):
        """Equilibrium values and force constants of all bonds (``bondset.py:533-539``)."""
        bonds = list(itertools.chain(*self._molecules.values()))
        state = self._state
        if state is not None and all(b._source is not None and b._source[0] is state for b in bonds):
            state.invert(self._temperature)
            for bond_id, bond in enumerate(bonds):
                if state.n_values(bond_id):
                    state.assign(bond, bond_id, self._temperature)
            return
        for bond in bonds:
            try:
                bond.boltzmann_invert(temp=self._temperature)
            except ValueError:
                pass

    # ------------------------------------------------------------------ sample dump (bondset.py:541-599)
    @staticmethod
    def _get_lines_for_bond_dump(bonds, target_number=None, rad2deg=False):
        lines = []
        for row in transpose_and_sample((bond.values for bond in bonds), n=target_number):
            if rad2deg:
                row = [math.degrees(val) for val in row]
            lines.append((len(row) * "{:12.5f}").format(*row))
        return lines

    def dump_values(self, target_number=10000, output_dir=None):
        import pathlib
        output_dir = pathlib.Path("." if output_dir is None else output_dir)
        for mol in self._molecules:
            if mol == "SOL":
                continue
            for bonds, suffix, deg in ((self.get_bond_lengths(mol, with_constr=True), "length", False),
                                       (self.get_bond_angles(mol), "angle", True),
                                       (self.get_bond_dihedrals(mol), "dihedral", True)):
                if bonds:
                    lines = BondSet._get_lines_for_bond_dump(bonds, target_number, rad2deg=deg)
                    file_write_lines(output_dir.joinpath(f"{mol}_{suffix}.dat"), lines)

    # ------------------------------------------------------------------ dict-like surface
    def __len__(self):
        return len(self._molecules)

    def __contains__(self, item):
        return item in self._molecules

    def __getitem__(self, item):
        return self._molecules[item]

    def __iter__(self):
        return iter(self._molecules)
