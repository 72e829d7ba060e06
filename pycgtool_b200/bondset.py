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
        self._histogram = None   # (counts, lo, hi) after boltzmann_invert; the edges are made on demand

    def __iter__(self):
        return iter(self.atoms)

    @property
    def histogram(self):
        """``(counts uint64 (nbins,), edges float64 (nbins + 1,))`` of the measured values after
        ``boltzmann_invert`` (``numpy.histogram`` semantics), or None."""
        if self._histogram is None:
            return None
        counts, lo, hi = self._histogram
        return counts, np.linspace(lo, hi, len(counts) + 1)

    @histogram.setter
    def histogram(self, value):
        if value is None:
            self._histogram = None
        else:
            counts, edges = value
            self._histogram = (counts, float(edges[0]), float(edges[-1]))

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
            state.invert(temp)      # collective when frames are sharded: every rank enters before any early exit
            if state.n_values_global(bond_id) == 0:
                raise ValueError("No bonds were measured between beads {0}".format(self.atoms))
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
                                          device.dptr(row), None, 0, stream), "cgb_circular_pass2")
    out = torch.empty(_lib.NOUT, dtype=torch.float64, device=dev)
    nat = torch.tensor([natoms], dtype=torch.int32, device=dev)
    form = torch.tensor([_lib.FORM_STATS_ONLY], dtype=torch.int32, device=dev)
    sh = torch.tensor([shift], dtype=torch.float32, device=dev)
    _lib.check(lib.cgb_boltzmann(device.dptr(row), device.dptr(sh), device.dptr(nat), device.dptr(form),
                                 310.0, 1, None, None, None, 0, 2, device.dptr(out), stream), "cgb_boltzmann")
    res = out.cpu().numpy()
    return float(res[2]), float(res[3])


class MeasurePlan:
    """Integer tables for one (bond set, topology) pair: the columns and the fused kernel's plan."""

    def __init__(self):
        self.n_beads = 0
        self.n_bonds = 0
        self.counts = (0, 0, 0)       # columns per kind (lengths, angles, dihedrals)
        self.col_beads = None         # int32 (M, 4)
        self.col_bond = None          # int32 (M,)
        self.bond_cols = []           # per bond: int64 positions in the values row, in instance order
        self.bond_columns = []        # per bond: int64 column numbers (rows of col_beads), in instance order
        self.bond_indices = []        # per bond: int32 (n_inst, natoms) -- what mdtraj would receive
        self.natoms = None            # int32 (n_bonds,)
        self.bond_first = None        # int32 (n_bonds, 4): beads of each bond's first instance
        self.natoms_measured = None   # int32 (n_bonds,): natoms, 0 for a bond without instances
        # fused kernel (cgb_measure_fused): structured arrays mirroring include/pycgtool_b200.h
        self.tiles = self.groups = self.lanes = self.tile_bonds = self.tile_lists = None
        self.shape = _lib.SHAPE_EAD   # kinds of the four chunks of every group (compile-time in the kernel)
        self.n_staged = 0             # columns [0, n_staged) of the values row come from the fused kernel
        self.max_span = self.max_slots = self.max_edge_chunks = self.max_tile_groups = 0
        self.col_out = None           # int32 (M,): position of every column in the values row
        # gather kernels (cgb_measure): columns whose beads are too far apart for a tile, in kind order
        self.wide_counts = (0, 0, 0)
        self.wide_beads = self.wide_bond = self.wide_slot = None
        self.wide_max_slots = 0
        # frame folding (small systems): a second plan for `fold` copies of the system, measured as one frame
        self.fold = 1
        self.folded = None            # MeasurePlan of the folded system (kernel arrays only)
        self.fold_bond_cols = None    # per bond: int64 (fold, n_inst) positions in the folded values row
        self._dev = {}

    @property
    def n_cols(self):
        return int(sum(self.counts))


class _MeasureState:
    """Device buffers of one ``BondSet.apply`` call.

    ``packed`` is ONE float64 buffer ``[partials (n_bonds, 8) | hist (n_bonds, nbins)]``: the only state
    that crosses GPUs, summed with a single all-reduce.
    """

    def __init__(self, bondset, plan, n_frames, values, packed, nbins, shift, hist_lo, hist_hi, dev, source):
        self.bondset, self.plan, self.n_frames = bondset, plan, n_frames
        self.values, self.packed, self.shift = values, packed, shift
        npart = plan.n_bonds * _lib.NPART
        self.partials = packed[:npart].view(plan.n_bonds, _lib.NPART)
        self.hist = packed[npart:].view(plan.n_bonds, nbins) if nbins else None
        self.hist_lo, self.hist_hi = hist_lo, hist_hi
        self.dev = dev
        self.source = source        # (xyz, box, ortho) on the device: lets `.values` be measured on demand
        self.reduced = False
        self.result = None
        self.result_temp = None
        self.hist_host = None
        self.counts_host = None     # (n_bonds,) number of values over all ranks (finite or not)

    def n_values(self, bond_id):
        return self.n_frames * len(self.plan.bond_cols[bond_id])

    def n_values_global(self, bond_id):
        """Values of the bond over all ranks (known after ``invert``)."""
        if self.counts_host is None:
            return self.n_values(bond_id)
        return int(self.counts_host[bond_id])

    def fetch_values(self, bond_id):
        import torch
        cols = self.plan.bond_cols[bond_id]
        if len(cols) == 0 or self.n_frames == 0:
            return np.zeros(0, dtype=np.float32)
        if self.values is None:
            # statistics-only pass: measure the values now (into scratch accumulators)
            self.values = torch.empty((self.n_frames, self.plan.n_cols), dtype=torch.float32, device=self.dev)
            scratch = torch.zeros_like(self.packed)
            self.bondset._launch_measure(self.plan, self.source, self.n_frames, self.shift, self.values, scratch,
                                         0, None, None)
        main = BondSet.folded_frames(self.plan, self.n_frames)
        parts = []
        if main:
            # the folded rows hold `fold` real frames each, in the folded plan's column order: copy r of the bond's
            # instances first for r = 0, then r = 1, ... = frame-major once flattened
            fold = self.plan.fold
            fidx = torch.from_numpy(np.ascontiguousarray(self.plan.fold_bond_cols[bond_id].reshape(-1))).to(self.dev)
            view = self.values[:main].reshape(main // fold, fold * self.plan.n_cols)
            parts.append(view.index_select(1, fidx).reshape(-1))
        if main < self.n_frames:
            idx = torch.from_numpy(cols).to(self.dev)
            parts.append(self.values[main:].index_select(1, idx).reshape(-1))
        return torch.cat(parts).cpu().numpy()

    def _boltzmann(self, temp, pass2, out):
        lib = _lib.load()
        plan = self.plan
        d = BondSet._device_plan(plan, self.dev)
        nbins = 0 if self.hist is None else int(self.hist.shape[1])
        _lib.check(lib.cgb_boltzmann(device.dptr(self.partials), device.dptr(self.shift), device.dptr(d["natoms"]),
                                     device.dptr(d["forms"]), float(temp), plan.n_bonds, device.dptr(self.hist),
                                     device.dptr(self.hist_lo), device.dptr(self.hist_hi), nbins, pass2,
                                     device.dptr(out), device.stream_handle()), "cgb_boltzmann")

    def invert(self, temp):
        """Cross-GPU reduction and the epilogue kernel (idempotent per ``temp``).

        Common case: one all-reduce of the packed buffer, one epilogue launch, one device-to-host copy.  Only
        when a dihedral Bond has samples next to the antipode of its mean (its one-pass circular variance is
        then flagged inexact) does a second pass run: over the stored values, or -- statistics-only -- a second
        measurement about the now known mean.
        """
        import torch
        if self.result is not None and self.result_temp == temp:
            return
        lib = _lib.load()
        plan = self.plan
        stream = device.stream_handle()
        d = BondSet._device_plan(plan, self.dev)
        if not self.reduced:
            dist.allreduce_packed(self.packed)      # frames are sharded across GPUs: one sum of [moments | hist]
            self.reduced = True
        out = torch.empty((plan.n_bonds, _lib.NOUT), dtype=torch.float64, device=self.dev)
        self._boltzmann(temp, 0, out)
        # one device-to-host copy: fitted parameters, value counts and (first time) the histograms
        want_hist = self.hist is not None and self.hist_host is None
        parts = [out.reshape(-1), self.partials[:, 6]] + ([self.hist.reshape(-1)] if want_hist else [])
        host = torch.cat(parts).cpu().numpy()
        n_out = plan.n_bonds * _lib.NOUT
        result = host[:n_out].reshape(plan.n_bonds, _lib.NOUT)
        self.counts_host = host[n_out:n_out + plan.n_bonds].astype(np.int64)
        if want_hist:
            self.hist_host = np.rint(host[n_out + plan.n_bonds:]).astype(np.uint64).reshape(self.hist.shape)
        flagged = np.nonzero(result[:, 4] == 1.0)[0]
        if len(flagged):
            n_cols = plan.n_cols
            if self.values is not None:
                # exact <wrap(v - mean)^2> of the flagged Bonds over the stored values (util.py:49-53)
                self.partials[:, 5].zero_()
                if n_cols and self.n_frames:
                    main = BondSet.folded_frames(plan, self.n_frames)
                    regions = []
                    if main:
                        fd = BondSet._device_plan(plan.folded, self.dev)
                        regions.append((self.values, main // plan.fold, plan.fold * n_cols, fd["col_bond_by_pos"]))
                    if main < self.n_frames:
                        regions.append((self.values[main:], self.n_frames - main, n_cols, d["col_bond_by_pos"]))
                    for vals, frames, width, by_pos in regions:
                        _lib.check(lib.cgb_circular_pass2(device.dptr(vals), frames, width, device.dptr(by_pos), width,
                                                          device.dptr(self.partials), out.data_ptr() + 4 * 8,
                                                          _lib.NOUT, stream), "cgb_circular_pass2")
                if dist.is_distributed():
                    wrapped = self.partials[:, 5].contiguous()
                    dist.allreduce_packed(wrapped)
                    self.partials[:, 5] = wrapped
                self._boltzmann(temp, 1, out)
            else:
                # statistics only: measure again about the mean (then e == 0 and the one-pass sum is exact)
                mean32 = out[:, 2].to(torch.float32)
                pick = torch.zeros(plan.n_bonds, dtype=torch.bool, device=self.dev)
                pick[torch.from_numpy(flagged).to(self.dev)] = True
                self.shift = torch.where(pick, mean32, self.shift).contiguous()
                again = torch.zeros_like(self.packed)
                nbins = 0 if self.hist is None else int(self.hist.shape[1])
                self.bondset._launch_measure(plan, self.source, self.n_frames, self.shift, None, again, nbins,
                                             self.hist_lo, self.hist_hi)
                dist.allreduce_packed(again)
                npart = plan.n_bonds * _lib.NPART
                fresh = again[:npart].view(plan.n_bonds, _lib.NPART)
                self.partials[pick] = fresh[pick]
                self._boltzmann(temp, 0, out)
            result = out.cpu().numpy()
        self.result = result
        self.result_temp = temp
        if self.hist is not None and self.hist_host is None:
            self.hist_host = np.rint(self.hist.cpu().numpy()).astype(np.uint64)

    def assign(self, bond, bond_id, temp):
        eqm, fconst, mean, var = (float(x) for x in self.result[bond_id][:4])
        if math.isnan(mean):
            return          # nothing finite was measured: leave eqm / fconst as None
        _finish_bond(bond, mean, var, temp, device_result=(eqm, fconst))
        if self.hist_host is not None:
            bond._histogram = (self.hist_host[bond_id], float(self.hist_lo_host[bond_id]),
                               float(self.hist_hi_host[bond_id]))


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
        self.staged = True            # fused kernel over shared-memory tiles (False: gather from global)
        self.fold_frames = True       # small systems: measure several frames as one (frame folding)
        self.tuning = None            # optional _lib.MeasureTuning, passed to every cgb_measure_fused call
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
    def compile_plan(self, topology, fused=None):
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
        plan.col_beads = np.ascontiguousarray(
            np.concatenate(col_beads, axis=0) if col_beads else np.zeros((0, 4)), dtype=np.int32)
        plan.col_bond = np.ascontiguousarray(np.concatenate(col_bond) if col_bond else np.zeros(0), dtype=np.int32)
        plan.bond_first = np.zeros((plan.n_bonds, 4), dtype=np.int32)
        plan.natoms_measured = np.zeros(plan.n_bonds, dtype=np.int32)
        for b, cols in enumerate(plan.bond_cols):
            if len(cols):
                plan.bond_first[b] = plan.col_beads[cols[0]]
                plan.natoms_measured[b] = plan.natoms[b]
        if not self._compile_kernel_plan(plan, fused):
            return None           # fused=(groups, beads): a unit of the system does not fit a fused tile
        # from column numbers to positions in the values row
        plan.bond_columns = plan.bond_cols
        plan.bond_cols = [plan.col_out[cols].astype(np.int64) for cols in plan.bond_columns]
        if fused is None and self.staged and self.fold_frames:
            self._compile_folded(plan)
        return plan

    # Small systems leave most of a warp idle (one small molecule: 23 of 32 lanes): `fold` consecutive frames are then
    # measured as ONE frame of fold * B beads -- the CG trajectory [F][B][3] is the same memory as
    # [F / fold][fold * B][3] -- with a plan compiled for `fold` copies of the system.
    FOLD_MAX_BEADS = 128          # only systems this small are folded
    FOLD_CANDIDATES = range(2, 25)

    def _compile_folded(self, plan):
        import ctypes
        lib = _lib.load()
        n_cols = plan.n_cols
        if not n_cols or plan.n_beads > self.FOLD_MAX_BEADS or sum(plan.wide_counts) or len(plan.tiles) != 1:
            return

        def cost(p, fold):
            # CTA time per real frame: a tile of ng groups spreads its frames over 7 // ng warps per group
            return sum(1.0 / max(1, _lib.MEAS_WARPS // int(ng)) for ng in p.tiles["ngroups"]) / fold
        best, best_cost = None, cost(plan, 1)
        n2, n3, n4 = plan.counts
        starts = (0, n2, n2 + n3)
        for fold in self.FOLD_CANDIDATES:
            if fold * plan.n_beads > 512 or fold * n_cols > 4096:
                break
            f = MeasurePlan()
            f.n_beads, f.n_bonds, f.natoms = fold * plan.n_beads, plan.n_bonds, plan.natoms
            f.counts = (fold * n2, fold * n3, fold * n4)
            beads, bond, origin = [], [], []           # origin: (copy, column) of every folded column
            for k0, nk in zip(starts, plan.counts):
                for r in range(fold):
                    beads.append(plan.col_beads[k0:k0 + nk] + r * plan.n_beads)
                    bond.append(plan.col_bond[k0:k0 + nk])
                    origin.append(np.stack([np.full(nk, r), np.arange(k0, k0 + nk)], axis=1))
            f.col_beads = np.ascontiguousarray(np.concatenate(beads), dtype=np.int32)
            f.col_bond = np.ascontiguousarray(np.concatenate(bond), dtype=np.int32)
            origin = np.concatenate(origin)
            f.bond_first, f.natoms_measured = plan.bond_first, plan.natoms_measured
            if not self._compile_kernel_plan(f) or sum(f.wide_counts) or not len(f.tiles):
                continue
            # the launch must fit a CTA's shared memory with the `fold` box-table entries per staged frame
            cfg = (ctypes.c_int64 * 8)()
            rc = lib.cgb_measure_fused_config(1 << 20, f.n_beads, len(f.tiles), f.max_span, f.max_slots,
                                              f.max_edge_chunks, f.max_tile_groups, int(self.histogram_bins), 1, None,
                                              148, cfg)
            if rc != 0 or cfg[1] + cfg[2] * cfg[3] * 48 * (fold - 1) > 110 * 1024:
                continue
            c = cost(f, fold)
            if c < 0.95 * best_cost:          # a larger fold (longer tail, more tiles) must pay by at least 5 %
                best, best_cost = (fold, f, origin), c
        if best is None:
            return
        fold, f, origin = best
        plan.fold, plan.folded = fold, f
        # position in the folded values row of (copy r, column c)
        pos = np.empty((fold, n_cols), dtype=np.int64)
        pos[origin[:, 0], origin[:, 1]] = f.col_out
        plan.fold_bond_cols = [pos[:, cols] for cols in plan.bond_columns]

    def _compile_kernel_plan(self, plan, fused=None):
        """Tiles / groups / lanes of the fused kernel (``cgb_measure_plan_*``, host code in the library) and the
        sub-table of the columns left to the gather kernels.  ``fused=(groups, beads)``: the unit-aligned plan of
        ``cgb_map_measure`` (tiles of whole molecule instances, at most ``beads`` beads and ``groups`` groups);
        returns False when the system cannot be tiled that way."""
        import ctypes
        lib = _lib.load()
        n2, n3, n4 = plan.counts
        n_cols = n2 + n3 + n4
        nbins = int(self.histogram_bins)
        # Bonds per tile: what the largest molecule needs (a tile then never mixes molecule types, which keeps
        # its shared-memory histograms small), within what fits
        per_molecule = max([len(bonds) for bonds in self._molecules.values()] + [8])
        max_span, max_slots = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(lib.cgb_measure_plan_limits(nbins, 1, per_molecule, ctypes.byref(max_span), ctypes.byref(max_slots)),
                   "cgb_measure_plan_limits")
        plan.col_out = np.arange(n_cols, dtype=np.int32)
        wide = np.arange(n_cols, dtype=np.int64)
        plan.tiles = np.zeros(0, dtype=_lib.MEAS_TILE_DTYPE)
        plan.groups = np.zeros(0, dtype=_lib.MEAS_GROUP_DTYPE)
        plan.lanes = np.zeros(0, dtype=_lib.MEAS_LANE_DTYPE)
        plan.tile_bonds = np.zeros(0, dtype=np.int32)
        plan.tile_lists = np.zeros(0, dtype=np.uint16)
        plan.shape = _lib.SHAPE_EAD if n4 else _lib.SHAPE_EEAA
        plan.tile_start = None
        if (self.staged or fused) and n_cols:
            info = _lib.MeasPlanInfo()
            slots = max_slots.value
            unit_groups, span = (0, max_span.value) if fused is None else (int(fused[0]), min(int(fused[1]), max_span.value))
            args = (_lib.ptr(plan.col_beads), _lib.ptr(plan.col_bond), n2, n3, n4, plan.n_beads,
                    span, slots, plan.shape, unit_groups)
            rc = lib.cgb_measure_plan_size(*args, ctypes.byref(info))
            if fused is not None and rc == -3:
                return False
            _lib.check(rc, "cgb_measure_plan_size")
            if fused is not None:
                plan.tile_start = np.zeros(plan.n_beads, dtype=np.uint8)
            plan.tiles = np.zeros(info.n_tiles, dtype=_lib.MEAS_TILE_DTYPE)
            plan.groups = np.zeros(info.n_groups, dtype=_lib.MEAS_GROUP_DTYPE)
            plan.lanes = np.zeros(info.n_groups * _lib.MEAS_CHUNKS * 32, dtype=_lib.MEAS_LANE_DTYPE)
            plan.tile_bonds = np.zeros(info.n_tile_bonds, dtype=np.int32)
            plan.tile_lists = np.zeros(info.n_tile_lists, dtype=np.uint16)
            wide = np.zeros(info.n_wide, dtype=np.int64)
            _lib.check(lib.cgb_measure_plan_build(*args, _lib.ptr(plan.tiles), _lib.ptr(plan.groups),
                                                  _lib.ptr(plan.lanes), _lib.ptr(plan.tile_bonds),
                                                  _lib.ptr(plan.tile_lists), _lib.ptr(plan.col_out), _lib.ptr(wide),
                                                  _lib.ptr(plan.tile_start)), "cgb_measure_plan_build")
            plan.n_staged = int(info.n_staged)
            plan.max_span, plan.max_slots = int(info.max_span), int(info.max_slots)
            plan.max_edge_chunks, plan.max_tile_groups = int(info.max_edge_chunks), int(info.max_tile_groups)
        # wide columns keep their kind order; they follow the staged columns in the values row
        w_counts = (int((wide < n2).sum()), int(((wide >= n2) & (wide < n2 + n3)).sum()), int((wide >= n2 + n3).sum()))
        plan.wide_counts = w_counts
        plan.wide_beads = np.ascontiguousarray(plan.col_beads[wide])
        plan.wide_bond = np.ascontiguousarray(plan.col_bond[wide])
        plan.wide_slot, plan.wide_max_slots = self._tile_slots(plan.wide_bond, w_counts)
        return True

    @staticmethod
    def _tile_slots(col_bond, counts):
        """Number the distinct Bonds inside every tile of ``CGB_MEASURE_TILE`` consecutive columns of a
        kind (see ``include/pycgtool_b200.h``): the gather CTAs keep per-Bond sums and histograms
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
        key = (id(topology), bool(self.staged), int(self.histogram_bins))
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
            by_pos = np.zeros(plan.n_cols, dtype=np.int32)
            by_pos[plan.col_out] = plan.col_bond
            plan._dev[key] = {
                "tiles": device.to_device(plan.tiles.view(np.int32), dev),
                "groups": device.to_device(plan.groups.view(np.int32), dev),
                "lanes": device.to_device(plan.lanes.view(np.int32), dev),
                "tile_bonds": device.to_device(plan.tile_bonds, dev),
                "tile_lists": device.to_device(plan.tile_lists.view(np.int16), dev),
                "wide_beads": device.to_device(plan.wide_beads, dev),
                "wide_bond": device.to_device(plan.wide_bond, dev),
                "wide_slot": device.to_device(plan.wide_slot, dev),
                "col_bond_by_pos": device.to_device(by_pos, dev),
                "bond_first": device.to_device(plan.bond_first, dev),
                "natoms_measured": device.to_device(plan.natoms_measured, dev),
                "natoms": device.to_device(plan.natoms, dev),
            }
        return plan._dev[key]

    @staticmethod
    def folded_frames(plan, n_frames):
        """Leading frames measured through the folded plan (a multiple of ``plan.fold``); the rest, and everything
        when the trajectory is short, goes through the plain plan."""
        if plan.folded is None or n_frames < 8 * plan.fold:
            return 0
        return (n_frames // plan.fold) * plan.fold

    def _launch_measure(self, plan, source, n_frames, shift, values, packed, nbins, hist_lo, hist_hi):
        """Enqueue the measurement of all columns: the fused kernel, then the gather kernels for wide columns.
        ``values`` None = statistics only; ``nbins`` 0 = no histogram.  Small systems: the leading
        ``folded_frames`` go through the folded plan (``fold`` real frames per measured frame), the tail through
        the plain one; both accumulate into ``packed``."""
        import torch
        xyz, box_dev, ortho = source
        main = self.folded_frames(plan, n_frames)

        def tail():
            self._launch_one(plan, xyz[main:], None if box_dev is None else box_dev[main:], ortho, n_frames - main, 1,
                             shift, None if values is None else values[main:], plan.n_cols, packed, nbins,
                             hist_lo, hist_hi)

        if not main:
            tail()
            return
        if main < n_frames:
            # the few frames that do not fill a fold run beside the folded launch, on a side stream (both only add
            # to `packed` with atomics and write disjoint rows of `values`)
            current = torch.cuda.current_stream(xyz.device)
            side = self._side_stream(xyz.device)
            side.wait_stream(current)
            with torch.cuda.stream(side):
                tail()
        self._launch_one(plan.folded, xyz, box_dev, ortho, main // plan.fold, plan.fold, shift, values,
                         plan.fold * plan.n_cols, packed, nbins, hist_lo, hist_hi)
        if main < n_frames:
            current.wait_stream(side)
            for t in (xyz, box_dev, values, packed, shift, hist_lo, hist_hi):
                if t is not None:
                    t.record_stream(side)

    def _side_stream(self, dev):
        import torch
        streams = self.__dict__.setdefault("_side_streams", {})
        key = str(dev)
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=dev)
        return streams[key]

    def _launch_one(self, plan, xyz, box_dev, ortho, n_frames, fold, shift, values, row_stride, packed, nbins,
                    hist_lo, hist_hi):
        lib = _lib.load()
        dev = xyz.device
        d = self._device_plan(plan, dev)
        stream = device.stream_handle()
        npart = plan.n_bonds * _lib.NPART
        partials = packed[:npart]
        hist = packed[npart:] if nbins else None
        if len(plan.tiles):
            _lib.check(lib.cgb_measure_fused(
                device.dptr(xyz), device.dptr(box_dev), ortho, n_frames, plan.n_beads, device.dptr(d["tiles"]),
                len(plan.tiles), device.dptr(d["groups"]), device.dptr(d["lanes"]), device.dptr(d["tile_bonds"]),
                device.dptr(d["tile_lists"]), plan.shape, plan.max_span, plan.max_slots, plan.max_edge_chunks,
                plan.max_tile_groups, int(fold), device.dptr(shift),
                device.dptr(values), row_stride, device.dptr(partials), device.dptr(hist), device.dptr(hist_lo),
                device.dptr(hist_hi), int(nbins), _lib.tuning_ptr(self.tuning), stream), "cgb_measure_fused")
        w2, w3, w4 = plan.wide_counts
        if w2 + w3 + w4:
            wide_values = None if values is None else values.reshape(-1)[plan.n_staged:]
            _lib.check(lib.cgb_measure(
                device.dptr(xyz), device.dptr(box_dev), ortho, n_frames, plan.n_beads, device.dptr(d["wide_beads"]),
                device.dptr(d["wide_bond"]), device.dptr(d["wide_slot"]), plan.wide_max_slots, w2, w3, w4,
                device.dptr(shift), device.dptr(wide_values), row_stride, device.dptr(partials), device.dptr(hist),
                device.dptr(hist_lo), device.dptr(hist_hi), int(nbins), stream), "cgb_measure")

    def _release(self):
        """Drop the device buffers of the previous measurement (state, and the Bonds' references to it)."""
        self._state = None
        for bond in itertools.chain(*self._molecules.values()):
            if bond._source is not None:
                bond._source = None
                bond._values = []

    # ------------------------------------------------------------------ apply (bondset.py:496-531)
    def apply(self, frame, histogram=True, store_values=True):
        """Measure every bonded term in every frame of ``frame`` on the GPU.

        Afterwards each ``bond.values`` is the frame-major float32 array the reference stores (fetched
        lazily), and the statistics accumulators needed by ``boltzmann_invert`` are already on the device.
        ``store_values=False`` is the statistics-only mode: no value is written per measure (the reference
        only needs them for ``dump_values``, ``bondset.py:562-599``); ``bond.values`` then triggers a
        measurement of its own on first use.
        """
        import torch
        device.require_cuda()
        lib = _lib.load()
        self._release()             # the previous call's buffers go before new ones are allocated
        plan = self._plan_for(frame._topology)
        traj = frame._trajectory
        xyz = traj.device_xyz()
        dev = xyz.device
        n_frames = int(xyz.shape[0])
        n_cols = plan.n_cols
        d = self._device_plan(plan, dev)
        if "forms" not in d:
            d["forms"] = device.to_device(self._form_ids(), dev)

        # mdtraj: orthorhombic fast path only if every angle of every frame is ~90 degrees; with the frames
        # sharded across GPUs both decisions are taken over the whole trajectory, like the reference's
        box_dev = traj.device_box_vectors(dev)
        _, has_box, ortho = traj.global_box_flags(dev)
        ortho = int(ortho)
        if not has_box:
            box_dev = None
        source = (xyz, box_dev, ortho)

        nbins = int(self.histogram_bins) if (histogram and plan.n_bonds) else 0
        values = torch.empty((n_frames, n_cols), dtype=torch.float32, device=dev) if store_values else None
        packed = torch.zeros(plan.n_bonds * (_lib.NPART + nbins), dtype=torch.float64, device=dev)
        hist_lo = hist_hi = lo = hi = None
        if nbins:
            key = ("hist", tuple(sorted(self.histogram_ranges.items())))
            if key not in d:
                lo = np.array([self.histogram_ranges[int(k)][0] for k in plan.natoms], dtype=np.float32)
                hi = np.array([self.histogram_ranges[int(k)][1] for k in plan.natoms], dtype=np.float32)
                d[key] = (lo, hi, device.to_device(lo, dev), device.to_device(hi, dev))
            lo, hi, hist_lo, hist_hi = d[key]
        stream = device.stream_handle()

        # Shift for the one-pass variance: each bond's first value in the first frame of rank 0, so
        # that identical samples give exactly zero variance and all shards share the same shift.
        shift = torch.empty(plan.n_bonds, dtype=torch.float32, device=dev)
        _lib.check(lib.cgb_measure_shift(device.dptr(xyz), device.dptr(box_dev), ortho, 0, n_frames, plan.n_beads,
                                         device.dptr(d["bond_first"]), device.dptr(d["natoms_measured"]),
                                         plan.n_bonds, device.dptr(shift), stream), "cgb_measure_shift")
        dist.broadcast_device(shift, src=0)
        if n_cols and n_frames:
            self._launch_measure(plan, source, n_frames, shift, values, packed, nbins, hist_lo, hist_hi)

        state = _MeasureState(self, plan, n_frames, values, packed, nbins, shift, hist_lo, hist_hi, dev, source)
        state.hist_lo_host, state.hist_hi_host = lo, hi
        self._state = state
        for bond_id, bond in enumerate(itertools.chain(*self._molecules.values())):
            bond._values = None
            bond._source = (state, bond_id)
            bond.eqm = bond.fconst = None
            bond._histogram = None

    def boltzmann_invert(self):
        """Equilibrium values and force constants of all bonds (``bondset.py:533-539``)."""
        bonds = list(itertools.chain(*self._molecules.values()))
        state = self._state
        if state is not None and all(b._source is not None and b._source[0] is state for b in bonds):
            state.invert(self._temperature)
            for bond_id, bond in enumerate(bonds):
                if state.n_values_global(bond_id):
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
