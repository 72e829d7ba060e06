"""Atom -> bead mapping with the reference's ``Mapping`` API, executed on the GPU.

``Mapping(filename, options, itp_filename)`` parses a ``.map`` file exactly as
``pycgtool/mapping.py:135-357`` does (bead lists, optional charges, ``@v``
virtual beads, geom / mass / first weights, mdtraj-style renamed aliases).
``Mapping.apply(frame)`` differs in *how*: instead of a Python loop over
residues x beads with NumPy temporaries (``mapping.py:411-438``), the name
lookups are resolved once per topology into integer tables (the "assignment",
bit-exact with the reference's lookups), compiled into a tile plan, and the
frames are mapped by the ``cgb_map_tiles`` / ``cgb_map_gather`` kernels.
"""

import copy
import logging

import numpy as np

from . import _lib, device, dist, formats
from .elements import mass_by_symbol
from .frame import Frame
from .parsers import CFG, ITP
from .topology import Topology

logger = logging.getLogger(__name__)

#: widest input span (atoms) one shared-memory tile may stage, and most beads per tile.
#: 32 beads = 96 output floats = three full warps of (bead, component) lanes per frame group.
MAX_SPAN_ATOMS = 512
MAX_TILE_BEADS = 32
#: host-resident inputs at least this large are streamed through the GPU in frame chunks
STREAM_THRESHOLD_BYTES = 512 << 20


class BeadMap:
    """One bead: its name, type, charge, mass and the atoms it is built from."""

    def __init__(self, name, num, type=None, atoms=None, charge=0, mass=0):  # noqa: A002 (reference signature)
        self.name = name
        self.num = num
        self.type = type
        self.mass = mass
        self.charge = charge
        self.atoms = atoms
        count = len(atoms)
        # float32, as in the reference (mapping.py:55-58); "mass" is added later if masses are known
        self.weights_dict = {
            "geom": np.full(count, 1.0 / count, dtype=np.float32),
            "first": np.array([1.0] + [0.0] * (count - 1), dtype=np.float32),
        }
        self.weights = self.weights_dict["geom"]

    def __repr__(self):
        return f"BeadMap #{self.num} {self.name} type: {self.type} mass: {self.mass} charge: {self.charge}"

    def __iter__(self):
        return iter(self.atoms)

    def __len__(self):
        return len(self.atoms)

    def __getitem__(self, item):
        return self.atoms[item]

    def guess_atom_masses(self):
        """Masses from atom names: two-letter element symbol first, then one letter
        (``mapping.py:77-106``); float32 weights."""
        masses = np.zeros(len(self.atoms), dtype=np.float32)
        for i, atom in enumerate(self.atoms):
            for symbol in (atom[:2], atom[0]):
                try:
                    masses[i] = mass_by_symbol(symbol)
                    break
                except KeyError:
                    continue
            else:
                raise RuntimeError(f"Mass of atom {atom} could not be automatically assigned, "
                                   "map_center=mass is not available.")
        self.mass = sum(masses)
        if not np.all(masses):
            raise RuntimeError("Some atom masses could not be automatically assigned, "
                               "map_center=mass is not available")
        masses /= self.mass
        self.weights_dict["mass"] = masses


class VirtualMap(BeadMap):
    """A virtual site placed from other CG beads of the same residue."""

    def __init__(self, name, num, type=None, atoms=None, charge=0):  # noqa: A002
        super().__init__(name, num, type=type, atoms=atoms, charge=charge, mass=0.0)
        self.gromacs_type_id_dict = {"geom": 1, "mass": 2}
        self.gromacs_type_id = self.gromacs_type_id_dict["geom"]

    def guess_atom_masses(self):
        """Virtual beads have no atoms of their own."""
        return


def _unique_rows(block):
    """Distinct rows of an integer matrix and, per row, the index of its distinct row.

    ``np.unique(axis=0)`` sorts the rows lexicographically (9 s for the 20 M-atom configuration); hashing each
    row to 64 bits and sorting the hashes is ~30x faster.  Collisions are detected and fall back."""
    n, width = block.shape
    if n == 0 or width == 0:
        return block[:1], np.zeros(n, dtype=np.int64)
    mult = (np.arange(1, width + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        hashed = (block.astype(np.uint64) * mult[None, :]).sum(axis=1, dtype=np.uint64)
    _, first, inverse = np.unique(hashed, return_index=True, return_inverse=True)
    templates = block[first]
    inverse = inverse.reshape(-1)
    if not np.array_equal(templates[inverse], block):      # two different rows with one hash
        templates, inverse = np.unique(block, axis=0, return_inverse=True)
        inverse = inverse.reshape(-1)
    return templates, inverse


class MapPlan:
    """Integer tables for one (mapping, topology) pair; everything frame-invariant."""

    def __init__(self):
        self.n_atoms = 0
        self.n_beads = 0
        self.bead_ptr = None      # int32 (B+1,)
        self.index = None         # int32 (nnz,)  atoms for real beads, CG beads for virtual beads
        self.weights = None       # float32 or float64 (nnz,), real-bead entries
        self.vweights = None      # float32 or float64 (nnz,), virtual-bead entries
        self.is_virtual = None    # uint8 (B,)
        self.cg_topology = None
        # device-side copies, created lazily
        self._dev = {}


class Mapping:
    """Dictionary ``molecule name -> [BeadMap]`` plus the GPU-backed ``apply``."""

    def __init__(self, filename, options, itp_filename=None):
        self._manual_charges = {}
        self._mappings = {}
        self._map_center = options.map_center
        self._virtual_map_center = options.virtual_map_center
        self._masses_are_set = False
        self._plans = {}
        self.tuning = None      # optional _lib.MapTuning, passed to every K1 launch of this object

        with CFG(filename) as cfg:
            for mol_name, rows in cfg.items():
                self._mappings[mol_name], self._manual_charges[mol_name] = self._mol_map_from_section(rows)

        if itp_filename is not None:
            self._load_itp(itp_filename)
        if "mass" in (self._map_center, self._virtual_map_center) and not self._masses_are_set:
            self._guess_atom_masses()
        self._set_bead_weights()
        self._rename_atoms()

    # ------------------------------------------------------------------ parsing (mapping.py:171-357)
    @staticmethod
    def _mol_map_from_section(mol_section):
        mol_map, manual_charges = [], False
        for i, row in enumerate(mol_section):
            bead_class = BeadMap
            if row[0].startswith("@"):
                if row[0] != "@v":
                    raise ValueError(f'Invalid line prefix "{row[0]}" in mapping file')
                bead_class = VirtualMap
                row = row[1:]
            name, typ, third, *atoms = row
            charge = 0
            try:
                # an optional numeric third column is the bead charge (mapping.py:271-279)
                charge = float(third)
                manual_charges = True
            except ValueError:
                atoms.insert(0, third)
            if not atoms:
                raise ValueError(f"Bead {name} specification contains no atoms")
            mol_map.append(bead_class(name, i, type=typ, atoms=atoms, charge=charge))
        return mol_map, manual_charges

    def _load_itp(self, itp_filename):
        with ITP(itp_filename) as itp:
            for mol_name, mol_map in self._mappings.items():
                if mol_name not in itp:
                    logger.warning("No itp information for molecule %s found in %s", mol_name, itp.filepath)
                    continue
                manual = self._manual_charges[mol_name]
                if manual:
                    logger.warning("Charges assigned in mapping for molecule %s, ignoring itp charges", mol_name)
                try:
                    self._itp_section_into_mol_map(mol_map, itp[mol_name], manual)
                except KeyError:
                    # an atom of the mapping is missing from the .itp: same message, no abort (mapping.py:178-196)
                    logger.warning("No itp information for molecule %s found in %s", mol_name, itp.filepath)
            self._masses_are_set = True

    @staticmethod
    def _itp_section_into_mol_map(mol_map, itp_mol_entry, manual_charges):
        """ITP masses give *float64* weights (``mapping.py:224-227``)."""
        charge_mass = {toks[4]: (float(toks[6]), float(toks[7])) for toks in itp_mol_entry["atoms"]}
        for bead in mol_map:
            if isinstance(bead, VirtualMap):
                continue
            masses = np.array([charge_mass[atom][1] for atom in bead])
            bead.mass = sum(masses)
            bead.weights_dict["mass"] = masses / bead.mass
            if not manual_charges:
                for atom in bead:
                    bead.charge += charge_mass[atom][0]
        for bead in mol_map:
            if isinstance(bead, VirtualMap):
                parts = [real for real in mol_map if real.name in bead]
                masses = np.array([real.mass for real in parts])
                bead.weights_dict["mass"] = masses / sum(masses)
                if not manual_charges:
                    bead.charge = sum(real.charge for real in parts)

    def _guess_atom_masses(self):
        for mol_map in self._mappings.values():
            for bead in mol_map:
                if not bead.mass:
                    bead.guess_atom_masses()
            for bead in mol_map:
                if isinstance(bead, VirtualMap):
                    masses = np.array([real.mass for real in mol_map if real.name in bead], dtype=np.float32)
                    bead.weights_dict["mass"] = masses / sum(masses)
        self._masses_are_set = True

    def _set_bead_weights(self):
        for mol_map in self._mappings.values():
            for bmap in mol_map:
                if isinstance(bmap, VirtualMap):
                    bmap.weights = bmap.weights_dict[self._virtual_map_center]
                    bmap.gromacs_type_id = bmap.gromacs_type_id_dict[self._virtual_map_center]
                else:
                    bmap.weights = bmap.weights_dict[self._map_center]

    def _rename_atoms(self):
        """Add aliases under the names the topology reader produces (``mapping.py:288-317``)."""
        for mol_name in list(self._mappings):
            if mol_name not in formats.RESIDUE_NAME_REPLACEMENTS:
                continue
            alias = formats.RESIDUE_NAME_REPLACEMENTS[mol_name]
            table = formats.ATOM_NAME_REPLACEMENTS.get(alias, {})
            mol_map = copy.deepcopy(self._mappings[mol_name])
            for bead in mol_map:
                bead.atoms = [table.get(atom, atom) for atom in bead.atoms]
            self._mappings[alias] = mol_map

    # ------------------------------------------------------------------ dict-like surface
    def __len__(self):
        return len(self._mappings)

    def __contains__(self, item):
        return item in self._mappings

    def __getitem__(self, item):
        return self._mappings[item]

    def __iter__(self):
        return iter(self._mappings)

    def items(self):
        return self._mappings.items()

    # ------------------------------------------------------------------ plan (the "assignment")
    def compile_plan(self, topology):
        """Resolve every bead's atom names to indices for ``topology``.

        Follows the lookups of ``mapping.py:411-438``: residues whose *name* has a
        mapping, in topology order; beads in file order; each atom name is the
        first atom of that name inside the residue (``KeyError`` if absent);
        virtual beads index CG beads of the same CG residue.  Residues that share
        (name, atom-name sequence) are resolved once and broadcast.
        """
        top = topology
        plan = MapPlan()
        plan.n_atoms = top.n_atoms
        res_len = np.diff(top.res_first)
        name_of = top.res_names
        mapped_name_ids = [i for i, nm in enumerate(name_of) if nm in self._mappings]
        missing = sorted({nm for nm in name_of if nm not in self._mappings})
        for nm in missing:
            logger.warning("A mapping has not been provided for '%s' residues, they will not be mapped", nm)

        mapped_mask = np.isin(top.res_name_id, mapped_name_ids)
        mapped_res = np.nonzero(mapped_mask)[0]
        n_mres = len(mapped_res)
        beads_per_name = {i: len(self._mappings[name_of[i]]) for i in mapped_name_ids}
        nnz_per_name = {i: sum(len(b) for b in self._mappings[name_of[i]]) for i in mapped_name_ids}
        nb_r = np.zeros(n_mres, dtype=np.int64)
        nnz_r = np.zeros(n_mres, dtype=np.int64)
        rid = top.res_name_id[mapped_res]
        for i in mapped_name_ids:
            sel = rid == i
            nb_r[sel] = beads_per_name[i]
            nnz_r[sel] = nnz_per_name[i]
        bead0_r = np.concatenate([[0], np.cumsum(nb_r)])
        nnz0_r = np.concatenate([[0], np.cumsum(nnz_r)])
        n_beads, nnz = int(bead0_r[-1]), int(nnz0_r[-1])
        if n_beads >= 2**31 - 1 or nnz >= 2**31 - 1:
            raise ValueError("mapping too large for int32 tables")

        bead_ptr = np.zeros(n_beads + 1, dtype=np.int32)
        index = np.zeros(nnz, dtype=np.int32)
        is_virtual = np.zeros(n_beads, dtype=np.uint8)
        real_dtype, virt_dtype = np.float32, np.float32
        for i in mapped_name_ids:
            for bead in self._mappings[name_of[i]]:
                if bead.weights.dtype == np.float64:
                    if isinstance(bead, VirtualMap):
                        virt_dtype = np.float64
                    else:
                        real_dtype = np.float64
        weights = np.zeros(nnz, dtype=real_dtype)
        vweights = np.zeros(nnz, dtype=virt_dtype)
        bead_name_id = np.zeros(n_beads, dtype=np.int32)
        bead_names = []

        for i in mapped_name_ids:
            mol_map = self._mappings[name_of[i]]
            names_in_mol = [b.name for b in mol_map]
            sel_all = np.nonzero(rid == i)[0]           # positions within mapped_res
            lens = res_len[mapped_res[sel_all]]
            for natoms in np.unique(lens):
                sel = sel_all[lens == natoms]
                firsts = top.res_first[mapped_res[sel]]
                if natoms:
                    block = top.atom_name_id[firsts[:, None] + np.arange(natoms)[None, :]]
                else:
                    block = np.zeros((len(sel), 0), dtype=np.int32)
                templates, inverse = _unique_rows(block)
                for t, template in enumerate(templates):
                    atom_names = [top.atom_names[a] for a in template]
                    members = sel[inverse == t]
                    # local CSR for this template
                    l_off, l_w, l_vw, l_ptr, l_virt, l_isv_entry = [], [], [], [0], [], []
                    for bead in mol_map:
                        virt = isinstance(bead, VirtualMap)
                        for k, atom in enumerate(bead.atoms):
                            if virt:
                                try:
                                    l_off.append(names_in_mol.index(atom))
                                except ValueError:
                                    raise KeyError(atom)
                            else:
                                try:
                                    l_off.append(atom_names.index(atom))
                                except ValueError:
                                    raise KeyError(atom)
                            l_w.append(0.0 if virt else bead.weights[k])
                            l_vw.append(bead.weights[k] if virt else 0.0)
                            l_isv_entry.append(virt)
                        l_ptr.append(len(l_off))
                        l_virt.append(virt)
                    l_off = np.array(l_off, dtype=np.int64)
                    l_isv_entry = np.array(l_isv_entry, dtype=bool)
                    nb, nz = len(mol_map), len(l_off)
                    b0 = bead0_r[members]
                    z0 = nnz0_r[members]
                    base = np.where(l_isv_entry[None, :], b0[:, None], top.res_first[mapped_res[members]][:, None])
                    index[(z0[:, None] + np.arange(nz)[None, :]).reshape(-1)] = (base + l_off[None, :]).reshape(-1)
                    weights[(z0[:, None] + np.arange(nz)[None, :]).reshape(-1)] = np.tile(
                        np.array(l_w, dtype=real_dtype), len(members))
                    vweights[(z0[:, None] + np.arange(nz)[None, :]).reshape(-1)] = np.tile(
                        np.array(l_vw, dtype=virt_dtype), len(members))
                    bead_ptr[(b0[:, None] + 1 + np.arange(nb)[None, :]).reshape(-1)] = (
                        z0[:, None] + np.array(l_ptr[1:], dtype=np.int64)[None, :]).reshape(-1)
                    is_virtual[(b0[:, None] + np.arange(nb)[None, :]).reshape(-1)] = np.tile(
                        np.array(l_virt, dtype=np.uint8), len(members))
            for b in mol_map:
                if b.name not in bead_names:
                    bead_names.append(b.name)
            ids = np.array([bead_names.index(b.name) for b in mol_map], dtype=np.int32)
            members_all = np.nonzero(rid == i)[0]
            bead_name_id[(bead0_r[members_all][:, None] + np.arange(len(mol_map))[None, :]).reshape(-1)] = np.tile(
                ids, len(members_all))

        plan.n_beads = n_beads
        plan.bead_ptr, plan.index = bead_ptr, index
        plan.weights, plan.vweights, plan.is_virtual = weights, vweights, is_virtual
        plan.cg_topology = Topology.from_arrays(
            [name_of[i] for i in mapped_name_ids],
            np.searchsorted(np.array(mapped_name_ids, dtype=np.int64), rid).astype(np.int32) if n_mres else [],
            bead0_r, bead_names, bead_name_id)
        return plan

    def _plan_for(self, topology):
        key = id(topology)
        cached = self._plans.get(key)
        if cached is None or cached[0] is not topology or cached[1].n_atoms != topology.n_atoms:
            cached = (topology, self.compile_plan(topology))
            self._plans = {key: cached}
        return cached[1]

    @staticmethod
    def _device_plan(plan, dev):
        """Tile plan + CSR tables on the device (built once per plan and device)."""
        key = str(dev)
        if key in plan._dev:
            return plan._dev[key]
        lib = _lib.load()
        import ctypes
        import torch
        d = {}
        span_of = None
        compute = (plan.is_virtual == 0).astype(np.uint8)
        if plan.n_beads:
            # beads wider than a tile go through the gather kernel
            cnt = np.diff(plan.bead_ptr)
            nz = cnt > 0
            lo = np.full(plan.n_beads, 0, dtype=np.int64)
            hi = np.full(plan.n_beads, 0, dtype=np.int64)
            if len(plan.index):
                starts = plan.bead_ptr[:-1][nz]
                lo[nz] = np.minimum.reduceat(plan.index, starts)
                hi[nz] = np.maximum.reduceat(plan.index, starts)
            span_of = hi - lo + 1
            wide = (span_of > MAX_SPAN_ATOMS) & (plan.is_virtual == 0)
            compute[wide] = 0
            compute[~nz] = 0
        else:
            wide = np.zeros(0, dtype=bool)
        w64 = plan.weights.dtype == np.float64
        n_tiles, n_ent = ctypes.c_int64(), ctypes.c_int64()
        max_span, max_ent = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(lib.cgb_map_plan_size(_lib.ptr(plan.bead_ptr), _lib.ptr(plan.index), _lib.ptr(compute),
                                         plan.n_beads, plan.n_atoms, MAX_SPAN_ATOMS, MAX_TILE_BEADS,
                                         ctypes.byref(n_tiles), ctypes.byref(n_ent),
                                         ctypes.byref(max_span), ctypes.byref(max_ent)), "cgb_map_plan_size")
        tiles = np.zeros(n_tiles.value * 8, dtype=np.int32)          # CgbMapTile = 8 x int32
        entries = np.zeros(n_ent.value * 2, dtype=np.int32)          # CgbMapEntry = 2 x int32
        ew64 = np.zeros(n_ent.value, dtype=np.float64) if w64 else None
        _lib.check(lib.cgb_map_plan_build(_lib.ptr(plan.bead_ptr), _lib.ptr(plan.index), _lib.ptr(plan.weights),
                                          int(w64), _lib.ptr(compute), plan.n_beads, plan.n_atoms,
                                          MAX_SPAN_ATOMS, MAX_TILE_BEADS, _lib.ptr(tiles), _lib.ptr(entries),
                                          _lib.ptr(ew64)), "cgb_map_plan_build")
        d["n_tiles"], d["max_span"], d["max_ent"] = n_tiles.value, max_span.value, max_ent.value
        d["tiles_host"], d["entries_host"] = tiles, entries
        d["tiles"] = device.to_device(tiles, dev) if n_tiles.value else None
        d["entries"] = device.to_device(entries, dev) if n_ent.value else None
        d["entries_w64"] = device.to_device(ew64, dev) if (w64 and n_ent.value) else None
        d["w64"] = w64
        gather_real = np.nonzero(wide)[0].astype(np.int32)
        gather_virt = np.nonzero(plan.is_virtual)[0].astype(np.int32)
        if len(gather_real) or len(gather_virt):
            d["bead_ptr"] = device.to_device(plan.bead_ptr, dev)
            d["index"] = device.to_device(plan.index, dev)
        if len(gather_real):
            d["wide_list"] = device.to_device(gather_real, dev)
            d["weights"] = device.to_device(plan.weights, dev)
        if len(gather_virt):
            d["virt_list"] = device.to_device(gather_virt, dev)
            d["vweights"] = device.to_device(plan.vweights, dev)
        d["n_wide"], d["n_virt"] = len(gather_real), len(gather_virt)
        torch.cuda.current_stream().synchronize()
        plan._dev[key] = d
        return d

    # ------------------------------------------------------------------ apply (mapping.py:391-443)
    def map_device(self, plan, xyz_dev, box_lengths, out=None, window=None):
        """Run the mapping kernels on device-resident coordinates.

        ``xyz_dev``: CUDA float32 tensor (F, N, 3); ``box_lengths``: numpy / tensor
        (F, 3) or None.  Returns the CUDA tensor (F, B, 3).  Asynchronous.
        ``window=(first_atom, n_atoms)``: ``xyz_dev`` holds only that atom range of every frame
        (``atom_window``), as uploaded by ``map_streamed``.
        """
        import torch
        lib = _lib.load()
        dev = xyz_dev.device
        n_frames = int(xyz_dev.shape[0])
        first_atom, held = (0, plan.n_atoms) if window is None else (int(window[0]), int(window[1]))
        if xyz_dev.dtype != torch.float32 or not xyz_dev.is_contiguous() or int(xyz_dev.shape[1]) != held:
            raise ValueError("xyz must be a contiguous float32 (F, N, 3) tensor matching the topology")
        if out is None:
            out = torch.empty((n_frames, plan.n_beads, 3), dtype=torch.float32, device=dev)
        if plan.n_beads == 0 or n_frames == 0:
            return out
        d = self._device_plan(plan, dev)
        box_dev = None
        if box_lengths is not None:
            box_dev = device.to_device(box_lengths, dev, dtype=np.float32) if not isinstance(
                box_lengths, torch.Tensor) else box_lengths.to(dev).contiguous()
        stream = device.stream_handle()
        if window is not None and d["n_wide"]:
            raise ValueError("an atom window cannot be combined with beads wider than a tile")
        if d["n_tiles"]:
            _lib.check(lib.cgb_map_tiles_ex(device.dptr(xyz_dev), device.dptr(box_dev), n_frames, held, first_atom,
                                            plan.n_beads, device.dptr(d["tiles"]), d["n_tiles"],
                                            device.dptr(d["entries"]), device.dptr(d["entries_w64"]),
                                            d["max_span"], d["max_ent"], device.dptr(out),
                                            _lib.tuning_ptr(getattr(self, "tuning", None)), stream),
                       "cgb_map_tiles_ex")
        if d["n_wide"]:
            _lib.check(lib.cgb_map_gather(device.dptr(xyz_dev), plan.n_atoms, device.dptr(box_dev), n_frames,
                                          plan.n_beads, device.dptr(d["bead_ptr"]), device.dptr(d["index"]),
                                          device.dptr(d["weights"]), int(d["w64"]), device.dptr(d["wide_list"]),
                                          d["n_wide"], device.dptr(out), stream), "cgb_map_gather")
        if d["n_virt"]:
            # virtual beads read the CG frame, after the real beads (mapping.py:431-438)
            _lib.check(lib.cgb_map_gather(device.dptr(out), plan.n_beads, device.dptr(box_dev), n_frames,
                                          plan.n_beads, device.dptr(d["bead_ptr"]), device.dptr(d["index"]),
                                          device.dptr(d["vweights"]), int(plan.vweights.dtype == np.float64),
                                          device.dptr(d["virt_list"]), d["n_virt"], device.dptr(out), stream),
                       "cgb_map_gather")
        return out

    @staticmethod
    def atom_window(plan):
        """``(first_atom, n_atoms)``: the contiguous range of input atoms the real beads are built from.  When a
        host trajectory is streamed only this range is uploaded -- unmapped solvent after (or before) the mapped
        molecules never crosses PCIe."""
        if getattr(plan, "_atom_window", None) is None:
            real_entry = np.repeat(plan.is_virtual == 0, np.diff(plan.bead_ptr))
            atoms = plan.index[real_entry]
            plan._atom_window = (0, plan.n_atoms) if len(atoms) == 0 else (int(atoms.min()), int(atoms.max()) + 1 - int(atoms.min()))
        return plan._atom_window

    def map_streamed(self, plan, xyz_host, box_lengths, chunk_bytes=256 << 20, out_host=None):
        """Map a host-resident trajectory through a double-buffered H2D -> kernel -> D2H pipeline.

        ``xyz_host``: numpy float32 (F, N, 3) (page-locked memory makes the copies asynchronous).
        Returns ``(cg_device, cg_host)``: the CG coordinates on the GPU and their host copy; both are
        complete when this function returns.  Frames are independent, so a chunk of frames is the unit
        of work; the input never has to fit in HBM.
        """
        import torch
        dev = device.default_device()
        n_frames, n_atoms = int(xyz_host.shape[0]), plan.n_atoms
        src = torch.from_numpy(xyz_host)
        lib = _lib.load()
        # upload only the atoms the mapping reads when that saves at least a tenth of the transfer
        first_atom, held = self.atom_window(plan)
        windowed = (held <= 0.9 * n_atoms and held * 12 > (64 << 10) and not self._device_plan(plan, dev)["n_wide"]
                    and src.is_contiguous())
        if not windowed:
            first_atom, held = 0, n_atoms
        chunk = int(max(1, min(n_frames, chunk_bytes // max(12 * held, 1))))
        cg_dev = torch.empty((n_frames, plan.n_beads, 3), dtype=torch.float32, device=dev)
        if out_host is None:
            out_host = torch.empty((n_frames, plan.n_beads, 3), dtype=torch.float32, pin_memory=True)
        box_dev = None if box_lengths is None else device.to_device(box_lengths, dev, dtype=np.float32)
        bufs = [torch.empty((chunk, held, 3), dtype=torch.float32, device=dev) for _ in range(2)]
        main = torch.cuda.current_stream()
        h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
        h2d.wait_stream(main)
        d2h.wait_stream(main)
        mapped = [None, None]           # event: kernel finished reading input buffer i
        for i, f0 in enumerate(range(0, n_frames, chunk)):
            f1 = min(n_frames, f0 + chunk)
            slot = i % 2
            with torch.cuda.stream(h2d):
                if mapped[slot] is not None:
                    h2d.wait_event(mapped[slot])
                if windowed:
                    _lib.check(lib.cgb_copy_window_h2d(device.dptr(bufs[slot]), src[f0:f1].data_ptr(), f1 - f0, n_atoms,
                                                       first_atom, held, h2d.cuda_stream), "cgb_copy_window_h2d")
                else:
                    bufs[slot][:f1 - f0].copy_(src[f0:f1], non_blocking=True)
                loaded = h2d.record_event()
            main.wait_event(loaded)
            self.map_device(plan, bufs[slot][:f1 - f0], None if box_dev is None else box_dev[f0:f1],
                            out=cg_dev[f0:f1], window=(first_atom, held) if windowed else None)
            mapped[slot] = main.record_event()
            with torch.cuda.stream(d2h):
                d2h.wait_event(mapped[slot])
                out_host[f0:f1].copy_(cg_dev[f0:f1], non_blocking=True)
        main.wait_stream(d2h)
        main.synchronize()
        return cg_dev, out_host

    def apply(self, frame, cg_frame=None, out=None):
        """Apply the AA->CG mapping to ``frame``; returns the CG ``Frame``.

        ``out``: optional CUDA float32 tensor ``(F, B, 3)`` to write the bead coordinates into (a pipeline that
        maps block after block reuses one buffer instead of allocating per call).

        The CG coordinates stay on the GPU (``cg._trajectory.device_xyz()``) and are
        copied to the host only when ``.xyz`` / ``atom.coords`` is read.  A large
        host-resident input is streamed through the GPU in frame chunks
        (``map_streamed``) instead of being uploaded in one piece.
        """
        device.require_cuda()
        logger.info("Applying AA->CG mapping")
        plan = self._plan_for(frame._topology)
        traj = frame._trajectory
        # mapping.py:406-408: lengths only, and no PBC at all if any length of any frame is zero
        box = frame.unitcell_lengths
        # evaluated over the whole trajectory, i.e. over every frame shard when sharded across GPUs
        if not traj.global_box_flags(device.default_device())[0]:
            box = None
        host_xyz = None
        if not traj.on_device and traj.xyz.nbytes >= STREAM_THRESHOLD_BYTES and plan.n_beads:
            out, host = self.map_streamed(plan, traj.xyz, box)
            host_xyz = host.numpy()
        else:
            xyz_dev = traj.device_xyz()
            if out is not None and (tuple(out.shape) != (traj.n_frames, plan.n_beads, 3) or not out.is_cuda):
                raise ValueError("out must be a CUDA float32 tensor of shape (n_frames, n_beads, 3)")
            out = self.map_device(plan, xyz_dev, None if box is None else traj.device_box_lengths(xyz_dev.device),
                                  out=out)
        new = Frame.from_arrays(plan.cg_topology if cg_frame is None else cg_frame._topology, xyz=host_xyz,
                                time=frame.time, xyz_device=out)
        new._trajectory.share_unitcell(traj)        # same unit cell: nothing is derived (or decided) twice
        new._attach(new._trajectory)
        if cg_frame is not None and hasattr(cg_frame, "_trajectory"):
            cg_frame._trajectory += new._trajectory
            cg_frame._attach(cg_frame._trajectory)
            return cg_frame
        logger.info("Finished applying AA->CG mapping")
        return new


    def apply_xtc(self, frame, path, block_bytes=128 << 20):
        """``apply`` on an .xtc file that is never expanded as a whole: the compressed image is uploaded in blocks of
        frames, every block is decoded on the GPU into a pipeline buffer and mapped from there, and only the CG
        trajectory is kept (reference: ``Frame(gro, xtc)`` through mdtraj's codec, ``pycgtool/frame.py:41-63``, then
        ``Mapping.apply``).  A trajectory whose raw coordinates exceed the GPU's memory flows through this way.
        ``frame``: a Frame holding the topology (e.g. ``Frame("system.gro")``)."""
        import torch
        from . import formats
        from .topology import lengths_and_angles
        device.require_cuda()
        dev = device.default_device()
        plan = self._plan_for(frame._topology)
        pinned, size, table, box, time, n_frames, n_atoms = formats.read_image_pinned(path)
        if n_atoms != plan.n_atoms:
            raise ValueError("Number of atoms does not match between topology and trajectory files")
        vectors = box if box.any() else None
        lengths = None
        if vectors is not None:
            lengths = lengths_and_angles(vectors)[0]
            if not dist.all_true(bool(np.all(lengths)), device=dev):        # mapping.py:406-408
                lengths = None
        lengths_dev = None if lengths is None else device.to_device(lengths, dev, dtype=np.float32)
        out = torch.empty((n_frames, plan.n_beads, 3), dtype=torch.float32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)

        def map_block(f0, f1, coords):
            self.map_device(plan, coords, None if lengths_dev is None else lengths_dev[f0:f1], out=out[f0:f1])

        if n_frames and plan.n_beads:
            formats.decode_image_device(pinned, size, table, n_frames, n_atoms, None, block_bytes=block_bytes,
                                        on_block=map_block, bad_frames=bad, device=dev)
        torch.cuda.current_stream().synchronize()
        if int(bad.item()):
            raise OSError(f"'{path}': {int(bad.item())} frame(s) hold a truncated or corrupt coordinate stream")
        return Frame.from_arrays(plan.cg_topology, xyz=None, unitcell_vectors=vectors, time=time, xyz_device=out)


def calc_coords_weight(ref_coords, coords, weights, box=None):
    """Single-bead host formula of ``mapping.py:446-463``, kept for API compatibility
    (tests and user scripts call it directly); the trajectory path does not use it."""
    vectors = coords - ref_coords
    if box is not None:
        vectors -= box * np.rint(vectors / box)
    result = np.sum(weights[np.newaxis, np.newaxis].T * vectors, axis=0)
    result += ref_coords
    return result
