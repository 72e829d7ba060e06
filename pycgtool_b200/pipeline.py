"""Mapping and bond measurement in one pass.

The reference runs ``Mapping.apply`` and then ``BondSet.apply`` on its result
(``PyCGTOOL.apply_mapping`` / ``measure_bonds``, ``pycgtool/__main__.py:69-112``): the CG
trajectory is written and read back.  ``map_and_measure`` does both in one kernel
(``cgb_map_measure``): the mapping warps of a CTA hand every mapped tile, still in shared
memory, to the measurement warps of the same CTA, so the CG trajectory is written to HBM once
(it is an output) and never re-read.

The fused kernel needs a plan in which every molecule instance ("unit": a bead range no
bonded term straddles) fits one tile of ``FUSED_TILE_BEADS`` beads together with its terms, bead
weights in float32 and no measured virtual bead.  When a system does not fit that mould,
``map_and_measure`` runs the two stages one after the other -- same results, one more pass over
the CG trajectory.
"""

import ctypes
import itertools
import logging

import numpy as np

from . import _lib, device, dist
from .frame import Frame

logger = logging.getLogger(__name__)

FUSED_TILE_BEADS = 48        # beads per fused tile: 144 output floats = 5 mapping warps
FUSED_SPAN_ATOMS = 1024      # input atoms a fused tile may span
FUSED_GROUPS = 2             # measurement groups (warps) per fused tile


class _FusedPlan:
    """Mapping tile plan that follows the measurement plan's units, and the pairing of the two."""

    def __init__(self):
        self.bond_plan = None         # MeasurePlan compiled with unit_groups
        self.n_tiles = 0
        self.max_span = self.max_ent = 0
        self.tiles = self.entries = self.meas_of_tile = None
        self._dev = {}


def _compile(mapping, bondset, map_plan):
    """Fused plan, or None when the system does not fit the fused kernel."""
    lib = _lib.load()
    if map_plan.weights.dtype != np.float32 or map_plan.n_beads == 0:
        return None
    bond_plan = bondset.compile_plan(map_plan.cg_topology, fused=(FUSED_GROUPS, FUSED_TILE_BEADS))
    if bond_plan is None or len(bond_plan.tiles) == 0 or sum(bond_plan.wide_counts):
        return None
    # every measured bead must be computed by a mapping tile (not virtual, not wider than a tile)
    cnt = np.diff(map_plan.bead_ptr)
    compute = ((map_plan.is_virtual == 0) & (cnt > 0)).astype(np.uint8)
    if len(map_plan.index):
        nz = cnt > 0
        starts = map_plan.bead_ptr[:-1][nz]
        span = np.zeros(map_plan.n_beads, dtype=np.int64)
        span[nz] = np.maximum.reduceat(map_plan.index, starts) - np.minimum.reduceat(map_plan.index, starts) + 1
        compute[span > FUSED_SPAN_ATOMS] = 0
    measured = np.unique(bond_plan.col_beads)
    if len(measured) and not compute[measured].all():
        return None
    if (compute == 0).any() and (map_plan.is_virtual == 0)[compute == 0].any():
        return None                   # a real bead wider than a tile: the separate path handles it
    fp = _FusedPlan()
    fp.bond_plan = bond_plan
    n_tiles, n_ent = ctypes.c_int64(), ctypes.c_int64()
    max_span, max_ent = ctypes.c_int32(), ctypes.c_int32()
    args = (_lib.ptr(map_plan.bead_ptr), _lib.ptr(map_plan.index))
    rc = lib.cgb_map_plan_size_cuts(*args, _lib.ptr(compute), map_plan.n_beads, map_plan.n_atoms, FUSED_SPAN_ATOMS,
                                    FUSED_TILE_BEADS, _lib.ptr(bond_plan.tile_start), ctypes.byref(n_tiles),
                                    ctypes.byref(n_ent), ctypes.byref(max_span), ctypes.byref(max_ent))
    if rc != 0:
        return None
    tiles = np.zeros(n_tiles.value * 8, dtype=np.int32)
    entries = np.zeros(n_ent.value * 2, dtype=np.int32)
    _lib.check(lib.cgb_map_plan_build_cuts(*args, _lib.ptr(map_plan.weights), 0, _lib.ptr(compute), map_plan.n_beads,
                                           map_plan.n_atoms, FUSED_SPAN_ATOMS, FUSED_TILE_BEADS,
                                           _lib.ptr(bond_plan.tile_start), _lib.ptr(tiles), _lib.ptr(entries), None),
               "cgb_map_plan_build_cuts")
    fp.n_tiles, fp.max_span, fp.max_ent = n_tiles.value, max_span.value, max_ent.value
    fp.tiles, fp.entries = tiles, entries
    # pair the plans: the mapping tile that starts at a measurement tile's first bead must cover its whole range
    t = tiles.reshape(-1, 8)
    bead0, nbeads = t[:, 2], t[:, 3]
    order = np.argsort(bead0, kind="stable")
    pos = np.searchsorted(bead0[order], bond_plan.tiles["bead0"])
    if (pos >= len(order)).any():
        return None
    match = order[pos]
    if not (np.array_equal(bead0[match], bond_plan.tiles["bead0"]) and
            (nbeads[match] >= bond_plan.tiles["span"]).all() and len(np.unique(match)) == len(match)):
        return None
    meas_of_tile = np.full(fp.n_tiles, -1, dtype=np.int32)
    meas_of_tile[match] = np.arange(len(match), dtype=np.int32)
    # launch order: tiles with bonded terms (instruction-bound) spread evenly among the others (HBM-bound), so that an
    # SM holds one of each at a time and the term arithmetic hides behind the streaming of the solvent
    with_terms = np.nonzero(meas_of_tile >= 0)[0]
    without = np.nonzero(meas_of_tile < 0)[0]
    if len(with_terms) and len(without):
        pos_a = (np.arange(len(with_terms)) + 0.5) / len(with_terms)
        pos_b = (np.arange(len(without)) + 0.5) / len(without)
        order = np.concatenate([with_terms, without])[np.argsort(np.concatenate([pos_a, pos_b]), kind="stable")]
        fp.tiles = np.ascontiguousarray(t[order]).reshape(-1)
        meas_of_tile = meas_of_tile[order]
    fp.meas_of_tile = np.ascontiguousarray(meas_of_tile)
    return fp


def _fused_plan(mapping, bondset, map_plan):
    key = (id(map_plan), id(bondset), int(bondset.histogram_bins))
    cache = getattr(mapping, "_fused_plans", None)
    if cache is None:
        cache = mapping._fused_plans = {}
    if key not in cache or cache[key][0] is not map_plan:
        cache.clear()
        cache[key] = (map_plan, _compile(mapping, bondset, map_plan))
    return cache[key][1]


def _device_fused(fp, dev):
    key = str(dev)
    if key not in fp._dev:
        fp._dev[key] = {
            "tiles": device.to_device(fp.tiles, dev),
            "entries": device.to_device(fp.entries, dev),
            "meas_of_tile": device.to_device(fp.meas_of_tile, dev),
        }
    return fp._dev[key]


def map_and_measure(mapping, bondset, frame, histogram=True, store_values=True, fused=None, out=None):
    """``cg = mapping.apply(frame); bondset.apply(cg)`` in one pass over the atoms; returns ``cg``.

    Afterwards ``bondset.boltzmann_invert()`` and every ``bond.values`` behave exactly as after the two
    separate calls.  ``fused=True`` asks for the single fused kernel (``cgb_map_measure``); the default runs the two
    kernels back to back on the device-resident trajectory, which is the faster of the two on B200: stage 2 is bound
    by instruction issue, not by HBM, so re-reading the beads costs less than sharing an SM's warps between mapping
    and measurement (measured in ``bench.py``'s ``pipeline`` leg).  A system the fused kernel does
    not support, or a host trajectory large enough to be streamed, always takes the two stages.  ``out``: optional
    CUDA tensor ``(F, B, 3)`` for the bead coordinates (see ``Mapping.apply``).
    """
    import torch
    device.require_cuda()
    lib = _lib.load()
    map_plan = mapping._plan_for(frame._topology)
    traj = frame._trajectory
    fp = _fused_plan(mapping, bondset, map_plan) if fused else None
    bondset._release()                  # the previous pass's CG trajectory and values go before new ones are allocated
    from . import mapping as _mapping_module
    streamed = (not traj.on_device) and traj.xyz.nbytes >= _mapping_module.STREAM_THRESHOLD_BYTES
    if fp is None or streamed:
        cg = mapping.apply(frame, out=out)
        bondset.apply(cg, histogram=histogram, store_values=store_values)
        state = bondset._state
        if not streamed and state is not None and state.n_frames:
            xyz_in, cg_out = traj.device_xyz(), cg._trajectory.device_xyz()
            box = frame.unitcell_lengths
            if not traj.global_box_flags(xyz_in.device)[0]:
                box = None
            box_len_dev = None if box is None else device.to_device(box, xyz_in.device, dtype=np.float32)
            nbins_ = 0 if state.hist is None else int(state.hist.shape[1])
            # the closure holds the buffers, not the state (state -> closure -> state would keep a whole CG
            # trajectory alive until the cycle collector runs)
            plan_, source_, nf_, shift_ = state.plan, state.source, state.n_frames, state.shift
            lo_, hi_ = state.hist_lo, state.hist_hi

            def relaunch(packed_=state.packed, values_=state.values):
                """The two launches again (K1 then K2) into the given accumulators; bench.py times this."""
                mapping.map_device(map_plan, xyz_in, box_len_dev, out=cg_out)
                bondset._launch_measure(plan_, source_, nf_, shift_, values_, packed_, nbins_, lo_, hi_)

            state.relaunch = relaunch
        return cg

    from .bondset import BondSet, _MeasureState
    xyz = traj.device_xyz()
    dev = xyz.device
    n_frames = int(xyz.shape[0])
    plan = fp.bond_plan
    n_cols = plan.n_cols
    # mapping: box lengths only, and none at all if any length of any frame is zero (mapping.py:406-408)
    box = frame.unitcell_lengths
    lengths_ok, has_vectors, ortho = traj.global_box_flags(dev)
    if not lengths_ok:
        box = None
    box_len = device.to_device(box, dev, dtype=np.float32) if box is not None else None
    # measurement: full box vectors, orthorhombic fast path as mdtraj decides it (see BondSet.apply)
    box_vec = traj.device_box_vectors(dev) if has_vectors else None
    ortho = int(ortho)

    d_map = mapping._device_plan(map_plan, dev)         # virtual beads, and K1 for the frame-0 shift
    d_fused = _device_fused(fp, dev)
    d = BondSet._device_plan(plan, dev)
    if "forms" not in d:
        d["forms"] = device.to_device(bondset._form_ids(), dev)
    stream = device.stream_handle()
    if out is None:
        out = torch.empty((n_frames, map_plan.n_beads, 3), dtype=torch.float32, device=dev)
    nbins = int(bondset.histogram_bins) if (histogram and plan.n_bonds) else 0
    values = torch.empty((n_frames, n_cols), dtype=torch.float32, device=dev) if store_values else None
    packed = torch.zeros(plan.n_bonds * (_lib.NPART + nbins), dtype=torch.float64, device=dev)
    npart = plan.n_bonds * _lib.NPART
    hist_lo = hist_hi = lo = hi = None
    if nbins:
        lo = np.array([bondset.histogram_ranges[int(k)][0] for k in plan.natoms], dtype=np.float32)
        hi = np.array([bondset.histogram_ranges[int(k)][1] for k in plan.natoms], dtype=np.float32)
        hist_lo, hist_hi = device.to_device(lo, dev), device.to_device(hi, dev)

    shift = torch.zeros(plan.n_bonds, dtype=torch.float32, device=dev)
    if n_frames:
        # shift of the one-pass statistics: the bonds' first values in frame 0 (of rank 0), from a one-frame K1
        mapping.map_device(map_plan, xyz[:1], None if box is None else box[:1], out=out[:1])
        _lib.check(lib.cgb_measure_shift(device.dptr(out), device.dptr(box_vec), ortho, 0, 1, map_plan.n_beads,
                                         device.dptr(d["bond_first"]), device.dptr(d["natoms_measured"]),
                                         plan.n_bonds, device.dptr(shift), stream), "cgb_measure_shift")
    dist.broadcast_device(shift, src=0)

    def launch(packed_=packed, values_=values):
        """One pass of the fused kernel (+ the virtual beads) into the given accumulators; bench.py times this."""
        _lib.check(lib.cgb_map_measure(
            device.dptr(xyz), device.dptr(box_len), device.dptr(box_vec), ortho, n_frames, map_plan.n_atoms,
            map_plan.n_beads, device.dptr(d_fused["tiles"]), fp.n_tiles, device.dptr(d_fused["entries"]), fp.max_span,
            fp.max_ent, FUSED_TILE_BEADS, device.dptr(out), device.dptr(d_fused["meas_of_tile"]),
            device.dptr(d["tiles"]), device.dptr(d["groups"]), device.dptr(d["lanes"]), device.dptr(d["tile_bonds"]),
            device.dptr(d["tile_lists"]), plan.shape, plan.max_slots, device.dptr(shift), device.dptr(values_), n_cols,
            device.dptr(packed_[:npart]), device.dptr(packed_[npart:] if nbins else None), device.dptr(hist_lo),
            device.dptr(hist_hi), nbins, stream), "cgb_map_measure")
        if d_map["n_virt"]:
            # virtual beads read the CG frame, after the real beads (mapping.py:431-438)
            _lib.check(lib.cgb_map_gather(device.dptr(out), map_plan.n_beads, device.dptr(box_len), n_frames,
                                          map_plan.n_beads, device.dptr(d_map["bead_ptr"]), device.dptr(d_map["index"]),
                                          device.dptr(d_map["vweights"]), int(map_plan.vweights.dtype == np.float64),
                                          device.dptr(d_map["virt_list"]), d_map["n_virt"], device.dptr(out), stream),
                       "cgb_map_gather")

    if n_frames:
        launch()

    cg = Frame.from_arrays(map_plan.cg_topology, xyz=None, time=frame.time, xyz_device=out)
    cg._trajectory.share_unitcell(traj)
    cg._attach(cg._trajectory)
    state = _MeasureState(bondset, plan, n_frames, values, packed, nbins, shift, hist_lo, hist_hi, dev,
                          (out, box_vec, ortho))
    state.hist_lo_host, state.hist_hi_host = lo, hi
    state.relaunch = launch
    bondset._state = state
    for bond_id, bond in enumerate(itertools.chain(*bondset._molecules.values())):
        bond._values = None
        bond._source = (state, bond_id)
        bond.eqm = bond.fconst = None
        bond._histogram = None
    return cg
