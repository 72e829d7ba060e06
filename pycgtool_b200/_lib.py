"""ctypes binding of ``libpycgtool_b200.so`` (declared in ``include/pycgtool_b200.h``).

There is no fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised.
"""

import ctypes
import pathlib

_CSRC = pathlib.Path(__file__).resolve().parent / "csrc"
LIBRARY_PATH = _CSRC / "libpycgtool_b200.so"

NPART = 8  # CGB_NPART
NOUT = 5  # CGB_NOUT
MEASURE_TILE = 224  # CGB_MEASURE_TILE
MEAS_WARPS = 7  # CGB_MEAS_WARPS
MEAS_CHUNKS = 4  # CGB_MEAS_CHUNKS
MEAS_EDGE_CHUNKS = 2  # CGB_MEAS_EDGE_CHUNKS
SHAPE_EAD, SHAPE_EEAA = 0, 1  # CGB_SHAPE_*: [edges | angles | dihedrals | -] / [edges | edges | angles | angles]
COMM_ID_BYTES = 128  # CGB_COMM_ID_BYTES
ABI_VERSION = 3  # CGB_ABI_VERSION

FORM_IDS = {
    "Harmonic": 0,
    "CosHarmonic": 1,
    "MartiniDefaultLength": 2,
    "MartiniDefaultAngle": 3,
    "MartiniDefaultDihedral": 4,
}
FORM_STATS_ONLY = 5


XTC_FRAME_BYTES = 48  # sizeof(CgbXtcFrame)


class MeasPlanInfo(ctypes.Structure):
    _fields_ = [
        ("n_tiles", ctypes.c_int64),
        ("n_groups", ctypes.c_int64),
        ("n_tile_bonds", ctypes.c_int64),
        ("n_tile_lists", ctypes.c_int64),
        ("n_wide", ctypes.c_int64),
        ("n_staged", ctypes.c_int64),
        ("max_span", ctypes.c_int32),
        ("max_slots", ctypes.c_int32),
        ("max_edge_chunks", ctypes.c_int32),
        ("max_tile_groups", ctypes.c_int32),
    ]


class MapTuning(ctypes.Structure):
    """``CgbMapTuning``: per-call benchmarking knobs of K1 (zero keeps the default)."""
    _fields_ = [
        ("stage_bytes", ctypes.c_int32),
        ("n_stages", ctypes.c_int32),
        ("frames_per_cta", ctypes.c_int32),
        ("consumer_threads", ctypes.c_int32),
        ("frame_group", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 3),
    ]


class MeasureTuning(ctypes.Structure):
    """``CgbMeasureTuning``: per-call benchmarking knobs of K2 (zero keeps the default)."""
    _fields_ = [
        ("stage_bytes", ctypes.c_int32),
        ("n_stages", ctypes.c_int32),
        ("max_run", ctypes.c_int32),
        ("ctas_per_sm", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 4),
    ]


#: numpy dtypes of the plan structs of include/pycgtool_b200.h (CgbMeasTile, CgbMeasGroup, CgbMeasLane)
MEAS_TILE_DTYPE = [("bead0", "<i4"), ("span", "<i4"), ("group0", "<i4"), ("ngroups", "<i4"),
                   ("slot0", "<i4"), ("nslots", "<i4"), ("list0", "<i4"), ("nlist", "<i4")]
MEAS_GROUP_DTYPE = [("kind", "u1", (4,)), ("nlanes", "u1", (4,)), ("nemit", "u1", (4,)),
                    ("vcol0", "<i4", (4,)), ("reserved", "<i4")]
MEAS_LANE_DTYPE = [("ref", "<u4"), ("slot", "<i4")]


class MapTile(ctypes.Structure):
    _fields_ = [
        ("atom0", ctypes.c_int32),
        ("span", ctypes.c_int32),
        ("bead0", ctypes.c_int32),
        ("nbeads", ctypes.c_int32),
        ("ent0", ctypes.c_int64),
        ("nslots", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_int = ctypes.c_int

# name -> (restype, argtypes); mirrors include/pycgtool_b200.h one to one
SIGNATURES = {
    "cgb_strerror": (ctypes.c_char_p, [_int]),
    "cgb_abi_version": (_int, []),
    "cgb_map_plan_size": (_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cgb_map_plan_build": (_int, [_vp, _vp, _vp, _int, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "cgb_map_plan_size_cuts": (_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "cgb_map_plan_build_cuts": (_int, [_vp, _vp, _vp, _int, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cgb_map_tiles": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "cgb_map_tiles_window": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "cgb_copy_window_h2d": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "cgb_map_gather": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _int, _vp, _i64, _vp, _vp]),
    "cgb_map_tiles_ex": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "cgb_measure_plan_limits": (_int, [_i32, _int, _i32, _vp, _vp]),
    "cgb_measure_plan_size": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp]),
    "cgb_measure_plan_build": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _vp, _vp]),
    "cgb_measure_fused": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32,
                                 _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "cgb_map_measure": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "cgb_measure_fused_config": (_int, [_i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _int, _vp, _i32, _vp]),
    "cgb_measure": (_int, [_vp, _vp, _int, _i64, _i64, _vp, _vp, _vp, _i32, _i64, _i64, _i64,
                           _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp]),
    "cgb_measure_shift": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    "cgb_circular_pass2": (_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _vp]),
    "cgb_value_stats": (_int, [_vp, _i64, _int, ctypes.c_float, _vp, _vp]),
    "cgb_boltzmann": (_int, [_vp, _vp, _vp, _vp, ctypes.c_double, _i64, _vp, _vp, _vp, _i32, _int, _vp, _vp]),
    "cgb_comm_unique_id": (_int, [_vp]),
    "cgb_comm_init": (_int, [_vp, _i32, _i32, _vp]),
    "cgb_comm_destroy": (_int, [_vp]),
    "cgb_allreduce_partials": (_int, [_vp, _vp, _i64, _vp]),
    "cgb_comm_broadcast": (_int, [_vp, _vp, _i64, _i32, _vp]),
    "cgb_xtc_scan": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cgb_xtc_workspace_bytes": (_i64, [_i64, _i64]),
    "cgb_xtc_decode": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    "cgb_xtc_decode_host": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "cgb_xtc_encode": (_int, [_vp, _vp, _vp, _i64, _i64, ctypes.c_float, _vp, _i64, _vp]),
}

_lib = None


def load():
    """Load the shared library once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIBRARY_PATH.exists():
        raise RuntimeError(
            f"{LIBRARY_PATH} is missing: build it with `python -m pycgtool_b200.build` "
            "(pycgtool_b200 has no CPU fallback)")
    lib = ctypes.CDLL(str(LIBRARY_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cgb_abi_version() != ABI_VERSION:
        raise RuntimeError("libpycgtool_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(code, what="libpycgtool_b200 call"):
    """Turn a non-zero return code into an exception."""
    if code != 0:
        text = load().cgb_strerror(code).decode()
        raise RuntimeError(f"{what} failed with code {code}: {text}")


def tuning_ptr(tuning):
    """``ctypes`` pointer of a tuning struct, or None (built-in defaults)."""
    return None if tuning is None else ctypes.cast(ctypes.pointer(tuning), ctypes.c_void_p)


def ptr(array):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if array is None:
        return None
    if not array.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return array.ctypes.data
