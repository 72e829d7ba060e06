/*
 * pycgtool_b200.h -- C ABI of libpycgtool_b200.so
 *
 * B200 (sm_100a) implementation of PyCGTOOL's per-frame hot path:
 *   stage 1  atom -> bead mapping            (reference pycgtool/mapping.py:391-463)
 *   stage 2  bonded-term measurement + stats (reference pycgtool/bondset.py:496-531, which calls
 *            mdtraj.compute_distances/angles/dihedrals, and pycgtool/util.py:24-53)
 *   stage 3  Boltzmann inversion             (reference pycgtool/bondset.py:62-91,533-539 and
 *            pycgtool/functionalforms.py:92-148)
 *
 * The reference is pure Python and has no FFI of its own; these entry points are what a ctypes
 * binding inside Mapping.apply / BondSet.apply / BondSet.boltzmann_invert would call (see
 * INTEGRATION.md).  Conventions:
 *   - plain pointers and sizes only; no C++ or torch types cross the boundary;
 *   - "device" pointers are CUDA device memory owned by the caller; "host" pointers are ordinary memory;
 *   - every launch is enqueued on `stream` (a cudaStream_t passed as void*, NULL = default stream)
 *     and returns without synchronising; the library allocates nothing and keeps no state besides the
 *     process-wide benchmarking knobs (cgb_map_set_tuning / cgb_map_set_shape / cgb_measure_set_tuning);
 *   - return 0 on success, a positive cudaError_t on a CUDA failure, a negative CGB_E* on bad arguments;
 *   - all index tables are int32; frames, atoms and beads are counted in int64.
 */
#ifndef PYCGTOOL_B200_H
#define PYCGTOOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_ABI_VERSION 1

enum {
    CGB_OK = 0,
    CGB_EINVAL = -1,      /* NULL pointer, negative size, inconsistent table */
    CGB_ERANGE = -2,      /* index outside the atom / bead range */
    CGB_ETOOBIG = -3,     /* a single bead spans more atoms than a shared-memory tile can hold */
    CGB_EALIGN = -4       /* pointer not aligned to its element type */
};

/* Human-readable text for any return code of this library (CUDA codes included). */
const char* cgb_strerror(int code);
int cgb_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * Stage 1: mapping.
 *
 * Bead b (output order = reference CG topology order, mapping.py:359-389) is defined by the CSR row
 * index[bead_ptr[b] .. bead_ptr[b+1]) with one weight per entry.  For a real bead the entries are
 * global atom indices inside one residue (mapping.py:419-429); the first entry is the reference atom.
 *   n == 1 : bead = x[a0]
 *   n  > 1 : v_i = x[a_i] - x[a0];  if box: v_i -= L * rint(v_i / L)   (L = unit-cell *lengths* of the
 *            frame; IEEE divide, round-half-even, no fused multiply-add);
 *            bead = (sum_i w_i * v_i, summed in entry order) + x[a0]      (mapping.py:446-463)
 * Arithmetic is float32 when the weights are float32 and float64 when they are float64 (ITP mass
 * weights, mapping.py:224-227), the result is stored as float32 either way (frame.py:248).
 * ---------------------------------------------------------------------------------------------- */

/* One shared-memory tile of the mapping plan (32 bytes). */
typedef struct CgbMapTile {
    int32_t atom0;    /* first atom of the contiguous input span the tile stages   */
    int32_t span;     /* atoms in that span                                         */
    int32_t bead0;    /* first output bead of the tile                              */
    int32_t nbeads;   /* output beads of the tile (contiguous in the output)        */
    int64_t ent0;     /* offset of the tile's first entry in the entry array        */
    int32_t nslots;   /* ELL depth: max atoms per bead in this tile                 */
    int32_t reserved;
} CgbMapTile;

/* One ELL entry (8 bytes).  Entries of a tile are slot-major: entry(slot, j) = ent0 + slot*nbeads + j.
 * slot 0: off3 = 3 * (reference atom - atom0), bits = atom count of the bead (0 = bead not computed by
 * this kernel, e.g. a virtual bead).  slot >= 1: off3 = 3 * (atom - atom0), bits = float32 weight bits
 * (padding entries repeat the reference atom with weight 0). */
typedef struct CgbMapEntry {
    int32_t off3;
    int32_t bits;
} CgbMapEntry;

/* Host-side plan compiler, pass 1: count tiles / entries for a CSR mapping restricted to the beads
 * with compute[b] != 0 (compute == NULL means all).  max_span_atoms bounds the input span of a tile
 * and max_tile_beads its bead count.  All pointers are HOST pointers. */
int cgb_map_plan_size(const int32_t* bead_ptr, const int32_t* index, const uint8_t* compute,
                      int64_t n_beads, int64_t n_atoms, int32_t max_span_atoms, int32_t max_tile_beads,
                      int64_t* n_tiles, int64_t* n_entries, int32_t* max_span, int32_t* max_tile_entries);

/* Pass 2: fill caller-allocated HOST arrays sized by pass 1.  `weights` is float32[nnz] or
 * float64[nnz] (weights_f64 != 0); entries_w64 (float64 per entry) is written only in the latter case
 * and may be NULL otherwise. */
int cgb_map_plan_build(const int32_t* bead_ptr, const int32_t* index, const void* weights, int weights_f64,
                       const uint8_t* compute, int64_t n_beads, int64_t n_atoms,
                       int32_t max_span_atoms, int32_t max_tile_beads,
                       CgbMapTile* tiles, CgbMapEntry* entries, double* entries_w64);

/* K1: tiled mapping kernel.  xyz: device float32 [F][N][3]; box_len: device float32 [F][3] or NULL
 * (no periodic unwrap -- the caller passes NULL when the frame has no box or any length of any frame
 * is zero, mapping.py:406-408); tiles/entries(/entries_w64): device copies of the plan; cg_xyz: device
 * float32 [F][B][3].  Beads not covered by the plan are left untouched. */
int cgb_map_tiles(const float* xyz, const float* box_len, int64_t n_frames, int64_t n_atoms, int64_t n_beads,
                  const CgbMapTile* tiles, int64_t n_tiles, const CgbMapEntry* entries, const double* entries_w64,
                  int32_t max_span, int32_t max_tile_entries, float* cg_xyz, void* stream);

/* K1 over an atom window: xyz holds only atoms [first_atom, first_atom + n_atoms) of every frame
 * ([F][n_atoms][3]); the tile plan keeps its absolute atom numbers and must not reference atoms outside the
 * window.  With first_atom = 0 this is cgb_map_tiles.  Used when a host trajectory is streamed: unmapped
 * solvent never crosses PCIe. */
int cgb_map_tiles_window(const float* xyz, const float* box_len, int64_t n_frames, int64_t n_atoms, int64_t first_atom,
                         int64_t n_beads, const CgbMapTile* tiles, int64_t n_tiles, const CgbMapEntry* entries,
                         const double* entries_w64, int32_t max_span, int32_t max_tile_entries, float* cg_xyz,
                         void* stream);

/* Host -> device copy of the atom window [first_atom, first_atom + n_window_atoms) of n_frames frames of a host
 * trajectory [F][n_atoms][3] (page-locked for an asynchronous copy) into a dense device buffer
 * [F][n_window_atoms][3]: one strided copy (cudaMemcpy2DAsync) on `stream`. */
int cgb_copy_window_h2d(float* dst, const float* src, int64_t n_frames, int64_t n_atoms, int64_t first_atom,
                        int64_t n_window_atoms, void* stream);

/* K1g: gather kernel over a CSR list, used for virtual beads (src = cg_xyz, mapping.py:431-438) and as
 * the path for beads too wide for a tile.  bead_list: device int32[n_list] of output beads to compute. */
int cgb_map_gather(const float* src, int64_t src_atoms, const float* box_len, int64_t n_frames, int64_t n_beads,
                   const int32_t* bead_ptr, const int32_t* index, const void* weights, int weights_f64,
                   const int32_t* bead_list, int64_t n_list, float* cg_xyz, void* stream);

/* Tunables of K1 (0 keeps the built-in default): frames per pipeline stage byte budget, pipeline
 * depth, frames per CTA.  Intended for benchmarking; not needed for correctness. */
int cgb_map_set_tuning(int32_t stage_bytes, int32_t n_stages, int32_t frames_per_cta);
/* More K1 benchmarking knobs: consumer threads per CTA (multiple of 32, <= 256) and frames per register
 * group (1, 2 or 4); 0 keeps the default. */
int cgb_map_set_shape(int32_t consumer_threads, int32_t frame_group);

/* ------------------------------------------------------------------------------------------------
 * Stage 2: bonded-term measurement with streaming statistics.
 *
 * A "column" is one instance of one bonded term (bondset.py:509-528); columns are ordered lengths,
 * then angles, then dihedrals (n2, n3, n4 of them).  col_beads: device int32 [M][4] bead indices
 * (unused trailing entries repeat the last bead); col_bond: device int32 [M] owning Bond.
 * values: device float32, F*M floats stored as one plane per kind, planes in kind order: the plane of a
 * kind with n columns starting at column c0 begins at float F*c0 and is [F][n] frame-major
 * (Bond.values = plane[:, columns of the bond].flatten()); each kernel then writes one contiguous region.
 *   length   |mic(x1 - x0)|                                         (mdtraj compute_distances)
 *   angle    acos(u.v / |u||v|), u = mic(x0 - x1), v = mic(x2 - x1)  (mdtraj compute_angles)
 *   dihedral atan2((b1.c1)|b2|, c1.c2), b_i = mic(x_i - x_{i-1}), c1 = b2 x b3, c2 = b1 x b2
 *                                                                   (mdtraj compute_dihedrals)
 * box_vec: device float32 [F][9] rows a,b,c, or NULL (no unit cell).  ortho != 0 selects the
 * orthorhombic minimum image (box diagonal), else the triclinic one (reduce, subtract, 27 images).
 *
 * partials: device float64 [n_bonds][CGB_NPART], ACCUMULATED into (caller zeroes it once; per-GPU
 * shards are summed with an allreduce):
 *   [0] count of finite values        [1] sum (v - shift)       [2] sum (v - shift)^2
 *   [3] sum cos v                     [4] sum sin v             [5] sum wrap(v - circular mean)^2 (pass 2)
 *   [6] count of all values           [7] reserved
 * shift: device float32 [n_bonds] or NULL (0) -- any value near the data, identical on all shards.
 * hist: device uint64 [n_bonds][n_bins] accumulated, or NULL; bin edges are
 * linspace(hist_lo[b], hist_hi[b], n_bins + 1) with numpy.histogram's edge rules (1 <= n_bins <= 512).
 * ---------------------------------------------------------------------------------------------- */
#define CGB_NPART 8
/* col_slot: device int32 [M] or NULL.  Inside every group of CGB_MEASURE_TILE consecutive columns of one
 * kind (a kind with fewer columns is one group) the distinct Bonds are numbered 0, 1, ... in any order and
 * col_slot[c] is the number of column c's Bond; max_slots is the largest count over all groups.  It lets a
 * CTA keep its per-Bond sums and histograms in shared memory sized by the group, not by the system.
 * NULL (or max_slots too large for shared memory) selects global atomics. */
#define CGB_MEASURE_TILE 224
/* tile_span: device int32 [T][2] or NULL, T = sum over kinds of ceil(n_kind / CGB_MEASURE_TILE), one row per
 * group in kind order: (first bead, number of beads) of the contiguous bead range that contains every bead
 * of the group's columns; max_span = the largest number of beads.  With tile_span (and col_slot) the staged
 * kernels run: the range is streamed through shared memory with bulk copies and gathered with LDS.  NULL, or
 * a range too large for shared memory, selects the kernels that gather from global memory. */

int cgb_measure(const float* cg_xyz, const float* box_vec, int ortho, int64_t n_frames, int64_t n_beads,
                const int32_t* col_beads, const int32_t* col_bond, const int32_t* col_slot, int32_t max_slots,
                const int32_t* tile_span, int32_t max_span, int64_t n2, int64_t n3, int64_t n4,
                int64_t n_bonds, const float* shift, float* values, double* partials,
                unsigned long long* hist, const float* hist_lo, const float* hist_hi, int32_t n_bins,
                void* stream);

/* Benchmarking knobs of the staged K2 kernels: bytes per pipeline stage (1 KiB .. 48 KiB) and pipeline
 * depth (2 .. 4).  Not needed for correctness. */
int cgb_measure_set_tuning(int32_t stage_bytes, int32_t n_stages);

/* Pass 2 for dihedral columns: partials[b][5] += sum over the bond's values of wrap(v - m_b)^2 with
 * m_b = atan2(partials[b][4], partials[b][3]) (util.py:40-53).  Call after partials[.][3..4] are final
 * (i.e. after the cross-GPU reduction).  values: device pointer to the first of n_cols columns of a
 * [F][row_stride] matrix (the dihedral plane: values + F*(n2+n3), row_stride = n4); col_bond: the Bond of
 * each of those columns (col_bond + n2 + n3). */
int cgb_circular_pass2(const float* values, int64_t n_frames, int64_t row_stride,
                       const int32_t* col_bond, int64_t n_cols, double* partials, void* stream);

/* Statistics of a caller-supplied value array for ONE Bond (Bond.boltzmann_invert on values that were
 * not measured by cgb_measure, bondset.py:72-74): accumulates slots [0..4] and [6] of one partials row.
 * values: device float32 [n]; natoms in {2,3,4} selects which sums are meaningful. */
int cgb_value_stats(const float* values, int64_t n, int natoms, float shift, double* partials_row, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3: Boltzmann inversion (one thread per Bond).
 *   natoms[b] in {2,3,4}; form[b] is a CGB_FORM_* id; RT = 8.314 * temperature / 1000.
 *   out: device float64 [n_bonds][4] = { eqm, fconst, mean, variance }
 *   lengths: arithmetic mean / population variance; angles and dihedrals: circular mean / variance;
 *   dihedral eqm = ((mean) mod 2pi) - pi (bondset.py:81-84); variance == 0 -> fconst = +inf
 *   (bondset.py:89-91); a bond with no finite values yields NaN in all four slots.
 * ---------------------------------------------------------------------------------------------- */
enum {
    CGB_FORM_HARMONIC = 0,            /* RT / var                       functionalforms.py:101-104 */
    CGB_FORM_COSHARMONIC = 1,         /* RT / (sin^2(mean) * var)       functionalforms.py:117-121 */
    CGB_FORM_MARTINI_LENGTH = 2,      /* 1250                           functionalforms.py:129-132 */
    CGB_FORM_MARTINI_ANGLE = 3,       /* 25                             functionalforms.py:138-141 */
    CGB_FORM_MARTINI_DIHEDRAL = 4,    /* 50                             functionalforms.py:147-148 */
    CGB_FORM_STATS_ONLY = 5           /* user-defined form: only mean/variance are produced */
};

int cgb_boltzmann(const double* partials, const float* shift, const int32_t* natoms, const int32_t* form,
                  double temperature, int64_t n_bonds, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Trajectory ingest: GROMACS .xtc (the caller of the hot path, reference pycgtool/frame.py:41-63, where the
 * reference calls mdtraj's C codec).  The file image is scanned on the host, the compressed coordinate
 * payloads are decoded on the GPU (a walk per frame, then one thread per 16 coordinate groups) straight into the
 * [F][N][3] float32 trajectory.
 * Coordinates are float32(int) * float32(1 / precision), the xdrfile convention.
 * ---------------------------------------------------------------------------------------------- */
typedef struct CgbXtcFrame {     /* 48 bytes */
    int64_t offset;              /* byte offset of the coordinate payload inside the file image */
    int32_t nbytes;              /* payload bytes (12 * natoms raw big-endian floats when natoms <= 9) */
    int32_t natoms;
    int32_t minint[3];
    int32_t maxint[3];
    int32_t smallidx;
    float precision;
} CgbXtcFrame;

/* Host: walk the frame headers of a file image.  frames == NULL only counts.  box: float32 [F][9] row
 * vectors, time: float32 [F], step: int32 [F] (each may be NULL).  CGB_EINVAL on a malformed image or a
 * changing atom count, CGB_ERANGE when max_frames is too small. */
int cgb_xtc_scan(const uint8_t* data, int64_t n_bytes, int64_t max_frames, CgbXtcFrame* frames,
                 float* box, float* time, int32_t* step, int64_t* n_frames, int32_t* n_atoms);

/* GPU: decode `n_frames` (<= 65535 per call) frames of the file image `data` (device memory, 4-byte aligned,
 * n_bytes long and readable up to 8 bytes past the end) into xyz (device float32 [n_frames][n_atoms][3]).
 * frames: device copy of the table from cgb_xtc_scan (its offsets are relative to `data`).  Two kernels: a
 * sequential walk per frame that records a checkpoint every 16 coordinate groups, then one thread per
 * checkpoint decodes in parallel.  workspace: device scratch of cgb_xtc_workspace_bytes(), 16-byte aligned. */
int64_t cgb_xtc_workspace_bytes(int64_t n_frames, int64_t n_atoms);
int cgb_xtc_decode(const uint8_t* data, int64_t n_bytes, const CgbXtcFrame* frames, int64_t n_frames,
                   int64_t n_atoms, float* xyz, void* workspace, int64_t workspace_bytes, void* stream);

/* Host: the same decoder on the CPU (all pointers host memory). */
int cgb_xtc_decode_host(const uint8_t* data, const CgbXtcFrame* frames, int64_t n_frames, int64_t n_atoms,
                        float* xyz);

/* Host: encode frames as .xtc into out[capacity]; *n_written receives the size of the image.  Returns
 * CGB_ERANGE (with *n_written set) when capacity is too small.  box: [F][9] or NULL, time: [F] or NULL. */
int cgb_xtc_encode(const float* xyz, const float* box, const float* time, int64_t n_frames, int64_t n_atoms,
                   float precision, uint8_t* out, int64_t capacity, int64_t* n_written);

#ifdef __cplusplus
}
#endif
#endif /* PYCGTOOL_B200_H */
