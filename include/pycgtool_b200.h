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
 *     and returns without synchronising; the library allocates nothing and keeps no mutable state (tuning
 *     knobs are per-call arguments), so calls are re-entrant across streams and threads;
 *   - return 0 on success, a positive cudaError_t on a CUDA failure, a negative CGB_E* on bad arguments;
 *   - all index tables are int32; frames, atoms and beads are counted in int64.
 */
#ifndef PYCGTOOL_B200_H
#define PYCGTOOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_ABI_VERSION 3

enum {
    CGB_OK = 0,
    CGB_EINVAL = -1,      /* NULL pointer, negative size, inconsistent table */
    CGB_ERANGE = -2,      /* index outside the atom / bead range */
    CGB_ETOOBIG = -3,     /* a single bead spans more atoms than a shared-memory tile can hold */
    CGB_EALIGN = -4,      /* pointer not aligned to its element type */
    CGB_ENCCL = -5        /* libnccl could not be loaded, or an NCCL call failed */
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

/* The same two passes with forced cuts: tile_start (HOST uint8 [n_beads], or NULL) is 1 where a tile must begin.
 * Used by the fused mapping + measurement path, whose mapping tiles follow the measurement plan's units. */
int cgb_map_plan_size_cuts(const int32_t* bead_ptr, const int32_t* index, const uint8_t* compute, int64_t n_beads,
                           int64_t n_atoms, int32_t max_span_atoms, int32_t max_tile_beads, const uint8_t* tile_start,
                           int64_t* n_tiles, int64_t* n_entries, int32_t* max_span, int32_t* max_tile_entries);
int cgb_map_plan_build_cuts(const int32_t* bead_ptr, const int32_t* index, const void* weights, int weights_f64,
                            const uint8_t* compute, int64_t n_beads, int64_t n_atoms, int32_t max_span_atoms,
                            int32_t max_tile_beads, const uint8_t* tile_start, CgbMapTile* tiles, CgbMapEntry* entries,
                            double* entries_w64);

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

/* Tunables of K1, passed per call (NULL or a zero field keeps the built-in default).  Intended for benchmarking;
 * not needed for correctness.  The library keeps no mutable global state. */
typedef struct CgbMapTuning {
    int32_t stage_bytes;       /* byte budget of one pipeline stage                      */
    int32_t n_stages;          /* pipeline depth (2..4)                                  */
    int32_t frames_per_cta;    /* frames one CTA works through                           */
    int32_t consumer_threads;  /* consumer threads per CTA (multiple of 32, <= 256)      */
    int32_t frame_group;       /* frames per register group (1, 2 or 4)                  */
    int32_t reserved[3];
} CgbMapTuning;

/* K1 with explicit tuning and atom window (cgb_map_tiles / cgb_map_tiles_window call this with tuning = NULL). */
int cgb_map_tiles_ex(const float* xyz, const float* box_len, int64_t n_frames, int64_t n_atoms, int64_t first_atom,
                     int64_t n_beads, const CgbMapTile* tiles, int64_t n_tiles, const CgbMapEntry* entries,
                     const double* entries_w64, int32_t max_span, int32_t max_tile_entries, float* cg_xyz,
                     const CgbMapTuning* tuning, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2: bonded-term measurement with streaming statistics.
 *
 * A "column" is one instance of one bonded term (bondset.py:509-528); the caller numbers its columns lengths,
 * then angles, then dihedrals (n2, n3, n4 of them).  col_beads: int32 [M][4] bead indices (unused trailing
 * entries repeat the last bead); col_bond: int32 [M] owning Bond.
 *   length   |mic(x1 - x0)|                                         (mdtraj compute_distances)
 *   angle    acos(u.v / |u||v|), u = mic(x0 - x1), v = mic(x2 - x1)  (mdtraj compute_angles)
 *   dihedral atan2((b1.c1)|b2|, c1.c2), b_i = mic(x_i - x_{i-1}), c1 = b2 x b3, c2 = b1 x b2
 *                                                                   (mdtraj compute_dihedrals)
 * box_vec: device float32 [F][9] rows a,b,c, or NULL (no unit cell).  ortho != 0 selects the
 * orthorhombic minimum image (box diagonal), else the triclinic one (reduce, subtract, 27 images).
 *
 * values: device float32 [F][row_stride] row-major, or NULL (statistics only: nothing is stored per
 * measure); Bond.values = values[:, columns of the bond].flatten().  The fused kernel places every column
 * at the position the plan compiler chose (col_out); the gather entry writes its column j at position j
 * of the pointer it is given.
 *
 * partials: device float64 [n_bonds][CGB_NPART], ACCUMULATED into (caller zeroes it once; per-GPU
 * shards are summed with an allreduce):
 *   [0] count of finite values        [1] sum d                 [2] sum d^2
 *   [3] sum cos v                     [4] sum sin v             [5] sum wrap(v - circular mean)^2 (pass 2)
 *   [6] count of all values           [7] reserved
 *   with d = v - shift for lengths and angles, and d = wrap(v - shift) into [-pi, pi] for dihedrals.
 * shift: device float32 [n_bonds] or NULL (0) -- a value near the data, identical on all shards
 * (cgb_measure_shift).  The statistics are accumulated in float32 over short runs, so their accuracy
 * relies on |d| being of the order of the spread of the data.
 * hist: device float64 [n_bonds][n_bins] integer-valued counts, accumulated, or NULL; bin edges are
 * linspace(hist_lo[b], hist_hi[b], n_bins + 1) with numpy.histogram's edge rules (1 <= n_bins <= 512).
 * Counts are float64 so that partials and histograms form ONE float64 buffer for the cross-GPU sum.
 * ---------------------------------------------------------------------------------------------- */
#define CGB_NPART 8

/* ---- fused kernel (K2): plan ---- */
#define CGB_MEAS_WARPS 7          /* consumer warps per CTA = groups per tile at most                    */
#define CGB_MEAS_CHUNKS 4         /* 32-lane chunks per group                                            */
#define CGB_MEAS_EDGE_CHUNKS 2    /* of which at most this many hold edges (bead pairs)                  */
/* Shape of every group of a plan: the kind of each of its four chunks, fixed at compile time in the kernel. */
#define CGB_SHAPE_EAD 0           /* [edges | angles | dihedrals | unused]: <= 32 of each                 */
#define CGB_SHAPE_EEAA 1          /* [edges | edges | angles | angles]: <= 64 of each, no dihedrals        */

typedef struct CgbMeasTile {      /* 32 bytes: the work of one CTA */
    int32_t bead0, span;          /* contiguous bead range staged per frame                              */
    int32_t group0, ngroups;      /* its groups                                                          */
    int32_t slot0, nslots;        /* tile_bonds[slot0 .. slot0 + nslots): Bond of every tile-local slot  */
    int32_t list0, nlist;         /* tile_lists[list0 ..): uint16 slot_ptr[nslots + 1], then nlist lane-roles
                                     (chunk * 224 + group-in-tile * 32 + lane) grouped by slot            */
} CgbMeasTile;

typedef struct CgbMeasGroup {     /* 32 bytes: the work of one warp */
    uint8_t kind[CGB_MEAS_CHUNKS];    /* 0 unused, 2 edges, 3 angles, 4 dihedrals; edge chunks come first */
    uint8_t nlanes[CGB_MEAS_CHUNKS];  /* active lanes of the chunk                                        */
    uint8_t nemit[CGB_MEAS_CHUNKS];   /* its first nemit lanes produce a column                           */
    int32_t vcol0[CGB_MEAS_CHUNKS];   /* position in the values row of lane 0's column                    */
    int32_t reserved;
} CgbMeasGroup;

/* One lane of a chunk (8 bytes).  Edge lane: ref = a | b << 16, beads relative to the tile's bead0; its record
 * holds the minimum image of x_b - x_a.  Angle lane: ref = e_u | s_u << 7 | e_v << 8 | s_v << 15, the records
 * (0..63, numbered through the edge chunks) of u = x0 - x1 and v = x2 - x1, s = 1 when the record is reversed.
 * Dihedral lane: three such bytes for b1, b2, b3.  slot: tile-local Bond slot, -1 for an edge without a column. */
typedef struct CgbMeasLane {
    uint32_t ref;
    int32_t slot;
} CgbMeasLane;

typedef struct CgbMeasPlanInfo {
    int64_t n_tiles, n_groups, n_tile_bonds, n_tile_lists;
    int64_t n_wide;               /* columns whose beads are too far apart for a tile (gather kernel)     */
    int64_t n_staged;             /* columns of the fused kernel: positions [0, n_staged) of the values row */
    int32_t max_span, max_slots, max_edge_chunks, max_tile_groups;
} CgbMeasPlanInfo;

/* Limits to compile a plan with, for a histogram of n_bins bins (with_hist = 0: no histogram): want_slots = the
 * Bonds the caller would like one tile to hold (those of its largest molecule); max_slots <= want_slots and
 * max_span, the widest bead range, are what fits the shared memory of a CTA together. */
int cgb_measure_plan_limits(int32_t n_bins, int with_hist, int32_t want_slots, int32_t* max_span, int32_t* max_slots);

/* Host-side plan compiler (all pointers HOST memory).  shape: CGB_SHAPE_EAD, or CGB_SHAPE_EEAA when n4 == 0.
 * Pass 1 sizes the arrays; pass 2 fills them: tiles[n_tiles], groups[n_groups],
 * lanes[n_groups][CGB_MEAS_CHUNKS][32], tile_bonds[n_tile_bonds], tile_lists[n_tile_lists] (uint16),
 * col_out[M] (position of every column in the values row; wide columns follow the staged ones, in kind
 * order) and wide_cols[n_wide] (the columns left to cgb_measure).
 * unit_groups = 0: tiles are packed freely (cgb_measure_fused).  unit_groups > 0: the plan of the fused mapping +
 * measurement kernel (cgb_map_measure): every tile is a run of whole units -- bead ranges no term straddles, e.g.
 * one molecule instance -- of at most max_span beads and unit_groups groups, so tiles are disjoint bead ranges;
 * tile_start (uint8 [n_beads], may be NULL) receives 1 where such a tile begins and right after it ends: the cuts the
 * mapping plan must respect (cgb_map_plan_size/build, tile_start).  CGB_ETOOBIG when a unit does not fit. */
int cgb_measure_plan_size(const int32_t* col_beads, const int32_t* col_bond, int64_t n2, int64_t n3, int64_t n4,
                          int64_t n_beads, int32_t max_span, int32_t max_slots, int32_t shape, int32_t unit_groups,
                          CgbMeasPlanInfo* info);
int cgb_measure_plan_build(const int32_t* col_beads, const int32_t* col_bond, int64_t n2, int64_t n3, int64_t n4,
                           int64_t n_beads, int32_t max_span, int32_t max_slots, int32_t shape, int32_t unit_groups,
                           CgbMeasTile* tiles, CgbMeasGroup* groups, CgbMeasLane* lanes, int32_t* tile_bonds,
                           uint16_t* tile_lists, int32_t* col_out, int64_t* wide_cols, uint8_t* tile_start);

/* Benchmarking knobs of K2, passed per call (NULL or a zero field keeps the default). */
typedef struct CgbMeasureTuning {
    int32_t stage_bytes;   /* bytes per pipeline stage                                                    */
    int32_t n_stages;      /* pipeline depth (2..4)                                                       */
    int32_t max_run;       /* frames a lane accumulates in float32 before its CTA reduces in float64      */
    int32_t ctas_per_sm;   /* persistent CTAs per SM (1 or 2)                                             */
    int32_t reserved[4];
} CgbMeasureTuning;

/* K2: one launch of persistent CTAs for all columns of the plan (device copies of its arrays).  The work items
 * (tile, block of frames) are dealt to the CTAs in equal contiguous shares.
 * fold >= 1, frame folding for systems too small to fill a warp (one small molecule: 23 of 32 lanes): `fold`
 * consecutive real frames are measured as ONE frame of fold * beads -- cg_xyz [F][B][3] is the same memory as
 * [F / fold][fold * B][3] -- with a plan compiled for `fold` copies of the system (bead indices of copy r offset by
 * r * B).  n_frames and n_beads are then those of the folded view, box_vec still holds one box per REAL frame
 * (fold * n_frames rows; copy r of folded frame v is real frame v * fold + r), and a row of `values` holds the
 * fold * M values of its real frames in the plan's column order. */
int cgb_measure_fused(const float* cg_xyz, const float* box_vec, int ortho, int64_t n_frames, int64_t n_beads,
                      const CgbMeasTile* tiles, int64_t n_tiles, const CgbMeasGroup* groups,
                      const CgbMeasLane* lanes, const int32_t* tile_bonds, const uint16_t* tile_lists,
                      int32_t shape, int32_t max_span, int32_t max_slots, int32_t max_edge_chunks,
                      int32_t max_tile_groups, int32_t fold, const float* shift, float* values, int64_t row_stride,
                      double* partials, double* hist, const float* hist_lo, const float* hist_hi, int32_t n_bins,
                      const CgbMeasureTuning* tuning, void* stream);

/* Host only (no CUDA call): the launch configuration cgb_measure_fused would choose on a device with sm_count
 * SMs.  out: int64[8] = {CTAs, shared-memory bytes per CTA, frames per stage, pipeline stages, frame blocks per
 * segment, frame blocks per tile, 1 if the one tile is the whole frame, 0}.  CGB_ETOOBIG when it does not fit. */
int cgb_measure_fused_config(int64_t n_frames, int64_t n_beads, int64_t n_tiles, int32_t max_span, int32_t max_slots,
                             int32_t max_edge_chunks, int32_t max_tile_groups, int32_t n_bins, int with_hist,
                             const CgbMeasureTuning* tuning, int32_t sm_count, int64_t* out);

/* ---- stages 1 + 2 fused: mapping warps hand every mapped tile, still in shared memory, to measurement warps, so
 * the CG trajectory is written once and never read back (reference call sequence: pycgtool/__main__.py:69-112).
 * Mapping plan: cgb_map_plan_build_cuts with the tile_start of a measurement plan compiled with unit_groups = 2,
 * max_span = max_tile_beads; meas_of_tile: device int32 [n_tiles], the measurement tile that covers exactly the bead
 * range of mapping tile i, or -1 (solvent).  float32 weights only; every measured bead must be computed by a tile.
 * box_len: device float32 [F][3] or NULL (mapping, mapping.py:406-408); box_vec / ortho as in cgb_measure_fused.
 * cg_xyz is written for all tiles; values / partials / hist as in cgb_measure_fused. */
int cgb_map_measure(const float* xyz, const float* box_len, const float* box_vec, int ortho, int64_t n_frames,
                    int64_t n_atoms, int64_t n_beads, const CgbMapTile* tiles, int64_t n_tiles,
                    const CgbMapEntry* entries, int32_t max_span, int32_t max_tile_entries, int32_t max_tile_beads,
                    float* cg_xyz, const int32_t* meas_of_tile, const CgbMeasTile* meas_tiles,
                    const CgbMeasGroup* groups, const CgbMeasLane* lanes, const int32_t* tile_bonds,
                    const uint16_t* tile_lists, int32_t shape, int32_t max_slots, const float* shift, float* values,
                    int64_t row_stride, double* partials, double* hist, const float* hist_lo, const float* hist_hi,
                    int32_t n_bins, void* stream);

/* ---- gather kernels: columns straight from global memory (wide columns, or callers without a plan) ----
 * col_slot: device int32 [M] or NULL.  Inside every group of CGB_MEASURE_TILE consecutive columns of one
 * kind (a kind with fewer columns is one group) the distinct Bonds are numbered 0, 1, ... in any order and
 * col_slot[c] is the number of column c's Bond; max_slots is the largest count over all groups.  It lets a
 * CTA keep its per-Bond sums and histograms in shared memory.  NULL selects global atomics. */
#define CGB_MEASURE_TILE 224
int cgb_measure(const float* cg_xyz, const float* box_vec, int ortho, int64_t n_frames, int64_t n_beads,
                const int32_t* col_beads, const int32_t* col_bond, const int32_t* col_slot, int32_t max_slots,
                int64_t n2, int64_t n3, int64_t n4, const float* shift, float* values, int64_t row_stride,
                double* partials, double* hist, const float* hist_lo, const float* hist_hi, int32_t n_bins,
                void* stream);

/* shift[b] = value of Bond b's first instance in frame `frame` (bond_beads: device int32 [n_bonds][4], natoms:
 * device int32 [n_bonds], 0 for a Bond without instances), 0 when it is not finite.  Evaluated with the routines
 * of the measurement kernels, so samples identical to it give d == 0 exactly. */
int cgb_measure_shift(const float* cg_xyz, const float* box_vec, int ortho, int64_t frame, int64_t n_frames,
                      int64_t n_beads, const int32_t* bond_beads, const int32_t* natoms, int64_t n_bonds,
                      float* shift, void* stream);

/* Pass 2 for dihedral columns: partials[b][5] += sum over the bond's values of wrap(v - m_b)^2 with
 * m_b = atan2(partials[b][4], partials[b][3]) (util.py:40-53), for the Bonds with flags[b * flag_stride] != 0
 * (all Bonds when flags == NULL).  Call after partials[.][3..4] are final (i.e. after the cross-GPU
 * reduction).  values: device pointer to the first of n_cols columns of a [F][row_stride] matrix; col_bond:
 * the Bond of each of those columns. */
int cgb_circular_pass2(const float* values, int64_t n_frames, int64_t row_stride, const int32_t* col_bond,
                       int64_t n_cols, double* partials, const double* flags, int64_t flag_stride, void* stream);

/* Statistics of a caller-supplied value array for ONE Bond (Bond.boltzmann_invert on values that were
 * not measured by cgb_measure, bondset.py:72-74): accumulates slots [0..4] and [6] of one partials row.
 * values: device float32 [n]; natoms in {2,3,4} selects which sums are meaningful. */
int cgb_value_stats(const float* values, int64_t n, int natoms, float shift, double* partials_row, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3: Boltzmann inversion (one thread per Bond).
 *   natoms[b] in {2,3,4}; form[b] is a CGB_FORM_* id; RT = 8.314 * temperature / 1000.
 *   out: device float64 [n_bonds][CGB_NOUT] = { eqm, fconst, mean, variance, flag }
 *   lengths: arithmetic mean / population variance; angles and dihedrals: circular mean / variance;
 *   dihedral eqm = ((mean) mod 2pi) - pi (bondset.py:81-84); variance == 0 -> fconst = +inf
 *   (bondset.py:89-91); a bond with no finite values yields NaN in the first four slots.
 *
 * Dihedral variance in one pass: with d = wrap(v - shift) and e = wrap(mean - shift),
 * <wrap(v - mean)^2> = <d^2> - 2 e <d> + e^2 provided no sample lies on the arc between the antipodes of the
 * shift and of the mean.  The histogram of the Bond (hist / hist_lo / hist_hi / n_bins as in stage 2; it must
 * hold every finite sample) proves that: the bins that touch the arc, widened by one bin, must be empty.
 * Otherwise -- or without a histogram, unless e == 0 -- flag = 1 and the caller runs cgb_circular_pass2 (or
 * measures again with shift = mean) and calls cgb_boltzmann with pass2 != 0.
 *   pass2 = 0: one-pass variance, flag output as above;
 *   pass2 = 1: dihedral Bonds whose out[b][4] is non-zero ON ENTRY take partials[b][5] / n (flag becomes 2);
 *   pass2 = 2: every dihedral Bond takes partials[b][5] / n.
 * ---------------------------------------------------------------------------------------------- */
#define CGB_NOUT 5
enum {
    CGB_FORM_HARMONIC = 0,            /* RT / var                       functionalforms.py:101-104 */
    CGB_FORM_COSHARMONIC = 1,         /* RT / (sin^2(mean) * var)       functionalforms.py:117-121 */
    CGB_FORM_MARTINI_LENGTH = 2,      /* 1250                           functionalforms.py:129-132 */
    CGB_FORM_MARTINI_ANGLE = 3,       /* 25                             functionalforms.py:138-141 */
    CGB_FORM_MARTINI_DIHEDRAL = 4,    /* 50                             functionalforms.py:147-148 */
    CGB_FORM_STATS_ONLY = 5           /* user-defined form: only mean/variance are produced */
};

int cgb_boltzmann(const double* partials, const float* shift, const int32_t* natoms, const int32_t* form,
                  double temperature, int64_t n_bonds, const double* hist, const float* hist_lo,
                  const float* hist_hi, int32_t n_bins, int pass2, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-GPU exchange (one process per GPU, frames sharded by rank): the only state that crosses GPUs is the
 * packed float64 buffer [partials | hist].  These entry points wrap NCCL (libnccl.so.2, resolved at run time
 * with dlopen so that the library loads without it); the reference has no counterpart (single process).
 *   cgb_comm_unique_id : rank 0 creates the 128-byte id, the caller distributes it to the other ranks;
 *   cgb_comm_init      : collective over all ranks, on the current CUDA device;
 *   cgb_allreduce_partials : ncclAllReduce(sum, float64) in place on `stream`, no host synchronisation;
 *   cgb_comm_broadcast : ncclBroadcast of raw bytes from `root` (the shift of the statistics).
 * ---------------------------------------------------------------------------------------------- */
#define CGB_COMM_ID_BYTES 128
int cgb_comm_unique_id(uint8_t* id);
int cgb_comm_init(const uint8_t* id, int32_t rank, int32_t world_size, void** comm);
int cgb_comm_destroy(void* comm);
int cgb_allreduce_partials(void* comm, double* packed, int64_t count, void* stream);
int cgb_comm_broadcast(void* comm, void* buffer, int64_t n_bytes, int32_t root, void* stream);

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
/* The decoder does not trust the bit stream: a frame whose stream leaves the small-difference table or runs past
 * its payload (truncated / corrupt file) is not decoded and counted in *bad_frames (device int32, accumulated;
 * may be NULL); no byte outside the frame's payload is read and none outside its rows of xyz written. */
int cgb_xtc_decode(const uint8_t* data, int64_t n_bytes, const CgbXtcFrame* frames, int64_t n_frames,
                   int64_t n_atoms, float* xyz, void* workspace, int64_t workspace_bytes, int32_t* bad_frames,
                   void* stream);

/* Host: the same decoder on the CPU (all pointers host memory).  CGB_EINVAL when a frame's stream is invalid. */
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
