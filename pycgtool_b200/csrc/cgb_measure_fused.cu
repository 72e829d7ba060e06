// Stage 2 -- bonded-term measurement with streaming statistics, fused kernel (K2).
//
// Reference: pycgtool/bondset.py:496-531 hands (F,B,3) bead coordinates and one index table per Bond to
// mdtraj.compute_distances / compute_angles / compute_dihedrals (mdtraj 1.9.7, periodic=True) and keeps the
// flattened float32 result; pycgtool/util.py:24-53 then makes ~15 more passes over those values for the
// (circular) mean and variance.  Here ONE launch covers all Bonds of all kinds: the CG frame is read from HBM
// once, every minimum-image displacement is evaluated once, and the same pass stores the values (optional) and
// accumulates count, sum and sum of squares about a shift, sum of cos / sin, and a fixed-bin histogram.
//
// Work decomposition (plan: cgb_measure_plan.cpp.inc):
//   * a CTA owns one tile -- a contiguous bead range of the CG frame -- and a share of the frames.  A producer
//     warp streams the range of k frames at a time into shared memory with 1-D bulk asynchronous copies (TMA,
//     SASS UBLKCP) through an mbarrier ring and prepares the frames' box table;
//   * a consumer warp owns one group of the tile: <= 64 bead pairs ("edges") and the angles / dihedrals built from
//     them.  It walks the stage four frames at a time, as two packed float32x2 frame pairs.  Edge lanes gather two
//     beads with LDS, apply the minimum image, leave the displacement and its squared norm in a per-warp
//     shared-memory record and finish the lengths; after a __syncwarp the angle and dihedral lanes read two or
//     three records each.  Tiles with fewer groups than warps spread the frames of a stage over the spare warps;
//   * statistics stay in packed float32 registers for the (bounded) life of the CTA and are reduced per Bond in
//     a fixed order in float64 at its end -- no floating-point atomics inside a CTA.

#include <algorithm>
#include <cstring>

#include "cgb_terms.cuh"
#include "cgb_measure_plan.cpp.inc"

namespace cgb {

constexpr int kWarps = CGB_MEAS_WARPS;            // consumer warps per CTA
constexpr int kConsumers = kWarps * 32;
constexpr int kChunks = CGB_MEAS_CHUNKS;
constexpr int kRoles = kChunks * kConsumers;      // lane-roles of a CTA (chunk, warp, lane)
constexpr int kStagesMax = 4;

struct FusedParams {
    const float* xyz;
    const float* box;
    int64_t F, B;
    const CgbMeasTile* tiles;
    const CgbMeasGroup* groups;
    const CgbMeasLane* lanes;
    const int32_t* tile_bonds;
    const uint16_t* tile_lists;
    const float* shift;
    const float* hlo;
    const float* hhi;
    float* values;
    int64_t row_stride;
    double* partials;
    double* hist;
    int64_t n_tiles;
    int64_t nblocks;       // frame blocks of k frames per tile
    int32_t seg_cap;       // blocks a CTA works through before it reduces its float32 sums
    int32_t nbins;
    int32_t max_slots;
    int32_t k;             // frames per stage (multiple of 4)
    int32_t nstage;
    int32_t slot_bytes;    // bytes between the staged frames of a stage (== 12 B for a whole-frame tile)
    int32_t tab_off;       // byte offset of the box table inside a stage
    int32_t stage_bytes;
    int32_t contiguous;    // the tile is the whole CG frame: one bulk copy per stage
    int32_t ecache_warp_bytes;
    // shared-memory map (byte offsets)
    int32_t off_sparam, off_shist, off_lists, off_eacc, off_nbad, off_ecache, off_stages;
};

// Kind of chunk c in a group of the given shape (0: unused).
template <int SHAPE>
__host__ __device__ constexpr int chunk_kind(int c) {
    return SHAPE == CGB_SHAPE_EAD ? (c == 0 ? 2 : (c == 1 ? 3 : (c == 2 ? 4 : 0))) : (c < 2 ? 2 : 3);
}

// Edge records of a warp: [frame pair][half of the record][edge][16 bytes].  Strides by shape (compile time).
template <int SHAPE>
struct RecLayout {
    static constexpr int kEdges = SHAPE == CGB_SHAPE_EAD ? 32 : 64;
    static constexpr int kHalf = kEdges * 16;       // bytes between the two halves of a record
    static constexpr int kPair = 2 * kHalf;         // bytes between the records of the two frame pairs
    static constexpr int kWarpBytes = 2 * kPair;
};

// Per-lane state of one chunk of the warp's group.
struct Role {
    uint32_t a0, a1, a2;   // edge: byte offsets of the two beads in a staged frame; angle / dihedral: record offsets
    float g1, g2;          // orientation signs
    f2 s1, s2, c, s;       // float32 sums of this CTA segment: d, d^2, cos, sin (two frames per register pair)
    int slot;              // tile-local Bond slot, -1: the lane produces no column
};

// The share of the work items (tile-major, frame-block-minor) of one CTA, cut into segments: consecutive blocks of
// one tile, at most `cap` of them.  Producer and consumers walk the same sequence.
struct Schedule {
    int64_t w, w_end, nblocks;
    int cap;
    __device__ Schedule(const FusedParams& q) {
        const int64_t total = q.n_tiles * q.nblocks;
        w = total * (int64_t)blockIdx.x / (int64_t)gridDim.x;
        w_end = total * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
        nblocks = q.nblocks;
        cap = q.seg_cap;
    }
    __device__ bool next(int& tile, int64_t& block0, int& count) {
        if (w >= w_end) return false;
        tile = (int)(w / nblocks);
        block0 = w - (int64_t)tile * nblocks;
        count = (int)min(min(w_end - w, nblocks - block0), (int64_t)cap);
        w += count;
        return true;
    }
};

template <int MIC, int SHAPE>
__global__ void __launch_bounds__(kConsumers + 32, 2) measure_fused_kernel(const FusedParams q) {
    static_assert(MIC == MIC_ORTHO || MIC == MIC_TRICLINIC, "no box = orthorhombic path with a zero box table");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kStagesMax;
    int* sbond = reinterpret_cast<int*>(smem + 128);
    // per Bond slot of the tile: {-shift, bins / (hi - lo), -lo * that - 0.5, unused} for the statistics / histogram
    float4* sparam = reinterpret_cast<float4*>(smem + q.off_sparam);
    unsigned int* shist = reinterpret_cast<unsigned int*>(smem + q.off_shist);
    uint16_t* slist = reinterpret_cast<uint16_t*>(smem + q.off_lists);     // slot_ptr[nslots + 1], then lane-roles
    // reduction scratch of the segment epilogue, over the (then idle) edge records: [4][2 chunks x 224] sums + counts
    float* eacc = reinterpret_cast<float*>(smem + q.off_ecache);
    int* ecnt = reinterpret_cast<int*>(eacc + 4 * 2 * kConsumers);
    unsigned int* snbad = reinterpret_cast<unsigned int*>(smem + q.off_nbad);  // [kRoles] NaN values (rare path)
    unsigned char* stages = smem + q.off_stages;
    const bool do_hist = q.hist != nullptr;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (tid == 0) {
        for (int s = 0; s < q.nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);      // every consumer warp arrives, working or not
        }
        fence_barrier_init();
    }
    __syncthreads();

    const size_t frame_fl = (size_t)q.B * 3;
    Schedule sched(q);
    int seg_tile, seg_count;
    int64_t seg_block0;

    if (warp == kWarps) {
        // ------------------------------------------------------------------ producer warp
        const uintptr_t lo = reinterpret_cast<uintptr_t>(q.xyz);
        const uintptr_t hi = lo + (uintptr_t)q.F * (uintptr_t)q.B * 12u;
        const uint64_t policy = policy_evict_first();
        int s = 0, filled = 0;
        uint32_t phase = 1;  // parity of the previous use of the slot: the first pass through the ring never waits
        while (sched.next(seg_tile, seg_block0, seg_count)) {
            const int bead0 = __ldg(&q.tiles[seg_tile].bead0), span = __ldg(&q.tiles[seg_tile].span);
            const float* const span0 = q.xyz + (size_t)bead0 * 3;     // the tile's range in frame 0
            // every stage starts at a frame that is a multiple of 4, so its 16-byte head is the same for all stages
            const uint32_t h0 = (uint32_t)(reinterpret_cast<uintptr_t>(span0) & 15u);
            for (int it = 0; it < seg_count; ++it) {
                const int64_t f0 = (seg_block0 + it) * q.k;
                if (filled >= q.nstage) mbar_wait(&empty[s], phase);
                ++filled;
                const int s_now = s;
                if (++s == q.nstage) {
                    s = 0;
                    phase ^= 1u;
                }
                const int kk = (int)min((int64_t)q.k, q.F - f0);
                unsigned char* sbase = stages + (size_t)s_now * q.stage_bytes;
                uint64_t* const full_s = &full[s_now];
                const int ncopies = q.contiguous ? 1 : kk;
                const uint32_t nfl = q.contiguous ? (uint32_t)kk * (uint32_t)frame_fl : (uint32_t)(3 * span);
                for (int j0 = 0; j0 < ncopies; j0 += 32) {
                    const int j = j0 + lane;
                    uint32_t bytes = 0;
                    bool edge = false;
                    const float* src = nullptr;
                    if (j < ncopies) {
                        src = span0 + (size_t)(f0 + j) * frame_fl;
                        bytes = span_bytes(src, nfl, lo, hi, &edge);
                    }
                    // bulk copies first (their latency is the long pole), then the hand-copied edge runs
                    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
                    if (lane == 0 && total) mbar_expect_tx(full_s, total);
                    __syncwarp();
                    // float i of frame j lands at sbase + j * slot_bytes + h0 + 4 i: slot_bytes == 12 B (mod 16), so
                    // this address has the low four bits of its source and the copy runs between 16-byte boundaries
                    unsigned char* dst = sbase + (size_t)j * q.slot_bytes + h0;
                    if (j < ncopies && !edge) {
                        const void* a0 = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15));
                        void* d0 = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(dst) & ~uintptr_t(15));
                        bulk_g2s_stream(d0, a0, bytes, full_s, policy);
                    }
                    unsigned edges = __ballot_sync(0xffffffffu, edge);
                    while (edges) {
                        const int e = __ffs(edges) - 1;
                        edges &= edges - 1;
                        const float* esrc = reinterpret_cast<const float*>(
                            __shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(src), e));
                        float* edst = reinterpret_cast<float*>(sbase + (size_t)(j0 + e) * q.slot_bytes + h0);
                        for (uint32_t i = lane; i < nfl; i += 32) edst[i] = __ldg(esrc + i);
                    }
                }
                {
                    // the frames' boxes, ready to use.  Orthorhombic (and no box: all zeros, which leaves every
                    // displacement as it is): one 12-float block per pair of adjacent frames,
                    // [-Lx_a -Lx_b -Ly_a -Ly_b | -Lz_a -Lz_b ix_a ix_b | iy_a iy_b iz_a iz_b] (i = 1 / L), so that three
                    // 16-byte loads give the six packed operands.  Triclinic: 12 floats per frame, [a 0 | b 0 | c 0] reduced.
                    float* tab = reinterpret_cast<float*>(sbase + q.tab_off);
                    for (int j = lane; j < kk; j += 32) {
                        Box bx;
                        if (q.box) {
                            RawBox<MIC> rb;
                            fetch_box<MIC>(q.box, f0 + j, rb);
                            make_box<MIC>(rb, bx);
                        } else {
#pragma unroll
                            for (int d = 0; d < 3; ++d) bx.nL[d] = bx.inv[d] = 0.f;
                        }
                        if (MIC == MIC_ORTHO) {
                            float* t = tab + (j >> 1) * 12 + (j & 1);
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                t[2 * d] = bx.nL[d];
                                t[6 + 2 * d] = bx.inv[d];
                            }
                        } else {
                            float* t = tab + j * 12;
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                t[d] = bx.a[d];
                                t[4 + d] = bx.b[d];
                                t[8 + d] = bx.c[d];
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(full_s);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    using Rec = RecLayout<SHAPE>;
    const uint32_t ecache32 = smem_u32(smem + q.off_ecache) + (uint32_t)warp * (uint32_t)Rec::kWarpBytes;
    const uint32_t shist32 = smem_u32(shist);
    const bool store = q.values != nullptr;  // uniform
    const int nbins = q.nbins;
    const int64_t row_stride = q.row_stride;
    int s = 0;
    uint32_t phase = 0;

    while (sched.next(seg_tile, seg_block0, seg_count)) {
        // ---- segment prologue: the tile's tables and this lane's roles
        asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");   // the previous segment's epilogue is over
        const CgbMeasTile tile = q.tiles[seg_tile];
        const int ng = tile.ngroups;
        const int fl = kWarps / ng;                  // warps that share a group (frame lanes)
        const bool working = warp < ng * fl;
        const int gi = working ? warp % ng : 0;
        const int fli = warp / ng;                   // this warp's frame lane
        const CgbMeasGroup grp = q.groups[tile.group0 + gi];
        const uint32_t h0 = (uint32_t)(reinterpret_cast<uintptr_t>(q.xyz + (size_t)tile.bead0 * 3) & 15u);
        for (int i = tid; i < tile.nslots; i += kConsumers) {
            const int bond = __ldg(q.tile_bonds + tile.slot0 + i);
            sbond[i] = bond;
            float4 prm = make_float4(q.shift ? -__ldg(q.shift + bond) : 0.f, 0.f, 0.f, 0.f);
            if (do_hist) {
                const float lo = __ldg(q.hlo + bond), hi = __ldg(q.hhi + bond);
                prm.y = (float)nbins / (hi - lo);
                prm.z = fmaf(-lo, prm.y, -0.5f);
            }
            sparam[i] = prm;
        }
        if (do_hist)
            for (int i = tid; i < tile.nslots * nbins; i += kConsumers) shist[i] = 0u;
        for (int i = tid; i < tile.nslots + 1 + tile.nlist; i += kConsumers) slist[i] = __ldg(q.tile_lists + tile.list0 + i);
        for (int i = tid; i < kRoles; i += kConsumers) snbad[i] = 0u;

        Role role[kChunks];
        int nl[kChunks], ne[kChunks];                // active / emitting lanes of each chunk (uniform)
        int vcol[kChunks];                           // this lane's position in the values row, per chunk
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            Role& r = role[c];
            r.a0 = r.a1 = r.a2 = 0;
            r.g1 = r.g2 = 1.f;
            r.s1 = r.s2 = r.c = r.s = 0;
            r.slot = -1;
            constexpr int kNone = 0;
            const int kind = chunk_kind<SHAPE>(c);
            nl[c] = (working && kind != kNone) ? grp.nlanes[c] : 0;
            ne[c] = (working && kind != kNone) ? grp.nemit[c] : 0;
            vcol[c] = grp.vcol0[c] + lane;
            if (lane >= nl[c]) continue;
            const CgbMeasLane ln = q.lanes[((size_t)(tile.group0 + gi) * kChunks + c) * 32 + lane];
            if (kind == 2) {
                r.a0 = (ln.ref & 0xffffu) * 12u + h0;
                r.a1 = (ln.ref >> 16) * 12u + h0;
                r.a2 = ecache32 + (uint32_t)(c * 32 + lane) * 16u;      // where this lane leaves its records
            } else {
                const uint32_t e0 = ln.ref & 0x7fu, e1 = (ln.ref >> 8) & 0x7fu, e2 = (ln.ref >> 16) & 0x7fu;
                const float s0 = (ln.ref & 0x80u) ? -1.f : 1.f, s1 = (ln.ref & 0x8000u) ? -1.f : 1.f,
                            s2 = (ln.ref & 0x800000u) ? -1.f : 1.f;
                r.a0 = ecache32 + e0 * 16u;                            // shared-memory addresses of the records
                r.a1 = ecache32 + e1 * 16u;
                r.a2 = ecache32 + e2 * 16u;
                if (kind == 3) {
                    r.g1 = s0 * s1;
                } else {
                    r.g1 = s0 * s1 * s2;
                    r.g2 = s0 * s2;
                }
            }
            r.slot = ln.slot;
        }
        int frames_done = 0;                     // frames this warp has evaluated in this segment (uniform)
        asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");

        // Exact-rule histogram update of one value (next to a bin edge, out of range or NaN: ~0.02 % of the values).
        auto hist_exact = [&](int slot, float v) {
            const int bond = sbond[slot];
            hist_add_exact(shist + slot * nbins, v, __ldg(q.hlo + bond), __ldg(q.hhi + bond), nbins);
        };
        // Statistics, histogram and stores of the two evaluated frame pairs of a quad (the common case: all four
        // frames valid).  Straight-line code up to the one rare branch at its end, so that the two pairs' dependency
        // chains -- and the evaluation before them -- interleave.
        auto finish_fast = [&](Role& r, int kind, f2 va, f2 cva, f2 sva, f2 vb, f2 cvb, f2 svb, float* out) {
            const float4 prm = sparam[r.slot];
            const f2 nshift2 = splat2(prm.x);
            const f2 da = kind == 4 ? wrap_about2(va, nshift2) : add2(va, nshift2);
            const f2 db = kind == 4 ? wrap_about2(vb, nshift2) : add2(vb, nshift2);
            r.s1 = add2(r.s1, add2(da, db));
            r.s2 = fma2(db, db, fma2(da, da, r.s2));
            if (kind != 2) {
                r.c = add2(r.c, add2(cva, cvb));
                r.s = add2(r.s, add2(sva, svb));
            }
            if (store) {
                __stcs(out, lo2(va));
                __stcs(out + row_stride, hi2(va));
                __stcs(out + 2 * row_stride, lo2(vb));
                __stcs(out + 3 * row_stride, hi2(vb));
            }
            if (do_hist) {
                const f2 hn = splat2(prm.y), hc = splat2(prm.z);
                const uint32_t row_adj = shist32 + (uint32_t)(r.slot * nbins) * 4u - 4u * 0x4b000000u;
                const unsigned redo = hist_add_pair_pred(row_adj, va, hn, hc, (unsigned)nbins) |
                                      (hist_add_pair_pred(row_adj, vb, hn, hc, (unsigned)nbins) << 2);
                if (redo) {
                    if (redo & 1u) hist_exact(r.slot, lo2(va));
                    if (redo & 2u) hist_exact(r.slot, hi2(va));
                    if (redo & 4u) hist_exact(r.slot, lo2(vb));
                    if (redo & 8u) hist_exact(r.slot, hi2(vb));
                }
            }
        };
        // One value on its own (rare): a frame next to the end of the trajectory or a degenerate / NaN operand.
        auto finish_one = [&](Role& r, int kind, int c, int half, float v, float cv, float sv, float* out) {
            if (store) __stcs(out, v);
            if (!(v == v)) {  // NaN-skipping statistics (np.nanmean / np.nanvar, util.py:36,53)
                snbad[c * kConsumers + tid] += 1u;
                return;
            }
            const float4 prm = sparam[r.slot];
            const f2 vv = half ? pack2(0.f, v) : pack2(v, 0.f);
            const f2 ns = half ? pack2(0.f, prm.x) : pack2(prm.x, 0.f);
            const f2 d = kind == 4 ? wrap_about2(vv, ns) : add2(vv, ns);
            r.s1 = add2(r.s1, d);
            r.s2 = fma2(d, d, r.s2);
            if (kind != 2) {
                r.c = add2(r.c, half ? pack2(0.f, cv) : pack2(cv, 0.f));
                r.s = add2(r.s, half ? pack2(0.f, sv) : pack2(sv, 0.f));
            }
            if (do_hist) {
                const int bond = sbond[r.slot];
                const int bin = hist_bin(v, __ldg(q.hlo + bond), __ldg(q.hhi + bond), prm.y, nbins);
                if (bin >= 0) atomicAdd(shist + r.slot * nbins + bin, 1u);
            }
        };
        auto pick = [](int i, f2 a, f2 b) {  // frame i (0..3) of the two pairs of a quad
            const f2 x = i < 2 ? a : b;
            return (i & 1) ? hi2(x) : lo2(x);
        };

        // One quad (frames 4 qd .. 4 qd + 3 of the stage, `nvalid` of them inside the trajectory) of this warp's group.
        auto do_quad = [&](const unsigned char* fq, const float* tq, float* vq, int nvalid) {
            const bool whole = nvalid == 4;
            // ---- edges: displacement records of both frame pairs, and the lengths
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                if (chunk_kind<SHAPE>(c) != 2) continue;
                Role& r = role[c];
                if (lane < nl[c]) {
                    EdgeRec e[2];
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        const unsigned char* fa = fq + (size_t)(2 * pp) * q.slot_bytes;
                        const unsigned char* fb = fa + q.slot_bytes;
                        if (MIC == MIC_TRICLINIC) {
                            float rr[2][3];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                Box bx;
                                const float4* t = reinterpret_cast<const float4*>(tq + (2 * pp + h) * 12);
                                const float4 t0 = t[0], t1 = t[1], t2 = t[2];
                                bx.a[0] = t0.x; bx.a[1] = t0.y; bx.a[2] = t0.z;
                                bx.b[0] = t1.x; bx.b[1] = t1.y; bx.b[2] = t1.z;
                                bx.c[0] = t2.x; bx.c[1] = t2.y; bx.c[2] = t2.z;
                                const unsigned char* fr = h ? fb : fa;
                                const float* xa = reinterpret_cast<const float*>(fr + r.a0);
                                const float* xb = reinterpret_cast<const float*>(fr + r.a1);
                                const float pa[3] = {xa[0], xa[1], xa[2]}, pb[3] = {xb[0], xb[1], xb[2]};
                                edge_triclinic(pa, pb, bx, rr[h]);
                            }
#pragma unroll
                            for (int d = 0; d < 3; ++d) e[pp].r[d] = pack2(rr[0][d], rr[1][d]);
                            edge_norm2(e[pp]);
                        } else {
                            const float* xa0 = reinterpret_cast<const float*>(fa + r.a0);
                            const float* xa1 = reinterpret_cast<const float*>(fb + r.a0);
                            const float* xb0 = reinterpret_cast<const float*>(fa + r.a1);
                            const float* xb1 = reinterpret_cast<const float*>(fb + r.a1);
                            f2 xa[3], xb[3];
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                xa[d] = pack2(xa0[d], xa1[d]);
                                xb[d] = pack2(xb0[d], xb1[d]);
                            }
                            const ulonglong2* t = reinterpret_cast<const ulonglong2*>(tq + pp * 12);
                            const ulonglong2 q0 = t[0], q1 = t[1], q2 = t[2];
                            const f2 nL[3] = {q0.x, q0.y, q1.x}, inv[3] = {q1.y, q2.x, q2.y};
                            edge_pair<MIC_ORTHO>(xa, xb, nL, inv, e[pp]);
                        }
                    }
                    // the records in two 16-byte halves, each in its own array: conflict-free 128-bit accesses
                    sts_v2u64<0>(r.a2, e[0].r[0], e[0].r[1]);
                    sts_v2u64<Rec::kHalf>(r.a2, e[0].r[2], e[0].n2);
                    sts_v2u64<Rec::kPair>(r.a2, e[1].r[0], e[1].r[1]);
                    sts_v2u64<Rec::kPair + Rec::kHalf>(r.a2, e[1].r[2], e[1].n2);
                    if (lane < ne[c]) {
                        float* out = vq + vcol[c];
                        f2 va, vb;
                        const bool oka = length_pair(e[0].n2, va), okb = length_pair(e[1].n2, vb);
                        if (oka && okb && whole) {
                            finish_fast(r, 2, va, 0, 0, vb, 0, 0, out);
                        } else {
#pragma unroll 1
                            for (int i = 0; i < nvalid; ++i) {
                                const bool ok = i < 2 ? oka : okb;
                                const float v = ok ? pick(i, va, vb) : __fsqrt_rn(pick(i, e[0].n2, e[1].n2));
                                finish_one(r, 2, c, i & 1, v, 0.f, 0.f, out + (size_t)i * row_stride);
                            }
                        }
                    }
                }
            }
            __syncwarp();

            // ---- angles and dihedrals from the records
            auto load_rec = [&](int pp, uint32_t addr, EdgeRec& e) {
                if (pp == 0) {
                    lds_v2u64<0>(addr, e.r[0], e.r[1]);
                    lds_v2u64<Rec::kHalf>(addr, e.r[2], e.n2);
                } else {
                    lds_v2u64<Rec::kPair>(addr, e.r[0], e.r[1]);
                    lds_v2u64<Rec::kPair + Rec::kHalf>(addr, e.r[2], e.n2);
                }
            };
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                constexpr int kNone = 0;
                const int kind = chunk_kind<SHAPE>(c);
                if (kind == kNone || kind == 2) continue;
                Role& r = role[c];
                if (lane < nl[c]) {
                    float* out = vq + vcol[c];
                    if (kind == 3) {
                        f2 v2[2], cv2[2], sv2[2], den[2];
                        bool ok[2];
#pragma unroll
                        for (int pp = 0; pp < 2; ++pp) {
                            EdgeRec u, v;
                            load_rec(pp, r.a0, u);
                            load_rec(pp, r.a1, v);
                            ok[pp] = angle_pair(u, v, r.g1, v2[pp], cv2[pp], sv2[pp], den[pp]);
                        }
                        if (ok[0] && ok[1] && whole) {
                            finish_fast(r, 3, v2[0], cv2[0], sv2[0], v2[1], cv2[1], sv2[1], out);
                        } else {
#pragma unroll 1
                            for (int i = 0; i < nvalid; ++i) {
                                const float dn = pick(i, den[0], den[1]);
                                const bool good = dn > 0.f && dn < 3e38f;
                                const float val = good ? pick(i, v2[0], v2[1]) : __int_as_float(0x7fc00000);
                                finish_one(r, 3, c, i & 1, val, pick(i, cv2[0], cv2[1]), pick(i, sv2[0], sv2[1]),
                                           out + (size_t)i * row_stride);
                            }
                        }
                    } else {
                        f2 v2[2], cv2[2], sv2[2], p1[2], p2[2];
                        bool ok[2];
#pragma unroll
                        for (int pp = 0; pp < 2; ++pp) {
                            EdgeRec b1, b2, b3;
                            load_rec(pp, r.a0, b1);
                            load_rec(pp, r.a1, b2);
                            load_rec(pp, r.a2, b3);
                            ok[pp] = dihedral_pair(b1, b2, b3, r.g1, r.g2, v2[pp], cv2[pp], sv2[pp], p1[pp], p2[pp]);
                        }
                        if (ok[0] && ok[1] && whole) {
                            finish_fast(r, 4, v2[0], cv2[0], sv2[0], v2[1], cv2[1], sv2[1], out);
                        } else {
#pragma unroll 1
                            for (int i = 0; i < nvalid; ++i) {
                                const float x1 = pick(i, p1[0], p1[1]), x2 = pick(i, p2[0], p2[1]);
                                const float rr = fmaf(x1, x1, x2 * x2);
                                float cv = pick(i, cv2[0], cv2[1]), sv = pick(i, sv2[0], sv2[1]);
                                float val = pick(i, v2[0], v2[1]);
                                if (!(rr > 0.f && rr < 3e38f)) val = dihedral_degenerate(x1, x2, cv, sv);
                                finish_one(r, 4, c, i & 1, val, cv, sv, out + (size_t)i * row_stride);
                            }
                        }
                    }
                }
            }
            __syncwarp();  // the records are rewritten by the next quad
        };

        // ---- main loop over the segment's frame blocks (32-bit counters and running pointers: this code runs once
        // per stage and warp)
        {
            const int64_t f_seg = seg_block0 * q.k;
            int frames_left = (int)min(q.F - f_seg, (int64_t)0x7fffffff);
            float* vstage = store ? q.values + (size_t)f_seg * row_stride : nullptr;
            const size_t vstage_step = (size_t)q.k * row_stride;
            const uint32_t fq_first = (uint32_t)(4 * fli) * (uint32_t)q.slot_bytes;
            const uint32_t fq_step = (uint32_t)(4 * fl) * (uint32_t)q.slot_bytes;
            const uint32_t tq_first = (uint32_t)q.tab_off + (MIC == MIC_ORTHO ? 96u : 192u) * (uint32_t)fli;
            const uint32_t tq_step = (MIC == MIC_ORTHO ? 96u : 192u) * (uint32_t)fl;
            const size_t vq_first = (size_t)(4 * fli) * row_stride, vq_step = (size_t)(4 * fl) * row_stride;
            const unsigned char* stage = stages + (size_t)s * q.stage_bytes;
            for (int it = 0; it < seg_count; ++it) {
                const int kk = min(q.k, frames_left);
                frames_left -= q.k;
                mbar_wait(&full[s], phase);
                if (working) {
                    const unsigned char* fq = stage + fq_first;
                    const unsigned char* tq = stage + tq_first;
                    float* vq = vstage + vq_first;
                    int left = kk - 4 * fli;                                  // frames from this warp's quad on
                    for (; left > 0; left -= 4 * fl) {
                        const int nvalid = min(4, left);
                        frames_done += nvalid;
                        do_quad(fq, reinterpret_cast<const float*>(tq), vq, nvalid);
                        fq += fq_step;
                        tq += tq_step;
                        vq += vq_step;
                    }
                }
                vstage += vstage_step;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                stage += q.stage_bytes;
                if (++s == q.nstage) {
                    s = 0;
                    phase ^= 1u;
                    stage = stages;
                }
            }
        }

        // ---- segment epilogue: per-lane float32 sums -> per-Bond float64 sums of the segment in a fixed order (no
        // floating-point atomics inside the CTA, so its contribution does not depend on scheduling); segments are then
        // combined with float64 atomics.  Two rounds of two chunks each; the scratch lives where the edge records were.
        double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // of the slots this warp reduces: sl = warp, warp + 7, ...
        constexpr int kHalf = 2 * kConsumers;
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");   // main loop (or previous round) is over
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const Role& r = role[2 * round + cc];
                const int i = cc * kConsumers + tid;
                float a, b;
                unpack2(r.s1, a, b);
                eacc[0 * kHalf + i] = a + b;
                unpack2(r.s2, a, b);
                eacc[1 * kHalf + i] = a + b;
                unpack2(r.c, a, b);
                eacc[2 * kHalf + i] = a + b;
                unpack2(r.s, a, b);
                eacc[3 * kHalf + i] = a + b;
                ecnt[i] = (working && r.slot >= 0) ? frames_done : 0;
            }
            asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");
            // one warp per Bond slot: lane-strided sums over the slot's lane-roles (and their copies in the warps that
            // share the group), then a fixed butterfly
            for (int sl = warp; sl < tile.nslots; sl += kWarps) {
                const int p0 = slist[sl], p1 = slist[sl + 1];
                const uint16_t* ent = slist + tile.nslots + 1;
                double a6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                for (int p = p0 + lane; p < p1; p += 32) {
                    const int e = ent[p];
                    const int c = e / kConsumers;
                    if ((c >> 1) != round) continue;
                    const int base = e - 2 * round * kConsumers;
                    for (int j = 0; j < fl; ++j) {
                        const int i = base + j * ng * 32;      // the same lane-role in the warp that is j frame lanes on
                        const int cnt = ecnt[i];
                        if (cnt == 0) continue;
                        a6[0] += (double)(cnt - (int)snbad[i + 2 * round * kConsumers]);
                        a6[1] += (double)eacc[0 * kHalf + i];
                        a6[2] += (double)eacc[1 * kHalf + i];
                        a6[3] += (double)eacc[2 * kHalf + i];
                        a6[4] += (double)eacc[3 * kHalf + i];
                        a6[5] += (double)cnt;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) a6[k] += __shfl_down_sync(0xffffffffu, a6[k], o);
                }
                if (lane == 0) {
                    double* dst = q.partials + (size_t)sbond[sl] * CGB_NPART;
                    const int where[6] = {0, 1, 2, 3, 4, 6};
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (a6[k] != 0.0) atomicAdd(dst + where[k], a6[k]);
                }
            }
        }
        (void)acc;
        if (do_hist) {
            for (int i = tid; i < tile.nslots * nbins; i += kConsumers) {
                const unsigned int v = shist[i];
                if (v) atomicAdd(q.hist + (size_t)sbond[i / nbins] * nbins + (i % nbins), (double)v);
            }
        }
    }
}

}  // namespace cgb

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" int cgb_measure_plan_size(const int32_t* col_beads, const int32_t* col_bond, int64_t n2, int64_t n3,
                                     int64_t n4, int64_t n_beads, int32_t max_span, int32_t max_slots,
                                     int32_t shape, CgbMeasPlanInfo* info) {
    if (!info) return CGB_EINVAL;
    cgb::plan::Compiler cc{col_beads, col_bond, n2, n3, n4, n_beads, max_span, max_slots, shape, nullptr, {}, {}};
    const int rc = cc.run();
    if (rc != CGB_OK) return rc;
    info->n_tiles = (int64_t)cc.out.tiles.size();
    info->n_groups = (int64_t)cc.out.groups.size();
    info->n_tile_bonds = (int64_t)cc.out.tile_bonds.size();
    info->n_tile_lists = (int64_t)cc.out.tile_lists.size();
    info->n_wide = (int64_t)cc.out.wide.size();
    info->n_staged = cc.out.n_staged;
    info->max_span = cc.out.max_span;
    info->max_slots = cc.out.max_slots;
    info->max_edge_chunks = cc.out.max_edge_chunks;
    info->max_tile_groups = 0;
    for (const CgbMeasTile& t : cc.out.tiles) info->max_tile_groups = std::max(info->max_tile_groups, t.ngroups);
    return CGB_OK;
}

extern "C" int cgb_measure_plan_build(const int32_t* col_beads, const int32_t* col_bond, int64_t n2, int64_t n3,
                                      int64_t n4, int64_t n_beads, int32_t max_span, int32_t max_slots,
                                      int32_t shape, CgbMeasTile* tiles, CgbMeasGroup* groups, CgbMeasLane* lanes,
                                      int32_t* tile_bonds, uint16_t* tile_lists, int32_t* col_out,
                                      int64_t* wide_cols) {
    const int64_t M = n2 + n3 + n4;
    if (M > 0 && !col_out) return CGB_EINVAL;
    cgb::plan::Compiler cc{col_beads, col_bond, n2, n3, n4, n_beads, max_span, max_slots, shape, col_out, {}, {}};
    const int rc = cc.run();
    if (rc != CGB_OK) return rc;
    const auto& o = cc.out;
    if ((!o.tiles.empty() && (!tiles || !groups || !lanes || !tile_bonds || !tile_lists)) ||
        (!o.wide.empty() && !wide_cols))
        return CGB_EINVAL;
    if (!o.tiles.empty()) {
        std::memcpy(tiles, o.tiles.data(), o.tiles.size() * sizeof(CgbMeasTile));
        std::memcpy(groups, o.groups.data(), o.groups.size() * sizeof(CgbMeasGroup));
        std::memcpy(lanes, o.lanes.data(), o.lanes.size() * sizeof(CgbMeasLane));
        std::memcpy(tile_bonds, o.tile_bonds.data(), o.tile_bonds.size() * sizeof(int32_t));
        std::memcpy(tile_lists, o.tile_lists.data(), o.tile_lists.size() * sizeof(uint16_t));
    }
    if (!o.wide.empty()) std::memcpy(wide_cols, o.wide.data(), o.wide.size() * sizeof(int64_t));
    return CGB_OK;
}

namespace cgb {

// Shared-memory map, stage geometry and grid of one launch.
static int fused_geometry(FusedParams& q, const CgbMeasureTuning* tune, int32_t shape, int32_t max_span,
                          int32_t max_tile_groups, int sm_count, unsigned* grid, size_t* smem_bytes) {
    const int min_fl = kWarps / std::max(1, std::min<int>(max_tile_groups, kWarps));   // frame lanes of the fullest tile
    const int ctas_per_sm = tune && tune->ctas_per_sm ? tune->ctas_per_sm : 2;
    if (ctas_per_sm < 1 || ctas_per_sm > 2) return CGB_EINVAL;
    const int kCtaSmem = ctas_per_sm == 2 ? 112 * 1024 : 224 * 1024;
    q.nstage = tune && tune->n_stages ? tune->n_stages : 3;
    if (q.nstage < 2 || q.nstage > kStagesMax) return CGB_EINVAL;
    auto up = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
    size_t off = 128 + (size_t)((q.max_slots + 3) & ~3) * 4;
    off = up(off, 16);
    q.off_sparam = (int)off;
    off += (size_t)q.max_slots * 16;
    q.off_shist = (int)off;
    off += q.hist ? (size_t)q.max_slots * q.nbins * 4 : 0;
    q.off_lists = (int)off;
    off = up(off + 2 * ((size_t)q.max_slots + 1 + kRoles), 16);
    q.off_eacc = 0;
    q.off_nbad = (int)off;
    off = up(off + (size_t)kRoles * 4, 128);
    q.off_ecache = (int)off;
    q.ecache_warp_bytes = shape == CGB_SHAPE_EAD ? RecLayout<CGB_SHAPE_EAD>::kWarpBytes : RecLayout<CGB_SHAPE_EEAA>::kWarpBytes;
    off += (size_t)kWarps * q.ecache_warp_bytes;
    q.off_stages = (int)off;
    const size_t fixed = off;
    const int64_t frame_bytes = q.B * 12;
    q.contiguous = (q.n_tiles == 1 && max_span == q.B) ? 1 : 0;
    int64_t slot;
    if (q.contiguous) {
        slot = frame_bytes;
    } else {
        slot = (int64_t)max_span * 12 + 48;
        slot += ((frame_bytes - slot) % 16 + 16) % 16;       // == 12 B (mod 16): see the producer
    }
    // frames per stage: whole quads, as many as fit the stage budget (12 bytes of box table per frame); the pipeline
    // is three stages deep unless only two stages of one quad of the widest tile fit
    const int64_t quad = (slot + 12) * 4 + 64 + 128;
    if ((int64_t)fixed + 2 * quad > kCtaSmem) return -1;    // two stages of one quad of the widest tile do not fit
    while (q.nstage > 2 && quad * q.nstage > (int64_t)kCtaSmem - (int64_t)fixed) --q.nstage;
    int64_t want = tune && tune->stage_bytes ? tune->stage_bytes : 20 * 1024;
    int64_t k = 4;
    for (;;) {
        const int64_t room = ((int64_t)kCtaSmem - (int64_t)fixed) / q.nstage - 128;
        k = (std::min(want, room) - 64) / (slot + 12);
        k -= k % 4;
        if (k >= 4 * min_fl) k -= k % (4 * min_fl);              // whole rounds of quads over the frame lanes
        if (k < 4) k = 4;
        k = std::min<int64_t>(k, 64);
        k = std::min<int64_t>(k, std::max<int64_t>(4, (q.F + 3) / 4 * 4));
        q.k = (int)k;
        q.slot_bytes = (int)slot;
        q.tab_off = (int)((k * slot + 32 + 15) & ~15LL);
        q.stage_bytes = (int)((q.tab_off + 12 * 4 * k + 127) & ~127LL);
        *smem_bytes = fixed + (size_t)q.nstage * q.stage_bytes;
        if (*smem_bytes <= (size_t)kCtaSmem) break;
        if (k > 4)
            want = (k - 4) * (slot + 12) + 64;                   // one quad less per stage
        else if (q.nstage > 2)
            --q.nstage;
        else
            return -1;
    }
    // persistent CTAs: the (tile, frame block) items are dealt out in equal contiguous shares; a CTA reduces its
    // float32 sums (segment epilogue) at every change of tile and after at most `max_run` frames per lane
    q.nblocks = (q.F + k - 1) / k;
    const int64_t max_run = tune && tune->max_run ? tune->max_run : 512;
    q.seg_cap = (int)std::max<int64_t>(1, max_run * min_fl / k);   // a lane sees k / fl frames per block
    const int64_t items = q.n_tiles * q.nblocks;
    *grid = (unsigned)std::min<int64_t>((int64_t)sm_count * ctas_per_sm, items);
    return CGB_OK;
}

}  // namespace cgb

extern "C" int cgb_measure_fused(const float* cg_xyz, const float* box_vec, int ortho, int64_t n_frames,
                                 int64_t n_beads, const CgbMeasTile* tiles, int64_t n_tiles,
                                 const CgbMeasGroup* groups, const CgbMeasLane* lanes, const int32_t* tile_bonds,
                                 const uint16_t* tile_lists, int32_t shape, int32_t max_span, int32_t max_slots,
                                 int32_t max_edge_chunks, int32_t max_tile_groups, const float* shift, float* values,
                                 int64_t row_stride, double* partials, double* hist, const float* hist_lo,
                                 const float* hist_hi, int32_t n_bins, const CgbMeasureTuning* tuning, void* stream) {
    using namespace cgb;
    if (n_frames < 0 || n_beads < 0 || n_tiles < 0) return CGB_EINVAL;
    if (n_frames == 0 || n_tiles == 0) return CGB_OK;
    if (!cg_xyz || !tiles || !groups || !lanes || !tile_bonds || !tile_lists || !partials) return CGB_EINVAL;
    if (hist && (!hist_lo || !hist_hi || n_bins < 1 || n_bins > 512)) return CGB_EINVAL;  // fast-bin error bound
    if (values && row_stride < 1) return CGB_EINVAL;
    if (shape != CGB_SHAPE_EAD && shape != CGB_SHAPE_EEAA) return CGB_EINVAL;
    if (max_span < 1 || max_slots < 1 || max_edge_chunks < 1 || max_edge_chunks > CGB_MEAS_EDGE_CHUNKS ||
        n_tiles > 0x7fffffffLL)
        return CGB_EINVAL;
    if (reinterpret_cast<uintptr_t>(cg_xyz) & 3u) return CGB_EALIGN;
    FusedParams q;
    std::memset(&q, 0, sizeof q);
    q.xyz = cg_xyz;
    q.box = box_vec;
    q.F = n_frames;
    q.B = n_beads;
    q.tiles = tiles;
    q.groups = groups;
    q.lanes = lanes;
    q.tile_bonds = tile_bonds;
    q.tile_lists = tile_lists;
    q.shift = shift;
    q.hlo = hist_lo;
    q.hhi = hist_hi;
    q.values = values;
    q.row_stride = row_stride;
    q.partials = partials;
    q.hist = hist;
    q.n_tiles = n_tiles;
    q.nbins = hist ? n_bins : 0;
    q.max_slots = max_slots;
    int dev = 0, sms = 0;
    CGB_CUDA_TRY(cudaGetDevice(&dev));
    CGB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned grid = 0;
    size_t smem = 0;
    (void)max_edge_chunks;
    const int rc = fused_geometry(q, tuning, shape, max_span, max_tile_groups, sms, &grid, &smem);
    if (rc == -1) return CGB_ETOOBIG;
    if (rc != CGB_OK) return rc;
    const bool triclinic = box_vec != nullptr && !ortho;     // no box: the orthorhombic path with a zero box table
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CGB_FUSED(M_, S_)                                                                                \
    do {                                                                                                 \
        auto kern = measure_fused_kernel<M_, S_>;                                                        \
        CGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, kConsumers + 32, smem, st>>>(q);                                                    \
        return (int)cudaGetLastError();                                                                  \
    } while (0)
    if (triclinic && shape == CGB_SHAPE_EAD) CGB_FUSED(MIC_TRICLINIC, CGB_SHAPE_EAD);
    if (triclinic) CGB_FUSED(MIC_TRICLINIC, CGB_SHAPE_EEAA);
    if (shape == CGB_SHAPE_EAD) CGB_FUSED(MIC_ORTHO, CGB_SHAPE_EAD);
    CGB_FUSED(MIC_ORTHO, CGB_SHAPE_EEAA);
#undef CGB_FUSED
}

/* Host only: the launch configuration cgb_measure_fused would use (no CUDA call), for tests and tools.
 * out = {grid, shared-memory bytes, frames per stage, stages, blocks per segment, frame blocks, whole-frame tile, 0}. */
extern "C" int cgb_measure_fused_config(int64_t n_frames, int64_t n_beads, int64_t n_tiles, int32_t max_span,
                                        int32_t max_slots, int32_t max_edge_chunks, int32_t max_tile_groups,
                                        int32_t n_bins, int with_hist, const CgbMeasureTuning* tuning,
                                        int32_t sm_count, int64_t* out) {
    using namespace cgb;
    if (!out || n_frames < 1 || n_beads < 1 || n_tiles < 1 || sm_count < 1) return CGB_EINVAL;
    FusedParams q;
    std::memset(&q, 0, sizeof q);
    q.F = n_frames;
    q.B = n_beads;
    q.n_tiles = n_tiles;
    q.nbins = with_hist ? n_bins : 0;
    q.hist = with_hist ? reinterpret_cast<double*>(sizeof(double)) : nullptr;   // only tested against NULL
    q.max_slots = max_slots;
    unsigned grid = 0;
    size_t smem = 0;
    const int rc = fused_geometry(q, tuning, max_edge_chunks > 1 ? CGB_SHAPE_EEAA : CGB_SHAPE_EAD, max_span,
                                  max_tile_groups, sm_count, &grid, &smem);
    if (rc == -1) return CGB_ETOOBIG;
    if (rc != CGB_OK) return rc;
    out[0] = grid;
    out[1] = (int64_t)smem;
    out[2] = q.k;
    out[3] = q.nstage;
    out[4] = q.seg_cap;
    out[5] = q.nblocks;
    out[6] = q.contiguous;
    out[7] = 0;
    return CGB_OK;
}

/* Limits to compile a plan with: the shared memory of a CTA (112 KB, two CTAs per SM) holds the Bond tables and
 * histograms of a tile, the reduction scratch, the edge records of seven warps and at least two stages of one quad
 * of the widest tile.  The caller says how many Bonds it would like a tile to hold (what its largest molecule has);
 * it gets max_slots <= that and the widest bead range that still fits next to them. */
extern "C" int cgb_measure_plan_limits(int32_t n_bins, int with_hist, int32_t want_slots, int32_t* max_span,
                                       int32_t* max_slots) {
    if (!max_span || !max_slots || n_bins < 0 || n_bins > 512 || want_slots < 1) return CGB_EINVAL;
    const int64_t budget = 112 * 1024;
    const int64_t per_slot = 22 + (with_hist ? 4 * (int64_t)n_bins : 0);
    // everything that does not depend on the plan: barriers, lists, reduction scratch, NaN counters, 7 x 4 KB records
    const int64_t fixed = 128 + 2 * (1 + cgb::kRoles) + 16 + cgb::kRoles * 4 + 128 +
                          (int64_t)cgb::kWarps * 2 * CGB_MEAS_EDGE_CHUNKS * 32 * 32 + 64;
    auto stages_for = [](int64_t span) { return 2 * (4 * (12 * span + 48 + 16 + 12) + 64 + 128 + 128); };
    const int64_t min_span = 64, best_span = 384;
    int64_t slots = std::min<int64_t>(want_slots, (budget - fixed - stages_for(min_span)) / per_slot);
    slots = std::max<int64_t>(slots, 1);
    int64_t span = best_span;
    while (span > min_span && fixed + slots * per_slot + stages_for(span) > budget) span -= 8;
    *max_slots = (int32_t)slots;
    *max_span = (int32_t)span;
    return CGB_OK;
}
