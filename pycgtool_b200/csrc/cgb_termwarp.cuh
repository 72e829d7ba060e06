// The work of one measurement warp, shared by the fused measurement kernel (K2, cgb_measure_fused.cu) and the fused
// mapping + measurement kernel (cgb_map_measure.cu).
//
// Reference: pycgtool/bondset.py:496-531 (mdtraj.compute_distances / compute_angles / compute_dihedrals on the CG
// frame) and the statistics of pycgtool/util.py:24-53.  A warp owns one *group* of the plan (cgb_measure_plan.cpp.inc):
// its edge lanes gather two beads each from a CG tile held in shared memory, leave the minimum-image displacement in
// a per-warp record and finish the lengths; after a __syncwarp its angle / dihedral lanes read two or three records.
// Four frames (two packed float32x2 frame pairs) per call.  Where the CG tile comes from -- bulk copies of the CG
// trajectory (K2) or the mapping warps of the same CTA (fused kernel) -- is the caller's business.
#pragma once

#include "cgb_terms.cuh"

namespace cgb {

constexpr int kTermWarps = CGB_MEAS_WARPS;        // measurement warps a CTA can have = groups per tile at most
constexpr int kTermThreads = kTermWarps * 32;
constexpr int kChunks = CGB_MEAS_CHUNKS;
constexpr int kRoles = kChunks * kTermThreads;    // lane-roles of a CTA (chunk, warp, lane)

// Words between the shared-memory histogram rows of two Bond slots (see TermTables::hstride).
__host__ __device__ constexpr int hist_stride(int nbins) { return nbins | 1; }

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

// Per-lane accumulators of one chunk of the warp's group: float32 sums of this CTA segment (two frames per register
// pair): d, d^2, cos, sin.  Everything else a lane needs to know about its role -- bead offsets or record offsets,
// orientation signs, Bond slot, output column -- lives in the warp's *role table* in shared memory (16 bytes per
// lane and chunk, one LDS.128 per chunk and quad), which keeps the registers for the arithmetic.
struct Role {
    f2 s1, s2, c, s;
};

// Role table entry (uint4) of lane l in chunk c, at role32 + (c * 32 + l) * 16:
//   edges:  x = byte offsets of the two beads inside a staged frame (a | b << 16), y = Bond slot (-1: no column),
//           z = float offset of the lane's box-table entry inside its frame (pair)'s block (12 * instance, when
//           real frames are folded; else 0), w = position in the values row (-1: the lane emits nothing);
//   terms:  x = byte offsets of the records of the first two edges inside the warp's record area (e0 | e1 << 16),
//           z = third record (low 16 bits) | bit 16: g1 < 0 | bit 17: g2 < 0, y and w as above.
// Lanes beyond the chunk's population hold a harmless role (bead 0 twice / record 0) and w = -1.
constexpr int kRoleTableBytes = CGB_MEAS_CHUNKS * 32 * 16;

// Shared-memory tables of the tile the CTA works on and the outputs of the launch (uniform over the CTA).
struct TermTables {
    int* sbond;                 // [nslots] global Bond of every tile-local slot
    float4* sparam;             // [nslots] {-shift, bins / (hi - lo), -lo * that - 0.5, biased address of the histogram row}
    unsigned int* shist;        // [nslots][nbins] histogram counts of the segment
    uint16_t* slist;            // slot_ptr[nslots + 1], then the lane-roles of every slot
    unsigned int* snbad;        // [kRoles] NaN values per lane-role (rare path)
    float* eacc;                // reduction scratch: [4][2 chunks x kTermThreads] sums ...
    int* ecnt;                  // ... and [2 x kTermThreads] frame counts (may alias the edge records)
    uint32_t shist32;           // shared-memory address of shist
    uint32_t dummy32;           // shared-memory address of a scratch word (see hist_add_pair_pred)
    uint32_t rs32;              // row_stride as 32 bits (row_stride * 4 frames < 2^32: checked by the host)
    const float* shift;         // per Bond (global), may be NULL
    const float* hlo;
    const float* hhi;
    float* values;              // [F][row_stride] or NULL
    int64_t row_stride;
    double* partials;
    double* hist;               // [n_bonds][nbins] float64 counts or NULL
    int nbins;
    int fold;                   // real frames folded into one staged frame (>= 1): the box table then holds `fold`
                                // entries per staged frame (pair), picked by the edge lane's instance
    int hstride;                // words between the histogram rows of two slots: odd, so that equal bins of
                                // different Bonds (similar distributions!) fall into different banks
};

// Fill the tile's tables; `vt` / `nthreads`: the calling thread's index among the threads that cooperate.
__device__ __forceinline__ void term_tables_fill(const TermTables& t, const CgbMeasTile& tile, const int32_t* tile_bonds,
                                                 const uint16_t* tile_lists, int vt, int nthreads) {
    const bool do_hist = t.hist != nullptr;
    for (int i = vt; i < tile.nslots; i += nthreads) {
        const int bond = __ldg(tile_bonds + tile.slot0 + i);
        t.sbond[i] = bond;
        float4 prm = make_float4(t.shift ? -__ldg(t.shift + bond) : 0.f, 0.f, 0.f, 0.f);
        if (do_hist) {
            const float lo = __ldg(t.hlo + bond), hi = __ldg(t.hhi + bond);
            prm.y = (float)t.nbins / (hi - lo);
            prm.z = fmaf(-lo, prm.y, -0.5f);
            // w: shared-memory address of the slot's histogram row, biased so that the float bits of 2^23 + bin
            // (0x4b000000 + bin) scale straight into the address of the bin (hist_add_pair_pred)
            prm.w = __uint_as_float(t.shist32 + (uint32_t)(i * t.hstride) * 4u - 4u * 0x4b000000u);
        }
        t.sparam[i] = prm;
    }
    if (do_hist)
        for (int i = vt; i < tile.nslots * t.hstride; i += nthreads) t.shist[i] = 0u;
    for (int i = vt; i < tile.nslots + 1 + tile.nlist; i += nthreads) t.slist[i] = __ldg(tile_lists + tile.list0 + i);
    for (int i = vt; i < kRoles; i += nthreads) t.snbad[i] = 0u;
}

template <int MIC, int SHAPE>
struct TermWarp {
    static_assert(MIC == MIC_ORTHO || MIC == MIC_TRICLINIC, "no box = orthorhombic path with a zero box table");
    using Rec = RecLayout<SHAPE>;

    Role role[kChunks];
    uint32_t role32;                  // shared-memory address of this lane's entry of chunk 0 in the role table
    uint32_t rec32;                   // shared-memory address of this warp's edge records
    int frames_done;                  // frames this warp has evaluated in the current segment (uniform)
    int vt;                           // this thread among the measurement threads: 32 * (warp among them) + lane

    // Roles of this lane for group `grp` of the tile; h0 = byte offset of bead 0 of the tile inside a staged frame;
    // ecache32 = shared-memory address of this warp's edge records; table32 = of this warp's role table.
    // bead0 / fold_beads: first bead of the tile and beads per real frame when `fold` real frames make one staged
    // frame (0: no folding) -- an edge lane's instance (which real frame, hence which box) is its bead / fold_beads.
    __device__ __forceinline__ void setup(bool working, const CgbMeasGroup& grp, const CgbMeasLane* lanes, int lane,
                                          uint32_t h0, uint32_t ecache32, uint32_t table32, int vt_, int bead0 = 0,
                                          int fold_beads = 0) {
        vt = vt_;
        frames_done = 0;
        rec32 = ecache32;
        role32 = table32 + (uint32_t)lane * 16u;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            Role& r = role[c];
            r.s1 = r.s2 = r.c = r.s = 0;
            constexpr int kNone = 0;
            const int kind = chunk_kind<SHAPE>(c);
            if (kind == kNone) continue;
            const int nl = working ? grp.nlanes[c] : 0, ne = working ? grp.nemit[c] : 0;
            uint4 e = make_uint4(kind == 2 ? (h0 | (h0 << 16)) : 0u, 0xffffffffu, 0u, 0xffffffffu);
            if (lane < nl) {
                const CgbMeasLane ln = lanes[c * 32 + lane];
                if (kind == 2) {
                    e.x = ((ln.ref & 0xffffu) * 12u + h0) | (((ln.ref >> 16) * 12u + h0) << 16);
                    if (fold_beads > 0) e.z = 12u * (uint32_t)((bead0 + (int)(ln.ref & 0xffffu)) / fold_beads);
                } else {
                    const uint32_t e0 = ln.ref & 0x7fu, e1 = (ln.ref >> 8) & 0x7fu, e2 = (ln.ref >> 16) & 0x7fu;
                    const uint32_t n0 = (ln.ref >> 7) & 1u, n1 = (ln.ref >> 15) & 1u, n2 = (ln.ref >> 23) & 1u;
                    e.x = (e0 * 16u) | ((e1 * 16u) << 16);
                    // angles: g1 = s0 s1.  dihedrals: g1 = s0 s1 s2, g2 = s0 s2.
                    const uint32_t g1neg = kind == 3 ? (n0 ^ n1) : (n0 ^ n1 ^ n2), g2neg = kind == 3 ? 0u : (n0 ^ n2);
                    e.z = (e2 * 16u) | (g1neg << 16) | (g2neg << 17);
                }
                e.y = (uint32_t)ln.slot;
                if (lane < ne) e.w = (uint32_t)(grp.vcol0[c] + lane);
            }
            sts_v4u32(role32 + (uint32_t)c * 512u, e);
        }
        __syncwarp();
    }

    // This lane's Bond slot in chunk c (-1: none), from the role table.
    __device__ __forceinline__ int slot_of(int c) const { return (int)lds_v4u32(role32 + (uint32_t)c * 512u).y; }

    // Exact-rule histogram update of one value (next to a bin edge, out of range or NaN: ~0.02 % of the values).
    __device__ __forceinline__ void hist_exact(const TermTables& t, int slot, float v) {
        const int bond = t.sbond[slot];
        hist_add_exact(t.shist + slot * t.hstride, v, __ldg(t.hlo + bond), __ldg(t.hhi + bond), t.nbins);
    }

    // Statistics, histogram and stores of the two evaluated frame pairs of a quad (the common case: all four frames
    // valid).  Straight-line code up to the one rare branch at its end, so that the two pairs' dependency chains --
    // and the evaluation before them -- interleave.
    template <int KIND, bool STORE, bool HIST>
    __device__ __forceinline__ void finish_fast(const TermTables& t, Role& r, int slot, f2 va, f2 cva, f2 sva, f2 vb,
                                                f2 cvb, f2 svb, float* vq, uint32_t col) {
        const float4 prm = t.sparam[slot];
        const f2 nshift2 = splat2(prm.x);
        const f2 da = KIND == 4 ? wrap_about2(va, nshift2) : add2(va, nshift2);
        const f2 db = KIND == 4 ? wrap_about2(vb, nshift2) : add2(vb, nshift2);
        r.s1 = add2(r.s1, add2(da, db));
        r.s2 = fma2(db, db, fma2(da, da, r.s2));
        if (KIND != 2) {
            r.c = add2(r.c, add2(cva, cvb));
            r.s = add2(r.s, add2(sva, svb));
        }
        if (STORE) {
            // 32-bit element offsets from the quad's first row: one IMAD.WIDE per store
            __stcs(vq + col, lo2(va));
            __stcs(vq + (col + t.rs32), hi2(va));
            __stcs(vq + (col + 2u * t.rs32), lo2(vb));
            __stcs(vq + (col + 3u * t.rs32), hi2(vb));
        }
        if (HIST) {
            const f2 hn = splat2(prm.y), hc = splat2(prm.z);
            const uint32_t row_adj = __float_as_uint(prm.w);   // address of the slot's histogram row - 4 * 0x4b000000
            bool f0, f1, f2_, f3;
            const bool fine_a = hist_add_pair_pred(row_adj, t.dummy32, va, hn, hc, (unsigned)t.nbins, f0, f1);
            const bool fine_b = hist_add_pair_pred(row_adj, t.dummy32, vb, hn, hc, (unsigned)t.nbins, f2_, f3);
            if (!(fine_a && fine_b)) {
                if (!f0) hist_exact(t, slot, lo2(va));
                if (!f1) hist_exact(t, slot, hi2(va));
                if (!f2_) hist_exact(t, slot, lo2(vb));
                if (!f3) hist_exact(t, slot, hi2(vb));
            }
        }
    }

    // One value on its own (rare): a frame next to the end of the trajectory or a degenerate / NaN operand.
    __device__ __forceinline__ void finish_one(const TermTables& t, Role& r, int slot, int kind, int c, int half, float v,
                                            float cv, float sv, float* out) {
        if (t.values != nullptr) __stcs(out, v);
        if (!(v == v)) {  // NaN-skipping statistics (np.nanmean / np.nanvar, util.py:36,53)
            t.snbad[c * kTermThreads + vt] += 1u;
            return;
        }
        const float4 prm = t.sparam[slot];
        const f2 vv = half ? pack2(0.f, v) : pack2(v, 0.f);
        const f2 ns = half ? pack2(0.f, prm.x) : pack2(prm.x, 0.f);
        const f2 d = kind == 4 ? wrap_about2(vv, ns) : add2(vv, ns);
        r.s1 = add2(r.s1, d);
        r.s2 = fma2(d, d, r.s2);
        if (kind != 2) {
            r.c = add2(r.c, half ? pack2(0.f, cv) : pack2(cv, 0.f));
            r.s = add2(r.s, half ? pack2(0.f, sv) : pack2(sv, 0.f));
        }
        if (t.hist != nullptr) {
            const int bond = t.sbond[slot];
            const int bin = hist_bin(v, __ldg(t.hlo + bond), __ldg(t.hhi + bond), prm.y, t.nbins);
            if (bin >= 0) atomicAdd(t.shist + slot * t.hstride + bin, 1u);
        }
    }

    static __device__ __forceinline__ float pick(int i, f2 a, f2 b) {  // frame i (0..3) of the two pairs of a quad
        const f2 x = i < 2 ? a : b;
        return (i & 1) ? hi2(x) : lo2(x);
    }

    static __device__ __forceinline__ void load_rec(int pp, uint32_t addr, EdgeRec& e) {
        if (pp == 0) {
            lds_v2u64<0>(addr, e.r[0], e.r[1]);
            lds_v2u64<Rec::kHalf>(addr, e.r[2], e.n2);
        } else {
            lds_v2u64<Rec::kPair>(addr, e.r[0], e.r[1]);
            lds_v2u64<Rec::kPair + Rec::kHalf>(addr, e.r[2], e.n2);
        }
    }

    // One quad of this warp's group: fq = the first of four staged frames `slot_bytes` apart, tq = their box table
    // (orthorhombic: 12 floats per frame pair; triclinic: 12 per frame), vq = row of the first frame in `values`,
    // nvalid = how many of the four frames lie inside the trajectory.  Every lane evaluates its role (lanes beyond a
    // chunk's population hold a harmless one), so the arithmetic is straight-line; only a lane that emits a column
    // goes on to the statistics.
    template <bool STORE, bool HIST>
    __device__ __forceinline__ void quad(const TermTables& t, int lane, const unsigned char* fq, uint32_t slot_bytes,
                                         const float* tq, float* vq, int nvalid) {
        const bool whole = nvalid == 4;
        frames_done += nvalid;
        // ---- edges: displacement records of both frame pairs, and the lengths
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            if (chunk_kind<SHAPE>(c) != 2) continue;
            Role& r = role[c];
            const uint4 ro = lds_v4u32(role32 + (uint32_t)c * 512u);
            const uint32_t a0 = ro.x & 0xffffu, a1 = ro.x >> 16;
            EdgeRec e[2];
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const unsigned char* fa = fq + (size_t)(2 * pp) * slot_bytes;
                const unsigned char* fb = fa + slot_bytes;
                if (MIC == MIC_TRICLINIC) {
                    float rr[2][3];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        Box bx;
                        const float4* tb = reinterpret_cast<const float4*>(tq + (2 * pp + h) * 12 * t.fold + ro.z);
                        const float4 t0 = tb[0], t1 = tb[1], t2 = tb[2];
                        bx.a[0] = t0.x; bx.a[1] = t0.y; bx.a[2] = t0.z;
                        bx.b[0] = t1.x; bx.b[1] = t1.y; bx.b[2] = t1.z;
                        bx.c[0] = t2.x; bx.c[1] = t2.y; bx.c[2] = t2.z;
                        const unsigned char* fr = h ? fb : fa;
                        const float* xa = reinterpret_cast<const float*>(fr + a0);
                        const float* xb = reinterpret_cast<const float*>(fr + a1);
                        const float pa[3] = {xa[0], xa[1], xa[2]}, pb[3] = {xb[0], xb[1], xb[2]};
                        edge_triclinic(pa, pb, bx, rr[h]);
                    }
#pragma unroll
                    for (int d = 0; d < 3; ++d) e[pp].r[d] = pack2(rr[0][d], rr[1][d]);
                    edge_norm2(e[pp]);
                } else {
                    const float* xa0 = reinterpret_cast<const float*>(fa + a0);
                    const float* xa1 = reinterpret_cast<const float*>(fb + a0);
                    const float* xb0 = reinterpret_cast<const float*>(fa + a1);
                    const float* xb1 = reinterpret_cast<const float*>(fb + a1);
                    f2 xa[3], xb[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        xa[d] = pack2(xa0[d], xa1[d]);
                        xb[d] = pack2(xb0[d], xb1[d]);
                    }
                    const ulonglong2* tb = reinterpret_cast<const ulonglong2*>(tq + pp * 12 * t.fold + ro.z);
                    const ulonglong2 q0 = tb[0], q1 = tb[1], q2 = tb[2];
                    const f2 nL[3] = {q0.x, q0.y, q1.x}, inv[3] = {q1.y, q2.x, q2.y};
                    edge_pair<MIC_ORTHO>(xa, xb, nL, inv, e[pp]);
                }
            }
            // the records in two 16-byte halves, each in its own array: conflict-free 128-bit accesses
            const uint32_t rec = rec32 + (uint32_t)(c * 32 + lane) * 16u;
            sts_v2u64<0>(rec, e[0].r[0], e[0].r[1]);
            sts_v2u64<Rec::kHalf>(rec, e[0].r[2], e[0].n2);
            sts_v2u64<Rec::kPair>(rec, e[1].r[0], e[1].r[1]);
            sts_v2u64<Rec::kPair + Rec::kHalf>(rec, e[1].r[2], e[1].n2);
            if ((int)ro.w >= 0) {
                f2 va, vb;
                const bool oka = length_pair(e[0].n2, va), okb = length_pair(e[1].n2, vb);
                if (oka && okb && whole) {
                    finish_fast<2, STORE, HIST>(t, r, (int)ro.y, va, 0, 0, vb, 0, 0, vq, ro.w);
                } else {
                    float* out = vq + ro.w;
#pragma unroll 1
                    for (int i = 0; i < nvalid; ++i) {
                        const bool ok = i < 2 ? oka : okb;
                        const float v = ok ? pick(i, va, vb) : __fsqrt_rn(pick(i, e[0].n2, e[1].n2));
                        finish_one(t, r, (int)ro.y, 2, c, i & 1, v, 0.f, 0.f, out + (size_t)i * t.row_stride);
                    }
                }
            }
        }
        __syncwarp();

        // ---- angles and dihedrals from the records
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            constexpr int kNone = 0;
            const int kind = chunk_kind<SHAPE>(c);
            if (kind == kNone || kind == 2) continue;
            Role& r = role[c];
            const uint4 ro = lds_v4u32(role32 + (uint32_t)c * 512u);
            const uint32_t r0 = rec32 + (ro.x & 0xffffu), r1 = rec32 + (ro.x >> 16);
            const float g1 = __uint_as_float(0x3f800000u | ((ro.z & 0x10000u) << 15));
            const bool emit = (int)ro.w >= 0;
            if (kind == 3) {
                f2 v2[2], cv2[2], sv2[2], den[2];
                bool ok[2];
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    EdgeRec u, v;
                    load_rec(pp, r0, u);
                    load_rec(pp, r1, v);
                    ok[pp] = angle_pair(u, v, g1, v2[pp], cv2[pp], sv2[pp], den[pp]);
                }
                if (emit) {
                    if (ok[0] && ok[1] && whole) {
                        finish_fast<3, STORE, HIST>(t, r, (int)ro.y, v2[0], cv2[0], sv2[0], v2[1], cv2[1], sv2[1], vq, ro.w);
                    } else {
                        float* out = vq + ro.w;
#pragma unroll 1
                        for (int i = 0; i < nvalid; ++i) {
                            const float dn = pick(i, den[0], den[1]);
                            const bool good = dn > 0.f && dn < 3e38f;
                            const float val = good ? pick(i, v2[0], v2[1]) : __int_as_float(0x7fc00000);
                            finish_one(t, r, (int)ro.y, 3, c, i & 1, val, pick(i, cv2[0], cv2[1]),
                                       pick(i, sv2[0], sv2[1]), out + (size_t)i * t.row_stride);
                        }
                    }
                }
            } else {
                const uint32_t r2 = rec32 + (ro.z & 0xffffu);
                const float g2 = __uint_as_float(0x3f800000u | ((ro.z & 0x20000u) << 14));
                f2 v2[2], cv2[2], sv2[2], p1[2], p2[2];
                bool ok[2];
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    EdgeRec b1, b2, b3;
                    load_rec(pp, r0, b1);
                    load_rec(pp, r1, b2);
                    load_rec(pp, r2, b3);
                    ok[pp] = dihedral_pair(b1, b2, b3, g1, g2, v2[pp], cv2[pp], sv2[pp], p1[pp], p2[pp]);
                }
                if (emit) {
                    if (ok[0] && ok[1] && whole) {
                        finish_fast<4, STORE, HIST>(t, r, (int)ro.y, v2[0], cv2[0], sv2[0], v2[1], cv2[1], sv2[1], vq, ro.w);
                    } else {
                        float* out = vq + ro.w;
#pragma unroll 1
                        for (int i = 0; i < nvalid; ++i) {
                            const float x1 = pick(i, p1[0], p1[1]), x2 = pick(i, p2[0], p2[1]);
                            const float rr = fmaf(x1, x1, x2 * x2);
                            float cv = pick(i, cv2[0], cv2[1]), sv = pick(i, sv2[0], sv2[1]);
                            float val = pick(i, v2[0], v2[1]);
                            if (!(rr > 0.f && rr < 3e38f)) val = dihedral_degenerate(x1, x2, cv, sv);
                            finish_one(t, r, (int)ro.y, 4, c, i & 1, val, cv, sv, out + (size_t)i * t.row_stride);
                        }
                    }
                }
            }
        }
        __syncwarp();  // the records are rewritten by the next quad
    }

    // The quad with the launch's (uniform) output options resolved at compile time.
    template <int MODE>
    __device__ __forceinline__ void quad_mode(const TermTables& t, int lane, const unsigned char* fq, uint32_t slot_bytes,
                                              const float* tq, float* vq, int nvalid) {
        quad<(MODE & 1) != 0, (MODE & 2) != 0>(t, lane, fq, slot_bytes, tq, vq, nvalid);
    }

    // Segment epilogue, executed by all `nwarps` measurement warps of the CTA (named barrier `bar` over 32 * nwarps
    // threads): per-lane float32 sums -> per-Bond float64 sums of the segment in a fixed order (no floating-point
    // atomics inside the CTA, so its contribution does not depend on scheduling); segments are then combined with
    // float64 atomics.  Two rounds of two chunks each; `ng` groups of the tile, `fl` warps per group.
    __device__ __forceinline__ void epilogue(const TermTables& t, const CgbMeasTile& tile, bool working, int warp,
                                             int lane, int nwarps, int ng, int fl, int bar) {
        constexpr int kHalfRoles = 2 * kTermThreads;
        const int nthreads = 32 * nwarps;
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");   // main loop (or previous round) is over
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const Role& r = role[2 * round + cc];
                const int i = cc * kTermThreads + vt;
                float a, b;
                unpack2(r.s1, a, b);
                t.eacc[0 * kHalfRoles + i] = a + b;
                unpack2(r.s2, a, b);
                t.eacc[1 * kHalfRoles + i] = a + b;
                unpack2(r.c, a, b);
                t.eacc[2 * kHalfRoles + i] = a + b;
                unpack2(r.s, a, b);
                t.eacc[3 * kHalfRoles + i] = a + b;
                t.ecnt[i] = (working && slot_of(2 * round + cc) >= 0) ? frames_done : 0;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");
            // one warp per Bond slot: lane-strided sums over the slot's lane-roles (and their copies in the warps that
            // share the group), then a fixed butterfly
            for (int sl = warp; sl < tile.nslots; sl += nwarps) {
                const int p0 = t.slist[sl], p1 = t.slist[sl + 1];
                const uint16_t* ent = t.slist + tile.nslots + 1;
                double a6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                for (int p = p0 + lane; p < p1; p += 32) {
                    const int e = ent[p];
                    const int c = e / kTermThreads;
                    if ((c >> 1) != round) continue;
                    const int base = e - 2 * round * kTermThreads;
                    for (int j = 0; j < fl; ++j) {
                        const int i = base + j * ng * 32;      // the same lane-role in the warp that is j frame lanes on
                        const int cnt = t.ecnt[i];
                        if (cnt == 0) continue;
                        a6[0] += (double)(cnt - (int)t.snbad[i + 2 * round * kTermThreads]);
                        a6[1] += (double)t.eacc[0 * kHalfRoles + i];
                        a6[2] += (double)t.eacc[1 * kHalfRoles + i];
                        a6[3] += (double)t.eacc[2 * kHalfRoles + i];
                        a6[4] += (double)t.eacc[3 * kHalfRoles + i];
                        a6[5] += (double)cnt;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) a6[k] += __shfl_down_sync(0xffffffffu, a6[k], o);
                }
                if (lane == 0) {
                    double* dst = t.partials + (size_t)t.sbond[sl] * CGB_NPART;
                    const int where[6] = {0, 1, 2, 3, 4, 6};
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (a6[k] != 0.0) atomicAdd(dst + where[k], a6[k]);
                }
            }
        }
        if (t.hist != nullptr) {
            for (int i = vt; i < tile.nslots * t.hstride; i += nthreads) {
                const unsigned int v = t.shist[i];      // (the padding word of a row is never written: zero)
                if (v) atomicAdd(t.hist + (size_t)t.sbond[i / t.hstride] * t.nbins + (i % t.hstride), (double)v);
            }
        }
    }
};

}  // namespace cgb
