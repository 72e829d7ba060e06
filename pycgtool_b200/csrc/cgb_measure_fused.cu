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
constexpr int kStagesMax = 4;
constexpr int kEpiAcc = 6;                        // per lane-role float64 handed to the CTA epilogue

struct FusedParams {
    const float* xyz;
    const float* box;
    int64_t F, B;
    const CgbMeasTile* tiles;
    const CgbMeasGroup* groups;
    const CgbMeasLane* lanes;
    const int32_t* tile_bonds;
    const float* shift;
    const float* hlo;
    const float* hhi;
    float* values;
    int64_t row_stride;
    double* partials;
    double* hist;
    int32_t nbins;
    int32_t max_slots;
    int32_t k;             // frames per stage (multiple of 4)
    int32_t nstage;
    int32_t slot_bytes;    // bytes between the staged frames of a stage (== 12 B for a whole-frame tile)
    int32_t tab_off;       // byte offset of the box table inside a stage
    int32_t stage_bytes;
    int32_t contiguous;    // the tile is the whole CG frame: one bulk copy per stage
    int32_t head_bytes;    // barriers + sbond + shist
    int32_t ecache_warp_bytes;
    int32_t blocks_per_cta;  // frame blocks (stages) one CTA works through
};

// Per-lane state of one chunk of the warp's group.
struct Role {
    uint32_t a0, a1, a2;   // edge: byte offsets of the two beads in a staged frame; angle / dihedral: record offsets
    float g1, g2;          // orientation signs
    unsigned int* hist;    // this lane's histogram row in shared memory
    float nshift, hnorm, hc;
    f2 s1, s2, c, s;
    int nbad;
    int slot;
};

template <int MIC>
__global__ void __launch_bounds__(kConsumers + 32, 2) measure_fused_kernel(const FusedParams q) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kStagesMax;
    int* sbond = reinterpret_cast<int*>(smem + 128);
    unsigned int* shist = reinterpret_cast<unsigned int*>(sbond + ((q.max_slots + 3) & ~3));
    unsigned char* ecache_all = smem + q.head_bytes;
    unsigned char* stages = ecache_all + (size_t)kWarps * q.ecache_warp_bytes;
    const bool do_hist = q.hist != nullptr;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    const CgbMeasTile tile = q.tiles[blockIdx.x];
    const int ng = tile.ngroups;
    const int fl = kWarps / ng;                  // warps that share a group (frame lanes)
    const int nactive = ng * fl;

    if (tid == 0) {
        for (int s = 0; s < q.nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nactive);
        }
        fence_barrier_init();
    }
    for (int i = tid; i < q.max_slots; i += blockDim.x) sbond[i] = i < tile.nslots ? __ldg(q.tile_bonds + tile.slot0 + i) : -1;
    if (do_hist)
        for (int i = tid; i < q.max_slots * q.nbins; i += blockDim.x) shist[i] = 0u;
    __syncthreads();

    // frame blocks of k frames: CTA y works on blocks y * bpc .. (y + 1) * bpc - 1
    const int64_t nblocks = (q.F + q.k - 1) / q.k;
    const int64_t blk0 = (int64_t)blockIdx.y * q.blocks_per_cta;
    const int nst = (int)max((int64_t)0, min((int64_t)q.blocks_per_cta, nblocks - blk0));
    const size_t frame_fl = (size_t)q.B * 3;
    const float* const span0 = q.xyz + (size_t)tile.bead0 * 3;     // the tile's range in frame 0
    // every stage starts at a frame that is a multiple of 4, so its 16-byte head is the same for all stages
    const uint32_t h0 = (uint32_t)(reinterpret_cast<uintptr_t>(span0) & 15u);

    if (warp == kWarps) {
        // ------------------------------------------------------------------ producer warp
        const uintptr_t lo = reinterpret_cast<uintptr_t>(q.xyz);
        const uintptr_t hi = lo + (uintptr_t)q.F * (uintptr_t)q.B * 12u;
        const uint64_t policy = policy_evict_first();
        int s = 0;
        uint32_t phase = 1;  // parity of the previous use of the slot: the first pass through the ring never waits
        for (int it = 0; it < nst; ++it) {
            const int64_t f0 = (blk0 + it) * q.k;
            if (it >= q.nstage) mbar_wait(&empty[s], phase);
            const int s_now = s;
            if (++s == q.nstage) {
                s = 0;
                phase ^= 1u;
            }
            const int kk = (int)min((int64_t)q.k, q.F - f0);
            unsigned char* sbase = stages + (size_t)s_now * q.stage_bytes;
            uint64_t* const full_s = &full[s_now];
            const int ncopies = q.contiguous ? 1 : kk;
            const uint32_t nfl = q.contiguous ? (uint32_t)kk * (uint32_t)frame_fl : (uint32_t)(3 * tile.span);
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
                // float i of frame j lands at sbase + j * slot_bytes + h0 + 4 i: slot_bytes == 12 B (mod 16), so this
                // address has the low four bits of its source and the copy runs between 16-byte boundaries
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
            if (MIC != MIC_NONE) {
                // the frames' boxes, ready to use.  Orthorhombic: one 12-float block per pair of adjacent frames,
                // [-Lx_a -Lx_b -Ly_a -Ly_b | -Lz_a -Lz_b ix_a ix_b | iy_a iy_b iz_a iz_b] (i = 1 / L), so that three
                // 16-byte loads give the six packed operands.  Triclinic: 12 floats per frame, [a 0 | b 0 | c 0] reduced.
                float* tab = reinterpret_cast<float*>(sbase + q.tab_off);
                for (int j = lane; j < kk; j += 32) {
                    RawBox<MIC> rb;
                    fetch_box<MIC>(q.box, f0 + j, rb);
                    Box bx;
                    make_box<MIC>(rb, bx);
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
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const bool working = warp < nactive;
    const int gi = working ? warp % ng : 0;
    const int fli = warp / ng;                                      // this warp's frame lane
    const CgbMeasGroup grp = q.groups[tile.group0 + gi];
    unsigned char* const ecache = ecache_all + (size_t)warp * q.ecache_warp_bytes;   // [2 pairs][edges][32 B]
    const uint32_t pair_stride = (uint32_t)q.ecache_warp_bytes >> 1;

    Role role[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        Role& r = role[c];
        r.a0 = r.a1 = r.a2 = 0;
        r.g1 = r.g2 = 1.f;
        r.hist = shist;
        r.nshift = r.hnorm = r.hc = 0.f;
        r.s1 = r.s2 = r.c = r.s = 0;
        r.nbad = 0;
        r.slot = -1;
        const int kind = grp.kind[c];
        if (!working || kind == 0 || lane >= grp.nlanes[c]) continue;
        const CgbMeasLane ln = q.lanes[((size_t)(tile.group0 + gi) * kChunks + c) * 32 + lane];
        if (kind == 2) {
            r.a0 = (ln.ref & 0xffffu) * 12u + h0;
            r.a1 = (ln.ref >> 16) * 12u + h0;
        } else {
            const uint32_t e0 = ln.ref & 0x7fu, e1 = (ln.ref >> 8) & 0x7fu, e2 = (ln.ref >> 16) & 0x7fu;
            const float s0 = (ln.ref & 0x80u) ? -1.f : 1.f, s1 = (ln.ref & 0x8000u) ? -1.f : 1.f,
                        s2 = (ln.ref & 0x800000u) ? -1.f : 1.f;
            r.a0 = e0 * 32u;
            r.a1 = e1 * 32u;
            r.a2 = e2 * 32u;
            if (kind == 3) {
                r.g1 = s0 * s1;
            } else {
                r.g1 = s0 * s1 * s2;
                r.g2 = s0 * s2;
            }
        }
        r.slot = ln.slot;
        if (ln.slot >= 0) {
            const int bond = sbond[ln.slot];
            r.nshift = q.shift ? -__ldg(q.shift + bond) : 0.f;
            r.hist = shist + (size_t)ln.slot * q.nbins;
            if (do_hist) {
                const float lo = __ldg(q.hlo + bond), hi = __ldg(q.hhi + bond);
                r.hnorm = (float)q.nbins / (hi - lo);
                r.hc = fmaf(-lo, r.hnorm, -0.5f);
            }
        }
    }
    const bool store = q.values != nullptr;  // uniform
    int frames_done = 0;                     // frames this warp has evaluated (uniform)

    // Exact-rule histogram update of one value (next to a bin edge, out of range or NaN: ~0.02 % of the values).
    auto hist_exact = [&](const Role& r, float v) {
        const int bond = sbond[r.slot];
        hist_add_exact(r.hist, v, __ldg(q.hlo + bond), __ldg(q.hhi + bond), q.nbins);
    };
    // Statistics, histogram and stores of one evaluated pair (the common case: both frames valid).
    auto finish_fast = [&](Role& r, int kind, f2 v2, f2 cv2, f2 sv2, float* out) {
        const f2 nshift2 = splat2(r.nshift);
        const f2 d = kind == 4 ? wrap_about2(v2, nshift2) : add2(v2, nshift2);
        r.s1 = add2(r.s1, d);
        r.s2 = fma2(d, d, r.s2);
        if (kind != 2) {
            r.c = add2(r.c, cv2);
            r.s = add2(r.s, sv2);
        }
        if (do_hist) {
            const unsigned redo = hist_add_pair(r.hist, v2, splat2(r.hnorm), splat2(r.hc), (unsigned)q.nbins);
            if (redo) {
                if (redo & 1u) hist_exact(r, lo2(v2));
                if (redo & 2u) hist_exact(r, hi2(v2));
            }
        }
        if (store) {
            __stcs(out, lo2(v2));
            __stcs(out + q.row_stride, hi2(v2));
        }
    };
    // One value on its own: a frame next to the end of the trajectory or a degenerate / NaN operand.
    auto finish_one = [&](Role& r, int kind, int half, float v, float cv, float sv, float* out) {
        if (store) __stcs(out, v);
        if (!(v == v)) {  // NaN-skipping statistics (np.nanmean / np.nanvar, util.py:36,53)
            r.nbad += 1;
            return;
        }
        const f2 vv = half ? pack2(0.f, v) : pack2(v, 0.f);
        const f2 ns = half ? pack2(0.f, r.nshift) : pack2(r.nshift, 0.f);
        const f2 d = kind == 4 ? wrap_about2(vv, ns) : add2(vv, ns);
        r.s1 = add2(r.s1, d);
        r.s2 = fma2(d, d, r.s2);
        if (kind != 2) {
            r.c = add2(r.c, half ? pack2(0.f, cv) : pack2(cv, 0.f));
            r.s = add2(r.s, half ? pack2(0.f, sv) : pack2(sv, 0.f));
        }
        if (do_hist) {
            const int bin = hist_bin(v, __ldg(q.hlo + sbond[r.slot]), __ldg(q.hhi + sbond[r.slot]), r.hnorm, q.nbins);
            if (bin >= 0) atomicAdd(r.hist + bin, 1u);
        }
    };

    int s = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nst; ++it) {
        const int64_t f0 = (blk0 + it) * q.k;
        const int kk = (int)min((int64_t)q.k, q.F - f0);
        const unsigned char* stage = stages + (size_t)s * q.stage_bytes;
        const float* tab = reinterpret_cast<const float*>(stage + q.tab_off);
        if (working) {
            mbar_wait(&full[s], phase);
            const int nquads = (kk + 3) >> 2;
            for (int qd = fli; qd < nquads; qd += fl) {
                const int nvalid = min(4, kk - 4 * qd);                   // frames of this quad inside the trajectory
                const bool whole = nvalid == 4;
                frames_done += nvalid;
                const unsigned char* fq = stage + (size_t)(4 * qd) * q.slot_bytes;
                float* const vq = store ? q.values + (size_t)(f0 + 4 * qd) * q.row_stride : nullptr;

                // ---- edges: displacement records of both frame pairs, and the lengths
#pragma unroll
                for (int c = 0; c < CGB_MEAS_EDGE_CHUNKS; ++c) {
                    if (grp.kind[c] != 2) continue;
                    Role& r = role[c];
                    if (lane < grp.nlanes[c]) {
                        EdgeRec e[2];
#pragma unroll
                        for (int pp = 0; pp < 2; ++pp) {
                            const unsigned char* fa = fq + (size_t)(2 * pp) * q.slot_bytes;
                            const unsigned char* fb = fa + q.slot_bytes;
                            if (MIC == MIC_TRICLINIC) {
                                float ra[3], rb[3];
                                {
                                    Box bx;
                                    const float4* t = reinterpret_cast<const float4*>(tab + (4 * qd + 2 * pp) * 12);
                                    const float4 t0 = t[0], t1 = t[1], t2 = t[2];
                                    bx.a[0] = t0.x; bx.a[1] = t0.y; bx.a[2] = t0.z;
                                    bx.b[0] = t1.x; bx.b[1] = t1.y; bx.b[2] = t1.z;
                                    bx.c[0] = t2.x; bx.c[1] = t2.y; bx.c[2] = t2.z;
                                    const float* xa = reinterpret_cast<const float*>(fa + r.a0);
                                    const float* xb = reinterpret_cast<const float*>(fa + r.a1);
                                    const float pa[3] = {xa[0], xa[1], xa[2]}, pb[3] = {xb[0], xb[1], xb[2]};
                                    edge_triclinic(pa, pb, bx, ra);
                                }
                                {
                                    Box bx;
                                    const float4* t = reinterpret_cast<const float4*>(tab + (4 * qd + 2 * pp + 1) * 12);
                                    const float4 t0 = t[0], t1 = t[1], t2 = t[2];
                                    bx.a[0] = t0.x; bx.a[1] = t0.y; bx.a[2] = t0.z;
                                    bx.b[0] = t1.x; bx.b[1] = t1.y; bx.b[2] = t1.z;
                                    bx.c[0] = t2.x; bx.c[1] = t2.y; bx.c[2] = t2.z;
                                    const float* xa = reinterpret_cast<const float*>(fb + r.a0);
                                    const float* xb = reinterpret_cast<const float*>(fb + r.a1);
                                    const float pa[3] = {xa[0], xa[1], xa[2]}, pb[3] = {xb[0], xb[1], xb[2]};
                                    edge_triclinic(pa, pb, bx, rb);
                                }
#pragma unroll
                                for (int d = 0; d < 3; ++d) e[pp].r[d] = pack2(ra[d], rb[d]);
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
                                f2 nL[3] = {0, 0, 0}, inv[3] = {0, 0, 0};
                                if (MIC == MIC_ORTHO) {
                                    const ulonglong2* t = reinterpret_cast<const ulonglong2*>(tab + (2 * qd + pp) * 12);
                                    const ulonglong2 q0 = t[0], q1 = t[1], q2 = t[2];
                                    nL[0] = q0.x; nL[1] = q0.y; nL[2] = q1.x;
                                    inv[0] = q1.y; inv[1] = q2.x; inv[2] = q2.y;
                                }
                                edge_pair<(MIC == MIC_TRICLINIC ? MIC_NONE : MIC)>(xa, xb, nL, inv, e[pp]);
                            }
                            ulonglong2* rec = reinterpret_cast<ulonglong2*>(ecache + pp * pair_stride + (c * 32 + lane) * 32);
                            rec[0] = make_ulonglong2(e[pp].r[0], e[pp].r[1]);
                            rec[1] = make_ulonglong2(e[pp].r[2], e[pp].n2);
                        }
                        if (lane < grp.nemit[c]) {
                            float* out = vq + grp.vcol0[c] + lane;
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp) {
                                f2 v2;
                                const bool ok = length_pair(e[pp].n2, v2);
                                if (ok && whole) {
                                    finish_fast(r, 2, v2, 0, 0, out + (size_t)(2 * pp) * q.row_stride);
                                } else {
#pragma unroll
                                    for (int h = 0; h < 2; ++h)
                                        if (2 * pp + h < nvalid) {
                                            const float x = h ? hi2(e[pp].n2) : lo2(e[pp].n2);
                                            const float v = ok ? (h ? hi2(v2) : lo2(v2)) : __fsqrt_rn(x);
                                            finish_one(r, 2, h, v, 0.f, 0.f, out + (size_t)(2 * pp + h) * q.row_stride);
                                        }
                                }
                            }
                        }
                    }
                }
                __syncwarp();

                // ---- angles and dihedrals from the records
#pragma unroll
                for (int c = 1; c < kChunks; ++c) {
                    const int kind = grp.kind[c];
                    if (kind < 3) continue;
                    Role& r = role[c];
                    if (lane < grp.nlanes[c]) {
                        float* out = vq + grp.vcol0[c] + lane;
                        if (kind == 3) {
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp) {
                                EdgeRec u, v;
                                {
                                    const ulonglong2* ru = reinterpret_cast<const ulonglong2*>(ecache + pp * pair_stride + r.a0);
                                    const ulonglong2* rv = reinterpret_cast<const ulonglong2*>(ecache + pp * pair_stride + r.a1);
                                    const ulonglong2 u0 = ru[0], u1 = ru[1], v0 = rv[0], v1 = rv[1];
                                    u.r[0] = u0.x; u.r[1] = u0.y; u.r[2] = u1.x; u.n2 = u1.y;
                                    v.r[0] = v0.x; v.r[1] = v0.y; v.r[2] = v1.x; v.n2 = v1.y;
                                }
                                f2 v2, cv2, sv2, den;
                                const bool ok = angle_pair(u, v, r.g1, v2, cv2, sv2, den);
                                if (ok && whole) {
                                    finish_fast(r, 3, v2, cv2, sv2, out + (size_t)(2 * pp) * q.row_stride);
                                } else {
#pragma unroll
                                    for (int h = 0; h < 2; ++h)
                                        if (2 * pp + h < nvalid) {
                                            const float dn = h ? hi2(den) : lo2(den);
                                            const bool good = dn > 0.f && dn < 3e38f;
                                            const float val = good ? (h ? hi2(v2) : lo2(v2)) : __int_as_float(0x7fc00000);
                                            finish_one(r, 3, h, val, h ? hi2(cv2) : lo2(cv2), h ? hi2(sv2) : lo2(sv2),
                                                       out + (size_t)(2 * pp + h) * q.row_stride);
                                        }
                                }
                            }
                        } else {
#pragma unroll
                            for (int pp = 0; pp < 2; ++pp) {
                                EdgeRec b1, b2, b3;
                                {
                                    const ulonglong2* r1 = reinterpret_cast<const ulonglong2*>(ecache + pp * pair_stride + r.a0);
                                    const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(ecache + pp * pair_stride + r.a1);
                                    const ulonglong2* r3 = reinterpret_cast<const ulonglong2*>(ecache + pp * pair_stride + r.a2);
                                    const ulonglong2 x0 = r1[0], x1 = r1[1], y0 = r2[0], y1 = r2[1], z0 = r3[0], z1 = r3[1];
                                    b1.r[0] = x0.x; b1.r[1] = x0.y; b1.r[2] = x1.x; b1.n2 = x1.y;
                                    b2.r[0] = y0.x; b2.r[1] = y0.y; b2.r[2] = y1.x; b2.n2 = y1.y;
                                    b3.r[0] = z0.x; b3.r[1] = z0.y; b3.r[2] = z1.x; b3.n2 = z1.y;
                                }
                                f2 v2, cv2, sv2, p1, p2;
                                const bool ok = dihedral_pair(b1, b2, b3, r.g1, r.g2, v2, cv2, sv2, p1, p2);
                                if (ok && whole) {
                                    finish_fast(r, 4, v2, cv2, sv2, out + (size_t)(2 * pp) * q.row_stride);
                                } else {
#pragma unroll
                                    for (int h = 0; h < 2; ++h)
                                        if (2 * pp + h < nvalid) {
                                            const float x1 = h ? hi2(p1) : lo2(p1), x2 = h ? hi2(p2) : lo2(p2);
                                            const float rr = fmaf(x1, x1, x2 * x2);
                                            float cv = h ? hi2(cv2) : lo2(cv2), sv = h ? hi2(sv2) : lo2(sv2);
                                            float val = h ? hi2(v2) : lo2(v2);
                                            if (!(rr > 0.f && rr < 3e38f)) val = dihedral_degenerate(x1, x2, cv, sv);
                                            finish_one(r, 4, h, val, cv, sv, out + (size_t)(2 * pp + h) * q.row_stride);
                                        }
                                }
                            }
                        }
                    }
                }
                __syncwarp();  // the records are rewritten by the next quad
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (++s == q.nstage) {
            s = 0;
            phase ^= 1u;
        }
    }

    // ---------------------------------------------------------------------- CTA epilogue
    // Per-lane float32 sums -> per-Bond float64 sums of the CTA in a fixed order (no floating-point atomics inside
    // the CTA, so its contribution does not depend on scheduling); CTAs are then combined with float64 atomics.
    // The stage memory is reused: every stage this CTA was sent has been consumed by every warp by now.
    double* tacc = reinterpret_cast<double*>(stages);                       // [kEpiAcc][kConsumers]
    int* tslot = reinterpret_cast<int*>(tacc + (size_t)kEpiAcc * kConsumers);
    const int nwarp_slots = kWarps;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");
        const Role& r = role[c];
        const bool mine = working && r.slot >= 0 && frames_done > 0;
        tslot[tid] = mine ? r.slot : -1;
        if (mine) {
            float a, b;
            tacc[0 * kConsumers + tid] = (double)(frames_done - r.nbad);
            unpack2(r.s1, a, b);
            tacc[1 * kConsumers + tid] = (double)a + (double)b;
            unpack2(r.s2, a, b);
            tacc[2 * kConsumers + tid] = (double)a + (double)b;
            unpack2(r.c, a, b);
            tacc[3 * kConsumers + tid] = (double)a + (double)b;
            unpack2(r.s, a, b);
            tacc[4 * kConsumers + tid] = (double)a + (double)b;
            tacc[5 * kConsumers + tid] = (double)frames_done;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");
        // one warp per Bond slot: lane-strided sums over the threads of the slot, then a fixed butterfly
        for (int sl = warp; sl < tile.nslots; sl += nwarp_slots) {
            double a[kEpiAcc] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            bool any = false;
            for (int t = lane; t < kConsumers; t += 32)
                if (tslot[t] == sl) {
                    any = true;
#pragma unroll
                    for (int k = 0; k < kEpiAcc; ++k) a[k] += tacc[k * kConsumers + t];
                }
            if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < kEpiAcc; ++k) a[k] += __shfl_down_sync(0xffffffffu, a[k], o);
            }
            if (lane == 0) {
                double* dst = q.partials + (size_t)sbond[sl] * CGB_NPART;
                const int where[kEpiAcc] = {0, 1, 2, 3, 4, 6};
#pragma unroll
                for (int k = 0; k < kEpiAcc; ++k)
                    if (a[k] != 0.0) atomicAdd(dst + where[k], a[k]);
            }
        }
    }
    if (do_hist) {
        asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");
        for (int i = tid; i < tile.nslots * q.nbins; i += kConsumers) {
            const unsigned int v = shist[i];
            if (v) atomicAdd(q.hist + (size_t)sbond[i / q.nbins] * q.nbins + (i % q.nbins), (double)v);
        }
    }
}

}  // namespace cgb

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" int cgb_measure_plan_size(const int32_t* col_beads, const int32_t* col_bond, int64_t n2, int64_t n3,
                                     int64_t n4, int64_t n_beads, int32_t max_span, int32_t max_slots,
                                     CgbMeasPlanInfo* info) {
    if (!info) return CGB_EINVAL;
    cgb::plan::Compiler cc{col_beads, col_bond, n2, n3, n4, n_beads, max_span, max_slots, nullptr, {}, {}};
    const int rc = cc.run();
    if (rc != CGB_OK) return rc;
    info->n_tiles = (int64_t)cc.out.tiles.size();
    info->n_groups = (int64_t)cc.out.groups.size();
    info->n_tile_bonds = (int64_t)cc.out.tile_bonds.size();
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
                                      CgbMeasTile* tiles, CgbMeasGroup* groups, CgbMeasLane* lanes,
                                      int32_t* tile_bonds, int32_t* col_out, int64_t* wide_cols) {
    const int64_t M = n2 + n3 + n4;
    if (M > 0 && !col_out) return CGB_EINVAL;
    cgb::plan::Compiler cc{col_beads, col_bond, n2, n3, n4, n_beads, max_span, max_slots, col_out, {}, {}};
    const int rc = cc.run();
    if (rc != CGB_OK) return rc;
    const auto& o = cc.out;
    if ((!o.tiles.empty() && (!tiles || !groups || !lanes || !tile_bonds)) || (!o.wide.empty() && !wide_cols))
        return CGB_EINVAL;
    if (!o.tiles.empty()) {
        std::memcpy(tiles, o.tiles.data(), o.tiles.size() * sizeof(CgbMeasTile));
        std::memcpy(groups, o.groups.data(), o.groups.size() * sizeof(CgbMeasGroup));
        std::memcpy(lanes, o.lanes.data(), o.lanes.size() * sizeof(CgbMeasLane));
        std::memcpy(tile_bonds, o.tile_bonds.data(), o.tile_bonds.size() * sizeof(int32_t));
    }
    if (!o.wide.empty()) std::memcpy(wide_cols, o.wide.data(), o.wide.size() * sizeof(int64_t));
    return CGB_OK;
}

namespace cgb {

// Shared-memory layout and grid of one launch.
static int fused_geometry(FusedParams& q, const CgbMeasureTuning* tune, int64_t n_tiles, int32_t max_span,
                          int32_t max_edge_chunks, int32_t max_tile_groups, int sm_count, dim3* grid,
                          size_t* smem_bytes) {
    const int min_fl = kWarps / std::max(1, std::min<int>(max_tile_groups, kWarps));   // frame lanes of the fullest tile
    const int kCtaSmem = 112 * 1024;                         // two CTAs per SM
    q.nstage = tune && tune->n_stages ? tune->n_stages : 3;
    if (q.nstage < 2 || q.nstage > kStagesMax) return CGB_EINVAL;
    const size_t hist_bytes = q.hist ? (size_t)q.max_slots * q.nbins * 4 : 0;
    q.head_bytes = (int)((128 + (size_t)((q.max_slots + 3) & ~3) * 4 + hist_bytes + 127) & ~(size_t)127);
    q.ecache_warp_bytes = 2 * std::max(1, max_edge_chunks) * 32 * 32;
    const size_t fixed = (size_t)q.head_bytes + (size_t)kWarps * q.ecache_warp_bytes;
    const size_t epilogue = (size_t)kConsumers * (kEpiAcc * 8 + 4);
    const int64_t frame_bytes = q.B * 12;
    q.contiguous = (n_tiles == 1 && max_span == q.B) ? 1 : 0;
    int64_t slot;
    if (q.contiguous) {
        slot = frame_bytes;
    } else {
        slot = (int64_t)max_span * 12 + 48;
        slot += ((frame_bytes - slot) % 16 + 16) % 16;       // == 12 B (mod 16): see the producer
    }
    if (fixed + epilogue > (size_t)kCtaSmem) return -1;
    // frames per stage: whole quads, as many as fit the stage budget (12 bytes of box table per frame); the pipeline
    // is three stages deep unless only two stages of one quad of the widest tile fit
    const int64_t quad = (slot + 12) * 4 + 64 + 128;
    if (!(tune && tune->n_stages) && quad * q.nstage > (int64_t)kCtaSmem - (int64_t)fixed) q.nstage = 2;
    const int64_t room = ((int64_t)kCtaSmem - (int64_t)fixed) / q.nstage - 128;
    if (quad > room + 128) return -1;                          // one quad of the widest tile does not fit
    int64_t want = tune && tune->stage_bytes ? tune->stage_bytes : 16 * 1024;
    want = std::min(want, room);
    int64_t k = (want - 64) / (slot + 12);
    k -= k % 4;
    if (k >= 4 * min_fl) k -= k % (4 * min_fl);              // whole rounds of quads over the frame lanes
    if (k < 4) k = 4;
    k = std::min<int64_t>(k, 64);
    k = std::min<int64_t>(k, std::max<int64_t>(4, (q.F + 3) / 4 * 4));
    q.k = (int)k;
    q.slot_bytes = (int)slot;
    q.tab_off = (int)((k * slot + 32 + 15) & ~15LL);
    q.stage_bytes = (int)((q.tab_off + 12 * 4 * k + 127) & ~127LL);
    *smem_bytes = fixed + std::max((size_t)q.nstage * q.stage_bytes, epilogue);
    if (*smem_bytes > (size_t)kCtaSmem) return -1;
    // frames per CTA: at most `max_run` frames per lane (the statistics stay in float32 registers for the life of
    // a CTA), and about two waves of CTAs when the trajectory allows it
    const int64_t nblocks = (q.F + k - 1) / k;
    const int64_t max_run = tune && tune->max_run ? tune->max_run : 512;
    const int64_t cap = std::max<int64_t>(1, max_run * min_fl / k);   // a lane sees k / fl frames per block
    const int64_t waves = tune && tune->waves ? tune->waves : 2;
    int64_t ny = std::max<int64_t>(1, ((int64_t)sm_count * 2 * waves + n_tiles - 1) / n_tiles);
    ny = std::min(ny, nblocks);
    int64_t bpc = (nblocks + ny - 1) / ny;
    bpc = std::min(bpc, cap);
    ny = (nblocks + bpc - 1) / bpc;
    if (ny > 65535 || n_tiles > INT32_MAX) return CGB_ETOOBIG;
    q.blocks_per_cta = (int)bpc;
    *grid = dim3((unsigned)n_tiles, (unsigned)ny);
    return CGB_OK;
}

}  // namespace cgb

extern "C" int cgb_measure_fused(const float* cg_xyz, const float* box_vec, int ortho, int64_t n_frames,
                                 int64_t n_beads, const CgbMeasTile* tiles, int64_t n_tiles,
                                 const CgbMeasGroup* groups, const CgbMeasLane* lanes, const int32_t* tile_bonds,
                                 int32_t max_span, int32_t max_slots, int32_t max_edge_chunks,
                                 int32_t max_tile_groups, const float* shift, float* values, int64_t row_stride,
                                 double* partials, double* hist,
                                 const float* hist_lo, const float* hist_hi, int32_t n_bins,
                                 const CgbMeasureTuning* tuning, void* stream) {
    using namespace cgb;
    if (n_frames < 0 || n_beads < 0 || n_tiles < 0) return CGB_EINVAL;
    if (n_frames == 0 || n_tiles == 0) return CGB_OK;
    if (!cg_xyz || !tiles || !groups || !lanes || !tile_bonds || !partials) return CGB_EINVAL;
    if (hist && (!hist_lo || !hist_hi || n_bins < 1 || n_bins > 512)) return CGB_EINVAL;  // fast-bin error bound
    if (values && row_stride < 1) return CGB_EINVAL;
    if (max_span < 1 || max_slots < 1 || max_edge_chunks < 1 || max_edge_chunks > CGB_MEAS_EDGE_CHUNKS) return CGB_EINVAL;
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
    q.shift = shift;
    q.hlo = hist_lo;
    q.hhi = hist_hi;
    q.values = values;
    q.row_stride = row_stride;
    q.partials = partials;
    q.hist = hist;
    q.nbins = hist ? n_bins : 0;
    q.max_slots = max_slots;
    int dev = 0, sms = 0;
    CGB_CUDA_TRY(cudaGetDevice(&dev));
    CGB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    dim3 grid;
    size_t smem = 0;
    const int rc = fused_geometry(q, tuning, n_tiles, max_span, max_edge_chunks, max_tile_groups, sms, &grid, &smem);
    if (rc == -1) return CGB_ETOOBIG;
    if (rc != CGB_OK) return rc;
    const int mic = box_vec ? (ortho ? MIC_ORTHO : MIC_TRICLINIC) : MIC_NONE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define CGB_FUSED(M_)                                                                                    \
    do {                                                                                                 \
        auto kern = measure_fused_kernel<M_>;                                                            \
        CGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, kConsumers + 32, smem, st>>>(q);                                                    \
        return (int)cudaGetLastError();                                                                  \
    } while (0)
    if (mic == MIC_ORTHO) CGB_FUSED(MIC_ORTHO);
    if (mic == MIC_TRICLINIC) CGB_FUSED(MIC_TRICLINIC);
    CGB_FUSED(MIC_NONE);
#undef CGB_FUSED
}

/* The largest bead range (`max_span`) and number of Bonds (`max_slots`) a tile may have for the given histogram
 * size: what the plan compiler should be called with. */
extern "C" int cgb_measure_plan_limits(int32_t n_bins, int with_hist, int32_t* max_span, int32_t* max_slots) {
    if (!max_span || !max_slots || n_bins < 0 || n_bins > 512) return CGB_EINVAL;
    // shared memory of a CTA (112 KB): head (Bond table + histograms), 7 x 4 KB edge records, three stages of
    // at least one quad.  Histograms get at most 40 KB; a quad of the widest tile must fit a 16 KB stage.
    const int per_slot = 4 + (with_hist ? 4 * n_bins : 0);
    *max_slots = std::max(8, std::min(1024, 40 * 1024 / per_slot));
    *max_span = 512;                              // a quad of 512 beads is a 24 KB stage
    return CGB_OK;
}
