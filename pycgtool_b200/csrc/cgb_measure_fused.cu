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

#include "cgb_termwarp.cuh"
#include "cgb_measure_plan.cpp.inc"

namespace cgb {

constexpr int kWarps = kTermWarps;                // consumer warps per CTA
constexpr int kConsumers = kTermThreads;
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
    int32_t fold;          // real frames per staged frame (frame folding, >= 1); fold_beads = beads per real frame
    int32_t fold_beads;
    int32_t ecache_warp_bytes;
    // shared-memory map (byte offsets)
    int32_t off_sparam, off_shist, off_lists, off_eacc, off_nbad, off_roles, off_ecache, off_stages;
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

template <int MIC, int SHAPE, int MODE>   // MODE: bit 0 = the values are stored, bit 1 = histograms
__global__ void __launch_bounds__(kConsumers + 32, 2) measure_fused_kernel(const FusedParams q) {
    static_assert(MIC == MIC_ORTHO || MIC == MIC_TRICLINIC, "no box = orthorhombic path with a zero box table");
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kStagesMax;
    unsigned int* scratch = reinterpret_cast<unsigned int*>(smem + 64);   // scratch words (hist_add_pair_pred)
    int* sbond = reinterpret_cast<int*>(smem + 192);
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
                    // with frame folding a staged frame holds q.fold real frames, each with its own box: entry
                    // (staged frame j, instance r) is the box of real frame (f0 + j) * fold + r.  Up to four entries
                    // per lane and round, their global loads all in flight before the first is used.
                    const int nent = kk * q.fold;
                    for (int i0 = 0; i0 < nent; i0 += 128) {
                        RawBox<MIC> rb[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + 32 * u + lane;
                            if (q.box && i < nent) fetch_box<MIC>(q.box, f0 * q.fold + i, rb[u]);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + 32 * u + lane;
                            if (i >= nent) continue;
                            const int j = i / q.fold, r = i - j * q.fold;
                            Box bx;
                            if (q.box) {
                                make_box<MIC>(rb[u], bx);
                            } else {
#pragma unroll
                                for (int d = 0; d < 3; ++d) bx.nL[d] = bx.inv[d] = 0.f;
                            }
                            if (MIC == MIC_ORTHO) {
                                float* t = tab + ((j >> 1) * q.fold + r) * 12 + (j & 1);
#pragma unroll
                                for (int d = 0; d < 3; ++d) {
                                    t[2 * d] = bx.nL[d];
                                    t[6 + 2 * d] = bx.inv[d];
                                }
                            } else {
                                float* t = tab + (j * q.fold + r) * 12;
#pragma unroll
                                for (int d = 0; d < 3; ++d) {
                                    t[d] = bx.a[d];
                                    t[4 + d] = bx.b[d];
                                    t[8 + d] = bx.c[d];
                                }
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
    const uint32_t table32 = smem_u32(smem + q.off_roles) + (uint32_t)warp * (uint32_t)kRoleTableBytes;
    TermTables tabs;
    tabs.sbond = sbond;
    tabs.sparam = sparam;
    tabs.shist = shist;
    tabs.slist = slist;
    tabs.snbad = snbad;
    tabs.eacc = eacc;
    tabs.ecnt = ecnt;
    tabs.shist32 = smem_u32(shist);
    tabs.dummy32 = smem_u32(scratch);
    tabs.rs32 = (uint32_t)q.row_stride;
    tabs.shift = q.shift;
    tabs.hlo = q.hlo;
    tabs.hhi = q.hhi;
    tabs.values = q.values;
    tabs.row_stride = q.row_stride;
    tabs.partials = q.partials;
    tabs.hist = q.hist;
    tabs.nbins = q.nbins;
    tabs.hstride = hist_stride(q.nbins);
    tabs.fold = q.fold;
    constexpr bool store = (MODE & 1) != 0;
    const int64_t row_stride = q.row_stride;
    TermWarp<MIC, SHAPE> tw;
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
        const uint32_t h0 = (uint32_t)(reinterpret_cast<uintptr_t>(q.xyz + (size_t)tile.bead0 * 3) & 15u);
        term_tables_fill(tabs, tile, q.tile_bonds, q.tile_lists, tid, kConsumers);
        tw.setup(working, q.groups[tile.group0 + gi], q.lanes + (size_t)(tile.group0 + gi) * kChunks * 32, lane, h0,
                 ecache32, table32, tid, tile.bead0, q.fold > 1 ? q.fold_beads : 0);
        asm volatile("bar.sync 1, %0;" ::"r"(kConsumers) : "memory");

        // ---- main loop over the segment's frame blocks (32-bit counters and running pointers: this code runs once
        // per stage and warp)
        {
            const int64_t f_seg = seg_block0 * q.k;
            int frames_left = (int)min(q.F - f_seg, (int64_t)0x7fffffff);
            float* vstage = store ? q.values + (size_t)f_seg * row_stride : nullptr;
            const size_t vstage_step = (size_t)q.k * row_stride;
            const uint32_t quad_bytes = 4u * (uint32_t)q.slot_bytes;
            const uint32_t tq_quad = (MIC == MIC_ORTHO ? 96u : 192u) * (uint32_t)q.fold;
            const size_t vq_quad = (size_t)4 * row_stride;
            const unsigned char* stage = stages + (size_t)s * q.stage_bytes;
            for (int it = 0; it < seg_count; ++it) {
                const int kk = min(q.k, frames_left);
                frames_left -= q.k;
                mbar_wait(&full[s], phase);
                if (working) {
                    // the warps that share a group take the quads of the stage in turn (a fixed assignment: the float32
                    // sums of a lane, hence the statistics, do not depend on scheduling)
                    const int nquads = (kk + 3) >> 2;
                    for (int qd = fli; qd < nquads; qd += fl)
                        tw.template quad_mode<MODE>(tabs, lane, stage + (uint32_t)qd * quad_bytes, (uint32_t)q.slot_bytes,
                                reinterpret_cast<const float*>(stage + (uint32_t)q.tab_off + (uint32_t)qd * tq_quad),
                                vstage + (size_t)qd * vq_quad, min(4, kk - 4 * qd));
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
        tw.epilogue(tabs, tile, working, warp, lane, kWarps, ng, fl, 1);
    }
}

}  // namespace cgb

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" int cgb_measure_plan_size(const int32_t* col_beads, const int32_t* col_bond, int64_t n2, int64_t n3,
                                     int64_t n4, int64_t n_beads, int32_t max_span, int32_t max_slots,
                                     int32_t shape, int32_t unit_groups, CgbMeasPlanInfo* info) {
    if (!info || unit_groups < 0 || unit_groups > CGB_MEAS_WARPS) return CGB_EINVAL;
    cgb::plan::Compiler cc{col_beads, col_bond, n2, n3, n4, n_beads, max_span, max_slots, shape, unit_groups, nullptr,
                           nullptr, {}, {}};
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
                                      int32_t shape, int32_t unit_groups, CgbMeasTile* tiles, CgbMeasGroup* groups,
                                      CgbMeasLane* lanes, int32_t* tile_bonds, uint16_t* tile_lists, int32_t* col_out,
                                      int64_t* wide_cols, uint8_t* tile_start) {
    const int64_t M = n2 + n3 + n4;
    if (M > 0 && !col_out) return CGB_EINVAL;
    if (unit_groups < 0 || unit_groups > CGB_MEAS_WARPS) return CGB_EINVAL;
    if (tile_start) std::memset(tile_start, 0, (size_t)std::max<int64_t>(n_beads, 0));
    cgb::plan::Compiler cc{col_beads, col_bond, n2, n3, n4, n_beads, max_span, max_slots, shape, unit_groups, tile_start,
                           col_out, {}, {}};
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
    size_t off = 192 + (size_t)((q.max_slots + 3) & ~3) * 4;
    off = up(off, 16);
    q.off_sparam = (int)off;
    off += (size_t)q.max_slots * 16;
    q.off_shist = (int)off;
    off += q.hist ? (size_t)q.max_slots * hist_stride(q.nbins) * 4 : 0;
    q.off_lists = (int)off;
    off = up(off + 2 * ((size_t)q.max_slots + 1 + kRoles), 16);
    q.off_eacc = 0;
    q.off_nbad = (int)off;
    off = up(off + (size_t)kRoles * 4, 128);
    q.off_roles = (int)off;
    off += (size_t)kWarps * kRoleTableBytes;
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
    const int64_t tab_frame = 48 * (int64_t)std::max(1, q.fold);     // box-table bytes per staged frame
    const int64_t quad = (slot + tab_frame) * 4 + 64 + 128;
    if ((int64_t)fixed + 2 * quad > kCtaSmem) return -1;    // two stages of one quad of the widest tile do not fit
    while (q.nstage > 2 && quad * q.nstage > (int64_t)kCtaSmem - (int64_t)fixed) --q.nstage;
    int64_t want = tune && tune->stage_bytes ? tune->stage_bytes : 20 * 1024;
    int64_t k = 4;
    for (;;) {
        const int64_t room = ((int64_t)kCtaSmem - (int64_t)fixed) / q.nstage - 128;
        k = (std::min(want, room) - 64) / (slot + tab_frame);
        k -= k % 4;
        if (k >= 4 * min_fl) k -= k % (4 * min_fl);              // whole rounds of quads over the frame lanes
        if (k < 4) k = 4;
        k = std::min<int64_t>(k, 64);
        k = std::min<int64_t>(k, std::max<int64_t>(4, (q.F + 3) / 4 * 4));
        q.k = (int)k;
        q.slot_bytes = (int)slot;
        q.tab_off = (int)((k * slot + 32 + 15) & ~15LL);
        q.stage_bytes = (int)((q.tab_off + tab_frame * k + 127) & ~127LL);
        *smem_bytes = fixed + (size_t)q.nstage * q.stage_bytes;
        if (*smem_bytes <= (size_t)kCtaSmem) break;
        if (k > 4)
            want = (k - 4) * (slot + tab_frame) + 64;            // one quad less per stage
        else if (q.nstage > 2)
            --q.nstage;
        else
            return -1;
    }
    // persistent CTAs: the (tile, frame block) items are dealt out in equal contiguous shares; a CTA reduces its
    // float32 sums (segment epilogue) at every change of tile and after at most `max_run` frames per lane
    q.nblocks = (q.F + k - 1) / k;
    // (bounded: a segment's 32-bit shared-memory histogram bins and float32 sums must not be overrun)
    const int64_t max_run = std::min<int64_t>(tune && tune->max_run ? tune->max_run : 512, 1 << 20);
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
                                 int32_t max_edge_chunks, int32_t max_tile_groups, int32_t fold, const float* shift,
                                 float* values, int64_t row_stride, double* partials, double* hist,
                                 const float* hist_lo, const float* hist_hi, int32_t n_bins,
                                 const CgbMeasureTuning* tuning, void* stream) {
    using namespace cgb;
    if (n_frames < 0 || n_beads < 0 || n_tiles < 0 || fold < 1 || fold > 64 || n_beads % fold) return CGB_EINVAL;
    if (n_frames == 0 || n_tiles == 0) return CGB_OK;
    if (!cg_xyz || !tiles || !groups || !lanes || !tile_bonds || !tile_lists || !partials) return CGB_EINVAL;
    if (hist && (!hist_lo || !hist_hi || n_bins < 1 || n_bins > 512)) return CGB_EINVAL;  // fast-bin error bound
    if (values && (row_stride < 1 || row_stride > 0x3fffffffLL)) return CGB_EINVAL;
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
    q.fold = fold;
    q.fold_beads = (int32_t)(n_beads / fold);
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
#define CGB_FUSED_MODE(M_, S_, MODE_)                                                                     \
    do {                                                                                                 \
        auto kern = measure_fused_kernel<M_, S_, MODE_>;                                                 \
        CGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, kConsumers + 32, smem, st>>>(q);                                                    \
        return (int)cudaGetLastError();                                                                  \
    } while (0)
#define CGB_FUSED(M_, S_)                         \
    do {                                          \
        if (mode == 3) CGB_FUSED_MODE(M_, S_, 3); \
        if (mode == 2) CGB_FUSED_MODE(M_, S_, 2); \
        if (mode == 1) CGB_FUSED_MODE(M_, S_, 1); \
        CGB_FUSED_MODE(M_, S_, 0);                \
    } while (0)
    const int mode = (values ? 1 : 0) | (hist ? 2 : 0);
    if (triclinic && shape == CGB_SHAPE_EAD) CGB_FUSED(MIC_TRICLINIC, CGB_SHAPE_EAD);
    if (triclinic) CGB_FUSED(MIC_TRICLINIC, CGB_SHAPE_EEAA);
    if (shape == CGB_SHAPE_EAD) CGB_FUSED(MIC_ORTHO, CGB_SHAPE_EAD);
    CGB_FUSED(MIC_ORTHO, CGB_SHAPE_EEAA);
#undef CGB_FUSED_MODE
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
    q.fold = 1;
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
    const int64_t per_slot = 22 + (with_hist ? 4 * (int64_t)cgb::hist_stride(n_bins) : 0);
    // everything that does not depend on the plan: barriers, lists, reduction scratch, NaN counters, 7 x 4 KB records
    const int64_t fixed = 192 + 2 * (1 + cgb::kRoles) + 16 + cgb::kRoles * 4 + 128 + cgb::kWarps * cgb::kRoleTableBytes +
                          (int64_t)cgb::kWarps * 2 * CGB_MEAS_EDGE_CHUNKS * 32 * 32 + 64;
    auto stages_for = [](int64_t span) { return 2 * (4 * (12 * span + 48 + 16 + 12) + 64 + 128 + 128); };
    const int64_t min_span = 64, best_span = 544;
    int64_t slots = std::min<int64_t>(want_slots, (budget - fixed - stages_for(min_span)) / per_slot);
    slots = std::max<int64_t>(slots, 1);
    int64_t span = best_span;
    while (span > min_span && fixed + slots * per_slot + stages_for(span) > budget) span -= 8;
    *max_slots = (int32_t)slots;
    *max_span = (int32_t)span;
    return CGB_OK;
}
