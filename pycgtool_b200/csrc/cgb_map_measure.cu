// Stages 1 + 2 in one kernel: mapping warps hand every mapped tile, still in shared memory, to measurement warps.
//
// Reference: the call sequence of PyCGTOOL.apply_mapping / measure_bonds (pycgtool/__main__.py:69-112):
// Mapping.apply (mapping.py:391-443) writes the CG frame, BondSet.apply (bondset.py:496-531) reads it back.  Here
// the CG frame is written to HBM once (it is an output: the reference saves it as .gro / .xtc) and never re-read:
//   * the producer warp and the mapping warps are K1 (cgb_mapwarp.cuh): a tile of output beads, four frames per
//     stage, results streamed to cg_xyz -- and, in the same instruction stream, into a CG tile in shared memory;
//   * two measurement warps (cgb_termwarp.cuh, the warps of K2) take the CG tile from there: edges, lengths,
//     angles, dihedrals, statistics, histograms, optional values -- per quad of frames;
//   * hand-over through a ring of three CG-tile slots and two mbarriers per slot (filled by the mapping warps,
//     drained by the measurement warps), so the two halves overlap.
// The plan makes this possible: measurement tiles are runs of whole units (molecule instances) and the mapping
// plan adopts them as its own tiles (cgb_measure_plan_build with unit_groups, cgb_map_plan_build_cuts).  Tiles
// without bonded terms (solvent) run as plain K1 tiles.  Systems the plan cannot express this way -- a unit wider
// than a tile, float64 weights, measured virtual beads -- take the separate K1 and K2 launches instead.

#include <algorithm>
#include <cstring>

#include "cgb_mapwarp.cuh"
#include "cgb_termwarp.cuh"

namespace cgb {

constexpr int kMMWarps = 2;      // measurement warps per CTA
constexpr int kMMSlots = 3;      // CG-tile slots between the mapping and the measurement warps
constexpr int kMMFG = 4;         // frames per stage = one quad

struct MMParams {
    const int32_t* meas_of_tile;  // [mapping tiles]: measurement tile of the same bead range, or -1
    const CgbMeasTile* mtiles;
    const CgbMeasGroup* groups;
    const CgbMeasLane* lanes;
    const int32_t* tile_bonds;
    const uint16_t* tile_lists;
    const float* box_vec;         // [F][9] or NULL
    const float* shift;
    const float* hlo;
    const float* hhi;
    float* values;
    int64_t row_stride;
    double* partials;
    double* hist;
    int32_t nbins;
    int32_t max_slots;
    int32_t ncw;                  // mapping warps
    int32_t cg_frame_bytes;       // bytes between the frames of a CG-tile slot (multiple of 16)
    int32_t cg_slot_bytes;        // one slot: four frames + their box table
    // shared-memory map of the measurement region (byte offsets from the start of shared memory)
    int32_t off_bars, off_sbond, off_sparam, off_shist, off_lists, off_nbad, off_eacc, off_roles, off_ecache, off_cg;
};

template <bool HAS_BOX, int MIC, int SHAPE>
__global__ void __launch_bounds__(256, 2) map_measure_kernel(const MapParams p, const MMParams m) {
    constexpr int FG = kMMFG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    CgbMapEntry* ents = reinterpret_cast<CgbMapEntry*>(smem_raw + 128);
    float* stages = reinterpret_cast<float*>(smem_raw + 128 + (size_t)p.ent_cap * sizeof(CgbMapEntry));
    uint64_t* cgfull = reinterpret_cast<uint64_t*>(smem_raw + m.off_bars);
    uint64_t* cgempty = cgfull + kMMSlots;
    unsigned char* cgbuf = smem_raw + m.off_cg;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ncw = m.ncw;
    const CgbMapTile tile = p.tiles[blockIdx.x];
    const int mt = __ldg(m.meas_of_tile + blockIdx.x);
    const bool has_meas = mt >= 0;
    CgbMeasTile mtile;
    if (has_meas) mtile = m.mtiles[mt];
    const int ng = has_meas ? mtile.ngroups : 1;
    const int64_t fbeg = (int64_t)blockIdx.y * p.fc;
    const int64_t fend = min(p.F, fbeg + p.fc);
    const int nst = (int)((fend - fbeg + p.k - 1) / p.k);
    const int nent = tile.nslots * tile.nbeads;

    if (tid == 0) {
        for (int s = 0; s < p.nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ncw);
        }
        for (int s = 0; s < kMMSlots; ++s) {
            mbar_init(&cgfull[s], ncw);    // every mapping warp arrives when its part of the CG tile is written
            mbar_init(&cgempty[s], ng);    // one measurement warp per group drains a slot
        }
        fence_barrier_init();
    }
    for (int i = tid; i < nent; i += blockDim.x) ents[i] = p.ent[tile.ent0 + i];
    __syncthreads();

    const int span_fl = 3 * tile.span;
    const int frame_fl = p.slot_floats;

    if (warp == ncw + kMMWarps) {
        // ------------------------------------------------------------------ producer warp (as in map_tiles_kernel)
        const uintptr_t lo = reinterpret_cast<uintptr_t>(p.xyz);
        const uintptr_t hi = lo + (uintptr_t)p.F * (uintptr_t)p.N * 12u;
        const uint64_t policy = policy_evict_first();
        for (int it = 0; it < nst; ++it) {
            const int s = it % p.nstage;
            if (it >= p.nstage) mbar_wait(&empty[s], ((it / p.nstage) - 1) & 1);
            const int64_t f0 = fbeg + (int64_t)it * p.k;
            const int kk = (int)min((int64_t)p.k, fend - f0);
            float* sbase = stages + (size_t)s * p.stage_floats;
            uint32_t bytes = 0;
            bool edge = false;
            const float* src = nullptr;
            if (lane < kk) {
                src = p.xyz + ((size_t)(f0 + lane) * p.N + (size_t)(tile.atom0 - p.base)) * 3;
                bytes = span_bytes(src, (uint32_t)span_fl, lo, hi, &edge);
            }
            unsigned edges = __ballot_sync(0xffffffffu, edge);
            while (edges) {
                const int j = __ffs(edges) - 1;
                edges &= edges - 1;
                const float* esrc = reinterpret_cast<const float*>(
                    __shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(src), j));
                float* dst = sbase + (size_t)j * p.slot_floats + head_floats(esrc);
                for (uint32_t i = lane; i < (uint32_t)span_fl; i += 32) dst[i] = __ldg(esrc + i);
            }
            if (HAS_BOX) {
                float2* tab = reinterpret_cast<float2*>(sbase + p.tab_off);
                for (int i = lane; i < kk * 3; i += 32) {
                    const float La = fabsf(__ldg(p.box + f0 * 3 + i));
                    tab[i] = make_float2(La, box_t05(La));
                }
            }
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&full[s], total);
            __syncwarp();
            if (lane < kk && !edge) {
                const void* a0 = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15));
                bulk_g2s_stream(sbase + (size_t)lane * p.slot_floats, a0, bytes, &full[s], policy);
            }
        }
        return;
    }

    if (warp < ncw) {
        // ------------------------------------------------------------------ mapping warps
        const int nchunks = (3 * tile.nbeads + 31) >> 5;
        const bool resident = nchunks <= ncw;
        LaneConst lc = lane_const(tile, ents, nullptr, warp < nchunks ? warp : 0, lane);
        const uint32_t hstep = (uint32_t)((12ull * (uint64_t)p.N) & 15ull);
        const int cg_stride = m.cg_frame_bytes >> 2;
        for (int it = 0; it < nst; ++it) {
            const int s = it % p.nstage;
            const int cs = it % kMMSlots;
            const int64_t f0 = fbeg + (int64_t)it * p.k;
            const int kk = (int)min((int64_t)p.k, fend - f0);
            const unsigned char* stage = reinterpret_cast<const unsigned char*>(stages + (size_t)s * p.stage_floats);
            const uint32_t h0b = (uint32_t)(reinterpret_cast<uintptr_t>(
                                     p.xyz + ((size_t)f0 * p.N + (size_t)(tile.atom0 - p.base)) * 3) & 15u);
            float* cg_tile = has_meas ? reinterpret_cast<float*>(cgbuf + (size_t)cs * m.cg_slot_bytes) : nullptr;
            mbar_wait(&full[s], (it / p.nstage) & 1);
            // the slot of the CG tile must have been drained by the measurement warps
            if (has_meas && it >= kMMSlots) mbar_wait(&cgempty[cs], ((it / kMMSlots) - 1) & 1);
            if (resident) {
                if (warp < nchunks)
                    map_task<FG, HAS_BOX, false>(p, tile, lc, stage, 0, kk, f0, frame_fl, 0, h0b, hstep, cg_tile, cg_stride);
            } else {
                for (int chunk = warp; chunk < nchunks; chunk += ncw) {
                    lc = lane_const(tile, ents, nullptr, chunk, lane);
                    map_task<FG, HAS_BOX, false>(p, tile, lc, stage, 0, kk, f0, frame_fl, 0, h0b, hstep, cg_tile, cg_stride);
                }
            }
            if (has_meas && warp == 0) {
                // the box of the stage's frames in the form the term code loads (see measure_fused_kernel)
                float* tab = cg_tile + FG * cg_stride;
                if (lane < kk) {
                    const int j = lane;
                    Box bx;
                    if (m.box_vec) {
                        RawBox<MIC> rb;
                        fetch_box<MIC>(m.box_vec, f0 + j, rb);
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
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                if (has_meas) mbar_arrive(&cgfull[cs]);   // release: the CG tile written above is visible to the waiters
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- measurement warps
    if (!has_meas) return;
    const int w = warp - ncw;                       // 0 .. kMMWarps - 1
    const int vt = w * 32 + lane;
    using Rec = RecLayout<SHAPE>;
    TermTables tabs;
    tabs.sbond = reinterpret_cast<int*>(smem_raw + m.off_sbond);
    tabs.sparam = reinterpret_cast<float4*>(smem_raw + m.off_sparam);
    tabs.shist = reinterpret_cast<unsigned int*>(smem_raw + m.off_shist);
    tabs.slist = reinterpret_cast<uint16_t*>(smem_raw + m.off_lists);
    tabs.snbad = reinterpret_cast<unsigned int*>(smem_raw + m.off_nbad);
    tabs.eacc = reinterpret_cast<float*>(smem_raw + m.off_eacc);
    tabs.ecnt = reinterpret_cast<int*>(tabs.eacc + 4 * 2 * kTermThreads);
    tabs.shist32 = smem_u32(tabs.shist);
    tabs.dummy32 = smem_u32(smem_raw + m.off_bars + 56);
    tabs.rs32 = (uint32_t)m.row_stride;
    tabs.shift = m.shift;
    tabs.hlo = m.hlo;
    tabs.hhi = m.hhi;
    tabs.values = m.values;
    tabs.row_stride = m.row_stride;
    tabs.partials = m.partials;
    tabs.hist = m.hist;
    tabs.nbins = m.nbins;
    tabs.hstride = hist_stride(m.nbins);
    tabs.fold = 1;
    const uint32_t ecache32 = smem_u32(smem_raw + m.off_ecache) + (uint32_t)w * (uint32_t)Rec::kWarpBytes;
    const int fl = kMMWarps / ng;                   // warps per group: they take every fl-th stage
    const bool working = w < ng * fl;
    const int gi = working ? w % ng : 0;
    const int fli = w / ng;
    term_tables_fill(tabs, mtile, m.tile_bonds, m.tile_lists, vt, kMMWarps * 32);
    TermWarp<MIC, SHAPE> tw;
    tw.setup(working, m.groups[mtile.group0 + gi], m.lanes + (size_t)(mtile.group0 + gi) * kChunks * 32, lane, 0u, ecache32,
             smem_u32(smem_raw + m.off_roles) + (uint32_t)w * (uint32_t)kRoleTableBytes, vt);
    asm volatile("bar.sync 2, %0;" ::"r"(kMMWarps * 32) : "memory");
    if (working) {
        for (int it = fli; it < nst; it += fl) {
            const int cs = it % kMMSlots;
            const int64_t f0 = fbeg + (int64_t)it * p.k;
            const int kk = (int)min((int64_t)p.k, fend - f0);
            const unsigned char* slot = cgbuf + (size_t)cs * m.cg_slot_bytes;
            float* vq = m.values ? m.values + (size_t)f0 * m.row_stride : nullptr;
            mbar_wait(&cgfull[cs], (it / kMMSlots) & 1);
            const float* tq = reinterpret_cast<const float*>(slot + (size_t)FG * m.cg_frame_bytes);
            // the launch's output options are uniform: one of four specialised copies of the quad runs
            if (m.values != nullptr) {
                if (m.hist != nullptr) tw.template quad_mode<3>(tabs, lane, slot, (uint32_t)m.cg_frame_bytes, tq, vq, kk);
                else tw.template quad_mode<1>(tabs, lane, slot, (uint32_t)m.cg_frame_bytes, tq, vq, kk);
            } else {
                if (m.hist != nullptr) tw.template quad_mode<2>(tabs, lane, slot, (uint32_t)m.cg_frame_bytes, tq, vq, kk);
                else tw.template quad_mode<0>(tabs, lane, slot, (uint32_t)m.cg_frame_bytes, tq, vq, kk);
            }
            if (lane == 0) mbar_arrive(&cgempty[cs]);
        }
    }
    tw.epilogue(tabs, mtile, working, w, lane, kMMWarps, ng, fl, 2);
}

}  // namespace cgb

extern "C" int cgb_map_measure(const float* xyz, const float* box_len, const float* box_vec, int ortho,
                               int64_t n_frames, int64_t n_atoms, int64_t n_beads, const CgbMapTile* tiles,
                               int64_t n_tiles, const CgbMapEntry* entries, int32_t max_span,
                               int32_t max_tile_entries, int32_t max_tile_beads, float* cg_xyz,
                               const int32_t* meas_of_tile, const CgbMeasTile* meas_tiles,
                               const CgbMeasGroup* groups, const CgbMeasLane* lanes, const int32_t* tile_bonds,
                               const uint16_t* tile_lists, int32_t shape, int32_t max_slots, const float* shift,
                               float* values, int64_t row_stride, double* partials, double* hist,
                               const float* hist_lo, const float* hist_hi, int32_t n_bins, void* stream) {
    using namespace cgb;
    if (n_frames < 0 || n_atoms < 0 || n_beads < 0 || n_tiles < 0 || max_span < 0 || max_tile_entries < 0 ||
        max_tile_beads < 1 || max_slots < 1)
        return CGB_EINVAL;
    if (n_frames == 0 || n_tiles == 0) return CGB_OK;
    if (!xyz || !tiles || !entries || !cg_xyz || !meas_of_tile || !meas_tiles || !groups || !lanes || !tile_bonds ||
        !tile_lists || !partials)
        return CGB_EINVAL;
    if (hist && (!hist_lo || !hist_hi || n_bins < 1 || n_bins > 512)) return CGB_EINVAL;
    if (values && (row_stride < 1 || row_stride > 0x3fffffffLL)) return CGB_EINVAL;
    if (shape != CGB_SHAPE_EAD && shape != CGB_SHAPE_EEAA) return CGB_EINVAL;
    if ((reinterpret_cast<uintptr_t>(xyz) & 3u) || (reinterpret_cast<uintptr_t>(cg_xyz) & 3u)) return CGB_EALIGN;
    if (n_tiles > 0x7fffffffLL) return CGB_EINVAL;

    MapParams p;
    std::memset(&p, 0, sizeof p);
    p.xyz = xyz;
    p.box = box_len;
    p.F = n_frames;
    p.N = n_atoms;
    p.base = 0;
    p.B = n_beads;
    p.tiles = tiles;
    p.ent = entries;
    p.w64 = nullptr;
    p.out = cg_xyz;
    p.nstage = 2;
    p.ent_cap = (max_tile_entries + 1) & ~1;
    p.contiguous = 0;
    p.slot_floats = (3 * max_span + 8 + 3) & ~3;
    p.k = kMMFG;                                    // one quad per stage
    p.tab_off = p.k * p.slot_floats;
    p.stage_floats = p.tab_off + ((6 * p.k + 3) & ~3);
    // frames per CTA: long runs amortise a CTA's start-up (plan entries, first copies), as long as the grid keeps
    // at least ~16 CTAs per SM to balance the tail
    p.fc = 128;
    {
        int dev = 0, sms = 0;
        CGB_CUDA_TRY(cudaGetDevice(&dev));
        CGB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        while (p.fc < 512 && n_tiles * ((n_frames + 2 * p.fc - 1) / (2 * p.fc)) >= 16 * (int64_t)sms) p.fc *= 2;
    }
    {
        const int64_t nchunks = (n_frames + p.fc - 1) / p.fc;
        if (nchunks > 65535) {
            int64_t need = (n_frames + 65534) / 65535;
            need = (need + p.k - 1) / p.k * p.k;
            p.fc = (int)need;
        }
    }
    const int64_t ny = (n_frames + p.fc - 1) / p.fc;

    MMParams m;
    std::memset(&m, 0, sizeof m);
    m.meas_of_tile = meas_of_tile;
    m.mtiles = meas_tiles;
    m.groups = groups;
    m.lanes = lanes;
    m.tile_bonds = tile_bonds;
    m.tile_lists = tile_lists;
    m.box_vec = box_vec;
    m.shift = shift;
    m.hlo = hist_lo;
    m.hhi = hist_hi;
    m.values = values;
    m.row_stride = row_stride;
    m.partials = partials;
    m.hist = hist;
    m.nbins = hist ? n_bins : 0;
    m.max_slots = max_slots;
    m.ncw = 5;                                      // 5 mapping + 2 measurement + 1 producer warp = 256 threads
    m.cg_frame_bytes = (3 * max_tile_beads * 4 + 15) & ~15;
    m.cg_slot_bytes = kMMFG * m.cg_frame_bytes + 256;
    auto up = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
    size_t off = 128 + (size_t)p.ent_cap * sizeof(CgbMapEntry) + (size_t)p.nstage * p.stage_floats * 4;
    off = up(off, 64);
    m.off_bars = (int)off;
    off += 64;
    m.off_sbond = (int)off;
    off = up(off + (size_t)max_slots * 4, 16);
    m.off_sparam = (int)off;
    off += (size_t)max_slots * 16;
    m.off_shist = (int)off;
    off += hist ? (size_t)max_slots * hist_stride(n_bins) * 4 : 0;
    m.off_lists = (int)off;
    off = up(off + 2 * ((size_t)max_slots + 1 + kRoles), 16);
    m.off_nbad = (int)off;
    off += (size_t)kRoles * 4;
    m.off_eacc = (int)off;
    off = up(off + (size_t)(4 * 2 * kTermThreads + 2 * kTermThreads) * 4, 128);
    m.off_roles = (int)off;
    off += (size_t)kMMWarps * kRoleTableBytes;
    m.off_ecache = (int)off;
    off += (size_t)kMMWarps * (shape == CGB_SHAPE_EAD ? RecLayout<CGB_SHAPE_EAD>::kWarpBytes
                                                      : RecLayout<CGB_SHAPE_EEAA>::kWarpBytes);
    off = up(off, 128);
    m.off_cg = (int)off;
    off += (size_t)kMMSlots * m.cg_slot_bytes;
    const size_t smem = off;
    if (smem > 227 * 1024) return CGB_ETOOBIG;

    dim3 grid((unsigned)n_tiles, (unsigned)ny);
    const int threads = (m.ncw + kMMWarps + 1) * 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool hb = box_len != nullptr;
    const bool triclinic = box_vec != nullptr && !ortho;
#define CGB_MM(HB_, M_, S_)                                                                               \
    do {                                                                                                  \
        auto kern = map_measure_kernel<HB_, M_, S_>;                                                      \
        CGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, threads, smem, st>>>(p, m);                                                          \
        return (int)cudaGetLastError();                                                                   \
    } while (0)
#define CGB_MM_BOX(M_, S_)             \
    do {                               \
        if (hb) CGB_MM(true, M_, S_);  \
        CGB_MM(false, M_, S_);         \
    } while (0)
    if (triclinic && shape == CGB_SHAPE_EAD) CGB_MM_BOX(MIC_TRICLINIC, CGB_SHAPE_EAD);
    if (triclinic) CGB_MM_BOX(MIC_TRICLINIC, CGB_SHAPE_EEAA);
    if (shape == CGB_SHAPE_EAD) CGB_MM_BOX(MIC_ORTHO, CGB_SHAPE_EAD);
    CGB_MM_BOX(MIC_ORTHO, CGB_SHAPE_EEAA);
#undef CGB_MM_BOX
#undef CGB_MM
}
