// Stage 1 -- atom -> bead mapping (reference: pycgtool/mapping.py:391-463).
//
// K1 `map_tiles_kernel`: a segmented weighted reduction over frame-major coordinates.
//   * The plan cuts the output beads into tiles whose atoms lie in one contiguous input span
//     (atoms of a residue are contiguous, beads never cross residues -- mapping.py:419-425).
//   * One CTA owns (tile, block of frames).  A producer warp streams the tile's span of every
//     frame into shared memory with 1-D bulk asynchronous copies (TMA, `cp.async.bulk`, SASS
//     UBLKCP) through a multi-stage mbarrier ring; every input byte is fetched from HBM once,
//     in full 16-byte-aligned runs, with an L2 evict-first policy.
//   * Consumer threads own one output float (bead, component) for FG frames at a time, walk the
//     bead's atom list (slot-major ELL entries held in shared memory, loaded once per CTA),
//     apply the reference's minimum-image update and accumulate in the reference's order, then
//     store coalesced.
// K1g `map_gather_kernel`: CSR gather straight from global memory, for virtual beads (which read
//     the CG frame, mapping.py:431-438) and for beads wider than a tile.
//
// Bit-exactness: every floating-point step of calc_coords_weight (mapping.py:455-462) is done
// with explicitly rounded, un-fused operations in the same order, so float32 results equal the
// NumPy reference bit for bit.

#include <algorithm>
#include <vector>

#include "cgb_mapwarp.cuh"

namespace cgb {

// ------------------------------------------------------------------------------------------
// Host: plan compiler
// ------------------------------------------------------------------------------------------
struct TileDraft {
    int64_t bead0 = 0;
    int64_t nbeads = 0;
    int64_t lo = INT64_MAX, hi = -1;
    int32_t nslots = 0;
};

static inline int64_t entry_cap(int32_t max_span_atoms) { return std::max<int64_t>(1024, 3LL * max_span_atoms); }

// Greedy left-to-right tiling; `emit(draft)` is called for every tile that computes >= 1 bead.
template <typename Emit>
static int tile_walk(const int32_t* bead_ptr, const int32_t* index, const uint8_t* compute, int64_t n_beads,
                     int64_t n_atoms, int32_t max_span_atoms, int32_t max_tile_beads, const uint8_t* tile_start,
                     Emit emit) {
    if (!bead_ptr || (!index && bead_ptr[n_beads] > 0) || n_beads < 0 || n_atoms < 0 || max_span_atoms < 1 ||
        max_tile_beads < 1)
        return CGB_EINVAL;
    const int64_t cap = entry_cap(max_span_atoms);
    TileDraft cur;
    cur.bead0 = 0;
    auto flush = [&](int64_t next_bead0) {
        if (cur.hi >= 0) emit(cur);
        cur = TileDraft();
        cur.bead0 = next_bead0;
    };
    for (int64_t b = 0; b < n_beads; ++b) {
        const int64_t p0 = bead_ptr[b], p1 = bead_ptr[b + 1];
        if (p1 < p0) return CGB_EINVAL;
        if (tile_start && tile_start[b] && cur.nbeads > 0) flush(b);   // a tile must begin at this bead
        const bool active = (compute == nullptr || compute[b]) && p1 > p0;
        if (!active) {
            // The bead keeps its output slot inside the current tile but is not computed here.
            if (cur.nbeads + 1 > max_tile_beads || (int64_t)cur.nslots * (cur.nbeads + 1) > cap) flush(b);
            cur.nbeads += 1;
            continue;
        }
        int64_t blo = INT64_MAX, bhi = -1;
        for (int64_t p = p0; p < p1; ++p) {
            const int64_t a = index[p];
            if (a < 0 || a >= n_atoms) return CGB_ERANGE;
            blo = std::min(blo, a);
            bhi = std::max(bhi, a);
        }
        const int64_t cnt = p1 - p0;
        if (bhi - blo + 1 > max_span_atoms || cnt > cap) return CGB_ETOOBIG;
        const int64_t nlo = std::min(cur.lo, blo), nhi = std::max(cur.hi, bhi);
        const int64_t nsl = std::max<int64_t>(cur.nslots, cnt);
        if (nhi - nlo + 1 > max_span_atoms || cur.nbeads + 1 > max_tile_beads || nsl * (cur.nbeads + 1) > cap) {
            flush(b);
            cur.lo = blo;
            cur.hi = bhi;
            cur.nslots = (int32_t)cnt;
        } else {
            cur.lo = nlo;
            cur.hi = nhi;
            cur.nslots = (int32_t)nsl;
        }
        cur.nbeads += 1;
    }
    flush(n_beads);
    return CGB_OK;
}

}  // namespace cgb

extern "C" int cgb_map_plan_size(const int32_t* bead_ptr, const int32_t* index, const uint8_t* compute,
                                 int64_t n_beads, int64_t n_atoms, int32_t max_span_atoms, int32_t max_tile_beads,
                                 int64_t* n_tiles, int64_t* n_entries, int32_t* max_span,
                                 int32_t* max_tile_entries) {
    return cgb_map_plan_size_cuts(bead_ptr, index, compute, n_beads, n_atoms, max_span_atoms, max_tile_beads, nullptr,
                                  n_tiles, n_entries, max_span, max_tile_entries);
}

extern "C" int cgb_map_plan_size_cuts(const int32_t* bead_ptr, const int32_t* index, const uint8_t* compute,
                                      int64_t n_beads, int64_t n_atoms, int32_t max_span_atoms,
                                      int32_t max_tile_beads, const uint8_t* tile_start, int64_t* n_tiles,
                                      int64_t* n_entries, int32_t* max_span, int32_t* max_tile_entries) {
    if (!n_tiles || !n_entries || !max_span || !max_tile_entries) return CGB_EINVAL;
    int64_t nt = 0, ne = 0;
    int32_t ms = 0, me = 0;
    int rc = cgb::tile_walk(bead_ptr, index, compute, n_beads, n_atoms, max_span_atoms, max_tile_beads, tile_start,
                            [&](const cgb::TileDraft& t) {
                                nt += 1;
                                const int64_t e = (int64_t)t.nslots * t.nbeads;
                                ne += e;
                                ms = std::max<int32_t>(ms, (int32_t)(t.hi - t.lo + 1));
                                me = std::max<int32_t>(me, (int32_t)e);
                            });
    if (rc != CGB_OK) return rc;
    *n_tiles = nt;
    *n_entries = ne;
    *max_span = ms;
    *max_tile_entries = me;
    return CGB_OK;
}

extern "C" int cgb_map_plan_build(const int32_t* bead_ptr, const int32_t* index, const void* weights,
                                  int weights_f64, const uint8_t* compute, int64_t n_beads, int64_t n_atoms,
                                  int32_t max_span_atoms, int32_t max_tile_beads, CgbMapTile* tiles,
                                  CgbMapEntry* entries, double* entries_w64) {
    return cgb_map_plan_build_cuts(bead_ptr, index, weights, weights_f64, compute, n_beads, n_atoms, max_span_atoms,
                                   max_tile_beads, nullptr, tiles, entries, entries_w64);
}

extern "C" int cgb_map_plan_build_cuts(const int32_t* bead_ptr, const int32_t* index, const void* weights,
                                       int weights_f64, const uint8_t* compute, int64_t n_beads, int64_t n_atoms,
                                       int32_t max_span_atoms, int32_t max_tile_beads, const uint8_t* tile_start,
                                       CgbMapTile* tiles, CgbMapEntry* entries, double* entries_w64) {
    if (!tiles || !entries || !weights || (weights_f64 && !entries_w64)) return CGB_EINVAL;
    const float* w32 = static_cast<const float*>(weights);
    const double* w64 = static_cast<const double*>(weights);
    int64_t nt = 0, ne = 0;
    return cgb::tile_walk(
        bead_ptr, index, compute, n_beads, n_atoms, max_span_atoms, max_tile_beads, tile_start,
        [&](const cgb::TileDraft& t) {
            CgbMapTile& out = tiles[nt++];
            out.atom0 = (int32_t)t.lo;
            out.span = (int32_t)(t.hi - t.lo + 1);
            out.bead0 = (int32_t)t.bead0;
            out.nbeads = (int32_t)t.nbeads;
            out.ent0 = ne;
            out.nslots = t.nslots;
            out.reserved = 0;
            CgbMapEntry* e = entries + ne;
            double* ew = weights_f64 ? entries_w64 + ne : nullptr;
            for (int64_t j = 0; j < t.nbeads; ++j) {
                const int64_t b = t.bead0 + j;
                const int64_t p0 = bead_ptr[b], p1 = bead_ptr[b + 1];
                const bool active = (compute == nullptr || compute[b]) && p1 > p0;
                const int64_t cnt = active ? p1 - p0 : 0;
                const int32_t ref3 = active ? (int32_t)(3 * (index[p0] - t.lo)) : 0;
                for (int64_t s = 0; s < t.nslots; ++s) {
                    CgbMapEntry& x = e[s * t.nbeads + j];
                    if (s == 0) {
                        x.off3 = ref3;
                        x.bits = (int32_t)cnt;
                        if (ew) ew[j] = 0.0;
                    } else if (s < cnt) {
                        x.off3 = (int32_t)(3 * (index[p0 + s] - t.lo));
                        if (weights_f64) {
                            x.bits = 0;
                            ew[s * t.nbeads + j] = w64[p0 + s];
                        } else {
                            float wv = w32[p0 + s];
                            x.bits = *reinterpret_cast<int32_t*>(&wv);
                        }
                    } else {
                        x.off3 = ref3;  // padding: v = 0, weight = 0
                        x.bits = 0;
                        if (ew) ew[s * t.nbeads + j] = 0.0;
                    }
                }
            }
            ne += (int64_t)t.nslots * t.nbeads;
        });
}

namespace cgb {

template <int FG, bool HAS_BOX, bool W64>
__global__ void __launch_bounds__(kProducerThreads + 256) map_tiles_kernel(const MapParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + kMaxStages;
    CgbMapEntry* ents = reinterpret_cast<CgbMapEntry*>(smem_raw + 128);
    double* entw = reinterpret_cast<double*>(smem_raw + 128 + (size_t)p.ent_cap * sizeof(CgbMapEntry));
    float* stages = reinterpret_cast<float*>(smem_raw + 128 + (size_t)p.ent_cap * sizeof(CgbMapEntry) * (W64 ? 2 : 1));

    const int ncons = blockDim.x - kProducerThreads;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // warp index through a shuffle so that the compiler knows it is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ncw = ncons >> 5;
    const CgbMapTile tile = p.tiles[blockIdx.x];
    const int64_t fbeg = (int64_t)blockIdx.y * p.fc;
    const int64_t fend = min(p.F, fbeg + p.fc);
    const int nst = (int)((fend - fbeg + p.k - 1) / p.k);
    const int nent = tile.nslots * tile.nbeads;

    if (tid == 0) {
        for (int s = 0; s < p.nstage; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], ncw);
        }
        fence_barrier_init();
    }
    // Plan entries of this tile: loaded once, reused for every frame of the CTA.
    for (int i = tid; i < nent; i += blockDim.x) {
        ents[i] = p.ent[tile.ent0 + i];
        if (W64) entw[i] = p.w64[tile.ent0 + i];
    }
    __syncthreads();

    // In contiguous mode the stage holds whole frames back to back, so the tile's atoms start
    // 3*atom0 floats into each frame.
    const int span_fl = 3 * tile.span;
    const int frame_fl = p.contiguous ? (int)(3 * p.N) : p.slot_floats;
    const int lead_fl = p.contiguous ? 3 * tile.atom0 : 0;

    if (warp == ncw) {
        // ------------------------------------------------------------------ producer warp
        const uintptr_t lo = reinterpret_cast<uintptr_t>(p.xyz);
        const uintptr_t hi = lo + (uintptr_t)p.F * (uintptr_t)p.N * 12u;
        const uint64_t policy = policy_evict_first();
        for (int it = 0; it < nst; ++it) {
            const int s = it % p.nstage;
            if (it >= p.nstage) mbar_wait(&empty[s], ((it / p.nstage) - 1) & 1);
            const int64_t f0 = fbeg + (int64_t)it * p.k;
            const int kk = (int)min((int64_t)p.k, fend - f0);
            float* sbase = stages + (size_t)s * p.stage_floats;
            const int ncopies = p.contiguous ? 1 : kk;
            const uint32_t nfl = p.contiguous ? (uint32_t)kk * (uint32_t)(3 * p.N) : (uint32_t)span_fl;
            // lane j owns copy j (a stage never has more than 32 frames unless it is contiguous)
            uint32_t bytes = 0;
            bool edge = false;
            const float* src = nullptr;
            if (lane < ncopies) {
                src = p.contiguous ? p.xyz + (size_t)f0 * p.N * 3
                                   : p.xyz + ((size_t)(f0 + lane) * p.N + (size_t)(tile.atom0 - p.base)) * 3;
                bytes = span_bytes(src, nfl, lo, hi, &edge);
            }
            // runs touching the first / last bytes of the whole array are copied by hand
            unsigned edges = __ballot_sync(0xffffffffu, edge);
            while (edges) {
                const int j = __ffs(edges) - 1;
                edges &= edges - 1;
                const float* esrc = reinterpret_cast<const float*>(
                    __shfl_sync(0xffffffffu, reinterpret_cast<uintptr_t>(src), j));
                float* dst = sbase + (size_t)j * p.slot_floats + head_floats(esrc);
                for (uint32_t i = lane; i < nfl; i += 32) dst[i] = __ldg(esrc + i);
            }
            if (HAS_BOX) {
                // per-frame box table of the stage: (|L|, t05) for each component
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
            if (lane < ncopies && !edge) {
                const void* a0 = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15));
                bulk_g2s_stream(sbase + (size_t)lane * p.slot_floats, a0, bytes, &full[s], policy);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    // Warp w owns the 32-lane chunks w, w + ncw, ... of the tile's output floats for every frame.
    const int nchunks = (3 * tile.nbeads + 31) >> 5;
    const bool resident = nchunks <= ncw;  // one chunk per warp: lane constants computed once
    LaneConst lc = lane_const(tile, ents, entw, warp < nchunks ? warp : 0, lane);
    const uint32_t hstep = (uint32_t)((12ull * (uint64_t)p.N) & 15ull);
    for (int it = 0; it < nst; ++it) {
        const int s = it % p.nstage;
        const int64_t f0 = fbeg + (int64_t)it * p.k;
        const int kk = (int)min((int64_t)p.k, fend - f0);
        const unsigned char* stage = reinterpret_cast<const unsigned char*>(stages + (size_t)s * p.stage_floats);
        const int ngroups = (kk + FG - 1) / FG;
        // byte misalignment of the first frame's tile (of the whole stage in contiguous mode)
        const uint32_t h0b = (uint32_t)(reinterpret_cast<uintptr_t>(
                                 p.contiguous ? p.xyz + (size_t)f0 * p.N * 3
                                              : p.xyz + ((size_t)f0 * p.N + (size_t)(tile.atom0 - p.base)) * 3) & 15u);
        mbar_wait(&full[s], (it / p.nstage) & 1);
        if (resident) {
            if (warp < nchunks)
                for (int g = 0; g < ngroups; ++g)
                    map_task<FG, HAS_BOX, W64>(p, tile, lc, stage, g, kk, f0, frame_fl, lead_fl, h0b, hstep);
        } else {
            for (int chunk = warp; chunk < nchunks; chunk += ncw) {
                lc = lane_const(tile, ents, entw, chunk, lane);
                for (int g = 0; g < ngroups; ++g)
                    map_task<FG, HAS_BOX, W64>(p, tile, lc, stage, g, kk, f0, frame_fl, lead_fl, h0b, hstep);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}

// ------------------------------------------------------------------------------------------
// Device: K1g (CSR gather from global memory)
// ------------------------------------------------------------------------------------------
template <bool W64>
__global__ void __launch_bounds__(256) map_gather_kernel(const float* __restrict__ src, int64_t src_atoms,
                                                         const float* __restrict__ box, int64_t F, int64_t B,
                                                         const int32_t* __restrict__ bead_ptr,
                                                         const int32_t* __restrict__ index,
                                                         const void* __restrict__ weights,
                                                         const int32_t* __restrict__ bead_list, int64_t n_list,
                                                         float* __restrict__ out) {
    const int64_t per_frame = 3 * n_list;
    const int64_t total = F * per_frame;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = t / per_frame;
        const int64_t r = t - f * per_frame;
        const int64_t item = r / 3;
        const int c = (int)(r - 3 * item);
        const int32_t b = bead_list[item];
        const int32_t p0 = bead_ptr[b], p1 = bead_ptr[b + 1];
        if (p1 <= p0) continue;
        const float* fr = src + (size_t)f * src_atoms * 3;
        const float ref = fr[(size_t)index[p0] * 3 + c];
        float L = 0.f, thr = 0.f;
        if (box) {
            L = box[f * 3 + c];
            thr = 0.49f * fabsf(L);
        }
        float acc = 0.f;
        double accd = 0.0;
        for (int32_t q = p0 + 1; q < p1; ++q) {
            float v = __fsub_rn(fr[(size_t)index[q] * 3 + c], ref);
            if (box) v = mic1(v, L, thr);
            if (W64)
                accd = __dadd_rn(accd, __dmul_rn(static_cast<const double*>(weights)[q], (double)v));
            else
                acc = __fadd_rn(acc, __fmul_rn(static_cast<const float*>(weights)[q], v));
        }
        const float res = W64 ? __double2float_rn(__dadd_rn(accd, (double)ref)) : __fadd_rn(acc, ref);
        out[((size_t)f * B + b) * 3 + c] = res;
    }
}

// ------------------------------------------------------------------------------------------
// Host: launchers
// ------------------------------------------------------------------------------------------
// Defaults from the sweep in profiles/ (S2, B200): 32-bead tiles, 2 stages of <= 24 KB and 3
// consumer warps give five CTAs (15 consumer warps, ~90 KB of loads in flight) per SM.
constexpr int kDefStageBytes = 24576, kDefStages = 2, kDefFramesPerCta = 128, kDefConsumers = 96, kDefFg = 4;

template <int FG>
static cudaError_t launch_tiles(const MapParams& p, dim3 grid, int threads, size_t smem, bool has_box, bool w64,
                                cudaStream_t st) {
#define CGB_LAUNCH(HB, WD)                                                                                     \
    do {                                                                                                       \
        auto kern = map_tiles_kernel<FG, HB, WD>;                                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        if (e != cudaSuccess) return e;                                                                        \
        kern<<<grid, threads, smem, st>>>(p);                                                                  \
        return cudaGetLastError();                                                                             \
    } while (0)
    if (has_box && w64) CGB_LAUNCH(true, true);
    if (has_box && !w64) CGB_LAUNCH(true, false);
    if (!has_box && w64) CGB_LAUNCH(false, true);
    CGB_LAUNCH(false, false);
#undef CGB_LAUNCH
}

}  // namespace cgb

extern "C" int cgb_map_tiles(const float* xyz, const float* box_len, int64_t n_frames, int64_t n_atoms,
                             int64_t n_beads, const CgbMapTile* tiles, int64_t n_tiles, const CgbMapEntry* entries,
                             const double* entries_w64, int32_t max_span, int32_t max_tile_entries, float* cg_xyz,
                             void* stream) {
    return cgb_map_tiles_ex(xyz, box_len, n_frames, n_atoms, 0, n_beads, tiles, n_tiles, entries, entries_w64,
                            max_span, max_tile_entries, cg_xyz, nullptr, stream);
}

extern "C" int cgb_copy_window_h2d(float* dst, const float* src, int64_t n_frames, int64_t n_atoms, int64_t first_atom,
                                   int64_t n_window_atoms, void* stream) {
    if (n_frames < 0 || n_atoms < 0 || first_atom < 0 || n_window_atoms < 0 || first_atom + n_window_atoms > n_atoms)
        return CGB_EINVAL;
    if (n_frames == 0 || n_window_atoms == 0) return CGB_OK;
    if (!dst || !src) return CGB_EINVAL;
    return (int)cudaMemcpy2DAsync(dst, (size_t)n_window_atoms * 12, src + (size_t)first_atom * 3, (size_t)n_atoms * 12,
                                  (size_t)n_window_atoms * 12, (size_t)n_frames, cudaMemcpyHostToDevice,
                                  static_cast<cudaStream_t>(stream));
}

extern "C" int cgb_map_tiles_window(const float* xyz, const float* box_len, int64_t n_frames, int64_t n_atoms,
                                    int64_t first_atom, int64_t n_beads, const CgbMapTile* tiles, int64_t n_tiles,
                                    const CgbMapEntry* entries, const double* entries_w64, int32_t max_span,
                                    int32_t max_tile_entries, float* cg_xyz, void* stream) {
    return cgb_map_tiles_ex(xyz, box_len, n_frames, n_atoms, first_atom, n_beads, tiles, n_tiles, entries,
                            entries_w64, max_span, max_tile_entries, cg_xyz, nullptr, stream);
}

extern "C" int cgb_map_tiles_ex(const float* xyz, const float* box_len, int64_t n_frames, int64_t n_atoms,
                                int64_t first_atom, int64_t n_beads, const CgbMapTile* tiles, int64_t n_tiles,
                                const CgbMapEntry* entries, const double* entries_w64, int32_t max_span,
                                int32_t max_tile_entries, float* cg_xyz, const CgbMapTuning* tuning, void* stream) {
    using namespace cgb;
    // per-call tuning (no process-wide state): zero fields keep the defaults
    const int g_stage_bytes = tuning && tuning->stage_bytes ? tuning->stage_bytes : kDefStageBytes;
    const int g_stage_user = tuning && tuning->stage_bytes ? 1 : 0;
    const int g_nstages = tuning && tuning->n_stages ? tuning->n_stages : kDefStages;
    const int g_frames_per_cta = tuning && tuning->frames_per_cta ? tuning->frames_per_cta : kDefFramesPerCta;
    const int g_consumers = tuning && tuning->consumer_threads ? tuning->consumer_threads : kDefConsumers;
    const int g_fg = tuning && tuning->frame_group ? tuning->frame_group : kDefFg;
    if (g_stage_bytes < 0 || g_nstages < 2 || g_nstages > kMaxStages || g_frames_per_cta < 0 || g_consumers < 32 ||
        g_consumers > 256 || g_consumers % 32 || (g_fg != 1 && g_fg != 2 && g_fg != 4))
        return CGB_EINVAL;
    if (n_frames < 0 || n_atoms < 0 || n_beads < 0 || n_tiles < 0 || max_span < 0 || max_tile_entries < 0 ||
        first_atom < 0)
        return CGB_EINVAL;
    if (n_frames == 0 || n_tiles == 0) return CGB_OK;
    if (!xyz || !tiles || !entries || !cg_xyz) return CGB_EINVAL;
    if ((reinterpret_cast<uintptr_t>(xyz) & 3u) || (reinterpret_cast<uintptr_t>(cg_xyz) & 3u)) return CGB_EALIGN;
    if (n_tiles > 0x7fffffffLL) return CGB_EINVAL;

    MapParams p;
    p.xyz = xyz;
    p.box = box_len;
    p.F = n_frames;
    p.N = n_atoms;
    p.base = first_atom;
    p.B = n_beads;
    p.tiles = tiles;
    p.ent = entries;
    p.w64 = entries_w64;
    p.out = cg_xyz;
    p.nstage = g_nstages;
    p.ent_cap = (max_tile_entries + 1) & ~1;  // keep the float64 weight array 16-byte aligned
    const int fg = g_fg;

    // A single tile in a frame no larger than a slot: stage whole frames, one bulk copy per stage.
    // Small stages keep many CTAs resident (the sweep under profiles/ favours ~12 KB there).
    const int64_t frame_bytes = n_atoms * 12;
    p.contiguous = (n_tiles == 1 && first_atom == 0 && frame_bytes * fg <= g_stage_bytes) ? 1 : 0;
    if (p.contiguous) {
        const int64_t budget = g_stage_user ? g_stage_bytes : g_stage_bytes / 2;
        int k = (int)std::max<int64_t>(fg, budget / std::max<int64_t>(frame_bytes, 1));
        k -= k % fg;
        k = (int)std::min<int64_t>(k, std::max<int64_t>(n_frames, 1));
        p.k = k;
        p.slot_floats = 0;
        p.tab_off = (int)(((int64_t)k * n_atoms * 3 + 8 + 3) & ~3LL);
        p.stage_floats = p.tab_off + ((6 * k + 3) & ~3);
    } else {
        p.slot_floats = (3 * max_span + 8 + 3) & ~3;
        int k = std::min(32, std::max(1, g_stage_bytes / (p.slot_floats * 4)));
        if (k >= fg) k -= k % fg;
        k = (int)std::min<int64_t>(k, n_frames);
        p.k = k;
        p.tab_off = k * p.slot_floats;
        p.stage_floats = p.tab_off + ((6 * k + 3) & ~3);
    }
    int fc = std::max(g_frames_per_cta, p.k);
    fc -= fc % p.k;
    p.fc = fc;
    const int64_t nchunks = (n_frames + fc - 1) / fc;
    if (nchunks > 65535) {
        // grid.y limit: widen the per-CTA frame block
        int64_t need = (n_frames + 65534) / 65535;
        need = (need + p.k - 1) / p.k * p.k;
        p.fc = (int)need;
    }
    const int64_t ny = (n_frames + p.fc - 1) / p.fc;

    const size_t smem = 128 + (size_t)p.ent_cap * sizeof(CgbMapEntry) * (entries_w64 ? 2 : 1) +
                        (size_t)p.nstage * p.stage_floats * 4;
    if (smem > 227 * 1024) return CGB_ETOOBIG;
    dim3 grid((unsigned)n_tiles, (unsigned)ny);
    const int threads = g_consumers + kProducerThreads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    const bool hb = box_len != nullptr, wd = entries_w64 != nullptr;
    if (fg == 1)
        e = launch_tiles<1>(p, grid, threads, smem, hb, wd, st);
    else if (fg == 2)
        e = launch_tiles<2>(p, grid, threads, smem, hb, wd, st);
    else
        e = launch_tiles<4>(p, grid, threads, smem, hb, wd, st);
    return (int)e;
}

extern "C" int cgb_map_gather(const float* src, int64_t src_atoms, const float* box_len, int64_t n_frames,
                              int64_t n_beads, const int32_t* bead_ptr, const int32_t* index, const void* weights,
                              int weights_f64, const int32_t* bead_list, int64_t n_list, float* cg_xyz,
                              void* stream) {
    using namespace cgb;
    if (n_frames < 0 || n_beads < 0 || n_list < 0 || src_atoms < 0) return CGB_EINVAL;
    if (n_frames == 0 || n_list == 0) return CGB_OK;
    if (!src || !bead_ptr || !index || !weights || !bead_list || !cg_xyz) return CGB_EINVAL;
    const int64_t total = n_frames * n_list * 3;
    const int threads = 256;
    const int64_t blocks = std::min<int64_t>((total + threads - 1) / threads, (int64_t)kNumSMs * 32);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (weights_f64)
        map_gather_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(src, src_atoms, box_len, n_frames, n_beads,
                                                                     bead_ptr, index, weights, bead_list, n_list,
                                                                     cg_xyz);
    else
        map_gather_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(src, src_atoms, box_len, n_frames, n_beads,
                                                                      bead_ptr, index, weights, bead_list, n_list,
                                                                      cg_xyz);
    return (int)cudaGetLastError();
}
