// Device code of the mapping warps (K1), shared by the mapping kernel (cgb_map.cu) and the fused mapping +
// measurement kernel (cgb_map_measure.cu).  Reference: pycgtool/mapping.py:446-463 (calc_coords_weight).
#pragma once

#include "cgb_common.cuh"

// ------------------------------------------------------------------------------------------
// Device: K1
// ------------------------------------------------------------------------------------------
namespace cgb {

constexpr int kMaxStages = 8;
constexpr int kProducerThreads = 32;

struct MapParams {
    const float* xyz;
    const float* box;  // [F][3] or nullptr
    int64_t F, N, B;
    int64_t base;      // first atom held by xyz: the buffer is the window [base, base + N) of every frame
    const CgbMapTile* tiles;
    const CgbMapEntry* ent;
    const double* w64;
    float* out;
    int32_t k;            // frames per stage
    int32_t nstage;       // ring depth
    int32_t fc;           // frames per CTA
    int32_t slot_floats;  // floats per frame slot in a stage (multiple of 4)
    int32_t stage_floats; // floats per stage (multiple of 4), box table included
    int32_t tab_off;      // float offset of the per-frame box table inside a stage
    int32_t ent_cap;      // entries held in shared memory
    int32_t contiguous;   // tile == whole frame: one bulk copy per stage
};

// Minimum image of one component, v - L * rint(v / L) (mapping.py:457), without a division on the
// common paths.  With t05 = the largest float32 a for which rint(fl(a / L)) == 0 (computed exactly
// by the producer warp, see box_t05):
//     |v| <= t05               -> rint(v / L) = 0, v unchanged
//     t05 < |v| <= 1.49 L      -> rint(v / L) = +-1 and L * (+-1) is exact: v -= copysign(L, v)
//     otherwise (never for whole molecules) -> the IEEE division below.
// Each case reproduces the reference's float32 result bit for bit.
static __device__ __noinline__ float mic_divide(float v, float La) {
    const float k = rintf(__fdiv_rn(v, La));
    return __fsub_rn(v, __fmul_rn(La, k));
}

// t = L/2 divides to exactly 0.5, which rounds to 0 (ties to even); the next float above t may
// still divide to 0.5 after rounding, the one after that never does.
__device__ __forceinline__ float box_t05(float La) {
    const float t = 0.5f * La;
    const float t1 = __int_as_float(__float_as_int(t) + 1);
    return (__fdiv_rn(t1, La) <= 0.5f) ? t1 : t;
}

// What a lane needs to know about its output float (bead, component); constant for the lifetime of
// the CTA when the tile has no more 32-lane chunks than the CTA has consumer warps.
struct LaneConst {
    const CgbMapEntry* ep;  // slot-0 entry of the lane's bead
    const double* wp;       // float64 weight of slot 0 (W64 only)
    int o;                  // output float inside the tile: 3 * bead + component
    int c4;                 // 4 * component
    int lo;                 // byte offset of the reference atom's component inside a frame of the tile
    int cnt;                // atoms of the bead (0: not computed here)
    int cmax;               // max cnt over the warp (uniform loop bound; ELL padding is exact)
};

__device__ __forceinline__ LaneConst lane_const(const CgbMapTile& tile, const CgbMapEntry* ents, const double* entw,
                                                int chunk, int lane) {
    LaneConst k;
    const int nb3 = 3 * tile.nbeads;
    k.o = chunk * 32 + lane;
    const bool valid = k.o < nb3;
    const int oo = valid ? k.o : 0;
    const int jb = oo / 3;
    const int c = oo - 3 * jb;
    k.ep = ents + jb;
    k.wp = entw + jb;
    const CgbMapEntry e0 = *k.ep;
    k.cnt = valid ? e0.bits : 0;
    k.cmax = __reduce_max_sync(0xffffffffu, k.cnt);
    k.c4 = c * 4;
    k.lo = e0.off3 * 4 + k.c4;
    return k;
}

// One warp-task: the lane's output float for the FG frames of group g of the current stage.
// g, the stage and everything derived from them are warp-uniform (frame bases live in uniform
// registers, so the shared-memory loads are [lane offset + uniform frame base]).
// cg_tile != NULL: the results also go to a CG tile in shared memory (frame q of the group at cg_tile +
// q * cg_stride floats), for the measurement warps of the fused kernel.
template <int FG, bool HAS_BOX, bool W64>
__device__ __forceinline__ void map_task(const MapParams& p, const CgbMapTile& tile, const LaneConst& k,
                                         const unsigned char* stage, int g, int kk, int64_t f0, int frame_fl,
                                         int lead_fl, uint32_t h0b, uint32_t hstep, float* cg_tile = nullptr,
                                         int cg_stride = 0) {
    if (k.cmax == 0) return;
    uint32_t base[FG];  // byte offset of the frame's tile inside the stage (warp-uniform)
    float ref[FG], La[FG], t05[FG];
    const float2* tab = reinterpret_cast<const float2*>(stage + (size_t)p.tab_off * 4);
    const int c = k.c4 >> 2;
#pragma unroll
    for (int q = 0; q < FG; ++q) {
        const int j = min(g * FG + q, kk - 1);
        // 16-byte-alignment head of frame j: frames are 12 N bytes apart
        const uint32_t h = p.contiguous ? (h0b >> 2) : (((h0b + (uint32_t)j * hstep) & 15u) >> 2);
        base[q] = (uint32_t)(j * frame_fl + (int)h + lead_fl) * 4u;
        ref[q] = *reinterpret_cast<const float*>(stage + base[q] + k.lo);
        if (HAS_BOX) {
            const float2 bt = tab[j * 3 + c];  // (|L|, t05) of this frame and component
            La[q] = bt.x;
            t05[q] = bt.y;
        }
    }
    float rare = 0.f;
    if (HAS_BOX) {
        rare = La[0];
#pragma unroll
        for (int q = 1; q < FG; ++q) rare = fminf(rare, La[q]);
        rare *= 1.49f;
    }

    float acc[FG];
    double accd[FG];
    float vmax = 0.f;
#pragma unroll
    for (int q = 0; q < FG; ++q) {
        acc[q] = 0.f;
        accd[q] = 0.0;
    }
    const CgbMapEntry* ep = k.ep;
    const double* wp = k.wp;
    for (int i = 1; i < k.cmax; ++i) {
        ep += tile.nbeads;
        const CgbMapEntry e = *ep;
        const int a = e.off3 * 4 + k.c4;
        float v[FG];
#pragma unroll
        for (int q = 0; q < FG; ++q) v[q] = *reinterpret_cast<const float*>(stage + base[q] + a);
        if (FG % 2 == 0) {
#pragma unroll
            for (int q = 0; q < FG; q += 2) {
                const uint64_t d = sub2(pack2(v[q], v[q + 1]), pack2(ref[q], ref[q + 1]));
                unpack2(d, v[q], v[q + 1]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < FG; ++q) v[q] = __fsub_rn(v[q], ref[q]);
        }
        if (HAS_BOX) {
            // common cases without a branch: v += t * copysign(|L|, -v) with t = (|v| > t05) ? 1 : 0.
            // t * |L| is exact, so the fused form rounds once, exactly like v - L * rint(v / L); the
            // running maximum decides after the loop whether any |v| was beyond 1.49 |L|.
            float t[FG], sh[FG];
#pragma unroll
            for (int q = 0; q < FG; ++q) {
                vmax = fmaxf(vmax, fabsf(v[q]));
                t[q] = fabsf(v[q]) > t05[q] ? 1.f : 0.f;
                // |L| with the sign opposite to v, as one three-input logic op
                sh[q] = __int_as_float((__float_as_int(La[q]) & 0x7fffffff) | (~__float_as_int(v[q]) & 0x80000000));
            }
            if (FG % 2 == 0) {
#pragma unroll
                for (int q = 0; q < FG; q += 2) {
                    const uint64_t r = fma2(pack2(t[q], t[q + 1]), pack2(sh[q], sh[q + 1]), pack2(v[q], v[q + 1]));
                    unpack2(r, v[q], v[q + 1]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < FG; ++q) v[q] = __fmaf_rn(t[q], sh[q], v[q]);
            }
        }
        if (W64) {
            wp += tile.nbeads;
            const double wd = *wp;
#pragma unroll
            for (int q = 0; q < FG; ++q) accd[q] = __dadd_rn(accd[q], __dmul_rn(wd, (double)v[q]));
        } else {
            const float wt = __int_as_float(e.bits);
            if (FG % 2 == 0) {
                const uint64_t w2 = pack2(wt, wt);
#pragma unroll
                for (int q = 0; q < FG; q += 2) {
                    const uint64_t s = add2(pack2(acc[q], acc[q + 1]), mul2(w2, pack2(v[q], v[q + 1])));
                    unpack2(s, acc[q], acc[q + 1]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < FG; ++q) acc[q] = __fadd_rn(acc[q], __fmul_rn(wt, v[q]));
            }
        }
    }
    if (HAS_BOX && __any_sync(0xffffffffu, vmax > rare)) {
        // some |v| reached 1.49 |L| (atoms of one bead more than a box apart): redo the task with the
        // general IEEE-division minimum image.  Never taken for whole or wrapped molecules.
#pragma unroll
        for (int q = 0; q < FG; ++q) {
            acc[q] = 0.f;
            accd[q] = 0.0;
        }
        ep = k.ep;
        wp = k.wp;
        for (int i = 1; i < k.cmax; ++i) {
            ep += tile.nbeads;
            wp += tile.nbeads;
            const CgbMapEntry e = *ep;
            const int a = e.off3 * 4 + k.c4;
#pragma unroll
            for (int q = 0; q < FG; ++q) {
                float v = __fsub_rn(*reinterpret_cast<const float*>(stage + base[q] + a), ref[q]);
                v = mic_divide(v, La[q]);
                if (W64)
                    accd[q] = __dadd_rn(accd[q], __dmul_rn(*wp, (double)v));
                else
                    acc[q] = __fadd_rn(acc[q], __fmul_rn(__int_as_float(e.bits), v));
            }
        }
    }
    if (k.cnt > 0) {
        const size_t fstride = (size_t)p.B * 3;
        float* op = p.out + ((size_t)(f0 + (int64_t)g * FG) * p.B + tile.bead0) * 3 + k.o;
#pragma unroll
        for (int q = 0; q < FG; ++q) {
            if (g * FG + q < kk) {
                const float r = W64 ? __double2float_rn(__dadd_rn(accd[q], (double)ref[q]))
                                    : __fadd_rn(acc[q], ref[q]);
                st_stream(op + q * fstride, r);
                if (cg_tile) cg_tile[q * cg_stride + k.o] = r;
            }
        }
    }
}


}  // namespace cgb
