// Trajectory ingest: GROMACS .xtc (XDR "xdr3dfcoord" compressed coordinates).
//
// The reference loads trajectories through mdtraj's C codec (pycgtool/frame.py:41-63); at scale that
// decode, not the mapping, dominates the wall clock.  This file provides
//   * cgb_xtc_scan    (host)  walk the frame headers of a file image, fill a frame table;
//   * cgb_xtc_decode  (GPU)   decodes the bit streams straight into the [F][N][3] float32 trajectory
//                             in HBM (a per-frame walk that drops checkpoints, then one thread per
//                             checkpoint) -- the compressed bytes are what crosses PCIe (about a
//                             third of the raw floats);
//   * cgb_xtc_decode_host / cgb_xtc_encode (host) the same codec on the CPU, for files read or written
//     without a GPU in the loop and as the check of the GPU decoder.
// The bit-stream format is the public xdrfile one (magic-number table, mixed-radix packed triples,
// run-length coded small differences with the first two atoms of a run interchanged); coordinates are
// float32(int) * float32(1 / precision), the xdrfile convention.

#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include "cgb_common.cuh"

namespace cgb {

constexpr int kFirstIdx = 9;
constexpr int kLastIdx = 72;  // entries in the table below

#define CGB_MAGICINTS                                                                                         \
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406,  \
        512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384,      \
        20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280,     \
        416127, 524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304,       \
        5284491, 6658042, 8388607, 10568983, 13316085, 16777216

__device__ __constant__ int d_magicints[] = {CGB_MAGICINTS};
static const int h_magicints[] = {CGB_MAGICINTS};

#ifdef __CUDA_ARCH__
#define CGB_MAGIC(i) d_magicints[i]
#else
#define CGB_MAGIC(i) h_magicints[i]
#endif

__host__ __device__ inline int bit_length_u32(uint32_t v) {
    int n = 0;
    while (v) {
        ++n;
        v >>= 1;
    }
    return n;
}

__host__ __device__ inline uint32_t load_be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

// MSB-first bit reader over a byte stream whose first byte is 4-byte aligned (XDR payloads are);
// refills 32 bits at a time and can start at any bit position.
struct BitReader {
    const uint8_t* base;
    uint64_t word;    // index of the next 32-bit word to load
    uint64_t buf;     // low `have` bits are valid
    uint64_t nwords;  // words of the payload: reads beyond it return zeros (a corrupt stream cannot leave the frame)
    int have;
    __host__ __device__ inline uint32_t load_word(uint64_t w) const {
        if (w >= nwords) return 0u;
#ifdef __CUDA_ARCH__
        return __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(base) + w), 0, 0x0123);
#else
        return load_be32(base + 4 * w);
#endif
    }
    __host__ __device__ BitReader(const uint8_t* data, uint64_t payload_bytes, uint64_t bitpos = 0)
        : base(data), word(bitpos >> 5), buf(0), nwords((payload_bytes + 3) >> 2), have(0) {
        const int skip = (int)(bitpos & 31);
        if (skip) {
            buf = load_word(word++) & (0xffffffffu >> skip);
            have = 32 - skip;
        }
    }
    __host__ __device__ inline uint64_t position() const { return word * 32 - (uint64_t)have; }
    __host__ __device__ inline uint32_t take(int n) {  // n <= 32
        if (n == 0) return 0;
        if (have < n) {
            buf = (buf << 32) | (uint64_t)load_word(word++);
            have += 32;
        }
        have -= n;
        return (uint32_t)((buf >> have) & ((n == 32) ? 0xffffffffull : ((1ull << n) - 1ull)));
    }
    __host__ __device__ inline void skip_bits(uint64_t n) {  // jump ahead without reading the bits
        const uint64_t pos = position() + n;
        word = pos >> 5;
        buf = 0;
        have = 0;
        const int skip = (int)(pos & 31);
        if (skip) {
            buf = load_word(word++) & (0xffffffffu >> skip);
            have = 32 - skip;
        }
    }
};

// xdrfile receiveints: `nbits` bits hold a mixed-radix number, sent as bytes least significant first.
__host__ __device__ inline void take_triple(BitReader& r, int nbits, const uint32_t sizes[3], int out[3]) {
    if (nbits <= 64) {
        uint64_t v = 0;
        int shift = 0;
        while (nbits > 8) {
            v |= (uint64_t)r.take(8) << shift;
            shift += 8;
            nbits -= 8;
        }
        if (nbits > 0) v |= (uint64_t)r.take(nbits) << shift;
        const uint64_t q2 = v / sizes[2];
        out[2] = (int)(v - q2 * sizes[2]);
        const uint64_t q1 = q2 / sizes[1];
        out[1] = (int)(q2 - q1 * sizes[1]);
        out[0] = (int)q1;
    } else {
        // up to 72 bits: byte-array long division, as xdrfile does
        uint8_t bytes[12];
        int nb = 0;
        while (nbits > 8) {
            bytes[nb++] = (uint8_t)r.take(8);
            nbits -= 8;
        }
        if (nbits > 0) bytes[nb++] = (uint8_t)r.take(nbits);
        for (int i = 2; i > 0; --i) {
            uint64_t num = 0;
            for (int j = nb - 1; j >= 0; --j) {
                num = (num << 8) | bytes[j];
                const uint64_t p = num / sizes[i];
                bytes[j] = (uint8_t)p;
                num -= p * sizes[i];
            }
            out[i] = (int)num;
        }
        out[0] = (int)((uint32_t)bytes[0] | ((uint32_t)bytes[1] << 8) | ((uint32_t)bytes[2] << 16) |
                       ((uint32_t)bytes[3] << 24));
    }
}

// Constant part of a frame's coding: field widths of a full coordinate.
struct FrameCoding {
    uint32_t sizeint[3];
    int bitsizeint[3];
    int bitsize;     // 0 when the three components are stored as separate bit fields ("wide")
    int fullbits;    // bits of one full coordinate
    bool wide;
};

__host__ __device__ inline FrameCoding frame_coding(const CgbXtcFrame& fr) {
    FrameCoding c;
    c.wide = false;
    for (int d = 0; d < 3; ++d) {
        c.sizeint[d] = (uint32_t)(fr.maxint[d] - fr.minint[d] + 1);
        c.bitsizeint[d] = 0;
        if (c.sizeint[d] > 0xffffffu) c.wide = true;
    }
    if (c.wide) {
        for (int d = 0; d < 3; ++d) c.bitsizeint[d] = bit_length_u32(c.sizeint[d]);
        c.bitsize = 0;
        c.fullbits = c.bitsizeint[0] + c.bitsizeint[1] + c.bitsizeint[2];
    } else {
        // bit length of the product of the three sizes (up to 72 bits)
        const uint64_t lo = (uint64_t)c.sizeint[0] * c.sizeint[1];  // < 2^48
        const uint64_t hi_part = (lo >> 32) * c.sizeint[2];
        const uint64_t lo_part = (lo & 0xffffffffull) * c.sizeint[2];
        const uint64_t carry = hi_part + (lo_part >> 32);
        const uint64_t top = carry >> 32;  // bits above 64
        const uint64_t low64 = (carry << 32) | (lo_part & 0xffffffffull);
        if (top) {
            c.bitsize = 64 + bit_length_u32((uint32_t)top);
        } else {
            c.bitsize = bit_length_u32((uint32_t)(low64 >> 32));
            c.bitsize = c.bitsize ? c.bitsize + 32 : bit_length_u32((uint32_t)low64);
        }
        c.fullbits = c.bitsize;
    }
    return c;
}

// Running state of the decoder between groups (a group = one full coordinate + its run of smalls).
struct GroupState {
    int atom;      // next atom to be written
    int smallidx;  // current entry of the small-difference table
    int run;       // current run length (persists while the flag bit is 0)
};

// Decode up to `max_groups` groups starting at the reader's position; returns the groups decoded, or -1 when the
// stream is not a valid one: the small-difference table entry leaves [kFirstIdx, kLastIdx] or the reader passes the
// end of the payload (truncated or corrupt file).  Nothing is read or written outside the frame in either case.
__host__ __device__ inline int decode_groups(BitReader& r, const CgbXtcFrame& fr, const FrameCoding& c,
                                             GroupState& st, int max_groups, float* xyz) {
    const int natoms = fr.natoms;
    const float inv = 1.0f / fr.precision;
    int smallidx = st.smallidx;
    int smaller = CGB_MAGIC(smallidx - 1 > kFirstIdx ? smallidx - 1 : kFirstIdx) / 2;
    int smallnum = CGB_MAGIC(smallidx) / 2;
    uint32_t sizesmall[3];
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (uint32_t)CGB_MAGIC(smallidx);
    int i = st.atom, run = st.run, groups = 0;
    while (i < natoms && groups < max_groups) {
        int cur[3];
        if (c.wide) {
            for (int d = 0; d < 3; ++d) cur[d] = (int)r.take(c.bitsizeint[d]);
        } else {
            take_triple(r, c.bitsize, c.sizeint, cur);
        }
        for (int d = 0; d < 3; ++d) cur[d] += fr.minint[d];
        int is_smaller = 0;
        if (r.take(1)) {
            run = (int)r.take(5);
            is_smaller = run % 3;
            run -= is_smaller;
            is_smaller -= 1;
        }
        if (run > 0) {
            int prev[3] = {cur[0], cur[1], cur[2]};
            for (int k = 0; k < run && i < natoms; k += 3) {
                int t[3];
                take_triple(r, smallidx, sizesmall, t);
                for (int d = 0; d < 3; ++d) t[d] += prev[d] - smallnum;
                if (k == 0) {
                    // first two atoms of a run are interchanged (better compression of water)
                    for (int d = 0; d < 3; ++d) xyz[(size_t)i * 3 + d] = (float)t[d] * inv;
                    ++i;
                    if (i < natoms) {
                        for (int d = 0; d < 3; ++d) xyz[(size_t)i * 3 + d] = (float)prev[d] * inv;
                        ++i;
                    }
                } else {
                    for (int d = 0; d < 3; ++d) xyz[(size_t)i * 3 + d] = (float)t[d] * inv;
                    ++i;
                }
                for (int d = 0; d < 3; ++d) prev[d] = t[d];
            }
        } else {
            for (int d = 0; d < 3; ++d) xyz[(size_t)i * 3 + d] = (float)cur[d] * inv;
            ++i;
        }
        smallidx += is_smaller;
        if (smallidx < kFirstIdx || smallidx > kLastIdx || r.position() > r.nwords * 32) return -1;
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > kFirstIdx ? CGB_MAGIC(smallidx - 1) / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = CGB_MAGIC(smallidx) / 2;
        }
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (uint32_t)CGB_MAGIC(smallidx);
        ++groups;
    }
    st.atom = i;
    st.smallidx = smallidx;
    st.run = run;
    return groups;
}

// Decode one frame's compressed coordinates into xyz[natoms][3] (sequential; host path).  False: invalid stream.
__host__ __device__ inline bool decode_frame(const uint8_t* payload, const CgbXtcFrame& fr, float* xyz) {
    const FrameCoding c = frame_coding(fr);
    BitReader r(payload, (uint64_t)fr.nbytes);
    GroupState st = {0, fr.smallidx, 0};
    return decode_groups(r, fr, c, st, 0x7fffffff, xyz) >= 0 && st.atom >= fr.natoms;
}

__host__ __device__ inline void decode_raw(const uint8_t* payload, int natoms, float* xyz) {
    for (int i = 0; i < 3 * natoms; ++i) {
        const uint32_t bits = load_be32(payload + 4 * i);
        float v;
        memcpy(&v, &bits, 4);
        xyz[i] = v;
    }
}

// GPU decode in two kernels.  The bit stream of a frame is sequential, but finding where each group
// starts only needs its flag / run bits: `xtc_walk_kernel` (one warp per frame) skips over the
// coordinate fields and drops a checkpoint (bit position, atom index, table entry, run length) every
// kCheckEvery groups; `xtc_decode_kernel` then decodes every checkpoint's groups with one thread each,
// which is where the divisions, conversions and stores are.
constexpr int kCheckEvery = 16;

struct XtcCheckpoint {  // 12 bytes
    uint32_t bitpos;
    uint32_t atom;
    uint32_t state;     // smallidx | run << 8
};

__host__ __device__ inline int64_t checks_per_frame(int64_t n_atoms) { return n_atoms / kCheckEvery + 2; }

// `n` (<= 25) bits at bit position `pos` of a 4-byte aligned big-endian stream, without reader state.
__device__ __forceinline__ uint32_t peek_bits(const uint32_t* __restrict__ words, uint32_t pos, int n) {
    const uint32_t w = pos >> 5, sh = pos & 31;
    const uint32_t hi = __byte_perm(__ldg(words + w), 0, 0x0123);
    const uint32_t lo = __byte_perm(__ldg(words + w + 1), 0, 0x0123);
    return __funnelshift_l(lo, hi, sh) >> (32 - n);
}

// One warp per frame.  The 32 lanes stage the bit stream into a two-half ring in shared memory (coalesced loads,
// bytes swapped to big-endian once); lane 0 walks it.  A group's position depends on the previous group's (flag,
// run) bits, so the walk is a dependent chain -- pointer chasing through the bit stream, which no prefix scan can
// break because the positions themselves are the unknowns (real trajectories set the flag in 40-100 % of the groups,
// so speculating on "flag = 0" across lanes does not pay either; measured).  What can be done is to make the chain
// short: one step is branch-free integer arithmetic (run / 3 as a multiply-shift, selects instead of branches) and
// ONE 64-bit shared-memory load -- ring entry i holds words i and i + 1 of the stream, so the six bits that straddle
// a word boundary come from a single LDS.64.  ~100 cycles per group instead of ~700 for the branchy version with two
// 32-bit loads and a speculative second peek.  Throughput comes from walking all frames of a block at once.
constexpr int kWalkHalfWords = 1024;                        // stream words per half of the ring (8 KB of entries)
constexpr uint32_t kWalkMarginBits = 2 * 96 + 8 * 72 + 64;  // two full coordinates, one run of smalls, slack

__device__ __forceinline__ uint32_t ring_peek6(const uint2* ring, uint32_t pos) {
    const uint2 e = ring[(pos >> 5) & (2 * kWalkHalfWords - 1)];     // .x = word w, .y = word w + 1
    return __funnelshift_l(e.y, e.x, pos & 31) >> 26;
}

__global__ void __launch_bounds__(32) xtc_walk_kernel(const uint8_t* __restrict__ data, int64_t n_bytes,
                                                      const CgbXtcFrame* __restrict__ frames, int64_t n_frames,
                                                      int64_t n_atoms, XtcCheckpoint* __restrict__ checks,
                                                      int32_t* __restrict__ n_checks, int32_t* __restrict__ bad_frames) {
    __shared__ uint2 ring[2 * kWalkHalfWords];
    const int64_t f = blockIdx.x;
    const int lane = threadIdx.x;
    const CgbXtcFrame fr = frames[f];
    XtcCheckpoint* out = checks + f * checks_per_frame(n_atoms);
    if (fr.natoms <= 9) {
        if (lane == 0) n_checks[f] = 0;
        return;
    }
    const FrameCoding c = frame_coding(fr);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(data + fr.offset);
    // words of the image readable from the payload start (the image is padded by 8 bytes), and of the payload
    const int64_t avail = min((n_bytes + 8 - fr.offset) >> 2, ((int64_t)fr.nbytes + 3) >> 2);
    const uint32_t end_bits = (uint32_t)((((int64_t)fr.nbytes + 3) >> 2) * 32);
    bool bad = false;   // lane 0: the stream is not a valid one (see decode_groups)

    auto load_half = [&](int h) {
        uint2* dst = ring + (h & 1) * kWalkHalfWords;
        const int64_t w0 = (int64_t)h * kWalkHalfWords;
#pragma unroll 8
        for (int i = lane; i < kWalkHalfWords; i += 32) {
            const int64_t w = w0 + i;
            // entry i = (word w, word w + 1); the neighbour's word comes from the same cache lines
            const uint32_t a = w < avail ? __byte_perm(__ldg(words + w), 0, 0x0123) : 0u;
            const uint32_t b = w + 1 < avail ? __byte_perm(__ldg(words + w + 1), 0, 0x0123) : 0u;
            dst[i] = make_uint2(a, b);
        }
    };
    load_half(0);
    load_half(1);
    __syncwarp();
    int hl = 1;  // last half loaded; the ring holds halves hl - 1 and hl

    uint32_t pos = 0, smalls = 0;         // smalls = run / 3: small-difference triples per group
    int smallidx = fr.smallidx, atom = 0, group = 0, nc = 0;
    const uint32_t fullbits = (uint32_t)c.fullbits;
    const int natoms = fr.natoms;
    for (;;) {
        bool done = false;
        if (lane == 0) {
            const uint32_t limit_bits = (uint32_t)(hl + 1) * kWalkHalfWords * 32u;
            uint32_t v = ring_peek6(ring, pos + fullbits);
            while (atom < natoms && pos + kWalkMarginBits < limit_bits) {
                if ((group & (kCheckEvery - 1)) == 0) {
                    XtcCheckpoint cp;
                    cp.bitpos = pos;
                    cp.atom = (uint32_t)atom;
                    cp.state = (uint32_t)smallidx | ((3u * smalls) << 8);
                    out[nc++] = cp;
                }
                // flag bit, then (if set) the 5-bit run code: run = code - code % 3, table entry += code % 3 - 1
                const uint32_t flag = v >> 5, code = v & 31u;
                const uint32_t third = (code * 11u) >> 5;                  // code / 3 for code < 32
                const int step = (int)code - 3 * (int)third - 1;           // code % 3 - 1
                smalls = flag ? third : smalls;
                pos += fullbits + 1u + 5u * flag + smalls * (uint32_t)smallidx;   // the smalls use the old entry
                atom += (int)smalls + 1;
                smallidx += flag ? step : 0;
                ++group;
                if ((uint32_t)(smallidx - kFirstIdx) > (uint32_t)(kLastIdx - kFirstIdx) || pos > end_bits) {
                    bad = true;
                    break;
                }
                v = ring_peek6(ring, pos + fullbits);
            }
            done = bad || atom >= natoms;
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done) break;
        // pos lies in half hl by now (the margin is far smaller than a half): overwrite the older half
        ++hl;
        __syncwarp();
        load_half(hl);
        __syncwarp();
    }
    if (lane == 0) {
        n_checks[f] = bad ? -1 : nc;      // a bad frame is not decoded
        if (bad && bad_frames) atomicAdd(bad_frames, 1);
    }
}

__global__ void __launch_bounds__(128) xtc_decode_kernel(const uint8_t* __restrict__ data,
                                                         const CgbXtcFrame* __restrict__ frames, int64_t n_frames,
                                                         int64_t n_atoms, const XtcCheckpoint* __restrict__ checks,
                                                         const int32_t* __restrict__ n_checks,
                                                         float* __restrict__ xyz, int32_t* __restrict__ bad_frames) {
    const int64_t f = blockIdx.y;
    const int64_t cpf = checks_per_frame(n_atoms);
    const CgbXtcFrame fr = frames[f];
    float* out = xyz + (size_t)f * n_atoms * 3;
    if (fr.natoms <= 9) {
        if (blockIdx.x == 0 && threadIdx.x == 0) decode_raw(data + fr.offset, fr.natoms, out);
        return;
    }
    const int nc = n_checks[f];     // -1: the walk found the stream invalid, the frame stays undecoded
    const FrameCoding c = frame_coding(fr);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nc; k += (int64_t)gridDim.x * blockDim.x) {
        const XtcCheckpoint cp = checks[f * cpf + k];
        BitReader r(data + fr.offset, (uint64_t)fr.nbytes, cp.bitpos);
        GroupState st = {(int)cp.atom, (int)(cp.state & 0xff), (int)(cp.state >> 8)};
        if (decode_groups(r, fr, c, st, kCheckEvery, out) < 0 && bad_frames) atomicAdd(bad_frames, 1);
    }
}

// ------------------------------------------------------------------------------------------ host
struct BitWriter {
    std::vector<uint8_t>& out;
    uint64_t buf = 0;
    int have = 0;
    explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
    void put(int n, uint32_t v) {  // n <= 32
        if (n == 0) return;
        buf = (buf << n) | (uint64_t)(v & ((n == 32) ? 0xffffffffu : ((1u << n) - 1u)));
        have += n;
        while (have >= 8) {
            have -= 8;
            out.push_back((uint8_t)(buf >> have));
        }
    }
    void flush() {
        if (have > 0) {
            out.push_back((uint8_t)(buf << (8 - have)));
            have = 0;
        }
    }
};

// xdrfile sendints: mixed-radix number of three values, bytes least significant first.
static void put_triple(BitWriter& w, int nbits, const uint32_t sizes[3], const uint32_t nums[3]) {
    uint8_t bytes[16];
    int nb = 0;
    uint32_t tmp = nums[0];
    do {
        bytes[nb++] = tmp & 0xff;
        tmp >>= 8;
    } while (tmp != 0);
    for (int i = 1; i < 3; ++i) {
        uint64_t carry = nums[i];
        int b = 0;
        for (; b < nb; ++b) {
            carry = (uint64_t)bytes[b] * sizes[i] + carry;
            bytes[b] = carry & 0xff;
            carry >>= 8;
        }
        while (carry != 0) {
            bytes[b++] = carry & 0xff;
            carry >>= 8;
        }
        nb = b;
    }
    if (nbits >= nb * 8) {
        for (int i = 0; i < nb; ++i) w.put(8, bytes[i]);
        int rest = nbits - nb * 8;
        while (rest > 0) {
            const int n = std::min(rest, 32);
            w.put(n, 0);
            rest -= n;
        }
    } else {
        for (int i = 0; i < nb - 1; ++i) w.put(8, bytes[i]);
        w.put(nbits - (nb - 1) * 8, bytes[nb - 1]);
    }
}

static int product_bits(const uint32_t s[3]) {
    // bit length of s0*s1*s2 through byte-array multiplication (as xdrfile's sizeofints)
    uint8_t bytes[16] = {1};
    int nb = 1;
    for (int i = 0; i < 3; ++i) {
        uint64_t carry = 0;
        int b = 0;
        for (; b < nb; ++b) {
            carry = (uint64_t)bytes[b] * s[i] + carry;
            bytes[b] = carry & 0xff;
            carry >>= 8;
        }
        while (carry != 0) {
            bytes[b++] = carry & 0xff;
            carry >>= 8;
        }
        nb = b;
    }
    return (nb - 1) * 8 + bit_length_u32(bytes[nb - 1]);
}

static void put_be32(std::vector<uint8_t>& out, uint32_t v) {
    out.push_back(v >> 24);
    out.push_back(v >> 16);
    out.push_back(v >> 8);
    out.push_back(v);
}
static void put_bef(std::vector<uint8_t>& out, float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    put_be32(out, b);
}

}  // namespace cgb

extern "C" int cgb_xtc_scan(const uint8_t* data, int64_t n_bytes, int64_t max_frames, CgbXtcFrame* frames,
                            float* box, float* time, int32_t* step, int64_t* n_frames, int32_t* n_atoms) {
    using namespace cgb;
    if (!data || n_bytes < 0 || !n_frames || !n_atoms) return CGB_EINVAL;
    int64_t pos = 0, nf = 0;
    int32_t natoms0 = -1;
    while (pos < n_bytes) {
        if (pos + 56 > n_bytes) return CGB_EINVAL;
        if (load_be32(data + pos) != 1995u) return CGB_EINVAL;
        const int32_t natoms = (int32_t)load_be32(data + pos + 4);
        if (natoms0 < 0) natoms0 = natoms;
        if (natoms != natoms0 || natoms < 0) return CGB_EINVAL;
        CgbXtcFrame fr;
        memset(&fr, 0, sizeof(fr));
        fr.natoms = natoms;
        int64_t next;
        if (natoms <= 9) {
            fr.offset = pos + 56;
            fr.nbytes = 12 * natoms;
            fr.precision = 1.0f;
            next = pos + 56 + 12 * (int64_t)natoms;
        } else {
            if (pos + 92 > n_bytes) return CGB_EINVAL;
            uint32_t pb = load_be32(data + pos + 56);
            memcpy(&fr.precision, &pb, 4);
            for (int d = 0; d < 3; ++d) {
                fr.minint[d] = (int32_t)load_be32(data + pos + 60 + 4 * d);
                fr.maxint[d] = (int32_t)load_be32(data + pos + 72 + 4 * d);
            }
            fr.smallidx = (int32_t)load_be32(data + pos + 84);
            fr.nbytes = (int32_t)load_be32(data + pos + 88);
            fr.offset = pos + 92;
            if (fr.smallidx < kFirstIdx || fr.smallidx > kLastIdx || fr.nbytes < 0) return CGB_EINVAL;
            if (!(fr.precision > 0.0f) || !std::isfinite(fr.precision)) return CGB_EINVAL;
            for (int d = 0; d < 3; ++d) {
                // sizeint = maxint - minint + 1 must be a positive 32-bit count (it is a divisor in the decoder)
                const int64_t size = (int64_t)fr.maxint[d] - (int64_t)fr.minint[d] + 1;
                if (size < 1 || size > 0xffffffffLL) return CGB_EINVAL;
            }
            if ((int64_t)fr.nbytes * 8 >= ((int64_t)1 << 32)) return CGB_ERANGE;  // 32-bit bit positions
            next = pos + 92 + (((int64_t)fr.nbytes + 3) / 4) * 4;
        }
        if (next > n_bytes) return CGB_EINVAL;
        if (frames) {
            if (nf >= max_frames) return CGB_ERANGE;
            frames[nf] = fr;
            if (step) step[nf] = (int32_t)load_be32(data + pos + 8);
            if (time) {
                uint32_t tb = load_be32(data + pos + 12);
                memcpy(time + nf, &tb, 4);
            }
            if (box)
                for (int k = 0; k < 9; ++k) {
                    uint32_t bb = load_be32(data + pos + 16 + 4 * k);
                    memcpy(box + nf * 9 + k, &bb, 4);
                }
        }
        ++nf;
        pos = next;
    }
    *n_frames = nf;
    *n_atoms = natoms0 < 0 ? 0 : natoms0;
    return CGB_OK;
}

namespace cgb {

// Run fn(f) for f in [0, n) on the host cores (frames of a trajectory are independent).
template <class Fn>
static void parallel_frames(int64_t n, Fn fn) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int64_t workers = std::min<int64_t>(std::min<int64_t>(hw, 64), n);
    if (workers <= 1) {
        for (int64_t f = 0; f < n; ++f) fn(f);
        return;
    }
    std::atomic<int64_t> next(0);
    std::vector<std::thread> pool;
    pool.reserve((size_t)workers);
    for (int64_t w = 0; w < workers; ++w)
        pool.emplace_back([&]() {
            for (int64_t f = next.fetch_add(1); f < n; f = next.fetch_add(1)) fn(f);
        });
    for (auto& t : pool) t.join();
}

}  // namespace cgb

extern "C" int cgb_xtc_decode_host(const uint8_t* data, const CgbXtcFrame* frames, int64_t n_frames,
                                   int64_t n_atoms, float* xyz) {
    using namespace cgb;
    if (n_frames < 0 || n_atoms < 0) return CGB_EINVAL;
    if (n_frames == 0) return CGB_OK;
    if (!data || !frames || !xyz) return CGB_EINVAL;
    for (int64_t f = 0; f < n_frames; ++f)
        if (frames[f].natoms != n_atoms) return CGB_EINVAL;
    for (int64_t f = 0; f < n_frames; ++f) {   // a table not made by cgb_xtc_scan: the same checks again
        const CgbXtcFrame& fr = frames[f];
        if (fr.natoms <= 9) continue;
        if (fr.smallidx < kFirstIdx || fr.smallidx > kLastIdx || fr.nbytes < 0 || !(fr.precision > 0.0f)) return CGB_EINVAL;
        for (int d = 0; d < 3; ++d)
            if ((int64_t)fr.maxint[d] - (int64_t)fr.minint[d] + 1 < 1) return CGB_EINVAL;
    }
    // frames are independent: decode them on all host cores
    std::atomic<int64_t> bad(0);
    parallel_frames(n_frames, [&](int64_t f) {
        float* out = xyz + (size_t)f * n_atoms * 3;
        if (frames[f].natoms <= 9)
            decode_raw(data + frames[f].offset, frames[f].natoms, out);
        else if (!decode_frame(data + frames[f].offset, frames[f], out))
            bad.fetch_add(1);
    });
    return bad.load() ? CGB_EINVAL : CGB_OK;   // truncated or corrupt coordinate stream
}

extern "C" int64_t cgb_xtc_workspace_bytes(int64_t n_frames, int64_t n_atoms) {
    if (n_frames < 0 || n_atoms < 0) return 0;
    return n_frames * cgb::checks_per_frame(n_atoms) * (int64_t)sizeof(cgb::XtcCheckpoint) + n_frames * 4 + 64;
}

extern "C" int cgb_xtc_decode(const uint8_t* data, int64_t n_bytes, const CgbXtcFrame* frames, int64_t n_frames,
                              int64_t n_atoms, float* xyz, void* workspace, int64_t workspace_bytes,
                              int32_t* bad_frames, void* stream) {
    using namespace cgb;
    if (n_frames < 0 || n_atoms < 0 || n_bytes < 0) return CGB_EINVAL;
    if (n_frames == 0 || n_atoms == 0) return CGB_OK;
    if (!data || !frames || !xyz || !workspace) return CGB_EINVAL;
    if (workspace_bytes < cgb_xtc_workspace_bytes(n_frames, n_atoms)) return CGB_ERANGE;
    if ((reinterpret_cast<uintptr_t>(data) & 3u) || (reinterpret_cast<uintptr_t>(workspace) & 15u)) return CGB_EALIGN;
    if (n_frames > 65535) return CGB_ERANGE;  // one launch decodes at most 65535 frames: call per block of frames
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XtcCheckpoint* checks = static_cast<XtcCheckpoint*>(workspace);
    int32_t* n_checks = reinterpret_cast<int32_t*>(checks + n_frames * checks_per_frame(n_atoms));
    xtc_walk_kernel<<<(unsigned)n_frames, 32, 0, st>>>(data, n_bytes, frames, n_frames, n_atoms, checks,
                                                                    n_checks, bad_frames);
    CGB_CUDA_TRY(cudaGetLastError());
    const int64_t max_checks = checks_per_frame(n_atoms);
    const unsigned bx = (unsigned)std::min<int64_t>((max_checks + 127) / 128, 1024);
    dim3 grid(bx, (unsigned)n_frames);
    xtc_decode_kernel<<<grid, 128, 0, st>>>(data, frames, n_frames, n_atoms, checks, n_checks, xyz, bad_frames);
    return (int)cudaGetLastError();
}

// Encode frames into a caller-provided buffer; returns the bytes written through *n_written, or
// CGB_ERANGE when `capacity` is too small (call again with a larger buffer).
namespace cgb {

// One frame (header + compressed coordinates) appended to `buf`; CGB_OK or CGB_ERANGE.
static int encode_frame(const float* x, const float* box9, float time, int64_t f, int64_t n_atoms, float precision,
                        std::vector<uint8_t>& buf) {
    std::vector<int> ip((size_t)n_atoms * 3);
    {
        put_be32(buf, 1995);
        put_be32(buf, (uint32_t)n_atoms);
        put_be32(buf, (uint32_t)f);
        put_bef(buf, time);
        for (int k = 0; k < 9; ++k) put_bef(buf, box9 ? box9[k] : 0.0f);
        put_be32(buf, (uint32_t)n_atoms);
        if (n_atoms <= 9) {
            for (int64_t i = 0; i < 3 * n_atoms; ++i) put_bef(buf, x[i]);
            return CGB_OK;
        }
        int minint[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, maxint[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
        for (int64_t i = 0; i < n_atoms; ++i)
            for (int d = 0; d < 3; ++d) {
                const float lf = x[i * 3 + d] * precision;
                if (std::fabs(lf) > 2e9f) return CGB_ERANGE;  // does not fit an int at this precision
                const int v = lf >= 0 ? (int)(lf + 0.5f) : (int)(lf - 0.5f);
                ip[(size_t)i * 3 + d] = v;
                minint[d] = std::min(minint[d], v);
                maxint[d] = std::max(maxint[d], v);
            }
        uint32_t sizeint[3];
        bool wide = false;
        for (int d = 0; d < 3; ++d) {
            sizeint[d] = (uint32_t)(maxint[d] - minint[d] + 1);
            if (sizeint[d] > 0xffffffu) wide = true;
        }
        int bitsizeint[3] = {0, 0, 0}, bitsize = 0;
        if (wide)
            for (int d = 0; d < 3; ++d) bitsizeint[d] = bit_length_u32(sizeint[d]);
        else
            bitsize = product_bits(sizeint);
        // small-difference table entry: large enough for most neighbour differences, constant per frame
        std::vector<int> diffs;
        diffs.reserve((size_t)n_atoms);
        for (int64_t i = 1; i < n_atoms; ++i) {
            int m = 0;
            for (int d = 0; d < 3; ++d) m = std::max(m, std::abs(ip[(size_t)i * 3 + d] - ip[(size_t)(i - 1) * 3 + d]));
            diffs.push_back(m);
        }
        std::nth_element(diffs.begin(), diffs.begin() + diffs.size() * 3 / 4, diffs.end());
        const int typical = diffs[diffs.size() * 3 / 4];
        int smallidx = kFirstIdx;
        while (smallidx < kLastIdx && h_magicints[smallidx] / 2 <= typical) ++smallidx;
        const int smallnum = h_magicints[smallidx] / 2;
        const uint32_t sizesmall[3] = {(uint32_t)h_magicints[smallidx], (uint32_t)h_magicints[smallidx],
                                       (uint32_t)h_magicints[smallidx]};
        const int top = h_magicints[smallidx] - 1 - smallnum;
        auto fits = [&](const int* a, const int* ref) {
            for (int d = 0; d < 3; ++d) {
                const int dl = a[d] - ref[d];
                if (dl < -smallnum || dl > top) return false;
            }
            return true;
        };

        std::vector<uint8_t> payload;
        BitWriter w(payload);
        int prevrun = -1;
        int64_t i = 0;
        while (i < n_atoms) {
            // a run is [a0, a1, a2, ...]: a1 is stored in full, a0 relative to a1, a2 relative to a0,
            // every later atom relative to its predecessor (the decoder interchanges a0 and a1)
            int m = 0;
            if (i + 1 < n_atoms && fits(&ip[(size_t)i * 3], &ip[(size_t)(i + 1) * 3])) {
                m = 1;
                while (m < 8 && i + 1 + m < n_atoms) {
                    const int* a = &ip[(size_t)(i + 1 + m) * 3];
                    const int* ref = (m == 1) ? &ip[(size_t)i * 3] : &ip[(size_t)(i + m) * 3];
                    if (!fits(a, ref)) break;
                    ++m;
                }
            }
            const int* full = m ? &ip[(size_t)(i + 1) * 3] : &ip[(size_t)i * 3];
            uint32_t rel[3];
            for (int d = 0; d < 3; ++d) rel[d] = (uint32_t)(full[d] - minint[d]);
            if (wide)
                for (int d = 0; d < 3; ++d) w.put(bitsizeint[d], rel[d]);
            else
                put_triple(w, bitsize, sizeint, rel);
            const int run = 3 * m;
            if (run != prevrun) {
                w.put(1, 1);
                w.put(5, (uint32_t)(run + 1));  // +1: keep the small-difference table entry
                prevrun = run;
            } else {
                w.put(1, 0);
            }
            for (int k = 0; k < m; ++k) {
                const int* a = (k == 0) ? &ip[(size_t)i * 3] : &ip[(size_t)(i + 1 + k) * 3];
                const int* ref = (k == 0) ? &ip[(size_t)(i + 1) * 3]
                                          : (k == 1 ? &ip[(size_t)i * 3] : &ip[(size_t)(i + k) * 3]);
                uint32_t sv[3];
                for (int d = 0; d < 3; ++d) sv[d] = (uint32_t)(a[d] - ref[d] + smallnum);
                put_triple(w, smallidx, sizesmall, sv);
            }
            i += m ? m + 1 : 1;
        }
        w.flush();
        put_bef(buf, precision);
        for (int d = 0; d < 3; ++d) put_be32(buf, (uint32_t)minint[d]);
        for (int d = 0; d < 3; ++d) put_be32(buf, (uint32_t)maxint[d]);
        put_be32(buf, (uint32_t)smallidx);
        put_be32(buf, (uint32_t)payload.size());
        buf.insert(buf.end(), payload.begin(), payload.end());
        while (buf.size() % 4) buf.push_back(0);
    }
    return CGB_OK;
}

}  // namespace cgb

extern "C" int cgb_xtc_encode(const float* xyz, const float* box, const float* time, int64_t n_frames,
                              int64_t n_atoms, float precision, uint8_t* out, int64_t capacity,
                              int64_t* n_written) {
    using namespace cgb;
    if (n_frames < 0 || n_atoms < 0 || !n_written || (n_frames && n_atoms && !xyz) || precision <= 0)
        return CGB_EINVAL;
    // frames are independent: encode them on all host cores, then concatenate
    std::vector<std::vector<uint8_t>> parts((size_t)n_frames);
    std::vector<int> status((size_t)n_frames, CGB_OK);
    parallel_frames(n_frames, [&](int64_t f) {
        status[(size_t)f] = encode_frame(xyz + (size_t)f * n_atoms * 3, box ? box + f * 9 : nullptr,
                                         time ? time[f] : (float)f, f, n_atoms, precision, parts[(size_t)f]);
    });
    size_t total = 0;
    for (int64_t f = 0; f < n_frames; ++f) {
        if (status[(size_t)f] != CGB_OK) return status[(size_t)f];
        total += parts[(size_t)f].size();
    }
    *n_written = (int64_t)total;
    if ((int64_t)total > capacity || (!out && total)) return CGB_ERANGE;
    size_t at = 0;
    for (int64_t f = 0; f < n_frames; ++f) {
        if (!parts[(size_t)f].empty()) memcpy(out + at, parts[(size_t)f].data(), parts[(size_t)f].size());
        at += parts[(size_t)f].size();
    }
    return CGB_OK;
}
