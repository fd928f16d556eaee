// circall_b200 — kernels shared by the session pipeline (cg_session.cu) and the collector surface
// (cg_api.cu): read packing and the per-thread strand set-up.
#pragma once
#include "cg_common.h"
#include "cg_pipeline.cuh"
#include "cg_pack_swar.h"

namespace cgk {
using namespace cg;

// Packed-read record of one read: WT = W2f + WMf uint64 words, read-major (record r at pk + r * WT):
//   W2f code words (32 bases each, base j at bits 2j) then WMf N-mask words (64 positions each).
// Positions past the read's end carry code 0 / mask 1.
struct PackGeom {
    int W2f, WMf, WT;
    __host__ __device__ static PackGeom make(int maxLen) {
        PackGeom g;
        g.W2f = (maxLen + 31) / 32; if (g.W2f < 1) g.W2f = 1;
        g.WMf = (g.W2f + 1) / 2;
        g.WT = g.W2f + g.WMf;
        return g;
    }
    // shared-memory words per thread for the read view, padded to an odd count so that lanes reading
    // the same word index of their own reads hit distinct banks
    __host__ __device__ int chunk() const { return (W2f + 1 + WMf + 1) | 1; }
};

// 4 ASCII bases -> 4 x 2-bit codes (bits 0..7), N flags (bits 0..3) and "other character" flags
__device__ __forceinline__ void pack4(uint32_t w, uint32_t& code8, uint32_t& n4, uint32_t& bad4) {
    uint32_t u = w & 0xDFDFDFDFu;
    uint32_t t = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    code8 = (t | (t >> 6) | (t >> 12) | (t >> 18)) & 0xFFu;
    uint32_t isb = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    uint32_t isn = __vcmpeq4(u, 0x4E4E4E4Eu);
    uint32_t nb = ~isb & 0x01010101u;               // not a base
    uint32_t ob = ~(isb | isn) & 0x01010101u;       // neither base nor n/N
    n4 = (nb | (nb >> 7) | (nb >> 14) | (nb >> 21)) & 0xFu;
    bad4 = (ob | (ob >> 7) | (ob >> 14) | (ob >> 21)) & 0xFu;
}

// the general formulation of one 32-base word: any character, N flags, partial words
constexpr int PACK_THREADS = 256;
static inline unsigned pack_grid(uint64_t n_reads, const PackGeom& g) {
    uint64_t rpb = PACK_THREADS / g.W2f;
    return (unsigned)((n_reads + rpb - 1) / rpb);
}
struct Word8 { uint32_t v[8]; };
struct PackedWord { uint64_t code; uint32_t nmask, bad; };
static __device__ __noinline__ PackedWord pack_word_general(Word8 in, int nb) {
    uint64_t code = 0; uint32_t nmask = 0, bad = 0;
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        uint32_t c8, n4, b4;
        pack4(in.v[x], c8, n4, b4);
        code |= (uint64_t)c8 << (8 * x);
        nmask |= n4 << (4 * x);
        bad |= b4 << (4 * x);
    }
    if (nb < 32) {
        uint32_t keep = (1u << nb) - 1;
        code &= (1ULL << (2 * nb)) - 1;
        nmask = (nmask & keep) | ~keep;
        bad &= keep;
    }
    // an N position carries code 0: spread the 32 mask bits to both bits of each 2-bit field
    uint64_t x = nmask;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x << 2)) & 0x3333333333333333ULL;
    x = (x | (x << 1)) & 0x5555555555555555ULL;
    code &= ~(x | (x << 1));
    return PackedWord{code, nmask, bad};
}

// K1: one thread per (read, 32-base word).  Consecutive threads read consecutive 32-byte pieces of
// the concatenated ASCII buffer (aligned 4-byte loads + funnel shift) and write consecutive words
// of the read-major packed records, so both sides are coalesced.
//   reads r = 2 * pair + mate;  status bit0: character outside ACGTNacgtn, bit1: read longer than maxLen
// The kernel is bound by its instruction count, so a word takes the short way when all of its characters are
// bases (ACGTacgt): four codes per 32-bit word with SWAR arithmetic, the characters the codes stand for rebuilt and
// compared with the input (one test per thread), the codes gathered by a multiply.  A word with any other character
// (N, or a character the reference rejects) goes through pack_word_general.  GENERAL = true: that way for every word
// (what cg_pack_reads' check compares the short way with).
template <bool GENERAL>
__device__ __forceinline__ void pack_body(const unsigned char* __restrict__ b1, const uint64_t* __restrict__ off1,
                       const unsigned char* __restrict__ b2, const uint64_t* __restrict__ off2,
                       uint64_t n_pairs, PackGeom g, int maxLen, uint64_t* __restrict__ pk,
                       uint16_t* __restrict__ lens, int* status) {
    // a block of PACK_THREADS threads takes PACK_THREADS / W2f whole reads (launch with pack_grid): no 64-bit division
    float inv = __fdividef(1.0f, (float)g.W2f);
    int rl = (int)(((float)threadIdx.x + 0.5f) * inv);                         // threadIdx.x / W2f: the half keeps the quotient
    int rpb = (int)(((float)PACK_THREADS + 0.5f) * inv);                       // away from an integer, so the approximation is exact
    int j = (int)threadIdx.x - rl * g.W2f;
    uint64_t r = blockIdx.x * (uint64_t)rpb + rl;
    if (rl >= rpb || r >= 2 * n_pairs) return;
    uint64_t i = r >> 1;
    const unsigned char* base = (r & 1) ? b2 : b1;
    const uint64_t* off = (r & 1) ? off2 : off1;
    uint64_t o0 = off[i], o1 = off[i + 1];
    int64_t len = (int64_t)(o1 - o0);
    if (len > maxLen) { atomicOr(status, 2); len = maxLen; }
    if (j == 0) lens[r] = (uint16_t)len;
    int nb = (int)len - 32 * j;
    nb = nb < 0 ? 0 : (nb > 32 ? 32 : nb);
    uint64_t code = 0; uint32_t nmask = 0xFFFFFFFFu, bad = 0;
    if (nb > 0) {
        const unsigned char* p = base + o0 + 32 * j;
        uintptr_t a = (uintptr_t)p;
        const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
        int sh = (int)(a & 3) * 8;
        int need = (int)(a & 3) + nb;                 // bytes from q that hold data
        uint32_t w[9];
        Word8 in;
        uint32_t (&v)[8] = in.v;
#pragma unroll
        for (int x = 0; x < 9; ++x) w[x] = (4 * x < need) ? __ldg(q + x) : 0u;
#pragma unroll
        for (int x = 0; x < 8; ++x) v[x] = __funnelshift_r(w[x], w[x + 1], sh);
        bool general = GENERAL;
        if (!GENERAL) {
            uint32_t diff = 0, m[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                uint32_t c = (4 * x + 4 <= nb) ? v[x] : 0x41414141u;     // words past the end (and a partial one) count as "AAAA"
                m[x] = cg_pack4_bases(c, diff);                                // the four codes side by side in the top byte
            }
            uint32_t lo = __byte_perm(__byte_perm(m[0], m[1], 0x0073), __byte_perm(m[2], m[3], 0x0073), 0x5410);
            uint32_t hi = __byte_perm(__byte_perm(m[4], m[5], 0x0073), __byte_perm(m[6], m[7], 0x0073), 0x5410);
            code = (uint64_t)lo | ((uint64_t)hi << 32);
            nmask = nb < 32 ? ~((1u << nb) - 1) : 0u;
            general = diff != 0;
            if (!general && (nb & 3)) {                                    // the bases of a partial word, one by one
                int b0 = nb & ~3;
                for (int y = b0; y < nb; ++y) {
                    uint32_t c = __ldg(p + y);
                    uint32_t t2 = ((c >> 1) ^ (c >> 2)) & 3u;
                    if (((0x54474341u >> (8 * t2)) & 0xFFu) != (c & 0xDFu)) general = true;
                    code |= (uint64_t)t2 << (2 * y);
                }
            }
        }
        if (general) { PackedWord o = pack_word_general(in, nb); code = o.code; nmask = o.nmask; bad = o.bad; }
    }
    if (bad) atomicOr(status, 1);
    uint64_t* rec = pk + r * (uint64_t)g.WT;
    rec[j] = code;
    uint32_t* mw = (uint32_t*)(rec + g.W2f + (j >> 1));
    mw[j & 1] = nmask;
    if ((g.W2f & 1) && j == g.W2f - 1) mw[1] = 0xFFFFFFFFu;
}
static __global__ void k_pack(const unsigned char* __restrict__ b1, const uint64_t* __restrict__ off1,
                       const unsigned char* __restrict__ b2, const uint64_t* __restrict__ off2,
                       uint64_t n_pairs, PackGeom g, int maxLen, uint64_t* __restrict__ pk,
                       uint16_t* __restrict__ lens, int* status) {
    pack_body<false>(b1, off1, b2, off2, n_pairs, g, maxLen, pk, lens, status);
}
static __global__ void k_pack_general(const unsigned char* __restrict__ b1, const uint64_t* __restrict__ off1,
                       const unsigned char* __restrict__ b2, const uint64_t* __restrict__ off2,
                       uint64_t n_pairs, PackGeom g, int maxLen, uint64_t* __restrict__ pk,
                       uint16_t* __restrict__ lens, int* status) {
    pack_body<true>(b1, off1, b2, off2, n_pairs, g, maxLen, pk, lens, status);
}

// set up this thread's read view in shared memory
__device__ __forceinline__ ReadView make_view(const uint64_t* __restrict__ rec, const PackGeom& g, int L, uint64_t* smem) {
    return build_view(rec, 1, L, g.W2f, g.WMf, smem + (size_t)threadIdx.x * g.chunk());
}

} // namespace cgk
