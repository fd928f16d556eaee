// circall_b200 — kernels shared by the session pipeline (cg_session.cu) and the collector surface
// (cg_api.cu): read packing and the per-thread strand set-up.
#pragma once
#include "cg_common.h"
#include "cg_pipeline.cuh"

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

// K1: one thread per (read, 32-base word).  Consecutive threads read consecutive 32-byte pieces of
// the concatenated ASCII buffer (aligned 4-byte loads + funnel shift) and write consecutive words
// of the read-major packed records, so both sides are coalesced.
//   reads r = 2 * pair + mate;  status bit0: character outside ACGTNacgtn, bit1: read longer than maxLen
static __global__ void k_pack(const unsigned char* __restrict__ b1, const uint64_t* __restrict__ off1,
                       const unsigned char* __restrict__ b2, const uint64_t* __restrict__ off2,
                       uint64_t n_pairs, PackGeom g, int maxLen, uint64_t* __restrict__ pk,
                       uint16_t* __restrict__ lens, int* status) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t r = t / (uint64_t)g.W2f;
    int j = (int)(t - r * (uint64_t)g.W2f);
    if (r >= 2 * n_pairs) return;
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
#pragma unroll
        for (int x = 0; x < 9; ++x) w[x] = (4 * x < need) ? __ldg(q + x) : 0u;
        nmask = 0;
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            uint32_t v = __funnelshift_r(w[x], w[x + 1], sh);
            uint32_t c8, n4, b4;
            pack4(v, c8, n4, b4);
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
        // an N position carries code 0
        uint64_t nm2 = 0;
        {   // spread the 32 mask bits to the even bit positions of a 64-bit word
            uint64_t x = nmask;
            x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL;
            x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
            x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL;
            x = (x | (x << 2)) & 0x3333333333333333ULL;
            x = (x | (x << 1)) & 0x5555555555555555ULL;
            nm2 = x | (x << 1);
        }
        code &= ~nm2;
    }
    if (bad) atomicOr(status, 1);
    uint64_t* rec = pk + r * (uint64_t)g.WT;
    rec[j] = code;
    uint32_t* mw = (uint32_t*)(rec + g.W2f + (j >> 1));
    mw[j & 1] = nmask;
    if ((g.W2f & 1) && j == g.W2f - 1) mw[1] = 0xFFFFFFFFu;
}

// set up this thread's read view in shared memory
__device__ __forceinline__ ReadView make_view(const uint64_t* __restrict__ rec, const PackGeom& g, int L, uint64_t* smem) {
    return build_view(rec, 1, L, g.W2f, g.WMf, smem + (size_t)threadIdx.x * g.chunk());
}

} // namespace cgk
