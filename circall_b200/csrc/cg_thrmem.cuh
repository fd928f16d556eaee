// circall_b200 — per-thread scratch and small helpers shared by the thread-per-read tiers (cg_pre.cuh, cg_thr.cuh).
//
// CG_HD throughout: tests/emu runs the same functions on the host in front of the emulated warp tiers.
#pragma once
#include "cg_pipeline.cuh"

namespace cgt {
using namespace cg;

#if defined(__CUDACC__) && !defined(CG_HOST_EMU)
typedef uint2 u32x2;
#else
struct u32x2 { uint32_t x, y; };
#endif

CG_HD uint32_t fsr(uint32_t lo, uint32_t hi, int sh) {      // funnel shift right, sh in 0..31
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// Per-thread scratch of 32-bit words, word j at base[j * stride] (stride = threads per block on the device, so
// that the lanes of a warp touching the same word index hit 32 different banks; 1 on the host).
//   [0, 2 W2)        forward code words, 16 bases each (W2 = 64-bit words incl. one zero pad word)
//   [2 W2, 4 W2)     reverse-complement strand, same layout
struct TMem {
    uint32_t* base; int stride; int W2;
    CG_HD uint32_t ld(int j) const { return base[(size_t)j * stride]; }
    CG_HD void st(int j, uint32_t v) const { base[(size_t)j * stride] = v; }
    // 32-base window of strand s at oriented position i (i < L)
    CG_HD uint64_t chunk(int s, int i) const {
        int p = (s ? 2 * W2 : 0) + (i >> 4), sh = (i & 15) * 2;
        uint32_t a = ld(p), b = ld(p + 1), c = ld(p + 2);
        return (uint64_t)fsr(a, b, sh) | ((uint64_t)fsr(b, c, sh) << 32);
    }
};

// transcriptAtPosition(p) from the second half of p's text block (mask, tid0, start0)
CG_HD uint32_t tid_from(const u32x4& t, uint32_t p) {
    uint64_t mask = (uint64_t)t.x | ((uint64_t)t.y << 32);
    uint32_t o = p & 63;
    return t.z + (uint32_t)popc64(o ? (mask & ((1ULL << o) - 1)) : 0ULL);
}
CG_HD const u32x2* mask_ptr(const IndexView& ix, uint32_t b) { return reinterpret_cast<const u32x2*>(&ix.text[2 * (uint64_t)b + 1]); }
CG_HD uint64_t half_codes(const u32x4& b, int half) { return half ? ((uint64_t)b.z | ((uint64_t)b.w << 32)) : ((uint64_t)b.x | ((uint64_t)b.y << 32)); }

// the reverse-complement strand of the read staged in m (forward words at [0, 2 W2)): word pair j = 32 bases from oriented
// position 32 j, zero past the end
template <class M>
CG_HD void make_rc(const M& m, int L) {
    for (int j = 0; j < m.W2; ++j) {
        int a = L - 32 - 32 * j;
        uint64_t v = 0;
        if (a >= 0) v = rc64(m.chunk(0, a));
        else if (32 + a > 0) v = rc64(m.chunk(0, 0)) >> (2 * (-a));
        m.st(2 * m.W2 + 2 * j, (uint32_t)v); m.st(2 * m.W2 + 2 * j + 1, (uint32_t)(v >> 32));
    }
}

} // namespace cgt
