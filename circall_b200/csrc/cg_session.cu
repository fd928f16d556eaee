// circall_b200 — the pipeline session: Circall_wt and Circall_bsj per-pair bodies fused on the device.
//
// One batch = one pass of (paths relative to /root/reference/Circall_v1.0.1/):
//   k_pack        K1  2-bit packing + N mask                       SACollectorCircall.hpp:117-128,243-260
//   k_wt          K2-K5 collector on txIdx for both mates, K6 flags src/Circall_wt.cpp:275-342,485-517
//   k_wt_commit   K6/K8 export decision + scalar counters          src/Circall_wt.cpp:350-372,467-470,525-547,642-645
//   k_flag_*      K6  ordered stream compaction of the export list
//   k_bsj         K2-K5 collector on bsjIdx for every exported record, K7 span filter, K8 counters
//                                                                   src/Circall_bsj.cpp:300-399,488-491
// Nothing here falls back to the host: without a CUDA device session creation fails.
#include "cg_kernels.cuh"
#include "cg_warp.cuh"
#include "cg_pre.cuh"
#include "cg_thr.cuh"
#include <algorithm>
#include <mutex>
#include <dlfcn.h>
#include <vector>
#include <cstring>
#include <cstddef>
#include <cstdlib>
#include <cstdio>

using namespace cg;
using namespace cgk;
using cgi::GrowBuf;
using cgi::PinBuf;

// page-locked host array that only ever grows; resize() does not keep the old contents
template <class T>
struct PinVec {
    PinBuf b; uint64_t n = 0;
    cudaError_t resize(uint64_t m) { cudaError_t e = b.reserve(m * sizeof(T) + 16); if (e == cudaSuccess) n = m; return e; }
    T* data() const { return b.as<T>(); }
    uint64_t size() const { return n; }
    T& operator[](uint64_t i) const { return b.as<T>()[i]; }
    T* begin() const { return data(); }
    T* end() const { return data() + n; }
    void release() { b.release(); n = 0; }
};

namespace {

enum { ST_BADBASE = 1, ST_TOOLONG = 2, ST_CAPACITY = 4 };
enum { C_WT_OBS = 0, C_WT_MAP, C_WT_HITS, C_WT_UPPER, C_BSJ_OBS, C_BSJ_MAP, C_BSJ_HITS, C_BSJ_UPPER, C_NUM };

// device-side per-batch header
struct BatchHdr {
    unsigned long long n_junc;     // rows the span filter produced (may exceed the buffer capacity)
    unsigned long long n_eq;       // multi-transcript classes recorded
    unsigned long long n_eq_tids;
    uint32_t n_exp;
    int status;
    uint32_t n_ovf_wt;             // mates / records the fast path handed to the general kernels
    uint32_t n_ovf_bsj;
    uint32_t n_t2_wt;              // mates the front kernel declined / records handed to the lockstep tier
    uint32_t n_t2_bsj;
    uint32_t n_t3_wt;              // mates / records the lockstep thread tier handed to the warp kernels
    uint32_t n_t3_bsj;
    uint32_t n_cand;               // BSJ candidate records (Circall_bsj.cpp:390)
};

// ---- debug surface (cg_opts::dump_intervals): the post-vote interval lists of every collector call, as the tier that answered it
// computed them — the row-A3 parity surface of the PRODUCTION tiers (tests compare them with the oracle's fwdSAInts / rcSAInts).
// head[item] = nF | nR << 12 | tier << 24 | found << 28 (0 = never written), off[item] = first entry in the pool.
struct IvDump { uint32_t* head; uint32_t* off; int4* pool; uint32_t* cursor; uint32_t cap; };
enum { TIER_PRE = 1, TIER_THR = 2, TIER_THR_VOTE = 3, TIER_WARP = 4, TIER_GEN = 5 };
__device__ __forceinline__ uint32_t iv_head(int nF, int nR, int tier, bool found) { return (uint32_t)nF | ((uint32_t)nR << 12) | ((uint32_t)tier << 24) | (found ? 1u << 28 : 0u); }
// one thread writes the lists it keeps in its scratch
template <class M>
__device__ __noinline__ void iv_dump_thread(IvDump d, uint64_t item, int tier, bool found, M m, int nF, int nR) {
    const uint32_t n = (uint32_t)(nF + nR), o = n ? atomicAdd(d.cursor, n) : 0u;
    d.head[item] = iv_head(nF, nR, tier, found); d.off[item] = o;
    if (o + n > d.cap) return;
    for (int sx = 0; sx < 2; ++sx)
        for (int i = 0; i < (sx ? nR : nF); ++i) {
            const uint32_t lq = m.ld(m.iv(sx, i) + 2);
            d.pool[o + (sx ? nF : 0) + i] = make_int4((int)m.ld(m.iv(sx, i)), (int)m.ld(m.iv(sx, i) + 1), (int)(lq & 0xffffu), (int)(lq >> 16));
        }
}
// one warp writes the lists of its collector state
__device__ __noinline__ void iv_dump_warp(IvDump d, uint64_t item, int tier, bool found, const Interval* fwd, int nF, const Interval* rc, int nR, int lane) {
    const uint32_t n = (uint32_t)(nF + nR);
    uint32_t o = 0;
    if (lane == 0) { o = n ? atomicAdd(d.cursor, n) : 0u; d.head[item] = iv_head(nF, nR, tier, found); d.off[item] = o; }
    o = __shfl_sync(0xffffffffu, o, 0);
    if (o + n > d.cap) return;
    for (int i = lane; i < (int)n; i += 32) {
        const Interval& I = i < nF ? fwd[i] : rc[i - nF];
        d.pool[o + i] = make_int4((int)I.begin, (int)I.end, (int)I.len, (int)I.qpos);
    }
}

// ---- warp-per-read kernels --------------------------------------------------------------------------
static const int WARPS_PER_BLOCK = 8;
typedef cgw::WStateT<cgw::FAST_FI, cgw::FAST_FK> FastState;
typedef cgw::WStateT<cgw::GEN_FI, cgw::GEN_FK> GenState;
// per-warp global scratch of the general kernels: collector state, two hit lists with their marks, merged list
static const size_t GEN_SCRATCH = ((sizeof(GenState) + 15) / 16) * 16 + 2 * (cgw::WSCR * 4 + cgw::WSCR) + 2 * cgw::WSCR * 4;
// shared-memory bytes of one warp: read view words (+ the fast kernels' collector state, + 64 transcript ids for bsj)
__host__ __device__ inline int warp_smem_bytes(const PackGeom& g, bool gen, bool bsj) {
    return (2 * (g.W2f + 1) + g.WMf + 1) * 8 + (gen ? 0 : (int)sizeof(FastState) + (bsj ? 64 * 4 : 0));
}
// stage the packed read of this warp into shared memory
__device__ __forceinline__ cgw::RV warp_view(const uint64_t* __restrict__ rec, const PackGeom& g, int L, uint64_t* dst, int lane) {
    const int W2 = g.W2f + 1;
    __syncwarp();
    for (int j = lane; j < g.WT; j += 32) {
        uint64_t v = __ldg(rec + j);
        dst[j < g.W2f ? j : 2 * W2 + (j - g.W2f)] = v;
    }
    if (lane == 0) { dst[g.W2f] = 0; dst[2 * W2 + g.WMf] = ~0ULL; }
    __syncwarp();
    // the reverse-complement strand, once per read: word j = 32 bases from oriented position 32 j (0 past the end)
    if (lane < W2) {
        ReadView fv; fv.w = dst; fv.L = L; fv.W2 = W2; fv.WM = 0; fv.anyN = false;
        dst[W2 + lane] = fv.chunk(1, 32 * lane);
    }
    // any N below L?  (mask bits past the read's end are set by construction)
    uint64_t m = 0;
    if (lane < g.WMf) {
        int lo = 64 * lane;
        m = dst[2 * W2 + lane];
        if (lo >= L) m = 0; else if (L - lo < 64) m &= (1ULL << (L - lo)) - 1;
    }
    cgw::RV rv; rv.w = dst; rv.L = L; rv.W2 = W2; rv.anyN = __any_sync(cgw::FULL, m != 0) ? 1 : 0;
    __syncwarp();
    return rv;
}
// hits of both strands merged by transcript (:567-579): number of distinct transcripts
__device__ __forceinline__ uint32_t warp_union_count(uint32_t tidF, bool aliveF, uint32_t tidR, bool aliveR) {
    unsigned fB = __ballot_sync(cgw::FULL, aliveF), rB = __ballot_sync(cgw::FULL, aliveR);
    uint32_t n = __popc(fB) + __popc(rB);
    if (fB && rB) {
        bool common = false;
#pragma unroll 1
        for (unsigned m = fB; m; m &= m - 1) common |= (__shfl_sync(cgw::FULL, tidF, __ffs(m) - 1) == tidR);
        n -= __popc(__ballot_sync(cgw::FULL, aliveR && common));
    }
    return n;
}
// where the general kernels keep their per-warp state
struct GenPtrs { GenState* st; uint32_t* scrA; uint8_t* actA; uint32_t* scrB; uint8_t* actB; uint32_t* merged; };
__device__ __forceinline__ GenPtrs gen_ptrs(unsigned char* base, uint64_t warp) {
    unsigned char* p = base + warp * GEN_SCRATCH;
    GenPtrs g;
    g.st = reinterpret_cast<GenState*>(p); p += ((sizeof(GenState) + 15) / 16) * 16;
    g.scrA = reinterpret_cast<uint32_t*>(p); p += cgw::WSCR * 4;
    g.scrB = reinterpret_cast<uint32_t*>(p); p += cgw::WSCR * 4;
    g.merged = reinterpret_cast<uint32_t*>(p); p += 2 * cgw::WSCR * 4;
    g.actA = p; p += cgw::WSCR;
    g.actB = p;
    return g;
}

// K2-K6: one warp per mate (persistent warps drawing their mates from a counter).
//   GEN = false: over all n mates, state in shared memory; ovf[r] = 1 hands the mate to the general kernel
//   GEN = true : over list[0..*n_list), worst-case state in global scratch
#ifndef CG_GEN_MINB
#define CG_GEN_MINB 3      // resident blocks per SM of the general kernels (latency-bound tails over a few thousand reads: more warps in flight)
#endif
#ifndef CG_WARP_MINB
#define CG_WARP_MINB 4     // resident blocks per SM asked of the fast warp kernels: 64 registers, 32 warps per SM (the kernels are latency-bound)
#endif
// next work item of a warp, drawn from a counter (zero at launch)
__device__ __forceinline__ uint64_t warp_draw(uint32_t* __restrict__ work, int lane) {
    uint32_t x = 0;
    if (lane == 0) x = atomicAdd(work, 1u);
    return (uint64_t)__shfl_sync(0xffffffffu, x, 0);
}
template <bool GEN, bool LIST>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, GEN ? CG_GEN_MINB : CG_WARP_MINB)
k_wt_warp(const uint64_t* __restrict__ pk, const uint16_t* __restrict__ lens, PackGeom g, uint64_t n_reads,
          const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list, int lceMode,
          uint8_t* __restrict__ flags, uint32_t* __restrict__ nhits, uint8_t* __restrict__ ovf, unsigned char* gscratch, IvDump dump, uint32_t* __restrict__ work) {
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint64_t* mine = smem + (size_t)wib * (warp_smem_bytes(g, GEN, false) / 8);
    const uint64_t nwarps = (uint64_t)gridDim.x * WARPS_PER_BLOCK, warp = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + wib;
    GenPtrs gp;
    if (GEN) gp = gen_ptrs(gscratch, warp);
    const uint64_t n = LIST ? (uint64_t)*n_list : n_reads;
    // the reads of a list differ a hundredfold in what they cost (repeat families): warps draw them from a counter instead of striding
    for (uint64_t x = warp_draw(work, lane); x < n; x = warp_draw(work, lane)) {
        const uint64_t r = LIST ? (uint64_t)list[x] : x;
        int L = lens[r];
        int pairLen = lens[r & ~1ULL];             // readLen is mate 1's length for both mates (Circall_wt.cpp:242)
        if ((CG_OPT_PF & 4) && !LIST && x + nwarps < n && lane * 4 < g.WT) cgw::prefetch_l2(pk + (x + nwarps) * (uint64_t)g.WT + lane * 4);   // this warp's next read
#ifdef CG_WARP_PROF
        const long long wp_read0 = clock64();
#endif
        cgw::RV rv = warp_view(pk + r * (uint64_t)g.WT, g, L, mine, lane);
        bool found, ok; int nF, nR;
        const Interval* fwd; const Interval* rc;
        if (GEN) {
            ok = cgw::w_collect<0, cgw::GEN_FI, cgw::GEN_FK>(rv, lane, lceMode, *gp.st, found, nF, nR);
            fwd = gp.st->fwd; rc = gp.st->rc;
        } else {
            FastState& st = *reinterpret_cast<FastState*>(mine + (2 * (g.W2f + 1) + g.WMf + 1));
            ok = cgw::w_collect<0, cgw::FAST_FI, cgw::FAST_FK>(rv, lane, lceMode, st, found, nF, nR);
            fwd = st.fwd; rc = st.rc;
        }
        uint8_t fl = 0; uint32_t nh = 0;
        if (ok && found) {
            if (nF + nR > 0) fl |= MATE_SEEDED;                                   // Circall_wt.cpp:311,486
            bool full = false;                                                    // :318-342,493-517
#pragma unroll 1
            for (int i = lane; i < nF; i += 32) full |= ((int)fwd[i].len == pairLen);
#pragma unroll 1
            for (int i = lane; i < nR; i += 32) full |= ((int)rc[i].len == pairLen);
            if (__any_sync(cgw::FULL, full)) fl |= MATE_FULL;
            WPROF_T0();
            if (GEN) {
                uint32_t a = cgw::w_hits_list<0>(lane, fwd, nF, gp.scrA, gp.actA);
                uint32_t b = cgw::w_hits_list<0>(lane, rc, nR, gp.scrB, gp.actB);
                nh = (a && b) ? cgw::w_merge(lane, gp.scrA, a, gp.scrB, b, gp.merged) : a + b;
            } else {
                cgw::Hits hF = cgw::w_hits_lanes<0>(lane, fwd, nF), hR = cgw::w_hits_lanes<0>(lane, rc, nR);
                ok = hF.ok && hR.ok;
                if (ok) nh = warp_union_count(hF.tid, hF.alive, hR.tid, hR.alive);
            }
            WPROF_ADD(cgw::WP_HITS, lane);
        }
#ifdef CG_WARP_PROF
        if (lane == 0) {
            const unsigned long long dt = (unsigned long long)(clock64() - wp_read0);
            atomicAdd(&cgw::g_wprof[7], dt); atomicAdd(&cgw::g_wprof[8 + (GEN ? 1 : 0)], 1ULL); atomicAdd(&cgw::g_wprof[16 + (GEN ? 24 : 0) + (63 - __clzll(dt | 1ULL))], 1ULL);
        }
#endif
        if (lane == 0) { flags[r] = ok ? fl : MATE_ERR; nhits[r] = ok ? nh : 0; if (!GEN) ovf[r] = ok ? 0 : 1; }
        if (dump.head && ok) iv_dump_warp(dump, r, GEN ? TIER_GEN : TIER_WARP, found, fwd, found ? nF : 0, rc, found ? nR : 0, lane);
    }
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* sh) {
    v = __reduce_add_sync(0xffffffffu, v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    uint32_t t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0;
    if (w == 0) t = __reduce_add_sync(0xffffffffu, t);
    return t;   // valid in warp 0
}

// one thread per pair: export flags, per-mate detail, scalar counters
__global__ void __launch_bounds__(256)
k_wt_commit(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ nhits, uint64_t n_pairs,
            uint8_t* __restrict__ expflag, uint8_t* __restrict__ mflags_out, uint16_t* __restrict__ mnhits_out,
            unsigned long long* ctr, BatchHdr* hdr) {
    __shared__ uint32_t sh[8];
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t upper = 0, mapped = 0, hits = 0, obs = 0;
    if (i < n_pairs) {
        uint8_t f0 = flags[2 * i], f1 = flags[2 * i + 1];
        uint32_t h0 = nhits[2 * i], h1 = nhits[2 * i + 1];
        if ((f0 | f1) & MATE_ERR) atomicOr(&hdr->status, ST_CAPACITY);
        upper = (h0 > 0) + (h1 > 0);                                               // Circall_wt.cpp:369,544
        uint32_t a = h0 > (uint32_t)MAX_READ_OCCS ? 0 : h0, b = h1 > (uint32_t)MAX_READ_OCCS ? 0 : h1;   // :372,547
        mapped = (a > 0) + ((a > 0) || (b > 0));                                   // mappedFrag is sticky across mates :467,642
        hits = a + b;                                                              // :468,643
        obs = 2;                                                                   // :470,645
        expflag[2 * i] = ((f0 & MATE_SEEDED) && !(f0 & MATE_FULL)) ? 1 : 0;        // :350
        expflag[2 * i + 1] = ((f1 & MATE_SEEDED) && !(f1 & MATE_FULL)) ? 1 : 0;    // :525
        if (mflags_out) {
            mflags_out[2 * i] = f0 & 3; mflags_out[2 * i + 1] = f1 & 3;
            mnhits_out[2 * i] = (uint16_t)a; mnhits_out[2 * i + 1] = (uint16_t)b;
        }
    }
    uint32_t s;
    s = block_sum(obs, sh);    if (threadIdx.x == 0 && s) atomicAdd(&ctr[C_WT_OBS], (unsigned long long)s);
    s = block_sum(mapped, sh); if (threadIdx.x == 0 && s) atomicAdd(&ctr[C_WT_MAP], (unsigned long long)s);
    s = block_sum(hits, sh);   if (threadIdx.x == 0 && s) atomicAdd(&ctr[C_WT_HITS], (unsigned long long)s);
    s = block_sum(upper, sh);  if (threadIdx.x == 0 && s) atomicAdd(&ctr[C_WT_UPPER], (unsigned long long)s);
}

// ---- ordered compaction of a byte-flag array: indices of the set flags, in order -----------------
static const int CP_THREADS = 256, CP_ITEMS = 16, CP_TILE = CP_THREADS * CP_ITEMS;     // (4 items: 4 x the tiles, and the single-block scan of 16 class counts per tile took 0.2 ms per stage)

__global__ void __launch_bounds__(CP_THREADS)
k_flag_count(const uint8_t* __restrict__ f, uint64_t n, uint32_t* __restrict__ bc) {
    __shared__ uint32_t sh[8];
    uint64_t base = blockIdx.x * (uint64_t)CP_TILE + threadIdx.x * CP_ITEMS;
    uint32_t c = 0;
#pragma unroll
    for (int x = 0; x < CP_ITEMS; ++x) if (base + x < n) c += f[base + x] ? 1 : 0;
    uint32_t s = block_sum(c, sh);
    if (threadIdx.x == 0) bc[blockIdx.x] = s;
}
// exclusive scan of nb block counts by one block; total -> *total
__global__ void __launch_bounds__(1024)
k_flag_scan(uint32_t* bc, uint32_t nb, uint32_t* total) {
    // one block, SCAN_ITEMS consecutive counts per thread and trip (the class sort scans 16 counts per tile: 250 k entries for 8 M pairs —
    // one entry per thread and trip took 0.22 ms, a launch-sized hole in the middle of every stage)
    const int SCAN_ITEMS = 8;
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024 * SCAN_ITEMS) {
        const uint32_t i0 = base + threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) { v[j] = i0 + j < nb ? bc[i0 + j] : 0; sum += v[j]; }
        uint32_t x = sum;
        int l = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (l >= d) x += y; }
        if (l == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t s = wsum[l], z = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, d); if (l >= d) z += y; }
            wsum[l] = z - s;
        }
        __syncthreads();
        uint32_t excl = carry + wsum[w] + x - sum;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) { if (i0 + j < nb) bc[i0 + j] = excl; excl += v[j]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(CP_THREADS)
k_flag_write(const uint8_t* __restrict__ f, uint64_t n, const uint32_t* __restrict__ bc, uint32_t* __restrict__ out) {
    __shared__ uint32_t wsum[8];
    uint64_t base = blockIdx.x * (uint64_t)CP_TILE + threadIdx.x * CP_ITEMS;
    uint8_t v[CP_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int x = 0; x < CP_ITEMS; ++x) { v[x] = (base + x < n) ? f[base + x] : 0; c += v[x] ? 1 : 0; }
    int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (l >= d) x += y; }
    if (l == 31) wsum[w] = x;
    __syncthreads();
    uint32_t pre = 0;
    for (int y = 0; y < w; ++y) pre += wsum[y];
    uint32_t o = bc[blockIdx.x] + pre + x - c;
#pragma unroll
    for (int y = 0; y < CP_ITEMS; ++y) if (v[y]) out[o++] = (uint32_t)(base + y);
}

// ---- counting sort of the non-zero entries of a class array (values 1..16): indices grouped by class ------------
// The lockstep thread tier runs one read per lane; a warp is efficient when its 32 reads take the same path, and the path
// is mostly decided by which of the read's four end k-mers (first, last, their reverse complements) are in the table.
// bc is class-major (bc[c * nb + block]) so that ONE exclusive scan (k_flag_scan) yields every block's write offsets.
static const int CLS_N = 16;
static const int CLS_VOTE = 12;       // first bucket of the two-strand classes
// bucket of class c (1..16): the four classes whose first k-mer AND its reverse complement are in the table (both strands
// seeded: the read needs the kmerScores vote) come last, so that the two-strand instantiation of the lockstep kernels
// runs over the tail [bc[CLS_VOTE * nb], total) of the sorted list and the lean one over the head
__device__ __forceinline__ int cls_bucket(uint8_t c) {
    const int v = (c - 1) & (CLS_N - 1), lo = v & 3, hi = v >> 2;
    return lo == 3 ? CLS_VOTE + hi : hi * 3 + lo;
}
__global__ void __launch_bounds__(CP_THREADS)
k_cls_count(const uint8_t* __restrict__ f, uint64_t n, uint32_t* __restrict__ bc, uint32_t nb) {
    __shared__ uint32_t sh[CLS_N];
    if (threadIdx.x < CLS_N) sh[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = blockIdx.x * (uint64_t)CP_TILE + threadIdx.x * CP_ITEMS;
#pragma unroll
    for (int x = 0; x < CP_ITEMS; ++x) if (base + x < n) { uint8_t c = f[base + x]; if (c) atomicAdd(&sh[cls_bucket(c)], 1u); }
    __syncthreads();
    if (threadIdx.x < CLS_N) bc[threadIdx.x * nb + blockIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(CP_THREADS)
k_cls_write(const uint8_t* __restrict__ f, uint64_t n, const uint32_t* __restrict__ bc, uint32_t nb, uint32_t* __restrict__ out) {
    __shared__ uint32_t cur[CLS_N];
    if (threadIdx.x < CLS_N) cur[threadIdx.x] = bc[threadIdx.x * nb + blockIdx.x];
    __syncthreads();
    uint64_t base = blockIdx.x * (uint64_t)CP_TILE + threadIdx.x * CP_ITEMS;
#pragma unroll
    for (int x = 0; x < CP_ITEMS; ++x) if (base + x < n) { uint8_t c = f[base + x]; if (c) out[atomicAdd(&cur[cls_bucket(c)], 1u)] = (uint32_t)(base + x); }
}

// Circall_bsj has no front kernel in front of the lockstep tier: this one only classifies the exported records
// (cls[rec] = 1 + f0 + 2 f1 + 4 f2 + 8 f3 for the k-mers read[0,k), its reverse complement, read[L-k,L), its reverse
// complement; 1 for a read with an N or shorter than k + 1)
__global__ void __launch_bounds__(256)
k_bsj_cls(const uint64_t* __restrict__ pk, const uint16_t* __restrict__ lens, PackGeom g, const uint32_t* __restrict__ exp_mate, uint64_t n_fixed,
          const uint32_t* __restrict__ n_dev, uint8_t* __restrict__ cls) {
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_fixed;
    const uint64_t rec = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (rec >= n) return;
    const IndexView& ix = cgw::c_ix[1];
    const uint64_t r = exp_mate ? (uint64_t)exp_mate[rec] : 2 * rec;
    const uint64_t* w = pk + r * (uint64_t)g.WT;
    const int L = lens[r], k = ix.k;
    uint8_t c = 1;
    bool anyN = false;
    for (int j = 0; j < g.WMf; ++j) {
        const int lo = 64 * j;
        if (lo >= L) break;
        uint64_t mk = __ldg(w + g.W2f + j);
        if (L - lo < 64) mk &= (1ULL << (L - lo)) - 1;
        anyN |= mk != 0;
    }
    if (!anyN && L > k) {
        const int j = (L - k) >> 5, sh = ((L - k) & 31) * 2;
        const uint64_t lo = __ldg(w + j), hi = sh ? __ldg(w + j + 1) : 0ULL;
        const uint64_t mer0 = __ldg(w) & ix.km, merL = (sh ? (lo >> sh) | (hi << (64 - sh)) : lo) & ix.km;
        const uint64_t rc0 = revcomp(mer0, k), rcL = revcomp(merL, k);
        const uint64_t h0 = home_slot(ix, mer0), h1 = home_slot(ix, rc0), h2 = home_slot(ix, merL), h3 = home_slot(ix, rcL);
        const u32x4 s0 = CG_LDG(&ix.hslots[h0]), s1 = CG_LDG(&ix.hslots[h1]), s2 = CG_LDG(&ix.hslots[h2]), s3 = CG_LDG(&ix.hslots[h3]);
        uint32_t b, e;
        c = (uint8_t)(1 + (resolve(ix, mer0, h0, s0, b, e) ? 1 : 0) + (resolve(ix, rc0, h1, s1, b, e) ? 2 : 0)
                        + (resolve(ix, merL, h2, s2, b, e) ? 4 : 0) + (resolve(ix, rcL, h3, s3, b, e) ? 8 : 0));
    }
    cls[rec] = c;
}

// ---- Circall_bsj ----------------------------------------------------------------------------------
struct JSink {
    cg_junction* buf; unsigned long long* cursor; uint64_t cap; uint32_t rec;
    __device__ __forceinline__ void line(uint32_t dir, uint32_t q, uint32_t len, int hitPos, uint32_t tid) {
        unsigned long long s = atomicAdd(cursor, 1ULL);
        if (s < cap) { cg_junction j; j.rec = rec; j.dir = dir; j.query_pos = q; j.len = len; j.hit_pos = hitPos; j.tid = tid; buf[s] = j; }
    }
};

struct BsjOut {
    cg_junction* junc; uint64_t junc_cap;
    uint4* eq_ent; uint64_t eq_cap;            // {rec, first tid slot, number of tids, 0}
    uint32_t* eq_tids; uint64_t eq_tids_cap;
    uint32_t* rec_nhits; uint8_t* rec_split;
    unsigned long long* per_bsj;               // dense single-transcript class counters
    unsigned long long* ctr;
    BatchHdr* hdr;
};

// K2-K5, K7, K8: one warp per exported record (persistent warps).  exp_mate == nullptr: record rec is mate 1 of
// pair rec (stage 2).
//   GEN = false: over all records (count read from the device), state in shared memory; a record that exceeds a
//                fast capacity is flagged in ovf[] and emits nothing here
//   GEN = true : over list[0..*n_list), worst-case state in global scratch
// do_counts bit0: accumulate the scalar and per-BSJ counters; bit1: record the multi-transcript classes.
template <bool GEN, bool LIST>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, GEN ? CG_GEN_MINB : CG_WARP_MINB)
k_bsj_warp(const uint64_t* __restrict__ pk, const uint16_t* __restrict__ lens, PackGeom g,
           const uint32_t* __restrict__ exp_mate, uint64_t n_fixed, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list,
           int lceMode, BsjOut o, int do_counts, uint8_t* __restrict__ ovf, unsigned char* gscratch, IvDump dump, uint32_t* __restrict__ work) {
    extern __shared__ uint64_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint64_t* mine = smem + (size_t)wib * (warp_smem_bytes(g, GEN, true) / 8);
    const uint64_t nwarps = (uint64_t)gridDim.x * WARPS_PER_BLOCK, warp = (uint64_t)blockIdx.x * WARPS_PER_BLOCK + wib;
    const IndexView& ix = cgw::c_ix[1];
    GenPtrs gp;
    if (GEN) gp = gen_ptrs(gscratch, warp);
    const uint64_t n = LIST ? (uint64_t)*n_list : (exp_mate ? (uint64_t)o.hdr->n_exp : n_fixed);
    uint32_t obs = 0, mapped = 0, hits = 0, upper = 0;          // per-warp totals (kept by every lane alike)
    for (uint64_t x = warp_draw(work, lane); x < n; x = warp_draw(work, lane)) {
        const uint64_t rec = LIST ? (uint64_t)list[x] : x;
        const uint64_t r = exp_mate ? (uint64_t)exp_mate[rec] : 2 * rec;
        int L = lens[r];
        cgw::RV rv = warp_view(pk + r * (uint64_t)g.WT, g, L, mine, lane);
        bool found, ok; int nF, nR;
        const Interval* fwd; const Interval* rc;
        uint32_t* tl;                                            // class key (ordered transcript list)
        if (GEN) {
            ok = cgw::w_collect<1, cgw::GEN_FI, cgw::GEN_FK>(rv, lane, lceMode, *gp.st, found, nF, nR);
            fwd = gp.st->fwd; rc = gp.st->rc; tl = gp.merged;
        } else {
            FastState& st = *reinterpret_cast<FastState*>(mine + (2 * (g.W2f + 1) + g.WMf + 1));
            ok = cgw::w_collect<1, cgw::FAST_FI, cgw::FAST_FK>(rv, lane, lceMode, st, found, nF, nR);
            fwd = st.fwd; rc = st.rc; tl = reinterpret_cast<uint32_t*>(&st + 1);
        }
        uint32_t nh = 0;
        if (ok && found) {
            if (GEN) {
                uint32_t a = cgw::w_hits_list<1>(lane, fwd, nF, gp.scrA, gp.actA);
                uint32_t b = cgw::w_hits_list<1>(lane, rc, nR, gp.scrB, gp.actB);
                if (a && b) nh = cgw::w_merge(lane, gp.scrA, a, gp.scrB, b, gp.merged);
                else { nh = a + b; tl = a ? gp.scrA : gp.scrB; }
            } else {
                cgw::Hits hF = cgw::w_hits_lanes<1>(lane, fwd, nF), hR = cgw::w_hits_lanes<1>(lane, rc, nR);
                ok = hF.ok && hR.ok;
                if (ok) {
                    nh = warp_union_count(hF.tid, hF.alive, hR.tid, hR.alive);
                    if (nh > 0 && nh <= 64) {
                        // gather the hit transcripts, sort and de-duplicate them on lane 0 (a handful of entries)
                        unsigned fB = __ballot_sync(cgw::FULL, hF.alive), rB = __ballot_sync(cgw::FULL, hR.alive);
                        __syncwarp();
                        if (hF.alive) tl[__popc(fB & ((1u << lane) - 1))] = hF.tid;
                        if (hR.alive) tl[__popc(fB) + __popc(rB & ((1u << lane) - 1))] = hR.tid;
                        __syncwarp();
                        if (lane == 0) {
                            int nn = __popc(fB) + __popc(rB);
                            for (int i = 1; i < nn; ++i) { uint32_t v = tl[i]; int j = i - 1; while (j >= 0 && tl[j] > v) { tl[j + 1] = tl[j]; --j; } tl[j + 1] = v; }
                            int m = 0;
                            for (int i = 0; i < nn; ++i) if (m == 0 || tl[m - 1] != tl[i]) tl[m++] = tl[i];
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (!ok) {
            if (GEN) {          // the general collector ran out of its worst-case capacities: reported, never dropped (as k_wt_warp / k_wt_commit do)
                if (lane == 0) { atomicOr(&o.hdr->status, ST_CAPACITY); o.rec_nhits[rec] = 0; o.rec_split[rec] = 0; }
            } else if (lane == 0) ovf[rec] = 1;
            continue;
        }
        if (dump.head) iv_dump_warp(dump, rec, GEN ? TIER_GEN : TIER_WARP, found, fwd, found ? nF : 0, rc, found ? nR : 0, lane);
        bool split = false;
        if (nh > 0) {                                                             // Circall_bsj.cpp:319
#pragma unroll 1
            for (int dir = 0; dir < 2; ++dir) {
                const Interval* ints = dir == 0 ? fwd : rc;
                const int ni = dir == 0 ? nF : nR;
#pragma unroll 1
                for (int xi = 0; xi < ni; ++xi) {
                    const Interval it = ints[xi];
#pragma unroll 1
                    for (uint32_t c0 = it.begin; c0 < it.end; c0 += 32) {         // :326 / :355, 32 positions per round
                        bool pred = false; uint32_t tid = 0; int hitPos = 0;
                        if (c0 + lane < it.end) {
                            uint32_t p = CG_LDG(&ix.sa[c0 + lane]), local;
                            if (ix.satid) { tid = CG_LDG(&ix.satid[c0 + lane]); local = p - CG_LDG(&ix.offsets[tid]); }
                            else tid_of(ix, p, tid, local);
                            uint32_t refLen = CG_LDG(&ix.offsets[tid + 1]) - CG_LDG(&ix.offsets[tid]) - 1;
                            hitPos = (int)local;
                            int junctPos = (int)(refLen / 2);                     // :333
                            pred = hitPos < junctPos - 5 && hitPos + (int)it.len > junctPos + 5;   // :336 / :364
                        }
                        unsigned pB = __ballot_sync(cgw::FULL, pred);
                        if (pB) {
                            split = true;
                            unsigned long long base = 0;
                            if (lane == 0) base = atomicAdd(&o.hdr->n_junc, (unsigned long long)__popc(pB));
                            base = cgw::shfl64(base, 0);
                            if (pred) {
                                unsigned long long slot = base + __popc(pB & ((1u << lane) - 1));
                                if (slot < o.junc_cap) {
                                    cg_junction j; j.rec = (uint32_t)rec; j.dir = (uint32_t)dir; j.query_pos = it.qpos; j.len = it.len;
                                    j.hit_pos = hitPos; j.tid = tid;
                                    o.junc[slot] = j;
                                }
                            }
                        }
                    }
                }
            }
        }
        // commit (Circall_bsj.cpp:383-491)
        if (lane == 0) { o.rec_nhits[rec] = nh; o.rec_split[rec] = split ? 1 : 0; }
        upper += nh > 0;                                                          // :383
        const uint32_t c = nh > (uint32_t)MAX_READ_OCCS ? 0 : nh;                 // :385
        const bool cand = c > 0 && split;                                         // :390
        obs += 1; mapped += cand; hits += c;                                      // :488-491
        if (cand) {
            // tl[0..c) = jointHits' transcripts in ascending order (the fast kernel fills it for nh <= 64; a fast
            // record with more hits than that is impossible: <= 32 per strand)
            if (c == 1) {
                if ((do_counts & 1) && lane == 0) atomicAdd(&o.per_bsj[tl[0]], 1ULL); // single-transcript class (EquivalenceClassBuilder.hpp:112-130)
            } else if (do_counts & 2) {
                unsigned long long e = 0, sidx = 0;
                if (lane == 0) { e = atomicAdd(&o.hdr->n_eq, 1ULL); sidx = atomicAdd(&o.hdr->n_eq_tids, (unsigned long long)c); }
                e = cgw::shfl64(e, 0); sidx = cgw::shfl64(sidx, 0);
                if (e < o.eq_cap && sidx + c <= o.eq_tids_cap) {
                    if (lane == 0) o.eq_ent[e] = make_uint4((uint32_t)rec, (uint32_t)sidx, c, 0u);
                    for (uint32_t t = lane; t < c; t += 32) o.eq_tids[sidx + t] = tl[t];
                }
            }
        }
    }
    if ((do_counts & 1) && lane == 0) {
        if (obs) atomicAdd(&o.ctr[C_BSJ_OBS], (unsigned long long)obs);
        if (mapped) atomicAdd(&o.ctr[C_BSJ_MAP], (unsigned long long)mapped);
        if (hits) atomicAdd(&o.ctr[C_BSJ_HITS], (unsigned long long)hits);
        if (upper) atomicAdd(&o.ctr[C_BSJ_UPPER], (unsigned long long)upper);
    }
}

// ---- front kernel of Circall_wt: one thread per mate, clean full-length matches only (cg_pre.cuh) -----------------
#ifndef CG_PRE_THREADS
#define CG_PRE_THREADS 128
#endif
static const int PRE_THREADS = CG_PRE_THREADS;
template <bool DUMP>        // DUMP: the debug surface of cg_opts::dump_intervals, its own instantiation so that the production kernel carries none of it
__global__ void __launch_bounds__(PRE_THREADS)
k_wt_pre(const uint64_t* __restrict__ pk, const uint16_t* __restrict__ lens, PackGeom g, uint64_t n_reads,
         uint8_t* __restrict__ flags, uint32_t* __restrict__ nhits, uint8_t* __restrict__ ovf, IvDump dump) {
    extern __shared__ uint32_t pre_smem[];
    const uint64_t r = (uint64_t)blockIdx.x * PRE_THREADS + threadIdx.x;
    if (r >= n_reads) return;
    const uint64_t* rec = pk + r * (uint64_t)g.WT;
    const int L = lens[r], pairLen = lens[r & ~1ULL];          // readLen is mate 1's length for both mates (Circall_wt.cpp:242)
    cgt::TMem m; m.base = pre_smem + threadIdx.x; m.stride = PRE_THREADS; m.W2 = g.W2f + 1;
    bool anyN = false;
    for (int j = 0; j < g.WMf; ++j) {
        const int lo = 64 * j;
        if (lo >= L) break;
        uint64_t mk = __ldg(rec + g.W2f + j);
        if (L - lo < 64) mk &= (1ULL << (L - lo)) - 1;
        anyN |= mk != 0;
    }
    bool done = false;
    uint8_t fl = 0, cls = 1; uint32_t nh = 0;
    if (!anyN) {
        for (int j = 0; j < g.W2f; ++j) { const uint64_t v = __ldg(rec + j); m.st(2 * j, (uint32_t)v); m.st(2 * j + 1, (uint32_t)(v >> 32)); }
        m.st(2 * g.W2f, 0); m.st(2 * g.W2f + 1, 0);
        cgp::PreIv piv;
        done = cgp::perfect_mate(cgw::c_ix[0], m, L, pairLen, fl, nh, cls, DUMP ? &piv : (cgp::PreIv*)nullptr);
        if (DUMP && done) {
            // the list this path stands for (cg_pre.cuh): the full-length interval, then the k-mer interval of the forced last position —
            // twice on the forward strand (Phase B re-runs it after setting lastSearch, :300-321), once on the other (:412)
            const int k = cgw::c_ix[0].k;
            const bool lastOk = piv.el > piv.bl && piv.el - piv.bl < 1000u;
            const int n = 1 + (lastOk ? (piv.s == 0 ? 2 : 1) : 0);
            const uint32_t o = atomicAdd(dump.cursor, (uint32_t)n);
            dump.head[r] = iv_head(piv.s == 0 ? n : 0, piv.s == 0 ? 0 : n, TIER_PRE, true); dump.off[r] = o;
            if (o + n <= dump.cap) {
                dump.pool[o] = make_int4((int)piv.lb, (int)piv.ub, L, 0);
                for (int i = 1; i < n; ++i) dump.pool[o + i] = make_int4((int)piv.bl, (int)piv.el, k, L - k);
            }
        }
    }
    if (done) { flags[r] = fl; nhits[r] = nh; }
    ovf[r] = done ? 0 : cls;          // declined: the look-up class (1..16), see k_cls_*
}

// ---- lockstep thread tier (cg_thr.cuh): one thread per read, single-strand reads ------------------------------------
#ifndef CG_THR_THREADS
#define CG_THR_THREADS 64        // measured in round 1, 32 / 64 / 128 / 256: bsj lockstep kernel 7.70 / 7.45 / 7.57 / 7.70 ms, the two-strand kernels 0.41 / 0.41 / 0.56 / 0.64; round 2 after the split: 64 -> 23.93 ms per step, 128 -> 25.12
#endif
static const int THR_THREADS = CG_THR_THREADS;
#ifndef CG_THR_VOTE_MINB
#define CG_THR_VOTE_MINB 5      // resident blocks per SM (x 128 threads) of the two-strand collector (measured, step of 8 M pairs: 3 -> 21.74 ms, 4 -> 21.48, 5 -> 21.41)
#endif
// stage the packed read of this thread into its scratch; false when the read holds an N (or is empty)
template <class M>
__device__ __forceinline__ bool thr_stage(const uint64_t* __restrict__ rec, const PackGeom& g, int L, const M& m) {
    bool anyN = L < 1;
    for (int j = 0; j < g.WMf; ++j) {
        const int lo = 64 * j;
        if (lo >= L) break;
        uint64_t mk = __ldg(rec + g.W2f + j);
        if (L - lo < 64) mk &= (1ULL << (L - lo)) - 1;
        anyN |= mk != 0;
    }
    if (anyN) return false;
    for (int j = 0; j < g.W2f; ++j) { const uint64_t v = __ldg(rec + j); m.st(2 * j, (uint32_t)v); m.st(2 * j + 1, (uint32_t)(v >> 32)); }
    m.st(2 * g.W2f, 0); m.st(2 * g.W2f + 1, 0);
    cgt::make_rc(m, L);
    return true;
}

// The lockstep tier runs as two kernels per list: k_thr_col (the collector: q_begin, the q_step loop, q_vote — its intervals go to a
// slot of the hand-over buffer) and k_wt_fin / k_bsj_fin (hit enumeration, full-length test / span filter, commit).  As one kernel
// (72 KB of code for Circall_bsj) the warps of an SM, each at its own place in the collector or in the hit enumeration, ran out of the
// 32 KB instruction cache: 56 % of the single Circall_bsj kernel's stall samples were instruction-fetch stalls (ncu r2s).  Split, each kernel's loop fits,
// and the second one starts with all lanes of a warp at the same instruction.
// Hand-over, slot x = position in the work list (cap slots): res[x] = status (int8) | nF << 8 | nR << 16, the intervals of the nF forward
// then the nR reverse-complement entries as three words each, word j of slot x at iv[j * cap + x] (coalesced across a warp).
static const int THR_IVW = 3 * cgq::Q_FI;          // words per slot; a two-strand read that keeps more than Q_FI intervals after the vote is declined

// Over the work list[..] (null: all items): Circall_wt mates (BSJ = false: item = mate), or Circall_bsj records (BSJ = true: item = record,
// read = exp_mate[item], or mate 1 of pair item when exp_mate is null).  The list is grouped by look-up class: the lean instantiation
// (VOTE = false) takes its head [0, *seg), the two-strand one its tail [*seg, *n_list); seg null = one instantiation over everything.
template <bool VOTE, bool BSJ>
__global__ void __launch_bounds__(THR_THREADS, (VOTE ? CG_THR_VOTE_MINB : CG_THR_MINB) * (128 / THR_THREADS))
k_thr_col(const uint64_t* __restrict__ pk, const uint16_t* __restrict__ lens, PackGeom g, const uint32_t* __restrict__ exp_mate, uint64_t n_items,
          const uint32_t* __restrict__ n_exp, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list, const uint32_t* __restrict__ seg,
          int lceMode, const uint8_t* __restrict__ cls, uint32_t* __restrict__ res, uint32_t* __restrict__ iv, uint64_t cap) {
    extern __shared__ uint32_t thr_smem[];
    const uint64_t total = list ? (uint64_t)*n_list : (n_exp ? (uint64_t)*n_exp : n_items);
    const uint64_t lo = (VOTE && seg) ? (uint64_t)*seg : 0, hi = (!VOTE && seg) ? (uint64_t)*seg : total;
    const uint64_t x = lo + (uint64_t)blockIdx.x * THR_THREADS + threadIdx.x;
    if (x >= hi) return;
    const uint64_t item = list ? (uint64_t)list[x] : x;
    const uint64_t r = BSJ ? (exp_mate ? (uint64_t)exp_mate[item] : 2 * item) : item;
    const int L = lens[r];
    const IndexView& ix = cgw::c_ix[BSJ ? 1 : 0];
    cgq::QMemT<VOTE> m; m.base = thr_smem + threadIdx.x; m.stride = THR_THREADS; m.W2 = g.W2f + 1;
    if (!VOTE) { m.giv = iv + x; m.gcap = cap; }
    cgq::QState st; st.nF = st.nR = 0;
    int status = thr_stage(pk + r * (uint64_t)g.WT, g, L, m) ? cgq::q_begin<VOTE>(ix, m, L, cls ? (int)cls[item] - 1 : -1, st) : (int)cgq::QD_N;
    while (status == cgq::Q_CONT) status = cgq::q_step<VOTE>(ix, m, lceMode, st);
    if (status == cgq::Q_DONE) status = cgq::q_vote<VOTE>(ix, m, st);
    int nF = status == cgq::Q_FOUND ? st.nF : 0, nR = status == cgq::Q_FOUND ? st.nR : 0;
    if (VOTE && nF + nR > cgq::Q_FI) { status = cgq::QD_NINT; nF = nR = 0; }
    res[x] = (uint32_t)(status & 0xff) | ((uint32_t)nF << 8) | ((uint32_t)nR << 16);
    if (VOTE) for (int i = 0; i < nF + nR; ++i) {
        const int at = i < nF ? m.iv(0, i) : m.iv(1, i - nF);
        iv[(uint64_t)(3 * i) * cap + x] = m.ld(at); iv[(uint64_t)(3 * i + 1) * cap + x] = m.ld(at + 1); iv[(uint64_t)(3 * i + 2) * cap + x] = m.ld(at + 2);
    }
}
// the slot of work item x back into a thread's scratch: status, nF, nR
template <bool VOTE>
__device__ __forceinline__ int thr_take(const uint32_t* __restrict__ res, const uint32_t* __restrict__ iv, uint64_t cap, uint64_t x, const cgq::QMemT<VOTE>& m, int& nF, int& nR) {
    const uint32_t w = res[x];
    nF = (int)((w >> 8) & 0xffu); nR = (int)((w >> 16) & 0xffu);
    for (int i = 0; i < nF + nR; ++i) {
        const int at = i < nF ? m.iv(0, i) : m.iv(1, i - nF);
        m.st(at, iv[(uint64_t)(3 * i) * cap + x]); m.st(at + 1, iv[(uint64_t)(3 * i + 1) * cap + x]); m.st(at + 2, iv[(uint64_t)(3 * i + 2) * cap + x]);
    }
    return (int)(int8_t)(w & 0xffu);
}

// Circall_wt after the collector: ovf[r] = 1 hands mate r to the warp kernels
template <bool VOTE, bool DUMP>
__global__ void __launch_bounds__(THR_THREADS, 12)
k_wt_fin(const uint16_t* __restrict__ lens, uint64_t n_reads, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list, const uint32_t* __restrict__ seg,
         const uint32_t* __restrict__ res, const uint32_t* __restrict__ iv, uint64_t cap, uint8_t* __restrict__ flags, uint32_t* __restrict__ nhits,
         uint8_t* __restrict__ ovf, uint8_t* __restrict__ retry, IvDump dump) {
    extern __shared__ uint32_t thr_smem[];
    const uint64_t total = list ? (uint64_t)*n_list : n_reads;
    const uint64_t lo = (VOTE && seg) ? (uint64_t)*seg : 0, hi = (!VOTE && seg) ? (uint64_t)*seg : total;
    const uint64_t x = lo + (uint64_t)blockIdx.x * THR_THREADS + threadIdx.x;
    if (x >= hi) return;
    const uint64_t r = list ? (uint64_t)list[x] : x;
    const int pairLen = lens[r & ~1ULL];          // readLen is mate 1's length for both mates (Circall_wt.cpp:242)
    cgq::QMemT<VOTE> m; m.base = thr_smem + threadIdx.x; m.stride = THR_THREADS; m.W2 = 0;
    int nF, nR;
    const int status = thr_take<VOTE>(res, iv, cap, x, m, nF, nR);
    uint8_t fl = 0; uint32_t nh = 0; int why = status;
    const bool done = status >= 0 && cgq::q_wt_finish<VOTE>(cgw::c_ix[0], m, status, pairLen, nF, nR, fl, nh, why);
    if (done) { flags[r] = fl; nhits[r] = nh; }
    else if (!VOTE && retry && (status == cgq::QD_BOTH_A || status == cgq::QD_OTHER)) retry[r] = 1;      // the other strand scores too: the two-strand instantiation's read
    else ovf[r] = 1;
    if (DUMP && done) iv_dump_thread(dump, r, VOTE ? TIER_THR_VOTE : TIER_THR, status == cgq::Q_FOUND, m, nF, nR);
}

// Circall_bsj after the collector: rows to the junction buffer, per-BSJ counts, multi-transcript classes, per-record hits / split;
// ovf[rec] = 1 hands the record to the warp kernels (nothing was emitted for it)
template <bool VOTE, bool DUMP>
__global__ void __launch_bounds__(THR_THREADS, 12)
k_bsj_fin(const uint32_t* __restrict__ exp_mate, uint64_t n_fixed, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list, const uint32_t* __restrict__ seg,
          const uint32_t* __restrict__ res, const uint32_t* __restrict__ iv, uint64_t cap, BsjOut o, int do_counts, uint8_t* __restrict__ ovf,
          uint8_t* __restrict__ retry, IvDump dump) {
    extern __shared__ uint32_t thr_smem[];
    const uint64_t total = list ? (uint64_t)*n_list : (exp_mate ? (uint64_t)o.hdr->n_exp : n_fixed);
    const uint64_t lo = (VOTE && seg) ? (uint64_t)*seg : 0, hi = (!VOTE && seg) ? (uint64_t)*seg : total;
    const uint64_t x = lo + (uint64_t)blockIdx.x * THR_THREADS + threadIdx.x;
    const bool active = x < hi;
    const uint64_t rec = active && list ? (uint64_t)list[x] : x;
    uint32_t obs = 0, mapped = 0, hits = 0, upper = 0;
    if (active) {
        cgq::QMemT<VOTE> m; m.base = thr_smem + threadIdx.x; m.stride = THR_THREADS; m.W2 = 0;
        int nF, nR;
        const int status = thr_take<VOTE>(res, iv, cap, x, m, nF, nR);
        JSink sink; sink.buf = o.junc; sink.cursor = &o.hdr->n_junc; sink.cap = o.junc_cap; sink.rec = (uint32_t)rec;
        uint32_t nh = 0; bool split = false; int why = status, at = 0;
        const bool done = status >= 0 && cgq::q_bsj_finish<VOTE>(cgw::c_ix[1], m, status, nF, nR, sink, nh, split, at, why);
        if (DUMP && done) iv_dump_thread(dump, rec, VOTE ? TIER_THR_VOTE : TIER_THR, status == cgq::Q_FOUND, m, nF, nR);
        if (!done) {
            if (!VOTE && retry && (status == cgq::QD_BOTH_A || status == cgq::QD_OTHER)) retry[rec] = 1;
            else ovf[rec] = 1;
        } else {
            // commit (Circall_bsj.cpp:383-491)
            o.rec_nhits[rec] = nh; o.rec_split[rec] = split ? 1 : 0;
            upper = nh > 0;                                                       // :383
            const uint32_t c = nh > (uint32_t)MAX_READ_OCCS ? 0 : nh;             // :385
            const bool cand = c > 0 && split;                                     // :390
            obs = 1; mapped = cand; hits = c;                                     // :488-491
            if (cand) {
                if (c == 1) { if (do_counts & 1) atomicAdd(&o.per_bsj[m.ld(at)], 1ULL); }          // EquivalenceClassBuilder.hpp:112-130
                else if (do_counts & 2) {
                    const unsigned long long e = atomicAdd(&o.hdr->n_eq, 1ULL), sidx = atomicAdd(&o.hdr->n_eq_tids, (unsigned long long)c);
                    if (e < o.eq_cap && sidx + c <= o.eq_tids_cap) {
                        o.eq_ent[e] = make_uint4((uint32_t)rec, (uint32_t)sidx, c, 0u);
                        for (uint32_t t = 0; t < c; ++t) o.eq_tids[sidx + t] = m.ld(at + t);
                    }
                }
            }
        }
    }
    if (do_counts & 1) {
        obs = __reduce_add_sync(0xffffffffu, obs); mapped = __reduce_add_sync(0xffffffffu, mapped);
        hits = __reduce_add_sync(0xffffffffu, hits); upper = __reduce_add_sync(0xffffffffu, upper);
        if ((threadIdx.x & 31) == 0) {
            if (obs) atomicAdd(&o.ctr[C_BSJ_OBS], (unsigned long long)obs);
            if (mapped) atomicAdd(&o.ctr[C_BSJ_MAP], (unsigned long long)mapped);
            if (hits) atomicAdd(&o.ctr[C_BSJ_HITS], (unsigned long long)hits);
            if (upper) atomicAdd(&o.ctr[C_BSJ_UPPER], (unsigned long long)upper);
        }
    }
}

// ---- result lists derived on the device (the host used to loop over every exported record for them) -------------------
// export list -> (pair, mate code 1 / 2) per record; exp == nullptr: record r is mate 1 of pair r (stage 2)
__global__ void __launch_bounds__(256)
k_exp_split(const uint32_t* __restrict__ exp, const uint32_t* __restrict__ n_dev, uint64_t n_fixed, uint32_t* __restrict__ pair, uint8_t* __restrict__ code) {
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_fixed;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t m = exp ? exp[r] : (uint32_t)(2 * r);
        pair[r] = m >> 1; code[r] = (uint8_t)((m & 1) + 1);
    }
}
// candidate flag of every record slot 0 .. worst (Circall_bsj.cpp:385-390): hits in 1 .. maxReadOccs and a junction row
__global__ void __launch_bounds__(256)
k_cand_flag(const uint32_t* __restrict__ nh, const uint8_t* __restrict__ sp, const uint32_t* __restrict__ n_dev, uint64_t n_fixed, uint64_t worst,
            uint8_t* __restrict__ flag) {
    const uint64_t n = n_dev ? (uint64_t)*n_dev : n_fixed;
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= worst) return;
    flag[r] = (r < n && sp[r] && nh[r] > 0 && nh[r] <= (uint32_t)MAX_READ_OCCS) ? 1 : 0;
}

inline unsigned grid_for(uint64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// A batch counts into its own staging vector [per-BSJ counts | scalar counters]; this kernel, last on the batch's stream, adds it
// to the session's accumulator only when the batch ended clean (status 0) and clears it either way: a batch that fails with
// CG_ERR_BASE / LENGTH / CAPACITY leaves no partial counts behind (cg_session_counters, cg_session_accumulator and the all-reduce
// across GPUs only ever see whole batches).
__global__ void __launch_bounds__(256)
k_acc_fold(unsigned long long* __restrict__ batch, unsigned long long* __restrict__ acc, uint64_t n, const BatchHdr* hdr) {
    const bool ok = hdr->status == 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long v = batch[i];
        if (v) { if (ok) acc[i] += v; batch[i] = 0; }
    }
}

// offsets of n reads of one length (cg_session_submit_fixed: the offsets are not worth their bytes on PCIe)
__global__ void k_fixed_offsets(uint64_t* __restrict__ off, uint64_t n_plus_1, uint64_t len) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n_plus_1) off[i] = i * len;
}

} // namespace

struct cg_session {
    const cg_index* tx = nullptr;
    const cg_index* bsj = nullptr;
    cg_opts opts{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp = nullptr, evw = nullptr, evc = nullptr;
    uint64_t launches = 0;
    int warp_grid_wt = 0, warp_grid_bsj = 0;     // persistent grids of the warp-per-read kernels (SMs x resident blocks)
    int gen_grid = 0;                            // grid of the general kernels (sized with the per-warp global scratch)
    int sms = 0;
    bool use_pre = true;                         // Circall_wt front kernel for clean full-length matches (CG_NO_PRE=1 switches it off)
    bool use_thr = true;                         // lockstep thread tier in front of the warp kernels (CG_NO_THR=1 switches it off)
    unsigned char* d_gscratch = nullptr;
    bool pending = false;
    // batch state
    uint64_t n_pairs = 0;
    int max_len = 0;
    PackGeom geom{};
    bool input_on_device = false;
    const unsigned char* in_r1 = nullptr; const unsigned char* in_r2 = nullptr;   // device pointers of the batch
    const uint64_t* in_off1 = nullptr; const uint64_t* in_off2 = nullptr;
    // device buffers
    GrowBuf d_r1, d_r2, d_off1, d_off2, d_pk, d_lens, d_flags, d_nhits, d_expflag, d_mflags, d_mnhits, d_bc, d_exp,
        d_junc, d_eqent, d_eqtids, d_recnh, d_recsp, d_hdr, d_ovf, d_list, d_ovf2, d_list2, d_ovf3, d_list3, d_exppair, d_expcode, d_cand;
    GrowBuf d_ivhead_wt, d_ivoff_wt, d_ivpool_wt, d_ivhead_bsj, d_ivoff_bsj, d_ivpool_bsj, d_ivcur;   // cg_opts::dump_intervals
    GrowBuf d_ovf4, d_list4, d_work;     // second pass of the two-strand instantiation: retry flags, list, counts
    bool use_retry = true;
    GrowBuf d_thr_res, d_thr_iv;         // hand-over between the collector and the finishing kernels of the lockstep tier (thr_cap slots)
    uint64_t thr_cap = 0;
    uint32_t iv_cap = 0;
    PinVec<uint32_t> h_ivhead_wt, h_ivoff_wt, h_ivhead_bsj, h_ivoff_bsj;
    PinVec<int4> h_ivpool_wt, h_ivpool_bsj;
    unsigned long long* d_acc = nullptr;   // [n_bsj | 8 scalars]: totals of the batches that ended clean
    unsigned long long* d_accb = nullptr;  // the same for the batch in flight (k_acc_fold)
    uint64_t n_bsj = 0;
    uint64_t junc_cap = 0, eq_cap = 0, eq_tids_cap = 0;
    // host result buffers
    PinBuf h_hdr;
    // device -> host landing buffers are page-locked (a copy into pageable memory is staged and blocks the caller)
    PinVec<uint32_t> h_exp_pair, h_cand, h_recnh, h_eq_tids_raw;
    PinVec<uint8_t> h_exp_code, h_recsp, h_mflags;
    PinVec<uint16_t> h_mnhits;
    PinVec<cg_junction> h_junc;
    PinVec<uint4> h_eqent;
    std::vector<uint32_t> h_eq_tids;
    std::vector<uint64_t> h_eq_off;
    uint32_t wt_read_len = 0, bsj_read_len = 0;
};

namespace {

// ---- who owns the per-device constant index view (cgw::c_ix) ------------------------------------------------------------
// The kernels read the two index views from constant memory (the warp kernels are written for code size and take them as immediate
// operands), and a session rewrites them on its own stream before each stage.  Sessions on the same indices write the same bytes;
// a session on OTHER indices must not do so while batches of the current owner are still running.  The registry below makes that a
// property of the library instead of a rule for the caller: a session whose indices differ from the ones the constants hold first
// makes its stream wait for every batch in flight on the device (their end-of-batch events), then takes the constants over.
struct DeviceIx {
    std::mutex mu;
    const cg_index* owner[2] = {nullptr, nullptr};       // [0] txIdx, [1] bsjIdx currently in c_ix
    std::vector<cg_session*> inflight;                   // sessions with a batch submitted and not yet fetched
};
DeviceIx g_devix[64];

int acquire_ix(cg_session* s) {
    if (s->device < 0 || s->device >= 64) return CG_OK;
    DeviceIx& D = g_devix[s->device];
    std::lock_guard<std::mutex> lock(D.mu);
    const bool uses0 = s->opts.stage != 2, uses1 = s->opts.stage != 1;
    const bool clash = (uses0 && D.owner[0] && D.owner[0] != s->tx) || (uses1 && D.owner[1] && D.owner[1] != s->bsj);
    if (clash) {
        for (cg_session* o : D.inflight)
            if (o != s) CG_CUDA(cudaStreamWaitEvent(s->stream, o->ev1, 0));
        // (the owners' later submits wait for this batch in turn: they find their indices displaced)
    }
    if (uses0) D.owner[0] = s->tx;
    if (uses1) D.owner[1] = s->bsj;
    if (std::find(D.inflight.begin(), D.inflight.end(), s) == D.inflight.end()) D.inflight.push_back(s);
    return CG_OK;
}
void release_ix(cg_session* s) {
    if (s->device < 0 || s->device >= 64) return;
    DeviceIx& D = g_devix[s->device];
    std::lock_guard<std::mutex> lock(D.mu);
    D.inflight.erase(std::remove(D.inflight.begin(), D.inflight.end(), s), D.inflight.end());
}
void forget_index(const cg_session* s) {          // a destroyed session's indices may be closed next: never compare against a dangling owner
    if (s->device < 0 || s->device >= 64) return;
    DeviceIx& D = g_devix[s->device];
    std::lock_guard<std::mutex> lock(D.mu);
    bool used[2] = {false, false};
    for (cg_session* o : D.inflight) { if (o == s) continue; if (o->opts.stage != 2 && o->tx == D.owner[0]) used[0] = true; if (o->opts.stage != 1 && o->bsj == D.owner[1]) used[1] = true; }
    if (!used[0] && D.owner[0] == s->tx) D.owner[0] = nullptr;
    if (!used[1] && D.owner[1] == s->bsj) D.owner[1] = nullptr;
}

template <class K, class... A>
int launch(cg_session* s, K kern, dim3 grid, dim3 block, size_t smem, A... args) {
    if (grid.x == 0) return CG_OK;      // empty batch: nothing to do (the batch header is already zero)
    if (smem > 48 * 1024) CG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, s->stream>>>(args...);
    CG_CUDA(cudaGetLastError());
    ++s->launches;
    return CG_OK;
}

// ordered compaction of the set flags of f[0..n) into out[], count into *total
int compact_flags(cg_session* s, const uint8_t* f, uint64_t n, uint32_t* out, uint32_t* total) {
    uint32_t nb = grid_for(n, CP_TILE);
    int rc = launch(s, k_flag_count, dim3(nb), dim3(CP_THREADS), 0, f, n, s->d_bc.as<uint32_t>());
    if (rc) return rc;
    rc = launch(s, k_flag_scan, dim3(1), dim3(1024), 0, s->d_bc.as<uint32_t>(), nb, total);
    if (rc) return rc;
    return launch(s, k_flag_write, dim3(nb), dim3(CP_THREADS), 0, f, n, s->d_bc.as<uint32_t>(), out);
}

// out[*n_out ..) = list[*seg .. *n_list); *n_new = *n_out + (*n_list - *seg): the two-strand tail of a class-sorted list appended to
// the list of the lean instantiation's other-strand declines, so that ONE two-strand pass runs over both
__global__ void k_list_append(const uint32_t* __restrict__ list, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ n_list,
                              uint32_t* __restrict__ out, const uint32_t* __restrict__ n_out, uint32_t* __restrict__ n_new) {
    const uint32_t lo = *seg, hi = *n_list, base = *n_out, n = hi > lo ? hi - lo : 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[base + i] = list[lo + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_new = base + n;
}

// indices of the non-zero entries of the class array f[0..n), grouped by class (1..16); count into *total
int compact_classes(cg_session* s, const uint8_t* f, uint64_t n, uint32_t* out, uint32_t* total) {
    uint32_t nb = grid_for(n, CP_TILE);
    int rc = launch(s, k_cls_count, dim3(nb), dim3(CP_THREADS), 0, f, n, s->d_bc.as<uint32_t>(), nb);
    if (rc) return rc;
    rc = launch(s, k_flag_scan, dim3(1), dim3(1024), 0, s->d_bc.as<uint32_t>(), nb * CLS_N, total);
    if (rc) return rc;
    return launch(s, k_cls_write, dim3(nb), dim3(CP_THREADS), 0, f, n, (const uint32_t*)s->d_bc.as<uint32_t>(), nb, out);
}

int run_bsj_stage(cg_session* s, int do_counts) {
    const uint64_t n_reads = 2 * s->n_pairs;
    const bool stage2 = s->opts.stage == 2;
    const uint64_t worst = stage2 ? s->n_pairs : n_reads;
    if (worst == 0) return CG_OK;
    BsjOut o;
    o.junc = s->d_junc.as<cg_junction>(); o.junc_cap = s->junc_cap;
    o.eq_ent = s->d_eqent.as<uint4>(); o.eq_cap = s->eq_cap;
    o.eq_tids = s->d_eqtids.as<uint32_t>(); o.eq_tids_cap = s->eq_tids_cap;
    o.rec_nhits = s->d_recnh.as<uint32_t>(); o.rec_split = s->d_recsp.as<uint8_t>();
    o.per_bsj = s->d_accb; o.ctr = s->d_accb + s->n_bsj; o.hdr = s->d_hdr.as<BatchHdr>();
    const uint32_t* em = stage2 ? nullptr : s->d_exp.as<uint32_t>();
    IvDump dbs; memset(&dbs, 0, sizeof dbs);
    if (s->opts.dump_intervals && do_counts == 3) {          // (not on the redo pass of a batch whose row buffers were too small)
        dbs.head = s->d_ivhead_bsj.as<uint32_t>(); dbs.off = s->d_ivoff_bsj.as<uint32_t>(); dbs.pool = s->d_ivpool_bsj.as<int4>();
        dbs.cursor = s->d_ivcur.as<uint32_t>() + 1; dbs.cap = s->iv_cap;
        CG_CUDA(cudaMemsetAsync(dbs.head, 0, worst * 4, s->stream));
    }
    IndexView v = s->bsj->view();
    BatchHdr* dh = s->d_hdr.as<BatchHdr>();
    CG_CUDA(cudaMemsetAsync(s->d_ovf.p, 0, worst, s->stream));
    CG_CUDA(cudaMemcpyToSymbolAsync(cgw::c_ix, &v, sizeof(IndexView), sizeof(IndexView), cudaMemcpyHostToDevice, s->stream));
    const dim3 blk(WARPS_PER_BLOCK * 32);
    const size_t smf = (size_t)WARPS_PER_BLOCK * warp_smem_bytes(s->geom, false, true), smg = (size_t)WARPS_PER_BLOCK * warp_smem_bytes(s->geom, true, true);
    int rc;
    // tiers: lockstep thread tier, fast warp kernel, general warp kernel — each over the records
    // the one before it declined
    const uint32_t* cur_list = nullptr; const uint32_t* cur_n = nullptr;            // records still to map (null: all of them)
    if (s->use_thr) {
        // classify the records by their end k-mers, group them by class, map them one per thread
        CG_CUDA(cudaMemsetAsync(s->d_ovf3.p, 0, worst, s->stream));
        rc = launch(s, k_bsj_cls, dim3(grid_for(worst, 256)), dim3(256), 0, s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), s->geom, em, s->n_pairs,
                    em ? (const uint32_t*)&dh->n_exp : (const uint32_t*)nullptr, s->d_ovf3.as<uint8_t>());
        if (rc) return rc;
        rc = compact_classes(s, s->d_ovf3.as<uint8_t>(), worst, s->d_list.as<uint32_t>(), &dh->n_t2_bsj);
        if (rc) return rc;
        const uint32_t* seg = (const uint32_t*)(s->d_bc.as<uint32_t>() + (size_t)CLS_VOTE * grid_for(worst, CP_TILE));
        {
            const size_t sm0 = (size_t)THR_THREADS * cgq::QMemT<false>::words_col(s->geom.W2f + 1) * 4, sm1 = (size_t)THR_THREADS * cgq::QMemT<true>::words_col(s->geom.W2f + 1) * 4;
            const size_t sf0 = (size_t)THR_THREADS * cgq::QMemT<false>::words(0) * 4, sf1 = (size_t)THR_THREADS * cgq::QMemT<true>::words(0) * 4;
            const dim3 grid(grid_for(worst, THR_THREADS)), blk2(THR_THREADS);
            const uint32_t* l2 = (const uint32_t*)s->d_list.as<uint32_t>(); const uint32_t* n2 = (const uint32_t*)&dh->n_t2_bsj;
            const uint32_t* nexp = em ? (const uint32_t*)&dh->n_exp : (const uint32_t*)nullptr;
            uint32_t* res = s->d_thr_res.as<uint32_t>(); uint32_t* ivb = s->d_thr_iv.as<uint32_t>();
            rc = launch(s, k_thr_col<false, true>, grid, blk2, sm0, s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), s->geom, em, s->n_pairs, nexp, l2, n2, seg,
                        s->opts.lce_mode, (const uint8_t*)s->d_ovf3.as<uint8_t>(), res, ivb, s->thr_cap);
            if (rc) return rc;
            CG_CUDA(cudaMemsetAsync(s->d_ovf4.p, 0, worst, s->stream));
            uint8_t* retry = s->use_retry ? s->d_ovf4.as<uint8_t>() : (uint8_t*)nullptr;
            rc = launch(s, dbs.head ? k_bsj_fin<false, true> : k_bsj_fin<false, false>, grid, blk2, sf0, em, s->n_pairs, l2, n2, seg,
                        (const uint32_t*)res, (const uint32_t*)ivb, s->thr_cap, o, do_counts, s->d_ovf.as<uint8_t>(), retry, dbs);
            if (rc) return rc;
            // ONE two-strand pass: the tail of the class-sorted list and what the lean instantiation declined because the other strand
            // scored (a sixth of the warp kernels' cost per read)
            const uint32_t* l2s = l2; const uint32_t* n2s = n2; const uint32_t* segs = seg;
            if (retry) {
                uint32_t* n4 = s->d_work.as<uint32_t>() + 1;
                rc = compact_flags(s, retry, worst, s->d_list4.as<uint32_t>(), n4);
                if (rc) return rc;
                rc = launch(s, k_list_append, dim3(s->sms * 2), dim3(256), 0, l2, seg, n2, s->d_list4.as<uint32_t>(), (const uint32_t*)n4, n4 + 2);
                if (rc) return rc;
                l2s = s->d_list4.as<uint32_t>(); n2s = n4 + 2; segs = nullptr;
            }
            rc = launch(s, k_thr_col<true, true>, grid, blk2, sm1, s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), s->geom, em, s->n_pairs, nexp, l2s, n2s, segs,
                        s->opts.lce_mode, (const uint8_t*)s->d_ovf3.as<uint8_t>(), res, ivb, s->thr_cap);
            if (rc) return rc;
            rc = launch(s, dbs.head ? k_bsj_fin<true, true> : k_bsj_fin<true, false>, grid, blk2, sf1, em, s->n_pairs, l2s, n2s, segs,
                        (const uint32_t*)res, (const uint32_t*)ivb, s->thr_cap, o, do_counts, s->d_ovf.as<uint8_t>(), (uint8_t*)nullptr, dbs);
            if (rc) return rc;
        }
        rc = compact_flags(s, s->d_ovf.as<uint8_t>(), worst, s->d_list3.as<uint32_t>(), &dh->n_t3_bsj);
        if (rc) return rc;
        cur_list = s->d_list3.as<uint32_t>(); cur_n = &dh->n_t3_bsj;
    }
    CG_CUDA(cudaMemsetAsync(s->d_ovf2.p, 0, worst, s->stream));
    CG_CUDA(cudaMemsetAsync(s->d_work.as<uint32_t>() + 4, 0, 16, s->stream));          // draw counters of the two warp kernels
    const uint8_t* ovf_fast = s->d_ovf2.as<uint8_t>();
    if (cur_list)
        rc = launch(s, k_bsj_warp<false, true>, dim3(s->warp_grid_bsj), blk, smf,
                    s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), s->geom, em, s->n_pairs, cur_list, cur_n,
                    s->opts.lce_mode, o, do_counts, s->d_ovf2.as<uint8_t>(), (unsigned char*)nullptr, dbs, s->d_work.as<uint32_t>() + 4);
    else
        rc = launch(s, k_bsj_warp<false, false>, dim3(s->warp_grid_bsj), blk, smf,
                    s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), s->geom, em, s->n_pairs, (const uint32_t*)nullptr, (const uint32_t*)nullptr,
                    s->opts.lce_mode, o, do_counts, s->d_ovf2.as<uint8_t>(), (unsigned char*)nullptr, dbs, s->d_work.as<uint32_t>() + 4);
    if (rc) return rc;
    rc = compact_flags(s, ovf_fast, worst, s->d_list2.as<uint32_t>(), &dh->n_ovf_bsj);
    if (rc) return rc;
    rc = launch(s, k_bsj_warp<true, true>, dim3(s->gen_grid), blk, smg,
                s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), s->geom, em, s->n_pairs, (const uint32_t*)s->d_list2.as<uint32_t>(),
                (const uint32_t*)&dh->n_ovf_bsj, s->opts.lce_mode, o, do_counts, (uint8_t*)nullptr, s->d_gscratch, dbs, s->d_work.as<uint32_t>() + 5);
    if (rc) return rc;
    // the candidate list (records with a junction row and 1 .. maxReadOccs hits), in record order; d_ovf2 is free again
    rc = launch(s, k_cand_flag, dim3(grid_for(worst, 256)), dim3(256), 0, (const uint32_t*)o.rec_nhits, (const uint8_t*)o.rec_split,
                em ? (const uint32_t*)&dh->n_exp : (const uint32_t*)nullptr, s->n_pairs, worst, s->d_ovf2.as<uint8_t>());
    if (rc) return rc;
    return compact_flags(s, s->d_ovf2.as<uint8_t>(), worst, s->d_cand.as<uint32_t>(), &dh->n_cand);
}

int run_batch(cg_session* s) {
    const uint64_t n_pairs = s->n_pairs, n_reads = 2 * n_pairs;
    const int stage = s->opts.stage;
    s->geom = PackGeom::make(s->max_len);
    const PackGeom g = s->geom;
    CG_CUDA(s->d_pk.reserve(n_reads * g.WT * 8));
    CG_CUDA(s->d_lens.reserve(n_reads * 2));
    CG_CUDA(s->d_flags.reserve(n_reads));
    CG_CUDA(s->d_nhits.reserve(n_reads * 4));
    CG_CUDA(s->d_expflag.reserve(n_reads + 16));
    CG_CUDA(s->d_exp.reserve(n_reads * 4));
    CG_CUDA(s->d_bc.reserve((n_reads / CP_TILE + 2) * 4 * CLS_N));
    CG_CUDA(s->d_ovf.reserve(n_reads + 16));
    CG_CUDA(s->d_list.reserve(n_reads * 4 + 16));
    CG_CUDA(s->d_ovf2.reserve(n_reads + 16));
    CG_CUDA(s->d_list2.reserve(n_reads * 4 + 16));
    CG_CUDA(s->d_ovf3.reserve(n_reads + 16));
    CG_CUDA(s->d_list3.reserve(n_reads * 4 + 16));
    CG_CUDA(s->d_exppair.reserve(n_reads * 4 + 16));
    CG_CUDA(s->d_expcode.reserve(n_reads + 16));
    if (stage != 1) CG_CUDA(s->d_cand.reserve(n_reads * 4 + 16));
    CG_CUDA(s->d_hdr.reserve(sizeof(BatchHdr)));
    CG_CUDA(s->d_ovf4.reserve(n_reads + 16)); CG_CUDA(s->d_list4.reserve(n_reads * 4 + 16)); CG_CUDA(s->d_work.reserve(64));
    if (s->use_thr) { s->thr_cap = n_reads; CG_CUDA(s->d_thr_res.reserve(n_reads * 4 + 16)); CG_CUDA(s->d_thr_iv.reserve(n_reads * 4 * (uint64_t)THR_IVW + 16)); }
    if (s->opts.keep_mate_detail) { CG_CUDA(s->d_mflags.reserve(n_reads)); CG_CUDA(s->d_mnhits.reserve(n_reads * 2)); }
    if (stage != 1) {
        CG_CUDA(s->d_recnh.reserve(n_reads * 4));
        CG_CUDA(s->d_recsp.reserve(n_reads));
        uint64_t want_j = std::max<uint64_t>(1u << 20, 2 * n_reads);
        if (s->junc_cap < want_j) { CG_CUDA(s->d_junc.reserve(want_j * sizeof(cg_junction))); s->junc_cap = want_j; }
        uint64_t want_e = std::max<uint64_t>(1u << 18, n_reads / 4);
        if (s->eq_cap < want_e) { CG_CUDA(s->d_eqent.reserve(want_e * 16)); s->eq_cap = want_e; }
        uint64_t want_t = want_e * 8;
        if (s->eq_tids_cap < want_t) { CG_CUDA(s->d_eqtids.reserve(want_t * 4)); s->eq_tids_cap = want_t; }
    }
    { int arc = acquire_ix(s); if (arc) return arc; }
    CG_CUDA(cudaMemsetAsync(s->d_hdr.p, 0, sizeof(BatchHdr), s->stream));
    IvDump dwt; memset(&dwt, 0, sizeof dwt);
    if (s->opts.dump_intervals) {
        s->iv_cap = (uint32_t)std::min<uint64_t>(0xfffffff0ull, 24 * n_reads + 4096);
        CG_CUDA(s->d_ivcur.reserve(8));
        CG_CUDA(cudaMemsetAsync(s->d_ivcur.p, 0, 8, s->stream));
        if (stage != 2) {
            CG_CUDA(s->d_ivhead_wt.reserve(n_reads * 4 + 4)); CG_CUDA(s->d_ivoff_wt.reserve(n_reads * 4 + 4)); CG_CUDA(s->d_ivpool_wt.reserve((uint64_t)s->iv_cap * 16));
            dwt.head = s->d_ivhead_wt.as<uint32_t>(); dwt.off = s->d_ivoff_wt.as<uint32_t>(); dwt.pool = s->d_ivpool_wt.as<int4>();
            dwt.cursor = s->d_ivcur.as<uint32_t>(); dwt.cap = s->iv_cap;
            CG_CUDA(cudaMemsetAsync(dwt.head, 0, n_reads * 4 + 4, s->stream));
        }
        if (stage != 1) { CG_CUDA(s->d_ivhead_bsj.reserve(n_reads * 4 + 4)); CG_CUDA(s->d_ivoff_bsj.reserve(n_reads * 4 + 4)); CG_CUDA(s->d_ivpool_bsj.reserve((uint64_t)s->iv_cap * 16)); }
    }
    CG_CUDA(cudaEventRecord(s->ev0, s->stream));
    int rc;
    // K1
    rc = launch(s, k_pack, dim3(pack_grid(n_reads, g)), dim3(PACK_THREADS), 0, s->in_r1, s->in_off1, s->in_r2, s->in_off2, n_pairs, g,
                s->max_len, s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), &s->d_hdr.as<BatchHdr>()->status);
    if (rc) return rc;
    CG_CUDA(cudaEventRecord(s->evp, s->stream));
    if (stage != 2 && n_reads) {
        IndexView v = s->tx->view();
        BatchHdr* dh = s->d_hdr.as<BatchHdr>();
        CG_CUDA(cudaMemcpyToSymbolAsync(cgw::c_ix, &v, sizeof(IndexView), 0, cudaMemcpyHostToDevice, s->stream));
        const dim3 blk(WARPS_PER_BLOCK * 32);
        const size_t smf = (size_t)WARPS_PER_BLOCK * warp_smem_bytes(g, false, false), smg = (size_t)WARPS_PER_BLOCK * warp_smem_bytes(g, true, false);
        // the tiers, each over the mates the one before it declined: front kernel (clean full-length matches, one thread
        // per mate), lockstep thread tier (single-strand reads with mismatches), fast
        // warp kernel, general warp kernel
        const uint32_t* cur_list = nullptr; const uint32_t* cur_n = nullptr;        // mates still to map (null: all of them)
        if (s->use_pre) {
            rc = launch(s, dwt.head ? k_wt_pre<true> : k_wt_pre<false>, dim3(grid_for(n_reads, PRE_THREADS)), dim3(PRE_THREADS), (size_t)PRE_THREADS * cgp::pre_words(g.W2f + 1) * 4,
                        s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), g, n_reads, s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), s->d_ovf.as<uint8_t>(), dwt);
            if (rc) return rc;
            // the front kernel leaves a look-up class per declined mate: the lockstep tier wants its work grouped by it
            rc = s->use_thr ? compact_classes(s, s->d_ovf.as<uint8_t>(), n_reads, s->d_list.as<uint32_t>(), &dh->n_t2_wt)
                                             : compact_flags(s, s->d_ovf.as<uint8_t>(), n_reads, s->d_list.as<uint32_t>(), &dh->n_t2_wt);
            if (rc) return rc;
            cur_list = s->d_list.as<uint32_t>(); cur_n = &dh->n_t2_wt;
        }
        if (s->use_thr) {
            CG_CUDA(cudaMemsetAsync(s->d_ovf3.p, 0, n_reads, s->stream));
            // the class-sorted list has its two-strand classes at the end (compact_classes): lean instantiation over the head,
            // two-strand instantiation over the tail; an unsorted list (or none) goes through the lean one alone
            const bool sorted = cur_list && s->use_pre;
            const uint32_t* seg = sorted ? (const uint32_t*)(s->d_bc.as<uint32_t>() + (size_t)CLS_VOTE * grid_for(n_reads, CP_TILE)) : (const uint32_t*)nullptr;
            const size_t sm0 = (size_t)THR_THREADS * cgq::QMemT<false>::words_col(g.W2f + 1) * 4, sm1 = (size_t)THR_THREADS * cgq::QMemT<true>::words_col(g.W2f + 1) * 4;
            const size_t sf0 = (size_t)THR_THREADS * cgq::QMemT<false>::words(0) * 4, sf1 = (size_t)THR_THREADS * cgq::QMemT<true>::words(0) * 4;
            const dim3 grid(grid_for(n_reads, THR_THREADS)), blk2(THR_THREADS);
            const uint8_t* clsp = sorted ? (const uint8_t*)s->d_ovf.as<uint8_t>() : (const uint8_t*)nullptr;
            uint32_t* res = s->d_thr_res.as<uint32_t>(); uint32_t* ivb = s->d_thr_iv.as<uint32_t>();
            // lean collector + finish over the head of the class-sorted list; then ONE two-strand pass over its tail and over what the
            // lean one declined because the other strand scored
            rc = launch(s, k_thr_col<false, false>, grid, blk2, sm0, s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), g, (const uint32_t*)nullptr, n_reads, (const uint32_t*)nullptr,
                        cur_list, cur_n, seg, s->opts.lce_mode, clsp, res, ivb, s->thr_cap);
            if (rc) return rc;
            CG_CUDA(cudaMemsetAsync(s->d_ovf4.p, 0, n_reads, s->stream));
            uint8_t* retry = (s->use_retry && sorted) ? s->d_ovf4.as<uint8_t>() : (uint8_t*)nullptr;
            rc = launch(s, dwt.head ? k_wt_fin<false, true> : k_wt_fin<false, false>, grid, blk2, sf0, s->d_lens.as<uint16_t>(), n_reads, cur_list, cur_n, seg,
                        (const uint32_t*)res, (const uint32_t*)ivb, s->thr_cap, s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), s->d_ovf3.as<uint8_t>(), retry, dwt);
            if (rc) return rc;
            if (sorted) {
                const uint32_t* l2s = cur_list; const uint32_t* n2s = cur_n; const uint32_t* segs = seg;      // the two-strand pass's list
                if (retry) {
                    uint32_t* n4 = s->d_work.as<uint32_t>() + 0;
                    rc = compact_flags(s, retry, n_reads, s->d_list4.as<uint32_t>(), n4);
                    if (rc) return rc;
                    rc = launch(s, k_list_append, dim3(s->sms * 2), dim3(256), 0, cur_list, seg, cur_n, s->d_list4.as<uint32_t>(), (const uint32_t*)n4, n4 + 2);
                    if (rc) return rc;
                    l2s = s->d_list4.as<uint32_t>(); n2s = n4 + 2; segs = nullptr;
                }
                rc = launch(s, k_thr_col<true, false>, grid, blk2, sm1, s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), g, (const uint32_t*)nullptr, n_reads, (const uint32_t*)nullptr,
                            l2s, n2s, segs, s->opts.lce_mode, clsp, res, ivb, s->thr_cap);
                if (rc) return rc;
                rc = launch(s, dwt.head ? k_wt_fin<true, true> : k_wt_fin<true, false>, grid, blk2, sf1, s->d_lens.as<uint16_t>(), n_reads, l2s, n2s, segs,
                            (const uint32_t*)res, (const uint32_t*)ivb, s->thr_cap, s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), s->d_ovf3.as<uint8_t>(), (uint8_t*)nullptr, dwt);
                if (rc) return rc;
            }
            rc = compact_flags(s, s->d_ovf3.as<uint8_t>(), n_reads, s->d_list3.as<uint32_t>(), &dh->n_t3_wt);
            if (rc) return rc;
            cur_list = s->d_list3.as<uint32_t>(); cur_n = &dh->n_t3_wt;
        }
        CG_CUDA(cudaMemsetAsync(s->d_ovf2.p, 0, n_reads, s->stream));
        CG_CUDA(cudaMemsetAsync(s->d_work.as<uint32_t>() + 4, 0, 16, s->stream));      // draw counters of the two warp kernels
        const uint8_t* ovf_fast = s->d_ovf2.as<uint8_t>();
        if (cur_list)
            rc = launch(s, k_wt_warp<false, true>, dim3(s->warp_grid_wt), blk, smf,
                        s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), g, n_reads, cur_list, cur_n, s->opts.lce_mode,
                        s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), s->d_ovf2.as<uint8_t>(), (unsigned char*)nullptr, dwt, s->d_work.as<uint32_t>() + 4);
        else
            rc = launch(s, k_wt_warp<false, false>, dim3(s->warp_grid_wt), blk, smf,
                        s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), g, n_reads, (const uint32_t*)nullptr, (const uint32_t*)nullptr, s->opts.lce_mode,
                        s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), s->d_ovf2.as<uint8_t>(), (unsigned char*)nullptr, dwt, s->d_work.as<uint32_t>() + 4);
        if (rc) return rc;
        rc = compact_flags(s, ovf_fast, n_reads, s->d_list2.as<uint32_t>(), &dh->n_ovf_wt);
        if (rc) return rc;
        rc = launch(s, k_wt_warp<true, true>, dim3(s->gen_grid), blk, smg,
                    s->d_pk.as<uint64_t>(), s->d_lens.as<uint16_t>(), g, n_reads, (const uint32_t*)s->d_list2.as<uint32_t>(), (const uint32_t*)&dh->n_ovf_wt,
                    s->opts.lce_mode, s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), (uint8_t*)nullptr, s->d_gscratch, dwt, s->d_work.as<uint32_t>() + 5);
        if (rc) return rc;
        CG_CUDA(cudaEventRecord(s->evw, s->stream));
        rc = launch(s, k_wt_commit, dim3(grid_for(n_pairs, 256)), dim3(256), 0, s->d_flags.as<uint8_t>(), s->d_nhits.as<uint32_t>(), n_pairs,
                    s->d_expflag.as<uint8_t>(), s->opts.keep_mate_detail ? s->d_mflags.as<uint8_t>() : (uint8_t*)nullptr,
                    s->opts.keep_mate_detail ? s->d_mnhits.as<uint16_t>() : (uint16_t*)nullptr, s->d_accb + s->n_bsj, dh);
        if (rc) return rc;
        rc = compact_flags(s, s->d_expflag.as<uint8_t>(), n_reads, s->d_exp.as<uint32_t>(), &dh->n_exp);
        if (rc) return rc;
        rc = launch(s, k_exp_split, dim3(s->sms * 8), dim3(256), 0, (const uint32_t*)s->d_exp.as<uint32_t>(), (const uint32_t*)&dh->n_exp, (uint64_t)0,
                    s->d_exppair.as<uint32_t>(), s->d_expcode.as<uint8_t>());
        if (rc) return rc;
    } else if (stage != 2) {
        CG_CUDA(cudaEventRecord(s->evw, s->stream));
    }
    if (stage == 2) {
        CG_CUDA(cudaEventRecord(s->evw, s->stream));
        rc = launch(s, k_exp_split, dim3(s->sms * 8), dim3(256), 0, (const uint32_t*)nullptr, (const uint32_t*)nullptr, n_pairs,
                    s->d_exppair.as<uint32_t>(), s->d_expcode.as<uint8_t>());
        if (rc) return rc;
    }
    CG_CUDA(cudaEventRecord(s->evc, s->stream));
    if (stage != 1) {
        rc = run_bsj_stage(s, 3);
        if (rc) return rc;
    }
    rc = launch(s, k_acc_fold, dim3(s->sms * 4), dim3(256), 0, s->d_accb, s->d_acc, s->n_bsj + (uint64_t)C_NUM, (const BatchHdr*)s->d_hdr.as<BatchHdr>());
    if (rc) return rc;
    CG_CUDA(cudaEventRecord(s->ev1, s->stream));
    CG_CUDA(cudaMemcpyAsync(s->h_hdr.p, s->d_hdr.p, sizeof(BatchHdr), cudaMemcpyDeviceToHost, s->stream));
    s->pending = true;
    return CG_OK;
}

template <class T>
int d2h(cg_session* s, PinVec<T>& dst, const void* src, uint64_t n) {
    CG_CUDA(dst.resize(n));
    if (n) CG_CUDA(cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    return CG_OK;
}

} // namespace

extern "C" {

int cg_session_create(const cg_index* tx, const cg_index* bsj, const cg_opts* opts, cg_session** out) {
    if (!out) CG_FAIL(CG_ERR_ARG, "null argument");
    cg_opts o{};
    if (opts) o = *opts;
    if (o.stage < 0 || o.stage > 2) CG_FAIL(CG_ERR_ARG, "stage must be 0, 1 or 2");
    if (o.stage != 2 && !tx) CG_FAIL(CG_ERR_ARG, "this stage needs the transcript index");
    if (o.stage != 1 && !bsj) CG_FAIL(CG_ERR_ARG, "this stage needs the BSJ index");
    if (tx && bsj && tx->device != bsj->device) CG_FAIL(CG_ERR_ARG, "both indices must live on the same device");
    if (o.lce_mode != 0 && o.lce_mode != 1) CG_FAIL(CG_ERR_ARG, "lce_mode must be 0 or 1");
    int device = tx ? tx->device : bsj->device;
    int nd = 0;
    if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) { cudaGetLastError(); CG_FAIL(CG_ERR_CUDA, "no CUDA device available (libcircall_gpu has no CPU fallback)"); }
    CG_CUDA(cudaSetDevice(device));
    cgi::device_tuning(device);
    cg_session* s = new cg_session();
    s->tx = tx; s->bsj = bsj; s->opts = o; s->device = device;
    s->n_bsj = (o.stage != 1 && bsj) ? bsj->n_tx : 0;
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&s->evp);
    if (e == cudaSuccess) e = cudaEventCreate(&s->evw);
    if (e == cudaSuccess) e = cudaEventCreate(&s->evc);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_acc, (s->n_bsj + C_NUM) * 8);
    if (e == cudaSuccess) e = cudaMemset(s->d_acc, 0, (s->n_bsj + C_NUM) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_accb, (s->n_bsj + C_NUM) * 8);
    if (e == cudaSuccess) e = cudaMemset(s->d_accb, 0, (s->n_bsj + C_NUM) * 8);
    if (e == cudaSuccess) e = s->h_hdr.reserve(sizeof(BatchHdr));
    if (e == cudaSuccess) {
        // persistent grids: as many blocks as stay resident (sized for the largest read geometry in common use)
        int sms = 0, bw = 0, bb = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        PackGeom g = PackGeom::make(160);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bw, k_wt_warp<false, false>, WARPS_PER_BLOCK * 32, (size_t)WARPS_PER_BLOCK * warp_smem_bytes(g, false, false));
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, k_bsj_warp<false, false>, WARPS_PER_BLOCK * 32, (size_t)WARPS_PER_BLOCK * warp_smem_bytes(g, false, true));
        s->sms = sms;
        s->use_pre = getenv("CG_NO_PRE") == nullptr;
        s->use_thr = getenv("CG_NO_THR") == nullptr;
        s->use_retry = getenv("CG_NO_RETRY") == nullptr;
        // (the lockstep tier keeps table SECTOR numbers in 32 bits, cg_thr.cuh::q_walk: good for 2^33 slots = 137 GB of table, more than
        // the index builder can allocate next to the rest of an index, so no table size switches a tier off)
        s->warp_grid_wt = sms * (bw > 0 ? bw : 1);
        s->warp_grid_bsj = sms * (bb > 0 ? bb : 1);
        s->gen_grid = sms * CG_GEN_MINB;
        if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_gscratch, (size_t)s->gen_grid * WARPS_PER_BLOCK * GEN_SCRATCH);
    }
    if (e != cudaSuccess) { int rc = cgi::cuda_fail(e, "session set-up", __FILE__, __LINE__); cg_session_destroy(s); return rc; }
    *out = s;
    return CG_OK;
}

void cg_session_destroy(cg_session* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    release_ix(s); forget_index(s);
#ifdef CG_WARP_PROF
    {
        unsigned long long h[64];
        if (cudaMemcpyFromSymbol(h, cgw::g_wprof, sizeof h) == cudaSuccess) {
            fprintf(stderr, "WPROF (k_wt_warp, cycles of lane 0 summed over warps): look %llu extend %llu lce %llu vote %llu hits %llu | per-read total %llu reads fast %llu gen %llu\n",
                    h[0], h[1], h[2], h[3], h[4], h[7], h[8], h[9]);
            fprintf(stderr, "WPROF reads by log2(cycles), fast kernel:");
            for (int b = 8; b < 24; ++b) fprintf(stderr, " %d:%llu", b, h[16 + b]);
            fprintf(stderr, "\nWPROF reads by log2(cycles), general kernel:");
            for (int b = 8; b < 24; ++b) fprintf(stderr, " %d:%llu", b, h[40 + b]);
            fprintf(stderr, "\n");
            memset(h, 0, sizeof h);
            cudaMemcpyToSymbol(cgw::g_wprof, h, sizeof h);
        }
    }
#endif
    for (GrowBuf* b : {&s->d_r1, &s->d_r2, &s->d_off1, &s->d_off2, &s->d_pk, &s->d_lens, &s->d_flags, &s->d_nhits, &s->d_expflag,
                       &s->d_mflags, &s->d_mnhits, &s->d_bc, &s->d_exp, &s->d_junc, &s->d_eqent, &s->d_eqtids, &s->d_recnh, &s->d_recsp, &s->d_hdr,
                       &s->d_ovf, &s->d_list, &s->d_ovf2, &s->d_list2, &s->d_ovf3, &s->d_list3, &s->d_exppair, &s->d_expcode, &s->d_cand,
                       &s->d_ivhead_wt, &s->d_ivoff_wt, &s->d_ivpool_wt, &s->d_ivhead_bsj, &s->d_ivoff_bsj, &s->d_ivpool_bsj, &s->d_ivcur, &s->d_thr_res, &s->d_thr_iv, &s->d_ovf4, &s->d_list4, &s->d_work})
        b->release();
    s->h_hdr.release();
    s->h_exp_pair.release(); s->h_exp_code.release(); s->h_cand.release(); s->h_recnh.release(); s->h_eq_tids_raw.release(); s->h_recsp.release(); s->h_mflags.release();
    s->h_mnhits.release(); s->h_junc.release(); s->h_eqent.release();
    s->h_ivhead_wt.release(); s->h_ivoff_wt.release(); s->h_ivhead_bsj.release(); s->h_ivoff_bsj.release(); s->h_ivpool_wt.release(); s->h_ivpool_bsj.release();
    if (s->d_acc) cudaFree(s->d_acc);
    if (s->d_accb) cudaFree(s->d_accb);
    if (s->d_gscratch) cudaFree(s->d_gscratch);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->evp) cudaEventDestroy(s->evp);
    if (s->evw) cudaEventDestroy(s->evw);
    if (s->evc) cudaEventDestroy(s->evc);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int cg_session_submit(cg_session* s, const char* r1, const uint64_t* off1, const char* r2, const uint64_t* off2, uint64_t n_pairs) {
    if (!s || !off1 || !off2 || (n_pairs && (!r1 || !r2))) CG_FAIL(CG_ERR_ARG, "null argument");
    if (s->pending) CG_FAIL(CG_ERR_STATE, "previous batch not fetched");
    if (n_pairs >= (1ULL << 30)) CG_FAIL(CG_ERR_ARG, "batch too large (limit 2^30 pairs)");
    CG_CUDA(cudaSetDevice(s->device));
    uint64_t bytes1 = off1[n_pairs] - off1[0], bytes2 = off2[n_pairs] - off2[0];
    if (off1[0] != 0 || off2[0] != 0) CG_FAIL(CG_ERR_ARG, "offsets must start at 0");
    uint64_t mx = 0;
    for (uint64_t i = 0; i < n_pairs; ++i) {
        if (off1[i + 1] < off1[i] || off2[i + 1] < off2[i]) CG_FAIL(CG_ERR_ARG, "offsets must be non-decreasing");
        mx = std::max(mx, std::max(off1[i + 1] - off1[i], off2[i + 1] - off2[i]));
    }
    if (mx > CG_MAX_READ_LEN) CG_FAIL(CG_ERR_LENGTH, "a read is longer than CG_MAX_READ_LEN");
    CG_CUDA(s->d_r1.reserve(bytes1 + 64)); CG_CUDA(s->d_r2.reserve(bytes2 + 64));
    CG_CUDA(s->d_off1.reserve((n_pairs + 1) * 8)); CG_CUDA(s->d_off2.reserve((n_pairs + 1) * 8));
    if (bytes1) CG_CUDA(cudaMemcpyAsync(s->d_r1.p, r1, bytes1, cudaMemcpyHostToDevice, s->stream));
    if (bytes2) CG_CUDA(cudaMemcpyAsync(s->d_r2.p, r2, bytes2, cudaMemcpyHostToDevice, s->stream));
    CG_CUDA(cudaMemcpyAsync(s->d_off1.p, off1, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s->stream));
    CG_CUDA(cudaMemcpyAsync(s->d_off2.p, off2, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s->stream));
    s->in_r1 = s->d_r1.as<unsigned char>(); s->in_r2 = s->d_r2.as<unsigned char>();
    s->in_off1 = s->d_off1.as<uint64_t>(); s->in_off2 = s->d_off2.as<uint64_t>();
    s->n_pairs = n_pairs; s->max_len = (int)mx; s->input_on_device = false;
    if (s->wt_read_len == 0 && n_pairs) s->wt_read_len = (uint32_t)(off1[1] - off1[0]);
    return run_batch(s);
}

int cg_session_submit_fixed(cg_session* s, const char* r1, const char* r2, uint64_t n_pairs, uint32_t read_len) {
    if (!s || (n_pairs && (!r1 || !r2))) CG_FAIL(CG_ERR_ARG, "null argument");
    if (s->pending) CG_FAIL(CG_ERR_STATE, "previous batch not fetched");
    if (n_pairs >= (1ULL << 30)) CG_FAIL(CG_ERR_ARG, "batch too large (limit 2^30 pairs)");
    if (read_len == 0) CG_FAIL(CG_ERR_ARG, "read_len must be positive");
    if (read_len > CG_MAX_READ_LEN) CG_FAIL(CG_ERR_LENGTH, "a read is longer than CG_MAX_READ_LEN");
    CG_CUDA(cudaSetDevice(s->device));
    uint64_t bytes = n_pairs * (uint64_t)read_len;
    CG_CUDA(s->d_r1.reserve(bytes + 64)); CG_CUDA(s->d_r2.reserve(bytes + 64));
    CG_CUDA(s->d_off1.reserve((n_pairs + 1) * 8));
    if (bytes) CG_CUDA(cudaMemcpyAsync(s->d_r1.p, r1, bytes, cudaMemcpyHostToDevice, s->stream));
    if (bytes) CG_CUDA(cudaMemcpyAsync(s->d_r2.p, r2, bytes, cudaMemcpyHostToDevice, s->stream));
    k_fixed_offsets<<<grid_for(n_pairs + 1, 256), 256, 0, s->stream>>>(s->d_off1.as<uint64_t>(), n_pairs + 1, read_len);
    CG_CUDA(cudaGetLastError());
    ++s->launches;
    s->in_r1 = s->d_r1.as<unsigned char>(); s->in_r2 = s->d_r2.as<unsigned char>();
    s->in_off1 = s->d_off1.as<uint64_t>(); s->in_off2 = s->d_off1.as<uint64_t>();
    s->n_pairs = n_pairs; s->max_len = (int)read_len; s->input_on_device = false;
    if (s->wt_read_len == 0 && n_pairs) s->wt_read_len = read_len;
    return run_batch(s);
}

int cg_session_submit_device(cg_session* s, const char* d_r1, const uint64_t* d_off1, const char* d_r2, const uint64_t* d_off2,
                             uint64_t n_pairs, uint32_t max_read_len) {
    if (!s || !d_off1 || !d_off2 || (n_pairs && (!d_r1 || !d_r2))) CG_FAIL(CG_ERR_ARG, "null argument");
    if (s->pending) CG_FAIL(CG_ERR_STATE, "previous batch not fetched");
    if (n_pairs >= (1ULL << 30)) CG_FAIL(CG_ERR_ARG, "batch too large (limit 2^30 pairs)");
    if (max_read_len == 0 || max_read_len > CG_MAX_READ_LEN) CG_FAIL(CG_ERR_LENGTH, "max_read_len must be in 1..CG_MAX_READ_LEN");
    CG_CUDA(cudaSetDevice(s->device));
    s->in_r1 = (const unsigned char*)d_r1; s->in_r2 = (const unsigned char*)d_r2; s->in_off1 = d_off1; s->in_off2 = d_off2;
    s->n_pairs = n_pairs; s->max_len = (int)max_read_len; s->input_on_device = true;
    return run_batch(s);
}

int cg_session_fetch(cg_session* s, int want_lists, cg_batch_result* out) {
    if (!s || !out) CG_FAIL(CG_ERR_ARG, "null argument");
    if (!s->pending) CG_FAIL(CG_ERR_STATE, "no batch submitted");
    CG_CUDA(cudaSetDevice(s->device));
    CG_CUDA(cudaStreamSynchronize(s->stream));
    s->pending = false;
    release_ix(s);
    BatchHdr hdr = *s->h_hdr.as<BatchHdr>();
    if (hdr.status & ST_BADBASE) CG_FAIL(CG_ERR_BASE, "a read holds a character outside ACGTNacgtn");
    if (hdr.status & ST_TOOLONG) CG_FAIL(CG_ERR_LENGTH, "a read is longer than the batch's max_read_len");
    if (hdr.status & ST_CAPACITY) CG_FAIL(CG_ERR_CAPACITY, "collector capacity exceeded (intervals / k-mer scores / hits)");
    float ms = 0, ms_k[4] = {0, 0, 0, 0};
    CG_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    CG_CUDA(cudaEventElapsedTime(&ms_k[0], s->ev0, s->evp));
    CG_CUDA(cudaEventElapsedTime(&ms_k[1], s->evp, s->evw));
    CG_CUDA(cudaEventElapsedTime(&ms_k[2], s->evw, s->evc));
    CG_CUDA(cudaEventElapsedTime(&ms_k[3], s->evc, s->ev1));
    const int stage = s->opts.stage;
    uint64_t n_exp = stage == 2 ? s->n_pairs : hdr.n_exp;
    // the span filter / class lists never drop rows: when a list outgrew its buffer, grow it and redo the
    // Circall_bsj stage of this batch with the counters switched off (they were already accumulated)
    if (stage != 1 && (hdr.n_junc > s->junc_cap || hdr.n_eq > s->eq_cap || hdr.n_eq_tids > s->eq_tids_cap)) {
        if (hdr.n_junc > s->junc_cap) { CG_CUDA(s->d_junc.reserve((hdr.n_junc + 1024) * sizeof(cg_junction))); s->junc_cap = hdr.n_junc + 1024; }
        if (hdr.n_eq > s->eq_cap) { CG_CUDA(s->d_eqent.reserve((hdr.n_eq + 1024) * 16)); s->eq_cap = hdr.n_eq + 1024; }
        if (hdr.n_eq_tids > s->eq_tids_cap) { CG_CUDA(s->d_eqtids.reserve((hdr.n_eq_tids + 1024) * 4)); s->eq_tids_cap = hdr.n_eq_tids + 1024; }
        BatchHdr* dh = s->d_hdr.as<BatchHdr>();
        CG_CUDA(cudaMemsetAsync(&dh->n_junc, 0, 3 * sizeof(unsigned long long), s->stream));
        int rc = acquire_ix(s);
        if (rc) return rc;
        rc = run_bsj_stage(s, 2);
        if (rc) { release_ix(s); return rc; }
        CG_CUDA(cudaEventRecord(s->ev1, s->stream));
        CG_CUDA(cudaMemcpyAsync(s->h_hdr.p, s->d_hdr.p, sizeof(BatchHdr), cudaMemcpyDeviceToHost, s->stream));
        CG_CUDA(cudaStreamSynchronize(s->stream));
        release_ix(s);
        BatchHdr h2 = *s->h_hdr.as<BatchHdr>();
        if (h2.n_junc != hdr.n_junc || h2.n_eq != hdr.n_eq || h2.n_eq_tids != hdr.n_eq_tids) CG_FAIL(CG_ERR_STATE, "redo of the bsj stage disagreed with the first pass");
    }
    memset(out, 0, sizeof *out);
    out->n_pairs = s->n_pairs;
    out->device_ms = ms;
    for (int x = 0; x < 4; ++x) out->kernel_ms[x] = ms_k[x];
    out->n_export = n_exp;
    out->n_junc = hdr.n_junc;
    out->n_eq = hdr.n_eq;
    out->n_cand = stage != 1 ? hdr.n_cand : 0;
    out->n_fallback_wt = hdr.n_ovf_wt;
    out->n_fallback_bsj = hdr.n_ovf_bsj;
    out->n_tier2_wt = s->use_thr ? hdr.n_t3_wt : hdr.n_t2_wt;          // reads that reached the warp kernels
    out->n_tier2_bsj = s->use_thr ? hdr.n_t3_bsj : hdr.n_t2_bsj;
    if (s->input_on_device && s->wt_read_len == 0 && s->n_pairs) {
        uint16_t l0 = 0;
        CG_CUDA(cudaMemcpy(&l0, s->d_lens.p, 2, cudaMemcpyDeviceToHost));
        s->wt_read_len = l0;
    }
    if (stage != 1 && s->bsj_read_len == 0 && n_exp) {
        uint32_t m0 = 0; uint16_t l0 = 0;
        if (stage != 2) CG_CUDA(cudaMemcpy(&m0, s->d_exp.p, 4, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(&l0, s->d_lens.as<uint16_t>() + m0, 2, cudaMemcpyDeviceToHost));
        s->bsj_read_len = l0;
    }
    if (!want_lists) return CG_OK;
    int rc;
    if ((rc = d2h(s, s->h_exp_pair, s->d_exppair.p, n_exp))) return rc;
    if ((rc = d2h(s, s->h_exp_code, s->d_expcode.p, n_exp))) return rc;
    if (s->opts.keep_mate_detail && stage != 2) {
        if ((rc = d2h(s, s->h_mflags, s->d_mflags.p, 2 * s->n_pairs))) return rc;
        if ((rc = d2h(s, s->h_mnhits, s->d_mnhits.p, 2 * s->n_pairs))) return rc;
    }
    if (stage != 1) {
        if ((rc = d2h(s, s->h_junc, s->d_junc.p, hdr.n_junc))) return rc;
        if ((rc = d2h(s, s->h_cand, s->d_cand.p, hdr.n_cand))) return rc;
        if ((rc = d2h(s, s->h_recnh, s->d_recnh.p, n_exp))) return rc;
        if ((rc = d2h(s, s->h_recsp, s->d_recsp.p, n_exp))) return rc;
        if ((rc = d2h(s, s->h_eqent, s->d_eqent.p, hdr.n_eq))) return rc;
        if ((rc = d2h(s, s->h_eq_tids_raw, s->d_eqtids.p, hdr.n_eq_tids))) return rc;
    }
    uint32_t ivn[2] = {0, 0};
    if (s->opts.dump_intervals) {
        CG_CUDA(cudaMemcpyAsync(ivn, s->d_ivcur.p, 8, cudaMemcpyDeviceToHost, s->stream));
        CG_CUDA(cudaStreamSynchronize(s->stream));
        if (ivn[0] > s->iv_cap || ivn[1] > s->iv_cap) CG_FAIL(CG_ERR_CAPACITY, "interval dump buffer too small for this batch");
        if (stage != 2) {
            if ((rc = d2h(s, s->h_ivhead_wt, s->d_ivhead_wt.p, 2 * s->n_pairs))) return rc;
            if ((rc = d2h(s, s->h_ivoff_wt, s->d_ivoff_wt.p, 2 * s->n_pairs))) return rc;
            if ((rc = d2h(s, s->h_ivpool_wt, s->d_ivpool_wt.p, ivn[0]))) return rc;
        }
        if (stage != 1) {
            if ((rc = d2h(s, s->h_ivhead_bsj, s->d_ivhead_bsj.p, n_exp))) return rc;
            if ((rc = d2h(s, s->h_ivoff_bsj, s->d_ivoff_bsj.p, n_exp))) return rc;
            if ((rc = d2h(s, s->h_ivpool_bsj, s->d_ivpool_bsj.p, ivn[1]))) return rc;
        }
    }
    CG_CUDA(cudaStreamSynchronize(s->stream));
    if (s->opts.dump_intervals) {
        if (stage != 2) { out->wt_iv_head = s->h_ivhead_wt.data(); out->wt_iv_off = s->h_ivoff_wt.data(); out->wt_iv = (const int32_t*)s->h_ivpool_wt.data(); out->n_wt_iv = ivn[0]; }
        if (stage != 1) { out->bsj_iv_head = s->h_ivhead_bsj.data(); out->bsj_iv_off = s->h_ivoff_bsj.data(); out->bsj_iv = (const int32_t*)s->h_ivpool_bsj.data(); out->n_bsj_iv = ivn[1]; }
    }
    out->exp_pair = s->h_exp_pair.data(); out->exp_code = s->h_exp_code.data();
    if (s->opts.keep_mate_detail && stage != 2) { out->mate_flags = s->h_mflags.data(); out->mate_nhits = s->h_mnhits.data(); }
    if (stage != 1) {
        out->junc = s->h_junc.data();
        out->rec_nhits = s->h_recnh.data(); out->rec_split = s->h_recsp.data();
        out->cand_rec = s->h_cand.data();
        // classes in record order
        std::sort(s->h_eqent.begin(), s->h_eqent.end(), [](const uint4& a, const uint4& b) { return a.x < b.x; });
        s->h_eq_off.assign(1, 0); s->h_eq_tids.clear();
        for (const uint4& e : s->h_eqent) {
            s->h_eq_tids.insert(s->h_eq_tids.end(), s->h_eq_tids_raw.begin() + e.y, s->h_eq_tids_raw.begin() + e.y + e.z);
            s->h_eq_off.push_back(s->h_eq_tids.size());
        }
        out->eq_off = s->h_eq_off.data(); out->eq_tids = s->h_eq_tids.data();
    }
    return CG_OK;
}

int cg_session_counters(cg_session* s, cg_counters* c, uint64_t* per_bsj, uint64_t n_bsj) {
    if (!s || !c) CG_FAIL(CG_ERR_ARG, "null argument");
    if (s->pending) CG_FAIL(CG_ERR_STATE, "fetch the pending batch first");
    if (per_bsj && n_bsj != s->n_bsj) CG_FAIL(CG_ERR_ARG, "per_bsj must hold one entry per BSJ reference");
    CG_CUDA(cudaSetDevice(s->device));
    unsigned long long v[C_NUM];
    CG_CUDA(cudaMemcpy(v, s->d_acc + s->n_bsj, sizeof v, cudaMemcpyDeviceToHost));
    c->wt_num_observed = v[C_WT_OBS]; c->wt_num_mapped = v[C_WT_MAP]; c->wt_num_hits = v[C_WT_HITS]; c->wt_upper_bound = v[C_WT_UPPER];
    c->bsj_num_observed = v[C_BSJ_OBS]; c->bsj_num_mapped = v[C_BSJ_MAP]; c->bsj_num_hits = v[C_BSJ_HITS]; c->bsj_upper_bound = v[C_BSJ_UPPER];
    c->wt_read_len = s->wt_read_len; c->bsj_read_len = s->bsj_read_len;
    if (per_bsj && n_bsj) CG_CUDA(cudaMemcpy(per_bsj, s->d_acc, n_bsj * 8, cudaMemcpyDeviceToHost));
    return CG_OK;
}

int cg_session_accumulator(cg_session* s, void** dptr, uint64_t* n_u64) {
    if (!s || !dptr || !n_u64) CG_FAIL(CG_ERR_ARG, "null argument");
    *dptr = s->d_acc; *n_u64 = s->n_bsj + C_NUM;
    return CG_OK;
}

uint64_t cg_session_launches(const cg_session* s) { return s ? s->launches : 0; }

// ---- the one collective of the path (SURVEY.md section 8 e): all-reduce of the accumulators -------------------------------------
// NCCL is bound at first use (dlopen of libnccl.so.2), so the library loads in processes that bring their own NCCL (bench.py's ranks
// are torch.distributed ranks and reduce through torch) and on boxes without one; only this entry point needs it.
namespace {
typedef void* nccl_comm_t;
struct Nccl {
    void* h = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll"); CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce"); GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd"); GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd && GetErrorString;
    }
};
Nccl g_nccl;
__global__ void __launch_bounds__(256)
k_acc_add(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
#define CG_NCCL(expr) do { int _r = (expr); if (_r != 0) { cgi::set_error(std::string("NCCL error: ") + g_nccl.GetErrorString(_r) + " in " #expr); return CG_ERR_CUDA; } } while (0)
} // namespace

int cg_sessions_allreduce(cg_session* const* ss, int n) {
    if (!ss || n <= 0) CG_FAIL(CG_ERR_ARG, "no session given");
    const uint64_t len = ss[0]->n_bsj + C_NUM;
    for (int i = 0; i < n; ++i) {
        if (!ss[i]) CG_FAIL(CG_ERR_ARG, "null session");
        if (ss[i]->pending) CG_FAIL(CG_ERR_STATE, "fetch the pending batch of every session first");
        if (ss[i]->n_bsj + C_NUM != len) CG_FAIL(CG_ERR_ARG, "the sessions count over different BSJ references");
    }
    // per device: the first session given is the leader, the others of that device are added into it
    std::vector<int> devs; std::vector<cg_session*> lead;
    for (int i = 0; i < n; ++i) {
        size_t d = std::find(devs.begin(), devs.end(), ss[i]->device) - devs.begin();
        if (d == devs.size()) { devs.push_back(ss[i]->device); lead.push_back(ss[i]); continue; }
        CG_CUDA(cudaSetDevice(ss[i]->device));
        CG_CUDA(cudaStreamSynchronize(ss[i]->stream));
        k_acc_add<<<ss[i]->sms * 4, 256, 0, lead[d]->stream>>>(lead[d]->d_acc, ss[i]->d_acc, len);
        CG_CUDA(cudaGetLastError());
    }
    if (devs.size() > 1 || getenv("CG_NCCL_ALWAYS")) {
        if (!g_nccl.load()) CG_FAIL(CG_ERR_CUDA, "NCCL (libnccl.so.2) is not available: the all-reduce across GPUs needs it");
        std::vector<nccl_comm_t> comm(devs.size(), nullptr);
        CG_NCCL(g_nccl.CommInitAll(comm.data(), (int)devs.size(), devs.data()));
        CG_NCCL(g_nccl.GroupStart());
        for (size_t d = 0; d < devs.size(); ++d)        // in place, uint64 sum (ncclUint64 = 5, ncclSum = 0)
            CG_NCCL(g_nccl.AllReduce(lead[d]->d_acc, lead[d]->d_acc, len, 5, 0, comm[d], lead[d]->stream));
        CG_NCCL(g_nccl.GroupEnd());
        for (size_t d = 0; d < devs.size(); ++d) { CG_CUDA(cudaSetDevice(devs[d])); CG_CUDA(cudaStreamSynchronize(lead[d]->stream)); }
        for (size_t d = 0; d < devs.size(); ++d) g_nccl.CommDestroy(comm[d]);
    }
    // every session ends up holding the total
    for (int i = 0; i < n; ++i) {
        size_t d = std::find(devs.begin(), devs.end(), ss[i]->device) - devs.begin();
        if (lead[d] == ss[i]) continue;
        CG_CUDA(cudaSetDevice(ss[i]->device));
        CG_CUDA(cudaMemcpyAsync(ss[i]->d_acc, lead[d]->d_acc, len * 8, cudaMemcpyDeviceToDevice, lead[d]->stream));
    }
    for (size_t d = 0; d < devs.size(); ++d) { CG_CUDA(cudaSetDevice(devs[d])); CG_CUDA(cudaStreamSynchronize(lead[d]->stream)); }
    // read lengths are "the first one any worker saw": take the first non-zero
    uint32_t wl = 0, bl = 0;
    for (int i = 0; i < n; ++i) { if (!wl) wl = ss[i]->wt_read_len; if (!bl) bl = ss[i]->bsj_read_len; }
    for (int i = 0; i < n; ++i) { ss[i]->wt_read_len = wl; ss[i]->bsj_read_len = bl; }
    return CG_OK;
}

} // extern "C"
