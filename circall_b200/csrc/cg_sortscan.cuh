// circall_b200 — the sort and scan primitives of the index builder (cg_index.cu), hand-written for sm_100a: a stable least-significant-
// digit radix sort of (uint64 key, uint32 value) pairs, prefix scans (sum / max) of uint32 arrays of any length, and the ordered
// selection of the indices of set byte flags.  Round 1 used CUB for these (a third of the index build's device time); they are
// set-up, not the timed path, so the aim here is correctness at 10^9 elements and streaming-rate passes, not the last 30 %.
//
//   radix sort   8-bit digits, one pass = histogram kernel (one 256-bin histogram per 2048-pair tile, digit-major in global memory),
//                one exclusive scan over all (digit, tile) counts, one scatter kernel.  Stability inside a tile: every warp owns a
//                contiguous 256-pair slice and walks it 32 pairs at a time; __match_any_sync groups the lanes of a step by digit, the
//                rank of a pair is its warp's running count for the digit plus the number of lower lanes in its group.  The tile is put in
//                sorted order in shared memory first, so that the pairs of a digit leave as one contiguous run.
//   scans        three phases: per-chunk reduction, single-block scan of the chunk totals, per-chunk scan + carry-in.
//   select       flag count per tile, scan, ordered write (as cg_session.cu's compaction, over up to 2^32 flags).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cgs {

static const int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_PER_WARP = 256, RS_TILE = RS_WARPS * RS_PER_WARP;      // 2048 pairs per block
static const int SC_THREADS = 256, SC_ITEMS = 16, SC_CHUNK = SC_THREADS * SC_ITEMS;                                        // 4096 entries per block

struct OpSum { __device__ __forceinline__ static uint32_t id() { return 0u; } __device__ __forceinline__ static uint32_t ap(uint32_t a, uint32_t b) { return a + b; } };
struct OpMax { __device__ __forceinline__ static uint32_t id() { return 0u; } __device__ __forceinline__ static uint32_t ap(uint32_t a, uint32_t b) { return a > b ? a : b; } };

// inclusive scan of one value per thread over a block of SC_THREADS threads; returns the inclusive value, total in *tot
template <class Op>
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t* sh /* 32 */, uint32_t* tot) {
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (l >= d) x = Op::ap(y, x); }
    if (l == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = l < (int)(blockDim.x >> 5) ? sh[l] : Op::id();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, d); if (l >= d) s = Op::ap(y, s); }
        sh[l] = s;
    }
    __syncthreads();
    if (tot) *tot = sh[(blockDim.x >> 5) - 1];
    const uint32_t carry = w ? sh[w - 1] : Op::id();
    __syncthreads();
    return Op::ap(carry, x);
}

// phase 1: total of every chunk
template <class Op>
__global__ void __launch_bounds__(SC_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * SC_CHUNK + (uint64_t)threadIdx.x * SC_ITEMS;
    uint32_t a = Op::id();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) if (base + j < n) a = Op::ap(a, in[base + j]);
    uint32_t tot;
    block_scan<Op>(a, sh, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
// phase 2: exclusive scan of m chunk totals by ONE block (m up to a few million: 8192 entries per trip); grand total -> *total (may be null)
template <class Op>
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* sums, uint64_t m, uint32_t* total) {
    __shared__ uint32_t sh[32];
    __shared__ uint32_t incs[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = Op::id();
    __syncthreads();
    for (uint64_t base = 0; base < m; base += 1024 * 8) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * 8;
        uint32_t v[8], a = Op::id();
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = i0 + j < m ? sums[i0 + j] : Op::id(); a = Op::ap(a, v[j]); }
        uint32_t tot;
        const uint32_t inc = block_scan<Op>(a, sh, &tot);
        incs[threadIdx.x] = inc;                                 // exclusive prefix of a thread = inclusive value of the thread before it
        __syncthreads();
        const uint32_t ex = threadIdx.x ? incs[threadIdx.x - 1] : Op::id();
        uint32_t run = Op::ap(carry, ex);
#pragma unroll
        for (int j = 0; j < 8; ++j) { if (i0 + j < m) sums[i0 + j] = run; run = Op::ap(run, v[j]); }
        __syncthreads();
        if (threadIdx.x == 0) carry = Op::ap(carry, tot);
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
// phase 3: every chunk scanned on its own, plus the carry-in of the chunks before it
template <class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(SC_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * SC_CHUNK + (uint64_t)threadIdx.x * SC_ITEMS;
    uint32_t v[SC_ITEMS], a = Op::id();
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) { v[j] = base + j < n ? in[base + j] : Op::id(); a = Op::ap(a, v[j]); }
    const uint32_t inc = block_scan<Op>(a, sh, nullptr);
    // exclusive prefix of the thread: inclusive value of the thread before it
    __shared__ uint32_t incs[SC_THREADS];
    incs[threadIdx.x] = inc;
    __syncthreads();
    uint32_t run = Op::ap(sums[blockIdx.x], threadIdx.x ? incs[threadIdx.x - 1] : Op::id());
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) {
        if (INCLUSIVE) { run = Op::ap(run, v[j]); if (base + j < n) out[base + j] = run; }
        else { if (base + j < n) out[base + j] = run; run = Op::ap(run, v[j]); }
    }
}

// scan of in[0..n) into out (may alias in); tmp: at least (n / SC_CHUNK + 1) uint32; total (device, may be null) receives the grand total
template <class Op, bool INCLUSIVE>
inline cudaError_t scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* tmp, uint32_t* total, cudaStream_t st = 0) {
    if (n == 0) return cudaSuccess;
    const uint64_t nc = (n + SC_CHUNK - 1) / SC_CHUNK;
    k_scan_reduce<Op><<<(unsigned)nc, SC_THREADS, 0, st>>>(in, n, tmp);
    k_scan_sums<Op><<<1, 1024, 0, st>>>(tmp, nc, total);
    k_scan_apply<Op, INCLUSIVE><<<(unsigned)nc, SC_THREADS, 0, st>>>(in, out, n, tmp);
    return cudaGetLastError();
}
inline uint64_t scan_tmp_entries(uint64_t n) { return n / SC_CHUNK + 2; }

// ---- ordered selection: out[] = indices i (ascending) with flags[i] != 0; count -> *n_sel (device uint64) -------------------
__global__ void __launch_bounds__(SC_THREADS) k_sel_count(const uint8_t* __restrict__ f, uint64_t n, uint32_t* __restrict__ cnt) {
    __shared__ uint32_t sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * SC_CHUNK + (uint64_t)threadIdx.x * SC_ITEMS;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) if (base + j < n) c += f[base + j] ? 1u : 0u;
    uint32_t tot;
    block_scan<OpSum>(c, sh, &tot);
    if (threadIdx.x == 0) cnt[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SC_THREADS) k_sel_write(const uint8_t* __restrict__ f, uint64_t n, const uint32_t* __restrict__ cnt, uint32_t* __restrict__ out,
                                                         const uint32_t* __restrict__ total32, unsigned long long* __restrict__ n_sel) {
    __shared__ uint32_t sh[32];
    __shared__ uint32_t incs[SC_THREADS];
    const uint64_t base = (uint64_t)blockIdx.x * SC_CHUNK + (uint64_t)threadIdx.x * SC_ITEMS;
    uint8_t v[SC_ITEMS]; uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) { v[j] = base + j < n ? f[base + j] : 0; c += v[j] ? 1u : 0u; }
    const uint32_t inc = block_scan<OpSum>(c, sh, nullptr);
    incs[threadIdx.x] = inc;
    __syncthreads();
    uint32_t o = cnt[blockIdx.x] + (threadIdx.x ? incs[threadIdx.x - 1] : 0u);
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) if (v[j]) out[o++] = (uint32_t)(base + j);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_sel) *n_sel = (unsigned long long)*total32;
}
// tmp: scan_tmp_entries(n) + 1 uint32
inline cudaError_t select_flagged(const uint8_t* flags, uint64_t n, uint32_t* out, unsigned long long* n_sel, uint32_t* tmp, cudaStream_t st = 0) {
    if (n == 0) return cudaMemsetAsync(n_sel, 0, 8, st);
    const uint64_t nc = (n + SC_CHUNK - 1) / SC_CHUNK;
    uint32_t* total32 = tmp + nc;
    k_sel_count<<<(unsigned)nc, SC_THREADS, 0, st>>>(flags, n, tmp);
    k_scan_sums<OpSum><<<1, 1024, 0, st>>>(tmp, nc, total32);
    k_sel_write<<<(unsigned)nc, SC_THREADS, 0, st>>>(flags, n, tmp, out, total32, n_sel);
    return cudaGetLastError();
}

// ---- radix sort of (uint64, uint32) pairs ---------------------------------------------------------------------------------------
// one pass: digit = (key >> shift) & mask (mask = 2^bits - 1, bits <= 8)
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t mask, uint32_t* __restrict__ hist, uint32_t nb) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int j = 0; j < RS_TILE / RS_THREADS; ++j) {
        const uint64_t i = base + (uint64_t)j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];       // digit-major: the scan of the whole array is the global offset of (digit, tile)
}
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, int shift, uint32_t mask,
                                                          const uint32_t* __restrict__ offs, uint32_t nb, uint64_t* __restrict__ okeys, uint32_t* __restrict__ ovals) {
    __shared__ uint32_t cnt[RS_WARPS][256];          // first the slice histograms, then the running local cursors of every warp
    __shared__ uint32_t gbase[256];                  // global position of a pair = gbase[digit] + its position in the tile's sorted order
    __shared__ uint32_t sh[32];
    __shared__ uint64_t skeys[RS_TILE];              // the tile in sorted order: the pairs of a digit leave as one contiguous run
    __shared__ uint32_t svals[RS_TILE];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t tbase = (uint64_t)blockIdx.x * RS_TILE, wbase = tbase + (uint64_t)w * RS_PER_WARP;
    // phase A: histogram of the warp's slice
    for (int c = 0; c < RS_PER_WARP / 32; ++c) {
        const uint64_t i = wbase + (uint64_t)c * 32 + lane;
        const bool in = i < n;
        const uint32_t d = in ? ((uint32_t)(keys[i] >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (in && lane == __ffs(peers) - 1) cnt[w][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase B: where the digit's run starts inside the tile (scan over the digits), the local cursor of every (warp, digit), and the
    // global base of the run
    {
        const int d = threadIdx.x;
        uint32_t tot = 0;
#pragma unroll
        for (int x = 0; x < RS_WARPS; ++x) tot += cnt[x][d];
        const uint32_t lstart = block_scan<OpSum>(tot, sh, nullptr) - tot;
        uint32_t run = lstart;
#pragma unroll
        for (int x = 0; x < RS_WARPS; ++x) { const uint32_t t = cnt[x][d]; cnt[x][d] = run; run += t; }
        gbase[d] = offs[(size_t)d * nb + blockIdx.x] - lstart;
    }
    __syncthreads();
    // phase C: the slice again, in order, into the tile's sorted order
    for (int c = 0; c < RS_PER_WARP / 32; ++c) {
        const uint64_t i = wbase + (uint64_t)c * 32 + lane;
        const bool in = i < n;
        uint64_t key = 0; uint32_t val = 0;
        if (in) { key = keys[i]; val = vals[i]; }
        const uint32_t d = in ? ((uint32_t)(key >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t pos = 0;
        if (in) pos = cnt[w][d] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (in && lane == __ffs(peers) - 1) cnt[w][d] += __popc(peers);
        __syncwarp();
        if (in) { skeys[pos] = key; svals[pos] = val; }
    }
    __syncthreads();
    // phase D: out, run by run
    const uint32_t have = (uint32_t)(n - tbase < (uint64_t)RS_TILE ? n - tbase : (uint64_t)RS_TILE);
    for (uint32_t i = threadIdx.x; i < have; i += RS_THREADS) {
        const uint64_t key = skeys[i];
        const uint32_t g = gbase[(uint32_t)(key >> shift) & mask] + i;
        okeys[g] = key; ovals[g] = svals[i];
    }
}

struct SortBufs { uint64_t* k[2]; uint32_t* v[2]; int cur; };       // the pairs live in k[cur], v[cur]
inline uint64_t sort_tmp_entries(uint64_t n) { const uint64_t nb = (n + RS_TILE - 1) / RS_TILE; return 256 * nb + scan_tmp_entries(256 * nb) + 8; }
// stable sort of the n pairs by key bits [begin_bit, end_bit); tmp: sort_tmp_entries(n) uint32
inline cudaError_t sort_pairs(SortBufs& b, uint64_t n, int begin_bit, int end_bit, uint32_t* tmp, cudaStream_t st = 0) {
    if (n == 0) return cudaSuccess;
    const uint64_t nb = (n + RS_TILE - 1) / RS_TILE;
    uint32_t* hist = tmp;
    uint32_t* stmp = tmp + 256 * nb;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        const int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        const uint32_t mask = (1u << bits) - 1u;
        k_rs_hist<<<(unsigned)nb, RS_THREADS, 0, st>>>(b.k[b.cur], n, shift, mask, hist, (uint32_t)nb);
        cudaError_t e = scan_u32<OpSum, false>(hist, hist, 256 * nb, stmp, nullptr, st);
        if (e != cudaSuccess) return e;
        k_rs_scatter<<<(unsigned)nb, RS_THREADS, 0, st>>>(b.k[b.cur], b.v[b.cur], n, shift, mask, hist, (uint32_t)nb, b.k[b.cur ^ 1], b.v[b.cur ^ 1]);
        b.cur ^= 1;
    }
    return cudaGetLastError();
}

} // namespace cgs
