// circall_b200 — quasi-index construction on the device and its directory I/O.
//
// Replaces (paths relative to /root/reference/Circall_v1.0.1/):
//   cg_index_build[_fasta]  SailfishIndex::build -> RapMap indexTranscriptsSA + buildHash [U]
//                           behind src/TxIndexer.cpp:202-227
//   cg_index_open / _save   SailfishIndex::load, include/ReadExperiment.hpp:66-67
//
// Device data produced (consumed by cg_core.cuh):
//   text    2 x uint4 per 64-base block: {c0, c1} 2-bit codes, {'$' mask, tid0, start0}
//   SA      uint32 suffix array of the whole '$'-joined text ('$' < A < C < G < T, suffixes compare
//           across '$' exactly like libdivsufsort on the raw bytes) — built by prefix doubling:
//           one 64-bit radix sort on the first 21 characters, then rounds of
//           (rank[i], rank[i+h]) sorts with h doubling until every suffix has its own rank
//   khash   open addressing, 16-B slots {k-mer word, begin, end}: one slot per run of SA entries
//           whose first k characters are equal and '$'-free
// The sorts and scans here use CUB (index construction is set-up, not the timed hot path).
#include "cg_common.h"
#include <cub/cub.cuh>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sys/stat.h>
#include <cerrno>
#include <stdexcept>

using namespace cg;
using cgi::DevBuf;

namespace {

__global__ void k_text_blocks(const unsigned char* __restrict__ ascii, uint64_t n, uint64_t nb, uint4* text,
                              uint32_t* cnt, int* bad) {
    uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    unsigned char buf[64];
    int64_t rem = (int64_t)n - (int64_t)(b * 64);
    int nv = rem >= 64 ? 64 : (rem > 0 ? (int)rem : 0);
    if (nv == 64) {
        const uint4* p = (const uint4*)(ascii + b * 64);
        uint4* q = (uint4*)buf;
        q[0] = __ldg(p); q[1] = __ldg(p + 1); q[2] = __ldg(p + 2); q[3] = __ldg(p + 3);
    } else {
        for (int i = 0; i < 64; ++i) buf[i] = i < nv ? ascii[b * 64 + i] : (unsigned char)'$';
    }
    uint64_t c0, c1, mask; bool isbad;
    make_block(buf, nv, c0, c1, mask, isbad);
    if (isbad) atomicExch(bad, 1);
    text[2 * b] = make_uint4((uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32));
    text[2 * b + 1] = make_uint4((uint32_t)mask, (uint32_t)(mask >> 32), 0u, 0u);
    uint64_t real = nv >= 64 ? mask : (nv > 0 ? (mask & ((1ULL << nv) - 1)) : 0ULL);
    cnt[b] = (uint32_t)__popcll(real);
}

__global__ void k_text_finish(uint4* text, const uint32_t* tid0, const uint32_t* offsets, uint32_t n_tx, uint64_t nb) {
    uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t t = tid0[b];
    uint4 v = text[2 * b + 1];
    v.z = t; v.w = offsets[t < n_tx ? t : n_tx];
    text[2 * b + 1] = v;
}

// 21 characters x 3 bits: pad 0 < '$' 1 < A 2 < C 3 < G 4 < T 5
__global__ void k_sa_init(const unsigned char* __restrict__ ascii, uint64_t n, uint64_t* keys, uint32_t* sa) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0;
#pragma unroll
    for (int j = 0; j < 21; ++j) {
        uint64_t c = 0;
        if (i + j < n) {
            int bc = base_class(ascii[i + j]);
            c = bc < 4 ? (uint64_t)(2 + bc) : 1;
        }
        key = (key << 3) | c;
    }
    keys[i] = key;
    sa[i] = (uint32_t)i;
}

// head[j] = j when sorted key j opens a new group, else 0; counts the groups
__global__ void k_sa_heads(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* head, unsigned long long* ngroups) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool f = false;
    if (j < n) {
        f = (j == 0) || (keys[j] != keys[j - 1]);
        head[j] = f ? (uint32_t)j : 0u;
    }
    unsigned m = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(ngroups, (unsigned long long)__popc(m));
}

__global__ void k_sa_scatter_rank(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ head, uint64_t n, uint32_t* rank) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j < n) rank[sa[j]] = head[j];
}

__global__ void k_sa_keys(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ rank, uint64_t n, uint64_t h, uint64_t* keys) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t s = sa[j];
    uint64_t hi = rank[s];
    uint64_t lo = (s + h < n) ? (uint64_t)rank[s + h] + 1 : 0;
    keys[j] = (hi << 32) | lo;
}

__global__ void k_kmers(IndexView ix, uint64_t n, uint64_t* km) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t w;
    km[j] = text_kmer(ix, ix.sa[j], w) ? w : ~0ULL;
}

__global__ void k_boundary(const uint64_t* __restrict__ km, uint64_t n, uint8_t* bnd) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t w = km[j];
    bool same = (w != ~0ULL) && j > 0 && km[j - 1] == w;
    bnd[j] = same ? 0 : 1;
}

__global__ void k_count_heads(const uint32_t* __restrict__ B, uint64_t nB, const uint64_t* __restrict__ km, unsigned long long* nk) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    bool f = i < nB && km[B[i]] != ~0ULL;
    unsigned m = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nk, (unsigned long long)__popc(m));
}

__global__ void k_insert(const uint32_t* __restrict__ B, uint64_t nB, const uint64_t* __restrict__ km, uint64_t n, uint4* slots, uint64_t hmask, int hshift) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nB) return;
    uint32_t j = B[i];
    uint64_t key = km[j];
    if (key == ~0ULL) return;
    uint32_t e = (i + 1 < nB) ? B[i + 1] : (uint32_t)n;
    uint64_t h = home_of(key, hshift);
    for (;;) {
        unsigned long long* kp = (unsigned long long*)&slots[h];
        unsigned long long old = atomicCAS(kp, ~0ULL, (unsigned long long)key);
        if (old == ~0ULL) {
            uint32_t* be = (uint32_t*)kp;
            be[2] = j; be[3] = e;
            return;
        }
        h = (h + 1) & hmask;
    }
}

inline unsigned grid_for(uint64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

int bits_for(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

// suffix array of ascii[0,n) into *out_sa (device, n uint32; caller frees)
int build_sa(const unsigned char* d_ascii, uint64_t n, uint32_t** out_sa) {
    DevBuf kA, kB, sA, sB, rk, tmp, cnt;
    CG_CUDA(kA.alloc(n * 8)); CG_CUDA(kB.alloc(n * 8));
    CG_CUDA(sA.alloc(n * 4)); CG_CUDA(sB.alloc(n * 4));
    CG_CUDA(rk.alloc(n * 4)); CG_CUDA(cnt.alloc(8));
    cub::DoubleBuffer<uint64_t> keys(kA.as<uint64_t>(), kB.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> vals(sA.as<uint32_t>(), sB.as<uint32_t>());
    size_t tb_sort = 0, tb_scan = 0;
    CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, keys, vals, (int64_t)n, 0, 64));
    CG_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb_scan, (uint32_t*)nullptr, (uint32_t*)nullptr, cub::Max(), (int64_t)n));
    CG_CUDA(tmp.alloc(tb_sort > tb_scan ? tb_sort : tb_scan));
    const int BS = 256;
    k_sa_init<<<grid_for(n, BS), BS>>>(d_ascii, n, keys.Current(), vals.Current());
    CG_CUDA(cudaGetLastError());
    size_t tb = tmp.bytes;
    CG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys, vals, (int64_t)n, 0, 63));
    const int rbits = bits_for(n + 1);
    uint64_t h = 21;
    for (int round = 0; round < 64; ++round) {
        uint32_t* head = (uint32_t*)keys.Alternate();   // free scratch until the next sort
        CG_CUDA(cudaMemset(cnt.p, 0, 8));
        k_sa_heads<<<grid_for(n, BS), BS>>>(keys.Current(), n, head, cnt.as<unsigned long long>());
        CG_CUDA(cudaGetLastError());
        unsigned long long ngroups = 0;
        CG_CUDA(cudaMemcpy(&ngroups, cnt.p, 8, cudaMemcpyDeviceToHost));
        if (ngroups == n) break;
        if (h >= n) CG_FAIL(CG_ERR_STATE, "suffix array construction did not converge");
        tb = tmp.bytes;
        CG_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, head, head, cub::Max(), (int64_t)n));
        k_sa_scatter_rank<<<grid_for(n, BS), BS>>>(vals.Current(), head, n, rk.as<uint32_t>());
        CG_CUDA(cudaGetLastError());
        k_sa_keys<<<grid_for(n, BS), BS>>>(vals.Current(), rk.as<uint32_t>(), n, h, keys.Current());
        CG_CUDA(cudaGetLastError());
        tb = tmp.bytes;
        CG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys, vals, (int64_t)n, 0, 32 + rbits));
        h *= 2;
    }
    CG_CUDA(cudaDeviceSynchronize());
    // hand the buffer that holds the result to the caller
    if (vals.Current() == sA.as<uint32_t>()) { *out_sa = sA.as<uint32_t>(); sA.p = nullptr; }
    else { *out_sa = sB.as<uint32_t>(); sB.p = nullptr; }
    return CG_OK;
}

int build_khash(cg_index* idx) {
    const uint64_t n = idx->n;
    const int BS = 256;
    DevBuf km, bnd, B, nsel, tmp, cnt;
    CG_CUDA(km.alloc(n * 8)); CG_CUDA(bnd.alloc(n)); CG_CUDA(B.alloc(n * 4)); CG_CUDA(nsel.alloc(8)); CG_CUDA(cnt.alloc(8));
    IndexView v = idx->view();
    k_kmers<<<grid_for(n, BS), BS>>>(v, n, km.as<uint64_t>());
    CG_CUDA(cudaGetLastError());
    k_boundary<<<grid_for(n, BS), BS>>>(km.as<uint64_t>(), n, bnd.as<uint8_t>());
    CG_CUDA(cudaGetLastError());
    size_t tb = 0;
    cub::CountingInputIterator<uint32_t> it(0);
    CG_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, it, bnd.as<uint8_t>(), B.as<uint32_t>(), nsel.as<uint64_t>(), (int64_t)n));
    CG_CUDA(tmp.alloc(tb));
    CG_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, it, bnd.as<uint8_t>(), B.as<uint32_t>(), nsel.as<uint64_t>(), (int64_t)n));
    uint64_t nB = 0;
    CG_CUDA(cudaMemcpy(&nB, nsel.p, 8, cudaMemcpyDeviceToHost));
    CG_CUDA(cudaMemset(cnt.p, 0, 8));
    if (nB) k_count_heads<<<grid_for(nB, BS), BS>>>(B.as<uint32_t>(), nB, km.as<uint64_t>(), cnt.as<unsigned long long>());
    CG_CUDA(cudaGetLastError());
    unsigned long long nk = 0;
    CG_CUDA(cudaMemcpy(&nk, cnt.p, 8, cudaMemcpyDeviceToHost));
#ifndef CG_TABLE_SLOTS_PER_KMER
#define CG_TABLE_SLOTS_PER_KMER 8       // measured: 4 -> 21.98 ms per 4 M pairs, 8 -> 20.86 (fewer collided home slots, so fewer second and third
                                        // resolution passes); costs 26 GB more of the 180 GB for config 2
#endif
    // load factor <= 0.125: unsuccessful look-ups (the bulk of the mapper's probes) stop after ~1.3 slots
    uint64_t cap = 16;
    int hshift = 60;
    while (cap < nk * CG_TABLE_SLOTS_PER_KMER + 2) { cap <<= 1; --hshift; }
    bnd.alloc(0); tmp.alloc(0);
    if (cap > (1ULL << 33)) CG_FAIL(CG_ERR_CAPACITY, "k-mer table would need more than 2^33 slots");
    CG_CUDA(cudaMalloc((void**)&idx->d_hslots, cap * 16));
    CG_CUDA(cudaMemset(idx->d_hslots, 0xFF, cap * 16));
    if (nB) k_insert<<<grid_for(nB, BS), BS>>>(B.as<uint32_t>(), nB, km.as<uint64_t>(), n, idx->d_hslots, cap - 1, hshift);
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaDeviceSynchronize());
    idx->hcap = cap; idx->n_kmers = nk;
    return CG_OK;
}

// seed filter (cg_core.cuh: IndexView::fbits): one thread per text position
__global__ void __launch_bounds__(256)
k_fbits(IndexView ix, int fm, uint32_t* __restrict__ bits) {
    const uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p >= ix.n) return;
    uint64_t w;
    if (!text_word(ix, (uint32_t)p, fm, w)) return;
    const uint32_t x = (uint32_t)w, y = fm_revcomp(x, fm);
    if (!((bits[x >> 5] >> (x & 31)) & 1u)) atomicOr(&bits[x >> 5], 1u << (x & 31));
    if (!((bits[y >> 5] >> (y & 31)) & 1u)) atomicOr(&bits[y >> 5], 1u << (y & 31));
}
// transcriptAtPosition(SA[i]) for every i (IndexView::satid)
__global__ void __launch_bounds__(256)
k_satid(IndexView ix, uint32_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= ix.n) return;
    uint32_t tid, local;
    tid_of(ix, ix.sa[i], tid, local);
    out[i] = tid;
}
int build_satid(cg_index* idx) {
    if (getenv("CG_NO_SATID")) return CG_OK;
    IndexView v = idx->view();
    CG_CUDA(cudaMalloc((void**)&idx->d_satid, idx->n * 4));
    k_satid<<<grid_for(idx->n, 256), 256>>>(v, idx->d_satid);
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaDeviceSynchronize());
    return CG_OK;
}

// min(255, lcp(SA[i - 1], SA[i])) for every i (IndexView::lcp)
__global__ void __launch_bounds__(256)
k_lcp(IndexView ix, uint8_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= ix.n) return;
    out[i] = i == 0 ? (uint8_t)0 : (uint8_t)lce(ix, (uint32_t)i - 1, (uint32_t)i, 0, 255, 1);
}
int build_lcp(cg_index* idx) {
    if (getenv("CG_NO_LCP")) return CG_OK;
    CG_CUDA(cudaMalloc((void**)&idx->d_lcp, idx->n));
    IndexView v = idx->view();
    v.lcp = nullptr;
    k_lcp<<<grid_for(idx->n, 256), 256>>>(v, idx->d_lcp);
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaDeviceSynchronize());
    return CG_OK;
}

int build_fbits(cg_index* idx) {
    int want = 16;
    if (const char* e = getenv("CG_FM")) want = atoi(e);
    idx->fm = fm_for_k(idx->k, want);
    if (!idx->fm) return CG_OK;
    const uint64_t words = (1ULL << (2 * idx->fm)) / 32;
    CG_CUDA(cudaMalloc((void**)&idx->d_fbits, words * 4));
    CG_CUDA(cudaMemset(idx->d_fbits, 0, words * 4));
    IndexView v = idx->view();
    k_fbits<<<grid_for(idx->n, 256), 256>>>(v, idx->fm, idx->d_fbits);
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cudaDeviceSynchronize());
    return CG_OK;
}

int check_device(int device) {
    int nd = 0;
    cudaError_t e = cudaGetDeviceCount(&nd);
    if (e != cudaSuccess || nd == 0) { cudaGetLastError(); CG_FAIL(CG_ERR_CUDA, "no CUDA device available (libcircall_gpu has no CPU fallback)"); }
    if (device < 0 || device >= nd) CG_FAIL(CG_ERR_ARG, "device ordinal out of range");
    CG_CUDA(cudaSetDevice(device));
    cgi::device_tuning(device);
    return CG_OK;
}

void free_index_device(cg_index* idx) {
    if (idx->d_text) cudaFree(idx->d_text);
    if (idx->d_sa) cudaFree(idx->d_sa);
    if (idx->d_hslots) cudaFree(idx->d_hslots);
    if (idx->d_offsets) cudaFree(idx->d_offsets);
    if (idx->d_fbits) cudaFree(idx->d_fbits);
    if (idx->d_satid) cudaFree(idx->d_satid);
    if (idx->d_lcp) cudaFree(idx->d_lcp);
    idx->d_fbits = nullptr; idx->d_satid = nullptr; idx->d_lcp = nullptr;
    idx->d_text = nullptr; idx->d_sa = nullptr; idx->d_hslots = nullptr; idx->d_offsets = nullptr;
}

void finish_sizes(cg_index* idx) {
    idx->device_bytes = idx->n_blocks * 32 + idx->n * 4 + idx->hcap * 16 + ((uint64_t)idx->n_tx + 1) * 4 + (idx->d_fbits ? (1ULL << (2 * idx->fm)) / 8 : 0) + (idx->d_satid ? idx->n * 4 : 0) + (idx->d_lcp ? idx->n : 0);
}

int build_impl(int device, const char* text, uint64_t n, int k, const std::vector<std::string>* names, cg_index** out) {
    if (!text || !out) CG_FAIL(CG_ERR_ARG, "null argument");
    if (n == 0 || text[n - 1] != '$') CG_FAIL(CG_ERR_ARG, "index text must be non-empty and end with '$'");
    if (k < 3 || k > 31 || (k & 1) == 0) CG_FAIL(CG_ERR_ARG, "k must be odd and <= 31 (TxIndexer.cpp:208-214)");
    if (n >= (1ULL << 31) - 128) CG_FAIL(CG_ERR_LENGTH, "text >= 2^31: the 64-bit index (RapMapSAIndex<int64_t>) is not implemented");
    int rc = check_device(device);
    if (rc) return rc;
    auto t0 = std::chrono::steady_clock::now();
    cg_index* idx = new cg_index();
    idx->device = device; idx->k = k; idx->n = n;
    {
        uint32_t start = 0;
        const char* p = text; const char* end = text + n;
        while (p < end) {
            const char* q = (const char*)memchr(p, '$', (size_t)(end - p));
            if (!q) break;
            idx->offsets.push_back(start);
            start = (uint32_t)(q - text) + 1;
            p = q + 1;
        }
        idx->n_tx = (uint32_t)idx->offsets.size();
        idx->offsets.push_back((uint32_t)n);
    }
    if (names) idx->names = *names;
    idx->names.resize(idx->n_tx);
    for (uint32_t t = 0; t < idx->n_tx; ++t)
        if (idx->names[t].empty()) idx->names[t] = "t" + std::to_string(t);
    idx->n_blocks = (n + 63) / 64 + 1;
    auto fail = [&](int code) { free_index_device(idx); delete idx; return code; };
#define CG_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(cgi::cuda_fail(_e, #expr, __FILE__, __LINE__)); } while (0)
    DevBuf ascii, cnt, bad;
    CG_TRY(ascii.alloc(n + 64));
    CG_TRY(cudaMemcpy(ascii.p, text, n, cudaMemcpyHostToDevice));
    CG_TRY(cudaMalloc((void**)&idx->d_text, idx->n_blocks * 32));
    CG_TRY(cudaMalloc((void**)&idx->d_offsets, ((uint64_t)idx->n_tx + 1) * 4));
    CG_TRY(cudaMemcpy(idx->d_offsets, idx->offsets.data(), ((uint64_t)idx->n_tx + 1) * 4, cudaMemcpyHostToDevice));
    CG_TRY(cnt.alloc(idx->n_blocks * 4));
    CG_TRY(bad.alloc(4));
    CG_TRY(cudaMemset(bad.p, 0, 4));
    const int BS = 256;
    k_text_blocks<<<grid_for(idx->n_blocks, BS), BS>>>(ascii.as<unsigned char>(), n, idx->n_blocks, idx->d_text, cnt.as<uint32_t>(), bad.as<int>());
    CG_TRY(cudaGetLastError());
    {
        size_t tb = 0;
        DevBuf tmp;
        CG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<uint32_t>(), cnt.as<uint32_t>(), (int64_t)idx->n_blocks));
        CG_TRY(tmp.alloc(tb));
        CG_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.as<uint32_t>(), cnt.as<uint32_t>(), (int64_t)idx->n_blocks));
    }
    k_text_finish<<<grid_for(idx->n_blocks, BS), BS>>>(idx->d_text, cnt.as<uint32_t>(), idx->d_offsets, idx->n_tx, idx->n_blocks);
    CG_TRY(cudaGetLastError());
    int hbad = 0;
    CG_TRY(cudaMemcpy(&hbad, bad.p, 4, cudaMemcpyDeviceToHost));
    if (hbad) { cgi::set_error("index text holds a character outside {A,C,G,T,$}"); return fail(CG_ERR_BASE); }
    cnt.alloc(0);
    rc = build_sa(ascii.as<unsigned char>(), n, &idx->d_sa);
    if (rc) return fail(rc);
    ascii.alloc(0);
    rc = build_khash(idx);
    if (rc) return fail(rc);
    rc = build_fbits(idx);
    if (rc) return fail(rc);
    rc = build_satid(idx);
    if (rc) return fail(rc);
    rc = build_lcp(idx);
    if (rc) return fail(rc);
#undef CG_TRY
    finish_sizes(idx);
    idx->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *out = idx;
    return CG_OK;
}

// ---- directory layout (this library's own) ----
//   header.json    human-readable summary; its presence marks a finished index (TxIndexer.cpp:192-193)
//   meta.bin       magic "CGIDX001", k, n, n_tx, n_kmers, hcap, n_blocks  (uint64 each)
//   offsets.bin    (n_tx + 1) uint32      names.txt   one transcript name per line
//   text.bin       n_blocks x 32 B        sa.bin      n uint32         hash.bin    hcap x 16 B
bool write_file(const std::string& path, const void* p, uint64_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = bytes == 0 || fwrite(p, 1, bytes, f) == bytes;
    return fclose(f) == 0 && ok;
}
bool read_file(const std::string& path, void* p, uint64_t bytes) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = bytes == 0 || fread(p, 1, bytes, f) == bytes;
    fclose(f);
    return ok;
}
int save_dev(const std::string& path, const void* d, uint64_t bytes) {
    std::vector<char> h(bytes);
    CG_CUDA(cudaMemcpy(h.data(), d, bytes, cudaMemcpyDeviceToHost));
    if (!write_file(path, h.data(), bytes)) CG_FAIL(CG_ERR_IO, "cannot write " + path);
    return CG_OK;
}
int load_dev(const std::string& path, void** d, uint64_t bytes) {
    std::vector<char> h(bytes);
    if (!read_file(path, h.data(), bytes)) CG_FAIL(CG_ERR_IO, "cannot read " + path);
    CG_CUDA(cudaMalloc(d, bytes ? bytes : 1));
    CG_CUDA(cudaMemcpy(*d, h.data(), bytes, cudaMemcpyHostToDevice));
    return CG_OK;
}

} // namespace

extern "C" {

int cg_index_build(int device, const char* text, uint64_t n, int k, cg_index** out) {
    try { return build_impl(device, text, n, k, nullptr, out); }
    catch (const std::exception& e) { CG_FAIL(CG_ERR_IO, std::string("index build failed: ") + e.what()); }
    catch (...) { CG_FAIL(CG_ERR_IO, "index build failed"); }
}

static int index_build_fasta_impl(int device, const char* fasta_path, int k, cg_index** out) {
    std::ifstream in(fasta_path);
    if (!in) CG_FAIL(CG_ERR_IO, std::string("cannot open ") + fasta_path);
    // RapMap record normalisation: name up to the first whitespace, upper case, non-ACGT ->
    // pseudo-random base from a fixed-seed engine, trailing poly-A (>= 10) clipped, short records dropped
    std::vector<std::string> names;
    std::string text, line, name, seq;
    std::mt19937 eng(271828);
    std::uniform_int_distribution<> dis(0, 3);
    const char bases[4] = {'A', 'C', 'G', 'T'};
    bool have = false;
    auto flush = [&]() {
        if (!have) return;
        for (auto& c : seq) {
            c = (char)::toupper((unsigned char)c);
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') c = bases[dis(eng)];
        }
        const size_t clip = 10;
        if (seq.size() > clip && seq.compare(seq.size() - clip, clip, std::string(clip, 'A')) == 0) {
            size_t e = seq.find_last_not_of('A');
            if (e == std::string::npos) seq.clear(); else seq.resize(e + 1);
        }
        if (seq.size() >= (size_t)k) { names.push_back(name); text += seq; text += '$'; }
        seq.clear();
    };
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            flush();
            have = true;
            size_t e = line.find_first_of(" \t", 1);
            name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
        } else if (have) {
            seq += line;
        }
    }
    flush();
    if (text.empty()) CG_FAIL(CG_ERR_IO, "no usable record in FASTA file");
    return build_impl(device, text.data(), text.size(), k, &names, out);
}
int cg_index_build_fasta(int device, const char* fasta_path, int k, cg_index** out) {
    if (!fasta_path || !out) CG_FAIL(CG_ERR_ARG, "null argument");
    try { return index_build_fasta_impl(device, fasta_path, k, out); }
    catch (const std::exception& e) { CG_FAIL(CG_ERR_IO, std::string("index build failed: ") + e.what()); }
    catch (...) { CG_FAIL(CG_ERR_IO, "index build failed"); }
}

static int index_save_impl(const cg_index* idx, const char* dir);
int cg_index_save(const cg_index* idx, const char* dir) {
    if (!idx || !dir) CG_FAIL(CG_ERR_ARG, "null argument");
    try { return index_save_impl(idx, dir); }
    catch (const std::exception& e) { CG_FAIL(CG_ERR_IO, std::string("index save failed: ") + e.what()); }
    catch (...) { CG_FAIL(CG_ERR_IO, "index save failed"); }
}
static int index_save_impl(const cg_index* idx, const char* dir) {
    CG_CUDA(cudaSetDevice(idx->device));
    std::string d(dir);
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) CG_FAIL(CG_ERR_IO, "cannot create directory " + d);
    uint64_t meta[8] = {0x3130305844494743ULL /* "CGIDX001" */, (uint64_t)idx->k, idx->n, idx->n_tx, idx->n_kmers, idx->hcap, idx->n_blocks, 0};
    if (!write_file(d + "/meta.bin", meta, sizeof meta)) CG_FAIL(CG_ERR_IO, "cannot write " + d + "/meta.bin");
    if (!write_file(d + "/offsets.bin", idx->offsets.data(), idx->offsets.size() * 4)) CG_FAIL(CG_ERR_IO, "cannot write offsets.bin");
    {
        std::string nm;
        for (auto& s : idx->names) { nm += s; nm += '\n'; }
        if (!write_file(d + "/names.txt", nm.data(), nm.size())) CG_FAIL(CG_ERR_IO, "cannot write names.txt");
    }
    int rc;
    if ((rc = save_dev(d + "/text.bin", idx->d_text, idx->n_blocks * 32))) return rc;
    if ((rc = save_dev(d + "/sa.bin", idx->d_sa, idx->n * 4))) return rc;
    if ((rc = save_dev(d + "/hash.bin", idx->d_hslots, idx->hcap * 16))) return rc;
    char js[512];
    snprintf(js, sizeof js,
             "{\n  \"IndexType\": \"circall_b200 quasi-index\",\n  \"IndexVersion\": \"CGIDX001\",\n  \"usesKmers\": true,\n"
             "  \"kmerLen\": %d,\n  \"bigSA\": false,\n  \"perfectHash\": false,\n  \"textLen\": %llu,\n  \"numTranscripts\": %u,\n"
             "  \"numKmers\": %llu\n}\n",
             idx->k, (unsigned long long)idx->n, idx->n_tx, (unsigned long long)idx->n_kmers);
    if (!write_file(d + "/header.json", js, strlen(js))) CG_FAIL(CG_ERR_IO, "cannot write header.json");
    return CG_OK;
}

static bool file_size_is(const std::string& path, uint64_t bytes) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 && (uint64_t)st.st_size == bytes;
}

static int index_open_impl(const char* dir, int device, cg_index** out) {
    std::string d(dir);
    uint64_t meta[8];
    if (!file_size_is(d + "/meta.bin", sizeof meta) || !read_file(d + "/meta.bin", meta, sizeof meta) || meta[0] != 0x3130305844494743ULL)
        CG_FAIL(CG_ERR_IO, "not a circall_b200 index directory: " + d);
    // nothing in meta.bin is trusted: every field is checked against the others and against the files' sizes before anything is
    // allocated or indexed with it
    const uint64_t k = meta[1], n = meta[2], n_tx = meta[3], n_kmers = meta[4], hcap = meta[5], n_blocks = meta[6];
    if (k < 3 || k > 31 || (k & 1) == 0) CG_FAIL(CG_ERR_IO, "meta.bin: k must be odd and <= 31");
    if (n == 0 || n >= (1ULL << 31) - 128) CG_FAIL(CG_ERR_IO, "meta.bin: text length out of range");
    if (n_tx == 0 || n_tx > n) CG_FAIL(CG_ERR_IO, "meta.bin: transcript count out of range");
    if (n_blocks != (n + 63) / 64 + 1) CG_FAIL(CG_ERR_IO, "meta.bin: text block count does not match the text length");
    if (hcap < 16 || (hcap & (hcap - 1)) != 0 || hcap > (1ULL << 33) || n_kmers >= hcap) CG_FAIL(CG_ERR_IO, "meta.bin: table capacity must be a power of two above the k-mer count");
    if (!file_size_is(d + "/offsets.bin", (n_tx + 1) * 4)) CG_FAIL(CG_ERR_IO, "offsets.bin has the wrong size");
    if (!file_size_is(d + "/text.bin", n_blocks * 32)) CG_FAIL(CG_ERR_IO, "text.bin has the wrong size");
    if (!file_size_is(d + "/sa.bin", n * 4)) CG_FAIL(CG_ERR_IO, "sa.bin has the wrong size");
    if (!file_size_is(d + "/hash.bin", hcap * 16)) CG_FAIL(CG_ERR_IO, "hash.bin has the wrong size");
    int rc = check_device(device);
    if (rc) return rc;
    cg_index* idx = new cg_index();
    idx->device = device; idx->k = (int)k; idx->n = n; idx->n_tx = (uint32_t)n_tx;
    idx->n_kmers = n_kmers; idx->hcap = hcap; idx->n_blocks = n_blocks;
    auto fail = [&](int code) { free_index_device(idx); delete idx; return code; };
    idx->offsets.resize((size_t)idx->n_tx + 1);
    if (!read_file(d + "/offsets.bin", idx->offsets.data(), idx->offsets.size() * 4)) { cgi::set_error("cannot read offsets.bin"); return fail(CG_ERR_IO); }
    if (idx->offsets[0] != 0 || idx->offsets[idx->n_tx] != n) { cgi::set_error("offsets.bin: first / last offset do not match the text"); return fail(CG_ERR_IO); }
    for (uint32_t t = 0; t < idx->n_tx; ++t)
        if (idx->offsets[t + 1] <= idx->offsets[t]) { cgi::set_error("offsets.bin: offsets must increase"); return fail(CG_ERR_IO); }
    {
        std::ifstream in(d + "/names.txt");
        std::string s;
        while (std::getline(in, s)) idx->names.push_back(s);
        if (idx->names.size() != idx->n_tx) { cgi::set_error("names.txt does not match the transcript count"); return fail(CG_ERR_IO); }
    }
    if ((rc = load_dev(d + "/text.bin", (void**)&idx->d_text, idx->n_blocks * 32))) return fail(rc);
    if ((rc = load_dev(d + "/sa.bin", (void**)&idx->d_sa, idx->n * 4))) return fail(rc);
    if ((rc = load_dev(d + "/hash.bin", (void**)&idx->d_hslots, idx->hcap * 16))) return fail(rc);
    cudaError_t e = cudaMalloc((void**)&idx->d_offsets, idx->offsets.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(idx->d_offsets, idx->offsets.data(), idx->offsets.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(cgi::cuda_fail(e, "upload offsets", __FILE__, __LINE__));
    if ((rc = build_fbits(idx))) return fail(rc);          // the seed filter and the per-suffix transcript array are derived from
    if ((rc = build_satid(idx))) return fail(rc);          // text + SA: not part of the directory
    if ((rc = build_lcp(idx))) return fail(rc);
    finish_sizes(idx);
    *out = idx;
    return CG_OK;
}

int cg_index_open(const char* dir, int device, cg_index** out) {
    if (!dir || !out) CG_FAIL(CG_ERR_ARG, "null argument");
    try { return index_open_impl(dir, device, out); }
    catch (const std::exception& e) { CG_FAIL(CG_ERR_IO, std::string("index open failed: ") + e.what()); }
    catch (...) { CG_FAIL(CG_ERR_IO, "index open failed"); }
}

void cg_index_close(cg_index* idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    free_index_device(idx);
    delete idx;
}

int cg_index_get_info(const cg_index* idx, cg_index_info* info) {
    if (!idx || !info) CG_FAIL(CG_ERR_ARG, "null argument");
    info->text_len = idx->n; info->n_tx = idx->n_tx; info->n_kmers = idx->n_kmers; info->hash_capacity = idx->hcap;
    info->k = idx->k; info->device = idx->device; info->device_bytes = idx->device_bytes; info->build_seconds = idx->build_seconds;
    return CG_OK;
}

const char* cg_index_tx_name(const cg_index* idx, uint32_t t) { return (idx && t < idx->n_tx) ? idx->names[t].c_str() : nullptr; }
int cg_index_set_tx_name(cg_index* idx, uint32_t t, const char* name) {
    if (!idx || !name || t >= idx->n_tx) CG_FAIL(CG_ERR_ARG, "bad transcript id");
    idx->names[t] = name;
    return CG_OK;
}
uint32_t cg_index_tx_len(const cg_index* idx, uint32_t t) { return (idx && t < idx->n_tx) ? idx->offsets[t + 1] - idx->offsets[t] - 1 : 0; }
uint32_t cg_index_tx_offset(const cg_index* idx, uint32_t t) { return (idx && t < idx->n_tx) ? idx->offsets[t] : 0; }

int cg_index_copy_sa(const cg_index* idx, int32_t* out) {
    if (!idx || !out) CG_FAIL(CG_ERR_ARG, "null argument");
    CG_CUDA(cudaSetDevice(idx->device));
    CG_CUDA(cudaMemcpy(out, idx->d_sa, idx->n * 4, cudaMemcpyDeviceToHost));
    return CG_OK;
}

} // extern "C"
