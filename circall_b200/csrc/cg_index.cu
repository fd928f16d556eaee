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
// The sorts, scans and selections are cg_sortscan.cuh's (hand-written; index construction is set-up, not the timed hot path).
#include "cg_common.h"
#include "cg_sortscan.cuh"
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
    cgs::SortBufs sb; sb.k[0] = kA.as<uint64_t>(); sb.k[1] = kB.as<uint64_t>(); sb.v[0] = sA.as<uint32_t>(); sb.v[1] = sB.as<uint32_t>(); sb.cur = 0;
    const uint64_t te = std::max(cgs::sort_tmp_entries(n), cgs::scan_tmp_entries(n));
    CG_CUDA(tmp.alloc(te * 4));
    const int BS = 256;
    k_sa_init<<<grid_for(n, BS), BS>>>(d_ascii, n, sb.k[0], sb.v[0]);
    CG_CUDA(cudaGetLastError());
    CG_CUDA(cgs::sort_pairs(sb, n, 0, 63, tmp.as<uint32_t>()));
    const int rbits = bits_for(n + 1);
    uint64_t h = 21;
    for (int round = 0; round < 64; ++round) {
        uint32_t* head = (uint32_t*)sb.k[sb.cur ^ 1];   // free scratch until the next sort
        CG_CUDA(cudaMemset(cnt.p, 0, 8));
        k_sa_heads<<<grid_for(n, BS), BS>>>(sb.k[sb.cur], n, head, cnt.as<unsigned long long>());
        CG_CUDA(cudaGetLastError());
        unsigned long long ngroups = 0;
        CG_CUDA(cudaMemcpy(&ngroups, cnt.p, 8, cudaMemcpyDeviceToHost));
        if (ngroups == n) break;
        if (h >= n) CG_FAIL(CG_ERR_STATE, "suffix array construction did not converge");
        CG_CUDA((cgs::scan_u32<cgs::OpMax, true>(head, head, n, tmp.as<uint32_t>(), nullptr)));
        k_sa_scatter_rank<<<grid_for(n, BS), BS>>>(sb.v[sb.cur], head, n, rk.as<uint32_t>());
        CG_CUDA(cudaGetLastError());
        k_sa_keys<<<grid_for(n, BS), BS>>>(sb.v[sb.cur], rk.as<uint32_t>(), n, h, sb.k[sb.cur]);
        CG_CUDA(cudaGetLastError());
        CG_CUDA(cgs::sort_pairs(sb, n, 0, 32 + rbits, tmp.as<uint32_t>()));
        h *= 2;
    }
    CG_CUDA(cudaDeviceSynchronize());
    // hand the buffer that holds the result to the caller
    if (sb.v[sb.cur] == sA.as<uint32_t>()) { *out_sa = sA.as<uint32_t>(); sA.p = nullptr; }
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
    CG_CUDA(tmp.alloc((cgs::scan_tmp_entries(n) + 2) * 4));
    CG_CUDA(cgs::select_flagged(bnd.as<uint8_t>(), n, B.as<uint32_t>(), nsel.as<unsigned long long>(), tmp.as<uint32_t>()));
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
        DevBuf tmp;
        CG_TRY(tmp.alloc(cgs::scan_tmp_entries(idx->n_blocks) * 4));
        CG_TRY((cgs::scan_u32<cgs::OpSum, false>(cnt.as<uint32_t>(), cnt.as<uint32_t>(), idx->n_blocks, tmp.as<uint32_t>(), nullptr)));
        CG_TRY(cudaDeviceSynchronize());
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
// ---- directory layout ---------------------------------------------------------------------------------------------------
// The file set RapMap's indexer writes behind TxIndexer (recollected — no sample index ships with the reference, so
// the layout is UNPINNED), which is what an existing -txIdx / -bsjIdx directory holds:
//   header.json        cereal JSON: IndexType, IndexVersion, UsesKmers, KmerLen, BigSA, PerfectHash
//   versionInfo.json   Sailfish: indexVersion, indexType, kmerLength            (read at Circall_wt.cpp:1691-1693)
//   txpInfo.bin        cereal binary: vector<string> names, vector<IndexT> offsets, string text ('$'-joined), vector<uint32> lengths
//                      (cereal binary = u64 element count, then the raw little-endian payload; a string = u64 length + bytes)
//   sa.bin             cereal binary vector<IndexT>            rsd.bin   bit array with a 1 at every '$': u64 bit count + bytes
//   hash.bin           sparsehash dense_hash_map dump — never read here and not written: the k-mer table is rebuilt from SA + text
// plus this library's own cache beside them, so that opening does not have to rebuild anything:
//   cg_meta.bin        magic "CGIDX002", k, n, n_tx, n_kmers, hcap, n_blocks   (uint64 each)
//   cg_text.bin        n_blocks x 32 B        cg_hash.bin    hcap x 16 B
// cg_index_open reads names / offsets from txpInfo.bin; with a valid cache it loads cg_text.bin, sa.bin and cg_hash.bin, without
// one (a directory the reference's TxIndexer wrote) it takes the text from txpInfo.bin and builds suffix array and table on the device.
namespace {

static const uint64_t CG_META_MAGIC = 0x3230305844494743ULL;   // "CGIDX002"

struct RefDir {
    std::vector<std::string> names;
    std::vector<uint32_t> offsets;     // n_tx + 1 (the text length appended)
    std::string text;                  // only when asked for
    uint64_t n = 0;
    int k = 0;
    bool big = false;
};

bool rd_u64(FILE* f, uint64_t& v) { return fread(&v, 8, 1, f) == 1; }

// KmerLen / kmerLength from header.json or versionInfo.json (0 when neither says)
int json_int(const std::string& path, const char* const* keys) {
    std::ifstream in(path);
    if (!in) return 0;
    std::string js((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    for (const char* const* k = keys; *k; ++k) {
        size_t p = js.find(std::string("\"") + *k + "\"");
        if (p == std::string::npos) continue;
        p = js.find(':', p);
        if (p == std::string::npos) continue;
        ++p;
        while (p < js.size() && (js[p] == ' ' || js[p] == '\t' || js[p] == '"')) ++p;
        if (js.compare(p, 4, "true") == 0) return 1;
        if (js.compare(p, 5, "false") == 0) return 0;
        int v = 0; bool any = false;
        while (p < js.size() && js[p] >= '0' && js[p] <= '9') { v = v * 10 + (js[p] - '0'); ++p; any = true; }
        if (any) return v;
    }
    return 0;
}

// txpInfo.bin (+ the k the JSON files state); the text itself only when want_text
int read_refdir(const std::string& d, bool want_text, RefDir& R) {
    static const char* const kkeys[] = {"KmerLen", "kmerLen", "kmerLength", nullptr};
    static const char* const bkeys[] = {"BigSA", "bigSA", nullptr};
    R.k = json_int(d + "/header.json", kkeys);
    if (!R.k) R.k = json_int(d + "/versionInfo.json", kkeys);
    if (!R.k) R.k = 31;                                                  // TxIndexer's default (TxIndexer.cpp:87)
    R.big = json_int(d + "/header.json", bkeys) != 0;
    struct stat st;
    if (stat((d + "/txpInfo.bin").c_str(), &st) != 0) CG_FAIL(CG_ERR_IO, "not an index directory (no txpInfo.bin): " + d);
    const uint64_t fsz = (uint64_t)st.st_size;
    FILE* f = fopen((d + "/txpInfo.bin").c_str(), "rb");
    if (!f) CG_FAIL(CG_ERR_IO, "cannot open " + d + "/txpInfo.bin");
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
    auto bad = [&](const char* what) { cgi::set_error(std::string("txpInfo.bin: ") + what); return (int)CG_ERR_IO; };
    uint64_t n_tx = 0;
    if (!rd_u64(f, n_tx) || n_tx == 0 || n_tx > fsz / 8) return bad("transcript count out of range");
    R.names.resize((size_t)n_tx);
    for (uint64_t t = 0; t < n_tx; ++t) {
        uint64_t len = 0;
        if (!rd_u64(f, len) || len > 65536) return bad("name length out of range");
        R.names[t].resize((size_t)len);
        if (len && fread(&R.names[t][0], 1, (size_t)len, f) != len) return bad("truncated name");
    }
    uint64_t n_off = 0;
    if (!rd_u64(f, n_off) || n_off != n_tx) return bad("offset count does not match the transcript count");
    const long at_off = ftell(f);
    // IndexT is int32 unless the text needed the 64-bit index (BigSA): take what header.json says, and check it against the
    // position of the text's length field either way
    int width = R.big ? 8 : 4;
    for (int attempt = 0; attempt < 2; ++attempt) {
        uint64_t tl = 0;
        if (fseek(f, at_off + (long)(n_off * (uint64_t)width), SEEK_SET) == 0 && rd_u64(f, tl) && tl > 0 &&
            (uint64_t)at_off + n_off * (uint64_t)width + 8 + tl + 8 + n_tx * 4 <= fsz) { R.n = tl; break; }
        width = 12 - width;
    }
    if (!R.n) return bad("cannot locate the text (neither 32- nor 64-bit offsets fit the file)");
    if (R.n >= (1ULL << 31) - 128) { cgi::set_error("text >= 2^31: the 64-bit index (RapMapSAIndex<int64_t>) is not implemented"); return CG_ERR_LENGTH; }
    fseek(f, at_off, SEEK_SET);
    R.offsets.resize((size_t)n_tx + 1);
    for (uint64_t t = 0; t < n_tx; ++t) {
        uint64_t v = 0;
        if (fread(&v, (size_t)width, 1, f) != 1) return bad("truncated offsets");
        if (v >= R.n || (t && v <= R.offsets[t - 1])) return bad("offsets must increase and lie inside the text");
        R.offsets[t] = (uint32_t)v;
    }
    R.offsets[n_tx] = (uint32_t)R.n;
    if (R.offsets[0] != 0) return bad("first offset must be 0");
    uint64_t tl = 0;
    rd_u64(f, tl);
    if (want_text) {
        R.text.resize((size_t)R.n);
        if (fread(&R.text[0], 1, (size_t)R.n, f) != R.n) return bad("truncated text");
        if (R.text[R.n - 1] != '$') return bad("text must end with '$'");
        for (uint64_t t = 1; t <= n_tx; ++t)
            if (R.text[R.offsets[t] - 1] != '$') return bad("a transcript does not end with '$' where the next offset says");
    }
    return CG_OK;
}

void put_u64(std::string& o, uint64_t v) { o.append((const char*)&v, 8); }

// the ASCII text ('$'-joined) back from the device's 2-bit blocks
int text_from_device(const cg_index* idx, std::string& text) {
    std::vector<uint4> blk((size_t)idx->n_blocks * 2);
    CG_CUDA(cudaMemcpy(blk.data(), idx->d_text, idx->n_blocks * 32, cudaMemcpyDeviceToHost));
    text.resize((size_t)idx->n);
    for (uint64_t p = 0; p < idx->n; ++p) {
        const uint4& c = blk[2 * (p >> 6)]; const uint4& m = blk[2 * (p >> 6) + 1];
        const uint32_t o = (uint32_t)(p & 63);
        const uint64_t codes = o < 32 ? ((uint64_t)c.x | ((uint64_t)c.y << 32)) : ((uint64_t)c.z | ((uint64_t)c.w << 32));
        const uint64_t mask = (uint64_t)m.x | ((uint64_t)m.y << 32);
        text[(size_t)p] = ((mask >> o) & 1ULL) ? '$' : "ACGT"[(codes >> (2 * (o & 31))) & 3ULL];
    }
    return CG_OK;
}

} // namespace

static int index_save_impl(const cg_index* idx, const char* dir) {
    CG_CUDA(cudaSetDevice(idx->device));
    std::string d(dir);
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) CG_FAIL(CG_ERR_IO, "cannot create directory " + d);
    int rc;
    // ---- the reference's file set
    std::string text;
    if ((rc = text_from_device(idx, text))) return rc;
    {
        std::string o;
        put_u64(o, idx->n_tx);
        for (auto& s : idx->names) { put_u64(o, s.size()); o += s; }
        put_u64(o, idx->n_tx);
        o.append((const char*)idx->offsets.data(), (size_t)idx->n_tx * 4);          // IndexT = int32 (text < 2^31)
        put_u64(o, idx->n);
        FILE* f = fopen((d + "/txpInfo.bin").c_str(), "wb");
        if (!f) CG_FAIL(CG_ERR_IO, "cannot write " + d + "/txpInfo.bin");
        bool ok = fwrite(o.data(), 1, o.size(), f) == o.size() && fwrite(text.data(), 1, text.size(), f) == text.size();
        std::string l;
        put_u64(l, idx->n_tx);
        for (uint32_t t = 0; t < idx->n_tx; ++t) { const uint32_t len = idx->offsets[t + 1] - idx->offsets[t] - 1; l.append((const char*)&len, 4); }
        ok = ok && fwrite(l.data(), 1, l.size(), f) == l.size();
        if (fclose(f) != 0 || !ok) CG_FAIL(CG_ERR_IO, "cannot write " + d + "/txpInfo.bin");
    }
    {
        std::vector<char> h(8 + idx->n * 4);
        memcpy(h.data(), &idx->n, 8);
        CG_CUDA(cudaMemcpy(h.data() + 8, idx->d_sa, idx->n * 4, cudaMemcpyDeviceToHost));
        if (!write_file(d + "/sa.bin", h.data(), h.size())) CG_FAIL(CG_ERR_IO, "cannot write sa.bin");
    }
    {
        std::vector<unsigned char> bits(8 + (idx->n + 7) / 8, 0);
        memcpy(bits.data(), &idx->n, 8);
        for (uint64_t p = 0; p < idx->n; ++p) if (text[(size_t)p] == '$') bits[8 + (p >> 3)] |= (unsigned char)(1u << (p & 7));
        if (!write_file(d + "/rsd.bin", bits.data(), bits.size())) CG_FAIL(CG_ERR_IO, "cannot write rsd.bin");
    }
    text.clear(); text.shrink_to_fit();
    char js[768];
    snprintf(js, sizeof js,
             "{\n    \"value0\": {\n        \"IndexType\": 1,\n        \"IndexVersion\": \"q3\",\n        \"UsesKmers\": true,\n"
             "        \"KmerLen\": %d,\n        \"BigSA\": false,\n        \"PerfectHash\": false,\n        \"writer\": \"circall_b200\",\n"
             "        \"textLen\": %llu,\n        \"numTranscripts\": %u,\n        \"numKmers\": %llu\n    }\n}\n",
             idx->k, (unsigned long long)idx->n, idx->n_tx, (unsigned long long)idx->n_kmers);
    {
        char vj[256];
        snprintf(vj, sizeof vj, "{\n    \"indexVersion\": 2,\n    \"indexType\": 1,\n    \"kmerLength\": %d\n}\n", idx->k);
        if (!write_file(d + "/versionInfo.json", vj, strlen(vj))) CG_FAIL(CG_ERR_IO, "cannot write versionInfo.json");
    }
    // ---- this library's cache
    if ((rc = save_dev(d + "/cg_text.bin", idx->d_text, idx->n_blocks * 32))) return rc;
    if ((rc = save_dev(d + "/cg_hash.bin", idx->d_hslots, idx->hcap * 16))) return rc;
    uint64_t meta[8] = {CG_META_MAGIC, (uint64_t)idx->k, idx->n, idx->n_tx, idx->n_kmers, idx->hcap, idx->n_blocks, 0};
    if (!write_file(d + "/cg_meta.bin", meta, sizeof meta)) CG_FAIL(CG_ERR_IO, "cannot write " + d + "/cg_meta.bin");
    // header.json last: its presence marks a finished index (TxIndexer.cpp:192-193)
    if (!write_file(d + "/header.json", js, strlen(js))) CG_FAIL(CG_ERR_IO, "cannot write header.json");
    return CG_OK;
}

static bool file_size_is(const std::string& path, uint64_t bytes) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 && (uint64_t)st.st_size == bytes;
}

// is the cache beside the reference's files usable for this text?  (nothing in cg_meta.bin is trusted: every field is checked
// against the others, against txpInfo.bin and against the files' sizes before anything is allocated or indexed with it)
static bool cache_valid(const std::string& d, const RefDir& R, uint64_t meta[8]) {
    if (!file_size_is(d + "/cg_meta.bin", 64) || !read_file(d + "/cg_meta.bin", meta, 64) || meta[0] != CG_META_MAGIC) return false;
    const uint64_t k = meta[1], n = meta[2], n_tx = meta[3], n_kmers = meta[4], hcap = meta[5], n_blocks = meta[6];
    if (k < 3 || k > 31 || (k & 1) == 0 || (int)k != R.k) return false;
    if (n != R.n || n_tx + 1 != R.offsets.size()) return false;
    if (n_blocks != (n + 63) / 64 + 1) return false;
    if (hcap < 16 || (hcap & (hcap - 1)) != 0 || hcap > (1ULL << 33) || n_kmers >= hcap) return false;
    return file_size_is(d + "/cg_text.bin", n_blocks * 32) && file_size_is(d + "/sa.bin", 8 + n * 4) && file_size_is(d + "/cg_hash.bin", hcap * 16);
}

static int index_open_impl(const char* dir, int device, cg_index** out) {
    std::string d(dir);
    RefDir R;
    int rc = read_refdir(d, false, R);
    if (rc) return rc;
    uint64_t meta[8];
    if (!cache_valid(d, R, meta)) {
        // a directory without this library's cache (the reference's TxIndexer wrote it): text from txpInfo.bin, everything else rebuilt
        if ((rc = read_refdir(d, true, R))) return rc;
        cg_index* idx = nullptr;
        rc = build_impl(device, R.text.data(), R.text.size(), R.k, &R.names, &idx);
        if (rc) return rc;
        idx->build_seconds = 0;
        *out = idx;
        return CG_OK;
    }
    rc = check_device(device);
    if (rc) return rc;
    cg_index* idx = new cg_index();
    idx->device = device; idx->k = (int)meta[1]; idx->n = meta[2]; idx->n_tx = (uint32_t)meta[3];
    idx->n_kmers = meta[4]; idx->hcap = meta[5]; idx->n_blocks = meta[6];
    auto fail = [&](int code) { free_index_device(idx); delete idx; return code; };
    idx->offsets = R.offsets;
    idx->names = R.names;
    if ((rc = load_dev(d + "/cg_text.bin", (void**)&idx->d_text, idx->n_blocks * 32))) return fail(rc);
    {
        std::vector<char> h(8 + idx->n * 4);
        uint64_t cnt = 0;
        if (!read_file(d + "/sa.bin", h.data(), h.size()) || (memcpy(&cnt, h.data(), 8), cnt != idx->n)) { cgi::set_error("sa.bin does not hold one 32-bit entry per text position"); return fail(CG_ERR_IO); }
        cudaError_t e = cudaMalloc((void**)&idx->d_sa, idx->n * 4);
        if (e == cudaSuccess) e = cudaMemcpy(idx->d_sa, h.data() + 8, idx->n * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(cgi::cuda_fail(e, "upload sa.bin", __FILE__, __LINE__));
    }
    if ((rc = load_dev(d + "/cg_hash.bin", (void**)&idx->d_hslots, idx->hcap * 16))) return fail(rc);
    cudaError_t e = cudaMalloc((void**)&idx->d_offsets, idx->offsets.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(idx->d_offsets, idx->offsets.data(), idx->offsets.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(cgi::cuda_fail(e, "upload offsets", __FILE__, __LINE__));
    if ((rc = build_fbits(idx))) return fail(rc);          // the seed filter, the per-suffix transcript array and the lcp bytes are derived
    if ((rc = build_satid(idx))) return fail(rc);          // from text + SA: not part of the directory
    if ((rc = build_lcp(idx))) return fail(rc);
    finish_sizes(idx);
    *out = idx;
    return CG_OK;
}

// what a directory holds, read on the host (no device needed): layout 1 = the reference's file set only, 2 = with this library's cache
int cg_index_dir_info(const char* dir, cg_index_info* info, int* layout) {
    if (!dir || !info) CG_FAIL(CG_ERR_ARG, "null argument");
    try {
        RefDir R;
        int rc = read_refdir(dir, false, R);
        if (rc) return rc;
        uint64_t meta[8];
        const bool cached = cache_valid(dir, R, meta);
        memset(info, 0, sizeof *info);
        info->text_len = R.n; info->n_tx = (uint32_t)(R.offsets.size() - 1); info->k = R.k; info->device = -1;
        if (cached) { info->n_kmers = meta[4]; info->hash_capacity = meta[5]; }
        if (layout) *layout = cached ? 2 : 1;
        return CG_OK;
    } catch (const std::exception& e) { CG_FAIL(CG_ERR_IO, std::string("index directory unreadable: ") + e.what()); }
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

// The index builder's own primitives (cg_sortscan.cuh) on caller-provided data, results back on the host: the tests compare them with
// numpy at sizes around the tile boundaries (the suffix-array tests only see them through the finished index).
//   keys / vals (n each): sorted in place by key bits [begin_bit, end_bit), stably
int cg_debug_sort_pairs(int device, uint64_t* keys, uint32_t* vals, uint64_t n, int begin_bit, int end_bit) {
    if ((n && (!keys || !vals)) || begin_bit < 0 || end_bit > 64 || begin_bit > end_bit) CG_FAIL(CG_ERR_ARG, "bad argument");
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) return CG_OK;
    DevBuf kA, kB, vA, vB, tmp;
    CG_CUDA(kA.alloc(n * 8)); CG_CUDA(kB.alloc(n * 8)); CG_CUDA(vA.alloc(n * 4)); CG_CUDA(vB.alloc(n * 4));
    CG_CUDA(tmp.alloc(cgs::sort_tmp_entries(n) * 4));
    CG_CUDA(cudaMemcpy(kA.p, keys, n * 8, cudaMemcpyHostToDevice));
    CG_CUDA(cudaMemcpy(vA.p, vals, n * 4, cudaMemcpyHostToDevice));
    cgs::SortBufs sb; sb.k[0] = kA.as<uint64_t>(); sb.k[1] = kB.as<uint64_t>(); sb.v[0] = vA.as<uint32_t>(); sb.v[1] = vB.as<uint32_t>(); sb.cur = 0;
    CG_CUDA(cgs::sort_pairs(sb, n, begin_bit, end_bit, tmp.as<uint32_t>()));
    CG_CUDA(cudaDeviceSynchronize());
    CG_CUDA(cudaMemcpy(keys, sb.k[sb.cur], n * 8, cudaMemcpyDeviceToHost));
    CG_CUDA(cudaMemcpy(vals, sb.v[sb.cur], n * 4, cudaMemcpyDeviceToHost));
    return CG_OK;
}
//   mode 0: exclusive sum, 1: inclusive max of data[0..n) in place; mode 2: data = flags (one byte each in the low byte of every word):
//   out_idx receives the indices of the non-zero ones, *n_out their number
int cg_debug_scan(int device, uint32_t* data, uint64_t n, int mode, uint32_t* out_idx, uint64_t* n_out) {
    if ((n && !data) || mode < 0 || mode > 2 || (mode == 2 && (!out_idx || !n_out))) CG_FAIL(CG_ERR_ARG, "bad argument");
    int rc = check_device(device);
    if (rc) return rc;
    if (n == 0) { if (n_out) *n_out = 0; return CG_OK; }
    DevBuf d, tmp, f, o, cnt;
    CG_CUDA(d.alloc(n * 4)); CG_CUDA(tmp.alloc((cgs::scan_tmp_entries(n) + 2) * 4));
    CG_CUDA(cudaMemcpy(d.p, data, n * 4, cudaMemcpyHostToDevice));
    if (mode == 0) CG_CUDA((cgs::scan_u32<cgs::OpSum, false>(d.as<uint32_t>(), d.as<uint32_t>(), n, tmp.as<uint32_t>(), nullptr)));
    else if (mode == 1) CG_CUDA((cgs::scan_u32<cgs::OpMax, true>(d.as<uint32_t>(), d.as<uint32_t>(), n, tmp.as<uint32_t>(), nullptr)));
    else {
        std::vector<uint8_t> hf(n);
        for (uint64_t i = 0; i < n; ++i) hf[i] = (uint8_t)(data[i] & 0xffu);
        CG_CUDA(f.alloc(n)); CG_CUDA(o.alloc(n * 4)); CG_CUDA(cnt.alloc(8));
        CG_CUDA(cudaMemcpy(f.p, hf.data(), n, cudaMemcpyHostToDevice));
        CG_CUDA(cgs::select_flagged(f.as<uint8_t>(), n, o.as<uint32_t>(), cnt.as<unsigned long long>(), tmp.as<uint32_t>()));
        CG_CUDA(cudaDeviceSynchronize());
        unsigned long long c = 0;
        CG_CUDA(cudaMemcpy(&c, cnt.p, 8, cudaMemcpyDeviceToHost));
        *n_out = c;
        CG_CUDA(cudaMemcpy(out_idx, o.p, c * 4, cudaMemcpyDeviceToHost));
        return CG_OK;
    }
    CG_CUDA(cudaDeviceSynchronize());
    CG_CUDA(cudaMemcpy(data, d.p, n * 4, cudaMemcpyDeviceToHost));
    return CG_OK;
}

} // extern "C"
