// circall_b200 — internal declarations shared by the translation units of libcircall_gpu.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/circall_gpu.h"
#include "cg_core.cuh"

struct cg_index {
    int device = 0;
    int k = 31;
    uint64_t n = 0;            // text length incl. '$'
    uint32_t n_tx = 0;
    uint64_t n_kmers = 0;
    uint64_t hcap = 0;         // hash slots
    uint64_t n_blocks = 0;     // text blocks incl. the pad block
    uint4* d_text = nullptr;
    uint32_t* d_sa = nullptr;
    uint4* d_hslots = nullptr;
    uint32_t* d_offsets = nullptr;
    uint32_t* d_satid = nullptr;   // transcriptAtPosition(SA[i]) per suffix-array index, rebuilt from SA + text at build / open time
    uint8_t* d_lcp = nullptr;      // min(255, lcp(SA[i-1], SA[i])) per suffix-array index, rebuilt from SA + text at build / open time
    uint32_t* d_fbits = nullptr;   // seed filter bit set (4^fm bits), rebuilt from the text at build / open time
    int fm = 0;
    std::vector<uint32_t> offsets;   // n_tx + 1
    std::vector<std::string> names;
    double build_seconds = 0;
    uint64_t device_bytes = 0;
    cg::IndexView view() const {
        cg::IndexView v;
        v.text = d_text; v.sa = d_sa; v.hslots = d_hslots; v.hmask = hcap - 1; v.offsets = d_offsets;
        v.hshift = 64; for (uint64_t c = hcap; c > 1; c >>= 1) --v.hshift;
        v.n = (uint32_t)n; v.n_tx = n_tx; v.k = k;
        v.km = cg::kmask(k); v.km3 = v.km / 3;
        v.fbits = d_fbits; v.fm = d_fbits ? fm : 0; v.satid = d_satid; v.lcp = d_lcp;
        return v;
    }
};

namespace cgi {
void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#ifndef CG_L2_FETCH_DEFAULT
#define CG_L2_FETCH_DEFAULT 0          // 0 = leave the driver's default; set by measurement (profiles/r2_experiments.md)
#endif
void device_tuning(int device);        // cg_api.cu

// RAII device buffer used for temporaries inside one API call
struct DevBuf {
    void* p = nullptr;
    uint64_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(uint64_t b) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = b;
        return cudaMalloc(&p, b ? b : 1);
    }
    template <class T> T* as() const { return (T*)p; }
};
// grow-only device buffer owned by a session
struct GrowBuf {
    void* p = nullptr;
    uint64_t cap = 0;
    cudaError_t reserve(uint64_t b) {
        if (b <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        uint64_t want = b + b / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};
struct PinBuf {
    void* p = nullptr;
    uint64_t cap = 0;
    cudaError_t reserve(uint64_t b) {
        if (b <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        uint64_t want = b + b / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};
}

#define CG_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return cgi::cuda_fail(_e, #expr, __FILE__, __LINE__); } while (0)
#define CG_FAIL(code, msg) do { cgi::set_error(msg); return (code); } while (0)
