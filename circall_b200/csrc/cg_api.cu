// circall_b200 — C-ABI odds and ends: error text, device helpers, the collector parity surface
// (cg_collect = SACollectorCircall::operator() on a batch of reads), the mate-hit intersection
// (cg_merge_pairs = mergeLeftRightHitsFuzzy) and the random-sector gather microbenchmark.
#include "cg_kernels.cuh"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>

using namespace cg;
using namespace cgk;
using cgi::DevBuf;

namespace cgi {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")";
    cudaGetLastError();
    return CG_ERR_CUDA;
}
// Device-wide settings of the random-access path, applied whenever the library touches a device.  The k-mer table, suffix array and
// text are read one 32-byte sector at a time at random addresses: the L2's default DRAM fetch granularity pulls in whole 128-byte
// lines for them (110 B of DRAM traffic per missed sector).  CG_L2_FETCH = 32 / 64 / 128 overrides the default below.
void device_tuning(int device) {
    static int applied[64] = {0};
    if (device < 0 || device >= 64 || applied[device]) return;
    applied[device] = 1;
    size_t want = CG_L2_FETCH_DEFAULT;
    if (const char* e = getenv("CG_L2_FETCH")) want = (size_t)atoi(e);
    if (want == 32 || want == 64 || want == 128) {
        if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, want) != cudaSuccess) cudaGetLastError();   // a hint: never fatal
    }
    if (getenv("CG_DEBUG")) {
        size_t got = 0;
        cudaError_t e = cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "circall_gpu: device %d cudaLimitMaxL2FetchGranularity asked %zu, now %zu (%s)\n", device, want, got, cudaGetErrorString(e));
    }
}
}

namespace {

inline unsigned grid_for(uint64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- collector surface ---------------------------------------------------------------------------
static const int CMAXINT = 232, CMAXKS = 928, CMAXHITS = 2 * MAXC;

struct CollectOut {
    uint8_t* found; uint32_t* nF; uint32_t* nR; uint32_t* nH;
    int4* fwd; int4* rc;                        // CMAXINT per read
    uint32_t* hit_tid; int32_t* hit_pos; uint8_t* hit_fwd;   // CMAXHITS per read
    int* status;
};

__global__ void __launch_bounds__(128)
k_collect(IndexView ix, const uint64_t* __restrict__ pk, const uint16_t* __restrict__ lens, PackGeom g, uint64_t n_reads,
          int strict, int lceMode, CollectOut o, int rstride) {
    extern __shared__ uint64_t smem[];
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int L = lens[r * rstride];
    ReadView rv = make_view(pk + r * (uint64_t)rstride * g.WT, g, L, smem);
    Scratch<CMAXINT, CMAXKS, true> sc;
    bool f = collect_hits<CMAXINT, CMAXKS, true>(ix, rv, strict != 0, lceMode, sc);
    if (sc.st.err) atomicOr(o.status, 4);
    o.found[r] = f;
    uint32_t nf = 0, nr = 0, nh = 0;
    if (f) {
        nf = sc.st.nF; nr = sc.st.nR;
        for (int i = 0; i < sc.st.nF; ++i) { const Interval& I = sc.st.fwd[i]; o.fwd[r * CMAXINT + i] = make_int4((int)I.begin, (int)I.end, I.len, I.qpos); }
        for (int i = 0; i < sc.st.nR; ++i) { const Interval& I = sc.st.rc[i]; o.rc[r * CMAXINT + i] = make_int4((int)I.begin, (int)I.end, I.len, I.qpos); }
        uint32_t* ht = o.hit_tid + r * (uint64_t)CMAXHITS; int32_t* hp = o.hit_pos + r * (uint64_t)CMAXHITS; uint8_t* hf = o.hit_fwd + r * (uint64_t)CMAXHITS;
        merged_hits(sc, [&](uint32_t t, bool fw, int i) {
            if (nh < (uint32_t)CMAXHITS) { ht[nh] = t; hp[nh] = (int32_t)(fw ? sc.posF[i] : sc.posR[i]); hf[nh] = fw; }
            ++nh;
        });
    }
    o.nF[r] = nf; o.nR[r] = nr; o.nH[r] = nh;
}

struct CollectHost {
    std::vector<uint8_t> found, hit_fwd;
    std::vector<uint64_t> fwd_off, rc_off, hits_off;
    std::vector<int32_t> fwd_ints, rc_ints, hit_pos;
    std::vector<uint32_t> hit_tid;
};
static thread_local CollectHost* g_collect = nullptr;

// ---- mate-hit intersection (K9) -------------------------------------------------------------------
// One thread per pair; pass 0 counts the joint hits, pass 1 writes them at joint_off[pair].
struct MergeIn {
    const uint64_t* loff; const uint32_t* ltid; const int32_t* lpos; const uint8_t* lfwd; const uint8_t* lseed;
    const uint64_t* roff; const uint32_t* rtid; const int32_t* rpos; const uint8_t* rfwd; const uint8_t* rseed;
};

__global__ void __launch_bounds__(256)
k_merge(MergeIn in, uint64_t n_pairs, uint32_t readLen, uint32_t maxNumHits, int pass, uint64_t* cnt,
        const uint64_t* __restrict__ joint_off, int32_t* __restrict__ joint, uint8_t* __restrict__ too_many) {
    uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    uint64_t l = in.loff[p], le = in.loff[p + 1], r = in.roff[p], re = in.roff[p + 1];
    int32_t* out = pass ? joint + 7 * joint_off[p] : nullptr;
    uint64_t n = 0;
    bool tm = false;
    auto orphan = [&](uint64_t b, uint64_t e, bool left) {
        for (uint64_t i = b; i < e; ++i, ++n)
            if (out) {
                int32_t* o = out + 7 * n;
                o[0] = (int32_t)(left ? in.ltid[i] : in.rtid[i]); o[1] = left ? in.lpos[i] : in.rpos[i];
                o[2] = left ? in.lfwd[i] : in.rfwd[i]; o[3] = 0; o[4] = 0; o[5] = 0; o[6] = left ? 1 : 2;
            }
    };
    if (l == le) {
        if (!in.lseed[p] && r != re) orphan(r, re, false);          // left mate had no seed at all: right orphans
    } else if (r == re) {
        if (!in.rseed[p]) orphan(l, le, true);
    } else {
        uint64_t a = l, b = r;
        while (a != le && b != re) {
            uint32_t ta = in.ltid[a], tb = in.rtid[b];
            if (ta < tb) ++a;
            else if (tb < ta) ++b;
            else {
                int32_t s1 = in.lpos[a] > 0 ? in.lpos[a] : 0, s2 = in.rpos[b] > 0 ? in.rpos[b] : 0;
                bool firstLeft = s1 < s2;
                int32_t fs = firstLeft ? s1 : s2;
                int32_t fe = firstLeft ? s2 + (int32_t)readLen : s1 + (int32_t)readLen;
                if (out) {
                    int32_t* o = out + 7 * n;
                    o[0] = (int32_t)ta; o[1] = s1; o[2] = in.lfwd[a]; o[3] = s2; o[4] = in.rfwd[b]; o[5] = fe - fs; o[6] = 3;
                }
                ++n;
                if (n > maxNumHits) { tm = true; break; }
                ++a; ++b;
            }
        }
        if (n == 0) { orphan(l, le, true); orphan(r, re, false); }  // nothing in common: all left then all right
    }
    if (pass == 0) cnt[p] = n;
    else too_many[p] = tm;
}

// ---- random-sector gather microbenchmark ------------------------------------------------------------
// MODE picks the load instruction: what the L2 fetches from DRAM for a missing 32-byte sector depends on it (profiles/r2_experiments.md)
//   0 ld.global.nc (what __ldg compiles to)   1 .nc.L2::64B   2 .nc.L2::128B   3 .nc.L2::256B   4 .nc.L1::no_allocate
//   5 ld.global.cg   6 ld.global (L1-allocating)   7 .nc.L1::no_allocate.L2::64B   8 ld.global.cv   9 both halves of the sector (2 x 16 B)
template <int MODE>
__device__ __forceinline__ uint4 gather_load(const uint4* p) {
    uint4 v;
    if (MODE == 1) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 5) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 6) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 7) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 8) asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 9) { uint4 a = __ldg(p), b = __ldg(p + 1); v.x = a.x ^ b.x; v.y = a.y; v.z = b.z; v.w = a.w ^ b.w; }
    else v = __ldg(p);
    return v;
}
template <int MODE>
__global__ void __launch_bounds__(256)
k_gather(const uint4* __restrict__ buf, uint64_t n_sectors_mask, uint32_t per_thread, uint64_t seed, uint32_t* sink) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    uint64_t x = hmix(t * 0x9E3779B97F4A7C15ULL + seed);
    for (uint32_t i = 0; i < per_thread; i += 4) {
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x = x * 6364136223846793005ULL + 1442695040888963407ULL;
            uint64_t s = (x >> 20) & n_sectors_mask;
            v[j] = gather_load<MODE>(&buf[2 * s]);   // first 16 B of a 32-B sector
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += v[j].x ^ v[j].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
typedef void (*gather_fn)(const uint4*, uint64_t, uint32_t, uint64_t, uint32_t*);
static gather_fn gather_kernel(int mode) {
    switch (mode) {
        case 1: return k_gather<1>; case 2: return k_gather<2>; case 3: return k_gather<3>; case 4: return k_gather<4>; case 5: return k_gather<5>;
        case 6: return k_gather<6>; case 7: return k_gather<7>; case 8: return k_gather<8>; case 9: return k_gather<9>; default: return k_gather<0>;
    }
}

} // namespace

extern "C" {

const char* cg_last_error(void) { return cgi::g_err.c_str(); }

int cg_device_count(int* n) {
    if (!n) CG_FAIL(CG_ERR_ARG, "null argument");
    int nd = 0;
    cudaError_t e = cudaGetDeviceCount(&nd);
    if (e != cudaSuccess) { cudaGetLastError(); nd = 0; }
    *n = nd;
    return CG_OK;
}

int cg_device_alloc(int device, uint64_t bytes, void** dptr) {
    if (!dptr) CG_FAIL(CG_ERR_ARG, "null argument");
    CG_CUDA(cudaSetDevice(device));
    CG_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return CG_OK;
}
int cg_device_free(int device, void* dptr) {
    CG_CUDA(cudaSetDevice(device));
    CG_CUDA(cudaFree(dptr));
    return CG_OK;
}
int cg_device_upload(int device, void* dptr, const void* src, uint64_t bytes) {
    CG_CUDA(cudaSetDevice(device));
    CG_CUDA(cudaMemcpy(dptr, src, bytes, cudaMemcpyHostToDevice));
    return CG_OK;
}
int cg_device_download(int device, void* dst, const void* dptr, uint64_t bytes) {
    CG_CUDA(cudaSetDevice(device));
    CG_CUDA(cudaMemcpy(dst, dptr, bytes, cudaMemcpyDeviceToHost));
    return CG_OK;
}
int cg_host_alloc(uint64_t bytes, void** ptr) {
    if (!ptr) CG_FAIL(CG_ERR_ARG, "null argument");
    CG_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return CG_OK;
}
int cg_host_free(void* ptr) {
    CG_CUDA(cudaFreeHost(ptr));
    return CG_OK;
}
int cg_device_flush_l2(int device) {
    // write a buffer larger than the 126 MB L2 so the next timed iteration starts cold
    CG_CUDA(cudaSetDevice(device));
    static thread_local void* buf = nullptr;
    const uint64_t bytes = 256ull << 20;
    if (!buf) CG_CUDA(cudaMalloc(&buf, bytes));
    static thread_local int v = 0;
    CG_CUDA(cudaMemset(buf, ++v & 0xff, bytes));
    CG_CUDA(cudaDeviceSynchronize());
    return CG_OK;
}

int cg_collect(const cg_index* idx, const char* bases, const uint64_t* offsets, uint64_t n_reads, int strict_check, int lce_mode,
               cg_collect_result* out) {
    if (!idx || !offsets || !out || (n_reads && !bases)) CG_FAIL(CG_ERR_ARG, "null argument");
    CG_CUDA(cudaSetDevice(idx->device));
    if (!g_collect) g_collect = new CollectHost();
    CollectHost& H = *g_collect;
    H = CollectHost();
    H.found.resize(n_reads);
    H.fwd_off.assign(1, 0); H.rc_off.assign(1, 0); H.hits_off.assign(1, 0);
    uint64_t mx = 0;
    for (uint64_t i = 0; i < n_reads; ++i) {
        if (offsets[i + 1] < offsets[i]) CG_FAIL(CG_ERR_ARG, "offsets must be non-decreasing");
        mx = std::max(mx, offsets[i + 1] - offsets[i]);
    }
    if (mx > CG_MAX_READ_LEN) CG_FAIL(CG_ERR_LENGTH, "a read is longer than CG_MAX_READ_LEN");
    const PackGeom g = PackGeom::make((int)mx);
    const uint64_t CH = 8192;   // reads per chunk (bounded scratch: every read owns worst-case output slots)
    DevBuf dB, dOff, dPk, dLens, dFound, dNF, dNR, dNH, dFwd, dRc, dHt, dHp, dHf, dSt;
    CG_CUDA(dOff.alloc((CH + 1) * 8)); CG_CUDA(dPk.alloc(2 * CH * g.WT * 8)); CG_CUDA(dLens.alloc(2 * CH * 2));
    CG_CUDA(dFound.alloc(CH)); CG_CUDA(dNF.alloc(CH * 4)); CG_CUDA(dNR.alloc(CH * 4)); CG_CUDA(dNH.alloc(CH * 4));
    CG_CUDA(dFwd.alloc(CH * CMAXINT * 16)); CG_CUDA(dRc.alloc(CH * CMAXINT * 16));
    CG_CUDA(dHt.alloc(CH * CMAXHITS * 4)); CG_CUDA(dHp.alloc(CH * CMAXHITS * 4)); CG_CUDA(dHf.alloc(CH * CMAXHITS)); CG_CUDA(dSt.alloc(4));
    std::vector<uint64_t> off2;
    std::vector<uint32_t> nF(CH), nR(CH), nH(CH), ht;
    std::vector<int4> fw, rc;
    std::vector<int32_t> hp;
    std::vector<uint8_t> hf;
    for (uint64_t c0 = 0; c0 < n_reads; c0 += CH) {
        uint64_t nc = std::min(CH, n_reads - c0);
        uint64_t b0 = offsets[c0], nbytes = offsets[c0 + nc] - b0;
        off2.resize(nc + 1);
        for (uint64_t i = 0; i <= nc; ++i) off2[i] = offsets[c0 + i] - b0;
        CG_CUDA(dB.alloc(nbytes + 64));
        if (nbytes) CG_CUDA(cudaMemcpy(dB.p, bases + b0, nbytes, cudaMemcpyHostToDevice));
        CG_CUDA(cudaMemcpy(dOff.p, off2.data(), (nc + 1) * 8, cudaMemcpyHostToDevice));
        CG_CUDA(cudaMemset(dSt.p, 0, 4));
        // k_pack takes pairs: hand it the chunk as both mate files and read the mate-1 records (stride 2)
        k_pack<<<pack_grid(2 * nc, g), PACK_THREADS>>>(dB.as<unsigned char>(), dOff.as<uint64_t>(), dB.as<unsigned char>(), dOff.as<uint64_t>(),
                                                       nc, g, (int)mx, dPk.as<uint64_t>(), dLens.as<uint16_t>(), dSt.as<int>());
        CG_CUDA(cudaGetLastError());
        CollectOut o;
        o.found = dFound.as<uint8_t>(); o.nF = dNF.as<uint32_t>(); o.nR = dNR.as<uint32_t>(); o.nH = dNH.as<uint32_t>();
        o.fwd = dFwd.as<int4>(); o.rc = dRc.as<int4>(); o.hit_tid = dHt.as<uint32_t>(); o.hit_pos = dHp.as<int32_t>(); o.hit_fwd = dHf.as<uint8_t>();
        o.status = dSt.as<int>();
        const int BS = 128;
        size_t smem = (size_t)BS * g.chunk() * 8;
        if (smem > 48 * 1024) CG_CUDA(cudaFuncSetAttribute(k_collect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_collect<<<grid_for(nc, BS), BS, smem>>>(idx->view(), dPk.as<uint64_t>(), dLens.as<uint16_t>(), g, nc, strict_check, lce_mode, o, 2);
        CG_CUDA(cudaGetLastError());
        int st = 0;
        CG_CUDA(cudaMemcpy(&st, dSt.p, 4, cudaMemcpyDeviceToHost));
        if (st & 1) CG_FAIL(CG_ERR_BASE, "a read holds a character outside ACGTNacgtn");
        if (st & 4) CG_FAIL(CG_ERR_CAPACITY, "collector capacity exceeded");
        CG_CUDA(cudaMemcpy(H.found.data() + c0, dFound.p, nc, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(nF.data(), dNF.p, nc * 4, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(nR.data(), dNR.p, nc * 4, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(nH.data(), dNH.p, nc * 4, cudaMemcpyDeviceToHost));
        fw.resize(nc * CMAXINT); rc.resize(nc * CMAXINT); ht.resize(nc * CMAXHITS); hp.resize(nc * CMAXHITS); hf.resize(nc * CMAXHITS);
        CG_CUDA(cudaMemcpy(fw.data(), dFwd.p, nc * CMAXINT * 16, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(rc.data(), dRc.p, nc * CMAXINT * 16, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(ht.data(), dHt.p, nc * CMAXHITS * 4, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(hp.data(), dHp.p, nc * CMAXHITS * 4, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(hf.data(), dHf.p, nc * CMAXHITS, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < nc; ++i) {
            if (nH[i] > (uint32_t)CMAXHITS) CG_FAIL(CG_ERR_CAPACITY, "hit list capacity exceeded");
            for (uint32_t x = 0; x < nF[i]; ++x) { const int4& v = fw[i * CMAXINT + x]; H.fwd_ints.insert(H.fwd_ints.end(), {v.x, v.y, v.z, v.w}); }
            for (uint32_t x = 0; x < nR[i]; ++x) { const int4& v = rc[i * CMAXINT + x]; H.rc_ints.insert(H.rc_ints.end(), {v.x, v.y, v.z, v.w}); }
            H.hit_tid.insert(H.hit_tid.end(), ht.begin() + i * CMAXHITS, ht.begin() + i * CMAXHITS + nH[i]);
            H.hit_pos.insert(H.hit_pos.end(), hp.begin() + i * CMAXHITS, hp.begin() + i * CMAXHITS + nH[i]);
            H.hit_fwd.insert(H.hit_fwd.end(), hf.begin() + i * CMAXHITS, hf.begin() + i * CMAXHITS + nH[i]);
            H.fwd_off.push_back(H.fwd_ints.size() / 4); H.rc_off.push_back(H.rc_ints.size() / 4); H.hits_off.push_back(H.hit_tid.size());
        }
    }
    out->n_reads = n_reads;
    out->found = H.found.data();
    out->fwd_off = H.fwd_off.data(); out->fwd_ints = H.fwd_ints.data();
    out->rc_off = H.rc_off.data(); out->rc_ints = H.rc_ints.data();
    out->hits_off = H.hits_off.data(); out->hit_tid = H.hit_tid.data(); out->hit_pos = H.hit_pos.data(); out->hit_fwd = H.hit_fwd.data();
    return CG_OK;
}

void cg_collect_free(void) { delete g_collect; g_collect = nullptr; }

int cg_merge_pairs(int device, uint64_t n_pairs, const uint64_t* left_off, const uint32_t* left_tid, const int32_t* left_pos,
                   const uint8_t* left_fwd, const uint8_t* left_seeded, const uint64_t* right_off, const uint32_t* right_tid,
                   const int32_t* right_pos, const uint8_t* right_fwd, const uint8_t* right_seeded, uint32_t read_len,
                   uint32_t max_num_hits, uint64_t* joint_off, int32_t* joint, uint64_t joint_cap, uint8_t* too_many) {
    if (!left_off || !right_off || !joint_off || !left_seeded || !right_seeded || !too_many) CG_FAIL(CG_ERR_ARG, "null argument");
    CG_CUDA(cudaSetDevice(device));
    uint64_t nl = left_off[n_pairs], nr = right_off[n_pairs];
    DevBuf lo, lt, lp, lf, ls, ro, rt, rp, rf, rs, cnt, jo, jt, tm;
    auto up = [](DevBuf& b, const void* src, uint64_t bytes) -> cudaError_t {
        cudaError_t e = b.alloc(bytes);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
        return e;
    };
    CG_CUDA(up(lo, left_off, (n_pairs + 1) * 8)); CG_CUDA(up(lt, left_tid, nl * 4)); CG_CUDA(up(lp, left_pos, nl * 4));
    CG_CUDA(up(lf, left_fwd, nl)); CG_CUDA(up(ls, left_seeded, n_pairs));
    CG_CUDA(up(ro, right_off, (n_pairs + 1) * 8)); CG_CUDA(up(rt, right_tid, nr * 4)); CG_CUDA(up(rp, right_pos, nr * 4));
    CG_CUDA(up(rf, right_fwd, nr)); CG_CUDA(up(rs, right_seeded, n_pairs));
    CG_CUDA(cnt.alloc((n_pairs + 1) * 8)); CG_CUDA(tm.alloc(n_pairs));
    MergeIn in{lo.as<uint64_t>(), lt.as<uint32_t>(), lp.as<int32_t>(), lf.as<uint8_t>(), ls.as<uint8_t>(),
               ro.as<uint64_t>(), rt.as<uint32_t>(), rp.as<int32_t>(), rf.as<uint8_t>(), rs.as<uint8_t>()};
    if (n_pairs) k_merge<<<grid_for(n_pairs, 256), 256>>>(in, n_pairs, read_len, max_num_hits, 0, cnt.as<uint64_t>(), nullptr, nullptr, nullptr);
    CG_CUDA(cudaGetLastError());
    std::vector<uint64_t> c(n_pairs + 1);
    if (n_pairs) CG_CUDA(cudaMemcpy(c.data(), cnt.p, n_pairs * 8, cudaMemcpyDeviceToHost));
    uint64_t tot = 0;
    for (uint64_t i = 0; i < n_pairs; ++i) { joint_off[i] = tot; tot += c[i]; }
    joint_off[n_pairs] = tot;
    if (tot > joint_cap) CG_FAIL(CG_ERR_CAPACITY, "joint-hit buffer too small (need " + std::to_string(tot) + " entries)");
    CG_CUDA(up(jo, joint_off, (n_pairs + 1) * 8));
    CG_CUDA(jt.alloc(tot * 28));
    if (n_pairs) k_merge<<<grid_for(n_pairs, 256), 256>>>(in, n_pairs, read_len, max_num_hits, 1, nullptr, jo.as<uint64_t>(), jt.as<int32_t>(), tm.as<uint8_t>());
    CG_CUDA(cudaGetLastError());
    if (tot) CG_CUDA(cudaMemcpy(joint, jt.p, tot * 28, cudaMemcpyDeviceToHost));
    if (n_pairs) CG_CUDA(cudaMemcpy(too_many, tm.p, n_pairs, cudaMemcpyDeviceToHost));
    return CG_OK;
}

int cg_bench_gather(int device, uint64_t bytes, uint32_t per_thread, int iters, double* gbps) {
    return cg_bench_gather_mode(device, bytes, per_thread, iters, 0, gbps);
}

int cg_bench_gather_mode(int device, uint64_t bytes, uint32_t per_thread, int iters, int mode, double* gbps) {
    if (mode < 0 || mode > 9) CG_FAIL(CG_ERR_ARG, "gather mode must be 0..9");
    if (!gbps || bytes < (1u << 20) || per_thread == 0 || iters <= 0) CG_FAIL(CG_ERR_ARG, "bad argument");
    CG_CUDA(cudaSetDevice(device));
    cgi::device_tuning(device);
    uint64_t ns = 1;
    while (ns * 2 * 32 <= bytes) ns *= 2;       // power-of-two number of 32-B sectors
    DevBuf buf, sink;
    CG_CUDA(buf.alloc(ns * 32)); CG_CUDA(sink.alloc(4));
    CG_CUDA(cudaMemset(buf.p, 1, ns * 32));
    per_thread = (per_thread + 3) & ~3u;
    const uint64_t threads = 148ull * 2048 * 4;
    cudaEvent_t e0, e1;
    CG_CUDA(cudaEventCreate(&e0)); CG_CUDA(cudaEventCreate(&e1));
    const gather_fn kern = gather_kernel(mode);
    kern<<<grid_for(threads, 256), 256>>>(buf.as<uint4>(), ns - 1, per_thread, 1, sink.as<uint32_t>());
    CG_CUDA(cudaDeviceSynchronize());
    CG_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) kern<<<grid_for(threads, 256), 256>>>(buf.as<uint4>(), ns - 1, per_thread, 2 + i, sink.as<uint32_t>());
    CG_CUDA(cudaEventRecord(e1));
    CG_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    CG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gbps = (double)threads * per_thread * iters * 32.0 / (ms * 1e-3) / 1e9;
    return CG_OK;
}

int cg_pack_reads(int device, const char* r1, const uint64_t* off1, const char* r2, const uint64_t* off2, uint64_t n_pairs,
                  int max_len, int general, int iters, uint64_t* packed, uint16_t* lens, int* status, double* ms_per_launch) {
    if (!off1 || !off2 || !packed || !lens || !status || max_len < 1 || max_len > 256 || iters < 1) CG_FAIL(CG_ERR_ARG, "bad argument");
    CG_CUDA(cudaSetDevice(device));
    PackGeom g = PackGeom::make(max_len);
    uint64_t nb1 = off1[n_pairs], nb2 = off2[n_pairs], n_reads = 2 * n_pairs;
    DevBuf d1, d2, o1, o2, pk, ln, st;
    CG_CUDA(d1.alloc(nb1 + 64)); CG_CUDA(d2.alloc(nb2 + 64));
    CG_CUDA(o1.alloc((n_pairs + 1) * 8)); CG_CUDA(o2.alloc((n_pairs + 1) * 8));
    CG_CUDA(pk.alloc(n_reads * g.WT * 8)); CG_CUDA(ln.alloc(n_reads * 2)); CG_CUDA(st.alloc(4));
    if (nb1) CG_CUDA(cudaMemcpy(d1.p, r1, nb1, cudaMemcpyHostToDevice));
    if (nb2) CG_CUDA(cudaMemcpy(d2.p, r2, nb2, cudaMemcpyHostToDevice));
    CG_CUDA(cudaMemcpy(o1.p, off1, (n_pairs + 1) * 8, cudaMemcpyHostToDevice));
    CG_CUDA(cudaMemcpy(o2.p, off2, (n_pairs + 1) * 8, cudaMemcpyHostToDevice));
    CG_CUDA(cudaMemset(st.p, 0, 4));
    cudaEvent_t e0, e1;
    CG_CUDA(cudaEventCreate(&e0)); CG_CUDA(cudaEventCreate(&e1));
    auto kern = general ? k_pack_general : k_pack;
    for (int it = 0; it <= iters && n_reads; ++it) {          // the first launch is the warm-up
        if (it == 1) CG_CUDA(cudaEventRecord(e0));
        kern<<<pack_grid(n_reads, g), PACK_THREADS>>>(d1.as<unsigned char>(), o1.as<uint64_t>(), d2.as<unsigned char>(), o2.as<uint64_t>(), n_pairs, g, max_len,
                                                       pk.as<uint64_t>(), ln.as<uint16_t>(), st.as<int>());
    }
    CG_CUDA(cudaGetLastError());
    float ms = 0;
    if (n_reads) {
        CG_CUDA(cudaEventRecord(e1));
        CG_CUDA(cudaEventSynchronize(e1));
        CG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_per_launch) *ms_per_launch = ms / iters;
    if (n_reads) {
        CG_CUDA(cudaMemcpy(packed, pk.p, n_reads * g.WT * 8, cudaMemcpyDeviceToHost));
        CG_CUDA(cudaMemcpy(lens, ln.p, n_reads * 2, cudaMemcpyDeviceToHost));
    }
    CG_CUDA(cudaMemcpy(status, st.p, 4, cudaMemcpyDeviceToHost));
    return CG_OK;
}

} // extern "C"
