// circall_b200 — device core of the quasi-mapping hot path (sm_100a).
//
// Everything here is written once as CG_HD functions: nvcc compiles them into the kernels of
// cg_kernels.cu; tests/ additionally compile this header with g++ (CG_HOST_EMU) to step through
// the exact device logic on a CPU box that has no GPU.  That emulation build is test scaffolding:
// it is not linked into libcircall_gpu.so and the library has no CPU path.
//
// What the functions replace (paths relative to /root/reference/Circall_v1.0.1/):
//   probe()            khash.find(mer.get_bits(0,2k))      include/SACollectorCircall.hpp:131-132,261,268,372,379,462,471
//   extend()           SASearcher::extendSearchNaive [U]    call sites :143-144,280-282,393-395  
//   lce()              SASearcher::lce [U]                  call sites :220,304,416               
//   collect()          SACollectorCircall::operator()       :24-582                               
//   hit_tids()         intersectSAHits + collectHitsSimpleSA [U]  :492-579                        
//
// Data layout:
//   text    32-byte blocks of 64 bases: c0,c1 = 2-bit codes (base j of a half at bits 2j), mask = '$'
//           bits, tid0 = transcriptAtPosition(block start), start0 = txpOffsets[tid0].  One 32-B
//           sector gives bases, terminators and the rank for any position inside it.
//   SA      uint32 suffix array.
//   khash   open addressing, 16-B slots {key(62-bit k-mer, base j at bits 2j), begin, end}.
//   reads   2-bit words (base j at bits 2j) + N-mask words of the forward strand, staged in shared
//           memory by the kernel prologue; the reverse-complement strand is derived by bit reversal.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__) && !defined(CG_HOST_EMU)
#define CG_HD __host__ __device__ __forceinline__
#define CG_HDN __host__ __device__ __noinline__
#else
#define CG_HD inline
#define CG_HDN inline
#endif

// Loads from the index (k-mer table, suffix array, text blocks, offsets): random addresses, read-only.  What the L2 pulls from DRAM
// for a missing 32-byte sector is set by the load's prefetch-size qualifier: a plain ld.global.nc fills the whole 128-byte line
// (130 B of DRAM traffic per load in the gather microbenchmark), .L2::64B fills 64 (65 B) at the same request rate
// (profiles/r2_experiments.md).  CG_LTC = 0 / 64 / 128 / 256 picks the qualifier at build time.
#ifndef CG_LTC
#define CG_LTC 64
#endif
#if defined(__CUDA_ARCH__)
#if CG_LTC == 64
#define CG_LTC_Q ".L2::64B"
#elif CG_LTC == 128
#define CG_LTC_Q ".L2::128B"
#elif CG_LTC == 256
#define CG_LTC_Q ".L2::256B"
#else
#define CG_LTC_Q ""
#endif
namespace cg {
__device__ __forceinline__ uint4 ldg_ix(const uint4* p) {
    uint4 v; asm("ld.global.nc" CG_LTC_Q ".v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ uint2 ldg_ix(const uint2* p) {
    uint2 v; asm("ld.global.nc" CG_LTC_Q ".v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v;
}
__device__ __forceinline__ uint32_t ldg_ix(const uint32_t* p) {
    uint32_t v; asm("ld.global.nc" CG_LTC_Q ".u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint8_t ldg_ix(const uint8_t* p) {
    uint32_t v; asm("ld.global.nc" CG_LTC_Q ".u8 %0, [%1];" : "=r"(v) : "l"(p)); return (uint8_t)v;
}
}
#define CG_LDG(p) cg::ldg_ix(p)
#else
#define CG_LDG(p) (*(p))
#endif

namespace cg {

#if defined(__CUDACC__) && !defined(CG_HOST_EMU)
typedef uint4 u32x4;
#else
struct u32x4 { uint32_t x, y, z, w; };
#endif

CG_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
CG_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
CG_HD int clz32(uint32_t x) {        // x != 0
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return __builtin_clz(x);
#endif
}
CG_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
CG_HD uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
template <class T> CG_HD T cmin(T a, T b) { return a < b ? a : b; }
template <class T> CG_HD T cmax(T a, T b) { return a > b ? a : b; }

static const uint64_t EVEN = 0x5555555555555555ULL;

// ---------------------------------------------------------------------------------------------
// index view
// ---------------------------------------------------------------------------------------------
struct IndexView {
    const u32x4* text;        // 2 x 16 B per 64-base block (+1 all-'$' pad block)
    const uint32_t* sa;
    const u32x4* hslots;      // {key_lo, key_hi, begin, end}; empty key = all ones
    uint64_t hmask;           // capacity - 1 (capacity is a power of two, >= 4 x the number of k-mers)
    int hshift;               // 64 - log2(capacity)
    const uint32_t* offsets;  // n_tx + 1 entries; offsets[n_tx] = n
    uint32_t n;               // text length (last character is '$')
    uint32_t n_tx;
    int k;
    uint64_t km;              // kmask(k)
    uint64_t km3;             // kmask(k) / 3: a homopolymer word is (w & 3) * km3
    // Seed filter (null / 0 = none): one bit per fm-mer code, set when the fm-mer or its reverse complement occurs '$'-free
    // in the text.  Every k-mer of the table is a '$'-free substring of the text, so a read window that contains an fm-mer
    // whose bit is clear — and its reverse complement, the set being closed under it — is certainly not in the table: one
    // random access settles the k - fm + 1 windows that contain the probed fm-mer (cg_thr.cuh::q_walk).
    const uint32_t* fbits;
    int fm;
    // transcriptAtPosition(SA[i]) for every suffix-array index i (null = none): the hit enumeration and the span filter walk SA
    // intervals, and with this array an interval's transcripts are one coalesced read instead of one random text block per suffix
    const uint32_t* satid;
    // min(255, longest common prefix of the suffixes SA[i - 1] and SA[i]) for every i >= 1 ('$' matches nothing; lcp[0] = 0; null = none).
    // extendSearchNaive compares the query with every suffix of a k-mer interval; the suffixes are sorted, so from the match length
    // of one suffix and this byte the match length of the next follows without touching the text, unless the two are equal
    // (cg_thr.cuh::q_step, cg_pre.cuh): one comparison per interval instead of one per suffix
    const uint8_t* lcp;
};

struct TB { uint64_t c0, c1, mask; uint32_t tid0, start0; };

CG_HD TB load_tb(const IndexView& ix, uint32_t b) {
    u32x4 a = CG_LDG(&ix.text[2 * (uint64_t)b]);
    u32x4 c = CG_LDG(&ix.text[2 * (uint64_t)b + 1]);
    TB t;
    t.c0 = (uint64_t)a.x | ((uint64_t)a.y << 32);
    t.c1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
    t.mask = (uint64_t)c.x | ((uint64_t)c.y << 32);
    t.tid0 = c.z; t.start0 = c.w;
    return t;
}

// transcriptAtPosition(p) and the transcript-local offset of p (rankDict + txpOffsets upstream)
CG_HD void tid_of(const IndexView& ix, uint32_t p, uint32_t& tid, uint32_t& local) {
    TB t = load_tb(ix, p >> 6);
    uint32_t o = p & 63;
    int c = popc64(o ? (t.mask & ((1ULL << o) - 1)) : 0ULL);
    tid = t.tid0 + (uint32_t)c;
    uint32_t start = c ? CG_LDG(&ix.offsets[tid]) : t.start0;
    local = p - start;
}

// SA[i0 .. i0+n) (n <= 4) -> transcript id and transcript-local offset of each position, with the
// suffix-array entries, then the text blocks, then the offsets fetched as three waves of independent loads
CG_HD void sa_tids4(const IndexView& ix, uint32_t i0, int n, uint32_t* tid, uint32_t* local) {
    if (ix.satid) {
        uint32_t p[4], st[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { p[x] = x < n ? CG_LDG(&ix.sa[i0 + x]) : 0u; tid[x] = x < n ? CG_LDG(&ix.satid[i0 + x]) : 0u; }
#pragma unroll
        for (int x = 0; x < 4; ++x) st[x] = x < n ? CG_LDG(&ix.offsets[tid[x]]) : 0u;
#pragma unroll
        for (int x = 0; x < 4; ++x) if (x < n) local[x] = p[x] - st[x];
        return;
    }
    uint32_t p[4]; TB t[4]; int c[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) p[x] = x < n ? CG_LDG(&ix.sa[i0 + x]) : 0u;
#pragma unroll
    for (int x = 0; x < 4; ++x) if (x < n) t[x] = load_tb(ix, p[x] >> 6);
#pragma unroll
    for (int x = 0; x < 4; ++x)
        if (x < n) {
            uint32_t o = p[x] & 63;
            c[x] = popc64(o ? (t[x].mask & ((1ULL << o) - 1)) : 0ULL);
            tid[x] = t[x].tid0 + (uint32_t)c[x];
        }
    uint32_t st[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) if (x < n) st[x] = c[x] ? CG_LDG(&ix.offsets[tid[x]]) : t[x].start0;
#pragma unroll
    for (int x = 0; x < 4; ++x) if (x < n) local[x] = p[x] - st[x];
}

// ---------------------------------------------------------------------------------------------
// packed read view.  Only the forward strand is stored: w[0..W2) code words (32 bases each, base j
// at bits 2j, last word a zero pad) then w[W2..W2+WM) N-mask words (64 positions each, last word an
// all-ones pad).  Positions >= L carry code 0 and mask 1.  The reverse-complement strand (s = 1,
// oriented position i = distance from the read's end) is derived on the fly by bit reversal.
// ---------------------------------------------------------------------------------------------
CG_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    return (uint32_t)(brev64((uint64_t)x) >> 32);
#endif
}
// reverse the order of the 32 2-bit codes of x and complement them
CG_HD uint64_t rc64(uint64_t x) {
    uint64_t r = brev64(x);
    return ~(((r >> 1) & EVEN) | ((r & EVEN) << 1));
}

struct ReadView {
    const uint64_t* w;
    int L;
    int W2;    // code words incl. the zero pad  = ceil(maxL/32) + 1
    int WM;    // mask words incl. the ones pad  = ceil(W2f/2) + 1
    bool anyN; // some position < L is not a base (false lets every N test return at once)
    CG_HD uint64_t code(int j) const { return w[j]; }
    CG_HD uint64_t msk(int j) const { return w[W2 + j]; }
    CG_HD uint64_t chunkF(int i) const {
        int j = i >> 5, sh = (i & 31) * 2;
        uint64_t lo = code(j);
        if (sh == 0) return lo;
        return (lo >> sh) | (code(j + 1) << (64 - sh));
    }
    CG_HD uint32_t mask32F(int i) const {
        int j = i >> 6, sh = i & 63;
        uint64_t lo = msk(j);
        if (sh == 0) return (uint32_t)lo;
        return (uint32_t)((lo >> sh) | (msk(j + 1) << (64 - sh)));
    }
    // 32 bases (64 bits) of strand s starting at oriented position i (i <= L)
    CG_HD uint64_t chunk(int s, int i) const {
        if (s == 0) return chunkF(i);
        int a = L - 32 - i;                       // forward window [a, a+32) read backwards
        if (a >= 0) return rc64(chunkF(a));
        int nv = 32 + a;
        if (nv <= 0) return 0;
        return rc64(chunkF(0)) >> (2 * (32 - nv));
    }
    // 32 N-mask bits of strand s starting at oriented position i
    CG_HD uint32_t mask32(int s, int i) const {
        if (s == 0) return mask32F(i);
        int a = L - 32 - i;
        if (a >= 0) return brev32(mask32F(a));
        int nv = 32 + a;
        if (nv <= 0) return 0xffffffffu;
        return (brev32(mask32F(0)) >> (32 - nv)) | (0xffffffffu << nv);
    }
    // first n/N at or after oriented position i on strand s (L when there is none)
    CG_HD int first_n(int s, int i) const {
        if (!anyN) return L;
        for (int p = i; p < L; p += 32) {
            uint32_t m = mask32(s, p);
            if (m) { int r = p + ctz32(m); return r < L ? r : L; }
        }
        return L;
    }
};

// words a view needs for reads of up to maxL bases
CG_HD int view_words(int W2f, int WMf) { return (W2f + 1) + (WMf + 1); }

// Copy one packed forward read (W2f code words then WMf mask words, word j at pk[j * pkstride]) into
// the padded layout ReadView expects.
CG_HD ReadView build_view(const uint64_t* pk, size_t pkstride, int L, int W2f, int WMf, uint64_t* dst) {
    const int W2 = W2f + 1, WM = WMf + 1;
    for (int j = 0; j < W2f; ++j) dst[j] = pk[(size_t)j * pkstride];
    dst[W2f] = 0;
    for (int j = 0; j < WMf; ++j) dst[W2 + j] = pk[(size_t)(W2f + j) * pkstride];
    dst[W2 + WMf] = ~0ULL;
    ReadView rv; rv.w = dst; rv.L = L; rv.W2 = W2; rv.WM = WM;
    rv.anyN = false;
    for (int j = 0; j < WMf; ++j) {
        int lo = 64 * j;
        if (lo >= L) break;
        uint64_t m = dst[W2 + j];
        if (L - lo < 64) m &= (1ULL << (L - lo)) - 1;
        if (m) rv.anyN = true;
    }
    return rv;
}

CG_HD uint64_t kmask(int k) { return (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1); }

CG_HD uint64_t revcomp(uint64_t w, int k) { return (rc64(w) >> (64 - 2 * k)) & kmask(k); }
CG_HD bool homopolymer(uint64_t w, int k) { return w == (w & 3) * (kmask(k) / 3); }
CG_HD bool homopolymer_ix(uint64_t w, const IndexView& ix) { return w == (w & 3) * ix.km3; }

// forward-strand k-mer word at pos; when the window holds a non-ACGT base the word is cut at the
// first one and the rest is zero ('A'), which is how the oracle defines Jellyfish's
// mer_dna(const char*) on such input (reached from SACollectorCircall.hpp:167,293,406)
CG_HD uint64_t kmer_trunc(const ReadView& rv, int pos, int k) {
    uint64_t w = rv.chunkF(pos) & kmask(k);
    uint32_t m = rv.mask32F(pos) & (uint32_t)((1ULL << k) - 1);
    if (m) w &= (1ULL << (2 * ctz32(m))) - 1;
    return w;
}

// ---------------------------------------------------------------------------------------------
// packing primitives shared by the read-pack kernel, the text-pack kernel and the k-mer table build
// ---------------------------------------------------------------------------------------------
// 0..3 = A C G T (either case), 4 = n/N, 5 = '$', 6 = anything else
CG_HD int base_class(unsigned char c) {
    unsigned char u = c & 0xDF;
    if (u == 'A' || u == 'C' || u == 'G' || u == 'T') return (int)(((c >> 1) ^ (c >> 2)) & 3);
    if (u == 'N') return 4;
    if (c == '$') return 5;
    return 6;
}
// up to 32 read bases -> 2-bit code word + N-mask bits; positions >= nb get code 0 / mask 1
CG_HD void pack_word(const unsigned char* s, int nb, uint64_t& code, uint32_t& nmask, bool& bad) {
    code = 0; nmask = 0; bad = false;
    for (int i = 0; i < 32; ++i) {
        if (i < nb) {
            int c = base_class(s[i]);
            if (c < 4) code |= (uint64_t)c << (2 * i);
            else { nmask |= 1u << i; if (c != 4) bad = true; }
        } else nmask |= 1u << i;
    }
}
// 64 text characters -> one text block's codes and terminator mask; positions >= nv count as '$'
CG_HD void make_block(const unsigned char* s, int nv, uint64_t& c0, uint64_t& c1, uint64_t& mask, bool& bad) {
    c0 = c1 = mask = 0; bad = false;
    for (int i = 0; i < 64; ++i) {
        int c = i < nv ? base_class(s[i]) : 5;
        if (c < 4) { if (i < 32) c0 |= (uint64_t)c << (2 * i); else c1 |= (uint64_t)c << (2 * (i - 32)); }
        else { mask |= 1ULL << i; if (c != 5) bad = true; }
    }
}
// word of text[p, p+k); false when the window holds a '$' or runs past the text
CG_HD bool text_word(const IndexView& ix, uint32_t p, int k, uint64_t& w) {
    if ((uint64_t)p + (uint64_t)k > (uint64_t)ix.n) return false;
    w = 0;
    int i = 0;
    while (i < k) {
        uint32_t tp = p + (uint32_t)i;
        TB tb = load_tb(ix, tp >> 6);
        int half = (tp >> 5) & 1, o5 = tp & 31;
        int n = cmin(32 - o5, k - i);
        uint64_t tw = (half ? tb.c1 : tb.c0) >> (2 * o5);
        uint32_t tm = (uint32_t)(tb.mask >> (32 * half + o5));
        if (n < 32) { tw &= (1ULL << (2 * n)) - 1; tm &= (1u << n) - 1; }
        if (tm) return false;
        w |= tw << (2 * i);
        i += n;
    }
    return true;
}

CG_HD bool text_kmer(const IndexView& ix, uint32_t p, uint64_t& w) { return text_word(ix, p, ix.k, w); }

// seed filter: m-mer codes are 2 m bits (m <= 16), base j at bits 2j like the k-mer words
CG_HD uint32_t fm_mask(int m) { return m >= 16 ? 0xffffffffu : ((1u << (2 * m)) - 1u); }
CG_HD uint32_t fm_revcomp(uint32_t x, int m) {
    uint32_t r = brev32(x);
    r = ~(((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1));
    return r >> (32 - 2 * m);
}
CG_HD bool fm_homopolymer(uint32_t x, int m) { return x == (x & 3u) * (fm_mask(m) / 3u); }
// m of an index with k-mers of k bases: 16 (a 512 MB bit set) when that leaves the 3-probe stage of the walk inside 64 windows
CG_HD int fm_for_k(int k, int want) { return (want >= 8 && want <= 16 && want < k && 3 * (k - want) + 1 < 64) ? want : 0; }

// ---------------------------------------------------------------------------------------------
// k-mer table probe
// ---------------------------------------------------------------------------------------------
CG_HD uint64_t hmix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// home slot of a k-mer: multiplicative hash (one 64-bit multiply), forced to the even slot of a 32-byte
// sector so that the first two slots a look-up examines cost one memory sector
CG_HD uint64_t home_of(uint64_t key, int hshift) { return ((key * 0x9E3779B97F4A7C15ULL) >> hshift) & ~1ULL; }
CG_HD uint64_t home_slot(const IndexView& ix, uint64_t key) { return home_of(key, ix.hshift); }
// finish a probe whose home slot `s` (at index h) has already been loaded
CG_HD bool resolve(const IndexView& ix, uint64_t key, uint64_t h, u32x4 s, uint32_t& b, uint32_t& e) {
    for (;;) {
        uint64_t kk = (uint64_t)s.x | ((uint64_t)s.y << 32);
        if (kk == key) { b = s.z; e = s.w; return true; }
        if (kk == ~0ULL) return false;
        h = (h + 1) & ix.hmask;
        s = CG_LDG(&ix.hslots[h]);
    }
}
CG_HD bool probe(const IndexView& ix, uint64_t key, uint32_t& b, uint32_t& e) {
    uint64_t h = home_slot(ix, key);
    return resolve(ix, key, h, CG_LDG(&ix.hslots[h]), b, e);
}
// two probes with both home-slot loads in flight together
CG_HD void probe2(const IndexView& ix, uint64_t k1, uint64_t k2, bool& f1, uint32_t& b1, uint32_t& e1,
                  bool& f2, uint32_t& b2, uint32_t& e2) {
    uint64_t h1 = home_slot(ix, k1), h2 = home_slot(ix, k2);
    u32x4 s1 = CG_LDG(&ix.hslots[h1]);
    u32x4 s2 = CG_LDG(&ix.hslots[h2]);
    f1 = resolve(ix, k1, h1, s1, b1, e1);
    f2 = resolve(ix, k2, h2, s2, b2, e2);
}
CG_HD bool present(const IndexView& ix, uint64_t key) { uint32_t b, e; return probe(ix, key, b, e); }

// ---------------------------------------------------------------------------------------------
// query-vs-suffix comparison
// ---------------------------------------------------------------------------------------------
// Advance i while query[qstart+i] == text[p+i], i < lim.  On a stop before lim, `ord` tells how
// the query character sorts against the text character by raw char value
// ('$' < A < C < G < N < T): -1 query smaller, +1 query larger.
struct TB2 { uint64_t c0, c1, mask; };     // codes + terminator mask of one text block (what comparisons need)
CG_HD TB2 load_tb2(const IndexView& ix, uint32_t b) {
    u32x4 a = CG_LDG(&ix.text[2 * (uint64_t)b]);
    u32x4 c = CG_LDG(&ix.text[2 * (uint64_t)b + 1]);
    TB2 t;
    t.c0 = (uint64_t)a.x | ((uint64_t)a.y << 32);
    t.c1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
    t.mask = (uint64_t)c.x | ((uint64_t)c.y << 32);
    return t;
}
// b0 / b1: text blocks already in registers (block index pb0 and pb0 + 1); pass pb0 = 0xffffffff for none
CG_HD int match_run_pre(const IndexView& ix, const ReadView& rv, int s, int qstart, uint32_t p, int i, int lim, int& ord,
                        uint32_t pb0, const TB2& b0, const TB2& b1) {
    uint32_t curb = 0xffffffffu;
    TB2 tb; tb.c0 = tb.c1 = tb.mask = 0;
    ord = 0;
    while (i < lim) {
        uint32_t tp = p + (uint32_t)i;
        uint32_t b = tp >> 6;
        if (b != curb) {
            if (b == pb0) tb = b0;
            else if (b == pb0 + 1 && pb0 != 0xffffffffu) tb = b1;
            else tb = load_tb2(ix, b);
            curb = b;
        }
        int half = (tp >> 5) & 1, o5 = tp & 31;
        uint64_t tw = (half ? tb.c1 : tb.c0) >> (2 * o5);
        uint32_t tm = (uint32_t)(tb.mask >> (32 * half + o5));
        int n = cmin(32 - o5, lim - i);
        uint64_t q = rv.chunk(s, qstart + i);
        uint32_t qm = rv.mask32(s, qstart + i);
        uint64_t x = tw ^ q;
        uint64_t d = (x | (x >> 1)) & EVEN;
        int j1 = d ? (ctz64(d) >> 1) : 32;
        uint32_t mm = tm | qm;
        int j2 = mm ? ctz32(mm) : 32;
        int j = cmin(j1, j2);
        if (j < n) {
            i += j;
            bool tdollar = (tm >> j) & 1, qn = (qm >> j) & 1;
            int tc = (int)((tw >> (2 * j)) & 3), qc = (int)((q >> (2 * j)) & 3);
            if (tdollar) ord = 1;
            else if (qn) ord = (tc == 3) ? -1 : 1;
            else ord = qc < tc ? -1 : 1;
            return i;
        }
        i += n;
    }
    return i;
}
CG_HD int match_run(const IndexView& ix, const ReadView& rv, int s, int qstart, uint32_t p, int i, int lim, int& ord) {
    uint32_t curb = 0xffffffffu;
    TB2 tb; tb.c0 = tb.c1 = tb.mask = 0;
    ord = 0;
    while (i < lim) {
        uint32_t tp = p + (uint32_t)i;
        uint32_t b = tp >> 6;
        if (b != curb) { tb = load_tb2(ix, b); curb = b; }
        int half = (tp >> 5) & 1, o5 = tp & 31;
        uint64_t tw = (half ? tb.c1 : tb.c0) >> (2 * o5);
        uint32_t tm = (uint32_t)(tb.mask >> (32 * half + o5));
        int n = cmin(32 - o5, lim - i);
        uint64_t q = rv.chunk(s, qstart + i);
        uint32_t qm = rv.anyN ? rv.mask32(s, qstart + i) : 0u;
        uint64_t x = tw ^ q;
        uint64_t d = (x | (x >> 1)) & EVEN;
        int j1 = d ? (ctz64(d) >> 1) : 32;
        uint32_t mm = tm | qm;
        int j2 = mm ? ctz32(mm) : 32;
        int j = cmin(j1, j2);
        if (j < n) {
            i += j;
            bool tdollar = (tm >> j) & 1, qn = (qm >> j) & 1;
            int tc = (int)((tw >> (2 * j)) & 3), qc = (int)((q >> (2 * j)) & 3);
            if (tdollar) ord = 1;
            else if (qn) ord = (tc == 3) ? -1 : 1;
            else ord = qc < tc ? -1 : 1;
            return i;
        }
        i += n;
    }
    return i;
}

// extendSearchNaive: among SA indices in the open interval (lbIn, ubIn) — all of which share
// query[0:startAt] — find the maximum-mappable-prefix length maxI and the sub-interval sharing
// query[0:maxI].  Narrow intervals are resolved by evaluating every candidate (independent
// gathers, two dependent memory steps in total); wide ones by the reference's three binary searches.
#ifndef CG_LINEAR_MAX
#define CG_LINEAR_MAX 8
#endif
CG_HD void extend(const IndexView& ix, const ReadView& rv, int s, int qstart, uint32_t lbIn, uint32_t ubIn,
                  int startAt, uint32_t& lbOut, uint32_t& ubOut, int& lenOut) {
    const int m = rv.L - qstart;
    if (m <= startAt) { lbOut = lbIn + 1; ubOut = ubIn; lenOut = startAt; return; }
    const uint32_t w = ubIn - lbIn - 1;
    int ord;
    if (w == 0) { lbOut = ubIn; ubOut = ubIn; lenOut = startAt; return; }     // empty open interval (reference: lower = upper = ubIn)
    if (w <= CG_LINEAR_MAX) {
        int best = -1; uint32_t first = 0, last = 0;
        const uint32_t nblk = (uint32_t)((ix.n + 63) >> 6);       // index of the pad block
        for (uint32_t g0 = lbIn + 1; g0 < ubIn; g0 += 4) {
            uint32_t p[4]; uint32_t pb[4]; TB2 t0[4], t1[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) p[c] = (g0 + c < ubIn) ? CG_LDG(&ix.sa[g0 + c]) : 0u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                pb[c] = 0xffffffffu; t0[c].c0 = t0[c].c1 = t0[c].mask = 0; t1[c] = t0[c];
                if (g0 + c < ubIn) {
                    pb[c] = (p[c] + (uint32_t)startAt) >> 6;
                    t0[c] = load_tb2(ix, pb[c]);
                    // the second block only when the compared range [p+startAt, p+m) reaches into it
                    if ((int)((p[c] + (uint32_t)startAt) & 63) + (m - startAt) > 64)
                        t1[c] = load_tb2(ix, pb[c] + 1 <= nblk ? pb[c] + 1 : nblk);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (g0 + c < ubIn) {
                    int l = match_run_pre(ix, rv, s, qstart, p[c], startAt, m, ord, pb[c], t0[c], t1[c]);
                    if (l > best) { best = l; first = g0 + c; last = g0 + c; }
                    else if (l == best) last = g0 + c;
                }
            }
        }
        lbOut = first; ubOut = last + 1; lenOut = best;
        return;
    }
    uint32_t lb = lbIn, ub = ubIn;
    int prevLow = startAt, prevHigh = startAt, maxI = startAt;
    while (ub - lb > 1) {
        uint32_t c = lb + (ub - lb) / 2;
        uint32_t p = CG_LDG(&ix.sa[c]);
        int i = match_run(ix, rv, s, qstart, p, cmin(prevLow, prevHigh), m, ord);
        if (i > maxI) maxI = i;
        if (i >= m || ord < 0) { ub = c; prevHigh = i; } else { lb = c; prevLow = i; }
    }
    // lower bound: first c whose suffix is >= query[0:maxI]
    uint32_t lo = lbIn, hi = ubIn;
    while (hi - lo > 1) {
        uint32_t c = lo + (hi - lo) / 2;
        uint32_t p = CG_LDG(&ix.sa[c]);
        int i = match_run(ix, rv, s, qstart, p, startAt, maxI, ord);
        // i >= maxI: shares the prefix; else ord>0 means query larger => suffix smaller
        if (i >= maxI || ord < 0) hi = c; else lo = c;
    }
    uint32_t lower = hi;
    lo = lower - 1; hi = ubIn;
    while (hi - lo > 1) {
        uint32_t c = lo + (hi - lo) / 2;
        uint32_t p = CG_LDG(&ix.sa[c]);
        int i = match_run(ix, rv, s, qstart, p, startAt, maxI, ord);
        if (i < maxI && ord < 0) hi = c; else lo = c;
    }
    lbOut = lower; ubOut = hi; lenOut = maxI;
}

// lce.  mode 0 = LCE_LITERAL: startAt applied twice; result =
// min(first len >= startAt where the characters differ / are '$' / the text ends, max(startAt, stopAt)).
CG_HD int lce(const IndexView& ix, uint32_t p1, uint32_t p2, int startAt, int stopAt, int mode) {
    if (stopAt <= startAt) return startAt;      // the loop can only stop at len == startAt
    uint32_t o1 = CG_LDG(&ix.sa[p1]), o2 = (p2 == p1) ? o1 : CG_LDG(&ix.sa[p2]);
    if (mode == 0) { o1 += (uint32_t)startAt; o2 += (uint32_t)startAt; }
    int len = startAt;
    uint32_t mx = cmax(o1, o2);
    int64_t room = (int64_t)ix.n - (int64_t)mx;          // maxIndex + len < textLen
    int lim = (int)cmin<int64_t>((int64_t)stopAt, room);
    uint32_t b1c = 0xffffffffu, b2c = 0xffffffffu;
    TB t1, t2; t1.c0 = t1.c1 = t1.mask = 0; t2 = t1; t1.tid0 = t1.start0 = t2.tid0 = t2.start0 = 0;
    while (len < lim) {
        uint32_t a = o1 + (uint32_t)len, b = o2 + (uint32_t)len;
        if ((a >> 6) != b1c) { b1c = a >> 6; t1 = load_tb(ix, b1c); }
        if ((b >> 6) != b2c) { b2c = b >> 6; t2 = (b2c == b1c) ? t1 : load_tb(ix, b2c); }
        int ha = (a >> 5) & 1, oa = a & 31, hb = (b >> 5) & 1, ob = b & 31;
        uint64_t wa = (ha ? t1.c1 : t1.c0) >> (2 * oa), wb = (hb ? t2.c1 : t2.c0) >> (2 * ob);
        uint32_t ma = (uint32_t)(t1.mask >> (32 * ha + oa)), mb = (uint32_t)(t2.mask >> (32 * hb + ob));
        int n = cmin(cmin(32 - oa, 32 - ob), lim - len);
        uint64_t x = wa ^ wb;
        uint64_t d = (x | (x >> 1)) & EVEN;
        int j1 = d ? (ctz64(d) >> 1) : 32;
        uint32_t mm = ma | mb;                      // differ-by-terminator or both '$': stop either way
        int j2 = mm ? ctz32(mm) : 32;
        int j = cmin(j1, j2);
        if (j < n) return len + j;
        len += n;
    }
    return len;
}

// ---------------------------------------------------------------------------------------------
// collector state
// ---------------------------------------------------------------------------------------------
struct Interval { uint32_t begin, end; uint16_t len, qpos; };

// kmerScores entry: kpos << 8 | (fwdScore+1) << 2 | (rcScore+1), scores in {-1 ABSENT, 0 UNTESTED, 1 PRESENT}
CG_HD uint32_t ks_pack(int kpos, int f, int r) { return ((uint32_t)kpos << 8) | ((uint32_t)(f + 1) << 2) | (uint32_t)(r + 1); }
CG_HD int ks_f(uint32_t e) { return (int)((e >> 2) & 3) - 1; }
CG_HD int ks_r(uint32_t e) { return (int)(e & 3) - 1; }
CG_HD uint32_t ks_key(uint32_t e) { return e >> 8; }

// libstdc++ std::sort(first, last) with operator< on kpos only, reproduced step for step so that
// the element surviving std::unique among equal kpos is the one the reference keeps
// (SACollectorCircall.hpp:450-451; SURVEY §7 hard part 2): introsort with depth limit 2*floor(log2 n),
// median-of-three to first, unguarded partition, heap-sort fallback, threshold 16, final insertion sort.
struct StdSort {
    uint32_t* a;
    CG_HD bool lt(uint32_t x, uint32_t y) const { return ks_key(x) < ks_key(y); }
    CG_HD void swp(int i, int j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }
    CG_HD void push_heap(int first, int hole, int top, uint32_t v) {
        int parent = (hole - 1) / 2;
        while (hole > top && lt(a[first + parent], v)) { a[first + hole] = a[first + parent]; hole = parent; parent = (hole - 1) / 2; }
        a[first + hole] = v;
    }
    CG_HD void adjust_heap(int first, int hole, int len, uint32_t v) {
        int top = hole, child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (lt(a[first + child], a[first + child - 1])) --child;
            a[first + hole] = a[first + child]; hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[first + hole] = a[first + child - 1]; hole = child - 1;
        }
        push_heap(first, hole, top, v);
    }
    CG_HD void heap_sort(int first, int last) {      // __partial_sort(first, last, last)
        int len = last - first;
        if (len >= 2) {
            int parent = (len - 2) / 2;
            for (;;) { uint32_t v = a[first + parent]; adjust_heap(first, parent, len, v); if (parent == 0) break; --parent; }
        }
        while (last - first > 1) {
            --last;
            uint32_t v = a[last]; a[last] = a[first];
            adjust_heap(first, 0, last - first, v);
        }
    }
    CG_HD void median_to_first(int r, int x, int y, int z) {
        if (lt(a[x], a[y])) {
            if (lt(a[y], a[z])) swp(r, y); else if (lt(a[x], a[z])) swp(r, z); else swp(r, x);
        } else if (lt(a[x], a[z])) swp(r, x);
        else if (lt(a[y], a[z])) swp(r, z);
        else swp(r, y);
    }
    CG_HD int partition(int first, int last, int pivot) {
        for (;;) {
            while (lt(a[first], a[pivot])) ++first;
            --last;
            while (lt(a[pivot], a[last])) --last;
            if (!(first < last)) return first;
            swp(first, last);
            ++first;
        }
    }
    CG_HD void linear_insert(int last) {
        uint32_t v = a[last]; int next = last - 1;
        while (lt(v, a[next])) { a[last] = a[next]; last = next; --next; }
        a[last] = v;
    }
    CG_HD void insertion(int first, int last) {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (lt(a[i], a[first])) { uint32_t v = a[i]; for (int j = i; j > first; --j) a[j] = a[j - 1]; a[first] = v; }
            else linear_insert(i);
        }
    }
    // BIG = false: the caller guarantees n <= 16 (plain insertion sort, no introsort stack)
    template <bool BIG>
    CG_HD void sort(int n) {
        if (n <= 1) return;
        if (BIG && n > 16) {
            int lg = 0; for (int t = n; t > 1; t >>= 1) ++lg;
            // explicit stack for __introsort_loop(cut, last, depth) recursion on the right part
            int stF[48], stL[48], stD[48]; int sp = 0;
            stF[0] = 0; stL[0] = n; stD[0] = 2 * lg; sp = 1;
            while (sp > 0) {
                --sp;
                int first = stF[sp], last = stL[sp], depth = stD[sp];
                // the recursive call on [cut,last) runs to completion before the loop continues on
                // [first,cut); the two ranges are disjoint, so only the order of work inside each
                // matters and a stack that handles the left part later gives identical results
                while (last - first > 16) {
                    if (depth == 0) { heap_sort(first, last); break; }
                    --depth;
                    int mid = first + (last - first) / 2;
                    median_to_first(first, first + 1, mid, last - 1);
                    int cut = partition(first + 1, last, first);
                    stF[sp] = cut; stL[sp] = last; stD[sp] = depth; ++sp;
                    last = cut;
                }
            }
            insertion(0, 16);
            for (int i = 16; i != n; ++i) linear_insert(i);
        } else {
            insertion(0, n);
        }
    }
};

template <int MAXINT, int MAXKS>
struct CollectState {
    Interval fwd[MAXINT];
    Interval rc[MAXINT];
    uint32_t ks[MAXKS];
    int nF, nR, nKS;
    int err;      // non-zero when a fixed capacity was exceeded (reported, never silently dropped)
};

enum { CG_ERR_INTERVALS = 1, CG_ERR_KSCORES = 2, CG_ERR_HITS = 4 };

// SACollectorCircall::operator() up to and including the strand vote.  Returns foundHit.
// After it, st.fwd[0..nF) / st.rc[0..nR) are the post-vote fwdSAInts / rcSAInts.
template <int MAXINT, int MAXKS>
CG_HD bool collect(const IndexView& ix, const ReadView& rv, bool strict, int lceMode, CollectState<MAXINT, MAXKS>& st) {
    const int L = rv.L, k = ix.k;
    const int skipOverlap = k - 1, homoSkip = k / 2;
    const uint32_t maxInterval = 1000;
    const uint64_t KM = kmask(k);
    st.nF = st.nR = st.nKS = 0; st.err = 0;
    uint32_t fwdHit = 0, rcHit = 0;
    bool found = false;
    uint32_t lbF = 0, ubF = 0;
    int matched = 0;

#define CG_PUSH_KS(kp, f, r) do { if (st.nKS < MAXKS) st.ks[st.nKS++] = ks_pack((kp), (f), (r)); else st.err |= CG_ERR_KSCORES; } while (0)
#define CG_PUSH_INT(arr, cnt, b_, e_, l_, q_) do { if (cnt < MAXINT) { arr[cnt].begin = (b_); arr[cnt].end = (e_); arr[cnt].len = (uint16_t)(l_); arr[cnt].qpos = (uint16_t)(q_); ++cnt; } else st.err |= CG_ERR_INTERVALS; } while (0)

    // ---- Phase A (:112-200)
    int rb = 0;
    bool preA = false, preF = false, preR = false; uint32_t pb_ = 0, pe_ = 0, prb_ = 0, pre_ = 0;
    while (rb + k < L) {
        int inv = rv.first_n(0, rb);
        if (inv < L && inv <= rb + k) { rb = inv + 1; preA = false; continue; }    // :118
        uint64_t mer = rv.chunk(0, rb) & KM;
        if (homopolymer(mer, k)) { rb += homoSkip; preA = false; continue; }        // :127
        uint64_t rcm = revcomp(mer, k);
        uint32_t b = pb_, e = pe_, rb_ = prb_, re_ = pre_;
        bool hasF = preF, hasR = preR;
        if (!preA) probe2(ix, mer, rcm, hasF, b, e, hasR, rb_, re_);                // :131-132
        preA = false;
        if (hasF) {
            uint32_t lbRestart = b > 0 ? b - 1 : 0;                                 // :140
            extend(ix, rv, 0, rb, lbRestart, e, k, lbF, ubF, matched);              // :143
            if (ubF > lbF && ubF - lbF < maxInterval) {                             // :149
                CG_PUSH_INT(st.fwd, st.nF, lbF, ubF, matched, rb);
                ++fwdHit;
                if (strict) {
                    if (hasR) { ++rcHit; CG_PUSH_KS(rb, 1, 1); } else { CG_PUSH_KS(rb, 1, -1); }
                    if (rb + matched < L) CG_PUSH_KS(rb + matched - skipOverlap, -1, 0);   // :165-168
                } else if (hasR) ++rcHit;
            }
        }
        if (hasR && re_ > rb_ && !fwdHit) {                                         // :178-191
            ++rcHit;
            if (strict) CG_PUSH_KS(rb, -1, 1);
        }
        if (fwdHit + rcHit > 0) { found = true; break; }
        ++rb;
        if (hasF) continue;     // the table holds the forward k-mer (its interval was too wide): plain step
        // neither k-mer is in the table: the loop now walks rb+1, rb+2, ... whatever the look-ups return,
        // so the next two positions (four look-ups) go out together
        for (;;) {
            int q[2]; uint64_t kf[2], kr[2], hf[2], hr[2]; u32x4 sf[2], sr[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                q[c] = -1; kf[c] = kr[c] = hf[c] = hr[c] = 0;
                sf[c].x = sf[c].y = sf[c].z = sf[c].w = 0; sr[c] = sf[c];
                while (rb + k < L) {
                    int iv = rv.first_n(0, rb);
                    if (iv < L && iv <= rb + k) { rb = iv + 1; continue; }
                    uint64_t kw = rv.chunk(0, rb) & KM;
                    if (homopolymer(kw, k)) { rb += homoSkip; continue; }
                    q[c] = rb; kf[c] = kw; kr[c] = revcomp(kw, k);
                    hf[c] = home_slot(ix, kf[c]); hr[c] = home_slot(ix, kr[c]);
                    sf[c] = CG_LDG(&ix.hslots[hf[c]]); sr[c] = CG_LDG(&ix.hslots[hr[c]]);
                    ++rb;
                    break;
                }
            }
            int got = -1;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (got < 0 && q[c] >= 0) {
                    uint32_t b1 = 0, e1 = 0, b2 = 0, e2 = 0;
                    bool f1 = resolve(ix, kf[c], hf[c], sf[c], b1, e1);
                    bool f2 = resolve(ix, kr[c], hr[c], sr[c], b2, e2);
                    if (f1 || f2) { got = q[c]; preF = f1; preR = f2; pb_ = b1; pe_ = e1; prb_ = b2; pre_ = e2; }
                }
            }
            if (got >= 0) { rb = got; preA = true; break; }
            if (q[1] < 0) break;                // ran off the read
        }
    }
    if (!found) return false;                                                       // :204

    // ---- Phases B (:208-329, forward strand) and C (:332-440, reverse-complement strand) share one
    //      loop body; `s` is the strand being scanned and `o` the position on that strand.
    for (int s = 0; s < 2; ++s) {
        int o;
        bool lastSearch = false;
        if (s == 0) {
            if (!fwdHit) continue;
            int matchLen = st.fwd[0].len;
            int q0 = st.fwd[0].qpos;
            int remaining = L - (q0 + matchLen);                                    // :219
            int l = lce(ix, lbF, ubF - 1, matchLen, remaining, lceMode);            // :220
            int fwdSkip = cmax(matchLen - skipOverlap, l - skipOverlap);
            o = cmin(cmax(0, L - k), q0 + fwdSkip);                                 // :224-228
        } else {
            if (!(rcHit >= fwdHit)) continue;                                       // :332
            o = 0;
        }
        bool pre = false; uint32_t preB = 0, preE = 0;     // k-mer at o already looked up by the scan-ahead below
        while (o + k <= L) {                                                        // :234 / :341-344
            int inv = rv.first_n(s, o);
            if (inv < o + k) { o = inv + 1; pre = false; continue; }               // :247 / :357 (inv < L here)
            uint64_t own = rv.chunk(s, o) & KM;         // k-mer as read on this strand
            if (homopolymer(own, k)) { o += homoSkip; pre = false; continue; }      // :260 / :370 (rc of a homopolymer is one too)
            uint32_t b = preB, e = preE;
            bool hit = pre ? true : probe(ix, own, b, e);                           // :261 / :372
            pre = false;
            if (hit) {
                int fpos = s == 0 ? o : L - (o + k);     // forward offset of this k-mer (:366)
                if (strict) {
                    bool other = present(ix, revcomp(own, k));                      // :268 / :379
                    if (s == 0) { ++fwdHit; if (other) ++rcHit; CG_PUSH_KS(fpos, 1, other ? 1 : 0); }
                    else { ++rcHit; if (other) ++fwdHit; CG_PUSH_KS(fpos, other ? 1 : 0, 1); }
                }
                uint32_t lb = b > 0 ? b - 1 : 0, ub = e;                            // :279 / :392
                extend(ix, rv, s, o, lb, ub, k, lb, ub, matched);                   // :280 / :393
                if (ub > lb && ub - lb < maxInterval) {                             // :285 / :398
                    if (s == 0) CG_PUSH_INT(st.fwd, st.nF, lb, ub, matched, o);
                    else CG_PUSH_INT(st.rc, st.nR, lb, ub, matched, o);
                    if (strict && o + matched < L) {                                // :291 / :404
                        int kp = s == 0 ? o + matched - skipOverlap : L - (o + matched);   // :292 / :405
                        CG_PUSH_KS(kp, 0, 0);
                    }
                }
                if (lastSearch) break;                                              // :300 / :412
                int mismatch = o + matched;
                if (mismatch < L) {
                    int l = lce(ix, lb, ub - 1, matched, L - mismatch, lceMode);    // :304 / :416
                    o = cmax(o + l - skipOverlap, mismatch - skipOverlap);
                    if (o > L - k) { o = L - k; lastSearch = true; }
                } else {
                    lastSearch = true; o = L - k;
                }
            } else {
                // sampFactor = 1 (:324 / :436): the loop would now try o+1, o+2, ... one dependent table
                // look-up at a time.  The positions it visits do not depend on the look-ups' results, so the
                // next four are looked up together and the first one present is where the loop resumes.
                o += 1;
                for (;;) {
                    int q[4]; uint64_t key[4], hh[4]; u32x4 sl[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        q[c] = -1; key[c] = 0; hh[c] = 0; sl[c].x = sl[c].y = sl[c].z = sl[c].w = 0;
                        while (o + k <= L) {
                            int iv = rv.first_n(s, o);
                            if (iv < o + k) { o = iv + 1; continue; }
                            uint64_t kw = rv.chunk(s, o) & KM;
                            if (homopolymer(kw, k)) { o += homoSkip; continue; }
                            q[c] = o; key[c] = kw; hh[c] = home_slot(ix, kw);
                            sl[c] = CG_LDG(&ix.hslots[hh[c]]);
                            ++o;
                            break;
                        }
                    }
                    int got = -1;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (got < 0 && q[c] >= 0) {
                            uint32_t bb, ee;
                            if (resolve(ix, key[c], hh[c], sl[c], bb, ee)) { got = q[c]; preB = bb; preE = ee; }
                        }
                    }
                    if (got >= 0) { o = got; pre = true; break; }
                    if (q[3] < 0) break;            // ran off the read: o + k > L, the outer loop ends
                }
            }
        }
    }

    // ---- strand vote (:442-490)
    if (strict) {
        if (fwdHit > 0 && rcHit == 0) st.nR = 0;
        else if (rcHit > 0 && fwdHit == 0) st.nF = 0;
        else {
            StdSort ss; ss.a = st.ks; ss.template sort<(MAXKS > 16)>(st.nKS);                              // :450
            int fs = 0, rs = 0;
            uint32_t prevKey = 0xffffffffu;
            for (int i = 0; i < st.nKS; ++i) {                                      // std::unique keeps the first of each run
                uint32_t en = st.ks[i];
                if (ks_key(en) == prevKey) continue;
                prevKey = ks_key(en);
                int f = ks_f(en), r = ks_r(en);
                if (f == 0 || r == 0) {
                    uint64_t w = kmer_trunc(rv, (int)ks_key(en), k);
                    if (f == 0) f = present(ix, w) ? 1 : -1;                        // :461-464
                    if (r == 0) r = present(ix, revcomp(w, k)) ? 1 : -1;            // :469-473
                }
                fs += f; rs += r;
            }
            if (fs > rs) st.nR = 0; else if (rs > fs) st.nF = 0;                    // :482-488
        }
    }
#undef CG_PUSH_KS
#undef CG_PUSH_INT
    return true;
}

// ---------------------------------------------------------------------------------------------
// intervals -> transcript hits (:492-579).  For one strand's interval list this produces the
// tid-sorted hit list that intersectSAHits + collectHitsSimpleSA (>1 interval) or the single-
// interval enumeration (:497-526) yields: tid[i], and when WITH_POS also pos[i] = minPos - its
// queryPos.  cand/act/(cpos,cq) are caller-provided scratch of capacity MAXC.
// ---------------------------------------------------------------------------------------------
template <bool WITH_POS>
struct HitScratch {
    uint32_t* tid;     // MAXC
    uint8_t* act;      // MAXC
    uint32_t* pos;     // MAXC (WITH_POS)
    uint16_t* qp;      // MAXC (WITH_POS)
};

template <bool WITH_POS>
CG_HD void hs_swap(HitScratch<WITH_POS>& h, int i, int j) {
    uint32_t t = h.tid[i]; h.tid[i] = h.tid[j]; h.tid[j] = t;
    if (WITH_POS) {
        uint32_t p = h.pos[i]; h.pos[i] = h.pos[j]; h.pos[j] = p;
        uint16_t q = h.qp[i]; h.qp[i] = h.qp[j]; h.qp[j] = q;
    }
}
template <bool WITH_POS>
CG_HD bool hs_less(const HitScratch<WITH_POS>& h, int i, int j) {
    if (h.tid[i] != h.tid[j]) return h.tid[i] < h.tid[j];
    if (WITH_POS) return h.pos[i] < h.pos[j];
    return false;
}
// in-place heap sort by (tid, pos); equal (tid,pos) pairs cannot occur inside one interval
template <bool WITH_POS>
CG_HD void hs_sort(HitScratch<WITH_POS>& h, int n) {
    if (n <= 24) {
        for (int i = 1; i < n; ++i)
            for (int j = i; j > 0 && hs_less(h, j, j - 1); --j) hs_swap(h, j, j - 1);
        return;
    }
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && hs_less(h, child, child + 1)) ++child;
            if (!hs_less(h, root, child)) break;
            hs_swap(h, root, child); root = child;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        hs_swap(h, 0, end);
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && hs_less(h, child, child + 1)) ++child;
            if (!hs_less(h, root, child)) break;
            hs_swap(h, root, child); root = child;
        }
    }
}

// returns the number of hits (tids ascending in h.tid[0..ret)); for WITH_POS h.pos holds the
// QuasiAlignment::pos of each (as int32 bit pattern)
template <bool WITH_POS>
CG_HD int hit_tids(const IndexView& ix, const Interval* ints, int n, HitScratch<WITH_POS>& h, int maxc, int& err) {
    if (n == 0) return 0;
    // smallest span, first on ties (intersectSAHits); with one interval it is that interval
    int mi = 0;
    for (int i = 1; i < n; ++i)
        if (ints[i].end - ints[i].begin < ints[mi].end - ints[mi].begin) mi = i;
    const Interval mh = ints[mi];
    int nc = 0;
    for (uint32_t i0 = mh.begin; i0 < mh.end; i0 += 4) {
        int nn = (int)cmin<uint32_t>(4u, mh.end - i0);
        uint32_t tt[4], ll[4];
        sa_tids4(ix, i0, nn, tt, ll);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            if (x < nn) {
                if (nc >= maxc) { err |= CG_ERR_HITS; }
                else {
                    h.tid[nc] = tt[x]; h.act[nc] = 1;
                    if (WITH_POS) { h.pos[nc] = ll[x]; h.qp[nc] = mh.qpos; }
                    ++nc;
                }
            }
        }
        if (err & CG_ERR_HITS) break;
    }
    hs_sort(h, nc);
    // unique by tid keeping the smallest pos (first after the (tid,pos) sort)
    int m = 0;
    for (int i = 0; i < nc; ++i) {
        if (m > 0 && h.tid[m - 1] == h.tid[i]) continue;
        if (m != i) { h.tid[m] = h.tid[i]; if (WITH_POS) { h.pos[m] = h.pos[i]; h.qp[m] = h.qp[i]; } }
        h.act[m] = 1;
        ++m;
    }
    nc = m;
    int counter = 2;
    for (int x = 0; x < n; ++x) {
        if (x == mi) continue;
        const Interval it = ints[x];
        // an interval equal to minHit or to one already processed cannot drop a live tid and
        // cannot lower a min position: pass every live candidate through without touching memory
        bool dup = (it.begin == mh.begin && it.end == mh.end);
        for (int y = 0; y < x && !dup; ++y)
            if (y != mi && ints[y].begin == it.begin && ints[y].end == it.end) dup = true;
        if (dup) {
            for (int c = 0; c < nc; ++c) if (h.act[c] == counter - 1) h.act[c] = (uint8_t)counter;
        } else {
            for (uint32_t i0 = it.begin; i0 < it.end; i0 += 4) {
                int nn = (int)cmin<uint32_t>(4u, it.end - i0);
                uint32_t tt[4], ll[4];
                sa_tids4(ix, i0, nn, tt, ll);
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    if (z < nn) {
                        uint32_t tid = tt[z], local = ll[z];
                        int lo = 0, hi = nc;
                        while (lo < hi) { int md = (lo + hi) >> 1; if (h.tid[md] < tid) lo = md + 1; else hi = md; }
                        if (lo < nc && h.tid[lo] == tid) {
                            if (h.act[lo] == counter - 1) h.act[lo] = (uint8_t)counter;
                            if (WITH_POS && h.act[lo] == counter && local < h.pos[lo]) { h.pos[lo] = local; h.qp[lo] = it.qpos; }
                        }
                    }
                }
            }
        }
        ++counter;
    }
    // active = numActive >= #intervals; compact
    m = 0;
    for (int c = 0; c < nc; ++c) {
        if (h.act[c] >= n) {
            h.tid[m] = h.tid[c];
            if (WITH_POS) h.pos[m] = (uint32_t)((int32_t)h.pos[c] - (int32_t)h.qp[c]);
            ++m;
        }
    }
    return m;
}

} // namespace cg
