// circall_b200 — warp-per-read collector (device only): the fast path of k_wt / k_bsj.
//
// One warp maps one read.  Control flow is warp-uniform and mirrors cg::collect() (and through it
// include/SACollectorCircall.hpp:24-582) statement by statement; the inner loops that dominate the
// reference's run time are spread over the 32 lanes:
//   * runs of absent k-mers (the ~31 windows that cover a mismatch; the hunt for a first seed): the
//     positions the reference would visit one dependent table look-up at a time do not depend on the
//     look-ups' results, so up to 32 of them are probed in one round and the first hit is taken (w_look)
//   * extendSearchNaive over narrow intervals: one candidate suffix per lane, max/first/last by ballot
//   * SA interval -> transcript ids, intersection over intervals, junction-span filter: one SA
//     position per lane
// Scalar steps (LCE, wide-interval binary searches) run redundantly on all lanes; same-address loads
// are broadcast, so they cost no extra memory traffic.
//
// The kernel is bound by instruction fetch as much as by memory when its body outgrows the
// instruction cache, so this file is written for code size: one out-of-line copy of every helper,
// arguments by value, the index view in constant memory.
// Two instantiations: the FAST one (state in shared memory, FAST_FI intervals per strand, FAST_FK k-mer
// scores, hit intervals of at most 32 positions) handles almost every read; a read that does not fit
// "overflows" — nothing is emitted — and is redone by the GENERAL one (worst-case capacities in per-warp
// global scratch, intervals of any width), which runs over the list of such reads.
#pragma once
#include "cg_core.cuh"

#ifndef CG_OPT_TWOBLOCK
#define CG_OPT_TWOBLOCK 0     // load the next text block together with the current one in w_match: measured slower (registers)
#endif
#ifndef CG_OPT_PF
#define CG_OPT_PF 1           // bit0: transcript-id half of a candidate's block (kept); bit1: last k-mer's slots, bit2: next read (measured slower)
#endif

namespace cgw {
using namespace cg;

static const unsigned FULL = 0xffffffffu;
// capacities: the fast instantiation keeps its state in shared memory (FI intervals per strand, FK kmerScores
// entries; <= 16 entries: libstdc++'s sort is a plain insertion sort); the general one, run on the reads the
// fast one hands over, keeps worst-case state (a 256-base read cannot produce more) in per-warp global scratch
static const int FAST_FI = 8, FAST_FK = 16, GEN_FI = 232, GEN_FK = 928;
#ifndef CG_W_32ARY
#define CG_W_32ARY 1          // wide open intervals: the three searches of extendSearchNaive 32 ways at a time instead of as binary searches
#endif
#ifndef CG_W_LINEAR
#define CG_W_LINEAR (CG_W_32ARY ? 64 : 256)
#endif
static const int W_LINEAR = CG_W_LINEAR;     // widest open interval the extension resolves candidate by candidate (32 per round)

__constant__ IndexView c_ix[2];          // [0] txIdx, [1] bsjIdx — written by the session before each stage

// -DCG_WARP_PROF: cycles the warp-per-read kernels spend per phase (lane 0 of every warp adds its clock64() deltas), and a log2
// histogram of the cycles per read; printed by cg_session_destroy.  Measurement build only (profiles/r2_experiments.md).
#ifdef CG_WARP_PROF
__device__ unsigned long long g_wprof[64];       // [0..7] phase cycles, [8] reads, [16 + b] reads whose cycles have bit length b
enum { WP_LOOK = 0, WP_EXTEND, WP_LCE, WP_VOTE, WP_HITS, WP_REST };
#define WPROF_T0() const long long _wp0 = clock64()
#define WPROF_ADD(ph, lane) do { if ((lane) == 0) atomicAdd(&cgw::g_wprof[ph], (unsigned long long)(clock64() - _wp0)); } while (0)
#else
#define WPROF_T0() do {} while (0)
#define WPROF_ADD(ph, lane) do {} while (0)
#endif

template <int FI, int FK>
struct WStateT {
    Interval fwd[FI];
    Interval rc[FI];
    uint32_t ks[FK];
};

struct RV { const uint64_t* w; int L; int W2; int anyN; };     // read view, passed by value

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src), hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// Read view in shared memory: W2 forward code words (32 bases each, the last a zero pad), W2 words of the
// reverse-complement strand in the same layout (built once per read by warp_view), then the N-mask words of the
// forward strand.
// 32-base window of strand s at oriented position i (i < L): three 32-bit words and two funnel shifts
__device__ __forceinline__ uint64_t w_chunk(RV r, int s, int i) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(r.w + (s ? r.W2 : 0)) + (i >> 4);
    const int sh = (i & 15) * 2;
    const uint32_t a = p[0], b = p[1], c = p[2];
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}
// 32 N-mask bits of strand s at oriented position i (only reads that hold an N get here)
__device__ __noinline__ uint32_t w_mask(RV r, int s, int i) {
    ReadView v; v.w = r.w + r.W2; v.L = r.L; v.W2 = r.W2; v.WM = 0; v.anyN = true;     // msk(j) = w[W2 + j]
    return v.mask32(s, i);
}

struct Hit { uint32_t b, e; bool found; };
// table look-up with a different key per lane; every lane of the warp must call
template <int X>
__device__ __noinline__ Hit w_probe(uint64_t key, bool active) {
    const IndexView& ix = c_ix[X];
    uint64_t h = home_slot(ix, key);
    Hit r; r.b = r.e = 0; r.found = false;
    bool pend = active;
    while (__any_sync(FULL, pend)) {
        if (pend) {
            u32x4 s = CG_LDG(&ix.hslots[h]);
            uint64_t kk = (uint64_t)s.x | ((uint64_t)s.y << 32);
            if (kk == key) { r.b = s.z; r.e = s.w; r.found = true; pend = false; }
            else if (kk == ~0ULL) pend = false;
            else h = (h + 1) & ix.hmask;
        }
    }
    return r;
}

// query-vs-suffix comparison (cg::match_run): returns 4 * matched length + (ord + 1)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

template <int X>
__device__ __noinline__ int w_match(RV r, int s, int qstart, uint32_t p, int i, int lim) {
    const IndexView& ix = c_ix[X];
    uint32_t curb = 0xffffffffu, nxb = 0xffffffffu;
    uint64_t c0 = 0, c1 = 0, tmk = 0, n0 = 0, n1 = 0, nmk = 0;
    while (i < lim) {
        uint32_t tp = p + (uint32_t)i;
        uint32_t b = tp >> 6;
        if (b != curb) {
            if (b == nxb) { c0 = n0; c1 = n1; tmk = nmk; }
            else {
                u32x4 a = CG_LDG(&ix.text[2 * (uint64_t)b]);
                u32x4 c = CG_LDG(&ix.text[2 * (uint64_t)b + 1]);
                c0 = (uint64_t)a.x | ((uint64_t)a.y << 32); c1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
                tmk = (uint64_t)c.x | ((uint64_t)c.y << 32);
            }
            curb = b;
            // the block after this one goes out now when the compared range reaches into it: its latency
            // overlaps this block's load / comparison instead of following it
            nxb = 0xffffffffu;
            if (CG_OPT_TWOBLOCK && (int)(tp & 63) + (lim - i) > 64) {
                u32x4 a = CG_LDG(&ix.text[2 * (uint64_t)b + 2]);
                u32x4 c = CG_LDG(&ix.text[2 * (uint64_t)b + 3]);
                n0 = (uint64_t)a.x | ((uint64_t)a.y << 32); n1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
                nmk = (uint64_t)c.x | ((uint64_t)c.y << 32);
                nxb = b + 1;
            }
        }
        int half = (tp >> 5) & 1, o5 = tp & 31;
        uint64_t tw = (half ? c1 : c0) >> (2 * o5);
        uint32_t tm = (uint32_t)(tmk >> (32 * half + o5));
        int n = cmin(32 - o5, lim - i);
        uint64_t q = w_chunk(r, s, qstart + i);
        uint32_t qm = r.anyN ? w_mask(r, s, qstart + i) : 0u;
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
            int ord;
            if (tdollar) ord = 1;
            else if (qn) ord = (tc == 3) ? -1 : 1;
            else ord = qc < tc ? -1 : 1;
            return 4 * i + ord + 1;
        }
        i += n;
    }
    return 4 * i + 1;
}

struct Ext { uint32_t lb, ub; int len; };
// extendSearchNaive over the open interval (lbIn, ubIn): one candidate per lane when narrow, the
// reference's three binary searches (the same on every lane) when wide
template <int X>
__device__ __noinline__ Ext w_extend(RV r, int lane, int s, int qstart, uint32_t lbIn, uint32_t ubIn, int startAt) {
    const IndexView& ix = c_ix[X];
    Ext o;
    const int m = r.L - qstart;
    if (m <= startAt) { o.lb = lbIn + 1; o.ub = ubIn; o.len = startAt; return o; }
    const uint32_t w = ubIn - lbIn - 1;
    if (w == 0) { o.lb = ubIn; o.ub = ubIn; o.len = startAt; return o; }        // empty open interval
    if (w <= (uint32_t)W_LINEAR) {
        // every candidate suffix by its own lane, 32 per round: the suffixes that share the longest prefix are contiguous in
        // SA order, so the answer is the first and the last index that reach the maximum.  Up to W_LINEAR candidates this
        // beats the three binary searches below (2 x 3 x log2(w) dependent loads) on latency.
        int best = -1; uint32_t first = 0, last = 0;
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < w; c0 += 32) {
            int l = -1;
            if (c0 + (uint32_t)lane < w) {
                const uint32_t p = CG_LDG(&ix.sa[lbIn + 1 + c0 + lane]);
                if ((CG_OPT_PF & 1) && w <= 32 && !ix.satid) prefetch_l2(&ix.text[2 * (uint64_t)(p >> 6) + 1]);   // transcript-id half of the block of p, for the hit enumeration
                l = w_match<X>(r, s, qstart, p, startAt, m) >> 2;
            }
            const int cb = __reduce_max_sync(FULL, l);
            const unsigned mk = __ballot_sync(FULL, l == cb);
            const uint32_t f = c0 + (uint32_t)(__ffs(mk) - 1), e = c0 + (uint32_t)(31 - __clz(mk));
            if (cb > best) { best = cb; first = f; last = e; }
            else if (cb == best) last = e;
        }
        o.lb = lbIn + 1 + first;
        o.ub = lbIn + 1 + last + 1;
        o.len = best;
        return o;
    }
#if CG_W_32ARY
    // Wide open interval (a repeat family's k-mer: thousands of suffixes).  The reference runs three binary searches here — the
    // query's insertion point (its neighbours hold the longest match), then the first suffix that shares query[0:maxI] and the first
    // one past them — 3 x log2(w) dependent SA + text round trips.  The same three searches 32 ways at a time: 32 evenly spaced
    // suffixes of the current range are compared by the 32 lanes at once, the predicate (monotone along the suffix array) is
    // balloted and the range shrinks to the gap between two samples, until at most 32 candidates are left for one last round.
    // Same results (the brute-force definition of SURVEY A6 / B2), a sixth of the dependent rounds.
    // sample j (0..31) of the open interval (lo, hi) with n = hi - lo - 1 > 32 candidates: distinct, increasing
#define W_SAMPLE(lo, n, j) ((lo) + 1u + (uint32_t)(((uint64_t)(j) * (uint64_t)((n) - 1u)) / 31u))
    int maxI = startAt;
    uint32_t wit = lbIn + 1;                              // a candidate that reaches maxI (every candidate shares the first startAt bases)
    {   // ---- 1: insertion point (pred: query <= suffix) and the longest match seen on the way
        uint32_t lo = lbIn, hi = ubIn;
        for (;;) {
            const uint32_t n = hi - lo - 1;
            const bool last = n <= 32;
            const bool have = last ? (uint32_t)lane < n : true;
            const uint32_t c = last ? lo + 1 + (uint32_t)lane : W_SAMPLE(lo, n, lane);
            int i = -1; bool le = false;
            if (have) {
                const int v = w_match<X>(r, s, qstart, CG_LDG(&ix.sa[c]), startAt, m);
                i = v >> 2; le = i >= m || ((v & 3) - 1) < 0;
            }
            const int cb = __reduce_max_sync(FULL, i);
            if (cb > maxI) {
                const unsigned mk = __ballot_sync(FULL, i == cb);
                maxI = cb; wit = __shfl_sync(FULL, c, __ffs(mk) - 1);
            }
            if (last) break;
            const unsigned leB = __ballot_sync(FULL, le);
            const int t = leB ? __ffs(leB) - 1 : 32;      // first sample the query does not sort after
            const uint32_t nlo = t == 0 ? lo : __shfl_sync(FULL, c, t - 1), nhi = t == 32 ? hi : __shfl_sync(FULL, c, t & 31);
            lo = nlo; hi = nhi;
            if (hi - lo <= 1) break;                      // the insertion point lies between two adjacent samples: both were compared
        }
    }
    // ---- 2: first suffix of (lbIn, wit] that is >= query[0:maxI]  (pass 0), first suffix of (wit, ubIn) that is > it (pass 1)
    uint32_t bound[2];
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        // invariant: pred(lo) false (or lo is the excluded end), pred(hi) true (or hi is the excluded end)
        uint32_t lo = pass == 0 ? lbIn : wit, hi = pass == 0 ? wit : ubIn;
        while (hi - lo > 1) {
            const uint32_t n = hi - lo - 1;
            const bool last = n <= 32;
            const bool have = last ? (uint32_t)lane < n : true;
            const uint32_t c = last ? lo + 1 + (uint32_t)lane : W_SAMPLE(lo, n, lane);
            bool pr = false;
            if (have) {
                const int v = w_match<X>(r, s, qstart, CG_LDG(&ix.sa[c]), startAt, maxI);
                const int i = v >> 2, ord = (v & 3) - 1;
                pr = pass == 0 ? (i >= maxI || ord < 0) : (i < maxI && ord < 0);
            }
            const unsigned pB = __ballot_sync(FULL, pr);
            const int t = pB ? __ffs(pB) - 1 : 32;
            if (last) { if (t < 32) hi = lo + 1 + (uint32_t)t; lo = hi - 1; break; }
            const uint32_t nlo = t == 0 ? lo : __shfl_sync(FULL, c, t - 1), nhi = t == 32 ? hi : __shfl_sync(FULL, c, t & 31);
            lo = nlo; hi = nhi;
        }
        bound[pass] = hi;
    }
#undef W_SAMPLE
    o.lb = bound[0]; o.ub = bound[1]; o.len = maxI;
    return o;
}
#else
    uint32_t lb = lbIn, ub = ubIn;
    int prevLow = startAt, prevHigh = startAt, maxI = startAt;
    while (ub - lb > 1) {
        uint32_t c = lb + (ub - lb) / 2;
        int v = w_match<X>(r, s, qstart, CG_LDG(&ix.sa[c]), cmin(prevLow, prevHigh), m);
        int i = v >> 2, ord = (v & 3) - 1;
        if (i > maxI) maxI = i;
        if (i >= m || ord < 0) { ub = c; prevHigh = i; } else { lb = c; prevLow = i; }
    }
    // pass 0: first c whose suffix is >= query[0:maxI]; pass 1: first c whose suffix is > it
    uint32_t lower = 0, lo = lbIn, hi = ubIn;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        while (hi - lo > 1) {
            uint32_t c = lo + (hi - lo) / 2;
            int v = w_match<X>(r, s, qstart, CG_LDG(&ix.sa[c]), startAt, maxI);
            int i = v >> 2, ord = (v & 3) - 1;
            bool goLeft = pass == 0 ? (i >= maxI || ord < 0) : (i < maxI && ord < 0);
            if (goLeft) hi = c; else lo = c;
        }
        if (pass == 0) { lower = hi; lo = lower - 1; hi = ubIn; }
    }
    o.lb = lower; o.ub = hi; o.len = maxI;
    return o;
}
#endif

template <int X>
__device__ __noinline__ int w_lce(uint32_t p1, uint32_t p2, int startAt, int stopAt, int mode) {
    return lce(c_ix[X], p1, p2, startAt, stopAt, mode);
}

// The reference's position walk without the look-up results (:112-128, :234-260, :341-372): starting at
// oriented position o on strand s visit positions <= lim — an N inside the window (nbits window bits)
// moves past it, a homopolymer window skips k/2, anything else is looked up — until a look-up hits.
//   both  : two keys per position (the k-mer and its reverse complement, lanes 2j / 2j+1), else one
//   once  : stop after the first position that was looked up, hit or not (single expected-present probe)
//   either: a position counts as a hit when either key is present (phase A), else only the first key
// Returns pos >= 0 (lane jl holds the first key's interval in b/e, fb = found-ballot) or pos < 0 with
// o_end = the position the walk reached.
struct Look { int pos, o_end, jl; unsigned fb; uint32_t b, e; };
template <int X>
__device__ __noinline__ Look w_look(RV r, int lane, int s, int o, int lim, uint32_t nbits, bool both, bool once, bool either) {
    const int k = c_ix[X].k, homoSkip = k / 2;
    const uint64_t KM = c_ix[X].km;
    const int sh = both ? 1 : 0;
    const int np = once ? 1 : (32 >> sh);
    Look R; R.pos = -1; R.jl = 0; R.fb = 0; R.b = R.e = 0; R.o_end = o;
    if (once && !r.anyN && o <= lim) {
        // the common single probe: no N anywhere in the read, so only the homopolymer rule can intervene
        uint64_t mer = w_chunk(r, s, o) & KM;
        if (!homopolymer_ix(mer, c_ix[X])) {
            Hit h = w_probe<X>((both && (lane & 1)) ? revcomp(mer, k) : mer, lane < (both ? 2 : 1));
            unsigned fb = __ballot_sync(FULL, h.found);
            if (both && either ? (fb & 3u) : (fb & 1u)) {
                R.jl = 0; R.pos = o; R.fb = fb; R.b = h.b; R.e = h.e;
                return R;
            }
            R.o_end = o + 1;
            return R;
        }
    }
    for (;;) {
        if (o > lim) { R.o_end = o; return R; }
        int j = lane >> sh;
        int p = o + j;
        bool inr = j < np && p <= lim;
        uint32_t m32 = inr ? (r.anyN ? (w_mask(r, s, p) & nbits) : 0u) : 1u;
        uint64_t mer = inr ? (w_chunk(r, s, p) & KM) : 0;
        bool nbad = m32 != 0;
        bool homo = !nbad && homopolymer_ix(mer, c_ix[X]);
        unsigned inrB = __ballot_sync(FULL, inr), nB = __ballot_sync(FULL, nbad) & inrB, hB = __ballot_sync(FULL, homo) & inrB;
        unsigned visit = 0;
        int q = 0;
        if (hB == 0 && !both) { visit = inrB & ~nB; q = __popc(inrB); }            // plain walk: every N-free window is looked up
        else {
            while (q < np && ((inrB >> (q << sh)) & 1)) {
                if ((nB >> (q << sh)) & 1) ++q;
                else if ((hB >> (q << sh)) & 1) q += homoSkip;
                else { visit |= (both ? 3u : 1u) << (q << sh); ++q; }
            }
        }
        uint64_t key = (both && (lane & 1)) ? revcomp(mer, k) : mer;
        Hit h = w_probe<X>(key, (visit >> lane) & 1);
        unsigned fb = __ballot_sync(FULL, h.found);
        unsigned pm = both ? ((either ? (fb | (fb >> 1)) : fb) & 0x55555555u) : fb;
        if (pm) {
            R.jl = __ffs(pm) - 1;
            R.pos = o + (R.jl >> sh);
            R.fb = fb; R.b = h.b; R.e = h.e;
            return R;
        }
        o += q;
        if (once && visit) { R.o_end = o; return R; }
    }
}

__device__ __noinline__ void w_put_int(Interval* arr, int idx, uint32_t b, uint32_t e, int len, int q) {
    arr[idx].begin = b; arr[idx].end = e; arr[idx].len = (uint16_t)len; arr[idx].qpos = (uint16_t)q;
}

// SACollectorCircall::operator() up to and including the strand vote, one read per warp.
// Returns false on overflow (the caller must hand the read to the general path).  found / nF / nR as in
// cg::collect(); the post-vote intervals are in st.fwd / st.rc.
template <int X, int WFI, int WFK>
__device__ __forceinline__ bool w_collect(RV rv, int lane, int lceMode, WStateT<WFI, WFK>& st, bool& found, int& nF, int& nR) {
    const int L = rv.L, k = c_ix[X].k;
    const int skipOverlap = k - 1;
    const uint32_t maxInterval = 1000;
    const uint32_t kbits = (k >= 32) ? 0xffffffffu : ((1u << k) - 1);
    const uint32_t k1bits = (k + 1 >= 32) ? 0xffffffffu : ((1u << (k + 1)) - 1);
    nF = nR = 0; found = false;
    int nKS = 0;
    bool ovf = false;
    // the forced last look-up (:300-321) probes the k-mer at L - k and its reverse complement: start those two
    // sectors on their way now
    if ((CG_OPT_PF & 2) && L > k && lane < 2 && !rv.anyN) {
        const uint64_t mer = w_chunk(rv, 0, L - k) & c_ix[X].km;
        prefetch_l2(&c_ix[X].hslots[home_slot(c_ix[X], lane ? revcomp(mer, k) : mer)]);
    }
    uint32_t fwdHit = 0, rcHit = 0;
    uint32_t lbF = 0, ubF = 0;
    int matched = 0;

#define W_PUSH_KS(kp, f, r) do { if (nKS < WFK) { if (lane == 0) st.ks[nKS] = ks_pack((kp), (f), (r)); ++nKS; } else ovf = true; } while (0)
#define W_PUSH_INT(arr, cnt, b_, e_, l_, q_) do { if (cnt < WFI) { if (lane == 0) w_put_int(arr, cnt, (b_), (e_), (l_), (q_)); ++cnt; } else ovf = true; } while (0)

    // ---- Phase A (:112-200): `re < read.end()`, N test inclusive of pos + k
    int rb = 0, q0 = 0;
    bool first = true;
    while (rb + k < L) {
        Look lk; { WPROF_T0(); lk = w_look<X>(rv, lane, 0, rb, L - k - 1, k1bits, true, first, true); WPROF_ADD(WP_LOOK, lane); }  // :117-132
        first = false;
        if (lk.pos < 0) { rb = lk.o_end; continue; }
        rb = lk.pos;
        const bool hasF = (lk.fb >> lk.jl) & 1, hasR = (lk.fb >> (lk.jl + 1)) & 1;
        if (hasF) {
            uint32_t b = __shfl_sync(FULL, lk.b, lk.jl), e = __shfl_sync(FULL, lk.e, lk.jl);
            Ext x; { WPROF_T0(); x = w_extend<X>(rv, lane, 0, rb, b > 0 ? b - 1 : 0, e, k); WPROF_ADD(WP_EXTEND, lane); }           // :140-143
            lbF = x.lb; ubF = x.ub; matched = x.len;
            if (ubF > lbF && ubF - lbF < maxInterval) {                              // :149
                W_PUSH_INT(st.fwd, nF, lbF, ubF, matched, rb);
                q0 = rb;
                ++fwdHit;
                if (hasR) { ++rcHit; W_PUSH_KS(rb, 1, 1); } else { W_PUSH_KS(rb, 1, -1); }
                if (rb + matched < L) W_PUSH_KS(rb + matched - skipOverlap, -1, 0);  // :165-168
            }
        }
        if (hasR && !fwdHit) {                                                       // :178-191
            ++rcHit;
            W_PUSH_KS(rb, -1, 1);
        }
        if (fwdHit + rcHit > 0) { found = true; break; }
        ++rb;
    }
    if (!found) return true;                                                         // :204

    // ---- Phases B (:208-329) and C (:332-440)
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        int o;
        bool lastSearch = false, expectAbsent = false;
        if (s == 0) {
            if (!fwdHit) continue;
            int remaining = L - (q0 + matched);                                      // :219
            int l; { WPROF_T0(); l = w_lce<X>(lbF, ubF - 1, matched, remaining, lceMode); WPROF_ADD(WP_LCE, lane); }             // :220
            int fwdSkip = cmax(matched - skipOverlap, l - skipOverlap);
            o = cmin(cmax(0, L - k), q0 + fwdSkip);                                  // :224-228
            expectAbsent = remaining > 0 && o < L - k;    // the match stopped on a mismatch: the next windows cover it
        } else {
            if (!(rcHit >= fwdHit)) continue;                                        // :332
            o = 0;
        }
        int prevO = -1; uint32_t pB = 0, pE = 0; bool pOther = false;
        while (o + k <= L) {                                                         // :234 / :341-344
            uint32_t b, e; bool other;
            if (o == prevO) { b = pB; e = pE; other = pOther; }      // the forced last position was just looked up
            else {
                // expected present: this k-mer and its reverse complement on lanes 0/1 (:261+:268 / :372+:379);
                // after a mismatch: up to 32 positions at once
                Look lk; { WPROF_T0(); lk = w_look<X>(rv, lane, s, o, L - k, kbits, !expectAbsent, !expectAbsent, false); WPROF_ADD(WP_LOOK, lane); }
                if (lk.pos < 0) { o = lk.o_end; expectAbsent = true; continue; }     // :324 / :436
                o = lk.pos;
                b = __shfl_sync(FULL, lk.b, lk.jl); e = __shfl_sync(FULL, lk.e, lk.jl);
                if (!expectAbsent) other = (lk.fb >> (lk.jl + 1)) & 1;
                else other = __shfl_sync(FULL, (int)w_probe<X>(revcomp(w_chunk(rv, s, o) & c_ix[X].km, k), lane == 0).found, 0);   // :268 / :379
                expectAbsent = false;
                prevO = o; pB = b; pE = e; pOther = other;
            }
            int fpos = s == 0 ? o : L - (o + k);                                     // :366
            if (s == 0) { ++fwdHit; if (other) ++rcHit; W_PUSH_KS(fpos, 1, other ? 1 : 0); }
            else { ++rcHit; if (other) ++fwdHit; W_PUSH_KS(fpos, other ? 1 : 0, 1); }
            Ext x; { WPROF_T0(); x = w_extend<X>(rv, lane, s, o, b > 0 ? b - 1 : 0, e, k); WPROF_ADD(WP_EXTEND, lane); }            // :279-280 / :392-393
            matched = x.len;
            if (x.ub > x.lb && x.ub - x.lb < maxInterval) {                          // :285 / :398
                if (s == 0) W_PUSH_INT(st.fwd, nF, x.lb, x.ub, matched, o);
                else W_PUSH_INT(st.rc, nR, x.lb, x.ub, matched, o);
                if (o + matched < L) {                                               // :291 / :404
                    int kp = s == 0 ? o + matched - skipOverlap : L - (o + matched); // :292 / :405
                    W_PUSH_KS(kp, 0, 0);
                }
            }
            if (lastSearch) break;                                                   // :300 / :412
            int mismatch = o + matched;
            if (mismatch < L) {
                int l; { WPROF_T0(); l = w_lce<X>(x.lb, x.ub - 1, matched, L - mismatch, lceMode); WPROF_ADD(WP_LCE, lane); }    // :304 / :416
                o = cmax(o + l - skipOverlap, mismatch - skipOverlap);
                if (o > L - k) { o = L - k; lastSearch = true; }
                else expectAbsent = true;
            } else {
                lastSearch = true; o = L - k;
            }
        }
    }
    if (ovf) return false;
    __syncwarp();

    // ---- strand vote (:442-490)
    WPROF_T0();
    if (fwdHit > 0 && rcHit == 0) nR = 0;
    else if (rcHit > 0 && fwdHit == 0) nF = 0;
    else {
        // std::sort: a stable insertion sort up to 16 entries, libstdc++'s introsort (replicated step by step by
        // cg::StdSort) above; std::unique keeps the first of each run of equal kpos
        if (lane == 0) {
            if (WFK <= 16 || nKS <= 16) {
#pragma unroll 1
                for (int i = 1; i < nKS; ++i) {
                    uint32_t v = st.ks[i]; int j = i - 1;
                    while (j >= 0 && ks_key(v) < ks_key(st.ks[j])) { st.ks[j + 1] = st.ks[j]; --j; }
                    st.ks[j + 1] = v;
                }
            } else {
                StdSort ss; ss.a = st.ks; ss.template sort<(WFK > 16)>(nKS);
            }
        }
        __syncwarp();
        int fs = 0, rs = 0;
#pragma unroll 1
        for (int base = 0; base < nKS; base += 32) {
            const int i = base + lane;
            uint32_t en = i < nKS ? st.ks[i] : 0;
            uint32_t prev = i > 0 && i < nKS ? st.ks[i - 1] : 0;
            bool keep = i < nKS && (i == 0 || ks_key(prev) != ks_key(en));
            int f = keep ? ks_f(en) : 0, r = keep ? ks_r(en) : 0;
            uint64_t w = 0;
            if (keep && (f == 0 || r == 0)) {         // k-mer word cut at the first non-base
                int pos = (int)ks_key(en);
                w = w_chunk(rv, 0, pos) & c_ix[X].km;
                uint32_t m = rv.anyN ? (w_mask(rv, 0, pos) & kbits) : 0u;
                if (m) w &= (1ULL << (2 * ctz32(m))) - 1;
            }
            bool pf = w_probe<X>(w, keep && f == 0).found;                           // :461-464
            bool pr = w_probe<X>(revcomp(w, k), keep && r == 0).found;               // :469-473
            if (keep && f == 0) f = pf ? 1 : -1;
            if (keep && r == 0) r = pr ? 1 : -1;
            fs += __reduce_add_sync(FULL, f); rs += __reduce_add_sync(FULL, r);
        }
        if (fs > rs) nR = 0; else if (rs > fs) nF = 0;                               // :482-488
    }
    WPROF_ADD(WP_VOTE, lane);
#undef W_PUSH_KS
#undef W_PUSH_INT
    return true;
}

// transcriptAtPosition(SA[i]): from the index's per-suffix transcript array when it has one, else through the text block
template <int X>
__device__ __forceinline__ uint32_t tid_at(uint32_t p);
template <int X>
__device__ __forceinline__ uint32_t sa_tid(uint32_t i) {
    if (c_ix[X].satid) return CG_LDG(&c_ix[X].satid[i]);
    return tid_at<X>(CG_LDG(&c_ix[X].sa[i]));
}
// transcriptAtPosition(p) alone (no offset look-up)
template <int X>
__device__ __forceinline__ uint32_t tid_at(uint32_t p) {
    u32x4 c = CG_LDG(&c_ix[X].text[2 * (uint64_t)(p >> 6) + 1]);
    uint64_t mask = (uint64_t)c.x | ((uint64_t)c.y << 32);
    uint32_t o = p & 63;
    return c.z + (uint32_t)__popcll(o ? (mask & ((1ULL << o) - 1)) : 0ULL);
}

// SA interval list of one strand -> its hits (intersectSAHits + collectHitsSimpleSA, or the single-interval
// enumeration, :492-563), narrow form: the narrowest maximal interval spans <= 32 positions, lane c holds transcript id `tid`
// and `alive` says whether it is a hit.  ok = false when the narrowest maximal interval is wider (the general kernel's job).
struct Hits { uint32_t tid; bool alive, ok; uint32_t count; };
template <int X>
__device__ __noinline__ Hits w_hits_lanes(int lane, const Interval* ints, int n) {
    const IndexView& ix = c_ix[X];
    Hits H; H.alive = false; H.tid = 0xffffffffu; H.ok = true; H.count = 0;
    if (n == 0) return H;
    // An interval whose query range lies inside that of a longer interval of the same strand cannot change the hit
    // set: every occurrence p of the longer match gives the occurrence p + (qpos' - qpos) of the shorter one in the
    // same transcript, so tids(longer) is a subset of tids(shorter) and span(longer) <= span(shorter); the
    // intersection over all intervals (intersectSAHits) equals the one over the maximal intervals.  Skip the rest.
    bool drop = false;
    if (lane < n) {
        const int ql = ints[lane].qpos, ll = ints[lane].len;
#pragma unroll 1
        for (int y = 0; y < n; ++y) {
            const int qy = ints[y].qpos, ly = ints[y].len;
            drop |= (ly > ll && qy <= ql && qy + ly >= ql + ll);
        }
    }
    const unsigned skipB = __ballot_sync(FULL, drop);
    int mi = -1;
    uint32_t msp = 0xffffffffu;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        if ((skipB >> i) & 1) continue;
        uint32_t sp = ints[i].end - ints[i].begin;
        if (sp < msp) { msp = sp; mi = i; }
    }
    if (msp > 32) { H.ok = false; return H; }       // the candidates (narrowest interval) must fit the lanes; the others may be wider
    const uint32_t mb = ints[mi].begin, me = ints[mi].end;
    bool have = (uint32_t)lane < msp;
    if (have) H.tid = sa_tid<X>(mb + lane);
    // one lane per distinct transcript
    unsigned same = __match_any_sync(FULL, have ? H.tid : (0x80000000u | (uint32_t)lane));
    H.alive = have && (__ffs(same) - 1 == lane);
#pragma unroll 1
    for (int x = 0; x < n; ++x) {
        if (x == mi || ((skipB >> x) & 1)) continue;
        const uint32_t ib = ints[x].begin, ie = ints[x].end;
        bool dup = (ib == mb && ie == me);                           // same positions: cannot drop a transcript
#pragma unroll 1
        for (int y = 0; y < x && !dup; ++y) dup = (y != mi && ints[y].begin == ib && ints[y].end == ie);
        if (dup) continue;
        // a candidate stays alive iff its transcript occurs among this interval's positions (32 of them per round)
        const unsigned aB = __ballot_sync(FULL, H.alive);
        unsigned seenB = 0;
#pragma unroll 1
        for (uint32_t c0 = ib; c0 < ie; c0 += 32) {
            uint32_t t = 0xffffffffu;
            if (c0 + (uint32_t)lane < ie) t = sa_tid<X>(c0 + lane);
#pragma unroll 1
            for (unsigned m = aB & ~seenB; m; m &= m - 1) {
                int a = __ffs(m) - 1;
                uint32_t ta = __shfl_sync(FULL, H.tid, a);
                if (__any_sync(FULL, t == ta)) seenB |= 1u << a;
            }
        }
        if (H.alive && !((seenB >> lane) & 1u)) H.alive = false;
    }
    H.count = __popc(__ballot_sync(FULL, H.alive));
    return H;
}

// The same for intervals of any width (general kernel): the candidates go through scratch (scr: WSCR transcript
// ids, act: WSCR bytes) — sorted with a bitonic network, de-duplicated, intersected interval by interval.
// Returns the number of hits; scr[0..count) holds them in ascending transcript order.
static const int WSCR = 1024;       // an interval kept by the collector spans < maxInterval = 1000 positions
template <int X>
__device__ __noinline__ uint32_t w_hits_list(int lane, const Interval* ints, int n, uint32_t* scr, uint8_t* act) {
    const IndexView& ix = c_ix[X];
    if (n == 0) return 0;
    int mi = 0;
    uint32_t msp = ints[0].end - ints[0].begin;
#pragma unroll 1
    for (int i = 1; i < n; ++i) {
        uint32_t sp = ints[i].end - ints[i].begin;
        if (sp < msp) { msp = sp; mi = i; }
    }
    const uint32_t mb = ints[mi].begin, me = ints[mi].end;
    __syncwarp();
    uint32_t N = 32;
    while (N < msp) N <<= 1;
    for (uint32_t c = lane; c < N; c += 32) scr[c] = c < msp ? sa_tid<X>(mb + c) : 0xffffffffu;
    __syncwarp();
    for (uint32_t k2 = 2; k2 <= N; k2 <<= 1)
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < N; i += 32) {
                uint32_t p = i ^ j;
                if (p > i) {
                    uint32_t a = scr[i], b = scr[p];
                    if ((a > b) == ((i & k2) == 0)) { scr[i] = b; scr[p] = a; }
                }
            }
            __syncwarp();
        }
    uint32_t cnt = 0;                                 // distinct transcripts, in place
    for (uint32_t base = 0; base < N; base += 32) {
        uint32_t v = scr[base + lane];
        uint32_t pv = (base + lane) ? scr[base + lane - 1] : ~v;
        bool keep = v != 0xffffffffu && v != pv;
        unsigned kb = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) { uint32_t d = cnt + __popc(kb & ((1u << lane) - 1)); scr[d] = v; act[d] = 1; }
        cnt += __popc(kb);
        __syncwarp();
    }
    uint32_t counter = 1;
#pragma unroll 1
    for (int x = 0; x < n; ++x) {
        if (x == mi) continue;
        const uint32_t ib = ints[x].begin, ie = ints[x].end;
        bool dup = (ib == mb && ie == me);                           // same positions: cannot drop a transcript
#pragma unroll 1
        for (int y = 0; y < x && !dup; ++y) dup = (y != mi && ints[y].begin == ib && ints[y].end == ie);
        if (dup) continue;
        // mark the candidates that occur among this interval's positions (binary search in the sorted list)
        for (uint32_t c = ib + lane; c < ie; c += 32) {
            uint32_t t = sa_tid<X>(c);
            uint32_t lo = 0, hi = cnt;
            while (lo < hi) { uint32_t md = (lo + hi) >> 1; if (scr[md] < t) lo = md + 1; else hi = md; }
            if (lo < cnt && scr[lo] == t && act[lo] == counter) act[lo] = (uint8_t)(counter + 1);
        }
        ++counter;
        __syncwarp();
    }
    uint32_t m = 0;                                   // keep the transcripts every interval confirmed
    for (uint32_t base = 0; base < cnt; base += 32) {
        uint32_t i = base + lane;
        uint32_t v = i < cnt ? scr[i] : 0;
        bool keep = i < cnt && act[i] == counter;
        unsigned kb = __ballot_sync(FULL, keep);
        __syncwarp();
        if (keep) scr[m + __popc(kb & ((1u << lane) - 1))] = v;
        m += __popc(kb);
        __syncwarp();
    }
    return m;
}

// two ascending transcript lists merged, duplicates dropped (:567-579) -> out; returns the merged size
__device__ __noinline__ uint32_t w_merge(int lane, const uint32_t* a, uint32_t na, const uint32_t* b, uint32_t nb, uint32_t* out) {
    uint32_t n = 0;
    if (lane == 0) {
        uint32_t i = 0, j = 0;
        while (i < na || j < nb) {
            if (j >= nb || (i < na && a[i] <= b[j])) { uint32_t t = a[i++]; if (j < nb && b[j] == t) ++j; out[n++] = t; }
            else out[n++] = b[j++];
        }
    }
    n = __shfl_sync(FULL, n, 0);
    __syncwarp();
    return n;
}

} // namespace cgw
