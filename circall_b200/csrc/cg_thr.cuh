// circall_b200 — thread-per-read collector ("lockstep tier"): second tier of Circall_wt (after cg_pre.cuh) and first tier of Circall_bsj.
// Kernels (cg_session.cu): k_thr_col (q_begin / q_step loop / q_vote) and k_wt_fin / k_bsj_fin (q_wt_finish / q_bsj_finish).
//
// The clean-match front kernel showed what this hardware wants: one THREAD per read with several independent
// random sectors in flight each, as long as the 32 reads of a warp run the same instructions.  This file gives
// the same treatment to the reads that carry mismatches (or, in Circall_bsj, lie near a junction): the whole of
// SACollectorCircall::operator() (include/SACollectorCircall.hpp:24-582) for a read whose hits sit on ONE strand,
// written so that every heavy operation occurs exactly once in the instruction stream —
//
//     for each segment of the read:   LOOK (seed-filter stage + table look-ups, QW per round)   ->  EXTEND (extendSearchNaive: one
//                                     text comparison + the lcp bytes of the interval)  ->  LCE  ->  next position
//
// — and the lanes of a warp meet at each of them whatever segment of whatever strand they are on.  (Round 1's resumable /
// sequential state-machine tiers failed on exactly this: every control state carried its own copy of the engines and
// the lanes ran them one after the other; they are gone from the tree, profiles/r1_experiments.md keeps their numbers.)
//
// What makes the single loop possible, following the reference statement by statement (k = 31, read without N):
//   * Phase A (:112-200) looks up the k-mer at 0 and its reverse complement; Phase B's / C's "expected present"
//     look-ups only ever happen at positions 0 and L - k of the strand being scanned (:224-228 after a match that
//     reached the read's end, :300-321 the forced last position, :332-344 the start of Phase C).  The four k-mers
//     involved — read[0,k), read[L-k,L) and their reverse complements — are looked up together up front (as the
//     clean-match kernel does) and answer all of those.
//   * every other look-up is part of a walk over positions whose k-mers are expected to be absent (the windows
//     that cover a mismatch, or the hunt for a first seed): the positions visited do not depend on the results, so
//     QW of them go out per round and the first present one is taken.
//   * a read with hits on both strands (`other` look-up present, :265-272 / :379-385, or both k-mers in Phase A)
//     needs the kmerScores vote (:442-490): declined.
// Declined as well (nothing is emitted, the warp-per-read kernels redo the read from scratch with identical
// results): N in the read, a homopolymer window among the looked-up positions, a k-mer interval starting at SA index
// 0, an open interval wider than Q_WL (Q_W without lcp bytes) in the extension, more than Q_FI intervals, a maximal interval wider than Q_WH in
// the hit enumeration, an interval wider than Q_SPAN in the span filter.
//
// Hit enumeration: cg_warp.cuh::w_hits_lanes' exact shortcut (intervals whose query range lies inside a longer one
// of the same strand never change the intersection) with the candidates in the thread's scratch.
//
// CG_HD throughout: tests/emu runs the same functions on the host in front of the emulated warp tiers.
#pragma once
#include "cg_thrmem.cuh"

#ifndef CG_THR_WALK
#define CG_THR_WALK 2      // table look-ups per thread in a walk round.  Round 1 (lockstep kernels of wt / bsj, ms per 4 M pairs): 8 -> 4.93 / 9.28,
                           // 4 -> 4.72 / 8.17, 2 -> 4.92 / 8.60.  Round 2, with the seed filter leaving 1.5 table rounds per read (step of 8 M pairs):
                           // 2 -> 20.44 ms, 4 -> 20.68, 8 -> 27.02 (spills): the narrower round over-fetches less when the walk ends and is less code
#endif
#ifndef CG_THR_MINB
#define CG_THR_MINB 7      // resident blocks per SM (x 128 threads) asked of the lockstep kernels.  Round 1, DRAM-bound kernels: 4 -> 5.85 / 11.4 ms, 5 -> 5.39 / 10.5,
                           // 6 -> 5.46 / 10.8.  Round 2 (seed filter, lcp bytes: a quarter of the DRAM bytes, the kernels wait on latency), step of 8 M pairs:
                           // 4 -> 30.14 ms, 5 -> 29.85, 6 -> 29.33 (80 registers), 7 -> 30.59 (72, spills), 8 -> 31.28; asking for the largest shared-memory
                           // carve-out instead costs 25 %: the kernels live on L1 hits for text / SA sectors.  After the collector / finish split and with the lean
                           // collector's intervals written straight to global memory (20 words of shared memory per read): 5 -> 21.41 ms, 6 -> 20.81, 7 -> 20.50, 8 -> 20.54
#endif

namespace cgq {
using namespace cg;

#if defined(CG_HOST_EMU)
// tests/emu: what one read costs, by operation (0 loop trips, 1 filter stages, 2 table rounds, 3 suffixes extended, 4 q_match word steps, 5 lce word
// steps, 6 q_tids4 calls, 7 span-filter chunks)
static uint32_t g_cost[8];
#define CG_COST(i, n) (g_cost[i] += (uint32_t)(n))
#else
#define CG_COST(i, n) ((void)0)
#endif

static const int QW = CG_THR_WALK;
#ifndef CG_THR_EXT
#define CG_THR_EXT 4       // suffixes per round of the extension
#endif
static const int QX = CG_THR_EXT;
#ifndef CG_THR_FI
#define CG_THR_FI 8
#endif
static const int Q_FI = CG_THR_FI;        // intervals kept (one strand)
static const int Q_W = 16;        // widest k-mer interval the extension compares suffix by suffix with the text (an index without lcp bytes)
static const int Q_WL = 256;      // widest k-mer interval the extension walks with the lcp bytes (a byte per suffix, the text only where two lengths tie)
static const int Q_WH = 32;       // widest maximal interval in the hit enumeration (candidates, intersection: one bit per candidate)
static const int Q_SPAN = 64;     // widest interval the span filter walks

static const int Q_FK = 16;       // kmerScores entries (two-strand instantiation; <= 16: libstdc++'s sort is a plain insertion sort)

// Per-thread scratch of 32-bit words, word j at base[j * stride] (stride = threads per block on the device; 1 on the host)
//   [0, 2 W2)         forward code words, 16 bases each (W2 = 64-bit words incl. one zero pad word)
//   [2 W2, 4 W2)      reverse-complement strand, same layout
//   then 3 Q_FI       intervals of a strand: begin, end, len | qpos << 16   (VOTE: forward list, then reverse-complement list)
//   VOTE: Q_FK        kmerScores entries (cg::ks_pack)
//   then Q_WH         hit candidates (transcript ids) of a strand            (VOTE: one area per strand; not in the collector kernel's scratch)
template <bool VOTE>
struct QMemT {
    uint32_t* base; int stride; int W2;
    // the lean collector kernel writes its intervals straight to its hand-over slot in global memory (word j of the slot at giv[j * gcap]):
    // it never reads them back, and the shared memory they took (a third of the kernel's footprint) is worth more as L1
    uint32_t* giv = nullptr; uint64_t gcap = 0;
    CG_HD uint32_t ld(int j) const { return base[(size_t)j * stride]; }
    CG_HD void st(int j, uint32_t v) const { base[(size_t)j * stride] = v; }
    // interval n of strand s (a single-strand read keeps one list: entry n of the slot)
    CG_HD void put_iv(int s, int n, uint32_t lb, uint32_t ub, uint32_t lq) const {
        if (!VOTE && giv) { giv[(uint64_t)(3 * n) * gcap] = lb; giv[(uint64_t)(3 * n + 1) * gcap] = ub; giv[(uint64_t)(3 * n + 2) * gcap] = lq; }
        else { st(iv(s, n), lb); st(iv(s, n) + 1, ub); st(iv(s, n) + 2, lq); }
    }
    CG_HD int iv(int s, int i) const { return 4 * W2 + 3 * ((VOTE ? s * Q_FI : 0) + i); }
    CG_HD int ks(int j) const { return 4 * W2 + 3 * Q_FI * 2 + j; }
    CG_HD int cb(int s) const { return 4 * W2 + 3 * Q_FI * (VOTE ? 2 : 1) + (VOTE ? Q_FK + s * Q_WH : 0); }
    CG_HD static int words(int W2) { return 4 * W2 + 3 * Q_FI * (VOTE ? 2 : 1) + (VOTE ? Q_FK + 2 * Q_WH : Q_WH); }
    CG_HD static int words_col(int W2) { return 4 * W2 + (VOTE ? 3 * Q_FI * 2 + Q_FK : 0); }      // the collector kernel: no candidates, and the lean one keeps no intervals (put_iv)
    // 32-base window of strand s at oriented position i (i < L)
    CG_HD uint64_t chunk(int s, int i) const {
        int p = (s ? 2 * W2 : 0) + (i >> 4), sh = (i & 15) * 2;
        uint32_t a = ld(p), b = ld(p + 1), c = ld(p + 2);
        return (uint64_t)cgt::fsr(a, b, sh) | ((uint64_t)cgt::fsr(b, c, sh) << 32);
    }
};
typedef QMemT<false> QMem;

enum { Q_NOTFOUND = 0, Q_FOUND = 1 };
// decline reasons (negative return values; tests/emu counts them)
enum { QD_HOMO = -1, QD_BOTH_A = -2, QD_OTHER = -3, QD_SA0 = -4, QD_WIDE_EXT = -5, QD_NINT = -6, QD_WIDE_HITS = -7, QD_WIDE_SPAN = -8, QD_ODD = -9, QD_N = -10, QD_NKS = -11 };

// longest common prefix of strand s of the read from query offset q and the text at p, compared over [i0, lim)
// (out of line, arguments by value: the lockstep kernels are as sensitive to their code size as the warp kernels — a
// quarter of the stall samples of the first version were instruction-fetch stalls)
template <class M>
CG_HDN int q_match(const u32x4* text, M m, int s, int q, uint32_t p, int i0, int lim) {
    int i = i0;
    uint32_t curb = 0xffffffffu;
    u32x4 cc; cc.x = cc.y = cc.z = cc.w = 0;
    cgt::u32x2 mk; mk.x = mk.y = 0;
    while (i < lim) {
        CG_COST(4, 1);
        const uint32_t tp = p + (uint32_t)i, b = tp >> 6;
        if (b != curb) { cc = CG_LDG(&text[2 * (uint64_t)b]); mk = CG_LDG(reinterpret_cast<const cgt::u32x2*>(&text[2 * (uint64_t)b + 1])); curb = b; }
        const int half = (tp >> 5) & 1, o5 = tp & 31;
        const uint64_t tw = cgt::half_codes(cc, half) >> (2 * o5);
        const uint32_t tm = (half ? mk.y : mk.x) >> o5;
        const int n = cmin(32 - o5, lim - i);
        const uint64_t x = tw ^ m.chunk(s, q + i);
        const uint64_t d = (x | (x >> 1)) & EVEN;
        const int j = cmin(d ? (ctz64(d) >> 1) : 32, tm ? ctz32(tm) : 32);
        if (j < n) return i + j;
        i += n;
    }
    return i;
}

// SASearcher::lce on two text positions (cg::lce after its suffix-array reads)
CG_HDN int q_lce(const u32x4* text, uint32_t n_text, uint32_t o1, uint32_t o2, int startAt, int stopAt, int mode) {
    if (mode == 0) { o1 += (uint32_t)startAt; o2 += (uint32_t)startAt; }
    int len = startAt;
    const uint32_t mx = cmax(o1, o2);
    const int64_t room = (int64_t)n_text - (int64_t)mx;
    const int lim = (int)cmin<int64_t>((int64_t)stopAt, room);
    uint32_t b1c = 0xffffffffu, b2c = 0xffffffffu;
    u32x4 c1, c2; c1.x = c1.y = c1.z = c1.w = 0; c2 = c1;
    cgt::u32x2 m1, m2; m1.x = m1.y = 0; m2 = m1;
    while (len < lim) {
        CG_COST(5, 1);
        const uint32_t a = o1 + (uint32_t)len, b = o2 + (uint32_t)len;
        if ((a >> 6) != b1c) { b1c = a >> 6; c1 = CG_LDG(&text[2 * (uint64_t)b1c]); m1 = CG_LDG(reinterpret_cast<const cgt::u32x2*>(&text[2 * (uint64_t)b1c + 1])); }
        if ((b >> 6) != b2c) { b2c = b >> 6; if (b2c == b1c) { c2 = c1; m2 = m1; } else { c2 = CG_LDG(&text[2 * (uint64_t)b2c]); m2 = CG_LDG(reinterpret_cast<const cgt::u32x2*>(&text[2 * (uint64_t)b2c + 1])); } }
        const int ha = (a >> 5) & 1, oa = a & 31, hb = (b >> 5) & 1, ob = b & 31;
        const uint64_t wa = cgt::half_codes(c1, ha) >> (2 * oa), wb = cgt::half_codes(c2, hb) >> (2 * ob);
        const uint32_t ma = (ha ? m1.y : m1.x) >> oa, mb = (hb ? m2.y : m2.x) >> ob;
        const int n = cmin(cmin(32 - oa, 32 - ob), lim - len);
        const uint64_t x = wa ^ wb;
        const uint64_t d = (x | (x >> 1)) & EVEN;
        const uint32_t mm = ma | mb;
        const int j = cmin(d ? (ctz64(d) >> 1) : 32, mm ? ctz32(mm) : 32);
        if (j < n) return len + j;
        len += n;
    }
    return len;
}

struct QProbe { bool f; uint32_t b, e; };

// the generic probe loop from slot h on (what the two straight-line passes of a walk round leave open): x = found, y / z = interval
CG_HDN u32x4 q_resolve_from(const u32x4* hslots, uint64_t hmask, uint64_t key, uint64_t h) {
    u32x4 r; r.x = r.y = r.z = r.w = 0;
    for (;;) {
        const u32x4 sl = CG_LDG(&hslots[h]);
        const uint64_t kk = (uint64_t)sl.x | ((uint64_t)sl.y << 32);
        if (kk == key) { r.x = 1; r.y = sl.z; r.z = sl.w; return r; }
        if (kk == ~0ULL) return r;
        h = (h + 1) & hmask;
    }
}

// The position walk (:112-132 with both = true: k-mer and reverse complement, a position counts when either is there;
// :234-261 / :341-372 after a mismatch with both = false): positions o .. lim of strand s.
// Returns the first position that hits (F: the k-mer's own look-up, hasR: its reverse complement's, both only) or -1;
// homo is set when a looked-up window is a homopolymer (the caller declines).
//
// A walk runs over windows that are NOT in the table — the 31 windows that cover a substitution, the stretch of a read that
// lies outside every BSJ reference — and each of them used to cost one random DRAM access (the table slot) to learn that.
// With the index's seed filter (IndexView::fbits: which fm-mers occur in the text, fm = 16) a stage of the walk first looks up
// Q_NP fm-mers of the read, S = k - fm positions apart, all in flight together: an fm-mer the text does not hold settles the
// S + 1 windows that contain it (no k-mer of the table can contain it, nor can a reverse complement: the bit set is closed
// under it), so three accesses answer 46 windows.  Only the windows left alive are looked up in the table, QW per round as
// before.  Exactness: the filter only ever removes windows that are absent from the table, and a window it removes is not a
// homopolymer (a homopolymer fm-mer is never used as evidence), so the positions visited, the first hit and the homopolymer
// decline are those of the unfiltered walk.
#if defined(CG_HOST_EMU)
struct WalkStats { uint64_t filter_loads = 0, table_loads = 0, dead_windows = 0, calls = 0, rounds = 0, found = 0, hist[24] = {0}; };     // tests/emu: what the walks of a run cost
static WalkStats g_walk;
#define CG_WALK_STAT(f, n) (g_walk.f += (uint64_t)(n))
struct WalkScope { int r = 0; ~WalkScope() { g_walk.calls += 1; g_walk.rounds += (uint64_t)r; g_walk.hist[r < 23 ? r : 23] += 1; } };
#define CG_WALK_SCOPE WalkScope walk_scope_
#define CG_WALK_ROUND (++walk_scope_.r, CG_COST(2, 1))
#else
#define CG_WALK_SCOPE ((void)0)
#define CG_WALK_ROUND ((void)0)
#define CG_WALK_STAT(f, n) ((void)0)
#endif
// Geometry of a filter stage: Q_NP fm-mers at o0 + Q_D0 + Q_SP j, all in flight together.  An fm-mer at p is part of the windows p - S .. p
// (S = k - fm), so the stage decides the windows o0 .. o0 + Q_D0 + Q_SP (Q_NP - 1); the table stage then goes on over Q_TAIL more windows
// taken as alive (the window after a run of dead ones is the one a walk usually ends on: no second stage for it).
#ifndef CG_THR_FNP
#define CG_THR_FNP 6       // measured (lockstep kernels of wt / bsj, ms per 4 M pairs of config 2): 3 fm-mers S apart 3.10 / 5.56, 6 fm-mers S / 3 apart 2.84 / 5.03, 4 fm-mers S / 2 apart
                           // in between: with S = 15 a stage of 6 decides 31 windows — the windows that cover one substitution — and a window holds three of the
                           // fm-mers, so the stretches where the read enters a region of the text (fm-mers present, k-mers not yet) leave 5 windows alive instead of 15
#endif
#ifndef CG_THR_FD0
#define CG_THR_FD0 0       // 0: S / CG_THR_FDIV
#endif
#ifndef CG_THR_FSP
#define CG_THR_FSP 0       // 0: S / CG_THR_FDIV
#endif
#ifndef CG_THR_FDIV
#define CG_THR_FDIV 3
#endif
#ifndef CG_THR_FTAIL
#define CG_THR_FTAIL 0
#endif
#ifndef CG_THR_F2L
#define CG_THR_F2L 3       // > 1: two-level filter stage, every CG_THR_F2L-th probe first (measured: 1 -> 20.40 ms per 8 M pairs, 3 -> 20.18; filter probes per
                           // Circall_bsj record 12.5 -> 8.1 in the host emulation)
#endif
static const int Q_F2L = CG_THR_F2L;
#if defined(CG_HOST_EMU)
static const int Q_NP = 8;
#else
static const int Q_NP = CG_THR_FNP;        // seed-filter look-ups per stage (Q_D0 + Q_SP (Q_NP - 1) + Q_TAIL + 1 windows must fit 64 bits)
#endif
#if defined(CG_HOST_EMU)
static int Q_D0X = CG_THR_FD0, Q_SPX = CG_THR_FSP, Q_TAIL = CG_THR_FTAIL, Q_NPR = CG_THR_FNP, Q_DIV = CG_THR_FDIV;      // tests/emu sweeps them (Q_NPR <= Q_NP)
#else
static const int Q_D0X = CG_THR_FD0, Q_SPX = CG_THR_FSP, Q_TAIL = CG_THR_FTAIL, Q_NPR = CG_THR_FNP, Q_DIV = CG_THR_FDIV;
#endif
template <class M>
CG_HD int q_walk(const IndexView& ix, const M& m, int L, int s, int o, int lim, bool both, QProbe& F, bool& hasR, bool& homo) {
    const int k = ix.k;
    const int sh = both ? 1 : 0, per = QW >> sh;
    const int fm = ix.fm, S = k - fm;
    const int Q_D0 = Q_D0X > 0 ? Q_D0X : cmax(1, S / Q_DIV), Q_SP = Q_SPX > 0 ? Q_SPX : cmax(1, S / Q_DIV);
    F.f = false; F.b = F.e = 0; hasR = false;
    CG_WALK_SCOPE;
    // ONE loop, one filter stage and one table round per trip at most, so that the lanes of a warp that are inside a walk do their
    // rounds together (with the table rounds in a loop of their own under the stages, and a trip of that loop spent on skipping the
    // windows the filter settled, the rounds of different lanes fell into different trips: 3.8 active lanes in a round, ncu r2j)
    const bool usef = fm && Q_D0 + Q_SP * (Q_NPR - 1) + Q_TAIL + QW < 64;
    int o0 = o, hi = usef ? o - 1 : lim;                     // the windows o0 .. hi are decided (dead bit j = window o0 + j is certainly not in the table)
    uint64_t dead = 0;
    while (o <= lim) {
        if (usef && o > hi) {                                // ---- filter stage
            o0 = o; dead = 0;
            CG_COST(1, 1);
            const uint32_t fmk = fm_mask(fm);
            uint32_t x[Q_NP], wv[Q_NP];
            bool ok[Q_NP];
#pragma unroll
            for (int j = 0; j < Q_NP; ++j) {
                const int p = o0 + Q_D0 + Q_SP * j;                            // this fm-mer lies inside the windows p - S .. p
                x[j] = 0; wv[j] = 0xffffffffu; ok[j] = false;
                if (j < Q_NPR && p + fm <= L && p - S <= lim) {
                    const int wb = (s ? 2 * m.W2 : 0) + (p >> 4);
                    x[j] = cgt::fsr(m.ld(wb), m.ld(wb + 1), (p & 15) * 2) & fmk;
                    ok[j] = !fm_homopolymer(x[j], fm);
                }
            }
            // Two levels (Q_F2L = n > 1): every n-th probe first — S / Q_DIV apart they are S apart when n = Q_DIV, so that together they
            // lie inside every window of the stage — and the others only when one of those is in the text.  After a substitution the
            // first level settles the whole stage four times out of five: 2 DRAM accesses instead of 6, and the collectors run at the
            // memory system's random-access rate.
            bool more = Q_F2L <= 1;
#pragma unroll
            for (int j = 0; j < Q_NP; ++j)
                if (Q_F2L > 1 && (j + 1) % Q_F2L == 0 && ok[j]) { wv[j] = CG_LDG(&ix.fbits[x[j] >> 5]); CG_WALK_STAT(filter_loads, 1); more |= ((wv[j] >> (x[j] & 31)) & 1u) != 0; }
            if (more) {
#pragma unroll
                for (int j = 0; j < Q_NP; ++j)
                    if ((Q_F2L <= 1 || (j + 1) % Q_F2L != 0) && ok[j]) { wv[j] = CG_LDG(&ix.fbits[x[j] >> 5]); CG_WALK_STAT(filter_loads, 1); }
            }
            const uint64_t full = (2ULL << S) - 1ULL;
#pragma unroll
            for (int j = 0; j < Q_NP; ++j)
                if (!((wv[j] >> (x[j] & 31)) & 1u)) { const int lo = Q_D0 + Q_SP * j - S; dead |= lo >= 0 ? full << lo : full >> -lo; }
            hi = cmin(lim, o0 + Q_D0 + Q_SP * (Q_NPR - 1) + Q_TAIL);
        }
        // ---- table round over the next windows that are still alive
        {
            const uint64_t dm0 = usef ? dead >> (o - o0) : 0ULL;
            const int sk = ctz64(~dm0);                       // windows the filter settled (bits past the decided range are 0)
            CG_WALK_STAT(dead_windows, cmin(sk, hi - o + 1));
            o += sk;
        }
        if (o <= hi) {
            const uint64_t dm = usef ? dead >> (o - o0) : 0ULL;
            uint64_t key[QW]; u32x4 sl[QW]; uint32_t hs[QW];
            CG_WALK_ROUND;
            // the (at most) QW windows of this round start within QW bases of each other: they are cut out of four words of the
            // strand read once, instead of three scratch loads per window
            const int wb = (s ? 2 * m.W2 : 0) + (o >> 4), sh0 = (o & 15) * 2;
            const uint32_t w0 = m.ld(wb), w1 = m.ld(wb + 1), w2 = m.ld(wb + 2), w3 = m.ld(wb + 3);
#pragma unroll
            for (int i = 0; i < QW; ++i) {
                const int j = i >> sh;                                             // position o + j
                const bool in = o + j <= lim && !((dm >> j) & 1ULL);
                const int bit = sh0 + 2 * j, hi5 = bit >> 5, sf = bit & 31;        // bit < 46: the window starts in word 0 or 1
                const uint32_t a = hi5 ? w1 : w0, b = hi5 ? w2 : w1, c = hi5 ? w3 : w2;
                const uint64_t mer = ((uint64_t)cgt::fsr(a, b, sf) | ((uint64_t)cgt::fsr(b, c, sf) << 32)) & ix.km;
                if (!(both && (i & 1))) homo |= in && homopolymer_ix(mer, ix);
                key[i] = (both && (i & 1)) ? revcomp(mer, k) : mer;
                hs[i] = (uint32_t)(home_slot(ix, key[i]) >> 1);                    // home slots are even: the sector number fits 32 bits up to 2^33 slots
                sl[i].x = sl[i].y = 0xffffffffu; sl[i].z = sl[i].w = 0;            // an empty slot: not found
                if (in) { sl[i] = CG_LDG(&ix.hslots[2 * (uint64_t)hs[i]]); CG_WALK_STAT(table_loads, 1); }
            }
            if (homo) return -1;
            // Resolve all of them together.  A home slot is the even slot of a 32-byte sector and holds another k-mer for a third
            // of the look-ups; the odd slot of the same sector (an L1 hit) settles almost all of those.  Both steps are
            // straight-line code for the whole warp; only what is still open afterwards takes the generic probe loop.
            uint32_t fb = 0, pend = 0;
#pragma unroll
            for (int i = 0; i < QW; ++i) {
                const uint64_t kk = (uint64_t)sl[i].x | ((uint64_t)sl[i].y << 32);
                if (kk == key[i]) fb |= 1u << i;
                else if (kk != ~0ULL) pend |= 1u << i;
            }
            if (pend) {
#pragma unroll
                for (int i = 0; i < QW; ++i) if ((pend >> i) & 1u) sl[i] = CG_LDG(&ix.hslots[2 * (uint64_t)hs[i] + 1]);
#pragma unroll
                for (int i = 0; i < QW; ++i) {
                    if ((pend >> i) & 1u) {
                        const uint64_t kk = (uint64_t)sl[i].x | ((uint64_t)sl[i].y << 32);
                        if (kk == key[i]) { fb |= 1u << i; pend &= ~(1u << i); }
                        else if (kk == ~0ULL) pend &= ~(1u << i);
                    }
                }
                // both slots of the home sector taken by other k-mers (up to 9 % of the look-ups at this table's load): the two slots
                // of the next sector, again for all look-ups together — one lane at a time in the probe loop this was 16 % of the
                // kernel's time — and the loop only for what is left after that
#pragma unroll 1
                for (int step = 2; step < 4 && pend; ++step) {
#pragma unroll
                    for (int i = 0; i < QW; ++i) if ((pend >> i) & 1u) sl[i] = CG_LDG(&ix.hslots[(2 * (uint64_t)hs[i] + step) & ix.hmask]);
#pragma unroll
                    for (int i = 0; i < QW; ++i) {
                        if ((pend >> i) & 1u) {
                            const uint64_t kk = (uint64_t)sl[i].x | ((uint64_t)sl[i].y << 32);
                            if (kk == key[i]) { fb |= 1u << i; pend &= ~(1u << i); }
                            else if (kk == ~0ULL) pend &= ~(1u << i);
                        }
                    }
                }
                if (pend) {
#pragma unroll
                    for (int i = 0; i < QW; ++i) {
                        if ((pend >> i) & 1u) {
                            const u32x4 r = q_resolve_from(ix.hslots, ix.hmask, key[i], (2 * (uint64_t)hs[i] + 4) & ix.hmask);
                            if (r.x) { fb |= 1u << i; sl[i].z = r.y; sl[i].w = r.z; }
                        }
                    }
                }
            }
            // the first position that hits; sl[i].z / .w are the interval of look-up i when it was found
            int pos = -1;
#pragma unroll
            for (int i = QW - 1; i >= 0; --i) {
                if (both) {
                    if (!(i & 1) && ((fb >> i) & 3u)) { pos = o + (i >> 1); F.f = (fb >> i) & 1u; F.b = sl[i].z; F.e = sl[i].w; hasR = (fb >> (i + 1)) & 1u; }
                } else if ((fb >> i) & 1u) { pos = o + i; F.f = true; F.b = sl[i].z; F.e = sl[i].w; }
            }
            if (pos >= 0) return pos;
            o += per;
        }
    }
    return -1;
}

// SACollectorCircall::operator() up to and including the strand vote, as three pieces: q_begin (the four up-front look-ups), q_step
// (one trip of the reference's scan loops: LOOK -> EXTEND -> LCE -> next position), q_vote (:442-490).  k_thr_col strings them
// together for one read per thread, q_collect for the host emulation.  (The pieces exist because a persistent-lane kernel — a lane
// that finishes a read draws the next one — was measured: 1.7 x slower, profiles/r2_experiments.md.)
//   VOTE = false: reads whose hits lie on one strand (a look-up that scores for the other strand declines the read)
//   VOTE = true : both strands — Phase B then Phase C when rcHit >= fwdHit (:332), the kmerScores list and its vote (:442-490)
// Returns Q_FOUND (nF / nR post-vote intervals in the scratch, in the reference's order), Q_NOTFOUND (:204) or a QD_* reason.
// known: which of the four up-front k-mers are in the table, as the kernel in front of this tier already found out (bit 0 read[0,k),
// bit 1 its reverse complement, bit 2 read[L-k,L), bit 3 its reverse complement; the look-up class minus one) — an absent one is
// not looked up again; < 0: not known.
enum { Q_CONT = 2, Q_DONE = 3 };
struct QState {
    QProbe F0, R0, FL, RL;            // read[0,k), its reverse complement, read[L-k,L), its reverse complement
    int L, s, o, nKS, nF, nR;
    uint32_t fwdHit, rcHit;
    bool phaseA, firstB, aOther, expectAbsent, lastSearch;
};

template <bool VOTE>
CG_HD int q_begin(const IndexView& ix, const QMemT<VOTE>& m, int L, int known, QState& st) {
    const int k = ix.k;
    st.L = L; st.nF = st.nR = 0;
    if (L > 0xffff) return QD_ODD;
    if (!(k < L)) return Q_NOTFOUND;       // Phase A's loop needs rb + k < L (:112)
    // the four look-ups every path starts with: read[0,k), read[L-k,L) and their reverse complements
    {
        const uint64_t mer0 = m.chunk(0, 0) & ix.km, merL = m.chunk(0, L - k) & ix.km;
        if (homopolymer_ix(mer0, ix) || homopolymer_ix(merL, ix)) return QD_HOMO;
        const uint64_t rc0 = revcomp(mer0, k), rcL = revcomp(merL, k);
        const uint64_t h0 = home_slot(ix, mer0), h1 = home_slot(ix, rc0), h2 = home_slot(ix, merL), h3 = home_slot(ix, rcL);
        u32x4 s0, s1, s2, s3;
        s0.x = s0.y = 0xffffffffu; s0.z = s0.w = 0; s1 = s0; s2 = s0; s3 = s0;             // an empty slot: not found
        if (known < 0 || (known & 1)) s0 = CG_LDG(&ix.hslots[h0]);
        if (known < 0 || (known & 2)) s1 = CG_LDG(&ix.hslots[h1]);
        if (known < 0 || (known & 4)) s2 = CG_LDG(&ix.hslots[h2]);
        if (known < 0 || (known & 8)) s3 = CG_LDG(&ix.hslots[h3]);
        st.F0.b = st.F0.e = st.R0.b = st.R0.e = st.FL.b = st.FL.e = st.RL.b = st.RL.e = 0;
        st.F0.f = resolve(ix, mer0, h0, s0, st.F0.b, st.F0.e); st.R0.f = resolve(ix, rc0, h1, s1, st.R0.b, st.R0.e);
        st.FL.f = resolve(ix, merL, h2, s2, st.FL.b, st.FL.e); st.RL.f = resolve(ix, rcL, h3, s3, st.RL.b, st.RL.e);
    }
    st.phaseA = true; st.firstB = false; st.aOther = false; st.expectAbsent = false; st.lastSearch = false;
    st.s = 0; st.o = 0; st.nKS = 0; st.fwdHit = 0; st.rcHit = 0;
    return Q_CONT;
}

// one trip of the scan loop: Q_CONT (more to do), Q_DONE (both strands finished: q_vote next), Q_NOTFOUND or a QD_* reason
template <bool VOTE>
CG_HD int q_step(const IndexView& ix, const QMemT<VOTE>& m, int lceMode, QState& st) {
    const int k = ix.k, skipOverlap = k - 1, L = st.L;
#define Q_PUSH_KS(kp, f, r) do { if (VOTE) { if (st.nKS >= Q_FK) return QD_NKS; m.st(m.ks(st.nKS), ks_pack((kp), (f), (r))); ++st.nKS; } } while (0)
    bool done = false;                                                         // this strand is finished
    bool homo = false;
    CG_COST(0, 1);
    // ---------------- LOOK: the next position of strand s whose k-mer is in the table ----------------
    int pos = -1;
    QProbe hit; hit.f = false; hit.b = hit.e = 0;
    bool other = false;
    bool walk; int wlim;
    if (st.phaseA) {
        wlim = L - k - 1;                                                      // `re < read.end()` (:112)
        if (st.F0.f || st.R0.f) { pos = 0; hit = st.F0; other = st.R0.f; walk = false; }   // :131-132 at position 0
        else { walk = true; st.o = 1; }
    } else {
        wlim = L - k;                                                          // :234 / :341-344
        walk = st.expectAbsent;
        if (!walk) {
            // expected present (:261+:268 / :372+:379): only ever asked at 0 or L - k, answered up front
            QProbe a, b;
            if (st.o == 0) { if (st.s == 0) { a = st.F0; b = st.R0; } else { a = st.RL; b = st.FL; } }
            else if (st.o == L - k) { if (st.s == 0) { a = st.FL; b = st.RL; } else { a = st.R0; b = st.F0; } }
            else return QD_ODD;
            if (a.f) { pos = st.o; hit = a; other = b.f; }
            else { st.o += 1; walk = true; }                                   // :324 / :436, then a walk
        }
    }
    if (walk) {
        bool hasR = false;
        pos = q_walk(ix, m, L, st.s, st.o, wlim, st.phaseA, hit, hasR, homo);
        if (homo) return QD_HOMO;
        if (pos >= 0) {
            if (st.phaseA) other = hasR;
            else {                                                             // :268 / :379
                uint32_t b2, e2;
                other = probe(ix, revcomp(m.chunk(st.s, pos) & ix.km, k), b2, e2);
            }
        }
    }
    if (pos < 0) {
        if (st.phaseA) return Q_NOTFOUND;                                      // :204
        done = true;                                                           // ran off the read
    }
    if (!done) {
        st.o = pos;
        bool extendIt = true;
        if (st.phaseA) {
            st.phaseA = false;
            if (!VOTE && hit.f && other) return QD_BOTH_A;                     // both strands seeded: the vote decides
            if (hit.f) { st.s = 0; st.firstB = true; st.aOther = other; }      // :134-176, Phase B follows
            else {                                                             // :178-191: rcHit only, Phase C from 0 (:332-346)
                ++st.rcHit; Q_PUSH_KS(pos, -1, 1);
                st.s = 1; st.o = 0; st.expectAbsent = false; extendIt = false;
            }
        } else {
            if (!VOTE && other) return QD_OTHER;                               // :265-272 / :381-385: the other strand scores too
            const int fpos = st.s == 0 ? st.o : L - (st.o + k);                // :366
            if (st.s == 0) { ++st.fwdHit; if (other) ++st.rcHit; Q_PUSH_KS(fpos, 1, other ? 1 : 0); }
            else { ++st.rcHit; if (other) ++st.fwdHit; Q_PUSH_KS(fpos, other ? 1 : 0, 1); }
        }
        if (extendIt) {
            // ---------------- EXTEND: extendSearchNaive(b - 1, e) from position o (:143 / :280 / :393) ----------------
            const int s = st.s, o = st.o;
            if (hit.b == 0) return QD_SA0;
            uint32_t lb = hit.b, ub = hit.e;
            int matched = k;
            const int mq = L - o;
            if (mq > k) {
                const uint32_t w = hit.e - hit.b;
                if (w > (uint32_t)(ix.lcp ? Q_WL : Q_W)) return QD_WIDE_EXT;
                CG_COST(3, w);
                int best = -1; uint32_t first = 0, last = 0;
                if (ix.lcp) {
                    // The suffixes are sorted: with lp the match length of the one before and h = lcp of the two, this one matches
                    // lp when h > lp and h when h < lp; only h == lp (or a saturated byte) needs the text, from there on.
                    int lp = q_match(ix.text, m, s, o, CG_LDG(&ix.sa[hit.b]), k, mq);
                    best = lp;
                    for (uint32_t c0 = 1; c0 < w; c0 += QX) {
                        int hb[QX];
#pragma unroll
                        for (int c = 0; c < QX; ++c) hb[c] = (c0 + c < w) ? (int)CG_LDG(&ix.lcp[hit.b + c0 + c]) : 0;
#pragma unroll
                        for (int c = 0; c < QX; ++c) {
                            if (c0 + c < w) {
                                const int h = hb[c];
                                int l;
                                if (h > lp) l = lp;
                                else if (h < lp && h < 255) l = h;
                                else l = q_match(ix.text, m, s, o, CG_LDG(&ix.sa[hit.b + c0 + c]), cmin(h, lp), mq);
                                if (l > best) { best = l; first = last = c0 + c; }
                                else if (l == best) last = c0 + c;
                                lp = l;
                            }
                        }
                    }
                } else
                for (uint32_t c0 = 0; c0 < w; c0 += QX) {
                    uint32_t p[QX];
#pragma unroll
                    for (int c = 0; c < QX; ++c) p[c] = (c0 + c < w) ? CG_LDG(&ix.sa[hit.b + c0 + c]) : 0u;
#pragma unroll
                    for (int c = 0; c < QX; ++c) {
                        if (c0 + c < w) {
                            const int l = q_match(ix.text, m, s, o, p[c], k, mq);
                            if (l > best) { best = l; first = last = c0 + c; }
                            else if (l == best) last = c0 + c;
                        }
                    }
                }
                lb = hit.b + first; ub = hit.b + last + 1; matched = best;
            }
            const bool valid = ub > lb && ub - lb < 1000u;                     // :149 / :285 / :398 (maxInterval)
            if (valid) {
                const int n = s == 0 ? st.nF : st.nR;
                if (n >= Q_FI) return QD_NINT;
                m.put_iv(s, n, lb, ub, (uint32_t)matched | ((uint32_t)o << 16));
                if (s == 0) st.nF = n + 1; else st.nR = n + 1;
            }
            if (st.firstB) {                                                   // Phase A's bookkeeping (:149-191)
                if (valid) {
                    ++st.fwdHit;
                    if (st.aOther) { ++st.rcHit; Q_PUSH_KS(o, 1, 1); } else Q_PUSH_KS(o, 1, -1);
                    if (o + matched < L) Q_PUSH_KS(o + matched - skipOverlap, -1, 0);      // :165-168
                } else return QD_ODD;                                          // (an interval of >= 1000 suffixes: not this tier's read)
            } else if (valid && o + matched < L) {                             // :291 / :404
                Q_PUSH_KS(s == 0 ? o + matched - skipOverlap : L - (o + matched), 0, 0);
            }
            if (st.lastSearch) done = true;                                    // :300 / :412
            else {
                // ---------------- LCE and the next position (:219-228 after Phase A, :302-321 / :414-433 in the loop) ----------------
                const int mismatch = o + matched;
                const int stopAt = L - mismatch;
                int l = matched;
                if (stopAt > matched) {                                        // else lce returns startAt at once
                    const uint32_t p1 = CG_LDG(&ix.sa[lb]), p2 = (ub - 1 == lb) ? p1 : CG_LDG(&ix.sa[ub - 1]);
                    l = q_lce(ix.text, ix.n, p1, p2, matched, stopAt, lceMode);
                }
                if (st.firstB) {
                    st.firstB = false;
                    const int fwdSkip = cmax(matched - skipOverlap, l - skipOverlap);
                    const int o2 = cmin(cmax(0, L - k), o + fwdSkip);          // :224-228
                    st.expectAbsent = stopAt > 0 && o2 < L - k;
                    st.o = o2;
                } else if (mismatch < L) {
                    st.o = cmax(o + l - skipOverlap, mismatch - skipOverlap);
                    if (st.o > L - k) { st.o = L - k; st.lastSearch = true; st.expectAbsent = false; }
                    else st.expectAbsent = true;
                } else { st.lastSearch = true; st.o = L - k; st.expectAbsent = false; }
                if (st.o + k > L) done = true;
            }
        }
    }
#undef Q_PUSH_KS
    if (done) {
        if (VOTE && st.s == 0 && st.rcHit >= st.fwdHit) {                      // Phase C after Phase B (:332-346)
            st.s = 1; st.o = 0; st.expectAbsent = false; st.lastSearch = false;
            return Q_CONT;
        }
        return Q_DONE;
    }
    return Q_CONT;
}

// the strand vote (:442-490)
template <bool VOTE>
CG_HD int q_vote(const IndexView& ix, const QMemT<VOTE>& m, QState& st) {
    const int k = ix.k, nKS = st.nKS;
    if (VOTE) {
        if (st.fwdHit > 0 && st.rcHit == 0) st.nR = 0;
        else if (st.rcHit > 0 && st.fwdHit == 0) st.nF = 0;
        else {
            for (int i = 1; i < nKS; ++i) {                                        // std::sort of <= 16 entries: a stable insertion sort
                const uint32_t v = m.ld(m.ks(i)); int j = i - 1;
                while (j >= 0 && ks_key(v) < ks_key(m.ld(m.ks(j)))) { m.st(m.ks(j + 1), m.ld(m.ks(j))); --j; }
                m.st(m.ks(j + 1), v);
            }
            int fs = 0, rs = 0;
            uint32_t prevKey = 0xffffffffu;
            for (int i = 0; i < nKS; ++i) {                                        // std::unique keeps the first of each run
                const uint32_t en = m.ld(m.ks(i));
                if (ks_key(en) == prevKey) continue;
                prevKey = ks_key(en);
                int f = ks_f(en), r = ks_r(en);
                if (f == 0 || r == 0) {
                    const uint64_t w = m.chunk(0, (int)ks_key(en)) & ix.km;       // (no N in this tier's reads)
                    bool pf = false, pr = false; uint32_t b1, e1, b2, e2;
                    probe2(ix, w, revcomp(w, k), pf, b1, e1, pr, b2, e2);
                    if (f == 0) f = pf ? 1 : -1;                                   // :461-464
                    if (r == 0) r = pr ? 1 : -1;                                   // :469-473
                }
                fs += f; rs += r;
            }
            if (fs > rs) st.nR = 0; else if (rs > fs) st.nF = 0;                   // :482-488
        }
    }
    return Q_FOUND;
}

template <bool VOTE>
CG_HD int q_collect(const IndexView& ix, const QMemT<VOTE>& m, int L, int lceMode, int& nF, int& nR, int known = -1) {
    QState st;
    int r = q_begin<VOTE>(ix, m, L, known, st);
    while (r == Q_CONT) r = q_step<VOTE>(ix, m, lceMode, st);
    if (r == Q_DONE) r = q_vote<VOTE>(ix, m, st);
    nF = st.nF; nR = st.nR;
    if (r != Q_FOUND) nF = nR = 0;
    return r;
}

// up to four SA positions -> transcript ids (x, y, z, w; 0xffffffff past nn), the suffix-array entries then the text sectors as
// two waves of loads
CG_HDN u32x4 q_tids4(const uint32_t* sa, const u32x4* text, const uint32_t* satid, uint32_t i0, int nn) {
    CG_COST(6, 1);
    if (satid) {                                         // the index keeps transcriptAtPosition(SA[i]) per suffix-array index: one wave, contiguous
        u32x4 r;
        r.x = CG_LDG(&satid[i0]);
        r.y = nn > 1 ? CG_LDG(&satid[i0 + 1]) : 0xffffffffu;
        r.z = nn > 2 ? CG_LDG(&satid[i0 + 2]) : 0xffffffffu;
        r.w = nn > 3 ? CG_LDG(&satid[i0 + 3]) : 0xffffffffu;
        return r;
    }
    uint32_t p[4], tid[4]; u32x4 t[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) p[x] = x < nn ? CG_LDG(&sa[i0 + x]) : 0u;
#pragma unroll
    for (int x = 0; x < 4; ++x) { t[x].x = t[x].y = t[x].z = t[x].w = 0; if (x < nn) t[x] = CG_LDG(&text[2 * (uint64_t)(p[x] >> 6) + 1]); }
#pragma unroll
    for (int x = 0; x < 4; ++x) tid[x] = x < nn ? cgt::tid_from(t[x], p[x]) : 0xffffffffu;
    u32x4 r; r.x = tid[0]; r.y = tid[1]; r.z = tid[2]; r.w = tid[3];
    return r;
}

CG_HD int q_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// the n intervals of strand s in the scratch -> number of hit transcripts (intersectSAHits + collectHitsSimpleSA, or the
// single-interval enumeration, :492-563); the candidates stay in the scratch (cb(s) + c, c < nc) with `alive` marking the
// hits.  < 0: declined.
template <class M>
CG_HD int q_hits(const IndexView& ix, const M& m, int s, int n, uint32_t& alive, int& nc) {
    alive = 0; nc = 0;
    if (n == 0) return 0;
    uint32_t keep = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t lqi = m.ld(m.iv(s, i) + 2);
        const int li = (int)(lqi & 0xffffu), qi = (int)(lqi >> 16);
        bool drop = false;
        for (int y = 0; y < n; ++y) {
            const uint32_t lqy = m.ld(m.iv(s, y) + 2);
            const int ly = (int)(lqy & 0xffffu), qy = (int)(lqy >> 16);
            drop |= (ly > li && qy <= qi && qy + ly >= qi + li);
        }
        if (!drop) keep |= 1u << i;
    }
    int mi = -1; uint32_t msp = 0xffffffffu; bool wide = false;
    for (int i = 0; i < n; ++i) {
        if (!((keep >> i) & 1u)) continue;
        const uint32_t sp = m.ld(m.iv(s, i) + 1) - m.ld(m.iv(s, i));
        if (sp > (uint32_t)Q_WH) wide = true;
        if (sp < msp) { msp = sp; mi = i; }
    }
    if (wide || mi < 0) return QD_WIDE_HITS;
    const uint32_t mb = m.ld(m.iv(s, mi)), me = m.ld(m.iv(s, mi) + 1);
    const int cb = m.cb(s);
    for (uint32_t c0 = 0; c0 < msp; c0 += 4) {
        const int nn = (int)cmin<uint32_t>(4u, msp - c0);
        const u32x4 t = q_tids4(ix.sa, ix.text, ix.satid, mb + c0, nn);
        m.st(cb + (int)c0, t.x);
        if (nn > 1) m.st(cb + (int)c0 + 1, t.y);
        if (nn > 2) m.st(cb + (int)c0 + 2, t.z);
        if (nn > 3) m.st(cb + (int)c0 + 3, t.w);
    }
    nc = (int)msp;
    for (int c = 0; c < nc; ++c) {                          // one candidate per distinct transcript
        const uint32_t tc = m.ld(cb + c);
        bool dup = false;
        for (int y = 0; y < c; ++y) dup |= m.ld(cb + y) == tc;
        if (!dup) alive |= 1u << c;
    }
    for (int x = 0; x < n; ++x) {
        if (x == mi || !((keep >> x) & 1u)) continue;
        const uint32_t ib = m.ld(m.iv(s, x)), ie = m.ld(m.iv(s, x) + 1);
        if (ib == mb && ie == me) continue;                  // same positions: cannot drop a transcript
        uint32_t seen = 0;
        for (uint32_t c0 = ib; c0 < ie; c0 += 4) {
            const int nn = (int)cmin<uint32_t>(4u, ie - c0);
            const u32x4 t = q_tids4(ix.sa, ix.text, ix.satid, c0, nn);
            for (int c = 0; c < nc; ++c) {
                const uint32_t tc = m.ld(cb + c);
                if (tc == t.x || tc == t.y || tc == t.z || tc == t.w) seen |= 1u << c;          // unused entries are 0xffffffff
            }
        }
        alive &= seen;
    }
    return q_popc(alive);
}

// hits of both strands: each strand's candidates compacted to its hits (cb(s)[0 .. h_s)); returns hF + hR - common, the size
// of the list merged by transcript (:567-579), or < 0.  VOTE = false: only the strand that holds intervals.
template <bool VOTE>
CG_HD int q_hits2(const IndexView& ix, const QMemT<VOTE>& m, int nF, int nR, int& hF, int& hR) {
    hF = hR = 0;
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int n = s == 0 ? nF : nR;
        if (n == 0) continue;
        uint32_t alive; int nc;
        const int h = q_hits(ix, m, s, n, alive, nc);
        if (h < 0) return h;
        int j = 0;
        for (int c = 0; c < nc; ++c) if ((alive >> c) & 1u) { m.st(m.cb(s) + j, m.ld(m.cb(s) + c)); ++j; }
        if (s == 0) hF = h; else hR = h;
    }
    int common = 0;
    if (VOTE && hF && hR)
        for (int a = 0; a < hF; ++a) { const uint32_t t = m.ld(m.cb(0) + a); for (int b = 0; b < hR; ++b) common += m.ld(m.cb(1) + b) == t; }
    return hF + hR - common;
}

// one mate of Circall_wt (src/Circall_wt.cpp:275-342 / 283-517) after the collector returned r (>= 0) with nF / nR intervals in the
// scratch: false = declined
template <bool VOTE>
CG_HD bool q_wt_finish(const IndexView& ix, const QMemT<VOTE>& m, int r, int pairLen, int nF, int nR, uint8_t& flags, uint32_t& nh, int& why) {
    flags = 0; nh = 0;
    if (r == Q_NOTFOUND) return true;
    bool full = false;
    for (int i = 0; i < nF; ++i) full |= (int)(m.ld(m.iv(0, i) + 2) & 0xffffu) == pairLen;       // :318-342,493-517
    for (int i = 0; i < nR; ++i) full |= (int)(m.ld(m.iv(1, i) + 2) & 0xffffu) == pairLen;
    int hF, hR;
    const int h = q_hits2<VOTE>(ix, m, nF, nR, hF, hR);
    if (h < 0) { why = h; return false; }
    if (nF + nR > 0) flags |= MATE_SEEDED;                                                       // :311,486
    if (full) flags |= MATE_FULL;
    nh = (uint32_t)h;
    return true;
}
template <bool VOTE>
CG_HD bool q_wt_mate(const IndexView& ix, const QMemT<VOTE>& m, int L, int pairLen, int lceMode, uint8_t& flags, uint32_t& nh, int& why, int& nF, int& nR, int known = -1) {
    const int r = q_collect<VOTE>(ix, m, L, lceMode, nF, nR, known);
    why = r;
    if (r < 0) return false;
    return q_wt_finish<VOTE>(ix, m, r, pairLen, nF, nR, flags, nh, why);
}

// one record of Circall_bsj (src/Circall_bsj.cpp:300-399): false = declined (nothing was emitted).  On success the
// hit transcripts are left in ascending order at tids_at (word index in the scratch), nh of them.
template <bool VOTE, class Sink>
CG_HD bool q_bsj_finish(const IndexView& ix, const QMemT<VOTE>& m, int r, int nF, int nR, Sink& sink, uint32_t& nh, bool& split, int& tids_at, int& why);
template <bool VOTE, class Sink>
CG_HD bool q_bsj_record(const IndexView& ix, const QMemT<VOTE>& m, int L, int lceMode, Sink& sink, uint32_t& nh, bool& split, int& tids_at, int& why, int& nF, int& nR, int known = -1) {
    const int r = q_collect<VOTE>(ix, m, L, lceMode, nF, nR, known);
    why = r;
    tids_at = m.cb(0);
    if (r < 0) return false;
    return q_bsj_finish<VOTE>(ix, m, r, nF, nR, sink, nh, split, tids_at, why);
}
// the record after the collector returned r (>= 0) with nF / nR intervals in the scratch
template <bool VOTE, class Sink>
CG_HD bool q_bsj_finish(const IndexView& ix, const QMemT<VOTE>& m, int r, int nF, int nR, Sink& sink, uint32_t& nh, bool& split, int& tids_at, int& why) {
    tids_at = m.cb(0);
    nh = 0; split = false;
    if (r == Q_NOTFOUND) return true;
    int hF, hR;
    const int h = q_hits2<VOTE>(ix, m, nF, nR, hF, hR);
    if (h < 0) { why = h; return false; }
    if (h == 0) return true;                                                                     // :319
    if (h > Q_WH) { why = QD_WIDE_HITS; return false; }
    for (int s = 0; s < 2; ++s)
        for (int i = 0; i < (s == 0 ? nF : nR); ++i)
            if (m.ld(m.iv(s, i) + 1) - m.ld(m.iv(s, i)) > (uint32_t)Q_SPAN) { why = QD_WIDE_SPAN; return false; }
    nh = (uint32_t)h;
    // span filter over every SA position of every interval, forward list then reverse-complement list (:323-374)
    for (int s = 0; s < 2; ++s) {
        for (int i = 0; i < (s == 0 ? nF : nR); ++i) {
            const uint32_t ib = m.ld(m.iv(s, i)), ie = m.ld(m.iv(s, i) + 1), lq = m.ld(m.iv(s, i) + 2);
            const int len = (int)(lq & 0xffffu);
            for (uint32_t i0 = ib; i0 < ie; i0 += 4) {                                           // :326 / :355
                const int nn = (int)cmin<uint32_t>(4u, ie - i0);
                CG_COST(7, 1);
                uint32_t tt[4], ll[4], o0[4], o1[4];
                sa_tids4(ix, i0, nn, tt, ll);
#pragma unroll
                for (int z = 0; z < 4; ++z) { o0[z] = o1[z] = 0; if (z < nn) { o0[z] = CG_LDG(&ix.offsets[tt[z]]); o1[z] = CG_LDG(&ix.offsets[tt[z] + 1]); } }
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    if (z < nn) {
                        const uint32_t refLen = o1[z] - o0[z] - 1;
                        const int hitPos = (int)ll[z];
                        const int junctPos = (int)(refLen / 2);                                  // :333
                        if (hitPos < junctPos - 5 && hitPos + len > junctPos + 5) {              // :336 / :364
                            sink.line((uint32_t)s, lq >> 16, (uint32_t)len, hitPos, tt[z]);
                            split = true;
                        }
                    }
                }
            }
        }
    }
    // class key: the hit transcripts of both strands merged, in ascending order
    int j;
    if (hF == 0) { tids_at = m.cb(VOTE ? 1 : 0); j = hR; }
    else {
        j = hF;
        if (VOTE)
            for (int b = 0; b < hR; ++b) {
                const uint32_t t = m.ld(m.cb(1) + b);
                bool dup = false;
                for (int a = 0; a < hF; ++a) dup |= m.ld(m.cb(0) + a) == t;
                if (!dup) { m.st(m.cb(0) + j, t); ++j; }                             // j <= h <= Q_WH: stays inside the forward area
            }
    }
    for (int i = 1; i < j; ++i) {
        const uint32_t v = m.ld(tids_at + i);
        int y = i - 1;
        while (y >= 0 && m.ld(tids_at + y) > v) { m.st(tids_at + y + 1, m.ld(tids_at + y)); --y; }
        m.st(tids_at + y + 1, v);
    }
    return true;
}

} // namespace cgq
