// circall_b200 — thread-per-read state machine ("TSM"): the first tier of k_wt / k_bsj.
//
// One THREAD maps one read.  The per-read algorithm — include/SACollectorCircall.hpp:24-582 as restated by
// cg::collect(), then the hit enumeration of :492-579 and, for Circall_bsj, the span filter of
// src/Circall_bsj.cpp:323-374 — is split into
//   * CONTROL: the reference's own sequential logic, written as a resumable routine (a protothread:
//     `switch (pc)` with resume labels inside its loops).  It never touches global memory; whenever it needs the
//     index it files a request with one of four ENGINES and suspends.
//   * ENGINES (table look-ups, extendSearchNaive, lce, SA position -> transcript): each is a short chain of
//     "issue loads into the context's registers" / "consume them one scheduling round later" steps.
// One call of tsm_step() = complete what was in flight, run the control code up to its next request, issue
// that request's loads.  The kernel's main loop calls it for the 32 lanes of a warp over and over.  All lanes
// go through the engine code together (it is the heavy part: hashing, 2-bit comparisons) whatever control
// state they are in, the control code in between is a few instructions, and every lane keeps its own loads in
// flight across the call — 32+ independent HBM requests per warp instead of the one or two of the
// warp-per-read kernels, whose 32 lanes mostly repeat scalar work.
//
// Scope: the common paths only.  Anything unusual makes the routine BAIL — nothing is emitted and the read is
// redone from scratch by the warp-per-read kernels, whose results are identical:
//   * an N in the read, a homopolymer k-mer window
//   * both strands holding hits at the vote (the kmerScores vote, :442-490, stays with the warp kernels)
//   * more than T_FI intervals on a strand, an open interval wider than T_WMAX in the extension,
//     more than T_HC hit candidates, an interval wider than T_OTHER in the intersection / the span filter
//
// Exact shortcut used by the hit enumeration (this tier and only this tier): an interval whose query range
// [qpos, qpos + len) lies inside the query range of a longer interval of the same strand cannot change the
// set of hit transcripts — every occurrence p of the longer match gives the occurrence p + (qpos' - qpos) of
// the shorter one in the same transcript, so tids(longer) is a subset of tids(shorter) and, the map being
// injective, span(longer) <= span(shorter).  intersectSAHits computes the intersection of the tid sets of all
// intervals, which therefore equals the intersection over the maximal ones; the others are never enumerated.
//
// CG_HD throughout: tests/emu steps the same code on the host against the oracle.
#pragma once
#include "cg_pipeline.cuh"

namespace cgt {
using namespace cg;

static const int T_FI = 8;        // intervals kept per strand
static const int T_HC = 16;       // hit candidates (positions of the narrowest maximal interval)
static const int T_WMAX = 32;     // widest open interval the extension resolves candidate by candidate
static const int T_OTHER = 64;    // widest interval walked by the intersection / the span filter

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
//   [4 W2, +6 T_FI)  intervals: forward i at 3 i, reverse-complement i at 3 (T_FI + i): begin, end, len | qpos << 16
//   then T_HC        hit candidates (transcript ids)
struct TMem {
    uint32_t* base; int stride; int W2;
    static const bool kAsync = false;
    CG_HD uint32_t ld(int j) const { return base[(size_t)j * stride]; }
    CG_HD void st(int j, uint32_t v) const { base[(size_t)j * stride] = v; }
    CG_HD int iv(int s, int i) const { return 4 * W2 + 3 * (s * T_FI + i); }
    CG_HD int cb() const { return 4 * W2 + 6 * T_FI; }
    CG_HD static int words(int W2) { return 4 * W2 + 6 * T_FI + T_HC; }
    // loads of the engines: value at p -> landing field (here: straight into the context's registers)
    CG_HD void ld16(u32x4& dst, int /*slot*/, const u32x4* p) const { dst = CG_LDG(p); }
    CG_HD void ld8(uint32_t& lo, uint32_t& hi, int /*slot*/, int /*half*/, const u32x2* p) const { u32x2 v = CG_LDG(p); lo = v.x; hi = v.y; }
    CG_HD void ld4(uint32_t& dst, int /*slot*/, const uint32_t* p) const { dst = CG_LDG(p); }
    CG_HD void ld4w(uint32_t& dst, int /*slot*/, int /*word*/, const uint32_t* p) const { dst = CG_LDG(p); }
    // 32-base window of strand s at oriented position i (i < L)
    CG_HD uint64_t chunk(int s, int i) const {
        int p = (s ? 2 * W2 : 0) + (i >> 4), sh = (i & 15) * 2;
        uint32_t a = ld(p), b = ld(p + 1), c = ld(p + 2);
        return (uint64_t)fsr(a, b, sh) | ((uint64_t)fsr(b, c, sh) << 32);
    }
};

enum { BAIL_N = 1, BAIL_HOMO, BAIL_VOTE, BAIL_NINT, BAIL_WIDE_EXT, BAIL_NCAND, BAIL_WIDE_OTHER, BAIL_NUM };

// control resume points
enum {
    PC_START = 0, PC_DONE = -1, PC_BAIL = -2,
    PC_A_PROBED = 1, PC_A_EXT, PC_B_LCE0, PC_B_PROBED, PC_B_RUN, PC_B_RC, PC_B_EXT, PC_B_LCE, PC_H_TID, PC_O_TID, PC_J_TID
};
// engine states: RQ_x = requested by the control code (loads not issued yet), RQ_x_y = loads y in flight
enum {
    RQ_NONE = 0,
    RQ_PROBE, RQ_PROBE_W,                   // table look-ups
    RQ_EXT, RQ_EXT_SA, RQ_EXT_BLK,          // extendSearchNaive, candidate by candidate
    RQ_LCE, RQ_LCE_SA, RQ_LCE_BLK,          // SASearcher::lce
    RQ_TID, RQ_TID_SA, RQ_TID_BLK, RQ_TID_OFF,  // up to four SA positions -> transcript ids (-> transcript bounds)
    RQ_FETCH_R, RQ_FETCH_W, RQ_DEAD             // owned by the kernel: record index / packed read in flight; out of reads
};
// look-up patterns
enum { PM_PAIR1 = 0,    // k-mer at pos and its reverse complement
       PM_PAIR2,        // the same for pos and pos + 1
       PM_RUN,          // the k-mers at pos .. pos + 3 (those that fit the read)
       PM_RC1 };        // the reverse complement of the k-mer at pos

// Load targets.  Every global load of the engines lands in one of these; nothing else in the context is written by a
// load, and the engines never keep a value here across rounds except through a load.  On the host (and in a
// thread-resident kernel) they are plain fields; in the block-sorted kernel they are the cp.async landing zone of the
// context in shared memory, re-read at the start of every round and never written back.
struct Land {
    u32x4 s0, s1, s2, s3;                  // table slots | code halves of text blocks A / B (s0 / s1) and their '$' masks
                                           // (s2: x,y = A; z,w = B) | second halves of four text blocks | transcript bounds
    uint32_t sa0, sa1, sa2, sa3;           // suffix-array entries
};
enum { LD_S0 = 0, LD_S1, LD_S2, LD_S3 };   // 16-byte landing slots; LD_S2 also takes two 8-byte halves (ld8 lo / hi)

struct Ctx {
    int pc, req;
    // the read
    int L, pairLen;
    // collector (names as in cg::collect / SACollectorCircall.hpp)
    int pos, s, nF, nR, matched, q0, prevO;
    uint32_t fwdHit, rcHit, lbF, ubF, pB, pE, ivB, ivE;
    bool found, lastSearch, expectAbsent, pOther, other, hasF, hasR, a_two;
    // look-up engine: request (strand p_s, position p_pos, pattern p_mode, limit p_lim = last position allowed)
    int p_s, p_pos, p_mode, p_lim;
    uint64_t key0, key1, key2, key3;
    uint32_t h0, h1, h2, h3, pend, fnd;
    // extension engine
    int x_s, x_q, x_i, x_best, x_m, x_len;
    uint32_t x_lbIn, x_ubIn, x_ci, x_w, x_first, x_last, x_p, x_lb, x_ub, x_b0, x_b1;
    bool x_two;
    // lce engine
    uint32_t l_i1, l_i2, l_o1, l_o2; int l_len, l_lim, l_res;
    // SA position -> transcript engine: positions e_c .. min(e_c + 4, e_end); out: d.sa0..3 (text positions), tid0..3
    uint32_t e_c, e_end, tid0, tid1, tid2, tid3;
    bool e_off;                            // also fetch offsets[tid], offsets[tid + 1] (into d.s0 / d.s1)
    // hit enumeration
    uint32_t keep, alive, seen, e_n;
    int e_x, e_mi, nc;
    // results
    uint32_t nh; uint8_t flags; bool split;
    int why;                               // BAIL_* when pc == PC_BAIL
    int zero;                              // always 0, but only known at run time (see T_REQ)
    uint32_t x, r;                         // owned by the kernel: work item (mate / export record) and the mate it maps
    Land d;                                // last: everything before it is the part of the context that is saved
};

CG_HD uint64_t slot_key(const u32x4& s) { return (uint64_t)s.x | ((uint64_t)s.y << 32); }
CG_HD uint32_t tid_from(const u32x4& t, uint32_t p) {
    uint64_t mask = (uint64_t)t.x | ((uint64_t)t.y << 32);
    uint32_t o = p & 63;
    return t.z + (uint32_t)popc64(o ? (mask & ((1ULL << o) - 1)) : 0ULL);
}
CG_HD const u32x2* mask_ptr(const IndexView& ix, uint32_t b) { return reinterpret_cast<const u32x2*>(&ix.text[2 * (uint64_t)b + 1]); }

// ---- engine pieces -------------------------------------------------------------------------------------------
// text blocks for the extension: the block that holds text position p + i and, when the compared range reaches
// into it, the next one
template <class M>
CG_HD void ext_issue_blocks(Ctx& c, const IndexView& ix, const M& m) {
    uint32_t tp = c.x_p + (uint32_t)c.x_i;
    c.x_b0 = tp >> 6;
    c.x_two = (int)(tp & 63) + (c.x_m - c.x_i) > 64;
    m.ld16(c.d.s0, LD_S0, &ix.text[2 * (uint64_t)c.x_b0]);
    m.ld8(c.d.s2.x, c.d.s2.y, LD_S2, 0, mask_ptr(ix, c.x_b0));
    if (c.x_two) {
        m.ld16(c.d.s1, LD_S1, &ix.text[2 * (uint64_t)c.x_b0 + 2]);
        m.ld8(c.d.s2.z, c.d.s2.w, LD_S2, 1, mask_ptr(ix, c.x_b0 + 1));
    }
}
CG_HD uint64_t half_codes(const u32x4& b, int half) { return half ? ((uint64_t)b.z | ((uint64_t)b.w << 32)) : ((uint64_t)b.x | ((uint64_t)b.y << 32)); }
// next candidate of the extension: its SA entry has landed in sa0; fetch the following one alongside its text blocks
template <class M>
CG_HD void ext_next_candidate(Ctx& c, const IndexView& ix, const M& m, int k) {
    c.x_p = c.d.sa0;
    if (c.x_ci + 1 < c.x_w) m.ld4(c.d.sa0, 0, &ix.sa[c.x_lbIn + 2 + c.x_ci]);
    c.x_i = k;
    ext_issue_blocks(c, ix, m);
}
template <class M>
CG_HD void lce_issue_blocks(Ctx& c, const IndexView& ix, const M& m) {
    c.x_b0 = (c.l_o1 + (uint32_t)c.l_len) >> 6; c.x_b1 = (c.l_o2 + (uint32_t)c.l_len) >> 6;
    m.ld16(c.d.s0, LD_S0, &ix.text[2 * (uint64_t)c.x_b0]);
    m.ld8(c.d.s2.x, c.d.s2.y, LD_S2, 0, mask_ptr(ix, c.x_b0));
    if (c.x_b1 != c.x_b0) {
        m.ld16(c.d.s1, LD_S1, &ix.text[2 * (uint64_t)c.x_b1]);
        m.ld8(c.d.s2.z, c.d.s2.w, LD_S2, 1, mask_ptr(ix, c.x_b1));
    }
}

// The request is filed as `constant ^ run-time zero` so that the three phases stay separate pieces of code whatever
// the compiler inlines: the kernel puts a full-warp meeting point between them (see k_tsm), because no reconvergence
// barrier is placed around a switch that jumps into loops, and without one every control state would run the heavy
// engine code by itself.
#define T_REQ(rq, label) do { c.req = (rq) ^ c.zero; c.pc = (label); return; case (label):; } while (0)
#define T_BAIL(why_) do { c.pc = PC_BAIL; c.why = (why_); c.req = c.zero; return; } while (0)
#define T_DONE() do { c.pc = PC_DONE; c.req = c.zero; return; } while (0)
// extendSearchNaive(x_lbIn, x_ubIn) of strand x_s at query position x_q, startAt = k: out x_lb, x_ub, x_len
#define T_EXT(label) do { \
        c.x_m = L - c.x_q; c.x_w = c.x_ubIn - c.x_lbIn - 1; \
        if (c.x_m <= k) { c.x_lb = c.x_lbIn + 1; c.x_ub = c.x_ubIn; c.x_len = k; } \
        else if (c.x_w == 0) { c.x_lb = c.x_ubIn; c.x_ub = c.x_ubIn; c.x_len = k; }      /* empty open interval */ \
        else { if (c.x_w > (uint32_t)T_WMAX) T_BAIL(BAIL_WIDE_EXT); T_REQ(RQ_EXT, label); } \
    } while (0)
// lce(SA index l_o1, SA index l_o2, startAt l_len, stopAt l_lim): out l_res
#define T_LCE(label) do { if (c.l_lim <= c.l_len) c.l_res = c.l_len; else T_REQ(RQ_LCE, label); } while (0)

// One scheduling round of a read = tsm_complete() (consume the loads that were in flight; true when the control code
// may run), tsm_control() (the reference's logic up to its next request) and tsm_issue() (that request's loads).
// Sink::line(dir, queryPos, len, hitPos, tid) receives the junction-table rows (BSJ only).  On PC_DONE: c.flags / c.nh
// (wt), c.nh / c.split and, for BSJ, the hit transcripts in ascending order in the candidate words of `m`.
template <class M>
CG_HD bool tsm_complete(Ctx& c, const IndexView& ix, const M& m, int lceMode) {
    const int k = ix.k;
    // ================= complete what was in flight =================
    bool ctl = true;                                    // false: this lane's engine needs another round
    if (c.req == RQ_PROBE_W) {
#define T_RESOLVE(bit, KEY, S, SLOT, H) if (c.pend & (bit)) { uint64_t kk = slot_key(S); if (kk == (KEY)) { c.fnd |= (bit); c.pend &= ~(bit); } else if (kk == ~0ULL) c.pend &= ~(bit); else { H = (uint32_t)(((uint64_t)H + 1) & ix.hmask); m.ld16(S, SLOT, &ix.hslots[H]); } }
        T_RESOLVE(1u, c.key0, c.d.s0, LD_S0, c.h0)
        T_RESOLVE(2u, c.key1, c.d.s1, LD_S1, c.h1)
        T_RESOLVE(4u, c.key2, c.d.s2, LD_S2, c.h2)
        T_RESOLVE(8u, c.key3, c.d.s3, LD_S3, c.h3)
#undef T_RESOLVE
        if (c.pend) ctl = false;                        // linear probing: the next slots are in flight
        else c.req = RQ_NONE;
    } else if (c.req == RQ_EXT_SA) {
        ext_next_candidate(c, ix, m, k);
        c.req = RQ_EXT_BLK;
        ctl = false;
    } else if (c.req == RQ_EXT_BLK) {
        bool stop = false;
        while (c.x_i < c.x_m) {
            uint32_t tp = c.x_p + (uint32_t)c.x_i;
            uint32_t bi = (tp >> 6) - c.x_b0;
            if (bi > (c.x_two ? 1u : 0u)) break;                             // past the loaded blocks
            int half = (tp >> 5) & 1, o5 = tp & 31;
            uint64_t tw = half_codes(bi ? c.d.s1 : c.d.s0, half) >> (2 * o5);
            uint32_t tm = (bi ? (half ? c.d.s2.w : c.d.s2.z) : (half ? c.d.s2.y : c.d.s2.x)) >> o5;
            int n = cmin(32 - o5, c.x_m - c.x_i);
            uint64_t q = m.chunk(c.x_s, c.x_q + c.x_i);
            uint64_t x = tw ^ q;
            uint64_t d = (x | (x >> 1)) & EVEN;
            int j1 = d ? (ctz64(d) >> 1) : 32;
            int j2 = tm ? ctz32(tm) : 32;
            int j = cmin(j1, j2);
            if (j < n) { c.x_i += j; stop = true; break; }
            c.x_i += n;
        }
        if (!stop && c.x_i < c.x_m) { ext_issue_blocks(c, ix, m); ctl = false; }     // the match runs on into further blocks
        else {
            // the suffixes sharing the longest prefix are contiguous in SA order
            if (c.x_i > c.x_best) { c.x_best = c.x_i; c.x_first = c.x_ci; c.x_last = c.x_ci; }
            else if (c.x_i == c.x_best) c.x_last = c.x_ci;
            if (++c.x_ci < c.x_w) { ext_next_candidate(c, ix, m, k); ctl = false; }
            else {
                c.x_lb = c.x_lbIn + 1 + c.x_first; c.x_ub = c.x_lbIn + 1 + c.x_last + 1; c.x_len = c.x_best;
                c.req = RQ_NONE;
            }
        }
    } else if (c.req == RQ_LCE_SA) {
        c.l_o1 = c.d.sa0; c.l_o2 = (c.l_i2 == c.l_i1) ? c.d.sa0 : c.d.sa1;
        if (lceMode == 0) { c.l_o1 += (uint32_t)c.l_len; c.l_o2 += (uint32_t)c.l_len; }
        uint32_t mx = cmax(c.l_o1, c.l_o2);
        int64_t room = (int64_t)ix.n - (int64_t)mx;                          // maxIndex + len < textLen
        c.l_lim = (int)cmin<int64_t>((int64_t)c.l_lim, room);
        if (c.l_len < c.l_lim) { lce_issue_blocks(c, ix, m); c.req = RQ_LCE_BLK; ctl = false; }
        else { c.l_res = c.l_len; c.req = RQ_NONE; }
    } else if (c.req == RQ_LCE_BLK) {
        const bool same = c.x_b1 == c.x_b0;
        bool stop = false;
        while (c.l_len < c.l_lim) {
            uint32_t a = c.l_o1 + (uint32_t)c.l_len, b = c.l_o2 + (uint32_t)c.l_len;
            if ((a >> 6) != c.x_b0 || (b >> 6) != c.x_b1) break;
            int ha = (a >> 5) & 1, oa = a & 31, hb = (b >> 5) & 1, ob = b & 31;
            uint64_t wa = half_codes(c.d.s0, ha) >> (2 * oa);
            uint64_t wb = half_codes(same ? c.d.s0 : c.d.s1, hb) >> (2 * ob);
            uint32_t ma = (ha ? c.d.s2.y : c.d.s2.x) >> oa;
            uint32_t mb = (same ? (hb ? c.d.s2.y : c.d.s2.x) : (hb ? c.d.s2.w : c.d.s2.z)) >> ob;
            int n = cmin(cmin(32 - oa, 32 - ob), c.l_lim - c.l_len);
            uint64_t x = wa ^ wb;
            uint64_t d = (x | (x >> 1)) & EVEN;
            int j1 = d ? (ctz64(d) >> 1) : 32;
            uint32_t mm = ma | mb;
            int j2 = mm ? ctz32(mm) : 32;
            int j = cmin(j1, j2);
            if (j < n) { c.l_len += j; stop = true; break; }
            c.l_len += n;
        }
        if (!stop && c.l_len < c.l_lim) { lce_issue_blocks(c, ix, m); ctl = false; }
        else { c.l_res = c.l_len; c.req = RQ_NONE; }
    } else if (c.req == RQ_TID_SA) {
        m.ld16(c.d.s0, LD_S0, &ix.text[2 * (uint64_t)(c.d.sa0 >> 6) + 1]);
        if (c.e_c + 1 < c.e_end) m.ld16(c.d.s1, LD_S1, &ix.text[2 * (uint64_t)(c.d.sa1 >> 6) + 1]);
        if (c.e_c + 2 < c.e_end) m.ld16(c.d.s2, LD_S2, &ix.text[2 * (uint64_t)(c.d.sa2 >> 6) + 1]);
        if (c.e_c + 3 < c.e_end) m.ld16(c.d.s3, LD_S3, &ix.text[2 * (uint64_t)(c.d.sa3 >> 6) + 1]);
        c.req = RQ_TID_BLK;
        ctl = false;
    } else if (c.req == RQ_TID_BLK) {
        c.tid0 = tid_from(c.d.s0, c.d.sa0);
        if (c.e_c + 1 < c.e_end) c.tid1 = tid_from(c.d.s1, c.d.sa1);
        if (c.e_c + 2 < c.e_end) c.tid2 = tid_from(c.d.s2, c.d.sa2);
        if (c.e_c + 3 < c.e_end) c.tid3 = tid_from(c.d.s3, c.d.sa3);
        if (c.e_off) {
            m.ld4w(c.d.s0.x, LD_S0, 0, &ix.offsets[c.tid0]); m.ld4w(c.d.s0.y, LD_S0, 1, &ix.offsets[c.tid0 + 1]);
            if (c.e_c + 1 < c.e_end) { m.ld4w(c.d.s0.z, LD_S0, 2, &ix.offsets[c.tid1]); m.ld4w(c.d.s0.w, LD_S0, 3, &ix.offsets[c.tid1 + 1]); }
            if (c.e_c + 2 < c.e_end) { m.ld4w(c.d.s1.x, LD_S1, 0, &ix.offsets[c.tid2]); m.ld4w(c.d.s1.y, LD_S1, 1, &ix.offsets[c.tid2 + 1]); }
            if (c.e_c + 3 < c.e_end) { m.ld4w(c.d.s1.z, LD_S1, 2, &ix.offsets[c.tid3]); m.ld4w(c.d.s1.w, LD_S1, 3, &ix.offsets[c.tid3 + 1]); }
            c.req = RQ_TID_OFF;
            ctl = false;
        } else c.req = RQ_NONE;
    } else if (c.req == RQ_TID_OFF) {
        c.req = RQ_NONE;
    }

    return ctl;
}

template <bool BSJ, class M, class Sink>
CG_HD void tsm_control(Ctx& c, const IndexView& ix, const M& m, Sink& sink) {
    const int k = ix.k, L = c.L;
    const int skipOverlap = k - 1;
    const uint32_t maxInterval = 1000;
    switch (c.pc) {
    case PC_START:
#include "cg_tsm_control.inc"

    default:
        break;
    }
}

template <class M>
CG_HD void tsm_issue(Ctx& c, const IndexView& ix, const M& m) {
    const int k = ix.k;
    const uint64_t KM = kmask(k);
    if (c.req == RQ_PROBE) {
        // keys of the pattern, then their home slots
        const int p = c.p_pos;
        bool homo = false;
        c.pend = 0; c.fnd = 0;
        if (c.p_mode == PM_RUN) {
            if (p <= c.p_lim) { c.key0 = m.chunk(c.p_s, p) & KM; c.pend |= 1u; homo |= homopolymer(c.key0, k); }
            if (p + 1 <= c.p_lim) { c.key1 = m.chunk(c.p_s, p + 1) & KM; c.pend |= 2u; homo |= homopolymer(c.key1, k); }
            if (p + 2 <= c.p_lim) { c.key2 = m.chunk(c.p_s, p + 2) & KM; c.pend |= 4u; homo |= homopolymer(c.key2, k); }
            if (p + 3 <= c.p_lim) { c.key3 = m.chunk(c.p_s, p + 3) & KM; c.pend |= 8u; homo |= homopolymer(c.key3, k); }
        } else {
            c.key0 = m.chunk(c.p_s, p) & KM;
            homo |= homopolymer(c.key0, k);
            c.key1 = revcomp(c.key0, k);
            c.pend = 3u;
            if (c.p_mode == PM_RC1) { c.key0 = c.key1; c.pend = 1u; }
            else if (c.p_mode == PM_PAIR2 && p + 1 <= c.p_lim) {
                c.key2 = m.chunk(c.p_s, p + 1) & KM;
                homo |= homopolymer(c.key2, k);
                c.key3 = revcomp(c.key2, k);
                c.pend = 15u;
            }
        }
        if (homo) { c.pc = PC_BAIL; c.why = BAIL_HOMO; c.req = RQ_NONE; }      // (no jump back to `issue`: the warp meets there once per step)
        else {
            if (c.pend & 1u) { c.h0 = (uint32_t)home_slot(ix, c.key0); m.ld16(c.d.s0, LD_S0, &ix.hslots[c.h0]); }
            if (c.pend & 2u) { c.h1 = (uint32_t)home_slot(ix, c.key1); m.ld16(c.d.s1, LD_S1, &ix.hslots[c.h1]); }
            if (c.pend & 4u) { c.h2 = (uint32_t)home_slot(ix, c.key2); m.ld16(c.d.s2, LD_S2, &ix.hslots[c.h2]); }
            if (c.pend & 8u) { c.h3 = (uint32_t)home_slot(ix, c.key3); m.ld16(c.d.s3, LD_S3, &ix.hslots[c.h3]); }
            c.req = RQ_PROBE_W;
        }
    } else if (c.req == RQ_EXT) {
        c.x_best = -1; c.x_first = c.x_last = 0; c.x_ci = 0;
        m.ld4(c.d.sa0, 0, &ix.sa[c.x_lbIn + 1]);
        c.req = RQ_EXT_SA;
    } else if (c.req == RQ_LCE) {
        m.ld4(c.d.sa0, 0, &ix.sa[c.l_i1]);
        if (c.l_i2 != c.l_i1) m.ld4(c.d.sa1, 1, &ix.sa[c.l_i2]);
        c.req = RQ_LCE_SA;
    } else if (c.req == RQ_TID) {
        m.ld4(c.d.sa0, 0, &ix.sa[c.e_c]);
        if (c.e_c + 1 < c.e_end) m.ld4(c.d.sa1, 1, &ix.sa[c.e_c + 1]);
        if (c.e_c + 2 < c.e_end) m.ld4(c.d.sa2, 2, &ix.sa[c.e_c + 2]);
        if (c.e_c + 3 < c.e_end) m.ld4(c.d.sa3, 3, &ix.sa[c.e_c + 3]);
        c.req = RQ_TID_SA;
    }
}

// the three phases back to back (host emulation; the kernel interleaves them with warp meeting points)
template <bool BSJ, class M, class Sink>
CG_HD void tsm_step(Ctx& c, const IndexView& ix, const M& m, int lceMode, Sink& sink) {
    if (tsm_complete(c, ix, m, lceMode)) tsm_control<BSJ>(c, ix, m, sink);
    tsm_issue(c, ix, m);
}

#undef T_REQ
#undef T_EXT
#undef T_LCE
#undef T_BAIL
#undef T_DONE

// ---- the same control code as plain sequential code: one thread maps one read start to finish -----------------------
// An engine request runs to completion on the spot (issue the loads, consume them, repeat while the engine wants
// another round).  No switch, no resume labels: ordinary structured loops, which is what SIMT execution wants when a
// warp holds 32 reads that mostly do the same thing — and the compiler keeps only the live part of the context in
// registers.  Result in c.pc (PC_DONE / PC_BAIL), c.flags, c.nh, c.split, the candidate words of `m`.
template <class M>
CG_HD void tsm_engine_sync(Ctx& c, const IndexView& ix, const M& m, int lceMode) {
    tsm_issue(c, ix, m);
    while (c.req != RQ_NONE) tsm_complete(c, ix, m, lceMode);
}
#define T_REQ(rq, label) do { c.req = (rq); tsm_engine_sync(c, ix, m, lceMode); if (c.pc == PC_BAIL) return; } while (0)
#define T_BAIL(why_) do { c.pc = PC_BAIL; c.why = (why_); return; } while (0)
#define T_DONE() do { c.pc = PC_DONE; return; } while (0)
#define T_EXT(label) do { \
        c.x_m = L - c.x_q; c.x_w = c.x_ubIn - c.x_lbIn - 1; \
        if (c.x_m <= k) { c.x_lb = c.x_lbIn + 1; c.x_ub = c.x_ubIn; c.x_len = k; } \
        else if (c.x_w == 0) { c.x_lb = c.x_ubIn; c.x_ub = c.x_ubIn; c.x_len = k; } \
        else { if (c.x_w > (uint32_t)T_WMAX) T_BAIL(BAIL_WIDE_EXT); T_REQ(RQ_EXT, label); } \
    } while (0)
#define T_LCE(label) do { if (c.l_lim <= c.l_len) c.l_res = c.l_len; else T_REQ(RQ_LCE, label); } while (0)
template <bool BSJ, class M, class Sink>
CG_HD void tsm_mate_sync(Ctx& c, const IndexView& ix, const M& m, int lceMode, Sink& sink) {
    const int k = ix.k, L = c.L;
    const int skipOverlap = k - 1;
    const uint32_t maxInterval = 1000;
    c.pc = PC_START; c.req = RQ_NONE; c.zero = 0; c.why = 0;
    {
#include "cg_tsm_control.inc"
    }
}
#undef T_REQ
#undef T_EXT
#undef T_LCE
#undef T_BAIL
#undef T_DONE

struct NoSink { CG_HD void line(uint32_t, uint32_t, uint32_t, int, uint32_t) const {} };

// reverse-complement strand from the forward words: word j = 32 bases from oriented position 32 j (ReadView::chunk(1, .))
template <class M>
CG_HD void tsm_make_rc(const M& m, int L) {
    for (int j = 0; j < m.W2; ++j) {
        int a = L - 32 - 32 * j;
        uint64_t v = 0;
        if (a >= 0) v = rc64(m.chunk(0, a));
        else if (32 + a > 0) v = rc64(m.chunk(0, 0)) >> (2 * (-a));
        m.st(2 * m.W2 + 2 * j, (uint32_t)v); m.st(2 * m.W2 + 2 * j + 1, (uint32_t)(v >> 32));
    }
}

// Stage a packed read (cgk::k_pack record: W2f code words then the N-mask words) into the thread's scratch and
// initialise the context (host side; the kernel stages the words with cp.async and does the rest itself).
// Returns false when the read must go to the warp kernels straight away (an N below L).
CG_HD bool tsm_begin(Ctx& c, const TMem& m, const uint64_t* rec, int W2f, int WMf, int L, int pairLen) {
    bool anyN = false;
    for (int j = 0; j < WMf; ++j) {
        int lo = 64 * j;
        if (lo >= L) break;
        uint64_t mk = CG_LDG(rec + W2f + j);
        if (L - lo < 64) mk &= (1ULL << (L - lo)) - 1;
        anyN |= mk != 0;
    }
    c.L = L; c.pairLen = pairLen; c.pc = PC_START; c.req = RQ_NONE; c.why = 0; c.zero = 0;
    if (anyN || L < 1) { c.why = BAIL_N; return false; }
    for (int j = 0; j < W2f; ++j) { uint64_t v = CG_LDG(rec + j); m.st(2 * j, (uint32_t)v); m.st(2 * j + 1, (uint32_t)(v >> 32)); }
    m.st(2 * W2f, 0); m.st(2 * W2f + 1, 0);
    tsm_make_rc(m, L);
    return true;
}

} // namespace cgt
