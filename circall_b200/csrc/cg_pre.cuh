// circall_b200 — thread-per-mate front kernel of Circall_wt: mates that match a transcript over their whole length.
//
// With 0.5 % substitution errors 60 % of 100-base mates are such clean matches.  For them
// SACollectorCircall::operator() (include/SACollectorCircall.hpp:24-582) always takes the same path, and what
// Circall_wt needs from it (src/Circall_wt.cpp:311-342: seeded, full-length, hits.size()) follows from four table
// look-ups that can all be issued at once, one maximal-mappable-prefix extension and one transcript-id read per
// surviving suffix.  One THREAD resolves one mate: no divergence to speak of, four independent random sectors per
// thread in flight, a few dozen warp instructions per mate instead of the two thousand of the warp-per-read kernel,
// which then only sees the mates this one declines.
//
// Which mates, exactly (k = 31, L = read length, read without N, no homopolymer k-mer at either end):
//   forward strand:  Phase A (:112-200) finds the k-mer at 0 and not its reverse complement; the extension (:143)
//     reaches L; lce (:220) returns at once because nothing remains; Phase B (:208-329) jumps to the forced last
//     position L - k, whose k-mer must be there (it is part of the match) and whose reverse complement must not be;
//     Phase C (:332) does not run (rcHit = 0); the vote (:442-490) keeps the forward list.
//   reverse-complement strand: Phase A finds only the reverse complement; Phase C starts at 0 on the other strand,
//     whose first k-mer is the reverse complement of the read's last one: it must be there and the read's last k-mer
//     itself must not; the extension reaches L; the forced last position looks up the two k-mers of Phase A again.
//   Both ways the interval list is {(len L, q 0), (len k, q L - k) x 1 or 2}; the short ones lie inside the long one,
//   so the hits are the distinct transcripts of the long interval's positions (see cg_warp.cuh::w_hits_lanes).
// Anything else — a mismatch, both strands present, an open interval wider than PRE_WMAX — is declined: the thread
// returns false, nothing is written, and the warp kernels map the mate from scratch.
//
// CG_HD: tests/emu runs the same function on the host in front of the emulated tiers.
#pragma once
#include "cg_thrmem.cuh"

namespace cgp {
using namespace cg;

static const int PRE_WMAX = 8;       // widest k-mer interval whose suffixes one thread walks

// scratch words a thread needs: forward + reverse-complement code words in cgt::TMem's layout
CG_HD int pre_words(int W2) { return 4 * W2; }

// longest common prefix of strand s of the read (from offset i0) and the text at p, up to L
CG_HD int pre_match(const IndexView& ix, const cgt::TMem& m, int s, uint32_t p, int i0, int L) {
    int i = i0;
    uint32_t curb = 0xffffffffu;
    u32x4 cc; cc.x = cc.y = cc.z = cc.w = 0;
    cgt::u32x2 mk; mk.x = mk.y = 0;
    while (i < L) {
        const uint32_t tp = p + (uint32_t)i, b = tp >> 6;
        if (b != curb) { cc = CG_LDG(&ix.text[2 * (uint64_t)b]); mk = CG_LDG(cgt::mask_ptr(ix, b)); curb = b; }
        const int half = (tp >> 5) & 1, o5 = tp & 31;
        const uint64_t tw = cgt::half_codes(cc, half) >> (2 * o5);
        const uint32_t tm = (half ? mk.y : mk.x) >> o5;
        const int n = cmin(32 - o5, L - i);
        const uint64_t x = tw ^ m.chunk(s, i);
        const uint64_t d = (x | (x >> 1)) & EVEN;
        const int j = cmin(d ? (ctz64(d) >> 1) : 32, tm ? ctz32(tm) : 32);
        if (j < n) return i + j;
        i += n;
    }
    return i;
}

// m holds the forward strand; the reverse-complement strand is derived here when it is needed
// cls (1..16) classifies a declined mate by which of the four k-mers are in the table (1 + f0 + 2 f1 + 4 f2 + 8 f3; 1 when it
// never got to the look-ups): the lockstep tier sorts its work by it so that a warp holds reads that take the same path
// what the clean-match path implies for the collector's interval list (debug surface of the session, row A3): strand, the
// full-length interval, the k-mer interval of the forced last position
struct PreIv { int s; uint32_t lb, ub, bl, el; };
CG_HD bool perfect_mate(const IndexView& ix, const cgt::TMem& m, int L, int pairLen, uint8_t& flags, uint32_t& nh, uint8_t& cls, PreIv* iv = nullptr) {
    const int k = ix.k;
    cls = 1;
    if (L < k + 1 || L > 0xffff) return false;
    const uint64_t mer0 = m.chunk(0, 0) & ix.km, merL = m.chunk(0, L - k) & ix.km;
    if (homopolymer_ix(mer0, ix) || homopolymer_ix(merL, ix)) return false;
    const uint64_t rc0 = revcomp(mer0, k), rcL = revcomp(merL, k);
    // the four look-ups of the path, their home slots in flight together
    const uint64_t h0 = home_slot(ix, mer0), h1 = home_slot(ix, rc0), h2 = home_slot(ix, merL), h3 = home_slot(ix, rcL);
    const u32x4 s0 = CG_LDG(&ix.hslots[h0]), s1 = CG_LDG(&ix.hslots[h1]), s2 = CG_LDG(&ix.hslots[h2]), s3 = CG_LDG(&ix.hslots[h3]);
    uint32_t b0 = 0, e0 = 0, b1 = 0, e1 = 0, b2 = 0, e2 = 0, b3 = 0, e3 = 0;
    const bool f0 = resolve(ix, mer0, h0, s0, b0, e0), f1 = resolve(ix, rc0, h1, s1, b1, e1);
    const bool f2 = resolve(ix, merL, h2, s2, b2, e2), f3 = resolve(ix, rcL, h3, s3, b3, e3);
    cls = (uint8_t)(1 + (f0 ? 1 : 0) + (f1 ? 2 : 0) + (f2 ? 4 : 0) + (f3 ? 8 : 0));
    int s; uint32_t b, e, bl, el;                                       // (bl, el): the k-mer interval of the forced last position
    if (f0 && !f1 && f2 && !f3) { s = 0; b = b0; e = e0; bl = b2; el = e2; }      // :131-132 then :261+:268 at L - k
    else if (!f0 && f1 && !f2 && f3) { s = 1; b = b3; e = e3; bl = b1; el = e1; } // :131-132 then :372+:379 at 0 and at L - k
    else return false;
    const uint32_t w = e - b;
    if (b == 0 || w == 0 || w > (uint32_t)PRE_WMAX) return false;
    if (s == 1) cgt::make_rc(m, L);
    // extendSearchNaive(b - 1, e) (:143 / :393): every suffix of the interval
    uint32_t fullB = 0;                                                 // candidates that match over the whole read
    uint32_t tid[PRE_WMAX];
    if (ix.lcp && ix.satid) {
        // sorted suffixes: the match length of one follows from that of the one before and the lcp of the two (cg_thr.cuh::q_step);
        // the text is only read for the first suffix and where the two lengths are equal
        int hb[PRE_WMAX];
#pragma unroll
        for (int c = 0; c < PRE_WMAX; ++c) {
            hb[c] = (c > 0 && (uint32_t)c < w) ? (int)CG_LDG(&ix.lcp[b + c]) : 0;
            tid[c] = (uint32_t)c < w ? CG_LDG(&ix.satid[b + c]) : 0xffffffffu;
        }
        int lp = pre_match(ix, m, s, CG_LDG(&ix.sa[b]), k, L);
        if (lp == L) fullB |= 1u;
#pragma unroll
        for (int c = 1; c < PRE_WMAX; ++c) {
            if ((uint32_t)c < w) {
                const int h = hb[c];
                int l;
                if (h > lp) l = lp;
                else if (h < lp && h < 255) l = h;
                else l = pre_match(ix, m, s, CG_LDG(&ix.sa[b + c]), cmin(h, lp), L);
                if (l == L) fullB |= 1u << c;
                lp = l;
            }
        }
        if (!fullB) return false;                                       // the match stops early: not this kernel's mate
#pragma unroll
        for (int c = 0; c < PRE_WMAX; ++c) if (!((fullB >> c) & 1u)) tid[c] = 0xffffffffu;
    } else {
        uint32_t p[PRE_WMAX];                                           // all of their SA entries first
#pragma unroll
        for (int c = 0; c < PRE_WMAX; ++c) p[c] = (uint32_t)c < w ? CG_LDG(&ix.sa[b + c]) : 0u;
#pragma unroll
        for (int c = 0; c < PRE_WMAX; ++c)
            if ((uint32_t)c < w && pre_match(ix, m, s, p[c], k, L) == L) fullB |= 1u << c;
        if (!fullB) return false;                                       // the match stops early: not this kernel's mate
        // hits: distinct transcripts of the full-length interval (contiguous in SA order)
#pragma unroll
        for (int c = 0; c < PRE_WMAX; ++c) tid[c] = ((fullB >> c) & 1u) ? cgt::tid_from(CG_LDG(&ix.text[2 * (uint64_t)(p[c] >> 6) + 1]), p[c]) : 0xffffffffu;
    }
    uint32_t n = 0;
#pragma unroll
    for (int c = 0; c < PRE_WMAX; ++c) {
        bool fresh = (fullB >> c) & 1u;
#pragma unroll
        for (int y = 0; y < c; ++y) fresh = fresh && !(((fullB >> y) & 1u) && tid[y] == tid[c]);
        n += fresh;
    }
    // full-length test against mate 1's length (Circall_wt.cpp:242,318-342): the long interval, or — when mate 1 is
    // exactly k bases long — the (len k) interval of the last position, which is in the list when it is narrower than
    // maxInterval (:285 / :398; its lower bound is b - 1 + 1)
    const bool full = L == pairLen || (pairLen == k && bl > 0 && el > bl && el - bl < 1000u);
    flags = (uint8_t)(MATE_SEEDED | (full ? MATE_FULL : 0));            // Circall_wt.cpp:311
    nh = n;
    if (iv) {                                                           // the suffixes that share the longest prefix are contiguous
        iv->s = s; iv->lb = b + (uint32_t)ctz32(fullB); iv->ub = b + 32u - (uint32_t)clz32(fullB); iv->bl = bl; iv->el = el;
    }
    return true;
}

} // namespace cgp
