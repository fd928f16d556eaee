// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [U] RapMap SASearcher::extendSearchNaive and SASearcher::lce.
// Call sites: SACollectorCircall.hpp:143-144,280-282,393-395 (extend) and :220,304,416 (lce).
// Kept in its own translation unit so it can be swapped if a real binary ever disagrees.
#include "oracle_types.hpp"
#include "u_mer.hpp"
#include <algorithm>
#include <cctype>

namespace orc {

char Query::at(int i) const {
    if (!rc) return (char)::toupper((unsigned char)read[start + i]);
    // reverse iterators + complementBases=true (SACollectorCircall.hpp:393-395)
    return mer_complement((char)::toupper((unsigned char)read[L - 1 - (start + i)]));
}

// Three binary searches, as upstream: (1) maximum-mappable-prefix length inside the open
// interval (lbIn, ubIn); (2) lower bound and (3) upper bound of query[0:maxI] (upstream uses
// the sentinels '#' < '$' and '}' > 'T' for these).  Characters compare by raw char value:
// '$'(36) < 'A' < 'C' < 'G' < 'N' < 'T'.
void extendSearchNaive(const Searcher& ss, OffsetT lbIn, OffsetT ubIn, OffsetT startAt,
                       const Query& q, OffsetT& lbOut, OffsetT& ubOut, OffsetT& lenOut) {
    const Index& ix = *ss.ix;
    const std::vector<OffsetT>& SA = ix.SA;
    const std::string& seq = ix.seq;
    const int64_t m = q.remaining();
    const int64_t n = (int64_t)seq.size();

    if (ubIn - lbIn == 2) {   // single candidate: linear compare
        OffsetT c = lbIn + 1;
        int64_t i = startAt;
        ss.wc.S += 1;
        while (i < m && SA[c] + i < n) {
            ss.wc.T += 1;
            if (q.at((int)i) != seq[SA[c] + i]) break;
            ++i;
        }
        lbOut = c; ubOut = ubIn; lenOut = (OffsetT)i;
        return;
    }

    OffsetT lb = lbIn, ub = ubIn;
    int64_t prevILow = startAt, prevIHigh = startAt, maxI = startAt;
    while (ub - lb > 1) {
        OffsetT c = lb + (ub - lb) / 2;
        int64_t i = std::min(prevILow, prevIHigh);
        bool plt = true;   // query sorts at-or-before suffix c
        ss.wc.S += 1;
        bool decided = false;
        while (i < m && SA[c] + i < n) {
            char qc = q.at((int)i);
            char tc = seq[SA[c] + i];
            ss.wc.T += 1;
            if (qc < tc) { plt = true; decided = true; break; }
            if (qc > tc) { plt = false; decided = true; break; }
            ++i;
        }
        if (!decided) plt = true;   // query exhausted (or text end): query <= suffix
        if (i > maxI) maxI = i;
        if (plt) { ub = c; prevIHigh = i; } else { lb = c; prevILow = i; }
    }

    // prefix comparison of suffix c against query[0:maxI] (first startAt chars known equal)
    auto cmp_prefix = [&](OffsetT c) -> int {   // <0 suffix smaller, 0 shares prefix, >0 suffix larger
        ss.wc.S += 1;
        for (int64_t i = startAt; i < maxI; ++i) {
            if (SA[c] + i >= n) return -1;
            char qc = q.at((int)i);
            char tc = seq[SA[c] + i];
            ss.wc.T += 1;
            if (tc < qc) return -1;
            if (tc > qc) return 1;
        }
        return 0;
    };
    OffsetT lo = lbIn, hi = ubIn;     // lower bound: first c in (lbIn, ubIn) with cmp >= 0
    while (hi - lo > 1) {
        OffsetT c = lo + (hi - lo) / 2;
        if (cmp_prefix(c) >= 0) hi = c; else lo = c;
    }
    OffsetT lower = hi;
    lo = lower - 1; hi = ubIn;        // upper bound: first c with cmp > 0
    while (hi - lo > 1) {
        OffsetT c = lo + (hi - lo) / 2;
        if (cmp_prefix(c) > 0) hi = c; else lo = c;
    }
    lbOut = lower; ubOut = hi; lenOut = (OffsetT)maxI;
}

// Semantic definition by brute force, for property tests.
void extendSearchBrute(const Searcher& ss, OffsetT lbIn, OffsetT ubIn, OffsetT startAt,
                       const Query& q, OffsetT& lbOut, OffsetT& ubOut, OffsetT& lenOut) {
    const Index& ix = *ss.ix;
    const int64_t m = q.remaining();
    const int64_t n = (int64_t)ix.seq.size();
    int64_t best = -1;
    OffsetT first = ubIn, last = lbIn;
    for (OffsetT c = lbIn + 1; c < ubIn; ++c) {
        int64_t i = 0;
        while (i < m && ix.SA[c] + i < n && q.at((int)i) == ix.seq[ix.SA[c] + i]) ++i;
        if (i > best) { best = i; first = c; last = c; }
        else if (i == best) last = c;
    }
    if (best < startAt) best = startAt;
    lbOut = first; ubOut = last + 1; lenOut = (OffsetT)best;
}

// SURVEY B1.  LCE_LITERAL: startAt is applied twice (o1/o2 already include it and len starts
// at it); the stopAt test happens only after an equal, non-'$' character.
int lce(const Searcher& ss, OffsetT p1, OffsetT p2, OffsetT startAt, OffsetT stopAt) {
    const Index& ix = *ss.ix;
    const std::string& seq = ix.seq;
    const int64_t textLen = (int64_t)seq.size();
    int64_t len = startAt;
    int64_t o1 = (int64_t)ix.SA[p1] + (ss.lceMode == 0 ? startAt : 0);
    int64_t o2 = (int64_t)ix.SA[p2] + (ss.lceMode == 0 ? startAt : 0);
    ss.wc.S += 2;
    int64_t maxIndex = std::max(o1, o2);
    while (maxIndex + len < textLen && seq[o1 + len] == seq[o2 + len]) {
        ss.wc.T += 2;
        if (seq[o1 + len] == '$') break;
        if (len >= stopAt) break;
        ++len;
    }
    return (int)len;
}

} // namespace orc
