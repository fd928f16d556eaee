// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [R] restatement of SACollectorCircall::operator() — include/SACollectorCircall.hpp:24-582.
// Iterators are replaced by integer positions; every quirk listed in SURVEY.md §8 A3 is kept and
// the reference line it comes from is cited at the statement.
#include "oracle_types.hpp"
#include "u_mer.hpp"
#include <algorithm>
#include <string>

namespace orc {

void extendSearchNaive(const Searcher&, OffsetT, OffsetT, OffsetT, const Query&, OffsetT&, OffsetT&, OffsetT&);
int lce(const Searcher&, OffsetT, OffsetT, OffsetT, OffsetT);
struct SATxpQueryPos; struct ProcessedSAHit;
}
#include <map>
namespace orc {
struct SATxpQueryPos { uint32_t pos; uint32_t queryPos; bool queryRC; };
struct ProcessedSAHit { std::vector<SATxpQueryPos> tqvec; uint32_t numActive = 0; bool active = false; };
using SAHitMap = std::map<int, ProcessedSAHit>;
SAHitMap intersectSAHits(std::vector<SAIntervalHit>&, const Index&, WorkCounters&);
void collectHitsSimpleSA(SAHitMap&, uint32_t, uint32_t, std::vector<QuasiAlignment>&, MateStatus);

enum HitStatus { ABSENT = -1, UNTESTED = 0, PRESENT = 1 };           // :70
struct KmerDirScore {                                                 // :73-86
    uint64_t kmer; int32_t kpos; HitStatus fwdScore; HitStatus rcScore;
    bool operator==(const KmerDirScore& o) const { return kpos == o.kpos; }
    bool operator<(const KmerDirScore& o) const { return kpos < o.kpos; }
};

static const size_t NPOS = std::string::npos;

static inline size_t find_first_nN(const char* read, size_t L, size_t from) {   // read.find_first_of("nN", pos)
    for (size_t i = from; i < L; ++i)
        if (read[i] == 'n' || read[i] == 'N') return i;
    return NPOS;
}

bool collect(const Searcher& ss, const char* read, int readLenIn,
             std::vector<QuasiAlignment>& hits,
             std::vector<SAIntervalHit>& fwdSAInts, std::vector<SAIntervalHit>& rcSAInts,
             MateStatus mateStatus, bool strictCheck) {
    const Index& ix = *ss.ix;
    const KHash& khash = ix.khash;
    ss.wc.calls += 1;

    const int64_t readLen = readLenIn;
    const int64_t k = ix.k;
    const uint32_t maxDist = (uint32_t)(1.5 * readLen);               // :46
    auto probe = [&](uint64_t w) { ss.wc.P += 1; return khash.find(w); };

    int64_t rb = 0;                                                    // :54  (re = rb + k)
    OffsetT lbLeftFwd = 0, ubLeftFwd = 0, lbRightFwd = 0, ubRightFwd = 0, lbRightRC = 0, ubRightRC = 0;
    OffsetT matchedLen = 0;
    uint32_t fwdHit = 0, rcHit = 0;
    bool foundHit = false;
    uint64_t mer = 0, rcMer = 0;
    std::vector<KmerDirScore> kmerScores;
    const OffsetT maxInterval = 1000;                                  // :99
    const OffsetT skipOverlap = (OffsetT)(k - 1);                      // :106
    const OffsetT homoPolymerSkip = (OffsetT)(k / 2);                  // :108
    const uint32_t sampFactor = 1;                                     // :42

    // ---- Phase A (:112-200): first k-mer with a fwd or rc table hit; `re < read.end()` so the
    //      last k-mer is never a first seed
    while (rb + k < readLen) {
        size_t pos = (size_t)rb;
        size_t invalidPos = find_first_nN(read, (size_t)readLen, pos);
        if (invalidPos <= pos + (size_t)k) {                           // :118 (note <=, Phase B uses <)
            rb = (int64_t)invalidPos + 1;
            continue;
        }
        mer = mer_from_chars(read + pos, (int)k);                      // :126
        if (mer_is_homopolymer(mer, (int)k)) { rb += homoPolymerSkip; continue; }   // :127
        rcMer = mer_revcomp(mer, (int)k);
        auto merIt = probe(mer);                                       // :131
        auto rcMerIt = probe(rcMer);                                   // :132
        if (merIt) {
            OffsetT lb = merIt->begin, ub = merIt->end;
            OffsetT lbRestart = std::max((OffsetT)0, lb - 1);          // :140
            Query q{read, (int)readLen, (int)rb, false};
            extendSearchNaive(ss, lbRestart, ub, (OffsetT)k, q, lbLeftFwd, ubLeftFwd, matchedLen);   // :143
            OffsetT diff = ubLeftFwd - lbLeftFwd;
            if (ubLeftFwd > lbLeftFwd && diff < maxInterval) {         // :149
                fwdSAInts.push_back({lbLeftFwd, ubLeftFwd, matchedLen, (uint32_t)rb, false});
                if (strictCheck) {
                    ++fwdHit;
                    if (rcMerIt) { ++rcHit; kmerScores.push_back({mer, (int32_t)pos, PRESENT, PRESENT}); }
                    else { kmerScores.push_back({mer, (int32_t)pos, PRESENT, ABSENT}); }
                    if (rb + matchedLen < readLen) {                   // :165
                        int64_t kmerPos = rb + matchedLen - skipOverlap;
                        mer = mer_from_chars(read + kmerPos, (int)k);  // :167 (no N check upstream)
                        kmerScores.push_back({mer, (int32_t)kmerPos, ABSENT, UNTESTED});
                    }
                } else {
                    ++fwdHit;
                    if (rcMerIt) ++rcHit;
                }
            }
        }
        if (rcMerIt) {                                                 // :178
            if (rcMerIt->end > rcMerIt->begin) {
                if (!fwdHit) {
                    ++rcHit;
                    if (strictCheck) kmerScores.push_back({mer, (int32_t)pos, ABSENT, PRESENT});
                }
            }
        }
        if (fwdHit + rcHit > 0) { foundHit = true; break; }            // :195
        ++rb;
    }
    if (!foundHit) return false;                                       // :204

    bool lastSearch = false;
    // ---- Phase B (:208-329)
    if (fwdHit) {
        OffsetT matchLen = fwdSAInts.front().len;
        rb = fwdSAInts.front().queryPos;
        OffsetT remainingLength = (OffsetT)(readLen - (rb + matchLen));          // :219
        int lceV = lce(ss, lbLeftFwd, ubLeftFwd - 1, matchLen, remainingLength);  // :220
        OffsetT fwdSkip = std::max((OffsetT)matchLen - skipOverlap, (OffsetT)lceV - skipOverlap);
        size_t nextInformativePosition = (size_t)std::min(
            std::max((OffsetT)0, (OffsetT)readLen - (OffsetT)k),
            (OffsetT)(rb + fwdSkip));                                  // :224-228
        rb = (int64_t)nextInformativePosition;
        size_t invalidPos = 0;
        while (rb + k <= readLen) {                                    // :234
            size_t pos = (size_t)rb;
            if (invalidPos != NPOS) invalidPos = find_first_nN(read, (size_t)readLen, pos);   // :242-244
            if (invalidPos < pos + (size_t)k) {                        // :247
                rb = (int64_t)invalidPos + 1;
                continue;
            }
            mer = mer_from_chars(read + pos, (int)k);                  // :259
            if (mer_is_homopolymer(mer, (int)k)) { rb += homoPolymerSkip; continue; }
            auto merIt = probe(mer);                                   // :261
            if (merIt) {
                if (strictCheck) {
                    ++fwdHit;
                    kmerScores.push_back({mer, (int32_t)pos, PRESENT, UNTESTED});
                    uint64_t rcm = mer_revcomp(mer, (int)k);
                    auto rcIt = probe(rcm);                            // :268
                    if (rcIt) { ++rcHit; kmerScores.back().rcScore = PRESENT; }
                }
                lbRightFwd = merIt->begin; ubRightFwd = merIt->end;
                lbRightFwd = std::max((OffsetT)0, lbRightFwd - 1);     // :279
                Query q{read, (int)readLen, (int)rb, false};
                extendSearchNaive(ss, lbRightFwd, ubRightFwd, (OffsetT)k, q, lbRightFwd, ubRightFwd, matchedLen);
                OffsetT diff = ubRightFwd - lbRightFwd;
                if (ubRightFwd > lbRightFwd && diff < maxInterval) {   // :285
                    fwdSAInts.push_back({lbRightFwd, ubRightFwd, matchedLen, (uint32_t)rb, false});
                    if (strictCheck && rb + matchedLen < readLen) {    // :291
                        int64_t kmerPos = rb + matchedLen - skipOverlap;
                        mer = mer_from_chars(read + kmerPos, (int)k);
                        kmerScores.push_back({mer, (int32_t)kmerPos, UNTESTED, UNTESTED});
                    }
                }
                if (lastSearch) break;                                 // :300
                int64_t mismatch = rb + matchedLen;
                if (mismatch < readLen) {
                    OffsetT remainingDistance = (OffsetT)(readLen - mismatch);
                    int lce2 = lce(ss, lbRightFwd, ubRightFwd - 1, matchedLen, remainingDistance);   // :304
                    int64_t skipMatch = mismatch - skipOverlap;
                    int64_t skipLCE = rb + lce2 - skipOverlap;
                    rb = std::max(skipLCE, skipMatch);
                    if (rb > readLen - k) { rb = readLen - k; lastSearch = true; }   // :312-315
                } else {
                    lastSearch = true;                                 // :318
                    rb = readLen - k;
                }
            } else {
                rb += sampFactor;                                      // :324
            }
        }
    }

    lastSearch = false;                                                // :331
    // ---- Phase C (:332-440): scan of the reversed read with complemented bases.
    //      `o` is the distance from read.rbegin(); pos = L - (o + k) is the forward offset.
    if (rcHit >= fwdHit) {
        int64_t o = 0;
        while (o + k <= readLen) {                                     // :341-344
            // first n/N inside the reversed window [o, o+k)            :349-362
            int64_t inv = -1;
            for (int64_t j = o; j < o + k; ++j) {
                char c = read[readLen - 1 - j];
                if (c == 'n' || c == 'N') { inv = j; break; }
            }
            if (inv >= 0) { o = inv + 1; continue; }
            size_t pos = (size_t)(readLen - (o + k));                  // :366
            mer = mer_from_chars(read + pos, (int)k);                  // :369
            if (mer_is_homopolymer(mer, (int)k)) { o += homoPolymerSkip; continue; }
            rcMer = mer_revcomp(mer, (int)k);
            auto rcMerIt = probe(rcMer);                               // :372
            if (rcMerIt) {
                if (strictCheck) {
                    ++rcHit;
                    kmerScores.push_back({mer, (int32_t)pos, UNTESTED, PRESENT});
                    auto merIt = probe(mer);                           // :379
                    if (merIt) { ++fwdHit; kmerScores.back().fwdScore = PRESENT; }
                }
                lbRightRC = rcMerIt->begin; ubRightRC = rcMerIt->end;
                lbRightRC = std::max((OffsetT)0, lbRightRC - 1);       // :392
                Query q{read, (int)readLen, (int)o, true};
                extendSearchNaive(ss, lbRightRC, ubRightRC, (OffsetT)k, q, lbRightRC, ubRightRC, matchedLen);   // :393
                OffsetT diff = ubRightRC - lbRightRC;
                if (ubRightRC > lbRightRC && diff < maxInterval) {     // :398
                    rcSAInts.push_back({lbRightRC, ubRightRC, matchedLen, (uint32_t)o, true});   // queryPos on the reversed read :399
                    if (strictCheck && o + matchedLen < readLen) {     // :404
                        int64_t kmerPos = readLen - (o + matchedLen);  // :405 distance(revRB + matchedLen, rend)
                        mer = mer_from_chars(read + kmerPos, (int)k);
                        kmerScores.push_back({mer, (int32_t)kmerPos, UNTESTED, UNTESTED});
                    }
                }
                if (lastSearch) break;                                 // :412
                int64_t mismatch = o + matchedLen;
                if (mismatch < readLen) {
                    OffsetT remainingDistance = (OffsetT)(readLen - mismatch);
                    int lce2 = lce(ss, lbRightRC, ubRightRC - 1, matchedLen, remainingDistance);     // :416
                    int64_t skipMatch = mismatch - skipOverlap;
                    int64_t skipLCE = o + lce2 - skipOverlap;
                    o = std::max(skipLCE, skipMatch);
                    if (o > readLen - k) { o = readLen - k; lastSearch = true; }
                } else {
                    lastSearch = true;
                    o = readLen - k;
                }
            } else {
                o += sampFactor;
            }
        }
    }

    // ---- strand vote (:442-490)
    if (strictCheck) {
        if (fwdHit > 0 && rcHit == 0) {
            rcSAInts.clear();
        } else if (rcHit > 0 && fwdHit == 0) {
            fwdSAInts.clear();
        } else {
            // libstdc++ std::sort on kpos only: insertion sort (stable) for <= 16 entries, introsort
            // above — the oracle uses the library call itself so that equal-kpos survivors are
            // whatever libstdc++ picks (SURVEY §7 hard part 2).
            std::sort(kmerScores.begin(), kmerScores.end());           // :450
            auto e = std::unique(kmerScores.begin(), kmerScores.end());
            int32_t fwdScore = 0, rcScore = 0;
            for (auto it = kmerScores.begin(); it != e; ++it) {
                auto& kms = *it;
                if (kms.fwdScore == UNTESTED) kms.fwdScore = probe(kms.kmer) ? PRESENT : ABSENT;   // :461-464
                fwdScore += kms.fwdScore;
                if (kms.rcScore == UNTESTED) {
                    uint64_t r = mer_revcomp(kms.kmer, (int)k);
                    kms.rcScore = probe(r) ? PRESENT : ABSENT;         // :469-473
                }
                rcScore += kms.rcScore;
            }
            if (fwdScore > rcScore) rcSAInts.clear();                  // :482
            else if (rcScore > fwdScore) fwdSAInts.clear();
        }
    }

    // ---- interval -> transcript hits (:492-579)
    auto single = [&](const SAIntervalHit& h, bool fwd, size_t initialSize) {
        for (OffsetT i = h.begin; i != h.end; ++i) {
            auto globalPos = ix.SA[i];
            auto txpID = ix.transcriptAtPosition(globalPos);
            ++ss.wc.H;
            int32_t pos = (int32_t)(globalPos - (OffsetT)ix.txpOffsets[txpID]);
            QuasiAlignment qa; qa.tid = txpID; qa.pos = pos - (int32_t)h.queryPos; qa.fwd = fwd;
            qa.readLen = (uint32_t)readLen; qa.mateStatus = mateStatus;
            hits.push_back(qa);
        }
        std::sort(hits.begin() + initialSize, hits.end(), [](const QuasiAlignment& a, const QuasiAlignment& b) {
            return a.tid == b.tid ? a.pos < b.pos : a.tid < b.tid; });
        auto ne = std::unique(hits.begin() + initialSize, hits.end(),
                              [](const QuasiAlignment& a, const QuasiAlignment& b) { return a.tid == b.tid; });
        hits.resize(ne - hits.begin());
    };
    size_t fwdHitsStart = hits.size();
    if (fwdSAInts.size() > 1) {
        auto processed = intersectSAHits(fwdSAInts, ix, ss.wc);
        collectHitsSimpleSA(processed, (uint32_t)readLen, maxDist, hits, mateStatus);
    } else if (fwdSAInts.size() == 1) {
        single(fwdSAInts.front(), true, hits.size());
    }
    size_t fwdHitsEnd = hits.size();
    size_t rcHitsStart = fwdHitsEnd;
    if (rcSAInts.size() > 1) {
        auto processed = intersectSAHits(rcSAInts, ix, ss.wc);
        collectHitsSimpleSA(processed, (uint32_t)readLen, maxDist, hits, mateStatus);
    } else if (rcSAInts.size() == 1) {
        single(rcSAInts.front(), false, hits.size());
    }
    size_t rcHitsEnd = hits.size();
    if (fwdHitsEnd > fwdHitsStart && rcHitsEnd > rcHitsStart) {        // :567-579
        std::inplace_merge(hits.begin() + fwdHitsStart, hits.begin() + fwdHitsEnd, hits.begin() + rcHitsEnd,
                           [](const QuasiAlignment& a, const QuasiAlignment& b) { return a.tid < b.tid; });
        auto ne = std::unique(hits.begin() + fwdHitsStart, hits.begin() + rcHitsEnd,
                              [](const QuasiAlignment& a, const QuasiAlignment& b) { return a.tid == b.tid; });
        hits.resize(ne - hits.begin());
    }
    return foundHit;
}

} // namespace orc
