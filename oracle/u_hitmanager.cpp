// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [U] rapmap::hit_manager::intersectSAHits / collectHitsSimpleSA, called at
// SACollectorCircall.hpp:495-496,532-533; and rapmap::utils::mergeLeftRightHitsFuzzy (SURVEY
// A13, B4), live only at Circall_pseudo.cpp:271-280.
#include "oracle_types.hpp"
#include <algorithm>

namespace orc {

struct SATxpQueryPos { uint32_t pos; uint32_t queryPos; bool queryRC; };
struct ProcessedSAHit { std::vector<SATxpQueryPos> tqvec; uint32_t numActive = 0; bool active = false; };
using SAHitMap = std::map<int, ProcessedSAHit>;

SAHitMap intersectSAHits(std::vector<SAIntervalHit>& inHits, const Index& ix, WorkCounters& wc) {
    SAHitMap outHits;
    if (inHits.size() < 2) return outHits;
    // smallest span, first on ties
    SAIntervalHit* minHit = &inHits[0];
    for (auto& h : inHits)
        if (h.span() < minHit->span()) minHit = &h;
    for (OffsetT i = minHit->begin; i < minHit->end; ++i) {
        auto globalPos = ix.SA[i];
        auto tid = ix.transcriptAtPosition(globalPos);
        auto txpPos = globalPos - ix.txpOffsets[tid];
        ++wc.H;
        auto& e = outHits[(int)tid];
        e.numActive = 1;
        e.tqvec.push_back({(uint32_t)txpPos, minHit->queryPos, minHit->queryRC});
    }
    size_t intervalCounter = 2;
    for (auto& h : inHits) {
        if (&h == minHit) continue;
        for (OffsetT i = h.begin; i < h.end; ++i) {
            auto globalPos = ix.SA[i];
            auto tid = ix.transcriptAtPosition(globalPos);
            ++wc.H;
            auto it = outHits.find((int)tid);
            if (it != outHits.end()) {
                it->second.numActive += (it->second.numActive == intervalCounter - 1) ? 1 : 0;
                if (it->second.numActive == intervalCounter) {
                    auto localPos = globalPos - ix.txpOffsets[tid];
                    it->second.tqvec.push_back({(uint32_t)localPos, h.queryPos, h.queryRC});
                }
            }
        }
        ++intervalCounter;
    }
    size_t required = inHits.size();
    for (auto& kv : outHits) kv.second.active = (kv.second.numActive >= required);   // consistentHits=false everywhere
    return outHits;
}

void collectHitsSimpleSA(SAHitMap& processed, uint32_t readLen, uint32_t /*maxDist*/,
                         std::vector<QuasiAlignment>& hits, MateStatus ms) {
    for (auto& ph : processed) {
        if (!ph.second.active) continue;
        auto minIt = std::min_element(ph.second.tqvec.begin(), ph.second.tqvec.end(),
                                      [](const SATxpQueryPos& a, const SATxpQueryPos& b) { return a.pos < b.pos; });
        QuasiAlignment qa;
        qa.tid = (uint32_t)ph.first;
        qa.pos = (int32_t)minIt->pos - (int32_t)minIt->queryPos;
        qa.fwd = !minIt->queryRC;
        qa.readLen = readLen;
        qa.mateStatus = ms;
        hits.push_back(qa);
    }
}

// SURVEY B4 (confidence: medium).  Two-pointer merge of tid-sorted left/right lists.
void mergeLeftRightHitsFuzzy(bool leftMatches, bool rightMatches,
                             std::vector<QuasiAlignment>& leftHits, std::vector<QuasiAlignment>& rightHits,
                             std::vector<QuasiAlignment>& jointHits, uint32_t readLen, uint32_t maxNumHits,
                             bool& tooManyHits) {
    (void)readLen;
    if (leftHits.empty()) {
        if (!leftMatches && !rightHits.empty()) {   // left mate had no seed at all: right orphans
            jointHits.insert(jointHits.end(), rightHits.begin(), rightHits.end());
        }
        return;
    }
    if (rightHits.empty()) {
        if (!rightMatches) jointHits.insert(jointHits.end(), leftHits.begin(), leftHits.end());
        return;
    }
    auto l = leftHits.begin(), le = leftHits.end();
    auto r = rightHits.begin(), re = rightHits.end();
    size_t numHits = 0;
    while (l != le && r != re) {
        if (l->tid < r->tid) ++l;
        else if (r->tid < l->tid) ++r;
        else {
            QuasiAlignment j;
            int32_t startRead1 = std::max(l->pos, 0);
            int32_t startRead2 = std::max(r->pos, 0);
            bool read1First = startRead1 < startRead2;
            int32_t fragStartPos = read1First ? startRead1 : startRead2;
            int32_t fragEndPos = read1First ? (startRead2 + (int32_t)r->readLen) : (startRead1 + (int32_t)l->readLen);
            j.tid = l->tid; j.pos = startRead1; j.fwd = l->fwd; j.readLen = l->readLen;
            j.fragLen = (uint32_t)(fragEndPos - fragStartPos);
            j.matePos = startRead2; j.mateIsFwd = r->fwd; j.mateStatus = PAIRED_END_PAIRED;
            jointHits.push_back(j);
            ++numHits;
            if (numHits > maxNumHits) { tooManyHits = true; break; }
            ++l; ++r;
        }
    }
    if (jointHits.empty()) {   // both mates hit but nothing in common: all left, then all right, as orphans
        jointHits.insert(jointHits.end(), leftHits.begin(), leftHits.end());
        jointHits.insert(jointHits.end(), rightHits.begin(), rightHits.end());
    }
}

} // namespace orc
