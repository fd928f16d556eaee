// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [R] restatement of the per-pair bodies of
//   Circall_wt  processReadsQuasi(paired)  src/Circall_wt.cpp:237-660   
//   Circall_bsj processReadsQuasi(paired)  src/Circall_bsj.cpp:225-505  
// and of EquivalenceClassBuilder::addGroup (include/EquivalenceClassBuilder.hpp:112-130, A10).
// Worker threads pull 1000-pair jobs like the reference (Circall_wt.cpp:98,1195-1276); results are
// merged in job order so the oracle itself is deterministic.
#include "oracle_tools.hpp"
#include <algorithm>
#include <thread>
#include <atomic>
#include <cmath>

namespace orc {

bool collect(const Searcher&, const char*, int, std::vector<QuasiAlignment>&,
             std::vector<SAIntervalHit>&, std::vector<SAIntervalHit>&, MateStatus, bool);

static const uint32_t maxReadOccs = 200;        // Circall_wt.cpp:1515
static const uint64_t readGroupSize = 1000;     // Circall_wt.cpp:98

template <class Local, class Body>
static void run_jobs(uint64_t n, int nthreads, std::vector<Local>& locals, Body body) {
    uint64_t njobs = (n + readGroupSize - 1) / readGroupSize;
    locals.resize(njobs);
    std::atomic<uint64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            uint64_t j = next.fetch_add(1);
            if (j >= njobs) break;
            uint64_t b = j * readGroupSize, e = std::min(n, b + readGroupSize);
            body(locals[j], b, e);
        }
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

void run_wt(const Index& ix, const ReadBatch& rb, int nthreads, int lceMode, WtResult& out) {
    std::vector<WtResult> locals;
    run_jobs(rb.n, nthreads, locals, [&](WtResult& L, uint64_t b, uint64_t e) {
        Searcher ss{&ix}; ss.lceMode = lceMode;
        std::vector<QuasiAlignment> leftHits, rightHits;
        std::vector<SAIntervalHit> leftFwd, leftRc, rightFwd, rightRc;
        std::vector<uint32_t> txpIDsCompat;
        for (uint64_t i = b; i < e; ++i) {
            const char* s1 = rb.r1 + rb.off1[i]; size_t l1 = rb.off1[i + 1] - rb.off1[i];
            const char* s2 = rb.r2 + rb.off2[i]; size_t l2 = rb.off2[i + 1] - rb.off2[i];
            size_t readLen = l1;                                          // :242 (mate 1's length, used for both mates)
            leftHits.clear(); rightHits.clear(); txpIDsCompat.clear();    // :244-250
            bool mappedFrag = false;                                      // :252 (not reset between mates)
            leftFwd.clear(); leftRc.clear(); rightFwd.clear(); rightRc.clear();
            collect(ss, s1, (int)l1, leftHits, leftFwd, leftRc, PAIRED_END_LEFT, true);    // :275
            collect(ss, s2, (int)l2, rightHits, rightFwd, rightRc, PAIRED_END_RIGHT, true); // :283
            if (L.readLen == 0) L.readLen = (uint32_t)readLen;            // :303
            // ---- left mate :310-365
            bool isKmerMapped = !(leftFwd.empty() && leftRc.empty());
            bool isUnmapped = true;
            for (auto& h : leftFwd) if ((size_t)h.len == readLen) { isUnmapped = false; break; }
            for (auto& h : leftRc) if ((size_t)h.len == readLen) { isUnmapped = false; break; }
            uint8_t fl = (uint8_t)((isKmerMapped ? 1 : 0) | (isUnmapped ? 0 : 2));
            if (isUnmapped && isKmerMapped) { L.expPair.push_back(i); L.expCode.push_back(1); }   // ___11 / ___22
            L.upperBound += (leftHits.size() > 0);                        // :369
            if (leftHits.size() > maxReadOccs) leftHits.clear();          // :372
            if (!leftHits.empty()) {
                // libType IU => compatibleHit() is true for every hit: all tids go to txpIDsCompat
                for (auto& h : leftHits) txpIDsCompat.push_back(h.tid);
                mappedFrag = true;
                L.eq[txpIDsCompat] += 1;                                  // :450-451
            }
            L.numMapped += mappedFrag ? 1 : 0;                            // :467
            L.numHits += leftHits.size();
            L.numObserved += 1;
            L.mateFlags.push_back(fl); L.mateNHits.push_back((uint16_t)std::min<size_t>(leftHits.size(), 65535));
            // ---- right mate :485-645
            isKmerMapped = !(rightFwd.empty() && rightRc.empty());
            isUnmapped = true;
            for (auto& h : rightFwd) if ((size_t)h.len == readLen) { isUnmapped = false; break; }
            for (auto& h : rightRc) if ((size_t)h.len == readLen) { isUnmapped = false; break; }
            fl = (uint8_t)((isKmerMapped ? 1 : 0) | (isUnmapped ? 0 : 2));
            if (isUnmapped && isKmerMapped) { L.expPair.push_back(i); L.expCode.push_back(2); }   // ___21 / ___12
            L.upperBound += (rightHits.size() > 0);
            if (rightHits.size() > maxReadOccs) rightHits.clear();
            if (!rightHits.empty()) {
                for (auto& h : rightHits) txpIDsCompat.push_back(h.tid);  // appended to mate 1's ids: :247-252 quirk
                mappedFrag = true;
                L.eq[txpIDsCompat] += 1;
            }
            L.numMapped += mappedFrag ? 1 : 0;                            // :642 (sticky across mates)
            L.numHits += rightHits.size();
            L.numObserved += 1;
            L.mateFlags.push_back(fl); L.mateNHits.push_back((uint16_t)std::min<size_t>(rightHits.size(), 65535));
        }
        L.wc = ss.wc;
    });
    for (auto& L : locals) {
        out.expPair.insert(out.expPair.end(), L.expPair.begin(), L.expPair.end());
        out.expCode.insert(out.expCode.end(), L.expCode.begin(), L.expCode.end());
        out.mateFlags.insert(out.mateFlags.end(), L.mateFlags.begin(), L.mateFlags.end());
        out.mateNHits.insert(out.mateNHits.end(), L.mateNHits.begin(), L.mateNHits.end());
        out.numObserved += L.numObserved; out.numMapped += L.numMapped; out.numHits += L.numHits; out.upperBound += L.upperBound;
        for (auto& kv : L.eq) out.eq[kv.first] += kv.second;
        if (out.readLen == 0) out.readLen = L.readLen;
        out.wc.add(L.wc);
    }
}

void run_bsj(const Index& ix, const ReadBatch& rb, int nthreads, int lceMode, BsjResult& out) {
    std::vector<BsjResult> locals;
    run_jobs(rb.n, nthreads, locals, [&](BsjResult& L, uint64_t b, uint64_t e) {
        Searcher ss{&ix}; ss.lceMode = lceMode;
        std::vector<QuasiAlignment> jointHits;
        std::vector<SAIntervalHit> leftFwd, leftRc;
        std::vector<uint32_t> txpIDsCompat;
        for (uint64_t i = b; i < e; ++i) {
            const char* s1 = rb.r1 + rb.off1[i]; size_t l1 = rb.off1[i + 1] - rb.off1[i];
            jointHits.clear(); leftFwd.clear(); leftRc.clear(); txpIDsCompat.clear();
            bool mappedFrag = false;
            collect(ss, s1, (int)l1, jointHits, leftFwd, leftRc, SINGLE_END, true);     // :300-307 (read 2 is never mapped)
            if (L.readLen == 0) L.readLen = (uint32_t)l1;                 // :310
            bool isSplit = false;
            if (!jointHits.empty()) {                                     // :319
                for (int dir = 0; dir < 2; ++dir) {
                    auto& ints = dir == 0 ? leftFwd : leftRc;
                    for (auto& h : ints) {
                        for (OffsetT sa_i = h.begin; sa_i != h.end; ++sa_i) {           // :326
                            auto globalPos = ix.SA[sa_i];
                            auto txpID = ix.transcriptAtPosition(globalPos);
                            ++ss.wc.H;
                            int32_t hitPos = (int32_t)(globalPos - (OffsetT)ix.txpOffsets[txpID]);   // seed offset, not pos-queryPos
                            int32_t junctPos = (int32_t)std::floor(ix.txpLens[txpID] / 2);          // :333 (integer division)
                            if (hitPos < junctPos - 5 && hitPos + h.len > junctPos + 5) {           // :336,364
                                L.junc.push_back({i, (uint8_t)dir, h.queryPos, h.len, hitPos, txpID});
                                isSplit = true;
                            }
                        }
                    }
                }
            }
            L.upperBound += (jointHits.size() > 0);                       // :383
            uint32_t nh0 = (uint32_t)jointHits.size();
            if (jointHits.size() > maxReadOccs) jointHits.clear();        // :385
            if (!jointHits.empty() && isSplit) {                          // :390
                L.cand.push_back(i);
                for (auto& h : jointHits) txpIDsCompat.push_back(h.tid);
                mappedFrag = true;
                L.eq[txpIDsCompat] += 1;
            }
            L.numMapped += mappedFrag ? 1 : 0;                            // :488
            L.numHits += jointHits.size();
            L.numObserved += 1;
            L.recNHits.push_back(nh0); L.recSplit.push_back(isSplit ? 1 : 0);
        }
        L.wc = ss.wc;
    });
    for (auto& L : locals) {
        out.junc.insert(out.junc.end(), L.junc.begin(), L.junc.end());
        out.cand.insert(out.cand.end(), L.cand.begin(), L.cand.end());
        out.recNHits.insert(out.recNHits.end(), L.recNHits.begin(), L.recNHits.end());
        out.recSplit.insert(out.recSplit.end(), L.recSplit.begin(), L.recSplit.end());
        out.numObserved += L.numObserved; out.numMapped += L.numMapped; out.numHits += L.numHits; out.upperBound += L.upperBound;
        for (auto& kv : L.eq) out.eq[kv.first] += kv.second;
        if (out.readLen == 0) out.readLen = L.readLen;
        out.wc.add(L.wc);
    }
}

// Step 1 output -> Step 2 input, as Circall_source.sh:273-291 does through UN_read{1,2}_*.fa:
// record = (unmapped mate first, the other mate second).
void export_records(const ReadBatch& rb, const WtResult& wt, std::string& r1, std::vector<uint64_t>& o1,
                    std::string& r2, std::vector<uint64_t>& o2) {
    o1.assign(1, 0); o2.assign(1, 0);
    for (size_t j = 0; j < wt.expPair.size(); ++j) {
        uint64_t i = wt.expPair[j];
        const char* s1 = rb.r1 + rb.off1[i]; size_t l1 = rb.off1[i + 1] - rb.off1[i];
        const char* s2 = rb.r2 + rb.off2[i]; size_t l2 = rb.off2[i + 1] - rb.off2[i];
        if (wt.expCode[j] == 1) { r1.append(s1, l1); r2.append(s2, l2); }
        else { r1.append(s2, l2); r2.append(s1, l1); }
        o1.push_back(r1.size()); o2.push_back(r2.size());
    }
}

} // namespace orc
