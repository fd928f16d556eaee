// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
#pragma once
#include "oracle_types.hpp"

namespace orc {

struct ReadBatch {            // concatenated bases + (n+1) offsets per mate file
    const char* r1; const uint64_t* off1;
    const char* r2; const uint64_t* off2;
    uint64_t n;
};

using EqMap = std::map<std::vector<uint32_t>, uint64_t>;

struct WtResult {
    std::vector<uint64_t> expPair;    // pair index of every exported record, in input order
    std::vector<uint8_t> expCode;     // 1: mate 1 unmapped (headers ___11/___22); 2: mate 2 unmapped (___21/___12)
    std::vector<uint8_t> mateFlags;   // per mate (2 per pair): bit0 = seeded (any interval), bit1 = some interval has len == readLen
    std::vector<uint16_t> mateNHits;  // per mate: hits after the >200 clear
    uint64_t numObserved = 0, numMapped = 0, numHits = 0, upperBound = 0;
    uint32_t readLen = 0;
    EqMap eq;
    WorkCounters wc;
};

struct JuncLine { uint64_t rec; uint8_t dir; uint32_t queryPos; int32_t len; int32_t hitPos; uint32_t tid; };

struct BsjResult {
    std::vector<JuncLine> junc;       // junction-table rows (Circall_bsj.cpp:339-341,366-368)
    std::vector<uint64_t> cand;       // records written to BSJ_read{1,2}_*.fa
    std::vector<uint32_t> recNHits;   // per record: jointHits.size() before the >200 clear
    std::vector<uint8_t> recSplit;    // per record: isSplit
    uint64_t numObserved = 0, numMapped = 0, numHits = 0, upperBound = 0;
    uint32_t readLen = 0;
    EqMap eq;
    WorkCounters wc;
};

void run_wt(const Index& ix, const ReadBatch& rb, int nthreads, int lceMode, WtResult& out);
void run_bsj(const Index& ix, const ReadBatch& rb, int nthreads, int lceMode, BsjResult& out);
void export_records(const ReadBatch& rb, const WtResult& wt, std::string& r1, std::vector<uint64_t>& o1,
                    std::string& r2, std::vector<uint64_t>& o2);

void index_from_text(Index& ix, const char* seq, uint64_t n, int k, int nthreads);
void index_from_text_with_sa(Index& ix, const char* seq, uint64_t n, int k, const OffsetT* sa, int nthreads);
void index_from_records(Index& ix, const std::vector<std::pair<std::string, std::string>>& recs, int k, int nthreads);

} // namespace orc
