// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// CPU restatement of the Circall quasi-mapping hot path.  Nothing under
// circall_b200/ may include, link or call this code; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Evidence grades used in comments (same as SURVEY.md):
//   [R] restates code that is readable under /root/reference (file:line cited,
//       paths relative to /root/reference/Circall_v1.0.1/)
//   [U] restates un-shipped upstream code (RapMap / Sailfish v0.10.0 / Jellyfish
//       2.2.5) from its published algorithm; SURVEY.md Appendix B item cited.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>
#include <atomic>

namespace orc {

using OffsetT = int32_t;   // reference picks int32 when the text is < 2^31 (Circall_wt.cpp:1168,1210-1268)

// [U] rapmap::utils::SAIntervalHit<OffsetT>; used SACollectorCircall.hpp:151,287,400
struct SAIntervalHit {
    OffsetT begin, end, len;
    uint32_t queryPos;
    bool queryRC;
    OffsetT span() const { return end - begin; }
};

// [U] rapmap::utils::QuasiAlignment (fields the hot path touches)
enum MateStatus { SINGLE_END = 0, PAIRED_END_LEFT = 1, PAIRED_END_RIGHT = 2, PAIRED_END_PAIRED = 3 };
struct QuasiAlignment {
    uint32_t tid;
    int32_t pos;
    bool fwd;
    uint32_t readLen;
    uint32_t fragLen = 0;
    int32_t matePos = 0;
    bool mateIsFwd = false;
    MateStatus mateStatus = SINGLE_END;
};

// open-addressing stand-in for google::dense_hash_map<uint64_t, SAInterval>
struct KHash {
    struct Slot { uint64_t key; OffsetT begin, end; };
    std::vector<Slot> slots;
    uint64_t mask = 0;
    uint64_t size = 0;
    static constexpr uint64_t EMPTY = ~0ULL;
    static inline uint64_t mix(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
        return x;
    }
    void init(uint64_t nkeys);
    void insert(uint64_t key, OffsetT b, OffsetT e);
    const Slot* find(uint64_t key) const {
        uint64_t h = mix(key) & mask;
        for (;;) {
            const Slot& s = slots[h];
            if (s.key == key) return &s;
            if (s.key == EMPTY) return nullptr;
            h = (h + 1) & mask;
        }
    }
};

// rank over the '$' bit vector ([U] rank9b in RapMap): rank(p) = number of '$' in seq[0,p)
struct RankDict {
    std::vector<uint64_t> bits;
    std::vector<uint32_t> cum;   // '$' count before each 64-bit word
    void build(const std::string& seq);
    inline uint32_t rank(uint64_t p) const {
        uint64_t w = p >> 6, o = p & 63;
        uint64_t m = o ? (bits[w] & ((1ULL << o) - 1)) : 0;
        return cum[w] + (uint32_t)__builtin_popcountll(m);
    }
};

// [U] RapMapSAIndex<int32_t, DenseHash>
struct Index {
    int k = 31;
    std::string seq;                    // transcripts upper-cased, each followed by '$'
    std::vector<OffsetT> SA;
    std::vector<uint32_t> txpOffsets, txpLens;
    std::vector<std::string> txpNames;
    KHash khash;
    RankDict rankDict;
    inline uint32_t transcriptAtPosition(uint64_t p) const { return rankDict.rank(p); }
};

// instrumentation for SURVEY §8(d) algorithmic-bytes formula (P, S, T, H)
struct WorkCounters {
    uint64_t P = 0;   // k-mer table probes
    uint64_t S = 0;   // SA entries read in extendSearchNaive / lce
    uint64_t T = 0;   // text bases compared in extendSearchNaive / lce
    uint64_t H = 0;   // SA positions enumerated in hit collection / span filter
    uint64_t calls = 0;
    void add(const WorkCounters& o) { P += o.P; S += o.S; T += o.T; H += o.H; calls += o.calls; }
};

struct Searcher {   // [U] SASearcher<RapMapIndexT>
    const Index* ix;
    mutable WorkCounters wc;
    int lceMode = 0;   // 0 = LCE_LITERAL (SURVEY B1 default), 1 = LCE_CLEAN
};

// query-character accessor: fwd -> toupper(read[pos+i]); rc -> complement(toupper(read[L-1-(pos+i)]))
struct Query {
    const char* read; int L; int start; bool rc;
    inline int remaining() const { return L - start; }
    char at(int i) const;
};

} // namespace orc
