// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [U] RapMap indexTranscriptsSA / buildHash, reached from
// TxIndexer.cpp:202-227 through SailfishIndex::build; consumed at
// SACollectorCircall.hpp:37-41 and ReadExperiment.hpp:122-135.
//
// The suffix array is the true suffix array of the whole '$'-joined text (what
// libdivsufsort returns; it is unique, so any correct sorter gives the same array).
// Built by induced sorting (u_sais.cpp); the bucketed comparison sort below is its cross-check
// at test sizes — both deliberately different algorithms from the product's GPU
// prefix-doubling builder so that they check each other.
#include "oracle_types.hpp"
#include "u_mer.hpp"
#include <algorithm>
#include <cstring>
#include <random>
#include <thread>
#include <stdexcept>
#include <atomic>

namespace orc {

void KHash::init(uint64_t nkeys) {
    uint64_t cap = 16;
    while (cap < nkeys * 2 + 2) cap <<= 1;
    slots.assign(cap, Slot{EMPTY, 0, 0});
    mask = cap - 1;
    size = 0;
}

void KHash::insert(uint64_t key, OffsetT b, OffsetT e) {
    uint64_t h = mix(key) & mask;
    while (slots[h].key != EMPTY) {
        if (slots[h].key == key) throw std::runtime_error("khash: duplicate k-mer interval");
        h = (h + 1) & mask;
    }
    slots[h] = Slot{key, b, e};
    ++size;
}

void RankDict::build(const std::string& seq) {
    size_t n = seq.size();
    size_t nw = n / 64 + 2;
    bits.assign(nw, 0);
    cum.assign(nw + 1, 0);
    for (size_t i = 0; i < n; ++i)
        if (seq[i] == '$') bits[i >> 6] |= 1ULL << (i & 63);
    uint32_t c = 0;
    for (size_t w = 0; w < nw; ++w) { cum[w] = c; c += (uint32_t)__builtin_popcountll(bits[w]); }
    cum[nw] = c;
}

static inline int sa_code(char c) {   // order-preserving: '$' < A < C < G < T, pad below '$'
    switch (c) { case '$': return 1; case 'A': return 2; case 'C': return 3; case 'G': return 4; case 'T': return 5; }
    throw std::runtime_error("index text holds a character outside {A,C,G,T,$}");
}

void build_sa_sais(const std::string& seq, std::vector<OffsetT>& SA);     // u_sais.cpp

// the bucketed comparison sort: quadratic-ish on repetitive text, kept as the cross-check of u_sais.cpp at test sizes
void build_sa_compare(const std::string& seq, std::vector<OffsetT>& SA, int nthreads) {
    const size_t n = seq.size();
    if (n >= (1ULL << 31)) throw std::runtime_error("text >= 2^31: 64-bit index not supported by the oracle");
    SA.resize(n);
    const int P = 5;                       // bucket prefix length
    size_t nb = 1; for (int i = 0; i < P; ++i) nb *= 6;
    std::vector<uint64_t> cnt(nb + 1, 0);
    const char* s = seq.data();
    auto bucket_of = [&](size_t i) -> size_t {
        size_t b = 0;
        for (int j = 0; j < P; ++j) b = b * 6 + (i + j < n ? sa_code(s[i + j]) : 0);
        return b;
    };
    std::vector<uint32_t> bid(n);
    for (size_t i = 0; i < n; ++i) { size_t b = bucket_of(i); bid[i] = (uint32_t)b; ++cnt[b + 1]; }
    for (size_t b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
    {
        std::vector<uint64_t> fill(cnt.begin(), cnt.end() - 1);
        for (size_t i = 0; i < n; ++i) SA[fill[bid[i]]++] = (OffsetT)i;
    }
    std::vector<uint32_t>().swap(bid);
    auto cmp = [s, n, P](OffsetT a, OffsetT b) -> bool {
        // the first P characters are equal inside a bucket (or both ran off the end)
        size_t la = n - (size_t)a, lb = n - (size_t)b;
        size_t m = std::min(la, lb);
        if (m > (size_t)P) {
            int r = std::memcmp(s + a + P, s + b + P, m - P);
            if (r != 0) return r < 0;
        }
        return la < lb;   // proper prefix sorts first
    };
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            size_t b = next.fetch_add(1);
            if (b >= nb) break;
            if (cnt[b + 1] - cnt[b] > 1) std::sort(SA.begin() + cnt[b], SA.begin() + cnt[b + 1], cmp);
        }
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

template <class F>
static void parallel_ranges(size_t n, int nthreads, F f) {
    if (nthreads < 1) nthreads = 1;
    if (n < (size_t)1 << 16) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() { f(n * (size_t)t / (size_t)nthreads, n * (size_t)(t + 1) / (size_t)nthreads, t); });
    for (auto& t : th) t.join();
}

// walk the SA once; consecutive suffixes whose first k characters are equal
// and '$'-free form one [begin,end).  The walk is split over threads: a first pass classifies
// every SA slot (0 = continues the previous slot's k-mer, 1 = opens a run, 2 = no valid k-mer),
// a second pass inserts every run (slot claimed with a compare-and-swap on its key).
static void build_khash(Index& ix, int nthreads) {
    const size_t n = ix.seq.size();
    const int k = ix.k;
    const char* s = ix.seq.data();
    auto valid = [&](OffsetT p) -> bool {
        if ((size_t)p + k > n) return false;
        return std::memchr(s + p, '$', k) == nullptr;
    };
    std::vector<uint8_t> cls(n);
    std::vector<uint64_t> heads((size_t)std::max(1, nthreads), 0);
    parallel_ranges(n, nthreads, [&](size_t b, size_t e, int t) {
        uint64_t h = 0;
        bool pv = b > 0 && valid(ix.SA[b - 1]);
        for (size_t i = b; i < e; ++i) {
            OffsetT p = ix.SA[i];
            bool v = valid(p);
            uint8_t c = 2;
            if (v) { c = (pv && std::memcmp(s + ix.SA[i - 1], s + p, k) == 0) ? 0 : 1; h += c; }
            cls[i] = c;
            pv = v;
        }
        heads[t] = h;
    });
    uint64_t nk = 0;
    for (auto h : heads) nk += h;
    ix.khash.init(nk);
    KHash& H = ix.khash;
    parallel_ranges(n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            if (cls[i] != 1) continue;
            size_t j = i + 1;
            while (j < n && cls[j] == 0) ++j;
            uint64_t key = mer_from_chars(s + ix.SA[i], k);
            uint64_t h = KHash::mix(key) & H.mask;
            for (;;) {
                uint64_t expect = KHash::EMPTY;
                if (__atomic_compare_exchange_n(&H.slots[h].key, &expect, key, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                    H.slots[h].begin = (OffsetT)i; H.slots[h].end = (OffsetT)j;
                    break;
                }
                if (expect == key) throw std::runtime_error("khash: duplicate k-mer interval");
                h = (h + 1) & H.mask;
            }
        }
    });
    H.size = nk;
}

void finish_index(Index& ix, int nthreads) {
    ix.rankDict.build(ix.seq);
    build_sa_sais(ix.seq, ix.SA);
    build_khash(ix, nthreads);
}

// text already normalised and '$'-joined: offsets/lengths recovered from the '$' positions
void index_from_text(Index& ix, const char* seq, uint64_t n, int k, int nthreads) {
    ix.k = k;
    ix.seq.assign(seq, n);
    if (n == 0 || ix.seq.back() != '$') throw std::runtime_error("text must end with '$'");
    uint32_t start = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (seq[i] == '$') { ix.txpOffsets.push_back(start); ix.txpLens.push_back((uint32_t)(i - start)); start = (uint32_t)(i + 1); }
    ix.txpNames.resize(ix.txpOffsets.size());
    for (size_t t = 0; t < ix.txpNames.size(); ++t) ix.txpNames[t] = "t" + std::to_string(t);
    finish_index(ix, nthreads);
}

// variant that takes an externally computed suffix array (bench.py at BASELINE config-2
// scale, where a CPU comparison sort would take minutes).  The array is NOT trusted: it is
// checked with the linear-time suffix-array criterion (permutation; first characters
// ordered; ties ordered by the rank of the next suffix) before use.
bool verify_sa(const std::string& seq, const std::vector<OffsetT>& SA, int nthreads) {
    const size_t n = seq.size();
    if (SA.size() != n) return false;
    std::vector<OffsetT> inv(n, -1);
    std::atomic<bool> ok{true};
    parallel_ranges(n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t i = b; i < e; ++i) {
            OffsetT p = SA[i];
            if (p < 0 || (size_t)p >= n) { ok = false; return; }
            inv[p] = (OffsetT)i;      // a value hit twice leaves another one unhit, caught below
        }
    });
    if (!ok) return false;
    parallel_ranges(n, nthreads, [&](size_t b, size_t e, int) {
        for (size_t p = b; p < e; ++p)
            if (inv[p] < 0 || (size_t)SA[inv[p]] != p) { ok = false; return; }
    });
    if (!ok) return false;
    parallel_ranges(n - 1, nthreads, [&](size_t b0, size_t e0, int) {
        for (size_t i = b0; i < e0; ++i) {
            OffsetT a = SA[i], b = SA[i + 1];
            unsigned char ca = seq[a], cb = seq[b];
            if (ca > cb) { ok = false; return; }
            if (ca == cb) {
                // suffix a+1 must sort before suffix b+1 (empty suffix sorts first)
                if ((size_t)a + 1 == n) continue;
                if ((size_t)b + 1 == n) { ok = false; return; }
                if (inv[a + 1] > inv[b + 1]) { ok = false; return; }
            }
        }
    });
    return ok;
}

void index_from_text_with_sa(Index& ix, const char* seq, uint64_t n, int k, const OffsetT* sa, int nthreads) {
    ix.k = k;
    ix.seq.assign(seq, n);
    if (n == 0 || ix.seq.back() != '$') throw std::runtime_error("text must end with '$'");
    uint32_t start = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (seq[i] == '$') { ix.txpOffsets.push_back(start); ix.txpLens.push_back((uint32_t)(i - start)); start = (uint32_t)(i + 1); }
    ix.txpNames.resize(ix.txpOffsets.size());
    for (size_t t = 0; t < ix.txpNames.size(); ++t) ix.txpNames[t] = "t" + std::to_string(t);
    ix.SA.assign(sa, sa + n);
    if (!verify_sa(ix.seq, ix.SA, nthreads)) throw std::runtime_error("imported suffix array failed verification");
    ix.rankDict.build(ix.seq);
    build_khash(ix, nthreads);
}

// FASTA records -> index text, per SURVEY B6 (name up to first whitespace, upper-case,
// non-ACGT -> base from a fixed-seed mt19937, trailing poly-A (>=10) clipped, records
// shorter than k dropped).  The random replacement cannot be matched to upstream's exact
// draw sequence; synthetic references contain only ACGT and no >=10-A tail.
void index_from_records(Index& ix, const std::vector<std::pair<std::string, std::string>>& recs, int k, int nthreads) {
    ix.k = k;
    std::mt19937 eng(271828);
    std::uniform_int_distribution<> dis(0, 3);
    const char bases[4] = {'A', 'C', 'G', 'T'};
    for (auto& r : recs) {
        std::string name = r.first.substr(0, r.first.find_first_of(" \t"));
        std::string sq = r.second;
        for (auto& c : sq) {
            c = (char)::toupper((unsigned char)c);
            if (mer_code(c) < 0) c = bases[dis(eng)];
        }
        const size_t polyAClip = 10;
        if (sq.size() > polyAClip && sq.compare(sq.size() - polyAClip, polyAClip, std::string(polyAClip, 'A')) == 0) {
            size_t e = sq.find_last_not_of("Aa");
            if (e == std::string::npos) sq.clear(); else sq.resize(e + 1);
        }
        if (sq.size() < (size_t)k) continue;
        ix.txpNames.push_back(name);
        ix.txpOffsets.push_back((uint32_t)ix.seq.size());
        ix.txpLens.push_back((uint32_t)sq.size());
        ix.seq += sq;
        ix.seq += '$';
    }
    finish_index(ix, nthreads);
}

} // namespace orc
