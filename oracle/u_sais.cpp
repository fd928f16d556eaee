// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [U] libdivsufsort's role in RapMap's indexer (src/CMakeLists.txt:137-138): the suffix array of the
// '$'-joined transcript text.  The array is unique (all suffixes differ), so any correct suffix sorter returns what
// divsufsort returns; this one is induced sorting (SA-IS, Nong / Zhang / Chan 2009), linear in the text length whatever
// its repetitiveness — the isoform families of the hg19-scale configurations make comparison sorts take many minutes.
// It shares nothing with the product's GPU builder (prefix doubling over radix sorts), so the two check each other
// (tests/test_oracle_units.py on the CPU, tests/test_gpu_pipeline.py::test_index_matches_oracle on the GPU) and
// bench.py --impl reference builds its indices without loading the product library.
#include "oracle_types.hpp"
#include <stdexcept>
#include <cstring>

namespace orc {
namespace {

typedef int32_t I;

struct Bits {
    std::vector<uint64_t> w;
    explicit Bits(size_t n) : w((n + 63) / 64, 0) {}
    inline bool get(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    inline void set(size_t i, bool v) { if (v) w[i >> 6] |= 1ULL << (i & 63); else w[i >> 6] &= ~(1ULL << (i & 63)); }
};

template <class S>
void bucket_bounds(const S* s, I n, I K, std::vector<I>& bkt, bool ends) {
    std::fill(bkt.begin(), bkt.end(), 0);
    for (I i = 0; i < n; ++i) ++bkt[(size_t)s[i]];
    I sum = 0;
    for (I c = 0; c < K; ++c) { sum += bkt[c]; bkt[c] = ends ? sum : sum - bkt[c]; }
}

// L-type suffixes from the sorted suffixes to their left in the array (left-to-right scan)
template <class S>
void induce_l(const S* s, I* SA, I n, I K, const Bits& t, std::vector<I>& bkt) {
    bucket_bounds(s, n, K, bkt, false);
    for (I i = 0; i < n; ++i) {
        const I j = SA[i] - 1;
        if (SA[i] > 0 && !t.get((size_t)j)) SA[bkt[(size_t)s[j]]++] = j;
    }
}
// S-type suffixes (right-to-left scan)
template <class S>
void induce_s(const S* s, I* SA, I n, I K, const Bits& t, std::vector<I>& bkt) {
    bucket_bounds(s, n, K, bkt, true);
    for (I i = n - 1; i >= 0; --i) {
        const I j = SA[i] - 1;
        if (SA[i] > 0 && t.get((size_t)j)) SA[--bkt[(size_t)s[j]]] = j;
    }
}

// s[0..n): symbols in [0, K); s[n-1] is the unique smallest symbol.  SA has room for n entries.
template <class S>
void sais(const S* s, I* SA, I n, I K) {
    if (n == 1) { SA[0] = 0; return; }
    Bits t((size_t)n);                      // 1 = S-type
    t.set((size_t)n - 1, true);
    t.set((size_t)n - 2, false);
    for (I i = n - 3; i >= 0; --i)
        t.set((size_t)i, s[i] < s[i + 1] || (s[i] == s[i + 1] && t.get((size_t)i + 1)));
    auto lms = [&](I i) { return i > 0 && t.get((size_t)i) && !t.get((size_t)i - 1); };
    std::vector<I> bkt((size_t)K);

    // stage 1: sort the LMS substrings by one round of induced sorting
    bucket_bounds(s, n, K, bkt, true);
    for (I i = 0; i < n; ++i) SA[i] = -1;
    for (I i = 1; i < n; ++i) if (lms(i)) SA[--bkt[(size_t)s[i]]] = i;
    induce_l(s, SA, n, K, t, bkt);
    induce_s(s, SA, n, K, t, bkt);
    I n1 = 0;
    for (I i = 0; i < n; ++i) if (lms(SA[i])) SA[n1++] = SA[i];
    for (I i = n1; i < n; ++i) SA[i] = -1;
    // name them (equal substrings get equal names)
    I name = 0, prev = -1;
    for (I i = 0; i < n1; ++i) {
        const I pos = SA[i];
        bool diff = false;
        for (I d = 0; d < n; ++d) {
            if (prev == -1 || s[pos + d] != s[prev + d] || t.get((size_t)(pos + d)) != t.get((size_t)(prev + d))) { diff = true; break; }
            if (d > 0 && (lms(pos + d) || lms(prev + d))) break;
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (I i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];
    I* s1 = SA + n - n1;
    I* SA1 = SA;

    // stage 2: the reduced problem
    if (name < n1) sais<I>(s1, SA1, n1, name);
    else for (I i = 0; i < n1; ++i) SA1[s1[i]] = i;

    // stage 3: induce the whole array from the sorted LMS suffixes
    bucket_bounds(s, n, K, bkt, true);
    for (I i = 1, j = 0; i < n; ++i) if (lms(i)) s1[j++] = i;
    for (I i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    for (I i = n1; i < n; ++i) SA[i] = -1;
    for (I i = n1 - 1; i >= 0; --i) { const I j = SA[i]; SA[i] = -1; SA[--bkt[(size_t)s[j]]] = j; }
    induce_l(s, SA, n, K, t, bkt);
    induce_s(s, SA, n, K, t, bkt);
}

} // namespace

// suffix array of seq over {'$' < A < C < G < T}; a proper prefix sorts first (a virtual terminator below '$')
void build_sa_sais(const std::string& seq, std::vector<OffsetT>& SA) {
    const size_t n = seq.size();
    if (n + 1 >= (1ULL << 31)) throw std::runtime_error("text >= 2^31: 64-bit index not supported by the oracle");
    std::vector<uint8_t> s(n + 1);
    for (size_t i = 0; i < n; ++i) {
        switch (seq[i]) {
            case '$': s[i] = 1; break; case 'A': s[i] = 2; break; case 'C': s[i] = 3; break; case 'G': s[i] = 4; break; case 'T': s[i] = 5; break;
            default: throw std::runtime_error("index text holds a character outside {A,C,G,T,$}");
        }
    }
    s[n] = 0;
    std::vector<I> full(n + 1);
    sais<uint8_t>(s.data(), full.data(), (I)(n + 1), 6);
    std::vector<uint8_t>().swap(s);
    if (full[0] != (I)n) throw std::runtime_error("sais: terminator not first");
    SA.resize(n);
    std::memcpy(SA.data(), full.data() + 1, n * sizeof(I));
}

} // namespace orc
