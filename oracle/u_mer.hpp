// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
// [U] Jellyfish 2.2.5 mer_dna as used through rapmap::utils::my_mer.
// Call sites: SACollectorCircall.hpp:126-132,167,259-268,293,369-372,406,462,470.
#pragma once
#include <cstdint>

namespace orc {

// codes A=0 C=1 G=2 T=3 (either case); anything else is "not DNA"
static inline int mer_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

// my_mer(const char*): first base most significant; from_chars stops at the first
// non-DNA character and leaves the remaining low bits zero (no throw) — SURVEY B5,
// "oracle-defined" for N-containing windows.
static inline uint64_t mer_from_chars(const char* s, int k) {
    uint64_t w = 0;
    for (int j = 0; j < k; ++j) {
        int c = mer_code(s[j]);
        if (c < 0) break;
        w |= (uint64_t)c << (2 * (k - 1 - j));
    }
    return w;
}

static inline uint64_t mer_revcomp(uint64_t w, int k) {
    uint64_t r = 0;
    for (int j = 0; j < k; ++j) {
        uint64_t c = (w >> (2 * j)) & 3;
        r |= (3 - c) << (2 * (k - 1 - j));
    }
    return r;
}

static inline bool mer_is_homopolymer(uint64_t w, int k) {
    uint64_t c = w & 3;
    for (int j = 1; j < k; ++j)
        if (((w >> (2 * j)) & 3) != c) return false;
    return true;
}

// mer_dna::complement(char): ACGT (either case) -> upper-case complement; other -> 'N'
static inline char mer_complement(char c) {
    switch (c) {
        case 'A': case 'a': return 'T';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        case 'T': case 't': return 'A';
        default: return 'N';
    }
}

} // namespace orc
