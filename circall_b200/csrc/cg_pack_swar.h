// circall_b200 — the all-bases short cut of the read-packing kernel (K1), four characters at a time.
// Plain integer arithmetic in a header of its own so that tests/test_pack_swar.py can compile it with gcc and try all
// 2^32 input words against the character-by-character definition (SACollectorCircall.hpp:117-128: A 0, C 1, G 2, T 3).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD static inline
#endif

// c: four ASCII characters.  Returns the four 2-bit codes side by side in the TOP byte (character 0 at bits 24-25) and
// ORs into `diff` a non-zero value unless all four characters are bases (ACGT in either case).
CG_HD uint32_t cg_pack4_bases(uint32_t c, uint32_t& diff) {
    uint32_t t2 = ((c >> 1) ^ (c >> 2)) & 0x03030303u;                       // A 0, C 1, G 2, T 3
    uint32_t hi = (t2 >> 1) & 0x01010101u;
    uint32_t e = 0x41414141u + 2u * t2 + 2u * hi + 11u * (t2 & hi);          // 'A' + {0, 2, 6, 19}: the upper-case base each code stands for
    diff |= (e ^ c) & 0xDFDFDFDFu;
    return t2 * 0x01041040u;
}
