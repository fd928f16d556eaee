"""The all-bases short cut of the read-packing kernel (circall_b200/csrc/cg_pack_swar.h, the same source the CUDA kernel
includes) compiled for the host and tried on ALL 2^32 four-character words against the character-by-character definition
of the codes (SACollectorCircall.hpp:117-128 / SURVEY.md §8 A3: A 0, C 1, G 2, T 3, either case; anything else is not a base)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
HDR = os.path.join(HERE, "..", "circall_b200", "csrc")

SRC = r'''
#include <stdio.h>
#include "cg_pack_swar.h"
static int code_of(unsigned c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; } return -1; }
int main() {
    unsigned long long bad = 0;
    for (unsigned long long w = 0; w < (1ull << 32); ++w) {
        uint32_t c = (uint32_t)w, diff = 0;
        uint32_t m = cg_pack4_bases(c, diff);
        int ok = 1; uint32_t want = 0;
        for (int b = 0; b < 4; ++b) { int k = code_of((c >> (8 * b)) & 0xFF); if (k < 0) { ok = 0; break; } want |= (uint32_t)k << (2 * b); }
        if ((diff == 0) != ok || (ok && (m >> 24) != want)) ++bad;
    }
    printf("%llu\n", bad);
    return 0;
}
'''


def test_pack4_bases_exhaustive(tmp_path):
    c = tmp_path / "swar.cpp"
    c.write_text(SRC)
    exe = tmp_path / "swar"
    subprocess.check_call(["g++", "-O3", "-I", HDR, "-o", str(exe), str(c)])
    out = subprocess.check_output([str(exe)], timeout=300).decode().strip()
    assert out == "0"
