"""Host side of the mt19937 jump-ahead (pdsat_b200/csrc/mt_jump.hpp), no GPU: the polynomials the
radix-8 jump tree asks for - x^(624*r) mod phi for r = d * 2^e, d = 1..7 - applied to a state as the
correlation the device evaluates (out[w] = XOR over the set bits i of x[i + w]) must give the state
r rounds later; for distances too long to step, jumps must compose (7 = 4 + 2 + 1)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <vector>
#include "mt_jump.hpp"
using namespace pdsat;

// the device's evaluation of a jump, word for word (kernels.cuh: mt19937_seq_kernel + mt19937_corr_kernel)
static std::vector<uint32_t> jump_apply(const uint32_t* poly, const std::vector<uint32_t>& st) {
    std::vector<uint32_t> x(624 * 34);
    for (int i = 0; i < 624; i++) x[i] = st[i];
    for (size_t n = 0; n + 624 < x.size(); n++) {
        const uint32_t y = (x[n] & 0x80000000u) | (x[n + 1] & 0x7fffffffu);
        x[n + 624] = x[n + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    std::vector<uint32_t> out(624, 0);
    for (int i = 0; i < 19937; i++)
        if ((poly[i >> 5] >> (i & 31)) & 1u)
            for (int w = 0; w < 624; w++) out[w] ^= x[i + w];
    // the generator's state is 19937 bits: the low 31 bits of word 0 are not part of it (the
    // recurrence reads only the top bit of x[n]); a jump leaves an arbitrary combination there
    out[0] &= 0x80000000u;
    return out;
}

static const uint32_t* poly_rounds(uint64_t r) {
    if (__builtin_popcountll(r) == 1) return MtJumpPolys::instance().poly(__builtin_ctzll(r));
    return MtJumpPolys::instance().poly_for_rounds(r);
}

int main() {
    std::vector<uint32_t> s0(624);
    mt19937_seed(5489u, s0.data());
    int bad = 0;
    // (a) against stepping the generator
    const uint64_t rs[] = {1, 2, 3, 5, 6, 7, 8, 24, 40, 56, 3 * 64, 7 * 8, 5 * 16, 320};
    for (uint64_t r : rs) {
        std::vector<uint32_t> want = s0;
        for (uint64_t k = 0; k < r; k++) mt19937_twist(want.data());
        want[0] &= 0x80000000u;
        if (jump_apply(poly_rounds(r), s0) != want) { printf("MISMATCH rounds %llu\n", (unsigned long long)r); bad++; }
    }
    // (b) composition where stepping is out of reach: d * 2^e for the exponents a 10^7 x 30 fill uses
    for (int e : {10, 13, 16, 19, 31}) {
        const std::vector<uint32_t> a1 = jump_apply(poly_rounds(1ull << e), s0);
        const std::vector<uint32_t> a2 = jump_apply(poly_rounds(2ull << e), s0);
        const std::vector<uint32_t> a3 = jump_apply(poly_rounds(3ull << e), s0);
        const std::vector<uint32_t> a4 = jump_apply(poly_rounds(4ull << e), s0);
        const std::vector<uint32_t> a5 = jump_apply(poly_rounds(5ull << e), s0);
        const std::vector<uint32_t> a6 = jump_apply(poly_rounds(6ull << e), s0);
        const std::vector<uint32_t> a7 = jump_apply(poly_rounds(7ull << e), s0);
        const uint32_t* p1 = poly_rounds(1ull << e);
        if (jump_apply(p1, a1) != a2) { printf("MISMATCH 1+1 e=%d\n", e); bad++; }
        if (jump_apply(p1, a2) != a3) { printf("MISMATCH 2+1 e=%d\n", e); bad++; }
        if (jump_apply(p1, a3) != a4) { printf("MISMATCH 3+1 e=%d\n", e); bad++; }
        if (jump_apply(p1, a4) != a5) { printf("MISMATCH 4+1 e=%d\n", e); bad++; }
        if (jump_apply(p1, a5) != a6) { printf("MISMATCH 5+1 e=%d\n", e); bad++; }
        if (jump_apply(p1, a6) != a7) { printf("MISMATCH 6+1 e=%d\n", e); bad++; }
        if (jump_apply(poly_rounds(8ull << e), s0) != jump_apply(p1, a7)) { printf("MISMATCH 7+1 e=%d\n", e); bad++; }
    }
    printf(bad ? "FAILED %d\n" : "OK %d\n", bad);
    return bad ? 1 : 0;
}
"""


def test_jump_polynomials_on_the_host(tmp_path):
    src = tmp_path / "jump_check.cpp"
    src.write_text(SRC)
    exe = tmp_path / "jump_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "pdsat_b200", "csrc"),
                           "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK 0"), out.stdout + out.stderr
