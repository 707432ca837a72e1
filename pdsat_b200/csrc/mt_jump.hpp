// mt_jump.hpp — host side of the mt19937 stream used for the reference-order sample bits
// (boost::random::mt19937 seeded with the worker rank, src_mpi/mpi_predicter.cpp:87;
// consumed one draw per bit by Addit_func::bool_rand, src_common/addit_func.cpp:193-196).
//
// mt19937 is a sequential generator; to let thousands of CTAs emit one stream in parallel
// the device needs the generator state at many positions of the sequence.  This file holds
// the GF(2) polynomial arithmetic for the standard jump-ahead (Haramoto, Matsumoto,
// Nishimura, Panneton, L'Ecuyer: "Efficient jump ahead for F2-linear random number
// generators"):
//     state(n + J) = g_J(A) state(n),   g_J(x) = x^J mod phi(x),
// phi = characteristic polynomial of the generator (degree 19937), A = one-word step.
// Applied to the 624-word window this is   out[w] = XOR_{i : g_J[i] = 1} x[n + i + w],
// which the device evaluates (kernels.cuh: mt19937_jump_kernel).  Only the polynomials
// x^(624 * 2^t) mod phi are ever needed (segment lengths are powers of two rounds); they
// come from one chain of squarings, computed lazily and cached for the process.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace pdsat {

inline void mt19937_seed(uint32_t seed, uint32_t* mt) {
    mt[0] = seed;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

inline void mt19937_twist(uint32_t* mt) {
    for (int k = 0; k < 624; k++) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
}

class MtJumpPolys {
public:
    static constexpr int DEG = 19937;
    static constexpr int W64 = (DEG + 63) / 64 + 1;  // 313 words hold degree <= 19937 + slack
    static constexpr int POLY_WORDS32 = 624;          // device format: 624 x 32 bits (bits >= DEG are zero)

    static MtJumpPolys& instance() {
        static MtJumpPolys inst;
        return inst;
    }

    // x^(624 * 2^t) mod phi as 624 little-endian 32-bit words
    const uint32_t* poly(int t) {
        std::lock_guard<std::mutex> lk(mu_);
        ensure_phi();
        while ((int)chain_.size() <= t) {
            if (chain_.empty()) {
                std::vector<uint64_t> p(W64, 0);
                p[624 / 64] |= 1ull << (624 % 64);  // x^624, already reduced
                chain_.push_back(p);
            } else {
                chain_.push_back(square_mod(chain_.back()));
            }
            std::vector<uint32_t> d(POLY_WORDS32, 0);
            const std::vector<uint64_t>& p = chain_.back();
            for (int i = 0; i < POLY_WORDS32; i++) d[i] = (uint32_t)(p[i >> 1] >> ((i & 1) * 32));
            dev_format_.push_back(d);
        }
        return dev_format_[t].data();
    }

    // x^(624 * rounds) mod phi for an arbitrary round count: product of the chain polynomials
    // of the set bits (one modular multiplication each, a few ms); cached per round count
    const uint32_t* poly_for_rounds(uint64_t rounds) {
        int top = 63;
        while (top > 0 && !((rounds >> top) & 1ull)) top--;
        poly(top);  // make sure the chain is long enough (takes and releases the lock)
        std::lock_guard<std::mutex> lk(mu_);
        auto it = composite_.find(rounds);
        if (it != composite_.end()) return it->second.data();
        std::vector<uint64_t> acc;
        for (int t = 0; t <= top; t++) {
            if (!((rounds >> t) & 1ull)) continue;
            if (acc.empty()) acc = chain_[t];
            else acc = mul_mod(acc, chain_[t]);
        }
        std::vector<uint32_t> d(POLY_WORDS32, 0);
        for (int i = 0; i < POLY_WORDS32; i++) d[i] = (uint32_t)(acc[i >> 1] >> ((i & 1) * 32));
        return composite_.emplace(rounds, std::move(d)).first->second.data();
    }

    int degree_of_phi() {
        std::lock_guard<std::mutex> lk(mu_);
        ensure_phi();
        return phi_deg_;
    }

private:
    std::mutex mu_;
    std::vector<uint64_t> phi_;                     // bit i = coefficient of x^i
    int phi_deg_ = -1;
    std::vector<std::vector<uint64_t>> phi_shift_;  // phi << s for s = 0..63 (W64 + 1 words)
    std::vector<std::vector<uint64_t>> chain_;
    std::vector<std::vector<uint32_t>> dev_format_;
    std::map<uint64_t, std::vector<uint32_t>> composite_;

    // Berlekamp-Massey over GF(2) on one output bit of the generator: the minimal polynomial
    // of that bit sequence is the characteristic polynomial (it is primitive of degree 19937).
    void ensure_phi() {
        if (phi_deg_ >= 0) return;
        const int N = 2 * DEG + 64;
        std::vector<uint8_t> s(N);
        {
            uint32_t mt[624];
            mt19937_seed(4357u, mt);
            int produced = 0;
            while (produced < N) {
                mt19937_twist(mt);
                for (int i = 0; i < 624 && produced < N; i++) s[produced++] = (uint8_t)(mt[i] & 1u);
            }
        }
        const int WN = (N + 63) / 64 + 2;
        // C, B: connection polynomials, bit i = coefficient c_i; srev: reversed sequence so
        // that the discrepancy is a parity of an AND of two aligned bit ranges.
        std::vector<uint64_t> C(WN, 0), B(WN, 0), T(WN, 0);
        C[0] = B[0] = 1;
        int L = 0, m = 1;
        // rev[j] = s[N-1-j]; window for step n: s[n-i], i = 0..L  ->  rev[(N-1-n) + i]
        std::vector<uint64_t> rev(WN + 1, 0);
        for (int j = 0; j < N; j++)
            if (s[N - 1 - j]) rev[j >> 6] |= 1ull << (j & 63);
        for (int n = 0; n < N; n++) {
            const int base = N - 1 - n;
            const int bw = base >> 6, bs = base & 63;
            uint64_t acc = 0;
            const int lw = L / 64 + 1;
            for (int w = 0; w < lw; w++) {
                uint64_t win = rev[bw + w] >> bs;
                if (bs) win |= rev[bw + w + 1] << (64 - bs);
                acc ^= win & C[w];
            }
            const int d = __builtin_parityll(acc);
            if (d) {
                const int mw = m >> 6, ms = m & 63;
                const bool grow = 2 * L <= n;
                if (grow) T = C;
                const int bwords = (n - L + 1 + m) / 64 + 2;  // degree of B is <= n - L... generous bound
                for (int w = 0; w < bwords && w + mw < WN; w++) {
                    uint64_t v = B[w];
                    if (!v) continue;
                    C[w + mw] ^= v << ms;
                    if (ms && w + mw + 1 < WN) C[w + mw + 1] ^= v >> (64 - ms);
                }
                if (grow) {
                    L = n + 1 - L;
                    B = T;
                    m = 1;
                } else {
                    m++;
                }
            } else {
                m++;
            }
        }
        // C(x) = sum c_i x^i is the connection polynomial: s_n = sum_{i>=1} c_i s_{n-i};
        // the characteristic polynomial is its reciprocal x^L C(1/x).
        phi_deg_ = L;
        phi_.assign(W64 + 1, 0);
        for (int i = 0; i <= L; i++)
            if ((C[i >> 6] >> (i & 63)) & 1ull) {
                int j = L - i;
                phi_[j >> 6] |= 1ull << (j & 63);
            }
        phi_shift_.assign(64, std::vector<uint64_t>(W64 + 2, 0));
        for (int sft = 0; sft < 64; sft++)
            for (int w = 0; w < W64 + 1; w++) {
                phi_shift_[sft][w] |= phi_[w] << sft;
                if (sft) phi_shift_[sft][w + 1] |= phi_[w] >> (64 - sft);
            }
    }

    // (p(x))^2 mod phi: squaring over GF(2) spreads the bits, then reduce from the top
    std::vector<uint64_t> square_mod(const std::vector<uint64_t>& p) {
        const int W2 = 2 * W64 + 4;
        std::vector<uint64_t> sq(W2, 0);
        auto spread = [](uint32_t x) -> uint64_t {
            uint64_t v = x;
            v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
            v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
            v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
            v = (v | (v << 2)) & 0x3333333333333333ull;
            v = (v | (v << 1)) & 0x5555555555555555ull;
            return v;
        };
        for (int w = 0; w < W64; w++) {
            sq[2 * w] = spread((uint32_t)p[w]);
            sq[2 * w + 1] = spread((uint32_t)(p[w] >> 32));
        }
        return reduce(sq);
    }

    // a(x) * b(x) mod phi (both already reduced): shift-and-xor product, then reduction
    std::vector<uint64_t> mul_mod(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
        const int W2 = 2 * W64 + 4;
        std::vector<uint64_t> pr(W2, 0);
        for (int i = 0; i < DEG; i++) {
            if (!((a[i >> 6] >> (i & 63)) & 1ull)) continue;
            const int sw = i >> 6, sh = i & 63;
            for (int w = 0; w < W64; w++) {
                const uint64_t v = b[w];
                if (!v) continue;
                pr[sw + w] ^= v << sh;
                if (sh) pr[sw + w + 1] ^= v >> (64 - sh);
            }
        }
        return reduce(pr);
    }

    // reduce a polynomial of degree < 2*DEG modulo phi, from the top bit down
    std::vector<uint64_t> reduce(std::vector<uint64_t>& sq) {
        for (int i = 2 * (DEG - 1); i >= DEG; i--) {
            if (!((sq[i >> 6] >> (i & 63)) & 1ull)) continue;
            const int sh = i - DEG;  // xor phi << sh
            const int sw = sh >> 6;
            const std::vector<uint64_t>& ph = phi_shift_[sh & 63];
            for (int w = 0; w < W64 + 2; w++) sq[sw + w] ^= ph[w];
        }
        std::vector<uint64_t> r(W64, 0);
        for (int w = 0; w < W64; w++) r[w] = sq[w];
        // clear anything at or above DEG (must already be zero)
        return r;
    }
};

}  // namespace pdsat
