// mt_jump.hpp — host side of the mt19937 stream used for the reference-order sample bits
// (boost::random::mt19937 seeded with the worker rank, src_mpi/mpi_predicter.cpp:87).
// Produces the 624-word generator state that precedes a given round of 624 draws, so a
// device stream segment can start anywhere in the sequence.
#pragma once
#include <cstdint>

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

// state BEFORE the twist that yields draws 624*round .. 624*round+623
inline void mt19937_state_at_round(uint32_t seed, uint64_t round, uint32_t* out) {
    mt19937_seed(seed, out);
    for (uint64_t r = 0; r < round; r++) mt19937_twist(out);
}

}  // namespace pdsat
