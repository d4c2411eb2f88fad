// rng_chacha12.h -- restatement of rand 0.8 `StdRng::seed_from_u64(seed)` + `gen::<f64>()`, the only RNG
// use on the reference's path (viterbi_solver/utils.rs:101,170; Cargo.toml:18 rand = "0.8.4").
//
// StdRng (rand 0.8) = rand_chacha::ChaCha12Rng.  rand_core's default seed_from_u64 expands the u64 with a
// PCG32 stream into the 32-byte key; ChaCha12 runs with a 64-bit block counter (words 12,13) and stream id 0
// (words 14,15); BlockRng hands out the 16-word blocks in order (rand_chacha buffers four blocks, which does
// not change the order); next_u64 = two consecutive words, low word first; Standard f64 =
// (next_u64 >> 11) * 2^-53.  No Rust toolchain exists here; the Python twin of this file (superseq.py StdRng) is
// pinned to the value-stability vectors the rand / rand_chacha / rand_pcg crates publish in their own tests
// (tests/test_cli.py::test_stdrng_matches_published_rand_vectors), and the CLI tests check that this file
// produces the same active masks as the Python twin for 0 < prop < 1.
#pragma once
#include <cstdint>

struct StdRng {
    uint32_t key[8]; uint64_t counter = 0; uint32_t buf[16]; int idx = 16;
    static StdRng seed_from_u64(uint64_t state)
    {
        StdRng r;
        const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
        for (int i = 0; i < 8; i++) {
            state = state * MUL + INC;
            uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            r.key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
        }
        return r;
    }
    static inline uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
    void block()
    {
        uint32_t in[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5],
                           key[6], key[7], (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
        uint32_t x[16];
        for (int i = 0; i < 16; i++) x[i] = in[i];
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 16);
            x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl(x[b], 12);
            x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 8);
            x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl(x[b], 7);
        };
        for (int r = 0; r < 6; r++) {   // 12 rounds
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) buf[i] = x[i] + in[i];
        counter++; idx = 0;
    }
    uint64_t next_u64()
    {
        if (idx >= 16) block();
        uint64_t lo = buf[idx], hi = buf[idx + 1];
        idx += 2;
        return (hi << 32) | lo;
    }
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
};
