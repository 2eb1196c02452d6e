// fast_host_check.cpp — CPU check of simulate_2048_b200/csrc/g2048_fast.cuh against the oracle.
// Build & run (tests/test_fast_path_host.py does this):
//   g++ -O2 -std=c++17 -o /tmp/fast_host_check tests/fast_host_check.cpp oracle/libg2048_oracle.so
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../simulate_2048_b200/csrc/g2048_fast.cuh"

extern "C" {
void orc_row_table(uint16_t* result, uint32_t* reward, uint8_t* moved);
void orc_step(const uint64_t* boards_in, const uint8_t* actions, const uint8_t* spawns, uint64_t seed,
              uint64_t env_offset, const uint32_t* step_ctr, uint32_t step_base, const uint64_t* env_keys,
              const uint8_t* done_in, uint64_t* boards_out, int32_t* rewards, uint8_t* flags, size_t n);
uint32_t orc_philox_word(uint64_t key, uint64_t env, uint32_t step, uint32_t tag);
}

using namespace g2048::fast;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return rng_state;
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 2000000;
    std::vector<uint16_t> res(65536); std::vector<uint32_t> rew(65536); std::vector<uint8_t> mv(65536);
    orc_row_table(res.data(), rew.data(), mv.data());
    std::vector<uint32_t> table(65536);
    for (int r = 0; r < 65536; ++r) table[hash_row16(r)] = (rew[r] >> 2) | ((uint32_t)res[r] << 16);
    auto lut = [&](u32 off) { return table[off >> 2]; };
    const Orient orient[4] = G2048_ORIENT_TABLE;

    std::vector<uint64_t> boards(n), out(n); std::vector<uint8_t> act(n), spn(n), fl(n); std::vector<int32_t> rw(n);
    for (size_t i = 0; i < n; ++i) {
        uint64_t b = 0;
        const int max_exp = (i % 3 == 0) ? 3 : 13;           // small exponents force merges / full boards
        const int fill = 1 + (int)(rnd() % 16);
        for (int c = 0; c < 16; ++c)
            if ((int)(rnd() % 16) < fill) b |= (uint64_t)(1 + rnd() % max_exp) << (4 * c);
        boards[i] = b;
        act[i] = (uint8_t)(rnd() & 3);
        spn[i] = (uint8_t)((rnd() & 0x1F));                  // rank 0..15 (maybe invalid), four bit
    }
    // a few hand-made edge boards
    boards[0] = 0; boards[1] = 0x1212212112122121ull; boards[2] = 0x1111111111111111ull;
    boards[3] = 0xDDDDDDDDDDDDDDDDull; boards[4] = 0x000000000000000Dull;

    // injected mode
    orc_step(boards.data(), act.data(), spn.data(), 0, 0, nullptr, 0, nullptr, nullptr, out.data(), rw.data(), fl.data(), n);
    size_t bad = 0, skipped = 0;
    for (size_t i = 0; i < n; ++i) {
        const u32 lo = (u32)boards[i], hi = (u32)(boards[i] >> 32);
        if (has_big_tile(lo, hi)) { ++skipped; continue; }
        FastOut f = transition_fast<false>(lo, hi, orient[act[i] & 3], spn[i] & 0xFu, (1u + ((spn[i] >> 4) & 1u)) << 29, 0, lut, 1u);
        const uint64_t got = ((uint64_t)f.hi << 32) | f.lo;
        if (got != out[i] || (int32_t)f.reward != rw[i] || (uint8_t)f.flags != fl[i]) {
            if (bad < 10)
                printf("MISMATCH inj i=%zu board=%016llx a=%u s=%02x: got %016llx r=%u f=%02x want %016llx r=%d f=%02x\n", i,
                       (unsigned long long)boards[i], act[i], spn[i], (unsigned long long)got, f.reward, f.flags,
                       (unsigned long long)out[i], rw[i], fl[i]);
            ++bad;
        }
    }
    // philox mode
    const uint64_t seed = 0x123456789ABCDEFull, off = 1ull << 40;
    orc_step(boards.data(), act.data(), nullptr, seed, off, nullptr, 7, nullptr, nullptr, out.data(), rw.data(), fl.data(), n);
    for (size_t i = 0; i < n; ++i) {
        const u32 lo = (u32)boards[i], hi = (u32)(boards[i] >> 32);
        if (has_big_tile(lo, hi)) continue;
        const uint32_t w = orc_philox_word(seed, off + i, 7, 0);
        FastOut f = transition_fast<true>(lo, hi, orient[act[i] & 3], 0, 0, w, lut, 1u);
        const uint64_t got = ((uint64_t)f.hi << 32) | f.lo;
        if (got != out[i] || (int32_t)f.reward != rw[i] || (uint8_t)f.flags != fl[i]) {
            if (bad < 20)
                printf("MISMATCH phx i=%zu board=%016llx a=%u: got %016llx r=%u f=%02x want %016llx r=%d f=%02x\n", i,
                       (unsigned long long)boards[i], act[i], (unsigned long long)got, f.reward, f.flags,
                       (unsigned long long)out[i], rw[i], fl[i]);
            ++bad;
        }
    }
    // has_big_tile is exact
    for (size_t i = 0; i < n; ++i) {
        bool big = false;
        for (int c = 0; c < 16; ++c) big |= ((boards[i] >> (4 * c)) & 0xF) >= 14;
        uint64_t b2 = boards[i] | ((rnd() % 7 == 0) ? (uint64_t)(14 + rnd() % 2) << (4 * (rnd() % 16)) : 0);
        bool big2 = false;
        for (int c = 0; c < 16; ++c) big2 |= ((b2 >> (4 * c)) & 0xF) >= 14;
        if (has_big_tile((u32)boards[i], (u32)(boards[i] >> 32)) != big || has_big_tile((u32)b2, (u32)(b2 >> 32)) != big2) ++bad;
    }
    printf("checked %zu boards x 2 modes (%zu skipped as big), mismatches: %zu\n", n, skipped, bad);
    return bad ? 1 : 0;
}
