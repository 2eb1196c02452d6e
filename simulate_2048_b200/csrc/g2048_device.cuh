// g2048_device.cuh — device primitives for packed 2048 boards (sm_100a).
//
// A board is one u64: nibble i = log2(tile) of flat cell i = 4r + c (include/g2048.h).
// The GPU is a 32-bit machine, so every primitive works on the two halves
//   lo = rows 0,1   hi = rows 2,3
// and is written so that ptxas emits LOP3 / SHF / PRMT / IMAD and nothing wider.
//
// Reference semantics implemented here (paths relative to the reference repo):
//   row move        twentyfortyeight/core/gameboard.py:23-69 merge_column (+ :72-102 padding),
//                   reinforce/game/core.py:43-101 merge_row
//   direction map   gameboard.py:105-129 / core.py:129-170 (rot90(k=action) + slide left)
//   legal mask      twentyfortyeight/core/gamemove.py:45-83, core.py:275-302
//   spawn           gameboard.py:247-288 fill_cells, core.py:173-231 (rank-th empty, row-major)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/g2048.h"

namespace g2048 {

typedef unsigned int u32;
typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// Row table entry (u32): bits 0..15 = (merge score) >> 2, bits 16..31 = row after a left move.
// Score of a row is a multiple of 4 and at most 131072, so score >> 2 fits 16 bits exactly.
// The shared-memory copy holds rows [0, kSmemRows): every row whose column-3 tile is <= 8192.
// Rows outside it (a 16384 / 32768 tile in the last column of the oriented board) take the
// exact slow path through the full table in global memory.
// ---------------------------------------------------------------------------------------------
constexpr u32 kRows = 65536u;
constexpr u32 kSmemRows = 0xE000u;             // 57344 entries = 229376 bytes
constexpr u32 kSmemTableBytes = kSmemRows * 4u;

struct Board {
    u32 lo, hi;
};

__device__ __forceinline__ Board split(u64 b) { return Board{(u32)b, (u32)(b >> 32)}; }
__device__ __forceinline__ u64 join(Board b) { return ((u64)b.hi << 32) | b.lo; }

// Transpose the 4x4 nibble matrix: 2x2 nibble blocks inside each half (xor-swap), then the
// 2x2 block of bytes across the halves (two PRMT).
__device__ __forceinline__ Board transpose(Board b) {
    u32 t = (b.lo ^ (b.lo >> 12)) & 0x0000F0F0u;
    u32 lo = b.lo ^ t ^ (t << 12);
    t = (b.hi ^ (b.hi >> 12)) & 0x0000F0F0u;
    u32 hi = b.hi ^ t ^ (t << 12);
    return Board{__byte_perm(lo, hi, 0x6240), __byte_perm(lo, hi, 0x7351)};
}

// Reverse the four nibbles of every row (mirror left <-> right).
__device__ __forceinline__ u32 reverse_rows(u32 h) {
    h = ((h & 0x0F0F0F0Fu) << 4) | ((h >> 4) & 0x0F0F0F0Fu);
    return __byte_perm(h, 0u, 0x2301);
}
__device__ __forceinline__ Board reverse_rows(Board b) {
    return Board{reverse_rows(b.lo), reverse_rows(b.hi)};
}

// bit 3 of every nibble = (nibble != 0); exact per nibble, no carries cross nibbles.
__device__ __forceinline__ u32 nz8(u32 v) {
    return (((v & 0x77777777u) + 0x77777777u) | v) & 0x88888888u;
}

// ---------------------------------------------------------------------------------------------
// Exact scalar row move, used to build the table and on the rare slow path.
// Returns the entry format above; *overflow is set when a 15+15 merge happened.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline u32 row_left_exact(u32 row, u32* score_out, bool* overflow) {
    u32 cells[4];
    int k = 0;
    for (int j = 0; j < 4; ++j) {
        u32 e = (row >> (4 * j)) & 0xFu;
        if (e) cells[k++] = e;  // drop empties (gameboard.py:46)
    }
    u32 out = 0, score = 0;
    bool ovf = false;
    int m = 0, i = 0;
    while (i < k) {  // greedy left-to-right pairing, each tile merges once (gameboard.py:56-64)
        if (i + 1 < k && cells[i] == cells[i + 1]) {
            u32 e = cells[i] + 1;
            score += 1u << e;
            if (e > 15u) ovf = true;
            out |= (e & 0xFu) << (4 * m);
            i += 2;
        } else {
            out |= cells[i] << (4 * m);
            i += 1;
        }
        ++m;
    }
    if (score_out) *score_out = score;
    if (overflow) *overflow = ovf;
    return out;
}

__host__ __device__ inline u32 make_entry(u32 row) {
    u32 score;
    bool ovf;
    u32 res = row_left_exact(row, &score, &ovf);
    return (score >> 2) | (res << 16);
}

// ---------------------------------------------------------------------------------------------
// Row-table accessors of the exact (first-generation) path: `GlobalTable` gathers from the full
// table through L1/L2, `RegisterTable` computes the row.  The staged shared-memory copy is read
// through `SmemLut` (g2048_kernels.cu) by the fast transition of g2048_fast.cuh.
// ---------------------------------------------------------------------------------------------
struct GlobalTable {
    const u32* __restrict__ tab;
    __device__ __forceinline__ u32 operator()(u32 row) const { return __ldg(tab + row); }
    // every row is in range
    __device__ __forceinline__ bool covers(Board) const { return true; }
};

struct MoveResult {
    Board after;
    u32 score;      // sum of merged tile values
    bool overflow;  // exact
};

// Exact out-of-line move for boards with a 16384 / 32768 tile in the oriented last column or a
// 15+15 merge.  Kept out of line so the hot path stays small.
__device__ __noinline__ MoveResult move_left_slow(Board x) {
    MoveResult r;
    u32 res[4];
    r.score = 0;
    r.overflow = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const u32 h = (i < 2) ? x.lo : x.hi;
        const u32 row = (i & 1) ? (h >> 16) : (h & 0xFFFFu);
        u32 s;
        bool o;
        res[i] = row_left_exact(row, &s, &o);
        r.score += s;
        r.overflow |= o;
    }
    r.after.lo = res[0] | (res[1] << 16);
    r.after.hi = res[2] | (res[3] << 16);
    return r;
}

// Left move of all four rows of an oriented board.
template <class Table>
__device__ __forceinline__ MoveResult move_left(Board x, const Table& T) {
    MoveResult r;
    if (T.covers(x)) {
        u32 e0 = T(x.lo & 0xFFFFu), e1 = T(x.lo >> 16), e2 = T(x.hi & 0xFFFFu), e3 = T(x.hi >> 16);
        r.after.lo = __byte_perm(e0, e1, 0x7632);
        r.after.hi = __byte_perm(e2, e3, 0x7632);
        r.score = ((e0 & 0xFFFFu) + (e1 & 0xFFFFu) + (e2 & 0xFFFFu) + (e3 & 0xFFFFu)) << 2;
        r.overflow = false;
        if (r.score < 65536u) return r;  // a 15+15 merge scores 65536 on its own
    }
    // slow path: tiles >= 16384 present; exact arithmetic, never taken in ordinary play
    return move_left_slow(x);
}

// ---------------------------------------------------------------------------------------------
// Pure register-arithmetic row move (no table): the comparison variant the north star asks for
// (G2048_IMPL_REGISTER).  Branch-free per row: six conditional "close the gap" steps compact the
// four cells, three compare-and-merge steps apply the greedy left-to-right pairing
// (gameboard.py:56-64), two more steps close the holes the merges leave.  About 70 select /
// compare / shift instructions per row — see DESIGN.md for the measured comparison.
// ---------------------------------------------------------------------------------------------
struct RegisterTable {
    __device__ __forceinline__ bool covers(Board) const { return true; }
};

__device__ __forceinline__ u32 row_left_register(u32 row, u32& score, bool& overflow) {
    u32 a = row & 0xFu, b = (row >> 4) & 0xFu, c = (row >> 8) & 0xFu, d = (row >> 12) & 0xFu;
#define G2048_CLOSE0 { const bool z = (a == 0u); a = z ? b : a; b = z ? c : b; c = z ? d : c; d = z ? 0u : d; }
#define G2048_CLOSE1 { const bool z = (b == 0u); b = z ? c : b; c = z ? d : c; d = z ? 0u : d; }
#define G2048_CLOSE2 { const bool z = (c == 0u); c = z ? d : c; d = z ? 0u : d; }
    G2048_CLOSE0 G2048_CLOSE0 G2048_CLOSE0 G2048_CLOSE1 G2048_CLOSE1 G2048_CLOSE2
    const bool m0 = (a != 0u) & (a == b);
    const bool m1 = (b != 0u) & (b == c) & !m0;
    const bool m2 = (c != 0u) & (c == d) & !m1;
    score = (m0 ? (2u << a) : 0u) + (m1 ? (2u << b) : 0u) + (m2 ? (2u << c) : 0u);
    overflow = (m0 & (a == 15u)) | (m1 & (b == 15u)) | (m2 & (c == 15u));
    a = m0 ? (a + 1u) & 0xFu : a;
    if (m0) b = 0u;
    b = m1 ? (b + 1u) & 0xFu : b;
    if (m1) c = 0u;
    c = m2 ? (c + 1u) & 0xFu : c;
    if (m2) d = 0u;
    // holes are closed by POSITION, not by value: an overflowed 15+15 merge wraps to nibble 0
    // but keeps its place, exactly as the packed format defines (include/g2048.h)
    if (m0) { b = c; c = d; d = 0u; }       // hole at 1
    if (m1) { c = d; d = 0u; }              // hole at 2
    if (m0 & m2) { c = 0u; }                // second merge landed at 1, its partner hole at 2
    return a | (b << 4) | (c << 8) | (d << 12);
#undef G2048_CLOSE0
#undef G2048_CLOSE1
#undef G2048_CLOSE2
}

__device__ __forceinline__ MoveResult move_left(Board x, const RegisterTable&) {
    MoveResult r;
    u32 s0, s1, s2, s3;
    bool o0, o1, o2, o3;
    const u32 r0 = row_left_register(x.lo & 0xFFFFu, s0, o0), r1 = row_left_register(x.lo >> 16, s1, o1);
    const u32 r2 = row_left_register(x.hi & 0xFFFFu, s2, o2), r3 = row_left_register(x.hi >> 16, s3, o3);
    r.after.lo = r0 | (r1 << 16);
    r.after.hi = r2 | (r3 << 16);
    r.score = s0 + s1 + s2 + s3;
    r.overflow = o0 | o1 | o2 | o3;
    return r;
}

// Afterstate of `b` under `action` (0 left, 1 up, 2 right, 3 down): orient so that the move is
// a left move (transpose for vertical, mirror for right/down), move, undo.
template <class Table>
__device__ __forceinline__ MoveResult move_action(Board b, u32 action, const Table& T) {
    const bool vertical = action & 1u, mirrored = action & 2u;
    Board x = b;
    if (vertical) x = transpose(x);
    if (mirrored) x = reverse_rows(x);
    MoveResult r = move_left(x, T);
    if (mirrored) r.after = reverse_rows(r.after);
    if (vertical) r.after = transpose(r.after);
    return r;
}

// ---------------------------------------------------------------------------------------------
// Legal-action mask, pure register arithmetic (gamemove.py:45-83): bit a = action a is legal.
//   merge possible : a non-empty cell equals its right / lower neighbour
//   slide possible : an empty cell has a non-empty neighbour on the far side
// All tests run nibble-parallel on the bit-3 indicators produced by nz8().
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 legal_mask(Board b, u32 nlo, u32 nhi) {
    // horizontal: neighbour j+1 aligned on j; column 3 is masked out by M
    const u32 elo = nz8(b.lo ^ (b.lo >> 4)), ehi = nz8(b.hi ^ (b.hi >> 4));  // cells differ
    const u32 slo = nlo >> 4, shi = nhi >> 4;                                // neighbour non-empty
    const u32 M = 0x08880888u;
    const u32 L = (((nlo & ~elo) | (~nlo & slo)) | ((nhi & ~ehi) | (~nhi & shi))) & M;
    const u32 R = (((nlo & ~elo) | (nlo & ~slo)) | ((nhi & ~ehi) | (nhi & ~shi))) & M;
    // vertical: row r+1 aligned on row r = (hi:lo) >> 16; row 3 has no lower neighbour
    const u32 vlo = __byte_perm(b.lo, b.hi, 0x5432), vhi = b.hi >> 16;
    const u32 flo = nz8(b.lo ^ vlo), fhi = nz8(b.hi ^ vhi);
    const u32 tlo = __byte_perm(nlo, nhi, 0x5432), thi = nhi >> 16;
    const u32 U = ((nlo & ~flo) | (~nlo & tlo)) | (((nhi & ~fhi) | (~nhi & thi)) & 0x0000FFFFu);
    const u32 D = ((nlo & ~flo) | (nlo & ~tlo)) | (((nhi & ~fhi) | (nhi & ~thi)) & 0x0000FFFFu);
    return (L != 0u ? 1u : 0u) | (U != 0u ? 2u : 0u) | (R != 0u ? 4u : 0u) | (D != 0u ? 8u : 0u);
}
__device__ __forceinline__ u32 legal_mask(Board b) { return legal_mask(b, nz8(b.lo), nz8(b.hi)); }

// flags byte of a board: legal mask + DONE when nothing is legal (core.py:305-321)
__device__ __forceinline__ u32 board_flags(u32 mask) {
    return mask | (mask == 0u ? G2048_FLAG_DONE : 0u);
}

// ---------------------------------------------------------------------------------------------
// Spawn: place a tile of exponent `exp` (1 or 2) on the rank-th empty cell in row-major order.
// `zlo/zhi` carry bit 0 of every empty nibble.  Within a half, a multiply by 0x11111111 turns
// the indicators into inclusive prefix counts (<= 8, no carry), adding (7 - r) per nibble sets
// bit 3 exactly where the count exceeds r, and the lowest such nibble is the target.
// Returns false (board untouched) when rank >= number of empties.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void empties(Board b, u32& zlo, u32& zhi) {
    zlo = (~nz8(b.lo) >> 3) & 0x11111111u;
    zhi = (~nz8(b.hi) >> 3) & 0x11111111u;
}

__device__ __forceinline__ bool spawn_rank(Board& b, u32 zlo, u32 zhi, u32 rank, u32 exp) {
    const u32 clo = __popc(zlo), n = clo + __popc(zhi);
    if (rank >= n) return false;
    const bool in_hi = rank >= clo;
    const u32 z = in_hi ? zhi : zlo;
    const u32 r = in_hi ? rank - clo : rank;
    const u32 g = (z * 0x11111111u + (0x77777777u - r * 0x11111111u)) & 0x88888888u;
    const u32 tile = ((g & (0u - g)) >> 3) * exp;
    if (in_hi) b.hi |= tile; else b.lo |= tile;
    return true;
}

__device__ __forceinline__ bool spawn_cell(Board& b, u32 cell, u32 exp) {
    const u32 sh = (cell & 7u) * 4u;
    u32& h = (cell & 8u) ? b.hi : b.lo;
    if ((h >> sh) & 0xFu) return false;
    h |= exp << sh;
    return true;
}

// Injected spawn byte (include/g2048.h).  Returns false -> G2048_FLAG_BADSPAWN.
__device__ __forceinline__ bool spawn_byte(Board& b, u32 zlo, u32 zhi, u32 byte) {
    const u32 exp = (byte & 0x10u) ? 2u : 1u;
    if (byte & 0x80u) return spawn_cell(b, byte & 0xFu, exp);
    return spawn_rank(b, zlo, zhi, byte & 0xFu, exp);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), stream layout v2 (include/g2048.h "Philox mode"):
// ONE block serves the four consecutive envs of a quad — counter = (quad_lo, quad_hi, step, tag)
// with quad = global env id >> 2, key = (k0, k1) — and env e reads word e & 3.  A thread of the step
// kernel owns exactly one quad, so the ~42 instructions of a block are paid once per four envs.
// The word is consumed as a mixed-radix fraction: rank = hi32(w * n_empty), f1 = lo32(w * n_empty);
// tile 4 iff hi32(f1 * 10) == 0, f2 = lo32(f1 * 10); the rollout policy's next action is
// hi32(f2 * n_legal).  Each step is one IMAD.WIDE.
// ---------------------------------------------------------------------------------------------
struct Philox4 {
    u32 x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const u32 h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const u32 h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// Same generator with the ten round keys precomputed on the host (they depend only on the seed):
// as kernel parameters they are constant-bank operands of the xor and cost no instruction.
struct PhiloxKeys {
    u32 k0[10], k1[10];
};

inline PhiloxKeys philox_key_schedule(u64 key) {
    PhiloxKeys s;
    u32 a = (u32)key, b = (u32)(key >> 32);
    for (int r = 0; r < 10; ++r) {
        s.k0[r] = a;
        s.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return s;
}

// (each product is ONE 32 x 32 -> 64 multiply, mul.wide.u32 / IMAD.WIDE.U32: written as __umulhi
// plus a low multiply ptxas emits IMAD.HI + IMAD, six instructions per round instead of four)
__device__ __forceinline__ void philox_mul(u32 m, u32 c, u32& hi, u32& lo) {
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(m), "r"(c));
}

__device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, const PhiloxKeys& ks) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        u32 h0, l0, h1, l1;
        philox_mul(0xD2511F53u, c0, h0, l0);
        philox_mul(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ ks.k0[r];
        c1 = l1;
        c2 = h0 ^ c3 ^ ks.k1[r];
        c3 = l0;
    }
    return Philox4{c0, c1, c2, c3};
}

// the block of the quad that holds `env`
__device__ __forceinline__ Philox4 philox_quad(const PhiloxKeys& ks, u64 env, u32 step, u32 tag) {
    const u64 q = env >> 2;
    return philox4x32_10((u32)q, (u32)(q >> 32), step, tag, ks);
}
__device__ __forceinline__ Philox4 philox_quad(u64 key, u64 env, u32 step, u32 tag) {
    const u64 q = env >> 2;
    return philox4x32_10((u32)q, (u32)(q >> 32), step, tag, (u32)key, (u32)(key >> 32));
}
// word j (0..3) of a block
__device__ __forceinline__ u32 philox_pick(const Philox4& p, u32 j) {
    const u32 a = (j & 1u) ? p.y : p.x, b = (j & 1u) ? p.w : p.z;
    return (j & 2u) ? b : a;
}
// the word of one env
__device__ __forceinline__ u32 philox_word(u64 key, u64 env, u32 step, u32 tag) {
    return philox_pick(philox_quad(key, env, step, tag), (u32)env & 3u);
}
__device__ __forceinline__ u32 philox_word(const PhiloxKeys& ks, u64 env, u32 step, u32 tag) {
    return philox_pick(philox_quad(ks, env, step, tag), (u32)env & 3u);
}

// The GAME stream of one env (include/g2048.h "Game stream"; g2048_reset and g2048_rollout): block b =
// Philox4x32-10 of counter (env_lo, env_hi, b, tag 6), and word p of the stream is word p & 3 of
// block p >> 2 — four consecutive draws of ONE game cost one block.  (The quad layout above gives
// four envs at the SAME step one block: the right shape for the lock-step kernels, the wrong one
// for a rollout lane whose games are at different steps.)
__device__ __forceinline__ Philox4 philox_game_block(const PhiloxKeys& ks, u32 env_lo, u32 env_hi, u32 b) {
    return philox4x32_10(env_lo, env_hi, b, G2048_PHILOX_TAG_GAME, ks);
}
__device__ __forceinline__ Philox4 philox_game_block(u64 key, u64 env, u32 b) {
    return philox4x32_10((u32)env, (u32)(env >> 32), b, G2048_PHILOX_TAG_GAME, (u32)key, (u32)(key >> 32));
}

// digits 1 and 2 of a word: spawn rank in [0, n) and the spawned exponent (2 for a tile 4, p = 0.1);
// returns the remainder f2 (the next digit is the rollout policy's action)
__device__ __forceinline__ u32 draw_spawn(u32 w, u32 n, u32& rank, u32& exp) {
    const u64 p1 = (u64)w * n;
    const u64 p2 = (u64)(u32)p1 * 10u;
    rank = (u32)(p1 >> 32);
    exp = (u32)(p2 >> 32) == 0u ? 2u : 1u;
    return (u32)p2;
}

// Philox spawn on a board with at least one empty cell.  Returns the remainder f2.
__device__ __forceinline__ u32 spawn_philox(Board& b, u32 zlo, u32 zhi, u32 w) {
    const u32 n = __popc(zlo) + __popc(zhi);
    if (n == 0u) return w;
    u32 rank, exp;
    const u32 rem = draw_spawn(w, n, rank, exp);
    spawn_rank(b, zlo, zhi, rank, exp);
    return rem;
}

__device__ __forceinline__ u32 max_exponent(Board b) {
    u32 m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m = max(m, (b.lo >> (4 * i)) & 0xFu);
        m = max(m, (b.hi >> (4 * i)) & 0xFu);
    }
    return m;
}

}  // namespace g2048
