// g2048_fast.cuh — the branch-free transition used by the HBM-bound step kernel.
//
// Measured on B200 (profiles/r01_microbench_pipes.txt): LOP3 / SHF / PRMT / ISETP / SEL issue at
// 2 warp-instr/clk/SM (the "ALU" pipe), IMAD at 2 on the separate FMA pipe, integer adds at 4,
// IMAD.HI at 1, POPC at 0.5, a random LDS.32 costs 3.5 clk.  ncu on the first version of the
// step kernel: ALU pipe 94 % busy, FMA pipe 10 %, 220 instructions per env (profiles/).  The
// HBM roofline leaves ~1 clk per env-step per SM, so the transition is written to
//   * never branch on data (actions are random: a divergent warp pays for every path),
//   * keep the ALU pipe for what only it can do (3-input logic, byte permutes, compares),
//   * put every shift on the FMA pipe (x << k is IMAD.SHL, x >> k is IMAD.HI by 2^(32-k)),
//     and use adds where the operands have disjoint bits.
//
// Orientation.  Up / right / down become "left" through a transform F_a and back through
// G_a = F_a^-1.  With  T = Bt.Nt  (Nt: swap nibbles 1<->4, 3<->6 inside each 32-bit half;
// Bt: 2x2 byte-block transpose across the halves) and  R = Br.Nr  (Nr: swap the nibbles of
// every byte; Br: swap the bytes of every halfword), Nr commutes with every byte permutation
// and Nt with Bt, so
//   F_a = P_a . Nr^[a&2] . Nt^[a&1]          G_a = Nt^[a&1] . Nr^[a&2] . Q_a
// with ONE byte permutation each (a PRMT pair whose selector is per-thread data) and the nibble
// stages made conditional by a mask that is zero when the stage is off.  The six per-action
// words (two masks, four selectors) come from a 4-entry shared-memory table: two conflict-free
// broadcast loads per env.
//
// Row-table entry (u32): bits 0..15 = (merge score of the row) >> 2, bits 16..31 = the row after
// a left move.  With the score in the LOW half the four entries of a board simply add up:
// on the fast path every row has nibbles <= 13, so a field is <= 8192 and the sum <= 32768.
//
// Every primitive has a host twin so the identical dataflow is unit-tested on the CPU
// (tests/test_fast_path_host.py builds tools/fast_host_check.cpp against the oracle).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define G2048_HD __host__ __device__ __forceinline__
#else
#define G2048_HD inline
#endif

namespace g2048 {
namespace fast {

typedef unsigned int u32;
typedef unsigned long long u64;

// ---- primitives with forced instruction selection on the device -----------------------------
constexpr u32 kA = 0xF0, kB = 0xCC, kC = 0xAA;  // lop3 truth-table generators

template <u32 kLut>
G2048_HD u32 lop3(u32 a, u32 b, u32 c) {
#if defined(__CUDA_ARCH__)
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(kLut));
    return r;
#else
    u32 r = 0;
    for (int m = 0; m < 8; ++m)
        if ((kLut >> m) & 1u)
            r |= ((m & 4) ? a : ~a) & ((m & 2) ? b : ~b) & ((m & 1) ? c : ~c);
    return r;
#endif
}

// byte permute without the selector masking __byte_perm adds (selector nibbles are 0..7 here)
G2048_HD u32 prmt(u32 a, u32 b, u32 sel) {
#if defined(__CUDA_ARCH__)
    u32 r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
#else
    u64 src = ((u64)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; ++i) r |= (u32)((src >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// x >> k on the FMA pipe: high word of x * 2^(32-k) (IMAD.HI)
template <int k>
G2048_HD u32 shr(u32 x) {
#if defined(__CUDA_ARCH__)
    return __umulhi(x, 1u << (32 - k));
#else
    return x >> k;
#endif
}

// a + b kept off the ALU pipe.  ptxas rewrites mad.lo(a, 1, b) into IADD3 / VIADD, which issue
// on the same 2/clk pipe as LOP3 (profiles/r01_microbench_pipes.txt), so the multiplier is a
// run-time 1 (`one`, a kernel parameter living in a uniform register): a genuine IMAD.
G2048_HD u32 add(u32 a, u32 b, u32 one) {
#if defined(__CUDA_ARCH__)
    u32 r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
#else
    return a * one + b;
#endif
}

// min(x, 1) at full issue rate (VIMNMX): any-bit-set -> 0/1 without a predicate round trip
G2048_HD u32 nonzero01(u32 x) {
#if defined(__CUDA_ARCH__)
    u32 r;
    asm("min.u32 %0, %1, 1;" : "=r"(r) : "r"(x));
    return r;
#else
    return x ? 1u : 0u;
#endif
}

// a * b + c on the FMA pipe (IMAD)
G2048_HD u32 mad(u32 a, u32 b, u32 c) {
#if defined(__CUDA_ARCH__)
    u32 r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a * b + c;
#endif
}

G2048_HD u32 popc(u32 x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return (u32)__builtin_popcount(x);
#endif
}
G2048_HD u32 mulhi(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}

// ---- per-action orientation words -----------------------------------------------------------
struct alignas(32) Orient {
    u32 mt, mr;        // nibble-stage masks (0 = stage off)
    u32 f_lo, f_hi;    // forward byte permutation (PRMT selectors)
    u32 b_lo, b_hi;    // backward byte permutation
    u32 pad0, pad1;    // 32-byte entries
};

// action 0 left, 1 up, 2 right, 3 down
#define G2048_ORIENT_TABLE                                                          \
    {                                                                               \
        {0x00000000u, 0x00000000u, 0x3210u, 0x7654u, 0x3210u, 0x7654u, 0u, 0u},     \
        {0x0000F0F0u, 0x00000000u, 0x6240u, 0x7351u, 0x6240u, 0x7351u, 0u, 0u},     \
        {0x00000000u, 0x0F0F0F0Fu, 0x2301u, 0x6745u, 0x2301u, 0x6745u, 0u, 0u},     \
        {0x0000F0F0u, 0x0F0F0F0Fu, 0x2604u, 0x3715u, 0x7351u, 0x6240u, 0u, 0u},     \
    }

// conditional nibble stages on one half: swap the nibble pairs selected by the mask
G2048_HD u32 stage_nt(u32 x, u32 mt) {
    const u32 t = lop3<(kA ^ kB) & kC>(x, shr<12>(x), mt);
    return lop3<kA ^ kB ^ kC>(x, t, t << 12);
}
G2048_HD u32 stage_nr(u32 x, u32 mr) {
    const u32 t = lop3<(kA ^ kB) & kC>(x, shr<4>(x), mr);
    return lop3<kA ^ kB ^ kC>(x, t, t << 4);
}

G2048_HD void orient_forward(u32& lo, u32& hi, const Orient& o) {
    const u32 l = stage_nr(stage_nt(lo, o.mt), o.mr), h = stage_nr(stage_nt(hi, o.mt), o.mr);
    lo = prmt(l, h, o.f_lo);
    hi = prmt(l, h, o.f_hi);
}
G2048_HD void orient_backward(u32& lo, u32& hi, const Orient& o) {
    const u32 l = prmt(lo, hi, o.b_lo), h = prmt(lo, hi, o.b_hi);
    lo = stage_nt(stage_nr(l, o.mr), o.mt);
    hi = stage_nt(stage_nr(h, o.mr), o.mt);
}

// bit 3 of every nibble = (nibble != 0); all other bits 0
G2048_HD u32 nz8(u32 v, u32 one) {
    return lop3<(kA | kB) & kC>(add(v & 0x77777777u, 0x77777777u, one), v, 0x88888888u);
}
// bit 3 of every nibble = (nibble of a != nibble of b); the OTHER bits are garbage.  Two LOP3
// instead of four: the xor is fused into both stages and the final mask is dropped, which is
// sound because every consumer ANDs bit-3-only indicator words into the result.
G2048_HD u32 differ8_dirty(u32 a, u32 b, u32 one) {
    const u32 t = add(lop3<(kA ^ kB) & kC>(a, b, 0x77777777u), 0x77777777u, one);
    return lop3<kA | (kB ^ kC)>(t, a, b);
}
// Bank-spreading row hash.  The staged table is indexed by h(row) = row ^ ((row >> 4) & 0x0FFF)
// (a bijection on 16 bits, applied to both rows of a 32-bit word at once).  The shared-memory
// bank of an entry is the low 5 bits of its index; un-hashed that is the column-0 nibble, which
// is 0 for about half of all cells in real boards, so distinct rows pile up on banks 0 and 16
// (ncu: 38.8 wavefronts per warp-step instead of ~20).  XOR-ing column 1 in spreads them.
G2048_HD u32 hash_rows(u32 w) { return lop3<kA ^ (kB & kC)>(w, shr<4>(w), 0x0FFF0FFFu); }
G2048_HD u32 hash_row16(u32 row) { return (row ^ ((row >> 4) & 0x0FFFu)) & 0xFFFFu; }

// true when some tile is >= 16384 (nibble >= 14): such boards leave the fast path, which keeps
// every row inside the staged table window and rules out a 15+15 merge.
G2048_HD u32 big_tile_bits(u32 lo, u32 hi) {
    const u32 a = lop3<kA & kB & kC>(lo, lo << 1, lo << 2), b = lop3<kA & kB & kC>(hi, hi << 1, hi << 2);
    return lop3<(kA | kB) & kC>(a, b, 0x88888888u);
}
G2048_HD bool has_big_tile(u32 lo, u32 hi) { return big_tile_bits(lo, hi) != 0u; }

// Legal mask of board (lo,hi) given its non-empty indicators (bit 3 per nibble).
// Every cell is compared with its LEFT / UPPER neighbour (left shifts: IMAD.SHL, FMA pipe).
//   left / up   : a tile whose neighbour is empty or equal            n & ~(e & s)
//   right / down: a tile equal to its neighbour, or an empty cell
//                 whose neighbour is a tile                           (n & ~e) | (~n & s)
// The shifted-in neighbour bits at column 0 / row 0 are garbage and masked out at the end.
G2048_HD u32 legal_mask(u32 lo, u32 hi, u32 nlo, u32 nhi, u32 one) {
    constexpr u32 kFwd = kA & ~(kB & kC) & 0xFF, kBwd = ((kA & ~kB) | (~kA & kC)) & 0xFF;
    // horizontal, valid at columns 1..3
    const u32 elo = differ8_dirty(lo, lo << 4, one), ehi = differ8_dirty(hi, hi << 4, one);  // differs from left neighbour
    const u32 slo = nlo << 4, shi = nhi << 4;                        // left neighbour non-empty
    const u32 L = lop3<(kA | kB) & kC>(lop3<kFwd>(nlo, elo, slo), lop3<kFwd>(nhi, ehi, shi), 0x88808880u);
    const u32 R = lop3<(kA | kB) & kC>(lop3<kBwd>(nlo, elo, slo), lop3<kBwd>(nhi, ehi, shi), 0x88808880u);
    // vertical, valid at rows 1..3: (hi:lo) << 16 puts row r-1 under row r
    const u32 vhi = prmt(lo, hi, 0x5432), thi = prmt(nlo, nhi, 0x5432);
    const u32 flo = differ8_dirty(lo, lo << 16, one), fhi = differ8_dirty(hi, vhi, one);
    const u32 tlo = nlo << 16;
    const u32 U = lop3<(kA & kB) | kC>(lop3<kFwd>(nlo, flo, tlo), 0xFFFF0000u, lop3<kFwd>(nhi, fhi, thi));
    const u32 D = lop3<(kA & kB) | kC>(lop3<kBwd>(nlo, flo, tlo), 0xFFFF0000u, lop3<kBwd>(nhi, fhi, thi));
    // any bit set -> 0/1 (min.u32, full rate), then l + 2u + 4r + 8d on the FMA pipe
    return mad(mad(nonzero01(D), 2u, nonzero01(R)), 4u, mad(nonzero01(U), 2u, nonzero01(L)));
}

struct FastOut {
    u32 lo, hi, reward, flags;
};

// One transition on the fast path.
//   lut(off) returns the u32 entry of the staged (hashed) row table at BYTE offset `off`.
//   Preconditions (checked by the caller): !has_big_tile(lo, hi).
//   rank = spawn rank among the empties (injected) — ignored when kPhilox (then w0 draws it);
//   exp  = spawned exponent (1 or 2);  one = the run-time constant 1 (see add()).
// ---- building blocks (also used by the fused rollout kernel) ---------------------------------

// Afterstate of (lo,hi) under the action whose orientation words are `o`.  Returns the merge
// score; `moved` is evaluated in the oriented frame (core.py:258).
template <class Lut>
G2048_HD u32 move_fast(u32 lo, u32 hi, const Orient& o, const Lut& lut, u32 one, u32& al, u32& ah,
                       bool& moved) {
    u32 xl = lo, xh = hi;
    orient_forward(xl, xh, o);
    const u32 yl = hash_rows(xl), yh = hash_rows(xh);
    // byte offsets: the x4 is an IMAD.SHL (FMA pipe), one ALU op per lookup for the extraction
    const u32 e0 = lut((yl << 2) & 0x3FFFCu), e1 = lut(prmt(yl, 0u, 0x4432u) << 2);
    const u32 e2 = lut((yh << 2) & 0x3FFFCu), e3 = lut(prmt(yh, 0u, 0x4432u) << 2);
    al = prmt(e0, e1, 0x7632);
    ah = prmt(e2, e3, 0x7632);
    moved = (al != xl) | (ah != xh);
    const u32 reward = (add(add(e0, e1, one), add(e2, e3, one), one) << 2) & 0x3FFFCu;
    orient_backward(al, ah, o);
    return reward;
}

// Marker (bit 3 of the target nibble, in the half `in_hi` selects) of the rank-th empty cell in
// row-major order, given the non-empty indicators and tile counts of the two halves.
// Inclusive prefix count of EMPTY cells at nibble k of a half = (k + 1) - prefix(tiles), so with
// t = tile indicators at bit 0:  g = 0x87654321 + (7 - r)*0x11111111 - t*0x11111111  has bit 3
// set exactly in the nibbles whose empty-count exceeds r (every nibble stays <= 15).
G2048_HD u32 spawn_marker(u32 nlo, u32 nhi, u32 tiles_lo, u32 rank, bool& in_hi) {
    const u32 clo = 8u - tiles_lo;
    in_hi = rank >= clo;
    const u32 nsel = in_hi ? nhi : nlo;
    const u32 r = in_hi ? rank - clo : rank;
    const u32 g = mad(shr<3>(nsel), 0xEEEEEEEFu, mad(r, 0xEEEEEEEFu, 0xFEDCBA98u)) & 0x88888888u;
    return g & (0u - g);
}

template <bool kPhilox, class Lut>
G2048_HD FastOut transition_fast(u32 lo, u32 hi, const Orient& o, u32 rank, u32 exp, u32 w0,
                                 const Lut& lut, u32 one) {
    u32 al, ah;
    bool moved;
    const u32 reward = move_fast(lo, hi, o, lut, one, al, ah, moved);

    // ---- spawn on the rank-th empty cell (row-major) of the afterstate ----
    u32 nlo = nz8(al, one), nhi = nz8(ah, one);
    const u32 tiles_lo = popc(nlo), ne = 16u - (tiles_lo + popc(nhi));
    if (kPhilox) rank = mulhi(w0, ne);
    const bool ok = rank < ne;
    bool in_hi;
    const u32 bit3 = spawn_marker(nlo, nhi, tiles_lo, rank, in_hi);
    const u32 tile = shr<3>(bit3) * exp;         // exponent 1 or 2 at bit 0 of that nibble
    if (moved & ok) {
        if (in_hi) { ah += tile; nhi += bit3; } else { al += tile; nlo += bit3; }
    }

    // ---- flags of the new board ----
    const u32 m = legal_mask(al, ah, nlo, nhi, one);
    u32 f = mad(nonzero01(m), 0xFFFFFFF0u, m) + 0x10u;   // + DONE (0x10) iff no action is legal
    if (moved) f += 0x20u;                               // MOVED
    if (moved & !ok) f += 0x80u;                         // BADSPAWN
    return FastOut{al, ah, reward, f};
}

}  // namespace fast
}  // namespace g2048
