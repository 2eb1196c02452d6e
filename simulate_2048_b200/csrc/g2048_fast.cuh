// g2048_fast.cuh — the branch-free transition used by the HBM-bound step kernel.
//
// Measured on B200 (profiles/r01_microbench_pipes.txt): LOP3 / SHF / PRMT / ISETP / SEL / IADD3
// issue at 2 warp-instr/clk/SM (the "ALU" pipe; the probe's "IADD 4.0" is ptxas folding two adds
// into one IADD3), IMAD at 2 on the separate FMA pipe, IMAD.HI / IMAD.WIDE at 1, POPC at 0.5, a
// random LDS.32 costs 3.5 clk.  ncu on the first version of the
// step kernel: ALU pipe 94 % busy, FMA pipe 10 %, 220 instructions per env (profiles/).  The
// HBM roofline leaves ~1 clk per env-step per SM, so the transition is written to
//   * never branch on data (actions are random: a divergent warp pays for every path),
//   * keep the ALU pipe for what only it can do (3-input logic, byte permutes, compares),
//   * put every shift on the FMA pipe (x << k is IMAD.SHL, x >> k is IMAD.HI by 2^(32-k)),
//     and use adds where the operands have disjoint bits.
//
// Orientation.  Up / right / down become "left" through a transform F_a and back through
// G_a = F_a^-1.  Lines are independent, so any line order serves, and each F_a factors into ONE
// nibble-level involution S_a applied inside each 32-bit half and ONE byte permutation P_a:
//   up    : transpose       S = swap nibbles 1<->4, 3<->6 (delta swap, distance 12 bits)
//   right : mirror          S = swap the nibbles of every byte (distance 4)
//   down  : anti-transpose  S = swap nibbles 0<->5, 2<->7 (distance 20) — columns taken right to
//           left and read bottom-up, which keeps S an involution (a plain transpose followed by a
//           mirror is a 4-cycle on the 2x2 nibble blocks and would need two stages)
//   F_a = P_a . S_a          G_a = S_a . P_a      (every P_a is an involution)
// The delta swap t = (x ^ x>>k) & m; x ^= t ^ t<<k takes its distance from two per-thread
// multipliers (x>>k = mulhi(x, 2^(32-k)), t<<k = t * 2^k) and is a no-op when m = 0, so every
// action runs the same 4 instructions per half; P_a is one PRMT pair whose selector is per-thread
// data.  The five per-action words come from a 4-entry shared-memory table.
//
// Row-table entry (u32): bits 0..15 = (merge score of the row) >> 2, bits 16..31 = the row after
// a left move.  With the score in the LOW half the four entries of a board simply add up:
// on the fast path every row has nibbles <= 13, so a field is <= 8192 and the sum <= 32768.
//
// Every primitive has a host twin so the identical dataflow is unit-tested on the CPU
// (tests/test_fast_path_host.py builds tests/fast_host_check.cpp against the oracle).
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

// acc + k if x != 0, as a predicated IMAD: a LOP3 that produces x can set the predicate itself
// (LOP3.LUT P, ...), so no ALU-pipe instruction is spent on the test.  k = 0 stands for `one`
// (a literal multiplier 1 would let ptxas fold the mad into an ALU-pipe add).
template <u32 k>
G2048_HD u32 add_if_nonzero(u32 acc, u32 x, u32 one) {
#if defined(__CUDA_ARCH__)
    if (k == 1u)
        asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p mad.lo.u32 %0, %2, %2, %0;\n\t}"
            : "+r"(acc) : "r"(x), "r"(one));
    else
        asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p mad.lo.u32 %0, %2, %3, %0;\n\t}"
            : "+r"(acc) : "r"(x), "r"(one), "n"(k));
    return acc;
#else
    return x ? acc + k * one : acc;
#endif
}

// x - 16 if x != ref (cancels the DONE bit once some mask bit was added), predicated IMAD
G2048_HD u32 sub16_if_differ(u32 x, u32 ref, u32 one) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, %1;\n\t@p mad.lo.u32 %0, %2, 0xFFFFFFF0, %0;\n\t}"
        : "+r"(x) : "r"(ref), "r"(one));
    return x;
#else
    return x != ref ? x - 16u * one : x;
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
// a * b as {hi, lo} in one instruction (IMAD.WIDE.U32): a digit of the mixed-radix random word
// and its remainder (g2048_device.cuh, Philox section)
G2048_HD void mul_wide(u32 a, u32 b, u32& hi, u32& lo) {
    const u64 p = (u64)a * b;
    hi = (u32)(p >> 32);
    lo = (u32)p;
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
    u32 mask;            // delta-swap mask (0 = no nibble stage)
    u32 mul_hi;          // 2^(32-k): x >> k as the high word of a product (IMAD.HI)
    u32 mul_lo;          // 2^k:      t << k as a product (IMAD)
    u32 sel_lo, sel_hi;  // byte permutation (PRMT selectors); every P_a is an involution, so the
                         // same pair serves forwards and backwards
    u32 shift;           // k itself, for the variant that shifts on the ALU pipe (SHF)
    u32 pad1, pad2;
};

// action 0 left, 1 up, 2 right, 3 down.  Derivation of the selectors: tools/derive_orient.py.
#define G2048_ORIENT_TABLE                                                             \
    {                                                                                  \
        {0x00000000u, 0x00000000u, 0x00000000u, 0x3210u, 0x7654u, 0u, 0u, 0u},         \
        {0x0000F0F0u, 0x00100000u, 0x00001000u, 0x6240u, 0x7351u, 12u, 0u, 0u},        \
        {0x0F0F0F0Fu, 0x10000000u, 0x00000010u, 0x2301u, 0x6745u, 4u, 0u, 0u},         \
        {0x00000F0Fu, 0x00001000u, 0x00100000u, 0x3715u, 0x2604u, 20u, 0u, 0u},        \
    }

// the one conditional nibble stage: delta swap of the nibble pairs selected by the mask, with the
// shift distance k (12 / 4 / 20 bits) carried by two per-thread multipliers, so that both shifts
// run on the FMA pipe and the stage is the same instruction sequence for every action
template <bool kAluShift = false>
G2048_HD u32 stage(u32 x, const Orient& o) {
    const u32 t = lop3<(kA ^ kB) & kC>(x, kAluShift ? x >> o.shift : mulhi(x, o.mul_hi), o.mask);
    return lop3<kA ^ kB ^ kC>(x, t, t * o.mul_lo);
}

template <bool kAluShift = false>
G2048_HD void orient_forward(u32& lo, u32& hi, const Orient& o) {
    const u32 l = stage<kAluShift>(lo, o), h = stage<kAluShift>(hi, o);
    lo = prmt(l, h, o.sel_lo);
    hi = prmt(l, h, o.sel_hi);
}
template <bool kAluShift = false>
G2048_HD void orient_backward(u32& lo, u32& hi, const Orient& o) {
    const u32 l = prmt(lo, hi, o.sel_lo), h = prmt(lo, hi, o.sel_hi);
    lo = stage<kAluShift>(l, o);
    hi = stage<kAluShift>(h, o);
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
// kAluShift: the shift as SHF on the ALU pipe instead of IMAD.HI (4 cycles of the FMA-heavy pipe per
// warp) — for the Philox-mode step kernel, whose FMA-heavy pipe carries the generator's IMAD.WIDEs.
template <bool kAluShift = false>
G2048_HD u32 hash_rows(u32 w) { return lop3<kA ^ (kB & kC)>(w, kAluShift ? w >> 4 : shr<4>(w), 0x0FFF0FFFu); }
G2048_HD u32 hash_row16(u32 row) { return (row ^ ((row >> 4) & 0x0FFFu)) & 0xFFFFu; }

// true when some tile is >= 16384 (nibble >= 14): such boards leave the fast path, which keeps
// every row inside the staged table window and rules out a 15+15 merge.
G2048_HD u32 big_tile_half(u32 x) { return lop3<kA & kB & kC>(x, x << 1, x << 2); }  // dirty outside bit 3
G2048_HD u32 big_tile_bits(u32 lo, u32 hi) {
    return lop3<(kA | kB) & kC>(big_tile_half(lo), big_tile_half(hi), 0x88888888u);
}
G2048_HD bool has_big_tile(u32 lo, u32 hi) { return big_tile_bits(lo, hi) != 0u; }

// Legal mask of board (lo,hi) given its non-empty indicators (bit 3 per nibble).
// Every cell is compared with its LEFT / UPPER neighbour (left shifts: IMAD.SHL, FMA pipe).
//   left / up   : a tile whose neighbour is empty or equal            n & ~(e & s)
//   right / down: a tile equal to its neighbour, or an empty cell
//                 whose neighbour is a tile                           (n & ~e) | (~n & s)
// The shifted-in neighbour bits at column 0 / row 0 are garbage and masked out at the end.
// Returns base + mask (base = 0 for the plain mask; the step kernel passes its status bits so
// that the first mask bit is a predicated add as well instead of a select).
G2048_HD u32 legal_mask(u32 lo, u32 hi, u32 nlo, u32 nhi, u32 one, u32 base = 0u) {
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
    // any bit set -> mask bit: four predicated IMADs, the predicates set by the LOP3s above
    u32 m = base;
    m = add_if_nonzero<1u>(m, L, one);
    m = add_if_nonzero<2u>(m, U, one);
    m = add_if_nonzero<4u>(m, R, one);
    m = add_if_nonzero<8u>(m, D, one);
    return m;
}

struct FastOut {
    u32 lo, hi, reward, flags;
};

// One transition on the fast path.
//   lut(off) returns the u32 entry of the staged (hashed) row table at BYTE offset `off`.
//   Preconditions (checked by the caller): !has_big_tile(lo, hi).
//   rank = spawn rank among the empties (injected), exp_mul = spawned exponent (1 or 2) << 29 (the
//   tile is the high word of marker * exp_mul) — both ignored when kPhilox: then they are the first
//   two mixed-radix digits of the env's random word w0;
//   one = the run-time constant 1 (see add()).
// ---- building blocks (also used by the fused rollout kernel) ---------------------------------

// Afterstate of (lo,hi) under the action whose orientation words are `o`.  Returns the merge
// score; `moved` is evaluated in the oriented frame (core.py:258).
template <bool kAluShift = false, class Lut>
G2048_HD u32 move_fast(u32 lo, u32 hi, const Orient& o, const Lut& lut, u32 one, u32& al, u32& ah,
                       bool& moved) {
    u32 xl = lo, xh = hi;
    orient_forward<kAluShift>(xl, xh, o);
    const u32 yl = hash_rows<kAluShift>(xl), yh = hash_rows<kAluShift>(xh);
    // byte offsets: one ALU op per lookup for the extraction, the x4 on the FMA pipe (IMAD.SHL;
    // for the high row a run-time multiplier, or ptxas folds shift + base into an ALU-pipe LEA)
    const u32 four = one << 2;
    const u32 e0 = lut((yl << 2) & 0x3FFFCu), e1 = lut(prmt(yl, 0u, 0x4432u) * four);
    const u32 e2 = lut((yh << 2) & 0x3FFFCu), e3 = lut(prmt(yh, 0u, 0x4432u) * four);
    al = prmt(e0, e1, 0x7632);
    ah = prmt(e2, e3, 0x7632);
    moved = (al != xl) | (ah != xh);
    const u32 reward = shr<14>(add(add(e0, e1, one), add(e2, e3, one), one) << 16);  // (sum & 0xFFFF) * 4
    orient_backward<kAluShift>(al, ah, o);
    return reward;
}

// Marker (bit 3 of the target nibble, in the half `in_hi` selects) of the rank-th empty cell in
// row-major order, given the non-empty indicators and tile counts of the two halves.
// Inclusive prefix count of EMPTY cells at nibble k of a half = (k + 1) - prefix(tiles), so with
// t = tile indicators at bit 0:  g = 0x87654321 + (7 - r)*0x11111111 - t*0x11111111  has bit 3
// set exactly in the nibbles whose empty-count exceeds r (every nibble stays <= 15).
G2048_HD u32 spawn_prefix(u32 nlo, u32 nhi, u32 tiles_lo, u32 rank, u32 one, bool& in_hi) {
    const u32 clo = mad(tiles_lo, 0u - one, 8u);   // 8 - tiles_lo as an IMAD
    in_hi = rank >= clo;
    const u32 nsel = in_hi ? nhi : nlo;
    const u32 r = in_hi ? rank - clo : rank;   // (ptxas turns a predicated subtract into SEL + IADD as well)
    return mad(shr<3>(nsel), 0xEEEEEEEFu, mad(r, 0xEEEEEEEFu, 0xFEDCBA98u));
}
// The empty-count is non-decreasing along the nibbles, so bit 3 of g is 0..0 1..1 from the target
// upwards: the lowest set indicator is the one whose lower neighbour is clear (one LOP3 with the
// shift on the FMA pipe, instead of mask, negate, and).
constexpr u32 kLowestIndicator = kA & ~kB & kC & 0xFF;
G2048_HD u32 spawn_marker(u32 nlo, u32 nhi, u32 tiles_lo, u32 rank, u32 one, bool& in_hi) {
    const u32 g = spawn_prefix(nlo, nhi, tiles_lo, rank, one, in_hi);
    return lop3<kLowestIndicator>(g, g << 4, 0x88888888u);
}

// The spawn commit of transition_fast, branch-free (a divergent branch here costs BSSY / BRA /
// BSYNC and a reconvergence per env).  g = spawn_prefix(...); the marker is taken from it only when
// the move is legal and the spawn valid (rank < ne; a Philox rank always is) and is 0 otherwise,
// so the predicated adds are no-ops then.  tile = exponent (1 or 2) at bit 0 of the marked nibble
// (one IMAD.HI: exp_mul = exponent << 29), added to the half `in_hi` selects together with its
// non-empty indicator.  status = the flag bits besides the mask (0x10 is cancelled by the caller
// when some action is legal): not moved -> 0x10, moved without a valid spawn -> MOVED | BADSPAWN,
// moved and spawned -> MOVED.
template <bool kCheckRank>
G2048_HD void spawn_commit(u32& al, u32& ah, u32& nlo, u32& nhi, u32& status, u32 g, u32 exp_mul,
                           bool moved, u32 rank, u32 ne, bool in_hi, u32 one) {
#if defined(__CUDA_ARCH__)
#define G2048_COMMIT_BODY                                    \
        "setp.ne.u32 ph, %9, 0;\n\t"                          \
        "selp.u32 %4, 0xB0, 0x10, pm;\n\t"                    \
        "@pd mov.u32 %4, 0x30;\n\t"                           \
        "shl.b32 g4, %5, 4;\n\t"                              \
        "lop3.b32 b, %5, g4, 0x88888888, 0x20;\n\t"           \
        "selp.u32 b, b, 0, pd;\n\t"                           \
        "mul.hi.u32 t, b, %6;\n\t"                            \
        "@ph mad.lo.u32 %1, t, %11, %1;\n\t"                  \
        "@ph mad.lo.u32 %3, b, %11, %3;\n\t"                  \
        "@!ph mad.lo.u32 %0, t, %11, %0;\n\t"                 \
        "@!ph mad.lo.u32 %2, b, %11, %2;\n\t"                 \
        "}"
    if (kCheckRank)
        asm("{\n\t.reg .pred pm, pd, ph;\n\t.reg .u32 b, t, g4;\n\t"
            "setp.ne.u32 pm, %7, 0;\n\t"
            "setp.lt.and.u32 pd, %8, %10, pm;\n\t" G2048_COMMIT_BODY
            : "+r"(al), "+r"(ah), "+r"(nlo), "+r"(nhi), "=r"(status)
            : "r"(g), "r"(exp_mul), "r"((u32)moved), "r"(rank), "r"((u32)in_hi), "r"(ne), "r"(one));
    else
        asm("{\n\t.reg .pred pm, pd, ph;\n\t.reg .u32 b, t, g4;\n\t"
            "setp.ne.u32 pm, %7, 0;\n\t"
            "setp.ne.u32 pd, %7, 0;\n\t" G2048_COMMIT_BODY
            : "+r"(al), "+r"(ah), "+r"(nlo), "+r"(nhi), "=r"(status)
            : "r"(g), "r"(exp_mul), "r"((u32)moved), "r"(rank), "r"((u32)in_hi), "r"(0u), "r"(one));
#undef G2048_COMMIT_BODY
#else
    const bool ok = !kCheckRank || rank < ne;
    status = moved ? (ok ? 0x30u : 0xB0u) : 0x10u;
    const u32 b = (moved && ok) ? lop3<kLowestIndicator>(g, g << 4, 0x88888888u) : 0u;
    const u32 t = mulhi(b, exp_mul);
    if (in_hi) { ah += t; nhi += b; } else { al += t; nlo += b; }
#endif
}

template <bool kPhilox, class Lut>
G2048_HD FastOut transition_fast(u32 lo, u32 hi, const Orient& o, u32 rank, u32 exp_mul, u32 w0,
                                 const Lut& lut, u32 one) {
    u32 al, ah;
    bool moved;
    const u32 reward = move_fast<kPhilox>(lo, hi, o, lut, one, al, ah, moved);

    // ---- spawn on the rank-th empty cell (row-major) of the afterstate ----
    u32 nlo = nz8(al, one), nhi = nz8(ah, one);
    const u32 tiles_lo = popc(nlo), ne = 16u - (tiles_lo + popc(nhi));
    if (kPhilox) {
        // digit 1 (base ne) = spawn rank, digit 2 (base 10) == 0 -> tile 4 (exponent 2)
        u32 f1;
        mul_wide(w0, ne, rank, f1);
        // hi32(f1 * 10) == 0  <=>  f1 < ceil(2^32 / 10): a compare and a select on the ALU pipe
        // instead of IMAD.HI + VIMNMX + IMAD (6 cycles of the FMA-heavy pipe, this mode's busiest)
        exp_mul = f1 < 0x1999999Au ? 0x40000000u : 0x20000000u;
    }
    bool in_hi;
    const u32 g = spawn_prefix(nlo, nhi, tiles_lo, rank, one, in_hi);
    // a Philox rank is < ne by construction (and a move always leaves an empty cell)
    u32 status;
    spawn_commit<!kPhilox>(al, ah, nlo, nhi, status, g, exp_mul, moved, rank, ne, in_hi, one);

    // ---- flags of the new board: mask | DONE (0x10) iff no action is legal | status ----
    const u32 f = legal_mask(al, ah, nlo, nhi, one, status);
    return FastOut{al, ah, reward, sub16_if_differ(f, status, one)};
}

}  // namespace fast
}  // namespace g2048
