// g2048_kernels.cu — sm_100a kernels and the C ABI of libg2048.so (include/g2048.h).
//
// Hot path (BASELINE.json north_star): one batched 2048 transition per env
//   board(u64) , action(u8) [, spawn(u8)]  ->  board(u64), reward(i32), flags(u8)
// HBM-bound by design: 22 B/env (Philox) or 23 B/env (injected spawns).  Design:
//   * persistent grid, one 640-thread CTA per SM (512 with Philox) over a contiguous range of
//     warp-tiles, 4 consecutive envs per thread so every stream is one wide access per thread:
//     256-bit board load/store (LDG.E.256 / STG.E.256), 128-bit reward store, 32-bit action /
//     spawn / flags; launched as a programmatic dependent, so the launch latency and table staging
//     of launch N+1 overlap the tail of launch N;
//   * the 57344-entry row table (224 KiB, hashed against bank conflicts) is staged into shared
//     memory once per CTA by TMA bulk copies (cp.async.bulk, UBLKCP) that complete on an mbarrier
//     while the first boards are already in flight; a move is then 4 LDS + 2 PRMT;
//   * up / right / down are reduced to "left" by ONE per-action nibble delta swap and one byte
//     permutation each way (g2048_fast.cuh), reward rides in the table entry, the legal mask /
//     done of the NEW board and the tile spawn are nibble-parallel register arithmetic in the
//     same pass, without a data-dependent branch;
//   * inputs of tile t+1 are loaded before tile t is computed (register double buffering).
// The other entry points (expand, rollout, env_step, observe, n-step returns, ...) follow below.
// No Triton, no CPU fallback: every entry point either launches these kernels or fails.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <vector>

#include "g2048_device.cuh"
#include "g2048_fast.cuh"

using namespace g2048;

__constant__ g2048::fast::Orient c_orient[4] = G2048_ORIENT_TABLE;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return G2048_ECUDA;
}
#define CU(call)                                               \
    do {                                                       \
        cudaError_t e_ = (call);                               \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);    \
    } while (0)

extern "C" int g2048_abi_version(void) { return G2048_ABI_VERSION; }
extern "C" const char* g2048_last_cuda_error(void) { return g_cuda_err; }
extern "C" const char* g2048_error_string(int code) {
    switch (code) {
        case G2048_OK: return "ok";
        case G2048_EINVAL: return "invalid argument (NULL or misaligned pointer, bad enum)";
        case G2048_ECUDA: return "CUDA runtime failure (see g2048_last_cuda_error)";
        case G2048_ENODEV: return "no usable CUDA device (sm_100 required)";
        case G2048_ERANGE: return "tile value not representable (0 or a power of two in [2, 32768])";
        default: return "unknown error";
    }
}

// ---------------------------------------------------------------------------------------------
// per-device state: the row table
// ---------------------------------------------------------------------------------------------
// tab: natural order (general kernels, slow path); tab_hashed: entry of `row` stored at
// fast::hash_row16(row) — the copy the step kernel stages into shared memory.
__global__ void build_table_kernel(u32* __restrict__ tab, u32* __restrict__ tab_hashed) {
    u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < kRows) {
        const u32 e = make_entry(row);
        tab[row] = e;
        tab_hashed[fast::hash_row16(row)] = e;
    }
}

// Library state (all of it; include/g2048.h "State"):
//   g_dev[]   per-device row tables, built once under the device's mutex and published with a
//             release store of `ready` (readers take the acquire load on every call);
//   g_impl / g_step_threads / g_step_groups   process-wide measurement knobs (g2048_set_impl,
//             g2048_set_tuning), atomics read once per call;
//   g_pipes   pool of host-pipeline resources of g2048_step_host (streams + staging buffers),
//             handed out under a mutex, returned after the call, freed by g2048_shutdown.
struct DeviceState {
    std::mutex mu;
    std::atomic<bool> ready{false};
    u32* table = nullptr;
    u32* table_hashed = nullptr;
    int sms = 0;
    size_t smem_optin = 0;
    int children_occupancy = 1;  // resident CTAs per SM of expand_children_kernel
    // Stream-ordered scratch (queue counters, scan aggregates, per-env rollout results) comes from
    // the library's own pool with an unlimited release threshold: the default pool hands its memory
    // back to the driver at every synchronisation, which made each call after a host sync pay for
    // fresh virtual-memory mappings (hundreds of microseconds).
    cudaMemPool_t pool = nullptr;
};
static DeviceState g_dev[64];
static std::atomic<int> g_impl{G2048_IMPL_AUTO};
static std::atomic<int> g_step_threads{0};  // 0 = automatic: 640 (5 warps per scheduler at 96 registers: +1 % over 512 in both spawn modes)
// programmatic dependent launch of the step kernel (environment G2048_PDL=0 turns it off)
static const int g_pdl = [] { const char* e = getenv("G2048_PDL"); return e ? (atoi(e) != 0 ? 1 : 0) : 1; }();
static std::atomic<int> g_step_groups{1};   // groups of 4 envs per thread (1 or 2)  // tuning knob (g2048_set_tuning)

static int device_occupancy(DeviceState& s);  // defined next to the kernels it measures

static int current_state(DeviceState** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDevice");
        return G2048_ENODEV;
    }
    if (dev < 0 || dev >= 64) return G2048_EINVAL;
    DeviceState& s = g_dev[dev];
    if (!s.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(s.mu);
        if (!s.ready.load(std::memory_order_relaxed)) {
            cudaDeviceProp p;
            CU(cudaGetDeviceProperties(&p, dev));
            if (p.major != 10) {
                snprintf(g_cuda_err, sizeof g_cuda_err, "device %d is sm_%d%d, need sm_100", dev,
                         p.major, p.minor);
                return G2048_ENODEV;
            }
            s.sms = p.multiProcessorCount;
            s.smem_optin = p.sharedMemPerBlockOptin;
            CU(cudaMalloc(&s.table, kRows * sizeof(u32)));
            CU(cudaMalloc(&s.table_hashed, kRows * sizeof(u32)));
            build_table_kernel<<<kRows / 256, 256>>>(s.table, s.table_hashed);
            CU(cudaGetLastError());
            CU(cudaDeviceSynchronize());
            int rc = device_occupancy(s);
            if (rc) return rc;
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            CU(cudaMemPoolCreate(&s.pool, &props));
            unsigned long long keep = ~0ull;
            CU(cudaMemPoolSetAttribute(s.pool, cudaMemPoolAttrReleaseThreshold, &keep));
            s.ready.store(true, std::memory_order_release);
        }
    }
    *out = &s;
    return G2048_OK;
}

extern "C" int g2048_init(int device) {
    if (device >= 0) CU(cudaSetDevice(device));
    DeviceState* s;
    return current_state(&s);
}

extern "C" int g2048_set_tuning(int step_threads, int envs_per_thread) {
    if (step_threads != 0 && step_threads != 512 && step_threads != 640 && step_threads != 768 && step_threads != 1024)
        return G2048_EINVAL;
    if (envs_per_thread != 4 && envs_per_thread != 8) return G2048_EINVAL;
    if (step_threads == 640 && envs_per_thread == 8) return G2048_EINVAL;  // not instantiated
    g_step_threads.store(step_threads);
    g_step_groups.store(envs_per_thread / 4);
    return G2048_OK;
}

extern "C" int g2048_set_impl(int impl) {
    if (impl < G2048_IMPL_AUTO || impl > G2048_IMPL_GATHER) return G2048_EINVAL;
    g_impl.store(impl);
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// wide global accesses (streaming: nothing is re-read, keep it out of L1)
// ---------------------------------------------------------------------------------------------
struct Boards4 {
    u64 b[4];
};

// (no .nc: boards_out may alias boards_in, and PTX requires .nc data to stay read-only for the
// whole kernel; a thread reads its four boards before it writes them, so plain loads are exact)
__device__ __forceinline__ Boards4 ld_boards4(const u64* p) {
    Boards4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.b[0]), "=l"(r.b[1]), "=l"(r.b[2]), "=l"(r.b[3])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_boards4(u64* p, const Boards4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.b[0]),
                 "l"(v.b[1]), "l"(v.b[2]), "l"(v.b[3])
                 : "memory");
}
__device__ __forceinline__ u32 ld_u32_stream(const void* p) {
    u32 r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_u32_stream(void* p, u32 v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_int4_stream(void* p, int4 v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy of the row table into shared memory, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) {
    return (u32)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void table_stage_begin(u32* smem_tab, u64* bar, const u32* gtab) {
    if (threadIdx.x == 0) {
        const u32 b = smem_addr(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b),
                     "r"(kSmemTableBytes)
                     : "memory");
        constexpr u32 kChunk = 32768u;  // 7 bulk copies of 32 KiB
#pragma unroll
        for (u32 off = 0; off < kSmemTableBytes; off += kChunk) {
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(smem_addr(smem_tab) + off),
                "l"(reinterpret_cast<const char*>(gtab) + off), "r"(kChunk), "r"(b)
                : "memory");
        }
    }
}

__device__ __forceinline__ void table_stage_wait(u64* bar) {
    __syncthreads();  // the barrier init by thread 0 is visible before anyone polls it
    const u32 b = smem_addr(bar);
    u32 done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(b)
            : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// one transition, shared by every step kernel
// ---------------------------------------------------------------------------------------------
template <class Table, bool kPhilox>
__device__ __forceinline__ void transition(Board b, u32 action, u32 spawn, u64 key, u64 env,
                                           u32 step, const Table& T, Board& out, u32& reward,
                                           u32& flags) {
    MoveResult m = move_action(b, action & 3u, T);
    const bool moved = (m.after.lo != b.lo) | (m.after.hi != b.hi);  // core.py:258
    Board nb = m.after;  // an illegal move leaves the board as it was (gameboard.py:325)
    u32 nlo = nz8(nb.lo), nhi = nz8(nb.hi);
    u32 f = (moved ? G2048_FLAG_MOVED : 0u) | (m.overflow ? G2048_FLAG_OVERFLOW : 0u);
    // spawn on the afterstate of a legal move (gameboard.py:323, core.py:267)
    const u32 zlo = (~nlo >> 3) & 0x11111111u, zhi = (~nhi >> 3) & 0x11111111u;
    const u32 clo = __popc(zlo), n_empty = clo + __popc(zhi);
    u32 rank, exp;
    bool ok = true;
    if (kPhilox) {
        draw_spawn(philox_word(key, env, step, G2048_PHILOX_TAG_STEP), n_empty, rank, exp);
    } else {
        exp = (spawn & 0x10u) ? 2u : 1u;
        rank = spawn & 0xFu;
        if (spawn & 0x80u) {  // absolute cell: translate to its rank among the empties
            const u32 cell = spawn & 0xFu, sh = (cell & 7u) * 4u;
            const u32 zh = (cell & 8u) ? zhi : zlo;
            ok = (zh >> sh) & 1u;
            rank = __popc(zh & ((1u << sh) - 1u)) + ((cell & 8u) ? clo : 0u);
        }
    }
    ok = ok & (rank < n_empty);
    if (moved & ok) {
        const bool in_hi = rank >= clo;
        const u32 z = in_hi ? zhi : zlo;
        const u32 r = in_hi ? rank - clo : rank;
        const u32 g = (z * 0x11111111u + (0x77777777u - r * 0x11111111u)) & 0x88888888u;
        const u32 bit3 = g & (0u - g);  // bit 3 of the target nibble
        const u32 tile = (bit3 >> 3) * exp;
        if (in_hi) {
            nb.hi |= tile;
            nhi |= bit3;
        } else {
            nb.lo |= tile;
            nlo |= bit3;
        }
    }
    if (moved & !ok) f |= G2048_FLAG_BADSPAWN;
    f |= board_flags(legal_mask(nb, nlo, nhi));
    out = nb;
    reward = m.score;
    flags = f;
}

struct StepArgs {
    const u64* boards_in;
    const uint8_t* actions;
    const uint8_t* spawns;
    u64 seed, env_offset;
    const u32* step_ctr;
    u32 step_base;
    const u64* env_keys;
    const uint8_t* done_in;
    u64* boards_out;
    int32_t* rewards;
    uint8_t* flags;
    size_t n;
    const u32* table;
    const u32* table_hashed;
    u32 one;  // always 1: a run-time multiplier that keeps adds on the FMA pipe (g2048_fast.cuh)
    PhiloxKeys keys;  // round keys of `seed`
};

// ---------------------------------------------------------------------------------------------
// fast step kernel: smem table, 4 envs / thread, persistent, double buffered
// ---------------------------------------------------------------------------------------------

struct TileIn {
    Boards4 boards;
    u32 actions, spawns;
};

// per-thread stream cursors: advanced once per loop trip, every access a compile-time offset away
struct StepCursor {
    const u64* boards_in;
    const uint8_t* actions;
    const uint8_t* spawns;
    u64* boards_out;
    int32_t* rewards;
    uint8_t* flags;
    u64 env;  // global env id of the thread's first env (Philox counter)
    __device__ __forceinline__ void advance(size_t d) {
        boards_in += d; actions += d; spawns += d; boards_out += d; rewards += d; flags += d; env += d;
    }
};

template <bool kPhilox>
__device__ __forceinline__ TileIn load_tile(const StepCursor& c, size_t off) {
    TileIn t;
    t.boards = ld_boards4(c.boards_in + off);
    t.actions = ld_u32_stream(c.actions + off);
    t.spawns = kPhilox ? 0u : ld_u32_stream(c.spawns + off);
    return t;
}

// exact out-of-line transition for the rare envs the fast path does not cover (a tile >= 16384
// on the board, or an absolute-cell spawn byte)
template <bool kPhilox>
__device__ __noinline__ uint4 transition_slow(u64 board, u32 action, u32 spawn, u64 key, u64 env,
                                              u32 step, const u32* __restrict__ table) {
    Board nb;
    u32 reward, flags;
    transition<GlobalTable, kPhilox>(split(board), action, spawn, key, env, step, GlobalTable{table},
                                     nb, reward, flags);
    return make_uint4(nb.lo, nb.hi, reward, flags);
}


// The dynamic shared memory of the step / rollout kernels: [row table | mbarrier | ...].
// Indexing the extern symbol directly (instead of a generic pointer) keeps the address space
// known to ptxas, so a lookup is `LDS R, [R + UR]` with the base in a uniform register.
extern __shared__ __align__(128) u32 smem_words[];


struct SmemLut {
    // staged (hashed) row table at the start of dynamic shared memory, indexed by BYTE offset
    __device__ __forceinline__ u32 operator()(u32 byte_off) const {
        return *reinterpret_cast<const u32*>(reinterpret_cast<const unsigned char*>(smem_words) + byte_off);
    }
};

// Four consecutive envs of one thread.  The byte streams are decoded once per tile (one LOP3
// per field for all four envs, then one PRMT per env), the four transitions are independent so
// the scheduler interleaves them, and the rare inputs the fast path does not cover (a tile
// >= 16384, an absolute-cell spawn byte) send the whole thread-tile to the exact slow path.
template <bool kPhilox>
__device__ __forceinline__ void step_tile4(const Boards4& in, u32 actions, u32 spawns,
                                           const StepArgs& a, u64 env0, const SmemLut& lut,
                                           const uint4* __restrict__ s_orient, u64* out_boards,
                                           int32_t* out_rewards, uint8_t* out_flags) {
    // Range test in two levels.  Cheap filter (one LOP3 per word): bit 3 of acc is set where some
    // nibble has bits 3 and 2 set, i.e. a tile >= 4096.  Only then the exact test (nibble >= 14)
    // runs — never on the benchmark set or in random play, rarely with a strong player.
    u32 acc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u32 lo = (u32)in.b[k], hi = (u32)(in.b[k] >> 32);
        acc = fast::lop3<fast::kA | (fast::kB & fast::kC)>(acc, lo, lo << 1);
        acc = fast::lop3<fast::kA | (fast::kB & fast::kC)>(acc, hi, hi << 1);
    }
    // One test on the common path: some tile >= 4096 or some absolute-cell spawn byte (bit 7)
    const u32 abs4 = kPhilox ? 0u : spawns & 0x80808080u;
    bool slow = false;
    if (((acc & 0x88888888u) | abs4) != 0u) {
        u32 big = abs4;
#pragma unroll
        for (int k = 0; k < 4; ++k) big |= fast::big_tile_bits((u32)in.b[k], (u32)(in.b[k] >> 32));
        slow = big != 0u;
    }
    // Each branch ends in its own stores: merging the results of the out-of-line calls with the
    // fast path's into one set of registers costs ~10 moves per tile in front of the 256-bit store.
    if (slow) {
        Boards4 ob;
        int rws[4];
        u32 fl = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // unrolled: static indices keep the arrays in registers
            const uint4 r = transition_slow<kPhilox>(in.b[k], (actions >> (8 * k)) & 0xFFu,
                                                     (spawns >> (8 * k)) & 0xFFu, a.seed, env0 + k,
                                                     a.step_base, a.table);
            ob.b[k] = ((u64)r.y << 32) | r.x;
            rws[k] = (int)r.z;
            fl |= r.w << (8 * k);
        }
        st_boards4(out_boards, ob);
        st_int4_stream(out_rewards, make_int4(rws[0], rws[1], rws[2], rws[3]));
        st_u32_stream(out_flags, fl);
        return;
    }
    Boards4 ob;
    int rws[4];
    u32 fl = 0;
    // Philox: the thread's four envs are one quad (env0 is a multiple of 4), i.e. ONE block
    u32 pw[4] = {0u, 0u, 0u, 0u};
    if (kPhilox) {
        const Philox4 p = philox_quad(a.keys, env0, a.step_base, G2048_PHILOX_TAG_STEP);
        pw[0] = p.x; pw[1] = p.y; pw[2] = p.z; pw[3] = p.w;
    }
    const u32 act4 = (actions << 5) & 0x60606060u;     // byte k = 32 * action_k (byte offset)
    const u32 rank4 = spawns & 0x0F0F0F0Fu;            // byte k = spawn rank
    const u32 exp4 = ((spawns & 0x10101010u) << 1) + 0x20202020u;       // byte k = exponent (1 or 2) << 5
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned char* op = reinterpret_cast<const unsigned char*>(s_orient) + fast::prmt(act4, 0u, 0x4440u + k);
        const uint4 o0 = *reinterpret_cast<const uint4*>(op);
        fast::Orient o;
        o.mask = o0.x; o.mul_hi = o0.y; o.mul_lo = o0.z; o.sel_lo = o0.w;
        if (kPhilox) {   // + the shift distance itself: this mode shifts on the ALU pipe
            const uint2 o1 = *reinterpret_cast<const uint2*>(op + 16);
            o.sel_hi = o1.x; o.shift = o1.y;
        } else {
            o.sel_hi = *reinterpret_cast<const u32*>(op + 16);
        }
        u32 rank = 0, exp_mul = 0, w0 = 0;  // exp_mul = exponent << 29 (Philox: drawn inside)
        if (kPhilox) {
            w0 = pw[k];
        } else {
            rank = fast::prmt(rank4, 0u, 0x4440u + k);
            exp_mul = fast::prmt(exp4, 0u, 0x0444u + (k << 12));  // byte k of exp4 into byte 3
        }
        const fast::FastOut r = fast::transition_fast<kPhilox>((u32)in.b[k], (u32)(in.b[k] >> 32), o,
                                                               rank, exp_mul, w0, lut, a.one);
        ob.b[k] = ((u64)r.hi << 32) | r.lo;
        rws[k] = (int)r.reward;
        fl = fast::mad(r.flags, 1u << (8 * k), fl);   // one IMAD: shift and accumulate
    }
    st_boards4(out_boards, ob);
    st_int4_stream(out_rewards, make_int4(rws[0], rws[1], rws[2], rws[3]));
    st_u32_stream(out_flags, fl);
}

constexpr u32 kStepSmemBytes = kSmemTableBytes + 64u + 4u * (u32)sizeof(fast::Orient);

// kGroups groups of four consecutive envs per thread (4 or 8 envs): more independent work per
// warp for the scheduler and less per-env decode / address / loop overhead.
template <bool kPhilox, int kGroups>
struct TileBuf {
    TileIn g[kGroups];
};

template <bool kPhilox, int kGroups>
__device__ __forceinline__ void load_tiles(TileBuf<kPhilox, kGroups>& t, const StepCursor& c, size_t off) {
#pragma unroll
    for (int j = 0; j < kGroups; ++j) t.g[j] = load_tile<kPhilox>(c, off + 4 * j);
}

template <bool kPhilox, int kGroups>
__device__ __forceinline__ void compute_store_tiles(const TileBuf<kPhilox, kGroups>& t, const StepArgs& a,
                                                    const StepCursor& c, size_t off, const SmemLut& lut,
                                                    const uint4* __restrict__ s_orient) {
#pragma unroll
    for (int j = 0; j < kGroups; ++j)
        step_tile4<kPhilox>(t.g[j].boards, t.g[j].actions, t.g[j].spawns, a, c.env + off + 4 * j,
                            lut, s_orient, c.boards_out + off + 4 * j, c.rewards + off + 4 * j,
                            c.flags + off + 4 * j);
}

// Persistent CTA (one per SM).  Each CTA owns a CONTIGUOUS range of warp-tiles so that the next
// tile of every stream is a compile-time offset away; the loop is unrolled by two tiles that
// ping-pong between two register buffers (the loads of one tile are in flight while the other
// computes).
template <bool kPhilox, int kThreads, int kGroups>
__global__ void __launch_bounds__(kThreads, 1) step_lut_kernel(const StepArgs a) {
    u32* smem_tab = smem_words;
    u64* bar = reinterpret_cast<u64*>(smem_words + kSmemRows);
    uint4* s_orient = reinterpret_cast<uint4*>(smem_words + kSmemRows + 16);
    constexpr size_t kPerThread = 4 * kGroups;
    constexpr size_t kTile = (size_t)kThreads * kPerThread;

    table_stage_begin(smem_tab, bar, a.table_hashed);
    if (threadIdx.x < 32)
        reinterpret_cast<u32*>(s_orient)[threadIdx.x] = reinterpret_cast<const u32*>(c_orient)[threadIdx.x];

    // Work is split in units of one WARP-tile (32 x kPerThread consecutive envs, so that every warp
    // access stays aligned to whole 128-byte lines): CTA b owns the contiguous units
    // [w_begin, w_end), warp w takes units w_begin + w, + kWarps, ...  The CTAs' shares then differ
    // by at most one warp-tile (a split in whole CTA-tiles leaves 2 or 3 tiles per CTA at 2^20
    // boards: 16 % of that launch); only the last trip of a CTA is partly filled.
    constexpr size_t kWarps = kThreads / 32, kWarpEnvs = 32 * kPerThread;
    const size_t wunits = a.n / kWarpEnvs;
    const size_t w_begin = wunits * blockIdx.x / gridDim.x;
    const size_t w_end = wunits * (blockIdx.x + 1) / gridDim.x;
    const size_t mine = w_end - w_begin, wi = threadIdx.x >> 5;
    const size_t tiles = mine > wi ? (mine - wi + kWarps - 1) / kWarps : 0;  // trips of this warp
    StepCursor c{a.boards_in, a.actions, a.spawns, a.boards_out, a.rewards, a.flags, a.env_offset};
    c.advance((w_begin + wi) * kWarpEnvs + (threadIdx.x & 31u) * kPerThread);

    // Programmatic dependent launch: everything above (table staging, orientation words, cursor
    // set-up) reads nothing an earlier kernel of the stream wrote, so it may run while that kernel
    // drains; its outputs are awaited here.  The dependents of THIS launch may be scheduled at
    // once: they cannot become resident before our CTA leaves the SM (shared memory is full), and
    // they wait at the same point for our stores.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");

    TileBuf<kPhilox, kGroups> bufA, bufB;
    if (tiles != 0) load_tiles(bufA, c, 0);

    table_stage_wait(bar);  // __syncthreads inside also publishes s_orient
    const SmemLut lut{};

    // Bounds are compares of the LOW WORD of the board cursor against precomputed ends (one ISETP
    // each; a 64-bit pointer compare is four instructions, and a 32-bit trip counter makes ptxas
    // build a jump table (BRX) for the three tests, which costs 12 %).  A CTA's range is far below
    // 4 GiB and a warp's cursor only moves in whole tiles, so equality of the low words is exact.
    constexpr u32 kTileBytes = (u32)(kTile * sizeof(u64));
    const u32 end0 = (u32)reinterpret_cast<uintptr_t>(c.boards_in + tiles * kTile);
    const u32 end1 = end0 - kTileBytes, end2 = end0 - 2u * kTileBytes;
#define G2048_CUR() ((u32) reinterpret_cast<uintptr_t>(c.boards_in))
    while (G2048_CUR() != end0) {
        const bool hasB = G2048_CUR() != end1;
        if (hasB) load_tiles(bufB, c, kTile);
        compute_store_tiles(bufA, a, c, 0, lut, s_orient);
        if (!hasB) break;
        if (G2048_CUR() != end2) load_tiles(bufA, c, 2 * kTile);
        compute_store_tiles(bufB, a, c, kTile, lut, s_orient);
        c.advance(2 * kTile);
    }
#undef G2048_CUR

    // ragged tail (< one warp-tile): one env per thread through the exact path
    for (size_t i = wunits * kWarpEnvs + blockIdx.x * (size_t)kThreads + threadIdx.x; i < a.n;
         i += (size_t)gridDim.x * kThreads) {
        const uint4 r = transition_slow<kPhilox>(a.boards_in[i], a.actions[i], kPhilox ? 0u : a.spawns[i],
                                                 a.seed, a.env_offset + i, a.step_base, a.table);
        a.boards_out[i] = ((u64)r.y << 32) | r.x;
        a.rewards[i] = (int)r.z;
        a.flags[i] = (uint8_t)r.w;
    }
}

// ---------------------------------------------------------------------------------------------
// general step kernel: every optional input, any alignment, table gathered from global/L2.
// Used for small batches (no 224 KiB staging) and for the rarely used options.
// ---------------------------------------------------------------------------------------------
template <class Table>
__device__ __forceinline__ Table make_table(const u32* tab);
template <>
__device__ __forceinline__ GlobalTable make_table<GlobalTable>(const u32* tab) { return GlobalTable{tab}; }
template <>
__device__ __forceinline__ RegisterTable make_table<RegisterTable>(const u32*) { return RegisterTable{}; }

template <bool kPhilox, class Table>
__global__ void __launch_bounds__(256) step_general_kernel(const StepArgs a) {
    const Table T = make_table<Table>(a.table);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.n;
         i += (size_t)gridDim.x * blockDim.x) {
        const Board b = split(a.boards_in[i]);
        if (a.done_in && a.done_in[i]) {  // env.py:103-104 pass-through
            a.boards_out[i] = join(b);
            a.rewards[i] = 0;
            a.flags[i] = (uint8_t)(legal_mask(b) | G2048_FLAG_DONE);
            continue;
        }
        Board nb;
        u32 r, f;
        transition<Table, kPhilox>(b, a.actions[i], kPhilox ? 0u : a.spawns[i],
                                   a.env_keys ? a.env_keys[i] : a.seed, a.env_offset + i,
                                   a.step_ctr ? a.step_ctr[i] : a.step_base, T, nb, r, f);
        a.boards_out[i] = join(nb);
        a.rewards[i] = (int)r;
        a.flags[i] = (uint8_t)f;
    }
}

static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static int grid_for(size_t n, int threads, int sms, int per_sm) {
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// smallest batch for which staging the 224 KiB table per SM beats gathering it through L2
constexpr size_t kLutMinEnvs = (size_t)1 << 18;

extern "C" int g2048_step(const uint64_t* boards_in, const uint8_t* actions, const uint8_t* spawns,
                          uint64_t seed, uint64_t env_offset, const uint32_t* step_ctr,
                          uint32_t step_base, const uint64_t* env_keys, const uint8_t* done_in,
                          uint64_t* boards_out, int32_t* rewards, uint8_t* flags, size_t n,
                          void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards_in || !actions || !boards_out || !rewards || !flags) return G2048_EINVAL;
    if (!aligned(boards_in, 8) || !aligned(boards_out, 8) || !aligned(rewards, 4)) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    StepArgs a{reinterpret_cast<const u64*>(boards_in), actions, spawns, seed, env_offset, step_ctr,
               step_base, reinterpret_cast<const u64*>(env_keys), done_in,
               reinterpret_cast<u64*>(boards_out), rewards, flags, n, s->table, s->table_hashed, 1u,
               philox_key_schedule(seed)};

    const bool plain = !step_ctr && !env_keys && !done_in;
    const bool vec_ok = aligned(boards_in, 32) && aligned(boards_out, 32) && aligned(rewards, 16) &&
                        aligned(actions, 4) && aligned(flags, 4) && (!spawns || aligned(spawns, 4));
    // Philox: a thread's four envs must be one quad of the stream (one block per thread)
    bool use_lut = plain && vec_ok && s->smem_optin >= kStepSmemBytes && (spawns || (env_offset & 3u) == 0u);
    const int impl = g_impl.load(), tune_threads = g_step_threads.load(), tune_groups = g_step_groups.load();
    if (impl == G2048_IMPL_AUTO) use_lut = use_lut && n >= kLutMinEnvs;
    if (impl == G2048_IMPL_REGISTER || impl == G2048_IMPL_GATHER) use_lut = false;

    if (use_lut) {
        const int smem = (int)kStepSmemBytes;
        const int threads = tune_threads ? tune_threads : 640, groups = tune_groups;  // 5 warps per scheduler at 96 registers
        // one CTA per SM as soon as there is a warp-tile for each (the kernel splits in warp-tiles)
        const size_t wunits = n / ((size_t)128 * groups);
        const int grid = (int)(wunits < (size_t)s->sms ? (wunits ? wunits : 1) : (size_t)s->sms);
        // launched with the programmatic-serialization attribute: back-to-back step launches
        // overlap the next launch's latency and table staging with this launch's tail
        cudaLaunchAttribute pdl[1];
        pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pdl[0].val.programmaticStreamSerializationAllowed = g_pdl;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cfg.attrs = pdl;
        cfg.numAttrs = 1;
#define G2048_LAUNCH_STEP(PH, TH, GR)                                                                     \
    do {                                                                                                  \
        CU(cudaFuncSetAttribute(step_lut_kernel<PH, TH, GR>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                smem));                                                                   \
        cfg.blockDim = dim3(TH);                                                                          \
        CU(cudaLaunchKernelEx(&cfg, step_lut_kernel<PH, TH, GR>, a));                                     \
    } while (0)
#define G2048_LAUNCH_STEP_T(PH)                                          \
    do {                                                                 \
        if (groups == 2) {                                               \
            if (threads == 512) G2048_LAUNCH_STEP(PH, 512, 2);           \
            else if (threads == 768) G2048_LAUNCH_STEP(PH, 768, 2);      \
            else G2048_LAUNCH_STEP(PH, 1024, 2);                         \
        } else {                                                         \
            if (threads == 512) G2048_LAUNCH_STEP(PH, 512, 1);           \
            else if (threads == 640) G2048_LAUNCH_STEP(PH, 640, 1);      \
            else if (threads == 768) G2048_LAUNCH_STEP(PH, 768, 1);      \
            else G2048_LAUNCH_STEP(PH, 1024, 1);                         \
        }                                                                \
    } while (0)
        if (spawns) G2048_LAUNCH_STEP_T(false);
        else G2048_LAUNCH_STEP_T(true);
#undef G2048_LAUNCH_STEP_T
#undef G2048_LAUNCH_STEP
    } else {
        int grid = grid_for(n, 256, s->sms, 8);
        if (impl == G2048_IMPL_REGISTER) {
            if (spawns) step_general_kernel<false, RegisterTable><<<grid, 256, 0, st>>>(a);
            else step_general_kernel<true, RegisterTable><<<grid, 256, 0, st>>>(a);
        } else {
            if (spawns) step_general_kernel<false, GlobalTable><<<grid, 256, 0, st>>>(a);
            else step_general_kernel<true, GlobalTable><<<grid, 256, 0, st>>>(a);
        }
    }
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// GameState step (env.py:79-133), one pass, in place
// ---------------------------------------------------------------------------------------------
struct EnvStepArgs {
    u64* boards;
    int32_t* step_count;
    uint8_t* done;
    float* total_reward;
    const uint8_t* actions;
    const uint8_t* spawns;
    u64 seed, env_offset;
    const u64* env_keys;
    float* rewards;
    uint8_t* flags;
    float* obs;
    int32_t* max_tile;
    uint8_t* legal;
    size_t n;
    const u32* table;
};

__device__ __forceinline__ void write_obs(float* obs, size_t i, Board b) {
    // core.py:371-375: log2(tile)/16 = exponent * 0.0625f, exact in f32
    float4* o = reinterpret_cast<float4*>(obs + i * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const u32 h = (q < 2) ? b.lo : b.hi;
        const u32 row = (h >> ((q & 1) * 16)) & 0xFFFFu;
        o[q] = make_float4((float)(row & 0xFu) * 0.0625f, (float)((row >> 4) & 0xFu) * 0.0625f,
                           (float)((row >> 8) & 0xFu) * 0.0625f, (float)(row >> 12) * 0.0625f);
    }
}

template <bool kPhilox>
__global__ void __launch_bounds__(256) env_step_kernel(const EnvStepArgs a) {
    const GlobalTable T{a.table};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.n;
         i += (size_t)gridDim.x * blockDim.x) {
        const Board b = split(a.boards[i]);
        Board nb = b;
        u32 r = 0, f;
        if (a.done[i]) {  // env.py:103-104
            f = legal_mask(b) | G2048_FLAG_DONE;
        } else {
            const int sc = a.step_count[i];
            transition<GlobalTable, kPhilox>(b, a.actions[i], kPhilox ? 0u : a.spawns[i],
                                             a.env_keys ? a.env_keys[i] : a.seed, a.env_offset + i,
                                             (u32)sc, T, nb, r, f);
            a.boards[i] = join(nb);
            a.step_count[i] = sc + 1;                             // env.py:116
            a.done[i] = (f & G2048_FLAG_DONE) ? 1 : 0;            // env.py:111
            a.total_reward[i] = a.total_reward[i] + (float)r;     // env.py:118
        }
        if (a.rewards) a.rewards[i] = (float)r;
        if (a.flags) a.flags[i] = (uint8_t)f;
        if (a.obs) write_obs(a.obs, i, nb);
        if (a.max_tile) {
            const u32 e = max_exponent(nb);
            a.max_tile[i] = e ? (1 << e) : 0;  // core.py:427 returns the tile value
        }
        if (a.legal)  // bool[4] as bytes: bit k of the mask -> byte k
            reinterpret_cast<u32*>(a.legal)[i] = (f & 1u) | ((f & 2u) << 7) | ((f & 4u) << 14) | ((f & 8u) << 21);
    }
}

extern "C" int g2048_env_step(uint64_t* boards, int32_t* step_count, uint8_t* done,
                              float* total_reward, const uint8_t* actions, const uint8_t* spawns,
                              uint64_t seed, uint64_t env_offset, const uint64_t* env_keys,
                              float* rewards, uint8_t* flags, float* obs, int32_t* max_tile,
                              uint8_t* legal, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards || !step_count || !done || !total_reward || !actions) return G2048_EINVAL;
    if ((obs && !aligned(obs, 16)) || (legal && !aligned(legal, 4))) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    EnvStepArgs a{reinterpret_cast<u64*>(boards), step_count, done, total_reward, actions, spawns,
                  seed, env_offset, reinterpret_cast<const u64*>(env_keys), rewards, flags, obs,
                  max_tile, legal, n, s->table};
    int grid = grid_for(n, 256, s->sms, 8);
    if (spawns) env_step_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else env_step_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// One self-play iteration in one launch: action sampling (core.py:381-424),
// GameState step (env.py:79-133), the next observation / legal mask, and the trajectory record
// (self_play.py:346-368: observation and policy the action was drawn from, action, reward, value).
// ---------------------------------------------------------------------------------------------
struct SelfplayArgs {
    u64* boards;
    int32_t* step_count;
    uint8_t* done;
    float* total_reward;
    const float* policy;        // f32[n*4] search / network policy (core.py:381 `policy`)
    const float* values;        // f32[n] root values to record, or NULL
    const uint8_t* actions_in;  // u8[n]: use these actions instead of sampling, or NULL
    const uint8_t* spawns;      // injected spawn bytes, or NULL (Philox)
    float temperature;
    u64 seed, env_offset;
    // trajectory buffers, time-major [capacity][n] (any may be NULL); row = *t_dev
    const int32_t* t_dev;
    u32 capacity;
    float* tr_obs;
    uint8_t* tr_legal;
    uint8_t* tr_actions;
    float* tr_rewards;
    float* tr_policies;
    float* tr_values;
    uint8_t* tr_active;
    int32_t* lengths;
    // outputs describing the NEW state (any may be NULL)
    float* rewards;
    float* norm_rewards;  // utils/normalize.py:6: log2(reward) / 16, 0 for reward 0
    float* obs;
    uint8_t* legal;
    int32_t* max_tile;
    uint8_t* actions_out;
    float* probs_out;     // f32[n*4]: the distribution the action was drawn from
    size_t n;
    const u32* table;
};

// core.py:401-421 in f32: mask, renormalise (uniform over the legal actions when the mass is below
// 1e-8), then softmax(log(p + 1e-8) / temperature) or the one-hot of the argmax when temperature < 0.01
__device__ __forceinline__ void action_distribution(const float4 pol, u32 mask, float temperature, float out[4]) {
    const float p[4] = {pol.x, pol.y, pol.z, pol.w};
    float m[4], total = 0.0f, n_legal = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const bool legal = (mask >> a) & 1u;
        m[a] = legal ? p[a] : 0.0f;
        total += m[a];
        n_legal += legal ? 1.0f : 0.0f;
    }
    const float inv_legal = 1.0f / fmaxf(n_legal, 1.0f);
#pragma unroll
    for (int a = 0; a < 4; ++a) m[a] = total < 1e-8f ? (((mask >> a) & 1u) ? inv_legal : 0.0f) : m[a] / total;
    if (temperature < 0.01f) {
        int best = 0;  // argmax, first maximum (jnp.argmax)
#pragma unroll
        for (int a = 1; a < 4; ++a)
            if (m[a] > m[best]) best = a;
#pragma unroll
        for (int a = 0; a < 4; ++a) out[a] = a == best ? 1.0f : 0.0f;
        return;
    }
    float z[4], zmax = -INFINITY, sum = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        z[a] = logf(m[a] + 1e-8f) / temperature;
        zmax = fmaxf(zmax, z[a]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        z[a] = expf(z[a] - zmax);
        sum += z[a];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) out[a] = z[a] / sum;
}

// inverse CDF with a 24-bit uniform from the env's word of tag 4; never returns an action of
// probability 0 (rounding in the running sum cannot push the draw past the last positive entry)
__device__ __forceinline__ u32 sample_from(const float q[4], u32 w) {
    const float u = (float)(w >> 8) * (1.0f / 16777216.0f);
    float c = 0.0f;
    u32 last = 0;
#pragma unroll
    for (u32 a = 0; a < 4; ++a) {
        if (q[a] > 0.0f) {
            c += q[a];
            last = a;
            if (u < c) return a;
        }
    }
    return last;
}

// core.py:381 sample_action on its own: policy f32[n*4] + legal u8[n*4] -> action u8[n] (and the
// distribution); the draw of env i is the word of (seed, env_offset + i, step, tag 4), as inside
// g2048_selfplay_step.
__global__ void __launch_bounds__(256) sample_action_kernel(const float* __restrict__ policy,
                                                            const uint8_t* __restrict__ legal, float temperature,
                                                            u64 seed, u64 env_offset, const int32_t* __restrict__ step_ctr,
                                                            u32 step_base, uint8_t* __restrict__ actions,
                                                            float* __restrict__ probs, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u32 l4 = reinterpret_cast<const u32*>(legal)[i];
        const u32 mask = ((l4 & 0xFFu) ? 1u : 0u) | ((l4 & 0xFF00u) ? 2u : 0u) | ((l4 & 0xFF0000u) ? 4u : 0u) |
                         ((l4 & 0xFF000000u) ? 8u : 0u);
        float q[4];
        action_distribution(*reinterpret_cast<const float4*>(policy + i * 4), mask, temperature, q);
        const u32 step = step_ctr ? (u32)step_ctr[i] : step_base;
        if (actions) actions[i] = (uint8_t)sample_from(q, philox_word(seed, env_offset + i, step, G2048_PHILOX_TAG_POLICY));
        if (probs) *reinterpret_cast<float4*>(probs + i * 4) = make_float4(q[0], q[1], q[2], q[3]);
    }
}

extern "C" int g2048_sample_action(const float* policy, const uint8_t* legal, float temperature, uint64_t seed,
                                   uint64_t env_offset, const int32_t* step_ctr, uint32_t step_base,
                                   uint8_t* actions_out, float* probs_out, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!policy || !legal || (!actions_out && !probs_out)) return G2048_EINVAL;
    if (!aligned(policy, 16) || !aligned(legal, 4) || !aligned(probs_out, 16)) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    sample_action_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        policy, legal, temperature, seed, env_offset, step_ctr, step_base, actions_out, probs_out, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

template <bool kPhilox>
__global__ void __launch_bounds__(256) selfplay_step_kernel(const SelfplayArgs a) {
    const GlobalTable T{a.table};
    const u32 t = a.t_dev ? (u32)*a.t_dev : 0u;
    const bool rec = a.t_dev && t < a.capacity;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.n;
         i += (size_t)gridDim.x * blockDim.x) {
        const Board b = split(a.boards[i]);
        const bool was_done = a.done[i] != 0;
        const int sc = a.step_count[i];
        const u32 mask = legal_mask(b);
        const size_t row = (size_t)t * a.n + i;
        // ---- the action (core.py:381-424) ----
        float q[4] = {0.25f, 0.25f, 0.25f, 0.25f};
        u32 action;
        if (a.actions_in) {
            action = a.actions_in[i] & 3u;
        } else {
            action_distribution(*reinterpret_cast<const float4*>(a.policy + i * 4), mask, a.temperature, q);
            action = sample_from(q, philox_word(a.seed, a.env_offset + i, (u32)sc, G2048_PHILOX_TAG_POLICY));
        }
        if (a.actions_out) a.actions_out[i] = (uint8_t)action;
        if (a.probs_out) *reinterpret_cast<float4*>(a.probs_out + i * 4) = make_float4(q[0], q[1], q[2], q[3]);
        // ---- record what the policy saw (self_play.py:346-353) ----
        if (rec) {
            if (a.tr_obs) write_obs(a.tr_obs, row, b);
            if (a.tr_legal)
                reinterpret_cast<u32*>(a.tr_legal)[row] = (mask & 1u) | ((mask & 2u) << 7) | ((mask & 4u) << 14) | ((mask & 8u) << 21);
            if (a.tr_actions) a.tr_actions[row] = (uint8_t)action;
            if (a.tr_policies) {
                const float4 pol = a.policy ? *reinterpret_cast<const float4*>(a.policy + i * 4) : make_float4(q[0], q[1], q[2], q[3]);
                *reinterpret_cast<float4*>(a.tr_policies + row * 4) = pol;
            }
            if (a.tr_values) a.tr_values[row] = a.values ? a.values[i] : 0.0f;
            if (a.tr_active) a.tr_active[row] = was_done ? 0 : 1;
            if (a.lengths && !was_done) a.lengths[i] += 1;
        }
        // ---- the step (env.py:79-133) ----
        Board nb = b;
        u32 r = 0, f;
        if (was_done) {
            f = mask | G2048_FLAG_DONE;
        } else {
            transition<GlobalTable, kPhilox>(b, action, kPhilox ? 0u : a.spawns[i], a.seed, a.env_offset + i, (u32)sc,
                                             T, nb, r, f);
            a.boards[i] = join(nb);
            a.step_count[i] = sc + 1;
            a.done[i] = (f & G2048_FLAG_DONE) ? 1 : 0;
            a.total_reward[i] = a.total_reward[i] + (float)r;
        }
        if (rec && a.tr_rewards) a.tr_rewards[row] = (float)r;
        if (a.rewards) a.rewards[i] = (float)r;
        if (a.norm_rewards) a.norm_rewards[i] = r ? log2f((float)r) * 0.0625f : 0.0f;
        if (a.obs) write_obs(a.obs, i, nb);
        if (a.max_tile) {
            const u32 e = max_exponent(nb);
            a.max_tile[i] = e ? (1 << e) : 0;
        }
        if (a.legal) reinterpret_cast<u32*>(a.legal)[i] = (f & 1u) | ((f & 2u) << 7) | ((f & 4u) << 14) | ((f & 8u) << 21);
    }
}

__global__ void advance_counter_kernel(int32_t* t) { *t += 1; }

extern "C" int g2048_selfplay_step(uint64_t* boards, int32_t* step_count, uint8_t* done, float* total_reward,
                                   const float* policy, const float* values, const uint8_t* actions_in,
                                   const uint8_t* spawns, float temperature, uint64_t seed, uint64_t env_offset,
                                   int32_t* t_dev, uint32_t capacity, float* tr_obs, uint8_t* tr_legal,
                                   uint8_t* tr_actions, float* tr_rewards, float* tr_policies, float* tr_values,
                                   uint8_t* tr_active, int32_t* lengths, float* rewards, float* norm_rewards,
                                   float* obs, uint8_t* legal, int32_t* max_tile, uint8_t* actions_out,
                                   float* probs_out, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards || !step_count || !done || !total_reward) return G2048_EINVAL;
    if (!policy && !actions_in) return G2048_EINVAL;
    if (!aligned(policy, 16) || !aligned(probs_out, 16) || !aligned(obs, 16) || !aligned(tr_obs, 16) ||
        !aligned(tr_policies, 16) || !aligned(legal, 4) || !aligned(tr_legal, 4))
        return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    const SelfplayArgs a{reinterpret_cast<u64*>(boards), step_count, done, total_reward, policy, values, actions_in,
                         spawns, temperature, seed, env_offset, t_dev, capacity, tr_obs, tr_legal, tr_actions,
                         tr_rewards, tr_policies, tr_values, tr_active, lengths, rewards, norm_rewards, obs, legal,
                         max_tile, actions_out, probs_out, n, s->table};
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n, 256, s->sms, 8);
    if (spawns) selfplay_step_kernel<false><<<grid, 256, 0, st>>>(a);
    else selfplay_step_kernel<true><<<grid, 256, 0, st>>>(a);
    if (t_dev) advance_counter_kernel<<<1, 1, 0, st>>>(t_dev);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// mask / observe / pack / unpack / reset
// ---------------------------------------------------------------------------------------------
// four consecutive boards per thread: one 256-bit load and one 32-bit store, enough bytes in
// flight per thread to cover the HBM latency (one board per thread reached 26 % of the peak)
__global__ void __launch_bounds__(256) mask_kernel(const u64* __restrict__ boards,
                                                   uint8_t* __restrict__ flags, size_t n, bool vec) {
    const size_t quads = vec ? n / 4 : 0;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < quads;
         q += (size_t)gridDim.x * blockDim.x) {
        const Boards4 b = ld_boards4(boards + 4 * q);
        u32 f = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // the fused-indicator mask of g2048_fast.cuh (38 instructions)
            const u32 lo = (u32)b.b[k], hi = (u32)(b.b[k] >> 32);
            f |= board_flags(fast::legal_mask(lo, hi, fast::nz8(lo, 1u), fast::nz8(hi, 1u), 1u)) << (8 * k);
        }
        st_u32_stream(flags + 4 * q, f);
    }
    for (size_t i = quads * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        flags[i] = (uint8_t)board_flags(legal_mask(split(boards[i])));
}

extern "C" int g2048_mask(const uint64_t* boards, uint8_t* flags, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards || !flags) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    mask_kernel<<<grid_for((n + 3) / 4, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const u64*>(boards), flags, n, aligned(boards, 32) && aligned(flags, 4));
    CU(cudaGetLastError());
    return G2048_OK;
}

// legal mask as one 0/1 byte per action (a bool[n,4] array: core.py:275, env.py:245) and done bytes
__global__ void __launch_bounds__(256) legal_bool_kernel(const u64* __restrict__ boards, u32* __restrict__ legal,
                                                         uint8_t* __restrict__ done, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 b = boards[i];
        const u32 lo = (u32)b, hi = (u32)(b >> 32);
        const u32 f = fast::legal_mask(lo, hi, fast::nz8(lo, 1u), fast::nz8(hi, 1u), 1u);
        if (legal) legal[i] = (f & 1u) | ((f & 2u) << 7) | ((f & 4u) << 14) | ((f & 8u) << 21);
        if (done) done[i] = f == 0u ? 1 : 0;
    }
}

extern "C" int g2048_legal_actions(const uint64_t* boards, uint8_t* legal, uint8_t* done, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards || (!legal && !done) || !aligned(legal, 4)) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    legal_bool_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const u64*>(boards), reinterpret_cast<u32*>(legal), done, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// Warp-cooperative 16-values-per-board writers (observe, unpack).  A warp takes 32 boards = 2 KiB
// of output; lane l writes the 16-byte chunks l, l + 32, l + 64, l + 96 of it, i.e. row (q & 3) of
// board (q >> 2) for chunk q, so every store instruction covers 512 contiguous bytes.  (One board
// per thread writes 64 bytes at a 64-byte stride per lane: 14-40 % of the HBM peak.)  The board of a
// chunk is re-read through L1 (8 distinct boards per load instruction).
template <class Op>
__device__ __forceinline__ void rows_of_tile(const u64* __restrict__ boards, size_t base, size_t n, u32 lane,
                                             const Op& op) {
#pragma unroll
    for (u32 j = 0; j < 4; ++j) {
        const u32 q = lane + 32u * j;
        const size_t i = base + (q >> 2);
        if (i < n) {
            const u64 b = __ldg(boards + i);
            op(base * 4 + q, (u32)(b >> (16u * (q & 3u))) & 0xFFFFu);   // chunk index, row nibbles
        }
    }
}

__global__ void __launch_bounds__(256) observe_kernel(const u64* __restrict__ boards,
                                                      float* __restrict__ obs,
                                                      uint8_t* __restrict__ max_exp, size_t n) {
    const u32 lane = threadIdx.x & 31u;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t base = warp * 32; base < n; base += n_warps * 32) {
        if (obs)  // core.py:371-375: log2(tile)/16 = exponent * 0.0625f, exact in f32
            rows_of_tile(boards, base, n, lane, [&](size_t chunk, u32 row) {
                reinterpret_cast<float4*>(obs)[chunk] =
                    make_float4((float)(row & 0xFu) * 0.0625f, (float)((row >> 4) & 0xFu) * 0.0625f,
                                (float)((row >> 8) & 0xFu) * 0.0625f, (float)(row >> 12) * 0.0625f);
            });
        if (max_exp && base + lane < n) max_exp[base + lane] = (uint8_t)max_exponent(split(__ldg(boards + base + lane)));
    }
}

extern "C" int g2048_observe(const uint64_t* boards, float* obs, uint8_t* max_exp, size_t n,
                             void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards || (obs && !aligned(obs, 16))) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    observe_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const u64*>(boards), obs, max_exp, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// One thread per 16 OUTPUT bytes (one 128-bit store): the bytes [16 g, 16 g + 16) of the flat
// u8[n * 16 * width] array belong to at most two cells (width >= 16), each of which contributes a
// single 1 at offset log2(tile) (binary.py:47-49 eye(width)[log2]).  (One thread per cell writing
// `width` single bytes reached 2.5 % of the HBM peak.)
template <int kWidth>   // 0 = run-time width
__global__ void __launch_bounds__(256) onehot_kernel(const u64* __restrict__ boards,
                                                     uint8_t* __restrict__ out, int width_rt, size_t n) {
    const u32 width = kWidth ? (u32)kWidth : (u32)width_rt;
    const size_t groups = n * width;   // n * 16 * width / 16
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < groups;
         g += (size_t)gridDim.x * blockDim.x) {
        const size_t first = g * 16;
        const size_t cell = first / width;
        const u32 k0 = (u32)(first - cell * width);   // position of byte 0 inside its cell
        const u32 e0 = (u32)(__ldg(boards + (cell >> 4)) >> (4 * (cell & 15))) & 0xFu;
        u32 w[4] = {0u, 0u, 0u, 0u};
        const u32 p0 = e0 - k0;                        // byte of this group that holds cell's 1
        if (p0 < 16u) w[p0 >> 2] |= 1u << (8u * (p0 & 3u));
        const u32 rest = width - k0;                   // bytes of `cell` in this group (>= 1)
        if (rest < 16u && cell + 1 < n * 16) {         // the group runs into the next cell
            const size_t c1 = cell + 1;
            const u32 e1 = (u32)(__ldg(boards + (c1 >> 4)) >> (4 * (c1 & 15))) & 0xFu;
            const u32 p1 = rest + e1;
            if (p1 < 16u) w[p1 >> 2] |= 1u << (8u * (p1 & 3u));
        }
        st_int4_stream(out + first, make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]));
    }
}

// fallback for an output that is not 16-byte aligned: one thread per (board, cell)
__global__ void __launch_bounds__(256) onehot_bytes_kernel(const u64* __restrict__ boards,
                                                           uint8_t* __restrict__ out, int width, size_t n) {
    const size_t total = n * 16;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const u32 e = (u32)(boards[t >> 4] >> (4 * (t & 15))) & 0xFu;
        uint8_t* row = out + t * width;
        for (int k = 0; k < width; ++k) row[k] = (uint8_t)(k == (int)e);
    }
}

extern "C" int g2048_observe_onehot(const uint64_t* boards, uint8_t* out, int width, size_t n,
                                    void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards || !out || width < 16 || width > 64) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    const u64* b = reinterpret_cast<const u64*>(boards);
    cudaStream_t st = (cudaStream_t)stream;
    if (!aligned(out, 16))
        onehot_bytes_kernel<<<grid_for(n * 16, 256, s->sms, 8), 256, 0, st>>>(b, out, width, n);
    else if (width == 31)   // the reference's encodage_size (envs/twentyfortyeight.py:66): division by a constant
        onehot_kernel<31><<<grid_for(n * 31, 256, s->sms, 8), 256, 0, st>>>(b, out, width, n);
    else
        onehot_kernel<0><<<grid_for(n * (size_t)width, 256, s->sms, 8), 256, 0, st>>>(b, out, width, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// tile value -> exponent; *bad is set for a value that is not 0 or a power of two in [2, 32768]
__device__ __forceinline__ u32 tile_exponent(int32_t t, bool& bad) {
    if (t == 0) return 0u;
    if (t < 2 || t > 32768 || (t & (t - 1)) != 0) {
        bad = true;
        return 0u;
    }
    return 31u - __clz(t);
}

// Warp-cooperative: lane l reads the 16-byte chunks l, l + 32, ... of the warp's 2 KiB of tiles
// (coalesced LDG.128), packs its row, and the four lanes of a board OR their rows together.
__global__ void __launch_bounds__(256) pack_kernel(const int32_t* __restrict__ tiles,
                                                   u64* __restrict__ boards,
                                                   int32_t* __restrict__ status, size_t n, bool vec) {
    const u32 lane = threadIdx.x & 31u;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (size_t base = warp * 32; base < n; base += n_warps * 32) {
        int4 t[4];   // the four loads are issued before any of them is consumed
#pragma unroll
        for (u32 j = 0; j < 4; ++j) {
            const u32 q = lane + 32u * j;
            t[j] = make_int4(0, 0, 0, 0);
            if (base + (q >> 2) < n) {
                if (vec) {
                    t[j] = __ldg(reinterpret_cast<const int4*>(tiles) + base * 4 + q);
                } else {
                    const int32_t* p = tiles + (base * 4 + q) * 4;
                    t[j] = make_int4(p[0], p[1], p[2], p[3]);
                }
            }
        }
#pragma unroll
        for (u32 j = 0; j < 4; ++j) {
            const u32 q = lane + 32u * j;
            const size_t i = base + (q >> 2);
            const u32 row = tile_exponent(t[j].x, bad) | (tile_exponent(t[j].y, bad) << 4) |
                            (tile_exponent(t[j].z, bad) << 8) | (tile_exponent(t[j].w, bad) << 12);
            const u32 r = q & 3u;
            u32 lo = (r < 2u) ? row << (16u * r) : 0u;
            u32 hi = (r < 2u) ? 0u : row << (16u * (r - 2u));
            lo |= __shfl_xor_sync(0xFFFFFFFFu, lo, 1);
            hi |= __shfl_xor_sync(0xFFFFFFFFu, hi, 1);
            lo |= __shfl_xor_sync(0xFFFFFFFFu, lo, 2);
            hi |= __shfl_xor_sync(0xFFFFFFFFu, hi, 2);
            if (i < n && (q & 3u) == 0u) boards[i] = ((u64)hi << 32) | lo;
        }
    }
    if (bad && status) *status = G2048_ERANGE;
}

__global__ void __launch_bounds__(256) unpack_kernel(const u64* __restrict__ boards,
                                                     int32_t* __restrict__ tiles, size_t n, bool vec) {
    const u32 lane = threadIdx.x & 31u;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t base = warp * 32; base < n; base += n_warps * 32)
        rows_of_tile(boards, base, n, lane, [&](size_t chunk, u32 row) {
            const u32 e0 = row & 0xFu, e1 = (row >> 4) & 0xFu, e2 = (row >> 8) & 0xFu, e3 = row >> 12;
            const int4 v = make_int4(e0 ? 1 << e0 : 0, e1 ? 1 << e1 : 0, e2 ? 1 << e2 : 0, e3 ? 1 << e3 : 0);
            if (vec) {
                st_int4_stream(reinterpret_cast<int4*>(tiles) + chunk, v);
            } else {
                int32_t* p = tiles + chunk * 4;
                p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
            }
        });
}

extern "C" int g2048_pack(const int32_t* tiles, uint64_t* boards, int32_t* status, size_t n,
                          void* stream) {
    if (n == 0) return G2048_OK;
    if (!tiles || !boards) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    pack_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        tiles, reinterpret_cast<u64*>(boards), status, n, aligned(tiles, 16));
    CU(cudaGetLastError());
    return G2048_OK;
}

extern "C" int g2048_unpack(const uint64_t* boards, int32_t* tiles, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!tiles || !boards) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    unpack_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const u64*>(boards), tiles, n, aligned(tiles, 16));
    CU(cudaGetLastError());
    return G2048_OK;
}

// two sequential spawns on the empty board (core.py:338-342)
// (Philox: words 0 and 1 of the env's game stream, include/g2048.h "Game stream")
__device__ __forceinline__ Board reset_board(u64 key, u64 env, const uint8_t* spawn2) {
    Board b{0u, 0u};
    Philox4 p{0u, 0u, 0u, 0u};
    if (!spawn2) p = philox_game_block(key, env, 0u);
#pragma unroll
    for (u32 t = 0; t < 2; ++t) {
        u32 zlo, zhi;
        empties(b, zlo, zhi);
        if (spawn2) {
            spawn_byte(b, zlo, zhi, spawn2[t]);
        } else {
            spawn_philox(b, zlo, zhi, t == 0u ? p.x : p.y);
        }
    }
    return b;
}

__global__ void __launch_bounds__(256) reset_kernel(u64 seed, u64 env_offset,
                                                    const u64* __restrict__ env_keys,
                                                    const uint8_t* __restrict__ spawns,
                                                    u64* __restrict__ boards, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        boards[i] = join(reset_board(env_keys ? env_keys[i] : seed, env_offset + i,
                                     spawns ? spawns + 2 * i : nullptr));
}

extern "C" int g2048_reset(uint64_t seed, uint64_t env_offset, const uint64_t* env_keys,
                           const uint8_t* spawns, uint64_t* boards_out, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards_out) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    reset_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        seed, env_offset, reinterpret_cast<const u64*>(env_keys), spawns,
        reinterpret_cast<u64*>(boards_out), n);
    CU(cudaGetLastError());
    return G2048_OK;
}

__global__ void __launch_bounds__(256) spawn_kernel(u64* __restrict__ boards,
                                                    const uint8_t* __restrict__ spawns, u64 seed,
                                                    u64 env_offset, const u64* __restrict__ env_keys,
                                                    u32 step, u32 tag, uint8_t* __restrict__ flags,
                                                    size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        Board b = split(boards[i]);
        u32 zlo, zhi;
        empties(b, zlo, zhi);
        bool ok = true;
        if ((zlo | zhi) != 0u) {  // a full board is returned unchanged (gameboard.py:277-278)
            if (spawns) ok = spawn_byte(b, zlo, zhi, spawns[i]);
            else spawn_philox(b, zlo, zhi, philox_word(env_keys ? env_keys[i] : seed, env_offset + i, step, tag));
            boards[i] = join(b);
        }
        if (flags) flags[i] = ok ? 0 : G2048_FLAG_BADSPAWN;
    }
}

extern "C" int g2048_spawn(uint64_t* boards_io, const uint8_t* spawns, uint64_t seed,
                           uint64_t env_offset, const uint64_t* env_keys, uint32_t step_base,
                           uint32_t tag, uint8_t* flags, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards_io) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    spawn_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<u64*>(boards_io), spawns, seed, env_offset,
        reinterpret_cast<const u64*>(env_keys), step_base, tag, flags, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// MCTS leaf expansion (gameboard.py:105 x4, :132 after_state, gamemove.py:45)
// One thread per board for the fixed-size outputs; children are written by the whole warp,
// lane j producing child j of the current (board, action), so stores are contiguous.
// ---------------------------------------------------------------------------------------------
struct ExpandArgs {
    const u64* boards;
    u64* after;
    int32_t* rewards;
    uint8_t* mask;
    uint8_t* num_empty;
    uint8_t* child_count;
    u64* children;
    const int64_t* child_offsets;
    int layout;
    size_t n;
    const u32* table;
    const u32* table_hashed;
    u32 one;  // run-time 1 (see StepArgs)
    int64_t* offsets_out;  // g2048_expand_compact: the offsets are an OUTPUT (device-side scan)
    size_t capacity;       // children buffer capacity in boards (0 = unlimited): a board whose
                           // children would run past it writes none
    // large-batch g2048_expand_compact: the moves kernel writes the child total of every group of
    // 32 boards, scan_top_kernel turns them into the exclusive prefix of every tile of 1024 boards,
    // and the children kernel derives each board's offset from the two (see scan_top_kernel)
    u32* group_sums;
    const unsigned long long* tile_prefix;
};

__device__ __forceinline__ u32 empty_cells16(Board b) {
    // one bit per cell: gather bit 3 of every nibble of ~nz8
    const u32 zlo = (~nz8(b.lo)) & 0x88888888u, zhi = (~nz8(b.hi)) & 0x88888888u;
    u32 m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        m |= ((zlo >> (4 * i + 3)) & 1u) << i;
        m |= ((zhi >> (4 * i + 3)) & 1u) << (8 + i);
    }
    return m;
}

// Expansion in two kernels.
//   expand_moves_*_kernel   lane = board: the four afterstates through the fast move of
//       g2048_fast.cuh with compile-time orientation words, rewards, empty counts, child counts
//       and the legal mask; all fixed-size outputs are one wide store per lane.  Large batches
//       run it persistent with the hashed row table staged in shared memory (TMA bulk copy, as
//       the step kernel); small ones gather the same table from L2.  A board holding a tile
//       >= 16384 takes the exact first-generation path.
//   expand_children_kernel   warp = board, lane = empty cell of one of the legal afterstates:
//       the (action, rank) of lane j follows from the four counts, the lane builds BOTH children
//       of its cell (tile 2 then 4: gameboard.py:154,162) with the branch-free rank select and
//       writes them with one 128-bit store.  In the compact layout the children of a board are
//       contiguous, so the warp writes one coalesced run of up to 512 bytes per pass.
// A legal move always leaves an empty cell, so child counts are even and compact offsets keep
// the 16-byte alignment; an odd caller-supplied offset falls back to two 64-bit stores.
struct GlobalLut {
    const u32* tab;  // hashed row table in global memory (L2 resident)
    __device__ __forceinline__ u32 operator()(u32 byte_off) const {
        return __ldg(reinterpret_cast<const u32*>(reinterpret_cast<const unsigned char*>(tab) + byte_off));
    }
};

struct MovesOut {
    uint2 after4[4];  // afterstates {lo, hi} per action
    int rew4[4];
    u32 ne4, cc4;     // empties / child count of each afterstate, one byte per action
    u32 flags;        // mask byte (LEGAL | DONE | OVERFLOW bits)
};

template <class Lut>
__device__ __forceinline__ MovesOut expand_moves_compute(const ExpandArgs& a, u64 packed, const Lut& lut) {
    constexpr fast::Orient kOrient[4] = G2048_ORIENT_TABLE;
    const u32 lo = (u32)packed, hi = (u32)(packed >> 32);
    u32 m = 0, ovf = 0, ne4 = 0, cc4 = 0;
    uint2 after4[4];
    int rew4[4];
    if (!fast::has_big_tile(lo, hi)) {
#pragma unroll
        for (u32 act = 0; act < 4; ++act) {
            u32 al, ah;
            bool moved;
            rew4[act] = (int)fast::move_fast(lo, hi, kOrient[act], lut, a.one, al, ah, moved);
            const u32 ne = 16u - __popc(fast::nz8(al, a.one)) - __popc(fast::nz8(ah, a.one));
            m |= moved ? (1u << act) : 0u;
            after4[act] = make_uint2(al, ah);
            ne4 |= ne << (8 * act);
            cc4 |= (moved ? 2u * ne : 0u) << (8 * act);  // a move always leaves an empty cell
        }
    } else {
        const GlobalTable T{a.table};
        const Board b{lo, hi};
        const Board bt = transpose(b);
#pragma unroll
        for (u32 act = 0; act < 4; ++act) {
            Board x = (act & 1u) ? bt : b;
            if (act & 2u) x = reverse_rows(x);
            MoveResult r = move_left(x, T);
            if (act & 2u) r.after = reverse_rows(r.after);
            if (act & 1u) r.after = transpose(r.after);
            const bool moved = (r.after.lo != b.lo) | (r.after.hi != b.hi);
            m |= moved ? (1u << act) : 0u;
            ovf |= r.overflow ? 1u : 0u;
            const u32 ne = 16u - __popc(nz8(r.after.lo)) - __popc(nz8(r.after.hi));
            after4[act] = make_uint2(r.after.lo, r.after.hi);
            rew4[act] = (int)r.score;
            ne4 |= ne << (8 * act);
            cc4 |= (moved ? (ne ? 2u * ne : 1u) : 0u) << (8 * act);
        }
    }
    MovesOut r;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        r.after4[k] = after4[k];
        r.rew4[k] = rew4[k];
    }
    r.ne4 = ne4;
    r.cc4 = cc4;
    r.flags = board_flags(m) | (ovf ? G2048_FLAG_OVERFLOW : 0u);
    return r;
}

// the fixed-size outputs of board i, one wide store per field (NULL pointers are skipped)
__device__ __forceinline__ void expand_moves_store(const ExpandArgs& a, size_t i, const MovesOut& r) {
    if (a.after) {
        uint4* o = reinterpret_cast<uint4*>(a.after + i * 4);
        o[0] = make_uint4(r.after4[0].x, r.after4[0].y, r.after4[1].x, r.after4[1].y);
        o[1] = make_uint4(r.after4[2].x, r.after4[2].y, r.after4[3].x, r.after4[3].y);
    }
    if (a.rewards) *reinterpret_cast<int4*>(a.rewards + i * 4) = make_int4(r.rew4[0], r.rew4[1], r.rew4[2], r.rew4[3]);
    if (a.num_empty) *reinterpret_cast<u32*>(a.num_empty + i * 4) = r.ne4;
    if (a.child_count) *reinterpret_cast<u32*>(a.child_count + i * 4) = r.cc4;
    if (a.mask) a.mask[i] = (uint8_t)r.flags;
}

template <class Lut>
__device__ __forceinline__ void expand_moves_board(const ExpandArgs& a, size_t i, const Lut& lut) {
    expand_moves_store(a, i, expand_moves_compute(a, a.boards[i], lut));
}

__device__ __forceinline__ u32 sum_bytes4(u32 x) { return (x * 0x01010101u) >> 24; }  // four counts <= 30

// persistent, hashed table staged in shared memory.  A warp takes 32 consecutive boards per trip;
// with `group_sums` it also writes their child total (the first level of the device-side scan of
// g2048_expand_compact: five shuffles here replace a kernel that re-read every child count).
__global__ void __launch_bounds__(1024, 1) expand_moves_smem_kernel(const ExpandArgs a) {
    u32* smem_tab = smem_words;
    u64* bar = reinterpret_cast<u64*>(smem_words + kSmemRows);
    table_stage_begin(smem_tab, bar, a.table_hashed);
    table_stage_wait(bar);
    const SmemLut lut{};
    const u32 lane = threadIdx.x & 31u;
    for (size_t base = blockIdx.x * (size_t)blockDim.x + threadIdx.x - lane; base < a.n;
         base += (size_t)gridDim.x * blockDim.x) {
        const size_t i = base + lane;
        u32 tot = 0u;
        if (i < a.n) {
            const MovesOut r = expand_moves_compute(a, a.boards[i], lut);
            expand_moves_store(a, i, r);
            tot = sum_bytes4(r.cc4);
        }
        if (a.group_sums) {  // kernel argument: warp-uniform
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
            if (lane == 0u) a.group_sums[base >> 5] = tot;
        }
    }
}

// small batches: the same table gathered from L2
__global__ void __launch_bounds__(256) expand_moves_gather_kernel(const ExpandArgs a) {
    const GlobalLut lut{a.table_hashed};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x)
        expand_moves_board(a, i, lut);
}

constexpr int kChildWarps = 8;

// One warp's record of a tile of 32 boards in shared memory (lane = board while it is written).
struct ChildTile {
    uint4 rec[32][4];  // {after.lo, after.hi, n_lo, n_hi} per action
    uint4 meta[32];    // {counts, first pair of each action, offset lo, offset hi}
};

__device__ __forceinline__ void expand_record(ChildTile& t, u32 lane, uint4 s01, uint4 s23, u32 cc4,
                                              unsigned long long off) {
    t.rec[lane][0] = make_uint4(s01.x, s01.y, fast::nz8(s01.x, 1u), fast::nz8(s01.y, 1u));
    t.rec[lane][1] = make_uint4(s01.z, s01.w, fast::nz8(s01.z, 1u), fast::nz8(s01.w, 1u));
    t.rec[lane][2] = make_uint4(s23.x, s23.y, fast::nz8(s23.x, 1u), fast::nz8(s23.y, 1u));
    t.rec[lane][3] = make_uint4(s23.z, s23.w, fast::nz8(s23.z, 1u), fast::nz8(s23.w, 1u));
    // pairs (= empty cells) in front of each action's segment, one byte per action
    const u32 h = (cc4 >> 1) & 0x7F7F7F7Fu;
    const u32 first4 = (h << 8) + (h << 16) + (h << 24);
    t.meta[lane] = make_uint4(cc4, first4, (u32)off, (u32)(off >> 32));
}

// The children of the first n_here boards of a tile, written by the whole warp board after board.
// A lane serves a FIXED (action, rank) pair — action = lane / 8, rank = lane % 8 (+ 8 in the
// second pass) — so no per-child search for the segment is needed and the record is read once
// per board.
template <int kLayout, class Tile>
__device__ __forceinline__ void expand_children_tile(const ExpandArgs& a, const Tile& t, u32 n_here,
                                                     u32 lane) {
    const u32 act = lane >> 3, r0 = lane & 7u, byte_sel = 0x4440u + act;
    for (u32 k = 0; k < n_here; ++k) {
        const uint4 mt = t.meta[k];  // warp-uniform
        const u32 cc4 = mt.x;
        // dense layout: every one of the board's 128 slots is written (children, then zeros), so
        // the output needs no memset and DRAM sees full lines instead of partial 256-byte slots
        if (kLayout == G2048_CHILDREN_COMPACT && cc4 == 0u) continue;
        u64* const board_dst = a.children + (((unsigned long long)mt.w << 32) | mt.z);
        if (((cc4 & 0x01010101u) | (kLayout == G2048_CHILDREN_COMPACT ? (mt.z & 1u) : 0u)) == 0u) {
            // even counts, 16-byte aligned destination: the common case
            const uint4 rec = t.rec[k][act];
            const u32 q = fast::prmt(cc4, 0u, byte_sel) >> 1;  // empty cells of my action
            const u32 slot0 = (kLayout == G2048_CHILDREN_COMPACT) ? 2u * fast::prmt(mt.y, 0u, byte_sel) : 32u * act;
            const u32 tiles_lo = __popc(rec.z);
            const bool second = ((cc4 + 0x6F6F6F6Fu) & 0x80808080u) != 0u;  // some count > 16
#pragma unroll
            for (u32 pass = 0; pass < 2u; ++pass) {
                if (kLayout == G2048_CHILDREN_COMPACT && pass == 1u && !second) break;  // warp-uniform
                const u32 rank = r0 + 8u * pass;
                if (kLayout == G2048_CHILDREN_DENSE && rank >= q) {
                    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(
                                     board_dst + slot0 + 2u * rank), "r"(0u)
                                 : "memory");
                } else if (rank < q) {
                    bool in_hi;
                    const u32 bit3 = fast::spawn_marker(rec.z, rec.w, tiles_lo, rank, 1u, in_hi);
                    const u32 tl1 = bit3 >> 3;           // exponent 1 (tile 2) in the target nibble
                    const u32 tl = in_hi ? 0u : tl1, th = in_hi ? tl1 : 0u;
                    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(
                                     board_dst + slot0 + 2u * rank),
                                 "r"(rec.x + tl), "r"(rec.y + th), "r"(rec.x + 2u * tl), "r"(rec.y + 2u * th)
                                 : "memory");
                }
            }
        } else {
            if (kLayout == G2048_CHILDREN_DENSE) {  // zero the board's slots first
                asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(board_dst + 2u * lane), "r"(0u) : "memory");
                asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(board_dst + 64u + 2u * lane), "r"(0u) : "memory");
                __syncwarp();
            }
            if (lane < 4u) {
                // Odd counts or an odd caller-supplied offset: one lane per action, serially.  A
                // count of 1 (an afterstate without an empty cell) cannot follow a move; the exact
                // path keeps the reference's convention (gameboard.py:158-159) for completeness.
                const u32 cc = (cc4 >> (8u * lane)) & 0xFFu;
                u32 before = 0;
                for (u32 q = 0; q < lane; ++q) before += (cc4 >> (8u * q)) & 0xFFu;
                u64* dst = board_dst + ((kLayout == G2048_CHILDREN_COMPACT) ? before : 32u * lane);
                const uint4 rec = t.rec[k][lane];
                const u64 st = ((u64)rec.y << 32) | rec.x;
                if (cc == 1u) {
                    *dst = st;
                } else {
                    u32 q = 0;
                    for (u32 cell = 0; cell < 16u && q < cc; ++cell) {
                        if ((st >> (4u * cell)) & 0xFull) continue;
                        dst[q++] = st | (1ull << (4u * cell));
                        dst[q++] = st | (2ull << (4u * cell));
                    }
                }
            }
        }
    }
}

template <int kLayout>
__global__ void __launch_bounds__(kChildWarps * 32) expand_children_kernel(const ExpandArgs a) {
    // per warp: the records of a tile of 32 boards, loaded coalesced once (lane = board), then
    // read back by the whole warp board after board
    __shared__ ChildTile s_tile[kChildWarps];
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    // The records of the NEXT tile are loaded into registers while the current one is expanded;
    // whatever consumes the loaded values (the offset scan, the capacity test) runs one trip later,
    // when they have arrived.
    struct TileRegs {
        uint4 s01, s23;
        u32 cc4, gsum;
        unsigned long long off;
    };
    auto load = [&](size_t tile_base) {
        TileRegs r;
        r.s01 = r.s23 = make_uint4(0u, 0u, 0u, 0u);
        r.cc4 = r.gsum = 0u;
        r.off = 0ull;
        const size_t mine = tile_base + lane;
        if (mine < a.n) {
            r.cc4 = *reinterpret_cast<const u32*>(a.child_count + mine * 4);
            const uint4* src = reinterpret_cast<const uint4*>(a.after + mine * 4);
            r.s01 = src[0];
            r.s23 = src[1];
            if (kLayout != G2048_CHILDREN_COMPACT) r.off = (unsigned long long)mine * 128ull;
            else if (!a.tile_prefix) r.off = (unsigned long long)a.child_offsets[mine];
        }
        if (kLayout == G2048_CHILDREN_COMPACT && a.tile_prefix && tile_base < a.n) {
            // device-side offsets: prefix of my tile of 1024 boards + the groups of 32 boards in
            // front of mine inside it (one coalesced load, lane = group) + a scan over my 32 boards
            const size_t g = tile_base >> 5;
            r.off = a.tile_prefix[g >> 5];
            if (lane < (u32)(g & 31u)) r.gsum = a.group_sums[(g & ~(size_t)31) + lane];
        }
        return r;
    };
    auto finish = [&](TileRegs& r, size_t tile_base) {
        if (kLayout != G2048_CHILDREN_COMPACT) return;
        const u32 tot = sum_bytes4(r.cc4);
        if (a.tile_prefix) {
            const u32 FULL = 0xFFFFFFFFu;
            u32 gp = r.gsum, inc = tot;
#pragma unroll
            for (int o = 16; o; o >>= 1) gp += __shfl_xor_sync(FULL, gp, o);
#pragma unroll
            for (u32 d = 1; d < 32u; d <<= 1) {
                const u32 v = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += v;
            }
            r.off += (unsigned long long)gp + (inc - tot);
            if (a.offsets_out && tile_base + lane < a.n) a.offsets_out[tile_base + lane] = (int64_t)r.off;
        }
        if (a.capacity != 0 && r.off + tot > a.capacity) r.cc4 = 0u;  // does not fit: nothing is written for this board
    };
    TileRegs cur = load(warp * 32);
    for (size_t tile = warp * 32; tile < a.n; tile += n_warps * 32) {
        finish(cur, tile);
        expand_record(s_tile[w], lane, cur.s01, cur.s23, cur.cc4, cur.off);
        cur = load(tile + n_warps * 32);  // in flight during the expansion below
        __syncwarp();
        expand_children_tile<kLayout>(a, s_tile[w], (u32)min((size_t)32, a.n - tile), lane);
        __syncwarp();
    }
}

// the second kernel of the expansion: one wave, as many blocks as are resident at once (a partial
// second wave costs ~14 %)
static int device_occupancy(DeviceState& s) {
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &occ, expand_children_kernel<G2048_CHILDREN_COMPACT>, kChildWarps * 32, 0));
    s.children_occupancy = occ < 1 ? 1 : occ;
    return G2048_OK;
}

static int launch_expand_children(const ExpandArgs& a, const DeviceState* s, cudaStream_t st) {
    const int grid = grid_for(a.n, kChildWarps * 32, s->sms, s->children_occupancy);
    if (a.layout == G2048_CHILDREN_COMPACT)
        expand_children_kernel<G2048_CHILDREN_COMPACT><<<grid, kChildWarps * 32, 0, st>>>(a);
    else
        expand_children_kernel<G2048_CHILDREN_DENSE><<<grid, kChildWarps * 32, 0, st>>>(a);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// compact expansion without a host round trip (g2048_expand_compact): the child offsets are an
// exclusive scan of the per-board child totals, computed on the device.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kAggValid = 1ull << 63;
constexpr unsigned long long kGroupArrival = 1ull << 48;  // group word: arrivals << 48 | sum of the child totals

// The aggregate and its valid bit travel in ONE 64-bit word, so relaxed accesses are enough: a
// release store would first wait for the CTA's earlier output stores to become visible, and an
// acquire load fences every poll (measured: 8 us of a 23 us launch at 65 536 boards).
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Small batches (n < 2^17): moves and children in ONE kernel, lane = (board, ACTION).  The first
// generation (lane = board, removed) kept only ~3.5 warps per scheduler busy at 65 536 boards (2048
// warps on 592 schedulers) and every lane walked four dependent table gathers: ncu showed 26-32 %
// issue activity and 18-23 us per launch, all of it latency
// (profiles/r02_expand_fused_before_ncu_summary.txt).  With one (board, action) per lane there are
// four times as many warps, each with a quarter of the chain (4096 boards: 7.4 -> 4.1 us, 16 384:
// 8.4 -> 5.2, 65 536: 11.2 -> 10.5 with given offsets): a warp owns 8
// boards, computes their 32 moves at once (per-lane orientation words from shared memory, as the
// step kernel), combines the four lanes of a board with two shuffles, and expands the 8 boards
// from shared memory.  kScan adds the device-side offsets of g2048_expand_compact: CTA t owns tile
// t of 64 boards, publishes its child total as ONE relaxed 64-bit word (value | valid bit) and
// sums the words of all earlier tiles (thread blocks are dispatched in blockIdx order, so a tile
// only waits for tiles that started before it — the assumption CUB's look-back makes as well).
constexpr int kQuadWarps = 8, kQuadBoards = 8 * kQuadWarps;  // 8 boards per warp, 64 per CTA

struct QuadTile {
    uint4 rec[8][4];  // {after.lo, after.hi, n_lo, n_hi} per board and action
    uint4 meta[8];    // {counts, first pair of each action, offset lo, offset hi}
};

template <int kLayout, bool kScan>
__global__ void __launch_bounds__(kQuadWarps * 32) expand_quad_kernel(const ExpandArgs a, unsigned long long* agg,
                                                                    unsigned long long* total_out) {
    __shared__ QuadTile s_tile[kQuadWarps];
    __shared__ __align__(16) uint4 s_orient[8];
    __shared__ u32 s_warp_tot[kQuadWarps];
    __shared__ unsigned long long s_red[kQuadWarps];
    const u32 FULL = 0xFFFFFFFFu, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32 b = lane >> 2, act = lane & 3u;
    if (threadIdx.x < 32) reinterpret_cast<u32*>(s_orient)[threadIdx.x] = reinterpret_cast<const u32*>(c_orient)[threadIdx.x];
    __syncthreads();
    const GlobalLut lut{a.table_hashed};
    for (size_t tile = blockIdx.x; tile * kQuadBoards < a.n; tile += gridDim.x) {  // kScan: one tile per CTA
        const size_t warp_base = tile * kQuadBoards + w * 8, mine = warp_base + b, ia = mine * 4 + act;
        const bool have = mine < a.n;
        u32 al = 0u, ah = 0u, cc = 0u, flags = 0u;  // flags: bit act = moved, bit 8 = overflow
        if (have) {
            const u64 packed = a.boards[mine];
            const u32 lo = (u32)packed, hi = (u32)(packed >> 32);
            u32 ne;
            int reward;
            if (!fast::has_big_tile(lo, hi)) {
                const uint4 o0 = s_orient[2u * act];
                fast::Orient o;
                o.mask = o0.x; o.mul_hi = o0.y; o.mul_lo = o0.z; o.sel_lo = o0.w;
                o.sel_hi = s_orient[2u * act + 1u].x;
                bool moved;
                reward = (int)fast::move_fast(lo, hi, o, lut, a.one, al, ah, moved);
                ne = 16u - __popc(fast::nz8(al, a.one)) - __popc(fast::nz8(ah, a.one));
                cc = moved ? 2u * ne : 0u;  // a move always leaves an empty cell
                flags = moved ? 1u << act : 0u;
            } else {  // exact path: a tile >= 16384 on the board
                const MoveResult r = move_action(Board{lo, hi}, act, GlobalTable{a.table});
                al = r.after.lo;
                ah = r.after.hi;
                const bool moved = (al != lo) | (ah != hi);
                reward = (int)r.score;
                ne = 16u - __popc(nz8(al)) - __popc(nz8(ah));
                cc = moved ? (ne ? 2u * ne : 1u) : 0u;
                flags = (moved ? 1u << act : 0u) | (r.overflow ? 0x100u : 0u);
            }
            if (a.after) a.after[ia] = ((u64)ah << 32) | al;
            if (a.rewards) a.rewards[ia] = reward;
            if (a.num_empty) a.num_empty[ia] = (uint8_t)ne;
            if (a.child_count) a.child_count[ia] = (uint8_t)cc;
        }
        // the four lanes of a board: counts packed one byte per action, flags or-ed
        u32 cc4 = cc << (8u * act);
        cc4 += __shfl_xor_sync(FULL, cc4, 1);
        cc4 += __shfl_xor_sync(FULL, cc4, 2);
        flags |= __shfl_xor_sync(FULL, flags, 1);
        flags |= __shfl_xor_sync(FULL, flags, 2);
        if (have && act == 0u && a.mask)
            a.mask[mine] = (uint8_t)(board_flags(flags & 0xFu) | ((flags & 0x100u) ? G2048_FLAG_OVERFLOW : 0u));
        const u32 tot = sum_bytes4(cc4);
        unsigned long long off;
        if (kScan) {
            u32 inc = act == 0u ? tot : 0u;  // inclusive scan over the warp; one entry per board
#pragma unroll
            for (u32 d = 1; d < 32u; d <<= 1) {
                const u32 v = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += v;
            }
            const u32 excl = __shfl_sync(FULL, inc, lane & ~3u) - tot;  // of my board, inside the warp
            if (lane == 31u) s_warp_tot[w] = inc;
            __syncthreads();
            u32 warp_off = 0u, cta_tot = 0u;
#pragma unroll
            for (u32 k = 0; k < (u32)kQuadWarps; ++k) {
                warp_off += k < w ? s_warp_tot[k] : 0u;
                cta_tot += s_warp_tot[k];
            }
            // Two-level publish: the tile's own word (value | valid bit) and an atomic add of
            // (value + one arrival) to the word of its GROUP of 32 tiles.  A tile then needs the
            // complete groups before its own (<= n_tiles / 32 words) and the earlier tiles of its
            // group (<= 31 words): ~60 polled words per tile instead of every earlier tile's word
            // (the all-predecessors form made 1024 tiles x 256 threads hammer the same 256 L2
            // sectors: +13 us at 65 536 boards).  Warp 0 polls, one word per lane and round.
            const size_t n_tiles = gridDim.x, grp = tile >> 5;
            unsigned long long* grp_word = agg + n_tiles;
            if (threadIdx.x == 0) {
                st_relaxed_u64(&agg[tile], kAggValid | cta_tot);
                atomicAdd(&grp_word[grp], (unsigned long long)cta_tot + kGroupArrival);
            }
            if (w == 0u) {
                unsigned long long part = 0ull;
                for (size_t gg = lane; gg < grp; gg += 32) {  // complete groups in front of mine (all hold 32 tiles)
                    unsigned long long v;
                    do v = ld_relaxed_u64(&grp_word[gg]); while ((v >> 48) != 32ull);
                    part += v & (kGroupArrival - 1ull);
                }
                if (lane < (u32)(tile & 31u)) {  // earlier tiles of my group
                    unsigned long long v;
                    do v = ld_relaxed_u64(&agg[(grp << 5) + lane]); while (!(v & kAggValid));
                    part += v & ~kAggValid;
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                if (lane == 0u) s_red[0] = part;
            }
            __syncthreads();
            const unsigned long long prefix = s_red[0];
            off = prefix + warp_off + excl;
            if (have && act == 0u && a.offsets_out) a.offsets_out[mine] = (int64_t)off;
            if (threadIdx.x == 0 && (tile + 1) * (size_t)kQuadBoards >= a.n && total_out) {  // the last tile
                total_out[0] = prefix + cta_tot;
                total_out[1] = (a.capacity != 0 && prefix + cta_tot > a.capacity) ? 1ull : 0ull;
            }
            if (a.capacity != 0 && off + tot > a.capacity) cc4 = 0u;  // does not fit: nothing is written
        } else {
            off = !have ? 0ull
                        : (kLayout == G2048_CHILDREN_COMPACT ? (unsigned long long)a.child_offsets[mine]
                                                             : (unsigned long long)mine * 128ull);
        }
        QuadTile& t = s_tile[w];
        t.rec[b][act] = make_uint4(al, ah, fast::nz8(al, 1u), fast::nz8(ah, 1u));
        if (act == 0u) {
            const u32 h = (cc4 >> 1) & 0x7F7F7F7Fu;  // pairs (= empty cells) in front of each action's segment
            t.meta[b] = make_uint4(cc4, (h << 8) + (h << 16) + (h << 24), (u32)off, (u32)(off >> 32));
        }
        __syncwarp();
        if (warp_base < a.n)
            expand_children_tile<kLayout>(a, t, (u32)min((size_t)8, a.n - warp_base), lane);
        __syncwarp();
    }
}

// Large batches: moves kernel (writes child_count and the child total of every group of 32 boards),
// scan_top_kernel (prefix of every tile of 1024 boards), children kernel (derives the offsets).
// exclusive scan of one value per thread over the CTA; returns the CTA total through `total`
template <int kThreads>
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* s_warp,
                                                                   unsigned long long& total) {
    const u32 FULL = 0xFFFFFFFFu, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (u32 d = 1; d < 32u; d <<= 1) {
        const unsigned long long x = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += x;
    }
    if (lane == 31u) s_warp[w] = inc;
    __syncthreads();
    unsigned long long warp_off = 0ull;
    total = 0ull;
    for (u32 k = 0; k < (u32)(kThreads / 32); ++k) {
        warp_off += k < w ? s_warp[k] : 0ull;
        total += s_warp[k];
    }
    __syncthreads();  // s_warp may be reused by the caller's next scan
    return warp_off + inc - v;
}

// One CTA: the exclusive prefix of every tile of 1024 boards from the group sums the moves kernel
// wrote (32 groups of 32 boards per tile, 128 bytes per thread), chunk by chunk with a running
// carry — 128 KiB of input per million boards.  The children kernel adds the groups in front of a
// warp's own group and a warp scan over its 32 boards, so no per-board pass is left between the
// moves and the children (round 2 started with three scan kernels — tile sums, top, apply —
// and a second and third read of the child counts: 40 -> 32 us at 2^17 boards, 150 -> 145 us at 2^20).
__global__ void __launch_bounds__(1024) scan_top_kernel(const u32* __restrict__ group_sums, size_t groups,
                                                        unsigned long long* __restrict__ tile_prefix, size_t tiles,
                                                        size_t capacity, unsigned long long* __restrict__ total_out) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long carry = 0ull;
    for (size_t base = 0; base < tiles; base += 1024) {
        const size_t j = base + threadIdx.x;
        unsigned long long v = 0ull;
        if (j < tiles) {
            const size_t g0 = j * 32;
            if (g0 + 32 <= groups) {
                const uint4* p = reinterpret_cast<const uint4*>(group_sums + g0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint4 q = p[k];
                    v += (unsigned long long)q.x + q.y + q.z + q.w;
                }
            } else {
                for (size_t g = g0; g < groups; ++g) v += group_sums[g];
            }
        }
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan<1024>(v, s_warp, total);
        if (j < tiles) tile_prefix[j] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) {
        total_out[0] = carry;
        total_out[1] = (capacity != 0 && carry > capacity) ? 1ull : 0ull;
    }
}

constexpr size_t kExpandSmemMinBoards = (size_t)1 << 17;

extern "C" int g2048_expand(const uint64_t* boards, uint64_t* after, int32_t* rewards,
                            uint8_t* mask, uint8_t* num_empty, uint8_t* child_count,
                            uint64_t* children, const int64_t* child_offsets, int child_layout,
                            size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards) return G2048_EINVAL;
    if (child_layout < G2048_CHILDREN_NONE || child_layout > G2048_CHILDREN_COMPACT) return G2048_EINVAL;
    if (child_layout != G2048_CHILDREN_NONE && !children) return G2048_EINVAL;
    if (child_layout == G2048_CHILDREN_COMPACT && !child_offsets) return G2048_EINVAL;
    if (!aligned(after, 16) || !aligned(rewards, 16) || !aligned(num_empty, 4) || !aligned(child_count, 4) ||
        !aligned(children, 16))
        return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (child_layout != G2048_CHILDREN_NONE && n < kExpandSmemMinBoards) {
        // small batch: one fused kernel, the records never leave the SM (no scratch either)
        const ExpandArgs a{reinterpret_cast<const u64*>(boards), reinterpret_cast<u64*>(after), rewards, mask,
                           num_empty, child_count, reinterpret_cast<u64*>(children), child_offsets,
                           child_layout, n, s->table, s->table_hashed, 1u};
        const size_t tiles = (n + kQuadBoards - 1) / kQuadBoards;
        const int grid = (int)(tiles < (size_t)s->sms * 8 ? tiles : (size_t)s->sms * 8);
        if (child_layout == G2048_CHILDREN_COMPACT)
            expand_quad_kernel<G2048_CHILDREN_COMPACT, false><<<grid, kQuadWarps * 32, 0, st>>>(a, nullptr, nullptr);
        else
            expand_quad_kernel<G2048_CHILDREN_DENSE, false><<<grid, kQuadWarps * 32, 0, st>>>(a, nullptr, nullptr);
        CU(cudaGetLastError());
        return G2048_OK;
    }
    // the children kernel reads the afterstates and counts back: stream-ordered scratch when the
    // caller does not want them
    void* scratch = nullptr;
    u64* after_buf = reinterpret_cast<u64*>(after);
    uint8_t* count_buf = child_count;
    if (!after_buf || !count_buf) {
        CU(cudaMallocFromPoolAsync(&scratch, n * 36, s->pool, st));
        if (!after_buf) after_buf = reinterpret_cast<u64*>(scratch);
        if (!count_buf) count_buf = reinterpret_cast<uint8_t*>(scratch) + n * 32;
    }
    ExpandArgs a{reinterpret_cast<const u64*>(boards), after_buf, rewards, mask,
                 num_empty, count_buf, reinterpret_cast<u64*>(children), child_offsets,
                 child_layout, n, s->table, s->table_hashed, 1u};
    // the launches in a helper so that the scratch buffer is released on every path
    auto launch = [&]() -> int {
        if (n >= kExpandSmemMinBoards) {
            CU(cudaFuncSetAttribute(expand_moves_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSmemTableBytes + 64u)));
            expand_moves_smem_kernel<<<s->sms, 1024, kSmemTableBytes + 64u, st>>>(a);
        } else {
            expand_moves_gather_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, st>>>(a);
        }
        CU(cudaGetLastError());
        if (child_layout == G2048_CHILDREN_NONE) return G2048_OK;
        return launch_expand_children(a, s, st);
    };
    rc = launch();
    if (scratch) {
        const cudaError_t fe = cudaFreeAsync(scratch, st);
        if (rc == G2048_OK && fe != cudaSuccess) return cuda_fail(fe, "cudaFreeAsync(scratch)");
    }
    return rc;
}

extern "C" int g2048_expand_children(const uint64_t* after, const uint8_t* child_count,
                                     const int64_t* child_offsets, uint64_t* children,
                                     int child_layout, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!after || !child_count || !children) return G2048_EINVAL;
    if (child_layout != G2048_CHILDREN_DENSE && child_layout != G2048_CHILDREN_COMPACT) return G2048_EINVAL;
    if (child_layout == G2048_CHILDREN_COMPACT && !child_offsets) return G2048_EINVAL;
    if (!aligned(after, 16) || !aligned(child_count, 4) || !aligned(children, 16)) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    ExpandArgs a{nullptr, const_cast<u64*>(reinterpret_cast<const u64*>(after)), nullptr, nullptr, nullptr,
                 const_cast<uint8_t*>(child_count), reinterpret_cast<u64*>(children), child_offsets,
                 child_layout, n, s->table, s->table_hashed, 1u};
    return launch_expand_children(a, s, (cudaStream_t)stream);
}

extern "C" int g2048_expand_compact(const uint64_t* boards, uint64_t* after, int32_t* rewards,
                                    uint8_t* mask, uint8_t* num_empty, uint8_t* child_count,
                                    int64_t* child_offsets, uint64_t* children, size_t capacity,
                                    uint64_t* total_out, size_t n, void* stream) {
    if (n == 0) {
        if (total_out) CU(cudaMemsetAsync(total_out, 0, 16, (cudaStream_t)stream));
        return G2048_OK;
    }
    if (!boards || !children || !child_offsets || capacity == 0) return G2048_EINVAL;
    if (!aligned(after, 16) || !aligned(rewards, 16) || !aligned(num_empty, 4) || !aligned(child_count, 4) ||
        !aligned(children, 16) || !aligned(child_offsets, 8) || !aligned(total_out, 8))
        return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(total_out);
    if (n < kExpandSmemMinBoards) {
        const size_t tiles = (n + kQuadBoards - 1) / kQuadBoards;
        unsigned long long* scratch = nullptr;
        CU(cudaMallocFromPoolAsync(&scratch, (tiles + tiles / 32 + 4) * sizeof(unsigned long long), s->pool, st));
        auto launch = [&]() -> int {
            CU(cudaMemsetAsync(scratch, 0, (tiles + tiles / 32 + 4) * sizeof(unsigned long long), st));
            const ExpandArgs a{reinterpret_cast<const u64*>(boards), reinterpret_cast<u64*>(after), rewards, mask,
                               num_empty, child_count, reinterpret_cast<u64*>(children), nullptr,
                               G2048_CHILDREN_COMPACT, n, s->table, s->table_hashed, 1u, child_offsets, capacity};
            expand_quad_kernel<G2048_CHILDREN_COMPACT, true><<<(unsigned)tiles, kQuadWarps * 32, 0, st>>>(a, scratch + 2, tot);
            CU(cudaGetLastError());
            return G2048_OK;
        };
        rc = launch();
        const cudaError_t fe = cudaFreeAsync(scratch, st);
        if (rc == G2048_OK && fe != cudaSuccess) return cuda_fail(fe, "cudaFreeAsync(scratch)");
        return rc;
    }
    // large batch: moves (+ group sums) -> prefix of the tiles -> children (+ offsets)
    const size_t groups = (n + 31) / 32, tiles = (groups + 31) / 32;
    unsigned char* scratch = nullptr;  // [tile prefix | group sums | afterstates | child counts] as needed
    const size_t prefix_bytes = (tiles * 8 + 15) / 16 * 16, sums_bytes = (groups * 4 + 15) / 16 * 16;
    const size_t need = prefix_bytes + sums_bytes + (after ? 0 : n * 32) + (child_count ? 0 : n * 4);
    CU(cudaMallocFromPoolAsync(&scratch, need, s->pool, st));
    auto launch = [&]() -> int {
        unsigned long long* prefix = reinterpret_cast<unsigned long long*>(scratch);
        u32* sums = reinterpret_cast<u32*>(scratch + prefix_bytes);
        unsigned char* p = scratch + prefix_bytes + sums_bytes;
        u64* after_buf = after ? reinterpret_cast<u64*>(after) : reinterpret_cast<u64*>(p);
        uint8_t* count_buf = child_count ? child_count : p + (after ? 0 : n * 32);
        ExpandArgs a{reinterpret_cast<const u64*>(boards), after_buf, rewards, mask, num_empty, count_buf,
                     reinterpret_cast<u64*>(children), nullptr, G2048_CHILDREN_COMPACT, n, s->table,
                     s->table_hashed, 1u, child_offsets, capacity, sums, prefix};
        CU(cudaFuncSetAttribute(expand_moves_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kSmemTableBytes + 64u)));
        expand_moves_smem_kernel<<<s->sms, 1024, kSmemTableBytes + 64u, st>>>(a);
        CU(cudaGetLastError());
        scan_top_kernel<<<1, 1024, 0, st>>>(sums, groups, prefix, tiles, capacity, tot);
        CU(cudaGetLastError());
        return launch_expand_children(a, s, st);
    };
    rc = launch();
    const cudaError_t fe = cudaFreeAsync(scratch, st);
    if (rc == G2048_OK && fe != cudaSuccess) return cuda_fail(fe, "cudaFreeAsync(scratch)");
    return rc;
}

// Warp-cooperative: a warp takes 32 boards = 8 KiB of children; lane l serves the FIXED rank
// l & 15 and, in pass j, board 2 j + (l >> 4) of the tile, so every store instruction writes the 32
// slots of two boards (512 contiguous bytes).  All 32 slots of a board are written — the children,
// then zeros — so the caller needs no memset.  (One board per thread, first 2 E slots only, over a
// zero-filled buffer reached 20 % of the HBM peak.)
__global__ void __launch_bounds__(256) after_state_kernel(const u64* __restrict__ boards,
                                                          u64* __restrict__ children,
                                                          uint8_t* __restrict__ num_empty,
                                                          uint16_t* __restrict__ empty_cells, size_t n) {
    const u32 lane = threadIdx.x & 31u, k = lane & 15u;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t base = warp * 32; base < n; base += n_warps * 32) {
        if ((num_empty || empty_cells) && base + lane < n) {
            const u32 e = empty_cells16(split(__ldg(boards + base + lane)));
            if (num_empty) num_empty[base + lane] = (uint8_t)__popc(e);
            if (empty_cells) empty_cells[base + lane] = (uint16_t)e;
        }
        if (!children) continue;
#pragma unroll 4
        for (u32 j = 0; j < 16; ++j) {
            const size_t bi = base + 2u * j + (lane >> 4);
            if (bi >= n) continue;
            const u64 p = __ldg(boards + bi);
            const u32 lo = (u32)p, hi = (u32)(p >> 32);
            const u32 nlo = fast::nz8(lo, 1u), nhi = fast::nz8(hi, 1u);
            const u32 tiles_lo = __popc(nlo), ne = 16u - tiles_lo - __popc(nhi);
            int4 v = make_int4(0, 0, 0, 0);
            if (k < ne) {
                bool in_hi;
                const u32 t = fast::spawn_marker(nlo, nhi, tiles_lo, k, 1u, in_hi) >> 3;  // exponent 1 in the k-th empty cell
                const u32 tl = in_hi ? 0u : t, th = in_hi ? t : 0u;
                v = make_int4((int)(lo + tl), (int)(hi + th), (int)(lo + 2u * tl), (int)(hi + 2u * th));  // tile 2, then tile 4 (gameboard.py:162)
            } else if (ne == 0u && k == 0u) {
                v = make_int4((int)lo, (int)hi, 0, 0);  // gameboard.py:158-159: a full board is its own child
            }
            st_int4_stream(children + bi * 32 + 2u * k, v);
        }
    }
}

// children pointer with 8-byte alignment only: one board per thread, every slot written
__global__ void __launch_bounds__(256) after_state_scalar_kernel(const u64* __restrict__ boards,
                                                                 u64* __restrict__ children, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const u64 p = boards[i];
        u32 e = empty_cells16(split(p));
        u64* dst = children + i * 32;
        u32 k = 0;
        if (e == 0u) dst[k++] = p;  // gameboard.py:158-159
        for (; e; e &= e - 1u) {
            const u32 cell = __ffs(e) - 1u;
            dst[k++] = p | (1ull << (4 * cell));  // tile 2 first (gameboard.py:162)
            dst[k++] = p | (2ull << (4 * cell));  // then tile 4
        }
        for (; k < 32u; ++k) dst[k] = 0ull;
    }
}

extern "C" int g2048_after_state(const uint64_t* boards, uint64_t* children, uint8_t* num_empty,
                                 uint16_t* empty_cells, size_t n, void* stream) {
    if (n == 0) return G2048_OK;
    if (!boards) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    const u64* b = reinterpret_cast<const u64*>(boards);
    u64* ch = reinterpret_cast<u64*>(children);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(n, 256, s->sms, 8);
    if (ch && !aligned(ch, 16)) {
        after_state_scalar_kernel<<<grid, 256, 0, st>>>(b, ch, n);
        ch = nullptr;  // the counts and masks below
        if (!num_empty && !empty_cells) {
            CU(cudaGetLastError());
            return G2048_OK;
        }
    }
    after_state_kernel<<<grid, 256, 0, st>>>(b, ch, num_empty, empty_cells, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// expectation over chance children (row f2): 8 lanes per (board, action), lane k of a group sums
// the cells k, k+8 of its segment, a 3-step shuffle reduction finishes the group.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expect_kernel(const float* __restrict__ v, const uint8_t* __restrict__ child_count,
                                                     const int64_t* __restrict__ child_offsets,
                                                     const int32_t* __restrict__ rewards, float discount,
                                                     float illegal_value, float* __restrict__ q, size_t n) {
    const u32 lane = threadIdx.x & 31u, act = lane >> 3, r0 = lane & 7u;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t i = warp; i < n; i += n_warps) {
        const u32 cc4 = *reinterpret_cast<const u32*>(child_count + i * 4);
        const u32 cc = (cc4 >> (8u * act)) & 0xFFu;
        u32 before = 0;
        for (u32 k = 0; k < act; ++k) before += (cc4 >> (8u * k)) & 0xFFu;
        const float* seg = v + child_offsets[i] + before;
        float acc = 0.0f;
        if (cc == 1u) {
            if (r0 == 0u) acc = seg[0];  // the afterstate itself, probability 1
        } else {
            for (u32 cell = r0; 2u * cell < cc; cell += 8u) acc += 0.9f * seg[2u * cell] + 0.1f * seg[2u * cell + 1u];
        }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
        if (r0 == 0u) {
            const float ne = cc == 1u ? 1.0f : (float)(cc >> 1);
            const float r = rewards ? (float)rewards[i * 4 + act] : 0.0f;
            q[i * 4 + act] = cc ? r + discount * (acc / ne) : illegal_value;
        }
    }
}

extern "C" int g2048_expect(const float* child_values, const uint8_t* child_count, const int64_t* child_offsets,
                            const int32_t* rewards, float discount, float illegal_value, float* q_out, size_t n,
                            void* stream) {
    if (n == 0) return G2048_OK;
    if (!child_values || !child_count || !child_offsets || !q_out || !aligned(child_count, 4)) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    expect_kernel<<<grid_for(n * 32, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        child_values, child_count, child_offsets, rewards, discount, illegal_value, q_out, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// progressive widening (gameboard.py:174 after_state_lazy + :205 generate_outcome): k chance
// children per legal afterstate DRAWN from the spawn distribution instead of all 2 E of them, and the
// matching Monte-Carlo backup.  One thread per (board, action, draw).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_children_kernel(const u64* __restrict__ after,
                                                              const uint8_t* __restrict__ child_count, u32 k,
                                                              u64 seed, u64 env_offset, u32 draw_base,
                                                              u64* __restrict__ children, size_t n) {
    const size_t total = n * 4 * (size_t)k;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t ia = t / k;             // board * 4 + action
        const u32 j = (u32)(t - ia * k);
        u64 out = 0ull;
        if (child_count[ia]) {
            Board b = split(after[ia]);
            u32 zlo, zhi;
            empties(b, zlo, zhi);
            // word of (seed, board, draw index, tag 5): digit 1 = the empty cell, digit 2 = the tile
            spawn_philox(b, zlo, zhi, philox_word(seed, env_offset + (ia >> 2), draw_base + (u32)(ia & 3) * k + j,
                                                  G2048_PHILOX_TAG_CHANCE));
            out = join(b);                   // an afterstate without an empty cell is its own child
        }
        children[t] = out;
    }
}

extern "C" int g2048_sample_children(const uint64_t* after, const uint8_t* child_count, uint32_t k, uint64_t seed,
                                     uint64_t env_offset, uint32_t draw_base, uint64_t* children, size_t n,
                                     void* stream) {
    if (n == 0 || k == 0) return G2048_OK;
    if (!after || !child_count || !children) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    sample_children_kernel<<<grid_for(n * 4 * (size_t)k, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const u64*>(after), child_count, k, seed, env_offset, draw_base,
        reinterpret_cast<u64*>(children), n);
    CU(cudaGetLastError());
    return G2048_OK;
}

__global__ void __launch_bounds__(256) expect_sampled_kernel(const float* __restrict__ v,
                                                             const uint8_t* __restrict__ child_count,
                                                             const int32_t* __restrict__ rewards, u32 k, float discount,
                                                             float illegal_value, float* __restrict__ q, size_t n) {
    for (size_t ia = blockIdx.x * (size_t)blockDim.x + threadIdx.x; ia < n * 4; ia += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.0f;
        for (u32 j = 0; j < k; ++j) acc += v[ia * k + j];
        const float r = rewards ? (float)rewards[ia] : 0.0f;
        q[ia] = child_count[ia] ? r + discount * (acc / (float)k) : illegal_value;
    }
}

extern "C" int g2048_expect_sampled(const float* child_values, const uint8_t* child_count, const int32_t* rewards,
                                    uint32_t k, float discount, float illegal_value, float* q_out, size_t n,
                                    void* stream) {
    if (n == 0) return G2048_OK;
    if (!child_values || !child_count || !q_out || k == 0) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    expect_sampled_kernel<<<grid_for(n * 4, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        child_values, child_count, rewards, k, discount, illegal_value, q_out, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// n-step TD(lambda) returns (self_play.py:524): one thread per (t, env), time-major buffers so
// that a warp reads consecutive envs of one time step; the n loop carries the discounted reward
// sum, the discount power and the lambda power.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nstep_returns_kernel(const float* __restrict__ rewards,
                                                            const float* __restrict__ values,
                                                            const int32_t* __restrict__ lengths, double discount,
                                                            int n_steps, double lam, float* __restrict__ out,
                                                            size_t T, size_t n) {
    const size_t total = T * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t t = idx / n, e = idx - t * n;
        const size_t len = lengths ? (size_t)max(lengths[e], 0) : T;
        if (t >= len) {
            out[idx] = 0.0f;
            continue;
        }
        const size_t n_max = min((size_t)max(n_steps, 0), len - t);
        double ret = 0.0, wsum = 0.0, rsum = 0.0, g = 1.0, lp = 1.0;  // g = discount^(k-1), lp = lambda^(k-1)
        for (size_t k = 1; k <= n_max; ++k) {
            rsum += g * (double)rewards[(t + k - 1) * n + e];
            g *= discount;
            double gk = rsum;
            if (t + k < len) gk += g * (double)values[(t + k) * n + e];
            const double w = (k < (size_t)n_steps) ? (1.0 - lam) * lp : lp;
            ret += w * gk;
            wsum += w;
            lp *= lam;
        }
        out[idx] = (float)(wsum > 0.0 ? ret / wsum : (double)values[idx]);
    }
}

extern "C" int g2048_nstep_returns(const float* rewards, const float* values, const int32_t* lengths,
                                   double discount, int n_steps, double td_lambda, float* out, size_t T,
                                   size_t n, void* stream) {
    if (T == 0 || n == 0) return G2048_OK;
    if (!rewards || !values || !out || n_steps < 0) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    nstep_returns_kernel<<<grid_for(T * n, 256, s->sms, 8), 256, 0, (cudaStream_t)stream>>>(
        rewards, values, lengths, discount, n_steps, td_lambda, out, T, n);
    CU(cudaGetLastError());
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// fused random rollout: boards, scores and step counters live in registers until the game ends
// ---------------------------------------------------------------------------------------------
struct RolloutArgs {
    u64 seed, env_offset;
    size_t n;
    size_t per_queue;  // env ids per CTA queue (roll_fetch)
    u32 max_steps;
    u32 t_end;  // max_steps + 2 (saturated): a slot's `t` at which a game is stopped
    u64* boards_io;
    int64_t* score_out;
    int32_t* steps_out;
    uint8_t* ovf_out;  // one byte per env, zero-filled: set when the episode saw a 15+15 merge
    uint8_t* trace;
    u32 trace_stride;
    const u32* table;
    const u32* table_hashed;
    u32 one;
    PhiloxKeys keys;
};

// pick[mask * 4 + k] = k-th set bit of the 4-bit mask (branch-free policy)
__constant__ uint8_t c_pick[64] = {
    0, 0, 0, 0,  0, 0, 0, 0,  1, 0, 0, 0,  0, 1, 0, 0,  2, 0, 0, 0,  0, 2, 0, 0,  1, 2, 0, 0,  0, 1, 2, 0,
    3, 0, 0, 0,  0, 3, 0, 0,  1, 3, 0, 0,  0, 1, 3, 0,  2, 3, 0, 0,  0, 2, 3, 0,  1, 2, 3, 0,  0, 1, 2, 3};

// Exact out-of-line move for boards holding a tile >= 16384 (the branch-free register row move of
// g2048_device.cuh, which also reports a 15+15 merge exactly).
__device__ __noinline__ uint4 rollout_move_slow(u32 lo, u32 hi, u32 action) {
    const MoveResult m = move_action(Board{lo, hi}, action, RegisterTable{});
    return make_uint4(m.after.lo, m.after.hi, m.score, m.overflow ? 1u : 0u);
}

// Every lane plays kRollSlots games at once, one step of each per loop iteration: the two
// dependent chains interleave (ILP 2 with 5 warps per scheduler instead of ILP 1 with 8).
//
// Random numbers: a game draws from its env's own GAME stream (include/g2048.h "Game stream"):
// word p of the stream is word p & 3 of block p >> 2, so ONE Philox4x32-10 block feeds four
// consecutive steps of a game.  To make that pay in SIMT, the loop runs in QUADS of four
// iterations: games only start at a quad boundary, so every lane of the warp is at word
// (iteration & 3) of its current block — the two blocks of a lane (one per slot) are computed
// once per quad by the whole warp, and inside the unrolled quad the word is a fixed register.
// (Round-2 first version: the quad-layout block of (env >> 2, step) per lane and step, three of
// its four words unused — 43 of 222 instructions per step.)  A game that ends inside a quad is
// frozen (`ended`: no move, no spawn, no step count) until the boundary, where its results are
// written and the next game is fetched: a lane idles 1.5 iterations per ~120-step game, and the
// divergent episode branch runs once per quad instead of once per iteration.  That branch is kept
// to three stores and one atomic: the per-episode results go to per-env arrays (the caller's or
// scratch) and a second, HBM-bound kernel (rollout_stats_kernel) derives the max tile, the capped
// flag and the 24 statistics words from them.  Results are keyed by the global env id and the
// position in the game, so they do not depend on which lane or slot played a game, or when.
//
// One step (roll_step) = the fast transition of g2048_fast.cuh with the action drawn one step
// ahead: the word of the step places the tile (digits 1, 2) and, once the mask of the new board
// is known, picks the NEXT action (digit 3) — so the policy costs one IMAD.HI and one LDS.  The two
// reset tiles run through the same code as phases t = 0, 1 (stream words 0, 1, no move): a new
// game costs two ordinary iterations instead of a divergent reset path.
constexpr int kRollSlots = 2;

struct RollSlot {
    u32 lo, hi;     // board
    u32 act;        // action of the next step (already drawn)
    u32 t;          // 0, 1: reset tiles still to place; t >= 2: playing step t - 2.  Bit 31 = the game
                    // is over (or there is none): the slot is frozen until the next quad boundary —
                    // in the counter's sign bit `playing` and `running` are one signed compare each
    u32 elo, ehi;   // global env id (Philox counter words 0, 1 of the game stream)
    u32 ovf;
    u32 k;          // position of this game in the CTA's queue (env = queue base + k)
    u32 pre;        // prefetched queue position of the NEXT game (kRollNoId: none)
    u32 score;      // < 2^32: even a game that builds a 65536 tile scores ~1.8 M
    u32 bigw;       // OR of the game's move scores: >= 16384 iff the board MAY hold a tile >= 16384 (rows
                    // outside the staged table window) — no smaller move can build such a tile; the
                    // out-of-line path such a game takes re-evaluates it exactly
    bool alive;     // the slot holds a game whose results have not been written yet
};
constexpr u32 kRollEnded = 0x80000000u;

// some tile >= 16384 (nibble >= 14).  Plain C: the lop3 asm of fast::has_big_tile crashes ptxas 12.9
// (SIGSEGV) once two slots are unrolled in the rollout kernel.
__device__ __forceinline__ bool roll_has_big(u32 lo, u32 hi) {
    return ((((lo << 1) & (lo << 2) & lo) | ((hi << 1) & (hi << 2) & hi)) & 0x88888888u) != 0u;
}

// The per-action orientation words and the action pick table live in DYNAMIC shared memory behind
// the row table (or at its start when the table is gathered from L2) and are indexed off the extern
// symbol: through a generic pointer to a static __shared__ array ptxas rebuilt the shared-window
// address (S2R SR_CgaCtaId + LEA) in front of every lookup — eight instructions per step.
template <bool kSmem>
struct RollShared {
    static constexpr u32 kOrientByte = ((kSmem ? kSmemRows : 0u) + 16u) * 4u;  // after the mbarrier slot
    static constexpr u32 kPickByte = kOrientByte + 4u * (u32)sizeof(fast::Orient);
    static constexpr u32 kBytes = kPickByte + 64u;
    __device__ __forceinline__ fast::Orient orient(u32 action) const {
        const unsigned char* op = reinterpret_cast<const unsigned char*>(smem_words) + kOrientByte + 32u * action;
        const uint4 o0 = *reinterpret_cast<const uint4*>(op);
        fast::Orient o;
        o.mask = o0.x; o.mul_hi = o0.y; o.mul_lo = o0.z; o.sel_lo = o0.w;
        o.sel_hi = *reinterpret_cast<const u32*>(op + 16);
        return o;
    }
    __device__ __forceinline__ u32 pick(u32 i) const {
        return *(reinterpret_cast<const unsigned char*>(smem_words) + kPickByte + i);
    }
};

// per-thread constants of the CTA's queue: its first env (local index), its length, and the output
// pointers moved to that base (the host side always provides boards / score / steps buffers)
struct RollQueue {
    size_t lo;
    u32 own;
    u64* boards;
    int64_t* score;
    int32_t* steps;
    uint8_t* ovf;
    uint8_t* trace;
    unsigned int* counter;
};

__device__ __forceinline__ RollQueue roll_queue(const RolloutArgs& a, unsigned int* next) {
    RollQueue q;
    q.lo = (size_t)blockIdx.x * a.per_queue;
    q.own = q.lo >= a.n ? 0u : (u32)min((size_t)a.per_queue, a.n - q.lo);
    q.boards = a.boards_io + q.lo;
    q.score = a.score_out + q.lo;
    q.steps = a.steps_out + q.lo;
    q.ovf = a.ovf_out ? a.ovf_out + q.lo : nullptr;
    q.trace = a.trace ? a.trace + q.lo * a.trace_stride : nullptr;
    q.counter = next + blockIdx.x;
    return q;
}

// the episode of slot `s` is over: its per-env results (rollout_stats_kernel does the rest)
template <bool kTrace>
__device__ __forceinline__ void roll_emit(const RollSlot& s, const RolloutArgs& a, const RollQueue& q) {
    const u32 step = (s.t & ~kRollEnded) - 2u;
    if (kTrace) q.trace[(size_t)s.k * a.trace_stride + min(step, a.trace_stride - 1u)] = 0;
    q.boards[s.k] = ((u64)s.hi << 32) | s.lo;
    q.score[s.k] = (int64_t)s.score;
    q.steps[s.k] = (int32_t)step;
    if (s.ovf && q.ovf) q.ovf[s.k] = 1;  // zero-filled by the host side
}

// Next game for slot `s`.  The env range is cut into one queue per CTA (`per_queue` consecutive
// ids and a counter each), so no counter is shared between SMs: ONE counter for the whole grid
// serialises in L2 (~0.7 G same-address atomics/s capped the round-1 kernel).  The queues hold
// equal numbers of games; a CTA plays ~14 000 of them, so the work per CTA differs by ~0.4 %
// (sqrt(14 000) x 55 steps against 14 000 x 118) and stealing between queues is not worth its
// probes (an earlier version with stealing spent 11 % of its stall samples in the final, empty
// probe round: ncu, profiles/r02_rollout_ncu_summary.txt).
// The id is PREFETCHED: the atomic for the game after this one is issued now and its result
// (s.pre) is not read before the next episode end, ~118 steps later, so the ~1 us round trip never
// stalls the warp.  (Inline PTX: the compiler would aggregate atomicAdd over the warp and shuffle
// the result at once, which waits for it.)  Prefetching stops kRollMargin ids before the end of
// the queue: the last ids are fetched on demand, so no slot reserves a game others could start.
// kFromReset: the game starts on the empty board in phase 0.  Otherwise it starts from boards_io
// with the first action drawn from the word of tag 3; a board that is already over is emitted here.
constexpr u32 kRollNoId = ~0u;
constexpr u32 kRollMargin = 4096;  // > 2 x slots of a CTA

// `lane_zero` is a per-lane value that is always 0 (the slot's overflow flag >> 1): it makes the
// address look lane-dependent, so ptxas does not aggregate the atomic over the warp — the
// aggregated form shuffles the result to the lanes at once, i.e. waits for the round trip.
__device__ __forceinline__ u32 queue_take(unsigned int* counter, u32 lane_zero) {
    u32 k;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(k) : "l"(counter + lane_zero) : "memory");
    return k;
}

template <bool kFromReset, bool kTrace, class Sh>
__device__ __forceinline__ void roll_fetch(RollSlot& s, const RolloutArgs& a, const Sh& sm,
                                           const RollQueue& q) {
    for (;;) {
        u32 k = s.pre;  // reserved at the previous game start (or none)
        const u32 lane_zero = s.ovf >> 1;  // ovf is 0 or 1
        if (k == kRollNoId) k = queue_take(q.counter, lane_zero);  // on demand
        const bool got = k < q.own;
        s.pre = kRollNoId;
        if (got && k + kRollMargin < q.own) s.pre = queue_take(q.counter, lane_zero);  // not read until the next game end
        s.lo = s.hi = s.act = s.ovf = s.bigw = 0u;
        s.t = got ? 0u : kRollEnded;
        s.score = 0;
        s.k = k;
        s.alive = got;
        if (!got) {
            s.elo = s.ehi = 0u;
            return;
        }
        const u64 env = a.env_offset + q.lo + k;
        s.elo = (u32)env;
        s.ehi = (u32)(env >> 32);
        if (kFromReset) return;
        const Board b = split(q.boards[k]);
        s.lo = b.lo;
        s.hi = b.hi;
        s.t = 2u;
        s.bigw = roll_has_big(b.lo, b.hi) ? 16384u : 0u;
        const u32 mask = fast::legal_mask(b.lo, b.hi, fast::nz8(b.lo, a.one), fast::nz8(b.hi, a.one), a.one);
        const u32 w = philox_word(a.keys, env, 0u, G2048_PHILOX_TAG_START);
        s.act = sm.pick(mask * 4u + __umulhi(w, (u32)__popc(mask)));
        if (mask != 0u && a.max_steps != 0u) return;
        s.t |= kRollEnded;
        roll_emit<kTrace>(s, a, q);  // nothing to play
    }
}

// the block that holds the next four words of the slot's game: stream position = t from reset
// (words 0, 1 are the reset tiles), t - 2 from given boards; a multiple of 4 at every quad boundary
template <bool kFromReset>
__device__ __forceinline__ Philox4 roll_block(const RollSlot& s, const RolloutArgs& a) {
    return philox_game_block(a.keys, s.elo, s.ehi, (kFromReset ? s.t : s.t - 2u) >> 2);
}

// one step (or reset phase) of slot `s` with the stream word `w`
template <bool kTrace, class Sh, class Lut>
__device__ __forceinline__ void roll_step(RollSlot& s, u32 w, const RolloutArgs& a, const Sh& sm,
                                          const RollQueue& q, const Lut& lut) {
    const u32 one = a.one;
    const bool run = (int)s.t >= 0;      // not frozen
    const bool playing = (int)s.t >= 2;  // and past the reset phases
    const bool maybe_big = s.bigw >= 16384u;
    // The fast move runs unconditionally — on the EMPTY board in a reset phase, in a frozen slot and
    // when the board may hold a tile >= 16384 (such a row lies outside the staged table window): the
    // move of the empty board is the empty board with score 0, so OR-ing the withheld board back in
    // restores it (one LOP3 per half, no selects on the results).
    const bool fast_ok = playing & !maybe_big;
    const u32 xl = fast_ok ? s.lo : 0u, xh = fast_ok ? s.hi : 0u;
    u32 al, ah, reward;
    {
        bool moved;
        reward = fast::move_fast(xl, xh, sm.orient(s.act), lut, one, al, ah, moved);
    }
    al |= s.lo ^ xl;
    ah |= s.hi ^ xh;
    if (maybe_big & playing) {  // never in random play: the exact out-of-line move
        const uint4 r = rollout_move_slow(s.lo, s.hi, s.act);
        al = r.x; ah = r.y; reward = r.z;
        s.ovf |= r.w;
        s.bigw = roll_has_big(al, ah) ? 16384u : 0u;
    }
    s.bigw |= reward;
    s.score += reward;
    u32 nlo = fast::nz8(al, one), nhi = fast::nz8(ah, one);
    const u32 tiles_lo = fast::popc(nlo), ne = 16u - (tiles_lo + fast::popc(nhi));
    u32 rank, f1, tenth, f2;
    fast::mul_wide(w, ne, rank, f1);      // digit 1: spawn rank
    fast::mul_wide(f1, 10u, tenth, f2);   // digit 2: 0 -> tile 4
    bool in_hi;
    u32 bit3 = fast::spawn_marker(nlo, nhi, tiles_lo, rank, one, in_hi);
    if (!run) bit3 = 0u;
    const u32 exp = 2u - fast::nonzero01(tenth);
    const u32 tile = fast::shr<3>(bit3) * exp;
    if (in_hi) { ah += tile; nhi += bit3; } else { al += tile; nlo += bit3; }
    if (kTrace) {
        const u32 step = s.t - 2u;
        if (playing && step + 1u < a.trace_stride) {
            const u32 cell = ((__ffs(bit3) - 1u) >> 2) + (in_hi ? 8u : 0u);
            q.trace[(size_t)s.k * a.trace_stride + step] =
                (uint8_t)(0x80u | s.act | (cell << 2) | ((exp - 1u) << 6));
        }
    }
    const u32 mask = fast::legal_mask(al, ah, nlo, nhi, one);
    // digit 3: the next action, the k-th legal one of the new board
    const u32 act = sm.pick(mask * 4u + fast::mulhi(f2, fast::popc(mask)));
    s.lo = al;
    s.hi = ah;
    s.act = act;
    // branch-free: the board of a reset phase always has a legal move, t_end is in [2, 2^31), and a
    // frozen counter is negative
    if (run) s.t += 1u;
    if ((mask == 0u) | ((int)s.t >= (int)a.t_end)) s.t |= kRollEnded;
}

// quad boundary: results of a finished game out, next game in
template <bool kFromReset, bool kTrace, class Sh>
__device__ __forceinline__ void roll_finish(RollSlot& s, const RolloutArgs& a, const Sh& sm,
                                            const RollQueue& q) {
    if (s.alive && (int)s.t < 0) {
        roll_emit<kTrace>(s, a, q);
        roll_fetch<kFromReset, kTrace>(s, a, sm, q);
    }
}

#define G2048_ROLL_WORD(p, u) ((u) == 0 ? (p).x : (u) == 1 ? (p).y : (u) == 2 ? (p).z : (p).w)

// The drain.  Once the queue is empty a lane is left with its last two games, and a warp stays in
// the loop until the longest of its 64 last games ends (~300 iterations) while its lanes fall idle.
// Two measures halve what that costs: a lane whose slot 0 runs dry takes over its own slot-1 game
// (so live games sit in slot 0), and the warp votes once per quad whether ANY lane still has a
// slot-1 game — if not, it runs the one-slot quad (half the instructions).  All lanes stay in the
// loop until the warp is done (a lane without games holds frozen slots), so the votes are
// full-warp.  The two-slot quad keeps both steps of an iteration in one basic block: ptxas
// interleaves the two dependent chains (ILP 2).
template <bool kFromReset, bool kTrace, class Sh, class Lut>
__device__ __forceinline__ void rollout_loop(const RolloutArgs& a, const Sh& sm, const Lut& lut,
                                             unsigned int* next) {
    const u32 FULL = 0xFFFFFFFFu;
    const RollQueue q = roll_queue(a, next);
    RollSlot slot[kRollSlots];
#pragma unroll
    for (int k = 0; k < kRollSlots; ++k) {
        slot[k].pre = kRollNoId;
        slot[k].ovf = 0u;
        roll_fetch<kFromReset, kTrace>(slot[k], a, sm, q);
    }
    for (;;) {
        if (__any_sync(FULL, slot[1].alive)) {
            const Philox4 p0 = roll_block<kFromReset>(slot[0], a), p1 = roll_block<kFromReset>(slot[1], a);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                roll_step<kTrace>(slot[0], G2048_ROLL_WORD(p0, u), a, sm, q, lut);
                roll_step<kTrace>(slot[1], G2048_ROLL_WORD(p1, u), a, sm, q, lut);
            }
            roll_finish<kFromReset, kTrace>(slot[0], a, sm, q);
            roll_finish<kFromReset, kTrace>(slot[1], a, sm, q);
            if (!slot[0].alive && slot[1].alive) {  // the queue is empty: keep the lane's last game in slot 0
                slot[0] = slot[1];
                slot[1].alive = false;
                slot[1].t |= kRollEnded;
            }
        } else {
            if (!__any_sync(FULL, slot[0].alive)) break;
            const Philox4 p0 = roll_block<kFromReset>(slot[0], a);
#pragma unroll
            for (int u = 0; u < 4; ++u) roll_step<kTrace>(slot[0], G2048_ROLL_WORD(p0, u), a, sm, q, lut);
            roll_finish<kFromReset, kTrace>(slot[0], a, sm, q);
        }
    }
}
#undef G2048_ROLL_WORD

template <bool kSmem, bool kFromReset, bool kTrace, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) rollout_kernel(const RolloutArgs a, unsigned int* next) {
    using Sh = RollShared<kSmem>;
    u64* bar = reinterpret_cast<u64*>(smem_words + kSmemRows);
    if (kSmem) table_stage_begin(smem_words, bar, a.table_hashed);
    unsigned char* sh = reinterpret_cast<unsigned char*>(smem_words);
    if (threadIdx.x < 32)
        reinterpret_cast<u32*>(sh + Sh::kOrientByte)[threadIdx.x] = reinterpret_cast<const u32*>(c_orient)[threadIdx.x];
    if (threadIdx.x < 64) sh[Sh::kPickByte + threadIdx.x] = c_pick[threadIdx.x];
    if (kSmem) table_stage_wait(bar); else __syncthreads();

    const Sh sm{};
    if (kSmem) rollout_loop<kFromReset, kTrace>(a, sm, SmemLut{}, next);
    else rollout_loop<kFromReset, kTrace>(a, sm, GlobalLut{a.table_hashed}, next);
}

// Second kernel of a rollout: max tile, capped flag and the 24 statistics words from the per-env
// results (HBM-bound: 21 bytes per episode).  An episode that ended with a legal action left was
// stopped by max_steps.
__global__ void __launch_bounds__(256) rollout_stats_kernel(const u64* __restrict__ boards,
                                                            const int64_t* __restrict__ score,
                                                            const int32_t* __restrict__ steps,
                                                            const uint8_t* __restrict__ ovf,
                                                            uint8_t* __restrict__ maxexp_out,
                                                            u64* __restrict__ stats, size_t n) {
    __shared__ u64 sh[G2048_STATS_WORDS];
    __shared__ u32 sh_hist[16];  // 32-bit shared atomics are native; the 64-bit form is a CAS loop
    if (threadIdx.x < G2048_STATS_WORDS) sh[threadIdx.x] = 0;
    if (threadIdx.x < 16) sh_hist[threadIdx.x] = 0;
    __syncthreads();
    u64 t_steps = 0, t_score = 0, t_maxscore = 0;
    u32 t_eps = 0, t_maxsteps = 0, t_capped = 0, t_ovf = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const Board b = split(boards[i]);
        const u32 mx = max_exponent(b);
        if (maxexp_out) maxexp_out[i] = (uint8_t)mx;
        if (!stats) continue;
        const u64 sc = (u64)score[i];
        const u32 st = (u32)steps[i];
        atomicAdd(&sh_hist[mx], 1u);  // < 2^32 episodes per CTA: the grid-stride loop gives a CTA n / grid of them
        t_eps += 1u;
        t_steps += st;
        t_score += sc;
        t_maxscore = max(t_maxscore, sc);
        t_maxsteps = max(t_maxsteps, st);
        t_capped += legal_mask(b) != 0u ? 1u : 0u;
        t_ovf += ovf[i] ? 1u : 0u;
    }
    if (!stats) return;
    const u32 FULL = 0xFFFFFFFFu;
    for (int off = 16; off; off >>= 1) {
        t_eps += __shfl_xor_sync(FULL, t_eps, off);
        t_steps += __shfl_xor_sync(FULL, t_steps, off);
        t_score += __shfl_xor_sync(FULL, t_score, off);
        t_capped += __shfl_xor_sync(FULL, t_capped, off);
        t_ovf += __shfl_xor_sync(FULL, t_ovf, off);
        t_maxscore = max(t_maxscore, (u64)__shfl_xor_sync(FULL, t_maxscore, off));
        t_maxsteps = max(t_maxsteps, __shfl_xor_sync(FULL, t_maxsteps, off));
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicAdd(&sh[0], (u64)t_eps);
        atomicAdd(&sh[1], t_steps);
        atomicAdd(&sh[2], t_score);
        atomicMax(&sh[3], t_maxscore);
        atomicMax(&sh[4], (u64)t_maxsteps);
        atomicAdd(&sh[5], (u64)t_capped);
        atomicAdd(&sh[6], (u64)t_ovf);
    }
    __syncthreads();
    if (threadIdx.x < G2048_STATS_WORDS) {
        const u64 v = threadIdx.x >= 8 ? (u64)sh_hist[threadIdx.x - 8] : sh[threadIdx.x];
        if (threadIdx.x == 3 || threadIdx.x == 4) atomicMax(&stats[threadIdx.x], v);
        else if (v) atomicAdd(&stats[threadIdx.x], v);
    }
}

// persistent, one CTA per SM, table in shared memory: 640 threads x 2 slots at 94 registers (no
// spills; 768 threads are capped at 80 registers and keep more games in flight at the end: 155.0 /
// 153.7 / 149.5 G env-steps/s at 640 / 768 / 512 threads on 2^24 envs, 127.5 / 123.6 / 126.8 on 2^20).
// (environment G2048_ROLL_THREADS = 512 | 640 | 768: a tuning knob for measurements)
static int g_roll_threads = [] { const char* e = getenv("G2048_ROLL_THREADS"); const int t = e ? atoi(e) : 0; return (t == 512 || t == 640 || t == 768) ? t : 0; }();
constexpr int kRollSmallThreads = 128;  // small batches: table gathered from L2

extern "C" int g2048_rollout(uint64_t seed, uint64_t env_offset, size_t n, uint32_t max_steps,
                             int from_reset, uint64_t* boards_io, int64_t* score_out,
                             int32_t* steps_out, uint8_t* maxexp_out, uint64_t* stats_out,
                             uint8_t* trace, uint32_t trace_stride, void* stream) {
    if (n == 0) return G2048_OK;
    if (!from_reset && !boards_io) return G2048_EINVAL;
    if (trace && trace_stride < 2) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool smem = n >= (size_t)s->sms * 512 && s->smem_optin >= RollShared<true>::kBytes;
    const int bytes = smem ? (int)RollShared<true>::kBytes : (int)RollShared<false>::kBytes;
    const int threads = smem ? (g_roll_threads ? g_roll_threads : 640) : kRollSmallThreads;
    const int grid = smem ? s->sms : grid_for((n + kRollSlots - 1) / kRollSlots, kRollSmallThreads, s->sms, 8);
    // scratch (stream-ordered): the queue counters, and the per-env result arrays the caller does not
    // ask for (the kernel always writes boards / score / steps; the statistics pass reads them)
    const bool post = stats_out || maxexp_out;
    const size_t per_queue = (n + (size_t)grid - 1) / (size_t)grid;
    if (per_queue >= ((size_t)1 << 31)) return G2048_EINVAL;
    const size_t counters = ((size_t)grid * 4 + 15) / 16 * 16;
    const size_t sz_boards = boards_io ? 0 : n * 8, sz_score = score_out ? 0 : n * 8;
    const size_t sz_steps = steps_out ? 0 : (n * 4 + 15) / 16 * 16, sz_ovf = post ? n : 0;
    unsigned char* scratch = nullptr;
    CU(cudaMallocFromPoolAsync(&scratch, counters + sz_boards + sz_score + sz_steps + sz_ovf, s->pool, st));
    auto launch = [&]() -> int {  // in a helper so that the scratch is released on every path
        unsigned int* next = reinterpret_cast<unsigned int*>(scratch);
        unsigned char* p = scratch + counters;
        u64* boards_buf = boards_io ? reinterpret_cast<u64*>(boards_io) : reinterpret_cast<u64*>(p);
        p += sz_boards;
        int64_t* score_buf = score_out ? score_out : reinterpret_cast<int64_t*>(p);
        p += sz_score;
        int32_t* steps_buf = steps_out ? steps_out : reinterpret_cast<int32_t*>(p);
        p += sz_steps;
        uint8_t* ovf_buf = sz_ovf ? p : nullptr;
        CU(cudaMemsetAsync(next, 0, (size_t)grid * sizeof(unsigned int), st));
        if (ovf_buf) CU(cudaMemsetAsync(ovf_buf, 0, n, st));
        const RolloutArgs a{seed, env_offset, n, per_queue, max_steps,
                            max_steps > 0x7FFFFFFDu ? 0x7FFFFFFFu : max_steps + 2u, boards_buf,
                            score_buf, steps_buf, ovf_buf, trace, trace_stride, s->table, s->table_hashed, 1u,
                            philox_key_schedule(seed)};
#define G2048_LAUNCH_ROLL(SM, FR, TR, TH)                                                                  \
    do {                                                                                                   \
        if (SM) CU(cudaFuncSetAttribute(rollout_kernel<SM, FR, TR, TH>,                                    \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));              \
        rollout_kernel<SM, FR, TR, TH><<<grid, threads, bytes, st>>>(a, next);                             \
    } while (0)
#define G2048_LAUNCH_ROLL_T(SM, TH)                                   \
    do {                                                              \
        if (from_reset) {                                             \
            if (trace) G2048_LAUNCH_ROLL(SM, true, true, TH);         \
            else G2048_LAUNCH_ROLL(SM, true, false, TH);              \
        } else {                                                      \
            if (trace) G2048_LAUNCH_ROLL(SM, false, true, TH);        \
            else G2048_LAUNCH_ROLL(SM, false, false, TH);             \
        }                                                             \
    } while (0)
        if (smem && threads == 768) G2048_LAUNCH_ROLL_T(true, 768);
        else if (smem && threads == 640) G2048_LAUNCH_ROLL_T(true, 640);
        else if (smem) G2048_LAUNCH_ROLL_T(true, 512);
        else G2048_LAUNCH_ROLL_T(false, kRollSmallThreads);
#undef G2048_LAUNCH_ROLL_T
#undef G2048_LAUNCH_ROLL
        CU(cudaGetLastError());
        if (post) {
            rollout_stats_kernel<<<grid_for(n, 256, s->sms, 8), 256, 0, st>>>(
                boards_buf, score_buf, steps_buf, ovf_buf, maxexp_out, reinterpret_cast<u64*>(stats_out), n);
            CU(cudaGetLastError());
        }
        return G2048_OK;
    };
    rc = launch();
    const cudaError_t fe = cudaFreeAsync(scratch, st);
    if (rc == G2048_OK && fe != cudaSuccess) return cuda_fail(fe, "cudaFreeAsync(rollout scratch)");
    return rc;
}

// ---------------------------------------------------------------------------------------------
// row table dump (tests: SURVEY.md §8c hashes)
// ---------------------------------------------------------------------------------------------
__global__ void row_table_kernel(uint16_t* res, u32* reward, uint8_t* moved) {
    u32 row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= kRows) return;
    u32 score;
    bool ovf;
    u32 r = row_left_exact(row, &score, &ovf);
    res[row] = (uint16_t)r;
    reward[row] = score;
    moved[row] = (uint8_t)(r != row);
}

extern "C" int g2048_row_table_host(uint16_t* result, uint32_t* reward, uint8_t* moved) {
    if (!result || !reward || !moved) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    uint16_t* d_res;
    u32* d_rew;
    uint8_t* d_mv;
    CU(cudaMalloc(&d_res, kRows * 2));
    CU(cudaMalloc(&d_rew, kRows * 4));
    CU(cudaMalloc(&d_mv, kRows));
    row_table_kernel<<<kRows / 256, 256>>>(d_res, d_rew, d_mv);
    CU(cudaGetLastError());
    // cross-check the packed entries actually used by the kernels
    std::vector<u32> h_tab(kRows);
    CU(cudaMemcpy(result, d_res, kRows * 2, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(reward, d_rew, kRows * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(moved, d_mv, kRows, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_tab.data(), s->table, kRows * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_res);
    cudaFree(d_rew);
    cudaFree(d_mv);
    for (u32 r = 0; r < kRows; ++r)
        if ((h_tab[r] >> 16) != result[r] || (h_tab[r] & 0xFFFFu) != (reward[r] >> 2)) {
            snprintf(g_cuda_err, sizeof g_cuda_err, "row table entry %u inconsistent", r);
            return G2048_ECUDA;
        }
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer transition: chunked H2D -> kernel -> D2H pipeline over 4 streams
// ---------------------------------------------------------------------------------------------
namespace {
struct HostPipe {
    static constexpr int kSlots = 4;
    static constexpr size_t kChunk = (size_t)1 << 21;  // capacity of a slot (envs); the stage size is chosen at run time
    int dev = -1;
    cudaStream_t stream[kSlots] = {};
    uint64_t *d_in[kSlots] = {}, *d_out[kSlots] = {};
    uint8_t *d_act[kSlots] = {}, *d_spn[kSlots] = {}, *d_flg[kSlots] = {};
    int32_t* d_rew[kSlots] = {};
};

// Pool of pipes: a call takes one for its device (creating it on first use), returns it when done.
// Concurrent callers each get their own; g2048_shutdown frees the pooled ones.
std::mutex g_pipe_mu;
std::vector<HostPipe*> g_pipes[64];

void pipe_destroy(HostPipe* p) {
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        if (p->stream[i]) cudaStreamDestroy(p->stream[i]);
        cudaFree(p->d_in[i]); cudaFree(p->d_out[i]); cudaFree(p->d_act[i]);
        cudaFree(p->d_spn[i]); cudaFree(p->d_flg[i]); cudaFree(p->d_rew[i]);
    }
    delete p;
}

int pipe_create(HostPipe* p) {
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        CU(cudaStreamCreateWithFlags(&p->stream[i], cudaStreamNonBlocking));
        CU(cudaMalloc(&p->d_in[i], HostPipe::kChunk * 8));
        CU(cudaMalloc(&p->d_out[i], HostPipe::kChunk * 8));
        CU(cudaMalloc(&p->d_act[i], HostPipe::kChunk));
        CU(cudaMalloc(&p->d_spn[i], HostPipe::kChunk));
        CU(cudaMalloc(&p->d_flg[i], HostPipe::kChunk));
        CU(cudaMalloc(&p->d_rew[i], HostPipe::kChunk * 4));
    }
    return G2048_OK;
}

int pipe_acquire(HostPipe** out) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return G2048_EINVAL;
    {
        std::lock_guard<std::mutex> lk(g_pipe_mu);
        if (!g_pipes[dev].empty()) {
            *out = g_pipes[dev].back();
            g_pipes[dev].pop_back();
            return G2048_OK;
        }
    }
    HostPipe* p = new HostPipe();
    p->dev = dev;
    const int rc = pipe_create(p);
    if (rc) {
        pipe_destroy(p);
        return rc;
    }
    *out = p;
    return G2048_OK;
}

void pipe_release(HostPipe* p) {
    std::lock_guard<std::mutex> lk(g_pipe_mu);
    g_pipes[p->dev].push_back(p);
}

// envs per pipeline stage (<= kChunk); environment G2048_HOST_CHUNK_LOG2 overrides
size_t host_chunk() {
    const char* e = getenv("G2048_HOST_CHUNK_LOG2");
    const int lg = e ? atoi(e) : 21;
    return (lg >= 12 && lg <= 21) ? ((size_t)1 << lg) : HostPipe::kChunk;
}

// kCompute = false: the same copies on the same streams without the kernel (the PCIe ceiling of
// the call, g2048_step_host_copies)
int step_host_impl(bool compute, const uint64_t* boards_in, const uint8_t* actions, const uint8_t* spawns,
                   uint64_t seed, uint64_t env_offset, uint32_t step_base, uint64_t* boards_out,
                   int32_t* rewards, uint8_t* flags, size_t n) {
    if (n == 0) return G2048_OK;
    if (!boards_in || !actions || !boards_out || !rewards || !flags) return G2048_EINVAL;
    DeviceState* s;
    int rc = current_state(&s);
    if (rc) return rc;
    HostPipe* p;
    rc = pipe_acquire(&p);
    if (rc) return rc;
    const size_t chunk = host_chunk();
    auto run = [&]() -> int {
        size_t done = 0;
        int slot = 0;
        while (done < n) {
            const size_t m = (n - done < chunk) ? n - done : chunk;
            cudaStream_t st = p->stream[slot];
            // stream order makes reuse of this slot's buffers safe (previous chunk's D2H precedes)
            CU(cudaMemcpyAsync(p->d_in[slot], boards_in + done, m * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(p->d_act[slot], actions + done, m, cudaMemcpyHostToDevice, st));
            if (spawns) CU(cudaMemcpyAsync(p->d_spn[slot], spawns + done, m, cudaMemcpyHostToDevice, st));
            if (compute) {
                const int r = g2048_step(p->d_in[slot], p->d_act[slot], spawns ? p->d_spn[slot] : nullptr, seed,
                                         env_offset + done, nullptr, step_base, nullptr, nullptr, p->d_out[slot],
                                         p->d_rew[slot], p->d_flg[slot], m, st);
                if (r) return r;
            }
            CU(cudaMemcpyAsync(boards_out + done, p->d_out[slot], m * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(rewards + done, p->d_rew[slot], m * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(flags + done, p->d_flg[slot], m, cudaMemcpyDeviceToHost, st));
            done += m;
            slot = (slot + 1) % HostPipe::kSlots;
        }
        return G2048_OK;
    };
    rc = run();
    // drain every stream before the pipe goes back to the pool, whatever happened above
    for (int i = 0; i < HostPipe::kSlots; ++i) {
        const cudaError_t e = cudaStreamSynchronize(p->stream[i]);
        if (e != cudaSuccess && rc == G2048_OK) rc = cuda_fail(e, "cudaStreamSynchronize(host pipe)");
    }
    pipe_release(p);
    return rc;
}
}  // namespace

extern "C" int g2048_step_host(const uint64_t* boards_in, const uint8_t* actions,
                               const uint8_t* spawns, uint64_t seed, uint64_t env_offset,
                               uint32_t step_base, uint64_t* boards_out, int32_t* rewards,
                               uint8_t* flags, size_t n) {
    return step_host_impl(true, boards_in, actions, spawns, seed, env_offset, step_base, boards_out, rewards,
                          flags, n);
}

extern "C" int g2048_step_host_copies(const uint64_t* boards_in, const uint8_t* actions,
                                      const uint8_t* spawns, uint64_t* boards_out, int32_t* rewards,
                                      uint8_t* flags, size_t n) {
    return step_host_impl(false, boards_in, actions, spawns, 0, 0, 0, boards_out, rewards, flags, n);
}

// ---------------------------------------------------------------------------------------------
// pinned host memory, placed on the NUMA node of the current GPU when the kernel allows it
// ---------------------------------------------------------------------------------------------
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <map>

namespace {
struct HostBlock {
    size_t bytes;
    bool mapped;  // mmap + cudaHostRegister (else cudaHostAlloc)
};
std::mutex g_host_mu;
std::map<void*, HostBlock> g_host_blocks;

// NUMA node of the PCI device behind CUDA device `dev` (sysfs), -1 if unknown
int gpu_numa_node(int dev) {
    char bus[64] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[160];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// node that holds the page at `p` (get_mempolicy with MPOL_F_NODE | MPOL_F_ADDR), -1 if unknown
int page_numa_node(const void* p) {
    int node = -1;
    if (syscall(SYS_get_mempolicy, &node, nullptr, 0ul, p, 3ul /* MPOL_F_NODE | MPOL_F_ADDR */) != 0) return -1;
    return node;
}
}  // namespace

extern "C" int g2048_host_alloc(void** ptr, size_t bytes) {
    if (!ptr || bytes == 0) return G2048_EINVAL;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const int node = gpu_numa_node(dev);
    const char* off = getenv("G2048_HOST_NUMA");
    if (node >= 0 && node < 64 && !(off && atoi(off) == 0)) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t len = (bytes + page - 1) / page * page;
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            const unsigned long mask = 1ul << node;
            // MPOL_PREFERRED (1): place on the GPU's node when the cpuset allows it, never fail hard
            (void)syscall(SYS_mbind, p, len, 1, &mask, 65ul, 0u);
            memset(p, 0, len);  // first touch under the policy
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_host_mu);
                g_host_blocks[p] = HostBlock{len, true};
                *ptr = p;
                return G2048_OK;
            }
            (void)cudaGetLastError();
            munmap(p, len);
        }
    }
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_host_blocks[*ptr] = HostBlock{bytes, false};
    return G2048_OK;
}

extern "C" int g2048_host_free(void* ptr) {
    if (!ptr) return G2048_OK;
    HostBlock b{0, false};
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_host_blocks.find(ptr);
        if (it == g_host_blocks.end()) return G2048_EINVAL;
        b = it->second;
        g_host_blocks.erase(it);
    }
    if (b.mapped) {
        CU(cudaHostUnregister(ptr));
        munmap(ptr, b.bytes);
    } else {
        CU(cudaFreeHost(ptr));
    }
    return G2048_OK;
}

extern "C" int g2048_host_numa_info(const void* ptr, int* gpu_node, int* mem_node) {
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if (gpu_node) *gpu_node = gpu_numa_node(dev);
    if (mem_node) *mem_node = ptr ? page_numa_node(ptr) : -1;
    return G2048_OK;
}

// ---------------------------------------------------------------------------------------------
// shutdown: release every device resource the library holds
// ---------------------------------------------------------------------------------------------
extern "C" int g2048_shutdown(void) {
    int keep = 0;
    const bool have_dev = cudaGetDevice(&keep) == cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_pipe_mu);
        for (int d = 0; d < 64; ++d) {
            if (g_pipes[d].empty()) continue;
            cudaSetDevice(d);
            for (HostPipe* p : g_pipes[d]) pipe_destroy(p);
            g_pipes[d].clear();
        }
    }
    for (int d = 0; d < 64; ++d) {
        DeviceState& s = g_dev[d];
        std::lock_guard<std::mutex> lk(s.mu);
        if (!s.ready.load(std::memory_order_acquire)) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        cudaFree(s.table);
        cudaFree(s.table_hashed);
        s.table = s.table_hashed = nullptr;
        if (s.pool) cudaMemPoolDestroy(s.pool);
        s.pool = nullptr;
        s.ready.store(false, std::memory_order_release);
    }
    if (have_dev) cudaSetDevice(keep);
    (void)cudaGetLastError();
    return G2048_OK;
}
