// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A small SIMT emulator that lets the .cu kernel sources of qoipp_b200/csrc be compiled with g++ and
// stepped on the CPU: every CUDA thread is a fiber, warp collectives (__shfl_sync, __ballot_sync,
// __match_any_sync, ...) and __syncthreads are rendezvous points, a bounded set of CTAs is "resident"
// at a time (launched in blockIdx order, like the hardware) and the fiber scheduler is seeded so that
// different interleavings of the decoupled look-back can be exercised.  It exists because the build
// container has no GPU: kernel LOGIC (indexing, carries, edge cases) is checked here against the
// oracle before GPU minutes are spent; memory-model and performance questions are answered only on
// the B200.  It is never linked into the product library and is not a fallback.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>

#include <sys/mman.h>

#define QB_EMU 1

// ---------------------------------------------------------------- keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static const

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return { a, b, c, d }; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return { a, b }; }
typedef void* cudaStream_t;

namespace emu
{
    constexpr int kStack = 96 * 1024;

    enum class Wait : uint8_t { None, Warp, Block, Spin, Grid };

    struct Cta;
    struct Fiber {
        void*    sp = nullptr;
        char*    stack = nullptr;
        Cta*     cta = nullptr;
        unsigned tid = 0;
        bool     done = false;
        Wait     wait = Wait::None;
        uint64_t wait_gen = 0;
        // the stack belongs to the fiber: a CTA taken from the pool for a smaller block gives the surplus back
        // (kernels of 256 and of 32 threads alternate in the resumable decode: 224 leaked stacks per call otherwise)
        Fiber() = default;
        Fiber(const Fiber&) = delete;
        Fiber& operator=(const Fiber&) = delete;
        Fiber(Fiber&& o) noexcept : sp(o.sp), stack(o.stack), cta(o.cta), tid(o.tid), done(o.done), wait(o.wait), wait_gen(o.wait_gen) { o.stack = nullptr; }
        Fiber& operator=(Fiber&& o) noexcept
        {
            if (this != &o) {
                release();
                sp = o.sp, stack = o.stack, cta = o.cta, tid = o.tid, done = o.done, wait = o.wait, wait_gen = o.wait_gen;
                o.stack = nullptr;
            }
            return *this;
        }
        ~Fiber() { release(); }
        inline void release();
    };

    inline void Fiber::release()
    {
        if (stack) munmap(stack, kStack);
        stack = nullptr;
    }

    struct WarpRv {  // rendezvous state of one warp
        uint32_t arrived = 0, mask = 0;
        int      op = 0;
        uint64_t gen = 0;
        uint64_t arg[32], arg2[32], res[32];
    };

    struct Cta {
        unsigned            bid = 0;
        std::vector<Fiber>  fibers;
        std::vector<WarpRv> warps;
        std::vector<uint8_t> smem;
        unsigned            bar_arrived = 0, bar_or = 0, bar_and = 1, bar_cnt = 0, bar_res_or = 0, bar_res_and = 0, bar_res_cnt = 0;
        uint64_t            bar_gen = 0;
        unsigned            live = 0;
    };

    struct Ctx {
        Fiber*                 cur = nullptr;
        void*                  sched_sp = nullptr;
        dim3                   grid, block;
        std::function<void()>  body;
        std::mt19937_64        rng{ 12345 };
        bool                   random_order = false;
        uint64_t               switches = 0;
        uint64_t               grid_gen = 0;
        unsigned               grid_arrived = 0, grid_expected = 0;
    };
    inline Ctx& ctx() { static Ctx c; return c; }

    extern "C" void emu_switch(void** save_sp, void* new_sp);
    asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

    inline void yield_to_sched() { Ctx& c = ctx(); ++c.switches; emu_switch(&c.cur->sp, c.sched_sp); }

    inline void trampoline()
    {
        Ctx& c = ctx();
        c.body();
        c.cur->done = true;
        --c.cur->cta->live;
        yield_to_sched();
        abort();
    }

    inline void fiber_init(Fiber& f)
    {
        if (!f.stack) {
            f.stack = (char*)mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (f.stack == (char*)MAP_FAILED) { perror("mmap"); abort(); }
        }
        uintptr_t top = ((uintptr_t)f.stack + kStack) & ~uintptr_t(15);
        void**    sp  = (void**)top;
        *--sp = nullptr;                // fake return address of trampoline
        *--sp = (void*)&trampoline;     // popped by `ret`
        for (int i = 0; i < 6; ++i) *--sp = nullptr;
        f.sp = sp;
        f.done = false;
        f.wait = Wait::None;
    }

    // ---- launch: at most `resident` CTAs alive; CTA ids handed out in increasing order
    inline void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body, int resident = 6, uint64_t seed = 0)
    {
        Ctx& c = ctx();
        c.grid = grid; c.block = block; c.body = std::move(body);
        c.random_order = seed != 0;
        c.rng.seed(seed ? seed : 1);
        const unsigned nthreads = block.x, nwarps = (nthreads + 31) / 32, total = grid.x;
        c.grid_arrived = 0; c.grid_expected = total * nthreads;
        static std::vector<Cta*> pool;  // fiber stacks are reused across launches
        std::vector<Cta*> live;
        unsigned next = 0;
        auto start_cta = [&](Cta* k) {
            k->bid = next++;
            k->fibers.resize(nthreads);
            k->warps.assign(nwarps, WarpRv{});
            k->smem.assign(smem_bytes + 128, 0xCD);
            k->bar_arrived = 0; k->bar_gen = 0; k->bar_or = 0; k->bar_and = 1; k->bar_cnt = 0;
            k->live = nthreads;
            for (unsigned t = 0; t < nthreads; ++t) { Fiber& f = k->fibers[t]; f.cta = k; f.tid = t; fiber_init(f); }
        };
        while (next < total || !live.empty()) {
            while (next < total && (int)live.size() < resident) {
                Cta* k;
                if (!pool.empty()) { k = pool.back(); pool.pop_back(); } else k = new Cta();
                start_cta(k);
                live.push_back(k);
            }
            // one scheduling sweep: run every runnable fiber once
            bool progressed = false, only_spin = true;
            size_t n = live.size();
            size_t first = c.random_order ? c.rng() % n : 0;
            for (size_t q = 0; q < n; ++q) {
                Cta* k = live[(first + q) % n];
                unsigned nt = (unsigned)k->fibers.size();
                unsigned off = c.random_order ? (unsigned)(c.rng() % nt) : 0;
                for (unsigned i = 0; i < nt; ++i) {
                    Fiber& f = k->fibers[(i + off) % nt];
                    if (f.done) continue;
                    if (f.wait == Wait::Warp && k->warps[f.tid / 32].gen == f.wait_gen) continue;
                    if (f.wait == Wait::Block && k->bar_gen == f.wait_gen) continue;
                    if (f.wait == Wait::Grid && c.grid_gen == f.wait_gen) continue;
                    bool was_spin = f.wait == Wait::Spin;
                    f.wait = Wait::None;
                    c.cur = &f;
                    emu_switch(&c.sched_sp, f.sp);
                    c.cur = nullptr;
                    progressed = true;
                    if (!(was_spin && f.wait == Wait::Spin)) only_spin = false;
                }
            }
            for (size_t q = 0; q < live.size();) {
                if (live[q]->live == 0) { pool.push_back(live[q]); live[q] = live.back(); live.pop_back(); only_spin = false; }
                else ++q;
            }
            if (!progressed && !live.empty()) {
                fprintf(stderr, "cuda_emu: DEADLOCK: %zu CTAs alive, nothing runnable (divergent collective or missing barrier)\n", live.size());
                for (Cta* k : live) {
                    fprintf(stderr, "  cta %u: live=%u bar_arrived=%u\n", k->bid, k->live, k->bar_arrived);
                    for (size_t w = 0; w < k->warps.size(); ++w)
                        if (k->warps[w].arrived) fprintf(stderr, "    warp %zu arrived=%08x mask=%08x op=%d\n", w, k->warps[w].arrived, k->warps[w].mask, k->warps[w].op);
                }
                abort();
            }
            static int spin_rounds = 0;
            if (progressed && only_spin && next >= total) {
                if (++spin_rounds > 2000000) { fprintf(stderr, "cuda_emu: LIVELOCK: all fibers spin-wait forever\n"); abort(); }
            } else spin_rounds = 0;
        }
    }

    // ---- collectives
    enum Op { OP_SYNCWARP = 1, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_BALLOT, OP_MATCH, OP_ANY, OP_ALL, OP_RED_OR, OP_RED_ADD, OP_RED_MAX, OP_RED_MIN, OP_RED_AND };

    inline uint64_t warp_collective(int op, uint32_t mask, uint64_t a, uint64_t b)
    {
        Fiber&   f    = *ctx().cur;
        unsigned lane = f.tid & 31;
        WarpRv&  w    = f.cta->warps[f.tid / 32];
        unsigned nth  = ctx().block.x;
        uint32_t exist = (f.tid / 32 == (nth - 1) / 32 && (nth & 31)) ? ((1u << (nth & 31)) - 1) : 0xffffffffu;
        mask &= exist;
        if (!(mask >> lane & 1)) { fprintf(stderr, "cuda_emu: lane %u not in its own mask %08x (op %d)\n", lane, mask, op); abort(); }
        if (w.arrived == 0) { w.mask = mask; w.op = op; }
        else if (w.mask != mask || w.op != op) {
            fprintf(stderr, "cuda_emu: MISMATCHED COLLECTIVE in cta %u warp %u: lane %u calls op %d mask %08x, pending op %d mask %08x arrived %08x\n",
                    f.cta->bid, f.tid / 32, lane, op, mask, w.op, w.mask, w.arrived);
            abort();
        }
        w.arrived |= 1u << lane;
        w.arg[lane] = a; w.arg2[lane] = b;
        if (w.arrived == w.mask) {
            for (unsigned l = 0; l < 32; ++l) {
                if (!(mask >> l & 1)) continue;
                uint64_t r = 0;
                switch (op) {
                case OP_SYNCWARP: break;
                case OP_SHFL: { unsigned s = (unsigned)w.arg2[l] & 31; r = (mask >> s & 1) ? w.arg[s] : w.arg[l]; } break;
                case OP_SHFL_UP: { int s = (int)l - (int)w.arg2[l]; r = (s >= 0 && (mask >> s & 1)) ? w.arg[s] : w.arg[l]; } break;
                case OP_SHFL_DOWN: { unsigned s = l + (unsigned)w.arg2[l]; r = (s < 32 && (mask >> s & 1)) ? w.arg[s] : w.arg[l]; } break;
                case OP_SHFL_XOR: { unsigned s = l ^ (unsigned)w.arg2[l]; r = (s < 32 && (mask >> s & 1)) ? w.arg[s] : w.arg[l]; } break;
                case OP_BALLOT: for (unsigned j = 0; j < 32; ++j) if ((mask >> j & 1) && w.arg[j]) r |= 1u << j; break;
                case OP_MATCH: for (unsigned j = 0; j < 32; ++j) if ((mask >> j & 1) && w.arg[j] == w.arg[l]) r |= 1u << j; break;
                case OP_ANY: for (unsigned j = 0; j < 32; ++j) if ((mask >> j & 1) && w.arg[j]) r = 1; break;
                case OP_ALL: r = 1; for (unsigned j = 0; j < 32; ++j) if ((mask >> j & 1) && !w.arg[j]) r = 0; break;
                case OP_RED_OR: for (unsigned j = 0; j < 32; ++j) if (mask >> j & 1) r |= w.arg[j]; break;
                case OP_RED_AND: r = ~0ull; for (unsigned j = 0; j < 32; ++j) if (mask >> j & 1) r &= w.arg[j]; break;
                case OP_RED_ADD: for (unsigned j = 0; j < 32; ++j) if (mask >> j & 1) r += w.arg[j]; break;
                case OP_RED_MAX: r = 0; for (unsigned j = 0; j < 32; ++j) if ((mask >> j & 1) && w.arg[j] > r) r = w.arg[j]; break;
                case OP_RED_MIN: r = ~0ull; for (unsigned j = 0; j < 32; ++j) if ((mask >> j & 1) && w.arg[j] < r) r = w.arg[j]; break;
                }
                w.res[l] = r;
            }
            w.arrived = 0;
            ++w.gen;
            return w.res[lane];
        }
        f.wait = Wait::Warp; f.wait_gen = w.gen;
        yield_to_sched();
        return w.res[lane];
    }

    inline unsigned block_barrier(int pred)
    {
        Fiber& f = *ctx().cur;
        Cta&   k = *f.cta;
        if (k.live != k.fibers.size()) { fprintf(stderr, "cuda_emu: __syncthreads after some threads of cta %u exited\n", k.bid); abort(); }
        k.bar_or |= (pred != 0); k.bar_and &= (pred != 0); k.bar_cnt += (pred != 0);
        if (++k.bar_arrived == k.fibers.size()) {
            k.bar_res_or = k.bar_or; k.bar_res_and = k.bar_and; k.bar_res_cnt = k.bar_cnt;
            k.bar_arrived = 0; k.bar_or = 0; k.bar_and = 1; k.bar_cnt = 0;
            ++k.bar_gen;
        } else {
            f.wait = Wait::Block; f.wait_gen = k.bar_gen;
            yield_to_sched();
        }
        return 0;
    }

    // cooperative-groups grid barrier: every thread of every CTA of the launch (all CTAs must be resident)
    inline void grid_barrier()
    {
        Ctx&   c = ctx();
        Fiber& f = *c.cur;
        if (++c.grid_arrived == c.grid_expected) {
            c.grid_arrived = 0;
            ++c.grid_gen;
        } else {
            f.wait = Wait::Grid; f.wait_gen = c.grid_gen;
            yield_to_sched();
        }
    }

    inline void spin_yield()
    {
        Fiber& f = *ctx().cur;
        f.wait = Wait::Spin;
        yield_to_sched();
    }
}  // namespace emu

// ---------------------------------------------------------------- built-in variables
#define threadIdx (uint3{ emu::ctx().cur->tid, 0, 0 })
#define blockIdx (uint3{ emu::ctx().cur->cta->bid, 0, 0 })
#define blockDim (uint3{ emu::ctx().block.x, 1, 1 })
#define gridDim (uint3{ emu::ctx().grid.x, 1, 1 })
#define warpSize 32
#define QB_DYN_SMEM (emu::ctx().cur->cta->smem.data() + ((128 - ((uintptr_t)emu::ctx().cur->cta->smem.data() & 127)) & 127))
#define QB_SPIN_YIELD() emu::spin_yield()
#define QB_GRID_SYNC() emu::grid_barrier()

// ---------------------------------------------------------------- intrinsics
static inline void     __syncthreads() { emu::block_barrier(0); }
static inline int      __syncthreads_or(int p) { emu::block_barrier(p); return emu::ctx().cur->cta->bar_res_or; }
static inline int      __syncthreads_and(int p) { emu::block_barrier(p); return emu::ctx().cur->cta->bar_res_and; }
static inline int      __syncthreads_count(int p) { emu::block_barrier(p); return emu::ctx().cur->cta->bar_res_cnt; }
static inline void     __syncwarp(unsigned m = 0xffffffffu) { emu::warp_collective(emu::OP_SYNCWARP, m, 0, 0); }
static inline void     __threadfence() {}
static inline void     __threadfence_block() {}
static inline void     __nanosleep(unsigned) {}
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) { uint64_t a = 0; memcpy(&a, &v, sizeof(T)); uint64_t r = emu::warp_collective(emu::OP_SHFL, m, a, (unsigned)src); T o; memcpy(&o, &r, sizeof(T)); return o; }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { uint64_t a = 0; memcpy(&a, &v, sizeof(T)); uint64_t r = emu::warp_collective(emu::OP_SHFL_UP, m, a, d); T o; memcpy(&o, &r, sizeof(T)); return o; }
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) { uint64_t a = 0; memcpy(&a, &v, sizeof(T)); uint64_t r = emu::warp_collective(emu::OP_SHFL_DOWN, m, a, d); T o; memcpy(&o, &r, sizeof(T)); return o; }
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int d) { uint64_t a = 0; memcpy(&a, &v, sizeof(T)); uint64_t r = emu::warp_collective(emu::OP_SHFL_XOR, m, a, (unsigned)d); T o; memcpy(&o, &r, sizeof(T)); return o; }
static inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)emu::warp_collective(emu::OP_BALLOT, m, p != 0, 0); }
static inline int      __any_sync(unsigned m, int p) { return (int)emu::warp_collective(emu::OP_ANY, m, p != 0, 0); }
static inline int      __all_sync(unsigned m, int p) { return (int)emu::warp_collective(emu::OP_ALL, m, p != 0, 0); }
template <typename T> static inline unsigned __match_any_sync(unsigned m, T v) { uint64_t a = 0; memcpy(&a, &v, sizeof(T)); return (unsigned)emu::warp_collective(emu::OP_MATCH, m, a, 0); }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)emu::warp_collective(emu::OP_RED_OR, m, v, 0); }
static inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return (unsigned)emu::warp_collective(emu::OP_RED_AND, m, v, 0); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)emu::warp_collective(emu::OP_RED_ADD, m, v, 0); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)emu::warp_collective(emu::OP_RED_MAX, m, v, 0); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)emu::warp_collective(emu::OP_RED_MIN, m, v, 0); }

static inline int      __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int      __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int      __popc(unsigned x) { return __builtin_popcount(x); }
static inline int      __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int      __ffs(int x) { return __builtin_ffs(x); }
static inline int      __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s > 32 ? 32 : s)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)((v << (s & 31)) >> 32); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) { for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u); return c; }
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) { return c + (a & 0xFFFFu) * (b & 255u) + (a >> 16) * ((b >> 8) & 255u); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    uint64_t v = ((uint64_t)y << 32) | x; unsigned r = 0;
    for (int i = 0; i < 4; ++i) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * sel)) & 255u) << (8 * i); }
    return r;
}
static inline unsigned __vsub4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; ++i) r |= (((a >> (8 * i)) - (b >> (8 * i))) & 255u) << (8 * i); return r; }
static inline unsigned __vadd4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; ++i) r |= (((a >> (8 * i)) + (b >> (8 * i))) & 255u) << (8 * i); return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }

// atomics (single OS thread, fibers switch only at yield points => plain ops are atomic)
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline unsigned atomicInc(unsigned* p, unsigned v) { unsigned o = *p; *p = (o >= v) ? 0 : o + 1; return o; }
