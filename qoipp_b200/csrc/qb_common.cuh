// qb_common.cuh -- shared device-side definitions of the qoipp_b200 kernels (sm_100a).
//
// The kernels are single-pass chained scans: every tile publishes small "aggregate" words, looks back over
// its predecessors' words and then publishes "inclusive" words (decoupled look-back).  Every word that
// crosses CTAs is ONE 64-bit value carrying {payload, status, launch epoch}, so a relaxed 64-bit load sees a
// consistent snapshot and no memset is needed between launches (a stale epoch reads as "not ready").
//
// The same sources compile under tests/emu/cuda_emu.h (QB_EMU) so the kernel logic can be stepped on a CPU.
#pragma once

#include <stdint.h>

#ifndef QB_EMU
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#define QB_GRID_SYNC() cooperative_groups::this_grid().sync()
#ifndef QB_SPIN_NS
#define QB_SPIN_NS 0  // measured (tools/ab_probe.py): 0 / 32 / 100 / 250 / 600 ns -> decode 4K 272 / 283 / 287 / 289 / 296 us
#endif
#define QB_SPIN_YIELD() __nanosleep(QB_SPIN_NS)
#define QB_DYN_SMEM qb_dyn_smem
extern __shared__ __align__(128) unsigned char qb_dyn_smem[];
#endif

namespace qb
{
    constexpr unsigned kFull = 0xffffffffu;

    // ---- QOI constants (reference: source/util.hpp:27-55, include/qoipp/common.hpp:17-23)
    constexpr unsigned kOpIndex = 0x00, kOpDiff = 0x40, kOpLuma = 0x80, kOpRun = 0xC0, kOpRgb = 0xFE, kOpRgba = 0xFF;
    constexpr unsigned kHeader = 14, kMarker = 8, kRunLimit = 62;
    constexpr unsigned kStartPixel = 0xFF000000u;  // {0,0,0,255} as little-endian r|g<<8|b<<16|a<<24

    // util::hash (source/util.hpp:347-351) reduced mod 64: one dp4a on the packed pixel
    __device__ __forceinline__ unsigned slot_of(unsigned px) { return __dp4a(px, 0x0B070503u, 0u) & 63u; }

    // a value that can never be stored in table slot `s` (its own slot differs): marks "no writer yet"
    __device__ __forceinline__ unsigned sentinel(unsigned s) { return s == 0 ? 1u : 0u; }

    // ---- cross-CTA words.  Layout: [31:0] payload-lo | [33:32] status | [53:34] epoch | [63:54] payload-hi(10b)
    // 42-bit payloads (byte offsets, pixel indices) use lo + hi.
    enum : unsigned { ST_NONE = 0, ST_AGG_EMPTY = 1, ST_AGG = 2, ST_INCL = 3 };
    constexpr unsigned kEpochBits = 20, kEpochMask = (1u << kEpochBits) - 1;

    __device__ __forceinline__ uint64_t pack_word(uint64_t payload42, unsigned status, unsigned epoch)
    {
        return (payload42 & 0xffffffffull) | ((uint64_t)status << 32) | ((uint64_t)(epoch & kEpochMask) << 34) |
               ((payload42 >> 32) << 54);
    }
    __device__ __forceinline__ unsigned word_status(uint64_t w, unsigned epoch)
    {
        return (((unsigned)(w >> 34)) & kEpochMask) == (epoch & kEpochMask) ? ((unsigned)(w >> 32) & 3u) : ST_NONE;
    }
    __device__ __forceinline__ uint64_t word_payload(uint64_t w) { return (w & 0xffffffffull) | ((w >> 54) << 32); }

    __device__ __forceinline__ uint64_t ld_word(const uint64_t* p)
    {
#ifdef QB_EMU
        return *p;
#else
        uint64_t v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        return v;
#endif
    }
    __device__ __forceinline__ void st_word(uint64_t* p, uint64_t v)
    {
#ifdef QB_EMU
        *p = v;
#else
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
    }
    // spin until the word is valid for this epoch
    __device__ __forceinline__ uint64_t wait_word(const uint64_t* p, unsigned epoch)
    {
        uint64_t w = ld_word(p);
        while (word_status(w, epoch) == ST_NONE) {
            QB_SPIN_YIELD();
            w = ld_word(p);
        }
        return w;
    }

    // shuffles of trivially copyable values made of 32-bit words
    template <class T>
    __device__ __forceinline__ T shfl_down_T(const T& v, unsigned d)
    {
        static_assert(sizeof(T) % 4 == 0, "word multiple");
        T        r;
        unsigned a[sizeof(T) / 4];
        __builtin_memcpy(a, &v, sizeof(T));
#pragma unroll
        for (unsigned i = 0; i < sizeof(T) / 4; ++i) a[i] = __shfl_down_sync(kFull, a[i], d);
        __builtin_memcpy(&r, a, sizeof(T));
        return r;
    }
    template <class T>
    __device__ __forceinline__ T shfl_T(const T& v, int src)
    {
        static_assert(sizeof(T) % 4 == 0, "word multiple");
        T        r;
        unsigned a[sizeof(T) / 4];
        __builtin_memcpy(a, &v, sizeof(T));
#pragma unroll
        for (unsigned i = 0; i < sizeof(T) / 4; ++i) a[i] = __shfl_sync(kFull, a[i], src);
        __builtin_memcpy(&r, a, sizeof(T));
        return r;
    }

    // ---- which launch epochs a reader accepts for the words of tile p.
    // A decode is up to kRounds + 1 launches with consecutive epochs base, base+1, ...: in round r the tiles from
    // `fresh_from` on are recomputed and must be read at epoch `cur`; earlier tiles are final and carry any epoch of this
    // decode.  Encode (and round 0) use fresh_from = 0, i.e. an exact match.
    struct Epochs {
        unsigned cur, base, fresh_from;
        __device__ __forceinline__ bool valid(uint64_t w, unsigned p) const
        {
            const unsigned e = ((unsigned)(w >> 34)) & kEpochMask;
            if (((unsigned)(w >> 32) & 3u) == ST_NONE) return false;
            return p >= fresh_from ? e == (cur & kEpochMask) : ((e - base) & kEpochMask) < ((cur - base) & kEpochMask);
        }
    };
    __device__ __forceinline__ uint64_t wait_word(const uint64_t* ptr, const Epochs& ep, unsigned p)
    {
        uint64_t w = ld_word(ptr);
        while (!ep.valid(w, p)) {
            QB_SPIN_YIELD();
            w = ld_word(ptr);
        }
        return w;
    }
    __device__ __forceinline__ unsigned raw_status(uint64_t w) { return (unsigned)(w >> 32) & 3u; }

    // ---- warp-cooperative decoupled look-back (all 32 lanes of ONE warp call this together)
    // Lane l inspects predecessor base - l of tile t.  `fetch(p, st)` waits for tile p's word and returns its value with
    // st = ST_AGG / ST_AGG_EMPTY (the tile's own aggregate; EMPTY values are ignored) or ST_INCL (inclusive prefix).
    // The virtual tile -1 is inclusive with value `init`.  Returns, in every lane, the exclusive prefix of tile t.
    // `comb(earlier, later)` must be associative with identity `empty`.
    template <class T, class Fetch, class Comb>
    __device__ __forceinline__ T warp_lookback(unsigned t, T init, T empty, Fetch fetch, Comb comb)
    {
        const unsigned lane = threadIdx.x & 31u;
        T              acc  = empty;
        for (int base = (int)t - 1;; base -= 32) {
            const int p  = base - (int)lane;
            T         v  = init;
            unsigned  st = ST_INCL;
            if (p >= 0) {
                v = fetch((unsigned)p, st);
                if (st == ST_AGG_EMPTY) v = empty;
            }
            const unsigned incl  = __ballot_sync(kFull, st == ST_INCL);
            const unsigned first = incl ? (unsigned)__ffs((int)incl) - 1u : 31u;  // nearest inclusive predecessor
            if (lane > first) v = empty;
            // ordered fold: lane 0 is the latest tile, higher lanes are earlier
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T o = shfl_down_T(v, d);
                if (lane + d < 32u) v = comb(o, v);
            }
            const T round = shfl_T(v, 0);
            acc           = comb(round, acc);
            if (incl) break;
        }
        return acc;
    }

    // Same contract as warp_lookback, but `peek(p, st)` never waits: it returns tile p's word as it is now (st = ST_NONE when
    // not published yet).  A round is accepted as soon as every predecessor CLOSER than the nearest inclusive one has
    // published; farther tiles are never waited for, so a slow tile 30 places back does not hold this one up when a nearer
    // tile already knows its inclusive prefix.
    template <class T, class Peek, class Comb>
    __device__ __forceinline__ T warp_lookback_lazy(unsigned t, T init, T empty, Peek peek, Comb comb)
    {
        const unsigned lane = threadIdx.x & 31u;
        T              acc  = empty;
        for (int base = (int)t - 1;; base -= 32) {
            const int p = base - (int)lane;
            T         v;
            unsigned  st, incl, first;
            for (;;) {
                v = init, st = ST_INCL;
                if (p >= 0) v = peek((unsigned)p, st);
                incl  = __ballot_sync(kFull, st == ST_INCL);
                first = incl ? (unsigned)__ffs((int)incl) - 1u : 32u;  // nearest inclusive predecessor of this window
                const unsigned pending = __ballot_sync(kFull, st == ST_NONE) & (first >= 32u ? kFull : (1u << first) - 1u);
                if (pending == 0) break;
                QB_SPIN_YIELD();
            }
            if (first == 0u) return comb(shfl_T(v, 0), acc);  // the direct predecessor is inclusive already: no fold
            if (st == ST_AGG_EMPTY || st == ST_NONE || lane > first) v = empty;
            // ordered fold over lanes 0 .. first (lane 0 is the latest tile): only as many steps as that span needs
            for (unsigned d = 1; d <= min(first, 31u); d <<= 1) {
                const T o = shfl_down_T(v, d);
                if (lane + d < 32u) v = comb(o, v);
            }
            const T round = shfl_T(v, 0);
            acc           = comb(round, acc);
            if (incl) break;
        }
        return acc;
    }

    // development aid (tools/phase_probe.py): with -DQB_TIMING thread 0 of every CTA stamps SM cycles at phase boundaries
    // into the padding words of its tile's carry record
#if defined(QB_TIMING) && !defined(QB_EMU)
#define QB_STAMP(desc_ptr, word, slot, t0)                                                        \
    do {                                                                                          \
        if (threadIdx.x == 0) reinterpret_cast<unsigned*>((desc_ptr) + (word))[slot] = (unsigned)(clock64() - (t0)); \
    } while (0)
#define QB_T0() clock64()
#else
#define QB_STAMP(desc_ptr, word, slot, t0) do { } while (0)
#define QB_T0() 0ll
#endif

    __device__ __forceinline__ unsigned lanemask_lt(unsigned lane) { return (1u << lane) - 1u; }
    __device__ __forceinline__ unsigned lanemask_gt(unsigned lane) { return lane == 31 ? 0u : ~((2u << lane) - 1u); }

    // lanes of the warp whose (6-bit) slot equals this lane's, among the `active` ones.  Measured on B200 (tools/ab_probe.py):
    // seven ballots instead of one __match_any_sync are no faster (4K RGB 129 vs 131 us, 8K RGBA 478 vs 459 us): the
    // vote pipe (ADU) is already the busiest one, so the single MATCH instruction stays.
    __device__ __forceinline__ unsigned match_slot(bool active, unsigned slot)
    {
#ifndef QB_MATCH_BY_BALLOTS
        const unsigned lane = threadIdx.x & 31u;
        const unsigned m    = __match_any_sync(kFull, active ? slot : 64u + lane);
        return active ? m : 0u;
#else
        unsigned m = __ballot_sync(kFull, active);
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const unsigned v = __ballot_sync(kFull, (slot >> b) & 1u);
            m &= ((slot >> b) & 1u) ? v : ~v;
        }
        return active ? m : 0u;
#endif
    }

    // bytewise (mod 256) subtract / add on packed pixels
    __device__ __forceinline__ unsigned sub4(unsigned a, unsigned b)
    {
        return ((a | 0x80808080u) - (b & 0x7f7f7f7fu)) ^ ((a ^ ~b) & 0x80808080u);
    }
    __device__ __forceinline__ unsigned add4(unsigned a, unsigned b)
    {
        return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
    }
}  // namespace qb
