// decode_wt.cuh -- warp-autonomous tile decoder for sm_100a (+ the exact sequential loop it falls back on).
//
// Replaces the serial loop of the reference, impl::decode (source/simple.cpp:100-171).  A tile is kDecTB = 32 * kWtChunk
// bytes of the chunk stream and belongs to ONE WARP: the warps of a CTA are independent persistent workers that draw
// tiles from a ticket counter, nothing in the kernel needs __syncthreads.  A lane owns kWtChunk consecutive stream bytes
// and walks its ops like the reference loop does -- cursor, delta accumulator, slot hash and pixel counter in registers.
// What a lane cannot know at the start of its chunk is carried symbolically:
//
//   parse   an op's length depends on its tag byte only (simple.cpp:118-165): a chunk is a map {entry offset 0..4 -> exit
//           offset}; maps compose across lanes (warp scan) and tiles (look-back 1).  No self-synchronisation is assumed.
//           The same walk counts ops / pixels and finds the last OP_RGBA, so op ordinals, pixel offsets and the inherited
//           alpha of every lane are known BEFORE the main walk (warp scans + look-back 2, pixels and alpha in one word).
//   walk    ONE pass over the lane's ops writes a record per op, compact in stream order (op ordinal k):
//             rec[k]  value relative to the op's base: absolute after a literal (OP_RGB / OP_RGBA), a delta otherwise
//             base[k] the base node: ABS, the lane's entry node E_l (ops before the lane's first literal / OP_INDEX),
//                     or the ordinal of the OP_INDEX op that roots the segment
//             slot[k] table slot of the value (util::hash is linear mod 64, util.hpp:347-351: slots follow from the
//                     literals by sums of 3dr+5dg+7db without knowing pixel values) | flags;  pix[k] pixel offset
//   nodes   E_l and the OP_INDEX ops are the only nodes that point at other ops (E_l = the op before the lane's first,
//           OP_INDEX = the last earlier op with the same slot, found by a backward search over slot[]).  Pointer jumping
//           over these few nodes ends in ABS or in an entry of the state entering the tile (EXT: prev or a table slot).
//   state   look-back 4: per tile a transfer function -- each of the 64 table slots and `prev` leaves the tile as a
//           constant or as (incoming entry) + delta.  Only the entries a tile really reads are followed through the
//           predecessors' words (lazy: an OP_INDEX-free tile reads `prev` alone).
//   emit    data-parallel over the compact records (lane = op): value = rec[k] (+ rec[base]), verification, pixels stored
//           straight to global memory (consecutive lanes = consecutive pixels), OP_RUN expanded by the whole warp.
//
// SPECULATION (unchanged from round 1).  The slot of an OP_RGB pixel needs its inherited alpha: assumed "alpha of the last
// OP_RGBA before it, else 255" (+ alphas learned by earlier rounds), and an OP_INDEX is assumed to read a written slot.
// Every tile VERIFIES both on the final values (OP_RGB: alpha equals its predecessor's; OP_INDEX: hash(value) == slot).
// Verified => exact, by induction over op order.  A refuted tile records the alphas it saw; decode_finish_kernel re-decodes
// from the first refuted tile on (up to kDecRounds rounds) and hands what still fails to the sequential loop.
#pragma once

#include "qb_common.cuh"

#include <type_traits>

// The decode kernel is bounded by instruction fetch when its hot path does not fit the SM's 32 KB instruction cache
// (profiles/r02_experiments.md: 98 KB of SASS, hit rate 88 %, time proportional to the number of tiles whatever the
// occupancy).  Helpers that are called once per tile, or are cold, are real calls instead of inlined copies.
#define QB_NOINLINE __noinline__
#ifdef QB_STATS  // development build: event counters in DecControl::stats (tools/stats_probe.py)
#define QB_COUNT(P, i) do { if ((threadIdx.x & 31u) == 0) atomicAdd(&(P).control->stats[i], 1u); } while (0)
#else
#define QB_COUNT(P, i) do { } while (0)
#endif
#if defined(QB_EMU) || defined(QB_NO_REPAIR_FENCE)
#define QB_FENCE_SC() do { } while (0)
#else
#define QB_FENCE_SC() __threadfence()  // fence.sc.gpu
#endif

namespace qb
{
    struct DecState {  // == StreamDecoder members (include/qoipp/stream.hpp:239-243), pixels packed
        uint32_t prev;
        uint32_t run;
        uint32_t table[64];
    };

    constexpr int kDecRounds = 4;  // verification-driven retry rounds before the sequential loop takes over

    struct DecResult {
        uint32_t bad;   // round 0 refuted a speculation somewhere in this image
        uint32_t path;  // retry rounds used; + 100 when the sequential loop produced (part of) the image
        uint64_t pixels;
        uint64_t processed;  // resumable decode: input bytes consumed / output bytes written / carry-out
        uint64_t written;
        uint32_t first_bad[kDecRounds + 1];  // per round: 0 = all verified, else 0xFFFFFFFF - first refuted tile
        uint32_t pad[3];
        DecState state;
    };

    // round 0 -> cascade (decode_finish_body): a tile that consumed a word its repaired predecessor retracted, with the table
    // entries / prev that had changed on entering it
    struct CascadeReq {
        uint32_t img, tile, lo, hi, prev, pad;
    };

    struct DecControl {  // zeroed with the results before every decode
        uint32_t tickets[kDecRounds + 1];  // tile tickets of round 0 and of the retry rounds
        uint32_t any_bad[kDecRounds + 1];  // some image needs round r + 1
        uint32_t rounds_needed;            // some image needs retry round 1 (a tile stayed refuted, or its cascade did not end)
        uint32_t n_req;                    // cascade requests of round 0 (may exceed the capacity: the surplus went to the rounds)
        uint32_t list_n, list_tiles;       // work list of a retry round: images on it, tiles in their ranges
        uint32_t stats[2];                 // development builds (QB_STATS)
    };

    struct DecParams {
        const uint8_t*  qoi;
        const uint64_t* offsets;     // [n_images][2] device: first byte of stream k, one past its last; null => single[]
        const uint32_t* tile_first;  // [n_images + 1] device; null => single image
        uint64_t        single[2];
        uint8_t*        out;
        uint64_t        out_stride;
        uint64_t        n_pixels;
        uint32_t        width, height, target, flip;
        uint32_t        n_images, n_tiles, epoch, round;  // epoch = first epoch of this decode, round r uses epoch + r
        DecResult*      results;
        DecControl*     control;
        uint64_t*       desc;
        uint32_t*       fix;  // [n_tiles][kFixWords]: alpha learned at OP_RGB ops by earlier rounds
        CascadeReq*     req;      // [req_cap] cascade requests of round 0
        uint32_t        req_cap;
        const DecState* init;  // resumable decode (wt_decode_tile<true>): the state the chunk stream is entered with; null otherwise
    };

#ifndef QB_WT_CHUNK
#define QB_WT_CHUNK 28  // stream bytes per lane: an odd number of words keeps the lanes' chunk reads on different banks
#endif
#ifndef QB_WT_WARPS
#define QB_WT_WARPS 8  // warp workers per CTA
#endif
#ifndef QB_WT_LOCKSTEP
#define QB_WT_LOCKSTEP 1  // the warps of a CTA start their tiles together (instruction cache, see decode_wt_kernel)
#endif
#ifndef QB_WT_CTAS
#define QB_WT_CTAS 3  // CTAs per SM the kernels are compiled for (24 warps, <= 85 registers)
#endif
    constexpr int kWtChunk = QB_WT_CHUNK, kDecTB = 32 * kWtChunk, kWtWarps = QB_WT_WARPS, kWtThreads = kWtWarps * 32;
    static_assert(kWtChunk >= 8 && kWtChunk <= 28 && kWtChunk % 4 == 0, "op-start masks are 32 bits wide");

    // node ids: op ordinals 0 .. kDecTB - 1, lane entry nodes, entries of the state entering the tile, "absolute"
    constexpr unsigned kIdE = kDecTB, kIdExt = kDecTB + 32, kIdAbs = 0xFFFFu, kNoOp = 0xFFFFu, kNoPos = 0xFFFFu;
    // "absolute r, g, b; alpha = the alpha entering the tile" (an OP_RGB before the tile's first alpha setter): a node whose
    // value is alpha << 24, so that a record based on it needs no special case; filled in when the alpha is gathered (late)
    constexpr unsigned kIdAbsA = kIdExt + 66u;
    constexpr int      kWtNodes = kDecTB + 32 + 67;
    constexpr unsigned kSlRgb = 0x40u, kSlIdx = 0x80u;  // flags beside the 6-bit slot

#ifdef QB_TIMING
    constexpr int kDecDescWords = 80;  // development build: words 72..79 hold phase stamps (tools/phase_probe_wt.py), no top-level totals
#else
    constexpr int kDecDescWords = 73;
#endif
    constexpr int kDwParse = 0, kDwPixA = 1, kDwSlot = 2, kDwState = 3;  // 3..67: 64 table entries, then prev
    constexpr int kDwNeedLo = 70, kDwNeedHi = 71;  // entries of the incoming state this tile read (slots 0..31 / 32..63, bit 32: prev)
    constexpr int kDwGrp = 68, kDwSup = 69, kDwTop = 72;  // totals of the 32 / 1024 / 32768 tiles ending with this tile (see wt_gather_pixa)
        constexpr unsigned kGrp = 32, kSup = 32 * kGrp, kTop = 32 * kSup;  // every level folds <= 31 words: one warp step
    constexpr int kFixWords = 16, kFixMax = kFixWords - 1;  // word 0: count | decode tag << 8; entries: pos | alpha << 16

    // parse map of a byte range: exit offset for each of the five possible entry offsets.  Entries 0..3 live in the
    // bytes of `lo`, entry 4 in `hi`, so that composing two maps is two PRMT instructions.
    struct Map {
        unsigned lo, hi;
    };
    __device__ __forceinline__ Map map_identity() { return Map{ 0x03020100u, 4u }; }
    __device__ __forceinline__ Map map_const(unsigned e) { return Map{ e * 0x01010101u, e }; }
    __device__ __forceinline__ Map map_compose(const Map& f, const Map& g)  // f first, then g
    {
        const unsigned sel = (f.lo & 7u) | ((f.lo >> 4) & 0x70u) | ((f.lo >> 8) & 0x700u) | ((f.lo >> 12) & 0x7000u);
        return Map{ __byte_perm(g.lo, g.hi, sel), __byte_perm(g.lo, g.hi, f.hi & 7u) & 0xFFu };
    }
    __device__ __forceinline__ unsigned map_at(const Map& f, unsigned e) { return e < 4u ? (f.lo >> (8u * e)) & 7u : f.hi & 7u; }
    __device__ __forceinline__ unsigned map_pack(const Map& f)
    {
        return (f.lo & 7u) | ((f.lo >> 8) & 7u) << 3 | ((f.lo >> 16) & 7u) << 6 | ((f.lo >> 24) & 7u) << 9 | (f.hi & 7u) << 12;
    }
    __device__ __forceinline__ Map map_unpack(unsigned v)
    {
        return Map{ (v & 7u) | ((v >> 3) & 7u) << 8 | ((v >> 6) & 7u) << 16 | ((v >> 9) & 7u) << 24, (v >> 12) & 7u };
    }

    __device__ __forceinline__ unsigned op_length(unsigned tag)  // simple.cpp:118-165
    {
        return 1u + ((tag >> 6) == 2u) + 3u * (tag == kOpRgb) + 4u * (tag == kOpRgba);
    }

    // ---- per-tag table shared by the warps of a CTA (2 KB): what the walks need from an op's first byte in one LDS.64
    //   x: delta of r (bits 0..7) and b (bits 16..23) as two 16-bit lanes: OP_DIFF dr, db; OP_LUMA dg - 8 in both (the second
    //      byte's nibbles are added by the walk); mod 256 per lane
    //   y: dg (bits 0..7) | 3dr+5dg+7db mod 64 of that delta (8..13) | op length (14..16) | run - 1 of an OP_RUN (17..22) |
    //      OP_LUMA (23) | slot-byte flags (24..31): kSlRgb, kSlIdx, both for OP_RUN
    __device__ __forceinline__ uint2 wt_lut_entry(unsigned tag)
    {
        unsigned rb = 0, g = 0, fl = 0, runx = 0, luma = 0;
        const unsigned hi = tag >> 6;
        if (tag == kOpRgb) fl = kSlRgb;
        else if (tag == kOpRgba) fl = 0;
        else if (hi == 0u) fl = kSlIdx;
        else if (hi == 1u) {  // simple.cpp:136-144
            rb = ((((tag >> 4) & 3u) + 254u) & 255u) | ((((tag & 3u) + 254u) & 255u) << 16);
            g  = (((tag >> 2) & 3u) + 254u) & 255u;
        } else if (hi == 2u) {  // simple.cpp:145-155
            g    = (tag + 0x60u) & 255u;  // dg = (tag & 63) - 32
            rb   = ((g + 248u) & 255u) * 0x00010001u;
            luma = 1;
        } else {
            fl = kSlRgb | kSlIdx, runx = tag & 63u;  // simple.cpp:156-163
        }
        const unsigned hc = (3u * (rb & 255u) + 5u * g + 7u * ((rb >> 16) & 255u)) & 63u;
        return make_uint2(rb, g | hc << 8 | op_length(tag) << 14 | runx << 17 | luma << 23 | fl << 24);
    }
    __device__ __forceinline__ void wt_build_lut(uint2* lut)  // all threads of the CTA; ends with a barrier
    {
        for (unsigned i = threadIdx.x; i < 256u; i += blockDim.x) lut[i] = wt_lut_entry(i);
        __syncthreads();
    }
    constexpr size_t kWtLutBytes = 256 * sizeof(uint2);

    // pixels before a tile (33 bits, saturating) and the alpha of the last OP_RGBA (0x100 | alpha, 0 = none so far)
    struct PixA {
        unsigned lo, hi, a;
    };
    constexpr uint64_t kPixSat = (1ull << 33) - 1;
    __device__ __forceinline__ PixA pixa_comb(const PixA& x, const PixA& y)  // x happens before y
    {
        uint64_t s = ((uint64_t)x.hi << 32 | x.lo) + ((uint64_t)y.hi << 32 | y.lo);
        if (s > kPixSat) s = kPixSat;
        return PixA{ (unsigned)s, (unsigned)(s >> 32), (y.a & 0x100u) ? y.a : x.a };
    }
    __device__ __forceinline__ uint64_t pixa_pack(const PixA& x) { return ((uint64_t)x.hi << 32 | x.lo) | (uint64_t)(x.a & 0x1FFu) << 33; }
    __device__ __forceinline__ PixA     pixa_unpack(uint64_t v) { return PixA{ (unsigned)v, (unsigned)(v >> 32) & 1u, (unsigned)(v >> 33) & 0x1FFu }; }

    struct WtSmem {
        alignas(16) unsigned char bytes[kDecTB + 48];  // tile bytes at [shift, shift + kDecTB + 8), zero padded
        unsigned       rec[kWtNodes + 3];              // per node: value relative to its base (ops, E nodes, EXT entries)
        unsigned short base[kDecTB + 32];              // per op / E node: base node id
        alignas(4) unsigned char slot[kDecTB + 4];     // per op: slot | flags; an OP_RUN keeps its tag byte (0xC0 | run - 1)
        unsigned       lastk[64];                      // 1 + last op per slot (0 = none), built with atomicMax
        unsigned       fixe[kFixWords];                // learned alphas: pos | alpha << 16
        unsigned       fails[kFixWords];               // OP_RGB ops refuted in this round: op ordinal | actual alpha << 16
        unsigned char  hE[32];                         // slot of the value entering each lane's chunk
        unsigned       nfail, fixn, fixn0, fix_dirty;
    };

    // store one pixel (target 3 or 4 bytes), optionally bottom-up rows (simple.cpp:401-408 done in place)
    __device__ QB_NOINLINE void store_pixel(uint8_t* out, uint64_t pix, unsigned val, const DecParams& P)
    {
        if (P.flip) {
            const uint64_t y = pix / P.width, x = pix - y * P.width;
            pix = (uint64_t)(P.height - 1 - y) * P.width + x;
        }
        if (P.target == 4) {
            uint8_t* d = out + pix * 4;
            if ((reinterpret_cast<uintptr_t>(out) & 3u) == 0) *reinterpret_cast<unsigned*>(d) = val;
            else d[0] = (uint8_t)val, d[1] = (uint8_t)(val >> 8), d[2] = (uint8_t)(val >> 16), d[3] = (uint8_t)(val >> 24);
        } else {
            uint8_t* d = out + pix * 3;
            d[0] = (uint8_t)val, d[1] = (uint8_t)(val >> 8), d[2] = (uint8_t)(val >> 16);
        }
    }

    __device__ __forceinline__ void locate_image(const DecParams& P, unsigned gt, unsigned& img, unsigned& t, unsigned& ntiles,
                                                 const uint8_t*& stream, uint64_t& size)
    {
        if (P.tile_first == nullptr) {
            img = 0, t = gt, ntiles = P.n_tiles;
            stream = P.qoi + P.single[0], size = P.single[1] - P.single[0];
            return;
        }
        unsigned lo = 0, hi = P.n_images;  // largest img with tile_first[img] <= gt
        while (hi - lo > 1) {
            const unsigned mid = (lo + hi) >> 1;
            if (__ldg(P.tile_first + mid) <= gt) lo = mid;
            else hi = mid;
        }
        img    = lo;
        const unsigned f = __ldg(P.tile_first + lo);
        t = gt - f, ntiles = __ldg(P.tile_first + lo + 1) - f;
        const uint64_t o0 = __ldg(P.offsets + 2u * lo);
        stream = P.qoi + o0, size = __ldg(P.offsets + 2u * lo + 1u) - o0;
    }

    // ---- pixels before tile t of its image and the alpha of the last OP_RGBA before it, WITHOUT a chain: every tile
    // publishes its own count (kDwPixA) right after its parse; the last tile of every 64 (4096) sums its group
    // (super-group, top-level group) from those words (kDwGrp / kDwSup / kDwTop, in its own descriptor).  A reader folds <= 31 tile words, <= 31
    // group words and the super-group words before it: it waits for tiles that started before it to pass their parse,
    // never for a predecessor's own prefix.  (A chained look-back of this sum moved at 32 tiles per L2 round trip: the
    // whole decode was bounded by it, profiles/r02_experiments.md.)  All lanes of one warp call this together.
    template <class Fetch>
    __device__ __forceinline__ PixA wt_fold32(unsigned n, Fetch fetch)  // elements 0 .. n-1 (n <= 32) in order, lane = element
    {
        const unsigned lane = threadIdx.x & 31u;
        const PixA     v    = lane < n ? fetch(lane) : PixA{ 0u, 0u, 0u };
        // the sum in three 16/16/1-bit slices (each slice sum fits 32 bits); the alpha of the last element that has one
        const uint64_t s = (uint64_t)__reduce_add_sync(kFull, v.lo & 0xFFFFu) + ((uint64_t)__reduce_add_sync(kFull, v.lo >> 16) << 16) +
                           ((uint64_t)__reduce_add_sync(kFull, v.hi) << 32);
        const unsigned has = __ballot_sync(kFull, (v.a & 0x100u) != 0);
        const unsigned a   = __shfl_sync(kFull, v.a, has ? 31 - __clz((int)has) : 0);
        const uint64_t c   = s > kPixSat ? kPixSat : s;
        return PixA{ (unsigned)c, (unsigned)(c >> 32), has ? a : 0u };
    }
    __device__ __forceinline__ PixA wt_wait_pixa(const uint64_t* d_t, unsigned t, unsigned p, int which, const Epochs& ep)
    {
        return pixa_unpack(word_payload(wait_word(d_t - (int64_t)(t - p) * kDecDescWords + which, ep, p)));
    }
    // sum over tiles [t0, t0 + n) (n <= 64) of word `which`, stride `step` tiles between elements
    __device__ QB_NOINLINE PixA wt_fold_words(const uint64_t* d_t, unsigned t, unsigned first, unsigned n, unsigned step, int which, const Epochs& ep)
    {
        PixA acc{ 0u, 0u, 0u };
#pragma unroll 1
        for (unsigned o = 0; o < n; o += 32u)
            acc = pixa_comb(acc, wt_fold32(min(n - o, 32u), [&](unsigned i) { return wt_wait_pixa(d_t, t, first + (o + i) * step, which, ep); }));
        return acc;
    }
    __device__ QB_NOINLINE PixA wt_gather_pixa(const uint64_t* d_t, unsigned t, const Epochs& ep, const DecState* init = nullptr)
    {
        // (A variant that issued the tile-word and group-word loads of a lane together, four in flight, decoded wrongly on
        // the B200 although it is equivalent on paper and passes in the emulator -- profiles/r02_experiments.md; the folds
        // stay one after the other.)
        PixA           acc{ 0u, 0u, 0x1FFu };  // before the stream: no pixels, alpha 255
        if (init) acc = PixA{ init->run, 0u, 0x100u | (init->prev >> 24) };  // resumable decode: the pending run comes first (stream.cpp:335-339)
        const unsigned top = t / kTop, s = t / kSup, g = t / kGrp, r = t % kGrp;
#ifndef QB_TIMING
        for (unsigned j = 0; j < top; j += 64u)  // top-level groups before mine (one per 32768 tiles: 29 MB of stream)
            acc = pixa_comb(acc, wt_fold_words(d_t, t, j * kTop + kTop - 1u, min(64u, top - j), kTop, kDwTop, ep));
        const unsigned s0 = top * (kTop / kSup);  // super-groups of my top-level group before mine
#else
        const unsigned s0 = 0;
        (void)top;
#endif
        for (unsigned j = s0; j < s; j += 64u)
            acc = pixa_comb(acc, wt_fold_words(d_t, t, j * kSup + kSup - 1u, min(64u, s - j), kSup, kDwSup, ep));
        const unsigned g0 = s * (kSup / kGrp);  // groups of my super-group before mine
        if (g > g0) acc = pixa_comb(acc, wt_fold_words(d_t, t, g0 * kGrp + kGrp - 1u, g - g0, kGrp, kDwGrp, ep));
        if (r) acc = pixa_comb(acc, wt_fold_words(d_t, t, g * kGrp, r, 1u, kDwPixA, ep));  // tiles of my group before me
        return acc;
    }
    // the last tile of a group / super-group publishes the totals (own count `mine` not yet visible through memory)
    __device__ QB_NOINLINE void wt_publish_groups(uint64_t* d_t, unsigned t, const PixA& mine, unsigned epoch, const Epochs& ep)
    {
        const unsigned lane = threadIdx.x & 31u;
        if (t % kGrp != kGrp - 1u) return;
        const PixA grp = pixa_comb(wt_fold_words(d_t, t, t - (kGrp - 1u), kGrp - 1u, 1u, kDwPixA, ep), mine);
        if (lane == 0) st_word(d_t + kDwGrp, pack_word(pixa_pack(grp), ST_INCL, epoch));
        if (t % kSup != kSup - 1u) return;
        const PixA sup = pixa_comb(wt_fold_words(d_t, t, t - (kSup - kGrp), kSup / kGrp - 1u, kGrp, kDwGrp, ep), grp);
        if (lane == 0) st_word(d_t + kDwSup, pack_word(pixa_pack(sup), ST_INCL, epoch));
#ifndef QB_TIMING
        if (t % kTop != kTop - 1u) return;
        const PixA tp = pixa_comb(wt_fold_words(d_t, t, t - (kTop - kSup), kTop / kSup - 1u, kSup, kDwSup, ep), sup);
        if (lane == 0) st_word(d_t + kDwTop, pack_word(pixa_pack(tp), ST_INCL, epoch));
#endif
    }

    // incoming value of state entry `e` (0..63 table slot, 64 prev) of tile `t`: follow the chain of transfer words
    // through the predecessors until a constant (inclusive word) or the start of the stream.  `d_t` = descriptor of tile t.
    __device__ QB_NOINLINE unsigned wt_resolve_entry(const uint64_t* d_t, unsigned t, unsigned e, const Epochs& ep, const DecState* init = nullptr)
    {
        unsigned acc = 0;
        for (int p = (int)t - 1;; --p) {
            if (p < 0 && init) return add4(e == 64u ? init->prev : init->table[e], acc);  // stream.hpp:239-243
            if (p < 0) return add4((e == 64u || e == 53u) ? kStartPixel : 0u, acc);  // simple.cpp:103-108
            const uint64_t wd = wait_word(d_t - (int64_t)(t - (unsigned)p) * kDecDescWords + kDwState + (int)e, ep, (unsigned)p);
            const uint64_t pl = word_payload(wd);
            if (raw_status(wd) == ST_INCL) return add4((unsigned)pl, acc);
            acc = add4(acc, (unsigned)pl);
            e   = (unsigned)(pl >> 32);
        }
    }

    // alpha learned by an earlier round for the OP_RGB at byte `p` of the tile, if any (cold: retry rounds only)
    __device__ QB_NOINLINE bool wt_fix_lookup(const WtSmem& sm, unsigned p, unsigned& alpha)
    {
        bool hit = false;
#pragma unroll 1
        for (unsigned f = 0; f < sm.fixn0; ++f)
            if ((sm.fixe[f] & 0xFFFFu) == p) alpha = (sm.fixe[f] >> 16) & 255u, hit = true;
        return hit;
    }

    // OP_INDEX ops of a tile (tiles without one never come here): every OP_INDEX finds the last earlier op with its slot
    // (backward search over the packed slot bytes), then pointer jumping over the lanes' entry nodes and the OP_INDEX ops
    // ends every node in ABS or in an entry of the state entering the tile.  One warp.
    __device__ QB_NOINLINE void wt_resolve_index(WtSmem& sm, unsigned idxm, unsigned opbase, unsigned& need_lo, unsigned& need_hi)
    {
        const unsigned lane = threadIdx.x & 31u;
        {
            const unsigned* W = reinterpret_cast<const unsigned*>(sm.slot);
            for (unsigned bits = idxm; bits; bits &= bits - 1u) {
                const unsigned q = opbase + (unsigned)__ffs((int)bits) - 1u, s = sm.slot[q] & 63u;
                const unsigned pat = s * 0x01010101u;
                unsigned       found = kNoOp;
                unsigned       keep  = (1u << (8u * (q & 3u))) - 1u;  // first word: only the ops before q
                for (int wi = (int)(q >> 2); wi >= 0; --wi) {
                    const unsigned wv = W[wi];
                    const unsigned x  = (wv & 0x3F3F3F3Fu) ^ pat;                // bytes 0..0x3F, zero where the slot matches
                    const unsigned nr = ~(wv & (wv << 1)) >> 1;                  // bit 6 of a byte: not an OP_RUN (0xC0 | run - 1)
                    const unsigned m  = (0x40404040u - x) & 0x40404040u & keep & nr;  // no borrow crosses a byte
                    if (m) {
                        found = (unsigned)wi * 4u + ((31u - (unsigned)__clz((int)m)) >> 3);
                        break;
                    }
                    keep = 0xFFFFFFFFu;
                }
                if (found != kNoOp) sm.base[q] = (unsigned short)found;
                else {
                    sm.base[q] = (unsigned short)(kIdExt + s);
                    if (s < 32u) need_lo |= 1u << s;
                    else need_hi |= 1u << (s - 32u);
                }
            }
        }
        __syncwarp();

        // ================= pointer jumping over the entry nodes and the OP_INDEX ops =================
        {
            const unsigned maxJ = __reduce_max_sync(kFull, 1u + (unsigned)__popc(idxm));
            for (;;) {
                bool     changed = false;
                unsigned bits    = idxm;
                for (unsigned i = 0; i < maxJ; ++i) {
                    unsigned x = kNoOp;
                    if (i == 0) x = kIdE + lane;
                    else if (bits) x = opbase + (unsigned)__ffs((int)bits) - 1u, bits &= bits - 1u;
                    unsigned nb = 0, nr = 0;
                    bool     upd = false;
                    if (x != kNoOp) {
                        const unsigned b = sm.base[x];
                        if (b < kIdExt) nb = sm.base[b], nr = add4(sm.rec[x], sm.rec[b]), upd = true;
                    }
                    __syncwarp();  // every lane has read its target before any node changes
                    if (upd) sm.base[x] = (unsigned short)nb, sm.rec[x] = nr, changed = true;
                    __syncwarp();
                }
                if (!__ballot_sync(kFull, changed)) break;
            }
            // entries of the incoming state the nodes end in
            unsigned bits = idxm;
            for (unsigned i = 0; i <= (unsigned)__popc(idxm); ++i) {
                unsigned x = kIdE + lane;
                if (i) x = opbase + (unsigned)__ffs((int)bits) - 1u, bits &= bits - 1u;
                const unsigned b = sm.base[x];
                if (b >= kIdExt && b < kIdExt + 64u) {
                    const unsigned e = b - kIdExt;
                    if (e < 32u) need_lo |= 1u << e;
                    else need_hi |= 1u << (e - 32u);
                }
            }
        }
    }

    // A tile whose verification failed: remember, for every refuted OP_RGB, the alpha it actually saw (exact if everything
    // before it was exact) -- in shared memory for the tile's own repair pass and in the tile's global list for the retry
    // rounds.  Cold.  One warp.
    // `from` = first tile the next round has to decode again (kNoRedo: none, the tile repairs itself)
    constexpr unsigned kNoRedo = 0xFFFFFFFFu;
    __device__ QB_NOINLINE void wt_flag_redo(const DecParams& P, unsigned round, DecResult* res, unsigned from)
    {
        if ((threadIdx.x & 31u) == 0) {
            if (round == 0) atomicOr(&res->bad, 1u), P.control->rounds_needed = 1;
            atomicMax(&res->first_bad[round], 0xFFFFFFFFu - from);
            P.control->any_bad[round] = 1;
        }
    }
    // Table entries (and prev) whose value changed on the way into a tile
    struct Changed {
        unsigned lo, hi, prev;
    };
    // Round 0: a repaired tile found that tile `from` of its image consumed one of its retracted words.  The request is served
    // after the round by the cascade of decode_finish_body: that tile alone is decoded again, then whatever read ITS changed
    // words, ... -- a handful of tiles instead of everything behind it.  `c` = what had changed on entering `from`.
    constexpr unsigned kCascadeFail = 0xFFFFFFFEu;
    __device__ QB_NOINLINE void wt_flag_cascade(const DecParams& P, DecResult* res, unsigned img, unsigned from, const Changed& c)
    {
        if ((threadIdx.x & 31u) == 0) {
            atomicOr(&res->bad, 2u);
            const unsigned i = atomicAdd(&P.control->n_req, 1u);
            if (i < P.req_cap) P.req[i] = CascadeReq{ img, from, c.lo, c.hi, c.prev, 0u };
            else atomicMax(&res->first_bad[0], 0xFFFFFFFFu - from), P.control->rounds_needed = 1;  // no room: the retry round, from there on
            P.control->any_bad[0] = 1;
        }
    }
    __device__ QB_NOINLINE void wt_record_failures(const DecParams& P, WtSmem& sm, uint32_t* fix, const unsigned char* B, unsigned first_pos,
                                                   unsigned opbase, unsigned nops)
    {
        const unsigned lane = threadIdx.x & 31u;
        __syncwarp();
        const unsigned nf = min(sm.nfail, (unsigned)kFixMax);
        for (unsigned f = 0; f < nf; ++f) {
            const unsigned k = sm.fails[f] & 0xFFFFu, actual = sm.fails[f] >> 16;
            if (k >= opbase && k < opbase + nops) {  // the owning lane walks to the op's byte position
                unsigned p = first_pos;
                for (unsigned i = opbase; i < k; ++i) p += op_length(B[p]);
                unsigned j = 0;
                for (; j < sm.fixn; ++j)
                    if ((sm.fixe[j] & 0xFFFFu) == p) break;
                if (j < (unsigned)kFixMax) {
                    sm.fixe[j] = p | actual << 16;
                    if (j == sm.fixn) sm.fixn = j + 1u;
                    sm.fix_dirty = 1;
                }
            }
            __syncwarp();
        }
        if (sm.fix_dirty && lane < (unsigned)kFixWords) {
            const unsigned n = min(sm.fixn, (unsigned)kFixMax);
            fix[lane] = lane == 0 ? (n | (P.epoch & 0xFFFFFFu) << 8) : (lane <= n ? sm.fixe[lane - 1] : 0u);
        }
    }

    // A tile repaired itself and some of its carry words changed.  Follow the changed state entries through the tiles behind
    // it: a tile that read a changed entry (its kDwNeed words) must be decoded again -- flag the retry round from there; an
    // entry a tile overwrites with a constant, or derives from an unchanged entry (the `ref` its state words keep), stops being
    // changed.  Usually nothing read the one or two table slots in question and they are overwritten within a tile or two.
    // A changed pixel / alpha or slot word is not followed (everything behind the tile is flagged).  Cold.  One warp.
    // Returns the first tile behind t that has to be decoded again (kNoRedo: none; kCascadeFail | tile: the scan could not
    // tell -- everything from that tile on).  `c`: in = entries that had already changed on entering t (a cascade carries them
    // along; empty after a plain repair), out = the entries that have changed on entering the returned tile.
    constexpr unsigned kScanGaveUp = 0x80000000u;
    __device__ QB_NOINLINE unsigned wt_repair_scan(const DecParams& P, uint64_t* desc, unsigned t, unsigned ntiles, const Epochs& ep,
                                                   const uint64_t (&old_word)[3], unsigned dirty, Changed& c)
    {
        constexpr unsigned kScan = 16;  // tiles followed before giving up
        const unsigned     lane  = threadIdx.x & 31u;
        unsigned dlo = 0, dhi = 0, dprev = 0, hard = 0;
        QB_FENCE_SC();  // this tile's republished words are ordered before the reads of the successors' need words (see wt_decode_tile)
        // what tile u hands on of a changed set: an entry is still changed if it is derived from a changed one
        auto hand_on = [&](const uint64_t* d_u, unsigned u, unsigned& lo, unsigned& hi, unsigned& pv) -> bool {
            bool d[3];
            bool missing = false;
#pragma unroll
            for (int hh = 0; hh < 3; ++hh) {
                const unsigned e = lane + 32u * hh;
                d[hh]            = false;
                if (e <= 64u) {
                    const uint64_t ws = ld_word(d_u + kDwState + e);
                    if (!ep.valid(ws, u)) missing = true;
                    const unsigned ref = (unsigned)(word_payload(ws) >> 32) & 0x7Fu;
                    d[hh] = ref < 32u ? (lo >> ref) & 1u : (ref < 64u ? (hi >> (ref - 32u)) & 1u : (ref == 64u ? pv != 0 : false));
                }
            }
            if (__ballot_sync(kFull, missing)) return false;
            lo = __ballot_sync(kFull, d[0]), hi = __ballot_sync(kFull, d[1]), pv = __ballot_sync(kFull, d[2]) & 1u;
            return true;
        };
        // (1) what this tile says differently from before
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const unsigned wi = 1u + 32u * i + lane;
            if (wi <= (unsigned)kDwState + 64u && (ld_word(desc + wi) != old_word[i] || ((dirty >> i) & 1u))) {
                if (wi < (unsigned)kDwState) hard = 1;
                else {
                    const unsigned e = wi - kDwState;
                    if (e < 32u) dlo |= 1u << e;
                    else if (e < 64u) dhi |= 1u << (e - 32u);
                    else dprev = 1;
                }
            }
        }
        dlo = __reduce_or_sync(kFull, dlo), dhi = __reduce_or_sync(kFull, dhi), dprev = __reduce_or_sync(kFull, dprev), hard = __reduce_or_sync(kFull, hard);
        // (2) what had changed before it and passes through it
        if (c.lo | c.hi | c.prev) {
            unsigned lo = c.lo, hi = c.hi, pv = c.prev;
            if (!hand_on(desc, t, lo, hi, pv)) return kScanGaveUp | t;
            dlo |= lo, dhi |= hi, dprev |= pv;
        }
        QB_COUNT(P, 0);  // repaired tiles
        c = Changed{ dlo, dhi, dprev };
        if (!(dlo | dhi | dprev | hard)) return kNoRedo;  // the repaired tile says exactly what it said before
        if (hard || dprev) QB_COUNT(P, 1);
        unsigned u = t + 1u;
#ifdef QB_REPAIR_ALWAYS_REDO
        return u < ntiles ? (kScanGaveUp | u) : kNoRedo;
#endif
        if (hard) return u < ntiles ? u : kNoRedo;  // a pixel / alpha or slot word changed: the next tile used it for certain
        for (; u < ntiles && u <= t + kScan; ++u) {
            const uint64_t* d_u = desc + (uint64_t)(u - t) * kDecDescWords;
            // never wait here: tile u may not even have been drawn yet, and the warp that would draw it may be this one's
            // neighbour stuck in the same scan.  A tile that has not said what it reads is decoded again, with all behind it.
            const uint64_t wl = ld_word(d_u + kDwNeedLo), wh = ld_word(d_u + kDwNeedHi);
            if (!ep.valid(wl, u) || !ep.valid(wh, u)) return kScanGaveUp | u;
            const uint64_t nl = word_payload(wl), nh = word_payload(wh);
            if (((unsigned)nl & dlo) | ((unsigned)nh & dhi) | dprev) {  // tile u read a changed entry (prev is always read)
                QB_COUNT(P, 1);
                c = Changed{ dlo, dhi, dprev };
                return u;
            }
            if (!hand_on(d_u, u, dlo, dhi, dprev)) return kScanGaveUp | u;
            if (!(dlo | dhi | dprev)) return kNoRedo;  // every changed entry has been overwritten before anything read it
        }
        if (u >= ntiles) return kNoRedo;  // the stream ended first
        return kScanGaveUp | u;
    }

    // one tile (global ticket `gticket`) of round `round`; one warp.
    // kStream = the resumable decode (StreamDecoder::decode, stream.cpp:312-424): the buffer has no header, the state the
    // stream is entered with is P.init instead of the constants, a pending run comes first, P.n_pixels is the room in the
    // output, an op that the input does not hold completely is not consumed (stream.cpp:341-392), and the tile that holds
    // the last consumed op ("final tile") reports bytes consumed, bytes written and the state to carry on with.  Tiles behind
    // it do nothing.  A refuted speculation is not retried here: the call falls back to the sequential loop.
    // `cascade`: the tile is decoded AGAIN after round 0 (decode_finish_body) because a predecessor retracted a word it had
    // consumed; everything around it is at rest.  Its carry words as they were count as "published before": what differs
    // afterwards is followed through its successors like after a repair.  Returns the next tile to decode again (kNoRedo: none,
    // kCascadeFail: this tile is still refuted -- the retry rounds take over); kNoRedo always outside a cascade.
    template <bool kStream = false>
    __device__ __forceinline__ unsigned wt_decode_tile(const DecParams& P, WtSmem& sm, const uint2* lut, unsigned round, unsigned gticket, unsigned img,
                                                       unsigned t, unsigned ntiles, const uint8_t* stream, uint64_t size, unsigned fresh_from,
                                                       Changed* casc = nullptr)
    {
        const bool cascade = casc != nullptr;
        const unsigned lane = threadIdx.x & 31u;
        [[maybe_unused]] const long long qb_t0 = QB_T0();
        constexpr unsigned kSkip   = kStream ? 0u : kHeader;
        const DecState*    init    = kStream ? P.init : nullptr;
        const uint64_t body_len = size - kSkip;  // every byte after the header is chunk data (simple.cpp:110-113)
        const uint64_t tile_b0  = (uint64_t)t * kDecTB;
        const unsigned limit    = (unsigned)(body_len - tile_b0 < (uint64_t)kDecTB ? body_len - tile_b0 : (uint64_t)kDecTB);
        uint64_t*      desc     = P.desc + (uint64_t)gticket * kDecDescWords;
        const unsigned epoch    = P.epoch + round;
        const Epochs   ep{ epoch, P.epoch, fresh_from };
        uint8_t*       out      = P.out + (uint64_t)img * P.out_stride;
        const uint64_t N        = P.n_pixels;
        DecResult*     res      = P.results + img;
        uint32_t*      fix      = P.fix + (uint64_t)gticket * kFixWords;
        auto word_of = [&](unsigned p, int which) { return desc - (int64_t)(t - p) * kDecDescWords + which; };

        // ---- alpha values learned by earlier rounds for OP_RGB ops of this tile
        if (lane < (unsigned)kFixWords) {
            unsigned n = 0;
            if (round > 0) {
                const unsigned h = fix[0];
                if ((h >> 8) == (P.epoch & 0xFFFFFFu)) n = min(h & 255u, (unsigned)kFixMax);
                if (lane >= 1 && lane <= n) sm.fixe[lane - 1] = fix[lane];
            }
            if (lane == 0) sm.fixn = n, sm.fixn0 = n, sm.fix_dirty = 0, sm.nfail = 0;
        }

        // ---- stage the tile: 16-byte aligned vectors land at the same misalignment in shared memory
        const uint8_t* src   = stream + kSkip + tile_b0;
        const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        {
            const uint64_t avail = body_len - tile_b0;  // bytes of the stream from the tile start
            const unsigned want  = (unsigned)(avail < (uint64_t)(kDecTB + 8) ? avail : (uint64_t)(kDecTB + 8));
            const unsigned nvec  = (shift + want + 15u) >> 4;
            const uint4*   vsrc  = reinterpret_cast<const uint4*>(src - shift);
            for (unsigned c = lane; c < nvec; c += 32u) reinterpret_cast<uint4*>(sm.bytes)[c] = __ldg(vsrc + c);
            __syncwarp();
            for (unsigned b = shift + want + lane; b < (unsigned)kDecTB + 48u; b += 32u) sm.bytes[b] = 0;  // zero padding, simple.cpp:106
            __syncwarp();
        }
        const unsigned char* B     = sm.bytes + shift;
        const unsigned       cbeg  = lane * kWtChunk;
        const unsigned       cend  = min(cbeg + (unsigned)kWtChunk, limit);  // ops of this lane start below cend
        QB_STAMP(desc, 72, 0, qb_t0);  // ticket + staging

        // ================= look-back 1: parse map; counts along the path from entry offset 0 =================
        unsigned M0 = 0, R0 = 0, X0 = 0, A0 = kNoPos;  // op starts / OP_RUN ops (bit = byte of the chunk), extra run pixels, last OP_RGBA
        Map      mymap;
        const unsigned* lut_y = reinterpret_cast<const unsigned*>(lut) + 1;  // word y of entry i at lut_y[2 i]
        {
            unsigned p = cbeg;
            while (p < cend) {
                const unsigned tag = B[p], y = lut_y[2u * tag], bit = 1u << (p - cbeg);
                const unsigned rx  = (y >> 17) & 63u;
                M0 |= bit;
                X0 += rx;
                if (rx) R0 |= bit;
                if (tag == kOpRgba) A0 = p;
                p += (y >> 14) & 7u;
            }
            const unsigned cfull = cbeg + kWtChunk;
            const unsigned exit0 = p > cfull ? p - cfull : 0u;
            unsigned       win   = exit0;
#ifdef QB_WT_ALL_ENTRIES
#pragma unroll 1
            for (unsigned e = 1; e <= 4u; ++e) {  // a late entry usually falls into the path of entry 0 after an op or two
                unsigned q = cbeg + e;
                while (q < cend && !((M0 >> (q - cbeg)) & 1u)) q += (lut_y[2u * B[q]] >> 14) & 7u;
                win |= (q < cend ? exit0 : (q > cfull ? q - cfull : 0u)) << (3u * e);
            }
#else
            // Only the entry offsets a chunk can really be entered at are walked: lane 0 can be entered anywhere (the tile's
            // entry is not known yet), lane l only at an exit of lane l - 1 over the entries IT can be entered at -- nearly
            // always the single exit of its path 0.  Entries that cannot occur keep a meaningless exit in the map: composition
            // never selects them.  The sets grow until nothing new appears (two rounds as a rule; walking all four late
            // entries of every lane was the most expensive line of the kernel, 4.8 % of its instructions).
            unsigned have = 1u, outs = 1u << exit0;  // entries walked / exits they produced, as sets
            auto walk_entry = [&](unsigned e) {
                unsigned q = cbeg + e;
                while (q < cend && !((M0 >> (q - cbeg)) & 1u)) q += (lut_y[2u * B[q]] >> 14) & 7u;
                const unsigned ex = q < cend ? exit0 : (q > cfull ? q - cfull : 0u);
                win |= ex << (3u * e), have |= 1u << e, outs |= 1u << ex;
            };
            if (lane == 0) {
#pragma unroll 1
                for (unsigned e = 1; e <= 4u; ++e) walk_entry(e);
            }
            for (;;) {
                unsigned need = __shfl_up_sync(kFull, outs, 1);
                if (lane == 0) need = 0;
                need &= ~have;
                if (__ballot_sync(kFull, need != 0) == 0) break;
#pragma unroll 1
                while (need) {
                    const unsigned e = (unsigned)__ffs((int)need) - 1u;
                    need &= need - 1u;
                    walk_entry(e);
                }
            }
#endif
            mymap = map_unpack(win);
        }
        Map incl_map = mymap;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Map o = Map{ __shfl_up_sync(kFull, incl_map.lo, d), __shfl_up_sync(kFull, incl_map.hi, d) };
            if ((int)lane >= d) incl_map = map_compose(o, incl_map);
        }
        Map excl_map = Map{ __shfl_up_sync(kFull, incl_map.lo, 1), __shfl_up_sync(kFull, incl_map.hi, 1) };
        if (lane == 0) excl_map = map_identity();
        const Map tile_map = Map{ __shfl_sync(kFull, incl_map.lo, 31), __shfl_sync(kFull, incl_map.hi, 31) };
        unsigned  tile_entry;
        {
            // a tile that re-synchronises (nearly all do) exits at the same offset whatever it was entered at: inclusive at once
            const bool const_map = tile_map.lo == (tile_map.lo & 0xFFu) * 0x01010101u && tile_map.hi == (tile_map.lo & 0xFFu);
            if (lane == 0 && t > 0) st_word(desc + kDwParse, pack_word(map_pack(tile_map), const_map ? ST_INCL : ST_AGG, epoch));
            const Map in = warp_lookback_lazy<Map>(
                t, map_const(0), map_identity(),
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(word_of(p, kDwParse));
                    st                = ep.valid(wd, p) ? raw_status(wd) : (unsigned)ST_NONE;
                    return map_unpack((unsigned)word_payload(wd));
                },
                [](const Map& a, const Map& b) { return map_compose(a, b); });
            tile_entry = in.lo & 7u;  // `in` is constant: every chain ended in an inclusive word
            if (lane == 0 && (t == 0 || !const_map)) st_word(desc + kDwParse, pack_word(map_pack(map_const(map_at(tile_map, tile_entry))), ST_INCL, epoch));
        }
        const unsigned my_entry = map_at(excl_map, tile_entry);
        QB_STAMP(desc, 72, 1, qb_t0);  // parse + look-back 1

        // ---- everything from here to the emit pass may run again ("repair"): a tile whose OP_RGB alpha speculation is
        // refuted learns the alpha it saw and decodes itself again at once.  Its carry words were published already; if the
        // repaired tile publishes the SAME words (the wrong alpha usually dies inside the tile: the next OP_RGBA resets it and
        // the table slots it touched are overwritten), nothing downstream changed and no retry round is needed.  One refuted
        // op in 18 000 tiles (RGBA photo with soft alpha blobs) used to cost a second decode of everything behind it.
#ifndef QB_REPAIR_PASSES
#define QB_REPAIR_PASSES 3
#endif
        constexpr unsigned kRepairPasses = QB_REPAIR_PASSES;
        uint64_t           old_word[3] = { 0, 0, 0 };  // this lane's share of words 1 .. 67 as published by the latest refuted pass
        unsigned           dirty = 0;                  // bit i: word i of this lane differed between two refuted passes
        if (cascade) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (1u + 32u * i + lane <= (unsigned)kDwState + 64u) old_word[i] = ld_word(desc + 1 + 32 * i + lane);
            if (lane == 0) fix[0] = 0;  // alphas learned under the retracted inputs are not trusted
        }
        uint64_t           pix_base_f = 0;
        unsigned           n_pix_f = 0, fixn0 = sm.fixn0;
        bool               tail_fill_f = false, still_bad = false;
        [[maybe_unused]] unsigned fin_val[3] = { 0, 0, 0 }, fin_rem = 0;
        [[maybe_unused]] uint64_t fin_made = 0, fin_used = 0;
        [[maybe_unused]] bool     fin_final = false;
        unsigned           pass = 0;
        for (;; ++pass) {
        // ================= counts of the true path, look-back 2: pixels and inherited alpha =================
        unsigned nops, npx_lane, a_sum;  // a_sum: 0x100 | alpha after this lane's last alpha setter, 0 = none
        {
            unsigned p = cbeg + my_entry, own = 0, ownx = 0, Apos = kNoPos, X;
            while (p < cend && !((M0 >> (p - cbeg)) & 1u)) {  // until the path of entry 0 is met
                const unsigned tag = B[p], y = lut_y[2u * tag];
                ++own;
                ownx += (y >> 17) & 63u;
                if (tag == kOpRgba) Apos = p;
                p += (y >> 14) & 7u;
            }
            if (p < cend) {
                const unsigned rel = p - cbeg;
                nops               = own + (unsigned)__popc(M0 >> rel);
                unsigned xb        = 0;
                for (unsigned bits = R0 & ((1u << rel) - 1u); bits; bits &= bits - 1u) xb += B[cbeg + (unsigned)__ffs((int)bits) - 1u] & 63u;
                X = ownx + X0 - xb;
                if (A0 != kNoPos && A0 >= p) Apos = A0;
            } else {
                nops = own, X = ownx;
            }
            npx_lane = nops + X;
            a_sum    = Apos != kNoPos ? 0x100u | B[Apos + 4u] : 0u;
            if (fixn0) {  // a learned alpha acts like an OP_RGBA from its op on
#pragma unroll 1
                for (unsigned j = 0; j < fixn0; ++j) {
                    const unsigned fp = sm.fixe[j] & 0xFFFFu;
                    if (fp >= cbeg + my_entry && fp < cend && (Apos == kNoPos || fp > Apos)) Apos = fp, a_sum = 0x100u | ((sm.fixe[j] >> 16) & 255u);
                }
            }
        }
        unsigned opbase, pix_lane, n_ops, n_pix, alpha_lane;  // alpha_lane: 0x100 | alpha when an earlier lane of this tile sets it, else 0
        {
            const unsigned mine = npx_lane | nops << 16;
            unsigned       inc  = mine, ai = a_sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(kFull, inc, d), oa = __shfl_up_sync(kFull, ai, d);
                if ((int)lane >= d) {
                    inc += o;
                    if (!(ai & 0x100u)) ai = oa;
                }
            }
            const unsigned tot = __shfl_sync(kFull, inc, 31), atot = __shfl_sync(kFull, ai, 31);
            unsigned       aex = __shfl_up_sync(kFull, ai, 1);
            if (lane == 0) aex = 0;
            opbase = (inc - mine) >> 16, pix_lane = (inc - mine) & 0xFFFFu;
            n_ops = tot >> 16, n_pix = tot & 0xFFFFu;
            const PixA agg{ n_pix, 0u, atot };
            if (lane == 0) st_word(desc + kDwPixA, pack_word(pixa_pack(agg), ST_INCL, epoch));
            if (t % kGrp == kGrp - 1u) wt_publish_groups(desc, t, agg, epoch, ep);
            alpha_lane = aex;
        }
        QB_STAMP(desc, 73, 0, qb_t0);  // counts

        // ================= the walk: one record per op =================
        // Branch-free for the common ops: the tag's table entry gives the delta (OP_DIFF / OP_LUMA), the slot contribution, the
        // length and the flags; a literal overrides by selects; only OP_INDEX and learned alphas branch.  r and b accumulate in
        // two 16-bit lanes of one register (carries stay inside a lane for the <= 28 ops of a chunk), g in another.
        unsigned idxm = 0;  // OP_INDEX ops of this lane (bit = op number within the lane)
        unsigned exit_bid, exit_h, exit_rec, exit_p;
        {
            const unsigned* S32 = reinterpret_cast<const unsigned*>(sm.bytes);
            unsigned        p = cbeg + my_entry, k = opbase, j = 0;
            unsigned        arb = 0, ag = 0, atop = 0, bid = kIdE + lane, h = 0, alpha = alpha_lane & 255u, recv = 0;
            unsigned        lit_bid = (alpha_lane & 0x100u) ? kIdAbs : kIdAbsA;  // base of a literal: its alpha is known / the tile's entry alpha
            while (p < cend) {
                const unsigned a   = shift + p;
                const unsigned x   = __funnelshift_r(S32[a >> 2], S32[(a >> 2) + 1u], (a & 3u) * 8u);  // the op's first four bytes
                const unsigned tag = x & 0xFFu;
                const uint2    L   = lut[tag];
                const unsigned lm  = 0u - ((L.y >> 23) & 1u);                                             // OP_LUMA: all ones
                const unsigned nib = (((x >> 12) & 15u) | ((x >> 8) & 15u) << 16) & lm;                   // dr - dg + 8, db - dg + 8
                arb += L.x + nib, ag += L.y;                                                              // bits above 7 of ag are ignored
                h += (L.y >> 8) + __dp2a_lo(nib, 0x00000703u, 0u);
                if (tag >= kOpRgb) {  // simple.cpp:119-129; OP_RGB keeps the alpha: speculated here, verified in the emit pass
                    if (tag == kOpRgba) alpha = sm.bytes[a + 4u], lit_bid = kIdAbs;
                    else if (fixn0 && wt_fix_lookup(sm, p, alpha)) lit_bid = kIdAbs;
                    const unsigned rgb = x >> 8;
                    arb = rgb & 0x00FF00FFu, ag = rgb >> 8, atop = alpha << 24, bid = lit_bid;  // alpha = 0 while it is provisional
                    h = __dp4a(rgb | atop, 0x0B070503u, 0u);
                } else if (tag < 0x40u) {  // simple.cpp:132-135: the value is found by the search below
                    arb = 0, ag = 0, atop = 0, bid = k, h = tag, idxm |= 1u << j;
                }
                recv = (__byte_perm(arb, ag, 0x3240) & 0x00FFFFFFu) | atop;
                const unsigned fl = L.y >> 24;
                sm.rec[k]  = recv;
                sm.base[k] = (unsigned short)bid;
                sm.slot[k] = (unsigned char)(fl == (kSlRgb | kSlIdx) ? tag : ((h & 63u) | fl));  // an OP_RUN keeps its tag
                ++k, ++j, p += (L.y >> 14) & 7u;
            }
            exit_bid = bid, exit_h = h & 63u, exit_rec = recv, exit_p = p;
        }
        QB_STAMP(desc, 77, 1, qb_t0);  // walk loop
        sm.lastk[lane] = 0, sm.lastk[lane + 32] = 0;
        const bool any_idx = __ballot_sync(kFull, idxm != 0) != 0;
        // ---- look-back 2 (no chain, see wt_gather_pixa): pixels before the tile and the alpha entering it.  Taken AFTER the
        // walk: by now the tiles before this one have long published their counts (before the walk, every tile waited here for
        // the slowest of all its predecessors to finish parsing: a third of the tile time).
        const PixA     pin      = wt_gather_pixa(desc, t, ep, init);
        const uint64_t pix_base = (uint64_t)pin.hi << 32 | pin.lo;
        const unsigned ain      = pin.a & 255u;
        if (lane == 0) sm.rec[kIdAbsA] = ain << 24;
        // ops of this tile that are executed: all of them, except in the final tile of a resumable decode.  (A second variable
        // beside n_ops in the one-shot instantiation cost 18 % of the 8K decode: registers across the whole tile body.)
        [[maybe_unused]] unsigned n_keep_s = n_ops;
        auto n_keep = [&]() -> unsigned {
            if constexpr (kStream) return n_keep_s;
            else return n_ops;
        };
        // ops of this tile that are executed (all of them, except in the final tile of a resumable decode)
        [[maybe_unused]] bool     final_tile = false;
        [[maybe_unused]] unsigned run_rem = 0, final_endp = 0;
        [[maybe_unused]] uint64_t final_made = 0, final_used = 0;  // pixels in the output / input bytes consumed when this is the final tile
        if constexpr (kStream) {
            (void)pix_lane, (void)exit_p;
            if (t == 0) {  // the pending run comes first (stream.cpp:335-339)
                const uint64_t pre = (uint64_t)init->run < N ? (uint64_t)init->run : N;
                for (uint64_t q = lane; q < pre; q += 32u) store_pixel(out, q, init->prev, P);
            }
            // Only the last op of the whole input can be incomplete; it is left for the next call.  It may START in the tile before
            // the last one and reach past the end of the input through the last tile, which then has nothing to decode.
            const uint64_t avail = body_len - tile_b0;
            if ((uint64_t)tile_entry > avail) return kNoRedo;
            const bool incomplete = __ballot_sync(kFull, nops != 0 && (uint64_t)exit_p > avail) != 0;
            const unsigned keep_ops = n_ops - (incomplete ? 1u : 0u), keep_pix = n_pix - (incomplete ? 1u : 0u);
            if (pix_base >= N) {
                if (t > 0) return kNoRedo;  // the output was full before this tile: nothing of it is consumed
                final_tile = true, n_keep_s = 0, run_rem = (unsigned)(pix_base - N), final_made = N, final_used = 0;  // not even the pending run fits
            } else if (pix_base + keep_pix >= N || incomplete || t == ntiles - 1u) {
                final_tile = true;
                // the op that holds the last pixel that fits (the output is the limit), else the last complete op (the input is)
                const bool     by_room = pix_base + keep_pix >= N;
                const unsigned target  = by_room ? (unsigned)(N - 1u - pix_base) : 0u;
                unsigned       kcut = 0, rem = 0, endp = 0;
                bool           hit = false;
                {
                    unsigned p = cbeg + my_entry, po = pix_lane;
                    for (unsigned j = 0, k = opbase; j < nops && k < keep_ops; ++j, ++k) {
                        const unsigned tag = B[p], len = op_length(tag), ext = (tag >= kOpRun && tag < kOpRgb) ? (tag & 63u) : 0u;
                        if (by_room ? (po <= target && target <= po + ext) : (k + 1u == keep_ops)) hit = true, kcut = k, rem = by_room ? po + ext - target : 0u, endp = p + len;
                        po += 1u + ext, p += len;
                    }
                }
                const unsigned who = __ballot_sync(kFull, hit);
                if (who) {
                    const int src = __ffs((int)who) - 1;
                    n_keep_s = __shfl_sync(kFull, kcut, src) + 1u, run_rem = __shfl_sync(kFull, rem, src), final_endp = __shfl_sync(kFull, endp, src);
                } else {
                    n_keep_s = 0, final_endp = tile_entry;  // no complete op starts in this tile
                }
                final_made = by_room ? N : pix_base + keep_pix, final_used = tile_b0 + final_endp;
            }
        }
        QB_STAMP(desc, 76, 0, qb_t0);  // walk + gather
        if (exit_bid == kIdAbsA) exit_h = (exit_h + 11u * ain) & 63u;  // util.hpp:347-351: the alpha's share of the slot
        // `prev` leaving the tile is known already when the last op follows a literal: inclusive at once, the next tile waits less
        bool prev_done = false;
        {
            const unsigned last_lane = 31u - (unsigned)__clz((int)(__ballot_sync(kFull, nops != 0) | 1u));
            const unsigned lb = __shfl_sync(kFull, exit_bid, (int)last_lane), lr = __shfl_sync(kFull, exit_rec, (int)last_lane);
            prev_done = n_ops != 0 && lb == kIdAbs && n_keep() == n_ops;
            if (prev_done && lane == 0) st_word(desc + kDwState + 64, pack_word((uint64_t)65u << 32 | lr, ST_INCL, epoch));
        }
        // entry nodes: the value entering lane l's chunk = the value of the op before its first one
        if (any_idx) {  // general form: an alias of that op, resolved by the pointer jumping below
            sm.base[kIdE + lane] = (unsigned short)(opbase ? opbase - 1u : kIdExt + 64u);
            sm.rec[kIdE + lane]  = 0;
        } else {  // no OP_INDEX in the tile: a lane leaves either an absolute value or (its entry + delta): one warp scan
            unsigned va = exit_bid == kIdAbs ? 1u : (exit_bid == kIdAbsA ? 2u : 0u), vv = nops ? exit_rec : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned oa = __shfl_up_sync(kFull, va, d), ov = __shfl_up_sync(kFull, vv, d);
                if ((int)lane >= d && !va) va = oa, vv = add4(ov, vv);
            }
            unsigned ea = __shfl_up_sync(kFull, va, 1), ev = __shfl_up_sync(kFull, vv, 1);
            if (lane == 0) ea = 0, ev = 0;
            sm.base[kIdE + lane] = (unsigned short)(ea == 1u ? kIdAbs : (ea == 2u ? kIdAbsA : kIdExt + 64u));
            sm.rec[kIdE + lane]  = ev;
        }
        QB_STAMP(desc, 73, 1, qb_t0);  // walk

        // ================= look-back 3: slot of the value entering the tile / every lane =================
        {
            const bool     rooted = exit_bid != kIdE + lane;  // a literal or an OP_INDEX made the slot absolute
            const unsigned mine   = exit_h | (rooted ? 64u : 0u);
            auto comb = [](unsigned a, unsigned b) { return (b & 64u) ? b : ((a & 64u) | ((a + b) & 63u)); };  // a before b
            unsigned inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(kFull, inc, d);
                if ((int)lane >= d) inc = comb(o, inc);
            }
            const unsigned tot = __shfl_sync(kFull, inc, 31);
            unsigned       ex  = __shfl_up_sync(kFull, inc, 1);
            if (lane == 0) ex = 0;
            if (lane == 0 && t > 0) st_word(desc + kDwSlot, pack_word(tot, (tot & 64u) ? ST_INCL : ST_AGG, epoch));
            const unsigned in = warp_lookback_lazy<unsigned>(
                t, kStream ? 64u | slot_of(init->prev) : 64u | 53u, 0u,  // {0,0,0,255}: slot 53 (simple.cpp:108); resumable: the carried pixel
                [&](unsigned p, unsigned& st) {
                    if (p < ep.fresh_from) {  // a tile finished by an earlier round: take the slot of its actual last pixel
                        const uint64_t wd = ld_word(word_of(p, kDwState + 64));
                        st                = (ep.valid(wd, p) && raw_status(wd) == ST_INCL) ? (unsigned)ST_INCL : (unsigned)ST_NONE;
                        return 64u | slot_of((unsigned)word_payload(wd));
                    }
                    const uint64_t wd = ld_word(word_of(p, kDwSlot));
                    st                = ep.valid(wd, p) ? raw_status(wd) : (unsigned)ST_NONE;
                    return (unsigned)word_payload(wd);
                },
                comb);
            if (lane == 0 && (t == 0 || !(tot & 64u))) st_word(desc + kDwSlot, pack_word(comb(in, tot), ST_INCL, epoch));
            sm.hE[lane] = (unsigned char)((ex & 64u) ? ex & 63u : (in + ex) & 63u);
        }
        __syncwarp();

        QB_STAMP(desc, 76, 1, qb_t0);  // entry nodes + look-back 3
        // ================= absolute slots, last op per slot =================
        // lastk[s] = 1 + the last op of the tile whose value is stored in slot s (0 = none).
        // Lane-serial: every lane runs over its own ops (the ones it recorded in the walk), so the lane's entry slot is a
        // register and "before the lane's first root" needs no search; the last writer of a slot is an atomicMax in shared
        // memory.  (Lane = op, this phase cost 32 ops per ~30 instructions with a store / read-back / retry loop for equal slots;
        // here it is 32 ops per ~10.)
        {
            const unsigned hEl = sm.hE[lane], ownE = kIdE + lane, a11 = 11u * ain;
            for (unsigned j = 0, k = opbase; j < nops && (!kStream || k < n_keep()); ++j, ++k) {
                const unsigned s8 = sm.slot[k];
                if (s8 >= 0xC0u) continue;  // an OP_RUN repeats its predecessor: never the only writer of a slot
                const unsigned b = sm.base[k];
                unsigned       s = s8 & 63u;
                if (b == ownE || b == kIdAbsA) {  // the walk knew the slot relative to the lane's entry / to the tile's entry alpha
                    s = (s + (b == ownE ? hEl : a11)) & 63u;
                    if (any_idx) sm.slot[k] = (unsigned char)(s | (s8 & 0xC0u));  // the searches below compare absolute slots
                }
                atomicMax(&sm.lastk[s], k + 1u);
            }
        }
        __syncwarp();
        QB_STAMP(desc, 74, 0, qb_t0);  // slots + last writers

        // ================= OP_INDEX ops: writers (backward search over slot[]) and pointer jumping =================
        unsigned need_lo = 0, need_hi = 0;  // table entries of the state entering the tile that are read
        if (any_idx) wt_resolve_index(sm, idxm, opbase, need_lo, need_hi);
        QB_STAMP(desc, 74, 1, qb_t0);  // index search + jumping

        // ================= the tile's transfer function: entries it can state now =================
        // entry e = lane, lane + 32 (table slots) and, in lane 0, 64 (prev)
        unsigned pub_ref[3], pub_add[3];  // ref: 0..64 = incoming entry, 65 = constant (published inclusive already)
        const bool tail_fill = !kStream && t == ntiles - 1 && pix_base + n_pix < N;  // the stream ends before the image does
#pragma unroll
        for (int hh = 0; hh < 3; ++hh) {
            const unsigned e = lane + 32u * hh;
            pub_ref[hh] = 66u, pub_add[hh] = 0;
            if (e > 64u) continue;
            const unsigned k = e < 64u ? (unsigned)sm.lastk[e] - 1u : n_keep() - 1u;  // 0xFFFFFFFF = none
            const bool     none = e < 64u ? sm.lastk[e] == 0 : n_keep() == 0;
            unsigned       ref = e, add = 0;
            if (!none) {
                unsigned b = sm.base[k];
                add        = sm.rec[k];
                if (b < kIdExt) add = add4(add, sm.rec[b]), b = sm.base[b];  // one hop: entry nodes and OP_INDEX ops are final
                if (b == kIdAbsA) add |= ain << 24, b = kIdAbs;
                ref = b == kIdAbs ? 65u : b - kIdExt;
            }
            pub_ref[hh] = ref, pub_add[hh] = add;
            if (ref == 65u) st_word(desc + kDwState + e, pack_word((uint64_t)65u << 32 | add, ST_INCL, epoch));
            else st_word(desc + kDwState + e, pack_word((uint64_t)ref << 32 | add, ST_AGG, epoch));
            if (!none && ref < 64u) {  // an entry this tile writes from an incoming one: resolve it, publish it inclusive below
                if (ref < 32u) need_lo |= 1u << ref;
                else need_hi |= 1u << (ref - 32u);
            }
        }
        if (tail_fill) need_lo |= 1u;  // the zero padding decodes as OP_INDEX 0
        if (kStream && final_tile) need_lo = need_hi = 0xFFFFFFFFu;  // the state to carry on with is all 65 entries
        QB_STAMP(desc, 77, 0, qb_t0);  // transfer function

        // ================= look-back 4: the entries of the incoming state that are read =================
        need_lo = __reduce_or_sync(kFull, need_lo), need_hi = __reduce_or_sync(kFull, need_hi);
        if (lane == 0) {  // what this tile reads of its predecessors' state: lets a repaired predecessor tell whether it mattered
            st_word(desc + kDwNeedLo, pack_word(need_lo, ST_INCL, epoch));
            st_word(desc + kDwNeedHi, pack_word((uint64_t)1u << 32 | need_hi, ST_INCL, epoch));
        }
        // The need words above and the reads of the predecessors' state words below are the two halves of a store-buffering
        // pattern with a repairing predecessor (it republishes its words, then reads our need words, wt_repair_scan): without a
        // sequentially consistent fence on both sides each could miss the other's store -- we would use a retracted word and
        // the predecessor would not notice.
        // (`prev` needs no fence: a predecessor whose `prev` changed sends everything behind it to the retry round, whatever was read.)
        if (need_lo | need_hi) QB_FENCE_SC();
        if ((need_lo >> lane) & 1u) sm.rec[kIdExt + lane] = wt_resolve_entry(desc, t, lane, ep, init);
        if ((need_hi >> lane) & 1u) sm.rec[kIdExt + 32u + lane] = wt_resolve_entry(desc, t, lane + 32u, ep, init);
        if (lane == 0) sm.rec[kIdExt + 64u] = wt_resolve_entry(desc, t, 64u, ep, init);  // prev: nearly every tile reads it
        __syncwarp();
        // now-known entries become inclusive words, so later tiles stop here
#pragma unroll
        for (int hh = 0; hh < 3; ++hh) {
            const unsigned e = lane + 32u * hh;
            if (pub_ref[hh] <= 64u) {
                const bool known = pub_ref[hh] == 64u || ((pub_ref[hh] < 32u ? need_lo >> pub_ref[hh] : need_hi >> (pub_ref[hh] - 32u)) & 1u);
                if (known) {
                    pub_add[hh] = add4(pub_add[hh], sm.rec[kIdExt + pub_ref[hh]]);
                    st_word(desc + kDwState + e, pack_word((uint64_t)pub_ref[hh] << 32 | pub_add[hh], ST_INCL, epoch));  // value; the entry it came from stays readable
                }
            }
        }
        if constexpr (kStream) {
            if (final_tile) {
#pragma unroll
                for (int hh = 0; hh < 3; ++hh) fin_val[hh] = pub_add[hh];  // every entry is resolved here: the state after the last consumed op
                fin_rem = run_rem, fin_made = final_made, fin_used = final_used, fin_final = true;
            }
        }
        // entry nodes and OP_INDEX ops become absolute
        {
            unsigned bits = idxm;
            for (unsigned i = 0; i <= (unsigned)__popc(idxm); ++i) {
                unsigned x = kIdE + lane;
                if (i) x = opbase + (unsigned)__ffs((int)bits) - 1u, bits &= bits - 1u;
                const unsigned b = sm.base[x];
                if (b != kIdAbs) sm.rec[x] = add4(sm.rec[x], sm.rec[b]), sm.base[x] = (unsigned short)kIdAbs;
            }
        }
        __syncwarp();
        QB_STAMP(desc, 75, 0, qb_t0);  // state look-back

        // ================= emit: values, verification, pixels (lane = op) =================
        // (A lane-serial emit -- every lane running over its own ops with the previous value and the pixel offset in registers,
        // ~20 instructions per 32 ops instead of ~80 -- was measured and lost: its 32 stores per instruction go to 32 different
        // places, ~20-30 cycles of LSU time each; 4K RGB 178 -> 233 us, 8K RGBA 694 -> 786 us.  profiles/r02_experiments.md)
        bool           bad    = false;
        const unsigned live_n = pix_base < N ? (unsigned)(N - pix_base < 0xFFFFFFFFull ? N - pix_base : 0xFFFFFFFFull) : 0u;  // pixels of this tile inside the image
        const unsigned tgt    = P.target;
        uint8_t* const obase  = out + pix_base * tgt;
        // FAST: rows top-down and (four-byte pixels) a word-aligned image; else the general store_pixel
        const bool fast = !P.flip && (tgt == 3u || (reinterpret_cast<uintptr_t>(out) & 3u) == 0);
        {
            unsigned       carry   = sm.rec[kIdExt + 64u];  // value of the op before this step's first one (prev entering the tile)
            unsigned       xcarry  = 0;        // extra OP_RUN pixels before this step
            const unsigned* recp   = sm.rec + lane;
            const unsigned short* basep = sm.base + lane;
            const unsigned char*  slotp = sm.slot + lane;
            auto put = [&](unsigned po, unsigned val) {  // pixel `po` of this tile
                if (!fast) store_pixel(out, pix_base + po, val, P);
                else if (tgt == 4u) reinterpret_cast<unsigned*>(obase)[po] = val;
                else {
                    uint8_t* d = obase + po * 3u;
                    d[0] = (uint8_t)val, d[1] = (uint8_t)(val >> 8), d[2] = (uint8_t)(val >> 16);
                }
            };
            for (unsigned kb = 0; kb < n_keep(); kb += 32u, recp += 32, basep += 32, slotp += 32) {
                const unsigned k     = kb + lane;
                const bool     valid = k < n_keep();
                unsigned       v = 0, sl = 0;
                if (valid) {
                    v                = *recp;
                    const unsigned b = *basep;
                    sl               = *slotp;
                    if (b != kIdAbs) v = add4(v, sm.rec[b]);
                }
                const unsigned ext = sl >= 0xC0u ? sl & 63u : 0u;  // OP_RUN: run - 1 more pixels
                // pixel offset of the op = its ordinal + the extra run pixels before it (a scan only in steps that hold an OP_RUN)
                unsigned       p0    = k + xcarry;
                const unsigned runs0 = __ballot_sync(kFull, ext != 0);
                if (runs0) {
                    unsigned inc = ext;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned o = __shfl_up_sync(kFull, inc, d);
                        if ((int)lane >= d) inc += o;
                    }
                    p0 += inc - ext;
                    xcarry += __shfl_sync(kFull, inc, 31);
                }
                unsigned up = __shfl_up_sync(kFull, v, 1);
                if (lane == 0) up = carry;
                carry = __shfl_sync(kFull, v, 31);
                const bool live = valid && p0 < live_n;  // ops past the image are never executed by the reference
                if (live) {
                    // flags: kSlRgb alone = OP_RGB (simple.cpp:119-123: the alpha is inherited from the previous pixel), kSlIdx alone =
                    // OP_INDEX (a never-written or mis-predicted slot was read when the value does not hash to it)
                    if ((sl & 0xC0u) == kSlRgb && ((v ^ up) >> 24) != 0) {
                        bad              = true;
                        const unsigned f = atomicAdd(&sm.nfail, 1u);
                        if (f < (unsigned)kFixMax) sm.fails[f] = k | (up >> 24) << 16;
                    }
                    if ((sl & 0xC0u) == kSlIdx && slot_of(v) != (sl & 63u)) bad = true;
                    put(p0, v);
                }
                if (runs0) {  // OP_RUN pixels are written by the whole warp (clamped, simple.cpp:158)
                    unsigned runs = runs0 & __ballot_sync(kFull, live);
                    while (runs) {
                        const int src = __ffs((int)runs) - 1;
                        runs &= runs - 1u;
                        const unsigned rp = __shfl_sync(kFull, p0, src), rn = __shfl_sync(kFull, ext, src), rv = __shfl_sync(kFull, v, src);
                        for (unsigned j = 1u + lane; j <= rn && rp + j < live_n; j += 32u) put(rp + j, rv);
                    }
                }
            }
        }
        pix_base_f = pix_base, n_pix_f = n_pix, tail_fill_f = tail_fill;
        still_bad  = __ballot_sync(kFull, bad) != 0;
        if (!still_bad) break;
        const unsigned nfail = sm.nfail;  // refuted OP_RGB ops (an OP_INDEX that read a never-written slot counts in `bad` only)
        if (nfail) wt_record_failures(P, sm, fix, B, cbeg + my_entry, opbase, nops);
        if (nfail == 0 || pass + 1u >= kRepairPasses) break;  // nothing learned, or not converging: the retry rounds take over
#pragma unroll
        for (int i = 0; i < 3; ++i)  // a successor may have read the words of ANY refuted pass: a word counts as changed if two passes disagree on it
            if (1u + 32u * i + lane <= (unsigned)kDwState + 64u) {
                const uint64_t now = ld_word(desc + 1 + 32 * i + lane);
                if ((pass > 0 || cascade) && now != old_word[i]) dirty |= 1u << i;
                old_word[i] = now;
            }
        __syncwarp();
        fixn0 = sm.fixn;
        if (lane == 0) sm.fixn0 = fixn0, sm.nfail = 0;
        __syncwarp();
        }  // repair loop
        unsigned next = kNoRedo;
        if (still_bad) {
            if (cascade) next = kCascadeFail;
            else wt_flag_redo(P, round, res, t);  // the image is eligible for the next round, from this tile on
        } else if (pass || cascade) {  // repaired / decoded again: did anything this tile told the others change, and did it matter?
            Changed c{ 0u, 0u, 0u };
            if (cascade) c = *casc;
            const unsigned nx = wt_repair_scan(P, desc, t, ntiles, ep, old_word, dirty, c);
            if (nx != kNoRedo) {
                if (cascade) next = nx, *casc = c;  // the caller goes on there (or gives up: kScanGaveUp)
                else if (round == 0 && P.req_cap && !(nx & kScanGaveUp)) wt_flag_cascade(P, res, img, nx, c);
                else wt_flag_redo(P, round, res, nx & ~kScanGaveUp);
            }
        }
        QB_STAMP(desc, 75, 1, qb_t0);  // emit

        if constexpr (kStream) {
            // ---- the final tile reports what the call consumed and produced and the state to carry on with (stream.cpp:418-423)
            if (fin_final && !still_bad) {
                res->state.table[lane] = fin_val[0], res->state.table[lane + 32] = fin_val[1];
                if (lane == 0) {
                    res->state.prev = fin_val[2], res->state.run = fin_rem;
                    res->processed = fin_used;
                    res->written   = fin_made * P.target;
                    res->pixels    = fin_made;
                }
            }
            __syncwarp();
            return next;
        }
        // ---- the stream ended before the image: the zero padding decodes as INDEX 0 forever (simple.cpp:106,132-135)
        if (t == ntiles - 1) {
            const uint64_t have = pix_base_f + n_pix_f;
            if (lane == 0) res->pixels = have < N ? have : N;
            if (tail_fill_f) {
                unsigned fill = sm.rec[kIdExt + 0u];  // table[0] after this tile
                if (sm.lastk[0]) {
                    const unsigned k0 = sm.lastk[0] - 1u;
                    fill             = sm.rec[k0];
                    const unsigned b = sm.base[k0];
                    if (b != kIdAbs) fill = add4(fill, sm.rec[b]);
                }
                if (lane == 0 && slot_of(fill) != 0) {  // cannot happen while the table invariant holds; be safe
                    atomicMax(&res->first_bad[round], 0xFFFFFFFFu - t);
                    P.control->any_bad[round] = 1;
                }
                for (uint64_t pix = have + lane; pix < N; pix += 32u) store_pixel(out, pix, fill, P);
            }
        }
        __syncwarp();
        return next;
    }

    template <bool kStream>
    __device__ __forceinline__ void wt_round0(const DecParams& P)
    {
        uint2*  lut = reinterpret_cast<uint2*>(QB_DYN_SMEM);
        WtSmem& sm  = reinterpret_cast<WtSmem*>(QB_DYN_SMEM + kWtLutBytes)[threadIdx.x >> 5];
        wt_build_lut(lut);
#if QB_WT_LOCKSTEP
        // The warps of a CTA draw kWtWarps consecutive tiles together and start them together: the tile body is ~90 KB of
        // straight-line code, far more than the 32 KB instruction cache of an SM, and warps that drift apart each keep their own
        // part of it in flight ("no instruction" was the largest stall of the kernel: 2.8 of 11 stall cycles per issued
        // instruction, 4.9 after the tile body had grown by 4 %).  Measured, 8K RGBA: independent warps 824 us, 4 / 8 / 12 warps
        // in step 698 / 631 / 620 us; barriers INSIDE the tile (after staging, look-back 1, the walk, look-back 3, the state
        // look-back) gain nothing more.  No tile waits for a warp of its own CTA at the barrier while that warp waits for one
        // of its words: every word a tile owes its successors is published before the tile ends.
        unsigned* base = reinterpret_cast<unsigned*>(QB_DYN_SMEM + kWtLutBytes + sizeof(WtSmem) * kWtWarps);
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) *base = atomicAdd(&P.control->tickets[0], (unsigned)kWtWarps);
            __syncthreads();
            const unsigned x0 = *base;
            if (x0 >= P.n_tiles) break;
            const unsigned x = x0 + (threadIdx.x >> 5);
            if (x >= P.n_tiles) continue;
#else
        const unsigned lane = threadIdx.x & 31u;
        for (;;) {
            unsigned x = 0;
            if (lane == 0) x = atomicAdd(&P.control->tickets[0], 1u);
            x = __shfl_sync(kFull, x, 0);
            if (x >= P.n_tiles) break;
#endif
            if constexpr (kStream) {
                wt_decode_tile<true>(P, sm, lut, 0u, x, 0u, x, P.n_tiles, P.qoi + P.single[0], P.single[1] - P.single[0], 0u);
            } else {
                unsigned       img, t, ntiles;
                const uint8_t* stream;
                uint64_t       size;
                locate_image(P, x, img, t, ntiles, stream, size);
                wt_decode_tile<false>(P, sm, lut, 0u, x, img, t, ntiles, stream, size, 0u);
            }
        }
    }
    // round 0: persistent warps draw tiles from a ticket counter in start order, so every tile a running warp waits for is
    // held by a warp that is running too or done
    __global__ void __launch_bounds__(kWtThreads, QB_WT_CTAS) decode_wt_kernel(const DecParams P) { wt_round0<false>(P); }
    // resumable decode (StreamDecoder::decode): the same over the tiles of one input buffer, carry-in P.init
    __global__ void __launch_bounds__(kWtThreads, QB_WT_CTAS) decode_wt_stream_kernel(const DecParams P) { wt_round0<true>(P); }

    // =====================================================================================================
    // Exact sequential decoder: the reference loop (simple.cpp:100-171, stream.cpp:312-447) with one decoding lane per
    // image; the other 31 lanes stage input and output through shared memory.  Runs for images the retry rounds could not
    // verify (mode 0), and for the resumable entry point (mode 1).
    // =====================================================================================================
    struct SerialParams {
        DecParams       d;
        uint32_t        mode;      // 0 = redo images flagged bad; 1 = resumable decode of one buffer
        const DecState* init;      // mode 1 carry-in
        uint64_t        in_size;   // mode 1: bytes available (no header), out capacity in bytes is d.out_stride
        uint32_t        only_if_bad;  // mode 1 behind the parallel kernels: run only when the last retry round still refuted a speculation
    };

    constexpr int kSerIn = 4096, kSerOut = 1024;
    struct SerialSmem {
        unsigned char in[kSerIn + 16];
        unsigned      px[kSerOut];
        unsigned      table[64];
        unsigned      ctl[8];
    };

    // one warp; `img` selects the image (mode 0) -- called by decode_finish_kernel and decode_serial_kernel
    __device__ __forceinline__ void decode_serial_body(const SerialParams& S, SerialSmem& sm, unsigned img)
    {
        const DecParams&      P    = S.d;
        const unsigned        lane = threadIdx.x & 31u;
        DecResult*            res  = P.results + img;
        unsigned restart = 0;  // mode 0: first tile to decode again (the tiles before it verified in some round)
        if (S.mode == 1 && S.only_if_bad && res->first_bad[kDecRounds] == 0) return;  // the parallel kernels' result stands
        if (S.mode == 0) {
            unsigned rounds = 0;
            for (int r = 0; r < kDecRounds; ++r)
                if (res->first_bad[r]) rounds = r + 1;
            const unsigned fb = res->first_bad[kDecRounds];
            __syncwarp();
            if (lane == 0) res->path = rounds + (fb ? 100u : 0u);
            if (fb == 0) return;
            restart = 0xFFFFFFFFu - fb;
        }

        const uint8_t* stream;
        uint64_t       size;
        unsigned       first_tile = 0;
        if (P.tile_first == nullptr) stream = P.qoi + P.single[0], size = P.single[1] - P.single[0];
        else stream = P.qoi + P.offsets[2u * img], size = P.offsets[2u * img + 1u] - P.offsets[2u * img], first_tile = P.tile_first[img];
        uint8_t*       out   = P.out + (uint64_t)img * (S.mode == 0 ? P.out_stride : 0);
        const uint8_t* body  = S.mode == 0 ? stream + kHeader : stream;
        const uint64_t blen  = S.mode == 0 ? size - kHeader : S.in_size;
        const uint64_t room  = S.mode == 0 ? P.n_pixels : P.out_stride / P.target;  // pixels that may be produced

        sm.table[lane] = 0, sm.table[lane + 32] = 0;
        __syncwarp();
        unsigned prev = kStartPixel, run = 0;
        uint64_t pos = 0, px = 0;  // consumed input bytes, produced pixels
        if (S.mode == 1) {
            prev = S.init->prev, run = S.init->run;
            sm.table[lane] = S.init->table[lane], sm.table[lane + 32] = S.init->table[lane + 32];
        } else if (restart > 0) {
            // resume behind the last verified tile: the decoder state there follows from the verified tiles' transfer words
            const uint64_t* d_r = P.desc + (uint64_t)(first_tile + restart) * kDecDescWords;  // descriptor of tile `restart`
            const uint64_t* d   = d_r - kDecDescWords;
            const Epochs    ep{ P.epoch + (unsigned)kDecRounds + 1u, P.epoch, restart };  // any epoch of this decode is final below `restart`
            pos                 = (uint64_t)restart * kDecTB + (word_payload(d[kDwParse]) & 7u);
            {
                const PixA before = wt_gather_pixa(d_r, restart, ep);
                px                = (uint64_t)before.hi << 32 | before.lo;
            }
            prev                = wt_resolve_entry(d_r, restart, 64u, ep);
            sm.table[lane] = wt_resolve_entry(d_r, restart, lane, ep), sm.table[lane + 32] = wt_resolve_entry(d_r, restart, lane + 32u, ep);
            prev = __shfl_sync(kFull, prev, 0);
        } else if (lane == 0) {
            sm.table[53] = kStartPixel;  // simple.cpp:108
        }
        __syncwarp();

        bool     stop = false;
        while (!stop && px < room) {
            // stage kSerIn bytes from `pos` (zero padded past the end: simple.cpp:106)
            for (unsigned b = lane; b < kSerIn + 16; b += 32) sm.in[b] = pos + b < blen ? body[pos + b] : 0;
            __syncwarp();
            if (lane == 0) {
                // The next bytes of the stream live in a 64-bit register window refilled from prefetched words, so the
                // tag -> length -> next tag chain never waits for shared memory.
                const unsigned* in32 = reinterpret_cast<const unsigned*>(sm.in);
                uint64_t        win  = (uint64_t)in32[0] | (uint64_t)in32[1] << 32;
                unsigned        avail = 8, wnext = 4, n0 = in32[2], n1 = in32[3];
                unsigned        ip = 0, op = 0;
                while (op < kSerOut && px + op < room) {
                    if (run) {  // pending run (stream.cpp:335-339)
                        --run, sm.px[op++] = prev;
                        continue;
                    }
                    if (ip >= kSerIn) break;
                    const unsigned tag = (unsigned)win & 0xFFu, len = op_length(tag);
                    if (S.mode == 1 && pos + ip + len > blen) { stop = true; break; }  // stream.cpp:341-392: incomplete op is not consumed
                    const unsigned pay = (unsigned)(win >> 8);  // the four bytes after the tag
                    unsigned       cur = prev;
                    bool           is_run = false;
                    if (tag == kOpRgb) cur = (pay & 0xFFFFFFu) | (prev & 0xFF000000u);  // simple.cpp:119-123
                    else if (tag == kOpRgba) cur = pay;
                    else if ((tag >> 6) == 0) cur = sm.table[tag & 63u];
                    else if ((tag >> 6) == 1)
                        cur = add4(prev, add4(((tag >> 4) & 3u) | ((tag >> 2) & 3u) << 8 | (tag & 3u) << 16, 0x00FEFEFEu));
                    else if ((tag >> 6) == 2) {
                        const unsigned rb = pay & 0xFFu, vg = ((tag & 63u) + 224u) & 255u;
                        cur = add4(prev, ((vg + (rb >> 4) + 248u) & 255u) | vg << 8 | ((vg + (rb & 15u) + 248u) & 255u) << 16);
                    } else {
                        run = tag & 63u, is_run = true;  // RUN: one pixel now, the rest pending (simple.cpp:156-163)
                    }
                    ip += len, avail -= len;
                    win = len == 8 ? 0 : win >> (8u * len);
                    while (avail <= 4) {
                        win |= (uint64_t)n0 << (8u * avail);
                        avail += 4, n0 = n1, n1 = in32[wnext < (kSerIn + 16) / 4 ? wnext : 0];
                        ++wnext;
                    }
                    sm.px[op++] = cur;
                    if (!is_run) sm.table[slot_of(cur)] = cur;  // simple.cpp:169
                    prev = cur;
                }
                sm.ctl[0] = ip, sm.ctl[1] = op, sm.ctl[2] = stop;
            }
            __syncwarp();
            const unsigned ip = sm.ctl[0], op = sm.ctl[1];
            stop = sm.ctl[2] != 0;
            for (unsigned j = lane; j < op; j += 32) store_pixel(out, px + j, sm.px[j], P);
            __syncwarp();
            pos += ip, px += op;
            if (ip == 0 && op == 0) break;
        }
        if (S.mode == 0) {
            if (lane == 0) res->pixels = px;
            return;
        }
        // mode 1 carry-out.  A run that is still pending when the input of mode 0 ends is dropped by the clamp
        // (simple.cpp:158); in mode 1 it stays in the state (stream.cpp:405-409).
        prev = __shfl_sync(kFull, prev, 0), run = __shfl_sync(kFull, run, 0);
        res->state.table[lane] = sm.table[lane], res->state.table[lane + 32] = sm.table[lane + 32];
        if (lane == 0) {
            res->state.prev = prev, res->state.run = run;
            res->processed = pos, res->written = px * P.target, res->path = 1;
        }
    }

    // resumable decode (mode 1): one warp
    __global__ void __launch_bounds__(32) decode_serial_kernel(const SerialParams S)
    {
        decode_serial_body(S, *reinterpret_cast<SerialSmem*>(QB_DYN_SMEM), 0u);
    }

    constexpr size_t kWtSmemBytes = kWtLutBytes + (sizeof(WtSmem) * kWtWarps > sizeof(SerialSmem) ? sizeof(WtSmem) * kWtWarps : sizeof(SerialSmem)) + 16;
    static_assert(sizeof(WtSmem) % 16 == 0, "per-warp areas stay 16-byte aligned");

    // Everything after round 0, in ONE cooperative launch of co-resident persistent CTAs (so that an image that
    // verified costs a single empty launch): rounds 1..kDecRounds re-decode, per image, the tiles from the first refuted
    // one on with the alphas learned by the round before (grid-wide barrier between rounds); what still fails after the
    // last round is decoded by the sequential loop, one warp per image, resuming behind the last verified tile.
    template <bool kStream>
    __device__ __forceinline__ void decode_finish_body(const DecParams& P)
    {
        uint2*         lut  = reinterpret_cast<uint2*>(QB_DYN_SMEM);
        WtSmem&        sm   = reinterpret_cast<WtSmem*>(QB_DYN_SMEM + kWtLutBytes)[threadIdx.x >> 5];
        const unsigned lane = threadIdx.x & 31u;
        if (P.control->any_bad[0] == 0 && P.control->any_bad[kDecRounds] == 0) return;  // everything verified in round 0 (same value in every CTA)
        wt_build_lut(lut);
        // ---- cascade: a tile that consumed a word its predecessor retracted in round 0 is decoded again, then whatever read ITS
        // changed words, ... -- typically two or three tiles per request.  One warp per image takes the image's requests in
        // ascending tile order (a served request leaves everything up to where it ends exact; two requests of one image never
        // run at the same time: a look-back of the later one may reach into the tiles the earlier one is rewriting).  An image
        // whose requests need more than kCascadeBudget tile decodes, a request that finds a tile still refuted, and a scan
        // that cannot tell send the image to the retry rounds from there.
        {
            constexpr unsigned kCascadeBudget = 48;
            const unsigned gw = blockIdx.x * kWtWarps + (threadIdx.x >> 5), nw = gridDim.x * kWtWarps;
            const unsigned total = min(P.control->n_req, P.req_cap);
            for (unsigned img = gw; img < P.n_images && total; img += nw) {
                DecResult* res = P.results + img;
                if (!(res->bad & 2u)) continue;
                unsigned       first = 0, ntiles = P.n_tiles;
                const uint8_t* stream = P.qoi + P.single[0];
                uint64_t       size   = P.single[1] - P.single[0];
                if (P.tile_first) {
                    first = P.tile_first[img], ntiles = P.tile_first[img + 1] - first;
                    stream = P.qoi + P.offsets[2u * img], size = P.offsets[2u * img + 1u] - P.offsets[2u * img];
                }
                unsigned fail = kNoRedo, budget = kCascadeBudget, last = 0;
                bool     any = false;
                const bool rounds_anyway = res->first_bad[0] != 0;  // a tile of round 0 stayed refuted: only the earliest request matters
                while (fail == kNoRedo) {
                    // the smallest requested tile above the last one served, with everything that was asked for it
                    unsigned u = kNoRedo;
                    for (unsigned j0 = 0; j0 < total; j0 += 32u) {
                        const unsigned j = j0 + lane;
                        unsigned tj = kNoRedo;
                        if (j < total && P.req[j].img == img && (!any || P.req[j].tile > last)) tj = P.req[j].tile;
                        u = min(u, __reduce_min_sync(kFull, tj));
                    }
                    if (u == kNoRedo) break;
                    if (rounds_anyway) { fail = u; break; }
                    Changed c{ 0u, 0u, 0u };
                    for (unsigned j0 = 0; j0 < total; j0 += 32u) {
                        const unsigned j = j0 + lane;
                        unsigned lo = 0, hi = 0, pv = 0;
                        if (j < total && P.req[j].img == img && P.req[j].tile == u) lo = P.req[j].lo, hi = P.req[j].hi, pv = P.req[j].prev;
                        c.lo |= __reduce_or_sync(kFull, lo), c.hi |= __reduce_or_sync(kFull, hi), c.prev |= __reduce_or_sync(kFull, pv);
                    }
                    any = true, last = u;
                    unsigned v = u;
                    while (v != kNoRedo && v < ntiles) {
                        if (budget-- == 0) { fail = v; break; }
                        const unsigned nx = wt_decode_tile<kStream>(P, sm, lut, 0u, first + v, img, v, ntiles, stream, size, 0u, &c);
                        if (nx == kCascadeFail) { fail = v; break; }
                        if (nx != kNoRedo && (nx & kScanGaveUp)) { fail = nx & ~kScanGaveUp; break; }
                        v = nx;
                    }
                }
                if (lane == 0 && fail != kNoRedo && fail < ntiles) atomicMax(&res->first_bad[0], 0xFFFFFFFFu - fail), P.control->rounds_needed = 1;
                __syncwarp();
            }
            QB_GRID_SYNC();
        }
        for (unsigned round = 1; round <= (unsigned)kDecRounds; ++round) {
            if ((round == 1 ? P.control->rounds_needed : P.control->any_bad[round - 1]) == 0) break;  // same value in every CTA: final since the last barrier
            // Work list of the round: the images that need it, each with the range of tiles from its first affected one on, in
            // the (dead by now) request array; tickets run over the concatenated ranges.  (One ticket and one image search per
            // tile of the whole batch -- 4.5 M tiles for 8192 images, a handful of them with work -- cost 12 ms.)
            const unsigned gw = blockIdx.x * kWtWarps + (threadIdx.x >> 5), nw = gridDim.x * kWtWarps;
            uint32_t*      nlist = &P.control->list_n;  // entries in the list; reset below for the next round
            bool           listed = P.n_images > 1 && P.req_cap >= 2;
            if (listed) {
                for (unsigned img = gw * 32u + lane; img < P.n_images; img += nw * 32u) {
                    const unsigned fb = P.results[img].first_bad[round - 1];
                    if (fb) {
                        const unsigned from = 0xFFFFFFFFu - fb, i = atomicAdd(nlist, 1u);
                        if (i < P.req_cap) P.req[i] = CascadeReq{ img, from, P.tile_first[img + 1] - P.tile_first[img] - from, 0u, 0u, 0u };
                    }
                }
                QB_GRID_SYNC();
                const unsigned K = *nlist;
                listed = K <= P.req_cap;  // else: the full loop below
                if (listed && gw == 0) {  // exclusive prefix of the range lengths -> .hi
                    unsigned run = 0;
                    for (unsigned j0 = 0; j0 < K; j0 += 32u) {
                        const unsigned j = j0 + lane, n = j < K ? P.req[j].lo : 0u;
                        unsigned inc = n;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const unsigned o = __shfl_up_sync(kFull, inc, d);
                            if ((int)lane >= d) inc += o;
                        }
                        if (j < K) P.req[j].hi = run + inc - n;
                        run += __shfl_sync(kFull, inc, 31);
                    }
                    if (lane == 0) P.control->list_tiles = run;
                }
                QB_GRID_SYNC();
            }
            const unsigned K = listed ? *nlist : 0u, W = listed ? P.control->list_tiles : P.n_tiles;
            for (;;) {
                unsigned x = 0;
                if (lane == 0) x = atomicAdd(&P.control->tickets[round], 1u);
                x = __shfl_sync(kFull, x, 0);
                if (x >= W) break;
                unsigned       img, t, ntiles, gt;
                const uint8_t* stream;
                uint64_t       size;
                if (listed) {
                    unsigned lo = 0, hi = K;  // the range that holds ticket x
                    while (hi - lo > 1) {
                        const unsigned mid = (lo + hi) >> 1;
                        if (P.req[mid].hi <= x) lo = mid;
                        else hi = mid;
                    }
                    img = P.req[lo].img, t = P.req[lo].tile + (x - P.req[lo].hi);
                    const unsigned f = P.tile_first[img];
                    gt = f + t, ntiles = P.tile_first[img + 1] - f;
                    const uint64_t o0 = P.offsets[2u * img];
                    stream = P.qoi + o0, size = P.offsets[2u * img + 1u] - o0;
                } else {
                    gt = x;
                    locate_image(P, x, img, t, ntiles, stream, size);
                }
                const unsigned fb = P.results[img].first_bad[round - 1];
                if (fb == 0 || t < 0xFFFFFFFFu - fb) continue;  // image verified, or a tile before the first refuted one: final
                wt_decode_tile<kStream>(P, sm, lut, round, gt, img, t, ntiles, stream, size, 0xFFFFFFFFu - fb);
            }
            if (listed) {
                QB_GRID_SYNC();
                if (gw == 0 && lane == 0) *nlist = 0;
            }
            QB_GRID_SYNC();
        }
        if constexpr (!kStream) {
            __syncthreads();
            if (threadIdx.x < 32) {
                SerialParams S{};
                S.d = P, S.mode = 0;
                for (unsigned img = blockIdx.x; img < P.n_images; img += gridDim.x)
                    decode_serial_body(S, *reinterpret_cast<SerialSmem*>(QB_DYN_SMEM), img);
            }
        }
    }
    __global__ void __launch_bounds__(kWtThreads, QB_WT_CTAS) decode_finish_kernel(const DecParams P) { decode_finish_body<false>(P); }
    // resumable decode: the retry rounds only; what the last round still refutes is decoded by decode_serial_kernel (mode 1) behind it
    __global__ void __launch_bounds__(kWtThreads, QB_WT_CTAS) decode_finish_stream_kernel(const DecParams P) { decode_finish_body<true>(P); }

    // device-side epilogue of the *_stream_*_dev entry points: results of the call -> the caller's result and state blocks
    struct StreamOut {
        uint64_t processed, written;
    };
    __global__ void stream_decode_epilogue_kernel(const DecResult* res, StreamOut* out, DecState* state)
    {
        const unsigned i = threadIdx.x;
        if (i < 64u) state->table[i] = res->state.table[i];
        if (i == 0) state->prev = res->state.prev, state->run = res->state.run, out->processed = res->processed, out->written = res->written;
    }

#ifndef QB_EMU
    inline cudaError_t dec_set_attrs()
    {
        cudaError_t e = cudaFuncSetAttribute(decode_wt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWtSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWtSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_finish_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWtSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_finish_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_wt_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWtSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_wt_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_finish_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(decode_wt_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
#endif
}  // namespace qb
