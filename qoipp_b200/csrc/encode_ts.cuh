// encode_ts.cuh -- thread-serial, warp-per-tile encoder: the one-shot fast path of the QOI encoder for sm_100a.
//
// Same result as encode_kernel (encode_kernel.cuh) and the reference loop impl::encode (source/simple.cpp:17-98),
// byte for byte; chosen by the host when the call is a plain one-shot encode (no carried state, capacity >= worst
// size, 16-byte aligned input).  encode_kernel gives every LANE one pixel per step and pays for it in shuffles,
// ballots and __match_any_sync (262 thread-instructions per pixel, issue bound).  Here every THREAD walks kTsK = 32
// consecutive pixels like the reference loop does -- previous pixel, run counter and output window live in
// registers -- and a tile is the 1024 pixels of ONE WARP, so nothing in the kernel needs __syncthreads: the warps of a
// CTA are independent, persistent workers that draw tiles from a ticket counter (measured with 128-thread tiles: a
// third of all warp cycles were spent at CTA barriers behind the look-backs).  Only the 64-entry "seen" table needs
// help, because a thread does not know the table at the start of its chunk:
//
//   W1     each lane records, per slot, the last pixel it would store (simple.cpp:57) in its own column of the warp's
//          tab[slot][lane] (row stride 33 words), other slots keep a sentinel;
//   merge  an exclusive "last writer wins" scan along the lanes turns column t into the table on entry to lane t's
//          chunk (lane = slot, two rounds of 32 slots, conflict free); the tile's own last writers fall out;
//   carry  per-slot last writer and last differing pixel are published and looked back as in encode_kernel (same
//          72-word record per tile, now per 1024 pixels);
//   W3     the reference loop proper: chunk for "differs, no hit" (DIFF / LUMA / RGB / RGBA), probe + store per
//          differing pixel, run counter; chunks are appended through a 64-bit register window to the lane's
//          word-aligned slice of the tile's SCRATCH record in global memory ([word][lane]: whole 128-byte lines);
//          then the tile's byte count is published and the warp is done with the tile;
//   copy   a second kernel (encode_ts_copy_kernel, one warp per tile) turns the records into the final bytes: the tile's
//          offset is a sum of 64-tile group totals (accumulated by the encode kernel with one atomicAdd per tile) and of
//          the byte counts of the earlier tiles of its group -- no chain, nothing to wait for --, then the 32 per-lane
//          counts are prefix-summed, the slices funnel-shifted into one contiguous run in shared memory and copied out
//          as realigned 16-byte stores.  Measured alternatives (profiles/r01_experiments.md): byte carry inside the encode
//          pass (every tile synchronises to the slowest of its 32 predecessors), copy role inside the same kernel behind
//          a ticket lag (the copying warps wait for stragglers half of their time).
#pragma once

#include "encode_kernel.cuh"

namespace qb
{
#ifndef QB_TS_WARPS
#define QB_TS_WARPS 4  // independent warp workers per CTA
#endif
#ifndef QB_TS_CTAS
#define QB_TS_CTAS 6  // CTAs per SM (24 warps; 80 registers per thread)
#endif
    constexpr int kTsWarps = QB_TS_WARPS, kTsThreads = kTsWarps * 32, kTsK = 32, kTsT = 32 * kTsK;  // tile = 1024 pixels
    constexpr int kTsRow = 33;  // words per table row: bank = (slot + lane) & 31, the merge (lane = slot) is conflict free

    template <int CH>
    struct TsCfg {
        static constexpr int kPrivWords = kTsK * (CH + 1) / 4 + 1;  // per lane: its chunks, word aligned (+ the partial word)
        static constexpr int kScrWords  = (kPrivWords + 1) * 32;    // scratch record of a tile: [word][lane], then the counts
    };

    constexpr int kTsCopyWarps = 8;  // encode_ts_copy_kernel: tiles (warps) per CTA
    struct TsCopySmem {
        alignas(16) unsigned tab[(kTsT * 5 + 64) / 4];  // the tile's staging bytes
    };

    struct TsWarpSmem {
        alignas(16) unsigned tab[64 * kTsRow];  // [slot][lane]
        unsigned gin[64];                       // table on entry to the tile
    };

    // chunk of a pixel that differs from its predecessor and missed the table (simple.cpp:59-79, util.hpp:163-225):
    // first four bytes in `chunk` (an RGBA chunk's fifth byte is the pixel's alpha), length in `len`.
    // r/b are handled as two 16-bit lanes of one register, biased so that no borrow crosses the lanes.
    template <int CH>
    __device__ __forceinline__ void ts_colour_chunk(unsigned cur, unsigned prv, unsigned& chunk, unsigned& len)
    {
        const unsigned tg  = ((cur >> 8) & 0xFFu) - ((prv >> 8) & 0xFFu) + 2u;           // low byte: dg + 2
        const unsigned trb = (cur & 0x00FF00FFu) - (prv & 0x00FF00FFu) + 0x04020402u;    // low byte of each lane: d + 2
        const bool     diff_ok = ((trb & 0x00FC00FCu) | (tg & 0xFCu)) == 0;              // util.hpp:102-107
        const unsigned diffb   = kOpDiff | (trb & 3u) << 4 | (tg & 3u) << 2 | ((trb >> 16) & 3u);
        const unsigned vg = (tg + 30u) & 0xFFu;                                           // dg + 32
        const unsigned x  = trb + 0x00080008u - (tg & 0xFFu) * 0x00010001u;               // low bytes: dr-dg+8, db-dg+8
        const bool     luma_ok = ((x & 0x00F000F0u) | (vg & 0xC0u)) == 0;                 // util.hpp:109-114
        const unsigned lumab   = kOpLuma | vg | ((x & 15u) << 4 | ((x >> 16) & 15u)) << 8;
        const bool     alpha_ne = CH == 4 && ((cur ^ prv) >> 24) != 0;
        const unsigned lit = __byte_perm(cur, alpha_ne ? kOpRgba : kOpRgb, 0x2104);       // tag r g b
        chunk = alpha_ne ? lit : (diff_ok ? diffb : (luma_ok ? lumab : lit));
        len   = alpha_ne ? 5u : (diff_ok ? 1u : (luma_ok ? 2u : 4u));
    }

    // G consecutive pixels of a lane's chunk (group g) from global memory.  Complete chunks use 16-byte loads; in the image's
    // last tile pixels past the end repeat the last valid one (`lastv`), so they never differ and `vmask` drops them.
    template <int CH, int G>
    __device__ __forceinline__ void ts_load_group(const uint8_t* in_img, uint64_t g0, unsigned g, unsigned nvalid, unsigned& lastv,
                                                  unsigned (&px)[G])
    {
        if (nvalid == (unsigned)kTsK) {
            if (CH == 4) {
                const uint4* src = reinterpret_cast<const uint4*>(in_img + g0 * 4) + g * (G / 4);
#pragma unroll
                for (int j = 0; j < G / 4; ++j) {
                    const uint4 v = __ldg(src + j);
                    px[4 * j] = v.x, px[4 * j + 1] = v.y, px[4 * j + 2] = v.z, px[4 * j + 3] = v.w;
                }
            } else {
                const uint2* src = reinterpret_cast<const uint2*>(in_img + g0 * 3) + g * (3 * G / 8);  // 3 G bytes, 8-byte aligned
                unsigned     wd[3 * G / 4];
#pragma unroll
                for (int j = 0; j < 3 * G / 8; ++j) {
                    const uint2 v = __ldg(src + j);
                    wd[2 * j] = v.x, wd[2 * j + 1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < G / 4; ++j) {  // four pixels from three words; util.hpp:325: alpha forced to 255
                    const unsigned a = wd[3 * j], b = wd[3 * j + 1], c = wd[3 * j + 2];
                    px[4 * j]     = a | 0xFF000000u;
                    px[4 * j + 1] = __funnelshift_r(a, b, 24) | 0xFF000000u;
                    px[4 * j + 2] = __funnelshift_r(b, c, 16) | 0xFF000000u;
                    px[4 * j + 3] = (c >> 8) | 0xFF000000u;
                }
            }
            lastv = px[G - 1];
        } else {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const unsigned k = g * G + i;
                if (k < nvalid) lastv = load_pixel<CH>(in_img, g0 + k, true);
                px[i] = lastv;
            }
        }
    }

    // ---- encode role: tile `gt` (global tile index) -> scratch record + carry words.  One warp.
    template <int CH>
    __device__ __forceinline__ void ts_encode_tile(const EncParams& P, TsWarpSmem& sm, unsigned gt)
    {
        using C          = TsCfg<CH>;
        constexpr int K  = kTsK, T = kTsT, R = kTsRow;
        const unsigned lane = threadIdx.x & 31u;
        [[maybe_unused]] const long long qb_t0 = QB_T0();

        const unsigned  img        = gt / P.tiles_per_image;
        const unsigned  t          = gt % P.tiles_per_image;
        const uint8_t*  in_img     = P.in + (uint64_t)img * P.in_stride;
        const uint64_t  N          = P.n_pixels;
        const uint64_t  tile_start = (uint64_t)t * T;
        const unsigned  n_here     = (unsigned)(N - tile_start < (uint64_t)T ? N - tile_start : (uint64_t)T);
        uint64_t*       desc       = P.desc + (uint64_t)gt * kEncDescWords;
        const unsigned  epoch      = P.epoch;
        const unsigned  first      = lane * K;  // tile-local index of this lane's first pixel
        const unsigned  nvalid     = first >= n_here ? 0u : min((unsigned)K, n_here - first);
        const uint64_t  g0         = tile_start + first;

        // The chunk is walked in groups of G pixels (loops, not 32 unrolled copies: unrolled, the kernel was 80 KB of code and
        // instruction fetch was its largest stall, ncu no_instruction 2.7 per issued instruction); W1 and W3 each read the
        // pixels once, the second time from L2.
        constexpr int G = 8, NG = K / G;
        // neighbours across the chunk boundary and the first group; the loads are in flight while the column is initialised
        const unsigned prev0 = g0 == 0 ? kStartPixel : (g0 <= N ? load_pixel<CH>(in_img, g0 - 1, true) : 0u);
        const unsigned nxt   = g0 + K < N ? load_pixel<CH>(in_img, g0 + K, true) : 0u;
        unsigned       px[G], nx[G];
        unsigned       lastv = prev0;
        ts_load_group<CH, G>(in_img, g0, 0u, nvalid, lastv, px);
        // own table column: nothing stored yet (sentinel(s))
#pragma unroll
        for (int s = 0; s < 64; ++s) sm.tab[s * R + lane] = s == 0 ? 1u : 0u;
        QB_STAMP(desc, 66, 0, qb_t0);  // setup

        // ================= W1: per-slot last store of this chunk, mask of differing pixels =================
        unsigned neMask = 0, last_px;
        {
            unsigned p = prev0;
#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                if (g + 1 < NG) ts_load_group<CH, G>(in_img, g0, (unsigned)g + 1u, nvalid, lastv, nx);  // in flight during this group
                unsigned bits = 0;
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const unsigned cur = px[i];
                    if (cur != p) {
                        bits |= 1u << i;
                        sm.tab[slot_of(cur) * R + lane] = cur;  // simple.cpp:54-57: every differing pixel ends up in its slot
                    }
                    p = cur;
                }
                neMask |= bits << (g * G);
#pragma unroll
                for (int i = 0; i < G; ++i) px[i] = nx[i];
            }
            last_px = p;
        }
        const bool nexteq_last = g0 + K < N && nxt == last_px;  // does a run continue into the next lane's chunk?
        // the first group of the second pass is fetched (from L2 now) before the merge and the look-backs, not after them
        lastv = prev0;
        ts_load_group<CH, G>(in_img, g0, 0u, nvalid, lastv, px);
        const unsigned vmask  = nvalid == (unsigned)K ? 0xFFFFFFFFu : (1u << nvalid) - 1u;
        const unsigned eqMask = ~neMask & vmask;
        // last differing pixel before this lane's chunk (tile-local index + 1, 0 = none): exclusive max-scan
        unsigned before, tile_last;
        {
            unsigned inc = neMask ? first + (31u - (unsigned)__clz((int)neMask)) + 1u : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(kFull, inc, d);
                if ((int)lane >= d) inc = max(inc, o);
            }
            before = __shfl_up_sync(kFull, inc, 1);
            if (lane == 0) before = 0;
            tile_last = __shfl_sync(kFull, inc, 31);
        }
        // carry (2), run: publish at once (nothing else is needed for it)
        if (lane == 0) {
            if (tile_last) st_word(desc + kWordLne, pack_word(tile_start + (tile_last - 1u) + kLneBias, ST_INCL, epoch));
            else st_word(desc + kWordLne, pack_word(0, ST_AGG_EMPTY, epoch));
        }
        __syncwarp();
        QB_STAMP(desc, 66, 1, qb_t0);  // W1

        // ================= merge: column t := table on entry to lane t's chunk; publish the tile's last writers =================
        unsigned own[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned s = 32u * h + lane, sent = sentinel(s);
            unsigned*      row = sm.tab + s * R;
            unsigned       cur = sent;
#pragma unroll
            for (int t0 = 0; t0 < 32; t0 += 8) {  // loads of a batch first, they are independent
                unsigned x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = row[t0 + i];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    row[t0 + i] = cur;
                    if (x[i] != sent) cur = x[i];
                }
            }
            own[h] = cur;
            st_word(desc + s, pack_word(cur, cur != sent ? ST_INCL : ST_AGG_EMPTY, epoch));  // carry (1), table
        }
        QB_STAMP(desc, 67, 0, qb_t0);  // merge

        // ================= look back: table on entry to the tile, run counter on entry =================
        // the first probes of all three look-backs are in flight together
        const uint64_t* prev_rec = desc - kEncDescWords;
        uint64_t        first_wd[2] = { 0, 0 };
        if (t > 0) first_wd[0] = ld_word(prev_rec + lane), first_wd[1] = ld_word(prev_rec + 32 + lane);
        unsigned base62;
        {
            // payload = last differing pixel index + bias (never 0); the later tile wins when it has one
            const uint64_t lp = warp_lookback_lazy<uint64_t>(
                t, (uint64_t)(kLneBias - 1), (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(desc - (int64_t)(t - p) * kEncDescWords + kWordLne);
                    st                = word_status(wd, epoch);
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return b ? b : a; });
            if (lane == 0 && !tile_last) st_word(desc + kWordLne, pack_word(lp, ST_INCL, epoch));
            base62 = (unsigned)((tile_start + kLneBias - lp) % kRunLimit);  // (tile_start - last differing) mod 62
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned s = 32u * h + lane;
            unsigned       e;
            int            p  = (int)t - 1;
            uint64_t       wd = first_wd[h];
            for (;;) {
                if (p < 0) { e = 0u; break; }  // simple.cpp:28: zero-initialised table
                const uint64_t* src = desc - (int64_t)(t - p) * kEncDescWords + s;
                while (word_status(wd, epoch) == ST_NONE) {
                    QB_SPIN_YIELD();
                    wd = ld_word(src);
                }
                if (word_status(wd, epoch) == ST_AGG_EMPTY) {
                    if (--p >= 0) wd = ld_word(src - kEncDescWords);
                    continue;
                }
                e = (unsigned)word_payload(wd);
                break;
            }
            if (own[h] == sentinel(s)) st_word(desc + s, pack_word(e, ST_INCL, epoch));
            sm.gin[s] = e;
        }
        __syncwarp();
        QB_STAMP(desc, 67, 1, qb_t0);  // table / run look-back

        // ================= W3: the reference loop with a known table; chunks go to this lane's slice of the record =================
        // equal pixels before this chunk, mod 62 (the run counter on entry, simple.cpp:39-44)
        unsigned r = before ? (first - before) % kRunLimit : (first + base62 + kRunLimit - 1u) % kRunLimit;
        unsigned* const priv = P.scratch + (uint64_t)gt * C::kScrWords + lane;  // word j of this lane at priv[j * 32]
        unsigned        nw = 0, fill = 0, alo = 0, ahi = 0;  // whole words stored, bytes pending in alo (ahi: overflow of one append)
        {
            unsigned p = prev0;
#pragma unroll 1
            for (int g = 0; g < NG; ++g) {
                if (g + 1 < NG) ts_load_group<CH, G>(in_img, g0, (unsigned)g + 1u, nvalid, lastv, nx);
                const unsigned nm = neMask >> (g * G), em = eqMask >> (g * G);
                const bool     nexteq_g = g + 1 < NG ? ((em >> G) & 1u) != 0 : nexteq_last;  // for the group's last pixel
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const unsigned cur = px[i];
                    const bool     ne  = (nm >> i) & 1u, eq = (em >> i) & 1u;
                    const bool     nexteq = i < G - 1 ? ((em >> (i < G - 1 ? i + 1 : i)) & 1u) != 0 : nexteq_g;
                    // run pixel: one byte when the counter reaches 62 or the run ends here (simple.cpp:39-49, 91-94)
                    const unsigned r1   = r + 1u;
                    const bool     full = r1 == kRunLimit;
                    const bool     emit = eq && (full || !nexteq);
                    r                   = (eq && !full) ? r1 : 0u;
                    unsigned chunk, len;
                    ts_colour_chunk<CH>(cur, p, chunk, len);
                    // table probe and store (simple.cpp:51-57); an entry this tile has not stored yet holds the sentinel
                    const unsigned slot = slot_of(cur);
                    unsigned*      te   = sm.tab + slot * R + lane;
                    unsigned       tv   = *te;
                    if (tv == (slot == 0 ? 1u : 0u)) tv = sm.gin[slot];
                    if (ne) *te = cur;
                    const bool hit = tv == cur;
                    chunk = ne ? (hit ? (kOpIndex | slot) : chunk) : (emit ? (kOpRun - 1u) + r1 : 0u);  // util.hpp:190-235
                    len   = ne ? (hit ? 1u : len) : (emit ? 1u : 0u);
                    // append
                    const unsigned sh = fill * 8u;
                    alo |= chunk << sh;
                    ahi = __funnelshift_l(chunk, (CH == 4 && len == 5u) ? cur >> 24 : 0u, sh);
                    fill += len;
                    if (fill >= 4u) {
                        priv[nw * 32] = alo;
                        ++nw, alo = ahi, fill -= 4u;
                        if (CH == 4 && fill >= 4u) {  // a five-byte chunk behind three pending bytes fills two words
                            priv[nw * 32] = alo;
                            ++nw, alo = 0u, fill -= 4u;
                        }
                    }
                    p = cur;
                }
#pragma unroll
                for (int i = 0; i < G; ++i) px[i] = nx[i];
            }
            priv[nw * 32] = alo;  // the partial last word (its upper bytes are zero)
        }
        const unsigned total = nw * 4u + fill;
        QB_STAMP(desc, 68, 0, qb_t0);  // encode loop

        // ================= the tile's byte count =================
        priv[C::kPrivWords * 32] = total;
        unsigned tile_bytes = total;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) tile_bytes += __shfl_xor_sync(kFull, tile_bytes, d);
        if (lane == 0) {
            P.tile_bytes[gt] = tile_bytes;
            atomicAdd(P.group_bytes + (uint64_t)img * P.groups_per_image + (t >> 6), tile_bytes);
        }
        QB_STAMP(desc, 68, 1, qb_t0);  // counts
    }

    // ---- copy role: scratch record of tile `gt` -> the tile's bytes at their final place.  One warp.
    template <int CH>
    __device__ __forceinline__ void ts_copy_tile(const EncParams& P, TsCopySmem& sm, unsigned gt)
    {
        using C             = TsCfg<CH>;
        const unsigned lane = threadIdx.x & 31u;
        [[maybe_unused]] const long long qb_t0 = QB_T0();
        const unsigned  img     = gt / P.tiles_per_image;
        const unsigned  t       = gt % P.tiles_per_image;
        uint8_t*        out_img = P.out + (uint64_t)img * P.out_stride;
        [[maybe_unused]] uint64_t* desc = P.desc + (uint64_t)gt * kEncDescWords;
        const unsigned* priv    = P.scratch + (uint64_t)gt * C::kScrWords + lane;

        // where the tile's bytes start: header + the totals of the earlier 64-tile groups + the earlier tiles of this group
        const unsigned tile_bytes = P.tile_bytes[gt];
        uint64_t       tile_off   = 0;
        {
            const uint32_t* gb = P.group_bytes + (uint64_t)img * P.groups_per_image;
            const unsigned  g  = t >> 6;
            for (unsigned i = lane; i < g; i += 32) tile_off += gb[i];
            const unsigned j0 = gt - (t & 63u);  // first tile of the group
            for (unsigned j = j0 + lane; j < gt; j += 32) tile_off += P.tile_bytes[j];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) tile_off += __shfl_xor_sync(kFull, tile_off, d);
            tile_off += kHeader;
        }
        // per-lane byte counts -> offsets inside the tile
        const unsigned total = priv[C::kPrivWords * 32];
        unsigned       off;
        {
            unsigned inc = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(kFull, inc, d);
                if ((int)lane >= d) inc += o;
            }
            off = inc - total;
        }
        QB_STAMP(desc, 69, 0, qb_t0);  // copy: offsets

        // ================= compaction: the lanes' word-aligned slices -> the tile's contiguous bytes =================
        unsigned char* const stage = reinterpret_cast<unsigned char*>(sm.tab);
        {
            // destination word m (from the word holding my first byte) = my bytes 4m - a .. 4m - a + 3: slice words m - 1 and m
            // funnel-shifted; only the first and the last destination word can be shared with a neighbour (byte stores).
            const unsigned a = off & 3u, rs = 32u - a * 8u;
            unsigned*      d32 = reinterpret_cast<unsigned*>(stage) + (off >> 2);
            const unsigned nwp = (total + 3u) >> 2;                       // slice words holding bytes
            const unsigned nd  = total ? (total + a + 3u) >> 2 : 0u;      // destination words touched
            const unsigned maxnd = __reduce_max_sync(kFull, nd);          // the warp walks the longest slice, eight words a round
            unsigned       cw[8], nx[8], lo = 0, v_first = 0, v_last = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) cw[i] = (unsigned)i < nwp ? priv[i * 32] : 0u;
            for (unsigned m0 = 0; m0 < maxnd; m0 += 8u) {
                if (m0 + 8u < maxnd) {  // the next round's words are in flight during this one
#pragma unroll
                    for (int i = 0; i < 8; ++i) nx[i] = m0 + 8u + i < nwp ? priv[(m0 + 8u + i) * 32] : 0u;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned m = m0 + i;
                    const unsigned v = __funnelshift_rc(lo, cw[i], rs);
                    lo               = cw[i];
                    if (m < nd) {
                        const bool whole = (m > 0 || a == 0) && 4u * m + 4u <= total + a;
                        if (whole) d32[m] = v;
                        else if (m == 0) v_first = v;
                        else v_last = v;
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) cw[i] = nx[i];
            }
            auto partial = [&](unsigned m, unsigned v) {
                const unsigned b0 = m == 0 ? a : 0u, b1 = min(4u, total + a - 4u * m);
                unsigned char* d = reinterpret_cast<unsigned char*>(d32 + m);
#pragma unroll
                for (unsigned bb = 0; bb < 4; ++bb)
                    if (bb >= b0 && bb < b1) d[bb] = (unsigned char)(v >> (8u * bb));
            };
            if (total) {
                if (!(a == 0 && total >= 4u)) partial(0u, v_first);
                if (nd > 1 && ((total + a) & 3u) != 0) partial(nd - 1u, v_last);
            }
        }
        __syncwarp();
        QB_STAMP(desc, 69, 1, qb_t0);  // copy: compaction

        // ================= realigned 16-byte copy-out =================
        const unsigned tile_total = tile_bytes;
        if (tile_total) {
            uint8_t*        dst  = out_img + tile_off;
            const unsigned  head = min(tile_total, (16u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
            const unsigned  nv   = (tile_total - head) >> 4;
            if (lane < head) dst[lane] = stage[lane];
            // global chunk c is 16-byte aligned; its source starts at stage[head + 16c], any alignment mod 4
            const unsigned* s32 = reinterpret_cast<const unsigned*>(stage);
            const unsigned  sh8 = (head & 3u) * 8u, w0 = head >> 2;
            for (unsigned c = lane; c < nv; c += 32) {
                const unsigned* q = s32 + w0 + 4 * c;
                const unsigned  a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
                reinterpret_cast<uint4*>(dst + head)[c] = make_uint4(__funnelshift_r(a0, a1, sh8), __funnelshift_r(a1, a2, sh8),
                                                                     __funnelshift_r(a2, a3, sh8), __funnelshift_r(a3, a4, sh8));
            }
            const unsigned done = head + (nv << 4);
            if (lane < tile_total - done) dst[done + lane] = stage[done + lane];
        }
        if (t == 0 && lane < kHeader) out_img[lane] = P.header[lane];
        if (t == P.tiles_per_image - 1 && lane == 0) {  // end marker (util.hpp:151-161) and the result
            const uint64_t written = tile_off + tile_total;
            for (unsigned b = 0; b < kMarker; ++b) out_img[written + b] = b == kMarker - 1 ? 1 : 0;
            EncResult* res = P.results + img;
            res->written   = written + kMarker;
            res->complete  = 1;
            res->processed = P.n_pixels;
        }
        QB_STAMP(desc, 70, 0, qb_t0);  // copy: copy-out
    }

    // Encode kernel: persistent, independent warps draw tiles from a ticket counter in start order, so every tile a running
    // warp waits for (table / run look-back) is held by a warp that is running too or done.  The counter never resets: the
    // host passes its value at launch, every warp draws exactly one ticket beyond the last tile.
    template <int CH>
    __global__ void __launch_bounds__(kTsThreads, QB_TS_CTAS) encode_ts_kernel(const EncParams P)
    {
        TsWarpSmem&    sm      = reinterpret_cast<TsWarpSmem*>(QB_DYN_SMEM)[threadIdx.x >> 5];
        const unsigned lane    = threadIdx.x & 31u;
        const unsigned n_tiles = P.tiles_per_image * P.n_images;
        for (;;) {
            // the ticket is drawn when the tile starts, never earlier: a claimed tile that is not running yet would stall
            // the look-backs of all its successors (measured: drawing one tile ahead to prefetch its pixels cost 20 %)
            unsigned x = 0;
            if (lane == 0) x = atomicAdd(P.ticket, 1u) - P.ticket_base[0];
            x = __shfl_sync(kFull, x, 0);
            if (x >= n_tiles) break;
            ts_encode_tile<CH>(P, sm, x);
            __syncwarp();
        }
    }

    // Copy kernel: one warp per tile, launched behind the encode kernel on the same stream.
    template <int CH>
    __global__ void __launch_bounds__(kTsCopyWarps * 32, 5) encode_ts_copy_kernel(const EncParams P)
    {
        TsCopySmem&    sm = reinterpret_cast<TsCopySmem*>(QB_DYN_SMEM)[threadIdx.x >> 5];
        const unsigned gt = blockIdx.x * kTsCopyWarps + (threadIdx.x >> 5);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < P.zero_n; i += gridDim.x * blockDim.x) P.zero_ptr[i] = 0;
        if (gt < P.tiles_per_image * P.n_images) ts_copy_tile<CH>(P, sm, gt);
    }
}  // namespace qb
