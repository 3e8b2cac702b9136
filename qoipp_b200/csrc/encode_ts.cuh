// encode_ts.cuh -- thread-serial tile encoder: the one-shot fast path of the QOI encoder for sm_100a.
//
// Same result as encode_kernel (encode_kernel.cuh) and the reference loop impl::encode (source/simple.cpp:17-98),
// byte for byte; chosen by the host when the call is a plain one-shot encode (no carried state, capacity >= worst
// size, 16-byte aligned input).  encode_kernel gives every LANE one pixel per step and pays for it in shuffles,
// ballots and __match_any_sync (262 thread-instructions per pixel, issue bound).  Here every THREAD walks kTsK = 32
// consecutive pixels like the reference loop does -- previous pixel, run counter and output cursor live in
// registers -- and only the 64-entry "seen" table needs help, because a thread does not know the table at the start
// of its chunk:
//
//   W1     each thread records, per slot, the last pixel it would store (simple.cpp:57) in its own column of
//          tab[slot][thread] (conflict-free: bank = thread), other slots keep a sentinel;
//   merge  an exclusive "last writer wins" scan along the threads turns column t into the table on entry to thread
//          t's chunk; lane = slot, skewed by one thread per lane so that the 32 lanes of a step hit 16 banks twice;
//          two halves of 64 threads are scanned concurrently, the second half falls back on gin[1][slot];
//   carry  the tile's per-slot last writer / last differing pixel / byte count are published and looked back
//          exactly as in encode_kernel (same 72-word record per tile);
//   W2     chunk for "differs, no table hit" (DIFF / LUMA / RGB / RGBA) for all 32 pixels -- table independent, so
//          it runs between publishing the tile's words and waiting for the predecessors';
//   W3     the reference loop proper: probe + store per differing pixel, run counter, final chunk per pixel;
//   W3     the reference loop proper: probe + store per differing pixel, run counter, final chunk per pixel, appended
//          through a 64-bit register window to the thread's word-aligned slice of the tile's SCRATCH record in global
//          memory ([word][thread], coalesced); the tile's byte count is published and the CTA is done with the tile;
//   copy   `lag` tickets later some CTA (whose own encode work is finished) turns that record into the final bytes:
//          look back over the byte counts -- by then every predecessor has long published, nobody waits for a slow
//          neighbour --, prefix-sum the 128 per-thread counts, funnel-shift the slices into one contiguous run in
//          shared memory, realigned 16-byte copy-out.  Measured: with the byte carry inside the encode pass every tile
//          synchronised to the slowest of its 32 predecessors (6.2k of 33k cycles per tile waiting, 4 CTAs per SM).
#pragma once

#include "encode_kernel.cuh"

namespace qb
{
    #ifndef QB_TS_THREADS
#define QB_TS_THREADS 128
#endif
#ifndef QB_TS_CTAS
#define QB_TS_CTAS (768 / QB_TS_THREADS)
#endif
    constexpr int kTsThreads = QB_TS_THREADS, kTsWarps = kTsThreads / 32, kTsK = 32, kTsT = kTsThreads * kTsK;
    constexpr int kTsHalves = kTsThreads / 64;  // the merge scans ranges of 64 threads concurrently
    static_assert(kTsThreads == 64 || kTsThreads == 128 || kTsThreads == 256, "64 slots = warps x (32 / ranges) lanes");

    template <int CH>
    struct TsSmem {
        static constexpr int kPrivWords = kTsK * (CH + 1) / 4 + 1;  // per thread: its chunks, word aligned (+ the partial word)
        static constexpr int kScrWords  = (kPrivWords + 1) * kTsThreads;  // scratch record of a tile: [word][thread], then the counts
        alignas(16) unsigned tab[64 * kTsThreads];  // [slot][thread]; in the copy phase: the tile's staging bytes
        unsigned endv[kTsHalves][64];  // last writer per slot in each range of 64 threads (sentinel = none)
        unsigned gin[kTsHalves][64];   // table on entry to each range
        unsigned wlast[kTsWarps];  // per warp: tile-local index + 1 of its last differing pixel, 0 = none
        unsigned wbytes[kTsWarps];
        uint64_t tile_off;
        unsigned ticket, base62, tile_bytes;
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

    // ---- encode role: tile `gt` (global tile index) -> scratch record + carry words
    template <int CH>
    __device__ __forceinline__ void ts_encode_tile(const EncParams& P, TsSmem<CH>& sm, unsigned gt)
    {
        using S          = TsSmem<CH>;
        constexpr int K  = kTsK, NT = kTsThreads, T = kTsT;
        const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
        [[maybe_unused]] const long long qb_t0 = QB_T0();

        const unsigned  img        = gt / P.tiles_per_image;
        const unsigned  t          = gt % P.tiles_per_image;
        const uint8_t*  in_img     = P.in + (uint64_t)img * P.in_stride;
        const uint64_t  N          = P.n_pixels;
        const uint64_t  tile_start = (uint64_t)t * T;
        const unsigned  n_here     = (unsigned)(N - tile_start < (uint64_t)T ? N - tile_start : (uint64_t)T);
        uint64_t*       desc       = P.desc + ((uint64_t)img * P.tiles_per_image + t) * kEncDescWords;
        const unsigned  epoch      = P.epoch;
        const unsigned  first      = tid * K;  // tile-local index of this thread's first pixel
        const unsigned  nvalid     = first >= n_here ? 0u : min((unsigned)K, n_here - first);
        const uint64_t  g0         = tile_start + first;
        QB_STAMP(desc, 66, 0, qb_t0);  // ticket, table init
#if defined(QB_TIMING) && !defined(QB_EMU)
        if (tid == 0) { unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns)); reinterpret_cast<unsigned*>(desc + 71)[0] = (unsigned)ns; }
#endif

        // ---- this thread's 32 pixels
        unsigned px[K];
        if (nvalid == (unsigned)K) {
            if (CH == 4) {
                const uint4* src = reinterpret_cast<const uint4*>(in_img + g0 * 4);
#pragma unroll
                for (int j = 0; j < K / 4; ++j) {
                    const uint4 v = __ldg(src + j);
                    px[4 * j] = v.x, px[4 * j + 1] = v.y, px[4 * j + 2] = v.z, px[4 * j + 3] = v.w;
                }
            } else {
                const uint4* src = reinterpret_cast<const uint4*>(in_img + g0 * 3);
                unsigned     wd[3 * K / 4];
#pragma unroll
                for (int j = 0; j < 3 * K / 16; ++j) {
                    const uint4 v = __ldg(src + j);
                    wd[4 * j] = v.x, wd[4 * j + 1] = v.y, wd[4 * j + 2] = v.z, wd[4 * j + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < K / 4; ++j) {  // four pixels from three words; util.hpp:325: alpha forced to 255
                    const unsigned a = wd[3 * j], b = wd[3 * j + 1], c = wd[3 * j + 2];
                    px[4 * j]     = a | 0xFF000000u;
                    px[4 * j + 1] = __funnelshift_r(a, b, 24) | 0xFF000000u;
                    px[4 * j + 2] = __funnelshift_r(b, c, 16) | 0xFF000000u;
                    px[4 * j + 3] = (c >> 8) | 0xFF000000u;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) px[k] = (unsigned)k < nvalid ? load_pixel<CH>(in_img, g0 + k, true) : 0u;
        }
        // own table column: nothing stored yet (sentinel(s)); issued behind the pixel loads so that it overlaps their latency
#pragma unroll
        for (int s = 0; s < 64; ++s) sm.tab[s * NT + tid] = s == 0 ? 1u : 0u;
        // neighbours across the chunk boundary
        const unsigned up = __shfl_up_sync(kFull, px[K - 1], 1);
        const unsigned dn = __shfl_down_sync(kFull, px[0], 1);
        unsigned       prev0 = up, nxt = dn;
        if (lane == 0) prev0 = g0 == 0 ? kStartPixel : (g0 <= N ? load_pixel<CH>(in_img, g0 - 1, true) : 0u);
        if (lane == 31) nxt = g0 + K < N ? load_pixel<CH>(in_img, g0 + K, true) : 0u;
        if (nvalid != (unsigned)K) {  // pixels past the image repeat the last one: they never differ, `vmask` drops them
            unsigned lastv = prev0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if ((unsigned)k < nvalid) lastv = px[k];
                else px[k] = lastv;
            }
        }
        const bool nexteq_last = g0 + K < N && nxt == px[K - 1];  // does a run continue into the next thread's chunk?

        // ================= W1: per-slot last store of this chunk, mask of differing pixels =================
        unsigned neMask = 0;
        {
            unsigned p = prev0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned cur = px[k];
                if (cur != p) {
                    neMask |= 1u << k;
                    sm.tab[slot_of(cur) * NT + tid] = cur;  // simple.cpp:54-57: every differing pixel ends up in its slot
                }
                p = cur;
            }
        }
        const unsigned vmask  = nvalid == (unsigned)K ? 0xFFFFFFFFu : (1u << nvalid) - 1u;
        const unsigned eqMask = ~neMask & vmask;
        // last differing pixel before this thread's chunk (tile-local index + 1, 0 = none): exclusive max-scan
        unsigned before;
        {
            unsigned inc = neMask ? first + (31u - (unsigned)__clz((int)neMask)) + 1u : 0u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(kFull, inc, d);
                if ((int)lane >= d) inc = max(inc, o);
            }
            before = __shfl_up_sync(kFull, inc, 1);
            if (lane == 0) before = 0;
            if (lane == 31) sm.wlast[w] = inc;
        }
        __syncthreads();
        QB_STAMP(desc, 66, 1, qb_t0);  // pixel loads, W1

        // ================= merge: column t := table on entry to thread t's chunk (relative to its half) =================
        {
            constexpr int SPW = 64 / kTsWarps;  // slots per warp; the other lanes take the same slots in the next range
            const unsigned j = lane % SPW, h = lane / SPW, s = SPW * w + j, sent = sentinel(s);
            unsigned*      row = sm.tab + s * NT + 64u * h;
            unsigned       cur = sent;
#pragma unroll
            for (int k = 0; k < SPW - 1; ++k) {
                const int tr = k - (int)j;
                if (tr >= 0) {
                    const unsigned x = row[tr];
                    row[tr]          = cur;
                    if (x != sent) cur = x;
                }
            }
            for (int k0 = SPW - 1; k0 < 64; k0 += 8) {  // all lanes in range: loads of a batch first, they are independent
                unsigned x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (k0 + i < 64) x[i] = row[k0 + i - (int)j];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (k0 + i < 64) {
                        row[k0 + i - (int)j] = cur;
                        if (x[i] != sent) cur = x[i];
                    }
            }
#pragma unroll
            for (int k = 64; k < 64 + SPW - 1; ++k) {
                const int tr = k - (int)j;
                if (tr < 64) {
                    const unsigned x = row[tr];
                    row[tr]          = cur;
                    if (x != sent) cur = x;
                }
            }
            sm.endv[h][s] = cur;
        }
        __syncthreads();
        QB_STAMP(desc, 67, 0, qb_t0);  // merge

        // ================= carries (1) table and (2) run: publish the tile's words, look back =================
        if (tid < 64) {
            const unsigned s = tid, sent = sentinel(s);
            unsigned       own = sent;
#pragma unroll
            for (int h = kTsHalves - 1; h >= 0; --h)
                if (own == sent) own = sm.endv[h][s];
            const bool present = own != sent;
            st_word(desc + s, pack_word(own, present ? ST_INCL : ST_AGG_EMPTY, epoch));
            unsigned e;
            int      p = (int)t - 1;
            for (;;) {
                if (p < 0) { e = 0u; break; }  // simple.cpp:28: zero-initialised table
                const uint64_t wd = wait_word(desc - (int64_t)(t - p) * kEncDescWords + s, epoch);
                if (word_status(wd, epoch) == ST_AGG_EMPTY) { --p; continue; }
                e = (unsigned)word_payload(wd);
                break;
            }
            if (!present) st_word(desc + s, pack_word(e, ST_INCL, epoch));
#pragma unroll
            for (int h = 0; h < kTsHalves; ++h) {
                sm.gin[h][s] = e;
                const unsigned v = sm.endv[h][s];
                if (v != sent) e = v;
            }
        }
        if (w == kTsWarps - 1) {
            unsigned m = 0;
#pragma unroll
            for (int ww = 0; ww < kTsWarps; ++ww) m = max(m, sm.wlast[ww]);
            const int tl = (int)m - 1;
            if (lane == 0) {
                if (tl >= 0) st_word(desc + kWordLne, pack_word(tile_start + (unsigned)tl + kLneBias, ST_INCL, epoch));
                else st_word(desc + kWordLne, pack_word(0, ST_AGG_EMPTY, epoch));
            }
            // payload = last differing pixel index + bias (never 0); the later tile wins when it has one
            const uint64_t lp = warp_lookback_lazy<uint64_t>(
                t, (uint64_t)(kLneBias - 1), (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(desc - (int64_t)(t - p) * kEncDescWords + kWordLne);
                    st                = word_status(wd, epoch);
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return b ? b : a; });
            if (lane == 0) {
                if (tl < 0) st_word(desc + kWordLne, pack_word(lp, ST_INCL, epoch));
                sm.base62 = (unsigned)((tile_start + kLneBias - lp) % kRunLimit);  // (tile_start - last differing) mod 62
            }
        }
        __syncthreads();
        QB_STAMP(desc, 67, 1, qb_t0);  // table / run look-back

        // ================= the reference loop with a known table; chunks go to this thread's private words =================
        // equal pixels before this chunk, mod 62 (the run counter on entry, simple.cpp:39-44)
        unsigned r;
        {
            unsigned bf = before;
            for (unsigned ww = 0; ww < w; ++ww) bf = max(bf, sm.wlast[ww]);
            r = bf ? (first - bf) % kRunLimit : (first + sm.base62 + kRunLimit - 1u) % kRunLimit;
        }
        unsigned* const scr  = P.scratch + (uint64_t)gt * S::kScrWords;
        unsigned* const priv = scr + tid;  // word j of this thread at priv[j * NT]: a warp stores whole 128-byte lines
        unsigned        nw = 0, fill = 0, alo = 0, ahi = 0;  // whole words stored, bytes pending in alo (ahi: overflow of one append)
        {
            const unsigned* gin = sm.gin[tid >> 6];
            unsigned        p   = prev0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned cur = px[k];
                const bool     ne  = (neMask >> k) & 1u, eq = (eqMask >> k) & 1u;
                const bool     nexteq = k < K - 1 ? ((eqMask >> (k < K - 1 ? k + 1 : k)) & 1u) != 0 : nexteq_last;
                // run pixel: one byte when the counter reaches 62 or the run ends here (simple.cpp:39-49, 91-94)
                const unsigned r1   = r + 1u;
                const bool     full = r1 == kRunLimit;
                const bool     emit = eq && (full || !nexteq);
                r                   = (eq && !full) ? r1 : 0u;
                unsigned chunk, len;
                ts_colour_chunk<CH>(cur, p, chunk, len);
                // table probe and store (simple.cpp:51-57); an entry this half has not stored yet holds the sentinel
                const unsigned slot = slot_of(cur);
                unsigned*      te   = sm.tab + slot * NT + tid;
                unsigned       tv   = *te;
                if (tv == (slot == 0 ? 1u : 0u)) tv = gin[slot];
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
                    priv[nw * NT] = alo;
                    ++nw, alo = ahi, fill -= 4u;
                    if (CH == 4 && fill >= 4u) {  // a five-byte chunk behind three pending bytes fills two words
                        priv[nw * NT] = alo;
                        ++nw, alo = 0u, fill -= 4u;
                    }
                }
                p = cur;
            }
            priv[nw * NT] = alo;  // the partial last word (its upper bytes are zero)
        }
        const unsigned total = nw * 4u + fill;
        QB_STAMP(desc, 68, 0, qb_t0);  // encode loop

        // ================= carry (3): the tile's byte count; the record is complete =================
        priv[S::kPrivWords * NT] = total;
        {
            unsigned sum = total;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(kFull, sum, d);
            if (lane == 0) sm.wbytes[w] = sum;
        }
        __syncthreads();  // every thread's scratch stores are issued
        if (tid == 0) {
            unsigned tile_bytes = 0;
#pragma unroll
            for (int ww = 0; ww < kTsWarps; ++ww) tile_bytes += sm.wbytes[ww];
            __threadfence();  // the record before the word that announces it
            st_word(desc + kWordBytes, pack_word(tile_bytes, ST_AGG, epoch));
        }
        QB_STAMP(desc, 68, 1, qb_t0);  // publish
#if defined(QB_TIMING) && !defined(QB_EMU)
        if (tid == 0) { unsigned long long ns; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns)); reinterpret_cast<unsigned*>(desc + 71)[1] = (unsigned)ns; }
#endif
    }

    // ---- copy role: scratch record of tile `gt` -> the tile's bytes at their final place
    template <int CH>
    __device__ __forceinline__ void ts_copy_tile(const EncParams& P, TsSmem<CH>& sm, unsigned gt)
    {
        using S          = TsSmem<CH>;
        constexpr int NT = kTsThreads;
        const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
        [[maybe_unused]] const long long qb_t0 = QB_T0();
        const unsigned  img     = gt / P.tiles_per_image;
        const unsigned  t       = gt % P.tiles_per_image;
        uint8_t*        out_img = P.out + (uint64_t)img * P.out_stride;
        uint64_t*       desc    = P.desc + (uint64_t)gt * kEncDescWords;
        const unsigned  epoch   = P.epoch;
        const unsigned* scr     = P.scratch + (uint64_t)gt * S::kScrWords;
        const unsigned* priv    = scr + tid;

        if (tid == 0) {  // the encoder of this tile holds a lower ticket: it is running or done
            sm.tile_bytes = (unsigned)word_payload(wait_word(desc + kWordBytes, epoch));
            __threadfence();  // the word before the record it announces
        }
        __syncthreads();
        const unsigned tile_bytes = sm.tile_bytes;
        QB_STAMP(desc, 69, 1, qb_t0);  // copy: waited for the record
        if (w == 0) {  // where the tile's bytes start: look back over the byte counts, 32 predecessors per round
            const uint64_t toff = warp_lookback_lazy<uint64_t>(
                t, (uint64_t)kHeader, (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(desc - (int64_t)(t - p) * kEncDescWords + kWordBytes);
                    st                = word_status(wd, epoch);
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return a + b; });
            if (lane == 0) {
                st_word(desc + kWordBytes, pack_word(toff + tile_bytes, ST_INCL, epoch));
                sm.tile_off = toff;
            }
        }
        // per-thread byte counts -> offsets inside the tile
        const unsigned total = __ldcg(priv + S::kPrivWords * NT);
        unsigned       off;
        {
            unsigned inc = total;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned o = __shfl_up_sync(kFull, inc, d);
                if ((int)lane >= d) inc += o;
            }
            off = inc - total;
            if (lane == 31) sm.wbytes[w] = inc;
        }
        __syncthreads();
        for (unsigned ww = 0; ww < w; ++ww) off += sm.wbytes[ww];
        QB_STAMP(desc, 70, 0, qb_t0);  // copy: look-back (warp 0), counts, scan

        // ================= compaction: the threads' word-aligned slices -> the tile's contiguous bytes =================
        unsigned char* const stage = reinterpret_cast<unsigned char*>(sm.tab);
        if (total) {
            // destination word m (from the word holding my first byte) = my bytes 4m - a .. 4m - a + 3: slice words m - 1 and m
            // funnel-shifted; only the first and the last destination word can be shared with a neighbour (byte stores)
            const unsigned a = off & 3u, rs = 32u - a * 8u;
            unsigned*      d32 = reinterpret_cast<unsigned*>(stage) + (off >> 2);
            const unsigned nwp = (total + 3u) >> 2;      // slice words holding bytes
            const unsigned nd  = (total + a + 3u) >> 2;  // destination words touched
            auto partial = [&](unsigned m, unsigned v) {
                const unsigned b0 = m == 0 ? a : 0u, b1 = min(4u, total + a - 4u * m);
                unsigned char* d = reinterpret_cast<unsigned char*>(d32 + m);
#pragma unroll
                for (unsigned bb = 0; bb < 4; ++bb)
                    if (bb >= b0 && bb < b1) d[bb] = (unsigned char)(v >> (8u * bb));
            };
            unsigned lo = __ldcg(priv);
            {
                const unsigned v = __funnelshift_rc(0u, lo, rs);
                if (a == 0 && total >= 4u) d32[0] = v;
                else partial(0u, v);
            }
            unsigned m = 1;
            for (; m + 4u < nd; m += 4u) {  // four whole words per round, loads first
                unsigned h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = __ldcg(priv + (m + i) * NT);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    d32[m + i] = __funnelshift_rc(lo, h[i], rs);
                    lo         = h[i];
                }
            }
            for (; m < nd; ++m) {
                const unsigned hi = m < nwp ? __ldcg(priv + m * NT) : 0u;
                const unsigned v  = __funnelshift_rc(lo, hi, rs);
                lo                = hi;
                if (m + 1u < nd || ((total + a) & 3u) == 0) d32[m] = v;
                else partial(m, v);
            }
        }
        __syncthreads();

        QB_STAMP(desc, 70, 1, qb_t0);  // copy: compaction
        // ================= realigned 16-byte copy-out =================
        const uint64_t tile_off   = sm.tile_off;
        const unsigned tile_total = tile_bytes;
        if (tile_total) {
            uint8_t*        dst  = out_img + tile_off;
            const unsigned  head = min(tile_total, (16u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
            const unsigned  nv   = (tile_total - head) >> 4;
            if (tid < head) dst[tid] = stage[tid];
            // global chunk c is 16-byte aligned; its source starts at stage[head + 16c], any alignment mod 4
            const unsigned* s32 = reinterpret_cast<const unsigned*>(stage);
            const unsigned  sh8 = (head & 3u) * 8u, w0 = head >> 2;
            for (unsigned c = tid; c < nv; c += NT) {
                const unsigned* q = s32 + w0 + 4 * c;
                const unsigned  a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
                reinterpret_cast<uint4*>(dst + head)[c] = make_uint4(__funnelshift_r(a0, a1, sh8), __funnelshift_r(a1, a2, sh8),
                                                                     __funnelshift_r(a2, a3, sh8), __funnelshift_r(a3, a4, sh8));
            }
            const unsigned done = head + (nv << 4);
            if (tid < tile_total - done) dst[done + tid] = stage[done + tid];
        }
        if (t == 0 && tid < kHeader) out_img[tid] = P.header[tid];
        if (t == P.tiles_per_image - 1 && tid == 0) {  // end marker (util.hpp:151-161) and the result
            const uint64_t written = tile_off + tile_total;
            for (unsigned b = 0; b < kMarker; ++b) out_img[written + b] = b == kMarker - 1 ? 1 : 0;
            EncResult* res = P.results + img;
            res->written   = written + kMarker;
            res->complete  = 1;
            res->processed = P.n_pixels;
        }
        QB_STAMP(desc, 69, 0, qb_t0);  // copy role
    }

    // Grid = tiles + lag CTAs.  Ticket x (handed out in start order) encodes tile x (x < tiles).  A CTA that is done encoding --
    // or has nothing to encode -- takes a copy ticket c (handed out in FINISHING order) and copies tile c - lag: at least c
    // CTAs finished before it, so that tile's encoder started long ago and, `lag` finishers later, is done in practice; the
    // copy role therefore finds every word it looks back on already published, and no tile waits for a slow neighbour.
    // Every CTA a running one can wait for has started (tickets are issued in order), so the waits cannot deadlock.
    template <int CH>
    __global__ void __launch_bounds__(kTsThreads, QB_TS_CTAS) encode_ts_kernel(const EncParams P)
    {
        TsSmem<CH>& sm = *reinterpret_cast<TsSmem<CH>*>(QB_DYN_SMEM);
        const unsigned n_tiles = P.tiles_per_image * P.n_images;
        if (threadIdx.x == 0) sm.ticket = atomicInc(P.ticket, n_tiles + P.lag - 1u);
        __syncthreads();
        const unsigned x = sm.ticket;
        if (x < n_tiles) ts_encode_tile<CH>(P, sm, x);
        __syncthreads();  // the table's memory becomes the staging bytes
        if (threadIdx.x == 0) sm.ticket = atomicInc(P.ticket + 1, n_tiles + P.lag - 1u);
        __syncthreads();
        const unsigned c = sm.ticket;
        if (c >= P.lag) ts_copy_tile<CH>(P, sm, c - P.lag);
    }
}  // namespace qb
