// encode_kernel.cuh -- single-pass, tile-parallel QOI encoder for sm_100a.
//
// Replaces the serial loop of the reference, impl::encode (source/simple.cpp:17-98) and its resumable twin
// StreamEncoder::encode (source/stream.cpp:138-239), byte for byte.  The per-pixel recurrences of the
// reference are restated so that every tile (T pixels, one CTA) needs only three small carries:
//
//   (1) table carry   -- the 64-entry "seen" array is written by EVERY pixel that differs from its
//                        predecessor (index hit or not, simple.cpp:54-57), so "index hit at pixel i" ==
//                        "the last earlier differing pixel with the same slot equals p[i]".  Per tile that is a
//                        right-biased 64-slot override map; across tiles a decoupled look-back per slot.
//   (2) run carry     -- index of the last differing pixel (a max-scan); run bytes follow from
//                        runpos = i - last_differing: one byte when runpos % 62 == 0 or the run ends at i.
//   (3) offset carry  -- prefix sum of the per-pixel chunk lengths {0,1,2,4,5}.
//
// (1) and (2) depend on no other carry, (3) needs both, so a tile does: classify -> publish (1),(2) -> look back
// -> fix the <=64 first-in-warp slot probes and the run bytes -> publish (3) -> look back -> stage bytes in
// shared memory -> 16-byte coalesced stores.  Every chunk is attributed to exactly one pixel, so the
// "first chunk that does not fit ends the output" rule of util::ChunkArray<Checked> (util.hpp:240-246)
// becomes a comparison of chunk end offsets with the capacity.
#pragma once

#include "qb_common.cuh"

namespace qb
{
    struct EncState {  // == StreamEncoder members (include/qoipp/stream.hpp:112-115), pixels packed r|g<<8|b<<16|a<<24
        uint32_t prev;
        uint32_t run;
        uint32_t table[64];
    };

    struct EncResult {
        uint64_t written;    // bytes stored to the output of this image / call
        uint64_t processed;  // pixels consumed (stream mode; one-shot: pixels whose chunks were stored)
        uint32_t complete;   // one-shot: EncodeStatus::complete; stream: 1 when every input pixel was consumed
        uint32_t pad;
        EncState state;      // stream mode: carry-out after `processed` pixels
    };

    enum : uint32_t { ENC_STREAM = 1u };

    struct EncParams {
        const uint8_t*  in;
        uint8_t*        out;
        uint64_t        n_pixels;  // per image
        uint64_t        in_stride, out_stride, out_cap;
        uint32_t        tiles_per_image, n_images, epoch, flags;
        uint8_t         header[16];  // one-shot: the 14 header bytes (util.hpp:125-149)
        const EncState* init_state;  // stream mode carry-in (device), may be null
        EncResult*      results;     // [n_images]
        uint64_t*       desc;        // [n_images * tiles_per_image][kEncDescWords]
        uint32_t*       ticket;
        uint32_t*       scratch;     // encode_ts_kernel: per-tile records, read by encode_ts_copy_kernel
        uint32_t*       tile_bytes;  // encode_ts_kernel: [tiles] byte count of every tile
        uint32_t*       group_bytes; // encode_ts_kernel: [n_images][groups_per_image] totals of 64-tile groups (zero at launch)
        uint32_t*       zero_ptr;    // encode_ts_copy_kernel clears these zero_n words: the group totals of the PREVIOUS encode, which the
        uint32_t        zero_n;      // next one will use (two buffers take turns, so no memset is launched between encodes)
        uint32_t        groups_per_image;
        uint32_t        ticket_base[1];  // encode_ts_kernel: value of *ticket at launch (never reset)
    };

    constexpr int kEncWarps = 8, kEncThreads = kEncWarps * 32;
    constexpr int kEncDescWords = 72;  // 64 table words, [64] run carry, [65] byte offset carry, padding to 576 B
    constexpr int kWordLne = 64, kWordBytes = 65;
    constexpr uint64_t kLneBias = 64;  // run carry payload = last differing pixel index + 64 (>= 1: run-in <= 62)

    template <int K>
    struct EncSmem {
        static constexpr int T = kEncThreads * K;
        static constexpr int kStage = ((T * 5 + 96) + 127) / 128 * 128;
        alignas(128) unsigned char stage[kStage];  // RGB input staging, later the tile's chunk bytes (tile-relative)
        unsigned wtab[kEncWarps * 64];             // per warp: last differing pixel value per slot (sentinel = none)
        unsigned inw[kEncWarps * 64];              // per warp: table state on entry to the warp's chunk
        unsigned excl[64];                         // table state on entry to the tile
        unsigned incl[64];                         // table state on exit from the tile
        int      wlne[kEncWarps];                  // tile-local index of the warp's last differing pixel, -1 none
        unsigned wbytes[kEncWarps];
        uint64_t tile_off;
        uint64_t lne_payload;  // inclusive run carry on entry (biased)
        unsigned ticket, base62, tile_total, pre, cut_rel, cut_px, cut_q, last_q, first_ne;
    };

    template <int CH>
    __device__ __forceinline__ unsigned load_pixel(const uint8_t* img, uint64_t g, bool aligned4)
    {
        if (CH == 4) {
            if (aligned4) return __ldg(reinterpret_cast<const unsigned*>(img) + g);
            const uint8_t* p = img + g * 4;
            return p[0] | (unsigned)p[1] << 8 | (unsigned)p[2] << 16 | (unsigned)p[3] << 24;
        }
        const uint8_t* p = img + g * 3;
        return p[0] | (unsigned)p[1] << 8 | (unsigned)p[2] << 16 | 0xFF000000u;  // util.hpp:325: alpha forced to 255
    }

    // chunk for a pixel that differs from its predecessor and missed the table: simple.cpp:59-79, util.hpp:163-225.
    // Returns the first four chunk bytes; the fifth byte of an RGBA chunk is the pixel's alpha.  len in `len`.
    template <int CH>
    __device__ __forceinline__ unsigned colour_chunk(unsigned cur, unsigned prv, unsigned& len)
    {
        const unsigned d = sub4(cur, prv);  // wrapping i8 deltas, simple.cpp:66-71
        if (CH == 4 && (d >> 24) != 0) { len = 5; return kOpRgba | cur << 8; }
        const unsigned t = add4(d, 0x00020202u);  // bias_op_diff
        if ((t & 0x00FCFCFCu) == 0) {             // util.hpp:102-107
            len = 1;
            return kOpDiff | (t & 3u) << 4 | ((t >> 8) & 3u) << 2 | ((t >> 16) & 3u);
        }
        const unsigned dr = d & 255u, dg = (d >> 8) & 255u, db = (d >> 16) & 255u;
        const unsigned vg = (dg + 32u) & 255u, vr = (dr - dg + 8u) & 255u, vb = (db - dg + 8u) & 255u;
        if (((vg >> 6) | (vr >> 4) | (vb >> 4)) == 0) {  // util.hpp:109-114
            len = 2;
            return kOpLuma | vg | (vr << 4 | vb) << 8;
        }
        len = 4;
        return kOpRgb | cur << 8;
    }

    template <int CH, int K>
    __global__ void __launch_bounds__(kEncThreads, 4) encode_kernel(const EncParams P)
    {
        using S            = EncSmem<K>;
        constexpr int T    = S::T;
        static_assert(K <= 8, "per-lane masks hold 4 bits per step");
        S&            sm   = *reinterpret_cast<S*>(QB_DYN_SMEM);
        const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
        const bool     stream = (P.flags & ENC_STREAM) != 0;
        [[maybe_unused]] const long long qb_t0 = QB_T0();

        // ---- tile ticket: ids are handed out in start order, so every predecessor is running or done
        if (tid == 0) {
            sm.ticket   = atomicInc(P.ticket, P.tiles_per_image * P.n_images - 1u);
            sm.cut_rel  = 0;
            sm.cut_px   = 0xffffffffu;
            sm.last_q   = 0;
            sm.first_ne = 0;
        }
        __syncthreads();
        const unsigned  img        = sm.ticket / P.tiles_per_image;
        const unsigned  t          = sm.ticket % P.tiles_per_image;
        const uint8_t*  in_img     = P.in + (uint64_t)img * P.in_stride;
        uint8_t*        out_img    = P.out + (uint64_t)img * P.out_stride;
        const uint64_t  N          = P.n_pixels;
        const uint64_t  tile_start = (uint64_t)t * T;
        const unsigned  n_here     = (unsigned)(N - tile_start < (uint64_t)T ? N - tile_start : (uint64_t)T);
        uint64_t*       desc       = P.desc + ((uint64_t)img * P.tiles_per_image + t) * kEncDescWords;
        const unsigned  epoch      = P.epoch;
        const bool      aligned4   = (reinterpret_cast<uintptr_t>(in_img) & 3u) == 0;
        const EncState* init       = stream ? P.init_state : nullptr;
        const unsigned  run_in     = init ? init->run : 0u;
        const unsigned  chunk0     = w * (K * 32);  // tile-local index of this warp's first pixel
        QB_STAMP(desc, 66, 0, qb_t0);  // after the ticket

        // ---- pixels of this warp's chunk into registers (striped: step k holds 32 consecutive pixels)
        unsigned px[K];
        if (CH == 3) {
            const uint8_t* src    = in_img + tile_start * 3;
            const unsigned nbytes = n_here * 3u;
            if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
                const unsigned nv = nbytes >> 4;
                for (unsigned c = tid; c < nv; c += kEncThreads)
                    reinterpret_cast<uint4*>(sm.stage)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
                for (unsigned b = (nv << 4) + tid; b < nbytes; b += kEncThreads) sm.stage[b] = src[b];
            } else {
                for (unsigned b = tid; b < nbytes; b += kEncThreads) sm.stage[b] = src[b];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned i = chunk0 + k * 32 + lane;
                const unsigned char* p = sm.stage + i * 3u;
                px[k] = i < n_here ? (p[0] | (unsigned)p[1] << 8 | (unsigned)p[2] << 16 | 0xFF000000u) : 0u;
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned i = chunk0 + k * 32 + lane;
                px[k] = i < n_here ? load_pixel<CH>(in_img, tile_start + i, aligned4) : 0u;
            }
        }
        // neighbours across the chunk boundary
        unsigned chunk_prev = kStartPixel;
        if (chunk0 < n_here) {
            if (tile_start + chunk0 == 0) chunk_prev = init ? init->prev : kStartPixel;
            else chunk_prev = load_pixel<CH>(in_img, tile_start + chunk0 - 1, aligned4);
        }
        const uint64_t g_next     = tile_start + chunk0 + K * 32;
        const unsigned chunk_next = g_next < N ? load_pixel<CH>(in_img, g_next, aligned4) : 0u;

        QB_STAMP(desc, 69, 0, qb_t0);  // after pixel loads issued
        // ================= phase A: classify every pixel =================
        // per-lane state of the K pixels it owns: first four chunk bytes, 4-bit length fields, 1-bit-per-step masks
        unsigned lo[K];
        unsigned lens = 0;
        unsigned unresmask = 0, leadmask = 0, nextmask = 0;
        unsigned probemask = 0, lastmask = 0;  // differing pixel with no in-step predecessor / last of its slot in the step
        sm.wtab[w * 64 + lane]      = sentinel(lane);
        sm.wtab[w * 64 + 32 + lane] = 0u;

        // A1: everything that does not touch the warp table -- independent across steps, so the loads, shuffles and
        // __match_any_sync of different steps overlap
        int wl = -1;  // chunk-local index of the last differing pixel so far (warp uniform)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned ic    = k * 32 + lane;  // chunk-local
            const unsigned i     = chunk0 + ic;
            const bool     valid = i < n_here;
            const unsigned cur   = px[k];
            const unsigned up    = __shfl_up_sync(kFull, cur, 1);
            const unsigned carry = k == 0 ? chunk_prev : __shfl_sync(kFull, px[k > 0 ? k - 1 : 0], 31);
            const unsigned prv   = lane ? up : carry;
            const unsigned dn    = __shfl_down_sync(kFull, cur, 1);
            const unsigned cnext = k == K - 1 ? chunk_next : __shfl_sync(kFull, px[k < K - 1 ? k + 1 : k], 0);
            const unsigned nxt   = lane == 31 ? cnext : dn;
            // does the run continue into the next pixel?  Past the image end: one-shot flushes (simple.cpp:91-94),
            // the resumable form keeps the run pending (stream.cpp:158-169)
            const bool nexteq = tile_start + i + 1 < N ? nxt == cur : stream;
            const bool eq     = valid && cur == prv;
            const bool ne     = valid && !eq;
            const unsigned slot  = slot_of(cur);
            const unsigned m     = __match_any_sync(kFull, ne ? slot : 64u + lane);
            const unsigned below = m & lanemask_lt(lane);
            const unsigned pl    = below ? 31u - __clz(below) : lane;
            const unsigned pv    = __shfl_sync(kFull, cur, pl);
            const unsigned bne   = __ballot_sync(kFull, ne);
            unsigned len = 0, bytes = 0;
            if (ne) {
                if (below && pv == cur) bytes = kOpIndex | slot, len = 1;  // hit on a pixel of the same step
                else bytes = colour_chunk<CH>(cur, prv, len);              // provisional when the table is still to be probed
                if (!below) probemask |= 1u << k;
                if ((m & lanemask_gt(lane)) == 0) lastmask |= 1u << k;
            } else if (eq) {
                // run pixel: position in its run = distance to the last differing pixel (simple.cpp:39-49)
                const unsigned bl = bne & lanemask_lt(lane);
                if (bl || wl >= 0) {
                    const unsigned q = bl ? lane - (31u - __clz(bl)) : (ic - (unsigned)wl) % kRunLimit;
                    if (q == 0 || !nexteq) bytes = kOpRun | ((q + kRunLimit - 1) % kRunLimit), len = 1;  // util.hpp:227-235
                    if (tile_start + i + 1 == N) sm.last_q = q;
                } else {
                    leadmask |= 1u << k;  // the run started before this warp's chunk: needs carry (2)
                }
            }
            if (bne) wl = k * 32 + 31 - __clz(bne);
            lo[k] = bytes;
            lens |= len << (4 * k);
            nextmask |= (unsigned)nexteq << k;
            if (i == 0 && t == 0) sm.first_ne = ne;
        }
        // A2: the serial part -- probe and update the warp's 64-entry table step by step (simple.cpp:51-57)
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned cur = px[k], slot = slot_of(cur);
            if ((probemask >> k) & 1u) {
                const unsigned tv = sm.wtab[w * 64 + slot];
                if (tv == sentinel(slot)) unresmask |= 1u << k;  // first pixel of this slot in the warp's chunk
                else if (tv == cur) lo[k] = kOpIndex | slot, lens = (lens & ~(15u << (4 * k))) | 1u << (4 * k);
            }
            __syncwarp();
            if ((lastmask >> k) & 1u) sm.wtab[w * 64 + slot] = cur;
            __syncwarp();
        }
        if (lane == 0) sm.wlne[w] = wl < 0 ? -1 : (int)chunk0 + wl;
        __syncthreads();
        QB_STAMP(desc, 66, 1, qb_t0);  // after phase A

        // ================= carries (1) and (2): publish, look back =================
        if (tid < 64) {
            const unsigned s   = tid;
            unsigned       own = sentinel(s);
            for (int ww = kEncWarps - 1; ww >= 0; --ww) {
                const unsigned v = sm.wtab[ww * 64 + s];
                if (v != sentinel(s)) { own = v; break; }
            }
            const bool present = own != sentinel(s);
            st_word(desc + s, pack_word(own, present ? ST_INCL : ST_AGG_EMPTY, epoch));
            unsigned e;
            int      p = (int)t - 1;
            for (;;) {
                if (p < 0) { e = init ? init->table[s] : 0u; break; }  // simple.cpp:28: zero-initialised table
                const uint64_t wd = wait_word(desc - (int64_t)(t - p) * kEncDescWords + s, epoch);
                if (word_status(wd, epoch) == ST_AGG_EMPTY) { --p; continue; }
                e = (unsigned)word_payload(wd);
                break;
            }
            if (!present) st_word(desc + s, pack_word(e, ST_INCL, epoch));
            sm.excl[s] = e;
            sm.incl[s] = present ? own : e;
        } else if (w == 2) {  // run carry: one warp inspects 32 predecessors per round
            int tl = -1;
            for (int ww = kEncWarps - 1; ww >= 0; --ww)
                if (sm.wlne[ww] >= 0) { tl = sm.wlne[ww]; break; }
            if (lane == 0) {
                if (tl >= 0) st_word(desc + kWordLne, pack_word(tile_start + (unsigned)tl + kLneBias, ST_INCL, epoch));
                else st_word(desc + kWordLne, pack_word(0, ST_AGG_EMPTY, epoch));
            }
            // payload = last differing pixel index + bias (never 0); the later tile wins when it has one
            const uint64_t lp = warp_lookback<uint64_t>(
                t, (uint64_t)(kLneBias - 1 - run_in), (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = wait_word(desc - (int64_t)(t - p) * kEncDescWords + kWordLne, epoch);
                    st                = word_status(wd, epoch);
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return b ? b : a; });
            if (lane == 0) {
                if (tl < 0) st_word(desc + kWordLne, pack_word(lp, ST_INCL, epoch));
                sm.lne_payload = lp;
                sm.base62      = (unsigned)((tile_start + kLneBias - lp) % kRunLimit);  // (tile_start - last_differing) mod 62
            }
        }
        __syncthreads();
        QB_STAMP(desc, 67, 0, qb_t0);  // after table / run look-back

        // ================= fix-ups: table probes that left the warp, runs that entered the warp =================
        {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const unsigned s = lane + 32 * h;
                unsigned       v = sm.excl[s];
                for (int ww = (int)w - 1; ww >= 0; --ww) {
                    const unsigned x = sm.wtab[ww * 64 + s];
                    if (x != sentinel(s)) { v = x; break; }
                }
                sm.inw[w * 64 + s] = v;
            }
            __syncwarp();
        }
        if (unresmask) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if ((unresmask >> k) & 1u) {
                    const unsigned cur = px[k], slot = slot_of(cur);
                    if (sm.inw[w * 64 + slot] == cur) {
                        lo[k] = kOpIndex | slot;
                        lens  = (lens & ~(15u << (4 * k))) | 1u << (4 * k);
                    }
                }
            }
        }
        if (__any_sync(kFull, leadmask != 0)) {
            // run position of tile-local pixel i is congruent to i + o (mod 62): o = -(last differing pixel) or the carry
            int o = (int)sm.base62;
            for (int ww = (int)w - 1; ww >= 0; --ww)
                if (sm.wlne[ww] >= 0) { o = -sm.wlne[ww]; break; }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if ((leadmask >> k) & 1u) {
                    const unsigned i = chunk0 + k * 32 + lane;
                    const unsigned q = (unsigned)((int)i + o) % kRunLimit;
                    if (q == 0 || !((nextmask >> k) & 1u)) {
                        lo[k] = kOpRun | ((q + kRunLimit - 1) % kRunLimit);
                        lens |= 1u << (4 * k);
                    }
                    if (tile_start + i + 1 == N) sm.last_q = q;
                }
            }
        }

        // ================= carry (3): chunk offsets (lengths are 0,1,2,4,5: three ballots give the prefix) =================
        unsigned offs[(K + 1) / 2];  // two 16-bit chunk-relative offsets per register
#pragma unroll
        for (int j = 0; j < (K + 1) / 2; ++j) offs[j] = 0;
        unsigned running = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned len = (lens >> (4 * k)) & 15u;
            const unsigned b1 = __ballot_sync(kFull, len & 1u), b2 = __ballot_sync(kFull, len & 2u), b4 = __ballot_sync(kFull, len & 4u);
            const unsigned lt = lanemask_lt(lane);
            const unsigned e  = __popc(b1 & lt) + 2u * __popc(b2 & lt) + 4u * __popc(b4 & lt);
            offs[k >> 1] |= (running + e) << (16 * (k & 1));
            running += __popc(b1) + 2u * __popc(b2) + 4u * __popc(b4);
        }
        if (lane == 0) sm.wbytes[w] = running;
        __syncthreads();
        QB_STAMP(desc, 67, 1, qb_t0);  // after fix-ups and offsets

        // warp 0 publishes the tile's byte count and looks back (32 predecessors per round) while the others stage
        if (w == 0) {
            const unsigned pre = (stream && t == 0 && run_in > 0 && sm.first_ne) ? 1u : 0u;  // stream.cpp:171-178
            unsigned       acc = pre;
            for (int ww = 0; ww < kEncWarps; ++ww) acc += sm.wbytes[ww];
            if (lane == 0 && t > 0) st_word(desc + kWordBytes, pack_word(acc, ST_AGG, epoch));
            const uint64_t off = warp_lookback<uint64_t>(
                t, (uint64_t)(stream ? 0u : kHeader), (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = wait_word(desc - (int64_t)(t - p) * kEncDescWords + kWordBytes, epoch);
                    st                = word_status(wd, epoch);
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return a + b; });
            if (lane == 0) {
                st_word(desc + kWordBytes, pack_word(off + acc, ST_INCL, epoch));
                sm.pre = pre, sm.tile_total = acc, sm.tile_off = off;
                if (pre) sm.stage[0] = (unsigned char)(kOpRun | (run_in - 1));
            }
        }
        unsigned wb = (stream && t == 0 && run_in > 0 && sm.first_ne) ? 1u : 0u;
        for (unsigned ww = 0; ww < w; ++ww) wb += sm.wbytes[ww];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned len = (lens >> (4 * k)) & 15u;
            if (len) {
                unsigned char* d = sm.stage + wb + ((offs[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
                const unsigned v = lo[k];
                d[0] = (unsigned char)v;
                if (len > 1) d[1] = (unsigned char)(v >> 8);
                if (len > 2) { d[2] = (unsigned char)(v >> 16); d[3] = (unsigned char)(v >> 24); }
                if (len > 4) d[4] = (unsigned char)(px[k] >> 24);
            }
        }
        __syncthreads();
        QB_STAMP(desc, 68, 0, qb_t0);  // after byte look-back and staging

        // ================= capacity cut (util.hpp:240-246), then realigned 16-byte copy-out =================
        const uint64_t tile_off   = sm.tile_off;
        const unsigned tile_total = sm.tile_total;
        const uint64_t cap        = P.out_cap;
        const bool     fits_all   = tile_off + tile_total <= cap;
        const bool     is_cut     = tile_off <= cap && !fits_all;
        if (is_cut) {  // rare: find the largest chunk boundary that still fits and the first chunk that does not
            const unsigned limit = (unsigned)(cap - tile_off);
            if (tid == 0 && sm.pre) atomicMax(&sm.cut_rel, 1u);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned len = (lens >> (4 * k)) & 15u;
                if (len) {
                    const unsigned end = wb + ((offs[k >> 1] >> (16 * (k & 1))) & 0xFFFFu) + len;
                    if (end <= limit) atomicMax(&sm.cut_rel, end);
                    else atomicMin(&sm.cut_px, chunk0 + k * 32 + lane);
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k)  // the owner of the first refused chunk records its run position for the carry-out
                if (chunk0 + k * 32 + lane == sm.cut_px) sm.cut_q = ((lo[k] & 63u) + 1u) % kRunLimit;
            __syncthreads();
        }
        const unsigned copy_len = fits_all ? tile_total : (is_cut ? sm.cut_rel : 0u);
        if (copy_len) {
            uint8_t*        dst  = out_img + tile_off;
            const unsigned  head = min(copy_len, (16u - (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
            const unsigned  nv   = (copy_len - head) >> 4;
            if (tid < head) dst[tid] = sm.stage[tid];
            // global chunk c is 16-byte aligned; its source starts at stage[head + 16c], any alignment mod 4
            const unsigned* s32 = reinterpret_cast<const unsigned*>(sm.stage);
            const unsigned  sh8 = (head & 3u) * 8u, w0 = head >> 2;
            for (unsigned c = tid; c < nv; c += kEncThreads) {
                const unsigned* q = s32 + w0 + 4 * c;
                const unsigned  a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4];
                reinterpret_cast<uint4*>(dst + head)[c] = make_uint4(__funnelshift_r(a0, a1, sh8), __funnelshift_r(a1, a2, sh8),
                                                                     __funnelshift_r(a2, a3, sh8), __funnelshift_r(a3, a4, sh8));
            }
            const unsigned done = head + (nv << 4);
            if (tid < copy_len - done) dst[done + tid] = sm.stage[done + tid];
        }
        if (!stream && t == 0 && tid < kHeader) out_img[tid] = P.header[tid];

        QB_STAMP(desc, 68, 1, qb_t0);  // after copy-out
        // ================= results =================
        const bool is_last = t == P.tiles_per_image - 1;
        if (!(is_cut || (is_last && fits_all))) return;
        EncResult* res = P.results + img;
        if (!stream) {
            if (tid == 0) {
                uint64_t written  = tile_off + copy_len;
                unsigned complete = 0;
                if (fits_all && written + kMarker <= cap) {  // util.hpp:151-161: all eight bytes or none
                    for (unsigned b = 0; b < kMarker; ++b) out_img[written + b] = b == kMarker - 1 ? 1 : 0;
                    written += kMarker;
                    complete = 1;
                }
                res->written   = written;
                res->complete  = complete;
                res->processed = fits_all ? N : 0;
            }
            return;
        }
        // stream mode: carry-out after the last consumed pixel (stream.cpp:222-238)
        if (fits_all) {
            if (tid < 64) res->state.table[tid] = sm.incl[tid];
            if (tid == 0) {
                const unsigned last = load_pixel<CH>(in_img, N - 1, aligned4);
                const unsigned prev = N >= 2 ? load_pixel<CH>(in_img, N - 2, aligned4) : (init ? init->prev : kStartPixel);
                res->written     = tile_off + copy_len;
                res->processed   = N;
                res->complete    = 1;
                res->state.prev  = last;
                res->state.run   = last == prev ? sm.last_q : 0u;
            }
            return;
        }
        // the first refused chunk belongs to tile-local pixel cut_px
        const unsigned ci   = sm.cut_px;
        const uint64_t cg   = tile_start + ci;
        const unsigned cpx  = load_pixel<CH>(in_img, cg, aligned4);
        const unsigned cprv = cg ? load_pixel<CH>(in_img, cg - 1, aligned4) : (init ? init->prev : kStartPixel);
        unsigned       n_done;  // tile-local count of consumed pixels
        unsigned       run_out, prev_out;
        if (cpx != cprv) {  // a colour / index chunk was refused: pixel not consumed, table slot restored
            n_done = ci, run_out = 0, prev_out = cprv;
        } else {
            const unsigned q = sm.cut_q;
            if (q == 0) n_done = ci, run_out = kRunLimit - 1, prev_out = cprv;  // refused RUN(62): counter back to 61
            else n_done = ci + 1, run_out = q, prev_out = cpx;                 // refused flush: the run stays pending
        }
        if (tid < 64) {
            // table after the consumed pixels of this tile: last differing pixel per slot, else the carry-in
            unsigned v = sm.excl[tid];
            if (n_done) {
                unsigned cur = load_pixel<CH>(in_img, tile_start + n_done - 1, aligned4);
                for (int j = (int)n_done - 1; j >= 0; --j) {
                    const uint64_t g = tile_start + (unsigned)j;
                    const unsigned p = g ? load_pixel<CH>(in_img, g - 1, aligned4) : (init ? init->prev : kStartPixel);
                    if (cur != p && slot_of(cur) == tid) { v = cur; break; }
                    cur = p;
                }
            }
            res->state.table[tid] = v;
        }
        if (tid == 0) {
            res->written    = tile_off + copy_len;
            res->processed  = tile_start + n_done;
            res->complete   = 0;
            res->state.prev = prev_out;
            res->state.run  = run_out;
        }
    }
}  // namespace qb
