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
    };

    constexpr int kEncWarps = 8, kEncThreads = kEncWarps * 32;
    constexpr int kEncDescWords = 72;  // 64 table words, [64] run carry, [65] byte offset carry, padding to 576 B
    constexpr int kWordLne = 64, kWordBytes = 65;
    constexpr uint64_t kLneBias = 64;  // run carry payload = last differing pixel index + 64 (>= 1: run-in <= 62)

    template <int K>
    struct EncSmem {
        static constexpr int T = kEncThreads * K;
        static constexpr int kStage = ((T * 5 + 48) + 127) / 128 * 128;
        alignas(128) unsigned char stage[kStage];  // RGB input staging, later the output staging
        uint2    rec[T];                           // per pixel: x = chunk bytes 0..3, y = byte4 | len << 8 | offset << 16
        unsigned wtab[kEncWarps * 64];             // per warp: last differing pixel value per slot (sentinel = none)
        unsigned inw[kEncWarps * 64];              // per warp: table state on entry to the warp's chunk
        unsigned excl[64];                         // table state on entry to the tile
        unsigned incl[64];                         // table state on exit from the tile
        int      wlne[kEncWarps];                  // tile-local index of the warp's last differing pixel, -1 none
        unsigned wbytes[kEncWarps];
        unsigned wbase[kEncWarps];
        uint64_t tile_off;
        uint64_t lne_payload;  // inclusive run carry on entry (biased)
        unsigned ticket, base62, tile_total, pre, cut_rel, cut_px, last_q, first_ne;
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

    // chunk for a pixel that differs from its predecessor and missed the table: simple.cpp:59-79, util.hpp:163-225
    template <int CH>
    __device__ __forceinline__ uint2 colour_chunk(unsigned cur, unsigned prv)
    {
        const unsigned d = sub4(cur, prv);  // wrapping i8 deltas, simple.cpp:66-71
        if (CH == 4 && (d >> 24) != 0) return make_uint2(kOpRgba | cur << 8, (cur >> 24) | 5u << 8);
        const unsigned t = add4(d, 0x00020202u);  // bias_op_diff
        if ((t & 0x00FCFCFCu) == 0)               // util.hpp:102-107
            return make_uint2(kOpDiff | (t & 3u) << 4 | ((t >> 8) & 3u) << 2 | ((t >> 16) & 3u), 1u << 8);
        const unsigned dr = d & 255u, dg = (d >> 8) & 255u, db = (d >> 16) & 255u;
        const unsigned vg = (dg + 32u) & 255u, vr = (dr - dg + 8u) & 255u, vb = (db - dg + 8u) & 255u;
        if (((vg >> 6) | (vr >> 4) | (vb >> 4)) == 0)  // util.hpp:109-114
            return make_uint2(kOpLuma | vg | (vr << 4 | vb) << 8, 2u << 8);
        return make_uint2(kOpRgb | cur << 8, 4u << 8);
    }

    template <int CH, int K>
    __global__ void __launch_bounds__(kEncThreads) encode_kernel(const EncParams P)
    {
        using S            = EncSmem<K>;
        constexpr int T    = S::T;
        S&            sm   = *reinterpret_cast<S*>(QB_DYN_SMEM);
        const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
        const bool     stream = (P.flags & ENC_STREAM) != 0;

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
        const uint64_t g_next    = tile_start + chunk0 + K * 32;
        const unsigned chunk_next = g_next < N ? load_pixel<CH>(in_img, g_next, aligned4) : 0u;

        // ================= phase A: classify every pixel against the warp-local table =================
        sm.wtab[w * 64 + lane]      = sentinel(lane);
        sm.wtab[w * 64 + 32 + lane] = 0u;
        __syncwarp();

        unsigned eqmask = 0, unresmask = 0;
        int      wl     = -1;  // chunk-local index of the last differing pixel (warp uniform)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned i     = chunk0 + k * 32 + lane;
            const bool     valid = i < n_here;
            const unsigned cur   = px[k];
            const unsigned up    = __shfl_up_sync(kFull, cur, 1);
            const unsigned carry = k == 0 ? chunk_prev : __shfl_sync(kFull, px[k > 0 ? k - 1 : 0], 31);
            const unsigned prv   = lane ? up : carry;
            const bool     eq    = valid && cur == prv;
            const bool     ne    = valid && !eq;
            const unsigned slot  = slot_of(cur);
            const unsigned m     = __match_any_sync(kFull, ne ? slot : 64u + lane);
            const unsigned below = m & lanemask_lt(lane);
            const unsigned pl    = below ? 31u - __clz(below) : lane;
            const unsigned pv    = __shfl_sync(kFull, cur, pl);
            const unsigned tv    = sm.wtab[w * 64 + slot];
            bool           hit = false, unres = false;
            if (ne) {
                if (below) hit = pv == cur;
                else if (tv == sentinel(slot)) unres = true;
                else hit = tv == cur;
            }
            __syncwarp();
            if (ne && (m & lanemask_gt(lane)) == 0) sm.wtab[w * 64 + slot] = cur;
            __syncwarp();
            const unsigned bne = __ballot_sync(kFull, ne);
            if (bne) wl = k * 32 + 31 - __clz(bne);
            eqmask |= (unsigned)eq << k;
            unresmask |= (unsigned)unres << k;
            if (valid) {
                uint2 r = make_uint2(0u, 0u);  // run pixels get their byte in phase B
                if (ne) r = hit ? make_uint2(kOpIndex | slot, 1u << 8) : colour_chunk<CH>(cur, prv);
                sm.rec[i] = r;
            }
            if (i == 0 && t == 0) sm.first_ne = ne;
        }
        if (lane == 0) sm.wlne[w] = wl < 0 ? -1 : (int)chunk0 + wl;
        __syncthreads();

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
        } else if (tid == 64) {
            int tl = -1;
            for (int ww = kEncWarps - 1; ww >= 0; --ww)
                if (sm.wlne[ww] >= 0) { tl = sm.wlne[ww]; break; }
            if (tl >= 0) st_word(desc + kWordLne, pack_word(tile_start + (unsigned)tl + kLneBias, ST_INCL, epoch));
            else st_word(desc + kWordLne, pack_word(0, ST_AGG_EMPTY, epoch));
            uint64_t lp;
            int      p = (int)t - 1;
            for (;;) {
                if (p < 0) { lp = kLneBias - 1 - run_in; break; }  // last differing pixel = -1 - pending run
                const uint64_t wd = wait_word(desc - (int64_t)(t - p) * kEncDescWords + kWordLne, epoch);
                if (word_status(wd, epoch) == ST_AGG_EMPTY) { --p; continue; }
                lp = word_payload(wd);
                break;
            }
            if (tl < 0) st_word(desc + kWordLne, pack_word(lp, ST_INCL, epoch));
            sm.lne_payload = lp;
            sm.base62      = (unsigned)((tile_start + kLneBias - lp) % kRunLimit);  // (tile_start - last_differing) mod 62
        }
        __syncthreads();

        // ================= fix-ups: table probes that left the warp, run bytes =================
        int o;  // run position of tile-local pixel i is congruent to i + o (mod 62), i + o >= 1
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
            o = (int)sm.base62;
            for (int ww = (int)w - 1; ww >= 0; --ww)
                if (sm.wlne[ww] >= 0) { o = -sm.wlne[ww]; break; }
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if ((unresmask >> k) & 1u) {
                const unsigned cur = px[k], slot = slot_of(cur);
                if (sm.inw[w * 64 + slot] == cur) sm.rec[chunk0 + k * 32 + lane] = make_uint2(kOpIndex | slot, 1u << 8);
            }
        }
        const unsigned eq0 = __shfl_sync(kFull, eqmask, 0);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned i     = chunk0 + k * 32 + lane;
            const bool     valid = i < n_here;
            const bool     eq    = (eqmask >> k) & 1u;
            const unsigned beq   = __ballot_sync(kFull, eq);
            const unsigned bne   = __ballot_sync(kFull, valid && !eq);
            if (beq) {
                if (eq) {
                    const unsigned bl = bne & lanemask_lt(lane);
                    const unsigned q  = bl ? lane - (31u - __clz(bl)) : (unsigned)((int)i + o) % kRunLimit;
                    bool           nexteq;
                    if (tile_start + i + 1 >= N) nexteq = stream;  // one-shot flushes at the image end (simple.cpp:91-94)
                    else if (lane < 31) nexteq = (beq >> (lane + 1)) & 1u;
                    else if (k < K - 1) nexteq = (eq0 >> (k + 1)) & 1u;
                    else nexteq = chunk_next == px[k];
                    if (q == 0 || !nexteq)  // util.hpp:227-235
                        sm.rec[i] = make_uint2(kOpRun | ((q + kRunLimit - 1) % kRunLimit), 1u << 8);
                    if (tile_start + i + 1 == N) sm.last_q = q;
                }
            }
            if (bne) o = -(int)(chunk0 + k * 32 + 31 - __clz(bne));
        }

        // ================= carry (3): chunk offsets =================
        unsigned running = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const unsigned i     = chunk0 + k * 32 + lane;
            const bool     valid = i < n_here;
            const unsigned y     = valid ? sm.rec[i].y : 0u;
            const unsigned len   = (y >> 8) & 0xFFu;
            unsigned       x     = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned n = __shfl_up_sync(kFull, x, d);
                if ((int)lane >= d) x += n;
            }
            if (valid) sm.rec[i].y = y | (running + x - len) << 16;
            running += __shfl_sync(kFull, x, 31);
        }
        if (lane == 0) sm.wbytes[w] = running;
        __syncthreads();

        if (tid == 0) {
            const unsigned pre = (stream && t == 0 && run_in > 0 && sm.first_ne) ? 1u : 0u;  // stream.cpp:171-178
            unsigned       acc = pre;
            for (int ww = 0; ww < kEncWarps; ++ww) {
                sm.wbase[ww] = acc;
                acc += sm.wbytes[ww];
            }
            sm.pre        = pre;
            sm.tile_total = acc;
            uint64_t off;
            if (t == 0) {
                off = stream ? 0u : kHeader;
            } else {
                st_word(desc + kWordBytes, pack_word(acc, ST_AGG, epoch));
                off   = 0;
                int p = (int)t - 1;
                for (;;) {
                    const uint64_t wd = wait_word(desc - (int64_t)(t - p) * kEncDescWords + kWordBytes, epoch);
                    off += word_payload(wd);
                    if (word_status(wd, epoch) == ST_INCL) break;
                    --p;
                }
            }
            st_word(desc + kWordBytes, pack_word(off + acc, ST_INCL, epoch));
            sm.tile_off = off;
        }
        __syncthreads();

        // ================= stage the chunks, then coalesced copy-out =================
        const uint64_t tile_off   = sm.tile_off;
        const unsigned tile_total = sm.tile_total;
        const uint64_t cap        = P.out_cap;
        const unsigned sh         = (unsigned)((reinterpret_cast<uintptr_t>(out_img) + tile_off) & 15u);
        const unsigned wb         = sm.wbase[w];
        if (tile_off <= cap) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned i = chunk0 + k * 32 + lane;
                if (i < n_here) {
                    const uint2    r   = sm.rec[i];
                    const unsigned len = (r.y >> 8) & 0xFFu;
                    if (len) {
                        const unsigned rel = wb + (r.y >> 16);
                        if (tile_off + rel + len <= cap) {
                            unsigned char* d = sm.stage + sh + rel;
                            d[0]             = (unsigned char)r.x;
                            if (len > 1) d[1] = (unsigned char)(r.x >> 8);
                            if (len > 2) { d[2] = (unsigned char)(r.x >> 16); d[3] = (unsigned char)(r.x >> 24); }
                            if (len > 4) d[4] = (unsigned char)r.y;
                            if (tile_off + tile_total > cap) atomicMax(&sm.cut_rel, rel + len);
                        } else {
                            atomicMin(&sm.cut_px, i);
                        }
                    }
                }
            }
            if (tid == 0 && sm.pre) {
                sm.stage[sh] = (unsigned char)(kOpRun | (run_in - 1));
                if (tile_off + tile_total > cap) atomicMax(&sm.cut_rel, 1u);
            }
        }
        __syncthreads();

        const bool     fits_all = tile_off + tile_total <= cap;
        const bool     is_cut   = tile_off <= cap && !fits_all;
        const unsigned copy_len = fits_all ? tile_total : (is_cut ? sm.cut_rel : 0u);
        {
            uint8_t*             dst  = out_img + tile_off;
            const unsigned char* src  = sm.stage + sh;
            const unsigned       head = min(copy_len, (16u - sh) & 15u);
            const unsigned       nv   = (copy_len - head) >> 4;
            if (tid < head) dst[tid] = src[tid];
            for (unsigned c = tid; c < nv; c += kEncThreads)
                reinterpret_cast<uint4*>(dst + head)[c] = reinterpret_cast<const uint4*>(src + head)[c];
            const unsigned done = head + (nv << 4);
            if (tid < copy_len - done) dst[done + tid] = src[done + tid];
        }
        if (!stream && t == 0 && tid < kHeader) out_img[tid] = P.header[tid];

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
            const unsigned q = ((sm.rec[ci].x & 63u) + 1u) % kRunLimit;
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
