// decode_kernel.cuh -- tile-parallel QOI decoder for sm_100a (+ the exact sequential kernel it falls back on).
//
// Replaces the serial loop of the reference, impl::decode (source/simple.cpp:100-171).  A tile is kDecTB bytes of
// the chunk stream, one CTA; three chained decoupled look-backs carry what the serial loop carries:
//
//   (1) parse carry  -- an op's length is a function of its tag byte only (simple.cpp:118-165), so a tile is a
//                       map {entry offset 0..4 -> exit offset 0..4}; maps compose, tiles scan them.  No
//                       self-synchronisation is assumed (a late parse of FE FE FE .. never re-syncs).
//   (2) count carry  -- pixels produced so far (sum), and the value-free "slot / alpha" carry: util::hash
//                       (util.hpp:347-351) is linear mod 64, so the table slot written by every op follows from
//                       the literal ops (roots) by a segmented prefix sum of 3dr+5dg+7db without knowing pixels.
//   (3) state carry  -- per tile a transfer function: each of the 64 table slots and `prev` leaves the tile
//                       either as a constant or as (incoming entry e) + delta (mod 256 per channel).  These
//                       compose, so the look-back follows one entry per thread until it meets a constant.
//
// Inside a tile: INDEX ops find their writer (last earlier colour op with the same slot) with
// __match_any_sync + per-warp tables, chains of INDEX -> INDEX are collapsed by pointer jumping in shared
// memory, DIFF/LUMA are segmented mod-256 prefix sums from their root.
//
// SPECULATION.  The slot of an OP_RGB pixel needs its inherited alpha.  The parallel path assumes "alpha =
// alpha of the last OP_RGBA before it, else 255" (exact for every stream whose INDEX ops never change alpha,
// e.g. all RGB images and opaque RGBA images) and that an INDEX op reads a slot that was written.  Every tile
// VERIFIES both against the final pixel values; if any check fails the image is decoded again by
// decode_serial_kernel, which is the reference loop verbatim in behaviour (one lane decodes, the warp stages I/O).
// Verified => exact: by induction over op order, if every root's assumed slot equals the hash of its computed
// value then every writer pointer, hence every value, is the true one.
#pragma once

#include "qb_common.cuh"

namespace qb
{
    struct DecState {  // == StreamDecoder members (include/qoipp/stream.hpp:239-243), pixels packed
        uint32_t prev;
        uint32_t run;
        uint32_t table[64];
    };

    constexpr int kDecRounds = 4;  // verification-driven retry rounds before the sequential kernel takes over

    struct DecResult {
        uint32_t bad;   // round 0 refuted a speculation somewhere in this image
        uint32_t path;  // retry rounds used; + 100 when the sequential kernel produced (part of) the image
        uint64_t pixels;
        uint64_t processed;  // resumable decode: input bytes consumed / output bytes written / carry-out
        uint64_t written;
        uint32_t first_bad[kDecRounds + 1];  // per round: 0 = all verified, else 0xFFFFFFFF - first refuted tile
        uint32_t pad[3];  // pad[0]: decode_ts_kernel (the fast path) was refuted for this image, decode it with decode_tile
        DecState state;
    };

    struct DecControl {  // zeroed with the results before every decode
        uint32_t tickets[kDecRounds + 1];  // tile tickets of the retry rounds
        uint32_t any_bad[kDecRounds + 1];  // some image needs round r + 1
        uint32_t fast_any_bad;             // decode_ts_kernel was refuted for some image
        uint32_t ticket0;                  // tile tickets of the round-0 pass over those images
    };

    struct DecParams {
        const uint8_t*  qoi;
        const uint64_t* offsets;     // [n_images + 1] device; null => single[]
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
        uint32_t*       ticket;
        uint32_t        fast_used;  // round 0 was decode_ts_kernel: decode_finish_kernel decodes the images it flagged
    };

    constexpr int kFixWords = 32, kFixMax = kFixWords - 1;  // word 0: count | decode tag << 8; entries: pos | alpha << 16

#ifndef QB_DEC_WARPS
#define QB_DEC_WARPS 8
#endif
#ifndef QB_DEC_SB
#define QB_DEC_SB 8
#endif
    constexpr int kDecWarps = QB_DEC_WARPS, kDecThreads = kDecWarps * 32, kDecSB = QB_DEC_SB, kDecTB = kDecThreads * kDecSB;
    constexpr int kDecMinCtas = kDecThreads * kDecSB >= 4096 ? 2 : 5;  // CTAs per SM the kernels are compiled for (shared memory bound)
    constexpr int kDecDescWords = 72;
    constexpr int kDwParse = 0, kDwPix = 1, kDwSlot = 2, kDwState = 3;  // 3..67: 64 table entries, then prev

    // link / transfer-function entry codes
    constexpr unsigned kPtrConst = 0xFE00u, kPtrLeafSlot = 0xFF00u, kPtrLeafPrev = 0xFF40u, kNoRoot = 0xFFFFu, kNoOp = 0xFFFFu;
    enum : unsigned { K_INDEX = 0, K_DIFF = 1, K_LUMA = 2, K_RUN = 3, K_RGB = 4, K_RGBA = 5 };

    struct Seg {       // scan element of carry (2), see combine()
        unsigned cnt;  // pixels | ops << 20
        unsigned da;   // delta since the segment root (r,g,b bytes) | alpha of the last OP_RGBA << 24
        unsigned fl;   // root byte position | c << 16 | has_root << 22 | uses_alpha_in << 23 | has_rgba << 24
    };
    constexpr unsigned kFlRoot = 1u << 22, kFlUses = 1u << 23, kFlRgba = 1u << 24;

    __device__ __forceinline__ Seg seg_identity() { return Seg{ 0u, 0u, 0u }; }

    // A happens before B
    __device__ __forceinline__ Seg combine(const Seg& A, const Seg& B)
    {
        Seg C;
        C.cnt                = A.cnt + B.cnt;
        const unsigned hasA  = A.fl & kFlRgba;
        const unsigned alphaA = A.da >> 24;
        unsigned       alpha  = (B.fl & kFlRgba) ? (B.da >> 24) : alphaA;
        unsigned       rgba   = (B.fl & kFlRgba) | hasA;
        unsigned       root, delta, c, flags;
        if (B.fl & kFlRoot) {
            root = B.fl & 0xFFFFu, delta = B.da & 0xFFFFFFu, c = (B.fl >> 16) & 63u, flags = kFlRoot | (B.fl & kFlUses);
            if ((B.fl & kFlUses) && hasA) c = (c + 11u * alphaA) & 63u, flags = kFlRoot;
        } else {
            root = A.fl & 0xFFFFu, delta = add4(A.da, B.da) & 0xFFFFFFu, c = ((A.fl >> 16) + (B.fl >> 16)) & 63u;
            flags = A.fl & (kFlRoot | kFlUses);
        }
        C.da = delta | alpha << 24;
        C.fl = root | c << 16 | flags | rgba;
        return C;
    }

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

    template <class T>
    __device__ __forceinline__ T shfl_up_T(const T& v, unsigned d)
    {
        T        r;
        unsigned a[sizeof(T) / 4];
        __builtin_memcpy(a, &v, sizeof(T));
#pragma unroll
        for (unsigned i = 0; i < sizeof(T) / 4; ++i) a[i] = __shfl_up_sync(kFull, a[i], d);
        __builtin_memcpy(&r, a, sizeof(T));
        return r;
    }

    // inclusive scan across the lanes of a warp, then across the (<= 8) warp totals kept in shared memory.
    // Returns the exclusive prefix of the calling thread; `total` receives the fold over the whole CTA.
    // Contains one __syncthreads().
    template <class T, class Comb>
    __device__ __forceinline__ T cta_exclusive_scan(const T& mine, T* wtot /* smem [kDecWarps] */, const T& identity, Comb comb, T& total)
    {
        const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
        T              incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T o = shfl_up_T(incl, d);
            if ((int)lane >= d) incl = comb(o, incl);
        }
        if (lane == 31) wtot[w] = incl;
        __syncthreads();
        T wv = lane < (unsigned)kDecWarps ? wtot[lane] : identity;  // every warp scans the warp totals redundantly
#pragma unroll
        for (int d = 1; d < kDecWarps; d <<= 1) {
            const T o = shfl_up_T(wv, d);
            if ((int)lane >= d) wv = comb(o, wv);
        }
        total         = shfl_T(wv, kDecWarps - 1);
        const T wpre  = shfl_T(wv, (int)((w + 31u) & 31u));
        T       excl  = shfl_up_T(incl, 1);
        if (lane == 0) excl = identity;
        return w ? comb(wpre, excl) : excl;
    }

    __device__ __forceinline__ unsigned op_length(unsigned tag)  // simple.cpp:118-165
    {
        return 1u + ((tag >> 6) == 2u) + 3u * (tag == kOpRgb) + 4u * (tag == kOpRgba);
    }

    struct DecSmem {
        alignas(16) unsigned char bytes[kDecTB + 48];  // tile bytes at [shift, shift + kDecTB + 8), zero padded
        uint64_t       link[kDecTB];                   // by byte position of a root's tag: ptr << 32 | add
        unsigned       op_meta[kDecTB];                // pixoff | (npix-1) << 17 | kind << 23 | slot << 26
        unsigned       op_val[kDecTB];                 // delta since the root, later the pixel value
        unsigned short op_pos[kDecTB];
        unsigned short op_root[kDecTB];                // byte position of the segment root, kNoRoot = before the tile
        unsigned short wtab[kDecWarps * 64];           // per warp: op index of the last colour op per slot
        unsigned       in_state[65], fn_code[65], fn_add[65];
        Seg            wseg[kDecWarps];
        Map            wmap[kDecWarps];
        uint64_t       pix_base;
        unsigned       fixe[kFixMax];
        unsigned       ticket, entry, slot_in, alpha_in, n_ops, n_pix, bad, changed, fixn, fixn0, fix_dirty;
    };

    // op walk of one thread's sub-chunk: f(pos, tag) for every op whose tag lies in [8*tid, 8*tid + 8) and below `limit`
    template <class F>
    __device__ __forceinline__ void walk_ops(const unsigned char* b, unsigned tid, unsigned entry, unsigned limit, F&& f)
    {
        unsigned p = tid * kDecSB + entry;
        const unsigned end = min((tid + 1u) * kDecSB, limit);
        while (p < end) {
            const unsigned tag = b[p];
            f(p, tag);
            p += op_length(tag);
        }
    }

    // element of a single op for carry (2); also returns kind / npix / literal-or-delta
    struct OpInfo {
        unsigned kind, npix, data, lin;  // data: literal (RGB: alpha byte 0) or delta bytes; lin: slot contribution
    };
    __device__ __forceinline__ OpInfo op_info(const unsigned char* b, unsigned p, unsigned tag)
    {
        OpInfo o;
        o.npix = 1, o.data = 0, o.lin = 0;
        if (tag == kOpRgb) {
            o.kind = K_RGB;
            o.data = b[p + 1] | (unsigned)b[p + 2] << 8 | (unsigned)b[p + 3] << 16;
            o.lin  = __dp4a(o.data, 0x00070503u, 0u) & 63u;
        } else if (tag == kOpRgba) {
            o.kind = K_RGBA;
            o.data = b[p + 1] | (unsigned)b[p + 2] << 8 | (unsigned)b[p + 3] << 16 | (unsigned)b[p + 4] << 24;
            o.lin  = slot_of(o.data);
        } else {
            const unsigned hi = tag >> 6;
            if (hi == 0) {
                o.kind = K_INDEX, o.lin = tag & 63u;
            } else if (hi == 1) {  // simple.cpp:136-144
                o.kind = K_DIFF;
                o.data = add4(((tag >> 4) & 3u) | ((tag >> 2) & 3u) << 8 | (tag & 3u) << 16, 0x00FEFEFEu);
                o.lin  = __dp4a(o.data, 0x00070503u, 0u) & 63u;
            } else if (hi == 2) {  // simple.cpp:145-155
                o.kind            = K_LUMA;
                const unsigned rb = b[p + 1];
                const unsigned vg = ((tag & 63u) + 224u) & 255u;
                const unsigned vr = (vg + (rb >> 4) + 248u) & 255u, vb = (vg + (rb & 15u) + 248u) & 255u;
                o.data = vr | vg << 8 | vb << 16;
                o.lin  = __dp4a(o.data, 0x00070503u, 0u) & 63u;
            } else {
                o.kind = K_RUN, o.npix = (tag & 63u) + 1u;  // simple.cpp:156-163
            }
        }
        return o;
    }

    __device__ __forceinline__ Seg op_seg(const OpInfo& o, unsigned p)
    {
        Seg s;
        s.cnt = o.npix | 1u << 20;
        s.da = 0, s.fl = 0;
        switch (o.kind) {
        case K_RGB: s.fl = p | o.lin << 16 | kFlRoot | kFlUses; break;
        case K_RGBA: s.fl = p | o.lin << 16 | kFlRoot | kFlRgba, s.da = o.data & 0xFF000000u; break;
        case K_INDEX: s.fl = p | o.lin << 16 | kFlRoot; break;
        case K_DIFF:
        case K_LUMA: s.da = o.data, s.fl = o.lin << 16; break;
        default: break;
        }
        return s;
    }

    __device__ __forceinline__ Seg seg_shfl_up(const Seg& s, unsigned d)
    {
        return Seg{ __shfl_up_sync(kFull, s.cnt, d), __shfl_up_sync(kFull, s.da, d), __shfl_up_sync(kFull, s.fl, d) };
    }

    // store one pixel (target 3 or 4 bytes), optionally bottom-up rows (simple.cpp:401-408 done in place)
    __device__ __forceinline__ void store_pixel(uint8_t* out, uint64_t pix, unsigned val, const DecParams& P)
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
        const uint64_t o0 = __ldg(P.offsets + lo);
        stream = P.qoi + o0, size = __ldg(P.offsets + lo + 1) - o0;
    }

    // learned alpha for the OP_RGB at byte position p of this tile (earlier rounds), if any
    __device__ __forceinline__ bool fix_lookup(const DecSmem& sm, unsigned p, unsigned& alpha)
    {
        for (unsigned j = 0; j < sm.fixn0; ++j)
            if ((sm.fixe[j] & 0xFFFFu) == p) { alpha = (sm.fixe[j] >> 16) & 255u; return true; }
        return false;
    }

    // one tile (global ticket sm.ticket) of round `round`; all threads of the CTA
    __device__ __forceinline__ void decode_tile(const DecParams& P, DecSmem& sm, unsigned round, unsigned img, unsigned t, unsigned ntiles,
                                                const uint8_t* stream, uint64_t size, unsigned fresh_from)
    {
        const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
        [[maybe_unused]] const long long qb_t0 = QB_T0();
        const uint64_t body_len = size - kHeader;  // every byte after the header is chunk data (simple.cpp:110-113)
        const uint64_t tile_b0  = (uint64_t)t * kDecTB;
        const unsigned limit    = (unsigned)(body_len - tile_b0 < (uint64_t)kDecTB ? body_len - tile_b0 : (uint64_t)kDecTB);
        uint64_t*      desc     = P.desc + (uint64_t)sm.ticket * kDecDescWords;
        const unsigned epoch    = P.epoch + round;
        const Epochs   ep{ epoch, P.epoch, fresh_from };
        uint8_t*       out      = P.out + (uint64_t)img * P.out_stride;
        const uint64_t N        = P.n_pixels;
        DecResult*     res      = P.results + img;
        uint32_t*      fix      = P.fix + (uint64_t)sm.ticket * kFixWords;
        auto word_of = [&](unsigned p, int which) { return desc - (int64_t)(t - p) * kDecDescWords + which; };

        // alpha values learned by earlier rounds for OP_RGB ops of this tile
        if (tid < kFixWords) {
            unsigned n = 0;
            if (round > 0) {
                const unsigned h = fix[0];
                if ((h >> 8) == (P.epoch & 0xFFFFFFu)) n = min(h & 255u, (unsigned)kFixMax);
                if (tid >= 1 && tid <= n) sm.fixe[tid - 1] = fix[tid];
            }
            if (tid == 0) sm.fixn = n, sm.fixn0 = n, sm.fix_dirty = 0, sm.bad = 0;
        }

        // ---- stage the tile: 16-byte aligned chunks land at the same misalignment in shared memory
        const uint8_t* src   = stream + kHeader + tile_b0;
        const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        {
            const uint64_t avail = body_len - tile_b0;  // bytes of the stream from the tile start
            const unsigned want  = (unsigned)(avail < (uint64_t)(kDecTB + 8) ? avail : (uint64_t)(kDecTB + 8));
            const unsigned nvec  = (shift + want + 15u) >> 4;
            const uint4*   vsrc  = reinterpret_cast<const uint4*>(src - shift);
            for (unsigned c = tid; c < nvec; c += kDecThreads) reinterpret_cast<uint4*>(sm.bytes)[c] = __ldg(vsrc + c);
            __syncthreads();
            for (unsigned b = shift + want + tid; b < kDecTB + 48; b += kDecThreads) sm.bytes[b] = 0;  // zero padding, simple.cpp:106
            __syncthreads();
        }
        const unsigned char* B = sm.bytes + shift;

        QB_STAMP(desc, 68, 0, qb_t0);  // ticket + staging
        // ================= carry (1): parse map =================
        Map mymap;
        {
            unsigned win = 0;  // exit offsets of positions j+1..j+5, 3 bits each
#pragma unroll
            for (int j = kDecSB - 1; j >= 0; --j) {
                const unsigned L   = op_length(B[tid * kDecSB + j]);
                const unsigned nxt = j + L;
                const unsigned e   = nxt >= (unsigned)kDecSB ? nxt - kDecSB : (win >> (3u * (L - 1u))) & 7u;
                win                = (win << 3 | e) & 0x7FFFu;
            }
            mymap = map_unpack(win);
        }
        Map       tile_map;
        const Map excl_map = cta_exclusive_scan(mymap, sm.wmap, map_identity(), [](const Map& a, const Map& b) { return map_compose(a, b); }, tile_map);
        if (w == 0) {  // 32 predecessors per look-back round; an inclusive word is the constant map "exit offset"
            if (lane == 0 && t > 0) st_word(desc + kDwParse, pack_word(map_pack(tile_map), ST_AGG, epoch));
            const Map in = warp_lookback_lazy<Map>(
                t, map_const(0), map_identity(),
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(word_of(p, kDwParse));
                    st                = ep.valid(wd, p) ? raw_status(wd) : (unsigned)ST_NONE;
                    return map_unpack((unsigned)word_payload(wd));
                },
                [](const Map& a, const Map& b) { return map_compose(a, b); });
            const unsigned entry = in.lo & 7u;  // `in` is constant: every chain ended in an inclusive word
            if (lane == 0) {
                st_word(desc + kDwParse, pack_word(map_pack(map_const(map_at(tile_map, entry))), ST_INCL, epoch));
                sm.entry = entry;
            }
        }
        __syncthreads();
        const unsigned my_entry = map_at(excl_map, sm.entry);

        QB_STAMP(desc, 68, 1, qb_t0);  // parse maps + look-back 1
        // ================= carry (2): counts, slot / alpha =================
        // one parse of the thread's ops: the decoded op is parked in link[p] (meta << 32 | data) for the record pass
        Seg      mine   = seg_identity();
        unsigned starts = 0;  // bit j: an op starts at byte j of this thread's sub-chunk
        {
            unsigned cnt = 0, delta = 0, c = 0, flags = 0, rootpos = 0, alpha = 0;
            walk_ops(B, tid, my_entry, limit, [&](unsigned p, unsigned tag) {
                const OpInfo o = op_info(B, p, tag);
                starts |= 1u << (p & (kDecSB - 1));
                sm.link[p] = (uint64_t)(o.kind | (o.npix - 1u) << 3 | o.lin << 9) << 32 | o.data;
                cnt += o.npix | 1u << 20;
                switch (o.kind) {
                case K_RGB:  // slot needs the inherited alpha: known if an OP_RGBA (or a learned alpha) came earlier here
                    rootpos = p, delta = 0;
                    if (sm.fixn0 && fix_lookup(sm, p, alpha)) flags |= kFlRgba;  // an earlier round learned the alpha at this op
                    if (flags & kFlRgba) c = o.lin + 11u * alpha, flags = kFlRoot | kFlRgba;
                    else c = o.lin, flags = kFlRoot | kFlUses;
                    break;
                case K_RGBA: rootpos = p, delta = 0, c = o.lin, alpha = o.data >> 24, flags = kFlRoot | kFlRgba; break;
                case K_INDEX: rootpos = p, delta = 0, c = o.lin, flags = kFlRoot | (flags & kFlRgba); break;
                case K_DIFF:
                case K_LUMA: delta = add4(delta, o.data), c += o.lin; break;
                default: break;
                }
            });
            mine.cnt = cnt;
            mine.da  = (delta & 0xFFFFFFu) | alpha << 24;
            mine.fl  = rootpos | (c & 63u) << 16 | flags;
        }
        Seg       tot;
        const Seg excl = cta_exclusive_scan(mine, sm.wseg, seg_identity(), [](const Seg& a, const Seg& b) { return combine(a, b); }, tot);
        if (w == 0) {  // pixels before this tile
            const unsigned npix = tot.cnt & 0xFFFFFu;
            if (lane == 0 && t > 0) st_word(desc + kDwPix, pack_word(npix, ST_AGG, epoch));
            const uint64_t base = warp_lookback_lazy<uint64_t>(
                t, (uint64_t)0, (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(word_of(p, kDwPix));
                    st                = ep.valid(wd, p) ? raw_status(wd) : (unsigned)ST_NONE;
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return a + b; });
            if (lane == 0) {
                const uint64_t total = base + npix;
                st_word(desc + kDwPix, pack_word(total < (1ull << 41) ? total : (1ull << 41), ST_INCL, epoch));
                sm.pix_base = base, sm.n_ops = tot.cnt >> 20, sm.n_pix = npix;
            }
        } else if (w == 1) {  // slot and alpha of the value entering the tile
            // payload: c | root << 6 | uses << 7 | rgba << 8 | alpha << 9
            auto pack = [](const Seg& q) {
                return ((q.fl >> 16) & 63u) | ((q.fl & kFlRoot) ? 64u : 0u) | ((q.fl & kFlUses) ? 128u : 0u) |
                       ((q.fl & kFlRgba) ? 256u : 0u) | (q.da >> 24) << 9;
            };
            auto unpack = [](uint64_t v64) {
                const unsigned v = (unsigned)v64;
                return Seg{ 0u, (v >> 9) << 24, (v & 63u) << 16 | ((v & 64u) ? kFlRoot : 0u) | ((v & 128u) ? kFlUses : 0u) | ((v & 256u) ? kFlRgba : 0u) };
            };
            const Seg start = Seg{ 0u, 255u << 24, 53u << 16 | kFlRoot | kFlRgba };  // {0,0,0,255}: slot 53 (simple.cpp:108)
            if (lane == 0 && t > 0) st_word(desc + kDwSlot, pack_word(pack(tot), ST_AGG, epoch));
            const Seg acc = warp_lookback_lazy<Seg>(
                t, start, seg_identity(),
                [&](unsigned p, unsigned& st) {
                    if (p < ep.fresh_from) {  // a tile finished by an earlier round: take slot and alpha of its actual last pixel
                        const uint64_t wd = ld_word(word_of(p, kDwState + 64));
                        const unsigned v  = (unsigned)word_payload(wd);
                        st                = ep.valid(wd, p) ? (unsigned)ST_INCL : (unsigned)ST_NONE;
                        return Seg{ 0u, v & 0xFF000000u, slot_of(v) << 16 | kFlRoot | kFlRgba };
                    }
                    const uint64_t wd = ld_word(word_of(p, kDwSlot));
                    st                = ep.valid(wd, p) ? raw_status(wd) : (unsigned)ST_NONE;
                    return unpack(word_payload(wd));
                },
                [](const Seg& a, const Seg& b) { return combine(a, b); });
            if (lane == 0) {
                sm.slot_in  = (acc.fl >> 16) & 63u;  // concrete: the chain ended in an inclusive word
                sm.alpha_in = acc.da >> 24;
                st_word(desc + kDwSlot, pack_word(pack(combine(acc, tot)), ST_INCL, epoch));
            }
        }
        __syncthreads();

        QB_STAMP(desc, 69, 0, qb_t0);  // walk + look-back 2
        // ================= per-op records (compact, stream order) =================
        const unsigned n_ops = sm.n_ops;
        {
            const unsigned alpha_in = sm.alpha_in, slot_in = sm.slot_in;
            unsigned k      = excl.cnt >> 20;
            unsigned pixoff = excl.cnt & 0xFFFFFu;
            unsigned alpha  = (excl.fl & kFlRgba) ? excl.da >> 24 : alpha_in;
            unsigned slot   = (excl.fl & kFlRoot) ? (((excl.fl >> 16) & 63u) + ((excl.fl & kFlUses) ? 11u * alpha_in : 0u)) & 63u
                                                  : (slot_in + ((excl.fl >> 16) & 63u)) & 63u;
            unsigned root   = (excl.fl & kFlRoot) ? (excl.fl & 0xFFFFu) : kNoRoot;
            unsigned delta  = excl.da & 0xFFFFFFu;
            for (unsigned bits = starts; bits; bits &= bits - 1u) {
                const unsigned p    = tid * kDecSB + (unsigned)__ffs((int)bits) - 1u;
                const uint64_t L    = sm.link[p];
                const unsigned meta = (unsigned)(L >> 32), data = (unsigned)L;
                const unsigned kind = meta & 7u, npix1 = (meta >> 3) & 63u, lin = (meta >> 9) & 63u;
                switch (kind) {
                case K_RGB: {
                    if (sm.fixn0) fix_lookup(sm, p, alpha);
                    const unsigned lit = data | alpha << 24;  // speculated alpha, verified below
                    slot = slot_of(lit), root = p, delta = 0;
                    sm.link[p] = (uint64_t)kPtrConst << 32 | lit;
                } break;
                case K_RGBA:
                    alpha = data >> 24, slot = lin, root = p, delta = 0;
                    sm.link[p] = (uint64_t)kPtrConst << 32 | data;
                    break;
                case K_INDEX: slot = lin, root = p, delta = 0; break;
                case K_DIFF:
                case K_LUMA: delta = add4(delta, data) & 0xFFFFFFu, slot = (slot + lin) & 63u; break;
                default: break;
                }
                sm.op_pos[k]  = (unsigned short)p;
                sm.op_root[k] = (unsigned short)root;
                sm.op_val[k]  = delta;
                sm.op_meta[k] = pixoff | npix1 << 17 | kind << 23 | slot << 26;
                ++k, pixoff += npix1 + 1u;
            }
        }
        sm.wtab[tid]               = (unsigned short)kNoOp;
        sm.wtab[tid + kDecThreads] = (unsigned short)kNoOp;
        __syncthreads();

        QB_STAMP(desc, 69, 1, qb_t0);  // records
        // ================= writers of the INDEX ops =================
        const unsigned per_warp = ((n_ops + kDecWarps * 32 - 1) / (kDecWarps * 32)) * 32;  // ops per warp, multiple of 32
        const unsigned k0 = w * per_warp, k1 = min(k0 + per_warp, n_ops);
        // value of colour op `wk` as a link: (its root's link) + its delta; roots before the tile hang off `prev`
        auto link_of_op = [&](unsigned wk) -> uint64_t {
            const unsigned r = sm.op_root[wk];
            return (uint64_t)r << 32 | sm.op_val[wk];  // ptr = root position (or kNoRoot), add = delta; jumped below
        };
        for (unsigned kb = k0; kb < k1; kb += 32) {
            const unsigned k      = kb + lane;
            const bool     valid  = k < k1;
            const unsigned meta   = valid ? sm.op_meta[k] : (K_RUN << 23);
            const unsigned kind   = (meta >> 23) & 7u, slot = meta >> 26;
            const bool     colour = kind != K_RUN;
            const unsigned m      = __match_any_sync(kFull, colour ? slot : 64u + lane);
            const unsigned below  = m & lanemask_lt(lane);
            const unsigned prevw  = sm.wtab[w * 64 + slot];
            if (kind == K_INDEX && valid) {
                const unsigned wk = below ? kb + (31u - __clz(below)) : prevw;
                // unresolved inside the warp: remember "look in earlier warps" as ptr = kPtrLeafSlot + slot (fixed below)
                sm.link[sm.op_pos[k]] = wk != kNoOp ? link_of_op(wk) : (uint64_t)(kPtrLeafSlot + slot) << 32;
            }
            __syncwarp();
            if (colour && (m & lanemask_gt(lane)) == 0) sm.wtab[w * 64 + slot] = (unsigned short)k;
            __syncwarp();
        }
        __syncthreads();
        // INDEX ops whose writer lies in an earlier warp of this tile
        for (unsigned kb = k0; kb < k1; kb += 32) {
            const unsigned k = kb + lane;
            if (k < k1) {
                const unsigned meta = sm.op_meta[k];
                if (((meta >> 23) & 7u) == K_INDEX) {
                    const unsigned p = sm.op_pos[k];
                    if ((unsigned)(sm.link[p] >> 32) == kPtrLeafSlot + (meta >> 26)) {
                        for (int ww = (int)w - 1; ww >= 0; --ww) {
                            const unsigned wk = sm.wtab[ww * 64 + (meta >> 26)];
                            if (wk != kNoOp) { sm.link[p] = link_of_op(wk); break; }
                        }
                    }
                }
            }
        }
        __syncthreads();

        QB_STAMP(desc, 70, 0, qb_t0);  // writers
        // ================= collapse INDEX -> INDEX chains (pointer jumping; 64-bit links are read/written whole) =================
        // a link's ptr is: a root position (< kPtrConst) still to be followed, kNoRoot (= value entering the tile as prev),
        // kPtrConst, kPtrLeafSlot + s or kPtrLeafPrev
        for (;;) {
            bool changed = false;
            for (unsigned k = tid; k < n_ops; k += kDecThreads) {
                if (((sm.op_meta[k] >> 23) & 7u) != K_INDEX) continue;
                const unsigned p = sm.op_pos[k];
                uint64_t       L = sm.link[p];
                unsigned       ptr = (unsigned)(L >> 32);
                if (ptr == kNoRoot) {
                    sm.link[p] = (uint64_t)kPtrLeafPrev << 32 | (unsigned)L;
                } else if (ptr < kPtrConst) {
                    const uint64_t Tl   = sm.link[ptr];
                    const unsigned tptr = (unsigned)(Tl >> 32);
                    // the target root's own link may itself still point at a root: follow again next round
                    sm.link[p] = (uint64_t)(tptr == kNoRoot ? kPtrLeafPrev : tptr) << 32 | add4((unsigned)L, (unsigned)Tl);
                    changed    = true;
                }
            }
            if (!__syncthreads_or(changed)) break;
        }

        QB_STAMP(desc, 70, 1, qb_t0);  // pointer jumping
        // ================= carry (3): the tile's transfer function, look-back, concrete state =================
        if (tid < 65) {
            // outgoing entry e: table slot e (< 64) or prev (64)
            unsigned wk = kNoOp;
            if (tid < 64) {
                for (int ww = kDecWarps - 1; ww >= 0; --ww) {
                    const unsigned x = sm.wtab[ww * 64 + tid];
                    if (x != kNoOp) { wk = x; break; }
                }
            } else if (n_ops) {
                wk = n_ops - 1;
            }
            unsigned code = tid == 64 ? kPtrLeafPrev : kPtrLeafSlot + tid, add = 0;
            if (wk != kNoOp) {
                const unsigned r = sm.op_root[wk];
                add              = sm.op_val[wk];
                if (r == kNoRoot) code = kPtrLeafPrev;
                else {
                    const uint64_t L = sm.link[r];
                    code = (unsigned)(L >> 32), add = add4(add, (unsigned)L);
                }
            }
            sm.fn_code[tid] = code, sm.fn_add[tid] = add;
            // payload: add | entry code << 32 (0..63 slot, 64 prev, 65 constant)
            auto enc = [](unsigned c) { return c == kPtrConst ? 65u : (c == kPtrLeafPrev ? 64u : c - kPtrLeafSlot); };
            if (code != kPtrConst) st_word(desc + kDwState + tid, pack_word((uint64_t)enc(code) << 32 | add, ST_AGG, epoch));
            else st_word(desc + kDwState + tid, pack_word(add, ST_INCL, epoch));  // a constant is already inclusive
            // incoming value of entry `tid`: follow the chain through the predecessors
            unsigned e = tid, acc = 0, v;
            for (int p = (int)t - 1;; --p) {
                if (p < 0) {  // simple.cpp:103-108: zero table, prev = start, start stored at its slot
                    v = add4((e == 64 || e == 53) ? kStartPixel : 0u, acc);
                    break;
                }
                const uint64_t wd = wait_word(word_of((unsigned)p, kDwState + (int)e), ep, (unsigned)p);
                const uint64_t pl = word_payload(wd);
                if (raw_status(wd) == ST_INCL) { v = add4((unsigned)pl, acc); break; }
                acc = add4(acc, (unsigned)pl);
                const unsigned c = (unsigned)(pl >> 32);
                if (c == 65u) { v = acc; break; }
                e = c;
            }
            sm.in_state[tid] = v;
        }
        __syncthreads();
        if (tid < 65) {
            const unsigned code = sm.fn_code[tid], add = sm.fn_add[tid];
            const unsigned v    = code == kPtrConst ? add : add4(sm.in_state[code == kPtrLeafPrev ? 64 : code - kPtrLeafSlot], add);
            st_word(desc + kDwState + tid, pack_word(v, ST_INCL, epoch));
            sm.fn_add[tid] = v;  // concrete outgoing state
        }

        QB_STAMP(desc, 71, 0, qb_t0);  // state look-back
        // ================= values, verification, pixels =================
        for (unsigned k = tid; k < n_ops; k += kDecThreads) {
            const unsigned r = sm.op_root[k];
            unsigned       base;
            if (r == kNoRoot) base = sm.in_state[64];
            else {
                const uint64_t L   = sm.link[r];
                const unsigned ptr = (unsigned)(L >> 32);
                base = ptr == kPtrConst ? (unsigned)L : add4(sm.in_state[ptr == kPtrLeafPrev ? 64 : ptr - kPtrLeafSlot], (unsigned)L);
            }
            sm.op_val[k] = add4(base, sm.op_val[k]);
        }
        __syncthreads();
        const uint64_t pix_base = sm.pix_base;
        bool           bad      = false;
        for (unsigned k = tid; k < n_ops; k += kDecThreads) {
            const unsigned meta = sm.op_meta[k], kind = (meta >> 23) & 7u, val = sm.op_val[k];
            if (pix_base + (meta & 0x1FFFFu) < N) {  // ops past the image are never executed by the reference
                if (kind == K_RGB) {  // simple.cpp:119-123: the alpha is inherited from the previous pixel
                    const unsigned actual = (k ? sm.op_val[k - 1] : sm.in_state[64]) >> 24;
                    if ((val >> 24) != actual) {
                        bad = true;
                        // remember the alpha seen here for the next round (exact if everything before this op was exact)
                        const unsigned p = sm.op_pos[k], e = p | actual << 16;
                        unsigned       j = 0;
                        for (; j < sm.fixn0; ++j)
                            if ((sm.fixe[j] & 0xFFFFu) == p) break;
                        if (j == sm.fixn0) j = atomicAdd(&sm.fixn, 1u);
                        if (j < (unsigned)kFixMax) sm.fixe[j] = e;
                        sm.fix_dirty = 1;
                    }
                }
                if (kind == K_INDEX) bad |= slot_of(val) != (meta >> 26);  // a never-written (or mis-predicted) slot was read
            }
        }
        if (bad) sm.bad = 1;
        if (P.flip) {  // bottom-up rows: pixels of a tile are not contiguous in the output, store them one by one
            for (unsigned k = tid; k < n_ops; k += kDecThreads) {
                const unsigned meta = sm.op_meta[k], val = sm.op_val[k];
                const uint64_t pix  = pix_base + (meta & 0x1FFFFu);
                const unsigned np1  = (meta >> 17) & 63u;
                for (unsigned j = 0; j <= np1 && pix + j < N; ++j) store_pixel(out, pix + j, val, P);  // OP_RUN clamped (simple.cpp:158)
            }
            __syncthreads();
        } else {
            // pixels of the tile are one contiguous byte range of the output: expand runs into shared memory (the link
            // array is dead by now) at the output's 16-byte phase, then leave as aligned uint4 stores
            __syncthreads();  // all readers of sm.link are done
            unsigned char* stg   = reinterpret_cast<unsigned char*>(sm.link);
            const unsigned tgt   = P.target;
            const unsigned pch   = tgt == 4 ? 4064u : 5440u;  // pixels per staging round: pch * tgt + 15 < sizeof(link)
            const uint64_t avail = pix_base < N ? N - pix_base : 0;
            const unsigned total = (unsigned)(avail < (uint64_t)sm.n_pix ? avail : (uint64_t)sm.n_pix);
            for (unsigned c0 = 0; c0 < total; c0 += pch) {
                const unsigned cn   = min(pch, total - c0);
                uint8_t*       gdst = out + (pix_base + c0) * tgt;
                const unsigned sh   = (unsigned)(reinterpret_cast<uintptr_t>(gdst) & 15u);
                if (total <= pch && (tgt == 3 || (sh & 3u) == 0)) {
                    // common case: the whole tile fits one staging round (and four-byte pixels are word aligned).  Every op
                    // stores its own pixel; the pixels of an OP_RUN (clamped to the image, simple.cpp:158) are written by the
                    // whole warp, 32 at a time -- one lane looping over its run kept the other 31 waiting (measured: 7 % of
                    // the kernel's instructions at a fifth of the lanes).
                    unsigned* s32 = reinterpret_cast<unsigned*>(stg + sh);
                    auto put = [&](unsigned p, unsigned val) {
                        if (tgt == 4) s32[p] = val;
                        else {
                            unsigned char* d = stg + sh + p * 3u;
                            d[0] = (unsigned char)val, d[1] = (unsigned char)(val >> 8), d[2] = (unsigned char)(val >> 16);
                        }
                    };
                    const unsigned lane_ = tid & 31u;
                    for (unsigned kb = tid - lane_; kb < n_ops; kb += kDecThreads) {  // uniform per warp
                        const unsigned k     = kb + lane_;
                        const bool     valid = k < n_ops;
                        const unsigned meta = valid ? sm.op_meta[k] : 0u, val = valid ? sm.op_val[k] : 0u;
                        const unsigned p0 = meta & 0x1FFFFu, np1 = (meta >> 17) & 63u;
                        if (valid && p0 < total) put(p0, val);
                        unsigned runs = __ballot_sync(kFull, valid && np1 != 0);
                        while (runs) {
                            const int src = __ffs((int)runs) - 1;
                            runs &= runs - 1u;
                            const unsigned rp = __shfl_sync(kFull, p0, src), rn = __shfl_sync(kFull, np1, src), rv = __shfl_sync(kFull, val, src);
                            for (unsigned j = 1u + lane_; j <= rn && rp + j < total; j += 32u) put(rp + j, rv);
                        }
                    }
                } else
                for (unsigned k = tid; k < n_ops; k += kDecThreads) {
                    const unsigned meta = sm.op_meta[k];
                    const unsigned p0 = meta & 0x1FFFFu, p1 = p0 + ((meta >> 17) & 63u) + 1u;  // tile-relative pixel range
                    if (p1 <= c0 || p0 >= c0 + cn) continue;
                    const unsigned val = sm.op_val[k];
                    const unsigned a = max(p0, c0) - c0, b = min(p1, c0 + cn) - c0;
                    for (unsigned j = a; j < b; ++j) {
                        unsigned char* d = stg + sh + j * tgt;
                        if (tgt == 4 && (sh & 3u) == 0) *reinterpret_cast<unsigned*>(d) = val;
                        else {
                            d[0] = (unsigned char)val, d[1] = (unsigned char)(val >> 8), d[2] = (unsigned char)(val >> 16);
                            if (tgt == 4) d[3] = (unsigned char)(val >> 24);
                        }
                    }
                }
                __syncthreads();
                const unsigned nbytes = cn * tgt;
                const unsigned head   = min(nbytes, (16u - sh) & 15u);
                const unsigned nv     = (nbytes - head) >> 4;
                if (tid < head) gdst[tid] = stg[sh + tid];
                for (unsigned c = tid; c < nv; c += kDecThreads)
                    reinterpret_cast<uint4*>(gdst + head)[c] = reinterpret_cast<const uint4*>(stg + sh + head)[c];
                const unsigned done = head + (nv << 4);
                if (tid < nbytes - done) gdst[done + tid] = stg[sh + done + tid];
                __syncthreads();
            }
        }
        // a refuted tile makes the image eligible for the next round, from the first such tile on
        if (sm.bad) {
            if (tid == 0) {
                if (round == 0) atomicOr(&res->bad, 1u);
                atomicMax(&res->first_bad[round], 0xFFFFFFFFu - t);
                P.control->any_bad[round] = 1;
            }
            if (sm.fix_dirty && tid < kFixWords) {
                const unsigned n = min(sm.fixn, (unsigned)kFixMax);
                fix[tid] = tid == 0 ? (n | (P.epoch & 0xFFFFFFu) << 8) : (tid <= n ? sm.fixe[tid - 1] : 0u);
            }
        }

        QB_STAMP(desc, 71, 1, qb_t0);  // values + stores
        // ---- the stream ended before the image: the zero padding decodes as INDEX 0 forever (simple.cpp:106,132-135)
        if (t == ntiles - 1) {
            const uint64_t have = pix_base + sm.n_pix;
            const unsigned fill = sm.fn_add[0];
            if (tid == 0) {
                res->pixels = have < N ? have : N;
                if (have < N && slot_of(fill) != 0) {  // cannot happen while the table invariant holds; be safe
                    atomicMax(&res->first_bad[round], 0xFFFFFFFFu - t);
                    P.control->any_bad[round] = 1;
                }
            }
            for (uint64_t pix = have + tid; pix < N; pix += kDecThreads) store_pixel(out, pix, fill, P);
        }
    }

    // round 0: one tile per CTA
    __global__ void __launch_bounds__(kDecThreads, kDecMinCtas) decode_kernel(const DecParams P)
    {
        DecSmem& sm = *reinterpret_cast<DecSmem*>(QB_DYN_SMEM);
        if (threadIdx.x == 0) sm.ticket = atomicInc(P.ticket, P.n_tiles - 1u);
        __syncthreads();
        unsigned       img, t, ntiles;
        const uint8_t* stream;
        uint64_t       size;
        locate_image(P, sm.ticket, img, t, ntiles, stream, size);
        decode_tile(P, sm, 0u, img, t, ntiles, stream, size, 0u);
    }

    // =====================================================================================================
    // Exact sequential decoder: the reference loop (simple.cpp:100-171, stream.cpp:312-447) with one decoding lane per
    // image; the other 31 lanes stage input and output through shared memory.  Runs only when `bad` is set
    // (mode 0), or always for the resumable entry point (mode 1) whose buffers are small by construction.
    // =====================================================================================================
    struct SerialParams {
        DecParams       d;
        uint32_t        mode;      // 0 = redo images flagged bad; 1 = resumable decode of one buffer
        const DecState* init;      // mode 1 carry-in
        uint64_t        in_size;   // mode 1: bytes available (no header), out capacity in bytes is d.out_stride
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
        else stream = P.qoi + P.offsets[img], size = P.offsets[img + 1] - P.offsets[img], first_tile = P.tile_first[img];
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
            // resume behind the last verified tile: its inclusive carry words are the decoder state at that point
            const uint64_t* d = P.desc + (uint64_t)(first_tile + restart - 1) * kDecDescWords;
            pos               = (uint64_t)restart * kDecTB + (word_payload(d[kDwParse]) & 7u);
            px                = word_payload(d[kDwPix]);
            prev              = (unsigned)word_payload(d[kDwState + 64]);
            sm.table[lane] = (unsigned)word_payload(d[kDwState + lane]), sm.table[lane + 32] = (unsigned)word_payload(d[kDwState + 32 + lane]);
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

    // Everything after round 0, in ONE cooperative launch of co-resident persistent CTAs (so that an image that
    // verified costs a single empty launch): rounds 1..kDecRounds re-decode, per image, the tiles from the first refuted
    // one on with the alphas learned by the round before (grid-wide barrier between rounds); what still fails after the
    // last round is decoded by the sequential loop, one warp per image, resuming behind the last verified tile.
    __global__ void __launch_bounds__(kDecThreads, kDecMinCtas) decode_finish_kernel(const DecParams P)
    {
        DecSmem& sm = *reinterpret_cast<DecSmem*>(QB_DYN_SMEM);
        if (P.fast_used && P.control->fast_any_bad) {  // same value in every CTA: final when this kernel starts
            // round 0 of the general path for the images the thread-serial fast path (decode_ts.cuh) could not verify
            for (;;) {
                __syncthreads();
                if (threadIdx.x == 0) sm.ticket = atomicAdd(&P.control->ticket0, 1u);
                __syncthreads();
                if (sm.ticket >= P.n_tiles) break;
                unsigned       img, t, ntiles;
                const uint8_t* stream;
                uint64_t       size;
                locate_image(P, sm.ticket, img, t, ntiles, stream, size);
                if (P.results[img].pad[0] == 0) continue;
                decode_tile(P, sm, 0u, img, t, ntiles, stream, size, 0u);
            }
            QB_GRID_SYNC();
        }
        for (unsigned round = 1; round <= (unsigned)kDecRounds; ++round) {
            if (P.control->any_bad[round - 1] == 0) break;  // same value in every CTA: final since the last barrier
            for (;;) {
                __syncthreads();  // the previous tile's shared memory is no longer in use
                if (threadIdx.x == 0) sm.ticket = atomicAdd(&P.control->tickets[round], 1u);
                __syncthreads();
                if (sm.ticket >= P.n_tiles) break;
                unsigned       img, t, ntiles;
                const uint8_t* stream;
                uint64_t       size;
                locate_image(P, sm.ticket, img, t, ntiles, stream, size);
                const unsigned fb = P.results[img].first_bad[round - 1];
                if (fb == 0 || t < 0xFFFFFFFFu - fb) continue;  // image verified, or a tile before the first refuted one: final
                decode_tile(P, sm, round, img, t, ntiles, stream, size, 0xFFFFFFFFu - fb);
            }
            QB_GRID_SYNC();
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            SerialParams S{};
            S.d = P, S.mode = 0;
            for (unsigned img = blockIdx.x; img < P.n_images; img += gridDim.x)
                decode_serial_body(S, *reinterpret_cast<SerialSmem*>(QB_DYN_SMEM), img);
        }
    }

#ifndef QB_EMU
    inline cudaError_t dec_set_attrs()
    {
        cudaError_t e = cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_finish_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
#endif
}  // namespace qb
