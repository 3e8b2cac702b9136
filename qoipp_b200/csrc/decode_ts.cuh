// decode_ts.cuh -- thread-serial, warp-per-tile decoder: the fast path of the one-shot QOI decode for sm_100a.
//
// Same pixels as decode_kernel (decode_kernel.cuh) and the reference loop impl::decode (source/simple.cpp:100-171).
// decode_kernel spreads a 2 KB tile over 256 threads (8 bytes each) and spends ~670 thread-instructions per pixel on
// per-op records, CTA scans and barriers.  Here a tile is the 2176 stream bytes of ONE WARP -- every lane owns 68
// consecutive bytes and walks its ops like the reference loop does, state in registers -- and the warps of a CTA are
// independent, persistent workers (no __syncthreads anywhere).  What a lane does not know at the start of its chunk is
// carried symbolically:
//
//   parse   exit offset for each of the five possible entry offsets of the chunk (backward DP over its 68 bytes, bytes
//           in registers), maps scanned across the lanes and looked back across tiles  -> every lane's entry offset;
//   W1      op walk: pixel / op counts and the value-free slot / alpha carry `Seg` (util::hash is linear mod 64)
//           -> per-lane pixel offset, table slot and alpha on entry (warp scan + two look-backs, as decode_kernel);
//   W2      symbolic walk: the lane's column of fin[slot][lane] holds, per table slot, the last value it stores as
//           {base, delta}: base = a constant, the lane's incoming `prev`, or "what an OP_INDEX of this lane read from
//           slot s before the lane stored to it" (an external read; those are listed per lane, at most kDtExt);
//   merge   per slot, the last lane before lane t that stored to it (lane = slot, a 32-step walk along the lanes);
//           incoming `prev` of every lane by a warp scan of {base, origin lane, delta};
//   chase   every base a lane needs -- its incoming prev, its external reads, and for the tile's transfer function the
//           last store per slot -- is followed through the lanes' columns until it is a constant or an entry of the
//           state on entry to the TILE (bounded hops);
//   state   the tile's transfer function is published and looked back exactly as in decode_kernel (65 words per tile);
//   W3      the reference loop proper with concrete values: pixels are stored straight to the image, the lane's column
//           is its private table, external reads come from the lane's list.
//
// SPECULATION (the same two assumptions as decode_kernel, see there): an OP_RGB inherits "alpha of the last OP_RGBA
// before it, else 255", and an OP_INDEX reads a slot that was stored to.  W3 VERIFIES both on the concrete values; a
// refuted image -- or one that overflows an external-read list or the chase bound, or whose stream ends early -- is
// flagged and decoded again by decode_kernel's machinery inside decode_finish_kernel.  Verified => exact.
#pragma once

#include "decode_kernel.cuh"

namespace qb
{
#ifndef QB_DT_WARPS
#define QB_DT_WARPS 4
#endif
#ifndef QB_DT_CTAS
#define QB_DT_CTAS 3
#endif
    constexpr int kDtWarps = QB_DT_WARPS, kDtThreads = kDtWarps * 32;
    constexpr int kDtSB = 68, kDtTB = 32 * kDtSB;  // 17 words per lane: the lanes' chunks start in 32 different banks
    constexpr int kDtRowV = 33, kDtRowC = 68;      // strides of finv[slot][lane] (words) and finc[lane][slot] (bytes)
    constexpr int kDtExt = 16, kDtHops = 12;

    // base codes of a symbolic value {code, val}: val is the value (constant) or the r,g,b delta to add to the base
    constexpr unsigned kCPrev = 64, kCConst = 65;  // 0..63: external read of that slot by the owning lane / tile-in slot
    constexpr unsigned kCCached = 0x80;            // | slot: the column caches an external read (not a store)
    constexpr unsigned kCNone = 0xFF;

    struct DtParams {
        DecParams       d;            // stream(s), output, results, control, epoch as for decode_kernel
        const uint32_t* tile_first;   // [n_images + 1] first fast tile of every image (batch), null for one image
        uint64_t*       desc;         // [n_tiles][kDecDescWords]
        uint32_t*       ticket;
        uint32_t        ticket_base;
        uint32_t        n_tiles;
    };

    struct DtWarpSmem {
        alignas(16) unsigned char bytes[kDtTB + 64];  // tile bytes at [shift, shift + kDtTB + 8), zero padded
        unsigned      finv[64 * kDtRowV];
        unsigned char finc[32 * kDtRowC];
        unsigned char lw[32 * kDtRowC];   // [lane][slot]: last lane before `lane` that stored to `slot`
        unsigned      tin[65];            // concrete state on entry to the tile (64 table slots, prev)
        unsigned      fcode[65], fadd[65];  // the tile's transfer function, tile relative
        unsigned      extv[kDtExt * 32];  // [k][lane] external reads of the lane, in order of first use
        unsigned char extc[kDtExt * 32];
        unsigned      piv[32];            // incoming prev of every lane: value / delta
        unsigned char pic[32], pio[32];   // its base code and the lane the base belongs to
        unsigned char tlw[64];            // last lane of the tile that stored to the slot
    };

    struct DtSym {
        unsigned code, lane, val;
    };

    __device__ __forceinline__ void dt_locate(const DtParams& P, unsigned gt, unsigned& img, unsigned& t, unsigned& ntiles,
                                              const uint8_t*& stream, uint64_t& size)
    {
        const DecParams& D = P.d;
        if (P.tile_first == nullptr) {
            img = 0, t = gt, ntiles = P.n_tiles;
            stream = D.qoi + D.single[0], size = D.single[1] - D.single[0];
            return;
        }
        unsigned lo = 0, hi = D.n_images;  // largest img with tile_first[img] <= gt
        while (hi - lo > 1) {
            const unsigned mid = (lo + hi) >> 1;
            if (__ldg(P.tile_first + mid) <= gt) lo = mid;
            else hi = mid;
        }
        img    = lo;
        const unsigned f = __ldg(P.tile_first + lo);
        t = gt - f, ntiles = __ldg(P.tile_first + lo + 1) - f;
        const uint64_t o0 = __ldg(D.offsets + lo);
        stream = D.qoi + o0, size = __ldg(D.offsets + lo + 1) - o0;
    }

    // One op as the walks see it; everything is computed without branches (the lanes of a warp hold different kinds).
    struct DtOp {
        unsigned tag, pay;  // pay = the four bytes behind the tag
        bool     rgb, rgba, index, run, delta;
    };
    __device__ __forceinline__ DtOp dt_op(unsigned lo, unsigned b4)
    {
        DtOp o;
        o.tag = lo & 0xFFu, o.pay = __funnelshift_r(lo, b4, 8);
        const unsigned k = o.tag >> 6;
        o.rgb = o.tag == kOpRgb, o.rgba = o.tag == kOpRgba;
        o.index = k == 0, o.run = k == 3 && o.tag < kOpRgb, o.delta = k == 1 || k == 2;
        return o;
    }

    // Ops of a lane's chunk in stream order.  `mlo` / `mhi` mark the chunk bytes (0..63 / 64..67) where an op starts (the
    // parse already knows them), so the next op's position never waits for the current op: its bytes are fetched from shared
    // memory one op ahead.
    template <class F>
    __device__ __forceinline__ void dt_walk(const unsigned* words, unsigned byte0, unsigned long long mlo, unsigned mhi, F&& f)
    {
        auto take = [&](unsigned& lo, unsigned& b4) {
            unsigned p;
            if (mlo) p = (unsigned)__ffsll((long long)mlo) - 1u, mlo &= mlo - 1ull;
            else p = 63u + (unsigned)__ffs((int)mhi), mhi &= mhi - 1u;
            const unsigned  a  = byte0 + p, sh = (a & 3u) * 8u;
            const unsigned* w  = words + (a >> 2);
            const unsigned  w0 = w[0], w1 = w[1];
            lo = __funnelshift_r(w0, w1, sh), b4 = w1 >> sh;
        };
        bool     more = (mlo | mhi) != 0;
        unsigned lo = 0, b4 = 0;
        if (more) take(lo, b4);
        while (more) {
            const unsigned clo = lo, cb4 = b4;
            more = (mlo | mhi) != 0;
            if (more) take(lo, b4);
            f(dt_op(clo, cb4));
        }
    }

    __device__ __forceinline__ unsigned dt_diff_delta(unsigned tag)  // simple.cpp:136-144
    {
        return add4(((tag >> 4) & 3u) | ((tag >> 2) & 3u) << 8 | (tag & 3u) << 16, 0x00FEFEFEu);
    }
    __device__ __forceinline__ unsigned dt_luma_delta(unsigned tag, unsigned pay)  // simple.cpp:145-155
    {
        const unsigned rb = pay & 0xFFu, vg = ((tag & 63u) + 224u) & 255u;
        return ((vg + (rb >> 4) + 248u) & 255u) | vg << 8 | ((vg + (rb & 15u) + 248u) & 255u) << 16;
    }
    __device__ __forceinline__ unsigned dt_lin(unsigned d) { return __dp4a(d, 0x00070503u, 0u) & 63u; }

    // one tile, one warp; flags the image (DecResult::pad[0]) when it has to take the general path
    __device__ __forceinline__ void dt_decode_tile(const DtParams& P, DtWarpSmem& sm, unsigned gt)
    {
        const DecParams& D    = P.d;
        const unsigned   lane = threadIdx.x & 31u;
        unsigned         img, t, ntiles;
        const uint8_t*   stream;
        uint64_t         size;
        dt_locate(P, gt, img, t, ntiles, stream, size);
        const uint64_t body_len = size - kHeader;  // every byte after the header is chunk data (simple.cpp:110-113)
        const uint64_t tile_b0  = (uint64_t)t * kDtTB;
        const unsigned limit    = (unsigned)(body_len - tile_b0 < (uint64_t)kDtTB ? body_len - tile_b0 : (uint64_t)kDtTB);
        uint64_t*      desc     = P.desc + (uint64_t)gt * kDecDescWords;
        const unsigned epoch    = D.epoch;
        uint8_t*       out      = D.out + (uint64_t)img * D.out_stride;
        const uint64_t N        = D.n_pixels;
        DecResult*     res      = D.results + img;
        auto word_of = [&](unsigned p, int which) { return desc - (int64_t)(t - p) * kDecDescWords + which; };
        auto status_of = [&](uint64_t wd) { return word_status(wd, epoch); };
        bool bad = false;
        [[maybe_unused]] const long long qb_t0 = QB_T0();

        // ---- stage the tile: 16-byte aligned chunks land at the same misalignment in shared memory
        const uint8_t* src   = stream + kHeader + tile_b0;
        const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u);
        {
            const uint64_t avail = body_len - tile_b0;
            const unsigned want  = (unsigned)(avail < (uint64_t)(kDtTB + 8) ? avail : (uint64_t)(kDtTB + 8));
            const unsigned nvec  = (shift + want + 15u) >> 4;
            const uint4*   vsrc  = reinterpret_cast<const uint4*>(src - shift);
            for (unsigned c = lane; c < nvec; c += 32) reinterpret_cast<uint4*>(sm.bytes)[c] = __ldg(vsrc + c);
            __syncwarp();
            for (unsigned b = shift + want + lane; b < kDtTB + 64; b += 32) sm.bytes[b] = 0;  // zero padding, simple.cpp:106
            __syncwarp();
        }
        const unsigned* words = reinterpret_cast<const unsigned*>(sm.bytes);
        const unsigned  byte0 = shift + kDtSB * lane;  // this lane's chunk in sm.bytes
        const unsigned  cend  = limit > kDtSB * lane ? min((unsigned)kDtSB, limit - kDtSB * lane) : 0u;

        QB_STAMP(desc, 68, 0, qb_t0);  // staged
        // ================= parse: entry offset of every lane =================
        Map      mymap;
        unsigned bw[kDtSB / 4 + 1];  // chunk bytes 4j .. 4j+3
        {
            const unsigned* w  = words + (byte0 >> 2);
            const unsigned  sh = (byte0 & 3u) * 8u;
            unsigned        pw = w[0];
#pragma unroll
            for (int j = 0; j < kDtSB / 4 + 1; ++j) {
                const unsigned nx = w[j + 1];
                bw[j]             = __funnelshift_r(pw, nx, sh);
                pw                = nx;
            }
            unsigned win = 0;  // exit offsets of positions j+1..j+5, 3 bits each
#pragma unroll
            for (int j = kDtSB - 1; j >= 0; --j) {
                const unsigned L   = op_length((bw[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                const unsigned nxt = j + L;
                const unsigned e   = nxt >= (unsigned)kDtSB ? nxt - kDtSB : (win >> (3u * (L - 1u))) & 7u;
                win                = (win << 3 | e) & 0x7FFFu;
            }
            mymap = map_unpack(win);
        }
        unsigned my_entry;
        {
            Map incl = mymap;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Map o = shfl_up_T(incl, d);
                if ((int)lane >= d) incl = map_compose(o, incl);
            }
            Map excl = shfl_up_T(incl, 1);
            if (lane == 0) excl = map_identity();
            const Map tile_map = shfl_T(incl, 31);
            if (lane == 0 && t > 0) st_word(desc + kDwParse, pack_word(map_pack(tile_map), ST_AGG, epoch));
            const Map in = warp_lookback_lazy<Map>(
                t, map_const(0), map_identity(),
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(word_of(p, kDwParse));
                    st                = status_of(wd);
                    return map_unpack((unsigned)word_payload(wd));
                },
                [](const Map& a, const Map& b) { return map_compose(a, b); });
            const unsigned entry = in.lo & 7u;  // `in` is constant: every chain ended in an inclusive word
            if (lane == 0) st_word(desc + kDwParse, pack_word(map_pack(map_const(map_at(tile_map, entry))), ST_INCL, epoch));
            my_entry = map_at(excl, entry);
        }
        // where the ops of this chunk start: forward over the bytes, `pend` = starts at byte j, j+1, .. j+4
        unsigned long long mlo = 0;
        unsigned           mhi = 0;
        {
            unsigned pend = 1u << my_entry;
#pragma unroll
            for (int j = 0; j < kDtSB; ++j) {
                const unsigned cur = pend & 1u;
                const unsigned L   = op_length((bw[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                pend               = (pend >> 1) | (cur << (L - 1u));
                if (j < 64) mlo |= (unsigned long long)cur << j;
                else mhi |= cur << (j - 64);
            }
            // only tags below the end of the stream are ops
            mlo = cend >= 64u ? mlo : mlo & ((1ull << cend) - 1ull);
            mhi = cend >= 64u ? mhi & ((1u << (cend - 64u)) - 1u) : 0u;
        }

        QB_STAMP(desc, 68, 1, qb_t0);  // parse + look-back
        // ================= W1: counts, slot / alpha carry =================
        Seg mine = seg_identity();
        {
            unsigned cnt = 0, delta = 0, c = 0, flags = 0, alpha = 0;
            dt_walk(words, byte0, mlo, mhi, [&](const DtOp& o) {
                const unsigned d    = (o.tag >> 6) == 1 ? dt_diff_delta(o.tag) : dt_luma_delta(o.tag, o.pay);
                const bool     seen = (flags & kFlRgba) != 0;  // an OP_RGBA came earlier in this chunk: its alpha is the OP_RGB's
                const unsigned lin3 = __dp4a(o.pay & 0xFFFFFFu, 0x00070503u, 0u);
                cnt += (o.run ? (o.tag & 63u) + 1u : 1u) | 1u << 20;
                const bool root = o.rgb || o.rgba || o.index;
                const unsigned c_root = o.rgb ? lin3 + (seen ? 11u * alpha : 0u) : (o.rgba ? slot_of(o.pay) : o.tag & 63u);
                const unsigned f_root = o.rgb ? (seen ? kFlRoot | kFlRgba : kFlRoot | kFlUses)
                                              : (o.rgba ? kFlRoot | kFlRgba : kFlRoot | (flags & kFlRgba));
                delta = root ? 0u : (o.delta ? add4(delta, d) : delta);
                c     = root ? c_root : (o.delta ? c + dt_lin(d) : c);
                flags = root ? f_root : flags;
                alpha = o.rgba ? o.pay >> 24 : alpha;
            });
            mine.cnt = cnt;
            mine.da  = (delta & 0xFFFFFFu) | alpha << 24;
            mine.fl  = (c & 63u) << 16 | flags;
        }
        uint64_t pix_base;
        unsigned pixoff, n_pix, slot_l, alpha_l;  // this lane: tile-relative pixel offset, slot and alpha of the value entering it
        {
            Seg incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Seg o = seg_shfl_up(incl, d);
                if ((int)lane >= d) incl = combine(o, incl);
            }
            Seg excl = seg_shfl_up(incl, 1);
            if (lane == 0) excl = seg_identity();
            const Seg tot = Seg{ __shfl_sync(kFull, incl.cnt, 31), __shfl_sync(kFull, incl.da, 31), __shfl_sync(kFull, incl.fl, 31) };
            n_pix = tot.cnt & 0xFFFFFu;
            // pixels before this tile
            if (lane == 0 && t > 0) st_word(desc + kDwPix, pack_word(n_pix, ST_AGG, epoch));
            // payload of the slot / alpha word: c | root << 6 | uses << 7 | rgba << 8 | alpha << 9
            auto pack = [](const Seg& q) {
                return ((q.fl >> 16) & 63u) | ((q.fl & kFlRoot) ? 64u : 0u) | ((q.fl & kFlUses) ? 128u : 0u) |
                       ((q.fl & kFlRgba) ? 256u : 0u) | (q.da >> 24) << 9;
            };
            auto unpack = [](uint64_t v64) {
                const unsigned v = (unsigned)v64;
                return Seg{ 0u, (v >> 9) << 24, (v & 63u) << 16 | ((v & 64u) ? kFlRoot : 0u) | ((v & 128u) ? kFlUses : 0u) | ((v & 256u) ? kFlRgba : 0u) };
            };
            if (lane == 0 && t > 0) st_word(desc + kDwSlot, pack_word(pack(tot), ST_AGG, epoch));
            pix_base = warp_lookback_lazy<uint64_t>(
                t, (uint64_t)0, (uint64_t)0,
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(word_of(p, kDwPix));
                    st                = status_of(wd);
                    return word_payload(wd);
                },
                [](uint64_t a, uint64_t b) { return a + b; });
            if (lane == 0) {
                const uint64_t total = pix_base + n_pix;
                st_word(desc + kDwPix, pack_word(total < (1ull << 41) ? total : (1ull << 41), ST_INCL, epoch));
            }
            const Seg start = Seg{ 0u, 255u << 24, 53u << 16 | kFlRoot | kFlRgba };  // {0,0,0,255}: slot 53 (simple.cpp:108)
            const Seg acc   = warp_lookback_lazy<Seg>(
                t, start, seg_identity(),
                [&](unsigned p, unsigned& st) {
                    const uint64_t wd = ld_word(word_of(p, kDwSlot));
                    st                = status_of(wd);
                    return unpack(word_payload(wd));
                },
                [](const Seg& a, const Seg& b) { return combine(a, b); });
            if (lane == 0) st_word(desc + kDwSlot, pack_word(pack(combine(acc, tot)), ST_INCL, epoch));
            const unsigned slot_in = (acc.fl >> 16) & 63u, alpha_in = acc.da >> 24;  // concrete: the chain ended in an inclusive word
            pixoff  = excl.cnt & 0xFFFFFu;
            alpha_l = (excl.fl & kFlRgba) ? excl.da >> 24 : alpha_in;
            slot_l  = (excl.fl & kFlRoot) ? (((excl.fl >> 16) & 63u) + ((excl.fl & kFlUses) ? 11u * alpha_in : 0u)) & 63u
                                          : (slot_in + ((excl.fl >> 16) & 63u)) & 63u;
        }

        QB_STAMP(desc, 69, 0, qb_t0);  // W1 + look-backs
        // ================= W2: symbolic walk, the lane's stores per slot =================
        unsigned char* const myc = sm.finc + lane * kDtRowC;  // my column of codes
        unsigned* const      myv = sm.finv + lane;             // value of slot s at myv[s * kDtRowV]
        {
            unsigned* c32 = reinterpret_cast<unsigned*>(myc);
#pragma unroll
            for (int j = 0; j < 16; ++j) c32[j] = 0xFFFFFFFFu;  // kCNone
        }
        unsigned pc = kCPrev, pv = 0, n_ext = 0;
        {
            unsigned alpha = alpha_l, slot = slot_l;
            dt_walk(words, byte0, mlo, mhi, [&](const DtOp& o) {
                const unsigned d   = (o.tag >> 6) == 1 ? dt_diff_delta(o.tag) : dt_luma_delta(o.tag, o.pay);
                const unsigned lit = o.rgba ? o.pay : (o.pay & 0xFFFFFFu) | alpha << 24;  // OP_RGB: speculated alpha, verified in W3
                // OP_INDEX: what the column holds for the slot (an entry of this lane, a cached external read, or nothing yet)
                const unsigned s  = o.tag & 63u;
                const unsigned ic = myc[s], iv = myv[s * kDtRowV];
                const bool     ext = o.index && ic == kCNone;  // external read: what the table held in slot s when this lane began
                if (ext) {
                    myc[s] = (unsigned char)(kCCached | s), myv[s * kDtRowV] = 0u;
                    if (n_ext < (unsigned)kDtExt) sm.extc[n_ext * 32 + lane] = (unsigned char)s;
                    ++n_ext;
                }
                const bool own = o.index && ic < kCCached;  // stored by this lane before
                const bool lit_op = o.rgb || o.rgba;
                pc    = lit_op ? kCConst : (o.index ? (own ? ic : s) : pc);
                pv    = lit_op ? lit : (o.index ? (own ? iv : 0u) : (o.delta ? add4(pv, d) : pv));
                alpha = o.rgba ? o.pay >> 24 : alpha;
                slot  = lit_op ? slot_of(lit) : (o.index ? s : (o.delta ? (slot + dt_lin(d)) & 63u : slot));
                if (lit_op || o.delta) myc[slot] = (unsigned char)pc, myv[slot * kDtRowV] = pv;  // OP_INDEX / OP_RUN store nothing new
            });
        }
        if (n_ext > (unsigned)kDtExt) bad = true;
        __syncwarp();

        QB_STAMP(desc, 69, 1, qb_t0);  // W2
        // ================= merge: last storing lane per slot; incoming prev of every lane =================
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned s   = 32u * h + lane;
            unsigned       cur = kCNone;
#pragma unroll
            for (int t0 = 0; t0 < 32; t0 += 8) {
                unsigned c[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = sm.finc[(t0 + i) * kDtRowC + s];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sm.lw[(t0 + i) * kDtRowC + s] = (unsigned char)cur;
                    if (c[i] < kCCached) cur = t0 + i;
                }
            }
            sm.tlw[s] = (unsigned char)cur;
        }
        DtSym po_tile;  // prev on exit from the tile, lane relative
        {
            DtSym incl = DtSym{ pc, lane, pv };
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const DtSym o = shfl_up_T(incl, d);
                if ((int)lane >= d && incl.code == kCPrev) incl = DtSym{ o.code, o.lane, add4(o.val, incl.val) };
            }
            DtSym excl = shfl_up_T(incl, 1);
            if (lane == 0) excl = DtSym{ kCPrev, 0u, 0u };
            sm.pic[lane] = (unsigned char)excl.code, sm.pio[lane] = (unsigned char)excl.lane, sm.piv[lane] = excl.val;
            po_tile = shfl_T(incl, 31);
        }
        __syncwarp();

        QB_STAMP(desc, 70, 0, qb_t0);  // merge
        // ================= chase: lane-relative bases -> constants or entries of the state entering the tile =================
        // returns code 0..63 (tile-in slot), kCPrev (tile-in prev) or kCConst
        auto chase = [&](DtSym v) {
            for (int hop = 0; hop < kDtHops; ++hop) {
                if (v.code == kCConst) return v;
                if (v.code == kCPrev) {  // incoming prev of lane v.lane
                    if (v.lane == 0) return v;
                    const unsigned c = sm.pic[v.lane];
                    v = DtSym{ c, sm.pio[v.lane], add4(sm.piv[v.lane], v.val) };
                    if (c == kCPrev) return DtSym{ kCPrev, 0u, v.val };  // the chain reaches the tile's incoming prev
                } else {  // what lane v.lane read from slot v.code before storing to it
                    const unsigned w = sm.lw[v.lane * kDtRowC + v.code];
                    if (w == kCNone) return DtSym{ v.code, 0u, v.val };  // nobody in this tile before: the tile's incoming table
                    v = DtSym{ sm.finc[w * kDtRowC + v.code], w, add4(sm.finv[v.code * kDtRowV + w], v.val) };
                }
            }
            bad = true;
            return DtSym{ kCConst, 0u, 0u };
        };
        // the tile's transfer function: 64 slots and prev (entry 64, lane 0)
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const unsigned e = 32u * h + lane;
            if (h < 2 || lane == 0) {
                DtSym f;
                if (h == 2) f = chase(po_tile);
                else {
                    const unsigned w = sm.tlw[e];
                    f = w == kCNone ? DtSym{ e, 0u, 0u } : chase(DtSym{ sm.finc[w * kDtRowC + e], w, sm.finv[e * kDtRowV + w] });
                }
                sm.fcode[e] = f.code, sm.fadd[e] = f.val;
                // payload: add | entry code << 32 (0..63 slot, 64 prev, 65 constant); a constant is already inclusive
                if (f.code != kCConst) st_word(desc + kDwState + e, pack_word((uint64_t)f.code << 32 | f.val, ST_AGG, epoch));
                else st_word(desc + kDwState + e, pack_word(f.val, ST_INCL, epoch));
            }
        }
        // this lane's incoming prev and external reads, tile relative
        DtSym my_prev = chase(DtSym{ kCPrev, lane, 0u });
        {
            const unsigned ne = min(n_ext, (unsigned)kDtExt);
            for (unsigned k = 0; k < ne; ++k) {
                const DtSym x = chase(DtSym{ sm.extc[k * 32 + lane], lane, 0u });
                sm.extc[k * 32 + lane] = (unsigned char)x.code, sm.extv[k * 32 + lane] = x.val;
            }
        }
        QB_STAMP(desc, 70, 1, qb_t0);  // chase + publish
        // ================= state look-back: concrete state entering the tile =================
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const unsigned e0 = 32u * h + lane;
            if (h < 2 || lane == 0) {
                unsigned e = e0, acc = 0, v;
                for (int p = (int)t - 1;; --p) {
                    if (p < 0) {  // simple.cpp:103-108: zero table, prev = start, start stored at its slot
                        v = add4((e == 64 || e == 53) ? kStartPixel : 0u, acc);
                        break;
                    }
                    const uint64_t wd = wait_word(word_of((unsigned)p, kDwState + (int)e), epoch);
                    const uint64_t pl = word_payload(wd);
                    if (status_of(wd) == ST_INCL) { v = add4((unsigned)pl, acc); break; }
                    acc = add4(acc, (unsigned)pl);
                    const unsigned c = (unsigned)(pl >> 32);
                    if (c == 65u) { v = acc; break; }
                    e = c;
                }
                sm.tin[e0] = v;
            }
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 3; ++h) {  // concrete state leaving the tile
            const unsigned e = 32u * h + lane;
            if (h < 2 || lane == 0) {
                const unsigned c = sm.fcode[e], a = sm.fadd[e];
                st_word(desc + kDwState + e, pack_word(c == kCConst ? a : add4(sm.tin[c], a), ST_INCL, epoch));
            }
        }
        auto concrete = [&](unsigned code, unsigned val) { return code == kCConst ? val : add4(sm.tin[code], val); };

        QB_STAMP(desc, 71, 0, qb_t0);  // state look-back
        // ================= W3: the reference loop with concrete values =================
        {
            unsigned* c32 = reinterpret_cast<unsigned*>(myc);
#pragma unroll
            for (int j = 0; j < 16; ++j) c32[j] = 0xFFFFFFFFu;
        }
        {
            unsigned prev = concrete(my_prev.code, my_prev.val), alpha = alpha_l, k_ext = 0;
            uint64_t pix = pix_base + pixoff;
            const unsigned tgt = D.target;
            const bool     w32 = (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
            // pixels leave in aligned groups of four (three or four 32-bit stores); q0..q2 hold the group under construction,
            // `gstart` is the position in its group of the first pixel this lane owns
            unsigned q0 = 0, q1 = 0, q2 = 0, gstart = (unsigned)pix & 3u;
            auto store1 = [&](uint64_t i, unsigned v) {
                if (tgt == 4 && w32) reinterpret_cast<unsigned*>(out)[i] = v;
                else {
                    uint8_t* d = out + i * tgt;
                    d[0] = (uint8_t)v, d[1] = (uint8_t)(v >> 8), d[2] = (uint8_t)(v >> 16);
                    if (tgt == 4) d[3] = (uint8_t)(v >> 24);
                }
            };
            auto put = [&](unsigned cur) {
                const unsigned ph = (unsigned)pix & 3u;
                if (ph == 3u) {
                    if (gstart == 0u && w32) {
                        if (tgt == 4) {
                            unsigned* d = reinterpret_cast<unsigned*>(out) + (pix - 3u);
                            d[0] = q0, d[1] = q1, d[2] = q2, d[3] = cur;
                        } else {
                            unsigned* d = reinterpret_cast<unsigned*>(out + (pix - 3u) * 3u);
                            d[0] = (q0 & 0xFFFFFFu) | q1 << 24;
                            d[1] = ((q1 >> 8) & 0xFFFFu) | q2 << 16;
                            d[2] = ((q2 >> 16) & 0xFFu) | cur << 8;
                        }
                    } else {
                        for (unsigned j = gstart; j < 3u; ++j) store1(pix - 3u + j, j == 0 ? q0 : (j == 1 ? q1 : q2));
                        store1(pix, cur);
                    }
                    gstart = 0u;
                } else {
                    q0 = ph == 0u ? cur : q0, q1 = ph == 1u ? cur : q1, q2 = ph == 2u ? cur : q2;
                }
                ++pix;
            };
            dt_walk(words, byte0, mlo, mhi, [&](const DtOp& o) {
                if (pix >= N) return;  // ops past the image are never executed by the reference
                const unsigned d = (o.tag >> 6) == 1 ? dt_diff_delta(o.tag) : dt_luma_delta(o.tag, o.pay);
                const unsigned s = o.tag & 63u;
                const unsigned ic = myc[s], iv = myv[s * kDtRowV];
                const bool     ext = o.index && ic == kCNone;
                unsigned       xv = 0;
                if (ext) {
                    xv = k_ext < (unsigned)kDtExt ? concrete(sm.extc[k_ext * 32 + lane], sm.extv[k_ext * 32 + lane]) : 0u;
                    ++k_ext;
                    myc[s] = 0, myv[s * kDtRowV] = xv;
                }
                // simple.cpp:119-123: an OP_RGB inherits the alpha of the previous pixel
                const unsigned cur = o.rgb ? (o.pay & 0xFFFFFFu) | (prev & 0xFF000000u)
                                   : o.rgba ? o.pay
                                   : o.index ? (ext ? xv : iv)
                                   : o.delta ? add4(prev, d) : prev;
                bad |= o.rgb && (prev >> 24) != alpha;     // the alpha W2 assumed for this literal
                bad |= o.index && slot_of(cur) != s;       // a never-stored (or mis-predicted) slot was read
                alpha = o.rgba ? o.pay >> 24 : alpha;
                if (o.rgb || o.rgba || o.delta) {
                    const unsigned ws = slot_of(cur);
                    myc[ws] = 0, myv[ws * kDtRowV] = cur;  // simple.cpp:169
                }
                const unsigned n = o.run ? (o.tag & 63u) + 1u : 1u;
                for (unsigned j = 0; j < n && pix < N; ++j) put(cur);  // OP_RUN clamped to the image (simple.cpp:158)
                prev = cur;
            });
            // pixels of an unfinished group
            for (unsigned j = gstart; j < ((unsigned)pix & 3u); ++j) store1((pix & ~3ull) + j, j == 0 ? q0 : (j == 1 ? q1 : q2));
        }
        // the stream ended before the image (the reference decodes the zero padding on): general path
        if (t == ntiles - 1 && pix_base + n_pix < N) bad = true;
        if (t == ntiles - 1 && lane == 0) res->pixels = pix_base + n_pix < N ? pix_base + n_pix : N;
        if (__any_sync(kFull, bad) && lane == 0) {
            res->pad[0]             = 1;  // fast path refuted for this image
            D.control->fast_any_bad = 1;
        }
        __syncwarp();
    }

    // persistent, independent warps draw tiles from a ticket counter in start order (every tile a running warp waits for
    // is held by a warp that is running too or done); the counter never resets, the host passes its value at launch
    __global__ void __launch_bounds__(kDtThreads, QB_DT_CTAS) decode_ts_kernel(const DtParams P)
    {
        DtWarpSmem&    sm   = reinterpret_cast<DtWarpSmem*>(QB_DYN_SMEM)[threadIdx.x >> 5];
        const unsigned lane = threadIdx.x & 31u;
        for (;;) {
            unsigned x = 0;
            if (lane == 0) x = atomicAdd(P.ticket, 1u) - P.ticket_base;
            x = __shfl_sync(kFull, x, 0);
            if (x >= P.n_tiles) break;
            dt_decode_tile(P, sm, x);
            __syncwarp();
        }
    }
}  // namespace qb
