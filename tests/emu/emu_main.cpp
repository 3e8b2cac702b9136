// emu_main.cpp -- TEST INFRASTRUCTURE ONLY: runs the product kernels of qoipp_b200/csrc under the CPU SIMT
// emulator (cuda_emu.h) so their logic can be checked against the oracle without a GPU.  Built by
// tests/emu/Makefile into tests/emu/libqb_emu.so; never part of libqoipp_b200.so.
#include "cuda_emu.h"

#include "../../qoipp_b200/csrc/decode_wt.cuh"
#include "../../qoipp_b200/csrc/encode_kernel.cuh"
#include "../../qoipp_b200/csrc/encode_ts.cuh"
#include "../../qoipp_b200/csrc/host_util.hpp"

#include <algorithm>
#include <vector>

using namespace qb;

namespace
{
    template <int CH, int K>
    void run_encode(const EncParams& P, int resident, uint64_t seed)
    {
        emu::launch(dim3(P.tiles_per_image * P.n_images), dim3(kEncThreads), sizeof(EncSmem<K>) + 128,
                    [=] { encode_kernel<CH, K>(P); }, resident, seed);
    }

    template <int CH>
    void run_encode_ts(const EncParams& P, int resident, uint64_t seed)
    {
        EncParams Q = P;
        const unsigned n_tiles = P.tiles_per_image * P.n_images, groups = (P.tiles_per_image + 63) / 64;
        std::vector<uint32_t> scratch((size_t)n_tiles * TsCfg<CH>::kScrWords, 0xDEADBEEFu), counts(n_tiles + (size_t)groups * P.n_images, 0);
        uint32_t ticket = 0xFFFFFFF0u + (uint32_t)seed;  // the counter never resets (and may wrap)
        Q.scratch = scratch.data(); Q.tile_bytes = counts.data(); Q.group_bytes = counts.data() + n_tiles; Q.groups_per_image = groups;
        Q.ticket = &ticket; Q.ticket_base[0] = ticket;
        const unsigned n_ctas = std::min<unsigned>((n_tiles + kTsWarps - 1) / kTsWarps, (unsigned)resident);
        emu::launch(dim3(n_ctas), dim3(kTsThreads), kTsWarps * sizeof(TsWarpSmem) + 128, [=] { encode_ts_kernel<CH>(Q); }, resident, seed);
        emu::launch(dim3((n_tiles + kTsCopyWarps - 1) / kTsCopyWarps), dim3(kTsCopyWarps * 32), kTsCopyWarps * sizeof(TsCopySmem) + 128,
                    [=] { encode_ts_copy_kernel<CH>(Q); }, resident, seed);
    }

    // K == kTsK selects the thread-serial kernel (encode_ts.cuh), any other K the general kernel with K pixels per lane
    uint64_t tile_pixels(int K) { return K == kTsK ? (uint64_t)kTsT : (uint64_t)kEncThreads * K; }

    void dispatch_encode(const EncParams& P, int ch, int K, int resident, uint64_t seed)
    {
        if (K == kTsK) return ch == 3 ? run_encode_ts<3>(P, resident, seed) : run_encode_ts<4>(P, resident, seed);
#define QB_CASE(CHV, KV) if (ch == CHV && K == KV) return run_encode<CHV, KV>(P, resident, seed);
        QB_CASE(3, 1) QB_CASE(4, 1) QB_CASE(3, 2) QB_CASE(4, 2) QB_CASE(3, 8) QB_CASE(4, 8)
#undef QB_CASE
        fprintf(stderr, "emu: unsupported ch=%d K=%d\n", ch, K);
        abort();
    }

    unsigned pack(const uint8_t* p) { return p[0] | p[1] << 8 | p[2] << 16 | (unsigned)p[3] << 24; }
    void     unpack(unsigned v, uint8_t* p) { p[0] = v, p[1] = v >> 8, p[2] = v >> 16, p[3] = v >> 24; }
}

extern "C"
{
    // one-shot / batch encode of n_images equally shaped images
    int emu_encode(const uint8_t* raw, uint64_t raw_stride, uint32_t n_images, uint32_t w, uint32_t h, uint8_t ch, uint8_t cs,
                   uint8_t* out, uint64_t out_stride, uint64_t cap, uint64_t* written, int* complete, int K, int resident,
                   uint64_t seed)
    {
        qoipp_b200_desc d{ w, h, ch, cs };
        if (cap < host::kHeaderSize) { for (uint32_t i = 0; i < n_images; ++i) written[i] = 0, complete[i] = 0; return 0; }
        EncParams P{};
        P.in = raw; P.out = out;
        P.n_pixels = (uint64_t)w * h;
        P.in_stride = raw_stride; P.out_stride = out_stride; P.out_cap = cap;
        const uint64_t T = tile_pixels(K);
        P.tiles_per_image = (uint32_t)((P.n_pixels + T - 1) / T);
        P.n_images = n_images; P.epoch = 7; P.flags = 0;
        host::write_header(d, P.header);
        std::vector<uint64_t>  desc((size_t)P.tiles_per_image * n_images * kEncDescWords, 0);
        std::vector<EncResult> res(n_images);
        uint32_t               ticket[2] = { 0, 0 };
        P.desc = desc.data(); P.results = res.data(); P.ticket = ticket; P.init_state = nullptr;
        dispatch_encode(P, ch, K, resident, seed);
        if (ticket[0] != 0 || ticket[1] != 0) return -1;
        for (uint32_t i = 0; i < n_images; ++i) written[i] = res[i].written, complete[i] = (int)res[i].complete;
        return 0;
    }

    int emu_stream_encode(qoipp_b200_state* st, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t cap,
                          uint64_t* processed, uint64_t* written, int K, int resident, uint64_t seed)
    {
        const unsigned ch = st->channels;
        const uint64_t n  = in_size / ch;
        if (n == 0) { *processed = 0; *written = 0; return 0; }
        EncState is{};
        is.prev = pack(st->prev); is.run = st->run;
        for (int s = 0; s < 64; ++s) is.table[s] = pack(st->seen[s]);
        EncParams P{};
        P.in = in; P.out = out; P.n_pixels = n; P.in_stride = 0; P.out_stride = 0; P.out_cap = cap;
        const uint64_t T = (uint64_t)kEncThreads * K;
        P.tiles_per_image = (uint32_t)((n + T - 1) / T);
        P.n_images = 1; P.epoch = 9; P.flags = ENC_STREAM;
        std::vector<uint64_t> desc((size_t)P.tiles_per_image * kEncDescWords, 0);
        EncResult             res{};
        uint32_t              ticket = 0;
        P.desc = desc.data(); P.results = &res; P.ticket = &ticket; P.init_state = &is;
        dispatch_encode(P, (int)ch, K, resident, seed);
        *processed = res.processed * ch;
        *written   = res.written;
        unpack(res.state.prev, st->prev);
        st->run = (uint8_t)res.state.run;
        for (int s = 0; s < 64; ++s) unpack(res.state.table[s], st->seen[s]);
        return 0;
    }

    // one-shot / batch decode.  offsets has n_images + 1 entries (streams include their headers).
    // out_path[k]: 0 = the parallel kernel's pixels stand, 1 = the sequential kernel re-decoded image k.
    int emu_decode(const uint8_t* qoi, const uint64_t* offsets, uint32_t n_images, uint32_t w, uint32_t h, uint8_t target,
                   int flip, uint8_t* out, uint64_t out_stride, int* out_path, int force_serial, int resident, uint64_t seed)
    {
        std::vector<uint32_t> tile_first(n_images + 1);
        uint64_t              tiles = 0;
        for (uint32_t k = 0; k < n_images; ++k) {
            tile_first[k] = (uint32_t)tiles;
            tiles += (offsets[k + 1] - offsets[k] - host::kHeaderSize + kDecTB - 1) / kDecTB;
        }
        tile_first[n_images] = (uint32_t)tiles;
        DecParams P{};
        P.qoi = qoi;
        if (n_images == 1) { P.offsets = nullptr; P.tile_first = nullptr; P.single[0] = offsets[0]; P.single[1] = offsets[1]; }
        std::vector<uint64_t> pairs;  // the kernels take {first byte, one past the last} per stream
        for (uint32_t k = 0; k < n_images; ++k) pairs.push_back(offsets[k]), pairs.push_back(offsets[k + 1]);
        if (n_images != 1) { P.offsets = pairs.data(); P.tile_first = tile_first.data(); }
        P.out = out; P.out_stride = out_stride; P.n_pixels = (uint64_t)w * h;
        P.width = w; P.height = h; P.target = target; P.flip = flip;
        P.n_images = n_images; P.n_tiles = (uint32_t)tiles; P.epoch = 3;
        std::vector<uint64_t>  desc((size_t)tiles * kDecDescWords, 0);
        std::vector<uint32_t>  fix((size_t)tiles * kFixWords, 0xDEADBEEFu);
        std::vector<DecResult> res(n_images);
        DecControl             ctl;
        memset(res.data(), 0, sizeof(DecResult) * n_images);
        memset(&ctl, 0, sizeof ctl);
        std::vector<CascadeReq> reqs(64);
        P.desc = desc.data(); P.results = res.data(); P.control = &ctl; P.fix = fix.data(); P.req = reqs.data(); P.req_cap = (uint32_t)reqs.size();
        P.epoch = 5; P.round = 0;
        const unsigned n_ctas = std::max(1u, std::min<unsigned>((P.n_tiles + kWtWarps - 1) / kWtWarps, (unsigned)resident));
        if (!force_serial) {
            emu::launch(dim3(n_ctas), dim3(kWtThreads), kWtSmemBytes + 128, [=] { decode_wt_kernel(P); }, (int)n_ctas, seed);
        } else {
            for (auto& r : res) r.first_bad[kDecRounds] = 0xFFFFFFFFu;  // "tile 0 refuted in the last round"
            ctl.any_bad[kDecRounds] = 1;
        }
        {   // cooperative launch: every CTA resident
            emu::launch(dim3(n_ctas), dim3(kWtThreads), kWtSmemBytes + 128, [=] { decode_finish_kernel(P); }, (int)n_ctas, seed + 1);
        }
        if (getenv("QB_EMU_DEBUG"))
            for (uint32_t k = 0; k < n_images; ++k)
                fprintf(stderr, "img %u: bad %u path %u first_bad %x %x %x %x %x pixels %llu\n", k, res[k].bad, res[k].path, res[k].first_bad[0],
                        res[k].first_bad[1], res[k].first_bad[2], res[k].first_bad[3], res[k].first_bad[4], (unsigned long long)res[k].pixels);
        for (uint32_t k = 0; k < n_images; ++k) out_path[k] = (int)res[k].path;
        return 0;
    }

    // resumable decode.  parallel != 0: decode_wt_stream_kernel first (whatever the input size), the sequential loop only
    // when that kernel refuted a speculation -- the launch sequence of qoipp_b200_stream_decode_dev.  *used_serial reports which.
    int emu_stream_decode(qoipp_b200_state* st, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t cap,
                          uint64_t* processed, uint64_t* written, int parallel, int resident, uint64_t seed, int* used_serial)
    {
        DecState is{};
        is.prev = pack(st->prev); is.run = st->run;
        for (int s = 0; s < 64; ++s) is.table[s] = pack(st->seen[s]);
        const unsigned ch = st->channels;
        DecResult    res{};
        DecControl   ctl;
        memset(&ctl, 0, sizeof ctl);
        SerialParams S{};
        const uint64_t room = cap / ch;
        // the kernels stage whole 16-byte aligned vectors: up to 15 bytes around the buffer are read (never across a page on
        // the device); give the host copy that slack, at an odd alignment
        std::vector<uint8_t> padded(in_size + 64, 0xEE);
        memcpy(padded.data() + 19, in, in_size);
        in = padded.data() + 19;
        S.d.qoi = in; S.d.single[0] = 0; S.d.single[1] = in_size; S.d.out = out; S.d.out_stride = room * ch;
        S.d.target = ch; S.d.flip = 0; S.d.n_images = 1; S.d.results = &res; S.d.n_pixels = room;
        S.mode = 1; S.init = &is; S.in_size = in_size; S.only_if_bad = 0;
        const uint64_t tiles = (in_size + kDecTB - 1) / kDecTB;
        std::vector<uint64_t> desc((size_t)tiles * kDecDescWords + 1, 0);
        std::vector<uint32_t> fix((size_t)tiles * kFixWords + 1, 0xDEADBEEFu);
        std::vector<CascadeReq> reqs(64);
        if (parallel && tiles > 0) {
            DecParams& P = S.d;
            P.offsets = nullptr; P.tile_first = nullptr; P.n_tiles = (uint32_t)tiles; P.epoch = 7; P.round = 0;
            P.control = &ctl; P.desc = desc.data(); P.fix = fix.data(); P.init = &is; P.req = reqs.data(); P.req_cap = (uint32_t)reqs.size();
            const unsigned n_ctas = std::max(1u, std::min<unsigned>((P.n_tiles + kWtWarps - 1) / kWtWarps, (unsigned)resident));
            const DecParams PP = P;
            emu::launch(dim3(n_ctas), dim3(kWtThreads), kWtSmemBytes + 128, [=] { decode_wt_stream_kernel(PP); }, (int)n_ctas, seed);
            emu::launch(dim3(n_ctas), dim3(kWtThreads), kWtSmemBytes + 128, [=] { decode_finish_stream_kernel(PP); }, (int)n_ctas, seed + 1);
            S.only_if_bad = 1;
        }
        if (used_serial) *used_serial = !(parallel && tiles > 0) || res.first_bad[kDecRounds] != 0;
        emu::launch(dim3(1), dim3(32), sizeof(SerialSmem) + 128, [=] { decode_serial_kernel(S); }, 1, 0);
        *processed = res.processed; *written = res.written;
        unpack(res.state.prev, st->prev);
        st->run = (uint8_t)res.state.run;
        for (int s = 0; s < 64; ++s) unpack(res.state.table[s], st->seen[s]);
        return 0;
    }

    uint64_t emu_switch_count() { return emu::ctx().switches; }
}
