// emu_main.cpp -- TEST INFRASTRUCTURE ONLY: runs the product kernels of qoipp_b200/csrc under the CPU SIMT
// emulator (cuda_emu.h) so their logic can be checked against the oracle without a GPU.  Built by
// tests/emu/Makefile into tests/emu/libqb_emu.so; never part of libqoipp_b200.so.
#include "cuda_emu.h"

#include "../../qoipp_b200/csrc/encode_kernel.cuh"
#include "../../qoipp_b200/csrc/host_util.hpp"

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

    void dispatch_encode(const EncParams& P, int ch, int K, int resident, uint64_t seed)
    {
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
        const uint64_t T = (uint64_t)kEncThreads * K;
        P.tiles_per_image = (uint32_t)((P.n_pixels + T - 1) / T);
        P.n_images = n_images; P.epoch = 7; P.flags = 0;
        host::write_header(d, P.header);
        std::vector<uint64_t>  desc((size_t)P.tiles_per_image * n_images * kEncDescWords, 0);
        std::vector<EncResult> res(n_images);
        uint32_t               ticket = 0;
        P.desc = desc.data(); P.results = res.data(); P.ticket = &ticket; P.init_state = nullptr;
        dispatch_encode(P, ch, K, resident, seed);
        if (ticket != 0) return -1;
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

    uint64_t emu_switch_count() { return emu::ctx().switches; }
}
