// decode_host.inl -- decode entry points of the C ABI (included at the end of qoipp_b200.cu).
namespace
{
    // tile bookkeeping shared by the single-image and batch launches
    constexpr size_t kCtrlBytes = 64;  // DecControl lives in front of the DecResult array
    static_assert(sizeof(DecControl) <= kCtrlBytes, "control block");

    int32_t launch_decode(qoipp_b200_ctx* c, DecParams& P, cudaStream_t s)
    {
        QB_CUDA(set_attrs(c));
        if (P.n_pixels >= kPixSat) return H::TooBig;  // the pixel carry word holds 33 bits
        const size_t res_bytes = kCtrlBytes + sizeof(DecResult) * P.n_images;
        QB_CUDA(c->results.reserve(res_bytes, s));
        QB_CUDA(cudaMemsetAsync(c->results.p, 0, res_bytes, s));
        P.req_cap = (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(64, P.n_tiles / 64));  // cascade requests are rare: one in ~10^4 tiles
        QB_CUDA(c->reqs.reserve(sizeof(CascadeReq) * P.req_cap, s));
        P.req = static_cast<CascadeReq*>(c->reqs.p);
        QB_CUDA(c->fix.reserve((size_t)P.n_tiles * kFixWords * sizeof(uint32_t), s, true));
        // one epoch per possible round; the learned-alpha lists are tagged with the first one
        QB_CUDA(c->next_epoch((uint64_t)P.n_tiles * kDecDescWords * sizeof(uint64_t), s, kDecRounds + 1));
        P.epoch   = c->epoch;
        P.round   = 0;
        P.control = static_cast<DecControl*>(c->results.p);
        P.results = reinterpret_cast<DecResult*>(static_cast<uint8_t*>(c->results.p) + kCtrlBytes);
        P.desc    = static_cast<uint64_t*>(c->carry.p);
        P.fix     = static_cast<uint32_t*>(c->fix.p);
        // persistent warps: one CTA per resident slot
        const unsigned want   = (P.n_tiles + kWtWarps - 1) / kWtWarps;
        const unsigned n_ctas = std::max(1u, std::min<unsigned>(want, (unsigned)c->dec_coresident));
        decode_wt_kernel<<<n_ctas, kWtThreads, kWtSmemBytes, s>>>(P);
        QB_CUDA(cudaGetLastError());
        // Retry rounds and the sequential part are ONE cooperative launch enqueued unconditionally: it returns at once when
        // round 0 verified every tile -- no host round trip, a few microseconds when nothing is to be done.
        void* args[] = { &P };
        QB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(decode_finish_kernel), dim3(n_ctas), dim3(kWtThreads), args,
                                            kWtSmemBytes, s));
        return 0;
    }

    // batch decode: stream k is d_qoi[start(k), start(k) + size(k))
    template <class Span>
    int32_t decode_batch_impl(qoipp_b200_ctx* c, const uint8_t* d_qoi, uint32_t n_images, Span span, const qoipp_b200_desc* desc,
                                     uint8_t target, uint8_t* d_out, uint64_t out_stride, cudaStream_t s)
    {
        if (n_images == 0) return H::Empty;
        uint64_t raw;
        if (int32_t e = H::count_bytes(*desc, &raw)) return e;
        const unsigned tgt = target ? target : desc->channels;
        if (tgt != 3 && tgt != 4) return H::InvalidDesc;
        const uint64_t n = (uint64_t)desc->width * desc->height;
        if (out_stride < n * tgt) return H::NotEnoughSpace;
        Guard g(c->device);
        // per-image table: {first byte, one past the last} (u64 pairs) then first-tile ids (u32).  The table is built in
        // ordinary host memory and compared with the one the device already holds: a repeated decode of the same batch
        // layout (the steady state of a pipeline, and of bench.py) uploads nothing and does not touch the stream.  A new
        // layout is staged through pinned memory; only then the stream is drained first (the pinned copy of the previous
        // table may be in flight).
        const size_t off_bytes = sizeof(uint64_t) * 2 * (size_t)n_images, tf_bytes = sizeof(uint32_t) * ((size_t)n_images + 2);
        std::vector<uint64_t>& tab = c->batch_table;
        tab.assign((off_bytes + tf_bytes + 7) / 8, 0);
        auto*    ho = tab.data();
        auto*    ht = reinterpret_cast<uint32_t*>(ho + 2 * (size_t)n_images);
        uint64_t tiles = 0;
        for (uint32_t k = 0; k < n_images; ++k) {
            uint64_t start, sz;
            span(k, start, sz);
            if (sz <= H::kHeaderSize + H::kMarkerSize) return sz == 0 ? H::Empty : H::TooShort;
            ho[2 * k] = start, ho[2 * k + 1] = start + sz, ht[k] = (uint32_t)tiles;
            tiles += (sz - H::kHeaderSize + kDecTB - 1) / kDecTB;
            if (tiles >= (1ull << 31)) return H::TooBig;
        }
        ht[n_images] = (uint32_t)tiles;
        const bool same = c->batch_table_dev_bytes == off_bytes + tf_bytes && c->batch_table_stream == s && c->aux.p && c->h_pin_in.p &&
                          std::memcmp(c->h_pin_in.p, ho, off_bytes + tf_bytes) == 0;
        if (!same) {
            c->batch_table_dev_bytes = 0;
            QB_CUDA(cudaStreamSynchronize(s));
            QB_CUDA(c->h_pin_in.reserve(off_bytes + tf_bytes));
            QB_CUDA(c->aux.reserve(off_bytes + tf_bytes, s));
            std::memcpy(c->h_pin_in.p, ho, off_bytes + tf_bytes);
            QB_CUDA(cudaMemcpyAsync(c->aux.p, c->h_pin_in.p, off_bytes + tf_bytes, cudaMemcpyHostToDevice, s));
            c->batch_table_dev_bytes = off_bytes + tf_bytes, c->batch_table_stream = s;
        }
        DecParams P{};
        P.qoi = d_qoi;
        P.offsets    = static_cast<uint64_t*>(c->aux.p);
        P.tile_first = reinterpret_cast<uint32_t*>(static_cast<uint64_t*>(c->aux.p) + 2 * (size_t)n_images);
        P.out = d_out, P.out_stride = out_stride, P.n_pixels = n;
        P.width = desc->width, P.height = desc->height, P.target = tgt, P.flip = 0;
        P.n_images = n_images, P.n_tiles = (uint32_t)tiles;
        return launch_decode(c, P, s);
    }
}

extern "C"
{
    int32_t qoipp_b200_decode_dev(qoipp_b200_ctx* c, const uint8_t* d_qoi, uint64_t qoi_size, const qoipp_b200_desc* desc,
                                  uint8_t target, int32_t flip, uint8_t* d_out, uint64_t out_cap, void* stream)
    {
        if (qoi_size == 0) return H::Empty;
        if (qoi_size <= H::kHeaderSize + H::kMarkerSize) return H::TooShort;  // simple.cpp:369
        uint64_t raw;
        if (int32_t e = H::count_bytes(*desc, &raw)) return e;
        const unsigned tgt = target ? target : desc->channels;
        if (tgt != 3 && tgt != 4) return H::InvalidDesc;
        const uint64_t n = (uint64_t)desc->width * desc->height;
        if (out_cap < n * tgt) return H::NotEnoughSpace;
        const uint64_t tiles = (qoi_size - H::kHeaderSize + kDecTB - 1) / kDecTB;
        if (tiles >= (1ull << 31)) return H::TooBig;
        Guard     g(c->device);
        DecParams P{};
        P.qoi = d_qoi, P.offsets = nullptr, P.tile_first = nullptr;
        P.single[0] = 0, P.single[1] = qoi_size;
        P.out = d_out, P.out_stride = 0, P.n_pixels = n;
        P.width = desc->width, P.height = desc->height, P.target = tgt, P.flip = flip != 0;
        P.n_images = 1, P.n_tiles = (uint32_t)tiles;
        return launch_decode(c, P, static_cast<cudaStream_t>(stream));
    }

    int32_t qoipp_b200_decode_status(qoipp_b200_ctx* c, void* stream, int32_t* path)
    {
        Guard g(c->device);
        auto  s = static_cast<cudaStream_t>(stream);
        auto* h = static_cast<DecResult*>(c->h_result.p);
        QB_CUDA(cudaMemcpyAsync(h, static_cast<uint8_t*>(c->results.p) + kCtrlBytes, 64, cudaMemcpyDeviceToHost, s));
        QB_CUDA(cudaStreamSynchronize(s));
        if (path) *path = (int32_t)h->path;
        return 0;
    }

    int32_t qoipp_b200_decode_status_batch(qoipp_b200_ctx* c, void* stream, int32_t* paths, uint32_t n_images)
    {
        Guard g(c->device);
        auto  s = static_cast<cudaStream_t>(stream);
        QB_CUDA(c->h_pin_out.reserve(sizeof(uint32_t) * (size_t)n_images));
        // DecResult[k].path -> a packed pinned array
        QB_CUDA(cudaMemcpy2DAsync(c->h_pin_out.p, sizeof(uint32_t), static_cast<uint8_t*>(c->results.p) + kCtrlBytes + offsetof(DecResult, path),
                                  sizeof(DecResult), sizeof(uint32_t), n_images, cudaMemcpyDeviceToHost, s));
        QB_CUDA(cudaStreamSynchronize(s));
        std::memcpy(paths, c->h_pin_out.p, sizeof(uint32_t) * (size_t)n_images);
        return 0;
    }

    int32_t qoipp_b200_decode_host(qoipp_b200_ctx* c, const uint8_t* h_qoi, uint64_t qoi_size, uint8_t target, int32_t flip,
                                   uint8_t* h_out, uint64_t out_cap, qoipp_b200_desc* desc)
    {
        // check order of qoipp::decode_into, source/simple.cpp:451-474
        if (qoi_size == 0) return H::Empty;
        if (qoi_size <= H::kHeaderSize + H::kMarkerSize) return H::TooShort;
        if (int32_t e = H::read_header(h_qoi, qoi_size, desc)) return e;
        uint64_t src_bytes;
        if (int32_t e = H::count_bytes(*desc, &src_bytes)) return e;
        if (out_cap < src_bytes) return H::NotEnoughSpace;  // :467-471 sized with the SOURCE channel count
        const unsigned tgt  = target ? target : desc->channels;
        const uint64_t need = (uint64_t)desc->width * desc->height * tgt;
        if (out_cap < need) return H::NotEnoughSpace;  // SURVEY hazard 3: the reference would write past the buffer
        Guard          g(c->device);
        cudaStream_t   s     = c->own_stream;
        const uint8_t* d_in  = mapped_host(h_qoi);  // page-locked buffers are used in place (zero-copy over PCIe)
        uint8_t*       d_out = mapped_host(h_out);
        if (!d_in) {
            QB_CUDA(c->stage_in.reserve(qoi_size + 64, s));
            QB_CUDA(pageable_to_device(c, c->stage_in.p, h_qoi, qoi_size, s));
            d_in = static_cast<uint8_t*>(c->stage_in.p);
        }
        const bool staged_out = d_out == nullptr;
        if (staged_out) {
            QB_CUDA(c->stage_out.reserve(need + 64, s));
            d_out = static_cast<uint8_t*>(c->stage_out.p);
        }
        if (int32_t e = qoipp_b200_decode_dev(c, d_in, qoi_size, desc, (uint8_t)tgt, flip, d_out, need, s)) return e;
        if (staged_out) QB_CUDA(device_to_pageable(c, h_out, c->stage_out.p, need, s));
        else QB_CUDA(cudaStreamSynchronize(s));
        desc->channels = (uint8_t)tgt;  // :476 the returned Desc carries the target
        return 0;
    }

    // decode into the context's staging buffer; returns once the work is enqueued (the caller allocates meanwhile)
    int32_t qoipp_b200_decode_staged(qoipp_b200_ctx* c, const uint8_t* h_qoi, uint64_t qoi_size, uint8_t target, int32_t flip,
                                     qoipp_b200_desc* desc, uint64_t* out_bytes)
    {
        if (qoi_size == 0) return H::Empty;  // check order of qoipp::decode, source/simple.cpp:367-395
        if (qoi_size <= H::kHeaderSize + H::kMarkerSize) return H::TooShort;
        if (int32_t e = H::read_header(h_qoi, qoi_size, desc)) return e;
        uint64_t src_bytes;
        if (int32_t e = H::count_bytes(*desc, &src_bytes)) return e;
        const unsigned tgt  = target ? target : desc->channels;
        const uint64_t need = (uint64_t)desc->width * desc->height * tgt;
        Guard          g(c->device);
        cudaStream_t   s    = c->own_stream;
        const uint8_t* d_in = mapped_host(h_qoi);
        if (!d_in) {
            QB_CUDA(c->stage_in.reserve(qoi_size + 64, s));
            QB_CUDA(pageable_to_device(c, c->stage_in.p, h_qoi, qoi_size, s));
            d_in = static_cast<uint8_t*>(c->stage_in.p);
        }
        QB_CUDA(c->stage_out.reserve(need + 64, s));
        if (int32_t e = qoipp_b200_decode_dev(c, d_in, qoi_size, desc, (uint8_t)tgt, flip, static_cast<uint8_t*>(c->stage_out.p), need, s)) return e;
        desc->channels  = (uint8_t)tgt;
        c->staged_bytes = need, *out_bytes = need;
        return 0;
    }

    int32_t qoipp_b200_decode_batch_dev(qoipp_b200_ctx* c, const uint8_t* d_qoi, const uint64_t* h_offsets, uint32_t n_images,
                                        const qoipp_b200_desc* desc, uint8_t target, uint8_t* d_out, uint64_t out_stride,
                                        void* stream)
    {
        return decode_batch_impl(
            c, d_qoi, n_images,
            [&](uint32_t k, uint64_t& start, uint64_t& sz) { start = h_offsets[k], sz = h_offsets[k + 1] >= h_offsets[k] ? h_offsets[k + 1] - h_offsets[k] : 0; },
            desc, target, d_out, out_stride, static_cast<cudaStream_t>(stream));
    }

    int32_t qoipp_b200_decode_batch_strided_dev(qoipp_b200_ctx* c, const uint8_t* d_qoi, uint64_t in_stride, const uint64_t* h_sizes,
                                                uint32_t n_images, const qoipp_b200_desc* desc, uint8_t target, uint8_t* d_out,
                                                uint64_t out_stride, void* stream)
    {
        return decode_batch_impl(
            c, d_qoi, n_images, [&](uint32_t k, uint64_t& start, uint64_t& sz) { start = (uint64_t)k * in_stride, sz = std::min(h_sizes[k], in_stride); }, desc,
            target, d_out, out_stride, static_cast<cudaStream_t>(stream));
    }

    int32_t qoipp_b200_decode_batch_host(qoipp_b200_ctx* c, const uint8_t* h_qoi, uint64_t in_stride, const uint64_t* h_sizes, uint32_t n_images,
                                         const qoipp_b200_desc* desc, uint8_t target, uint8_t* h_out, uint64_t out_stride)
    {
        if (n_images == 0) return H::Empty;
        uint64_t raw;
        if (int32_t e = H::count_bytes(*desc, &raw)) return e;
        const unsigned tgt  = target ? target : desc->channels;
        if (tgt != 3 && tgt != 4) return H::InvalidDesc;
        const uint64_t need = (uint64_t)desc->width * desc->height * tgt;
        if (out_stride < need) return H::NotEnoughSpace;
        Guard          g(c->device);
        cudaStream_t   s     = c->own_stream;
        const uint64_t out_bytes = out_stride * (n_images - 1) + need;
        const uint8_t* d_in  = mapped_host(h_qoi);
        uint8_t*       d_out = mapped_host(h_out);
        if (!d_in) {  // only the bytes of each stream travel
            QB_CUDA(c->stage_in.reserve(in_stride * n_images + 64, s));
            for (uint32_t k = 0; k < n_images; ++k)
                QB_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(c->stage_in.p) + k * in_stride, h_qoi + k * in_stride, std::min(h_sizes[k], in_stride),
                                        cudaMemcpyHostToDevice, s));
            d_in = static_cast<uint8_t*>(c->stage_in.p);
        }
        const bool staged_out = d_out == nullptr;
        if (staged_out) {
            QB_CUDA(c->stage_out.reserve(out_bytes + 64, s));
            d_out = static_cast<uint8_t*>(c->stage_out.p);
        }
        if (int32_t e = qoipp_b200_decode_batch_strided_dev(c, d_in, in_stride, h_sizes, n_images, desc, (uint8_t)tgt, d_out, out_stride, s)) return e;
        if (staged_out) QB_CUDA(device_to_pageable(c, h_out, c->stage_out.p, out_bytes, s));
        else QB_CUDA(cudaStreamSynchronize(s));
        return 0;
    }

    // ---- resumable decode on device buffers: everything is enqueued on `stream`, nothing waits for the device
    int32_t qoipp_b200_stream_decode_dev(qoipp_b200_ctx* c, uint8_t channels, qoipp_b200_dev_state* d_state, const uint8_t* d_in,
                                         uint64_t in_size, uint8_t* d_out, uint64_t out_cap, qoipp_b200_stream_result* d_result, void* stream)
    {
        if (channels != 3 && channels != 4) return H::NotInitialized;  // error order of StreamDecoder::decode, stream.cpp:314-320
        if (out_cap == 0) return H::Empty;
        if (out_cap < channels) return H::TooShort;
        Guard g(c->device);
        auto  s = static_cast<cudaStream_t>(stream);
        QB_CUDA(set_attrs(c));
        const uint64_t room = out_cap / channels;  // pixels
        if (room >= kPixSat) return H::TooBig;
        const uint64_t tiles = (in_size + kDecTB - 1) / kDecTB;
        if (tiles >= (1ull << 31)) return H::TooBig;
        const size_t res_bytes = kCtrlBytes + sizeof(DecResult);
        QB_CUDA(c->results.reserve(res_bytes, s));
        QB_CUDA(cudaMemsetAsync(c->results.p, 0, res_bytes, s));
        auto* d_res = reinterpret_cast<DecResult*>(static_cast<uint8_t*>(c->results.p) + kCtrlBytes);
        SerialParams S{};
        S.d.qoi = d_in, S.d.single[0] = 0, S.d.single[1] = in_size;
        S.d.out = d_out, S.d.out_stride = room * channels;  // mode 1: capacity in bytes
        S.d.target = channels, S.d.flip = 0, S.d.n_images = 1, S.d.n_pixels = room;
        S.d.results = d_res;
        S.mode = 1, S.init = reinterpret_cast<const DecState*>(d_state), S.in_size = in_size, S.only_if_bad = 0;
        // The parallel kernel needs a few tiles to be worth its launch; a short buffer (and a call that has nothing to read) is
        // decoded by the sequential loop alone.
        if (in_size >= c->stream_parallel_min && tiles > 0) {
            DecParams& P = S.d;
            QB_CUDA(c->fix.reserve((size_t)tiles * kFixWords * sizeof(uint32_t), s, true));
            QB_CUDA(c->next_epoch(tiles * kDecDescWords * sizeof(uint64_t), s, kDecRounds + 1));
            P.offsets = nullptr, P.tile_first = nullptr;
            P.n_tiles = (uint32_t)tiles, P.epoch = c->epoch, P.round = 0;
            P.control = static_cast<DecControl*>(c->results.p);
            P.desc    = static_cast<uint64_t*>(c->carry.p);
            P.fix     = static_cast<uint32_t*>(c->fix.p);
            P.init    = S.init;
            P.req_cap = (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(64, tiles / 64));
            QB_CUDA(c->reqs.reserve(sizeof(CascadeReq) * P.req_cap, s));
            P.req = static_cast<CascadeReq*>(c->reqs.p);
            const unsigned want   = ((unsigned)tiles + kWtWarps - 1) / kWtWarps;
            const unsigned n_ctas = std::max(1u, std::min<unsigned>(want, (unsigned)c->dec_coresident));
            decode_wt_stream_kernel<<<n_ctas, kWtThreads, kWtSmemBytes, s>>>(P);
            QB_CUDA(cudaGetLastError());
            void* args[] = { &P };  // the retry rounds (returns at once when round 0 verified every tile)
            QB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(decode_finish_stream_kernel), dim3(n_ctas), dim3(kWtThreads), args, kWtSmemBytes, s));
            S.only_if_bad = 1;  // the sequential loop runs only for what the last round still refutes
        }
        decode_serial_kernel<<<1, 32, sizeof(SerialSmem), s>>>(S);
        QB_CUDA(cudaGetLastError());
        stream_decode_epilogue_kernel<<<1, 64, 0, s>>>(d_res, reinterpret_cast<StreamOut*>(d_result), reinterpret_cast<DecState*>(d_state));
        QB_CUDA(cudaGetLastError());
        return 0;
    }

    int32_t qoipp_b200_stream_decode_host(qoipp_b200_ctx* c, qoipp_b200_state* st, const uint8_t* h_in, uint64_t in_size,
                                          uint8_t* h_out, uint64_t out_cap, uint64_t* processed, uint64_t* written)
    {
        // error order of StreamDecoder::decode, source/stream.cpp:314-320
        if (!st->channels) return H::NotInitialized;
        if (out_cap == 0) return H::Empty;
        if (out_cap < st->channels) return H::TooShort;
        *processed = 0, *written = 0;
        Guard          g(c->device);
        const unsigned ch = st->channels;
        // a call produces at most 62 pixels per input byte plus the pending run
        const uint64_t cap = std::min<uint64_t>(out_cap, (in_size * 62 + st->run) * ch);
        if (cap < ch) return 0;  // nothing to read and nothing pending
        cudaStream_t s  = c->own_stream;
        QB_CUDA(c->stage_in.reserve(in_size + 64, s));
        QB_CUDA(c->stage_out.reserve(cap + 64, s));
        QB_CUDA(c->state.reserve(sizeof(DecState) + sizeof(StreamOut), s));
        auto* d_state = static_cast<DecState*>(c->state.p);
        auto* d_sres  = reinterpret_cast<StreamOut*>(d_state + 1);
        // pinned block: [0] carry-in, [1] carry-out + result
        auto* hs = reinterpret_cast<DecState*>(static_cast<uint8_t*>(c->h_result.p) + 1024);
        hs->prev = pack_px(st->prev), hs->run = st->run;
        for (int i = 0; i < 64; ++i) hs->table[i] = pack_px(st->seen[i]);
        QB_CUDA(cudaMemcpyAsync(d_state, hs, sizeof(DecState), cudaMemcpyHostToDevice, s));
        if (in_size) {
            if (const uint8_t* m = mapped_host(h_in); m && in_size < (1u << 20)) QB_CUDA(cudaMemcpyAsync(c->stage_in.p, m, in_size, cudaMemcpyDefault, s));
            else QB_CUDA(pageable_to_device(c, c->stage_in.p, h_in, in_size, s));
        }
        if (int32_t e = qoipp_b200_stream_decode_dev(c, (uint8_t)ch, reinterpret_cast<qoipp_b200_dev_state*>(d_state), static_cast<uint8_t*>(c->stage_in.p), in_size,
                                                     static_cast<uint8_t*>(c->stage_out.p), cap, reinterpret_cast<qoipp_b200_stream_result*>(d_sres), s))
            return e;
        auto* hr = reinterpret_cast<DecState*>(static_cast<uint8_t*>(c->h_result.p) + 2048);
        QB_CUDA(cudaMemcpyAsync(hr, d_state, sizeof(DecState) + sizeof(StreamOut), cudaMemcpyDeviceToHost, s));
        QB_CUDA(cudaStreamSynchronize(s));
        const auto* ho = reinterpret_cast<const StreamOut*>(hr + 1);
        if (ho->written) QB_CUDA(device_to_pageable(c, h_out, c->stage_out.p, ho->written, s));
        *processed = ho->processed, *written = ho->written;
        unpack_px(hr->prev, st->prev);
        st->run = (uint8_t)hr->run;
        for (int i = 0; i < 64; ++i) unpack_px(hr->table[i], st->seen[i]);
        return 0;
    }
}
