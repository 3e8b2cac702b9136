// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" doorway into the UNMODIFIED reference (mrizaln/qoipp), compiled by oracle/Makefile from the
// sources where they lie under /root/reference into oracle/_ref/libqoipp_ref.so.  No reference source is
// copied into this repository; this file only calls the reference's public API (include/qoipp/*.hpp).
// Used to pin oracle/qoi_oracle.c, as the strongest parity checker on the GPU box (the .so travels,
// /root/reference does not) and as the CPU baseline of bench.py (cpu_baseline.kind = "reference").
#include <qoipp/simple.hpp>
#include <qoipp/stream.hpp>

#include <atomic>
#include <chrono>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

namespace
{
    qoipp::Desc make_desc(uint32_t w, uint32_t h, uint8_t ch, uint8_t cs)
    {
        return { w, h, static_cast<qoipp::Channels>(ch), static_cast<qoipp::Colorspace>(cs) };
    }

    std::optional<qoipp::Channels> make_target(uint8_t target)
    {
        return target ? std::optional{ static_cast<qoipp::Channels>(target) } : std::nullopt;
    }

    void put_desc(const qoipp::Desc& d, uint32_t* out4)
    {
        out4[0] = d.width;
        out4[1] = d.height;
        out4[2] = static_cast<uint32_t>(d.channels);
        out4[3] = static_cast<uint32_t>(d.colorspace);
    }
}

extern "C"
{
    // every entry returns 0 on success or the qoipp::Error value

    int ref_worst_size(uint32_t w, uint32_t h, uint8_t ch, uint8_t cs, uint64_t* out)
    {
        auto r = qoipp::worst_size(make_desc(w, h, ch, cs));
        if (not r) return static_cast<int>(r.error());
        *out = *r;
        return 0;
    }

    int ref_read_header(const uint8_t* in, uint64_t size, uint32_t* desc4)
    {
        auto r = qoipp::read_header(qoipp::ByteCSpan{ in, size });
        if (not r) return static_cast<int>(r.error());
        put_desc(*r, desc4);
        return 0;
    }

    // qoipp::encode_into(ByteSpan, ByteCSpan, Desc)
    int ref_encode_into(uint8_t* out, uint64_t cap, const uint8_t* raw, uint64_t raw_size, uint32_t w, uint32_t h,
                        uint8_t ch, uint8_t cs, uint64_t* written, int* complete)
    {
        auto r = qoipp::encode_into(qoipp::ByteSpan{ out, cap }, qoipp::ByteCSpan{ raw, raw_size }, make_desc(w, h, ch, cs));
        if (not r) return static_cast<int>(r.error());
        *written  = r->written;
        *complete = r->complete;
        return 0;
    }

    // qoipp::decode_into(ByteSpan, ByteCSpan, target, flip)
    int ref_decode_into(uint8_t* out, uint64_t cap, const uint8_t* in, uint64_t size, uint8_t target, int flip,
                        uint32_t* desc4)
    {
        auto r = qoipp::decode_into(qoipp::ByteSpan{ out, cap }, qoipp::ByteCSpan{ in, size }, make_target(target), flip != 0);
        if (not r) return static_cast<int>(r.error());
        put_desc(*r, desc4);
        return 0;
    }

    // qoipp::decode(ByteCSpan, target, flip): allocating form; copies into `out` (cap must hold w*h*target)
    int ref_decode(uint8_t* out, uint64_t cap, const uint8_t* in, uint64_t size, uint8_t target, int flip, uint32_t* desc4)
    {
        auto r = qoipp::decode(qoipp::ByteCSpan{ in, size }, make_target(target), flip != 0);
        if (not r) return static_cast<int>(r.error());
        put_desc(r->desc, desc4);
        if (r->data.size() > cap) return -1;
        std::memcpy(out, r->data.data(), r->data.size());
        return 0;
    }

    // ---- StreamEncoder
    void* ref_senc_new() { return new (std::nothrow) qoipp::StreamEncoder{}; }
    void  ref_senc_delete(void* p) { delete static_cast<qoipp::StreamEncoder*>(p); }
    int   ref_senc_initialize(void* p, uint8_t* out, uint64_t cap, uint32_t w, uint32_t h, uint8_t ch, uint8_t cs,
                              uint64_t* written)
    {
        auto r = static_cast<qoipp::StreamEncoder*>(p)->initialize({ out, cap }, make_desc(w, h, ch, cs));
        if (not r) return static_cast<int>(r.error());
        *written = *r;
        return 0;
    }
    int ref_senc_encode(void* p, uint8_t* out, uint64_t cap, const uint8_t* in, uint64_t in_size, uint64_t* processed,
                        uint64_t* written)
    {
        auto r = static_cast<qoipp::StreamEncoder*>(p)->encode({ out, cap }, { in, in_size });
        if (not r) return static_cast<int>(r.error());
        *processed = r->processed;
        *written   = r->written;
        return 0;
    }
    int ref_senc_finalize(void* p, uint8_t* out, uint64_t cap, uint64_t* written)
    {
        auto r = static_cast<qoipp::StreamEncoder*>(p)->finalize({ out, cap });
        if (not r) return static_cast<int>(r.error());
        *written = *r;
        return 0;
    }
    void ref_senc_reset(void* p) { static_cast<qoipp::StreamEncoder*>(p)->reset(); }
    int  ref_senc_has_run(void* p) { return static_cast<qoipp::StreamEncoder*>(p)->has_run_count(); }

    // ---- StreamDecoder
    void* ref_sdec_new() { return new (std::nothrow) qoipp::StreamDecoder{}; }
    void  ref_sdec_delete(void* p) { delete static_cast<qoipp::StreamDecoder*>(p); }
    int   ref_sdec_initialize(void* p, const uint8_t* in, uint64_t size, uint8_t target, uint32_t* desc4)
    {
        auto r = static_cast<qoipp::StreamDecoder*>(p)->initialize({ in, size }, make_target(target));
        if (not r) return static_cast<int>(r.error());
        put_desc(*r, desc4);
        return 0;
    }
    int ref_sdec_decode(void* p, uint8_t* out, uint64_t cap, const uint8_t* in, uint64_t in_size, uint64_t* processed,
                        uint64_t* written)
    {
        auto r = static_cast<qoipp::StreamDecoder*>(p)->decode({ out, cap }, { in, in_size });
        if (not r) return static_cast<int>(r.error());
        *processed = r->processed;
        *written   = r->written;
        return 0;
    }
    int ref_sdec_drain_run(void* p, uint8_t* out, uint64_t cap, uint64_t* written)
    {
        auto r = static_cast<qoipp::StreamDecoder*>(p)->drain_run({ out, cap });
        if (not r) return static_cast<int>(r.error());
        *written = *r;
        return 0;
    }
    void ref_sdec_reset(void* p) { static_cast<qoipp::StreamDecoder*>(p)->reset(); }
    int  ref_sdec_run_count(void* p) { return static_cast<qoipp::StreamDecoder*>(p)->run_count(); }

    // ---- CPU baseline harness (bench.py `cpu_baseline` / `--impl reference`), the method of the reference's own benchmark
    // (example/source/04_bench.cpp:445-510, 733-754): qoipp::encode_into into a PRE-ALLOCATED,
    // PRE-TOUCHED worst_size buffer, qoipp::decode_into into a pre-allocated buffer, one untimed call + 3 warm-ups, then
    // `reps` timed calls, steady_clock.  thread-per-image: thread i owns images i, i + T, ...; the time of a direction is
    // the wall time from the moment all threads are released until the last one finishes its timed calls.
    //   raws[k] : n_images raw images of raw_size bytes each (w x h x ch)
    //   threads : 0 = std::thread::hardware_concurrency()
    // Returns 0, or the qoipp::Error of a failing call, or -2 when a decode does not reproduce its input.
    int ref_bench(const uint8_t* const* raws, uint32_t n_images, uint64_t raw_size, uint32_t w, uint32_t h, uint8_t ch, uint8_t cs,
                  int threads, int warmups, int reps, double* enc_seconds, double* dec_seconds, uint64_t* enc_bytes_total,
                  int* threads_used)
    {
        using clock = std::chrono::steady_clock;
        const qoipp::Desc desc = make_desc(w, h, ch, cs);
        auto              ws   = qoipp::worst_size(desc);
        if (not ws) return static_cast<int>(ws.error());
        const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
        const unsigned T  = std::min<unsigned>(threads > 0 ? (unsigned)threads : hc, n_images);
        *threads_used     = (int)T;
        std::vector<std::vector<uint8_t>> enc(n_images), dec(n_images);
        std::vector<uint64_t>             enc_len(n_images, 0);
        std::atomic<int>                  err{ 0 };
        std::atomic<unsigned>             ready{ 0 }, go{ 0 }, done{ 0 };
        std::vector<clock::time_point>    t_end(T);
        clock::time_point                 t_go;
        auto run_phase = [&](bool decode_phase) -> double {
            ready = 0, go = 0, done = 0;
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    auto one = [&](uint32_t k) {
                        if (!decode_phase) {
                            auto r = qoipp::encode_into(qoipp::ByteSpan{ enc[k].data(), enc[k].size() }, qoipp::ByteCSpan{ raws[k], raw_size }, desc);
                            if (not r) err = static_cast<int>(r.error());
                            else enc_len[k] = r->written;
                        } else {
                            auto r = qoipp::decode_into(qoipp::ByteSpan{ dec[k].data(), dec[k].size() }, qoipp::ByteCSpan{ enc[k].data(), enc_len[k] },
                                                        std::nullopt, false);
                            if (not r) err = static_cast<int>(r.error());
                        }
                    };
                    for (uint32_t k = t; k < n_images; k += T) {  // buffers allocated and touched by the thread that uses them
                        if (!decode_phase) enc[k].assign(*ws, 0xAA);
                        else dec[k].assign(raw_size, 0xAA);
                    }
                    for (int i = 0; i < 1 + warmups; ++i)
                        for (uint32_t k = t; k < n_images; k += T) one(k);
                    ready.fetch_add(1);
                    while (go.load(std::memory_order_acquire) == 0) std::this_thread::yield();
                    for (int i = 0; i < reps; ++i)
                        for (uint32_t k = t; k < n_images; k += T) one(k);
                    t_end[t] = clock::now();
                    done.fetch_add(1);
                });
            while (ready.load() < T) std::this_thread::yield();
            t_go = clock::now();
            go.store(1, std::memory_order_release);
            for (auto& x : th) x.join();
            clock::time_point last = t_go;
            for (auto& e : t_end) last = std::max(last, e);
            return std::chrono::duration<double>(last - t_go).count();
        };
        *enc_seconds = run_phase(false);
        if (err) return err;
        *dec_seconds = run_phase(true);
        if (err) return err;
        uint64_t total = 0;
        for (uint32_t k = 0; k < n_images; ++k) {
            total += enc_len[k];
            if (std::memcmp(dec[k].data(), raws[k], raw_size) != 0) return -2;
        }
        *enc_bytes_total = total;
        return 0;
    }
}
