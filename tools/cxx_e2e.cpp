// cxx_e2e.cpp -- bench.py's `e2e_pageable` leg: the calls a user of the reference makes, qoipp::encode(ByteCSpan, Desc)
// and qoipp::decode(ByteCSpan) on ordinary (pageable) host memory, through libqoipp.so (-> libqoipp_b200.so -> the CUDA
// kernels).  Built by qoipp_b200/csrc/cxx/Makefile into qoipp_b200/libqoipp_e2e.so; a plain C door so that bench.py can
// time it with ctypes.  Allocation of the returned vectors is inside the timed region, as it is for the reference
// (source/simple.cpp:190-204, 390-413).
#include "qoipp/qoipp.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>

extern "C" int qoipp_cxx_roundtrip(const uint8_t* raw, uint64_t raw_size, uint32_t w, uint32_t h, uint8_t ch, int device, int warmups, int reps,
                                   double* enc_seconds, double* dec_seconds, uint64_t* enc_bytes)
{
    using clock = std::chrono::steady_clock;
    qoipp::b200::set_device(device);
    const qoipp::Desc desc{ w, h, static_cast<qoipp::Channels>(ch), qoipp::Colorspace::sRGB };
    const auto        in = qoipp::ByteCSpan{ raw, raw_size };
    double            te = 0, td = 0;
    for (int i = 0; i < warmups + reps; ++i) {
        const auto t0  = clock::now();
        auto       enc = qoipp::encode(in, desc);
        const auto t1  = clock::now();
        if (not enc) return static_cast<int>(enc.error());
        const auto t2  = clock::now();
        auto       dec = qoipp::decode(qoipp::ByteCSpan{ enc->data(), enc->size() });
        const auto t3  = clock::now();
        if (not dec) return 100 + static_cast<int>(dec.error());
        if (i >= warmups) {
            te += std::chrono::duration<double>(t1 - t0).count();
            td += std::chrono::duration<double>(t3 - t2).count();
        }
        *enc_bytes = enc->size();
        if (i + 1 == warmups + reps)
            if (dec->data.size() != raw_size or std::memcmp(dec->data.data(), raw, raw_size) != 0) return -2;
    }
    *enc_seconds = te, *dec_seconds = td;
    return 0;
}
