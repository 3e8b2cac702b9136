// host_util.hpp -- host-only helpers shared by the C ABI (qoipp_b200.cu) and the C++ API glue.
// Descriptor rules restate include/qoipp/common.hpp:346-412 and source/common.cpp:13-50 of the reference.
#pragma once

#include "../../include/qoipp_b200.h"

#include <cstdint>
#include <cstring>

namespace qb::host
{
    enum Error : int32_t {  // qoipp::Error numbering, common.hpp:78-94
        Ok = 0, Empty = 1, TooShort, TooBig, NotQoi, InvalidDesc, MismatchedDesc, NotEnoughSpace, NotInitialized,
        AlreadyInitialized, NotRegularFile, FileExists, FileNotExists, IoError, BadAlloc
    };

    constexpr uint64_t kHeaderSize = 14, kMarkerSize = 8;

    inline bool is_valid(const qoipp_b200_desc& d)
    {
        return d.width > 0 && d.height > 0 && (d.channels == 3 || d.channels == 4) && d.colorspace <= 1;
    }

    inline int32_t count_bytes(const qoipp_b200_desc& d, uint64_t* out)
    {
        if (!is_valid(d)) return InvalidDesc;
        const uint64_t px = (uint64_t)d.width * d.height;  // < 2^64 always
        const uint64_t by = px * d.channels;
        if (by / d.channels != px) return TooBig;
        *out = by;
        return Ok;
    }

    inline int32_t worst_size(const qoipp_b200_desc& d, uint64_t* out)
    {
        uint64_t n;
        if (int32_t e = count_bytes(d, &n)) return e;
        // (channels + 1) * width * height + 22 (common.hpp:394-412), refused when it does not fit 64 bits: the reference would
        // hand back a wrapped value for such a descriptor; a capacity test against a wrapped value must never pass
        const uint64_t px = (uint64_t)d.width * d.height, mul = (uint64_t)d.channels + 1;
        if (px > (~0ull - (kHeaderSize + kMarkerSize)) / mul) return TooBig;
        *out = mul * px + kHeaderSize + kMarkerSize;
        return Ok;
    }

    inline int32_t read_header(const uint8_t* in, uint64_t size, qoipp_b200_desc* out)
    {
        if (size == 0) return Empty;
        if (size < kHeaderSize) return TooShort;
        if (std::memcmp(in, "qoif", 4) != 0) return NotQoi;
        auto be32 = [&](int o) { return (uint32_t)in[o] << 24 | (uint32_t)in[o + 1] << 16 | (uint32_t)in[o + 2] << 8 | in[o + 3]; };
        const uint32_t w = be32(4), h = be32(8);
        if ((in[12] != 3 && in[12] != 4) || in[13] > 1 || w == 0 || h == 0) return InvalidDesc;
        *out = qoipp_b200_desc{ w, h, in[12], in[13] };
        return Ok;
    }

    inline void write_header(const qoipp_b200_desc& d, uint8_t* out14)
    {
        std::memcpy(out14, "qoif", 4);
        for (int i = 0; i < 4; ++i) {
            out14[4 + i] = (uint8_t)(d.width >> (24 - 8 * i));
            out14[8 + i] = (uint8_t)(d.height >> (24 - 8 * i));
        }
        out14[12] = d.channels;
        out14[13] = d.colorspace;
    }

    inline const char* error_string(int32_t code)
    {
        switch (code) {
        case Ok: return "Ok";
        case Empty: return "Data is empty";
        case TooShort: return "Data is too short";
        case TooBig: return "Image is too big to process";
        case NotQoi: return "Not a QOI file";
        case InvalidDesc: return "Image description is invalid";
        case MismatchedDesc: return "Image description does not match the data";
        case NotEnoughSpace: return "Buffer does not have enough space";
        case NotInitialized: return "Stream encoder/decoder is not initialized yet";
        case AlreadyInitialized: return "Stream encoder/decoder already initialized";
        case NotRegularFile: return "Not a regular file";
        case FileExists: return "File already exists";
        case FileNotExists: return "File does not exist";
        case IoError: return "Unable to do read or write operation";
        case BadAlloc: return "Failed to allocate memory";
        default: return code < 0 ? "CUDA error" : "Unknown";
        }
    }
}  // namespace qb::host
