// qoipp_api.cpp -- the C++20 `qoipp::` API (include/qoipp/qoipp.hpp) as thin host glue over the C ABI of
// libqoipp_b200 (include/qoipp_b200.h).  Validation order, error values and file semantics follow the reference
// (mrizaln/qoipp v0.5.0, source/simple.cpp:174-569, source/stream.cpp:103-459, source/common.cpp:9-73); every
// pixel and chunk byte is produced by the CUDA kernels -- nothing here encodes or decodes on the CPU.
#include "qoipp/qoipp.hpp"

#include "qoipp_b200.h"

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <new>
#include <sstream>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#include <unistd.h>
#endif

namespace fs = std::filesystem;

namespace
{
    using namespace qoipp;

    // one device context per host thread AND device: concurrent callers never share a stream or a workspace.  The device
    // is the one the calling thread selected (qoipp::b200::set_device, else QOIPP_B200_DEVICE, else the thread's current
    // CUDA device), so a thread-per-GPU consumer of qoipp::encode / decode reaches every GPU of the box.
    struct ThreadCtx {
        qoipp_b200_ctx* ctx  = nullptr;
        int32_t         code = 0;
        int             device = -1;
    };

    thread_local int tl_device = -1;  // qoipp::b200::set_device

    int env_device()
    {
        static const int v = [] {
            const char* e = std::getenv("QOIPP_B200_DEVICE");
            return e && *e ? std::atoi(e) : -1;
        }();
        return v;
    }

    int wanted_device()
    {
        if (tl_device >= 0) return tl_device;
        if (env_device() >= 0) return env_device();
        return qoipp_b200_current_device();
    }

    struct ThreadCtxSet {
        std::vector<ThreadCtx> all;
        ~ThreadCtxSet()
        {
            for (auto& t : all)
                if (t.ctx) qoipp_b200_ctx_destroy(t.ctx);
        }
    };

    ThreadCtx& thread_ctx()
    {
        thread_local ThreadCtxSet set;
        const int                 dev = wanted_device();
        for (auto& t : set.all)
            if (t.device == dev) return t;
        ThreadCtx t;
        t.device = dev;
        t.code   = qoipp_b200_ctx_create(dev, &t.ctx);
        set.all.push_back(t);
        return set.all.back();
    }

    // A zero-filled buffer for a whole image or stream (what the reference allocates too, simple.cpp:190, 390).  For large
    // sizes the pages are requested as transparent huge pages before they are touched: the zero fill and the copy that follows
    // then take 2 MiB faults instead of 4 KiB ones (a 133 MB image: 32 k page faults were most of the allocation time).
    ByteVec big_buffer(std::size_t n)
    {
        ByteVec v;
#if defined(__linux__) && defined(MADV_HUGEPAGE)
        if (n >= (8u << 20)) {
            v.reserve(n);
            const auto page = static_cast<std::uintptr_t>(2u << 20);
            const auto lo   = (reinterpret_cast<std::uintptr_t>(v.data()) + page - 1) & ~(page - 1);
            const auto hi   = (reinterpret_cast<std::uintptr_t>(v.data()) + n) & ~(page - 1);
            if (hi > lo) (void)::madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);  // advice only: failure changes nothing
        }
#endif
        v.resize(n);
        return v;
    }

    // C ABI code -> qoipp::Error (the enum has no device member: CUDA failures surface as IoError)
    Error to_error(int32_t code) { return code >= 1 and code <= 14 ? static_cast<Error>(code) : Error::IoError; }

    qoipp_b200_desc to_c(const Desc& d)
    {
        return { d.width, d.height, static_cast<uint8_t>(d.channels), static_cast<uint8_t>(d.colorspace) };
    }

    Desc from_c(const qoipp_b200_desc& d)
    {
        return { d.width, d.height, static_cast<Channels>(d.channels), static_cast<Colorspace>(d.colorspace) };
    }

    uint8_t target_byte(std::optional<Channels> t) { return t ? static_cast<uint8_t>(*t) : uint8_t{ 0 }; }

    // PixelGenFun sources are host callbacks invoked once per index in order (ref: util.hpp:336-343); they are
    // materialised into a host image that the kernels then encode.
    Result<ByteVec> materialise(const PixelGenFun& gen, const Desc& desc) noexcept
    {
        const auto bytes = count_bytes(desc);
        if (not bytes) return make_error<ByteVec>(bytes.error());
        try {
            auto       raw = ByteVec(*bytes);
            const auto ch  = static_cast<std::size_t>(desc.channels);
            const auto n   = *bytes / ch;
            for (std::size_t i = 0; i < n; ++i) {
                const Pixel p = gen(i);
                std::memcpy(raw.data() + i * ch, &p, ch);  // alpha is dropped for RGB, i.e. forced to 255 downstream
            }
            return raw;
        } catch (const std::bad_alloc&) {
            return make_error<ByteVec>(Error::BadAlloc);
        }
    }

    Result<ByteVec> slurp(const fs::path& path) noexcept  // ref: simple.cpp:422-441
    {
        if (not fs::exists(path)) return make_error<ByteVec>(Error::FileNotExists);
        if (not fs::is_regular_file(path)) return make_error<ByteVec>(Error::NotRegularFile);
        auto file = std::ifstream{ path, std::ios::binary };
        if (not file.is_open()) return make_error<ByteVec>(Error::IoError);
        try {
            auto buf = std::stringstream{};
            buf << file.rdbuf();
            if (not file) {
                // an empty file sets failbit on `<<` in libstdc++; the reference reports Empty for it through decode()
                if (fs::file_size(path) == 0) return ByteVec{};
                return make_error<ByteVec>(Error::IoError);
            }
            const auto view = buf.view();
            return ByteVec(view.begin(), view.end());
        } catch (...) {
            return make_error<ByteVec>(Error::IoError);
        }
    }

    Result<std::size_t> spill(const fs::path& path, const ByteVec& bytes) noexcept  // ref: simple.cpp:316-328
    {
        auto file = std::ofstream{ path, std::ios::binary | std::ios::trunc };
        if (not file.is_open()) return make_error<std::size_t>(Error::IoError);
        file.write(reinterpret_cast<const char*>(bytes.data()), static_cast<std::streamsize>(bytes.size()));
        if (not file) return make_error<std::size_t>(Error::IoError);
        return bytes.size();
    }

    Result<EncodeStatus> encode_span(ByteSpan out, ByteCSpan in, const Desc& desc)
    {
        auto& t = thread_ctx();
        if (not t.ctx) {
            // argument errors still win over the missing device, as in the reference's check order
            if (in.size() == 0) return make_error<EncodeStatus>(Error::Empty);
            if (auto n = count_bytes(desc); not n) return make_error<EncodeStatus>(n.error());
            return make_error<EncodeStatus>(to_error(t.code));
        }
        const auto cd      = to_c(desc);
        uint64_t   written = 0;
        int32_t    complete = 0;
        const auto code = qoipp_b200_encode_host(t.ctx, in.data(), in.size(), &cd, out.data(), out.size(), &written, &complete);
        if (code) return make_error<EncodeStatus>(to_error(code));
        return EncodeStatus{ static_cast<std::size_t>(written), complete != 0 };
    }
}

namespace qoipp
{
    // ------------------------------------------------------------------ header
    Result<Desc> read_header(ByteCSpan in_data) noexcept
    {
        qoipp_b200_desc d{};
        if (const auto code = qoipp_b200_read_header(in_data.data(), in_data.size(), &d)) return make_error<Desc>(to_error(code));
        return from_c(d);
    }

    Result<Desc> read_header(const fs::path& in_path) noexcept  // ref: common.cpp:52-72 (short file -> IoError, hazard 5)
    {
        if (not fs::exists(in_path)) return make_error<Desc>(Error::FileNotExists);
        if (not fs::is_regular_file(in_path)) return make_error<Desc>(Error::NotRegularFile);
        auto file = std::ifstream{ in_path, std::ios::binary };
        if (not file.is_open()) return make_error<Desc>(Error::IoError);
        auto head = ByteArr<constants::header_size>{};
        file.read(reinterpret_cast<char*>(head.data()), head.size());
        if (not file) return make_error<Desc>(Error::IoError);
        return read_header(ByteCSpan{ head });
    }

    // ------------------------------------------------------------------ encode
    Result<ByteVec> encode(ByteCSpan in_data, Desc desc) noexcept
    {
        // ref: simple.cpp:182-188 -- Empty, then the descriptor, then the size match
        if (in_data.size() == 0) return make_error<ByteVec>(Error::Empty);
        const auto bytes = count_bytes(desc);
        if (not bytes) return make_error<ByteVec>(bytes.error());
        if (in_data.size() != *bytes) return make_error<ByteVec>(Error::MismatchedDesc);
        try {
            // the stream stays on the device until its size is known: the result is allocated at that size, not at
            // worst_size (the reference allocates and zero-fills the worst case, then shrinks: simple.cpp:190-204)
            auto& t = thread_ctx();
            if (not t.ctx) return make_error<ByteVec>(to_error(t.code));
            const auto cd      = to_c(desc);
            uint64_t   written = 0;
            if (const auto code = qoipp_b200_encode_staged(t.ctx, in_data.data(), in_data.size(), &cd, &written)) return make_error<ByteVec>(to_error(code));
            auto out = big_buffer(written);
            if (const auto code = qoipp_b200_fetch_staged(t.ctx, out.data(), written)) return make_error<ByteVec>(to_error(code));
            return out;
        } catch (const std::bad_alloc&) {
            return make_error<ByteVec>(Error::BadAlloc);
        }
    }

    Result<ByteVec> encode(PixelGenFun in_func, Desc desc) noexcept
    {
        auto raw = materialise(in_func, desc);
        if (not raw) return make_error<ByteVec>(raw.error());
        return encode(ByteCSpan{ *raw }, desc);
    }

    Result<EncodeStatus> encode_into(ByteSpan out_buf, ByteCSpan in_data, Desc desc)
    {
        if (in_data.size() == 0) return make_error<EncodeStatus>(Error::Empty);
        const auto bytes = count_bytes(desc);
        if (not bytes) return make_error<EncodeStatus>(bytes.error());
        if (in_data.size() != *bytes) return make_error<EncodeStatus>(Error::MismatchedDesc);
        return encode_span(out_buf, in_data, desc);
    }

    Result<EncodeStatus> encode_into(ByteSpan out_buf, PixelGenFun in_func, Desc desc)
    {
        auto raw = materialise(in_func, desc);
        if (not raw) return make_error<EncodeStatus>(raw.error());
        return encode_span(out_buf, *raw, desc);
    }

    Result<std::size_t> encode_into(ByteSinkFun out_func, ByteCSpan in_data, Desc desc)
    {
        auto encoded = encode(in_data, desc);
        if (not encoded) return make_error<std::size_t>(encoded.error());
        for (const Byte b : *encoded) out_func(b);  // one call per byte, in order (ref: util.hpp:263-269)
        return encoded->size();
    }

    Result<std::size_t> encode_into(ByteSinkFun out_func, PixelGenFun in_func, Desc desc)
    {
        auto encoded = encode(std::move(in_func), desc);
        if (not encoded) return make_error<std::size_t>(encoded.error());
        for (const Byte b : *encoded) out_func(b);
        return encoded->size();
    }

    namespace
    {
        std::optional<Error> path_writable(const fs::path& p, const Desc& desc, bool overwrite)  // ref: simple.cpp:304-310
        {
            std::error_code ec;
            if (fs::exists(p, ec) and not overwrite) return Error::FileExists;
            if (fs::exists(p, ec) and not fs::is_regular_file(p, ec)) return Error::NotRegularFile;
            if (const auto n = count_bytes(desc); not n) return n.error();
            return std::nullopt;
        }
    }

    Result<std::size_t> encode_into(const fs::path& out_path, ByteCSpan in_data, Desc desc, bool overwrite) noexcept
    {
        if (const auto err = path_writable(out_path, desc, overwrite)) return make_error<std::size_t>(*err);
        auto encoded = encode(in_data, desc);
        if (not encoded) return make_error<std::size_t>(encoded.error());  // nothing is created on failure
        return spill(out_path, *encoded);
    }

    Result<std::size_t> encode_into(const fs::path& out_path, PixelGenFun in_func, Desc desc, bool overwrite) noexcept
    {
        if (const auto err = path_writable(out_path, desc, overwrite)) return make_error<std::size_t>(*err);
        auto encoded = encode(std::move(in_func), desc);
        if (not encoded) return make_error<std::size_t>(encoded.error());
        return spill(out_path, *encoded);
    }

    // ------------------------------------------------------------------ decode
    Result<Desc> decode_into(ByteSpan out_buf, ByteCSpan in_data, std::optional<Channels> target, bool flip_vertically)
    {
        auto& t = thread_ctx();
        if (not t.ctx) {
            if (in_data.size() == 0) return make_error<Desc>(Error::Empty);
            if (in_data.size() <= constants::header_size + constants::end_marker_size) return make_error<Desc>(Error::TooShort);
            if (auto h = read_header(in_data); not h) return make_error<Desc>(h.error());
            return make_error<Desc>(to_error(t.code));
        }
        qoipp_b200_desc d{};
        const auto      code = qoipp_b200_decode_host(t.ctx, in_data.data(), in_data.size(), target_byte(target), flip_vertically,
                                                      out_buf.data(), out_buf.size(), &d);
        if (code) return make_error<Desc>(to_error(code));
        return from_c(d);
    }

    Result<Image> decode(ByteCSpan in_data, std::optional<Channels> target, bool flip_vertically) noexcept
    {
        // ref: simple.cpp:367-395 -- the output is sized with the TARGET channel count
        if (in_data.size() == 0) return make_error<Image>(Error::Empty);
        if (in_data.size() <= constants::header_size + constants::end_marker_size) return make_error<Image>(Error::TooShort);
        auto header = read_header(in_data);
        if (not header) return make_error<Image>(header.error());
        const auto src   = header->channels;
        header->channels = target.value_or(src);
        const auto bytes = count_bytes(*header);
        if (not bytes) return make_error<Image>(bytes.error());
        try {
            // transfer and kernels are enqueued first; the image is allocated while they run, then fetched
            auto& t = thread_ctx();
            if (not t.ctx) return make_error<Image>(to_error(t.code));
            qoipp_b200_desc cd{};
            uint64_t        need = 0;
            if (const auto code = qoipp_b200_decode_staged(t.ctx, in_data.data(), in_data.size(), target_byte(target), flip_vertically, &cd, &need))
                return make_error<Image>(to_error(code));
            auto buf = big_buffer(*bytes);
            if (need != *bytes) return make_error<Image>(Error::IoError);
            if (const auto code = qoipp_b200_fetch_staged(t.ctx, buf.data(), need)) return make_error<Image>(to_error(code));
            return Image{ std::move(buf), from_c(cd) };
        } catch (const std::bad_alloc&) {
            return make_error<Image>(Error::BadAlloc);
        }
    }

    Result<Image> decode(const fs::path& in_path, std::optional<Channels> target, bool flip_vertically) noexcept
    {
        auto bytes = slurp(in_path);
        if (not bytes) return make_error<Image>(bytes.error());
        return decode(ByteCSpan{ *bytes }, target, flip_vertically);
    }

    Result<Desc> decode_into(PixelSinkFun out_func, ByteCSpan in_data)
    {
        // the sink receives whole pixels including the alpha that flows through the stream (ref: util.hpp:303-311)
        auto image = decode(in_data, Channels::RGBA, false);
        if (not image) return make_error<Desc>(image.error());
        const auto n = image->data.size() / 4;
        for (std::size_t i = 0; i < n; ++i) {
            Pixel p;
            std::memcpy(&p, image->data.data() + i * 4, 4);
            out_func(p);
        }
        auto header = read_header(in_data);  // the returned Desc keeps the file's channel count (ref: simple.cpp:514)
        return *header;
    }

    Result<Desc> decode_into(ByteSpan out_buf, const fs::path& in_path, std::optional<Channels> target, bool flip_vertically) noexcept
    {
        auto bytes = slurp(in_path);
        if (not bytes) return make_error<Desc>(bytes.error());
        try {
            return decode_into(out_buf, ByteCSpan{ *bytes }, target, flip_vertically);
        } catch (...) {
            return make_error<Desc>(Error::BadAlloc);
        }
    }

    Result<Desc> decode_into(PixelSinkFun out_func, const fs::path& in_path) noexcept
    {
        auto bytes = slurp(in_path);
        if (not bytes) return make_error<Desc>(bytes.error());
        try {
            return decode_into(std::move(out_func), ByteCSpan{ *bytes });
        } catch (...) {
            return make_error<Desc>(Error::BadAlloc);
        }
    }

    // ------------------------------------------------------------------ resumable encoder (ref: stream.cpp:105-277)
    namespace
    {
        constexpr Pixel start_pixel{ 0x00, 0x00, 0x00, 0xFF };  // ref: util.hpp:42

        template <typename Self>
        qoipp_b200_state pack_state(uint8_t channels, uint8_t target, Byte run, Pixel prev, const Self& seen)
        {
            qoipp_b200_state s{};
            s.channels = channels, s.target = target, s.run = run;
            std::memcpy(s.prev, &prev, 4);
            std::memcpy(s.seen, seen.data(), sizeof(s.seen));
            return s;
        }
    }

    StreamEncoder::StreamEncoder() noexcept
        : m_channels{}
        , m_run{ 0 }
        , m_prev{ start_pixel }
        , m_seen{}
    {
    }

    Result<std::size_t> StreamEncoder::initialize(ByteSpan out_buf, Desc desc) noexcept
    {
        if (m_channels) return make_error<std::size_t>(Error::AlreadyInitialized);
        if (out_buf.size() == 0) return make_error<std::size_t>(Error::Empty);
        if (out_buf.size() < constants::header_size) return make_error<std::size_t>(Error::TooShort);
        if (const auto n = count_bytes(desc); not n) return make_error<std::size_t>(n.error());
        Byte* o = out_buf.data();
        std::memcpy(o, constants::magic.data(), 4);
        for (int i = 0; i < 4; ++i) {
            o[4 + i] = static_cast<Byte>(desc.width >> (24 - 8 * i));
            o[8 + i] = static_cast<Byte>(desc.height >> (24 - 8 * i));
        }
        o[12]      = static_cast<Byte>(desc.channels);
        o[13]      = static_cast<Byte>(desc.colorspace);
        m_channels = desc.channels;
        return constants::header_size;
    }

    Result<StreamResult> StreamEncoder::encode(ByteSpan out_buf, ByteCSpan in_buf) noexcept
    {
        if (not m_channels) return make_error<StreamResult>(Error::NotInitialized);
        if (out_buf.empty() or in_buf.empty()) return make_error<StreamResult>(Error::Empty);
        if (out_buf.size() < 5) return make_error<StreamResult>(Error::TooShort);
        auto& t = thread_ctx();
        if (not t.ctx) return make_error<StreamResult>(to_error(t.code));
        auto     st        = pack_state(static_cast<uint8_t>(*m_channels), 0, m_run, m_prev, m_seen);
        uint64_t processed = 0, written = 0;
        const auto code = qoipp_b200_stream_encode_host(t.ctx, &st, in_buf.data(), in_buf.size(), out_buf.data(), out_buf.size(),
                                                        &processed, &written);
        if (code) return make_error<StreamResult>(to_error(code));
        m_run = st.run;
        std::memcpy(&m_prev, st.prev, 4);
        std::memcpy(m_seen.data(), st.seen, sizeof(st.seen));
        return StreamResult{ static_cast<std::size_t>(processed), static_cast<std::size_t>(written) };
    }

    Result<std::size_t> StreamEncoder::finalize(ByteSpan out_buf) noexcept
    {
        if (not m_channels) return make_error<std::size_t>(Error::NotInitialized);
        if (out_buf.size() == 0) return make_error<std::size_t>(Error::Empty);
        const std::size_t need = constants::end_marker_size + (m_run > 0 ? 1 : 0);
        if (out_buf.size() < need) return make_error<std::size_t>(Error::TooShort);
        Byte* o = out_buf.data();
        if (m_run > 0) *o++ = static_cast<Byte>(0xC0 | (m_run - 1));  // pending QOI_OP_RUN (ref: util.hpp:227-235)
        std::memset(o, 0, 7);
        o[7] = 1;
        reset();
        return need;
    }

    void StreamEncoder::reset() noexcept
    {
        m_channels.reset();
        m_run  = 0;
        m_prev = start_pixel;
        m_seen.fill(Pixel{});
    }

    // ------------------------------------------------------------------ resumable decoder (ref: stream.cpp:282-458)
    StreamDecoder::StreamDecoder() noexcept
        : m_channels{}
        , m_target{}
        , m_run{ 0 }
        , m_prev{ start_pixel }
        , m_seen{}
    {
    }

    Result<Desc> StreamDecoder::initialize(ByteCSpan in_buf, std::optional<Channels> target) noexcept
    {
        if (m_channels) return make_error<Desc>(Error::AlreadyInitialized);
        auto desc = read_header(in_buf);
        if (not desc) return desc;
        if (const auto n = count_bytes(*desc); not n) return make_error<Desc>(n.error());
        m_target       = target.value_or(desc->channels);
        m_channels     = m_target;  // ref: stream.cpp:302-304 -- both become the target
        desc->channels = *m_channels;
        m_seen[(m_prev.r * 3 + m_prev.g * 5 + m_prev.b * 7 + m_prev.a * 11) % constants::running_array_size] = m_prev;
        return desc;
    }

    Result<StreamResult> StreamDecoder::decode(ByteSpan out_buf, ByteCSpan in_buf) noexcept
    {
        if (not m_channels) return make_error<StreamResult>(Error::NotInitialized);
        if (out_buf.size() == 0) return make_error<StreamResult>(Error::Empty);
        if (out_buf.size() < static_cast<std::size_t>(*m_channels)) return make_error<StreamResult>(Error::TooShort);
        auto& t = thread_ctx();
        if (not t.ctx) return make_error<StreamResult>(to_error(t.code));
        auto st = pack_state(static_cast<uint8_t>(*m_channels), static_cast<uint8_t>(*m_target), m_run, m_prev, m_seen);
        uint64_t processed = 0, written = 0;
        const auto code = qoipp_b200_stream_decode_host(t.ctx, &st, in_buf.data(), in_buf.size(), out_buf.data(), out_buf.size(),
                                                        &processed, &written);
        if (code) return make_error<StreamResult>(to_error(code));
        m_run = st.run;
        std::memcpy(&m_prev, st.prev, 4);
        std::memcpy(m_seen.data(), st.seen, sizeof(st.seen));
        return StreamResult{ static_cast<std::size_t>(processed), static_cast<std::size_t>(written) };
    }

    Result<std::size_t> StreamDecoder::drain_run(ByteSpan out_buf) noexcept
    {
        if (not m_channels) return make_error<std::size_t>(Error::NotInitialized);
        if (out_buf.size() == 0) return make_error<std::size_t>(Error::Empty);
        const auto  ch = static_cast<std::size_t>(*m_channels);
        std::size_t n  = 0;
        while (m_run > 0 and (n + 1) * ch <= out_buf.size()) {  // replicate the pending pixel: a byte copy, not a decode
            std::memcpy(out_buf.data() + n * ch, &m_prev, ch);
            ++n, --m_run;
        }
        return n * ch;
    }

    void StreamDecoder::reset() noexcept
    {
        m_channels.reset();
        m_target.reset();
        m_run  = 0;
        m_prev = start_pixel;
        m_seen.fill(Pixel{});
    }
}

// ---- extension (not in the reference): which GPU serves the calling thread
namespace qoipp::b200
{
    void set_device(int device) noexcept { tl_device = device; }
    int  device() noexcept { return wanted_device(); }
    int  device_count() noexcept { return qoipp_b200_device_count(); }
}
