// qoipp.hpp -- the public C++20 API of qoipp, served by the B200-native backend (libqoipp_b200).
//
// Source-compatible with mrizaln/qoipp v0.5.0: the same names, argument meaning, error values and check order as
// the reference headers include/qoipp/{common,simple,stream}.hpp (cited per declaration as ref:<file>:<line>).
// <qoipp/common.hpp>, <qoipp/simple.hpp> and <qoipp/stream.hpp> forward to this file.  Every codec call runs
// the CUDA kernels behind include/qoipp_b200.h; there is no CPU implementation in this library.
#ifndef QOIPP_B200_QOIPP_HPP
#define QOIPP_B200_QOIPP_HPP

#include <array>
#include <concepts>
#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <functional>
#include <optional>
#include <span>
#include <string_view>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

#if defined(__cpp_lib_expected)
#    include <expected>
#endif

namespace qoipp
{
    // ------------------------------------------------------------------ constants (ref:common.hpp:17-23)
    namespace constants
    {
        inline constexpr std::string_view magic              = "qoif";
        inline constexpr std::size_t      header_size        = 14;
        inline constexpr std::size_t      end_marker_size    = 8;
        inline constexpr std::size_t      running_array_size = 64;
    }

    // ------------------------------------------------------------------ vocabulary types (ref:common.hpp:29-159)
    struct Pixel;

    using Byte      = std::uint8_t;
    using ByteVec   = std::vector<Byte>;
    using ByteSpan  = std::span<Byte>;
    using ByteCSpan = std::span<const Byte>;
    template <std::size_t N>
    using ByteArr = std::array<Byte, N>;

    using PixelVec   = std::vector<Pixel>;
    using PixelSpan  = std::span<Pixel>;
    using PixelCSpan = std::span<const Pixel>;
    template <std::size_t N>
    using PixelArr = std::array<Pixel, N>;

    using PixelGenFun  = std::function<Pixel(std::size_t index)>;  // called once per index, in order
    using PixelSinkFun = std::function<void(Pixel pixel)>;
    using ByteSinkFun  = std::function<void(std::uint8_t byte)>;

    enum class Colorspace : Byte { sRGB = 0, Linear = 1 };  // header byte 13 only, never affects the chunks
    enum class Channels : Byte { RGB = 3, RGBA = 4 };       // bytes per pixel

    enum class Error  // ref:common.hpp:78-94, numbering is part of the C ABI (include/qoipp_b200.h)
    {
        Empty = 1,
        TooShort,
        TooBig,
        NotQoi,
        InvalidDesc,
        MismatchedDesc,
        NotEnoughSpace,
        NotInitialized,
        AlreadyInitialized,
        NotRegularFile,
        FileExists,
        FileNotExists,
        IoError,   // also: any CUDA failure other than out-of-memory
        BadAlloc,  // also: cudaErrorMemoryAllocation
    };

    struct Pixel {
        Byte r, g, b, a;
        constexpr auto operator<=>(const Pixel&) const = default;
    };

    struct Desc {
        std::uint32_t width;
        std::uint32_t height;
        Channels      channels;
        Colorspace    colorspace;
        constexpr auto operator<=>(const Desc&) const = default;
    };

    struct Image {
        ByteVec data;  // width * height * desc.channels bytes
        Desc    desc;
    };

    struct EncodeStatus {
        std::size_t written;   // bytes stored: always a whole number of chunks
        bool        complete;  // false when the buffer ended before the end marker
    };

    struct StreamResult {
        std::size_t processed;  // input bytes consumed
        std::size_t written;    // output bytes produced
    };

    // ------------------------------------------------------------------ Result<T> (ref:common.hpp:161-253)
#if defined(__cpp_lib_expected)
    template <typename T>
    using Result = std::expected<T, Error>;
#else
    // std::expected stand-in for C++20: holds either a T or an Error.
    template <typename T>
    class [[nodiscard]] Result
    {
    public:
        Result() = default;

        template <typename U>
            requires std::constructible_from<T, U> or std::same_as<std::decay_t<U>, Error>
        Result(U&& u)
            : m_state{ std::forward<U>(u) }
        {
        }

        bool     has_value() const noexcept { return m_state.index() == 0; }
        explicit operator bool() const noexcept { return has_value(); }

        T&       value() & { return std::get<0>(m_state); }
        const T& value() const& { return std::get<0>(m_state); }
        T&&      value() && { return std::get<0>(std::move(m_state)); }

        Error&       error() & { return std::get<1>(m_state); }
        const Error& error() const& { return std::get<1>(m_state); }
        Error&&      error() && { return std::get<1>(std::move(m_state)); }

        T&       operator*() & noexcept { return value(); }
        const T& operator*() const& noexcept { return value(); }
        T&&      operator*() && noexcept { return std::move(value()); }
        T*       operator->() noexcept { return &value(); }
        const T* operator->() const noexcept { return &value(); }

    private:
        std::variant<T, Error> m_state;
    };

    template <>
    class Result<void>
    {
    public:
        Result() = default;
        Result(Error e)
            : m_error{ e }
        {
        }
        bool     has_value() const noexcept { return not m_error.has_value(); }
        explicit operator bool() const noexcept { return has_value(); }
        Error&       error() & { return *m_error; }
        const Error& error() const& { return *m_error; }

    private:
        std::optional<Error> m_error;
    };
#endif

    template <typename T, typename... Args>
    Result<T> make_result(Args&&... args)
    {
#if defined(__cpp_lib_expected)
        return Result<T>{ std::in_place, std::forward<Args>(args)... };
#else
        return Result<T>{ std::forward<Args>(args)... };
#endif
    }

    template <typename T>
    Result<T> make_error(Error error)
    {
#if defined(__cpp_lib_expected)
        return Result<T>{ std::unexpect, error };
#else
        return Result<T>{ error };
#endif
    }

    // ------------------------------------------------------------------ small helpers (ref:common.hpp:260-412)
    inline std::string_view to_string(Error error) noexcept
    {
        switch (error) {
        case Error::Empty: return "Data is empty";
        case Error::TooShort: return "Data is too short";
        case Error::TooBig: return "Image is too big to process";
        case Error::NotQoi: return "Not a QOI file";
        case Error::InvalidDesc: return "Image description is invalid";
        case Error::MismatchedDesc: return "Image description does not match the data";
        case Error::NotEnoughSpace: return "Buffer does not have enough space";
        case Error::NotInitialized: return "Stream encoder/decoder is not initialized yet";
        case Error::AlreadyInitialized: return "Stream encoder/decoder already initialized";
        case Error::NotRegularFile: return "Not a regular file";
        case Error::FileExists: return "File already exists";
        case Error::FileNotExists: return "File does not exist";
        case Error::IoError: return "Unable to do read or write operation";
        case Error::BadAlloc: return "Failed to allocate memory";
        }
        return "Unknown";
    }

    template <std::integral T>
    constexpr std::optional<Channels> to_channels(T n) noexcept
    {
        if (n == 3) return Channels::RGB;
        if (n == 4) return Channels::RGBA;
        return std::nullopt;
    }

    template <std::integral T>
    constexpr std::optional<Colorspace> to_colorspace(T n) noexcept
    {
        if (n == 0) return Colorspace::sRGB;
        if (n == 1) return Colorspace::Linear;
        return std::nullopt;
    }

    // view `size` elements (bytes when T is void) as bytes
    template <typename T>
        requires std::same_as<T, void> or std::is_trivially_copyable_v<T>
    ByteCSpan to_span(T* t, std::size_t size)
    {
        if constexpr (std::same_as<T, void>) return { reinterpret_cast<const Byte*>(t), size };
        else return { reinterpret_cast<const Byte*>(t), size * sizeof(T) };
    }

    inline bool is_valid(const Desc& desc)
    {
        const bool ch = desc.channels == Channels::RGB or desc.channels == Channels::RGBA;
        const bool cs = desc.colorspace == Colorspace::sRGB or desc.colorspace == Colorspace::Linear;
        return desc.width > 0 and desc.height > 0 and ch and cs;
    }

    // width * height * channels, InvalidDesc / TooBig otherwise (ref:common.hpp:364-388)
    inline Result<std::size_t> count_bytes(const Desc& desc)
    {
        if (not is_valid(desc)) return make_error<std::size_t>(Error::InvalidDesc);
        const auto mul_overflows = [](std::size_t a, std::size_t b) { return a != 0 and (a * b) / a != b; };
        if (mul_overflows(desc.width, desc.height)) return make_error<std::size_t>(Error::TooBig);
        const auto pixels = static_cast<std::size_t>(desc.width) * desc.height;
        const auto bpp    = static_cast<std::size_t>(desc.channels);
        if (mul_overflows(pixels, bpp)) return make_error<std::size_t>(Error::TooBig);
        return pixels * bpp;
    }

    // (channels + 1) * width * height + header + end marker (ref:common.hpp:402-412)
    inline Result<std::size_t> worst_size(const Desc& desc)
    {
        if (const auto n = count_bytes(desc); not n) return make_error<std::size_t>(n.error());
        return (static_cast<std::size_t>(desc.channels) + 1) * desc.width * desc.height + constants::header_size
             + constants::end_marker_size;
    }

    Result<Desc> read_header(ByteCSpan in_data) noexcept;                     // ref:common.hpp:426, common.cpp:13-50
    Result<Desc> read_header(const std::filesystem::path& in_path) noexcept;  // ref:common.hpp:443, common.cpp:52-72

    // ------------------------------------------------------------------ one-shot API (ref:simple.hpp:23-324)
    Result<ByteVec> encode(ByteCSpan in_data, Desc desc) noexcept;
    Result<ByteVec> encode(PixelGenFun in_func, Desc desc) noexcept;

    // span targets: stores the longest whole-chunk prefix that fits, never NotEnoughSpace (ref:simple.hpp:52-61)
    Result<EncodeStatus> encode_into(ByteSpan out_buf, ByteCSpan in_data, Desc desc);
    Result<EncodeStatus> encode_into(ByteSpan out_buf, PixelGenFun in_func, Desc desc);
    Result<std::size_t>  encode_into(ByteSinkFun out_func, ByteCSpan in_data, Desc desc);
    Result<std::size_t>  encode_into(ByteSinkFun out_func, PixelGenFun in_func, Desc desc);
    Result<std::size_t>  encode_into(const std::filesystem::path& out_path, ByteCSpan in_data, Desc desc, bool overwrite = false) noexcept;
    Result<std::size_t>  encode_into(const std::filesystem::path& out_path, PixelGenFun in_func, Desc desc, bool overwrite = false) noexcept;

    Result<Image> decode(ByteCSpan in_data, std::optional<Channels> target = std::nullopt, bool flip_vertically = false) noexcept;
    Result<Image> decode(const std::filesystem::path& in_path, std::optional<Channels> target = std::nullopt, bool flip_vertically = false) noexcept;

    Result<Desc> decode_into(ByteSpan out_buf, ByteCSpan in_data, std::optional<Channels> target = std::nullopt, bool flip_vertically = false);
    Result<Desc> decode_into(PixelSinkFun out_func, ByteCSpan in_data);
    Result<Desc> decode_into(ByteSpan out_buf, const std::filesystem::path& in_path, std::optional<Channels> target = std::nullopt, bool flip_vertically = false) noexcept;
    Result<Desc> decode_into(PixelSinkFun out_func, const std::filesystem::path& in_path) noexcept;

    // ------------------------------------------------------------------ resumable API (ref:stream.hpp:23-244)
    // initialize() -> encode()... -> finalize().  Not thread-safe.  The object carries {channels, run, prev, 64-slot
    // table}; device workspace is cached per host thread, so the "never allocates" promise of the reference holds
    // only after the first call on a thread.
    class StreamEncoder
    {
    public:
        StreamEncoder() noexcept;

        Result<std::size_t>  initialize(ByteSpan out_buf, Desc desc) noexcept;         // writes the 14 header bytes
        Result<StreamResult> encode(ByteSpan out_buf, ByteCSpan in_buf) noexcept;       // out_buf.size() >= 5
        Result<std::size_t>  finalize(ByteSpan out_buf) noexcept;                       // pending run + end marker, resets
        void                 reset() noexcept;

        bool                    has_run_count() const noexcept { return m_run > 0; }
        std::optional<Channels> channels() const noexcept { return m_channels; }
        bool                    is_initialized() const noexcept { return m_channels.has_value(); }

    private:
        std::optional<Channels>                     m_channels;
        Byte                                        m_run;
        Pixel                                       m_prev;
        PixelArr<constants::running_array_size>     m_seen;
    };

    // initialize() -> decode()... -> drain_run()...  Not thread-safe.
    class StreamDecoder
    {
    public:
        StreamDecoder() noexcept;

        Result<Desc>         initialize(ByteCSpan in_buf, std::optional<Channels> target = std::nullopt) noexcept;
        Result<StreamResult> decode(ByteSpan out_buf, ByteCSpan in_buf) noexcept;  // out_buf.size() >= target channels
        Result<std::size_t>  drain_run(ByteSpan out_buf) noexcept;
        void                 reset() noexcept;

        bool                    has_run_count() const noexcept { return m_run > 0; }
        Byte                    run_count() const noexcept { return m_run; }
        std::optional<Channels> channels() const noexcept { return m_channels; }
        std::optional<Channels> target() const noexcept { return m_target; }
        bool                    is_initialized() const noexcept { return m_channels.has_value(); }

    private:
        std::optional<Channels>                 m_channels;
        std::optional<Channels>                 m_target;
        Byte                                    m_run;
        Pixel                                   m_prev;
        PixelArr<constants::running_array_size> m_seen;
    };

    // ---- extension of the B200 backend (not in the reference): GPU selection.  Every host thread owns one device context
    // per GPU it uses.  The GPU of a call is, in this order: set_device() of the calling thread, the environment variable
    // QOIPP_B200_DEVICE, the thread's current CUDA device (cudaGetDevice, 0 unless the application changed it).
    namespace b200
    {
        void set_device(int device) noexcept;  // < 0: back to the default rule
        int  device() noexcept;
        int  device_count() noexcept;
    }
}

#endif
