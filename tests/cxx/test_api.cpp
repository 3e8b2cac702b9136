// test_api.cpp -- replays the assertions of the reference's own tests (test/source/simple_test.cpp and
// test/source/stream_test.cpp of mrizaln/qoipp v0.5.0) against the B200-backed qoipp:: C++ API, with a
// dependency-free harness.  Usage: test_api <dir with image_{raw,qoi}_{3,4}.bin, image_qoi_{3,4}_incomplete.bin>
// Driven by tests/test_gpu_cxx_api.py (needs a GPU: there is no CPU path to test).
#include <qoipp/simple.hpp>
#include <qoipp/stream.hpp>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <string>
#include <thread>
#include <unistd.h>

namespace fs = std::filesystem;
using qoipp::Byte;
using qoipp::ByteCSpan;
using qoipp::ByteSpan;
using qoipp::ByteVec;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        ++g_checks;                                                              \
        if (!(cond)) {                                                           \
            ++g_fail;                                                            \
            std::fprintf(stderr, "FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
        }                                                                        \
    } while (0)
#define REQUIRE(cond)                                                                    \
    do {                                                                                 \
        ++g_checks;                                                                      \
        if (!(cond)) {                                                                   \
            std::fprintf(stderr, "FATAL %s:%d  %s\n", __FILE__, __LINE__, #cond);        \
            std::exit(2);                                                                \
        }                                                                                \
    } while (0)

static ByteVec read_file(const fs::path& p)
{
    std::ifstream f(p, std::ios::binary);
    REQUIRE(f.is_open());
    return ByteVec(std::istreambuf_iterator<char>(f), {});
}

static ByteVec to_rgb(ByteCSpan d)  // test/source/util.hpp:61-75
{
    ByteVec r;
    for (size_t i = 0; i + 3 < d.size() + 0; i += 4) r.insert(r.end(), d.begin() + i, d.begin() + i + 3);
    return r;
}
static ByteVec to_rgba(ByteCSpan d)  // test/source/util.hpp:77-92
{
    ByteVec r;
    for (size_t i = 0; i + 2 < d.size(); i += 3) {
        r.insert(r.end(), d.begin() + i, d.begin() + i + 3);
        r.push_back(0xFF);
    }
    return r;
}
static bool eq(ByteCSpan a, ByteCSpan b) { return a.size() == b.size() && std::equal(a.begin(), a.end(), b.begin()); }

struct Case {
    qoipp::Desc desc;
    ByteVec     raw, qoi, qoi_incomplete;
};

constexpr size_t chunk_boundary = 1007;  // simple_test.cpp:24-25

static void simple_tests(const Case& c)
{
    const auto& [desc, raw, qoi, incomplete] = c;
    const bool rgba = desc.channels == qoipp::Channels::RGBA;
    auto gen = [&](std::size_t i) -> qoipp::Pixel {
        auto o = i * static_cast<size_t>(desc.channels);
        return { raw[o], raw[o + 1], raw[o + 2], rgba ? raw[o + 3] : Byte{ 0xFF } };
    };

    {  // simple image encode (:77-83)
        auto e = qoipp::encode(raw, desc);
        REQUIRE(e.has_value());
        CHECK(eq(*e, qoi));
        auto e2 = qoipp::encode(gen, desc);
        REQUIRE(e2.has_value());
        CHECK(eq(*e2, qoi));
    }
    {  // encode into buffer from span / function (:85-143)
        for (int from_fn = 0; from_fn < 2; ++from_fn) {
            ByteVec big(qoipp::worst_size(desc).value());
            auto    r = from_fn ? qoipp::encode_into(big, gen, desc) : qoipp::encode_into(big, raw, desc);
            REQUIRE(r.has_value());
            CHECK(r->complete);
            CHECK(eq(ByteCSpan{ big.data(), r->written }, qoi));

            ByteVec small(chunk_boundary);
            auto    p = from_fn ? qoipp::encode_into(small, gen, desc) : qoipp::encode_into(small, raw, desc);
            REQUIRE(p.has_value());
            CHECK(!p->complete);
            CHECK(p->written == chunk_boundary);
            CHECK(eq(small, ByteCSpan{ qoi.data(), chunk_boundary }));
        }
    }
    {  // encode into byte sink (:145-177)
        ByteVec got;
        auto    r = qoipp::encode_into([&](Byte b) { got.push_back(b); }, raw, desc);
        REQUIRE(r.has_value());
        CHECK(*r == qoi.size());
        CHECK(eq(got, qoi));
        got.clear();
        auto r2 = qoipp::encode_into([&](Byte b) { got.push_back(b); }, gen, desc);
        REQUIRE(r2.has_value());
        CHECK(*r2 == qoi.size() && eq(got, qoi));
    }
    {  // decode, wants RGB, wants RGBA (:179-210)
        auto d = qoipp::decode(qoi);
        REQUIRE(d.has_value());
        CHECK(d->desc == desc);
        CHECK(eq(d->data, raw));
        auto d3 = qoipp::decode(qoi, qoipp::Channels::RGB);
        REQUIRE(d3.has_value());
        CHECK(d3->desc.channels == qoipp::Channels::RGB);
        CHECK(eq(d3->data, rgba ? to_rgb(raw) : raw));
        auto d4 = qoipp::decode(qoi, qoipp::Channels::RGBA);
        REQUIRE(d4.has_value());
        CHECK(d4->desc.channels == qoipp::Channels::RGBA);
        CHECK(eq(d4->data, rgba ? raw : to_rgba(raw)));
        auto flipped = qoipp::decode(qoi, std::nullopt, true);
        REQUIRE(flipped.has_value());
        const size_t line = desc.width * static_cast<size_t>(desc.channels);
        bool         ok   = true;
        for (size_t y = 0; y < desc.height; ++y)
            ok = ok && std::equal(raw.begin() + y * line, raw.begin() + (y + 1) * line, flipped->data.begin() + (desc.height - 1 - y) * line);
        CHECK(ok);
    }
    {  // decode into buffer / pixel sink (:212-242)
        ByteVec buf(raw.size());
        auto    r = qoipp::decode_into(buf, qoi);
        REQUIRE(r.has_value());
        CHECK(*r == desc);
        CHECK(eq(buf, raw));
        ByteVec small(raw.size() - 1);
        auto    s = qoipp::decode_into(small, qoi);
        CHECK(!s.has_value() && s.error() == qoipp::Error::NotEnoughSpace);
        ByteVec got;
        auto    k = qoipp::decode_into(
            [&](qoipp::Pixel p) {
                got.push_back(p.r), got.push_back(p.g), got.push_back(p.b);
                if (rgba) got.push_back(p.a);
                else CHECK(p.a == 0xFF);
            },
            qoi);
        REQUIRE(k.has_value());
        CHECK(*k == desc);
        CHECK(eq(got, raw));
    }
    {  // file round trip (:244-280)
        auto path = fs::temp_directory_path() / ("qoipp_b200_test_" + std::to_string(::getpid()) + "_" + std::to_string((int)desc.channels) + ".qoi");
        fs::remove(path);
        auto w = qoipp::encode_into(path, raw, desc, false);
        REQUIRE(w.has_value());
        CHECK(*w == qoi.size());
        CHECK(eq(read_file(path), qoi));
        auto again = qoipp::encode_into(path, raw, desc, false);
        CHECK(!again.has_value() && again.error() == qoipp::Error::FileExists);
        auto over = qoipp::encode_into(path, gen, desc, true);
        CHECK(over.has_value() && *over == qoi.size());
        auto d = qoipp::decode(path);
        REQUIRE(d.has_value());
        CHECK(d->desc == desc && eq(d->data, raw));
        ByteVec buf(raw.size());
        auto    di = qoipp::decode_into(buf, path);
        CHECK(di.has_value() && eq(buf, raw));
        size_t n  = 0;
        auto   ds = qoipp::decode_into([&](qoipp::Pixel) { ++n; }, path);
        CHECK(ds.has_value() && n == desc.width * desc.height);
        auto hp = qoipp::read_header(path);
        CHECK(hp.has_value() && *hp == desc);
        fs::remove(path);

        std::ofstream(path).close();  // empty file
        auto e = qoipp::decode(path);
        CHECK(!e.has_value() && e.error() == qoipp::Error::Empty);
        fs::remove(path);
        auto m = qoipp::decode(path);
        CHECK(!m.has_value() && m.error() == qoipp::Error::FileNotExists);
        auto bad = qoipp::encode_into(path, raw, qoipp::Desc{ 0, 1, desc.channels, desc.colorspace });
        CHECK(!bad.has_value() && bad.error() == qoipp::Error::InvalidDesc);
        CHECK(!fs::exists(path));  // no file is created on failure
        auto dir = qoipp::decode(fs::temp_directory_path());
        CHECK(!dir.has_value() && dir.error() == qoipp::Error::NotRegularFile);
    }
    {  // header read (:282-295)
        auto h = qoipp::read_header(qoi);
        CHECK(h.has_value() && *h == desc);
        auto e = qoipp::read_header(ByteCSpan{});
        CHECK(!e.has_value() && e.error() == qoipp::Error::Empty);
        Byte junk[4] = { 1, 2, 3, 4 };
        auto t       = qoipp::read_header(ByteCSpan{ junk, 4 });
        CHECK(!t.has_value() && t.error() == qoipp::Error::TooShort);
        ByteVec notqoi(qoi);
        notqoi[0] = 'x';
        auto nq   = qoipp::decode(notqoi);
        CHECK(!nq.has_value() && nq.error() == qoipp::Error::NotQoi);
    }
    {  // decode incomplete data (:316-322)
        auto d = qoipp::decode(incomplete);
        REQUIRE(d.has_value());
        CHECK(d->desc == desc);
        CHECK(d->data.size() == raw.size());
    }
    {  // error order of encode (source/simple.cpp:182-188)
        auto e0 = qoipp::encode(ByteCSpan{}, desc);
        CHECK(!e0 && e0.error() == qoipp::Error::Empty);
        auto e1 = qoipp::encode(raw, qoipp::Desc{ desc.width, 0, desc.channels, desc.colorspace });
        CHECK(!e1 && e1.error() == qoipp::Error::InvalidDesc);
        auto e2 = qoipp::encode(ByteCSpan{ raw.data(), raw.size() - 1 }, desc);
        CHECK(!e2 && e2.error() == qoipp::Error::MismatchedDesc);
    }
}

// the canonical calling protocol, stream_test.cpp:43-123
static ByteVec stream_encode(qoipp::StreamEncoder& enc, qoipp::Desc desc, ByteSpan out, ByteCSpan input)
{
    REQUIRE(!enc.is_initialized());
    ByteVec encoded(qoipp::constants::header_size);
    auto    res = enc.initialize(encoded, desc);
    REQUIRE(res.has_value() && *res == qoipp::constants::header_size);
    size_t off = 0;
    while (off < input.size()) {
        auto in = ByteCSpan{ input.data() + off, std::min(out.size(), input.size() - off) };
        auto r  = enc.encode(out, in);
        REQUIRE(r.has_value());
        off += r->processed;
        encoded.insert(encoded.end(), out.begin(), out.begin() + static_cast<long>(r->written));
    }
    auto size = encoded.size(), extra = qoipp::constants::end_marker_size + enc.has_run_count();
    encoded.resize(size + extra);
    auto fin = enc.finalize({ encoded.data() + size, extra });
    REQUIRE(fin.has_value() && *fin == extra);
    return encoded;
}

static ByteVec stream_decode(qoipp::StreamDecoder& dec, qoipp::Desc ref, ByteSpan out, ByteCSpan input, std::optional<qoipp::Channels> target = std::nullopt)
{
    REQUIRE(!dec.is_initialized());
    ByteVec decoded;
    if (target) ref.channels = *target;
    auto parsed = dec.initialize({ input.data(), qoipp::constants::header_size }, target);
    REQUIRE(parsed.has_value());
    CHECK(*parsed == ref);
    size_t off = qoipp::constants::header_size, end = input.size() - qoipp::constants::end_marker_size;
    while (off < end) {
        auto in = ByteCSpan{ input.data() + off, std::min(out.size(), end - off) };
        auto r  = dec.decode(out, in);
        REQUIRE(r.has_value());
        off += r->processed;
        decoded.insert(decoded.end(), out.begin(), out.begin() + static_cast<long>(r->written));
    }
    while (dec.has_run_count()) {
        auto n = dec.drain_run(out).value();
        decoded.insert(decoded.end(), out.begin(), out.begin() + static_cast<long>(n));
    }
    dec.reset();
    return decoded;
}

static void stream_tests(const Case cases[2], unsigned stride)
{
    qoipp::StreamEncoder enc;  // one object reused for everything (stream_test.cpp:188-190)
    qoipp::StreamDecoder dec;
    for (unsigned i = 5; i <= 1024; i += stride) {
        for (int k = 0; k < 2; ++k) {
            const auto& c = cases[k];
            const bool  rgba = c.desc.channels == qoipp::Channels::RGBA;
            ByteVec     buf(i);
            CHECK(eq(stream_encode(enc, c.desc, buf, c.raw), c.qoi));
            CHECK(eq(stream_decode(dec, c.desc, buf, c.qoi), c.raw));
            CHECK(eq(stream_decode(dec, c.desc, buf, c.qoi, qoipp::Channels::RGB), rgba ? to_rgb(c.raw) : c.raw));
            CHECK(eq(stream_decode(dec, c.desc, buf, c.qoi, qoipp::Channels::RGBA), rgba ? c.raw : to_rgba(c.raw)));
            auto part = stream_decode(dec, c.desc, buf, c.qoi_incomplete);
            CHECK(part.size() != c.raw.size());
            CHECK(part.size() <= c.raw.size() && std::equal(part.begin(), part.end(), c.raw.begin()));
        }
    }
    // error order (source/stream.cpp:113-146, 290-320)
    Byte     tiny[4] = {};
    ByteVec  room(64);
    auto     ne = enc.encode(room, room);
    CHECK(!ne && ne.error() == qoipp::Error::NotInitialized);
    CHECK(enc.initialize(room, cases[1].desc).has_value());
    auto twice = enc.initialize(room, cases[1].desc);
    CHECK(!twice && twice.error() == qoipp::Error::AlreadyInitialized);
    auto em = enc.encode(room, ByteCSpan{});
    CHECK(!em && em.error() == qoipp::Error::Empty);
    auto ts = enc.encode(ByteSpan{ tiny, 4 }, room);
    CHECK(!ts && ts.error() == qoipp::Error::TooShort);
    enc.reset();
    auto nd = dec.decode(room, room);
    CHECK(!nd && nd.error() == qoipp::Error::NotInitialized);
    auto dd = dec.drain_run(room);
    CHECK(!dd && dd.error() == qoipp::Error::NotInitialized);
}

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <fixture dir> [stream sweep stride]\n", argv[0]);
        return 2;
    }
    const fs::path dir    = argv[1];
    const unsigned stride = argc > 2 ? (unsigned)std::atoi(argv[2]) : 1;
    Case cases[2] = {
        { { 29, 17, qoipp::Channels::RGB, qoipp::Colorspace::sRGB }, read_file(dir / "image_raw_3.bin"), read_file(dir / "image_qoi_3.bin"), read_file(dir / "image_qoi_3_incomplete.bin") },
        { { 24, 14, qoipp::Channels::RGBA, qoipp::Colorspace::sRGB }, read_file(dir / "image_raw_4.bin"), read_file(dir / "image_qoi_4.bin"), read_file(dir / "image_qoi_4_incomplete.bin") },
    };
    for (const auto& c : cases) simple_tests(c);
    stream_tests(cases, stride);

    // free functions are re-entrant: thread-per-image, every thread gets its own device context
    {
        std::vector<std::thread> th;
        std::vector<int>         ok(6, 0);
        for (int t = 0; t < 6; ++t)
            th.emplace_back([&, t] {
                const auto& c = cases[t & 1];
                bool        good = true;
                for (int r = 0; r < 20; ++r) {
                    auto e = qoipp::encode(c.raw, c.desc);
                    good   = good && e.has_value() && eq(*e, c.qoi);
                    auto d = qoipp::decode(c.qoi);
                    good   = good && d.has_value() && eq(d->data, c.raw);
                }
                ok[t] = good;
            });
        for (auto& x : th) x.join();
        for (int v : ok) CHECK(v == 1);
    }
    std::printf("%d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
