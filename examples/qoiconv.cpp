// qoiconv.cpp -- raw <-> .qoi conversion through the B200-backed qoipp:: C++ API (the on-disk container round trip of the
// reference's example/source/02_conv.cpp, without the PNG side: stb / fpng are third-party and out of scope).
//
//   qoiconv encode <in.raw> <width> <height> <channels 3|4> <out.qoi> [linear]
//   qoiconv decode <in.qoi> <out.raw> [channels 3|4] [flip]
//   qoiconv info   <in.qoi>
//
// Every codec call runs the CUDA kernels behind include/qoipp/qoipp.hpp; errors are the reference's qoipp::Error values.
#include <qoipp/simple.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iterator>
#include <string>

namespace fs = std::filesystem;

static int fail(const char* what, qoipp::Error e)
{
    std::fprintf(stderr, "qoiconv: %s: %s\n", what, qoipp::to_string(e).data());
    return 1;
}

static int usage()
{
    std::fprintf(stderr, "usage: qoiconv encode <in.raw> <w> <h> <3|4> <out.qoi> [linear]\n"
                         "       qoiconv decode <in.qoi> <out.raw> [3|4] [flip]\n"
                         "       qoiconv info   <in.qoi>\n");
    return 2;
}

int main(int argc, char** argv)
{
    if (argc < 3) return usage();
    const std::string cmd = argv[1];
    if (cmd == "info") {
        auto d = qoipp::read_header(fs::path{ argv[2] });
        if (!d) return fail("read_header", d.error());
        std::printf("%u x %u, %d channels, %s\n", d->width, d->height, (int)d->channels,
                    d->colorspace == qoipp::Colorspace::sRGB ? "sRGB" : "linear");
        return 0;
    }
    if (cmd == "encode") {
        if (argc < 7) return usage();
        std::ifstream f(argv[2], std::ios::binary);
        if (!f) { std::fprintf(stderr, "qoiconv: cannot open %s\n", argv[2]); return 1; }
        qoipp::ByteVec raw(std::istreambuf_iterator<char>(f), {});
        qoipp::Desc    desc{ (unsigned)std::strtoul(argv[3], nullptr, 10), (unsigned)std::strtoul(argv[4], nullptr, 10),
                             std::atoi(argv[5]) == 4 ? qoipp::Channels::RGBA : qoipp::Channels::RGB,
                             argc > 7 && std::strcmp(argv[7], "linear") == 0 ? qoipp::Colorspace::Linear : qoipp::Colorspace::sRGB };
        auto n = qoipp::encode_into(fs::path{ argv[6] }, raw, desc, true);
        if (!n) return fail("encode_into", n.error());
        std::printf("%zu -> %zu bytes\n", raw.size(), *n);
        return 0;
    }
    if (cmd == "decode") {
        if (argc < 4) return usage();
        std::optional<qoipp::Channels> target;
        bool                           flip = false;
        for (int i = 4; i < argc; ++i) {
            if (std::strcmp(argv[i], "3") == 0) target = qoipp::Channels::RGB;
            else if (std::strcmp(argv[i], "4") == 0) target = qoipp::Channels::RGBA;
            else if (std::strcmp(argv[i], "flip") == 0) flip = true;
        }
        auto img = qoipp::decode(fs::path{ argv[2] }, target, flip);
        if (!img) return fail("decode", img.error());
        std::ofstream o(argv[3], std::ios::binary | std::ios::trunc);
        o.write(reinterpret_cast<const char*>(img->data.data()), (std::streamsize)img->data.size());
        if (!o) { std::fprintf(stderr, "qoiconv: cannot write %s\n", argv[3]); return 1; }
        std::printf("%u x %u x %d -> %zu bytes\n", img->desc.width, img->desc.height, (int)img->desc.channels, img->data.size());
        return 0;
    }
    return usage();
}
