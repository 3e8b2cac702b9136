// batch_multi_gpu.cpp -- a batch of images sharded by image over the GPUs of one box, one host thread per GPU
// (image k of B goes to GPU floor(k * G / B); nothing is exchanged between GPUs).
// Every thread selects its GPU with qoipp::b200::set_device and then uses the ordinary qoipp::encode / qoipp::decode.
//
//   batch_multi_gpu [images=64] [width=512] [height=512] [channels=4] [gpus=all] [threads_per_gpu=1]
// With threads_per_gpu > 1 several host threads drive the same GPU at once, each through its own context (the reference's
// free functions are re-entrant; so are these).
#include "qoipp/qoipp.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

static qoipp::ByteVec make_image(unsigned k, unsigned w, unsigned h, unsigned ch)
{
    qoipp::ByteVec raw((size_t)w * h * ch);
    uint64_t       s = 0x51F0 + k;
    for (size_t i = 0; i < raw.size(); i += ch) {
        if ((i / ch) % 7 == 0) s = s * 6364136223846793005ull + 1442695040888963407ull;  // short runs of one colour
        raw[i] = (uint8_t)(s >> 33), raw[i + 1] = (uint8_t)(s >> 41), raw[i + 2] = (uint8_t)(s >> 49);
        if (ch == 4) raw[i + 3] = 255;
    }
    return raw;
}

int main(int argc, char** argv)
{
    const unsigned B = argc > 1 ? std::atoi(argv[1]) : 64, w = argc > 2 ? std::atoi(argv[2]) : 512, h = argc > 3 ? std::atoi(argv[3]) : 512;
    const unsigned ch = argc > 4 ? std::atoi(argv[4]) : 4;
    int            G  = qoipp::b200::device_count();
    if (argc > 5 && std::atoi(argv[5]) > 0) G = std::min(G, std::atoi(argv[5]));
    const int T = argc > 6 ? std::max(1, std::atoi(argv[6])) : 1;
    if (G <= 0) {
        std::fprintf(stderr, "no CUDA device\n");
        return 2;
    }
    std::vector<qoipp::ByteVec> images;
    for (unsigned k = 0; k < B; ++k) images.push_back(make_image(k, w, h, ch));
    const qoipp::Desc desc{ w, h, static_cast<qoipp::Channels>(ch), qoipp::Colorspace::sRGB };

    std::vector<int>         failures(G, 0), served(G, 0);
    std::vector<std::thread> threads;
    const auto               t0 = std::chrono::steady_clock::now();
    std::vector<int> fail_t(G * T, 0), served_t(G * T, 0);
    for (int gt = 0; gt < G * T; ++gt)
        threads.emplace_back([&, gt] {
            const int g = gt / T, sub = gt % T;
            auto&     failures = fail_t;  // per thread slots: no sharing between threads
            auto&     served   = served_t;
            qoipp::b200::set_device(g);
            unsigned mine = 0;
            for (unsigned k = 0; k < B; ++k) {
                if ((int)((uint64_t)k * G / B) != g) continue;  // not this GPU's image
                if ((int)(mine++ % (unsigned)T) != sub) continue;  // not this thread's share of the GPU's images
                auto enc = qoipp::encode(images[k], desc);
                if (not enc) { ++failures[gt]; continue; }
                auto dec = qoipp::decode(*enc);
                if (not dec or dec->data != images[k] or qoipp::b200::device() != g) ++failures[gt];
                ++served[gt];
            }
        });
    for (auto& t : threads) t.join();
    for (int gt = 0; gt < G * T; ++gt) failures[gt / T] += fail_t[gt], served[gt / T] += served_t[gt];
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int          bad = 0, total = 0;
    for (int g = 0; g < G; ++g) {
        std::printf("gpu %d: %d images, %d failures\n", g, served[g], failures[g]);
        bad += failures[g], total += served[g];
    }
    std::printf("%d images on %d GPUs in %.3f s (%.2f GB/s raw, encode + decode), %d failures\n", total, G, sec,
                2.0 * total * w * h * ch / sec / 1e9, bad);
    return bad == 0 and total == (int)B ? 0 : 1;
}
