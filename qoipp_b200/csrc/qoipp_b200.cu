// qoipp_b200.cu -- the C ABI of include/qoipp_b200.h: context/workspace management and kernel launches.
// Compiled for sm_100a only (see __graft_entry__.build()).  No CPU fallback: every codec entry point needs a
// CUDA device and fails with a negative cudaError_t (or BadAlloc) otherwise.
#include "../../include/qoipp_b200.h"

#include "decode_wt.cuh"
#include "encode_kernel.cuh"
#include "encode_ts.cuh"
#include "host_util.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

using namespace qb;
namespace H = qb::host;

namespace
{
#ifndef QB_ENC_K
#define QB_ENC_K 8
#endif
    constexpr int kEncK = QB_ENC_K;  // pixels per thread per tile: tile = 256 * K pixels

    int32_t cuda_code(cudaError_t e)
    {
        if (e == cudaSuccess) return 0;
        if (e == cudaErrorMemoryAllocation) return H::BadAlloc;
        return -(int32_t)e;
    }
#define QB_CUDA(expr)                                  \
    do {                                               \
        cudaError_t qb_e_ = (expr);                    \
        if (qb_e_ != cudaSuccess) {                    \
            (void)cudaGetLastError();                  \
            return cuda_code(qb_e_);                   \
        }                                              \
    } while (0)

    struct DevBuf {
        void*  p   = nullptr;
        size_t cap = 0;
        // grow-only; contents are NOT preserved.  `zero` clears the new allocation, ordered on the call's stream `s`
        // (the kernels that read it are launched on `s`; a plain cudaMemset would run on the legacy default stream,
        // unordered against a non-blocking stream).  Growing waits for `s` first: an earlier launch may still use the old buffer.
        cudaError_t reserve(size_t n, cudaStream_t s, bool zero = false)
        {
            if (n <= cap) return cudaSuccess;
            if (p) {
                cudaError_t e = cudaStreamSynchronize(s);
                if (e != cudaSuccess) return e;
                cudaFree(p);
            }
            p = nullptr, cap = 0;
            size_t      want = std::max<size_t>(n + n / 4, 4096);
            cudaError_t e    = cudaMalloc(&p, want);
            if (e != cudaSuccess) { p = nullptr; return e; }
            cap = want;
            if (zero) return cudaMemsetAsync(p, 0, want, s);
            return cudaSuccess;
        }
        void release()
        {
            if (p) cudaFree(p);
            p = nullptr, cap = 0;
        }
    };

    struct PinnedBuf {
        void*  p   = nullptr;
        size_t cap = 0;
        cudaError_t reserve(size_t n)
        {
            if (n <= cap) return cudaSuccess;
            if (p) cudaFreeHost(p);
            p = nullptr, cap = 0;
            size_t      want = std::max<size_t>(n + n / 4, 4096);
            cudaError_t e    = cudaMallocHost(&p, want);
            if (e != cudaSuccess) { p = nullptr; return e; }
            cap = want;
            return cudaSuccess;
        }
        void release()
        {
            if (p) cudaFreeHost(p);
            p = nullptr, cap = 0;
        }
    };
    // Worker threads of a context for the host side of pageable transfers: a pageable buffer reaches the GPU through
    // cudaMemcpy at the speed of ONE CPU copy thread (measured 12 GB/s H2D, 17.5 GB/s D2H on the B200 box, against 51 / 54
    // GB/s from page-locked memory).  The staging pipeline below copies through a page-locked ring with several threads.
    class CopyPool
    {
    public:
        explicit CopyPool(unsigned n_threads)
        {
            for (unsigned i = 0; i < n_threads; ++i) th_.emplace_back([this, i] { work(i); });
        }
        ~CopyPool()
        {
            {
                std::lock_guard<std::mutex> l(m_);
                stop_ = true;
                ++gen_;
            }
            cv_.notify_all();
            for (auto& t : th_) t.join();
        }
        // memcpy(dst, src, n) split over the workers and the calling thread
        void copy(void* dst, const void* src, size_t n)
        {
            const size_t parts = th_.size() + 1, slice = (n / parts + 63) & ~size_t(63);
            if (th_.empty() || n < (256u << 10)) {
                std::memcpy(dst, src, n);
                return;
            }
            {
                std::lock_guard<std::mutex> l(m_);
                dst_ = static_cast<uint8_t*>(dst), src_ = static_cast<const uint8_t*>(src), n_ = n, slice_ = slice;
                pending_ = (unsigned)th_.size();
                ++gen_;
            }
            cv_.notify_all();
            const size_t off = slice * th_.size();  // the caller takes the last slice
            if (off < n) std::memcpy(dst_ + off, src_ + off, n - off);
            std::unique_lock<std::mutex> l(m_);
            done_.wait(l, [this] { return pending_ == 0; });
        }

    private:
        void work(unsigned i)
        {
            unsigned seen = 0;
            for (;;) {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                const size_t off = slice_ * i, len = off < n_ ? std::min(slice_, n_ - off) : 0;
                uint8_t*       d = dst_;
                const uint8_t* s = src_;
                l.unlock();
                if (len) std::memcpy(d + off, s + off, len);
                l.lock();
                if (--pending_ == 0) done_.notify_one();
            }
        }
        std::vector<std::thread> th_;
        std::mutex               m_;
        std::condition_variable  cv_, done_;
        uint8_t*                 dst_ = nullptr;
        const uint8_t*           src_ = nullptr;
        size_t                   n_ = 0, slice_ = 0;
        unsigned                 gen_ = 0, pending_ = 0;
        bool                     stop_ = false;
    };
}  // namespace

struct qoipp_b200_ctx {
    int      device = 0;
    int      sm_count = 148;
    uint32_t epoch  = 0;  // launch counter for the cross-CTA words (20 bits, see qb_common.cuh)
    DevBuf   carry;       // tile carry words of the kernel in flight (encode or decode)
    DevBuf   tickets;     // [0] encode ticket, [1..] decode tickets
    DevBuf   results;     // EncResult[n_images] / DecResult
    DevBuf   state;       // EncState / DecState carry-in for the resumable calls
    DevBuf   aux;         // decode: per-image offsets and first-tile ids of a batch
    DevBuf   fix;         // decode: per-tile lists of alphas learned by the retry rounds
    DevBuf   reqs;        // decode: cascade requests of round 0
    DevBuf   scratch;     // encode_ts_kernel: per-tile records, read by encode_ts_copy_kernel
    DevBuf   counts;      // encode_ts_kernel: byte counts per tile
    DevBuf   groups[2];   // encode_ts_kernel: totals per 64-tile group; the two take turns, each cleared by the other's copy kernel
    size_t   groups_dirty[2] = { 0, 0 };  // leading words of groups[i] that are not known to be zero
    int      groups_cur = 0;
    DevBuf   stage_in, stage_out;  // device staging of the host-pointer calls
    PinnedBuf h_result;   // pinned landing zone for result structs
    PinnedBuf h_pin_in, h_pin_out;
    std::vector<uint64_t> batch_table;      // decode_batch_dev: the offset / first-tile table of this call (host)
    uint64_t staged_bytes = 0;  // valid bytes of stage_out left by the last *_staged call
    cudaStream_t batch_table_stream = nullptr;  // the stream that upload was ordered on
    size_t   batch_table_dev_bytes = 0;     // bytes of the table `aux` holds (a copy of it stays in h_pin_in), 0 = none
    PinnedBuf ring;              // page-locked ring of the pageable staging pipeline (2 slots)
    cudaEvent_t ring_ev[2] = { nullptr, nullptr };
    CopyPool* pool = nullptr;    // created with the first large pageable transfer
    unsigned  copy_threads = 3;  // QOIPP_B200_COPY_THREADS (0 = plain cudaMemcpyAsync from / to pageable memory)
    cudaStream_t own_stream = nullptr;
    bool     enc_trivial = false;  // last encode needed no launch (capacity below the header)
    bool     attrs_set   = false;
    uint32_t ts_ticket = 0;  // encode_ts_kernel: current value of its ticket counter
    bool     force_general = false;  // QOIPP_B200_GENERAL=1: always the general kernels (A/B measurements, tests)
    uint64_t stream_parallel_min = 4 * kDecTB;  // resumable decode: shorter inputs take the sequential loop (QOIPP_B200_STREAM_PARALLEL_MIN)
    int      dec_coresident = 148;  // CTAs of decode_finish_kernel that fit on the device at once

    // reserves `count` consecutive epochs and returns with `epoch` = the first; the carry buffer (and the tagged
    // learned-alpha lists) are cleared when the 20-bit counter would wrap or the buffer was (re)allocated
    cudaError_t next_epoch(size_t carry_bytes, cudaStream_t s, unsigned count = 1)
    {
        const bool grew = carry_bytes > carry.cap;
        if (grew) {
            cudaError_t e = carry.reserve(carry_bytes, s, true);
            if (e != cudaSuccess) return e;
        }
        epoch += span;  // skip the epochs the previous call reserved
        span = count;
        if (epoch == 0 || epoch + count > kEpochMask) {
            cudaError_t e = cudaMemsetAsync(carry.p, 0, carry.cap, s);
            if (e == cudaSuccess && fix.p) e = cudaMemsetAsync(fix.p, 0, fix.cap, s);
            if (e != cudaSuccess) return e;
            epoch = 1;
        }
        return cudaSuccess;
    }
    unsigned span = 1;
};

namespace
{
    template <typename Kern>
    cudaError_t allow_smem(Kern k, size_t bytes)
    {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        // the kernels hide look-back latency with resident CTAs: give shared memory the whole carve-out
        return cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }

    cudaError_t set_attrs(qoipp_b200_ctx* c)
    {
        if (c->attrs_set) return cudaSuccess;
        cudaError_t e;
        if ((e = allow_smem(encode_kernel<3, kEncK>, sizeof(EncSmem<kEncK>))) != cudaSuccess) return e;
        if ((e = allow_smem(encode_kernel<4, kEncK>, sizeof(EncSmem<kEncK>))) != cudaSuccess) return e;
        if ((e = allow_smem(encode_ts_kernel<3>, kTsWarps * sizeof(TsWarpSmem))) != cudaSuccess) return e;
        if ((e = allow_smem(encode_ts_kernel<4>, kTsWarps * sizeof(TsWarpSmem))) != cudaSuccess) return e;
        if ((e = allow_smem(encode_ts_copy_kernel<3>, kTsCopyWarps * sizeof(TsCopySmem))) != cudaSuccess) return e;
        if ((e = allow_smem(encode_ts_copy_kernel<4>, kTsCopyWarps * sizeof(TsCopySmem))) != cudaSuccess) return e;
        if ((e = dec_set_attrs()) != cudaSuccess) return e;
        int per_sm = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_finish_kernel, kWtThreads, kWtSmemBytes)) != cudaSuccess) return e;
        int per_sm_stream = 0;  // both cooperative kernels are launched with the same grid: it must be co-resident for either
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_stream, decode_finish_stream_kernel, kWtThreads, kWtSmemBytes)) != cudaSuccess) return e;
        per_sm = std::min(per_sm, per_sm_stream);
        if (const char* g = std::getenv("QOIPP_B200_DEC_CTAS")) per_sm = std::max(1, std::min(per_sm, std::atoi(g)));  // A/B measurements
        c->dec_coresident = std::max(1, per_sm) * c->sm_count;
        c->attrs_set = true;
        return cudaSuccess;
    }

    struct Guard {  // selects the context's device for the duration of a call
        int prev = -1;
        explicit Guard(int dev)
        {
            cudaGetDevice(&prev);
            if (prev != dev) cudaSetDevice(dev);
            else prev = -1;
        }
        ~Guard()
        {
            if (prev >= 0) cudaSetDevice(prev);
        }
    };

    // ---- encode launch shared by the one-shot, batch and resumable entry points
    int32_t launch_encode(qoipp_b200_ctx* c, const uint8_t* d_in, uint64_t in_stride, uint32_t n_images, uint64_t n_pixels,
                          unsigned ch, const uint8_t* header14, uint8_t* d_out, uint64_t out_stride, uint64_t out_cap,
                          uint32_t flags, const EncState* d_init, cudaStream_t s, bool buffers_on_device = true)
    {
        QB_CUDA(set_attrs(c));
        // Thread-serial kernels (encode_ts.cuh) for the plain one-shot case: nothing can be refused (capacity >= worst size,
        // util.hpp:240-246 never triggers), no carried state, every image 16-byte aligned.  Everything else -- partial
        // buffers, the resumable form, unaligned spans -- takes the general kernel.  So do page-locked HOST buffers used in
        // place: the general kernel reads and writes in one pass, so both PCIe directions are busy at once, while the
        // encode + copy pair would use them one after the other (measured: 4K RGB e2e 37 GB/s against 27 GB/s).
        const bool worst_fits = n_pixels <= (~0ull - (H::kHeaderSize + H::kMarkerSize)) / (ch + 1);  // no wrapped comparison below
        const bool ts = flags == 0 && d_init == nullptr && worst_fits && out_cap >= n_pixels * (ch + 1) + H::kHeaderSize + H::kMarkerSize &&
                        (reinterpret_cast<uintptr_t>(d_in) & 15u) == 0 && (n_images == 1 || (in_stride & 15u) == 0) &&
                        buffers_on_device && !c->force_general;
        const uint64_t T     = ts ? (uint64_t)kTsT : (uint64_t)kEncThreads * kEncK;
        const uint64_t tiles = (n_pixels + T - 1) / T;
        if (tiles * n_images >= (1ull << 31)) return H::TooBig;
        QB_CUDA(c->results.reserve(sizeof(EncResult) * n_images, s));
        QB_CUDA(c->tickets.reserve(64, s, true));
        QB_CUDA(c->next_epoch(tiles * n_images * kEncDescWords * sizeof(uint64_t), s));
        EncParams P{};
        P.in = d_in, P.out = d_out;
        P.n_pixels = n_pixels, P.in_stride = in_stride, P.out_stride = out_stride, P.out_cap = out_cap;
        P.tiles_per_image = (uint32_t)tiles, P.n_images = n_images, P.epoch = c->epoch, P.flags = flags;
        if (header14) std::memcpy(P.header, header14, 14);
        P.init_state = d_init;
        P.results    = static_cast<EncResult*>(c->results.p);
        P.desc       = static_cast<uint64_t*>(c->carry.p);
        P.ticket     = static_cast<uint32_t*>(c->tickets.p);
        const dim3 grid((unsigned)(tiles * n_images));
        if (ts) {
            const uint64_t n_tiles   = tiles * n_images;
            const uint64_t scr_words = ch == 3 ? TsCfg<3>::kScrWords : TsCfg<4>::kScrWords;
            const uint64_t groups    = (tiles + 63) / 64;
            QB_CUDA(c->scratch.reserve(n_tiles * scr_words * 4, s));
            QB_CUDA(c->counts.reserve(n_tiles * 4, s));
            P.scratch          = static_cast<uint32_t*>(c->scratch.p);
            P.tile_bytes       = static_cast<uint32_t*>(c->counts.p);
            P.groups_per_image = (uint32_t)groups;
            // group totals: this encode adds into one buffer (zero by now) and its copy kernel clears what the previous encode
            // left in the other, so that no memset sits between two encodes
            {
                const int    cur = c->groups_cur, oth = cur ^ 1;
                const size_t need = groups * n_images;
                if (need * 4 > c->groups[cur].cap) {
                    QB_CUDA(c->groups[cur].reserve(need * 4, s, true));  // new memory arrives cleared
                    c->groups_dirty[cur] = 0;
                }
                if (c->groups_dirty[cur]) {  // first use after a call that could not be followed by a copy kernel
                    QB_CUDA(cudaMemsetAsync(c->groups[cur].p, 0, c->groups_dirty[cur] * 4, s));
                    c->groups_dirty[cur] = 0;
                }
                P.group_bytes = static_cast<uint32_t*>(c->groups[cur].p);
                P.zero_ptr    = static_cast<uint32_t*>(c->groups[oth].p);
                P.zero_n      = (uint32_t)c->groups_dirty[oth];
                c->groups_dirty[oth] = 0, c->groups_dirty[cur] = need, c->groups_cur = oth;
            }
            // persistent warps: one CTA per resident slot; the ticket counter is never reset, every warp draws exactly one
            // ticket beyond the last tile
            const unsigned n_ctas = (unsigned)std::min<uint64_t>((n_tiles + kTsWarps - 1) / kTsWarps, (uint64_t)c->sm_count * QB_TS_CTAS);
            P.ticket         = static_cast<uint32_t*>(c->tickets.p) + 8;
            P.ticket_base[0] = c->ts_ticket;
            c->ts_ticket += (uint32_t)n_tiles + n_ctas * kTsWarps;
            const dim3 grid2((unsigned)((n_tiles + kTsCopyWarps - 1) / kTsCopyWarps)), block2(kTsCopyWarps * 32);
            if (ch == 3) {
                encode_ts_kernel<3><<<dim3(n_ctas), dim3(kTsThreads), kTsWarps * sizeof(TsWarpSmem), s>>>(P);
                encode_ts_copy_kernel<3><<<grid2, block2, kTsCopyWarps * sizeof(TsCopySmem), s>>>(P);
            } else {
                encode_ts_kernel<4><<<dim3(n_ctas), dim3(kTsThreads), kTsWarps * sizeof(TsWarpSmem), s>>>(P);
                encode_ts_copy_kernel<4><<<grid2, block2, kTsCopyWarps * sizeof(TsCopySmem), s>>>(P);
            }
        } else {
            if (ch == 3) encode_kernel<3, kEncK><<<grid, dim3(kEncThreads), sizeof(EncSmem<kEncK>), s>>>(P);
            else encode_kernel<4, kEncK><<<grid, dim3(kEncThreads), sizeof(EncSmem<kEncK>), s>>>(P);
        }
        if (cudaError_t le = cudaGetLastError(); le != cudaSuccess) {
            // the host-side shadow of the ticket counter was advanced for a launch that did not happen: start both from zero again
            (void)cudaMemsetAsync(c->tickets.p, 0, 64, s);
            c->ts_ticket = 0;
            for (int i = 0; i < 2; ++i) c->groups_dirty[i] = c->groups[i].cap / 4;  // neither buffer of group totals is known to be clear
            return cuda_code(le);
        }
        return 0;
    }

    // A host pointer the device can dereference (page-locked memory under unified addressing: cudaHostAlloc /
    // cudaHostRegister, e.g. torch pinned tensors), else nullptr.  Such buffers are read and written by the kernels
    // directly over PCIe (zero-copy), so transfer and codec work overlap tile by tile with no staging copy.
    template <class T>
    T* mapped_host(T* p)
    {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        if (a.type == cudaMemoryTypeHost && a.devicePointer) return static_cast<T*>(a.devicePointer);
        return nullptr;
    }

    // ---- pageable staging pipeline: chunks travel through a page-locked ring of two slots; the CPU copy of
    // chunk k + 1 (several threads) overlaps the DMA of chunk k.  Everything is enqueued on `s`.
    // Slot size: an eighth of the transfer, between 1 and 8 MiB (measured: 8 MiB slots are best for a 133 MB image, 22-25 GB/s,
    // but leave a 25 MB image with three chunks and no overlap; per-chunk cost is an event wait and a pool wake-up).
    constexpr size_t kRingSlot = 8u << 20;
    size_t ring_chunk(size_t n) { return std::min(kRingSlot, std::max<size_t>(1u << 20, (n / 8 + 0xFFFF) & ~size_t(0xFFFF))); }

    cudaError_t ring_ready(qoipp_b200_ctx* c)
    {
        if (cudaError_t e = c->ring.reserve(2 * kRingSlot); e != cudaSuccess) return e;
        for (auto& ev : c->ring_ev)
            if (!ev)
                if (cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); e != cudaSuccess) return e;
        if (!c->pool) c->pool = new (std::nothrow) CopyPool(c->copy_threads);
        return c->pool ? cudaSuccess : cudaErrorMemoryAllocation;
    }

    bool use_ring(const qoipp_b200_ctx* c, size_t n) { return c->copy_threads > 0 && n >= (4u << 20); }

    cudaError_t pageable_to_device(qoipp_b200_ctx* c, void* d_dst, const void* h_src, size_t n, cudaStream_t s)
    {
        if (!use_ring(c, n)) return cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyDefault, s);
        if (cudaError_t e = ring_ready(c); e != cudaSuccess) return e;
        auto*        ring       = static_cast<uint8_t*>(c->ring.p);
        const size_t kRingChunk = ring_chunk(n);
        for (size_t off = 0, k = 0; off < n; off += kRingChunk, ++k) {
            const size_t   len  = std::min(kRingChunk, n - off);
            const unsigned slot = k & 1;
            if (k >= 2)
                if (cudaError_t e = cudaEventSynchronize(c->ring_ev[slot]); e != cudaSuccess) return e;  // its last DMA has read the slot
            c->pool->copy(ring + slot * kRingSlot, static_cast<const uint8_t*>(h_src) + off, len);
            if (cudaError_t e = cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off, ring + slot * kRingSlot, len, cudaMemcpyHostToDevice, s); e != cudaSuccess) return e;
            if (cudaError_t e = cudaEventRecord(c->ring_ev[slot], s); e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }

    // returns after the data is in h_dst (synchronises `s`)
    cudaError_t device_to_pageable(qoipp_b200_ctx* c, void* h_dst, const void* d_src, size_t n, cudaStream_t s)
    {
        if (!use_ring(c, n)) {
            if (cudaError_t e = cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDefault, s); e != cudaSuccess) return e;
            return cudaStreamSynchronize(s);
        }
        if (cudaError_t e = ring_ready(c); e != cudaSuccess) return e;
        auto*        ring       = static_cast<uint8_t*>(c->ring.p);
        const size_t kRingChunk = ring_chunk(n);
        const size_t total      = (n + kRingChunk - 1) / kRingChunk;
        for (size_t k = 0; k <= total; ++k) {
            if (k < total) {  // DMA of chunk k into its slot (the slot was emptied two rounds ago, by this thread)
                const size_t off = k * kRingChunk, len = std::min(kRingChunk, n - off);
                if (cudaError_t e = cudaMemcpyAsync(ring + (k & 1) * kRingSlot, static_cast<const uint8_t*>(d_src) + off, len, cudaMemcpyDeviceToHost, s); e != cudaSuccess) return e;
                if (cudaError_t e = cudaEventRecord(c->ring_ev[k & 1], s); e != cudaSuccess) return e;
            }
            if (k >= 1) {  // chunk k - 1 has arrived: out of the ring while chunk k is in flight
                const size_t j = k - 1, off = j * kRingChunk, len = std::min(kRingChunk, n - off);
                if (cudaError_t e = cudaEventSynchronize(c->ring_ev[j & 1]); e != cudaSuccess) return e;
                c->pool->copy(static_cast<uint8_t*>(h_dst) + off, ring + (j & 1) * kRingSlot, len);
            }
        }
        return cudaSuccess;
    }

    __global__ void stream_encode_epilogue_kernel(const EncResult* res, unsigned ch, StreamOut* out, EncState* state)
    {
        const unsigned i = threadIdx.x;
        if (i < 64u) state->table[i] = res->state.table[i];
        if (i == 0) state->prev = res->state.prev, state->run = res->state.run, out->processed = res->processed * ch, out->written = res->written;
    }

    unsigned pack_px(const uint8_t* p) { return p[0] | (unsigned)p[1] << 8 | (unsigned)p[2] << 16 | (unsigned)p[3] << 24; }
    void     unpack_px(unsigned v, uint8_t* p) { p[0] = (uint8_t)v, p[1] = (uint8_t)(v >> 8), p[2] = (uint8_t)(v >> 16), p[3] = (uint8_t)(v >> 24); }
}  // namespace

extern "C"
{
    int32_t     qoipp_b200_version(void) { return QOIPP_B200_VERSION; }
    const char* qoipp_b200_error_string(int32_t code) { return H::error_string(code); }

    int32_t qoipp_b200_device_count(void)
    {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
            (void)cudaGetLastError();
            return 0;
        }
        return n;
    }

    int32_t qoipp_b200_current_device(void)
    {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) {
            (void)cudaGetLastError();
            return 0;
        }
        return d;
    }

    int32_t qoipp_b200_count_bytes(const qoipp_b200_desc* desc, uint64_t* out) { return H::count_bytes(*desc, out); }
    int32_t qoipp_b200_worst_size(const qoipp_b200_desc* desc, uint64_t* out) { return H::worst_size(*desc, out); }
    int32_t qoipp_b200_read_header(const uint8_t* h_qoi, uint64_t size, qoipp_b200_desc* out) { return H::read_header(h_qoi, size, out); }

    int32_t qoipp_b200_ctx_create(int32_t device, qoipp_b200_ctx** out)
    {
        *out = nullptr;
        int n = 0;
        QB_CUDA(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) return -(int32_t)cudaErrorInvalidDevice;
        auto* c = new (std::nothrow) qoipp_b200_ctx();
        if (!c) return H::BadAlloc;
        c->device = device;
        Guard g(device);
        cudaDeviceProp prop{};
        cudaError_t    e = cudaGetDeviceProperties(&prop, device);
        if (e == cudaSuccess && prop.major < 10) {
            std::fprintf(stderr, "qoipp_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
            e = cudaErrorNoKernelImageForDevice;
        }
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = c->h_result.reserve(4096);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            delete c;
            return cuda_code(e);
        }
        c->sm_count = prop.multiProcessorCount;
        if (const char* g = std::getenv("QOIPP_B200_GENERAL")) c->force_general = g[0] == '1';
        if (const char* g = std::getenv("QOIPP_B200_STREAM_PARALLEL_MIN")) c->stream_parallel_min = (uint64_t)std::max(1ll, std::atoll(g));
        if (const char* g = std::getenv("QOIPP_B200_COPY_THREADS")) c->copy_threads = (unsigned)std::max(0, std::min(16, std::atoi(g)));
        c->copy_threads = std::min(c->copy_threads, std::max(1u, std::thread::hardware_concurrency()) - 1u);
        *out        = c;
        return 0;
    }

    int32_t qoipp_b200_ctx_destroy(qoipp_b200_ctx* c)
    {
        if (!c) return 0;
        Guard g(c->device);
        cudaDeviceSynchronize();
        c->carry.release(), c->tickets.release(), c->results.release(), c->state.release(), c->aux.release(), c->fix.release(), c->reqs.release(), c->scratch.release(), c->counts.release(), c->groups[0].release(), c->groups[1].release();
        c->stage_in.release(), c->stage_out.release();
        c->h_result.release(), c->h_pin_in.release(), c->h_pin_out.release(), c->ring.release();
        for (auto& e : c->ring_ev)
            if (e) cudaEventDestroy(e);
        delete c->pool;
        if (c->own_stream) cudaStreamDestroy(c->own_stream);
        delete c;
        return 0;
    }

    // ------------------------------------------------------------------ encode
    static int32_t encode_one(qoipp_b200_ctx* c, const uint8_t* d_raw, const qoipp_b200_desc* desc, uint8_t* d_out, uint64_t out_cap,
                              void* stream, bool buffers_on_device)
    {
        uint64_t raw;
        if (int32_t e = H::count_bytes(*desc, &raw)) return e;
        Guard g(c->device);
        c->enc_trivial = out_cap < H::kHeaderSize;  // util.hpp:125-131: the header chunk does not fit, nothing is stored
        if (c->enc_trivial) return 0;
        uint8_t hdr[14];
        H::write_header(*desc, hdr);
        return launch_encode(c, d_raw, 0, 1, (uint64_t)desc->width * desc->height, desc->channels, hdr, d_out, 0, out_cap, 0,
                             nullptr, static_cast<cudaStream_t>(stream), buffers_on_device);
    }

    int32_t qoipp_b200_encode_dev(qoipp_b200_ctx* c, const uint8_t* d_raw, const qoipp_b200_desc* desc, uint8_t* d_out,
                                  uint64_t out_cap, void* stream)
    {
        return encode_one(c, d_raw, desc, d_out, out_cap, stream, true);
    }

    int32_t qoipp_b200_encode_status(qoipp_b200_ctx* c, void* stream, uint64_t* written, int32_t* complete)
    {
        if (c->enc_trivial) {
            *written = 0, *complete = 0;
            return 0;
        }
        Guard g(c->device);
        auto  s = static_cast<cudaStream_t>(stream);
        auto* h = static_cast<EncResult*>(c->h_result.p);
        QB_CUDA(cudaMemcpyAsync(h, c->results.p, 24, cudaMemcpyDeviceToHost, s));
        QB_CUDA(cudaStreamSynchronize(s));
        *written  = h->written;
        *complete = (int32_t)h->complete;
        return 0;
    }

    int32_t qoipp_b200_encode_host(qoipp_b200_ctx* c, const uint8_t* h_raw, uint64_t raw_size, const qoipp_b200_desc* desc,
                                   uint8_t* h_out, uint64_t out_cap, uint64_t* written, int32_t* complete)
    {
        // validation order of qoipp::encode_into, source/simple.cpp:235-244
        if (raw_size == 0) return H::Empty;
        uint64_t need, worst;
        if (int32_t e = H::count_bytes(*desc, &need)) return e;
        if (raw_size != need) return H::MismatchedDesc;
        if (int32_t e = H::worst_size(*desc, &worst)) return e;
        Guard          g(c->device);
        const uint64_t cap = std::min(out_cap, worst);
        cudaStream_t   s   = c->own_stream;
        const uint8_t* d_in  = mapped_host(h_raw);
        uint8_t*       d_out = mapped_host(h_out);
        if (!d_in) {
            QB_CUDA(c->stage_in.reserve(raw_size + 16, s));
            QB_CUDA(pageable_to_device(c, c->stage_in.p, h_raw, raw_size, s));
            d_in = static_cast<uint8_t*>(c->stage_in.p);
        }
        const bool staged_out = d_out == nullptr;
        if (staged_out) {
            QB_CUDA(c->stage_out.reserve(cap + 16, s));
            d_out = static_cast<uint8_t*>(c->stage_out.p);
        }
        const bool in_place = mapped_host(h_raw) != nullptr || !staged_out;  // a kernel touches host memory directly
        if (int32_t e = encode_one(c, d_in, desc, d_out, cap, s, !in_place)) return e;
        if (int32_t e = qoipp_b200_encode_status(c, s, written, complete)) return e;
        if (staged_out && *written) QB_CUDA(device_to_pageable(c, h_out, c->stage_out.p, *written, s));
        return 0;
    }

    // ---- staged forms for callers that allocate their result (qoipp::encode / qoipp::decode): the output stays in the
    // context's device staging buffer until its size is known / while the caller allocates, then _fetch_staged brings it over
    int32_t qoipp_b200_encode_staged(qoipp_b200_ctx* c, const uint8_t* h_raw, uint64_t raw_size, const qoipp_b200_desc* desc, uint64_t* written)
    {
        if (raw_size == 0) return H::Empty;  // validation order of qoipp::encode, source/simple.cpp:182-188
        uint64_t need, worst;
        if (int32_t e = H::count_bytes(*desc, &need)) return e;
        if (raw_size != need) return H::MismatchedDesc;
        if (int32_t e = H::worst_size(*desc, &worst)) return e;
        Guard          g(c->device);
        cudaStream_t   s    = c->own_stream;
        const uint8_t* d_in = mapped_host(h_raw);
        if (!d_in) {
            QB_CUDA(c->stage_in.reserve(raw_size + 16, s));
            QB_CUDA(pageable_to_device(c, c->stage_in.p, h_raw, raw_size, s));
            d_in = static_cast<uint8_t*>(c->stage_in.p);
        }
        QB_CUDA(c->stage_out.reserve(worst + 16, s));
        if (int32_t e = encode_one(c, d_in, desc, static_cast<uint8_t*>(c->stage_out.p), worst, s, mapped_host(h_raw) == nullptr)) return e;
        int32_t complete = 0;
        if (int32_t e = qoipp_b200_encode_status(c, s, written, &complete)) return e;
        c->staged_bytes = *written;
        return 0;
    }

    int32_t qoipp_b200_fetch_staged(qoipp_b200_ctx* c, uint8_t* h_out, uint64_t n)
    {
        if (n > c->staged_bytes) return H::NotEnoughSpace;
        Guard        g(c->device);
        cudaStream_t s = c->own_stream;
        if (n == 0) {
            QB_CUDA(cudaStreamSynchronize(s));
            return 0;
        }
        QB_CUDA(device_to_pageable(c, h_out, c->stage_out.p, n, s));
        return 0;
    }

    int32_t qoipp_b200_encode_batch_dev(qoipp_b200_ctx* c, const uint8_t* d_raw, uint64_t raw_stride, uint32_t n_images,
                                        const qoipp_b200_desc* desc, uint8_t* d_out, uint64_t out_stride, uint64_t out_cap,
                                        uint64_t* d_written, void* stream)
    {
        uint64_t raw;
        if (int32_t e = H::count_bytes(*desc, &raw)) return e;
        if (n_images == 0) return H::Empty;
        if (out_cap < H::kHeaderSize) return H::NotEnoughSpace;
        Guard g(c->device);
        auto  s = static_cast<cudaStream_t>(stream);
        uint8_t hdr[14];
        H::write_header(*desc, hdr);
        c->enc_trivial = false;
        if (int32_t e = launch_encode(c, d_raw, raw_stride, n_images, (uint64_t)desc->width * desc->height, desc->channels, hdr,
                                      d_out, out_stride, out_cap, 0, nullptr, s))
            return e;
        if (d_written)  // EncResult[k].written -> d_written[k]
            QB_CUDA(cudaMemcpy2DAsync(d_written, sizeof(uint64_t), c->results.p, sizeof(EncResult), sizeof(uint64_t), n_images,
                                      cudaMemcpyDeviceToDevice, s));
        return 0;
    }

    int32_t qoipp_b200_encode_batch_host(qoipp_b200_ctx* c, const uint8_t* h_raw, uint64_t raw_stride, uint32_t n_images,
                                         const qoipp_b200_desc* desc, uint8_t* h_out, uint64_t out_stride, uint64_t out_cap,
                                         uint64_t* h_written)
    {
        uint64_t raw, worst;
        if (int32_t e = H::count_bytes(*desc, &raw)) return e;
        if (n_images == 0) return H::Empty;
        if (raw_stride < raw || out_stride < out_cap) return H::MismatchedDesc;
        if (out_cap < H::kHeaderSize) return H::NotEnoughSpace;
        if (int32_t e = H::worst_size(*desc, &worst)) return e;
        const uint64_t cap = std::min(out_cap, worst);
        Guard          g(c->device);
        cudaStream_t   s     = c->own_stream;
        const uint8_t* d_in  = mapped_host(h_raw);
        uint8_t*       d_out = mapped_host(h_out);
        const uint64_t in_bytes = raw_stride * (n_images - 1) + raw, out_bytes = out_stride * (n_images - 1) + cap;
        const bool     in_place = d_in != nullptr || d_out != nullptr;  // a kernel touches host memory directly
        if (!d_in) {
            QB_CUDA(c->stage_in.reserve(in_bytes + 16, s));
            QB_CUDA(pageable_to_device(c, c->stage_in.p, h_raw, in_bytes, s));
            d_in = static_cast<uint8_t*>(c->stage_in.p);
        }
        const bool staged_out = d_out == nullptr;
        if (staged_out) {
            QB_CUDA(c->stage_out.reserve(out_bytes + 16, s));
            d_out = static_cast<uint8_t*>(c->stage_out.p);
        }
        uint8_t hdr[14];
        H::write_header(*desc, hdr);
        c->enc_trivial = false;
        if (int32_t e = launch_encode(c, d_in, raw_stride, n_images, (uint64_t)desc->width * desc->height, desc->channels, hdr, d_out,
                                      out_stride, cap, 0, nullptr, s, !in_place))
            return e;
        QB_CUDA(c->h_pin_out.reserve(sizeof(uint64_t) * (size_t)n_images));
        QB_CUDA(cudaMemcpy2DAsync(c->h_pin_out.p, sizeof(uint64_t), c->results.p, sizeof(EncResult), sizeof(uint64_t), n_images,
                                  cudaMemcpyDeviceToHost, s));
        QB_CUDA(cudaStreamSynchronize(s));
        const auto* w = static_cast<const uint64_t*>(c->h_pin_out.p);
        if (h_written) std::memcpy(h_written, w, sizeof(uint64_t) * (size_t)n_images);
        if (staged_out) {  // only the bytes each image produced travel back
            for (uint32_t k = 0; k < n_images; ++k)
                if (w[k]) QB_CUDA(cudaMemcpyAsync(h_out + k * out_stride, d_out + k * out_stride, w[k], cudaMemcpyDeviceToHost, s));
            QB_CUDA(cudaStreamSynchronize(s));
        }
        return 0;
    }

    // ---- resumable encode on device buffers: everything is enqueued on `stream`, nothing waits for the device
    int32_t qoipp_b200_stream_encode_dev(qoipp_b200_ctx* c, uint8_t channels, qoipp_b200_dev_state* d_state, const uint8_t* d_in,
                                         uint64_t in_size, uint8_t* d_out, uint64_t out_cap, qoipp_b200_stream_result* d_result, void* stream)
    {
        if (channels != 3 && channels != 4) return H::NotInitialized;  // error order of StreamEncoder::encode, stream.cpp:140-146
        if (out_cap == 0 || in_size == 0) return H::Empty;
        if (out_cap < 5) return H::TooShort;
        const uint64_t n = in_size / channels;  // whole pixels only (stream.cpp:59)
        Guard g(c->device);
        auto  s = static_cast<cudaStream_t>(stream);
        if (n == 0) {
            QB_CUDA(cudaMemsetAsync(d_result, 0, sizeof(*d_result), s));
            return 0;
        }
        const uint64_t cap = std::min<uint64_t>(out_cap, n * (channels + 1) + 1);  // worst case of this input + a pending run flush
        c->enc_trivial = false;
        if (int32_t e = launch_encode(c, d_in, 0, 1, n, channels, nullptr, d_out, 0, cap, ENC_STREAM, reinterpret_cast<const EncState*>(d_state), s))
            return e;
        stream_encode_epilogue_kernel<<<1, 64, 0, s>>>(static_cast<const EncResult*>(c->results.p), channels, reinterpret_cast<StreamOut*>(d_result),
                                                       reinterpret_cast<EncState*>(d_state));
        QB_CUDA(cudaGetLastError());
        return 0;
    }

    int32_t qoipp_b200_stream_encode_host(qoipp_b200_ctx* c, qoipp_b200_state* st, const uint8_t* h_in, uint64_t in_size,
                                          uint8_t* h_out, uint64_t out_cap, uint64_t* processed, uint64_t* written)
    {
        // error order of StreamEncoder::encode, source/stream.cpp:140-146
        if (!st->channels) return H::NotInitialized;
        if (out_cap == 0 || in_size == 0) return H::Empty;
        if (out_cap < 5) return H::TooShort;
        const unsigned ch = st->channels;
        const uint64_t n  = in_size / ch;
        *processed = 0, *written = 0;
        if (n == 0) return 0;
        Guard g(c->device);
        // a call can never store more than the worst case of its input (+1 for a pending run flush)
        const uint64_t cap = std::min<uint64_t>(out_cap, n * (ch + 1) + 1);
        cudaStream_t s  = c->own_stream;
        QB_CUDA(c->stage_in.reserve(n * ch + 16, s));
        QB_CUDA(c->stage_out.reserve(cap + 16, s));
        QB_CUDA(c->state.reserve(sizeof(EncState) + sizeof(StreamOut), s));
        auto* d_state = static_cast<EncState*>(c->state.p);
        auto* d_sres  = reinterpret_cast<StreamOut*>(d_state + 1);
        auto* hs = reinterpret_cast<EncState*>(static_cast<uint8_t*>(c->h_result.p) + 1024);
        hs->prev = pack_px(st->prev), hs->run = st->run;
        for (int i = 0; i < 64; ++i) hs->table[i] = pack_px(st->seen[i]);
        QB_CUDA(cudaMemcpyAsync(d_state, hs, sizeof(EncState), cudaMemcpyHostToDevice, s));
        QB_CUDA(pageable_to_device(c, c->stage_in.p, h_in, n * ch, s));
        if (int32_t e = qoipp_b200_stream_encode_dev(c, (uint8_t)ch, reinterpret_cast<qoipp_b200_dev_state*>(d_state), static_cast<uint8_t*>(c->stage_in.p), n * ch,
                                                     static_cast<uint8_t*>(c->stage_out.p), cap, reinterpret_cast<qoipp_b200_stream_result*>(d_sres), s))
            return e;
        auto* hr = reinterpret_cast<EncState*>(static_cast<uint8_t*>(c->h_result.p) + 2048);
        QB_CUDA(cudaMemcpyAsync(hr, d_state, sizeof(EncState) + sizeof(StreamOut), cudaMemcpyDeviceToHost, s));
        QB_CUDA(cudaStreamSynchronize(s));
        const auto* ho = reinterpret_cast<const StreamOut*>(hr + 1);
        if (ho->written) QB_CUDA(device_to_pageable(c, h_out, c->stage_out.p, ho->written, s));
        *processed = ho->processed;
        *written   = ho->written;
        unpack_px(hr->prev, st->prev);
        st->run = (uint8_t)hr->run;
        for (int i = 0; i < 64; ++i) unpack_px(hr->table[i], st->seen[i]);
        return 0;
    }
}

#include "decode_host.inl"

#ifdef QB_STATS
extern "C" int32_t qoipp_b200_debug_stats(qoipp_b200_ctx* c, uint32_t* out4)  // development aid only
{
    out4[2] = out4[3] = 0;
    return cuda_code(cudaMemcpy(out4, static_cast<uint8_t*>(c->results.p) + offsetof(DecControl, stats), 8, cudaMemcpyDeviceToHost));
}
#endif
#ifdef QB_TIMING
extern "C" int32_t qoipp_b200_debug_carry(qoipp_b200_ctx* c, void** ptr, uint64_t* bytes)  // development aid only
{
    *ptr = c->carry.p, *bytes = c->carry.cap;
    return 0;
}
#endif
