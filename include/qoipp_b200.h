/*
 * qoipp_b200.h -- C ABI of the B200-native QOI hot path (libqoipp_b200.so).
 *
 * This is the drop-in boundary: plain C types, caller-owned buffers, no CUDA or torch types in any
 * signature (a CUDA stream travels as `void*`).  Each entry point names the reference interface it
 * replaces (mrizaln/qoipp v0.5.0, paths relative to the reference root).  The C++20 `qoipp::` API in
 * include/qoipp/ is a thin layer over these calls; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Return value of every call: 0 = ok, 1..14 = the qoipp::Error enumerator of include/qoipp/common.hpp:78-94
 * (same numbering), negative = -(cudaError_t) for device failures the reference has no enumerator for.
 * There is no CPU fallback anywhere behind this header: without a CUDA device the calls fail.
 *
 * Pointers named d_* are device pointers, h_* host pointers.  *_dev calls are asynchronous on `stream`;
 * their results are fetched with the matching *_status call, which synchronises the stream.
 */
#ifndef QOIPP_B200_H
#define QOIPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QOIPP_B200_VERSION 100 /* 0.1.0 */

/* qoipp::Desc (include/qoipp/common.hpp:114-122) */
typedef struct qoipp_b200_desc {
    uint32_t width;
    uint32_t height;
    uint8_t  channels;   /* 3 = RGB, 4 = RGBA   (qoipp::Channels,   common.hpp:66-70) */
    uint8_t  colorspace; /* 0 = sRGB, 1 = Linear (qoipp::Colorspace, common.hpp:54-58) */
} qoipp_b200_desc;

/* Resumable codec state: the members of qoipp::StreamEncoder / StreamDecoder
 * (include/qoipp/stream.hpp:112-115 and :239-243).  Pixels are r,g,b,a bytes. */
typedef struct qoipp_b200_state {
    uint8_t channels; /* 0 = not initialised */
    uint8_t target;   /* decoder only */
    uint8_t run;
    uint8_t reserved;
    uint8_t prev[4];
    uint8_t seen[64][4];
} qoipp_b200_state;

/* The same state resident in device memory (the resumable *_dev entry points): pixels packed r | g<<8 | b<<16 | a<<24,
 * `run` = pending run pixels (decoder) / current run length (encoder). */
typedef struct qoipp_b200_dev_state {
    uint32_t prev;
    uint32_t run;
    uint32_t seen[64];
} qoipp_b200_dev_state;

/* StreamResult{processed, written} (include/qoipp/common.hpp:148-152), both in bytes, written by the device */
typedef struct qoipp_b200_stream_result {
    uint64_t processed;
    uint64_t written;
} qoipp_b200_stream_result;

/* Opaque per-thread context: device workspace for the tile carries, pinned staging, result slots.
 * One context serves one host thread / stream at a time; create as many as there are driving threads. */
typedef struct qoipp_b200_ctx qoipp_b200_ctx;

int32_t     qoipp_b200_version(void);
const char* qoipp_b200_error_string(int32_t code); /* qoipp::to_string(Error), common.hpp:260-280 */
int32_t     qoipp_b200_device_count(void);
int32_t     qoipp_b200_current_device(void); /* cudaGetDevice of the calling thread (0 when there is none) */

int32_t qoipp_b200_ctx_create(int32_t device, qoipp_b200_ctx** out);
int32_t qoipp_b200_ctx_destroy(qoipp_b200_ctx* ctx);

/* ---- host-side helpers that never touch the device (common.hpp:346-412, common.cpp:13-50) */
int32_t qoipp_b200_count_bytes(const qoipp_b200_desc* desc, uint64_t* out);
int32_t qoipp_b200_worst_size(const qoipp_b200_desc* desc, uint64_t* out);
int32_t qoipp_b200_read_header(const uint8_t* h_qoi, uint64_t size, qoipp_b200_desc* out);

/* ---- one-shot encode: replaces impl::encode<Checked> (source/simple.cpp:17-98) behind
 * qoipp::encode_into(ByteSpan, ByteCSpan, Desc) (source/simple.cpp:231-252).
 * Stores the largest whole-chunk prefix of the stream that fits `out_cap` (header and end marker are
 * chunks).  d_raw must hold width*height*channels bytes. */
int32_t qoipp_b200_encode_dev(qoipp_b200_ctx* ctx, const uint8_t* d_raw, const qoipp_b200_desc* desc, uint8_t* d_out,
                              uint64_t out_cap, void* stream);
/* EncodeStatus{written, complete} (common.hpp:142-146) of the last encode issued on this context */
int32_t qoipp_b200_encode_status(qoipp_b200_ctx* ctx, void* stream, uint64_t* written, int32_t* complete);

/* same, host buffers: validation order of simple.cpp:231-252, H2D -> kernels -> D2H of `written` bytes */
int32_t qoipp_b200_encode_host(qoipp_b200_ctx* ctx, const uint8_t* h_raw, uint64_t raw_size, const qoipp_b200_desc* desc,
                               uint8_t* h_out, uint64_t out_cap, uint64_t* written, int32_t* complete);

/* ---- staged forms for callers that allocate their result (qoipp::encode(ByteCSpan, Desc), source/simple.cpp:178-205, and
 * qoipp::decode(ByteCSpan, target, flip), :365-414): the result stays in the context's device staging buffer until
 * qoipp_b200_fetch_staged copies its first n bytes to h_out.  encode_staged returns the exact size, so the caller allocates
 * `written` bytes instead of worst_size; decode_staged returns as soon as the work is enqueued, so the caller's allocation of
 * the image overlaps the transfer and the kernels. */
int32_t qoipp_b200_encode_staged(qoipp_b200_ctx* ctx, const uint8_t* h_raw, uint64_t raw_size, const qoipp_b200_desc* desc,
                                 uint64_t* written);
int32_t qoipp_b200_decode_staged(qoipp_b200_ctx* ctx, const uint8_t* h_qoi, uint64_t qoi_size, uint8_t target_channels,
                                 int32_t flip_vertically, qoipp_b200_desc* desc, uint64_t* out_bytes);
int32_t qoipp_b200_fetch_staged(qoipp_b200_ctx* ctx, uint8_t* h_out, uint64_t n);

/* ---- batch encode (extension): n_images equally shaped images, image k at d_raw + k*raw_stride,
 * output k at d_out + k*out_stride with capacity out_cap each; d_written[k] / d_complete[k] are device arrays
 * filled by the kernels (either may be NULL). */
int32_t qoipp_b200_encode_batch_dev(qoipp_b200_ctx* ctx, const uint8_t* d_raw, uint64_t raw_stride, uint32_t n_images,
                                    const qoipp_b200_desc* desc, uint8_t* d_out, uint64_t out_stride, uint64_t out_cap,
                                    uint64_t* d_written, void* stream);

/* same, host buffers (page-locked buffers are used in place; h_written[k] may be NULL); returns when the outputs are in h_out */
int32_t qoipp_b200_encode_batch_host(qoipp_b200_ctx* ctx, const uint8_t* h_raw, uint64_t raw_stride, uint32_t n_images,
                                     const qoipp_b200_desc* desc, uint8_t* h_out, uint64_t out_stride, uint64_t out_cap,
                                     uint64_t* h_written);

/* ---- resumable encode: replaces StreamEncoder::encode (source/stream.cpp:138-239).  `state` is carried by the
 * caller between calls; in_size is truncated to whole pixels (stream.cpp:59). */
/* device buffers; asynchronous on `stream`: *d_state is read and replaced, *d_result is written; no host synchronisation */
int32_t qoipp_b200_stream_encode_dev(qoipp_b200_ctx* ctx, uint8_t channels, qoipp_b200_dev_state* d_state, const uint8_t* d_in,
                                     uint64_t in_size, uint8_t* d_out, uint64_t out_cap, qoipp_b200_stream_result* d_result,
                                     void* stream);
int32_t qoipp_b200_stream_encode_host(qoipp_b200_ctx* ctx, qoipp_b200_state* state, const uint8_t* h_in, uint64_t in_size,
                                      uint8_t* h_out, uint64_t out_cap, uint64_t* processed, uint64_t* written);

/* ---- one-shot decode: replaces impl::decode (source/simple.cpp:100-171) behind
 * qoipp::decode_into(ByteSpan, ByteCSpan, target, flip) (source/simple.cpp:444-494).
 * `desc` is the parsed header of the stream (qoipp_b200_read_header); target_channels 0 keeps desc->channels;
 * d_out receives width*height*target bytes, rows bottom-up when flip_vertically != 0. */
int32_t qoipp_b200_decode_dev(qoipp_b200_ctx* ctx, const uint8_t* d_qoi, uint64_t qoi_size, const qoipp_b200_desc* desc,
                              uint8_t target_channels, int32_t flip_vertically, uint8_t* d_out, uint64_t out_cap,
                              void* stream);
/* 0 when the last decode on this context completed; *path: 0 = verified in round 0, 1..4 = retry rounds used,
 * + 100 = the sequential loop produced part of the image */
int32_t qoipp_b200_decode_status(qoipp_b200_ctx* ctx, void* stream, int32_t* path);

/* same for the last batch decode: paths[k] of image k (n_images as given to the batch call) */
int32_t qoipp_b200_decode_status_batch(qoipp_b200_ctx* ctx, void* stream, int32_t* paths, uint32_t n_images);

int32_t qoipp_b200_decode_host(qoipp_b200_ctx* ctx, const uint8_t* h_qoi, uint64_t qoi_size, uint8_t target_channels,
                               int32_t flip_vertically, uint8_t* h_out, uint64_t out_cap, qoipp_b200_desc* desc);

/* ---- batch decode (extension): stream k is d_qoi + offsets[k] .. offsets[k+1] (h_offsets has n_images+1 entries),
 * all with the same `desc`; output k at d_out + k*out_stride. */
int32_t qoipp_b200_decode_batch_dev(qoipp_b200_ctx* ctx, const uint8_t* d_qoi, const uint64_t* h_offsets, uint32_t n_images,
                                    const qoipp_b200_desc* desc, uint8_t target_channels, uint8_t* d_out,
                                    uint64_t out_stride, void* stream);

/* same with stream k at d_qoi + k*in_stride, h_sizes[k] bytes long: the layout qoipp_b200_encode_batch_dev writes
 * (out_stride, d_written), so an encoded batch is decoded where it lies */
int32_t qoipp_b200_decode_batch_strided_dev(qoipp_b200_ctx* ctx, const uint8_t* d_qoi, uint64_t in_stride, const uint64_t* h_sizes,
                                            uint32_t n_images, const qoipp_b200_desc* desc, uint8_t target_channels, uint8_t* d_out,
                                            uint64_t out_stride, void* stream);
/* host buffers (page-locked ones are used in place); returns when the pixels are in h_out */
int32_t qoipp_b200_decode_batch_host(qoipp_b200_ctx* ctx, const uint8_t* h_qoi, uint64_t in_stride, const uint64_t* h_sizes,
                                     uint32_t n_images, const qoipp_b200_desc* desc, uint8_t target_channels, uint8_t* h_out,
                                     uint64_t out_stride);

/* ---- resumable decode: replaces StreamDecoder::decode / drain_run (source/stream.cpp:312-447).
 * Inputs of a few KB and more are decoded by the same parallel tile kernel as a whole image, with `state` as the carry-in
 * and the carry-out taken behind the last consumed op; an incomplete op at the end of the input is not consumed
 * (stream.cpp:341-392).  Shorter inputs, and calls whose content refutes the kernel's speculation, take a sequential loop. */
/* device buffers; asynchronous on `stream`: *d_state is read and replaced, *d_result is written; no host synchronisation */
int32_t qoipp_b200_stream_decode_dev(qoipp_b200_ctx* ctx, uint8_t channels, qoipp_b200_dev_state* d_state, const uint8_t* d_in,
                                     uint64_t in_size, uint8_t* d_out, uint64_t out_cap, qoipp_b200_stream_result* d_result,
                                     void* stream);
int32_t qoipp_b200_stream_decode_host(qoipp_b200_ctx* ctx, qoipp_b200_state* state, const uint8_t* h_in, uint64_t in_size,
                                      uint8_t* h_out, uint64_t out_cap, uint64_t* processed, uint64_t* written);

#ifdef __cplusplus
}
#endif
#endif
