/*
 * qoi_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the codec loops of mrizaln/qoipp v0.5.0
 * (the reference).  It exists so that the CUDA path can be checked bit-for-bit on machines
 * where /root/reference is absent.  Nothing in the product library (qoipp_b200/csrc, include/)
 * may include, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the reference's own golden fixtures (test/resources/image_{raw,qoi}_{3,4}.txt and the two
 *       *_incomplete prefixes, the 1007-byte partial-encode boundary, the 5..1024 stream sweep), and
 *   (2) the unmodified reference compiled from /root/reference/source/{common,simple,stream}.cpp
 *       into oracle/_ref/libqoipp_ref.so (recipe: oracle/Makefile), on randomized inputs.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 */
#ifndef QOI_ORACLE_H
#define QOI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* qoipp::Error, include/qoipp/common.hpp:78-94 (values start at 1; 0 = success here). */
enum {
    QO_OK = 0,
    QO_EMPTY = 1,
    QO_TOO_SHORT,
    QO_TOO_BIG,
    QO_NOT_QOI,
    QO_INVALID_DESC,
    QO_MISMATCHED_DESC,
    QO_NOT_ENOUGH_SPACE,
    QO_NOT_INITIALIZED,
    QO_ALREADY_INITIALIZED,
    QO_NOT_REGULAR_FILE,
    QO_FILE_EXISTS,
    QO_FILE_NOT_EXISTS,
    QO_IO_ERROR,
    QO_BAD_ALLOC
};

typedef struct {
    uint32_t width, height;
    uint8_t  channels;   /* 3 | 4 */
    uint8_t  colorspace; /* 0 | 1 */
} qo_desc;

typedef struct {
    uint8_t r, g, b, a;
} qo_pixel;

/* resumable codec state == members of StreamEncoder / StreamDecoder
 * (include/qoipp/stream.hpp:112-115, 239-243) */
typedef struct {
    uint8_t  channels; /* 0 = not initialised */
    uint8_t  target;   /* decoder only */
    uint8_t  run;
    qo_pixel prev;
    qo_pixel seen[64];
} qo_state;

/* common.hpp:346-412, common.cpp:13-50 */
int qo_is_valid(const qo_desc* d);
int qo_count_bytes(const qo_desc* d, size_t* out);
int qo_worst_size(const qo_desc* d, size_t* out);
int qo_read_header(const uint8_t* in, size_t size, qo_desc* out);

/* impl::encode<Checked> (source/simple.cpp:17-98). `checked` != 0 selects the bounds-checked
 * ChunkArray; returns bytes written, *complete as EncodeStatus::complete. */
size_t qo_encode_core(const uint8_t* raw, const qo_desc* d, uint8_t* out, size_t cap, int checked, int* complete);

/* impl::decode (source/simple.cpp:100-171) writing `target` channels per pixel.  Defined
 * behaviour only: stops after width*height pixels (the reference keeps looping while
 * data_index < size-22, which is out-of-bounds UB -- SURVEY hazard 2). */
void qo_decode_core(const uint8_t* in, size_t size, uint32_t width, uint32_t height, uint8_t target, uint8_t* out);

/* qoipp::encode_into(ByteSpan, ByteCSpan, Desc) (source/simple.cpp:231-252): validation order,
 * Checked iff cap < worst_size. */
int qo_encode_into(uint8_t* out, size_t cap, const uint8_t* raw, size_t raw_size, const qo_desc* d, size_t* written,
                   int* complete);

/* qoipp::decode_into(ByteSpan, ByteCSpan, target, flip) (source/simple.cpp:444-494).  target = 0 keeps the
 * source channels.  Hazard 3 is resolved by additionally refusing when cap < w*h*target. */
int qo_decode_into(uint8_t* out, size_t cap, const uint8_t* in, size_t size, uint8_t target, int flip, qo_desc* desc);

/* StreamEncoder (source/stream.cpp:105-277) */
void qo_senc_init(qo_state* s);
int  qo_senc_initialize(qo_state* s, uint8_t* out, size_t cap, const qo_desc* d, size_t* written);
int  qo_senc_encode(qo_state* s, uint8_t* out, size_t cap, const uint8_t* in, size_t in_size, size_t* processed,
                    size_t* written);
int  qo_senc_finalize(qo_state* s, uint8_t* out, size_t cap, size_t* written);
void qo_senc_reset(qo_state* s);

/* StreamDecoder (source/stream.cpp:282-458) */
void qo_sdec_init(qo_state* s);
int  qo_sdec_initialize(qo_state* s, const uint8_t* in, size_t size, uint8_t target, qo_desc* desc);
int  qo_sdec_decode(qo_state* s, uint8_t* out, size_t cap, const uint8_t* in, size_t in_size, size_t* processed,
                    size_t* written);
int  qo_sdec_drain_run(qo_state* s, uint8_t* out, size_t cap, size_t* written);
void qo_sdec_reset(qo_state* s);

#ifdef __cplusplus
}
#endif
#endif
