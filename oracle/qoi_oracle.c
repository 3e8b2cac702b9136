/*
 * qoi_oracle.c -- TEST INFRASTRUCTURE ONLY (see qoi_oracle.h).  Parity status: PINNED.
 *
 * Scalar restatement of the reference (mrizaln/qoipp v0.5.0) codec loops.  Written as a
 * cursor-based emitter rather than the reference's adaptor templates; the observable
 * behaviour (bytes, counts, error codes, state after back-out) is what is restated.
 */
#include "qoi_oracle.h"

#include <string.h>

enum { OP_INDEX = 0x00, OP_DIFF = 0x40, OP_LUMA = 0x80, OP_RUN = 0xC0, OP_RGB = 0xFE, OP_RGBA = 0xFF };
enum { HEADER = 14, MARKER = 8, RUN_LIMIT = 62 }; /* common.hpp:17-23, util.hpp:27-43 */

static const qo_pixel START = { 0, 0, 0, 255 }; /* util.hpp:42 */

/* util.hpp:347-351 (caller reduces mod 64) */
static unsigned slot_of(qo_pixel p) { return (p.r * 3u + p.g * 5u + p.b * 7u + p.a * 11u) & 63u; }
static int      same(qo_pixel a, qo_pixel b) { return a.r == b.r && a.g == b.g && a.b == b.b && a.a == b.a; }

/* ---------------------------------------------------------------- descriptors */

int qo_is_valid(const qo_desc* d) /* common.hpp:346-352 */
{
    return d->width > 0 && d->height > 0 && (d->channels == 3 || d->channels == 4) && d->colorspace <= 1;
}

int qo_count_bytes(const qo_desc* d, size_t* out) /* common.hpp:364-388 */
{
    if (!qo_is_valid(d)) return QO_INVALID_DESC;
    size_t px = (size_t)d->width * d->height; /* u32*u32 never overflows a 64-bit size_t */
    if (sizeof(size_t) < 8 && d->width != 0 && px / d->width != d->height) return QO_TOO_BIG;
    size_t bytes = px * d->channels;
    if (bytes / d->channels != px) return QO_TOO_BIG;
    *out = bytes;
    return QO_OK;
}

int qo_worst_size(const qo_desc* d, size_t* out) /* common.hpp:402-412 */
{
    size_t n;
    int    e = qo_count_bytes(d, &n);
    if (e) return e;
    *out = ((size_t)d->channels + 1) * d->width * d->height + HEADER + MARKER;
    return QO_OK;
}

int qo_read_header(const uint8_t* in, size_t size, qo_desc* out) /* common.cpp:13-50 */
{
    if (size == 0) return QO_EMPTY;
    if (size < HEADER) return QO_TOO_SHORT;
    if (memcmp(in, "qoif", 4) != 0) return QO_NOT_QOI;
    uint32_t w = (uint32_t)in[4] << 24 | (uint32_t)in[5] << 16 | (uint32_t)in[6] << 8 | in[7];
    uint32_t h = (uint32_t)in[8] << 24 | (uint32_t)in[9] << 16 | (uint32_t)in[10] << 8 | in[11];
    if ((in[12] != 3 && in[12] != 4) || in[13] > 1 || w == 0 || h == 0) return QO_INVALID_DESC;
    out->width      = w;
    out->height     = h;
    out->channels   = in[12];
    out->colorspace = in[13];
    return QO_OK;
}

/* ---------------------------------------------------------------- chunk emitter */

/* util::ChunkArray<Out, Checked> (util.hpp:116-252): a chunk is stored only if all of it fits;
 * the first refusal latches `ok` to 0 and every later chunk is refused too. */
typedef struct {
    uint8_t* out;
    size_t   cap, pos;
    int      checked, ok;
} emitter;

static int emit(emitter* e, const uint8_t* bytes, size_t n)
{
    if (e->checked && (!e->ok || e->pos + n > e->cap)) return e->ok = 0;
    memcpy(e->out + e->pos, bytes, n);
    e->pos += n;
    return 1;
}

static void emit_header(emitter* e, const qo_desc* d) /* util.hpp:125-149 */
{
    uint8_t b[HEADER] = { 'q', 'o', 'i', 'f' };
    for (int i = 0; i < 4; ++i) {
        b[4 + i] = (uint8_t)(d->width >> (24 - 8 * i));
        b[8 + i] = (uint8_t)(d->height >> (24 - 8 * i));
    }
    b[12] = d->channels;
    b[13] = d->colorspace;
    emit(e, b, HEADER);
}

static void emit_marker(emitter* e) /* util.hpp:151-161, :41 */
{
    static const uint8_t m[MARKER] = { 0, 0, 0, 0, 0, 0, 0, 1 };
    emit(e, m, MARKER);
}

static void emit_run(emitter* e, unsigned run) /* util.hpp:227-235 */
{
    uint8_t b = (uint8_t)(OP_RUN | (run - 1));
    emit(e, &b, 1);
}

/* Colour chunk selection for a pixel that is neither a run continuation nor an index hit:
 * simple.cpp:59-79 / stream.cpp:183-214; byte layouts util.hpp:163-225. */
static void emit_colour(emitter* e, qo_pixel cur, qo_pixel prev, int rgba_input)
{
    uint8_t b[5];
    if (rgba_input && cur.a != prev.a) {
        b[0] = OP_RGBA, b[1] = cur.r, b[2] = cur.g, b[3] = cur.b, b[4] = cur.a;
        emit(e, b, 5);
        return;
    }
    int8_t dr = (int8_t)(cur.r - prev.r), dg = (int8_t)(cur.g - prev.g), db = (int8_t)(cur.b - prev.b);
    int8_t dr_dg = (int8_t)(dr - dg), db_dg = (int8_t)(db - dg);
    if (dr >= -2 && dr <= 1 && dg >= -2 && dg <= 1 && db >= -2 && db <= 1) { /* util.hpp:102-107 */
        b[0] = (uint8_t)(OP_DIFF | (dr + 2) << 4 | (dg + 2) << 2 | (db + 2));
        emit(e, b, 1);
    } else if (dr_dg >= -8 && dr_dg <= 7 && db_dg >= -8 && db_dg <= 7 && dg >= -32 && dg <= 31) { /* :109-114 */
        b[0] = (uint8_t)(OP_LUMA | (dg + 32));
        b[1] = (uint8_t)((dr_dg + 8) << 4 | (db_dg + 8));
        emit(e, b, 2);
    } else {
        b[0] = OP_RGB, b[1] = cur.r, b[2] = cur.g, b[3] = cur.b;
        emit(e, b, 4);
    }
}

static qo_pixel load_px(const uint8_t* raw, size_t i, unsigned ch) /* util.hpp:319-327 */
{
    const uint8_t* p = raw + i * ch;
    qo_pixel       v = { p[0], p[1], p[2], ch == 4 ? p[3] : (uint8_t)255 };
    return v;
}

/* ---------------------------------------------------------------- one-shot codec */

size_t qo_encode_core(const uint8_t* raw, const qo_desc* d, uint8_t* out, size_t cap, int checked, int* complete)
{ /* simple.cpp:17-98 */
    emitter  e      = { out, cap, 0, checked, 1 };
    qo_pixel tab[64];
    qo_pixel prev = START;
    unsigned run  = 0;
    memset(tab, 0, sizeof tab); /* simple.cpp:28: value-initialised, NOT seeded with START */

    emit_header(&e, d);

    /* simple.cpp:36 multiplies two u32 in 32 bits (SURVEY hazard 1); every tested size is < 2^32
     * pixels so the wrapped and the exact products agree.  We use the exact one. */
    size_t n = (size_t)d->width * d->height;
    for (size_t i = 0; i < n; ++i) {
        qo_pixel cur = load_px(raw, i, d->channels);
        if (same(cur, prev)) { /* :39-44 */
            if (++run == RUN_LIMIT) {
                emit_run(&e, run);
                run = 0;
            }
        } else {
            if (run) { /* :46-49 */
                emit_run(&e, run);
                run = 0;
            }
            unsigned s = slot_of(cur);
            if (same(tab[s], cur)) { /* :54-55 */
                uint8_t b = (uint8_t)(OP_INDEX | s);
                emit(&e, &b, 1);
            } else {
                tab[s] = cur; /* :57 -- stored before the alpha test */
                emit_colour(&e, cur, prev, d->channels == 4);
            }
        }
        prev = cur;
        if (checked && !e.ok) { /* :84-88 */
            *complete = 0;
            return e.pos;
        }
    }
    if (run) emit_run(&e, run); /* :91-94 */
    emit_marker(&e);
    *complete = e.ok;
    return e.pos;
}

void qo_decode_core(const uint8_t* in, size_t size, uint32_t width, uint32_t height, uint8_t target, uint8_t* out)
{ /* simple.cpp:100-171 */
    qo_pixel tab[64];
    qo_pixel prev = START;
    memset(tab, 0, sizeof tab);
    tab[slot_of(prev)] = prev; /* :108 */

    size_t n = (size_t)width * height, pos = HEADER, px = 0;
#define NEXT() (pos < size ? in[pos++] : (pos++, (uint8_t)0)) /* :106 zero padding past the end */
    while (px < n) {
        uint8_t  tag = NEXT();
        qo_pixel cur = prev;
        if (tag == OP_RGB) { /* :119-123 alpha is kept */
            cur.r = NEXT(), cur.g = NEXT(), cur.b = NEXT();
        } else if (tag == OP_RGBA) { /* :124-129 */
            cur.r = NEXT(), cur.g = NEXT(), cur.b = NEXT(), cur.a = NEXT();
        } else if ((tag & 0xC0) == OP_INDEX) { /* :132-135 */
            cur = tab[tag & 63];
        } else if ((tag & 0xC0) == OP_DIFF) { /* :136-144 */
            cur.r = (uint8_t)(prev.r + ((tag >> 4) & 3) - 2);
            cur.g = (uint8_t)(prev.g + ((tag >> 2) & 3) - 2);
            cur.b = (uint8_t)(prev.b + (tag & 3) - 2);
        } else if ((tag & 0xC0) == OP_LUMA) { /* :145-155 */
            uint8_t rb = NEXT();
            int     dg = (tag & 63) - 32;
            cur.r      = (uint8_t)(prev.r + dg + (rb >> 4) - 8);
            cur.g      = (uint8_t)(prev.g + dg);
            cur.b      = (uint8_t)(prev.b + dg + (rb & 15) - 8);
        } else { /* OP_RUN :156-163 -- clamped to the image, no table update */
            unsigned run = (tag & 63) + 1;
            while (run-- && px < n) {
                memcpy(out + px * target, &prev, target);
                ++px;
            }
            continue;
        }
        memcpy(out + px * target, &cur, target); /* util.hpp:281-296 */
        ++px;
        tab[slot_of(cur)] = cur; /* :169 */
        prev              = cur;
    }
#undef NEXT
}

/* ---------------------------------------------------------------- one-shot API rules */

int qo_encode_into(uint8_t* out, size_t cap, const uint8_t* raw, size_t raw_size, const qo_desc* d, size_t* written,
                   int* complete)
{ /* simple.cpp:231-252 */
    size_t need, worst;
    int    e;
    if (raw_size == 0) return QO_EMPTY;
    if ((e = qo_count_bytes(d, &need))) return e;
    if (raw_size != need) return QO_MISMATCHED_DESC;
    qo_worst_size(d, &worst);
    *written = qo_encode_core(raw, d, out, cap, cap < worst, complete);
    return QO_OK;
}

int qo_decode_into(uint8_t* out, size_t cap, const uint8_t* in, size_t size, uint8_t target, int flip, qo_desc* desc)
{ /* simple.cpp:444-494 */
    int e;
    if (size == 0) return QO_EMPTY;
    if (size <= HEADER + MARKER) return QO_TOO_SHORT;
    if ((e = qo_read_header(in, size, desc))) return e;
    size_t src_bytes;
    if ((e = qo_count_bytes(desc, &src_bytes))) return e;
    if (cap < src_bytes) return QO_NOT_ENOUGH_SPACE; /* :467-471 sized with the SOURCE channels */
    uint8_t dest = target ? target : desc->channels;
    size_t  need = (size_t)desc->width * desc->height * dest;
    if (cap < need) return QO_NOT_ENOUGH_SPACE; /* hazard 3: the reference would overflow here */
    desc->channels = dest;
    qo_decode_core(in, size, desc->width, desc->height, dest, out);
    if (flip) { /* :484-491 */
        size_t line = (size_t)desc->width * dest;
        for (size_t y = 0; y < desc->height / 2; ++y) {
            uint8_t *a = out + y * line, *b = out + (desc->height - 1 - y) * line;
            for (size_t i = 0; i < line; ++i) {
                uint8_t t = a[i];
                a[i]      = b[i];
                b[i]      = t;
            }
        }
    }
    return QO_OK;
}

/* ---------------------------------------------------------------- stream encoder */

static void state_clear(qo_state* s)
{
    memset(s, 0, sizeof *s);
    s->prev = START;
}

void qo_senc_init(qo_state* s) { state_clear(s); } /* stream.cpp:105-111 */

int qo_senc_initialize(qo_state* s, uint8_t* out, size_t cap, const qo_desc* d, size_t* written)
{ /* stream.cpp:113-136 */
    size_t n;
    int    e;
    if (s->channels) return QO_ALREADY_INITIALIZED;
    if (cap == 0) return QO_EMPTY;
    if (cap < HEADER) return QO_TOO_SHORT;
    if ((e = qo_count_bytes(d, &n))) return e;
    emitter em = { out, cap, 0, 0, 1 };
    emit_header(&em, d);
    s->channels = d->channels;
    *written    = HEADER;
    return QO_OK;
}

int qo_senc_encode(qo_state* s, uint8_t* out, size_t cap, const uint8_t* in, size_t in_size, size_t* processed,
                   size_t* written)
{ /* stream.cpp:138-239 */
    if (!s->channels) return QO_NOT_INITIALIZED;
    if (cap == 0 || in_size == 0) return QO_EMPTY;
    if (cap < 5) return QO_TOO_SHORT;

    unsigned ch = s->channels;
    size_t   n  = in_size / ch; /* :59 whole pixels only */
    emitter  e  = { out, cap, 0, 1, 1 };
    size_t   i  = 0;
    for (; i < n; ++i) {
        qo_pixel cur = load_px(in, i, ch);
        if (same(cur, s->prev)) {
            if (s->run + 1 == RUN_LIMIT) { /* :158-169: a refused RUN(62) leaves the counter at 61 */
                emit_run(&e, RUN_LIMIT);
                if (!e.ok) break;
                s->run = 0;
            } else {
                ++s->run;
            }
        } else {
            if (s->run) { /* :171-178 */
                emit_run(&e, s->run);
                if (!e.ok) break;
                s->run = 0;
            }
            unsigned slot = slot_of(cur);
            if (same(s->seen[slot], cur)) {
                uint8_t b = (uint8_t)(OP_INDEX | slot);
                emit(&e, &b, 1);
                if (!e.ok) break;
            } else {
                emit_colour(&e, cur, s->prev, ch == 4);
                if (!e.ok) break; /* :228-236: the slot is restored, i.e. never stored */
                s->seen[slot] = cur;
            }
        }
        s->prev = cur;
    }
    *processed = i * ch;
    *written   = e.pos;
    return QO_OK;
}

int qo_senc_finalize(qo_state* s, uint8_t* out, size_t cap, size_t* written)
{ /* stream.cpp:241-267 */
    if (!s->channels) return QO_NOT_INITIALIZED;
    if (cap == 0) return QO_EMPTY;
    size_t need = MARKER + (s->run > 0);
    if (cap < need) return QO_TOO_SHORT;
    emitter e = { out, cap, 0, 0, 1 };
    if (s->run) emit_run(&e, s->run);
    emit_marker(&e);
    *written = need;
    state_clear(s);
    return QO_OK;
}

void qo_senc_reset(qo_state* s) { state_clear(s); } /* stream.cpp:269-277 */

/* ---------------------------------------------------------------- stream decoder */

void qo_sdec_init(qo_state* s) { state_clear(s); } /* stream.cpp:282-288 */

int qo_sdec_initialize(qo_state* s, const uint8_t* in, size_t size, uint8_t target, qo_desc* desc)
{ /* stream.cpp:290-310 */
    size_t n;
    int    e;
    if (s->channels) return QO_ALREADY_INITIALIZED;
    if ((e = qo_read_header(in, size, desc))) return e;
    if ((e = qo_count_bytes(desc, &n))) return e;
    s->target = s->channels = target ? target : desc->channels; /* :302-304 both become the target */
    desc->channels          = s->channels;
    s->seen[slot_of(s->prev)] = s->prev; /* :306 */
    return QO_OK;
}

int qo_sdec_decode(qo_state* s, uint8_t* out, size_t cap, const uint8_t* in, size_t in_size, size_t* processed,
                   size_t* written)
{ /* stream.cpp:312-424 */
    if (!s->channels) return QO_NOT_INITIALIZED;
    if (cap == 0) return QO_EMPTY;
    if (cap < s->channels) return QO_TOO_SHORT;

    unsigned ch = s->channels;
    size_t   room = cap / ch, px = 0, pos = 0;
    while (px < room) {
        if (s->run) { /* :335-339 pending run first */
            --s->run;
            memcpy(out + px++ * ch, &s->prev, ch);
            continue;
        }
        if (pos >= in_size) break; /* :341-344 */
        uint8_t  tag = in[pos];
        size_t   len = tag == OP_RGB ? 4 : tag == OP_RGBA ? 5 : (tag & 0xC0) == OP_LUMA ? 2 : 1;
        if (pos + len > in_size) break; /* :352,364,392 incomplete op is rewound */
        const uint8_t* p   = in + pos;
        qo_pixel       cur = s->prev;
        if (tag == OP_RGB) {
            cur.r = p[1], cur.g = p[2], cur.b = p[3];
        } else if (tag == OP_RGBA) {
            cur.r = p[1], cur.g = p[2], cur.b = p[3], cur.a = p[4];
        } else if ((tag & 0xC0) == OP_INDEX) {
            cur = s->seen[tag & 63];
        } else if ((tag & 0xC0) == OP_DIFF) {
            cur.r = (uint8_t)(cur.r + ((tag >> 4) & 3) - 2);
            cur.g = (uint8_t)(cur.g + ((tag >> 2) & 3) - 2);
            cur.b = (uint8_t)(cur.b + (tag & 3) - 2);
        } else if ((tag & 0xC0) == OP_LUMA) {
            int dg = (tag & 63) - 32;
            cur.r  = (uint8_t)(cur.r + dg + (p[1] >> 4) - 8);
            cur.g  = (uint8_t)(cur.g + dg);
            cur.b  = (uint8_t)(cur.b + dg + (p[1] & 15) - 8);
        } else {
            s->run = (uint8_t)(tag & 63); /* :405-409: len-1 stays pending, one pixel is emitted now */
        }
        pos += len;
        memcpy(out + px++ * ch, &cur, ch);
        s->seen[slot_of(cur)] = cur; /* :415 (also for RUN: idempotent) */
        s->prev               = cur;
    }
    *processed = pos;
    *written   = px * ch;
    return QO_OK;
}

int qo_sdec_drain_run(qo_state* s, uint8_t* out, size_t cap, size_t* written)
{ /* stream.cpp:426-447 */
    if (!s->channels) return QO_NOT_INITIALIZED;
    if (cap == 0) return QO_EMPTY;
    size_t px = 0, ch = s->channels;
    while (s->run && (px + 1) * ch <= cap) {
        memcpy(out + px++ * ch, &s->prev, ch);
        --s->run;
    }
    *written = px * ch;
    return QO_OK;
}

void qo_sdec_reset(qo_state* s) { state_clear(s); } /* stream.cpp:449-458 */
