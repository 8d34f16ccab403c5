/*
 * ORACLE (test infrastructure only) — Zstandard frame decoder restated from RFC 8878.
 *
 * On the reference path this arithmetic sits behind `codec::decompress_into`
 * (znippy-common/src/codec.rs:67-78: zl_get_decompressed_size -> zl_decompress), i.e. inside the
 * third-party crate openzl-sys-rs 0.2.0 -> facebook/openzl -> zstd, none of it vendored under
 * /root/reference.  The restatement follows the published format; it is pinned against libzstd 1.5.5
 * (this image) on every corpus in tests/test_oracle_zstd.py and the fixtures in tests/golden/.
 * Single-threaded and written for readability, not speed.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define ZSTD_MAGIC 0xFD2FB528u
#define BLOCK_MAX (128u * 1024u)
#define ERR(e) do { return (e); } while (0)

/* ------------------------------------------------------------------ bit readers */

/* forward little-endian bit reader (FSE table descriptions) */
typedef struct { const uint8_t* p; size_t len; size_t bitpos; } fbits;
static uint32_t fb_peek(const fbits* b, int n) {
  uint64_t v = 0;
  size_t byte = b->bitpos >> 3;
  for (int i = 0; i < 5; i++)
    if (byte + i < b->len) v |= (uint64_t)b->p[byte + i] << (8 * i);
  return (uint32_t)((v >> (b->bitpos & 7)) & ((1ull << n) - 1));
}

/* backward bit reader (Huffman + FSE payload streams): bits are consumed from the end of the buffer
 * towards its start; reading past the start yields zeros and drives `bits_left` negative. */
typedef struct { const uint8_t* p; long bits_left; } bbits;
static int bb_init(bbits* b, const uint8_t* p, size_t len) {
  if (len == 0) return -1;
  uint8_t last = p[len - 1];
  if (last == 0) return -1;
  int hb = 7;
  while (!(last >> hb)) hb--;
  b->p = p;
  b->bits_left = (long)(len - 1) * 8 + hb; /* bits below the end marker */
  return 0;
}
/* returns the next n bits (n <= 32) as they would appear MSB-first, without consuming */
static uint32_t bb_peek(const bbits* b, int n) {
  if (n == 0) return 0;
  uint64_t v = 0;
  long hi = b->bits_left; /* exclusive bit index of the first unread bit */
  for (int i = 0; i < n; i++) {
    long bit = hi - 1 - i;
    uint32_t x = 0;
    if (bit >= 0) x = (b->p[bit >> 3] >> (bit & 7)) & 1u;
    v = (v << 1) | x;
  }
  return (uint32_t)v;
}
static uint32_t bb_read(bbits* b, int n) {
  uint32_t v = bb_peek(b, n);
  b->bits_left -= n;
  return v;
}

/* ------------------------------------------------------------------ FSE */

typedef struct { uint8_t sym; uint8_t nbits; uint16_t base; } fse_entry;
typedef struct { fse_entry e[512]; int log; int valid; } fse_table;

static int highbit(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }

/* RFC 8878 §4.1.1: parse normalized counts. returns bytes consumed or <0 */
static long fse_read_ncount(const uint8_t* p, size_t len, int max_log, int max_sym, int16_t* norm,
                            int* out_log, int* out_nsym) {
  fbits b = {p, len, 0};
  if (len == 0) return -1;
  int log = (int)fb_peek(&b, 4) + 5;
  b.bitpos += 4;
  if (log > max_log) return -1;
  int remaining = (1 << log) + 1, threshold = 1 << log, nbits = log + 1, sym = 0;
  for (int i = 0; i <= max_sym; i++) norm[i] = 0;
  while (remaining > 1 && sym <= max_sym) {
    int max = (2 * threshold - 1) - remaining;
    int low = (int)fb_peek(&b, nbits - 1), value;
    if (low < max) {
      value = low;
      b.bitpos += nbits - 1;
    } else {
      value = (int)fb_peek(&b, nbits);
      if (value >= threshold) value -= max;
      b.bitpos += nbits;
    }
    int count = value - 1; /* -1 => "less than one" probability */
    remaining -= count < 0 ? -count : count;
    norm[sym++] = (int16_t)count;
    if (count == 0) {
      for (;;) {
        int rep = (int)fb_peek(&b, 2);
        b.bitpos += 2;
        sym += rep; /* that many additional zero-probability symbols */
        if (rep != 3) break;
      }
    }
    if (remaining < 1) return -1;
    while (remaining < threshold) { nbits--; threshold >>= 1; }
  }
  if (remaining != 1 || sym > max_sym + 1) return -1;
  size_t used = (b.bitpos + 7) >> 3;
  if (used > len) return -1;
  *out_log = log;
  *out_nsym = sym;
  return (long)used;
}

/* RFC 8878 §4.1.1 "from normalized distribution to decoding tables" */
static void fse_build(fse_table* t, const int16_t* norm, int nsym, int log) {
  int size = 1 << log, high = size - 1;
  uint16_t next[256];
  for (int s = 0; s < nsym; s++) {
    if (norm[s] == -1) { t->e[high--].sym = (uint8_t)s; next[s] = 1; }
    else next[s] = (uint16_t)norm[s];
  }
  int step = (size >> 1) + (size >> 3) + 3, mask = size - 1, pos = 0;
  for (int s = 0; s < nsym; s++)
    for (int i = 0; i < norm[s]; i++) {
      t->e[pos].sym = (uint8_t)s;
      do pos = (pos + step) & mask; while (pos > high);
    }
  for (int u = 0; u < size; u++) {
    uint16_t ns = next[t->e[u].sym]++;
    int nb = log - highbit(ns);
    t->e[u].nbits = (uint8_t)nb;
    t->e[u].base = (uint16_t)((ns << nb) - size);
  }
  t->log = log;
  t->valid = 1;
}
static void fse_build_rle(fse_table* t, uint8_t sym) {
  t->e[0].sym = sym; t->e[0].nbits = 0; t->e[0].base = 0; t->log = 0; t->valid = 1;
}

static const int16_t LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                       2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                       1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                       1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1,
                                       1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
static const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18,
                                     20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096,
                                     8192, 16384, 32768, 65536};
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
                                    1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20,
                                     21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41,
                                     43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195,
                                     16387, 32771, 65539};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1,
                                    2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

/* ------------------------------------------------------------------ Huffman */

typedef struct { uint8_t sym[2048]; uint8_t nbits[2048]; int max_bits; int valid; } huf_table;

/* RFC 8878 §4.2.1: returns bytes consumed by the tree description or <0 */
static long huf_read_table(huf_table* h, const uint8_t* p, size_t len, zn_ref_zstd_stats* st) {
  uint8_t w[256];
  int n = 0;
  if (len < 1) return -1;
  int hb = p[0];
  long used;
  if (hb >= 128) { /* direct 4-bit weights */
    n = hb - 127;
    size_t bytes = (size_t)(n + 1) / 2;
    if (1 + bytes > len) return -1;
    for (int i = 0; i < n; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
    used = 1 + (long)bytes;
    if (st) st->huf_weights_direct++;
  } else { /* FSE-compressed weights, two interleaved states */
    if ((size_t)hb + 1 > len || hb == 0) return -1;
    int16_t norm[16];
    int log, nsym;
    long hd = fse_read_ncount(p + 1, (size_t)hb, 6, 11, norm, &log, &nsym);
    if (hd < 0) return -1;
    fse_table t;
    fse_build(&t, norm, nsym, log);
    bbits b;
    if (bb_init(&b, p + 1 + hd, (size_t)hb - (size_t)hd) < 0) return -1;
    uint32_t s1 = bb_read(&b, log), s2 = bb_read(&b, log);
    if (b.bits_left < 0) return -1;
    for (;;) { /* at most 255 explicit weights (the 256th symbol's weight is implied) */
      if (n >= 254) return -1;
      w[n++] = t.e[s1].sym;
      s1 = t.e[s1].base + bb_read(&b, t.e[s1].nbits);
      if (b.bits_left < 0) { w[n++] = t.e[s2].sym; break; }
      if (n >= 254) return -1;
      w[n++] = t.e[s2].sym;
      s2 = t.e[s2].base + bb_read(&b, t.e[s2].nbits);
      if (b.bits_left < 0) { w[n++] = t.e[s1].sym; break; }
    }
    if (n > 255) return -1;
    used = 1 + hb;
    if (st) st->huf_weights_fse++;
  }
  /* last weight is implied so that sum 2^(w-1) is a power of two */
  uint32_t sum = 0;
  for (int i = 0; i < n; i++) {
    if (w[i] > 11) return -1;
    if (w[i]) sum += 1u << (w[i] - 1);
  }
  if (sum == 0) return -1;
  int max_bits = highbit(sum) + 1;
  if (max_bits > 11) return -1;
  uint32_t left = (1u << max_bits) - sum;
  if (left & (left - 1)) return -1; /* must be a power of two */
  w[n++] = (uint8_t)(highbit(left) + 1);
  /* canonical assignment: ascending weight (longest codes first), ascending symbol within a weight */
  uint32_t rank_start[13] = {0}, count[13] = {0};
  for (int i = 0; i < n; i++) count[w[i]]++;
  uint32_t pos = 0;
  for (int wt = 1; wt <= max_bits; wt++) { rank_start[wt] = pos; pos += count[wt] << (wt - 1); }
  if (pos != (1u << max_bits)) return -1;
  for (int s = 0; s < n; s++) {
    if (!w[s]) continue;
    uint32_t span = 1u << (w[s] - 1), start = rank_start[w[s]];
    for (uint32_t k = 0; k < span; k++) {
      h->sym[start + k] = (uint8_t)s;
      h->nbits[start + k] = (uint8_t)(max_bits + 1 - w[s]);
    }
    rank_start[w[s]] += span;
  }
  h->max_bits = max_bits;
  h->valid = 1;
  return used;
}

static int huf_decode_stream(const huf_table* h, const uint8_t* p, size_t len, uint8_t* out, size_t n) {
  bbits b;
  if (bb_init(&b, p, len) < 0) return -1;
  for (size_t i = 0; i < n; i++) {
    uint32_t idx = bb_peek(&b, h->max_bits);
    out[i] = h->sym[idx];
    b.bits_left -= h->nbits[idx];
  }
  return b.bits_left == 0 ? 0 : -1; /* stream must be consumed exactly */
}

/* ------------------------------------------------------------------ frame state */

typedef struct {
  huf_table huf;
  fse_table ll, of, ml;
  uint32_t rep[3];
  uint8_t* lit; /* BLOCK_MAX scratch */
} frame_ctx;

static int decode_literals(frame_ctx* fc, const uint8_t* p, size_t len, size_t* consumed,
                           const uint8_t** lit_out, size_t* lit_len, zn_ref_zstd_stats* st) {
  if (len < 1) ERR(ZN_REF_ERR_CORRUPT);
  int type = p[0] & 3, sf = (p[0] >> 2) & 3;
  if (type < 2) { /* raw / RLE */
    size_t hdr, regen;
    if ((sf & 1) == 0) { hdr = 1; regen = p[0] >> 3; }
    else if (sf == 1) { if (len < 2) ERR(ZN_REF_ERR_CORRUPT); hdr = 2; regen = (p[0] >> 4) | ((size_t)p[1] << 4); }
    else { if (len < 3) ERR(ZN_REF_ERR_CORRUPT); hdr = 3; regen = (p[0] >> 4) | ((size_t)p[1] << 4) | ((size_t)p[2] << 12); }
    if (regen > BLOCK_MAX) ERR(ZN_REF_ERR_CORRUPT);
    if (type == 0) {
      if (hdr + regen > len) ERR(ZN_REF_ERR_CORRUPT);
      *lit_out = p + hdr; *lit_len = regen; *consumed = hdr + regen;
      if (st) st->lit_raw++;
    } else {
      if (hdr + 1 > len) ERR(ZN_REF_ERR_CORRUPT);
      memset(fc->lit, p[hdr], regen);
      *lit_out = fc->lit; *lit_len = regen; *consumed = hdr + 1;
      if (st) st->lit_rle++;
    }
    return 0;
  }
  /* Huffman-compressed (2) or treeless (3) */
  size_t hdr, regen, comp;
  int streams;
  if (sf <= 1) {
    if (len < 3) ERR(ZN_REF_ERR_CORRUPT);
    uint32_t v = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    hdr = 3; regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; streams = sf == 0 ? 1 : 4;
  } else if (sf == 2) {
    if (len < 4) ERR(ZN_REF_ERR_CORRUPT);
    uint32_t v = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    hdr = 4; regen = (v >> 4) & 0x3FFF; comp = v >> 18; streams = 4;
  } else {
    if (len < 5) ERR(ZN_REF_ERR_CORRUPT);
    uint64_t v = p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24) | ((uint64_t)p[4] << 32);
    hdr = 5; regen = (v >> 4) & 0x3FFFF; comp = (size_t)(v >> 22); streams = 4;
  }
  if (regen > BLOCK_MAX || hdr + comp > len) ERR(ZN_REF_ERR_CORRUPT);
  const uint8_t* q = p + hdr;
  size_t qlen = comp;
  if (type == 2) {
    long used = huf_read_table(&fc->huf, q, qlen, st);
    if (used < 0) ERR(ZN_REF_ERR_CORRUPT);
    q += used; qlen -= (size_t)used;
  } else {
    if (!fc->huf.valid) ERR(ZN_REF_ERR_CORRUPT);
    if (st) st->lit_treeless++;
  }
  if (streams == 1) {
    if (huf_decode_stream(&fc->huf, q, qlen, fc->lit, regen) < 0) ERR(ZN_REF_ERR_CORRUPT);
    if (st) st->lit_huf_1stream++;
  } else {
    if (qlen < 6) ERR(ZN_REF_ERR_CORRUPT);
    size_t s1 = q[0] | ((size_t)q[1] << 8), s2 = q[2] | ((size_t)q[3] << 8), s3 = q[4] | ((size_t)q[5] << 8);
    if (6 + s1 + s2 + s3 > qlen) ERR(ZN_REF_ERR_CORRUPT);
    size_t s4 = qlen - 6 - s1 - s2 - s3, seg = (regen + 3) / 4;
    if (seg * 3 > regen) ERR(ZN_REF_ERR_CORRUPT);
    const uint8_t* d = q + 6;
    if (huf_decode_stream(&fc->huf, d, s1, fc->lit, seg) < 0) ERR(ZN_REF_ERR_CORRUPT);
    if (huf_decode_stream(&fc->huf, d + s1, s2, fc->lit + seg, seg) < 0) ERR(ZN_REF_ERR_CORRUPT);
    if (huf_decode_stream(&fc->huf, d + s1 + s2, s3, fc->lit + 2 * seg, seg) < 0) ERR(ZN_REF_ERR_CORRUPT);
    if (huf_decode_stream(&fc->huf, d + s1 + s2 + s3, s4, fc->lit + 3 * seg, regen - 3 * seg) < 0) ERR(ZN_REF_ERR_CORRUPT);
    if (st) st->lit_huf_4stream++;
  }
  *lit_out = fc->lit; *lit_len = regen; *consumed = hdr + comp;
  return 0;
}

static int setup_seq_table(fse_table* t, int mode, const uint8_t** pp, const uint8_t* end, int max_log,
                           int max_sym, const int16_t* def, int def_n, int def_log, zn_ref_zstd_stats* st) {
  if (mode == 0) { fse_build(t, def, def_n, def_log); if (st) st->mode_predefined++; return 0; }
  if (mode == 1) {
    if (*pp >= end) return -1;
    if (**pp > max_sym) return -1;
    fse_build_rle(t, *(*pp)++);
    if (st) st->mode_rle++;
    return 0;
  }
  if (mode == 2) {
    int16_t norm[64];
    int log, nsym;
    long used = fse_read_ncount(*pp, (size_t)(end - *pp), max_log, max_sym, norm, &log, &nsym);
    if (used < 0) return -1;
    fse_build(t, norm, nsym, log);
    *pp += used;
    if (st) st->mode_fse++;
    return 0;
  }
  if (!t->valid) return -1; /* repeat with no previous table */
  if (st) st->mode_repeat++;
  return 0;
}

static int decode_block(frame_ctx* fc, const uint8_t* p, size_t len, uint8_t* dst_base, size_t dst_cap,
                        size_t* pos_io, size_t frame_start, zn_ref_zstd_stats* st) {
  const uint8_t* lit;
  size_t lit_len, used;
  int rc = decode_literals(fc, p, len, &used, &lit, &lit_len, st);
  if (rc) return rc;
  const uint8_t* q = p + used;
  const uint8_t* end = p + len;
  size_t pos = *pos_io, block_start = pos;
  if (q >= end) ERR(ZN_REF_ERR_CORRUPT);
  size_t nseq = *q++;
  if (nseq >= 128) {
    if (nseq == 255) { if (end - q < 2) ERR(ZN_REF_ERR_CORRUPT); nseq = q[0] + ((size_t)q[1] << 8) + 0x7F00; q += 2; }
    else { if (end - q < 1) ERR(ZN_REF_ERR_CORRUPT); nseq = ((nseq - 128) << 8) + q[0]; q += 1; }
  }
  size_t lit_pos = 0;
  if (nseq > 0) {
    if (q >= end) ERR(ZN_REF_ERR_CORRUPT);
    int modes = *q++;
    if (modes & 3) ERR(ZN_REF_ERR_CORRUPT);
    if (setup_seq_table(&fc->ll, modes >> 6, &q, end, 9, 35, LL_DEFAULT, 36, 6, st)) ERR(ZN_REF_ERR_CORRUPT);
    if (setup_seq_table(&fc->of, (modes >> 4) & 3, &q, end, 8, 31, OF_DEFAULT, 29, 5, st)) ERR(ZN_REF_ERR_CORRUPT);
    if (setup_seq_table(&fc->ml, (modes >> 2) & 3, &q, end, 9, 52, ML_DEFAULT, 53, 6, st)) ERR(ZN_REF_ERR_CORRUPT);
    bbits b;
    if (bb_init(&b, q, (size_t)(end - q)) < 0) ERR(ZN_REF_ERR_CORRUPT);
    uint32_t sl = bb_read(&b, fc->ll.log), so = bb_read(&b, fc->of.log), sm = bb_read(&b, fc->ml.log);
    if (b.bits_left < 0) ERR(ZN_REF_ERR_CORRUPT);
    for (size_t i = 0; i < nseq; i++) {
      int oc = fc->of.e[so].sym, mc = fc->ml.e[sm].sym, lc = fc->ll.e[sl].sym;
      if (oc > 31 || mc > 52 || lc > 35) ERR(ZN_REF_ERR_CORRUPT);
      uint64_t ov = ((uint64_t)1 << oc) + bb_read(&b, oc);
      uint32_t ml = ML_BASE[mc] + bb_read(&b, ML_BITS[mc]);
      uint32_t ll = LL_BASE[lc] + bb_read(&b, LL_BITS[lc]);
      if (b.bits_left < 0) ERR(ZN_REF_ERR_CORRUPT);
      if (i + 1 < nseq) {
        sl = fc->ll.e[sl].base + bb_read(&b, fc->ll.e[sl].nbits);
        sm = fc->ml.e[sm].base + bb_read(&b, fc->ml.e[sm].nbits);
        so = fc->of.e[so].base + bb_read(&b, fc->of.e[so].nbits);
        if (b.bits_left < 0) ERR(ZN_REF_ERR_CORRUPT);
      }
      /* repeat-offset resolution, RFC 8878 §3.1.1.5 */
      uint64_t offset;
      if (ov > 3) {
        offset = ov - 3;
        fc->rep[2] = fc->rep[1]; fc->rep[1] = fc->rep[0]; fc->rep[0] = (uint32_t)offset;
      } else {
        uint32_t idx = (uint32_t)ov - 1 + (ll == 0 ? 1 : 0); /* 0..3 */
        if (st) st->repcode_uses++;
        if (idx == 0) offset = fc->rep[0];
        else {
          offset = idx == 3 ? (uint64_t)fc->rep[0] - 1 : fc->rep[idx];
          if (offset == 0) ERR(ZN_REF_ERR_CORRUPT);
          if (idx != 1) fc->rep[2] = fc->rep[1];
          fc->rep[1] = fc->rep[0];
          fc->rep[0] = (uint32_t)offset;
        }
      }
      /* execute */
      if (ll > lit_len - lit_pos) ERR(ZN_REF_ERR_CORRUPT);
      if ((size_t)ll + ml > dst_cap - pos) ERR(ZN_REF_ERR_DST_TOO_SMALL);
      memcpy(dst_base + pos, lit + lit_pos, ll);
      pos += ll; lit_pos += ll;
      if (offset > pos - frame_start) ERR(ZN_REF_ERR_CORRUPT);
      if (st) { st->sequences++; st->match_bytes += ml; if (offset < ml) st->overlap_matches++; }
      const uint8_t* s = dst_base + pos - offset;
      uint8_t* d = dst_base + pos;
      for (uint32_t k = 0; k < ml; k++) d[k] = s[k]; /* byte order makes overlapping matches replicate */
      pos += ml;
    }
    if (b.bits_left != 0) ERR(ZN_REF_ERR_CORRUPT);
  }
  size_t rest = lit_len - lit_pos;
  if (rest > dst_cap - pos) ERR(ZN_REF_ERR_DST_TOO_SMALL);
  memcpy(dst_base + pos, lit + lit_pos, rest);
  pos += rest;
  if (st) st->literal_bytes += lit_len;
  if (pos - block_start > BLOCK_MAX) ERR(ZN_REF_ERR_CORRUPT);
  *pos_io = pos;
  return 0;
}

/* parses a frame header; returns header size or <0 */
static long parse_frame_header(const uint8_t* p, size_t len, uint64_t* fcs, int* has_fcs, uint64_t* window,
                               int* checksum) {
  if (len < 5) return ZN_REF_ERR_SRC_TRUNCATED;
  uint32_t magic = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  if (magic != ZSTD_MAGIC) return ZN_REF_ERR_BAD_MAGIC;
  int fhd = p[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
  if (fhd & 0x08) return ZN_REF_ERR_UNSUPPORTED; /* reserved bit */
  *checksum = (fhd >> 2) & 1;
  size_t pos = 5;
  uint64_t win = 0;
  if (!single) {
    if (len < pos + 1) return ZN_REF_ERR_SRC_TRUNCATED;
    int wd = p[pos++], e = wd >> 3, m = wd & 7;
    uint64_t base = 1ull << (10 + e);
    win = base + (base >> 3) * m;
  }
  static const int did_bytes[4] = {0, 1, 2, 4};
  int db = did_bytes[did_flag];
  if (len < pos + db) return ZN_REF_ERR_SRC_TRUNCATED;
  uint32_t did = 0;
  for (int i = 0; i < db; i++) did |= (uint32_t)p[pos + i] << (8 * i);
  pos += db;
  if (did != 0) return ZN_REF_ERR_UNSUPPORTED;
  int fb = fcs_flag == 0 ? (single ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
  if (len < pos + fb) return ZN_REF_ERR_SRC_TRUNCATED;
  uint64_t v = 0;
  for (int i = 0; i < fb; i++) v |= (uint64_t)p[pos + i] << (8 * i);
  if (fb == 2) v += 256;
  pos += fb;
  *has_fcs = fb != 0;
  *fcs = v;
  if (single) win = v;
  *window = win;
  return (long)pos;
}

int zn_ref_zstd_frame_content_size(const uint8_t* src, size_t src_len, uint64_t* fcs) {
  uint64_t win;
  int has, cs;
  long h = parse_frame_header(src, src_len, fcs, &has, &win, &cs);
  if (h < 0) return (int)h;
  return has ? 0 : 1;
}

int zn_ref_zstd_decompress(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap,
                           size_t* out_len, zn_ref_zstd_stats* st) {
  size_t ip = 0, pos = 0;
  if (st) memset(st, 0, sizeof *st);
  *out_len = 0;
  if (src_len == 0) ERR(ZN_REF_ERR_SRC_TRUNCATED);
  frame_ctx* fc = (frame_ctx*)malloc(sizeof *fc);
  fc->lit = (uint8_t*)malloc(BLOCK_MAX + 64);
  int rc = 0;
  while (ip < src_len && rc == 0) {
    if (src_len - ip >= 8) { /* skippable frame */
      uint32_t magic = src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
      if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
        uint32_t sz = src[ip + 4] | ((uint32_t)src[ip + 5] << 8) | ((uint32_t)src[ip + 6] << 16) | ((uint32_t)src[ip + 7] << 24);
        if ((size_t)sz > src_len - ip - 8) { rc = ZN_REF_ERR_SRC_TRUNCATED; break; }
        ip += 8 + sz;
        if (st) st->skippable_frames++;
        continue;
      }
    }
    uint64_t fcs, window;
    int has_fcs, checksum;
    long h = parse_frame_header(src + ip, src_len - ip, &fcs, &has_fcs, &window, &checksum);
    if (h < 0) { rc = (int)h; break; }
    ip += (size_t)h;
    if (st) st->frames++;
    size_t frame_start = pos;
    size_t block_max = window < BLOCK_MAX ? (size_t)window : BLOCK_MAX;
    fc->huf.valid = fc->ll.valid = fc->of.valid = fc->ml.valid = 0;
    fc->rep[0] = 1; fc->rep[1] = 4; fc->rep[2] = 8;
    for (;;) {
      if (src_len - ip < 3) { rc = ZN_REF_ERR_SRC_TRUNCATED; break; }
      uint32_t bh = src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
      ip += 3;
      int last = bh & 1, type = (bh >> 1) & 3;
      size_t bsize = bh >> 3;
      if (type == 3) { rc = ZN_REF_ERR_CORRUPT; break; }
      if (type == 0) {
        if (bsize > src_len - ip) { rc = ZN_REF_ERR_SRC_TRUNCATED; break; }
        if (bsize > dst_cap - pos) { rc = ZN_REF_ERR_DST_TOO_SMALL; break; }
        memcpy(dst + pos, src + ip, bsize);
        pos += bsize; ip += bsize;
        if (st) st->blocks_raw++;
      } else if (type == 1) {
        if (src_len - ip < 1) { rc = ZN_REF_ERR_SRC_TRUNCATED; break; }
        if (bsize > dst_cap - pos) { rc = ZN_REF_ERR_DST_TOO_SMALL; break; }
        memset(dst + pos, src[ip], bsize);
        pos += bsize; ip += 1;
        if (st) st->blocks_rle++;
      } else {
        if (bsize > src_len - ip) { rc = ZN_REF_ERR_SRC_TRUNCATED; break; }
        if (bsize > block_max || bsize < 2) { rc = ZN_REF_ERR_CORRUPT; break; }
        rc = decode_block(fc, src + ip, bsize, dst, dst_cap, &pos, frame_start, st);
        if (rc) break;
        ip += bsize;
        if (st) st->blocks_compressed++;
      }
      if (last) break;
    }
    if (rc) break;
    if (has_fcs && pos - frame_start != fcs) { rc = ZN_REF_ERR_SIZE_MISMATCH; break; }
    if (checksum) {
      if (src_len - ip < 4) { rc = ZN_REF_ERR_SRC_TRUNCATED; break; }
      uint32_t want = src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16) | ((uint32_t)src[ip + 3] << 24);
      uint32_t got = (uint32_t)zn_ref_xxh64(dst + frame_start, pos - frame_start, 0);
      ip += 4;
      if (want != got) { rc = ZN_REF_ERR_CHECKSUM; break; }
      if (st) st->checksums_verified++;
    }
  }
  free(fc->lit);
  free(fc);
  *out_len = pos;
  return rc;
}
