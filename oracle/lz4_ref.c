/*
 * ORACLE (test infrastructure only) — LZ4 block and LZ4 frame decoders restated from the public
 * LZ4 block format / frame format (v1.6.x) descriptions.  On the reference path LZ4 is one of the
 * two payload codecs OpenZL wraps (README.md:10, codec.rs:1-2) via openzl-sys-rs 0.2.0, not vendored.
 * Pinned against liblz4 1.9.4 (LZ4_compress_default / LZ4_compress_HC / LZ4F_compressFrame outputs)
 * in tests/test_oracle_lz4.py.
 */
#include "oracle.h"
#include <string.h>

long zn_ref_lz4_block_decompress(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap) {
  size_t ip = 0, op = 0;
  if (src_len == 0) return ZN_REF_ERR_SRC_TRUNCATED;
  for (;;) {
    if (ip >= src_len) return ZN_REF_ERR_SRC_TRUNCATED;
    unsigned token = src[ip++];
    size_t ll = token >> 4;
    if (ll == 15) {
      unsigned b;
      do {
        if (ip >= src_len) return ZN_REF_ERR_SRC_TRUNCATED;
        b = src[ip++];
        ll += b;
      } while (b == 255);
    }
    if (ll > src_len - ip) return ZN_REF_ERR_SRC_TRUNCATED;
    if (ll > dst_cap - op) return ZN_REF_ERR_DST_TOO_SMALL;
    memcpy(dst + op, src + ip, ll);
    ip += ll;
    op += ll;
    if (ip == src_len) break; /* last sequence: literals only */
    if (src_len - ip < 2) return ZN_REF_ERR_SRC_TRUNCATED;
    size_t offset = src[ip] | ((size_t)src[ip + 1] << 8);
    ip += 2;
    if (offset == 0 || offset > op) return ZN_REF_ERR_CORRUPT;
    size_t ml = token & 15;
    if (ml == 15) {
      unsigned b;
      do {
        if (ip >= src_len) return ZN_REF_ERR_SRC_TRUNCATED;
        b = src[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (ml > dst_cap - op) return ZN_REF_ERR_DST_TOO_SMALL;
    for (size_t k = 0; k < ml; k++) dst[op + k] = dst[op - offset + k];
    op += ml;
  }
  return (long)op;
}

#define LZ4F_MAGIC 0x184D2204u
static uint32_t rd32(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* returns header size or <0; fills flags */
static long lz4f_header(const uint8_t* src, size_t len, uint64_t* csize, int* has_csize, int* bchk, int* cchk,
                        int* indep, size_t* bmax) {
  if (len < 7) return ZN_REF_ERR_SRC_TRUNCATED;
  if (rd32(src) != LZ4F_MAGIC) return ZN_REF_ERR_BAD_MAGIC;
  int flg = src[4], bd = src[5];
  if ((flg >> 6) != 1) return ZN_REF_ERR_UNSUPPORTED; /* version */
  if (flg & 0x02) return ZN_REF_ERR_UNSUPPORTED;       /* reserved */
  *indep = (flg >> 5) & 1;
  *bchk = (flg >> 4) & 1;
  *has_csize = (flg >> 3) & 1;
  *cchk = (flg >> 2) & 1;
  int dict = flg & 1;
  if (bd & 0x8F) return ZN_REF_ERR_UNSUPPORTED;
  int bs = (bd >> 4) & 7;
  if (bs < 4) return ZN_REF_ERR_UNSUPPORTED;
  *bmax = (size_t)1 << (8 + 2 * bs); /* 4:64K 5:256K 6:1M 7:4M */
  size_t pos = 6;
  *csize = 0;
  if (*has_csize) {
    if (len < pos + 8) return ZN_REF_ERR_SRC_TRUNCATED;
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)src[pos + i] << (8 * i);
    *csize = v;
    pos += 8;
  }
  if (dict) return ZN_REF_ERR_UNSUPPORTED;
  if (len < pos + 1) return ZN_REF_ERR_SRC_TRUNCATED;
  uint8_t hc = (uint8_t)(zn_ref_xxh32(src + 4, pos - 4, 0) >> 8);
  if (hc != src[pos]) return ZN_REF_ERR_CHECKSUM;
  return (long)(pos + 1);
}

int zn_ref_lz4_frame_content_size(const uint8_t* src, size_t src_len, uint64_t* fcs) {
  int has, b, c, ind;
  size_t bmax;
  long h = lz4f_header(src, src_len, fcs, &has, &b, &c, &ind, &bmax);
  if (h < 0) return (int)h;
  return has ? 0 : 1;
}

int zn_ref_lz4_frame_decompress(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_cap,
                                size_t* out_len) {
  uint64_t csize;
  int has_csize, bchk, cchk, indep;
  size_t bmax;
  *out_len = 0;
  long h = lz4f_header(src, src_len, &csize, &has_csize, &bchk, &cchk, &indep, &bmax);
  if (h < 0) return (int)h;
  if (!indep) return ZN_REF_ERR_UNSUPPORTED; /* linked blocks are not produced on this path */
  size_t ip = (size_t)h, op = 0;
  for (;;) {
    if (src_len - ip < 4) return ZN_REF_ERR_SRC_TRUNCATED;
    uint32_t bs = rd32(src + ip);
    ip += 4;
    if (bs == 0) break; /* EndMark */
    int raw = bs >> 31;
    size_t n = bs & 0x7FFFFFFFu;
    if (n > bmax || n > src_len - ip) return ZN_REF_ERR_SRC_TRUNCATED;
    if (raw) {
      if (n > dst_cap - op) return ZN_REF_ERR_DST_TOO_SMALL;
      memcpy(dst + op, src + ip, n);
      op += n;
    } else {
      size_t cap = dst_cap - op < bmax ? dst_cap - op : bmax;
      long w = zn_ref_lz4_block_decompress(src + ip, n, dst + op, cap);
      if (w < 0) return (int)w;
      op += (size_t)w;
    }
    ip += n;
    if (bchk) {
      if (src_len - ip < 4) return ZN_REF_ERR_SRC_TRUNCATED;
      if (rd32(src + ip) != zn_ref_xxh32(src + ip - n, n, 0)) return ZN_REF_ERR_CHECKSUM;
      ip += 4;
    }
  }
  if (cchk) {
    if (src_len - ip < 4) return ZN_REF_ERR_SRC_TRUNCATED;
    if (rd32(src + ip) != zn_ref_xxh32(dst, op, 0)) return ZN_REF_ERR_CHECKSUM;
    ip += 4;
  }
  if (has_csize && csize != op) return ZN_REF_ERR_SIZE_MISMATCH;
  *out_len = op;
  return ZN_REF_OK;
}
