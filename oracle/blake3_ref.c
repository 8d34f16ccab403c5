/*
 * ORACLE (test infrastructure only) — BLAKE3 default-mode hash, restated from the public
 * BLAKE3 specification.  Replaces, for checking purposes, `blake3::hash(&[u8])` as called by
 * the reference at znippy-common/src/decompress.rs:172, znippy-compress/src/stream_packer.rs:219
 * and znippy-compress/src/slot_packer.rs:553 (crate blake3 1.8.5, not vendored in /root/reference).
 * Pinned by the KAT table in SURVEY.md §8(c) and the official Python bindings (tests/test_oracle_blake3.py).
 */
#include "oracle.h"
#include <string.h>

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

#define G(a, b, c, d, mx, my)      \
  do {                             \
    v[a] = v[a] + v[b] + (mx);     \
    v[d] = rotr(v[d] ^ v[a], 16);  \
    v[c] = v[c] + v[d];            \
    v[b] = rotr(v[b] ^ v[c], 12);  \
    v[a] = v[a] + v[b] + (my);     \
    v[d] = rotr(v[d] ^ v[a], 8);   \
    v[c] = v[c] + v[d];            \
    v[b] = rotr(v[b] ^ v[c], 7);   \
  } while (0)

/* one compression; writes the 8-word chaining value */
static void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                     uint32_t block_len, uint32_t flags, uint32_t out[8]) {
  uint32_t v[16], m[16], t[16];
  for (int i = 0; i < 8; i++) v[i] = cv[i];
  for (int i = 0; i < 4; i++) v[8 + i] = IV[i];
  v[12] = (uint32_t)counter;
  v[13] = (uint32_t)(counter >> 32);
  v[14] = block_len;
  v[15] = flags;
  memcpy(m, block, 64);
  for (int r = 0; r < 7; r++) {
    G(0, 4, 8, 12, m[0], m[1]);
    G(1, 5, 9, 13, m[2], m[3]);
    G(2, 6, 10, 14, m[4], m[5]);
    G(3, 7, 11, 15, m[6], m[7]);
    G(0, 5, 10, 15, m[8], m[9]);
    G(1, 6, 11, 12, m[10], m[11]);
    G(2, 7, 8, 13, m[12], m[13]);
    G(3, 4, 9, 14, m[14], m[15]);
    for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
    memcpy(m, t, 64);
  }
  for (int i = 0; i < 8; i++) out[i] = v[i] ^ v[i + 8];
}

static void load_block(const uint8_t* p, size_t n, uint32_t w[16]) {
  uint8_t buf[64];
  memset(buf, 0, 64);
  memcpy(buf, p, n);
  for (int i = 0; i < 16; i++)
    w[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) |
           ((uint32_t)buf[4 * i + 3] << 24);
}

/* chaining value of one chunk (<= 1024 bytes); `root` sets ROOT on the final block */
static void chunk_cv(const uint8_t* p, size_t len, uint64_t chunk_index, int root, uint32_t out[8]) {
  uint32_t cv[8], w[16];
  memcpy(cv, IV, 32);
  size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
  for (size_t b = 0; b < nblocks; b++) {
    size_t off = b * 64, n = len - off < 64 ? len - off : 64;
    uint32_t flags = 0;
    if (b == 0) flags |= CHUNK_START;
    if (b == nblocks - 1) flags |= CHUNK_END | (root ? ROOT : 0);
    load_block(p + off, n, w);
    compress(cv, w, chunk_index, (uint32_t)n, flags, cv);
  }
  memcpy(out, cv, 32);
}

static void parent_cv(const uint32_t l[8], const uint32_t r[8], int root, uint32_t out[8]) {
  uint32_t w[16];
  memcpy(w, l, 32);
  memcpy(w + 8, r, 32);
  compress(IV, w, 0, 64, PARENT | (root ? ROOT : 0), out);
}

static void store_digest(const uint32_t cv[8], uint8_t out[32]) {
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)cv[i];
    out[4 * i + 1] = (uint8_t)(cv[i] >> 8);
    out[4 * i + 2] = (uint8_t)(cv[i] >> 16);
    out[4 * i + 3] = (uint8_t)(cv[i] >> 24);
  }
}

/* Merge a level array of chaining values by adjacent pairing with odd carry-up.  This is the same
 * tree as the spec's "left subtree = largest power of two chunks strictly less than the total". */
static void reduce_tree(uint32_t (*cvs)[8], size_t n, uint8_t out[32]) {
  while (n > 1) {
    size_t m = 0;
    for (size_t i = 0; i + 1 < n; i += 2) {
      uint32_t t[8];
      parent_cv(cvs[i], cvs[i + 1], n == 2, t);
      memcpy(cvs[m++], t, 32);
    }
    if (n & 1) memcpy(cvs[m++], cvs[n - 1], 32);
    n = m;
  }
  store_digest(cvs[0], out);
}

#include <stdlib.h>

void zn_ref_blake3(const uint8_t* data, size_t len, uint8_t out[32]) {
  size_t nchunks = len == 0 ? 1 : (len + 1023) / 1024;
  if (nchunks == 1) {
    uint32_t cv[8];
    chunk_cv(data, len, 0, 1, cv);
    store_digest(cv, out);
    return;
  }
  uint32_t(*cvs)[8] = (uint32_t(*)[8])malloc(nchunks * 32);
  for (size_t c = 0; c < nchunks; c++) {
    size_t off = c * 1024, n = len - off < 1024 ? len - off : 1024;
    chunk_cv(data + off, n, c, 0, cvs[c]);
  }
  reduce_tree(cvs, nchunks, out);
  free(cvs);
}

/* ------------------------------------------------------------------------------------------
 * SIMD-across-chunks variant (W chunks per vector lane set) using GCC vector extensions, so the
 * CPU baseline is not handicapped relative to the reference's AVX2/AVX-512 blake3 crate.
 * ------------------------------------------------------------------------------------------ */
#define VW 16
typedef uint32_t vec __attribute__((vector_size(VW * 4)));

static inline vec vrotr(vec x, int n) { return (x >> n) | (x << (32 - n)); }

#define VG(a, b, c, d, mx, my)     \
  do {                             \
    v[a] = v[a] + v[b] + (mx);     \
    v[d] = vrotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d];            \
    v[b] = vrotr(v[b] ^ v[c], 12); \
    v[a] = v[a] + v[b] + (my);     \
    v[d] = vrotr(v[d] ^ v[a], 8);  \
    v[c] = v[c] + v[d];            \
    v[b] = vrotr(v[b] ^ v[c], 7);  \
  } while (0)

static const uint8_t SCHED[7][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
    {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1},
    {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
    {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4},
    {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
    {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};

/* hashes VW full 1024-byte chunks starting at p (consecutive), chunk indices c0..c0+VW-1 */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void chunks_cv_wide(const uint8_t* p, uint64_t c0, uint32_t (*out)[8]) {
  vec cv[8];
  for (int i = 0; i < 8; i++)
    for (int l = 0; l < VW; l++) cv[i][l] = IV[i];
  for (int b = 0; b < 16; b++) {
    vec m[16], v[16];
    for (int w = 0; w < 16; w++)
      for (int l = 0; l < VW; l++) {
        uint32_t x;
        memcpy(&x, p + (size_t)l * 1024 + b * 64 + w * 4, 4);
        m[w][l] = x;
      }
    for (int i = 0; i < 8; i++) v[i] = cv[i];
    for (int i = 0; i < 4; i++)
      for (int l = 0; l < VW; l++) v[8 + i][l] = IV[i];
    for (int l = 0; l < VW; l++) {
      v[12][l] = (uint32_t)(c0 + l);
      v[13][l] = (uint32_t)((c0 + l) >> 32);
      v[14][l] = 64;
      v[15][l] = (b == 0 ? CHUNK_START : 0) | (b == 15 ? CHUNK_END : 0);
    }
    for (int r = 0; r < 7; r++) {
      const uint8_t* s = SCHED[r];
      VG(0, 4, 8, 12, m[s[0]], m[s[1]]);
      VG(1, 5, 9, 13, m[s[2]], m[s[3]]);
      VG(2, 6, 10, 14, m[s[4]], m[s[5]]);
      VG(3, 7, 11, 15, m[s[6]], m[s[7]]);
      VG(0, 5, 10, 15, m[s[8]], m[s[9]]);
      VG(1, 6, 11, 12, m[s[10]], m[s[11]]);
      VG(2, 7, 8, 13, m[s[12]], m[s[13]]);
      VG(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; i++) cv[i] = v[i] ^ v[i + 8];
  }
  for (int l = 0; l < VW; l++)
    for (int i = 0; i < 8; i++) out[l][i] = cv[i][l];
}

void zn_ref_blake3_fast(const uint8_t* data, size_t len, uint8_t out[32]) {
  size_t nchunks = len == 0 ? 1 : (len + 1023) / 1024;
  if (nchunks < 2 * VW) {
    zn_ref_blake3(data, len, out);
    return;
  }
  uint32_t(*cvs)[8] = (uint32_t(*)[8])malloc(nchunks * 32);
  size_t full = len / 1024, c = 0;
  for (; c + VW <= full; c += VW) chunks_cv_wide(data + c * 1024, c, cvs + c);
  for (; c < nchunks; c++) {
    size_t off = c * 1024, n = len - off < 1024 ? len - off : 1024;
    chunk_cv(data + off, n, c, 0, cvs[c]);
  }
  reduce_tree(cvs, nchunks, out);
  free(cvs);
}
