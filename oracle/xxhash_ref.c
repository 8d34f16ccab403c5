/*
 * ORACLE (test infrastructure only) — XXH32 / XXH64 restated from the public xxHash specification.
 * zstd's optional Content_Checksum is the low 32 bits of XXH64(content, 0) (RFC 8878 §3.1.1);
 * the LZ4 frame header checksum byte is (XXH32(descriptor, 0) >> 8) & 0xFF (LZ4 frame format 1.6).
 * On the reference path these live inside openzl-sys-rs 0.2.0 -> zstd / lz4 (codec.rs:67-78), not vendored.
 * Pinned against the python `xxhash` package in tests/test_oracle_misc.py.
 */
#include "oracle.h"
#include <string.h>

static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P64_5 0x27D4EB2F165667C5ULL

static inline uint64_t round64(uint64_t acc, uint64_t in) {
  acc += in * P64_2;
  acc = rotl64(acc, 31);
  return acc * P64_1;
}
static inline uint64_t merge64(uint64_t h, uint64_t v) {
  h ^= round64(0, v);
  return h * P64_1 + P64_4;
}

uint64_t zn_ref_xxh64(const uint8_t* p, size_t len, uint64_t seed) {
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P64_1 + P64_2, v2 = seed + P64_2, v3 = seed, v4 = seed - P64_1;
    do {
      v1 = round64(v1, rd64(p));
      v2 = round64(v2, rd64(p + 8));
      v3 = round64(v3, rd64(p + 16));
      v4 = round64(v4, rd64(p + 24));
      p += 32;
    } while (p + 32 <= end);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = merge64(h, v1);
    h = merge64(h, v2);
    h = merge64(h, v3);
    h = merge64(h, v4);
  } else {
    h = seed + P64_5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) {
    h ^= round64(0, rd64(p));
    h = rotl64(h, 27) * P64_1 + P64_4;
    p += 8;
  }
  if (p + 4 <= end) {
    h ^= (uint64_t)rd32(p) * P64_1;
    h = rotl64(h, 23) * P64_2 + P64_3;
    p += 4;
  }
  while (p < end) {
    h ^= (*p++) * P64_5;
    h = rotl64(h, 11) * P64_1;
  }
  h ^= h >> 33;
  h *= P64_2;
  h ^= h >> 29;
  h *= P64_3;
  h ^= h >> 32;
  return h;
}

#define P32_1 0x9E3779B1U
#define P32_2 0x85EBCA77U
#define P32_3 0xC2B2AE3DU
#define P32_4 0x27D4EB2FU
#define P32_5 0x165667B1U

static inline uint32_t round32(uint32_t acc, uint32_t in) {
  acc += in * P32_2;
  acc = rotl32(acc, 13);
  return acc * P32_1;
}

uint32_t zn_ref_xxh32(const uint8_t* p, size_t len, uint32_t seed) {
  const uint8_t* end = p + len;
  uint32_t h;
  if (len >= 16) {
    uint32_t v1 = seed + P32_1 + P32_2, v2 = seed + P32_2, v3 = seed, v4 = seed - P32_1;
    do {
      v1 = round32(v1, rd32(p));
      v2 = round32(v2, rd32(p + 4));
      v3 = round32(v3, rd32(p + 8));
      v4 = round32(v4, rd32(p + 12));
      p += 16;
    } while (p + 16 <= end);
    h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
  } else {
    h = seed + P32_5;
  }
  h += (uint32_t)len;
  while (p + 4 <= end) {
    h += rd32(p) * P32_3;
    h = rotl32(h, 17) * P32_4;
    p += 4;
  }
  while (p < end) {
    h += (*p++) * P32_5;
    h = rotl32(h, 11) * P32_1;
  }
  h ^= h >> 15;
  h *= P32_2;
  h ^= h >> 13;
  h *= P32_3;
  h ^= h >> 16;
  return h;
}
