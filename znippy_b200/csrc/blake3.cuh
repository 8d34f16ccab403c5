// BLAKE3 (default hash mode) device primitives for sm_100a.
//
// Replaces blake3::hash(&[u8]) at the reference's three call sites (znippy-common/src/decompress.rs:172,
// znippy-compress/src/stream_packer.rs:219, znippy-compress/src/slot_packer.rs:553).
//
// Mapping: one lane hashes one 1 KiB chunk (16 chained 64-byte compressions); the chunk chaining values are
// then merged level by level (adjacent pairs, odd one carried up), which is the spec's left-full tree.
// The compression function is written so that ptxas sees: 16/8-bit rotates as PRMT, 12/7-bit rotates as
// SHF.R.W, the 3-input adds as IADD3, and the message schedule as pure register renaming (every index below is
// a compile-time constant).
#pragma once
#include "common.cuh"

namespace zn {
namespace b3 {

enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

#define ZN_IV0 0x6A09E667u
#define ZN_IV1 0xBB67AE85u
#define ZN_IV2 0x3C6EF372u
#define ZN_IV3 0xA54FF53Au
#define ZN_IV4 0x510E527Fu
#define ZN_IV5 0x9B05688Cu
#define ZN_IV6 0x1F83D9ABu
#define ZN_IV7 0x5BE0CD19u

ZN_D uint32_t rotr16(uint32_t x) { return __byte_perm(x, x, 0x1032); }
ZN_D uint32_t rotr8(uint32_t x) { return __byte_perm(x, x, 0x0321); }
ZN_D uint32_t rotr12(uint32_t x) { return __funnelshift_r(x, x, 12); }
ZN_D uint32_t rotr7(uint32_t x) { return __funnelshift_r(x, x, 7); }

// Pipe balance (measured with tools/ubench_b3.cu on B200): xor / rotate can only issue on the ALU pipe, which is the
// binding limit of this function; the message add is therefore written as a multiply-add by an opaque 1 so that it
// goes to the otherwise idle FMA pipe (IMAD) instead of becoming a 3-input IADD3 on the ALU pipe.  +19 % throughput.
ZN_D uint32_t add_fma(uint32_t x, uint32_t one, uint32_t acc) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(one), "r"(acc));
  return d;
}

#define ZN_G(a, b, c, d, mx, my)   \
  a = add_fma((mx), one, a + b);   \
  d = rotr16(d ^ a);               \
  c = c + d;                       \
  b = rotr12(b ^ c);               \
  a = add_fma((my), one, a + b);   \
  d = rotr8(d ^ a);                \
  c = c + d;                       \
  b = rotr7(b ^ c);

#define ZN_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  ZN_G(v0, v4, v8, v12, m[s0], m[s1])                                                  \
  ZN_G(v1, v5, v9, v13, m[s2], m[s3])                                                  \
  ZN_G(v2, v6, v10, v14, m[s4], m[s5])                                                 \
  ZN_G(v3, v7, v11, v15, m[s6], m[s7])                                                 \
  ZN_G(v0, v5, v10, v15, m[s8], m[s9])                                                 \
  ZN_G(v1, v6, v11, v12, m[s10], m[s11])                                               \
  ZN_G(v2, v7, v8, v13, m[s12], m[s13])                                                \
  ZN_G(v3, v4, v9, v14, m[s14], m[s15])

// cv <- compress(cv, m, counter, block_len, flags)[0..8].  `one` must be the value 1 held in a register the compiler
// cannot see through (a kernel argument).
ZN_D void compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint32_t ctr_lo, uint32_t ctr_hi,
                   uint32_t block_len, uint32_t flags, uint32_t one) {
  uint32_t v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
  uint32_t v8 = ZN_IV0, v9 = ZN_IV1, v10 = ZN_IV2, v11 = ZN_IV3;
  uint32_t v12 = ctr_lo, v13 = ctr_hi, v14 = block_len, v15 = flags;
  ZN_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
  ZN_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
  ZN_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
  ZN_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
  ZN_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
  ZN_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
  ZN_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
  cv[0] = v0 ^ v8;
  cv[1] = v1 ^ v9;
  cv[2] = v2 ^ v10;
  cv[3] = v3 ^ v11;
  cv[4] = v4 ^ v12;
  cv[5] = v5 ^ v13;
  cv[6] = v6 ^ v14;
  cv[7] = v7 ^ v15;
}

ZN_D void set_iv(uint32_t (&cv)[8]) {
  cv[0] = ZN_IV0; cv[1] = ZN_IV1; cv[2] = ZN_IV2; cv[3] = ZN_IV3;
  cv[4] = ZN_IV4; cv[5] = ZN_IV5; cv[6] = ZN_IV6; cv[7] = ZN_IV7;
}

// Loads one full 64-byte block from global memory at any byte alignment into 16 little-endian words.
// Misaligned sources read 17 aligned words and funnel-shift; the extra word always overlaps valid bytes.
ZN_D void load_block_full(const uint8_t* __restrict__ p, uint32_t (&m)[16]) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((a & 15) == 0) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint4 t = __ldg(q + i);
      m[4 * i] = t.x; m[4 * i + 1] = t.y; m[4 * i + 2] = t.z; m[4 * i + 3] = t.w;
    }
  } else if ((a & 3) == 0) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = __ldg(q + i);
  } else {
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    uint32_t lo = __ldg(q);
#pragma unroll
    for (int i = 0; i < 16; i++) {
      uint32_t hi = __ldg(q + i + 1);
      m[i] = __funnelshift_r(lo, hi, sh);
      lo = hi;
    }
  }
}

// Loads a partial block (n < 64 valid bytes, zero padded).  Byte loads; runs at most once per chunk.
ZN_D void load_block_partial(const uint8_t* __restrict__ p, uint32_t n, uint32_t (&m)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((uint32_t)(4 * i + k) < n) w |= (uint32_t)__ldg(p + 4 * i + k) << (8 * k);
    m[i] = w;
  }
}

// Chaining value of one chunk: `len` in 0..1024 bytes at p, chunk counter `ctr`; `root` marks a single-chunk input.
ZN_D void hash_chunk(const uint8_t* __restrict__ p, uint32_t len, uint64_t ctr, bool root, uint32_t (&cv)[8], uint32_t one) {
  set_iv(cv);
  const uint32_t nblocks = len == 0 ? 1u : (len + 63u) >> 6;
  const uint32_t clo = (uint32_t)ctr, chi = (uint32_t)(ctr >> 32);
  uint32_t m[16];
#pragma unroll 1
  for (uint32_t b = 0; b < nblocks; b++) {
    const uint32_t n = min(64u, len - b * 64u);
    uint32_t flags = (b == 0 ? CHUNK_START : 0u);
    if (b + 1 == nblocks) flags |= CHUNK_END | (root ? ROOT : 0u);
    if (n == 64) load_block_full(p + b * 64u, m);
    else load_block_partial(p + b * 64u, n, m);
    compress(cv, m, clo, chi, n, flags, one);
  }
}

// parent node: cv_out = compress(IV, left || right, 0, 64, PARENT [| ROOT])
ZN_D void parent(const uint32_t (&l)[8], const uint32_t (&r)[8], bool root, uint32_t (&out)[8], uint32_t one) {
  uint32_t m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
  set_iv(out);
  compress(out, m, 0, 0, 64, PARENT | (root ? ROOT : 0u), one);
}

ZN_D void load_cv(const uint32_t* p, uint32_t (&cv)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  cv[0] = a.x; cv[1] = a.y; cv[2] = a.z; cv[3] = a.w; cv[4] = b.x; cv[5] = b.y; cv[6] = b.z; cv[7] = b.w;
}
ZN_D void store_cv(uint32_t* p, const uint32_t (&cv)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(cv[0], cv[1], cv[2], cv[3]);
  *reinterpret_cast<uint4*>(p + 4) = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

}  // namespace b3
}  // namespace zn
