// Bit-stream readers and byte-granular global-memory helpers shared by the codec kernels.
// Everything here is warp-uniform scalar code unless it takes a `lane` argument.  It also compiles for the
// host (tests/host_decode_check drives the same parsing logic on the CPU).
#pragma once
#include "common.cuh"

namespace zn {

#if defined(__CUDA_ARCH__)
ZN_D uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
ZN_D int hibit32(uint32_t v) { return 31 - __clz((int)v); }
#else
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31;
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
inline int hibit32(uint32_t v) { int r = 0; while (v >>= 1) r++; return r; }
#endif

// Non-blocking prefetch of the line that holds p.  A lane that walks its own bit stream has ONE dependent load in flight;
// a sector it has not seen yet costs thousands of cycles (DRAM + address translation, every lane in another page), so the
// lane-per-stream decoders ask for their bytes a few hundred bytes before they get there.
#if defined(__CUDA_ARCH__)
ZN_D void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
ZN_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
inline void prefetch_l1(const void*) {}
inline void prefetch_l2(const void*) {}
#endif

ZN_HD uint32_t ld8(const uint8_t* p) { return *p; }
ZN_HD uint32_t ld16le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
ZN_HD uint32_t ld24le(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
ZN_HD uint32_t ld32le(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Aligned 32-bit word that contains byte address `a` is always safe to read when byte `a` itself is valid.
ZN_HD uint32_t ld_word_aligned(const uint8_t* a) {
  return *reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)3);
}

// ---------------------------------------------------------------------------------------------------------
// Backward bit reader (RFC 8878 §4.1 / §4.2.2): the stream occupies bytes [p, p+len); its last byte holds the
// end marker (highest set bit).  Bits are consumed from the top down.  A 64-bit window is refilled with ALIGNED
// 32-bit words; bytes below the stream start are masked to zero, so over-reads see zeros and are detected by
// bits_left going negative (never by a fault).
// ---------------------------------------------------------------------------------------------------------
// ZN_BACKBITS_DEPTH (per translation unit; the struct's name carries it, so the two forms never meet): how many words the
// reader keeps loaded ahead of the window.  2 in the device-wide pipeline's lane-per-block sequence decoder (one L2 round
// trip must fit between two refills); 1 in the one-team decoders — their thread 0 shares an SM with the hash warps of the
// fused kernel, where the extra register showed up as spills (k_decode_ws: 468 -> 496 B) and ~2 % of the metric.
#ifndef ZN_BACKBITS_DEPTH
#define ZN_BACKBITS_DEPTH 2
#endif
#define ZN_BB_CAT2(a, b) a##b
#define ZN_BB_CAT(a, b) ZN_BB_CAT2(a, b)
#define BackBits ZN_BB_CAT(BackBits_depth, ZN_BACKBITS_DEPTH)
struct BackBits {
  static constexpr int kDepth = ZN_BACKBITS_DEPTH;
  const uint32_t* wbase;  // aligned word holding the first stream byte
  uint32_t lowmask;       // clears the bytes of word 0 that precede the stream
  int32_t widx;           // index of the word held in `pre` (descending); < 0 -> zeros
  uint32_t pre2;          // word widx - 1, RAW, loaded two refills ahead
  uint32_t pre;           // word widx, RAW (lowmask not applied yet), loaded one refill ahead: nothing touches it until the
                          // next refill, so its memory latency stays off the critical path
  uint64_t win;           // unread bits, MSB-aligned
  int32_t navail;         // valid bits in win
  int32_t bits_left;      // unread bits in the stream; < 0 == over-read

  ZN_HD uint32_t fetch(int32_t i) const {
    if (i < 0) return 0u;
    const uint32_t w = wbase[i];
    return i == 0 ? (w & lowmask) : w;
  }
  ZN_HD uint32_t fetch_raw(int32_t i) const { return i < 0 ? 0u : wbase[i]; }
  // after widx moved down by one: the word for the next refill
  ZN_HD void advance() {
    if (kDepth > 1) {
      pre = pre2; pre2 = fetch_raw(widx - 1);
      const int32_t f1 = widx - 64, f2 = widx - 256;  // lane-per-stream decoders: 256 B / 1 KiB ahead (backward stream)
      prefetch_l1(wbase + (f1 > 0 ? f1 : 0));
      prefetch_l2(wbase + (f2 > 0 ? f2 : 0));
    }
    else pre = fetch_raw(widx);
  }
  ZN_HD uint32_t pre_word() const { return widx == 0 ? (pre & lowmask) : pre; }  // the word in `pre`, masked at the time of use
  ZN_HD bool init(const uint8_t* p, uint32_t len) {
    if (len == 0) return false;
    const uint32_t last = p[len - 1];
    if (last == 0) return false;
    const int hb = hibit32(last);
    bits_left = (int32_t)(len - 1) * 8 + hb;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), s_al = a & ~(uintptr_t)3, e = a + len;
    wbase = reinterpret_cast<const uint32_t*>(s_al);
    lowmask = 0xFFFFFFFFu << ((a & 3) * 8);
    const int32_t t = (int32_t)((((e - 1) & ~(uintptr_t)3) - s_al) >> 2);
    const int nb_top = (int)(e - (s_al + 4 * (uintptr_t)t));  // 1..4 bytes of the top word belong to the stream
    const uint32_t w = fetch(t);
    const int nvalid = (nb_top - 1) * 8 + hb;  // bits below the end marker
    win = nvalid ? ((uint64_t)w << (64 - nvalid)) : 0;
    navail = nvalid;
    widx = t - 1;
    pre = fetch_raw(widx);
    pre2 = kDepth > 1 ? fetch_raw(widx - 1) : 0u;
    return true;
  }
#if defined(__CUDA_ARCH__)
  // Device fast path: the 64-bit window is handled as two 32-bit halves with clamped funnel shifts (one SHF each)
  // instead of multi-instruction 64-bit variable shifts.  When navail <= 32 every valid bit sits in the high half.
  ZN_D void refill() {  // afterwards navail > 32, so any read of <= 32 bits is served from the window
    if (navail <= 32) {
      uint32_t hi = (uint32_t)(win >> 32);
      const uint32_t n = (uint32_t)navail;
      const uint32_t p = pre_word();
      hi |= __funnelshift_rc(p, 0u, n);                      // p >> n          (0 when n == 32)
      const uint32_t lo = __funnelshift_rc(0u, p, n);        // p << (32 - n)   (0 when n == 0)
      win = ((uint64_t)hi << 32) | lo;
      navail += 32;
      widx--;
      advance();
      if (navail <= 32) {  // the window was empty: take a second word
        const uint32_t n2 = (uint32_t)navail;
        const uint32_t p2 = pre_word();
        uint32_t h2 = (uint32_t)(win >> 32);
        h2 |= __funnelshift_rc(p2, 0u, n2);
        win = ((uint64_t)h2 << 32) | __funnelshift_rc(0u, p2, n2);
        navail += 32;
        widx--;
        advance();
      }
    }
  }
  ZN_D uint32_t peek(uint32_t n) const { return __funnelshift_lc((uint32_t)(win >> 32), 0u, n); }  // n <= 32
  ZN_D void skip(uint32_t n) {
    const uint32_t hi = (uint32_t)(win >> 32), lo = (uint32_t)win;
    win = ((uint64_t)__funnelshift_lc(lo, hi, n) << 32) | __funnelshift_lc(0u, lo, n);
    navail -= (int32_t)n;
    bits_left -= (int32_t)n;
  }
#else
  ZN_HD void refill() {  // afterwards navail > 32, so any read of <= 32 bits is served from the window
    while (navail <= 32) {
      win |= (uint64_t)pre_word() << (32 - navail);
      navail += 32;
      widx--;
      advance();
    }
  }
  ZN_HD uint32_t peek(uint32_t n) const { return n ? (uint32_t)(win >> (64 - n)) : 0u; }  // n <= 32
  ZN_HD void skip(uint32_t n) { win <<= n; navail -= (int32_t)n; bits_left -= (int32_t)n; }
#endif
  ZN_HD uint32_t read(uint32_t n) {  // n <= 32; caller guarantees a refill() since the last 32 bits consumed
    const uint32_t v = peek(n);
    skip(n);
    return v;
  }
};

// Forward little-endian bit reader over a byte range (FSE table descriptions; tiny, byte loads are fine).
struct FwdBits {
  const uint8_t* p;
  uint32_t len;
  uint32_t bitpos;
  ZN_HD uint32_t peek(uint32_t n) const {  // n <= 16
    const uint32_t byte = bitpos >> 3;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (byte + i < len) v |= (uint32_t)p[byte + i] << (8 * i);
    return (v >> (bitpos & 7)) & ((1u << n) - 1u);
  }
};

}  // namespace zn
