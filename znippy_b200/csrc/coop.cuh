// Team-cooperative byte movers used by the codec kernels: unaligned bulk copy, fill, and LZ77 match copy with
// overlap (periodic) handling.  A "team" is every thread of the CTA that decodes one blob; the bytes of one
// operation are spread over the team with 128-bit stores.  The host versions (team of one, plain loops) let the
// surrounding team-uniform parsing code be exercised on the CPU (tests/host_emu).
#pragma once
#include "bitio.cuh"

namespace zn {

constexpr uint32_t kPatWords = 144;   // periodic-match pattern buffer: periods <= kPatMaxOff plus 20 bytes run-out
constexpr uint32_t kPatMaxOff = 512;
constexpr uint32_t kShortCopy = 48;   // below this a copy is one byte per thread
constexpr uint32_t kTileBytes = 8192;   // shared-memory tile that long periodic matches are bulk-stored from
constexpr uint32_t kBulkMin = 32768;    // shortest periodic match that takes the bulk-store path
constexpr uint32_t kBulkIssuers = 32;   // threads that may issue (and must therefore wait for) bulk stores

struct Team {
  uint32_t tid, n;
  uint32_t bar = 0;  // 0: the team is the whole CTA (bar.sync 0); else the named barrier of a sub-CTA team of n threads
};

#if defined(__CUDA_ARCH__)

// A team is either a whole CTA or, for small blobs, a CTA of exactly one warp.
ZN_D void team_sync(const Team& t) {
  if (t.n == 32) __syncwarp();
  else if (t.bar == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(t.bar), "r"(t.n) : "memory");
}

// ---- TMA bulk stores (cp.async.bulk shared -> global): one instruction moves up to a whole tile, so a 128 KiB
// periodic match is ~8 instructions from one thread instead of 8192 STG.128 spread over the team.
ZN_D void bulk_store(uint8_t* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(__cvta_generic_to_global(gdst)),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
ZN_D void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
ZN_D void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
ZN_D void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// 16 bytes from any alignment, assembled from aligned 32-bit words (the extra word always overlaps valid bytes)
ZN_D uint4 load16_any(const uint8_t* s) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
  if ((a & 15) == 0) return *reinterpret_cast<const uint4*>(s);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
  if (sh == 0) return make_uint4(w0, w1, w2, w3);
  const uint32_t w4 = q[4];
  return make_uint4(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh), funnel_r(w2, w3, sh), funnel_r(w3, w4, sh));
}

// 16 bytes at byte address (aligned word pointer q, bit shift sh = 8 * misalignment), sh != 0
ZN_D uint4 load16_shift(const uint32_t* q, uint32_t sh) {
  const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
  return make_uint4(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh), funnel_r(w2, w3, sh), funnel_r(w3, w4, sh));
}

// dst[0..n) = src[0..n).  Every source byte is final and visible to the team; dst does not feed src.
// The vector body keeps kCopyUnroll 16-byte loads in flight per thread before the first store: a lone CTA copying
// 128 KiB would otherwise pay one full memory round trip per 4 KiB.
constexpr int kCopyUnroll = 8;
ZN_D void team_copy(const Team& t, uint8_t* dst, const uint8_t* src, uint32_t n) {
  if (n <= kShortCopy) {
    for (uint32_t i = t.tid; i < n; i += t.n) dst[i] = src[i];
    return;
  }
  const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;  // < n
  if (t.tid < head) dst[t.tid] = src[t.tid];
  const uint32_t nvec = (n - head) >> 4;
  const uint8_t* s = src + head;
  uint4* d = reinterpret_cast<uint4*>(dst + head);
  const uintptr_t sa = reinterpret_cast<uintptr_t>(s);
  uint32_t v = t.tid;
  if ((sa & 3) == 0) {  // word-aligned source: 128-bit (or 4 x 32-bit) loads
    if ((sa & 15) == 0) {
      const uint4* q = reinterpret_cast<const uint4*>(s);
      for (; v + (kCopyUnroll - 1) * t.n < nvec; v += kCopyUnroll * t.n) {
        uint4 r[kCopyUnroll];
#pragma unroll
        for (int u = 0; u < kCopyUnroll; u++) r[u] = q[v + u * t.n];
#pragma unroll
        for (int u = 0; u < kCopyUnroll; u++) d[v + u * t.n] = r[u];
      }
      for (; v < nvec; v += t.n) d[v] = q[v];
    } else {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(s);
      for (; v + (kCopyUnroll - 1) * t.n < nvec; v += kCopyUnroll * t.n) {
        uint4 r[kCopyUnroll];
#pragma unroll
        for (int u = 0; u < kCopyUnroll; u++) {
          const uint32_t* w = q + 4u * (v + u * t.n);
          r[u] = make_uint4(w[0], w[1], w[2], w[3]);
        }
#pragma unroll
        for (int u = 0; u < kCopyUnroll; u++) d[v + u * t.n] = r[u];
      }
      for (; v < nvec; v += t.n) {
        const uint32_t* w = q + 4u * v;
        d[v] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  } else {  // byte-misaligned source: aligned words + funnel shifts (the extra word always overlaps valid bytes)
    const uint32_t* q = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(sa & 3) * 8;
    for (; v + (kCopyUnroll / 2 - 1) * t.n < nvec; v += (kCopyUnroll / 2) * t.n) {
      uint4 r[kCopyUnroll / 2];
#pragma unroll
      for (int u = 0; u < kCopyUnroll / 2; u++) r[u] = load16_shift(q + 4u * (v + u * t.n), sh);
#pragma unroll
      for (int u = 0; u < kCopyUnroll / 2; u++) d[v + u * t.n] = r[u];
    }
    for (; v < nvec; v += t.n) d[v] = load16_shift(q + 4u * v, sh);
  }
  const uint32_t k = head + (nvec << 4) + t.tid;
  if (k < n) dst[k] = src[k];
}

ZN_D void team_fill(const Team& t, uint8_t* dst, uint32_t byte, uint32_t n) {
  if (n <= kShortCopy) {
    for (uint32_t i = t.tid; i < n; i += t.n) dst[i] = (uint8_t)byte;
    return;
  }
  const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
  if (t.tid < head) dst[t.tid] = (uint8_t)byte;
  const uint32_t nvec = (n - head) >> 4;
  const uint32_t w = byte * 0x01010101u;
  uint4* d = reinterpret_cast<uint4*>(dst + head);
  for (uint32_t v = t.tid; v < nvec; v += t.n) d[v] = make_uint4(w, w, w, w);
  const uint32_t k = head + (nvec << 4) + t.tid;
  if (k < n) dst[k] = (uint8_t)byte;
}

// LZ77 match with byte-serial semantics: dst[k] = dst[k - off] for k in [0, ml).  Because every byte of the match
// equals window[k mod off] with window = dst[-off .. 0), the whole match is a gather from bytes that existed before
// it started: no ordering between the team's stores is needed.  Precondition: the window is final and visible to
// the team (the caller barriers when it was written since the last barrier).  `pat` = kPatWords of shared memory.
// Contains team_sync() on the small-period path, so every thread of the team must call it with the same arguments.
// `tile` = kTileBytes of 128-byte aligned shared memory.  Returns true when part of the match was issued as bulk
// (async-proxy) stores that are still in flight: the caller must bulk_wait_all() + barrier before anyone reads those
// bytes or before tile / pat are rewritten (zstd_decode.cuh: mem_sync).
ZN_D bool team_match(const Team& t, uint8_t* dst, uint32_t off, uint32_t ml, uint32_t* pat, uint8_t* tile) {
  if (off >= ml) {
    team_copy(t, dst, dst - off, ml);
    return false;
  }
  if (ml <= kShortCopy) {
    for (uint32_t i = t.tid; i < ml; i += t.n) dst[i] = dst[(int32_t)(i % off) - (int32_t)off];
    return false;
  }
  if (off <= kPatMaxOff && ml >= kBulkMin && tile != nullptr) {
    // The aligned body of the match is periodic with period lcm(off, 16): build one tile holding a whole number of
    // such periods in shared memory, then bulk-store the same tile back to back over the body.
    uint8_t* pat8 = reinterpret_cast<uint8_t*>(pat);
    for (uint32_t x = t.tid; x < off + 20; x += t.n) pat8[x] = dst[(int32_t)(x % off) - (int32_t)off];
    team_sync(t);
    ZN_TP(10);
    const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
    const uint32_t body = (ml - head) & ~15u;
    const uint32_t g = min(16u, off & (0u - off));      // gcd(off, 16)
    const uint32_t period = off * (16u / g);             // lcm(off, 16) <= 8192
    const uint32_t tlen = period * (kTileBytes / period);
    const uint32_t fill = min(tlen, body);
    {
      uint32_t phase = (head + 16u * t.tid) % off;
      const uint32_t step = (16u * t.n) % off;
      uint4* tv = reinterpret_cast<uint4*>(tile);
      for (uint32_t v = t.tid; v < (fill >> 4); v += t.n) {
        const uint32_t sh = (phase & 3) * 8, wi = phase >> 2;
        const uint32_t w0 = pat[wi], w1 = pat[wi + 1], w2 = pat[wi + 2], w3 = pat[wi + 3], w4 = pat[wi + 4];
        tv[v] = make_uint4(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh), funnel_r(w2, w3, sh), funnel_r(w3, w4, sh));
        phase += step;
        if (phase >= off) phase -= off;
      }
    }
    if (t.tid < head) dst[t.tid] = pat8[t.tid % off];
    const uint32_t k = head + body + t.tid;
    if (k < ml) dst[k] = pat8[k % off];
    fence_async_smem();
    team_sync(t);
    ZN_TP(11);
    // one bulk store per thread (bulk groups are per thread: mem_sync makes threads < kBulkIssuers wait)
    for (uint32_t o = t.tid * tlen; o < body; o += kBulkIssuers * tlen) {
      if (t.tid < kBulkIssuers) {
        bulk_store(dst + head + o, tile, min(tlen, body - o));
        bulk_commit();
      }
    }
    ZN_TP(12);
    return true;
  }
  const uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
  const uint32_t nvec = (ml - head) >> 4;
  const uint32_t step = (16u * t.n) % off;
  if (off <= kPatMaxOff) {
    // stage one period (plus 20 bytes of wrap-around) in shared memory, then every thread streams 16-byte stores
    // whose source is the pattern at its own phase.
    uint8_t* pat8 = reinterpret_cast<uint8_t*>(pat);
    for (uint32_t x = t.tid; x < off + 20; x += t.n) pat8[x] = dst[(int32_t)(x % off) - (int32_t)off];
    team_sync(t);
    if (t.tid < head) dst[t.tid] = pat8[t.tid % off];
    uint32_t phase = (head + 16u * t.tid) % off;
    uint4* dv = reinterpret_cast<uint4*>(dst + head);
#pragma unroll 2
    for (uint32_t v = t.tid; v < nvec; v += t.n) {
      const uint32_t sh = (phase & 3) * 8, wi = phase >> 2;
      const uint32_t w0 = pat[wi], w1 = pat[wi + 1], w2 = pat[wi + 2], w3 = pat[wi + 3], w4 = pat[wi + 4];
      dv[v] = make_uint4(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh), funnel_r(w2, w3, sh), funnel_r(w3, w4, sh));
      phase += step;
      if (phase >= off) phase -= off;
    }
    const uint32_t k = head + (nvec << 4) + t.tid;
    if (k < ml) dst[k] = pat8[k % off];
    return false;
  }
  // long period: every span of `off` bytes is a copy of the same window, which existed before the match started —
  // so the spans are independent vectorised copies (no barrier between them)
  for (uint32_t done = 0; done < ml; done += off) team_copy(t, dst + done, dst - off, min(off, ml - done));
  return false;
}

#else  // ------------------------------------------------------------------ host emulation (team of one)

inline void team_sync(const Team&) {}
inline void team_copy(const Team&, uint8_t* dst, const uint8_t* src, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) dst[i] = src[i];
}
inline void team_fill(const Team&, uint8_t* dst, uint32_t byte, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) dst[i] = (uint8_t)byte;
}
inline bool team_match(const Team&, uint8_t* dst, uint32_t off, uint32_t ml, uint32_t*, uint8_t*) {
  for (uint32_t i = 0; i < ml; i++) dst[i] = dst[(int64_t)i - (int64_t)off];
  return false;
}
inline void bulk_wait_all() {}

#endif

}  // namespace zn
