// Warp-cooperative byte movers used by the codec kernels: unaligned bulk copy, fill, and LZ77 match copy with
// overlap (periodic) handling.  Device versions spread the bytes over the 32 lanes with 128-bit stores; the host
// versions are plain loops so the surrounding (warp-uniform) parsing code can be exercised on the CPU.
#pragma once
#include "bitio.cuh"

namespace zn {

constexpr uint32_t kPatWords = 144;  // periodic-match pattern buffer: offsets < 512 plus 20 bytes of run-out

#if defined(__CUDA_ARCH__)

ZN_D uint32_t lane_id() { return threadIdx.x & 31u; }
ZN_D void warp_sync() { __syncwarp(); }

// 16 bytes from any alignment, assembled from aligned 32-bit words (the extra word always overlaps valid bytes)
ZN_D uint4 load16_any(const uint8_t* s) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(s);
  if ((a & 15) == 0) return *reinterpret_cast<const uint4*>(s);
  const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
  if (sh == 0) return make_uint4(w0, w1, w2, w3);
  uint32_t w4 = q[4];
  return make_uint4(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh), funnel_r(w2, w3, sh), funnel_r(w3, w4, sh));
}

// dst[0..n) = src[0..n); ranges must not overlap in a way that makes src depend on this call's writes.
ZN_D void coop_copy(uint8_t* dst, const uint8_t* src, uint32_t n) {
  const uint32_t lane = lane_id();
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
  if (head > n) head = n;
  if (lane < head) dst[lane] = src[lane];
  const uint32_t nvec = (n - head) >> 4;
  const uint8_t* s = src + head;
  uint4* d = reinterpret_cast<uint4*>(dst + head);
#pragma unroll 4
  for (uint32_t v = lane; v < nvec; v += 32) d[v] = load16_any(s + 16u * v);
  const uint32_t k = head + (nvec << 4) + lane;
  if (k < n) dst[k] = src[k];
}

ZN_D void coop_fill(uint8_t* dst, uint32_t byte, uint32_t n) {
  const uint32_t lane = lane_id();
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
  if (head > n) head = n;
  if (lane < head) dst[lane] = (uint8_t)byte;
  const uint32_t nvec = (n - head) >> 4;
  const uint32_t w = byte * 0x01010101u;
  uint4* d = reinterpret_cast<uint4*>(dst + head);
  for (uint32_t v = lane; v < nvec; v += 32) d[v] = make_uint4(w, w, w, w);
  const uint32_t k = head + (nvec << 4) + lane;
  if (k < n) dst[k] = (uint8_t)byte;
}

// LZ77 match: out[d .. d+ml) = out[d-off .. ), byte-serial semantics (off < ml replicates the last `off` bytes).
// All bytes below d are final and visible to the warp.  `pat` is a per-warp shared-memory scratch of kPatWords.
ZN_D void coop_match(uint8_t* out, uint64_t d, uint32_t off, uint32_t ml, uint32_t* pat) {
  const uint32_t lane = lane_id();
  uint8_t* dst = out + d;
  if (off >= ml) {
    coop_copy(dst, dst - off, ml);
    return;
  }
  if (off >= 512) {  // spans of `off` bytes: each span only reads bytes completed by earlier spans
    for (uint32_t done = 0; done < ml; done += off) {
      const uint32_t n = min(off, ml - done);
      coop_copy(dst + done, dst + done - off, n);
      warp_sync();
    }
    return;
  }
  // periodic fill: stage the period (plus 20 bytes of wrap-around) in shared memory once, then every lane
  // streams 16-byte stores whose source is the pattern at its own phase.
  uint8_t* pat8 = reinterpret_cast<uint8_t*>(pat);
  for (uint32_t x = lane; x < off + 20; x += 32) pat8[x] = dst[(int64_t)(x % off) - (int64_t)off];
  warp_sync();
  uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15)) & 15u;
  if (head > ml) head = ml;
  if (lane < head) dst[lane] = pat8[lane % off];
  const uint32_t nvec = (ml - head) >> 4;
  uint32_t phase = (head + 16u * lane) % off;
  const uint32_t step = 512u % off;
  uint4* dv = reinterpret_cast<uint4*>(dst + head);
#pragma unroll 2
  for (uint32_t v = lane; v < nvec; v += 32) {
    const uint32_t sh = (phase & 3) * 8, wi = phase >> 2;
    const uint32_t w0 = pat[wi], w1 = pat[wi + 1], w2 = pat[wi + 2], w3 = pat[wi + 3], w4 = pat[wi + 4];
    dv[v] = make_uint4(funnel_r(w0, w1, sh), funnel_r(w1, w2, sh), funnel_r(w2, w3, sh), funnel_r(w3, w4, sh));
    phase += step;
    if (phase >= off) phase -= off;
  }
  const uint32_t k = head + (nvec << 4) + lane;
  if (k < ml) dst[k] = pat8[k % off];
  warp_sync();
}

#else  // ------------------------------------------------------------------ host emulation (one "lane")

inline uint32_t lane_id() { return 0; }
inline void warp_sync() {}
inline void coop_copy(uint8_t* dst, const uint8_t* src, uint32_t n) { for (uint32_t i = 0; i < n; i++) dst[i] = src[i]; }
inline void coop_fill(uint8_t* dst, uint32_t byte, uint32_t n) { for (uint32_t i = 0; i < n; i++) dst[i] = (uint8_t)byte; }
inline void coop_match(uint8_t* out, uint64_t d, uint32_t off, uint32_t ml, uint32_t*) {
  for (uint32_t i = 0; i < ml; i++) out[d + i] = out[d + i - off];
}

#endif

}  // namespace zn
