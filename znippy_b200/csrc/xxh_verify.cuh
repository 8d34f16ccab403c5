// Content checksums of Zstandard frames (RFC 8878 §3.1.1: Content_Checksum = low 32 bits of XXH64(content, seed 0)).
// libzstd — and so the reference's decode (codec.rs:67-78 through OpenZL's zstd) — rejects a frame whose stored checksum
// disagrees with the decoded content; so does this path: right before a row's blake3 digest is compared (the tree kernels,
// one thread per row at that point) xxh_row_bad walks the frame and block headers of the row if its first frame carries
// the checksum flag, hashes the decoded bytes and turns a mismatch into S_DECODE_ERROR.  Rows without the flag (everything
// libzstd's one-shot API and this library's own compressor write) cost one 8-byte read and no kernel launch.
//
// XXH64 is four independent accumulator chains over 32-byte stripes and strictly serial along the stripes: one thread
// interleaving the four is all the parallelism a row has (an 8 MiB row takes ~5 ms; rows are independent).
// LZ4 frames: their optional XXH32 content / block checksums stay unverified (blake3 of the content supersedes them).
#pragma once
#include "common.cuh"

namespace zn {

namespace xx {
constexpr uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull,
                   P5 = 0x27D4EB2F165667C5ull;
ZN_D uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
ZN_D uint64_t round(uint64_t acc, uint64_t in) { return rotl(acc + in * P2, 31) * P1; }
ZN_D uint64_t merge(uint64_t h, uint64_t v) { return (h ^ round(0, v)) * P1 + P4; }
ZN_D uint64_t rd64(const uint8_t* p) {
  if ((reinterpret_cast<uintptr_t>(p) & 7u) == 0) return *reinterpret_cast<const uint64_t*>(p);
  uint64_t v = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
  return v;
}
ZN_D uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

}  // namespace xx

// XXH64(p[0..len), 0) by ONE thread: the four accumulators are independent chains, so one thread interleaving them is as
// fast as four lanes with one each.
ZN_D uint64_t xxh64_1(const uint8_t* p, uint64_t len) {
  uint64_t h;
  uint64_t done = 0;
  if (len >= 32) {
    uint64_t v1 = xx::P1 + xx::P2, v2 = xx::P2, v3 = 0ull, v4 = 0ull - xx::P1;
    const uint64_t stripes = len / 32;
    const uint8_t* q = p;
    for (uint64_t s = 0; s < stripes; s++, q += 32) {
      v1 = xx::round(v1, xx::rd64(q));
      v2 = xx::round(v2, xx::rd64(q + 8));
      v3 = xx::round(v3, xx::rd64(q + 16));
      v4 = xx::round(v4, xx::rd64(q + 24));
    }
    done = stripes * 32;
    h = xx::rotl(v1, 1) + xx::rotl(v2, 7) + xx::rotl(v3, 12) + xx::rotl(v4, 18);
    h = xx::merge(h, v1); h = xx::merge(h, v2); h = xx::merge(h, v3); h = xx::merge(h, v4);
  } else {
    h = xx::P5;
  }
  h += len;
  const uint8_t* q = p + done;
  const uint8_t* end = p + len;
  while (q + 8 <= end) { h ^= xx::round(0, xx::rd64(q)); h = xx::rotl(h, 27) * xx::P1 + xx::P4; q += 8; }
  if (q + 4 <= end) { h ^= (uint64_t)xx::rd32(q) * xx::P1; h = xx::rotl(h, 23) * xx::P2 + xx::P3; q += 4; }
  while (q < end) { h ^= (uint64_t)(*q++) * xx::P5; h = xx::rotl(h, 11) * xx::P1; }
  h ^= h >> 33; h *= xx::P2; h ^= h >> 29; h *= xx::P3; h ^= h >> 32;
  return h;
}

// One thread, one decoded row: true when a frame of the row carries a content checksum that disagrees with the decoded
// bytes.  Called from the tree kernels right before a row's digest is compared (no launch of its own).
ZN_D bool xxh_row_bad(const BlobDesc& d, const uint8_t* __restrict__ blobs_base, const uint8_t* __restrict__ out_base) {
  if (!(d.flags & F_COMPRESSED) || (d.flags & F_LZ4_BLOCK)) return false;
  if (d.src_len < 9 || d.src_len >= 0xFFFFFFF0ull) return false;
  const uint8_t* src = blobs_base + d.src_off;
  const uint32_t src_len = (uint32_t)d.src_len;
  const uint8_t* out = out_base + d.dst_off;
  uint32_t ip = 0, frames = 0;
  uint64_t opos = 0;
  while (ip + 5 <= src_len) {
    const uint32_t magic = xx::rd32(src + ip);
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
      if (src_len - ip < 8) break;
      const uint32_t sz = xx::rd32(src + ip + 4);
      if (sz > src_len - ip - 8) break;
      ip += 8 + sz;
      continue;
    }
    if (magic != 0xFD2FB528u) break;  // LZ4 frame or not a frame: nothing to do here
    const uint32_t fhd = src[ip + 4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
    const bool checksum = (fhd >> 2) & 1;
    // the common case ends here.  (A later frame of a multi-frame blob could still carry a checksum; such blobs are walked
    // only when the first frame announces one — libzstd writes the flag per stream, not per frame.)
    if (frames == 0 && !checksum) break;
    uint32_t hp = ip + 5 + (single ? 0u : 1u) + (did_flag == 3 ? 4u : did_flag);
    const uint32_t fb = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2u : (fcs_flag == 2 ? 4u : 8u));
    if (hp + fb > src_len) break;
    uint64_t fcs = ~0ull;
    if (fb) {
      fcs = 0;
      for (uint32_t i = 0; i < fb; i++) fcs |= (uint64_t)src[hp + i] << (8 * i);
      if (fb == 2) fcs += 256;
    }
    hp += fb;
    bool ok = false;
    for (;;) {  // block headers
      if (hp + 3 > src_len) break;
      const uint32_t bh = (uint32_t)src[hp] | (uint32_t)src[hp + 1] << 8 | (uint32_t)src[hp + 2] << 16;
      const uint32_t type = (bh >> 1) & 3u, size = bh >> 3;
      hp += 3;
      const uint32_t adv = type == 1 ? 1u : size;
      if (adv > src_len - hp) break;
      hp += adv;
      if (bh & 1u) { ok = true; break; }
    }
    if (!ok) break;
    uint32_t stored = 0;
    if (checksum) {
      if (hp + 4 > src_len) break;
      stored = xx::rd32(src + hp);
      hp += 4;
    }
    // content range of this frame: from its header, or — a lone frame without one — everything the row decoded to
    uint64_t flen;
    if (fcs != ~0ull) flen = fcs;
    else if (frames == 0 && hp == src_len) flen = d.dst_cap;
    else break;
    if (opos + flen > d.dst_cap) break;
    if (checksum && (uint32_t)xxh64_1(out + opos, flen) != stored) return true;
    opos += flen;
    ip = hp;
    frames++;
  }
  return false;
}

}  // namespace zn
