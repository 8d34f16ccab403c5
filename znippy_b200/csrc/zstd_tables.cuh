// Zstandard entropy tables (RFC 8878 §3.1.1.3.2.2, §4.1, §4.2.1): FSE normalized-count parsing, FSE and
// Huffman decoding-table construction, predefined distributions and the length/offset code baselines.
// Warp-uniform scalar code; also compiled for the host so it can be checked against the oracle on the CPU.
//
// On the reference path this is inside codec::decompress_into (znippy-common/src/codec.rs:67-78) ->
// openzl-sys-rs 0.2.0 -> zstd, not vendored; this is a from-the-spec implementation.
#pragma once
#include "bitio.cuh"

namespace zn {
namespace zs {

// decoding-table entry: sym | nbits << 8 | base << 16
ZN_HD uint32_t fse_pack(uint32_t sym, uint32_t nbits, uint32_t base) { return sym | (nbits << 8) | (base << 16); }
ZN_HD uint32_t fse_sym(uint32_t e) { return e & 0xFF; }
ZN_HD uint32_t fse_nbits(uint32_t e) { return (e >> 8) & 0xFF; }
ZN_HD uint32_t fse_base(uint32_t e) { return e >> 16; }

struct FseTable {  // lives in shared memory (or host memory in the CPU check)
  uint32_t e[512];
  uint32_t log;
  uint32_t valid;
};

#if defined(__CUDA_ARCH__)
#define ZN_CONST __constant__
#else
#define ZN_CONST static const
#endif

ZN_CONST int16_t kLLDefault[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                   2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
ZN_CONST int16_t kMLDefault[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                   1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
ZN_CONST int16_t kOFDefault[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1,
                                   1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
ZN_CONST uint32_t kLLBase[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18,
                                 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096,
                                 8192, 16384, 32768, 65536};
ZN_CONST uint8_t kLLBits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1,
                                1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
ZN_CONST uint32_t kMLBase[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20,
                                 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37, 39, 41,
                                 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195,
                                 16387, 32771, 65539};
ZN_CONST uint8_t kMLBits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1,
                                2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

// Parses an FSE table description. Returns bytes consumed, or -1.
ZN_HD int fse_read_ncount(const uint8_t* p, uint32_t len, int max_log, int max_sym, int16_t* norm, int* out_log,
                          int* out_nsym) {
  if (len == 0) return -1;
  FwdBits b{p, len, 0};
  const int log = (int)b.peek(4) + 5;
  b.bitpos += 4;
  if (log > max_log) return -1;
  int remaining = (1 << log) + 1, threshold = 1 << log, nbits = log + 1, sym = 0;
  for (int i = 0; i <= max_sym; i++) norm[i] = 0;
  while (remaining > 1 && sym <= max_sym) {
    const int max = (2 * threshold - 1) - remaining;
    const int low = (int)b.peek(nbits - 1);
    int value;
    if (low < max) {
      value = low;
      b.bitpos += nbits - 1;
    } else {
      value = (int)b.peek(nbits);
      if (value >= threshold) value -= max;
      b.bitpos += nbits;
    }
    const int count = value - 1;
    remaining -= count < 0 ? -count : count;
    norm[sym++] = (int16_t)count;
    if (count == 0) {
      for (;;) {
        const int rep = (int)b.peek(2);
        b.bitpos += 2;
        sym += rep;
        if (rep != 3) break;
        if ((b.bitpos >> 3) > len) return -1;
      }
    }
    if (remaining < 1) return -1;
    while (remaining < threshold) { nbits--; threshold >>= 1; }
    if ((b.bitpos >> 3) > len) return -1;
  }
  if (remaining != 1 || sym > max_sym + 1) return -1;
  const uint32_t used = (b.bitpos + 7) >> 3;
  if (used > len) return -1;
  *out_log = log;
  *out_nsym = sym;
  return (int)used;
}

// Builds the decoding table from normalized counts. `next` is scratch for nsym uint16 (nsym <= 256).
ZN_HD void fse_build(FseTable* t, const int16_t* norm, int nsym, int log, uint16_t* next) {
  const int size = 1 << log;
  int high = size - 1;
  for (int s = 0; s < nsym; s++) {
    if (norm[s] == -1) { t->e[high--] = (uint32_t)s; next[s] = 1; }
    else next[s] = (uint16_t)norm[s];
  }
  const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
  int pos = 0;
  for (int s = 0; s < nsym; s++)
    for (int i = 0; i < norm[s]; i++) {
      t->e[pos] = (uint32_t)s;
      do pos = (pos + step) & mask; while (pos > high);
    }
  for (int u = 0; u < size; u++) {
    const uint32_t s = t->e[u];
    const uint32_t ns = next[s]++;
    const uint32_t nb = (uint32_t)(log - hibit32(ns));
    t->e[u] = fse_pack(s, nb, (ns << nb) - (uint32_t)size);
  }
  t->log = (uint32_t)log;
  t->valid = 1;
}

ZN_HD void fse_build_rle(FseTable* t, uint32_t sym) {
  t->e[0] = fse_pack(sym, 0, 0);
  t->log = 0;
  t->valid = 1;
}

// Huffman decoding table: entry = sym | nbits << 8, indexed by the next max_bits bits of the stream.
struct HufTable {
  uint16_t e[2048];
  uint32_t max_bits;
  uint32_t valid;
};

// Huffman tree description (RFC 8878 §4.2.1). `w` is scratch for 256 weights, fse/next scratch for the
// FSE-compressed form.  Returns bytes consumed or -1.
ZN_HD int huf_read_table(HufTable* h, const uint8_t* p, uint32_t len, uint8_t* w, FseTable* fse_scratch,
                         uint16_t* next_scratch) {
  if (len < 1) return -1;
  const uint32_t hb = p[0];
  int n = 0, used;
  if (hb >= 128) {
    n = (int)hb - 127;
    const uint32_t bytes = (uint32_t)(n + 1) / 2;
    if (1 + bytes > len) return -1;
    for (int i = 0; i < n; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
    used = 1 + (int)bytes;
  } else {
    if (hb == 0 || hb + 1 > len) return -1;
    int16_t norm[16];
    int log, nsym;
    const int hd = fse_read_ncount(p + 1, hb, 6, 11, norm, &log, &nsym);
    if (hd < 0 || (uint32_t)hd >= hb) return -1;
    fse_build(fse_scratch, norm, nsym, log, next_scratch);
    BackBits b;
    if (!b.init(p + 1 + hd, hb - (uint32_t)hd)) return -1;
    b.refill();
    uint32_t s1 = b.read((uint32_t)log), s2 = b.read((uint32_t)log);
    if (b.bits_left < 0) return -1;
    for (;;) {
      if (n >= 254) return -1;
      b.refill();
      uint32_t e1 = fse_scratch->e[s1];
      w[n++] = (uint8_t)fse_sym(e1);
      s1 = fse_base(e1) + b.read(fse_nbits(e1));
      if (b.bits_left < 0) { w[n++] = (uint8_t)fse_sym(fse_scratch->e[s2]); break; }
      if (n >= 254) return -1;
      uint32_t e2 = fse_scratch->e[s2];
      w[n++] = (uint8_t)fse_sym(e2);
      s2 = fse_base(e2) + b.read(fse_nbits(e2));
      if (b.bits_left < 0) { w[n++] = (uint8_t)fse_sym(fse_scratch->e[s1]); break; }
    }
    if (n > 255) return -1;
    used = 1 + (int)hb;
  }
  uint32_t sum = 0;
  for (int i = 0; i < n; i++) {
    if (w[i] > 11) return -1;
    if (w[i]) sum += 1u << (w[i] - 1);
  }
  if (sum == 0) return -1;
  const int max_bits = hibit32(sum) + 1;
  if (max_bits > 11) return -1;
  const uint32_t left = (1u << max_bits) - sum;
  if (left & (left - 1)) return -1;
  w[n++] = (uint8_t)(hibit32(left) + 1);
  uint32_t rank_start[13], count[13];
  for (int i = 0; i < 13; i++) count[i] = 0;
  for (int i = 0; i < n; i++) count[w[i]]++;
  uint32_t pos = 0;
  for (int wt = 1; wt <= max_bits; wt++) { rank_start[wt] = pos; pos += count[wt] << (wt - 1); }
  if (pos != (1u << max_bits)) return -1;
  for (int s = 0; s < n; s++) {
    const uint32_t wt = w[s];
    if (!wt) continue;
    const uint32_t span = 1u << (wt - 1), start = rank_start[wt];
    const uint16_t ent = (uint16_t)((uint32_t)s | ((uint32_t)(max_bits + 1 - wt) << 8));
    for (uint32_t k = 0; k < span; k++) h->e[start + k] = ent;
    rank_start[wt] = start + span;
  }
  h->max_bits = (uint32_t)max_bits;
  h->valid = 1;
  return used;
}

}  // namespace zs
}  // namespace zn
