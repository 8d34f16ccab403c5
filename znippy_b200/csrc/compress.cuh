// Per-slice compressors (write side): LZ4 blocks and Zstandard compressed blocks, one warp per block.
//
// Replaces CompressCtx::compress_into (znippy-common/src/codec.rs:43-55) as called from the barrel / worker bodies
// (znippy-compress/src/stream_packer.rs:230, slot_packer.rs:570).  The reference reaches zstd level 19 through OpenZL
// (not vendored); this is a GPU-shaped design written from the LZ4 block/frame and RFC 8878 formats:
//
//   * a slice is cut into independent blocks (LZ4: 64 KiB frame blocks; zstd: 128 KiB blocks) and every block of
//     every slice of the batch is compressed by its own warp — 8 MiB slices give 64-128 blocks each, so a 500 MiB
//     batch is ~4-8 k warps of work;
//   * match finder: the 32 lanes hash 32 consecutive positions at once into a shared-memory table, ballot picks the
//     first lane with a verified 4-byte match, and the match is extended 128 bytes per step across the warp;
//   * zstd blocks carry raw literals and sequences coded with the PREDEFINED FSE tables (modes byte 0), never a
//     repeat-offset code, so blocks do not depend on each other and stock libzstd decodes the frame;
//     the match window of a block is primed with the tail of the previous block (the window is the whole slice);
//   * every block lands in a scratch slot; a second kernel lays the blocks out back to back behind the frame header.
//
// Everything here is warp-uniform code over a `Warp` abstraction (32 lanes on the device, 1 lane on the host) so the
// format logic is exercised on the CPU against stock libzstd / liblz4 (tests/host_emu).
#pragma once
#include "bitio.cuh"
#include "zstd_tables.cuh"

namespace zn {
namespace cz {

struct Warp {
  uint32_t lane, n;
};

#if defined(__CUDA_ARCH__)
ZN_D uint32_t w_ballot(const Warp&, bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
ZN_D uint32_t w_shfl(const Warp&, uint32_t v, uint32_t src) { return __shfl_sync(0xFFFFFFFFu, v, src); }
ZN_D void w_sync(const Warp&) { __syncwarp(); }
ZN_D uint32_t ffs32(uint32_t v) { return (uint32_t)__ffs((int)v); }
#else
inline uint32_t w_ballot(const Warp&, bool p) { return p ? 1u : 0u; }
inline uint32_t w_shfl(const Warp&, uint32_t v, uint32_t) { return v; }
inline void w_sync(const Warp&) {}
inline uint32_t ffs32(uint32_t v) { return v ? (uint32_t)__builtin_ffs((int)v) : 0u; }
#endif

ZN_HD uint32_t hash4(uint32_t v, uint32_t hlog) { return (v * 2654435761u) >> (32 - hlog); }

constexpr uint32_t kLz4Block = 64u * 1024u;
constexpr uint32_t kLz4Slot = kLz4Block + kLz4Block / 255u + 32u;  // worst-case LZ4 block
constexpr uint32_t kLz4HashLog = 12;                                 // 4096 x u16 = 8 KiB per warp
constexpr uint32_t kZstdHashLog = 12;                                // 4096 x u32 = 16 KiB per warp
constexpr uint32_t kZstdPrime = 1024;                                // bytes of the previous block hashed in first
constexpr uint32_t kZstdMaxSeq = 32768;                              // sequences per 128 KiB block (min match 4)
constexpr uint32_t kZstdSlot = kZstdBlockMax + 64u;                  // staged block payload

// Length of the common prefix of a[0..max) and b[0..max), found 4 bytes per lane per step.
ZN_HD uint32_t match_extend(const Warp& w, const uint8_t* a, const uint8_t* b, uint32_t max) {
  uint32_t len = 0;
  for (;;) {
    const uint32_t idx = len + 4u * w.lane;
    uint32_t neq;  // number of equal leading bytes in this lane's word, 0..4
    if (idx + 4u <= max) {
      const uint32_t x = ld32le(a + idx) ^ ld32le(b + idx);
      neq = x == 0 ? 4u : ((ffs32(x) - 1u) >> 3);
    } else {
      neq = 0;
      while (idx + neq < max && a[idx + neq] == b[idx + neq]) neq++;
    }
    const uint32_t stop = w_ballot(w, neq < 4u);
    if (stop) {
      const uint32_t first = ffs32(stop) - 1u;
      return len + 4u * first + w_shfl(w, neq, first);
    }
    len += 4u * w.n;
  }
}

// dst[0..n) = src[0..n), spread over the warp (byte granular; the compressed side is small by construction)
ZN_HD void w_copy(const Warp& w, uint8_t* dst, const uint8_t* src, uint32_t n) {
  for (uint32_t i = w.lane; i < n; i += w.n) dst[i] = src[i];
}

// LZ4 length extension: value v (>= 0) as 255,255,...,rest.  Returns bytes written.
ZN_HD uint32_t lz4_put_len(const Warp& w, uint8_t* dst, uint32_t v) {
  const uint32_t n255 = v / 255u;
  for (uint32_t i = w.lane; i < n255; i += w.n) dst[i] = 255;
  if (w.lane == 0) dst[n255] = (uint8_t)(v - n255 * 255u);
  return n255 + 1u;
}

// ------------------------------------------------------------------------------------------------- LZ4 block
// Compresses in[0..n) (n <= 64 KiB) into out (capacity kLz4Slot).  tab = 4096 x u16 (shared memory), zeroed here.
// Returns the compressed size (may exceed n for incompressible input; the caller then stores the block raw).
ZN_HD uint32_t lz4_compress_block(const Warp& w, const uint8_t* in, uint32_t n, uint8_t* out, uint16_t* tab) {
  for (uint32_t i = w.lane; i < (1u << kLz4HashLog); i += w.n) tab[i] = 0;
  w_sync(w);
  uint32_t op = 0, anchor = 0, pos = 0;
  if (n >= 13) {
    const uint32_t mflimit = n - 12, matchlimit = n - 5;
    while (pos < mflimit) {
      const uint32_t p = pos + w.lane;
      const bool valid = p < mflimit;
      const uint32_t v = valid ? ld32le(in + p) : 0u;
      const uint32_t h = hash4(v, kLz4HashLog);
      const uint32_t cand = valid ? tab[h] : 0u;
      const bool ok = valid && cand < p && ld32le(in + cand) == v;
      const uint32_t m = w_ballot(w, ok);
      const uint32_t f = m ? ffs32(m) - 1u : w.n;
      // insert only the positions up to the match: later lanes are looked up again next step (or lie inside the
      // match) and must not find themselves in the table instead of their real candidate
      w_sync(w);
      if (valid && w.lane <= f) tab[h] = (uint16_t)p;
      w_sync(w);
      if (!m) {
        pos += w.n;
        continue;
      }
      const uint32_t mp = pos + f, mc = w_shfl(w, cand, f);
      const uint32_t ml = 4u + match_extend(w, in + mp + 4, in + mc + 4, matchlimit - (mp + 4));
      // ---- emit: token, literal length, literals, offset, match length
      const uint32_t ll = mp - anchor, mlc = ml - 4u;
      if (w.lane == 0) out[op] = (uint8_t)(((ll < 15u ? ll : 15u) << 4) | (mlc < 15u ? mlc : 15u));
      op += 1;
      if (ll >= 15u) op += lz4_put_len(w, out + op, ll - 15u);
      w_copy(w, out + op, in + anchor, ll);
      op += ll;
      if (w.lane == 0) {
        out[op] = (uint8_t)(mp - mc);
        out[op + 1] = (uint8_t)((mp - mc) >> 8);
      }
      op += 2;
      if (mlc >= 15u) op += lz4_put_len(w, out + op, mlc - 15u);
      pos = anchor = mp + ml;
    }
  }
  // last sequence: literals only
  const uint32_t ll = n - anchor;
  if (w.lane == 0) out[op] = (uint8_t)((ll < 15u ? ll : 15u) << 4);
  op += 1;
  if (ll >= 15u) op += lz4_put_len(w, out + op, ll - 15u);
  w_copy(w, out + op, in + anchor, ll);
  op += ll;
  w_sync(w);
  return op;
}

// LZ4 frame header for a slice of `content` bytes (15 bytes): independent 64 KiB blocks, content size present.
ZN_HD uint32_t lz4_frame_header(uint8_t* dst, uint64_t content) {
  dst[0] = 0x04; dst[1] = 0x22; dst[2] = 0x4D; dst[3] = 0x18;
  dst[4] = 0x68;  // version 01, B.Indep, C.Size
  dst[5] = 0x40;  // 64 KiB blocks
  for (int i = 0; i < 8; i++) dst[6 + i] = (uint8_t)(content >> (8 * i));
  // XXH32 of the 10 descriptor bytes (len < 16 path)
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  uint32_t h = P5 + 10u;
  for (int i = 0; i < 8; i += 4) {
    h += ld32le(dst + 4 + i) * P3;
    h = ((h << 17) | (h >> 15)) * P4;
  }
  for (int i = 8; i < 10; i++) {
    h += (uint32_t)dst[4 + i] * P5;
    h = ((h << 11) | (h >> 21)) * P1;
  }
  h ^= h >> 15; h *= P2;
  h ^= h >> 13; h *= P3;
  h ^= h >> 16;
  dst[14] = (uint8_t)(h >> 8);
  return 15;
}

// ------------------------------------------------------------------------------------------------- FSE encoding
// Compression tables for the three predefined distributions (RFC 8878 §3.1.1.3.2.2.1), built once on the host.
struct FseCTable {
  uint16_t state[64];      // next-state table (table size <= 64 for the predefined logs 6/5/6)
  int32_t delta_nb[53];    // per symbol: (maxBitsOut << 16) - minStatePlus
  int32_t delta_state[53]; // per symbol: first state slot - count
  uint32_t log;
};
struct PredefCTables {
  FseCTable ll, of, ml;
};

inline void fse_build_ctable(FseCTable* ct, const int16_t* norm, int nsym, int log) {
  const int size = 1 << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
  int cumul[64] = {0};
  uint8_t symtab[64];
  int high = size - 1;
  for (int s = 0; s < nsym; s++) {
    if (norm[s] == -1) { cumul[s + 1] = cumul[s] + 1; symtab[high--] = (uint8_t)s; }
    else cumul[s + 1] = cumul[s] + norm[s];
  }
  int pos = 0;
  for (int s = 0; s < nsym; s++)
    for (int i = 0; i < norm[s]; i++) {
      symtab[pos] = (uint8_t)s;
      do pos = (pos + step) & mask; while (pos > high);
    }
  for (int u = 0; u < size; u++) { const int s = symtab[u]; ct->state[cumul[s]++] = (uint16_t)(size + u); }
  int total = 0;
  for (int s = 0; s < nsym; s++) {
    const int n = norm[s];
    if (n == 0) { ct->delta_nb[s] = ((log + 1) << 16) - (1 << log); ct->delta_state[s] = 0; }
    else if (n == -1 || n == 1) { ct->delta_nb[s] = (log << 16) - (1 << log); ct->delta_state[s] = total - 1; total++; }
    else {
      const int max_bits = log - hibit32((uint32_t)(n - 1));
      const int min_plus = n << max_bits;
      ct->delta_nb[s] = (max_bits << 16) - min_plus;
      ct->delta_state[s] = total - n;
      total += n;
    }
  }
  ct->log = (uint32_t)log;
}

inline void build_predef_ctables(PredefCTables* p) {
  fse_build_ctable(&p->ll, zs::kLLDefault, 36, 6);
  fse_build_ctable(&p->of, zs::kOFDefault, 29, 5);
  fse_build_ctable(&p->ml, zs::kMLDefault, 53, 6);
}

#if defined(__CUDACC__)
__device__ PredefCTables g_predef_c;  // filled by zn_ctx_create
#endif
#if defined(__CUDA_ARCH__)
ZN_D const PredefCTables* predef_ctables() { return &g_predef_c; }
#else
inline const PredefCTables* predef_ctables() {
  static PredefCTables p;
  static bool init = false;
  if (!init) { build_predef_ctables(&p); init = true; }
  return &p;
}
#endif

// forward bit writer (the decoder reads it backwards): bits accumulate LSB first
struct BitWriter {
  uint8_t* p;
  uint64_t acc;
  uint32_t nbits;
  ZN_HD void add(uint32_t v, uint32_t n) {  // n <= 32
    acc |= (uint64_t)(v & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u))) << nbits;
    nbits += n;
  }
  ZN_HD void flush() {  // keeps < 8 bits pending
    while (nbits >= 8) { *p++ = (uint8_t)acc; acc >>= 8; nbits -= 8; }
  }
  ZN_HD void close() {  // end mark
    add(1, 1);
    flush();
    if (nbits) { *p++ = (uint8_t)acc; acc = 0; nbits = 0; }
  }
};

// code of a literal length / match length: the code c with base[c] <= v < base[c] + 2^bits[c]
ZN_HD uint32_t len_code(uint32_t v, const uint32_t* base, const uint8_t* bits, uint32_t direct, uint32_t ncodes) {
  if (v < direct + base[0]) return v - base[0];
  uint32_t c = direct;
  while (c + 1 < ncodes && v >= base[c + 1]) c++;
  (void)bits;
  return c;
}

struct FseCState {
  uint32_t value;
  ZN_HD void init(const FseCTable* ct, uint32_t sym) {
    const int32_t dnb = ct->delta_nb[sym];
    const uint32_t nb = (uint32_t)(dnb + (1 << 15)) >> 16;
    const uint32_t v = (nb << 16) - (uint32_t)dnb;
    value = ct->state[(int32_t)(v >> nb) + ct->delta_state[sym]];
  }
  ZN_HD void encode(BitWriter& bw, const FseCTable* ct, uint32_t sym) {
    const uint32_t nb = (uint32_t)((int32_t)value + ct->delta_nb[sym]) >> 16;
    bw.add(value, nb);
    value = ct->state[(int32_t)(value >> nb) + ct->delta_state[sym]];
  }
  ZN_HD void flush(BitWriter& bw, const FseCTable* ct) { bw.add(value, ct->log); }
};

// packed sequence: ll | ml << 20 | off << 40 (each < 2^20)
ZN_HD uint64_t seq_pack(uint32_t ll, uint32_t ml, uint32_t off) { return (uint64_t)ll | ((uint64_t)ml << 20) | ((uint64_t)off << 40); }

// Sequences section with predefined tables.  One thread.  Returns bytes written, or 0 as soon as the section
// would exceed `limit` bytes (the block is then stored raw); the writer overshoots `limit` by < 16 bytes.
ZN_HD uint32_t zstd_encode_sequences(uint8_t* dst, const uint64_t* seqs, uint32_t nseq, uint32_t limit) {
  uint8_t* p = dst;
  if (nseq < 128) *p++ = (uint8_t)nseq;
  else if (nseq < 0x7F00) { *p++ = (uint8_t)((nseq >> 8) + 128); *p++ = (uint8_t)nseq; }
  else { *p++ = 255; *p++ = (uint8_t)(nseq - 0x7F00); *p++ = (uint8_t)((nseq - 0x7F00) >> 8); }
  if (nseq == 0) return (uint32_t)(p - dst);
  *p++ = 0;  // LL, OF, ML all predefined
  const PredefCTables* ct = predef_ctables();
  BitWriter bw{p, 0, 0};
  FseCState sl, so, sm;
  for (uint32_t k = nseq; k-- > 0;) {
    const uint64_t s = seqs[k];
    const uint32_t ll = (uint32_t)(s & 0xFFFFF), ml = (uint32_t)((s >> 20) & 0xFFFFF), off = (uint32_t)(s >> 40);
    const uint32_t ofb = off + 3u;  // never a repeat code
    const uint32_t oc = (uint32_t)hibit32(ofb);
    const uint32_t lc = len_code(ll, zs::kLLBase, zs::kLLBits, 16, 36);
    const uint32_t mc = len_code(ml, zs::kMLBase, zs::kMLBits, 32, 53);
    if (k == nseq - 1) {
      sm.init(&ct->ml, mc);
      so.init(&ct->of, oc);
      sl.init(&ct->ll, lc);
    } else {
      so.encode(bw, &ct->of, oc);
      sm.encode(bw, &ct->ml, mc);
      sl.encode(bw, &ct->ll, lc);
      bw.flush();
    }
    bw.add(ll - zs::kLLBase[lc], zs::kLLBits[lc]);
    bw.add(ml - zs::kMLBase[mc], zs::kMLBits[mc]);
    bw.flush();
    bw.add(ofb - (1u << oc), oc);
    bw.flush();
    if ((uint32_t)(bw.p - dst) > limit) return 0;
  }
  sm.flush(bw, &ct->ml);
  bw.flush();
  so.flush(bw, &ct->of);
  bw.flush();
  sl.flush(bw, &ct->ll);
  bw.close();
  return (uint32_t)(bw.p - dst);
}

// ------------------------------------------------------------------------------------------------- zstd block
// Compresses slice[bstart .. bstart+n) (n <= 128 KiB) into the payload of one compressed block.
//   stage   kZstdSlot bytes: literals are written from stage+3, the block payload ends up at stage + *payload_off
//   seqs    kZstdMaxSeq packed sequences (global scratch)
//   tab     2^kZstdHashLog x u32 (shared memory)
// Returns the payload size, or 0 when the block should be stored raw.
ZN_HD uint32_t zstd_compress_block(const Warp& w, const uint8_t* slice, uint32_t bstart, uint32_t n, uint8_t* stage,
                                   uint64_t* seqs, uint32_t* tab, uint32_t* payload_off) {
  const uint32_t wbase = bstart > kZstdPrime ? bstart - kZstdPrime : 0u;  // table positions are relative to wbase
  const uint8_t* in = slice + wbase;
  const uint32_t s0 = bstart - wbase, end = s0 + n;
  for (uint32_t i = w.lane; i < (1u << kZstdHashLog); i += w.n) tab[i] = 0xFFFFFFFFu;
  w_sync(w);
  // prime with the tail of the previous block
  for (uint32_t p = w.lane; p + 4u <= s0; p += w.n) tab[hash4(ld32le(in + p), kZstdHashLog)] = p;
  w_sync(w);
  uint8_t* lit = stage + 3;
  uint32_t nlit = 0, nseq = 0, anchor = s0, pos = s0;
  if (n >= 8) {
    const uint32_t mflimit = end - 7;  // a match needs 4 bytes to verify
    while (pos < mflimit && nseq < kZstdMaxSeq) {
      const uint32_t p = pos + w.lane;
      const bool valid = p < mflimit;
      const uint32_t v = valid ? ld32le(in + p) : 0u;
      const uint32_t h = hash4(v, kZstdHashLog);
      const uint32_t cand = valid ? tab[h] : 0xFFFFFFFFu;
      const bool ok = valid && cand < p && ld32le(in + cand) == v;
      const uint32_t m = w_ballot(w, ok);
      const uint32_t f = m ? ffs32(m) - 1u : w.n;
      w_sync(w);
      if (valid && w.lane <= f) tab[h] = p;  // see lz4_compress_block
      w_sync(w);
      if (!m) {
        pos += w.n;
        continue;
      }
      const uint32_t mp = pos + f, mc = w_shfl(w, cand, f);
      const uint32_t ml = 4u + match_extend(w, in + mp + 4, in + mc + 4, end - (mp + 4));
      const uint32_t ll = mp - anchor;
      w_copy(w, lit + nlit, in + anchor, ll);
      nlit += ll;
      if (w.lane == 0) seqs[nseq] = seq_pack(ll, ml, mp - mc);
      nseq++;
      pos = anchor = mp + ml;
    }
  }
  const uint32_t rest = end - anchor;
  w_copy(w, lit + nlit, in + anchor, rest);
  nlit += rest;
  w_sync(w);
  if (nseq == 0) return 0;  // nothing found: raw block
  // literals section header (raw literals), placed right before the literal bytes
  const uint32_t hs = nlit < 32 ? 1u : (nlit < 4096 ? 2u : 3u);
  uint8_t* hp = stage + 3 - hs;
  uint32_t total = 0;
  if (w.lane == 0) {
    if (hs == 1) hp[0] = (uint8_t)(nlit << 3);
    else if (hs == 2) { hp[0] = (uint8_t)(0x04 | ((nlit & 0xF) << 4)); hp[1] = (uint8_t)(nlit >> 4); }
    else { hp[0] = (uint8_t)(0x0C | ((nlit & 0xF) << 4)); hp[1] = (uint8_t)(nlit >> 4); hp[2] = (uint8_t)(nlit >> 12); }
    // the block must beat its raw form, which also keeps the section inside the slot
    if (hs + nlit + 16u < n) {
      const uint32_t ss = zstd_encode_sequences(lit + nlit, seqs, nseq, n - (hs + nlit) - 16u);
      if (ss) total = hs + nlit + ss;
    }
  }
  total = w_shfl(w, total, 0);
  w_sync(w);
  *payload_off = 3 - hs;
  return (total && total < n) ? total : 0u;
}

// Zstandard frame header: single segment, 8-byte content size (13 bytes); empty content uses the 1-byte form.
ZN_HD uint32_t zstd_frame_header(uint8_t* dst, uint64_t content) {
  dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
  dst[4] = 0xE0;
  for (int i = 0; i < 8; i++) dst[5 + i] = (uint8_t)(content >> (8 * i));
  return 13;
}
ZN_HD void zstd_block_header(uint8_t* dst, uint32_t last, uint32_t type, uint32_t size) {
  const uint32_t v = last | (type << 1) | (size << 3);
  dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16);
}

}  // namespace cz
}  // namespace zn
