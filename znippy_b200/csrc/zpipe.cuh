// Device-wide Zstandard decode pipeline ("zpipe"): the entropy-coded path of codec::decompress_into
// (znippy-common/src/codec.rs:67-78, called per row by decompress.rs:156-166 and archive.rs:159-164) for a whole
// batch of index rows at once.  The arithmetic on the reference path lives in openzl-sys-rs 0.2.0 -> zstd (not
// vendored); this is written from RFC 8878.
//
// The one-CTA-per-blob decoders (zstd_decode.cuh, zstd_par.cuh) leave an entropy-coded batch bound by a handful of
// serial bit-stream decoders per SM.  Here every phase runs over EVERY block of EVERY blob of the batch at once:
//
//   walk     one thread per blob hops over frame and block headers and fills a global block table: where each block's
//            literals / Huffman tree / sequence tables are DEFINED (own section, an earlier block for treeless and
//            repeat modes, or the predefined distribution) — which removes the table dependence between blocks — and
//            bump-allocates the blob's sequence records, literal bytes and table sets from batch-wide pools;
//   tables   one warp per compressed block: lane 0 parses the three FSE table descriptions, then all lanes build each
//            decoding table position by position (closed form of FSE's spread walk); 4-byte entries hold next-state
//            base, state bits, extra-bit count and symbol, in global memory;
//   seq      the three interleaved FSE state machines of every block.  Large blobs: in two phases — ONE LANE PER BLOCK
//            runs the state chain alone (tables and bit stream staged in shared memory) and leaves {bit cursor, states}
//            per sequence; then a warp per block decodes the values of 128 sequences per step and gets positions and
//            repeat-offset histories from prefix scans.  Many small blobs: one pass, ONE LANE PER BLOCK, every block of
//            the batch in flight.  Repeat offsets that reach back before the block are kept SYMBOLIC (history entry
//            i minus k), so no block waits for its predecessor.  Output: one 16-byte record per sequence;
//   lit      ONE LANE PER HUFFMAN STREAM (4 per block), the block's decoding table in shared memory;
//   chain    one thread per blob: block output offsets (prefix sum of the regenerated sizes), true repeat-offset
//            history entering every block, frame content sizes;
//   exec     one CTA per blob executes the sequences group by group in shared memory (zpipe_kernels.cuh).
//
// Anything irregular (LZ4 frames, dictionaries, more blocks than budgeted, any malformed field) marks the blob for the
// legacy one-team decoder, which then produces the status exactly as before: the pipeline only has to be right for
// well-formed frames and conservative for everything else.
//
// Everything in this file is scalar code that also compiles for the host (tests/host_emu runs the phases serially and
// checks them against the oracle on the CPU).
#pragma once
#include "zstd_tables.cuh"
#if !defined(__CUDA_ARCH__)
#include <cstring>
#endif

namespace zn {
namespace zp {

constexpr uint32_t kDefPredef = 0xFFFFFFFEu;  // table source: predefined distribution
constexpr uint32_t kDefNone = 0xFFFFFFFFu;    // table source: nothing defined yet in this frame
constexpr uint32_t kSymBit = 0x80000000u;     // offset is symbolic: incoming history entry (v & 3) minus ((v >> 2) & 0x1FFFFFFF)
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

constexpr uint32_t kTabLL = 512, kTabOF = 256, kTabML = 512;      // entries per table (max table logs 9 / 8 / 9)
constexpr uint32_t kTabSet = kTabLL + kTabOF + kTabML;             // entries of one block's table set
constexpr uint32_t kTabOffLL = 0, kTabOffOF = kTabLL, kTabOffML = kTabLL + kTabOF;

// FSE decoding-table entry, 4 bytes: next-state base (10 bits) | state bits << 10 (4) | extra bits << 14 (5) | symbol << 19
// (6).  Everything the bit-position chain needs sits in the entry; the value baseline is looked up by symbol
// (value_base()), off the critical path.  A block's three tables are 5 KiB: 32 streams fit one SM's shared memory.
typedef uint32_t FseD;
ZN_HD uint32_t fd_pack(uint32_t base, uint32_t nbits, uint32_t extra, uint32_t sym) { return base | (nbits << 10) | (extra << 14) | (sym << 19); }
ZN_HD uint32_t fd_base(uint32_t e) { return e & 0x3FFu; }
ZN_HD uint32_t fd_nbits(uint32_t e) { return (e >> 10) & 0xFu; }
ZN_HD uint32_t fd_extra(uint32_t e) { return (e >> 14) & 0x1Fu; }
ZN_HD uint32_t fd_sym(uint32_t e) { return e >> 19; }

// block flags
enum : uint32_t {
  ZB_TYPE_MASK = 3u,      // 0 raw, 1 RLE, 2 compressed
  ZB_LAST = 4u,           // last block of its frame
  ZB_FIRST = 8u,          // first block of its frame
  ZB_LIT_SHIFT = 4,       // bits 4-5: literals type (0 raw, 1 RLE, 2 Huffman, 3 treeless)
  ZB_STREAMS4 = 64u,      // four Huffman streams
  ZB_HAS_FCS = 128u,      // (first block) the frame header carries a content size
  ZB_MODES_SHIFT = 8      // bits 8-15: Symbol_Compression_Modes byte
};

struct ZBlock {
  uint32_t src_off;     // block content, offset from the blob's first byte
  uint32_t len;         // Block_Size (RLE: regenerated size)
  uint32_t flags;
  uint32_t lit_regen;   // regenerated literal bytes
  uint32_t lit_off;     // raw: offset of the literal bytes; RLE: the byte; Huffman: offset of the streams area
  uint32_t lit_clen;    // Huffman: bytes of the streams area (jump table + streams)
  uint32_t huf_desc;    // Huffman: offset of the tree description (type 2 only)
  uint32_t huf_dlen;    //          and its length
  uint32_t huf_def;     // blob-relative index of the block whose tree description this block uses
  uint32_t nseq;
  uint32_t seq_off;     // offset of the first byte after the modes byte
  uint32_t seq_def[3];  // LL / OF / ML: blob-relative index of the block whose table set holds the table, kDefPredef, kDefNone
  uint32_t seq_base;    // first record of this block in the sequence pool
  uint32_t lit_base16;  // Huffman: literal pool offset of the regenerated literals, in 16-byte units
  uint32_t tab_slot;    // table set of this block (it defines at least one table), else kNoSlot
  uint32_t fcs_lo, fcs_hi;  // first block of a frame: Frame_Content_Size
  // ---- tables phase
  uint32_t bits_off;    // offset of the sequence bit stream; ~0u: the table descriptions are malformed
  uint32_t tlogs;       // table logs of this block's set, LL | OF << 8 | ML << 16
  // ---- seq / lit phases
  uint32_t matched;     // bytes produced by the sequences (literals + matches)
  uint32_t lit_used;    // literals consumed by the sequences
  uint32_t rep_fin[3];  // history after the block (possibly symbolic)
  uint32_t st_seq;      // 0 = sequences decoded fine
  uint32_t st_lit;      // Huffman streams still to be decoded fine (0 = literals ready)
  // ---- chain phase
  uint32_t out_base;    // first output byte of the block, relative to the blob's output
  uint32_t rep_in[3];   // true history entering the block
  uint32_t frame_start; // output offset where the block's frame began
  uint32_t pad[3];
};
static_assert(sizeof(ZBlock) == 144, "ZBlock layout");

// per pipeline blob (host fills blob / slot0 / slot_cap, the walk fills the rest)
struct ZBlob {
  uint32_t blob;      // index of the BlobDesc
  uint32_t slot0;     // first ZBlock of this blob
  uint32_t slot_cap;  // ZBlocks available
  uint32_t n_blocks;
  uint32_t state;     // 0: pipeline; != 0: legacy decoder takes the blob
  uint32_t pad[3];
};

// batch-wide bump allocators + capacities
struct ZPools {
  uint32_t seq_used, seq_cap;      // sequence records
  uint32_t lit_used16, lit_cap16;  // Huffman literal bytes / 16
  uint32_t tab_used, tab_cap;      // table sets
  uint32_t comp_used, comp_cap;    // compressed blocks (work list of the tables / seq / lit phases)
};

// what phase 1 of the sequence stage leaves per sequence: bit cursor before it, LL | OF << 9 | ML << 18 states
struct alignas(8) SeqP1 {
  uint32_t cursor, states;
};
// 16-byte sequence record
struct alignas(16) SeqRec16 {
  uint32_t w0, w1, w2, w3;
};
// everything the pipeline kernels share
struct ZArgs {
  const BlobDesc* blobs;
  const uint8_t* blobs_base;
  ZBlob* zb;
  uint32_t nzb;
  ZBlock* blocks;
  ZPools* pools;
  uint32_t* comp_list;
  FseD* tabs;
  SeqRec16* recs;
  uint8_t* lits;
  struct SeqP1* p1;  // phase-1 output of the sequence stage, one entry per record slot
};

ZN_HD SeqRec16 rec_pack(uint32_t out_rel, uint32_t lit_rel, uint32_t ll, uint32_t ml, uint32_t off) {
  SeqRec16 r;
  r.w0 = off;
  r.w1 = out_rel | ((ll & 0x3FFFu) << 18);
  r.w2 = lit_rel | ((ml & 0x3FFFu) << 18);
  r.w3 = (ll >> 14) | ((ml >> 14) << 8);
  return r;
}
ZN_HD uint32_t rec_off(const SeqRec16& r) { return r.w0; }
ZN_HD uint32_t rec_out(const SeqRec16& r) { return r.w1 & 0x3FFFFu; }
ZN_HD uint32_t rec_lit(const SeqRec16& r) { return r.w2 & 0x3FFFFu; }
ZN_HD uint32_t rec_ll(const SeqRec16& r) { return (r.w1 >> 18) | ((r.w3 & 0xFFu) << 14); }
ZN_HD uint32_t rec_ml(const SeqRec16& r) { return (r.w2 >> 18) | ((r.w3 >> 8) << 14); }

// records are written once and read once, much later: stream them past the L2-resident decoding tables
#if defined(__CUDA_ARCH__)
ZN_D void rec_store(SeqRec16* p, const SeqRec16& r) { __stcs(reinterpret_cast<uint4*>(p), make_uint4(r.w0, r.w1, r.w2, r.w3)); }
#else
inline void rec_store(SeqRec16* p, const SeqRec16& r) { *p = r; }
#endif

ZN_HD bool is_sym(uint32_t v) { return (v & kSymBit) != 0; }
ZN_HD uint32_t sym_make(uint32_t i) { return kSymBit | i; }
ZN_HD uint32_t sym_minus1(uint32_t v) { return v + 4u; }
// value of a (possibly symbolic) offset given the true incoming history; 0 = corrupt
ZN_HD uint32_t sym_resolve(uint32_t v, uint32_t r0, uint32_t r1, uint32_t r2) {
  if (!is_sym(v)) return v;
  const uint32_t i = v & 3u, k = (v >> 2) & 0x1FFFFFFFu;
  const uint32_t base = i == 0 ? r0 : (i == 1 ? r1 : r2);
  return base > k ? base - k : 0u;
}

// ------------------------------------------------------------------------------------------------------------ walk
// Literals section header (RFC 8878 §3.1.1.3.1.1).  Fills the literal fields of `b`; *consumed = bytes of the whole
// literals section.  false = malformed.
ZN_HD bool walk_literals(const uint8_t* p, uint32_t len, uint32_t blk_off, ZBlock* b, uint32_t* lit_type, uint32_t* consumed) {
  if (len < 1) return false;
  const uint32_t b0 = p[0], type = b0 & 3, sf = (b0 >> 2) & 3;
  *lit_type = type;
  b->lit_clen = 0; b->huf_desc = 0; b->huf_dlen = 0;
  if (type < 2) {
    uint32_t hdr, regen;
    if ((sf & 1) == 0) { hdr = 1; regen = b0 >> 3; }
    else if (sf == 1) { if (len < 2) return false; hdr = 2; regen = (b0 >> 4) | ((uint32_t)p[1] << 4); }
    else { if (len < 3) return false; hdr = 3; regen = (b0 >> 4) | ((uint32_t)p[1] << 4) | ((uint32_t)p[2] << 12); }
    if (regen > kZstdBlockMax) return false;
    b->lit_regen = regen;
    if (type == 0) {
      if (hdr + regen > len) return false;
      b->lit_off = blk_off + hdr;
      *consumed = hdr + regen;
    } else {
      if (hdr + 1 > len) return false;
      b->lit_off = p[hdr];
      *consumed = hdr + 1;
    }
    return true;
  }
  uint32_t hdr, regen, comp, streams;
  if (sf <= 1) {
    if (len < 3) return false;
    const uint32_t v = ld24le(p);
    hdr = 3; regen = (v >> 4) & 0x3FF; comp = (v >> 14) & 0x3FF; streams = sf == 0 ? 1 : 4;
  } else if (sf == 2) {
    if (len < 4) return false;
    const uint32_t v = ld32le(p);
    hdr = 4; regen = (v >> 4) & 0x3FFF; comp = v >> 18; streams = 4;
  } else {
    if (len < 5) return false;
    const uint64_t v = (uint64_t)ld32le(p) | ((uint64_t)p[4] << 32);
    hdr = 5; regen = (uint32_t)(v >> 4) & 0x3FFFF; comp = (uint32_t)(v >> 22); streams = 4;
  }
  if (regen > kZstdBlockMax || hdr + comp > len) return false;
  b->lit_regen = regen;
  uint32_t desc = 0;
  if (type == 2) {
    if (comp < 1) return false;
    const uint32_t hb = p[hdr];
    desc = hb < 128 ? 1 + hb : 1 + ((hb - 127) + 1) / 2;
    if (desc > comp) return false;
    b->huf_desc = blk_off + hdr;
    b->huf_dlen = desc;
  }
  b->lit_off = blk_off + hdr + desc;
  b->lit_clen = comp - desc;
  if (streams == 4) b->flags |= ZB_STREAMS4;
  *consumed = hdr + comp;
  return true;
}

// Number_of_Sequences at q; returns bytes used (0 = malformed).
ZN_HD uint32_t walk_nseq(const uint8_t* q, uint32_t avail, uint32_t* nseq) {
  if (avail < 1) return 0;
  const uint32_t n = q[0];
  if (n < 128) { *nseq = n; return 1; }
  if (n == 255) { if (avail < 3) return 0; *nseq = (uint32_t)q[1] + ((uint32_t)q[2] << 8) + 0x7F00u; return 3; }
  if (avail < 2) return 0;
  *nseq = ((n - 128) << 8) + q[1];
  return 2;
}

struct WalkNeeds {
  uint32_t nseq, lit16, tabs, comp;
};

// Walks every frame of one blob.  blocks[0 .. slot_cap) receive the block table (allocation fields blob-relative:
// seq_base / lit_base16 / tab_slot count from 0 and are rebased by the caller once the pools have been bumped).
// Returns the number of blocks, or ~0u when the blob has to go to the legacy decoder.
ZN_HD uint32_t walk_blob(const uint8_t* src, uint32_t src_len, uint32_t slot_cap, ZBlock* blocks, WalkNeeds* needs) {
  needs->nseq = needs->lit16 = needs->tabs = needs->comp = 0;
  uint32_t ip = 0, nb = 0;
  if (src_len == 0) return ~0u;
  while (ip < src_len) {
    if (src_len - ip >= 8) {
      const uint32_t magic = ld32le(src + ip);
      if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
        const uint32_t sz = ld32le(src + ip + 4);
        if (sz > src_len - ip - 8) return ~0u;
        ip += 8 + sz;
        continue;
      }
    }
    if (src_len - ip < 5) return ~0u;
    if (ld32le(src + ip) != 0xFD2FB528u) return ~0u;  // LZ4 frame, unknown magic: legacy decoder decides
    const uint32_t fhd = src[ip + 4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
    if (fhd & 0x08) return ~0u;
    const uint32_t checksum = (fhd >> 2) & 1;
    uint32_t hp = ip + 5;
    uint64_t window = 0;
    if (!single) {
      if (src_len < hp + 1) return ~0u;
      const uint32_t wd = src[hp++], e = wd >> 3, m = wd & 7;
      const uint64_t base = 1ull << (10 + e);
      window = base + (base >> 3) * m;
    }
    const uint32_t db = did_flag == 3 ? 4u : did_flag;
    if (src_len < hp + db) return ~0u;
    uint32_t did = 0;
    for (uint32_t i = 0; i < db; i++) did |= (uint32_t)src[hp + i] << (8 * i);
    hp += db;
    if (did != 0) return ~0u;
    const uint32_t fb = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2u : (fcs_flag == 2 ? 4u : 8u));
    if (src_len < hp + fb) return ~0u;
    uint64_t fcs = 0;
    for (uint32_t i = 0; i < fb; i++) fcs |= (uint64_t)src[hp + i] << (8 * i);
    if (fb == 2) fcs += 256;
    hp += fb;
    if (single) window = fcs;
    ip = hp;
    const uint32_t block_max = window < kZstdBlockMax ? (uint32_t)window : kZstdBlockMax;
    uint32_t def_huf = kDefNone, def_seq[3] = {kDefNone, kDefNone, kDefNone};
    bool first = true;
    for (;;) {
      if (src_len - ip < 3) return ~0u;
      if (nb >= slot_cap) return ~0u;
      const uint32_t bh = ld24le(src + ip);
      ip += 3;
      const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
      if (type == 3) return ~0u;
      ZBlock* b = &blocks[nb];
      b->src_off = ip; b->len = bsize;
      b->flags = type | (last ? ZB_LAST : 0u) | (first ? ZB_FIRST : 0u) | (first && fb ? ZB_HAS_FCS : 0u);
      b->fcs_lo = (uint32_t)fcs; b->fcs_hi = (uint32_t)(fcs >> 32);
      b->lit_regen = 0; b->lit_off = 0; b->lit_clen = 0; b->huf_desc = 0; b->huf_dlen = 0; b->huf_def = kDefNone;
      b->nseq = 0; b->seq_off = 0; b->seq_def[0] = b->seq_def[1] = b->seq_def[2] = kDefNone;
      b->seq_base = needs->nseq; b->lit_base16 = needs->lit16; b->tab_slot = kNoSlot;
      b->bits_off = ~0u; b->tlogs = 0;
      b->matched = 0; b->lit_used = 0;
      b->rep_fin[0] = sym_make(0); b->rep_fin[1] = sym_make(1); b->rep_fin[2] = sym_make(2);
      b->st_seq = 0; b->st_lit = 0;
      b->out_base = 0; b->rep_in[0] = b->rep_in[1] = b->rep_in[2] = 0; b->frame_start = 0;
      first = false;
      if (type == 1) {
        if (src_len - ip < 1) return ~0u;
        ip += 1;
      } else {
        if (bsize > src_len - ip) return ~0u;
        if (type == 2) {
          if (bsize > block_max || bsize < 2) return ~0u;
          const uint8_t* p = src + ip;
          uint32_t lit_type, consumed;
          if (!walk_literals(p, bsize, ip, b, &lit_type, &consumed)) return ~0u;
          b->flags |= lit_type << ZB_LIT_SHIFT;
          if (lit_type == 2) def_huf = nb;
          if (lit_type >= 2) {
            if (def_huf == kDefNone) return ~0u;
            b->huf_def = def_huf;
            b->st_lit = (b->flags & ZB_STREAMS4) ? 4u : 1u;
            needs->lit16 += (b->lit_regen + 15u) / 16u + 1u;  // + 1: word stores may touch the next aligned word
          }
          uint32_t nseq;
          const uint32_t used = walk_nseq(p + consumed, bsize - consumed, &nseq);
          if (!used) return ~0u;
          b->nseq = nseq;
          if (nseq == 0) {
            if (consumed + used != bsize) return ~0u;
          } else {
            if (consumed + used >= bsize) return ~0u;
            if (nseq > kZstdBlockMax / 3 + 1) return ~0u;
            const uint32_t modes = p[consumed + used];
            if (modes & 3) return ~0u;
            b->flags |= modes << ZB_MODES_SHIFT;
            b->seq_off = ip + consumed + used + 1;
            bool defines = false;
            for (int k = 0; k < 3; k++) {
              const uint32_t m = (modes >> (6 - 2 * k)) & 3;
              if (m == 0) def_seq[k] = kDefPredef;
              else if (m != 3) { def_seq[k] = nb; defines = true; }
              else if (def_seq[k] == kDefNone) return ~0u;
              b->seq_def[k] = def_seq[k];
            }
            if (defines) b->tab_slot = needs->tabs++;
            b->st_seq = 1;
            needs->nseq += nseq;
          }
          needs->comp++;
        }
        ip += bsize;
      }
      nb++;
      if (last) break;
    }
    if (checksum) {  // XXH64 content checksum: present but not verified (blake3 of the content is)
      if (src_len - ip < 4) return ~0u;
      ip += 4;
    }
  }
  return nb;
}

// ---------------------------------------------------------------------------------------------------------- tables
ZN_HD void table_params(int k, int* max_log, int* max_sym) {
  *max_log = k == 1 ? 8 : 9;
  *max_sym = k == 0 ? 35 : (k == 1 ? 31 : 52);
}

// value baseline and extra-bit count of symbol s of table kind k (0 LL, 1 OF, 2 ML); false = symbol not decodable here
ZN_HD bool sym_value(int k, uint32_t s, uint32_t* base, uint32_t* extra) {
  if (k == 0) { if (s > 35) return false; *base = zs::kLLBase[s]; *extra = zs::kLLBits[s]; }
  else if (k == 1) { if (s > 30) return false; *base = 1u << s; *extra = s; }
  else { if (s > 52) return false; *base = zs::kMLBase[s]; *extra = zs::kMLBits[s]; }
  return true;
}

// Decoding table from normalized counts (same spreading as zs::fse_build).  `next` = scratch for nsym uint16.
ZN_HD bool build_fat_table(FseD* t, int k, const int16_t* norm, int nsym, int log, uint16_t* next) {
  const int size = 1 << log;
  int high = size - 1;
  for (int s = 0; s < nsym; s++) {
    if (norm[s] == -1) { t[high--] = (uint32_t)s; next[s] = 1; }
    else next[s] = (uint16_t)norm[s];
  }
  const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
  int pos = 0;
  for (int s = 0; s < nsym; s++)
    for (int i = 0; i < norm[s]; i++) {
      t[pos] = (uint32_t)s;
      do pos = (pos + step) & mask; while (pos > high);
    }
  for (int u = 0; u < size; u++) {
    const uint32_t s = t[u];
    const uint32_t ns = next[s]++;
    const uint32_t nb = (uint32_t)(log - hibit32(ns));
    uint32_t base, extra;
    if (!sym_value(k, s, &base, &extra)) return false;
    t[u] = fd_pack((ns << nb) - (uint32_t)size, nb, extra, s);
  }
  return true;
}

// ---- the same table, position by position (what lets a whole warp build it: k_ztables).  FSE spreads the occurrences of
// the symbols over the table by walking pos = (pos + step) & mask from 0 and skipping the positions above `high` (those
// hold the symbols of probability "less than one", one each, from the top down).  step is odd, so the unfiltered walk
// visits position p at time t(p) = p * step^-1 mod size; the filtered walk reaches it after t(p) minus the number of
// skipped positions visited earlier.  That rank j says which occurrence lands on p, and the inclusive prefix sums of
// the counts say whose occurrence it is.  The state's number within its symbol (what FSE calls `next`) is then the
// count of lower positions holding the same symbol.
ZN_HD uint32_t fat_step_inv(uint32_t size) {  // inverse of the spread step modulo size (a power of two, >= 32)
  const uint32_t step = (size >> 1) + (size >> 3) + 3;
  uint32_t x = step;  // correct to 3 bits for any odd number; every Newton step doubles that
  x *= 2u - step * x;
  x *= 2u - step * x;
  x *= 2u - step * x;
  return x & (size - 1);
}
// symbol at position p.  cum[s] = inclusive prefix sum of the positive counts (0xFFFF from nsym up to 63), lowsym[i] = the
// i-th symbol of count -1
ZN_HD uint32_t fat_symbol_at(uint32_t p, uint32_t size, uint32_t n_low, uint32_t inv, const uint16_t* cum, const uint16_t* lowsym) {
  const uint32_t mask = size - 1;
  if (p + n_low >= size) return lowsym[size - 1 - p];  // above `high` (n_low may be the whole table)
  const uint32_t t = (p * inv) & mask;
  uint32_t skipped = 0;
  for (uint32_t i = 0; i < n_low; i++) skipped += (((size - 1 - i) * inv) & mask) < t ? 1u : 0u;
  const uint32_t j = t - skipped;
  uint32_t s = 0;  // smallest s with cum[s] > j
#pragma unroll
  for (uint32_t w = 32; w; w >>= 1)
    if (cum[s + w - 1] <= j) s += w;
  return s;
}
ZN_HD bool fat_entry(int k, uint32_t sym, uint32_t ns, int log, FseD* e) {
  const uint32_t nb = (uint32_t)(log - hibit32(ns));
  uint32_t base, extra;
  if (!sym_value(k, sym, &base, &extra)) return false;
  *e = fd_pack((ns << nb) - (1u << log), nb, extra, sym);
  return true;
}
#if !defined(__CUDA_ARCH__)
// host mirror of the warp's build (k_ztables): same helpers, positions in ascending order
inline bool build_fat_table_by_position(FseD* t, int k, const int16_t* norm, int nsym, int log) {
  const uint32_t size = 1u << log, inv = fat_step_inv(size);
  uint16_t cum[64], lowsym[64], run[64];
  uint32_t acc = 0, n_low = 0;
  for (int s = 0; s < 64; s++) {
    run[s] = 0;
    if (s < nsym && norm[s] > 0) acc += (uint32_t)norm[s];
    if (s < nsym && norm[s] == -1) lowsym[n_low++] = (uint16_t)s;
    cum[s] = s < nsym ? (uint16_t)acc : (uint16_t)0xFFFF;
  }
  for (uint32_t p = 0; p < size; p++) {
    const uint32_t sym = fat_symbol_at(p, size, n_low, inv, cum, lowsym);
    if (sym >= (uint32_t)nsym) return false;
    const uint32_t ns = (norm[sym] == -1 ? 1u : (uint32_t)norm[sym]) + run[sym]++;
    if (!fat_entry(k, sym, ns, log, &t[p])) return false;
  }
  return true;
}
#endif

// Step 1 (one thread): parses the table descriptions of block b in order.  For mode-2 tables the normalized counts go
// to norm[k][0..64), nsym[k], log[k]; mode-1 tables are written directly (one entry).  Sets *bits_off.  false = malformed.
ZN_HD bool parse_table_descs(const uint8_t* src, const ZBlock* b, FseD* set, int16_t (*norm)[64], int* nsym, int* log,
                             uint32_t* bits_off) {
  const uint32_t modes = (b->flags >> ZB_MODES_SHIFT) & 0xFFu;
  const uint8_t* q = src + b->seq_off;
  const uint8_t* end = src + b->src_off + b->len;
  const uint32_t offs[3] = {kTabOffLL, kTabOffOF, kTabOffML};
  for (int k = 0; k < 3; k++) {
    const uint32_t m = (modes >> (6 - 2 * k)) & 3;
    int max_log, max_sym;
    table_params(k, &max_log, &max_sym);
    log[k] = -1;
    if (m == 1) {
      if (q >= end) return false;
      const uint32_t sym = *q++;
      if ((int)sym > max_sym) return false;
      uint32_t base, extra;
      if (!sym_value(k, sym, &base, &extra)) return false;
      set[offs[k]] = fd_pack(0, 0, extra, sym);
      log[k] = 0;
      nsym[k] = 0;  // nothing to build
    } else if (m == 2) {
      if (q >= end) return false;
      const int used = zs::fse_read_ncount(q, (uint32_t)(end - q), max_log, max_sym, norm[k], &log[k], &nsym[k]);
      if (used < 0) return false;
      q += used;
    }
  }
  if (q >= end) return false;  // the bit stream must hold at least its end-marker byte
  *bits_off = (uint32_t)(q - src);
  return true;
}

// ------------------------------------------------------------------------------------------------------------- seq
// Per-block table view used by the sequence decoder: Acc::ld(k, i) returns entry i of table k (0 LL, 1 OF, 2 ML) and
// Acc::base(k, sym) the value baseline of a symbol — on the device both are shared-memory reads of the lane's own
// table copy and of the baseline LUTs, on the host plain arrays.
struct HostTabs {
  const FseD* t[3];
  ZN_HD uint32_t ld(int k, uint32_t i) const { return t[k][i]; }
  ZN_HD void ld1(int k, uint32_t i, uint32_t* nb, uint32_t* tot) const {  // phase-1 view of the same entry (two-phase form below)
    const FseD e = t[k][i];
    *nb = e & 0x3FFFu; *tot = fd_nbits(e) + fd_extra(e);
  }
  ZN_HD uint32_t base(int k, uint32_t s) const { return k == 0 ? zs::kLLBase[s] : zs::kMLBase[s]; }
};

#if defined(__CUDA_ARCH__)
ZN_D uint32_t shl_c(uint32_t w, uint32_t n) { return __funnelshift_lc(0u, w, n); }  // w << n, 0 for n >= 32
ZN_D uint32_t shr_c(uint32_t w, uint32_t n) { return __funnelshift_rc(w, 0u, n); }  // w >> n, 0 for n >= 32
#else
inline uint32_t shl_c(uint32_t w, uint32_t n) { return n >= 32 ? 0u : w << n; }
inline uint32_t shr_c(uint32_t w, uint32_t n) { return n >= 32 ? 0u : w >> n; }
#endif

// One thread decodes every sequence of block b into rec[0 .. nseq).  Fills matched / lit_used / rep_fin and returns
// true, or false when anything is off (the blob then goes to the legacy decoder).
//
// Bit reading: the three table entries of a sequence say how many bits each of its six fields takes before any bit is
// read, so the common case (<= 32 bits in all) is ONE refill of the 64-bit window, six shift-pairs out of its top word
// and one skip; longer sequences (huge offsets) take the field-by-field path.
template <class Acc>
ZN_HD bool decode_sequences(const uint8_t* src, const ZBlock* b, const Acc& tabs, const uint32_t* logs, SeqRec16* rec,
                            uint32_t* matched, uint32_t* lit_used, uint32_t* rep_fin) {
  const uint32_t nseq = b->nseq, lit_len = b->lit_regen;
  const uint32_t end = b->src_off + b->len;
  if (b->bits_off >= end) return false;
  BackBits bb;
  if (!bb.init(src + b->bits_off, end - b->bits_off)) return false;
  bb.refill();
  uint32_t sl = bb.read(logs[0]), so = bb.read(logs[1]), sm = bb.read(logs[2]);
  if (bb.bits_left < 0) return false;
  uint32_t h0 = sym_make(0), h1 = sym_make(1), h2 = sym_make(2);
  uint32_t lit_pos = 0, out_pos = 0;
  uint32_t bad = 0;  // sticky: no early exit from the lane-per-block loop (a diverged warp pays for both sides)
  for (uint32_t i = 0; i < nseq; i++) {
    const uint32_t el = tabs.ld(0, sl), eo = tabs.ld(1, so), em = tabs.ld(2, sm);
    const uint32_t ofx = fd_extra(eo), mlx = fd_extra(em), llx = fd_extra(el);
    const bool lastq = i + 1 == nseq;
    const uint32_t nl = lastq ? 0u : fd_nbits(el), nm = lastq ? 0u : fd_nbits(em), no = lastq ? 0u : fd_nbits(eo);
    // the offset's extra bits (up to 31, ~20 for a far match) are read on their own; the other five fields (two length
    // extras + three state updates, typically ~20 bits) come out of ONE 32-bit window
    bb.refill();
    const uint32_t ofv = bb.read(ofx);
    bb.refill();
    const uint32_t a1 = mlx, a2 = a1 + llx, a3 = a2 + nl, a4 = a3 + nm, rest = a4 + no;
    uint32_t mlv, llv, vl, vm, vo;
    if (rest <= 32) {
      const uint32_t w = (uint32_t)(bb.win >> 32);
      mlv = shr_c(w, 32 - mlx);
      llv = shr_c(shl_c(w, a1), 32 - llx);
      vl = shr_c(shl_c(w, a2), 32 - nl);
      vm = shr_c(shl_c(w, a3), 32 - nm);
      vo = shr_c(shl_c(w, a4), 32 - no);
      bb.skip(rest);
    } else {
      mlv = bb.read(mlx);
      llv = bb.read(llx);
      bb.refill();
      vl = bb.read(nl);
      vm = bb.read(nm);
      vo = bb.read(no);
    }
    const uint32_t oc = fd_sym(eo);
    const uint32_t ov = (1u << oc) + ofv;
    const uint32_t ml = tabs.base(2, fd_sym(em)) + mlv;
    const uint32_t ll = tabs.base(0, fd_sym(el)) + llv;
    sl = fd_base(el) + vl;
    sm = fd_base(em) + vm;
    so = fd_base(eo) + vo;
    uint32_t offset;
    if (ov > 3) {
      offset = ov - 3;
      h2 = h1; h1 = h0; h0 = offset;
    } else {
      const uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);
      if (idx == 0) offset = h0;
      else {
        if (idx == 3) {
          if (is_sym(h0)) offset = sym_minus1(h0);
          else { offset = h0 - 1; bad |= offset == 0; }
        } else offset = idx == 1 ? h1 : h2;
        if (idx != 1) h2 = h1;
        h1 = h0;
        h0 = offset;
      }
    }
    // (an over-read shows up as bits_left < 0 at the end: the reader returns zeros below the stream start, never faults)
    bad |= (ll > lit_len - lit_pos) | (out_pos + ll + ml > kZstdBlockMax);
    rec_store(rec + i, rec_pack(out_pos & 0x3FFFFu, lit_pos & 0x3FFFFu, ll & 0x3FFFFu, ml & 0x3FFFFu, offset));
    lit_pos += ll;
    out_pos += ll + ml;
  }
  if (bad) return false;
  if (bb.bits_left != 0) return false;
  *matched = out_pos;
  *lit_used = lit_pos;
  rep_fin[0] = h0; rep_fin[1] = h1; rep_fin[2] = h2;
  return true;
}

// ----------------------------------------------------------------------------------------------- seq, two-phase form
// decode_sequences() above keeps ~200 instructions per sequence on ONE lane's in-order chain: the only true serial
// dependence of the sequence section is state -> table entry -> bit count -> next state, everything else (extra-bit
// values, baselines, repeat offsets, positions, record packing) merely rides along and stalls the chain whenever one of
// its own loads is late.  The two-phase form splits them:
//   phase 1 (seq_phase1, lane per block)   the state chain alone: per sequence three table reads, one add, one 32-bit
//            field out of the stream at a computed bit position, three state updates.  It leaves {bit cursor, three
//            states} — 8 bytes, densely packed so that four sequences complete a 32-byte sector (a half-written sector
//            costs a read-modify-write in ECC memory) — in the p1 pool;
//   phase 2 (k_zseq2 on the device, seq_phase2_host here: a warp per block, 128 consecutive sequences per step, four
//            per lane)  with cursor and states known every sequence decodes independently and every access
//            is coalesced: values out of the stream, then positions as a prefix sum and repeat-offset histories as a
//            prefix scan — one sequence's effect on the history is a map "slot i minus k | fixed value" per entry, and
//            such maps compose (sym_compose) — then the final 16-byte records, bit-identical to what decode_sequences()
//            writes (tests/host_emu compares the two on every block).

// Random-access view of a backward bit stream: bit positions count from the aligned word that holds the first stream
// byte; bits below the stream (and any word outside it) read as zero, so over-reads are harmless and show up as a
// final cursor != bias.
struct SeqBits {
  const uint32_t* wbase;
  uint32_t lowmask;
  int32_t top;   // last aligned word holding stream bytes
  int32_t bias;  // cursor when every bit has been read (bits of word 0 that precede the stream)
  int32_t c0;    // cursor at the start: the end marker's position
  ZN_HD bool init(const uint8_t* p, uint32_t len) {
    if (len == 0) return false;
    const uint32_t last = p[len - 1];
    if (last == 0) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), s_al = a & ~(uintptr_t)3;
    wbase = reinterpret_cast<const uint32_t*>(s_al);
    bias = (int32_t)(a & 3) * 8;
    lowmask = 0xFFFFFFFFu << bias;
    top = (int32_t)((((a + len - 1) & ~(uintptr_t)3) - s_al) >> 2);
    c0 = bias + (int32_t)(len - 1) * 8 + hibit32(last);
    return true;
  }
  ZN_HD uint32_t word(int32_t i) const {
    if (i < 0 || i > top) return 0u;
#if defined(__CUDA_ARCH__)
    const uint32_t w = __ldg(wbase + i);  // the blobs are read-only for the whole batch
#else
    const uint32_t w = wbase[i];
#endif
    return i == 0 ? (w & lowmask) : w;
  }
  ZN_HD void advance(int32_t) {}  // (the device's phase-1 reader stages the stream ahead of the cursor here: RingBits)
  ZN_HD uint32_t rel(int32_t c) const { return (uint32_t)c; }
  ZN_HD uint32_t bits32(int32_t lo) const {  // the 32 bits at positions lo .. lo + 31 (lo may be negative)
    const int32_t wi = lo >> 5;
    return funnel_r(word(wi), word(wi + 1), (uint32_t)lo & 31u);
  }
};
ZN_HD uint32_t lowbits(uint32_t v, uint32_t n) { return v & ((1u << n) - 1u); }  // n <= 31

// Phase-1 view of the lane's bit stream (same interface as SeqBits): the stream is staged through a 64-word ring in
// shared memory by asynchronous 16-byte copies issued 13 units (208 bytes, ~55 sequences) ahead of the cursor.  With the
// words read straight from global memory a warp waits, at EVERY sequence, for whichever of its 32 lanes happens to touch
// a new sector (each lane does so every ~9 sequences: 1 - 0.9^32 = 97 % of the steps see a miss); staged, the chain only
// ever reads shared memory.  Bit positions count from the 16-byte unit that holds the first stream byte.  A unit that
// holds a valid byte lies inside the allocation (allocations are 16-byte granular), like the aligned words elsewhere.
#if defined(__CUDACC__)
#define ZN_RING ZN_D
#else
#define ZN_RING inline
#endif
struct RingBits {
#if defined(__CUDACC__)
  uint32_t ring_s;       // shared address of the lane's ring (kRingBytes)
#else
  uint32_t ring_h[64];   // host emulation: the copies are synchronous
#endif
  const uint8_t* g16;    // unit 0
  uint32_t lowmask;
  int32_t first_w, top;  // first / last word holding stream bytes
  int32_t bias, c0;
  int32_t u_have;        // lowest unit asked for so far
  // a lane's own copy groups complete in order, so "at most 12 pending" covers everything 13 or more units below the
  // cursor whether the hardware counts groups per thread (PTX) or per warp (then ~12 of the warp's steps, still far
  // more than a DRAM round trip)
  static constexpr int32_t kAhead = 13;
  static constexpr uint32_t kRingBytes = 256;
#if defined(__CUDACC__)
  ZN_D void fetch_unit(int32_t u) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_s + ((uint32_t)u & 15u) * 16u), "l"(__cvta_generic_to_global(g16 + (size_t)u * 16)) : "memory");
  }
  ZN_D void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  ZN_D void wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  ZN_D void wait_12() { asm volatile("cp.async.wait_group 12;" ::: "memory"); }
  ZN_D uint32_t ring_ld(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ring_s + i * 4u)); return v; }
#else
  void fetch_unit(int32_t u) { memcpy(ring_h + ((uint32_t)u & 15u) * 4u, g16 + (size_t)u * 16, 16); }
  void commit() {}
  void wait_all() {}
  void wait_12() {}
  uint32_t ring_ld(uint32_t i) const { return ring_h[i]; }
#endif
  ZN_RING bool init(const uint8_t* p, uint32_t len) {
    if (len == 0) return false;
    const uint32_t last = p[len - 1];
    if (last == 0) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), s16 = a & ~(uintptr_t)15;
    g16 = reinterpret_cast<const uint8_t*>(s16);
    first_w = (int32_t)((a - s16) >> 2);
    bias = (int32_t)(a - s16) * 8;
    lowmask = 0xFFFFFFFFu << ((a & 3) * 8);
    top = (int32_t)(((a + len - 1) - s16) >> 2);
    c0 = bias + (int32_t)(len - 1) * 8 + hibit32(last);
    const int32_t u_top = top >> 2;
    u_have = u_top > kAhead + 1 ? u_top - (kAhead + 1) : 0;
    for (int32_t u = u_top; u >= u_have; u--) fetch_unit(u);
    commit();
    wait_all();
    return true;
  }
  ZN_RING uint32_t rel(int32_t c) const { return (uint32_t)(c - first_w * 32); }  // cursor as SeqBits (phase 2) counts it
  // the cursor is about to be `lo`: ask for the unit kAhead below it, make sure everything at and above it has arrived
  ZN_RING void advance(int32_t lo) {
    if (u_have > 0 && (lo >> 7) < u_have + kAhead) {
      u_have--;
      fetch_unit(u_have);
      commit();
      if (u_have == 0) wait_all();  // nothing will be asked for any more: from here on every unit is in
    }
    wait_12();  // groups = units in descending order: all but the 12 lowest are in
  }
  ZN_RING uint32_t word(int32_t i) const {
    if (i < first_w || i > top) return 0u;
    const uint32_t w = ring_ld((uint32_t)i & 63u);
    return i == first_w ? (w & lowmask) : w;
  }
  ZN_RING uint32_t bits32(int32_t lo) const {
    const int32_t wi = lo >> 5;
    return funnel_r(word(wi), word(wi + 1), (uint32_t)lo & 31u);
  }
};
#undef ZN_RING

// Phase-1 view of a decoding-table entry, Acc::ld1(k, state, &nb, &tot): nb = next-state base | state bits << 10 (the low
// 14 bits of the FseD entry), tot = state bits + extra bits (what the sequence takes out of the stream for this table).
// In shared memory that is 3 bytes per entry (u16 + u8), which is what lets two rounds of lanes cover a 2 GiB batch.
ZN_HD void p1_store(SeqP1* r, uint32_t cursor, uint32_t states) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<uint2*>(r) = make_uint2(cursor, states);
#else
  r->cursor = cursor; r->states = states;
#endif
}

// Phase 1.  Leaves {cursor before the sequence, LL | OF << 9 | ML << 18 states} in rec[i].  false = the stream does not
// end exactly where the sequences do.
template <class Bits, class Acc>
ZN_HD bool seq_phase1(Bits& sb, const uint8_t* src, const ZBlock* b, const Acc& tabs, const uint32_t* logs, SeqP1* rec) {
  const uint32_t nseq = b->nseq;
  const uint32_t end = b->src_off + b->len;
  if (b->bits_off >= end) return false;
  if (!sb.init(src + b->bits_off, end - b->bits_off)) return false;
  int32_t c = sb.c0;
  c -= (int32_t)logs[0];
  uint32_t sl = lowbits(sb.bits32(c), logs[0]);
  c -= (int32_t)logs[1];
  uint32_t so = lowbits(sb.bits32(c), logs[1]);
  c -= (int32_t)logs[2];
  uint32_t sm = lowbits(sb.bits32(c), logs[2]);
  if (c < sb.bias) return false;
  uint32_t el, eo, em, tl, to, tm;
  for (uint32_t i = 0; i + 1 < nseq; i++) {
    tabs.ld1(0, sl, &el, &tl); tabs.ld1(1, so, &eo, &to); tabs.ld1(2, sm, &em, &tm);
    p1_store(rec + i, sb.rel(c), sl | (so << 9) | (sm << 18));
    c -= (int32_t)(tl + to + tm);
    sb.advance(c);
    const uint32_t f = sb.bits32(c);  // from the bottom: OF state bits, ML state bits, LL state bits (read in the opposite order)
    const uint32_t no = eo >> 10, nm = em >> 10, nl = el >> 10;
    so = (eo & 0x3FFu) + lowbits(f, no);
    sm = (em & 0x3FFu) + lowbits(f >> no, nm);
    sl = (el & 0x3FFu) + lowbits(f >> (no + nm), nl);
  }
  tabs.ld1(0, sl, &el, &tl); tabs.ld1(1, so, &eo, &to); tabs.ld1(2, sm, &em, &tm);
  p1_store(rec + (nseq - 1), sb.rel(c), sl | (so << 9) | (sm << 18));
  c -= (int32_t)(tl + to + tm);  // the last sequence has no state update
  c += (int32_t)((el >> 10) + (eo >> 10) + (em >> 10));
  return c == sb.bias;
}

// Positions and repeat-offset history: at the start of a run (after the scan) or a run's own effect (before it).
struct RunSum {
  uint32_t lit, out;    // literals consumed / bytes produced; sums saturate at kRunSat
  uint32_t h0, h1, h2;  // history, possibly symbolic (relative to whatever precedes)
};
constexpr uint32_t kRunSat = 1u << 30;
ZN_HD uint32_t sat_add(uint32_t a, uint32_t b) { const uint32_t s = a + b; return s < kRunSat ? s : kRunSat; }  // a, b <= 2^30

// v as left by a run that starts with history (r0, r1, r2) — themselves symbolic or fixed; 0 = invalid, as in sym_resolve
ZN_HD uint32_t sym_compose(uint32_t v, uint32_t r0, uint32_t r1, uint32_t r2) {
  if (!is_sym(v)) return v;
  const uint32_t i = v & 3u, k = (v >> 2) & 0x1FFFFFFFu;
  const uint32_t base = i == 0 ? r0 : (i == 1 ? r1 : r2);
  if (is_sym(base)) return base + (k << 2);
  return base > k ? base - k : 0u;
}

// One sequence's effect on the history (RFC 8878 §3.1.1.5); returns its offset.  *zero is set when a fixed offset
// reaches 0 (corrupt) — only meaningful when the history is relative to the block start.
ZN_HD uint32_t rep_step(uint32_t ov, uint32_t ll, uint32_t& h0, uint32_t& h1, uint32_t& h2, uint32_t* zero) {
  uint32_t offset;
  if (ov > 3) {
    offset = ov - 3;
    h2 = h1; h1 = h0; h0 = offset;
  } else {
    const uint32_t idx = ov - 1 + (ll == 0 ? 1u : 0u);
    if (idx == 0) offset = h0;
    else {
      if (idx == 3) {
        if (is_sym(h0)) offset = sym_minus1(h0);
        else { offset = h0 - 1; *zero |= offset == 0; }
      } else offset = idx == 1 ? h1 : h2;
      if (idx != 1) h2 = h1;
      h1 = h0;
      h0 = offset;
    }
  }
  return offset;
}

// Phase 2, values of one sequence from its phase-1 slot.  Acc::ld(k, state) = FseD entry, Acc::base(k, sym) = baseline.
template <class Acc>
ZN_HD void seq_values(const SeqBits& sb, const Acc& tabs, const SeqP1* slot, uint32_t* ll, uint32_t* ml, uint32_t* ov) {
#if defined(__CUDA_ARCH__)
  const uint2 p = *reinterpret_cast<const uint2*>(slot);
  const uint32_t w0 = p.x, w1 = p.y;
#else
  const uint32_t w0 = slot->cursor, w1 = slot->states;
#endif
  const uint32_t el = tabs.ld(0, w1 & 511u), eo = tabs.ld(1, (w1 >> 9) & 511u), em = tabs.ld(2, w1 >> 18);
  const uint32_t ofx = fd_extra(eo), mlx = fd_extra(em), llx = fd_extra(el);
  const int32_t c1 = (int32_t)w0 - (int32_t)ofx;              // offset extra bits: [c1, c1 + ofx)
  const int32_t c2 = c1 - (int32_t)(mlx + llx);                // then ML extras above LL extras: [c2, c2 + llx + mlx)
  const uint32_t ofv = lowbits(sb.bits32(c1), ofx);
  const uint32_t g = sb.bits32(c2);
  *ov = (1u << fd_sym(eo)) + ofv;
  *ml = tabs.base(2, fd_sym(em)) + lowbits(g >> llx, mlx);
  *ll = tabs.base(0, fd_sym(el)) + lowbits(g, llx);
}

constexpr uint32_t kSeqPerLane = 4;  // phase 2: consecutive sequences per lane and step (one scan per 128 sequences)

#if !defined(__CUDA_ARCH__)
// Phase 2 as k_zseq2 runs it, with the warp's 32 lanes as arrays: steps of 32 x kSeqPerLane consecutive sequences; a
// lane walks its own kSeqPerLane sequences serially (values, its own effect on the history, local positions), then
// inclusive Hillis-Steele scans over the lanes give the two sums and the history maps; the state after a step is carried
// into the next.  Returns false when a sequence breaks a block limit (the conditions decode_sequences() checks).
template <class Acc>
inline bool seq_phase2_host(const SeqBits& sb, const Acc& tabs, const SeqP1* p1, SeqRec16* rec, uint32_t nseq, uint32_t lit_len, RunSum* fin) {
  RunSum carry;
  carry.lit = 0; carry.out = 0; carry.h0 = sym_make(0); carry.h1 = sym_make(1); carry.h2 = sym_make(2);
  uint32_t bad = 0;
  const uint32_t K = kSeqPerLane;
  for (uint32_t base = 0; base < nseq; base += 32 * K) {
    uint32_t ll[32][K], ml[32][K], offx[32][K], pl[32][K], po[32][K], sl[32], so[32], tl[32], to[32], m0[32], m1[32], m2[32];
    for (uint32_t l = 0; l < 32; l++) {
      m0[l] = sym_make(0); m1[l] = sym_make(1); m2[l] = sym_make(2);
      sl[l] = so[l] = 0;
      for (uint32_t j = 0; j < K; j++) {
        const uint32_t i = base + l * K + j;
        ll[l][j] = ml[l][j] = 0; offx[l][j] = 0;
        if (i < nseq) {
          uint32_t ov, z = 0;
          seq_values(sb, tabs, p1 + i, &ll[l][j], &ml[l][j], &ov);
          offx[l][j] = rep_step(ov, ll[l][j], m0[l], m1[l], m2[l], &z);
        }
        pl[l][j] = sl[l]; po[l][j] = so[l];
        sl[l] += ll[l][j]; so[l] += ll[l][j] + ml[l][j];
      }
      tl[l] = sl[l]; to[l] = so[l];
    }
    for (uint32_t d = 1; d < 32; d <<= 1)
      for (uint32_t l = 31; l >= d; l--) {  // descending: lane l - d still holds the previous step's value
        sl[l] += sl[l - d]; so[l] += so[l - d];
        const uint32_t t0 = m0[l - d], t1 = m1[l - d], t2 = m2[l - d];
        const uint32_t n0 = sym_compose(m0[l], t0, t1, t2), n1 = sym_compose(m1[l], t0, t1, t2), n2 = sym_compose(m2[l], t0, t1, t2);
        m0[l] = n0; m1[l] = n1; m2[l] = n2;
      }
    for (uint32_t l = 0; l < 32; l++) {
      const uint32_t e0 = l ? m0[l - 1] : sym_make(0), e1 = l ? m1[l - 1] : sym_make(1), e2 = l ? m2[l - 1] : sym_make(2);
      const uint32_t lit_base = carry.lit + (sl[l] - tl[l]), out_base = carry.out + (so[l] - to[l]);
      for (uint32_t j = 0; j < K; j++) {
        const uint32_t i = base + l * K + j;
        if (i >= nseq) break;
        const uint32_t offset = sym_compose(sym_compose(offx[l][j], e0, e1, e2), carry.h0, carry.h1, carry.h2);
        const uint32_t lit_pos = lit_base + pl[l][j], out_pos = out_base + po[l][j];
        bad |= (offset == 0) | (lit_pos + ll[l][j] > lit_len) | (out_pos + ll[l][j] + ml[l][j] > kZstdBlockMax);
        rec_store(rec + i, rec_pack(out_pos & 0x3FFFFu, lit_pos & 0x3FFFFu, ll[l][j] & 0x3FFFFu, ml[l][j] & 0x3FFFFu, offset));
      }
    }
    const uint32_t n0 = sym_compose(m0[31], carry.h0, carry.h1, carry.h2), n1 = sym_compose(m1[31], carry.h0, carry.h1, carry.h2),
                   n2 = sym_compose(m2[31], carry.h0, carry.h1, carry.h2);
    carry.h0 = n0; carry.h1 = n1; carry.h2 = n2;
    carry.lit = sat_add(carry.lit, sl[31]);
    carry.out = sat_add(carry.out, so[31]);
  }
  *fin = carry;
  return bad == 0;
}
#endif

// ------------------------------------------------------------------------------------------------------------- lit
// Huffman decoding table for tree description p[0..len) into h (entries sym | nbits << 8; 1 << max_bits of them).
// Returns max_bits, or 0 when malformed.  Scratch lives on the caller's stack.
ZN_HD uint32_t huf_build(uint16_t* h, const uint8_t* p, uint32_t len) {
  uint8_t w[256];
  if (len < 1) return 0;
  const uint32_t hb = p[0];
  int n = 0;
  if (hb >= 128) {
    n = (int)hb - 127;
    const uint32_t bytes = (uint32_t)(n + 1) / 2;
    if (1 + bytes > len) return 0;
    for (int i = 0; i < n; i++) w[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
  } else {
    if (hb == 0 || hb + 1 > len) return 0;
    int16_t norm[16];
    int log, nsym;
    const int hd = zs::fse_read_ncount(p + 1, hb, 6, 11, norm, &log, &nsym);
    if (hd < 0 || (uint32_t)hd >= hb) return 0;
    uint32_t tab[64];
    uint16_t next[16];
    {  // zs::fse_build on a 64-entry table
      const int size = 1 << log;
      int high = size - 1;
      for (int s = 0; s < nsym; s++) {
        if (norm[s] == -1) { tab[high--] = (uint32_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
      }
      const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
      int pos = 0;
      for (int s = 0; s < nsym; s++)
        for (int i = 0; i < norm[s]; i++) {
          tab[pos] = (uint32_t)s;
          do pos = (pos + step) & mask; while (pos > high);
        }
      for (int u = 0; u < size; u++) {
        const uint32_t s = tab[u];
        const uint32_t ns = next[s]++;
        const uint32_t nb = (uint32_t)(log - hibit32(ns));
        tab[u] = zs::fse_pack(s, nb, (ns << nb) - (uint32_t)size);
      }
    }
    BackBits b;
    if (!b.init(p + 1 + hd, hb - (uint32_t)hd)) return 0;
    b.refill();
    uint32_t s1 = b.read((uint32_t)log), s2 = b.read((uint32_t)log);
    if (b.bits_left < 0) return 0;
    for (;;) {
      if (n >= 254) return 0;
      b.refill();
      const uint32_t e1 = tab[s1];
      w[n++] = (uint8_t)zs::fse_sym(e1);
      s1 = zs::fse_base(e1) + b.read(zs::fse_nbits(e1));
      if (b.bits_left < 0) { w[n++] = (uint8_t)zs::fse_sym(tab[s2]); break; }
      if (n >= 254) return 0;
      const uint32_t e2 = tab[s2];
      w[n++] = (uint8_t)zs::fse_sym(e2);
      s2 = zs::fse_base(e2) + b.read(zs::fse_nbits(e2));
      if (b.bits_left < 0) { w[n++] = (uint8_t)zs::fse_sym(tab[s1]); break; }
    }
    if (n > 255) return 0;
  }
  uint32_t sum = 0;
  for (int i = 0; i < n; i++) {
    if (w[i] > 11) return 0;
    if (w[i]) sum += 1u << (w[i] - 1);
  }
  if (sum == 0) return 0;
  const int max_bits = hibit32(sum) + 1;
  if (max_bits > 11) return 0;
  const uint32_t left = (1u << max_bits) - sum;
  if (left & (left - 1)) return 0;
  w[n++] = (uint8_t)(hibit32(left) + 1);
  uint32_t rank_start[13], count[13];
  for (int i = 0; i < 13; i++) count[i] = 0;
  for (int i = 0; i < n; i++) count[w[i]]++;
  uint32_t pos = 0;
  for (int wt = 1; wt <= max_bits; wt++) { rank_start[wt] = pos; pos += count[wt] << (wt - 1); }
  if (pos != (1u << max_bits)) return 0;
  for (int s = 0; s < n; s++) {
    const uint32_t wt = w[s];
    if (!wt) continue;
    const uint32_t span = 1u << (wt - 1), start = rank_start[wt];
    const uint16_t ent = (uint16_t)((uint32_t)s | ((uint32_t)(max_bits + 1 - wt) << 8));
    for (uint32_t k = 0; k < span; k++) h[start + k] = ent;
    rank_start[wt] = start + span;
  }
  return (uint32_t)max_bits;
}

// Stream k (of `streams`) of block b: source range and output range.  false = malformed.
ZN_HD bool huf_stream_ranges(const uint8_t* src, const ZBlock* b, uint32_t k, uint32_t* s_off, uint32_t* s_len,
                             uint32_t* o_off, uint32_t* o_len) {
  const uint8_t* q = src + b->lit_off;
  const uint32_t qlen = b->lit_clen, regen = b->lit_regen;
  if (!(b->flags & ZB_STREAMS4)) {
    *s_off = 0; *s_len = qlen; *o_off = 0; *o_len = regen;
    return k == 0;
  }
  if (qlen < 6) return false;
  const uint32_t s1 = ld16le(q), s2 = ld16le(q + 2), s3 = ld16le(q + 4);
  if (6 + s1 + s2 + s3 > qlen) return false;
  const uint32_t seg = (regen + 3) / 4;
  if (seg * 3 > regen) return false;
  const uint32_t so[4] = {6, 6 + s1, 6 + s1 + s2, 6 + s1 + s2 + s3};
  const uint32_t sl[4] = {s1, s2, s3, qlen - (6 + s1 + s2 + s3)};
  *s_off = so[k]; *s_len = sl[k];
  *o_off = seg * k; *o_len = k < 3 ? seg : regen - 3 * seg;
  return true;
}

// One Huffman stream of n symbols, decoded by the calling thread; `out` may have any alignment, full words are
// stored as words (a lane-per-stream warp would otherwise issue 32 scattered byte stores per symbol).
ZN_HD bool huf_decode_stream_w(const uint16_t* h, uint32_t mb, const uint8_t* p, uint32_t len, uint8_t* out, uint32_t n) {
  BackBits b;
  if (!b.init(p, len)) return false;
  uint32_t i = 0;
  while (i < n && (reinterpret_cast<uintptr_t>(out + i) & 3)) {
    b.refill();
    const uint32_t e0 = h[b.peek(mb)];
    b.skip(e0 >> 8);
    out[i++] = (uint8_t)e0;
  }
  for (; i + 4 <= n; i += 4) {
    b.refill();
    const uint32_t e0 = h[b.peek(mb)];
    b.skip(e0 >> 8);
    const uint32_t e1 = h[b.peek(mb)];
    b.skip(e1 >> 8);
    b.refill();
    const uint32_t e2 = h[b.peek(mb)];
    b.skip(e2 >> 8);
    const uint32_t e3 = h[b.peek(mb)];
    b.skip(e3 >> 8);
    *reinterpret_cast<uint32_t*>(out + i) = (e0 & 0xFFu) | ((e1 & 0xFFu) << 8) | ((e2 & 0xFFu) << 16) | (e3 << 24);
  }
  for (; i < n; i++) {
    b.refill();
    const uint32_t e0 = h[b.peek(mb)];
    b.skip(e0 >> 8);
    out[i] = (uint8_t)e0;
  }
  return b.bits_left == 0;
}

// ----------------------------------------------------------------------------------------------------------- chain
// One thread per blob: output offsets, true histories, frame sizes.  Returns false when the blob has to go to the
// legacy decoder (any phase reported a problem, sizes do not add up).
ZN_HD bool chain_blob(ZBlock* blocks, uint32_t n_blocks, uint32_t dst_cap) {
  uint32_t pos = 0, r0 = 1, r1 = 4, r2 = 8, frame_start = 0;
  uint64_t fcs = 0;
  bool has_fcs = false;
  for (uint32_t j = 0; j < n_blocks; j++) {
    ZBlock* b = &blocks[j];
    const uint32_t type = b->flags & ZB_TYPE_MASK;
    if (b->flags & ZB_FIRST) {
      r0 = 1; r1 = 4; r2 = 8;
      frame_start = pos;
      has_fcs = (b->flags & ZB_HAS_FCS) != 0;
      fcs = (uint64_t)b->fcs_lo | ((uint64_t)b->fcs_hi << 32);
    }
    uint32_t dec;
    if (type != 2) dec = b->len;
    else {
      if (b->st_seq != 0 || b->st_lit != 0) return false;
      if (b->lit_used > b->lit_regen) return false;
      dec = b->matched + (b->lit_regen - b->lit_used);
      if (dec > kZstdBlockMax) return false;
    }
    if (dec > dst_cap - pos) return false;
    b->out_base = pos;
    b->rep_in[0] = r0; b->rep_in[1] = r1; b->rep_in[2] = r2;
    b->frame_start = frame_start;
    if (type == 2 && b->nseq) {
      const uint32_t n0 = sym_resolve(b->rep_fin[0], r0, r1, r2), n1 = sym_resolve(b->rep_fin[1], r0, r1, r2),
                     n2 = sym_resolve(b->rep_fin[2], r0, r1, r2);
      if (!n0 || !n1 || !n2) return false;
      r0 = n0; r1 = n1; r2 = n2;
    }
    pos += dec;
    if ((b->flags & ZB_LAST) && has_fcs && (uint64_t)(pos - frame_start) != fcs) return false;
  }
  return pos == dst_cap;
}

// predefined table set in the fat format (kTabSet entries; LL log 6, OF log 5, ML log 6)
inline void build_predef_set(FseD* set) {
  uint16_t next[64];
  build_fat_table(set + kTabOffLL, 0, zs::kLLDefault, 36, 6, next);
  build_fat_table(set + kTabOffOF, 1, zs::kOFDefault, 29, 5, next);
  build_fat_table(set + kTabOffML, 2, zs::kMLDefault, 53, 6, next);
}

#if !defined(__CUDA_ARCH__)
}  // namespace zp
}  // namespace zn
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>
namespace zn {
namespace zp {
// ---- host emulation of the whole pipeline for ONE blob (serial), used by tests/host_emu.  Returns 0 when the pipeline
// decoded the blob (bytes in out), 1 when it would hand the blob to the legacy decoder.
inline int host_pipeline(const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap, const FseD* predef_set,
                         uint64_t* stats /* nullable: [0] blocks [1] sequences [2] huffman literal bytes */) {
  const uint32_t slot_cap = cap / 16384 + 8;  // level >= 16 frames split their 128 KiB blocks
  std::vector<ZBlock> blocks(slot_cap);
  WalkNeeds needs;
  const uint32_t nb = walk_blob(src, src_len, slot_cap, blocks.data(), &needs);
  if (nb == ~0u) return 1;
  std::vector<SeqRec16> recs(needs.nseq + 1);
  std::vector<uint8_t> lits((size_t)needs.lit16 * 16 + 16);
  std::vector<FseD> tabs((size_t)(needs.tabs + 1) * kTabSet);
  // tables
  for (uint32_t j = 0; j < nb; j++) {
    ZBlock* b = &blocks[j];
    if ((b->flags & ZB_TYPE_MASK) != 2 || b->nseq == 0) continue;
    FseD* set = b->tab_slot != kNoSlot ? tabs.data() + (size_t)b->tab_slot * kTabSet : nullptr;  // null: defines nothing
    int16_t norm[3][64];
    int nsym[3], log[3];
    uint32_t bits_off;
    if (!parse_table_descs(src, b, set, norm, nsym, log, &bits_off)) continue;
    bool ok = true;
    const uint32_t offs[3] = {kTabOffLL, kTabOffOF, kTabOffML};
    uint32_t tlogs = 0;
    for (int k = 0; k < 3; k++) {
      if (log[k] >= 5) {  // FSE-described (mode 2); RLE tables (log 0) were written by the parser
        uint16_t next[64];
        ok = ok && build_fat_table(set + offs[k], k, norm[k], nsym[k], log[k], next);
        if (ok) {  // the warp-parallel construction (k_ztables) must give the same table
          FseD alt[512];
          if (!build_fat_table_by_position(alt, k, norm[k], nsym[k], log[k]) || memcmp(alt, set + offs[k], sizeof(FseD) << log[k]) != 0) return 2;
        }
      }
      if (log[k] >= 0) tlogs |= (uint32_t)log[k] << (8 * k);
    }
    if (!ok) continue;
    b->tlogs = tlogs;
    b->bits_off = bits_off;
  }
  // seq
  for (uint32_t j = 0; j < nb; j++) {
    ZBlock* b = &blocks[j];
    if ((b->flags & ZB_TYPE_MASK) != 2 || b->nseq == 0) continue;
    if (b->bits_off == ~0u) continue;
    HostTabs st;
    uint32_t logs[3];
    bool ok = true;
    const uint32_t offs[3] = {kTabOffLL, kTabOffOF, kTabOffML};
    const uint32_t plog[3] = {6, 5, 6};
    for (int k = 0; k < 3; k++) {
      const uint32_t def = b->seq_def[k];
      if (def == kDefPredef) { st.t[k] = predef_set + offs[k]; logs[k] = plog[k]; }
      else if (def == kDefNone) ok = false;
      else {
        const ZBlock* d = &blocks[def];
        if (d->bits_off == ~0u || d->tab_slot == kNoSlot) ok = false;
        else { st.t[k] = tabs.data() + (size_t)d->tab_slot * kTabSet + offs[k]; logs[k] = (d->tlogs >> (8 * k)) & 0xFFu; }
      }
    }
    if (!ok) continue;
    // the one-pass form is the cross-check, the two-phase form (as k_zseq1 / k_zseq2 run it) is what the pipeline uses
    std::vector<SeqRec16> ref(b->nseq);
    uint32_t ref_matched = 0, ref_lit = 0, ref_rep[3] = {0, 0, 0};
    const bool ok_ref = decode_sequences(src, b, st, logs, ref.data(), &ref_matched, &ref_lit, ref_rep);
    SeqRec16* rec = recs.data() + b->seq_base;
    std::vector<SeqP1> p1(b->nseq);
    SeqBits sb1;
    bool ok2 = seq_phase1(sb1, src, b, st, logs, p1.data());
    {  // the device's staged reader must see the same stream
      std::vector<SeqP1> p1r(b->nseq);
      RingBits rb;
      if (seq_phase1(rb, src, b, st, logs, p1r.data()) != ok2) return 2;
      if (ok2 && memcmp(p1r.data(), p1.data(), (size_t)b->nseq * sizeof(SeqP1)) != 0) return 2;
    }
    if (ok2) {
      SeqBits sb;
      sb.init(src + b->bits_off, b->src_off + b->len - b->bits_off);
      RunSum tot;
      ok2 = seq_phase2_host(sb, st, p1.data(), rec, b->nseq, b->lit_regen, &tot);
      if (ok2) {
        b->matched = tot.out; b->lit_used = tot.lit;
        b->rep_fin[0] = tot.h0; b->rep_fin[1] = tot.h1; b->rep_fin[2] = tot.h2;
      }
    }
    if (ok2 != ok_ref) return 2;
    if (ok2) {
      if (memcmp(ref.data(), rec, (size_t)b->nseq * sizeof(SeqRec16)) != 0) return 2;
      if (b->matched != ref_matched || b->lit_used != ref_lit || memcmp(b->rep_fin, ref_rep, sizeof ref_rep) != 0) return 2;
      b->st_seq = 0;
    }
  }
  // lit
  for (uint32_t j = 0; j < nb; j++) {
    ZBlock* b = &blocks[j];
    const uint32_t lt = (b->flags >> ZB_LIT_SHIFT) & 3u;
    if ((b->flags & ZB_TYPE_MASK) != 2 || lt < 2) continue;
    const ZBlock* d = &blocks[b->huf_def];
    uint16_t h[2048];
    const uint32_t mb = huf_build(h, src + d->huf_desc, d->huf_dlen);
    if (!mb) continue;
    const uint32_t streams = (b->flags & ZB_STREAMS4) ? 4u : 1u;
    for (uint32_t k = 0; k < streams; k++) {
      uint32_t so, sl, oo, ol;
      if (!huf_stream_ranges(src, b, k, &so, &sl, &oo, &ol)) continue;
      if (huf_decode_stream_w(h, mb, src + b->lit_off + so, sl, lits.data() + (size_t)b->lit_base16 * 16 + oo, ol)) b->st_lit--;
    }
  }
  if (getenv("ZP_DEBUG")) for (uint32_t j = 0; j < nb; j++) fprintf(stderr, "blk %u type %u nseq %u st_seq %u st_lit %u bits_off %u tab %u defs %x %x %x lt %u\n", j, blocks[j].flags & 3, blocks[j].nseq, blocks[j].st_seq, blocks[j].st_lit, blocks[j].bits_off, blocks[j].tab_slot, blocks[j].seq_def[0], blocks[j].seq_def[1], blocks[j].seq_def[2], (blocks[j].flags >> 4) & 3);
  if (!chain_blob(blocks.data(), nb, cap)) return 1;
  // exec (plain serial loops; the device kernel's group logic is checked on the GPU against the same oracle)
  uint64_t nseq_total = 0, nlit = 0;
  for (uint32_t j = 0; j < nb; j++) {
    const ZBlock* b = &blocks[j];
    const uint32_t type = b->flags & ZB_TYPE_MASK;
    uint8_t* o = out + b->out_base;
    if (type == 0) memcpy(o, src + b->src_off, b->len);
    else if (type == 1) memset(o, src[b->src_off], b->len);
    else {
      const uint32_t lt = (b->flags >> ZB_LIT_SHIFT) & 3u;
      const uint8_t* lit = lt == 0 ? src + b->lit_off : lits.data() + (size_t)b->lit_base16 * 16;
      const int rle = lt == 1 ? (int)b->lit_off : -1;
      if (lt >= 2) nlit += b->lit_regen;
      nseq_total += b->nseq;
      if (stats && getenv("ZP_STATS")) {  // development: true dependence depth of the block's matches (sources inside the block)
        std::vector<uint16_t> bd(kZstdBlockMax + 16, 0);
        uint32_t maxd = 0;
        for (uint32_t i = 0; i < b->nseq; i++) {
          const SeqRec16 r = recs[b->seq_base + i];
          const uint32_t off = sym_resolve(rec_off(r), b->rep_in[0], b->rep_in[1], b->rep_in[2]);
          const int32_t d = (int32_t)(rec_out(r) + rec_ll(r)), sp = d - (int32_t)off, ml = (int32_t)rec_ml(r);
          const int32_t se = (int32_t)off >= ml ? sp + ml : d;
          uint32_t dep = 0;
          for (int32_t p = sp < 0 ? 0 : sp; p < se; p++) dep = std::max<uint32_t>(dep, bd[p]);
          dep += 1;
          for (int32_t k = 0; k < ml; k++) bd[d + k] = (uint16_t)dep;
          maxd = std::max(maxd, dep);
        }
        stats[26] += maxd;
        // wavefront passes with groups of <= 2048 sequences / 32 KiB: depth counted inside each group only
        {
          const uint32_t GBW = 32768, GS = 2048;
          uint32_t s0 = 0, gpos = 0;
          while (s0 < b->nseq) {
            uint32_t cnt = 0;
            while (cnt < GS && s0 + cnt < b->nseq) {
              const SeqRec16 r = recs[b->seq_base + s0 + cnt];
              if (rec_out(r) + rec_ll(r) + rec_ml(r) - gpos > GBW) break;
              cnt++;
            }
            if (cnt == 0) { const SeqRec16 r = recs[b->seq_base + s0]; gpos = rec_out(r) + rec_ll(r) + rec_ml(r); s0++; continue; }
            std::fill(bd.begin(), bd.end(), 0);
            uint32_t gd = 0;
            for (uint32_t i = 0; i < cnt; i++) {
              const SeqRec16 r = recs[b->seq_base + s0 + i];
              const uint32_t off = sym_resolve(rec_off(r), b->rep_in[0], b->rep_in[1], b->rep_in[2]);
              const int32_t d = (int32_t)(rec_out(r) + rec_ll(r)), sp = d - (int32_t)off, ml = (int32_t)rec_ml(r);
              const int32_t se = (int32_t)off >= ml ? sp + ml : d;
              uint32_t dep = 0;
              for (int32_t p = sp < (int32_t)gpos ? (int32_t)gpos : sp; p < se; p++) dep = std::max<uint32_t>(dep, bd[p]);
              const bool farm = (int32_t)off >= ml && se <= (int32_t)gpos;
              if (!farm) dep += 1;
              for (int32_t k = 0; k < ml; k++) bd[d + k] = (uint16_t)dep;
              gd = std::max(gd, dep);
            }
            stats[27] += gd; stats[28]++;
            const SeqRec16 rl = recs[b->seq_base + s0 + cnt - 1];
            gpos = rec_out(rl) + rec_ll(rl) + rec_ml(rl);
            s0 += cnt;
          }
        }
      }
      if (stats && getenv("ZP_STATS")) {  // development: dependence structure as the exec kernel sees it
        const uint32_t GB = 16384, NT = 512;
        uint32_t s0 = 0, gpos = 0;
        while (s0 < b->nseq) {
          uint32_t cnt = 0;
          while (cnt < NT && s0 + cnt < b->nseq) {
            const SeqRec16 r = recs[b->seq_base + s0 + cnt];
            if (rec_out(r) + rec_ll(r) + rec_ml(r) - gpos > GB) break;
            cnt++;
          }
          if (cnt == 0) { stats[4]++; const SeqRec16 r = recs[b->seq_base + s0]; gpos = rec_out(r) + rec_ll(r) + rec_ml(r); s0++; continue; }
          stats[5]++;  // groups
          for (uint32_t w = 0; w < cnt; w += 32) {
            const uint32_t n = cnt - w < 32 ? cnt - w : 32;
            uint32_t pend = 0; int32_t dst[32], send[32];
            for (uint32_t l = 0; l < n; l++) {
              const SeqRec16 r = recs[b->seq_base + s0 + w + l];
              const uint32_t off = sym_resolve(rec_off(r), b->rep_in[0], b->rep_in[1], b->rep_in[2]);
              dst[l] = (int32_t)(rec_out(r) + rec_ll(r));
              const int32_t sp = dst[l] - (int32_t)off;
              send[l] = off >= rec_ml(r) ? sp + (int32_t)rec_ml(r) : dst[l];
              stats[10 + (rec_ml(r) > 64 ? 3 : rec_ml(r) > 32 ? 2 : rec_ml(r) > 16 ? 1 : 0)]++;
              stats[14 + (rec_ll(r) > 64 ? 3 : rec_ll(r) > 32 ? 2 : rec_ll(r) > 16 ? 1 : 0)]++;
              if (rec_ml(r) && send[l] > (int32_t)gpos) { pend |= 1u << l; stats[6]++; } else stats[7]++;
            }
            if (pend) stats[18]++;  // warp turns with work
            {  // exact-dependence rounds for comparison
              uint32_t pe = pend; int32_t srcp[32], mlv[32];
              for (uint32_t l = 0; l < n; l++) { const SeqRec16 r = recs[b->seq_base + s0 + w + l]; mlv[l] = (int32_t)rec_ml(r); srcp[l] = dst[l] - (int32_t)sym_resolve(rec_off(r), b->rep_in[0], b->rep_in[1], b->rep_in[2]); }
              while (pe) {
                stats[9]++;
                uint32_t ready = 0;
                for (uint32_t l = 0; l < n; l++) if (pe >> l & 1) {
                  bool ok = true;
                  for (uint32_t j = 0; j < l; j++) if ((pe >> j & 1) && dst[j] < send[l] && dst[j] + mlv[j] > srcp[l]) ok = false;
                  if (ok) ready |= 1u << l;
                }
                pe &= ~ready;
              }
            }
            while (pend) {
              stats[8]++;  // rounds
              const uint32_t f = (uint32_t)__builtin_ctz(pend);
              uint32_t ready = 0;
              for (uint32_t l = f; l < n; l++) if ((pend >> l & 1) && (l == f || send[l] <= dst[f])) ready |= 1u << l;
              pend &= ~ready;
            }
          }
          {  // compacted near list, turns of 32, watermark rounds
            std::vector<int32_t> nd, ns;
            for (uint32_t l = 0; l < cnt; l++) {
              const SeqRec16 r = recs[b->seq_base + s0 + l];
              const uint32_t off = sym_resolve(rec_off(r), b->rep_in[0], b->rep_in[1], b->rep_in[2]);
              const int32_t d = (int32_t)(rec_out(r) + rec_ll(r)), sp = d - (int32_t)off;
              const int32_t se = off >= rec_ml(r) ? sp + (int32_t)rec_ml(r) : d;
              if (rec_ml(r) && !(off >= rec_ml(r) && se <= (int32_t)gpos)) {
                nd.push_back(d); ns.push_back(se);
                stats[21 + (off < 4 ? 0 : off < 16 ? 1 : off < rec_ml(r) ? 2 : 3)]++;
                if (off < 16) stats[25] += rec_ml(r);
              }
            }
            for (size_t w = 0; w < nd.size(); w += 32) {
              const uint32_t n = (uint32_t)std::min<size_t>(32, nd.size() - w);
              uint32_t pend = n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1);
              stats[19]++;
              while (pend) {
                stats[20]++;
                const uint32_t f = (uint32_t)__builtin_ctz(pend);
                uint32_t ready = 0;
                for (uint32_t l = f; l < n; l++) if ((pend >> l & 1) && (l == f || ns[w + l] <= nd[w + f])) ready |= 1u << l;
                pend &= ~ready;
              }
            }
          }
          const SeqRec16 rl = recs[b->seq_base + s0 + cnt - 1];
          gpos = rec_out(rl) + rec_ll(rl) + rec_ml(rl);
          s0 += cnt;
        }
      }
      for (uint32_t i = 0; i < b->nseq; i++) {
        const SeqRec16 r = recs[b->seq_base + i];
        const uint32_t ll = rec_ll(r), ml = rec_ml(r), orl = rec_out(r), lr = rec_lit(r);
        for (uint32_t k = 0; k < ll; k++) o[orl + k] = rle >= 0 ? (uint8_t)rle : lit[lr + k];
        const uint32_t off = sym_resolve(rec_off(r), b->rep_in[0], b->rep_in[1], b->rep_in[2]);
        const uint32_t d = b->out_base + orl + ll;
        if (off == 0 || off > d - b->frame_start) return 1;
        for (uint32_t k = 0; k < ml; k++) out[d + k] = out[d - off + k];
      }
      for (uint32_t k = b->lit_used; k < b->lit_regen; k++) o[b->matched + k - b->lit_used] = rle >= 0 ? (uint8_t)rle : lit[k];
    }
  }
  if (stats) { stats[0] = nb; stats[1] = nseq_total; stats[2] = nlit; }
  return 0;
}

#endif

}  // namespace zp
}  // namespace zn
