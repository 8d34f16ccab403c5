// Block-parallel Zstandard decode for large blobs (8 MiB slices = 64+ blocks each).
//
// The one-team-per-blob decoder (zstd_decode.cuh) runs a frame's blocks strictly one after another, so entropy-coded
// data is bound by ONE thread's FSE state machine (~300 cycles per sequence).  Here a CTA still owns one blob, but
// works on groups of kGroup blocks:
//
//   walk   thread 0 hops over the next kGroup block headers, recording for every compressed block where its Huffman
//          tree and its three sequence tables are DEFINED (its own section, an earlier block's section for
//          treeless / repeat modes, or the predefined distribution) — this removes the table dependence between
//          blocks;
//   A      warp 2j decodes block j's sequences (FSE) while warp 2j+1 decodes its literals (Huffman): 2*kGroup
//          independent serial decoders per CTA.  Repeat offsets that reach back before the block start are kept
//          SYMBOLIC (initial history entry i, minus k), so no block waits for its predecessor;
//   B0     thread 0 chains the groups' blocks: output bases, true repeat-offset history, the few symbolic offsets;
//   B      blocks execute in order.  Long-match blocks (pattern data) go through the team executor with TMA bulk
//          stores; blocks of short sequences are executed lane-parallel: one lane per sequence, literals first (no
//          dependences), then matches in batches of 32 with an in-order completion watermark between warps.
//
// Replaces the same reference call as zstd_decode.cuh (codec::decompress_into, codec.rs:67-78).
#pragma once
#include "zstd_decode.cuh"

namespace zn {
namespace par {

constexpr int kGroup = 8;                     // blocks in flight per CTA
constexpr int kParThreads = kGroup * 64;      // warp pair per block
constexpr uint32_t kMaxSeq = 44032;           // > 128 KiB / 3 (minimum match)
constexpr uint32_t kSymCap = 24;              // symbolic repeat offsets remembered per block
constexpr uint32_t kDefPredef = 0xFFFFFFFEu;  // table source: predefined distribution
constexpr uint32_t kDefNone = 0xFFFFFFFFu;    // table source: nothing defined yet in this frame
constexpr uint32_t kSymBit = 0x80000000u;     // offset is symbolic: initial history entry (v & 3) minus ((v >> 2) & 0x1FFFFFFF)
constexpr uint32_t kLaneMax = 64;             // copies longer than this are done by the whole warp, not one lane
constexpr uint32_t kTeamAvg = 256;            // blocks averaging >= this many bytes per sequence use the team executor

struct SeqOut {
  uint32_t out_rel;  // block-relative output position of this sequence's literals
  uint32_t lit_rel;  // offset of its literals in the block's literal source
  uint32_t ll, ml;
  uint32_t off;      // match distance, possibly symbolic (kSymBit)
};

struct BlockRec {
  uint32_t off, len, type;       // content offset in src, content size, block type (0 raw, 1 RLE, 2 compressed)
  uint32_t huf_def;              // content offset of the block whose literals section holds this block's Huffman tree
  uint32_t seq_def[3];           // LL / OF / ML: content offset of the defining block, kDefPredef or kDefNone
  // ---- filled by phase A
  uint32_t status_seq, status_lit;
  uint32_t nseq, matched;        // sequences; bytes produced by them (literals + matches)
  uint32_t lit_len, lit_used;    // regenerated literals; literals consumed by the sequences
  int32_t lit_rle;               // >= 0: every literal is this byte
  uint32_t lit_in_src;           // raw literals: offset in src; ~0u: literals are in the block's scratch slot
  uint32_t rep_fin[3];           // history after the block (possibly symbolic)
  uint32_t nsym, sym_overflow;
  uint32_t sym_idx[kSymCap];
  // ---- filled by B0
  uint32_t base_out;
};

ZN_HD bool is_sym(uint32_t v) { return (v & kSymBit) != 0; }
ZN_HD uint32_t sym_make(uint32_t i) { return kSymBit | i; }
ZN_HD uint32_t sym_minus1(uint32_t v) { return v + 4u; }
// value of a (possibly symbolic) offset given the true incoming history; 0 = corrupt
ZN_HD uint32_t sym_resolve(uint32_t v, const uint32_t* r) {
  if (!is_sym(v)) return v;
  const uint32_t base = r[v & 3u], k = (v >> 2) & 0x1FFFFFFFu;
  return base > k ? base - k : 0u;
}

// Locates the sequences section of the compressed block p[0..len): returns false when malformed.
// *lit_type, *seq (pointer just behind the literals section).
ZN_HD bool skip_literals(const uint8_t* p, uint32_t len, uint32_t* lit_type, uint32_t* consumed) {
  if (len < 1) return false;
  const uint32_t b0 = p[0], type = b0 & 3, sf = (b0 >> 2) & 3;
  *lit_type = type;
  if (type < 2) {
    uint32_t hdr, regen;
    if ((sf & 1) == 0) { hdr = 1; regen = b0 >> 3; }
    else if (sf == 1) { if (len < 2) return false; hdr = 2; regen = (b0 >> 4) | ((uint32_t)p[1] << 4); }
    else { if (len < 3) return false; hdr = 3; regen = (b0 >> 4) | ((uint32_t)p[1] << 4) | ((uint32_t)p[2] << 12); }
    *consumed = type == 0 ? hdr + regen : hdr + 1;
    return *consumed <= len;
  }
  uint32_t hdr, comp;
  if (sf <= 1) { if (len < 3) return false; hdr = 3; comp = (ld24le(p) >> 14) & 0x3FF; }
  else if (sf == 2) { if (len < 4) return false; hdr = 4; comp = ld32le(p) >> 18; }
  else { if (len < 5) return false; hdr = 5; comp = (uint32_t)((((uint64_t)ld32le(p) | ((uint64_t)p[4] << 32)) >> 22)); }
  *consumed = hdr + comp;
  return *consumed <= len;
}

// Parses the Number_of_Sequences field at q; returns bytes used (0 = malformed).
ZN_HD uint32_t read_nseq(const uint8_t* q, uint32_t avail, uint32_t* nseq) {
  if (avail < 1) return 0;
  uint32_t n = q[0];
  if (n < 128) { *nseq = n; return 1; }
  if (n == 255) { if (avail < 3) return 0; *nseq = (uint32_t)q[1] + ((uint32_t)q[2] << 8) + 0x7F00u; return 3; }
  if (avail < 2) return 0;
  *nseq = ((n - 128) << 8) + q[1];
  return 2;
}

// Frame-level table provenance carried by the walker.
struct Defs {
  uint32_t huf, seq[3];
};

// Walker step for one compressed block: fills rec->huf_def / seq_def and advances `defs`.  Returns false when malformed.
ZN_HD bool walk_block(const uint8_t* src, BlockRec* rec, Defs& defs) {
  const uint8_t* p = src + rec->off;
  uint32_t lit_type, consumed;
  if (rec->len < 2 || !skip_literals(p, rec->len, &lit_type, &consumed)) return false;
  if (lit_type == 2) defs.huf = rec->off;
  rec->huf_def = lit_type >= 2 ? defs.huf : kDefNone;
  if (lit_type == 3 && defs.huf == kDefNone) return false;
  uint32_t nseq;
  const uint32_t used = read_nseq(p + consumed, rec->len - consumed, &nseq);
  if (!used) return false;
  rec->seq_def[0] = rec->seq_def[1] = rec->seq_def[2] = kDefNone;
  if (nseq) {
    if (consumed + used >= rec->len) return false;
    const uint32_t modes = p[consumed + used];
    if (modes & 3) return false;
    for (int k = 0; k < 3; k++) {
      const uint32_t m = (modes >> (6 - 2 * k)) & 3;
      if (m == 0) defs.seq[k] = kDefPredef;
      else if (m != 3) defs.seq[k] = rec->off;
      else if (defs.seq[k] == kDefNone) return false;
      rec->seq_def[k] = defs.seq[k];
    }
  }
  return true;
}

ZN_HD void table_params(int k, int* max_log, int* max_sym, int* plog) {
  *max_log = k == 1 ? 8 : 9;
  *max_sym = k == 0 ? 35 : (k == 1 ? 31 : 52);
  *plog = k == 1 ? 5 : 6;
}

// Builds sequence table k of the block whose content starts at src+def_off (content size unknown here: bounded by
// src_len) into t.  One thread.
ZN_HD bool build_external_table(zs::FseTable* t, int k, const uint8_t* src, uint32_t src_len, uint32_t def_off, DecShared* sh) {
  // the defining block's header sits 3 bytes before its content
  const uint32_t bh = ld24le(src + def_off - 3);
  const uint32_t len = bh >> 3;
  if (def_off + len > src_len) return false;
  const uint8_t* p = src + def_off;
  uint32_t lit_type, consumed, nseq;
  if (!skip_literals(p, len, &lit_type, &consumed)) return false;
  const uint32_t used = read_nseq(p + consumed, len - consumed, &nseq);
  if (!used || !nseq || consumed + used >= len) return false;
  const uint32_t modes = p[consumed + used];
  const uint8_t* q = p + consumed + used + 1;
  const uint8_t* end = p + len;
  for (int j = 0; j <= k; j++) {
    const uint32_t m = (modes >> (6 - 2 * j)) & 3;
    int max_log, max_sym, plog;
    table_params(j, &max_log, &max_sym, &plog);
    if (j == k) {
      if (m != 1 && m != 2) return false;
      return zs::setup_seq_table(t, m, q, end, max_log, max_sym, nullptr, plog, sh);
    }
    if (m == 1) {
      if (q >= end) return false;
      q++;
    } else if (m == 2) {
      int log, nsym;
      const int u = zs::fse_read_ncount(q, (uint32_t)(end - q), max_log, max_sym, sh->norm, &log, &nsym);
      if (u < 0) return false;
      q += u;
    }
  }
  return false;
}

// Phase A, sequence half: one thread decodes every sequence of block `rec` into `out` (kMaxSeq records).
// `predef_mask` (per slot, persistent) tells which tables of `sh` already hold the predefined distribution.
ZN_HD void decode_block_sequences(const uint8_t* src, uint32_t src_len, BlockRec* rec, DecShared* sh, SeqOut* out,
                                  uint32_t& predef_mask) {
  rec->status_seq = S_DECODE_ERROR;
  rec->nseq = 0; rec->matched = 0; rec->lit_used = 0; rec->nsym = 0; rec->sym_overflow = 0;
  rec->rep_fin[0] = sym_make(0); rec->rep_fin[1] = sym_make(1); rec->rep_fin[2] = sym_make(2);
  const uint8_t* p = src + rec->off;
  const uint32_t len = rec->len;
  uint32_t lit_type, consumed, nseq;
  if (!skip_literals(p, len, &lit_type, &consumed)) return;
  const uint32_t used = read_nseq(p + consumed, len - consumed, &nseq);
  if (!used) return;
  rec->nseq = nseq;
  if (nseq == 0) { rec->status_seq = (consumed + used == len) ? S_OK : S_DECODE_ERROR; return; }
  if (nseq > kMaxSeq) return;
  const uint32_t modes = p[consumed + used];
  const uint8_t* q = p + consumed + used + 1;
  const uint8_t* end = p + len;
  const PredefTables* pd = predef_tables();
  zs::FseTable* tabs[3] = {&sh->ll, &sh->of, &sh->ml};
  const uint32_t* pdt[3] = {pd->ll, pd->of, pd->ml};
  for (int k = 0; k < 3; k++) {
    const uint32_t m = (modes >> (6 - 2 * k)) & 3;
    int max_log, max_sym, plog;
    table_params(k, &max_log, &max_sym, &plog);
    const uint32_t def = rec->seq_def[k];
    if (m == 0 || (m == 3 && def == kDefPredef)) {
      if (!(predef_mask & (1u << k))) {
        const int n = 1 << plog;
        for (int i = 0; i < n; i++) tabs[k]->e[i] = pdt[k][i];
        predef_mask |= 1u << k;
      }
      tabs[k]->log = (uint32_t)plog;
      tabs[k]->valid = 1;
    } else if (m == 3) {
      predef_mask &= ~(1u << k);
      if (def == kDefNone || !build_external_table(tabs[k], k, src, src_len, def, sh)) return;
    } else {
      predef_mask &= ~(1u << k);
      if (!zs::setup_seq_table(tabs[k], m, q, end, max_log, max_sym, nullptr, plog, sh)) return;
    }
  }
  BackBits b;
  if (!b.init(q, (uint32_t)(end - q))) return;
  b.refill();
  uint32_t sl = b.read(sh->ll.log), so = b.read(sh->of.log), sm = b.read(sh->ml.log);
  if (b.bits_left < 0) return;
  uint32_t h0 = sym_make(0), h1 = sym_make(1), h2 = sym_make(2);
  uint32_t lit_pos = 0, out_pos = 0, nsym = 0;
  const uint32_t* tl = sh->ll.e;
  const uint32_t* to = sh->of.e;
  const uint32_t* tm = sh->ml.e;
  for (uint32_t i = 0; i < nseq; i++) {
    const uint32_t el = tl[sl], eo = to[so], em = tm[sm];
    const uint32_t lc = zs::fse_sym(el), oc = zs::fse_sym(eo), mc = zs::fse_sym(em);
    if (oc > 30 || mc > 52 || lc > 35) return;
    b.refill();
    const uint32_t ov = (1u << oc) + b.read(oc);
    b.refill();
    const uint32_t xm = sh->lut_ml[mc], xl = sh->lut_ll[lc];
    const uint32_t ml = (xm & 0xFFFFFFu) + b.read(xm >> 24);
    const uint32_t ll = (xl & 0xFFFFFFu) + b.read(xl >> 24);
    if (i + 1 < nseq) {
      b.refill();
      sl = zs::fse_base(el) + b.read(zs::fse_nbits(el));
      sm = zs::fse_base(em) + b.read(zs::fse_nbits(em));
      so = zs::fse_base(eo) + b.read(zs::fse_nbits(eo));
    }
    if (b.bits_left < 0) return;
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
          else { offset = h0 - 1; if (offset == 0) return; }
        } else offset = idx == 1 ? h1 : h2;
        if (idx != 1) h2 = h1;
        h1 = h0;
        h0 = offset;
      }
    }
    if (is_sym(offset)) {
      if (nsym < kSymCap) rec->sym_idx[nsym] = i;
      nsym++;
    }
    if (ll > rec->lit_len - lit_pos) return;  // lit_len was published by the walker from the literals header
    if (ll > kZstdBlockMax || ml > kZstdBlockMax + 3u || out_pos + ll + ml > kZstdBlockMax) return;
    SeqOut r;
    r.out_rel = out_pos; r.lit_rel = lit_pos; r.ll = ll; r.ml = ml; r.off = offset;
    out[i] = r;
    lit_pos += ll;
    out_pos += ll + ml;
  }
  if (b.bits_left != 0) return;
  rec->matched = out_pos;
  rec->lit_used = lit_pos;
  rec->rep_fin[0] = h0; rec->rep_fin[1] = h1; rec->rep_fin[2] = h2;
  rec->nsym = nsym < kSymCap ? nsym : kSymCap;
  rec->sym_overflow = nsym > kSymCap;
  rec->status_seq = S_OK;
}

// Regenerated size of the literals section of block p (header only); false when malformed.
ZN_HD bool literal_regen_size(const uint8_t* p, uint32_t len, uint32_t* regen) {
  if (len < 1) return false;
  const uint32_t b0 = p[0], type = b0 & 3, sf = (b0 >> 2) & 3;
  if (type < 2) {
    if ((sf & 1) == 0) *regen = b0 >> 3;
    else if (sf == 1) { if (len < 2) return false; *regen = (b0 >> 4) | ((uint32_t)p[1] << 4); }
    else { if (len < 3) return false; *regen = (b0 >> 4) | ((uint32_t)p[1] << 4) | ((uint32_t)p[2] << 12); }
  } else {
    if (sf <= 1) { if (len < 3) return false; *regen = (ld24le(p) >> 4) & 0x3FF; }
    else if (sf == 2) { if (len < 4) return false; *regen = (ld32le(p) >> 4) & 0x3FFF; }
    else { if (len < 5) return false; *regen = (uint32_t)(((uint64_t)ld32le(p) | ((uint64_t)p[4] << 32)) >> 4) & 0x3FFFF; }
  }
  return *regen <= kZstdBlockMax;
}

// Phase A, literal half: a 32-thread team (one warp; tid 0 = lane 0) regenerates block `rec`'s literals.
ZN_HD void decode_block_literals(const Team& t, const uint8_t* src, uint32_t src_len, BlockRec* rec, DecShared* sh,
                                 uint8_t* lit_slot) {
  const uint8_t* p = src + rec->off;
  const uint32_t type = p[0] & 3;
  if (type == 3) {  // treeless: rebuild the tree from the block that defined it
    if (t.tid == 0) {
      const uint32_t def = rec->huf_def;
      uint32_t ok = 0;
      if (def != kDefNone) {
        const uint32_t dlen = ld24le(src + def - 3) >> 3;
        if (def + dlen <= src_len) {
          const uint8_t* dp = src + def;
          const uint32_t sf = (dp[0] >> 2) & 3;
          const uint32_t hdr = sf <= 1 ? 3u : (sf == 2 ? 4u : 5u);
          uint32_t lt, cons;
          if (skip_literals(dp, dlen, &lt, &cons) && lt == 2 && cons > hdr)
            ok = zs::huf_read_table(&sh->huf, dp + hdr, cons - hdr, sh->weights, &sh->wt, sh->next) > 0;
        }
      }
      if (!ok) sh->huf.valid = 0;
    }
    team_sync(t);
  }
  zs::LitInfo li;
  const uint32_t rc = zs::decode_literals(t, sh, p, rec->len, lit_slot, li);
  if (t.tid == 0) {
    rec->status_lit = rc;
    rec->lit_rle = li.rle;
    rec->lit_in_src = (rc == S_OK && li.rle < 0 && li.base != lit_slot) ? (uint32_t)(li.base - src) : 0xFFFFFFFFu;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Walker: thread 0 collects up to kGroup blocks starting at *ip.  Returns S_OK or an error status.
struct GroupInfo {
  uint32_t n;       // blocks in the group
  uint32_t last;    // the frame's last block is in the group
  uint32_t next_ip;
  uint32_t status;
};

ZN_HD void walk_group(const uint8_t* src, uint32_t src_len, uint32_t ip, uint32_t block_max, Defs& defs, BlockRec* recs,
                      GroupInfo* gi) {
  gi->n = 0; gi->last = 0; gi->status = S_OK;
  while (gi->n < (uint32_t)kGroup) {
    if (src_len - ip < 3) { gi->status = S_DECODE_ERROR; break; }
    const uint32_t bh = ld24le(src + ip);
    ip += 3;
    const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsize = bh >> 3;
    BlockRec* r = &recs[gi->n];
    r->off = ip; r->len = bsize; r->type = type;
    r->status_seq = r->status_lit = S_OK;
    r->nseq = 0; r->matched = 0; r->lit_len = 0; r->lit_used = 0; r->nsym = 0; r->sym_overflow = 0;
    if (type == 3) { gi->status = S_DECODE_ERROR; break; }
    if (type == 1) {
      if (src_len - ip < 1) { gi->status = S_DECODE_ERROR; break; }
      ip += 1;
    } else {
      if (bsize > src_len - ip) { gi->status = S_DECODE_ERROR; break; }
      if (type == 2) {
        if (bsize > block_max || bsize < 2 || !walk_block(src, r, defs) ||
            !literal_regen_size(src + r->off, r->len, &r->lit_len)) { gi->status = S_DECODE_ERROR; break; }
      }
      ip += bsize;
    }
    gi->n++;
    if (last) { gi->last = 1; break; }
  }
  gi->next_ip = ip;
}

// B0 for one group: output bases, capacity, true repeat history, symbolic offsets.  One thread.
// rep[3] = history entering the group (updated).  *pos = output cursor entering the group (updated).
ZN_HD uint32_t chain_group(BlockRec* recs, uint32_t n, SeqOut* const* seqs, uint32_t* rep, uint32_t* pos, uint32_t cap,
                           uint32_t frame_start) {
  for (uint32_t j = 0; j < n; j++) {
    BlockRec* r = &recs[j];
    r->base_out = *pos;
    uint32_t dec;
    if (r->type != 2) dec = r->len;
    else {
      if (r->status_lit != S_OK) return r->status_lit;
      if (r->status_seq != S_OK) return r->status_seq;
      if (r->lit_used > r->lit_len) return S_DECODE_ERROR;
      dec = r->matched + (r->lit_len - r->lit_used);
      if (dec > kZstdBlockMax) return S_DECODE_ERROR;
    }
    if (dec > cap - *pos) return S_DST_TOO_SMALL;
    if (r->type == 2 && r->nseq) {
      SeqOut* so = seqs[j];
      if (r->sym_overflow) {
        for (uint32_t i = 0; i < r->nseq; i++)
          if (is_sym(so[i].off)) { const uint32_t v = sym_resolve(so[i].off, rep); if (!v) return S_DECODE_ERROR; so[i].off = v; }
      } else {
        for (uint32_t k = 0; k < r->nsym; k++) {
          const uint32_t i = r->sym_idx[k];
          const uint32_t v = sym_resolve(so[i].off, rep);
          if (!v) return S_DECODE_ERROR;
          so[i].off = v;
        }
      }
      const uint32_t n0 = sym_resolve(r->rep_fin[0], rep), n1 = sym_resolve(r->rep_fin[1], rep), n2 = sym_resolve(r->rep_fin[2], rep);
      if (!n0 || !n1 || !n2) return S_DECODE_ERROR;
      rep[0] = n0; rep[1] = n1; rep[2] = n2;
    }
    *pos += dec;
    (void)frame_start;
  }
  return S_OK;
}

#if !defined(__CUDA_ARCH__)
// ---- host emulation of the whole block-parallel pipeline (serial), used by tests/host_emu to validate the walker,
// the symbolic repeat offsets, external table rebuilds and the chaining on the CPU.
inline uint32_t host_decode_frames_par(const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap, uint32_t* produced) {
  static DecShared slots[kGroup];
  static SeqOut* seqbuf[kGroup] = {nullptr};
  static uint8_t* litbuf[kGroup] = {nullptr};
  static BlockRec recs[kGroup];
  if (!seqbuf[0])
    for (int j = 0; j < kGroup; j++) { seqbuf[j] = new SeqOut[kMaxSeq]; litbuf[j] = new uint8_t[kZstdBlockMax + 64]; }
  uint32_t predef_mask[kGroup] = {0};
  const Team t{0, 1};
  uint32_t ip = 0, pos = 0;
  *produced = 0;
  if (src_len == 0) return S_DECODE_ERROR;
  while (ip < src_len) {
    if (src_len - ip >= 8 && (ld32le(src + ip) & 0xFFFFFFF0u) == 0x184D2A50u) {
      const uint32_t sz = ld32le(src + ip + 4);
      if (sz > src_len - ip - 8) return S_DECODE_ERROR;
      ip += 8 + sz;
      continue;
    }
    if (src_len - ip < 5) return S_DECODE_ERROR;
    if (ld32le(src + ip) != 0xFD2FB528u) return ip == 0 ? S_UNSUPPORTED : S_DECODE_ERROR;
    const uint32_t fhd = src[ip + 4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did_flag = fhd & 3;
    if (fhd & 0x08) return S_UNSUPPORTED;
    const uint32_t checksum = (fhd >> 2) & 1;
    uint32_t hp = ip + 5;
    uint64_t window = 0;
    if (!single) { if (src_len < hp + 1) return S_DECODE_ERROR; const uint32_t wd = src[hp++]; const uint64_t b = 1ull << (10 + (wd >> 3)); window = b + (b >> 3) * (wd & 7); }
    const uint32_t db = did_flag == 3 ? 4u : did_flag;
    if (src_len < hp + db) return S_DECODE_ERROR;
    uint32_t did = 0;
    for (uint32_t i = 0; i < db; i++) did |= (uint32_t)src[hp + i] << (8 * i);
    hp += db;
    if (did) return S_UNSUPPORTED;
    const uint32_t fb = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2u : (fcs_flag == 2 ? 4u : 8u));
    if (src_len < hp + fb) return S_DECODE_ERROR;
    uint64_t fcs = 0;
    for (uint32_t i = 0; i < fb; i++) fcs |= (uint64_t)src[hp + i] << (8 * i);
    if (fb == 2) fcs += 256;
    hp += fb;
    if (single) window = fcs;
    ip = hp;
    const uint32_t frame_start = pos;
    const uint32_t block_max = window < kZstdBlockMax ? (uint32_t)window : kZstdBlockMax;
    Defs defs{kDefNone, {kDefNone, kDefNone, kDefNone}};
    uint32_t rep[3] = {1, 4, 8};
    for (int j = 0; j < kGroup; j++) { slots[j].huf.valid = 0; zs::init_luts(t, &slots[j]); }
    for (;;) {
      GroupInfo gi;
      walk_group(src, src_len, ip, block_max, defs, recs, &gi);
      // a malformed block ends the group early: the blocks before it still decode (byte-exact prefix, as the serial path)
      for (uint32_t j = 0; j < gi.n; j++)
        if (recs[j].type == 2) {
          decode_block_literals(t, src, src_len, &recs[j], &slots[j], litbuf[j]);
          decode_block_sequences(src, src_len, &recs[j], &slots[j], seqbuf[j], predef_mask[j]);
        }
      const uint32_t rc = chain_group(recs, gi.n, seqbuf, rep, &pos, cap, frame_start);
      if (rc != S_OK) return rc;
      for (uint32_t j = 0; j < gi.n; j++) {
        BlockRec* r = &recs[j];
        uint8_t* o = out + r->base_out;
        if (r->type == 0) memcpy(o, src + r->off, r->len);
        else if (r->type == 1) memset(o, src[r->off], r->len);
        else {
          const uint8_t* lit = r->lit_in_src != 0xFFFFFFFFu ? src + r->lit_in_src : litbuf[j];
          for (uint32_t i = 0; i < r->nseq; i++) {
            const SeqOut s = seqbuf[j][i];
            for (uint32_t k = 0; k < s.ll; k++) o[s.out_rel + k] = r->lit_rle >= 0 ? (uint8_t)r->lit_rle : lit[s.lit_rel + k];
            const uint32_t d = r->base_out + s.out_rel + s.ll;
            if (s.off > d - frame_start) return S_DECODE_ERROR;
            for (uint32_t k = 0; k < s.ml; k++) out[d + k] = out[d - s.off + k];
          }
          for (uint32_t k = r->lit_used; k < r->lit_len; k++)
            o[r->matched + k - r->lit_used] = r->lit_rle >= 0 ? (uint8_t)r->lit_rle : lit[k];
        }
      }
      *produced = pos;
      if (gi.status != S_OK) return gi.status;
      ip = gi.next_ip;
      if (gi.last) break;
    }
    if (fb != 0 && (uint64_t)(pos - frame_start) != fcs) return S_SIZE_MISMATCH;
    if (checksum) { if (src_len - ip < 4) return S_DECODE_ERROR; ip += 4; }
  }
  *produced = pos;
  return S_OK;
}
#endif

}  // namespace par
}  // namespace zn
