// Kernels of the device-wide Zstandard decode pipeline (see zpipe.cuh for the phases and what they replace:
// codec::decompress_into, znippy-common/src/codec.rs:67-78, for every row of a batch at once).
//
//   k_zwalk    thread per blob      block table + pool allocation
//   k_ztables  warp per block       FSE table descriptions -> fat decoding tables (lanes 0-2 build LL / OF / ML)
//   k_zseq     LANE per block       three interleaved FSE state machines, 16-byte sequence records
//   k_zlit     LANE per stream      Huffman literals, decoding table in shared memory, word stores
//   k_zchain   thread per blob      output offsets, repeat-offset histories, size checks
//   k_zexec    CTA per blob         sequence execution in shared memory, group by group
//
// Every kernel is latency-bound integer work; the design goal is the number of independent dependent-chains in flight
// (32 per warp in seq / lit, versus one per warp in the one-team-per-blob decoders), not bytes per instruction.
#pragma once
#include "coop.cuh"
#include "zpipe.cuh"

namespace zn {
namespace zp {

__device__ FseD g_zpredef[kTabSet];  // predefined distributions in the fat format (filled by zn_ctx_create)

constexpr uint32_t kTabWarps = 8;    // k_ztables: warps per CTA
constexpr uint32_t kLitBlocks = 16;  // k_zlit: blocks per CTA (4 lanes each), one 4 KiB Huffman table per block
constexpr uint32_t kLitSmem = kLitBlocks * 2048 * 2;

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
};

// ---------------------------------------------------------------------------------------------------------------- walk
__global__ void __launch_bounds__(64) k_zwalk(ZArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nzb) return;
  ZBlob z = a.zb[i];
  const BlobDesc d = a.blobs[z.blob];
  uint32_t state = 1, nb = 0;
  if (d.src_len < 0xFFFFFFF0ull && d.dst_cap < 0xFFFFFFF0ull && !(d.flags & F_LZ4_BLOCK)) {
    WalkNeeds needs;
    ZBlock* blocks = a.blocks + z.slot0;
    nb = walk_blob(a.blobs_base + d.src_off, (uint32_t)d.src_len, z.slot_cap, blocks, &needs);
    if (nb != ~0u) {
      ZPools* p = a.pools;
      const uint32_t comp0 = atomicAdd(&p->comp_used, needs.comp);
      bool ok = comp0 + needs.comp <= p->comp_cap;
      uint32_t seq0 = 0, lit0 = 0, tab0 = 0;
      if (ok) { seq0 = atomicAdd(&p->seq_used, needs.nseq); ok = (uint64_t)seq0 + needs.nseq <= p->seq_cap; }
      if (ok) { lit0 = atomicAdd(&p->lit_used16, needs.lit16); ok = (uint64_t)lit0 + needs.lit16 <= p->lit_cap16; }
      if (ok) { tab0 = atomicAdd(&p->tab_used, needs.tabs); ok = (uint64_t)tab0 + needs.tabs <= p->tab_cap; }
      uint32_t c = 0;
      for (uint32_t j = 0; j < nb; j++) {
        ZBlock* b = &blocks[j];
        b->seq_base += seq0;
        b->lit_base16 += lit0;
        if (b->tab_slot != kNoSlot) b->tab_slot += tab0;
        b->pad[0] = i;  // owning pipeline blob
        if ((b->flags & ZB_TYPE_MASK) == 2) {
          if (comp0 + c < p->comp_cap) a.comp_list[comp0 + c] = ok ? z.slot0 + j : kNoSlot;
          c++;
        }
      }
      if (ok) state = 0;
    } else nb = 0;
  }
  a.zb[i].n_blocks = nb;
  a.zb[i].state = state;
}

// -------------------------------------------------------------------------------------------------------------- tables
__global__ void __launch_bounds__(kTabWarps * 32) k_ztables(ZArgs a) {
  __shared__ int16_t s_norm[kTabWarps][3][64];
  __shared__ uint16_t s_next[kTabWarps][3][64];
  __shared__ int s_nsym[kTabWarps][3], s_log[kTabWarps][3];
  __shared__ uint32_t s_bits[kTabWarps], s_ok[kTabWarps];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t n_comp = min(a.pools->comp_used, a.pools->comp_cap);
  for (uint32_t it = blockIdx.x * kTabWarps + warp; it < n_comp; it += gridDim.x * kTabWarps) {
    const uint32_t slot = a.comp_list[it];
    if (slot == kNoSlot) continue;
    ZBlock* b = &a.blocks[slot];
    if (b->nseq == 0) continue;
    const BlobDesc d = a.blobs[a.zb[b->pad[0]].blob];
    const uint8_t* src = a.blobs_base + d.src_off;
    FseD* set = b->tab_slot != kNoSlot ? a.tabs + (size_t)b->tab_slot * kTabSet : nullptr;
    if (lane == 0) {
      uint32_t bits_off = 0;
      s_ok[warp] = parse_table_descs(src, b, set, s_norm[warp], s_nsym[warp], s_log[warp], &bits_off) ? 1u : 0u;
      s_bits[warp] = bits_off;
    }
    __syncwarp();
    bool ok = s_ok[warp] != 0;
    if (ok && lane < 3 && s_log[warp][lane] >= 5) {
      const uint32_t offs = lane == 0 ? kTabOffLL : (lane == 1 ? kTabOffOF : kTabOffML);
      ok = build_fat_table(set + offs, (int)lane, s_norm[warp][lane], s_nsym[warp][lane], s_log[warp][lane], s_next[warp][lane]);
    }
    ok = __all_sync(0xFFFFFFFFu, ok);
    if (lane == 0 && ok) {
      uint32_t tl = 0;
      for (int k = 0; k < 3; k++)
        if (s_log[warp][k] >= 0) tl |= (uint32_t)s_log[warp][k] << (8 * k);
      b->tlogs = tl;
      b->bits_off = s_bits[warp];
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------------------------------------------- seq
__global__ void __launch_bounds__(64) k_zseq(ZArgs a) {
  const uint32_t n_comp = min(a.pools->comp_used, a.pools->comp_cap);
  for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n_comp; it += gridDim.x * blockDim.x) {
    const uint32_t slot = a.comp_list[it];
    if (slot == kNoSlot) continue;
    ZBlock* b = &a.blocks[slot];
    if (b->nseq == 0 || b->bits_off == ~0u) continue;
    const ZBlob z = a.zb[b->pad[0]];
    const BlobDesc d = a.blobs[z.blob];
    const uint8_t* src = a.blobs_base + d.src_off;
    SeqTabs st;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const uint32_t offs = k == 0 ? kTabOffLL : (k == 1 ? kTabOffOF : kTabOffML);
      const uint32_t def = b->seq_def[k];
      st.t[k] = g_zpredef + offs;
      st.log[k] = k == 1 ? 5u : 6u;
      if (def == kDefNone) ok = false;
      else if (def != kDefPredef) {
        const ZBlock* db = &a.blocks[z.slot0 + def];
        if (db->bits_off == ~0u || db->tab_slot == kNoSlot) ok = false;
        else {
          st.t[k] = a.tabs + (size_t)db->tab_slot * kTabSet + offs;
          st.log[k] = (db->tlogs >> (8 * k)) & 0xFFu;
        }
      }
    }
    if (!ok) continue;
    uint32_t matched, lit_used, rep[3];
    if (decode_sequences(src, b, st, a.recs + b->seq_base, &matched, &lit_used, rep)) {
      b->matched = matched;
      b->lit_used = lit_used;
      b->rep_fin[0] = rep[0]; b->rep_fin[1] = rep[1]; b->rep_fin[2] = rep[2];
      b->st_seq = 0;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------- lit
__global__ void __launch_bounds__(kLitBlocks * 4) k_zlit(ZArgs a) {
  extern __shared__ __align__(16) uint8_t lit_smem[];
  __shared__ uint32_t s_mb[kLitBlocks];
  uint16_t* tables = reinterpret_cast<uint16_t*>(lit_smem);
  const uint32_t q = threadIdx.x >> 2, k = threadIdx.x & 3u;
  const uint32_t n_comp = min(a.pools->comp_used, a.pools->comp_cap);
  for (uint32_t base = blockIdx.x * kLitBlocks; base < n_comp; base += gridDim.x * kLitBlocks) {
    const uint32_t it = base + q;
    ZBlock* b = nullptr;
    const uint8_t* src = nullptr;
    uint32_t slot0 = 0;
    if (it < n_comp) {
      const uint32_t slot = a.comp_list[it];
      if (slot != kNoSlot) {
        b = &a.blocks[slot];
        if (((b->flags >> ZB_LIT_SHIFT) & 3u) < 2) b = nullptr;
      }
    }
    if (b) {
      const ZBlob z = a.zb[b->pad[0]];
      src = a.blobs_base + a.blobs[z.blob].src_off;
      slot0 = z.slot0;
    }
    uint16_t* h = tables + q * 2048;
    if (k == 0) {
      uint32_t mb = 0;
      if (b) {
        const ZBlock* db = &a.blocks[slot0 + b->huf_def];
        mb = huf_build(h, src + db->huf_desc, db->huf_dlen);
      }
      s_mb[q] = mb;
    }
    __syncthreads();
    const uint32_t mb = s_mb[q];
    if (b && mb && k < ((b->flags & ZB_STREAMS4) ? 4u : 1u)) {
      uint32_t so, sl, oo, ol;
      if (huf_stream_ranges(src, b, k, &so, &sl, &oo, &ol) &&
          huf_decode_stream_w(h, mb, src + b->lit_off + so, sl, a.lits + (size_t)b->lit_base16 * 16 + oo, ol))
        atomicSub(&b->st_lit, 1u);
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------------------------------------------- chain
__global__ void __launch_bounds__(64) k_zchain(ZArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nzb) return;
  const ZBlob z = a.zb[i];
  if (z.state) return;
  const BlobDesc d = a.blobs[z.blob];
  if (!chain_blob(a.blocks + z.slot0, z.n_blocks, (uint32_t)d.dst_cap)) a.zb[i].state = 1;
}

// ---------------------------------------------------------------------------------------------------------------- exec
// One CTA per blob, blocks in order.  A block's sequences are executed in GROUPS: the next <= NT sequences whose
// output spans <= kGroupBytes.  The group's output is assembled in shared memory:
//   stage    the group's literals (one contiguous range of the block's literal source) come in with coalesced loads;
//   phase 1  one lane per sequence, all warps in parallel: literal runs, and every match whose source lies entirely
//            BEFORE the group (final bytes, read from global memory);
//   phase 2  matches that read bytes of the group itself.  Warps take turns in sequence order (a token in shared
//            memory); inside a warp, rounds: every pending lane whose source ends below the first pending lane's
//            destination copies (those bytes are final), the rest wait for the next round.  Measured on python
//            sources: 40 % of the matches are of this kind, 3.6 rounds per warp turn;
//   flush    one coalesced copy to global memory.
// A sequence longer than the group buffer is executed alone, straight in global memory, by the whole team.
constexpr uint32_t kGroupBytes = 16384;
constexpr uint32_t kLaneFar = 64;   // far matches up to this length are copied by their own lane (memory latency bound)
constexpr uint32_t kLaneNear = 32;  // near matches up to this length likewise (two 16-byte chunks); longer: whole warp

template <int NT>
struct ExecShared {
  alignas(16) uint8_t buf[kGroupBytes + 16];
  alignas(16) uint8_t lits[kGroupBytes + 32];
  uint32_t pat[kPatWords];
  uint32_t wcnt[NT / 32];
  SeqRec16 big;
  uint32_t gend, lit_lo, lit_hi;
  uint32_t token;
  uint32_t err;
  uint32_t item;
};

// byte p (block-relative, may be negative = earlier blocks) of the blob's output while group [gpos, ..) is assembled
#define ZN_SRC_BYTE(p) ((p) >= (int32_t)gpos ? sh->buf[(p) - (int32_t)gpos] : gout[(p)])

template <int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 6) k_zexec(ZArgs a, uint8_t* out_base, uint32_t* produced, uint32_t* work_counter) {
  extern __shared__ __align__(16) uint8_t exec_smem[];
  ExecShared<NT>* sh = reinterpret_cast<ExecShared<NT>*>(exec_smem);
  const Team t{threadIdx.x, (uint32_t)NT};
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  for (;;) {
    if (tid == 0) sh->item = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t item = sh->item;
    __syncthreads();
    if (item >= a.nzb) break;
    const ZBlob z = a.zb[item];
    if (z.state) continue;
    const BlobDesc d = a.blobs[z.blob];
    const uint8_t* src = a.blobs_base + d.src_off;
    uint8_t* out = out_base + d.dst_off;
    if (tid == 0) sh->err = 0;
    __syncthreads();
    bool bad = false;
    for (uint32_t j = 0; j < z.n_blocks && !bad; j++) {
      const ZBlock* b = &a.blocks[z.slot0 + j];
      const uint32_t flags = b->flags, type = flags & ZB_TYPE_MASK;
      uint8_t* gout = out + b->out_base;
      if (type == 0) { team_copy(t, gout, src + b->src_off, b->len); __syncthreads(); continue; }
      if (type == 1) { team_fill(t, gout, src[b->src_off], b->len); __syncthreads(); continue; }
      const uint32_t lt = (flags >> ZB_LIT_SHIFT) & 3u;
      const uint8_t* lit = lt == 0 ? src + b->lit_off : a.lits + (size_t)b->lit_base16 * 16;
      const int rle = lt == 1 ? (int)b->lit_off : -1;
      const SeqRec16* seqs = a.recs + b->seq_base;
      const uint32_t nseq = b->nseq, out_base_blk = b->out_base, frame_start = b->frame_start;
      const uint32_t r0 = b->rep_in[0], r1 = b->rep_in[1], r2 = b->rep_in[2];
      uint32_t s0 = 0, gpos = 0;
      while (s0 < nseq) {
        // ---- the group: leading sequences that fit the buffer
        const bool have = s0 + tid < nseq;
        SeqRec16 r;
        r.w0 = r.w1 = r.w2 = r.w3 = 0;
        if (have) r = seqs[s0 + tid];
        const uint32_t orl = rec_out(r), ll = rec_ll(r), lr = rec_lit(r), ml = rec_ml(r);
        const uint32_t endp = orl + ll + ml;
        const bool in = have && endp - gpos <= kGroupBytes;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, in);
        if (lane == 0) sh->wcnt[warp] = bal == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~bal) - 1u;
        if (tid == 0) { sh->big = r; sh->token = 0; }
        __syncthreads();
        uint32_t count = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) {
          const uint32_t c = sh->wcnt[w];
          if (count == (uint32_t)w * 32u) count += c;
        }
        if (count == 0) {
          // ---- a sequence longer than the group buffer: alone, in global memory, by the whole team
          const SeqRec16 q = sh->big;
          const uint32_t qo = rec_out(q), qll = rec_ll(q), qlr = rec_lit(q), qml = rec_ml(q);
          if (qll) {
            if (rle >= 0) team_fill(t, gout + qo, (uint32_t)rle, qll);
            else team_copy(t, gout + qo, lit + qlr, qll);
          }
          const uint32_t off = sym_resolve(rec_off(q), r0, r1, r2);
          const uint32_t dabs = out_base_blk + qo + qll;
          if (qml && (off == 0 || off > dabs - frame_start)) { bad = true; break; }
          __syncthreads();
          if (qml) team_match(t, gout + qo + qll, off, qml, sh->pat, nullptr);
          __syncthreads();
          gpos = qo + qll + qml;
          s0 += 1;
          continue;
        }
        if (tid == count - 1) { sh->gend = endp; sh->lit_hi = lr + ll; }
        if (tid == 0) sh->lit_lo = lr;
        __syncthreads();
        const uint32_t gend = sh->gend, lit_lo = sh->lit_lo, lit_n = sh->lit_hi - lit_lo;
        if (rle < 0 && lit_n) team_copy(t, sh->lits, lit + lit_lo, lit_n);
        __syncthreads();
        // ---- phase 1
        const bool mine = tid < count;
        bool pending = false, far = false;
        int32_t dst_rel = 0, src_rel = 0, src_end = 0;
        uint32_t off = 1;
        if (mine) {
          uint8_t* o = sh->buf + (orl - gpos);
          if (ll <= kLaneFar) {
            if (rle >= 0) for (uint32_t i = 0; i < ll; i++) o[i] = (uint8_t)rle;
            else { const uint8_t* ls = sh->lits + (lr - lit_lo); for (uint32_t i = 0; i < ll; i++) o[i] = ls[i]; }
          }
          if (ml) {
            off = sym_resolve(r.w0, r0, r1, r2);
            const uint32_t dabs = out_base_blk + orl + ll;
            if (off == 0 || off > dabs - frame_start) sh->err = 1;
            else {
              dst_rel = (int32_t)(orl + ll);
              src_rel = dst_rel - (int32_t)off;
              src_end = off >= ml ? src_rel + (int32_t)ml : dst_rel;
              // far: the whole source range lies before the group and the match does not feed itself
              far = off >= ml && src_rel + (int32_t)ml <= (int32_t)gpos;
              pending = !far;
            }
          }
        }
        if (far && ml <= kLaneFar) {  // source entirely before the group: final bytes in global memory
          const uint8_t* s = gout + src_rel;
          uint8_t* o = sh->buf + (dst_rel - (int32_t)gpos);
          for (uint32_t c = 0; c < ml; c += 16) {
            uint8_t v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) if (c + i < ml) v[i] = s[c + i];
#pragma unroll
            for (int i = 0; i < 16; i++) if (c + i < ml) o[c + i] = v[i];
          }
        }
        {  // long literal runs and long far matches: whole warp per copy
          uint32_t m = __ballot_sync(0xFFFFFFFFu, mine && ll > kLaneFar);
          while (m) {
            const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            const uint32_t dd = __shfl_sync(0xFFFFFFFFu, orl, sl) - gpos, lr2 = __shfl_sync(0xFFFFFFFFu, lr, sl) - lit_lo,
                           l = __shfl_sync(0xFFFFFFFFu, ll, sl);
            for (uint32_t k = lane; k < l; k += 32) sh->buf[dd + k] = rle >= 0 ? (uint8_t)rle : sh->lits[lr2 + k];
          }
          m = __ballot_sync(0xFFFFFFFFu, far && ml > kLaneFar);
          while (m) {
            const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            const int32_t dd = __shfl_sync(0xFFFFFFFFu, dst_rel, sl), ss = __shfl_sync(0xFFFFFFFFu, src_rel, sl);
            const uint32_t l = __shfl_sync(0xFFFFFFFFu, ml, sl);
            for (uint32_t k = lane; k < l; k += 32) sh->buf[dd - (int32_t)gpos + (int32_t)k] = gout[ss + (int32_t)k];
          }
        }
        __syncthreads();
        if (sh->err) { bad = true; break; }
        // ---- phase 2: warps in sequence order
        if (warp * 32u < count) {
          if (lane == 0) while (*(volatile uint32_t*)&sh->token != warp) {}
          __syncwarp();
          __threadfence_block();
          for (;;) {
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pending);
            if (!pm) break;
            const uint32_t f = (uint32_t)__ffs((int)pm) - 1u;
            const int32_t dstf = __shfl_sync(0xFFFFFFFFu, dst_rel, f);
            const bool ready = pending && (lane == f || src_end <= dstf);
            if (ready && ml <= kLaneNear) {
              uint8_t* o = sh->buf + (dst_rel - (int32_t)gpos);
              if (off >= 16 && src_rel >= (int32_t)gpos) {  // 16 loads in flight, then 16 stores
                const uint8_t* s = sh->buf + (src_rel - (int32_t)gpos);
                for (uint32_t c = 0; c < ml; c += 16) {
                  uint8_t v[16];
#pragma unroll
                  for (int i = 0; i < 16; i++) if (c + i < ml) v[i] = s[c + i];
#pragma unroll
                  for (int i = 0; i < 16; i++) if (c + i < ml) o[c + i] = v[i];
                }
              } else {  // short distance (bytes feed later bytes) or a source that starts before the group
                for (uint32_t i = 0; i < ml; i++) {
                  const int32_t p = src_rel + (int32_t)i;
                  o[i] = ZN_SRC_BYTE(p);
                }
              }
            }
            uint32_t m = __ballot_sync(0xFFFFFFFFu, ready && ml > kLaneNear);
            while (m) {  // long match: byte k reads window[k mod off], which existed before the match began
              const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
              m &= m - 1u;
              const int32_t dd = __shfl_sync(0xFFFFFFFFu, dst_rel, sl);
              const uint32_t oo = __shfl_sync(0xFFFFFFFFu, off, sl), l = __shfl_sync(0xFFFFFFFFu, ml, sl);
              for (uint32_t k = lane; k < l; k += 32) {
                const int32_t p = dd - (int32_t)oo + (int32_t)(oo >= l ? k : k % oo);
                sh->buf[dd - (int32_t)gpos + (int32_t)k] = ZN_SRC_BYTE(p);
              }
            }
            pending = pending && !ready;
            __syncwarp();
          }
          __threadfence_block();
          __syncwarp();
          if (lane == 0) *(volatile uint32_t*)&sh->token = warp + 1u;
        }
        __syncthreads();
        // ---- flush
        team_copy(t, gout + gpos, sh->buf, gend - gpos);
        __syncthreads();
        gpos = gend;
        s0 += count;
      }
      if (bad) break;
      const uint32_t rest = b->lit_regen - b->lit_used;
      if (rest) {
        if (rle >= 0) team_fill(t, gout + b->matched, (uint32_t)rle, rest);
        else team_copy(t, gout + b->matched, lit + b->lit_used, rest);
      }
      __syncthreads();
    }
    if (tid == 0) {
      if (bad) a.zb[item].state = 1;
      else produced[z.blob] = (uint32_t)d.dst_cap;
    }
  }
}
#undef ZN_SRC_BYTE

}  // namespace zp
}  // namespace zn
