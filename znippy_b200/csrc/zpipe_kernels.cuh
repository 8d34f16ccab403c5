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
// One CTA per blob, blocks in order.  A block's sequences are executed in GROUPS (the next <= NT-32 sequences whose
// output spans <= kGroupBytes) through a ring buffer in shared memory that holds the last 32 KiB of the blob's output.
// The only serial dependence of LZ77 execution — a match that reads bytes produced by matches just before it — is
// isolated in ONE warp; everything else is prepared around it by the other warps:
//
//   producers (warps 1..)  for group g: sequence records (prefetched a group ahead), the group's literals (staged with
//                          coalesced loads), one lane per sequence: literal runs and every FAR match (source entirely
//                          below the start of group g-1, i.e. already flushed: read from global memory) into the ring,
//                          and a compacted list of the NEAR matches; then, once the executor is done with group g-1,
//                          one coalesced flush of that group from the ring to global memory;
//   executor (warp 0)      near matches of group g, 32 per turn, in rounds: every pending lane whose source ends below
//                          the first pending lane's destination copies (those bytes are final), the rest wait a round.
//                          Measured on python sources: 40 % of the matches are near, ~1 100 rounds per 128 KiB block.
//
// Producers and executor meet at hardware named barriers (bar.arrive / bar.sync: no polling, waiting warps cost no
// issue slots): READY[g & 1] (group g prepared) and DONE[g & 1] (its near matches executed).  The executor works on
// group g while the producers prepare g+1 and flush g-1; three groups (<= 24 KiB) are live in the 32 KiB ring.
// A sequence longer than a group, raw / RLE blocks and a block's trailing literals are written straight to global
// memory by the producer team after draining the pipeline.
constexpr uint32_t kRing = 32768, kRingMask = kRing - 1;
constexpr uint32_t kGroupBytes = 8192;
constexpr uint32_t kLitWin = 8192;  // staged literal window
constexpr uint32_t kLaneFar = 64;   // far matches up to this length are copied by their own lane (memory latency bound)
constexpr uint32_t kLaneNear = 32;  // near matches up to this length likewise (two 16-byte chunks); longer: whole warp
constexpr uint32_t kNearEnd = 0xFFFFFFFFu;

template <int NT>
struct ExecShared {
  alignas(16) uint8_t ring[kRing];
  alignas(16) uint8_t lits[kLitWin + 32];
  uint32_t n_off[2][NT - 32], n_dst[2][NT - 32], n_ml[2][NT - 32];  // near lists of the two groups in flight
  uint32_t gi_n[2], gi_lo[2];  // entries; lowest position still valid in the ring for that group's sources
  uint32_t pat[kPatWords];
  uint32_t wcnt[NT / 32], wnear[NT / 32];
  SeqRec16 big;
  uint32_t gend, lit_lo, lit_hi;
  uint32_t err, item;
};

ZN_D void bar_sync_n(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
ZN_D void bar_arrive_n(uint32_t id, uint32_t n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr uint32_t kBarProd = 1, kBarReady = 2, kBarDone = 4;  // named barriers: 1, 2-3, 4-5

// ring -> global: output bytes [lo, hi) of the blob
ZN_D void ring_flush(const Team& t, uint8_t* out, const uint8_t* ring, uint32_t lo, uint32_t hi) {
  if (hi <= lo) return;
  const uint32_t r0 = lo & kRingMask, n = hi - lo;
  if (r0 + n <= kRing) team_copy(t, out + lo, ring + r0, n);
  else {
    const uint32_t n0 = kRing - r0;
    team_copy(t, out + lo, ring + r0, n0);
    team_copy(t, out + lo + n0, ring, n - n0);
  }
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 6) k_zexec(ZArgs a, uint8_t* out_base, uint32_t* produced, uint32_t* work_counter) {
  extern __shared__ __align__(16) uint8_t exec_smem[];
  ExecShared<NT>* sh = reinterpret_cast<ExecShared<NT>*>(exec_smem);
  constexpr uint32_t NP = NT - 32;  // producer threads
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  for (;;) {
    if (tid == 0) sh->item = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t item = sh->item;
    if (tid == 0) sh->err = 0;
    __syncthreads();
    if (item >= a.nzb) break;
    const ZBlob z = a.zb[item];
    if (z.state) continue;
    const BlobDesc d = a.blobs[z.blob];
    const uint8_t* src = a.blobs_base + d.src_off;
    uint8_t* out = out_base + d.dst_off;
    if (warp == 0) {
      // ================================================================ executor: near matches, group after group
      for (uint32_t g = 0;; g++) {
        const uint32_t par = g & 1u;
        bar_sync_n(kBarReady + par, NT);
        const uint32_t n = sh->gi_n[par], vlo = sh->gi_lo[par];
        if (n == kNearEnd) break;
        for (uint32_t base = 0; base < n; base += 32) {
          const uint32_t i = base + lane;
          bool pending = i < n;
          uint32_t off = 1, dst = 0, ml = 0;
          if (pending) { off = sh->n_off[par][i]; dst = sh->n_dst[par][i]; ml = sh->n_ml[par][i]; }
          const uint32_t srcp = dst - off;
          const uint32_t src_end = off >= ml ? srcp + ml : dst;
          for (;;) {
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pending);
            if (!pm) break;
            const uint32_t f = (uint32_t)__ffs((int)pm) - 1u;
            const uint32_t dstf = __shfl_sync(0xFFFFFFFFu, dst, f);
            const bool ready = pending && (lane == f || src_end <= dstf);
            if (ready && ml <= kLaneNear) {
              const uint32_t ro = dst & kRingMask, rs = srcp & kRingMask;
              if (off >= 16 && srcp >= vlo && ro + ml <= kRing && rs + ml <= kRing) {  // 16 loads in flight, then 16 stores
                const uint8_t* s = sh->ring + rs;
                uint8_t* o = sh->ring + ro;
                for (uint32_t c = 0; c < ml; c += 16) {
                  uint8_t v[16];
#pragma unroll
                  for (int k = 0; k < 16; k++) if (c + k < ml) v[k] = s[c + k];
#pragma unroll
                  for (int k = 0; k < 16; k++) if (c + k < ml) o[c + k] = v[k];
                }
              } else {  // short distance (bytes feed later bytes), ring wrap, or a source that starts below the ring
                for (uint32_t k = 0; k < ml; k++) {
                  const uint32_t p = srcp + k;
                  sh->ring[(dst + k) & kRingMask] = p >= vlo ? sh->ring[p & kRingMask] : out[p];
                }
              }
            }
            uint32_t m = __ballot_sync(0xFFFFFFFFu, ready && ml > kLaneNear);
            while (m) {  // long match: byte k reads window[k mod off], which existed before the match began
              const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
              m &= m - 1u;
              const uint32_t dd = __shfl_sync(0xFFFFFFFFu, dst, sl), oo = __shfl_sync(0xFFFFFFFFu, off, sl),
                             l = __shfl_sync(0xFFFFFFFFu, ml, sl);
              for (uint32_t k = lane; k < l; k += 32) {
                const uint32_t p = dd - oo + (oo >= l ? k : k % oo);
                sh->ring[(dd + k) & kRingMask] = p >= vlo ? sh->ring[p & kRingMask] : out[p];
              }
            }
            pending = pending && !ready;
            __syncwarp();
          }
        }
        __threadfence_block();
        bar_arrive_n(kBarDone + par, NT);
      }
    } else {
      // ================================================================ producers
      const uint32_t pt = tid - 32, pw = warp - 1;
      const Team t{pt, NP, kBarProd};
      uint32_t g = 0;              // groups handed to the executor so far
      bool pend = false;           // group g-1 is with the executor / not flushed yet
      uint32_t pend_lo = 0, pend_hi = 0;
      uint32_t ring_lo = 0;        // lowest output position whose bytes are valid in the ring for the NEXT group's near matches
      bool bad = false;
      // waits for the executor to finish the group in flight and flushes it: afterwards everything produced so far is
      // in global memory and visible to the producer team
      auto drain = [&]() {
        if (pend) {
          bar_sync_n(kBarDone + ((g - 1) & 1u), NT);
          ring_flush(t, out, sh->ring, pend_lo, pend_hi);
          pend = false;
        }
        bar_sync_n(kBarProd, NP);
      };
      for (uint32_t j = 0; j < z.n_blocks && !bad; j++) {
        const ZBlock* b = &a.blocks[z.slot0 + j];
        const uint32_t flags = b->flags, type = flags & ZB_TYPE_MASK;
        const uint32_t blk0 = b->out_base;
        if (type != 2) {
          drain();
          if (type == 0) team_copy(t, out + blk0, src + b->src_off, b->len);
          else team_fill(t, out + blk0, src[b->src_off], b->len);
          bar_sync_n(kBarProd, NP);
          ring_lo = blk0 + b->len;
          continue;
        }
        const uint32_t lt = (flags >> ZB_LIT_SHIFT) & 3u;
        const uint8_t* lit = lt == 0 ? src + b->lit_off : a.lits + (size_t)b->lit_base16 * 16;
        const int rle = lt == 1 ? (int)b->lit_off : -1;
        const SeqRec16* seqs = a.recs + b->seq_base;
        const uint32_t nseq = b->nseq, frame_start = b->frame_start;
        const uint32_t r0 = b->rep_in[0], r1 = b->rep_in[1], r2 = b->rep_in[2];
        uint32_t s0 = 0, gpos = blk0;       // gpos: blob-absolute output position where the next group starts
        uint32_t lw_lo = 0, lw_hi = 0;      // literal window staged in sh->lits: literals [lw_lo, lw_hi) of this block
        SeqRec16 rn;                        // prefetched record of sequence s0 + pt
        rn.w0 = rn.w1 = rn.w2 = rn.w3 = 0;
        if (pt < nseq) rn = seqs[pt];
        while (s0 < nseq) {
          // ---- the group: leading sequences that fit
          const bool have = s0 + pt < nseq;
          const SeqRec16 r = rn;
          const uint32_t orl = rec_out(r), ll = rec_ll(r), lr = rec_lit(r), ml = rec_ml(r);
          const uint32_t endp = blk0 + orl + ll + ml;
          const bool in = have && endp - gpos <= kGroupBytes;
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, in);
          if (lane == 0) sh->wcnt[pw] = bal == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs((int)~bal) - 1u;
          if (pt == 0) sh->big = r;
          bar_sync_n(kBarProd, NP);
          uint32_t count = 0;
#pragma unroll
          for (int w = 0; w < (int)(NP / 32); w++) {
            const uint32_t c = sh->wcnt[w];
            if (count == (uint32_t)w * 32u) count += c;
          }
          if (count == 0) {
            // ---- a sequence longer than a group: alone, in global memory, by the producer team
            drain();
            const SeqRec16 q = sh->big;
            const uint32_t qo = blk0 + rec_out(q), qll = rec_ll(q), qlr = rec_lit(q), qml = rec_ml(q);
            if (qll) {
              if (rle >= 0) team_fill(t, out + qo, (uint32_t)rle, qll);
              else team_copy(t, out + qo, lit + qlr, qll);
            }
            const uint32_t off = sym_resolve(rec_off(q), r0, r1, r2);
            const uint32_t dabs = qo + qll;
            if (qml && (off == 0 || off > dabs - frame_start)) { bad = true; break; }
            bar_sync_n(kBarProd, NP);
            if (qml) team_match(t, out + dabs, off, qml, sh->pat, nullptr);
            bar_sync_n(kBarProd, NP);
            gpos = dabs + qml;
            ring_lo = gpos;
            s0 += 1;
            rn.w0 = rn.w1 = rn.w2 = rn.w3 = 0;
            if (s0 + pt < nseq) rn = seqs[s0 + pt];
            continue;
          }
          // prefetch the next group's records: their latency hides behind this group's work
          rn.w0 = rn.w1 = rn.w2 = rn.w3 = 0;
          if (s0 + count + pt < nseq) rn = seqs[s0 + count + pt];
          if (pt == count - 1) { sh->gend = endp; sh->lit_hi = lr + ll; }
          if (pt == 0) sh->lit_lo = lr;
          bar_sync_n(kBarProd, NP);
          const uint32_t gend = sh->gend, lit_lo = sh->lit_lo, lit_hi = sh->lit_hi;
          if (rle < 0 && lit_hi > lit_lo && (lit_lo < lw_lo || lit_hi > lw_hi)) {  // (re)stage the literal window
            lw_lo = lit_lo;
            lw_hi = min(b->lit_regen, lit_lo + kLitWin);
            bar_sync_n(kBarProd, NP);  // nobody still reads the old window
            team_copy(t, sh->lits, lit + lw_lo, lw_hi - lw_lo);
            bar_sync_n(kBarProd, NP);
          }
          // ---- phase 1: literal runs and far matches into the ring; near matches into the list
          const uint32_t par = g & 1u;
          const bool mine = pt < count;
          bool near_m = false, far = false;
          uint32_t dabs = 0, sabs = 0, off = 1;
          if (mine) {
            const uint32_t oabs = blk0 + orl;
            if (ll <= kLaneFar) {
              if (rle >= 0) for (uint32_t i = 0; i < ll; i++) sh->ring[(oabs + i) & kRingMask] = (uint8_t)rle;
              else { const uint8_t* ls = sh->lits + (lr - lw_lo); for (uint32_t i = 0; i < ll; i++) sh->ring[(oabs + i) & kRingMask] = ls[i]; }
            }
            if (ml) {
              off = sym_resolve(r.w0, r0, r1, r2);
              dabs = oabs + ll;
              if (off == 0 || off > dabs - frame_start) sh->err = 1;
              else {
                sabs = dabs - off;
                // far: the whole source lies below the ring's valid range (flushed, final) and the match does not feed itself
                far = off >= ml && sabs + ml <= ring_lo;
                near_m = !far;
              }
            }
          }
          if (far && ml <= kLaneFar) {
            const uint8_t* s = out + sabs;
            const uint32_t ro = dabs & kRingMask;
            if (ro + ml <= kRing) {
              uint8_t* o = sh->ring + ro;
              for (uint32_t c = 0; c < ml; c += 16) {
                uint8_t v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) if (c + i < ml) v[i] = s[c + i];
#pragma unroll
                for (int i = 0; i < 16; i++) if (c + i < ml) o[c + i] = v[i];
              }
            } else for (uint32_t i = 0; i < ml; i++) sh->ring[(dabs + i) & kRingMask] = s[i];
          }
          {  // long literal runs and long far matches: whole warp per copy
            uint32_t m = __ballot_sync(0xFFFFFFFFu, mine && ll > kLaneFar);
            while (m) {
              const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
              m &= m - 1u;
              const uint32_t dd = blk0 + __shfl_sync(0xFFFFFFFFu, orl, sl), lr2 = __shfl_sync(0xFFFFFFFFu, lr, sl) - lw_lo,
                             l = __shfl_sync(0xFFFFFFFFu, ll, sl);
              for (uint32_t k = lane; k < l; k += 32) sh->ring[(dd + k) & kRingMask] = rle >= 0 ? (uint8_t)rle : sh->lits[lr2 + k];
            }
            m = __ballot_sync(0xFFFFFFFFu, far && ml > kLaneFar);
            while (m) {
              const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
              m &= m - 1u;
              const uint32_t dd = __shfl_sync(0xFFFFFFFFu, dabs, sl), ss = __shfl_sync(0xFFFFFFFFu, sabs, sl),
                             l = __shfl_sync(0xFFFFFFFFu, ml, sl);
              for (uint32_t k = lane; k < l; k += 32) sh->ring[(dd + k) & kRingMask] = out[ss + k];
            }
          }
          // near list, in sequence order
          const uint32_t nb = __ballot_sync(0xFFFFFFFFu, near_m);
          if (lane == 0) sh->wnear[pw] = (uint32_t)__popc(nb);
          bar_sync_n(kBarProd, NP);
          if (sh->err) { bad = true; break; }
          uint32_t before = 0, total = 0;
#pragma unroll
          for (int w = 0; w < (int)(NP / 32); w++) {
            const uint32_t c = sh->wnear[w];
            if ((uint32_t)w < pw) before += c;
            total += c;
          }
          if (near_m) {
            const uint32_t slot = before + (uint32_t)__popc(nb & ((1u << lane) - 1u));
            sh->n_off[par][slot] = off;
            sh->n_dst[par][slot] = dabs;
            sh->n_ml[par][slot] = ml;
          }
          if (pt == 0) { sh->gi_n[par] = total; sh->gi_lo[par] = ring_lo; }
          __threadfence_block();
          bar_arrive_n(kBarReady + par, NT);
          // ---- the previous group is done once the executor says so: flush it
          if (pend) {
            bar_sync_n(kBarDone + ((g - 1) & 1u), NT);
            ring_flush(t, out, sh->ring, pend_lo, pend_hi);
          }
          bar_sync_n(kBarProd, NP);
          // far sources of the next group may reach up to the start of THIS group's predecessor... which is what was
          // just flushed: everything below gpos is now final in global memory except this group itself
          ring_lo = gpos;
          pend = true;
          pend_lo = gpos;
          pend_hi = gend;
          gpos = gend;
          s0 += count;
          g++;
        }
        if (bad) break;
        const uint32_t rest = b->lit_regen - b->lit_used;
        if (rest) {
          drain();
          if (rle >= 0) team_fill(t, out + blk0 + b->matched, (uint32_t)rle, rest);
          else team_copy(t, out + blk0 + b->matched, lit + b->lit_used, rest);
          bar_sync_n(kBarProd, NP);
          ring_lo = blk0 + b->matched + rest;
        }
      }
      drain();
      if (pt == 0) { sh->gi_n[g & 1u] = kNearEnd; sh->gi_lo[g & 1u] = 0; }
      __threadfence_block();
      bar_arrive_n(kBarReady + (g & 1u), NT);
      if (pt == 0) {
        if (bad) a.zb[item].state = 1;
        else produced[z.blob] = (uint32_t)d.dst_cap;
      }
    }
  }
}

}  // namespace zp
}  // namespace zn
