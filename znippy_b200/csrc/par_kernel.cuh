// k_decode_par: the block-parallel decode kernel for large blobs (see zstd_par.cuh for the pipeline).
#pragma once
#include "decode_kernels.cuh"
#include "zstd_par.cuh"

namespace zn {
namespace par {

constexpr size_t kParSeqBytes = (size_t)kGroup * kMaxSeq * sizeof(SeqOut);
constexpr size_t kParLitBytes = (size_t)kGroup * kLitStride;
constexpr size_t kParScratchPerCta = kParSeqBytes + kParLitBytes;

// dynamic shared memory layout of one CTA
struct ParShared {
  DecShared slot[kGroup];  // slot j: tables of the block j in flight; slot 0 also serves the team executor / LZ4 path
  BlockRec recs[kGroup];
  GroupInfo gi;
  uint32_t predef[kGroup];
  uint32_t item, status, pos_after, err, done;
  alignas(128) uint8_t tile[kTileBytes];
  alignas(16) uint8_t src[kSrcStage + 32];
};

static_assert(sizeof(ParShared) <= 227 * 1024, "ParShared must fit one SM's shared memory");

// Lane-parallel execution of one block of short sequences.  Team-uniform entry; returns S_OK / S_DECODE_ERROR.
//
// The block is assembled IN SHARED MEMORY: during phase B nobody needs the slots' entropy tables, so their 148 KB
// hold the block's (<= 128 KiB) output.  Every dependence between the block's sequences is then served at
// shared-memory latency instead of a round trip through L2, bytes that precede the block come from global memory
// (final since the previous block), and the finished block leaves with one coalesced copy.
ZN_D uint32_t exec_block_lanes(const Team& t, ParShared* ps, const BlockRec& r, const SeqOut* seqs, uint8_t* out,
                               const uint8_t* lit, uint32_t frame_start) {
  static_assert(sizeof(ps->slot) >= kZstdBlockMax + 16, "slots must be able to hold one decoded block");
  uint8_t* obuf = reinterpret_cast<uint8_t*>(ps->slot);
  const uint32_t warp = t.tid >> 5, lane = t.tid & 31u, nw = t.n >> 5;
  const uint32_t base = r.base_out, nseq = r.nseq, nb = (nseq + 31u) >> 5;
  const uint32_t dec = r.matched + (r.lit_len - r.lit_used);
  const int rle = r.lit_rle;
  const uint8_t* gout = out + base;  // gout[p] for p < 0: bytes of earlier blocks
  if (t.tid == 0) ps->err = 0;
  team_sync(t);
  // ---- pass 1: literals (no dependences) + offset validation
  for (uint32_t b = warp; b < nb; b += nw) {
    const uint32_t s = b * 32u + lane;
    const bool act = s < nseq;
    SeqOut q;
    q.out_rel = q.lit_rel = q.ll = q.ml = 0; q.off = 1;
    if (act) q = seqs[s];
    if (act && q.ml && (q.off == 0 || q.off > base + q.out_rel + q.ll - frame_start)) ps->err = 1;
    if (act && q.ll <= kLaneMax) {
      if (rle >= 0) for (uint32_t i = 0; i < q.ll; i++) obuf[q.out_rel + i] = (uint8_t)rle;
      else for (uint32_t i = 0; i < q.ll; i++) obuf[q.out_rel + i] = lit[q.lit_rel + i];
    }
    uint32_t m = __ballot_sync(0xFFFFFFFFu, act && q.ll > kLaneMax);
    while (m) {
      const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
      m &= m - 1u;
      const uint32_t dd = __shfl_sync(0xFFFFFFFFu, q.out_rel, sl), lr = __shfl_sync(0xFFFFFFFFu, q.lit_rel, sl),
                     l = __shfl_sync(0xFFFFFFFFu, q.ll, sl);
      for (uint32_t k = lane; k < l; k += 32) obuf[dd + k] = rle >= 0 ? (uint8_t)rle : lit[lr + k];
    }
  }
  for (uint32_t k = t.tid; k < r.lit_len - r.lit_used; k += t.n)
    obuf[r.matched + k] = rle >= 0 ? (uint8_t)rle : lit[r.lit_used + k];
  team_sync(t);
  ZN_TP(34);
  uint32_t rc = ps->err ? S_DECODE_ERROR : S_OK;
  // ---- pass 2: matches, one lane per sequence, 32 sequences per batch, batches spread over the warps.
  // Readiness is tracked exactly: a lane may copy as soon as every EARLIER BATCH that produced bytes of its source
  // range is flagged done (bdone[]), and the part of its source inside its own batch lies below the first unfinished
  // lane's match.  Far matches therefore never wait, and a batch only waits for the few lanes that really depend on
  // its predecessors.  bstart[c] = block-relative output position where batch c begins.
  // both tables live in the bulk-store tile, which is idle here (the caller drained bulk stores with mem_sync); the
  // staged-source buffer ps->src must NOT be touched: the walker of the next group still parses from it
  constexpr uint32_t kMaxBatches = kMaxSeq / 32;
  uint32_t* bstart = reinterpret_cast<uint32_t*>(ps->tile);                                      // nb + 1 entries
  volatile uint8_t* bdone = reinterpret_cast<volatile uint8_t*>(ps->tile + (kMaxBatches + 2) * 4);  // nb flags
  static_assert(kTileBytes >= (kMaxBatches + 2) * 4 + kMaxBatches + 1, "tile too small for the batch tables");
  for (uint32_t c = t.tid; c <= nb; c += t.n) {
    bstart[c] = c < nb ? seqs[c * 32u].out_rel : r.matched;
    if (c < nb) bdone[c] = 0;
  }
  team_sync(t);
  for (uint32_t b = warp; rc == S_OK && b < nb; b += nw) {
    const uint32_t s = b * 32u + lane;
    const bool act = s < nseq;
    SeqOut q;
    q.out_rel = q.lit_rel = q.ll = q.ml = 0; q.off = 1;
    if (act) q = seqs[s];
    const int32_t dst = (int32_t)(q.out_rel + q.ll);                  // block-relative
    const int32_t src = dst - (int32_t)q.off;                         // may be negative: earlier blocks (final)
    const int32_t src_end = q.off >= q.ml ? src + (int32_t)q.ml : dst;
    bool pending = act && q.ml > 0;
    // producer batches of the in-block part of the source: [c_lo, c_hi], restricted to batches before this one
    uint32_t c_lo = 1, c_hi = 0;  // empty
    if (pending && src_end > 0) {
      const uint32_t plo = src > 0 ? (uint32_t)src : 0u, phi = (uint32_t)(src_end - 1);
      uint32_t lo = 0, hi = b;  // batches > b cannot hold bytes below dst
      while (hi > lo) { const uint32_t mid = (lo + hi + 1) >> 1; if (bstart[mid] <= plo) lo = mid; else hi = mid - 1; }
      c_lo = lo;
      lo = c_lo; hi = b;
      while (hi > lo) { const uint32_t mid = (lo + hi + 1) >> 1; if (bstart[mid] <= phi) lo = mid; else hi = mid - 1; }
      c_hi = lo;
    }
    const bool in_batch = pending && c_hi == b && c_lo <= c_hi;       // part of the source lies in this batch
    const uint32_t c_end = c_hi == b ? b - 1u : c_hi;                  // (b == 0 and c_hi == 0: wraps; guarded below)
    // earlier lanes of this batch whose match output [dst_j, dst_j + ml_j) overlaps my source range (computed once)
    uint32_t dep_mask = 0;
    if (__any_sync(0xFFFFFFFFu, in_batch)) {
#pragma unroll 4
      for (uint32_t j = 0; j < 31; j++) {
        const int32_t dj = __shfl_sync(0xFFFFFFFFu, dst, j);
        const int32_t ej = dj + (int32_t)__shfl_sync(0xFFFFFFFFu, q.ml, j);
        if (in_batch && j < lane && dj < src_end && ej > src) dep_mask |= 1u << j;
      }
    }
    uint32_t pm;
    ZN_CNT(0, 1);
    const long long t_b0 = clock64();
    while ((pm = __ballot_sync(0xFFFFFFFFu, pending)) != 0) {
      ZN_CNT(1, 1);
      bool ext_ok = true;
      if (pending && c_lo <= c_hi && !(c_hi == b && b == 0) && c_lo <= c_end) {
        while (c_lo <= c_end && bdone[c_lo]) c_lo++;
        ext_ok = c_lo > c_end;
      }
      // inside the batch a lane waits only for PENDING earlier lanes whose match output overlaps its source
      const bool ready = pending && ext_ok && (dep_mask & pm) == 0;
      if (!__any_sync(0xFFFFFFFFu, ready)) {
        ZN_CNT(2, 1);
        __nanosleep(20);
        continue;
      }
      { const uint32_t nr__ = __popc(__ballot_sync(0xFFFFFFFFu, ready)); ZN_CNT(3, nr__); (void)nr__; }
      __threadfence_block();  // order the reads below after the observation of bdone[]
      if (ready && q.ml <= kLaneMax) {
        uint32_t i = 0;
        if (q.off >= 4) {  // 4 bytes per step: the loads of a step do not depend on its stores (distance >= 4)
          for (; i + 4 <= q.ml; i += 4) {
            const int32_t p = src + (int32_t)i;
            uint8_t v0, v1, v2, v3;
            if (p >= 0) { v0 = obuf[p]; v1 = obuf[p + 1]; v2 = obuf[p + 2]; v3 = obuf[p + 3]; }
            else {
              v0 = gout[p]; v1 = p + 1 >= 0 ? obuf[p + 1] : gout[p + 1]; v2 = p + 2 >= 0 ? obuf[p + 2] : gout[p + 2];
              v3 = p + 3 >= 0 ? obuf[p + 3] : gout[p + 3];
            }
            uint8_t* o = obuf + dst + (int32_t)i;
            o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
          }
        }
        for (; i < q.ml; i++) {  // byte-serial tail / short distances: also correct for self-overlapping matches
          const int32_t p = src + (int32_t)i;
          obuf[dst + (int32_t)i] = p >= 0 ? obuf[p] : gout[p];
        }
      }
      uint32_t m = __ballot_sync(0xFFFFFFFFu, ready && q.ml > kLaneMax);
      while (m) {  // long match: whole warp; byte j reads window[j mod off], which existed before the match began
        const uint32_t sl = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        const int32_t dd = __shfl_sync(0xFFFFFFFFu, dst, sl);
        const uint32_t oo = __shfl_sync(0xFFFFFFFFu, q.off, sl), l = __shfl_sync(0xFFFFFFFFu, q.ml, sl);
        for (uint32_t j = lane; j < l; j += 32) {
          const int32_t p = dd - (int32_t)oo + (int32_t)(oo >= l ? j : j % oo);
          obuf[dd + (int32_t)j] = p >= 0 ? obuf[p] : gout[p];
        }
      }
      pending = pending && !ready;
      __syncwarp();
    }
    ZN_CNT(4, clock64() - t_b0);
    __threadfence_block();
    if (lane == 0) bdone[b] = 1;
    __syncwarp();
  }
  team_sync(t);
  if (rc == S_OK) team_copy(t, out + base, obuf, dec);
  team_sync(t);
  // the slots were used as the output buffer: restore what the next walk / team executor expects
  if (t.tid < (uint32_t)kGroup) { ps->predef[t.tid] = 0; ps->slot[t.tid].huf.valid = 0; ps->slot[t.tid].tile = nullptr; }
  for (int j = 0; j < kGroup; j++) zs::init_luts(t, &ps->slot[j]);
  team_sync(t);
  if (t.tid == 0) ps->slot[0].tile = ps->tile;
  team_sync(t);
  return rc;
}

// Team executor for a block of long sequences: records -> ring -> exec_batch (vector copies, TMA bulk stores).
ZN_D uint32_t exec_block_team(const Team& t, ParShared* ps, const BlockRec& r, const SeqOut* seqs, uint8_t* out,
                              const uint8_t* lit, uint32_t frame_start, zs::ExecState& es) {
  DecShared* sh = &ps->slot[0];
  for (uint32_t first = 0; first < r.nseq; first += kSeqBatch) {
    const uint32_t count = r.nseq - first < kSeqBatch ? r.nseq - first : kSeqBatch;
    if (t.tid == 0) ps->err = 0;
    team_sync(t);  // previous batch's readers of the ring are done
    if (t.tid < count) {
      const SeqOut q = seqs[first + t.tid];
      const uint32_t d = r.base_out + q.out_rel + q.ll;
      if (q.ml && (q.off == 0 || q.off > d - frame_start)) ps->err = 1;
      SeqRec rr;
      rr.lit = q.lit_rel; rr.ll = q.ll; rr.ml = q.ml; rr.off = q.off;
      sh->ring[t.tid] = rr;
    }
    team_sync(t);
    if (ps->err) return S_DECODE_ERROR;
    zs::exec_batch(t, sh, count, out, lit, r.lit_rle, es);
  }
  const uint32_t rest = r.lit_len - r.lit_used;
  if (rest) {
    if (r.lit_rle >= 0) team_fill(t, out + es.pos, (uint32_t)r.lit_rle, rest);
    else team_copy(t, out + es.pos, lit + r.lit_used, rest);
    es.pos += rest;
  }
  return S_OK;
}

// One blob, block-parallel.  Team-uniform.
ZN_D uint32_t decode_blob_par(const Team& t, ParShared* ps, const uint8_t* src, uint32_t src_len, uint8_t* out, uint32_t cap,
                              uint8_t* scratch, uint32_t* produced) {
  *produced = 0;
  if (src_len >= 4 && ld32le(src) == 0x184D2204u) {  // LZ4 frame: team path
    return lz::decode_frame(t, &ps->slot[0], src, src_len, out, cap, produced);
  }
  SeqOut* seqbuf = reinterpret_cast<SeqOut*>(scratch);
  uint8_t* litbuf = scratch + kParSeqBytes;
  const uint32_t warp = t.tid >> 5, lane = t.tid & 31u;
  uint32_t ip = 0;
  zs::ExecState es;
  es.pos = 0; es.wm = 0; es.bulk = 0;
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
    if (!single) {
      if (src_len < hp + 1) return S_DECODE_ERROR;
      const uint32_t wd = src[hp++];
      const uint64_t bs = 1ull << (10 + (wd >> 3));
      window = bs + (bs >> 3) * (wd & 7);
    }
    const uint32_t db = did_flag == 3 ? 4u : did_flag;
    if (src_len < hp + db) return S_DECODE_ERROR;
    uint32_t did = 0;
    for (uint32_t i = 0; i < db; i++) did |= (uint32_t)src[hp + i] << (8 * i);
    hp += db;
    if (did != 0) return S_UNSUPPORTED;
    const uint32_t fb = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2u : (fcs_flag == 2 ? 4u : 8u));
    if (src_len < hp + fb) return S_DECODE_ERROR;
    uint64_t fcs = 0;
    for (uint32_t i = 0; i < fb; i++) fcs |= (uint64_t)src[hp + i] << (8 * i);
    if (fb == 2) fcs += 256;
    hp += fb;
    if (single) window = fcs;
    ip = hp;
    const uint32_t frame_start = es.pos;
    const uint32_t block_max = window < kZstdBlockMax ? (uint32_t)window : kZstdBlockMax;
    Defs defs{kDefNone, {kDefNone, kDefNone, kDefNone}};  // thread 0's copy is the live one
    uint32_t rep[3] = {1, 4, 8};                          // likewise
    team_sync(t);
    if (t.tid < (uint32_t)kGroup) ps->slot[t.tid].huf.valid = 0;
    for (;;) {
      // ---- walk
      if (t.tid == 0) walk_group(src, src_len, ip, block_max, defs, ps->recs, &ps->gi);
      team_sync(t);
      ZN_TP(30);
      const uint32_t gn = ps->gi.n;
      // ---- phase A: warp pair j -> block j
      {
        const uint32_t j = warp >> 1;
        if (j < gn && ps->recs[j].type == 2) {
          if ((warp & 1u) == 0) {
            if (lane == 0)
              decode_block_sequences(src, src_len, &ps->recs[j], &ps->slot[j], seqbuf + (size_t)j * kMaxSeq, ps->predef[j]);
          } else {
            decode_block_literals(Team{lane, 32u}, src, src_len, &ps->recs[j], &ps->slot[j], litbuf + (size_t)j * kLitStride);
          }
        }
      }
      team_sync(t);
      ZN_TP(31);
      // ---- B0: chain the group
      if (t.tid == 0) {
        SeqOut* ptrs[kGroup];
        for (int j = 0; j < kGroup; j++) ptrs[j] = seqbuf + (size_t)j * kMaxSeq;
        uint32_t pos = es.pos;
        ps->status = chain_group(ps->recs, gn, ptrs, rep, &pos, cap, frame_start);
        ps->pos_after = pos;
      }
      team_sync(t);
      ZN_TP(32);
      if (ps->status != S_OK) return ps->status;
      // ---- phase B: execute the blocks in order
      for (uint32_t j = 0; j < gn; j++) {
        const BlockRec r = ps->recs[j];
        if (r.type == 0) {
          team_copy(t, out + es.pos, src + r.off, r.len);
          es.pos += r.len;
        } else if (r.type == 1) {
          team_fill(t, out + es.pos, src[r.off], r.len);
          es.pos += r.len;
        } else {
          const SeqOut* seqs = seqbuf + (size_t)j * kMaxSeq;
          const uint8_t* lit = r.lit_in_src != 0xFFFFFFFFu ? src + r.lit_in_src : litbuf + (size_t)j * kLitStride;
          const uint32_t dec = r.matched + (r.lit_len - r.lit_used);
          uint32_t rc;
          if (r.nseq == 0 || dec >= kTeamAvg * r.nseq) {
            rc = exec_block_team(t, ps, r, seqs, out, lit, frame_start, es);
          } else {
            zs::mem_sync(t, es);  // everything written so far is visible; no bulk store in flight
            rc = exec_block_lanes(t, ps, r, seqs, out, lit, frame_start);
            es.pos = r.base_out + dec;
            es.wm = es.pos;
          }
          if (rc != S_OK) return rc;
        }
        *produced = es.pos;
        ZN_TP(33);
      }
      if (ps->gi.status != S_OK) return ps->gi.status;
      ip = ps->gi.next_ip;
      const uint32_t last = ps->gi.last;
      team_sync(t);  // recs / gi are rewritten by the next walk
      if (last) break;
    }
    if (fb != 0 && (uint64_t)(es.pos - frame_start) != fcs) return S_SIZE_MISMATCH;
    if (checksum) {
      if (src_len - ip < 4) return S_DECODE_ERROR;
      ip += 4;
    }
  }
  *produced = es.pos;
  return S_OK;
}

__global__ void __launch_bounds__(kParThreads, 1) k_decode_par(const BlobDesc* __restrict__ blobs, const uint32_t* __restrict__ list,
                                                                uint32_t n_list, const uint8_t* blobs_base, uint8_t* out_base,
                                                                uint8_t* scratch, uint32_t* status, uint32_t* produced,
                                                                uint32_t* work_counter) {
  extern __shared__ __align__(128) uint8_t par_smem[];
  ParShared* ps = reinterpret_cast<ParShared*>(par_smem);
  const Team t{threadIdx.x, (uint32_t)kParThreads};
  uint8_t* my_scratch = scratch + (size_t)blockIdx.x * kParScratchPerCta;
  if (threadIdx.x < (uint32_t)kGroup) { ps->predef[threadIdx.x] = 0; ps->slot[threadIdx.x].tile = nullptr; }
  if (threadIdx.x == 0) ps->slot[0].tile = ps->tile;
  for (int j = 0; j < kGroup; j++) zs::init_luts(t, &ps->slot[j]);
  for (;;) {
    if (threadIdx.x == 0) ps->item = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t item = ps->item;
    __syncthreads();
    if (item >= n_list) break;
    const uint32_t blob = list[item];
    const BlobDesc d = blobs[blob];
    uint32_t st, got = 0;
    if (d.src_len >= 0xFFFFFFF0ull || d.dst_cap >= 0xFFFFFFF0ull) {
      st = S_UNSUPPORTED;
    } else {
      const uint8_t* src = blobs_base + d.src_off;
      if (d.src_len <= kSrcStage) {
        team_copy(t, ps->src + 16, src, (uint32_t)d.src_len);
        __syncthreads();
        src = ps->src + 16;
      }
      if (d.flags & F_LZ4_BLOCK)
        st = decode_lz4_block(t, &ps->slot[0], src, (uint32_t)d.src_len, out_base + d.dst_off, (uint32_t)d.dst_cap, &got);
      else
        st = decode_blob_par(t, ps, src, (uint32_t)d.src_len, out_base + d.dst_off, (uint32_t)d.dst_cap, my_scratch, &got);
      if (st == S_OK && got != (uint32_t)d.dst_cap) st = S_SIZE_MISMATCH;
    }
    if (threadIdx.x < kBulkIssuers) bulk_wait_all();
    __syncthreads();
    if (threadIdx.x == 0) {
      status[blob] = st;
      produced[blob] = got;
    }
  }
}

}  // namespace par
}  // namespace zn
