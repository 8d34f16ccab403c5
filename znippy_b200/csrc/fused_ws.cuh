// K4w: decode + BLAKE3 chunk hashing in ONE kernel with specialised warps (decompress.rs:157 + :172 for a batch of
// large, highly compressible blobs — the 2 GiB pattern file of BASELINE configs[1]).
//
// Decoding such blobs is bound by HBM writes (a few TMA bulk stores per 128 KiB block, issued by a handful of threads
// that then wait), hashing them by the integer ALU pipe: different resources, so the two run at the same time on the
// same SMs.  One CTA per SM = two 128-thread decode teams (named barriers 1 and 2) + 16 hash warps.
//   * A decode team takes blobs from the work counter and decodes them exactly like k_decode<128>.  After every zstd
//     block it turns the bytes that became final and visible (the executor's watermark) into 32-chunk TILES and appends
//     them to a device-wide queue: release fence, reserve slots with one atomicAdd on the tail, store the entries.
//   * Every hash warp of every SM takes the next queue index (one atomicAdd on the head), waits for that entry to be
//     filled, and hashes the tile (same b3_warp_tile as k_b3_chunks).  Hash work is therefore balanced over the whole
//     GPU however the blobs are spread over the SMs, and it starts a few microseconds after the first block is out.
//   * When the work list is empty the decode warps become hash warps too (their decode state in shared memory becomes
//     their staging buffers), so the tail hashes with all 24 warps of every SM.
// Consumers wait for producers only; producers never wait for anybody, and all CTAs of the launch are resident (grid <=
// SM count, one CTA per SM), so the waiting cannot deadlock and never crosses a kernel launch.
// The chunk chaining values land where k_b3_chunks would have put them (the blobs carry F_HASHED, so that kernel skips
// them); the tree kernels are unchanged.
#pragma once
#include "blake3_kernels.cuh"
#include "decode_kernels.cuh"

namespace zn {

constexpr int kWsTeamThreads = 128;
constexpr int kWsTeams = 2;
constexpr int kWsTeamWarps = kWsTeamThreads / 32;
constexpr int kWsDecWarps = kWsTeams * kWsTeamWarps;
constexpr int kWsHashWarps = 16;
constexpr int kWsThreads = 32 * (kWsDecWarps + kWsHashWarps);
constexpr uint32_t kWsTileBytes = 32u * kChunk;

struct WsDecRegion {
  alignas(128) uint8_t tile[kTileBytes];
  alignas(16) uint8_t src[kSrcStage + 32];
  DecShared sh;
  uint32_t base;
};

// per decode team: its decode state, later the staging buffers of its 4 warps
constexpr uint32_t kWsTeamBytes =
    ((sizeof(WsDecRegion) > (size_t)kWsTeamWarps * kB3SmemPerWarp ? sizeof(WsDecRegion) : (size_t)kWsTeamWarps * kB3SmemPerWarp) + 127u) & ~127u;
constexpr uint32_t kWsSmemBytes = kWsTeams * kWsTeamBytes + kWsHashWarps * kB3SmemPerWarp;
static_assert(kWsSmemBytes <= 227u * 1024u, "one CTA per SM");

// Device-wide tile queue of one plan run: ctl[0] = head (consumers), ctl[1] = tail (producers), ctl[2] = stall flag
// (watchdog), ent[i] = 0 until filled,
// then (list index << 32) | (tile + 1).  Zeroed before every launch.
struct WsQueue {
  uint32_t* ctl;
  unsigned long long* ent;
  uint32_t total;  // tiles of the whole launch
};

#ifdef ZN_WS_DEBUG
__device__ unsigned long long g_ws_dbg[4];  // first start, last decode end, last hash end (globaltimer ns), tiles hashed by then
ZN_D unsigned long long ws_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

ZN_D void ws_push(const WsQueue& q, uint32_t list_idx, uint32_t first_tile, uint32_t count) {  // one thread
  if (!count) return;
  __threadfence();  // the tile's bytes before its queue entry, device-wide
  const uint32_t pos = atomicAdd(q.ctl + 1, count);
  for (uint32_t i = 0; i < count; i++)
    *reinterpret_cast<volatile unsigned long long*>(q.ent + pos + i) = ((unsigned long long)list_idx << 32) | (first_tile + i + 1u);
}

// decode_frames hook: after every block, queue the tiles that became final
struct WsHook {
  WsQueue q;
  uint32_t list_idx, cap, pushed;
  ZN_D void after_block(const Team& t, zs::ExecState& es) {
    // The watermark advances whenever a match reaches back past it (every block of periodic data): the tiles queued here
    // lag the decoder by about a block, and its bulk stores keep draining while the next block is parsed.  Only data
    // that never synchronises on its own gets a barrier here, so that the hash warps are never far behind.
    if (es.pos - es.wm >= (1u << 20)) zs::mem_sync(t, es);
    const uint32_t w = es.wm < cap ? es.wm : cap;
    const uint32_t ready = w / kWsTileBytes;
    // The LAST thread of the team does the queue work (fence, atomic, stores: 2-3 us of global round trips) while
    // thread 0 is already walking the next block's headers; the rest of the team waits at that block's first barrier
    // either way.  The team's writes are ordered before this thread's fence by the barrier in mem_sync.
    // (Making the teams pause while more than 64 MiB of tiles are queued — to keep the fresh output in L2 for the hash —
    // was tried: DRAM reads per launch stayed at 2.1 GB, L2 hit rate 25 %, and the kernel got 4 % slower.)
    if (t.tid == t.n - 1u && ready > pushed) ws_push(q, list_idx, pushed, ready - pushed);
    if (ready > pushed) pushed = ready;
  }
  ZN_D void finish(const Team&, zs::ExecState&) {}  // the kernel queues the remaining tiles on every exit path
};

__global__ void __launch_bounds__(kWsThreads, 1)
    k_decode_ws(const BlobDesc* __restrict__ blobs, const uint32_t* __restrict__ list, uint32_t n_list, const uint8_t* blobs_base,
                uint8_t* out_base, uint8_t* lit_scratch, uint32_t* status, uint32_t* produced, uint32_t* work_counter, uint32_t* cvs,
                WsQueue q, uint32_t one) {
  extern __shared__ __align__(128) uint8_t ws_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
#ifdef ZN_WS_DEBUG
  if (threadIdx.x == 0) atomicMin(&g_ws_dbg[0], ws_now());
#endif

  if (warp < (uint32_t)kWsDecWarps) {
    // ------------------------------------------------------------------------------------------ decode teams
    const uint32_t team = warp / kWsTeamWarps;
    WsDecRegion* dr = reinterpret_cast<WsDecRegion*>(ws_smem + team * kWsTeamBytes);
    const Team t{threadIdx.x - team * kWsTeamThreads, (uint32_t)kWsTeamThreads, 1u + team};
    DecShared* sh = &dr->sh;
    if (t.tid == 0) sh->tile = dr->tile;
    zs::init_luts(t, sh);
    uint8_t* lit = lit_scratch + (size_t)(blockIdx.x * kWsTeams + team) * kLitStride;
    uint32_t predef = 0;
    for (;;) {
      if (t.tid == 0) dr->base = atomicAdd(work_counter, 1u);
      team_sync(t);
      const uint32_t base = dr->base;
      if (base >= n_list) break;
      const uint32_t blob = list[base];
      const BlobDesc d = blobs[blob];
      WsHook hook{q, base, (uint32_t)d.dst_cap, 0u};
      uint32_t st, got = 0;
      if (d.src_len >= 0xFFFFFFF0ull || d.dst_cap >= 0xFFFFFFF0ull) {
        st = S_UNSUPPORTED;
      } else {
        const uint8_t* src = blobs_base + d.src_off;
        if (d.src_len <= kSrcStage) {
          team_copy(t, dr->src + 16, src, (uint32_t)d.src_len);
          team_sync(t);
          src = dr->src + 16;
        }
        if (d.flags & F_LZ4_BLOCK) {
          st = decode_lz4_block(t, sh, src, (uint32_t)d.src_len, out_base + d.dst_off, (uint32_t)d.dst_cap, &got);
        } else {
          st = decode_blob(t, sh, src, (uint32_t)d.src_len, out_base + d.dst_off, (uint32_t)d.dst_cap, lit, predef, &got, hook);
        }
        if (st == S_OK && got != (uint32_t)d.dst_cap) st = S_SIZE_MISMATCH;
      }
      if (t.tid < kBulkIssuers) bulk_wait_all();  // drain: the tile buffer is reused and the bytes are about to be declared final
      team_sync(t);
      if (t.tid == 0) {
        status[blob] = st;
        produced[blob] = got;
      }
      if (t.tid == t.n - 1u) {
        // every tile of the blob is queued exactly once, on errors too: the digest of a failed row is never looked at,
        // and the consumers count on the launch's tile total
        const uint32_t n_tiles = (d.n_chunks + 31u) / 32u;
        ws_push(q, base, hook.pushed, n_tiles - hook.pushed);
      }
    }
#ifdef ZN_WS_DEBUG
    if (t.tid == 0) { atomicMax(&g_ws_dbg[1], ws_now()); atomicMax(&g_ws_dbg[3], (unsigned long long)*(volatile uint32_t*)q.ctl); }
#endif
    team_sync(t);  // nobody of the team touches its decode region any more: it becomes staging for these four warps
  }

  // ---------------------------------------------------------------------------------------------- hashing
  uint8_t* wbuf = warp < (uint32_t)kWsDecWarps
                      ? ws_smem + (warp / kWsTeamWarps) * kWsTeamBytes + (warp % kWsTeamWarps) * kB3SmemPerWarp
                      : ws_smem + kWsTeams * kWsTeamBytes + (warp - kWsDecWarps) * kB3SmemPerWarp;
  for (;;) {
    uint32_t idx = 0, e_lo = 0, e_hi = 0;
    if (lane == 0) {
      idx = atomicAdd(q.ctl, 1u);
      if (idx < q.total) {
        unsigned long long e;
        uint32_t spins = 0;
        while ((e = *reinterpret_cast<volatile unsigned long long*>(q.ent + idx)) == 0ull) {
          __nanosleep(100);
          // Watchdog: producers never wait, so an entry that stays empty for seconds means a bug, not load.  Give up
          // (ctl[2] != 0 makes zn_plan_results fail) instead of hanging the GPU.
          if (++spins > (1u << 25) || ((spins & 4095u) == 0 && *reinterpret_cast<volatile uint32_t*>(q.ctl + 2))) {
            atomicExch(q.ctl + 2, 1u);
            idx = 0xFFFFFFFFu;
            break;
          }
        }
        __threadfence();
        e_lo = (uint32_t)e;
        e_hi = (uint32_t)(e >> 32);
      }
    }
    idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
    if (idx >= q.total) break;
    const uint32_t tile = __shfl_sync(0xFFFFFFFFu, e_lo, 0) - 1u, li = __shfl_sync(0xFFFFFFFFu, e_hi, 0);
    const BlobDesc d = blobs[list[li]];
    const uint32_t cap = (uint32_t)d.dst_cap;
    const uint32_t g = tile * 32u + lane;
    const bool act = g < d.n_chunks;
    const uint32_t rem = act ? cap - min(cap, g * kChunk) : 0u;
    uint32_t cv[8];
#ifdef ZN_WS_NOHASH
    for (int k = 0; k < 8; k++) cv[k] = 0;
#else
    b3_warp_tile(wbuf, out_base + d.dst_off + (size_t)g * kChunk, rem < kChunk ? rem : kChunk, g, d.n_chunks == 1, act, lane, one, cv);
#endif
    if (act) b3::store_cv(cvs + (size_t)d.cv_base * 8 + (size_t)g * 8, cv);
  }
#ifdef ZN_WS_DEBUG
  if (lane == 0) atomicMax(&g_ws_dbg[2], ws_now());
#endif
}

}  // namespace zn
