// K6/K7: batched per-slice compression kernels (one warp per block) and frame assembly; host launcher.
// See compress.cuh for the algorithms and the reference call sites this replaces.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "compress.cuh"
#include "coop.cuh"

namespace zn {

struct SliceDesc {
  uint64_t src_off, src_len;
  uint64_t dst_off;     // frame goes to dst_base + dst_off
  uint32_t blk_first;   // first global block index of this slice
  uint32_t n_blocks;
};

constexpr int kLz4WarpsPerCta = 4;
constexpr int kZstdWarpsPerCta = 1;
template <class WN>
struct ZstdLaunch {
  static constexpr uint32_t kWindow = (WN::kSmem + 127u) & ~127u;
  static constexpr uint32_t kSmemBytes = kWindow + (2u << WN::kHashLog);
  static constexpr uint32_t kCtasPerSm = (228u * 1024u) / (kSmemBytes + 1024u);
};

// which slice owns global block j (slices' blk_first ascending)
ZN_D uint32_t slice_of_block(const SliceDesc* __restrict__ s, uint32_t n_slices, uint32_t j) {
  uint32_t lo = 0, hi = n_slices;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (s[mid].blk_first <= j) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- LZ4: every 64 KiB block of every slice -> tmp slot; csize[j] = compressed size (>= n means "store raw")
__global__ void __launch_bounds__(kLz4WarpsPerCta * 32) k_lz4_blocks(const SliceDesc* __restrict__ slices, uint32_t n_slices,
                                                                     uint32_t total_blocks, const uint8_t* __restrict__ src_base,
                                                                     uint8_t* tmp, uint32_t* csize) {
  __shared__ uint16_t tabs[kLz4WarpsPerCta][1u << cz::kLz4HashLog];
  const uint32_t warp = threadIdx.x >> 5;
  const cz::Warp w{threadIdx.x & 31u, 32u};
  for (uint32_t j = blockIdx.x * kLz4WarpsPerCta + warp; j < total_blocks; j += gridDim.x * kLz4WarpsPerCta) {
    const uint32_t si = slice_of_block(slices, n_slices, j);
    const SliceDesc s = slices[si];
    const uint64_t o = (uint64_t)(j - s.blk_first) * cz::kLz4Block;
    const uint32_t n = (uint32_t)(s.src_len - o < cz::kLz4Block ? s.src_len - o : cz::kLz4Block);
    const uint32_t c = cz::lz4_compress_block(w, src_base + s.src_off + o, n, tmp + (size_t)j * cz::kLz4Slot, tabs[warp]);
    if (w.lane == 0) csize[j] = c;
    __syncwarp();
  }
}

// ---- LZ4 frame assembly: one CTA per slice
__global__ void __launch_bounds__(256) k_lz4_assemble(const SliceDesc* __restrict__ slices, const uint8_t* __restrict__ src_base,
                                                      const uint8_t* tmp, const uint32_t* __restrict__ csize, uint8_t* dst_base,
                                                      uint64_t* dst_len) {
  const SliceDesc s = slices[blockIdx.x];
  uint8_t* dst = dst_base + s.dst_off;
  const Team t{threadIdx.x, blockDim.x};
  if (threadIdx.x == 0) cz::lz4_frame_header(dst, s.src_len);
  uint64_t op = 15;
  for (uint32_t b = 0; b < s.n_blocks; b++) {
    const uint32_t j = s.blk_first + b;
    const uint64_t o = (uint64_t)b * cz::kLz4Block;
    const uint32_t n = (uint32_t)(s.src_len - o < cz::kLz4Block ? s.src_len - o : cz::kLz4Block);
    const uint32_t c = csize[j];
    const bool raw = c >= n;
    const uint32_t sz = raw ? n : c;
    if (threadIdx.x == 0) {
      const uint32_t word = sz | (raw ? 0x80000000u : 0u);
      dst[op] = (uint8_t)word; dst[op + 1] = (uint8_t)(word >> 8); dst[op + 2] = (uint8_t)(word >> 16); dst[op + 3] = (uint8_t)(word >> 24);
    }
    team_copy(t, dst + op + 4, raw ? src_base + s.src_off + o : tmp + (size_t)j * cz::kLz4Slot, sz);
    op += 4 + sz;
  }
  if (threadIdx.x == 0) {
    dst[op] = dst[op + 1] = dst[op + 2] = dst[op + 3] = 0;  // EndMark
    dst_len[blockIdx.x] = op + 4;
  }
}

// ---- zstd: every 128 KiB block of every slice -> staged payload; meta[2j] = payload offset in the slot, meta[2j+1] = size (0 = raw)
template <class WN>
__global__ void __launch_bounds__(32) k_zstd_blocks(const SliceDesc* __restrict__ slices, uint32_t n_slices, uint32_t total_blocks,
                                                    const uint8_t* __restrict__ src_base, uint8_t* tmp, uint64_t* seq_scratch,
                                                    uint32_t* meta) {
  extern __shared__ __align__(16) uint8_t zsm[];  // one warp per CTA: the input window, then the hash table
  uint8_t* D = zsm;
  uint16_t* tab = reinterpret_cast<uint16_t*>(zsm + ZstdLaunch<WN>::kWindow);
  const cz::Warp w{threadIdx.x, 32u};
  uint64_t* seqs = seq_scratch + (size_t)blockIdx.x * cz::kZstdMaxSeq;
  for (uint32_t j = blockIdx.x; j < total_blocks; j += gridDim.x) {
    const uint32_t si = slice_of_block(slices, n_slices, j);
    const SliceDesc s = slices[si];
    const uint64_t o = (uint64_t)(j - s.blk_first) * cz::kZstdCBlock;
    const uint32_t n = (uint32_t)(s.src_len - o < cz::kZstdCBlock ? s.src_len - o : cz::kZstdCBlock);
    uint32_t poff = 0;
    const uint32_t c = cz::zstd_compress_block<WN>(w, src_base + s.src_off, (uint32_t)o, n, tmp + (size_t)j * cz::kZstdSlot, seqs, D, tab,
                                               &poff);
    if (w.lane == 0) { meta[2 * j] = poff; meta[2 * j + 1] = c; }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) k_zstd_assemble(const SliceDesc* __restrict__ slices, const uint8_t* __restrict__ src_base,
                                                       const uint8_t* tmp, const uint32_t* __restrict__ meta, uint8_t* dst_base,
                                                       uint64_t* dst_len) {
  const SliceDesc s = slices[blockIdx.x];
  uint8_t* dst = dst_base + s.dst_off;
  const Team t{threadIdx.x, blockDim.x};
  if (s.src_len == 0) {  // same bytes as libzstd: single-segment, 1-byte content size 0, empty raw last block
    if (threadIdx.x == 0) {
      const uint8_t e[9] = {0x28, 0xB5, 0x2F, 0xFD, 0x20, 0x00, 0x01, 0x00, 0x00};
      for (int i = 0; i < 9; i++) dst[i] = e[i];
      dst_len[blockIdx.x] = 9;
    }
    return;
  }
  if (threadIdx.x == 0) cz::zstd_frame_header(dst, s.src_len);
  uint64_t op = 13;
  for (uint32_t b = 0; b < s.n_blocks; b++) {
    const uint32_t j = s.blk_first + b;
    const uint64_t o = (uint64_t)b * cz::kZstdCBlock;
    const uint32_t n = (uint32_t)(s.src_len - o < cz::kZstdCBlock ? s.src_len - o : cz::kZstdCBlock);
    const uint32_t poff = meta[2 * j], c = meta[2 * j + 1];
    const uint32_t last = b + 1 == s.n_blocks ? 1u : 0u;
    if (c == 0) {
      if (threadIdx.x == 0) cz::zstd_block_header(dst + op, last, 0, n);
      team_copy(t, dst + op + 3, src_base + s.src_off + o, n);
      op += 3 + n;
    } else {
      if (threadIdx.x == 0) cz::zstd_block_header(dst + op, last, 2, c);
      team_copy(t, dst + op + 3, tmp + (size_t)j * cz::kZstdSlot + poff, c);
      op += 3 + c;
    }
  }
  if (threadIdx.x == 0) dst_len[blockIdx.x] = op;
}

// ---------------------------------------------------------------------------------------------- host side
struct CompressScratch {
  uint8_t* tmp = nullptr;
  size_t tmp_cap = 0;
  uint64_t* seqs = nullptr;
  size_t seqs_cap = 0;
  void* small = nullptr;  // slices + meta + dst_len
  size_t small_cap = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float last_ms = 0.f;  // device time of the kernels of the last compress_run
  void release() {
    if (tmp) cudaFree(tmp);
    if (seqs) cudaFree(seqs);
    if (small) cudaFree(small);
    if (ev[0]) cudaEventDestroy(ev[0]);
    if (ev[1]) cudaEventDestroy(ev[1]);
    ev[0] = ev[1] = nullptr;
    tmp = nullptr; seqs = nullptr; small = nullptr;
    tmp_cap = seqs_cap = small_cap = 0;
  }
};

inline void compress_init_attrs() {
  cz::PredefCTables ct;
  cz::build_predef_ctables(&ct);
  cudaMemcpyToSymbol(cz::g_predef_c, &ct, sizeof ct);
  cudaFuncSetAttribute(k_zstd_blocks<cz::WinFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZstdLaunch<cz::WinFast>::kSmemBytes);
  cudaFuncSetAttribute(k_zstd_blocks<cz::WinMid>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZstdLaunch<cz::WinMid>::kSmemBytes);
  cudaFuncSetAttribute(k_zstd_blocks<cz::WinHigh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZstdLaunch<cz::WinHigh>::kSmemBytes);
}

// zn_compress_bound: raw-block fallback makes this exact
inline size_t compress_bound(size_t n, int codec) {
  if (codec == 2) return n + 4 * ((n + cz::kLz4Block - 1) / cz::kLz4Block) + 32;
  return n + 3 * ((n + cz::kZstdCBlock - 1) / cz::kZstdCBlock) + 32;
}

template <typename T>
static bool grow(T** p, size_t* cap, size_t need_bytes) {
  if (*cap >= need_bytes) return true;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  if (cudaMalloc((void**)p, need_bytes) != cudaSuccess) { cudaGetLastError(); return false; }
  *cap = need_bytes;
  return true;
}

// Compresses n slices resident on the device into frames at d_dst + dst_off[i]; synchronises `st` before returning.
inline int compress_run(CompressScratch* cs, cudaStream_t st, int sm_count, const uint8_t* d_src, const uint64_t* src_off,
                        const uint64_t* src_len, uint32_t n, int level, int codec, uint8_t* d_dst, const uint64_t* dst_off,
                        const uint64_t* /*dst_cap*/, uint64_t* out_len, uint32_t* status, uint32_t* launches, std::string* err) {
  // zstd: the level picks the window geometry of the match finder (three efforts); LZ4 has one effort
  const int effort = level <= 2 ? 0 : (level <= 9 ? 1 : 2);
  const bool lz4 = codec == 2;
  const uint64_t bsz = lz4 ? cz::kLz4Block : cz::kZstdCBlock;
  std::vector<SliceDesc> sl(n);
  uint64_t total_blocks = 0;
  for (uint32_t i = 0; i < n; i++) {
    sl[i].src_off = src_off[i];
    sl[i].src_len = src_len[i];
    sl[i].dst_off = dst_off[i];
    sl[i].blk_first = (uint32_t)total_blocks;
    sl[i].n_blocks = (uint32_t)((src_len[i] + bsz - 1) / bsz);
    total_blocks += sl[i].n_blocks;
    status[i] = src_len[i] >= 0xFFFFFFF0ull ? 4u /*UNSUPPORTED*/ : 0u;
    if (status[i]) { *err = "slice too large"; return -1; }
  }
  if (total_blocks > 0x7FFFFFFFull) { *err = "too many blocks"; return -1; }
  const uint32_t nb = (uint32_t)total_blocks;
  const size_t slot = lz4 ? cz::kLz4Slot : cz::kZstdSlot;
  const uint32_t wpc = lz4 ? kLz4WarpsPerCta : kZstdWarpsPerCta;
  const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>((nb + wpc - 1) / wpc, (uint32_t)sm_count * (lz4 ? 6u : (effort == 0 ? ZstdLaunch<cz::WinFast>::kCtasPerSm
                                                                  : effort == 1 ? ZstdLaunch<cz::WinMid>::kCtasPerSm
                                                                                : ZstdLaunch<cz::WinHigh>::kCtasPerSm))));
  const size_t small_bytes = (size_t)n * sizeof(SliceDesc) + (size_t)nb * 8 + (size_t)n * 8 + 64;
  if (!grow(&cs->tmp, &cs->tmp_cap, std::max<size_t>(1, (size_t)nb * slot)) ||
      !grow((uint8_t**)&cs->small, &cs->small_cap, small_bytes) ||
      (!lz4 && !grow(&cs->seqs, &cs->seqs_cap, (size_t)grid * wpc * cz::kZstdMaxSeq * 8))) {
    *err = "compress scratch allocation failed";
    return -3;
  }
  SliceDesc* d_sl = (SliceDesc*)cs->small;
  uint32_t* d_meta = (uint32_t*)((uint8_t*)cs->small + (size_t)n * sizeof(SliceDesc));
  uint64_t* d_len = (uint64_t*)((uint8_t*)d_meta + (size_t)nb * 8);
  d_len = (uint64_t*)(((uintptr_t)d_len + 7) & ~(uintptr_t)7);
  if (cudaMemcpyAsync(d_sl, sl.data(), (size_t)n * sizeof(SliceDesc), cudaMemcpyHostToDevice, st) != cudaSuccess) {
    *err = "compress H2D failed";
    return -2;
  }
  if (!cs->ev[0]) { cudaEventCreate(&cs->ev[0]); cudaEventCreate(&cs->ev[1]); }
  *launches = 0;
  cudaEventRecord(cs->ev[0], st);
  if (nb) {
    if (lz4) k_lz4_blocks<<<grid, kLz4WarpsPerCta * 32, 0, st>>>(d_sl, n, nb, d_src, cs->tmp, d_meta);
    else if (effort == 0)
      k_zstd_blocks<cz::WinFast><<<grid, 32, ZstdLaunch<cz::WinFast>::kSmemBytes, st>>>(d_sl, n, nb, d_src, cs->tmp, cs->seqs, d_meta);
    else if (effort == 1)
      k_zstd_blocks<cz::WinMid><<<grid, 32, ZstdLaunch<cz::WinMid>::kSmemBytes, st>>>(d_sl, n, nb, d_src, cs->tmp, cs->seqs, d_meta);
    else
      k_zstd_blocks<cz::WinHigh><<<grid, 32, ZstdLaunch<cz::WinHigh>::kSmemBytes, st>>>(d_sl, n, nb, d_src, cs->tmp, cs->seqs, d_meta);
    (*launches)++;
  }
  if (lz4) k_lz4_assemble<<<n, 256, 0, st>>>(d_sl, d_src, cs->tmp, d_meta, d_dst, d_len);
  else k_zstd_assemble<<<n, 256, 0, st>>>(d_sl, d_src, cs->tmp, d_meta, d_dst, d_len);
  (*launches)++;
  cudaEventRecord(cs->ev[1], st);
  if (cudaMemcpyAsync(out_len, d_len, (size_t)n * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    *err = std::string("compress kernels: ") + cudaGetErrorString(cudaGetLastError());
    return -2;
  }
  cudaEventElapsedTime(&cs->last_ms, cs->ev[0], cs->ev[1]);
  return 0;
}

}  // namespace zn
