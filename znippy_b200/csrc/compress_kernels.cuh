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

}  // namespace zn
