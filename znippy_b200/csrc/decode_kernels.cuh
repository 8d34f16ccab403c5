// K1/K2: batched blob decode.  A persistent grid of CTAs pulls blobs from a work counter (largest first, so the
// tail of the batch is made of small blobs); each CTA is the "team" of zstd_decode.cuh / lz4_decode.cuh.
// Replaces the codec call of the reference's read loop (znippy-common/src/decompress.rs:156-166) and of
// ZnippyArchive::extract_file (archive.rs:159-164) for a whole batch of index rows at once.
#pragma once
#include "lz4_decode.cuh"

namespace zn {

constexpr uint32_t kLitStride = kZstdBlockMax + 256;  // per-CTA Huffman literal scratch
constexpr uint32_t kSrcStage = 8192;                  // blobs up to this size are parsed out of shared memory

template <int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_decode(const BlobDesc* __restrict__ blobs, const uint32_t* __restrict__ list,
                                               uint32_t n_list, const uint8_t* blobs_base, uint8_t* out_base,
                                               uint8_t* lit_scratch, uint32_t* status, uint32_t* produced,
                                               uint32_t* work_counter) {
  __shared__ DecShared sh;
  __shared__ uint32_t s_item;
  // Small blobs (every blob of the pattern corpora and of the 10 KiB-file corpus is < 1 KiB) are staged in shared
  // memory once, so that the serial header / bit-stream parsing costs ~30 cycles per access instead of an HBM trip.
  __shared__ __align__(16) uint8_t s_src[kSrcStage + 32];
  const Team t{threadIdx.x, (uint32_t)NT};
  uint8_t* lit = lit_scratch + (size_t)blockIdx.x * kLitStride;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t item = s_item;
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
        team_copy(t, s_src + 16, src, (uint32_t)d.src_len);
        __syncthreads();
        src = s_src + 16;
      }
      st = decode_blob(t, &sh, src, (uint32_t)d.src_len, out_base + d.dst_off, (uint32_t)d.dst_cap, lit, &got);
      if (st == S_OK && got != (uint32_t)d.dst_cap) st = S_SIZE_MISMATCH;
    }
    if (threadIdx.x < kBulkIssuers) bulk_wait_all();  // bulk stores read sh.tile: drain before the next blob (or exit) reuses it
    __syncthreads();  // every path out of decode_blob is team-uniform; this also fences the blob's last stores
    if (threadIdx.x == 0) {
      status[blob] = st;
      produced[blob] = got;
    }
  }
}

// Store-as-is rows with an output buffer (decompress.rs:164-166 + the pwrite source): plain gather, one CTA per
// 64 KiB piece so that a single 200 MiB jar still spreads over the whole machine.
constexpr uint32_t kGatherPiece = 64u * 1024u;
__global__ void __launch_bounds__(256) k_gather_raw(const BlobDesc* __restrict__ blobs,
                                                    const uint32_t* __restrict__ piece_blob,
                                                    const uint32_t* __restrict__ piece_idx, uint32_t n_pieces,
                                                    const uint8_t* __restrict__ blobs_base, uint8_t* out_base) {
  for (uint32_t p = blockIdx.x; p < n_pieces; p += gridDim.x) {
    const BlobDesc d = blobs[piece_blob[p]];
    const uint64_t o = (uint64_t)piece_idx[p] * kGatherPiece;
    const uint64_t rem = d.dst_cap - o;
    const uint32_t n = rem < kGatherPiece ? (uint32_t)rem : kGatherPiece;
    const Team t{threadIdx.x, blockDim.x};
    team_copy(t, out_base + d.dst_off + o, blobs_base + d.src_off + o, n);
  }
}

}  // namespace zn
