// BLAKE3 batch kernels: chunk chaining values (flat over every chunk of every blob) and tree merge + compare.
#pragma once
#include "blake3.cuh"

namespace zn {

// Where a blob's content lives: decoded rows in the output buffer, store-as-is rows in the blobs buffer.
ZN_D const uint8_t* content_ptr(const BlobDesc& d, const uint8_t* blobs_base, const uint8_t* out_base) {
  return (d.flags & F_COMPRESSED) ? out_base + d.dst_off : blobs_base + d.src_off;
}

// K3a: one lane per 1 KiB chunk, flat over the whole batch (so 100k x 10 KiB files fill warps as well as
// 256 x 8 MiB slices do).  chunk_prefix[b] = first flat chunk index of blob b (n_blobs+1 entries).
__global__ void __launch_bounds__(256) k_b3_chunks(const BlobDesc* __restrict__ blobs,
                                                   const uint32_t* __restrict__ chunk_prefix, uint32_t n_blobs,
                                                   uint32_t total_chunks, const uint8_t* __restrict__ blobs_base,
                                                   const uint8_t* __restrict__ out_base,
                                                   uint32_t* __restrict__ cvs) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_chunks) return;
  // binary search: largest b with chunk_prefix[b] <= g
  uint32_t lo = 0, hi = n_blobs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(chunk_prefix + mid) <= g) lo = mid; else hi = mid;
  }
  const BlobDesc d = blobs[lo];
  const uint32_t c = g - __ldg(chunk_prefix + lo);
  const uint8_t* p = content_ptr(d, blobs_base, out_base) + (uint64_t)c * kChunk;
  const uint64_t remain = d.dst_cap - (uint64_t)c * kChunk;
  const uint32_t len = remain < kChunk ? (uint32_t)remain : kChunk;
  uint32_t cv[8];
  b3::hash_chunk(p, len, c, d.n_chunks == 1, cv);
  b3::store_cv(cvs + (uint64_t)g * 8, cv);
}

// digest = cv words little-endian; compare with expect and fold into status (decode errors win).
ZN_D void finish_blob(const BlobDesc& d, uint32_t blob, const uint32_t (&cv)[8], uint32_t* __restrict__ digests,
                      const uint32_t* __restrict__ expect, uint32_t* __restrict__ status) {
  b3::store_cv(digests + (uint64_t)blob * 8, cv);
  if ((d.flags & F_HAS_EXPECT) && status[blob] == S_OK) {
    bool eq = true;
#pragma unroll
    for (int i = 0; i < 8; i++) eq &= (__ldg(expect + (uint64_t)blob * 8 + i) == cv[i]);
    if (!eq) status[blob] = S_DIGEST_MISMATCH;
  }
}

// K3b: small blobs (<= kTreeSmallMax chunks): one lane walks the levels of its own blob in place.
__global__ void __launch_bounds__(128) k_b3_tree_small(const BlobDesc* __restrict__ blobs,
                                                       const uint32_t* __restrict__ list, uint32_t n_list,
                                                       uint32_t* __restrict__ cvs, uint32_t* __restrict__ digests,
                                                       const uint32_t* __restrict__ expect,
                                                       uint32_t* __restrict__ status) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_list) return;
  const uint32_t blob = list[t];
  const BlobDesc d = blobs[blob];
  uint32_t* base = cvs + d.cv_base * 8;
  uint32_t n = d.n_chunks;
  uint32_t l[8], r[8], o[8];
  while (n > 1) {
    const uint32_t pairs = n >> 1;
    for (uint32_t j = 0; j < pairs; j++) {
      b3::load_cv(base + (uint64_t)(2 * j) * 8, l);
      b3::load_cv(base + (uint64_t)(2 * j + 1) * 8, r);
      b3::parent(l, r, n == 2, o);
      b3::store_cv(base + (uint64_t)j * 8, o);
    }
    if (n & 1) {
      b3::load_cv(base + (uint64_t)(n - 1) * 8, l);
      b3::store_cv(base + (uint64_t)pairs * 8, l);
    }
    n = pairs + (n & 1);
  }
  b3::load_cv(base, o);
  finish_blob(d, blob, o, digests, expect, status);
}

// K3c: large blobs: one CTA per blob, every level spread over the CTA, in place (read -> barrier -> write).
__global__ void __launch_bounds__(256) k_b3_tree_large(const BlobDesc* __restrict__ blobs,
                                                       const uint32_t* __restrict__ list,
                                                       uint32_t* __restrict__ cvs, uint32_t* __restrict__ digests,
                                                       const uint32_t* __restrict__ expect,
                                                       uint32_t* __restrict__ status) {
  const uint32_t blob = list[blockIdx.x];
  const BlobDesc d = blobs[blob];
  uint32_t* base = cvs + d.cv_base * 8;
  uint32_t n = d.n_chunks;
  uint32_t l[8], r[8], o[8];
  while (n > 1) {
    const uint32_t pairs = n >> 1, items = pairs + (n & 1);  // the odd tail is a pass-through item
    for (uint32_t j0 = 0; j0 < items; j0 += blockDim.x) {
      const uint32_t j = j0 + threadIdx.x;
      const bool act = j < items;
      if (act) {
        b3::load_cv(base + (uint64_t)(2 * j) * 8, l);
        if (j < pairs) {
          b3::load_cv(base + (uint64_t)(2 * j + 1) * 8, r);
          b3::parent(l, r, n == 2, o);
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++) o[i] = l[i];
        }
      }
      __syncthreads();  // all reads of slots [2*j0, 2*j0+2*blockDim) done before slots [j0, j0+blockDim) are overwritten
      if (act) b3::store_cv(base + (uint64_t)j * 8, o);
    }
    __syncthreads();
    n = items;
  }
  if (threadIdx.x == 0) {
    b3::load_cv(base, o);
    finish_blob(d, blob, o, digests, expect, status);
  }
}

constexpr uint32_t kTreeSmallMax = 64;

}  // namespace zn
