// Translation unit of the compression kernels (compress_kernels.cuh) + their host launcher.
#include "host_api.h"
#include "compress_kernels.cuh"

namespace zn {

void compress_init_attrs() {
  cz::PredefCTables ct;
  cz::build_predef_ctables(&ct);
  cudaMemcpyToSymbol(cz::g_predef_c, &ct, sizeof ct);
  cudaFuncSetAttribute(k_zstd_blocks<cz::WinFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZstdLaunch<cz::WinFast>::kSmemBytes);
  cudaFuncSetAttribute(k_zstd_blocks<cz::WinMid>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZstdLaunch<cz::WinMid>::kSmemBytes);
  cudaFuncSetAttribute(k_zstd_blocks<cz::WinHigh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZstdLaunch<cz::WinHigh>::kSmemBytes);
}

size_t compress_bound(size_t n, int codec) {
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
int compress_run(CompressScratch* cs, cudaStream_t st, int sm_count, const uint8_t* d_src, const uint64_t* src_off,
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
