// Host-side interfaces between the translation units of libznippy_cuda.so (znippy_cuda.cu: C ABI, decode and hash
// kernels; compress_tu.cu: compression kernels; zpipe_tu.cu: the device-wide zstd decode pipeline).  Splitting the
// library keeps a change to one kernel family from recompiling the other two.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>

#include "zpipe.cuh"

namespace zn {

// ---------------------------------------------------------------------------------------------- compress_tu.cu
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
void compress_init_attrs();
size_t compress_bound(size_t n, int codec);  // zn_compress_bound: raw-block fallback makes this exact
// Compresses n slices resident on the device into frames at d_dst + dst_off[i]; synchronises `st` before returning.
int compress_run(CompressScratch* cs, cudaStream_t st, int sm_count, const uint8_t* d_src, const uint64_t* src_off,
                 const uint64_t* src_len, uint32_t n, int level, int codec, uint8_t* d_dst, const uint64_t* dst_off,
                 const uint64_t* dst_cap, uint64_t* out_len, uint32_t* status, uint32_t* launches, std::string* err);

// ---------------------------------------------------------------------------------------------- zpipe_tu.cu
namespace zp {
bool pipeline_init();  // predefined tables + kernel attributes, once per context
struct PipelineLaunch {
  ZArgs a;
  uint32_t slots, seq_cap, lit_cap16, tab_cap;  // pool capacities of this batch
  uint32_t sm_count;
  uint64_t mean_bytes;  // mean decoded size of the batch's blobs: picks the exec team size
  uint8_t* d_out;
  uint32_t* produced;
  uint32_t* exec_counter;
};
// Enqueues init, walk, tables, seq (one kernel, or phase 1 + phase 2 for batches of large blobs), lit, chain, exec on
// `st` and returns the number of launches (7 or 8).  `marks` (nullable): 9 events; [0..6] are recorded before the first
// kernel and after walk / tables / seq / lit / chain / exec, [8] after seq phase 1 when ZN_ZPROF_SEQ1 is set.
uint32_t pipeline_enqueue(const PipelineLaunch& L, cudaStream_t st, cudaEvent_t* marks);
void pipeline_trace_dump();  // development builds (ZN_TRACE_BUILD=1): prints and clears the exec kernel's phase counters
}  // namespace zp

}  // namespace zn
