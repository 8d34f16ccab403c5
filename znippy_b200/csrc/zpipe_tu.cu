// Translation unit of the device-wide zstd decode pipeline: kernels (zpipe_kernels.cuh) + their launcher.
#include <algorithm>

#include "host_api.h"
#include "zpipe_kernels.cuh"

namespace zn {
namespace zp {

__global__ void k_zinit(ZPools* pools, uint32_t seq_cap, uint32_t lit_cap16, uint32_t tab_cap, uint32_t comp_cap) {
  ZPools z;
  z.seq_used = 0; z.seq_cap = seq_cap;
  z.lit_used16 = 0; z.lit_cap16 = lit_cap16;
  z.tab_used = 0; z.tab_cap = tab_cap;
  z.comp_used = 0; z.comp_cap = comp_cap;
  *pools = z;
}

bool pipeline_init() {
  static FseD zset[kTabSet];
  build_predef_set(zset);
  if (cudaMemcpyToSymbol(g_zpredef, zset, sizeof zset) != cudaSuccess) return false;
  cudaFuncSetAttribute(k_zlit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLitSmem);
  cudaFuncSetAttribute(k_zexec<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExecShared<512>));
  cudaFuncSetAttribute(k_zexec<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExecShared<128>));
  return true;
}

void pipeline_enqueue(const PipelineLaunch& L, cudaStream_t st, cudaEvent_t* marks) {
  const ZArgs& a = L.a;
  const uint32_t sms = L.sm_count, slots = L.slots;
  int m = 0;
  auto mark = [&]() { if (marks) cudaEventRecord(marks[m++], st); };
  mark();
  k_zinit<<<1, 1, 0, st>>>(a.pools, L.seq_cap, L.lit_cap16, L.tab_cap, slots);
  k_zwalk<<<(a.nzb + 63) / 64, 64, 0, st>>>(a);
  mark();
  k_ztables<<<std::min<uint32_t>((slots + kTabWarps - 1) / kTabWarps, sms * 8), kTabWarps * 32, 0, st>>>(a);
  mark();
  k_zseq<<<(slots + 63) / 64, 64, 0, st>>>(a);
  mark();
  k_zlit<<<std::min<uint32_t>((slots + kLitBlocks - 1) / kLitBlocks, sms * 3), kLitBlocks * 4, kLitSmem, st>>>(a);
  mark();
  k_zchain<<<(a.nzb + 63) / 64, 64, 0, st>>>(a);
  mark();
  if (L.mean_bytes >= (256u << 10))
    k_zexec<512><<<std::min<uint32_t>(a.nzb, sms * 2), 512, sizeof(ExecShared<512>), st>>>(a, L.d_out, L.produced, L.exec_counter);
  else
    k_zexec<128><<<std::min<uint32_t>(a.nzb, sms * 6), 128, sizeof(ExecShared<128>), st>>>(a, L.d_out, L.produced, L.exec_counter);
  mark();
}

}  // namespace zp
}  // namespace zn
