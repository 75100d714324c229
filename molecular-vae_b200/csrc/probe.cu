// L2 <-> SM fabric probe: the roofline denominator of the persistent recurrence sweeps.  The sweeps keep the
// tensor pipe ~25 % busy because every step moves 70-95 MB between L2 and the SMs (streamed operand re-read by the unit
// slices, saved gates, dG); what bounds them is the chip-wide L2 -> SM delivery rate, which no public number states.  These
// two kernels measure it on the device the bench runs on: all SMs stream an L2-resident buffer (ld.global.cg, 16 B per
// thread per access, L1 bypassed) -- read-only, and read + write (a copy inside L2).  bench.py times them with CUDA events.
#include "../../include/mvae_b200.h"
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(512) l2_read_kernel(const uint4* __restrict__ buf, long long n16, int passes, unsigned int* sink) {
  uint4 acc = make_uint4(0u, 0u, 0u, 0u);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p) {
    // every pass walks the whole buffer; a CTA touches a different part each pass (rotation) so nothing is L1 / register resident
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x + (long long)p * 7919 * blockDim.x) % stride;
    for (; i < n16; i += stride) {
      const uint4 v = __ldcg(buf + i);
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9E3779B9u) *sink = acc.x;   // keeps the loads alive
}

__global__ void __launch_bounds__(512) l2_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n16, int passes) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; ++p) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x + (long long)p * 7919 * blockDim.x) % stride;
    for (; i < n16; i += stride) __stcg(dst + i, __ldcg(src + i));
  }
}

}  // namespace

extern "C" int mvae_l2_probe(void* buf, size_t bytes, int passes, int mode, int ctas, mvae_stream_t stream) {
  if (!buf || bytes < (1u << 20) || (reinterpret_cast<uintptr_t>(buf) & 15) || passes < 1 || ctas < 1) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long n16 = (long long)(bytes / 16);
  if (mode == 0) {
    l2_read_kernel<<<ctas, 512, 0, st>>>(reinterpret_cast<const uint4*>(buf), n16, passes, reinterpret_cast<unsigned int*>(buf));
  } else {
    const long long half = n16 / 2;
    l2_copy_kernel<<<ctas, 512, 0, st>>>(reinterpret_cast<const uint4*>(buf), reinterpret_cast<uint4*>(buf) + half, half, passes);
  }
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
