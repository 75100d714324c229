// Device helpers shared by the persistent recurrence kernels (gru_rec.cu, gru_rec2.cu).
#pragma once
#include "common.cuh"

namespace rec {

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, int* err_flag) {
  if (ptx::mbar_try_wait(bar, parity)) return true;
  if (*(volatile int*)err_flag) return false;
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0) {
      const unsigned long long now = gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicExch(err_flag, 2); return false; }
      if (*(volatile int*)err_flag) return false;
    }
  }
  return true;
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool wait_counter(const unsigned int* ctr, unsigned int target, int* err_flag) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (ld_acquire(ctr) < target) {
    __nanosleep(32);
    if ((++spins & 0xFF) == 0) {
      const unsigned long long now = gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicExch(err_flag, 3); return false; }
      if (*(volatile int*)err_flag) return false;
    }
  }
  return true;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st16(__nv_bfloat16* dst, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf2(v[0], v[1]); a.y = pack_bf2(v[2], v[3]); a.z = pack_bf2(v[4], v[5]); a.w = pack_bf2(v[6], v[7]);
  b.x = pack_bf2(v[8], v[9]); b.y = pack_bf2(v[10], v[11]); b.z = pack_bf2(v[12], v[13]); b.w = pack_bf2(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}
__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&v)[16]) {
  v[0] = bf_lo(a.x); v[1] = bf_hi(a.x); v[2] = bf_lo(a.y); v[3] = bf_hi(a.y);
  v[4] = bf_lo(a.z); v[5] = bf_hi(a.z); v[6] = bf_lo(a.w); v[7] = bf_hi(a.w);
  v[8] = bf_lo(b.x); v[9] = bf_hi(b.x); v[10] = bf_lo(b.y); v[11] = bf_hi(b.y);
  v[12] = bf_lo(b.z); v[13] = bf_hi(b.z); v[14] = bf_lo(b.w); v[15] = bf_hi(b.w);
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// FAST: one MUFU per gate (tanh.approx, |err| ~ 5e-4); else exp/rcp based (|err| ~ 1e-6)
template <bool FAST> __device__ __forceinline__ float gate_sigmoid_t(float x) {
  if (FAST) return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f);
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
template <bool FAST> __device__ __forceinline__ float gate_tanh_t(float x) {
  if (FAST) return tanh_approx(x);
  return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}
__device__ __forceinline__ float gate_sigmoid(float x) { return gate_sigmoid_t<false>(x); }
__device__ __forceinline__ float gate_tanh(float x) { return gate_tanh_t<false>(x); }


struct __align__(32) u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg256(const void* p) {
  u32x8 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(void* p, const float (&v)[16]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(pack_bf2(v[0], v[1])),
               "r"(pack_bf2(v[2], v[3])), "r"(pack_bf2(v[4], v[5])), "r"(pack_bf2(v[6], v[7])), "r"(pack_bf2(v[8], v[9])),
               "r"(pack_bf2(v[10], v[11])), "r"(pack_bf2(v[12], v[13])), "r"(pack_bf2(v[14], v[15]))
               : "memory");
}
// 32 bytes per thread as two 16-byte halves 512 B apart: with lane l at +16*l each warp instruction covers 512
// contiguous bytes in full 32-byte sectors (a 256-bit access costs two partial sector transactions per sector in L2).
__device__ __forceinline__ u32x8 ldg2x128(const __nv_bfloat16* p) {
  u32x8 r;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p + 256));
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void stg2x128(__nv_bfloat16* p, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf2(v[0], v[1]); a.y = pack_bf2(v[2], v[3]); a.z = pack_bf2(v[4], v[5]); a.w = pack_bf2(v[6], v[7]);
  b.x = pack_bf2(v[8], v[9]); b.y = pack_bf2(v[10], v[11]); b.z = pack_bf2(v[12], v[13]); b.w = pack_bf2(v[14], v[15]);
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 256) = b;
}
__device__ __forceinline__ void unpack16(const u32x8& a, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[2 * i] = bf_lo(a.v[i]); v[2 * i + 1] = bf_hi(a.v[i]); }
}

}  // namespace rec
