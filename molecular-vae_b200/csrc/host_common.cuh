// Host-side helpers shared by the orchestration units (moses.cu, cfga.cu): workspace carving, the GEMM dispatcher
// (tcgen05 path in bf16 mode, CUDA-core SGEMM in fp32 check mode), and the token-table kernels.
#pragma once
#include "common.cuh"
#include "simt_kernels.cuh"
#include "umma_gemm.h"
#include <stdlib.h>

void mvae_count_launches(int n);   // cfgb.cu
struct mvae_graph;
int mvae_capture_into_graph(int (*fn)(void*, cudaStream_t), void* ctx, mvae_graph** out_graph);   // cfgb.cu

#define RC(expr) do { int _rc = (expr); if (_rc != MVAE_OK) return _rc; } while (0)
#define KCHECK() do { mvae_count_launches(1); MVAE_CUDA_CHECK(cudaGetLastError()); } while (0)

namespace {

struct Carver {
  uint8_t* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline int memset_async(void* p, size_t bytes, cudaStream_t st) {
  mvae_count_launches(1);
  MVAE_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  return MVAE_OK;
}
template <typename TA>
int gemm(int* err_flag, cudaStream_t st, const TA* A, long long lda, bool a_trans, const TA* B, long long ldb,
         bool b_kmajor, void* out, long long ldc, bool out_is_ta, int M, int N, int K, const float* bias,
         bool accumulate, int splits, int bn = 0, const mvae_umma_varlen* vl = nullptr, bool out_rb = false) {
  mvae_count_launches(1);
  if constexpr (sizeof(TA) == 4) {
    (void)out_is_ta; (void)bn; (void)err_flag; (void)vl; (void)out_rb;
    return simt::sgemm(st, reinterpret_cast<const float*>(A), a_trans ? 1 : lda, a_trans ? lda : 1,
                       reinterpret_cast<const float*>(B), b_kmajor ? 1 : ldb, b_kmajor ? ldb : 1,
                       reinterpret_cast<float*>(out), ldc, M, N, K, bias, simt::ACT_NONE, accumulate ? 1 : 0, splits);
  } else {
    mvae_umma_operand a{A, a_trans ? 1 : 0, M, K, lda, 1, 0, 0, 0};
    mvae_umma_operand b{B, b_kmajor ? 0 : 1, N, K, ldb, 1, 0, 0, 0};
    mvae_umma_out o{out, ldc, out_is_ta ? 1 : 0, accumulate ? 1 : 0, bias, out_rb ? 1 : 0};
    return mvae_umma_gemm(&a, &b, &o, M, N, K, bn, splits, 0, err_flag, st, nullptr, nullptr, nullptr, vl);
  }
}
// bf16 mode: the small fp32 GEMMs (latent / encoder Linears, token tables) go through the tensor cores (bf16x3 split,
// fp32-class accuracy, umma_gemm.h); the context is set per call by the unit's check_ws.  MVAE_TC_SGEMM=0: CUDA cores only.
thread_local mvae_tc_ctx g_tc{nullptr, 0, nullptr};
inline bool tc_sgemm_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MVAE_TC_SGEMM"); v = e ? (atoi(e) != 0) : 1; }
  return v != 0;
}
inline int sg(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk,
              long long sbn, float* C, long long ldc, int M, int N, int K, const float* bias, int act, int accumulate,
              int splits = 1) {
  if (g_tc.scratch) {
    int n = 0;
    const int rc = mvae_tc_sgemm(&g_tc, st, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act, accumulate, &n);
    if (rc != MVAE_ERR_UNSUPPORTED) { mvae_count_launches(n); return rc; }
  }
  mvae_count_launches(1);
  return simt::sgemm(st, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act, accumulate, splits);
}
inline int sg_wgrad(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk,
                    long long sbn, float* C, long long ldc, int M, int N, int K) {
  if (ldc == N) RC(memset_async(C, (size_t)M * N * 4, st));
  else { mvae_count_launches(1); MVAE_CUDA_CHECK(cudaMemset2DAsync(C, ldc * 4, 0, (size_t)N * 4, M, st)); }
  const int splits = K >= 1024 ? 16 : (K >= 256 ? 4 : 1);
  return sg(st, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, nullptr, simt::ACT_NONE, 1, splits);
}

// out[t][b][c] = tbl[ids[b][t]][c] (+ add[b][c]);  rows b >= B are zero.  tbl is [V][W] fp32, out [T][Bp][W] TA.
// reverse: slab s holds time T-1-s (the processing order of a reverse-direction RNN).
template <typename TA>
__global__ void gather_rows_kernel(const float* __restrict__ tbl, int W, const uint8_t* __restrict__ ids, int ids_ld,
                                   const float* __restrict__ add, int B, int Bp, int T, TA* __restrict__ out,
                                   int reverse = 0) {
  const long long total = (long long)T * Bp * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W);
    const long long rb = i / W;
    const int b = (int)(rb % Bp);
    int t = (int)(rb / Bp);
    if (reverse) t = T - 1 - t;
    float v = 0.f;
    if (b < B) {
      v = tbl[(long long)ids[(long long)b * ids_ld + t] * W + c];
      if (add) v += add[(long long)b * W + c];
    }
    out[i] = from_f32<TA>(v);
  }
}
// OH[t*Bp + b][v] = 1 iff b < B and ids[b][t] == v   (CP columns)
// reverse 1: the whole time axis is reversed (t -> T-1-t, padding first); reverse 2 (needs lens): every row is reversed
// inside its own length (t -> lens[b]-1-t for t < lens[b], zero rows past it), so the padding stays at the end
template <typename TA>
__global__ void onehot_rows_kernel(const uint8_t* __restrict__ ids, int ids_ld, int B, int Bp, int T, int CP,
                                   TA* __restrict__ out, int reverse = 0, const int* __restrict__ lens = nullptr) {
  const long long total = (long long)T * Bp * CP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CP);
    const long long rb = i / CP;
    const int b = (int)(rb % Bp);
    int t = (int)(rb / Bp);
    bool live = b < B;
    if (reverse == 1) t = T - 1 - t;
    else if (reverse == 2 && live) { const int L = lens[b]; live = t < L; t = L - 1 - t; }
    out[i] = from_f32<TA>((live && ids[(long long)b * ids_ld + t] == c) ? 1.f : 0.f);
  }
}

}  // namespace
