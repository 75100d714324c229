// Standalone checker for the tcgen05 GEMM (run on the GPU box): every operand-major combination,
// tile width, split-K, bias, accumulate, bf16/fp32 output and ragged edges against a CPU reference.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../molecular-vae_b200/csrc/umma_gemm.h"

extern "C" const char* mvae_last_cuda_error(void);

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

struct Case { int M, N, K, a_mn, b_mn, bn, splits, out_bf16, accumulate, bias, slabs, rb_a, rb_b, rb_out; };
// element (r, c) of a stored [rows][ld] matrix: plain row-major or the row-blocked layout [r/32][ld/16][32][16]
static inline long long IDX(int rb, long long r, long long c, long long ld) {
  return rb ? ((r >> 5) * (ld >> 3) + (c >> 3)) * 256 + (r & 31) * 8 + (c & 7) : r * ld + c;
}

static int run_case(const Case& c, int verbose) {
  const int slabs = c.slabs > 0 ? c.slabs : 1;
  const int slab = slabs - 1;
  // stored shapes
  const long long a_rows = c.a_mn ? c.K : c.M, a_cols = c.a_mn ? c.M : c.K;
  const long long b_rows = c.b_mn ? c.K : c.N, b_cols = c.b_mn ? c.N : c.K;
  const long long lda = c.rb_a ? (a_cols + 15) / 16 * 16 + 16 : (a_cols + 7) / 8 * 8 + 8;
  const long long ldb = c.rb_b ? (b_cols + 15) / 16 * 16 + 32 : (b_cols + 7) / 8 * 8 + 16;  // padded leading dims
  const long long a_slab = (c.rb_a ? (a_rows + 31) / 32 * 32 : a_rows) * lda, b_slab = (c.rb_b ? (b_rows + 31) / 32 * 32 : b_rows) * ldb;
  std::vector<__nv_bfloat16> hA(a_slab * slabs), hB(b_slab * slabs);
  std::vector<float> fA(a_slab * slabs), fB(b_slab * slabs);
  for (size_t i = 0; i < hA.size(); ++i) { float v = bf16_round(frand()); fA[i] = v; hA[i] = __float2bfloat16_rn(v); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = bf16_round(frand()); fB[i] = v; hB[i] = __float2bfloat16_rn(v); }
  const long long ldc = c.rb_out ? (c.N + 15) / 16 * 16 : (c.N + 7) / 8 * 8;
  const long long Mp = c.rb_out ? (c.M + 31) / 32 * 32 : c.M;
  std::vector<float> hC0((size_t)Mp * ldc), hBias(c.N);
  for (auto& v : hC0) v = (c.accumulate || c.splits > 1) ? frand() : 777.0f;
  for (auto& v : hBias) v = frand();
  __nv_bfloat16 *dA, *dB; void* dC; float* dBias; int* dErr;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dC, (size_t)Mp * ldc * 4); cudaMalloc(&dBias, c.N * 4); cudaMalloc(&dErr, 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dBias, hBias.data(), c.N * 4, cudaMemcpyHostToDevice);
  cudaMemset(dErr, 0, 4);
  if (c.out_bf16) cudaMemset(dC, 0x7f, (size_t)Mp * ldc * 4);
  else cudaMemcpy(dC, hC0.data(), hC0.size() * 4, cudaMemcpyHostToDevice);
  mvae_umma_operand A{dA, c.a_mn, c.M, c.K, lda, slabs, a_slab, slab, c.rb_a};
  mvae_umma_operand B{dB, c.b_mn, c.N, c.K, ldb, slabs, b_slab, slab, c.rb_b};
  mvae_umma_out D{dC, ldc, c.out_bf16, c.accumulate, c.bias ? dBias : nullptr, c.rb_out};
  int rc = mvae_umma_gemm(&A, &B, &D, c.M, c.N, c.K, c.bn, c.splits, 0, dErr, 0);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
  printf("case M=%d N=%d K=%d a_mn=%d b_mn=%d bn=%d splits=%d obf16=%d acc=%d bias=%d slabs=%d rb=%d%d%d : rc=%d cuda=%s errflag=%d",
         c.M, c.N, c.K, c.a_mn, c.b_mn, c.bn, c.splits, c.out_bf16, c.accumulate, c.bias, slabs, c.rb_a, c.rb_b, c.rb_out, rc,
         cudaGetErrorString(e), herr);
  int bad = (rc != 0 || e != cudaSuccess || herr != 0);
  if (rc != 0) printf(" [%s]", mvae_last_cuda_error());
  if (!bad) {
    std::vector<float> out((size_t)Mp * ldc);
    if (c.out_bf16) {
      std::vector<__nv_bfloat16> ob((size_t)Mp * ldc);
      cudaMemcpy(ob.data(), dC, ob.size() * 2, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < ob.size(); ++i) out[i] = __bfloat162float(ob[i]);
    } else {
      cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost);
    }
    // sample rows/cols when the full reference would be slow
    const long long full_cost = (long long)c.M * c.N * c.K;
    const int rstep = full_cost > (1ll << 31) ? 13 : 1, cstep = full_cost > (1ll << 31) ? 7 : 1;
    double maxerr = 0; long long nbad = 0, ncheck = 0; int fm = -1, fn = -1; double fgot = 0, fexp = 0;
    const float* pa = fA.data() + slab * a_slab; const float* pb = fB.data() + slab * b_slab;
    for (int m = 0; m < c.M; m += rstep)
      for (int n = 0; n < c.N; n += cstep) {
        double acc = 0;
        for (int k = 0; k < c.K; ++k) {
          float a = c.a_mn ? pa[IDX(c.rb_a, k, m, lda)] : pa[IDX(c.rb_a, m, k, lda)];
          float b = c.b_mn ? pb[IDX(c.rb_b, k, n, ldb)] : pb[IDX(c.rb_b, n, k, ldb)];
          acc += (double)a * b;
        }
        if (c.bias) acc += hBias[n];
        if (c.accumulate || c.splits > 1) acc += hC0[(size_t)m * ldc + n];
        double got = out[IDX(c.rb_out, m, n, ldc)];
        double tol = 2e-3 * sqrt((double)c.K) * 0.1 + (c.out_bf16 ? 0.01 * fabs(acc) + 1e-2 : 1e-4 * fabs(acc));
        double err = fabs(got - acc);
        if (err > maxerr) maxerr = err;
        ++ncheck;
        if (!(err <= tol)) { if (nbad == 0) { fm = m; fn = n; fgot = got; fexp = acc; } ++nbad; }
      }
    // untouched padding columns must keep their sentinel (fp32 non-accumulate path)
    printf(" maxerr=%.3e bad=%lld/%lld", maxerr, nbad, ncheck);
    if (nbad) {
      bad = 1;
      printf(" first bad (%d,%d) got %.5f exp %.5f", fm, fn, fgot, fexp);
      if (verbose) {
        printf("\n  row-block error map (rows/8 x cols/8, first 128x128; '.'=ok 'X'=bad)\n");
        for (int mb = 0; mb < 16 && mb * 8 < c.M; ++mb) {
          printf("  ");
          for (int nb = 0; nb < 16 && nb * 8 < c.N; ++nb) {
            int anybad = 0;
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) {
              int m = mb * 8 + i, n = nb * 8 + j;
              if (m >= c.M || n >= c.N || (m % rstep) || (n % cstep)) continue;
              double acc = 0;
              for (int k = 0; k < c.K; ++k) {
                float a = c.a_mn ? pa[IDX(c.rb_a, k, m, lda)] : pa[IDX(c.rb_a, m, k, lda)];
                float b = c.b_mn ? pb[IDX(c.rb_b, k, n, ldb)] : pb[IDX(c.rb_b, n, k, ldb)];
                acc += (double)a * b;
              }
              if (c.bias) acc += hBias[n];
              if (c.accumulate || c.splits > 1) acc += hC0[(size_t)m * ldc + n];
              if (fabs(out[IDX(c.rb_out, m, n, ldc)] - acc) > 0.05 + 0.02 * fabs(acc)) anybad = 1;
            }
            putchar(anybad ? 'X' : '.');
          }
          putchar('\n');
        }
      }
    }
  }
  printf(" -> %s\n", bad ? "FAIL" : "ok");
  fflush(stdout);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dBias); cudaFree(dErr);
  if (e != cudaSuccess) { printf("fatal CUDA error, stopping\n"); exit(3); }
  return bad;
}

static void bench_case(int M, int N, int K, int a_mn, int b_mn, int bn, int splits, int out_bf16, int rb = 0) {
  const long long a_rows = a_mn ? K : M, a_cols = a_mn ? M : K, b_rows = b_mn ? K : N, b_cols = b_mn ? N : K;
  __nv_bfloat16 *dA, *dB; void* dC; int* dErr;
  cudaMalloc(&dA, a_rows * a_cols * 2); cudaMalloc(&dB, b_rows * b_cols * 2);
  cudaMalloc(&dC, (size_t)M * N * 4); cudaMalloc(&dErr, 4);
  cudaMemset(dA, 0x11, a_rows * a_cols * 2); cudaMemset(dB, 0x11, b_rows * b_cols * 2); cudaMemset(dC, 0, (size_t)M * N * 4);
  cudaMemset(dErr, 0, 4);
  mvae_umma_operand A{dA, a_mn, M, K, a_cols, 1, 0, 0, rb & 1};
  mvae_umma_operand B{dB, b_mn, N, K, b_cols, 1, 0, 0, (rb >> 1) & 1};
  mvae_umma_out D{dC, N, out_bf16, 0, nullptr, (rb >> 2) & 1};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) mvae_umma_gemm(&A, &B, &D, M, N, K, bn, splits, 0, dErr, 0);
  cudaEventRecord(e0);
  const int iters = 10;
  for (int i = 0; i < iters; ++i) mvae_umma_gemm(&A, &B, &D, M, N, K, bn, splits, 0, dErr, 0);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  int herr = 0; cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
  printf("bench rb=%d M=%d N=%d K=%d a_mn=%d b_mn=%d bn=%d splits=%d obf16=%d : %.3f ms  %.1f TFLOP/s  (cuda=%s err=%d)\n", rb, M, N, K,
         a_mn, b_mn, bn, splits, out_bf16, ms, 2.0 * M * N * K / ms * 1e-9, cudaGetErrorString(e), herr);
  fflush(stdout);
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dErr);
}

int main(int argc, char** argv) {
  int stage = argc > 1 ? atoi(argv[1]) : 0;
  int fails = 0;
  // stage 0: the simplest possible case first (one tile, one k-block) so that a layout bug shows up in isolation
  std::vector<Case> cases = {
      {128, 64, 64, 0, 0, 64, 1, 0, 0, 0, 1},
      {128, 64, 64, 1, 0, 64, 1, 0, 0, 0, 1},
      {128, 64, 64, 0, 1, 64, 1, 0, 0, 0, 1},
      {128, 64, 64, 1, 1, 64, 1, 0, 0, 0, 1},
      {128, 128, 256, 0, 0, 128, 1, 0, 0, 0, 1},
      {128, 256, 256, 0, 0, 256, 1, 0, 0, 0, 1},
      {128, 192, 256, 0, 0, 192, 1, 0, 0, 0, 1},
      {128, 256, 256, 0, 1, 256, 1, 0, 0, 0, 1},
      {128, 256, 256, 1, 1, 256, 1, 0, 0, 0, 1},
      {128, 192, 128, 1, 1, 192, 1, 0, 0, 0, 1},
  };
  if (stage >= 1) {
    std::vector<Case> more = {
        {384, 512, 512, 0, 0, 256, 1, 0, 0, 1, 1},    // multi-tile, bias
        {384, 512, 512, 0, 0, 256, 1, 1, 0, 1, 1},    // bf16 out
        {300, 200, 500, 0, 0, 128, 1, 0, 0, 1, 1},    // ragged M,N,K (K%8==4 -> ld padded)
        {300, 200, 504, 0, 0, 0, 1, 1, 0, 0, 1},      // ragged, bf16 out, auto bn
        {256, 1536, 512, 0, 0, 256, 1, 1, 0, 1, 3},   // projection shape, slabs
        {256, 512, 1536, 0, 0, 256, 1, 0, 1, 0, 1},   // accumulate
        {256, 512, 1536, 0, 1, 256, 1, 0, 0, 0, 1},   // B MN-major (dX = dG * W)
        {1536, 512, 4096, 1, 1, 256, 8, 0, 0, 0, 1},  // wgrad: both MN-major, split-K
        {1503, 501, 2000, 1, 1, 256, 5, 0, 0, 0, 1},  // wgrad ragged (ld padded to 8)
        {1536, 512, 4096, 1, 1, 128, 3, 0, 0, 0, 2},
        {128, 35, 512, 0, 0, 64, 1, 0, 0, 1, 1},      // head-like skinny N
        {4096, 1536, 512, 0, 0, 192, 1, 1, 0, 1, 1},
        {40000, 1536, 512, 0, 0, 256, 1, 1, 0, 1, 1}, // persistent loop with many units per CTA
    };
    cases.insert(cases.end(), more.begin(), more.end());
  }
  if (stage == 3) {
    cases = {
        {128, 64, 64, 0, 0, 64, 1, 1, 0, 0, 1, 0, 0, 1},
        {384, 512, 512, 0, 0, 256, 1, 1, 0, 1, 2, 0, 0, 1},
        {256, 1536, 512, 0, 0, 256, 1, 1, 0, 1, 3, 0, 0, 1},
        {256, 512, 1536, 0, 1, 256, 1, 1, 0, 0, 1, 0, 0, 1},
        {4096, 512, 64, 0, 1, 256, 1, 1, 0, 0, 1, 0, 0, 1},
        {40000 / 32 * 32, 1536, 512, 0, 0, 256, 1, 1, 0, 1, 1, 0, 0, 1},
    };
  }
  for (auto& c : cases) fails += run_case(c, 1);
  printf("SUMMARY: %d failing of %zu\n", fails, cases.size());
  if (stage == 3 && fails == 0) {
    bench_case(491520, 1536, 512, 0, 0, 256, 1, 1, 0);
    bench_case(491520, 1536, 512, 0, 0, 256, 1, 1, 4);   // row-blocked output
    bench_case(491520, 512, 1536, 0, 0, 256, 1, 1, 4);
    return 0;
  }
  if (stage >= 2 && fails == 0) {
    bench_case(8192, 8192, 8192, 0, 0, 256, 1, 1);
    bench_case(8192, 8192, 8192, 0, 0, 128, 1, 1);
    bench_case(491520, 1536, 512, 0, 0, 256, 1, 1);   // GRU input projection over all timesteps
    bench_case(491520, 512, 1536, 0, 0, 256, 1, 1);   // dX
    bench_case(1536, 512, 491520, 1, 1, 256, 37, 0);  // wgrad
    bench_case(4096, 1536, 512, 0, 0, 192, 1, 1);     // one recurrent step
    bench_case(4096, 1536, 512, 0, 0, 256, 1, 1);
  }
  return fails ? 1 : 0;
}
