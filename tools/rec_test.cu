// Standalone timing / trace tool for the persistent GRU recurrence kernel (run on the GPU box).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../molecular-vae_b200/csrc/gru_rec.h"
extern "C" const char* mvae_last_cuda_error(void);

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 1;
  const int B = argc > 2 ? atoi(argv[2]) : 4096;
  const int T = argc > 3 ? atoi(argv[3]) : 120;
  const int fast = argc > 4 ? atoi(argv[4]) : 0;
  const int debug = argc > 5 ? atoi(argv[5]) : 0;
  const int boxr = argc > 6 ? atoi(argv[6]) : 0;
  const int Hp = 512;
  const size_t slab = (size_t)B * Hp;
  __nv_bfloat16 *W, *WT, *gi, *hs, *sv, *dX, *dG; float* bhh; unsigned* ctr; int* err; unsigned long long* trace;
  cudaMalloc(&W, 3 * Hp * Hp * 2); cudaMalloc(&WT, 3 * Hp * Hp * 2);
  cudaMalloc(&gi, (size_t)T * B * 3 * Hp * 2); cudaMalloc(&hs, (T + 1) * slab * 2); cudaMalloc(&sv, T * slab * 5 * 2);
  cudaMalloc(&dX, T * slab * 2); cudaMalloc(&dG, T * slab * 4 * 2); cudaMalloc(&bhh, 3 * Hp * 4);
  cudaMalloc(&ctr, 4096); cudaMalloc(&err, 4); cudaMalloc(&trace, (size_t)T * 4 * 12 * 8);
  // small random-ish contents
  std::vector<__nv_bfloat16> h(3 * Hp * Hp);
  unsigned s = 1;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = __float2bfloat16(((s >> 8) & 0xFFFF) / 65536.0f * 0.08f - 0.04f); }
  cudaMemcpy(W, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(WT, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(gi, 0x3c, (size_t)T * B * 3 * Hp * 2); cudaMemset(hs, 0, (T + 1) * slab * 2);
  cudaMemset(dX, 0x30, T * slab * 2); cudaMemset(bhh, 0, 3 * Hp * 4); cudaMemset(err, 0, 4);
  cudaMemset(trace, 0, (size_t)T * 4 * 12 * 8);
  for (int bwd = 0; bwd < 2; ++bwd) {
    mvae_gru_rec_args a{};
    a.backward = bwd; a.variant = variant; a.Bp = B; a.Hp = Hp; a.T = T;
    a.W = bwd ? WT : W; a.gi = gi; a.gi_tstride = (long long)B * 3 * Hp; a.bhh = bhh; a.hs = hs; a.sv = sv; a.dX = dX; a.dG = dG;
    a.counters = ctr; a.err_flag = err; a.trace = nullptr; a.debug = debug; a.ones_col = -1; a.a_box_rows = boxr;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto launch = [&](const mvae_gru_rec_args* aa) { return variant >= 3 ? mvae_gru_rec2_launch(aa, fast, 0) : mvae_gru_rec_launch(aa, 0); };
    if (variant >= 3) printf("max active clusters (%s): cl2=%d cl4=%d cl8=%d\n", bwd ? "bwd" : "fwd", mvae_gru_rec2_max_clusters(bwd, 2), mvae_gru_rec2_max_clusters(bwd, 4), mvae_gru_rec2_max_clusters(bwd, 8));
    int rc = launch(&a);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) rc |= launch(&a);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    int herr = 0; cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    printf("%s variant %d B=%d T=%d: rc=%d cuda=%s err=%d  %.3f ms  (%.2f us/step, %.1f TFLOP/s)\n", bwd ? "bwd" : "fwd", variant,
           B, T, rc, cudaGetErrorString(e), herr, ms, ms * 1e3 / T, 2.0 * B * 3 * Hp * Hp * (T - 1) / ms * 1e-9);
    if (rc) printf("  %s\n", mvae_last_cuda_error());
    // traced run
    a.trace = trace;
    launch(&a);
    cudaDeviceSynchronize();
    const int NT = variant == 2 ? 4 : 2;
    const int NS = variant >= 3 ? 12 : 8;
    std::vector<unsigned long long> tr((size_t)T * NT * NS);
    cudaMemcpy(tr.data(), trace, tr.size() * 8, cudaMemcpyDeviceToHost);
    printf("  trace CTA(0,0), ns relative to counter-ok of (step, tile 0): [ctr_ok, tma_issued, first_full, mma_issued, tfull_seen, epi_done, fenced, red_done]\n");
    for (int step : {1, 2, 60, 61}) {
      if (step >= T) continue;
      unsigned long long base = tr[((size_t)step * NT + 0) * NS + 0];
      for (int i = 0; i < NT; ++i) {
        printf("  step %3d tile %d:", step, i);
        for (int k = 0; k < NS; ++k) printf(" %7lld", (long long)(tr[((size_t)step * NT + i) * NS + k] - base));
        printf("\n");
      }
    }
    fflush(stdout);
  }
  if (variant == 32) {
    // cross-check: BPTT of the K-split variant against the pair kernel on identical inputs
    const size_t n = (size_t)T * B * 4 * Hp;
    {
      std::vector<__nv_bfloat16> hx((size_t)T * slab);
      unsigned s2 = 7;
      for (auto& v : hx) { s2 = s2 * 1664525u + 1013904223u; v = __float2bfloat16(((s2 >> 8) & 0xFFFF) / 65536.0f - 0.5f); }
      cudaMemcpy(dX, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    }
    std::vector<__nv_bfloat16> ref(n), got(n);
    mvae_gru_rec_args a{};
    a.backward = 1; a.Bp = B; a.Hp = Hp; a.T = T; a.W = WT; a.gi = gi; a.gi_tstride = (long long)B * 3 * Hp; a.bhh = bhh;
    a.hs = hs; a.sv = sv; a.dX = dX; a.dG = dG; a.counters = ctr; a.err_flag = err; a.ones_col = -1;
    for (int pass = 0; pass < 2; ++pass) {
      a.variant = pass ? 32 : 3;
      cudaMemset(dG, 0, n * 2);
      mvae_gru_rec2_launch(&a, 0, 0);
      cudaDeviceSynchronize();
      cudaMemcpy(pass ? got.data() : ref.data(), dG, n * 2, cudaMemcpyDeviceToHost);
    }
    size_t bad = 0; double maxd = 0, maxr = 0;
    std::vector<size_t> by_t(T, 0), by_pair(8, 0), by_rowblk(B / 128, 0), by_blk(4, 0);
    for (size_t i = 0; i < n; ++i) {
      const float r = __bfloat162float(ref[i]), g = __bfloat162float(got[i]);
      const double d = fabs((double)r - g);
      if (fabs(r) > maxr) maxr = fabs(r);
      if (d > maxd) maxd = d;
      if (d > 2e-2 * fabs(r) + 2e-3) {
        if (bad < 12) {
          const size_t t = i / ((size_t)B * 4 * Hp), row = (i / (4 * Hp)) % B, col = i % (4 * Hp);
          printf("  mismatch t=%zu row=%zu col=%zu (block %zu unit %zu): pair %g ksplit %g\n", t, row, col, col / Hp, col % Hp, r, g);
        }
        ++bad;
        const size_t t = i / ((size_t)B * 4 * Hp), row = (i / (4 * Hp)) % B, col = i % (4 * Hp);
        by_t[t]++; by_pair[(col % Hp) / 64]++; by_rowblk[row / 128]++; by_blk[col / Hp]++;
      }
    }
    printf("  by t:"); for (auto v : by_t) printf(" %zu", v);
    printf("\n  by pair:"); for (auto v : by_pair) printf(" %zu", v);
    printf("\n  by row block (128):"); for (auto v : by_rowblk) printf(" %zu", v);
    printf("\n  by dG block:"); for (auto v : by_blk) printf(" %zu", v);
    printf("\n");
    printf("K-split vs pair kernel dG: %zu mismatching of %zu, max |diff| %g, max |ref| %g\n", bad, n, maxd, maxr);
  }
  return 0;
}
