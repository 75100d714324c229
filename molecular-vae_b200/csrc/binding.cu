// BindingModel property head (mosesvae.py:6-25): an MLP on the latent code,
//   Linear(Z,256) -> BatchNorm1d(256) -> Tanh -> Linear(256,256) -> ReLU -> Linear(256,64) -> BatchNorm1d(64) -> ReLU -> Linear(64,1)
// forward and backward (autograd of the module) on one B200.  The callers that consume it are
// moses_train_distrib.py:274, trainbinding.py:216 and mosesanalyize.py:192.  Everything here is tiny next to the VAE
// step (0.12 MMAC per molecule), so it runs on the CUDA cores in fp32; BatchNorm uses batch statistics in train mode
// (biased variance for the normalisation, unbiased for the running estimate, torch semantics) and the running
// statistics in eval mode.
#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "umma_gemm.h"
#include "host_common.cuh"

namespace {

constexpr int H1 = 256, H2 = 256, H3 = 64;

struct BWS {
  float *a1, *t1, *r2, *a3, *r3, *xh1, *xh3, *stat1, *stat3;   // stat: [mean | inv_std] per column
  float *d3, *d2, *d1, *red;
  size_t total;
};
void carve(int B, int Z, void* base, BWS* w) {
  (void)Z;
  Carver c{reinterpret_cast<uint8_t*>(base), 0};
  const size_t b = B;
  w->a1 = c.take<float>(b * H1); w->t1 = c.take<float>(b * H1); w->xh1 = c.take<float>(b * H1);
  w->r2 = c.take<float>(b * H2);
  w->a3 = c.take<float>(b * H3); w->r3 = c.take<float>(b * H3); w->xh3 = c.take<float>(b * H3);
  w->stat1 = c.take<float>(2 * H1); w->stat3 = c.take<float>(2 * H3);
  w->d3 = c.take<float>(b * H3); w->d2 = c.take<float>(b * H2); w->d1 = c.take<float>(b * H1);
  w->red = c.take<float>(2 * H1);
  w->total = (c.off + 255) & ~size_t(255);
}

// one block per column: batch mean and biased variance -> stat[c] = mean, stat[C + c] = 1/sqrt(var + eps);
// train mode also updates the running estimates (momentum, unbiased variance); eval mode reads them instead.
__global__ void bn_stats_kernel(const float* __restrict__ x, int B, int C, float eps, float momentum, int train,
                                float* __restrict__ run_mean, float* __restrict__ run_var, float* __restrict__ stat) {
  const int c = blockIdx.x;
  __shared__ double red[64];
  if (!train) {
    if (threadIdx.x == 0) { stat[c] = run_mean[c]; stat[C + c] = rsqrtf(run_var[c] + eps); }
    return;
  }
  double s = 0.0, s2 = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) { const double v = x[(long long)b * C + c]; s += v; s2 += v * v; }
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red[32 + (threadIdx.x >> 5)] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, a2 = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += red[w]; a2 += red[32 + w]; }
    const double mean = a / B;
    double var = a2 / B - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[c] = (float)mean;
    stat[C + c] = (float)(1.0 / sqrt(var + (double)eps));
    if (run_mean) {
      const double unb = B > 1 ? var * B / (B - 1) : var;
      run_mean[c] = (float)((1.0 - momentum) * run_mean[c] + momentum * mean);
      run_var[c] = (float)((1.0 - momentum) * run_var[c] + momentum * unb);
    }
  }
}
// xhat = (x - mean) * inv_std ; y = act(gamma * xhat + beta)   act: 1 tanh, 2 relu
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ stat, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int B, int C, int act, float* __restrict__ xhat,
                                float* __restrict__ y) {
  const long long n = (long long)B * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float xh = (x[i] - stat[c]) * stat[C + c];
    const float v = fmaf(gamma[c], xh, beta[c]);
    xhat[i] = xh;
    y[i] = act == 1 ? tanhf(v) : fmaxf(v, 0.f);
  }
}
// dy (grad wrt the activation output y) -> dv = dy * act'(y) in place; per-column sums: red[c] = sum dv, red[C+c] = sum dv*xhat
__global__ void bn_bwd_reduce_kernel(float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ xhat, int B,
                                     int C, int act, float* __restrict__ red) {
  const int c = blockIdx.x;
  __shared__ double sh[64];
  double s = 0.0, s2 = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long i = (long long)b * C + c;
    const float yy = y[i];
    const float dv = dy[i] * (act == 1 ? (1.f - yy * yy) : (yy > 0.f ? 1.f : 0.f));
    dy[i] = dv;
    s += dv; s2 += (double)dv * xhat[i];
  }
  for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = s; sh[32 + (threadIdx.x >> 5)] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, a2 = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += sh[w]; a2 += sh[32 + w]; }
    red[c] = (float)a; red[C + c] = (float)a2;
  }
}
// dx = gamma * inv_std * (dv - mean(dv) - xhat * mean(dv * xhat))  (train)  |  gamma * inv_std * dv  (eval); dgamma, dbeta
__global__ void bn_bwd_apply_kernel(const float* dv /* may alias dx */, const float* __restrict__ xhat, const float* __restrict__ stat,
                                    const float* __restrict__ gamma, const float* __restrict__ red, int B, int C, int train,
                                    float* dx, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const long long n = (long long)B * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float g = gamma[c] * stat[C + c];
    dx[i] = train ? g * (dv[i] - red[c] / B - xhat[i] * red[C + c] / B) : g * dv[i];
    if (i < C) { dgamma[c] = red[C + c]; dbeta[c] = red[c]; }
  }
}
__global__ void relu_mask_kernel(const float* __restrict__ out, float* __restrict__ d, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(out[i] > 0.f)) d[i] = 0.f;
}

enum { P_W1 = 0, P_B1, P_G1, P_BE1, P_W2, P_B2, P_W3, P_B3, P_G3, P_BE3, P_W4, P_B4 };

int check(const mvae_binding_desc* d, void* ws, size_t ws_bytes, BWS* w) {
  if (!d || d->batch <= 0 || d->z_size <= 0 || d->bn_eps <= 0.f) return MVAE_ERR_INVALID;
  if (d->train && d->batch < 2) return MVAE_ERR_INVALID;   // torch raises for a single row in train mode
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) return MVAE_ERR_INVALID;
  carve(d->batch, d->z_size, ws, w);
  return ws_bytes < w->total ? MVAE_ERR_WORKSPACE : MVAE_OK;
}

}  // namespace

extern "C" {

size_t mvae_binding_workspace_bytes(const mvae_binding_desc* d) {
  if (!d || d->batch <= 0 || d->z_size <= 0) return 0;
  BWS w;
  carve(d->batch, d->z_size, nullptr, &w);
  return w.total;
}

int mvae_binding_forward(const mvae_binding_desc* d, const float* const* P, float* const* running, const float* z, float* out,
                         void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  BWS w;
  RC(check(d, workspace, workspace_bytes, &w));
  if (!P || !running || !z || !out) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = d->batch, Z = d->z_size, tr = d->train ? 1 : 0;
  RC(sg(st, z, Z, 1, P[P_W1], 1, Z, w.a1, H1, B, H1, Z, P[P_B1], simt::ACT_NONE, 0));
  bn_stats_kernel<<<H1, 256, 0, st>>>(w.a1, B, H1, d->bn_eps, d->bn_momentum, tr, running[0], running[1], w.stat1); KCHECK();
  bn_apply_kernel<<<grid_for((long long)B * H1), 256, 0, st>>>(w.a1, w.stat1, P[P_G1], P[P_BE1], B, H1, 1, w.xh1, w.t1); KCHECK();
  RC(sg(st, w.t1, H1, 1, P[P_W2], 1, H1, w.r2, H2, B, H2, H1, P[P_B2], simt::ACT_RELU, 0));
  RC(sg(st, w.r2, H2, 1, P[P_W3], 1, H2, w.a3, H3, B, H3, H2, P[P_B3], simt::ACT_NONE, 0));
  bn_stats_kernel<<<H3, 256, 0, st>>>(w.a3, B, H3, d->bn_eps, d->bn_momentum, tr, running[2], running[3], w.stat3); KCHECK();
  bn_apply_kernel<<<grid_for((long long)B * H3), 256, 0, st>>>(w.a3, w.stat3, P[P_G3], P[P_BE3], B, H3, 2, w.xh3, w.r3); KCHECK();
  RC(sg(st, w.r3, H3, 1, P[P_W4], 1, H3, out, 1, B, 1, H3, P[P_B4], simt::ACT_NONE, 0));
  return MVAE_OK;
}

int mvae_binding_backward(const mvae_binding_desc* d, const float* const* P, float* const* G, const float* z, const float* dout,
                          float* dz, void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  BWS w;
  RC(check(d, workspace, workspace_bytes, &w));
  if (!P || !G || !z || !dout) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int B = d->batch, Z = d->z_size, tr = d->train ? 1 : 0;
  // Linear(64,1)
  RC(sg_wgrad(st, dout, 1, 1, w.r3, H3, 1, G[P_W4], H3, 1, H3, B));
  RC(memset_async(G[P_B4], 4, st));
  RC(simt::colsum<float>(st, dout, B, 1, 1, G[P_B4])); mvae_count_launches(1);
  RC(sg(st, dout, 1, 1, P[P_W4], H3, 1, w.d3, H3, B, H3, 1, nullptr, simt::ACT_NONE, 0));
  // ReLU + BatchNorm1d(64)
  bn_bwd_reduce_kernel<<<H3, 256, 0, st>>>(w.d3, w.r3, w.xh3, B, H3, 2, w.red); KCHECK();
  bn_bwd_apply_kernel<<<grid_for((long long)B * H3), 256, 0, st>>>(w.d3, w.xh3, w.stat3, P[P_G3], w.red, B, H3, tr, w.d3, G[P_G3], G[P_BE3]); KCHECK();
  // Linear(256,64)
  RC(sg_wgrad(st, w.d3, 1, H3, w.r2, H2, 1, G[P_W3], H2, H3, H2, B));
  RC(memset_async(G[P_B3], (size_t)H3 * 4, st));
  RC(simt::colsum<float>(st, w.d3, B, H3, H3, G[P_B3])); mvae_count_launches(1);
  RC(sg(st, w.d3, H3, 1, P[P_W3], H2, 1, w.d2, H2, B, H2, H3, nullptr, simt::ACT_NONE, 0));
  // ReLU + Linear(256,256)
  relu_mask_kernel<<<grid_for((long long)B * H2), 256, 0, st>>>(w.r2, w.d2, (long long)B * H2); KCHECK();
  RC(sg_wgrad(st, w.d2, 1, H2, w.t1, H1, 1, G[P_W2], H1, H2, H1, B));
  RC(memset_async(G[P_B2], (size_t)H2 * 4, st));
  RC(simt::colsum<float>(st, w.d2, B, H2, H2, G[P_B2])); mvae_count_launches(1);
  RC(sg(st, w.d2, H2, 1, P[P_W2], H1, 1, w.d1, H1, B, H1, H2, nullptr, simt::ACT_NONE, 0));
  // Tanh + BatchNorm1d(256)
  bn_bwd_reduce_kernel<<<H1, 256, 0, st>>>(w.d1, w.t1, w.xh1, B, H1, 1, w.red); KCHECK();
  bn_bwd_apply_kernel<<<grid_for((long long)B * H1), 256, 0, st>>>(w.d1, w.xh1, w.stat1, P[P_G1], w.red, B, H1, tr, w.d1, G[P_G1], G[P_BE1]); KCHECK();
  // Linear(Z,256)
  RC(sg_wgrad(st, w.d1, 1, H1, z, Z, 1, G[P_W1], Z, H1, Z, B));
  RC(memset_async(G[P_B1], (size_t)H1 * 4, st));
  RC(simt::colsum<float>(st, w.d1, B, H1, H1, G[P_B1])); mvae_count_launches(1);
  if (dz) RC(sg(st, w.d1, H1, 1, P[P_W1], Z, 1, dz, Z, B, Z, H1, nullptr, simt::ACT_NONE, 0));
  return MVAE_OK;
}

}  // extern "C"
