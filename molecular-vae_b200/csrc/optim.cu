// Flat-buffer optimiser kernels: global gradient-norm clipping
// (torch.nn.utils.clip_grad_norm, train.py:102 / train_distributed.py:91) and the Adam (train.py:81) /
// SGD-momentum (train_distributed.py:73) updates, over the single flat fp32 parameter / gradient buffers
// used for data-parallel training.  No host synchronisation: the clip coefficient stays on the device.
#include "../../include/mvae_b200.h"
#include "common.cuh"

namespace {
long long g_opt_launches = 0;

__global__ void sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double local = 0.0;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    local += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[n4 * 4 + threadIdx.x]; local += (double)v * v; }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(out, s);
  }
}
// coef = min(1, max_norm / (norm + 1e-6)) as torch does; norm_out optional
__global__ void clip_coef_kernel(const double* __restrict__ sumsq, float max_norm, float* __restrict__ coef,
                                 float* __restrict__ norm_out) {
  const float norm = (float)sqrt(sumsq[0]);
  const float c = max_norm / (norm + 1e-6f);
  coef[0] = c < 1.f ? c : 1.f;
  if (norm_out) norm_out[0] = norm;
}
__global__ void scale_kernel(float* __restrict__ g, long long n, const float* __restrict__ coef) {
  const float c = coef[0];
  if (c == 1.f) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) g[i] *= c;
}
// torch.optim.Adam (no amsgrad, L2 weight decay added to the gradient), gradient pre-multiplied by *coef (may be null)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float bias1, float bias2, const float* __restrict__ coef) {
  const float c = coef ? coef[0] : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * c;
    if (weight_decay != 0.f) gi = fmaf(weight_decay, p[i], gi);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bias2) + eps;
    p[i] -= (lr / bias1) * (mi / denom);
  }
}
// torch.optim.SGD with momentum (dampening 0, no nesterov): buf = mu*buf + g (buf = g on the first step); p -= lr*buf
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                    long long n, float lr, float momentum, float weight_decay, int first_step,
                                    const float* __restrict__ coef) {
  const float c = coef ? coef[0] : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * c;
    if (weight_decay != 0.f) gi = fmaf(weight_decay, p[i], gi);
    const float b = first_step ? gi : fmaf(momentum, buf[i], gi);
    buf[i] = b;
    p[i] -= lr * b;
  }
}
inline int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}
}  // namespace

extern "C" {

// scratch: device, 8 bytes (double) + 4 bytes (coef) at scratch+8; norm_out optional device float.
int mvae_clip_grad_norm(float* grads, long long n, float max_norm, void* scratch16, float* norm_out, int apply_scale,
                        mvae_stream_t stream) {
  if (!grads || n <= 0 || !scratch16 || (reinterpret_cast<uintptr_t>(grads) & 15)) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* sumsq = reinterpret_cast<double*>(scratch16);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch16) + 8);
  MVAE_CUDA_CHECK(cudaMemsetAsync(sumsq, 0, 8, st));
  sumsq_kernel<<<grid_for(n / 4 + 1), 256, 0, st>>>(grads, n, sumsq);
  clip_coef_kernel<<<1, 1, 0, st>>>(sumsq, max_norm, coef, norm_out);
  if (apply_scale) scale_kernel<<<grid_for(n), 256, 0, st>>>(grads, n, coef);
  MVAE_CUDA_CHECK(cudaGetLastError());
  g_opt_launches += 3 + (apply_scale ? 1 : 0);
  return MVAE_OK;
}

int mvae_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, const float* clip_coef,
                   mvae_stream_t stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return MVAE_ERR_INVALID;
  // bias corrections in double on the host, as torch.optim.Adam computes them (1 - beta ** step in Python floats)
  const float bias1 = (float)(1.0 - pow((double)beta1, (double)step)), bias2 = (float)(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr,
                                                                              beta1, beta2, eps, weight_decay, bias1,
                                                                              bias2, clip_coef);
  MVAE_CUDA_CHECK(cudaGetLastError());
  ++g_opt_launches;
  return MVAE_OK;
}

int mvae_sgd_momentum_step(float* params, const float* grads, float* momentum_buf, long long n, float lr,
                           float momentum, float weight_decay, int first_step, const float* clip_coef,
                           mvae_stream_t stream) {
  if (!params || !grads || !momentum_buf || n <= 0) return MVAE_ERR_INVALID;
  sgd_momentum_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, momentum_buf, n, lr,
                                                                                      momentum, weight_decay,
                                                                                      first_step, clip_coef);
  MVAE_CUDA_CHECK(cudaGetLastError());
  ++g_opt_launches;
  return MVAE_OK;
}

}  // extern "C"
