// CUDA-core kernels of the ELBO path: fp32 strided SGEMM (check mode + the tiny encoder / latent GEMMs),
// the conv encoder over one-hot SMILES, reparametrise + KL, GRU gate math (forward and BPTT), the
// softmax + BCE head, and small reductions.  Everything here is HBM/L2-bound elementwise or <1 % of the
// step's FLOPs; the dense contractions of bf16 mode live in umma_gemm.cu / the fused recurrent kernel.
#pragma once
#include "common.cuh"

namespace simt {
namespace {   // internal linkage: this header is included by several translation units

constexpr float SELU_ALPHA = 1.6732632423543772848170429916717f;
constexpr float SELU_SCALE = 1.0507009873554804934193349852946f;

enum Act { ACT_NONE = 0, ACT_SELU = 1, ACT_RELU = 2 };

__device__ __forceinline__ float selu_f(float x) {
  return SELU_SCALE * (x > 0.f ? x : SELU_ALPHA * expm1f(x));
}
// selu'(a) expressed through y = selu(a):  a > 0 -> scale ;  else scale*alpha*exp(a) = y + scale*alpha
__device__ __forceinline__ float selu_grad_from_out(float y) {
  return y > 0.f ? SELU_SCALE : y + SELU_SCALE * SELU_ALPHA;
}

// ------------------------------------------------------------------------------------------
// SGEMM:  C[m][n] (+)= sum_k A(m,k) * B(k,n) (+ bias[n]) (act)      A(m,k) = A[m*sam + k*sak]
//                                                                    B(k,n) = B[k*sbk + n*sbn]
// 64x64x16 tiles, 256 threads, 4x4 micro-tile, optional split-K over blockIdx.z (atomicAdd).
// ------------------------------------------------------------------------------------------
struct SgemmArgs {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  float* C; long long ldc;
  const float* bias;
  int M, N, K;
  int act;
  int accumulate;
  int k_per_split;
};

__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmArgs a) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * a.k_per_split;
  const int kend = min(a.K, kbeg + a.k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // micro-tile coords: rows ty*4.., cols tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (a.sak == 1);
  const bool b_kfast = (a.sbk == 1);
  // 16-byte loads along the contiguous dimension for full, aligned tiles (k-fast: 4 consecutive k of one row; otherwise 4
  // consecutive rows / columns of one k); edge tiles and odd strides take the scalar path below
  const bool a_vec = ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) && m0 + 64 <= a.M &&
                     (a_kfast ? (a.sam % 4 == 0) : (a.sam == 1 && a.sak % 4 == 0));
  const bool b_vec = ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0) && n0 + 64 <= a.N &&
                     (b_kfast ? (a.sbn % 4 == 0) : (a.sbn == 1 && a.sbk % 4 == 0));
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    const bool kfull = k0 + 16 <= kend;
    if (a_vec && kfull) {
      if (a_kfast) {
        const int m = tid >> 2, k4 = (tid & 3) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.A + (long long)(m0 + m) * a.sam + (k0 + k4)));
        As[k4][m] = v.x; As[k4 + 1][m] = v.y; As[k4 + 2][m] = v.z; As[k4 + 3][m] = v.w;
      } else {
        const int k = tid >> 4, m4 = (tid & 15) * 4;
        *reinterpret_cast<float4*>(&As[k][m4]) = __ldg(reinterpret_cast<const float4*>(a.A + (m0 + m4) + (long long)(k0 + k) * a.sak));
      }
    }
    if (b_vec && kfull) {
      if (b_kfast) {
        const int n = tid >> 2, k4 = (tid & 3) * 4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.B + (long long)(n0 + n) * a.sbn + (k0 + k4)));
        Bs[k4][n] = v.x; Bs[k4 + 1][n] = v.y; Bs[k4 + 2][n] = v.z; Bs[k4 + 3][n] = v.w;
      } else {
        const int k = tid >> 4, n4 = (tid & 15) * 4;
        *reinterpret_cast<float4*>(&Bs[k][n4]) = __ldg(reinterpret_cast<const float4*>(a.B + (n0 + n4) + (long long)(k0 + k) * a.sbk));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (a_kfast) { k = tid & 15; m = (tid >> 4) + 16 * i; }
      else { m = tid & 63; k = (tid >> 6) + 4 * i; }
      const int gm = m0 + m, gk = k0 + k;
      if (!(a_vec && kfull)) As[k][m] = (gm < a.M && gk < kend) ? __ldg(a.A + gm * a.sam + gk * a.sak) : 0.f;
      int n, kk;
      if (b_kfast) { kk = tid & 15; n = (tid >> 4) + 16 * i; }
      else { n = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int gn = n0 + n, gk2 = k0 + kk;
      if (!(b_vec && kfull)) Bs[kk][n] = (gn < a.N && gk2 < kend) ? __ldg(a.B + gk2 * a.sbk + gn * a.sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float v = acc[i][j];
      float* c = a.C + gm * a.ldc + gn;
      if (split) {
        if (a.bias && blockIdx.z == 0) v += a.bias[gn];
        atomicAdd(c, v);
      } else {
        if (a.bias) v += a.bias[gn];
        if (a.accumulate) v += *c;
        if (a.act == ACT_SELU) v = selu_f(v);
        else if (a.act == ACT_RELU) v = fmaxf(v, 0.f);
        *c = v;
      }
    }
  }
}

inline int sgemm(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk,
                 long long sbn, float* C, long long ldc, int M, int N, int K, const float* bias, int act,
                 int accumulate, int splits) {
  if (M <= 0 || N <= 0 || K <= 0) return MVAE_OK;
  SgemmArgs a{A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, act, accumulate, K};
  if (splits < 1) splits = 1;
  int kps = round_up(ceil_div(K, splits), 16);
  splits = ceil_div(K, kps);
  a.k_per_split = kps;
  if (splits > 1 && act != ACT_NONE) return MVAE_ERR_INVALID;
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64), splits);
  sgemm_kernel<<<grid, 256, 0, st>>>(a);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

// ------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------
// float one-hot (B,T,C) -> u8 ids; flags rows that are not exactly one-hot (the kernels are id-driven).
__global__ void onehot_to_ids_kernel(const float* __restrict__ x, int rows, int C, uint8_t* __restrict__ ids,
                                     int* __restrict__ bad_flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* p = x + (long long)r * C;
  int arg = -1, ones = 0, other = 0;
  for (int c = 0; c < C; ++c) {
    const float v = p[c];
    if (v == 1.0f) { ++ones; if (arg < 0) arg = c; }
    else if (v != 0.0f) ++other;
  }
  if (ones != 1 || other != 0) { atomicExch(bad_flag, 1); arg = arg < 0 ? 0 : arg; }
  ids[r] = (uint8_t)arg;
}

// dst[rows_p][cols_p] (T) <- src[rows][cols] (fp32), zero padded; optional row-block permutation used to
// lay GRU gate blocks out as padded [g][Hp]:  dst row = perm_block(g)*Hp_rows + j  for src row g*H + j.
template <typename T>
__global__ void pad_gate_matrix_kernel(const float* __restrict__ src, int H, int cols, T* __restrict__ dst, int Hp,
                                       int cols_p, int g0, int g1, int g2) {
  // src: [3H][cols]  dst: [3Hp][cols_p]; dst gate block order (g0,g1,g2) names the SOURCE gate of each block
  const long long total = 3ll * Hp * cols_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols_p);
    const int r = (int)(i / cols_p);
    const int blk = r / Hp, j = r - blk * Hp;
    const int g = blk == 0 ? g0 : (blk == 1 ? g1 : g2);
    float v = 0.f;
    if (j < H && c < cols) v = src[((long long)g * H + j) * cols + c];
    dst[i] = from_f32<T>(v);
  }
}
// inverse for gradients: dst[3H][cols] fp32 <- src[3Hp][cols_p] fp32 (src block order names gate of each block)
__global__ void unpad_gate_matrix_kernel(const float* __restrict__ src, int Hp, int cols_p, float* __restrict__ dst,
                                         int H, int cols, int g0, int g1, int g2) {
  const long long total = 3ll * H * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int r = (int)(i / cols);
    const int g = r / H, j = r - g * H;
    const int blk = (g == g0) ? 0 : ((g == g1) ? 1 : 2);
    dst[i] = src[((long long)blk * Hp + j) * cols_p + c];
  }
}
template <typename T>
__global__ void pad_matrix_kernel(const float* __restrict__ src, int rows, int cols, T* __restrict__ dst, int rows_p,
                                  int cols_p) {
  const long long total = (long long)rows_p * cols_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols_p);
    const int r = (int)(i / cols_p);
    dst[i] = from_f32<T>((r < rows && c < cols) ? src[(long long)r * cols + c] : 0.f);
  }
}
__global__ void unpad_matrix_kernel(const float* __restrict__ src, int cols_p, float* __restrict__ dst, int rows,
                                    int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int r = (int)(i / cols);
    dst[i] = src[(long long)r * cols_p + c];
  }
}
// padded gate bias: dst[3Hp] <- src[3H]
__global__ void pad_gate_vector_kernel(const float* __restrict__ src, int H, float* __restrict__ dst, int Hp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * Hp) return;
  const int g = i / Hp, j = i - g * Hp;
  dst[i] = j < H ? src[g * H + j] : 0.f;
}

// column sums of a [rows][ld] matrix (T) over rows -> fp32 out[cols]   (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long rows, long long ld, int cols, float* __restrict__ out,
                              int rows_per_block) {
  // block (32 x 8): 32 consecutive columns, 8 row lanes
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < cols)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += to_f32<T>(x[r * ld + c]);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}
// packed sequences: the rows form slabs of rps rows (blockIdx.y = slab); only the first lim[slab] rows of a slab count
// (the rest may hold anything, it is not read)
template <typename T>
__global__ void colsum_slabs_kernel(const T* __restrict__ x, int rps, long long ld, int cols, float* __restrict__ out,
                                    const int* __restrict__ lim) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int n = min(rps, __ldg(lim + blockIdx.y));
  // blockIdx.z: chunk of the slab's rows (the active part of a slab is split over gridDim.z blocks)
  const int per = (rps + gridDim.z - 1) / gridDim.z;
  const int r0 = blockIdx.z * per, r1 = min(n, r0 + per);
  const T* xs = x + (long long)blockIdx.y * rps * ld;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < cols) {
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {      // four independent loads in flight per thread
      s0 += to_f32<T>(xs[(long long)r * ld + c]);
      s1 += to_f32<T>(xs[(long long)(r + 8) * ld + c]);
      s2 += to_f32<T>(xs[(long long)(r + 16) * ld + c]);
      s3 += to_f32<T>(xs[(long long)(r + 24) * ld + c]);
    }
    for (; r < r1; r += 8) s0 += to_f32<T>(xs[(long long)r * ld + c]);
  }
  const float s = (s0 + s1) + (s2 + s3);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols && r1 > r0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}
// bf16 fast path of the slab-wise column sums: 16-byte loads (one thread = 8 consecutive columns, a warp = 256 columns of one
// row), 8 row lanes per block, four rows in flight per thread.  cols % 256 == 0, ld % 8 == 0, x 16-byte aligned.
__global__ void __launch_bounds__(256) colsum_slabs_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, int rps, long long ld, int cols,
                                                                   float* __restrict__ out, const int* __restrict__ lim) {
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int c8 = blockIdx.x * 256 + lane * 8;
  const int n = lim ? min(rps, __ldg(lim + blockIdx.y)) : rps;   // lim == nullptr: every row of the (single) slab counts
  const int per = (rps + gridDim.z - 1) / gridDim.z;
  const int r0 = blockIdx.z * per, r1 = min(n, r0 + per);
  if (r1 <= r0) return;                                        // uniform over the block
  const __nv_bfloat16* xs = x + (long long)blockIdx.y * rps * ld + c8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto add = [&](const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[2 * k] += __uint_as_float(w[k] << 16); s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u); }
  };
  int r = r0 + wy;
  for (; r + 24 < r1; r += 32) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(xs + (long long)r * ld));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(xs + (long long)(r + 8) * ld));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(xs + (long long)(r + 16) * ld));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(xs + (long long)(r + 24) * ld));
    add(a); add(b); add(c); add(d);
  }
  for (; r < r1; r += 8) add(__ldg(reinterpret_cast<const uint4*>(xs + (long long)r * ld)));
#pragma unroll
  for (int k = 0; k < 8; ++k) red[wy][lane * 8 + k] = s[k];
  __syncthreads();
  const int c = threadIdx.x;                                   // 256 threads -> 256 columns of the block
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i][c];
  atomicAdd(out + blockIdx.x * 256 + c, t);
}
template <typename T>
inline int colsum(cudaStream_t st, const T* x, long long rows, long long ld, int cols, float* out_zeroed,
                  const int* lim = nullptr, int rps = 1) {
  if (lim) {
    bool fast = false;
    if constexpr (sizeof(T) == 2) {
      if (cols % 256 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        dim3 grid(cols / 256, (unsigned)(rows / rps), (unsigned)(rps >= 1024 ? 4 : 1));
        colsum_slabs_bf16x8_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), rps, ld, cols, out_zeroed, lim);
        fast = true;
      }
    }
    if (!fast) {
      dim3 grid(ceil_div(cols, 32), (unsigned)(rows / rps), (unsigned)(rps >= 1024 ? 8 : 1));
      colsum_slabs_kernel<T><<<grid, dim3(32, 8), 0, st>>>(x, rps, ld, cols, out_zeroed, lim);
    }
  } else {
    if constexpr (sizeof(T) == 2) {
      if (cols % 256 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && rows >= 4096 && rows < (1ll << 31)) {
        dim3 grid(cols / 256, 1, (unsigned)ceil_div64(rows, 1024));
        colsum_slabs_bf16x8_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), (int)rows, ld, cols, out_zeroed, nullptr);
        MVAE_CUDA_CHECK(cudaGetLastError());
        return MVAE_OK;
      }
    }
    const int rpb = rows >= 65536 ? 2048 : 256;   // small inputs (per-molecule bias sums): enough blocks to fill the GPU
    dim3 grid(ceil_div(cols, 32), (unsigned)ceil_div64(rows, rpb));
    colsum_kernel<T><<<grid, dim3(32, 8), 0, st>>>(x, rows, ld, cols, out_zeroed, rpb);
  }
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

// ------------------------------------------------------------------------------------------
// Encoder convs (models2d.py:23-27).  NB the Keras-port quirk: conv CHANNELS are the T sequence positions
// and the conv LENGTH axis is the charset.  The input is one-hot, so conv1 is a gather-sum of W1 columns:
//   a1[oc][l] = b1[oc] + sum_ic W1[oc][ic][ids[ic]-l]   for 0 <= ids[ic]-l < K1.
// One block per molecule (grid-stride), W2/W3 and the per-molecule activations live in shared memory.
// ------------------------------------------------------------------------------------------
struct ConvDims { int T, C, L1, L2, L3; };  // L1=C-8, L2=L1-8, L3=L2-10 ; channels 9,9,10 ; kernels 9,9,11

__global__ void __launch_bounds__(256) enc_conv_fwd_kernel(const uint8_t* __restrict__ ids, int B, ConvDims d,
                                                           const float* __restrict__ W1, const float* __restrict__ b1,
                                                           const float* __restrict__ W2, const float* __restrict__ b2,
                                                           const float* __restrict__ W3, const float* __restrict__ b3,
                                                           float* __restrict__ h1o, float* __restrict__ h2o,
                                                           float* __restrict__ h3o) {
  extern __shared__ float sm[];
  float* sW2 = sm;                 // 9*9*9
  float* sW3 = sW2 + 729;          // 10*9*11
  float* sh1 = sW3 + 990;          // 9*L1
  float* sh2 = sh1 + 9 * d.L1;     // 9*L2
  uint8_t* sid = reinterpret_cast<uint8_t*>(sh2 + 9 * d.L2);  // T
  for (int i = threadIdx.x; i < 729; i += blockDim.x) sW2[i] = W2[i];
  for (int i = threadIdx.x; i < 990; i += blockDim.x) sW3[i] = W3[i];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < d.T; i += blockDim.x) sid[i] = ids[(long long)b * d.T + i];
    __syncthreads();
    for (int o = threadIdx.x; o < 9 * d.L1; o += blockDim.x) {
      const int oc = o / d.L1, l = o - oc * d.L1;
      float acc = b1[oc];
      const float* w = W1 + (long long)oc * d.T * 9;
      for (int ic = 0; ic < d.T; ++ic) {
        const int k = (int)sid[ic] - l;
        if (k >= 0 && k < 9) acc += __ldg(w + ic * 9 + k);
      }
      acc = fmaxf(acc, 0.f);
      sh1[o] = acc;
      h1o[(long long)b * 9 * d.L1 + o] = acc;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 9 * d.L2; o += blockDim.x) {
      const int oc = o / d.L2, l = o - oc * d.L2;
      float acc = b2[oc];
      for (int ic = 0; ic < 9; ++ic)
#pragma unroll
        for (int k = 0; k < 9; ++k) acc = fmaf(sW2[(oc * 9 + ic) * 9 + k], sh1[ic * d.L1 + l + k], acc);
      acc = fmaxf(acc, 0.f);
      sh2[o] = acc;
      h2o[(long long)b * 9 * d.L2 + o] = acc;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 10 * d.L3; o += blockDim.x) {
      const int oc = o / d.L3, l = o - oc * d.L3;
      float acc = b3[oc];
      for (int ic = 0; ic < 9; ++ic)
#pragma unroll
        for (int k = 0; k < 11; ++k) acc = fmaf(sW3[(oc * 9 + ic) * 11 + k], sh2[ic * d.L2 + l + k], acc);
      h3o[(long long)b * 10 * d.L3 + o] = fmaxf(acc, 0.f);
    }
  }
}

// Backward of the three convs.  Each block walks its molecules, accumulates dW/db in shared memory and
// flushes once with atomics (dW1 alone is 9*T*9 floats = 38.9 KB at T=120).
__global__ void __launch_bounds__(256) enc_conv_bwd_kernel(const uint8_t* __restrict__ ids, int B, ConvDims d,
                                                           const float* __restrict__ W2, const float* __restrict__ W3,
                                                           const float* __restrict__ h1, const float* __restrict__ h2,
                                                           const float* __restrict__ h3,
                                                           const float* __restrict__ dflat,
                                                           float* __restrict__ dW1, float* __restrict__ db1,
                                                           float* __restrict__ dW2, float* __restrict__ db2,
                                                           float* __restrict__ dW3, float* __restrict__ db3) {
  extern __shared__ float sm[];
  const int nW1 = 9 * d.T * 9;
  float* sdW1 = sm;                 // nW1
  float* sdW2 = sdW1 + nW1;         // 729
  float* sdW3 = sdW2 + 729;         // 990
  float* sdb = sdW3 + 990;          // 9+9+10 (pad 32)
  float* sW2 = sdb + 32;            // 729
  float* sW3 = sW2 + 729;           // 990
  float* sh1 = sW3 + 990;           // 9*L1
  float* sh2 = sh1 + 9 * d.L1;      // 9*L2
  float* sda3 = sh2 + 9 * d.L2;     // 10*L3
  float* sda2 = sda3 + 10 * d.L3;   // 9*L2
  float* sda1 = sda2 + 9 * d.L2;    // 9*L1
  uint8_t* sid = reinterpret_cast<uint8_t*>(sda1 + 9 * d.L1);
  for (int i = threadIdx.x; i < nW1 + 729 + 990 + 32; i += blockDim.x) sm[i] = 0.f;
  for (int i = threadIdx.x; i < 729; i += blockDim.x) sW2[i] = W2[i];
  for (int i = threadIdx.x; i < 990; i += blockDim.x) sW3[i] = W3[i];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < d.T; i += blockDim.x) sid[i] = ids[(long long)b * d.T + i];
    for (int i = threadIdx.x; i < 9 * d.L1; i += blockDim.x) sh1[i] = h1[(long long)b * 9 * d.L1 + i];
    for (int i = threadIdx.x; i < 9 * d.L2; i += blockDim.x) sh2[i] = h2[(long long)b * 9 * d.L2 + i];
    for (int i = threadIdx.x; i < 10 * d.L3; i += blockDim.x) {
      const float h = h3[(long long)b * 10 * d.L3 + i];
      sda3[i] = h > 0.f ? dflat[(long long)b * 10 * d.L3 + i] : 0.f;
    }
    __syncthreads();
    // conv3: dW3[oc][ic][k] += sum_l da3[oc][l] * h2[ic][l+k] ; db3 ; dh2[ic][m] = sum_{oc,k} da3[oc][m-k] W3[oc][ic][k]
    for (int i = threadIdx.x; i < 990; i += blockDim.x) {
      const int oc = i / 99, rem = i - oc * 99, ic = rem / 11, k = rem - ic * 11;
      float s = 0.f;
      for (int l = 0; l < d.L3; ++l) s = fmaf(sda3[oc * d.L3 + l], sh2[ic * d.L2 + l + k], s);
      sdW3[i] += s;
    }
    if (threadIdx.x < 10) {
      float s = 0.f;
      for (int l = 0; l < d.L3; ++l) s += sda3[threadIdx.x * d.L3 + l];
      sdb[18 + threadIdx.x] += s;
    }
    for (int i = threadIdx.x; i < 9 * d.L2; i += blockDim.x) {
      const int ic = i / d.L2, m = i - ic * d.L2;
      float s = 0.f;
      for (int oc = 0; oc < 10; ++oc)
        for (int k = 0; k < 11; ++k) {
          const int l = m - k;
          if (l >= 0 && l < d.L3) s = fmaf(sda3[oc * d.L3 + l], sW3[(oc * 9 + ic) * 11 + k], s);
        }
      sda2[i] = sh2[i] > 0.f ? s : 0.f;
    }
    __syncthreads();
    // conv2
    for (int i = threadIdx.x; i < 729; i += blockDim.x) {
      const int oc = i / 81, rem = i - oc * 81, ic = rem / 9, k = rem - ic * 9;
      float s = 0.f;
      for (int l = 0; l < d.L2; ++l) s = fmaf(sda2[oc * d.L2 + l], sh1[ic * d.L1 + l + k], s);
      sdW2[i] += s;
    }
    if (threadIdx.x < 9) {
      float s = 0.f;
      for (int l = 0; l < d.L2; ++l) s += sda2[threadIdx.x * d.L2 + l];
      sdb[9 + threadIdx.x] += s;
    }
    for (int i = threadIdx.x; i < 9 * d.L1; i += blockDim.x) {
      const int ic = i / d.L1, m = i - ic * d.L1;
      float s = 0.f;
      for (int oc = 0; oc < 9; ++oc)
        for (int k = 0; k < 9; ++k) {
          const int l = m - k;
          if (l >= 0 && l < d.L2) s = fmaf(sda2[oc * d.L2 + l], sW2[(oc * 9 + ic) * 9 + k], s);
        }
      sda1[i] = sh1[i] > 0.f ? s : 0.f;
    }
    __syncthreads();
    // conv1 (one-hot input): dW1[oc][ic][k] += da1[oc][ids[ic]-k]
    for (int i = threadIdx.x; i < nW1; i += blockDim.x) {
      const int oc = i / (d.T * 9), rem = i - oc * d.T * 9, ic = rem / 9, k = rem - ic * 9;
      const int l = (int)sid[ic] - k;
      if (l >= 0 && l < d.L1) sdW1[i] += sda1[oc * d.L1 + l];
    }
    if (threadIdx.x < 9) {
      float s = 0.f;
      for (int l = 0; l < d.L1; ++l) s += sda1[threadIdx.x * d.L1 + l];
      sdb[threadIdx.x] += s;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nW1; i += blockDim.x) if (sdW1[i] != 0.f) atomicAdd(dW1 + i, sdW1[i]);
  for (int i = threadIdx.x; i < 729; i += blockDim.x) atomicAdd(dW2 + i, sdW2[i]);
  for (int i = threadIdx.x; i < 990; i += blockDim.x) atomicAdd(dW3 + i, sdW3[i]);
  if (threadIdx.x < 9) atomicAdd(db1 + threadIdx.x, sdb[threadIdx.x]);
  else if (threadIdx.x < 18) atomicAdd(db2 + threadIdx.x - 9, sdb[threadIdx.x]);
  else if (threadIdx.x < 28) atomicAdd(db3 + threadIdx.x - 18, sdb[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------
// reparametrise (models2d.py:31-38 / models.py:92-94) + the swapped KL of train.py:36-37
//   z = mu + eps_scale*eps*exp(logvar/2) (train) | mu ;  kl_sum += sum(1 + mu - logvar^2 - exp(mu))
// ------------------------------------------------------------------------------------------
__global__ void reparam_kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                  const float* __restrict__ eps, float eps_scale, int train, long long n,
                                  float* __restrict__ z, double* __restrict__ kl_sum) {
  double local = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], l = lv[i];
    z[i] = train ? fmaf(eps_scale * eps[i], expf(0.5f * l), m) : m;
    local += (double)(1.0f + m - l * l - expf(m));
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    if (kl_sum) atomicAdd(kl_sum, s);
  }
}
// d(mu), d(logvar): from z (dz), the swapped-KL term (scaled by kl_scale = 1/(B*Z), 0 to disable) and
// optional external upstream grads (the drop-in autograd path).
__global__ void reparam_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                      const float* __restrict__ eps, float eps_scale, int train,
                                      const float* __restrict__ dz, float kl_scale, const float* __restrict__ ext_dmu,
                                      const float* __restrict__ ext_dlv, long long n, float* __restrict__ dmu,
                                      float* __restrict__ dlv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], l = lv[i], g = dz[i];
    float gm = g + kl_scale * (-0.5f * (1.0f - expf(m)));
    float gl = kl_scale * l;
    if (train) gl += g * (eps_scale * eps[i]) * 0.5f * expf(0.5f * l);
    if (ext_dmu) gm += ext_dmu[i];
    if (ext_dlv) gl += ext_dlv[i];
    dmu[i] = gm;
    dlv[i] = gl;
  }
}
// y <- dy * selu'(out)     (in place on dy allowed)
__global__ void selu_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dy, float* __restrict__ dx,
                                long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * selu_grad_from_out(out[i]);
}

// ------------------------------------------------------------------------------------------
// GRU gate math, one thread per (row b, hidden unit j) of a [Bp][Hp] slab (torch.nn.GRU, gates r,z,n):
//   r = s(gi_r+gh_r)  z = s(gi_z+gh_z)  n = tanh(gi_n + r*gh_n)  h' = (1-z) n + z h
// gi: per-step [Bp][3Hp] (TG) or the time-invariant layer-0 projection (fp32, same shape, reused each t).
// Saves (r,z,n,gh_n) for BPTT into sv[Bp][4Hp].
// ------------------------------------------------------------------------------------------
template <typename TA, typename TG>
__global__ void gru_gate_fwd_kernel(const TG* __restrict__ gi, const float* __restrict__ gh,
                                    const float* __restrict__ hprev32, const TA* __restrict__ hprevA,
                                    TA* __restrict__ hnextA, float* __restrict__ hnext32, TA* __restrict__ sv, int Bp,
                                    int Hp, const int* __restrict__ lens = nullptr, float* __restrict__ hlast = nullptr,
                                    int t = 0, int B = 0, int lead_pad_T = 0) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * Hp) return;
  const int b = (int)(idx / Hp), j = (int)(idx - (long long)b * Hp);
  const long long g3 = (long long)b * 3 * Hp + j;
  const float ir = to_f32<TG>(gi[g3]), iz = to_f32<TG>(gi[g3 + Hp]), in_ = to_f32<TG>(gi[g3 + 2 * Hp]);
  const float hr = gh[g3], hz = gh[g3 + Hp], hn = gh[g3 + 2 * Hp];
  const float r = sigmoid_acc(ir + hr);
  const float z = sigmoid_acc(iz + hz);
  const float n = tanhf(fmaf(r, hn, in_));
  const float hp = hprev32 ? hprev32[idx] : to_f32<TA>(hprevA[idx]);
  // reverse direction of a packed bidirectional GRU (mosesfile.py:21-28): the sequence is walked from its last token, so
  // in processing order its padding comes FIRST (lead_pad_T - lens[b] steps): the state stays at h0 there and the saved
  // gates (r=0, z=1, n=0) make the BPTT kernel emit exact zeros for these steps.
  if (lead_pad_T > 0 && b < B && t < lead_pad_T - lens[b]) {
    hnextA[idx] = from_f32<TA>(hp);
    if (hnext32) hnext32[idx] = hp;
    if (sv) {
      const long long s4 = (long long)b * 4 * Hp + j;
      sv[s4] = from_f32<TA>(0.f); sv[s4 + Hp] = from_f32<TA>(1.f); sv[s4 + 2 * Hp] = from_f32<TA>(0.f);
      sv[s4 + 3 * Hp] = from_f32<TA>(0.f);
    }
    return;
  }
  const float h = fmaf(z, hp - n, n);  // (1-z) n + z h
  hnextA[idx] = from_f32<TA>(h);
  if (hnext32) hnext32[idx] = h;
  // packed-sequence final state (mosesvae.py:153-156): sequence b ends after lens[b] steps
  if (hlast && lens && b < B && t + 1 == lens[b]) hlast[idx] = h;
  if (sv) {
    const long long s4 = (long long)b * 4 * Hp + j;
    sv[s4] = from_f32<TA>(r);
    sv[s4 + Hp] = from_f32<TA>(z);
    sv[s4 + 2 * Hp] = from_f32<TA>(n);
    sv[s4 + 3 * Hp] = from_f32<TA>(hn);
  }
}

// BPTT gate step: dh = dh_carry + dX[t];  writes dG[t] = [da_n | da_r | da_z | da_n*r] (so that the dgi window
// is columns [0,3Hp) in (n,r,z) order and the dgh window is columns [Hp,4Hp) in (r,z,n) order), then
// dh_carry <- dh*z (the W_hh term is added by the following GEMM).  Layer 0 also accumulates sum_t dgi.
template <typename TA>
__global__ void gru_gate_bwd_kernel(const TA* __restrict__ sv, const TA* __restrict__ hprevA,
                                    const TA* __restrict__ dX, float* __restrict__ dh_carry, TA* __restrict__ dG,
                                    float* __restrict__ dgi_sum /*[Bp][3Hp] rzn or null*/, int Bp, int Hp,
                                    int active_rows = 1 << 30) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * Hp) return;
  const int b = (int)(idx / Hp), j = (int)(idx - (long long)b * Hp);
  const long long s4 = (long long)b * 4 * Hp + j;
  if (b >= active_rows) {   // packed sequences: this row's sequence has ended before step t -> no gradient, nothing read
    const TA zero = from_f32<TA>(0.f);
    dG[s4] = zero; dG[s4 + Hp] = zero; dG[s4 + 2 * Hp] = zero; dG[s4 + 3 * Hp] = zero;
    return;
  }
  const float r = to_f32<TA>(sv[s4]), z = to_f32<TA>(sv[s4 + Hp]), n = to_f32<TA>(sv[s4 + 2 * Hp]),
              hn = to_f32<TA>(sv[s4 + 3 * Hp]);
  const float hp = to_f32<TA>(hprevA[idx]);
  const float dh = dh_carry[idx] + to_f32<TA>(dX[idx]);
  const float dan = dh * (1.f - z) * (1.f - n * n);
  const float daz = dh * (hp - n) * z * (1.f - z);
  const float dar = dan * hn * r * (1.f - r);
  dG[s4] = from_f32<TA>(dan);
  dG[s4 + Hp] = from_f32<TA>(dar);
  dG[s4 + 2 * Hp] = from_f32<TA>(daz);
  dG[s4 + 3 * Hp] = from_f32<TA>(dan * r);
  dh_carry[idx] = dh * z;
  if (dgi_sum) {
    const long long g3 = (long long)b * 3 * Hp + j;
    dgi_sum[g3] += dar;
    dgi_sum[g3 + Hp] += daz;
    dgi_sum[g3 + 2 * Hp] += dan;
  }
}


// ------------------------------------------------------------------------------------------
// LSTM cell (torch.nn.LSTM, gate rows i,f,g,o; models.py:128,164 name these modules "gru" but they are nn.LSTM):
//   a = gi + gh ; i,f,o = sigmoid, g = tanh ; c' = f c + i g ; h' = o tanh(c')
// One thread per (row b, unit j) of a [Bp][Hp] slab; gi (incl. both biases) is per-step [Bp][4Hp] (TG) or the
// time-invariant layer-0 projection; gh fp32 [Bp][4Hp]; c: fp32 state, updated in place.
// Saves (i,f,g,o,c_prev,tanh c') into sv[Bp][6Hp] for BPTT.
// ------------------------------------------------------------------------------------------
template <typename TA, typename TG>
__global__ void lstm_gate_fwd_kernel(const TG* __restrict__ gi, const float* __restrict__ gh, float* __restrict__ c,
                                     TA* __restrict__ hnextA, TA* __restrict__ sv, int Bp, int Hp) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * Hp) return;
  const int b = (int)(idx / Hp), j = (int)(idx - (long long)b * Hp);
  const long long g4 = (long long)b * 4 * Hp + j;
  const float ai = to_f32<TG>(gi[g4]) + gh[g4], af = to_f32<TG>(gi[g4 + Hp]) + gh[g4 + Hp];
  const float ag = to_f32<TG>(gi[g4 + 2 * Hp]) + gh[g4 + 2 * Hp], ao = to_f32<TG>(gi[g4 + 3 * Hp]) + gh[g4 + 3 * Hp];
  const float i = sigmoid_acc(ai), f = sigmoid_acc(af), g = tanhf(ag), o = sigmoid_acc(ao);
  const float cp = c[idx];
  const float cn = fmaf(f, cp, i * g);
  const float tc = tanhf(cn);
  c[idx] = cn;
  hnextA[idx] = from_f32<TA>(o * tc);
  if (sv) {
    const long long s6 = (long long)b * 6 * Hp + j;
    sv[s6] = from_f32<TA>(i); sv[s6 + Hp] = from_f32<TA>(f); sv[s6 + 2 * Hp] = from_f32<TA>(g);
    sv[s6 + 3 * Hp] = from_f32<TA>(o); sv[s6 + 4 * Hp] = from_f32<TA>(cp); sv[s6 + 5 * Hp] = from_f32<TA>(tc);
  }
}
// BPTT step: dh = dh_carry + dX[t]; writes dG[t] = [di|df|dg|do] (pre-activation grads, shared by the ih and hh paths),
// dc_carry <- dc_total * f.  dh_carry for the previous step is then dG[t] * W_hh (GEMM, not accumulated).
template <typename TA>
__global__ void lstm_gate_bwd_kernel(const TA* __restrict__ sv, const TA* __restrict__ dX,
                                     const float* __restrict__ dh_carry, float* __restrict__ dc_carry,
                                     TA* __restrict__ dG, int Bp, int Hp) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * Hp) return;
  const int b = (int)(idx / Hp), j = (int)(idx - (long long)b * Hp);
  const long long s6 = (long long)b * 6 * Hp + j;
  const float i = to_f32<TA>(sv[s6]), f = to_f32<TA>(sv[s6 + Hp]), g = to_f32<TA>(sv[s6 + 2 * Hp]),
              o = to_f32<TA>(sv[s6 + 3 * Hp]), cp = to_f32<TA>(sv[s6 + 4 * Hp]), tc = to_f32<TA>(sv[s6 + 5 * Hp]);
  const float dh = dh_carry[idx] + to_f32<TA>(dX[idx]);
  const float dct = dc_carry[idx] + dh * o * (1.f - tc * tc);
  const long long g4 = (long long)b * 4 * Hp + j;
  dG[g4] = from_f32<TA>(dct * g * i * (1.f - i));
  dG[g4 + Hp] = from_f32<TA>(dct * cp * f * (1.f - f));
  dG[g4 + 2 * Hp] = from_f32<TA>(dct * i * (1.f - g * g));
  dG[g4 + 3 * Hp] = from_f32<TA>(dh * tc * o * (1.f - o));
  dc_carry[idx] = dct * f;
}
// Eight units per thread (16-byte accesses, every load of the thread issued before the first use): the bf16 path of Config A's
// decoder (B = 4096, H = 1024: 184 MB per forward call, 136 MB per BPTT call) is HBM-bound, and with exact expf / tanhf /
// IEEE divisions the forward cell was ALU-bound on top of it (~165 lane instructions per unit).  bf16 mode only: the gates use
// ex2.approx + rcp.approx (absolute error ~2e-7, three orders of magnitude under the bf16 rounding of the saved gates and of h);
// the fp32 check mode keeps the scalar kernels above.  Needs Hp % 8 == 0 and 16-byte aligned arrays.
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  F8 r;
#pragma unroll
  for (int k = 0; k < 4; ++k) { r.v[2 * k] = __uint_as_float(w[k] << 16); r.v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
  return r;
}
__device__ __forceinline__ F8 ld8(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  return F8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& f) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f.v[2 * k], f.v[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void st8(float* p, const F8& f) {
  *reinterpret_cast<float4*>(p) = make_float4(f.v[0], f.v[1], f.v[2], f.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f.v[4], f.v[5], f.v[6], f.v[7]);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, sigmoid_fast(2.0f * x), -1.0f); }

template <typename TG, bool FAST>
__global__ void __launch_bounds__(256) lstm_gate_fwd_x8_kernel(const TG* __restrict__ gi, const float* __restrict__ gh,
                                                               float* __restrict__ c, __nv_bfloat16* __restrict__ hnextA,
                                                               __nv_bfloat16* __restrict__ sv, int Bp, int Hp) {
  const int H8 = Hp >> 3;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * H8) return;
  const int b = (int)(idx / H8), j = (int)(idx - (long long)b * H8) * 8;
  const long long g4 = (long long)b * 4 * Hp + j, e = (long long)b * Hp + j;
  const F8 xi = ld8(gi + g4), xf = ld8(gi + g4 + Hp), xg = ld8(gi + g4 + 2 * Hp), xo = ld8(gi + g4 + 3 * Hp);
  const F8 hi = ld8(gh + g4), hf = ld8(gh + g4 + Hp), hg = ld8(gh + g4 + 2 * Hp), ho = ld8(gh + g4 + 3 * Hp);
  const F8 cp = ld8(c + e);
  F8 i, f, g, o, cn, tc, h;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if constexpr (FAST) {
      i.v[k] = sigmoid_fast(xi.v[k] + hi.v[k]);
      f.v[k] = sigmoid_fast(xf.v[k] + hf.v[k]);
      g.v[k] = tanh_fast(xg.v[k] + hg.v[k]);
      o.v[k] = sigmoid_fast(xo.v[k] + ho.v[k]);
    } else {   // bit-identical to lstm_gate_fwd_kernel
      i.v[k] = sigmoid_acc(xi.v[k] + hi.v[k]);
      f.v[k] = sigmoid_acc(xf.v[k] + hf.v[k]);
      g.v[k] = tanhf(xg.v[k] + hg.v[k]);
      o.v[k] = sigmoid_acc(xo.v[k] + ho.v[k]);
    }
    cn.v[k] = fmaf(f.v[k], cp.v[k], i.v[k] * g.v[k]);
    tc.v[k] = FAST ? tanh_fast(cn.v[k]) : tanhf(cn.v[k]);
    h.v[k] = o.v[k] * tc.v[k];
  }
  st8(c + e, cn);
  st8(hnextA + e, h);
  if (sv) {
    const long long s6 = (long long)b * 6 * Hp + j;
    st8(sv + s6, i); st8(sv + s6 + Hp, f); st8(sv + s6 + 2 * Hp, g);
    st8(sv + s6 + 3 * Hp, o); st8(sv + s6 + 4 * Hp, cp); st8(sv + s6 + 5 * Hp, tc);
  }
}
__global__ void __launch_bounds__(256) lstm_gate_bwd_x8_kernel(const __nv_bfloat16* __restrict__ sv, const __nv_bfloat16* __restrict__ dX,
                                                               const float* __restrict__ dh_carry, float* __restrict__ dc_carry,
                                                               __nv_bfloat16* __restrict__ dG, int Bp, int Hp) {
  const int H8 = Hp >> 3;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * H8) return;
  const int b = (int)(idx / H8), j = (int)(idx - (long long)b * H8) * 8;
  const long long s6 = (long long)b * 6 * Hp + j, g4 = (long long)b * 4 * Hp + j, e = (long long)b * Hp + j;
  const F8 i = ld8(sv + s6), f = ld8(sv + s6 + Hp), g = ld8(sv + s6 + 2 * Hp), o = ld8(sv + s6 + 3 * Hp),
           cp = ld8(sv + s6 + 4 * Hp), tc = ld8(sv + s6 + 5 * Hp);
  const F8 dx = ld8(dX + e), dhc = ld8(dh_carry + e), dcc = ld8(dc_carry + e);
  F8 di, df, dg, dO, dc;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float dh = dhc.v[k] + dx.v[k];
    const float dct = dcc.v[k] + dh * o.v[k] * (1.f - tc.v[k] * tc.v[k]);
    di.v[k] = dct * g.v[k] * i.v[k] * (1.f - i.v[k]);
    df.v[k] = dct * cp.v[k] * f.v[k] * (1.f - f.v[k]);
    dg.v[k] = dct * i.v[k] * (1.f - g.v[k] * g.v[k]);
    dO.v[k] = dh * tc.v[k] * o.v[k] * (1.f - o.v[k]);
    dc.v[k] = dct * f.v[k];
  }
  st8(dG + g4, di); st8(dG + g4 + Hp, df); st8(dG + g4 + 2 * Hp, dg); st8(dG + g4 + 3 * Hp, dO);
  st8(dc_carry + e, dc);
}
// 4-gate padding helpers: dst[4Hp][cols_p] <- src[4H][cols] ; inverse ; bias vectors
template <typename T>
__global__ void pad_gates4_kernel(const float* __restrict__ src, int H, int cols, T* __restrict__ dst, int Hp, int cols_p) {
  const long long total = 4ll * Hp * cols_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols_p);
    const int r = (int)(i / cols_p);
    const int g = r / Hp, j = r - g * Hp;
    dst[i] = from_f32<T>((j < H && c < cols) ? src[((long long)g * H + j) * cols + c] : 0.f);
  }
}
__global__ void unpad_gates4_kernel(const float* __restrict__ src, int Hp, int cols_p, float* __restrict__ dst, int H, int cols) {
  const long long total = 4ll * H * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int r = (int)(i / cols);
    const int g = r / H, j = r - g * H;
    dst[i] = src[((long long)g * Hp + j) * cols_p + c];
  }
}
// padded combined bias: dst[4Hp] = b_ih + b_hh
__global__ void pad_bias4_sum_kernel(const float* __restrict__ bih, const float* __restrict__ bhh, int H, float* __restrict__ dst, int Hp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * Hp) return;
  const int g = i / Hp, j = i - g * Hp;
  dst[i] = j < H ? bih[g * H + j] + bhh[g * H + j] : 0.f;
}
// dst_a[4H] = dst_b[4H] = src[4Hp] (unpadded): the two LSTM biases receive the same gradient
__global__ void unpad_bias4_dup_kernel(const float* __restrict__ src, int Hp, float* __restrict__ a, float* __restrict__ b, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * H) return;
  const float v = src[(i / H) * Hp + (i % H)];
  a[i] = v; b[i] = v;
}

// ------------------------------------------------------------------------------------------
// Head: softmax over the charset + max_len * BCE(mean) (train.py:31-35) and its gradient wrt the logits
// (SURVEY.md A.3).  One warp per (t,b) row of logits[T*Bp][CP]; lanes cover columns c, c+32.
// Rows are time-major (row = t*Bp + b).  Also counts per-molecule argmax hits (train.py:110-112).
// ------------------------------------------------------------------------------------------
template <typename TA>
__global__ void head_softmax_bce_kernel(const float* __restrict__ logits, int CP, int C, const uint8_t* __restrict__ ids,
                                        int B, int Bp, int T, float gscale /* max_len/(B*T*C) */,
                                        TA* __restrict__ dlogits, float* __restrict__ probs /*[B][T][C] or null*/,
                                        double* __restrict__ bce_sum, int* __restrict__ hit_count /*[B] or null*/) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)T * Bp;
  double local = 0.0;
  if (warp < rows) {
    const int t = warp / Bp, b = warp - t * Bp;
    const long long row = warp;
    if (b >= B) {
      if (dlogits) for (int c = lane; c < CP; c += 32) dlogits[row * CP + c] = from_f32<TA>(0.f);
    } else {
      const int y = ids[(long long)b * T + t];
      float a0 = lane < C ? logits[row * CP + lane] : -INFINITY;
      float a1 = (lane + 32) < C ? logits[row * CP + lane + 32] : -INFINITY;
      float m = fmaxf(a0, a1);
      int arg = a0 >= a1 ? lane : lane + 32;
      for (int o = 16; o; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
      }
      const float e0 = lane < C ? expf(a0 - m) : 0.f;
      const float e1 = (lane + 32) < C ? expf(a1 - m) : 0.f;
      float s = e0 + e1;
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float p0 = e0 / s, p1 = e1 / s;
      float g0 = 0.f, g1 = 0.f, l = 0.f;
      if (lane < C) {
        const float x = (lane == y) ? 1.f : 0.f;
        l += x > 0.f ? -fmaxf(logf(p0), -100.f) : -fmaxf(log1pf(-p0), -100.f);
        g0 = gscale * (p0 - x) / fmaxf(p0 * (1.f - p0), 1e-12f);
      }
      if (lane + 32 < C) {
        const float x = (lane + 32 == y) ? 1.f : 0.f;
        l += x > 0.f ? -fmaxf(logf(p1), -100.f) : -fmaxf(log1pf(-p1), -100.f);
        g1 = gscale * (p1 - x) / fmaxf(p1 * (1.f - p1), 1e-12f);
      }
      float gp = g0 * p0 + g1 * p1;
      for (int o = 16; o; o >>= 1) {
        gp += __shfl_xor_sync(0xffffffffu, gp, o);
        l += __shfl_xor_sync(0xffffffffu, l, o);
      }
      if (dlogits) {
        dlogits[row * CP + lane] = from_f32<TA>(lane < C ? p0 * (g0 - gp) : 0.f);
        if (lane + 32 < CP) dlogits[row * CP + lane + 32] = from_f32<TA>((lane + 32) < C ? p1 * (g1 - gp) : 0.f);
      }
      if (probs) {
        float* pr = probs + ((long long)b * T + t) * C;
        if (lane < C) pr[lane] = p0;
        if (lane + 32 < C) pr[lane + 32] = p1;
      }
      if (lane == 0) {
        local = (double)l;
        if (hit_count && arg == y) atomicAdd(hit_count + b, 1);
      }
    }
  }
  // block reduce of the loss partials
  __shared__ double red[32];
  if (lane == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0 && bce_sum) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(bce_sum, s);
  }
}

// dlogits from an EXTERNAL gradient wrt the probabilities (drop-in autograd path):
//   dlogit_j = p_j * (dp_j - sum_c dp_c p_c)      probs/dprobs are [B][T][C] fp32
template <typename TA>
__global__ void head_softmax_bwd_kernel(const float* __restrict__ logits, int CP, int C, const float* __restrict__ dprobs,
                                        int B, int Bp, int T, TA* __restrict__ dlogits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)T * Bp) return;
  const int t = warp / Bp, b = warp - t * Bp;
  const long long row = warp;
  if (b >= B) {
    for (int c = lane; c < CP; c += 32) dlogits[row * CP + c] = from_f32<TA>(0.f);
    return;
  }
  float a0 = lane < C ? logits[row * CP + lane] : -INFINITY;
  float a1 = (lane + 32) < C ? logits[row * CP + lane + 32] : -INFINITY;
  float m = fmaxf(a0, a1);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e0 = lane < C ? expf(a0 - m) : 0.f, e1 = (lane + 32) < C ? expf(a1 - m) : 0.f;
  float s = e0 + e1;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float p0 = e0 / s, p1 = e1 / s;
  const float* dp = dprobs + ((long long)b * T + t) * C;
  const float d0 = lane < C ? dp[lane] : 0.f, d1 = (lane + 32) < C ? dp[lane + 32] : 0.f;
  float gp = d0 * p0 + d1 * p1;
  for (int o = 16; o; o >>= 1) gp += __shfl_xor_sync(0xffffffffu, gp, o);
  dlogits[row * CP + lane] = from_f32<TA>(lane < C ? p0 * (d0 - gp) : 0.f);
  if (lane + 32 < CP) dlogits[row * CP + lane + 32] = from_f32<TA>((lane + 32) < C ? p1 * (d1 - gp) : 0.f);
}

// greedy decode: argmax over the charset of each logits row -> ids[b][t] (ties -> lowest index)
__global__ void head_argmax_kernel(const float* __restrict__ logits, int CP, int C, int B, int Bp, int T,
                                   uint8_t* __restrict__ out_ids) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)T * Bp) return;
  const int t = warp / Bp, b = warp - t * Bp;
  if (b >= B) return;
  const long long row = warp;
  float a0 = lane < C ? logits[row * CP + lane] : -INFINITY;
  float a1 = (lane + 32) < C ? logits[row * CP + lane + 32] : -INFINITY;
  float m = fmaxf(a0, a1);
  int arg = a0 >= a1 ? lane : lane + 32;
  for (int o = 16; o; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
  }
  if (lane == 0) out_ids[(long long)b * T + t] = (uint8_t)arg;
}

// out[0]=loss, [1]=max_len*bce_mean, [2]=kl, [3]=#molecules reconstructed exactly
__global__ void finalize_scalars_kernel(const double* __restrict__ bce_sum, const double* __restrict__ kl_sum,
                                        const int* __restrict__ hit_count, int B, int T, double bce_scale,
                                        double kl_scale, float* __restrict__ out, const int* __restrict__ err_flag = nullptr) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int local = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) local += (hit_count[b] == T) ? 1 : 0;
  atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0) {
    const double bce = bce_sum[0] * bce_scale;
    const double kl = -0.5 * kl_sum[0] * kl_scale;
    // a fired pipeline watchdog (bounded waits of the tcgen05 / persistent kernels) poisons the returned loss, so a caller
    // that never polls the device error flag still cannot train on garbage silently
    const bool bad = err_flag && *err_flag != 0;
    out[0] = bad ? __int_as_float(0x7fc00000) : (float)(bce + kl);
    out[1] = (float)bce;
    out[2] = (float)kl;
    out[3] = (float)cnt;
  }
}
// same guard for step functions whose scalars are final before their last kernels ran
__global__ void nan_if_error_kernel(const int* __restrict__ err_flag, float* __restrict__ out) {
  if (*err_flag != 0) out[0] = __int_as_float(0x7fc00000);
}

}  // namespace
}  // namespace simt
