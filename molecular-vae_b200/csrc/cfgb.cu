// Orchestration + C ABI of the Config-B ELBO step (models2d.py:8-52 + train.py:31-38) on one B200.
//
// HBM layout (all recurrent tensors are TIME-MAJOR so that step t of a layer is one contiguous
// [Bp][Hp] matrix = one TMA slab; Bp = B rounded up to 128 rows, Hp = H rounded up to 64 columns,
// pad rows/columns are exact zeros and stay zero through the recurrence):
//   hs[l]   [(T+1)][Bp][Hp]   TA   hidden states, slab 0 = h0 = 0; slabs 1..T feed the next layer / head
//   sv[l]   [T][Bp][4Hp]      TA   saved (r, z, n, W_hn h + b_hn) for BPTT
//   gi_all  [T][Bp][3Hp]      TA   x_t W_ih^T + b_ih of the current layer (l >= 1); layer 0's projection
//                                  is time-invariant (the Repeat(T) of models2d.py:42 is never built)
//   dG      [T][Bp][4Hp]      TA   [da_n | da_r | da_z | da_n*r]: columns [0,3Hp) = dgi in (n,r,z) order,
//                                  columns [Hp,4Hp) = dgh in (r,z,n) order -> both GEMM windows contiguous
//   dX      [T][Bp][Hp]       TA   gradient flowing into a layer's outputs (from the head or the layer above)
// TA = float in MVAE_PREC_FP32 (CUDA-core SGEMM everywhere) and bf16 in MVAE_PREC_BF16 (tcgen05 GEMMs,
// fp32 accumulation, fp32 master copy of h across steps).
#include <new>
#include <stdio.h>
#include <string.h>

#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "umma_gemm.h"
#include "gru_rec.h"
#include <stdlib.h>

namespace {

long long g_launches = 0;
inline void count(int n = 1) { g_launches += n; }

// ---- optional per-kernel timing (bench.py roofline): CUDA events recorded on the launching stream around the
// persistent recurrence kernels of DIRECT (non-captured) launches.  Off by default; nothing is recorded while a
// stream is being captured.
constexpr int PROF_TAGS = 4;
struct ProfPair { cudaEvent_t e0, e1; int tag; };
bool g_prof_on = false;
ProfPair g_prof[4096];
int g_prof_n = 0;
struct ProfScope {
  cudaStream_t st; int idx;
  ProfScope(int tag, cudaStream_t s) : st(s), idx(-1) {
    if (!g_prof_on || g_prof_n >= 4096) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    ProfPair& p = g_prof[g_prof_n];
    if (cudaEventCreate(&p.e0) != cudaSuccess) return;
    if (cudaEventCreate(&p.e1) != cudaSuccess) { cudaEventDestroy(p.e0); return; }
    p.tag = tag;
    idx = g_prof_n++;
    cudaEventRecord(p.e0, st);
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(g_prof[idx].e1, st); }
};

struct Dims {
  int B, Bp, T, C, CP, Z, H, Hp, L, F0, FLAT, L1, L2, L3;
  bool bf16, train;
  float max_len, eps_scale;
};

int make_dims(const mvae_cfgb_desc* d, Dims* o) {
  if (!d) return MVAE_ERR_INVALID;
  if (d->batch <= 0 || d->seq_len <= 0 || d->seq_len > 255 || d->charset < 27 || d->charset > 64 || d->latent <= 0 ||
      d->hidden <= 0 || d->layers < 1 || d->layers > 4 || d->fc0 <= 0)
    return MVAE_ERR_INVALID;
  if (d->precision != MVAE_PREC_FP32 && d->precision != MVAE_PREC_BF16) return MVAE_ERR_INVALID;
  o->B = d->batch; o->Bp = round_up(d->batch, 256);
  o->T = d->seq_len; o->C = d->charset; o->CP = 64;
  o->Z = d->latent; o->H = d->hidden; o->Hp = round_up(d->hidden, 64); o->L = d->layers; o->F0 = d->fc0;
  o->L1 = o->C - 8; o->L2 = o->L1 - 8; o->L3 = o->L2 - 10;
  if (o->L3 <= 0) return MVAE_ERR_INVALID;
  o->FLAT = 10 * o->L3;
  o->bf16 = d->precision == MVAE_PREC_BF16;
  o->train = d->train != 0;
  o->max_len = d->max_len; o->eps_scale = d->eps_scale;
  return MVAE_OK;
}

// ---- workspace carve-up ---------------------------------------------------------------------
struct WS {
  // control
  int* err_flag; int* bad_input; double* bce_sum; double* kl_sum; int* hit_count;
  // encoder / latent (fp32)
  float *h1, *h2, *h3, *h4, *mu, *lv, *z, *zr, *dzr, *da5, *dz, *dmu, *dlv, *dh4, *dflat;
  float* gi0;      // [Bp][3Hp]
  float* dgi0sum;  // [Bp][3Hp]
  // recurrent
  void* hs[4]; void* sv[4];
  void* gi_all; void* dG; void* dX;
  float* gh;        // [Bp][3Hp]
  float* h32[2];    // [Bp][Hp] fp32 master h (bf16 mode)
  float* dh_carry;  // [Bp][Hp]
  // head
  float* logits;    // [T*Bp][CP]
  void* dlogits;    // [T*Bp][CP] TA
  // padded weights (TA) and biases (fp32)
  void* Whh_p[4]; void* Wih_p[4]; void* Wih_nrz[4]; void* W3_p;
  void* WhhT_p[4];          // bf16 [Hp][3Hp] (fused BPTT kernel operand)
  void* gi0_bf;             // bf16 [Bp][3Hp] copy of the layer-0 projection (fused forward kernel)
  // tensor-core layer-0 path (fused recurrence only): bf16 operands of gi0 = zr W_ih0^T and of its two gradient GEMMs
  void* zr_bf;              // [Bp][Zp]
  void* Wih0_p;             // [3Hp][Zp] gate-padded W_ih_l0
  void* dgi0sum_bf;         // [Bp][3Hp] (r,z,n) time-summed dgi of layer 0
  float* dWih0_p;           // [3Hp][Zp] staging
  unsigned int* counters;   // [Bp/128] inter-CTA step counters of the fused recurrence
  float* bih_p[4]; float* bhh_p[4]; float* b3_p;
  float* bcomb_p[4];        // b_ih + (b_hr, b_hz, 0): projection bias when the fused kernel only adds b_hn
  // padded gradient staging (fp32)
  float* dW_p;   // [3Hp][Hp]
  float* dW3_p;  // [CP][Hp]
  float* csum;   // [4Hp]
  void* tcs; size_t tcs_bytes;   // bf16 mode: converted operands of the bf16x3 tensor-core path of the small fp32 GEMMs
  size_t total;
};

struct Carver {
  uint8_t* base; size_t off;
  template <typename T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void carve(const Dims& d, void* base, WS* w) {
  Carver c{reinterpret_cast<uint8_t*>(base), 0};
  const size_t es = d.bf16 ? 2 : 4;
  const size_t B = d.B, Bp = d.Bp, T = d.T, Hp = d.Hp;
  w->err_flag = c.take<int>(1);
  w->bad_input = c.take<int>(1);
  w->bce_sum = c.take<double>(1);
  w->kl_sum = c.take<double>(1);
  w->hit_count = c.take<int>(B);
  w->h1 = c.take<float>(B * 9 * d.L1);
  w->h2 = c.take<float>(B * 9 * d.L2);
  w->h3 = c.take<float>(B * d.FLAT);
  w->h4 = c.take<float>(B * d.F0);
  w->mu = c.take<float>(B * d.Z);
  w->lv = c.take<float>(B * d.Z);
  w->z = c.take<float>(B * d.Z);
  w->zr = c.take<float>(B * d.Z);
  w->dzr = c.take<float>(B * d.Z);
  w->da5 = c.take<float>(B * d.Z);
  w->dz = c.take<float>(B * d.Z);
  w->dmu = c.take<float>(B * d.Z);
  w->dlv = c.take<float>(B * d.Z);
  w->dh4 = c.take<float>(B * d.F0);
  w->dflat = c.take<float>(B * d.FLAT);
  w->gi0 = c.take<float>(Bp * 3 * Hp);
  w->dgi0sum = c.take<float>(Bp * 3 * Hp);
  for (int l = 0; l < d.L; ++l) {
    w->hs[l] = c.take<uint8_t>((T + 1) * Bp * Hp * es);
    w->sv[l] = c.take<uint8_t>(T * Bp * 5 * Hp * es);   // 5th block: h_{t-1} copy (fused kernel, fragment layout)
  }
  w->gi_all = c.take<uint8_t>(T * Bp * 3 * Hp * es);
  w->dG = c.take<uint8_t>(T * Bp * 4 * Hp * es);
  w->dX = c.take<uint8_t>(T * Bp * Hp * es);
  w->gh = c.take<float>(Bp * 3 * Hp);
  w->h32[0] = c.take<float>(Bp * Hp);
  w->h32[1] = c.take<float>(Bp * Hp);
  w->dh_carry = c.take<float>(Bp * Hp);
  w->logits = c.take<float>(T * Bp * d.CP);
  w->dlogits = c.take<uint8_t>(T * Bp * d.CP * es);
  for (int l = 0; l < d.L; ++l) {
    w->Whh_p[l] = c.take<uint8_t>(3 * Hp * Hp * es);
    w->Wih_p[l] = c.take<uint8_t>(3 * Hp * Hp * es);
    w->Wih_nrz[l] = c.take<uint8_t>(3 * Hp * Hp * es);
    w->WhhT_p[l] = c.take<uint8_t>(3 * Hp * Hp * es);
    w->bih_p[l] = c.take<float>(3 * Hp);
    w->bhh_p[l] = c.take<float>(3 * Hp);
    w->bcomb_p[l] = c.take<float>(3 * Hp);
  }
  w->W3_p = c.take<uint8_t>(d.CP * Hp * es);
  w->b3_p = c.take<float>(d.CP);
  w->dW_p = c.take<float>(3 * Hp * Hp);
  w->dW3_p = c.take<float>(d.CP * Hp);
  w->csum = c.take<float>(4 * Hp);
  w->gi0_bf = c.take<uint8_t>(Bp * 3 * Hp * 2);
  {
    const size_t Zp = (size_t)round_up(d.Z, 8);
    w->zr_bf = c.take<uint8_t>(Bp * Zp * 2);
    w->Wih0_p = c.take<uint8_t>(3 * Hp * Zp * 2);
    w->dgi0sum_bf = c.take<uint8_t>(Bp * 3 * Hp * 2);
    w->dWih0_p = c.take<float>(3 * Hp * Zp);
  }
  w->counters = c.take<unsigned int>(Bp / 128 + 1);
  w->tcs_bytes = d.bf16 ? mvae_tc_sgemm_scratch_bytes((long long)max((size_t)Bp, 3 * Hp), (long long)max(max(d.F0, d.Z), max(d.FLAT, (int)Hp))) : 0;
  w->tcs = c.take<uint8_t>(w->tcs_bytes);
  w->total = (c.off + 255) & ~size_t(255);
}

// parameter indices (state_dict order)
enum { P_C1W = 0, P_C1B, P_C2W, P_C2B, P_C3W, P_C3B, P_FC0W, P_FC0B, P_FC11W, P_FC11B, P_FC12W, P_FC12B, P_FC2W, P_FC2B, P_GRU0 };
inline int P_WIH(int l) { return P_GRU0 + 4 * l; }
inline int P_WHH(int l) { return P_GRU0 + 4 * l + 1; }
inline int P_BIH(int l) { return P_GRU0 + 4 * l + 2; }
inline int P_BHH(int l) { return P_GRU0 + 4 * l + 3; }
inline int P_FC3W(int L) { return P_GRU0 + 4 * L; }
inline int P_FC3B(int L) { return P_GRU0 + 4 * L + 1; }

#define RC(expr) do { int _rc = (expr); if (_rc != MVAE_OK) return _rc; } while (0)
#define KCHECK() do { count(); MVAE_CUDA_CHECK(cudaGetLastError()); } while (0)

inline int grid_for(long long n, int block = 256, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
inline int memset_async(void* p, size_t bytes, cudaStream_t st) {
  count();
  MVAE_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  return MVAE_OK;
}

// ---- the GEMM dispatcher --------------------------------------------------------------------
// out[M][N] (+)= A * B (+bias).  a_trans: A stored [K][M] (else [M][K]); b_kmajor: B stored [N][K] (else [K][N]).
template <typename TA>
int gemm(const Dims& d, const WS& w, cudaStream_t st, const TA* A, long long lda, bool a_trans, const TA* B,
         long long ldb, bool b_kmajor, void* out, long long ldc, bool out_is_ta, int M, int N, int K, const float* bias,
         bool accumulate, int splits, int bn = 0, bool out_rb = false) {
  count();
  if constexpr (sizeof(TA) == 4) {
    (void)out_is_ta; (void)bn; (void)d; (void)w; (void)out_rb;
    return simt::sgemm(st, reinterpret_cast<const float*>(A), a_trans ? 1 : lda, a_trans ? lda : 1,
                       reinterpret_cast<const float*>(B), b_kmajor ? 1 : ldb, b_kmajor ? ldb : 1,
                       reinterpret_cast<float*>(out), ldc, M, N, K, bias, simt::ACT_NONE, accumulate ? 1 : 0, splits);
  } else {
    (void)d;
    mvae_umma_operand a{A, a_trans ? 1 : 0, M, K, lda, 1, 0, 0, 0};
    mvae_umma_operand b{B, b_kmajor ? 0 : 1, N, K, ldb, 1, 0, 0, 0};
    mvae_umma_out o{out, ldc, out_is_ta ? 1 : 0, accumulate ? 1 : 0, bias, out_rb ? 1 : 0};
    return mvae_umma_gemm(&a, &b, &o, M, N, K, bn, splits, 0, w.err_flag, st);
  }
}

// bf16 mode: the small fp32 GEMMs of the latent / encoder Linears go through the tensor cores (bf16x3 split, fp32-class
// accuracy, umma_gemm.h); set per call by check_ws.  MVAE_TC_SGEMM=0 keeps them on the CUDA cores.
thread_local mvae_tc_ctx g_tc{nullptr, 0, nullptr};
bool tc_sgemm_enabled() {
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
    if (rc != MVAE_ERR_UNSUPPORTED) { count(n); return rc; }
  }
  count();
  return simt::sgemm(st, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act, accumulate, splits);
}
// dW[M][N] = A^T B with the contraction over the batch rows: zero + split-K
inline int sg_wgrad(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk,
                    long long sbn, float* C, long long ldc, int M, int N, int K) {
  // C rows are contiguous blocks of N only when ldc == N; zero row by row otherwise
  if (ldc == N) RC(memset_async(C, (size_t)M * N * 4, st));
  else {
    count();
    MVAE_CUDA_CHECK(cudaMemset2DAsync(C, ldc * 4, 0, (size_t)N * 4, M, st));
  }
  int splits = K >= 1024 ? 16 : (K >= 256 ? 4 : 1);
  return sg(st, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, nullptr, simt::ACT_NONE, 1, splits);
}

__global__ void gate_bias_grads_kernel(const float* __restrict__ csum, int H, int Hp, float* __restrict__ db_ih,
                                       float* __restrict__ db_hh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * H) return;
  const int g = i / H, j = i - g * H;
  // csum blocks: 0 = da_n, 1 = da_r, 2 = da_z, 3 = da_n*r
  const int blk_ih = (g == 0) ? 1 : (g == 1 ? 2 : 0);
  const int blk_hh = (g == 0) ? 1 : (g == 1 ? 2 : 3);
  db_ih[i] = csum[blk_ih * Hp + j];
  db_hh[i] = csum[blk_hh * Hp + j];
}
__global__ void copy_prefix_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}


// W_hh^T padded: dst[j][g*Hp + i] = W_hh[g*H + i][j]   (K-major operand of the fused BPTT kernel)
__global__ void pad_gate_matrix_T_kernel(const float* __restrict__ src, int H, __nv_bfloat16* __restrict__ dst, int Hp) {
  const long long total = 3ll * Hp * Hp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % (3 * Hp));
    const int j = (int)(idx / (3 * Hp));
    const int g = k / Hp, i = k - g * Hp;
    float v = 0.f;
    if (i < H && j < H) v = src[((long long)g * H + i) * H + j];
    dst[idx] = __float2bfloat16_rn(v);
  }
}
// layer 0: sum over time of the dgi window of dG ([T][Bp][4Hp], blocks n,r,z) -> fp32 [Bp][3Hp] in (r,z,n) order
__global__ void dgi_time_sum_kernel(const __nv_bfloat16* __restrict__ dG, int T, int Bp, int Hp,
                                    float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf = nullptr) {
  // one thread per (row b, 8 consecutive columns of the dgi window): 16-byte loads, fp32 accumulation over t
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cpr = 3 * Hp / 8;
  if (idx >= (long long)Bp * cpr) return;
  const int c8 = (int)(idx % cpr) * 8;          // column inside the (n,r,z) dgi window of dG
  const int b = (int)(idx / cpr);
  const int blk = c8 / Hp, j = c8 - blk * Hp;   // blk: 0 = n, 1 = r, 2 = z
  const int g = (blk == 0) ? 2 : (blk - 1);     // -> (r,z,n) order of the output
  const uint4* p = reinterpret_cast<const uint4*>(dG + (long long)b * 4 * Hp + c8);
  const long long tstride = (long long)Bp * 4 * Hp / 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = 0; t < T; ++t) {
    const uint4 v = __ldg(p + (long long)t * tstride);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[2 * k] += __uint_as_float(w[k] << 16);
      s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
    }
  }
  float* o = out + (long long)b * 3 * Hp + g * Hp + j;
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = s[k];
  if (out_bf) {
    __nv_bfloat16* ob = out_bf + (long long)b * 3 * Hp + g * Hp + j;
#pragma unroll
    for (int k = 0; k < 8; ++k) ob[k] = __float2bfloat16_rn(s[k]);
  }
}

// ones-column helpers (fused path): column Hp-1 of every hidden-state slab is the constant 1
__global__ void set_column_kernel(__nv_bfloat16* __restrict__ x, long long rows, int ld, int col, float v) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r < rows) x[r * ld + col] = __float2bfloat16_rn(v);
}
// bias grads read from the ones column of the padded weight gradients: src [3Hp][Hp] blocks named by (g0,g1,g2)
__global__ void gate_bias_from_dw_kernel(const float* __restrict__ dW_p, int H, int Hp, int g0, int g1, int g2,
                                         float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * H) return;
  const int g = i / H, j = i - g * H;
  const int blk = (g == g0) ? 0 : ((g == g1) ? 1 : 2);
  db[i] = dW_p[((long long)blk * Hp + j) * Hp + (Hp - 1)];
}
__global__ void unpad_gate_vector_kernel(const float* __restrict__ src, int Hp, float* __restrict__ dst, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * H) dst[i] = src[(i / H) * Hp + (i % H)];
}
__global__ void strided_copy_kernel(const float* __restrict__ src, long long stride, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[(long long)i * stride];
}

int g_sm_count = 0;
// 0: per-step GEMM + gate kernels; 1/2: persistent fused recurrence variants (gru_rec.cu)
int rec_variant(const Dims& d) {
  if (!d.bf16) return 0;
  const char* e = getenv("MVAE_REC");
  int v = e ? atoi(e) : 32;   // 32: pair kernel forward + K-split BPTT (gru_rec2.cu); 3: pair kernel both ways
  if (v <= 0) return 0;
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  }
  if (v >= 3) {  // 3: pair clusters, 34 / 38: clusters of 4 / 8 with operand multicast
    if (d.Hp == 256 || d.Hp == 512) return v;
    v = 1;
  }
  if (d.Bp > mvae_gru_rec_max_rows(d.Hp, v, g_sm_count)) return 0;
  return v;
}
// the fused path keeps a constant-1 pad column in every hidden-state slab (needs H < Hp)
bool ones_column(const Dims& d) { return rec_variant(d) >= 3 && d.H < d.Hp; }
// MVAE_FUSED_HEAD=0 falls back to logits in HBM + the separate softmax/BCE kernel (the parity cross-check)
bool fuse_head_enabled() {
  const char* e = getenv("MVAE_FUSED_HEAD");
  return e ? atoi(e) != 0 : true;
}
// gate non-linearities of the fused forward sweep through tanh.approx (one MUFU op per gate, |err| ~ 5e-4, below the
// bf16 rounding of the saved gates); MVAE_FAST_GATES=0 selects the exp/rcp forms.  fp32 check mode never uses it.
int fast_gates() {
  const char* e = getenv("MVAE_FAST_GATES");
  return e ? atoi(e) : 1;
}
__global__ void combine_bias_kernel(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ out,
                                    int Hp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * Hp) out[i] = bih[i] + (i < 2 * Hp ? bhh[i] : 0.f);
}
// layer-0 projection -> bf16, optionally adding the recurrent r/z biases (fused kernel variant 3)
// rb: write the row-blocked layout [row/32][3Hp/8][32][8] the fused kernel's epilogue reads
__global__ void gi0_to_bf16_kernel(const float* __restrict__ src, const float* __restrict__ bhh_rz, int Hp,
                                   __nv_bfloat16* __restrict__ dst, long long n, int rb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (3 * Hp));
    const long long r = i / (3 * Hp);
    const long long o = rb ? ((r >> 5) * (3 * Hp / 8) + (c >> 3)) * 256 + (r & 31) * 8 + (c & 7) : i;
    dst[o] = __float2bfloat16_rn(src[i] + ((bhh_rz && c < 2 * Hp) ? bhh_rz[c] : 0.f));
  }
}

// ---- weight preparation ---------------------------------------------------------------------
// bf16 fast path: every padded / permuted / transposed weight copy of the step in ONE launch (a table of up to 32 jobs in the
// kernel parameters, blockIdx.y = job) instead of ~22 launches of a few microseconds each.
struct PrepJob {
  const float* src; const float* src2; void* dst;
  int type;            // 0 gate matrix -> bf16, 1 W_hh^T -> bf16, 2 gate vector, 3 b_ih + (b_hr, b_hz, 0), 4 matrix -> bf16, 5 matrix fp32
  int H, cols, Hp, cols_p, g0, g1, g2;
  long long total;
};
struct PrepJobs { PrepJob j[32]; int n; };
__global__ void prep_weights_fused_kernel(const __grid_constant__ PrepJobs jobs) {
  const PrepJob& jb = jobs.j[blockIdx.y];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < jb.total; i += (long long)gridDim.x * blockDim.x) {
    if (jb.type == 0) {          // src [3H][cols] -> dst [3Hp][cols_p], block order (g0,g1,g2) names the source gate
      const int c = (int)(i % jb.cols_p), r = (int)(i / jb.cols_p);
      const int blk = r / jb.Hp, j = r - blk * jb.Hp;
      const int g = blk == 0 ? jb.g0 : (blk == 1 ? jb.g1 : jb.g2);
      float v = 0.f;
      if (j < jb.H && c < jb.cols) v = jb.src[((long long)g * jb.H + j) * jb.cols + c];
      reinterpret_cast<__nv_bfloat16*>(jb.dst)[i] = __float2bfloat16_rn(v);
    } else if (jb.type == 1) {   // dst[j][g*Hp + i] = W_hh[g*H + i][j]
      const int k = (int)(i % (3 * jb.Hp)), j = (int)(i / (3 * jb.Hp));
      const int g = k / jb.Hp, ii = k - g * jb.Hp;
      float v = 0.f;
      if (ii < jb.H && j < jb.H) v = jb.src[((long long)g * jb.H + ii) * jb.H + j];
      reinterpret_cast<__nv_bfloat16*>(jb.dst)[i] = __float2bfloat16_rn(v);
    } else if (jb.type == 2) {   // padded gate vector
      const int g = (int)(i / jb.Hp), j = (int)(i - (long long)g * jb.Hp);
      reinterpret_cast<float*>(jb.dst)[i] = j < jb.H ? jb.src[g * jb.H + j] : 0.f;
    } else if (jb.type == 3) {   // b_ih + (b_hr, b_hz, 0), padded
      const int g = (int)(i / jb.Hp), j = (int)(i - (long long)g * jb.Hp);
      reinterpret_cast<float*>(jb.dst)[i] = j < jb.H ? jb.src[g * jb.H + j] + (g < 2 ? jb.src2[g * jb.H + j] : 0.f) : 0.f;
    } else {                     // src [H][cols] -> dst [Hp][cols_p] (H / Hp = rows here)
      const int c = (int)(i % jb.cols_p), r = (int)(i / jb.cols_p);
      const float v = (r < jb.H && c < jb.cols) ? jb.src[(long long)r * jb.cols + c] : 0.f;
      if (jb.type == 4) reinterpret_cast<__nv_bfloat16*>(jb.dst)[i] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(jb.dst)[i] = v;
    }
  }
}
bool prep_fused_enabled() {
  const char* e = getenv("MVAE_PREP_FUSED");
  return e ? atoi(e) != 0 : true;
}

template <typename TA>
int prep_weights(const Dims& d, const WS& w, const float* const* P, cudaStream_t st, bool need_bwd) {
  const int H = d.H, Hp = d.Hp;
  if constexpr (sizeof(TA) == 2) {
    if (rec_variant(d) >= 3 && prep_fused_enabled()) {
      PrepJobs jobs{};
      int n = 0;
      auto add = [&](int type, const float* src, const float* src2, void* dst, int h, int cols, int hp, int cols_p, int g0, int g1, int g2, long long total) {
        jobs.j[n++] = PrepJob{src, src2, dst, type, h, cols, hp, cols_p, g0, g1, g2, total};
      };
      const int Zp = round_up(d.Z, 8);
      for (int l = 0; l < d.L; ++l) {
        add(0, P[P_WHH(l)], nullptr, w.Whh_p[l], H, H, Hp, Hp, 0, 1, 2, 3ll * Hp * Hp);
        if (need_bwd) add(1, P[P_WHH(l)], nullptr, w.WhhT_p[l], H, H, Hp, Hp, 0, 1, 2, 3ll * Hp * Hp);
        add(2, P[P_BHH(l)], nullptr, w.bhh_p[l], H, 1, Hp, 1, 0, 1, 2, 3ll * Hp);
        add(2, P[P_BIH(l)], nullptr, w.bih_p[l], H, 1, Hp, 1, 0, 1, 2, 3ll * Hp);
        add(3, P[P_BIH(l)], P[P_BHH(l)], w.bcomb_p[l], H, 1, Hp, 1, 0, 1, 2, 3ll * Hp);
        if (l >= 1) {
          add(0, P[P_WIH(l)], nullptr, w.Wih_p[l], H, H, Hp, Hp, 0, 1, 2, 3ll * Hp * Hp);
          if (need_bwd) add(0, P[P_WIH(l)], nullptr, w.Wih_nrz[l], H, H, Hp, Hp, 2, 0, 1, 3ll * Hp * Hp);
        } else {
          add(0, P[P_WIH(0)], nullptr, w.Wih0_p, H, d.Z, Hp, Zp, 0, 1, 2, 3ll * Hp * Zp);
        }
      }
      add(4, P[P_FC3W(d.L)], nullptr, w.W3_p, d.C, H, d.CP, Hp, 0, 1, 2, (long long)d.CP * Hp);
      add(5, P[P_FC3B(d.L)], nullptr, w.b3_p, 1, d.C, 1, d.CP, 0, 1, 2, (long long)d.CP);
      jobs.n = n;
      prep_weights_fused_kernel<<<dim3(96, n), 256, 0, st>>>(jobs);
      KCHECK();
      return MVAE_OK;
    }
  }
  const int g = grid_for(3ll * Hp * Hp);
  for (int l = 0; l < d.L; ++l) {
    simt::pad_gate_matrix_kernel<TA><<<g, 256, 0, st>>>(P[P_WHH(l)], H, H, (TA*)w.Whh_p[l], Hp, Hp, 0, 1, 2);
    KCHECK();
    if (need_bwd && rec_variant(d) > 0) {
      pad_gate_matrix_T_kernel<<<g, 256, 0, st>>>(P[P_WHH(l)], H, (__nv_bfloat16*)w.WhhT_p[l], Hp);
      KCHECK();
    }
    if (l >= 1) {
      simt::pad_gate_matrix_kernel<TA><<<g, 256, 0, st>>>(P[P_WIH(l)], H, H, (TA*)w.Wih_p[l], Hp, Hp, 0, 1, 2);
      KCHECK();
      if (need_bwd) {
        simt::pad_gate_matrix_kernel<TA><<<g, 256, 0, st>>>(P[P_WIH(l)], H, H, (TA*)w.Wih_nrz[l], Hp, Hp, 2, 0, 1);
        KCHECK();
      }
      simt::pad_gate_vector_kernel<<<ceil_div(3 * Hp, 256), 256, 0, st>>>(P[P_BIH(l)], H, w.bih_p[l], Hp);
      KCHECK();
    }
    simt::pad_gate_vector_kernel<<<ceil_div(3 * Hp, 256), 256, 0, st>>>(P[P_BHH(l)], H, w.bhh_p[l], Hp);
    KCHECK();
    if (l >= 1 && rec_variant(d) >= 3) {
      combine_bias_kernel<<<ceil_div(3 * Hp, 256), 256, 0, st>>>(w.bih_p[l], w.bhh_p[l], w.bcomb_p[l], Hp);
      KCHECK();
    }
    if constexpr (sizeof(TA) == 2) {
      if (l == 0 && rec_variant(d) >= 3) {
        const int Zp = round_up(d.Z, 8);
        simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hp * Zp), 256, 0, st>>>(P[P_WIH(0)], H, d.Z, (TA*)w.Wih0_p, Hp, Zp, 0, 1, 2);
        KCHECK();
        simt::pad_gate_vector_kernel<<<ceil_div(3 * Hp, 256), 256, 0, st>>>(P[P_BIH(0)], H, w.bih_p[0], Hp);
        KCHECK();
        combine_bias_kernel<<<ceil_div(3 * Hp, 256), 256, 0, st>>>(w.bih_p[0], w.bhh_p[0], w.bcomb_p[0], Hp);
        KCHECK();
      }
    }
  }
  simt::pad_matrix_kernel<TA><<<grid_for((long long)d.CP * Hp), 256, 0, st>>>(P[P_FC3W(d.L)], d.C, H, (TA*)w.W3_p, d.CP, Hp);
  KCHECK();
  simt::pad_matrix_kernel<float><<<1, 64, 0, st>>>(P[P_FC3B(d.L)], 1, d.C, w.b3_p, 1, d.CP);
  KCHECK();
  return MVAE_OK;
}

// ---- forward --------------------------------------------------------------------------------
// from_z: decode-only entry (z given in w.z); save: keep BPTT state
template <typename TA>
int run_forward(const Dims& d, const WS& w, const float* const* P, const uint8_t* ids, const float* eps,
                cudaStream_t st, bool decode_only, bool save, bool fuse_head = false) {
  const int B = d.B, Bp = d.Bp, T = d.T, Z = d.Z, H = d.H, Hp = d.Hp;
  const size_t slab = (size_t)Bp * Hp;
  RC(memset_async(w.err_flag, 4, st));
  RC(memset_async(w.bce_sum, 8, st));
  RC(memset_async(w.kl_sum, 8, st));
  RC(memset_async(w.hit_count, (size_t)B * 4, st));
  RC(memset_async(w.gi0, (size_t)Bp * 3 * Hp * 4, st));
  for (int l = 0; l < d.L; ++l) {
    RC(memset_async(w.hs[l], slab * sizeof(TA), st));
    if constexpr (sizeof(TA) == 2) {
      if (ones_column(d)) {
        set_column_kernel<<<ceil_div(Bp, 256), 256, 0, st>>>((__nv_bfloat16*)w.hs[l], Bp, Hp, Hp - 1, 1.0f);
        KCHECK();
      }
    }
  }
  if (!decode_only) {
    simt::ConvDims cd{T, d.C, d.L1, d.L2, d.L3};
    const size_t smem = (729 + 990 + 9 * d.L1 + 9 * d.L2) * 4 + round_up(T, 4);
    simt::enc_conv_fwd_kernel<<<min(B, 148 * 4), 256, smem, st>>>(ids, B, cd, P[P_C1W], P[P_C1B], P[P_C2W], P[P_C2B],
                                                                   P[P_C3W], P[P_C3B], w.h1, w.h2, w.h3);
    KCHECK();
    // fc0 + SELU (models2d.py:28)
    RC(sg(st, w.h3, d.FLAT, 1, P[P_FC0W], 1, d.FLAT, w.h4, d.F0, B, d.F0, d.FLAT, P[P_FC0B], simt::ACT_SELU, 0));
    // fc11 / fc12 (models2d.py:29)
    RC(sg(st, w.h4, d.F0, 1, P[P_FC11W], 1, d.F0, w.mu, Z, B, Z, d.F0, P[P_FC11B], simt::ACT_NONE, 0));
    RC(sg(st, w.h4, d.F0, 1, P[P_FC12W], 1, d.F0, w.lv, Z, B, Z, d.F0, P[P_FC12B], simt::ACT_NONE, 0));
    simt::reparam_kl_kernel<<<grid_for((long long)B * Z, 256, 592), 256, 0, st>>>(w.mu, w.lv, eps, d.eps_scale,
                                                                                  d.train ? 1 : 0, (long long)B * Z,
                                                                                  w.z, w.kl_sum);
    KCHECK();
  }
  // fc2 + SELU (models2d.py:41); layer-0 input projection computed once per molecule
  RC(sg(st, w.z, Z, 1, P[P_FC2W], 1, Z, w.zr, Z, B, Z, Z, P[P_FC2B], simt::ACT_SELU, 0));
  bool gi0_tc = false;
  if constexpr (sizeof(TA) == 2) {
    if (rec_variant(d) >= 3) {
      // layer-0 projection on the tensor cores straight into the row-blocked bf16 operand of the fused recurrence
      // (bias = b_ih + (b_hr, b_hz, 0)); pad rows get the bias only, which is harmless (their gradients are zero)
      const int Zp = round_up(Z, 8);
      simt::pad_matrix_kernel<TA><<<grid_for((long long)Bp * Zp), 256, 0, st>>>(w.zr, B, Z, (TA*)w.zr_bf, Bp, Zp);
      KCHECK();
      RC(gemm<TA>(d, w, st, (const TA*)w.zr_bf, Zp, false, (const TA*)w.Wih0_p, Zp, true, w.gi0_bf, 3 * Hp, true, Bp, 3 * Hp, Zp,
                  w.bcomb_p[0], false, 1, 0, true));
      gi0_tc = true;
    }
  }
  if (!gi0_tc)
    for (int g = 0; g < 3; ++g)
      RC(sg(st, w.zr, Z, 1, P[P_WIH(0)] + (size_t)g * H * Z, 1, Z, w.gi0 + (size_t)g * Hp, 3 * Hp, B, H, Z,
            P[P_BIH(0)] + (size_t)g * H, simt::ACT_NONE, 0));
  const int gate_grid = ceil_div(Bp * Hp, 256);
  for (int l = 0; l < d.L; ++l) {
    TA* hs = (TA*)w.hs[l];
    TA* sv = (TA*)w.sv[l];
    if (l >= 1) {
      const TA* X = (const TA*)w.hs[l - 1] + slab;
      RC(gemm<TA>(d, w, st, X, Hp, false, (const TA*)w.Wih_p[l], Hp, true, w.gi_all, 3 * Hp, true, T * Bp, 3 * Hp, Hp,
                  rec_variant(d) >= 3 ? w.bcomb_p[l] : w.bih_p[l], false, 1, 0, rec_variant(d) >= 3));
    }
    const int rv = rec_variant(d);
    if (rv > 0) {
      if constexpr (sizeof(TA) == 2) {
        const __nv_bfloat16* gi = (const __nv_bfloat16*)w.gi_all;
        long long gstride = (long long)Bp * 3 * Hp;
        if (l == 0) {
          if (!gi0_tc) {
            gi0_to_bf16_kernel<<<grid_for((long long)Bp * 3 * Hp), 256, 0, st>>>(
                w.gi0, rv >= 3 ? w.bhh_p[0] : nullptr, Hp, (__nv_bfloat16*)w.gi0_bf, (long long)Bp * 3 * Hp, rv >= 3 ? 1 : 0);
            KCHECK();
          }
          gi = (const __nv_bfloat16*)w.gi0_bf;
          gstride = 0;
        }
        mvae_gru_rec_args ra{};
        ra.backward = 0; ra.variant = rv; ra.Bp = Bp; ra.Hp = Hp; ra.T = T;
        ra.W = (const __nv_bfloat16*)w.Whh_p[l]; ra.gi = gi; ra.gi_tstride = gstride; ra.bhh = w.bhh_p[l];
        ra.hs = (__nv_bfloat16*)hs; ra.sv = save ? (__nv_bfloat16*)sv : nullptr; ra.counters = w.counters;
        ra.err_flag = w.err_flag;
        ra.ones_col = ones_column(d) ? Hp - 1 : -1;
        count(2);
        ProfScope prof(0, st);
        if (rv >= 3) {
          ra.bhh = w.bhh_p[l] + 2 * Hp;
          RC(mvae_gru_rec2_launch(&ra, fast_gates(), st));
        } else {
          RC(mvae_gru_rec_launch(&ra, st));
        }
      }
      continue;
    }
    if (d.bf16) RC(memset_async(w.h32[0], slab * 4, st));
    for (int t = 0; t < T; ++t) {
      RC(gemm<TA>(d, w, st, hs + t * slab, Hp, false, (const TA*)w.Whh_p[l], Hp, true, w.gh, 3 * Hp, false, Bp, 3 * Hp,
                  Hp, w.bhh_p[l], false, 1));
      TA* svt = save ? sv + (size_t)t * Bp * 4 * Hp : nullptr;
      float* hp32 = d.bf16 ? w.h32[t & 1] : nullptr;
      float* hn32 = d.bf16 ? w.h32[(t + 1) & 1] : nullptr;
      if (l == 0)
        simt::gru_gate_fwd_kernel<TA, float><<<gate_grid, 256, 0, st>>>(w.gi0, w.gh, hp32, hs + t * slab,
                                                                        hs + (t + 1) * slab, hn32, svt, Bp, Hp);
      else
        simt::gru_gate_fwd_kernel<TA, TA><<<gate_grid, 256, 0, st>>>((const TA*)w.gi_all + (size_t)t * Bp * 3 * Hp,
                                                                     w.gh, hp32, hs + t * slab, hs + (t + 1) * slab,
                                                                     hn32, svt, Bp, Hp);
      KCHECK();
    }
  }
  // vocabulary head logits (models2d.py:44-45)
  const TA* top = (const TA*)w.hs[d.L - 1] + slab;
  if constexpr (sizeof(TA) == 2) {
    if (fuse_head) {
      // fused head (models2d.py:44-46 + train.py:31-35): logits -> softmax -> BCE -> d(logits) inside the GEMM epilogue;
      // the (T*B, C) logits / probabilities never reach HBM
      mvae_umma_operand a{top, 0, (long long)T * Bp, Hp, Hp, 1, 0, 0, 0};
      mvae_umma_operand b{w.W3_p, 0, d.CP, Hp, Hp, 1, 0, 0, 0};
      mvae_umma_out o{w.logits, d.CP, 0, 0, w.b3_p, 0};
      mvae_umma_head h{ids, B, Bp, T, d.C, d.max_len / ((float)B * (float)T * (float)d.C), w.dlogits, w.bce_sum, w.hit_count};
      count();
      return mvae_umma_gemm(&a, &b, &o, T * Bp, d.CP, Hp, 64, 1, 0, w.err_flag, st, &h);
    }
  }
  RC(gemm<TA>(d, w, st, top, Hp, false, (const TA*)w.W3_p, Hp, true, w.logits, d.CP, false, T * Bp, d.CP, Hp, w.b3_p,
              false, 1, 64));
  return MVAE_OK;
}

// ---- backward -------------------------------------------------------------------------------
// expects w.dlogits filled.  kl_internal: add the swapped-KL gradient; ext_dmu/ext_dlv optional.
template <typename TA>
// phase: -1 = everything; otherwise only the part whose gradients become final in phase p of L (data-parallel bucket
// order): p = 0 head + top GRU layer, p = k GRU layer L-1-k, p = L-1 additionally latent + encoder.
int run_backward(const Dims& d, const WS& w, const float* const* P, float* const* G, const uint8_t* ids,
                 const float* eps, cudaStream_t st, bool kl_internal, const float* ext_dmu, const float* ext_dlv,
                 int phase = -1) {
  const int B = d.B, Bp = d.Bp, T = d.T, Z = d.Z, H = d.H, Hp = d.Hp, CP = d.CP, L = d.L;
  const size_t slab = (size_t)Bp * Hp;
  const int TB = T * Bp;
  const int wsplits = d.bf16 ? max(1, min(64, (148 * 2) / (ceil_div(3 * Hp, 128) * ceil_div(Hp, 256)))) : 64;
  const TA* dlog = (const TA*)w.dlogits;
  const bool ones = ones_column(d);
  if (phase <= 0) {
  // head: dX = dlogits * W3 ; dW3 = dlogits^T * h_top ; db3 = colsum(dlogits)
  RC(gemm<TA>(d, w, st, dlog, CP, false, (const TA*)w.W3_p, Hp, false, w.dX, Hp, true, TB, Hp, CP, nullptr, false, 1, 0,
              rec_variant(d) >= 3));
  RC(memset_async(w.dW3_p, (size_t)CP * Hp * 4, st));
  RC(gemm<TA>(d, w, st, dlog, CP, true, (const TA*)w.hs[L - 1] + slab, Hp, false, w.dW3_p, Hp, false, CP, Hp, TB,
              nullptr, true, d.bf16 ? 148 : 64, 256));
  simt::unpad_matrix_kernel<<<grid_for((long long)d.C * H), 256, 0, st>>>(w.dW3_p, Hp, G[P_FC3W(L)], d.C, H);
  KCHECK();
  if (ones) {  // db3 = ones column of dW3
    strided_copy_kernel<<<1, 64, 0, st>>>(w.dW3_p + (Hp - 1), Hp, G[P_FC3B(L)], d.C);
    KCHECK();
  } else {
    RC(memset_async(w.csum, (size_t)4 * Hp * 4, st));
    RC(simt::colsum<TA>(st, dlog, TB, CP, CP, w.csum)); count();
    copy_prefix_kernel<<<1, 64, 0, st>>>(w.csum, G[P_FC3B(L)], d.C);
    KCHECK();
  }
  }

  const int gate_grid = ceil_div(Bp * Hp, 256);
  for (int l = L - 1; l >= 0; --l) {
    if (phase >= 0 && l != L - 1 - phase) continue;
    const TA* hs = (const TA*)w.hs[l];
    const TA* sv = (const TA*)w.sv[l];
    TA* dG = (TA*)w.dG;
    TA* dX = (TA*)w.dX;
    const int rv = rec_variant(d);
    if (rv > 0) {
      if constexpr (sizeof(TA) == 2) {
        mvae_gru_rec_args ra{};
        ra.backward = 1; ra.variant = rv; ra.Bp = Bp; ra.Hp = Hp; ra.T = T;
        ra.W = (const __nv_bfloat16*)w.WhhT_p[l]; ra.hs = (__nv_bfloat16*)w.hs[l]; ra.sv = (__nv_bfloat16*)w.sv[l];
        ra.dX = (const __nv_bfloat16*)dX; ra.dG = (__nv_bfloat16*)dG; ra.counters = w.counters;
        ra.err_flag = w.err_flag;
        count(2);
        {
          ProfScope prof(1, st);
          if (rv >= 3) RC(mvae_gru_rec2_launch(&ra, 0, st));
          else RC(mvae_gru_rec_launch(&ra, st));
        }
        if (l == 0) {
          dgi_time_sum_kernel<<<(unsigned)ceil_div64((long long)Bp * 3 * Hp / 8, 256), 256, 0, st>>>(
              (const __nv_bfloat16*)dG, T, Bp, Hp, w.dgi0sum, rv >= 3 ? (__nv_bfloat16*)w.dgi0sum_bf : nullptr);
          KCHECK();
        }
      }
    } else {
    RC(memset_async(w.dh_carry, slab * 4, st));
    if (l == 0) RC(memset_async(w.dgi0sum, (size_t)Bp * 3 * Hp * 4, st));
    for (int t = T - 1; t >= 0; --t) {
      TA* dGt = dG + (size_t)t * Bp * 4 * Hp;
      simt::gru_gate_bwd_kernel<TA><<<gate_grid, 256, 0, st>>>(sv + (size_t)t * Bp * 4 * Hp, hs + t * slab,
                                                               dX + t * slab, w.dh_carry, dGt,
                                                               l == 0 ? w.dgi0sum : nullptr, Bp, Hp);
      KCHECK();
      if (t > 0)  // dh_{t-1} += dgh_t * W_hh
        RC(gemm<TA>(d, w, st, dGt + Hp, 4 * Hp, false, (const TA*)w.Whh_p[l], Hp, false, w.dh_carry, Hp, false, Bp, Hp,
                    3 * Hp, nullptr, true, 1));
    }
    }
    // dW_hh = dgh^T * h_{t-1}   (K = T*Bp)
    RC(memset_async(w.dW_p, (size_t)3 * Hp * Hp * 4, st));
    RC(gemm<TA>(d, w, st, dG + Hp, 4 * Hp, true, hs, Hp, false, w.dW_p, Hp, false, 3 * Hp, Hp, TB, nullptr, true,
                wsplits, 256));
    simt::unpad_gate_matrix_kernel<<<grid_for(3ll * H * H), 256, 0, st>>>(w.dW_p, Hp, Hp, G[P_WHH(l)], H, H, 0, 1, 2);
    KCHECK();
    if (ones) {
      // bias grads = ones column of the padded weight gradients (h pad column Hp-1 is the constant 1)
      gate_bias_from_dw_kernel<<<ceil_div(3 * H, 256), 256, 0, st>>>(w.dW_p, H, Hp, 0, 1, 2, G[P_BHH(l)]);
      KCHECK();
      if (l == 0) {
        RC(memset_async(w.csum, (size_t)4 * Hp * 4, st));
        RC(simt::colsum<float>(st, w.dgi0sum, Bp, 3 * Hp, 3 * Hp, w.csum)); count();
        unpad_gate_vector_kernel<<<ceil_div(3 * H, 256), 256, 0, st>>>(w.csum, Hp, G[P_BIH(0)], H);
        KCHECK();
      }
    } else {
      // bias grads from the column sums of dG
      RC(memset_async(w.csum, (size_t)4 * Hp * 4, st));
      RC(simt::colsum<TA>(st, dG, TB, 4 * Hp, 4 * Hp, w.csum)); count();
      gate_bias_grads_kernel<<<ceil_div(3 * H, 256), 256, 0, st>>>(w.csum, H, Hp, G[P_BIH(l)], G[P_BHH(l)]);
      KCHECK();
    }
    if (l >= 1) {
      const TA* X = (const TA*)w.hs[l - 1] + slab;
      RC(memset_async(w.dW_p, (size_t)3 * Hp * Hp * 4, st));
      RC(gemm<TA>(d, w, st, dG, 4 * Hp, true, X, Hp, false, w.dW_p, Hp, false, 3 * Hp, Hp, TB, nullptr, true, wsplits,
                  256));
      simt::unpad_gate_matrix_kernel<<<grid_for(3ll * H * H), 256, 0, st>>>(w.dW_p, Hp, Hp, G[P_WIH(l)], H, H, 2, 0, 1);
      KCHECK();
      if (ones) {
        gate_bias_from_dw_kernel<<<ceil_div(3 * H, 256), 256, 0, st>>>(w.dW_p, H, Hp, 2, 0, 1, G[P_BIH(l)]);
        KCHECK();
      }
      // gradient into the layer below: dX = dgi * W_ih
      RC(gemm<TA>(d, w, st, dG, 4 * Hp, false, (const TA*)w.Wih_nrz[l], Hp, false, w.dX, Hp, true, TB, Hp, 3 * Hp,
                  nullptr, false, 1, 0, rec_variant(d) >= 3));
    } else {
      // time-invariant layer-0 input: dW_ih0 = (sum_t dgi)^T zr ; dzr = (sum_t dgi) W_ih0
      bool tc = false;
      if constexpr (sizeof(TA) == 2) {
        if (rec_variant(d) >= 3) {
          const int Zp = round_up(Z, 8);
          RC(memset_async(w.dWih0_p, (size_t)3 * Hp * Zp * 4, st));
          RC(gemm<TA>(d, w, st, (const TA*)w.dgi0sum_bf, 3 * Hp, true, (const TA*)w.zr_bf, Zp, false, w.dWih0_p, Zp, false, 3 * Hp, Zp,
                      Bp, nullptr, true, max(1, min(8, Bp / 512)), 256));
          simt::unpad_gate_matrix_kernel<<<grid_for(3ll * H * Z), 256, 0, st>>>(w.dWih0_p, Hp, Zp, G[P_WIH(0)], H, Z, 0, 1, 2);
          KCHECK();
          RC(gemm<TA>(d, w, st, (const TA*)w.dgi0sum_bf, 3 * Hp, false, (const TA*)w.Wih0_p, Zp, false, w.dzr, Z, false, B, Z, 3 * Hp,
                      nullptr, false, 1));
          tc = true;
        }
      }
      for (int g = 0; g < 3 && !tc; ++g) {
        RC(sg_wgrad(st, w.dgi0sum + (size_t)g * Hp, 1, 3 * Hp, w.zr, Z, 1, G[P_WIH(0)] + (size_t)g * H * Z, Z, H, Z, B));
        RC(sg(st, w.dgi0sum + (size_t)g * Hp, 3 * Hp, 1, P[P_WIH(0)] + (size_t)g * H * Z, Z, 1, w.dzr, Z, B, Z, H,
              nullptr, simt::ACT_NONE, g > 0 ? 1 : 0));
      }
    }
  }
  if (phase >= 0 && phase != L - 1) return MVAE_OK;
  // fc2 (+SELU)
  const long long nBZ = (long long)B * Z;
  simt::selu_bwd_kernel<<<grid_for(nBZ), 256, 0, st>>>(w.zr, w.dzr, w.da5, nBZ);
  KCHECK();
  RC(sg_wgrad(st, w.da5, 1, Z, w.z, Z, 1, G[P_FC2W], Z, Z, Z, B));
  RC(memset_async(G[P_FC2B], (size_t)Z * 4, st));
  RC(simt::colsum<float>(st, w.da5, B, Z, Z, G[P_FC2B])); count();
  RC(sg(st, w.da5, Z, 1, P[P_FC2W], Z, 1, w.dz, Z, B, Z, Z, nullptr, simt::ACT_NONE, 0));
  // reparametrisation + KL
  simt::reparam_kl_bwd_kernel<<<grid_for(nBZ), 256, 0, st>>>(w.mu, w.lv, eps, d.eps_scale, d.train ? 1 : 0, w.dz,
                                                            kl_internal ? 1.0f / (float)nBZ : 0.f, ext_dmu, ext_dlv,
                                                            nBZ, w.dmu, w.dlv);
  KCHECK();
  // fc11 / fc12
  RC(sg_wgrad(st, w.dmu, 1, Z, w.h4, d.F0, 1, G[P_FC11W], d.F0, Z, d.F0, B));
  RC(sg_wgrad(st, w.dlv, 1, Z, w.h4, d.F0, 1, G[P_FC12W], d.F0, Z, d.F0, B));
  RC(memset_async(G[P_FC11B], (size_t)Z * 4, st));
  RC(simt::colsum<float>(st, w.dmu, B, Z, Z, G[P_FC11B])); count();
  RC(memset_async(G[P_FC12B], (size_t)Z * 4, st));
  RC(simt::colsum<float>(st, w.dlv, B, Z, Z, G[P_FC12B])); count();
  RC(sg(st, w.dmu, Z, 1, P[P_FC11W], d.F0, 1, w.dh4, d.F0, B, d.F0, Z, nullptr, simt::ACT_NONE, 0));
  RC(sg(st, w.dlv, Z, 1, P[P_FC12W], d.F0, 1, w.dh4, d.F0, B, d.F0, Z, nullptr, simt::ACT_NONE, 1));
  const long long nBF = (long long)B * d.F0;
  simt::selu_bwd_kernel<<<grid_for(nBF), 256, 0, st>>>(w.h4, w.dh4, w.dh4, nBF);
  KCHECK();
  // fc0
  RC(sg_wgrad(st, w.dh4, 1, d.F0, w.h3, d.FLAT, 1, G[P_FC0W], d.FLAT, d.F0, d.FLAT, B));
  RC(memset_async(G[P_FC0B], (size_t)d.F0 * 4, st));
  RC(simt::colsum<float>(st, w.dh4, B, d.F0, d.F0, G[P_FC0B])); count();
  RC(sg(st, w.dh4, d.F0, 1, P[P_FC0W], d.FLAT, 1, w.dflat, d.FLAT, B, d.FLAT, d.F0, nullptr, simt::ACT_NONE, 0));
  // convs
  RC(memset_async(G[P_C1W], (size_t)9 * T * 9 * 4, st));
  RC(memset_async(G[P_C1B], 9 * 4, st));
  RC(memset_async(G[P_C2W], 729 * 4, st));
  RC(memset_async(G[P_C2B], 9 * 4, st));
  RC(memset_async(G[P_C3W], 990 * 4, st));
  RC(memset_async(G[P_C3B], 10 * 4, st));
  {
    simt::ConvDims cd{T, d.C, d.L1, d.L2, d.L3};
    const size_t smem = (size_t)(9 * T * 9 + 729 + 990 + 32 + 729 + 990 + 9 * d.L1 + 9 * d.L2 + 10 * d.L3 + 9 * d.L2 +
                                 9 * d.L1) * 4 + round_up(T, 4);
    static size_t attr_cache[64] = {0};
    MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(simt::enc_conv_bwd_kernel), 100 * 1024, attr_cache));
    if (smem > 100 * 1024) return MVAE_ERR_UNSUPPORTED;
    simt::enc_conv_bwd_kernel<<<min(B, 148 * 2), 256, smem, st>>>(ids, B, cd, P[P_C2W], P[P_C3W], w.h1, w.h2, w.h3,
                                                                   w.dflat, G[P_C1W], G[P_C1B], G[P_C2W], G[P_C2B],
                                                                   G[P_C3W], G[P_C3B]);
    KCHECK();
  }
  return MVAE_OK;
}

template <typename TA>
int head_fused(const Dims& d, const WS& w, const uint8_t* ids, float* probs, bool want_dlogits, cudaStream_t st) {
  const long long rows = (long long)d.T * d.Bp;
  const float gscale = d.max_len / ((float)d.B * (float)d.T * (float)d.C);
  simt::head_softmax_bce_kernel<TA><<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(
      w.logits, d.CP, d.C, ids, d.B, d.Bp, d.T, gscale, want_dlogits ? (TA*)w.dlogits : nullptr, probs, w.bce_sum,
      w.hit_count);
  KCHECK();
  return MVAE_OK;
}

int finalize(const Dims& d, const WS& w, float* out_scalars, float* mu_out, float* lv_out, cudaStream_t st) {
  if (out_scalars) {
    simt::finalize_scalars_kernel<<<1, 256, 0, st>>>(w.bce_sum, w.kl_sum, w.hit_count, d.B, d.T,
                                                    (double)d.max_len / ((double)d.B * d.T * d.C),
                                                    1.0 / ((double)d.B * d.Z), out_scalars, w.err_flag);
    KCHECK();
  }
  const size_t n = (size_t)d.B * d.Z * 4;
  if (mu_out) { count(); MVAE_CUDA_CHECK(cudaMemcpyAsync(mu_out, w.mu, n, cudaMemcpyDeviceToDevice, st)); }
  if (lv_out) { count(); MVAE_CUDA_CHECK(cudaMemcpyAsync(lv_out, w.lv, n, cudaMemcpyDeviceToDevice, st)); }
  return MVAE_OK;
}

template <typename TA>
int elbo_step_t(const Dims& d, const WS& w, const float* const* P, float* const* G, const uint8_t* ids,
                const float* eps, float* out_scalars, float* mu_out, float* lv_out, cudaStream_t st, int phase = -1) {
  if (phase >= d.L) return MVAE_ERR_INVALID;
  if (phase <= 0) {
    RC(prep_weights<TA>(d, w, P, st, true));
    const bool fuse = sizeof(TA) == 2 && fuse_head_enabled();
    RC(run_forward<TA>(d, w, P, ids, eps, st, false, true, fuse));
    if (!fuse) RC(head_fused<TA>(d, w, ids, nullptr, true, st));
  }
  RC(run_backward<TA>(d, w, P, G, ids, eps, st, true, nullptr, nullptr, phase));
  if (phase < 0 || phase == d.L - 1) RC(finalize(d, w, out_scalars, mu_out, lv_out, st));
  return MVAE_OK;
}

int check_ws(const mvae_cfgb_desc* desc, void* ws, size_t ws_bytes, Dims* d, WS* w) {
  RC(make_dims(desc, d));
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) return MVAE_ERR_INVALID;
  carve(*d, ws, w);
  if (ws_bytes < w->total) return MVAE_ERR_WORKSPACE;
  g_tc = mvae_tc_ctx{(d->bf16 && tc_sgemm_enabled()) ? w->tcs : nullptr, w->tcs_bytes, w->err_flag};
  return MVAE_OK;
}

}  // namespace

void mvae_count_launches(int n) { g_launches += n; }   // used by the other orchestration units (moses.cu)

struct mvae_graph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  long long kernel_nodes;
  long long counted_launches;
};

// Capture whatever `fn` enqueues on a fresh stream into an instantiated CUDA graph (shared by cfgb.cu / cfga.cu).
int mvae_capture_into_graph(int (*fn)(void*, cudaStream_t), void* ctx, mvae_graph** out_graph) {
  if (!out_graph || !fn) return MVAE_ERR_INVALID;
  cudaStream_t cs;
  MVAE_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  const long long before = g_launches;
  MVAE_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  int rc = fn(ctx, cs);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(cs, &graph);
  const long long counted = g_launches - before;
  g_launches = before;
  cudaStreamDestroy(cs);
  if (rc != MVAE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  MVAE_CUDA_CHECK(e);
  mvae_graph* g = new (std::nothrow) mvae_graph();
  if (!g) return MVAE_ERR_INVALID;
  g->graph = graph;
  g->counted_launches = counted;
  e = cudaGraphInstantiate(&g->exec, graph, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(graph); delete g; MVAE_CUDA_CHECK(e); }
  size_t n = 0;
  cudaGraphGetNodes(graph, nullptr, &n);
  g->kernel_nodes = (long long)n;
  *out_graph = g;
  return MVAE_OK;
}

extern "C" {

int mvae_profile_enable(int on) {
  for (int i = 0; i < g_prof_n; ++i) { cudaEventDestroy(g_prof[i].e0); cudaEventDestroy(g_prof[i].e1); }
  g_prof_n = 0;
  g_prof_on = on != 0;
  return MVAE_OK;
}
int mvae_profile_read(int tag, float* total_ms, int* launches) {
  if (!total_ms || !launches || tag < 0 || tag >= PROF_TAGS) return MVAE_ERR_INVALID;
  float tot = 0.f; int n = 0;
  for (int i = 0; i < g_prof_n; ++i) {
    if (g_prof[i].tag != tag) continue;
    MVAE_CUDA_CHECK(cudaEventSynchronize(g_prof[i].e1));
    float ms = 0.f;
    MVAE_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof[i].e0, g_prof[i].e1));
    tot += ms; ++n;
  }
  *total_ms = tot; *launches = n;
  return MVAE_OK;
}
long long mvae_launch_count(void) { return g_launches; }
void mvae_reset_launch_count(void) { g_launches = 0; }

size_t mvae_cfgb_workspace_bytes(const mvae_cfgb_desc* desc) {
  Dims d; WS w;
  if (make_dims(desc, &d) != MVAE_OK) return 0;
  carve(d, nullptr, &w);
  return w.total;
}

int mvae_cfgb_elbo_step(const mvae_cfgb_desc* desc, const float* const* params, float* const* grads,
                        const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out, float* logvar_out,
                        void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  Dims d; WS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !grads || !ids || (d.train && !eps)) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d.bf16 ? elbo_step_t<__nv_bfloat16>(d, w, params, grads, ids, eps, out_scalars, mu_out, logvar_out, st)
                : elbo_step_t<float>(d, w, params, grads, ids, eps, out_scalars, mu_out, logvar_out, st);
}

int mvae_cfgb_elbo_step_phase(const mvae_cfgb_desc* desc, const float* const* params, float* const* grads,
                              const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out, float* logvar_out,
                              void* workspace, size_t workspace_bytes, int phase, mvae_stream_t stream) {
  Dims d; WS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !grads || !ids || (d.train && !eps) || phase < 0 || phase >= d.L) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d.bf16 ? elbo_step_t<__nv_bfloat16>(d, w, params, grads, ids, eps, out_scalars, mu_out, logvar_out, st, phase)
                : elbo_step_t<float>(d, w, params, grads, ids, eps, out_scalars, mu_out, logvar_out, st, phase);
}

int mvae_cfgb_elbo_step_phase_graph_create(const mvae_cfgb_desc* desc, const float* const* params, float* const* grads,
                                           const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                                           float* logvar_out, void* workspace, size_t workspace_bytes, int phase,
                                           mvae_graph** out_graph) {
  struct Ctx {
    const mvae_cfgb_desc* desc; const float* const* params; float* const* grads; const uint8_t* ids; const float* eps;
    float *out_scalars, *mu_out, *logvar_out; void* ws; size_t ws_bytes; int phase;
  } c{desc, params, grads, ids, eps, out_scalars, mu_out, logvar_out, workspace, workspace_bytes, phase};
  return mvae_capture_into_graph(
      [](void* p, cudaStream_t cs) {
        Ctx* c = static_cast<Ctx*>(p);
        return mvae_cfgb_elbo_step_phase(c->desc, c->params, c->grads, c->ids, c->eps, c->out_scalars, c->mu_out,
                                         c->logvar_out, c->ws, c->ws_bytes, c->phase, reinterpret_cast<mvae_stream_t>(cs));
      },
      &c, out_graph);
}

int mvae_cfgb_elbo_step_graph_create(const mvae_cfgb_desc* desc, const float* const* params, float* const* grads,
                                     const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                                     float* logvar_out, void* workspace, size_t workspace_bytes,
                                     mvae_graph** out_graph) {
  struct Ctx {
    const mvae_cfgb_desc* desc; const float* const* params; float* const* grads; const uint8_t* ids; const float* eps;
    float *out_scalars, *mu_out, *logvar_out; void* ws; size_t ws_bytes;
  } c{desc, params, grads, ids, eps, out_scalars, mu_out, logvar_out, workspace, workspace_bytes};
  return mvae_capture_into_graph(
      [](void* p, cudaStream_t cs) {
        Ctx* c = static_cast<Ctx*>(p);
        return mvae_cfgb_elbo_step(c->desc, c->params, c->grads, c->ids, c->eps, c->out_scalars, c->mu_out, c->logvar_out,
                                   c->ws, c->ws_bytes, reinterpret_cast<mvae_stream_t>(cs));
      },
      &c, out_graph);
}

int mvae_graph_launch(mvae_graph* g, mvae_stream_t stream) {
  if (!g) return MVAE_ERR_INVALID;
  MVAE_CUDA_CHECK(cudaGraphLaunch(g->exec, reinterpret_cast<cudaStream_t>(stream)));
  g_launches += g->counted_launches;
  return MVAE_OK;
}
long long mvae_graph_num_kernel_nodes(const mvae_graph* g) { return g ? g->kernel_nodes : 0; }
void mvae_graph_destroy(mvae_graph* g) {
  if (!g) return;
  cudaGraphExecDestroy(g->exec);
  cudaGraphDestroy(g->graph);
  delete g;
}

int mvae_cfgb_forward(const mvae_cfgb_desc* desc, const float* const* params, const uint8_t* ids, const float* eps,
                      float* probs, float* mu, float* logvar, void* workspace, size_t workspace_bytes,
                      mvae_stream_t stream) {
  Dims d; WS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !ids || (d.train && !eps)) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (d.bf16) {
    RC(prep_weights<__nv_bfloat16>(d, w, params, st, true));
    RC(run_forward<__nv_bfloat16>(d, w, params, ids, eps, st, false, true));
    RC(head_fused<__nv_bfloat16>(d, w, ids, probs, false, st));
  } else {
    RC(prep_weights<float>(d, w, params, st, true));
    RC(run_forward<float>(d, w, params, ids, eps, st, false, true));
    RC(head_fused<float>(d, w, ids, probs, false, st));
  }
  return finalize(d, w, nullptr, mu, logvar, st);
}

int mvae_cfgb_backward(const mvae_cfgb_desc* desc, const float* const* params, float* const* grads,
                       const uint8_t* ids, const float* eps, const float* dprobs, const float* dmu,
                       const float* dlogvar, void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  Dims d; WS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !grads || !ids || (d.train && !eps)) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = (long long)d.T * d.Bp;
  const size_t es = d.bf16 ? 2 : 4;
  if (!dprobs) {
    RC(memset_async(w.dlogits, (size_t)rows * d.CP * es, st));
  } else if (d.bf16) {
    simt::head_softmax_bwd_kernel<__nv_bfloat16><<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(
        w.logits, d.CP, d.C, dprobs, d.B, d.Bp, d.T, (__nv_bfloat16*)w.dlogits);
    KCHECK();
  } else {
    simt::head_softmax_bwd_kernel<float><<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(
        w.logits, d.CP, d.C, dprobs, d.B, d.Bp, d.T, (float*)w.dlogits);
    KCHECK();
  }
  return d.bf16 ? run_backward<__nv_bfloat16>(d, w, params, grads, ids, eps, st, false, dmu, dlogvar)
                : run_backward<float>(d, w, params, grads, ids, eps, st, false, dmu, dlogvar);
}

int mvae_cfgb_decode_greedy(const mvae_cfgb_desc* desc, const float* const* params, const float* z, uint8_t* ids_out,
                            float* probs_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  Dims d; WS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !z || !ids_out) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  count();
  MVAE_CUDA_CHECK(cudaMemcpyAsync(w.z, z, (size_t)d.B * d.Z * 4, cudaMemcpyDeviceToDevice, st));
  if (d.bf16) {
    RC(prep_weights<__nv_bfloat16>(d, w, params, st, false));
    RC(run_forward<__nv_bfloat16>(d, w, params, nullptr, nullptr, st, true, false));
  } else {
    RC(prep_weights<float>(d, w, params, st, false));
    RC(run_forward<float>(d, w, params, nullptr, nullptr, st, true, false));
  }
  const long long rows = (long long)d.T * d.Bp;
  simt::head_argmax_kernel<<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(w.logits, d.CP, d.C, d.B, d.Bp, d.T,
                                                                                 ids_out);
  KCHECK();
  if (probs_out) {
    // reuse the fused head in probs-only mode (ids only select the BCE target, which is discarded here)
    if (d.bf16) RC(head_fused<__nv_bfloat16>(d, w, ids_out, probs_out, false, st));
    else RC(head_fused<float>(d, w, ids_out, probs_out, false, st));
  }
  return MVAE_OK;
}

int mvae_onehot_to_ids(const float* onehot, long long rows, int charset, uint8_t* ids, int* not_onehot,
                       mvae_stream_t stream) {
  if (!onehot || !ids || !not_onehot || rows <= 0 || charset <= 0 || charset > 255) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  simt::onehot_to_ids_kernel<<<(unsigned)ceil_div64(rows, 256), 256, 0, st>>>(onehot, (int)rows, charset, ids, not_onehot);
  KCHECK();
  return MVAE_OK;
}

int mvae_cfgb_read_error(const mvae_cfgb_desc* desc, void* workspace, size_t workspace_bytes, int* flag,
                         mvae_stream_t stream) {
  Dims d; WS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!flag) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MVAE_CUDA_CHECK(cudaMemcpyAsync(flag, w.err_flag, 4, cudaMemcpyDeviceToHost, st));
  MVAE_CUDA_CHECK(cudaStreamSynchronize(st));
  return MVAE_OK;
}

int mvae_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major, void* D,
                   long long ldd, int d_is_bf16, int accumulate, const float* bias, int M, int N, int K, int tile_n,
                   int splits, int* err_flag, mvae_stream_t stream) {
  mvae_umma_operand a{A, a_mn_major ? 1 : 0, M, K, lda, 1, 0, 0, 0};
  mvae_umma_operand b{B, b_mn_major ? 1 : 0, N, K, ldb, 1, 0, 0, 0};
  mvae_umma_out o{D, ldd, d_is_bf16, accumulate, bias, 0};
  count();
  return mvae_umma_gemm(&a, &b, &o, M, N, K, tile_n, splits, 0, err_flag, reinterpret_cast<cudaStream_t>(stream));
}

int mvae_sgemm(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
               long long ldc, int M, int N, int K, const float* bias, int act, int accumulate, int splits,
               mvae_stream_t stream) {
  count();
  return simt::sgemm(reinterpret_cast<cudaStream_t>(stream), A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act,
                     accumulate, splits);
}

int mvae_sgemm_tc(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                  long long ldc, int M, int N, int K, const float* bias, int act, int accumulate, void* scratch,
                  size_t scratch_bytes, int* err_flag, mvae_stream_t stream) {
  mvae_tc_ctx ctx{scratch, scratch_bytes, err_flag};
  int n = 0;
  const int rc = mvae_tc_sgemm(&ctx, reinterpret_cast<cudaStream_t>(stream), A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act,
                               accumulate, &n);
  if (rc == MVAE_OK) count(n);
  return rc;
}
size_t mvae_sgemm_tc_scratch_bytes(long long rows_max, long long cols_max) { return mvae_tc_sgemm_scratch_bytes(rows_max, cols_max); }

}  // extern "C"
