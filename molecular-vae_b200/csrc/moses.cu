// MOSES-style character VAE (mosesvae.py:27-199): fused forward + loss + backward of one batch on one B200.
//
//   encoder  (mosesvae.py:142-164): trainable one-hot-initialised embedding -> packed GRU(V -> Hq) -> last state ->
//            two 2-layer MLP heads (mu, logvar) -> z = mu + exp(logvar/2) eps,  kl = 0.5 mean_b sum_d (e^lv + mu^2 - 1 - lv)
//   decoder  (mosesvae.py:166-199): input [emb(x_t) | z], h0 = decoder_lat(z) for every layer, GRU L x Hd, fc -> V,
//            recon = CE(y[:, :-1], x[:, 1:], ignore_index = pad)  (mean over the batch's non-pad targets)
// Layout choices (same conventions as cfgb.cu): time-major slabs [T][Bp][H], Bp = B rounded to 256; sequences are
// right-padded to T = max length.  Padded steps are simply computed: nothing a valid position needs depends on them
// (a sequence's padded steps come AFTER its valid ones) and every gradient that enters them is zero (ignore_index /
// final-state scatter), so no masking is needed inside the recurrence; the encoder's final state is captured at
// t = L_b - 1 by the gate kernel.  The embedding makes the layer-0 input projections table look-ups:
//   x_t W_ih^T = (E W_ih[:, :V]^T)[x_t]  (+ z W_ih[:, V:]^T + b_ih for the decoder, computed once per molecule),
// and their gradients one skinny tensor-core GEMM  dTBL = onehot^T * dgi  (K = T*B).
// The recurrence here is the per-step engine (tcgen05 GEMM + gate kernel per step); the persistent 2-CTA kernel of
// gru_rec2.cu needs an h0 / dh0 port before it can serve this path (next round).
#include <stdlib.h>

#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "umma_gemm.h"
#include "host_common.cuh"
#include "gru_rec.h"
#include "decode_persist.h"

namespace {

struct MDims {
  int B, Bp, T, V, CP, Z, Hq, Hd, L, MLP, pad;
  int bidir, lin, Hin;   // mosesfile.py variant: bidirectional encoder (:21-28), single-Linear heads (:31-32); Hin = Hq*(1+bidir)
  bool bf16;
  float kl_w, rec_w;
  float drop; unsigned int drop_seed;   // train-mode dropout between decoder layers (0 = off)
};

bool balanced_splits_enabled() {
  const char* e = getenv("MVAE_BALANCED_SPLITS");
  return e ? atoi(e) != 0 : true;
}
bool moses_fused_head_enabled() {
  const char* e = getenv("MVAE_FUSED_HEAD");
  return e ? atoi(e) != 0 : true;
}

int make_dims(const mvae_moses_desc* d, MDims* o) {
  if (!d) return MVAE_ERR_INVALID;
  if (d->batch <= 0 || d->max_len < 2 || d->max_len > 512 || d->vocab < 5 || d->vocab > 256 || d->d_z <= 0 ||
      d->q_hidden <= 0 || d->d_hidden <= 0 || d->d_layers < 1 || d->d_layers > 4 || d->mlp_hidden <= 0)
    return MVAE_ERR_INVALID;
  if ((d->q_hidden & 63) || (d->d_hidden & 63)) return MVAE_ERR_UNSUPPORTED;   // hidden sizes must be multiples of 64
  if (d->precision != MVAE_PREC_FP32 && d->precision != MVAE_PREC_BF16) return MVAE_ERR_INVALID;
  o->B = d->batch; o->Bp = round_up(d->batch, 256); o->T = d->max_len; o->V = d->vocab; o->CP = round_up(d->vocab, 64);   // logits / one-hot rows in 64-wide tiles
  o->Z = d->d_z; o->Hq = d->q_hidden; o->Hd = d->d_hidden; o->L = d->d_layers; o->MLP = d->mlp_hidden;
  o->pad = d->pad_id; o->bf16 = d->precision == MVAE_PREC_BF16; o->kl_w = d->kl_weight; o->rec_w = d->recon_weight;
  if (!(d->d_dropout >= 0.f && d->d_dropout < 1.f)) return MVAE_ERR_INVALID;
  o->drop = d->d_dropout; o->drop_seed = d->dropout_seed;
  o->bidir = d->q_bidir ? 1 : 0; o->lin = d->q_linear_heads ? 1 : 0; o->Hin = o->Hq * (1 + o->bidir);
  return MVAE_OK;
}

struct MWS {
  int* err_flag; double* kl_sum; double* nll_sum; int* M;
  int* act_dev;    // [T] packed-sequence batch sizes on the device (for the tile / k-block skipping of the big GEMMs)
  int* act256_dev; // [T] the same rounded up to the 256-row tiles of the persistent sweeps
  int* tileT;      // [Bp/256] time steps row tile j of the persistent sweeps has to run (= its longest sequence)
  float *TBLe, *TBLd, *hlast, *rmu, *rlv, *mu, *lv, *z, *h0, *zproj, *dh0, *dz, *dmu, *dlv, *dr, *dhenc, *dgisum;
  float *dTBL;      // [CP][3Hd] fp32, columns in (n,r,z) order
  float *dWT;       // [3Hd][V] staging
  void *OH;         // [T*Bp][CP] TA one-hot of the tokens
  void *gi;         // [T][Bp][3Hd] TA
  void *hs_enc, *sv_enc, *hs[4], *sv[4], *dG, *dX, *dlogits;
  // sampler with the fused GRU-cell GEMM epilogue (bf16 mode): [x | h] operands (ping-pong by step parity), permuted
  // concatenated weights [W_ih | W_hh] in tiles of 64 units x 4 gate blocks, fp32 master states
  void *xh[4][2]; void *Wcat[4]; float *bcat[4]; float *hm[4][2];
  void *WhhC[4]; float *bhhC[4];         // training decoder: W_hh / b_hh in the fused-cell tile order (G = 3)
  void *hdrop[4];                        // [l]: [T][Bp][Hd] dropped copy of decoder layer l-1's outputs (input of layer l), kept for backward
  void *hs_encr, *sv_encr, *Whh_encr;   // reverse direction of a bidirectional encoder
  float *bhh_encr, *TBLer, *hlastr, *hcat, *dhcat;
  float *gh, *h32[2], *dh_carry, *logits;
  void *Whh_enc, *Whh[4], *Wih[4], *Wih_nrz[4], *Wfc;
  float *bhh_enc, *bih[4], *bhh[4], *bfc;
  float *dW_p, *dWfc_p, *csum;
  // persistent decoder recurrence (gru_rec2.cu): W_hh^T [Hd][3Hd] bf16, b_ih + (b_hr, b_hz, 0), inter-CTA step counters
  void *WhhT[4]; float *bcomb[4]; unsigned int* counters;
  uint8_t* tokTr;   // ids^T with every row reversed inside its own length (reverse encoder direction on the persistent kernel)
  void *WhhT_enc; float *tbl_comb; uint8_t* tokT; void* zproj_rb;   // encoder W_hh^T, table + (b_hr, b_hz, 0), ids^T [T][Bp], zproj RB bf16
  void* tcs; size_t tcs_bytes;   // bf16 mode: converted operands of the bf16x3 tensor-core path of the small fp32 GEMMs
  unsigned int* dec_scratch;     // persistent decode kernel: completion counters + unit schedule
  size_t total;
};

void carve(const MDims& d, void* base, MWS* w) {
  Carver c{reinterpret_cast<uint8_t*>(base), 0};
  const size_t es = d.bf16 ? 2 : 4;
  const size_t B = d.B, Bp = d.Bp, T = d.T, Hd = d.Hd, Hq = d.Hq;
  w->err_flag = c.take<int>(1); w->kl_sum = c.take<double>(1); w->nll_sum = c.take<double>(1); w->M = c.take<int>(1);
  w->act_dev = c.take<int>(512); w->act256_dev = c.take<int>(512); w->tileT = c.take<int>(Bp / 256 + 8);
  w->TBLe = c.take<float>((size_t)d.V * 3 * Hq); w->TBLd = c.take<float>((size_t)d.V * 3 * Hd);
  w->hlast = c.take<float>(Bp * Hq); w->rmu = c.take<float>(B * d.MLP); w->rlv = c.take<float>(B * d.MLP);
  w->mu = c.take<float>(B * d.Z); w->lv = c.take<float>(B * d.Z); w->z = c.take<float>(B * d.Z);
  w->h0 = c.take<float>(Bp * Hd); w->zproj = c.take<float>(B * 3 * Hd); w->dh0 = c.take<float>(Bp * Hd);
  w->dz = c.take<float>(B * d.Z); w->dmu = c.take<float>(B * d.Z); w->dlv = c.take<float>(B * d.Z);
  w->dr = c.take<float>(B * d.MLP); w->dhenc = c.take<float>(B * d.Hin); w->dgisum = c.take<float>(Bp * 3 * Hd);
  w->dTBL = c.take<float>((size_t)d.CP * 3 * Hd); w->dWT = c.take<float>((size_t)3 * Hd * d.CP);
  w->hcat = c.take<float>(B * d.Hin); w->dhcat = c.take<float>(B * d.Hin);
  if (d.bidir) {
    w->TBLer = c.take<float>((size_t)d.V * 3 * Hq); w->hlastr = c.take<float>(Bp * Hq);
    w->hs_encr = c.take<uint8_t>((T + 1) * Bp * Hq * es); w->sv_encr = c.take<uint8_t>(T * Bp * 5 * Hq * es);
    w->Whh_encr = c.take<uint8_t>(3 * Hq * Hq * es); w->bhh_encr = c.take<float>(3 * Hq);
  } else {
    w->TBLer = nullptr; w->hlastr = nullptr; w->hs_encr = nullptr; w->sv_encr = nullptr; w->Whh_encr = nullptr; w->bhh_encr = nullptr;
  }
  w->OH = c.take<uint8_t>(T * Bp * d.CP * es);
  w->gi = c.take<uint8_t>(T * Bp * 3 * Hd * es);
  w->hs_enc = c.take<uint8_t>((T + 1) * Bp * Hq * es); w->sv_enc = c.take<uint8_t>(T * Bp * 5 * Hq * es);
  for (int l = 0; l < d.L; ++l) {
    w->hs[l] = c.take<uint8_t>((T + 1) * Bp * Hd * es);
    w->sv[l] = c.take<uint8_t>(T * Bp * 5 * Hd * es);   // 5: the persistent kernels also save h_{t-1} (fragment layout)
  }
  w->dG = c.take<uint8_t>(T * Bp * 4 * Hd * es); w->dX = c.take<uint8_t>(T * Bp * Hd * es);
  for (int l = 1; l < d.L; ++l) w->hdrop[l] = d.drop > 0.f ? c.take<uint8_t>(T * Bp * Hd * es) : nullptr;
  w->hdrop[0] = nullptr;
  for (int l = 0; l < d.L; ++l) {
    for (int k = 0; k < 2; ++k) { w->xh[l][k] = c.take<uint8_t>(Bp * 2 * Hd * 2); w->hm[l][k] = c.take<float>(Bp * Hd); }
    w->Wcat[l] = c.take<uint8_t>((size_t)4 * Hd * 2 * Hd * 2);
    w->bcat[l] = c.take<float>(4 * Hd);
    w->WhhC[l] = c.take<uint8_t>((size_t)3 * Hd * Hd * 2);
    w->bhhC[l] = c.take<float>(3 * Hd);
  }
  w->dlogits = c.take<uint8_t>(T * Bp * d.CP * es);
  w->gh = c.take<float>(Bp * 3 * Hd); w->h32[0] = c.take<float>(Bp * Hd); w->h32[1] = c.take<float>(Bp * Hd);
  w->dh_carry = c.take<float>(Bp * Hd); w->logits = c.take<float>(T * Bp * d.CP);
  w->Whh_enc = c.take<uint8_t>(3 * Hq * Hq * es); w->bhh_enc = c.take<float>(3 * Hq);
  for (int l = 0; l < d.L; ++l) {
    w->Whh[l] = c.take<uint8_t>(3 * Hd * Hd * es); w->Wih[l] = c.take<uint8_t>(3 * Hd * Hd * es);
    w->Wih_nrz[l] = c.take<uint8_t>(3 * Hd * Hd * es);
    w->bih[l] = c.take<float>(3 * Hd); w->bhh[l] = c.take<float>(3 * Hd);
  }
  w->Wfc = c.take<uint8_t>(d.CP * Hd * es); w->bfc = c.take<float>(d.CP);
  for (int l = 0; l < d.L; ++l) { w->WhhT[l] = c.take<uint8_t>(3 * Hd * Hd * 2); w->bcomb[l] = c.take<float>(3 * Hd); }
  w->counters = c.take<unsigned int>(2 * (Bp / 256) + 64);
  w->WhhT_enc = c.take<uint8_t>(3 * Hq * Hq * 2); w->tbl_comb = c.take<float>((size_t)d.CP * 3 * (Hd > Hq ? Hd : Hq));
  w->tokT = c.take<uint8_t>(T * Bp); w->zproj_rb = c.take<uint8_t>(Bp * 3 * Hd * 2);
  w->tokTr = d.bidir ? c.take<uint8_t>(T * Bp) : nullptr;
  w->dW_p = c.take<float>(3 * Hd * Hd); w->dWfc_p = c.take<float>(d.CP * Hd); w->csum = c.take<float>(4 * Hd);
  {
    const long long rmax = (long long)max(Bp, 3 * Hd), cmax = max(max((long long)3 * Hd, (long long)d.Hin), max((long long)d.MLP, max((long long)d.Z, (long long)d.CP)));
    w->tcs_bytes = d.bf16 ? mvae_tc_sgemm_scratch_bytes(rmax, cmax) : 0;
    w->tcs = c.take<uint8_t>(w->tcs_bytes);
  }
  w->dec_scratch = reinterpret_cast<unsigned int*>(c.take<uint8_t>(mvae_decode_persistent_scratch_bytes((int)Bp, d.L)));
  w->total = (c.off + 255) & ~size_t(255);
}

// parameter order = state_dict order (first occurrence of each tensor): oracle.moses_oracle.moses_shapes for mosesvae.VAE,
// oracle.moses_oracle.mosesfile_shapes for the bidirectional / single-Linear-head variant of mosesfile.py
struct MP {
  int bidir, lin, L;
  int emb() const { return 0; }
  int e_wih(int rev) const { return 1 + 4 * rev; }
  int e_whh(int rev) const { return 2 + 4 * rev; }
  int e_bih(int rev) const { return 3 + 4 * rev; }
  int e_bhh(int rev) const { return 4 + 4 * rev; }
  int heads() const { return 5 + 4 * bidir; }
  // MLP heads: mu {0.weight, 0.bias, 2.weight, 2.bias} then logvar {...}; Linear heads: mu {weight, bias}, logvar {weight, bias}
  int head_w0(int h) const { return heads() + (lin ? 2 : 4) * h; }
  int dec0() const { return heads() + (lin ? 4 : 8); }
  int wih(int l) const { return dec0() + 4 * l; }
  int whh(int l) const { return dec0() + 4 * l + 1; }
  int bih(int l) const { return dec0() + 4 * l + 2; }
  int bhh(int l) const { return dec0() + 4 * l + 3; }
  int latw() const { return dec0() + 4 * L; }
  int latb() const { return dec0() + 4 * L + 1; }
  int fcw() const { return dec0() + 4 * L + 2; }
  int fcb() const { return dec0() + 4 * L + 3; }
};

// ---------------------------------------------------------------------------------------------------------
// kernels specific to this path
// ---------------------------------------------------------------------------------------------------------
// z = mu + exp(lv/2) eps ; kl_sum += sum 0.5 (e^lv + mu^2 - 1 - lv)     (mosesvae.py:159-162)
__global__ void reparam_kl_std_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                      const float* __restrict__ eps, long long n, float* __restrict__ z,
                                      double* __restrict__ kl_sum) {
  double local = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], l = lv[i];
    z[i] = fmaf(eps[i], expf(0.5f * l), m);
    local += 0.5 * (double)(expf(l) + m * m - 1.0f - l);
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(kl_sum, s);
  }
}
__global__ void reparam_kl_std_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                          const float* __restrict__ eps, const float* __restrict__ dz, float klw_over_b,
                                          long long n, float* __restrict__ dmu, float* __restrict__ dlv) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mu[i], l = lv[i], g = dz[i];
    dmu[i] = g + klw_over_b * m;
    dlv[i] = g * eps[i] * 0.5f * expf(0.5f * l) + klw_over_b * 0.5f * (expf(l) - 1.0f);
  }
}
__global__ void relu_bwd_kernel(const float* __restrict__ out, float* __restrict__ d, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(out[i] > 0.f)) d[i] = 0.f;
}
// act[t] = number of sequences longer than t (lengths sorted descending -> they are the rows [0, act[t]))
__global__ void count_active_kernel(const int* __restrict__ lens, int B, int T, int* __restrict__ act, int* __restrict__ act256,
                                    int* __restrict__ tileT, int Bp) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < Bp / 256) tileT[t] = t * 256 < B ? min(T, lens[t * 256]) : 0;   // sorted descending: the tile's first row is its longest
  if (t >= T) return;
  int lo = 0, hi = B;                  // first index with lens <= t
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (lens[mid] > t) lo = mid + 1; else hi = mid; }
  act[t] = lo;
  act256[t] = min(Bp, (lo + 255) & ~255);
}
__global__ void count_targets_kernel(const int* __restrict__ lens, int B, int* __restrict__ M) {
  int local = 0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) local += lens[b] - 1;
  atomicAdd(M, local);
}
// slab-0 initialisation: hsA[b][j] = h0[b][j] (rows >= B zero), h32 copy
template <typename TA>
__global__ void init_h0_kernel(const float* __restrict__ h0, int B, int Bp, int H, TA* __restrict__ hsA,
                               float* __restrict__ h32) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * H) return;
  const int b = (int)(idx / H);
  const float v = b < B ? h0[idx] : 0.f;
  hsA[idx] = from_f32<TA>(v);
  if (h32) h32[idx] = v;
}
// shifted cross entropy with ignore_index (mosesvae.py:193-197): row (t,b) is a target position iff t+1 < L_b.
// One warp per row of logits[T*Bp][CP]; dlogits = (softmax - onehot(x[b][t+1])) / M on target rows, 0 elsewhere.
// y (optional, [B][T][V]): the reference's returned logits (decoder_fc(0) = bias at padded positions).
template <typename TA>
__global__ void head_ce_kernel(const float* __restrict__ logits, int CP, int V, const uint8_t* __restrict__ ids,
                               int ids_ld, const int* __restrict__ lens, int B, int Bp, int T,
                               const int* __restrict__ Mcount, float rec_w, const float* __restrict__ bias,
                               TA* __restrict__ dlogits,
                               float* __restrict__ y, double* __restrict__ nll_sum) {
  // lane l owns the vocabulary ids l, l + 32, ... (NV = CP / 32 <= 8 of them)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int NV = CP >> 5;
  double local = 0.0;
  if (warp < (long long)T * Bp) {
    const int t = warp / Bp, b = warp - t * Bp;
    const long long row = warp;
    const int L = b < B ? lens[b] : 0;
    const bool target = b < B && (t + 1 < L);
    float a[8], dv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int v = lane + 32 * k;
      a[k] = (k < NV && v < V) ? logits[row * CP + v] : -INFINITY;
      dv[k] = 0.f;
    }
    if (y && b < B) {
      float* yr = y + ((long long)b * T + t) * V;
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int v = lane + 32 * k; if (k < NV && v < V) yr[v] = t < L ? a[k] : bias[v]; }
    }
    if (target) {
      float m = a[0];
#pragma unroll
      for (int k = 1; k < 8; ++k) m = fmaxf(m, a[k]);
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float e[8], s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { e[k] = (k < NV && lane + 32 * k < V) ? expf(a[k] - m) : 0.f; s += e[k]; }
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int tgt = ids[(long long)b * ids_ld + t + 1];
      const float inv = rec_w / (float)Mcount[0];
      float at_l = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const bool hit = (lane + 32 * k) == tgt;
        dv[k] = (e[k] / s - (hit ? 1.f : 0.f)) * inv;
        if (hit) at_l = a[k];
      }
      const float at = __shfl_sync(0xffffffffu, at_l, tgt & 31);
      if (lane == 0) local = (double)(m + logf(s) - at);
    }
    if (dlogits) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < NV) dlogits[row * CP + lane + 32 * k] = from_f32<TA>((lane + 32 * k) < V ? dv[k] : 0.f);
    }
  }
  __shared__ double red[32];
  if (lane == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0 && nll_sum) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(nll_sum, s);
  }
}
// dX_enc[L_b - 1][b][:] = dh_enc[b][:] (the only gradient entering the encoder GRU); dX zeroed beforehand
template <typename TA>
__global__ void scatter_final_grad_kernel(const float* __restrict__ dh, int dh_ld, const int* __restrict__ lens, int B, int Bp,
                                          int H, TA* __restrict__ dX) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx - (long long)b * H);
  dX[((long long)(lens[b] - 1) * Bp + b) * H + j] = from_f32<TA>(dh[(long long)b * dh_ld + j]);
}
// sum over time of the dgi window of dG ([T][Bp][4H], blocks n,r,z) -> fp32 [Bp][3H] in (r,z,n) order
template <typename TA>
__global__ void dgi_time_sum_t_kernel(const TA* __restrict__ dG, int T, int Bp, int H, float* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * 3 * H) return;
  const int c = (int)(idx % (3 * H));
  const int b = (int)(idx / (3 * H));
  const int g = c / H, j = c - g * H;
  const int blk = (g == 0) ? 1 : (g == 1 ? 2 : 0);
  const TA* p = dG + (long long)b * 4 * H + blk * H + j;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += to_f32<TA>(p[(long long)t * Bp * 4 * H]);
  out[idx] = s;
}
// bf16 variant: one thread per (row, 8 consecutive columns of the (n,r,z) dgi window), 16-byte loads
// tileT (optional): rows of 256-row tile j only hold valid data for t < tileT[j]
__global__ void dgi_time_sum_bf16x8_kernel(const __nv_bfloat16* __restrict__ dG, int T, int Bp, int H, float* __restrict__ out,
                                           const int* __restrict__ tileT = nullptr) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int cpr = 3 * H / 8;
  if (idx >= (long long)Bp * cpr) return;
  const int c8 = (int)(idx % cpr) * 8, b = (int)(idx / cpr);
  const int blk = c8 / H, j = c8 - blk * H;      // blk: 0 = n, 1 = r, 2 = z
  const int g = (blk == 0) ? 2 : (blk - 1);      // -> (r,z,n) order of the output
  const uint4* p = reinterpret_cast<const uint4*>(dG + (long long)b * 4 * H + c8);
  const long long tstride = (long long)Bp * 4 * H / 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (tileT) T = min(T, tileT[b >> 8]);
#pragma unroll 4
  for (int t = 0; t < T; ++t) {
    const uint4 v = __ldg(p + (long long)t * tstride);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { s[2 * k] += __uint_as_float(w[k] << 16); s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u); }
  }
  float4* o = reinterpret_cast<float4*>(out + (long long)b * 3 * H + g * H + j);
  o[0] = make_float4(s[0], s[1], s[2], s[3]);
  o[1] = make_float4(s[4], s[5], s[6], s[7]);
}
// dW[(r,z,n) row g*H+j][v] = dTBL[v][blk(g)*H + j]   : transposes the (n,r,z)-ordered table gradient into torch's order
__global__ void tbl_grad_to_rzn_T_kernel(const float* __restrict__ dTBL, int H, int V, float* __restrict__ out /*[3H][V]*/) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= 3ll * H * V) return;
  const int v = (int)(idx % V);
  const int r = (int)(idx / V);
  const int g = r / H, j = r - g * H;
  const int blk = (g == 0) ? 1 : (g == 1 ? 2 : 0);
  out[idx] = dTBL[(long long)v * 3 * H + blk * H + j];
}
__global__ void gate_bias_grads_kernel(const float* __restrict__ csum, int H, float* __restrict__ db_ih,
                                       float* __restrict__ db_hh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * H) return;
  const int g = i / H, j = i - g * H;
  const int blk_ih = (g == 0) ? 1 : (g == 1 ? 2 : 0);
  const int blk_hh = (g == 0) ? 1 : (g == 1 ? 2 : 3);
  if (db_ih) db_ih[i] = csum[blk_ih * H + j];
  db_hh[i] = csum[blk_hh * H + j];
}
__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] += b[i];
}
__global__ void zero_row_kernel(float* __restrict__ a, int row, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cols) a[(long long)row * cols + i] = 0.f;
}
__global__ void finalize_kernel(const double* __restrict__ kl_sum, const double* __restrict__ nll_sum,
                                const int* __restrict__ M, int B, float klw, float recw, float* __restrict__ out) {
  const double kl = kl_sum[0] / B, rec = nll_sum[0] / (double)M[0];
  out[0] = (float)(klw * kl + recw * rec); out[1] = (float)kl; out[2] = (float)rec; out[3] = (float)M[0];
}

template <typename TA> __device__ __forceinline__ void load8(const TA* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
}
template <typename TA> __device__ __forceinline__ void store8(TA* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  __nv_bfloat162 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(o);
}
// counter-based uniform in (0,1): the same generator the sampler uses (restated in oracle/moses_oracle.u01_hash)
__device__ __forceinline__ float u01_hash(unsigned long long seed, unsigned int b, unsigned int i);
// dropout between decoder layers: element (t, b, j) of layer l is kept iff its 16-bit uniform (see the kernel) >= p.
// mode 0: out = keep ? x / (1 - p) : 0 (forward copy);  mode 1: x *= keep / (1 - p) in place (backward)
template <typename TA>
__global__ void dropout_kernel(const TA* x, TA* out /* may alias x */, unsigned long long seed, float p, int B, int Bp,
                               int H, int T, int rb) {
  // One thread = 8 consecutive hidden units of one (t, b) row (one 16-byte piece in bf16); blockIdx.y = t.
  // rb: x / out use the row-blocked layout [row/32][H/8][32][8] inside each time slab (dX of the persistent BPTT sweep).
  const int t = blockIdx.y;
  const int H8 = H >> 3;
  const unsigned pieces = (unsigned)Bp * (unsigned)H8;
  const float sc = 1.0f / (1.0f - p);
  const TA* xs = x + (size_t)t * Bp * H;
  TA* os = out + (size_t)t * Bp * H;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += gridDim.x * blockDim.x) {
    int b, cg;
    if (rb) { const unsigned q = i >> 5; cg = (int)(q % (unsigned)H8); b = (int)(q / (unsigned)H8) * 32 + (int)(i & 31u); }
    else { cg = (int)(i % (unsigned)H8); b = (int)(i / (unsigned)H8); }
    float v[8];
    load8<TA>(xs + (size_t)i * 8, v);
    // one 64-bit hash (the generator of u01_hash with counter j / 4) serves four consecutive units, 16 bits each:
    // keep(t, b, j) iff ((h >> 16 (j & 3)) & 0xFFFF + 0.5) / 65536 >= p      (restated in oracle/moses_oracle.dropout_masks)
    unsigned long long c = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)(unsigned)(t * B + b) * 1000003ull + (unsigned)(cg * 2) + 1);
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      unsigned long long h = c;
      h ^= h >> 30; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 27; h *= 0x94D049BB133111EBull; h ^= h >> 31;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float u = ((float)((unsigned)(h >> (16 * k)) & 0xFFFFu) + 0.5f) * (1.0f / 65536.0f);
        v[hq * 4 + k] = (b < B && u >= p) ? v[hq * 4 + k] * sc : 0.f;
      }
      c += 0x9E3779B97F4A7C15ull;
    }
    store8<TA>(os + (size_t)i * 8, v);
  }
}
template <typename TA>
int dropout_launch(cudaStream_t st, const TA* x, TA* out, unsigned long long seed, float p, int B, int Bp, int H, int T, int rb = 0) {
  const unsigned pieces = (unsigned)Bp * (unsigned)(H / 8);
  const long long gx = ceil_div64((long long)pieces, 256);
  dim3 grid((unsigned)(gx < 148 * 8 ? gx : 148 * 8), (unsigned)T);
  dropout_kernel<TA><<<grid, 256, 0, st>>>(x, out, seed, p, B, Bp, H, T, rb);
  KCHECK();
  return MVAE_OK;
}
template <typename TA>
__global__ void final_state_kernel(const TA* __restrict__ hsA, const float* __restrict__ h32, long long n, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = h32 ? h32[i] : to_f32<TA>(hsA[i]);
}
// dst[r][c] = src[r][c] for r < rows, c < cols (row strides sld / dld)
__global__ void copy_rows_kernel(const float* __restrict__ src, int sld, int rows, int cols, float* __restrict__ dst, int dld) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long long r = i / cols;
    dst[r * dld + c] = src[r * sld + c];
  }
}
// ---------------------------------------------------------------------------------------------------------
// per-step GRU engine (one layer), time-major, optional h0 / final-state capture / dh0
// ---------------------------------------------------------------------------------------------------------
// Weights of one decoder layer for the fused GRU-cell GEMM (umma_gemm.h mvae_umma_cell): rows in tiles of 64 units x G gate
// blocks.  first layer (G = 3, K = H): [W_hr | W_hz | W_hn];  other layers (G = 4, K = 2H, operand [x | h]):
// r: [W_ir | W_hr], z: [W_iz | W_hz], in: [W_in | 0], hn: [0 | W_hn].  bias in the same row order.
__global__ void cell_weights_kernel(const float* __restrict__ wih, const float* __restrict__ whh, const float* __restrict__ bih,
                                    const float* __restrict__ bhh, int H, int G, __nv_bfloat16* __restrict__ W,
                                    float* __restrict__ bias) {
  const int K = G == 4 ? 2 * H : H;
  const long long total = (long long)G * H * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % K);
    const int n = (int)(i / K);
    const int tile = n / (G * 64), g = (n / 64) % G, j = n % 64;
    const int unit = tile * 64 + j;
    float v = 0.f;
    if (G == 3) {
      v = whh[((long long)g * H + unit) * H + c];
    } else {
      const int gate = g < 2 ? g : 2;                       // source gate row block: r, z, n
      if (c < H) { if (g != 3) v = wih[((long long)gate * H + unit) * H + c]; }
      else if (g != 2) v = whh[((long long)gate * H + unit) * H + (c - H)];
    }
    W[i] = __float2bfloat16_rn(v);
    if (c == 0) {
      float b;
      if (G == 3) b = bhh[g * H + unit];
      else b = g < 2 ? bih[g * H + unit] + bhh[g * H + unit] : (g == 2 ? bih[2 * H + unit] : bhh[2 * H + unit]);
      bias[n] = b;
    }
  }
}
// Decoder recurrences on the persistent kernels of gru_rec2.cu (forward pair kernel + K-split BPTT sweep, one launch per
// layer and direction instead of one GEMM per step); bf16 mode, hidden size 256 or 512.  MVAE_MOSES_REC=0 keeps the
// per-step engine (the cross-check).
bool persistent_decoder(const MDims& d) {
  const char* e = getenv("MVAE_MOSES_REC");
  if (e && atoi(e) == 0) return false;
  return d.bf16 && (d.Hd == 256 || d.Hd == 512) && d.Bp % 256 == 0;
}
bool persistent_encoder(const MDims& d) {
  const char* e = getenv("MVAE_MOSES_REC");
  if (e && atoi(e) == 0) return false;
  return d.bf16 && (d.Hq == 256 || d.Hq == 512) && d.Bp % 256 == 0;
}
// W_hh^T for the BPTT sweep: dst[j][g*H + i] = W_hh[g*H + i][j]
__global__ void whh_transpose_kernel(const float* __restrict__ src, int H, __nv_bfloat16* __restrict__ dst) {
  const long long total = 3ll * H * H;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % (3 * H)), j = (int)(idx / (3 * H));
    dst[idx] = __float2bfloat16_rn(src[(long long)k * H + j]);
  }
}
__global__ void combine_gate_bias_kernel(const float* __restrict__ bih, const float* __restrict__ bhh, float* __restrict__ out, int H) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * H) out[i] = (bih ? bih[i] : 0.f) + (i < 2 * H ? bhh[i] : 0.f);
}
// per-molecule part of the decoder's layer-0 projection in the row-blocked layout [row/32][W/8][32][8] the persistent
// kernel's epilogue reads (time-invariant: gi_tstride 0): out[b][c] = src[b][c] + brz[c]; pad rows carry brz only
__global__ void rows_to_rb_kernel(const float* __restrict__ src, const float* __restrict__ brz, int B, int Bp, int W,
                                  __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)Bp * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % W), b = (int)(i / W);
    const long long o = ((long long)(b >> 5) * (W / 8) + (c >> 3)) * 256 + (b & 31) * 8 + (c & 7);
    out[o] = __float2bfloat16_rn((b < B ? src[i] : 0.f) + brz[c]);
  }
}
// Input projection of an embedding layer materialised for the persistent sweeps when the token table does not fit their
// shared memory (V > 64): out[t][b][:] = tbl[tok[t][b]][:] (+ add[b][:], row-blocked bf16) in the row-blocked layout
// [t][row/32][W/8][32][8]; one thread = 8 columns of one (t, row).  lim (optional, [T]): rows >= lim[t] are not written.
__global__ void gather_rb_kernel(const float* __restrict__ tbl, const uint8_t* __restrict__ tokT, const __nv_bfloat16* __restrict__ add_rb,
                                 int T, int Bp, int W, __nv_bfloat16* __restrict__ out, const int* __restrict__ lim) {
  const long long per_t = (long long)Bp * (W / 8);
  const long long total = (long long)T * per_t;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / per_t);
    const long long j = i - (long long)t * per_t;          // ((rblk * (W/8) + cg) * 32 + r)
    const int r = (int)(j & 31);
    const long long q = j >> 5;
    const int cg = (int)(q % (W / 8));
    const int b = (int)(q / (W / 8)) * 32 + r;
    if (lim && b >= lim[t]) continue;
    const float* src = tbl + (long long)tokT[(long long)t * Bp + b] * W + cg * 8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = src[k];
    if (add_rb) {
      float a[8];
      load8<__nv_bfloat16>(add_rb + j * 8, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += a[k];
    }
    store8<__nv_bfloat16>(out + (long long)t * Bp * W + j * 8, v);
  }
}
// table rows + (b_hr, b_hz, 0): what the persistent kernel stages into shared memory
__global__ void table_add_bias_kernel(const float* __restrict__ tbl, const float* __restrict__ brz, int V, int W, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V * W) out[i] = tbl[i] + (brz ? brz[i % W] : 0.f);
}
// tokT[t][b] = ids[b][t] (token 0 for pad rows); with lens: ids[b][lens[b]-1-t] for t < lens[b] (each row reversed inside its
// own length: the reverse GRU direction as a forward sweep whose padding still comes last), token 0 past the row's end
__global__ void transpose_ids_kernel(const uint8_t* __restrict__ ids, int ids_ld, int B, int Bp, int T, uint8_t* __restrict__ out,
                                     const int* __restrict__ lens = nullptr) {
  const long long total = (long long)T * Bp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i % Bp), t = (int)(i / Bp);
    uint8_t v = 0;
    if (b < B) {
      if (!lens) v = ids[(long long)b * ids_ld + t];
      else if (t < lens[b]) v = ids[(long long)b * ids_ld + (lens[b] - 1 - t)];
    }
    out[i] = v;
  }
}
// dX_enc (row-blocked [row/32][H/8][32][8] per slab) [L_b - 1][b][:] = dh_enc[b][:]; dX zeroed beforehand
__global__ void scatter_final_grad_rb_kernel(const float* __restrict__ dh, int dh_ld, const int* __restrict__ lens, int B, int Bp,
                                             int H, __nv_bfloat16* __restrict__ dX) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H) return;
  const int b = (int)(idx / H), j = (int)(idx - (long long)b * H);
  const long long o = (long long)(lens[b] - 1) * Bp * H + ((long long)(b >> 5) * (H / 8) + (j >> 3)) * 256 + (b & 31) * 8 + (j & 7);
  dX[o] = __float2bfloat16_rn(dh[(long long)b * dh_ld + j]);
}
bool varlen_gemm_enabled() {
  const char* e = getenv("MVAE_VARLEN_GEMM");
  return e ? atoi(e) != 0 : true;
}
bool train_cell_fused_enabled() {
  const char* e = getenv("MVAE_TRAIN_CELL_FUSED");
  return e ? atoi(e) != 0 : true;
}
// lens + hlast: capture each sequence's state after its last valid step (right-padded forward direction).
// lead_pad: reverse direction in processing order (padding first); `final32` then receives the state after the last step.
// WhhC / bhhC (bf16 mode, decoder layers): W_hh / b_hh in the fused-cell tile order -> the GRU cell runs in the epilogue of
// the step's GEMM (no gate pre-activations in HBM, no separate gate kernel).
template <typename TA>
int gru_fwd(const MDims& d, const MWS& w, cudaStream_t st, const TA* gi, const TA* Whh, const float* bhh, TA* hs, TA* sv,
            int H, const float* h0, const int* lens, float* hlast, bool lead_pad = false, float* final32 = nullptr,
            const void* WhhC = nullptr, const float* bhhC = nullptr, const int* act = nullptr) {
  // act (host, T entries, optional): rows [0, act[t]) are the sequences still running at step t (the batch is sorted by
  // length, so they form a prefix -- torch's packed-sequence batch sizes); the step's GEMM / cell work covers only them.
  const int Bp = d.Bp, T = d.T;
  const size_t slab = (size_t)Bp * H;
  if (act) RC(memset_async(hs, (size_t)(T + 1) * slab * sizeof(TA), st));   // rows past a sequence's end read as zeros
  if constexpr (sizeof(TA) == 2) {
    if (WhhC && !lens && !lead_pad && !final32 && train_cell_fused_enabled()) {
      if (h0) { init_h0_kernel<TA><<<ceil_div((int)slab, 256), 256, 0, st>>>(h0, d.B, Bp, H, hs, w.h32[0]); KCHECK(); }
      else { RC(memset_async(hs, slab * sizeof(TA), st)); RC(memset_async(w.h32[0], slab * 4, st)); }
      for (int t = 0; t < T; ++t) {
        const int M = act ? act[t] : Bp;
        if (M <= 0) break;
        mvae_umma_operand a{hs + t * slab, 0, M, H, H, 1, 0, 0, 0};
        mvae_umma_operand b{WhhC, 0, 3ll * H, H, H, 1, 0, 0, 0};
        mvae_umma_out o{w.gh, 3ll * H, 0, 0, bhhC, 0};
        mvae_umma_cell c{};
        c.gates = 3; c.H = H; c.gi = gi + (size_t)t * Bp * 3 * H; c.h_prev32 = w.h32[t & 1]; c.h_next32 = w.h32[(t + 1) & 1];
        c.out_a = hs + (t + 1) * slab; c.ld_a = H; c.out_b = nullptr; c.ld_b = 0; c.sv = sv + (size_t)t * Bp * 4 * H;
        mvae_count_launches(1);
        RC(mvae_umma_gemm(&a, &b, &o, M, 3 * H, H, 192, 1, 0, w.err_flag, st, nullptr, &c));
      }
      return MVAE_OK;
    }
  }
  if (h0) {
    init_h0_kernel<TA><<<ceil_div((int)slab, 256), 256, 0, st>>>(h0, d.B, Bp, H, hs, d.bf16 ? w.h32[0] : nullptr);
    KCHECK();
  } else {
    RC(memset_async(hs, slab * sizeof(TA), st));
    if (d.bf16) RC(memset_async(w.h32[0], slab * 4, st));
  }
  for (int t = 0; t < T; ++t) {
    const int M = act ? act[t] : Bp;
    if (M <= 0) break;
    RC(gemm<TA>(w.err_flag, st, hs + t * slab, H, false, Whh, H, true, w.gh, 3 * H, false, M, 3 * H, H, bhh, false, 1));
    simt::gru_gate_fwd_kernel<TA, TA><<<(unsigned)ceil_div64((long long)M * H, 256), 256, 0, st>>>(
        gi + (size_t)t * Bp * 3 * H, w.gh, d.bf16 ? w.h32[t & 1] : nullptr, hs + t * slab, hs + (t + 1) * slab,
        d.bf16 ? w.h32[(t + 1) & 1] : nullptr, sv + (size_t)t * Bp * 4 * H, M, H, lens, hlast, t, d.B, lead_pad ? T : 0);
    KCHECK();
  }
  if (final32) {
    final_state_kernel<TA><<<ceil_div((int)slab, 256), 256, 0, st>>>(hs + (size_t)T * slab, d.bf16 ? w.h32[T & 1] : nullptr, (long long)slab,
                                                                     final32);
    KCHECK();
  }
  return MVAE_OK;
}
// after the call: dG filled for all t; if dh0_acc != null, dh0_acc += dL/dh0
// carry_init ([B][carry_ld] fp32, optional): gradient wrt the state after the LAST processed step (reverse encoder direction)
template <typename TA>
int gru_bwd(const MDims& d, const MWS& w, cudaStream_t st, const TA* Whh, const TA* hs, const TA* sv, const TA* dX, TA* dG,
            int H, float* dh0_acc, const float* carry_init = nullptr, int carry_ld = 0, const int* act = nullptr) {
  const int Bp = d.Bp, T = d.T;
  const size_t slab = (size_t)Bp * H;
  RC(memset_async(w.dh_carry, slab * 4, st));
  if (carry_init) {
    copy_rows_kernel<<<grid_for((long long)d.B * H), 256, 0, st>>>(carry_init, carry_ld, d.B, H, w.dh_carry, H);
    KCHECK();
  }
  const int gate_grid = ceil_div((int)slab, 256);
  for (int t = T - 1; t >= 0; --t) {
    TA* dGt = dG + (size_t)t * Bp * 4 * H;
    const int M = act ? act[t] : Bp;   // sequences still running at step t; the other rows get zero gradients
    simt::gru_gate_bwd_kernel<TA><<<gate_grid, 256, 0, st>>>(sv + (size_t)t * Bp * 4 * H, hs + t * slab, dX + t * slab,
                                                             w.dh_carry, dGt, nullptr, Bp, H, M);
    KCHECK();
    if ((t > 0 || dh0_acc) && M > 0)
      RC(gemm<TA>(w.err_flag, st, dGt + H, 4 * H, false, Whh, H, false, w.dh_carry, H, false, M, H, 3 * H, nullptr, true, 1));
  }
  if (dh0_acc) {
    add_inplace_kernel<<<grid_for((long long)slab), 256, 0, st>>>(dh0_acc, w.dh_carry, (long long)slab);
    KCHECK();
  }
  return MVAE_OK;
}

// Property head on z (BindingModel, mosesvae.py:6-25) run between the VAE's forward and backward halves of ONE step: the
// historical VAE.forward(x, binding) -> (kl, recon, binding_loss, z) that moses_train_distrib.py:274 / trainbinding.py:216
// call.  binding_loss = weight * mean_b (BindingModel(z)_b - target_b)^2; its gradient wrt z joins the decoder's.
struct JointHead {
  const mvae_binding_desc* desc; const float* const* params; float* const* grads; float* const* running;
  const float* target; float weight; float* pred; float* dout; float* dz; float* loss_out; void* ws; size_t ws_bytes;
};
__global__ void mse_head_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, float weight,
                                float* __restrict__ dout, float* __restrict__ loss_out) {
  // one block: loss = weight * mean (pred - target)^2 ; dout = weight * 2 (pred - target) / B
  double local = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float e = pred[b] - target[b];
    local += (double)e * e;
    dout[b] = weight * 2.0f * e / (float)B;
  }
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    *loss_out = (float)(weight * s / B);
  }
}

// phase: -1 = the whole step; 0..L-1 = the part whose gradients become final in phase p (data-parallel bucket order, as
// mvae_cfgb_elbo_step_phase): 0 = forward + head + top decoder layer, k = decoder layer L-1-k, L-1 additionally the
// layer-0 input weights, decoder_lat, the encoder heads / GRU and x_emb.
// dz_ext (optional, [B][Z]): gradient wrt z arriving from outside (autograd of a consumer of z); jh (optional): property head.
template <typename TA>
int step_t(const MDims& d, const MWS& w, const float* const* P, float* const* G, const uint8_t* ids,
           const int* lens, const float* eps, float* out_scalars, float* z_out, float* lv_out, float* y_out,
           bool backward, cudaStream_t st, const int* act, int phase = -1, const float* dz_ext = nullptr,
           const JointHead* jh = nullptr) {
  const int B = d.B, Bp = d.Bp, T = d.T, V = d.V, CP = d.CP, Z = d.Z, Hq = d.Hq, Hd = d.Hd, L = d.L, ML = d.MLP;
  const int TB = T * Bp;
  const int IN0 = V + Z;   // decoder layer-0 input width
  const int Hin = d.Hin;
  const MP ix{d.bidir, d.lin, L};
  if (phase >= L) return MVAE_ERR_INVALID;
  // packed sequences: the big time-major GEMMs skip output tiles (mode 1) / k-blocks (mode 2) past the running sequences
  mvae_umma_varlen vlm{w.act_dev, Bp, 1, nullptr}, vlk{w.act_dev, Bp, 2, balanced_splits_enabled() ? act : nullptr};
  const mvae_umma_varlen* VLM = (act && d.bf16 && varlen_gemm_enabled()) ? &vlm : nullptr;
  const mvae_umma_varlen* VLK = (act && d.bf16 && varlen_gemm_enabled()) ? &vlk : nullptr;
  const bool prec_enc = sizeof(TA) == 2 && persistent_encoder(d);
  const bool prec = sizeof(TA) == 2 && persistent_decoder(d);
  const bool drop = d.drop > 0.f;
  if (phase <= 0) {
  // ---- control + weight preparation
  RC(memset_async(w.err_flag, 4, st)); RC(memset_async(w.kl_sum, 8, st)); RC(memset_async(w.nll_sum, 8, st));
  RC(memset_async(w.M, 4, st));
  count_targets_kernel<<<1, 256, 0, st>>>(lens, B, w.M); KCHECK();
  if (VLM) { count_active_kernel<<<ceil_div(max(T, Bp / 256), 128), 128, 0, st>>>(lens, B, T, w.act_dev, w.act256_dev, w.tileT, Bp); KCHECK(); }
  }
  // persistent sweeps over packed sequences: row tile j runs only tileT[j] steps; the GEMMs that feed / follow them skip the
  // same (t, 256-row tile) regions, which are then neither written nor read
  mvae_umma_varlen vlm256{w.act256_dev, Bp, 1, nullptr};
  const mvae_umma_varlen* VLS = VLM ? &vlm256 : nullptr;
  const int* tileT = VLM ? w.tileT : nullptr;
  const int* lim256 = VLM ? w.act256_dev : nullptr;
  if (phase <= 0) {
  simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hq * Hq), 256, 0, st>>>(P[ix.e_whh(0)], Hq, Hq, (TA*)w.Whh_enc, Hq, Hq, 0, 1, 2); KCHECK();
  simt::pad_gate_vector_kernel<<<ceil_div(3 * Hq, 256), 256, 0, st>>>(P[ix.e_bhh(0)], Hq, w.bhh_enc, Hq); KCHECK();
  if (d.bidir) {
    simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hq * Hq), 256, 0, st>>>(P[ix.e_whh(1)], Hq, Hq, (TA*)w.Whh_encr, Hq, Hq, 0, 1, 2); KCHECK();
    simt::pad_gate_vector_kernel<<<ceil_div(3 * Hq, 256), 256, 0, st>>>(P[ix.e_bhh(1)], Hq, w.bhh_encr, Hq); KCHECK();
  }
  for (int l = 0; l < L; ++l) {
    simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(P[ix.whh(l)], Hd, Hd, (TA*)w.Whh[l], Hd, Hd, 0, 1, 2); KCHECK();
    simt::pad_gate_vector_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(P[ix.bhh(l)], Hd, w.bhh[l], Hd); KCHECK();
    if (l >= 1) {
      simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(P[ix.wih(l)], Hd, Hd, (TA*)w.Wih[l], Hd, Hd, 0, 1, 2); KCHECK();
      simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(P[ix.wih(l)], Hd, Hd, (TA*)w.Wih_nrz[l], Hd, Hd, 2, 0, 1); KCHECK();
      simt::pad_gate_vector_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(P[ix.bih(l)], Hd, w.bih[l], Hd); KCHECK();
    }
  }
  simt::pad_matrix_kernel<TA><<<grid_for((long long)CP * Hd), 256, 0, st>>>(P[ix.fcw()], V, Hd, (TA*)w.Wfc, CP, Hd); KCHECK();
  simt::pad_matrix_kernel<float><<<1, 256, 0, st>>>(P[ix.fcb()], 1, V, w.bfc, 1, CP); KCHECK();

  // ---- encoder: table look-up projection, GRU, final state, MLP heads, reparametrise + KL
  RC(sg(st, P[ix.emb()], V, 1, P[ix.e_wih(0)], 1, V, w.TBLe, 3 * Hq, V, 3 * Hq, V, P[ix.e_bih(0)], simt::ACT_NONE, 0));
  RC(memset_async(w.hlast, (size_t)Bp * Hq * 4, st));
  if (prec_enc || prec) {
    transpose_ids_kernel<<<grid_for((long long)TB), 256, 0, st>>>(ids, T, B, Bp, T, w.tokT); KCHECK();
  }
  bool enc_swept = false;
  if constexpr (sizeof(TA) == 2) {
    if (prec_enc) {
      // forward direction of the encoder GRU as one persistent launch: x_t W_ih^T + b_ih is a row of the V x 3Hq table
      // (x_emb . W_ih^T), looked up by token inside the kernel; each row's state after its last token goes to hlast
      combine_gate_bias_kernel<<<ceil_div(3 * Hq, 256), 256, 0, st>>>(nullptr, w.bhh_enc, w.bcomb[0], Hq); KCHECK();
      table_add_bias_kernel<<<ceil_div(V * 3 * Hq, 256), 256, 0, st>>>(w.TBLe, w.bcomb[0], V, 3 * Hq, w.tbl_comb); KCHECK();
      RC(memset_async(w.hs_enc, (size_t)Bp * Hq * sizeof(TA), st));
      mvae_gru_rec_args ra{};
      ra.backward = 0; ra.variant = 32; ra.Bp = Bp; ra.Hp = Hq; ra.T = T;
      ra.W = (const __nv_bfloat16*)w.Whh_enc; ra.gi = nullptr; ra.gi_tstride = 0; ra.bhh = w.bhh_enc + 2 * Hq;
      ra.hs = (__nv_bfloat16*)w.hs_enc; ra.sv = (__nv_bfloat16*)w.sv_enc; ra.counters = w.counters; ra.err_flag = w.err_flag;
      ra.ones_col = -1; ra.lens = lens; ra.hlast = w.hlast; ra.nrows = B;
      if (V <= 64) { ra.tbl = w.tbl_comb; ra.tok = w.tokT; ra.V = V; }
      else {   // large vocabulary: the projection is materialised (table gather) instead of looked up inside the sweep
        gather_rb_kernel<<<grid_for((long long)TB * (3 * Hq / 8)), 256, 0, st>>>(w.tbl_comb, w.tokT, nullptr, T, Bp, 3 * Hq,
                                                                               (__nv_bfloat16*)w.gi, lim256); KCHECK();
        ra.gi = (const __nv_bfloat16*)w.gi; ra.gi_tstride = (long long)Bp * 3 * Hq;
      }
      ra.tile_T = tileT;
      mvae_count_launches(2);
      RC(mvae_gru_rec2_launch(&ra, 1, st));
      enc_swept = true;
    }
  }
  if (!enc_swept) {
    gather_rows_kernel<TA><<<grid_for((long long)TB * 3 * Hq), 256, 0, st>>>(w.TBLe, 3 * Hq, ids, T, nullptr, B, Bp, T, (TA*)w.gi); KCHECK();
    RC(gru_fwd<TA>(d, w, st, (const TA*)w.gi, (const TA*)w.Whh_enc, w.bhh_enc, (TA*)w.hs_enc, (TA*)w.sv_enc, Hq, nullptr, lens, w.hlast,
                   false, nullptr, nullptr, nullptr, act));
  }
  copy_rows_kernel<<<grid_for((long long)B * Hq), 256, 0, st>>>(w.hlast, Hq, B, Hq, w.hcat, Hin); KCHECK();
  if (d.bidir) {
    // reverse direction (mosesfile.py:21-28,112-116): same engine over the time-reversed token stream, padding first
    RC(sg(st, P[ix.emb()], V, 1, P[ix.e_wih(1)], 1, V, w.TBLer, 3 * Hq, V, 3 * Hq, V, P[ix.e_bih(1)], simt::ACT_NONE, 0));
    bool rev_swept = false;
    if constexpr (sizeof(TA) == 2) {
      if (prec_enc) {
        // the reverse direction as a FORWARD sweep of the persistent kernel over every row reversed inside its own length:
        // the padding still comes last, the direction's final state (after the row's first token) is the state after step
        // lens[b]-1, and the per-tile step windows of the packed batch apply unchanged
        transpose_ids_kernel<<<grid_for((long long)TB), 256, 0, st>>>(ids, T, B, Bp, T, w.tokTr, lens); KCHECK();
        combine_gate_bias_kernel<<<ceil_div(3 * Hq, 256), 256, 0, st>>>(nullptr, w.bhh_encr, w.bcomb[0], Hq); KCHECK();
        table_add_bias_kernel<<<ceil_div(V * 3 * Hq, 256), 256, 0, st>>>(w.TBLer, w.bcomb[0], V, 3 * Hq, w.tbl_comb); KCHECK();
        RC(memset_async(w.hs_encr, (size_t)Bp * Hq * sizeof(TA), st));
        RC(memset_async(w.hlastr, (size_t)Bp * Hq * 4, st));
        mvae_gru_rec_args ra{};
        ra.backward = 0; ra.variant = 32; ra.Bp = Bp; ra.Hp = Hq; ra.T = T;
        ra.W = (const __nv_bfloat16*)w.Whh_encr; ra.gi = nullptr; ra.gi_tstride = 0; ra.bhh = w.bhh_encr + 2 * Hq;
        ra.hs = (__nv_bfloat16*)w.hs_encr; ra.sv = (__nv_bfloat16*)w.sv_encr; ra.counters = w.counters; ra.err_flag = w.err_flag;
        ra.ones_col = -1; ra.lens = lens; ra.hlast = w.hlastr; ra.nrows = B;
        if (V <= 64) { ra.tbl = w.tbl_comb; ra.tok = w.tokTr; ra.V = V; }
        else {
          gather_rb_kernel<<<grid_for((long long)TB * (3 * Hq / 8)), 256, 0, st>>>(w.tbl_comb, w.tokTr, nullptr, T, Bp, 3 * Hq,
                                                                                 (__nv_bfloat16*)w.gi, lim256); KCHECK();
          ra.gi = (const __nv_bfloat16*)w.gi; ra.gi_tstride = (long long)Bp * 3 * Hq;
        }
        ra.tile_T = tileT;
        mvae_count_launches(2);
        RC(mvae_gru_rec2_launch(&ra, 1, st));
        rev_swept = true;
      }
    }
    if (!rev_swept) {
      gather_rows_kernel<TA><<<grid_for((long long)TB * 3 * Hq), 256, 0, st>>>(w.TBLer, 3 * Hq, ids, T, nullptr, B, Bp, T, (TA*)w.gi, 1); KCHECK();
      RC(gru_fwd<TA>(d, w, st, (const TA*)w.gi, (const TA*)w.Whh_encr, w.bhh_encr, (TA*)w.hs_encr, (TA*)w.sv_encr, Hq, nullptr, lens,
                     nullptr, true, w.hlastr));
    }
    copy_rows_kernel<<<grid_for((long long)B * Hq), 256, 0, st>>>(w.hlastr, Hq, B, Hq, w.hcat + Hq, Hin); KCHECK();
  }
  if (d.lin) {   // mosesfile.py:31-32,118: single Linear heads on cat(h_fwd, h_bwd)
    RC(sg(st, w.hcat, Hin, 1, P[ix.head_w0(0)], 1, Hin, w.mu, Z, B, Z, Hin, P[ix.head_w0(0) + 1], simt::ACT_NONE, 0));
    RC(sg(st, w.hcat, Hin, 1, P[ix.head_w0(1)], 1, Hin, w.lv, Z, B, Z, Hin, P[ix.head_w0(1) + 1], simt::ACT_NONE, 0));
  } else {       // mosesvae.py:68-69: Linear -> ReLU -> Linear
    RC(sg(st, w.hcat, Hin, 1, P[ix.head_w0(0)], 1, Hin, w.rmu, ML, B, ML, Hin, P[ix.head_w0(0) + 1], simt::ACT_RELU, 0));
    RC(sg(st, w.rmu, ML, 1, P[ix.head_w0(0) + 2], 1, ML, w.mu, Z, B, Z, ML, P[ix.head_w0(0) + 3], simt::ACT_NONE, 0));
    RC(sg(st, w.hcat, Hin, 1, P[ix.head_w0(1)], 1, Hin, w.rlv, ML, B, ML, Hin, P[ix.head_w0(1) + 1], simt::ACT_RELU, 0));
    RC(sg(st, w.rlv, ML, 1, P[ix.head_w0(1) + 2], 1, ML, w.lv, Z, B, Z, ML, P[ix.head_w0(1) + 3], simt::ACT_NONE, 0));
  }
  reparam_kl_std_kernel<<<grid_for((long long)B * Z, 256, 592), 256, 0, st>>>(w.mu, w.lv, eps, (long long)B * Z, w.z, w.kl_sum); KCHECK();

  // ---- decoder: h0, layer-0 projection = table + per-molecule z part, GRU stack, head
  RC(memset_async(w.h0, (size_t)Bp * Hd * 4, st));
  RC(sg(st, w.z, Z, 1, P[ix.latw()], 1, Z, w.h0, Hd, B, Hd, Z, P[ix.latb()], simt::ACT_NONE, 0));
  RC(sg(st, P[ix.emb()], V, 1, P[ix.wih(0)], 1, IN0, w.TBLd, 3 * Hd, V, 3 * Hd, V, nullptr, simt::ACT_NONE, 0));
  RC(sg(st, w.z, Z, 1, P[ix.wih(0)] + V, 1, IN0, w.zproj, 3 * Hd, B, 3 * Hd, Z, P[ix.bih(0)], simt::ACT_NONE, 0));
  if (prec) {
    // layer 0: the token part of the projection is looked up inside the kernel (table in shared memory), the per-molecule
    // z part (+ b_ih + (b_hr, b_hz, 0)) is one time-invariant row-blocked slab
    combine_gate_bias_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(nullptr, w.bhh[0], w.bcomb[0], Hd); KCHECK();
    rows_to_rb_kernel<<<grid_for((long long)Bp * 3 * Hd), 256, 0, st>>>(w.zproj, w.bcomb[0], B, Bp, 3 * Hd, (__nv_bfloat16*)w.zproj_rb); KCHECK();
  } else {
    gather_rows_kernel<TA><<<grid_for((long long)TB * 3 * Hd), 256, 0, st>>>(w.TBLd, 3 * Hd, ids, T, w.zproj, B, Bp, T, (TA*)w.gi); KCHECK();
  }
  for (int l = 0; l < L; ++l) {
    if (l >= 1) {
      const TA* X = (const TA*)w.hs[l - 1] + (size_t)Bp * Hd;
      if (drop) {   // nn.GRU(dropout=p) in train mode: dropout on the outputs of every layer but the last (mosesvae.py:78)
        RC(dropout_launch<TA>(st, X, (TA*)w.hdrop[l], (unsigned long long)d.drop_seed + (l - 1), d.drop, B, Bp, Hd, T));
        X = (const TA*)w.hdrop[l];
      }
      if (prec) {
        // every row of every slab is written (no tile skipping): the persistent sweep runs all rows through all T steps, and
        // what it computes past a sequence's end must stay finite (it meets zero gradients in the K = T*B wgrad GEMMs)
        combine_gate_bias_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(w.bih[l], w.bhh[l], w.bcomb[l], Hd); KCHECK();
        RC(gemm<TA>(w.err_flag, st, X, Hd, false, (const TA*)w.Wih[l], Hd, true, w.gi, 3 * Hd, true, TB, 3 * Hd, Hd, w.bcomb[l], false, 1, 0,
                    VLS, true));
      } else {
        RC(gemm<TA>(w.err_flag, st, X, Hd, false, (const TA*)w.Wih[l], Hd, true, w.gi, 3 * Hd, true, TB, 3 * Hd, Hd, w.bih[l], false, 1, 0, VLM));
      }
    }
    if constexpr (sizeof(TA) == 2) {
      if (prec) {
        // one launch = all T steps of this layer: h0 = decoder_lat(z) (mosesvae.py:180-181), saved gates in the fragment layout
        init_h0_kernel<TA><<<ceil_div(Bp * Hd, 256), 256, 0, st>>>(w.h0, B, Bp, Hd, (TA*)w.hs[l], nullptr); KCHECK();
        mvae_gru_rec_args ra{};
        ra.backward = 0; ra.variant = 32; ra.Bp = Bp; ra.Hp = Hd; ra.T = T;
        ra.W = (const __nv_bfloat16*)w.Whh[l]; ra.gi = (const __nv_bfloat16*)w.gi; ra.gi_tstride = (long long)Bp * 3 * Hd;
        if (l == 0 && V <= 64) { ra.gi = (const __nv_bfloat16*)w.zproj_rb; ra.gi_tstride = 0; ra.tbl = w.TBLd; ra.tok = w.tokT; ra.V = V; }
        else if (l == 0) {   // large vocabulary: token part + per-molecule z part materialised once for all steps
          gather_rb_kernel<<<grid_for((long long)TB * (3 * Hd / 8)), 256, 0, st>>>(w.TBLd, w.tokT, (const __nv_bfloat16*)w.zproj_rb, T, Bp,
                                                                                 3 * Hd, (__nv_bfloat16*)w.gi, lim256); KCHECK();
        }
        ra.bhh = w.bhh[l] + 2 * Hd; ra.hs = (__nv_bfloat16*)w.hs[l]; ra.sv = (__nv_bfloat16*)w.sv[l]; ra.counters = w.counters;
        ra.err_flag = w.err_flag; ra.ones_col = -1; ra.h0 = w.h0; ra.tile_T = tileT;
        mvae_count_launches(2);
        RC(mvae_gru_rec2_launch(&ra, 1, st));
        continue;
      }
    }
    if (d.bf16) {
      cell_weights_kernel<<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(nullptr, P[ix.whh(l)], nullptr, P[ix.bhh(l)], Hd, 3,
                                                                   (__nv_bfloat16*)w.WhhC[l], w.bhhC[l]); KCHECK();
    }
    RC(gru_fwd<TA>(d, w, st, (const TA*)w.gi, (const TA*)w.Whh[l], w.bhh[l], (TA*)w.hs[l], (TA*)w.sv[l], Hd, w.h0, nullptr, nullptr,
                   false, nullptr, d.bf16 ? w.WhhC[l] : nullptr, w.bhhC[l], act));
  }
  bool ce_fused = false;
  if constexpr (sizeof(TA) == 2) {
    // fused CE head (mosesvae.py:190-197): logits -> log-softmax -> NLL -> d(logits) inside the epilogue of the vocabulary GEMM;
    // the (T*B, V) logits never reach HBM.  Needs the vocabulary in one 64-wide tile and nobody asking for the logits (y_out);
    // MVAE_FUSED_HEAD=0 keeps the two-kernel form (the cross-check).  Row tiles past every running sequence are skipped by the
    // GEMM, so their d(logits) rows are zeroed first (the bias sum below reads every row).
    if (CP == 64 && !y_out && backward && moses_fused_head_enabled()) {
      if (VLM) RC(memset_async(w.dlogits, (size_t)TB * CP * sizeof(TA), st));
      mvae_umma_operand a{(const TA*)w.hs[L - 1] + (size_t)Bp * Hd, 0, (long long)TB, Hd, Hd, 1, 0, 0, 0};
      mvae_umma_operand b{w.Wfc, 0, CP, Hd, Hd, 1, 0, 0, 0};
      mvae_umma_out o{w.logits, CP, 0, 0, w.bfc, 0};
      mvae_umma_head h{ids, B, Bp, T, V, 0.f, w.dlogits, w.nll_sum, nullptr, 1, lens, w.M, d.rec_w};
      mvae_count_launches(1);
      RC(mvae_umma_gemm(&a, &b, &o, TB, CP, Hd, 64, 1, 0, w.err_flag, st, &h, nullptr, nullptr, VLM));
      ce_fused = true;
    }
  }
  if (!ce_fused) {
  RC(gemm<TA>(w.err_flag, st, (const TA*)w.hs[L - 1] + (size_t)Bp * Hd, Hd, false, (const TA*)w.Wfc, Hd, true, w.logits, CP, false,
              TB, CP, Hd, w.bfc, false, 1, 64, VLM));
  head_ce_kernel<TA><<<(unsigned)ceil_div64((long long)TB * 32, 256), 256, 0, st>>>(
      w.logits, CP, V, ids, T, lens, B, Bp, T, w.M, d.rec_w, w.bfc, backward ? (TA*)w.dlogits : nullptr, y_out, w.nll_sum);
  KCHECK();
  }
  finalize_kernel<<<1, 1, 0, st>>>(w.kl_sum, w.nll_sum, w.M, B, d.kl_w, d.rec_w, out_scalars); KCHECK();
  if (z_out) { mvae_count_launches(1); MVAE_CUDA_CHECK(cudaMemcpyAsync(z_out, w.z, (size_t)B * Z * 4, cudaMemcpyDeviceToDevice, st)); }
  if (lv_out) { mvae_count_launches(1); MVAE_CUDA_CHECK(cudaMemcpyAsync(lv_out, w.lv, (size_t)B * Z * 4, cudaMemcpyDeviceToDevice, st)); }
  if (jh) {
    // property head between the two halves of the step: prediction, MSE against the target, its backward (parameter
    // gradients + dL/dz, which joins the decoder's gradient wrt z below)
    RC(mvae_binding_forward(jh->desc, jh->params, jh->running, w.z, jh->pred, jh->ws, jh->ws_bytes, reinterpret_cast<mvae_stream_t>(st)));
    mse_head_kernel<<<1, 256, 0, st>>>(jh->pred, jh->target, B, jh->weight, jh->dout, jh->loss_out); KCHECK();
    if (backward)
      RC(mvae_binding_backward(jh->desc, jh->params, jh->grads, w.z, jh->dout, jh->dz, jh->ws, jh->ws_bytes, reinterpret_cast<mvae_stream_t>(st)));
    mvae_count_launches(24);
  }
  if (!backward) { simt::nan_if_error_kernel<<<1, 1, 0, st>>>(w.err_flag, out_scalars); KCHECK(); return MVAE_OK; }
  }   // phase <= 0: forward half

  // =================================== backward ===================================
  const TA* dlog = (const TA*)w.dlogits;
  const int wsplits = d.bf16 ? 12 : 64;
  if (phase <= 0) {
  onehot_rows_kernel<TA><<<grid_for((long long)TB * CP), 256, 0, st>>>(ids, T, B, Bp, T, CP, (TA*)w.OH); KCHECK();
  // head
  // persistent BPTT: dX in the row-blocked layout, every row written (zeros past a sequence's end, where dlogits is zero)
  RC(gemm<TA>(w.err_flag, st, dlog, CP, false, (const TA*)w.Wfc, Hd, false, w.dX, Hd, true, TB, Hd, CP, nullptr, false, 1, 0,
              prec ? VLS : VLM, prec));
  RC(memset_async(w.dWfc_p, (size_t)CP * Hd * 4, st));
  RC(gemm<TA>(w.err_flag, st, dlog, CP, true, (const TA*)w.hs[L - 1] + (size_t)Bp * Hd, Hd, false, w.dWfc_p, Hd, false, CP, Hd, TB,
              nullptr, true, d.bf16 ? 148 : 64, 256, VLK));
  simt::unpad_matrix_kernel<<<grid_for((long long)V * Hd), 256, 0, st>>>(w.dWfc_p, Hd, G[ix.fcw()], V, Hd); KCHECK();
  RC(memset_async(w.csum, (size_t)4 * Hd * 4, st));
  RC(simt::colsum<TA>(st, dlog, TB, CP, CP, w.csum)); mvae_count_launches(1);
  simt::pad_matrix_kernel<float><<<1, 256, 0, st>>>(w.csum, 1, V, G[ix.fcb()], 1, V); KCHECK();
  RC(memset_async(w.dh0, (size_t)Bp * Hd * 4, st));
  }   // phase <= 0: head
  // decoder GRU stack
  for (int l = L - 1; l >= 0; --l) {
    if (phase >= 0 && l != L - 1 - phase) continue;
    const TA* hs = (const TA*)w.hs[l];
    TA* dG = (TA*)w.dG;
    bool swept = false;
    if constexpr (sizeof(TA) == 2) {
      if (prec) {
        whh_transpose_kernel<<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(P[ix.whh(l)], Hd, (__nv_bfloat16*)w.WhhT[l]); KCHECK();
        mvae_gru_rec_args ra{};
        ra.backward = 1; ra.variant = 32; ra.Bp = Bp; ra.Hp = Hd; ra.T = T;
        ra.W = (const __nv_bfloat16*)w.WhhT[l]; ra.hs = (__nv_bfloat16*)w.hs[l]; ra.sv = (__nv_bfloat16*)w.sv[l];
        ra.dX = (const __nv_bfloat16*)w.dX; ra.dG = (__nv_bfloat16*)dG; ra.counters = w.counters; ra.err_flag = w.err_flag;
        ra.carry_out = w.dh_carry; ra.tile_T = tileT;
        mvae_count_launches(2);
        RC(mvae_gru_rec2_launch(&ra, 0, st));
        // dL/dh0 of this layer = dh_0 * z_0 (carry_out) + dgh_0 * W_hh
        RC(gemm<TA>(w.err_flag, st, dG + Hd, 4 * Hd, false, (const TA*)w.Whh[l], Hd, false, w.dh_carry, Hd, false, Bp, Hd, 3 * Hd, nullptr,
                    true, 1));
        add_inplace_kernel<<<grid_for((long long)Bp * Hd), 256, 0, st>>>(w.dh0, w.dh_carry, (long long)Bp * Hd); KCHECK();
        swept = true;
      }
    }
    if (!swept)
      RC(gru_bwd<TA>(d, w, st, (const TA*)w.Whh[l], hs, (const TA*)w.sv[l], (const TA*)w.dX, dG, Hd, w.dh0, nullptr, 0, act));
    RC(memset_async(w.dW_p, (size_t)3 * Hd * Hd * 4, st));
    RC(gemm<TA>(w.err_flag, st, dG + Hd, 4 * Hd, true, hs, Hd, false, w.dW_p, Hd, false, 3 * Hd, Hd, TB, nullptr, true, wsplits, 256, VLK));
    simt::unpad_gate_matrix_kernel<<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(w.dW_p, Hd, Hd, G[ix.whh(l)], Hd, Hd, 0, 1, 2); KCHECK();
    RC(memset_async(w.csum, (size_t)4 * Hd * 4, st));
    RC(simt::colsum<TA>(st, dG, TB, 4 * Hd, 4 * Hd, w.csum, swept ? lim256 : nullptr, Bp)); mvae_count_launches(1);
    gate_bias_grads_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(w.csum, Hd, G[ix.bih(l)], G[ix.bhh(l)]); KCHECK();
    if (l >= 1) {
      const TA* X = (const TA*)w.hs[l - 1] + (size_t)Bp * Hd;
      if (drop) X = (const TA*)w.hdrop[l];   // the dropped input of this layer, kept by the forward pass
      RC(memset_async(w.dW_p, (size_t)3 * Hd * Hd * 4, st));
      RC(gemm<TA>(w.err_flag, st, dG, 4 * Hd, true, X, Hd, false, w.dW_p, Hd, false, 3 * Hd, Hd, TB, nullptr, true, wsplits, 256, VLK));
      simt::unpad_gate_matrix_kernel<<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(w.dW_p, Hd, Hd, G[ix.wih(l)], Hd, Hd, 2, 0, 1); KCHECK();
      RC(gemm<TA>(w.err_flag, st, dG, 4 * Hd, false, (const TA*)w.Wih_nrz[l], Hd, false, w.dX, Hd, true, TB, Hd, 3 * Hd, nullptr, false, 1, 0,
                  prec ? VLS : VLM, prec));
      if (drop) {   // gradient wrt the undropped outputs of layer l-1
        RC(dropout_launch<TA>(st, (const TA*)w.dX, (TA*)w.dX, (unsigned long long)d.drop_seed + (l - 1), d.drop, B, Bp, Hd, T, prec ? 1 : 0));
      }
    } else {
      // layer 0: table gradient (tensor-core GEMM onehot^T * dgi) and the per-molecule z part (time sum)
      RC(memset_async(w.dTBL, (size_t)CP * 3 * Hd * 4, st));
      RC(gemm<TA>(w.err_flag, st, (const TA*)w.OH, CP, true, dG, 4 * Hd, false, w.dTBL, 3 * Hd, false, CP, 3 * Hd, TB, nullptr, true,
                  d.bf16 ? 24 : 64, 256, VLK));
      if constexpr (sizeof(TA) == 2) {
        dgi_time_sum_bf16x8_kernel<<<(unsigned)ceil_div64((long long)Bp * 3 * Hd / 8, 256), 256, 0, st>>>((const __nv_bfloat16*)dG, T, Bp, Hd, w.dgisum,
                                                                                                             swept ? tileT : nullptr); KCHECK();
      } else {
        dgi_time_sum_t_kernel<TA><<<(unsigned)ceil_div64((long long)Bp * 3 * Hd, 256), 256, 0, st>>>(dG, T, Bp, Hd, w.dgisum); KCHECK();
      }
    }
  }
  if (phase >= 0 && phase != L - 1) return MVAE_OK;
  // decoder layer-0 input weights: W_ih[:, :V] through the table, W_ih[:, V:] through z
  tbl_grad_to_rzn_T_kernel<<<(unsigned)ceil_div64(3ll * Hd * V, 256), 256, 0, st>>>(w.dTBL, Hd, V, w.dWT); KCHECK();
  //   dW_ih[:, :V] = dTBL_rzn^T(3Hd x V) * E (V x V)
  RC(sg(st, w.dWT, V, 1, P[ix.emb()], V, 1, G[ix.wih(0)], IN0, 3 * Hd, V, V, nullptr, simt::ACT_NONE, 0));
  //   dE = dTBL_rzn (V x 3Hd) * W_ih[:, :V] (3Hd x V)
  //   (a V x V output with K = 3Hd: split-K over 24 blocks instead of one block walking the whole contraction)
  RC(memset_async(G[ix.emb()], (size_t)V * V * 4, st));
  RC(sg(st, w.dWT, 1, V, P[ix.wih(0)], IN0, 1, G[ix.emb()], V, V, V, 3 * Hd, nullptr, simt::ACT_NONE, 1, 24));
  //   dW_ih[:, V:] = dgisum^T * z ; dz = dgisum * W_ih[:, V:]
  RC(sg_wgrad(st, w.dgisum, 1, 3 * Hd, w.z, Z, 1, G[ix.wih(0)] + V, IN0, 3 * Hd, Z, B));
  //   (skinny output, long contraction: split-K so that more than 3 x B/64 blocks share the work)
  RC(memset_async(w.dz, (size_t)B * Z * 4, st));
  RC(sg(st, w.dgisum, 3 * Hd, 1, P[ix.wih(0)] + V, IN0, 1, w.dz, Z, B, Z, 3 * Hd, nullptr, simt::ACT_NONE, 1, 4));
  // decoder_lat: h0 = z W^T + b, shared by all layers (dh0 already summed over layers)
  RC(sg_wgrad(st, w.dh0, 1, Hd, w.z, Z, 1, G[ix.latw()], Z, Hd, Z, B));
  RC(memset_async(G[ix.latb()], (size_t)Hd * 4, st));
  RC(simt::colsum<float>(st, w.dh0, B, Hd, Hd, G[ix.latb()])); mvae_count_launches(1);
  RC(sg(st, w.dh0, Hd, 1, P[ix.latw()], Z, 1, w.dz, Z, B, Z, Hd, nullptr, simt::ACT_NONE, 1, 4));
  // gradient wrt z from consumers of z outside the VAE (autograd) / from the property head
  if (dz_ext) { add_inplace_kernel<<<grid_for((long long)B * Z), 256, 0, st>>>(w.dz, dz_ext, (long long)B * Z); KCHECK(); }
  if (jh) { add_inplace_kernel<<<grid_for((long long)B * Z), 256, 0, st>>>(w.dz, jh->dz, (long long)B * Z); KCHECK(); }
  // reparametrisation + KL
  const long long nBZ = (long long)B * Z;
  reparam_kl_std_bwd_kernel<<<grid_for(nBZ), 256, 0, st>>>(w.mu, w.lv, eps, w.dz, d.kl_w / (float)B, nBZ, w.dmu, w.dlv); KCHECK();
  // heads (mu then logvar), accumulating into dhenc = dL/d cat(h_fwd, h_bwd)
  for (int head = 0; head < 2; ++head) {
    const float* dout = head == 0 ? w.dmu : w.dlv;
    const int W0 = ix.head_w0(head);
    if (d.lin) {
      RC(sg_wgrad(st, dout, 1, Z, w.hcat, Hin, 1, G[W0], Hin, Z, Hin, B));
      RC(memset_async(G[W0 + 1], (size_t)Z * 4, st));
      RC(simt::colsum<float>(st, dout, B, Z, Z, G[W0 + 1])); mvae_count_launches(1);
      RC(sg(st, dout, Z, 1, P[W0], Hin, 1, w.dhenc, Hin, B, Hin, Z, nullptr, simt::ACT_NONE, head));
      continue;
    }
    const float* r = head == 0 ? w.rmu : w.rlv;
    const int B0 = W0 + 1, W2 = W0 + 2, B2 = W0 + 3;
    RC(sg_wgrad(st, dout, 1, Z, r, ML, 1, G[W2], ML, Z, ML, B));
    RC(memset_async(G[B2], (size_t)Z * 4, st));
    RC(simt::colsum<float>(st, dout, B, Z, Z, G[B2])); mvae_count_launches(1);
    RC(sg(st, dout, Z, 1, P[W2], ML, 1, w.dr, ML, B, ML, Z, nullptr, simt::ACT_NONE, 0));
    relu_bwd_kernel<<<grid_for((long long)B * ML), 256, 0, st>>>(r, w.dr, (long long)B * ML); KCHECK();
    RC(sg_wgrad(st, w.dr, 1, ML, w.hcat, Hin, 1, G[W0], Hin, ML, Hin, B));
    RC(memset_async(G[B0], (size_t)ML * 4, st));
    RC(simt::colsum<float>(st, w.dr, B, ML, ML, G[B0])); mvae_count_launches(1);
    RC(sg(st, w.dr, ML, 1, P[W0], Hin, 1, w.dhenc, Hin, B, Hin, ML, nullptr, simt::ACT_NONE, head));
  }
  // encoder GRU(s): the gradient enters only at each direction's final state
  for (int rev = 0; rev <= d.bidir; ++rev) {
    TA* dXe = (TA*)w.dX;
    TA* dG = (TA*)w.dG;
    const TA* Whh = (const TA*)(rev ? w.Whh_encr : w.Whh_enc);
    const TA* hs = (const TA*)(rev ? w.hs_encr : w.hs_enc);
    const TA* sv = (const TA*)(rev ? w.sv_encr : w.sv_enc);
    RC(memset_async(dXe, (size_t)TB * Hq * sizeof(TA), st));
    bool swept = false;
    if constexpr (sizeof(TA) == 2) {
      if (prec_enc) {
        // both directions: the gradient enters at the row's last processed step (the reverse direction runs over the row
        // reversed inside its own length, see the forward pass)
        scatter_final_grad_rb_kernel<<<(unsigned)ceil_div64((long long)B * Hq, 256), 256, 0, st>>>(w.dhenc + (size_t)rev * Hq, Hin, lens, B,
                                                                                                 Bp, Hq, (__nv_bfloat16*)dXe); KCHECK();
        whh_transpose_kernel<<<grid_for(3ll * Hq * Hq), 256, 0, st>>>(P[ix.e_whh(rev)], Hq, (__nv_bfloat16*)w.WhhT_enc); KCHECK();
        mvae_gru_rec_args ra{};
        ra.backward = 1; ra.variant = 32; ra.Bp = Bp; ra.Hp = Hq; ra.T = T;
        ra.W = (const __nv_bfloat16*)w.WhhT_enc; ra.hs = (__nv_bfloat16*)(rev ? w.hs_encr : w.hs_enc);
        ra.sv = (__nv_bfloat16*)(rev ? w.sv_encr : w.sv_enc);
        ra.dX = (const __nv_bfloat16*)dXe; ra.dG = (__nv_bfloat16*)dG; ra.counters = w.counters; ra.err_flag = w.err_flag;
        ra.tile_T = tileT;
        mvae_count_launches(2);
        RC(mvae_gru_rec2_launch(&ra, 0, st));
        swept = true;
      }
    }
    if (swept) {
    } else if (!rev) {
      scatter_final_grad_kernel<TA><<<(unsigned)ceil_div64((long long)B * Hq, 256), 256, 0, st>>>(w.dhenc, Hin, lens, B, Bp, Hq, dXe); KCHECK();
      RC(gru_bwd<TA>(d, w, st, Whh, hs, sv, dXe, dG, Hq, nullptr, nullptr, 0, act));
    } else {
      RC(gru_bwd<TA>(d, w, st, Whh, hs, sv, dXe, dG, Hq, nullptr, w.dhenc + Hq, Hin));
    }
    RC(memset_async(w.dW_p, (size_t)3 * Hq * Hq * 4, st));
    RC(gemm<TA>(w.err_flag, st, dG + Hq, 4 * Hq, true, hs, Hq, false, w.dW_p, Hq, false, 3 * Hq, Hq, TB, nullptr, true, wsplits, 256,
                (rev && !swept) ? nullptr : VLK));
    simt::unpad_gate_matrix_kernel<<<grid_for(3ll * Hq * Hq), 256, 0, st>>>(w.dW_p, Hq, Hq, G[ix.e_whh(rev)], Hq, Hq, 0, 1, 2); KCHECK();
    RC(memset_async(w.csum, (size_t)4 * Hd * 4, st));
    RC(simt::colsum<TA>(st, dG, TB, 4 * Hq, 4 * Hq, w.csum, swept ? lim256 : nullptr, Bp)); mvae_count_launches(1);
    gate_bias_grads_kernel<<<ceil_div(3 * Hq, 256), 256, 0, st>>>(w.csum, Hq, G[ix.e_bih(rev)], G[ix.e_bhh(rev)]); KCHECK();
    if (rev) {   // one-hot rows of the reverse direction's token stream (per-row reversal when it ran on the persistent kernel)
      onehot_rows_kernel<TA><<<grid_for((long long)TB * CP), 256, 0, st>>>(ids, T, B, Bp, T, CP, (TA*)w.OH, swept ? 2 : 1, lens); KCHECK();
    }
    RC(memset_async(w.dTBL, (size_t)CP * 3 * Hd * 4, st));
    RC(gemm<TA>(w.err_flag, st, (const TA*)w.OH, CP, true, dG, 4 * Hq, false, w.dTBL, 3 * Hq, false, CP, 3 * Hq, TB, nullptr, true,
                d.bf16 ? 24 : 64, 256, (rev && !swept) ? nullptr : VLK));
    tbl_grad_to_rzn_T_kernel<<<(unsigned)ceil_div64(3ll * Hq * V, 256), 256, 0, st>>>(w.dTBL, Hq, V, w.dWT); KCHECK();
    RC(sg(st, w.dWT, V, 1, P[ix.emb()], V, 1, G[ix.e_wih(rev)], V, 3 * Hq, V, V, nullptr, simt::ACT_NONE, 0));
    RC(sg(st, w.dWT, 1, V, P[ix.e_wih(rev)], V, 1, G[ix.emb()], V, V, V, 3 * Hq, nullptr, simt::ACT_NONE, 1, 24));   // accumulate onto the decoder part
  }
  if (d.pad >= 0 && d.pad < V) {   // nn.Embedding(padding_idx = pad): the pad row receives no gradient
    zero_row_kernel<<<1, 256, 0, st>>>(G[ix.emb()], d.pad, V); KCHECK();
  }
  // a fired pipeline watchdog poisons the returned loss (the scalars were finalised before the backward kernels ran)
  simt::nan_if_error_kernel<<<1, 1, 0, st>>>(w.err_flag, out_scalars); KCHECK();
  return MVAE_OK;
}

// one decode step of VAE.sample (mosesvae.py:243-255): y/temp -> argmax (greedy) or inverse-CDF draw (multinomial) ->
// masked write + EOS bookkeeping.  One warp per sequence; logits row = [CP] fp32.
__device__ __forceinline__ float u01_hash(unsigned long long seed, unsigned int b, unsigned int i) {
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)b * 1000003ull + i + 1);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
  return (float)((x >> 40) + 0.5) * (1.0f / 16777216.0f);
}
__global__ void sample_step_kernel(const float* __restrict__ logits, int CP, int V, int B, int step, int max_len,
                                   int eos, int mode, float inv_temp, unsigned long long seed, const unsigned long long* seed_dev,
                                   uint8_t* __restrict__ w_cur, uint8_t* __restrict__ x, int* __restrict__ end,
                                   uint8_t* __restrict__ done) {
  // one warp per sequence; lane l owns the ids l, l + 32, ... (NV = CP / 32 <= 8 of them)
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int b = warp, NV = CP >> 5;
  float a[8];
  float m = -INFINITY;
  int arg = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int v = lane + 32 * k;
    a[k] = (k < NV && v < V) ? logits[(long long)b * CP + v] * inv_temp : -INFINITY;
    if (a[k] > m) { m = a[k]; arg = v; }          // strict: ties keep the lower id (k ascending = id ascending per lane)
  }
  if (arg == 0x7fffffff) arg = lane;
  for (int o = 16; o; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
  }
  int tok = arg;
  if (mode == 1) {
    // inverse-CDF draw over the ids in ascending order: chunk k holds the ids [32k, 32k + 32); inclusive prefix sums inside a
    // chunk over the lanes, chunk totals accumulated in order
    float c[8], tot[8], base = 0.f, total = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      c[k] = (k < NV && lane + 32 * k < V) ? expf(a[k] - m) : 0.f;
      for (int o = 1; o < 32; o <<= 1) {
        const float t0 = __shfl_up_sync(0xffffffffu, c[k], o);
        if (lane >= o) c[k] += t0;
      }
      tot[k] = __shfl_sync(0xffffffffu, c[k], 31);
      total += tot[k];
    }
    const float u = u01_hash(seed_dev ? *seed_dev : seed, (unsigned)b, (unsigned)step) * total;
    tok = -1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned mk = __ballot_sync(0xffffffffu, base + c[k] > u);
      if (tok < 0 && mk) tok = 32 * k + __ffs(mk) - 1;
      base += tot[k];
    }
    if (tok < 0 || tok >= V) tok = arg;
  }
  if (lane == 0) {
    w_cur[b] = (uint8_t)tok;
    if (!done[b]) {
      x[(long long)b * max_len + step] = (uint8_t)tok;
      if (tok == eos) { end[b] = step + 1; done[b] = 1; }
    }
  }
}
__global__ void sample_init_kernel(int B, int max_len, int bos, int pad, uint8_t* __restrict__ w_cur,
                                   uint8_t* __restrict__ x, int* __restrict__ end, uint8_t* __restrict__ done) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < (long long)B * max_len) x[i] = (i % max_len == 0) ? (uint8_t)bos : (uint8_t)pad;
  if (i < B) { w_cur[i] = (uint8_t)bos; end[i] = max_len; done[i] = 0; }
}

// dst[b][0..H) (row stride ld) = h0[b][:] as bf16, rows >= B zero; optional fp32 copy
__global__ void init_state_kernel(const float* __restrict__ h0, int B, int Bp, int H, __nv_bfloat16* __restrict__ dst, long long ld,
                                  float* __restrict__ h32) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)Bp * H) return;
  const int b = (int)(idx / H), j = (int)(idx - (long long)b * H);
  const float v = b < B ? h0[idx] : 0.f;
  dst[(long long)b * ld + j] = __float2bfloat16_rn(v);
  if (h32) h32[idx] = v;
}

// Layer 0 of the sampler as ONE K-concatenated contraction (no look-ups in the epilogue): operand row = [onehot(token) : 64 |
// z : ZP | h : H], weights [TBL^T | W_ih[:, V:] | W_hh] in tiles of 64 units x 4 gate blocks (r and z summed over all three
// parts, `in` from the input parts only, `hn` from the hidden part only).  TBL = x_emb . W_ih[:, :V]^T ([V][3H], fp32).
__global__ void cell_weights_l0_kernel(const float* __restrict__ tbl, const float* __restrict__ wih, const float* __restrict__ whh,
                                       const float* __restrict__ bih, const float* __restrict__ bhh, int H, int V, int Z, int ZP,
                                       __nv_bfloat16* __restrict__ W, float* __restrict__ bias) {
  const int K = 64 + ZP + H;
  const long long total = 4ll * H * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % K);
    const int n = (int)(i / K);
    const int tile = n / 256, g = (n / 64) % 4, j = n % 64;
    const int unit = tile * 64 + j;
    const int gate = g < 2 ? g : 2;
    float v = 0.f;
    if (c < 64) { if (g != 3 && c < V) v = tbl[(long long)c * 3 * H + gate * H + unit]; }
    else if (c < 64 + ZP) { if (g != 3 && c - 64 < Z) v = wih[((long long)gate * H + unit) * (V + Z) + V + (c - 64)]; }
    else if (g != 2) v = whh[((long long)gate * H + unit) * H + (c - 64 - ZP)];
    W[i] = __float2bfloat16_rn(v);
    if (c == 0) bias[n] = g < 2 ? bih[g * H + unit] + bhh[g * H + unit] : (g == 2 ? bih[2 * H + unit] : bhh[2 * H + unit]);
  }
}
// constant parts of the layer-0 operand rows: one-hot(bos) in columns [0,64) (buffer `oh` only) and z in [64, 64+ZP)
__global__ void sample_l0_operand_kernel(const float* __restrict__ z, int B, int Bp, int Z, int ZP, long long ld, int bos,
                                         __nv_bfloat16* __restrict__ buf0, __nv_bfloat16* __restrict__ buf1) {
  const long long total = (long long)Bp * (64 + ZP);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (64 + ZP)), b = (int)(i / (64 + ZP));
    float v0 = 0.f, v1 = 0.f;
    if (c >= 64) { if (b < B && c - 64 < Z) v0 = v1 = z[(long long)b * Z + (c - 64)]; }
    else if (c == bos) v1 = 1.f;                       // the first step reads buffer 1
    buf0[(long long)b * ld + c] = __float2bfloat16_rn(v0);
    buf1[(long long)b * ld + c] = __float2bfloat16_rn(v1);
  }
}
// MVAE_SAMPLE_PERSISTENT: 0 = per-step launches, 1 = persistent kernel with one CTA per unit, 2 (default) = CTA pairs
int sample_persistent_variant() {
  const char* e = getenv("MVAE_SAMPLE_PERSISTENT");
  return e ? atoi(e) : 2;
}
bool sample_persistent_enabled() { return sample_persistent_variant() != 0; }
bool sample_fused_enabled() {
  const char* e = getenv("MVAE_SAMPLE_FUSED");
  return e ? atoi(e) != 0 : true;
}

// VAE.sample with one tensor-core GEMM per decoder layer and step: the GRU cell runs in the GEMM epilogue (r/z
// contributions of x and h summed by a K = 2H contraction over the concatenated operand), so no gate pre-activations
// reach HBM and a step is L + 1 GEMMs + 2 small kernels instead of 2L GEMMs + L gate kernels + 3.
int sample_fused(const MDims& d, const MWS& w, const float* const* P, const float* z, int bos, int eos, int mode, float temp,
                 unsigned long long seed, const unsigned long long* seed_dev, uint8_t* ids_out, int* len_out, uint8_t* w_cur,
                 uint8_t* done, cudaStream_t st) {
  typedef __nv_bfloat16 TA;
  const int B = d.B, Bp = d.Bp, V = d.V, CP = d.CP, Z = d.Z, Hd = d.Hd, L = d.L, max_len = d.T;
  const int IN0 = V + Z;
  const MP ix{d.bidir, d.lin, L};
  RC(memset_async(w.err_flag, 4, st));
  // layer 0 as one contraction over [onehot | z | h] when that operand fits the [x | h] buffers (64 + ZP <= Hd)
  const int ZP = round_up(Z, 64), K0 = 64 + ZP + Hd;
  const bool kcat = V <= 64 && 64 + ZP <= Hd && getenv("MVAE_SAMPLE_KCAT0") == nullptr;
  for (int l = kcat ? 1 : 0; l < L; ++l) {
    cell_weights_kernel<<<grid_for((long long)(l ? 4 : 3) * Hd * (l ? 2 : 1) * Hd), 256, 0, st>>>(
        l ? P[ix.wih(l)] : nullptr, P[ix.whh(l)], l ? P[ix.bih(l)] : nullptr, P[ix.bhh(l)], Hd, l ? 4 : 3, (TA*)w.Wcat[l], w.bcat[l]);
    KCHECK();
  }
  simt::pad_matrix_kernel<TA><<<grid_for((long long)CP * Hd), 256, 0, st>>>(P[ix.fcw()], V, Hd, (TA*)w.Wfc, CP, Hd); KCHECK();
  simt::pad_matrix_kernel<float><<<1, 256, 0, st>>>(P[ix.fcb()], 1, V, w.bfc, 1, CP); KCHECK();
  RC(memset_async(w.h0, (size_t)Bp * Hd * 4, st));
  RC(sg(st, z, Z, 1, P[ix.latw()], 1, Z, w.h0, Hd, B, Hd, Z, P[ix.latb()], simt::ACT_NONE, 0));
  RC(sg(st, P[ix.emb()], V, 1, P[ix.wih(0)], 1, IN0, w.TBLd, 3 * Hd, V, 3 * Hd, V, nullptr, simt::ACT_NONE, 0));
  if (kcat) {
    cell_weights_l0_kernel<<<grid_for(4ll * Hd * K0), 256, 0, st>>>(w.TBLd, P[ix.wih(0)], P[ix.whh(0)], P[ix.bih(0)], P[ix.bhh(0)], Hd, V, Z,
                                                                   ZP, (TA*)w.Wcat[0], w.bcat[0]); KCHECK();
  } else {
    RC(sg(st, z, Z, 1, P[ix.wih(0)] + V, 1, IN0, w.zproj, 3 * Hd, B, 3 * Hd, Z, P[ix.bih(0)], simt::ACT_NONE, 0));
  }
  const int sgrid = (int)ceil_div64((long long)Bp * Hd, 256);
  const int h_off0 = kcat ? 64 + ZP : 0, ld0 = kcat ? K0 : Hd;   // where h sits in layer 0's operand rows
  for (int l = 0; l < L; ++l) {
    // the first step (i = 1) reads parity 1; layer 0's operand is [h] (ld Hd) or [onehot | z | h] (ld K0), the others [x | h]
    // (ld 2Hd, h in the right half)
    for (int k = 0; k < 2; ++k) RC(memset_async(w.xh[l][k], (size_t)Bp * 2 * Hd * 2, st));
    init_state_kernel<<<sgrid, 256, 0, st>>>(w.h0, B, Bp, Hd, (TA*)w.xh[l][1] + (l ? Hd : h_off0), l ? 2 * Hd : ld0, w.hm[l][1]); KCHECK();
  }
  if (kcat) {
    sample_l0_operand_kernel<<<grid_for((long long)Bp * (64 + ZP)), 256, 0, st>>>(z, B, Bp, Z, ZP, K0, bos, (TA*)w.xh[0][0], (TA*)w.xh[0][1]); KCHECK();
  }
  sample_init_kernel<<<(unsigned)ceil_div64((long long)B * max_len, 256), 256, 0, st>>>(B, max_len, bos, d.pad, w_cur, ids_out, len_out, done); KCHECK();
  if (kcat && CP == 64 && L >= 2 && sample_persistent_enabled()) {
    // the whole decode loop as ONE launch (decode_persist.cu): same GEMM pipeline and epilogues, units of all layers and steps
    // in one dependency-ordered list; falls through to the per-step launches when the shape is not supported
    mvae_decode_args da{};
    da.B = B; da.Bp = Bp; da.Hd = Hd; da.L = L; da.V = V; da.K0 = K0; da.max_len = max_len; da.eos = eos; da.mode = mode;
    da.inv_temp = 1.0f / temp; da.seed = seed; da.seed_dev = seed_dev;
    for (int l = 0; l < L; ++l) {
      da.Wcat[l] = w.Wcat[l]; da.bcat[l] = w.bcat[l];
      for (int k = 0; k < 2; ++k) { da.xh[l][k] = w.xh[l][k]; da.hm[l][k] = w.hm[l][k]; }
    }
    da.Wfc = w.Wfc; da.bfc = w.bfc; da.w_cur = w_cur; da.x = ids_out; da.end = len_out; da.done = done;
    da.counters = w.dec_scratch; da.sched = w.dec_scratch + (size_t)(L + 1) * (Bp / 128); da.err_flag = w.err_flag;
    da.variant = sample_persistent_variant();
    const int rc = mvae_decode_persistent_launch(&da, st);
    if (rc == MVAE_OK) { mvae_count_launches(3); return MVAE_OK; }
    if (rc != MVAE_ERR_UNSUPPORTED) return rc;
  }
  for (int i = 1; i < max_len; ++i) {
    const int cur = i & 1, nxt = cur ^ 1;
    for (int l = 0; l < L; ++l) {
      const int G = (l || kcat) ? 4 : 3, K = l ? 2 * Hd : ld0;
      mvae_umma_operand a{w.xh[l][cur], 0, Bp, K, K, 1, 0, 0, 0};
      mvae_umma_operand b{w.Wcat[l], 0, (long long)G * Hd, K, K, 1, 0, 0, 0};
      mvae_umma_out o{w.logits, (long long)G * Hd, 0, 0, w.bcat[l], 0};
      mvae_umma_cell c{};
      c.gates = G; c.H = Hd; c.h_prev32 = w.hm[l][cur]; c.h_next32 = w.hm[l][nxt];
      if (l == 0 && !kcat) { c.tbl = w.TBLd; c.add = w.zproj; c.tok = w_cur; c.tok_rows = B; }   // emb(w) | z through W_ih_l0: look-up + per-sequence part
      c.out_a = (TA*)w.xh[l][nxt] + (l ? Hd : h_off0); c.ld_a = K;
      c.out_b = (l + 1 < L) ? w.xh[l + 1][cur] : nullptr; c.ld_b = 2 * Hd;
      mvae_count_launches(1);
      RC(mvae_umma_gemm(&a, &b, &o, Bp, G * Hd, K, G * 64, 1, 0, w.err_flag, st, nullptr, &c));
    }
    // vocabulary GEMM with the sampling step in its epilogue: logits never reach HBM
    const TA* top = (const TA*)w.xh[L - 1][nxt] + (L > 1 ? Hd : 0);
    mvae_umma_operand a{top, 0, Bp, Hd, L > 1 ? 2 * Hd : Hd, 1, 0, 0, 0};
    mvae_umma_operand b{w.Wfc, 0, CP, Hd, Hd, 1, 0, 0, 0};
    mvae_umma_out o{w.logits, CP, 0, 0, w.bfc, 0};
    mvae_count_launches(1);
    if (CP == 64) {
      mvae_umma_sample sp{V, B, i, max_len, eos, mode, 1.0f / temp, seed, seed_dev, w_cur, kcat ? w.xh[0][nxt] : nullptr, K0, ids_out, len_out, done};
      RC(mvae_umma_gemm(&a, &b, &o, Bp, CP, Hd, 64, 1, 0, w.err_flag, st, nullptr, nullptr, &sp));
    } else {   // large vocabulary: logits of all CP / 64 tiles to memory, then one warp per sequence picks the token
      RC(mvae_umma_gemm(&a, &b, &o, Bp, CP, Hd, 64, 1, 0, w.err_flag, st));
      sample_step_kernel<<<ceil_div(B * 32, 256), 256, 0, st>>>(w.logits, CP, V, B, i, max_len, eos, mode, 1.0f / temp, seed, seed_dev, w_cur,
                                                                ids_out, len_out, done);
      KCHECK();
    }
  }
  return MVAE_OK;
}

template <typename TA>
int sample_t(const MDims& d, const MWS& w, const float* const* P, const float* z, int bos, int eos, int mode, float temp,
             unsigned long long seed, const unsigned long long* seed_dev, uint8_t* ids_out, int* len_out, uint8_t* w_cur,
             uint8_t* done, cudaStream_t st) {
  const int B = d.B, Bp = d.Bp, V = d.V, CP = d.CP, Z = d.Z, Hd = d.Hd, L = d.L, max_len = d.T;
  const int IN0 = V + Z;
  const MP ix{d.bidir, d.lin, L};
  const size_t slab = (size_t)Bp * Hd;
  RC(memset_async(w.err_flag, 4, st));
  for (int l = 0; l < L; ++l) {
    simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(P[ix.whh(l)], Hd, Hd, (TA*)w.Whh[l], Hd, Hd, 0, 1, 2); KCHECK();
    simt::pad_gate_vector_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(P[ix.bhh(l)], Hd, w.bhh[l], Hd); KCHECK();
    if (l >= 1) {
      simt::pad_gate_matrix_kernel<TA><<<grid_for(3ll * Hd * Hd), 256, 0, st>>>(P[ix.wih(l)], Hd, Hd, (TA*)w.Wih[l], Hd, Hd, 0, 1, 2); KCHECK();
      simt::pad_gate_vector_kernel<<<ceil_div(3 * Hd, 256), 256, 0, st>>>(P[ix.bih(l)], Hd, w.bih[l], Hd); KCHECK();
    }
  }
  simt::pad_matrix_kernel<TA><<<grid_for((long long)CP * Hd), 256, 0, st>>>(P[ix.fcw()], V, Hd, (TA*)w.Wfc, CP, Hd); KCHECK();
  simt::pad_matrix_kernel<float><<<1, 256, 0, st>>>(P[ix.fcb()], 1, V, w.bfc, 1, CP); KCHECK();
  // h0 (mosesvae.py:229-230), token table and the per-sequence z part of the layer-0 projection
  RC(memset_async(w.h0, (size_t)Bp * Hd * 4, st));
  RC(sg(st, z, Z, 1, P[ix.latw()], 1, Z, w.h0, Hd, B, Hd, Z, P[ix.latb()], simt::ACT_NONE, 0));
  RC(sg(st, P[ix.emb()], V, 1, P[ix.wih(0)], 1, IN0, w.TBLd, 3 * Hd, V, 3 * Hd, V, nullptr, simt::ACT_NONE, 0));
  RC(sg(st, z, Z, 1, P[ix.wih(0)] + V, 1, IN0, w.zproj, 3 * Hd, B, 3 * Hd, Z, P[ix.bih(0)], simt::ACT_NONE, 0));
  // per layer two hidden-state slabs (ping-pong by step parity): hs[l] slab (step & 1) holds h before the step
  for (int l = 0; l < L; ++l) {
    init_h0_kernel<TA><<<ceil_div((int)slab, 256), 256, 0, st>>>(w.h0, B, Bp, Hd, (TA*)w.hs[l] + slab, nullptr); KCHECK();
  }
  // fp32 master states: reuse the saved-gate buffers (not needed when sampling): sv[l] holds 2 fp32 slabs
  float* h32[4][2];
  for (int l = 0; l < L; ++l) {
    h32[l][0] = reinterpret_cast<float*>(w.sv[l]);
    h32[l][1] = h32[l][0] + slab;
    if (d.bf16) { init_h0_kernel<float><<<ceil_div((int)slab, 256), 256, 0, st>>>(w.h0, B, Bp, Hd, h32[l][1], nullptr); KCHECK(); }
  }
  sample_init_kernel<<<(unsigned)ceil_div64((long long)B * max_len, 256), 256, 0, st>>>(B, max_len, bos, d.pad, w_cur, ids_out, len_out, done); KCHECK();
  const int gate_grid = ceil_div((int)slab, 256);
  TA* gi0 = (TA*)w.gi;                       // [Bp][3Hd]
  TA* gi1 = gi0 + (size_t)Bp * 3 * Hd;       // projection of the lower layer's new state
  for (int i = 1; i < max_len; ++i) {
    const int cur = i & 1, nxt = cur ^ 1;
    gather_rows_kernel<TA><<<grid_for((long long)Bp * 3 * Hd), 256, 0, st>>>(w.TBLd, 3 * Hd, w_cur, 1, w.zproj, B, Bp, 1, gi0); KCHECK();
    for (int l = 0; l < L; ++l) {
      TA* hcur = (TA*)w.hs[l] + cur * slab;
      TA* hnxt = (TA*)w.hs[l] + nxt * slab;
      const TA* gi = gi0;
      if (l >= 1) {
        RC(gemm<TA>(w.err_flag, st, (const TA*)w.hs[l - 1] + nxt * slab, Hd, false, (const TA*)w.Wih[l], Hd, true, gi1, 3 * Hd, true,
                    Bp, 3 * Hd, Hd, w.bih[l], false, 1));
        gi = gi1;
      }
      RC(gemm<TA>(w.err_flag, st, hcur, Hd, false, (const TA*)w.Whh[l], Hd, true, w.gh, 3 * Hd, false, Bp, 3 * Hd, Hd, w.bhh[l], false, 1));
      simt::gru_gate_fwd_kernel<TA, TA><<<gate_grid, 256, 0, st>>>(gi, w.gh, d.bf16 ? h32[l][cur] : nullptr, hcur, hnxt,
                                                                   d.bf16 ? h32[l][nxt] : nullptr, nullptr, Bp, Hd);
      KCHECK();
    }
    RC(gemm<TA>(w.err_flag, st, (const TA*)w.hs[L - 1] + nxt * slab, Hd, false, (const TA*)w.Wfc, Hd, true, w.logits, CP, false, Bp, CP, Hd,
                w.bfc, false, 1, 64));
    sample_step_kernel<<<ceil_div(B * 32, 256), 256, 0, st>>>(w.logits, CP, V, B, i, max_len, eos, mode, 1.0f / temp, seed, seed_dev, w_cur,
                                                              ids_out, len_out, done);
    KCHECK();
  }
  return MVAE_OK;
}

int check_ws(const mvae_moses_desc* desc, void* ws, size_t ws_bytes, MDims* d, MWS* w) {
  RC(make_dims(desc, d));
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) return MVAE_ERR_INVALID;
  carve(*d, ws, w);
  if (ws_bytes < w->total) return MVAE_ERR_WORKSPACE;
  g_tc = mvae_tc_ctx{(d->bf16 && tc_sgemm_enabled()) ? w->tcs : nullptr, w->tcs_bytes, w->err_flag};
  return MVAE_OK;
}

}  // namespace

extern "C" {

size_t mvae_moses_workspace_bytes(const mvae_moses_desc* desc) {
  MDims d; MWS w;
  if (make_dims(desc, &d) != MVAE_OK) return 0;
  carve(d, nullptr, &w);
  return w.total;
}

static int moses_step_impl(const mvae_moses_desc* desc, const float* const* params, float* const* grads, const uint8_t* ids,
                           const int32_t* lengths, const int32_t* lengths_host, const float* eps, float* out_scalars, float* z_out,
                           float* logvar_out, float* y_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream,
                           int phase, const float* dz_ext, const JointHead* jh) {
  MDims d; MWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !ids || !lengths || !eps || !out_scalars) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool backward = grads != nullptr;
  if (phase >= 0 && !backward) return MVAE_ERR_INVALID;
  // packed-sequence batch sizes (torch pack_sequence): act[t] = number of sequences longer than t; with the batch sorted
  // by length they are the rows [0, act[t]).  Needs the host copy of the lengths; without it every step covers all rows.
  int act_buf[512];
  const int* act = nullptr;
  if (lengths_host) {
    for (int b = 0; b + 1 < d.B; ++b)
      if (lengths_host[b] < lengths_host[b + 1]) return MVAE_ERR_INVALID;          // pack_sequence contract: sorted descending
    if (lengths_host[0] > d.T || lengths_host[d.B - 1] < 2) return MVAE_ERR_INVALID;
    int b = d.B;
    for (int t = 0; t < d.T; ++t) {
      while (b > 0 && lengths_host[b - 1] <= t) --b;
      act_buf[t] = b;
    }
    act = act_buf;
  }
  return d.bf16 ? step_t<__nv_bfloat16>(d, w, params, grads, ids, lengths, eps, out_scalars, z_out, logvar_out, y_out, backward, st, act,
                                        phase, dz_ext, jh)
                : step_t<float>(d, w, params, grads, ids, lengths, eps, out_scalars, z_out, logvar_out, y_out, backward, st, act, phase,
                                dz_ext, jh);
}

int mvae_moses_step(const mvae_moses_desc* desc, const float* const* params, float* const* grads, const uint8_t* ids,
                    const int32_t* lengths, const int32_t* lengths_host, const float* eps, float* out_scalars, float* z_out,
                    float* logvar_out, float* y_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  return moses_step_impl(desc, params, grads, ids, lengths, lengths_host, eps, out_scalars, z_out, logvar_out, y_out, workspace,
                         workspace_bytes, stream, -1, nullptr, nullptr);
}

int mvae_moses_step_ex(const mvae_moses_desc* desc, const float* const* params, float* const* grads, const uint8_t* ids,
                       const int32_t* lengths, const int32_t* lengths_host, const float* eps, const float* dz_ext,
                       float* out_scalars, float* z_out, float* logvar_out, float* y_out, void* workspace,
                       size_t workspace_bytes, int phase, mvae_stream_t stream) {
  return moses_step_impl(desc, params, grads, ids, lengths, lengths_host, eps, out_scalars, z_out, logvar_out, y_out, workspace,
                         workspace_bytes, stream, phase, dz_ext, nullptr);
}

size_t mvae_moses_joint_extra_bytes(const mvae_moses_desc* desc) {
  // pred [B] + dout [B] + dz [B][Z] + loss (16 B), each 256-byte aligned
  if (!desc || desc->batch <= 0 || desc->d_z <= 0) return 0;
  const size_t B = desc->batch, Z = desc->d_z;
  auto up = [](size_t n) { return (n + 255) & ~size_t(255); };
  return up(B * 4) * 2 + up(B * Z * 4) + 256;
}

int mvae_moses_joint_step(const mvae_moses_desc* desc, const float* const* params, float* const* grads, const uint8_t* ids,
                          const int32_t* lengths, const int32_t* lengths_host, const float* eps,
                          const mvae_binding_desc* bdesc, const float* const* bparams, float* const* bgrads,
                          float* const* brunning, const float* target, float binding_weight, float* out_scalars,
                          float* binding_loss_out, float* z_out, void* workspace, size_t workspace_bytes,
                          void* binding_workspace, size_t binding_workspace_bytes, void* extra, size_t extra_bytes, int phase,
                          mvae_stream_t stream) {
  if (!bdesc || !bparams || !brunning || !target || !binding_loss_out || !binding_workspace || !extra) return MVAE_ERR_INVALID;
  if (grads && !bgrads) return MVAE_ERR_INVALID;
  if (bdesc->batch != desc->batch || bdesc->z_size != desc->d_z) return MVAE_ERR_INVALID;
  if (extra_bytes < mvae_moses_joint_extra_bytes(desc) || (reinterpret_cast<uintptr_t>(extra) & 255)) return MVAE_ERR_WORKSPACE;
  const size_t B = desc->batch, Z = desc->d_z;
  auto up = [](size_t n) { return (n + 255) & ~size_t(255); };
  uint8_t* e = reinterpret_cast<uint8_t*>(extra);
  JointHead jh{bdesc, bparams, bgrads, brunning, target, binding_weight, nullptr, nullptr, nullptr, binding_loss_out,
               binding_workspace, binding_workspace_bytes};
  jh.pred = reinterpret_cast<float*>(e); e += up(B * 4);
  jh.dout = reinterpret_cast<float*>(e); e += up(B * 4);
  jh.dz = reinterpret_cast<float*>(e);
  (void)Z;
  return moses_step_impl(desc, params, grads, ids, lengths, lengths_host, eps, out_scalars, z_out, nullptr, nullptr, workspace,
                         workspace_bytes, stream, phase, nullptr, &jh);
}

int mvae_moses_sample(const mvae_moses_desc* desc, const float* const* params, const float* z, int bos_id, int eos_id,
                      int mode, float temp, unsigned long long seed, const unsigned long long* seed_device, uint8_t* ids_out,
                      int32_t* lengths_out, void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  MDims d; MWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !z || !ids_out || !lengths_out || temp <= 0.f || (mode != 0 && mode != 1)) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // small per-sequence state lives at the start of the (unused while sampling) one-hot buffer
  uint8_t* w_cur = reinterpret_cast<uint8_t*>(w.OH);
  uint8_t* done = w_cur + d.Bp;
  if (d.bf16 && sample_fused_enabled())
    return sample_fused(d, w, params, z, bos_id, eos_id, mode, temp, seed, seed_device, ids_out, lengths_out, w_cur, done, st);
  return d.bf16 ? sample_t<__nv_bfloat16>(d, w, params, z, bos_id, eos_id, mode, temp, seed, seed_device, ids_out, lengths_out, w_cur, done, st)
                : sample_t<float>(d, w, params, z, bos_id, eos_id, mode, temp, seed, seed_device, ids_out, lengths_out, w_cur, done, st);
}

int mvae_moses_sample_graph_create(const mvae_moses_desc* desc, const float* const* params, const float* z, int bos_id,
                                   int eos_id, int mode, float temp, const unsigned long long* seed_device,
                                   uint8_t* ids_out, int32_t* lengths_out, void* workspace, size_t workspace_bytes,
                                   mvae_graph** out_graph) {
  if (!seed_device) return MVAE_ERR_INVALID;
  struct Ctx {
    const mvae_moses_desc* desc; const float* const* params; const float* z; int bos, eos, mode; float temp;
    const unsigned long long* seed_dev; uint8_t* ids; int32_t* lens; void* ws; size_t ws_bytes;
  } c{desc, params, z, bos_id, eos_id, mode, temp, seed_device, ids_out, lengths_out, workspace, workspace_bytes};
  return mvae_capture_into_graph(
      [](void* p, cudaStream_t cs) {
        Ctx* c = static_cast<Ctx*>(p);
        return mvae_moses_sample(c->desc, c->params, c->z, c->bos, c->eos, c->mode, c->temp, 0ull, c->seed_dev, c->ids,
                                 c->lens, c->ws, c->ws_bytes, reinterpret_cast<mvae_stream_t>(cs));
      },
      &c, out_graph);
}

int mvae_moses_read_error(const mvae_moses_desc* desc, void* workspace, size_t workspace_bytes, int* flag,
                          mvae_stream_t stream) {
  MDims d; MWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!flag) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MVAE_CUDA_CHECK(cudaMemcpyAsync(flag, w.err_flag, 4, cudaMemcpyDeviceToHost, st));
  MVAE_CUDA_CHECK(cudaStreamSynchronize(st));
  return MVAE_OK;
}

}  // extern "C"
