// "Config A": models.py exactly as shipped (models.py:97-165) + loss_function (train.py:31-38), fused forward + loss +
// backward of one batch on one B200.
//
//   encoder (models.py:126-135): Embedding(C,E) -> LSTM EL x EH -> ConvSELU(T->120,k18) -> ConvSELU(120->64,k18) ->
//            ConvSELU(64->64,k18) -> Flatten -> Linear(.,512)+SELU -> Lambda (mu, log_v, z = mu + exp(log_v/2) 1e-2 eps)
//   decoder (models.py:161-165): Linear(Z,Z)+SELU -> Repeat(T) -> LSTM DL x DH -> Linear(DH,C) -> Softmax
//   (the attributes are called "gru" in the reference but are nn.LSTM: gate rows i,f,g,o; c' = f c + i g; h' = o tanh c')
// Layout: recurrent tensors time-major [T][Bp][Hp] (Bp = B rounded to 128, Hp = H rounded to 64, pads are zeros and stay
// zero).  The embedding makes the encoder's layer-0 projection a table look-up (E W_ih^T + b, C rows); the decoder's
// layer-0 projection is time-invariant (Repeat(T) is never materialised) and computed once per molecule.
// The convolutions see the T sequence positions as channels and the LSTM feature axis as length (the reference's
// Keras-port quirk): each one is an im2col + one tensor-core GEMM [B*Lout, Cin*18] x [Cin*18, Cout]; their backward is two
// GEMMs (dW = da^T cols, dcols = da W) + a gather-form col2im.
// Every dense contraction with K >= 64 goes through the tcgen05 GEMM in bf16 mode; fp32 mode is CUDA-core FMA only.
#include <stdlib.h>

#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "umma_gemm.h"
#include "host_common.cuh"

namespace {

constexpr int KS = 18;          // kernel size of all three convolutions (models.py:118-120)
constexpr int C1 = 120, C2 = 64, C3 = 64, D1 = 512;

struct ADims {
  int B, Bp, T, C, CP, E, EH, EHp, EL, Z, DH, DL;
  int L1, L2, L3, FLAT, K1, K1p, K2, K2p, K3, K3p;
  int Hmax;
  bool bf16;
  float max_len, eps_scale;
};

int make_dims(const mvae_cfga_desc* d, ADims* o) {
  if (!d) return MVAE_ERR_INVALID;
  if (d->batch <= 0 || d->seq_len < 1 || d->seq_len > 512 || d->charset < 2 || d->charset > 64 || d->embed <= 0 ||
      d->enc_hidden < 3 * (KS - 1) + 1 || d->enc_layers < 1 || d->enc_layers > 4 || d->latent <= 0 || d->dec_hidden <= 0 ||
      d->dec_layers < 1 || d->dec_layers > 4)
    return MVAE_ERR_INVALID;
  if (d->dec_hidden & 63) return MVAE_ERR_UNSUPPORTED;   // decoder hidden size must be a multiple of 64 (1024 in models.py:150)
  if (d->precision != MVAE_PREC_FP32 && d->precision != MVAE_PREC_BF16) return MVAE_ERR_INVALID;
  o->B = d->batch; o->Bp = round_up(d->batch, 128); o->T = d->seq_len; o->C = d->charset; o->CP = 64; o->E = d->embed;
  o->EH = d->enc_hidden; o->EHp = round_up(d->enc_hidden, 64); o->EL = d->enc_layers; o->Z = d->latent;
  o->DH = d->dec_hidden; o->DL = d->dec_layers;
  o->L1 = o->EH - (KS - 1); o->L2 = o->L1 - (KS - 1); o->L3 = o->L2 - (KS - 1); o->FLAT = C3 * o->L3;
  o->K1 = o->T * KS; o->K1p = round_up(o->K1, 8); o->K2 = C1 * KS; o->K2p = round_up(o->K2, 8);
  o->K3 = C2 * KS; o->K3p = round_up(o->K3, 8);
  o->Hmax = o->DH > o->EHp ? o->DH : o->EHp;
  o->bf16 = d->precision == MVAE_PREC_BF16; o->max_len = d->max_len; o->eps_scale = d->eps_scale;
  return MVAE_OK;
}

struct AWS {
  int* err_flag; double* bce_sum; double* kl_sum; int* hit_count;
  // weights in operand form
  void *Whh_e[4], *Wih_e[4], *Whh_d[4], *Wih_d[4], *Wc[3], *Wfc;
  void *WhhC_e[4], *WhhC_d[4];   // W_hh in the fused-cell tile order (tiles of 64 units x [i|f|g|o]), bf16 mode
  float *bsum_e[4], *bsum_d[4], *bfc, *TBLe;
  // activations
  void *gi, *hs_e[4], *sv_e[4], *hs_d[4], *sv_d[4], *cols[3], *OH, *dlogits, *dG, *dX, *da;
  float *h[3], *flat, *h4, *mu, *lv, *z, *zr, *gi0, *gh, *c, *logits;
  // backward scratch
  float *dh_carry, *dc_carry, *dgisum, *dW_p, *dWc_p, *dWfc_p, *csum, *dTBL, *dhc, *dflat, *dh4, *dmu, *dlv, *dz, *dzr, *da5;
  void* tcs; size_t tcs_bytes;   // bf16 mode: converted operands of the bf16x3 tensor-core path of the small fp32 GEMMs
  size_t total;
};

void carve(const ADims& d, void* base, AWS* w) {
  Carver c{reinterpret_cast<uint8_t*>(base), 0};
  const size_t es = d.bf16 ? 2 : 4;
  const size_t B = d.B, Bp = d.Bp, T = d.T, EHp = d.EHp, DH = d.DH, Hm = d.Hmax, Z = d.Z;
  w->err_flag = c.take<int>(1); w->bce_sum = c.take<double>(1); w->kl_sum = c.take<double>(1);
  w->hit_count = c.take<int>(B);
  for (int l = 0; l < d.EL; ++l) {
    w->Whh_e[l] = c.take<uint8_t>(4 * EHp * EHp * es); w->Wih_e[l] = c.take<uint8_t>(4 * EHp * EHp * es);
    w->bsum_e[l] = c.take<float>(4 * EHp);
    w->WhhC_e[l] = c.take<uint8_t>(4 * EHp * EHp * 2);
  }
  for (int l = 0; l < d.DL; ++l) {
    w->Whh_d[l] = c.take<uint8_t>(4 * DH * DH * es); w->Wih_d[l] = c.take<uint8_t>(4 * DH * DH * es);
    w->bsum_d[l] = c.take<float>(4 * DH);
    w->WhhC_d[l] = c.take<uint8_t>(4 * DH * DH * 2);
  }
  w->Wc[0] = c.take<uint8_t>((size_t)C1 * d.K1p * es); w->Wc[1] = c.take<uint8_t>((size_t)C2 * d.K2p * es);
  w->Wc[2] = c.take<uint8_t>((size_t)C3 * d.K3p * es);
  w->Wfc = c.take<uint8_t>((size_t)d.CP * DH * es); w->bfc = c.take<float>(d.CP);
  w->TBLe = c.take<float>((size_t)d.CP * 4 * EHp);
  w->gi = c.take<uint8_t>(T * Bp * 4 * Hm * es);
  for (int l = 0; l < d.EL; ++l) {
    w->hs_e[l] = c.take<uint8_t>((T + 1) * Bp * EHp * es); w->sv_e[l] = c.take<uint8_t>(T * Bp * 6 * EHp * es);
  }
  for (int l = 0; l < d.DL; ++l) {
    w->hs_d[l] = c.take<uint8_t>((T + 1) * Bp * DH * es); w->sv_d[l] = c.take<uint8_t>(T * Bp * 6 * DH * es);
  }
  w->cols[0] = c.take<uint8_t>(B * d.L1 * d.K1p * es); w->cols[1] = c.take<uint8_t>(B * d.L2 * d.K2p * es);
  w->cols[2] = c.take<uint8_t>(B * d.L3 * d.K3p * es);
  w->OH = c.take<uint8_t>(T * Bp * d.CP * es); w->dlogits = c.take<uint8_t>(T * Bp * d.CP * es);
  w->dG = c.take<uint8_t>(T * Bp * 4 * Hm * es); w->dX = c.take<uint8_t>(T * Bp * Hm * es);
  w->da = c.take<uint8_t>(B * d.L1 * C1 * es);
  w->h[0] = c.take<float>(B * d.L1 * C1); w->h[1] = c.take<float>(B * d.L2 * C2); w->h[2] = c.take<float>(B * d.L3 * C3);
  w->flat = c.take<float>(B * d.FLAT); w->h4 = c.take<float>(B * D1);
  w->mu = c.take<float>(B * Z); w->lv = c.take<float>(B * Z); w->z = c.take<float>(B * Z); w->zr = c.take<float>(B * Z);
  w->gi0 = c.take<float>(Bp * 4 * DH); w->gh = c.take<float>(Bp * 4 * Hm); w->c = c.take<float>(Bp * Hm);
  w->logits = c.take<float>(T * Bp * d.CP);
  w->dh_carry = c.take<float>(Bp * Hm); w->dc_carry = c.take<float>(Bp * Hm); w->dgisum = c.take<float>(Bp * 4 * DH);
  w->dW_p = c.take<float>(4 * EHp * EHp);
  const size_t kmax = d.K1p > d.K2p ? d.K1p : d.K2p;
  w->dWc_p = c.take<float>((size_t)C1 * (kmax > (size_t)d.K3p ? kmax : (size_t)d.K3p));
  w->dWfc_p = c.take<float>((size_t)d.CP * DH); w->csum = c.take<float>(4 * Hm + 256);
  w->dTBL = c.take<float>((size_t)d.CP * 4 * EHp);
  w->dhc = c.take<float>(B * d.L1 * C1);
  w->dflat = c.take<float>(B * d.FLAT); w->dh4 = c.take<float>(B * D1);
  w->dmu = c.take<float>(B * Z); w->dlv = c.take<float>(B * Z); w->dz = c.take<float>(B * Z);
  w->dzr = c.take<float>(B * Z); w->da5 = c.take<float>(B * Z);
  {
    const long long rmax = (long long)max(Bp, 4 * DH), cmax = max(max((long long)4 * DH, (long long)d.FLAT), max((long long)D1, (long long)Z));
    w->tcs_bytes = d.bf16 ? mvae_tc_sgemm_scratch_bytes(rmax, cmax) : 0;
    w->tcs = c.take<uint8_t>(w->tcs_bytes);
  }
  w->total = (c.off + 255) & ~size_t(255);
}

// parameter order = state_dict order of models.MolecularVAE (oracle/cfga_oracle.cfga_shapes)
struct PIdx {
  int EL, DL;
  int emb() const { return 0; }
  int e_wih(int l) const { return 1 + 4 * l; }
  int e_whh(int l) const { return 2 + 4 * l; }
  int e_bih(int l) const { return 3 + 4 * l; }
  int e_bhh(int l) const { return 4 + 4 * l; }
  int conv_w(int i) const { return 1 + 4 * EL + 2 * i; }
  int conv_b(int i) const { return 2 + 4 * EL + 2 * i; }
  int d1_w() const { return 7 + 4 * EL; }
  int d1_b() const { return 8 + 4 * EL; }
  int mu_w() const { return 9 + 4 * EL; }
  int mu_b() const { return 10 + 4 * EL; }
  int lv_w() const { return 11 + 4 * EL; }
  int lv_b() const { return 12 + 4 * EL; }
  int li_w() const { return 13 + 4 * EL; }
  int li_b() const { return 14 + 4 * EL; }
  int d_wih(int l) const { return 15 + 4 * EL + 4 * l; }
  int d_whh(int l) const { return 16 + 4 * EL + 4 * l; }
  int d_bih(int l) const { return 17 + 4 * EL + 4 * l; }
  int d_bhh(int l) const { return 18 + 4 * EL + 4 * l; }
  int fc_w() const { return 15 + 4 * EL + 4 * DL; }
  int fc_b() const { return 16 + 4 * EL + 4 * DL; }
};

// ---------------------------------------------------------------------------------------------------------
// kernels specific to this path
// ---------------------------------------------------------------------------------------------------------
// cols[(b*L + l)][ci*KS + k] = x(b, ci, l + k);  x(b, ci, p) = x[b*sb + ci*sc + p*sp];  columns >= Cin*KS are zero.
template <typename TIN, typename TA>
__global__ void im2col_kernel(const TIN* __restrict__ x, long long sb, long long sc, long long sp, int B, int Cin, int L,
                              int Kp, TA* __restrict__ cols) {
  const long long total = (long long)B * L * Kp;
  const int K = Cin * KS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % Kp);
    const long long r = i / Kp;
    const int l = (int)(r % L);
    const long long b = r / L;
    float v = 0.f;
    if (col < K) {
      const int ci = col / KS, k = col - ci * KS;
      v = to_f32<TIN>(x[b * sb + ci * sc + (long long)(l + k) * sp]);
    }
    cols[i] = from_f32<TA>(v);
  }
}
// gather-form transpose of im2col: dx(b, ci, p) = sum_{k, 0 <= p-k < L} dcols[(b*L + p-k)][ci*KS + k]
template <typename TA, typename TOUT>
__global__ void col2im_kernel(const TA* __restrict__ dcols, int B, int Cin, int Lin, int L, int Kp, TOUT* __restrict__ dx,
                              long long sb, long long sc, long long sp) {
  const long long total = (long long)B * Cin * Lin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int p = (int)(r % Lin);
    const long long b = r / Lin;
    float s = 0.f;
    const int k0 = p - (L - 1) > 0 ? p - (L - 1) : 0, k1 = p < KS - 1 ? p : KS - 1;
    for (int k = k0; k <= k1; ++k) s += to_f32<TA>(dcols[(b * L + (p - k)) * Kp + ci * KS + k]);
    dx[b * sb + ci * sc + (long long)p * sp] = from_f32<TOUT>(s);
  }
}
__global__ void selu_inplace_kernel(float* __restrict__ a, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    a[i] = simt::selu_f(a[i]);
}
// Flatten of (B, C3, L3) (models.py:6-10,132): flat[b][co*L + l] = h[(b*L + l)*Cc + co]; inverse when `back`.
__global__ void flatten_kernel(float* __restrict__ h, float* __restrict__ flat, int B, int L, int Cc, int back) {
  const long long total = (long long)B * L * Cc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cc);
    const long long r = i / Cc;
    const int l = (int)(r % L);
    const long long b = r / L;
    const long long f = b * ((long long)L * Cc) + (long long)co * L + l;
    if (back) h[i] = flat[f]; else flat[f] = h[i];
  }
}
// da = dh * selu'(h)  (from the SELU output), fp32 [M][Cout] -> operand type
template <typename TA>
__global__ void selu_bwd_to_kernel(const float* __restrict__ h, const float* __restrict__ dh, TA* __restrict__ da, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    da[i] = from_f32<TA>(dh[i] * simt::selu_grad_from_out(h[i]));
}
// conv weight (Cout, Cin, KS) fp32 -> [Cout][Kp] operand (pad columns zero) and back
template <typename TA>
__global__ void pad_cols_kernel(const float* __restrict__ src, int rows, int cols, TA* __restrict__ dst, int cols_p) {
  const long long total = (long long)rows * cols_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols_p);
    const long long r = i / cols_p;
    dst[i] = from_f32<TA>(c < cols ? src[r * cols + c] : 0.f);
  }
}
// sum over time of dG [T][Bp][W] -> fp32 [Bp][W]
template <typename TA>
__global__ void time_sum_kernel(const TA* __restrict__ dG, int T, long long slab, float* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= slab) return;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += to_f32<TA>(dG[(long long)t * slab + idx]);
  out[idx] = s;
}
__global__ void copy_prefix2_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// padded W_hh [4Hp][Hp] (gate-major rows) -> fused-cell tile order: row n = tile*256 + gate*64 + j  <-  gate*Hp + tile*64 + j
__global__ void lstm_cell_weights_kernel(const __nv_bfloat16* __restrict__ src, int Hp, __nv_bfloat16* __restrict__ dst) {
  const long long total = 4ll * Hp * Hp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Hp);
    const int n = (int)(i / Hp);
    const int tile = n / 256, g = (n / 64) & 3, j = n & 63;
    dst[i] = src[((long long)g * Hp + tile * 64 + j) * Hp + c];
  }
}
// Off by default: with H = 1024 the step GEMM has only 512 tiles and the 8 epilogue warps of a CTA become the bottleneck
// once they also run the full-precision cell math (measured 152 ms/step either way at B = 4096); the GRU variant of the
// same epilogue pays off in the MOSES path (moses.cu), where tanh.approx gates are within budget.
#ifndef MVAE_LSTM_GATE_X8_DEFAULT
#define MVAE_LSTM_GATE_X8_DEFAULT 2
#endif
bool cell_fused_enabled() {
  const char* e = getenv("MVAE_LSTM_CELL_FUSED");
  return e ? atoi(e) != 0 : false;
}

// ---------------------------------------------------------------------------------------------------------
// per-step LSTM engine (one layer): one tcgen05 GEMM (h_{t-1} W_hh^T) per step; in bf16 mode the LSTM cell runs in the
// GEMM's epilogue (umma_gemm.h mvae_umma_cell, lstm = 1), in fp32 check mode a separate cell kernel follows the SGEMM
// ---------------------------------------------------------------------------------------------------------
// bf16 mode, LSTM cell kernels with eight units per thread (simt_kernels.cuh).  MVAE_LSTM_GATE_X8: 0 = the one-unit-per-thread
// kernels, 1 = x8 with exact expf / tanhf, 2 (default) = x8 with ex2 / rcp gate math in layers of H >= 256 (the decoder),
// 3 = ex2 / rcp gate math everywhere.  Returns 0 (scalar kernel), 1 (x8 exact) or 2 (x8 fast).  fp32 check mode: always 0.
// Measured at B = 4096 (tools/bench_cfga.py, profiles/r02_cfga_gate_variants.txt): 142.7 / 133.1 / 129.7 ms per step for
// 0 / 1 / 2.  The encoder keeps the exact math: its gradients sit at 0.0090-0.0095 of the 1e-2 bf16 budget from the bf16
// rounding of operands alone (tools/cfga_margin.py), mode 2 leaves them there (0.0089-0.0094), mode 3 moves the B = 64 case
// to 0.0103.
template <typename TA>
int gate_x8_mode(int H, long long gi_tstride) {
  if (sizeof(TA) != 2 || (H & 7) || (gi_tstride & 7)) return 0;
  static const int mode = [] { const char* e = getenv("MVAE_LSTM_GATE_X8"); return e ? atoi(e) : MVAE_LSTM_GATE_X8_DEFAULT; }();
  if (mode <= 0) return 0;
  if (mode == 1) return 1;
  return (mode >= 3 || H >= 256) ? 2 : 1;
}
template <typename TA, typename TG>
int lstm_fwd(const ADims& d, const AWS& w, cudaStream_t st, const TG* gi, long long gi_tstride, const TA* Whh, TA* hs, TA* sv,
             int H, const void* WhhC = nullptr) {
  const int Bp = d.Bp, T = d.T;
  const size_t slab = (size_t)Bp * H;
  RC(memset_async(hs, slab * sizeof(TA), st));
  RC(memset_async(w.c, slab * 4, st));
  if constexpr (sizeof(TA) == 2) {
    if (WhhC && cell_fused_enabled()) {
      for (int t = 0; t < T; ++t) {
        mvae_umma_operand a{hs + t * slab, 0, Bp, H, H, 1, 0, 0, 0};
        mvae_umma_operand b{WhhC, 0, 4ll * H, H, H, 1, 0, 0, 0};
        mvae_umma_out o{w.gh, 4ll * H, 0, 0, nullptr, 0};
        mvae_umma_cell c{};
        c.gates = 4; c.H = H; c.lstm = 1; c.gi = gi + (size_t)t * gi_tstride; c.gi_f32 = sizeof(TG) == 4 ? 1 : 0; c.cstate = w.c;
        c.out_a = hs + (t + 1) * slab; c.ld_a = H; c.sv = sv ? sv + (size_t)t * Bp * 6 * H : nullptr;
        mvae_count_launches(1);
        RC(mvae_umma_gemm(&a, &b, &o, Bp, 4 * H, H, 256, 1, 0, w.err_flag, st, nullptr, &c));
      }
      return MVAE_OK;
    }
  }
  const int gate_grid = (int)ceil_div64((long long)slab, 256);
  const int x8 = gate_x8_mode<TA>(H, gi_tstride);
  for (int t = 0; t < T; ++t) {
    RC(gemm<TA>(w.err_flag, st, hs + t * slab, H, false, Whh, H, true, w.gh, 4 * H, false, Bp, 4 * H, H, nullptr, false, 1));
    if constexpr (sizeof(TA) == 2) {
      if (x8) {
        const int g8 = (int)ceil_div64((long long)(slab / 8), 256);
        if (x8 == 2)
          simt::lstm_gate_fwd_x8_kernel<TG, true><<<g8, 256, 0, st>>>(gi + (size_t)t * gi_tstride, w.gh, w.c, hs + (t + 1) * slab,
                                                                      sv ? sv + (size_t)t * Bp * 6 * H : nullptr, Bp, H);
        else
          simt::lstm_gate_fwd_x8_kernel<TG, false><<<g8, 256, 0, st>>>(gi + (size_t)t * gi_tstride, w.gh, w.c, hs + (t + 1) * slab,
                                                                       sv ? sv + (size_t)t * Bp * 6 * H : nullptr, Bp, H);
        KCHECK();
        continue;
      }
    }
    simt::lstm_gate_fwd_kernel<TA, TG><<<gate_grid, 256, 0, st>>>(gi + (size_t)t * gi_tstride, w.gh, w.c, hs + (t + 1) * slab,
                                                                  sv ? sv + (size_t)t * Bp * 6 * H : nullptr, Bp, H);
    KCHECK();
  }
  return MVAE_OK;
}
template <typename TA>
int lstm_bwd(const ADims& d, const AWS& w, cudaStream_t st, const TA* Whh, const TA* sv, const TA* dX, TA* dG, int H) {
  const int Bp = d.Bp, T = d.T;
  const size_t slab = (size_t)Bp * H;
  RC(memset_async(w.dh_carry, slab * 4, st));
  RC(memset_async(w.dc_carry, slab * 4, st));
  const int gate_grid = (int)ceil_div64((long long)slab, 256);
  const int x8 = gate_x8_mode<TA>(H, 0);
  for (int t = T - 1; t >= 0; --t) {
    TA* dGt = dG + (size_t)t * Bp * 4 * H;
    bool done = false;
    if constexpr (sizeof(TA) == 2) {
      if (x8) {
        simt::lstm_gate_bwd_x8_kernel<<<(int)ceil_div64((long long)(slab / 8), 256), 256, 0, st>>>(
            sv + (size_t)t * Bp * 6 * H, dX + t * slab, w.dh_carry, w.dc_carry, dGt, Bp, H);
        done = true;
      }
    }
    if (!done)
      simt::lstm_gate_bwd_kernel<TA><<<gate_grid, 256, 0, st>>>(sv + (size_t)t * Bp * 6 * H, dX + t * slab, w.dh_carry,
                                                                w.dc_carry, dGt, Bp, H);
    KCHECK();
    if (t > 0)   // dh_{t-1} (recurrent part) = dG_t W_hh
      RC(gemm<TA>(w.err_flag, st, dGt, 4 * H, false, Whh, H, false, w.dh_carry, H, false, Bp, H, 4 * H, nullptr, false, 1));
  }
  return MVAE_OK;
}

template <typename TA>
int prep_weights(const ADims& d, const AWS& w, const float* const* P, cudaStream_t st, bool with_encoder) {
  const PIdx ix{d.EL, d.DL};
  const int EH = d.EH, EHp = d.EHp, DH = d.DH;
  RC(memset_async(w.err_flag, 4, st));
  if (with_encoder) {
    for (int l = 0; l < d.EL; ++l) {
      simt::pad_gates4_kernel<TA><<<grid_for(4ll * EHp * EHp), 256, 0, st>>>(P[ix.e_whh(l)], EH, EH, (TA*)w.Whh_e[l], EHp, EHp); KCHECK();
      if (l >= 1) {
        simt::pad_gates4_kernel<TA><<<grid_for(4ll * EHp * EHp), 256, 0, st>>>(P[ix.e_wih(l)], EH, EH, (TA*)w.Wih_e[l], EHp, EHp); KCHECK();
      }
      simt::pad_bias4_sum_kernel<<<ceil_div(4 * EHp, 256), 256, 0, st>>>(P[ix.e_bih(l)], P[ix.e_bhh(l)], EH, w.bsum_e[l], EHp); KCHECK();
      if constexpr (sizeof(TA) == 2) {
        lstm_cell_weights_kernel<<<grid_for(4ll * EHp * EHp), 256, 0, st>>>((const TA*)w.Whh_e[l], EHp, (TA*)w.WhhC_e[l]); KCHECK();
      }
    }
    const int Kc[3] = {d.K1, d.K2, d.K3}, Kcp[3] = {d.K1p, d.K2p, d.K3p}, Co[3] = {C1, C2, C3};
    for (int i = 0; i < 3; ++i) {
      pad_cols_kernel<TA><<<grid_for((long long)Co[i] * Kcp[i]), 256, 0, st>>>(P[ix.conv_w(i)], Co[i], Kc[i], (TA*)w.Wc[i], Kcp[i]); KCHECK();
    }
  }
  for (int l = 0; l < d.DL; ++l) {
    simt::pad_gates4_kernel<TA><<<grid_for(4ll * DH * DH), 256, 0, st>>>(P[ix.d_whh(l)], DH, DH, (TA*)w.Whh_d[l], DH, DH); KCHECK();
    if (l >= 1) {
      simt::pad_gates4_kernel<TA><<<grid_for(4ll * DH * DH), 256, 0, st>>>(P[ix.d_wih(l)], DH, DH, (TA*)w.Wih_d[l], DH, DH); KCHECK();
    }
    simt::pad_bias4_sum_kernel<<<ceil_div(4 * DH, 256), 256, 0, st>>>(P[ix.d_bih(l)], P[ix.d_bhh(l)], DH, w.bsum_d[l], DH); KCHECK();
    if constexpr (sizeof(TA) == 2) {
      lstm_cell_weights_kernel<<<grid_for(4ll * DH * DH), 256, 0, st>>>((const TA*)w.Whh_d[l], DH, (TA*)w.WhhC_d[l]); KCHECK();
    }
  }
  simt::pad_matrix_kernel<TA><<<grid_for((long long)d.CP * DH), 256, 0, st>>>(P[ix.fc_w()], d.C, DH, (TA*)w.Wfc, d.CP, DH); KCHECK();
  simt::pad_matrix_kernel<float><<<1, 64, 0, st>>>(P[ix.fc_b()], 1, d.C, w.bfc, 1, d.CP); KCHECK();
  return MVAE_OK;
}

// conv i (0..2) forward: cols <- im2col(x); h = selu(cols W^T + b)   ([B*L][Cout] fp32)
template <typename TA, typename TIN>
int conv_fwd(const ADims& d, const AWS& w, const float* const* P, cudaStream_t st, int i, const TIN* x, long long sb,
             long long sc, long long sp, int Cin, int L) {
  const PIdx ix{d.EL, d.DL};
  const int Kp[3] = {d.K1p, d.K2p, d.K3p}, Co[3] = {C1, C2, C3};
  const long long M = (long long)d.B * L;
  im2col_kernel<TIN, TA><<<grid_for(M * Kp[i]), 256, 0, st>>>(x, sb, sc, sp, d.B, Cin, L, Kp[i], (TA*)w.cols[i]); KCHECK();
  RC(gemm<TA>(w.err_flag, st, (const TA*)w.cols[i], Kp[i], false, (const TA*)w.Wc[i], Kp[i], true, w.h[i], Co[i], false, (int)M,
              Co[i], Kp[i], P[ix.conv_b(i)], false, 1));
  selu_inplace_kernel<<<grid_for(M * Co[i]), 256, 0, st>>>(w.h[i], M * Co[i]); KCHECK();
  return MVAE_OK;
}

// forward up to the logits.  decode_only: start from w.z (no encoder, no saved activations needed for BPTT).
template <typename TA>
int run_forward(const ADims& d, const AWS& w, const float* const* P, const uint8_t* ids, const float* eps, cudaStream_t st,
                bool decode_only, bool save, bool fuse_head = false) {
  const PIdx ix{d.EL, d.DL};
  const int B = d.B, Bp = d.Bp, T = d.T, C = d.C, E = d.E, EH = d.EH, EHp = d.EHp, Z = d.Z, DH = d.DH;
  const int TB = T * Bp;
  if (!decode_only) {
    RC(memset_async(w.bce_sum, 8, st)); RC(memset_async(w.kl_sum, 8, st)); RC(memset_async(w.hit_count, (size_t)B * 4, st));
    // ---- encoder LSTM: layer-0 projection = table E W_ih^T + (b_ih + b_hh), gathered by token id
    RC(memset_async(w.TBLe, (size_t)d.CP * 4 * EHp * 4, st));
    for (int g = 0; g < 4; ++g)
      RC(sg(st, P[ix.emb()], E, 1, P[ix.e_wih(0)] + (size_t)g * EH * E, 1, E, w.TBLe + (size_t)g * EHp, 4 * EHp, C, EH, E,
            w.bsum_e[0] + (size_t)g * EHp, simt::ACT_NONE, 0));
    gather_rows_kernel<TA><<<grid_for((long long)TB * 4 * EHp), 256, 0, st>>>(w.TBLe, 4 * EHp, ids, T, nullptr, B, Bp, T, (TA*)w.gi); KCHECK();
    for (int l = 0; l < d.EL; ++l) {
      if (l >= 1)
        RC(gemm<TA>(w.err_flag, st, (const TA*)w.hs_e[l - 1] + (size_t)Bp * EHp, EHp, false, (const TA*)w.Wih_e[l], EHp, true, w.gi,
                    4 * EHp, true, TB, 4 * EHp, EHp, w.bsum_e[l], false, 1));
      RC((lstm_fwd<TA, TA>(d, w, st, (const TA*)w.gi, (long long)Bp * 4 * EHp, (const TA*)w.Whh_e[l], (TA*)w.hs_e[l],
                           (TA*)w.sv_e[l], EHp, d.bf16 ? w.WhhC_e[l] : nullptr)));
    }
    // ---- convolutions: channels = the T positions, length axis = the LSTM features (models.py:129-131)
    const TA* enc_out = (const TA*)w.hs_e[d.EL - 1] + (size_t)Bp * EHp;   // [T][Bp][EHp]
    RC((conv_fwd<TA, TA>(d, w, P, st, 0, enc_out, EHp, (long long)Bp * EHp, 1, T, d.L1)));
    RC((conv_fwd<TA, float>(d, w, P, st, 1, w.h[0], (long long)d.L1 * C1, 1, C1, C1, d.L2)));
    RC((conv_fwd<TA, float>(d, w, P, st, 2, w.h[1], (long long)d.L2 * C2, 1, C2, C2, d.L3)));
    flatten_kernel<<<grid_for((long long)B * d.FLAT), 256, 0, st>>>(w.h[2], w.flat, B, d.L3, C3, 0); KCHECK();
    // ---- dense_1 + SELU, Lambda (models.py:89-94,133)
    RC(sg(st, w.flat, d.FLAT, 1, P[ix.d1_w()], 1, d.FLAT, w.h4, D1, B, D1, d.FLAT, P[ix.d1_b()], simt::ACT_SELU, 0));
    RC(sg(st, w.h4, D1, 1, P[ix.mu_w()], 1, D1, w.mu, Z, B, Z, D1, P[ix.mu_b()], simt::ACT_NONE, 0));
    RC(sg(st, w.h4, D1, 1, P[ix.lv_w()], 1, D1, w.lv, Z, B, Z, D1, P[ix.lv_b()], simt::ACT_NONE, 0));
    simt::reparam_kl_kernel<<<grid_for((long long)B * Z, 256, 592), 256, 0, st>>>(w.mu, w.lv, eps, d.eps_scale, 1,
                                                                                  (long long)B * Z, w.z, w.kl_sum);
    KCHECK();
  }
  // ---- decoder: latent_input + SELU; time-invariant layer-0 projection (models.py:162-164)
  RC(sg(st, w.z, Z, 1, P[ix.li_w()], 1, Z, w.zr, Z, B, Z, Z, P[ix.li_b()], simt::ACT_SELU, 0));
  RC(memset_async(w.gi0, (size_t)Bp * 4 * DH * 4, st));
  RC(sg(st, w.zr, Z, 1, P[ix.d_wih(0)], 1, Z, w.gi0, 4 * DH, B, 4 * DH, Z, w.bsum_d[0], simt::ACT_NONE, 0));
  for (int l = 0; l < d.DL; ++l) {
    TA* sv = save ? (TA*)w.sv_d[l] : nullptr;
    if (l == 0) {
      RC((lstm_fwd<TA, float>(d, w, st, w.gi0, 0, (const TA*)w.Whh_d[0], (TA*)w.hs_d[0], sv, DH, d.bf16 ? w.WhhC_d[0] : nullptr)));
    } else {
      RC(gemm<TA>(w.err_flag, st, (const TA*)w.hs_d[l - 1] + (size_t)Bp * DH, DH, false, (const TA*)w.Wih_d[l], DH, true, w.gi,
                  4 * DH, true, TB, 4 * DH, DH, w.bsum_d[l], false, 1));
      RC((lstm_fwd<TA, TA>(d, w, st, (const TA*)w.gi, (long long)Bp * 4 * DH, (const TA*)w.Whh_d[l], (TA*)w.hs_d[l], sv, DH,
                           d.bf16 ? w.WhhC_d[l] : nullptr)));
    }
  }
  if constexpr (sizeof(TA) == 2) {
    if (fuse_head) {
      // fused head (models.py:165 softmax + train.py:31-35 BCE): logits -> softmax -> BCE -> d(logits) inside the epilogue of the
      // vocabulary GEMM; the (T*B, C) logits / probabilities never reach HBM in the training step
      mvae_umma_operand a{(const TA*)w.hs_d[d.DL - 1] + (size_t)Bp * DH, 0, (long long)TB, DH, DH, 1, 0, 0, 0};
      mvae_umma_operand b{w.Wfc, 0, d.CP, DH, DH, 1, 0, 0, 0};
      mvae_umma_out o{w.logits, d.CP, 0, 0, w.bfc, 0};
      mvae_umma_head h{ids, B, Bp, T, d.C, d.max_len / ((float)B * (float)T * (float)d.C), w.dlogits, w.bce_sum, w.hit_count};
      mvae_count_launches(1);
      return mvae_umma_gemm(&a, &b, &o, TB, d.CP, DH, 64, 1, 0, w.err_flag, st, &h);
    }
  }
  RC(gemm<TA>(w.err_flag, st, (const TA*)w.hs_d[d.DL - 1] + (size_t)Bp * DH, DH, false, (const TA*)w.Wfc, DH, true, w.logits, d.CP,
              false, TB, d.CP, DH, w.bfc, false, 1, 64));
  return MVAE_OK;
}

// one LSTM layer's parameter gradients from dG [T][Bp][4H]:  dW_hh = dG^T h_{t-1},  db_ih = db_hh = colsum(dG)
template <typename TA>
int lstm_param_grads(const ADims& d, const AWS& w, cudaStream_t st, const TA* dG, const TA* hs, int H, int Hreal, float* gWhh,
                     float* gbih, float* gbhh, const TA* X, int XH, int XHreal, float* gWih) {
  const int TB = d.T * d.Bp;
  const int wsplits = d.bf16 ? max(1, min(64, (148 * 2) / (ceil_div(4 * H, 128) * ceil_div(H, 256)))) : 64;
  const bool padded = H != Hreal;
  float* out = padded ? w.dW_p : gWhh;
  RC(memset_async(out, (size_t)4 * H * H * 4, st));
  RC(gemm<TA>(w.err_flag, st, dG, 4 * H, true, hs, H, false, out, H, false, 4 * H, H, TB, nullptr, true, wsplits, 256));
  if (padded) { simt::unpad_gates4_kernel<<<grid_for(4ll * Hreal * Hreal), 256, 0, st>>>(w.dW_p, H, H, gWhh, Hreal, Hreal); KCHECK(); }
  RC(memset_async(w.csum, (size_t)4 * H * 4, st));
  RC(simt::colsum<TA>(st, dG, TB, 4 * H, 4 * H, w.csum)); mvae_count_launches(1);
  simt::unpad_bias4_dup_kernel<<<ceil_div(4 * Hreal, 256), 256, 0, st>>>(w.csum, H, gbih, gbhh, Hreal); KCHECK();
  if (X) {
    const bool xpad = padded || XH != XHreal;
    float* o2 = xpad ? w.dW_p : gWih;
    RC(memset_async(o2, (size_t)4 * H * XH * 4, st));
    RC(gemm<TA>(w.err_flag, st, dG, 4 * H, true, X, XH, false, o2, XH, false, 4 * H, XH, TB, nullptr, true, wsplits, 256));
    if (xpad) { simt::unpad_gates4_kernel<<<grid_for(4ll * Hreal * XHreal), 256, 0, st>>>(w.dW_p, H, XH, gWih, Hreal, XHreal); KCHECK(); }
  }
  return MVAE_OK;
}

// conv i backward.  dh: fp32 [B*L][Cout] gradient wrt the SELU output; writes dW, db and dcols (over w.cols[i]).
template <typename TA>
int conv_bwd(const ADims& d, const AWS& w, const float* const* P, float* const* G, cudaStream_t st, int i, const float* dh, int L) {
  (void)P;
  const PIdx ix{d.EL, d.DL};
  const int Kc[3] = {d.K1, d.K2, d.K3}, Kp[3] = {d.K1p, d.K2p, d.K3p}, Co[3] = {C1, C2, C3};
  const long long M = (long long)d.B * L;
  TA* da = (TA*)w.da;
  selu_bwd_to_kernel<TA><<<grid_for(M * Co[i]), 256, 0, st>>>(w.h[i], dh, da, M * Co[i]); KCHECK();
  // dW [Cout][K] = da^T cols
  RC(memset_async(w.dWc_p, (size_t)Co[i] * Kp[i] * 4, st));
  const int splits = d.bf16 ? (int)max(1ll, min(64ll, M / 2048)) : (int)max(1ll, min(64ll, M / 256));
  RC(gemm<TA>(w.err_flag, st, da, Co[i], true, (const TA*)w.cols[i], Kp[i], false, w.dWc_p, Kp[i], false, Co[i], Kp[i], (int)M,
              nullptr, true, splits, 256));
  simt::unpad_matrix_kernel<<<grid_for((long long)Co[i] * Kc[i]), 256, 0, st>>>(w.dWc_p, Kp[i], G[ix.conv_w(i)], Co[i], Kc[i]); KCHECK();
  RC(memset_async(w.csum, (size_t)256 * 4, st));
  RC(simt::colsum<TA>(st, da, M, Co[i], Co[i], w.csum)); mvae_count_launches(1);
  copy_prefix2_kernel<<<1, 128, 0, st>>>(w.csum, G[ix.conv_b(i)], Co[i]); KCHECK();
  // dcols [M][Kp] = da W   (overwrites cols, no longer needed)
  RC(gemm<TA>(w.err_flag, st, da, Co[i], false, (const TA*)w.Wc[i], Kp[i], false, w.cols[i], Kp[i], true, (int)M, Kp[i], Co[i],
              nullptr, false, 1));
  return MVAE_OK;
}

// expects w.dlogits filled.  kl_internal: add the swapped-KL gradient; ext_dmu / ext_dlv: optional upstream grads.
template <typename TA>
int run_backward(const ADims& d, const AWS& w, const float* const* P, float* const* G, const uint8_t* ids, const float* eps,
                 cudaStream_t st, bool kl_internal, const float* ext_dmu, const float* ext_dlv) {
  const PIdx ix{d.EL, d.DL};
  const int B = d.B, Bp = d.Bp, T = d.T, C = d.C, CP = d.CP, E = d.E, EH = d.EH, EHp = d.EHp, Z = d.Z, DH = d.DH;
  const int TB = T * Bp;
  const TA* dlog = (const TA*)w.dlogits;
  TA* dG = (TA*)w.dG;
  TA* dX = (TA*)w.dX;
  // ---- head: dX = dlogits W_fc ; dW_fc = dlogits^T h_top ; db = colsum(dlogits)
  RC(gemm<TA>(w.err_flag, st, dlog, CP, false, (const TA*)w.Wfc, DH, false, dX, DH, true, TB, DH, CP, nullptr, false, 1));
  RC(memset_async(w.dWfc_p, (size_t)CP * DH * 4, st));
  RC(gemm<TA>(w.err_flag, st, dlog, CP, true, (const TA*)w.hs_d[d.DL - 1] + (size_t)Bp * DH, DH, false, w.dWfc_p, DH, false, CP, DH,
              TB, nullptr, true, d.bf16 ? 148 : 64, 256));
  simt::unpad_matrix_kernel<<<grid_for((long long)C * DH), 256, 0, st>>>(w.dWfc_p, DH, G[ix.fc_w()], C, DH); KCHECK();
  RC(memset_async(w.csum, (size_t)256 * 4, st));
  RC(simt::colsum<TA>(st, dlog, TB, CP, CP, w.csum)); mvae_count_launches(1);
  copy_prefix2_kernel<<<1, 128, 0, st>>>(w.csum, G[ix.fc_b()], C); KCHECK();
  // ---- decoder LSTM stack
  for (int l = d.DL - 1; l >= 0; --l) {
    const TA* hs = (const TA*)w.hs_d[l];
    RC(lstm_bwd<TA>(d, w, st, (const TA*)w.Whh_d[l], (const TA*)w.sv_d[l], dX, dG, DH));
    const TA* X = l >= 1 ? (const TA*)w.hs_d[l - 1] + (size_t)Bp * DH : nullptr;
    RC(lstm_param_grads<TA>(d, w, st, dG, hs, DH, DH, G[ix.d_whh(l)], G[ix.d_bih(l)], G[ix.d_bhh(l)], X, DH, DH,
                            l >= 1 ? G[ix.d_wih(l)] : nullptr));
    if (l >= 1) {
      RC(gemm<TA>(w.err_flag, st, dG, 4 * DH, false, (const TA*)w.Wih_d[l], DH, false, dX, DH, true, TB, DH, 4 * DH, nullptr, false, 1));
    } else {
      // time-invariant input: dW_ih0 = (sum_t dG)^T zr ; dzr = (sum_t dG) W_ih0
      const long long slab = (long long)Bp * 4 * DH;
      time_sum_kernel<TA><<<(unsigned)ceil_div64(slab, 256), 256, 0, st>>>(dG, T, slab, w.dgisum); KCHECK();
      RC(sg_wgrad(st, w.dgisum, 1, 4 * DH, w.zr, Z, 1, G[ix.d_wih(0)], Z, 4 * DH, Z, B));
      RC(sg(st, w.dgisum, 4 * DH, 1, P[ix.d_wih(0)], Z, 1, w.dzr, Z, B, Z, 4 * DH, nullptr, simt::ACT_NONE, 0));
    }
  }
  // ---- latent_input (+SELU)
  const long long nBZ = (long long)B * Z;
  simt::selu_bwd_kernel<<<grid_for(nBZ), 256, 0, st>>>(w.zr, w.dzr, w.da5, nBZ); KCHECK();
  RC(sg_wgrad(st, w.da5, 1, Z, w.z, Z, 1, G[ix.li_w()], Z, Z, Z, B));
  RC(memset_async(G[ix.li_b()], (size_t)Z * 4, st));
  RC(simt::colsum<float>(st, w.da5, B, Z, Z, G[ix.li_b()])); mvae_count_launches(1);
  RC(sg(st, w.da5, Z, 1, P[ix.li_w()], Z, 1, w.dz, Z, B, Z, Z, nullptr, simt::ACT_NONE, 0));
  // ---- Lambda: reparametrisation + (swapped) KL
  simt::reparam_kl_bwd_kernel<<<grid_for(nBZ), 256, 0, st>>>(w.mu, w.lv, eps, d.eps_scale, 1, w.dz,
                                                            kl_internal ? 1.0f / (float)nBZ : 0.f, ext_dmu, ext_dlv, nBZ, w.dmu, w.dlv);
  KCHECK();
  RC(sg_wgrad(st, w.dmu, 1, Z, w.h4, D1, 1, G[ix.mu_w()], D1, Z, D1, B));
  RC(sg_wgrad(st, w.dlv, 1, Z, w.h4, D1, 1, G[ix.lv_w()], D1, Z, D1, B));
  RC(memset_async(G[ix.mu_b()], (size_t)Z * 4, st));
  RC(simt::colsum<float>(st, w.dmu, B, Z, Z, G[ix.mu_b()])); mvae_count_launches(1);
  RC(memset_async(G[ix.lv_b()], (size_t)Z * 4, st));
  RC(simt::colsum<float>(st, w.dlv, B, Z, Z, G[ix.lv_b()])); mvae_count_launches(1);
  RC(sg(st, w.dmu, Z, 1, P[ix.mu_w()], D1, 1, w.dh4, D1, B, D1, Z, nullptr, simt::ACT_NONE, 0));
  RC(sg(st, w.dlv, Z, 1, P[ix.lv_w()], D1, 1, w.dh4, D1, B, D1, Z, nullptr, simt::ACT_NONE, 1));
  const long long nBD = (long long)B * D1;
  simt::selu_bwd_kernel<<<grid_for(nBD), 256, 0, st>>>(w.h4, w.dh4, w.dh4, nBD); KCHECK();
  // ---- dense_1
  RC(sg_wgrad(st, w.dh4, 1, D1, w.flat, d.FLAT, 1, G[ix.d1_w()], d.FLAT, D1, d.FLAT, B));
  RC(memset_async(G[ix.d1_b()], (size_t)D1 * 4, st));
  RC(simt::colsum<float>(st, w.dh4, B, D1, D1, G[ix.d1_b()])); mvae_count_launches(1);
  RC(sg(st, w.dh4, D1, 1, P[ix.d1_w()], d.FLAT, 1, w.dflat, d.FLAT, B, d.FLAT, D1, nullptr, simt::ACT_NONE, 0));
  // ---- convolutions
  flatten_kernel<<<grid_for((long long)B * d.FLAT), 256, 0, st>>>(w.dhc, w.dflat, B, d.L3, C3, 1); KCHECK();
  RC(conv_bwd<TA>(d, w, P, G, st, 2, w.dhc, d.L3));
  col2im_kernel<TA, float><<<grid_for((long long)B * C2 * d.L2), 256, 0, st>>>((const TA*)w.cols[2], B, C2, d.L2, d.L3, d.K3p, w.dhc,
                                                                              (long long)d.L2 * C2, 1, C2); KCHECK();
  RC(conv_bwd<TA>(d, w, P, G, st, 1, w.dhc, d.L2));
  col2im_kernel<TA, float><<<grid_for((long long)B * C1 * d.L1), 256, 0, st>>>((const TA*)w.cols[1], B, C1, d.L1, d.L2, d.K2p, w.dhc,
                                                                              (long long)d.L1 * C1, 1, C1); KCHECK();
  RC(conv_bwd<TA>(d, w, P, G, st, 0, w.dhc, d.L1));
  // gradient wrt the encoder LSTM output, straight into the time-major dX (pads zero)
  RC(memset_async(dX, (size_t)TB * EHp * sizeof(TA), st));
  col2im_kernel<TA, TA><<<grid_for((long long)B * T * EH), 256, 0, st>>>((const TA*)w.cols[0], B, T, EH, d.L1, d.K1p, dX, EHp,
                                                                        (long long)Bp * EHp, 1); KCHECK();
  // ---- encoder LSTM stack
  for (int l = d.EL - 1; l >= 0; --l) {
    RC(lstm_bwd<TA>(d, w, st, (const TA*)w.Whh_e[l], (const TA*)w.sv_e[l], dX, dG, EHp));
    const TA* X = l >= 1 ? (const TA*)w.hs_e[l - 1] + (size_t)Bp * EHp : nullptr;
    RC(lstm_param_grads<TA>(d, w, st, dG, (const TA*)w.hs_e[l], EHp, EH, G[ix.e_whh(l)], G[ix.e_bih(l)], G[ix.e_bhh(l)], X, EHp, EH,
                            l >= 1 ? G[ix.e_wih(l)] : nullptr));
    if (l >= 1) {
      RC(gemm<TA>(w.err_flag, st, dG, 4 * EHp, false, (const TA*)w.Wih_e[l], EHp, false, dX, EHp, true, TB, EHp, 4 * EHp, nullptr,
                  false, 1));
    } else {
      // table gradient dTBL [CP][4EHp] = onehot^T dG, then through E and W_ih_l0
      onehot_rows_kernel<TA><<<grid_for((long long)TB * CP), 256, 0, st>>>(ids, T, B, Bp, T, CP, (TA*)w.OH); KCHECK();
      RC(memset_async(w.dTBL, (size_t)CP * 4 * EHp * 4, st));
      RC(gemm<TA>(w.err_flag, st, (const TA*)w.OH, CP, true, dG, 4 * EHp, false, w.dTBL, 4 * EHp, false, CP, 4 * EHp, TB, nullptr, true,
                  d.bf16 ? 24 : 64, 256));
      for (int g = 0; g < 4; ++g) {
        // dW_ih0[g*EH + j][e] = sum_c dTBL[c][g*EHp + j] E[c][e]
        RC(sg(st, w.dTBL + (size_t)g * EHp, 1, 4 * EHp, P[ix.emb()], E, 1, G[ix.e_wih(0)] + (size_t)g * EH * E, E, EH, E, C, nullptr,
              simt::ACT_NONE, 0));
        // dE[c][e] += sum_j dTBL[c][g*EHp + j] W_ih0[g*EH + j][e]
        RC(sg(st, w.dTBL + (size_t)g * EHp, 4 * EHp, 1, P[ix.e_wih(0)] + (size_t)g * EH * E, E, 1, G[ix.emb()], E, C, E, EH, nullptr,
              simt::ACT_NONE, g > 0 ? 1 : 0));
      }
    }
  }
  return MVAE_OK;
}

template <typename TA>
int head_fused(const ADims& d, const AWS& w, const uint8_t* ids, float* probs, bool want_dlogits, cudaStream_t st) {
  const long long rows = (long long)d.T * d.Bp;
  const float gscale = d.max_len / ((float)d.B * (float)d.T * (float)d.C);
  simt::head_softmax_bce_kernel<TA><<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(
      w.logits, d.CP, d.C, ids, d.B, d.Bp, d.T, gscale, want_dlogits ? (TA*)w.dlogits : nullptr, probs, w.bce_sum, w.hit_count);
  KCHECK();
  return MVAE_OK;
}

int finalize(const ADims& d, const AWS& w, float* out_scalars, float* mu_out, float* lv_out, cudaStream_t st) {
  if (out_scalars) {
    simt::finalize_scalars_kernel<<<1, 256, 0, st>>>(w.bce_sum, w.kl_sum, w.hit_count, d.B, d.T,
                                                    (double)d.max_len / ((double)d.B * d.T * d.C), 1.0 / ((double)d.B * d.Z),
                                                    out_scalars, w.err_flag);
    KCHECK();
  }
  const size_t n = (size_t)d.B * d.Z * 4;
  if (mu_out) { mvae_count_launches(1); MVAE_CUDA_CHECK(cudaMemcpyAsync(mu_out, w.mu, n, cudaMemcpyDeviceToDevice, st)); }
  if (lv_out) { mvae_count_launches(1); MVAE_CUDA_CHECK(cudaMemcpyAsync(lv_out, w.lv, n, cudaMemcpyDeviceToDevice, st)); }
  return MVAE_OK;
}

template <typename TA>
int elbo_step_t(const ADims& d, const AWS& w, const float* const* P, float* const* G, const uint8_t* ids, const float* eps,
                float* out_scalars, float* mu_out, float* lv_out, cudaStream_t st) {
  RC(prep_weights<TA>(d, w, P, st, true));
  const char* fe = getenv("MVAE_FUSED_HEAD");
  const bool fuse = sizeof(TA) == 2 && d.CP == 64 && (fe ? atoi(fe) != 0 : true);
  RC(run_forward<TA>(d, w, P, ids, eps, st, false, true, fuse));
  if (!fuse) RC(head_fused<TA>(d, w, ids, nullptr, true, st));
  RC(run_backward<TA>(d, w, P, G, ids, eps, st, true, nullptr, nullptr));
  return finalize(d, w, out_scalars, mu_out, lv_out, st);
}

int check_ws(const mvae_cfga_desc* desc, void* ws, size_t ws_bytes, ADims* d, AWS* w) {
  RC(make_dims(desc, d));
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 255)) return MVAE_ERR_INVALID;
  carve(*d, ws, w);
  if (ws_bytes < w->total) return MVAE_ERR_WORKSPACE;
  g_tc = mvae_tc_ctx{(d->bf16 && tc_sgemm_enabled()) ? w->tcs : nullptr, w->tcs_bytes, w->err_flag};
  return MVAE_OK;
}

}  // namespace

extern "C" {

size_t mvae_cfga_workspace_bytes(const mvae_cfga_desc* desc) {
  ADims d; AWS w;
  if (make_dims(desc, &d) != MVAE_OK) return 0;
  carve(d, nullptr, &w);
  return w.total;
}

int mvae_cfga_elbo_step(const mvae_cfga_desc* desc, const float* const* params, float* const* grads, const uint8_t* ids,
                        const float* eps, float* out_scalars, float* mu_out, float* logvar_out, void* workspace,
                        size_t workspace_bytes, mvae_stream_t stream) {
  ADims d; AWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !grads || !ids || !eps) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return d.bf16 ? elbo_step_t<__nv_bfloat16>(d, w, params, grads, ids, eps, out_scalars, mu_out, logvar_out, st)
                : elbo_step_t<float>(d, w, params, grads, ids, eps, out_scalars, mu_out, logvar_out, st);
}

int mvae_cfga_elbo_step_graph_create(const mvae_cfga_desc* desc, const float* const* params, float* const* grads,
                                     const uint8_t* ids, const float* eps, float* out_scalars, float* mu_out,
                                     float* logvar_out, void* workspace, size_t workspace_bytes, mvae_graph** out_graph) {
  struct Ctx {
    const mvae_cfga_desc* desc; const float* const* params; float* const* grads; const uint8_t* ids; const float* eps;
    float *out_scalars, *mu_out, *logvar_out; void* ws; size_t ws_bytes;
  } c{desc, params, grads, ids, eps, out_scalars, mu_out, logvar_out, workspace, workspace_bytes};
  return mvae_capture_into_graph(
      [](void* p, cudaStream_t cs) {
        Ctx* c = static_cast<Ctx*>(p);
        return mvae_cfga_elbo_step(c->desc, c->params, c->grads, c->ids, c->eps, c->out_scalars, c->mu_out, c->logvar_out, c->ws,
                                   c->ws_bytes, reinterpret_cast<mvae_stream_t>(cs));
      },
      &c, out_graph);
}

int mvae_cfga_forward(const mvae_cfga_desc* desc, const float* const* params, const uint8_t* ids, const float* eps, float* probs,
                      float* mu, float* logvar, void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  ADims d; AWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !ids || !eps) return MVAE_ERR_INVALID;
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

int mvae_cfga_backward(const mvae_cfga_desc* desc, const float* const* params, float* const* grads, const uint8_t* ids,
                       const float* eps, const float* dprobs, const float* dmu, const float* dlogvar, void* workspace,
                       size_t workspace_bytes, mvae_stream_t stream) {
  ADims d; AWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !grads || !ids || !eps) return MVAE_ERR_INVALID;
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

int mvae_cfga_decode(const mvae_cfga_desc* desc, const float* const* params, const float* z, uint8_t* ids_out, float* probs_out,
                     void* workspace, size_t workspace_bytes, mvae_stream_t stream) {
  ADims d; AWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!params || !z || !ids_out) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  mvae_count_launches(1);
  MVAE_CUDA_CHECK(cudaMemcpyAsync(w.z, z, (size_t)d.B * d.Z * 4, cudaMemcpyDeviceToDevice, st));
  if (d.bf16) {
    RC(prep_weights<__nv_bfloat16>(d, w, params, st, false));
    RC(run_forward<__nv_bfloat16>(d, w, params, nullptr, nullptr, st, true, false));
  } else {
    RC(prep_weights<float>(d, w, params, st, false));
    RC(run_forward<float>(d, w, params, nullptr, nullptr, st, true, false));
  }
  const long long rows = (long long)d.T * d.Bp;
  simt::head_argmax_kernel<<<(unsigned)ceil_div64(rows * 32, 256), 256, 0, st>>>(w.logits, d.CP, d.C, d.B, d.Bp, d.T, ids_out);
  KCHECK();
  if (probs_out) {
    if (d.bf16) RC(head_fused<__nv_bfloat16>(d, w, ids_out, probs_out, false, st));
    else RC(head_fused<float>(d, w, ids_out, probs_out, false, st));
  }
  return MVAE_OK;
}

int mvae_cfga_read_error(const mvae_cfga_desc* desc, void* workspace, size_t workspace_bytes, int* flag, mvae_stream_t stream) {
  ADims d; AWS w;
  RC(check_ws(desc, workspace, workspace_bytes, &d, &w));
  if (!flag) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  MVAE_CUDA_CHECK(cudaMemcpyAsync(flag, w.err_flag, 4, cudaMemcpyDeviceToHost, st));
  MVAE_CUDA_CHECK(cudaStreamSynchronize(st));
  return MVAE_OK;
}

}  // extern "C"
