// Internal host interface of the persistent fused GRU recurrence (gru_rec.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

struct mvae_gru_rec_args {
  int backward;              // 0: forward sweep, 1: BPTT sweep
  int variant;               // 1: 64 units x 2 tiles x 2 stages per CTA; 2: 32 units x 4 tiles x 8 stages
  int Bp, Hp, T;             // Bp % 128 == 0, Hp % 64 == 0, Hp <= 512
  const __nv_bfloat16* W;    // fwd: padded W_hh [3Hp][Hp]; bwd: padded W_hh^T [Hp][3Hp]
  const __nv_bfloat16* gi;   // fwd: input projections incl. b_ih, [T][Bp][3Hp] (or [Bp][3Hp] with gi_tstride 0)
  long long gi_tstride;
  const float* bhh;          // fwd: padded b_hh [3Hp]
  __nv_bfloat16* hs;         // [(T+1)][Bp][Hp]; slab 0 = h0 (fwd writes slabs 1..T, bwd reads slabs 0..T-1)
  __nv_bfloat16* sv;         // [T][Bp][4Hp]
  const __nv_bfloat16* dX;   // bwd: [T][Bp][Hp]
  __nv_bfloat16* dG;         // bwd: [T][Bp][4Hp]
  unsigned int* counters;    // [Bp/128]
  int* err_flag;
  int a_box_rows;            // gru_rec2: rows per operand TMA box (0/128, 64, 32)
  int ones_col;              // gru_rec2 fwd: hidden-unit index whose h is forced to 1.0 (-1: none)
  int debug;                 // timing experiments only (gru_rec2): bit0 skip counter waits, bit1 de-share operand rows
  unsigned long long* trace; // optional debug timestamps [T][tiles per CTA][8] of CTA (0,0), may be null
  const float* h0;           // gru_rec2 fwd, optional: fp32 initial state [Bp][Hp]; hs slab 0 must hold its bf16 copy
  float* carry_out;          // gru_rec2 bwd, optional: fp32 [Bp][Hp] <- dh_0 * z_0 (add dgh_0 * W_hh for dL/dh0)
  // gru_rec2 fwd, optional token-table input projection: gi(row, t) = tbl[tok[t][row]] (+ gi when gi != null)
  const float* tbl;          // [V][3Hp] fp32 (r,z,n), biases folded in
  const unsigned char* tok;  // [T][Bp]
  int V;                     // <= 64
  // gru_rec2 fwd, optional: hlast[row][:] = fp32 state after step lens[row]-1, rows < nrows
  const int* lens; float* hlast; int nrows;
  // gru_rec2, optional (packed sequences, batch sorted by length descending): DEVICE array [Bp/256], tile j runs only its
  // first tile_T[j] time steps (what lies beyond is neither read nor written); null = every tile runs T steps
  const int* tile_T;
};

// rows (molecules) one cooperative launch can cover on a device with num_sms SMs (0: shape unsupported)
int mvae_gru_rec_max_rows(int Hp, int variant, int num_sms);
int mvae_gru_rec_launch(const mvae_gru_rec_args* a, cudaStream_t stream);
// 2-CTA / cluster-multicast version (gru_rec2.cu): Bp % 256 == 0, Hp in {256, 512}; a->bhh = n-gate slice of the
// padded b_hh, b_hr / b_hz already folded into gi.
int mvae_gru_rec2_launch(const mvae_gru_rec_args* a, int fast_gates, cudaStream_t stream);
int mvae_gru_rec2_max_clusters(int backward, int cluster);
