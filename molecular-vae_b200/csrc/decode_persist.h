// Internal host interface of the persistent decode kernel (decode_persist.cu): VAE.sample's whole decode loop
// (mosesvae.py:239-251) as ONE launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct mvae_decode_args {
  int B, Bp;                 // sequences, rows of the operand buffers (multiple of 128)
  int Hd, L;                 // decoder hidden size (multiple of 64), layers (1..4)
  int V;                     // vocabulary (<= 64: one tile of logits)
  int K0;                    // row length of layer 0's operand [onehot(64) | z (padded) | h]
  int max_len, eos, mode;    // steps 1 .. max_len-1; mode 0 greedy, 1 multinomial
  float inv_temp;
  unsigned long long seed; const unsigned long long* seed_dev;
  // per layer l: operand buffers xh[l][parity] (bf16 [Bp][ldx]: layer 0 ldx = K0 with h at column K0 - Hd, layers >= 1
  // ldx = 2 Hd with [x | h]), fp32 master states hm[l][parity] [Bp][Hd], permuted weights Wcat[l] ([4 Hd][ldx], tiles of
  // 64 units x (r, z, in, hn)) and biases bcat[l] [4 Hd]
  void* xh[4][2]; float* hm[4][2]; const void* Wcat[4]; const float* bcat[4];
  const void* Wfc; const float* bfc;   // vocabulary head: bf16 [64][Hd], fp32 [64]
  unsigned char* w_cur; unsigned char* x; int* end; unsigned char* done;   // as mvae_umma_sample
  unsigned int* counters;    // [(L + 1) * Bp / 128] completion counters, zeroed by the launcher
  unsigned int* sched;       // [(8 L + 1) * Bp / 128] unit order of one step, written by the launcher
  int* err_flag;
  int variant;               // 2: CTA pairs (tcgen05 cta_group::2, units of 256 rows; Bp % 256 == 0), else one CTA per unit of 128 rows
};

// workspace the caller must provide for counters + schedule (bytes)
size_t mvae_decode_persistent_scratch_bytes(int Bp, int L);
// 0 on success; MVAE_ERR_UNSUPPORTED when the shape is outside what the kernel handles (the caller keeps its per-step path)
int mvae_decode_persistent_launch(const mvae_decode_args* a, cudaStream_t stream);
