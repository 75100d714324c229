// Internal host interface of the tcgen05 GEMM (umma_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct mvae_umma_operand {
  const void* ptr;        // bf16, 16-byte aligned
  int mn_major;           // 0: stored [mn][k] (k contiguous); 1: stored [k][mn] (mn contiguous)
  long long mn;           // extent of the M (for A) or N (for B) dimension
  long long k;            // extent of the contraction dimension
  long long ld;           // elements between consecutive rows of the stored matrix
  int slabs;              // number of 2-D slabs (third TMA dimension), >= 1
  long long slab_stride;  // elements between slabs
  int slab;               // which slab this call uses
  int rb;                 // reserved (row-blocked operands are not supported by the TMA path); must be 0
};

// Optional fused "vocabulary head" epilogue (tile_n must be 64, splits 1): the logits of a row never leave the SM.
// Per row (t, b) = (row / Bp, row % Bp): softmax over the first C columns, max_len*BCE(mean) loss partial
// (train.py:31-35) accumulated into bce_sum, the gradient wrt the logits written as bf16 [M][64]
// (zero for pad rows / columns), and the per-molecule argmax hit count (train.py:110-112).
struct mvae_umma_head {
  const unsigned char* ids;   // u8 [B][T] targets
  int B, Bp, T, C;
  float gscale;               // max_len / (B*T*C)
  void* dlogits;              // bf16 [M][64]
  double* bce_sum;
  int* hit_count;             // [B]
  // mode 1: shifted cross entropy with ignore_index (mosesvae.py:190-197) instead of the BCE: row (t, b) is a target
  // position iff t + 1 < lens[b]; target token ids[b][t + 1] (ids rows of T entries); bce_sum receives sum(-log softmax[tgt]);
  // dlogits = (softmax - onehot) * rec_w / *Mcount on target rows, zeros elsewhere (C = vocabulary size <= 64)
  int mode;
  const int* lens;            // [B]
  const int* Mcount;          // device scalar: number of target positions of the batch
  float rec_w;
};

// Optional fused GRU-cell epilogue (decode / per-step engines): the GEMM's N dimension is laid out in tiles of 64 hidden
// units x G gate blocks ([gate][64 units] inside a tile, weights / bias permuted accordingly by the caller):
//   gates 3 (tile_n 192): acc = h_prev W_hh^T blocks [r | z | hn];  r,z,n input parts come from `gi` (bf16 [M][3H], (r,z,n))
//   gates 4 (tile_n 256): acc = [x | h_prev] [W_ih | W_hh]^T blocks [r | z | in | hn]  (r,z already summed over both inputs)
//   r = s(.), z = s(.), n = tanh(in + r * hn), h' = (1 - z) n + z h_prev     (torch.nn.GRU; bias added to every block)
// h_prev32 / h_next32: fp32 master state [M][H]; out_a / out_b: bf16 copies of h' (row strides ld_a / ld_b, out_b optional).
struct mvae_umma_cell {
  int gates;                  // 3 or 4
  int H;                      // hidden size (multiple of 64)
  const void* gi;             // gates == 3 only
  const float* h_prev32;
  float* h_next32;
  void* out_a; long long ld_a;
  void* out_b; long long ld_b;
  void* sv;                   // optional (training): bf16 [M][4H] saved (r, z, n, W_hn h + b_hn) for BPTT
  // LSTM variant (lstm = 1, gates = 4, tile_n 256): acc = h_prev W_hh^T blocks [i | f | g | o]; gi = x W_ih^T + b_ih + b_hh
  // ([M][4H], bf16, or fp32 when gi_f32);
  // c' = f c + i g (fp32 `cstate` [M][H], updated in place), h' = o tanh(c') -> out_a; sv (optional) bf16 [M][6H] =
  // (i, f, g, o, c_prev, tanh c').  h_prev32 / h_next32 / out_b are unused.
  int lstm;
  int gi_f32;
  float* cstate;
  // gates == 3 alternative to `gi` (sampler): gi(row) = tbl[tok[row]] + add[row], both fp32 with rows of 3H (r,z,n);
  // rows >= tok_rows use token 0.  (mosesvae.py:244-245: emb(w) | z through W_ih_l0, as a table look-up + per-sequence part)
  const float* tbl;
  const float* add;
  const unsigned char* tok;
  int tok_rows;
};

// Optional fused sampling epilogue of the vocabulary GEMM (tile_n 64, mosesvae.py:247-255): per row softmax(y / temp) ->
// argmax (mode 0) or inverse-CDF draw with the counter-based generator (mode 1) -> masked write + EOS bookkeeping.
struct mvae_umma_sample {
  int V, B, step, max_len, eos, mode;
  float inv_temp;
  unsigned long long seed;
  const unsigned long long* seed_dev;   // optional: read the seed from device memory (a captured CUDA graph can then be
                                        // replayed with fresh draws)
  unsigned char* w_cur;       // [B] token fed to the next step
  void* onehot_next;          // optional: bf16 [rows][oh_ld], columns [0,64) of row b <- one-hot of the sampled token (the token
  long long oh_ld;            // part of the next step's [onehot | z | h] layer-0 operand)
  unsigned char* x;           // [B][max_len]
  int* end;                   // [B]
  unsigned char* done;        // [B]
};

struct mvae_umma_out {
  void* ptr;          // fp32 or bf16 [M][ld]
  long long ld;
  int bf16;           // output element type
  int accumulate;     // fp32 only: D += A*B
  const float* bias;  // optional per-column bias (fp32)
  int rb;             // bf16 output only: row-blocked layout [M/32][ld/8][32][8] (consumed by the recurrence epilogues)
  int act;            // fp32 output, splits 1, no accumulate: 0 none, 1 SELU, 2 ReLU applied after the bias
};

// Optional packed-sequence skipping for GEMMs over time-major slabs ([T][rows_per_slab][.], rows_per_slab % 128 == 0):
// act[t] (DEVICE array) = number of leading rows of slab t that belong to sequences still running at step t.
//   mode 1: the M dimension runs over (t, row): output tiles that lie entirely in the inactive part of their slab are
//           skipped (their output rows keep whatever they held);
//   mode 2: the K dimension runs over (t, row): 64-row k-blocks entirely in the inactive part are skipped (the caller
//           guarantees that they would contribute zeros).
struct mvae_umma_varlen {
  const int* act;
  int rows_per_slab;
  int mode;
  // optional HOST copy of act (mode 2 with split-K): the split boundaries are then placed so that every split gets the same
  // number of ACTIVE k-blocks -- with uniform boundaries the splits over the late time slabs of a packed batch are nearly
  // empty and the early ones carry the work
  const int* act_host;
};

// bn: 0 = auto, else 64/128/192/256.  splits: split-K factor (fp32 atomics epilogue when > 1; the
// caller zeroes D or passes an existing value to accumulate onto).  max_ctas: 0 = #SMs.
int mvae_umma_gemm(const mvae_umma_operand* A, const mvae_umma_operand* B, const mvae_umma_out* D, int M, int N, int K,
                   int bn, int splits, int max_ctas, int* err_flag, cudaStream_t stream, const mvae_umma_head* head = nullptr,
                   const mvae_umma_cell* cell = nullptr, const mvae_umma_sample* sample = nullptr,
                   const mvae_umma_varlen* varlen = nullptr);

// fp32-in / fp32-out GEMM with simt::sgemm's strided interface (A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]) on the
// tensor cores: operands split into bf16 (hi, lo) parts on the fly ("bf16x3", ~2^-17 relative per product), one tcgen05 GEMM
// over K' = 3 * roundup(K, 8).  `ctx->scratch` holds the two converted operands.  Returns MVAE_ERR_UNSUPPORTED (nothing
// enqueued) when the strides are not one of the two contiguous forms, the product is tiny, or the scratch is too small --
// the caller then takes the CUDA-core path.  *launches (optional) receives the number of enqueued operations.
struct mvae_tc_ctx { void* scratch; size_t bytes; int* err_flag; };
size_t mvae_tc_sgemm_scratch_bytes(long long rows_max, long long cols_max);
int mvae_tc_sgemm(const mvae_tc_ctx* ctx, cudaStream_t st, const float* A, long long sam, long long sak, const float* B,
                  long long sbk, long long sbn, float* C, long long ldc, int M, int N, int K, const float* bias, int act,
                  int accumulate, int* launches);

// 2-CTA (tcgen05 cta_group::2) variant: K- or MN-major operands, M % 256 == N % 256 == 0, bf16 output (+ bias) or fp32 split-K red.add
// (umma_gemm2.cu).  mvae_umma_gemm routes matching calls there (MVAE_GEMM_PAIRS=0 disables it); returns
// MVAE_ERR_UNSUPPORTED, with nothing enqueued, for anything else.
int mvae_umma_gemm_pairs(const mvae_umma_operand* A, const mvae_umma_operand* B, const mvae_umma_out* D, int M, int N, int K,
                         int splits, int* err_flag, cudaStream_t stream);
