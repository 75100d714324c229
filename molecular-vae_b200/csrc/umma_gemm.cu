// tcgen05 / TMA GEMM for sm_100a:   D[M,N] (+)= A * B  (+ bias[N]),  bf16 operands, fp32 accumulate in TMEM.
//
// One kernel family serves every dense contraction on the ELBO path that is not inside the
// fused recurrence (GRU input projections over all timesteps, dX of those projections, the
// K = B*T weight-gradient reductions, the vocabulary head):
//   * A is [M,K] "K-major" (row-major, K contiguous) or [K,M] "MN-major" (M contiguous);
//     same for B with N.  MN-major operands are what the weight-gradient GEMMs need
//     (dW = dG^T * X contracts over the row index of both stored matrices) and are fed to
//     the tensor core through transposing shared-memory descriptors -- nothing is transposed
//     in HBM.
//   * warp-specialised, persistent: warp 0 = TMA producer, warp 1 = single-thread tcgen05.mma
//     issuer (+ TMEM allocator), warps 2..9 = epilogue (tcgen05.ld -> registers -> global).
//     smem ring of STAGES {A 128x64, B BNx64} SWIZZLE_128B tiles, two TMEM accumulator
//     buffers so the epilogue of unit i overlaps the main loop of unit i+1.
//   * split-K work units (fp32 red.add epilogue) give the skinny wgrad GEMMs a full grid.
#include "common.cuh"
#include "umma_gemm.h"
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;   // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB

template <int BN> struct Cfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192) ? 5 : (BN == 128) ? 6 : 8;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct KernelParams {
  int M, N, K;
  int slabA, slabB;
  int splits;          // split-K factor (>=1)
  int kb_per_split;    // BK-blocks per split
  int tiles_m, tiles_n;
  void* out;
  long long ldc;
  const float* bias;   // per-N, may be null
  int out_rb;          // bf16 output in the row-blocked layout
  int out_act;         // generic epilogue, splits == 1: 0 none, 1 SELU, 2 ReLU (after the bias)
  int out_bf16;        // 1: bf16 output, 0: fp32
  int accumulate;      // fp32 only: D += result (plain RMW, or red.add when splits > 1)
  int* err_flag;
  int head_mode;       // fused vocabulary-head epilogue (BN == 64 only), see mvae_umma_head
  mvae_umma_head head;
  int cell_mode;       // fused GRU-cell epilogue (BN == 192 / 256), see mvae_umma_cell
  mvae_umma_cell cell;
  int sample_mode;     // fused sampling epilogue (BN == 64), see mvae_umma_sample
  mvae_umma_sample sample;
  int vl_mode;         // packed-sequence skipping, see mvae_umma_varlen
  const int* vl_act;
  int vl_tiles;        // M tiles (mode 1) or k-blocks (mode 2) per slab
  int vl_nb;           // > 0: explicit split-K boundaries (k-block indices) in vl_bounds[0 .. splits]
  int vl_bounds[160];
};

// k-block range [kb0, kb1) of split `split`
__device__ __forceinline__ void split_range(const KernelParams& p, int split, int kb_total, int& kb0, int& kb1) {
  if (p.vl_nb) { kb0 = p.vl_bounds[split]; kb1 = p.vl_bounds[split + 1]; }
  else { kb0 = split * p.kb_per_split; kb1 = min(kb_total, kb0 + p.kb_per_split); }
}

// mode 1: is output tile m_blk entirely past the running sequences of its slab?
__device__ __forceinline__ bool vl_skip_tile(const KernelParams& p, int m_blk) {
  if (p.vl_mode != 1) return false;
  const int t = m_blk / p.vl_tiles, r = m_blk - t * p.vl_tiles;
  return r * BM >= __ldg(p.vl_act + t);
}
// mode 2: is k-block kb entirely past the running sequences of its slab?  (the first block of a split is always kept so
// that the accumulator is written at least once; it contributes zeros)
__device__ __forceinline__ bool vl_skip_kb(const KernelParams& p, int kb, int kb0) {
  if (p.vl_mode != 2 || kb == kb0) return false;
  const int t = kb / p.vl_tiles, r = kb - t * p.vl_tiles;
  return r * BK >= __ldg(p.vl_act + t);
}
// first k-block of the slab after the one kb lies in (once one block of a slab is inactive, the remaining ones are as well)
__device__ __forceinline__ int vl_slab_end(const KernelParams& p, int kb) { return (kb / p.vl_tiles + 1) * p.vl_tiles; }
// the first k-block of a split is never skipped; when it lies in the inactive part (whose contents are undefined) the
// producer loads it from beyond the K extent instead, where TMA fills zeros
__device__ __forceinline__ bool vl_first_kb_inactive(const KernelParams& p, int kb, int kb0) {
  if (p.vl_mode != 2 || kb != kb0) return false;
  const int t = kb / p.vl_tiles, r = kb - t * p.vl_tiles;
  return r * BK >= __ldg(p.vl_act + t);
}

// counter-based uniform in (0,1); must stay identical to u01_hash in moses.cu / oracle/moses_oracle.u01_hash
__device__ __forceinline__ float u01_hash_gemm(unsigned long long seed, unsigned int b, unsigned int i) {
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)b * 1000003ull + i + 1);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
  return (float)((x >> 40) + 0.5) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float sigmoid_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(0.5f * x));
  return fmaf(0.5f, y, 0.5f);
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) {  // 2 s: report instead of hanging the device
        if (err_flag) atomicExch(err_flag, 1);
        return false;
      }
    }
  }
  return true;
}

template <int BN, int A_MN, int B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ KernelParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tfull_bar = bars + 2 * C::STAGES;
  uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], NUM_EPI_WARPS * 32);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int kb_total = (p.K + BK - 1) / BK;
  const int n_units = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int tile = unit / p.splits, split = unit - tile * p.splits;
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        int kb0, kb1;
        split_range(p, split, kb_total, kb0, kb1);
        if (vl_skip_tile(p, m_blk)) continue;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (vl_skip_kb(p, kb, kb0)) { kb = vl_slab_end(p, kb) - 1; continue; }   // the rest of this slab is inactive too
          if (!wait_bar(&empty_bar[s], ph ^ 1, p.err_flag)) goto done;
          ptx::mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          uint8_t* a_dst = sA + s * A_STAGE_BYTES;
          uint8_t* b_dst = sB + s * C::B_STAGE_BYTES;
          const int kc = (vl_first_kb_inactive(p, kb, kb0) ? kb_total : kb) * BK;   // k coordinate (past K: zero fill)
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              ptx::tma_load_3d(a_dst + c * 8192, &tmA, &full_bar[s], m_blk * BM + c * 64, kc, p.slabA);
          } else {
            ptx::tma_load_3d(a_dst, &tmA, &full_bar[s], kc, m_blk * BM, p.slabA);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              ptx::tma_load_3d(b_dst + c * 8192, &tmB, &full_bar[s], n_blk * BN + c * 64, kc, p.slabB);
          } else {
            ptx::tma_load_3d(b_dst, &tmB, &full_bar[s], kc, n_blk * BN, p.slabB);
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const int tile = unit / p.splits, split = unit - tile * p.splits;
        int kb0, kb1;
        split_range(p, split, kb_total, kb0, kb1);
        if (kb0 >= kb1) continue;
        if (vl_skip_tile(p, tile / p.tiles_n)) continue;
        if (!wait_bar(&tempty_bar[acc], acc_ph ^ 1, p.err_flag)) goto done;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (vl_skip_kb(p, kb, kb0)) { kb = vl_slab_end(p, kb) - 1; continue; }
          if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + s * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major SW128: 8-row groups 1024 B apart, advance 32 B per K=16 inside the swizzle row.
            // MN-major SW128: 64-wide MN chunks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO),
            //                 advance two k-groups (2048 B) per K=16.
            const uint64_t adesc = A_MN ? ptx::umma_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                        : ptx::umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? ptx::umma_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                        : ptx::umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::tc_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        ptx::tc_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int chalf = (warp - 2) >> 2;   // which half of the tile's columns this warp drains
    constexpr int CH = BN / 2;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
      const int tile = unit / p.splits, split = unit - tile * p.splits;
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      int kb0, kb1;
      split_range(p, split, kb_total, kb0, kb1);
      if (kb0 >= kb1) continue;
      if (vl_skip_tile(p, m_blk)) continue;
      if (!wait_bar(&tfull_bar[acc], acc_ph, p.err_flag)) goto done;
      ptx::tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool add_bias = p.bias != nullptr && split == 0;
      if constexpr (BN == 192 || BN == 256) {
        if (p.cell_mode) {
          // ---- fused GRU cell: this warp owns 32 rows x 32 hidden units (units n_blk*64 + chalf*32 ..), 16 at a time
          constexpr int G = BN / 64;
          const mvae_umma_cell& c = p.cell;
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
          if (BN == 256 && c.lstm) {
            // ---- fused LSTM cell (torch.nn.LSTM, gate rows i,f,g,o; models.py:128,164)
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              const int ul = chalf * 32 + half * 16;
              const int u = n_blk * 64 + ul;
              uint32_t a4[4][16];
#pragma unroll
              for (int g = 0; g < 4; ++g) ptx::tmem_ld_32x16(tacc + g * 64 + ul, a4[g]);
              ptx::tmem_ld_wait();
              if (row_ok && u < c.H) {
                float pre[4][16];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  if (c.gi_f32) {
                    const float* gi = reinterpret_cast<const float*>(c.gi) + (long long)row * 4 * c.H + (long long)g * c.H + u;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                      const float4 v = __ldg(reinterpret_cast<const float4*>(gi) + k4);
                      pre[g][4 * k4] = v.x; pre[g][4 * k4 + 1] = v.y; pre[g][4 * k4 + 2] = v.z; pre[g][4 * k4 + 3] = v.w;
                    }
                  } else {
                    const __nv_bfloat16* gi = reinterpret_cast<const __nv_bfloat16*>(c.gi) + (long long)row * 4 * c.H + (long long)g * c.H + u;
                    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(gi));
                    const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(gi) + 1);
                    const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) { pre[g][2 * k] = __uint_as_float(w[k] << 16); pre[g][2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u); }
                  }
#pragma unroll
                  for (int k = 0; k < 16; ++k) pre[g][k] += __uint_as_float(a4[g][k]);
                }
                float* cp = c.cstate + (long long)row * c.H + u;
                float cprev[16], tc[16], hn[16];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const float4 v = *reinterpret_cast<const float4*>(cp + 4 * k4);
                  cprev[4 * k4] = v.x; cprev[4 * k4 + 1] = v.y; cprev[4 * k4 + 2] = v.z; cprev[4 * k4 + 3] = v.w;
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  // full-precision expf / tanhf: the encoder's gradients are within a hair of the 1e-2 bf16 budget
                  const float ig = 1.f / (1.f + expf(-pre[0][k]));
                  const float fg = 1.f / (1.f + expf(-pre[1][k]));
                  const float gg = tanhf(pre[2][k]);
                  const float og = 1.f / (1.f + expf(-pre[3][k]));
                  const float cn = fmaf(fg, cprev[k], ig * gg);
                  const float t = tanhf(cn);
                  pre[0][k] = ig; pre[1][k] = fg; pre[2][k] = gg; pre[3][k] = og;
                  tc[k] = t; hn[k] = og * t;
                  a4[0][k] = __float_as_uint(cn);
                }
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  *reinterpret_cast<float4*>(cp + 4 * k4) = make_float4(__uint_as_float(a4[0][4 * k4]), __uint_as_float(a4[0][4 * k4 + 1]),
                                                                       __uint_as_float(a4[0][4 * k4 + 2]), __uint_as_float(a4[0][4 * k4 + 3]));
                auto pack16 = [](const float* v, uint4& lo, uint4& hi) {
                  uint32_t w[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) { __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]); w[k] = *reinterpret_cast<uint32_t*>(&t); }
                  lo = make_uint4(w[0], w[1], w[2], w[3]); hi = make_uint4(w[4], w[5], w[6], w[7]);
                };
                uint4 lo, hi;
                pack16(hn, lo, hi);
                uint4* oa = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(c.out_a) + (long long)row * c.ld_a + u);
                oa[0] = lo; oa[1] = hi;
                if (c.sv) {
                  __nv_bfloat16* svp = reinterpret_cast<__nv_bfloat16*>(c.sv) + (long long)row * 6 * c.H + u;
#pragma unroll
                  for (int blk = 0; blk < 6; ++blk) {
                    pack16(blk < 4 ? pre[blk] : (blk == 4 ? cprev : tc), lo, hi);
                    uint4* o = reinterpret_cast<uint4*>(svp + (long long)blk * c.H);
                    o[0] = lo; o[1] = hi;
                  }
                }
              }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            continue;
          }
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            const int ul = chalf * 32 + half * 16;            // unit offset inside the 64-unit tile
            const int u = n_blk * 64 + ul;                    // first hidden unit of this group of 16
            uint32_t ar[16], az[16], ai[16], ah[16];
            ptx::tmem_ld_32x16(tacc + 0 * 64 + ul, ar);
            ptx::tmem_ld_32x16(tacc + 1 * 64 + ul, az);
            if (G == 4) ptx::tmem_ld_32x16(tacc + 2 * 64 + ul, ai);
            ptx::tmem_ld_32x16(tacc + (G - 1) * 64 + ul, ah);
            ptx::tmem_ld_wait();
            if (row_ok && u < c.H) {
              const float* bias = p.bias + (long long)n_blk * BN + ul;   // permuted like the weights: [gate][64]
              float gr[16], gz[16], gn[16];
              if (G == 3 && c.tbl) {
                const int tok = row < c.tok_rows ? (int)c.tok[row] : 0;
                const float* tb = c.tbl + (long long)tok * 3 * c.H + u;
                const float* ad = c.add + (long long)(row < c.tok_rows ? row : 0) * 3 * c.H + u;
#pragma unroll
                for (int gk = 0; gk < 3; ++gk) {
                  float* dst = gk == 0 ? gr : (gk == 1 ? gz : gn);
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(tb + (long long)gk * c.H) + k4);
                    const float4 b = __ldg(reinterpret_cast<const float4*>(ad + (long long)gk * c.H) + k4);
                    dst[4 * k4] = a.x + b.x; dst[4 * k4 + 1] = a.y + b.y; dst[4 * k4 + 2] = a.z + b.z; dst[4 * k4 + 3] = a.w + b.w;
                  }
                }
              } else if (G == 3) {
                const __nv_bfloat16* gi = reinterpret_cast<const __nv_bfloat16*>(c.gi) + (long long)row * 3 * c.H + u;
#pragma unroll
                for (int gk = 0; gk < 3; ++gk) {
                  const uint4 a = __ldg(reinterpret_cast<const uint4*>(gi + (long long)gk * c.H));
                  const uint4 b = __ldg(reinterpret_cast<const uint4*>(gi + (long long)gk * c.H) + 1);
                  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                  float* dst = gk == 0 ? gr : (gk == 1 ? gz : gn);
#pragma unroll
                  for (int k = 0; k < 8; ++k) { dst[2 * k] = __uint_as_float(w[k] << 16); dst[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u); }
                }
              } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) { gr[k] = 0.f; gz[k] = 0.f; gn[k] = __uint_as_float(ai[k]) + __ldg(bias + 2 * 64 + k); }
              }
              const float* hp = c.h_prev32 + (long long)row * c.H + u;
              float hn[16];
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const float4 h4 = *reinterpret_cast<const float4*>(hp + 4 * k4);
                const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                  const int k = 4 * k4 + k2;
                  const float r = sigmoid_fast(gr[k] + __uint_as_float(ar[k]) + __ldg(bias + k));
                  const float z = sigmoid_fast(gz[k] + __uint_as_float(az[k]) + __ldg(bias + 64 + k));
                  const float ghn = __uint_as_float(ah[k]) + __ldg(bias + (G - 1) * 64 + k);
                  const float n = tanh_fast(fmaf(r, ghn, gn[k]));
                  hn[k] = fmaf(z, hv[k2] - n, n);
                  gr[k] = r; gz[k] = z; gn[k] = n; ar[k] = __float_as_uint(ghn);   // reuse as the saved gates
                }
              }
              if (c.sv) {
                __nv_bfloat16* svp = reinterpret_cast<__nv_bfloat16*>(c.sv) + (long long)row * 4 * c.H + u;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                  uint32_t w[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float lo = blk == 0 ? gr[2 * k] : blk == 1 ? gz[2 * k] : blk == 2 ? gn[2 * k] : __uint_as_float(ar[2 * k]);
                    const float hi = blk == 0 ? gr[2 * k + 1] : blk == 1 ? gz[2 * k + 1] : blk == 2 ? gn[2 * k + 1] : __uint_as_float(ar[2 * k + 1]);
                    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
                    w[k] = *reinterpret_cast<uint32_t*>(&t);
                  }
                  uint4* o = reinterpret_cast<uint4*>(svp + (long long)blk * c.H);
                  o[0] = make_uint4(w[0], w[1], w[2], w[3]);
                  o[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
              }
              float* ho = c.h_next32 + (long long)row * c.H + u;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                *reinterpret_cast<float4*>(ho + 4 * k4) = make_float4(hn[4 * k4], hn[4 * k4 + 1], hn[4 * k4 + 2], hn[4 * k4 + 3]);
              uint4 p0, p1;
              {
                __nv_bfloat162 t0 = __floats2bfloat162_rn(hn[0], hn[1]), t1 = __floats2bfloat162_rn(hn[2], hn[3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(hn[4], hn[5]), t3 = __floats2bfloat162_rn(hn[6], hn[7]);
                __nv_bfloat162 t4 = __floats2bfloat162_rn(hn[8], hn[9]), t5 = __floats2bfloat162_rn(hn[10], hn[11]);
                __nv_bfloat162 t6 = __floats2bfloat162_rn(hn[12], hn[13]), t7 = __floats2bfloat162_rn(hn[14], hn[15]);
                p0 = make_uint4(*reinterpret_cast<uint32_t*>(&t0), *reinterpret_cast<uint32_t*>(&t1), *reinterpret_cast<uint32_t*>(&t2), *reinterpret_cast<uint32_t*>(&t3));
                p1 = make_uint4(*reinterpret_cast<uint32_t*>(&t4), *reinterpret_cast<uint32_t*>(&t5), *reinterpret_cast<uint32_t*>(&t6), *reinterpret_cast<uint32_t*>(&t7));
              }
              uint4* oa = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(c.out_a) + (long long)row * c.ld_a + u);
              oa[0] = p0; oa[1] = p1;
              if (c.out_b) {
                uint4* ob = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(c.out_b) + (long long)row * c.ld_b + u);
                ob[0] = p0; ob[1] = p1;
              }
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tempty_bar[acc]);
          if (++acc == 2) { acc = 0; acc_ph ^= 1; }
          continue;
        }
      }
      if constexpr (BN == 64) {
        if (p.sample_mode) {
          // ---- fused sampling step: one thread owns one sequence's logits (the quarter's second warp only signals)
          if (chalf == 0) {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, r0);
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + 32, r1);
            ptx::tmem_ld_wait();
            const mvae_umma_sample& sp = p.sample;
            if (row < sp.B) {
              float v[64];
#pragma unroll
              for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
              float m = -INFINITY;
              int arg = 0;
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                if (j >= sp.V) break;
                v[j] = (v[j] + (p.bias ? __ldg(p.bias + j) : 0.f)) * sp.inv_temp;
                if (v[j] > m) { m = v[j]; arg = j; }
              }
              int tok = arg;
              if (sp.mode == 1) {
                float tot = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                  if (j >= sp.V) break;
                  v[j] = expf(v[j] - m);
                  tot += v[j];
                }
                const float uu = u01_hash_gemm(sp.seed_dev ? *sp.seed_dev : sp.seed, (unsigned)row, (unsigned)sp.step) * tot;
                float cum = 0.f;
                bool found = false;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                  if (j >= sp.V) break;
                  cum += v[j];
                  if (!found && cum > uu) { tok = j; found = true; }
                }
              }
              sp.w_cur[row] = (unsigned char)tok;
              if (sp.onehot_next) {
                uint4* oh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(sp.onehot_next) + (long long)row * sp.oh_ld);
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8) {
                  uint4 o = make_uint4(0u, 0u, 0u, 0u);
                  if ((tok >> 3) == j8) {
                    const uint32_t one = (tok & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in the high / low half
                    const int w2 = (tok & 7) >> 1;
                    o.x = w2 == 0 ? one : 0u; o.y = w2 == 1 ? one : 0u; o.z = w2 == 2 ? one : 0u; o.w = w2 == 3 ? one : 0u;
                  }
                  oh[j8] = o;
                }
              }
              if (!sp.done[row]) {
                sp.x[(long long)row * sp.max_len + sp.step] = (unsigned char)tok;
                if (tok == sp.eos) { sp.end[row] = sp.step + 1; sp.done[row] = 1; }
              }
            }
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tempty_bar[acc]);
          if (++acc == 2) { acc = 0; acc_ph ^= 1; }
          continue;
        }
        if (p.head_mode) {
          // ---- fused vocabulary head: one thread owns one row's 64 logits (the quarter's second warp only signals)
          if (chalf == 0) {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, r0);
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + 32, r1);
            ptx::tmem_ld_wait();
            const mvae_umma_head& h = p.head;
            double loss = 0.0;
            if (row_ok) {
              const int t = row / h.Bp, b = row - t * h.Bp;
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(h.dlogits) + (long long)row * 64);
              if (h.mode == 1) {
                // ---- shifted cross entropy (MOSES VAE): log-softmax + NLL + its gradient, full-precision exp / log
                const int L = b < h.B ? __ldg(h.lens + b) : 0;
                if (!(b < h.B && t + 1 < L)) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) o[j] = make_uint4(0u, 0u, 0u, 0u);
                } else {
                  const int y = h.ids[(long long)b * h.T + t + 1];
                  float v[64];
#pragma unroll
                  for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
                  float m = -INFINITY, at = 0.f;
#pragma unroll
                  for (int j = 0; j < 64; ++j) {
                    if (j >= h.C) break;
                    if (p.bias) v[j] += __ldg(p.bias + j);
                    m = fmaxf(m, v[j]);
                    if (j == y) at = v[j];
                  }
                  float ssum = 0.f;
#pragma unroll
                  for (int j = 0; j < 64; ++j) {
                    if (j >= h.C) break;
                    v[j] = expf(v[j] - m);
                    ssum += v[j];
                  }
                  const float sc = h.rec_w / (float)__ldg(h.Mcount);
                  const float inv = sc / ssum;
#pragma unroll
                  for (int j8 = 0; j8 < 8; ++j8) {
                    float d[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                      const int j = j8 * 8 + k;
                      d[k] = j < h.C ? fmaf(v[j], inv, j == y ? -sc : 0.f) : 0.f;
                    }
                    uint4 pk;
                    __nv_bfloat162 b0 = __floats2bfloat162_rn(d[0], d[1]);
                    __nv_bfloat162 b1 = __floats2bfloat162_rn(d[2], d[3]);
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(d[4], d[5]);
                    __nv_bfloat162 b3 = __floats2bfloat162_rn(d[6], d[7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&b0);
                    pk.y = *reinterpret_cast<uint32_t*>(&b1);
                    pk.z = *reinterpret_cast<uint32_t*>(&b2);
                    pk.w = *reinterpret_cast<uint32_t*>(&b3);
                    o[j8] = pk;
                  }
                  loss = (double)(m + logf(ssum) - at);
                }
              } else if (b >= h.B) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = make_uint4(0u, 0u, 0u, 0u);
              } else {
                const int y = h.ids[(long long)b * h.T + t];
                // v: logits -> probabilities, g: d(loss)/d(p); every loop stops at the (warp-uniform) charset size so the
                // 4 epilogue warps of a CTA do not pay for the 29 pad columns
                float v[64], g[64];
#pragma unroll
                for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
                float m = -INFINITY;
                int arg = 0;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                  if (j >= h.C) break;
                  if (p.bias) v[j] += __ldg(p.bias + j);
                  if (v[j] > m) { m = v[j]; arg = j; }
                }
                float ssum = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                  if (j >= h.C) break;
                  v[j] = __expf(v[j] - m);
                  ssum += v[j];
                }
                const float inv = 1.0f / ssum;
                float l = 0.f, gp = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                  if (j >= h.C) break;
                  const float pj = v[j] * inv;
                  const float x = (j == y) ? 1.f : 0.f;
                  const float q1 = 1.f - pj;
                  l -= fmaxf(__logf(x > 0.f ? pj : q1), -100.f);      // BCELoss clamps each log at -100
                  const float gj = h.gscale * (pj - x) / fmaxf(pj * q1, 1e-12f);
                  gp = fmaf(gj, pj, gp);
                  v[j] = pj; g[j] = gj;
                }
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8) {
                  float d[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const int j = j8 * 8 + k;
                    d[k] = j < h.C ? v[j] * (g[j] - gp) : 0.f;
                  }
                  uint4 pk;
                  __nv_bfloat162 b0 = __floats2bfloat162_rn(d[0], d[1]);
                  __nv_bfloat162 b1 = __floats2bfloat162_rn(d[2], d[3]);
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(d[4], d[5]);
                  __nv_bfloat162 b3 = __floats2bfloat162_rn(d[6], d[7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&b0);
                  pk.y = *reinterpret_cast<uint32_t*>(&b1);
                  pk.z = *reinterpret_cast<uint32_t*>(&b2);
                  pk.w = *reinterpret_cast<uint32_t*>(&b3);
                  o[j8] = pk;
                }
                loss = (double)l;
                if (h.hit_count && arg == y) atomicAdd(h.hit_count + b, 1);
              }
            }
            for (int off = 16; off; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
            if (lane == 0 && h.bce_sum && loss != 0.0) atomicAdd(h.bce_sum, loss);
          }
          ptx::tc_fence_before();
          ptx::mbar_arrive(&tempty_bar[acc]);
          if (++acc == 2) { acc = 0; acc_ph ^= 1; }
          continue;
        }
      }
      // Fast path (full bf16 output tile, aligned): the TMEM load of chunk c+1 is in flight while chunk c is biased, packed and
      // stored; the bias comes in 16-byte pieces.  K-short GEMMs (the input projections: 8 k-blocks per tile) are otherwise
      // bound by this epilogue, not by the tensor pipe.
      if (p.out_bf16 && p.splits == 1 && (m_blk + 1) * BM <= p.M && (long long)(n_blk + 1) * BN <= p.N && (p.ldc & 7) == 0 &&
          (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
        constexpr int NCH = CH / 32;
        const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + chalf * CH;
        uint32_t r0[32], r1[32];
        auto process = [&](const uint32_t (&rr)[32], int c) {
          const int col0 = n_blk * BN + chalf * CH + c * 32;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
            if (add_bias) {
              ba = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + h * 8));
              bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + h * 8) + 1);
            }
            uint4 pk;
            __nv_bfloat162 b0 = __floats2bfloat162_rn(__uint_as_float(rr[h * 8 + 0]) + ba.x, __uint_as_float(rr[h * 8 + 1]) + ba.y);
            __nv_bfloat162 b1 = __floats2bfloat162_rn(__uint_as_float(rr[h * 8 + 2]) + ba.z, __uint_as_float(rr[h * 8 + 3]) + ba.w);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(rr[h * 8 + 4]) + bb.x, __uint_as_float(rr[h * 8 + 5]) + bb.y);
            __nv_bfloat162 b3 = __floats2bfloat162_rn(__uint_as_float(rr[h * 8 + 6]) + bb.z, __uint_as_float(rr[h * 8 + 7]) + bb.w);
            pk.x = *reinterpret_cast<uint32_t*>(&b0);
            pk.y = *reinterpret_cast<uint32_t*>(&b1);
            pk.z = *reinterpret_cast<uint32_t*>(&b2);
            pk.w = *reinterpret_cast<uint32_t*>(&b3);
            const int col = col0 + h * 8;
            __nv_bfloat16* o = p.out_rb ? reinterpret_cast<__nv_bfloat16*>(p.out) +
                                              ((long long)(row >> 5) * (p.ldc >> 3) + (col >> 3)) * 256 + (row & 31) * 8
                                        : reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldc + col;
            *reinterpret_cast<uint4*>(o) = pk;
          }
        };
        ptx::tmem_ld_32x32(tb, r0);
#pragma unroll 1
        for (int c = 0; c < NCH; c += 2) {
          ptx::tmem_ld_wait();
          if (c + 1 < NCH) ptx::tmem_ld_32x32(tb + (c + 1) * 32, r1);
          process(r0, c);
          if (c + 1 < NCH) {
            ptx::tmem_ld_wait();
            if (c + 2 < NCH) ptx::tmem_ld_32x32(tb + (c + 2) * 32, r0);
            process(r1, c + 1);
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        continue;
      }
#pragma unroll 1
      for (int c0 = chalf * CH; c0 < (chalf + 1) * CH; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c0, r);
        ptx::tmem_ld_wait();
        const int col0 = n_blk * BN + c0;
        if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (add_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
          if (p.out_act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = 1.0507009873554804934193349852946f * (v[j] > 0.f ? v[j] : 1.6732632423543772848170429916717f * expm1f(v[j]));
          } else if (p.out_act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          const bool full = (col0 + 32 <= p.N);
          if (p.out_bf16 && p.out_rb) {
            // row-blocked output [row/32][ldc/8][row%32][8]: a warp's 32 rows x 8 cols are 512 B contiguous (full sectors)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int col = col0 + h * 8;
              if (col < p.N) {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                   ((long long)(row >> 5) * (p.ldc >> 3) + (col >> 3)) * 256 + (row & 31) * 8;
                uint4 pk;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(v[h * 8 + 0], v[h * 8 + 1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(v[h * 8 + 2], v[h * 8 + 3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v[h * 8 + 4], v[h * 8 + 5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(v[h * 8 + 6], v[h * 8 + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&b0);
                pk.y = *reinterpret_cast<uint32_t*>(&b1);
                pk.z = *reinterpret_cast<uint32_t*>(&b2);
                pk.w = *reinterpret_cast<uint32_t*>(&b3);
                *reinterpret_cast<uint4*>(o) = pk;
              }
            }
          } else if (p.out_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldc + col0;
            if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(v[j + 0], v[j + 1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&b0);
                pk.y = *reinterpret_cast<uint32_t*>(&b1);
                pk.z = *reinterpret_cast<uint32_t*>(&b2);
                pk.w = *reinterpret_cast<uint32_t*>(&b3);
                *reinterpret_cast<uint4*>(o + j) = pk;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16_rn(v[j]);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + col0;
            if (p.splits > 1) {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(o + j, v[j]);
            } else if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 w = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (p.accumulate) {
                  float4 old = *reinterpret_cast<float4*>(o + j);
                  w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
                }
                *reinterpret_cast<float4*>(o + j) = w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = p.accumulate ? o[j] + v[j] : v[j];
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Build the 3-D tensor map of one bf16 operand.  K-major: dims {K, rows, slabs}, box {64, box_rows, 1}.
// MN-major: dims {rows(MN), K, slabs}, box {64, 64, 1}.
int make_map(CUtensorMap* map, const mvae_umma_operand& op, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MVAE_ERR_DRIVER;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const long long slabs = op.slabs > 0 ? op.slabs : 1;
  // stored matrix: rows x cols with `cols` contiguous in the plain layout
  const long long rows = op.mn_major ? op.k : op.mn;
  const long long cols = op.mn_major ? op.mn : op.k;
  const int brow = op.mn_major ? 64 : box_rows;   // box extent along the stored rows
  if (reinterpret_cast<uintptr_t>(op.ptr) & 15) return MVAE_ERR_INVALID;
  CUresult r;
  if (op.rb) {
    // a 5-D tensor map over the row-blocked layout needs non-monotonic strides, which the TMA unit does not honour
    return MVAE_ERR_UNSUPPORTED;
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
    cuuint64_t strides[2] = {(cuuint64_t)op.ld * 2, (cuuint64_t)(slabs > 1 ? op.slab_stride : op.ld * rows) * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)brow, 1};
    if ((strides[0] & 15) || (strides[1] & 15)) return MVAE_ERR_INVALID;
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(op.ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_DRIVER;
}

template <int BN, int A_MN, int B_MN>
int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const KernelParams& kp, int grid, cudaStream_t st) {
  auto kern = umma_gemm_kernel<BN, A_MN, B_MN>;
  static size_t attr_cache[64] = {0};
  MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(kern), Cfg<BN>::SMEM_BYTES, attr_cache));
  kern<<<grid, NUM_THREADS, Cfg<BN>::SMEM_BYTES, st>>>(tmA, tmB, kp);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}

template <int BN>
int launch_bn(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const KernelParams& kp, int grid,
              cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_t<BN, 0, 0>(tmA, tmB, kp, grid, st);
  if (!a_mn && b_mn) return launch_t<BN, 0, 1>(tmA, tmB, kp, grid, st);
  if (a_mn && !b_mn) return launch_t<BN, 1, 0>(tmA, tmB, kp, grid, st);
  return launch_t<BN, 1, 1>(tmA, tmB, kp, grid, st);
}

int g_num_sms = 0;

}  // namespace

int mvae_umma_gemm(const mvae_umma_operand* A, const mvae_umma_operand* B, const mvae_umma_out* D, int M, int N, int K,
                   int bn, int splits, int max_ctas, int* err_flag, cudaStream_t stream, const mvae_umma_head* head,
                   const mvae_umma_cell* cell, const mvae_umma_sample* sample, const mvae_umma_varlen* varlen) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K <= 0) return MVAE_ERR_INVALID;
  if (varlen && (!varlen->act || (varlen->mode != 1 && varlen->mode != 2) || varlen->rows_per_slab <= 0 ||
                 (varlen->rows_per_slab % BM) || (varlen->mode == 1 ? M : K) % varlen->rows_per_slab))
    return MVAE_ERR_INVALID;
  if (sample && (bn != 64 || splits > 1 || N > 64 || sample->V > N || !sample->w_cur || !sample->x || !sample->end || !sample->done))
    return MVAE_ERR_INVALID;
  if (cell && ((cell->gates != 3 && cell->gates != 4) || bn != cell->gates * 64 || splits > 1 || (cell->H & 63) ||
               N != cell->gates * cell->H || !cell->out_a || (cell->ld_a & 7) || (cell->out_b && (cell->ld_b & 7))))
    return MVAE_ERR_INVALID;
  if (cell && cell->lstm && (cell->gates != 4 || !cell->gi || !cell->cstate)) return MVAE_ERR_INVALID;
  if (cell && !cell->lstm && (!D->bias || !cell->h_prev32 || !cell->h_next32 ||
                              (cell->gates == 3 && !cell->gi && !(cell->tbl && cell->add && cell->tok))))
    return MVAE_ERR_INVALID;
  if (head && (bn != 64 || splits > 1 || N > 64 || head->C > N || !head->ids || !head->dlogits)) return MVAE_ERR_INVALID;
  if (head && head->mode == 1 && (!head->lens || !head->Mcount)) return MVAE_ERR_INVALID;
  if (g_num_sms == 0) {
    int dev = 0;
    MVAE_CUDA_CHECK(cudaGetDevice(&dev));
    MVAE_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!head && !cell && !sample && !varlen && (bn == 0 || bn == 256) && max_ctas == 0) {
    // CTA pairs (256 x 256 tiles, half a B tile per CTA) when the shape allows: projections with a bf16 result, dX, and the
    // split-K weight gradients (fp32 red.add)
    const char* e = getenv("MVAE_GEMM_PAIRS");
    if (!e || atoi(e) != 0) {
      const int rc = mvae_umma_gemm_pairs(A, B, D, M, N, K, splits, err_flag, stream);
      if (rc != MVAE_ERR_UNSUPPORTED) return rc;
    }
  }
  if (bn == 0) bn = (N > 128) ? 256 : (N > 64 ? 128 : 64);
  if (bn != 64 && bn != 128 && bn != 192 && bn != 256) return MVAE_ERR_INVALID;
  if (splits < 1) splits = 1;
  if (splits > 1 && (D->bf16 || D->bias)) return MVAE_ERR_INVALID;  // split-K reduces with fp32 red.add
  if (D->bf16 && D->accumulate) return MVAE_ERR_INVALID;
  if (D->act && (splits > 1 || D->bf16 || D->accumulate || head || cell || sample)) return MVAE_ERR_INVALID;
  KernelParams kp{};
  kp.M = M; kp.N = N; kp.K = K;
  kp.slabA = A->slab; kp.slabB = B->slab;
  const int kb_total = ceil_div(K, BK);
  if (splits > kb_total) splits = kb_total;
  kp.kb_per_split = ceil_div(kb_total, splits);
  kp.splits = ceil_div(kb_total, kp.kb_per_split);  // no empty splits
  kp.tiles_m = ceil_div(M, BM);
  kp.tiles_n = ceil_div(N, bn);
  kp.out = D->ptr; kp.ldc = D->ld; kp.bias = D->bias; kp.out_bf16 = D->bf16; kp.accumulate = D->accumulate;
  kp.err_flag = err_flag;
  kp.out_rb = D->rb;
  kp.out_act = D->act;
  kp.head_mode = head ? 1 : 0;
  if (head) kp.head = *head;
  kp.cell_mode = cell ? 1 : 0;
  if (cell) kp.cell = *cell;
  kp.sample_mode = sample ? 1 : 0;
  if (sample) kp.sample = *sample;
  kp.vl_mode = varlen ? varlen->mode : 0;
  kp.vl_act = varlen ? varlen->act : nullptr;
  kp.vl_tiles = varlen ? varlen->rows_per_slab / (varlen->mode == 1 ? BM : BK) : 1;
  kp.vl_nb = 0;
  if (varlen && varlen->mode == 2 && varlen->act_host && kp.splits > 1 && kp.splits < 160) {
    // balanced split-K: equal numbers of active k-blocks per split (slab t has ceil(act[t] / 64) active blocks, the leading ones)
    const int slabs = K / varlen->rows_per_slab;
    long long total_active = 0;
    for (int t = 0; t < slabs; ++t) {
      const int a = ceil_div(varlen->act_host[t], BK);
      total_active += a < kp.vl_tiles ? a : kp.vl_tiles;
    }
    if (total_active >= kp.splits) {
      int sidx = 1;
      long long cum = 0;
      kp.vl_bounds[0] = 0;
      for (int t = 0; t < slabs && sidx < kp.splits; ++t) {
        int a = ceil_div(varlen->act_host[t], BK);
        if (a > kp.vl_tiles) a = kp.vl_tiles;
        for (int r = 0; r < a && sidx < kp.splits; ++r) {
          ++cum;
          if (cum * kp.splits >= total_active * sidx) kp.vl_bounds[sidx++] = t * kp.vl_tiles + r + 1;
        }
      }
      while (sidx <= kp.splits) kp.vl_bounds[sidx++] = kb_total;
      kp.vl_nb = kp.splits;
    }
  }
  if (D->rb && (!D->bf16 || (D->ld & 7) || (N & 7))) return MVAE_ERR_INVALID;
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, *A, BM);
  if (rc) return rc;
  rc = make_map(&tmB, *B, bn);
  if (rc) return rc;
  const long long units = (long long)kp.tiles_m * kp.tiles_n * kp.splits;
  int cap = max_ctas > 0 ? max_ctas : g_num_sms;
  int grid = (int)(units < cap ? units : cap);
  const int a_mn = A->mn_major, b_mn = B->mn_major;
  switch (bn) {
    case 64: return launch_bn<64>(a_mn, b_mn, tmA, tmB, kp, grid, stream);
    case 128: return launch_bn<128>(a_mn, b_mn, tmA, tmB, kp, grid, stream);
    case 192: return launch_bn<192>(a_mn, b_mn, tmA, tmB, kp, grid, stream);
    default: return launch_bn<256>(a_mn, b_mn, tmA, tmB, kp, grid, stream);
  }
}

// ---------------------------------------------------------------------------------------
// fp32 GEMM on the tensor cores at fp32-class accuracy ("bf16x3"): every fp32 operand element x is split into
// hi = bf16(x), lo = bf16(x - hi); A*B ~ A_hi*B_hi + A_hi*B_lo + A_lo*B_hi is ONE bf16 GEMM over a three times longer
// contraction, K' = [hi | hi | lo] (A) against [hi | lo | hi] (B), fp32 accumulation in TMEM.  The dropped terms are
// ~2^-17 relative per product.  Used for the small latent / encoder Linears (K, N of a few hundred; M = the batch) that
// were CUDA-core SGEMMs: two conversion launches + one tcgen05 GEMM instead of a latency-bound SIMT kernel.
// ---------------------------------------------------------------------------------------
namespace {

// One thread converts 8 consecutive elements along the source's contiguous dimension.
//   layout 0 (K-major operand):  X(mn, k) = src[mn * so + k];  dst[mn][seg * Kp + k]            (ldd = 3 * Kp)
//   layout 1 (MN-major operand): X(mn, k) = src[k * so + mn];  dst[seg * Kp + k][mn]            (ldd = roundup(MN, 8))
// role 0: segments (hi, hi, lo); role 1: (hi, lo, hi).  Pads (k in [K, Kp), mn in [MN, ldd)) are zero-filled.
__global__ void split3_kernel(const float* __restrict__ src, long long so, int n_outer, int n_outer_p, int n_inner, int n_inner_p,
                              __nv_bfloat16* __restrict__ dst, int layout, int Kp, long long ldd, int role) {
  const int groups = n_inner_p >> 3;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)n_outer_p * groups) return;
  const int o = (int)(idx / groups), c0 = (int)(idx % groups) * 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = c0 + 2 * j + e;
      x[e] = (o < n_outer && c < n_inner) ? __ldg(src + (long long)o * so + c) : 0.f;
    }
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[0]), h1 = __float2bfloat16_rn(x[1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x[0] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x[1] - __bfloat162float(h1));
    hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  const uint4 H = make_uint4(hi[0], hi[1], hi[2], hi[3]), L = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
  for (int seg = 0; seg < 3; ++seg) {
    const bool low = role == 0 ? seg == 2 : seg == 1;
    __nv_bfloat16* d = layout == 0 ? dst + (long long)o * ldd + (long long)seg * Kp + c0
                                   : dst + ((long long)seg * Kp + o) * ldd + c0;
    *reinterpret_cast<uint4*>(d) = low ? L : H;
  }
}

}  // namespace

size_t mvae_tc_sgemm_scratch_bytes(long long rows_max, long long cols_max) {
  const size_t r = (size_t)((rows_max + 7) & ~7ll), c = (size_t)((cols_max + 7) & ~7ll);
  return 2 * (3 * r * c * 2 + 256);
}

int mvae_tc_sgemm(const mvae_tc_ctx* ctx, cudaStream_t st, const float* A, long long sam, long long sak, const float* B,
                  long long sbk, long long sbn, float* C, long long ldc, int M, int N, int K, const float* bias, int act,
                  int accumulate, int* launches) {
  if (!ctx || !ctx->scratch || M <= 0 || N <= 0 || K <= 0) return MVAE_ERR_UNSUPPORTED;
  if ((long long)M * N * K < (1ll << 22)) return MVAE_ERR_UNSUPPORTED;     // tiny products: one SIMT launch is cheaper
  const bool a_k = sak == 1, b_k = sbk == 1;
  if ((!a_k && sam != 1) || (!b_k && sbn != 1)) return MVAE_ERR_UNSUPPORTED;
  if (act && accumulate) return MVAE_ERR_UNSUPPORTED;
  const int Kp = (K + 7) & ~7, Mp = (M + 7) & ~7, Np = (N + 7) & ~7;
  const size_t a_bytes = (a_k ? (size_t)M * 3 * Kp : (size_t)3 * Kp * Mp) * 2;
  const size_t b_bytes = (b_k ? (size_t)N * 3 * Kp : (size_t)3 * Kp * Np) * 2;
  const size_t a_off = 0, b_off = (a_bytes + 255) & ~(size_t)255;
  if (b_off + b_bytes > ctx->bytes) return MVAE_ERR_UNSUPPORTED;
  __nv_bfloat16* dA = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(ctx->scratch) + a_off);
  __nv_bfloat16* dB = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(ctx->scratch) + b_off);
  auto conv = [&](const float* src, bool kfast, long long so, int MN, int MNp, __nv_bfloat16* dst, int role) {
    // kfast: outer = mn (no pad rows), inner = k;  else: outer = k (padded to Kp), inner = mn
    const int n_outer = kfast ? MN : K, n_outer_p = kfast ? MN : Kp, n_inner = kfast ? K : MN, n_inner_p = kfast ? Kp : MNp;
    const long long total = (long long)n_outer_p * (n_inner_p >> 3);
    split3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, so, n_outer, n_outer_p, n_inner, n_inner_p, dst,
                                                                   kfast ? 0 : 1, Kp, kfast ? 3ll * Kp : (long long)MNp, role);
  };
  conv(A, a_k, a_k ? sam : sak, M, Mp, dA, 0);
  conv(B, b_k, b_k ? sbn : sbk, N, Np, dB, 1);
  MVAE_CUDA_CHECK(cudaGetLastError());
  int n_launch = 3;
  const int bn = N <= 64 ? 64 : (N <= 1024 ? 128 : 256);
  const int units = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  const int kb_total = (3 * Kp + BK - 1) / BK;
  int splits = 1;
  if (!bias && !act && units <= 48 && kb_total >= 16) {
    splits = (148 + units - 1) / units;
    if (splits > kb_total / 4) splits = kb_total / 4;
    if (splits < 1) splits = 1;
  }
  if (splits > 1 && !accumulate) {
    if (ldc == N) MVAE_CUDA_CHECK(cudaMemsetAsync(C, 0, (size_t)M * N * 4, st));
    else MVAE_CUDA_CHECK(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, st));
    ++n_launch;
  }
  if (launches) *launches = n_launch;
  mvae_umma_operand a{dA, a_k ? 0 : 1, M, 3ll * Kp, a_k ? 3ll * Kp : (long long)Mp, 1, 0, 0, 0};
  mvae_umma_operand b{dB, b_k ? 0 : 1, N, 3ll * Kp, b_k ? 3ll * Kp : (long long)Np, 1, 0, 0, 0};
  mvae_umma_out o{C, ldc, 0, (accumulate || splits > 1) ? 1 : 0, bias, 0, act};
  return mvae_umma_gemm(&a, &b, &o, M, N, 3 * Kp, bn, splits, 0, ctx->err_flag, st);
}
