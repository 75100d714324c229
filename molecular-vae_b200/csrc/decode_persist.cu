// Persistent decode kernel: VAE.sample's whole decode loop (mosesvae.py:239-251 -- max_len - 1 steps x (L GRU layers +
// vocabulary head + token choice)) as ONE launch.
//
// The per-step path (moses.cu: sample_fused) launches L cell-fused GEMMs + 1 sampling GEMM per step.  Each of those GEMMs
// has 512 work units at batch 8192 (3.46 waves of the 148 SMs: the last wave is half empty) and drains completely before
// the next one starts.  A decode step has no dependency ACROSS row tiles: unit (step i, layer l, row tile m, unit slice n)
// only needs layer l-1 of step i and layer l of step i-1 FOR THE SAME 128 ROWS.  So here every CTA walks one global list
// of work units -- step-major, inside a step layer-major, the head units of a row tile trailing that tile's top-layer
// units by 20 tiles -- round-robin (unit g belongs to CTA g mod grid), and a unit starts as soon as the per-(layer, row
// tile) completion counters say its rows are ready.  Units of the next layer / step fill what would have been the empty
// tail of a wave, nothing drains, and there is one launch instead of 396.
//
// Same arithmetic as the per-step path (same tcgen05 GEMM pipeline: TMA producer warp, single-thread MMA issuer, 8
// epilogue warps, 4 x 48 KB operand stages, two TMEM accumulators; same fused GRU-cell / sampling epilogues), so the decoded
// tokens are identical (tests/test_gpu_moses.py).  State stays in global memory / L2 (3 layers x 512 units x 128 rows of
// fp32 master state alone are 768 KB per row tile: it cannot live on an SM), weights are re-read from L2.
#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "decode_persist.h"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16, BN_CELL = 256, BN_HEAD = 64;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int B_STAGE_BYTES = BN_CELL * BK * 2;   // 32 KB (the head uses the first 8 KB)
constexpr int STAGES = 4;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + 256;
constexpr int HEAD_LAG = 20;   // row tiles the head units trail the top layer's units by (> 148 / 8 units in flight)

struct DecodeMaps {
  CUtensorMap A[4][2];   // layer operands by parity
  CUtensorMap AH[2];     // top layer's h half by parity (head operand)
  CUtensorMap B[4];      // permuted cell weights
  CUtensorMap BH;        // vocabulary weights
};

struct DecodeParams {
  int B, Bp, Hd, L, V, K0, max_len, eos, mode, tiles_m, units_per_step;
  float inv_temp;
  unsigned long long seed; const unsigned long long* seed_dev;
  void* xh[4][2]; float* hm[4][2]; const float* bcat[4]; const float* bfc;
  unsigned char* w_cur; unsigned char* x; int* end; unsigned char* done;
  unsigned int* counters; const unsigned int* sched; int* err_flag;
};

// counter-based uniform in (0,1); must stay identical to u01_hash_gemm in umma_gemm.cu / oracle/moses_oracle.u01_hash
__device__ __forceinline__ float u01_hash_dec(unsigned long long seed, unsigned int b, unsigned int i) {
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
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0) {
      const unsigned long long now = gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicExch(err_flag, 5); return false; }   // 2 s: report instead of hanging
      if (*(volatile int*)err_flag) return false;
    }
  }
  return true;
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool wait_count(const unsigned int* ctr, unsigned int target, int* err_flag) {
  if (ld_acquire(ctr) >= target) return true;
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (ld_acquire(ctr) < target) {
    __nanosleep(64);
    if ((++spins & 0xFF) == 0) {
      const unsigned long long now = gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicExch(err_flag, 6); return false; }
      if (*(volatile int*)err_flag) return false;
    }
  }
  return true;
}

// accumulator hand-back: the tcgen05 fences order the TMEM reads, no generic-proxy data rides on these arrives
__device__ __forceinline__ void arrive_relaxed_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void remote_arrive_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// schedule entry: layer (8 bits, L = head) | unit slice n (8 bits) | row tile m (16 bits)
__global__ void build_sched_kernel(unsigned int* sched, int tiles_m, int L) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int u = 0;
  for (int l = 0; l + 1 < L; ++l)
    for (int m = 0; m < tiles_m; ++m)
      for (int n = 0; n < 8; ++n) sched[u++] = ((unsigned)l << 24) | ((unsigned)n << 16) | (unsigned)m;
  const int lag = tiles_m > HEAD_LAG ? HEAD_LAG : tiles_m;
  for (int m = 0; m < tiles_m; ++m) {
    for (int n = 0; n < 8; ++n) sched[u++] = ((unsigned)(L - 1) << 24) | ((unsigned)n << 16) | (unsigned)m;
    if (m >= lag) sched[u++] = ((unsigned)L << 24) | (unsigned)(m - lag);
  }
  for (int m = tiles_m - lag; m < tiles_m; ++m) sched[u++] = ((unsigned)L << 24) | (unsigned)m;
}

// the fused GRU cell of one warp (32 rows x 32 units of a 64-unit tile) and the token choice: shared by both kernel versions
// `release_acc()` hands the accumulator back to the MMA issuer; it is called as soon as this warp's LAST TMEM load of the unit
// has completed -- before the second half's gate math and global stores -- so that the next-but-one unit's MMAs do not wait
// for this unit's stores (ncu: the tensor pipe was 48 % active with the epilogue warps 24 % of their time in release fences).
template <typename F>
__device__ __forceinline__ void cell_epilogue(const DecodeParams& p, uint32_t tacc, int layer, int n, int row, int chalf, int cur, int nxt,
                                              F release_acc) {
  const int H = p.Hd, L = p.L;
  const int ldx = layer == 0 ? p.K0 : 2 * H, hoff = layer == 0 ? p.K0 - H : H;
  const float* hprev = p.hm[layer][cur];
  float* hnext = p.hm[layer][nxt];
  __nv_bfloat16* out_a = reinterpret_cast<__nv_bfloat16*>(p.xh[layer][nxt]) + hoff;
  __nv_bfloat16* out_b = layer + 1 < L ? reinterpret_cast<__nv_bfloat16*>(p.xh[layer + 1][cur]) : nullptr;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int ul = chalf * 32 + half * 16;
    const int u = n * 64 + ul;
    uint32_t ar[16], az[16], ai[16], ah[16];
    ptx::tmem_ld_32x16(tacc + 0 * 64 + ul, ar);
    ptx::tmem_ld_32x16(tacc + 1 * 64 + ul, az);
    ptx::tmem_ld_32x16(tacc + 2 * 64 + ul, ai);
    ptx::tmem_ld_32x16(tacc + 3 * 64 + ul, ah);
    ptx::tmem_ld_wait();
    if (half == 1) release_acc();
    const float* bias = p.bcat[layer] + (long long)n * BN_CELL + ul;   // permuted like the weights: [gate][64]
    float bg[4][16];                                                     // the 16 biases of each gate block as 16-byte loads
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + g * 64) + k4);
        bg[g][4 * k4] = b4.x; bg[g][4 * k4 + 1] = b4.y; bg[g][4 * k4 + 2] = b4.z; bg[g][4 * k4 + 3] = b4.w;
      }
    float gn[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) gn[k] = __uint_as_float(ai[k]) + bg[2][k];
    const float* hp = hprev + (long long)row * H + u;
    float hn[16];
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const float4 h4 = *reinterpret_cast<const float4*>(hp + 4 * k4);
      const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const int k = 4 * k4 + k2;
        const float r = sigmoid_fast(0.f + __uint_as_float(ar[k]) + bg[0][k]);
        const float z = sigmoid_fast(0.f + __uint_as_float(az[k]) + bg[1][k]);
        const float ghn = __uint_as_float(ah[k]) + bg[3][k];
        const float nn = tanh_fast(fmaf(r, ghn, gn[k]));
        hn[k] = fmaf(z, hv[k2] - nn, nn);
      }
    }
    float* ho = hnext + (long long)row * H + u;
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
    uint4* oa = reinterpret_cast<uint4*>(out_a + (long long)row * ldx + u);
    oa[0] = p0; oa[1] = p1;
    if (out_b) {
      uint4* ob = reinterpret_cast<uint4*>(out_b + (long long)row * 2 * H + u);
      ob[0] = p0; ob[1] = p1;
    }
  }
}
template <typename F>
__device__ __forceinline__ void head_epilogue(const DecodeParams& p, uint32_t tacc, int row, int i, int nxt, F release_acc) {
  uint32_t r0[32], r1[32];
  ptx::tmem_ld_32x32(tacc, r0);
  ptx::tmem_ld_32x32(tacc + 32, r1);
  ptx::tmem_ld_wait();
  release_acc();
  if (row >= p.B) return;
  float v[64];
#pragma unroll
  for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
  float mx = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    if (j >= p.V) break;
    v[j] = (v[j] + (p.bfc ? __ldg(p.bfc + j) : 0.f)) * p.inv_temp;
    if (v[j] > mx) { mx = v[j]; arg = j; }
  }
  int tok = arg;
  if (p.mode == 1) {
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      if (j >= p.V) break;
      v[j] = expf(v[j] - mx);
      tot += v[j];
    }
    const float uu = u01_hash_dec(p.seed_dev ? *p.seed_dev : p.seed, (unsigned)row, (unsigned)i) * tot;
    float cum = 0.f;
    bool found = false;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      if (j >= p.V) break;
      cum += v[j];
      if (!found && cum > uu) { tok = j; found = true; }
    }
  }
  p.w_cur[row] = (unsigned char)tok;
  {   // one-hot of the chosen token into the next step's layer-0 operand row
    uint4* oh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.xh[0][nxt]) + (long long)row * p.K0);
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
  if (!p.done[row]) {
    p.x[(long long)row * p.max_len + i] = (unsigned char)tok;
    if (tok == p.eos) { p.end[row] = i + 1; p.done[row] = 1; }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
decode_persist_kernel(const __grid_constant__ DecodeMaps maps, const DecodeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], NUM_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int L = p.L, U = p.units_per_step, tm = p.tiles_m;
  const long long total = (long long)(p.max_len - 1) * U;

  if (warp == 0) {
    // ===================== TMA producer: waits for the rows of a unit to be ready, then streams its operands =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long g = blockIdx.x; g < total; g += gridDim.x) {
        const int i = 1 + (int)(g / U);
        const unsigned int e = p.sched[(int)(g % U)];
        const int layer = (int)(e >> 24), n = (int)((e >> 16) & 0xFF), m = (int)(e & 0xFFFF);
        const int cur = i & 1, nxt = cur ^ 1;
        const bool head = layer == L;
        if (head) {
          if (!wait_count(p.counters + (L - 1) * tm + m, 64u * (unsigned)i, p.err_flag)) goto done;
        } else {
          // own layer's previous step (h of these rows, and nobody still reads what this unit overwrites)
          if (i > 1 && !wait_count(p.counters + layer * tm + m, 64u * (unsigned)(i - 1), p.err_flag)) goto done;
          // input of this step: the layer below, or (layer 0) the token the head chose at the previous step
          if (layer > 0) { if (!wait_count(p.counters + (layer - 1) * tm + m, 64u * (unsigned)i, p.err_flag)) goto done; }
          else if (i > 1) { if (!wait_count(p.counters + L * tm + m, 8u * (unsigned)(i - 1), p.err_flag)) goto done; }
          // the x half of the next layer's operand (written by this unit) was last read two steps ago
          if (layer + 1 < L && i > 2 && !wait_count(p.counters + (layer + 1) * tm + m, 64u * (unsigned)(i - 2), p.err_flag)) goto done;
        }
        ptx::fence_proxy_async_all();   // other CTAs' generic-proxy stores (acquired above) before this CTA's TMA reads
        const CUtensorMap* mA = head ? &maps.AH[nxt] : &maps.A[layer][cur];
        const CUtensorMap* mB = head ? &maps.BH : &maps.B[layer];
        const int K = head ? p.Hd : (layer == 0 ? p.K0 : 2 * p.Hd);
        const uint32_t bytes = (uint32_t)(A_STAGE_BYTES + (head ? BN_HEAD : BN_CELL) * BK * 2);
        for (int kb = 0; kb < K / BK; ++kb) {
          if (!wait_bar(&empty_bar[s], ph ^ 1, p.err_flag)) goto done;
          ptx::mbar_arrive_expect_tx(&full_bar[s], bytes);
          ptx::tma_load_3d(sA + s * A_STAGE_BYTES, mA, &full_bar[s], kb * BK, m * BM, 0);
          ptx::tma_load_3d(sB + s * B_STAGE_BYTES, mB, &full_bar[s], kb * BK, head ? 0 : n * BN_CELL, 0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc_cell = ptx::umma_idesc_bf16(BM, BN_CELL, 0, 0);
      constexpr uint32_t idesc_head = ptx::umma_idesc_bf16(BM, BN_HEAD, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (long long g = blockIdx.x; g < total; g += gridDim.x) {
        const unsigned int e = p.sched[(int)(g % U)];
        const int layer = (int)(e >> 24);
        const bool head = layer == L;
        const int K = head ? p.Hd : (layer == 0 ? p.K0 : 2 * p.Hd);
        const uint32_t idesc = head ? idesc_head : idesc_cell;
        if (!wait_bar(&tempty_bar[acc], acc_ph ^ 1, p.err_flag)) goto done;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_CELL;
        for (int kb = 0; kb < K / BK; ++kb) {
          if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = ptx::umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = ptx::umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::tc_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        ptx::tc_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps: GRU cell / token choice out of TMEM, then publish the unit =====================
    const int q = warp & 3;              // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;   // which half of the tile's units / whether this warp owns the head's rows
    int acc = 0;
    uint32_t acc_ph = 0;
    for (long long g = blockIdx.x; g < total; g += gridDim.x) {
      const int i = 1 + (int)(g / U);
      const unsigned int e = p.sched[(int)(g % U)];
      const int layer = (int)(e >> 24), n = (int)((e >> 16) & 0xFF), m = (int)(e & 0xFFFF);
      const int cur = i & 1, nxt = cur ^ 1;
      const bool head = layer == L;
      if (!wait_bar(&tfull_bar[acc], acc_ph, p.err_flag)) goto done;
      ptx::tc_fence_after();
      const int row = m * BM + q * 32 + lane;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN_CELL;
      auto release_acc = [&]() {     // this warp has read its part of accumulator `acc` out of TMEM
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_relaxed_local(&tempty_bar[acc]);
      };
      if (!head) cell_epilogue(p, tacc, layer, n, row, chalf, cur, nxt, release_acc);
      else if (chalf == 0) head_epilogue(p, tacc, row, i, nxt, release_acc);
      else release_acc();
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      // publish: this warp's stores of the unit are ordered before the counter increment (warp barrier + release).  (Deferring
      // this into the next unit's epilogue, where the fence would find the stores long completed, was measured slower:
      // 771 k -> 749 k SMILES/s -- the units that wait for the counter wait longer than the fence costs.)
      __syncwarp();
      if (lane == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.counters + layer * tm + m), "r"(1u) : "memory");
    }
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================================
// Version 2: CTA PAIRS (tcgen05.mma.cta_group::2).  A unit is 256 rows x 256 columns: each CTA of a pair loads its own 128
// operand rows and HALF of the weight tile (the MMA reads both halves out of both CTAs' shared memory), so a CTA streams
// 32 KB per k-block instead of 48 KB (the decode is bound by operand delivery from L2, not by the tensor pipe) and six
// stages fit.  Each CTA runs the epilogue for its own 128 rows out of its own TMEM.
// =====================================================================================================================
constexpr int STAGES2 = 6;
constexpr int B2_STAGE_BYTES = (BN_CELL / 2) * BK * 2;   // 16 KB: this CTA's half of the weight tile
constexpr int SMEM2_BYTES = STAGES2 * (A_STAGE_BYTES + B2_STAGE_BYTES) + 1024 + 256;
constexpr int HEAD_LAG2 = 10;   // pair tiles (74 pairs / 8 units per pair tile in flight)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(ptx::smem_u32(holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   ptx::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// schedule of the pair kernel: same order as build_sched_kernel over pair tiles (256 rows)
__global__ void build_sched2_kernel(unsigned int* sched, int tiles2, int L) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int u = 0;
  for (int l = 0; l + 1 < L; ++l)
    for (int m = 0; m < tiles2; ++m)
      for (int n = 0; n < 8; ++n) sched[u++] = ((unsigned)l << 24) | ((unsigned)n << 16) | (unsigned)m;
  const int lag = tiles2 > HEAD_LAG2 ? HEAD_LAG2 : tiles2;
  for (int m = 0; m < tiles2; ++m) {
    for (int n = 0; n < 8; ++n) sched[u++] = ((unsigned)(L - 1) << 24) | ((unsigned)n << 16) | (unsigned)m;
    if (m >= lag) sched[u++] = ((unsigned)L << 24) | (unsigned)(m - lag);
  }
  for (int m = tiles2 - lag; m < tiles2; ++m) sched[u++] = ((unsigned)L << 24) | (unsigned)m;
}

struct DecodeMaps2 {
  CUtensorMap A[4][2];   // layer operands by parity, box 128 rows
  CUtensorMap AH[2];     // top layer's h half by parity
  CUtensorMap B[4];      // permuted cell weights, box 128 rows (half a tile)
  CUtensorMap BH;        // vocabulary weights, box 32 rows (half of the 64-row tile)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
decode_persist2_kernel(const __grid_constant__ DecodeMaps2 maps, const DecodeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES2 * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * (A_STAGE_BYTES + B2_STAGE_BYTES));
  uint64_t* full_bar = bars;                       // leader: both CTAs' stage s landed (tx bytes)
  uint64_t* empty_bar = bars + STAGES2;            // every CTA: the MMAs reading stage s have completed
  uint64_t* tfull_bar = bars + 2 * STAGES2;        // every CTA: accumulator complete
  uint64_t* tempty_bar = bars + 2 * STAGES2 + 2;   // leader: the 16 epilogue warps of the pair drained the accumulator
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 2 * NUM_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(tmem_holder, 512);
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int L = p.L, U = p.units_per_step, tm2 = p.tiles_m;   // tiles_m counts PAIR tiles (256 rows) here
  const long long total = (long long)(p.max_len - 1) * U;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own operand rows, own half of the weight tile) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long g = pair; g < total; g += npairs) {
        const int i = 1 + (int)(g / U);
        const unsigned int e = p.sched[(int)(g % U)];
        const int layer = (int)(e >> 24), n = (int)((e >> 16) & 0xFF), m = (int)(e & 0xFFFF);
        const int cur = i & 1, nxt = cur ^ 1;
        const bool head = layer == L;
        if (head) {
          if (!wait_count(p.counters + (L - 1) * tm2 + m, 128u * (unsigned)i, p.err_flag)) goto done;
        } else {
          if (i > 1 && !wait_count(p.counters + layer * tm2 + m, 128u * (unsigned)(i - 1), p.err_flag)) goto done;
          if (layer > 0) { if (!wait_count(p.counters + (layer - 1) * tm2 + m, 128u * (unsigned)i, p.err_flag)) goto done; }
          else if (i > 1) { if (!wait_count(p.counters + L * tm2 + m, 16u * (unsigned)(i - 1), p.err_flag)) goto done; }
          if (layer + 1 < L && i > 2 && !wait_count(p.counters + (layer + 1) * tm2 + m, 128u * (unsigned)(i - 2), p.err_flag)) goto done;
        }
        ptx::fence_proxy_async_all();
        const CUtensorMap* mA = head ? &maps.AH[nxt] : &maps.A[layer][cur];
        const CUtensorMap* mB = head ? &maps.BH : &maps.B[layer];
        const int K = head ? p.Hd : (layer == 0 ? p.K0 : 2 * p.Hd);
        const int bhalf = (head ? BN_HEAD : BN_CELL) / 2;                       // weight rows this CTA loads
        const uint32_t bytes = 2u * (uint32_t)(A_STAGE_BYTES + bhalf * BK * 2);   // both CTAs' loads report to the leader
        for (int kb = 0; kb < K / BK; ++kb) {
          if (!wait_bar(&empty_bar[s], ph ^ 1, p.err_flag)) goto done;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], bytes);
          const uint32_t fb = mapa(ptx::smem_u32(&full_bar[s]), 0u);
          tma_load_3d_2sm(sA + s * A_STAGE_BYTES, mA, fb, kb * BK, m * 2 * BM + (int)rank * BM, 0);
          tma_load_3d_2sm(sB + s * B2_STAGE_BYTES, mB, fb, kb * BK, (head ? 0 : n * BN_CELL) + (int)rank * bhalf, 0);
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader, one thread) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc_cell = ptx::umma_idesc_bf16(2 * BM, BN_CELL, 0, 0);
      constexpr uint32_t idesc_head = ptx::umma_idesc_bf16(2 * BM, BN_HEAD, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (long long g = pair; g < total; g += npairs) {
        const unsigned int e = p.sched[(int)(g % U)];
        const int layer = (int)(e >> 24);
        const bool head = layer == L;
        const int K = head ? p.Hd : (layer == 0 ? p.K0 : 2 * p.Hd);
        const uint32_t idesc = head ? idesc_head : idesc_cell;
        if (!wait_bar(&tempty_bar[acc], acc_ph ^ 1, p.err_flag)) goto done;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_CELL;
        for (int kb = 0; kb < K / BK; ++kb) {
          if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + s * B2_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = ptx::umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = ptx::umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma2_bf16(d_tmem, adesc, bdesc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          commit2_mc(&empty_bar[s], (uint16_t)3);   // stage s is free in both CTAs once these MMAs have read it
          if (++s == STAGES2) { s = 0; ph ^= 1; }
        }
        commit2_mc(&tfull_bar[acc], (uint16_t)3);   // accumulator complete -> both CTAs' epilogues
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (every CTA: its own 128 rows) =====================
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_ph = 0;
    const uint32_t tempty_leader = mapa(ptx::smem_u32(&tempty_bar[0]), 0u);
    for (long long g = pair; g < total; g += npairs) {
      const int i = 1 + (int)(g / U);
      const unsigned int e = p.sched[(int)(g % U)];
      const int layer = (int)(e >> 24), n = (int)((e >> 16) & 0xFF), m = (int)(e & 0xFFFF);
      const int cur = i & 1, nxt = cur ^ 1;
      const bool head = layer == L;
      if (!wait_bar(&tfull_bar[acc], acc_ph, p.err_flag)) goto done;
      ptx::tc_fence_after();
      const int row = m * 2 * BM + (int)rank * BM + q * 32 + lane;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN_CELL;
      auto release_acc = [&]() {     // this warp has read its part of accumulator `acc` out of TMEM -> pair leader's MMA thread
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) remote_arrive_relaxed(tempty_leader + (uint32_t)(acc * 8));
      };
      if (!head) cell_epilogue(p, tacc, layer, n, row, chalf, cur, nxt, release_acc);
      else if (chalf == 0) head_epilogue(p, tacc, row, i, nxt, release_acc);
      else release_acc();
      __syncwarp();
      if (lane == 0)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.counters + layer * tm2 + m), "r"(1u) : "memory");
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while the peer may still signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

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
// K-major bf16 operand [rows][cols] (row stride ld elements) as a 3-D map {cols, rows, 1}, box {64, box_rows, 1}, SWIZZLE_128B
int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MVAE_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 7)) return MVAE_ERR_INVALID;
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * rows * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_DRIVER;
}

}  // namespace

size_t mvae_decode_persistent_scratch_bytes(int Bp, int L) {
  const size_t tm = (size_t)Bp / BM;
  return ((size_t)(L + 1) * tm + (size_t)(8 * L + 1) * tm) * 4 + 512;
}

// pair version: clusters of two CTAs, units of 256 rows
static int launch_pairs(const mvae_decode_args* a, cudaStream_t st) {
  if (a->Bp % (2 * BM)) return MVAE_ERR_UNSUPPORTED;
  const int tm2 = a->Bp / (2 * BM);
  DecodeMaps2 maps;
  for (int l = 0; l < a->L; ++l) {
    const long long ldx = l == 0 ? a->K0 : 2ll * a->Hd;
    for (int k = 0; k < 2; ++k) {
      int rc = make_map(&maps.A[l][k], a->xh[l][k], a->Bp, ldx, ldx, BM);
      if (rc) return rc;
    }
    int rc = make_map(&maps.B[l], a->Wcat[l], 4ll * a->Hd, ldx, ldx, BN_CELL / 2);
    if (rc) return rc;
  }
  for (int l = a->L; l < 4; ++l) { maps.A[l][0] = maps.A[0][0]; maps.A[l][1] = maps.A[0][1]; maps.B[l] = maps.B[0]; }
  for (int k = 0; k < 2; ++k) {
    const __nv_bfloat16* top = reinterpret_cast<const __nv_bfloat16*>(a->xh[a->L - 1][k]) + a->Hd;
    int rc = make_map(&maps.AH[k], top, a->Bp, a->Hd, 2ll * a->Hd, BM);
    if (rc) return rc;
  }
  {
    int rc = make_map(&maps.BH, a->Wfc, 64, a->Hd, a->Hd, BN_HEAD / 2);
    if (rc) return rc;
  }
  DecodeParams p{};
  p.B = a->B; p.Bp = a->Bp; p.Hd = a->Hd; p.L = a->L; p.V = a->V; p.K0 = a->K0; p.max_len = a->max_len; p.eos = a->eos;
  p.mode = a->mode; p.tiles_m = tm2; p.units_per_step = (8 * a->L + 1) * tm2; p.inv_temp = a->inv_temp; p.seed = a->seed;
  p.seed_dev = a->seed_dev;
  for (int l = 0; l < 4; ++l) {
    p.bcat[l] = a->bcat[l];
    for (int k = 0; k < 2; ++k) { p.xh[l][k] = a->xh[l][k]; p.hm[l][k] = a->hm[l][k]; }
  }
  p.bfc = a->bfc; p.w_cur = a->w_cur; p.x = a->x; p.end = a->end; p.done = a->done; p.counters = a->counters; p.sched = a->sched;
  p.err_flag = a->err_flag;
  int dev = 0, sms = 0;
  MVAE_CUDA_CHECK(cudaGetDevice(&dev));
  MVAE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static size_t attr_cache[64] = {0};
  MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(decode_persist2_kernel), SMEM2_BYTES, attr_cache));
  const long long total = (long long)(a->max_len - 1) * p.units_per_step;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(NUM_THREADS, 1, 1); cfg.dynamicSmemBytes = SMEM2_BYTES; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
  // every pair must be resident at once (units wait for units of other pairs)
  cfg.gridDim = dim3(2 * (sms / 2), 1, 1);
  int max_clusters = 0;
  MVAE_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&max_clusters, decode_persist2_kernel, &cfg));
  long long pairs = max_clusters < sms / 2 ? max_clusters : sms / 2;
  if (pairs > total) pairs = total;
  if (pairs < 1) return MVAE_ERR_UNSUPPORTED;
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
  MVAE_CUDA_CHECK(cudaMemsetAsync(a->counters, 0, (size_t)(a->L + 1) * tm2 * 4, st));
  build_sched2_kernel<<<1, 32, 0, st>>>(a->sched, tm2, a->L);
  MVAE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, decode_persist2_kernel, maps, p));
  return MVAE_OK;
}

int mvae_decode_persistent_launch(const mvae_decode_args* a, cudaStream_t st) {
  if (!a || a->L < 2 || a->L > 4 || a->V < 1 || a->V > 64 || (a->Hd & 63) || (a->Bp % BM) || (a->K0 & 63) || a->K0 <= a->Hd ||
      a->max_len < 2 || !a->counters || !a->sched)
    return MVAE_ERR_UNSUPPORTED;
  if (a->variant == 2) {
    const int rc = launch_pairs(a, st);
    if (rc != MVAE_ERR_UNSUPPORTED) return rc;
  }
  const int tm = a->Bp / BM;
  if (tm > 65535) return MVAE_ERR_UNSUPPORTED;
  DecodeMaps maps;
  for (int l = 0; l < a->L; ++l) {
    const long long ldx = l == 0 ? a->K0 : 2ll * a->Hd;
    for (int k = 0; k < 2; ++k) {
      int rc = make_map(&maps.A[l][k], a->xh[l][k], a->Bp, ldx, ldx, BM);
      if (rc) return rc;
    }
    int rc = make_map(&maps.B[l], a->Wcat[l], 4ll * a->Hd, ldx, ldx, BN_CELL);
    if (rc) return rc;
  }
  for (int l = a->L; l < 4; ++l) { maps.A[l][0] = maps.A[0][0]; maps.A[l][1] = maps.A[0][1]; maps.B[l] = maps.B[0]; }
  for (int k = 0; k < 2; ++k) {
    const __nv_bfloat16* top = reinterpret_cast<const __nv_bfloat16*>(a->xh[a->L - 1][k]) + a->Hd;   // the h half of [x | h]
    int rc = make_map(&maps.AH[k], top, a->Bp, a->Hd, 2ll * a->Hd, BM);
    if (rc) return rc;
  }
  {
    int rc = make_map(&maps.BH, a->Wfc, 64, a->Hd, a->Hd, BN_HEAD);
    if (rc) return rc;
  }
  DecodeParams p{};
  p.B = a->B; p.Bp = a->Bp; p.Hd = a->Hd; p.L = a->L; p.V = a->V; p.K0 = a->K0; p.max_len = a->max_len; p.eos = a->eos;
  p.mode = a->mode; p.tiles_m = tm; p.units_per_step = (8 * a->L + 1) * tm; p.inv_temp = a->inv_temp; p.seed = a->seed;
  p.seed_dev = a->seed_dev;
  for (int l = 0; l < 4; ++l) {
    p.bcat[l] = a->bcat[l];
    for (int k = 0; k < 2; ++k) { p.xh[l][k] = a->xh[l][k]; p.hm[l][k] = a->hm[l][k]; }
  }
  p.bfc = a->bfc; p.w_cur = a->w_cur; p.x = a->x; p.end = a->end; p.done = a->done; p.counters = a->counters; p.sched = a->sched;
  p.err_flag = a->err_flag;
  int dev = 0, sms = 0;
  MVAE_CUDA_CHECK(cudaGetDevice(&dev));
  MVAE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static size_t attr_cache[64] = {0};
  MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(decode_persist_kernel), SMEM_BYTES, attr_cache));
  // every CTA must be resident at once (units wait for units of other CTAs): one CTA per SM by shared memory, grid <= #SMs
  int per_sm = 0;
  MVAE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persist_kernel, NUM_THREADS, SMEM_BYTES));
  if (per_sm < 1) return MVAE_ERR_UNSUPPORTED;
  const long long total = (long long)(a->max_len - 1) * p.units_per_step;
  const int grid = (int)(total < sms ? total : sms);
  MVAE_CUDA_CHECK(cudaMemsetAsync(a->counters, 0, (size_t)(a->L + 1) * tm * 4, st));
  build_sched_kernel<<<1, 32, 0, st>>>(a->sched, tm, a->L);
  decode_persist_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(maps, p);
  MVAE_CUDA_CHECK(cudaGetLastError());
  return MVAE_OK;
}
