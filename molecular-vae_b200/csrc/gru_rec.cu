// Persistent fused GRU recurrence for sm_100a (forward sweep and backward-through-time sweep).
//
// One cooperative launch runs ALL T steps of one layer.  CTA (j, y) owns hidden units [j*RU, (j+1)*RU) for
// NTILES batch tiles of 128 molecules.  Its slice of W_hh stays resident in shared memory for the whole
// sweep (forward: the 3*RU gate rows x Hp;  backward: RU rows of W_hh^T x 3Hp -- both 3*RU*Hp bf16), the
// per-step operand (h_{t-1} forward, dgh_{t+1} backward) streams through a TMA ring, tcgen05.mma accumulates
// in TMEM, and the epilogue warps do the gate math straight out of TMEM:
//   forward : r,z,n gates, h_t = (1-z) n + z h_{t-1}; fp32 master copy of h lives in TMEM columns (never in HBM),
//             bf16 h_t goes to hs[t+1] (next step's operand, next layer's input), (r,z,n,W_hn h+b_hn) to sv[t].
//   backward: dh_t = dgh_{t+1} W_hh + dh_{t+1}*z_{t+1} (fp32 carry in TMEM) + dX[t]; writes dG[t].
// The nj = Hp/RU CTAs that share a batch tile exchange their slices through L2 and synchronise with one
// global counter per tile (release add after the epilogue, acquire spin before the next step's TMA), so the
// tiles of a CTA are independent recurrences whose epilogue / exchange latency overlaps the other tiles' MMAs.
// All waits are bounded (2 s) and report through err_flag instead of hanging the device.
#include "common.cuh"
#include "gru_rec.h"
#include "rec_common.cuh"

namespace {

constexpr int A_STAGE_BYTES = 128 * 64 * 2;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int EPI_BAR_ID = 1;

struct RecParams {
  int Bp, Hp, T, tiles_total, nj;
  const __nv_bfloat16* gi;   // fwd: [T or 1][Bp][3Hp]
  long long gi_tstride;      // elements between steps (0 for the time-invariant layer-0 projection)
  const float* bhh;          // fwd: [3Hp]
  __nv_bfloat16* hs;         // [(T+1)][Bp][Hp]
  __nv_bfloat16* sv;         // [T][Bp][4Hp] (r,z,n,ghn)   fwd: written (may be null), bwd: read
  const __nv_bfloat16* dX;   // bwd: [T][Bp][Hp]
  __nv_bfloat16* dG;         // bwd: [T][Bp][4Hp]
  unsigned int* counters;    // [tiles_total], zeroed before launch
  int* err_flag;
  unsigned long long* trace; // optional [T][NTILES][8] timestamps of CTA (0,0) (debug)
};

using namespace rec;

template <int RU, int NTILES, int STAGES, bool BWD> struct Cfg {
  static constexpr int NB = BWD ? RU : 3 * RU;           // MMA N = resident B rows
  static constexpr int CHUNK_BYTES = NB * 128;           // one 64-wide K chunk of the resident operand
  static constexpr int ACC_COLS = NB;                    // per tile
  static constexpr int MASTER_COL0 = NTILES * NB;        // fp32 master h (fwd) / dh carry (bwd)
  static constexpr int TMEM_USED = NTILES * (NB + RU);
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128
                                   : TMEM_USED <= 256 ? 256 : 512;
  static constexpr int UH = RU / 2;                      // units per epilogue thread (2 threads per row)
  static_assert(TMEM_USED <= 512, "TMEM overflow");
  static_assert(UH % 16 == 0, "epilogue works in 16-unit chunks");
};

template <int RU, int NTILES, int STAGES, bool BWD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gru_rec_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const RecParams p) {
  using C = Cfg<RU, NTILES, STAGES, BWD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = (BWD ? 3 * p.Hp : p.Hp) / 64;  // K chunks of the recurrent GEMM
  uint8_t* sW = smem;
  uint8_t* sA = smem + (size_t)KC * C::CHUNK_BYTES;
  uint8_t* tail = sA + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + NTILES;
  uint64_t* wfull_bar = tempty_bar + NTILES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  float* sBias = reinterpret_cast<float*>(tmem_holder + 2);  // fwd: [3][RU]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * RU;                      // first hidden unit of this CTA
  const int tile0 = blockIdx.y * NTILES;
  const int ntiles = min(NTILES, p.tiles_total - tile0);

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&tmW);
    ptx::tma_prefetch_desc(&tmA);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < NTILES; ++i) { ptx::mbar_init(&tfull_bar[i], 1); ptx::mbar_init(&tempty_bar[i], NUM_EPI_WARPS * 32); }
    ptx::mbar_init(wfull_bar, 1);
    ptx::fence_mbar_init();
  }
  if (!BWD) for (int i = threadIdx.x; i < 3 * RU; i += NUM_THREADS) sBias[i] = p.bhh[(i / RU) * p.Hp + j0 + (i % RU)];
  if (warp == 1) { ptx::tmem_alloc(tmem_holder, C::TMEM_COLS); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // resident weight slice: KC chunks of [NB rows x 64 k]
      ptx::mbar_arrive_expect_tx(wfull_bar, (uint32_t)(KC * C::CHUNK_BYTES));
      for (int kc = 0; kc < KC; ++kc) {
        uint8_t* dst = sW + (size_t)kc * C::CHUNK_BYTES;
        if (BWD) {
          tma_load_2d(dst, &tmW, wfull_bar, kc * 64, j0);                 // W_hh^T rows j0.. (RU rows)
        } else {
#pragma unroll
          for (int g = 0; g < 3; ++g) tma_load_2d(dst + g * RU * 128, &tmW, wfull_bar, kc * 64, g * p.Hp + j0);
        }
      }
      int s = 0; uint32_t ph = 0;
      for (int step = 1; step < p.T; ++step) {           // step 0 has a zero operand: no MMA
        const int slab = BWD ? (p.T - step) : step;      // fwd: hs slab t ; bwd: dG slab t+1 (t = T-1-step)
        for (int i = 0; i < ntiles; ++i) {
          if (!wait_counter(p.counters + tile0 + i, (unsigned)(p.nj * step), p.err_flag)) goto done;
          ptx::fence_proxy_async_all();
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 8 + 0] = gtime();
          for (int kc = 0; kc < KC; ++kc) {
            if (!wait_bar(&empty_bar[s], ph ^ 1, p.err_flag)) goto done;
            ptx::mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES);
            ptx::tma_load_3d(sA + s * A_STAGE_BYTES, &tmA, &full_bar[s], kc * 64, (tile0 + i) * 128, slab);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 8 + 1] = gtime();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, C::NB, 0, 0);
      if (!wait_bar(wfull_bar, 0, p.err_flag)) goto done;
      int s = 0; uint32_t ph = 0;
      for (int step = 0; step < p.T; ++step) {
        for (int i = 0; i < ntiles; ++i) {
          if (step > 0) {
            // epilogue of step-1 (its arrivals complete phase step-1) has drained accumulator i
            if (!wait_bar(&tempty_bar[i], (uint32_t)((step - 1) & 1), p.err_flag)) goto done;
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + i * C::ACC_COLS;
            for (int kc = 0; kc < KC; ++kc) {
              if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
              ptx::tc_fence_after();
              if (p.trace && kc == 0 && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 8 + 2] = gtime();
              const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE_BYTES);
              const uint32_t b_addr = ptx::smem_u32(sW + (size_t)kc * C::CHUNK_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t adesc = ptx::umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                const uint64_t bdesc = ptx::umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, (kc > 0 || k > 0) ? 1u : 0u);
              }
              ptx::tc_commit(&empty_bar[s]);
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          }
          ptx::tc_commit(&tfull_bar[i]);
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 8 + 3] = gtime();
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 2) >> 2;       // which half of the RU units
    const int u0 = half * C::UH;            // first unit (within the CTA slice) of this thread
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t fph = 0;
    for (int step = 0; step < p.T; ++step) {
      const int t = BWD ? (p.T - 1 - step) : step;
      for (int i = 0; i < ntiles; ++i) {
        const int tile = tile0 + i;
        const long long row = (long long)tile * 128 + q * 32 + lane;
        // ---- prefetch this thread's global operands before waiting for the accumulator ----
        uint4 pre[BWD ? (6 * C::UH / 8) : (3 * C::UH / 8)];
        if (!BWD) {
          const __nv_bfloat16* g = p.gi + (long long)t * p.gi_tstride + row * 3 * p.Hp + j0 + u0;
#pragma unroll
          for (int gate = 0; gate < 3; ++gate)
#pragma unroll
            for (int v = 0; v < C::UH / 8; ++v)
              pre[gate * (C::UH / 8) + v] = __ldg(reinterpret_cast<const uint4*>(g + (long long)gate * p.Hp) + v);
        } else {
          const __nv_bfloat16* s4 = p.sv + ((long long)t * p.Bp + row) * 4 * p.Hp + j0 + u0;
#pragma unroll
          for (int blk = 0; blk < 4; ++blk)
#pragma unroll
            for (int v = 0; v < C::UH / 8; ++v)
              pre[blk * (C::UH / 8) + v] = __ldg(reinterpret_cast<const uint4*>(s4 + (long long)blk * p.Hp) + v);
          const __nv_bfloat16* hp = p.hs + ((long long)t * p.Bp + row) * p.Hp + j0 + u0;
          const __nv_bfloat16* dx = p.dX + ((long long)t * p.Bp + row) * p.Hp + j0 + u0;
#pragma unroll
          for (int v = 0; v < C::UH / 8; ++v) {
            pre[4 * (C::UH / 8) + v] = __ldg(reinterpret_cast<const uint4*>(hp) + v);
            pre[5 * (C::UH / 8) + v] = __ldg(reinterpret_cast<const uint4*>(dx) + v);
          }
        }
        (void)wait_bar(&tfull_bar[i], fph, p.err_flag);  // on failure keep walking: every barrier below must be reached
        ptx::tc_fence_after();
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && warp == 2 && lane == 0;
        if (tr) p.trace[((size_t)step * NTILES + i) * 8 + 4] = gtime();
#pragma unroll
        for (int c = 0; c < C::UH; c += 16) {
          const int uc = u0 + c;  // unit offset inside the CTA slice
          const uint32_t master_addr = tmem_base + lane_off + C::MASTER_COL0 + i * RU + uc;
          if (!BWD) {
            uint32_t ar[16], az[16], an[16], hm[16];
            if (step > 0) {
              const uint32_t acc = tmem_base + lane_off + i * C::ACC_COLS + uc;
              ptx::tmem_ld_32x16(acc, ar);
              ptx::tmem_ld_32x16(acc + RU, az);
              ptx::tmem_ld_32x16(acc + 2 * RU, an);
              ptx::tmem_ld_32x16(master_addr, hm);
              ptx::tmem_ld_wait();
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) { ar[k] = 0u; az[k] = 0u; an[k] = 0u; hm[k] = 0u; }
            }
            float gr[16], gz[16], gn[16];
            unpack16(pre[0 * (C::UH / 8) + c / 8], pre[0 * (C::UH / 8) + c / 8 + 1], gr);
            unpack16(pre[1 * (C::UH / 8) + c / 8], pre[1 * (C::UH / 8) + c / 8 + 1], gz);
            unpack16(pre[2 * (C::UH / 8) + c / 8], pre[2 * (C::UH / 8) + c / 8 + 1], gn);
            float h[16], r[16], z[16], n[16], ghn[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float hr = __uint_as_float(ar[k]) + sBias[0 * RU + uc + k];
              const float hz = __uint_as_float(az[k]) + sBias[1 * RU + uc + k];
              ghn[k] = __uint_as_float(an[k]) + sBias[2 * RU + uc + k];
              r[k] = gate_sigmoid(gr[k] + hr);
              z[k] = gate_sigmoid(gz[k] + hz);
              n[k] = gate_tanh(fmaf(r[k], ghn[k], gn[k]));
              const float hp = __uint_as_float(hm[k]);
              h[k] = fmaf(z[k], hp - n[k], n[k]);
            }
            tmem_st_32x16(master_addr, h);
            st16(p.hs + ((long long)(t + 1) * p.Bp + row) * p.Hp + j0 + uc, h);
            if (p.sv) {
              __nv_bfloat16* s4 = p.sv + ((long long)t * p.Bp + row) * 4 * p.Hp + j0 + uc;
              st16(s4, r);
              st16(s4 + p.Hp, z);
              st16(s4 + 2 * p.Hp, n);
              st16(s4 + 3 * p.Hp, ghn);
            }
          } else {
            uint32_t acc[16], cm[16];
            if (step > 0) {
              ptx::tmem_ld_32x16(tmem_base + lane_off + i * C::ACC_COLS + uc, acc);
              ptx::tmem_ld_32x16(master_addr, cm);
              ptx::tmem_ld_wait();
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) { acc[k] = 0u; cm[k] = 0u; }
            }
            float r[16], z[16], n[16], ghn[16], hp[16], dx[16];
            unpack16(pre[0 * (C::UH / 8) + c / 8], pre[0 * (C::UH / 8) + c / 8 + 1], r);
            unpack16(pre[1 * (C::UH / 8) + c / 8], pre[1 * (C::UH / 8) + c / 8 + 1], z);
            unpack16(pre[2 * (C::UH / 8) + c / 8], pre[2 * (C::UH / 8) + c / 8 + 1], n);
            unpack16(pre[3 * (C::UH / 8) + c / 8], pre[3 * (C::UH / 8) + c / 8 + 1], ghn);
            unpack16(pre[4 * (C::UH / 8) + c / 8], pre[4 * (C::UH / 8) + c / 8 + 1], hp);
            unpack16(pre[5 * (C::UH / 8) + c / 8], pre[5 * (C::UH / 8) + c / 8 + 1], dx);
            float dan[16], dar[16], daz[16], danr[16], carry[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float dh = __uint_as_float(acc[k]) + __uint_as_float(cm[k]) + dx[k];
              dan[k] = dh * (1.f - z[k]) * (1.f - n[k] * n[k]);
              daz[k] = dh * (hp[k] - n[k]) * z[k] * (1.f - z[k]);
              dar[k] = dan[k] * ghn[k] * r[k] * (1.f - r[k]);
              danr[k] = dan[k] * r[k];
              carry[k] = dh * z[k];
            }
            tmem_st_32x16(master_addr, carry);
            __nv_bfloat16* g4 = p.dG + ((long long)t * p.Bp + row) * 4 * p.Hp + j0 + uc;
            st16(g4, dan);
            st16(g4 + p.Hp, dar);
            st16(g4 + 2 * p.Hp, daz);
            st16(g4 + 3 * p.Hp, danr);
          }
        }
        tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty_bar[i]);
        if (tr) p.trace[((size_t)step * NTILES + i) * 8 + 5] = gtime();
        // publish this CTA's slice of the tile: all epilogue threads' global writes -> gpu scope, then one
        // release-add on the tile counter
        __threadfence();
        asm volatile("bar.sync %0, %1;" ::"n"(EPI_BAR_ID), "n"(NUM_EPI_WARPS * 32) : "memory");
        if (warp == 2 && lane == 0) {
          if (tr) p.trace[((size_t)step * NTILES + i) * 8 + 6] = gtime();
          ptx::fence_proxy_async_all();
          red_release_add(p.counters + tile, 1u);
          if (tr) p.trace[((size_t)step * NTILES + i) * 8 + 7] = gtime();
        }
      }
      fph ^= 1;
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
int encode(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MVAE_ERR_DRIVER;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_DRIVER;
}

template <int RU, int NTILES, int STAGES, bool BWD>
int launch_t(const mvae_gru_rec_args& a, cudaStream_t st) {
  using C = Cfg<RU, NTILES, STAGES, BWD>;
  const int Hp = a.Hp, Bp = a.Bp, T = a.T;
  const int KC = (BWD ? 3 * Hp : Hp) / 64;
  const size_t smem = (size_t)KC * C::CHUNK_BYTES + (size_t)STAGES * A_STAGE_BYTES + 1024 + 512 + 3 * RU * 4;
  if (smem > 232448) return MVAE_ERR_UNSUPPORTED;
  CUtensorMap tmW, tmA;
  {
    // fwd: W_hh padded [3Hp][Hp] ; bwd: W_hh^T padded [Hp][3Hp]; both K-major, box {64, RU}
    cuuint64_t dims[2] = {(cuuint64_t)(BWD ? 3 * Hp : Hp), (cuuint64_t)(BWD ? Hp : 3 * Hp)};
    cuuint64_t str[1] = {(cuuint64_t)(BWD ? 3 * Hp : Hp) * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)RU};
    int rc = encode(&tmW, a.W, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    cuuint32_t box[3] = {64, 128, 1};
    if (BWD) {
      cuuint64_t dims[3] = {(cuuint64_t)3 * Hp, (cuuint64_t)Bp, (cuuint64_t)T};
      cuuint64_t str[2] = {(cuuint64_t)4 * Hp * 2, (cuuint64_t)Bp * 4 * Hp * 2};
      int rc = encode(&tmA, a.dG + Hp, 3, dims, str, box);
      if (rc) return rc;
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)Hp, (cuuint64_t)Bp, (cuuint64_t)(T + 1)};
      cuuint64_t str[2] = {(cuuint64_t)Hp * 2, (cuuint64_t)Bp * Hp * 2};
      int rc = encode(&tmA, a.hs, 3, dims, str, box);
      if (rc) return rc;
    }
  }
  RecParams p{};
  p.Bp = Bp; p.Hp = Hp; p.T = T; p.tiles_total = Bp / 128; p.nj = Hp / RU;
  p.gi = a.gi; p.gi_tstride = a.gi_tstride; p.bhh = a.bhh; p.hs = a.hs; p.sv = a.sv; p.dX = a.dX; p.dG = a.dG;
  p.counters = a.counters; p.err_flag = a.err_flag; p.trace = a.trace;
  auto kern = gru_rec_kernel<RU, NTILES, STAGES, BWD>;
  static size_t attr_cache[64] = {0};
  MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, attr_cache));
  dim3 grid(Hp / RU, ceil_div(Bp / 128, NTILES));
  MVAE_CUDA_CHECK(cudaMemsetAsync(a.counters, 0, sizeof(unsigned int) * (Bp / 128), st));
  void* args[3] = {(void*)&tmW, (void*)&tmA, (void*)&p};
  MVAE_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(NUM_THREADS), args, smem, st));
  return MVAE_OK;
}

}  // namespace

int mvae_gru_rec_max_rows(int Hp, int variant, int num_sms) {
  const int ru = variant == 2 ? 32 : 64, nt = variant == 2 ? 4 : 2;
  if (Hp % ru || Hp > 512) return 0;
  const int nj = Hp / ru;
  return (num_sms / nj) * nt * 128;
}

int mvae_gru_rec_launch(const mvae_gru_rec_args* a, cudaStream_t stream) {
  if (!a || a->Bp % 128 || a->Hp % 64 || a->Hp > 512 || a->T < 1) return MVAE_ERR_INVALID;
  if (a->variant == 2) {
    return a->backward ? launch_t<32, 4, 8, true>(*a, stream) : launch_t<32, 4, 8, false>(*a, stream);
  }
  return a->backward ? launch_t<64, 2, 2, true>(*a, stream) : launch_t<64, 2, 2, false>(*a, stream);
}
