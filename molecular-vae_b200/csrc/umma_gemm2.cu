// 2-CTA (tcgen05.mma.cta_group::2) variant of the tcgen05 / TMA GEMM.
//
//   D[M,N] = A * B + bias[N]     bf16 operands, K-major or MN-major (as umma_gemm.cu), M % 256 == N % 256 == 0;
//   D bf16 (row-major or row-blocked), or fp32 accumulated with red.add over split-K work units (the K = T*B weight gradients)
//
// Why: the projection GEMMs (x W_ih^T over all T*B rows, K = 512) are 8 k-blocks per output tile; with one CTA per
// 128 x 256 tile every k-block costs 48 KB of operand traffic per SM and the kernel runs at the rate L2 delivers them
// (ncu: 50 % tensor-pipe active, no unit saturated).  A CTA PAIR computes a 256 x 256 tile: each CTA loads its
// own 128 rows of A and only HALF of the B tile (the MMA reads both halves out of both CTAs' shared memory), i.e. 32 KB per
// k-block and SM, and six stages fit.  Same pipeline as umma_gemm.cu otherwise: TMA producer warp, single-thread MMA issuer
// (pair leader), 8 epilogue warps per CTA that drain the CTA's own 128 rows of the accumulator (TMEM loads pipelined with the
// packed bf16 stores), two accumulators in TMEM, persistent over the output tiles.
#include "common.cuh"
#include "umma_gemm.h"
#include <stdlib.h>

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;
constexpr int A_STAGE_BYTES = BM * BK * 2;          // 16 KB: this CTA's 128 rows
constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;    // 16 KB: this CTA's half of the B tile
constexpr int STAGES = 6;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 + 256;

struct Params2 {
  int M, N, K, tiles_m, tiles_n;   // tiles_m counts 256-row pair tiles
  int splits, kb_per_split;        // split-K (fp32 output only)
  void* out; long long ldc; const float* bias; int out_rb; int out_bf16;
  int accumulate;                  // fp32 output, splits == 1: D += result (plain read-modify-write)
  int late_release;                // A/B switch (MVAE_GEMM_EARLY_RELEASE=0): hand the accumulator back after the tile's stores
  int* err_flag;
};

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
      else if (now - t0 > 2000000000ull) { if (err_flag) atomicExch(err_flag, 7); return false; }   // 2 s: report instead of hanging
      if (err_flag && *(volatile int*)err_flag) return false;
    }
  }
  return true;
}
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
// accumulator hand-back: the tcgen05 fences order the TMEM reads, no generic-proxy data rides on this arrive
__device__ __forceinline__ void remote_arrive_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params2 p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                      // leader: both CTAs' stage landed
  uint64_t* empty_bar = bars + STAGES;            // every CTA: the MMAs reading the stage completed
  uint64_t* tfull_bar = bars + 2 * STAGES;        // every CTA: accumulator complete
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;   // leader: all 16 epilogue warps of the pair drained the accumulator
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&tmA);
    ptx::tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
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

  const int kb_total = (p.K + BK - 1) / BK;
  const int n_units = p.tiles_m * p.tiles_n * p.splits;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = pair; unit < n_units; unit += npairs) {
        const int tile = unit / p.splits, split = unit - tile * p.splits;
        const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
        const int kb0 = split * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        const int m0 = m_blk * 2 * BM + (int)rank * BM, n0 = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!wait_bar(&empty_bar[s], ph ^ 1, p.err_flag)) goto done;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], 2u * (A_STAGE_BYTES + B_STAGE_BYTES));
          const uint32_t fb = mapa(ptx::smem_u32(&full_bar[s]), 0u);
          uint8_t* a_dst = sA + s * A_STAGE_BYTES;
          uint8_t* b_dst = sB + s * B_STAGE_BYTES;
          if (A_MN) {     // stored [K][M]: 64 (m) x 64 (k) boxes, 8 KB apart
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_3d_2sm(a_dst + c * 8192, &tmA, fb, m0 + c * 64, kb * BK, 0);
          } else {
            tma_load_3d_2sm(a_dst, &tmA, fb, kb * BK, m0, 0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 2 / 64; ++c) tma_load_3d_2sm(b_dst + c * 8192, &tmB, fb, n0 + c * 64, kb * BK, 0);
          } else {
            tma_load_3d_2sm(b_dst, &tmB, fb, kb * BK, n0, 0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader, one thread) =====================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int unit = pair; unit < n_units; unit += npairs) {
        const int split = unit % p.splits;
        const int kb0 = split * p.kb_per_split, kb1 = min(kb_total, kb0 + p.kb_per_split);
        if (!wait_bar(&tempty_bar[acc], acc_ph ^ 1, p.err_flag)) goto done;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major SW128: 8-row groups 1024 B apart, 32 B per K = 16 inside the swizzle row; MN-major SW128: 64-wide MN
            // chunks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), two k-groups (2048 B) per K = 16
            const uint64_t adesc = A_MN ? ptx::umma_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                        : ptx::umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? ptx::umma_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                        : ptx::umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma2_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          commit2_mc(&empty_bar[s], (uint16_t)3);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        commit2_mc(&tfull_bar[acc], (uint16_t)3);
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (every CTA: its own 128 rows x 256 columns) =====================
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int CH = BN / 2, NCH = CH / 32;
    int acc = 0;
    uint32_t acc_ph = 0;
    const uint32_t tempty_leader = mapa(ptx::smem_u32(&tempty_bar[0]), 0u);
    for (int unit = pair; unit < n_units; unit += npairs) {
      const int tile = unit / p.splits, split = unit - tile * p.splits;
      const int m_blk = tile / p.tiles_n, n_blk = tile - m_blk * p.tiles_n;
      if (!wait_bar(&tfull_bar[acc], acc_ph, p.err_flag)) goto done;
      ptx::tc_fence_after();
      const int row = m_blk * 2 * BM + (int)rank * BM + q * 32 + lane;
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + chalf * CH;
      uint32_t r0[32], r1[32];
      // the accumulator goes back to the MMA issuer as soon as this warp's LAST TMEM load of the tile has completed, i.e. before
      // the final chunk's global stores / atomics (a release arrive after them would wait for those stores)
      auto release_now = [&]() {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) remote_arrive_relaxed(tempty_leader + (uint32_t)(acc * 8));
      };
      auto release_acc = [&]() { if (!p.late_release) release_now(); };
      if (!p.out_bf16 && p.splits == 1) {
        // fp32 result of a whole-K unit: plain 16-byte stores (or read-modify-write when accumulating; nobody else owns the tile)
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          ptx::tmem_ld_32x32(tb + c * 32, r0);
          ptx::tmem_ld_wait();
          if (c == NCH - 1) release_acc();
          const int col0 = n_blk * BN + chalf * CH + c * 32;
          float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 w = make_float4(__uint_as_float(r0[j]), __uint_as_float(r0[j + 1]), __uint_as_float(r0[j + 2]), __uint_as_float(r0[j + 3]));
            if (p.bias) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              w.x += b4.x; w.y += b4.y; w.z += b4.z; w.w += b4.w;
            }
            if (p.accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(o + j);
              w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
            }
            *reinterpret_cast<float4*>(o + j) = w;
          }
        }
        if (p.late_release) release_now();
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        continue;
      }
      if (!p.out_bf16) {
        // fp32 result accumulated over the split-K units with red.add (the caller zeroes D or passes what to add onto)
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          ptx::tmem_ld_32x32(tb + c * 32, r0);
          ptx::tmem_ld_wait();
          if (c == NCH - 1) release_acc();
          float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + n_blk * BN + chalf * CH + c * 32;
          const bool add_bias = p.bias != nullptr && split == 0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            atomicAdd(o + j, __uint_as_float(r0[j]) + (add_bias ? __ldg(p.bias + n_blk * BN + chalf * CH + c * 32 + j) : 0.f));
        }
        if (p.late_release) release_now();
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        continue;
      }
      auto process = [&](const uint32_t (&rr)[32], int c) {
        const int col0 = n_blk * BN + chalf * CH + c * 32;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
          if (p.bias) {
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
        else release_acc();
        process(r0, c);
        if (c + 1 < NCH) {
          ptx::tmem_ld_wait();
          if (c + 2 < NCH) ptx::tmem_ld_32x32(tb + (c + 2) * 32, r0);
          else release_acc();
          process(r1, c + 1);
        }
      }
      if (p.late_release) release_now();
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
// K-major: dims {K, rows(MN)}, box {64, box_rows}.  MN-major (stored [K][MN]): dims {MN, K}, box {64, 64}.
int make_map(CUtensorMap* map, const mvae_umma_operand& op, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MVAE_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) || (op.ld & 7)) return MVAE_ERR_INVALID;
  const long long slabs = op.slabs > 0 ? op.slabs : 1;
  const long long rows = op.mn_major ? op.k : op.mn, cols = op.mn_major ? op.mn : op.k;
  cuuint32_t estr[3] = {1, 1, 1};
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  cuuint64_t strides[2] = {(cuuint64_t)op.ld * 2, (cuuint64_t)(slabs > 1 ? op.slab_stride : op.ld * rows) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)(op.mn_major ? 64 : box_rows), 1};
  if ((strides[0] & 15) || (strides[1] & 15)) return MVAE_ERR_INVALID;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_DRIVER;
}

}  // namespace

template <int A_MN, int B_MN>
static int launch_pairs_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params2& p, int pairs, cudaStream_t stream) {
  auto kern = umma_gemm2_kernel<A_MN, B_MN>;
  static size_t attr_cache[64] = {0};
  MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(kern), SMEM_BYTES, attr_cache));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1); cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = 1;
  MVAE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  return MVAE_OK;
}

// Returns MVAE_ERR_UNSUPPORTED (nothing enqueued) unless: single-slab operands, M and N multiples of 256, and either a bf16
// result without accumulation / split-K (16-byte aligned rows and bias) or an fp32 result that is ACCUMULATED (red.add, any
// split-K factor; the caller zeroed D or passes the value to add onto).
int mvae_umma_gemm_pairs(const mvae_umma_operand* A, const mvae_umma_operand* B, const mvae_umma_out* D, int M, int N, int K,
                         int splits, int* err_flag, cudaStream_t stream) {
  if (!A || !B || !D || A->slabs > 1 || B->slabs > 1 || A->slab || B->slab || D->act || D->rb && !D->bf16 ||
      (M % (2 * BM)) || (N % BN) || K < 1)
    return MVAE_ERR_UNSUPPORTED;
  if (splits < 1) splits = 1;
  if (D->bf16) {
    if (D->accumulate || splits > 1 || (D->ld & 7) || (reinterpret_cast<uintptr_t>(D->ptr) & 15) || (reinterpret_cast<uintptr_t>(D->bias) & 15))
      return MVAE_ERR_UNSUPPORTED;
  } else {
    if (splits > 1 ? !D->accumulate : ((D->ld & 3) || (reinterpret_cast<uintptr_t>(D->ptr) & 15) || (reinterpret_cast<uintptr_t>(D->bias) & 15)))
      return MVAE_ERR_UNSUPPORTED;   // split-K: red.add onto D; whole-K: 16-byte stores / read-modify-write
  }
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, *A, BM);
  if (rc) return rc;
  rc = make_map(&tmB, *B, BN / 2);
  if (rc) return rc;
  Params2 p{};
  p.M = M; p.N = N; p.K = K; p.tiles_m = M / (2 * BM); p.tiles_n = N / BN;
  const int kb_total = (K + BK - 1) / BK;
  if (splits > kb_total) splits = kb_total;
  p.kb_per_split = (kb_total + splits - 1) / splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.out = D->ptr; p.ldc = D->ld; p.bias = D->bias; p.out_rb = D->rb; p.out_bf16 = D->bf16; p.accumulate = D->accumulate; p.err_flag = err_flag;
  { const char* e = getenv("MVAE_GEMM_EARLY_RELEASE"); p.late_release = (e && atoi(e) == 0) ? 1 : 0; }
  int dev = 0, sms = 0;
  MVAE_CUDA_CHECK(cudaGetDevice(&dev));
  MVAE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long units = (long long)p.tiles_m * p.tiles_n * p.splits;
  long long pairs = sms / 2;
  if (pairs > units) pairs = units;
  if (!A->mn_major && !B->mn_major) return launch_pairs_t<0, 0>(tmA, tmB, p, (int)pairs, stream);
  if (!A->mn_major && B->mn_major) return launch_pairs_t<0, 1>(tmA, tmB, p, (int)pairs, stream);
  if (A->mn_major && !B->mn_major) return launch_pairs_t<1, 0>(tmA, tmB, p, (int)pairs, stream);
  return launch_pairs_t<1, 1>(tmA, tmB, p, (int)pairs, stream);
}
