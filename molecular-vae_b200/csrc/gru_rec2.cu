// Persistent fused GRU recurrence, 2-CTA tensor-core version (tcgen05 cta_group::2 + TMA multicast in clusters of 8).
//
// Why this shape (measured on B200, profiles/r01_rec_trace_v1.txt): with one CTA holding the whole W_hh slice
// of 64 hidden units (192 KB) only 32 KB of shared memory is left for the streamed operand, and the 8 CTAs that
// share a batch tile each pull the same h_{t-1} / dgh_{t+1} rows out of L2 -- the sweep ran at the L2 -> SM
// bandwidth / latency limit (~32 us per step), not at the tensor pipe.  Here
//   * a CTA PAIR owns 64 hidden units and issues M=256 MMAs (tcgen05.mma.cta_group::2): the resident weight
//     slice is split over the pair (96 KB per CTA), leaving 8 x 16 KB operand stages per CTA;
//   * 4 pairs form a cluster; the operand rows a CTA needs are also needed by the 3 same-parity CTAs of the
//     cluster, so each loads a quarter of every stage and TMA-multicasts it (L2 reads / 4);
//   * each CTA runs the gate epilogue for its own 128 rows x 64 units out of its own TMEM (16 epilogue warps),
//     fp32 master copy of h (forward) / of the dh carry (backward) stays in TMEM columns.
// Cross-pair dependencies (all 8 pairs of a row group must have published step t before step t+1 loads) go
// through per-(tile, parity) global counters with release/acquire; everything inside a cluster uses mbarriers.
// Every wait is bounded (2 s) and reports through err_flag.
//
// KS (BPTT only, "K-split"): the sweep is bound by the L2 -> SM delivery of the streamed operand (8 unit slices each
// re-read the whole [Bp x 3Hp] dgh slab: 100 MB per step).  With KS a cluster of TWO pairs owns 128 units: each pair
// keeps the W_hh^T rows of those 128 units for ONE HALF of K (same 96 KB per CTA), streams only that half of dgh
// (operand traffic halves), and the two pairs exchange the fp32 partial sums of the 64 units the other one finalises
// through distributed shared memory (32 KB per CTA per tile-step, off the L2 path).
#include "common.cuh"
#include "gru_rec.h"
#include "rec_common.cuh"

namespace {
using namespace rec;

constexpr int STAGES = 8;
constexpr int NTILES = 2;                 // pair-tiles (256 rows) per pair
constexpr int RU = 64;                    // hidden units per pair
constexpr int A_STAGE = 128 * 64 * 2;     // this CTA's 128 rows x 64 k, bf16
constexpr int EPI_WARPS = 16;
// warps 0..3 = control warpgroup (TMA producer, MMA issuer, publisher, one idle), warps 4..19 = epilogue.  The control
// warpgroup hands most of its registers to the epilogue warps (setmaxnreg), whose gate math otherwise spills at the
// 96 registers a 640-thread CTA starts with.
constexpr int CTRL_WARPS = 4;
constexpr int THREADS = (CTRL_WARPS + EPI_WARPS) * 32;
constexpr int PUB_WARP = 2;
constexpr int CTRL_REGS = 32, EPI_REGS = 112;   // (96 - 32) * 128 released >= (112 - 96) * 512 claimed

// Saved gates (forward -> BPTT): per (t, tile, pair, parity, part, q) -- i.e. per epilogue WARP -- one contiguous block of
// SV_NARR arrays x 1 KB, array a at +a*512 elements, inside it [half][lane][8 units] (thread (q, lane) owns two 16-byte
// halves 512 B apart).  Only the BPTT epilogue with the same thread mapping reads it.
//   MVAE_SV_BULK 1: the forward epilogue stages its arrays in shared memory and writes them with cp.async.bulk
//     (shared -> global): the 80 KB per tile-step no longer sit in the LSU / outstanding-store queue in front of the h'
//     stores and the release that the next step of the other CTAs waits for (profiles/r01_rec_trace_notes.txt: the plain
//     sv stores were the forward sweep's largest single cost).  Paid for with two operand stages (8 -> 6).
//   MVAE_SV_HP 0: h_{t-1} is not saved; the BPTT epilogue reads it from the hs slab (same bf16 value).  Measured (B200,
//     B=4096, T=120): forward 10.55 -> 10.45 us/step, BPTT 18.98 -> 20.58 (32 scattered 32-byte sectors per warp load
//     instead of one contiguous KB), so the copy stays (default 1).
#ifndef MVAE_SV_BULK
#define MVAE_SV_BULK 1
#endif
#ifndef MVAE_SV_HP
#define MVAE_SV_HP 1
#endif
//   MVAE_OUT_TMA 1: the row-major outputs other CTAs stream next step (forward: h' into the hs slab; BPTT: the dG blocks) are
//     staged per warp ([32 rows][32 B]) and written with TMA tensor stores instead of 32 scattered 32-byte sectors per warp
//     store instruction; the issuing lane waits for their completion before the tile is handed to the publisher.
#ifndef MVAE_OUT_TMA
#define MVAE_OUT_TMA 1
#endif
// operand stages.  Measured (B200, B=4096, T=120, us/step fwd / K-split BPTT): 6 stages 10.60 / 19.11, 5: 10.62 / 19.18,
// 4: 10.13 / 17.71, 3: 10.57 / 17.99 -- deeper prefetch only adds queueing in front of the epilogue's traffic.
#ifndef MVAE_NST_FWD
#define MVAE_NST_FWD 4
#endif
#ifndef MVAE_NST_BWD
#define MVAE_NST_BWD 3
#endif
//   MVAE_XCHG_ASYNC 1 (K-split BPTT): the partial sums cross to the partner CTA with st.async (the store itself reports its
//     bytes to the partner's mbarrier, complete_tx), so the sending warp needs no release fence -- ncu's stall sampling put
//     21 % of the epilogue warps' time into the MEMBAR + ERRBAR of the two mbarrier.arrive.release.cluster per tile-step
//     (profiles/r02_rec2_stalls.txt).  The receiving warp arms its own barrier (arrive.expect_tx) and hands the buffer back
//     with a relaxed remote arrive that is data-dependent on the values it read.  0: st.shared::cluster + release arrives.
//     Measured (B200, B=4096, T=120): K-split BPTT 17.47 -> 16.45 us/step, results identical (tools/rec_test 32).
#ifndef MVAE_XCHG_ASYNC
#define MVAE_XCHG_ASYNC 1
#endif
//   MVAE_XBUF_STAGE 1 (K-split BPTT with MVAE_OUT_TMA and MVAE_XCHG_ASYNC): once a warp has read the partner's partial sums, its
//     2 KB slot of the exchange buffer doubles as the staging area of two of the three dgh blocks; the slot is handed back to
//     the partner (xfree) only after those TMA stores have completed, and the slots are contiguous 2 KB blocks per warp.
//     Measured (B200, B=4096, T=120, us/step): 16.16 -> 15.08 at 3 operand stages; the 32 KB of dG staging it frees would
//     allow 5 stages, which are slower again (4: 15.58, 5: 15.55, 2: 16.83) -- 3 stays.
//   MVAE_XCHG_V4 1 (with MVAE_XBUF_STAGE): 16-byte st.async per lane (4 per warp and tile-step instead of 16 of 4 bytes).
//     Measured: 14.75 -> 14.87 us/step, no gain (the exchange sits at the DSMEM rate, 32 KB per CTA and tile-step), so off.
#ifndef MVAE_XCHG_V4
#define MVAE_XCHG_V4 0
#endif
#ifndef MVAE_XBUF_STAGE
#define MVAE_XBUF_STAGE 1
#endif
constexpr int SV_NARR = MVAE_SV_HP ? 5 : 4;
//   MVAE_SV_STAGE2 1 (forward, with MVAE_SV_BULK): two 2 KB staging buffers per epilogue warp, one per bulk store ([r | z] and
//     [n | W_hn h]), and the h_{t-1} copy goes out as a plain store right after the accumulators have been read -- no
//     cp.async.bulk.wait_group.read between the stores of one tile-step (ncu stall sampling: 14 % of all warp samples of the
//     forward sweep sat in those waits: the bulk engine takes ~1 us to report that it has read its source)
//     Measured (B200, B=4096, T=120): forward 10.12 -> 11.60 us/step -- the LSU store of the fifth array costs more than the
//     two waits it removes (the same per-SM outstanding-store throttling that made MVAE_SV_BULK pay), so it stays off.
#ifndef MVAE_SV_STAGE2
#define MVAE_SV_STAGE2 0
#endif
constexpr int SV_STAGE_BYTES = MVAE_SV_STAGE2 ? 4096 : 2048;      // per epilogue warp: two arrays per bulk store

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(ptx::smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
// TMA tensor store of a [32 rows][32 B] staging block (dense, no swizzle) to (col c0, row c1, slab c2)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* ssrc, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(ssrc)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// this thread's 16 bf16 (32 B) into row `lane` of a dense [32][32 B] block
__device__ __forceinline__ void sts_row32(uint8_t* blk, int lane, const float (&v)[16]) {
  uint4 a, b;
  a.x = rec::pack_bf2(v[0], v[1]); a.y = rec::pack_bf2(v[2], v[3]); a.z = rec::pack_bf2(v[4], v[5]); a.w = rec::pack_bf2(v[6], v[7]);
  b.x = rec::pack_bf2(v[8], v[9]); b.y = rec::pack_bf2(v[10], v[11]); b.z = rec::pack_bf2(v[12], v[13]); b.w = rec::pack_bf2(v[14], v[15]);
  *reinterpret_cast<uint4*>(blk + lane * 32) = a;
  *reinterpret_cast<uint4*>(blk + lane * 32 + 16) = b;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// 16 bf16 of this thread into a warp's 1 KB staging array: [half][lane][16 B]
__device__ __forceinline__ void sts2x128(uint8_t* arr_base, int lane, const float (&v)[16]) {
  uint4 a, b;
  a.x = rec::pack_bf2(v[0], v[1]); a.y = rec::pack_bf2(v[2], v[3]); a.z = rec::pack_bf2(v[4], v[5]); a.w = rec::pack_bf2(v[6], v[7]);
  b.x = rec::pack_bf2(v[8], v[9]); b.y = rec::pack_bf2(v[10], v[11]); b.z = rec::pack_bf2(v[12], v[13]); b.w = rec::pack_bf2(v[14], v[15]);
  *reinterpret_cast<uint4*>(arr_base + lane * 16) = a;
  *reinterpret_cast<uint4*>(arr_base + 512 + lane * 16) = b;
}

struct Params2 {
  int Bp, Hp, T, pair_tiles, npairs;
  const __nv_bfloat16* gi; long long gi_tstride;
  const float* bhn;          // fwd: b_hh of the n gate, padded [Hp]  (b_hr, b_hz are folded into gi)
  __nv_bfloat16* hs; __nv_bfloat16* sv;
  const __nv_bfloat16* dX; __nv_bfloat16* dG;
  unsigned int* counters;    // [pair_tiles][2]
  int* err_flag;
  unsigned long long* trace; // optional debug timestamps [T][NTILES][12] of CTA (0,0)
  int a_box_rows;            // rows per operand TMA box (128, 64 or 32): a stage is loaded as 128/a_box_rows boxes
  int ones_col;              // fwd: hidden-unit index forced to 1.0 (bias-gradient trick) or -1
  int debug;                 // timing experiments only: bit0 skip counter waits, bit1 de-share operand rows
  const float* h0;           // fwd, optional: fp32 initial state [Bp][Hp] (hs slab 0 holds its bf16 copy); null = zeros
  float* carry_out;          // bwd, optional: fp32 [Bp][Hp], receives dh_t * z_t of the last processed step (t = 0); the
                             // gradient wrt the initial state is carry_out + dgh_0 * W_hh (one GEMM, done by the caller)
  // fwd, optional token-table input projection (embedding layers: x_t W_ih^T is a row of a [V][3Hp] table): the pair's
  // slice of the table lives in shared memory (bf16) and gi(row, t) = table[tok[t][row]] (+ gi if gi != null)
  const float* tbl;          // [V][3Hp] fp32, (r,z,n) blocks, biases folded in by the caller
  const unsigned char* tok;  // [T][Bp] token ids (time-major)
  int V;
  int nst;                   // operand stages actually used (<= NST; the table takes the place of the others)
  // fwd, optional: state after each row's last valid step, hlast[row][:] = h_{lens[row]-1} (fp32, rows < nrows)
  const int* lens; float* hlast; int nrows;
  int tpp;                   // row tiles per pair actually used (1 or NTILES)
  // packed sequences (batch sorted by length, descending): tile j only runs its first tile_T[j] time steps (forward: steps
  // [0, tile_T[j]); BPTT: t from tile_T[j]-1 down to 0).  With `mirror` pair-row y owns tiles y and pair_tiles-1-y, so a
  // long tile shares its SMs with a short one.
  const int* tile_T; int mirror;
};

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
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, uint16_t mask,
                                               int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6}], [%2], %3;\n" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 2-CTA TMA loads: data lands in the executing CTA's smem, the transaction bytes are reported to an mbarrier that
// may live in the peer CTA (the pair leader) -- address given in the shared::cluster window.
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// 2-CTA multicast load: the box is written at the same smem offset of every CTA in `mask`; the transaction bytes are
// reported, for each destination, to the mbarrier at `bar` offset in that destination's PAIR LEADER (peer bit cleared).
__device__ __forceinline__ void tma_load_3d_2sm_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, uint16_t mask,
                                                   int c0, int c1, int c2) {
  const uint32_t bar_addr = ptx::smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6}], [%2], %3;\n" ::"r"(ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// asynchronous remote store: 4 bytes into the partner CTA's shared memory, reported (complete_tx) to an mbarrier of that CTA
__device__ __forceinline__ void st_async_b32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(v),
               "r"(cluster_bar)
               : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar)
               : "memory");
}
// relaxed remote arrive that cannot be issued before `dep` has been computed (orders the loads feeding `dep` in front of it)
__device__ __forceinline__ void remote_arrive_relaxed_dep(uint32_t cluster_addr, uint32_t dep) {
  asm volatile("{\n\t.reg .b32 t;\n\tand.b32 t, %1, 0;\n\tadd.u32 t, t, %0;\n\t"
               "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [t];\n\t}" ::"r"(cluster_addr), "r"(dep)
               : "memory");
}
// wait on a LOCAL mbarrier whose arrivals come from another CTA of the cluster (release.cluster): acquire at cluster scope
__device__ __forceinline__ bool wait_bar_cluster(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((++spins & 0x3FF) == 0) {
      const unsigned long long now = gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicExch(err_flag, 4); return false; }
      if (*(volatile int*)err_flag) return false;
    }
  }
}
__device__ __forceinline__ void remote_arrive_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(ptx::smem_u32(holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          ptx::smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// IN (forward only): where the input projection of a step comes from -- 0: the gi tensor (GEMM output); 1: gi (time-
// invariant per-row part) + a token table in shared memory; 2: the token table only, plus the final-state capture (lens / hlast)
// VL: packed sequences -- per-tile step windows (tile_T) and the mirrored tile assignment; without it every tile runs all T
// steps and the window arithmetic folds away
template <bool BWD, bool FAST, int CL, bool KS, int IN = 0, bool VL = false>
__global__ void __launch_bounds__(THREADS, 1)
gru_rec2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA,
                const __grid_constant__ CUtensorMap tmS, const Params2 p) {
  static_assert(!KS || (BWD && CL == 2), "K-split is a BPTT variant of the pair kernel");
  constexpr int NST = STAGES;                  // upper bound of the operand stages (barrier arrays); p.nst are in use
  constexpr int XBUF = 128 * RU * 4;           // KS: one tile's incoming partial sums (128 rows x 64 units fp32)
  constexpr int NB = BWD ? (KS ? 2 * RU : RU) : 3 * RU;        // MMA N of the pair
  constexpr int NBH = NB / 2;                  // resident rows per CTA
  constexpr int CHUNK = NBH * 128;             // bytes of one 64-wide K chunk of the resident half
  constexpr int MASTER0 = NTILES * NB;
  constexpr int TMEM_COLS = (BWD && !KS) ? 256 : 512;
  constexpr int NSAME = CL / 2;                // same-parity CTAs (= pairs) per cluster
  constexpr int A_PART = A_STAGE / NSAME;      // bytes of a stage this CTA loads and multicasts
  constexpr int A_PART_ROWS = 128 / NSAME;
  static_assert(NTILES * (NB + RU) <= TMEM_COLS, "TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = (BWD ? 3 * p.Hp : p.Hp) / 64 / (KS ? 2 : 1);   // k chunks this pair contracts over
  const int KCS = (p.debug & 256) ? KC / 2 : KC;   // timing experiment: stream only half of K (wrong results)
  uint8_t* sW = smem;
  uint8_t* sA = smem + (size_t)KC * CHUNK;
  const int nst = p.nst;                                        // stages in use (NST unless a token table needs the room)
  float* xbuf = reinterpret_cast<float*>(sA + nst * A_STAGE);   // KS: [64 units][128 rows], reused by consecutive tile-steps
  __nv_bfloat16* sTbl = reinterpret_cast<__nv_bfloat16*>(sA + nst * A_STAGE + (KS ? XBUF : 0));   // fwd: [V][3][64]
  const int tbl_bytes = (!BWD && IN > 0) ? ((p.V * 3 * RU * 2 + 1023) & ~1023) : 0;
  constexpr bool SVB = !BWD && MVAE_SV_BULK;                     // saved gates leave through cp.async.bulk
  constexpr int STG_BYTES = SVB ? EPI_WARPS * SV_STAGE_BYTES : 0;
  // row-major outputs through TMA tensor stores: BPTT only (measured: BPTT 17.91 -> 17.55 us/step, forward 10.22 -> 10.62,
  // where the single h' block per warp does not pay for the completion wait)
  constexpr bool OTMA = BWD && MVAE_OUT_TMA != 0;
  constexpr bool XST = KS && OTMA && MVAE_XCHG_ASYNC != 0 && MVAE_XBUF_STAGE != 0;   // dgh staging inside the exchange slots
  constexpr int OUT_BYTES = OTMA ? EPI_WARPS * (BWD ? (XST ? 1 : 3) : 1) * 1024 : 0;
  uint8_t* sStage = sA + nst * A_STAGE + (KS ? XBUF : 0) + tbl_bytes;
  uint8_t* sOut = sStage + STG_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + OUT_BYTES);
  uint64_t* full_bar = bars;                         // [NST] own operand stage landed (tx)
  uint64_t* empty_bar = bars + NST;                  // [NST] all pairs that share the stage consumed it
  uint64_t* pfull_bar = bars + 2 * NST;              // [NST] leader: peer's stage landed (relayed)
  uint64_t* tfull_bar = bars + 3 * NST;              // [NTILES] accumulator ready (both CTAs)
  uint64_t* tempty_bar = tfull_bar + NTILES;         // [NTILES] leader: both CTAs drained the accumulator
  uint64_t* wfull_bar = tempty_bar + NTILES;         // resident weights landed
  uint64_t* pwfull_bar = wfull_bar + 1;              // leader: peer's weights landed
  uint64_t* epi_bar = pwfull_bar + 1;                // [NTILES] all epilogue threads finished the tile
  // KS exchange barriers, one per epilogue warp: warp w of this CTA trades only with warp w of the partner CTA (same rows,
  // same unit group), so no CTA-wide rendezvous sits in the middle of the epilogue
  uint64_t* xfull_bar = epi_bar + NTILES;            // [EPI_WARPS] the partner warp's partial sums landed in our xbuf
  uint64_t* xfree_bar = xfull_bar + EPI_WARPS;       // [EPI_WARPS] the partner warp consumed what we last wrote into ITS xbuf
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(xfree_bar + EPI_WARPS);
  float* sBias = reinterpret_cast<float*>(tmem_holder + 2);  // fwd: b_hn for the pair's 64 units

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int parity = rank & 1;
  const bool leader = parity == 0;
  const int pair = blockIdx.x >> 1;                   // global pair index along the hidden dimension
  const int u0 = pair * RU;                           // first of the 64 units this pair finalises
  const int kh = KS ? (pair & 1) : 0;                 // KS: which half of K this pair contracts over
  const int koff = KS ? kh * (3 * p.Hp / 2) : 0;      // first dgh column (= first W_hh^T column) of that half
  const int U0 = KS ? (pair >> 1) * 2 * RU : u0;      // KS: first of the 128 units of the two-pair cluster
  const bool mirror = VL && p.mirror;
  const int ntiles = mirror ? NTILES : min(p.tpp, p.pair_tiles - (int)blockIdx.y * p.tpp);
  // per slot (scalars, selected with i ? b : a, so that nothing is indexed dynamically): row tile, first global step at
  // which it is active, one past its last active step
  static_assert(NTILES == 2, "slot scalars");
  const int tid_a = mirror ? (int)blockIdx.y : (int)blockIdx.y * p.tpp;
  const int tid_b = mirror ? p.pair_tiles - 1 - (int)blockIdx.y : (int)blockIdx.y * p.tpp + 1;
  const int Tn_a = (VL && p.tile_T) ? min(p.T, p.tile_T[tid_a]) : p.T;
  const int Tn_b = (VL && p.tile_T && ntiles > 1) ? min(p.T, p.tile_T[tid_b]) : p.T;
  const int fs_a = (VL && BWD) ? p.T - Tn_a : 0, fs_b = (VL && BWD) ? p.T - Tn_b : 0;
  const int es_a = (VL && !BWD) ? Tn_a : p.T, es_b = (VL && !BWD) ? Tn_b : p.T;
#define TILE_ID(i) ((i) ? tid_b : tid_a)
#define F_STEP(i) ((i) ? fs_b : fs_a)
#define E_STEP(i) ((i) ? es_b : es_a)
  const uint16_t mask_par = (uint16_t)((CL == 8 ? 0x55 : CL == 4 ? 0x5 : 0x1) << parity);  // same-parity CTAs
  const uint16_t mask_pair = (uint16_t)(3u << (rank & ~1u));
  const int qd = CL == 2 ? 0 : (int)(rank >> 1);      // multicast clusters: which part of the operand tile this CTA loads

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&tmW);
    ptx::tma_prefetch_desc(&tmA);
    if (OTMA) ptx::tma_prefetch_desc(&tmS);
    for (int s = 0; s < NST; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], NSAME);
      ptx::mbar_init(&pfull_bar[s], 1);
    }
    for (int i = 0; i < NTILES; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 2);
      ptx::mbar_init(&epi_bar[i], EPI_WARPS * 32);
    }
    for (int i = 0; i < EPI_WARPS; ++i) {
      ptx::mbar_init(&xfull_bar[i], 1);
      ptx::mbar_init(&xfree_bar[i], 1);
    }
    ptx::mbar_init(wfull_bar, 1);
    ptx::mbar_init(pwfull_bar, 1);
    ptx::fence_mbar_init();
  }
  if (!BWD) for (int i = threadIdx.x; i < RU; i += THREADS) sBias[i] = p.bhn[u0 + i];
  if (!BWD && IN > 0)
    for (int i = threadIdx.x; i < p.V * 3 * RU; i += THREADS) {
      const int v = i / (3 * RU), g = (i / RU) % 3, u = i % RU;
      sTbl[i] = __float2bfloat16_rn(p.tbl[(size_t)v * 3 * p.Hp + g * p.Hp + u0 + u]);
    }
  if (warp == 1) tmem_alloc2(tmem_holder, TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < CTRL_WARPS) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(CTRL_REGS));
  if (warp == 0) {
    // ===================== TMA producer (every CTA) =====================
    if (lane == 0) {
      constexpr bool DIRECT = true;        // both CTAs' TMA report straight to the pair leader's barriers (no relay)
      const uint32_t lrank = rank & ~1u;
      const uint32_t wbar_addr = DIRECT ? mapa(ptx::smem_u32(wfull_bar), lrank) : ptx::smem_u32(wfull_bar);
      if (DIRECT) { if (leader) ptx::mbar_arrive_expect_tx(wfull_bar, (uint32_t)(2 * KC * CHUNK)); }
      else ptx::mbar_arrive_expect_tx(wfull_bar, (uint32_t)(KC * CHUNK));
      for (int kc = 0; kc < KC; ++kc) {
        uint8_t* dst = sW + (size_t)kc * CHUNK;
        if (BWD && KS) {
#pragma unroll
          for (int b = 0; b < 2; ++b)                                       // 64 rows (units) of W_hh^T, this pair's K half
            tma_load_2d_2sm(dst + b * 4096, &tmW, wbar_addr, koff + kc * 64, U0 + 64 * parity + 32 * b);
        } else if (BWD) {
          if (DIRECT) tma_load_2d_2sm(dst, &tmW, wbar_addr, kc * 64, u0 + 32 * parity);
          else tma_load_2d(dst, &tmW, wfull_bar, kc * 64, u0 + 32 * parity);     // 32 rows of W_hh^T
        } else {
#pragma unroll
          for (int b = 0; b < 3; ++b) {                                     // 3 boxes of 32 gate rows
            const int n = 96 * parity + 32 * b;                             // row inside the pair's [r|z|n] x 64 block
            if (DIRECT) tma_load_2d_2sm(dst + b * 4096, &tmW, wbar_addr, kc * 64, (n >> 6) * p.Hp + u0 + (n & 63));
            else tma_load_2d(dst + b * 4096, &tmW, wfull_bar, kc * 64, (n >> 6) * p.Hp + u0 + (n & 63));
          }
        }
      }
      int s = 0; uint32_t ph = 0;
      // forward: step 0 contracts over slab 0 (the initial state; zeros when there is none) like every other step
      for (int step = BWD ? 1 : 0; step < p.T; ++step) {
        int slab = BWD ? (p.T - step) : step;
        if (p.debug & 2) slab = (slab + pair * 5) % p.T;
        if (p.debug & 16) slab = 0;   // timing experiment: operand that nobody writes during the sweep
        for (int i = 0; i < ntiles; ++i) {
          const int tile = TILE_ID(i);
          const int ls = step - F_STEP(i);                       // steps this tile has completed so far
          if (step >= E_STEP(i) || ls < (BWD ? 1 : 0)) continue;   // BPTT: the tile's first step has no operand
          if (!(p.debug & 1) && !wait_counter(p.counters + tile * 2 + parity, (unsigned)(p.npairs * ls), p.err_flag)) goto done;
          ptx::fence_proxy_async_all();
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 12 + 0] = gtime();
          const int row0 = tile * 256 + parity * 128 + qd * A_PART_ROWS;
          for (int kc0 = 0; kc0 < KCS; ++kc0) {
            const int kc = (p.debug & 128) ? (kc0 + pair * (KC / 8)) % KC : kc0;   // experiment: de-phase the pairs' chunk order
            if (!wait_bar(&empty_bar[s], ph ^ 1, p.err_flag)) goto done;
            if (p.debug & 4) {
              if (leader) ptx::mbar_arrive(&full_bar[s]);   // timing experiment: no operand traffic
            } else if (DIRECT) {
              if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * A_STAGE);
              if (CL == 2)
              {
                const uint32_t fb = mapa(ptx::smem_u32(&full_bar[s]), lrank);
                for (int r0 = 0; r0 < 128; r0 += p.a_box_rows)
                  tma_load_3d_2sm(sA + s * A_STAGE + r0 * 128, &tmA, fb, koff + kc * 64, row0 + r0, slab);
              }
              else
                tma_load_3d_2sm_mc(sA + s * A_STAGE + qd * A_PART, &tmA, &full_bar[s], mask_par, kc * 64, row0, slab);
            } else {
              ptx::mbar_arrive_expect_tx(&full_bar[s], A_STAGE);
              tma_load_3d_mc(sA + s * A_STAGE + qd * A_PART, &tmA, &full_bar[s], mask_par, kc * 64, row0, slab);
            }
            if (++s == nst) { s = 0; ph ^= 1; }
          }
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 12 + 1] = gtime();
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if (!leader) {
        // no relay needed
      } else if (false) {
        // pair-only clusters need no relay: the peer's TMA reports directly to the leader's barriers
      } else if (!leader) {
        // ===================== relay (odd CTA): forward "my stage landed" to the pair leader =====================
        const uint32_t lrank = rank & ~1u;
        if (!wait_bar(wfull_bar, 0, p.err_flag)) goto done;
        remote_arrive_relaxed(mapa(ptx::smem_u32(pwfull_bar), lrank));
        int s = 0; uint32_t ph = 0;
        for (int step = 1; step < p.T; ++step)
          for (int i = 0; i < ntiles; ++i)
            for (int kc = 0; kc < KC; ++kc) {
              if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
              remote_arrive_relaxed(mapa(ptx::smem_u32(&pfull_bar[s]), lrank));
              if (++s == nst) { s = 0; ph ^= 1; }
            }
      } else {
        // ===================== MMA issuer (even CTA, one thread for the pair) =====================
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(256, NB, 0, 0);
        if (!wait_bar(wfull_bar, 0, p.err_flag)) goto done;

        int s = 0; uint32_t ph = 0;
        for (int step = 0; step < p.T; ++step) {
          for (int i = 0; i < ntiles; ++i) {
            const int ls = step - F_STEP(i);
            if (ls < 0 || step >= E_STEP(i)) continue;
            if (ls > 0 || !BWD) {
              if (ls > 0 && !wait_bar(&tempty_bar[i], (uint32_t)((ls - 1) & 1), p.err_flag)) goto done;
              ptx::tc_fence_after();
              const uint32_t d_tmem = tmem_base + i * NB;
              for (int kc0 = 0; kc0 < KCS; ++kc0) {
                const int kc = (p.debug & 128) ? (kc0 + pair * (KC / 8)) % KC : kc0;
                const bool trm = p.trace && blockIdx.x == 0 && blockIdx.y == 0;
                if (!wait_bar(&full_bar[s], ph, p.err_flag)) goto done;
                if (trm && kc == 0) p.trace[((size_t)step * NTILES + i) * 12 + 2] = gtime();
                if (trm && kc == KC - 1) p.trace[((size_t)step * NTILES + i) * 12 + 9] = gtime();
                if (trm && kc == 0) p.trace[((size_t)step * NTILES + i) * 12 + 8] = gtime();
                if (trm && kc == KC - 1) p.trace[((size_t)step * NTILES + i) * 12 + 10] = gtime();
                ptx::tc_fence_after();
                const uint32_t a_addr = ptx::smem_u32(sA + s * A_STAGE);
                const uint32_t b_addr = ptx::smem_u32(sW + (size_t)kc * CHUNK);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t adesc = ptx::umma_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                  const uint64_t bdesc = ptx::umma_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                  umma2_bf16(d_tmem, adesc, bdesc, idesc, (kc0 > 0 || k > 0) ? 1u : 0u);
                }
                // stage s is free (for this pair) in every CTA that shares it: the pair itself, or the whole multicast cluster
                commit2_mc(&empty_bar[s], CL == 2 ? mask_pair : (uint16_t)((1u << CL) - 1));
                if (++s == nst) { s = 0; ph ^= 1; }
              }
            }
            const unsigned long long t_issue = (p.debug & 8) ? gtime() : 0ull;
            commit2_mc(&tfull_bar[i], mask_pair);     // accumulator i complete -> both epilogues
            if ((p.debug & 8) && p.trace && blockIdx.x == 0 && blockIdx.y == 0) {
              // timing experiment: how long until the tensor pipe has really finished this tile's MMAs
              wait_bar(&tfull_bar[i], (uint32_t)(ls & 1), p.err_flag);
              p.trace[((size_t)step * NTILES + i) * 12 + 9] = t_issue;
              p.trace[((size_t)step * NTILES + i) * 12 + 10] = gtime();
            }
            if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[((size_t)step * NTILES + i) * 12 + 3] = gtime();
          }
        }
      }
    }
  } else if (warp == PUB_WARP) {
    // ===================== publisher: frees the accumulator and publishes the tile's step to the other pairs =====================
    if (lane == 0) {
      const uint32_t tempty_remote = mapa(ptx::smem_u32(&tempty_bar[0]), rank & ~1u);
      for (int step = 0; step < p.T; ++step)
        for (int i = 0; i < ntiles; ++i) {
          const int tile = TILE_ID(i);
          const int ls = step - F_STEP(i);
          if (ls < 0 || step >= E_STEP(i)) continue;
          if (!wait_bar(&epi_bar[i], (uint32_t)(ls & 1), p.err_flag)) goto done;
          const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0;
          if (tr) p.trace[((size_t)step * NTILES + i) * 12 + 6] = gtime();
          // The release of the red is cumulative over the epilogue threads' stores this thread observed through the
          // epi_bar acquire; the consumer issues the generic->async proxy fence after its acquire.  (A separate
          // __threadfence + writer-side proxy fence here cost 0.9 us per step; debug&512 restores them.)
          if (p.debug & 512) {
            __threadfence();
            ptx::fence_proxy_async_all();
          }
          red_release_add(p.counters + tile * 2 + parity, 1u);
          if (tr) p.trace[((size_t)step * NTILES + i) * 12 + 7] = gtime();
          // accumulator i drained in this CTA -> pair leader.  After the publication: the other pairs' next step waits for the
          // counter, while the leader's MMA into this accumulator is a whole tile-stream away (the release fence of this arrive
          // cost the chain ~0.8 us per tile-step when it came first)
          remote_arrive(tempty_remote + (uint32_t)(i * 8));
        }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(EPI_REGS));
    // ===================== epilogue warps (every CTA: own 128 rows x 64 units) =====================
    const int q = warp & 3;
    const int part = (warp - CTRL_WARPS) >> 2;   // 0..3 -> 16 units each
    const int uc = part * 16;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t nx = 0;   // KS: exchanges this warp has done so far (the partner warp runs the same tile / step sequence)
    uint32_t nts = 0;  // XST: tile-steps this warp has finished (every one of them ends with an xfree arrive)
    int lastv_a = -1, lastv_b = -1;   // IN == 2: step after which this thread's row of slot a / b hands its state to hlast
    if constexpr (!BWD && (IN == 2 || (IN == 0 && VL))) {   // IN 0 + packed sequences: embedding layer with a materialised projection
      if (p.hlast)
        for (int i = 0; i < ntiles; ++i) {
          const long long row = (long long)TILE_ID(i) * 256 + parity * 128 + q * 32 + lane;
          if (row < p.nrows) { if (i) lastv_b = p.lens[row] - 1; else lastv_a = p.lens[row] - 1; }
        }
    }
    if constexpr (!BWD) {
      // fp32 master copy of the initial state (zeros without h0) into this thread's TMEM columns of every tile
      for (int i = 0; i < ntiles; ++i) {
        const long long row = (long long)TILE_ID(i) * 256 + parity * 128 + q * 32 + lane;
        float h[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) h[k] = 0.f;
        if (p.h0) {
          const float4* h0p = reinterpret_cast<const float4*>(p.h0 + row * p.Hp + u0 + uc);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 v = __ldg(h0p + k);
            h[4 * k] = v.x; h[4 * k + 1] = v.y; h[4 * k + 2] = v.z; h[4 * k + 3] = v.w;
          }
        }
        tmem_st_32x16(tmem_base + lane_off + MASTER0 + i * RU + uc, h);
      }
      tmem_st_wait();
    }
    for (int step = 0; step < p.T; ++step) {
      const int t = BWD ? (p.T - 1 - step) : step;
      for (int i = 0; i < ntiles; ++i) {
        const int tile = TILE_ID(i);
        const int ls = step - F_STEP(i);
        if (ls < 0 || step >= E_STEP(i)) continue;
        const long long row = (long long)tile * 256 + parity * 128 + q * 32 + lane;
        // saved activations use a per-thread "fragment" layout (consumed only by the BPTT epilogue with the same
        // thread mapping): block (t, tile, pair, parity, part, array) of 4 KB, thread (q, lane) owns 32 B -> every
        // warp-wide 256-bit access is 1 KB contiguous.
        const size_t sv_blk = (((((size_t)t * p.pair_tiles + tile) * p.npairs + pair) * 2 + parity) * 4 + part) * 4 + q;
        // this warp's block: SV_NARR arrays of [half][lane][8 units] -> a thread's two 16-byte halves are 512 B apart
        __nv_bfloat16* svw = p.sv ? p.sv + sv_blk * (SV_NARR * 512) : nullptr;
        __nv_bfloat16* svp = svw ? svw + lane * 8 : nullptr;
        // gi / dX are produced by the GEMM epilogues in the row-blocked layout [row/32][width/8][32][8]:
        // this thread's 16 units of its row are two 16-byte pieces 512 B apart; a warp access is 512 B contiguous.
        const long long rblk = row >> 5;   // = tile*8 + parity*4 + q ; row & 31 == lane
        u32x8 pre[BWD ? 6 : 3];
        int tokv = 0;
        const bool nomem = (p.debug & 32) != 0;   // timing experiment: epilogue without global traffic
        if (nomem) {
#pragma unroll
          for (int a = 0; a < (BWD ? 6 : 3); ++a)
#pragma unroll
            for (int k = 0; k < 8; ++k) pre[a].v[k] = 0x3c003c00u;
        } else if (!BWD) {
          if constexpr (IN != 2) {
            const __nv_bfloat16* g = p.gi + (long long)t * p.gi_tstride +
                                     (rblk * (3 * p.Hp / 8) + ((u0 + uc) >> 3)) * 256 + lane * 8;
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) pre[gate] = ldg2x128(g + (long long)gate * (p.Hp / 8) * 256);
          }
          if constexpr (IN > 0) tokv = p.tok[(long long)t * p.Bp + row];
        } else {
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) pre[blk] = ldg2x128(svp + blk * 512);
          if (MVAE_SV_HP) pre[4] = ldg2x128(svp + 4 * 512);   // h_{t-1} saved by the forward sweep next to the gates
          else pre[4] = ldg256(p.hs + ((long long)t * p.Bp + row) * p.Hp + u0 + uc);   // h_{t-1} = slab t of the hidden states
          pre[5] = ldg2x128(p.dX + (long long)t * p.Bp * p.Hp + (rblk * (p.Hp / 8) + ((u0 + uc) >> 3)) * 256 + lane * 8);
        }
        (void)wait_bar(&tfull_bar[i], (uint32_t)(ls & 1), p.err_flag);   // on failure keep walking: barriers below must be reached
        ptx::tc_fence_after();
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && warp == CTRL_WARPS && lane == 0;
        if (tr) p.trace[((size_t)step * NTILES + i) * 12 + 4] = gtime();
        const uint32_t master_addr = tmem_base + lane_off + MASTER0 + i * RU + uc;
        if (!BWD) {
          uint32_t ar[16], az[16], an[16], hm[16];
          {
            const uint32_t acc = tmem_base + lane_off + i * NB + uc;
            ptx::tmem_ld_32x16(acc, ar);
            ptx::tmem_ld_32x16(acc + RU, az);
            ptx::tmem_ld_32x16(acc + 2 * RU, an);
            ptx::tmem_ld_32x16(master_addr, hm);
            ptx::tmem_ld_wait();
          }
          const bool tr2 = tr && (p.debug & 64);
          if (tr2) p.trace[((size_t)step * NTILES + i) * 12 + 8] = gtime();
          float gr[16], gz[16], gn[16];
          if constexpr (IN != 2) {
            unpack16(pre[0], gr);
            unpack16(pre[1], gz);
            unpack16(pre[2], gn);
          }
          if constexpr (IN > 0) {
            // token-table part of the input projection: this row's token selects one bf16 row per gate in shared memory
            const uint4* tr = reinterpret_cast<const uint4*>(sTbl + (size_t)tokv * 3 * RU + uc);
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
              float* dst = gate == 0 ? gr : gate == 1 ? gz : gn;
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const uint4 a = tr[gate * (RU / 8) + hf];
                const uint32_t w4[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float lo = __uint_as_float(w4[k] << 16), hi = __uint_as_float(w4[k] & 0xFFFF0000u);
                  if constexpr (IN == 2) { dst[hf * 8 + 2 * k] = lo; dst[hf * 8 + 2 * k + 1] = hi; }
                  else { dst[hf * 8 + 2 * k] += lo; dst[hf * 8 + 2 * k + 1] += hi; }
                }
              }
            }
          }
          float h[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float ghn = __uint_as_float(an[k]) + sBias[uc + k];
            const float r = gate_sigmoid_t<FAST>(gr[k] + __uint_as_float(ar[k]));   // b_hr folded into gi
            const float z = gate_sigmoid_t<FAST>(gz[k] + __uint_as_float(az[k]));   // b_hz folded into gi
            const float n = gate_tanh_t<FAST>(fmaf(r, ghn, gn[k]));
            h[k] = fmaf(z, __uint_as_float(hm[k]) - n, n);
            // "ones column": the last pad unit carries the constant 1 so that the K = T*B weight-gradient GEMMs
            // also produce the bias gradients (its weights are zero padding, so it never feeds the recurrence)
            if (p.ones_col >= 0 && u0 + uc + k == p.ones_col) h[k] = 1.0f;
            gr[k] = r; gz[k] = z; gn[k] = n;
            an[k] = __float_as_uint(ghn);
          }
          if (tr2) p.trace[((size_t)step * NTILES + i) * 12 + 9] = gtime();
          if ((IN == 2 || (IN == 0 && VL)) && t == (i ? lastv_b : lastv_a)) {
            float4* hl = reinterpret_cast<float4*>(p.hlast + row * p.Hp + u0 + uc);
#pragma unroll
            for (int k = 0; k < 4; ++k) hl[k] = make_float4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
          }
          tmem_st_32x16(master_addr, h);
          if (!nomem) {
            if constexpr (OTMA) {
              uint8_t* ob = sOut + (warp - CTRL_WARPS) * 1024;
              sts_row32(ob, lane, h);          // (the previous tile-step's store of this buffer completed before its arrive)
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&tmS, ob, u0 + uc, (int)(row - lane), t + 1);
                bulk_commit();
                bulk_wait_all0();              // h' is in memory: the tile can be published
              }
              __syncwarp();
            } else {
              stg256(p.hs + ((long long)(t + 1) * p.Bp + row) * p.Hp + u0 + uc, h);
            }
          }
          // everything other CTAs / the MMA wait for is issued: let the publisher go before the saved-gate stores
          tmem_st_wait();
          ptx::tc_fence_before();
          ptx::mbar_arrive(&epi_bar[i]);
          if (tr2) p.trace[((size_t)step * NTILES + i) * 12 + 10] = gtime();
          if (svp && !nomem && !(p.debug & 1024)) {   // debug 1024: timing experiment without the saved-gate stores
            if constexpr (SVB) {
              // stage two arrays (2 KB per warp), hand them to the bulk-copy engine, re-use the staging buffer once the
              // engine has READ it; the global writes themselves complete asynchronously, off the LSU path
              // (lane 1 issues them: bulk groups are per thread, and lane 0's wait for the h' store must not wait for these)
              uint8_t* stg = sStage + (warp - CTRL_WARPS) * SV_STAGE_BYTES;
#if MVAE_SV_STAGE2 == 2
              // two staging buffers, three bulk stores: [r | z] from A, [n | W_hn h] from B, h_{t-1} from A again -- the only
              // wait inside a tile-step is for the oldest store (A), issued two stagings earlier
              if (lane == 1) bulk_wait_read0();
              __syncwarp();
              sts2x128(stg, lane, gr);
              sts2x128(stg + 1024, lane, gz);
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 1) { bulk_s2g(svw, stg, 2048); bulk_commit(); }
#pragma unroll
              for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(an[k]);
              sts2x128(stg + 2048, lane, gn);
              sts2x128(stg + 3072, lane, h);
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 1) { bulk_s2g(svw + 2 * 512, stg + 2048, 2048); bulk_commit(); }
              if (MVAE_SV_HP) {
                if (lane == 1) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");   // A's store has read its source
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(hm[k]);
                sts2x128(stg, lane, h);
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 1) { bulk_s2g(svw + 4 * 512, stg, 1024); bulk_commit(); }
              }
#elif MVAE_SV_STAGE2
              if (lane == 1) bulk_wait_read0();          // the previous tile-step's two stores have read their buffers
              __syncwarp();
              sts2x128(stg, lane, gr);
              sts2x128(stg + 1024, lane, gz);
#pragma unroll
              for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(an[k]);
              sts2x128(stg + 2048, lane, gn);
              sts2x128(stg + 3072, lane, h);
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 1) { bulk_s2g(svw, stg, 4096); bulk_commit(); }
              if (MVAE_SV_HP) {
#pragma unroll
                for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(hm[k]);
                stg2x128(svp + 4 * 512, h);               // h_{t-1}: 1 KB per warp straight through the LSU
              }
#else
              if (lane == 1) bulk_wait_read0();
              __syncwarp();
              sts2x128(stg, lane, gr);
              sts2x128(stg + 1024, lane, gz);
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 1) { bulk_s2g(svw, stg, 2048); bulk_commit(); bulk_wait_read0(); }
              __syncwarp();
#pragma unroll
              for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(an[k]);
              sts2x128(stg, lane, gn);
              sts2x128(stg + 1024, lane, h);
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 1) { bulk_s2g(svw + 2 * 512, stg, 2048); bulk_commit(); }
              if (MVAE_SV_HP) {
                if (lane == 1) bulk_wait_read0();
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(hm[k]);
                sts2x128(stg, lane, h);
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 1) { bulk_s2g(svw + 4 * 512, stg, 1024); bulk_commit(); }
              }
#endif
            } else {
              stg2x128(svp, gr);
              stg2x128(svp + 512, gz);
              stg2x128(svp + 2 * 512, gn);
#pragma unroll
              for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(an[k]);
              stg2x128(svp + 3 * 512, h);
              if (MVAE_SV_HP) {
#pragma unroll
                for (int k = 0; k < 16; ++k) h[k] = __uint_as_float(hm[k]);
                stg2x128(svp + 4 * 512, h);   // h_{t-1}: lets the BPTT epilogue skip the row-major hs read
              }
            }
          }
        } else {
          uint32_t acc[16], cm[16];
          uint32_t xdep = 0u;   // XST: what the xfree arrive of this tile-step is made data-dependent on
          if (ls > 0) {
            if (KS) {
              // this pair contracted over ONE half of K for all 128 units of the cluster: hand the partial sums of the 64
              // units the partner pair finalises to the partner CTA (same parity -> same rows) through its shared memory,
              // and add the partner's partial sums for our own 64 units.
              uint32_t oth[16];
              ptx::tmem_ld_32x16(tmem_base + lane_off + i * NB + RU * (1 - kh) + uc, oth);
              ptx::tmem_ld_32x16(tmem_base + lane_off + i * NB + RU * kh + uc, acc);
              ptx::tmem_ld_32x16(master_addr, cm);
              ptx::tmem_ld_wait();
              const uint32_t partner = rank ^ 2u;
              const uint32_t ev = nx++;                                  // exchange number
              // xbuf[unit][row]: a warp store covers 32 consecutive rows of one unit (128 contiguous bytes = one DSMEM
              // transaction).  (A [row][unit] layout with 128-bit stores -- 4 instructions instead of 16, but 32 separate
              // 16-byte remote transactions per instruction -- was measured slower: push 1.5 -> 2.4 us, sweep 17.55 -> 18.9 us/step.)
              const int xw = warp - CTRL_WARPS;                                          // this warp's exchange slot
              // XST: the slot is one contiguous 2 KB block per warp, [unit][lane] (it doubles as TMA-store staging below)
              const float* xin = XST ? xbuf + (size_t)xw * 512 + lane : xbuf + (size_t)uc * 128 + q * 32 + lane;
              constexpr uint32_t XSTRIDE = XST ? 128u : 512u;                            // bytes between units in a slot
              constexpr int XIN_STRIDE = XST ? 32 : 128;
              const uint32_t xloc = ptx::smem_u32(xin);
              const uint32_t xrem = mapa(xloc, partner);
              if constexpr (XST) {
                if (nts > 0) (void)wait_bar_cluster(&xfree_bar[xw], (nts - 1) & 1u, p.err_flag);   // partner's previous tile-step left its slot
              } else {
                if (ev > 0) (void)wait_bar_cluster(&xfree_bar[xw], (ev - 1) & 1u, p.err_flag);     // partner warp read exchange ev-1
              }
#if MVAE_XCHG_ASYNC
              // arm our own barrier for the 2 KB the partner warp sends (bytes that land before this only drive the tx-count
              // negative), then fire the stores: no fence, no arrive on the sending side
              if (lane == 0) ptx::mbar_arrive_expect_tx(&xfull_bar[xw], 16u * 32u * 4u);
              {
                const uint32_t xbar_rem = mapa(ptx::smem_u32(&xfull_bar[xw]), partner);
#if MVAE_XCHG_V4
                if constexpr (XST) {
                  // slot layout [4 unit groups][lane][4 units]: one 16-byte store per lane and group = 512 contiguous bytes per
                  // warp instruction (4 instead of 16 remote stores)
                  const uint32_t xr4 = mapa(ptx::smem_u32(xbuf + (size_t)xw * 512 + lane * 4), partner);
#pragma unroll
                  for (int g = 0; g < 4; ++g)
                    st_async_v4(xr4 + (uint32_t)g * 512u, oth[4 * g], oth[4 * g + 1], oth[4 * g + 2], oth[4 * g + 3], xbar_rem);
                } else
#endif
                {
#pragma unroll
                  for (int k = 0; k < 16; ++k) st_async_b32(xrem + (uint32_t)k * XSTRIDE, oth[k], xbar_rem);
                }
              }
              (void)wait_bar_cluster(&xfull_bar[xw], ev & 1u, p.err_flag);
              uint32_t seen = 0u;
#if MVAE_XCHG_V4
              if constexpr (XST) {
                const float4* x4 = reinterpret_cast<const float4*>(xbuf + (size_t)xw * 512 + lane * 4);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const float4 v = x4[g * 32];
                  seen |= __float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w);
                  acc[4 * g] = __float_as_uint(__uint_as_float(acc[4 * g]) + v.x);
                  acc[4 * g + 1] = __float_as_uint(__uint_as_float(acc[4 * g + 1]) + v.y);
                  acc[4 * g + 2] = __float_as_uint(__uint_as_float(acc[4 * g + 2]) + v.z);
                  acc[4 * g + 3] = __float_as_uint(__uint_as_float(acc[4 * g + 3]) + v.w);
                }
              } else
#endif
              {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const float xv = xin[k * XIN_STRIDE];
                  seen |= __float_as_uint(xv);
                  acc[k] = __float_as_uint(__uint_as_float(acc[k]) + xv);
                }
              }
              // every lane's reads are complete once the warp-wide OR of what they returned exists
              seen = __reduce_or_sync(0xffffffffu, seen);
              if constexpr (!XST) {
                if (lane == 0) remote_arrive_relaxed_dep(mapa(ptx::smem_u32(&xfree_bar[xw]), partner), seen);
              } else {
                xdep = seen;   // the slot goes back at the end of the tile-step, after it has served as store staging
              }
#else
#pragma unroll
              for (int k = 0; k < 16; ++k) st_cluster_f32(xrem + (uint32_t)k * 512u, oth[k]);
              __syncwarp();
              if (lane == 0) remote_arrive(mapa(ptx::smem_u32(&xfull_bar[xw]), partner));   // release.cluster
              (void)wait_bar_cluster(&xfull_bar[xw], ev & 1u, p.err_flag);
#pragma unroll
              for (int k = 0; k < 16; ++k) acc[k] = __float_as_uint(__uint_as_float(acc[k]) + xin[k * 128]);
              __syncwarp();
              if (lane == 0) remote_arrive(mapa(ptx::smem_u32(&xfree_bar[xw]), partner));
#endif
            } else {
              ptx::tmem_ld_32x16(tmem_base + lane_off + i * NB + uc, acc);
              ptx::tmem_ld_32x16(master_addr, cm);
              ptx::tmem_ld_wait();
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) { acc[k] = 0u; cm[k] = 0u; }
          }
          if (tr && (p.debug & 64)) p.trace[((size_t)step * NTILES + i) * 12 + 8] = gtime();    // exchange done
          float r[16], z[16], n[16], ghn[16], hp[16], dx[16];
          unpack16(pre[0], r);
          unpack16(pre[1], z);
          unpack16(pre[2], n);
          unpack16(pre[3], ghn);
          unpack16(pre[4], hp);
          unpack16(pre[5], dx);
          float carry[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float dh = __uint_as_float(acc[k]) + __uint_as_float(cm[k]) + dx[k];
            const float dan = dh * (1.f - z[k]) * (1.f - n[k] * n[k]);
            const float daz = dh * (hp[k] - n[k]) * z[k] * (1.f - z[k]);
            const float dar = dan * ghn[k] * r[k] * (1.f - r[k]);
            carry[k] = dh * z[k];
            hp[k] = dan * r[k];   // da_n * r
            n[k] = dan; ghn[k] = dar; dx[k] = daz;
          }
          tmem_st_32x16(master_addr, carry);
          if (p.carry_out && t == 0) {
            float4* co = reinterpret_cast<float4*>(p.carry_out + row * p.Hp + u0 + uc);
#pragma unroll
            for (int k = 0; k < 4; ++k) co[k] = make_float4(carry[4 * k], carry[4 * k + 1], carry[4 * k + 2], carry[4 * k + 3]);
          }
          if (tr && (p.debug & 64)) p.trace[((size_t)step * NTILES + i) * 12 + 9] = gtime();    // gate math done
          __nv_bfloat16* g4 = p.dG + ((long long)t * p.Bp + row) * 4 * p.Hp + u0 + uc;
          // staging blocks of the three dgh stores: own 3 KB, or (XST) this warp's exchange slot for the first two + own 1 KB
          uint8_t* ob = XST ? sOut + (warp - CTRL_WARPS) * 1024 : sOut + (warp - CTRL_WARPS) * 3072;
          uint8_t* ob_r = XST ? reinterpret_cast<uint8_t*>(xbuf + (size_t)(warp - CTRL_WARPS) * 512) : ob;
          uint8_t* ob_z = XST ? ob_r + 1024 : ob + 1024;
          uint8_t* ob_nr = XST ? ob : ob + 2048;
          if (!nomem) {
            // the three dgh blocks the other pairs stream next step go first; da_n (only read by the later wgrad / dX
            // GEMMs) is stored after this tile has been handed to the publisher
            if constexpr (OTMA) {
              if (lane == 1) bulk_wait_read0();     // the da_n store of the previous tile-step (lane 1's group) has read block 0
              __syncwarp();
              sts_row32(ob_r, lane, ghn);
              sts_row32(ob_z, lane, dx);
              sts_row32(ob_nr, lane, hp);
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                const int r0 = (int)(row - lane);
                tma_store_3d(&tmS, ob_r, p.Hp + u0 + uc, r0, t);
                tma_store_3d(&tmS, ob_z, 2 * p.Hp + u0 + uc, r0, t);
                tma_store_3d(&tmS, ob_nr, 3 * p.Hp + u0 + uc, r0, t);
                bulk_commit();
                if (tr && (p.debug & 64)) p.trace[((size_t)step * NTILES + i) * 12 + 10] = gtime();   // stores issued
                bulk_wait_all0();                   // dgh of this tile-step is in memory: the tile can be published
              }
              __syncwarp();
            } else {
              stg256(g4 + p.Hp, ghn);
              stg256(g4 + 2 * p.Hp, dx);
              stg256(g4 + 3 * p.Hp, hp);
            }
          }
          if constexpr (XST) {
            // the exchange slot is free again (its stores completed, or nothing was staged): the partner may push its next sums
            ++nts;
            if (lane == 0) remote_arrive_relaxed_dep(mapa(ptx::smem_u32(&xfree_bar[warp - CTRL_WARPS]), rank ^ 2u), xdep);
          }
          tmem_st_wait();
          ptx::tc_fence_before();
          if (tr) p.trace[((size_t)step * NTILES + i) * 12 + 5] = gtime();
          ptx::mbar_arrive(&epi_bar[i]);
          if (!nomem) {
            if constexpr (OTMA) {
              sts_row32(ob, lane, n);               // block 0 is free: its dgh store completed above
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 1) { tma_store_3d(&tmS, ob, u0 + uc, (int)(row - lane), t); bulk_commit(); }
            } else {
              stg256(g4, n);
            }
          }
        }
        if (tr && (p.debug & 64)) p.trace[((size_t)step * NTILES + i) * 12 + 11] = gtime();
        if (!BWD && tr) p.trace[((size_t)step * NTILES + i) * 12 + 5] = gtime();
      }
    }
    if constexpr (SVB || OTMA) { if (lane < 2) bulk_wait_all0(); }   // the last bulk / TMA stores must be in memory at kernel end
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while cluster peers may still signal its barriers / multicast into its smem
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
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
           const cuuint32_t* box, bool swizzle = true) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return MVAE_ERR_DRIVER;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_DRIVER;
}

template <bool BWD, bool FAST, int CL, bool KS = false, int IN = 0, bool VL = false>
int launch2(const mvae_gru_rec_args& a, cudaStream_t st) {
  const int Hp = a.Hp, Bp = a.Bp, T = a.T;
  constexpr int NBH = (BWD ? (KS ? 2 * RU : RU) : 3 * RU) / 2;
  const int KC = (BWD ? 3 * Hp : Hp) / 64 / (KS ? 2 : 1);
  const bool use_tbl = !BWD && IN > 0;
  if (use_tbl && (!a.tbl || !a.tok || a.V < 1 || a.V > 64)) return MVAE_ERR_INVALID;
  if (!BWD && IN != 2 && !a.gi) return MVAE_ERR_INVALID;
  const size_t tbl_bytes = use_tbl ? (((size_t)a.V * 3 * RU * 2 + 1023) & ~(size_t)1023) : 0;
  int nst = BWD ? MVAE_NST_BWD : MVAE_NST_FWD;
  const size_t stg_bytes = ((!BWD && MVAE_SV_BULK) ? (size_t)EPI_WARPS * SV_STAGE_BYTES : 0) +        // saved-gate staging (forward)
                           ((BWD && MVAE_OUT_TMA) ? (size_t)EPI_WARPS * ((KS && MVAE_XCHG_ASYNC && MVAE_XBUF_STAGE) ? 1 : 3) * 1024 : 0);   // output staging (BPTT)
  const size_t fixed = (size_t)KC * NBH * 128 + (KS ? 128 * RU * 4 : 0) + tbl_bytes + stg_bytes + 1024 + 1024;
  while (nst > 2 && fixed + (size_t)nst * A_STAGE > 232448) --nst;   // the token table takes the room of operand stages
  if (((a.debug >> 16) & 0xF) >= 2 && ((a.debug >> 16) & 0xF) < nst) nst = (a.debug >> 16) & 0xF;   // timing experiment: fewer stages
  const size_t smem = fixed + (size_t)nst * A_STAGE;
  if (smem > 232448) return MVAE_ERR_UNSUPPORTED;
  CUtensorMap tmW, tmA, tmS;
  {   // store map: [32 rows][16 columns] boxes of the row-major output (forward: hs slabs; BPTT: dG), dense shared memory
    cuuint32_t box[3] = {16, 32, 1};
    if (BWD) {
      cuuint64_t dims[3] = {(cuuint64_t)4 * Hp, (cuuint64_t)Bp, (cuuint64_t)T};
      cuuint64_t str[2] = {(cuuint64_t)4 * Hp * 2, (cuuint64_t)Bp * 4 * Hp * 2};
      int rc = encode(&tmS, a.dG, 3, dims, str, box, false);
      if (rc) return rc;
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)Hp, (cuuint64_t)Bp, (cuuint64_t)(T + 1)};
      cuuint64_t str[2] = {(cuuint64_t)Hp * 2, (cuuint64_t)Bp * Hp * 2};
      int rc = encode(&tmS, a.hs, 3, dims, str, box, false);
      if (rc) return rc;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(BWD ? 3 * Hp : Hp), (cuuint64_t)(BWD ? Hp : 3 * Hp)};
    cuuint64_t str[1] = {(cuuint64_t)(BWD ? 3 * Hp : Hp) * 2};
    cuuint32_t box[2] = {64, 32};
    int rc = encode(&tmW, a.W, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const int box_rows = (CL == 2) ? (a.a_box_rows > 0 ? a.a_box_rows : 128) : 128 / (CL / 2);
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    if (BWD) {
      cuuint64_t dims[3] = {(cuuint64_t)3 * Hp, (cuuint64_t)Bp, (cuuint64_t)T};
      cuuint64_t str[2] = {(cuuint64_t)4 * Hp * 2, (cuuint64_t)Bp * 4 * Hp * 2};
      if (a.debug & 8192) str[0] = 128;   // timing experiment (wrong data): every operand stage is one contiguous 16 KB region
      int rc = encode(&tmA, a.dG + Hp, 3, dims, str, box);
      if (rc) return rc;
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)Hp, (cuuint64_t)Bp, (cuuint64_t)(T + 1)};
      cuuint64_t str[2] = {(cuuint64_t)Hp * 2, (cuuint64_t)Bp * Hp * 2};
      if (a.debug & 8192) str[0] = 128;
      int rc = encode(&tmA, a.hs, 3, dims, str, box);
      if (rc) return rc;
    }
  }
  Params2 p{};
  p.Bp = Bp; p.Hp = Hp; p.T = T; p.pair_tiles = Bp / 256; p.npairs = Hp / RU;
  p.gi = a.gi; p.gi_tstride = a.gi_tstride; p.bhn = a.bhh; p.hs = a.hs; p.sv = a.sv; p.dX = a.dX; p.dG = a.dG;
  p.counters = a.counters; p.err_flag = a.err_flag; p.trace = a.trace; p.debug = a.debug; p.ones_col = a.ones_col;
  p.a_box_rows = a.a_box_rows > 0 ? a.a_box_rows : 128;
  p.h0 = BWD ? nullptr : a.h0; p.carry_out = BWD ? a.carry_out : nullptr;
  p.nst = nst;
  p.tbl = use_tbl ? a.tbl : nullptr; p.tok = use_tbl ? a.tok : nullptr; p.V = use_tbl ? a.V : 0;
  p.lens = BWD ? nullptr : a.lens; p.hlast = (BWD || !a.lens) ? nullptr : a.hlast; p.nrows = a.nrows;
  auto kern = gru_rec2_kernel<BWD, FAST, CL, KS, IN, VL>;
  static size_t attr_cache[64] = {0};
  MVAE_CUDA_CHECK(mvae_ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, attr_cache));
  MVAE_CUDA_CHECK(cudaMemsetAsync(a.counters, 0, sizeof(unsigned int) * 2 * (Bp / 256), st));
  cudaLaunchConfig_t cfg{};
  // one row tile per pair when every CTA is co-resident anyway (small hidden sizes): twice the SMs, half the chain per step
  const int tpp = ((Hp / 32) * (Bp / 256) <= 148 && !(a.debug & 4096)) ? 1 : NTILES;
  p.tpp = tpp;
  p.tile_T = VL ? a.tile_T : nullptr;
  p.mirror = (VL && a.tile_T && tpp == NTILES && (Bp / 256) % 2 == 0) ? 1 : 0;
  cfg.gridDim = dim3(Hp / 32, ceil_div(Bp / 256, tpp), 1);
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = KS ? 4 : CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  MVAE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmW, tmA, tmS, p));
  return MVAE_OK;
}

template <int CL> int max_clusters_t(int backward) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(16, 8, 1); cfg.blockDim = dim3(THREADS, 1, 1); cfg.dynamicSmemBytes = 231424;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = -1;
  if (backward) { cudaFuncSetAttribute(gru_rec2_kernel<true, false, CL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231424); cudaOccupancyMaxActiveClusters(&n, gru_rec2_kernel<true, false, CL, false>, &cfg); }
  else { cudaFuncSetAttribute(gru_rec2_kernel<false, false, CL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231424); cudaOccupancyMaxActiveClusters(&n, gru_rec2_kernel<false, false, CL, false>, &cfg); }
  return n;
}
}  // namespace

int mvae_gru_rec2_max_clusters(int backward, int cluster) {
  return cluster == 8 ? max_clusters_t<8>(backward) : cluster == 4 ? max_clusters_t<4>(backward) : max_clusters_t<2>(backward);
}

// variant 3.  Requires Bp % 256 == 0 and Hp in {256, 512}.  a->variant 32: K-split BPTT sweep (clusters of two pairs; the
// forward sweep of that variant is the plain pair kernel).  a->bhh must point at the n-gate slice of the padded
// b_hh (b_hr / b_hz are expected to be folded into gi by the caller).  a->variant: 3 -> clusters of 2 (pair only),
// 34 -> clusters of 4, 38 -> clusters of 8 (operand multicast across the pairs of a cluster).
int mvae_gru_rec2_launch(const mvae_gru_rec_args* a, int fast_gates, cudaStream_t stream) {
  if (!a || a->Bp % 256 || (a->Hp != 256 && a->Hp != 512) || a->T < 1) return MVAE_ERR_INVALID;
  const int cl = a->variant == 38 ? 8 : a->variant == 34 ? 4 : 2;
  if (a->tile_T) {   // packed sequences: K-split BPTT / pair forward kernel with per-tile step windows (fast gates only)
    if (a->backward) return a->variant == 32 ? launch2<true, false, 2, true, 0, true>(*a, stream) : MVAE_ERR_UNSUPPORTED;
    if (cl != 2 || !fast_gates) return MVAE_ERR_UNSUPPORTED;
    if (!a->tbl) return launch2<false, true, 2, false, 0, true>(*a, stream);
    return a->gi ? launch2<false, true, 2, false, 1, true>(*a, stream) : launch2<false, true, 2, false, 2, true>(*a, stream);
  }
  if (a->variant == 32 && a->backward) return launch2<true, false, 2, true>(*a, stream);
  if (!a->backward && a->tbl) {   // token-table input projection (pair clusters, fast gates only)
    if (cl != 2) return MVAE_ERR_UNSUPPORTED;
    return a->gi ? launch2<false, true, 2, false, 1>(*a, stream) : launch2<false, true, 2, false, 2>(*a, stream);
  }
#define MVAE_DISPATCH(CLV)                                                                     \
  if (a->backward) return launch2<true, false, CLV>(*a, stream);                                \
  return fast_gates ? launch2<false, true, CLV>(*a, stream) : launch2<false, false, CLV>(*a, stream);
  if (cl == 8) { MVAE_DISPATCH(8) }
  if (cl == 4) { MVAE_DISPATCH(4) }
  MVAE_DISPATCH(2)
#undef MVAE_DISPATCH
}
