// Device-side text assembly for decoded / sampled id rows: replaces the per-row Python loops of
// mosesvae.py:258-262 (slice + tensor2string -> vocab.ids2string, vocab.py:62-73), hugesample.py:31-35 ("[sym]" join)
// and featurizer.py:26-37 (charset join + strip) by one kernel sequence and ONE device->host copy of a packed byte
// buffer + row offsets.
#include "../../include/mvae_b200.h"
#include "common.cuh"
#include "simt_kernels.cuh"
#include "umma_gemm.h"
#include "host_common.cuh"

namespace {

struct RowSpan { int first, last; };   // token positions [first, last] that are emitted

// positions to emit for row b: drop a leading `rem_first` id / trailing `rem_last` id (ids2string's rem_bos / rem_eos),
// then (strip) leading / trailing tokens whose text is a single space
__device__ __forceinline__ RowSpan row_span(const uint8_t* __restrict__ row, int len, int rem_first, int rem_last, int strip,
                                            const uint8_t* __restrict__ table, int tok_stride, const uint8_t* __restrict__ tok_len) {
  int first = 0, last = len - 1;
  if (len > 0 && rem_first >= 0 && row[0] == rem_first) first = 1;
  if (last >= first && rem_last >= 0 && row[last] == rem_last) --last;
  if (strip) {
    while (first <= last && tok_len[row[first]] == 1 && table[(int)row[first] * tok_stride] == ' ') ++first;
    while (last >= first && tok_len[row[last]] == 1 && table[(int)row[last] * tok_stride] == ' ') --last;
  }
  return RowSpan{first, last};
}

__global__ void text_row_len_kernel(const uint8_t* __restrict__ ids, const int* __restrict__ lengths, int B, int L, int rem_first,
                                    int rem_last, int strip, const uint8_t* __restrict__ table, int tok_stride,
                                    const uint8_t* __restrict__ tok_len, int* __restrict__ row_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint8_t* row = ids + (long long)b * L;
  const int len = lengths ? min(max(lengths[b], 0), L) : L;
  const RowSpan s = row_span(row, len, rem_first, rem_last, strip, table, tok_stride, tok_len);
  int n = 0;
  for (int i = s.first; i <= s.last; ++i) n += tok_len[row[i]];
  row_len[b] = n;
}
// exclusive scan of row_len[B] -> offsets[B+1], one block
__global__ void text_scan_kernel(const int* __restrict__ row_len, int B, int* __restrict__ offsets) {
  __shared__ int carry;
  __shared__ int warp_tot[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < B ? row_len[i] : 0;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += t; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += warp_tot[w];
    if (i < B) offsets[i] = carry + wbase + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += wbase + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[B] = carry;
}
__global__ void text_write_kernel(const uint8_t* __restrict__ ids, const int* __restrict__ lengths, int B, int L, int rem_first,
                                  int rem_last, int strip, const uint8_t* __restrict__ table, int tok_stride,
                                  const uint8_t* __restrict__ tok_len, const int* __restrict__ offsets, long long capacity,
                                  uint8_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint8_t* row = ids + (long long)b * L;
  const int len = lengths ? min(max(lengths[b], 0), L) : L;
  const RowSpan s = row_span(row, len, rem_first, rem_last, strip, table, tok_stride, tok_len);
  long long o = offsets[b];
  for (int i = s.first; i <= s.last; ++i) {
    const int id = row[i], n = tok_len[id];
    for (int k = 0; k < n; ++k, ++o)
      if (o < capacity) out[o] = table[id * tok_stride + k];
  }
}

// ---- featurisation (featurizer.py:8-24, data_loader.py:26-31): packed SMILES bytes -> u8 ids (B,T), right-padded ----
// one thread per output position; a character without an id (lut == 255) or a string longer than T raises bad_flag
__global__ void text_to_ids_kernel(const uint8_t* __restrict__ text, const int* __restrict__ offsets, int B, int T,
                                   const uint8_t* __restrict__ lut, int pad_id, uint8_t* __restrict__ ids,
                                   int* __restrict__ bad_flag) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * T) return;
  const int b = (int)(i / T), t = (int)(i - (long long)b * T);
  const int o0 = offsets[b], n = offsets[b + 1] - o0;
  if (t == 0 && n > T) atomicOr(bad_flag, 2);
  int id = pad_id;
  if (t < n) {
    id = lut[text[o0 + t]];
    if (id == 255) { atomicOr(bad_flag, 1); id = pad_id; }
  }
  ids[i] = (uint8_t)id;
}

}  // namespace

extern "C" int mvae_ids_to_text(const uint8_t* ids, const int32_t* lengths, int B, int L, const uint8_t* table, int tok_stride,
                                const uint8_t* tok_len, int rem_first_id, int rem_last_id, int strip, uint8_t* out_bytes,
                                long long capacity, int32_t* out_offsets, int32_t* scratch_row_len, mvae_stream_t stream) {
  if (!ids || !table || !tok_len || !out_bytes || !out_offsets || !scratch_row_len || B <= 0 || L <= 0 || tok_stride <= 0 ||
      capacity < 0)
    return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  text_row_len_kernel<<<ceil_div(B, 128), 128, 0, st>>>(ids, lengths, B, L, rem_first_id, rem_last_id, strip, table, tok_stride,
                                                        tok_len, scratch_row_len); KCHECK();
  text_scan_kernel<<<1, 1024, 0, st>>>(scratch_row_len, B, out_offsets); KCHECK();
  text_write_kernel<<<ceil_div(B, 128), 128, 0, st>>>(ids, lengths, B, L, rem_first_id, rem_last_id, strip, table, tok_stride,
                                                      tok_len, out_offsets, capacity, out_bytes); KCHECK();
  return MVAE_OK;
}

extern "C" int mvae_text_to_ids(const uint8_t* text, const int32_t* offsets, int B, int T, const uint8_t* lut, int pad_id,
                                uint8_t* ids_out, int32_t* bad_flag, mvae_stream_t stream) {
  if (!text || !offsets || !lut || !ids_out || !bad_flag || B <= 0 || T <= 0 || pad_id < 0 || pad_id > 254) return MVAE_ERR_INVALID;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  mvae_count_launches(1);
  MVAE_CUDA_CHECK(cudaMemsetAsync(bad_flag, 0, 4, st));
  text_to_ids_kernel<<<(unsigned)ceil_div64((long long)B * T, 256), 256, 0, st>>>(text, offsets, B, T, lut, pad_id, ids_out, bad_flag); KCHECK();
  return MVAE_OK;
}
