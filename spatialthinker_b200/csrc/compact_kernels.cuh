// Ragged micro-batches: drop the padded response slots before the GEMMs and put the results back afterwards.
//
// Reference: verl/workers/actor/dp_actor.py:136-139 computes log-probs for every slot of the padded [B, T] response
// block and multiplies the padded ones by 0 afterwards; its padding-free branch (:86-104, flash-attn bert_padding
// unpad_input / pad_input) does this gather / scatter for the transformer body only. Here the head itself runs on the
// unmasked rows: stable compaction index (three small kernels, no host sync), row gather, row scatter with zero fill.
// All HBM-bound: 4 B / slot for the index, 2 * row_bytes per moved row.
#pragma once
#include "loss_kernels.cuh"

namespace grpo {

constexpr int kCompactThreads = 256;
constexpr int kCompactPerThread = 8;
constexpr int kCompactBlock = kCompactThreads * kCompactPerThread;  // slots per block

// block_counts[b] = number of unmasked slots among [b * 2048, (b + 1) * 2048)
__global__ void compact_count_kernel(const void* __restrict__ mask, int mask_dtype, size_t n,
                                     int32_t* __restrict__ block_counts) {
  __shared__ int32_t warp_tot[kCompactThreads / 32];
  const size_t base = static_cast<size_t>(blockIdx.x) * kCompactBlock + threadIdx.x * kCompactPerThread;
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < kCompactPerThread; ++i)
    if (base + i < n && load_mask(mask, mask_dtype, base + i) != 0.f) ++c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
#pragma unroll
    for (int w = 0; w < kCompactThreads / 32; ++w) t += warp_tot[w];
    block_counts[blockIdx.x] = t;
  }
}

// in place: block_counts[b] -> exclusive prefix; count[0] = total. One block, any number of entries.
__global__ void compact_scan_kernel(int32_t* __restrict__ block_counts, uint32_t nb, int32_t* __restrict__ count) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t b0 = 0; b0 < nb; b0 += blockDim.x) {
    const uint32_t b = b0 + threadIdx.x;
    const int32_t v = b < nb ? block_counts[b] : 0;
    int32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= static_cast<uint32_t>(o)) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
      int32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= static_cast<uint32_t>(o)) wi += t;
      }
      warp_tot[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int32_t carry = carry_s;
    if (b < nb) block_counts[b] = carry + warp_tot[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[warp] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = carry_s;
}

// gather_idx[j] = slot of the j-th unmasked entry (original order); inverse[slot] = j, or -1 for a masked slot
__global__ void compact_write_kernel(const void* __restrict__ mask, int mask_dtype, size_t n,
                                     const int32_t* __restrict__ block_offsets, int32_t* __restrict__ gather_idx,
                                     int32_t* __restrict__ inverse) {
  __shared__ int32_t warp_tot[kCompactThreads / 32];
  const size_t base = static_cast<size_t>(blockIdx.x) * kCompactBlock + threadIdx.x * kCompactPerThread;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t bits = 0;
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < kCompactPerThread; ++i)
    if (base + i < n && load_mask(mask, mask_dtype, base + i) != 0.f) {
      bits |= 1u << i;
      ++c;
    }
  int32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= static_cast<uint32_t>(o)) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  int32_t pos = block_offsets[blockIdx.x] + incl - c;
  for (uint32_t w = 0; w < warp; ++w) pos += warp_tot[w];
#pragma unroll
  for (int i = 0; i < kCompactPerThread; ++i) {
    if (base + i >= n) break;
    if (bits & (1u << i)) {
      gather_idx[pos] = static_cast<int32_t>(base + i);
      inverse[base + i] = pos++;
    } else {
      inverse[base + i] = -1;
    }
  }
}

// out[j][:] = in[gather_idx[j]][:], j < m.  kVec = uint4 (rows of 16-byte multiples: hidden states) or uint32_t
// (labels, log-probs, advantages, masks). One warp per row for wide rows, one thread per element for narrow ones.
template <class Vec>
__global__ void gather_rows_kernel(const Vec* __restrict__ in, const int32_t* __restrict__ gather_idx, size_t m,
                                   uint32_t vecs_per_row, Vec* __restrict__ out) {
  if (vecs_per_row >= 32) {
    const size_t j = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    if (j >= m) return;
    const Vec* src = in + static_cast<size_t>(gather_idx[j]) * vecs_per_row;
    Vec* dst = out + j * vecs_per_row;
    for (uint32_t i = threadIdx.x & 31; i < vecs_per_row; i += 32) dst[i] = src[i];
  } else {
    const size_t total = m * vecs_per_row;
    for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t j = e / vecs_per_row;
      out[e] = in[static_cast<size_t>(gather_idx[j]) * vecs_per_row + (e - j * vecs_per_row)];
    }
  }
}

// out[i][:] = inverse[i] >= 0 ? in[inverse[i]][:] : 0, i < n  (every output row is written: no separate memset)
template <class Vec>
__global__ void scatter_rows_kernel(const Vec* __restrict__ in, const int32_t* __restrict__ inverse, size_t n,
                                    uint32_t vecs_per_row, Vec* __restrict__ out) {
  Vec zero;
  memset(&zero, 0, sizeof(Vec));
  if (vecs_per_row >= 32) {
    const size_t i = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5;
    if (i >= n) return;
    const int32_t j = inverse[i];
    const Vec* src = in + static_cast<size_t>(j < 0 ? 0 : j) * vecs_per_row;
    Vec* dst = out + i * vecs_per_row;
    for (uint32_t k = threadIdx.x & 31; k < vecs_per_row; k += 32) dst[k] = j < 0 ? zero : src[k];
  } else {
    const size_t total = n * vecs_per_row;
    for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<size_t>(gridDim.x) * blockDim.x) {
      const size_t i = e / vecs_per_row;
      const int32_t j = inverse[i];
      out[e] = j < 0 ? zero : in[static_cast<size_t>(j) * vecs_per_row + (e - i * vecs_per_row)];
    }
  }
}

}  // namespace grpo
