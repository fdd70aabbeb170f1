// GRPO outcome advantages (group-normalised sequence scores), HBM-bound.
//
// Reference: verl/trainer/core_algos.py:137-175 compute_grpo_outcome_advantage
//   scores = token_level_rewards.sum(-1); per uid group: mean, unbiased std; a_i = (s_i - mean) / (std + eps);
//   advantages[i, t] = a_i * response_mask[i, t]
// Group membership is arbitrary (rows are permuted by the trainer's sequence balancing before this runs), so the host
// hands over a CSR view of the groups: `order` = row ids sorted by group, `offsets` = group boundaries.
#pragma once
#include "loss_kernels.cuh"

namespace grpo {

// one warp per sequence: scores[i] = sum_t rewards[i][t]. Accumulated in fp64 and rounded once: with one reward per
// sequence (the reference's reward functions) that is the reference's fp32 sum exactly, with per-token rewards (after
// KL shaping) it is the correctly rounded sum, whatever order torch's own fp32 reduction happens to use.
__global__ void row_score_kernel(const float* __restrict__ rewards, uint32_t bsz, uint32_t t_len,
                                 float* __restrict__ scores) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= bsz) return;
  const float* r = rewards + static_cast<size_t>(row) * t_len;
  double s = 0.0;
  if ((t_len & 3u) == 0 && (reinterpret_cast<uintptr_t>(r) & 15u) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(r);
    for (uint32_t i = lane; i < (t_len >> 2); i += 32) {
      const float4 q = r4[i];
      s += (static_cast<double>(q.x) + static_cast<double>(q.y)) + (static_cast<double>(q.z) + static_cast<double>(q.w));
    }
  } else {
    for (uint32_t i = lane; i < t_len; i += 32) s += static_cast<double>(r[i]);
  }
  s = warp_sum(s);
  if (lane == 0) scores[row] = static_cast<float>(s);
}

// one warp per group: warp-shuffle mean / unbiased std in fp64, then the per-sequence normalised score
__global__ void group_stats_kernel(const float* __restrict__ scores, const int32_t* __restrict__ order,
                                   const int32_t* __restrict__ offsets, uint32_t n_groups, float eps,
                                   float* __restrict__ seq_adv) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  const int32_t beg = offsets[g], end = offsets[g + 1];
  const int32_t n = end - beg;
  double sum = 0.0;
  for (int32_t i = beg + lane; i < end; i += 32) sum += static_cast<double>(scores[order[i]]);
  sum = warp_sum(sum);
  const double mean = sum / static_cast<double>(n);
  double ss = 0.0;
  for (int32_t i = beg + lane; i < end; i += 32) {
    const double d = static_cast<double>(scores[order[i]]) - mean;
    ss += d * d;
  }
  ss = warp_sum(ss);
  // torch.std: unbiased (n - 1); n == 1 is rejected on the host exactly like the reference's assert
  const float mean32 = static_cast<float>(mean);
  const float std32 = static_cast<float>(sqrt(ss / static_cast<double>(n > 1 ? n - 1 : 1)));
  for (int32_t i = beg + lane; i < end; i += 32) {
    const int32_t row = order[i];
    seq_adv[row] = (scores[row] - mean32) / (std32 + eps);
  }
}

// advantages[i][t] = seq_adv[i] * mask[i][t]     grid (column blocks, row blocks): no per-element index division
__global__ void broadcast_adv_kernel(const float* __restrict__ seq_adv, const void* __restrict__ mask, int mask_dtype,
                                     uint32_t bsz, uint32_t t_len, float* __restrict__ adv) {
  const uint32_t t_step = gridDim.x * blockDim.x;
  for (uint32_t row = blockIdx.y; row < bsz; row += gridDim.y) {
    const float a = seq_adv[row];
    const size_t base = static_cast<size_t>(row) * t_len;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (; t + 3 * t_step < t_len; t += 4 * t_step) {
      float m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u] = load_mask(mask, mask_dtype, base + t + u * t_step);
#pragma unroll
      for (int u = 0; u < 4; ++u) adv[base + t + u * t_step] = a * m[u];
    }
    for (; t < t_len; t += t_step) adv[base + t] = a * load_mask(mask, mask_dtype, base + t);
  }
}

}  // namespace grpo
