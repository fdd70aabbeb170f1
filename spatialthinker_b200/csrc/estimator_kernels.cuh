// The other advantage estimators, masked whitening, the value loss and the KL reward shaping of the reference's
// trainer, on the same HBM-bound skeleton as the GRPO kernels: grid-stride elementwise passes,
// warp-per-sequence / warp-per-group reductions, block reductions into fp64 accumulators.
//
// Reference arithmetic being restated on the device:
//   verl/trainer/core_algos.py:93-133    compute_gae_advantage_return                   (reverse recurrence + whiten)
//   verl/trainer/core_algos.py:179-214   compute_rloo_outcome_advantage                 (leave-one-out group baseline)
//   verl/trainer/core_algos.py:217-245   compute_reinforce_plus_plus_outcome_advantage  (discounted return + whiten)
//   verl/trainer/core_algos.py:248-273   compute_remax_outcome_advantage                (score - greedy baseline)
//   verl/trainer/core_algos.py:356-391   compute_value_loss                             (clipped value loss)
//   verl/utils/torch_functional.py:74-97 masked_var / masked_whiten
//   verl/trainer/ray_trainer.py:125-145  apply_kl_penalty                               (token rewards - beta * KL)
#pragma once
#include "advantage_kernels.cuh"

namespace grpo {

// ------------------------------------------------------------------------------------------ RLOO / ReMax
// one warp per group: a_i = s_i - (sum_g - s_i) / (n_g - 1)
__global__ void group_rloo_kernel(const float* __restrict__ scores, const int32_t* __restrict__ order,
                                  const int32_t* __restrict__ offsets, uint32_t n_groups,
                                  float* __restrict__ seq_adv) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (g >= n_groups) return;
  const int32_t beg = offsets[g], end = offsets[g + 1];
  const int32_t n = end - beg;
  double sum = 0.0;
  for (int32_t i = beg + lane; i < end; i += 32) sum += static_cast<double>(scores[order[i]]);
  const float sum32 = static_cast<float>(warp_sum(sum));  // the reference's group sum is an fp32 tensor
  const float nm1 = static_cast<float>(n > 1 ? n - 1 : 1);
  for (int32_t i = beg + lane; i < end; i += 32) {
    const int32_t row = order[i];
    const float s = scores[row];
    seq_adv[row] = __fsub_rn(s, __fdiv_rn(__fsub_rn(sum32, s), nm1));
  }
}

// seq_adv[i] = scores[i] - baselines[i]
__global__ void remax_seq_kernel(const float* __restrict__ scores, const float* __restrict__ baselines, uint32_t bsz,
                                 float* __restrict__ seq_adv) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < bsz) seq_adv[i] = __fsub_rn(scores[i], baselines[i]);
}

// ------------------------------------------------------------------------------------------ reverse recurrences
// The reference walks the response from its last token to its first with one fp32 recurrence per sequence. Here a
// warp owns 32 sequences: 32 x 32 tiles travel through shared memory so that global accesses are coalesced along the
// token axis while lane r runs sequence r's recurrence in the reference's own operation order (no FMA contraction:
// the results are the reference's bit for bit). The kernel is latency-bound, not bandwidth-bound (one dependency chain
// of T steps per sequence, 32 sequences per SM-resident warp), so the loads of the NEXT tile are issued into registers
// before the current tile's recurrence runs, all 64 of them independent.
constexpr int kScanTile = 32;

// mode 0 - GAE (core_algos.py:122-130):   delta_t = (r_t + gamma * v_{t+1}) - v_t;   A_t = delta_t + (gamma*lam) * A_{t+1}
//                                         out_a = A (before whitening), out_b = A + v (returns)
// mode 1 - REINFORCE++ (:236-242):        R_t = r_t + gamma * (R_{t+1} * mask_{t+1});   out_b = R (returns)
// MaskT / kHasMask: element type and presence of the mask (mode 1; no mask = all ones). Compile-time, not load_mask's
// run-time switch: a branch per element between the 64 loads of a tile serialises them, each waiting out a DRAM round
// trip (measured: 4x slower).
template <int kMode, typename MaskT, bool kHasMask>
__global__ void __launch_bounds__(32)
reverse_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const MaskT* __restrict__ mask,
                    uint32_t bsz, uint32_t t_len, float gamma, float gamma_lam, float* __restrict__ out_a,
                    float* __restrict__ out_b) {
  __shared__ float x[kScanTile][kScanTile + 1];
  __shared__ float y[kScanTile][kScanTile + 1];
  const uint32_t lane = threadIdx.x;
  const uint32_t row0 = blockIdx.x * kScanTile;
  if (row0 >= bsz) return;
  const uint32_t nrows = min(static_cast<uint32_t>(kScanTile), bsz - row0);
  const uint32_t tiles = (t_len + kScanTile - 1) / kScanTile;
  float rx[kScanTile], ry[kScanTile];
  auto fetch = [&](uint32_t tb) {
    const uint32_t t = tb * kScanTile + lane;
    const bool col_ok = t < t_len;
#pragma unroll
    for (int r = 0; r < kScanTile; ++r) {
      const bool ok = col_ok && static_cast<uint32_t>(r) < nrows;
      const size_t idx = static_cast<size_t>(row0 + r) * t_len + t;
      rx[r] = ok ? rewards[idx] : 0.f;
      if (kMode == 0) ry[r] = ok ? values[idx] : 0.f;
      else if (kHasMask) ry[r] = ok ? static_cast<float>(mask[idx]) : 0.f;
      else ry[r] = ok ? 1.f : 0.f;
    }
  };
  float carry = 0.f;       // A_{t+1} or R_{t+1} * mask_{t+1}
  float next_value = 0.f;  // v_{t+1} (GAE)
  fetch(tiles - 1);
  for (uint32_t tb = tiles; tb-- > 0;) {
    const uint32_t t0 = tb * kScanTile;
#pragma unroll
    for (int r = 0; r < kScanTile; ++r) {
      x[r][lane] = rx[r];
      y[r][lane] = ry[r];
    }
    __syncwarp();
    if (tb > 0) fetch(tb - 1);  // in flight while the recurrence below runs
    if (lane < nrows) {
      // this lane's 32 tokens go to registers first (64 independent shared loads), so that the dependency chain below
      // is arithmetic only. Slots past t_len hold zeros and leave the carried state at +0, exactly as if skipped.
      float xs[kScanTile], ys[kScanTile];
#pragma unroll
      for (int j = 0; j < kScanTile; ++j) {
        xs[j] = x[lane][j];
        ys[j] = y[lane][j];
      }
#pragma unroll
      for (int j = kScanTile - 1; j >= 0; --j) {
        if (kMode == 0) {
          const float v = ys[j];
          const float delta = __fsub_rn(__fadd_rn(xs[j], __fmul_rn(gamma, next_value)), v);
          carry = __fadd_rn(delta, __fmul_rn(gamma_lam, carry));
          next_value = v;
          xs[j] = carry;
          ys[j] = __fadd_rn(carry, v);
        } else {
          const float ret = __fadd_rn(xs[j], __fmul_rn(gamma, carry));
          carry = __fmul_rn(ret, ys[j]);
          ys[j] = ret;
        }
      }
#pragma unroll
      for (int j = 0; j < kScanTile; ++j) {
        if (kMode == 0) x[lane][j] = xs[j];
        y[lane][j] = ys[j];
      }
    }
    __syncwarp();
    const uint32_t t = t0 + lane;
    if (t < t_len) {
#pragma unroll 8
      for (uint32_t r = 0; r < nrows; ++r) {
        const size_t idx = static_cast<size_t>(row0 + r) * t_len + t;
        if (kMode == 0) out_a[idx] = x[r][lane];
        out_b[idx] = y[r][lane];
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------ masked variance / whitening
// acc[0] = sum(x*m), acc[1] = sum(m) come from masked_sum_kernel; this pass adds acc[2] = sum((x - mean)^2 * m).
__device__ __forceinline__ float whiten_mean(const double* acc, float eps) {
  return static_cast<float>(acc[0]) / (static_cast<float>(acc[1]) + eps);
}
__global__ void masked_centered_sq_kernel(const float* __restrict__ x, const void* __restrict__ mask, int mask_dtype,
                                          size_t n, double* __restrict__ acc) {
  const float mean = whiten_mean(acc, 1e-8f);  // masked_var calls masked_mean with its default eps
  float v[1] = {0.f};
  batched_grid_stride<kEwBatch>(
      n, [&](size_t i) { return make_float2(x[i], load_mask(mask, mask_dtype, i)); },
      [&](size_t, const float2& in) {
        if (in.y != 0.f) {
          const float d = in.x - mean;
          v[0] += d * d * in.y;
        }
      });
  const int slot[1] = {2};
  block_accumulate<1>(v, acc, slot);
}
// torch_functional.py:74-89: biased variance = masked_mean(centered^2); Bessel's correction unless sum(mask) <= 1
__device__ __forceinline__ float whiten_var(const double* acc, int unbiased) {
  const float msum = static_cast<float>(acc[1]);
  float var = static_cast<float>(acc[2]) / (msum + 1e-8f);
  if (unbiased && msum > 1.f) var = var * (msum / (msum - 1.f));
  return var;
}
__global__ void masked_var_finalize_kernel(const double* __restrict__ acc, int unbiased, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = whiten_var(acc, unbiased);
    out[1] = whiten_mean(acc, 1e-8f);
  }
}
// out = (x - mean) * rsqrt(var + eps)   (torch_functional.py:92-95; every position, masked or not, as the reference)
// x and out may alias (grpo_gae_advantage / grpo_masked_whiten whiten in place): no __restrict__ on them.
__global__ void whiten_apply_kernel(const float* x, size_t n, const double* __restrict__ acc, float eps, float* out) {
  const float mean = whiten_mean(acc, 1e-8f);
  const float scale = __fdiv_rn(1.f, __fsqrt_rn(whiten_var(acc, 1) + eps));
  batched_grid_stride<2 * kEwBatch>(
      n, [&](size_t i) { return x[i]; }, [&](size_t i, float xi) { out[i] = (xi - mean) * scale; });
}

// ------------------------------------------------------------------------------------------ value loss
// core_algos.py:386-390. acc[0] = sum(mask) must be complete (mask_sum_kernel) before this runs when dvpreds is wanted.
//   acc[1] += sum(max(l1, l2) * m), acc[2] += sum((l1 < l2) * m);   dvpreds = d(0.5 * masked_mean(max(l1, l2))) / dvpreds
__global__ void value_loss_kernel(const float* __restrict__ vpreds, const float* __restrict__ returns,
                                  const float* __restrict__ values, const void* __restrict__ mask, int mask_dtype,
                                  size_t n, float cliprange, double* __restrict__ acc, float* __restrict__ dvpreds) {
  const float wnorm = 0.5f / (static_cast<float>(acc[0]) + 1e-8f);
  float v[2] = {0.f, 0.f};
  struct In {
    float m, vp, ret, old;
  };
  batched_grid_stride<kEwBatch>(
      n,
      [&](size_t i) {
        In in;
        in.m = load_mask(mask, mask_dtype, i);
        in.vp = vpreds[i];
        in.ret = returns[i];
        in.old = values[i];
        return in;
      },
      [&](size_t i, const In& in) {
        const float m = in.m, vp = in.vp;
        const float lo = in.old - cliprange, hi = in.old + cliprange;
        const float vc = fminf(fmaxf(vp, lo), hi);
        const bool in_range = vp >= lo && vp <= hi;
        const float e1 = vp - in.ret, e2 = vc - in.ret;
        const float l1 = e1 * e1, l2 = e2 * e2;
        if (m != 0.f) {
          v[0] += fmaxf(l1, l2) * m;
          v[1] += (l1 < l2) ? m : 0.f;
        }
        if (dvpreds) {
          const float d1 = 2.f * e1, d2 = in_range ? 2.f * e2 : 0.f;
          const float d = (l1 > l2) ? d1 : ((l1 < l2) ? d2 : 0.5f * d1 + 0.5f * d2);  // torch.max splits ties evenly
          dvpreds[i] = (m != 0.f) ? m * wnorm * d : 0.f;
        }
      });
  const int slot[2] = {1, 2};
  block_accumulate<2>(v, acc, slot);
}
__global__ void value_loss_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float denom = static_cast<float>(acc[0]) + 1e-8f;
    out[0] = 0.5f * (static_cast<float>(acc[1]) / denom);
    out[1] = static_cast<float>(acc[2]) / denom;
  }
}

// ------------------------------------------------------------------------------------------ KL reward shaping
// ray_trainer.py:125-145, one warp per sequence:
//   kld = compute_kl(old_logp, ref_logp) * mask;  rewards = scores - beta * kld;
//   acc[0] += masked_mean(kld, mask, dim=-1) of this sequence   (current_kl = acc[0] / bsz)
__global__ void kl_reward_kernel(const float* __restrict__ scores, const float* __restrict__ logp,
                                 const float* __restrict__ ref, const void* __restrict__ mask, int mask_dtype,
                                 uint32_t bsz, uint32_t t_len, int kl_mode, float beta, float* __restrict__ rewards,
                                 double* __restrict__ acc) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= bsz) return;
  const size_t base = static_cast<size_t>(row) * t_len;
  float num = 0.f, den = 0.f;
  for (uint32_t t = lane; t < t_len; t += 32) {
    const size_t i = base + t;
    const float m = load_mask(mask, mask_dtype, i);
    float kld = 0.f;
    if (ref != nullptr) {
      float d;
      kl_term(kl_mode, logp[i], ref[i], kld, d);
      kld = (m != 0.f) ? kld * m : 0.f;
    }
    rewards[i] = __fsub_rn(scores[i], __fmul_rn(beta, kld));
    num += kld * m;
    den += m;
  }
  num = warp_sum(num);
  den = warp_sum(den);
  if (lane == 0) atomicAdd(&acc[0], static_cast<double>(num / (den + 1e-8f)));
}
__global__ void kl_reward_finalize_kernel(const double* __restrict__ acc, uint32_t bsz, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(acc[0] / static_cast<double>(bsz));
}

}  // namespace grpo
