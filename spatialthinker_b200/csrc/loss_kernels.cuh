// Token-level GRPO loss kernels (HBM-bound elementwise + reductions).
//
// Reference arithmetic being restated on the device:
//   verl/trainer/core_algos.py:291-353   compute_policy_loss  (ratio / clip / dual-clip, four masked means)
//   verl/trainer/core_algos.py:394-436   compute_kl           (kl, abs, mse, low_var_kl, chi2)
//   verl/utils/torch_functional.py:69-71 masked_mean          sum(x*mask) / (sum(mask) + 1e-8)
//   verl/workers/actor/dp_actor.py:253-278  entropy estimator, loss = (pg + kl_coef*kl) / grad_accum, backward
#pragma once
#include "ptx.cuh"

namespace grpo {

enum KlMode : int { KL_NONE = -1, KL_LOW_VAR = 0, KL_KL = 1, KL_ABS = 2, KL_MSE = 3, KL_CHI2 = 4 };
enum MaskDtype : int { MASK_F32 = 0, MASK_I64 = 1, MASK_U8 = 2, MASK_NONE = 3 };

struct LossCfg {
  float log_clip_lo;  // ln(1 - clip_ratio_low), rounded to fp32 like torch.clamp does with its scalar bounds
  float log_clip_hi;  // ln(1 + clip_ratio_high)
  float clip_dual;
  float kl_coef;
  int kl_mode;
  float inv_grad_accum;
  float entropy_coef;  // loss -= entropy_coef * masked_mean(true entropy); 0 in the reference (entropy is only logged)
};

// accumulator slots (double)
enum { ACC_MASK = 0, ACC_PG = 1, ACC_CF_HI = 2, ACC_CF_LO = 3, ACC_NEG_X = 4, ACC_KL = 5, ACC_LOGP = 6, ACC_ENT = 7, ACC_SAT = 8, ACC_N = 9 };
// log p at or below which the fused head's exp2 clamp (kClampLog2 = 100 in lmhead_kernels.cuh) may have bound: a clamped
// element contributes exactly 2^100 to the row sum, so sum >= 2^100 <=> log p[label] <= -100 ln 2
constexpr float kSaturatedLogp = -69.3f;
// metric slots (float) written by loss_finalize_kernel
enum {
  MET_PG_LOSS = 0,      // masked_mean(policy loss)                      (core_algos.py:349)
  MET_CLIPFRAC_HI = 1,  // masked_mean(pg_loss < pg_loss2)               (:350)
  MET_CLIPFRAC_LO = 2,  // masked_mean(clipped_hi > pg_loss3 and A < 0)  (:351)
  MET_PPO_KL = 3,       // masked_mean(-(logp - old))                    (:352)
  MET_KL_LOSS = 4,      // masked_mean(compute_kl(...))                  (dp_actor.py:270)
  MET_ENTROPY = 5,      // -masked_mean(logp)                            (dp_actor.py:253)
  MET_TOTAL = 6,        // pg_loss + kl_coef * kl_loss                   (dp_actor.py:271)
  MET_SCALED = 7,       // total / grad_accum - the value that is back-propagated (dp_actor.py:277)
  MET_TRUE_ENTROPY = 8, // masked_mean(lse - sum p z) when the per-token entropy was requested, else 0
  MET_MASK_SUM = 9,     // sum(mask): the micro-batch's valid-token count
  MET_SATURATED = 10,   // number of unmasked tokens with log p <= -69.3: the fused head's lse may be saturated there
  MET_N = 11
};

__device__ __forceinline__ float load_mask(const void* m, int dtype, size_t i) {
  switch (dtype) {
    case MASK_F32: return static_cast<const float*>(m)[i];
    case MASK_I64: return static_cast<float>(static_cast<const long long*>(m)[i]);
    case MASK_U8: return static_cast<float>(static_cast<const unsigned char*>(m)[i]);
    default: return 1.f;
  }
}

// Grid-stride loop with kBatch iterations in flight: every load of a batch is issued before the first use, so each
// thread keeps kBatch x (number of input streams) independent requests outstanding - the memory-level parallelism a
// 4-byte-per-thread elementwise kernel needs to cover HBM latency at 148 SMs x 2048 threads.
//   ld(i) -> In (all global loads of element i)      use(i, in) (arithmetic + stores of element i)
template <int kBatch, class Ld, class Use>
__device__ __forceinline__ void batched_grid_stride(size_t n, Ld&& ld, Use&& use) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + (kBatch - 1) * stride < n; i += kBatch * stride) {
    decltype(ld(i)) in[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) in[u] = ld(i + u * stride);
#pragma unroll
    for (int u = 0; u < kBatch; ++u) use(i + u * stride, in[u]);
  }
  for (; i < n; i += stride) {
    const auto in = ld(i);
    use(i, in);
  }
}
constexpr int kEwBatch = 4;

// value and d/dlogp of one KL estimator (core_algos.py:408-434)
__device__ __forceinline__ void kl_term(int mode, float logp, float ref, float& val, float& dlogp) {
  switch (mode) {
    case KL_LOW_VAR: {
      const float k = ref - logp;
      const float ek = expf(k);
      const float raw = ek - k - 1.f;
      val = fminf(fmaxf(raw, -10.f), 10.f);
      dlogp = (raw >= -10.f && raw <= 10.f) ? (1.f - ek) : 0.f;
      break;
    }
    case KL_KL: val = logp - ref; dlogp = 1.f; break;
    case KL_ABS: {
      const float d = logp - ref;
      val = fabsf(d);
      dlogp = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
      break;
    }
    case KL_MSE: {
      const float d = logp - ref;
      val = 0.5f * d * d;
      dlogp = d;
      break;
    }
    case KL_CHI2: {
      const float r = expf(ref - logp);
      const float raw = (r - 1.f) * (r - 1.f);
      val = fminf(fmaxf(raw, 0.f), 20.f);
      dlogp = (raw >= 0.f && raw <= 20.f) ? (-2.f * (r - 1.f) * r) : 0.f;
      break;
    }
    default: val = 0.f; dlogp = 0.f;
  }
}

// value and d/dlogp of the clipped policy-gradient term for one token (core_algos.py:331-347)
__device__ __forceinline__ void pg_term(const LossCfg& c, float logp, float old, float adv, float& loss, float& dlogp,
                                        float& cf_hi, float& cf_lo, float& x_out) {
  const float x = logp - old;
  const float r = expf(x);
  const float xc = fminf(fmaxf(x, c.log_clip_lo), c.log_clip_hi);
  const bool in_range = (x >= c.log_clip_lo) && (x <= c.log_clip_hi);
  const float rc = expf(xc);
  const float l1 = -adv * r, l2 = -adv * rc, l3 = -adv * c.clip_dual;
  const float hi = fmaxf(l1, l2);
  // d hi / dx : torch.max routes the gradient to the larger branch and splits it evenly on ties
  const float d1 = -adv * r;
  const float d2 = in_range ? -adv * rc : 0.f;
  float dhi;
  if (l1 > l2) dhi = d1;
  else if (l1 < l2) dhi = d2;
  else dhi = 0.5f * d1 + 0.5f * d2;
  cf_hi = (l1 < l2) ? 1.f : 0.f;
  if (adv < 0.f) {
    loss = fminf(hi, l3);
    dlogp = (hi < l3) ? dhi : ((hi == l3) ? 0.5f * dhi : 0.f);
    cf_lo = (hi > l3) ? 1.f : 0.f;
  } else {
    loss = hi;
    dlogp = dhi;
    cf_lo = 0.f;
  }
  x_out = x;
}

template <int N>
__device__ __forceinline__ void block_accumulate(float (&v)[N], double* acc, const int (&slot)[N]) {
  __shared__ float red[N][32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) red[i][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float s = (lane < nwarps) ? red[i][lane] : 0.f;
      s = warp_sum(s);
      if (lane == 0 && s != 0.f) atomicAdd(&acc[slot[i]], static_cast<double>(s));
    }
  }
}

// acc[ACC_MASK] += sum(mask)
__global__ void mask_sum_kernel(const void* __restrict__ mask, int mask_dtype, size_t n, double* __restrict__ acc) {
  float v[1] = {0.f};
  batched_grid_stride<2 * kEwBatch>(
      n, [&](size_t i) { return load_mask(mask, mask_dtype, i); }, [&](size_t, float m) { v[0] += m; });
  const int slot[1] = {ACC_MASK};
  block_accumulate<1>(v, acc, slot);
}

// Per token: loss terms, metric partial sums, and dL/dlogp with the masked-mean and grad-accum factors folded in.
// acc[ACC_MASK] must already hold sum(mask) over the whole micro-batch (the normaliser is per micro-batch,
// dp_actor.py:255-277), so this kernel can be launched chunk by chunk as log-probs become available.
__global__ void token_loss_kernel(const float* __restrict__ logp, const float* __restrict__ old_logp,
                                  const float* __restrict__ adv, const float* __restrict__ ref_logp,
                                  const float* __restrict__ entropy, const void* __restrict__ mask, int mask_dtype,
                                  size_t n, LossCfg cfg, double* __restrict__ acc, float* __restrict__ dlogp_out,
                                  float* __restrict__ dent_out) {
  const float denom = static_cast<float>(acc[ACC_MASK]) + 1e-8f;
  const float wnorm = cfg.inv_grad_accum / denom;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  struct In {
    float m, lp, old, adv, ref, ent;
  };
  const bool use_kl = cfg.kl_mode != KL_NONE && ref_logp != nullptr;
  batched_grid_stride<kEwBatch>(
      n,
      [&](size_t i) {
        In in;
        in.m = load_mask(mask, mask_dtype, i);
        in.lp = logp[i];
        in.old = old_logp[i];
        in.adv = adv[i];
        in.ref = use_kl ? ref_logp[i] : 0.f;
        in.ent = entropy ? entropy[i] : 0.f;
        return in;
      },
      [&](size_t i, const In& in) {
        const float m = in.m, lp = in.lp;
        float loss, dpg, cfh, cfl, x;
        pg_term(cfg, lp, in.old, in.adv, loss, dpg, cfh, cfl, x);
        float klv = 0.f, dkl = 0.f;
        if (use_kl) kl_term(cfg.kl_mode, lp, in.ref, klv, dkl);
        if (m != 0.f) {  // reference multiplies by the mask: garbage at padded positions never contributes
          v[0] += loss * m;
          v[1] += cfh * m;
          v[2] += cfl * m;
          v[3] += -x * m;
          v[4] += klv * m;
          v[5] += lp * m;
          if (entropy) v[6] += in.ent * m;
          if (lp <= kSaturatedLogp) v[7] += 1.f;
        }
        if (dlogp_out) dlogp_out[i] = (m != 0.f) ? m * wnorm * (dpg + cfg.kl_coef * dkl) : 0.f;
        if (dent_out) dent_out[i] = (m != 0.f) ? -cfg.entropy_coef * m * wnorm : 0.f;
      });
  const int slot[8] = {ACC_PG, ACC_CF_HI, ACC_CF_LO, ACC_NEG_X, ACC_KL, ACC_LOGP, ACC_ENT, ACC_SAT};
  block_accumulate<8>(v, acc, slot);
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, LossCfg cfg, float* __restrict__ metrics) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float denom = static_cast<float>(acc[ACC_MASK]) + 1e-8f;
  const float pg = static_cast<float>(acc[ACC_PG]) / denom;
  const float kl = static_cast<float>(acc[ACC_KL]) / denom;
  metrics[MET_PG_LOSS] = pg;
  metrics[MET_CLIPFRAC_HI] = static_cast<float>(acc[ACC_CF_HI]) / denom;
  metrics[MET_CLIPFRAC_LO] = static_cast<float>(acc[ACC_CF_LO]) / denom;
  metrics[MET_PPO_KL] = static_cast<float>(acc[ACC_NEG_X]) / denom;
  metrics[MET_KL_LOSS] = kl;
  metrics[MET_ENTROPY] = -static_cast<float>(acc[ACC_LOGP]) / denom;
  const float ent = static_cast<float>(acc[ACC_ENT]) / denom;
  float total = (cfg.kl_mode != KL_NONE) ? pg + kl * cfg.kl_coef : pg;
  if (cfg.entropy_coef != 0.f) total -= cfg.entropy_coef * ent;
  metrics[MET_TOTAL] = total;
  metrics[MET_SCALED] = total * cfg.inv_grad_accum;
  metrics[MET_TRUE_ENTROPY] = ent;
  metrics[MET_MASK_SUM] = static_cast<float>(acc[ACC_MASK]);
  metrics[MET_SATURATED] = static_cast<float>(acc[ACC_SAT]);
}

// Elementwise KL estimator with its derivative (standalone compute_kl surface).
__global__ void kl_elementwise_kernel(const float* __restrict__ logp, const float* __restrict__ ref, size_t n, int mode,
                                      float* __restrict__ out, float* __restrict__ dout_dlogp) {
  batched_grid_stride<kEwBatch>(
      n, [&](size_t i) { return make_float2(logp[i], ref[i]); },
      [&](size_t i, const float2& in) {
        float v, d;
        kl_term(mode, in.x, in.y, v, d);
        out[i] = v;
        if (dout_dlogp) dout_dlogp[i] = d;
      });
}

// masked_mean over all elements: out[0] = sum(x*mask) / (sum(mask) + eps)
__global__ void masked_sum_kernel(const float* __restrict__ x, const void* __restrict__ mask, int mask_dtype, size_t n,
                                  double* __restrict__ acc /*[2]: sum(x*m), sum(m)*/) {
  float v[2] = {0.f, 0.f};
  batched_grid_stride<kEwBatch>(
      n, [&](size_t i) { return make_float2(x[i], load_mask(mask, mask_dtype, i)); },
      [&](size_t, const float2& in) {
        if (in.y != 0.f) v[0] += in.x * in.y;  // a masked-out value may be anything, NaN included
        v[1] += in.y;
      });
  const int slot[2] = {0, 1};
  block_accumulate<2>(v, acc, slot);
}
__global__ void masked_mean_finalize_kernel(const double* __restrict__ acc, float eps, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<float>(acc[0]) / (static_cast<float>(acc[1]) + eps);
}

}  // namespace grpo
