// Stand-alone token log-prob / entropy from MATERIALISED logits (API parity with the reference's call surface;
// the fused lm_head path never builds this tensor). HBM-bound: one pass over the logits forward, one read + one
// write backward.
//
// Reference: verl/utils/torch_functional.py:34-66 log_probs_from_logits (flash-attn Triton cross-entropy branch:
// fp32 math on the stored logits, returns log p[label] as a negative number).
#pragma once
#include <cuda_fp16.h>
#include "ptx.cuh"

namespace grpo {

enum LogitsDtype : int { LOGITS_F32 = 0, LOGITS_BF16 = 1, LOGITS_F16 = 2 };

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

struct OnlineStat {
  float m, s, t;  // running max, sum exp(z - m), sum exp(z - m) * z
};
__device__ __forceinline__ void stat_push(OnlineStat& a, float z) {
  if (z == -INFINITY) return;  // masked-out logit: contributes nothing
  if (z > a.m) {
    const float w = __expf(a.m - z);
    a.s = a.s * w + 1.f;
    a.t = a.t * w + z;
    a.m = z;
  } else {
    const float e = __expf(z - a.m);
    a.s += e;
    a.t = fmaf(e, z, a.t);
  }
}
__device__ __forceinline__ OnlineStat stat_merge(const OnlineStat& a, const OnlineStat& b) {
  OnlineStat r;
  r.m = fmaxf(a.m, b.m);
  const float wa = (a.m == -INFINITY) ? 0.f : __expf(a.m - r.m);
  const float wb = (b.m == -INFINITY) ? 0.f : __expf(b.m - r.m);
  r.s = a.s * wa + b.s * wb;
  r.t = a.t * wa + b.t * wb;
  return r;
}

// A vector of logits at once: one max, at most one rescale of the running sums, then kVec exponentials with no
// data-dependent branch per element (the scalar form above diverges on every new maximum and spends a second
// exponential on it).
template <int kVec, bool kWantT>
__device__ __forceinline__ void stat_push_vec(OnlineStat& a, const float (&z)[kVec]) {
  float mx = z[0];
#pragma unroll
  for (int k = 1; k < kVec; ++k) mx = fmaxf(mx, z[k]);
  if (mx == -INFINITY) return;
  if (mx > a.m) {
    const float w = (a.m == -INFINITY) ? 0.f : __expf(a.m - mx);
    a.s *= w;
    a.t *= w;
    a.m = mx;
  }
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const float e = __expf(z[k] - a.m);  // exp(-inf) = 0 for masked-out logits
    a.s += e;
    if (kWantT && z[k] != -INFINITY) a.t = fmaf(e, z[k], a.t);  // sum e * z, only needed for the entropy
  }
}

template <typename T, int kVec>
__device__ __forceinline__ void unpack_vec(const uint4& q, float (&z)[kVec]) {
  const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
  for (int k = 0; k < kVec; ++k) z[k] = to_f32<T>(e[k]);
}

// one block per row; 16-byte loads, four of them in flight per thread. kWantT: also carry sum exp(z) * z (entropy).
template <typename T, bool kWantT>
__global__ void logprob_from_logits_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                           uint32_t rows, uint32_t vocab, int64_t ld, float* __restrict__ logp,
                                           float* __restrict__ entropy, float* __restrict__ lse_out) {
  const uint32_t r = blockIdx.x;
  if (r >= rows) return;
  const T* z = logits + static_cast<int64_t>(r) * ld;
  constexpr int kVec = 16 / sizeof(T);
  OnlineStat st{-INFINITY, 0.f, 0.f};
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) && (vocab % kVec == 0);
  if (vec_ok) {
    const uint4* z4 = reinterpret_cast<const uint4*>(z);
    const uint32_t nvec = vocab / kVec, step = blockDim.x;
    uint32_t i = threadIdx.x;
    for (; i + 3 * step < nvec; i += 4 * step) {
      const uint4 q0 = __ldcs(z4 + i), q1 = __ldcs(z4 + i + step), q2 = __ldcs(z4 + i + 2 * step), q3 = __ldcs(z4 + i + 3 * step);
      float f[kVec];
      unpack_vec<T, kVec>(q0, f);
      stat_push_vec<kVec, kWantT>(st, f);
      unpack_vec<T, kVec>(q1, f);
      stat_push_vec<kVec, kWantT>(st, f);
      unpack_vec<T, kVec>(q2, f);
      stat_push_vec<kVec, kWantT>(st, f);
      unpack_vec<T, kVec>(q3, f);
      stat_push_vec<kVec, kWantT>(st, f);
    }
    for (; i < nvec; i += step) {
      float f[kVec];
      unpack_vec<T, kVec>(__ldcs(z4 + i), f);
      stat_push_vec<kVec, kWantT>(st, f);
    }
  } else {
    for (uint32_t i = threadIdx.x; i < vocab; i += blockDim.x) stat_push(st, to_f32<T>(z[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    OnlineStat other;
    other.m = __shfl_xor_sync(0xffffffffu, st.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, st.s, o);
    other.t = __shfl_xor_sync(0xffffffffu, st.t, o);
    st = stat_merge(st, other);
  }
  __shared__ OnlineStat red[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) red[warp] = st;
  __syncthreads();
  if (warp == 0) {
    st = (lane < nwarps) ? red[lane] : OnlineStat{-INFINITY, 0.f, 0.f};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      OnlineStat other;
      other.m = __shfl_xor_sync(0xffffffffu, st.m, o);
      other.s = __shfl_xor_sync(0xffffffffu, st.s, o);
      other.t = __shfl_xor_sync(0xffffffffu, st.t, o);
      st = stat_merge(st, other);
    }
    if (lane == 0) {
      const float lse = st.m + logf(st.s);
      if (lse_out) lse_out[r] = lse;
      if (logp) {
        const int64_t lab = labels[r];
        logp[r] = (lab >= 0 && lab < static_cast<int64_t>(vocab)) ? to_f32<T>(z[lab]) - lse : 0.f;
      }
      if (entropy) entropy[r] = lse - st.t / st.s;
    }
  }
}

// dlogits[r][v] = dlogp[r] * (1[v == label] - softmax(z_r)[v]) + dent[r] * (-p (log p + H))   (dent optional)
__device__ __forceinline__ float dlogit_value(float zi, float l, float g, float ge, float h, bool is_label) {
  const float lp = zi - l;
  const float p = __expf(lp);
  float d = -g * p;
  if (is_label) d += g;
  if (ge != 0.f && p > 0.f) d -= ge * p * (lp + h);
  return d;
}

// grid (column blocks, rows); kVector: 16-byte loads and stores (rows 16-byte aligned, vocab a multiple of the vector)
template <typename T, bool kVector>
__global__ void logprob_from_logits_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels,
                                               const float* __restrict__ lse, const float* __restrict__ dlogp,
                                               const float* __restrict__ dent, const float* __restrict__ ent,
                                               uint32_t rows, uint32_t vocab, int64_t ld, T* __restrict__ dlogits,
                                               int64_t ld_out) {
  constexpr int kVec = 16 / sizeof(T);
  for (uint32_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const T* z = logits + static_cast<int64_t>(r) * ld;
    T* o = dlogits + static_cast<int64_t>(r) * ld_out;
    const float g = dlogp ? dlogp[r] : 0.f;
    const float ge = dent ? dent[r] : 0.f;
    const float h = (dent && ent) ? ent[r] : 0.f;
    const float l = lse[r];
    const int64_t lab = labels ? labels[r] : -1;
    if (kVector) {
      const uint4* z4 = reinterpret_cast<const uint4*>(z);
      uint4* o4 = reinterpret_cast<uint4*>(o);
      for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < vocab / kVec; i += gridDim.x * blockDim.x) {
        const uint4 q = z4[i];  // may alias o4[i] (in-place backward): plain load, read before the store below
        const T* e = reinterpret_cast<const T*>(&q);
        uint4 w;
        T* d = reinterpret_cast<T*>(&w);
        const int64_t c0 = static_cast<int64_t>(i) * kVec;
#pragma unroll
        for (int k = 0; k < kVec; ++k) d[k] = from_f32<T>(dlogit_value(to_f32<T>(e[k]), l, g, ge, h, c0 + k == lab));
        o4[i] = w;
      }
    } else {
      for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < vocab; i += gridDim.x * blockDim.x)
        o[i] = from_f32<T>(dlogit_value(to_f32<T>(z[i]), l, g, ge, h, static_cast<int64_t>(i) == lab));
    }
  }
}

}  // namespace grpo
