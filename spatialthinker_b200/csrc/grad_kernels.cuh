// Per-optimizer-step passes over the fp32 lm_head weight-gradient accumulator, and the run-to-run reproducible variant
// of the one-hot part of dW.
//
// Reference being replaced: verl/workers/actor/dp_actor.py:155-167 (_optimizer_step: clip_grad_norm_ over the module's
// parameters, skip on a non-finite norm) - there autograd's bf16 .grad tensors are walked by torch's foreach kernels; here
// the head's gradient lives in ONE fp32 [V, H] buffer (2.18 GB at the 7B head), so every pass over it is HBM-bound and
// the passes are fused: sum of squares (+ zeroing when nothing else needs the values), scale + cast to the parameter
// dtype + zeroing for the optimizer.
#pragma once
#include "ptx.cuh"

namespace grpo {

constexpr int kGradBlocks = 148 * 4;  // partial sums: one double per block, reduced in a fixed order (reproducible)
constexpr int kGradThreads = 512;

// partial[blockIdx] = sum over this block's grid-stride share of x^2 (fp32 products, fp64 accumulation per thread);
// zero_after: the same pass writes zeros back (the accumulator is ready for the next optimizer step).
__global__ void __launch_bounds__(kGradThreads)
grad_sumsq_kernel(float* __restrict__ x, size_t n, int zero_after, double* __restrict__ partial) {
  const size_t n4 = n >> 2;
  float4* x4 = reinterpret_cast<float4*>(x);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  double acc = 0.0;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = x4[i + u * stride];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) s += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
    acc += static_cast<double>(s);
    if (zero_after) {
#pragma unroll
      for (int u = 0; u < 4; ++u) x4[i + u * stride] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (; i < n4; i += stride) {
    const float4 v = x4[i];
    acc += static_cast<double>(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    if (zero_after) x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail elements when n is not a multiple of 4
    const size_t j = (n4 << 2) + threadIdx.x;
    acc += static_cast<double>(x[j] * x[j]);
    if (zero_after) x[j] = 0.f;
  }
  __shared__ double red[kGradThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = threadIdx.x < kGradThreads / 32 ? red[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}
// out[0] (+)= sum of the partials, in index order
__global__ void grad_sumsq_finalize_kernel(const double* __restrict__ partial, int n_partial, int accumulate,
                                           double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = accumulate ? out[0] : 0.0;
  for (int i = 0; i < n_partial; ++i) s += partial[i];
  out[0] = s;
}

// out[i] = bf16(x[i] * scale) (scale = scale_dev[0] when given, else scale_host); zero_after: x[i] = 0 in the same pass.
__global__ void __launch_bounds__(kGradThreads)
grad_scale_cast_kernel(float* __restrict__ x, size_t n, const float* __restrict__ scale_dev, float scale_host,
                       __nv_bfloat16* __restrict__ out, int zero_after) {
  const float sc = scale_dev ? scale_dev[0] : scale_host;
  const size_t n8 = n >> 3;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  float4* x4 = reinterpret_cast<float4*>(x);
  uint4* o4 = reinterpret_cast<uint4*>(out);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += stride) {
    const float4 a = x4[2 * i], b = x4[2 * i + 1];
    uint4 pk;
    pk.x = pack_bf16x2(a.x * sc, a.y * sc);
    pk.y = pack_bf16x2(a.z * sc, a.w * sc);
    pk.z = pack_bf16x2(b.x * sc, b.y * sc);
    pk.w = pack_bf16x2(b.z * sc, b.w * sc);
    o4[i] = pk;
    if (zero_after) {
      x4[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
      x4[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const size_t j = (n8 << 3) + threadIdx.x;
    out[j] = __float2bfloat16(x[j] * sc);
    if (zero_after) x[j] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// One-hot part of dW in a fixed summation order (option "deterministic"). scale_scatter_kernel adds
// onehot[r] * hidden[r][:] into dW[label_r][:] with fp32 atomics: rows of a chunk that share a label (about a thousand
// pairs per 18 944 rows of uniformly drawn labels) then add in completion order and the last bit of those dW rows can
// differ from run to run. Here the rows of a chunk are linked per label in row order (an O(rows^2 / 32) scan of the label
// vector, a few tens of microseconds from L1/L2) and the first row of every label sums its chain and makes the only
// read-modify-write of that dW row.
// ------------------------------------------------------------------------------------------
// next[r] = smallest r' > r with labels[r'] == labels[r] (or -1); has_prev[next[r]] = 1 (has_prev zeroed by the host)
__global__ void onehot_links_kernel(const int64_t* __restrict__ labels, uint32_t rows, uint32_t vocab,
                                    int32_t* __restrict__ next, int32_t* __restrict__ has_prev) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int64_t lab = labels[r];
  int32_t found = -1;
  if (lab >= 0 && lab < static_cast<int64_t>(vocab)) {
    for (uint32_t base = r + 1; base < rows; base += 32) {
      const uint32_t q = base + lane;
      const bool hit = q < rows && labels[q] == lab;
      const uint32_t m = __ballot_sync(0xffffffffu, hit);
      if (m != 0) {
        found = static_cast<int32_t>(base + __ffs(m) - 1);
        break;
      }
    }
  }
  if (lane == 0) {
    next[r] = found;
    if (found >= 0) has_prev[found] = 1;
  }
}
// the first row of every label walks its chain: dW[label][:] += sum_q onehot[q] * hidden[q][:], q in row order
__global__ void onehot_ordered_kernel(const __nv_bfloat16* __restrict__ hidden, const int64_t* __restrict__ labels,
                                      const float* __restrict__ onehot, const int32_t* __restrict__ next,
                                      const int32_t* __restrict__ has_prev, uint32_t rows, uint32_t hdim, uint32_t vocab,
                                      float* __restrict__ dweight) {
  const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (r >= rows || has_prev[r] != 0) return;
  const int64_t lab = labels[r];
  if (lab < 0 || lab >= static_cast<int64_t>(vocab)) return;
  float* dwp = dweight + static_cast<size_t>(lab) * hdim;
  for (uint32_t i = lane; i < (hdim >> 3); i += 32) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool any = false;
    for (int32_t q = static_cast<int32_t>(r); q >= 0; q = next[q]) {
      const float o = onehot[q];
      if (o == 0.f) continue;  // masked rows contribute nothing
      any = true;
      const uint4 a = reinterpret_cast<const uint4*>(hidden + static_cast<size_t>(q) * hdim)[i];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] += __uint_as_float(aw[k] << 16) * o;  // same products as scale_scatter_kernel (f * o), summed in row order
        acc[2 * k + 1] += __uint_as_float(aw[k] & 0xffff0000u) * o;
      }
    }
    if (!any) continue;
    float4* d = reinterpret_cast<float4*>(dwp + i * 8);
    float4 d0 = d[0], d1 = d[1];
    d0.x += acc[0]; d0.y += acc[1]; d0.z += acc[2]; d0.w += acc[3];
    d1.x += acc[4]; d1.y += acc[5]; d1.z += acc[6]; d1.w += acc[7];
    d[0] = d0;
    d[1] = d1;
  }
}

}  // namespace grpo
