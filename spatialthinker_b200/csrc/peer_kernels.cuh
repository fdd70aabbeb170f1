// Gradient exchange of the fp32 dW_lm_head accumulator over the GPUs of one NVSwitch box, through peer-mapped memory.
//
// Replaces: the gradient averaging FSDP does for the reference (mp_reduce_dtype = fp32, verl/workers/actor/config.py:58;
// verl/workers/fsdp_workers.py:242-280) and the optimizer-step passes that follow it (clip_grad_norm_ + the bf16 gradient
// the optimizer consumes, verl/workers/actor/dp_actor.py:155-167) - here for the replicated lm_head weight, once per
// optimizer step.
//
// An all-reduce moves 2 * (W-1)/W of the buffer over every GPU's links in each direction whoever issues it (3.8 GB at the
// 7B head; NCCL's ring already runs that at 73 % of the link rate, profiles/r2_nccl_probe_8gpu.log), so the gain is
// not in re-implementing it but in not sending what the optimizer step does not need in fp32:
//   1. peer_reduce_scatter_sumsq_kernel: rank r loads slab r of all W copies (W-1 of them over NVLink), sums them in rank
//      order in fp32, scales by 1/W, keeps the result in ITS OWN copy and adds up the squares of its slab; the last block
//      publishes that partial sum of squares to every rank. (W-1)/W of the buffer per direction.
//   2. host: global norm from the W partials (identical on all ranks, summed in rank order) -> clip coefficient.
//   3. peer_scale_cast_allgather_kernel: rank r scales its fp32 slab by the coefficient, rounds to bf16 - the dtype of the
//      parameter and therefore of the gradient the optimizer consumes - and stores it into all W bf16 gradient buffers;
//      the same pass zeroes the whole local fp32 accumulator for the next optimizer step. Half the bytes of an fp32
//      all-gather, and the clip / cast / zero passes over HBM disappear into it.
// Together 0.75x the wire bytes of the all-reduce and three HBM passes less. Every element is reduced by exactly one rank
// in a fixed order: bit-identical on all ranks and from run to run. Cross-rank ordering is a flag barrier
// (peer_barrier_kernel) before, between and after the passes, all stream-ordered: no host synchronisation.
#pragma once
#include "ptx.cuh"

namespace grpo {

constexpr int kMaxPeers = 8;
constexpr int kPeerBlocks = 148 * 4;
constexpr int kPeerThreads = 512;
struct PeerPtrs {
  void* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* ptr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* ptr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
// data loads from a peer's copy: system scope (served by the owner's L2, never by a stale line of this SM's L1)
__device__ __forceinline__ float4 ld_peer_f4(const float4* ptr) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr)
               : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t ns;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
  return ns;
}

// flags[q] of rank p = the last epoch rank q has announced to rank p. One block, one thread per peer: announce `epoch`
// to everyone (after a system-scope fence: everything this GPU wrote before - in earlier kernels of the stream - is
// visible to whoever sees the flag), then wait for everyone's announcement. A peer that never arrives traps the kernel
// after timeout_ns of wall-clock time instead of hanging the GPU for ever.
__global__ void peer_barrier_kernel(PeerPtrs flags, int rank, int world, uint32_t epoch, uint64_t timeout_ns) {
  const int q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();
  st_release_sys_u32(static_cast<uint32_t*>(flags.p[q]) + rank, epoch);
  const uint32_t* mine = static_cast<const uint32_t*>(flags.p[rank]) + q;
  const uint64_t t0 = global_timer_ns();
  while (static_cast<int32_t>(ld_acquire_sys_u32(mine) - epoch) < 0) {
    if (global_timer_ns() - t0 > timeout_ns) {
      printf("grpo: peer barrier timeout: rank %d waited %llu ms for rank %d (epoch %u)\n", rank,
             static_cast<unsigned long long>(timeout_ns / 1000000ull), q, epoch);
      __trap();
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

template <int W>
__device__ __forceinline__ float4 sum_copies_scaled(const float4 (&v)[W], float inv_world) {
  float4 s = v[0];
#pragma unroll
  for (int q = 1; q < W; ++q) {
    s.x += v[q].x;
    s.y += v[q].y;
    s.z += v[q].z;
    s.w += v[q].w;
  }
  s.x *= inv_world;
  s.y *= inv_world;
  s.z *= inv_world;
  s.w *= inv_world;
  return s;
}

// In-place mean all-reduce (general buffers): bufs.p[q] = rank q's copy of the [n] fp32 buffer; this rank reduces vectors
// [v0, v1) and writes the result into all copies.
template <int W>
__global__ void __launch_bounds__(kPeerThreads, 2)
peer_allreduce_mean_kernel(PeerPtrs bufs, size_t v0, size_t v1, float inv_world) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = v0 + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < v1; i += stride) {
    float4 v[W];
#pragma unroll
    for (int q = 0; q < W; ++q) v[q] = ld_peer_f4(static_cast<const float4*>(bufs.p[q]) + i);  // W loads in flight
    const float4 s = sum_copies_scaled<W>(v, inv_world);
#pragma unroll
    for (int q = 0; q < W; ++q) static_cast<float4*>(bufs.p[q])[i] = s;
  }
}

// Step 1. Vectors [v0, v1) = this rank's slab. mean -> own copy (own = bufs.p[rank]); sum of squares of the slab ->
// partial_out.p[q][rank] for every q (written by the last block to finish; `scratch` = kPeerBlocks doubles + one ticket
// word at scratch[kPeerBlocks], zero before the first launch and left zero).
template <int W>
__global__ void __launch_bounds__(kPeerThreads, 2)
peer_reduce_scatter_sumsq_kernel(PeerPtrs bufs, float4* __restrict__ own, int rank, size_t v0, size_t v1, float inv_world,
                                 PeerPtrs partial_out, double* __restrict__ scratch) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  double acc = 0.0;
  for (size_t i = v0 + blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < v1; i += stride) {
    float4 v[W];
#pragma unroll
    for (int q = 0; q < W; ++q) v[q] = ld_peer_f4(static_cast<const float4*>(bufs.p[q]) + i);
    const float4 s = sum_copies_scaled<W>(v, inv_world);
    own[i] = s;
    acc += static_cast<double>(s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w);
  }
  __shared__ double red[kPeerThreads / 32];
  __shared__ bool last;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = threadIdx.x < kPeerThreads / 32 ? red[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      scratch[blockIdx.x] = s;
      __threadfence();
      unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + gridDim.x);
      last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double total = 0.0;
  for (unsigned int b = 0; b < gridDim.x; ++b) total += *(volatile double*)(scratch + b);  // block order: reproducible
#pragma unroll
  for (int q = 0; q < W; ++q) static_cast<double*>(partial_out.p[q])[rank] = total;
  *reinterpret_cast<unsigned int*>(scratch + gridDim.x) = 0u;
}

// Step 3. Units of 8 elements; [u0, u1) = this rank's slab (the same elements as step 1's vectors [2 u0, 2 u1)).
// outs.p[q][slab] = bf16(grad[slab] * scale) for every q; zero_after: the WHOLE local accumulator is zeroed (every peer
// has finished reading it: the barrier after step 1).
template <int W>
__global__ void __launch_bounds__(kPeerThreads)
peer_scale_cast_allgather_kernel(float* __restrict__ grad, size_t units, size_t u0, size_t u1,
                                 const float* __restrict__ scale_dev, float scale_host, PeerPtrs outs, int zero_after) {
  const float sc = scale_dev ? scale_dev[0] : scale_host;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t first = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  float4* g4 = reinterpret_cast<float4*>(grad);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = u0 + first; i < u1; i += stride) {
    const float4 a = g4[2 * i], b = g4[2 * i + 1];
    uint4 pk;
    pk.x = pack_bf16x2(a.x * sc, a.y * sc);
    pk.y = pack_bf16x2(a.z * sc, a.w * sc);
    pk.z = pack_bf16x2(b.x * sc, b.y * sc);
    pk.w = pack_bf16x2(b.z * sc, b.w * sc);
#pragma unroll
    for (int q = 0; q < W; ++q) static_cast<uint4*>(outs.p[q])[i] = pk;
    if (zero_after) {
      g4[2 * i] = zero;
      g4[2 * i + 1] = zero;
    }
  }
  if (!zero_after) return;
  for (size_t i = first; i < units; i += stride) {  // the other ranks' slabs of the local accumulator
    if (i >= u0 && i < u1) continue;
    g4[2 * i] = zero;
    g4[2 * i + 1] = zero;
  }
}

}  // namespace grpo
