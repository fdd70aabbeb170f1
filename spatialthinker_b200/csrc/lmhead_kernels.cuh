// lm_head kernels: the softmax-statistics epilogue of the logits GEMM, the per-row combine, and the in-place
// transform of the stashed exp tile values into dL/dlogits.
//
// Reference path being replaced (logits fully materialised there):
//   verl/workers/actor/dp_actor.py:125-128   logits = lm_head(hidden); logits.div_(temperature); log_probs_from_logits
//   verl/utils/torch_functional.py:45-66     log p[label] = z[label] - logsumexp(z)
#pragma once
#include "gemm_core.cuh"

namespace grpo {

constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------
// Forward epilogue. One thread owns one token row of the 128 x BLOCK_N accumulator tile:
//   pass 1: tile max of the raw accumulator
//   pass 2: e = exp(z - tile max) -> running sum, sum(e*z) (for the entropy), optional bf16 stash of e, and capture of
//           the label logit if the label column falls in this tile
// Per-(vocab tile, row) partials go to a [tiles][rows_pad] workspace; lse/entropy are finished by combine_rows_kernel.
// ------------------------------------------------------------------------------------------
template <int kCta, int BLOCK_N>
struct EpiSoftmax {
  struct Params {
    uint32_t rows;       // token rows in this launch
    uint32_t vocab;      // V
    uint32_t rows_pad;   // leading dimension of the partial arrays
    float scale;         // 1 / temperature
    float* part_max;     // [n_tiles][rows_pad]  max_v z           (z = accumulator * scale)
    float* part_sum;     // [n_tiles][rows_pad]  sum_v exp(z - max)
    float* part_ez;      // [n_tiles][rows_pad]  sum_v exp(z - max) * z      (nullptr: skip)
    const int64_t* labels;  // [rows]
    float* target_z;        // [rows] z at the label column
    __nv_bfloat16* stash;   // [rows][ld_stash] exp(z - tile max) in bf16    (nullptr: forward only)
    int64_t ld_stash;
  };
  static constexpr int kSmemBytes = 0;

  __device__ static void run(const Params& p, const EpiCtx& c, uint8_t*) {
    const uint32_t row = c.m_blk * (kBlockM * kCta) + c.row_in_tile;
    const bool row_ok = row < p.rows;
    const uint32_t col0 = c.n_blk * BLOCK_N;
    const uint32_t ncols = min(static_cast<uint32_t>(BLOCK_N), p.vocab - col0);
    const uint32_t ngroups = (ncols + 31) >> 5;

    // ---- pass 1: tile max
    float mx = -INFINITY;
#pragma unroll 1
    for (uint32_t g = 0; g < ngroups; ++g) {
      uint32_t v[32];
      tmem_ld_32x32(c.tmem_acc + g * 32, v);
      tmem_ld_wait();
      const uint32_t valid = ncols - g * 32;  // >= 1
      if (valid >= 32) {
        float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]),
              m3 = __uint_as_float(v[3]);
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i]));
          m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (static_cast<uint32_t>(i) < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
    }
    // scale > 0, so max commutes with the temperature scaling
    const float c1 = p.scale * kLog2e;
    const float off = mx * c1;

    int32_t tl = -1;  // label column relative to this tile, if it lands here
    if (row_ok) {
      const int64_t lab = p.labels[row] - static_cast<int64_t>(col0);
      if (lab >= 0 && lab < static_cast<int64_t>(ncols)) tl = static_cast<int32_t>(lab);
    }
    const bool want_ez = p.part_ez != nullptr;
    const bool want_stash = p.stash != nullptr;
    __nv_bfloat16* srow = want_stash ? p.stash + static_cast<int64_t>(row) * p.ld_stash + col0 : nullptr;

    // ---- pass 2: exp, sums, stash, label capture
    float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f, zt = 0.f;
#pragma unroll 1
    for (uint32_t g = 0; g < ngroups; ++g) {
      uint32_t v[32];
      tmem_ld_32x32(c.tmem_acc + g * 32, v);
      tmem_ld_wait();
      const uint32_t valid = ncols - g * 32;
      float e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float a = __uint_as_float(v[i]);
        float ex = fast_exp2(fmaf(a, c1, -off));
        if (valid < 32 && static_cast<uint32_t>(i) >= valid) ex = 0.f;
        e[i] = ex;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        s0 += e[i];
        s1 += e[i + 1];
      }
      if (want_ez) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          t0 = fmaf(e[i], __uint_as_float(v[i]), t0);
          t1 = fmaf(e[i + 1], __uint_as_float(v[i + 1]), t1);
        }
      }
      if (__any_sync(0xffffffffu, (tl >> 5) == static_cast<int32_t>(g) && tl >= 0)) {
        const int32_t j = tl - static_cast<int32_t>(g * 32);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (j == i) zt = __uint_as_float(v[i]);
      }
      if (want_stash && row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (static_cast<uint32_t>(q * 8) < valid) {  // vocab % 8 == 0
            uint4 pk;
            pk.x = pack_bf16x2(e[8 * q + 0], e[8 * q + 1]);
            pk.y = pack_bf16x2(e[8 * q + 2], e[8 * q + 3]);
            pk.z = pack_bf16x2(e[8 * q + 4], e[8 * q + 5]);
            pk.w = pack_bf16x2(e[8 * q + 6], e[8 * q + 7]);
            *reinterpret_cast<uint4*>(srow + g * 32 + q * 8) = pk;
          }
        }
      }
    }
    if (row_ok) {
      const size_t o = static_cast<size_t>(c.n_blk) * p.rows_pad + row;
      p.part_max[o] = mx * p.scale;
      p.part_sum[o] = s0 + s1;
      if (want_ez) p.part_ez[o] = (t0 + t1) * p.scale;
      if (tl >= 0) p.target_z[row] = zt * p.scale;
    }
  }
};

// ------------------------------------------------------------------------------------------
// Per-row combine of the vocab-tile partials: lse, log p[label], entropy = lse - sum p z.
// ------------------------------------------------------------------------------------------
__global__ void combine_rows_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                    const float* __restrict__ part_ez, const float* __restrict__ target_z,
                                    const int64_t* __restrict__ labels, uint32_t rows, uint32_t rows_pad,
                                    uint32_t n_tiles, uint32_t vocab, float* __restrict__ lse_out,
                                    float* __restrict__ logp_out, float* __restrict__ ent_out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m = -INFINITY;
  for (uint32_t j = 0; j < n_tiles; ++j) m = fmaxf(m, part_max[static_cast<size_t>(j) * rows_pad + r]);
  float s = 0.f, ez = 0.f;
  for (uint32_t j = 0; j < n_tiles; ++j) {
    const size_t o = static_cast<size_t>(j) * rows_pad + r;
    const float w = __expf(part_max[o] - m);
    s = fmaf(part_sum[o], w, s);
    if (part_ez) ez = fmaf(part_ez[o], w, ez);
  }
  const float lse = m + logf(s);
  if (lse_out) lse_out[r] = lse;
  if (logp_out) {  // a label outside [0, vocab) (e.g. an ignore index) yields 0, like the cross-entropy it replaces
    const int64_t lab = labels[r];
    logp_out[r] = (lab >= 0 && lab < static_cast<int64_t>(vocab)) ? target_z[r] - lse : 0.f;
  }
  if (ent_out) ent_out[r] = lse - ez / s;
}

// ------------------------------------------------------------------------------------------
// stash (exp(z - tile max), bf16)  ->  dL/dz in place (bf16):
//    p[r][v]  = exp(tile max - lse r) * e[r][v]
//    dz[r][v] = scale * ( dlogp[r] * (1[v == label r] - p) - dent[r] * p * (log p + H r) )
// (the 1/temperature of z = h.W / T is folded in here, so the two backward GEMMs are plain products; the dent term is
// the gradient of a per-token entropy output, H = lse - sum p z, and is skipped when dent == nullptr).
// Rows whose upstream gradients are all zero (masked tokens) are written as zeros without reading the stash.
// ------------------------------------------------------------------------------------------
template <int BLOCK_N>
__global__ void stash_to_dlogits_kernel(__nv_bfloat16* __restrict__ stash, int64_t ld_stash, uint32_t rows,
                                        uint32_t vocab, const float* __restrict__ part_max, uint32_t rows_pad,
                                        const float* __restrict__ lse, const float* __restrict__ dlogp,
                                        const float* __restrict__ dent, const float* __restrict__ ent,
                                        const int64_t* __restrict__ labels, float scale) {
  const uint32_t vecs_per_row = vocab >> 3;
  for (uint32_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const float g = dlogp[r] * scale;
    const float ge = dent ? dent[r] * scale : 0.f;
    const float h = dent ? ent[r] : 0.f;
    const float l = lse[r];
    const int64_t lab = labels[r];
    uint4* rowp = reinterpret_cast<uint4*>(stash + static_cast<int64_t>(r) * ld_stash);
    for (uint32_t vi = blockIdx.x * blockDim.x + threadIdx.x; vi < vecs_per_row; vi += gridDim.x * blockDim.x) {
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (g != 0.f || ge != 0.f) {
        const uint32_t col = vi << 3;
        const float shift = part_max[static_cast<size_t>(col / BLOCK_N) * rows_pad + r] - l;  // log of the tile scale
        const float cexp = __expf(shift);
        const uint4 in = rowp[vi];
        const uint32_t w[4] = {in.x, in.y, in.z, in.w};
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f[2 * i] = __uint_as_float(w[i] << 16);
          f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e = f[i];
          const float p = cexp * e;
          float d = -g * p;
          if (ge != 0.f && e > 0.f) d -= ge * p * (shift + __logf(e) + h);
          f[i] = d;
        }
        const int64_t d = lab - static_cast<int64_t>(col);
        if (d >= 0 && d < 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (d == i) f[i] += g;
        }
        out.x = pack_bf16x2(f[0], f[1]);
        out.y = pack_bf16x2(f[2], f[3]);
        out.z = pack_bf16x2(f[4], f[5]);
        out.w = pack_bf16x2(f[6], f[7]);
      }
      rowp[vi] = out;
    }
  }
}

}  // namespace grpo
