// lm_head kernels: the label-logit reference, the softmax-statistics epilogue of the logits GEMM, the per-row combine,
// and the two ways of turning the stashed exponentials into the backward GEMMs' operands.
//
// Reference path being replaced (logits fully materialised there):
//   verl/workers/actor/dp_actor.py:125-128   logits = lm_head(hidden); logits.div_(temperature); log_probs_from_logits
//   verl/utils/torch_functional.py:45-66     log p[label] = z[label] - logsumexp(z)
#pragma once
#include "gemm_core.cuh"

namespace grpo {

constexpr float kLog2e = 1.4426950408889634f;
// exp2 argument clamp of the softmax epilogue: keeps every stashed value finite (2^100 ~ 1.3e30, row sums < 2e35) so that
// masked rows can never inject inf * 0 = NaN into the backward GEMMs. It only binds when some logit exceeds the label's
// logit by more than 69 nats, i.e. for a label whose probability is below e^-69 (1e-30); such a row's lse saturates.
constexpr float kClampLog2 = 100.f;

// ------------------------------------------------------------------------------------------
// Per-row reference for the softmax: the label's own logit a_label = h_r . W[label_r] (raw accumulator units, before
// the 1/temperature scale), one warp per row. With it the forward epilogue is a single pass, exp(z - z_label) needs no
// running max, log p[label] = -log sum_v exp(z_v - z_label), and the stash differs from the softmax only by a per-row
// factor - which the backward GEMMs apply in their epilogue / operand instead of a separate pass over the stash.
// ------------------------------------------------------------------------------------------
__global__ void label_dot_kernel(const __nv_bfloat16* __restrict__ hidden, const __nv_bfloat16* __restrict__ weight,
                                 const int64_t* __restrict__ labels, uint32_t rows, uint32_t hdim, uint32_t vocab,
                                 float* __restrict__ a_label) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t lab = labels[row];
  float acc = 0.f;
  if (lab >= 0 && lab < static_cast<int64_t>(vocab)) {
    const uint4* hp = reinterpret_cast<const uint4*>(hidden + static_cast<size_t>(row) * hdim);
    const uint4* wp = reinterpret_cast<const uint4*>(weight + static_cast<size_t>(lab) * hdim);
    for (uint32_t i = lane; i < (hdim >> 3); i += 32) {
      const uint4 a = hp[i], b = wp[i];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc = fmaf(__uint_as_float(aw[k] << 16), __uint_as_float(bw[k] << 16), acc);
        acc = fmaf(__uint_as_float(aw[k] & 0xffff0000u), __uint_as_float(bw[k] & 0xffff0000u), acc);
      }
    }
    acc = warp_sum(acc);
  }
  if (lane == 0) a_label[row] = acc;  // 0 for an out-of-range (ignored) label
}

// ------------------------------------------------------------------------------------------
// Forward epilogue. One thread owns one token row of the 128 x BLOCK_N accumulator tile and makes ONE pass over it:
//   e = exp((a - a_label) / T)  ->  sum e, sum e*z (for the entropy), optional bf16 stash of e.
// Per-(vocab tile, row) partial sums go to a [tiles][rows_pad] workspace and are finished by combine_rows_kernel.
// ------------------------------------------------------------------------------------------
template <int kCta, int BLOCK_N>
struct EpiSoftmax {
  struct alignas(64) Params {
    CUtensorMap stash_map;  // store map over the blocked stash, box (64 cols, 32 rows, 1, 1), SWIZZLE_128B; iff mode & 2
    CUtensorMap stash_map_half;  // the same with box (32 cols, 32 rows, 1, 1), SWIZZLE_64B; iff mode & 4
    uint32_t rows;       // token rows in this launch
    uint32_t vocab;      // V
    uint32_t rows_pad;   // leading dimension of the partial arrays
    float scale;         // 1 / temperature
    const float* ref;    // [rows] a_label (label_dot_kernel)
    float* part_sum;     // [n_tiles][rows_pad]  sum_v exp(z - z_label)
    float* part_ez;      // [n_tiles][rows_pad]  sum_v exp(z - z_label) * z      (nullptr: skip)
    __nv_bfloat16* stash;   // blocked [rows/64][stash_vb][64][64]: exp(z - z_label) in bf16 (nullptr: forward only)
    uint32_t stash_vb;      // 64-column blocks per row block
    // bit 0: full tiles drain TMEM software-pipelined (the load of column group g+1 is in flight while g is processed)
    // bit 1: (with bit 0) the stash leaves through shared memory and bulk tensor stores, 4 KB per warp and 64 columns,
    //        instead of 16-byte st.global.cs pieces in 32 different 128-byte lines per warp instruction
    // bit 2: (with bits 0-1) the warp's 4 KB staging buffer is used as two 2 KB halves, one bulk store per 32-column
    //        group: a half is only rewritten two groups later, so the warp never waits on the store it just issued
    uint32_t mode;
    uint64_t policy;  // L2 eviction priority of the stash stores: it is next read by another kernel, after 5.8 GB went by
  };
  static constexpr int kSmemBytes = 8 * kEpiStageBytes;
  static constexpr bool kCanShare = true;  // a call can cover BLOCK_N / 2 columns of an accumulator (TileSched::epi_share)
  __device__ static void finish(const Params& p, uint32_t lane) {
    if ((p.mode & 2) && lane == 0) bulk_wait_all();
  }

  // One 32-column group of a full tile: exponentials, running sums, bf16 pack; the packed row piece goes either to the
  // warp's staging buffer (swizzled 16-byte chunks `chunk0 .. chunk0+3` of this thread's 128-byte row) or to global.
  template <bool kWantEz>
  __device__ static __forceinline__ void group_full(const uint32_t (&v)[32], float c1, float off, float& s0, float& s1,
                                                    float& t0, float& t1, bool want_stash, uint8_t* smem_row,
                                                    uint32_t sw, uint32_t chunk0, __nv_bfloat16* gdst) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = fast_exp2(fminf(fmaf(__uint_as_float(v[8 * q + i]), c1, -off), kClampLog2));
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        s0 += e[i];
        s1 += e[i + 1];
      }
      if (kWantEz) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          t0 = fmaf(e[i], __uint_as_float(v[8 * q + i]), t0);
          t1 = fmaf(e[i + 1], __uint_as_float(v[8 * q + i + 1]), t1);
        }
      }
      if (want_stash) {
        const uint32_t x = pack_bf16x2(e[0], e[1]), y = pack_bf16x2(e[2], e[3]);
        const uint32_t z = pack_bf16x2(e[4], e[5]), w = pack_bf16x2(e[6], e[7]);
        if (smem_row) st_shared_v4(smem_row + (((chunk0 + q) ^ sw) << 4), x, y, z, w);
        else __stcs(reinterpret_cast<uint4*>(gdst + q * 8), make_uint4(x, y, z, w));
      }
    }
  }

  template <bool kWantEz, int kCols, class Release>
  __device__ static __forceinline__ void run_full(const Params& p, const EpiCtx& c, uint8_t* smem_epi, Release&& release,
                                                  float c1, float off, __nv_bfloat16* srow, float& s0, float& s1,
                                                  float& t0, float& t1) {
    const bool want_stash = p.stash != nullptr;
    const bool via_tma = want_stash && (p.mode & 2);
    uint8_t* buf = smem_epi + c.epi_warp * kEpiStageBytes;
    uint8_t* smem_row = via_tma ? buf + c.lane * 128 : nullptr;
    const uint32_t sw = c.lane & 7;
    const int32_t rb = static_cast<int32_t>(c.row0 >> 6), r_in = static_cast<int32_t>(c.row0 & 63);
    const int32_t vb0 = static_cast<int32_t>(c.col0 >> 6);
    uint32_t va[32], vb[32];
    static_assert((kCols / 32) % 2 == 0, "column groups are drained in pairs (one 64-column stash block)");
    tmem_ld_32x32(c.tmem_acc, va);
#pragma unroll 1
    for (int g = 0; g < kCols / 32; g += 2) {
      tmem_ld_wait();
      tmem_ld_32x32(c.tmem_acc + (g + 1) * 32, vb);
      if (via_tma) {  // the previous 64-column block must have left the staging buffer
        if (c.lane == 0) bulk_wait_read_all();
        __syncwarp();
      }
      __nv_bfloat16* gdst = srow + (g >> 1) * 4096;
      group_full<kWantEz>(va, c1, off, s0, s1, t0, t1, want_stash, smem_row, sw, 0, gdst);
      tmem_ld_wait();
      if (g + 2 < kCols / 32) tmem_ld_32x32(c.tmem_acc + (g + 2) * 32, va);
      else release();  // last TMEM read of this accumulator has landed
      group_full<kWantEz>(vb, c1, off, s0, s1, t0, t1, want_stash, smem_row, sw, 4, gdst + 32);
      if (via_tma) {
        fence_proxy_async_smem();
        __syncwarp();
        if (c.lane == 0) {
          tma_store_4d(&p.stash_map, buf, 0, r_in, vb0 + (g >> 1), rb, p.policy);
          bulk_commit();
        }
      }
    }
  }

  // exponentials, running sums and bf16 pack of one 32-column group of a full tile (registers only)
  template <bool kWantEz>
  __device__ static __forceinline__ void group_pack(const uint32_t (&v)[32], float c1, float off, float& s0, float& s1,
                                                    float& t0, float& t1, uint32_t (&pk)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float e[8];
#pragma unroll
#if defined(GRPO_EPI_DBG) && (GRPO_EPI_DBG & 1)  // timing experiment only: no MUFU (results are wrong)
      for (int i = 0; i < 8; ++i) e[i] = fminf(fmaf(__uint_as_float(v[8 * q + i]), c1, -off), kClampLog2);
#else
      for (int i = 0; i < 8; ++i) e[i] = fast_exp2(fminf(fmaf(__uint_as_float(v[8 * q + i]), c1, -off), kClampLog2));
#endif
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        s0 += e[i];
        s1 += e[i + 1];
      }
      if (kWantEz) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          t0 = fmaf(e[i], __uint_as_float(v[8 * q + i]), t0);
          t1 = fmaf(e[i + 1], __uint_as_float(v[8 * q + i + 1]), t1);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) pk[4 * q + i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
    }
  }

  // mode bits 0-2: pipelined TMEM drain, one bulk store per 32-column group out of alternating 2 KB staging halves
  template <bool kWantEz, int kCols, class Release>
  __device__ static __forceinline__ void run_half(const Params& p, const EpiCtx& c, uint8_t* smem_epi, Release&& release,
                                                  float c1, float off, float& s0, float& s1, float& t0, float& t1) {
    uint8_t* buf = smem_epi + c.epi_warp * kEpiStageBytes;
    uint8_t* smem_row = buf + c.lane * 64;          // 32 rows x 64 bytes per half
    const uint32_t sw = (c.lane >> 1) & 3;          // SWIZZLE_64B: 16-byte chunk index ^= bits 7-8 of the address
    const int32_t rb = static_cast<int32_t>(c.row0 >> 6), r_in = static_cast<int32_t>(c.row0 & 63);
    const int32_t vb0 = static_cast<int32_t>(c.col0 >> 6);
    uint32_t va[32], vb[32];
    auto emit = [&](const uint32_t (&v)[32], int g) {
      uint32_t pk[16];
      group_pack<kWantEz>(v, c1, off, s0, s1, t0, t1, pk);
      uint8_t* half = buf + (g & 1) * 2048;
#if defined(GRPO_EPI_DBG) && (GRPO_EPI_DBG & 2)  // timing experiment only: nothing leaves the registers
      if (pk[0] == 0x12345678u && pk[7] == 0x9abcdef0u) s0 += 1.f;
      return;
#endif
      if (c.lane == 0) bulk_wait_read_1();  // the store issued two groups ago (same half) has left shared memory
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4(smem_row + (g & 1) * 2048 + ((q ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (c.lane == 0) {
        tma_store_4d(&p.stash_map_half, half, (g & 1) * 32, r_in, vb0 + (g >> 1), rb, p.policy);
        bulk_commit();
      }
    };
    tmem_ld_32x32(c.tmem_acc, va);
#pragma unroll 1
    for (int g = 0; g < kCols / 32; g += 2) {
      tmem_ld_wait();
      tmem_ld_32x32(c.tmem_acc + (g + 1) * 32, vb);
      emit(va, g);
      tmem_ld_wait();
      if (g + 2 < kCols / 32) tmem_ld_32x32(c.tmem_acc + (g + 2) * 32, va);
      else release();  // last TMEM read of this accumulator has landed
      emit(vb, g + 1);
    }
  }

  template <class Release>
  __device__ static void run(const Params& p, const EpiCtx& c, uint8_t* smem_epi, Release&& release) {
    run_cols<BLOCK_N>(p, c, smem_epi, release);
  }

  // kCols columns of the accumulator starting at global column c.col0 / TMEM address c.tmem_acc
  template <int kCols, class Release>
  __device__ static void run_cols(const Params& p, const EpiCtx& c, uint8_t* smem_epi, Release&& release) {
    const uint32_t row = c.row, col0 = c.col0;
    const bool row_ok = row < p.rows;
    const uint32_t ncols = col0 < p.vocab ? min(static_cast<uint32_t>(kCols), p.vocab - col0) : 0u;
    const uint32_t ngroups = (ncols + 31) >> 5;  // column groups of 32 that hold at least one real vocabulary entry
    const float c1 = p.scale * kLog2e;
    // rows past the last token get an infinite offset: every exponential is then exactly 0 (zeros in the stash) with no
    // per-element select
    const float off = row_ok ? p.ref[row] * c1 : INFINITY;
    const bool want_ez = p.part_ez != nullptr;
    // Every row of the tile is stored (zeros past the last token / past the vocabulary) so that the blocked stash
    // never exposes stale bytes to the backward GEMMs' zero-padded K tails.
    const bool want_stash = p.stash != nullptr;
    __nv_bfloat16* srow = want_stash
        ? p.stash + (static_cast<size_t>(row >> 6) * p.stash_vb + (col0 >> 6)) * 4096 + (row & 63) * 64
        : nullptr;
    const uint32_t vb_here = col0 >> 6;  // first 64-column stash block of this call (may lie past the stash's last block)
    const uint32_t store_groups =
        (want_stash && vb_here < p.stash_vb) ? min(static_cast<uint32_t>(kCols / 32), 2 * (p.stash_vb - vb_here)) : 0;

    float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
    if ((p.mode & 7) == 7 && want_stash && ncols == kCols) {
      if (want_ez) run_half<true, kCols>(p, c, smem_epi, release, c1, off, s0, s1, t0, t1);
      else run_half<false, kCols>(p, c, smem_epi, release, c1, off, s0, s1, t0, t1);
    } else if ((p.mode & 1) && ncols == kCols) {  // every tile but the last one of the vocabulary
      if (want_ez) run_full<true, kCols>(p, c, smem_epi, release, c1, off, srow, s0, s1, t0, t1);
      else run_full<false, kCols>(p, c, smem_epi, release, c1, off, srow, s0, s1, t0, t1);
    } else {
      if (ngroups == 0) release();  // nothing to read: columns wholly past the vocabulary
#pragma unroll 1
      for (uint32_t g = 0; g < ngroups; ++g) {
        uint32_t v[32];
        tmem_ld_32x32(c.tmem_acc + g * 32, v);
        tmem_ld_wait();
        if (g + 1 == ngroups) release();  // last TMEM read of this accumulator
        const uint32_t valid = ncols - g * 32;  // >= 1
        float e[32];
        if (valid >= 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) e[i] = fast_exp2(fminf(fmaf(__uint_as_float(v[i]), c1, -off), kClampLog2));
        } else {  // only the last group of the last vocabulary tile: zero-filled columns past the vocabulary
#pragma unroll
          for (int i = 0; i < 32; ++i)
            e[i] = static_cast<uint32_t>(i) < valid ? fast_exp2(fminf(fmaf(__uint_as_float(v[i]), c1, -off), kClampLog2)) : 0.f;
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
        if (want_stash) {
          __nv_bfloat16* dst = srow + (g >> 1) * 4096 + (g & 1) * 32;  // 64-column block g/2, half g%2 of its 128-byte row
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk;
            pk.x = pack_bf16x2(e[8 * q + 0], e[8 * q + 1]);
            pk.y = pack_bf16x2(e[8 * q + 2], e[8 * q + 3]);
            pk.z = pack_bf16x2(e[8 * q + 4], e[8 * q + 5]);
            pk.w = pack_bf16x2(e[8 * q + 6], e[8 * q + 7]);
            __stcs(reinterpret_cast<uint4*>(dst + q * 8), pk);  // streamed: keep the hidden panel in L2
          }
        }
      }
      // column groups of the last vocabulary block that lie wholly past the vocabulary: zeros
      for (uint32_t g = ngroups; g < store_groups; ++g) {
        __nv_bfloat16* dst = srow + (g >> 1) * 4096 + (g & 1) * 32;
#pragma unroll
        for (int q = 0; q < 4; ++q) __stcs(reinterpret_cast<uint4*>(dst + q * 8), make_uint4(0u, 0u, 0u, 0u));
      }
    }
    if (row_ok) {
      const size_t o = static_cast<size_t>(c.part) * p.rows_pad + row;
      p.part_sum[o] = s0 + s1;
      if (want_ez) p.part_ez[o] = (t0 + t1) * p.scale;
    }
  }
};

// ------------------------------------------------------------------------------------------
// Per-row combine of the vocab-tile partial sums. blockDim = (32 rows, 8 tile slices).
//   S = sum_v exp(z_v - z_label);  lse = z_label + log S;  log p[label] = -log S;  entropy = lse - (sum e z) / S
// ------------------------------------------------------------------------------------------
__global__ void combine_rows_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_ez,
                                    const float* __restrict__ a_label, const int64_t* __restrict__ labels,
                                    uint32_t rows, uint32_t rows_pad, uint32_t n_tiles, uint32_t vocab, float scale,
                                    float* __restrict__ lse_out, float* __restrict__ logp_out,
                                    float* __restrict__ ent_out, float* __restrict__ inv_sum_out) {
  __shared__ float red_s[8][33], red_t[8][33];
  const uint32_t r = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f, t = 0.f;
  if (r < rows) {
    for (uint32_t j = threadIdx.y; j < n_tiles; j += 8) {
      const size_t o = static_cast<size_t>(j) * rows_pad + r;
      s += part_sum[o];
      if (part_ez) t += part_ez[o];
    }
  }
  red_s[threadIdx.y][threadIdx.x] = s;
  red_t[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y != 0 || r >= rows) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    s += red_s[k][threadIdx.x];
    t += red_t[k][threadIdx.x];
  }
  const float log_s = logf(s);
  const float lse = a_label[r] * scale + log_s;
  if (lse_out) lse_out[r] = lse;
  if (logp_out) {  // a label outside [0, vocab) (e.g. an ignore index) yields 0, like the cross-entropy it replaces
    const int64_t lab = labels[r];
    logp_out[r] = (lab >= 0 && lab < static_cast<int64_t>(vocab)) ? -log_s : 0.f;
  }
  if (ent_out) ent_out[r] = lse - t / s;
  if (inv_sum_out) inv_sum_out[r] = 1.f / s;
}

// ------------------------------------------------------------------------------------------
// Fast backward preparation (no entropy gradient). With p = e / S the gradient of the logits factorises per row:
//     dz[r][v] = (g_r / T) * (1[v == label_r] - e[r][v] / S_r)
// so the stash is used AS IS by both backward GEMMs and only O(rows * H) work remains, done here, one warp per row:
//     row_scale[r] = -g_r / (T S_r)          -> applied by the dHidden GEMM epilogue to (e . W)[r][:]
//     onehot[r]    =  g_r / T                -> dHidden epilogue adds onehot[r] * W[label_r][:]
//     hd_scaled[r] = bf16(row_scale[r] * hidden[r][:])      -> B operand of the dW GEMM:  dW += e^T . hd_scaled
//     dW[label_r][:] += onehot[r] * hidden[r][:]            -> the one-hot part of dW (fp32 red.add)
// ------------------------------------------------------------------------------------------
__global__ void scale_scatter_kernel(const __nv_bfloat16* __restrict__ hidden, const int64_t* __restrict__ labels,
                                     const float* __restrict__ dlogp, const float* __restrict__ inv_sum, float scale,
                                     uint32_t rows, uint32_t hdim, uint32_t vocab, float* __restrict__ row_scale,
                                     float* __restrict__ onehot, __nv_bfloat16* __restrict__ hd_scaled,
                                     float* __restrict__ dweight) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float o = dlogp[row] * scale;
  const float sc = -o * inv_sum[row];
  const int64_t lab = labels[row];
  const bool lab_ok = lab >= 0 && lab < static_cast<int64_t>(vocab);
  if (lane == 0) {
    row_scale[row] = sc;
    onehot[row] = lab_ok ? o : 0.f;
  }
  const uint4* hp = reinterpret_cast<const uint4*>(hidden + static_cast<size_t>(row) * hdim);
  uint4* op = reinterpret_cast<uint4*>(hd_scaled + static_cast<size_t>(row) * hdim);
  float* dwp = (lab_ok && o != 0.f) ? dweight + static_cast<size_t>(lab) * hdim : nullptr;
  for (uint32_t i = lane; i < (hdim >> 3); i += 32) {
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (sc != 0.f) {
      const uint4 a = hp[i];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      float f[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(aw[k] << 16);
        f[2 * k + 1] = __uint_as_float(aw[k] & 0xffff0000u);
      }
      out.x = pack_bf16x2(f[0] * sc, f[1] * sc);
      out.y = pack_bf16x2(f[2] * sc, f[3] * sc);
      out.z = pack_bf16x2(f[4] * sc, f[5] * sc);
      out.w = pack_bf16x2(f[6] * sc, f[7] * sc);
      if (dwp) {
        atomicAdd(reinterpret_cast<float4*>(dwp + i * 8), make_float4(f[0] * o, f[1] * o, f[2] * o, f[3] * o));
        atomicAdd(reinterpret_cast<float4*>(dwp + i * 8 + 4), make_float4(f[4] * o, f[5] * o, f[6] * o, f[7] * o));
      }
    }
    op[i] = out;  // masked rows (g == 0) contribute an all-zero operand row
  }
}

// ------------------------------------------------------------------------------------------
// Fix-up pass of the dHidden GEMM's split-K path: the K slices were summed in fp32 by the reduce-add epilogue; this
// applies what the direct bf16 epilogue (EpiBF16) does to a finished accumulator,
//     dhidden[r][:] = bf16(row_scale[r] * acc[r][:] + onehot[r] * W[label_r][:])
// (row_scale == nullptr: 1, onehot == nullptr: no gather term - the entropy-gradient mode). One thread per 8 columns.
// ------------------------------------------------------------------------------------------
__global__ void dh_fixup_kernel(const float* __restrict__ acc, const float* __restrict__ row_scale,
                                const float* __restrict__ onehot, const int64_t* __restrict__ labels,
                                const __nv_bfloat16* __restrict__ weight, uint32_t rows, uint32_t hdim,
                                __nv_bfloat16* __restrict__ dhidden) {
  const uint32_t vpr = hdim >> 3;  // 8-column vectors per row (hdim % 64 == 0)
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t row = idx / vpr;
  if (row >= rows) return;
  const uint32_t col = (idx - row * vpr) << 3;
  const size_t off = static_cast<size_t>(row) * hdim + col;
  const float4 a0 = *reinterpret_cast<const float4*>(acc + off);
  const float4 a1 = *reinterpret_cast<const float4*>(acc + off + 4);
  const float sc = row_scale ? row_scale[row] : 1.f;
  const float oh = onehot ? onehot[row] : 0.f;
  float f[8] = {a0.x * sc, a0.y * sc, a0.z * sc, a0.w * sc, a1.x * sc, a1.y * sc, a1.z * sc, a1.w * sc};
  if (oh != 0.f) {  // onehot is 0 for rows whose label is out of range (scale_scatter_kernel)
    const uint4 wv = *reinterpret_cast<const uint4*>(weight + static_cast<size_t>(labels[row]) * hdim + col);
    const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = fmaf(oh, __uint_as_float(ww[i] << 16), f[2 * i]);
      f[2 * i + 1] = fmaf(oh, __uint_as_float(ww[i] & 0xffff0000u), f[2 * i + 1]);
    }
  }
  uint4 pk;
  pk.x = pack_bf16x2(f[0], f[1]);
  pk.y = pack_bf16x2(f[2], f[3]);
  pk.z = pack_bf16x2(f[4], f[5]);
  pk.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(dhidden + off) = pk;
}

// ------------------------------------------------------------------------------------------
// General backward preparation (entropy gradient requested): stash -> dL/dz in place (bf16)
//    p = e / S;   dz[r][v] = scale * ( dlogp[r] * (1[v == label r] - p) - dent[r] * p * (log p + H r) )
// after which the backward GEMMs run with unit scales. Rows whose upstream gradients are all zero are written as zeros
// without reading the stash.
// ------------------------------------------------------------------------------------------
// Walks the stash in its storage order: a CTA owns one block of 64 rows and a run of 64-column blocks, each of them 8 KB
// contiguous in HBM (a row-by-row walk would touch 2374 different 8 KB blocks with 128 bytes each per row). The 64
// rows' constants sit in shared memory; every thread moves two 16-byte vectors per block, two blocks in flight.
constexpr int kDlogitsColBlocksPerCta = 8;
__global__ void __launch_bounds__(256)
stash_to_dlogits_kernel(__nv_bfloat16* __restrict__ stash, uint32_t stash_vb, uint32_t rows, uint32_t vocab,
                        const float* __restrict__ inv_sum, const float* __restrict__ dlogp,
                        const float* __restrict__ dent, const float* __restrict__ ent,
                        const int64_t* __restrict__ labels, float scale) {
  // dz = e * (A - B * log2 e) with per-row A = c * (-g - ge * (ln c + H)), B = c * ge * ln 2  (p = c * e, c = 1 / S):
  // one MUFU and two FMA-pipe operations per element; the one-hot term adds g at the label's column
  __shared__ float s_g[64], s_a[64], s_b[64];
  __shared__ int64_t s_lab[64];
  const uint32_t rb = blockIdx.y;
  if (threadIdx.x < 64) {
    const uint32_t r = rb * 64 + threadIdx.x;
    const bool ok = r < rows;
    const float c = ok ? inv_sum[r] : 0.f;
    const float g = ok ? dlogp[r] * scale : 0.f;
    const float ge = (ok && dent) ? dent[r] * scale : 0.f;
    const float h = (ok && dent) ? ent[r] : 0.f;
    s_g[threadIdx.x] = g;
    s_a[threadIdx.x] = (ge != 0.f) ? c * (-g - ge * (__logf(c) + h)) : -g * c;
    s_b[threadIdx.x] = c * ge * 0.6931471805599453f;
    s_lab[threadIdx.x] = ok ? labels[r] : -1;
  }
  __syncthreads();
  const uint32_t n_vb = (vocab + 63) >> 6;  // column blocks that hold real columns (stash_vb may be padded beyond)
  const uint32_t vb0 = blockIdx.x * kDlogitsColBlocksPerCta;
  const uint32_t vb1 = min(vb0 + kDlogitsColBlocksPerCta, n_vb);
  uint4* base = reinterpret_cast<uint4*>(stash + static_cast<size_t>(rb) * stash_vb * 4096);
  // vector q of a block: row q / 8 of the block, columns (q % 8) * 8 .. + 7
  auto transform = [&](uint32_t vb, uint32_t q, const uint4& in) {
    const uint32_t ri = q >> 3;
    if (vb * 64 + (q & 7) * 8 >= vocab || rb * 64 + ri >= rows) return in;  // padding of the last blocks: left as it is
    const float g = s_g[ri], a = s_a[ri], b = s_b[ri];
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (g != 0.f || b != 0.f) {  // rows without upstream gradient (dlogp = dent = 0) become zeros
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
        f[i] = (e > 0.f) ? e * fmaf(-b, __log2f(e), a) : 0.f;  // e == 0: p = 0, no contribution (and no 0 * inf)
      }
      const int64_t d = s_lab[ri] - static_cast<int64_t>(vb * 64 + (q & 7) * 8);
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
    return out;
  };
  const uint32_t q0 = threadIdx.x, q1 = threadIdx.x + 256;
  uint32_t vb = vb0;
  for (; vb + 1 < vb1; vb += 2) {
    uint4* b0 = base + static_cast<size_t>(vb) * 512;
    uint4* b1 = b0 + 512;
    const uint4 a0 = b0[q0], a1 = b0[q1], a2 = b1[q0], a3 = b1[q1];
    b0[q0] = transform(vb, q0, a0);
    b0[q1] = transform(vb, q1, a1);
    b1[q0] = transform(vb + 1, q0, a2);
    b1[q1] = transform(vb + 1, q1, a3);
  }
  if (vb < vb1) {
    uint4* b0 = base + static_cast<size_t>(vb) * 512;
    const uint4 a0 = b0[q0], a1 = b0[q1];
    b0[q0] = transform(vb, q0, a0);
    b0[q1] = transform(vb, q1, a1);
  }
}

}  // namespace grpo
