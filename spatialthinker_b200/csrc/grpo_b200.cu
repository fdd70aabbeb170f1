// C ABI (include/grpo_b200.h) over the sm_100a kernels. Host side only: argument checks, TMA descriptors,
// workspace carving, launch sequencing on the caller's stream. No allocation, no synchronisation.
#include "../../include/grpo_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "advantage_kernels.cuh"
#include "compact_kernels.cuh"
#include "estimator_kernels.cuh"
#include "gemm_core.cuh"
#include "grad_kernels.cuh"
#include "lmhead_kernels.cuh"
#include "logits_kernels.cuh"
#include "loss_kernels.cuh"
#include "peer_kernels.cuh"

namespace grpo {

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define GRPO_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(static_cast<int>(e__), "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define GRPO_TRY(expr)        \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != 0) return rc__; \
  } while (0)

// ------------------------------------------------------------------------------------------ launch counter / phase timer
// Evidence plumbing for bench.py: how many of OUR kernels were enqueued, and (when enabled) CUDA-event timing of each
// pipeline phase on the caller's stream. Disabled by default; costs nothing then.
static std::atomic<long long> g_launches{0};
static inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

enum Phase { PH_LOGITS_GEMM = 0, PH_ROW_STATS, PH_LOSS, PH_GRAD_PREP, PH_DH_GEMM, PH_DW_GEMM, PH_N };
struct Profiler {
  std::atomic<bool> on{false};
  std::mutex mu;
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[PH_N];
  double ms[PH_N] = {};
  long long cnt[PH_N] = {};
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
static Profiler g_prof;
struct PhaseScope {
  Phase ph;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr;
  PhaseScope(Phase p, cudaStream_t s) : ph(p), st(s) {
    if (g_prof.on.load(std::memory_order_relaxed)) {
      std::lock_guard<std::mutex> lk(g_prof.mu);
      e0 = g_prof.get();
      cudaEventRecord(e0, st);
    }
  }
  ~PhaseScope() {
    if (e0) {
      std::lock_guard<std::mutex> lk(g_prof.mu);
      cudaEvent_t e1 = g_prof.get();
      cudaEventRecord(e1, st);
      g_prof.pending[ph].push_back({e0, e1});
    }
  }
};

// ------------------------------------------------------------------------------------------ TMA descriptors
static inline uint64_t cdiv64(uint64_t x) { return (x + 63) / 64; }
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D bf16 tensor [outer][inner] (inner contiguous, row pitch `pitch_elems`), box = [box_outer][box_inner],
// 128-byte swizzle, out-of-bounds reads return zeros.
static int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(GRPO_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch_elems * 2) % 16 != 0)
    return fail(GRPO_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {pitch_elems * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GRPO_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// Blocked bf16 operand [blk3][blk2][64][64] (64 x 64 blocks of 8 KB, inner row = 128 B); box = (64, 64, box2, box3).
static int make_tmap_blocked(CUtensorMap* map, const void* base, uint64_t n_blk2, uint64_t n_blk3, uint64_t blk3_pitch,
                             uint32_t box2, uint32_t box3, uint32_t box_rows = 64, uint32_t box_cols = 64) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(GRPO_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if (reinterpret_cast<uintptr_t>(base) & 15u) return fail(GRPO_ERR_ARG, "TMA operand must be 16-byte aligned");
  const cuuint64_t dims[4] = {64, 64, n_blk2, n_blk3};
  const cuuint64_t strides[3] = {128, 8192, blk3_pitch * 8192};
  const cuuint32_t box[4] = {box_cols, box_rows, box2, box3};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GRPO_ERR_DRIVER, "cuTensorMapEncodeTiled (blocked) failed with CUresult %d", (int)r);
  return 0;
}

// fp32 matrix [rows][cols] (row pitch `pitch_elems`), box 32 x 32 (128-byte rows, 128-byte swizzle): the epilogue's
// bulk store / reduce-add target. (Two 16-column boxes per group out of alternating staging halves were measured
// slower: the reduce-adds of all CTAs arrive together and are bound by L2 atomic throughput, ~5.6 TB/s.)
static int make_tmap_f32_out(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(GRPO_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch_elems * 4) % 16 != 0)
    return fail(GRPO_ERR_ARG, "fp32 output must be 16-byte aligned with a 16-byte multiple row pitch");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch_elems * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GRPO_ERR_DRIVER, "cuTensorMapEncodeTiled (fp32 out) failed with CUresult %d", (int)r);
  return 0;
}

// A/B operand descriptor with the box the producer warp expects.
//   plain : [rows][k] (K-major) or [k][rows] (MN-major), row pitch `pitch` elements
//   blocked (A only): 64 x 64 blocks; `pitch` = number of blocks along the buffer's inner block dimension
static int make_operand_tmap(CUtensorMap* map, const void* base, uint64_t rows, uint64_t k, uint64_t pitch, int mode,
                             uint32_t rows_per_load) {
  switch (mode) {
    case A_K_MAJOR: return make_tmap_bf16(map, base, k, rows, pitch, kBlockK, rows_per_load);
    case A_MN_MAJOR: return make_tmap_bf16(map, base, rows, k, pitch, 64, kBlockK);
    case A_BLOCKED_K:  // [rows/64][k/64][64][64]: blk2 = K blocks, blk3 = row blocks
      return make_tmap_blocked(map, base, cdiv64(k), cdiv64(rows), pitch, 1, rows_per_load / 64);
    default:           // A_BLOCKED_MN: [k/64][rows/64][64][64]: blk2 = row blocks, blk3 = K blocks
      return make_tmap_blocked(map, base, cdiv64(rows), cdiv64(k), pitch, rows_per_load / 64, 1);
  }
}

// ------------------------------------------------------------------------------------------ device info / knobs
// Tuning knobs. Defaults are what profiles/ measures; GRPO_* environment variables (read once) or grpo_set_option()
// override them for experiments.
struct Knobs {
  int cta_group = 2;   // 2: one 256 x 256 tile per CTA pair (cta_group::2); 1: 128 x 256 per CTA (debug)
  int ksub = 2;        // 128-row accumulators per CTA and tile: 2 = wide 512 x 256 pair tile, 1 = 256 x 256 double-buffered
  int fwd_panel = 4864;  // rows of hidden the logits GEMM keeps L2-resident under one vocab sweep (35 MB at H = 3584)
  // progress-window periods in K-blocks (0 = free running). Persistent CTA pairs drift apart, and tiles that share an
  // operand panel then each stream their own copy from HBM (dHidden GEMM: 32 GB instead of 11 GB per chunk, measured
  // with ncu). Bounding the drift to two short windows keeps the shared panels L2-hot; with the wide tile (ksub = 2)
  // that is worth 6-7 % end to end (profiles/r1_knobs.md). Values: interleaved A/B sweep on B200.
  int sync_fwd = 28, sync_dh = 8, sync_dw = 8;
  // L2 eviction priorities on the TMA loads, bit 0: logits GEMM (hidden panel evict-last, W evict-first), bit 1: dW GEMM
  // (scaled hidden evict-last, stash evict-first). Measured (profiles/r1_knobs.md): both cost 2-3 % - default off.
  int l2_hints = 0;
  int dh_m_fast = 0;   // tile order of the dHidden GEMM (experiment)
  int chunk_rows = 0;  // rows per chunk of the pipeline (0: 37 row tiles, see default_chunk_rows)
  int wait_hint_ns = 10000000;  // suspend hint of the epilogue warps' accumulator-ready wait (0: busy poll)
  // softmax epilogue of the logits GEMM, bit 0: software-pipelined TMEM drain, bit 1: exp-stash through shared memory +
  // bulk tensor stores, bit 2: 2 KB staging halves (EpiSoftmax::Params::mode)
  int epi_mode = 7;
  int acc_lead = 2;    // wide tile: accumulator 0's lead over accumulator 1 at the tile ends, in K-blocks (TileSched::acc_lead)
  int st_hint = 3;     // bit 0: stash bulk stores evict-first, bit 1: dW bulk reduce-adds evict-first (else normal)
  int clk_probe = 0;   // 1: the GEMM kernels record clock64 / globaltimer at entry and exit (grpo_debug_probe_offset)
  int dw_tma = 1;      // dW GEMM epilogue: 1 = bulk tensor reduce-add from shared memory, 0 = per-thread red.global.add
  int dw_split = 1;    // dW GEMM: split-K tail for the last, partial round of tiles (TileSched::split_tail; 2 = multi-round plan)
  int epi_share = 0;   // logits GEMM, wide tile: both epilogue warpgroups drain accumulator 0, then 1 (TileSched::epi_share)
  // dHidden GEMM: when its tile count is not a whole number of rounds over the CTA pairs (micro-batches that are not a
  // multiple of 37 row tiles, e.g. the reference's 4-sequence micro-batches), accumulate in fp32 with a split-K tail and
  // convert in a fix-up pass (chunk_backward). 0: always the direct bf16 epilogue. Measured on B200
  // (profiles/r1_ab_dh_split.log): dHidden GEMM 3.93 -> 3.16 ms at 4096 rows, 1.68 -> 0.76 ms at 1024 rows.
  int dh_split = 1;
  // Run-to-run bit-reproducible gradients: no split-K (dw_split / dh_split are ignored: the K slices of a tile add into
  // the same fp32 words in completion order) and the one-hot rows of dW summed in row order instead of with atomics
  // (grad_kernels.cuh). Costs ~1 % (profiles/); off by default.
  int deterministic = 0;
  // serpentine K order (TileSched::serpentine), bit 0: dW GEMM, bit 1: dHidden GEMM
  int k_serp = 0;
};
static Knobs g_knobs;
static std::once_flag g_knobs_once;
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
static void init_knobs() {
  std::call_once(g_knobs_once, [] {
    g_knobs.cta_group = env_int("GRPO_CTA_GROUP", g_knobs.cta_group) == 1 ? 1 : 2;
    g_knobs.ksub = env_int("GRPO_KSUB", g_knobs.ksub) == 1 ? 1 : 2;
    g_knobs.fwd_panel = env_int("GRPO_FWD_PANEL", g_knobs.fwd_panel);
    g_knobs.sync_fwd = env_int("GRPO_SYNC_FWD", g_knobs.sync_fwd);
    g_knobs.sync_dh = env_int("GRPO_SYNC_DH", g_knobs.sync_dh);
    g_knobs.sync_dw = env_int("GRPO_SYNC_DW", g_knobs.sync_dw);
    g_knobs.l2_hints = env_int("GRPO_L2_HINTS", g_knobs.l2_hints);
    g_knobs.dh_m_fast = env_int("GRPO_DH_M_FAST", g_knobs.dh_m_fast);
    g_knobs.chunk_rows = env_int("GRPO_CHUNK_ROWS", g_knobs.chunk_rows);
    g_knobs.wait_hint_ns = env_int("GRPO_WAIT_HINT_NS", g_knobs.wait_hint_ns);
    g_knobs.epi_mode = env_int("GRPO_EPI_MODE", g_knobs.epi_mode) & 7;
    g_knobs.dw_tma = env_int("GRPO_DW_TMA", g_knobs.dw_tma) != 0;
    g_knobs.dw_split = env_int("GRPO_DW_SPLIT", g_knobs.dw_split);
    g_knobs.dw_split = g_knobs.dw_split < 0 ? 0 : (g_knobs.dw_split > 2 ? 2 : g_knobs.dw_split);
    g_knobs.epi_share = env_int("GRPO_EPI_SHARE", g_knobs.epi_share) != 0;
    g_knobs.dh_split = env_int("GRPO_DH_SPLIT", g_knobs.dh_split) != 0;
    g_knobs.acc_lead = env_int("GRPO_ACC_LEAD", g_knobs.acc_lead);
    g_knobs.st_hint = env_int("GRPO_ST_HINT", g_knobs.st_hint) & 3;
    g_knobs.deterministic = env_int("GRPO_DETERMINISTIC", g_knobs.deterministic) != 0;
    g_knobs.k_serp = env_int("GRPO_K_SERP", g_knobs.k_serp) & 3;
  });
}

struct DevInfo : Knobs {
  int sms = 0;
};
static int get_dev(DevInfo* out) {
  static int sms_cache[64] = {};
  init_knobs();
  int dev = 0;
  GRPO_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(GRPO_ERR_ARG, "device ordinal out of range");
  if (sms_cache[dev] == 0) {
    cudaDeviceProp p;
    GRPO_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) return fail(GRPO_ERR_ARG, "this library targets sm_100a (B200); found sm_%d%d", p.major, p.minor);
    sms_cache[dev] = p.multiProcessorCount;
  }
  static_cast<Knobs&>(*out) = g_knobs;
  if (out->deterministic) out->dw_split = out->dh_split = 0;
  out->sms = sms_cache[dev];
  return 0;
}

// ------------------------------------------------------------------------------------------ GEMM launch
static inline uint32_t cdiv(uint64_t a, uint64_t b) { return static_cast<uint32_t>((a + b - 1) / b); }

// Split-K plan for a partial round of `rem` tiles on `groups` persistent CTA groups (TileSched::split_tail == 2): every
// tail tile is cut along K into `slices` units and the rem * slices units are walked slice-major in `rounds` short
// rounds. Picks the slice count (<= 16) with the least rounds x (slice length + one accumulator drain), in K-blocks.
constexpr uint32_t kSplitDrainKb = 8;  // fp32 reduce-add drain of one wide tile ~ 8.4 k cycles ~ 8 K-blocks (r1_tile_trace.log)
struct TailPlan {
  uint32_t slices, kpb, rounds;
  uint64_t cost;  // K-block times
};
static TailPlan plan_tail(uint32_t rem, uint32_t groups, uint32_t k_blocks) {
  TailPlan best{1, k_blocks, (rem + groups - 1) / groups, 0};
  best.cost = static_cast<uint64_t>(best.rounds) * (k_blocks + kSplitDrainKb);
  for (uint32_t want = 2; want <= 16 && want <= k_blocks; ++want) {
    const uint32_t kpb = (k_blocks + want - 1) / want;
    const uint32_t slices = (k_blocks + kpb - 1) / kpb;  // every slice holds at least one K-block
    const uint32_t rounds = static_cast<uint32_t>((static_cast<uint64_t>(rem) * slices + groups - 1) / groups);
    const uint64_t cost = static_cast<uint64_t>(rounds) * (kpb + kSplitDrainKb);
    if (cost < best.cost) best = TailPlan{slices, kpb, rounds, cost};
  }
  return best;
}
// Estimated duration (K-block times) of a GEMM of `tiles` tiles x `k_blocks` on `groups` groups: whole-tile rounds
// only, or whole rounds + the planned split-K tail through the accumulating fp32 epilogue.
static uint64_t gemm_cost_plain(uint32_t tiles, uint32_t groups, uint32_t k_blocks) {
  return static_cast<uint64_t>((tiles + groups - 1) / groups) * k_blocks;
}
static uint64_t gemm_cost_split(uint32_t tiles, uint32_t groups, uint32_t k_blocks) {
  const uint32_t rem = tiles % groups;
  uint64_t c = static_cast<uint64_t>(tiles / groups) * (k_blocks + kSplitDrainKb);
  if (rem != 0) c += plan_tail(rem, groups, k_blocks).cost;
  return c;
}

// Work units of one launch: fills the unit fields of `sched` for `tiles` output tiles on at most `groups_avail`
// persistent CTA groups and returns the number of groups to launch. Units are whole tiles, plus - for accumulating
// epilogues (sched.split_tail != 0) - K slices of the tiles of the last, partial round.
static uint32_t plan_units(TileSched& sched, uint32_t tiles, uint32_t groups_avail) {
  uint32_t groups = groups_avail;
  TailPlan plan{1, 0, 0, 0};
  if (sched.split_tail == 2 && tiles % groups != 0 && sched.k_blocks >= 2)
    plan = plan_tail(tiles % groups, groups, sched.k_blocks);  // may use every group even when tiles < groups
  if (plan.slices <= 1 && groups > tiles) groups = tiles;
  const uint32_t rem = tiles % groups;
  sched.num_units = sched.whole_tiles = tiles;
  sched.split_slices = sched.split_kpb = sched.split_tiles = 0;
  sched.max_progress = cdiv(tiles, groups) * sched.k_blocks;
  if (plan.slices > 1) {  // mode 2: several short rounds of K slices, slice-major
    sched.whole_tiles = tiles - rem;
    sched.split_slices = plan.slices;
    sched.split_kpb = plan.kpb;
    sched.split_tiles = rem;
    sched.num_units = sched.whole_tiles + rem * plan.slices;
    sched.max_progress = (tiles / groups) * sched.k_blocks + plan.rounds * plan.kpb;
  } else if (sched.split_tail == 1 && rem != 0 && 2 * rem <= groups && sched.k_blocks >= 2) {
    // mode 1: the last round would leave more than half of the groups idle for a full tile time - one short round instead
    uint32_t slices = groups / rem;
    if (slices > sched.k_blocks) slices = sched.k_blocks;
    const uint32_t kpb = cdiv(sched.k_blocks, slices);
    slices = cdiv(sched.k_blocks, kpb);  // every slice holds at least one K-block
    sched.whole_tiles = tiles - rem;
    sched.split_slices = slices;
    sched.split_kpb = kpb;
    sched.num_units = sched.whole_tiles + rem * slices;  // rem * slices <= groups: one more (short) round
    sched.max_progress = (tiles / groups) * sched.k_blocks + kpb;
  }
  return groups;
}

template <int kCta, int kSub, int BLOCK_N, int kStages, int kAMode, bool kBMn, class Epi>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, TileSched sched,
                       const typename Epi::Params& ep, int sms, cudaStream_t stream) {
  using Cfg = GemmCfg<kCta, kSub, BLOCK_N, kStages>;
  auto kern = gemm_kernel<kCta, kSub, BLOCK_N, kStages, kAMode, kBMn, Epi>;
  const size_t smem = Cfg::smem_bytes(Epi::kSmemBytes);
  GRPO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const uint32_t tiles = sched.m_blocks * sched.n_blocks;
  if (tiles == 0) return 0;
  if (sms / kCta <= 0) return fail(GRPO_ERR_ARG, "no CTA groups on this device");
  const uint32_t groups = plan_units(sched, tiles, static_cast<uint32_t>(sms / kCta));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * kCta);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GRPO_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, sched, ep));
  count_launch();
  return 0;
}

constexpr int kBlockN = 256;
// smem ring depth per configuration (stage = A rows + B rows, 128 B each):
constexpr int kStages11 = 4;  // 1 CTA,  128 + 256 rows = 48 KB
constexpr int kStages21 = 6;  // pair,   128 + 128 rows = 32 KB
constexpr int kStages22 = 4;  // pair,   256 + 128 rows = 48 KB   (wide tile)

// Tile geometry for a knob setting: rows of A per CTA-group tile.
static inline int tile_rows(int cta_group, int ksub) { return kBlockM * cta_group * (cta_group == 2 ? ksub : 1); }

// Rows per chunk of the chunked lm_head pipeline. With 74 CTA pairs the dHidden GEMM's (rows / tile_rows) x (H / 256)
// output tiles should be a whole number of waves: 37 row blocks give 4 / 7 waves for H = 2048 / 3584. That is 9472
// rows with 256-row tiles (exp-stash 2.9 GB at V = 151936) and 18944 rows with the wide 512-row tiles (5.8 GB).
static inline int64_t default_chunk_rows(int cta_group, int ksub) { return 37ll * tile_rows(cta_group == 1 ? 2 : cta_group, ksub); }

template <int kAMode, bool kBMn, class Epi1, class Epi2>
static int launch_gemm_any(const DevInfo& dev, const void* a, uint64_t a_rows, uint64_t a_pitch, const void* b,
                           uint64_t b_rows, uint64_t b_pitch, uint64_t k, TileSched sched,
                           const typename Epi1::Params& ep1, const typename Epi2::Params& ep2, cudaStream_t stream) {
  const int cta_group = dev.cta_group, ksub = (cta_group == 2) ? dev.ksub : 1;
  const int trows = tile_rows(cta_group, ksub);
  CUtensorMap ta, tb;
  GRPO_TRY(make_operand_tmap(&ta, a, a_rows, k, a_pitch, kAMode, trows / cta_group));
  GRPO_TRY(make_operand_tmap(&tb, b, b_rows, k, b_pitch, kBMn ? A_MN_MAJOR : A_K_MAJOR, kBlockN / cta_group));
  sched.m_blocks = cdiv(a_rows, trows);
  sched.n_blocks = cdiv(b_rows, kBlockN);
  sched.k_blocks = cdiv(k, kBlockK);
  if (sched.panel_m == 0 || sched.panel_m > sched.m_blocks) sched.panel_m = sched.m_blocks;
  if (sched.hint_a == 0) sched.hint_a = kEvictNormal;
  if (sched.hint_b == 0) sched.hint_b = kEvictNormal;
  sched.wait_hint_ns = static_cast<uint32_t>(dev.wait_hint_ns);
  sched.acc_lead = static_cast<uint32_t>(dev.acc_lead);
  if (cta_group == 1) return launch_gemm<1, 1, kBlockN, kStages11, kAMode, kBMn, Epi1>(ta, tb, sched, ep1, dev.sms, stream);
  if (ksub == 1) return launch_gemm<2, 1, kBlockN, kStages21, kAMode, kBMn, Epi2>(ta, tb, sched, ep2, dev.sms, stream);
  return launch_gemm<2, 2, kBlockN, kStages22, kAMode, kBMn, Epi2>(ta, tb, sched, ep2, dev.sms, stream);
}

// ------------------------------------------------------------------------------------------ workspace
struct Workspace {
  // sizes for one chunk of rows
  int64_t chunk_rows = 0, rows_pad = 0, n_tiles = 0;
  int64_t stash_vb = 0;                // 64-column blocks per stash row block = ceil(V / 64)
  int64_t stash_rows = 0;              // rows the stash holds from its base pointer (rows_pad; less in a slot view)
  // exp(z - z_label) (entropy-gradient mode: rewritten to dL/dz), BLOCKED: [rows_pad/64][stash_vb][64 rows][64 cols].
  // 8 KB contiguous per block, so both backward GEMMs - one walking vocab blocks as K, the other row blocks as K -
  // stream it from HBM in whole DRAM pages instead of 128-byte pieces 300 KB apart.
  __nv_bfloat16* stash = nullptr;
  __nv_bfloat16* hd_scaled = nullptr;  // [chunk_rows][H] row-scaled hidden: B operand of the dW GEMM
  float* dh_acc = nullptr;             // [chunk_rows][H] fp32 accumulator of the dHidden GEMM's split-K path
  float *part_sum = nullptr, *part_ez = nullptr;  // [n_tiles][rows_pad]
  float *a_label = nullptr, *lse = nullptr, *inv_sum = nullptr, *dlogp = nullptr, *dent = nullptr, *ent = nullptr;
  float *row_scale = nullptr, *onehot = nullptr;
  int32_t *oh_next = nullptr, *oh_has_prev = nullptr;  // per-label row chains of the deterministic one-hot pass
  double* acc = nullptr;
  uint32_t* sync = nullptr;  // progress-barrier counters, one per GEMM of the chunk pipeline (first 64 bytes)
  unsigned long long* probe = nullptr;  // clock probes of the three GEMMs, 1024 x u64 each (measurement aid)
  size_t probe_offset = 0;
  size_t bytes = 0;
};
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Workspace carve(void* base, int64_t rows, int64_t hdim, int64_t vocab, bool with_stash) {
  Workspace w;
  init_knobs();
  const int64_t chunk_cap = g_knobs.chunk_rows > 0 ? g_knobs.chunk_rows : default_chunk_rows(g_knobs.cta_group, g_knobs.ksub);
  w.chunk_rows = rows < chunk_cap ? rows : chunk_cap;
  if (w.chunk_rows < 1) w.chunk_rows = 1;
  w.rows_pad = static_cast<int64_t>(align_up(static_cast<size_t>(w.chunk_rows), 512));
  w.stash_rows = w.rows_pad;
  w.n_tiles = (vocab + kBlockN - 1) / kBlockN;
  size_t off = 0;
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t nbytes) {
    uint8_t* r = p ? p + off : nullptr;
    off = align_up(off + nbytes, 1024);
    return r;
  };
  w.stash_vb = (vocab + 63) / 64;
  if (with_stash) {
    w.stash = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(w.rows_pad) * w.stash_vb * 64 * 2));
    w.hd_scaled = reinterpret_cast<__nv_bfloat16*>(take(static_cast<size_t>(w.chunk_rows) * hdim * 2));
    w.dh_acc = reinterpret_cast<float*>(take(static_cast<size_t>(w.chunk_rows) * hdim * 4));
  }
  const size_t part = 2 * static_cast<size_t>(w.n_tiles) * w.rows_pad * 4;  // two column parts per tile when epi_share
  w.part_sum = reinterpret_cast<float*>(take(part));
  w.part_ez = reinterpret_cast<float*>(take(part));
  const size_t vec = static_cast<size_t>(w.rows_pad) * 4;
  w.a_label = reinterpret_cast<float*>(take(vec));
  w.lse = reinterpret_cast<float*>(take(vec));
  w.inv_sum = reinterpret_cast<float*>(take(vec));
  w.dlogp = reinterpret_cast<float*>(take(vec));
  w.dent = reinterpret_cast<float*>(take(vec));
  w.ent = reinterpret_cast<float*>(take(vec));
  w.row_scale = reinterpret_cast<float*>(take(vec));
  w.onehot = reinterpret_cast<float*>(take(vec));
  w.oh_next = reinterpret_cast<int32_t*>(take(vec));
  w.oh_has_prev = reinterpret_cast<int32_t*>(take(vec));
  w.acc = reinterpret_cast<double*>(take(ACC_N * sizeof(double)));
  w.probe_offset = off + 64;
  w.sync = reinterpret_cast<uint32_t*>(take(64 + 3 * 8192));
  w.probe = p ? reinterpret_cast<unsigned long long*>(p + w.probe_offset) : nullptr;
  w.bytes = off;
  return w;
}

// A window of a chunk workspace that starts `row0` rows in (row0 % 512 == 0: whole tiles): the stash, the scaled hidden
// rows and every per-row array are shifted, leading dimensions stay. Several small micro-batches can then each run
// forward + dHidden in their own window of ONE chunk's stash and share a single dW GEMM over all of its rows
// (grpo_fused_loss_fwd_bwd_slot / grpo_deferred_dw_flush).
static Workspace slot_view(const Workspace& w, int64_t row0, int64_t hdim) {
  Workspace v = w;
  v.stash = w.stash + static_cast<size_t>(row0 / 64) * w.stash_vb * 4096;
  v.stash_rows = w.rows_pad - row0;
  v.hd_scaled = w.hd_scaled + static_cast<size_t>(row0) * hdim;
  v.part_sum = w.part_sum + row0;
  v.part_ez = w.part_ez + row0;
  v.a_label = w.a_label + row0;
  v.lse = w.lse + row0;
  v.inv_sum = w.inv_sum + row0;
  v.dlogp = w.dlogp + row0;
  v.dent = w.dent + row0;
  v.ent = w.ent + row0;
  v.row_scale = w.row_scale + row0;
  v.onehot = w.onehot + row0;
  v.oh_next = w.oh_next + row0;
  v.oh_has_prev = w.oh_has_prev + row0;
  return v;
}

static int check_head_args(const void* hidden, const void* weight, int64_t rows, int64_t h, int64_t v, float temp) {
  if (!hidden || !weight) return fail(GRPO_ERR_ARG, "hidden / weight must not be null");
  if (rows < 0 || h <= 0 || v <= 0) return fail(GRPO_ERR_ARG, "negative or zero dimension");
  if (h % 64 != 0) return fail(GRPO_ERR_ARG, "hidden_dim must be a multiple of 64 (got %lld)", (long long)h);
  if (v % 8 != 0) return fail(GRPO_ERR_ARG, "vocab must be a multiple of 8 (got %lld)", (long long)v);
  if (!(temp > 0.f)) return fail(GRPO_ERR_ARG, "temperature must be positive");
  if (rows > 0x7fffffffll || v > 0x7fffffffll) return fail(GRPO_ERR_ARG, "dimension exceeds 2^31");
  return 0;
}

// ------------------------------------------------------------------------------------------ chunk steps
// label logit -> logits GEMM + softmax sums (+ optional exp stash) -> per-row combine, for rows [r0, r0 + n)
static int chunk_forward(const DevInfo& dev, const Workspace& w, const __nv_bfloat16* hidden,
                         const __nv_bfloat16* weight, const int64_t* labels, int64_t r0, int64_t n, int64_t h,
                         int64_t v, float temperature, bool want_entropy, bool want_stash, float* logp_out,
                         float* ent_out, float* lse_out, cudaStream_t stream) {
  {
    PhaseScope ps(PH_ROW_STATS, stream);
    label_dot_kernel<<<cdiv(n * 32, 256), 256, 0, stream>>>(hidden + r0 * h, weight, labels + r0,
                                                            static_cast<uint32_t>(n), static_cast<uint32_t>(h),
                                                            static_cast<uint32_t>(v), w.a_label);
    count_launch();
    GRPO_CUDA(cudaGetLastError());
  }
  EpiSoftmax<1, kBlockN>::Params p1;
  memset(&p1, 0, sizeof(p1));
  p1.mode = static_cast<uint32_t>(dev.epi_mode);
  p1.policy = (dev.st_hint & 1) ? kEvictFirst : kEvictNormal;
  if (!want_stash || !(p1.mode & 1)) p1.mode &= 1u;
  if (!(p1.mode & 2)) p1.mode &= 3u;
  if (p1.mode & 2)  // store view of the blocked stash: 32 rows x 64 columns (4 KB, contiguous in HBM) per bulk store
    GRPO_TRY(make_tmap_blocked(&p1.stash_map, w.stash, static_cast<uint64_t>(w.stash_vb),
                               static_cast<uint64_t>(w.stash_rows / 64), static_cast<uint64_t>(w.stash_vb), 1, 1, 32));
  if (p1.mode & 4)  // ... or 32 rows x 32 columns (64-byte pieces of 32 consecutive 128-byte rows) per bulk store
    GRPO_TRY(make_tmap_blocked(&p1.stash_map_half, w.stash, static_cast<uint64_t>(w.stash_vb),
                               static_cast<uint64_t>(w.stash_rows / 64), static_cast<uint64_t>(w.stash_vb), 1, 1, 32, 32));
  p1.rows = static_cast<uint32_t>(n);
  p1.vocab = static_cast<uint32_t>(v);
  p1.rows_pad = static_cast<uint32_t>(w.rows_pad);
  p1.scale = 1.f / temperature;
  p1.ref = w.a_label;
  p1.part_sum = w.part_sum;
  p1.part_ez = want_entropy ? w.part_ez : nullptr;
  p1.stash = want_stash ? w.stash : nullptr;
  p1.stash_vb = static_cast<uint32_t>(w.stash_vb);
  EpiSoftmax<2, kBlockN>::Params p2;
  static_assert(sizeof(p1) == sizeof(p2), "epilogue params layout");
  memcpy(&p2, &p1, sizeof(p1));
  TileSched s{};
  s.m_fast = 1;  // walk the row blocks of a panel under one vocab tile: the hidden panel stays in L2, W streams by
  {
    const int trows = tile_rows(dev.cta_group, dev.cta_group == 2 ? dev.ksub : 1);
    s.panel_m = static_cast<uint32_t>((dev.fwd_panel + trows / 2) / trows);
    if (s.panel_m == 0) s.panel_m = 1;
  }
  s.sync_period = static_cast<uint32_t>(dev.sync_fwd);
  s.sync_ctr = w.sync;
  const bool share = dev.epi_share != 0 && dev.cta_group == 2 && dev.ksub == 2;  // wide tile only
  s.epi_share = share ? 1u : 0u;
  s.probe = dev.clk_probe ? w.probe : nullptr;
  if (dev.l2_hints & 1) {  // the hidden panel is re-read under every vocab tile; a W tile is dead after one panel pass
    s.hint_a = kEvictLast;
    s.hint_b = kEvictFirst;
  }
  if (dev.l2_hints & 8) s.hint_a = kEvictLast;  // only the hidden panel pinned; W at normal priority
  GRPO_CUDA(cudaMemsetAsync(w.sync, 0, 64, stream));
  {
    PhaseScope ps(PH_LOGITS_GEMM, stream);
    GRPO_TRY((launch_gemm_any<A_K_MAJOR, false, EpiSoftmax<1, kBlockN>, EpiSoftmax<2, kBlockN>>(
        dev, hidden + r0 * h, n, h, weight, v, h, h, s, p1, p2, stream)));
  }
  PhaseScope ps(PH_ROW_STATS, stream);
  combine_rows_kernel<<<cdiv(n, 32), dim3(32, 8), 0, stream>>>(
      w.part_sum, want_entropy ? w.part_ez : nullptr, w.a_label, labels + r0, static_cast<uint32_t>(n),
      static_cast<uint32_t>(w.rows_pad), static_cast<uint32_t>(w.n_tiles * (share ? 2 : 1)), static_cast<uint32_t>(v),
      1.f / temperature, lse_out, logp_out, ent_out, w.inv_sum);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

// dW[v][h] += E^T[v][n] . hd[n][h]    A = the stash's first n rows read transposed (MN-major), B = (scaled) hidden read
// transposed. The fp32 epilogue accumulates (bulk reduce-add), so a call adds one group of rows' contribution to `dweight`.
static int dw_gemm(const DevInfo& dev, const Workspace& w, const __nv_bfloat16* b_op, int64_t n, int64_t h, int64_t v,
                   float* dweight, cudaStream_t stream) {
  const uint32_t uh = static_cast<uint32_t>(h), uv = static_cast<uint32_t>(v);
  PhaseScope ps(PH_DW_GEMM, stream);
  EpiF32<1, kBlockN>::Params p1;
  memset(&p1, 0, sizeof(p1));
  p1.c = dweight;
  p1.ldc = h;
  p1.m = uv;
  p1.n = uh;
  p1.accumulate = 1u;
  p1.use_tma = dev.dw_tma ? 1u : 0u;
  p1.policy = (dev.st_hint & 2) ? kEvictFirst : kEvictNormal;
  if (p1.use_tma) GRPO_TRY(make_tmap_f32_out(&p1.c_map, dweight, uh, uv, uh));
  EpiF32<2, kBlockN>::Params p2;
  static_assert(sizeof(p1) == sizeof(p2), "epilogue params layout");
  memcpy(&p2, &p1, sizeof(p1));
  TileSched s{};
  s.m_fast = 0;  // the H column blocks of one vocab block run together: the stash panel is read from HBM once
  s.sync_period = static_cast<uint32_t>(dev.sync_dw);
  s.sync_ctr = w.sync + 2;
  s.split_tail = static_cast<uint32_t>(dev.dw_split);  // the epilogue accumulates (reduce-add): K slices just add up
  s.serpentine = static_cast<uint32_t>(dev.k_serp & 1);
  s.probe = dev.clk_probe ? w.probe + 2048 : nullptr;
  if (dev.l2_hints & 2) {  // the (scaled) hidden chunk is re-read for every vocab block; the stash streams through once
    s.hint_a = kEvictFirst;
    s.hint_b = kEvictLast;
  }
  if (dev.l2_hints & 4) s.hint_b = kEvictLast;  // only the scaled hidden pinned; the stash panel (shared by the 14 column
                                                // tiles of a vocab block running side by side) at normal priority
  GRPO_TRY((launch_gemm_any<A_BLOCKED_MN, true, EpiF32<1, kBlockN>, EpiF32<2, kBlockN>>(
      dev, w.stash, v, w.stash_vb, b_op, h, h, n, s, p1, p2, stream)));
  return 0;
}

// dHidden = dlogits . W and dW += dlogits^T . hidden for rows [r0, r0 + n), from the stash of chunk_forward.
//   dent == nullptr (the GRPO loss): dlogits factorises per row, the stash is used as is (scale_scatter_kernel).
//   dent != nullptr (entropy gradient): the stash is first rewritten into dlogits (stash_to_dlogits_kernel).
static int chunk_backward(const DevInfo& dev, const Workspace& w, const __nv_bfloat16* hidden,
                          const __nv_bfloat16* weight, const int64_t* labels, const float* dlogp, const float* dent,
                          const float* ent, int64_t r0, int64_t n, int64_t h, int64_t v, float temperature,
                          __nv_bfloat16* dhidden, float* dweight, cudaStream_t stream, bool skip_dw = false) {
  const bool factorised = dent == nullptr;
  const uint32_t un = static_cast<uint32_t>(n), uh = static_cast<uint32_t>(h), uv = static_cast<uint32_t>(v);
  {
    PhaseScope ps(PH_GRAD_PREP, stream);
    if (factorised) {
      scale_scatter_kernel<<<cdiv(n * 32, 256), 256, 0, stream>>>(hidden + r0 * h, labels + r0, dlogp, w.inv_sum,
                                                                  1.f / temperature, un, uh, uv, w.row_scale, w.onehot,
                                                                  w.hd_scaled, dev.deterministic ? nullptr : dweight);
      if (dev.deterministic) {  // the one-hot rows of dW in row order instead of with atomics
        GRPO_CUDA(cudaMemsetAsync(w.oh_has_prev, 0, static_cast<size_t>(n) * sizeof(int32_t), stream));
        onehot_links_kernel<<<cdiv(n * 32, 256), 256, 0, stream>>>(labels + r0, un, uv, w.oh_next, w.oh_has_prev);
        onehot_ordered_kernel<<<cdiv(n * 32, 256), 256, 0, stream>>>(hidden + r0 * h, labels + r0, w.onehot, w.oh_next,
                                                                     w.oh_has_prev, un, uh, uv, dweight);
        count_launch(2);
      }
    } else {
      dim3 grid(cdiv(cdiv(v, 64), kDlogitsColBlocksPerCta), cdiv(n, 64));  // (runs of column blocks, row blocks)
      stash_to_dlogits_kernel<<<grid, 256, 0, stream>>>(w.stash, static_cast<uint32_t>(w.stash_vb), un, uv, w.inv_sum,
                                                        dlogp, dent, ent, labels + r0, 1.f / temperature);
    }
    count_launch();
    GRPO_CUDA(cudaGetLastError());
  }
  // dHidden[n][h] = row_scale * (E[n][v] . W[v][h]) + onehot * W[label]   A = stash (K-major), B = W read transposed
  bool dh_split = false;
  if (dev.dh_split) {  // worth it when the tile count leaves the last round of CTA groups mostly idle
    const uint32_t groups = static_cast<uint32_t>(dev.sms / dev.cta_group);
    const uint32_t tiles = cdiv(n, tile_rows(dev.cta_group, dev.cta_group == 2 ? dev.ksub : 1)) * cdiv(h, kBlockN);
    const uint32_t kb = static_cast<uint32_t>(w.stash_vb);
    dh_split = groups > 0 && tiles % groups != 0 &&
               100 * gemm_cost_split(tiles, groups, kb) <= 93 * gemm_cost_plain(tiles, groups < tiles ? groups : tiles, kb);
  }
  if (dh_split) {  // fp32 accumulator + split-K tail (the K slices of a tile add up in the reduce-add epilogue), then
                   // the scale / one-hot / bf16 arithmetic of EpiBF16 in a fix-up pass over [n][h]
    PhaseScope ps(PH_DH_GEMM, stream);
    GRPO_CUDA(cudaMemsetAsync(w.dh_acc, 0, static_cast<size_t>(n) * h * sizeof(float), stream));
    EpiF32<1, kBlockN>::Params p1;
    memset(&p1, 0, sizeof(p1));
    p1.c = w.dh_acc;
    p1.ldc = h;
    p1.m = un;
    p1.n = uh;
    p1.accumulate = 1u;
    p1.use_tma = dev.dw_tma ? 1u : 0u;
    p1.policy = kEvictNormal;  // read back by the fix-up pass right away
    if (p1.use_tma) GRPO_TRY(make_tmap_f32_out(&p1.c_map, w.dh_acc, uh, un, uh));
    EpiF32<2, kBlockN>::Params p2;
    memcpy(&p2, &p1, sizeof(p1));
    TileSched s{};
    s.m_fast = static_cast<uint32_t>(dev.dh_m_fast);
    s.sync_period = static_cast<uint32_t>(dev.sync_dh);
    s.sync_ctr = w.sync + 1;
    s.split_tail = 2;
    s.probe = dev.clk_probe ? w.probe + 1024 : nullptr;
    GRPO_TRY((launch_gemm_any<A_BLOCKED_K, true, EpiF32<1, kBlockN>, EpiF32<2, kBlockN>>(
        dev, w.stash, n, w.stash_vb, weight, h, h, v, s, p1, p2, stream)));
    const uint32_t vecs = un * (uh >> 3);
    dh_fixup_kernel<<<cdiv(vecs, 256), 256, 0, stream>>>(w.dh_acc, factorised ? w.row_scale : nullptr,
                                                         factorised ? w.onehot : nullptr, labels + r0, weight, un, uh,
                                                         dhidden + r0 * h);
    count_launch();
    GRPO_CUDA(cudaGetLastError());
  } else {
    PhaseScope ps(PH_DH_GEMM, stream);
    EpiBF16<1, kBlockN>::Params p1{dhidden + r0 * h, h, un, uh, factorised ? w.row_scale : nullptr,
                                   factorised ? w.onehot : nullptr, labels + r0, weight, h};
    EpiBF16<2, kBlockN>::Params p2{dhidden + r0 * h, h, un, uh, factorised ? w.row_scale : nullptr,
                                   factorised ? w.onehot : nullptr, labels + r0, weight, h};
    TileSched s{};
    s.m_fast = static_cast<uint32_t>(dev.dh_m_fast);  // 0: all H column blocks of a few row blocks run together
    s.sync_period = static_cast<uint32_t>(dev.sync_dh);
    s.sync_ctr = w.sync + 1;
    s.probe = dev.clk_probe ? w.probe + 1024 : nullptr;
    s.serpentine = static_cast<uint32_t>((dev.k_serp >> 1) & 1);
    if (dev.l2_hints & 16) s.hint_b = kEvictLast;  // W is re-read by every round of tiles; the stash streams through once
    GRPO_TRY((launch_gemm_any<A_BLOCKED_K, true, EpiBF16<1, kBlockN>, EpiBF16<2, kBlockN>>(
        dev, w.stash, n, w.stash_vb, weight, h, h, v, s, p1, p2, stream)));
  }
  if (!skip_dw) {
    const __nv_bfloat16* b_op = factorised ? w.hd_scaled : hidden + r0 * h;
    GRPO_TRY(dw_gemm(dev, w, b_op, n, h, v, dweight, stream));
  }
  return 0;
}

static LossCfg make_loss_cfg(float clip_lo, float clip_hi, float clip_dual, int kl_mode, float kl_coef,
                             float grad_accum) {
  LossCfg c;
  c.log_clip_lo = static_cast<float>(log(1.0 - static_cast<double>(clip_lo)));
  c.log_clip_hi = static_cast<float>(log(1.0 + static_cast<double>(clip_hi)));
  c.clip_dual = clip_dual;
  c.kl_coef = kl_coef;
  c.kl_mode = kl_mode;
  c.inv_grad_accum = 1.f / grad_accum;
  return c;
}
// grid of broadcast_adv_kernel: up to 4 x 256 columns per block and pass, rows folded onto gridDim.y
static inline dim3 bcast_grid(int64_t bsz, int64_t t_len) {
  const int64_t gx = (t_len + 1023) / 1024;
  const int64_t cap = (148 * 8 * 4 + gx - 1) / gx;  // a few waves of blocks, then rows are looped over
  return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(bsz < cap ? bsz : cap));
}
static inline uint32_t ew_blocks(int64_t n, int threads, int sms) {
  const int64_t want = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sms > 0 ? sms : 148) * 8;
  return static_cast<uint32_t>(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace grpo

using namespace grpo;

// ============================================================================================ C ABI
extern "C" {

int grpo_abi_version(void) { return 2; }
const char* grpo_last_error(void) { return g_err; }
long long grpo_launch_count(void) { return g_launches.load(); }

int grpo_set_option(const char* name, int value) {
  init_knobs();
  if (!name) return fail(GRPO_ERR_ARG, "null option name");
  if (!strcmp(name, "cta_group")) g_knobs.cta_group = value == 1 ? 1 : 2;
  else if (!strcmp(name, "ksub")) g_knobs.ksub = value == 1 ? 1 : 2;
  else if (!strcmp(name, "fwd_panel")) g_knobs.fwd_panel = value > 0 ? value : 4864;
  else if (!strcmp(name, "sync_fwd")) g_knobs.sync_fwd = value;
  else if (!strcmp(name, "sync_dh")) g_knobs.sync_dh = value;
  else if (!strcmp(name, "sync_dw")) g_knobs.sync_dw = value;
  else if (!strcmp(name, "l2_hints")) g_knobs.l2_hints = value;
  else if (!strcmp(name, "dh_m_fast")) g_knobs.dh_m_fast = value;
  else if (!strcmp(name, "wait_hint_ns")) g_knobs.wait_hint_ns = value < 0 ? 0 : value;
  else if (!strcmp(name, "epi_mode")) g_knobs.epi_mode = value & 7;
  else if (!strcmp(name, "dw_tma")) g_knobs.dw_tma = value != 0;
  else if (!strcmp(name, "st_hint")) g_knobs.st_hint = value & 3;
  else if (!strcmp(name, "clk_probe")) g_knobs.clk_probe = value != 0;
  else if (!strcmp(name, "acc_lead")) g_knobs.acc_lead = value < 0 ? 0 : value;
  else if (!strcmp(name, "dw_split")) g_knobs.dw_split = value < 0 ? 0 : (value > 2 ? 2 : value);  // 2: grpo_debug_gemm runs the multi-round plan
  else if (!strcmp(name, "epi_share")) g_knobs.epi_share = value != 0;
  else if (!strcmp(name, "dh_split")) g_knobs.dh_split = value != 0;
  else if (!strcmp(name, "deterministic")) g_knobs.deterministic = value != 0;
  else if (!strcmp(name, "k_serp")) g_knobs.k_serp = value & 3;
  else if (!strcmp(name, "chunk_rows")) g_knobs.chunk_rows = value > 0 ? (value + 511) / 512 * 512 : 0;
  else return fail(GRPO_ERR_ARG, "unknown option '%s'", name);
  return 0;
}

int grpo_profile_enable(int on) {
  g_prof.on.store(on != 0);
  return 0;
}
int grpo_profile_read(double* ms_out, long long* count_out, int reset) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int p = 0; p < PH_N; ++p) {
    for (auto& pr : g_prof.pending[p]) {
      GRPO_CUDA(cudaEventSynchronize(pr.second));
      float ms = 0.f;
      GRPO_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
      g_prof.ms[p] += ms;
      g_prof.cnt[p] += 1;
      g_prof.pool.push_back(pr.first);
      g_prof.pool.push_back(pr.second);
    }
    g_prof.pending[p].clear();
    if (ms_out) ms_out[p] = g_prof.ms[p];
    if (count_out) count_out[p] = g_prof.cnt[p];
    if (reset) {
      g_prof.ms[p] = 0.0;
      g_prof.cnt[p] = 0;
    }
  }
  return 0;
}

size_t grpo_debug_probe_offset(int64_t rows, int64_t hidden_dim, int64_t vocab, int with_stash) {
  return carve(nullptr, rows, hidden_dim, vocab, with_stash != 0).probe_offset;
}
size_t grpo_lmhead_fwd_workspace_bytes(int64_t rows, int64_t hidden_dim, int64_t vocab) {
  return carve(nullptr, rows, hidden_dim, vocab, false).bytes;
}
size_t grpo_lmhead_bwd_workspace_bytes(int64_t rows, int64_t hidden_dim, int64_t vocab) {
  return carve(nullptr, rows, hidden_dim, vocab, true).bytes;
}
size_t grpo_fused_loss_workspace_bytes(int64_t rows, int64_t hidden_dim, int64_t vocab) {
  return carve(nullptr, rows, hidden_dim, vocab, true).bytes;
}

int grpo_lmhead_logprob_fwd(const void* hidden, const void* weight, const int64_t* labels, int64_t rows,
                            int64_t hidden_dim, int64_t vocab, float temperature, float* logp, float* entropy,
                            float* lse, void* workspace, size_t workspace_bytes, grpo_stream_t stream) {
  if (rows == 0) return 0;
  GRPO_TRY(check_head_args(hidden, weight, rows, hidden_dim, vocab, temperature));
  if (!labels || !logp) return fail(GRPO_ERR_ARG, "labels / logp must not be null");
  DevInfo dev;
  GRPO_TRY(get_dev(&dev));
  const Workspace w = carve(workspace, rows, hidden_dim, vocab, false);
  if (!workspace || workspace_bytes < w.bytes)
    return fail(GRPO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  auto* hp = static_cast<const __nv_bfloat16*>(hidden);
  auto* wp = static_cast<const __nv_bfloat16*>(weight);
  for (int64_t r0 = 0; r0 < rows; r0 += w.chunk_rows) {
    const int64_t n = (rows - r0 < w.chunk_rows) ? rows - r0 : w.chunk_rows;
    GRPO_TRY(chunk_forward(dev, w, hp, wp, labels, r0, n, hidden_dim, vocab, temperature, entropy != nullptr, false,
                           logp + r0, entropy ? entropy + r0 : nullptr, lse ? lse + r0 : nullptr, stream));
  }
  return 0;
}

int grpo_lmhead_bwd(const void* hidden, const void* weight, const int64_t* labels, const float* dlogp,
                    const float* dentropy, int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature,
                    void* dhidden, float* dweight, void* workspace, size_t workspace_bytes, grpo_stream_t stream) {
  if (rows == 0) return 0;
  GRPO_TRY(check_head_args(hidden, weight, rows, hidden_dim, vocab, temperature));
  if (!labels || !dlogp || !dhidden || !dweight)
    return fail(GRPO_ERR_ARG, "labels / dlogp / dhidden / dweight must not be null");
  DevInfo dev;
  GRPO_TRY(get_dev(&dev));
  const Workspace w = carve(workspace, rows, hidden_dim, vocab, true);
  if (!workspace || workspace_bytes < w.bytes)
    return fail(GRPO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  auto* hp = static_cast<const __nv_bfloat16*>(hidden);
  auto* wp = static_cast<const __nv_bfloat16*>(weight);
  for (int64_t r0 = 0; r0 < rows; r0 += w.chunk_rows) {
    const int64_t n = (rows - r0 < w.chunk_rows) ? rows - r0 : w.chunk_rows;
    GRPO_TRY(chunk_forward(dev, w, hp, wp, labels, r0, n, hidden_dim, vocab, temperature, dentropy != nullptr, true,
                           nullptr, dentropy ? w.ent : nullptr, w.lse, stream));
    GRPO_TRY(chunk_backward(dev, w, hp, wp, labels, dlogp + r0, dentropy ? dentropy + r0 : nullptr, w.ent, r0, n,
                            hidden_dim, vocab, temperature, static_cast<__nv_bfloat16*>(dhidden), dweight, stream));
  }
  return 0;
}

// Argument checks shared by the fused-loss entry points.
static int check_fused_loss_args(const void* hidden, const void* weight, const int64_t* labels, const float* old_logp,
                                 const float* advantages, const float* ref_logp, const void* mask, int mask_dtype,
                                 int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature, int kl_mode,
                                 float entropy_coef, float grad_accum, const float* logp_out, const float* entropy_out,
                                 const void* dhidden, const float* dweight, const float* metrics) {
  GRPO_TRY(check_head_args(hidden, weight, rows, hidden_dim, vocab, temperature));
  if (!labels || !old_logp || !advantages || !logp_out || !metrics)
    return fail(GRPO_ERR_ARG, "labels / old_logp / advantages / logp_out / metrics must not be null");
  if ((dhidden == nullptr) != (dweight == nullptr))
    return fail(GRPO_ERR_ARG, "dhidden and dweight must be given together");
  if (mask_dtype < 0 || mask_dtype > 3 || (mask_dtype != MASK_NONE && !mask))
    return fail(GRPO_ERR_ARG, "bad mask / mask_dtype");
  if (kl_mode < GRPO_KL_NONE || kl_mode > GRPO_KL_CHI2) return fail(GRPO_ERR_ARG, "unknown kl_mode %d", kl_mode);
  if (kl_mode != GRPO_KL_NONE && !ref_logp) return fail(GRPO_ERR_ARG, "kl_mode set but ref_logp is null");
  if (entropy_coef != 0.f && !entropy_out) return fail(GRPO_ERR_ARG, "entropy_coef != 0 needs entropy_out");
  if (!(grad_accum > 0.f)) return fail(GRPO_ERR_ARG, "grad_accum must be positive");
  return 0;
}

// One micro-batch through workspace (view) `w`: mask sum, then per chunk forward -> token loss -> backward, then the
// metric vector. skip_dw: leave the stash-dependent part of dW to a later dw_gemm over the whole workspace.
static int fused_loss_body(const DevInfo& dev, const Workspace& w, const void* hidden, const void* weight,
                           const int64_t* labels, const float* old_logp, const float* advantages,
                           const float* ref_logp, const void* mask, int mask_dtype, int64_t rows, int64_t hidden_dim,
                           int64_t vocab, float temperature, const LossCfg& cfg, float* logp_out, float* entropy_out,
                           void* dhidden, float* dweight, float* metrics, bool skip_dw, cudaStream_t stream) {
  const bool want_bwd = dhidden != nullptr;
  const float entropy_coef = cfg.entropy_coef;
  auto* hp = static_cast<const __nv_bfloat16*>(hidden);
  auto* wp = static_cast<const __nv_bfloat16*>(weight);
  const size_t esz = (mask_dtype == MASK_F32) ? 4 : (mask_dtype == MASK_I64 ? 8 : 1);

  GRPO_CUDA(cudaMemsetAsync(w.acc, 0, ACC_N * sizeof(double), stream));
  if (rows > 0) {
    // the normaliser sum(mask) spans the whole micro-batch and is needed before the first dL/dlogp (dp_actor.py:255)
    mask_sum_kernel<<<ew_blocks(rows, 256, dev.sms), 256, 0, stream>>>(mask, mask_dtype, static_cast<size_t>(rows),
                                                                        w.acc);
    count_launch();
    GRPO_CUDA(cudaGetLastError());
  }
  const bool want_ent = entropy_out != nullptr;
  for (int64_t r0 = 0; r0 < rows; r0 += w.chunk_rows) {
    const int64_t n = (rows - r0 < w.chunk_rows) ? rows - r0 : w.chunk_rows;
    GRPO_TRY(chunk_forward(dev, w, hp, wp, labels, r0, n, hidden_dim, vocab, temperature, want_ent, want_bwd,
                           logp_out + r0, want_ent ? entropy_out + r0 : nullptr, w.lse, stream));
    const void* mchunk = (mask_dtype == MASK_NONE) ? nullptr : static_cast<const uint8_t*>(mask) + r0 * esz;
    {
      PhaseScope ps(PH_LOSS, stream);
      token_loss_kernel<<<ew_blocks(n, 256, dev.sms), 256, 0, stream>>>(
        logp_out + r0, old_logp + r0, advantages + r0, ref_logp ? ref_logp + r0 : nullptr,
        want_ent ? entropy_out + r0 : nullptr, mchunk, mask_dtype, static_cast<size_t>(n), cfg, w.acc,
        want_bwd ? w.dlogp : nullptr, (want_bwd && entropy_coef != 0.f) ? w.dent : nullptr);
      count_launch();
    }
    GRPO_CUDA(cudaGetLastError());
    if (want_bwd)
      GRPO_TRY(chunk_backward(dev, w, hp, wp, labels, w.dlogp, entropy_coef != 0.f ? w.dent : nullptr,
                              want_ent ? entropy_out + r0 : nullptr, r0, n, hidden_dim, vocab, temperature,
                              static_cast<__nv_bfloat16*>(dhidden), dweight, stream, skip_dw));
  }
  loss_finalize_kernel<<<1, 32, 0, stream>>>(w.acc, cfg, metrics);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_fused_loss_fwd_bwd(const void* hidden, const void* weight, const int64_t* labels, const float* old_logp,
                            const float* advantages, const float* ref_logp, const void* mask, int mask_dtype,
                            int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature, float clip_ratio_low,
                            float clip_ratio_high, float clip_ratio_dual, int kl_mode, float kl_coef,
                            float entropy_coef, float grad_accum, float* logp_out, float* entropy_out, void* dhidden,
                            float* dweight, float* metrics, void* workspace, size_t workspace_bytes,
                            grpo_stream_t stream) {
  GRPO_TRY(check_fused_loss_args(hidden, weight, labels, old_logp, advantages, ref_logp, mask, mask_dtype, rows,
                                 hidden_dim, vocab, temperature, kl_mode, entropy_coef, grad_accum, logp_out,
                                 entropy_out, dhidden, dweight, metrics));
  DevInfo dev;
  GRPO_TRY(get_dev(&dev));
  const Workspace w = carve(workspace, rows, hidden_dim, vocab, true);
  if (!workspace || workspace_bytes < w.bytes)
    return fail(GRPO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  LossCfg cfg = make_loss_cfg(clip_ratio_low, clip_ratio_high, clip_ratio_dual, kl_mode, kl_coef, grad_accum);
  cfg.entropy_coef = entropy_coef;
  return fused_loss_body(dev, w, hidden, weight, labels, old_logp, advantages, ref_logp, mask, mask_dtype, rows,
                         hidden_dim, vocab, temperature, cfg, logp_out, entropy_out, dhidden, dweight, metrics, false,
                         stream);
}

long long grpo_chunk_capacity_rows(void) {
  init_knobs();
  return g_knobs.chunk_rows > 0 ? g_knobs.chunk_rows : default_chunk_rows(g_knobs.cta_group, g_knobs.ksub);
}

// Workspace of the deferred-dW entry points: one chunk of `capacity_rows` rows (a multiple of 512, at most one chunk).
static int carve_deferred(Workspace* w, void* workspace, size_t workspace_bytes, int64_t capacity_rows,
                          int64_t hidden_dim, int64_t vocab) {
  if (capacity_rows <= 0 || capacity_rows % 512 != 0 || capacity_rows > grpo_chunk_capacity_rows())
    return fail(GRPO_ERR_ARG, "capacity_rows must be a positive multiple of 512, at most grpo_chunk_capacity_rows()");
  *w = carve(workspace, capacity_rows, hidden_dim, vocab, true);
  if (!workspace || workspace_bytes < w->bytes)
    return fail(GRPO_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w->bytes, workspace_bytes);
  return 0;
}

int grpo_fused_loss_fwd_bwd_slot(const void* hidden, const void* weight, const int64_t* labels, const float* old_logp,
                                 const float* advantages, const float* ref_logp, const void* mask, int mask_dtype,
                                 int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature,
                                 float clip_ratio_low, float clip_ratio_high, float clip_ratio_dual, int kl_mode,
                                 float kl_coef, float grad_accum, float* logp_out, float* entropy_out, void* dhidden,
                                 float* dweight, float* metrics, int64_t slot_row0, int64_t capacity_rows,
                                 void* workspace, size_t workspace_bytes, grpo_stream_t stream) {
  GRPO_TRY(check_fused_loss_args(hidden, weight, labels, old_logp, advantages, ref_logp, mask, mask_dtype, rows,
                                 hidden_dim, vocab, temperature, kl_mode, 0.f, grad_accum, logp_out, entropy_out,
                                 dhidden, dweight, metrics));
  if (!dhidden) return fail(GRPO_ERR_ARG, "the slot entry is the training path: dhidden / dweight must be given");
  if (rows <= 0) return fail(GRPO_ERR_ARG, "a slot needs at least one row");
  if (slot_row0 < 0 || slot_row0 % 512 != 0 || slot_row0 + rows > capacity_rows)
    return fail(GRPO_ERR_ARG, "slot [%lld, %lld) does not fit the workspace of %lld rows (slot_row0 %% 512 == 0)",
                (long long)slot_row0, (long long)(slot_row0 + rows), (long long)capacity_rows);
  DevInfo dev;
  GRPO_TRY(get_dev(&dev));
  Workspace full;
  GRPO_TRY(carve_deferred(&full, workspace, workspace_bytes, capacity_rows, hidden_dim, vocab));
  const Workspace w = slot_view(full, slot_row0, hidden_dim);
  const LossCfg cfg = [&] {
    LossCfg c = make_loss_cfg(clip_ratio_low, clip_ratio_high, clip_ratio_dual, kl_mode, kl_coef, grad_accum);
    c.entropy_coef = 0.f;  // an entropy gradient rewrites the stash and needs the hidden rows at dW time: not deferrable
    return c;
  }();
  GRPO_TRY(fused_loss_body(dev, w, hidden, weight, labels, old_logp, advantages, ref_logp, mask, mask_dtype, rows,
                           hidden_dim, vocab, temperature, cfg, logp_out, entropy_out, dhidden, dweight, metrics, true,
                           stream));
  // rows between this slot's end and the next 512-row boundary: their stash rows were written as zeros by the logits
  // GEMM (it stores every row of its tiles); the scaled-hidden rows they meet in the deferred dW GEMM must be finite
  const int64_t end = slot_row0 + rows;
  int64_t gap_end = (end + 511) / 512 * 512;
  if (gap_end > capacity_rows) gap_end = capacity_rows;
  if (gap_end > end)
    GRPO_CUDA(cudaMemsetAsync(full.hd_scaled + static_cast<size_t>(end) * hidden_dim, 0,
                              static_cast<size_t>(gap_end - end) * hidden_dim * sizeof(__nv_bfloat16), stream));
  return 0;
}

int grpo_deferred_dw_flush(int64_t total_rows, int64_t capacity_rows, int64_t hidden_dim, int64_t vocab, float* dweight,
                           void* workspace, size_t workspace_bytes, grpo_stream_t stream) {
  if (!dweight) return fail(GRPO_ERR_ARG, "dweight must not be null");
  if (hidden_dim <= 0 || hidden_dim % 64 != 0 || vocab <= 0 || vocab % 8 != 0)
    return fail(GRPO_ERR_ARG, "bad hidden_dim / vocab");
  if (total_rows < 0 || total_rows > capacity_rows) return fail(GRPO_ERR_ARG, "total_rows exceeds capacity_rows");
  if (total_rows == 0) return 0;
  DevInfo dev;
  GRPO_TRY(get_dev(&dev));
  Workspace w;
  GRPO_TRY(carve_deferred(&w, workspace, workspace_bytes, capacity_rows, hidden_dim, vocab));
  return dw_gemm(dev, w, w.hd_scaled, total_rows, hidden_dim, vocab, dweight, stream);
}

int grpo_grad_sumsq(float* grad, int64_t n, int zero_after, int accumulate, double* out, double* scratch,
                    grpo_stream_t stream) {
  if (!grad || !out || !scratch) return fail(GRPO_ERR_ARG, "grad / out / scratch must not be null");
  if (n < 0) return fail(GRPO_ERR_ARG, "negative length");
  if (reinterpret_cast<uintptr_t>(grad) & 15u) return fail(GRPO_ERR_ARG, "grad must be 16-byte aligned");
  grad_sumsq_kernel<<<kGradBlocks, kGradThreads, 0, stream>>>(grad, static_cast<size_t>(n), zero_after, scratch);
  grad_sumsq_finalize_kernel<<<1, 32, 0, stream>>>(scratch, kGradBlocks, accumulate, out);
  count_launch(2);
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_grad_scale_cast(float* grad, int64_t n, const float* scale_dev, float scale_host, void* out_bf16,
                         int zero_after, grpo_stream_t stream) {
  if (!grad || !out_bf16) return fail(GRPO_ERR_ARG, "grad / out must not be null");
  if (n < 0) return fail(GRPO_ERR_ARG, "negative length");
  if ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(out_bf16)) & 15u)
    return fail(GRPO_ERR_ARG, "grad / out must be 16-byte aligned");
  if (n == 0) return 0;
  grad_scale_cast_kernel<<<kGradBlocks * 2, kGradThreads, 0, stream>>>(grad, static_cast<size_t>(n), scale_dev, scale_host,
                                                                       static_cast<__nv_bfloat16*>(out_bf16), zero_after);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

// CUDA IPC plumbing for the peer-mapped buffers: the exporter describes a device pointer as (handle of the allocation it
// lies in, byte offset), the importer maps that allocation into ITS OWN device's address space (peer access is enabled on
// the way) and gets the pointer back. Host-side only, called once per buffer at set-up. Two exported pointers may lie in
// one allocation (a caching allocator carves tensors out of larger segments): a handle is opened once per process and
// reference-counted here.
namespace {
struct IpcMapping {
  char handle[64];
  void* base;
  int refs;
};
std::mutex g_ipc_mutex;
std::vector<IpcMapping> g_ipc_open;
}  // namespace

int grpo_ipc_export(const void* ptr, void* handle_out_64, int64_t* offset_out) {
  if (!ptr || !handle_out_64 || !offset_out) return fail(GRPO_ERR_ARG, "null argument");
  using RangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range_fn = nullptr;
  if (!range_fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return fail(GRPO_ERR_DRIVER, "cuMemGetAddressRange entry point not available");
    range_fn = reinterpret_cast<RangeFn>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) return fail(GRPO_ERR_DRIVER, "cuMemGetAddressRange failed with CUresult %d", (int)r);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  GRPO_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
  memcpy(handle_out_64, &h, 64);
  *offset_out = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return 0;
}
int grpo_ipc_open(const void* handle_64, int64_t offset, void** ptr_out) {
  if (!handle_64 || !ptr_out || offset < 0) return fail(GRPO_ERR_ARG, "bad argument");
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  for (IpcMapping& m : g_ipc_open) {
    if (!memcmp(m.handle, handle_64, 64)) {
      ++m.refs;
      *ptr_out = static_cast<uint8_t*>(m.base) + offset;
      return 0;
    }
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, 64);
  void* base = nullptr;
  GRPO_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  IpcMapping m;
  memcpy(m.handle, handle_64, 64);
  m.base = base;
  m.refs = 1;
  g_ipc_open.push_back(m);
  *ptr_out = static_cast<uint8_t*>(base) + offset;
  return 0;
}
int grpo_ipc_close(void* ptr, int64_t offset) {
  if (!ptr) return 0;
  void* base = static_cast<uint8_t*>(ptr) - offset;
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  for (size_t i = 0; i < g_ipc_open.size(); ++i) {
    if (g_ipc_open[i].base != base) continue;
    if (--g_ipc_open[i].refs == 0) {
      g_ipc_open.erase(g_ipc_open.begin() + static_cast<long>(i));
      GRPO_CUDA(cudaIpcCloseMemHandle(base));
    }
    return 0;
  }
  return fail(GRPO_ERR_ARG, "pointer was not opened with grpo_ipc_open");
}

namespace {
// the rank-ordered pointer table of one peer-mapped buffer
int peer_table(void* const* ptrs, int rank, int world, unsigned align_mask, const char* what, PeerPtrs* out) {
  if (!ptrs || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(GRPO_ERR_ARG, "%s: need 1..%d ranks and a valid rank", what, kMaxPeers);
  *out = PeerPtrs{};
  for (int q = 0; q < world; ++q) {
    if (!ptrs[q] || (reinterpret_cast<uintptr_t>(ptrs[q]) & align_mask))
      return fail(GRPO_ERR_ARG, "%s: pointer of rank %d is null or misaligned", what, q);
    out->p[q] = ptrs[q];
  }
  return 0;
}
// slab of `rank`: items [rank * per, min((rank + 1) * per, items)), per = ceil(items / world)
inline void peer_slab(size_t items, int rank, int world, size_t* i0, size_t* i1) {
  const size_t per = (items + static_cast<size_t>(world) - 1) / static_cast<size_t>(world);
  *i0 = per * static_cast<size_t>(rank) < items ? per * static_cast<size_t>(rank) : items;
  *i1 = *i0 + per < items ? *i0 + per : items;
}
}  // namespace

// element range [e0, e1) of the slab `rank` reduces in grpo_peer_reduce_scatter_sumsq / grpo_peer_scale_cast_allgather
// (host-only; lets the host-side mirror peer.slab_bounds be checked against the partition the kernels are launched with)
int grpo_debug_peer_slab(int64_t n, int rank, int world, int64_t* e0, int64_t* e1) {
  if (!e0 || !e1 || n < 0 || n % 8 != 0 || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(GRPO_ERR_ARG, "peer slab: bad argument");
  size_t u0, u1;
  peer_slab(static_cast<size_t>(n) / 8, rank, world, &u0, &u1);
  *e0 = static_cast<int64_t>(u0 * 8);
  *e1 = static_cast<int64_t>(u1 * 8);
  return 0;
}

#define GRPO_PEER_DISPATCH(world, LAUNCH) \
  switch (world) {                        \
    case 1: LAUNCH(1); break;             \
    case 2: LAUNCH(2); break;             \
    case 3: LAUNCH(3); break;             \
    case 4: LAUNCH(4); break;             \
    case 5: LAUNCH(5); break;             \
    case 6: LAUNCH(6); break;             \
    case 7: LAUNCH(7); break;             \
    default: LAUNCH(8); break;            \
  }

int grpo_peer_barrier(void* const* flag_ptrs, int rank, int world, unsigned int epoch, int timeout_ms,
                      grpo_stream_t stream) {
  PeerPtrs f;
  if (int rc = peer_table(flag_ptrs, rank, world, 3u, "peer barrier", &f)) return rc;
  const uint64_t timeout_ns = static_cast<uint64_t>(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, stream>>>(f, rank, world, epoch, timeout_ns);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_peer_allreduce_mean(void* const* buf_ptrs, int rank, int world, int64_t n, grpo_stream_t stream) {
  PeerPtrs b;
  if (int rc = peer_table(buf_ptrs, rank, world, 15u, "peer all-reduce", &b)) return rc;
  if (n < 0 || n % 4 != 0) return fail(GRPO_ERR_ARG, "element count must be a non-negative multiple of 4");
  if (n == 0 || world == 1) return 0;
  size_t v0, v1;
  peer_slab(static_cast<size_t>(n) / 4, rank, world, &v0, &v1);
  if (v0 >= v1) return 0;
  const float inv = 1.f / static_cast<float>(world);
#define GRPO_PEER_AR(W) peer_allreduce_mean_kernel<W><<<kPeerBlocks, kPeerThreads, 0, stream>>>(b, v0, v1, inv)
  GRPO_PEER_DISPATCH(world, GRPO_PEER_AR)
#undef GRPO_PEER_AR
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_peer_reduce_scatter_sumsq(void* const* buf_ptrs, void* const* partial_ptrs, int rank, int world, int64_t n,
                                   double* scratch, grpo_stream_t stream) {
  PeerPtrs b, parts;
  if (int rc = peer_table(buf_ptrs, rank, world, 15u, "peer reduce-scatter", &b)) return rc;
  if (int rc = peer_table(partial_ptrs, rank, world, 7u, "peer reduce-scatter partials", &parts)) return rc;
  if (!scratch) return fail(GRPO_ERR_ARG, "scratch must not be null");
  if (n < 0 || n % 8 != 0) return fail(GRPO_ERR_ARG, "element count must be a non-negative multiple of 8");
  size_t u0, u1;
  peer_slab(static_cast<size_t>(n) / 8, rank, world, &u0, &u1);
  const float inv = 1.f / static_cast<float>(world);
#define GRPO_PEER_RS(W)                                                                                             \
  peer_reduce_scatter_sumsq_kernel<W><<<kPeerBlocks, kPeerThreads, 0, stream>>>(                                    \
      b, static_cast<float4*>(b.p[rank]), rank, 2 * u0, 2 * u1, inv, parts, scratch)
  GRPO_PEER_DISPATCH(world, GRPO_PEER_RS)
#undef GRPO_PEER_RS
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_peer_scale_cast_allgather(float* grad, void* const* out_ptrs, int rank, int world, int64_t n,
                                   const float* scale_dev, float scale_host, int zero_after, grpo_stream_t stream) {
  PeerPtrs outs;
  if (int rc = peer_table(out_ptrs, rank, world, 15u, "peer all-gather", &outs)) return rc;
  if (!grad || (reinterpret_cast<uintptr_t>(grad) & 15u)) return fail(GRPO_ERR_ARG, "grad is null or not 16-byte aligned");
  if (n < 0 || n % 8 != 0) return fail(GRPO_ERR_ARG, "element count must be a non-negative multiple of 8");
  if (n == 0) return 0;
  const size_t units = static_cast<size_t>(n) / 8;
  size_t u0, u1;
  peer_slab(units, rank, world, &u0, &u1);
#define GRPO_PEER_AG(W)                                                                                           \
  peer_scale_cast_allgather_kernel<W><<<kPeerBlocks, kPeerThreads, 0, stream>>>(grad, units, u0, u1, scale_dev, \
                                                                                scale_host, outs, zero_after)
  GRPO_PEER_DISPATCH(world, GRPO_PEER_AG)
#undef GRPO_PEER_AG
#undef GRPO_PEER_DISPATCH
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_policy_loss_fwd_bwd(const float* logp, const float* old_logp, const float* advantages, const float* ref_logp,
                             const void* mask, int mask_dtype, int64_t n, float clip_ratio_low, float clip_ratio_high,
                             float clip_ratio_dual, int kl_mode, float kl_coef, float grad_accum, float* dlogp,
                             float* metrics, double* acc_scratch, grpo_stream_t stream) {
  if (!logp || !old_logp || !advantages || !metrics || !acc_scratch)
    return fail(GRPO_ERR_ARG, "logp / old_logp / advantages / metrics / acc_scratch must not be null");
  if (n < 0) return fail(GRPO_ERR_ARG, "negative length");
  if (mask_dtype < 0 || mask_dtype > 3 || (mask_dtype != MASK_NONE && !mask))
    return fail(GRPO_ERR_ARG, "bad mask / mask_dtype");
  if (kl_mode < GRPO_KL_NONE || kl_mode > GRPO_KL_CHI2) return fail(GRPO_ERR_ARG, "unknown kl_mode %d", kl_mode);
  if (kl_mode != GRPO_KL_NONE && !ref_logp) return fail(GRPO_ERR_ARG, "kl_mode set but ref_logp is null");
  if (!(grad_accum > 0.f)) return fail(GRPO_ERR_ARG, "grad_accum must be positive");
  LossCfg cfg = make_loss_cfg(clip_ratio_low, clip_ratio_high, clip_ratio_dual, kl_mode, kl_coef, grad_accum);
  cfg.entropy_coef = 0.f;
  GRPO_CUDA(cudaMemsetAsync(acc_scratch, 0, ACC_N * sizeof(double), stream));
  if (n > 0) {
    const uint32_t blocks = ew_blocks(n, 256, 0);
    mask_sum_kernel<<<blocks, 256, 0, stream>>>(mask, mask_dtype, static_cast<size_t>(n), acc_scratch);
    count_launch();
    token_loss_kernel<<<blocks, 256, 0, stream>>>(logp, old_logp, advantages, ref_logp, nullptr, mask, mask_dtype,
                                                  static_cast<size_t>(n), cfg, acc_scratch, dlogp, nullptr);
    count_launch();
  }
  loss_finalize_kernel<<<1, 32, 0, stream>>>(acc_scratch, cfg, metrics);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_compute_kl(const float* logp, const float* ref_logp, int64_t n, int kl_mode, float* out, float* dout_dlogp,
                    grpo_stream_t stream) {
  if (!logp || !ref_logp || !out) return fail(GRPO_ERR_ARG, "logp / ref_logp / out must not be null");
  if (kl_mode < GRPO_KL_LOW_VAR || kl_mode > GRPO_KL_CHI2) return fail(GRPO_ERR_ARG, "unknown kl_mode %d", kl_mode);
  if (n <= 0) return n == 0 ? 0 : fail(GRPO_ERR_ARG, "negative length");
  kl_elementwise_kernel<<<ew_blocks(n, 256, 0), 256, 0, stream>>>(logp, ref_logp, static_cast<size_t>(n), kl_mode, out,
                                                                  dout_dlogp);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_masked_mean(const float* x, const void* mask, int mask_dtype, int64_t n, float eps, float* out,
                     double* acc_scratch, grpo_stream_t stream) {
  if (!x || !out || !acc_scratch) return fail(GRPO_ERR_ARG, "x / out / acc_scratch must not be null");
  if (mask_dtype < 0 || mask_dtype > 3 || (mask_dtype != MASK_NONE && !mask))
    return fail(GRPO_ERR_ARG, "bad mask / mask_dtype");
  if (n < 0) return fail(GRPO_ERR_ARG, "negative length");
  GRPO_CUDA(cudaMemsetAsync(acc_scratch, 0, 2 * sizeof(double), stream));
  if (n > 0) {
    masked_sum_kernel<<<ew_blocks(n, 256, 0), 256, 0, stream>>>(x, mask, mask_dtype, static_cast<size_t>(n),
                                                                acc_scratch);
    count_launch();
  }
  masked_mean_finalize_kernel<<<1, 32, 0, stream>>>(acc_scratch, eps, out);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_sequence_scores(const float* rewards, int64_t bsz, int64_t t_len, float* scores, grpo_stream_t stream) {
  if (bsz == 0) return 0;
  if (!rewards || !scores) return fail(GRPO_ERR_ARG, "rewards / scores must not be null");
  if (bsz < 0 || t_len <= 0 || bsz > 0x7fffffffll || t_len > 0x7fffffffll) return fail(GRPO_ERR_ARG, "bad dimensions");
  row_score_kernel<<<cdiv(bsz * 32, 256), 256, 0, stream>>>(rewards, static_cast<uint32_t>(bsz),
                                                            static_cast<uint32_t>(t_len), scores);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_advantage_from_scores(const float* scores_all, const int32_t* order, const int32_t* offsets, int64_t bsz_all,
                               int64_t n_groups, float eps, int64_t row_begin, const void* mask, int mask_dtype,
                               int64_t bsz_local, int64_t t_len, float* advantages, float* seq_scratch,
                               grpo_stream_t stream) {
  if (!scores_all || !order || !offsets || !advantages || !seq_scratch)
    return fail(GRPO_ERR_ARG, "scores_all / order / offsets / advantages / seq_scratch must not be null");
  if (mask_dtype < 0 || mask_dtype > 3 || (mask_dtype != MASK_NONE && !mask))
    return fail(GRPO_ERR_ARG, "bad mask / mask_dtype");
  if (bsz_all < 0 || n_groups < 0 || row_begin < 0 || bsz_local < 0 || row_begin + bsz_local > bsz_all || t_len < 0 ||
      bsz_all > 0x7fffffffll || t_len > 0x7fffffffll)
    return fail(GRPO_ERR_ARG, "bad dimensions");
  if (bsz_all == 0) return 0;
  group_stats_kernel<<<cdiv(n_groups * 32, 256), 256, 0, stream>>>(scores_all, order, offsets,
                                                                   static_cast<uint32_t>(n_groups), eps, seq_scratch);
  count_launch();
  if (bsz_local > 0 && t_len > 0) {
    broadcast_adv_kernel<<<bcast_grid(bsz_local, t_len), 256, 0, stream>>>(
        seq_scratch + row_begin, mask, mask_dtype, static_cast<uint32_t>(bsz_local), static_cast<uint32_t>(t_len),
        advantages);
    count_launch();
  }
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_advantage(const float* rewards, const void* mask, int mask_dtype, const int32_t* order,
                   const int32_t* offsets, int64_t bsz, int64_t t_len, int64_t n_groups, float eps, float* advantages,
                   float* seq_scratch, grpo_stream_t stream) {
  if (!rewards || !order || !offsets || !advantages || !seq_scratch)
    return fail(GRPO_ERR_ARG, "rewards / order / offsets / advantages / seq_scratch must not be null");
  if (mask_dtype < 0 || mask_dtype > 3 || (mask_dtype != MASK_NONE && !mask))
    return fail(GRPO_ERR_ARG, "bad mask / mask_dtype");
  if (bsz < 0 || t_len < 0 || n_groups < 0 || bsz > 0x7fffffffll || t_len > 0x7fffffffll)
    return fail(GRPO_ERR_ARG, "bad dimensions");
  if (bsz == 0 || t_len == 0) return 0;
  float* scores = seq_scratch;
  float* seq_adv = seq_scratch + bsz;
  const uint32_t b = static_cast<uint32_t>(bsz), t = static_cast<uint32_t>(t_len);
  row_score_kernel<<<cdiv(bsz * 32, 256), 256, 0, stream>>>(rewards, b, t, scores);
  count_launch();
  group_stats_kernel<<<cdiv(n_groups * 32, 256), 256, 0, stream>>>(scores, order, offsets,
                                                                   static_cast<uint32_t>(n_groups), eps, seq_adv);
  count_launch();
  broadcast_adv_kernel<<<bcast_grid(bsz, t_len), 256, 0, stream>>>(seq_adv, mask, mask_dtype, b, t,
                                                                           advantages);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ other estimators
static int check_seq_args(const void* mask, int mask_dtype, int64_t bsz, int64_t t_len) {
  if (mask_dtype < 0 || mask_dtype > 3 || (mask_dtype != MASK_NONE && !mask))
    return fail(GRPO_ERR_ARG, "bad mask / mask_dtype");
  if (bsz < 0 || t_len < 0 || bsz > 0x7fffffffll || t_len > 0x7fffffffll) return fail(GRPO_ERR_ARG, "bad dimensions");
  return 0;
}

int grpo_rloo_advantage(const float* rewards, const void* mask, int mask_dtype, const int32_t* order,
                        const int32_t* offsets, int64_t bsz, int64_t t_len, int64_t n_groups, float* advantages,
                        float* seq_scratch, grpo_stream_t stream) {
  if (!rewards || !order || !offsets || !advantages || !seq_scratch)
    return fail(GRPO_ERR_ARG, "rewards / order / offsets / advantages / seq_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, bsz, t_len));
  if (n_groups < 0) return fail(GRPO_ERR_ARG, "bad dimensions");
  if (bsz == 0 || t_len == 0) return 0;
  float* scores = seq_scratch;
  float* seq_adv = seq_scratch + bsz;
  const uint32_t b = static_cast<uint32_t>(bsz), t = static_cast<uint32_t>(t_len);
  row_score_kernel<<<cdiv(bsz * 32, 256), 256, 0, stream>>>(rewards, b, t, scores);
  group_rloo_kernel<<<cdiv(n_groups * 32, 256), 256, 0, stream>>>(scores, order, offsets,
                                                                  static_cast<uint32_t>(n_groups), seq_adv);
  broadcast_adv_kernel<<<bcast_grid(bsz, t_len), 256, 0, stream>>>(seq_adv, mask, mask_dtype, b, t,
                                                                           advantages);
  count_launch(3);
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_remax_advantage(const float* rewards, const float* reward_baselines, const void* mask, int mask_dtype,
                         int64_t bsz, int64_t t_len, float* advantages, float* seq_scratch, grpo_stream_t stream) {
  if (!rewards || !reward_baselines || !advantages || !seq_scratch)
    return fail(GRPO_ERR_ARG, "rewards / reward_baselines / advantages / seq_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, bsz, t_len));
  if (bsz == 0 || t_len == 0) return 0;
  float* scores = seq_scratch;
  float* seq_adv = seq_scratch + bsz;
  const uint32_t b = static_cast<uint32_t>(bsz), t = static_cast<uint32_t>(t_len);
  row_score_kernel<<<cdiv(bsz * 32, 256), 256, 0, stream>>>(rewards, b, t, scores);
  remax_seq_kernel<<<cdiv(bsz, 256), 256, 0, stream>>>(scores, reward_baselines, b, seq_adv);
  broadcast_adv_kernel<<<bcast_grid(bsz, t_len), 256, 0, stream>>>(seq_adv, mask, mask_dtype, b, t,
                                                                           advantages);
  count_launch(3);
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

// whiten `x` over the mask into `out` (may alias x); acc_scratch: 3 doubles
static int whiten_launch(const float* x, const void* mask, int mask_dtype, int64_t n, float eps, float* out,
                         double* acc, cudaStream_t stream) {
  GRPO_CUDA(cudaMemsetAsync(acc, 0, 3 * sizeof(double), stream));
  const uint32_t blocks = ew_blocks(n, 256, 0);
  const size_t sn = static_cast<size_t>(n);
  masked_sum_kernel<<<blocks, 256, 0, stream>>>(x, mask, mask_dtype, sn, acc);
  masked_centered_sq_kernel<<<blocks, 256, 0, stream>>>(x, mask, mask_dtype, sn, acc);
  if (out != nullptr) {
    whiten_apply_kernel<<<blocks, 256, 0, stream>>>(x, sn, acc, eps, out);
    count_launch();
  }
  count_launch(2);
  return 0;
}

int grpo_masked_whiten(const float* values, const void* mask, int mask_dtype, int64_t n, float eps, float* out,
                       double* acc_scratch, grpo_stream_t stream) {
  if (!values || !out || !acc_scratch) return fail(GRPO_ERR_ARG, "values / out / acc_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, n, 1));
  if (n == 0) return 0;
  GRPO_TRY(whiten_launch(values, mask, mask_dtype, n, eps, out, acc_scratch, stream));
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_masked_var(const float* values, const void* mask, int mask_dtype, int64_t n, int unbiased, float* out,
                    double* acc_scratch, grpo_stream_t stream) {
  if (!values || !out || !acc_scratch) return fail(GRPO_ERR_ARG, "values / out / acc_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, n, 1));
  if (n == 0) {
    GRPO_CUDA(cudaMemsetAsync(acc_scratch, 0, 3 * sizeof(double), stream));
  } else {
    GRPO_TRY(whiten_launch(values, mask, mask_dtype, n, 0.f, nullptr, acc_scratch, stream));
  }
  masked_var_finalize_kernel<<<1, 32, 0, stream>>>(acc_scratch, unbiased, out);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_reinforce_pp_advantage(const float* rewards, const void* mask, int mask_dtype, int64_t bsz, int64_t t_len,
                                float gamma, float* advantages, float* returns, double* acc_scratch,
                                grpo_stream_t stream) {
  if (!rewards || !advantages || !returns || !acc_scratch)
    return fail(GRPO_ERR_ARG, "rewards / advantages / returns / acc_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, bsz, t_len));
  if (bsz == 0 || t_len == 0) return 0;
  const uint32_t sb = cdiv(bsz, kScanTile), b32 = static_cast<uint32_t>(bsz), t32 = static_cast<uint32_t>(t_len);
  switch (mask_dtype) {
    case MASK_I64:
      reverse_scan_kernel<1, long long, true><<<sb, 32, 0, stream>>>(rewards, nullptr, static_cast<const long long*>(mask), b32,
                                                               t32, gamma, 0.f, nullptr, returns);
      break;
    case MASK_U8:
      reverse_scan_kernel<1, unsigned char, true><<<sb, 32, 0, stream>>>(rewards, nullptr, static_cast<const unsigned char*>(mask),
                                                                   b32, t32, gamma, 0.f, nullptr, returns);
      break;
    case MASK_F32:
      reverse_scan_kernel<1, float, true><<<sb, 32, 0, stream>>>(rewards, nullptr, static_cast<const float*>(mask), b32, t32,
                                                                 gamma, 0.f, nullptr, returns);
      break;
    default:  // no mask: all ones
      reverse_scan_kernel<1, float, false><<<sb, 32, 0, stream>>>(rewards, nullptr, nullptr, b32, t32, gamma, 0.f, nullptr,
                                                                  returns);
  }
  count_launch();
  GRPO_TRY(whiten_launch(returns, mask, mask_dtype, bsz * t_len, 1e-8f, advantages, acc_scratch, stream));
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_gae_advantage(const float* rewards, const float* values, const void* mask, int mask_dtype, int64_t bsz,
                       int64_t t_len, float gamma, float gamma_lam, float* advantages, float* returns,
                       double* acc_scratch, grpo_stream_t stream) {
  if (!rewards || !values || !advantages || !returns || !acc_scratch)
    return fail(GRPO_ERR_ARG, "rewards / values / advantages / returns / acc_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, bsz, t_len));
  if (bsz == 0 || t_len == 0) return 0;
  reverse_scan_kernel<0, float, false><<<cdiv(bsz, kScanTile), 32, 0, stream>>>(
      rewards, values, nullptr, static_cast<uint32_t>(bsz), static_cast<uint32_t>(t_len), gamma, gamma_lam, advantages,
      returns);
  count_launch();
  GRPO_TRY(whiten_launch(advantages, mask, mask_dtype, bsz * t_len, 1e-8f, advantages, acc_scratch, stream));
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_value_loss_fwd_bwd(const float* vpreds, const float* returns, const float* values, const void* mask,
                            int mask_dtype, int64_t n, float cliprange_value, float* dvpreds, float* out,
                            double* acc_scratch, grpo_stream_t stream) {
  if (!vpreds || !returns || !values || !out || !acc_scratch)
    return fail(GRPO_ERR_ARG, "vpreds / returns / values / out / acc_scratch must not be null");
  GRPO_TRY(check_seq_args(mask, mask_dtype, n, 1));
  GRPO_CUDA(cudaMemsetAsync(acc_scratch, 0, 3 * sizeof(double), stream));
  if (n > 0) {
    const uint32_t blocks = ew_blocks(n, 256, 0);
    mask_sum_kernel<<<blocks, 256, 0, stream>>>(mask, mask_dtype, static_cast<size_t>(n), acc_scratch);
    value_loss_kernel<<<blocks, 256, 0, stream>>>(vpreds, returns, values, mask, mask_dtype, static_cast<size_t>(n),
                                                  cliprange_value, acc_scratch, dvpreds);
    count_launch(2);
  }
  value_loss_finalize_kernel<<<1, 32, 0, stream>>>(acc_scratch, out);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_kl_penalty_rewards(const float* token_level_scores, const float* logp, const float* ref_logp,
                            const void* mask, int mask_dtype, int64_t bsz, int64_t t_len, int kl_mode, float kl_coef,
                            float* token_level_rewards, float* current_kl, double* acc_scratch,
                            grpo_stream_t stream) {
  if (!token_level_scores || !token_level_rewards || !current_kl || !acc_scratch)
    return fail(GRPO_ERR_ARG, "token_level_scores / token_level_rewards / current_kl / acc_scratch must not be null");
  if (ref_logp && !logp) return fail(GRPO_ERR_ARG, "ref_logp given without logp");
  if (ref_logp && (kl_mode < GRPO_KL_LOW_VAR || kl_mode > GRPO_KL_CHI2))
    return fail(GRPO_ERR_ARG, "unknown kl_mode %d", kl_mode);
  GRPO_TRY(check_seq_args(mask, mask_dtype, bsz, t_len));
  if (bsz == 0) return 0;
  GRPO_CUDA(cudaMemsetAsync(acc_scratch, 0, sizeof(double), stream));
  if (t_len > 0) {
    kl_reward_kernel<<<cdiv(bsz * 32, 256), 256, 0, stream>>>(token_level_scores, logp, ref_logp, mask, mask_dtype,
                                                              static_cast<uint32_t>(bsz), static_cast<uint32_t>(t_len),
                                                              kl_mode, kl_coef, token_level_rewards, acc_scratch);
    count_launch();
  }
  kl_reward_finalize_kernel<<<1, 32, 0, stream>>>(acc_scratch, static_cast<uint32_t>(bsz), current_kl);
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

size_t grpo_compact_scratch_bytes(int64_t n) {
  return n <= 0 ? sizeof(int32_t) : (static_cast<size_t>((n + kCompactBlock - 1) / kCompactBlock) + 1) * sizeof(int32_t);
}

int grpo_compact_index(const void* mask, int mask_dtype, int64_t n, int32_t* gather_idx, int32_t* inverse,
                       int32_t* count, void* scratch, size_t scratch_bytes, grpo_stream_t stream) {
  if (!mask || !gather_idx || !inverse || !count || !scratch)
    return fail(GRPO_ERR_ARG, "mask / gather_idx / inverse / count / scratch must not be null");
  if (mask_dtype < 0 || mask_dtype > 2) return fail(GRPO_ERR_ARG, "bad mask_dtype");
  if (n < 0 || n > 0x7fffffffll) return fail(GRPO_ERR_ARG, "bad length");
  if (scratch_bytes < grpo_compact_scratch_bytes(n)) return fail(GRPO_ERR_WORKSPACE, "compaction scratch too small");
  auto* blocks = static_cast<int32_t*>(scratch);
  const uint32_t nb = cdiv(n, kCompactBlock);
  if (nb > 0) {
    compact_count_kernel<<<nb, kCompactThreads, 0, stream>>>(mask, mask_dtype, static_cast<size_t>(n), blocks);
    count_launch();
  }
  compact_scan_kernel<<<1, 1024, 0, stream>>>(blocks, nb, count);
  count_launch();
  if (nb > 0) {
    compact_write_kernel<<<nb, kCompactThreads, 0, stream>>>(mask, mask_dtype, static_cast<size_t>(n), blocks,
                                                             gather_idx, inverse);
    count_launch();
  }
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

static int check_rows_args(const void* in, const int32_t* idx, void* out, int64_t n, int64_t row_bytes) {
  if (!in || !idx || !out) return fail(GRPO_ERR_ARG, "in / index / out must not be null");
  if (n < 0 || n > 0x7fffffffll || row_bytes <= 0 || row_bytes % 4 != 0 || row_bytes > (1ll << 30))
    return fail(GRPO_ERR_ARG, "bad row count / row_bytes (must be a positive multiple of 4)");
  return 0;
}
static inline bool rows_vec16(const void* in, const void* out, int64_t row_bytes) {
  return row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
}
static inline uint32_t rows_grid(int64_t rows, uint32_t vecs_per_row) {
  if (vecs_per_row >= 32) return cdiv(rows * 32, 256);
  return ew_blocks(rows * vecs_per_row, 256, 0);
}

int grpo_gather_rows(const void* in, const int32_t* gather_idx, int64_t m, int64_t row_bytes, void* out,
                     grpo_stream_t stream) {
  GRPO_TRY(check_rows_args(in, gather_idx, out, m, row_bytes));
  if (m == 0) return 0;
  if (rows_vec16(in, out, row_bytes)) {
    const uint32_t vpr = static_cast<uint32_t>(row_bytes / 16);
    gather_rows_kernel<uint4><<<rows_grid(m, vpr), 256, 0, stream>>>(
        static_cast<const uint4*>(in), gather_idx, static_cast<size_t>(m), vpr, static_cast<uint4*>(out));
  } else {
    const uint32_t vpr = static_cast<uint32_t>(row_bytes / 4);
    gather_rows_kernel<uint32_t><<<rows_grid(m, vpr), 256, 0, stream>>>(
        static_cast<const uint32_t*>(in), gather_idx, static_cast<size_t>(m), vpr, static_cast<uint32_t*>(out));
  }
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_scatter_rows(const void* in, const int32_t* inverse, int64_t n, int64_t row_bytes, void* out,
                      grpo_stream_t stream) {
  GRPO_TRY(check_rows_args(in, inverse, out, n, row_bytes));
  if (n == 0) return 0;
  if (rows_vec16(in, out, row_bytes)) {
    const uint32_t vpr = static_cast<uint32_t>(row_bytes / 16);
    scatter_rows_kernel<uint4><<<rows_grid(n, vpr), 256, 0, stream>>>(
        static_cast<const uint4*>(in), inverse, static_cast<size_t>(n), vpr, static_cast<uint4*>(out));
  } else {
    const uint32_t vpr = static_cast<uint32_t>(row_bytes / 4);
    scatter_rows_kernel<uint32_t><<<rows_grid(n, vpr), 256, 0, stream>>>(
        static_cast<const uint32_t*>(in), inverse, static_cast<size_t>(n), vpr, static_cast<uint32_t*>(out));
  }
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_logprob_from_logits(const void* logits, int logits_dtype, const int64_t* labels, int64_t rows, int64_t vocab,
                             int64_t ld, float* logp, float* entropy, float* lse, grpo_stream_t stream) {
  if (rows == 0) return 0;
  if (!logits || (logp && !labels)) return fail(GRPO_ERR_ARG, "logits (and labels when logp is wanted) must not be null");
  if (rows < 0 || vocab <= 0 || ld < vocab || rows > 0x7fffffffll || vocab > 0x7fffffffll)
    return fail(GRPO_ERR_ARG, "bad dimensions");
  const uint32_t r = static_cast<uint32_t>(rows), v = static_cast<uint32_t>(vocab);
  const int threads = vocab >= 8192 ? 256 : 128;  // eight resident row-blocks per SM hide each other's reduction tails
#define GRPO_LOGITS_FWD(T, E)                                                                                       \
  logprob_from_logits_kernel<T, E><<<r, threads, 0, stream>>>(static_cast<const T*>(logits), labels, r, v, ld, logp, \
                                                              entropy, lse)
  switch (logits_dtype) {
    case LOGITS_F32:
      if (entropy) GRPO_LOGITS_FWD(float, true); else GRPO_LOGITS_FWD(float, false);
      break;
    case LOGITS_BF16:
      if (entropy) GRPO_LOGITS_FWD(__nv_bfloat16, true); else GRPO_LOGITS_FWD(__nv_bfloat16, false);
      break;
    case LOGITS_F16:
      if (entropy) GRPO_LOGITS_FWD(__half, true); else GRPO_LOGITS_FWD(__half, false);
      break;
    default: return fail(GRPO_ERR_ARG, "unknown logits dtype %d", logits_dtype);
  }
#undef GRPO_LOGITS_FWD
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_logprob_from_logits_bwd(const void* logits, int logits_dtype, const int64_t* labels, const float* lse,
                                 const float* dlogp, const float* dentropy, const float* entropy, int64_t rows,
                                 int64_t vocab, int64_t ld, void* dlogits, int64_t ld_out, grpo_stream_t stream) {
  if (rows == 0) return 0;
  if (!logits || !lse || !dlogits) return fail(GRPO_ERR_ARG, "logits / lse / dlogits must not be null");
  if (dlogp && !labels) return fail(GRPO_ERR_ARG, "dlogp needs labels");
  if (dentropy && !entropy) return fail(GRPO_ERR_ARG, "dentropy needs the forward entropy");
  if (rows < 0 || vocab <= 0 || ld < vocab || ld_out < vocab || rows > 0x7fffffffll || vocab > 0x7fffffffll)
    return fail(GRPO_ERR_ARG, "bad dimensions");
  if (rows == 0) return 0;
  const uint32_t r = static_cast<uint32_t>(rows), v = static_cast<uint32_t>(vocab);
  const size_t esz = logits_dtype == LOGITS_F32 ? 4 : 2;
  const int64_t vec = 16 / static_cast<int64_t>(esz);
  const bool vector = vocab % vec == 0 && ld % vec == 0 && ld_out % vec == 0 &&
                      (reinterpret_cast<uintptr_t>(logits) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dlogits) & 15u) == 0;
  // four 16-byte vectors (or 8 scalars) per thread: the per-row constants are amortised, the grid stays in the 1e5 range
  dim3 grid(cdiv(vocab, 256 * (vector ? 4 * vec : 8)), r < 65535u ? r : 65535u);
#define GRPO_LOGITS_BWD(T, V)                                                                                         \
  logprob_from_logits_bwd_kernel<T, V><<<grid, 256, 0, stream>>>(static_cast<const T*>(logits), labels, lse, dlogp,    \
                                                                 dentropy, entropy, r, v, ld, static_cast<T*>(dlogits), \
                                                                 ld_out)
  switch (logits_dtype) {
    case LOGITS_F32:
      if (vector) GRPO_LOGITS_BWD(float, true); else GRPO_LOGITS_BWD(float, false);
      break;
    case LOGITS_BF16:
      if (vector) GRPO_LOGITS_BWD(__nv_bfloat16, true); else GRPO_LOGITS_BWD(__nv_bfloat16, false);
      break;
    case LOGITS_F16:
      if (vector) GRPO_LOGITS_BWD(__half, true); else GRPO_LOGITS_BWD(__half, false);
      break;
    default: return fail(GRPO_ERR_ARG, "unknown logits dtype %d", logits_dtype);
  }
#undef GRPO_LOGITS_BWD
  count_launch();
  GRPO_CUDA(cudaGetLastError());
  return 0;
}

int grpo_debug_plan_units(int64_t tiles, int64_t k_blocks, int groups_avail, int split_mode, int32_t* units,
                          int64_t max_units, int32_t* info) {
  if (tiles <= 0 || k_blocks <= 0 || groups_avail <= 0 || tiles > 0x7fffffffll || k_blocks > 0x7fffffffll)
    return fail(GRPO_ERR_ARG, "bad dimensions");
  if (split_mode < 0 || split_mode > 2) return fail(GRPO_ERR_ARG, "split_mode must be 0, 1 or 2");
  TileSched s{};
  s.k_blocks = static_cast<uint32_t>(k_blocks);
  s.split_tail = static_cast<uint32_t>(split_mode);
  const uint32_t groups = plan_units(s, static_cast<uint32_t>(tiles), static_cast<uint32_t>(groups_avail));
  if (info) {
    info[0] = static_cast<int32_t>(groups);
    info[1] = static_cast<int32_t>(s.num_units);
    info[2] = static_cast<int32_t>(s.max_progress);
    info[3] = static_cast<int32_t>(s.split_slices);
  }
  if (units) {
    if (static_cast<int64_t>(s.num_units) > max_units) return fail(GRPO_ERR_WORKSPACE, "units buffer too small");
    for (uint32_t u = 0; u < s.num_units; ++u) {
      uint32_t t, kb0, kb1;
      decode_unit(s, u, t, kb0, kb1);
      units[3 * u] = static_cast<int32_t>(t);
      units[3 * u + 1] = static_cast<int32_t>(kb0);
      units[3 * u + 2] = static_cast<int32_t>(kb1);
    }
  }
  return 0;
}

int grpo_debug_gemm(const void* a, const void* b, float* c, int64_t m, int64_t n, int64_t k, int a_mn_major,
                    int b_mn_major, int cta_group, int accumulate, grpo_stream_t stream) {
  if (!a || !b || !c) return fail(GRPO_ERR_ARG, "null operand");
  if (m <= 0 || n <= 0 || k <= 0 || n % 4 != 0) return fail(GRPO_ERR_ARG, "bad dimensions (n must be a multiple of 4)");
  if (cta_group != 1 && cta_group != 2) return fail(GRPO_ERR_ARG, "cta_group must be 1 or 2");
  DevInfo dev;
  GRPO_TRY(get_dev(&dev));
  dev.cta_group = cta_group;  // explicit for the debug entry; the tile shape (ksub) follows the process-wide knob
  EpiF32<1, kBlockN>::Params p1;
  memset(&p1, 0, sizeof(p1));
  p1.c = c;
  p1.ldc = n;
  p1.m = static_cast<uint32_t>(m);
  p1.n = static_cast<uint32_t>(n);
  p1.accumulate = static_cast<uint32_t>(accumulate != 0);
  p1.policy = kEvictNormal;
  p1.use_tma = dev.dw_tma ? 1u : 0u;  // the debug entry exercises whichever fp32 epilogue the process-wide knob selects
  if (p1.use_tma) GRPO_TRY(make_tmap_f32_out(&p1.c_map, c, static_cast<uint64_t>(n), static_cast<uint64_t>(m),
                                             static_cast<uint64_t>(n)));
  EpiF32<2, kBlockN>::Params p2;
  memcpy(&p2, &p1, sizeof(p1));
  TileSched s{};
  s.m_fast = 1;
  s.split_tail = (accumulate != 0) ? static_cast<uint32_t>(dev.dw_split) : 0u;  // 2: the multi-round plan (dHidden split path)
  // A: 0 = [m][k], 1 = [k][m], 2 = blocked [m/64][k/64][64][64], 3 = blocked [k/64][m/64][64 k][64 m]
  const uint64_t a_pitch = a_mn_major == 0 ? k : (a_mn_major == 1 ? m : (a_mn_major == 2 ? (k + 63) / 64 : (m + 63) / 64));
  const uint64_t b_pitch = b_mn_major ? n : k;
  using E1 = EpiF32<1, kBlockN>;
  using E2 = EpiF32<2, kBlockN>;
#define GRPO_DBG(AM, BM) return launch_gemm_any<AM, BM, E1, E2>(dev, a, m, a_pitch, b, n, b_pitch, k, s, p1, p2, stream)
  switch (a_mn_major * 2 + (b_mn_major ? 1 : 0)) {
    case 0: GRPO_DBG(A_K_MAJOR, false);
    case 1: GRPO_DBG(A_K_MAJOR, true);
    case 2: GRPO_DBG(A_MN_MAJOR, false);
    case 3: GRPO_DBG(A_MN_MAJOR, true);
    case 4: GRPO_DBG(A_BLOCKED_K, false);
    case 5: GRPO_DBG(A_BLOCKED_K, true);
    case 6: GRPO_DBG(A_BLOCKED_MN, false);
    case 7: GRPO_DBG(A_BLOCKED_MN, true);
    default: return fail(GRPO_ERR_ARG, "a_mn_major must be 0..3");
  }
#undef GRPO_DBG
}

}  // extern "C"
