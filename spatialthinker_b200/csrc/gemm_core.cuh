// Persistent, warp-specialised tcgen05 GEMM main loop for sm_100a (bf16 x bf16 -> fp32 in TMEM).
//
//   warp 0      : TMA producer (one lane) - cp.async.bulk.tensor into a kStages-deep smem ring
//   warp 1      : MMA issuer (one lane, pair leader only) - tcgen05.mma, accumulators double-buffered in TMEM
//   warp 2      : TMEM allocator
//   warps 4..7  : epilogue - tcgen05.ld; thread i of the warpgroup owns accumulator row i (TMEM lane i)
//
// kCta == 2 runs one 256 x BLOCK_N tile on a CTA pair (cta_group::2): each CTA loads its own 128 rows of A and half
// of the B tile, the leader issues the MMAs for both, each CTA drains its own 128 accumulator rows.
//
// Operand majorness is a template parameter: K-major operands are [rows][K] row-major in global memory, MN-major
// operands are [K][rows] row-major (the "transposed" reads the backward GEMMs need) - both are fed by plain 2-D TMA
// boxes in 128-byte-swizzle layout, so no transposed copies are ever materialised in HBM. The A operand can also live
// in a BLOCKED layout [blk][blk][64][64] (8 KB contiguous per 64 x 64 block): one 4-D TMA box then fetches the same
// smem image as the 2-D boxes would, from contiguous HBM, whichever of the two dimensions the GEMM treats as K.
#pragma once
#include "ptx.cuh"

namespace grpo {

constexpr int kBlockM = 128;  // accumulator rows per CTA (= TMEM lanes)
constexpr int kBlockK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiThreads = 128;

// A-operand storage / majorness (bit 0 = MN-major for the MMA descriptor)
enum AMode : int {
  A_K_MAJOR = 0,     // [rows][K] row-major
  A_MN_MAJOR = 1,    // [K][rows] row-major
  A_BLOCKED_K = 2,   // [rows/64][K/64][64 rows][64 k]   : K-major tile image
  A_BLOCKED_MN = 3,  // [K/64][rows/64][64 k][64 rows]   : MN-major tile image (the same buffer, other dimension as K)
};

// Tile walk: tiles are grouped in panels of `panel_m` row-blocks; inside a panel either the row-block index (m_fast)
// or the column-block index runs fastest. Chosen per GEMM so that the operand re-read by neighbouring CTAs is the
// one that stays resident in L2.
struct TileSched {
  uint32_t m_blocks;  // row blocks of 128*kCta
  uint32_t n_blocks;  // column blocks of BLOCK_N
  uint32_t k_blocks;  // K blocks of 64
  uint32_t panel_m;   // row blocks per panel
  uint32_t m_fast;
  // Optional lock-step: every `sync_period` K-blocks of progress all CTA groups meet at a global counter, so tiles that
  // share an operand panel read it from L2 while it is still resident (persistent CTAs otherwise drift apart, every
  // group then streams its own copy from HBM, and the extra DRAM traffic costs SM clock under the power cap).
  uint32_t sync_period;  // in K-blocks; 0 = free running
  uint32_t* sync_ctr;    // zeroed by the host before the launch
  // L2 eviction priority of the two operand streams (kEvictNormal / kEvictFirst / kEvictLast)
  uint64_t hint_a, hint_b;
  uint32_t wait_hint_ns;  // suspend-time hint of the epilogue warps' mbarrier waits (0: poll)
  // measurement aid: block 0 writes {clock64, globaltimer} at entry and exit (4 x u64) so that the SM clock a kernel
  // actually ran at inside a long pipeline can be read back (nullptr: off)
  unsigned long long* probe;
  uint32_t acc_lead;      // wide tile: K-blocks accumulator 0 runs ahead of accumulator 1 at the tile ends (0..kStages-1)
  // Split-K tail (accumulating epilogues only; filled in by launch_gemm). The persistent groups walk `num_units` work
  // units: the first `whole_tiles` are whole tiles; when the tile count is not a multiple of the group count, the
  // remaining tiles - which would occupy a fraction of the machine for a full tile time - are cut along K into
  // `split_slices` units of `split_kpb` K-blocks each, so the last round is short and (nearly) full.
  // Wide tile, epilogues that support it (Epi::kCanShare): both epilogue warpgroups drain accumulator 0 first (half of
  // its columns each), then accumulator 1, instead of one warpgroup per accumulator. Accumulator 0 - which the MMA warp
  // finishes `acc_lead` K-blocks early - is then free again after half the drain time, and the next tile's first MMAs
  // overlap the drain of accumulator 1.
  uint32_t epi_share;
  uint32_t split_tail;    // request (caller): 0 = off, 1 = one short extra round (dW GEMM), 2 = best slice count over
                          // several short rounds, slice-major (dHidden GEMM on its fp32 split path)
  uint32_t num_units, whole_tiles, split_slices, split_kpb;
  uint32_t split_tiles;   // 0: tail units are ordered tile-major (all slices of a tile are neighbours); otherwise the
                          // number of tail tiles, and the units are ordered slice-major: the groups that run side by
                          // side work on the SAME K range of neighbouring tiles, so the operand panels those tiles share
                          // are fetched from HBM once, as in the whole-tile rounds
  uint32_t max_progress;  // most K-blocks any one group loads (progress-window bookkeeping)
  // Serpentine K order: every other whole tile of a CTA group walks its K-blocks backwards. All groups are in the same
  // round at the same time (and in lock-step within it), so the operand that EVERY round re-reads along K - the scaled
  // hidden chunk in the dW GEMM (136 MB, just over the L2), W in the dHidden GEMM - is met again first where it was
  // touched last: the tail of one round is still L2-resident at the head of the next (a cyclic walk evicts exactly what
  // it needs next). Each tile still sums its K-blocks in a fixed order: results stay reproducible.
  uint32_t serpentine;
};

// work unit u -> (tile, K-block range)
__host__ __device__ __forceinline__ void decode_unit(const TileSched& s, uint32_t u, uint32_t& tile, uint32_t& kb0, uint32_t& kb1) {
  if (u < s.whole_tiles) {
    tile = u;
    kb0 = 0;
    kb1 = s.k_blocks;
  } else {
    const uint32_t j = u - s.whole_tiles;
    if (s.split_tiles != 0) {  // slice-major
      const uint32_t sl = j / s.split_tiles;
      tile = s.whole_tiles + (j - sl * s.split_tiles);
      kb0 = sl * s.split_kpb;
    } else {
      const uint32_t q = j / s.split_slices;
      tile = s.whole_tiles + q;
      kb0 = (j - q * s.split_slices) * s.split_kpb;
    }
    kb1 = (kb0 + s.split_kpb < s.k_blocks) ? kb0 + s.split_kpb : s.k_blocks;
  }
}

// Progress window `round` (1-based) starts: announce it, then make sure every one of the `groups` producers has at least
// STARTED the previous window. That bounds the skew between CTA groups to two windows without ever stalling a producer
// that is merely level with the others (a full barrier here costs ~10 % of tensor-pipe time, measured). A producer that
// waits too long stops waiting for the rest of the kernel (it still announces, so nobody can hang on it).
__device__ __forceinline__ void progress_sync(uint32_t* ctr, uint32_t round, uint32_t groups, bool& give_up) {
  atomicAdd(ctr, 1u);
  if (give_up || round < 2) return;
  const uint32_t target = (round - 1) * groups;
  const long long t0 = clock64();
  while (true) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    if (clock64() - t0 > 400000) {  // ~0.25 ms: someone is not co-resident; run free from here on
      give_up = true;
      break;
    }
    __nanosleep(64);
  }
}

__device__ __forceinline__ void decode_tile(const TileSched& s, uint32_t t, uint32_t& m_blk, uint32_t& n_blk) {
  const uint32_t per_panel = s.panel_m * s.n_blocks;
  const uint32_t p = t / per_panel;
  const uint32_t r = t - p * per_panel;
  const uint32_t m0 = p * s.panel_m;
  const uint32_t pm = min(s.panel_m, s.m_blocks - m0);
  if (s.m_fast) {
    n_blk = r / pm;
    m_blk = m0 + (r - n_blk * pm);
  } else {
    const uint32_t mm = r / s.n_blocks;
    m_blk = m0 + mm;
    n_blk = r - mm * s.n_blocks;
  }
}

// Build with -DGRPO_TRACE (tools/trace_tiles.py; never the product build) to have CTA 0 log clock64 at the pipeline
// events of its first 16 tiles into the probe area: [256 + 16 * event + tile].
//   0/1 first MMA of accumulator 0/1 issued   2/3 last MMA of accumulator 0/1 issued
//   4/5 epilogue 0/1 saw its accumulator complete   6/7 epilogue 0/1 released its TMEM slot   8/9 epilogue 0/1 done
//   10  producer issued the tile's first load
#ifdef GRPO_TRACE
#define GRPO_TR(ev, tile)                                                                        \
  do {                                                                                           \
    if (sched.probe != nullptr && blockIdx.x == 0 && (tile) < 16u)                               \
      sched.probe[256 + 16 * (ev) + (tile)] = static_cast<unsigned long long>(clock64());        \
  } while (0)
#else
#define GRPO_TR(ev, tile) \
  do {                    \
  } while (0)
#endif

// What an epilogue sees for one 128 x BLOCK_N accumulator.
struct EpiCtx {
  uint32_t row;       // global output row owned by this thread (TMEM lane)
  uint32_t n_blk;     // column block index
  uint32_t col0;      // first global column of the accumulator
  uint32_t warp, lane;
  uint32_t tmem_acc;  // TMEM address of this warp's lanes, column 0 of the accumulator
  uint32_t epi_warp;  // index of this warp among the CTA's epilogue warps (its 4 KB staging buffer in smem_epi)
  uint32_t row0;      // first global output row of this warp (row - lane)
  uint32_t part;      // index of this call's (column block, column part) among a row's partial results
};
constexpr int kEpiStageBytes = 4096;  // per-epilogue-warp staging buffer for shared -> global bulk stores

// kSub = number of 128-row accumulators a CTA computes per tile (its share of the tile is 128 * kSub rows):
//   kSub == 1: 128 x BLOCK_N per CTA; the two TMEM accumulator slots double-buffer consecutive tiles, so the epilogue
//              of tile t fully overlaps the MMAs of tile t+1. Shared-memory traffic per MMA cycle is (128 + BLOCK_N /
//              kCta) operand rows in plus the MMA's own operand reads - on a CTA pair exactly the 128 B/clk the SM can
//              move, which caps the tensor pipe near 84 % (measured, profiles/).
//   kSub == 2: 256 x BLOCK_N per CTA; both TMEM slots are live at once, every B stage is used by two MMAs, operand
//              traffic per MMA cycle drops by a quarter (the tile shape cuBLAS picks) and the main loop can run the
//              tensor pipe flat out; the price is that the epilogue (two warpgroups, one per accumulator) no longer
//              hides behind the next tile's MMAs.
template <int kCta, int kSub, int BLOCK_N, int kStages>
struct GemmCfg {
  static constexpr int kRowsPerCta = kBlockM * kSub;
  static constexpr int kTileRows = kRowsPerCta * kCta;
  static constexpr int kLoadN = BLOCK_N / kCta;  // B columns each CTA loads
  static constexpr int kABytes = kRowsPerCta * kBlockK * 2;
  static constexpr int kBBytes = kLoadN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kEpiWarps = 4 * kSub;
  static constexpr int kThreads = 128 + 32 * kEpiWarps;
  static_assert(kSub == 1 || kSub == 2, "one or two accumulators per CTA");
  static_assert(kTmemCols == 512 || kTmemCols == 256 || kTmemCols == 128 || kTmemCols == 64, "TMEM columns");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N <= 256, "UMMA N");
  static constexpr size_t smem_bytes(int epi_bytes) { return 1024 + kRingBytes + epi_bytes + kBarBytes; }
};

template <int kCta, int kSub, int BLOCK_N, int kStages, int kAMode, bool kBMn, class Epi>
__global__ void __launch_bounds__(128 + 128 * kSub, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const TileSched sched, const __grid_constant__ typename Epi::Params ep) {
  using Cfg = GemmCfg<kCta, kSub, BLOCK_N, kStages>;
  constexpr bool kAMn = (kAMode & 1) != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_epi = smem + Cfg::kRingBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes + Epi::kSmemBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = (kCta == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  // warp 3 has no other role. Layout: [0..3] block 0 {clock64, ns} at entry / exit; [8 + 2g], [9 + 2g] entry / exit ns of
  // CTA group g (how far apart the persistent groups start and finish)
  const bool probing = sched.probe != nullptr && threadIdx.x == 96 && rank == 0;
  if (probing) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    if (blockIdx.x == 0) {
      sched.probe[0] = static_cast<unsigned long long>(clock64());
      sched.probe[1] = ns;
    }
    if (blockIdx.x / kCta < 120) sched.probe[8 + 2 * (blockIdx.x / kCta)] = ns;
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], kCta);  // one arrive per CTA producer (+ transaction bytes), on the leader
      mbar_init(&empty_bar[i], 1);    // one tcgen05.commit (multicast to both CTAs)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);                        // one tcgen05.commit
      // the warpgroup(s) draining this slot, in every CTA of the pair
      mbar_init(&tmem_empty[i], kCta * kNumEpiThreads * ((kSub == 2 && Epi::kCanShare && sched.epi_share) ? 2 : 1));
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (kCta == 2) cluster_sync_all();  // peer must be resident before a pair-wide TMEM allocation
  if (warp == 2) tmem_alloc<kCta>(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  if (kCta == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t num_units = sched.num_units;
  const uint32_t first_tile = blockIdx.x / kCta;
  const uint32_t tile_step = gridDim.x / kCta;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const bool do_sync = leader && sched.sync_period != 0 && sched.sync_ctr != nullptr;
      uint32_t progress = 0, rounds_done = 0;
      bool give_up = false;
      uint32_t round = 0;
      for (uint32_t u = first_tile; u < num_units; u += tile_step, ++round) {
        uint32_t t, kb0, kb1, m_blk, n_blk;
        decode_unit(sched, u, t, kb0, kb1);
        decode_tile(sched, t, m_blk, n_blk);
        const int32_t m0 = static_cast<int32_t>(m_blk * Cfg::kTileRows + rank * Cfg::kRowsPerCta);
        const int32_t n0 = static_cast<int32_t>(n_blk * BLOCK_N + rank * Cfg::kLoadN);
        const bool backwards = sched.serpentine != 0 && (round & 1u) != 0 && u < sched.whole_tiles;
        for (uint32_t step = kb0; step < kb1; ++step, ++progress) {
          const uint32_t kb = backwards ? kb0 + (kb1 - 1 - step) : step;  // the K-block this stage carries
          if (do_sync && progress != 0 && progress % sched.sync_period == 0)
            progress_sync(sched.sync_ctr, ++rounds_done, tile_step, give_up);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (step == kb0) GRPO_TR(10, (u - first_tile) / tile_step);
          const int32_t k0 = static_cast<int32_t>(kb * kBlockK);
          uint8_t* sa = smem_a + stage * Cfg::kABytes;
          uint8_t* sb = smem_b + stage * Cfg::kBBytes;
          if constexpr (kAMode == A_K_MAJOR) {
            tma_load_2d<kCta>(&tmap_a, &full_bar[stage], sa, k0, m0, sched.hint_a);
          } else if constexpr (kAMode == A_MN_MAJOR) {
#pragma unroll
            for (int i = 0; i < Cfg::kRowsPerCta / 64; ++i)
              tma_load_2d<kCta>(&tmap_a, &full_bar[stage], sa + i * (kBlockK * 128), m0 + i * 64, k0, sched.hint_a);
          } else if constexpr (kAMode == A_BLOCKED_K) {  // box (64, 64, 1 K-block, kRowsPerCta / 64 row blocks)
            tma_load_4d<kCta>(&tmap_a, &full_bar[stage], sa, 0, 0, static_cast<int32_t>(kb), m0 >> 6, sched.hint_a);
          } else {  // A_BLOCKED_MN: box (64, 64, kRowsPerCta / 64 row blocks, 1 K-block)
            tma_load_4d<kCta>(&tmap_a, &full_bar[stage], sa, 0, 0, m0 >> 6, static_cast<int32_t>(kb), sched.hint_a);
          }
          if constexpr (!kBMn) {
            tma_load_2d<kCta>(&tmap_b, &full_bar[stage], sb, k0, n0, sched.hint_b);
          } else {
#pragma unroll
            for (int i = 0; i < Cfg::kLoadN / 64; ++i)
              tma_load_2d<kCta>(&tmap_b, &full_bar[stage], sb + i * (kBlockK * 128), n0 + i * 64, k0, sched.hint_b);
          }
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes * kCta);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (do_sync) {  // groups with less work keep arriving so the others never wait on them
        const uint32_t total = sched.max_progress;
        const uint32_t rounds = total == 0 ? 0 : (total - 1) / sched.sync_period;
        for (; rounds_done < rounds; ++rounds_done) atomicAdd(sched.sync_ctr, 1u);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair leader only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM * kCta, BLOCK_N, kAMn, kBMn);
      constexpr uint32_t a_lbo = kAMn ? kBlockK * 128 : 0, b_lbo = kBMn ? kBlockK * 128 : 0;
      constexpr uint32_t a_kstep = kAMn ? kUmmaK * 128 : kUmmaK * 2, b_kstep = kBMn ? kUmmaK * 128 : kUmmaK * 2;
      constexpr uint32_t a_sub = kBlockM * 128;  // 128 rows further on: 16 KB in either majorness (two 8 KB MN atoms)
      uint32_t it = 0, base = 0;  // base: K-blocks issued before this unit (ring position of its first K-block)
      uint32_t kblocks = 0;       // K-blocks of the current unit; `kb` below counts from the unit's first K-block
      for (uint32_t u = first_tile; u < num_units; u += tile_step, ++it, base += kblocks) {
        {
          uint32_t t_, kb0_, kb1_;
          decode_unit(sched, u, t_, kb0_, kb1_);
          kblocks = kb1_ - kb0_;
        }
        // accumulator slots: kSub == 1 alternates them tile by tile, kSub == 2 uses both for every tile
        const uint32_t slot0 = (kSub == 1) ? (it & 1) : 0u;
        const uint32_t ap = (kSub == 1) ? ((it >> 1) & 1) : (it & 1);
        // the MMAs of accumulator `sub` for K-block `kb` of this tile
        auto issue = [&](uint32_t sub, uint32_t kb) {
          const uint32_t cnt = base + kb, stage = cnt % kStages, phase = (cnt / kStages) & 1;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb == 0) {  // the epilogue of the tile that used this slot last must have drained it
            mbar_wait(&tmem_empty[slot0 + sub], ap ^ 1);
            tc_fence_after();
          }
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::kBBytes);
          if (kb == 0) GRPO_TR(sub, it);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = make_smem_desc(a_addr + sub * a_sub + k * a_kstep, 1024, a_lbo);
            const uint64_t db = make_smem_desc(b_addr + k * b_kstep, 1024, b_lbo);
            umma_bf16<kCta>(tmem_base + (slot0 + sub) * BLOCK_N, da, db, idesc, (kb | k) != 0);
          }
          if (kb + 1 == kblocks) {
            umma_commit<kCta>(&tmem_full[slot0 + sub]);  // this accumulator is complete
            GRPO_TR(2 + sub, it);
          }
        };
        // frees the smem slot of K-block `kb` (both CTAs) when the MMAs issued so far retire
        auto free_stage = [&](uint32_t kb) { umma_commit<kCta>(&empty_bar[(base + kb) % kStages]); };
        if constexpr (kSub == 1) {
          for (uint32_t kb = 0; kb < kblocks; ++kb) {
            issue(0, kb);
            free_stage(kb);
          }
        } else {
          // Accumulator 0 runs `lead` K-blocks ahead of accumulator 1 at both ends of the tile (in between they share
          // every B stage back to back): its epilogue starts - and its slot is free again for the next tile - that much
          // earlier, so the two drains overlap the other accumulator's MMAs instead of both idling the tensor pipe.
          // The ring simply holds `lead` stages a little longer at the two ends.
          uint32_t lead = sched.acc_lead < kStages - 1 ? sched.acc_lead : kStages - 1;
          if (kblocks < 2 * lead) lead = 0;
          for (uint32_t kb = 0; kb < lead; ++kb) issue(0, kb);
          for (uint32_t kb = 0; kb < lead; ++kb) {
            issue(1, kb);
            free_stage(kb);
          }
          for (uint32_t kb = lead; kb + lead < kblocks; ++kb) {
            issue(0, kb);
            issue(1, kb);
            free_stage(kb);
          }
          for (uint32_t kb = kblocks - lead; kb < kblocks; ++kb) issue(0, kb);
          for (uint32_t kb = kblocks - lead; kb < kblocks; ++kb) {
            issue(1, kb);
            free_stage(kb);
          }
        }
      }
      // the peer's last remote arrivals must land before this CTA's barriers go away
      if (kCta == 2 && it > 0) {
        const uint32_t last = it - 1;
        if (kSub == 1) {
          mbar_wait(&tmem_empty[last & 1], (last >> 1) & 1);
        } else {
          mbar_wait(&tmem_empty[0], last & 1);
          mbar_wait(&tmem_empty[1], last & 1);
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------------ epilogue: one warpgroup per accumulator
    const uint32_t ew = warp - kEpiWarp0;
    const uint32_t grp = ew >> 2, quad = ew & 3;  // warp (4 + ew) may touch TMEM lanes 32 * (ew % 4) .. +31
    uint32_t it = 0;
    for (uint32_t u = first_tile; u < num_units; u += tile_step, ++it) {
      uint32_t t, kb0_, kb1_;
      decode_unit(sched, u, t, kb0_, kb1_);
      if constexpr (kSub == 2 && Epi::kCanShare) {
        if (sched.epi_share) {  // both warpgroups on accumulator 0, then both on accumulator 1; `grp` picks the columns
          uint32_t m_blk, n_blk;
          decode_tile(sched, t, m_blk, n_blk);
#pragma unroll 1
          for (uint32_t slot = 0; slot < 2; ++slot) {
            EpiCtx c;
            c.n_blk = n_blk;
            c.part = n_blk * 2 + grp;
            c.warp = quad;
            c.lane = lane;
            c.epi_warp = ew;
            c.row0 = m_blk * Cfg::kTileRows + rank * Cfg::kRowsPerCta + slot * kBlockM + quad * 32;
            c.row = c.row0 + lane;
            c.col0 = n_blk * BLOCK_N + grp * (BLOCK_N / 2);
            c.tmem_acc = tmem_base + slot * BLOCK_N + grp * (BLOCK_N / 2) + ((quad * 32u) << 16);
            mbar_wait(&tmem_full[slot], it & 1, sched.wait_hint_ns);
            tc_fence_after();
            if (ew == 0 && lane == 0) GRPO_TR(4 + slot, it);
            auto release = [&]() {
              tc_fence_before();
              if (leader) mbar_arrive(&tmem_empty[slot]);
              else mbar_arrive_cluster(&tmem_empty[slot], 0);
              if (ew == 0 && lane == 0) GRPO_TR(6 + slot, it);
            };
            Epi::template run_cols<BLOCK_N / 2>(ep, c, smem_epi, release);
            if (ew == 0 && lane == 0) GRPO_TR(8 + slot, it);
          }
          continue;
        }
      }
      const uint32_t slot = (kSub == 1) ? (it & 1) : grp;
      const uint32_t ap = (kSub == 1) ? ((it >> 1) & 1) : (it & 1);
      EpiCtx c;
      uint32_t m_blk;
      decode_tile(sched, t, m_blk, c.n_blk);
      c.warp = quad;
      c.lane = lane;
      c.epi_warp = ew;
      c.row0 = m_blk * Cfg::kTileRows + rank * Cfg::kRowsPerCta + grp * kBlockM + quad * 32;
      c.row = c.row0 + lane;
      c.col0 = c.n_blk * BLOCK_N;
      c.tmem_acc = tmem_base + slot * BLOCK_N + ((quad * 32u) << 16);
      c.part = c.n_blk;
      mbar_wait(&tmem_full[slot], ap, sched.wait_hint_ns);
      tc_fence_after();
      if (quad == 0 && lane == 0) GRPO_TR(4 + (slot & 1), it);
      // the epilogue calls `release` as soon as its last TMEM load has completed, before it finishes the arithmetic
      // and the stores of that last column group: the MMA warp can refill the slot that much earlier
      auto release = [&]() {
        tc_fence_before();
        if (leader) mbar_arrive(&tmem_empty[slot]);
        else mbar_arrive_cluster(&tmem_empty[slot], 0);
        if (quad == 0 && lane == 0) GRPO_TR(6 + (slot & 1), it);
      };
      Epi::run(ep, c, smem_epi, release);
      if (quad == 0 && lane == 0) GRPO_TR(8 + (slot & 1), it);
    }
    Epi::finish(ep, lane);  // outstanding bulk stores of this warp
  }

  __syncwarp();  // re-converge the single-lane roles before the aligned barriers below
  tc_fence_before();
  if (kCta == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<kCta>(tmem_base, Cfg::kTmemCols);
  if (probing) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    if (blockIdx.x == 0) {
      sched.probe[2] = static_cast<unsigned long long>(clock64());
      sched.probe[3] = ns;
    }
    if (blockIdx.x / kCta < 120) sched.probe[9 + 2 * (blockIdx.x / kCta)] = ns;
  }
}

// ------------------------------------------------------------------------------------------
// Generic epilogues
// ------------------------------------------------------------------------------------------
// fp32 result, plain store or accumulate into C[M][ldc].
//   use_tma == 0: every thread stores / red.global.add's its own row, 16 bytes at a time (32 rows per warp instruction).
//   use_tma != 0: each warp stages 32 rows x 32 columns (4 KB, 128-byte swizzle, conflict-free st.shared) and one lane
//                 issues a bulk tensor store / reduce-add: the accumulation happens at L2 in whole 128-byte rows instead
//                 of 16-byte pieces in 32 different lines per instruction, and the SM's LSU sees none of it.
template <int kCta, int BLOCK_N>
struct EpiF32 {
  struct alignas(64) Params {
    CUtensorMap c_map;  // fp32 [m][n] (row pitch ldc), box 32 x 32, SWIZZLE_128B; valid iff use_tma
    float* c;
    int64_t ldc;
    uint32_t m, n;
    uint32_t accumulate;
    uint32_t use_tma;
    uint64_t policy;  // L2 eviction priority of the bulk reduce-add (kEvictNormal / kEvictFirst)
  };
  static constexpr int kSmemBytes = 8 * kEpiStageBytes;
  static constexpr bool kCanShare = false;
  __device__ static void finish(const Params& p, uint32_t lane) {
    if (p.use_tma && lane == 0) bulk_wait_all();
  }
  template <class Release>
  __device__ static void run(const Params& p, const EpiCtx& c, uint8_t* smem_epi, Release&& release) {
    const uint32_t row = c.row, col0 = c.col0;
    if (p.use_tma) {
      uint8_t* buf = smem_epi + c.epi_warp * kEpiStageBytes;
      uint8_t* myrow = buf + c.lane * 128;
      const uint32_t sw = c.lane & 7;
      const bool rows_live = c.row0 < p.m;  // warp-uniform; rows past m inside a live box are clipped by the TMA unit
      uint32_t va[32], vb[32];
      auto emit = [&](const uint32_t (&v)[32], int g) {
        const uint32_t col = col0 + g * 32;
        if (!rows_live || col >= p.n) return;
        if (c.lane == 0) bulk_wait_read_all();  // the previous copy out of this buffer has been read
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          st_shared_v4(myrow + ((q ^ sw) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (c.lane == 0) {
          if (p.accumulate) tma_reduce_add_2d(&p.c_map, buf, static_cast<int32_t>(col), static_cast<int32_t>(c.row0), p.policy);
          else tma_store_2d(&p.c_map, buf, static_cast<int32_t>(col), static_cast<int32_t>(c.row0));
          bulk_commit();
        }
      };
      static_assert((BLOCK_N / 32) % 2 == 0, "column groups are drained in pairs");
      tmem_ld_32x32(c.tmem_acc, va);
#pragma unroll 1
      for (int g = 0; g < BLOCK_N / 32; g += 2) {
        tmem_ld_wait();
        tmem_ld_32x32(c.tmem_acc + (g + 1) * 32, vb);
        emit(va, g);
        tmem_ld_wait();
        if (g + 2 < BLOCK_N / 32) tmem_ld_32x32(c.tmem_acc + (g + 2) * 32, va);
        else release();
        emit(vb, g + 1);
      }
      return;
    }
    float* out = p.c + static_cast<int64_t>(row) * p.ldc + col0;
#pragma unroll 1
    for (int g = 0; g < BLOCK_N / 32; ++g) {
      uint32_t v[32];
      tmem_ld_32x32(c.tmem_acc + g * 32, v);
      tmem_ld_wait();
      if (g == BLOCK_N / 32 - 1) release();
      if (row < p.m) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t col = col0 + g * 32 + q * 4;
          if (col < p.n) {  // n % 4 == 0
            float4 f = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                   __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            float4* dst = reinterpret_cast<float4*>(out + g * 32 + q * 4);
            if (p.accumulate) atomicAdd(dst, f);
            else *dst = f;
          }
        }
      }
    }
  }
};

// bf16 result into C[M][ldc]:  C[r][:] = row_scale[r] * acc[r][:] + onehot[r] * gather[idx[r]][:]
// (row_scale == nullptr: 1; onehot == nullptr: no gather term). The gather term is how the dHidden GEMM adds the
// one-hot part of dL/dlogits, (g_r / T) * W[label_r][:], without it ever being written into the stash.
template <int kCta, int BLOCK_N>
struct EpiBF16 {
  struct Params {
    __nv_bfloat16* c;
    int64_t ldc;
    uint32_t m, n;
    const float* row_scale;          // [m] or nullptr
    const float* onehot;             // [m] or nullptr
    const int64_t* idx;              // [m] row of `gather` to add (ignored where onehot[r] == 0)
    const __nv_bfloat16* gather;     // [*][ld_gather]
    int64_t ld_gather;
  };
  static constexpr int kSmemBytes = 0;
  static constexpr bool kCanShare = false;
  __device__ static void finish(const Params&, uint32_t) {}
  template <class Release>
  __device__ static void run(const Params& p, const EpiCtx& c, uint8_t*, Release&& release) {
    const uint32_t row = c.row, col0 = c.col0;
    const bool row_ok = row < p.m;
    __nv_bfloat16* out = p.c + static_cast<int64_t>(row) * p.ldc + col0;
    const float sc = (row_ok && p.row_scale) ? p.row_scale[row] : 1.f;
    const float oh = (row_ok && p.onehot) ? p.onehot[row] : 0.f;
    const __nv_bfloat16* grow = (oh != 0.f) ? p.gather + p.idx[row] * p.ld_gather + col0 : nullptr;
#pragma unroll 1
    for (int g = 0; g < BLOCK_N / 32; ++g) {
      uint32_t v[32];
      tmem_ld_32x32(c.tmem_acc + g * 32, v);
      tmem_ld_wait();
      if (g == BLOCK_N / 32 - 1) release();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t col = col0 + g * 32 + q * 8;
          if (col < p.n) {  // n % 8 == 0
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * q + i]) * sc;
            if (grow) {
              const uint4 w = *reinterpret_cast<const uint4*>(grow + g * 32 + q * 8);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
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
            *reinterpret_cast<uint4*>(out + g * 32 + q * 8) = pk;
          }
        }
      }
    }
  }
};

}  // namespace grpo
