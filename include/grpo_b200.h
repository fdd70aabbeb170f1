/* grpo_b200.h - C ABI of the B200-native GRPO policy-loss path.
 *
 * Every entry point replaces a piece of the reference's (hunarbatra/SpatialThinker, an EasyR1/veRL fork) Python hot
 * path; the reference has no FFI of its own (it is pure PyTorch), so the "binding" a maintainer adds is the ctypes
 * stub in INTEGRATION.md / spatialthinker_b200/_lib.py.  Cited lines are relative to the reference checkout.
 *
 * Conventions
 *   - all tensor arguments are DEVICE pointers (row-major, contiguous); the caller owns every buffer, the library
 *     never allocates or frees device memory and keeps no pointer after returning;
 *   - every call only ENQUEUES work on `stream` (no host synchronisation, no internal streams);
 *   - return value: 0 on success, a positive cudaError_t on a CUDA failure, a negative GRPO_ERR_* on bad arguments;
 *     grpo_last_error() returns a thread-local description; nothing throws across the ABI;
 *   - `mask_dtype`: 0 = float32, 1 = int64 (the reference's attention_mask slice), 2 = uint8/bool, 3 = no mask (all 1);
 *   - log-probabilities are log p (negative numbers) - the flash-attn branch of the reference, torch_functional.py:42.
 */
#ifndef GRPO_B200_H
#define GRPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* grpo_stream_t;

#define GRPO_ERR_ARG (-1)       /* shape / alignment / null-pointer violation */
#define GRPO_ERR_WORKSPACE (-2) /* workspace too small */
#define GRPO_ERR_DRIVER (-3)    /* TMA descriptor / driver entry point failure */

/* KL estimator selector - verl/trainer/core_algos.py:394-436 compute_kl(kl_penalty: str) */
#define GRPO_KL_NONE (-1)
#define GRPO_KL_LOW_VAR 0 /* "low_var_kl" (shipped default, scripts/config.yaml:22) */
#define GRPO_KL_KL 1      /* "kl"   */
#define GRPO_KL_ABS 2     /* "abs"  */
#define GRPO_KL_MSE 3     /* "mse"  */
#define GRPO_KL_CHI2 4    /* "chi2" */

/* metric slots written by the loss entry points (float[GRPO_NUM_METRICS], device memory) -
 * the "actor/..." keys of verl/workers/actor/dp_actor.py:274-286 */
#define GRPO_MET_PG_LOSS 0     /* masked_mean(policy loss), before the KL term   core_algos.py:349 */
#define GRPO_MET_CLIPFRAC_HI 1 /* actor/pg_clipfrac_higher                        core_algos.py:350 */
#define GRPO_MET_CLIPFRAC_LO 2 /* actor/pg_clipfrac_lower                         core_algos.py:351 */
#define GRPO_MET_PPO_KL 3      /* actor/ppo_kl                                    core_algos.py:352 */
#define GRPO_MET_KL_LOSS 4     /* actor/kl_loss                                   dp_actor.py:270   */
#define GRPO_MET_ENTROPY 5     /* actor/entropy_loss = -masked_mean(log_probs)    dp_actor.py:253   */
#define GRPO_MET_TOTAL 6       /* actor/pg_loss as logged: pg + kl_coef * kl      dp_actor.py:271   */
#define GRPO_MET_SCALED 7      /* total / grad_accum, the back-propagated scalar  dp_actor.py:277   */
#define GRPO_MET_TRUE_ENTROPY 8 /* masked_mean(lse - sum p z) when entropy was requested, else 0      */
#define GRPO_MET_MASK_SUM 9     /* sum(mask): valid tokens of the micro-batch                        */
#define GRPO_MET_SATURATED 10   /* unmasked tokens with log p <= -69.3 (= -100 ln 2): the fused head references its
                                 * softmax to the label's own logit and clamps exp2 arguments at 100, so the row sum of
                                 * such a token may be saturated; 0 for anything a policy could have sampled. The
                                 * reference (max-subtracted cross-entropy, torch_functional.py:45-66) has no such limit. */
#define GRPO_NUM_METRICS 11

int grpo_abi_version(void);
const char* grpo_last_error(void);

/* Measurement evidence (bench.py): number of this library's kernels enqueued so far by the process, and optional
 * CUDA-event timing of the pipeline phases on the caller's stream.
 *   phases: 0 logits GEMM (+softmax sums, exp stash), 1 row statistics (label logit + combine), 2 token loss,
 *           3 gradient preparation (row scales + one-hot scatter, or stash -> dlogits), 4 dHidden GEMM, 5 dW GEMM.   grpo_profile_read synchronises on the recorded events (call it outside the
 *           timed region); ms_out / count_out are [GRPO_NUM_PHASES] totals since the last reset. */
#define GRPO_NUM_PHASES 6
long long grpo_launch_count(void);
/* Tuning knobs for experiments (process-wide): "cta_group" 1|2, "fwd_panel" row blocks, "sync_fwd" / "sync_dh" /
 * "sync_dw" progress-barrier periods in K-blocks (0 = off), "l2_hints" 0|1, "dw_split" 0|1|2 (split-K tail of the dW GEMM;
 * 2 = the multi-round plan), "dh_split" 0|1 (dHidden GEMM: fp32 split-K path when its tiles are not a whole number of
 * rounds), "deterministic" 0|1 (run-to-run bit-reproducible gradients: no split-K, the one-hot rows of dW summed in row
 * order instead of with atomics). Defaults are the measured configuration. */
int grpo_set_option(const char* name, int value);
/* Measurement aid: with option "clk_probe" = 1 the three GEMM kernels of the chunk pipeline write
 * 1024 x uint64 each (order logits / dHidden / dW) at this byte offset of the caller's workspace:
 * [0..3] = block 0 {clock64 at entry, globaltimer ns at entry, clock64 at exit, globaltimer ns at exit} (cycles / ns = the
 * SM clock the kernel really ran at), [8 + 2g], [9 + 2g] = entry / exit ns of persistent CTA group g. */
size_t grpo_debug_probe_offset(int64_t rows, int64_t hidden_dim, int64_t vocab, int with_stash);
int grpo_profile_enable(int on);
int grpo_profile_read(double* ms_out, long long* count_out, int reset);

/* ------------------------------------------------------------------------------------------------------------------
 * lm_head -> token log-prob / entropy, forward only.
 * Replaces: HF `self.lm_head(hidden_states)` reached from dp_actor.py:118-125, `logits.div_(temperature)` :126 and
 * `log_probs_from_logits` torch_functional.py:45-66 - without materialising logits.
 *   hidden  bf16 [rows][hidden_dim]      weight bf16 [vocab][hidden_dim]      labels int64 [rows]
 *   logp    f32  [rows] (out)            entropy f32 [rows] (out, nullable: lse - sum p z)
 *   lse     f32  [rows] (out, nullable)
 * Requires hidden_dim % 64 == 0, vocab % 8 == 0, 16-byte aligned hidden / weight.
 * ------------------------------------------------------------------------------------------------------------------ */
size_t grpo_lmhead_fwd_workspace_bytes(int64_t rows, int64_t hidden_dim, int64_t vocab);
int grpo_lmhead_logprob_fwd(const void* hidden, const void* weight, const int64_t* labels, int64_t rows,
                            int64_t hidden_dim, int64_t vocab, float temperature, float* logp, float* entropy,
                            float* lse, void* workspace, size_t workspace_bytes, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * lm_head backward for arbitrary upstream gradients (the autograd of the call above).
 * Replaces: flash-attn CE backward + div_ backward + the two lm_head backward GEMMs triggered by `loss.backward()`
 * dp_actor.py:278.  Recomputes the logits tiles chunk by chunk, never holding more than one chunk's exp-stash.
 *   dlogp    f32 [rows]  dL/dlogp                       dentropy f32 [rows] dL/dentropy (nullable)
 *   dhidden  bf16 [rows][hidden_dim] (out, overwritten) dweight  f32 [vocab][hidden_dim] (ACCUMULATED into)
 * ------------------------------------------------------------------------------------------------------------------ */
size_t grpo_lmhead_bwd_workspace_bytes(int64_t rows, int64_t hidden_dim, int64_t vocab);
int grpo_lmhead_bwd(const void* hidden, const void* weight, const int64_t* labels, const float* dlogp,
                    const float* dentropy, int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature,
                    void* dhidden, float* dweight, void* workspace, size_t workspace_bytes, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused GRPO loss for ONE micro-batch: lm_head -> log-prob -> ratio / clip / dual-clip / KL / masked means ->
 * dL/dlogp -> dHidden, dW, in one pass (3 GEMM units, no recompute).
 * Replaces the body of the micro-batch loop, dp_actor.py:247-278 (+ compute_policy_loss core_algos.py:291-353,
 * compute_kl :394-436, masked_mean torch_functional.py:69-71).
 *   rows = micro_batch * response_length; old_logp / advantages / ref_logp (nullable) / mask are [rows]
 *   loss = (pg_loss + kl_coef * kl_loss - entropy_coef * masked_mean(entropy)) / grad_accum
 *   logp_out f32 [rows] (out)    entropy_out f32 [rows] (out, nullable unless entropy_coef != 0)
 *   dhidden bf16 [rows][hidden_dim] (out; nullable together with dweight => forward + metrics only)
 *   dweight f32 [vocab][hidden_dim] (ACCUMULATED into)       metrics f32 [GRPO_NUM_METRICS] (out)
 * ------------------------------------------------------------------------------------------------------------------ */
size_t grpo_fused_loss_workspace_bytes(int64_t rows, int64_t hidden_dim, int64_t vocab);
int grpo_fused_loss_fwd_bwd(const void* hidden, const void* weight, const int64_t* labels, const float* old_logp,
                            const float* advantages, const float* ref_logp, const void* mask, int mask_dtype,
                            int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature, float clip_ratio_low,
                            float clip_ratio_high, float clip_ratio_dual, int kl_mode, float kl_coef,
                            float entropy_coef, float grad_accum, float* logp_out, float* entropy_out, void* dhidden,
                            float* dweight, float* metrics, void* workspace, size_t workspace_bytes,
                            grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Deferred dW for SMALL micro-batches (the reference ships micro_batch_size_per_device_for_update = 4,
 * scripts/config.yaml; the gradient-accumulation loop is dp_actor.py:242-290). The weight gradient of the loop is a sum
 * over micro-batches, so its GEMM need not run once per micro-batch: the exp-stash workspace holds one chunk of
 * grpo_chunk_capacity_rows() rows anyway, and several small micro-batches can each run forward + loss + dHidden in
 * their own window ("slot") of it and share ONE dW GEMM over all collected rows - dW is then read-modify-written once
 * and the GEMM's K is long enough to hide its fp32 drain.
 *   grpo_fused_loss_fwd_bwd_slot: grpo_fused_loss_fwd_bwd for rows placed at [slot_row0, slot_row0 + rows) of a
 *     workspace of capacity_rows rows (capacity_rows % 512 == 0, <= grpo_chunk_capacity_rows(); slot_row0 % 512 == 0;
 *     workspace sized by grpo_fused_loss_workspace_bytes(capacity_rows, ...)). Everything except the stash-dependent
 *     part of dW is final on return (log-probs, metrics, dhidden, the one-hot rows of dW). No entropy gradient
 *     (entropy_out is output only). The caller must not use the workspace for anything else until the flush.
 *   grpo_deferred_dw_flush: dweight += stash^T . scaled hidden over rows [0, total_rows), where total_rows is the end of
 *     the last slot and consecutive slots start at the previous end rounded up to 512.
 * ------------------------------------------------------------------------------------------------------------------ */
long long grpo_chunk_capacity_rows(void);
int grpo_fused_loss_fwd_bwd_slot(const void* hidden, const void* weight, const int64_t* labels, const float* old_logp,
                                 const float* advantages, const float* ref_logp, const void* mask, int mask_dtype,
                                 int64_t rows, int64_t hidden_dim, int64_t vocab, float temperature,
                                 float clip_ratio_low, float clip_ratio_high, float clip_ratio_dual, int kl_mode,
                                 float kl_coef, float grad_accum, float* logp_out, float* entropy_out, void* dhidden,
                                 float* dweight, float* metrics, int64_t slot_row0, int64_t capacity_rows,
                                 void* workspace, size_t workspace_bytes, grpo_stream_t stream);
int grpo_deferred_dw_flush(int64_t total_rows, int64_t capacity_rows, int64_t hidden_dim, int64_t vocab, float* dweight,
                           void* workspace, size_t workspace_bytes, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-optimizer-step passes over the fp32 weight-gradient accumulator (the head's share of _optimizer_step,
 * dp_actor.py:155-167: clip_grad_norm_ + zero_grad). HBM-bound, one pass each.
 *   grpo_grad_sumsq     : out[0] (+)= sum(grad^2) in fp64, partial sums reduced in a fixed order; zero_after != 0 writes
 *                         zeros back in the same pass; accumulate != 0 adds to out[0] (global norm over several tensors).
 *                         scratch: GRPO_GRAD_SCRATCH_DOUBLES doubles of device memory.
 *   grpo_grad_scale_cast: out_bf16[i] = bf16(grad[i] * scale), scale = scale_dev[0] (device, e.g. the clip coefficient)
 *                         when non-null else scale_host; zero_after != 0 zeroes grad in the same pass.
 * ------------------------------------------------------------------------------------------------------------------ */
#define GRPO_GRAD_SCRATCH_DOUBLES 1024
int grpo_grad_sumsq(float* grad, int64_t n, int zero_after, int accumulate, double* out, double* scratch,
                    grpo_stream_t stream);
int grpo_grad_scale_cast(float* grad, int64_t n, const float* scale_dev, float scale_host, void* out_bf16,
                         int zero_after, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Gradient exchange of the fp32 dW accumulator over the GPUs of one NVSwitch box through PEER-MAPPED memory (replaces the
 * fp32 gradient averaging FSDP performs for the reference, actor/config.py:58 + fsdp_workers.py:242-280, for the replicated
 * lm_head weight, and the passes of _optimizer_step that follow it, dp_actor.py:155-167). One process per GPU; every
 * process maps every rank's buffers (CUDA IPC) and passes the `world` pointers in rank order. All calls are
 * stream-ordered; the caller brackets every data pass with grpo_peer_barrier on the same stream, on every rank:
 *     barrier -> reduce_scatter_sumsq -> barrier -> [norm, clip coefficient] -> scale_cast_allgather -> barrier
 *   grpo_peer_barrier  : flag_ptrs[q] = rank q's uint32[world] flag array (zero-initialised, peer-mapped). Announces
 *                        `epoch` (increasing by one per call, never 0) to every rank and waits for theirs. A peer that
 *                        does not arrive within timeout_ms (<= 0: 30 s) traps the kernel.
 *   grpo_peer_reduce_scatter_sumsq : buf_ptrs[q] = rank q's copy of the fp32 [n] buffer (n % 8 == 0, 16-byte aligned).
 *                        Slab `rank` (units of 8 elements, ceil(n / 8 / world) units per rank) of all copies is summed in
 *                        rank order, scaled by 1/world and stored into THIS rank's copy; the slab's sum of squares goes
 *                        to partial_ptrs[q][rank] (double[world] per rank, peer-mapped) for every q. scratch:
 *                        GRPO_GRAD_SCRATCH_DOUBLES doubles of local device memory, ZERO before the first call.
 *   grpo_peer_scale_cast_allgather : out_ptrs[q][slab] = bf16(grad[slab] * scale) for every q (out: bf16 [n] per rank,
 *                        peer-mapped; scale = scale_dev[0] when given, else scale_host); zero_after: the whole local
 *                        grad buffer is zeroed in the same pass.
 *   grpo_peer_allreduce_mean : general in-place mean all-reduce of a peer-mapped fp32 [n] buffer (n % 4 == 0): this rank
 *                        reduces its slab and writes the result into all copies. Barrier before and after.
 *   Results are bit-identical on all ranks and from run to run (every element is summed by one rank, in rank order).
 *   grpo_ipc_export / grpo_ipc_open / grpo_ipc_close: host-side CUDA IPC plumbing for those mappings. export describes a
 *     device pointer as (64-byte handle of the allocation it lies in, byte offset); another PROCESS opens it on its own
 *     current device (peer access is enabled on the way; a handle is mapped once per process and reference-counted) and
 *     gets the mapped pointer; close drops the reference.
 * ------------------------------------------------------------------------------------------------------------------ */
#define GRPO_MAX_PEERS 8
int grpo_ipc_export(const void* ptr, void* handle_out_64, int64_t* offset_out);
int grpo_ipc_open(const void* handle_64, int64_t offset, void** ptr_out);
int grpo_ipc_close(void* ptr, int64_t offset);
int grpo_peer_barrier(void* const* flag_ptrs, int rank, int world, unsigned int epoch, int timeout_ms,
                      grpo_stream_t stream);
int grpo_peer_reduce_scatter_sumsq(void* const* buf_ptrs, void* const* partial_ptrs, int rank, int world, int64_t n,
                                   double* scratch, grpo_stream_t stream);
int grpo_peer_scale_cast_allgather(float* grad, void* const* out_ptrs, int rank, int world, int64_t n,
                                   const float* scale_dev, float scale_host, int zero_after, grpo_stream_t stream);
int grpo_peer_allreduce_mean(void* const* buf_ptrs, int rank, int world, int64_t n, grpo_stream_t stream);
/* host-only: element range [e0, e1) of the slab `rank` owns in the two fused passes above (n % 8 == 0) */
int grpo_debug_peer_slab(int64_t n, int rank, int world, int64_t* e0, int64_t* e1);

/* ------------------------------------------------------------------------------------------------------------------
 * Token-level policy loss on given log-probs (no lm_head): the four masked means of compute_policy_loss
 * (core_algos.py:291-353), optionally the KL term (compute_kl :394-436) and dL/dlogp with
 * L = (pg + kl_coef * kl) / grad_accum.   acc_scratch: 16 doubles of device scratch.
 * ------------------------------------------------------------------------------------------------------------------ */
int grpo_policy_loss_fwd_bwd(const float* logp, const float* old_logp, const float* advantages, const float* ref_logp,
                             const void* mask, int mask_dtype, int64_t n, float clip_ratio_low, float clip_ratio_high,
                             float clip_ratio_dual, int kl_mode, float kl_coef, float grad_accum, float* dlogp,
                             float* metrics, double* acc_scratch, grpo_stream_t stream);

/* compute_kl (core_algos.py:394-436): out[i] = kl(logp[i], ref[i]); dout_dlogp nullable (d out / d logp). */
int grpo_compute_kl(const float* logp, const float* ref_logp, int64_t n, int kl_mode, float* out, float* dout_dlogp,
                    grpo_stream_t stream);

/* masked_mean over all elements (torch_functional.py:69-71): out[0] = sum(x*mask) / (sum(mask) + eps).
 * acc_scratch: 2 doubles of device scratch. */
int grpo_masked_mean(const float* x, const void* mask, int mask_dtype, int64_t n, float eps, float* out,
                     double* acc_scratch, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * compute_grpo_outcome_advantage (core_algos.py:137-175; caller ray_trainer.py:148-175).
 *   rewards f32 [bsz][t_len]; mask [bsz][t_len]; order int32 [bsz] = row ids sorted by group;
 *   offsets int32 [n_groups+1] = group boundaries in `order` (the host maps uid strings to this CSR view);
 *   advantages f32 [bsz][t_len] (out); seq_scratch f32 [2*bsz] device scratch (scores, normalised scores).
 * ------------------------------------------------------------------------------------------------------------------ */
int grpo_advantage(const float* rewards, const void* mask, int mask_dtype, const int32_t* order,
                   const int32_t* offsets, int64_t bsz, int64_t t_len, int64_t n_groups, float eps, float* advantages,
                   float* seq_scratch, grpo_stream_t stream);

/* The same in two steps, for a batch sharded by sequence over ranks (groups straddle ranks after the trainer's
 * sequence balancing, ray_trainer.py:628): each rank scores its own rows, the scores are all-gathered (bsz_all floats),
 * every rank runs the group statistics and broadcasts its rows [row_begin, row_begin + bsz_local) over its mask.
 *   seq_scratch f32 [bsz_all] device scratch (normalised score of every sequence). */
int grpo_sequence_scores(const float* rewards, int64_t bsz, int64_t t_len, float* scores, grpo_stream_t stream);
int grpo_advantage_from_scores(const float* scores_all, const int32_t* order, const int32_t* offsets, int64_t bsz_all,
                               int64_t n_groups, float eps, int64_t row_begin, const void* mask, int mask_dtype,
                               int64_t bsz_local, int64_t t_len, float* advantages, float* seq_scratch,
                               grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * The other advantage estimators of the reference's trainer, on the same skeleton.
 * Masks, `order` / `offsets` and `seq_scratch` (f32 [2*bsz]) as for grpo_advantage; acc_scratch: 4 doubles.
 *   grpo_rloo_advantage          core_algos.py:179-214  a_i = s_i - (sum_group - s_i) / (n_group - 1), broadcast over mask
 *   grpo_remax_advantage         core_algos.py:248-273  a_i = s_i - reward_baselines[i], broadcast over mask
 *   grpo_reinforce_pp_advantage  core_algos.py:217-245  R_t = r_t + gamma * R_{t+1} * mask_{t+1} (returns, out);
 *                                                      advantages = masked_whiten(R, mask)
 *   grpo_gae_advantage           core_algos.py:93-133   delta_t = r_t + gamma * v_{t+1} - v_t,
 *                                                      A_t = delta_t + gamma_lam * A_{t+1}; returns = A + v;
 *                                                      advantages = masked_whiten(A, mask). gamma_lam = gamma * lambda.
 * The recurrences run in the reference's fp32 operation order (no FMA contraction): bit-identical to the reference.
 * ------------------------------------------------------------------------------------------------------------------ */
int grpo_rloo_advantage(const float* rewards, const void* mask, int mask_dtype, const int32_t* order,
                        const int32_t* offsets, int64_t bsz, int64_t t_len, int64_t n_groups, float* advantages,
                        float* seq_scratch, grpo_stream_t stream);
int grpo_remax_advantage(const float* rewards, const float* reward_baselines, const void* mask, int mask_dtype,
                         int64_t bsz, int64_t t_len, float* advantages, float* seq_scratch, grpo_stream_t stream);
int grpo_reinforce_pp_advantage(const float* rewards, const void* mask, int mask_dtype, int64_t bsz, int64_t t_len,
                                float gamma, float* advantages, float* returns, double* acc_scratch,
                                grpo_stream_t stream);
int grpo_gae_advantage(const float* rewards, const float* values, const void* mask, int mask_dtype, int64_t bsz,
                       int64_t t_len, float gamma, float gamma_lam, float* advantages, float* returns,
                       double* acc_scratch, grpo_stream_t stream);

/* masked_var / masked_whiten over all elements (torch_functional.py:74-97). acc_scratch: 4 doubles.
 *   grpo_masked_whiten: out[i] = (values[i] - mean) * rsqrt(var + eps), var unbiased over the mask (out may alias values)
 *   grpo_masked_var   : out[0] = variance (Bessel-corrected iff unbiased and sum(mask) > 1), out[1] = masked mean */
int grpo_masked_whiten(const float* values, const void* mask, int mask_dtype, int64_t n, float eps, float* out,
                       double* acc_scratch, grpo_stream_t stream);
int grpo_masked_var(const float* values, const void* mask, int mask_dtype, int64_t n, int unbiased, float* out,
                    double* acc_scratch, grpo_stream_t stream);

/* compute_value_loss (core_algos.py:356-391): out[0] = vf_loss = 0.5 * masked_mean(max((vp - ret)^2, (clip(vp) - ret)^2)),
 * out[1] = vf_clipfrac; dvpreds f32 [n] (nullable) = d vf_loss / d vpreds. acc_scratch: 4 doubles. */
int grpo_value_loss_fwd_bwd(const float* vpreds, const float* returns, const float* values, const void* mask,
                            int mask_dtype, int64_t n, float cliprange_value, float* dvpreds, float* out,
                            double* acc_scratch, grpo_stream_t stream);

/* apply_kl_penalty (ray_trainer.py:125-145; compute_rewards core_algos.py:276-283 is its kl_mode = GRPO_KL_KL case
 * without a mask): token_level_rewards = token_level_scores - kl_coef * compute_kl(logp, ref_logp) * mask;
 * current_kl[0] = mean over sequences of masked_mean(kld, mask, dim=-1). ref_logp == NULL: kld = 0.
 * All tensors [bsz][t_len]; acc_scratch: 4 doubles. */
int grpo_kl_penalty_rewards(const float* token_level_scores, const float* logp, const float* ref_logp,
                            const void* mask, int mask_dtype, int64_t bsz, int64_t t_len, int kl_mode, float kl_coef,
                            float* token_level_rewards, float* current_kl, double* acc_scratch, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Ragged micro-batches: run the head on the unmasked response slots only.
 * Replaces: the reference computes log-probs for every slot of the padded [B, T] block and multiplies the padded ones
 * by zero (dp_actor.py:136-139, :253-273); its padding-free branch gathers / scatters tokens for the transformer body
 * only (flash-attn bert_padding unpad_input / pad_input, dp_actor.py:86-104).
 *   grpo_compact_index: gather_idx int32 [n] (gather_idx[j] = slot of the j-th unmasked entry, original order; entries
 *     past the count are unspecified), inverse int32 [n] (inverse[slot] = j or -1), count int32 [1] (device),
 *     scratch of grpo_compact_scratch_bytes(n). mask_dtype 0..2. No host synchronisation.
 *   grpo_gather_rows : out[j][:] = in[gather_idx[j]][:] for j < m           (rows of row_bytes bytes, multiple of 4)
 *   grpo_scatter_rows: out[i][:] = inverse[i] >= 0 ? in[inverse[i]][:] : 0   for i < n   (every output row written)
 * ------------------------------------------------------------------------------------------------------------------ */
size_t grpo_compact_scratch_bytes(int64_t n);
int grpo_compact_index(const void* mask, int mask_dtype, int64_t n, int32_t* gather_idx, int32_t* inverse,
                       int32_t* count, void* scratch, size_t scratch_bytes, grpo_stream_t stream);
int grpo_gather_rows(const void* in, const int32_t* gather_idx, int64_t m, int64_t row_bytes, void* out,
                     grpo_stream_t stream);
int grpo_scatter_rows(const void* in, const int32_t* inverse, int64_t n, int64_t row_bytes, void* out,
                      grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * log_probs_from_logits on MATERIALISED logits (torch_functional.py:45-66), API parity only.
 *   logits_dtype: 0 = f32, 1 = bf16, 2 = f16;  logits [rows][ld];  logp / entropy / lse f32 [rows] (each nullable)
 * Backward: dlogits = dlogp * (onehot - softmax) - dentropy * p * (log p + H)   (dlogits may alias logits).
 * ------------------------------------------------------------------------------------------------------------------ */
int grpo_logprob_from_logits(const void* logits, int logits_dtype, const int64_t* labels, int64_t rows, int64_t vocab,
                             int64_t ld, float* logp, float* entropy, float* lse, grpo_stream_t stream);
int grpo_logprob_from_logits_bwd(const void* logits, int logits_dtype, const int64_t* labels, const float* lse,
                                 const float* dlogp, const float* dentropy, const float* entropy, int64_t rows,
                                 int64_t vocab, int64_t ld, void* dlogits, int64_t ld_out, grpo_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Debug / test entry: plain tcgen05 GEMM  C[m][n] (f32) = A . B^T-style product with selectable operand majorness.
 *   a_mn_major == 0: A is [m][k] row-major, else A is [k][m] row-major; same for B with n.
 *   cta_group: 1 or 2.   accumulate != 0: C += (red.global.add).
 * ------------------------------------------------------------------------------------------------------------------ */
int grpo_debug_gemm(const void* a, const void* b, float* c, int64_t m, int64_t n, int64_t k, int a_mn_major,
                    int b_mn_major, int cta_group, int accumulate, grpo_stream_t stream);

/* Debug / test entry (host only, no device work): the work-unit plan a persistent GEMM launch would use for `tiles`
 * output tiles of `k_blocks` K-blocks on `groups_avail` CTA groups. split_mode: 0 = whole tiles, 1 = one short split-K
 * round for a small remainder (dW GEMM), 2 = best slice count over several short rounds (dHidden GEMM, fp32 split path).
 *   units (nullable) int32 [max_units][3] = {tile, first K-block, end K-block} per unit, in walk order (unit u runs on
 *   group u % groups);  info (nullable) int32 [4] = {groups launched, units, progress-window bound, slices per tile}. */
int grpo_debug_plan_units(int64_t tiles, int64_t k_blocks, int groups_avail, int split_mode, int32_t* units,
                          int64_t max_units, int32_t* info);

#ifdef __cplusplus
}
#endif
#endif /* GRPO_B200_H */
