/* omnisafe_b200 -- C ABI of the B200-native on-policy SafeRL hot path.
 *
 * The reference (PKU-Alignment/omnisafe) is pure Python: its seam for this path is the set of
 * Python methods listed below, not an FFI.  Each entry point here is what a ctypes binding of the
 * corresponding reference method body would call (see INTEGRATION.md for the stub).  All pointers
 * are DEVICE pointers to contiguous row-major arrays the caller owns; `stream` is a cudaStream_t
 * (NULL = default stream).  Every function returns 0 on success, non-zero on failure;
 * osb_last_error() returns the reason.  No function synchronises the host unless stated.
 *
 * Slab layout ("time-major"): per-step scalars are [T][N] (env index contiguous), observations
 * [T][N][O], actions [T][N][A].  Sample k of the reference's env-major order
 * (vector_onpolicy_buffer.py:L125-129, k = i*T + t) lives at slab row t*N + i.
 *
 * Flat parameter vector theta = [actor | reward_critic | cost_critic], each in the reference's
 * named_parameters() order (utils/tools.py:L35-129): actor = log_std[A], W1[64][O], b1[64],
 * W2[64][64], b2[64], W3[A][64], b3[A]; critic = W1, b1, W2, b2, W3[1][64], b3[1].
 */
#ifndef OMNISAFE_B200_H
#define OMNISAFE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- plumbing ---------------------------------------------------------------------------- */
const char* osb_last_error(void);
int osb_abi_version(void);
int osb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- dual GAE (segmented reverse scan) ----------------------------------------------------
 * replaces OnPolicyBuffer.finish_path            omnisafe/common/buffer/onpolicy_buffer.py:L148-203
 *          _calculate_adv_and_value_targets('gae')                                     :L299-303
 *          discount_cumsum (fp64 carry)           omnisafe/utils/math.py:L59-82
 * flags bit0 = terminated, bit1 = truncated; a path also ends at t == T-1.  At a path end the
 * bootstrap value is 0 if terminated, else boot_{r,c}[t][i].  disc_ret may be NULL.
 * workspace: osb_gae_workspace_doubles(N) doubles.  sums[4] <- {sum adv_r, sum adv_r^2,
 * sum adv_c, count} over this rank's samples (fp64; all-reduce them across ranks, then call
 * osb_adv_moments). */
int osb_gae_workspace_doubles(int n_envs);
int osb_gae_dual(const float* rew, const float* cost, const float* val_r, const float* val_c,
                 const unsigned char* flags, const float* boot_r, const float* boot_c, int T, int N,
                 double gamma, double lam, double lam_c, double penalty_coef, float* adv_r,
                 float* adv_c, float* tv_r, float* tv_c, float* disc_ret, double* workspace,
                 double* sums, void* stream);
/* moments[4] <- {mean_r, std_r + 1e-8, mean_c, 1}: the statistics VectorOnPolicyBuffer.get()
 * standardises with (vector_onpolicy_buffer.py:L131-136, utils/distributed.py:L382-388). */
int osb_adv_moments(const double* sums, int standardize_r, int standardize_c, float* moments,
                    void* stream);
/* out_r = (adv_r - mean_r) / (std_r + 1e-8), out_c = adv_c - mean_c  (what get() returns). */
int osb_adv_standardize(const float* adv_r, const float* adv_c, const float* moments, long long n,
                        float* out_r, float* out_c, void* stream);
/* discount_cumsum over `rows` independent vectors of length `len` (utils/math.py:L59-82);
 * x is fp32 (x_is_f64 = 0) or fp64, out is fp64. */
int osb_discount_cumsum(const void* x, int x_is_f64, int rows, int len, double discount,
                        double* out, void* stream);

/* ---- rollout: fused env step + 3 MLP forwards + sample + slab append ------------------------
 * replaces, per step, ConstraintActorCritic.step  models/actor_critic/constraint_actor_critic.py:L84-109
 *          ActionScale.step / ObsNormalize.step    envs/wrapper.py:L510-514, L231-241
 *          Normalizer.normalize / _push            common/normalizer.py:L88-139
 *          VectorOnPolicyBuffer.store              common/buffer/vector_onpolicy_buffer.py:L96-99
 *          the per-env done loop                   adapter/onpolicy_adapter.py:L114-136
 * Env state arrays: s_raw[2][N][O], final_raw[2][N][O], ep_step[N], episode[N], gstep[N],
 * ep_ret[N], ep_cost[N], ep_len[N], bias[O].  Normaliser state: mean/sumsq/std/mean1/std1 [O],
 * count[2], acc_all[2][O], acc_fin[2][O] (int64 fixed point), fin_count[1], had_fin[1], ticket[1]
 * (all zero-initialised by the caller).  osb_env_reset = OnPolicyAdapter.reset() at epoch start
 * (onpolicy_adapter.py:L80).  osb_rollout_step with t in [0, T) performs step t; t == T is the
 * epoch-end bootstrap launch (critics only).  eps = [N][A] standard-normal draws of this step
 * (parity mode) or NULL (in-kernel Philox keyed by noise_seed / global_step). */
int osb_env_reset(int O, int A, int max_episode_steps, unsigned seed, unsigned term_threshold,
                  unsigned env_id_offset, float cost_threshold, int obs_normalize, int N,
                  float* s_raw, float* final_raw, int* ep_step, unsigned* episode, unsigned* gstep,
                  float* ep_ret, float* ep_cost, int* ep_len, const float* bias,
                  float* norm_mean, float* norm_sumsq, float* norm_std, float* norm_mean1,
                  float* norm_std1, long long* norm_count, long long* acc_all, long long* acc_fin,
                  int* fin_count, int* had_fin, unsigned* ticket, void* stream);
int osb_rollout_step(int O, int A, int max_episode_steps, unsigned seed, unsigned term_threshold,
                     unsigned env_id_offset, float cost_threshold, int obs_normalize, int N, int T,
                     int t, float* s_raw, float* final_raw, int* ep_step, unsigned* episode,
                     unsigned* gstep, float* ep_ret, float* ep_cost, int* ep_len, const float* bias,
                     float* norm_mean, float* norm_sumsq, float* norm_std, float* norm_mean1,
                     float* norm_std1, long long* norm_count, long long* acc_all,
                     long long* acc_fin, int* fin_count, int* had_fin, unsigned* ticket,
                     float* obs, float* act, float* logp, float* rew, float* cost, float* val_r,
                     float* val_c, float* boot_r, float* boot_c, unsigned char* flags, float* epfin,
                     const float* theta, const float* eps, unsigned noise_seed,
                     unsigned global_step, void* stream);
/* Whole-epoch rollout in one call: reset, T step launches (eps_all = [T][N][A] or NULL), the
 * epoch-end bootstrap launch and the episode window (= OnPolicyAdapter.rollout,
 * adapter/onpolicy_adapter.py:L58-136).  Philox counter = epoch_index * T + t. */
int osb_rollout_epoch(int O, int A, int max_episode_steps, unsigned seed, unsigned term_threshold,
                      unsigned env_id_offset, float cost_threshold, int obs_normalize, int N, int T,
                      float* s_raw, float* final_raw, int* ep_step, unsigned* episode,
                      unsigned* gstep, float* ep_ret, float* ep_cost, int* ep_len, const float* bias,
                      float* norm_mean, float* norm_sumsq, float* norm_std, float* norm_mean1,
                      float* norm_std1, long long* norm_count, long long* acc_all,
                      long long* acc_fin, int* fin_count, int* had_fin, unsigned* ticket,
                      float* obs, float* act, float* logp, float* rew, float* cost, float* val_r,
                      float* val_c, float* boot_r, float* boot_c, unsigned char* flags, float* epfin,
                      const float* theta, const float* eps_all, unsigned noise_seed,
                      unsigned epoch_index, int W, float* ring, int* meta, double* window_sums,
                      void* stream);
/* Logger window of the last <= W finished episodes in (step, env) order
 * (common/logger.py:L253-282, adapter/onpolicy_adapter.py:L159-175).  ring[3][W], meta[2] persist
 * across epochs; window_sums[4] <- {sum EpRet, sum EpCost, sum EpLen, count} (fp64). */
int osb_episode_window(const unsigned char* flags, const float* epfin, int T, int N, int W,
                       float* ring, int* meta, double* window_sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMNISAFE_B200_H */
