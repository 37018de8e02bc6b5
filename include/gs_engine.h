/*
 * gs_engine.h — C-ABI of the B200 rollout-and-update engine for gymnasium-solver's
 * collect -> GAE / MC returns -> PPO / REINFORCE update hot path.
 *
 * The reference (tsilva/gymnasium-solver) is 100 % Python and has NO FFI for this path; its extension
 * points are Python protocols.  Each entry point below therefore cites the reference *Python* interface
 * whose arithmetic it replaces (path:line relative to the reference root).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; gs_last_error() gives a thread-local message.
 *     Nothing throws, aborts or synchronises the device unless stated.
 *   - every pointer argument is a CALLER-OWNED DEVICE pointer (torch.Tensor.data_ptr()) unless the
 *     name ends in _host.  The library allocates only inside opaque handles (gs_env_t).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - rollout arrays are TIME-MAJOR: element (t, n) of a (T, N) array lives at t*N + n; observations
 *     are (T, N, D) row-major, exactly the layout of RolloutBuffer.obs_buf (utils/rollout_buffer.py:48-56).
 *   - "env-major sample id" i = n*T + t is the index space of the reference's flattened
 *     RolloutTrajectory (utils/rollout_buffer.py:11-13) and of its MultiPassRandomSampler.
 */
#ifndef GS_ENGINE_H
#define GS_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_VERSION 100

/* ---- environment kinds (gym.make ids at utils/environment.py:94-96) ------------------------------ */
#define GS_ENV_CARTPOLE_V1    0
#define GS_ENV_ACROBOT_V1     1
#define GS_ENV_MOUNTAINCAR_V0 2

/* ---- per-env wrappers (gym_wrappers/__init__.py:29-46 registry ids) ------------------------------ */
#define GS_WRAP_MOUNTAINCAR_STATE_COUNT_BONUS 1 /* gym_wrappers/MountainCarV0/state_count_bonus.py:96-126 */
#define GS_WRAP_CARTPOLE_REWARD_SHAPER        2 /* gym_wrappers/CartPoleV1/reward_shaper.py:43-77        */
#define GS_WRAP_MOUNTAINCAR_REWARD_SHAPER     3 /* gym_wrappers/MountainCarV0/reward_shaper.py:60-102    */
/* Scripted replay over MountainCar-v0's spaces (2 observations, 3 actions): the env ignores its actions and replays per-step tables --
 * what the scripted fake envs of the reference's collector tests do (tests/test_rollouts_extra.py:161-240, test_mc_baseline_mask.py,
 * test_rollouts.py:95-123), so their known answers run on the DEVICE collector.  params = {L, reward[L], terminated[L], truncated[L],
 * obs[(L+1)*2]}: step k (0-based, k clamps at L-1) pays reward[k] with those flags and shows obs[k+1]; reset shows obs[0].  No
 * autoreset and no time limit (the tables say when an episode ends); episode statistics restart after a done like everywhere else. */
#define GS_WRAP_SCRIPTED_REPLAY               4

/* ---- activations (utils/torch.py:15-25 ACTIVATION_MAPPING; engine supports these two) ------------ */
#define GS_ACT_RELU 0
#define GS_ACT_TANH 1

/* ---- metric slots written by gs_ppo_step / gs_reinforce_step (keys: agents/ppo/ppo_agent.py:132-143,
 *      agents/base_agent.py:120-127, utils/torch.py:140-143,170-173) -------------------------------- */
enum {
    GS_M_LOSS_TOTAL = 0,      /* opt/loss/total            */
    GS_M_LOSS_POLICY,         /* opt/loss/policy           */
    GS_M_LOSS_ENTROPY,        /* opt/loss/entropy  (= -H)  */
    GS_M_ENTROPY,             /* opt/policy/entropy        */
    GS_M_LOSS_ENTROPY_SCALED, /* opt/loss/entropy_scaled   */
    GS_M_LOSS_VALUE,          /* opt/loss/value            */
    GS_M_LOSS_VALUE_SCALED,   /* opt/loss/value_scaled     */
    GS_M_CLIP_FRACTION,       /* opt/ppo/clip_fraction     */
    GS_M_CLIP_FRACTION_VF,    /* opt/ppo/clip_fraction_vf  */
    GS_M_EXPLAINED_VAR,       /* opt/value/explained_var   */
    GS_M_KL,                  /* opt/ppo/kl                */
    GS_M_APPROX_KL,           /* opt/ppo/approx_kl         */
    GS_M_ADV_NORM_MEAN,       /* roll/adv/norm/mean        */
    GS_M_ADV_NORM_STD,        /* roll/adv/norm/std         */
    GS_M_TARGETS_MEAN,        /* policy_targets_mean (REINFORCE) */
    GS_M_TARGETS_STD,         /* policy_targets_std  (REINFORCE) */
    GS_M_ACT0_MEAN,           /* opt/activations/backbone.0/mean     */
    GS_M_ACT0_STD,            /* opt/activations/backbone.0/std      */
    GS_M_ACT0_DEAD_PCT,       /* opt/activations/backbone.0/dead_pct */
    GS_M_ACT0_DEAD_MAX,       /* opt/activations/backbone.0/dead_max */
    GS_M_ACT1_MEAN,           /* opt/activations/backbone.2/...      */
    GS_M_ACT1_STD,
    GS_M_ACT1_DEAD_PCT,
    GS_M_ACT1_DEAD_MAX,
    GS_M_GRAD_NORM_ALL,       /* opt/grads/norm/all         (written by gs_clip_grad_norm) */
    GS_M_GRAD_NORM_BACKBONE,  /* opt/grads/norm/backbone    */
    GS_M_GRAD_NORM_POLICY,    /* opt/grads/norm/policy_head */
    GS_M_GRAD_NORM_VALUE,     /* opt/grads/norm/value_head  */
    GS_M_CLIP_COEF,           /* scale applied by the global-norm clip */
    GS_M_RET_NORM_MEAN,       /* roll/return/norm/mean (REINFORCE, normalize_returns == batch) */
    GS_M_RET_NORM_STD,        /* roll/return/norm/std  */
    GS_M_BATCH_COUNT,         /* samples that contributed */
    GS_M_SCRATCH = 32,        /* [32..39] device scratch of gs_clip_grad_norm; not a metric */
    GS_N_METRICS = 40
};

typedef struct gs_env gs_env_t; /* opaque: SoA fp64 state, elapsed steps, episode accumulators, RNG key, wrapper tables */

/* MLPActorCritic / MLPPolicy weights (utils/models.py:233-346): torch nn.Linear layout weight[out][in].
 * The flat parameter / gradient vector is the concatenation in nn.Module.parameters() order:
 * w1,b1,w2,b2,wp,bp,wv,bv  (backbone.0, backbone.2, policy_head, value_head).                        */
typedef struct gs_mlp {
    int32_t obs_dim;    /* D                                              */
    int32_t hidden1;    /* H1                                             */
    int32_t hidden2;    /* H2, 0 => single hidden layer (mlp_tiny)         */
    int32_t n_actions;  /* A                                              */
    int32_t has_value;  /* 1 = MLPActorCritic, 0 = MLPPolicy              */
    int32_t activation; /* GS_ACT_*                                       */
    const float* w1; const float* b1;
    const float* w2; const float* b2; /* NULL when hidden2 == 0 */
    const float* wp; const float* bp;
    const float* wv; const float* bv; /* NULL when has_value == 0 */
} gs_mlp_t;

/* Time-major device rollout storage (utils/rollout_buffer.py:48-56).  Nullable members are skipped. */
typedef struct gs_rollout {
    int32_t T;
    int32_t obs_dim;
    int64_t N;
    float*   obs;         /* (T,N,D)                                     */
    float*   next_obs;    /* (T,N,D) nullable                            */
    int32_t* actions;     /* (T,N)   (int64 only on the Python surface)  */
    float*   logprobs;    /* (T,N)                                       */
    float*   values;      /* (T,N)                                       */
    float*   rewards;     /* (T,N)                                       */
    uint8_t* dones;       /* (T,N)  terminated|truncated                 */
    uint8_t* timeouts;    /* (T,N)  truncated                            */
    float*   last_obs;    /* (N,D)  obs after the final step             */
    float*   last_values; /* (N,)   V(last_obs), nullable                */
    /* episode accounting (RecordEpisodeStatistics + rollout_collector.py:210-294) */
    double*  ep_return;   /* (T,N) nullable: episode return where done, else 0 */
    int32_t* ep_length;   /* (T,N) nullable: episode length where done, else 0 */
} gs_rollout_t;

/* One minibatch for the update kernels.  Arrays are the (T,N) time-major rollout arrays; sample ids are
 * env-major (n*T+t) as produced by MultiPassRandomSampler (utils/samplers.py:29-34) and consumed by
 * RolloutCollector.slice_trajectories (utils/rollout_collector.py:657-682).
 *   idx != NULL          : explicit sample ids (n entries)
 *   idx == NULL, perm_len: sample id = bijection_{perm_key}(perm_offset + i) over [0, perm_len)
 *   idx == NULL, else    : sample id = perm_offset + i
 * idx_map (nullable) is the MC valid-index remap of rollout_collector.py:664-669.                      */
typedef struct gs_batch {
    int64_t n;
    const int64_t* idx;
    uint64_t perm_key;
    int64_t  perm_offset;
    int64_t  perm_len;
    const int64_t* idx_map;
    int32_t T;
    int32_t obs_dim;
    int64_t N;
    const float*   obs;
    const int32_t* actions;
    const float*   logp_old;
    const float*   values_old; /* PPO only */
    const float*   adv;
    const float*   ret;
    /* Optional (nullable): the same six arrays as ONE 64-byte record per (t,n) in time-major order, written by
     * gs_rollout_pack once per rollout: {obs[0..7] (zero padded), action bits, logp_old, values_old, adv, ret, 0, 0, 0}.
     * With it the minibatch gather of the tensor-core update kernel is one aligned 64-byte access per sample instead of
     * seven scattered ones (one TLB lookup and two sectors instead of seven of each). */
    const float*   packed;
    /* 1 = gs_batch_prepare already ran for exactly this minibatch on this workspace (sample offsets are in place): the
     * step skips its gather pass.  Anything else about the batch must be unchanged between the two calls. */
    int32_t prepared;
    /* 1 = stop after the update kernel: the per-CTA partial gradients / metric partials stay in the workspace and
     * gs_update_finish completes the step (ordered reduction, gradient all-reduce, metrics, clip, Adam) in one launch.
     * 2 = the same, and the caller guarantees a CLEAN workspace: zero-initialised before its first use and only ever used by complete
     * steps (every finishing kernel leaves the counters it consumed zeroed).  The step then enqueues no memset, so the update kernel is
     * launched programmatically right under the previous gs_update_finish. */
    int32_t defer_reduce;
    /* Optional (nullable) caller-owned buffer of n uint32 time-major sample offsets.  gs_batch_prepare writes it, a step
     * with prepared = 1 reads it; NULL = the workspace's own buffer (one minibatch at a time).  Lets a caller prepare every
     * minibatch of a rollout up front and exchange all their moments between ranks in ONE collective. */
    uint32_t* offsets;
} gs_batch_t;
#define GS_RECORD_FLOATS 16

typedef struct gs_ppo_hparams {
    float clip_range;      /* agents/ppo/ppo_agent.py:59          */
    float clip_range_vf;   /* :83-87                              */
    float vf_coef;         /* :105                                */
    float ent_coef;        /* :106                                */
    int32_t normalize_adv; /* 0 off, 1 "batch" (utils/torch.py:97-99, unbiased std + 1e-8) */
    int32_t track_activations; /* utils/models.py:121-194 stats into GS_M_ACT*           */
} gs_ppo_hparams_t;

typedef struct gs_reinforce_hparams {
    float ent_coef;              /* agents/reinforce/reinforce_agent.py:65-66 */
    int32_t policy_targets;      /* 0 returns, 1 advantages (:31-36)          */
    int32_t normalize_returns;   /* 0 off, 1 batch (:23-25)                   */
    int32_t normalize_adv;       /* 0 off, 1 batch (:28)                      */
    int32_t track_activations;
} gs_reinforce_hparams_t;

/* ---- library ------------------------------------------------------------------------------------ */
int         gs_version(void);
const char* gs_last_error(void);
/* number of SMs of `device` (grid sizing on the host side) */
int         gs_device_sm_count(int device);

/* ---- vector environments: gym.make + TimeLimit + SyncVectorEnv(NEXT_STEP autoreset) +
 *      RecordEpisodeStatistics, utils/environment.py:94-96,136,212,410-415 (gymnasium 1.1.1, external) */
int gs_env_create(int env_kind, int64_t n_envs, int64_t env_id_offset /* global id of local env 0 */,
                  uint64_t seed, int max_episode_steps /* 0 -> 500/500/200 */, int device, gs_env_t** out);
int gs_env_destroy(gs_env_t* env);
/* VecNormalizeStatic (gym_wrappers/vec_normalize_static.py:20-60, applied by utils/environment.py:215-216 for normalize_obs="static"),
 * fused into every kernel that emits observations: dims with finite low < high become (x - low) / ((high - low) + 1e-8) in fp32, finite
 * degenerate dims 0, dims with non-finite bounds pass through.  low / high: HOST arrays of the env's Box bounds (NULL, NULL = off). */
int gs_env_set_obs_normalization(gs_env_t* env, const float* low_host, const float* high_host, int n_dims);
int gs_env_obs_dim(int env_kind);
int gs_env_state_dim(int env_kind);
int gs_env_n_actions(int env_kind);
int64_t gs_env_num_envs(const gs_env_t* env);
/* parity hooks: inject / read the fp64 SoA state [S][N] (+ TimeLimit elapsed steps, nullable) */
int gs_env_set_state(gs_env_t* env, const double* state, const int32_t* elapsed, void* stream);
int gs_env_get_state(gs_env_t* env, double* state, int32_t* elapsed, void* stream);
/* Exact snapshot / restore of everything the handle holds on the device (physics state, TimeLimit and
 * RecordEpisodeStatistics accumulators, autoreset flags, reset-stream counters, StateCountBonus tables): checkpoint / resume of
 * the env side, which the reference cannot do (agents/base_agent.py:658-732 saves no env state; TODO.md:29).  blob is a
 * caller-owned device buffer of gs_env_snapshot_bytes(); a blob only loads into a handle of the same kind, size and wrapper. */
int64_t gs_env_snapshot_bytes(const gs_env_t* env);
int gs_env_save(gs_env_t* env, void* blob, void* stream);
int gs_env_load(gs_env_t* env, const void* blob, void* stream);
/* VectorEnv.reset(): new episode in every env (rollout_collector.py:317); obs is (N,D) */
int gs_env_reset(gs_env_t* env, float* obs, void* stream);
/* VectorEnv.step(actions) (rollout_collector.py:504). ep_return/ep_length (nullable) carry
 * infos["episode"]["r"|"l"] where terminated|truncated, 0 elsewhere. */
int gs_env_step(gs_env_t* env, const int32_t* actions, float* obs, float* reward, uint8_t* terminated,
                uint8_t* truncated, double* ep_return, int32_t* ep_length, void* stream);
/* EnvWrapperRegistry.apply (gym_wrappers/env_wrapper_registry.py:10-16) for the device wrappers.
 * params_host: StateCountBonus {position_bins, velocity_bins, bonus_scale, bonus_type(0 count,1 inverse,2 log), min_count}
 *              CartPole shaper {angle_reward_scale, position_reward_scale, clip_potential}
 *              MountainCar shaper {position_reward_scale, velocity_reward_scale, height_reward_scale} */
int gs_wrapper_attach(gs_env_t* env, int wrapper_kind, const double* params_host, int n_params);

/* ---- policy_act (utils/policy_ops.py:14-34 -> utils/models.py:328-346 -> Categorical) -------------
 * uniforms (nullable, (n,)): injected U[0,1) draws for parity tests; otherwise Philox(rng_seed, row, rng_offset).
 * Sampling is inverse-CDF over softmax(logits): action = min{k : cdf_k > u}.  deterministic -> argmax (dist.mode). */
int gs_policy_act(const gs_mlp_t* mlp, const float* obs, int64_t n, uint64_t rng_seed, uint64_t rng_offset,
                  int64_t row_offset, int deterministic, const float* uniforms, int32_t* actions, float* logp,
                  float* value, float* logits_out /* nullable (n,A) */, void* stream);
/* policy_predict_values (utils/policy_ops.py:36-41) */
int gs_policy_values(const gs_mlp_t* mlp, const float* obs, int64_t n, float* value, void* stream);

/* ---- fused RolloutCollector._collect hot loop (utils/rollout_collector.py:474-542): for n_steps vector
 * steps: obs -> policy_act -> env.step -> buffer.add, then V(last_obs).  One persistent launch. */
int gs_rollout_collect(gs_env_t* env, const gs_mlp_t* mlp, const gs_rollout_t* buf, float* cur_obs /* (N,D) in/out */,
                       uint64_t rng_seed, uint64_t rng_offset /* vector-step counter at rollout start */,
                       int deterministic, void* stream);

/* ---- returns / advantages (utils/returns_advantages.py) ------------------------------------------- */
/* compute_batched_gae_advantages_and_returns :115-155 (fp32 arithmetic, reverse scan over t) */
int gs_gae(const float* values, const float* rewards, const uint8_t* dones, const uint8_t* timeouts,
           const float* last_values, const float* bootstrapped /* nullable (T,N) */, int T, int64_t N,
           double gamma, double gae_lambda /* Python floats of the reference, rounded to fp32 exactly as numpy does */,
           float* adv, float* ret, void* stream);
/* The same with bootstrapped == an all-zero (T,N) array, WITHOUT reading one (18 instead of 22 bytes per element): the value the
 * reference's collector holds whenever the vector env emits no final observation (NEXT_STEP autoreset: rollout_collector.py:262-284
 * never fires), i.e. next value of a truncated step = 0.  Bit-identical to gs_gae given a zero array. */
int gs_gae_zero_boot(const float* values, const float* rewards, const uint8_t* dones, const uint8_t* timeouts,
                     const float* last_values, int T, int64_t N, double gamma, double gae_lambda, float* adv, float* ret, void* stream);
/* compute_batched_mc_returns :67-91 (+ convert_returns_to_full_episode :93-113 when episode_mode).
 * timeouts nullable == all False (mc_treat_timeouts_as_terminals, rollout_collector.py:392-393).
 * last_terminal (nullable, (N,) int32): index of the last real terminal per env, -1 if none —
 * the per-env fact behind _build_valid_mask_and_index_map :33-52. */
int gs_mc_returns(const float* rewards, const uint8_t* dones, const uint8_t* timeouts, int T, int64_t N,
                  double gamma, int episode_mode, float* ret, int32_t* last_terminal, void* stream);
/* convert_returns_to_full_episode :93-113 as a standalone in-place pass over reward-to-go returns (T,N) */
int gs_returns_to_full_episode(float* ret, const uint8_t* dones, const uint8_t* timeouts /* nullable */, int T, int64_t N,
                               void* stream);
/* _build_valid_mask_and_index_map :33-52 + _build_idx_map_from_valid_mask :19-30, env-major (N*T,).
 * n_valid (device int64[1]) receives the number of valid entries (0 => reference returns None). */
int gs_valid_index_map(const int32_t* last_terminal, int T, int64_t N, uint8_t* valid_mask, int64_t* idx_map,
                       int64_t* n_valid, void* workspace, int64_t workspace_bytes, void* stream);
int64_t gs_valid_index_map_workspace_bytes(int64_t N);
/* masked moments over a (T,N) array: out[0]=sum, out[1]=sumsq, out[2]=count (double).  last_terminal nullable:
 * when given only t <= last_terminal[n] contributes (RunningStats over valid returns, rollout_collector.py:415-418).
 * Accumulates INTO out (caller zeroes). */
int gs_moments(const float* x, const int32_t* last_terminal, int T, int64_t N, double* out, void* stream);
/* The same over the valid entries of a rollout, with the reference's fallback: when n_valid (device int64[1], written by
 * gs_valid_index_map) is 0 the mask is None there and the statistics cover EVERY element (rollout_collector.py:435-455). */
int gs_moments_valid(const float* x, const int32_t* last_terminal, const int64_t* n_valid, int T, int64_t N, double* out, void* stream);
/* y = (x - mean)/(std + eps) with mean/std derived on device from moments (population std, numpy semantics,
 * returns_advantages.py:55-64); in place allowed. */
int gs_normalize(const float* x, int64_t n, const double* moments, float eps, float* y, void* stream);
/* y = x - mean(moments)  (MC baseline, rollout_collector.py:423-425) */
int gs_shift_by_mean(const float* x, int64_t n, const double* moments, float* y, void* stream);

/* ---- minibatch statistics for "batch" advantage normalisation (utils/torch.py:97-99,148-174) ---------
 * out[0..2] += sum, sumsq, count of field[sample ids of the batch]. */
int gs_batch_moments(const gs_batch_t* batch, const float* field /* (T,N) */, double* out, void* stream);

/* ---- PPOAgent.losses_for_batch fwd + bwd (agents/ppo/ppo_agent.py:21-152) ---------------------------
 * partials: workspace of gs_update_workspace_bytes(); grads_flat (P,) receives dLoss/dtheta (deterministic
 * two-stage reduction); metrics (double[GS_N_METRICS]) receives the finalised scalars of the metric enum.
 * adv_moments: device double[3] {sum, sum of squares, count} the "batch" advantage normalisation uses (gs_batch_moments,
 * all-reduced over ranks by the caller when the minibatch is sharded); NULL with hp.normalize_adv == 1 = take them over
 * this minibatch inside the call (fused into its gather pass). */
/* Bytes of device scratch gs_ppo_step / gs_reinforce_step need for minibatches of up to max_batch samples (per-CTA partial
 * gradients and metrics + 4 bytes per sample for the translated sample offsets). */
int64_t gs_update_workspace_bytes(const gs_mlp_t* mlp, int device, int64_t max_batch);
/* Builds the packed sample records of a rollout (see gs_batch_t.packed) from batch->{obs,actions,logp_old,values_old,adv,ret}
 * (values_old nullable: 0): packed is (T*N, GS_RECORD_FLOATS) floats, 16-byte aligned.  Streaming pass, once per rollout. */
int gs_rollout_pack(const gs_batch_t* batch, float* packed, void* stream);
/* First half of a step for callers that must exchange the minibatch moments between ranks before the update (sharded
 * minibatches): runs the gather pass of gs_ppo_step / gs_reinforce_step (sample offsets into the workspace) and accumulates
 * {sum, sumsq, count} of batch->adv into moments[0..2] (want_adv) and of batch->ret into moments[3..5] (want_ret); moments is
 * zeroed by the call.  All-reduce moments, then call the step with them and batch->prepared = 1. */
int gs_batch_prepare(const gs_mlp_t* mlp, const gs_batch_t* batch, int want_adv, int want_ret, double* moments,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* Kernel selection for gs_ppo_step / gs_reinforce_step: 0 (default) = tcgen05 tensor-core kernels where they exist (64x64, 128x128 and
 * relu 256x256 MLPs with obs_dim <= 7: kind::f16, fp16x3 split, fp32 TMEM accumulators -- the presets of utils/model_registry.py:20-31),
 * 1 = fp32 FMA-pipe kernel everywhere.  Env GS_UPDATE_IMPL=simt|tc sets the default; GS_ROLLOUT_IMPL=simt does the same for the collect /
 * policy_act kernels. */
int gs_set_update_impl(int impl);
int64_t gs_mlp_param_count(const gs_mlp_t* mlp);
int gs_ppo_step(const gs_mlp_t* mlp, const gs_batch_t* batch, const gs_ppo_hparams_t* hp, const double* adv_moments,
                float* grads_flat, double* metrics, void* workspace, int64_t workspace_bytes, void* stream);
/* REINFORCEAgent.losses_for_batch (agents/reinforce/reinforce_agent.py:11-88); ret_moments/adv_moments as above */
int gs_reinforce_step(const gs_mlp_t* mlp, const gs_batch_t* batch, const gs_reinforce_hparams_t* hp,
                      const double* ret_moments, const double* adv_moments, float* grads_flat, double* metrics,
                      void* workspace, int64_t workspace_bytes, void* stream);
/* BaseModel.compute_grad_norms (utils/models.py:196-230) + clip_gradients(..., "norm")
 * (agents/base_agent.py:604-617): writes GS_M_GRAD_NORM_* / GS_M_CLIP_COEF into metrics and scales
 * grads_flat by min(1, max_norm/(norm+1e-6)) when max_norm > 0. */
int gs_clip_grad_norm(const gs_mlp_t* mlp, float* grads_flat, float max_norm, double* metrics, void* stream);
/* torch.optim.Adam single-tensor math (utils/optimizer_factory.py:6-29, eps 1e-8) on the flat vectors; the
 * default host path keeps torch.optim, this is the graph-capturable variant. step_count: device int64[1]. */
int gs_adam_step(float* params_flat, const float* grads_flat, float* exp_avg, float* exp_avg_sq, int64_t n,
                 int64_t* step_count, float lr, float beta1, float beta2, float eps, void* stream);

/* ---- fused step tail: BaseAgent._backpropagate_and_step (agents/base_agent.py:591-621) in ONE launch -----------------
 * After a step called with batch->defer_reduce = 1, gs_update_finish
 *   1. sums the per-CTA partial gradient vectors in CTA order (deterministic),
 *   2. with a peer group: writes the local gradient into every rank's receive slot over NVLink (P2P stores), signals,
 *      waits for the other ranks' signals and sums the slots in rank order, divided by world_size (every rank computes the
 *      bit-identical mean: weights stay identical without a broadcast) -- the gradient all-reduce of SURVEY.md 8(e),
 *   3. finalises the metric vector (as gs_ppo_step / gs_reinforce_step do) and adds it into metrics_sum (nullable),
 *   4. takes the group norms (utils/models.py:196-230), applies the global-norm clip (agents/base_agent.py:604-617),
 *      leaves the clipped mean gradient in grads_flat,
 *   5. applies torch.optim.Adam's update (utils/optimizer_factory.py:6-29) when adam != NULL and bumps *adam->step_count.
 * Every rank of a peer group must make the same sequence of gs_update_finish calls. */
typedef struct gs_adam {
    float*   params_flat;
    float*   exp_avg;
    float*   exp_avg_sq;
    int64_t* step_count; /* device int64[1]: steps taken so far */
    float lr, beta1, beta2, eps;
} gs_adam_t;

typedef struct gs_finish {
    int32_t algo;              /* 0 = PPO, 1 = REINFORCE (which metric slots the deferred step fills)        */
    int32_t track_activations; /* as passed to the deferred step                                             */
    int32_t normalize_adv;
    int32_t normalize_ret;
    float   vf_coef;
    float   ent_coef;
    float   max_grad_norm;     /* <= 0: no clip (norms are still reported)                                   */
    int32_t reserved_;
} gs_finish_t;

/* NVLink peer group for the gradient exchange: every rank allocates one device buffer of receive slots -- [2 phases][world_size][max_floats]
 * 8-byte words {fp32 bits, call epoch}: the data carries its own arrival flags -- (gs_peer_create), the 64-byte CUDA IPC handles are exchanged by the caller (any host channel, e.g. an all_gather over
 * the torch.distributed store), gs_peer_connect maps the other ranks' buffers.  world_size <= 8 (one NVSwitch domain). */
typedef struct gs_peer gs_peer_t;
#define GS_PEER_HANDLE_BYTES 64
#define GS_PEER_MAX_WORLD 8
int gs_peer_create(int rank, int world_size, int64_t max_floats, int device, gs_peer_t** out, void* handle_out_host);
int gs_peer_connect(gs_peer_t* peer, const void* all_handles_host /* world_size * GS_PEER_HANDLE_BYTES, rank order */);
int gs_peer_destroy(gs_peer_t* peer);
/* In-place sum over the ranks of a short fp64 vector (n <= 8192; e.g. the (sum, sumsq, count) moments of a pass's minibatches, so that
 * "batch" advantage normalisation -- agents/ppo/ppo_agent.py:47-48 -- sees the GLOBAL minibatch), exchanged through the peer buffers by a
 * one-block kernel that co-resides with the update kernel; rank-ordered sum: every rank holds the identical result.  Every rank must make
 * the same sequence of calls. */
int gs_peer_allreduce_f64(gs_peer_t* peer, double* data, int64_t n, void* stream);
int gs_update_finish(const gs_mlp_t* mlp, const gs_batch_t* batch /* the deferred step's minibatch */, const gs_finish_t* fin,
                     float* grads_flat, const gs_adam_t* adam /* nullable: no optimizer step */,
                     gs_peer_t* peer /* nullable: single rank */, double* metrics, double* metrics_sum /* nullable */,
                     void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GS_ENGINE_H */
