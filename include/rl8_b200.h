/*
 * rl8_b200 — C-ABI of the B200-native PPO rollout-and-update engine.
 *
 * The upstream reference (theOGognf/rl8) is pure Python on stock PyTorch and has no FFI
 * of its own; its plug-in points are Python protocols.  This header is
 * the boundary those Python classes bind underneath: every entry point below replaces
 * the body of one reference function (cited as path:line relative to the upstream
 * repository root) and is what a ctypes / cffi stub in the reference would call
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 (RL8_OK) or a negative rl8_status; nothing throws;
 *   - no function allocates, frees or synchronises; all tensor pointers are DEVICE
 *     pointers owned by the caller (torch tensors on the host side);
 *   - the last argument is the cudaStream_t to launch on (the caller's current stream);
 *   - `N` = number of environments, `T` = horizon, `D` = observation width,
 *     `P` = policy-head width (number of discrete actions, or 2 = {mean, log_std raw}),
 *     `H` = hidden width (256 in the reference's default models);
 *   - rollout storage is HORIZON-MAJOR: a field is `[T+1][N]` floats (obs `[T+1][D][N]`),
 *     so one time step is one contiguous, coalesced slab.  The reference's env-major
 *     `[N, T+1, ...]` tensors are strided views of the same memory.
 *   - flattened transition index (the reference's `buffer.reshape(-1)` row,
 *     src/rl8/algorithms/_feedforward.py:473-481) is `row = n*T + t`.
 */
#ifndef RL8_B200_H
#define RL8_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rl8_stream_t; /* cudaStream_t */

typedef enum {
  RL8_OK = 0,
  RL8_ERR_ARG = -1,         /* bad argument (null pointer, unsupported size) */
  RL8_ERR_CUDA = -2,        /* a CUDA launch failed; see rl8_last_error() */
  RL8_ERR_UNSUPPORTED = -3, /* combination outside the fused path */
  RL8_ERR_WORKSPACE = -4    /* caller workspace too small */
} rl8_status;

typedef enum {
  RL8_ENV_DISCRETE_DUMMY = 0,   /* src/rl8/env.py:233-259 */
  RL8_ENV_CONTINUOUS_DUMMY = 1, /* src/rl8/env.py:206-230 */
  RL8_ENV_CARTPOLE = 2,         /* examples/cartpole/env.py */
  RL8_ENV_MOUNTAIN_CAR = 3,     /* examples/mountain_car/env.py */
  RL8_ENV_PENDULUM = 4          /* examples/pendulum/env.py */
} rl8_env_kind;

typedef enum {
  RL8_DIST_CATEGORICAL = 0,    /* src/rl8/distributions.py:125-132 */
  RL8_DIST_NORMAL = 1,         /* src/rl8/distributions.py:135-144 */
  RL8_DIST_SQUASHED_NORMAL = 2 /* src/rl8/distributions.py:147-170 */
} rl8_dist_kind;

typedef enum {
  RL8_PREC_FP32 = 0,   /* CUDA-core fp32 GEMMs (the cross-check of the mode below) */
  RL8_PREC_BF16 = 1,   /* tcgen05 bf16 GEMMs, fp32 accumulate (reference `enable_amp=True`) */
  RL8_PREC_FP32_TC = 2 /* fp32 results on tcgen05 (reference `enable_amp=False`): every fp32 operand of the
                        * 256 x 256 contractions is split into two fp16 pieces after a power-of-two scaling
                        * derived from a measured bound of the tensor (max |obs|, max |W2|, max |dOut|), three
                        * piece products per fp32 product (2e-6 of fp64 at K = 256), fp32 accumulation in tensor
                        * memory, pair MMAs (cta_group::2); everything else is fp32 on CUDA cores as in
                        * RL8_PREC_FP32.  (`make X3_BF16=1`: three bf16 pieces and six products, no scales.) */
} rl8_precision;

/* Environment constants, already rounded the way the reference rounds them (Python
 * double arithmetic between scalars, then one cast to f32 per tensor op).  Index map:
 *   dummy        p[0]=bounds
 *   cartpole     p[0]=force_mag p[1]=gravity p[2]=length p[3]=pole_mass
 *                p[4]=pole_mass_length p[5]=total_mass p[6]=tau p[7]=(4.0/3.0)
 *                p[8]=1.0 if kinematics_integrator != "euler" else 0.0
 *   mountain_car p[0]=force_mag p[1]=goal_position p[2]=goal_velocity p[3]=gravity
 *                p[4]=max_position p[5]=max_speed p[6]=min_position
 *   pendulum     p[0]=dt p[1]=3g/(2l) p[2]=3/(m l^2) p[3]=max_speed p[4]=max_torque
 *                p[5]=pi p[6]=2*pi
 */
typedef struct {
  float p[16];
} rl8_env_cfg;

/* Default feedforward model (src/rl8/models/_feedforward.py:234-383): two independent
 * MLPs D->H->H->P (policy) and D->H->H->1 (value), ReLU after the first two layers.
 * All pointers address one flat fp32 parameter (or gradient) buffer. */
typedef struct {
  int32_t D, H, P;
  const float *pi_w1, *pi_b1, *pi_w2, *pi_b2, *pi_w3, *pi_b3; /* [H,D] [H] [H,H] [H] [P,H] [P] */
  const float *vf_w1, *vf_b1, *vf_w2, *vf_b2, *vf_w3, *vf_b3; /* ... [1,H] [1] */
} rl8_model;

/* PPO hyper-parameters of one update (src/rl8/nn/functional.py:259-363). */
typedef struct {
  float clip_param;
  float dual_clip_param; /* <= 0: disabled (the reference tests truthiness, :335) */
  float entropy_coeff;
  float vf_clip_param;
  float vf_coeff;
  float loss_scale; /* 1 / grad_accumulation_steps (src/rl8/algorithms/_feedforward.py:545) */
} rl8_ppo_hparams;

/* ---- library ------------------------------------------------------------------- */

/* ABI version of this header; bumped on any signature change. */
int rl8_abi_version(void);
/* Text of the last CUDA error seen by this library on the calling thread ("" if none). */
const char* rl8_last_error(void);

/* ---- environments: Env.reset / Env.step (src/rl8/env.py:100-128) ------------------ */

/* state[S][N] <- reset distribution applied to `noise[S][N]` (standard normal for
 * cartpole / mountain_car, U[0,1) for dummy / pendulum), and the initial observation.
 * obs element (n, d) is written at obs[n*obs_stride_n + d*obs_stride_d].
 * Replaces DummyEnv.reset (src/rl8/env.py:197-203), CartPole.reset
 * (examples/cartpole/env.py:128-136), MountainCar.reset (examples/mountain_car/env.py:92-102),
 * Pendulum.reset (examples/pendulum/env.py:94-106). */
int rl8_env_reset(int env_kind, const rl8_env_cfg* cfg, const float* noise, float* state,
                  float* obs, int64_t obs_stride_n, int64_t obs_stride_d, int64_t N,
                  rl8_stream_t stream);

/* obs <- observation of the current state[S][N] (the tail of every reference reset/step,
 * e.g. examples/cartpole/env.py:133-136); used after installing a state by hand. */
int rl8_env_observe(int env_kind, const float* state, float* obs, int64_t obs_stride_n,
                    int64_t obs_stride_d, int64_t N, rl8_stream_t stream);

/* One transition of every environment.  `action` is int64[N] for discrete envs and
 * f32[N] for continuous ones.  Replaces DiscreteDummyEnv.step (src/rl8/env.py:253-259),
 * ContinuousDummyEnv.step (:224-230), examples/cartpole/env.py:12-64,
 * examples/mountain_car/env.py:12-38, examples/pendulum/env.py:12-39. */
int rl8_env_step(int env_kind, const rl8_env_cfg* cfg, float* state, const void* action,
                 float* obs, int64_t obs_stride_n, int64_t obs_stride_d, float* reward,
                 int64_t N, rl8_stream_t stream);

/* ---- action distributions (src/rl8/distributions.py:98-170) ------------------------ */

/* features[B][P] (logits, or {mean, log_std} with log_std already tanh'ed) + injected
 * noise (Exp(1) [B][P] for categorical, N(0,1) [B] otherwise; ignored when
 * `deterministic`) -> action (int64[B] or f32[B]) and logp[B] (may be NULL). */
int rl8_dist_sample(int dist_kind, const float* features, int32_t P, const float* noise,
                    int deterministic, void* action, float* logp, int64_t B,
                    rl8_stream_t stream);

/* logp[B] and entropy[B] (either may be NULL) of given actions.  Entropy of the squashed
 * normal is undefined in the reference (:153-157) -> RL8_ERR_UNSUPPORTED. */
int rl8_dist_logp_entropy(int dist_kind, const float* features, int32_t P, const void* action,
                          float* logp, float* entropy, int64_t B, rl8_stream_t stream);

/* ---- GAE (src/rl8/nn/functional.py:50-123) ----------------------------------------- */

/* rewards/values/advantages/returns hold (T+1) x N elements addressed as
 * ptr[n*stride_n + t*stride_t].  rewards are divided by (reward_scale + 1e-8) IN PLACE,
 * advantages/returns are written for all T+1 slots (A_T = 0, ret_T = V_T).  When
 * `normalize` is set, the first T slots of the advantages are standardised with the
 * unbiased std over N*T elements; `moments` (3 doubles: sum, sum of squares, count) is
 * scratch, zeroed by the caller, and may be all-reduced across ranks between the two
 * calls.  `returns` and `moments` may be NULL.  gamma / gae_lambda / reward_scale are the
 * reference's Python doubles: `gamma*gae_lambda` and `reward_scale + 1e-8` are formed in
 * double and rounded to f32 once, as torch does for scalar operands.
 * Layout decides the kernel: stride_n == 1 (horizon-major) -> one env per lane, sequential
 * in t, 128-bit coalesced; stride_t == 1 (env-major, the reference layout) -> one env row
 * per warp, warp-level reverse scan over the horizon; anything else -> strided sequential. */
int rl8_gae_scan(float* rewards, const float* values, float* advantages, float* returns,
                 int64_t N, int32_t T, int64_t stride_n, int64_t stride_t, double gamma,
                 double gae_lambda, double reward_scale, double* moments, rl8_stream_t stream);
int rl8_gae_normalize(float* advantages, int64_t N, int32_t T, int64_t stride_n, int64_t stride_t,
                      const double* moments, rl8_stream_t stream);

/* The device-resident form (no host round trip between collect() and step()): rl8_reward_scale
 * turns the (all-reduced) accumulator of rl8_collect_stats into out[0] = f32(std(rdr[:, 1:])) -- the reference's
 * float(torch.std(...)), src/rl8/algorithms/_feedforward.py:428-436; 1 when !normalize_rewards -- and
 * out[1] = f32(out[0] + 1e-8); rl8_gae_scan_dev is rl8_gae_scan reading that divisor from `reward_scale_dev[1]`.
 * write_scaled_rewards = 0 (horizon-major layout only) leaves `rewards` untouched: Algorithm.step() discards the
 * rewards right after GAE (src/rl8/algorithms/_feedforward.py:476-478), so the side effect of functional.py:106 is
 * unobservable there and the scan moves SURVEY.md §8d's 16 B per transition instead of 20. */
int rl8_reward_scale(const double* acc, double count, int normalize_rewards, float* out, rl8_stream_t stream);
int rl8_gae_scan_dev(float* rewards, const float* values, float* advantages, float* returns,
                     int64_t N, int32_t T, int64_t stride_n, int64_t stride_t, double gamma,
                     double gae_lambda, const float* reward_scale_dev, int write_scaled_rewards, double* moments,
                     rl8_stream_t stream);

/* ---- collect statistics (src/rl8/algorithms/_feedforward.py:410-436) ----------------- */

/* From horizon-major rewards[T+1][N] and reversed discounted returns rdr[T+1][N] (may be
 * NULL) accumulate into acc[16] doubles (caller sets acc[0..5] = 0, acc[6] = acc[8] = +inf,
 * acc[7] = acc[9] = -inf; sums all-reduce with SUM, extrema with MIN / MAX:
 *   0 sum r, 1 sum r^2, 2 sum R, 3 sum R^2, 4 sum rdr, 5 sum rdr^2 (slots 1..T),
 *   6 min r, 7 max r, 8 min R, 9 max R) where R = per-env sum of rewards over t < T. */
int rl8_collect_stats(const float* rewards, const float* rdr, int64_t N, int32_t T, double* acc,
                      rl8_stream_t stream);
/* Same, with the reward statistics (acc 0..3, 6..9) taken over slots reward_t0 <= t < T: the
 * recurrent algorithm summarises `rewards[:, 1:-1]` (src/rl8/algorithms/_recurrent.py:449), i.e.
 * reward_t0 = 1; the rdr moments still cover slots 1..T. */
int rl8_collect_stats_from(const float* rewards, const float* rdr, int64_t N, int32_t T,
                           int32_t reward_t0, double* acc, rl8_stream_t stream);

/* ---- policy / value networks ------------------------------------------------------- */

/* Bytes of workspace rl8_mlp_forward needs for `rows` rows. */
int64_t rl8_mlp_forward_workspace(int32_t H, int64_t rows);

/* out[rows][P_out] = MLP(obs) for one of the two networks of `model` (which = 0 policy,
 * 1 value).  obs element (r, d) is read at obs[r*obs_stride_r + d*obs_stride_d].
 * Replaces DefaultDiscreteModel.forward / value_function and the continuous twin
 * (src/rl8/models/_feedforward.py:292-310, 365-383).  For the continuous policy head the
 * second output is tanh'ed (log_std, :299) when `apply_tanh_log_std` is set. */
int rl8_mlp_forward(const rl8_model* model, int which, const float* obs, int64_t obs_stride_r,
                    int64_t obs_stride_d, int64_t rows, float* out, int apply_tanh_log_std,
                    int precision, void* workspace, int64_t workspace_bytes, rl8_stream_t stream);

/* ---- collect(): the T-step rollout (src/rl8/algorithms/_feedforward.py:359-408) ------- */

typedef struct {
  int32_t env_kind, dist_kind, T, deterministic;
  int64_t N;
  float gamma;
  int32_t normalize_rewards; /* maintain rdr (src/rl8/algorithms/_feedforward.py:378-383) */
  rl8_env_cfg env_cfg;
  float* env_state;    /* [S][N], advanced in place */
  float* obs;          /* [T+1][D][N]; slab 0 holds the initial observation */
  void* actions;       /* [T+1][N] int64 (discrete) or f32 (continuous) */
  float* logp;         /* [T+1][N] */
  float* values;       /* [T+1][N]; all T+1 slabs are written (bootstrap value included) */
  float* rewards;      /* [T+1][N] */
  float* rdr;          /* [T+1][N] or NULL; slab 0 is an input */
  const float* noise;  /* [T][N][P] Exp(1) (categorical) or [T][N] N(0,1); NULL if deterministic */
} rl8_rollout;

int64_t rl8_collect_workspace(const rl8_model* model, int64_t N, int32_t T, int precision);
int rl8_collect(const rl8_model* model, const rl8_rollout* ro, int precision, void* workspace,
                int64_t workspace_bytes, rl8_stream_t stream);

/* ---- step(): one PPO minibatch (src/rl8/algorithms/_feedforward.py:512-585) ------------ */

typedef struct {
  int32_t dist_kind, T;
  int64_t N;
  const float* obs;        /* [T+1][D][N] */
  const void* actions;     /* [T+1][N] */
  const float* logp;       /* [T+1][N] */
  const float* advantages; /* [T+1][N] */
  const float* returns;    /* [T+1][N] */
} rl8_batch;

int64_t rl8_ppo_workspace(const rl8_model* model, int64_t max_rows, int precision);

/* Forward + PPO losses + backward for the `M` transitions whose flattened indices
 * (row = n*T + t) are rows[0..M) (int64 device array), or row_begin..row_begin+M when
 * `rows` is NULL.  Gradients of `loss_scale * total / world` are ACCUMULATED into
 * `grads` (same flat layout as the parameters; the caller zeroes them at an optimizer
 * step boundary); `mean_denominator` is the global minibatch size the means divide by
 * (M on one GPU, M * world_size when ranks all-reduce gradients).  `loss_sums` receives
 * 5 doubles: sum entropy, sum policy surrogate, sum vf loss, sum kl, count -- raw sums
 * over the M rows, so ranks can all-reduce them. */
int rl8_ppo_minibatch(const rl8_model* model, const rl8_model* grads, const rl8_batch* batch,
                      const int64_t* rows, int64_t row_begin, int64_t M, double mean_denominator,
                      const rl8_ppo_hparams* hp, double* loss_sums, int precision,
                      void* workspace, int64_t workspace_bytes, rl8_stream_t stream);

/* ppo_losses (src/rl8/nn/functional.py:259-363) on given network outputs: features[B][P]
 * (logits | {mean, log_std}), values[B], and the stored batch columns (all [B], unit
 * stride).  loss_sums as above; d_features[B][P] / d_values[B] (either may be NULL) receive
 * d(loss_scale * total)/d(output) with means over `mean_denominator` -- for the continuous
 * head the second column is the gradient w.r.t. the PRE-tanh log_std output. */
int rl8_ppo_losses(int dist_kind, const float* features, int32_t P, const float* values,
                   const void* actions, const float* logp_old, const float* advantages,
                   const float* returns, int64_t B, double mean_denominator,
                   const rl8_ppo_hparams* hp, double* loss_sums, float* d_features,
                   float* d_values, rl8_stream_t stream);

/* Same, for features produced by a user-defined model (the reference's `Model` plug-in point,
 * src/rl8/models/_feedforward.py:146-162, whose `log_std` is the model's own output): the second
 * column of d_features is the gradient w.r.t. log_std itself. */
int rl8_ppo_losses_direct(int dist_kind, const float* features, int32_t P, const float* values,
                          const void* actions, const float* logp_old, const float* advantages,
                          const float* returns, int64_t B, double mean_denominator,
                          const rl8_ppo_hparams* hp, double* loss_sums, float* d_features,
                          float* d_values, rl8_stream_t stream);

/* ---- recurrent policy: RecurrentAlgorithm (src/rl8/algorithms/_recurrent.py) ----------------- */

/* Default recurrent models (src/rl8/models/_recurrent.py:169-341): ONE nn.LSTM(D, H) whose
 * latent h_t feeds the policy head ([P,H]: A logits, or {action_mean; action_log_std} stacked)
 * and the value head ([1,H]).  torch's gate packing: rows [0,H) input, [H,2H) forget,
 * [2H,3H) cell ("g"), [3H,4H) output.  All pointers address one flat fp32 buffer. */
typedef struct {
  int32_t D, H, P;
  const float *w_ih, *w_hh, *b_ih, *b_hh; /* [4H,D] [4H,H] [4H] [4H] */
  const float *pi_w, *pi_b;               /* [P,H] [P] */
  const float *vf_w, *vf_b;               /* [1,H] [1] */
} rl8_lstm_model;

/* RecurrentAlgorithm.collect (src/rl8/algorithms/_recurrent.py:356-445).  `ro` is the
 * feedforward rollout block (same fields, same layouts).  hidden / cell are horizon-major
 * [T+1][N][H]: slab t is the state the policy CONSUMES at step t (slab 0 is an input: the host
 * copies slab T of the previous collect into it, :380-382); slab t+1 receives the state the
 * LSTM produced at step t.  Slab t is zeroed first when `t % seq_len == 0 and (seqs + t /
 * seq_len) % seqs_per_state_reset == 0` -- unless seqs_per_state_reset < 0 and the running
 * count is non-zero (:384-392).  `seqs` is RecurrentAlgorithmState.seqs at entry; the host
 * advances it by T / seq_len afterwards (:430-431). */
typedef struct {
  rl8_rollout ro;
  float* hidden; /* [T+1][N][H] */
  float* cell;   /* [T+1][N][H] */
  int32_t seq_len, seqs_per_state_reset;
  int64_t seqs;
} rl8_recurrent_rollout;

int64_t rl8_lstm_collect_workspace(const rl8_lstm_model* model, int64_t N, int32_t T, int precision);
int rl8_lstm_collect(const rl8_lstm_model* model, const rl8_recurrent_rollout* rro, int precision,
                     void* workspace, int64_t workspace_bytes, rl8_stream_t stream);

/* One step of the LSTM + heads on B rows (RecurrentPolicy.sample with a length-1 sequence,
 * src/rl8/policies/_recurrent.py:68-164): obs (r, d) at obs[r*obs_stride_r + d*obs_stride_d],
 * h_in/c_in [B][H] -> h_out/c_out [B][H] (may alias the inputs), features[B][P] (log_std
 * column tanh'ed when apply_tanh_log_std), values[B].  workspace: B*4H floats. */
int rl8_lstm_forward(const rl8_lstm_model* model, const float* obs, int64_t obs_stride_r,
                     int64_t obs_stride_d, const float* h_in, const float* c_in, float* h_out,
                     float* c_out, float* features, float* values, int64_t B,
                     int apply_tanh_log_std, int precision, void* workspace,
                     int64_t workspace_bytes, rl8_stream_t stream);

/* One PPO minibatch of SEQUENCES with truncated back-propagation through time
 * (src/rl8/algorithms/_recurrent.py:517-600).  The [N, T] transitions are cut into
 * N*T/seq_len sequences, sequence s = n*(T/seq_len) + chunk covering steps chunk*seq_len ..
 * +seq_len-1 of env n (`buffer.reshape(-1, seq_len)`, :518).  The M sequences seqs[0..M)
 * (int64 device array; seq_begin.. when NULL) are replayed from their stored chunk-start state
 * hidden/cell[chunk*seq_len][n] (no gradient flows into stored states), the PPO losses are
 * taken over all M*seq_len rows, and gradients of `loss_scale * total / world` are ACCUMULATED
 * into `grads`.  mean_denominator = global number of ROWS the means divide by (M*seq_len on
 * one GPU).  loss_sums as rl8_ppo_minibatch. */
typedef struct {
  rl8_batch b;
  const float* hidden; /* [T+1][N][H] */
  const float* cell;   /* [T+1][N][H] */
  int32_t seq_len;
} rl8_recurrent_batch;

int64_t rl8_lstm_ppo_workspace(const rl8_lstm_model* model, int64_t max_seqs, int32_t seq_len,
                               int precision);
int rl8_lstm_ppo_minibatch(const rl8_lstm_model* model, const rl8_lstm_model* grads,
                           const rl8_recurrent_batch* batch, const int64_t* seqs,
                           int64_t seq_begin, int64_t M, double mean_denominator,
                           const rl8_ppo_hparams* hp, double* loss_sums, int precision,
                           void* workspace, int64_t workspace_bytes, rl8_stream_t stream);

/* ---- optimizer: clip_grad_norm_ + Adam (src/rl8/algorithms/_feedforward.py:586-593) ----- */

/* norm_out[0] = global L2 norm of grads (device float).  Then
 *   g *= min(1, max_norm / (norm + 1e-6));  Adam(lr, betas, eps) with torch's default
 *   (non-amsgrad, no weight decay) update order.  `step` is the 1-based update count.
 * Scalars are the reference's Python doubles; bias corrections are formed in double. */
int rl8_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                  double max_norm, double lr, double beta1, double beta2, double eps,
                  int64_t step, float* norm_out, rl8_stream_t stream);

/* rl8_clip_adam with the per-step scalars in device memory, for CUDA-graph replays of the update epochs: the step
 * counter `step_dev[0]` is advanced by the call itself (bias corrections are formed from the new value, in double,
 * as rl8_clip_adam forms them on the host) and the learning rate is read from `lr_dev[0]`.  `scratch`: 16 floats
 * ([0] receives the gradient norm like rl8_clip_adam's norm_out).  Same reference lines as rl8_clip_adam
 * (src/rl8/algorithms/_feedforward.py:586-593). */
int rl8_clip_adam_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                      double max_norm, const double* lr_dev, double beta1, double beta2, double eps,
                      long long* step_dev, float* scratch, rl8_stream_t stream);

/* The clip alone -- torch.nn.utils.clip_grad_norm_(parameters, max_norm) of _feedforward.py:587-589 on the flat
 * gradient buffer: norm_out[0] = ||g||_2, g *= min(1, max_norm / (norm + 1e-6)).  For `optimizer_cls` other than
 * the fused Adam (_feedforward.py:257-260): the caller's torch optimizer then steps on parameters whose .grad are
 * views of `grads`. */
int rl8_clip_grads(float* grads, int64_t count, double max_norm, float* norm_out, rl8_stream_t stream);

/* ---- view requirements (src/rl8/views.py) ------------------------------------------------ */

/* Sliding windows over the time axis of one rollout field:
 *
 *     out[b][w][s][f] = x[b][t_first + w + s][f]      w < count, s < size, f < F
 *     mask[b][w][s]   = (t_first + w + s < 0)          (zero-filled `out` there; mask may be NULL)
 *
 * `x` is [B][T][F] with element strides (stride_b, stride_t, stride_f) -- the horizon-major
 * buffer ([T+1][F][N]: stride_b = 1) and env-major tensors ([N][T+1][F]: stride_f = 1) are both
 * read coalesced; `out` is contiguous [B * count][size][F], `mask` contiguous [B * count][size]
 * bytes.  elem_bytes is 4 or 8 (f32 / i64 fields).  Requires t_first + count + size - 2 < T.
 * One call replaces the unfold + permute + reshape of
 *   RollingWindow.apply_all        src/rl8/views.py:158-193   (t_first = 0, count = T - size + 1)
 *   PaddedRollingWindow.apply_all  :245-281 incl. pad_whole_sequence :91-118
 *                                                          (t_first = -(size - 1), count = T)
 *   pad_last_sequence              :57-88                   (t_first = T - size, count = 1)
 *   pad_whole_sequence             :91-118                  (size = 1, t_first = -pad, count = T + pad)
 */
int rl8_view_windows(const void* x, int32_t elem_bytes, int64_t B, int64_t T, int64_t F,
                     int64_t stride_b, int64_t stride_t, int64_t stride_f, int32_t size,
                     int64_t t_first, int64_t count, void* out, uint8_t* mask, rl8_stream_t stream);

/* ---- test hook ------------------------------------------------------------------------ */

/* D[128][N] = A[128][K] * B[N][K]^T through ONE tcgen05 GEMM (bf16 operands, fp32 accumulate)
 * with each operand staged K-major or MN-major in the shared-memory operand format every
 * tensor-core kernel of this library uses; pins the descriptor encodings in the parity
 * tests.  N % 8 == 0, N <= 256; K % 16 == 0, K <= 256. */
int rl8_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K,
                    int a_mn_major, int b_mn_major, rl8_stream_t stream);

/* D[128][N] = A[128][8] * B[N][8]^T through ONE kind::tf32 instruction (operands rounded to
 * tf32, fp32 accumulate): the layer-1 contraction [obs, 1] * [W1, b1]^T of the tensor-core
 * kernels.  N % 16 == 0, 16 <= N <= 256. */
int rl8_tc_selftest_tf32(const float* A, const float* B, float* D, int32_t N, rl8_stream_t stream);

/* D[256][256] = A[256][K] * B[256][K]^T through tcgen05.mma.cta_group::2 (one cluster of two CTAs, M = 256,
 * each CTA staging its 128 rows of A and its 128-row half of B) with fp32 operands split into three bf16
 * pieces and the piece products selected by `terms` accumulated in fp32 tensor memory (bit 0 a0b0, 1 a0b1,
 * 2 a1b0, 3 a1b1, 4 a0b2, 5 a2b0; 63 = the fp32-accurate bf16 form, 7 = the two-piece form).  Bit 6: the
 * pieces are fp16 (two per operand, term bits 0..3; 64 + 7 = the update kernels' form; the caller keeps the
 * operands inside fp16's range).  Pins the pair plumbing and measures each term set's accuracy.  K % 32 == 0. */
int rl8_tc3_selftest(const float* A, const float* B, float* D, int32_t K, int32_t terms, rl8_stream_t stream);

/* out[128][128] = in[128][128] (32-bit words) through a tensor-memory store / load round trip
 * (tcgen05.st / tcgen05.ld 32x32b.x32), as used to park packed bf16 activations in TMEM. */
int rl8_tc_selftest_tmem(const uint32_t* in, uint32_t* out, rl8_stream_t stream);
/* Test hook: register layout of tcgen05.ld.16x256b.x4.  in[128][32] goes to tensor memory with the 32x32b store
 * (thread = lane); out[tid][16 h + j] is register j of warp tid / 32's load of its lanes 16 h .. 16 h + 15. */
int rl8_tc_selftest_tmem_16x256b(const uint32_t* in, uint32_t* out, rl8_stream_t stream);

/* Microbenchmark: out_cycles[0] = SM cycles for `iters` rounds of tensor-memory reads by `nwarps`
 * warps of one CTA (mode 0: 32x32b.x32 + wait; 1: two loads per wait; 2: x16 loads).  Sizes the
 * epilogue budgets in DESIGN.md. */
int rl8_tc_bench_tmem(long long* out_cycles, int32_t nwarps, int32_t iters, int32_t mode,
                      rl8_stream_t stream);

/* C[M][N] (ldc) = (accumulate ? C : 0) + sum_k A(m, k) * B(k, n) on tcgen05: fp32 operands converted to bf16
 * while staged, fp32 accumulation.  a_kmajor: A(m, k) = A[m * lda + k] else A[k * lda + m]; b_kmajor:
 * B(k, n) = B[n * ldb + k] else B[k * ldb + n].  accumulate splits K over `splits` CTAs (vector
 * reductions into C).  The GEMM the recurrent path uses for `enable_amp=True`; exported for the tests. */
int rl8_tc_gemm(int a_kmajor, int b_kmajor, int accumulate, const float* A, const float* B, float* C,
                int64_t M, int32_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int32_t splits,
                rl8_stream_t stream);

/* Microbenchmark: out_cycles[0] = SM cycles from the first issue to the mbarrier completion of `reps`
 * back-to-back tcgen05 GEMMs D[128][N] += A[128][k_total] * B[N][k_total]^T (bf16, one commit at the end,
 * operands in the library's chunked shared-memory format); out_cycles[1] = cycles of the issue loop.
 * Gives the per-instruction pace by N / operand major and the fixed issue -> commit -> wait latency. */
int rl8_tc_bench_mma(long long* out_cycles, int32_t N, int32_t k_total, int32_t reps, int a_mn_major,
                     int b_mn_major, rl8_stream_t stream);

/* Development hook: `pairs` clusters of two CTAs each issue reps x 2 x terms tcgen05.mma.cta_group::2 instructions
 * (M = 256, N = n_cols, K = 16, bf16) back to back; out[2 p] = SM cycles, out[2 p + 1] = nanoseconds of pair p.
 * Gives the pace of the split kernels' instruction stream with the whole chip busy (power cap included).
 * terms + 16: both operands MN-major ([K rows][128 columns] tiles, the weight-gradient kernel's form). */
int rl8_tc3_bench_pace(long long* out, int32_t pairs, int32_t reps, int32_t terms, int32_t n_cols,
                       rl8_stream_t stream);

/* Development hook (meaningful in `make ABLATE=1` builds): 32 x uint64 on the device (zeroed by the caller) receive, for pair 0 of each network
 * ([0..15] policy, [16..31] value) of x3_update_f_kernel, the SM cycles its MMA warp waited for ring stage kc
 * ([0..7]) and for a free accumulator ([8]), its total cycles ([9]) and its tile count ([10]); NULL switches it off. */
int rl8_x3_debug_buffer(unsigned long long* device_counters);

/* Debug hook: `device_counters` (24 x uint64 on the device, zeroed by the caller) receives the SM
 * cycles CTA 0 of each network spends in the 8 phases of the tensor-core update's activation
 * kernel ([0..7] policy, [8..15] value; [16..20] weight-gradient kernel CTA 0) for every later
 * rl8_ppo_minibatch call; NULL switches the
 * stamping off.  Feeds profiles/ (where the tile time goes), not the product path. */
int rl8_tc_phase_buffer(unsigned long long* device_counters);

#ifdef __cplusplus
}
#endif
#endif /* RL8_B200_H */
