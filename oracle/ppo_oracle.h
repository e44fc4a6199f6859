/*
 * ppo_oracle.h — CPU restatement of ppo_cpp's PPO hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and there only as the checker / the CPU baseline — never on the product path
 * (ppo_cpp_b200/ never imports, links or dlopens anything under oracle/).
 *
 * PARITY UNPINNED: all network arithmetic of the reference lives in TensorFlow 1.14 (not vendored,
 * not installable here) and its host arithmetic in Eigen (absent), and the reference ships no test
 * that pins numerical results of this path.  This file therefore restates the
 * algorithm from the reference's sources and its .meta.txt graph; it is pinned only on
 *   - the graph's baked initial weights / the shipped checkpoint (analytic known answers),
 *   - glibc rand()/libstdc++ std::random_shuffle run in this container (bit-exact),
 *   - Random123 Philox4x32-10 known-answer vectors,
 *   - an independent PyTorch fp64 autograd evaluation of the same loss (gradients).
 *
 * Every function exists in two precisions: *_f32 mirrors the reference's fp32 operation order
 * (compiled with -ffp-contract=off so x86 FMA does not change roundings), *_f64 does the same math
 * in double and is the "truth" the ≤1e-5 relative tolerance is measured against.
 *
 * Citations are file:line in /root/reference; GRAPH:n is a line of
 * resources/ppo_cl/graphs/ppo_cpp_[4_5]_lr_0.0004_cr_0.1610_ent_0.0007.meta.txt.
 */
#ifndef PPO_ORACLE_H
#define PPO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- network shape and flat parameter layout (gradient order of GRAPH:23738-24074, then q) ---- */
typedef struct {
    int obs_dim, act_dim, h1, h2;
} oracle_dims;

enum {
    OT_PI_FC0_W, OT_PI_FC0_B, OT_VF_FC0_W, OT_VF_FC0_B, OT_PI_FC1_W, OT_PI_FC1_B, OT_VF_FC1_W, OT_VF_FC1_B,
    OT_VF_W, OT_VF_B, OT_PI_W, OT_PI_B, OT_LOGSTD, OT_Q_W, OT_Q_B, OT_COUNT
};
/* offset (floats) of tensor t in the flat vector; t == OT_COUNT gives the total size (with q head),
 * t == OT_Q_W gives the trainable size P. */
int oracle_param_offset(const oracle_dims *d, int t);

/* ---- RNG ---- */
/* Philox4x32-10 (Salmon et al. SC'11; same generator TF's RandomStandardNormal uses). */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The core's N(0,1) stream: 2*ceil(act_dim/4)... see DESIGN.md "Philox stream".  Fills eps[act_dim] for
 * (seed, global env id, global step counter).  Box-Muller on pairs of 32-bit words, TF-style
 * (u1 = max(u32->[0,1), 1e-7), r = sqrt(-2 ln u1), theta = 2 pi u2, (r sin, r cos)). */
void oracle_normal_eps(uint64_t seed, uint32_t env_id, uint32_t step, int act_dim, float *eps);
/* U[0,1) from a 32-bit word exactly as TF's Uint32ToFloat (23 mantissa bits). */
float oracle_u32_to_unit_float(uint32_t x);

/* glibc TYPE_3 rand() restated (stdlib/random_r.c), state is explicit so tests can interleave. */
typedef struct {
    uint32_t ring[31]; /* r[i] = r[i-3] + r[i-31] additive feedback ring */
    int fi, ri;        /* front / rear positions */
} oracle_glibc_rand;
void oracle_srand(oracle_glibc_rand *st, unsigned seed);
int oracle_rand(oracle_glibc_rand *st);
/* libstdc++ std::random_shuffle(first,last) (bits/stl_algo.h:4581-4600): j = rand() % (i+1), swap. */
void oracle_random_shuffle(oracle_glibc_rand *st, int *a, int n);
/* Eigen `perm * v` semantics used at ppo2/ppo2.hpp:291-296: out[perm[i]] = in[i].  Returns the gather
 * list src[] with permuted_row[s] = original_row[src[s]]. */
void oracle_perm_to_gather(const int *perm, int n, int *src);

/* ---- policy (ppo2/policies.hpp:33-77 + GRAPH:1859-6866) ---- */
/* eps may be NULL (treated as 0 => action = mean).  Any output may be NULL. */
void oracle_policy_step_f32(const oracle_dims *d, const float *params, const float *obs, int n, const float *eps,
                            float *action, float *value, float *neglogp, float *mean);
void oracle_policy_step_f64(const oracle_dims *d, const float *params, const float *obs, int n, const float *eps,
                            double *action, double *value, double *neglogp, double *mean);

/* ---- VecNormalize (env/env_normalize.hpp:64-116, common/running_statistics.hpp:26-104,
 *      common/matrix_clamp.hpp:32-35) ---- */
typedef struct {
    int dim;
    float *mean;  /* [dim] */
    float *var;   /* [dim] */
    double count; /* initialised to 1e-6 (running_statistics.hpp:17-20) */
} oracle_rstats;
void oracle_rstats_update_f32(oracle_rstats *s, const float *batch, int rows);
/* double-precision statistics kept in caller arrays */
void oracle_rstats_update_f64(double *mean, double *var, double *count, int dim, const double *batch, int rows);

typedef struct {
    int n_envs, obs_dim;
    int training, norm_obs, norm_reward;
    float clip_obs, clip_reward, gamma, epsilon;
    oracle_rstats obs_rms, ret_rms;
    float *ret; /* [n_envs] discounted return accumulator */
} oracle_vecnorm;
oracle_vecnorm *oracle_vecnorm_create(int n_envs, int obs_dim, int training);
void oracle_vecnorm_destroy(oracle_vecnorm *v);
/* EnvNormalize::reset (env_normalize.hpp:111-116) */
void oracle_vecnorm_reset_f32(oracle_vecnorm *v, const float *raw_obs, float *obs_out);
/* EnvNormalize::step after the inner env returned (env_normalize.hpp:64-92) */
void oracle_vecnorm_step_f32(oracle_vecnorm *v, const float *raw_obs, const float *raw_rew, const float *done,
                             float *obs_out, float *rew_out);
void oracle_matrix_clamp_f32(const float *x, int n, float lo, float hi, float *out);

/* ---- GAE (ppo2/runner.hpp:159-191); all [n_steps, n_envs] row-major, time-major ---- */
void oracle_gae_f32(const float *rewards, const float *values, const float *dones, const float *last_values,
                    const float *last_dones, int n_steps, int n_envs, float gamma, float lam, float *advs,
                    float *returns);
void oracle_gae_f64(const float *rewards, const float *values, const float *dones, const float *last_values,
                    const float *last_dones, int n_steps, int n_envs, double gamma, double lam, double *advs,
                    double *returns);

/* ---- advantage normalisation per minibatch (ppo2/ppo2.hpp:401-406) ---- */
void oracle_advnorm_f32(const float *returns, const float *values, int n, float *advs);
void oracle_advnorm_f64(const float *returns, const float *values, int n, double *advs);

/* ---- loss + hand-derived backward (GRAPH:9210-23699) ----
 * losses[5] = pg_loss, vf_loss, entropy, approxkl, clipfrac; grads has P (trainable) entries. */
typedef struct {
    float ent_coef, vf_coef, clip_norm, beta1, beta2, adam_eps;
} oracle_hparams;
void oracle_loss_grad_f32(const oracle_dims *d, const oracle_hparams *hp, const float *params, const float *obs,
                          const float *actions, const float *advs, const float *returns, const float *old_neglogp,
                          const float *old_values, int B, float cliprange, float *grads, float *losses);
void oracle_loss_grad_f64(const oracle_dims *d, const oracle_hparams *hp, const float *params, const float *obs,
                          const float *actions, const double *advs, const float *returns, const float *old_neglogp,
                          const float *old_values, int B, double cliprange, double *grads, double *losses);

/* ---- clip_by_global_norm + ApplyAdam (GRAPH:23738-31383; TF 1.14 training_ops.cc ApplyAdam) ----
 * grads is scaled in place; returns the pre-clip global norm. */
float oracle_clip_adam_f32(const oracle_hparams *hp, int P, float lr, float *params, float *m, float *v, float *grads,
                           float *beta1_power, float *beta2_power);
double oracle_clip_adam_f64(const oracle_hparams *hp, int P, double lr, double *params, double *m, double *v,
                            double *grads, double *beta1_power, double *beta2_power);

/* ---- synthetic 18-dim env (SURVEY §8d "Synthetic inputs"; mirrors the hexapod's reward/episode shape) ---- */
typedef struct {
    int n_envs, dim;
    uint64_t seed;
    uint32_t env_id0; /* global id of local env 0 (rank offset) */
    float *state;     /* [n_envs, dim] */
    uint32_t *t_env;  /* per-env step counter (includes the phase offset) */
    uint32_t *resets; /* per-env number of resets so far */
} oracle_synth_env;
oracle_synth_env *oracle_synth_env_create(int n_envs, int dim, uint64_t seed, uint32_t env_id0);
void oracle_synth_env_destroy(oracle_synth_env *e);
void oracle_synth_env_reset(oracle_synth_env *e, float *obs);
void oracle_synth_env_step(oracle_synth_env *e, const float *actions, float *obs, float *rew, float *done);

/* ---- whole learner, structured like the reference (ppo2/ppo2.hpp:239-377, runner.hpp:56-157):
 *      one policy call per env step at batch n_envs, sequential GAE, per-epoch random_shuffle with
 *      full-buffer permute + slice copies, one train step per minibatch.  Used as the CPU baseline
 *      ("port") and as the end-to-end oracle at small sizes. ---- */
typedef struct oracle_learner oracle_learner;
typedef struct {
    oracle_dims dims;
    oracle_hparams hp;
    int n_envs, n_steps, nminibatches, noptepochs;
    float gamma, lam, lr, cliprange;
    uint64_t seed;        /* Philox seed for action noise and the synthetic env */
    unsigned shuffle_seed; /* srand() seed */
    int env_kind;         /* 0 = synthetic env, 1 = EnvMock(scaling 1.0) (env/env_mock.hpp:43-60) */
    int threads;          /* OpenMP threads for the batched math (<=0: all) */
} oracle_learner_desc;
oracle_learner *oracle_learner_create(const oracle_learner_desc *desc, const float *params_with_q);
void oracle_learner_destroy(oracle_learner *L);
/* Runner::run(): fills the rollout buffers (flat row = env*n_steps + t). */
void oracle_learner_rollout(oracle_learner *L);
/* the epoch/minibatch loop of PPO2::learn for the current rollout; mean losses[5] out. */
void oracle_learner_train(oracle_learner *L, float *mean_losses);
/* buffers: 0 obs,1 returns,2 dones,3 actions,4 values,5 neglogpacs,6 true_rewards,7 unnormalized_rewards */
const float *oracle_learner_buffer(oracle_learner *L, int which);
float *oracle_learner_params(oracle_learner *L);
int oracle_learner_threads(oracle_learner *L);
void oracle_learner_get_norm(oracle_learner *L, float *obs_mean, float *obs_var, double *obs_count, float *ret_mean,
                             float *ret_var, double *ret_count);

#ifdef __cplusplus
}
#endif
#endif
