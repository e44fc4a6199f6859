/*
 * ppo_oracle.c — CPU restatement of ppo_cpp's PPO hot path.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED
 * (see ppo_oracle.h for what that means and what the oracle IS pinned on).
 * Build: see oracle/Makefile (exact build: -O2 -ffp-contract=off; baseline build: -O3 -march=native -fopenmp).
 */
#define _GNU_SOURCE
#include "ppo_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* 0.5*log(2*pi) and 0.5*log(2*pi*e) as the fp32 constants baked in the graph (GRAPH:6103-6672, 10021-10180) */
#define HALF_LOG_2PI_F32 0.9189385175704956
#define HALF_LOG_2PIE_F32 1.4189385175704956

int oracle_param_offset(const oracle_dims *d, int t) {
    const int O = d->obs_dim, A = d->act_dim, H1 = d->h1, H2 = d->h2;
    const int sizes[OT_COUNT] = {O * H1, H1, O * H1, H1, H1 * H2, H2, H1 * H2, H2, H2, 1, H2 * A, A, A, H2 * A, A};
    int off = 0;
    for (int i = 0; i < t && i < OT_COUNT; ++i) off += sizes[i];
    return off;
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

float oracle_u32_to_unit_float(uint32_t x) {
    uint32_t bits = (x & 0x7fffffu) | 0x3f800000u; /* [1,2) */
    float f;
    memcpy(&f, &bits, 4);
    return f - 1.0f;
}

#define TAG_ACTION 0x50504F32u /* "PPO2" */
#define TAG_ENVNOISE 0x454E5631u /* "ENV1" */
#define TAG_ENVRESET 0x52535431u /* "RST1" */

static void box_muller(uint32_t w0, uint32_t w1, float *n0, float *n1) {
    float u1 = oracle_u32_to_unit_float(w0);
    if (u1 < 1.0e-7f) u1 = 1.0e-7f;
    float u2 = oracle_u32_to_unit_float(w1);
    float r = sqrtf(-2.0f * logf(u1));
    float th = 6.2831853071795864769f * u2;
    *n0 = r * sinf(th);
    *n1 = r * cosf(th);
}

static void normal_stream(uint64_t seed, uint32_t a, uint32_t b, uint32_t tag, int count, float *out) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int blk = 0; blk * 4 < count; ++blk) {
        uint32_t ctr[4] = {a, b, (uint32_t)blk, tag}, w[4];
        float n[4];
        oracle_philox4x32_10(ctr, key, w);
        box_muller(w[0], w[1], &n[0], &n[1]);
        box_muller(w[2], w[3], &n[2], &n[3]);
        for (int i = 0; i < 4 && blk * 4 + i < count; ++i) out[blk * 4 + i] = n[i];
    }
}

void oracle_normal_eps(uint64_t seed, uint32_t env_id, uint32_t step, int act_dim, float *eps) {
    normal_stream(seed, env_id, step, TAG_ACTION, act_dim, eps);
}

/* ------------------------------------------------------------------ glibc rand / random_shuffle */
void oracle_srand(oracle_glibc_rand *st, unsigned seed) {
    int32_t word = seed ? (int32_t)seed : 1;
    st->ring[0] = (uint32_t)word;
    for (int i = 1; i < 31; ++i) {
        long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        st->ring[i] = (uint32_t)word;
    }
    st->fi = 3;
    st->ri = 0;
    for (int k = 0; k < 310; ++k) (void)oracle_rand(st);
}

int oracle_rand(oracle_glibc_rand *st) {
    st->ring[st->fi] += st->ring[st->ri];
    uint32_t result = st->ring[st->fi] >> 1;
    if (++st->fi >= 31) {
        st->fi = 0;
        ++st->ri;
    } else if (++st->ri >= 31) {
        st->ri = 0;
    }
    return (int)result;
}

void oracle_random_shuffle(oracle_glibc_rand *st, int *a, int n) {
    for (int i = 1; i < n; ++i) {
        int j = oracle_rand(st) % (i + 1);
        if (i != j) {
            int t = a[i];
            a[i] = a[j];
            a[j] = t;
        }
    }
}

void oracle_perm_to_gather(const int *perm, int n, int *src) {
    for (int i = 0; i < n; ++i) src[perm[i]] = i;
}

/* ------------------------------------------------------------------ precision-generic part */
#define REAL float
#define SUF(x) x##_f32
#define R_TANH tanhf
#define R_EXP expf
#define R_SQRT sqrtf
#define R_FABS fabsf
#include "ppo_oracle_impl.inc"
#undef REAL
#undef SUF
#undef R_TANH
#undef R_EXP
#undef R_SQRT
#undef R_FABS

#define REAL double
#define SUF(x) x##_f64
#define R_TANH tanh
#define R_EXP exp
#define R_SQRT sqrt
#define R_FABS fabs
#include "ppo_oracle_impl.inc"
#undef REAL
#undef SUF
#undef R_TANH
#undef R_EXP
#undef R_SQRT
#undef R_FABS

/* ------------------------------------------------------------------ RunningStatistics / EnvNormalize */
/* common/running_statistics.hpp:26-54,88-104.  `Mat * double` converts the double to the matrix
 * scalar (float) before multiplying, one scalar at a time, left to right. */
void oracle_rstats_update_f32(oracle_rstats *s, const float *batch, int rows) {
    const int D = s->dim;
    const double batch_count = (double)rows;
    const double total = s->count + batch_count;
    for (int c = 0; c < D; ++c) {
        float sum = 0.f;
        for (int r = 0; r < rows; ++r) sum += batch[(size_t)r * D + c];
        float bmean = sum / (float)rows; /* colwise().mean() */
        float m2 = 0.f;
        for (int r = 0; r < rows; ++r) {
            float dlt = batch[(size_t)r * D + c] - bmean;
            m2 += dlt * dlt;
        }
        float bvar = m2 / (float)batch_count; /* get_m2(...) / double */
        float delta = bmean - s->mean[c];
        float new_mean = s->mean[c] + delta * (float)batch_count / (float)total;
        float m_a = s->var[c] * (float)s->count;
        float m_b = bvar * (float)batch_count;
        float m_2 = m_a + m_b + delta * delta * (float)s->count * (float)batch_count / (float)total;
        s->mean[c] = new_mean;
        s->var[c] = m_2 / (float)total;
    }
    s->count = batch_count + s->count;
}

void oracle_rstats_update_f64(double *mean, double *var, double *count, int dim, const double *batch, int rows) {
    const double bc = (double)rows, total = *count + bc;
    for (int c = 0; c < dim; ++c) {
        double sum = 0;
        for (int r = 0; r < rows; ++r) sum += batch[(size_t)r * dim + c];
        double bmean = sum / bc, m2 = 0;
        for (int r = 0; r < rows; ++r) {
            double dl = batch[(size_t)r * dim + c] - bmean;
            m2 += dl * dl;
        }
        double bvar = m2 / bc, delta = bmean - mean[c];
        double nm = mean[c] + delta * bc / total;
        double M2 = var[c] * *count + bvar * bc + delta * delta * *count * bc / total;
        mean[c] = nm;
        var[c] = M2 / total;
    }
    *count = total;
}

static void rstats_init(oracle_rstats *s, int dim) {
    s->dim = dim;
    s->mean = (float *)calloc((size_t)dim, sizeof(float));
    s->var = (float *)malloc(sizeof(float) * dim);
    for (int i = 0; i < dim; ++i) s->var[i] = 1.f;
    s->count = (double)1e-6f; /* `float epsilon=1e-6` converted to the double member (running_statistics.hpp:17-20) */
}

oracle_vecnorm *oracle_vecnorm_create(int n_envs, int obs_dim, int training) {
    oracle_vecnorm *v = (oracle_vecnorm *)calloc(1, sizeof(*v));
    v->n_envs = n_envs;
    v->obs_dim = obs_dim;
    v->training = training;
    v->norm_obs = v->norm_reward = 1;
    v->clip_obs = v->clip_reward = 10.f;
    v->gamma = 0.99f;
    v->epsilon = 1e-8f;
    rstats_init(&v->obs_rms, obs_dim);
    rstats_init(&v->ret_rms, 1);
    v->ret = (float *)calloc((size_t)n_envs, sizeof(float));
    return v;
}

void oracle_vecnorm_destroy(oracle_vecnorm *v) {
    if (!v) return;
    free(v->obs_rms.mean); free(v->obs_rms.var); free(v->ret_rms.mean); free(v->ret_rms.var); free(v->ret);
    free(v);
}

void oracle_matrix_clamp_f32(const float *x, int n, float lo, float hi, float *out) {
    for (int i = 0; i < n; ++i) {
        float y = x[i] > lo ? x[i] : lo; /* cwiseMax(lo) */
        out[i] = y < hi ? y : hi;        /* .cwiseMin(hi)  (matrix_clamp.hpp:32-35) */
    }
}

static void normalize_obs(oracle_vecnorm *v, const float *raw, float *out) {
    const int D = v->obs_dim, N = v->n_envs;
    if (!v->norm_obs) {
        memcpy(out, raw, sizeof(float) * (size_t)N * D);
        return;
    }
    if (v->training) oracle_rstats_update_f32(&v->obs_rms, raw, N);
    for (int c = 0; c < D; ++c) {
        float inv = 1.0f / sqrtf(v->obs_rms.var[c] + v->epsilon); /* cwiseSqrt().cwiseInverse() */
        for (int r = 0; r < N; ++r) out[(size_t)r * D + c] = (raw[(size_t)r * D + c] - v->obs_rms.mean[c]) * inv;
    }
    oracle_matrix_clamp_f32(out, N * D, -v->clip_obs, v->clip_obs, out);
}

void oracle_vecnorm_reset_f32(oracle_vecnorm *v, const float *raw_obs, float *obs_out) {
    memset(v->ret, 0, sizeof(float) * (size_t)v->n_envs);
    normalize_obs(v, raw_obs, obs_out);
}

void oracle_vecnorm_step_f32(oracle_vecnorm *v, const float *raw_obs, const float *raw_rew, const float *done,
                             float *obs_out, float *rew_out) {
    const int N = v->n_envs;
    for (int i = 0; i < N; ++i) v->ret[i] = v->ret[i] * v->gamma + raw_rew[i]; /* env_normalize.hpp:71 */
    normalize_obs(v, raw_obs, obs_out);                                         /* :74 */
    if (v->norm_reward) {
        if (v->training) oracle_rstats_update_f32(&v->ret_rms, v->ret, N); /* :77 */
        float inv = 1.0f / sqrtf(v->ret_rms.var[0] + v->epsilon);
        for (int i = 0; i < N; ++i) rew_out[i] = raw_rew[i] * inv; /* :80 (no mean subtraction) */
        oracle_matrix_clamp_f32(rew_out, N, -v->clip_reward, v->clip_reward, rew_out);
    } else {
        memcpy(rew_out, raw_rew, sizeof(float) * (size_t)N);
    }
    for (int i = 0; i < N; ++i) v->ret[i] = v->ret[i] * (1.0f - done[i]); /* :88 */
}

/* ------------------------------------------------------------------ synthetic env */
oracle_synth_env *oracle_synth_env_create(int n_envs, int dim, uint64_t seed, uint32_t env_id0) {
    oracle_synth_env *e = (oracle_synth_env *)calloc(1, sizeof(*e));
    e->n_envs = n_envs;
    e->dim = dim;
    e->seed = seed;
    e->env_id0 = env_id0;
    e->state = (float *)calloc((size_t)n_envs * dim, sizeof(float));
    e->t_env = (uint32_t *)calloc((size_t)n_envs, sizeof(uint32_t));
    e->resets = (uint32_t *)calloc((size_t)n_envs, sizeof(uint32_t));
    return e;
}

void oracle_synth_env_destroy(oracle_synth_env *e) {
    if (!e) return;
    free(e->state); free(e->t_env); free(e->resets); free(e);
}

static void synth_reset_one(oracle_synth_env *e, int i) {
    uint32_t key[2] = {(uint32_t)e->seed, (uint32_t)(e->seed >> 32)};
    uint32_t gid = e->env_id0 + (uint32_t)i;
    for (int blk = 0; blk * 4 < e->dim; ++blk) {
        uint32_t ctr[4] = {gid, e->resets[i], (uint32_t)blk, TAG_ENVRESET}, w[4];
        oracle_philox4x32_10(ctr, key, w);
        for (int k = 0; k < 4 && blk * 4 + k < e->dim; ++k) {
            float u = 2.0f * oracle_u32_to_unit_float(w[k]) - 1.0f;
            e->state[(size_t)i * e->dim + blk * 4 + k] = 0.1f * u; /* reset to 0.1*U(-1,1) */
        }
    }
    e->resets[i] += 1;
}

void oracle_synth_env_reset(oracle_synth_env *e, float *obs) {
    for (int i = 0; i < e->n_envs; ++i) {
        e->resets[i] = 0;
        e->t_env[i] = (e->env_id0 + (uint32_t)i) % 334u; /* per-env phase offset */
        synth_reset_one(e, i);
    }
    memcpy(obs, e->state, sizeof(float) * (size_t)e->n_envs * e->dim);
}

void oracle_synth_env_step(oracle_synth_env *e, const float *actions, float *obs, float *rew, float *done) {
    const int D = e->dim;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < e->n_envs; ++i) {
        float xi[32];
        float *s = e->state + (size_t)i * D;
        normal_stream(e->seed, e->env_id0 + (uint32_t)i, e->t_env[i], TAG_ENVNOISE, D, xi);
        float s0 = s[0];
        for (int k = 0; k < D; ++k) {
            float a = actions[(size_t)i * D + k];
            a = a < -1.f ? -1.f : (a > 1.f ? 1.f : a);
            s[k] = (0.9f * s[k] + 0.1f * a) + 0.01f * xi[k];
        }
        rew[i] = s[0] - s0;
        e->t_env[i] += 1;
        if (e->t_env[i] % 334u == 0u) {
            done[i] = 1.f;
            synth_reset_one(e, i);
        } else {
            done[i] = 0.f;
        }
        memcpy(obs + (size_t)i * D, s, sizeof(float) * D);
    }
}

/* ------------------------------------------------------------------ whole learner (reference structure) */
struct oracle_learner {
    oracle_learner_desc d;
    int P, Pq, n_batch;
    float *params, *m, *v, b1p, b2p;
    oracle_vecnorm *norm;
    oracle_synth_env *env;
    long mock_total_step;
    oracle_glibc_rand rng;
    float *cur_obs, *cur_dones; /* Runner::obs, Runner::dones */
    uint32_t global_step;
    /* rollout buffers, final layout flat row = env*n_steps + t */
    float *buf[8];
    int threads;
};

static void env_reset(oracle_learner *L, float *raw_obs) {
    const int N = L->d.n_envs, O = L->d.dims.obs_dim;
    if (L->d.env_kind == 0) {
        oracle_synth_env_reset(L->env, raw_obs);
    } else {
        for (int i = 0; i < N * O; ++i) raw_obs[i] = 1.0f;
    }
}

static void env_step(oracle_learner *L, const float *actions, float *raw_obs, float *raw_rew, float *done) {
    const int N = L->d.n_envs, O = L->d.dims.obs_dim;
    if (L->d.env_kind == 0) {
        oracle_synth_env_step(L->env, actions, raw_obs, raw_rew, done);
    } else { /* EnvMock(1.0): env/env_mock.hpp:43-60 */
        L->mock_total_step += 1;
        for (int i = 0; i < N * O; ++i) raw_obs[i] = 1.0f;
        for (int i = 0; i < N; ++i) {
            raw_rew[i] = 1.0f;
            done[i] = (L->mock_total_step % 300 == 0) ? 1.f : 0.f;
        }
    }
}

oracle_learner *oracle_learner_create(const oracle_learner_desc *desc, const float *params_with_q) {
    oracle_learner *L = (oracle_learner *)calloc(1, sizeof(*L));
    L->d = *desc;
    L->P = oracle_param_offset(&desc->dims, OT_Q_W);
    L->Pq = oracle_param_offset(&desc->dims, OT_COUNT);
    L->n_batch = desc->n_envs * desc->n_steps;
    L->params = (float *)malloc(sizeof(float) * L->Pq);
    memcpy(L->params, params_with_q, sizeof(float) * L->Pq);
    L->m = (float *)calloc((size_t)L->P, sizeof(float));
    L->v = (float *)calloc((size_t)L->P, sizeof(float));
    L->b1p = desc->hp.beta1; /* beta powers start at beta (GRAPH:25426,25579) */
    L->b2p = desc->hp.beta2;
    L->norm = oracle_vecnorm_create(desc->n_envs, desc->dims.obs_dim, 1);
    L->norm->gamma = 0.99f; /* EnvNormalize default, independent of PPO2's gamma (env_normalize.hpp:27) */
    if (desc->env_kind == 0) L->env = oracle_synth_env_create(desc->n_envs, desc->dims.obs_dim, desc->seed ^ 0x1234u, 0);
    oracle_srand(&L->rng, desc->shuffle_seed);
    const int N = desc->n_envs, O = desc->dims.obs_dim, A = desc->dims.act_dim;
    L->cur_obs = (float *)malloc(sizeof(float) * (size_t)N * O);
    L->cur_dones = (float *)calloc((size_t)N, sizeof(float));
    const size_t nb = (size_t)L->n_batch;
    const size_t w[8] = {(size_t)O, 1, 1, (size_t)A, 1, 1, 1, 1};
    for (int i = 0; i < 8; ++i) L->buf[i] = (float *)calloc(nb * w[i], sizeof(float));
    L->threads = 1;
#ifdef _OPENMP
    L->threads = desc->threads > 0 ? desc->threads : omp_get_max_threads();
    omp_set_num_threads(L->threads);
#endif
    /* Runner ctor: obs = env.reset() through EnvNormalize::reset (runner.hpp:48) */
    float *raw = (float *)malloc(sizeof(float) * (size_t)N * O);
    env_reset(L, raw);
    oracle_vecnorm_reset_f32(L->norm, raw, L->cur_obs);
    free(raw);
    return L;
}

void oracle_learner_destroy(oracle_learner *L) {
    if (!L) return;
    for (int i = 0; i < 8; ++i) free(L->buf[i]);
    free(L->params); free(L->m); free(L->v); free(L->cur_obs); free(L->cur_dones);
    oracle_vecnorm_destroy(L->norm);
    oracle_synth_env_destroy(L->env);
    free(L);
}

/* transposeInPlace of a [T,N] buffer into [N,T] (runner.hpp:146-152) */
static void transpose_tn(float *a, int T, int N) {
    float *tmp = (float *)malloc(sizeof(float) * (size_t)T * N);
    for (int t = 0; t < T; ++t)
        for (int e = 0; e < N; ++e) tmp[(size_t)e * T + t] = a[(size_t)t * N + e];
    memcpy(a, tmp, sizeof(float) * (size_t)T * N);
    free(tmp);
}

void oracle_learner_rollout(oracle_learner *L) {
    const int N = L->d.n_envs, T = L->d.n_steps, O = L->d.dims.obs_dim, A = L->d.dims.act_dim;
    float *obs = L->buf[0], *returns = L->buf[1], *dones = L->buf[2], *actions = L->buf[3], *values = L->buf[4],
          *nlp = L->buf[5], *rews = L->buf[6], *urews = L->buf[7];
    float *eps = (float *)malloc(sizeof(float) * (size_t)N * A);
    float *act = (float *)malloc(sizeof(float) * (size_t)N * A);
    float *val = (float *)malloc(sizeof(float) * N), *nl = (float *)malloc(sizeof(float) * N);
    float *raw_obs = (float *)malloc(sizeof(float) * (size_t)N * O), *raw_rew = (float *)malloc(sizeof(float) * N);
    float *done = (float *)malloc(sizeof(float) * N), *nrew = (float *)malloc(sizeof(float) * N);
    for (int t = 0; t < T; ++t) {
        /* mb.obs block (env-major [env][t][O], runner.hpp:77-78) */
        for (int e = 0; e < N; ++e) memcpy(obs + ((size_t)e * T + t) * O, L->cur_obs + (size_t)e * O, sizeof(float) * O);
        for (int e = 0; e < N; ++e) oracle_normal_eps(L->d.seed, (uint32_t)e, L->global_step, A, eps + (size_t)e * A);
        oracle_policy_step_f32(&L->d.dims, L->params, L->cur_obs, N, eps, act, val, nl, NULL); /* model.step, :81 */
        for (int e = 0; e < N; ++e) memcpy(actions + ((size_t)e * T + t) * A, act + (size_t)e * A, sizeof(float) * A);
        memcpy(values + (size_t)t * N, val, sizeof(float) * N);     /* time-major rows until the final transpose */
        memcpy(nlp + (size_t)t * N, nl, sizeof(float) * N);
        memcpy(dones + (size_t)t * N, L->cur_dones, sizeof(float) * N); /* done of the PREVIOUS step (:110) */
        env_step(L, act, raw_obs, raw_rew, done);
        oracle_vecnorm_step_f32(L->norm, raw_obs, raw_rew, done, L->cur_obs, nrew); /* EnvNormalize::step */
        memcpy(L->cur_dones, done, sizeof(float) * N);
        memcpy(rews + (size_t)t * N, nrew, sizeof(float) * N);
        memcpy(urews + (size_t)t * N, raw_rew, sizeof(float) * N);
        L->global_step += 1;
    }
    /* set_returns (:159-191) */
    float *last_values = (float *)malloc(sizeof(float) * N);
    float *advs = (float *)malloc(sizeof(float) * (size_t)T * N);
    oracle_policy_step_f32(&L->d.dims, L->params, L->cur_obs, N, NULL, NULL, last_values, NULL, NULL);
    oracle_gae_f32(rews, values, dones, last_values, L->cur_dones, T, N, L->d.gamma, L->d.lam, advs, returns);
    /* flatten: 1-D buffers are transposed to [N,T] (:146-152) */
    transpose_tn(returns, T, N); transpose_tn(dones, T, N); transpose_tn(values, T, N); transpose_tn(nlp, T, N);
    transpose_tn(rews, T, N); transpose_tn(urews, T, N);
    free(eps); free(act); free(val); free(nl); free(raw_obs); free(raw_rew); free(done); free(nrew);
    free(last_values); free(advs);
}

void oracle_learner_train(oracle_learner *L, float *mean_losses) {
    const int nb = L->n_batch, O = L->d.dims.obs_dim, A = L->d.dims.act_dim;
    const int B = nb / L->d.nminibatches;
    const int widths[6] = {O, 1, 1, A, 1, 1};
    const int which[6] = {0, 1, 2, 3, 4, 5}; /* get_train_input(): obs, returns, dones, actions, values, neglogpacs */
    int *perm = (int *)malloc(sizeof(int) * nb);
    for (int i = 0; i < nb; ++i) perm[i] = i; /* setIdentity once per update (ppo2.hpp:274-275) */
    float *permuted[6], *slice[6];
    for (int k = 0; k < 6; ++k) {
        permuted[k] = (float *)malloc(sizeof(float) * (size_t)nb * widths[k]);
        slice[k] = (float *)malloc(sizeof(float) * (size_t)B * widths[k]);
    }
    float *advs = (float *)malloc(sizeof(float) * B), *grads = (float *)malloc(sizeof(float) * L->P);
    double acc[5] = {0, 0, 0, 0, 0};
    float accf[5] = {0, 0, 0, 0, 0};
    (void)acc;
    for (int epoch = 0; epoch < L->d.noptepochs; ++epoch) {
        oracle_random_shuffle(&L->rng, perm, nb); /* :288 */
        for (int k = 0; k < 6; ++k) {             /* tmp = perm * buf : out[perm[i]] = in[i]  (:291-296) */
            const float *src = L->buf[which[k]];
            const int w = widths[k];
#pragma omp parallel for schedule(static)
            for (int i = 0; i < nb; ++i) memcpy(permuted[k] + (size_t)perm[i] * w, src + (size_t)i * w, sizeof(float) * w);
        }
        for (int start = 0; start < nb; start += B) {
            for (int k = 0; k < 6; ++k) /* slices (:304-307) */
                memcpy(slice[k], permuted[k] + (size_t)start * widths[k], sizeof(float) * (size_t)B * widths[k]);
            float losses[5];
            oracle_advnorm_f32(slice[1], slice[4], B, advs); /* _train_step :401-406 */
            oracle_loss_grad_f32(&L->d.dims, &L->d.hp, L->params, slice[0], slice[3], advs, slice[1], slice[5], slice[4], B,
                                 L->d.cliprange, grads, losses);
            oracle_clip_adam_f32(&L->d.hp, L->P, L->d.lr, L->params, L->m, L->v, grads, &L->b1p, &L->b2p);
            for (int i = 0; i < 5; ++i) accf[i] += losses[i];
        }
    }
    const float rows = (float)(L->d.noptepochs * L->d.nminibatches);
    for (int i = 0; i < 5; ++i) mean_losses[i] = accf[i] / rows; /* colwise().mean() (:335) */
    for (int k = 0; k < 6; ++k) { free(permuted[k]); free(slice[k]); }
    free(perm); free(advs); free(grads);
}

const float *oracle_learner_buffer(oracle_learner *L, int which) { return (which >= 0 && which < 8) ? L->buf[which] : NULL; }
float *oracle_learner_params(oracle_learner *L) { return L->params; }
int oracle_learner_threads(oracle_learner *L) { return L->threads; }
void oracle_learner_get_norm(oracle_learner *L, float *obs_mean, float *obs_var, double *obs_count, float *ret_mean,
                             float *ret_var, double *ret_count) {
    memcpy(obs_mean, L->norm->obs_rms.mean, sizeof(float) * L->norm->obs_dim);
    memcpy(obs_var, L->norm->obs_rms.var, sizeof(float) * L->norm->obs_dim);
    *obs_count = L->norm->obs_rms.count;
    *ret_mean = L->norm->ret_rms.mean[0];
    *ret_var = L->norm->ret_rms.var[0];
    *ret_count = L->norm->ret_rms.count;
}
