/*
 * oracle/envs.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU fp64 restatement of the environment side of gymnasium-solver's rollout path:
 *   gym.make(CartPole-v1 | Acrobot-v1 | MountainCar-v0)            reference call site utils/environment.py:94-96
 *   + per-env wrappers applied through EnvWrapperRegistry           utils/environment.py:396, gym_wrappers/
 *   + TimeLimit                                                     utils/environment.py:401
 *   + SyncVectorEnv with NEXT_STEP autoreset                        utils/environment.py:410-415
 *   + vector RecordEpisodeStatistics                                utils/environment.py:136,212
 * The arithmetic lives in the third-party dependency gymnasium==1.1.1 (reference uv.lock:897-898), which
 * is NOT vendored under /root/reference and NOT installed in the build image.  This file restates the
 * published algorithm of gymnasium/envs/classic_control/{cartpole,acrobot,mountain_car}.py,
 * gymnasium/wrappers/common.py::TimeLimit, gymnasium/vector/sync_vector_env.py and
 * gymnasium/wrappers/vector/common.py::RecordEpisodeStatistics at tag v1.1.1 (SURVEY.md §8c spec).
 *
 * PARITY UNPINNED for the physics: the reference's test-suite holds no golden trajectory for these
 * environments and gymnasium cannot be executed here, so this restatement is checked only
 * against hand-derived known answers (tests/test_oracle_envs.py).  The wrappers
 * (gym_wrappers/MountainCarV0/state_count_bonus.py:51-126, gym_wrappers/CartPoleV1/reward_shaper.py:43-77,
 * gym_wrappers/MountainCarV0/reward_shaper.py:27-102) ARE reference code and are restated with the pinned
 * numpy 1.26.4 scalar-promotion rules (np.float32 scalar op python float -> float64).
 *
 * Reset noise: gymnasium draws from numpy PCG64(seed+i); that stream is not reproducible on a GPU, so the
 * engine defines its own counter-based stream, Philox4x32-10 (Salmon et al., SC'11) keyed by
 * (seed, global env id, reset count).  This file carries an INDEPENDENT implementation of the same
 * published generator, pinned by the Random123 known-answer vectors, so CPU and GPU trajectories can be
 * compared over many episodes without injecting states.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, IEEE fp64 like numpy)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_CARTPOLE 0
#define ORC_ACROBOT 1
#define ORC_MOUNTAINCAR 2

#define ORC_WRAP_COUNT_BONUS 1
#define ORC_WRAP_CARTPOLE_SHAPER 2
#define ORC_WRAP_MOUNTAINCAR_SHAPER 3

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

/* 53-bit uniform in [0,1) from two 32-bit words (same construction numpy uses for random()). */
static double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

#define ORC_TAG_RESET 0x5E5E0000u

/* uniform doubles for reset number `reset_idx` of global env `gid`: block b supplies draws 2b, 2b+1 */
static void reset_uniforms(uint64_t seed, uint64_t gid, uint32_t reset_idx, int n, double* u) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int b = 0; 2 * b < n; ++b) {
        uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), reset_idx, ORC_TAG_RESET | (uint32_t)b};
        uint32_t r[4];
        philox4x32_10(ctr, key, r);
        u[2 * b] = u53(r[0], r[1]);
        if (2 * b + 1 < n) u[2 * b + 1] = u53(r[2], r[3]);
    }
}

void orc_reset_uniforms(uint64_t seed, uint64_t gid, uint32_t reset_idx, int n, double* u) {
    reset_uniforms(seed, gid, reset_idx, n, u);
}

/* ------------------------------------------------------------------ handle */
typedef struct orc_env {
    int kind;
    int64_t n;
    int64_t gid0;
    uint64_t seed;
    int max_steps;
    int sdim, odim;
    double* state;         /* SoA [sdim][n] */
    int32_t* elapsed;      /* TimeLimit._elapsed_steps */
    uint8_t* autoreset;    /* SyncVectorEnv._autoreset_envs */
    uint8_t* prev_done;    /* RecordEpisodeStatistics.prev_dones */
    uint32_t* reset_count;
    double* ep_ret;
    int32_t* ep_len;
    int wrapper;
    double wp[8];
    int64_t* counts;       /* StateCountBonus tables [n][pb*vb] (int64 like the reference) */
} orc_env_t;

static int sdim_of(int kind) { return kind == ORC_MOUNTAINCAR ? 2 : 4; }
static int odim_of(int kind) { return kind == ORC_CARTPOLE ? 4 : (kind == ORC_ACROBOT ? 6 : 2); }

orc_env_t* orc_env_create(int kind, int64_t n, int64_t gid0, uint64_t seed, int max_steps) {
    if (kind < 0 || kind > 2 || n <= 0) return NULL;
    orc_env_t* e = (orc_env_t*)calloc(1, sizeof(orc_env_t));
    e->kind = kind; e->n = n; e->gid0 = gid0; e->seed = seed;
    e->max_steps = max_steps > 0 ? max_steps : (kind == ORC_MOUNTAINCAR ? 200 : 500);
    e->sdim = sdim_of(kind); e->odim = odim_of(kind);
    e->state = (double*)calloc((size_t)e->sdim * n, sizeof(double));
    e->elapsed = (int32_t*)calloc(n, sizeof(int32_t));
    e->autoreset = (uint8_t*)calloc(n, 1);
    e->prev_done = (uint8_t*)calloc(n, 1);
    e->reset_count = (uint32_t*)calloc(n, sizeof(uint32_t));
    e->ep_ret = (double*)calloc(n, sizeof(double));
    e->ep_len = (int32_t*)calloc(n, sizeof(int32_t));
    return e;
}

void orc_env_destroy(orc_env_t* e) {
    if (!e) return;
    free(e->state); free(e->elapsed); free(e->autoreset); free(e->prev_done);
    free(e->reset_count); free(e->ep_ret); free(e->ep_len); free(e->counts); free(e);
}

int orc_env_attach_wrapper(orc_env_t* e, int wrapper, const double* params, int n_params) {
    if (e->wrapper) return -1;
    if (n_params > 8) return -1;
    if (wrapper == ORC_WRAP_COUNT_BONUS) {
        if (e->kind != ORC_MOUNTAINCAR || n_params != 5) return -1;
    } else if (wrapper == ORC_WRAP_CARTPOLE_SHAPER) {
        if (e->kind != ORC_CARTPOLE || n_params != 3) return -1;
    } else if (wrapper == ORC_WRAP_MOUNTAINCAR_SHAPER) {
        if (e->kind != ORC_MOUNTAINCAR || n_params != 3) return -1;
    } else return -1;
    e->wrapper = wrapper;
    memcpy(e->wp, params, sizeof(double) * n_params);
    if (wrapper == ORC_WRAP_COUNT_BONUS) {
        size_t cells = (size_t)e->wp[0] * (size_t)e->wp[1];
        e->counts = (int64_t*)calloc(cells * e->n, sizeof(int64_t));
    }
    return 0;
}

void orc_env_set_state(orc_env_t* e, const double* state, const int32_t* elapsed) {
    memcpy(e->state, state, sizeof(double) * e->sdim * e->n);
    if (elapsed) memcpy(e->elapsed, elapsed, sizeof(int32_t) * e->n);
}
void orc_env_get_state(const orc_env_t* e, double* state, int32_t* elapsed) {
    memcpy(state, e->state, sizeof(double) * e->sdim * e->n);
    if (elapsed) memcpy(elapsed, e->elapsed, sizeof(int32_t) * e->n);
}
int64_t* orc_env_counts(orc_env_t* e) { return e->counts; }

/* ------------------------------------------------------------------ observations */
static void write_obs(const orc_env_t* e, int64_t i, float* obs) {
    const double* s = e->state;
    int64_t n = e->n;
    float* o = obs + i * e->odim;
    if (e->kind == ORC_CARTPOLE) {
        for (int k = 0; k < 4; ++k) o[k] = (float)s[k * n + i];
    } else if (e->kind == ORC_MOUNTAINCAR) {
        o[0] = (float)s[i]; o[1] = (float)s[n + i];
    } else { /* acrobot _get_ob */
        double t1 = s[i], t2 = s[n + i];
        o[0] = (float)cos(t1); o[1] = (float)sin(t1);
        o[2] = (float)cos(t2); o[3] = (float)sin(t2);
        o[4] = (float)s[2 * n + i]; o[5] = (float)s[3 * n + i];
    }
}

/* env.reset() of one sub-env: new state from its Philox stream; TimeLimit.reset zeroes elapsed */
static void reset_one(orc_env_t* e, int64_t i) {
    double u[4];
    int64_t n = e->n;
    uint64_t gid = (uint64_t)(e->gid0 + i);
    uint32_t ridx = e->reset_count[i]++;
    if (e->kind == ORC_CARTPOLE) {
        reset_uniforms(e->seed, gid, ridx, 4, u);
        for (int k = 0; k < 4; ++k) e->state[k * n + i] = -0.05 + (0.05 - (-0.05)) * u[k];
    } else if (e->kind == ORC_ACROBOT) {
        reset_uniforms(e->seed, gid, ridx, 4, u);
        /* uniform(-0.1, 0.1, size=4).astype(np.float32): state holds float32-rounded values */
        for (int k = 0; k < 4; ++k) e->state[k * n + i] = (double)(float)(-0.1 + (0.1 - (-0.1)) * u[k]);
    } else {
        reset_uniforms(e->seed, gid, ridx, 1, u);
        e->state[i] = -0.6 + (-0.4 - (-0.6)) * u[0];
        e->state[n + i] = 0.0;
    }
    e->elapsed[i] = 0;
}

void orc_env_reset(orc_env_t* e, float* obs) {
    for (int64_t i = 0; i < e->n; ++i) {
        reset_one(e, i);
        e->autoreset[i] = 0;
        /* RecordEpisodeStatistics.reset */
        e->prev_done[i] = 0; e->ep_ret[i] = 0.0; e->ep_len[i] = 0;
        write_obs(e, i, obs);
    }
}

/* ------------------------------------------------------------------ dynamics */
static int cartpole_step(double* x, double* x_dot, double* theta, double* theta_dot, int action) {
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1;
    const double total_mass = masspole + masscart;
    const double length = 0.5;
    const double polemass_length = masspole * length;
    const double force_mag = 10.0, tau = 0.02;
    const double theta_threshold = 12 * 2 * M_PI / 360;
    const double x_threshold = 2.4;
    double force = action == 1 ? force_mag : -force_mag;
    double costheta = cos(*theta), sintheta = sin(*theta);
    double temp = (force + polemass_length * ((*theta_dot) * (*theta_dot)) * sintheta) / total_mass;
    double thetaacc = (gravity * sintheta - costheta * temp) /
                      (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
    double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
    *x = *x + tau * (*x_dot);
    *x_dot = *x_dot + tau * xacc;
    *theta = *theta + tau * (*theta_dot);
    *theta_dot = *theta_dot + tau * thetaacc;
    return (*x < -x_threshold) || (*x > x_threshold) || (*theta < -theta_threshold) || (*theta > theta_threshold);
}

static int mountaincar_step(double* position, double* velocity, int action) {
    const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5;
    const double goal_velocity = 0.0, force = 0.001, gravity = 0.0025;
    double p = *position, v = *velocity;
    v += (double)(action - 1) * force + cos(3 * p) * (-gravity);
    v = v < -max_speed ? -max_speed : (v > max_speed ? max_speed : v);
    p += v;
    p = p < min_position ? min_position : (p > max_position ? max_position : p);
    if (p == min_position && v < 0) v = 0;
    *position = p; *velocity = v;
    return (p >= goal_position) && (v >= goal_velocity);
}

static void acrobot_dsdt(const double s[5], double d[5]) {
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0, g = 9.8;
    double a = s[4];
    double theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
    double d1 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2;
    double d2 = m2 * (lc2 * lc2 + l1 * lc2 * cos(theta2)) + I2;
    double phi2 = m2 * lc2 * g * cos(theta1 + theta2 - M_PI / 2.0);
    double phi1 = -m2 * l1 * lc2 * (dtheta2 * dtheta2) * sin(theta2)
                  - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2)
                  + (m1 * lc1 + m2 * l1) * g * cos(theta1 - M_PI / 2)
                  + phi2;
    /* "book" variant */
    double ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * (dtheta1 * dtheta1) * sin(theta2) - phi2) /
                      (m2 * (lc2 * lc2) + I2 - (d2 * d2) / d1);
    double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    d[0] = dtheta1; d[1] = dtheta2; d[2] = ddtheta1; d[3] = ddtheta2; d[4] = 0.0;
}

static double wrap_angle(double x, double m, double M) {
    double diff = M - m;
    while (x > M) x = x - diff;
    while (x < m) x = x + diff;
    return x;
}
static double bound(double x, double m, double M) { double y = x > m ? x : m; return y < M ? y : M; }

static int acrobot_step(double s4[4], int action, double* reward) {
    const double dt = 0.2, MAX_VEL_1 = 4 * M_PI, MAX_VEL_2 = 9 * M_PI;
    const double torque_tab[3] = {-1.0, 0.0, +1.0};
    double y0[5] = {s4[0], s4[1], s4[2], s4[3], torque_tab[action]};
    double k1[5], k2[5], k3[5], k4[5], y[5];
    double dtt = dt - 0, dt2 = dtt / 2.0;
    acrobot_dsdt(y0, k1);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt2 * k1[i];
    acrobot_dsdt(y, k2);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dt2 * k2[i];
    acrobot_dsdt(y, k3);
    for (int i = 0; i < 5; ++i) y[i] = y0[i] + dtt * k3[i];
    acrobot_dsdt(y, k4);
    double ns[4];
    for (int i = 0; i < 4; ++i) ns[i] = y0[i] + dtt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    ns[0] = wrap_angle(ns[0], -M_PI, M_PI);
    ns[1] = wrap_angle(ns[1], -M_PI, M_PI);
    ns[2] = bound(ns[2], -MAX_VEL_1, MAX_VEL_1);
    ns[3] = bound(ns[3], -MAX_VEL_2, MAX_VEL_2);
    for (int i = 0; i < 4; ++i) s4[i] = ns[i];
    int terminated = (-cos(ns[0]) - cos(ns[1] + ns[0])) > 1.0;
    *reward = terminated ? 0.0 : -1.0;
    return terminated;
}

/* ------------------------------------------------------------------ wrappers (reference code restated) */
static double clip01(double v, double hi) { return v < 0.0 ? 0.0 : (v > hi ? hi : v); }

/* gym_wrappers/CartPoleV1/reward_shaper.py:43-57; obs entries are float32 values promoted by float() */
static double cartpole_phi(const orc_env_t* e, double x_f32, double theta_f32) {
    const double x_threshold = 2.4, theta_threshold = 12 * 2 * M_PI / 360;
    double pos_term = 1.0 - fabs(x_f32) / (x_threshold > 1e-6 ? x_threshold : 1e-6);
    double angle_term = 1.0 - fabs(theta_f32) / (theta_threshold > 1e-6 ? theta_threshold : 1e-6);
    if (e->wp[2] != 0.0) { pos_term = clip01(pos_term, 1.0); angle_term = clip01(angle_term, 1.0); }
    return e->wp[0] * angle_term + e->wp[1] * pos_term;
}

/* gym_wrappers/MountainCarV0/reward_shaper.py:27-48,60-102 */
static double mc_shaping(const orc_env_t* e, double p0, double v0, double p1, double v1) {
    const double min_position = -1.2, goal_position = 0.5, min_velocity = -0.07, max_velocity = 0.07;
    double h0 = sin(3 * p0), h1 = sin(3 * p1);
    double cp = (p1 - min_position) / (goal_position - min_position);
    double pp = (p0 - min_position) / (goal_position - min_position);
    double position_shaping = e->wp[0] * (cp - pp);
    double cv = (v1 - min_velocity) / (max_velocity - min_velocity);
    double pv = (v0 - min_velocity) / (max_velocity - min_velocity);
    double velocity_shaping = e->wp[1] * (cv - pv);
    double ch = (h1 + 1) / 2, ph = (h0 + 1) / 2;
    double height_shaping = e->wp[2] * (ch - ph);
    return position_shaping + velocity_shaping + height_shaping;
}

/* gym_wrappers/MountainCarV0/state_count_bonus.py:51-86,96-126 */
static double count_bonus(orc_env_t* e, int64_t i, double p_f32, double v_f32) {
    const double min_position = -1.2, max_position = 0.6, min_velocity = -0.07, max_velocity = 0.07;
    int pb = (int)e->wp[0], vb = (int)e->wp[1];
    double scale = e->wp[2];
    int btype = (int)e->wp[3];
    int64_t min_count = (int64_t)e->wp[4];
    double pos_norm = clip01((p_f32 - min_position) / (max_position - min_position), 0.999999);
    double vel_norm = clip01((v_f32 - min_velocity) / (max_velocity - min_velocity), 0.999999);
    int pos_bin = (int)(pos_norm * pb), vel_bin = (int)(vel_norm * vb);
    int64_t* cell = e->counts + ((size_t)i * pb + pos_bin) * vb + vel_bin;
    int64_t count = *cell;
    int64_t eff = count > min_count ? count : min_count;
    double bonus;
    if (btype == 0) bonus = 1.0 / sqrt((double)eff);
    else if (btype == 1) bonus = 1.0 / (double)eff;
    else bonus = 1.0 / log((double)(eff + 1));
    *cell = count + 1;
    return scale * bonus;
}

/* ------------------------------------------------------------------ SyncVectorEnv.step */
/* steps sub-envs [lo, hi): envs are independent, so host threads may call disjoint ranges concurrently */
void orc_env_step_range(orc_env_t* e, int64_t lo, int64_t hi, const int32_t* actions, float* obs, double* reward,
                        uint8_t* terminated, uint8_t* truncated, double* ep_return, int32_t* ep_length) {
    int64_t n = e->n;
    for (int64_t i = lo; i < hi; ++i) {
        double r = 0.0;
        int term = 0, trunc = 0;
        if (e->autoreset[i]) {
            /* NEXT_STEP autoreset: reset instead of step, reward 0, flags False, action ignored */
            reset_one(e, i);
        } else {
            double* s = e->state;
            if (e->kind == ORC_CARTPOLE) {
                double x = s[i], xd = s[n + i], th = s[2 * n + i], thd = s[3 * n + i];
                double phi0 = 0.0;
                if (e->wrapper == ORC_WRAP_CARTPOLE_SHAPER) phi0 = cartpole_phi(e, (double)(float)x, (double)(float)th);
                term = cartpole_step(&x, &xd, &th, &thd, actions[i]);
                s[i] = x; s[n + i] = xd; s[2 * n + i] = th; s[3 * n + i] = thd;
                r = 1.0;
                if (e->wrapper == ORC_WRAP_CARTPOLE_SHAPER)
                    r = r + (cartpole_phi(e, (double)(float)x, (double)(float)th) - phi0);
            } else if (e->kind == ORC_MOUNTAINCAR) {
                double p = s[i], v = s[n + i];
                double p0 = (double)(float)p, v0 = (double)(float)v;
                term = mountaincar_step(&p, &v, actions[i]);
                s[i] = p; s[n + i] = v;
                r = -1.0;
                if (e->wrapper == ORC_WRAP_COUNT_BONUS) r = r + count_bonus(e, i, (double)(float)p, (double)(float)v);
                else if (e->wrapper == ORC_WRAP_MOUNTAINCAR_SHAPER)
                    r = r + mc_shaping(e, p0, v0, (double)(float)p, (double)(float)v);
            } else {
                double s4[4] = {s[i], s[n + i], s[2 * n + i], s[3 * n + i]};
                term = acrobot_step(s4, actions[i], &r);
                s[i] = s4[0]; s[n + i] = s4[1]; s[2 * n + i] = s4[2]; s[3 * n + i] = s4[3];
            }
            /* TimeLimit.step */
            e->elapsed[i] += 1;
            if (e->elapsed[i] >= e->max_steps) trunc = 1;
        }
        /* RecordEpisodeStatistics.step */
        if (e->prev_done[i]) { e->ep_ret[i] = 0.0; e->ep_len[i] = 0; }
        else { e->ep_ret[i] += r; e->ep_len[i] += 1; }
        int done = term || trunc;
        e->prev_done[i] = (uint8_t)done;
        e->autoreset[i] = (uint8_t)done;
        write_obs(e, i, obs);
        reward[i] = r;
        terminated[i] = (uint8_t)term;
        truncated[i] = (uint8_t)trunc;
        if (ep_return) ep_return[i] = done ? e->ep_ret[i] : 0.0;
        if (ep_length) ep_length[i] = done ? e->ep_len[i] : 0;
    }
}

void orc_env_step(orc_env_t* e, const int32_t* actions, float* obs, double* reward, uint8_t* terminated,
                  uint8_t* truncated, double* ep_return, int32_t* ep_length) {
    orc_env_step_range(e, 0, e->n, actions, obs, reward, terminated, truncated, ep_return, ep_length);
}
