// env_dynamics.cuh — per-env device functions: classic-control physics (fp64, IEEE op order of Gymnasium's
// Python), TimeLimit, NEXT_STEP autoreset, RecordEpisodeStatistics and the reference's reward wrappers.
//
// Replaces (one thread per env, state in registers):
//   gym.make(env_id).step/reset           gymnasium 1.1.1 classic_control (external; call site utils/environment.py:94-96)
//   TimeLimit                             utils/environment.py:401
//   SyncVectorEnv.step (NEXT_STEP)        utils/environment.py:410-415
//   RecordEpisodeStatistics               utils/environment.py:136,212
//   MountainCarV0_StateCountBonus.step    gym_wrappers/MountainCarV0/state_count_bonus.py:96-126
//   CartPoleV1_RewardShaper.step          gym_wrappers/CartPoleV1/reward_shaper.py:43-77
//   MountainCarV0_RewardShaper.step       gym_wrappers/MountainCarV0/reward_shaper.py:60-102
// All fp64 products/sums go through __dmul_rn/__dadd_rn so ptxas cannot fuse them: the sequence of roundings is
// the one numpy / CPython performs, which is what keeps termination flags and episode lengths bit-exact.
#pragma once

#include "common.cuh"

namespace gs {

constexpr double kPi = 3.141592653589793;

struct EnvParams {
    int kind;
    int max_steps;
    int wrapper;
    double wp[5];
    uint32_t* counts;  // StateCountBonus tables [N][pb*vb]
    uint64_t seed;
    int64_t gid0;
    // VecNormalizeStatic (gym_wrappers/vec_normalize_static.py:44-60): per observation dim 0 = pass through (non-finite bounds),
    // 1 = (x - low) / den with den = (high - low) + 1e-8 in fp32, 2 = constant 0 (degenerate bounds)
    int obs_norm;
    int on_mode[6];
    float on_low[6], on_den[6];
};

// registers of one env between vector steps
struct EnvRegs {
    double s[4];
    double ep_ret;
    int32_t elapsed;
    int32_t ep_len;
    uint32_t reset_count;
    uint32_t needs_reset;  // SyncVectorEnv._autoreset_envs == RecordEpisodeStatistics.prev_dones
};

template <int KIND> struct EnvDims;
template <> struct EnvDims<GS_ENV_CARTPOLE_V1> { static constexpr int S = 4, D = 4, A = 2; };
template <> struct EnvDims<GS_ENV_ACROBOT_V1> { static constexpr int S = 4, D = 6, A = 3; };
template <> struct EnvDims<GS_ENV_MOUNTAINCAR_V0> { static constexpr int S = 2, D = 2, A = 3; };

// ---- reset noise: uniform(low, high) from Philox(seed; gid, reset_idx, block) ---------------------------
template <int KIND>
__device__ __forceinline__ void env_reset_state(double s[4], uint64_t seed, uint64_t gid, uint32_t ridx) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 r0 = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), ridx, kTagReset | 0u), key);
    if (KIND == GS_ENV_MOUNTAINCAR_V0) {
        s[0] = dadd(-0.6, dmul(dsub(-0.4, -0.6), u53(r0.x, r0.y)));
        s[1] = 0.0;
        s[2] = 0.0; s[3] = 0.0;
        return;
    }
    const uint4 r1 = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), ridx, kTagReset | 1u), key);
    const double u[4] = {u53(r0.x, r0.y), u53(r0.z, r0.w), u53(r1.x, r1.y), u53(r1.z, r1.w)};
    if (KIND == GS_ENV_CARTPOLE_V1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = dadd(-0.05, dmul(dsub(0.05, -0.05), u[k]));
    } else {  // Acrobot: uniform(-0.1, 0.1).astype(np.float32)
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = (double)(float)dadd(-0.1, dmul(dsub(0.1, -0.1), u[k]));
    }
}

// ---- observations -----------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void env_obs(const double s[4], float* o) {
    if (KIND == GS_ENV_CARTPOLE_V1) {
        o[0] = (float)s[0]; o[1] = (float)s[1]; o[2] = (float)s[2]; o[3] = (float)s[3];
    } else if (KIND == GS_ENV_MOUNTAINCAR_V0) {
        o[0] = (float)s[0]; o[1] = (float)s[1];
    } else {
        double s1, c1, s2, c2;
        sincos(s[0], &s1, &c1);
        sincos(s[1], &s2, &c2);
        o[0] = (float)c1; o[1] = (float)s1; o[2] = (float)c2; o[3] = (float)s2;
        o[4] = (float)s[2]; o[5] = (float)s[3];
    }
}

// fused VecNormalizeStatic: the observation every consumer sees (policy, rollout buffer, current-obs array) is the normalised one;
// the per-env reward wrappers read the raw state, as in the reference where they sit below the vector wrapper
template <int D>
__device__ __forceinline__ void obs_normalize(const EnvParams& P, float* o) {
    if (!P.obs_norm) return;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (P.on_mode[d] == 1) o[d] = __fdiv_rn(__fsub_rn(o[d], P.on_low[d]), P.on_den[d]);
        else if (P.on_mode[d] == 2) o[d] = 0.0f;
    }
}

// ---- CartPole-v1 (euler) ------------------------------------------------------------------------------
__device__ __forceinline__ bool cartpole_dynamics(double s[4], int action) {
    const double gravity = 9.8, masspole = 0.1, total_mass = 0.1 + 1.0, length = 0.5;
    const double polemass_length = 0.1 * 0.5, force_mag = 10.0, tau = 0.02;
    const double theta_threshold = 12 * 2 * kPi / 360, x_threshold = 2.4;
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    const double force = action == 1 ? force_mag : -force_mag;
    double sintheta, costheta;
    sincos(theta, &sintheta, &costheta);
    const double temp = ddiv(dadd(force, dmul(dmul(polemass_length, dmul(theta_dot, theta_dot)), sintheta)), total_mass);
    const double thetaacc = ddiv(dsub(dmul(gravity, sintheta), dmul(costheta, temp)),
                                 dmul(length, dsub(4.0 / 3.0, ddiv(dmul(masspole, dmul(costheta, costheta)), total_mass))));
    const double xacc = dsub(temp, ddiv(dmul(dmul(polemass_length, thetaacc), costheta), total_mass));
    x = dadd(x, dmul(tau, x_dot));
    x_dot = dadd(x_dot, dmul(tau, xacc));
    theta = dadd(theta, dmul(tau, theta_dot));
    theta_dot = dadd(theta_dot, dmul(tau, thetaacc));
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;
    return (x < -x_threshold) || (x > x_threshold) || (theta < -theta_threshold) || (theta > theta_threshold);
}

// ---- MountainCar-v0 ---------------------------------------------------------------------------------
__device__ __forceinline__ bool mountaincar_dynamics(double s[4], int action) {
    const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5;
    const double goal_velocity = 0.0, force = 0.001, gravity = 0.0025;
    double p = s[0], v = s[1];
    v = dadd(v, dadd(dmul((double)(action - 1), force), dmul(cos(dmul(3.0, p)), -gravity)));
    v = v < -max_speed ? -max_speed : (v > max_speed ? max_speed : v);
    p = dadd(p, v);
    p = p < min_position ? min_position : (p > max_position ? max_position : p);
    if (p == min_position && v < 0) v = 0;
    s[0] = p; s[1] = v;
    return (p >= goal_position) && (v >= goal_velocity);
}

// ---- Acrobot-v1 ("book" dynamics, one RK4 step over [0, dt]) -----------------------------------------------
__device__ __forceinline__ void acrobot_dsdt(const double s[4], double a, double d[4]) {
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0, g = 9.8;
    const double theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
    double sin2, cos2;
    sincos(theta2, &sin2, &cos2);
    // d1 = m1*lc1**2 + m2*(l1**2 + lc2**2 + 2*l1*lc2*cos(theta2)) + I1 + I2
    const double d1 = dadd(dadd(dadd(dmul(m1, lc1 * lc1), dmul(m2, dadd(dadd(l1 * l1, lc2 * lc2), dmul(2 * l1 * lc2, cos2)))), I1), I2);
    // d2 = m2*(lc2**2 + l1*lc2*cos(theta2)) + I2
    const double d2 = dadd(dmul(m2, dadd(lc2 * lc2, dmul(l1 * lc2, cos2))), I2);
    // phi2 = m2*lc2*g*cos(theta1 + theta2 - pi/2)
    const double phi2 = dmul(m2 * lc2 * g, cos(dsub(dadd(theta1, theta2), kPi / 2.0)));
    // phi1 = -m2*l1*lc2*dtheta2**2*sin(theta2) - 2*m2*l1*lc2*dtheta2*dtheta1*sin(theta2) + (m1*lc1+m2*l1)*g*cos(theta1-pi/2) + phi2
    const double t_a = dmul(dmul(-m2 * l1 * lc2, dmul(dtheta2, dtheta2)), sin2);
    const double t_b = dmul(dmul(dmul(2 * m2 * l1 * lc2, dtheta2), dtheta1), sin2);
    const double t_c = dmul((m1 * lc1 + m2 * l1) * g, cos(dsub(theta1, kPi / 2)));
    const double phi1 = dadd(dadd(dsub(t_a, t_b), t_c), phi2);
    // ddtheta2 = (a + d2/d1*phi1 - m2*l1*lc2*dtheta1**2*sin(theta2) - phi2) / (m2*lc2**2 + I2 - d2**2/d1)
    const double num = dsub(dsub(dadd(a, dmul(ddiv(d2, d1), phi1)), dmul(dmul(m2 * l1 * lc2, dmul(dtheta1, dtheta1)), sin2)), phi2);
    const double den = dsub(dadd(m2 * (lc2 * lc2), I2), ddiv(dmul(d2, d2), d1));
    const double ddtheta2 = ddiv(num, den);
    const double ddtheta1 = ddiv(-dadd(dmul(d2, ddtheta2), phi1), d1);
    d[0] = dtheta1; d[1] = dtheta2; d[2] = ddtheta1; d[3] = ddtheta2;
}

__device__ __forceinline__ bool acrobot_dynamics(double s[4], int action, double& reward) {
    const double dt = 0.2, dt2 = 0.2 / 2.0, dt6 = 0.2 / 6.0;
    const double MAX_VEL_1 = 4 * kPi, MAX_VEL_2 = 9 * kPi;
    const double a = (double)(action - 1);  // AVAIL_TORQUE = [-1, 0, +1]
    double k1[4], k2[4], k3[4], k4[4], y[4];
    acrobot_dsdt(s, a, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = dadd(s[i], dmul(dt2, k1[i]));
    acrobot_dsdt(y, a, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = dadd(s[i], dmul(dt2, k2[i]));
    acrobot_dsdt(y, a, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = dadd(s[i], dmul(dt, k3[i]));
    acrobot_dsdt(y, a, k4);
    double ns[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        ns[i] = dadd(s[i], dmul(dt6, dadd(dadd(dadd(k1[i], dmul(2.0, k2[i])), dmul(2.0, k3[i])), k4[i])));
    const double diff = dsub(kPi, -kPi);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double x = ns[i];
        while (x > kPi) x = dsub(x, diff);
        while (x < -kPi) x = dadd(x, diff);
        ns[i] = x;
    }
    ns[2] = fmin(fmax(ns[2], -MAX_VEL_1), MAX_VEL_1);
    ns[3] = fmin(fmax(ns[3], -MAX_VEL_2), MAX_VEL_2);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = ns[i];
    const bool terminated = dsub(-cos(ns[0]), cos(dadd(ns[1], ns[0]))) > 1.0;
    reward = terminated ? 0.0 : -1.0;
    return terminated;
}

// ---- wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double clip0(double v, double hi) { return v < 0.0 ? 0.0 : (v > hi ? hi : v); }

__device__ __forceinline__ double cartpole_phi(const EnvParams& P, double x, double theta) {
    const double x_threshold = 2.4, theta_threshold = 12 * 2 * kPi / 360;
    double pos_term = dsub(1.0, ddiv(fabs(x), x_threshold));
    double angle_term = dsub(1.0, ddiv(fabs(theta), theta_threshold));
    if (P.wp[2] != 0.0) { pos_term = clip0(pos_term, 1.0); angle_term = clip0(angle_term, 1.0); }
    return dadd(dmul(P.wp[0], angle_term), dmul(P.wp[1], pos_term));
}

__device__ __forceinline__ double mountaincar_shaping(const EnvParams& P, double p0, double v0, double p1, double v1) {
    const double min_position = -1.2, goal_position = 0.5, min_velocity = -0.07, max_velocity = 0.07;
    const double h0 = sin(dmul(3.0, p0)), h1 = sin(dmul(3.0, p1));
    const double pr = dsub(goal_position, min_position), vr = dsub(max_velocity, min_velocity);
    const double ps = dmul(P.wp[0], dsub(ddiv(dsub(p1, min_position), pr), ddiv(dsub(p0, min_position), pr)));
    const double vs = dmul(P.wp[1], dsub(ddiv(dsub(v1, min_velocity), vr), ddiv(dsub(v0, min_velocity), vr)));
    const double hs = dmul(P.wp[2], dsub(ddiv(dadd(h1, 1.0), 2.0), ddiv(dadd(h0, 1.0), 2.0)));
    return dadd(dadd(ps, vs), hs);
}

__device__ __forceinline__ double count_bonus(const EnvParams& P, int64_t local, double p, double v) {
    const double min_position = -1.2, max_position = 0.6, min_velocity = -0.07, max_velocity = 0.07;
    const int pb = (int)P.wp[0], vb = (int)P.wp[1];
    const double pos_norm = clip0(ddiv(dsub(p, min_position), dsub(max_position, min_position)), 0.999999);
    const double vel_norm = clip0(ddiv(dsub(v, min_velocity), dsub(max_velocity, min_velocity)), 0.999999);
    const int pos_bin = (int)dmul(pos_norm, (double)pb), vel_bin = (int)dmul(vel_norm, (double)vb);
    uint32_t* cell = P.counts + ((size_t)local * pb + pos_bin) * vb + vel_bin;
    const uint32_t count = *cell;
    const uint32_t min_count = (uint32_t)P.wp[4];
    const double eff = (double)(count > min_count ? count : min_count);
    const int btype = (int)P.wp[3];
    double bonus;
    if (btype == 0) bonus = ddiv(1.0, sqrt(eff));
    else if (btype == 1) bonus = ddiv(1.0, eff);
    else bonus = ddiv(1.0, log(dadd(eff, 1.0)));
    if (count != 0xFFFFFFFFu) *cell = count + 1;  // saturating
    return dmul(P.wp[2], bonus);
}

// ---- one sub-env of SyncVectorEnv.step + TimeLimit + RecordEpisodeStatistics ------------------------------------
// Outputs: obs (D floats), reward (fp64, stored as fp32 by the caller like rewards_buf), terminated/truncated,
// ep_r/ep_l = infos["episode"]["r"|"l"] where done else 0.
template <int KIND>
__device__ __forceinline__ void env_vec_step(EnvRegs& e, const EnvParams& P, int64_t local, int action, float* obs,
                                             double& reward, bool& terminated, bool& truncated, double& ep_r, int& ep_l) {
    double r = 0.0;
    bool term = false, trunc = false;
    if (KIND == GS_ENV_MOUNTAINCAR_V0 && P.wrapper == GS_WRAP_SCRIPTED_REPLAY) {
        // table replay (see gs_engine.h): elapsed is the step index, the state holds the observation being shown
        const float* tab = reinterpret_cast<const float*>(P.counts);
        const int L = (int)P.wp[0];
        if (e.needs_reset) { e.ep_ret = 0.0; e.ep_len = 0; }             // RecordEpisodeStatistics: the step after a done starts a new episode
        const int k = e.elapsed < L - 1 ? e.elapsed : L - 1;
        r = (double)tab[k];
        term = tab[L + k] != 0.f;
        trunc = tab[2 * L + k] != 0.f;
        const int ko = e.elapsed + 1 < L ? e.elapsed + 1 : L;
        e.s[0] = (double)tab[3 * L + 2 * ko];
        e.s[1] = (double)tab[3 * L + 2 * ko + 1];
        e.elapsed += 1;
        e.ep_ret = dadd(e.ep_ret, r);
        e.ep_len += 1;
    } else if (e.needs_reset) {
        env_reset_state<KIND>(e.s, P.seed, (uint64_t)(P.gid0 + local), e.reset_count);
        e.reset_count += 1;
        e.elapsed = 0;
        e.ep_ret = 0.0;
        e.ep_len = 0;
    } else {
        if (KIND == GS_ENV_CARTPOLE_V1) {
            double phi0 = 0.0;
            if (P.wrapper == GS_WRAP_CARTPOLE_REWARD_SHAPER) phi0 = cartpole_phi(P, (double)(float)e.s[0], (double)(float)e.s[2]);
            term = cartpole_dynamics(e.s, action);
            r = 1.0;
            if (P.wrapper == GS_WRAP_CARTPOLE_REWARD_SHAPER)
                r = dadd(r, dsub(cartpole_phi(P, (double)(float)e.s[0], (double)(float)e.s[2]), phi0));
        } else if (KIND == GS_ENV_MOUNTAINCAR_V0) {
            const double p0 = (double)(float)e.s[0], v0 = (double)(float)e.s[1];
            term = mountaincar_dynamics(e.s, action);
            r = -1.0;
            if (P.wrapper == GS_WRAP_MOUNTAINCAR_STATE_COUNT_BONUS)
                r = dadd(r, count_bonus(P, local, (double)(float)e.s[0], (double)(float)e.s[1]));
            else if (P.wrapper == GS_WRAP_MOUNTAINCAR_REWARD_SHAPER)
                r = dadd(r, mountaincar_shaping(P, p0, v0, (double)(float)e.s[0], (double)(float)e.s[1]));
        } else {
            term = acrobot_dynamics(e.s, action, r);
        }
        e.elapsed += 1;
        if (e.elapsed >= P.max_steps) trunc = true;
        e.ep_ret = dadd(e.ep_ret, r);
        e.ep_len += 1;
    }
    const bool done = term || trunc;
    e.needs_reset = done ? 1u : 0u;
    env_obs<KIND>(e.s, obs);
    obs_normalize<EnvDims<KIND>::D>(P, obs);
    reward = r;
    terminated = term;
    truncated = trunc;
    ep_r = done ? e.ep_ret : 0.0;
    ep_l = done ? e.ep_len : 0;
}

}  // namespace gs
