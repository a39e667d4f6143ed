// shipenv_kernels.cuh -- sm_100a device code of the batched ship-in-transit environment.
//
// Execution model: one ship asset per thread, lane pairs (2e, 2e+1) = (ship under test, obstacle
// ship) of environment e.  A lane keeps its ship's whole state in registers (6 hull states, shaft
// speed, clock, LOS and controller integrators, the current route segment) for as many simulator
// steps as the call needs; the two ships of an environment only meet in the reward/termination
// evaluation, which exchanges a handful of values with __shfl_xor_sync(.., 1).  Ship-type constants,
// the fixed routes and the map polygons are staged once per CTA in shared memory; per-ship and
// per-env state is structure-of-arrays in HBM (include/shipenv.h) so every load/store of a warp is
// one contiguous 256-byte (FP64) or 128-byte (int32) segment.
//
// The run_colav IW env without collision avoidance additionally skips the per-step event tests while every lane of
// the warp is provably far from all of its thresholds ("quiet steps", k_env): same results bit for bit, checked
// against the twin instantiation that evaluates everything at every step.
//
// Arithmetic is IEEE FP64 with the reference's forward-Euler scheme and operation order; the file is
// compiled with -fmad=false so a*b+c is two roundings, as in CPython/NumPy.
// Citations are paths relative to the reference root.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "shipenv.h"
#include "shipenv_launch.h"

// This header is compiled twice (kernels_strict.cu / kernels_fast.cu):
//   SENV_NS         namespace of the instantiation
//   SENV_FAST_MATH  0: the reference's formulas statement by statement, compiled with -fmad=false;
//                   1: algebraically identical rewrites that drop transcendental calls (wind force
//                      without atan2/sincos/sin, relative-wind angle by the angle-difference identity),
//                      compiled with FMA contraction.  Both meet the same parity tests.
#ifndef SENV_NS
#error "define SENV_NS and SENV_FAST_MATH before including shipenv_kernels.cuh"
#endif

#ifndef SENV_REFILL_MIN
#define SENV_REFILL_MIN 1   // lane pairs of a warp that must be free before they fetch (1 = fetch at once;
                            // batching 2/4/8 measured slower: 9.63 / 9.87 / 11.5 ms vs 9.35 ms per episode)
#endif
#ifndef SENV_MIN_BLOCKS
#define SENV_MIN_BLOCKS 4   // resident CTAs per SM the env kernel is compiled for (128 registers/thread)
#endif
#ifndef SENV_MIN_BLOCKS_QUIET
#define SENV_MIN_BLOCKS_QUIET SENV_MIN_BLOCKS   // the same for the instantiations with quiet steps (SENV_QUIET)
#endif
#ifndef SENV_ENV_BLOCK
#define SENV_ENV_BLOCK 128  // threads per CTA of the env kernel
#endif
#ifndef SENV_VMAG_RSQRT
#define SENV_VMAG_RSQRT 1   // fast build: relative wind speed as x * rsqrt(x), within 2 ulp (3 links fewer: +0.6 %)
#endif
#ifndef SENV_RUDDER_REASSOC
#define SENV_RUDDER_REASSOC 1
#endif
#ifndef SENV_SEG_SMEM
#define SENV_SEG_SMEM 0     // LOS segment cache in shared memory instead of 14 registers (measured: colav_iw +0.9 %,
                            // rl -2 %, one step per launch +4 %; profiles/r02_ncu_summary.md part 5)
#endif
#ifndef SENV_QUIET
#define SENV_QUIET 1        // env kernel: event tests skipped while every lane of the warp is provably far from any event (k_env)
#endif
#ifndef SENV_LOS_RSQRT
#define SENV_LOS_RSQRT 1    // fast build: e_ct / sqrt(R^2 - e_ct^2) as e_ct * rsqrt(.) (11 dependent FP64 links fewer: +5 %)
#endif

namespace SENV_NS {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr double kPi = 3.141592653589793;

#include "shipenv_math.cuh"

using MapGrid = SenvGrid;
using DevView = SenvView;

enum Mode { MODE_STEP = 0, MODE_SUBSTEPS = 1 };

// ------------------------------------------------------------------------------------------------
// registers of one ship
// ------------------------------------------------------------------------------------------------
// Current LOS segment (wp[k-1] -> wp[k]) and its bearing: seven doubles that change only at a waypoint switch.
// SENV_SEG_SMEM = 1 keeps them in a per-lane shared-memory slot ([field][thread], conflict-free) instead of fourteen
// registers: the simulator step re-reads them (LDS, off the dependent chains) and the registers go to the overlap
// of the step's two long FP64 chains, which ptxas otherwise serialises at 128 registers per thread.
#if SENV_SEG_SMEM
constexpr int kSegStride = 128;     // threads per CTA of every kernel that holds a Ship
struct Seg {
  double* slot;
  __device__ __forceinline__ double& pn() const { return slot[0]; }
  __device__ __forceinline__ double& pe() const { return slot[kSegStride]; }
  __device__ __forceinline__ double& sin_a() const { return slot[2 * kSegStride]; }
  __device__ __forceinline__ double& cos_a() const { return slot[3 * kSegStride]; }
  __device__ __forceinline__ double& wn() const { return slot[4 * kSegStride]; }
  __device__ __forceinline__ double& we() const { return slot[5 * kSegStride]; }
  __device__ __forceinline__ double& alpha() const { return slot[6 * kSegStride]; }
};
#define SENV_SEG_SLOT(ship) __shared__ double seg_slots_[7][kSegStride]; (ship).seg.slot = &seg_slots_[0][threadIdx.x]
#else
struct Seg {
  double pn_, pe_, wn_, we_, alpha_, sin_a_, cos_a_;
  __device__ __forceinline__ double& pn() { return pn_; }
  __device__ __forceinline__ double& pe() { return pe_; }
  __device__ __forceinline__ double& sin_a() { return sin_a_; }
  __device__ __forceinline__ double& cos_a() { return cos_a_; }
  __device__ __forceinline__ double& wn() { return wn_; }
  __device__ __forceinline__ double& we() { return we_; }
  __device__ __forceinline__ double& alpha() { return alpha_; }
  __device__ __forceinline__ const double& pn() const { return pn_; }
  __device__ __forceinline__ const double& pe() const { return pe_; }
  __device__ __forceinline__ const double& sin_a() const { return sin_a_; }
  __device__ __forceinline__ const double& cos_a() const { return cos_a_; }
  __device__ __forceinline__ const double& wn() const { return wn_; }
  __device__ __forceinline__ const double& we() const { return we_; }
  __device__ __forceinline__ const double& alpha() const { return alpha_; }
};
#define SENV_SEG_SLOT(ship) ((void)0)
#endif

struct Ship {
  double north, east, yaw, u, v, r, omega, time;
  double e_ct, e_ct_int;
  double hdg_err_i, hdg_prev_err, spd_err_i, spd_aux;
  Seg seg;
  int k;          // next waypoint index
  int n_wp;       // route length (file waypoints + sampled intermediate waypoints)
  int stop;
};

struct Route {            // route of this lane: fixed file waypoints + per-env sampled waypoints
  const double* file_n;   // shared memory
  const double* file_e;
  const double* iw_n;     // global, stride num_envs; nullptr when the route is fixed
  const double* iw_e;
  long long stride;
  int n_file;
  // bearings: of the file route's segments (shared memory, [k][3] = alpha, sin, cos of wp[k-1] -> wp[k]) and of
  // the two segments at the newest sampled waypoint (env_f64 rows SEG_NEW_* / SEG_END_* of this environment)
  const double* seg_file;
  const double* seg_env;
};

__device__ __forceinline__ void route_wp(const Route& rt, int n_iw, int idx, double& n, double& e) {
  // list.insert(-1, .) keeps the file's last waypoint last (controllers.py:417-422)
  const int head = rt.n_file - 1;
  if (idx < head) { n = rt.file_n[idx]; e = rt.file_e[idx]; }
  else if (idx < head + n_iw) { n = rt.iw_n[(long long)(idx - head) * rt.stride]; e = rt.iw_e[(long long)(idx - head) * rt.stride]; }
  else { n = rt.file_n[head]; e = rt.file_e[head]; }
}

// end points of the current segment (the bearing and its sin / cos come from the cache rows of ship_f64)
__device__ __forceinline__ void load_segment_points(const Route& rt, int n_iw, Ship& s) {
  route_wp(rt, n_iw, s.k - 1, s.seg.pn(), s.seg.pe());
  route_wp(rt, n_iw, s.k, s.seg.wn(), s.seg.we());
}

// bearing of a segment and its sin / cos (LOS_guidance.py:105-107).  (Moving this and the polygon edge loops out
// of line to shrink the simulator loop's instruction footprint was measured: -13 % code, but the call overhead
// and extra spills cost 1-2 %; see profiles/r01_ncu_summary.md part 2.)
__device__ __forceinline__ double3 segment_bearing(double dx, double dy) {
  double3 out;
  out.x = senv_atan2(dy, dx);
  senv_sincos(out.x, &out.y, &out.z);
  return out;                                                // by value: alpha, sin, cos stay in registers
}

// New target waypoint: end points and bearing of the segment.  The bearings of the file route's segments are
// tabulated once per CTA, those of the segments at the newest sampled waypoint by the step() prologue -- the values
// LOS_guidance.py:105-110 computes, without an atan2 + sincos for a single lane of the warp at every waypoint switch.
__device__ __forceinline__ void refresh_segment(const Route& rt, int n_iw, Ship& s) {
  load_segment_points(rt, n_iw, s);
  const int head = rt.n_file - 1;
  if (n_iw == 0 || s.k < head) {
    s.seg.alpha() = rt.seg_file[3 * s.k]; s.seg.sin_a() = rt.seg_file[3 * s.k + 1]; s.seg.cos_a() = rt.seg_file[3 * s.k + 2];
  } else if (s.k >= head + n_iw - 1 && rt.seg_env) {
    const int row = (s.k == head + n_iw) ? SHIPENV_EF_SEG_END_ALPHA : SHIPENV_EF_SEG_NEW_ALPHA;
    s.seg.alpha() = rt.seg_env[(long long)row * rt.stride];
    s.seg.sin_a() = rt.seg_env[(long long)(row + 1) * rt.stride];
    s.seg.cos_a() = rt.seg_env[(long long)(row + 2) * rt.stride];
  } else {                                                   // an older sampled segment: not tabulated
    const double3 b = segment_bearing(s.seg.wn() - s.seg.pn(), s.seg.we() - s.seg.pe());
    s.seg.alpha() = b.x; s.seg.sin_a() = b.y; s.seg.cos_a() = b.z;
  }
}

__device__ __forceinline__ double sat(double val, double low, double hi) {   // controllers.py:67-72
  const double m = (hi < val) ? hi : val;
  return (m > low) ? m : low;
}

// Per-ship constants the simulator step would otherwise recompute at every step (products and sums of
// parameters only, evaluated once per CTA by stage_params with the same IEEE operations, so the values are the
// ones the step computed inline).  derived_of(P) finds the block of the ship whose parameters P refers to: the
// two blocks are laid out in shared memory with the stride of ShipEnvShipParams, at a fixed distance from it.
struct alignas(16) Derived {
  // -- in the order the simulator step reads them (adjacent pairs load as one LDS.128)
  double los_ra2, los_r2;                 // ra * ra (LOS_guidance.py:88-92), R * R
  double los_r99, los_limit;              // 0.99 * R (LOS_guidance.py:111-113)
  double los_ki;
  double inv_ctrl_dt, ctrl_dt, hdg_kp, hdg_kd, hdg_ki, max_rudder, desired_speed, spd_kp, spd_kd, spd_ki, max_thrust, dt, cur_n, cur_e, c_rudder_v, c_rudder_r, cos_wind_dir, sin_wind_dir, wind_speed;
  double wind_cu, wind_cv, wind_cn;       // -0.3 A_f, -0.42 A_l, -0.096 A_l L (fast build's wind force)
  double mass, y_dv, x_du, lin_damp_u, ku, lin_damp_v, kv, lin_damp_r, kr, inv_m_u, inv_m_v, inv_m_r;
  double hz_min_n, hz_max_n, hz_min_e, hz_max_e;   // map horizon moved in by half a ship length (check_condition.py:96-119)
  double l_ship, nav_fail_tol, sim_time;
  // -- machinery models (detailed / simplified)
  double kp_ship_speed, ki_ship_speed, max_shaft_speed, kp_shaft_speed, ki_shaft_speed, p_me, p_el, tq_me_max, tq_el_max, d_me, r_me, d_hsg, r_hsg, k_torque, jp, thrust_coeff, dt_shaft, k_thrust, thrust_tau;
};
struct alignas(16) DerivedSlot {
  Derived d;
  char pad[sizeof(ShipEnvShipParams) - sizeof(Derived)];
};
struct SharedBlock;
__device__ __forceinline__ const Derived& derived_of(const ShipEnvShipParams& P);

// ------------------------------------------------------------------------------------------------
// one simulator step of one ship: autopilot (LOS + heading PID), speed controller, hull + machinery
// derivatives, forward Euler.  SURVEY.md Appendix A; ship_model.py:351-416, rl_env ship_model.py:
// 834-901, ship_engine.py:403-443, controllers.py:106-125,183-189,286-295,425-433,
// LOS_guidance.py:83-117, utils.py:42-53.
// ------------------------------------------------------------------------------------------------
// NavigationSystem.los_guidance on the cached segment wp[k-1] -> wp[k] (LOS_guidance.py:100-117): updates
// e_ct and the LOS integrator, returns the heading reference.
__device__ __forceinline__ double los_guidance(const ShipEnvShipParams& P, Ship& s) {
  const Derived& D = derived_of(P);
  double e_ct = -(s.north - s.seg.pn()) * s.seg.sin_a() + (s.east - s.seg.pe()) * s.seg.cos_a();
  const double R2 = D.los_r2;
  if (e_ct * e_ct >= R2) e_ct = D.los_r99;
  s.e_ct = e_ct;
#if SENV_LOS_RSQRT && defined(SENV_HAVE_OWN_SQRT_DIV)
  // e_ct / sqrt(R^2 - e_ct^2) as e_ct * rsqrt(..): the seed and the first refinement of the root (5 dependent FP64
  // links instead of the 16 of root + division).  The refined reciprocal root is within 0.5 ulp + 2^-60 of the
  // exact one (the correction term is ~2^-21 of it), the product adds one rounding: the same 2^-52 relative error
  // bound as the reference's correctly rounded root followed by its correctly rounded division.  R^2 - e_ct^2 >=
  // 0.0199 R^2 here, so the reference's clamp of the root to >= 1e-6 cannot bind for R > 1e-5 (checked on the host).
  const double q = e_ct * senv_rsqrt(R2 - e_ct * e_ct);
#else
  double delta = SENV_SQRT(R2 - e_ct * e_ct);
  if (!(delta > 1e-6)) delta = 1e-6;
  const double q = SENV_DIV(e_ct, delta);
#endif
  if (fabs(s.e_ct_int + q) <= D.los_limit) s.e_ct_int += q;
  const double chi_r = senv_atan(-q - s.e_ct_int * D.los_ki);
  return s.seg.alpha() + chi_r;
}

// heading_offset / speed_factor: SBMPC's course offset (already negated, "pos == clockwise in sim",
// env.py:394) and speed factor; (-0.0, 1.0) without SBMPC.
// `hook(north, east)` is called with the ship's new position as soon as the kinematics have it -- the same
// expressions as the Euler update at the end of the step -- so the caller can start loads that depend on it (the
// env kernel's culling-grid lookup) before the ~100 instructions of kinetics instead of after them.
struct NoStepHook { __device__ __forceinline__ void operator()(double, double) const {} };

// NavigationSystem.next_wpt (LOS_guidance.py:83-98): the waypoint switch at the top of the autopilot call.
__device__ __forceinline__ bool wpt_reached(const Derived& H, const Ship& s) {
  const double dn = s.seg.wn() - s.north, de = s.seg.we() - s.east;
  return (dn * dn + de * de <= H.los_ra2) && (s.n_wp > s.k + 1);
}

// The step is written as ONE basic block from the LOS guidance to the derivatives: the simulator loop is bound by
// the latency of its dependent FP64 chains (8 cycles per link), and the two long ones -- cross-track error -> root
// -> division -> atan -> heading PID -> rudder forces, and sincos(psi) -> kinematics -> current / wind -> root ->
// kinetics -- are independent until the force balance, so the scheduler can overlap them when no branch lies
// between them (profiles/r02_ncu_summary.md).  Hence:
//   SWITCH_AT_TOP = false: the caller performs the waypoint switch (wpt_reached + refresh_segment) after the
//     previous step's integration -- the same position the reference tests at the top of this step -- inside its
//     rarely taken event path, instead of a branch in front of every step;
//   COLLAV_SIMPLE: the 'simple' collision avoidance block exists only in the instantiations that need it;
//   the trajectory-log stores (a run-time option) sit behind everything else, right before the Euler update.
template <int MODEL, bool SWITCH_AT_TOP = true, bool COLLAV_SIMPLE = false, class Hook = NoStepHook>
__device__ __forceinline__ void ship_step(const ShipEnvShipParams& P, const Route& rt, int n_iw, Ship& s,
                                          bool collav_hit, double collav_bias, double heading_offset,
                                          double speed_factor, double* log_row = nullptr, Hook hook = Hook(),
                                          double2* yaw_sc = nullptr) {
  const Derived& H = derived_of(P);   // the step's parameters in reading order (see Derived)
  // --- NavigationSystem.next_wpt
  if (SWITCH_AT_TOP) {
    if (wpt_reached(H, s)) { s.k += 1; refresh_segment(rt, n_iw, s); }
  }
  // --- kinematics
  double spsi, cpsi;
  // yaw_sc (optional): sin / cos of the heading carried from the previous step -- the caller needs them for the
  // heading the step ends with (the encounter type of the AST reward), and that heading is the next step's input
  if (yaw_sc) { spsi = yaw_sc->x; cpsi = yaw_sc->y; }
  else senv_sincos(s.yaw, &spsi, &cpsi);
  const double u = s.u, v = s.v, r = s.r;
  const double dt = H.dt;
  const double d_north = cpsi * u + (-spsi) * v;
  const double d_east = spsi * u + cpsi * v;
  const double new_north = s.north + d_north * dt, new_east = s.east + d_east * dt, new_yaw = s.yaw + r * dt;
  hook(new_north, new_east);
  double2 new_sc = make_double2(0.0, 1.0);
  if (yaw_sc) senv_sincos(new_yaw, &new_sc.x, &new_sc.y);
  // --- NavigationSystem.los_guidance
  const double heading_ref = los_guidance(P, s);
  // --- heading PID -> rudder angle
  double rudder;
  {
    const double error = (heading_ref + heading_offset) - s.yaw;
#if SENV_FAST_MATH
    const double d_error = (error - s.hdg_prev_err) * H.inv_ctrl_dt;
#else
    const double d_error = (error - s.hdg_prev_err) / H.ctrl_dt;
#endif
    const double error_i = s.hdg_err_i + error * H.ctrl_dt;
    s.hdg_prev_err = error;
    s.hdg_err_i = error_i;
    const double out = error * H.hdg_kp + d_error * H.hdg_kd + error_i * H.hdg_ki;
    rudder = sat(-out, -H.max_rudder, H.max_rudder);
  }
  // --- speed controller
  double cmd;
  const double speed_set_point = H.desired_speed * speed_factor;
  if (MODEL == SHIPENV_MODEL_SIMPLE) {
    const double error = speed_set_point - s.u;
#if SENV_FAST_MATH
    const double d_error = (error - s.spd_aux) * H.inv_ctrl_dt;
#else
    const double d_error = (error - s.spd_aux) / H.ctrl_dt;
#endif
    const double error_i = s.spd_err_i + error * H.ctrl_dt;
    s.spd_aux = error;
    s.spd_err_i = error_i;
    const double out = error * H.spd_kp + d_error * H.spd_kd + error_i * H.spd_ki;
    cmd = sat(out, -H.max_thrust, H.max_thrust);
  } else if (MODEL == SHIPENV_MODEL_SIMPLIFIED) {
    // ThrottleFromSpeedSetPointSimplifiedPropulsion.throttle (rl_env controllers.py:229-232)
    const double error = speed_set_point - s.u;
    const double error_i = s.spd_err_i + error * H.ctrl_dt;
    s.spd_err_i = error_i;
    cmd = sat(error * H.kp_ship_speed + error_i * H.ki_ship_speed, 0.0, 1.1);
  } else {
    const double error = speed_set_point - s.u;
    const double error_i = s.spd_err_i + error * H.ctrl_dt;
    s.spd_err_i = error_i;
    const double w_d = sat(error * H.kp_ship_speed + error_i * H.ki_ship_speed, 0.0, H.max_shaft_speed);
    // measured_shaft_speed = forward_speed (rl_env env.py:397-401)
    const double error2 = w_d - s.u;
    const double error2_i = s.spd_aux + error2 * H.ctrl_dt;
    s.spd_aux = error2_i;
    cmd = sat(error2 * H.kp_shaft_speed + error2_i * H.ki_shaft_speed, 0.0, 1.1);
  }
  // --- collav 'simple' (run_colav env.py:1189-1202, rl_env env.py:405-418)
  if (COLLAV_SIMPLE && collav_hit) {
    cmd *= 0.5;
    cmd = (cmd < 0.0) ? 0.0 : ((cmd > 1.1) ? 1.1 : cmd);
    rudder += collav_bias;
    rudder = (rudder < -H.max_rudder) ? -H.max_rudder : ((rudder > H.max_rudder) ? H.max_rudder : rudder);
  }
  // --- machinery
  double thrust, d_omega = 0.0;
  if (MODEL == SHIPENV_MODEL_SIMPLE) {
    thrust = cmd;
  } else if (MODEL == SHIPENV_MODEL_SIMPLIFIED) {
    // SimplifiedMachineryModel.update_thrust_force (ship_engine.py:508-513); the thrust state feeds the
    // kinetics before it is integrated
    const double power = cmd * (H.p_me + H.p_el);
    thrust = s.omega;
    d_omega = SENV_DIV(-H.k_thrust * s.omega + power, H.thrust_tau);
  } else {
    const double w = s.omega;
    // (the divisions stay IEEE divisions in both builds -- replacing them by reciprocal multiplications moved the
    //  ill-conditioned detailed model past 1e-9 on one golden episode, rl_dt4_PTO; the fast build takes them by
    //  the library's fast-path sequence without its slow-path branch, same bits, shipenv_math.cuh)
    const double a_me = SENV_DIV(cmd * H.p_me, w + 0.1);
    const double tq_me = (H.tq_me_max < a_me) ? H.tq_me_max : a_me;
    const double a_el = SENV_DIV(cmd * H.p_el, w + 0.1);
    const double tq_el = (H.tq_el_max < a_el) ? H.tq_el_max : a_el;
    const double eq_me = SENV_DIV(tq_me - H.d_me * w, H.r_me);
    const double eq_hsg = SENV_DIV(tq_el - H.d_hsg * w, H.r_hsg);
    d_omega = SENV_DIV(eq_me + eq_hsg - H.k_torque * (w * w), H.jp);
    thrust = H.thrust_coeff * w * fabs(w);
  }
  // --- current in the body frame, rudder forces
  const double u_c = cpsi * H.cur_n + spsi * H.cur_e;
  const double v_c = (-spsi) * H.cur_n + cpsi * H.cur_e;
  const double u_r = u - u_c, v_r = v - v_c;
  double f_rudder_v, f_rudder_r;
  if (SENV_FAST_MATH && SENV_RUDDER_REASSOC && MODEL == SHIPENV_MODEL_SIMPLE) {
    // (-c (u - u_c)) * rudder instead of (-c rudder) * (u - u_c): the factor without the rudder angle is ready early, so
    // the force follows the angle -- the end of the step's longest chain -- by one FP64 link instead of two.  Simple
    // model only: its block's static schedule shrinks from 767 to 718 cycles, the detailed model's grows (745 -> 767).
    f_rudder_v = (-H.c_rudder_v * (u - u_c)) * rudder;
    f_rudder_r = (-H.c_rudder_r * (u - u_c)) * rudder;
  } else {
    f_rudder_v = -H.c_rudder_v * rudder * (u - u_c);
    f_rudder_r = -H.c_rudder_r * rudder * (u - u_c);
  }
  // --- wind (get_wind_force, ship_model.py:162-175)
#if SENV_FAST_MATH
  // u_rw = ws*cos(wd - psi) - u, v_rw = ws*sin(wd - psi) - v with the angle-difference identity;
  // gamma = -atan2(v_rw, u_rw)  =>  cos(gamma) = u_rw/|V|, sin(gamma) = -v_rw/|V|,
  // sin(2 gamma) = -2 u_rw v_rw/|V|^2, so with q = 0.6 |V|^2:
  //   tau_u = q*(-0.5 cos g)*A_f = -0.3 A_f |V| u_rw,  tau_v = q*(0.7 sin g)*A_l = -0.42 A_l |V| v_rw,
  //   tau_n = q*(0.08 sin 2g)*A_l*L = -0.096 A_l L u_rw v_rw
  const double cw = H.cos_wind_dir * cpsi + H.sin_wind_dir * spsi;
  const double sw = H.sin_wind_dir * cpsi - H.cos_wind_dir * spsi;
  const double u_rw = H.wind_speed * cw - u;
  const double v_rw = H.wind_speed * sw - v;
#if SENV_VMAG_RSQRT && defined(SENV_HAVE_OWN_SQRT_DIV)
  const double vmag = senv_sqrt_2ulp(u_rw * u_rw + v_rw * v_rw);
#else
  const double vmag = SENV_SQRT(u_rw * u_rw + v_rw * v_rw);
#endif
  const double tau_u = H.wind_cu * vmag * u_rw;
  const double tau_v = H.wind_cv * vmag * v_rw;
  const double tau_n = H.wind_cn * u_rw * v_rw;
#else
  double sw, cw;
  sincos(P.wind_dir - s.yaw, &sw, &cw);
  const double u_rw = H.wind_speed * cw - u;
  const double v_rw = H.wind_speed * sw - v;
  const double gamma_rw = -atan2(v_rw, u_rw);
  const double wind_rw2 = u_rw * u_rw + v_rw * v_rw;
  double sg, cg;
  sincos(gamma_rw, &sg, &cg);
  const double c_x = -0.5 * cg;
  const double c_y = 0.7 * sg;
  const double c_n = 0.08 * sin(2 * gamma_rw);
  const double tau_coeff = 0.5 * 1.2 * wind_rw2;
  const double tau_u = tau_coeff * c_x * P.proj_area_f;
  const double tau_v = tau_coeff * c_y * P.proj_area_l;
  const double tau_n = tau_coeff * c_n * P.proj_area_l * P.l_ship;
#endif
  // --- kinetics (x_g = 0, diagonal mass matrix)
  const double m = H.mass;
  const double crb0 = (-m * v) * r;
  const double crb1 = (m * u) * r;
  const double crb2 = (m * v) * u + (-m * u) * v;
  const double ca0 = (H.y_dv * v_r) * r;
  const double ca1 = (-H.x_du * u_r) * r;
  const double ca2 = (-H.y_dv * v_r) * u_r + (H.x_du * u_r) * v_r;
  const double dmp0 = (H.lin_damp_u + H.ku * u) * u_r;
  const double dmp1 = (H.lin_damp_v + H.kv * v) * v_r;
  const double dmp2 = (H.lin_damp_r + H.kr * r) * r;
  const double f0 = -crb0 - ca0 - dmp0 + tau_u + 0.0 + thrust;
  const double f1 = -crb1 - ca1 - dmp1 + tau_v + 0.0 + f_rudder_v;
  const double f2 = -crb2 - ca2 - dmp2 + tau_n + 0.0 + f_rudder_r;
  const double d_u = H.inv_m_u * f0;
  const double d_v = H.inv_m_v * f1;
  const double d_r = H.inv_m_r * f2;
  // --- store_simulation_data (ship_model.py:418-429): the row holds the state before the integration
  if (log_row) {
    log_row[SHIPENV_LOG_TIME] = s.time; log_row[SHIPENV_LOG_NORTH] = s.north; log_row[SHIPENV_LOG_EAST] = s.east;
    log_row[SHIPENV_LOG_YAW] = s.yaw; log_row[SHIPENV_LOG_RUDDER] = rudder; log_row[SHIPENV_LOG_U] = s.u;
    log_row[SHIPENV_LOG_V] = s.v; log_row[SHIPENV_LOG_R] = s.r; log_row[SHIPENV_LOG_OMEGA] = s.omega;
    log_row[SHIPENV_LOG_CMD] = cmd; log_row[SHIPENV_LOG_E_CT] = s.e_ct;
    log_row[SHIPENV_LOG_E_PSI] = fabs(s.yaw - heading_ref);            // get_heading_error, controllers.py:396-397
  }
  // --- forward Euler
  s.north = new_north;
  s.east = new_east;
  s.yaw = new_yaw;
  s.u = s.u + d_u * dt;
  s.v = s.v + d_v * dt;
  s.r = s.r + d_r * dt;
  if (MODEL != SHIPENV_MODEL_SIMPLE) s.omega = s.omega + d_omega * H.dt_shaft;
  s.time = s.time + dt;
  if (yaw_sc) *yaw_sc = new_sc;
}

// ------------------------------------------------------------------------------------------------
// map geometry (obstacle.py:126-141, check_condition.py:16-119); Shapely semantics: contains() is the
// strict interior (even-odd rule), exterior.distance() the distance to the closed ring.
// ------------------------------------------------------------------------------------------------
struct MapView {
  const double* ve;
  const double* vn;
  const int* start;
  const double* bbox;   // [n_poly][4] = min_e, max_e, min_n, max_n (shared memory)
  const unsigned char* next;   // [n_vert] index of the vertex that ends ring segment i (shared memory)
  int n_poly;
  MapGrid grid;
};

// bits 0..15: candidate polygons for contains(); bits 16..31: candidates for the clipped distance
__device__ __forceinline__ unsigned map_cell_masks(const MapView& mp, double n_pos, double e_pos) {
  const unsigned all = (mp.n_poly >= 16) ? 0xffffu : ((1u << mp.n_poly) - 1u);
  const double fx = (e_pos - mp.grid.e0) * mp.grid.inv_cell;
  const double fy = (n_pos - mp.grid.n0) * mp.grid.inv_cell;
  // outside the grid (or NaN): fall back to every polygon.  No branch: the load always happens (cell 0 then), so
  // it can be issued early and overlap other work.
  const bool inside = ((int)(fx >= 0.0) & (int)(fy >= 0.0) & (int)(fx < mp.grid.nx_f) & (int)(fy < mp.grid.ny_f)) != 0;
  const int ix = (int)(inside ? fx : 0.0), iy = (int)(inside ? fy : 0.0);
  const unsigned m = __ldg(mp.grid.cells + iy * mp.grid.nx + ix);
  return inside ? m : (all | (all << 16));
}

// Safe radius of the cell the point lies in (SenvGrid::safe): a ship that has moved less than this from the point
// cannot be grounded.  0 outside the grid (or NaN).  Single precision: a position within a millimetre of a cell border
// may read the neighbouring cell's radius, which is off by that millimetre at most.
__device__ __forceinline__ float map_safe_radius(const MapView& mp, float n_pos, float e_pos) {
  const float fx = (e_pos - (float)mp.grid.e0) * (float)mp.grid.inv_cell;
  const float fy = (n_pos - (float)mp.grid.n0) * (float)mp.grid.inv_cell;
  if (!(fx >= 0.0f && fy >= 0.0f && fx < (float)mp.grid.nx && fy < (float)mp.grid.ny)) return 0.0f;
  return __ldg(mp.grid.safe + (int)fy * mp.grid.nx + (int)fx);
}

// Polygon.contains(Point(x, y)) for one polygon: even-odd crossing rule
__device__ __forceinline__ bool poly_contains(const MapView& mp, int p, double x, double y) {
  // a point outside the polygon's bounding box crosses an even number of edges (exact skip; the step() prologue
  // stages no bounding boxes and walks the edges)
  if (mp.bbox) {
    const double* bb = mp.bbox + 4 * p;
    if (x < bb[0] || x > bb[1] || y < bb[2] || y > bb[3]) return false;
  }
  const int a = mp.start[p], b = mp.start[p + 1];
  bool inside = false;
  double xj = mp.ve[b - 1], yj = mp.vn[b - 1];
#pragma unroll 1
  for (int i = a; i < b; ++i) {
    const double xi = mp.ve[i], yi = mp.vn[i];
    if ((yi > y) != (yj > y)) {
      if (x < SENV_DIV((xj - xi) * (y - yi), yj - yi) + xi) inside = !inside;   // (yj != yi: the edge crosses y)
    }
    xj = xi; yj = yi;
  }
  return inside;
}

// PolygonObstacle.if_pos_inside_obstacles (obstacle.py:126-129) over the polygons in `mask`
__device__ __forceinline__ bool map_contains(const MapView& mp, unsigned mask, double n_pos, double e_pos) {
  while (mask) {
    const int p = __ffs(mask) - 1;
    mask &= mask - 1;
    if (poly_contains(mp, p, e_pos, n_pos)) return true;
  }
  return false;
}

// PolygonObstacle.obstacles_distance (obstacle.py:138-141): min over polygons of the ring distance.
// Only its value when <= 1000 m matters to the caller (reward_function.py:386-389, 454-457).  The grid
// cell lists the ring segments that can be nearest to a point of the cell within that clip (a segment
// whose lower distance bound over the cell exceeds the smallest upper bound cannot be the nearest one),
// so the minimum over the listed segments equals the minimum over all segments whenever it is <= 1000 m,
// and when no segment is listed the reward term is 0 either way.
__device__ __forceinline__ double map_distance(const MapView& mp, double n_pos, double e_pos) {
  const double px = e_pos, py = n_pos;
  unsigned long long m0, m1;
  {
    const double fx = (e_pos - mp.grid.e0) * mp.grid.inv_cell;
    const double fy = (n_pos - mp.grid.n0) * mp.grid.inv_cell;
    if (!(fx >= 0.0 && fy >= 0.0 && fx < mp.grid.nx_f && fy < mp.grid.ny_f)) {
      // outside the grid (or NaN): every segment
      const int nv = mp.start[mp.n_poly];
      m0 = (nv >= 64) ? ~0ull : ((1ull << nv) - 1ull);
      m1 = (nv > 64) ? ((nv >= 128) ? ~0ull : ((1ull << (nv - 64)) - 1ull)) : 0ull;
    } else {
      const ulonglong2 mm = __ldg(reinterpret_cast<const ulonglong2*>(mp.grid.edges) + ((int)fy * mp.grid.nx + (int)fx));
      m0 = mm.x; m1 = mm.y;
    }
  }
  double best2 = INFINITY;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    unsigned long long m = half ? m1 : m0;
    while (m) {
      const int i = __ffsll((long long)m) - 1 + 64 * half;
      m &= m - 1;
      // segment i -> next vertex of the same polygon (mp.next[i])
      const int k = mp.next[i];
      const double ax = mp.ve[i], ay = mp.vn[i], bx = mp.ve[k], by = mp.vn[k];
      const double dx = bx - ax, dy = by - ay;
      const double l2 = dx * dx + dy * dy;
      double t = 0.0;
      if (l2 != 0.0) {
        t = SENV_DIV((px - ax) * dx + (py - ay) * dy, l2);
        t = (t < 1.0) ? t : 1.0;
        t = (t > 0.0) ? t : 0.0;
      }
      const double cx = ax + t * dx, cy = ay + t * dy;
      const double d2 = (px - cx) * (px - cx) + (py - cy) * (py - cy);
      best2 = (d2 < best2) ? d2 : best2;
    }
  }
  // sqrt is monotonic and correctly rounded: sqrt(min d2) == min sqrt(d2); no listed segment: inf
  return (best2 < INFINITY) ? SENV_SQRT(best2) : INFINITY;
}

// is_pos_inside_obstacles (check_condition.py:48-78): is any of the four corners of the L x L square
// around the ship inside any polygon.  One pass over a polygon's edges serves all four corners (two
// distinct y values -> two crossing abscissae per edge, each compared with the two x values).
__device__ __forceinline__ bool pos_inside_obstacles_slow(const double* ve, const double* vn, const int* start,
                                                          const double* bbox, unsigned mask, double n, double e,
                                                          double ship_length);

// the common case (no polygon near the ship's grid cell) stays inline; the edge loops are out of line
__device__ __forceinline__ bool pos_inside_obstacles(const MapView& mp, unsigned mask, double n, double e,
                                                     double ship_length) {
  if (mask == 0) return false;
  return pos_inside_obstacles_slow(mp.ve, mp.vn, mp.start, mp.bbox, mask, n, e, ship_length);
}

__device__ __forceinline__ bool pos_inside_obstacles_slow(const double* ve, const double* vn, const int* start,
                                                          const double* bbox, unsigned mask, double n, double e,
                                                          double ship_length) {
  struct { const double* ve; const double* vn; const int* start; const double* bbox; } mp{ve, vn, start, bbox};
  const double margin = ship_length / 2;
  const double y0 = n - margin, x0 = e - margin, y1 = n + margin, x1 = e + margin;
  while (mask) {
    const int p = __ffs(mask) - 1;
    mask &= mask - 1;
    const double* bb = mp.bbox + 4 * p;
    // the square misses the polygon's bounding box: every corner crosses an even number of edges
    if (x1 < bb[0] || x0 > bb[1] || y1 < bb[2] || y0 > bb[3]) continue;
    const int a = mp.start[p], b = mp.start[p + 1];
    unsigned in = 0;   // bit 0: (y0,x0)  bit 1: (y0,x1)  bit 2: (y1,x0)  bit 3: (y1,x1)
    double xj = mp.ve[b - 1], yj = mp.vn[b - 1];
    // not unrolled: the loop runs only for ships next to a polygon, and its unrolled body (8 FP64 divisions) put
    // 11 KB of rarely executed code into the simulator loop's instruction-cache footprint
#pragma unroll 1
    for (int i = a; i < b; ++i) {
      const double xi = mp.ve[i], yi = mp.vn[i];
      if ((yi > y0) != (yj > y0)) {
        const double xc = SENV_DIV((xj - xi) * (y0 - yi), yj - yi) + xi;
        in ^= (x0 < xc ? 1u : 0u) | (x1 < xc ? 2u : 0u);
      }
      if ((yi > y1) != (yj > y1)) {
        const double xc = SENV_DIV((xj - xi) * (y1 - yi), yj - yi) + xi;
        in ^= (x0 < xc ? 4u : 0u) | (x1 < xc ? 8u : 0u);
      }
      xj = xi; yj = yi;
    }
    if (in) return true;
  }
  return false;
}

__device__ __forceinline__ double py_mod(double a, double b) {   // Python / NumPy float modulo
  double m = senv_fmod(a, b, 1.0 / b);                          // (b is a literal at both call sites: 1 / b folds)
  if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; }
  else m = copysign(0.0, b);
  return m;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int lane_mask) {
  return __shfl_xor_sync(FULL_MASK, v, lane_mask);
}

// ------------------------------------------------------------------------------------------------
// SBMPC collision avoidance (sbmpc.py:90-314, sbmpc_misc.py:3-123) as the envs call it: SBMPC(tf=1000,
// dt=20) with the default SBMPCParams, one dynamic obstacle.  7 course offsets x 4 speed factors = 28
// control behaviours, each a 50-sample straight-line prediction of both ships; cost = max over the
// samples of collision cost x risk, plus the manoeuvring penalties.  KAPPA_ = 0, so the COLREGs term mu
// never contributes and is not evaluated.
//
// One warp evaluates one environment: lane b < 28 takes behaviour (Chi_ca_[b / 4], P_ca_[b % 4]), the
// reference's loop order, and the first minimum in that order wins (strict `<` in sbmpc.py:176).
// ------------------------------------------------------------------------------------------------
constexpr int kSbSamples = 50;            // int(T_ / DT_) = int(1000 / 20)
constexpr double kSbDt = 20.0;
constexpr int kSbBehaviours = 28;

struct SbmpcIn {
  double os_x, os_y, os_v;                // own ship (ship under test): east, north, sway speed
  double ob_x, ob_y, ob_psi, ob_u, ob_v;  // obstacle ship: east, north, -yaw, surge, sway
  double u_d, chi_d;                      // nominal speed / course references of the calling ship
  double chi_last, p_last;                // SBMPCParams.Chi_ca_last_, P_ca_last_
};

__device__ __forceinline__ double wrap_pmpi(double a) {     // wrap_angle_to_pmpi, sbmpc_misc.py:3-33
  return -kPi + py_mod(a - (-kPi), kPi - (-kPi));
}

#ifndef SENV_SBMPC_INLINE
#define SENV_SBMPC_INLINE __forceinline__
#endif
// cost of behaviour b (SBMPC.cost_func, sbmpc.py:190-296, for the prediction of linear_pred,
// sbmpc_misc.py:102-122, against Obstacle.calculate_trajectory, sbmpc_misc.py:58-83)
__device__ SENV_SBMPC_INLINE double sbmpc_behaviour_cost(const SbmpcIn in, int b, double obs_l, double obs_w,
                                                        const double* __restrict__ inv_t) {
  const int ic = b >> 2, jp = b & 3;
  const double chi_ca = (-30.0 + 10.0 * (double)ic) * (kPi / 180.0);       // np.deg2rad(Chi_ca_[ic])
  const double p_ca = (jp == 0) ? 0.4 : ((jp == 1) ? 0.6 : ((jp == 2) ? 0.8 : 1.0));
  const double os_l = 25.0;                                 // ShipLinearModel defaults, sbmpc_misc.py:86
  const double d_safe = 1000.0, d_close = 2000.0;
  const double PHI = 1.1955505376161157;                    // PHI_AH_ = PHI_OT_ = np.deg2rad(68.5)
  const double cos_ot = 0.9997823068017366;                 // np.cos(np.deg2rad(PHI_OT_)): degrees applied twice
  // obstacle: constant velocity along its heading
  double so, co;
  senv_sincos(in.ob_psi, &so, &co);
  const double ob_dx = ((-so) * in.ob_u + co * in.ob_v) * kSbDt;
  const double ob_dy = (co * in.ob_u + so * in.ob_v) * kSbDt;
  const double vo0 = (-so) * in.ob_u + co * in.ob_v, vo1 = co * in.ob_u + so * in.ob_v;   // rot2d, sbmpc.py:312-314
  const double n_vo = SENV_SQRT(vo0 * vo0 + vo1 * vo1);
  // own ship: heading psi_d from sample 1 on, wrapped heading and the measured sway speed at sample 0
  const double ud = in.u_d * p_ca, psi_d = in.chi_d + chi_ca;
  double s0, c0, sd, cd;
  senv_sincos(psi_d, &sd, &cd);
#if SENV_FAST_MATH
  // sample 0 uses the wrapped heading wrap(psi_d) (sbmpc_misc.py:109): same angle modulo 2 pi, so the same sin / cos
  // up to the rounding of the wrap
  s0 = sd; c0 = cd;
#else
  senv_sincos(wrap_pmpi(psi_d), &s0, &c0);
#endif
  const double zero = 0.0;
  const double os_dx1 = kSbDt * ((-sd) * ud + cd * in.os_v), os_dy1 = kSbDt * (cd * ud + sd * in.os_v);   // 0 -> 1
  const double os_dx = kSbDt * ((-sd) * ud + cd * zero), os_dy = kSbDt * (cd * ud + sd * zero);           // i -> i+1
  // world-frame velocities and what depends only on them: sample 0 (A) and samples >= 1 (B)
  const double vsA0 = (-s0) * ud + c0 * in.os_v, vsA1 = c0 * ud + s0 * in.os_v;
  const double vsB0 = (-sd) * ud + cd * zero, vsB1 = cd * ud + sd * zero;
  const double n_vsA = SENV_SQRT(vsA0 * vsA0 + vsA1 * vsA1), n_vsB = SENV_SQRT(vsB0 * vsB0 + vsB1 * vsB1);
  const bool otA = (vsA0 * vo0 + vsA1 * vo1) > cos_ot * n_vsA * n_vo && n_vsA > n_vo;
  const bool otB = (vsB0 * vo0 + vsB1 * vo1) > cos_ot * n_vsB * n_vo && n_vsB > n_vo;
  const double k_coll = 1e-6 * os_l * obs_l;
  const double nrA = SENV_SQRT((vsA0 - vo0) * (vsA0 - vo0) + (vsA1 - vo1) * (vsA1 - vo1));
  const double nrB = SENV_SQRT((vsB0 - vo0) * (vsB0 - vo0) + (vsB1 - vo1) * (vsB1 - vo1));
  const double ccA = k_coll * (nrA * nrA), ccB = k_coll * (nrB * nrB);   // K_COLL * |v_s - v_o| ** 2
  // safety distance by the sector the own ship is seen in from the obstacle (sbmpc.py:232-245)
  const double ds_ahead = d_safe + obs_l / 2, ds_behind = 0.5 * d_safe + obs_l / 2, ds_beam = d_safe + obs_w / 2;
  const double ds_ot = d_safe + os_l / 2 + obs_l / 2;
  const double ds_min = fmin(ds_ahead, fmin(ds_behind, ds_beam)), ds_max = fmax(ds_ahead, fmax(ds_behind, ds_beam));

  // Beyond every safety distance a sample contributes H0 = 0 (R = C = 0), which cannot raise H1 >= 0; sqrt is
  // monotonic, so d2 >= (largest safety distance + 1 m)^2 decides that without taking the root.
  const double ds_far = fmax(ds_max, ds_ot) + 1.0;
  const double far2 = fmin(ds_far * ds_far, 4.0e6);        // never beyond D_CLOSE_ either
  double H1 = 0.0;
  // one sample of cost_func's loop: e = obstacle - own ship, t = (i + 1) * DT_
  auto sample = [&](double e0, double e1, double t, bool ot, double cc) {
    const double d2 = e0 * e0 + e1 * e1;
    if (!(d2 < far2)) return;
    const double dist = SENV_SQRT(d2);
    if (!(dist < d_close)) return;
    bool within;
    if (ot) within = dist < ds_ot;
    else if (dist < ds_min) within = true;                   // inside every sector's safety distance
    else if (!(dist < ds_max)) within = false;
    else {
#if SENV_FAST_MATH
      // phi_o = wrap(atan2(-e1, -e0) - psi_o + pi/2) in [-pi, pi): cos(phi_o) = (e1 co - e0 so) / dist,
      // sin(phi_o) = -(e0 co + e1 so) / dist, so phi_o > PHI  <=>  sin(phi_o) > 0 and cos(phi_o) < cos(PHI)
      // (PHI = 68.5 deg lies in (0, pi)); phi_o == PHI has measure zero and falls to the "ahead" distance like
      // every phi_o < PHI.  Same decision as the atan2 form up to the rounding of either.
      const double xc = e1 * co - e0 * so, ys = -(e0 * co + e1 * so);
      const bool behind = (ys > 0.0) && (xc < 0.3665012267242973 * dist);       // cos(np.deg2rad(68.5))
      within = dist < (behind ? ds_behind : ds_ahead);
#else
      const double phi_o = wrap_pmpi(atan2(-e1, -e0) - in.ob_psi + kPi / 2);
      const double d_safe_i = (phi_o < PHI) ? ds_ahead : ((phi_o > PHI) ? ds_behind : ds_beam);
      within = dist < d_safe_i;
#endif
    }
    if (within) {
      const double q = SENV_DIV(d_safe, dist);
      const double R = SENV_DIV(1.0, t) * ((q * q) * (q * q));     // (1 / |t - t0| ** P_) * (d_safe / dist) ** Q_
      const double H0 = cc * R + 0.0;                       // + KAPPA_ * mu, KAPPA_ = 0
      if (H0 > H1) H1 = H0;
    }
  };
  double sx = in.os_x, sy = in.os_y, ox = in.ob_x, oy = in.ob_y;
  sample(ox - sx, oy - sy, kSbDt, otA, ccA);                                   // i = 0: measured state
  sx = sx + os_dx1; sy = sy + os_dy1; ox = ox + ob_dx; oy = oy + ob_dy;
  sample(ox - sx, oy - sy, 2 * kSbDt, otB, ccB);                               // i = 1
#if SENV_FAST_MATH
  // From sample 1 on both ships move on straight lines: the offset is e_1 + k w (k = i - 1), so the samples that
  // can lie within ds_far form one window of k, the roots of |e_1 + k w|^2 = far^2 widened by a sample on either
  // side.  Samples outside it contribute nothing (see above) and are not visited; positions inside it are taken as
  // e_1 + k w instead of k sequential additions (differs by the rounding of the additions, ~1e-12 m).
  {
    const double e10 = ox - sx, e11 = oy - sy;
    const double w0 = ob_dx - os_dx, w1 = ob_dy - os_dy;
    const double qa = w0 * w0 + w1 * w1, qb = e10 * w0 + e11 * w1, qc = e10 * e10 + e11 * e11 - far2 * 1.000001;
    int k_lo = 1, k_hi = kSbSamples - 2;                       // k = 0 (sample 1) is done
    if (qa > 0.0) {
      const double disc = qb * qb - qa * qc;
      if (disc < 0.0) k_hi = 0;                                // never that close
      else {
        const double root = SENV_SQRT(disc), inv = SENV_DIV(1.0, qa);
        const double lo = (-qb - root) * inv - 1.0, hi = (-qb + root) * inv + 1.0;
        if (lo > (double)k_lo) k_lo = (lo < 1.0e6) ? (int)lo : kSbSamples;
        if (hi < (double)k_hi) k_hi = (hi > -1.0e6) ? (int)ceil(hi) : 0;
      }
    } else if (!(qc < 0.0)) k_hi = 0;                          // no relative motion and out of range
    // The window's samples without branches, two at a time: almost every sample inside the window goes the whole
    // way (root, sector, both divisions), and the loop was bound by the latency of that one dependent chain.  The
    // sector logic collapses to one comparison -- ds_min <= the sector's safety distance <= ds_max, so the two
    // shortcuts of sample() decide what `dist < ds_sector` decides -- and 1 / t comes from a table (exactly rounded).
    auto sample_window = [&](int k) {
      const double kk = (double)k;
      const double e0 = e10 + kk * w0, e1 = e11 + kk * w1;
      const double d2 = e0 * e0 + e1 * e1;
      const double dist = SENV_SQRT(d2);
      const bool in_range = ((int)(d2 < far2) & (int)(dist < d_close)) != 0;
      const double xc = e1 * co - e0 * so, ys = -(e0 * co + e1 * so);
      const bool behind = ((int)(ys > 0.0) & (int)(xc < 0.3665012267242973 * dist)) != 0;
      const double ds_i = otB ? ds_ot : (behind ? ds_behind : ds_ahead);
      const double q = SENV_DIV(d_safe, dist);
      const double R = inv_t[k + 1] * ((q * q) * (q * q));
      const double H0 = ccB * R + 0.0;
      return (in_range && dist < ds_i) ? H0 : 0.0;
    };
    int k = k_lo;
    for (; k + 1 <= k_hi; k += 2) {
      const double Ha = sample_window(k), Hb = sample_window(k + 1);
      if (Ha > H1) H1 = Ha;
      if (Hb > H1) H1 = Hb;
    }
    if (k <= k_hi) {
      const double Ha = sample_window(k);
      if (Ha > H1) H1 = Ha;
    }
  }
#else
  double t = 2 * kSbDt;
#pragma unroll 2
  for (int i = 2; i < kSbSamples; ++i) {
    sx = sx + os_dx; sy = sy + os_dy; ox = ox + ob_dx; oy = oy + ob_dy;
    t += kSbDt;
    sample(ox - sx, oy - sy, t, otB, ccB);
  }
#endif
  const double d_chi = chi_ca - in.chi_last;                // delta_Chi, sbmpc.py:303-310
  double dl_chi = 0.0;
  if (d_chi > 0) dl_chi = 20.0 * (d_chi * d_chi);
  else if (d_chi < 0) dl_chi = 30.0 * (d_chi * d_chi);
  const double H2 = 25.0 * (1 - p_ca) + 30.0 * (chi_ca * chi_ca) + 20.0 * fabs(in.p_last - p_ca) + dl_chi;
  return H1 + H2;
}

// get_optimal_ctrl_offset (sbmpc.py:113-185) for the environment whose inputs lane `src` holds: all 32
// lanes take part, every lane returns the winning behaviour index (-1: no finite cost, keep (1, 0)).
// Out of line: the evaluation's registers then do not count against the simulator loop's allocation (spills of the
// SBMPC instantiations 990 -> 440 B; steps without an active pair 13 % faster, active ones 3 % slower for the call).
#ifndef SENV_SBMPC_ARGMIN_INLINE
#define SENV_SBMPC_ARGMIN_INLINE __noinline__
#endif
__device__ SENV_SBMPC_ARGMIN_INLINE int sbmpc_warp_argmin(const SbmpcIn& mine, int src, int lane, double obs_l, double obs_w,
                                                         const double* __restrict__ inv_t) {
  SbmpcIn in;
  in.os_x = __shfl_sync(FULL_MASK, mine.os_x, src); in.os_y = __shfl_sync(FULL_MASK, mine.os_y, src);
  in.os_v = __shfl_sync(FULL_MASK, mine.os_v, src);
  in.ob_x = __shfl_sync(FULL_MASK, mine.ob_x, src); in.ob_y = __shfl_sync(FULL_MASK, mine.ob_y, src);
  in.ob_psi = __shfl_sync(FULL_MASK, mine.ob_psi, src);
  in.ob_u = __shfl_sync(FULL_MASK, mine.ob_u, src); in.ob_v = __shfl_sync(FULL_MASK, mine.ob_v, src);
  in.u_d = __shfl_sync(FULL_MASK, mine.u_d, src); in.chi_d = __shfl_sync(FULL_MASK, mine.chi_d, src);
  in.chi_last = __shfl_sync(FULL_MASK, mine.chi_last, src); in.p_last = __shfl_sync(FULL_MASK, mine.p_last, src);
  double cost = INFINITY;
  if (lane < kSbBehaviours) cost = sbmpc_behaviour_cost(in, lane, obs_l, obs_w, inv_t);
  if (!(cost < INFINITY)) cost = INFINITY;      // NaN / inf never win the reference's `cost_i < cost`
  int idx = lane;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double oc = __shfl_xor_sync(FULL_MASK, cost, off);
    const int oi = __shfl_xor_sync(FULL_MASK, idx, off);
    if (oc < cost || (oc == cost && oi < idx)) { cost = oc; idx = oi; }
  }
  return (cost < INFINITY) ? idx : -1;
}

// shared-memory staging of the parameter block
struct alignas(16) SharedBlock {
  ShipEnvParams p;
  DerivedSlot drv[2];                    // see derived_of()
  double roa2, seg_len2;                 // roa * roa (check_condition.py:181-204), 2 * AB segment length
  double sb_inv_t[50];                   // SBMPC: 1 / t of prediction sample i, t = (i + 1) * DT_ (sbmpc.py:262)
  double bbox[SHIPENV_MAX_POLY * 4];
  double seg[2][SHIPENV_MAX_WP][3];      // per ship: bearing, sin, cos of the file route's segment wp[k-1] -> wp[k]
  unsigned char next[SHIPENV_MAX_VERT];
};

__device__ __forceinline__ const Derived& derived_of(const ShipEnvShipParams& P) {
  constexpr long long kOffset = (long long)offsetof(SharedBlock, drv) - (long long)offsetof(SharedBlock, p.ship);
  static_assert(sizeof(DerivedSlot) == sizeof(ShipEnvShipParams), "derived blocks must have the parameter stride");
  return *reinterpret_cast<const Derived*>(reinterpret_cast<const char*>(&P) + kOffset);
}

// Builds the block from the parameters: bounding boxes, ring successors, the per-ship derived constants, SBMPC's
// 1 / t table and the bearing tables of the file routes (64 atan2 + sincos).  Runs ONCE per parameter upload
// (k_build_staged, from shipenv_create / shipenv_set_params); the kernels' CTAs copy the finished block.
__device__ __forceinline__ void build_staged(SharedBlock& sb, const ShipEnvParams* gp) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gp);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(&sb.p);
  constexpr int words = sizeof(ShipEnvParams) / 8;
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  for (int p = threadIdx.x; p < sb.p.n_poly; p += blockDim.x) {
    double mne = INFINITY, mxe = -INFINITY, mnn = INFINITY, mxn = -INFINITY;
#pragma unroll 1
    for (int i = sb.p.poly_start[p]; i < sb.p.poly_start[p + 1]; ++i) {
      mne = fmin(mne, sb.p.vert_e[i]); mxe = fmax(mxe, sb.p.vert_e[i]);
      mnn = fmin(mnn, sb.p.vert_n[i]); mxn = fmax(mxn, sb.p.vert_n[i]);
    }
    sb.bbox[4 * p + 0] = mne; sb.bbox[4 * p + 1] = mxe; sb.bbox[4 * p + 2] = mnn; sb.bbox[4 * p + 3] = mxn;
    for (int i = sb.p.poly_start[p]; i < sb.p.poly_start[p + 1]; ++i)
      sb.next[i] = (unsigned char)((i + 1 < sb.p.poly_start[p + 1]) ? i + 1 : sb.p.poly_start[p]);
  }
  if (threadIdx.x < 2) {
    const ShipEnvShipParams& P = sb.p.ship[threadIdx.x];
    Derived& D = sb.drv[threadIdx.x].d;
    D.los_r2 = P.los_r * P.los_r; D.los_r99 = 0.99 * P.los_r; D.los_ra2 = P.los_ra * P.los_ra;
    D.wind_cu = -0.3 * P.proj_area_f; D.wind_cv = -0.42 * P.proj_area_l; D.wind_cn = -0.096 * P.proj_area_l * P.l_ship;
    const double margin = P.l_ship / 2;
    D.hz_min_n = sb.p.map_min_n + margin; D.hz_max_n = sb.p.map_max_n - margin;
    D.hz_min_e = sb.p.map_min_e + margin; D.hz_max_e = sb.p.map_max_e - margin;
    // copies of the parameters the step reads, in its order
    D.los_limit = P.los_limit;
    D.los_ki = P.los_ki;
    D.inv_ctrl_dt = P.inv_ctrl_dt;
    D.ctrl_dt = P.ctrl_dt;
    D.hdg_kp = P.hdg_kp;
    D.hdg_kd = P.hdg_kd;
    D.hdg_ki = P.hdg_ki;
    D.max_rudder = P.max_rudder;
    D.desired_speed = P.desired_speed;
    D.spd_kp = P.spd_kp;
    D.spd_kd = P.spd_kd;
    D.spd_ki = P.spd_ki;
    D.max_thrust = P.max_thrust;
    D.dt = P.dt;
    D.cur_n = P.cur_n;
    D.cur_e = P.cur_e;
    D.c_rudder_v = P.c_rudder_v;
    D.c_rudder_r = P.c_rudder_r;
    D.cos_wind_dir = P.cos_wind_dir;
    D.sin_wind_dir = P.sin_wind_dir;
    D.wind_speed = P.wind_speed;
    D.mass = P.mass;
    D.y_dv = P.y_dv;
    D.x_du = P.x_du;
    D.lin_damp_u = P.lin_damp_u;
    D.ku = P.ku;
    D.lin_damp_v = P.lin_damp_v;
    D.kv = P.kv;
    D.lin_damp_r = P.lin_damp_r;
    D.kr = P.kr;
    D.inv_m_u = P.inv_m_u;
    D.inv_m_v = P.inv_m_v;
    D.inv_m_r = P.inv_m_r;
    D.l_ship = P.l_ship;
    D.nav_fail_tol = P.nav_fail_tol;
    D.sim_time = P.sim_time;
    D.kp_ship_speed = P.kp_ship_speed;
    D.ki_ship_speed = P.ki_ship_speed;
    D.max_shaft_speed = P.max_shaft_speed;
    D.kp_shaft_speed = P.kp_shaft_speed;
    D.ki_shaft_speed = P.ki_shaft_speed;
    D.p_me = P.p_me;
    D.p_el = P.p_el;
    D.tq_me_max = P.tq_me_max;
    D.tq_el_max = P.tq_el_max;
    D.d_me = P.d_me;
    D.r_me = P.r_me;
    D.d_hsg = P.d_hsg;
    D.r_hsg = P.r_hsg;
    D.k_torque = P.k_torque;
    D.jp = P.jp;
    D.thrust_coeff = P.thrust_coeff;
    D.dt_shaft = P.dt_shaft;
    D.k_thrust = P.k_thrust;
    D.thrust_tau = P.thrust_tau;
    if (threadIdx.x == 0) { sb.roa2 = sb.p.roa * sb.p.roa; sb.seg_len2 = sb.p.ab_segment_length * 2; }
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 50) {
    const int i = (int)threadIdx.x - 64;
    sb.sb_inv_t[i] = 1.0 / ((double)(i + 1) * 20.0);            // IEEE division: the value 1.0 / t has in the cost function
  }
  for (int i = threadIdx.x; i < 2 * SHIPENV_MAX_WP; i += blockDim.x) {
    const int r = i / SHIPENV_MAX_WP, k = i % SHIPENV_MAX_WP;
    const ShipEnvShipParams& P = sb.p.ship[r];
    if (k >= 1 && k < P.n_wp) {
      const double3 b = segment_bearing(P.wp_north[k] - P.wp_north[k - 1], P.wp_east[k] - P.wp_east[k - 1]);
      sb.seg[r][k][0] = b.x; sb.seg[r][k][1] = b.y; sb.seg[r][k][2] = b.z;
    }
  }
  __syncthreads();
}

static_assert(sizeof(SharedBlock) % 16 == 0, "the staged block is moved as one 16-byte-granular bulk copy");

__global__ void __launch_bounds__(128) k_build_staged(const ShipEnvParams* gp, SharedBlock* out) {
  __shared__ SharedBlock sb;
  build_staged(sb, gp);
  const uint4* src = reinterpret_cast<const uint4*>(&sb);
  uint4* dst = reinterpret_cast<uint4*>(out);
  for (int i = threadIdx.x; i < (int)(sizeof(SharedBlock) / 16); i += blockDim.x) dst[i] = src[i];
}

// CTA prologue of every kernel: the finished block (11.9 KB) arrives as ONE bulk copy -- thread 0 arms an mbarrier
// with the byte count and issues cp.async.bulk (the TMA engine moves the bytes), every thread waits on the barrier's
// phase.  Before, each CTA of each launch re-derived the block (64 atan2 + sincos among it): fixed work that
// dominated a one-step-per-launch kernel of 24 us.
__device__ __forceinline__ void stage_params(SharedBlock& sb, const void* staged) {
  __shared__ alignas(8) unsigned long long mbar;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(&sb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)sizeof(SharedBlock)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(staged), "r"((unsigned)sizeof(SharedBlock)), "r"(bar) : "memory");
  }
  __syncthreads();                                   // the barrier is initialised before anyone polls it
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_STAGED:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@!p bra WAIT_STAGED;\n"
      "}\n" ::"r"(bar) : "memory");
}

// trajectory log (shipenv_set_trajectory_log): next row of ship `sidx`, or nullptr when the ship is not logged
// or its log is full; log_n < 0 means "not logged"
__device__ __forceinline__ double* log_next_row(const DevView& dv, long long sidx, int& log_n) {
  if (log_n < 0) return nullptr;
  double* row = (log_n < dv.log_capacity)
                    ? dv.log_f64 + ((long long)sidx * dv.log_capacity + log_n) * SHIPENV_LOG_COLS : nullptr;
  log_n += 1;
  return row;
}

// store_last_simulation_data (ship_model.py:433-445): the previous row again, with the current time
__device__ __forceinline__ void log_repeat_row(const DevView& dv, long long sidx, int& log_n, double time) {
  if (log_n < 0) return;
  if (log_n > 0 && log_n < dv.log_capacity) {
    double* row = dv.log_f64 + ((long long)sidx * dv.log_capacity + log_n) * SHIPENV_LOG_COLS;
    const double* prev = row - SHIPENV_LOG_COLS;
    for (int c = 0; c < SHIPENV_LOG_COLS; ++c) row[c] = prev[c];
    row[SHIPENV_LOG_TIME] = time;
  }
  log_n += 1;
}

__device__ __forceinline__ int log_begin(const DevView& dv, long long env, long long sidx) {
  return (dv.log_f64 && env < dv.log_envs) ? dv.log_count[sidx] : -1;
}

// The rows of ship_f64 are walked with one pointer bumped by the row stride (two integer instructions per row);
// indexing every row as f[ROW * n_ships + sidx] cost four to five instructions of 64-bit address arithmetic per
// access, and these loads / stores run for a single lane pair of the warp when it changes environments.
static_assert(SHIPENV_SF_NORTH == 0 && SHIPENV_SF_EAST == 1 && SHIPENV_SF_YAW == 2 && SHIPENV_SF_U == 3 &&
              SHIPENV_SF_V == 4 && SHIPENV_SF_R == 5 && SHIPENV_SF_OMEGA == 6 && SHIPENV_SF_TIME == 7 &&
              SHIPENV_SF_E_CT == 8 && SHIPENV_SF_E_CT_INT == 9 && SHIPENV_SF_HDG_ERR_I == 10 &&
              SHIPENV_SF_HDG_PREV_ERR == 11 && SHIPENV_SF_SPD_ERR_I == 12 && SHIPENV_SF_SPD_AUX == 13 &&
              SHIPENV_SF_SEG_ALPHA == 14 && SHIPENV_SF_SEG_SIN == 15 && SHIPENV_SF_SEG_COS == 16,
              "load_ship / store_ship walk the rows of ship_f64 in this order");

__device__ __forceinline__ void load_ship(const DevView& dv, long long n_ships, long long sidx, Ship& s) {
  const double* p = dv.buf.ship_f64 + sidx;
  s.north = *p; p += n_ships;
  s.east = *p; p += n_ships;
  s.yaw = *p; p += n_ships;
  s.u = *p; p += n_ships;
  s.v = *p; p += n_ships;
  s.r = *p; p += n_ships;
  s.omega = *p; p += n_ships;
  s.time = *p; p += n_ships;
  s.e_ct = *p; p += n_ships;
  s.e_ct_int = *p; p += n_ships;
  s.hdg_err_i = *p; p += n_ships;
  s.hdg_prev_err = *p; p += n_ships;
  s.spd_err_i = *p; p += n_ships;
  s.spd_aux = *p; p += n_ships;
  s.seg.alpha() = *p; p += n_ships;
  s.seg.sin_a() = *p; p += n_ships;
  s.seg.cos_a() = *p;
  const int packed = dv.buf.ship_i32[sidx];
  s.k = packed & 0xff;
  s.stop = (packed >> 8) & 1;
}

__device__ __forceinline__ void store_ship(const DevView& dv, long long n_ships, long long sidx, const Ship& s) {
  double* p = dv.buf.ship_f64 + sidx;
  *p = s.north; p += n_ships;
  *p = s.east; p += n_ships;
  *p = s.yaw; p += n_ships;
  *p = s.u; p += n_ships;
  *p = s.v; p += n_ships;
  *p = s.r; p += n_ships;
  *p = s.omega; p += n_ships;
  *p = s.time; p += n_ships;
  *p = s.e_ct; p += n_ships;
  *p = s.e_ct_int; p += n_ships;
  *p = s.hdg_err_i; p += n_ships;
  *p = s.hdg_prev_err; p += n_ships;
  *p = s.spd_err_i; p += n_ships;
  *p = s.spd_aux; p += n_ships;
  *p = s.seg.alpha(); p += n_ships;
  *p = s.seg.sin_a(); p += n_ships;
  *p = s.seg.cos_a();
  dv.buf.ship_i32[sidx] = (s.k & 0xff) | (s.stop << 8);
}

__device__ __forceinline__ void init_ship_regs(const ShipEnvShipParams& P, const double* init_dev, long long n_ships,
                                               long long sidx, Ship& s) {
  // BaseShipModel.__init__ / reset (ship_model.py:100-118, 303-316), controller resets
  // (controllers.py:82-90,141-151), NavigationSystem.reset (LOS_guidance.py:129-136)
  if (init_dev) {
    s.north = init_dev[0 * n_ships + sidx]; s.east = init_dev[1 * n_ships + sidx];
    s.yaw = init_dev[2 * n_ships + sidx]; s.u = init_dev[3 * n_ships + sidx];
    s.v = init_dev[4 * n_ships + sidx]; s.r = init_dev[5 * n_ships + sidx];
    s.omega = init_dev[6 * n_ships + sidx];
  } else {
    s.north = P.init_north; s.east = P.init_east; s.yaw = P.init_yaw;
    s.u = P.init_u; s.v = P.init_v; s.r = P.init_r; s.omega = P.init_omega;
  }
  s.time = 0.0;
  s.e_ct = 0.0; s.e_ct_int = 0.0;
  s.hdg_err_i = 0.0; s.hdg_prev_err = 0.0;
  s.spd_err_i = 0.0;
  s.spd_aux = (P.model_kind == SHIPENV_MODEL_DETAILED) ? P.init_shaft_err_i : 0.0;
  s.k = 1;
  s.n_wp = P.n_wp;
  s.stop = 0;
}

// ------------------------------------------------------------------------------------------------
// construct / reset (+ init_step) kernel
// ------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(128)
k_reset(DevView dv, const uint8_t* __restrict__ mask, const double* __restrict__ init_dev, int do_init_step,
        int reinit) {
  __shared__ SharedBlock sb;
  stage_params(sb, dv.staged);
  const long long n_ships = 2 * dv.num_envs;
  const long long sidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= n_ships) return;
  const long long env = sidx >> 1;
  const int role = (int)(sidx & 1);
  if (mask && !mask[env]) return;
  const ShipEnvShipParams& P = sb.p.ship[role];
  const bool dynamic_route = (role == 1) && (sb.p.env_kind != SHIPENV_ENV_COLAV_NONIW || sb.p.obs_sampled_route != 0);
  Route rt{P.wp_north, P.wp_east, dynamic_route ? dv.buf.iw_f64 + env : nullptr,
           dynamic_route ? dv.buf.iw_f64 + (long long)SHIPENV_MAX_IW * dv.num_envs + env : nullptr, dv.num_envs, P.n_wp,
           &sb.seg[role][0][0], dynamic_route ? dv.buf.env_f64 + env : nullptr};
  Ship s;
  SENV_SEG_SLOT(s);
  int log_n = -1;
  if (reinit) {
    init_ship_regs(P, init_dev, n_ships, sidx, s);
    refresh_segment(rt, 0, s);
    if (dv.log_f64 && env < dv.log_envs) { dv.log_count[sidx] = 0; log_n = 0; }   // simulation_results = defaultdict(list)
    // obs rows <- initial_states (env.py:107-110); built from both ships of the pair
    float* orow = dv.buf.obs_f32 + env * 8;
    if (role == 0) { orow[0] = (float)s.north; orow[1] = (float)s.east; orow[2] = 0.0f; }
    else { orow[3] = (float)s.north; orow[4] = (float)s.east; orow[5] = (float)s.yaw; orow[6] = 0.0f; orow[7] = (float)s.u; }
    if (role == 1) {
      // init_get_intermediate_waypoints + results snapshot (env.py:143-184)
      double* ef = dv.buf.env_f64;
      ef[SHIPENV_EF_TRAVEL_DIST * dv.num_envs + env] = 0.0;
      ef[SHIPENV_EF_TRAVEL_TIME * dv.num_envs + env] = 0.0;
      ef[SHIPENV_EF_ACC_REWARD * dv.num_envs + env] = 0.0;
      ef[SHIPENV_EF_N_BASE * dv.num_envs + env] = sb.p.n_base0;
      ef[SHIPENV_EF_E_BASE * dv.num_envs + env] = sb.p.e_base0;
      ef[SHIPENV_EF_LOG_NORTH * dv.num_envs + env] = s.north;
      ef[SHIPENV_EF_LOG_EAST * dv.num_envs + env] = s.east;
      if (!do_init_step) {
        // Env.__init__ creates the SBMPC object (env.py:123); reset() never touches its memory
        ef[SHIPENV_EF_SB_P_LAST * dv.num_envs + env] = 1.0;
        ef[SHIPENV_EF_SB_CHI_LAST * dv.num_envs + env] = 0.0;
      }
      int* ei = dv.buf.env_i32;
      ei[SHIPENV_EI_SAMPLING_COUNT * dv.num_envs + env] = 0;
      ei[SHIPENV_EI_SNAPSHOT_INFO * dv.num_envs + env] = 0;
      ei[SHIPENV_EI_FLAGS * dv.num_envs + env] = 0;
      dv.buf.reward[env] = 0.0;
      dv.buf.info_i32[env] = 0;
      dv.buf.nsub_i32[env] = 0;
    }
  } else {
    load_ship(dv, n_ships, sidx, s);
    const int n_iw = dynamic_route ? dv.buf.env_i32[SHIPENV_EI_SAMPLING_COUNT * dv.num_envs + env] : 0;
    s.n_wp = P.n_wp + n_iw;
    load_segment_points(rt, n_iw, s);
    log_n = log_begin(dv, env, sidx);
  }
  if (do_init_step) {
    // init_step (env.py:297-342): controllers, log row, one integration step, tracker on
    if (role == 1) {
      dv.buf.env_f64[SHIPENV_EF_LOG_NORTH * dv.num_envs + env] = s.north;
      dv.buf.env_f64[SHIPENV_EF_LOG_EAST * dv.num_envs + env] = s.east;
      dv.buf.env_i32[SHIPENV_EI_FLAGS * dv.num_envs + env] |= SHIPENV_FLAG_TRACKER;
    }
    ship_step<MODEL>(P, rt, s.n_wp - P.n_wp, s, false, 0.0, -0.0, 1.0, log_next_row(dv, sidx, log_n));
    if (log_n >= 0) dv.log_count[sidx] = log_n;
  }
  store_ship(dv, n_ships, sidx, s);
}

// self.states is float32 and is only (re)initialised by __init__ (env.py:110): separate kernel so
// that reset() leaves it alone.
__global__ void k_init_prev_states(DevView dv) {
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= dv.num_envs) return;
  const float* orow = dv.buf.obs_f32 + env * 8;
  dv.buf.prev_f32[0 * dv.num_envs + env] = orow[0];
  dv.buf.prev_f32[1 * dv.num_envs + env] = orow[1];
  dv.buf.prev_f32[2 * dv.num_envs + env] = orow[3];
  dv.buf.prev_f32[3 * dv.num_envs + env] = orow[4];
}

// ------------------------------------------------------------------------------------------------
// step(action) prologue, one thread per environment (rl_env env.py:641-696, run_colav env.py:1430-1474):
// obs_ship_uses_scoping_angle -> get_intermediate_waypoints -> update_route -> sampling-failure test.
// Runs before k_env<MODE_STEP> on the same stream, at full lane utilisation (inside the persistent env
// kernel this work ran for one lane pair of a warp at a time).  An environment whose sampled waypoint fails
// the test is finished here (it returns its results snapshot unchanged); for the others the obstacle ship's
// segment cache is refreshed when the insertion changed the waypoint it is heading for.
// ------------------------------------------------------------------------------------------------
template <int ENVKIND>
__global__ void __launch_bounds__(128)
k_prologue(DevView dv, const double* __restrict__ actions, unsigned long long* __restrict__ queue) {
  // the work-queue counters of the k_env launch that follows on the same stream restart at 0
  if (blockIdx.x == 0 && threadIdx.x < 2) queue[threadIdx.x] = 0ull;
  // The parameter block is read where it lies (a dozen scalars per thread, the same addresses for every thread:
  // L1 / L2 hits): staging 7 KB into shared memory per 128 environments cost more than the prologue's own work.
  const ShipEnvParams& G = *dv.params;
  const long long B = dv.num_envs;
  const long long env = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  constexpr bool IS_RL = ENVKIND == SHIPENV_ENV_RL;
  int* ei = dv.buf.env_i32;
  double* ef = dv.buf.env_f64;
  int flags = (env < B) ? ei[SHIPENV_EI_FLAGS * B + env] : SHIPENV_FLAG_DONE;
  int sampling_count = (env < B) ? ei[SHIPENV_EI_SAMPLING_COUNT * B + env] : -1;
  if (dv.call_filter != 0) {
    // queue[2] (zeroed by the host): environments the second launch of this call will take (SenvView::call_filter) --
    // those whose sampling count, after this prologue, is the maximum; either launch returns at once when it has none
    const bool live = (env < B) && !(flags & SHIPENV_FLAG_DONE);
    const int after = sampling_count + ((live && sampling_count < G.max_sampling_frequency) ? 1 : 0);
    const unsigned last = __ballot_sync(FULL_MASK, env < B && after == G.max_sampling_frequency);
    if ((threadIdx.x & 31) == 0 && last) atomicAdd(&queue[2], (unsigned long long)__popc(last));
  }
  if (env >= B) return;
  if (flags & SHIPENV_FLAG_DONE) return;                      // k_env reports it as already done
  flags &= ~SHIPENV_FLAG_HAVE_IW;
  if (sampling_count < G.max_sampling_frequency) {
    const MapView mp{G.vert_e, G.vert_n, G.poly_start, nullptr, nullptr, G.n_poly, dv.grid};
    const double a = actions[env];
    sampling_count += 1;
    // get_intermediate_waypoints (env.py:198-236)
    const double n_base = ef[SHIPENV_EF_N_BASE * B + env], e_base = ef[SHIPENV_EF_E_BASE * B + env];
    const double l_s = fabs(G.ab_segment_length * tan(a));
    double e_s = l_s * G.cos_omega;
    double n_s = l_s * G.sin_omega;
    if (a > 0) e_s *= -1; else n_s *= -1;
    const double rn = n_base + n_s, re = e_base + e_s;
    ef[SHIPENV_EF_N_BASE * B + env] = rn + G.ab_north_segment_length;
    ef[SHIPENV_EF_E_BASE * B + env] = re + G.ab_east_segment_length;
    // update_route: insert before the last waypoint (controllers.py:417-422)
    dv.buf.iw_f64[(long long)(sampling_count - 1) * B + env] = rn;
    dv.buf.iw_f64[((long long)SHIPENV_MAX_IW + sampling_count - 1) * B + env] = re;
    ei[SHIPENV_EI_SAMPLING_COUNT * B + env] = sampling_count;
    ef[SHIPENV_EF_TRAVEL_DIST * B + env] = 0.0;
    ef[SHIPENV_EF_TRAVEL_TIME * B + env] = 0.0;
    flags |= SHIPENV_FLAG_HAVE_IW;
    // is_route_inside_obstacles / is_route_outside_horizon (check_condition.py:80-119)
    const bool fail = map_contains(mp, map_cell_masks(mp, rn, re) & 0xffffu, rn, re) ||
                      ((rn < G.map_min_n || rn > G.map_max_n) || (re < G.map_min_e || re > G.map_max_e));
    if (fail) {
      double out_reward = 0.0;
      if (IS_RL) {
        const double acc = ef[SHIPENV_EF_ACC_REWARD * B + env];    // reward_function.py:499-527, multiplier 2
        out_reward = (acc >= 0) ? (-acc * 2.0) : (acc * 2.0);
      }
      const int snapshot = (ei[SHIPENV_EI_SNAPSHOT_INFO * B + env] & 0x7ff) | SHIPENV_EV_SAMPLING_FAILURE |
                           SHIPENV_INFO_TERMINAL;
      ei[SHIPENV_EI_SNAPSHOT_INFO * B + env] = snapshot;
      dv.buf.info_i32[env] = snapshot | SHIPENV_INFO_DONE;
      dv.buf.reward[env] = out_reward;
      dv.buf.nsub_i32[env] = 0;
      flags |= SHIPENV_FLAG_DONE;
      if (dv.buf.counters) atomicAdd(&dv.buf.counters[1], 1ull);
    } else {
      if (IS_RL) ef[SHIPENV_EF_ACC_REWARD * B + env] = 0.0;      // rl_env env.py:696
      // The new waypoint took the route's second-to-last slot: bearings of the two segments it creates (previous
      // waypoint -> new waypoint -> route end), for the autopilot's next waypoint switches.
      const ShipEnvShipParams& P = G.ship[1];
      const long long n_ships = 2 * B, sidx = 2 * env + 1;
      const int head = P.n_wp - 1;
      const int j = sampling_count - 1;                              // index of the new waypoint among the sampled ones
      double pn, pe;
      if (j == 0) { pn = P.wp_north[head - 1]; pe = P.wp_east[head - 1]; }
      else {
        pn = dv.buf.iw_f64[(long long)(j - 1) * B + env];
        pe = dv.buf.iw_f64[((long long)SHIPENV_MAX_IW + j - 1) * B + env];
      }
      const double3 bn = segment_bearing(rn - pn, re - pe);         // LOS_guidance.py:105-107
      const double3 be = segment_bearing(P.wp_north[head] - rn, P.wp_east[head] - re);
      ef[SHIPENV_EF_SEG_NEW_ALPHA * B + env] = bn.x; ef[SHIPENV_EF_SEG_NEW_SIN * B + env] = bn.y;
      ef[SHIPENV_EF_SEG_NEW_COS * B + env] = bn.z;
      ef[SHIPENV_EF_SEG_END_ALPHA * B + env] = be.x; ef[SHIPENV_EF_SEG_END_SIN * B + env] = be.y;
      ef[SHIPENV_EF_SEG_END_COS * B + env] = be.z;
      // If the obstacle ship is already heading for the route's last waypoint (index head + j), that index now holds
      // the new waypoint: its current segment becomes previous -> new.
      const int k = dv.buf.ship_i32[sidx] & 0xff;
      if (k == head + j) {
        double* f = dv.buf.ship_f64;
        f[SHIPENV_SF_SEG_ALPHA * n_ships + sidx] = bn.x;
        f[SHIPENV_SF_SEG_SIN * n_ships + sidx] = bn.y;
        f[SHIPENV_SF_SEG_COS * n_ships + sidx] = bn.z;
      }
    }
  }
  ei[SHIPENV_EI_FLAGS * B + env] = flags;
}

// ------------------------------------------------------------------------------------------------
// the env kernel: step(action) [MODE_STEP] or k x _step() [MODE_SUBSTEPS]
//
// Persistent grid with lane-pair refill: the grid is sized to what is resident on the GPU; every lane
// pair first takes the environment with its own slot index and, whenever its environment has finished
// the call (reached the next radius of acceptance, terminated, was already done, ...), stores it and
// pulls the next environment index from a device-wide counter (one warp-aggregated atomicAdd).  A warp
// therefore keeps all 16 pairs busy until the queue is empty instead of waiting for its slowest
// environment, and environments that are already done cost one fetch.
// ------------------------------------------------------------------------------------------------
enum LaneState { LS_FETCH = 0, LS_LOAD = 1, LS_RUN = 2, LS_IDLE = 3 };


#ifndef SENV_MIN_BLOCKS_SBMPC
#define SENV_MIN_BLOCKS_SBMPC 4   // measured: 3 CTAs/SM (168 registers, fewer spills) speeds the inactive steps up but slows the evaluation
#endif
template <int MODEL, int ENVKIND, int MODE, int COLLAV, int QUIET_STEPS = 1>
#ifdef SENV_MAXNREG
__global__ void __maxnreg__(SENV_MAXNREG)
#else
__global__ void __launch_bounds__(SENV_ENV_BLOCK, COLLAV == SHIPENV_COLLAV_SBMPC ? SENV_MIN_BLOCKS_SBMPC
                                  : ((SENV_QUIET != 0) && (QUIET_STEPS != 0) && COLLAV == SHIPENV_COLLAV_NONE && ENVKIND == SHIPENV_ENV_COLAV_IW)
                                        ? SENV_MIN_BLOCKS_QUIET : SENV_MIN_BLOCKS)
#endif
k_env(DevView dv, const double* __restrict__ actions, int k_substeps, unsigned long long* __restrict__ queue) {
  constexpr bool SBMPC = COLLAV == SHIPENV_COLLAV_SBMPC;
  constexpr bool SIMPLE = COLLAV == SHIPENV_COLLAV_SIMPLE;
  // (the collision-avoidance mode is a template parameter: its branches would otherwise split the simulator step's
  //  basic block in every instantiation, see ship_step)
  // Without SBMPC the waypoint switch of step t + 1 (NavigationSystem.next_wpt at the top of the autopilot call) is
  // decided right after the integration of step t and carried out in the event path below; SBMPC's extra
  // los_guidance call runs on the segment BEFORE the switch (quirk 2), so those instantiations keep it at the top.
  constexpr bool EARLY_SWITCH = !SBMPC;
  if (MODE == MODE_STEP && dv.call_filter != 0) {
    // one of the two launches of a step(action) call (SenvView::call_filter): nothing to do when the other has it all
    const unsigned long long n_last = queue[2];
    if (dv.call_filter == 2 ? (n_last == 0ull) : (n_last == (unsigned long long)dv.num_envs)) return;
  }
  // work-queue index: queue[0]; the second launch of a split call draws from queue[3] (both zeroed before the first)
  unsigned long long* const work_q = queue + ((MODE == MODE_STEP && dv.call_filter == 2) ? 3 : 0);
  __shared__ SharedBlock sb_static;
  stage_params(sb_static, dv.staged);
  const int lane = (int)(threadIdx.x & 31);
  const int role_tid = (int)(threadIdx.x & 1);
  // The shared-memory addresses of the parameter block and of this lane's ship parameters are held in two
  // registers the compiler cannot rematerialise: left alone it recomputed them five times per simulator step
  // (S2R SR_CgaCtaId + S2R SR_TID.X + LEA + LOP3 + IMAD each time: 30 of the loop's 545 instructions, with the
  // S2R latency in front of the parameter loads that follow).  Measured +5.7 % (colav_iw) / +8.5 % (rl).  Doing the
  // same to `role` costs more in spills than the S2R + LOP3 it saves (measured: no gain).
  unsigned sb_addr = (unsigned)__cvta_generic_to_shared(&sb_static);
  unsigned p_addr = (unsigned)__cvta_generic_to_shared(&sb_static.p.ship[role_tid]);
  asm volatile("" : "+r"(sb_addr), "+r"(p_addr));
  // the role, where it is needed again, from those two registers (a subtraction and a compare) rather than from
  // S2R SR_TID.X, whose latency sat in front of every role-dependent select
  const int role = (p_addr - sb_addr) != (unsigned)offsetof(SharedBlock, p.ship[0]) ? 1 : 0;
  SharedBlock& sb = *reinterpret_cast<SharedBlock*>(__cvta_shared_to_generic(sb_addr));
  const ShipEnvParams& G = sb.p;
  const ShipEnvShipParams& P = *reinterpret_cast<const ShipEnvShipParams*>(__cvta_shared_to_generic(p_addr));
  const long long B = dv.num_envs;
  const long long n_ships = 2 * B;
  const long long n_slots = ((long long)gridDim.x * blockDim.x) >> 1;
  constexpr bool IS_RL = ENVKIND == SHIPENV_ENV_RL;
  constexpr bool IS_IW = ENVKIND != SHIPENV_ENV_COLAV_NONIW;
  // Quiet steps (see the simulator loop): instantiations without collision avoidance and without a per-step reward
  // (QUIET_STEPS = 0: the instantiation that evaluates every test at every step -- launches of a few steps, where the
  //  limits would be taken and never used, and SHIPENV_QUIET=0)
  constexpr bool QUIET = (SENV_QUIET != 0) && (QUIET_STEPS != 0) && COLLAV == SHIPENV_COLLAV_NONE && IS_IW && !IS_RL;
  // (the NonIW env samples intermediate waypoints too when it is driven with step(action), run_colav/env.py:678-800)
  const bool dynamic_route = (IS_IW || G.obs_sampled_route != 0) && role == 1;
  const MapView mp{G.vert_e, G.vert_n, G.poly_start, sb.bbox, sb.next, G.n_poly, dv.grid};
  const bool has_stop_branch = (role == 1) || !IS_RL;          // rl_env test_step has none (env.py:345-445)
  const bool collav_lane = SIMPLE && (role == 0 || !IS_IW);
  const double collav_bias = IS_RL ? (-15.0 * (kPi / 180.0)) : (15.0 * (kPi / 180.0));
  const double route_end_n = P.wp_north[P.n_wp - 1], route_end_e = P.wp_east[P.n_wp - 1];

  long long env = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  int lstate = (env < B) ? LS_LOAD : LS_IDLE;
  // QUIET: distance this ship has travelled since it was loaded (`odo`), the reading at which one of the tests on its
  // own position could change its outcome at the earliest (`lim`: map, route end, travel tracker, see own_budget; < 0:
  // to be taken anew by the full evaluation), and the reading at which the pair could collide at the earliest
  double odo = 0.0, lim = (env < B) ? -1.0 : 1e30, coll_lim = 1e30;
  // QUIET: squared distance to this ship's next waypoint at or below which a waypoint test can fire (-1: none applies)
  double wp_thr2 = -1.0;

  // registers of the environment currently held by this lane
  Ship s = Ship{};
  SENV_SEG_SLOT(s);
  s.k = 1;
  Route rt{P.wp_north, P.wp_east, nullptr, nullptr, B, P.n_wp, &sb.seg[role][0][0], nullptr};
  double travel_dist = 0.0, travel_time = 0.0, acc_reward = 0.0;
  // The travel tracker adds |row_t - row_{t-1}| of the obstacle ship's logged (pre-integration) positions
  // (env.py:526-534).  That distance is known one step early -- right after the integration of step t-1, as
  // |position after - position before| (the same subtraction) -- so it is carried as one value (`pending_dist`)
  // instead of the previous position, which would have to be rotated through registers at every step.  The
  // previous position itself (LOG_NORTH / LOG_EAST rows of env_f64) and the surge speed before the integration
  // (obs[6]) are only needed when the environment is stored: they are written to a per-lane shared-memory slot.
  // ([field][thread]: a warp's 8-byte stores of one field are conflict-free)
  __shared__ double lane_scratch[3][SENV_ENV_BLOCK];
  unsigned scratch_addr = (unsigned)__cvta_generic_to_shared(&lane_scratch[0][threadIdx.x]);
  asm volatile("" : "+r"(scratch_addr));
  double* const scratch_base = reinterpret_cast<double*>(__cvta_shared_to_generic(scratch_addr));
  struct { double& log_n; double& log_e; double& u_pre; } scratch{scratch_base[0], scratch_base[SENV_ENV_BLOCK], scratch_base[2 * SENV_ENV_BLOCK]};
  double pending_dist = 0.0;
  // rl_env, fast build: sin / cos of this ship's heading, carried from step to step (see ship_step's yaw_sc)
  constexpr bool CARRY_YAW_SC = IS_RL && (SENV_FAST_MATH != 0);
  double2 yaw_sc = make_double2(0.0, 1.0);
  int tlog_n = -1;             // rows in this ship's trajectory log (-1: not logged)
  double sb_p_last = 1.0, sb_chi_last = 0.0;   // SBMPCParams.P_ca_last_ / Chi_ca_last_ (both lanes of the pair)
  int sampling_count = 0, flags = 0, n_iw = 0;
  float ps_tn = 0.f, ps_te = 0.f, ps_on = 0.f, ps_oe = 0.f;
  bool last_stop_branch = false;
  double out_reward = 0.0;
  int out_info = 0, nsub = 0;
  bool have_obs = false;      // next_observations was assigned by this call
  int stage = 0;              // MODE_STEP: 0 main loop, 1 extra step after RoA, 2 run to completion
  bool have_iw = false;
  int k_left = (QUIET && !(env < B)) ? (1 << 30) : 0;   // (QUIET, MODE_STEP: simulator steps until the time limit can bind)
  // metric counters over every environment this lane handled
  int total_sub = 0, total_fin = 0, total_dead = 0;
#ifdef SENV_QUIET_STATS
  int qs_total = 0, qs_quiet = 0;
#endif
  // QUIET: distance this ship may still travel from where it is before one of the tests on its OWN position can change
  // its outcome: grounding (safe radius of the culling-grid cell), route end (max(|dn|, |de|) <= distance) and, for the
  // obstacle ship, the travel tracker (which lags the position by the step just taken).  A stopped ship's bits stay as
  // they are.  (The map horizon is not among them: ships sail along it for hundreds of steps, so the quiet steps test
  // it as it stands.)
  // (single precision: the margin of 2 m covers its rounding at map scale many times over; a NaN position makes the
  // odometer NaN, which no limit passes)
  auto own_budget = [&](double north, double east) -> float {
    if (has_stop_branch && s.stop) return 1e30f;
    const float n = (float)north, e = (float)east;
    float b = map_safe_radius(mp, n, e);
    b = fminf(b, fmaxf(fabsf(n - (float)route_end_n), fabsf(e - (float)route_end_e)) - 200.0f);
    if (role == 1) b = fminf(b, (float)(sb.seg_len2 - travel_dist - pending_dist));
    return 0.999f * b - 2.0f;
  };
  // QUIET: everything a lane holds for the quiet steps, taken anew from the state after a step (or after the load);
  // (pn, pe) is where the partner ship is.  Every term is a lower bound, 1-Lipschitz in the ship's position, of the
  // distance to a threshold, less a margin that covers the rounding of the running sum.
  auto take_limit = [&](double pn, double pe) {
    const Derived& D = derived_of(P);
    const bool stopped = has_stop_branch && s.stop;           // this ship no longer moves: its own bits stay as they are
    // the waypoint tests of the quiet steps: `<= ra^2` for the autopilot's switch, `< roa^2` for the env's radius of
    // acceptance (taken as `<=`: a step with equality is evaluated in full)
    wp_thr2 = -1.0;
    if (!stopped) {
      if (s.n_wp > s.k + 1) wp_thr2 = D.los_ra2;
      if (MODE == MODE_STEP && role == 1 && stage == 0) wp_thr2 = fmax(wp_thr2, sb.roa2);
    }
    // collision (d < 50): either ship may close half the gap; max(|dn|, |de|) <= d
    const float gap = fmaxf(fabsf((float)(pn - s.north)), fabsf((float)(pe - s.east)));
    coll_lim = odo + (double)(0.4995f * (gap - 50.0f) - 2.0f);
    lim = odo + (double)own_budget(s.north, s.east);
    // simulation time limit on the test ship's clock (a stopped ship's clock advances twice per step)
    int cap = 1 << 30;
    if (role == 0) {
      const float left = __fdividef((float)(D.sim_time - s.time), (float)(2.0 * D.dt)) - 4.0f;
      cap = (left > 0.0f) ? ((left < 1e9f) ? (int)left : (1 << 30)) : 0;
    }
    if (MODE == MODE_STEP) {
      k_left = cap;
    } else if (cap < k_left) {
      lim = -1.0;
    }
  };

  for (;;) {
    // ---------------- (1) fetch: lane pairs without an environment pull the next index
    {
      unsigned want = __ballot_sync(FULL_MASK, lstate == LS_FETCH && role == 0);
      if (SENV_REFILL_MIN > 1) {   // optional batching of refills
        const unsigned busy = __ballot_sync(FULL_MASK, lstate == LS_RUN || lstate == LS_LOAD);
        if (__popc(want) < SENV_REFILL_MIN && busy != 0) want = 0;
      }
      if (want) {
        const int leader = __ffs(want) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(work_q, (unsigned long long)__popc(want));
        base = __shfl_sync(FULL_MASK, base, leader);
        long long mine = n_slots + (long long)base + __popc(want & ((1u << lane) - 1u));
        mine = __shfl_sync(FULL_MASK, mine, lane & ~1);       // the obstacle lane takes its partner's index
        if (lstate == LS_FETCH) {
          env = mine;
          lstate = (env < B) ? LS_LOAD : LS_IDLE;
          if (QUIET && lstate == LS_IDLE) { lim = 1e30; coll_lim = 1e30; odo = 0.0; k_left = 1 << 30; wp_thr2 = -1.0; s.e_ct = 0.0; }   // an idle lane never vetoes a quiet step
        }
      }
    }
    if (!__any_sync(FULL_MASK, lstate != LS_IDLE)) break;

    // ---------------- (2) load the environment and run the step(action) prologue
    bool finalize = false;                // store this environment at the end of the iteration
    if (MODE == MODE_STEP && dv.call_filter != 0 && lstate == LS_LOAD) {
      // this launch takes only one kind of environment (SenvView::call_filter); the other launch of the call has the rest
      const bool last_call = dv.buf.env_i32[env] == G.max_sampling_frequency;     // row SHIPENV_EI_SAMPLING_COUNT
      if ((dv.call_filter == 2) != last_call) lstate = LS_FETCH;
    }
    if (lstate == LS_LOAD) {
      const long long sidx = 2 * env + role;
      load_ship(dv, n_ships, sidx, s);
      {
        // rows of env_f64 / env_i32 with a bumped pointer (see load_ship)
        static_assert(SHIPENV_EF_TRAVEL_DIST == 0 && SHIPENV_EF_TRAVEL_TIME == 1 && SHIPENV_EF_ACC_REWARD == 2 &&
                      SHIPENV_EF_LOG_NORTH == 5 && SHIPENV_EF_LOG_EAST == 6 && SHIPENV_EF_SB_P_LAST == 7 &&
                      SHIPENV_EF_SB_CHI_LAST == 8 && SHIPENV_EI_SAMPLING_COUNT == 0 && SHIPENV_EI_FLAGS == 2,
                      "row order of env_f64 / env_i32");
        const double* ef = dv.buf.env_f64 + env;
        travel_dist = *ef; ef += B;
        travel_time = *ef; ef += B;
        acc_reward = *ef; ef += 3 * B;
        const double log_n = *ef; ef += B;
        const double log_e = *ef; ef += B;
        scratch.log_n = log_n; scratch.log_e = log_e; scratch.u_pre = 0.0;
        const double tn = s.north - log_n, te = s.east - log_e;
        pending_dist = SENV_SQRT(tn * tn + te * te);
        if (SBMPC) {
          sb_p_last = *ef; ef += B;
          sb_chi_last = *ef;
        }
        const int* ei = dv.buf.env_i32 + env;
        sampling_count = *ei; ei += 2 * B;
        flags = *ei;
      }
      if (SIMPLE) {
        ps_tn = dv.buf.prev_f32[0 * B + env]; ps_te = dv.buf.prev_f32[1 * B + env];
        ps_on = dv.buf.prev_f32[2 * B + env]; ps_oe = dv.buf.prev_f32[3 * B + env];
      }
      rt.iw_n = dynamic_route ? dv.buf.iw_f64 + env : nullptr;
      rt.iw_e = dynamic_route ? dv.buf.iw_f64 + (long long)SHIPENV_MAX_IW * B + env : nullptr;
      rt.seg_env = dynamic_route ? dv.buf.env_f64 + env : nullptr;
      n_iw = dynamic_route ? sampling_count : 0;
      s.n_wp = P.n_wp + n_iw;
      last_stop_branch = false; out_reward = 0.0; out_info = 0; nsub = 0;
      have_obs = false; stage = 0; have_iw = false; k_left = k_substeps;
      odo = 0.0; lim = -1.0;
      lstate = LS_RUN;
      if (flags & SHIPENV_FLAG_DONE) {
        // already done before this call: nothing changes except the step count of the call
        if (role == 1) { dv.buf.nsub_i32[env] = 0; total_dead += 1; }
        lstate = LS_FETCH;
      } else if (MODE == MODE_SUBSTEPS && k_left <= 0) {
        if (role == 1) dv.buf.nsub_i32[env] = 0;
        lstate = LS_FETCH;
      } else if (MODE == MODE_STEP) {
        // the step(action) prologue (action -> intermediate waypoint, route insertion, sampling-failure test)
        // already ran for every environment in k_prologue
        have_iw = (flags & SHIPENV_FLAG_HAVE_IW) != 0;
      }
      if (lstate == LS_RUN) {
        load_segment_points(rt, n_iw, s);
        // the waypoint switch the first step of this call starts with (the route may have grown since the store)
        if (EARLY_SWITCH && !(has_stop_branch && s.stop) && wpt_reached(derived_of(P), s)) {
          s.k += 1;
          refresh_segment(rt, n_iw, s);
        }
        tlog_n = log_begin(dv, env, sidx);
        if (CARRY_YAW_SC) senv_sincos(s.yaw, &yaw_sc.x, &yaw_sc.y);
        if (QUIET) {
          // the limits of the quiet steps, so that the first step after the load can already be one (the two lanes of
          // the pair are here together)
          const unsigned pair = 3u << (lane & ~1);
          const double q_n = __shfl_xor_sync(pair, s.north, 1), q_e = __shfl_xor_sync(pair, s.east, 1);
          take_limit(q_n, q_e);
        }
      }
    }

    // ---------------- (3) simulator steps of the running lanes, until some pair has finished its call
    do {
    quiet_next:;
    const bool running = (lstate == LS_RUN) && !finalize;
    // QUIET: does this lane pass the coming step as a quiet one (see below); lanes that do not move keep their reading
    if (QUIET) k_left -= 1;
    bool lane_quiet = QUIET && (odo < lim) && (k_left > 0);
    // The assets step in list order (test_step, then obs_step).  The lanes of a pair run them side by
    // side, except under SBMPC in the NonIW env, where obs_step's SBMPC call reads the state the ship under
    // test has just integrated to (run_colav env.py:502-527): two phases there.
    constexpr int N_PHASE = (SBMPC && !IS_IW) ? 2 : 1;
    unsigned cell = 0;          // culling-grid masks of the cell the ship is in after this step (set by every running lane)
#pragma unroll 1
    for (int phase = 0; phase < N_PHASE; ++phase) {
    const bool stepping = running && (N_PHASE == 1 || role == phase);
    const bool stop_branch = stepping && has_stop_branch && s.stop;
    double heading_offset = -0.0, speed_factor = 1.0;
    if (SBMPC) {
      // rl_env env.py:360-385, run_colav env.py:1145-1170 (and :371-396, :502-527 in the NonIW env)
      const int caller_role = (N_PHASE == 1) ? 0 : phase;
      const bool caller = stepping && !stop_branch && role == caller_role;
      SbmpcIn in;
      in.chi_d = 0.0;
      // next_wpt()'s result is discarded and los_guidance runs on the autopilot's current waypoint index:
      // the LOS integrator advances a first time here (quirk 2 of SURVEY.md section 8)
      if (caller) in.chi_d = -los_guidance(P, s);
      const double q_east = shfl_xor_f64(s.east, 1), q_north = shfl_xor_f64(s.north, 1);
      // own ship = assets[0] (ship under test), obstacle = assets[1], whichever asset calls
      in.os_x = role == 0 ? s.east : q_east; in.os_y = role == 0 ? s.north : q_north;
      in.ob_x = role == 1 ? s.east : q_east; in.ob_y = role == 1 ? s.north : q_north;
      // D_INIT_ test (sbmpc.py:154-166); both lanes of the pair see the same two positions, so the lane that
      // does not call resets its copy of the SBMPC memory alongside the caller
      const bool pair_calls = __shfl_sync(FULL_MASK, (int)caller, (lane & ~1) | caller_role) != 0;
      bool active = false;
      if (pair_calls) {
        const double d0 = in.ob_x - in.os_x, d1 = in.ob_y - in.os_y;
        active = SENV_SQRT(d0 * d0 + d1 * d1) < 2000.0;
        if (!active) { sb_p_last = 1.0; sb_chi_last = 0.0; }
      }
      unsigned todo = __ballot_sync(FULL_MASK, caller && active);
      if (todo) {                                                // warp-uniform: most steps have no active pair
        const double q_yaw = shfl_xor_f64(s.yaw, 1), q_u = shfl_xor_f64(s.u, 1), q_v = shfl_xor_f64(s.v, 1);
        in.os_v = role == 0 ? s.v : q_v;
        in.ob_psi = -(role == 1 ? s.yaw : q_yaw);
        in.ob_u = role == 1 ? s.u : q_u; in.ob_v = role == 1 ? s.v : q_v;
        in.u_d = P.desired_speed;
        in.chi_last = sb_chi_last; in.p_last = sb_p_last;
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int best = sbmpc_warp_argmin(in, src, lane, G.ship[1].l_ship, G.ship[1].w_ship, sb.sb_inv_t);
          if (lane == src) {
            double u_best = 1.0, chi_best = 0.0;
            if (best >= 0) {
              const int ic = best >> 2, jp = best & 3;
              chi_best = (-30.0 + 10.0 * (double)ic) * (kPi / 180.0);
              u_best = (jp == 0) ? 0.4 : ((jp == 1) ? 0.6 : ((jp == 2) ? 0.8 : 1.0));
            }
            sb_p_last = u_best; sb_chi_last = chi_best;
            speed_factor = u_best; heading_offset = -chi_best;
          }
        }
        // the pair shares one SBMPC object: the other lane takes over the caller's copy of its memory
        sb_p_last = __shfl_sync(FULL_MASK, sb_p_last, (lane & ~1) | caller_role);
        sb_chi_last = __shfl_sync(FULL_MASK, sb_chi_last, (lane & ~1) | caller_role);
      }
    }
    if (stepping) {
      const double dt = derived_of(P).dt;
      if (stop_branch) {
        // stopped ship: log row repeated, clock advanced twice (env.py:451-479)
        log_repeat_row(dv, 2 * env + role, tlog_n, s.time);
        s.time = s.time + dt;
        s.time = s.time + dt;
        last_stop_branch = true;
        if (!QUIET) cell = map_cell_masks(mp, s.north, s.east);
      } else {
        scratch.u_pre = s.u;
        last_stop_branch = false;
        bool hit = false;
        if (collav_lane) {
          // is_collision_imminent on the float32 self.states (check_condition.py:130-140)
          const float dn = ps_tn - ps_on, de = ps_te - ps_oe;
          hit = (dn * dn + de * de) < 9000000.0f;
        }
        const double pre_n = s.north, pre_e = s.east;
        ship_step<MODEL, !EARLY_SWITCH, SIMPLE>(P, rt, n_iw, s, hit, collav_bias, heading_offset, speed_factor,
                                                log_next_row(dv, 2 * env + role, tlog_n),
                                                [&](double new_n, double new_e) {
                                                  if (!QUIET) { cell = map_cell_masks(mp, new_n, new_e); return; }
                                                  // QUIET: the travel tracker (below) and this lane's part of the
                                                  // quiet-step test, evaluated as soon as the kinematics have the new
                                                  // position -- beside the step's long chains instead of behind them
                                                  const bool track = (role == 1) && (flags & SHIPENV_FLAG_TRACKER);
                                                  travel_dist += track ? pending_dist : 0.0;
                                                  travel_time += track ? dt : 0.0;
                                                  const double tn = new_n - s.north, te = new_e - s.east;
                                                  pending_dist = SENV_SQRT(tn * tn + te * te);
                                                  odo += pending_dist;
                                                  const double qn = new_n - s.seg.wn(), qe = new_e - s.seg.we();
                                                  const Derived& D = derived_of(P);
                                                  const bool outside = ((int)(new_n < D.hz_min_n) | (int)(new_n > D.hz_max_n) |
                                                                        (int)(new_e < D.hz_min_e) | (int)(new_e > D.hz_max_e)) != 0;
                                                  // (combined without short-circuit branches: one basic block)
                                                  lane_quiet = ((int)(k_left > 0) & (int)!(qn * qn + qe * qe <= wp_thr2) & (int)!outside) != 0;
                                                },
                                                CARRY_YAW_SC ? &yaw_sc : nullptr);
        if (QUIET) {
          // a lane that reaches its limit first takes its own part anew where it is now (a few instructions, no
          // exchange with the partner lane): only a ship that really is next to one of its thresholds asks for the
          // full evaluation.  (Behind the step, not inside it: a branch would split the step's basic block.)
          if (!(odo < lim) && lim >= 0.0) lim = odo + (double)own_budget(s.north, s.east);
          lane_quiet = lane_quiet && (odo < lim) && !(fabs(s.e_ct) > derived_of(P).nav_fail_tol);
        }
        if (IS_IW && !QUIET) {
          // travel tracker on the two last logged rows of the obstacle ship (env.py:526-534); evaluated without a
          // branch on the role (the test lane's copies are never read: a divergent branch would cost the same issue
          // slots and end the step's basic block)
          const bool track = (role == 1) && (flags & SHIPENV_FLAG_TRACKER);
          travel_dist += track ? pending_dist : 0.0;
          travel_time += track ? dt : 0.0;
          const double tn = s.north - pre_n, te = s.east - pre_e;
          pending_dist = SENV_SQRT(tn * tn + te * te);
        }
        scratch.log_n = pre_n; scratch.log_e = pre_e;
      }
    }
    }
    if (QUIET) {
      // ---- quiet steps.  Every event test below is a predicate on where the two ships are (grounding, map horizon,
      // route end, radius of acceptance, waypoint switch, collision, cross-track error, travelled distance) or on the
      // clock.  The two that fire routinely are evaluated here as they stand: the distance to the ship's next waypoint
      // against the larger of the radii that apply to it (`wp_thr2`: radius of acceptance of the env, of the autopilot)
      // and the cross-track error against the navigation-failure tolerance.  For all the others each lane took, at its
      // last full evaluation, the smallest distance its ship would have to travel before one of ITS predicates could
      // flip (`lim`, see own_budget and the end of the loop body) and the number of steps before the time limit can bind
      // (`k_left`).  While every lane of the warp passes after the step it just took, all the tests below are known to
      // come out as they did -- "nothing holds" or, for a stopped ship, the same bits as before -- and the bookkeeping
      // they drive changes nothing, so the warp goes straight to the next step.  Idle lanes hold an infinite limit.
      // (MODE_SUBSTEPS: k_left is the launch's own step counter, so the last step of a launch is always evaluated in
      // full.)
#ifdef SENV_QUIET_STATS
      qs_total += 1;
#endif
      if (__all_sync(FULL_MASK, lane_quiet && (odo < coll_lim))) {
        nsub += 1;
#ifdef SENV_QUIET_STATS
        qs_quiet += 1;
#endif
        goto quiet_next;
      }
      if (__all_sync(FULL_MASK, lane_quiet)) {
        // only collision limits ran out (two ships passing each other stay a few steps apart for a long time): the
        // collision test as it stands -- same expressions as below -- and new limits from the gap that is left
        const double c_north = shfl_xor_f64(s.north, 1), c_east = shfl_xor_f64(s.east, 1);
        const double cx = c_north - s.north, cy = c_east - s.east;
        const bool hit = running && (cx * cx + cy * cy < 2500.0);
        if (running) coll_lim = odo + (double)(0.4995f * (fmaxf(fabsf((float)cx), fabsf((float)cy)) - 50.0f) - 2.0f);
        if (!__any_sync(FULL_MASK, hit)) {
          nsub += 1;
#ifdef SENV_QUIET_STATS
          qs_quiet += 1;
#endif
          goto quiet_next;
        }
      }
    }
    // ---- the ships meet: positions both ways, then one word of partial flags per ship.  Almost every
    // simulator step is "plain" (no termination / stop condition holds for either ship, no collision, no
    // radius of acceptance reached): those steps skip the event bookkeeping below altogether.  (The instantiations
    // with quiet steps, above, come here only when some lane of the warp cannot rule its tests out.)
    const double p_north = shfl_xor_f64(s.north, 1);
    const double p_east = shfl_xor_f64(s.east, 1);
    // bits of my_flags: 1 grounding, 2 navigation failure, 4 reached the route end, 8 outside the map horizon,
    // 16 simulation time limit (set by the test lane), 32 radius of acceptance reached (set by the obstacle lane)
    int my_flags = 0;
    double ra = 0.0, rb = 0.0;
    if (running) {
      const Derived& D = derived_of(P);
      bool nav_fail = fabs(s.e_ct) > derived_of(P).nav_fail_tol;
      // QUIET: a moving ship still inside its limit is known to fail every test on its own position (that is what the
      // limit bounds), so a full evaluation that another lane of the warp asked for skips them -- and the culling-grid
      // load in front of them -- for this lane
      const bool known_clear = QUIET && (odo < lim) && !(has_stop_branch && s.stop);
      // (four comparisons combined without short-circuit branches)
      const bool outside = ((int)(s.north < D.hz_min_n) | (int)(s.north > D.hz_max_n) |
                            (int)(s.east < D.hz_min_e) | (int)(s.east > D.hz_max_e)) != 0;
      if (!known_clear) {
        if (QUIET) cell = map_cell_masks(mp, s.north, s.east);
        const double len = derived_of(P).l_ship;
        const bool grounding = pos_inside_obstacles(mp, cell & 0xffffu, s.north, s.east, len);
        const double dn = s.north - route_end_n, de = s.east - route_end_e;
        // is_reaches_endpoint: sqrt(d2) <= 200  <=>  d2 <= 40000 exactly (sqrt is correctly rounded and
        // sqrt(nextafter(40000)) rounds above 200)
        const bool reached = (dn * dn + de * de) <= 40000.0;
        if (role == 1) nav_fail = (travel_dist > sb.seg_len2) || (travel_time > INFINITY) || nav_fail;
        my_flags = (grounding ? 1 : 0) | (reached ? 4 : 0);
      }
      my_flags |= (nav_fail ? 2 : 0) | (outside ? 8 : 0);
      // is_within_simu_time_limit on the test ship's clock (check_condition.py:206-213)
      if (role == 0 && s.time > derived_of(P).sim_time) my_flags |= 16;
      if (MODE == MODE_STEP && role == 1 && stage == 0) {
        // is_reach_radius_of_acceptance on the obstacle ship's next waypoint (check_condition.py:181-204)
        const double rn = s.north - s.seg.wn(), re = s.east - s.seg.we();
        if ((rn * rn + re * re) < sb.roa2) my_flags |= 32;
      }
      // bit 64: this ship's autopilot will switch to its next waypoint at the top of the next step
      if (EARLY_SWITCH && wpt_reached(D, s)) my_flags |= 64;
      if (IS_RL) {
        const double gd = map_distance(mp, s.north, s.east);
        const double aect = fabs(s.e_ct);
        // test_ship_grounding_reward / test_ship_nav_failure_reward (reward_function.py:359-425) on the test lane,
        // obs_ship_grounding_reward / obs_ship_nav_failure_reward (:427-494, negated) on the obstacle lane:
        // RewardDesign3/4 with per-role constants, one exp() call site for both roles
        const double g_scale = role == 0 ? 175000.0 : 50000.0;
        const double n_tol = role == 0 ? 3000.0 : 500.0, n_scale = role == 0 ? 1250000.0 : 12500.0;
        if (gd <= 1000.0) ra = (gd < 0.0) ? 1.0 : senv_exp(SENV_DIV(-(gd * gd), g_scale));
        rb = (aect < n_tol) ? senv_exp(SENV_DIV(-((aect - n_tol) * (aect - n_tol)), n_scale)) : 1.0;
        if (role == 1) { ra = -ra; rb = -rb; }
      }
    }
    const int p_flags = __shfl_xor_sync(FULL_MASK, my_flags, 1);
    const int p_stop = __shfl_xor_sync(FULL_MASK, s.stop, 1);
    double p_ra = 0.0, p_rb = 0.0;
    if (IS_RL) { p_ra = shfl_xor_f64(ra, 1); p_rb = shfl_xor_f64(rb, 1); }

    if (running) {
      nsub += 1;
      const int t_flags = role == 0 ? my_flags : p_flags, o_flags = role == 1 ? my_flags : p_flags;
      if (SIMPLE) {   // self.states = next_states (float32)
        const double t_n = role == 0 ? s.north : p_north, t_e = role == 0 ? s.east : p_east;
        const double o_n = role == 1 ? s.north : p_north, o_e = role == 1 ? s.east : p_east;
        ps_tn = (float)t_n; ps_te = (float)t_e; ps_on = (float)o_n; ps_oe = (float)o_e;
      }
      // ship-ship terms (compute_distance.py:16-40, check_condition.py:142-158); (a - b)^2 == (b - a)^2
      // exactly, so both lanes get the same d2
      const double dx = p_north - s.north, dy = p_east - s.east;       // obs - test on the test lane
      const double d2 = dx * dx + dy * dy;
      const bool is_collision = d2 < 2500.0;
      // The AST reward is accumulated on the TEST lane only (role 0: it has the test ship's heading for the
      // encounter type); the obstacle lane's r_total / acc_reward are never read (it fetches them from its
      // partner when the environment is stored).
      double r_total = 0.0;
      if (IS_RL) {
        const double distance = SENV_SQRT(d2);
        double r1 = 0.0;
        if (distance < 10000.0) {
#if SENV_FAST_MATH
          // encounter type (reward_function.py:82-112): beta = wrap(atan2(dy, dx) - psi) lies in [-pi, pi], and the
          // only class that changes the reward is |beta| > 165 deg, i.e. cos(beta) < cos(165 deg) with
          // cos(beta) = (dx cos psi + dy sin psi) / distance: one sincos of the heading instead of atan2 + a modulo.
          // Same decision as the angle form up to the rounding of either (beta exactly at 165 deg has measure zero).
          double s_yaw, c_yaw;
          if (CARRY_YAW_SC) { s_yaw = yaw_sc.x; c_yaw = yaw_sc.y; }   // of the heading after this step (ship_step)
          else senv_sincos(s.yaw, &s_yaw, &c_yaw);
          const bool overtaking = (dx * c_yaw + dy * s_yaw) < -0.9659258262890682 * distance;   // cos(165.0 * (pi / 180))
#else
          const double phi = senv_atan2(dy, dx);
          double beta = phi - s.yaw;
          beta = py_mod(beta + kPi, 2 * kPi) - kPi;
          const bool overtaking = !(fabs(beta) < 15.0 * (kPi / 180.0)) && (fabs(beta) > 165.0 * (kPi / 180.0));
#endif
          // head-on or crossing -> RewardDesign4(target 0, 2e8); the "overtake" branch is dead code
          if (!overtaking) r1 = (distance < 0.0) ? 1.0 : senv_exp(SENV_DIV(-(distance * distance), 200000000.0));
        }
        r_total = SENV_DIV((((r1 + ra) + rb) + p_ra) + p_rb, 5.0);               // test terms, then obstacle terms
      }
      bool st_done = false, st_terminal = false;
      out_info = 0;
      // One test for the common simulator step of a step(action) call: main loop (stage 0), no event bit of either
      // ship, radius of acceptance not reached, no collision and (run_colav, whose `done` needs both stop flags) no
      // stop flag set.  Nothing below changes anything then except the reward accumulators.
      const bool plain_step = (MODE == MODE_STEP) && (stage == 0) && (((t_flags | o_flags) & 127) == 0) &&
                              !is_collision && (IS_RL || ((s.stop | p_stop) == 0));
      if (plain_step) {
        if (IS_RL) { acc_reward += r_total; out_reward = r_total; }
      } else {
      if (((t_flags | o_flags) & 31) != 0 || is_collision) {
        // ---- something holds: termination reward, events and env_info (reward_function.py:204-314)
        const bool t_ground = t_flags & 1, t_nav = t_flags & 2, o_ground = o_flags & 1, o_nav = o_flags & 2;
        if (IS_RL) {
          // get_reward_due_to_ships_termination (reward_function.py:272-314)
          if (is_collision || t_ground || t_nav || o_ground || o_nav) {
            const double reward = r_total + acc_reward;
            r_total = 0;
            if (acc_reward > 0) {
              if (is_collision) r_total += reward * 10.0;
              if (t_ground) r_total += reward * 5.0;
              if (t_nav) r_total += reward * 5.0;
              if (o_ground) r_total += reward * -2.5;
              if (o_nav) r_total += reward * -2.5;
            } else if (acc_reward < 0) {
              if (is_collision) r_total += reward * -10.0;
              if (t_ground) r_total += reward * -5.0;
              if (t_nav) r_total += reward * -5.0;
              if (o_ground) r_total += reward * 2.5;
              if (o_nav) r_total += reward * 2.5;
            }
          }
        }
        int ev = 0;
        bool terminal = false, ts = false, os = false;
        if (is_collision) { ev |= SHIPENV_EV_COLLISION; terminal = ts = os = true; }
        if (t_ground) { ev |= SHIPENV_EV_TEST_GROUNDING; terminal = ts = true; }
        if (t_nav) { ev |= SHIPENV_EV_TEST_NAV_FAILURE; terminal = ts = true; }
        if (o_ground) { ev |= SHIPENV_EV_OBS_GROUNDING; terminal = os = true; }
        if (o_nav) { ev |= SHIPENV_EV_OBS_NAV_FAILURE; terminal = os = true; }
        if (t_flags & 4) { ev |= SHIPENV_EV_TEST_REACHED; ts = true; }
        if (t_flags & 8) { ev |= SHIPENV_EV_TEST_OUTSIDE; ts = true; }
        if (o_flags & 4) { ev |= SHIPENV_EV_OBS_REACHED; os = true; }
        if (o_flags & 8) { ev |= SHIPENV_EV_OBS_OUTSIDE; os = true; }
        if (t_flags & 16) { ev |= SHIPENV_EV_TIME_LIMIT; ts = os = true; }
        st_terminal = terminal;
        int partner_stop = p_stop;
        if (IS_RL) {                                              // rl_env env.py:603-610
          st_done = ts && !terminal;
          if (role == 1 && os && !terminal) s.stop = 1;
        } else {                                                  // run_colav env.py:1385-1399
          const int stop_before = s.stop;
          if (ts && !terminal) { if (role == 0) s.stop = 1; else partner_stop = 1; }
          if (os && !terminal) { if (role == 1) s.stop = 1; else partner_stop = 1; }
          if (QUIET && s.stop != stop_before) lim = -1.0;         // this ship stops moving: its limit is taken anew
          st_done = s.stop && partner_stop;                       // done needs both stop flags
        }
        out_info = ev | (terminal ? SHIPENV_INFO_TERMINAL : 0) | (ts ? SHIPENV_INFO_TEST_STOP : 0) |
                   (os ? SHIPENV_INFO_OBS_STOP : 0);
      } else if (!IS_RL) {
        st_done = s.stop && p_stop;
      }
      if (IS_RL) {
        if (MODE == MODE_STEP) acc_reward += r_total;
        out_reward = r_total;
      }
      const bool combined_done = st_terminal || st_done;
      if (combined_done) out_info |= SHIPENV_INFO_DONE;
      if (MODE == MODE_SUBSTEPS) {
        have_obs = true;
        if (!QUIET) k_left -= 1;
        if (combined_done) { flags |= SHIPENV_FLAG_DONE; finalize = true; }
        else if (k_left <= 0) finalize = true;
      } else {
        // step() control flow: rl_env env.py:700-771, run_colav env.py:1478-1533
        if (stage == 0) {
          if (combined_done) { have_obs = true; flags |= SHIPENV_FLAG_DONE; finalize = true; }
          else if (o_flags & 32) {
            // the obstacle ship is inside the radius of acceptance of its next waypoint
            if (have_iw) { stage = 1; if (QUIET) lim = -1.0; }
            else { out_info |= SHIPENV_INFO_UNBOUND | SHIPENV_INFO_DONE; flags |= SHIPENV_FLAG_DONE; finalize = true; }
          }
        } else if (stage == 1) {
          have_obs = true;
          if (combined_done) { flags |= SHIPENV_FLAG_DONE; finalize = true; }
          else if (sampling_count == G.max_sampling_frequency) { travel_dist = 0.0; travel_time = 0.0; stage = 2; if (QUIET) lim = -1.0; }
          else finalize = true;
        } else {
          have_obs = true;
          if (combined_done) { flags |= SHIPENV_FLAG_DONE; finalize = true; }
        }
      }
      // NavigationSystem.next_wpt of the next step (LOS_guidance.py:83-98), unless this ship stops stepping (a
      // stopped ship's autopilot is not called again) or the environment is stored (the next call's route may differ:
      // the test is repeated when it is loaded)
      if (EARLY_SWITCH && (my_flags & 64) && !finalize && !(has_stop_branch && s.stop)) {
        s.k += 1;
        refresh_segment(rt, n_iw, s);
        if (QUIET) lim = -1.0;                                 // new segment, new next waypoint
      }
      }   // !plain_step
      if (QUIET && !finalize && !(MODE == MODE_STEP && stage == 1) &&
          !((odo < lim) && (odo < coll_lim) && (k_left > 0 || MODE == MODE_SUBSTEPS))) {
        // ---- a lane without a valid limit takes everything anew from the state after this step
        take_limit(p_north, p_east);
      }
    }
    } while (!__any_sync(FULL_MASK, finalize || lstate == LS_FETCH));

    // ---------------- (4) store finished environments and free the lane pair
    {
      // this lane's entries of next_states (test_step / obs_step return values, env.py:440-443, 472-477,
      // 519-524): stop branch and the obstacle ship of the IW envs return [N, E, psi, u, e_ct], the others
      // [N, E, e_ct]
      const float o0 = (float)s.north, o1 = (float)s.east;
      float o2, o3 = 0.f, o4 = 0.f;
      if (last_stop_branch) { o2 = (float)s.yaw; o3 = 0.0f; o4 = (float)s.e_ct; }
      else if (role == 1 && IS_IW) { o2 = (float)s.yaw; o3 = (float)scratch.u_pre; o4 = (float)s.e_ct; }
      else { o2 = (float)s.e_ct; }
      const float t0 = __shfl_xor_sync(FULL_MASK, o0, 1);
      const float t1 = __shfl_xor_sync(FULL_MASK, o1, 1);
      const float t2 = __shfl_xor_sync(FULL_MASK, o2, 1);
      if (IS_RL) {
        // the reward accumulators live on the test lane; the obstacle lane writes the environment
        const double acc_t = shfl_xor_f64(acc_reward, 1), out_t = shfl_xor_f64(out_reward, 1);
        if (role == 1) { acc_reward = acc_t; out_reward = out_t; }
      }
      if (finalize) {
        store_ship(dv, n_ships, 2 * env + role, s);
        if (tlog_n >= 0) dv.log_count[2 * env + role] = tlog_n;
        if (role == 1) {
          int* ei = dv.buf.env_i32;
          if (have_obs) {
            float4* orow = reinterpret_cast<float4*>(dv.buf.obs_f32 + env * 8);
            orow[0] = make_float4(t0, t1, t2, o0);
            orow[1] = IS_IW ? make_float4(o1, o2, o3, o4) : make_float4(o1, o2, 0.f, 0.f);
            ei[SHIPENV_EI_SNAPSHOT_INFO * B + env] = out_info & ~SHIPENV_INFO_DONE;     // results snapshot
          }
          double* ef = dv.buf.env_f64 + env;
          *ef = travel_dist; ef += B;
          *ef = travel_time; ef += B;
          *ef = acc_reward; ef += 3 * B;
          *ef = scratch.log_n; ef += B;
          *ef = scratch.log_e; ef += B;
          if (SBMPC) {
            *ef = sb_p_last; ef += B;
            *ef = sb_chi_last;
          }
          ei[SHIPENV_EI_FLAGS * B + env] = flags;
          if (SIMPLE) {
            dv.buf.prev_f32[0 * B + env] = ps_tn; dv.buf.prev_f32[1 * B + env] = ps_te;
            dv.buf.prev_f32[2 * B + env] = ps_on; dv.buf.prev_f32[3 * B + env] = ps_oe;
          }
          if (IS_RL && MODE == MODE_STEP) out_reward = acc_reward;
          dv.buf.reward[env] = out_reward;
          dv.buf.info_i32[env] = out_info;
          dv.buf.nsub_i32[env] = nsub;
          total_sub += nsub;
          if (flags & SHIPENV_FLAG_DONE) total_fin += 1;
        }
        lstate = LS_FETCH;
      }
    }
  }

#ifdef SENV_QUIET_STATS
  // measurement build: warp iterations of the simulator loop and how many of them were quiet (spare counters 2, 3)
  if (lane == 0 && dv.buf.counters) {
    atomicAdd(&dv.buf.counters[2], (unsigned long long)qs_total);
    atomicAdd(&dv.buf.counters[3], (unsigned long long)qs_quiet);
  }
#endif
  // metric counters: one atomic per warp
  {
    const int sub = __reduce_add_sync(FULL_MASK, total_sub);
    const int fin = __reduce_add_sync(FULL_MASK, total_fin);
    const int dead = __reduce_add_sync(FULL_MASK, total_dead);
    if (lane == 0) {
      if (dv.buf.counters) {
        if (sub) atomicAdd(&dv.buf.counters[0], (unsigned long long)sub);
        if (fin) atomicAdd(&dv.buf.counters[1], (unsigned long long)fin);
      }
      if (dead + fin) atomicAdd(&queue[1], (unsigned long long)(dead + fin));   // done after this launch
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bare ship loop (no env logic): every lane integrates its own ship k steps
// ------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(128)
k_ship_rollout(DevView dv, int k_steps) {
  __shared__ SharedBlock sb;
  stage_params(sb, dv.staged);
  const long long n_ships = 2 * dv.num_envs;
  const long long sidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= n_ships) return;
  const int role = (int)(sidx & 1);
  const ShipEnvShipParams& P = sb.p.ship[role];
  const Route rt{P.wp_north, P.wp_east, nullptr, nullptr, dv.num_envs, P.n_wp, &sb.seg[role][0][0], nullptr};
  Ship s;
  SENV_SEG_SLOT(s);
  load_ship(dv, n_ships, sidx, s);
  s.n_wp = P.n_wp;
  load_segment_points(rt, 0, s);
  int log_n = log_begin(dv, sidx >> 1, sidx);
  for (int i = 0; i < k_steps; ++i) ship_step<MODEL>(P, rt, 0, s, false, 0.0, -0.0, 1.0, log_next_row(dv, sidx, log_n));
  if (log_n >= 0) dv.log_count[sidx] = log_n;
  store_ship(dv, n_ships, sidx, s);
  if (dv.buf.counters) atomicAdd(&dv.buf.counters[2], (unsigned long long)k_steps);
}

// ------------------------------------------------------------------------------------------------
// map geometry probe (shipenv_map_query): the env kernel's own geometry routines on caller-given points, so that
// tests can pin them against exact arithmetic (tests/test_map_geometry.py).  One thread per point.
//   contains[i]   PolygonObstacle.if_pos_inside_obstacles(north, east) (obstacle.py:126-129), as step()'s prologue
//                 tests a sampled waypoint
//   square[i]     is_pos_inside_obstacles (check_condition.py:48-78): any corner of the ship_length square inside
//   distance[i]   PolygonObstacle.obstacles_distance (obstacle.py:138-141) where it is <= 1000 m (the reward's clip);
//                 beyond the clip the value is only guaranteed to be > 1000 (possibly inf), see map_distance
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_map_query(DevView dv, long long n, const double* __restrict__ north, const double* __restrict__ east,
            double ship_length, int* __restrict__ contains, int* __restrict__ square, double* __restrict__ distance) {
  __shared__ SharedBlock sb;
  stage_params(sb, dv.staged);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ShipEnvParams& G = sb.p;
  const MapView mp{G.vert_e, G.vert_n, G.poly_start, sb.bbox, sb.next, G.n_poly, dv.grid};
  const double pn = north[i], pe = east[i];
  const unsigned cell = map_cell_masks(mp, pn, pe);
  contains[i] = map_contains(mp, cell & 0xffffu, pn, pe) ? 1 : 0;
  square[i] = pos_inside_obstacles(mp, cell & 0xffffu, pn, pe, ship_length) ? 1 : 0;
  distance[i] = map_distance(mp, pn, pe);
}

// the quiet steps' safe radius (map_safe_radius) at caller-given points: test probe
__global__ void __launch_bounds__(128)
k_map_safe_radius(DevView dv, long long n, const double* __restrict__ north, const double* __restrict__ east,
                  float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const MapView mp{nullptr, nullptr, nullptr, nullptr, nullptr, 0, dv.grid};
  out[i] = map_safe_radius(mp, (float)north[i], (float)east[i]);
}

#ifndef SENV_ONLY_ONE
// ------------------------------------------------------------------------------------------------
// launch wrappers (declared in shipenv_launch.h)
// ------------------------------------------------------------------------------------------------
constexpr int kBlock = 128;
constexpr int kEnvBlock = SENV_ENV_BLOCK;

inline int ship_grid(const DevView& v) { return (int)((2 * v.num_envs + kBlock - 1) / kBlock); }

cudaError_t launch_reset(const SenvView& v, int model, const uint8_t* mask, const double* init, int do_init,
                         int reinit, cudaStream_t st) {
  if (model == SHIPENV_MODEL_SIMPLE)
    k_reset<SHIPENV_MODEL_SIMPLE><<<ship_grid(v), kBlock, 0, st>>>(v, mask, init, do_init, reinit);
  else if (model == SHIPENV_MODEL_DETAILED)
    k_reset<SHIPENV_MODEL_DETAILED><<<ship_grid(v), kBlock, 0, st>>>(v, mask, init, do_init, reinit);
  else
    k_reset<SHIPENV_MODEL_SIMPLIFIED><<<ship_grid(v), kBlock, 0, st>>>(v, mask, init, do_init, reinit);
  return cudaGetLastError();
}

cudaError_t launch_init_prev(const SenvView& v, cudaStream_t st) {
  k_init_prev_states<<<(int)((v.num_envs + 255) / 256), 256, 0, st>>>(v);
  return cudaGetLastError();
}

// persistent grid: as many CTAs as are resident at once (queried per instantiation), never more than
// the environments need
template <int MODEL, int ENVKIND, int MODE, int COLLAV, int QUIET_STEPS = 1>
static void launch_env_inst2(const SenvView& v, const double* actions, int k, unsigned long long* queue,
                             int sm_count, int persistent, cudaStream_t st) {
  // the instantiations with quiet steps have a twin without them (see k_env)
  constexpr bool kHasQuiet = (SENV_QUIET != 0) && COLLAV == SHIPENV_COLLAV_NONE && ENVKIND == SHIPENV_ENV_COLAV_IW;
  if (kHasQuiet && QUIET_STEPS != 0 && v.no_quiet) {
    launch_env_inst2<MODEL, ENVKIND, MODE, COLLAV, kHasQuiet ? 0 : 1>(v, actions, k, queue, sm_count, persistent, st);
    return;
  }
  static int per_sm = 0;
  if (per_sm == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_env<MODEL, ENVKIND, MODE, COLLAV, QUIET_STEPS>, kEnvBlock, 0) != cudaSuccess || n < 1)
      n = SENV_MIN_BLOCKS;
    per_sm = n;
  }
  const long long need = (2 * v.num_envs + kEnvBlock - 1) / kEnvBlock;
  const long long resident = (long long)per_sm * (sm_count > 0 ? sm_count : 148);
  // persistent: resident CTAs only, lane pairs refill from the queue; otherwise one slot per environment
  const int grid = (int)((persistent && resident < need) ? resident : need);
  k_env<MODEL, ENVKIND, MODE, COLLAV, QUIET_STEPS><<<grid, kEnvBlock, 0, st>>>(v, actions, k, queue);
}

template <int MODEL, int ENVKIND, int MODE>
static void launch_env_inst(const SenvView& v, const double* actions, int k, unsigned long long* queue,
                            int sm_count, int persistent, cudaStream_t st) {
  if (v.collav == SHIPENV_COLLAV_SBMPC)
    launch_env_inst2<MODEL, ENVKIND, MODE, SHIPENV_COLLAV_SBMPC>(v, actions, k, queue, sm_count, persistent, st);
  else if (v.collav == SHIPENV_COLLAV_SIMPLE)
    launch_env_inst2<MODEL, ENVKIND, MODE, SHIPENV_COLLAV_SIMPLE>(v, actions, k, queue, sm_count, persistent, st);
  else
    launch_env_inst2<MODEL, ENVKIND, MODE, SHIPENV_COLLAV_NONE>(v, actions, k, queue, sm_count, persistent, st);
}

template <int MODEL, int MODE>
static void launch_env_kind(const SenvView& v, int env_kind, const double* actions, int k, unsigned long long* queue,
                            int sm_count, int persistent, cudaStream_t st) {
  switch (env_kind) {
    case SHIPENV_ENV_COLAV_NONIW:
      launch_env_inst<MODEL, SHIPENV_ENV_COLAV_NONIW, MODE>(v, actions, k, queue, sm_count, persistent, st);
      break;
    case SHIPENV_ENV_COLAV_IW:
      launch_env_inst<MODEL, SHIPENV_ENV_COLAV_IW, MODE>(v, actions, k, queue, sm_count, persistent, st);
      break;
    default:
      launch_env_inst<MODEL, SHIPENV_ENV_RL, MODE>(v, actions, k, queue, sm_count, persistent, st);
      break;
  }
}

cudaError_t launch_prologue(const SenvView& v, int env_kind, const double* actions, unsigned long long* queue,
                            cudaStream_t st) {
  const int grid = (int)((v.num_envs + kBlock - 1) / kBlock);
  if (env_kind == SHIPENV_ENV_RL) k_prologue<SHIPENV_ENV_RL><<<grid, kBlock, 0, st>>>(v, actions, queue);
  else k_prologue<SHIPENV_ENV_COLAV_IW><<<grid, kBlock, 0, st>>>(v, actions, queue);
  return cudaGetLastError();
}

cudaError_t launch_env(const SenvView& v, int model, int env_kind, int mode, const double* actions, int k,
                       unsigned long long* queue, int sm_count, int persistent, int clear_queue, cudaStream_t st) {
  // queue[0] = work-queue counter, queue[1] = environments found already done by this launch; both
  // restart at 0 for every launch (ordered on the same stream)
  if (clear_queue) {
    cudaError_t e = cudaMemsetAsync(queue, 0, 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
  }
  if (model == SHIPENV_MODEL_SIMPLE) {
    if (mode == MODE_STEP) launch_env_kind<SHIPENV_MODEL_SIMPLE, MODE_STEP>(v, env_kind, actions, k, queue, sm_count, persistent, st);
    else launch_env_kind<SHIPENV_MODEL_SIMPLE, MODE_SUBSTEPS>(v, env_kind, actions, k, queue, sm_count, persistent, st);
  } else if (model == SHIPENV_MODEL_DETAILED) {
    if (mode == MODE_STEP) launch_env_kind<SHIPENV_MODEL_DETAILED, MODE_STEP>(v, env_kind, actions, k, queue, sm_count, persistent, st);
    else launch_env_kind<SHIPENV_MODEL_DETAILED, MODE_SUBSTEPS>(v, env_kind, actions, k, queue, sm_count, persistent, st);
  } else {
    if (mode == MODE_STEP) launch_env_kind<SHIPENV_MODEL_SIMPLIFIED, MODE_STEP>(v, env_kind, actions, k, queue, sm_count, persistent, st);
    else launch_env_kind<SHIPENV_MODEL_SIMPLIFIED, MODE_SUBSTEPS>(v, env_kind, actions, k, queue, sm_count, persistent, st);
  }
  return cudaGetLastError();
}

cudaError_t launch_math_selftest(long long n, unsigned long long seed, unsigned long long* mismatches_dev,
                                 cudaStream_t st) {
  k_math_selftest<<<(int)((n + 255) / 256), 256, 0, st>>>(n, seed, mismatches_dev);
  return cudaGetLastError();
}

size_t staged_bytes() { return sizeof(SharedBlock); }

cudaError_t launch_build_staged(const ShipEnvParams* params_dev, void* staged_dev, cudaStream_t st) {
  k_build_staged<<<1, 128, 0, st>>>(params_dev, reinterpret_cast<SharedBlock*>(staged_dev));
  return cudaGetLastError();
}

cudaError_t launch_map_query(const SenvView& v, long long n, const double* north, const double* east,
                             double ship_length, int* contains, int* square, double* distance, cudaStream_t st) {
  k_map_query<<<(int)((n + kBlock - 1) / kBlock), kBlock, 0, st>>>(v, n, north, east, ship_length, contains, square,
                                                                   distance);
  return cudaGetLastError();
}

cudaError_t launch_map_safe_radius(const SenvView& v, long long n, const double* north, const double* east, float* out,
                                   cudaStream_t st) {
  k_map_safe_radius<<<(int)((n + kBlock - 1) / kBlock), kBlock, 0, st>>>(v, n, north, east, out);
  return cudaGetLastError();
}

cudaError_t launch_rollout(const SenvView& v, int model, int k, cudaStream_t st) {
  if (model == SHIPENV_MODEL_SIMPLE) k_ship_rollout<SHIPENV_MODEL_SIMPLE><<<ship_grid(v), kBlock, 0, st>>>(v, k);
  else if (model == SHIPENV_MODEL_DETAILED) k_ship_rollout<SHIPENV_MODEL_DETAILED><<<ship_grid(v), kBlock, 0, st>>>(v, k);
  else k_ship_rollout<SHIPENV_MODEL_SIMPLIFIED><<<ship_grid(v), kBlock, 0, st>>>(v, k);
  return cudaGetLastError();
}

#else
template __global__ void k_env<0, 1, 0, 0, 1>(DevView, const double*, int, unsigned long long*);
template __global__ void k_env<1, 2, 0, 0, 1>(DevView, const double*, int, unsigned long long*);
#endif
}  // namespace SENV_NS
