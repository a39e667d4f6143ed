/*
 * shipsim_oracle.c -- CPU restatement (plain C, scalar, FP64) of the reference simulator step.
 *
 * TEST INFRASTRUCTURE ONLY (see shipsim_oracle.h).  Build: make -C oracle  (gcc -O2
 * -ffp-contract=off, no fast-math: every operation is an individually rounded IEEE double
 * operation, as in CPython / NumPy scalar arithmetic).
 *
 * The 3x3 matrix products of the reference (np.dot / np.linalg.inv, ship_model.py:367-378) are
 * written out in closed form using x_g = 0 (ship_model.py:77), R^-1 = R^T and the diagonal mass
 * matrix; SURVEY.md Appendix A.  Everything else follows the reference statement by statement;
 * each function cites the lines it restates (paths relative to /root/reference).
 */
#define _DEFAULT_SOURCE
#include "shipsim_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_PI 3.141592653589793

int orc_sizeof_ship_config(void) { return (int)sizeof(OrcShipConfig); }
int orc_sizeof_env_config(void) { return (int)sizeof(OrcEnvConfig); }
int orc_sizeof_ship_state(void) { return (int)sizeof(OrcShipState); }
int orc_sizeof_env_state(void) { return (int)sizeof(OrcEnvState); }
int orc_sizeof_step_result(void) { return (int)sizeof(OrcStepResult); }

/* ---------------------------------------------------------------------------------------------
 * derived constants: BaseShipModel.__init__, run_colav ship_model.py:70-132 (= rl_env :412-474)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  double mass, i_z, x_du, y_dv, n_dr, l_ship;
  double proj_area_f, proj_area_l;
  double p_me, p_el, tq_me_max, tq_el_max, dp4;
} Derived;

static void derive(const OrcShipConfig* c, Derived* d) {
  double payload = 0.9 * (c->dead_weight_tonnage - c->bunkers);
  double lsw = c->dead_weight_tonnage / c->coefficient_of_deadweight_to_displacement - c->dead_weight_tonnage;
  d->mass = lsw + payload + c->bunkers + c->ballast;
  d->l_ship = c->length_of_ship;
  d->i_z = d->mass * (c->length_of_ship * c->length_of_ship + c->width_of_ship * c->width_of_ship) / 12;
  d->x_du = d->mass * c->added_mass_coefficient_in_surge;       /* set_added_mass :137-153 */
  d->y_dv = d->mass * c->added_mass_coefficient_in_sway;
  d->n_dr = d->i_z * c->added_mass_coefficient_in_yaw;
  d->proj_area_f = c->width_of_ship * 8.0;                      /* h_f = h_s = 8.0 :126-129 */
  d->proj_area_l = c->length_of_ship * 8.0;
  /* MachineryMode.update_available_propulsion_power, ship_engine.py:32-44 */
  if (c->shaft_generator_state == ORC_HSG_MOTOR) {
    d->p_me = c->main_engine_capacity;
    d->p_el = c->electrical_capacity - c->hotel_load;
  } else if (c->shaft_generator_state == ORC_HSG_GEN) {
    d->p_me = c->main_engine_capacity - c->hotel_load;
    d->p_el = 0;
  } else {
    d->p_me = c->main_engine_capacity;
    d->p_el = 0;
  }
  d->tq_me_max = d->p_me / 5 * ORC_PI / 30;                     /* ship_engine.py:423,432 */
  d->tq_el_max = d->p_el / 5 * ORC_PI / 30;
  d->dp4 = pow(c->propeller_diameter, 4.0);                     /* self.dp ** 4, ship_engine.py:414 */
}

static double sat(double val, double low, double hi) {          /* controllers.py:67-72: max(low, min(val, hi)) */
  double m = (hi < val) ? hi : val;                             /* Python min(val, hi) */
  return (m > low) ? m : low;                                   /* Python max(low, m)  */
}

/* ---------------------------------------------------------------------------------------------
 * construction / reset of one asset
 * ------------------------------------------------------------------------------------------- */
void orc_ship_init(const OrcShipConfig* c, OrcShipState* s) {
  memset(s, 0, sizeof(*s));
  s->north = c->initial_north_position_m;                       /* ship_model.py:100-105 */
  s->east = c->initial_east_position_m;
  s->yaw = c->initial_yaw_angle_rad;
  s->u = c->initial_forward_speed_m_per_s;
  s->v = c->initial_sideways_speed_m_per_s;
  s->r = c->initial_yaw_rate_rad_per_s;
  s->omega = c->initial_propeller_shaft_speed_rad_per_s;        /* ship_engine.py:370 */
  s->time = 0.0;                                                /* EulerInt.__init__, utils.py:20-25 */
  s->shaft_err_i = c->initial_shaft_speed_integral_error;       /* rl_env controllers.py:175-180 */
  s->n_wp = c->n_wp;                                            /* load_waypoints, LOS_guidance.py:58-81 */
  for (int i = 0; i < c->n_wp; ++i) { s->wp_north[i] = c->wp_north[i]; s->wp_east[i] = c->wp_east[i]; }
  s->next_wpt = 1;                                              /* controllers.py:405-406 */
  s->prev_wpt = 0;
}

/* ---------------------------------------------------------------------------------------------
 * NavigationSystem.next_wpt + los_guidance, LOS_guidance.py:83-117
 * ------------------------------------------------------------------------------------------- */
/* NavigationSystem.los_guidance(k, N, E), LOS_guidance.py:100-117 */
static double los_guidance(const OrcShipConfig* c, OrcShipState* s, int k);

static double los_heading_ref(const OrcShipConfig* c, OrcShipState* s) {
  int k = s->next_wpt;
  double N = s->north, E = s->east;
  double dn = s->wp_north[k] - N, de = s->wp_east[k] - E;
  if (dn * dn + de * de <= c->radius_of_acceptance * c->radius_of_acceptance) {
    if (s->n_wp > k + 1) { s->next_wpt = k + 1; s->prev_wpt = k; }
    else { s->next_wpt = k; s->prev_wpt = k; }
  } else { s->next_wpt = k; s->prev_wpt = k - 1; }
  return los_guidance(c, s, s->next_wpt);
}

static double los_guidance(const OrcShipConfig* c, OrcShipState* s, int k) {
  double N = s->north, E = s->east;
  double dx = s->wp_north[k] - s->wp_north[k - 1];
  double dy = s->wp_east[k] - s->wp_east[k - 1];
  double alpha_k = atan2(dy, dx);
  double e_ct = -(N - s->wp_north[k - 1]) * sin(alpha_k) + (E - s->wp_east[k - 1]) * cos(alpha_k);
  double R = c->lookahead_distance;
  s->e_ct = e_ct;
  if (e_ct * e_ct >= R * R) { e_ct = 0.99 * R; s->e_ct = e_ct; }
  double delta = sqrt(R * R - e_ct * e_ct);
  if (!(delta > 1e-6)) delta = 1e-6;                            /* max(1e-6, sqrt(...)) */
  if (fabs(s->e_ct_int + e_ct / delta) <= c->integrator_windup_limit) s->e_ct_int += e_ct / delta;
  double chi_r = atan(-e_ct / delta - s->e_ct_int * c->integral_gain);
  return alpha_k + chi_r;
}

/* PidController.pid_ctrl, controllers.py:106-118 */
static double pid_ctrl(double kp, double kd, double ki, double dt, double* err_i, double* prev_err,
                       double setpoint, double measurement) {
  double error = setpoint - measurement;
  double d_error = (error - *prev_err) / dt;
  double error_i = *err_i + error * dt;
  *prev_err = error;
  *err_i = error_i;
  return error * kp + d_error * kd + error_i * ki;
}

/* PiController.pi_ctrl, controllers.py:55-65 */
static double pi_ctrl(double kp, double ki, double dt, double* err_i, double setpoint, double measurement) {
  double error = setpoint - measurement;
  double error_i = *err_i + error * dt;
  *err_i = error_i;
  return error * kp + error_i * ki;
}

/* HeadingBySampledRouteController.rudder_angle_from_sampled_route, controllers.py:425-433
 * (= HeadingByRouteController.rudder_angle_from_route :344-352 when the offset is 0) */
static double autopilot_rudder(const OrcShipConfig* c, OrcShipState* s, double heading_offset) {
  double heading_ref = los_heading_ref(c, s);
  double out = pid_ctrl(c->hdg_kp, c->hdg_kd, c->hdg_ki, c->ctrl_time_step, &s->hdg_err_i, &s->hdg_prev_err,
                        heading_ref + heading_offset, s->yaw);
  return sat(-out, -c->max_rudder_angle, c->max_rudder_angle);  /* controllers.py:294-295 */
}

/* ThrustFromSpeedSetPoint.thrust (run_colav controllers.py:183-185) or
 * EngineThrottleFromSpeedSetPoint.throttle (rl_env controllers.py:185-189) with
 * measured_shaft_speed = forward_speed (rl_env env.py:322-326,397-401,494-498). */
static double speed_command(const OrcShipConfig* c, OrcShipState* s, double set_point) {
  if (c->model_kind == ORC_MODEL_SIMPLIFIED) {
    /* ThrottleFromSpeedSetPointSimplifiedPropulsion.throttle, rl_env controllers.py:212-232 */
    double thr = pi_ctrl(c->kp_ship_speed, c->ki_ship_speed, c->ctrl_time_step, &s->spd_err_i, set_point, s->u);
    return sat(thr, 0, 1.1);
  }
  if (c->model_kind == ORC_MODEL_SIMPLE) {
    double out = pid_ctrl(c->spd_kp, c->spd_kd, c->spd_ki, c->ctrl_time_step, &s->spd_err_i, &s->spd_prev_err,
                          set_point, s->u);
    return sat(out, -c->max_thrust, c->max_thrust);
  }
  double w_d = pi_ctrl(c->kp_ship_speed, c->ki_ship_speed, c->ctrl_time_step, &s->spd_err_i, set_point, s->u);
  w_d = sat(w_d, 0, c->max_shaft_speed);
  double thr = pi_ctrl(c->kp_shaft_speed, c->ki_shaft_speed, c->ctrl_time_step, &s->shaft_err_i, w_d, s->u);
  return sat(thr, 0, 1.1);
}

/* update_differentials + integrate_differentials + int.next_time:
 * SimpleShipModel ship_model.py:351-416; ShipModelAST rl_env ship_model.py:834-901;
 * ShipMachineryModel ship_engine.py:403-443; EulerInt utils.py:42-53. */
static const OrcSimplifiedMachinery* g_simplified = NULL;   /* set by orc_simplified_rollout / orc_set_simplified */
static OrcSimplifiedMachinery g_simplified_env;              /* copy made by orc_set_simplified */

/* env-level runs of the thrust-state model (A8'): the machinery constants the ship configuration struct does not
 * hold, for every ORC_MODEL_SIMPLIFIED ship stepped afterwards (not re-entrant; NULL switches it off).  The initial
 * thrust force is the ship's initial_propeller_shaft_speed_rad_per_s field (the machinery state's slot). */
void orc_set_simplified(const OrcSimplifiedMachinery* m) {
  if (m) { g_simplified_env = *m; g_simplified = &g_simplified_env; }
  else g_simplified = NULL;
}

static void ship_dynamics(const OrcShipConfig* c, const Derived* d, OrcShipState* s, double command,
                          double rudder_angle) {
  const double dt = c->integration_step;
  double cpsi = cos(s->yaw), spsi = sin(s->yaw);
  double u = s->u, v = s->v, r = s->r;
  /* three_dof_kinematics :177-186 */
  double d_north = cpsi * u + (-spsi) * v;
  double d_east = spsi * u + cpsi * v;
  double d_yaw = r;
  /* machinery */
  double thrust, d_omega = 0.0;
  if (c->model_kind == ORC_MODEL_SIMPLE) {
    thrust = command;
  } else if (c->model_kind == ORC_MODEL_SIMPLIFIED) {
    /* SimplifiedMachineryModel.update_thrust_force, ship_engine.py:508-513; the machinery state (thrust
     * force, in s->omega's slot) feeds the kinetics before it is integrated, like ShipModelAST's shaft speed
     * (rl_env ship_model.py:882-901) */
    double power = command * (d->p_me + d->p_el);
    thrust = s->omega;
    d_omega = (-(2160.0 / 790.0) * s->omega + power) / g_simplified->thrust_force_dynamic_time_constant;
  } else {
    double w = s->omega;
    double a_me = command * d->p_me / (w + 0.1);
    double tq_me = (d->tq_me_max < a_me) ? d->tq_me_max : a_me;       /* min(a, b) */
    double a_el = command * d->p_el / (w + 0.1);
    double tq_el = (d->tq_el_max < a_el) ? d->tq_el_max : a_el;
    double eq_me = (tq_me - c->linear_friction_main_engine * w) / c->gear_ratio_between_main_engine_and_propeller;
    double eq_hsg = (tq_el - c->linear_friction_hybrid_shaft_generator * w) /
                    c->gear_ratio_between_hybrid_shaft_generator_and_propeller;
    d_omega = (eq_me + eq_hsg - c->propeller_speed_to_torque_coefficient * (w * w)) / c->propeller_inertia;
    thrust = d->dp4 * c->propeller_speed_to_thrust_force_coefficient * w * fabs(w);
  }
  /* current in body frame: inv(R) . vel_c, :367-369 */
  double vn = c->current_velocity_component_from_north, ve = c->current_velocity_component_from_east;
  double u_c = cpsi * vn + spsi * ve;
  double v_c = (-spsi) * vn + cpsi * ve;
  double u_r = u - u_c, v_r = v - v_c;
  /* rudder :394-397 */
  double f_rudder_v = -c->rudder_angle_to_sway_force_coefficient * rudder_angle * (u - u_c);
  double f_rudder_r = -c->rudder_angle_to_yaw_force_coefficient * rudder_angle * (u - u_c);
  /* get_wind_force :162-175 */
  double uw = c->wind_speed * cos(c->wind_direction - s->yaw);
  double vw = c->wind_speed * sin(c->wind_direction - s->yaw);
  double u_rw = uw - u, v_rw = vw - v;
  double gamma_rw = -atan2(v_rw, u_rw);
  double wind_rw2 = u_rw * u_rw + v_rw * v_rw;
  double c_x = -0.5 * cos(gamma_rw);
  double c_y = 0.7 * sin(gamma_rw);
  double c_n = 0.08 * sin(2 * gamma_rw);
  double tau_coeff = 0.5 * 1.2 * wind_rw2;
  double tau_u = tau_coeff * c_x * d->proj_area_f;
  double tau_v = tau_coeff * c_y * d->proj_area_l;
  double tau_n = tau_coeff * c_n * d->proj_area_l * d->l_ship;
  /* kinetics :373-381 */
  double m = d->mass;
  double crb0 = (-m * v) * r;
  double crb1 = (m * u) * r;
  double crb2 = (m * v) * u + (-m * u) * v;
  double ca0 = (d->y_dv * v_r) * r;
  double ca1 = (-d->x_du * u_r) * r;
  double ca2 = (-d->y_dv * v_r) * u_r + (d->x_du * u_r) * v_r;
  double dmp0 = (m / c->mass_over_linear_friction_coefficient_in_surge + c->nonlinear_friction_coefficient__in_surge * u) * u_r;
  double dmp1 = (m / c->mass_over_linear_friction_coefficient_in_sway + c->nonlinear_friction_coefficient__in_sway * v) * v_r;
  double dmp2 = (d->i_z / c->mass_over_linear_friction_coefficient_in_yaw + c->nonlinear_friction_coefficient__in_yaw * r) * r;
  double f0 = -crb0 - ca0 - dmp0 + tau_u + 0.0 + thrust;
  double f1 = -crb1 - ca1 - dmp1 + tau_v + 0.0 + f_rudder_v;
  double f2 = -crb2 - ca2 - dmp2 + tau_n + 0.0 + f_rudder_r;
  double d_u = (1.0 / (m + d->x_du)) * f0;                      /* inv(diag) . f */
  double d_v = (1.0 / (m + d->y_dv)) * f1;
  double d_r = (1.0 / (d->i_z + d->n_dr)) * f2;
  /* integrate_differentials :411-416 / :895-901 */
  s->north = s->north + d_north * dt;
  s->east = s->east + d_east * dt;
  s->yaw = s->yaw + d_yaw * dt;
  s->u = s->u + d_u * dt;
  s->v = s->v + d_v * dt;
  s->r = s->r + d_r * dt;
  if (c->model_kind != ORC_MODEL_SIMPLE) s->omega = s->omega + d_omega * c->dt_shaft;
  s->time = s->time + dt;                                       /* next_time, utils.py:42-48 */
  s->last_rudder = rudder_angle;
  s->last_thrust = command;
}

/* store_simulation_data: only the entries the step path reads back (e_ct and positions) */
static void log_row(OrcShipState* s) {
  s->log_prev_north = s->log_north; s->log_prev_east = s->log_east;
  s->log_north = s->north; s->log_east = s->east;
  s->log_e_ct = s->e_ct;
  s->n_log += 1;
}

void orc_simplified_rollout(const OrcShipConfig* c, const OrcSimplifiedMachinery* m, OrcShipState* s, int64_t n_steps,
                            int record_every, double* out_states, int32_t* out_wpt) {
  /* bare loop of a hull driven by SimplifiedMachineryModel (A8' of SURVEY.md section 8a); not re-entrant */
  const OrcSimplifiedMachinery* saved = g_simplified;
  g_simplified = m;
  orc_ship_rollout(c, s, n_steps, record_every, out_states, out_wpt);
  g_simplified = saved;
}

void orc_ship_rollout(const OrcShipConfig* c, OrcShipState* s, int64_t n_steps, int record_every,
                      double* out_states, int32_t* out_wpt) {
  Derived d; derive(c, &d);
  if (record_every <= 0) record_every = 1;
  int64_t row = 0;
  for (int64_t i = 1; i <= n_steps; ++i) {
    double rudder = autopilot_rudder(c, s, 0.0);
    double cmd = speed_command(c, s, c->desired_forward_speed);
    ship_dynamics(c, &d, s, cmd, rudder);
    if (i % record_every == 0) {
      if (out_states) {
        double* o = out_states + 8 * row;
        o[0] = s->north; o[1] = s->east; o[2] = s->yaw; o[3] = s->u; o[4] = s->v; o[5] = s->r;
        o[6] = s->omega; o[7] = s->e_ct;
      }
      if (out_wpt) out_wpt[row] = s->next_wpt;
      ++row;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * map geometry: PolygonObstacle, obstacle.py:126-141 (Shapely semantics, see header)
 * ------------------------------------------------------------------------------------------- */
int orc_map_contains(const OrcMap* m, double n_pos, double e_pos) {
  double x = e_pos, y = n_pos;                                  /* Point(e_pos, n_pos) */
  for (int p = 0; p < m->n_poly; ++p) {
    int a = m->poly_start[p], b = m->poly_start[p + 1];
    int inside = 0;
    for (int i = a, j = b - 1; i < b; j = i++) {
      double xi = m->vert_e[i], yi = m->vert_n[i], xj = m->vert_e[j], yj = m->vert_n[j];
      if ((yi > y) != (yj > y)) {
        if (x < (xj - xi) * (y - yi) / (yj - yi) + xi) inside = !inside;
      }
    }
    if (inside) return 1;
  }
  return 0;
}

double orc_map_distance(const OrcMap* m, double n_pos, double e_pos) {
  double px = e_pos, py = n_pos, best = INFINITY;
  for (int p = 0; p < m->n_poly; ++p) {
    int a = m->poly_start[p], b = m->poly_start[p + 1];
    for (int i = a; i < b; ++i) {
      int k = (i + 1 < b) ? i + 1 : a;
      double ax = m->vert_e[i], ay = m->vert_n[i], bx = m->vert_e[k], by = m->vert_n[k];
      double dx = bx - ax, dy = by - ay;
      double l2 = dx * dx + dy * dy, t = 0.0;
      if (l2 != 0.0) {
        t = ((px - ax) * dx + (py - ay) * dy) / l2;
        t = (t < 1.0) ? t : 1.0;
        t = (t > 0.0) ? t : 0.0;
      }
      double cx = ax + t * dx, cy = ay + t * dy;
      double dd = sqrt((px - cx) * (px - cx) + (py - cy) * (py - cy));
      if (dd < best) best = dd;
    }
  }
  return best;
}

static void map_bounds(const OrcMap* m, double* min_n, double* max_n, double* min_e, double* max_e) {
  int nv = m->poly_start[m->n_poly];                            /* map_boundaries, obstacle.py:111-124 */
  *min_n = *min_e = INFINITY; *max_n = *max_e = -INFINITY;
  for (int i = 0; i < nv; ++i) {
    if (m->vert_e[i] < *min_e) *min_e = m->vert_e[i];
    if (m->vert_e[i] > *max_e) *max_e = m->vert_e[i];
    if (m->vert_n[i] < *min_n) *min_n = m->vert_n[i];
    if (m->vert_n[i] > *max_n) *max_n = m->vert_n[i];
  }
}

/* check_condition.py:16-46 */
static int is_pos_outside_horizon(const OrcMap* m, double n, double e, double ship_length) {
  double min_n, max_n, min_e, max_e; map_bounds(m, &min_n, &max_n, &min_e, &max_e);
  double margin = ship_length / 2;
  int outside_n = n < min_n + margin || n > max_n - margin;
  int outside_e = e < min_e + margin || e > max_e - margin;
  return outside_n || outside_e;
}
/* check_condition.py:48-78 */
static int is_pos_inside_obstacles(const OrcMap* m, double n, double e, double ship_length) {
  double margin = ship_length / 2;
  double mn = n - margin, me = e - margin, xn = n + margin, xe = e + margin;
  int inside = 0;
  if (orc_map_contains(m, mn, me)) inside = 1;
  if (orc_map_contains(m, mn, xe)) inside = 1;
  if (orc_map_contains(m, xn, me)) inside = 1;
  if (orc_map_contains(m, xn, xe)) inside = 1;
  return inside;
}
/* check_condition.py:80-107 */
static int is_route_outside_horizon(const OrcMap* m, double n, double e) {
  double min_n, max_n, min_e, max_e; map_bounds(m, &min_n, &max_n, &min_e, &max_e);
  return (n < min_n || n > max_n) || (e < min_e || e > max_e);
}

/* Python float modulo (numpy npy_divmod semantics) */
static double py_mod(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; }
  else m = copysign(0.0, b);
  return m;
}

/* RewardDesign3 / RewardDesign4, rl_env/reward_designs.py:33-55 */
static double reward_design3(double val, double target, double offset) {
  return (val < target) ? exp(-((val - target) * (val - target)) / offset) : 1.0;
}
static double reward_design4(double val, double target, double offset) {
  return (val < target) ? 1.0 : exp(-((val - target) * (val - target)) / offset);
}

/* get_env_info (run_colav get_env_info.py:53-224) and get_reward_and_env_info
 * (rl_env reward_function.py:59-270) share the same flags; the RL variant adds the reward. */
static double evaluate(const OrcEnvConfig* cfg, OrcEnvState* st, int* events, int* terminal, int* test_stop,
                       int* obs_stop) {
  const OrcShipState* t = &st->ship[0];
  const OrcShipState* o = &st->ship[1];
  double t_len = cfg->ship[0].length_of_ship, o_len = cfg->ship[1].length_of_ship;
  double test_e_ct = t->log_e_ct, obs_e_ct = o->log_e_ct;
  /* compute_distance.py:16-40 */
  double dx = o->north - t->north, dy = o->east - t->east;
  double distance = sqrt(dx * dx + dy * dy);
  double phi = atan2(dy, dx);
  double beta = phi - t->yaw;
  beta = py_mod(beta + ORC_PI, 2 * ORC_PI) - ORC_PI;
  int enc;  /* 0 head-on, 1 overtaking, 2 crossing */
  if (fabs(beta) < 15.0 * (ORC_PI / 180.0)) enc = 0;
  else if (fabs(beta) > 165.0 * (ORC_PI / 180.0)) enc = 1;
  else enc = 2;
  /* check_condition.py:142-158 */
  double sd = (t->north - o->north) * (t->north - o->north) + (t->east - o->east) * (t->east - o->east);
  int is_collision = sd < 50.0 * 50.0;
  double test_gd = orc_map_distance(&cfg->map, t->north, t->east);
  int is_test_grounding = is_pos_inside_obstacles(&cfg->map, t->north, t->east, t_len);
  double obs_gd = orc_map_distance(&cfg->map, o->north, o->east);
  int is_obs_grounding = is_pos_inside_obstacles(&cfg->map, o->north, o->east, o_len);
  int is_test_nav_failure = fabs(test_e_ct) > 3000.0;
  int is_obs_nav_failure = (st->travel_dist > st->ab_segment_length * 2) || (st->travel_time > INFINITY) ||
                           (fabs(obs_e_ct) > 500.0);
  double r_total = 0.0;
  if (cfg->env_kind == ORC_ENV_RL) {
    /* ships_collision_reward :316-357 -- the "overtake" branch never matches "overtaking" */
    double r1 = 0.0;
    if (distance < 10000.0 && (enc == 0 || enc == 2)) r1 = reward_design4(distance, 0.0, 200000000.0);
    double r2 = 0.0;                                            /* test_ship_grounding_reward :359-393 */
    if (test_gd <= 1000.0) r2 = reward_design4(test_gd, 0.0, 175000.0);
    double r3 = reward_design3(fabs(test_e_ct), 3000.0, 1250000.0);      /* :395-425 */
    double r4 = 0.0;                                            /* obs_ship_grounding_reward :427-461 */
    if (obs_gd <= 1000.0) r4 = -reward_design4(obs_gd, 0.0, 50000.0);
    double r5 = -reward_design3(fabs(obs_e_ct), 500.0, 12500.0);        /* :463-494 */
    r_total = ((((r1 + r2) + r3) + r4) + r5) / 5;               /* np.sum of 5 then / len */
    /* get_reward_due_to_ships_termination :272-314 */
    int conds[5] = {is_collision, is_test_grounding, is_test_nav_failure, is_obs_grounding, is_obs_nav_failure};
    double mult[5] = {10.0, 5.0, 5.0, -2.5, -2.5};
    if (conds[0] || conds[1] || conds[2] || conds[3] || conds[4]) {
      double acc = st->accumulated_rewards;
      double reward = r_total + acc;
      r_total = 0;
      for (int i = 0; i < 5; ++i) {
        if (acc > 0 && conds[i]) r_total += reward * mult[i];
        else if (acc < 0 && conds[i]) r_total += reward * -mult[i];
      }
    }
  }
  const OrcShipConfig* tc = &cfg->ship[0];
  const OrcShipConfig* oc = &cfg->ship[1];
  (void)tc; (void)oc;
  /* check_condition.py:5-14 with route ends navigate.north[-1] */
  double tdn = t->north - t->wp_north[t->n_wp - 1], tde = t->east - t->wp_east[t->n_wp - 1];
  int t6 = sqrt(tdn * tdn + tde * tde) <= 200.0;
  int t7 = is_pos_outside_horizon(&cfg->map, t->north, t->east, t_len);
  double odn = o->north - o->wp_north[o->n_wp - 1], ode = o->east - o->wp_east[o->n_wp - 1];
  int t8 = sqrt(odn * odn + ode * ode) <= 200.0;
  int t9 = is_pos_outside_horizon(&cfg->map, o->north, o->east, o_len);
  int t10 = t->time > cfg->ship[0].simulation_time;             /* is_within_simu_time_limit(test) */
  int ev = 0, term = 0, ts = 0, os = 0;
  if (is_collision) { ev |= ORC_EV_COLLISION; term = 1; ts = 1; os = 1; }
  if (is_test_grounding) { ev |= ORC_EV_TEST_GROUNDING; term = 1; ts = 1; }
  if (is_test_nav_failure) { ev |= ORC_EV_TEST_NAV_FAILURE; term = 1; ts = 1; }
  if (is_obs_grounding) { ev |= ORC_EV_OBS_GROUNDING; term = 1; os = 1; }
  if (is_obs_nav_failure) { ev |= ORC_EV_OBS_NAV_FAILURE; term = 1; os = 1; }
  if (t6) { ev |= ORC_EV_TEST_REACHED; ts = 1; }
  if (t7) { ev |= ORC_EV_TEST_OUTSIDE; ts = 1; }
  if (t8) { ev |= ORC_EV_OBS_REACHED; os = 1; }
  if (t9) { ev |= ORC_EV_OBS_OUTSIDE; os = 1; }
  if (t10) { ev |= ORC_EV_TIME_LIMIT; ts = 1; os = 1; }
  *events = ev; *terminal = term; *test_stop = ts; *obs_stop = os;
  return r_total;
}

/* ---------------------------------------------------------------------------------------------
 * env: __init__ / reset / init_step / _step / step
 * ------------------------------------------------------------------------------------------- */
/* init_get_intermediate_waypoints: run_colav env.py:904-930, rl_env env.py:143-169 */
static void init_iw(const OrcEnvConfig* cfg, OrcEnvState* st) {
  const OrcShipState* o = &st->ship[1];
  double ab_n = o->wp_north[o->n_wp - 1] - o->wp_north[0];
  double ab_e = o->wp_east[o->n_wp - 1] - o->wp_east[0];
  double ab_len = sqrt(ab_n * ab_n + ab_e * ab_e);
  int div = cfg->max_sampling_frequency + 1;
  st->ab_segment_length = ab_len / div;
  st->ab_north_segment_length = ab_n / div;
  st->ab_east_segment_length = ab_e / div;
  double ab_alpha = atan2(ab_e, ab_n);
  double ab_beta = ORC_PI / 2 - ab_alpha;
  st->omega = ORC_PI / 2 - ab_beta;
  st->n_base = st->ab_north_segment_length + o->wp_north[0];
  st->e_base = st->ab_east_segment_length + o->wp_east[0];
  st->sampling_count = 0;
  st->tracker_active = 0;
  st->travel_dist = 0;
  st->travel_time = 0;
}

static void snapshot_init(OrcEnvState* st) {                    /* results_snapshot :932-943 */
  memcpy(st->next_observations, st->initial_states, sizeof(st->initial_states));
  st->accumulated_rewards = 0;
  st->snapshot_events = 0; st->snapshot_terminal = 0; st->snapshot_test_stop = 0; st->snapshot_obs_stop = 0;
}

void orc_env_construct(const OrcEnvConfig* cfg, OrcEnvState* st) {
  memset(st, 0, sizeof(*st));
  orc_ship_init(&cfg->ship[0], &st->ship[0]);
  orc_ship_init(&cfg->ship[1], &st->ship[1]);
  /* initial_states, run_colav env.py:871-873 / rl_env env.py:107-109 */
  st->initial_states[0] = (float)st->ship[0].north; st->initial_states[1] = (float)st->ship[0].east;
  st->initial_states[2] = 0.0f;
  st->initial_states[3] = (float)st->ship[1].north; st->initial_states[4] = (float)st->ship[1].east;
  st->initial_states[5] = (float)st->ship[1].yaw; st->initial_states[6] = 0.0f;
  st->initial_states[7] = (float)st->ship[1].u;
  memcpy(st->states, st->initial_states, sizeof(st->states));
  init_iw(cfg, st);
  snapshot_init(st);
  st->sb_p_last = 1.0; st->sb_chi_last = 0.0; st->sb_active = 0;   /* SBMPCParams defaults */
}

void orc_env_init_step(const OrcEnvConfig* cfg, OrcEnvState* st) {   /* :1053-1097 / :297-342 */
  for (int i = 0; i < 2; ++i) {
    const OrcShipConfig* c = &cfg->ship[i];
    OrcShipState* s = &st->ship[i];
    Derived d; derive(c, &d);
    double rudder = autopilot_rudder(c, s, 0.0);
    double cmd = speed_command(c, s, c->desired_forward_speed);
    log_row(s);
    ship_dynamics(c, &d, s, cmd, rudder);
  }
  st->tracker_active = 1;
}

void orc_env_reset(OrcEnvConfig* cfg, OrcEnvState* st) {        /* :997-1051 / :238-295 */
  int64_t n = st->n_substeps;
  for (int i = 0; i < 2; ++i) {
    /* BaseMachineryModel.reset leaves the shaft integrator at EulerInt's default dt = 0.01
     * (ship_engine.py:331-333 via :476-481) */
    if (cfg->ship[i].model_kind == ORC_MODEL_DETAILED) cfg->ship[i].dt_shaft = 0.01;
    orc_ship_init(&cfg->ship[i], &st->ship[i]);
  }
  init_iw(cfg, st);
  snapshot_init(st);
  /* NOTE: self.states is NOT reset by reset() (only __init__ sets it) */
  orc_env_init_step(cfg, st);
  st->n_substeps = n;
}

/* is_collision_imminent on the float32 self.states, check_condition.py:130-140 */
static int collision_risk_f32(const OrcEnvState* st) {
  float dn = st->states[0] - st->states[3], de = st->states[1] - st->states[4];
  float d2 = dn * dn + de * de;
  return d2 < 9000000.0f;
}

/* ---------------------------------------------------------------------------------------------
 * SBMPC collision avoidance: sbmpc.py:90-314, sbmpc_misc.py:3-123, as the env calls it
 * (rl_env env.py:360-385, run_colav env.py:1145-1170) with SBMPC(tf=1000, dt=20) and default SBMPCParams
 * (KAPPA_ = 0, so the COLREGs term mu never contributes and is not restated).
 * ------------------------------------------------------------------------------------------- */
#define SB_NSAMP 50                      /* int(T / DT) = int(1000 / 20) */
#define SB_DT 20.0

static double wrap_pmpi(double a) {      /* wrap_angle_to_pmpi, sbmpc_misc.py:3-33 */
  return -ORC_PI + py_mod(a - (-ORC_PI), ORC_PI - (-ORC_PI));
}

static void sbmpc_offsets(const OrcEnvConfig* cfg, OrcEnvState* st, double u_d, double chi_d,
                          double* speed_factor, double* heading_offset) {
  const OrcShipState* os = &st->ship[0];
  const OrcShipState* ob = &st->ship[1];
  /* os_state = [east, north, -yaw, u, v, r]; obstacle state = [east, north, -yaw, u, v] */
  const double os_x = os->east, os_y = os->north, os_v = os->v;
  const double obs_l = cfg->ship[1].length_of_ship, obs_w = cfg->ship[1].width_of_ship;
  const double os_l = 25.0, os_w = 80.0;                   /* ShipLinearModel defaults, sbmpc_misc.py:86 */
  (void)os_w;
  /* Obstacle.__init__ + calculate_trajectory, sbmpc_misc.py:34-83 */
  double ox[SB_NSAMP], oy[SB_NSAMP];
  const double opsi = -ob->yaw, ou = ob->u, ov = ob->v;
  const double o11 = -sin(opsi), o12 = cos(opsi), o21 = cos(opsi), o22 = sin(opsi);
  ox[0] = ob->east; oy[0] = ob->north;
  for (int i = 1; i < SB_NSAMP; ++i) {
    ox[i] = ox[i - 1] + (o11 * ou + o12 * ov) * SB_DT;
    oy[i] = oy[i - 1] + (o21 * ou + o22 * ov) * SB_DT;
  }
  /* activation, sbmpc.py:154-166 */
  const double d0 = ox[0] - os_x, d1 = oy[0] - os_y;
  st->sb_active = sqrt(d0 * d0 + d1 * d1) < 2000.0;
  if (!st->sb_active) {
    st->sb_p_last = 1; st->sb_chi_last = 0;
    *speed_factor = 1; *heading_offset = 0;
    return;
  }
  static const double chi_deg[7] = {-30.0, -20.0, -10.0, 0.0, 10.0, 20.0, 30.0};
  static const double p_ca[4] = {0.4, 0.6, 0.8, 1.0};
  const double PHI = 68.5 * (ORC_PI / 180.0);               /* PHI_AH_ = PHI_OT_ */
  const double cos_ot = cos(PHI * (ORC_PI / 180.0));        /* np.cos(np.deg2rad(PHI_OT_)): degrees twice */
  const double d_safe = 1000.0, d_close = 2000.0;
  /* obstacle velocity in the world frame: rot2d(obstacle.psi_, [u, v]), sbmpc.py:312-314 */
  const double vo0 = -sin(opsi) * ou + cos(opsi) * ov, vo1 = cos(opsi) * ou + sin(opsi) * ov;
  const double n_vo = sqrt(vo0 * vo0 + vo1 * vo1);
  double cost = INFINITY, u_best = 1, chi_best = 0;
  for (int ic = 0; ic < 7; ++ic) {
    const double chi_ca = chi_deg[ic] * (ORC_PI / 180.0);   /* np.deg2rad */
    for (int jp = 0; jp < 4; ++jp) {
      /* ShipLinearModel.linear_pred(os_state, u_d * P, chi_d + Chi), sbmpc_misc.py:102-122 */
      const double ud = u_d * p_ca[jp], psi_d = chi_d + chi_ca;
      const double psi0 = wrap_pmpi(psi_d);
      const double r11 = -sin(psi_d), r12 = cos(psi_d), r21 = cos(psi_d), r22 = sin(psi_d);
      double sx = os_x, sy = os_y, su = ud, sv = os_v;
      double H1 = 0, t = 0;
      for (int i = 0; i < SB_NSAMP; ++i) {
        if (i > 0) {
          sx = sx + SB_DT * (r11 * su + r12 * sv);
          sy = sy + SB_DT * (r21 * su + r22 * sv);
          su = ud; sv = 0;
        }
        const double spsi = (i == 0) ? psi0 : psi_d;
        t += SB_DT;
        const double e0 = ox[i] - sx, e1 = oy[i] - sy;
        const double dist = sqrt(e0 * e0 + e1 * e1);
        double R = 0, Cc = 0;
        if (dist < d_close) {
          const double vs0 = -sin(spsi) * su + cos(spsi) * sv, vs1 = cos(spsi) * su + sin(spsi) * sv;
          double d_safe_i;
          const double phi_o = wrap_pmpi(atan2(-e1, -e0) - opsi + ORC_PI / 2);
          if (phi_o < PHI) d_safe_i = d_safe + obs_l / 2;
          else if (phi_o > PHI) d_safe_i = 0.5 * d_safe + obs_l / 2;
          else d_safe_i = d_safe + obs_w / 2;
          const double n_vs = sqrt(vs0 * vs0 + vs1 * vs1);
          const double dot = vs0 * vo0 + vs1 * vo1;
          if (dot > cos_ot * n_vs * n_vo && n_vs > n_vo) d_safe_i = d_safe + os_l / 2 + obs_l / 2;
          if (dist < d_safe_i) {
            R = (1 / pow(fabs(t - 0), 1.0)) * pow(d_safe / dist, 4.0);
            const double k_coll = 1e-6 * os_l * obs_l;
            const double w0 = vs0 - vo0, w1 = vs1 - vo1;
            Cc = k_coll * pow(sqrt(w0 * w0 + w1 * w1), 2.0);
          }
        }
        const double H0 = Cc * R + 0.0;                      /* + KAPPA_ * mu with KAPPA_ = 0 */
        if (H0 > H1) H1 = H0;
      }
      const double d_chi = chi_ca - st->sb_chi_last;         /* delta_Chi, sbmpc.py:303-310 */
      double dl_chi = 0;
      if (d_chi > 0) dl_chi = 20 * (d_chi * d_chi);
      else if (d_chi < 0) dl_chi = 30 * (d_chi * d_chi);
      const double H2 = 25 * (1 - p_ca[jp]) + 30 * (chi_ca * chi_ca) + 20 * fabs(st->sb_p_last - p_ca[jp]) + dl_chi;
      const double cost_i = H1 + H2;                         /* single obstacle: max over k */
      if (cost_i < cost) { cost = cost_i; u_best = p_ca[jp]; chi_best = chi_ca; }
    }
  }
  st->sb_p_last = u_best; st->sb_chi_last = chi_best;
  *speed_factor = u_best; *heading_offset = chi_best;
}

/* one asset's part of _step(): test_step / obs_step.  out5 = the returned next_states array
 * (3 entries in the normal branch of test_step / NonIW obs_step, 5 otherwise). */
static int asset_step(const OrcEnvConfig* cfg, OrcEnvState* st, int who, float* out5) {
  const OrcShipConfig* c = &cfg->ship[who];
  OrcShipState* s = &st->ship[who];
  const double dt = c->integration_step;
  int has_stop_branch = (who == 1) || (cfg->env_kind != ORC_ENV_RL);   /* quirk 6 */
  if (has_stop_branch && s->stop_flag) {
    /* :1104-1132: store_last_simulation_data, next_time twice */
    s->log_prev_north = s->log_north; s->log_prev_east = s->log_east; s->n_log += 1;
    s->time = s->time + dt;
    s->time = s->time + dt;
    out5[0] = (float)s->north; out5[1] = (float)s->east; out5[2] = (float)s->yaw;
    out5[3] = 0.0f; out5[4] = (float)s->log_e_ct;
    return 5;
  }
  double forward_speed = s->u;
  int collav_here = (who == 0) || (cfg->env_kind == ORC_ENV_COLAV_NONIW);
  double speed_factor = 1.0, heading_offset = 0.0;
  if (collav_here && cfg->collav == ORC_COLLAV_SBMPC) {
    /* rl_env env.py:360-385: next_wpt's result is discarded, los_guidance runs on the OLD waypoint
     * index and advances the LOS integrator a first time (quirk 2 of SURVEY.md section 8) */
    double chi_d = los_guidance(c, s, s->next_wpt);
    sbmpc_offsets(cfg, st, c->desired_forward_speed, -chi_d, &speed_factor, &heading_offset);
  }
  double rudder = autopilot_rudder(c, s, -heading_offset);
  double cmd = speed_command(c, s, c->desired_forward_speed * speed_factor);
  if (collav_here && cfg->collav == ORC_COLLAV_SIMPLE && collision_risk_f32(st)) {
    cmd *= 0.5;
    cmd = (cmd < 0.0) ? 0.0 : ((cmd > 1.1) ? 1.1 : cmd);         /* np.clip(x, 0.0, 1.1) */
    double bias = (cfg->env_kind == ORC_ENV_RL) ? (-15.0 * (ORC_PI / 180.0)) : (15.0 * (ORC_PI / 180.0));
    rudder += bias;
    rudder = (rudder < -c->max_rudder_angle) ? -c->max_rudder_angle
                                             : ((rudder > c->max_rudder_angle) ? c->max_rudder_angle : rudder);
  }
  Derived d; derive(c, &d);
  log_row(s);
  ship_dynamics(c, &d, s, cmd, rudder);
  out5[0] = (float)s->north; out5[1] = (float)s->east;
  int full = (who == 1) && (cfg->env_kind != ORC_ENV_COLAV_NONIW);
  if (!full) { out5[2] = (float)s->log_e_ct; return 3; }
  out5[2] = (float)s->yaw; out5[3] = (float)forward_speed; out5[4] = (float)s->log_e_ct;
  /* travel tracker :1309-1317 / :526-534 */
  if (st->tracker_active) {
    double tn = s->log_north - s->log_prev_north, te = s->log_east - s->log_prev_east;
    st->travel_dist += sqrt(tn * tn + te * te);
    st->travel_time += dt;
  }
  return 5;
}

void orc_env_substep(const OrcEnvConfig* cfg, OrcEnvState* st, OrcStepResult* out) {   /* _step() */
  float t5[5] = {0}, o5[5] = {0};
  asset_step(cfg, st, 0, t5);
  asset_step(cfg, st, 1, o5);
  float ns[8] = {0};
  ns[0] = t5[0]; ns[1] = t5[1]; ns[2] = t5[2];
  ns[3] = o5[0]; ns[4] = o5[1]; ns[5] = o5[2];
  if (cfg->env_kind != ORC_ENV_COLAV_NONIW) { ns[6] = o5[3]; ns[7] = o5[4]; }
  memcpy(st->states, ns, sizeof(ns));
  int ev, term, ts, os;
  double reward = evaluate(cfg, st, &ev, &term, &ts, &os);
  int done;
  if (cfg->env_kind == ORC_ENV_RL) {                            /* rl_env env.py:603-610 */
    done = ts && !term;
    if (os && !term) st->ship[1].stop_flag = 1;
  } else {                                                      /* run_colav env.py:1385-1399 */
    if (ts && !term) st->ship[0].stop_flag = 1;
    if (os && !term) st->ship[1].stop_flag = 1;
    done = st->ship[0].stop_flag && st->ship[1].stop_flag;
  }
  memcpy(out->obs, ns, sizeof(ns));
  out->last_step_reward = reward;
  out->done = term || done;
  out->events = ev; out->terminal = term; out->test_ship_stop = ts; out->obs_ship_stop = os;
  st->n_substeps += 1;
}

static int is_reach_roa(const OrcEnvConfig* cfg, const OrcEnvState* st) {   /* check_condition.py:181-204 */
  const OrcShipState* o = &st->ship[1];
  int k = o->next_wpt;
  double dn = o->north - o->wp_north[k], de = o->east - o->wp_east[k];
  return dn * dn + de * de < cfg->radius_of_acceptance * cfg->radius_of_acceptance;
}

void orc_env_step(const OrcEnvConfig* cfg, OrcEnvState* st, double action, OrcStepResult* out) {
  memset(out, 0, sizeof(*out));
  int is_roa = 0, combined_done = 0, have_iw = 0;
  int is_rl = cfg->env_kind == ORC_ENV_RL;
  OrcShipState* o = &st->ship[1];
  if (st->sampling_count < cfg->max_sampling_frequency) {
    /* obs_ship_uses_scoping_angle :1321-1344 + get_intermediate_waypoints :957-995 */
    st->sampling_count += 1;
    double l_s = fabs(st->ab_segment_length * tan(action));
    double e_s = l_s * cos(st->omega);
    double n_s = l_s * sin(st->omega);
    if (action > 0) e_s *= -1; else n_s *= -1;
    double rn = st->n_base + n_s, re = st->e_base + e_s;
    st->n_base = rn + st->ab_north_segment_length;
    st->e_base = re + st->ab_east_segment_length;
    /* update_route: list.insert(-1, .) controllers.py:417-422 */
    int n = o->n_wp;
    o->wp_north[n] = o->wp_north[n - 1]; o->wp_east[n] = o->wp_east[n - 1];
    o->wp_north[n - 1] = rn; o->wp_east[n - 1] = re;
    o->n_wp = n + 1;
    st->travel_dist = 0; st->travel_time = 0;
    have_iw = 1;                                                /* a non-empty list is truthy */
    int fail = orc_map_contains(&cfg->map, rn, re) || is_route_outside_horizon(&cfg->map, rn, re);
    if (fail) {                                                 /* :1462-1474 / :673-693 */
      memcpy(out->obs, st->next_observations, sizeof(out->obs));
      if (is_rl) {
        double acc = st->accumulated_rewards;                   /* reward_function.py:499-527, mult 2.0 */
        out->reward = (acc >= 0) ? (-acc * 2.0) : (acc * 2.0);
      }
      st->snapshot_events |= ORC_EV_SAMPLING_FAILURE;
      st->snapshot_terminal = 1; st->snapshot_test_stop = 0; st->snapshot_obs_stop = 0;
      out->done = 1; out->events = st->snapshot_events; out->terminal = 1;
      return;
    }
    if (is_rl) st->accumulated_rewards = 0;                     /* rl_env env.py:696 */
  }
  OrcStepResult sub;
  int have_obs = 0;
  while (!is_roa && !combined_done) {
    orc_env_substep(cfg, st, &sub); out->n_substeps++;
    combined_done = sub.done;
    if (is_rl) st->accumulated_rewards += sub.last_step_reward;
    is_roa = is_reach_roa(cfg, st);
    if (combined_done) { have_obs = 1; break; }
    if (is_roa && have_iw) {
      orc_env_substep(cfg, st, &sub); out->n_substeps++;
      combined_done = sub.done;
      if (is_rl) st->accumulated_rewards += sub.last_step_reward;
      have_obs = 1;
      if (st->sampling_count == cfg->max_sampling_frequency) {
        st->travel_dist = 0; st->travel_time = 0;
        while (!combined_done) {
          orc_env_substep(cfg, st, &sub); out->n_substeps++;
          combined_done = sub.done;
          if (is_rl) st->accumulated_rewards += sub.last_step_reward;
        }
      }
      break;
    }
  }
  if (!have_obs) { out->error = 1; return; }                    /* UnboundLocalError in the reference */
  memcpy(out->obs, sub.obs, sizeof(out->obs));
  memcpy(st->next_observations, sub.obs, sizeof(sub.obs));
  st->snapshot_events = sub.events; st->snapshot_terminal = sub.terminal;
  st->snapshot_test_stop = sub.test_ship_stop; st->snapshot_obs_stop = sub.obs_ship_stop;
  out->reward = is_rl ? st->accumulated_rewards : 0.0;
  out->last_step_reward = sub.last_step_reward;
  out->done = combined_done;
  out->events = sub.events; out->terminal = sub.terminal;
  out->test_ship_stop = sub.test_ship_stop; out->obs_ship_stop = sub.obs_ship_stop;
}

/* ---------------------------------------------------------------------------------------------
 * batch driver for the CPU baseline (bench.py cpu_baseline / --impl reference); pthreads, one
 * chunk of environments at a time per worker.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const OrcEnvConfig* cfg; int64_t n_envs; int n_rl_steps; const double* actions; const double* jitter_ne;
  double* out_return; int32_t* out_events;
  int64_t next; int64_t total; pthread_mutex_t mu;
} BenchJob;

static void* bench_worker(void* arg) {
  BenchJob* job = (BenchJob*)arg;
  OrcEnvConfig* c = (OrcEnvConfig*)malloc(sizeof(OrcEnvConfig));
  OrcEnvState* st = (OrcEnvState*)malloc(sizeof(OrcEnvState));
  int64_t local = 0;
  for (;;) {
    pthread_mutex_lock(&job->mu);
    int64_t b = job->next; job->next += 4;
    pthread_mutex_unlock(&job->mu);
    if (b >= job->n_envs) break;
    int64_t e_end = (b + 4 < job->n_envs) ? b + 4 : job->n_envs;
    for (int64_t e = b; e < e_end; ++e) {
      memcpy(c, job->cfg, sizeof(*c));
      if (job->jitter_ne) {
        for (int s = 0; s < 2; ++s) {
          c->ship[s].initial_north_position_m += job->jitter_ne[(e * 2 + s) * 2 + 0];
          c->ship[s].initial_east_position_m += job->jitter_ne[(e * 2 + s) * 2 + 1];
        }
      }
      orc_env_construct(c, st);
      orc_env_reset(c, st);
      OrcStepResult res; memset(&res, 0, sizeof(res));
      double ret = 0.0;
      for (int j = 0; j < job->n_rl_steps; ++j) {
        orc_env_step(c, st, job->actions[e * job->n_rl_steps + j], &res);
        ret += res.reward;
        if (res.done || res.error) break;
      }
      local += st->n_substeps;
      if (job->out_return) job->out_return[e] = ret;
      if (job->out_events) job->out_events[e] = res.events;
    }
  }
  pthread_mutex_lock(&job->mu); job->total += local; pthread_mutex_unlock(&job->mu);
  free(c); free(st);
  return NULL;
}

int64_t orc_bench_episodes(const OrcEnvConfig* cfg, int64_t n_envs, int n_rl_steps, const double* actions,
                           const double* jitter_ne, int n_threads, double* out_return, int32_t* out_events) {
  if (n_threads <= 0) { long n = sysconf(_SC_NPROCESSORS_ONLN); n_threads = n > 0 ? (int)n : 1; }
  if (n_threads > 256) n_threads = 256;
  BenchJob job = {cfg, n_envs, n_rl_steps, actions, jitter_ne, out_return, out_events, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  pthread_t th[256];
  for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, bench_worker, &job);
  for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
  return job.total;
}
