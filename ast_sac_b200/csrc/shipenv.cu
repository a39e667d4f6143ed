// shipenv.cu -- host side of the C ABI declared in include/shipenv.h (handle, validation, launches).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared ...
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <cmath>
#include <vector>

#include "shipenv.h"
#include "shipenv_launch.h"

#ifndef SHIPENV_STREAMING_K
#define SHIPENV_STREAMING_K 0
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) return fail(SHIPENV_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace

struct shipenv {
  int device = 0;
  long long num_envs = 0;
  ShipEnvParams params;
  ShipEnvParams* params_dev = nullptr;
  void* staged_dev = nullptr;         // the kernels' shared-memory block, rebuilt at every parameter upload
  ShipEnvBuffers buf{};
  bool bound = false;
  bool owns = false;
  bool constructed = false;
  const double* init_dev = nullptr;   // per-ship initial states given to shipenv_construct (caller-owned)
  unsigned long long* queue_dev = nullptr;   // [0] work-queue counter, [1] environments done after the launch,
                                             // [2], [3] split step() calls (launch_env): environments of the second
                                             // launch, its work-queue counter
  unsigned long long* done_host = nullptr;   // pinned copy of queue_dev[1] of the most recent completed launch
  int streaming_k = SHIPENV_STREAMING_K;    // _step() launches of at most this many steps use a static grid
  int quiet_min_k = 8;                       // _step() launches of fewer steps take no quiet steps; SHIPENV_QUIET_MIN_K
  int split_calls = 1;                       // step(action) as two launches (see launch_env); SHIPENV_SPLIT_CALLS=0: one
  int no_quiet = 0;                          // SHIPENV_QUIET=0: the env kernel takes no quiet steps (comparison runs)
  int persist_mode = 1;                      // 1 persistent grid + lane-pair refill (default), 0 one slot per environment, -1 auto
  // CUDA events around the env kernel itself (k_env), for shipenv_env_kernel_ms: the step() / _step() entry
  // points also launch the prologue kernel and a memset, which a caller's own events would include
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
  int time_kernels = 0;
  double kernel_ms_sum = 0.0;                // accumulated by shipenv_env_kernel_ms
  bool kernel_pending = false;
  int sm_count = 0;
  double* log_dev = nullptr;          // optional trajectory log (caller-owned), see shipenv_set_trajectory_log
  int32_t* log_count_dev = nullptr;
  long long log_envs = 0, log_capacity = 0;
  unsigned* grid_dev = nullptr;       // culling grid cells
  unsigned long long* edges_dev = nullptr;   // per-cell ring-segment masks
  float* safe_dev = nullptr;                 // per-cell safe radius (SenvGrid::safe)
  SenvGrid grid{};
  // staging for the *_host entry points
  double* act_dev = nullptr;
  uint8_t* mask_dev = nullptr;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_caller = nullptr;    // recorded on the caller's stream by every device-pointer entry point
  bool caller_pending = false;
  // caller buffers of the *_host entry points that were page-locked with cudaHostRegister: the copies then
  // go straight between the caller's memory and the device (no staging memcpy)
  struct HostReg { const void* ptr; size_t bytes; bool owned; };   // owned: registered by us (unregister at release)
  std::vector<HostReg> host_regs;
};

namespace {

int validate(const ShipEnvParams* p, long long num_envs) {
  if (!p) return fail(SHIPENV_E_ARG, "params is NULL");
  if (p->abi_version != SHIPENV_ABI_VERSION)
    return fail(SHIPENV_E_ARG, "params.abi_version %d != %d", p->abi_version, SHIPENV_ABI_VERSION);
  if (num_envs <= 0) return fail(SHIPENV_E_ARG, "num_envs must be positive");
  if (p->env_kind < SHIPENV_ENV_COLAV_NONIW || p->env_kind > SHIPENV_ENV_RL) return fail(SHIPENV_E_ARG, "bad env_kind");
  if (p->collav != SHIPENV_COLLAV_NONE && p->collav != SHIPENV_COLLAV_SIMPLE && p->collav != SHIPENV_COLLAV_SBMPC)
    return fail(SHIPENV_E_ARG, "collav mode %d not supported (none=0, simple=1, sbmpc=2)", p->collav);
  if (p->max_sampling_frequency < 0 || p->max_sampling_frequency > SHIPENV_MAX_IW)
    return fail(SHIPENV_E_ARG, "max_sampling_frequency must be in [0, %d]", SHIPENV_MAX_IW);
  if (p->n_poly < 0 || p->n_poly > SHIPENV_MAX_POLY) return fail(SHIPENV_E_ARG, "n_poly out of range");
  if (p->math_mode != SHIPENV_MATH_STRICT && p->math_mode != SHIPENV_MATH_FAST)
    return fail(SHIPENV_E_ARG, "math_mode must be SHIPENV_MATH_STRICT or SHIPENV_MATH_FAST");
  if (p->n_poly > 0 && (p->poly_start[0] != 0 || p->poly_start[p->n_poly] > SHIPENV_MAX_VERT))
    return fail(SHIPENV_E_ARG, "polygon vertex table out of range");
  for (int i = 0; i < p->n_poly; ++i)
    if (p->poly_start[i + 1] - p->poly_start[i] < 3) return fail(SHIPENV_E_ARG, "polygon %d has < 3 vertices", i);
  if (p->ship[0].model_kind != p->ship[1].model_kind)
    return fail(SHIPENV_E_ARG, "both ships of an environment must use the same ship model class");
  for (int s = 0; s < 2; ++s) {
    const ShipEnvShipParams& q = p->ship[s];
    if (q.model_kind < SHIPENV_MODEL_SIMPLE || q.model_kind > SHIPENV_MODEL_SIMPLIFIED)
      return fail(SHIPENV_E_ARG, "ship[%d].model_kind invalid", s);
    if (q.model_kind == SHIPENV_MODEL_SIMPLIFIED && !(q.thrust_tau > 0.0))
      return fail(SHIPENV_E_ARG, "ship[%d].thrust_tau must be positive", s);
    if (q.n_wp < 2 || q.n_wp > SHIPENV_MAX_WP) return fail(SHIPENV_E_ARG, "ship[%d].n_wp must be in [2, %d]", s, SHIPENV_MAX_WP);
    if (!(q.dt > 0.0) || !(q.ctrl_dt > 0.0)) return fail(SHIPENV_E_ARG, "ship[%d] time steps must be positive", s);
    // the fast build forms e_ct / sqrt(R^2 - e_ct^2) without the reference's clamp of the root to >= 1e-6
    // (LOS_guidance.py:113-115), which cannot bind for R > 1e-5 m: the root is >= 0.141 R
    if (p->math_mode == SHIPENV_MATH_FAST && !(q.los_r > 1e-5))
      return fail(SHIPENV_E_ARG, "ship[%d].los_r (lookahead distance) must exceed 1e-5 m in the fast build", s);
    if (q.model_kind != SHIPENV_MODEL_SIMPLE && !(q.dt_shaft > 0.0))
      return fail(SHIPENV_E_ARG, "ship[%d].dt_shaft must be positive", s);
  }
  if ((p->env_kind != SHIPENV_ENV_COLAV_NONIW || p->obs_sampled_route) && p->ship[1].n_wp + p->max_sampling_frequency > 255)
    return fail(SHIPENV_E_ARG, "route too long");
  return SHIPENV_OK;
}

SenvView view(const shipenv* h) {
  return SenvView{h->params_dev, h->staged_dev, h->buf, h->num_envs, h->grid, h->params.collav, h->sm_count, h->no_quiet, 0,
                  h->log_dev, h->log_count_dev, h->log_envs, h->log_capacity};
}

int check_ready(const shipenv* h, bool need_constructed) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  if (!h->bound) return fail(SHIPENV_E_STATE, "no buffers: call shipenv_bind or shipenv_alloc first");
  if (need_constructed && !h->constructed)
    return fail(SHIPENV_E_STATE, "environment state not initialised: call shipenv_construct or shipenv_reset first");
  return SHIPENV_OK;
}

// The *_host entry points run on the handle's own stream.  Work submitted through the device-pointer entry points
// on a caller stream is tracked with an event so that a later *_host call is ordered after it (a C caller may mix the
// two families without a synchronisation of its own).
int note_caller_stream(shipenv* h, cudaStream_t st) {
  if (h->stream && st == h->stream) return SHIPENV_OK;
  if (!h->ev_caller) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_caller, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(h->ev_caller, st));
  h->caller_pending = true;
  return SHIPENV_OK;
}

int launch_reset(shipenv* h, const uint8_t* mask, const double* init, int do_init, int reinit, cudaStream_t st) {
  const int model = h->params.ship[0].model_kind;
  cudaError_t e = (h->params.math_mode == SHIPENV_MATH_FAST)
                      ? senv_fast::launch_reset(view(h), model, mask, init, do_init, reinit, st)
                      : senv_strict::launch_reset(view(h), model, mask, init, do_init, reinit, st);
  if (e != cudaSuccess) return fail(SHIPENV_E_CUDA, "reset kernel launch: %s", cudaGetErrorString(e));
  return note_caller_stream(h, st);
}

int launch_env(shipenv* h, int mode, const double* actions, int k, cudaStream_t st) {
  const int model = h->params.ship[0].model_kind;
  // Grid policy.  With every environment alive a static grid (one slot per environment) is fastest;
  // once a noticeable share is done, a persistent grid whose lane pairs refill from the work queue
  // wins (finished environments cost one fetch instead of an idle lane pair).  The share comes from the
  // previous launch's count (copied to pinned memory without a sync, so it may lag by a launch).
  int persistent = h->persist_mode;
  if (persistent < 0) persistent = (double)(*h->done_host) > 0.08 * (double)h->num_envs ? 1 : 0;
  // step(action) as two launches (SenvView::call_filter, below)
  const bool has_quiet_twin = h->params.env_kind == SHIPENV_ENV_COLAV_IW && h->params.collav == SHIPENV_COLLAV_NONE;
  const bool split = mode == 0 && has_quiet_twin && !h->no_quiet && h->split_calls && persistent;
  if (mode == 0) {
    // step(action): the per-environment prologue first, at full width (csrc/shipenv_kernels.cuh k_prologue)
    // (it also restarts the work-queue counters of the env kernel)
    SenvView pv = view(h);
    if (split) {
      pv.call_filter = 1;                    // the prologue counts the environments of the second launch in queue[2]
      CUDA_TRY(cudaMemsetAsync(h->queue_dev + 2, 0, 2 * sizeof(unsigned long long), st));   // + the second launch's work index
    }
    cudaError_t pe = (h->params.math_mode == SHIPENV_MATH_FAST)
                         ? senv_fast::launch_prologue(pv, h->params.env_kind, actions, h->queue_dev, st)
                         : senv_strict::launch_prologue(pv, h->params.env_kind, actions, h->queue_dev, st);
    if (pe != cudaSuccess) return fail(SHIPENV_E_CUDA, "prologue kernel launch: %s", cudaGetErrorString(pe));
  } else {
    CUDA_TRY(cudaMemsetAsync(h->queue_dev, 0, 2 * sizeof(unsigned long long), st));
  }
  // (A static grid -- one slot per environment -- for launches of a few _step() was measured and is not the default:
  //  one step per launch reaches 0.42 / 0.79 of the HBM peak at 1e5 / 1e6 environments with the persistent grid, 0.39 /
  //  0.68 with the static one.  SHIPENV_STREAMING_K=k selects it for launches of at most k steps.)
  if (mode == 1 && k <= h->streaming_k) persistent = 0;
  if (h->time_kernels) {
    if (h->kernel_pending) {                 // fold the previous launch into the sum before reusing the events
      float ms = 0.f;
      if (cudaEventSynchronize(h->ev_k1) == cudaSuccess && cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1) == cudaSuccess)
        h->kernel_ms_sum += ms;
    }
    CUDA_TRY(cudaEventRecord(h->ev_k0, st));
  }
  SenvView v = view(h);
  // quiet steps pay for themselves over a run of steps; a launch of a few steps evaluates every test at every step
  if (mode == 1 && k < h->quiet_min_k) v.no_quiet = 1;
  auto launch = [&](const SenvView& w) {
    return (h->params.math_mode == SHIPENV_MATH_FAST)
               ? senv_fast::launch_env(w, model, h->params.env_kind, mode, actions, k, h->queue_dev, h->sm_count,
                                       persistent, 0, st)
               : senv_strict::launch_env(w, model, h->params.env_kind, mode, actions, k, h->queue_dev, h->sm_count,
                                         persistent, 0, st);
  };
  cudaError_t e = cudaSuccess;
  if (split) {
    // step(action) as two launches (SenvView::call_filter): environments in the last call of their episode go to the
    // kernel with quiet steps, the others to its twin; each launch has its own work-queue index, the count of
    // finished environments is shared
    SenvView a = v, b = v;
    a.no_quiet = 1; a.call_filter = 1;
    b.call_filter = 2;
    e = launch(a);
    if (e == cudaSuccess) e = launch(b);
  } else {
    e = launch(v);
  }
  if (e != cudaSuccess) return fail(SHIPENV_E_CUDA, "env kernel launch: %s", cudaGetErrorString(e));
  if (h->time_kernels) {
    CUDA_TRY(cudaEventRecord(h->ev_k1, st));
    h->kernel_pending = true;
  }
  // only the automatic grid policy reads the count of finished environments back
  if (h->persist_mode < 0)
    CUDA_TRY(cudaMemcpyAsync(h->done_host, h->queue_dev + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  return note_caller_stream(h, st);
}

// ---- culling grid ------------------------------------------------------------------------------
bool host_poly_contains(const ShipEnvParams& p, int poly, double x, double y) {
  const int a = p.poly_start[poly], b = p.poly_start[poly + 1];
  bool inside = false;
  for (int i = a, j = b - 1; i < b; j = i++) {
    const double xi = p.vert_e[i], yi = p.vert_n[i], xj = p.vert_e[j], yj = p.vert_n[j];
    if ((yi > y) != (yj > y)) {
      if (x < (xj - xi) * (y - yi) / (yj - yi) + xi) inside = !inside;
    }
  }
  return inside;
}

double host_ring_distance(const ShipEnvParams& p, int poly, double x, double y) {
  const int a = p.poly_start[poly], b = p.poly_start[poly + 1];
  double best = INFINITY;
  for (int i = a; i < b; ++i) {
    const int k = (i + 1 < b) ? i + 1 : a;
    const double ax = p.vert_e[i], ay = p.vert_n[i], dx = p.vert_e[k] - ax, dy = p.vert_n[k] - ay;
    const double l2 = dx * dx + dy * dy;
    double t = l2 > 0 ? ((x - ax) * dx + (y - ay) * dy) / l2 : 0.0;
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    best = std::fmin(best, std::hypot(x - (ax + t * dx), y - (ay + t * dy)));
  }
  return best;
}

// A polygon can only contain a corner of the ship square (centre in the cell, corners at most
// margin*sqrt(2) from the centre) if the cell centre is inside it or its ring is within
// halfdiag + margin*sqrt(2) of the cell centre (the ring distance is 1-Lipschitz); likewise a ring
// can only be within `clip` of a point of the cell if it is within clip + halfdiag of the centre.
int build_grid(shipenv* h) {
  const ShipEnvParams& p = h->params;
  const double w = p.map_max_e - p.map_min_e, ht = p.map_max_n - p.map_min_n;
  int nx = 1, ny = 1;
  double cell = 1.0;
  if (w > 0 && ht > 0) {   // (without polygons the masks are empty; the safe radius still knows the map horizon)
    cell = std::fmax(std::fmax(w, ht) / 256.0, 50.0);
    nx = (int)std::ceil(w / cell);
    ny = (int)std::ceil(ht / cell);
  }
  std::vector<unsigned> cells((size_t)nx * ny, 0u);
  std::vector<unsigned long long> edges((size_t)nx * ny * 2, 0ull);
  const int n_vert = p.n_poly > 0 ? p.poly_start[p.n_poly] : 0;
  // segment i: vertex i -> next vertex of its polygon
  std::vector<int> seg_next(n_vert, 0);
  for (int q = 0; q < p.n_poly; ++q)
    for (int i = p.poly_start[q]; i < p.poly_start[q + 1]; ++i)
      seg_next[i] = (i + 1 < p.poly_start[q + 1]) ? i + 1 : p.poly_start[q];
  auto seg_dist = [&](int i, double x, double y) {
    const int k = seg_next[i];
    const double ax = p.vert_e[i], ay = p.vert_n[i], dx = p.vert_e[k] - ax, dy = p.vert_n[k] - ay;
    const double l2 = dx * dx + dy * dy;
    double t = l2 > 0 ? ((x - ax) * dx + (y - ay) * dy) / l2 : 0.0;
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    return std::hypot(x - (ax + t * dx), y - (ay + t * dy));
  };
  const double half_len = 0.5 * std::fmax(p.ship[0].l_ship, p.ship[1].l_ship);
  const double halfdiag = cell * 0.70710678118654757 + 1e-6 * cell;
  const double r_contains = halfdiag + half_len * 1.4142135623730951 + 1.0;
  const double r_dist = 1000.0 + halfdiag + 1.0;
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
      const double cx = p.map_min_e + (ix + 0.5) * cell, cy = p.map_min_n + (iy + 0.5) * cell;
      unsigned m = 0;
      for (int q = 0; q < p.n_poly && q < 16; ++q) {
        const double d = host_ring_distance(p, q, cx, cy);
        if (d <= r_contains || host_poly_contains(p, q, cx, cy)) m |= 1u << q;
        if (d <= r_dist) m |= 1u << (16 + q);
      }
      cells[(size_t)iy * nx + ix] = m;
      // ring segments that can be the nearest one to a point of the cell: the distance to a segment is
      // convex, so its maximum over the cell is attained at a corner (upper bound U = min over segments of
      // that maximum) and it is 1-Lipschitz (lower bound = distance at the centre - half diagonal)
      const double x0 = p.map_min_e + ix * cell, x1 = x0 + cell, y0 = p.map_min_n + iy * cell, y1 = y0 + cell;
      double U = INFINITY;
      for (int i = 0; i < n_vert; ++i) {
        const double mx = std::fmax(std::fmax(seg_dist(i, x0, y0), seg_dist(i, x1, y0)),
                                    std::fmax(seg_dist(i, x0, y1), seg_dist(i, x1, y1)));
        U = std::fmin(U, mx);
      }
      const double slack = 1e-6 * cell + 1e-3;
      for (int i = 0; i < n_vert && i < 128; ++i) {
        const double lo = seg_dist(i, cx, cy) - halfdiag;
        if (lo <= U + slack && lo <= 1000.0 + slack)
          edges[((size_t)iy * nx + ix) * 2 + (i >> 6)] |= 1ull << (i & 63);
      }
    }
  // safe radius per cell (SenvGrid::safe): the ring distance is 1-Lipschitz, so every point of the cell is at least
  // d(centre) - half diagonal from every ring; a ship whose centre is outside every polygon and further than half a
  // diagonal of its L x L square from every ring has no corner inside one
  std::vector<float> safe((size_t)nx * ny, 0.f);
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
      const double cx = p.map_min_e + (ix + 0.5) * cell, cy = p.map_min_n + (iy + 0.5) * cell;
      double dmin = 1e9;
      bool in_poly = false;
      for (int q = 0; q < p.n_poly; ++q) {
        dmin = std::fmin(dmin, host_ring_distance(p, q, cx, cy));
        in_poly = in_poly || host_poly_contains(p, q, cx, cy);
      }
      const double r = 0.999 * (dmin - halfdiag - half_len * 1.4142135623730951 - 1.0);
      safe[(size_t)iy * nx + ix] = (!in_poly && r > 0.0) ? (float)r : 0.f;
    }
  // allocate and fill the new tables first; the handle keeps its old ones if anything fails
  unsigned* cells_dev = nullptr;
  unsigned long long* edges_dev = nullptr;
  float* safe_dev = nullptr;
  cudaError_t e = cudaMalloc(&cells_dev, cells.size() * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMalloc(&safe_dev, safe.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(safe_dev, safe.data(), safe.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&edges_dev, edges.size() * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpy(cells_dev, cells.data(), cells.size() * sizeof(unsigned), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(edges_dev, edges.data(), edges.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(cells_dev);
    cudaFree(edges_dev);
    cudaFree(safe_dev);
    return fail(SHIPENV_E_CUDA, "uploading the map culling grid: %s", cudaGetErrorString(e));
  }
  cudaFree(h->grid_dev);
  cudaFree(h->edges_dev);
  cudaFree(h->safe_dev);
  h->grid_dev = cells_dev;
  h->edges_dev = edges_dev;
  h->safe_dev = safe_dev;
  h->grid = SenvGrid{h->grid_dev, h->edges_dev, h->safe_dev, p.map_min_e, p.map_min_n, 1.0 / cell, (double)nx, (double)ny, nx, ny};
  return SHIPENV_OK;
}

// (re)build the kernels' shared-memory block from the uploaded parameters, with the device code of the handle's build
int build_staged(shipenv* h) {
  if (!h->staged_dev) CUDA_TRY(cudaMalloc(&h->staged_dev, senv_fast::staged_bytes()));
  CUDA_TRY((h->params.math_mode == SHIPENV_MATH_FAST)
               ? senv_fast::launch_build_staged(h->params_dev, h->staged_dev, nullptr)
               : senv_strict::launch_build_staged(h->params_dev, h->staged_dev, nullptr));
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  return SHIPENV_OK;
}

// DFMA peak microbenchmark (roofline denominator)
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

int ensure_staging(shipenv* h) {
  if (h->pinned) {
    if (h->caller_pending) {                // order this call after what the caller submitted on its own stream
      CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_caller, 0));
      h->caller_pending = false;
    }
    return SHIPENV_OK;
  }
  const size_t B = (size_t)h->num_envs;
  // actions f64 | obs 8 x f32 | reward f64 | info i32 | nsub i32 | mask u8
  h->pinned_bytes = B * (8 + 32 + 8 + 4 + 4 + 1) + 64;
  CUDA_TRY(cudaMallocHost(&h->pinned, h->pinned_bytes));
  CUDA_TRY(cudaMalloc(&h->act_dev, B * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->mask_dev, B));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  if (h->caller_pending) {
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_caller, 0));
    h->caller_pending = false;
  }
  return SHIPENV_OK;
}

struct Pinned {
  double* actions; float* obs; double* reward; int32_t* info; int32_t* nsub; uint8_t* mask;
};

Pinned pinned_view(shipenv* h) {
  const size_t B = (size_t)h->num_envs;
  char* p = (char*)h->pinned;
  Pinned v;
  v.actions = (double*)p; p += B * 8;
  v.reward = (double*)p; p += B * 8;
  v.obs = (float*)p; p += B * 32;
  v.info = (int32_t*)p; p += B * 4;
  v.nsub = (int32_t*)p; p += B * 4;
  v.mask = (uint8_t*)p;
  return v;
}

// Is [ptr, ptr + bytes) inside a range the caller page-locked through shipenv_register_host?  Only then do the
// *_host copies go straight between the caller's memory and the device; every other buffer is staged through the
// handle's own pinned area.  (The library never registers caller memory on its own: a registration outlives the
// array it was made for, and a later allocation at the same address would be DMA'd through a stale mapping.)
bool host_locked(shipenv* h, const void* ptr, size_t bytes) {
  const char* p = static_cast<const char*>(ptr);
  for (const auto& r : h->host_regs) {
    const char* b = static_cast<const char*>(r.ptr);
    if (p >= b && p + bytes <= b + r.bytes) return true;
  }
  return false;
}

int fetch_outputs(shipenv* h, float* obs_host, double* reward_host, int32_t* info_host, int32_t* nsub_host) {
  const size_t B = (size_t)h->num_envs;
  Pinned pv = pinned_view(h);
  // One DMA instead of four when the outputs lie back to back (obs | reward | info | nsub) on both sides -- the layout
  // ast_sac_b200.env allocates: a D2H copy of 48 B per environment is bandwidth-bound, four small ones are
  // latency-bound as well.
  {
    const char* d0 = reinterpret_cast<const char*>(h->buf.obs_f32);
    char* h0 = reinterpret_cast<char*>(obs_host);
    const bool packed_dev = (const char*)h->buf.reward == d0 + B * 32 && (const char*)h->buf.info_i32 == d0 + B * 40 &&
                            (const char*)h->buf.nsub_i32 == d0 + B * 44;
    const bool packed_host = obs_host && (char*)reward_host == h0 + B * 32 && (char*)info_host == h0 + B * 40 &&
                             (char*)nsub_host == h0 + B * 44;
    if (packed_dev && packed_host && host_locked(h, obs_host, B * 48)) {
      CUDA_TRY(cudaMemcpyAsync(obs_host, h->buf.obs_f32, B * 48, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      return SHIPENV_OK;
    }
  }
  const bool d_obs = obs_host && host_locked(h, obs_host, B * 32);
  const bool d_rew = reward_host && host_locked(h, reward_host, B * 8);
  const bool d_info = info_host && host_locked(h, info_host, B * 4);
  const bool d_nsub = nsub_host && host_locked(h, nsub_host, B * 4);
  if (obs_host) CUDA_TRY(cudaMemcpyAsync(d_obs ? obs_host : pv.obs, h->buf.obs_f32, B * 32, cudaMemcpyDeviceToHost, h->stream));
  if (reward_host)
    CUDA_TRY(cudaMemcpyAsync(d_rew ? reward_host : pv.reward, h->buf.reward, B * 8, cudaMemcpyDeviceToHost, h->stream));
  if (info_host)
    CUDA_TRY(cudaMemcpyAsync(d_info ? info_host : pv.info, h->buf.info_i32, B * 4, cudaMemcpyDeviceToHost, h->stream));
  if (nsub_host)
    CUDA_TRY(cudaMemcpyAsync(d_nsub ? nsub_host : pv.nsub, h->buf.nsub_i32, B * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (obs_host && !d_obs) memcpy(obs_host, pv.obs, B * 32);
  if (reward_host && !d_rew) memcpy(reward_host, pv.reward, B * 8);
  if (info_host && !d_info) memcpy(info_host, pv.info, B * 4);
  if (nsub_host && !d_nsub) memcpy(nsub_host, pv.nsub, B * 4);
  return SHIPENV_OK;
}

}  // namespace

extern "C" {

int shipenv_abi_version(void) { return SHIPENV_ABI_VERSION; }
int shipenv_sizeof_params(void) { return (int)sizeof(ShipEnvParams); }
const char* shipenv_last_error(void) { return g_err; }

int shipenv_create(const ShipEnvParams* params, int64_t num_envs, int device, shipenv_t** out) {
  if (!out) return fail(SHIPENV_E_ARG, "out is NULL");
  *out = nullptr;
  int rc = validate(params, num_envs);
  if (rc) return rc;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(SHIPENV_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(SHIPENV_E_ARG, "device %d out of range [0, %d)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  shipenv* h = new (std::nothrow) shipenv();
  if (!h) return fail(SHIPENV_E_NOMEM, "out of host memory");
  h->device = device;
  h->num_envs = num_envs;
  h->params = *params;
  cudaError_t ce = cudaMalloc(&h->params_dev, sizeof(ShipEnvParams));
  if (ce == cudaSuccess) ce = cudaMemcpy(h->params_dev, params, sizeof(ShipEnvParams), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    delete h;
    return fail(SHIPENV_E_CUDA, "uploading parameters: %s", cudaGetErrorString(ce));
  }
  rc = build_grid(h);
  if (rc == SHIPENV_OK) rc = build_staged(h);
  if (rc == SHIPENV_OK && (cudaMalloc(&h->queue_dev, 4 * sizeof(unsigned long long)) != cudaSuccess ||
                           cudaMallocHost(&h->done_host, sizeof(unsigned long long)) != cudaSuccess))
    rc = fail(SHIPENV_E_CUDA, "allocating the work-queue counters failed");
  if (rc) {
    cudaFree(h->params_dev);
    cudaFree(h->staged_dev);
    cudaFree(h->grid_dev);
    cudaFree(h->edges_dev);
    cudaFree(h->safe_dev);
    cudaFree(h->queue_dev);
    if (h->done_host) cudaFreeHost(h->done_host);
    cudaGetLastError();
    delete h;
    return rc;
  }
  *h->done_host = 0;
  if (const char* pm = getenv("SHIPENV_PERSISTENT")) h->persist_mode = atoi(pm);   // 1 (default), 0, -1 auto
  if (const char* sk = getenv("SHIPENV_STREAMING_K")) h->streaming_k = atoi(sk);    // measurement aid
  if (const char* q = getenv("SHIPENV_QUIET")) h->no_quiet = (atoi(q) == 0) ? 1 : 0;  // comparison aid (tests)
  if (const char* q = getenv("SHIPENV_QUIET_MIN_K")) h->quiet_min_k = atoi(q);         // measurement aid
  if (const char* q = getenv("SHIPENV_SPLIT_CALLS")) h->split_calls = atoi(q);         // measurement aid
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = h;
  return SHIPENV_OK;
}

int shipenv_destroy(shipenv_t* h) {
  if (!h) return SHIPENV_OK;
  cudaSetDevice(h->device);
  if (h->owns) {
    cudaFree(h->buf.ship_f64); cudaFree(h->buf.ship_i32); cudaFree(h->buf.env_f64); cudaFree(h->buf.env_i32);
    cudaFree(h->buf.iw_f64); cudaFree(h->buf.prev_f32); cudaFree(h->buf.obs_f32); cudaFree(h->buf.reward);
    cudaFree(h->buf.info_i32); cudaFree(h->buf.nsub_i32); cudaFree(h->buf.counters);
  }
  cudaFree(h->params_dev);
  cudaFree(h->staged_dev);
  cudaFree(h->grid_dev);
  cudaFree(h->edges_dev);
  cudaFree(h->safe_dev);
  cudaFree(h->queue_dev);
  if (h->done_host) cudaFreeHost(h->done_host);
  cudaFree(h->act_dev);
  cudaFree(h->mask_dev);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev_k0) { cudaEventDestroy(h->ev_k0); cudaEventDestroy(h->ev_k1); }
  if (h->ev_caller) cudaEventDestroy(h->ev_caller);
  for (const auto& r : h->host_regs)
    if (r.owned) cudaHostUnregister(const_cast<void*>(r.ptr));
  cudaGetLastError();   // a range the caller already freed is not an error of destroy
  delete h;
  return SHIPENV_OK;
}

int shipenv_layout(const shipenv_t* h, ShipEnvLayout* out) {
  if (!h || !out) return fail(SHIPENV_E_ARG, "NULL argument");
  const int64_t B = h->num_envs;
  out->ship_f64 = (int64_t)SHIPENV_SF_COUNT * 2 * B;
  out->ship_i32 = 2 * B;
  out->env_f64 = (int64_t)SHIPENV_EF_COUNT * B;
  out->env_i32 = (int64_t)SHIPENV_EI_COUNT * B;
  out->iw_f64 = (int64_t)2 * SHIPENV_MAX_IW * B;
  out->prev_f32 = 4 * B;
  out->obs_f32 = 8 * B;
  out->reward = B;
  out->info_i32 = B;
  out->nsub_i32 = B;
  out->counters = 4;
  return SHIPENV_OK;
}

int shipenv_bind(shipenv_t* h, const ShipEnvBuffers* b) {
  if (!h || !b) return fail(SHIPENV_E_ARG, "NULL argument");
  if (h->owns) return fail(SHIPENV_E_STATE, "buffers already allocated by shipenv_alloc");
  if (!b->ship_f64 || !b->ship_i32 || !b->env_f64 || !b->env_i32 || !b->iw_f64 || !b->prev_f32 || !b->obs_f32 ||
      !b->reward || !b->info_i32 || !b->nsub_i32)
    return fail(SHIPENV_E_ARG, "every buffer except counters must be non-NULL");
  if (((uintptr_t)b->obs_f32) % 16) return fail(SHIPENV_E_ARG, "obs_f32 must be 16-byte aligned");
  h->buf = *b;
  h->bound = true;
  h->constructed = false;
  return SHIPENV_OK;
}

int shipenv_alloc(shipenv_t* h) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  if (h->bound) return fail(SHIPENV_E_STATE, "buffers already bound");
  CUDA_TRY(cudaSetDevice(h->device));
  ShipEnvLayout L;
  shipenv_layout(h, &L);
  ShipEnvBuffers b{};
  cudaError_t e = cudaMalloc(&b.ship_f64, L.ship_f64 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&b.ship_i32, L.ship_i32 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b.env_f64, L.env_f64 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&b.env_i32, L.env_i32 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b.iw_f64, L.iw_f64 * 8);
  if (e == cudaSuccess) e = cudaMalloc(&b.prev_f32, L.prev_f32 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b.obs_f32, L.obs_f32 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b.reward, L.reward * 8);
  if (e == cudaSuccess) e = cudaMalloc(&b.info_i32, L.info_i32 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b.nsub_i32, L.nsub_i32 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b.counters, L.counters * 8);
  if (e == cudaSuccess) e = cudaMemset(b.counters, 0, L.counters * 8);
  if (e == cudaSuccess) e = cudaMemset(b.iw_f64, 0, L.iw_f64 * 8);
  if (e != cudaSuccess) {
    cudaFree(b.ship_f64); cudaFree(b.ship_i32); cudaFree(b.env_f64); cudaFree(b.env_i32); cudaFree(b.iw_f64);
    cudaFree(b.prev_f32); cudaFree(b.obs_f32); cudaFree(b.reward); cudaFree(b.info_i32); cudaFree(b.nsub_i32);
    cudaFree(b.counters);
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? SHIPENV_E_NOMEM : SHIPENV_E_CUDA, "allocating the state buffers: %s",
                cudaGetErrorString(e));
  }
  h->buf = b;
  h->bound = true;
  h->owns = true;
  return SHIPENV_OK;
}

int shipenv_buffers(const shipenv_t* h, ShipEnvBuffers* out) {
  if (!h || !out) return fail(SHIPENV_E_ARG, "NULL argument");
  if (!h->bound) return fail(SHIPENV_E_STATE, "no buffers bound");
  *out = h->buf;
  return SHIPENV_OK;
}

int shipenv_set_params(shipenv_t* h, const ShipEnvParams* params) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  int rc = validate(params, h->num_envs);
  if (rc) return rc;
  if (h->constructed) {
    // the state buffers hold environments of one model class / env class / route length: a parameter update may
    // change numbers (gains, time steps, the dt_shaft of the reset quirk, the map), not what the state means
    if (params->ship[0].model_kind != h->params.ship[0].model_kind || params->env_kind != h->params.env_kind ||
        params->ship[0].n_wp != h->params.ship[0].n_wp || params->ship[1].n_wp != h->params.ship[1].n_wp ||
        params->max_sampling_frequency != h->params.max_sampling_frequency || params->math_mode != h->params.math_mode ||
        params->obs_sampled_route != h->params.obs_sampled_route)
      return fail(SHIPENV_E_STATE, "model_kind, env_kind, route lengths, max_sampling_frequency, obs_sampled_route and math_mode cannot "
                                   "change once the environments are constructed: create a new handle");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaDeviceSynchronize());   // kernels in flight on any stream may still read the old block
  const ShipEnvParams old = h->params;
  h->params = *params;
  rc = build_grid(h);                  // (keeps the old grid when it fails)
  if (rc) { h->params = old; return rc; }
  CUDA_TRY(cudaMemcpy(h->params_dev, params, sizeof(ShipEnvParams), cudaMemcpyHostToDevice));
  return build_staged(h);
}

int shipenv_construct(shipenv_t* h, const double* init_dev, void* stream) {
  int rc = check_ready(h, false);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_reset(h, nullptr, init_dev, 0, 1, st);
  if (rc) return rc;
  CUDA_TRY(senv_strict::launch_init_prev(view(h), st));
  h->constructed = true;
  h->init_dev = init_dev;
  return SHIPENV_OK;
}

int shipenv_reset(shipenv_t* h, const uint8_t* mask_dev, const double* init_dev, void* stream) {
  int rc = check_ready(h, false);
  if (rc) return rc;
  if (!h->constructed) {
    if (mask_dev) return fail(SHIPENV_E_STATE, "the first reset must cover all environments (mask = NULL)");
    rc = shipenv_construct(h, init_dev, stream);
    if (rc) return rc;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!mask_dev) *h->done_host = 0;
  return launch_reset(h, mask_dev, init_dev, 1, 1, st);
}

int shipenv_init_step(shipenv_t* h, void* stream) {
  int rc = check_ready(h, true);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return launch_reset(h, nullptr, nullptr, 1, 0, st);
}

int shipenv_step(shipenv_t* h, const double* actions_dev, void* stream) {
  int rc = check_ready(h, true);
  if (rc) return rc;
  if (h->params.env_kind == SHIPENV_ENV_COLAV_NONIW && !h->params.obs_sampled_route)
    return fail(SHIPENV_E_STATE, "step(action) on the NonIW env kind needs params.obs_sampled_route (an obstacle ship "
                                 "with a HeadingBySampledRouteController)");
  if (!actions_dev) return fail(SHIPENV_E_ARG, "actions_dev is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return launch_env(h, 0, actions_dev, 0, st);
}

int shipenv_substeps(shipenv_t* h, int k, void* stream) {
  int rc = check_ready(h, true);
  if (rc) return rc;
  if (k < 0) return fail(SHIPENV_E_ARG, "k must be >= 0");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  return launch_env(h, 1, nullptr, k, st);
}

int shipenv_ship_rollout(shipenv_t* h, int k, void* stream) {
  int rc = check_ready(h, true);
  if (rc) return rc;
  if (k < 0) return fail(SHIPENV_E_ARG, "k must be >= 0");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int model = h->params.ship[0].model_kind;
  CUDA_TRY((h->params.math_mode == SHIPENV_MATH_FAST) ? senv_fast::launch_rollout(view(h), model, k, st)
                                                      : senv_strict::launch_rollout(view(h), model, k, st));
  return note_caller_stream(h, st);
}

int shipenv_reset_host(shipenv_t* h, const uint8_t* mask_host, float* obs_host) {
  int rc = check_ready(h, false);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  rc = ensure_staging(h);
  if (rc) return rc;
  const uint8_t* mask_dev = nullptr;
  if (mask_host) {
    Pinned pv = pinned_view(h);
    memcpy(pv.mask, mask_host, (size_t)h->num_envs);
    CUDA_TRY(cudaMemcpyAsync(h->mask_dev, pv.mask, (size_t)h->num_envs, cudaMemcpyHostToDevice, h->stream));
    mask_dev = h->mask_dev;
  }
  rc = shipenv_reset(h, mask_dev, h->init_dev, h->stream);
  if (rc) return rc;
  return fetch_outputs(h, obs_host, nullptr, nullptr, nullptr);
}

int shipenv_step_host(shipenv_t* h, const double* actions_host, float* obs_host, double* reward_host,
                      int32_t* info_host, int32_t* nsub_host) {
  int rc = check_ready(h, true);
  if (rc) return rc;
  if (!actions_host) return fail(SHIPENV_E_ARG, "actions_host is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  rc = ensure_staging(h);
  if (rc) return rc;
  Pinned pv = pinned_view(h);
  const void* src = actions_host;
  if (!host_locked(h, actions_host, (size_t)h->num_envs * 8)) {
    memcpy(pv.actions, actions_host, (size_t)h->num_envs * 8);
    src = pv.actions;
  }
  CUDA_TRY(cudaMemcpyAsync(h->act_dev, src, (size_t)h->num_envs * 8, cudaMemcpyHostToDevice, h->stream));
  rc = shipenv_step(h, h->act_dev, h->stream);
  if (rc) return rc;
  return fetch_outputs(h, obs_host, reward_host, info_host, nsub_host);
}

int shipenv_substeps_host(shipenv_t* h, int k, float* obs_host, double* reward_host, int32_t* info_host,
                          int32_t* nsub_host) {
  int rc = check_ready(h, true);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(h->device));
  rc = ensure_staging(h);
  if (rc) return rc;
  rc = shipenv_substeps(h, k, h->stream);
  if (rc) return rc;
  return fetch_outputs(h, obs_host, reward_host, info_host, nsub_host);
}

int shipenv_register_host(shipenv_t* h, void* ptr, size_t bytes) {
  if (!h || !ptr || bytes == 0) return fail(SHIPENV_E_ARG, "bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  for (const auto& r : h->host_regs)
    if (r.ptr == ptr && r.bytes >= bytes) return SHIPENV_OK;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
  bool owned = true;
  if (e == cudaErrorHostMemoryAlreadyRegistered) {   // page-locked by someone else (e.g. torch pin_memory): usable, not ours
    cudaGetLastError();
    owned = false;
  } else if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SHIPENV_E_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
  }
  h->host_regs.push_back({ptr, bytes, owned});
  return SHIPENV_OK;
}

int shipenv_unregister_host(shipenv_t* h, void* ptr) {
  if (!h || !ptr) return fail(SHIPENV_E_ARG, "bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  for (size_t i = 0; i < h->host_regs.size(); ++i)
    if (h->host_regs[i].ptr == ptr) {
      if (h->stream) cudaStreamSynchronize(h->stream);     // no copy of ours is still using the range
      if (h->host_regs[i].owned) cudaHostUnregister(ptr);
      cudaGetLastError();
      h->host_regs.erase(h->host_regs.begin() + (long)i);
      return SHIPENV_OK;
    }
  return fail(SHIPENV_E_ARG, "range was not registered through shipenv_register_host");
}

int shipenv_set_trajectory_log(shipenv_t* h, double* log_dev, int32_t* count_dev, int64_t log_envs, int64_t capacity) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  if (!log_dev || !count_dev || log_envs <= 0 || capacity <= 0) {
    h->log_dev = nullptr; h->log_count_dev = nullptr; h->log_envs = 0; h->log_capacity = 0;
    return SHIPENV_OK;
  }
  if (log_envs > h->num_envs) return fail(SHIPENV_E_ARG, "log_envs %lld > num_envs %lld", (long long)log_envs, h->num_envs);
  h->log_dev = log_dev; h->log_count_dev = count_dev; h->log_envs = log_envs; h->log_capacity = capacity;
  return SHIPENV_OK;
}

int shipenv_time_env_kernel(shipenv_t* h, int enable) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  if (enable && !h->ev_k0) {
    CUDA_TRY(cudaEventCreate(&h->ev_k0));
    CUDA_TRY(cudaEventCreate(&h->ev_k1));
  }
  h->time_kernels = enable ? 1 : 0;
  h->kernel_ms_sum = 0.0;
  h->kernel_pending = false;
  return SHIPENV_OK;
}

int shipenv_env_kernel_ms(shipenv_t* h, double* ms_out) {
  if (!h || !ms_out) return fail(SHIPENV_E_ARG, "NULL argument");
  if (!h->time_kernels) return fail(SHIPENV_E_STATE, "kernel timing is off (shipenv_time_env_kernel)");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->kernel_pending) {
    float ms = 0.f;
    CUDA_TRY(cudaEventSynchronize(h->ev_k1));
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1));
    h->kernel_ms_sum += ms;
    h->kernel_pending = false;
  }
  *ms_out = h->kernel_ms_sum;
  h->kernel_ms_sum = 0.0;
  return SHIPENV_OK;
}

int shipenv_read_counters(shipenv_t* h, unsigned long long* out_host) {
  int rc = check_ready(h, false);
  if (rc) return rc;
  if (!out_host) return fail(SHIPENV_E_ARG, "out_host is NULL");
  if (!h->buf.counters) return fail(SHIPENV_E_STATE, "no counters buffer bound");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpy(out_host, h->buf.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return SHIPENV_OK;
}

int shipenv_map_query(shipenv_t* h, int64_t n, const double* north_dev, const double* east_dev, double ship_length,
                      int32_t* contains_dev, int32_t* square_dev, double* distance_dev, void* stream) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  if (n <= 0 || !north_dev || !east_dev || !contains_dev || !square_dev || !distance_dev)
    return fail(SHIPENV_E_ARG, "bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY((h->params.math_mode == SHIPENV_MATH_FAST)
               ? senv_fast::launch_map_query(view(h), n, north_dev, east_dev, ship_length, contains_dev, square_dev,
                                             distance_dev, st)
               : senv_strict::launch_map_query(view(h), n, north_dev, east_dev, ship_length, contains_dev, square_dev,
                                               distance_dev, st));
  return SHIPENV_OK;
}

int shipenv_map_safe_radius(shipenv_t* h, int64_t n, const double* north_dev, const double* east_dev, float* out_dev,
                            void* stream) {
  if (!h) return fail(SHIPENV_E_ARG, "handle is NULL");
  if (n <= 0 || !north_dev || !east_dev || !out_dev) return fail(SHIPENV_E_ARG, "bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY((h->params.math_mode == SHIPENV_MATH_FAST)
               ? senv_fast::launch_map_safe_radius(view(h), n, north_dev, east_dev, out_dev, st)
               : senv_strict::launch_map_safe_radius(view(h), n, north_dev, east_dev, out_dev, st));
  return SHIPENV_OK;
}

int shipenv_selftest_math(int device, int64_t n, uint64_t seed, unsigned long long* mismatches_host) {
  if (!mismatches_host || n <= 0) return fail(SHIPENV_E_ARG, "bad arguments");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return fail(SHIPENV_E_CUDA, "no such CUDA device %d", device);
  CUDA_TRY(cudaSetDevice(device));
  unsigned long long* dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, 18 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(dev, 0, 18 * sizeof(unsigned long long)));
  CUDA_TRY(senv_fast::launch_math_selftest(n, seed, dev, nullptr));
  CUDA_TRY(senv_strict::launch_math_selftest(n, seed, dev + 9, nullptr));
  CUDA_TRY(cudaMemcpy(mismatches_host, dev, 18 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(dev);
  return SHIPENV_OK;
}

int shipenv_measure_fp64_peak(int device, int repeats, double* tflops_out) {
  if (!tflops_out) return fail(SHIPENV_E_ARG, "tflops_out is NULL");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return fail(SHIPENV_E_CUDA, "no such CUDA device %d", device);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, 8));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const int iters = 1 << 16, threads = 256, blocks = prop.multiProcessorCount * 8;
  double best_ms = 1e30;
  if (repeats < 1) repeats = 1;
  for (int r = 0; r < repeats + 1; ++r) {          // first launch is the warm-up
    CUDA_TRY(cudaEventRecord(e0));
    k_dfma_peak<<<blocks, threads>>>(out, iters, 0.9999999, 1e-7);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best_ms) best_ms = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
  *tflops_out = flops / (best_ms * 1e-3) / 1e12;
  return SHIPENV_OK;
}

}  // extern "C"
