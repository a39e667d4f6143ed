// shipenv_launch.h -- internal interface between the host side of the C ABI (shipenv.cu) and the two
// builds of the device code (kernels_strict.cu: -fmad=false, kernels_fast.cu: FMA contraction).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "shipenv.h"

// Culling grid over the map bounding box (built on the host at create time, read-only, lives in
// global memory and stays L1/L2 resident): one 32-bit word per cell; bits 0..15 = polygons that can
// contain a corner of a ship whose centre lies in the cell, bits 16..31 = polygons whose ring can be
// within the reward's 1000 m clipping distance of a point in the cell.  Conservative, so the exact
// tests give the same answers as testing every polygon.
struct SenvGrid {
  const unsigned* cells;
  // per cell, 128-bit mask over the map's ring segments (segment i runs from vertex i to the next vertex
  // of its polygon): the segments that can be the nearest one, within the 1000 m clip, to some point of
  // the cell.  edges[2*c] = segments 0..63, edges[2*c+1] = segments 64..127.
  const unsigned long long* edges;
  // per cell, the safe radius in metres: a ship whose centre is closer than this to ANY point of the cell has no
  // corner of its L x L square inside a polygon (no grounding); 0 in and next to polygons.
  const float* safe;
  double e0, n0, inv_cell;
  double nx_f, ny_f;   // nx, ny as doubles (the bounds test of the cell lookup runs at every simulator step)
  int nx, ny;
};

struct SenvView {
  const ShipEnvParams* params;
  const void* staged;   // the kernels' shared-memory block, built once per parameter upload (launch_build_staged)
  ShipEnvBuffers buf;
  long long num_envs;
  SenvGrid grid;
  int collav;  // params->collav (selects the kernel instantiation)
  int sm_count;  // multiprocessors of the device
  int no_quiet;  // 1: launch the env kernel's twin that evaluates every event test at every step (see k_env, quiet steps)
  // step(action) launches only -- 0: every environment; 1: only environments NOT in the last step() call of their
  // episode (sampling count below the maximum); 2: only those that are.  The last call runs the environment to its end
  // (hundreds of steps: quiet steps pay), the others end after ~85 steps at the next waypoint (they do not), so the host
  // issues the call as two launches, one per twin; an environment is taken by exactly one of them.
  int call_filter;
  // optional trajectory log (shipenv_set_trajectory_log)
  double* log_f64;
  int32_t* log_count;
  long long log_envs, log_capacity;
};

#define SENV_DECLARE(ns)                                                                                       \
  namespace ns {                                                                                               \
  cudaError_t launch_reset(const SenvView& v, int model, const uint8_t* mask, const double* init, int do_init, \
                           int reinit, cudaStream_t st);                                                       \
  cudaError_t launch_init_prev(const SenvView& v, cudaStream_t st);                                            \
  cudaError_t launch_prologue(const SenvView& v, int env_kind, const double* actions,                          \
                              unsigned long long* queue, cudaStream_t st);                                     \
  cudaError_t launch_env(const SenvView& v, int model, int env_kind, int mode, const double* actions, int k,   \
                         unsigned long long* queue, int sm_count, int persistent, int clear_queue,           \
                         cudaStream_t st);                                                                     \
  cudaError_t launch_rollout(const SenvView& v, int model, int k, cudaStream_t st);                            \
  size_t staged_bytes();                                                                                       \
  cudaError_t launch_build_staged(const ShipEnvParams* params_dev, void* staged_dev, cudaStream_t st);         \
  cudaError_t launch_map_query(const SenvView& v, long long n, const double* north, const double* east,        \
                               double ship_length, int* contains, int* square, double* distance,               \
                               cudaStream_t st);                                                               \
  cudaError_t launch_map_safe_radius(const SenvView& v, long long n, const double* north, const double* east,  \
                                     float* out, cudaStream_t st);                                             \
  cudaError_t launch_math_selftest(long long n, unsigned long long seed, unsigned long long* mismatches_dev,   \
                                   cudaStream_t st);                                                           \
  }

SENV_DECLARE(senv_strict)
SENV_DECLARE(senv_fast)
