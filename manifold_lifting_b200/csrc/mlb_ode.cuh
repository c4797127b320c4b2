// Constants of an ODE-constrained model as the kernels see them.
#pragma once
#include <utility>

#include "mlb_common.cuh"

namespace mlb {

constexpr int kOdeConsts = 40;

// Passed BY VALUE as a __grid_constant__ kernel parameter: the scalar constants are then
// constant-bank operands; the arrays live in one device blob per model handle (uploaded
// on first use, mlb_api.cu).
struct OdeData {
  const double* y_obs;   // [Y]
  const double* dt_seq;  // [n_steps] integrator step sizes
  const int* obs_step;   // [Y] index of the step whose end point is observation i
  const double* aux;     // [n_steps] per-step forcing (HH stimulus current) or nullptr
  int Y, n_steps;
  double consts[kOdeConsts];  // model constants (HH: rate constants and reciprocals)
};

}  // namespace mlb
