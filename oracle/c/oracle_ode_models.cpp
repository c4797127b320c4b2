// ODE-constrained models for the CPU oracle (FitzHugh-Nagumo, Hodgkin-Huxley).
// TEST INFRASTRUCTURE ONLY.  Filled in after the non-ODE models.
#include "oracle_models.hpp"

namespace orc {
Model* make_fn_model(int, int, const double*, int) { return nullptr; }
Model* make_hh_model(int, int, const double*, int) { return nullptr; }
}  // namespace orc
