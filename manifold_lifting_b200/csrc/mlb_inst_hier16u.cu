// Kernel instantiations for one model (compiled as its own translation unit).
#include "mlb_launch.cuh"
namespace mlb {
extern const ModelOps ops_hier16u = make_ops<HierModel<16, 0>>(MLB_MODEL_HIER, 16, 0);
}
