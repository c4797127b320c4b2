// Kernel instantiations for one model (compiled as its own translation unit).
#include "mlb_launch.cuh"
namespace mlb {
extern const ModelOps ops_hier8n = make_ops<HierModel<8, 1>>(MLB_MODEL_HIER, 8, 1);
}
