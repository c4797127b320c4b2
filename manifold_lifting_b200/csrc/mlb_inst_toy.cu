// Kernel instantiations for one model (compiled as its own translation unit).
#include "mlb_launch.cuh"
namespace mlb {
extern const ModelOps ops_toy = make_ops<ToyModel>(MLB_MODEL_TOY, 1, 0);
}
