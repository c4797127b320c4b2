// Kernel instantiations for the FitzHugh-Nagumo model (streaming family).
#include "mlb_stream_nuts.cuh"
namespace mlb {
extern const ModelOps ops_fn_u = make_stream_ops<FnOde<0>>(MLB_MODEL_FN, 0);
extern const ModelOps ops_fn_n = make_stream_ops<FnOde<1>>(MLB_MODEL_FN, 1);
}
