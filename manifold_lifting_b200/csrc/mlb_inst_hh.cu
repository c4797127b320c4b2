// Kernel instantiations for the Hodgkin-Huxley model (streaming family).
#include "mlb_stream_nuts.cuh"
namespace mlb {
extern const ModelOps ops_hh_u = make_stream_ops<HhOde<0>>(MLB_MODEL_HH, 0);
extern const ModelOps ops_hh_n = make_stream_ops<HhOde<1>>(MLB_MODEL_HH, 1);
}
