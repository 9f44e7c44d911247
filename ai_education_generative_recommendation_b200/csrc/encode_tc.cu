// encode_tc.cu — tensor-core encoder (tcgen05) + margin gate.  Placeholder until the kernel lands:
// the FAST mode refuses to run rather than silently taking another path.
#include "common.cuh"

namespace rqb {
int get_indices_fast(rqb200_model *m, const float *x, int64_t n, int64_t *codes, float *z_out,
                     int64_t *stats_host, cudaStream_t s) {
    set_error("RQB200_ENCODE_FAST is not available in this build");
    return RQB200_EINVAL;
}
}  // namespace rqb
