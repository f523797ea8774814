// placeholder until the tcgen05 variant lands (replaced below in this round)
#include "common.cuh"
#include "render_common.cuh"
#include "../../include/ssdnerf_b200.h"
namespace ssdnerf {
size_t dec_s_blob_floats() { return 0; }
int render_s_launch(const RenderParams&, int, uint32_t*, int, cudaStream_t) {
    return set_error_msg(SSDNERF_ERR_ARG, "render_fwd: variant S not built");
}
}
