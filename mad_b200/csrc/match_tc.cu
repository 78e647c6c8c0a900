// placeholder, replaced below
#include "common.cuh"
#include "match_common.cuh"
int mad_match_tc(const void*, int, int, const void*, int, int, const int32_t*, const int32_t*, double, int, int32_t*,
                 const int64_t*, int32_t*, int32_t*, double*, int, int, int32_t*, double*, cudaStream_t) {
    mad_set_error("tcgen05 matching kernel not built");
    return MAD_ERR_NODEVICE;
}
