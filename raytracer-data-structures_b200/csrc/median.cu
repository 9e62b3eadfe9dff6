// median.cu — K6 placeholder (replaced below in this round).
#include "rtds_internal.cuh"
int rtds_build_median(rtds_ctx*, int, rtds_build_stats*)
{
    rtds_set_error("median-split BVH builder not implemented yet");
    return RTDS_ERR_UNSUPPORTED;
}
