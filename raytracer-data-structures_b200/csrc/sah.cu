// sah.cu — K7 placeholder.
#include "rtds_internal.cuh"
int rtds_build_sah(rtds_ctx*, const rtds_build_params*, rtds_build_stats*)
{
    rtds_set_error("SAH-binned BVH builder not implemented yet");
    return RTDS_ERR_UNSUPPORTED;
}
