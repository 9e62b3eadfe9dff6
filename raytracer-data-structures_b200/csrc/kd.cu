// kd.cu — K8 placeholder.
#include "rtds_internal.cuh"
int rtds_build_kd(rtds_ctx*, const rtds_build_params*, rtds_build_stats*)
{
    rtds_set_error("KD-tree builder not implemented yet");
    return RTDS_ERR_UNSUPPORTED;
}
