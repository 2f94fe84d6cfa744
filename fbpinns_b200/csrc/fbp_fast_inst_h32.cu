// Instantiations of the tiled kernels for hidden width 32 (one translation unit per width to parallelise the build).
#include "fbp_fast.cuh"

int fbp_fast_launch_h32(int nhid, int na2, int na1, bool backward, const FastArgs& a, int grid, cudaStream_t st) {
    if (nhid == 1) { FBP_FAST_JET_SWITCH(32, 1) }
    else if (nhid == 2) { FBP_FAST_JET_SWITCH(32, 2) }
    fbp_set_error("fbp_fast: no tiled instance for H=32 nhid=%d jets=(%d,%d)", nhid, na2, na1);
    return 3;
}
