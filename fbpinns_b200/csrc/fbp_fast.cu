// placeholder: tiled kernels are added in fbp_fast_*.cu
#include "fbp_common.cuh"
int fbp_fast_lookup(const fbp_plan_desc*, FastSpec*) { return -1; }
int fbp_fast_forward(const fbp_plan*, const fbp_takes_view*, const float*, const float*, const float*, float*, cudaStream_t) { fbp_set_error("no tiled kernel"); return 3; }
int64_t fbp_fast_backward_workspace(const fbp_plan*, const fbp_takes_view*) { return 0; }
int fbp_fast_backward(const fbp_plan*, const fbp_takes_view*, const float*, const float*, const float*, const float*, float*, int, float*, cudaStream_t) { fbp_set_error("no tiled kernel"); return 3; }
