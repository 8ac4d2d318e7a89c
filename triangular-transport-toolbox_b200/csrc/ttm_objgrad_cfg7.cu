// instantiation 7 of the K-objgrad kernel template (see ttm_objgrad_impl.cuh)
#include "ttm_objgrad_impl.cuh"

cudaError_t ttm_objgrad_cfg7(const ObjArgs& a, bool grad, int grid, size_t smem, cudaStream_t st) {
    return ttm_obj::launch_cfg<3, false, true, 0, true, true, 4, 1>(a, grad, grid, smem, st);
}
