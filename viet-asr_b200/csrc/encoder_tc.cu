// tcgen05 fused sub-block (placeholder until the tensor path lands)
#include "common.cuh"
#include "kernels.cuh"
namespace vasr {
int tc_init() { return VASR_OK; }
bool subblock_tc_supported(const SubBlock&) { return false; }
int launch_subblock_tc(const SubBlock&, const float*, const float*, float*, int, int, int, const int*, const int*, int, cudaStream_t)
{ return set_error(VASR_EINVAL, "tcgen05 path not built"); }
}
