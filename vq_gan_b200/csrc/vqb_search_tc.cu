// placeholder until the tcgen05 kernel lands (next commit): routes to the fp32 tile kernel
#include "vqb_common.cuh"
namespace vqb {
size_t search_tc_workspace_bytes(int64_t, int, int) { return 0; }
int launch_search_tc(const float* z, int64_t B, int D, int64_t HW, const float* E, int K, const void* pack,
                     int64_t* idx_out, float* dmin_out, void*, size_t, int64_t*, cudaStream_t s) {
    return launch_search_fp32(z, B, D, HW, E, K, pack, nullptr, nullptr, 0, idx_out, dmin_out, s);
}
}  // namespace vqb
