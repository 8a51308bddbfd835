// Placeholder until the tcgen05 kernels land: nothing is routed to tensor cores.
#include "common.cuh"
namespace otvm {
int conv2d_tc(const otvm_conv_params*, cudaStream_t) { return OTVM_ERR_UNSUPPORTED; }
bool conv2d_tc_supported(const otvm_conv_params*) { return false; }
int memory_read_tc(const otvm_read_params*, cudaStream_t) { return OTVM_ERR_UNSUPPORTED; }
bool memory_read_tc_supported(const otvm_read_params*) { return false; }
int64_t memory_read_tc_workspace(int, int, int, int) { return 0; }
}
