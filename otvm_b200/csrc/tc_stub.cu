// Placeholder until the tcgen05 memory-read kernel lands.
#include "common.cuh"
namespace otvm {
int memory_read_tc(const otvm_read_params*, cudaStream_t) { return OTVM_ERR_UNSUPPORTED; }
bool memory_read_tc_supported(const otvm_read_params*) { return false; }
int64_t memory_read_tc_workspace(int, int, int, int) { return 0; }
}
