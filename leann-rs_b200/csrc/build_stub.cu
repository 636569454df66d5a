// temporary: the Vamana builder arrives in vamana_build.cu
#include "internal.h"
namespace leann {
void gpu_vamana_build(leann_cuda_index*, size_t, size_t, float, uint64_t) { throw Error(LEANN_ERR_INVALID_ARG, "vamana build: not implemented yet"); }
}
