// lsm_tiled.cu — performance kernels (placeholder until the tiled kernels land).
#include "lsm_dev.cuh"
#include "lsm_kernels.h"

namespace lsm {

template <class T> bool stage_tiled_supported(int, const StageParams<T>&) { return false; }
template <class T> cudaError_t launch_stage_tiled(int, const StageParams<T>&, int, cudaStream_t) { return cudaErrorNotSupported; }

template bool stage_tiled_supported<float>(int, const StageParams<float>&);
template bool stage_tiled_supported<double>(int, const StageParams<double>&);
template cudaError_t launch_stage_tiled<float>(int, const StageParams<float>&, int, cudaStream_t);
template cudaError_t launch_stage_tiled<double>(int, const StageParams<double>&, int, cudaStream_t);

}  // namespace lsm
