// bf16 tcgen05 rollout kernel — placeholder until the tensor-core path lands.
#include "rollout_params.cuh"

namespace simba {
bool rollout_tc_supported(int O, int A, int L, int U, int H) { return false; }
cudaError_t launch_rollout_tc(const RolloutParams& prm, int n_tiles, cudaStream_t stream) {
  return cudaErrorNotSupported;
}
}  // namespace simba
