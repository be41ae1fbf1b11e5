// Tensor-core (tcgen05, bf16 operands, fp32 accumulate in TMEM) path.  Placeholder entry
// points until the kernels land: they fail loudly, there is no fallback.
#include "mlp_fp32.cuh"

namespace rl8 {

int64_t collect_tc_workspace(const rl8_model*, int64_t, int32_t) { return RL8_ERR_UNSUPPORTED; }
int collect_tc(const rl8_model*, const rl8_rollout*, void*, int64_t, cudaStream_t) {
  return RL8_ERR_UNSUPPORTED;
}
int64_t ppo_tc_workspace(const rl8_model*, int64_t) { return RL8_ERR_UNSUPPORTED; }
int ppo_minibatch_tc(const rl8_model*, const rl8_model*, const rl8_batch*, const int64_t*, int64_t,
                     int64_t, double, const rl8_ppo_hparams*, double*, void*, int64_t,
                     cudaStream_t) {
  return RL8_ERR_UNSUPPORTED;
}
int mlp_forward_tc(const rl8_model*, int, const RowMap&, int64_t, float*, int, void*, int64_t,
                   cudaStream_t) {
  return RL8_ERR_UNSUPPORTED;
}

}  // namespace rl8
