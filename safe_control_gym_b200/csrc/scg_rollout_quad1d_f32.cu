// fused collector, SCG_SYS_QUAD1D / float: one translation unit per (system, precision) so that nvcc builds them in parallel
#define SCG_ROLLOUT_IMPL
#include "scg_rollout_launch.cuh"
RolloutLaunchResult scg_rollout_quad1d_f32(const RolloutLaunch& e, const scg_rollout_args* a, RolloutParams P, cudaStream_t s) {
    return dispatch_rollout<SCG_SYS_QUAD1D, float>(e, a, P, s);
}
