// scg_rollout_launch.cuh -- launch + variant dispatch of rollout_kernel, shared by the per-(system, precision) translation
// units scg_rollout_*.cu.  The rollout kernel is by far the most expensive template to compile; one translation unit per
// (system, precision) lets nvcc build them in parallel (build.py: --threads).  scg_api.cu only sees the entry points below.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "scg_kernels.cuh"

struct RolloutLaunch {          // what a launch needs from the handle
    int num_envs, obs_dim, nu, nx, L;
    DevTables tab;
    int extended;               // the env has an adversary input or non-default physics: only the extended collector implements those
    int fast_ok;                // the FAST schedule's preconditions hold (rollout_kernel: pre-drawn reset candidates, no observation noise, no
                                // done_on_violation / constraint penalty): known on the host, compiled into the default collector
};
struct RolloutLaunchResult {    // code: SCG_OK / SCG_EINVAL (msg) / SCG_ECUDA (cuda, msg = failing call)
    int code; cudaError_t cuda; const char* msg;
};

RolloutLaunchResult scg_rollout_cartpole_f64(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_cartpole_f32(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_quad1d_f64(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_quad1d_f32(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_quad2d_f64(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_quad2d_f32(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_quad3d_f64(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);
RolloutLaunchResult scg_rollout_quad3d_f32(const RolloutLaunch&, const scg_rollout_args*, RolloutParams, cudaStream_t);

#ifndef SCG_TC_DEFAULT
#define SCG_TC_DEFAULT 0      // 1: hidden-64 rollouts use the tensor-core collector unless the caller asks for the CUDA-core one
#endif
#ifdef SCG_ROLLOUT_IMPL   // defined by the scg_rollout_*.cu units only
#include "scg_rollout_tc.cuh"
#define RL_CUDA(call) do { cudaError_t st_ = (call); if (st_ != cudaSuccess) return RolloutLaunchResult{SCG_ECUDA, st_, #call}; } while (0)

template <int SYS, typename R, int HID, int E, int NT, bool SAFE = false, bool XT = false, int FM = SCG_FM_RUNTIME>
static RolloutLaunchResult launch_rollout(const RolloutLaunch& e, RolloutParams P, const scg_rollout_args* a, cudaStream_t s) {
    const RolloutSmem<HID, E> S(e.obs_dim, e.nu, e.L * e.nx * (int)(sizeof(R) / 4), NT / 64, (XT && P.adv_policy) ? P.adv_dim : 0, SysDims<SYS>::NADV);
    const size_t smem = (size_t)S.total_floats * sizeof(float);
    auto kern = rollout_kernel<SYS, R, HID, E, NT, SAFE, XT, FM>;
    RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the kernel lives out of shared memory (weights, tiles, tables): ask for the largest carve-out so several CTAs fit per SM
    RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    const int G = (e.num_envs + E - 1) / E;
#ifdef SCG_PROF
    static long long* d_prof = nullptr;
    if (!d_prof) cudaMalloc((void**)&d_prof, (16 + 2048) * sizeof(long long));
    cudaMemset(d_prof, 0, (16 + 2048) * sizeof(long long));
    P.prof = d_prof;
    { long long z[64] = {0}; cudaMemcpyToSymbol(scg_ticks, z, sizeof(z)); }
#endif
    kern<<<G, NT, smem, s>>>(e.tab, P);
    RL_CUDA(cudaGetLastError());
#ifdef SCG_PROF
    {
        long long hp[16]; cudaMemcpy(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost);
        const char* nm[8] = {"eps+obs_write", "mlp", "env step (owner)", "-", "-", "end barrier", "-", "tail(mlp+gae)"};
        fprintf(stderr, "[SCG_PROF] CTA0 cycles per step (T=%d):", a->T);
        for (int i = 0; i < 8; ++i) if (nm[i][0] != '-') fprintf(stderr, " %s=%.0f", nm[i], i == 7 ? (double)hp[i] : (double)hp[i] / a->T);
        fprintf(stderr, "\n");
        fprintf(stderr, "[SCG_PROF] env-phase cycles per step by warp (7 = owner, 6 = role 1 constraints + policy record, 5 = role 2 reward / mse, 4 = role 3 reset candidates, 0-3 = obs record + normals):");
        for (int i = 0; i < 8; ++i) fprintf(stderr, " w%d=%.0f", i, (double)hp[8 + i] / a->T);
        fprintf(stderr, "\n");
        long long tk[64]; cudaMemcpyFromSymbol(tk, scg_ticks, sizeof(tk));
        fprintf(stderr, "[SCG_PROF] ticks/step:");
        for (int i = 0; i < 25; ++i) fprintf(stderr, " %d:%.0f", i, (double)tk[i] / a->T);
        fprintf(stderr, "\n");
        std::vector<long long> ct(2048); cudaMemcpy(ct.data(), d_prof + 16, 2048 * sizeof(long long), cudaMemcpyDeviceToHost);
        int GG = G > 1024 ? 1024 : G;
        long long mn = ct[0], mx = ct[0]; double sum = 0;
        for (int i = 0; i < GG; ++i) { if (ct[i] < mn) mn = ct[i]; if (ct[i] > mx) mx = ct[i]; sum += ct[i]; }
        fprintf(stderr, "[SCG_PROF] per-CTA horizon cycles: min=%lld mean=%.0f max=%lld\n", mn, sum / GG, mx);
    }
#else
    (void)a;
#endif
    return RolloutLaunchResult{SCG_OK, cudaSuccess, nullptr};
}

// tensor-core collector (tcgen05 / TMEM, 3xTF32): 32 envs x 256 threads, hidden 64
template <int SYS, typename R, int HID>
static RolloutLaunchResult launch_rollout_tc(const RolloutLaunch& e, RolloutParams P, cudaStream_t s) {
    const typename tc::Variant<HID>::Plan S(e.obs_dim, e.nu, e.L * e.nx * (int)(sizeof(R) / 4));
    const size_t smem = (size_t)S.total_floats * sizeof(float);
    auto kern = rollout_tc_kernel<SYS, R, HID>;
    cudaFuncAttributes fa;
    RL_CUDA(cudaFuncGetAttributes(&fa, kern));
    if (smem + fa.sharedSizeBytes > 227 * 1024) return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: the tensor-core collector does not fit this configuration in shared memory"};
    RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int G = (e.num_envs + tc::E - 1) / tc::E;
#ifdef SCG_TC_PROF
    { long long z[16] = {0}; cudaMemcpyToSymbol(tc_ticks, z, sizeof(z)); }
#endif
    kern<<<G, tc::NT, smem, s>>>(e.tab, P);
    RL_CUDA(cudaGetLastError());
#ifdef SCG_TC_PROF
    {
        long long tk[16]; cudaMemcpyFromSymbol(tk, tc_ticks, sizeof(tk));
        const char* nm[11] = {"L1 issue", "L1 wait", "E1", "E1 fence+sync", "L2 issue", "L2 wait", "E2", "E2 sync", "owner env", "owner->barrier", "env phase (warp 0)"};
        fprintf(stderr, "[SCG_TC_PROF] cycles / forward (T=%d, ~T+1 forwards):", P.T);
        for (int i = 0; i < 11; ++i) fprintf(stderr, " %s=%.0f", nm[i], (double)tk[i] / P.T);
        fprintf(stderr, "\n");
    }
#endif
    return RolloutLaunchResult{SCG_OK, cudaSuccess, nullptr};
}

// envs_per_block -> (E, threads): a CTA runs 8 threads per env (4x4 register tiles at hidden 64).  Default: one 8-warp CTA of
// 32 envs (owner warp, helper warps, idle GEMM warps: see rollout_kernel); 16 envs x 128 threads at very large batches.
template <int SYS, typename R>
static RolloutLaunchResult dispatch_rollout(const RolloutLaunch& e, const scg_rollout_args* a, const RolloutParams& P, cudaStream_t s) {
    int E = a->envs_per_block & 0xffff;
    const int NTq = (a->envs_per_block >> 16) & 0xffff;   // requested threads per CTA (0 = 8 per env)
    (void)NTq;
#ifndef SCG_PROF
    if (e.extended || a->d_adv_policy) {  // adversary (RARL / RAP), adversary input, pyb_gnd / pyb_drag / pyb_dw: the extended collector
        if (a->safety_constraints > 0) return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: the safety-layer collector does not take an adversary / non-default physics"};
        if (a->hidden_dim != 64 || !(E == 0 || E == 32)) return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: the extended collector (adversary / non-default physics) is built for hidden_dim 64, 32 envs per CTA"};
        return launch_rollout<SYS, R, 64, 32, 256, false, true>(e, P, a, s);
    }
    if (a->safety_constraints > 0) {      // SafeExplorer collector: CUDA-core MLP, 32 envs per CTA, hidden 64 (the reference's default width)
        if (a->hidden_dim != 64 || !(E == 0 || E == 32)) return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: the safety-layer collector is built for hidden_dim 64, 32 envs per CTA"};
        return launch_rollout<SYS, R, 64, 32, 256, true>(e, P, a, s);
    }
    // mlp_engine: 0 = library default, 1 = CUDA cores (FFMA2), 2 = tensor cores (tcgen05, 3xTF32; hidden 64, 32 envs per CTA)
    // library default (SCG_MLP_AUTO): tensor cores for hidden 128 (where they win by a wide margin), CUDA cores for hidden 32 / 64
    // (measured: DESIGN.md 4); a configuration that does not fit the tensor-core kernel's shared-memory plan falls back to FFMA2
    const bool force_tc = a->mlp_engine == SCG_MLP_TCGEN05 || a->mlp_engine == SCG_MLP_TCGEN05_TF32;
    const bool want_tc = force_tc ||
                         (a->mlp_engine == SCG_MLP_AUTO && (E == 0 || E == 32) && (a->hidden_dim == 128 || (SCG_TC_DEFAULT && a->hidden_dim == 64)));
    if (want_tc) {
        if (a->hidden_dim != 64 && a->hidden_dim != 128) return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: the tensor-core collector is available for hidden_dim 64 and 128"};
        if (e.obs_dim > 32 && force_tc) return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: tensor-core collector: observation too wide"};
        if (e.obs_dim <= 32) {
            // hidden 64: fp16 fused-net form (half the MMAs of 3xTF32); SCG_MLP_TCGEN05_TF32 selects the TF32 form
            const bool tf32_64 = a->mlp_engine == SCG_MLP_TCGEN05_TF32;
            RolloutLaunchResult r = a->hidden_dim == 128 ? launch_rollout_tc<SYS, R, 128>(e, P, s)
                                    : tf32_64 ? launch_rollout_tc<SYS, R, 64>(e, P, s) : launch_rollout_tc<SYS, R, tc::HID64F>(e, P, s);
            if (r.code == SCG_OK || force_tc) return r;
        }
    }
#endif
#ifdef SCG_PROF
    if (a->hidden_dim == 64 && (E == 0 || E == 32)) return e.fast_ok ? launch_rollout<SYS, R, 64, 32, 256, false, false, SCG_FM_FAST>(e, P, a, s)
                                                                     : launch_rollout<SYS, R, 64, 32, 256>(e, P, a, s);
    if (a->hidden_dim == 64 && E == 16) return launch_rollout<SYS, R, 64, 16, 128>(e, P, a, s);
    return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "SCG_PROF build: only hidden 64 / E 32 or 16"};
#else
    switch (a->hidden_dim) {
#ifndef SCG_DEV_HID64_ONLY
        case 32:
            if (E == 0) E = 32;
            if (E == 32) return launch_rollout<SYS, R, 32, 32, 256>(e, P, a, s);
            if (E == 16) return launch_rollout<SYS, R, 32, 16, 128>(e, P, a, s);
            break;
#endif
        case 64:
            if (E == 0) E = 32;
#ifdef SCG_DEV_HID64_ONLY   // 16 threads per env, two CTAs per SM: measured slower than the 8-thread-per-env variants (round 1)
            if (E == 16 && NTq == 256) return launch_rollout<SYS, R, 64, 16, 256>(e, P, a, s);
            if (E == 8 && NTq == 128) return launch_rollout<SYS, R, 64, 8, 128>(e, P, a, s);
#endif
            // the default collector: the schedule is chosen here (the kernel's own test, evaluated on the host) and compiled in
            if (E == 32) return e.fast_ok ? launch_rollout<SYS, R, 64, 32, 256, false, false, SCG_FM_FAST>(e, P, a, s)
                                          : launch_rollout<SYS, R, 64, 32, 256, false, false, SCG_FM_SERIAL>(e, P, a, s);
            if (E == 16) return launch_rollout<SYS, R, 64, 16, 128>(e, P, a, s);
            if (E == 8) return launch_rollout<SYS, R, 64, 8, 64>(e, P, a, s);
            break;
#ifndef SCG_DEV_HID64_ONLY
        case 128:
            if (E == 0) E = 16;
            if (E == 16) return launch_rollout<SYS, R, 128, 16, 256>(e, P, a, s);
            if (E == 8) return launch_rollout<SYS, R, 128, 8, 128>(e, P, a, s);
            break;
#endif
        default: return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: hidden_dim must be 32, 64 or 128"};
    }
    return RolloutLaunchResult{SCG_EINVAL, cudaSuccess, "scg_rollout: unsupported envs_per_block for this hidden_dim"};
#endif
}
#endif  // SCG_ROLLOUT_IMPL
