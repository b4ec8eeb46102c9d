// scg_ppo.cu -- fused PPO update for sm_100a.
//
// Reference: PPOAgent.update, safe_control_gym/controllers/ppo/ppo_utils.py:82-146 -- for every epoch and every minibatch:
//   policy_loss / entropy_loss / approx_kl (compute_policy_loss :56-69), KL-gated actor Adam step (:128-131),
//   value_loss (compute_value_loss :71-80), critic Adam step (:133-137).
// The reference pays ~40 framework kernel launches + one host sync per minibatch.  Here the whole
// opt_epochs x num_mini_batch sequence is ONE cooperative kernel: a minibatch is cut into chunks of MS samples, chunk c goes
// to CTA c % G; a CTA keeps both networks' parameters and the chunk's activations in shared memory, runs forward, loss and
// backward as register-tiled FP32 GEMMs and accumulates its weight gradients in registers.  The CTAs then publish their partial
// gradients, cross a grid barrier, and each CTA reduces + applies Adam to its own slice of the parameter vector (fixed summation
// order => bit-reproducible for a given grid), crosses a second barrier and reloads the parameters.
//
// Shared-memory tiles are [feature][sample] with a row pitch of MS + 4 floats; weights keep the nn.Linear [out][in] layout with
// a pitch of in + 4, which serves the forward GEMM (vector loads along `in`), the input-gradient GEMM (vector loads along `in`)
// and needs no transposed copy.  Unit ownership inside a thread tile is interleaved so that the lanes of a warp hit distinct
// 16-byte bank groups.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "scg.h"
#include "scg_kernels.cuh"

namespace cg = cooperative_groups;

// -DSCG_PPO_PROF: per-phase cycle accounting of CTA 0 / thread 0 (printed by the kernel at its end)
#ifdef SCG_PPO_PROF
#define PP_T(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long c_ = clock64(); pp_acc[i] += c_ - pp_last; pp_last = c_; } } while (0)
#else
#define PP_T(i) do { } while (0)
#endif

void scg_internal_set_error(const std::string& msg);   // scg_api.cu (message returned by scg_last_error(NULL))

namespace {

constexpr int PPO_NT = 256;     // threads per CTA: 128 per network
constexpr int PPO_NETT = 128;
constexpr int PPO_MAXA = 4;     // act_dim <= 4 on this path
constexpr int PPO_NSTAT = 8;
constexpr int PPO_CLUSTER_NPT = 2;   // float4 parameter groups per thread a CTA of the cluster kernel owns (registers: weights, two Adam moments)

struct TrainLayout {
    int a_w1, a_b1, a_w2, a_b2, a_w3, a_b3, a_ls;
    int c_w1, c_b1, c_w2, c_b2, c_w3, c_b3;
    int critic0, total;
};
__host__ __device__ inline TrainLayout train_layout(int obs, int act, int hid) {
    TrainLayout L;
    int o = 0;
    auto take = [&](int n) { int s = o; o += (n + 3) & ~3; return s; };
    L.a_w1 = take(hid * obs); L.a_b1 = take(hid); L.a_w2 = take(hid * hid); L.a_b2 = take(hid);
    L.a_w3 = take(act * hid); L.a_b3 = take(act); L.a_ls = take(act);
    L.critic0 = o;
    L.c_w1 = take(hid * obs); L.c_b1 = take(hid); L.c_w2 = take(hid * hid); L.c_b2 = take(hid);
    L.c_w3 = take(hid); L.c_b3 = take(1);
    L.total = o;
    return L;
}

struct PpoParams {
    int obs, act, hid, activation;
    int mb, n_mb, epochs, clipped_value;
    long long block_rows;
    long long block_stride[6];
    const float* key[6];
    const int* perm;
    long long perm_stride;
    double clip, target_kl, ent_coef, lr_a, lr_c, beta1, beta2, eps;
    float *params, *m, *v;
    long long* steps;
    float* stats;
    float* ws;          // [G][ws_pitch]
    int ws_pitch;
    float* grad_out;
};

// shared-memory plan (floats)
template <int HID, int MS>
struct PpoSmem {
    static constexpr int MSP = MS + 4;
    static constexpr int W2S = HID + 4;
    int obsp, w1s;
    int net_floats;                   // parameters of one net
    int o_w1, o_b1, o_w2, o_b2, o_w3, o_b3, o_ls;   // inside a net block
    int o_x, o_h1, o_h2, o_dout, o_act, o_sc, o_red, o_row, total;
    __host__ __device__ PpoSmem(int obs) {
        obsp = (obs + 3) & ~3;
        w1s = (obsp % 32 == 0) ? obsp + 4 : obsp;
        int o = 0;
        o_w1 = o; o += HID * w1s;
        o_b1 = o; o += HID;
        o_w2 = o; o += HID * W2S;
        o_b2 = o; o += HID;
        o_w3 = o; o += PPO_MAXA * HID;
        o_b3 = o; o += 4;
        o_ls = o; o += 4;
        net_floats = o;
        o = 2 * net_floats;
        o_x = o; o += obsp * MSP;
        o_h1 = o; o += 2 * HID * MSP;          // [net][HID][MSP]
        o_h2 = o; o += 2 * HID * MSP;
        o_dout = o; o += 2 * PPO_MAXA * MSP;   // [net][4][MSP]: layer-3 outputs, then d(loss)/d(outputs)
        o_act = o; o += PPO_MAXA * MSP;        // actions [d][s]
        o_sc = o; o += 4 * MSP;                // logp_old, adv, ret, v_old
        o_red = o; o += 64;                    // block-reduction scratch
        o_row = o; o += 2 * MS;                // gathered rows: (block, index in block), ints; block < 0 = padding sample
        total = o;
    }
};

__device__ __forceinline__ float act_deriv(float h, int act) {
    // derivative of the hidden activation expressed through its OUTPUT h
    if (act == SCG_ACT_TANH) return fmaf(-h, h, 1.0f);
    if (act == SCG_ACT_RELU) return h > 0.0f ? 1.0f : 0.0f;
    return h > 0.0f ? 1.0f : 0.01f;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// out[u][s] = act(b[u] + sum_k W[u][k] in[k][s]) for this thread's UT interleaved units x 4 samples
template <int HID, int MS, int UT, int TUN>
__device__ __forceinline__ void dense_forward(const float* __restrict__ W, int wpitch, const float* __restrict__ b, const float* __restrict__ in,
                                              int K /*multiple of 4*/, float* __restrict__ out, int tu, int ts, int act) {
    constexpr int MSP = MS + 4;
    float acc[UT][4];
#pragma unroll
    for (int j = 0; j < UT; ++j) {
        const float bb = b[tu + TUN * j];
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = bb;
    }
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const float4 x0 = ld4(in + (k + 0) * MSP + 4 * ts), x1 = ld4(in + (k + 1) * MSP + 4 * ts);
        const float4 x2 = ld4(in + (k + 2) * MSP + 4 * ts), x3 = ld4(in + (k + 3) * MSP + 4 * ts);
#pragma unroll
        for (int j = 0; j < UT; ++j) {
            const float4 w = ld4(W + (tu + TUN * j) * wpitch + k);
            acc[j][0] = fmaf(w.x, x0.x, acc[j][0]); acc[j][1] = fmaf(w.x, x0.y, acc[j][1]); acc[j][2] = fmaf(w.x, x0.z, acc[j][2]); acc[j][3] = fmaf(w.x, x0.w, acc[j][3]);
            acc[j][0] = fmaf(w.y, x1.x, acc[j][0]); acc[j][1] = fmaf(w.y, x1.y, acc[j][1]); acc[j][2] = fmaf(w.y, x1.z, acc[j][2]); acc[j][3] = fmaf(w.y, x1.w, acc[j][3]);
            acc[j][0] = fmaf(w.z, x2.x, acc[j][0]); acc[j][1] = fmaf(w.z, x2.y, acc[j][1]); acc[j][2] = fmaf(w.z, x2.z, acc[j][2]); acc[j][3] = fmaf(w.z, x2.w, acc[j][3]);
            acc[j][0] = fmaf(w.w, x3.x, acc[j][0]); acc[j][1] = fmaf(w.w, x3.y, acc[j][1]); acc[j][2] = fmaf(w.w, x3.z, acc[j][2]); acc[j][3] = fmaf(w.w, x3.w, acc[j][3]);
        }
    }
#pragma unroll
    for (int j = 0; j < UT; ++j) {
        if (act == SCG_ACT_TANH) {
            acc[j][0] = tanhf(acc[j][0]); acc[j][1] = tanhf(acc[j][1]); acc[j][2] = tanhf(acc[j][2]); acc[j][3] = tanhf(acc[j][3]);
        } else if (act == SCG_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaxf(acc[j][i], 0.0f);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = acc[j][i] > 0.0f ? acc[j][i] : 0.01f * acc[j][i];
        }
        st4(out + (tu + TUN * j) * MSP + 4 * ts, make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]));
    }
}

// train-layout index -> offset in the CTA's padded shared-memory parameter copy (-1: layout padding, nothing to store)
template <int HID, int MS>
__device__ __forceinline__ int train_to_smem(int p, const TrainLayout& L, const PpoSmem<HID, MS>& S, int obs, int A) {
    constexpr int W2S = HID + 4;
    const int net = p >= L.critic0;
    const int base = net * S.net_floats;
    const int w1 = net ? L.c_w1 : L.a_w1, b1 = net ? L.c_b1 : L.a_b1, w2 = net ? L.c_w2 : L.a_w2, b2 = net ? L.c_b2 : L.a_b2;
    const int w3 = net ? L.c_w3 : L.a_w3, b3 = net ? L.c_b3 : L.a_b3, nout = net ? 1 : A;
    if (p >= w1 && p < w1 + HID * obs) { const int i = p - w1; return base + S.o_w1 + (i / obs) * S.w1s + (i % obs); }
    if (p >= b1 && p < b1 + HID) return base + S.o_b1 + (p - b1);
    if (p >= w2 && p < w2 + HID * HID) { const int i = p - w2; return base + S.o_w2 + (i / HID) * W2S + (i % HID); }
    if (p >= b2 && p < b2 + HID) return base + S.o_b2 + (p - b2);
    if (p >= w3 && p < w3 + nout * HID) return base + S.o_w3 + (p - w3);
    if (p >= b3 && p < b3 + nout) return base + S.o_b3 + (p - b3);
    if (!net && p >= L.a_ls && p < L.a_ls + A) return S.o_ls + (p - L.a_ls);
    return -1;
}

// CLUSTER = false: one cooperative grid, partial gradients through global memory, two grid barriers per minibatch (large minibatches).
// CLUSTER = true : ONE thread-block cluster works on the minibatch (ppo.yaml's default minibatch of 64 = 8 CTAs x 8 samples): partial
//                  gradients stay in shared memory, every CTA reduces its slice of the parameter vector over the cluster through
//                  distributed shared memory, keeps that slice's Adam moments and master weights in REGISTERS for the whole launch and
//                  stores the updated weights straight into every CTA's shared-memory copy -- no global round trip, no parameter
//                  reload, two cluster barriers per minibatch instead of two grid barriers + ~5 L2 round trips.
template <int HID, int MS, bool CLUSTER>
__global__ void __launch_bounds__(PPO_NT, 1) ppo_update_kernel(PpoParams P) {
    constexpr int MSP = MS + 4;
    constexpr int W2S = HID + 4;
    constexpr int SG = MS / 4;              // sample groups of 4
    constexpr int TUN = PPO_NETT / SG;      // unit groups of the forward tile
    constexpr int UT = HID / TUN;           // units per thread (interleaved, stride TUN)
    constexpr int TKN = HID / 4;            // k groups of the weight-gradient tile (4 interleaved k per thread, stride TKN)
    constexpr int TUN2 = PPO_NETT / TKN;    // unit groups of the weight-gradient tile
    constexpr int UT2 = HID / TUN2;
    constexpr int NW1 = (HID * 32 + PPO_NETT - 1) / PPO_NETT;      // dW1 accumulators per thread (obs <= 32)
    constexpr int NW3 = (PPO_MAXA * HID + PPO_NETT - 1) / PPO_NETT;
    static_assert(UT >= 1 && UT2 >= 1 && HID % TUN == 0 && HID % TUN2 == 0, "tile");
    static_assert(HID <= PPO_NETT, "bias gradients use one thread per unit");

    extern __shared__ __align__(16) float smem[];
    const PpoSmem<HID, MS> S(P.obs);
    const TrainLayout L = train_layout(P.obs, P.act, HID);
    cg::grid_group grid = cg::this_grid();
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, net = tid / PPO_NETT, t = tid % PPO_NETT;
    const int G = gridDim.x, cta = blockIdx.x;            // CLUSTER: the grid is exactly one cluster
    float* gbuf = smem + S.total;                          // CLUSTER: this CTA's partial gradient (train layout) + loss sums, read by the peers
    const int obs = P.obs, A = P.act, act = P.activation, mb = P.mb;
    const int nout = net == 0 ? A : 1;
    const int n_chunks = (mb + MS - 1) / MS;
    const unsigned obs_magic = (unsigned)((0x100000000ull + (unsigned)obs - 1) / (unsigned)obs);     // ceil(2^32 / obs)
    float* sp = smem + net * S.net_floats;           // this net's parameters
    float* sX = smem + S.o_x;
    float* sH1 = smem + S.o_h1 + net * HID * MSP;
    float* sH2 = smem + S.o_h2 + net * HID * MSP;
    float* sD = smem + S.o_dout + net * PPO_MAXA * MSP;
    float* sAct = smem + S.o_act;
    float* sSc = smem + S.o_sc;
    float* sRed = smem + S.o_red;
    int* sRow = reinterpret_cast<int*>(smem + S.o_row);   // block index (or -1)
    int* sIdx = sRow + MS;                                  // row inside the block
    const int g_w1 = net ? L.c_w1 : L.a_w1, g_b1 = net ? L.c_b1 : L.a_b1, g_w2 = net ? L.c_w2 : L.a_w2, g_b2 = net ? L.c_b2 : L.a_b2;
    const int g_w3 = net ? L.c_w3 : L.a_w3, g_b3 = net ? L.c_b3 : L.a_b3;

    auto load_params = [&]() {
        // global train layout -> padded shared layout (both nets, all 256 threads).  Loads first, stores after: the L2 round trips overlap.
        {
            constexpr int Q = HID * HID / 4;              // float4 per W2 matrix
            constexpr int PER = 2 * Q / PPO_NT;
            static_assert((2 * Q) % PPO_NT == 0, "W2 staging");
            float4 r[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = tid + PPO_NT * j, n = i / Q, q = i % Q;
                r[j] = __ldcg(reinterpret_cast<const float4*>(P.params + (n ? L.c_w2 : L.a_w2)) + q);
            }
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = tid + PPO_NT * j, n = i / Q, q = i % Q, u = q / (HID / 4), k4 = q % (HID / 4);
                st4(smem + n * S.net_floats + S.o_w2 + u * W2S + 4 * k4, r[j]);
            }
        }
        for (int n = 0; n < 2; ++n) {
            float* d = smem + n * S.net_floats;
            const float* g = P.params;
            const int w1 = n ? L.c_w1 : L.a_w1, b1 = n ? L.c_b1 : L.a_b1, b2 = n ? L.c_b2 : L.a_b2;
            const int w3 = n ? L.c_w3 : L.a_w3, b3 = n ? L.c_b3 : L.a_b3, no = n ? 1 : A;
#pragma unroll 4
            for (int i = tid; i < HID * obs; i += PPO_NT) {
                const int u = i / obs, k = i % obs;
                d[S.o_w1 + u * S.w1s + k] = __ldcg(g + w1 + i);
            }
            if (tid < HID) { d[S.o_b1 + tid] = __ldcg(g + b1 + tid); d[S.o_b2 + tid] = __ldcg(g + b2 + tid); }
            for (int i = tid; i < no * HID; i += PPO_NT) d[S.o_w3 + i] = __ldcg(g + w3 + i);
            if (tid < 4) {
                d[S.o_b3 + tid] = tid < no ? __ldcg(g + b3 + tid) : 0.0f;
                d[S.o_ls + tid] = (n == 0 && tid < A) ? __ldcg(g + L.a_ls + tid) : 0.0f;
            }
        }
        __syncthreads();
    };
    // one-time zero fill of the padding the staging loop never touches (W1 columns >= obs, W2 pad columns, W3 rows >= n_out)
    for (int i = tid; i < 2 * S.net_floats; i += PPO_NT) smem[i] = 0.0f;
    if constexpr (CLUSTER) for (int i = tid; i < L.total + PPO_NSTAT; i += PPO_NT) gbuf[i] = 0.0f;     // layout padding is never published: keep it 0
    __syncthreads();
    load_params();

    // CLUSTER: this CTA owns the float4 groups [q0, q1) of the parameter vector; thread tid owns q0 + tid + 256 j.  Master weights and
    // Adam moments of those groups live in registers for the whole launch.
    constexpr int NPT = PPO_CLUSTER_NPT;                   // float4 groups per thread: checked on the host (grid kernel otherwise)
    float4 own_x[NPT], own_m[NPT], own_v[NPT]; int own_off[NPT];     // own_off: shared-memory offset of the group (-1: none, -2: not contiguous)
    const int c_nq = (L.total + 3) / 4, c_perq = (c_nq + G - 1) / G;
    const int c_q0 = cta * c_perq, c_q1 = min(c_nq, c_q0 + c_perq);
    if constexpr (CLUSTER) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            const int q = c_q0 + tid + PPO_NT * j;
            own_x[j] = own_m[j] = own_v[j] = make_float4(0.f, 0.f, 0.f, 0.f); own_off[j] = -1;
            if (q < c_q1) {
                own_x[j] = __ldcg(reinterpret_cast<const float4*>(P.params) + q);
                own_m[j] = __ldcg(reinterpret_cast<const float4*>(P.m) + q);
                own_v[j] = __ldcg(reinterpret_cast<const float4*>(P.v) + q);
                int o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = train_to_smem<HID, MS>(4 * q + i, L, S, obs, A);
                const bool contig = o[0] >= 0 && (o[0] & 3) == 0 && o[1] == o[0] + 1 && o[2] == o[0] + 2 && o[3] == o[0] + 3;
                const bool none = o[0] < 0 && o[1] < 0 && o[2] < 0 && o[3] < 0;
                own_off[j] = contig ? o[0] : none ? -1 : -2;
            }
        }
    }
#ifdef SCG_PPO_PROF
    long long pp_acc[16]; for (int i = 0; i < 16; ++i) pp_acc[i] = 0;
    long long pp_last = clock64();
#endif
    long long step_a = P.steps[0], step_c = P.steps[1];   // optimizer step counts: every CTA tracks them identically
    // thread 0: beta1 ** step, beta2 ** step of the actor / critic optimizers, advanced by one multiplication per step
    double bpow[4] = {1.0, 1.0, 1.0, 1.0};
    float stat[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // CTA 0, thread 0: sums over minibatches
    if (tid == 0) {
        bpow[0] = pow(P.beta1, (double)step_a); bpow[1] = pow(P.beta2, (double)step_a);
        bpow[2] = pow(P.beta1, (double)step_c); bpow[3] = pow(P.beta2, (double)step_c);
    }
    const int ts = t % SG, tu = t / SG;                // forward / input-gradient tile
    const int tk = t % TKN, tu2 = t / TKN;             // weight-gradient tile

    for (int ep = 0; ep < P.epochs; ++ep) {
        for (int mi = 0; mi < P.n_mb; ++mi) {
            // ---- gradient accumulators of this CTA for this minibatch
            float gW2[UT2][4], gW1[NW1], gW3[NW3], gb1 = 0.f, gb2 = 0.f, gb3 = 0.f;
#pragma unroll
            for (int j = 0; j < UT2; ++j) gW2[j][0] = gW2[j][1] = gW2[j][2] = gW2[j][3] = 0.f;
#pragma unroll
            for (int i = 0; i < NW1; ++i) gW1[i] = 0.f;
#pragma unroll
            for (int i = 0; i < NW3; ++i) gW3[i] = 0.f;
            float gls[PPO_MAXA] = {0.f, 0.f, 0.f, 0.f};      // thread 0 only
            float st_pl = 0.f, st_kl = 0.f, st_vl = 0.f;     // thread 0 only
            const float inv_mb = 1.0f / (float)mb;
            const int* perm = P.perm + (long long)ep * P.perm_stride + (long long)mi * mb;

            for (int c = cta; c < n_chunks; c += G) {
                const int base = c * MS;
                // ---- gather the chunk
                for (int i = tid; i < MS; i += PPO_NT) {
                    const int r = (base + i < mb) ? perm[base + i] : -1;
                    sRow[i] = r >= 0 ? (int)(r / P.block_rows) : -1;
                    sIdx[i] = r >= 0 ? (int)(r % P.block_rows) : 0;
                }
                __syncthreads();
                PP_T(0);
                for (int i = tid; i < MS * S.obsp; i += PPO_NT) {
                    const int s = i / S.obsp, k = i % S.obsp, r = sRow[s];
                    float v = 0.f;
                    if (r >= 0 && k < obs) v = __ldg(P.key[0] + r * P.block_stride[0] + (long long)sIdx[s] * obs + k);
                    sX[k * MSP + s] = v;
                }
                for (int i = tid; i < MS * A; i += PPO_NT) {
                    const int s = i / A, d = i % A, r = sRow[s];
                    sAct[d * MSP + s] = r >= 0 ? __ldg(P.key[1] + r * P.block_stride[1] + (long long)sIdx[s] * A + d) : 0.f;
                }
                for (int i = tid; i < MS * 4; i += PPO_NT) {
                    const int s = i / 4, q = i % 4, r = sRow[s];
                    sSc[q * MSP + s] = r >= 0 ? __ldg(P.key[2 + q] + r * P.block_stride[2 + q] + sIdx[s]) : 0.f;
                }
                __syncthreads();
                PP_T(1);
                // ---- forward (neural_networks.py:49-54)
                dense_forward<HID, MS, UT, TUN>(sp + S.o_w1, S.w1s, sp + S.o_b1, sX, S.obsp, sH1, tu, ts, act);
                __syncthreads();
                dense_forward<HID, MS, UT, TUN>(sp + S.o_w2, W2S, sp + S.o_b2, sH1, HID, sH2, tu, ts, act);
                __syncthreads();
                for (int i = t; i < nout * MS; i += PPO_NETT) {
                    const int o = i / MS, s = i % MS;
                    float a = sp[S.o_b3 + o];
                    const float* w = sp + S.o_w3 + o * HID;
#pragma unroll 8
                    for (int k = 0; k < HID; ++k) a = fmaf(w[k], sH2[k * MSP + s], a);
                    sD[o * MSP + s] = a;
                }
                __syncthreads();
                PP_T(2);
                // ---- losses and d(loss)/d(network outputs)
                float r_pl = 0.f, r_kl = 0.f, r_vl = 0.f, r_ls[PPO_MAXA] = {0.f, 0.f, 0.f, 0.f};
                if (t < MS) {
                    const int s = t;
                    const bool valid = sRow[s] >= 0;
                    if (net == 0) {
                        // compute_policy_loss (ppo_utils.py:56-69), Normal.log_prob (distributions.py:9-33)
                        float logp = 0.f, diff[PPO_MAXA], ivar[PPO_MAXA];
                        for (int d = 0; d < A; ++d) {
                            const float ls = sp[S.o_ls + d], sd = expf(ls), var = sd * sd;
                            diff[d] = sAct[d * MSP + s] - sD[d * MSP + s];
                            ivar[d] = 1.0f / var;
                            logp += -(diff[d] * diff[d]) / (2.0f * var) - logf(sd) - 0.91893853320467274f;
                        }
                        const float lp_old = sSc[0 * MSP + s], adv = sSc[1 * MSP + s];
                        const float ratio = expf(logp - lp_old);
                        const float lo = 1.0f - (float)P.clip, hi = 1.0f + (float)P.clip;
                        const float s1 = ratio * adv, s2 = fminf(fmaxf(ratio, lo), hi) * adv;
                        const float inside = (ratio >= lo && ratio <= hi) ? 1.0f : 0.0f;
                        const float gf = s1 < s2 ? 1.0f : (s1 > s2 ? inside : 0.5f + 0.5f * inside);    // torch.min / clamp sub-gradients
                        const float dlogp = valid ? -(adv * ratio * gf) * inv_mb : 0.0f;
                        for (int d = 0; d < A; ++d) {
                            sD[d * MSP + s] = dlogp * diff[d] * ivar[d];                                 // d/d mean
                            r_ls[d] = dlogp * (diff[d] * diff[d] * ivar[d] - 1.0f);                      // d/d logstd
                        }
                        if (valid) { r_pl = -fminf(s1, s2); r_kl = lp_old - logp; }
                    } else {
                        // compute_value_loss (ppo_utils.py:71-80)
                        const float v = sD[s], ret = sSc[2 * MSP + s];
                        float dv, vl;
                        if (P.clipped_value) {
                            const float vo = sSc[3 * MSP + s], cl = (float)P.clip;
                            const float dlt = v - vo, vc = vo + fminf(fmaxf(dlt, -cl), cl);
                            const float l1 = (v - ret) * (v - ret), l2 = (vc - ret) * (vc - ret);
                            const float inside = (dlt >= -cl && dlt <= cl) ? 1.0f : 0.0f;
                            vl = 0.5f * fmaxf(l1, l2);
                            dv = l1 > l2 ? (v - ret) : (l1 < l2 ? (vc - ret) * inside : 0.5f * (v - ret) + 0.5f * (vc - ret) * inside);
                        } else {
                            vl = 0.5f * (v - ret) * (v - ret);
                            dv = v - ret;
                        }
                        sD[s] = valid ? dv * inv_mb : 0.0f;
                        if (valid) r_vl = vl;
                    }
                }
                // block reduction of the scalar sums: warps 0..MS/32-1 of each half hold data
                {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        r_pl += __shfl_xor_sync(0xffffffffu, r_pl, o); r_kl += __shfl_xor_sync(0xffffffffu, r_kl, o);
                        r_vl += __shfl_xor_sync(0xffffffffu, r_vl, o);
#pragma unroll
                        for (int d = 0; d < PPO_MAXA; ++d) r_ls[d] += __shfl_xor_sync(0xffffffffu, r_ls[d], o);
                    }
                    const int w = tid / 32;
                    if ((tid & 31) == 0) {
                        sRed[w * 8 + 0] = r_pl; sRed[w * 8 + 1] = r_kl; sRed[w * 8 + 2] = r_vl; sRed[w * 8 + 3] = 0.f;
#pragma unroll
                        for (int d = 0; d < PPO_MAXA; ++d) sRed[w * 8 + 4 + d] = r_ls[d];
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 0; w < PPO_NT / 32; ++w) {      // fixed order
                        st_pl += sRed[w * 8 + 0]; st_kl += sRed[w * 8 + 1]; st_vl += sRed[w * 8 + 2];
                        for (int d = 0; d < PPO_MAXA; ++d) gls[d] += sRed[w * 8 + 4 + d];
                    }
                }
                PP_T(3);
                // ---- backward, layer 3: dW3[o][k] += sum_s dout[o][s] h2[k][s]; db3[o] += sum_s dout[o][s]
#pragma unroll
                for (int i = 0; i < NW3; ++i) {
                    const int idx = t + PPO_NETT * i;
                    if (idx < nout * HID) {
                        const int o = idx / HID, k = idx % HID;
                        float a = 0.f;
#pragma unroll 4
                        for (int s = 0; s < MS; s += 4) {
                            const float4 dd = ld4(sD + o * MSP + s), hh = ld4(sH2 + k * MSP + s);
                            a = fmaf(dd.x, hh.x, a); a = fmaf(dd.y, hh.y, a); a = fmaf(dd.z, hh.z, a); a = fmaf(dd.w, hh.w, a);
                        }
                        gW3[i] += a;
                    }
                }
                if (t < nout) {
                    float a = 0.f;
                    for (int s = 0; s < MS; ++s) a += sD[t * MSP + s];
                    gb3 += a;
                }
                __syncthreads();
                // ---- dz2 = (W3^T dout) * act'(h2), in place over h2
                {
                    float4 dd[PPO_MAXA];
#pragma unroll
                    for (int o = 0; o < PPO_MAXA; ++o) dd[o] = o < nout ? ld4(sD + o * MSP + 4 * ts) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < UT; ++j) {
                        const int u = tu + TUN * j;
                        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int o = 0; o < PPO_MAXA; ++o) {
                            const float w = sp[S.o_w3 + o * HID + u];       // rows >= nout are zero
                            g4.x = fmaf(w, dd[o].x, g4.x); g4.y = fmaf(w, dd[o].y, g4.y); g4.z = fmaf(w, dd[o].z, g4.z); g4.w = fmaf(w, dd[o].w, g4.w);
                        }
                        float4 h = ld4(sH2 + u * MSP + 4 * ts);
                        h.x = g4.x * act_deriv(h.x, act); h.y = g4.y * act_deriv(h.y, act); h.z = g4.z * act_deriv(h.z, act); h.w = g4.w * act_deriv(h.w, act);
                        st4(sH2 + u * MSP + 4 * ts, h);
                    }
                }
                __syncthreads();
                PP_T(4);
                // ---- dW2[u][k] += sum_s dz2[u][s] h1[k][s]   (u = tu2 + TUN2*j, k = tk + TKN*i)
#pragma unroll 2
                for (int s = 0; s < MS; s += 4) {
                    float4 hk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) hk[i] = ld4(sH1 + (tk + TKN * i) * MSP + s);
#pragma unroll
                    for (int j = 0; j < UT2; ++j) {
                        const float4 dz = ld4(sH2 + (tu2 + TUN2 * j) * MSP + s);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float a = gW2[j][i];
                            a = fmaf(dz.x, hk[i].x, a); a = fmaf(dz.y, hk[i].y, a); a = fmaf(dz.z, hk[i].z, a); a = fmaf(dz.w, hk[i].w, a);
                            gW2[j][i] = a;
                        }
                    }
                }
                if (t < HID) {
                    float a = 0.f;
#pragma unroll 4
                    for (int s = 0; s < MS; s += 4) { const float4 dz = ld4(sH2 + t * MSP + s); a += (dz.x + dz.y) + (dz.z + dz.w); }
                    gb2 += a;
                }
                __syncthreads();
                PP_T(5);
                // ---- dz1 = (W2^T dz2) * act'(h1), in place over h1   (k = UT*tu + j contiguous, 4 samples)
                {
                    float acc[UT][4];
#pragma unroll
                    for (int j = 0; j < UT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 2
                    for (int u = 0; u < HID; ++u) {
                        const float4 dz = ld4(sH2 + u * MSP + 4 * ts);
                        const float* wrow = sp + S.o_w2 + u * W2S + UT * tu;
#pragma unroll
                        for (int j = 0; j < UT; ++j) {
                            const float w = wrow[j];
                            acc[j][0] = fmaf(w, dz.x, acc[j][0]); acc[j][1] = fmaf(w, dz.y, acc[j][1]); acc[j][2] = fmaf(w, dz.z, acc[j][2]); acc[j][3] = fmaf(w, dz.w, acc[j][3]);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < UT; ++j) {
                        const int k = UT * tu + j;
                        float4 h = ld4(sH1 + k * MSP + 4 * ts);
                        h.x = acc[j][0] * act_deriv(h.x, act); h.y = acc[j][1] * act_deriv(h.y, act); h.z = acc[j][2] * act_deriv(h.z, act); h.w = acc[j][3] * act_deriv(h.w, act);
                        st4(sH1 + k * MSP + 4 * ts, h);
                    }
                }
                __syncthreads();
                PP_T(6);
                // ---- dW1[u][k] += sum_s dz1[u][s] x[k][s]; db1
#pragma unroll
                for (int i = 0; i < NW1; ++i) {
                    const int idx = t + PPO_NETT * i;
                    if (idx < HID * obs) {
                        const int u = (int)__umulhi((unsigned)idx, obs_magic), k = idx - u * obs;      // idx / obs, idx % obs (idx < 2^16)
                        float a = 0.f;
#pragma unroll 4
                        for (int s = 0; s < MS; s += 4) {
                            const float4 dz = ld4(sH1 + u * MSP + s), xx = ld4(sX + k * MSP + s);
                            a = fmaf(dz.x, xx.x, a); a = fmaf(dz.y, xx.y, a); a = fmaf(dz.z, xx.z, a); a = fmaf(dz.w, xx.w, a);
                        }
                        gW1[i] += a;
                    }
                }
                if (t < HID) {
                    float a = 0.f;
#pragma unroll 4
                    for (int s = 0; s < MS; s += 4) { const float4 dz = ld4(sH1 + t * MSP + s); a += (dz.x + dz.y) + (dz.z + dz.w); }
                    gb1 += a;
                }
                __syncthreads();
                PP_T(7);
            }   // chunks

            // ---- publish this CTA's partial gradient (train layout) and loss sums
            {
                float* w = CLUSTER ? gbuf : P.ws + (size_t)cta * P.ws_pitch;
                auto put = [](float* dst, float val) { if constexpr (CLUSTER) *dst = val; else __stcg(dst, val); };   // shared vs global (L2) destination
#pragma unroll
                for (int j = 0; j < UT2; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) put(w + g_w2 + (tu2 + TUN2 * j) * HID + (tk + TKN * i), gW2[j][i]);
#pragma unroll
                for (int i = 0; i < NW1; ++i) { const int idx = t + PPO_NETT * i; if (idx < HID * obs) put(w + g_w1 + idx, gW1[i]); }
#pragma unroll
                for (int i = 0; i < NW3; ++i) { const int idx = t + PPO_NETT * i; if (idx < nout * HID) put(w + g_w3 + idx, gW3[i]); }
                if (t < HID) { put(w + g_b1 + t, gb1); put(w + g_b2 + t, gb2); }
                if (t < nout) put(w + g_b3 + t, gb3);
                if (tid == 0) {
                    for (int d = 0; d < A; ++d) put(w + L.a_ls + d, gls[d]);
                    put(w + L.total + 0, st_pl); put(w + L.total + 1, st_vl); put(w + L.total + 2, st_kl);
                }
            }
            PP_T(8);
            if constexpr (CLUSTER) cluster.sync();
            else { __threadfence(); if (G > 1) grid.sync(); else __syncthreads(); }
            PP_T(9);

            // ---- reduce over CTAs (fixed order) + KL gate + Adam on this CTA's slice
            {
                __shared__ double s_gate[8];
                if (tid < 32) {
                    float pl = 0.f, vl = 0.f, kl = 0.f;
                    const int Gu = n_chunks < G ? n_chunks : G;       // CTAs that had work
                    for (int c2 = tid; c2 < Gu; c2 += 32) {
                        if constexpr (CLUSTER) {
                            const float* w = cluster.map_shared_rank(gbuf, c2) + L.total;
                            pl += w[0]; vl += w[1]; kl += w[2];
                        } else {
                            const float* w = P.ws + (size_t)c2 * P.ws_pitch + L.total;
                            pl += __ldcg(w + 0); vl += __ldcg(w + 1); kl += __ldcg(w + 2);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {                // fixed tree: identical in every CTA
                        pl += __shfl_xor_sync(0xffffffffu, pl, o); vl += __shfl_xor_sync(0xffffffffu, vl, o); kl += __shfl_xor_sync(0xffffffffu, kl, o);
                    }
                    if (tid == 0) {
                        pl *= inv_mb; vl *= inv_mb; kl *= inv_mb;         // .mean() over the minibatch
                        const bool apply_actor = !(P.target_kl > 0.0) || ((double)kl <= 1.5 * P.target_kl);     // ppo_utils.py:128
                        if (apply_actor) { bpow[0] *= P.beta1; bpow[1] *= P.beta2; }     // beta ** step of the actor optimizer
                        bpow[2] *= P.beta1; bpow[3] *= P.beta2;
                        s_gate[0] = apply_actor ? 1.0 : 0.0;
                        s_gate[1] = P.lr_a / (1.0 - bpow[0]);      // step_size = lr / bias_correction1
                        s_gate[2] = sqrt(1.0 - bpow[1]);           // bias_correction2_sqrt
                        s_gate[3] = P.lr_c / (1.0 - bpow[2]);
                        s_gate[4] = sqrt(1.0 - bpow[3]);
                        if (cta == 0) {
                            float ent = 0.f;
                            for (int d = 0; d < A; ++d) ent += 0.5f + 0.91893853320467274f + logf(expf(smem[S.o_ls + d]));
                            stat[0] += pl; stat[1] += vl; stat[2] += -ent; stat[3] += kl;
                            stat[4] += apply_actor ? 1.f : 0.f; stat[5] += 1.f;
                        }
                    }
                }
                __syncthreads();
                const bool apply_actor = s_gate[0] != 0.0;
                const int Gu = n_chunks < G ? n_chunks : G;
                const float b2f = (float)P.beta2, w1f = (float)(1.0 - P.beta1), w2f = (float)(1.0 - P.beta2);
                const float epsf = (float)P.eps;
                const bool want_grad = P.grad_out && ep == P.epochs - 1 && mi == P.n_mb - 1;
                auto adam = [&](float& mm, float& vv, float& xx, float gg, float step_size, float bc2s) {
                    mm = fmaf(w1f, gg - mm, mm);                           // exp_avg.lerp_(grad, 1 - beta1)
                    vv = fmaf(w2f * gg, gg, b2f * vv);                     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                    const float denom = sqrtf(vv) / bc2s + epsf;
                    xx = xx - step_size * (mm / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
                };
                if constexpr (CLUSTER) {
#pragma unroll
                    for (int j = 0; j < NPT; ++j) {
                        const int q = c_q0 + tid + PPO_NT * j;
                        if (q >= c_q1) continue;
                        // the 8 remote (distributed shared memory) loads are independent: issue them together, then sum in a fixed order
                        float4 w[8];
#pragma unroll
                        for (int c2 = 0; c2 < 8; ++c2)
                            w[c2] = c2 < Gu ? *(reinterpret_cast<const float4*>(cluster.map_shared_rank(gbuf, c2)) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        float4 g = w[0];
#pragma unroll
                        for (int c2 = 1; c2 < 8; ++c2) { g.x += w[c2].x; g.y += w[c2].y; g.z += w[c2].z; g.w += w[c2].w; }
                        const int p = 4 * q;
                        if (p == L.a_ls) {                                   // d(entropy_coef * entropy_loss)/d logstd = -entropy_coef
                            const float e = -(float)P.ent_coef;
                            g.x += e; if (A > 1) g.y += e; if (A > 2) g.z += e; if (A > 3) g.w += e;
                        }
                        if (want_grad) reinterpret_cast<float4*>(P.grad_out)[q] = g;
                        const bool actor = p < L.critic0;
                        if (actor && !apply_actor) continue;
                        const float step_size = (float)(actor ? s_gate[1] : s_gate[3]), bc2s = (float)(actor ? s_gate[2] : s_gate[4]);
                        adam(own_m[j].x, own_v[j].x, own_x[j].x, g.x, step_size, bc2s); adam(own_m[j].y, own_v[j].y, own_x[j].y, g.y, step_size, bc2s);
                        adam(own_m[j].z, own_v[j].z, own_x[j].z, g.z, step_size, bc2s); adam(own_m[j].w, own_v[j].w, own_x[j].w, g.w, step_size, bc2s);
                        if (own_off[j] >= 0) {                               // the group is contiguous in the padded shared layout: one 16-byte store per CTA
#pragma unroll
                            for (int c2 = 0; c2 < 8; ++c2) if (c2 < G) *reinterpret_cast<float4*>(cluster.map_shared_rank(smem, c2) + own_off[j]) = own_x[j];
                        } else if (own_off[j] == -2) {                       // straddles a row / section boundary: element by element
                            const float xs[4] = {own_x[j].x, own_x[j].y, own_x[j].z, own_x[j].w};
                            for (int i = 0; i < 4; ++i) {
                                const int o = train_to_smem<HID, MS>(p + i, L, S, obs, A);
                                if (o >= 0) for (int c2 = 0; c2 < G; ++c2) cluster.map_shared_rank(smem, c2)[o] = xs[i];
                            }
                        }
                    }
                } else {
                const int per = (((L.total + G - 1) / G) + 3) & ~3;
                const int p0 = cta * per, p1 = min(L.total, p0 + per);
                const float4* __restrict__ ws4 = reinterpret_cast<const float4*>(P.ws);
                float4* __restrict__ m4p = reinterpret_cast<float4*>(P.m);
                float4* __restrict__ v4p = reinterpret_cast<float4*>(P.v);
                float4* __restrict__ x4p = reinterpret_cast<float4*>(P.params);
                const int pitch4 = P.ws_pitch / 4;
#pragma unroll 2
                for (int q = p0 / 4 + tid; q < p1 / 4; q += PPO_NT) {
                    const int p = 4 * q;
                    const bool actor = p < L.critic0;
                    float4 m = __ldcg(m4p + q), v = __ldcg(v4p + q), x = __ldcg(x4p + q);
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int c2 = 0; c2 < Gu; ++c2) {
                        const float4 w = __ldcg(ws4 + (size_t)c2 * pitch4 + q);
                        g.x += w.x; g.y += w.y; g.z += w.z; g.w += w.w;
                    }
                    if (p == L.a_ls) {                                   // d(entropy_coef * entropy_loss)/d logstd = -entropy_coef
                        const float e = -(float)P.ent_coef;
                        g.x += e; if (A > 1) g.y += e; if (A > 2) g.z += e; if (A > 3) g.w += e;
                    }
                    if (want_grad) reinterpret_cast<float4*>(P.grad_out)[q] = g;
                    if (actor && !apply_actor) continue;
                    const float step_size = (float)(actor ? s_gate[1] : s_gate[3]), bc2s = (float)(actor ? s_gate[2] : s_gate[4]);
                    adam(m.x, v.x, x.x, g.x, step_size, bc2s); adam(m.y, v.y, x.y, g.y, step_size, bc2s);
                    adam(m.z, v.z, x.z, g.z, step_size, bc2s); adam(m.w, v.w, x.w, g.w, step_size, bc2s);
                    __stcg(m4p + q, m); __stcg(v4p + q, v); __stcg(x4p + q, x);
                }
                }
                step_a += apply_actor ? 1 : 0;
                step_c += 1;
            }
            PP_T(10);
            if constexpr (CLUSTER) cluster.sync();          // every CTA's shared-memory weights are up to date (remote stores), gbuf may be overwritten
            else { __threadfence(); if (G > 1) grid.sync(); else __syncthreads(); load_params(); }
            PP_T(11);
        }
    }
    if constexpr (CLUSTER) {                                // the slice owners write weights and moments back once per launch
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            const int q = c_q0 + tid + PPO_NT * j;
            if (q < c_q1) { reinterpret_cast<float4*>(P.params)[q] = own_x[j]; reinterpret_cast<float4*>(P.m)[q] = own_m[j]; reinterpret_cast<float4*>(P.v)[q] = own_v[j]; }
        }
    }
#ifdef SCG_PPO_PROF
    if (cta == 0 && tid == 0) {
        const double n = (double)P.epochs * P.n_mb;
        printf("[SCG_PPO_PROF] cycles / minibatch (MS=%d cluster=%d G=%d): idx=%.0f gather=%.0f fwd=%.0f loss=%.0f dW3+dz2=%.0f dW2=%.0f dz1=%.0f dW1=%.0f publish=%.0f sync1=%.0f reduce+adam=%.0f sync2(+reload)=%.0f\n",
               MS, (int)CLUSTER, G, pp_acc[0] / n, pp_acc[1] / n, pp_acc[2] / n, pp_acc[3] / n, pp_acc[4] / n, pp_acc[5] / n, pp_acc[6] / n, pp_acc[7] / n, pp_acc[8] / n, pp_acc[9] / n, pp_acc[10] / n, pp_acc[11] / n);
    }
#endif
    if (cta == 0 && tid == 0) {
        P.steps[0] = step_a; P.steps[1] = step_c;
        for (int i = 0; i < 6; ++i) P.stats[i] += stat[i];
    }
}

// train layout -> packed rollout blob (scg_kernels.cuh: blob_layout)
__global__ void pack_blob_kernel(int obs, int act, int hid, const float* __restrict__ p, float* __restrict__ blob) {
    const TrainLayout T = train_layout(obs, act, hid);
    const BlobLayout B = blob_layout(obs, act, hid);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.total; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < B.b1) {                       // W1P [k/2][2H][2]
            const int j = i - B.w1t, par = j & 1, uu = (j >> 1) % (2 * hid), kp = (j >> 1) / (2 * hid), k = 2 * kp + par;
            v = uu < hid ? p[T.a_w1 + uu * obs + k] : p[T.c_w1 + (uu - hid) * obs + k];
        } else if (i < B.w2t) {
            const int u = i - B.b1;
            v = u < hid ? p[T.a_b1 + u] : p[T.c_b1 + u - hid];
        } else if (i < B.b2) {                // W2P [net][k/2][H][2]
            const int j = i - B.w2t, n = j / (hid * hid), r = j % (hid * hid), par = r & 1, u = (r >> 1) % hid, kp = (r >> 1) / hid, k = 2 * kp + par;
            v = p[(n ? T.c_w2 : T.a_w2) + u * hid + k];
        } else if (i < B.w3) {
            const int u = i - B.b2;
            v = u < hid ? p[T.a_b2 + u] : p[T.c_b2 + u - hid];
        } else if (i < B.b3) {                // W3 [A+1][H]
            const int j = i - B.w3, o = j / hid, k = j % hid;
            v = o < act ? p[T.a_w3 + o * hid + k] : p[T.c_w3 + k];
        } else if (i < B.logstd) {
            const int o = i - B.b3;
            v = o < act ? p[T.a_b3 + o] : (o == act ? p[T.c_b3] : 0.f);
        } else {
            const int o = i - B.logstd;
            v = o < act ? p[T.a_ls + o] : 0.f;
        }
        blob[i] = v;
    }
}

int ppo_fail(int code, const std::string& msg) {
    scg_internal_set_error(msg);
    return code;
}

template <int HID, int MS>
int ppo_launch(const scg_ppo_args* a, PpoParams& P, int G, cudaStream_t st) {
    const PpoSmem<HID, MS> S(a->obs_dim);
    const size_t smem = sizeof(float) * S.total;
    cudaError_t e = cudaFuncSetAttribute(ppo_update_kernel<HID, MS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ppo_fail(SCG_ECUDA, std::string("scg_ppo_update: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    void* kargs[] = {(void*)&P};
    e = cudaLaunchCooperativeKernel((void*)ppo_update_kernel<HID, MS, false>, dim3(G), dim3(PPO_NT), kargs, smem, st);
    if (e != cudaSuccess) return ppo_fail(SCG_ECUDA, std::string("scg_ppo_update: launch: ") + cudaGetErrorString(e));
    return SCG_OK;
}

// small minibatches: ONE thread-block cluster of C CTAs (C = 1, 2, 4 or 8), MS samples per CTA and chunk
constexpr int PPO_CLUSTER_MAX_MB = 256;
template <int HID, int MS>
int ppo_launch_cluster(const scg_ppo_args* a, PpoParams& P, cudaStream_t st) {
    const PpoSmem<HID, MS> S(a->obs_dim);
    const TrainLayout L = train_layout(a->obs_dim, a->act_dim, HID);
    const int chunks = (a->mini_batch_size + MS - 1) / MS;
    int C = 1; while (C < 8 && C < chunks) C *= 2;
    const int nq = (L.total + 3) / 4;
    if ((nq + C - 1) / C > PPO_CLUSTER_NPT * PPO_NT) C = 8;                  // the slice of a CTA must fit NPT float4 groups per thread
    if ((nq + C - 1) / C > PPO_CLUSTER_NPT * PPO_NT) return -100;            // does not fit: the caller uses the grid kernel
    const size_t smem = sizeof(float) * ((size_t)S.total + L.total + PPO_NSTAT);
    auto kern = ppo_update_kernel<HID, MS, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return ppo_fail(SCG_ECUDA, std::string("scg_ppo_update: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(PPO_NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, P);
    if (e != cudaSuccess) return ppo_fail(SCG_ECUDA, std::string("scg_ppo_update: cluster launch: ") + cudaGetErrorString(e));
    return SCG_OK;
}

int ppo_chunk(int hid) { return hid == 128 ? 32 : 64; }

int ppo_grid(const scg_ppo_args* a) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    const int ms = ppo_chunk(a->hidden_dim);
    const int chunks = (a->mini_batch_size + ms - 1) / ms;
    return chunks < sms ? chunks : sms;
}

int ppo_validate(const scg_ppo_args* a) {
    if (!a) return ppo_fail(SCG_EINVAL, "scg_ppo_update: null argument");
    if (a->hidden_dim != 32 && a->hidden_dim != 64) return ppo_fail(SCG_EINVAL, "scg_ppo_update: hidden_dim must be 32 or 64");
    if (a->obs_dim < 1 || a->obs_dim > 32) return ppo_fail(SCG_EINVAL, "scg_ppo_update: obs_dim must be in [1, 32]");
    if (a->act_dim < 1 || a->act_dim > PPO_MAXA) return ppo_fail(SCG_EINVAL, "scg_ppo_update: act_dim must be in [1, 4]");
    if (a->activation < SCG_ACT_TANH || a->activation > SCG_ACT_LEAKY_RELU) return ppo_fail(SCG_EINVAL, "scg_ppo_update: unknown activation");
    if (a->mini_batch_size < 1 || a->num_mini_batch < 1 || a->opt_epochs < 1)
        return ppo_fail(SCG_EINVAL, "scg_ppo_update: mini_batch_size, num_mini_batch (ppo_utils.py:94 'num_mini_batch is 0') and opt_epochs must be >= 1");
    if (a->block_rows < 1) return ppo_fail(SCG_EINVAL, "scg_ppo_update: block_rows must be >= 1");
    if (a->perm_stride < (int64_t)a->mini_batch_size * a->num_mini_batch) return ppo_fail(SCG_EINVAL, "scg_ppo_update: perm_stride < mini_batch_size * num_mini_batch");
    return SCG_OK;
}

}  // namespace

extern "C" {

int scg_ppo_param_layout(int32_t obs, int32_t act, int32_t hid, int32_t o[15]) {
    if (!o || obs < 1 || act < 1 || hid < 1) return SCG_EINVAL;
    const TrainLayout L = train_layout(obs, act, hid);
    const int v[15] = {L.a_w1, L.a_b1, L.a_w2, L.a_b2, L.a_w3, L.a_b3, L.a_ls, L.c_w1, L.c_b1, L.c_w2, L.c_b2, L.c_w3, L.c_b3, L.critic0, L.total};
    for (int i = 0; i < 15; ++i) o[i] = v[i];
    return SCG_OK;
}

size_t scg_ppo_workspace_bytes(const scg_ppo_args* a) {
    if (ppo_validate(a) != SCG_OK) return 0;
    const int G = ppo_grid(a);
    if (G < 1) return 0;
    const TrainLayout L = train_layout(a->obs_dim, a->act_dim, a->hidden_dim);
    return sizeof(float) * (size_t)G * (size_t)(L.total + PPO_NSTAT);
}

int scg_ppo_update(const scg_ppo_args* a, void* stream) {
    int rc = ppo_validate(a);
    if (rc != SCG_OK) return rc;
    for (int k = 0; k < 6; ++k)
        if (!a->d_key[k]) return ppo_fail(SCG_EINVAL, "scg_ppo_update: null record pointer");
    if (!a->d_perm || !a->d_params || !a->d_exp_avg || !a->d_exp_avg_sq || !a->d_adam_steps || !a->d_stats || !a->d_workspace)
        return ppo_fail(SCG_EINVAL, "scg_ppo_update: null pointer");
    const int G = ppo_grid(a);
    if (G < 1) return ppo_fail(SCG_ECUDA, "scg_ppo_update: no CUDA device");
    const TrainLayout L = train_layout(a->obs_dim, a->act_dim, a->hidden_dim);
    if (a->workspace_bytes < sizeof(float) * (size_t)G * (size_t)(L.total + PPO_NSTAT)) return ppo_fail(SCG_EINVAL, "scg_ppo_update: workspace too small");
    PpoParams P;
    P.obs = a->obs_dim; P.act = a->act_dim; P.hid = a->hidden_dim; P.activation = a->activation;
    P.mb = a->mini_batch_size; P.n_mb = a->num_mini_batch; P.epochs = a->opt_epochs; P.clipped_value = a->use_clipped_value;
    P.block_rows = a->block_rows;
    for (int k = 0; k < 6; ++k) { P.block_stride[k] = a->block_stride[k]; P.key[k] = a->d_key[k]; }
    P.perm = a->d_perm; P.perm_stride = a->perm_stride;
    P.clip = a->clip_param; P.target_kl = a->target_kl; P.ent_coef = a->entropy_coef;
    P.lr_a = a->actor_lr; P.lr_c = a->critic_lr; P.beta1 = a->beta1; P.beta2 = a->beta2; P.eps = a->adam_eps;
    P.params = a->d_params; P.m = a->d_exp_avg; P.v = a->d_exp_avg_sq; P.steps = (long long*)a->d_adam_steps;
    P.stats = a->d_stats; P.ws = a->d_workspace; P.ws_pitch = L.total + PPO_NSTAT; P.grad_out = a->d_grad_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (a->mini_batch_size <= PPO_CLUSTER_MAX_MB && !getenv("SCG_PPO_NO_CLUSTER")) {
        const int rc2 = a->hidden_dim == 64 ? ppo_launch_cluster<64, 8>(a, P, st) : ppo_launch_cluster<32, 16>(a, P, st);
        if (rc2 != -100) return rc2;
    }
    if (a->hidden_dim == 64) return ppo_launch<64, 64>(a, P, G, st);
    return ppo_launch<32, 64>(a, P, G, st);
}

int scg_policy_pack_device(int32_t obs, int32_t act, int32_t hid, const float* d_params, float* d_blob, void* stream) {
    if (!d_params || !d_blob || obs < 2 || (obs & 1) || (hid & 3) || act < 1) return ppo_fail(SCG_EINVAL, "scg_policy_pack_device: bad argument");
    const BlobLayout B = blob_layout(obs, act, hid);
    pack_blob_kernel<<<(B.total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(obs, act, hid, d_params, d_blob);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ppo_fail(SCG_ECUDA, std::string("scg_policy_pack_device: ") + cudaGetErrorString(e));
    return SCG_OK;
}

}  // extern "C"
