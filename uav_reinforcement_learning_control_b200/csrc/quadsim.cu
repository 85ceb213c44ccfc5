// quadsim.cu -- C-ABI entry points of libquadsim.so (include/quadsim_abi.h) and kernel dispatch.
// Built for sm_100a only; there is no CPU code path behind any of these calls.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <new>

#include "qs_kernels.cuh"
#include "qs_step2.cuh"
#include "qs_rollout.cuh"
#include "qs_rollout_tc.cuh"
#include "qs_ppo.cuh"
#include "qs_traj.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    if (ce != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

#define QS_CUDA(call)                                                      \
    do {                                                                   \
        cudaError_t ce_ = (call);                                          \
        if (ce_ != cudaSuccess) return fail(QS_ECUDA, #call, ce_);         \
    } while (0)

inline int nblocks(int n, int block) { return (n + block - 1) / block; }

// Every handle-taking entry point runs on the handle's device and leaves the caller's current device as it found it
// (two engines on different GPUs in one process, or an Engine(device=1) used after another cudaSetDevice).
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace

#ifndef QS_HOST_STREAMS
#define QS_HOST_STREAMS 5
#endif

struct QsEngine {
    QsParams P;
    int32_t n;
    int device;
    float* target_table;    // device, [max_episode_steps][3] or null
    double* waypoints;      // device, [shapes][QS_MAX_WP][3] or null
    float* scratch;         // device staging for qs_step_host: action | obs | reward | done
    cudaStream_t hs[QS_HOST_STREAMS];     // qs_step_host: copy/compute streams (H2D + kernel of later chunks under the D2H of earlier ones)
    cudaEvent_t hev[QS_HOST_STREAMS + 1];
    qs::Tables tables() const { return qs::Tables{target_table, waypoints}; }
};

#define QS_DISPATCH_MODE(mode, CALL)                                          \
    switch (mode) {                                                           \
        case QS_MODE_MJX_BRAX: { constexpr int M_ = QS_MODE_MJX_BRAX; CALL; break; }             \
        case QS_MODE_HOVER_GYM: { constexpr int M_ = QS_MODE_HOVER_GYM; CALL; break; }           \
        case QS_MODE_TRAJ_GYM: { constexpr int M_ = QS_MODE_TRAJ_GYM; CALL; break; }             \
        case QS_MODE_HOVER_BRAX: { constexpr int M_ = QS_MODE_HOVER_BRAX; CALL; break; }         \
        case QS_MODE_MJX_PLAYGROUND: { constexpr int M_ = QS_MODE_MJX_PLAYGROUND; CALL; break; } \
        default: return fail(QS_EINVAL, "unknown mode");                      \
    }

static int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return fail(QS_ECUDA, what, ce);
    return QS_OK;
}

extern "C" {

int qs_abi_version(void) { return QS_ABI_VERSION; }
const char* qs_last_error_string(void) { return g_err; }
int qs_params_size(void) { return (int)sizeof(QsParams); }
uint64_t qs_launch_count(void) { return g_launches.load(); }

int qs_create(const QsParams* params, int32_t num_envs, int32_t device, const float* target_table_host,
              const double* waypoints_host, QsHandle* out) {
    if (!params || !out || num_envs <= 0) return fail(QS_EINVAL, "qs_create: null argument or num_envs <= 0");
    const QsParams& P = *params;
    if (P.mode < 0 || P.mode > QS_MODE_MJX_PLAYGROUND) return fail(QS_EINVAL, "qs_create: bad mode");
    const bool gym = (P.mode == QS_MODE_HOVER_GYM || P.mode == QS_MODE_TRAJ_GYM);
    if (P.obs_dim != (gym ? 12 : 21)) return fail(QS_EINVAL, "qs_create: obs_dim does not match mode");
    const bool table = (P.mode == QS_MODE_MJX_BRAX || P.mode == QS_MODE_MJX_PLAYGROUND);
    if (table && (!target_table_host || P.max_episode_steps <= 0))
        return fail(QS_EINVAL, "qs_create: mjx modes need a target table");
    if (P.waypoint_mode) {
        if (!gym || !waypoints_host || P.wp_num_shapes <= 0 || P.wp_num_shapes > QS_MAX_SHAPES)
            return fail(QS_EINVAL, "qs_create: bad waypoint configuration");
        for (int s = 0; s < P.wp_num_shapes; ++s)
            if (P.wp_count[s] <= 0 || P.wp_count[s] > QS_MAX_WP) return fail(QS_EINVAL, "qs_create: bad wp_count");
    }
    if (!(P.dt > 0.f) || !(P.mass > 0.f)) return fail(QS_EINVAL, "qs_create: dt and mass must be positive");
    int ndev = 0;
    QS_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(QS_EINVAL, "qs_create: no such device");
    DeviceGuard guard(device);              // the caller's current device is restored on return
    cudaDeviceProp prop;
    QS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(QS_EUNSUPPORTED, "qs_create: libquadsim is built for sm_100a (B200) only");
    QsEngine* e = new (std::nothrow) QsEngine();
    if (!e) return fail(QS_ENOMEM, "qs_create: host allocation failed");
    e->P = P; e->n = num_envs; e->device = device; e->target_table = nullptr; e->waypoints = nullptr; e->scratch = nullptr;
    for (int k = 0; k < QS_HOST_STREAMS; ++k) e->hs[k] = nullptr;
    for (int k = 0; k <= QS_HOST_STREAMS; ++k) e->hev[k] = nullptr;
    if (table) {
        const size_t bytes = (size_t)P.max_episode_steps * 3 * sizeof(float);
        if (cudaMalloc(&e->target_table, bytes) != cudaSuccess) { delete e; return fail(QS_ENOMEM, "cudaMalloc target table"); }
        cudaMemcpy(e->target_table, target_table_host, bytes, cudaMemcpyHostToDevice);
    }
    if (P.waypoint_mode) {
        const size_t bytes = (size_t)P.wp_num_shapes * QS_MAX_WP * 3 * sizeof(double);
        if (cudaMalloc(&e->waypoints, bytes) != cudaSuccess) { cudaFree(e->target_table); delete e; return fail(QS_ENOMEM, "cudaMalloc waypoints"); }
        cudaMemcpy(e->waypoints, waypoints_host, bytes, cudaMemcpyHostToDevice);
    }
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) { cudaFree(e->target_table); cudaFree(e->waypoints); delete e; return fail(QS_ECUDA, "qs_create", ce); }
    *out = e;
    return QS_OK;
}

int qs_destroy(QsHandle h) {
    if (!h) return QS_OK;
    DeviceGuard guard(h->device);
    cudaFree(h->target_table); cudaFree(h->waypoints); cudaFree(h->scratch);
    for (int k = 0; k < QS_HOST_STREAMS; ++k) if (h->hs[k]) cudaStreamDestroy(h->hs[k]);
    for (int k = 0; k <= QS_HOST_STREAMS; ++k) if (h->hev[k]) cudaEventDestroy(h->hev[k]);
    delete h;
    return QS_OK;
}

int qs_num_envs(QsHandle h) { return h ? h->n : QS_EINVAL; }

int qs_get_params(QsHandle h, QsParams* out) {
    if (!h || !out) return fail(QS_EINVAL, "qs_get_params: null");
    *out = h->P;
    return QS_OK;
}

int qs_reset(QsHandle h, float* state, const uint8_t* mask, float* obs, float* first_state, void* stream) {
    if (!h || !state) return fail(QS_EINVAL, "qs_reset: null");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    QS_DISPATCH_MODE(h->P.mode, (qs::reset_kernel<M_><<<nblocks(h->n, qs::kBlock), qs::kBlock, 0, s>>>(
        h->P, h->tables(), h->n, state, mask, obs, first_state)));
    return check_launch("reset_kernel");
}

static int launch_step(QsHandle h, int lo, int count, float* state, const float* action, float* obs, float* reward,
                       float* done, float* truncated, float* metrics, float* terminal_obs, const float* first_state,
                       cudaStream_t s) {
    // Programmatic dependent launch (QS_USE_PDL): the step kernel signals launch_dependents at its top and waits for
    // its predecessors (griddepcontrol.wait) before its first global access, so the launch latency, CTA ramp-up and
    // parameter set-up of step k+1 overlap the drain of whatever ran before it on the stream.
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3((unsigned)nblocks(count, qs::kBlock)); lc.blockDim = dim3(qs::kBlock); lc.dynamicSmemBytes = 0; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = QS_USE_PDL ? 1 : 0;
    lc.attrs = at; lc.numAttrs = 1;
    const float4* a4 = (const float4*)action;
    const QsParams& Pq = h->P;
    if (Pq.mode == QS_MODE_HOVER_GYM && !Pq.battery && !Pq.rate_wrapper && !Pq.waypoint_mode && !Pq.pre_clip_action &&
        !metrics && !terminal_obs) {
        // plain north-star configuration: feature-folded instantiation (qs_env.cuh: FeatLean)
        // two adjacent envs per thread on the packed FP32 pipe (qs_step2.cuh) whenever the pairs are 8-byte aligned; an
        // odd last env (and any unaligned call) goes through the one-env-per-thread kernel
        static const int use_step2 = getenv("QS_STEP2") ? atoi(getenv("QS_STEP2")) : 1;      // A/B knob
        if (use_step2 && (h->n & 1) == 0 && (lo & 1) == 0 && count >= 2 &&
            (((uintptr_t)state | (uintptr_t)reward | (uintptr_t)done | (uintptr_t)truncated) & 7u) == 0) {
            const int npairs = count / 2;
            cudaLaunchConfig_t l2 = lc;
            l2.gridDim = dim3((unsigned)nblocks(npairs, qs::kBlock2)); l2.blockDim = dim3(qs::kBlock2);
            cudaLaunchKernelEx(&l2, qs::step2_kernel, h->P, (int)h->n, lo / 2, npairs, state, a4, obs, reward, done, truncated);
            if ((count & 1) == 0) return check_launch("step2_kernel");
            g_launches.fetch_add(1, std::memory_order_relaxed);
            lc.gridDim = dim3(1u);
            lo += 2 * npairs; count = 1;
        }
        cudaLaunchKernelEx(&lc, qs::step_kernel<QS_MODE_HOVER_GYM, qs::FeatLean>, h->P, h->tables(), (int)h->n, lo, count, state, a4,
                           obs, reward, done, truncated, metrics, terminal_obs, first_state);
        return check_launch("step_kernel<lean>");
    }
    QS_DISPATCH_MODE(h->P.mode, (cudaLaunchKernelEx(&lc, qs::step_kernel<M_, qs::FeatAll>, h->P, h->tables(), (int)h->n, lo, count,
        state, a4, obs, reward, done, truncated, metrics, terminal_obs, first_state)));
    return check_launch("step_kernel");
}

int qs_step(QsHandle h, float* state, const float* action, float* obs, float* reward, float* done,
            float* truncated, float* metrics, float* terminal_obs, const float* first_state, void* stream) {
    if (!h || !state || !action || !obs || !reward || !done) return fail(QS_EINVAL, "qs_step: null");
    DeviceGuard guard(h->device);
    if (h->P.auto_reset == QS_RESET_RESTORE_FIRST && !first_state)
        return fail(QS_EINVAL, "qs_step: auto_reset=restore_first needs first_state");
    if (((uintptr_t)action & 15u) != 0) return fail(QS_EINVAL, "qs_step: action must be 16-byte aligned");
    if (h->P.obs_dim == 12 && ((uintptr_t)obs & 15u) != 0) return fail(QS_EINVAL, "qs_step: obs must be 16-byte aligned");
    return launch_step(h, 0, h->n, state, action, obs, reward, done, truncated, metrics, terminal_obs, first_state,
                       (cudaStream_t)stream);
}

namespace qs {
// one thread per env: ~9 Philox blocks, three 3x3 tridiagonal solves and spline evaluations in float64
__global__ void __launch_bounds__(kBlock)
traj_info_kernel(const __grid_constant__ QsParams P, int n, const float* __restrict__ state,
                 const uint32_t* __restrict__ episode, const int32_t* __restrict__ sample_index, float* __restrict__ out9) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    const uint32_t epi = episode ? episode[i] : f2u_(state[26 * (size_t)n + i]);
    const int idx = sample_index ? sample_index[i] : f2i_(state[24 * (size_t)n + i]) - 1;
    float o[9];
    traj_info_eval(P, P.env_id_offset + (uint32_t)i, epi, idx, o);
#pragma unroll
    for (int k = 0; k < 9; ++k) out9[(size_t)i * 9 + k] = o[k];
}
}  // namespace qs

int qs_traj_info(QsHandle h, const float* state, const uint32_t* episode, const int32_t* sample_index, float* out9,
                 void* stream) {
    if (!h || !state || !out9) return fail(QS_EINVAL, "qs_traj_info: null");
    if (h->P.mode != QS_MODE_TRAJ_GYM) return fail(QS_EINVAL, "qs_traj_info: the spline reference exists in QS_MODE_TRAJ_GYM only");
    DeviceGuard guard(h->device);
    qs::traj_info_kernel<<<nblocks(h->n, qs::kBlock), qs::kBlock, 0, (cudaStream_t)stream>>>(h->P, h->n, state, episode,
                                                                                             sample_index, out9);
    return check_launch("traj_info_kernel");
}

int qs_observe(QsHandle h, const float* state, const float* action, float* obs, float* reward, float* done,
               void* stream) {
    if (!h || !state) return fail(QS_EINVAL, "qs_observe: null");
    DeviceGuard guard(h->device);
    if (action && ((uintptr_t)action & 15u) != 0) return fail(QS_EINVAL, "qs_observe: action must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    QS_DISPATCH_MODE(h->P.mode, (qs::observe_kernel<M_><<<nblocks(h->n, qs::kBlock), qs::kBlock, 0, s>>>(
        h->P, h->tables(), h->n, state, (const float4*)action, obs, reward, done)));
    return check_launch("observe_kernel");
}

int qs_physics_step(QsHandle h, float* state, const float* ctrl, void* stream) {
    if (!h || !state || !ctrl) return fail(QS_EINVAL, "qs_physics_step: null");
    DeviceGuard guard(h->device);
    if (((uintptr_t)ctrl & 15u) != 0) return fail(QS_EINVAL, "qs_physics_step: ctrl must be 16-byte aligned");
    qs::physics_kernel<<<nblocks(h->n, qs::kBlock), qs::kBlock, 0, (cudaStream_t)stream>>>(
        h->P, h->n, state, (const float4*)ctrl);
    return check_launch("physics_kernel");
}

int qs_rollout_random(QsHandle h, float* state, int32_t T, uint32_t t0, float* stats, const float* first_state,
                      void* stream) {
    if (!h || !state || T < 0) return fail(QS_EINVAL, "qs_rollout_random: bad argument");
    DeviceGuard guard(h->device);
    if (h->P.auto_reset == QS_RESET_RESTORE_FIRST && !first_state)
        return fail(QS_EINVAL, "qs_rollout_random: auto_reset=restore_first needs first_state");
    cudaStream_t s = (cudaStream_t)stream;
    {
        // plain north-star configuration, even env count: two envs per thread on the packed FP32 pipe (qs_step2.cuh)
        const QsParams& Pq = h->P;
        static const int use_step2 = getenv("QS_STEP2") ? atoi(getenv("QS_STEP2")) : 1;      // A/B knob
        if (use_step2 && Pq.mode == QS_MODE_HOVER_GYM && !Pq.battery && !Pq.rate_wrapper && !Pq.waypoint_mode && !Pq.pre_clip_action &&
            (h->n & 1) == 0 && (((uintptr_t)state | (uintptr_t)stats) & 7u) == 0) {
            qs::rollout_random2_kernel<<<nblocks(h->n / 2, qs::kBlock2), qs::kBlock2, 0, s>>>(h->P, h->n, h->n / 2, state, T, t0, stats);
            return check_launch("rollout_random2_kernel");
        }
    }
    QS_DISPATCH_MODE(h->P.mode, (qs::rollout_random_kernel<M_><<<nblocks(h->n, qs::kBlock), qs::kBlock, 0, s>>>(
        h->P, h->tables(), h->n, state, T, t0, stats, first_state)));
    return check_launch("rollout_random_kernel");
}

int qs_step_host(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                 float* done_host, void* stream) {
    return qs_step_host_ex(h, state, action_host, obs_host, reward_host, done_host, nullptr, stream);
}

namespace qs {
// float 0/1 flags -> bytes (Gymnasium's `terminated` / `truncated` are bool arrays): 4 flags per thread
__global__ void __launch_bounds__(256)
flags_u8_kernel(const float4* __restrict__ done, const float4* __restrict__ trunc, uchar4* __restrict__ done8,
                uchar4* __restrict__ trunc8, int n4) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    const float4 d = done[i];
    done8[i] = make_uchar4(d.x != 0.f, d.y != 0.f, d.z != 0.f, d.w != 0.f);
    if (trunc) {
        const float4 t = trunc[i];
        trunc8[i] = make_uchar4(t.x != 0.f, t.y != 0.f, t.z != 0.f, t.w != 0.f);
    }
}
}  // namespace qs

static int step_host_impl(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                          void* done_host, void* trunc_host, bool flags_u8, void* stream);

int qs_step_host_ex(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                    float* done_host, float* trunc_host, void* stream) {
    return step_host_impl(h, state, action_host, obs_host, reward_host, done_host, trunc_host, false, stream);
}

int qs_step_host_bytes(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                    uint8_t* done_host, uint8_t* trunc_host, void* stream) {
    return step_host_impl(h, state, action_host, obs_host, reward_host, done_host, trunc_host, true, stream);
}

static int step_host_impl(QsHandle h, float* state, const float* action_host, float* obs_host, float* reward_host,
                          void* done_host, void* trunc_host, bool flags_u8, void* stream) {
    if (!h || !state || !action_host || !obs_host || !reward_host || !done_host)
        return fail(QS_EINVAL, "qs_step_host: null");
    if (h->P.auto_reset == QS_RESET_RESTORE_FIRST) return fail(QS_EUNSUPPORTED, "qs_step_host: use qs_step for brax auto-reset");
    DeviceGuard guard(h->device);
    {
        // the schedule below relies on truly asynchronous copies: pageable buffers would silently serialise it
        const void* hb[5] = {action_host, obs_host, reward_host, done_host, trunc_host};
        for (int k = 0; k < (trunc_host ? 5 : 4); ++k) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, hb[k]) != cudaSuccess) { cudaGetLastError(); return fail(QS_EINVAL, "qs_step_host: cannot query a host buffer"); }
            if (at.type != cudaMemoryTypeHost)
                return fail(QS_EINVAL, "qs_step_host: host buffers must be page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)");
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)h->n, D = (size_t)h->P.obs_dim;
    if (!h->scratch) {
        QS_CUDA(cudaMalloc(&h->scratch, n * (4 + D + 3) * sizeof(float) + 2 * ((n + 15) & ~(size_t)15)));   // + byte flags
        for (int k = 0; k < QS_HOST_STREAMS; ++k) QS_CUDA(cudaStreamCreateWithFlags(&h->hs[k], cudaStreamNonBlocking));
        for (int k = 0; k <= QS_HOST_STREAMS; ++k) QS_CUDA(cudaEventCreateWithFlags(&h->hev[k], cudaEventDisableTiming));
    }
    float* d_act = h->scratch; float* d_obs = d_act + 4 * n; float* d_rew = d_obs + D * n; float* d_done = d_rew + n;
    float* d_trunc = trunc_host ? d_done + n : nullptr;
    uint8_t* d_done8 = reinterpret_cast<uint8_t*>(h->scratch + n * (4 + D + 3));
    uint8_t* d_trunc8 = d_done8 + ((n + 15) & ~(size_t)15);
    if (flags_u8 && (n & 3u) != 0) return fail(QS_EUNSUPPORTED, "qs_step_host_bytes: num_envs must be a multiple of 4");
    // Chunked, double-streamed: PCIe is full duplex, so the H2D of chunk k+1 and the kernel of chunk k+1 run under the
    // D2H of chunk k.  The D2H of the observations (48 of the 56 bytes per env) is what bounds the call, so the schedule
    // is built around keeping that copy engine busy: geometrically growing chunks (1/16, 1/16, 1/8, 1/4, 1/2 of the
    // batch) start the first D2H after ~20 us instead of ~90 us and keep the number of copies -- each costs ~4.6 us of
    // set-up on this platform -- small; reward and done leave in one copy each at the end.  Chunk boundaries are
    // multiples of the block size; small batches go through as one chunk.
    static const int chunks_env = getenv("QS_HOST_CHUNKS") ? atoi(getenv("QS_HOST_CHUNKS")) : 0;   // tuning override: equal chunks
    size_t bounds[34];
    int chunks = 0;
    bounds[0] = 0;
    if (chunks_env > 0 || n < ((size_t)1 << 17)) {
        chunks = chunks_env > 0 ? (chunks_env > 32 ? 32 : chunks_env) : 1;
        const size_t per = ((n + chunks - 1) / chunks + qs::kBlock - 1) / qs::kBlock * qs::kBlock;
        for (int c = 1; c <= chunks; ++c) bounds[c] = (size_t)c * per < n ? (size_t)c * per : n;
    } else {
        const size_t unit = (n / 16 + qs::kBlock - 1) / qs::kBlock * qs::kBlock;
        const int mult[5] = {1, 2, 4, 8, 16};
        chunks = 5;
        for (int c = 1; c <= chunks; ++c) bounds[c] = unit * mult[c - 1] < n ? unit * mult[c - 1] : n;
        bounds[chunks] = n;
    }
    QS_CUDA(cudaEventRecord(h->hev[0], s));                       // state is ready when the caller's stream gets here
    for (int k = 0; k < QS_HOST_STREAMS; ++k) QS_CUDA(cudaStreamWaitEvent(h->hs[k], h->hev[0], 0));
    // one stream per chunk (round robin): nothing of chunk k+1 queues behind the D2H of chunk k, the copy engines see
    // every H2D / D2H as soon as its own producer is done
    for (int c = 0; c < chunks; ++c) {
        const size_t lo = bounds[c];
        if (lo >= n || bounds[c + 1] <= lo) continue;
        const size_t cnt = bounds[c + 1] - lo;
        cudaStream_t cs = h->hs[c % QS_HOST_STREAMS];
        QS_CUDA(cudaMemcpyAsync(d_act + 4 * lo, action_host + 4 * lo, 4 * cnt * sizeof(float), cudaMemcpyHostToDevice, cs));
        int rc = launch_step(h, (int)lo, (int)cnt, state, d_act, d_obs, d_rew, d_done, d_trunc, nullptr, nullptr, nullptr, cs);
        if (rc != QS_OK) return rc;
        QS_CUDA(cudaMemcpyAsync(obs_host + D * lo, d_obs + D * lo, D * cnt * sizeof(float), cudaMemcpyDeviceToHost, cs));
        if (flags_u8) {
            // chunk boundaries are multiples of the block size (128) except the last one, and n % 4 == 0
            const int n4 = (int)(cnt / 4);
            qs::flags_u8_kernel<<<nblocks(n4, 256), 256, 0, cs>>>(
                reinterpret_cast<const float4*>(d_done + lo), d_trunc ? reinterpret_cast<const float4*>(d_trunc + lo) : nullptr,
                reinterpret_cast<uchar4*>(d_done8 + lo), reinterpret_cast<uchar4*>(d_trunc8 + lo), n4);
            int rc2 = check_launch("flags_u8_kernel");
            if (rc2 != QS_OK) return rc2;
        }
    }
    // reward / done: one copy each on stream 0, after every stream's last kernel (d_rew and d_done are adjacent)
    for (int k = 1; k < QS_HOST_STREAMS; ++k) {
        QS_CUDA(cudaEventRecord(h->hev[1 + k], h->hs[k]));
        QS_CUDA(cudaStreamWaitEvent(h->hs[0], h->hev[1 + k], 0));
    }
    QS_CUDA(cudaMemcpyAsync(reward_host, d_rew, n * sizeof(float), cudaMemcpyDeviceToHost, h->hs[0]));
    if (flags_u8) {
        QS_CUDA(cudaMemcpyAsync(done_host, d_done8, n, cudaMemcpyDeviceToHost, h->hs[0]));
        if (trunc_host) QS_CUDA(cudaMemcpyAsync(trunc_host, d_trunc8, n, cudaMemcpyDeviceToHost, h->hs[0]));
    } else {
        QS_CUDA(cudaMemcpyAsync(done_host, d_done, n * sizeof(float), cudaMemcpyDeviceToHost, h->hs[0]));
        if (trunc_host) QS_CUDA(cudaMemcpyAsync(trunc_host, d_trunc, n * sizeof(float), cudaMemcpyDeviceToHost, h->hs[0]));
    }
    QS_CUDA(cudaEventRecord(h->hev[1], h->hs[0]));
    QS_CUDA(cudaStreamWaitEvent(s, h->hev[1], 0));                // later work on the caller's stream sees the new state
    QS_CUDA(cudaStreamSynchronize(s));
    return QS_OK;
}

// ---- policy rollout + GAE: implemented in qs_rollout.cuh -------------------------------
int qs_policy_param_count(const QsPolicyDesc* d) {
    if (!d) return QS_EINVAL;
    return qs::policy_param_count(*d);
}

int qs_rollout_policy(QsHandle h, float* state, const QsPolicyDesc* desc, const float* policy_params,
                      int32_t T, uint32_t t0, float* last_obs, float* traj_obs, float* traj_act, float* traj_logp,
                      float* traj_value, float* traj_reward, float* traj_done, float* traj_trunc,
                      float* last_value, const float* first_state, void* stream) {
    if (!h || !state || !desc || !policy_params || T <= 0) return fail(QS_EINVAL, "qs_rollout_policy: bad argument");
    DeviceGuard guard(h->device);
    if (desc->obs_dim != h->P.obs_dim || desc->hidden != 128 || desc->act_dim != 4 || desc->dist < 0 || desc->dist > 1)
        return fail(QS_EUNSUPPORTED, "qs_rollout_policy: policy must be obs_dim->128->128->4 (dist 0|1)");
    if (h->P.auto_reset == QS_RESET_RESTORE_FIRST && !first_state)
        return fail(QS_EINVAL, "qs_rollout_policy: auto_reset=restore_first needs first_state");
    qs::RolloutBuffers rb{last_obs, traj_obs, traj_act, traj_logp, traj_value, traj_reward, traj_done, traj_trunc, last_value};
    int rc;
    if (desc->tensor_cores) {
        rc = qs::tc::launch_rollout_policy_tc(h->P, h->tables(), h->n, state, *desc, policy_params, T, t0, rb, first_state,
                                              (cudaStream_t)stream);
    } else {
        rc = qs::launch_rollout_policy(h->P, h->tables(), h->n, state, *desc, policy_params, T, t0, rb, first_state,
                                       (cudaStream_t)stream);
    }
    if (rc == -100) return fail(QS_EUNSUPPORTED, "qs_rollout_policy: unsupported mode/dist combination");
    if (rc != 0) return fail(QS_ECUDA, "qs_rollout_policy: launch configuration failed", (cudaError_t)rc);
    return check_launch("rollout_policy_kernel");
}

int qs_gae(int32_t T, int32_t B, const float* reward, const float* value, const float* done, const float* trunc,
           const float* last_value, float gamma, float lam, int32_t brax_form, float* adv, float* ret, void* stream) {
    if (T <= 0 || B <= 0 || !reward || !value || !done || !last_value || !adv || !ret)
        return fail(QS_EINVAL, "qs_gae: bad argument");
    if (brax_form && !trunc) return fail(QS_EINVAL, "qs_gae: brax form needs truncation flags");
    qs::gae_kernel<<<nblocks(B, qs::kGaeBlock), qs::kGaeBlock, 0, (cudaStream_t)stream>>>(T, B, reward, value, done, trunc, last_value,
                                                                      gamma, lam, brax_form, adv, ret);
    return check_launch("gae_kernel");
}

// ---- PPO update on device: implemented in qs_ppo.cuh ------------------------------------------------
namespace {
constexpr size_t kPpoWsHeader = 128;   // bytes: 2 doubles + 1 counter (advantage sums) | 2 floats (mean, 1/std) at byte 32
// tail of the workspace (behind the partial-gradient rows): per-minibatch records for the one-launch-per-epoch advantage
// statistics of qs_ppo_update_epoch: kPpoMaxEpochMb x {2 doubles + counter (32 B)} | kPpoMaxEpochMb x {mean, 1/std}
constexpr int kPpoMaxEpochMb = 1024;
constexpr size_t kPpoWsTail = (size_t)kPpoMaxEpochMb * (32 + 8);

int ppo_device_sms(int* sms) {
    int dev = 0, major = 0;
    QS_CUDA(cudaGetDevice(&dev));
    QS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return fail(QS_EUNSUPPORTED, "qs_ppo_*: libquadsim is built for sm_100a (B200) only");
    QS_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    return QS_OK;
}
bool ppo_desc_ok(const QsPolicyDesc* d) {
    return d && (d->obs_dim == 12 || d->obs_dim == 21) && d->hidden == 128 && d->act_dim == 4 && (d->dist == 0 || d->dist == 1);
}
const char* kPpoDescMsg = "the fused update covers obs_dim 12 | 21 -> 128 -> 128 -> 4 with dist 0 (SB3 Gaussian) | 1 (Brax tanh-normal)";

extern "C++" {
template <int D, int DIST>
int launch_ppo_grad_generic(const qs::ppo::Batch& b, const qs::ppo::HyperG& hp, const float* params, const float* adv_norm,
                            float* partial, int grid, int mn_swap, cudaStream_t s) {
    using S = qs::ppo::SmemPT<(D + 2 <= 16) ? 16 : 32>;
    auto kern = qs::ppo::ppo_grad_tc_kernel<D, DIST>;
    QS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    kern<<<grid, qs::tc::kM, S::TOTAL, s>>>(b, hp, params, adv_norm, partial, mn_swap);
    return QS_OK;
}
}  // extern "C++"
}  // namespace

int64_t qs_ppo_workspace_bytes(const QsPolicyDesc* desc) {
    if (!ppo_desc_ok(desc)) return fail(QS_EUNSUPPORTED, kPpoDescMsg);
    int sms = 0;
    const int rc = ppo_device_sms(&sms);
    if (rc != QS_OK) return rc;
    const size_t row = (size_t)qs::ppo::partial_stride(qs::policy_param_count(*desc));
    return (int64_t)(kPpoWsHeader + (size_t)sms * row * sizeof(float) + kPpoWsTail);
}

static int ppo_grad_impl(const QsPolicyDesc* desc, const float* policy_params, const qs::ppo::Batch& b, float clip_range,
                         float vf_coef, float ent_coef, int32_t normalize_adv, void* workspace, float* grad, void* stream,
                         const float* adv_norm_ready = nullptr);

int qs_ppo_grad(const QsPolicyDesc* desc, const float* policy_params, const float* obs, const float* act,
                const float* old_logp, const float* adv, const float* ret, const int32_t* idx, int32_t n,
                float clip_range, float vf_coef, float ent_coef, int32_t normalize_adv, void* workspace, float* grad,
                void* stream) {
    if (!ppo_desc_ok(desc)) return fail(QS_EUNSUPPORTED, kPpoDescMsg);
    if (!policy_params || !obs || !act || !old_logp || !adv || !ret || !workspace || !grad || n <= 0)
        return fail(QS_EINVAL, "qs_ppo_grad: bad argument");
    if ((((uintptr_t)act | (uintptr_t)workspace | (uintptr_t)grad) & 15u) != 0 || (desc->obs_dim % 4 == 0 && ((uintptr_t)obs & 15u) != 0))
        return fail(QS_EINVAL, "qs_ppo_grad: obs (12-D), act, workspace and grad must be 16-byte aligned");
    qs::ppo::Batch b{obs, act, old_logp, adv, ret, idx, n, nullptr};
    return ppo_grad_impl(desc, policy_params, b, clip_range, vf_coef, ent_coef, normalize_adv, workspace, grad, stream);
}

int qs_ppo_pack(const QsPolicyDesc* desc, const float* obs, const float* act, const float* old_logp, const float* adv,
                const float* ret, int64_t n, float* packed, void* stream) {
    if (!ppo_desc_ok(desc)) return fail(QS_EUNSUPPORTED, kPpoDescMsg);
    if (!obs || !act || !old_logp || !adv || !ret || !packed || n <= 0) return fail(QS_EINVAL, "qs_ppo_pack: bad argument");
    if (((uintptr_t)packed & 127u) != 0) return fail(QS_EINVAL, "qs_ppo_pack: packed rows must be 128-byte aligned");
    if ((((uintptr_t)obs | (uintptr_t)act) & 15u) != 0) return fail(QS_EINVAL, "qs_ppo_pack: obs and act must be 16-byte aligned");
    const unsigned blocks = (unsigned)((n + qs::ppo::kPackRows - 1) / qs::ppo::kPackRows);
    if (desc->obs_dim == 12)
        qs::ppo::ppo_pack_kernel<12><<<blocks, qs::ppo::kPackRows, 0, (cudaStream_t)stream>>>(obs, act, old_logp, adv, ret, (long long)n, packed);
    else
        qs::ppo::ppo_pack_kernel<21><<<blocks, qs::ppo::kPackRows, 0, (cudaStream_t)stream>>>(obs, act, old_logp, adv, ret, (long long)n, packed);
    return check_launch("ppo_pack_kernel");
}

int qs_ppo_grad_packed(const QsPolicyDesc* desc, const float* policy_params, const float* packed, const float* adv,
                       const int32_t* idx, int32_t n, float clip_range, float vf_coef, float ent_coef, int32_t normalize_adv,
                       void* workspace, float* grad, void* stream) {
    if (!ppo_desc_ok(desc)) return fail(QS_EUNSUPPORTED, kPpoDescMsg);
    if (!policy_params || !packed || !workspace || !grad || n <= 0 || (normalize_adv && !adv))
        return fail(QS_EINVAL, "qs_ppo_grad_packed: bad argument (the advantage statistics read the contiguous adv array)");
    if (((uintptr_t)packed & 127u) != 0 || (((uintptr_t)workspace | (uintptr_t)grad) & 15u) != 0)
        return fail(QS_EINVAL, "qs_ppo_grad_packed: packed rows must be 128-byte, workspace and grad 16-byte aligned");
    qs::ppo::Batch b{nullptr, nullptr, nullptr, adv, nullptr, idx, n, packed};
    return ppo_grad_impl(desc, policy_params, b, clip_range, vf_coef, ent_coef, normalize_adv, workspace, grad, stream);
}

static int ppo_grad_impl(const QsPolicyDesc* desc, const float* policy_params, const qs::ppo::Batch& b, float clip_range,
                         float vf_coef, float ent_coef, int32_t normalize_adv, void* workspace, float* grad, void* stream,
                         const float* adv_norm_ready) {
    const float* adv = b.adv;
    const int32_t* idx = b.idx;
    const int32_t n = b.n;
    if (normalize_adv < 0 || normalize_adv > 2) return fail(QS_EINVAL, "qs_ppo_grad: normalize_adv is 0 (off), 1 (unbiased std) or 2 (population std)");
    int sms = 0;
    const int rc = ppo_device_sms(&sms);
    if (rc != QS_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    double* adv_scratch = (double*)ws;
    const float* adv_norm = adv_norm_ready ? adv_norm_ready : (const float*)(ws + 32);
    float* partial = (float*)(ws + kPpoWsHeader);
    if (normalize_adv && !adv_norm_ready) {
        const int blocks = nblocks(n, 256) < 4 * sms ? nblocks(n, 256) : 4 * sms;
        qs::ppo::ppo_adv_stats_kernel<<<blocks, 256, 0, s>>>(adv, idx, n, normalize_adv == 2 ? 0 : 1, adv_scratch, (float*)(ws + 32), 0, n);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    const int ntiles = nblocks(n, qs::tc::kM);
    static const int mn_swap = getenv("QS_PPO_MN_SWAP") ? atoi(getenv("QS_PPO_MN_SWAP")) : 0;   // descriptor debug knob
    static const int use_v1 = getenv("QS_PPO_V1") ? atoi(getenv("QS_PPO_V1")) : 0;              // A/B: the single-tile schedule
    const int P = qs::policy_param_count(*desc);
    const qs::PolicyLayout L = qs::policy_layout(desc->obs_dim, desc->dist);
    const int len = P + qs::ppo::kPartialStats;
    int rows_a, rows_c;
    if (use_v1 || desc->obs_dim != 12 || desc->dist != 0) {
        // single-tile schedule, templated on observation size / distribution (qs_ppo_generic.cuh)
        const int grid = ntiles < sms ? ntiles : sms;
        qs::ppo::HyperG hg{clip_range, vf_coef, ent_coef, normalize_adv, (uint32_t)desc->sample_seed};
        int rc2;
        if (desc->obs_dim == 12) rc2 = desc->dist == 0 ? launch_ppo_grad_generic<12, 0>(b, hg, policy_params, adv_norm, partial, grid, mn_swap, s)
                                                       : launch_ppo_grad_generic<12, 1>(b, hg, policy_params, adv_norm, partial, grid, mn_swap, s);
        else rc2 = desc->dist == 0 ? launch_ppo_grad_generic<21, 0>(b, hg, policy_params, adv_norm, partial, grid, mn_swap, s)
                                   : launch_ppo_grad_generic<21, 1>(b, hg, policy_params, adv_norm, partial, grid, mn_swap, s);
        if (rc2 != QS_OK) return rc2;
        rows_a = rows_c = grid;
    } else {
        qs::ppo::Hyper hp{clip_range, vf_coef, ent_coef, normalize_adv};
        // one network per CTA, two tiles in flight, at most one CTA per SM; the actor's tiles cost ~13 % more than the
        // critic's (loss math, more gathered columns), so it gets ~53 % of the SMs
        const int pairs_needed = (ntiles + 1) / 2;
        const int a_max = (sms * 53 + 50) / 100 > sms - 1 ? sms - 1 : (sms * 53 + 50) / 100, c_max = sms - a_max;
        rows_a = pairs_needed < a_max ? pairs_needed : a_max;
        rows_c = pairs_needed < c_max ? pairs_needed : c_max;
        QS_CUDA(cudaFuncSetAttribute(qs::ppo::ppo_grad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, qs::ppo::SmemQ::TOTAL));
        qs::ppo::ppo_grad_tc2_kernel<<<rows_a + rows_c, qs::ppo::kThreads2, qs::ppo::SmemQ::TOTAL, s>>>(b, hp, policy_params, adv_norm, partial, rows_a);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    // entries [cW1, end of the critic) and statistic 1 (value loss) come from the critic rows, everything else from the actor rows
    qs::ppo::ppo_reduce_kernel<<<nblocks(len, 256), 256, 0, s>>>(partial, rows_a, rows_c, L.cW1, L.log_std, P + 1,
                                                                 qs::ppo::partial_stride(P), len, grad);
    return check_launch("ppo_grad");
}

// Running observation normaliser (brax normalize_observations=True, train_brax_ppo.py:611): merges the batch obs [n][obs_dim]
// into running = {count, mean[obs_dim], M2[obs_dim]} (device, doubles, zero-initialised by the caller) and refreshes the
// policy's obs_mean / obs_inv_std entries (mean_out / inv_std_out: pointers INTO the packed parameter vector).
int qs_obs_stats_update(const float* obs, int64_t n, int32_t obs_dim, double* running, float* mean_out, float* inv_std_out,
                        float std_min, float std_max, void* workspace, void* stream) {
    if (!obs || !running || !mean_out || !inv_std_out || !workspace || n <= 0) return fail(QS_EINVAL, "qs_obs_stats_update: bad argument");
    if (obs_dim != 12 && obs_dim != 21) return fail(QS_EUNSUPPORTED, "qs_obs_stats_update: obs_dim 12 | 21");
    cudaStream_t s = (cudaStream_t)stream;
    double* part = (double*)workspace;
    const int blocks = (int)((n + 255) / 256 < qs::ppo::kObsStatBlocks ? (n + 255) / 256 : qs::ppo::kObsStatBlocks);
    if (obs_dim == 12) {
        qs::ppo::obs_stats_partial_kernel<12><<<blocks, 256, 0, s>>>(obs, (long long)n, part);
        qs::ppo::obs_stats_merge_kernel<12><<<1, 32, 0, s>>>(obs, (long long)n, part, blocks, running, mean_out, inv_std_out, std_min, std_max);
    } else {
        qs::ppo::obs_stats_partial_kernel<21><<<blocks, 256, 0, s>>>(obs, (long long)n, part);
        qs::ppo::obs_stats_merge_kernel<21><<<1, 32, 0, s>>>(obs, (long long)n, part, blocks, running, mean_out, inv_std_out, std_min, std_max);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_launch("obs_stats");
}
int64_t qs_obs_stats_workspace_bytes(int32_t obs_dim) { return (int64_t)qs::ppo::kObsStatBlocks * 2 * (obs_dim > 0 ? obs_dim : 1) * sizeof(double); }

int qs_ppo_permutation(int32_t n, uint64_t seed, uint32_t epoch, int32_t* out, void* stream) {
    if (n <= 0 || !out) return fail(QS_EINVAL, "qs_ppo_permutation: bad argument");
    int bits = 1;
    while (bits < 31 && (1u << bits) < (uint32_t)n) ++bits;
    const int half_bits = (bits + 1) / 2 < 1 ? 1 : (bits + 1) / 2;
    const uint32_t k0 = qs::ppo::perm_mix((uint32_t)seed ^ 0xA511E9B3u) + epoch * 0x9E3779B9u;
    const uint32_t k1 = qs::ppo::perm_mix((uint32_t)(seed >> 32) + 0x632BE5ABu) ^ qs::ppo::perm_mix(epoch + 0x85157AF5u);
    qs::ppo::PermKeys K;
    K.k0 = k0;
    for (uint32_t r = 0; r < 4; ++r) K.rk[r] = qs::ppo::perm_mix(k1 + r);
    // one warp per chunk of kPermChunk elements, 8 warps per block
    qs::ppo::ppo_permutation_kernel<<<nblocks(nblocks(n, qs::ppo::kPermChunk), 8), 256, 0, (cudaStream_t)stream>>>((uint32_t)n, half_bits, K, out);
    return check_launch("ppo_permutation_kernel");
}

static int launch_adam(const QsPolicyDesc* desc, float* policy_params, const float* grad, float* m, float* v, int32_t step,
                       float lr, float beta1, float beta2, float eps, float max_grad_norm, float grad_scale, float* norm_out,
                       float* stats_acc, void* stream) {
    const qs::PolicyLayout L = qs::policy_layout(desc->obs_dim, desc->dist);
    qs::ppo::AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_grad_norm = max_grad_norm; a.grad_scale = grad_scale;
    a.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias2 = (float)(1.0 - pow((double)beta2, (double)step));
    a.n_train = L.mean;
    if (a.n_train > qs::ppo::kAdamPerThread * 1024) return fail(QS_EUNSUPPORTED, "qs_ppo_adam: policy too large for the single-CTA optimiser step");
    qs::ppo::ppo_adam_kernel<<<nblocks(a.n_train, 1024), 1024, 0, (cudaStream_t)stream>>>(
        a, policy_params, grad, m, v, norm_out, grad + qs::policy_param_count(*desc), stats_acc);
    return check_launch("ppo_adam_kernel");
}

int qs_ppo_adam(const QsPolicyDesc* desc, float* policy_params, const float* grad, float* m, float* v, int32_t step,
                float lr, float beta1, float beta2, float eps, float max_grad_norm, float grad_scale, float* norm_out,
                void* stream) {
    if (!desc || desc->hidden != 128 || desc->act_dim != 4) return fail(QS_EINVAL, "qs_ppo_adam: bad policy description");
    if (!policy_params || !grad || !m || !v || step <= 0) return fail(QS_EINVAL, "qs_ppo_adam: bad argument");
    return launch_adam(desc, policy_params, grad, m, v, step, lr, beta1, beta2, eps, max_grad_norm, grad_scale, norm_out, nullptr, stream);
}

// ---- multi-GPU PPO update over NVLink peer memory (qs_ppo.cuh: ppo_peer_adam_kernel) ------------------------------------
struct QsPpoComm {
    int world, rank, P, device;
    qs::ppo::PeerLayout L;
    unsigned char* local;
    qs::ppo::PeerPtrs peers;
    bool opened[qs::ppo::kMaxPeers];
};

int qs_ppo_comm_create(const QsPolicyDesc* desc, int32_t world, int32_t rank, QsPpoComm** out) {
    if (!ppo_desc_ok(desc) || !out || world < 1 || world > qs::ppo::kMaxPeers || rank < 0 || rank >= world)
        return fail(QS_EINVAL, "qs_ppo_comm_create: bad argument (1 <= world <= 8, supported policy)");
    int sms = 0;
    const int rc = ppo_device_sms(&sms);
    if (rc != QS_OK) return rc;
    QsPpoComm* c = new (std::nothrow) QsPpoComm();
    if (!c) return fail(QS_ENOMEM, "qs_ppo_comm_create: host allocation failed");
    c->world = world; c->rank = rank; c->P = qs::policy_param_count(*desc);
    QS_CUDA(cudaGetDevice(&c->device));
    c->L.F = qs::ppo::partial_stride(c->P);
    for (int p = 0; p < qs::ppo::kMaxPeers; ++p) { c->peers.base[p] = nullptr; c->opened[p] = false; }
    {
        // the optimiser kernel's grid barrier spins on a counter: all of its CTAs must be co-resident
        int per_sm = 0;
        const int grid = nblocks(c->L.F, 1024);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qs::ppo::ppo_peer_adam_kernel, 1024, 0) != cudaSuccess ||
            per_sm * sms < grid) { delete c; return fail(QS_EUNSUPPORTED, "qs_ppo_comm_create: the optimiser grid cannot be co-resident on this device"); }
    }
    if (cudaMalloc(&c->local, c->L.total()) != cudaSuccess) { delete c; return fail(QS_ENOMEM, "qs_ppo_comm_create: cudaMalloc"); }
    QS_CUDA(cudaMemset(c->local, 0, c->L.total()));
    QS_CUDA(cudaDeviceSynchronize());
    c->peers.base[rank] = c->local;
    *out = c;
    return QS_OK;
}

int qs_ppo_comm_export(QsPpoComm* c, void* handle64) {
    if (!c || !handle64) return fail(QS_EINVAL, "qs_ppo_comm_export: null");
    DeviceGuard guard(c->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    QS_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle64, &h, sizeof(h));
    return QS_OK;
}

int qs_ppo_comm_import(QsPpoComm* c, int32_t peer, const void* handle64) {
    if (!c || !handle64 || peer < 0 || peer >= c->world) return fail(QS_EINVAL, "qs_ppo_comm_import: bad argument");
    if (peer == c->rank) return QS_OK;
    DeviceGuard guard(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* ptr = nullptr;
    QS_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peers.base[peer] = (unsigned char*)ptr;
    c->opened[peer] = true;
    return QS_OK;
}

void* qs_ppo_comm_slot(QsPpoComm* c, uint32_t epoch) {
    if (!c) return nullptr;
    return c->local + c->L.slot((int)(epoch & 1u));
}

int qs_ppo_adam_peer(const QsPolicyDesc* desc, QsPpoComm* c, uint32_t epoch, float* policy_params, float* m, float* v,
                     int32_t step, float lr, float beta1, float beta2, float eps, float max_grad_norm, float* norm_out,
                     float* stats_acc, void* stream) {
    if (!ppo_desc_ok(desc) || !c || !policy_params || !m || !v || step <= 0 || epoch == 0)
        return fail(QS_EINVAL, "qs_ppo_adam_peer: bad argument (epoch counts from 1)");
    for (int p = 0; p < c->world; ++p)
        if (!c->peers.base[p]) return fail(QS_EINVAL, "qs_ppo_adam_peer: a peer buffer has not been imported");
    DeviceGuard guard(c->device);
    const qs::PolicyLayout L = qs::policy_layout(desc->obs_dim, desc->dist);
    qs::ppo::AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_grad_norm = max_grad_norm; a.grad_scale = 1.0f / (float)c->world;
    a.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias2 = (float)(1.0 - pow((double)beta2, (double)step));
    a.n_train = L.mean;
    const int grid = nblocks(c->L.F, 1024);
    if (grid > 64) return fail(QS_EUNSUPPORTED, "qs_ppo_adam_peer: policy too large");
    static const unsigned long long timeout_ns =
        (unsigned long long)(getenv("QS_PEER_TIMEOUT_MS") ? atof(getenv("QS_PEER_TIMEOUT_MS")) : 20000.0) * 1000000ull;
    qs::ppo::ppo_peer_adam_kernel<<<grid, 1024, 0, (cudaStream_t)stream>>>(a, c->L, c->peers, c->world, c->rank, epoch, c->P,
                                                                           policy_params, m, v, norm_out, stats_acc, timeout_ns);
    return check_launch("ppo_peer_adam_kernel");
}

// One epoch of minibatch updates launched back to back from native code: for the small minibatches of the reference's own
// geometry (train_brax_ppo.py:589-620: 10 240 samples) the four launches of an update cost less GPU time than their
// Python / ctypes round trips, and the statistics accumulate inside the optimiser kernel instead of a separate torch add.
int qs_ppo_update_epoch(const QsPolicyDesc* desc, float* policy_params, const float* packed, const float* adv,
                        const int32_t* perm, int32_t n_total, int32_t num_minibatches, float clip_range, float vf_coef,
                        float ent_coef, int32_t normalize_adv, float* m, float* v, int32_t step0, float lr, float beta1,
                        float beta2, float eps, float max_grad_norm, QsPpoComm* comm, uint32_t epoch0, void* workspace,
                        float* grad, float* stats_acc, float* norm_out, uint64_t sample_seed0, void* stream) {
    if (!ppo_desc_ok(desc)) return fail(QS_EUNSUPPORTED, kPpoDescMsg);
    if (!policy_params || !packed || !perm || !m || !v || !workspace || n_total <= 0 || num_minibatches <= 0 ||
        num_minibatches > n_total || step0 <= 0 || (normalize_adv && !adv) || (!comm && !grad) || (comm && epoch0 == 0))
        return fail(QS_EINVAL, "qs_ppo_update_epoch: bad argument");
    if (((uintptr_t)packed & 127u) != 0 || (((uintptr_t)workspace | (uintptr_t)grad) & 15u) != 0)
        return fail(QS_EINVAL, "qs_ppo_update_epoch: packed rows must be 128-byte, workspace and grad 16-byte aligned");
    const int32_t mb = n_total / num_minibatches;
    QsPolicyDesc d = *desc;
    // the advantage statistics of ALL minibatches of the epoch in one launch (they depend on the shuffle only, not on the
    // parameters): one record per minibatch in the tail of the workspace
    const float* norm_all = nullptr;
    if (normalize_adv && num_minibatches <= kPpoMaxEpochMb) {
        int sms = 0;
        const int rc0 = ppo_device_sms(&sms);
        if (rc0 != QS_OK) return rc0;
        const size_t row = (size_t)qs::ppo::partial_stride(qs::policy_param_count(*desc));
        unsigned char* tail = (unsigned char*)workspace + kPpoWsHeader + (size_t)sms * row * sizeof(float);
        float* norms = (float*)(tail + (size_t)kPpoMaxEpochMb * 32);
        const int blocks = nblocks(mb, 256) < 4 * sms ? nblocks(mb, 256) : 4 * sms;
        qs::ppo::ppo_adv_stats_kernel<<<dim3((unsigned)blocks, (unsigned)num_minibatches), 256, 0, (cudaStream_t)stream>>>(
            adv, perm, mb, normalize_adv == 2 ? 0 : 1, (double*)tail, norms, mb, n_total);
        const int rc1 = check_launch("ppo_adv_stats_kernel (epoch)");
        if (rc1 != QS_OK) return rc1;
        norm_all = norms;
    }
    for (int32_t k = 0; k < num_minibatches; ++k) {
        // (the last minibatch takes the n_total % num_minibatches remainder rows, as SB3's RolloutBuffer.get does)
        const int32_t n = k + 1 < num_minibatches ? mb : n_total - k * mb;
        d.sample_seed = (int32_t)((sample_seed0 + (uint64_t)k) & 0x7FFFFFFFull);
        float* out = comm ? (float*)qs_ppo_comm_slot(comm, epoch0 + (uint32_t)k) : grad;
        qs::ppo::Batch b{nullptr, nullptr, nullptr, adv, nullptr, perm + (size_t)k * mb, n, packed};
        int rc = ppo_grad_impl(&d, policy_params, b, clip_range, vf_coef, ent_coef, normalize_adv, workspace, out, stream,
                               norm_all ? norm_all + 2 * k : nullptr);
        if (rc != QS_OK) return rc;
        rc = comm ? qs_ppo_adam_peer(&d, comm, epoch0 + (uint32_t)k, policy_params, m, v, step0 + k, lr, beta1, beta2, eps,
                                     max_grad_norm, norm_out, stats_acc, stream)
                  : launch_adam(&d, policy_params, grad, m, v, step0 + k, lr, beta1, beta2, eps, max_grad_norm, 1.0f, norm_out,
                                stats_acc, stream);
        if (rc != QS_OK) return rc;
    }
    return QS_OK;
}

int qs_ppo_comm_error(QsPpoComm* c) {
    if (!c) return fail(QS_EINVAL, "qs_ppo_comm_error: null");
    DeviceGuard guard(c->device);
    uint32_t w = 0;
    QS_CUDA(cudaMemcpy(&w, c->local + c->L.counters() + 4 * sizeof(uint32_t), sizeof(w), cudaMemcpyDeviceToHost));
    if (w == 0) return QS_OK;
    char msg[160];
    if (w >= 0x100u) snprintf(msg, sizeof(msg), "qs_ppo_adam_peer: the optimiser grid did not complete (a CTA bailed out or was never scheduled)");
    else snprintf(msg, sizeof(msg), "qs_ppo_adam_peer: rank %u did not post its gradient within the timeout (QS_PEER_TIMEOUT_MS); parameters untouched", w - 1u);
    return fail(QS_ECUDA, msg);
}

int qs_ppo_comm_close_peers(QsPpoComm* c) {
    if (!c) return QS_OK;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; ++p)
        if (c->opened[p]) { cudaIpcCloseMemHandle(c->peers.base[p]); c->opened[p] = false; c->peers.base[p] = nullptr; }
    return QS_OK;
}

int qs_ppo_comm_destroy(QsPpoComm* c) {
    if (!c) return QS_OK;
    qs_ppo_comm_close_peers(c);
    DeviceGuard guard(c->device);
    cudaFree(c->local);
    delete c;
    return QS_OK;
}

}  // extern "C"
