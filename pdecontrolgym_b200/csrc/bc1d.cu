// bc1d.cu -- TransportPDE1D and ReactionDiffusionPDE1D: reset kernel, validation, dispatch and C ABI (sm_100a).
//
// Reference behaviour restated here (no code shared): hyperbolic.py:126-227, parabolic.py:126-221,
// base_env_1d.py:20-45, tuned_reward_1d.py:25-40, norm_reward.py:29-73 of lukebhan/PDEControlGym.
// The step kernel templates live in bc1d_kernels.cuh.
#include "bc1d_kernels.cuh"

namespace pdecg {

// ---- reset -------------------------------------------------------------------------------------
template <typename S, typename I>
__global__ void __launch_bounds__(kThreads) bc1d_reset_kernel(const Bc1dArgs a, const I* __restrict__ init,
                                                             int64_t init_stride, const uint8_t* __restrict__ mask,
                                                             const RandomIc rnd) {
    const int lane = threadIdx.x & 31;
    const int64_t env = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (env >= a.buf.n_envs) return;
    if (mask != nullptr && mask[env] == 0) return;
    reset_env<S, I, 32>(a, init, init_stride, rnd, env, lane, kFull, false);
}

// ---- host-side dispatch --------------------------------------------------------------------------
static thread_local int64_t g_map_envs = -1;      // set by pdecg_bc1d_step_host while it launches the chunks of one batch
struct MapEnvsScope {
    explicit MapEnvsScope(int64_t n) { g_map_envs = n; }
    ~MapEnvsScope() { g_map_envs = -1; }
};

static int validate(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, Bc1dArgs* out) {
    if (!cfg || !buf) return fail(PDECG_ERR_INVALID, "null cfg/buffers");
    if (cfg->abi_version != PDECG_ABI_VERSION) return fail(PDECG_ERR_INVALID, "abi_version mismatch (got %d, want %d)", cfg->abi_version, PDECG_ABI_VERSION);
    if (cfg->env != PDECG_ENV_TRANSPORT && cfg->env != PDECG_ENV_REACTION_DIFFUSION) return fail(PDECG_ERR_INVALID, "env must be 0 (Transport) or 1 (ReactionDiffusion)");
    if (cfg->precision < PDECG_PREC_REF || cfg->precision > PDECG_PREC_FP32) return fail(PDECG_ERR_INVALID, "bad precision %d", cfg->precision);
    if (cfg->nx < 2 || cfg->nt < 2 || cfg->substeps < 0) return fail(PDECG_ERR_INVALID, "need nx >= 2, nt >= 2, substeps >= 0 (nx=%d nt=%d substeps=%d)", cfg->nx, cfg->nt, cfg->substeps);
    if (!(cfg->dt > 0) || !(cfg->dx > 0)) return fail(PDECG_ERR_INVALID, "dt and dx must be positive");
    if (cfg->control_type != PDECG_CONTROL_DIRICHLET && cfg->control_type != PDECG_CONTROL_NEUMANN)
        return fail(PDECG_ERR_INVALID, "Invalid control_type parameter. Please use 'Neumann' or 'Dirchilet'. See documentation for details.");
    if (cfg->sensing_loc < PDECG_SENSE_FULL || cfg->sensing_loc > PDECG_SENSE_OPPOSITE)
        return fail(PDECG_ERR_INVALID, "Invalid sensing_loc parameter. Please use 'full', 'collocated', or 'opposite'. See documentation for details.");
    if (cfg->sensing_loc == PDECG_SENSE_OPPOSITE) {
        if (cfg->sensing_type != PDECG_SENSETYPE_DIRICHLET && cfg->sensing_type != PDECG_SENSETYPE_NEUMANN)
            return fail(PDECG_ERR_INVALID, "Invalid sensing_type parameter. Please use 'Neumann' or 'Dirchilet'. See documentation for details.");
        if (cfg->env == PDECG_ENV_REACTION_DIFFUSION && cfg->sensing_type == PDECG_SENSETYPE_DIRICHLET)
            return fail(PDECG_ERR_INVALID, "In the parabolic PDE system, u(0, t)=0 and so Dirchilet sensing at u(0, t) is not viable. See documentation for details.");
    }
    if (cfg->action_dtype != PDECG_F32 && cfg->action_dtype != PDECG_F64) return fail(PDECG_ERR_INVALID, "bad action_dtype");
    if (cfg->precision == PDECG_PREC_REF && cfg->beta_dtype != PDECG_F32 && cfg->beta_dtype != PDECG_F64) return fail(PDECG_ERR_INVALID, "bad beta_dtype");
    if (cfg->reward_kind == PDECG_REWARD_TUNED1D) {
        // the look-back row t-100 wraps onto an already written row when nt <= 100: not representable without history
        if (cfg->nt <= kLookback) return fail(PDECG_ERR_UNSUPPORTED, "TunedReward1D: nt must exceed 100");
    } else if (cfg->reward_kind == PDECG_REWARD_NORM) {
        if (cfg->norm_ord < PDECG_NORM_L1 || cfg->norm_ord > PDECG_NORM_LINF) return fail(PDECG_ERR_INVALID, "NormReward: norm must be 1, 2 or inf");
        if (cfg->horizon < PDECG_HORIZON_TEMPORAL || cfg->horizon > PDECG_HORIZON_T) return fail(PDECG_ERR_INVALID, "NormReward: horizon must be temporal, differential or t-horizon");
        if (cfg->horizon == PDECG_HORIZON_T && (cfg->t_horizon_length < 1 || cfg->t_horizon_length > cfg->substeps))
            return fail(PDECG_ERR_UNSUPPORTED, "NormReward t-horizon: need 1 <= t_horizon_length <= substeps");
        if (cfg->norm_coeff == 0.0) return fail(PDECG_ERR_INVALID, "NormReward: norm_coeff must be non-zero");
    } else {
        return fail(PDECG_ERR_INVALID, "bad reward_kind %d", cfg->reward_kind);
    }
    if (buf->n_envs < 0) return fail(PDECG_ERR_INVALID, "n_envs < 0");
    if (buf->n_envs > 0 && (!buf->u || !buf->reward || !buf->terminated || !buf->truncated || !buf->time_index || !buf->carry))
        return fail(PDECG_ERR_INVALID, "null device buffer");
    if (buf->n_envs > 0 && cfg->sensing_loc != PDECG_SENSE_FULL && !buf->obs) return fail(PDECG_ERR_INVALID, "scalar sensing needs an obs buffer");
    out->cfg = *cfg;
    out->buf = *buf;
    out->npts = pdecg_bc1d_npts(cfg);
    out->tuned_general = (cfg->reward_kind == PDECG_REWARD_TUNED1D && tuned_needs_general_ring(cfg->substeps, cfg->nt)) ? 1 : 0;
    out->ring_len = out->tuned_general ? kLookback + 1 : ring_len_for(cfg->substeps);
    out->carry_len = 1 + out->ring_len;
    out->ar = AutoReset{};
    out->map_envs = g_map_envs >= 0 ? g_map_envs : buf->n_envs;
    if (out->npts > kCtaMaxPoints) return fail(PDECG_ERR_UNSUPPORTED, "grid of %d points exceeds the shared-memory-resident limit of %d points per env", out->npts, kCtaMaxPoints);
    return PDECG_OK;
}

static int launch_mode(const Bc1dArgs& a, cudaStream_t st) {
    switch (a.cfg.precision) {
        case PDECG_PREC_REF: return a.cfg.beta_dtype == PDECG_F32 ? launch_bc1d_refb32(a, st) : launch_bc1d_refb64(a, st);
        case PDECG_PREC_FP64: return launch_bc1d_fp64(a, st);
        default: return launch_bc1d_fp32(a, st);
    }
}

}  // namespace pdecg

using namespace pdecg;

extern "C" int32_t pdecg_bc1d_npts(const pdecg_bc1d_cfg* cfg) {
    if (!cfg) return 0;
    return cfg->env == PDECG_ENV_REACTION_DIFFUSION ? cfg->nx + 1 : cfg->nx;   // parabolic.py:124
}

extern "C" int32_t pdecg_bc1d_carry_len(const pdecg_bc1d_cfg* cfg) {
    if (!cfg) return 0;
    if (cfg->reward_kind == PDECG_REWARD_TUNED1D && tuned_needs_general_ring(cfg->substeps, cfg->nt)) return 1 + kLookback + 1;
    return 1 + ring_len_for(cfg->substeps);
}

extern "C" int pdecg_bc1d_step(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, pdecg_stream_t stream) {
    Bc1dArgs a;
    int rc = validate(cfg, buf, &a);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    if (!buf->beta) return fail(PDECG_ERR_INVALID, "null beta buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return launch_mode(a, st);
}

extern "C" int pdecg_bc1d_step_autoreset(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf,
                                         const pdecg_bc1d_autoreset* ar, pdecg_stream_t stream) {
    Bc1dArgs a;
    int rc = validate(cfg, buf, &a);
    if (rc != PDECG_OK) return rc;
    if (!ar) return fail(PDECG_ERR_INVALID, "null autoreset descriptor");
    if (buf->n_envs == 0) return PDECG_OK;
    if (!buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    if (!buf->beta) return fail(PDECG_ERR_INVALID, "null beta buffer");
    if (!ar->init) return fail(PDECG_ERR_INVALID, "null init buffer");
    if (ar->init_dtype != PDECG_F32 && ar->init_dtype != PDECG_F64) return fail(PDECG_ERR_INVALID, "bad init_dtype");
    if (ar->random && !ar->episode) return fail(PDECG_ERR_INVALID, "null episode counter buffer");
    a.ar.enabled = 1;
    a.ar.init_dtype = ar->init_dtype;
    a.ar.init = ar->init;
    a.ar.init_stride = ar->random ? 0 : ar->init_env_stride;
    a.ar.rnd = ar->random ? RandomIc{1, ar->lo, ar->hi, ar->seed, ar->env_id_offset, ar->episode} : RandomIc{};
    a.ar.final_obs = ar->final_obs;
    return launch_mode(a, static_cast<cudaStream_t>(stream));
}

static int reset_impl(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, const void* init, int32_t init_dtype,
                      int64_t init_env_stride, const uint8_t* mask, const RandomIc& rnd, pdecg_stream_t stream) {
    Bc1dArgs a;
    int rc = validate(cfg, buf, &a);
    if (rc != PDECG_OK) return rc;
    if (!init) return fail(PDECG_ERR_INVALID, "null init buffer");
    if (init_dtype != PDECG_F32 && init_dtype != PDECG_F64) return fail(PDECG_ERR_INVALID, "bad init_dtype");
    if (buf->n_envs == 0) return PDECG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((buf->n_envs + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const bool s64 = cfg->precision == PDECG_PREC_FP64;
    if (s64) {
        if (init_dtype == PDECG_F32) bc1d_reset_kernel<double, float><<<grid, kThreads, 0, st>>>(a, static_cast<const float*>(init), init_env_stride, mask, rnd);
        else bc1d_reset_kernel<double, double><<<grid, kThreads, 0, st>>>(a, static_cast<const double*>(init), init_env_stride, mask, rnd);
    } else {
        if (init_dtype == PDECG_F32) bc1d_reset_kernel<float, float><<<grid, kThreads, 0, st>>>(a, static_cast<const float*>(init), init_env_stride, mask, rnd);
        else bc1d_reset_kernel<float, double><<<grid, kThreads, 0, st>>>(a, static_cast<const double*>(init), init_env_stride, mask, rnd);
    }
    count_launch();
    return check_launch("bc1d_reset_kernel");
}

extern "C" int pdecg_bc1d_reset(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, const void* init,
                                int32_t init_dtype, int64_t init_env_stride, const uint8_t* mask,
                                pdecg_stream_t stream) {
    RandomIc rnd{};
    return reset_impl(cfg, buf, init, init_dtype, init_env_stride, mask, rnd, stream);
}

extern "C" int pdecg_bc1d_reset_random(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf, const void* profile,
                                       int32_t profile_dtype, double lo, double hi, uint64_t seed, int64_t env_id_offset,
                                       int32_t* episode, const uint8_t* mask, pdecg_stream_t stream) {
    if (!episode && buf && buf->n_envs > 0) return fail(PDECG_ERR_INVALID, "null episode counter buffer");
    RandomIc rnd{1, lo, hi, seed, env_id_offset, episode};
    return reset_impl(cfg, buf, profile, profile_dtype, 0, mask, rnd, stream);
}

// Host-buffer step, pipelined: the batch is cut into chunks that rotate over four library-owned streams, so
// the D2H copy of chunk c overlaps the kernels of the chunks behind it (the observation copy-out is the
// PCIe-bound part of an end-to-end step with full-state sensing).
namespace {
constexpr int kPipeStreams = 4;
struct HostPipe {
    cudaStream_t s[kPipeStreams] = {};
    cudaStream_t small = nullptr;                    // reward / flag copies: one transfer each for the whole batch
    cudaEvent_t start = nullptr, done[kPipeStreams] = {}, kernels[kPipeStreams] = {}, small_done = nullptr;
    bool ready = false;
};
constexpr int kMaxDevices = 64;
thread_local HostPipe g_pipes[kMaxDevices];     // one pair of streams per (host thread, device), created on first use

int ensure_pipe(HostPipe** out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    if (dev < 0 || dev >= kMaxDevices) return fail(PDECG_ERR_UNSUPPORTED, "device index %d out of range", dev);
    HostPipe& p = g_pipes[dev];
    if (!p.ready) {
        for (int i = 0; i < kPipeStreams; ++i) {
            if ((e = cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
            if ((e = cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
            if ((e = cudaEventCreateWithFlags(&p.kernels[i], cudaEventDisableTiming)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
        }
        if ((e = cudaStreamCreateWithFlags(&p.small, cudaStreamNonBlocking)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        if ((e = cudaEventCreateWithFlags(&p.small_done, cudaEventDisableTiming)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
        if ((e = cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
        p.ready = true;
    }
    *out = &p;
    return PDECG_OK;
}
}  // namespace

extern "C" int pdecg_bc1d_step_host(const pdecg_bc1d_cfg* cfg, const pdecg_bc1d_buffers* buf,
                                    const void* action_host, void* obs_host, double* reward_host,
                                    uint8_t* terminated_host, uint8_t* truncated_host,
                                    const pdecg_bc1d_autoreset* autoreset, pdecg_stream_t stream) {
    Bc1dArgs a;
    int rc = validate(cfg, buf, &a);
    if (rc != PDECG_OK) return rc;
    if (buf->n_envs == 0) return PDECG_OK;
    if (!action_host || !buf->action) return fail(PDECG_ERR_INVALID, "null action buffer");
    HostPipe* pipe = nullptr;
    if ((rc = ensure_pipe(&pipe)) != PDECG_OK) return rc;
    HostPipe& g_pipe = *pipe;
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    const int64_t n = buf->n_envs;
    const size_t asz = cfg->action_dtype == PDECG_F32 ? 4 : 8;
    const size_t ssz = cfg->precision == PDECG_PREC_FP64 ? 8 : 4;
    const size_t bsz = (cfg->precision == PDECG_PREC_REF) ? (cfg->beta_dtype == PDECG_F64 ? 8 : 4) : ssz;
    const bool full = cfg->sensing_loc == PDECG_SENSE_FULL;
    // FULL sensing: the observation leaves through obs_copy when the caller gave one (possibly float32), else from the state
    const bool via_copy = full && buf->obs_copy != nullptr;
    const size_t osz = (via_copy && buf->obs_copy_f32) ? 4 : ssz;
    const size_t obs_row = full ? static_cast<size_t>(a.npts) * osz : ssz;
    // chunks of >= 4096 envs, at most 8 of them (PDECG_HOST_CHUNKS overrides the cap: A/B runs)
    static const int max_chunks = [] { const char* e = getenv("PDECG_HOST_CHUNKS"); const int v = e ? atoi(e) : 8; return v < 1 ? 1 : (v > 64 ? 64 : v); }();
    int nchunk = static_cast<int>(n / 4096);
    nchunk = nchunk < 1 ? 1 : (nchunk > max_chunks ? max_chunks : nchunk);
    const int64_t per = ((n + nchunk - 1) / nchunk + 3) / 4 * 4;
    // a short first chunk (a quarter of the others) gets the device->host queue going early: nothing can be copied out
    // before the first kernel has finished
    const int64_t first = (nchunk > 1) ? (per / 4 + 3) / 4 * 4 : per;
    // all actions in ONE transfer on the caller's stream (ordered after whatever the caller queued); the chunks fan out behind it
    cudaError_t e = cudaMemcpyAsync(const_cast<void*>(buf->action), action_host, n * asz, cudaMemcpyHostToDevice, user);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "H2D actions: %s", cudaGetErrorString(e));
    if ((e = cudaEventRecord(g_pipe.start, user)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
    for (int i = 0; i < kPipeStreams; ++i) cudaStreamWaitEvent(g_pipe.s[i], g_pipe.start, 0);
    MapEnvsScope map_scope(n);                                // every chunk uses the lane mapping of the whole batch
    auto off = [](const void* p, size_t bytes) { return p ? static_cast<const char*>(p) + bytes : nullptr; };
    int c = 0;
    for (int64_t lo = 0; lo < n; ++c) {
        const int64_t want = (c == 0) ? first : per, cnt = (lo + want <= n) ? want : n - lo;
        const bool last_on_stream = lo + cnt + (kPipeStreams - 1) * per >= n;      // no later chunk lands on this stream
        cudaStream_t st = g_pipe.s[c % kPipeStreams];
        pdecg_bc1d_buffers b = *buf;
        b.n_envs = cnt;
        b.u = const_cast<char*>(off(buf->u, lo * a.npts * ssz));
        b.beta = off(buf->beta, lo * buf->beta_env_stride * bsz);
        b.action = off(buf->action, lo * asz);
        b.obs = const_cast<char*>(off(buf->obs, lo * ssz));
        b.reward = buf->reward + lo;
        b.terminated = buf->terminated + lo;
        b.truncated = buf->truncated + lo;
        b.time_index = buf->time_index + lo;
        b.carry = buf->carry + lo * a.carry_len;
        b.obs_copy = via_copy ? const_cast<char*>(off(buf->obs_copy, lo * obs_row)) : nullptr;
        if (autoreset) {
            pdecg_bc1d_autoreset ar = *autoreset;
            const size_t isz = ar.init_dtype == PDECG_F64 ? 8 : 4;
            if (!ar.random) ar.init = off(ar.init, lo * ar.init_env_stride * isz);
            ar.env_id_offset += lo;
            if (ar.episode) ar.episode += lo;
            if (ar.final_obs) ar.final_obs = const_cast<char*>(off(ar.final_obs, lo * (full ? a.npts : 1) * ((buf->obs_copy_f32 && ssz == 8) ? 4 : ssz)));
            rc = pdecg_bc1d_step_autoreset(cfg, &b, &ar, st);
        } else {
            rc = pdecg_bc1d_step(cfg, &b, st);
        }
        if (rc != PDECG_OK) return rc;
        // the stream's last kernel: from here on only this chunk's observation copy follows on it
        if (last_on_stream) cudaEventRecord(g_pipe.kernels[c % kPipeStreams], st);
        if (obs_host) {
            const void* src = via_copy ? static_cast<const void*>(b.obs_copy) : (full ? static_cast<const void*>(b.u) : static_cast<const void*>(b.obs));
            if ((e = cudaMemcpyAsync(static_cast<char*>(obs_host) + lo * obs_row, src, cnt * obs_row, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
                return fail(PDECG_ERR_CUDA, "D2H obs: %s", cudaGetErrorString(e));
        }
        lo += cnt;
    }
    // reward / flags: one transfer each for the whole batch, once every chunk's kernel has run -- on their own stream, so
    // they overlap the last observation copies instead of fragmenting the D2H queue chunk by chunk
    const int used = c < kPipeStreams ? c : kPipeStreams;
    if (reward_host || terminated_host || truncated_host) {
        for (int i = 0; i < used; ++i) cudaStreamWaitEvent(g_pipe.small, g_pipe.kernels[i], 0);
        if (reward_host && (e = cudaMemcpyAsync(reward_host, buf->reward, n * 8, cudaMemcpyDeviceToHost, g_pipe.small)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H reward: %s", cudaGetErrorString(e));
        if (terminated_host && (e = cudaMemcpyAsync(terminated_host, buf->terminated, n, cudaMemcpyDeviceToHost, g_pipe.small)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H terminated: %s", cudaGetErrorString(e));
        if (truncated_host && (e = cudaMemcpyAsync(truncated_host, buf->truncated, n, cudaMemcpyDeviceToHost, g_pipe.small)) != cudaSuccess) return fail(PDECG_ERR_CUDA, "D2H truncated: %s", cudaGetErrorString(e));
        cudaEventRecord(g_pipe.small_done, g_pipe.small);
        cudaStreamWaitEvent(user, g_pipe.small_done, 0);
    }
    for (int i = 0; i < kPipeStreams; ++i) {                   // the caller's stream continues after every pipe
        cudaEventRecord(g_pipe.done[i], g_pipe.s[i]);
        cudaStreamWaitEvent(user, g_pipe.done[i], 0);
    }
    e = cudaStreamSynchronize(user);
    if (e != cudaSuccess) return fail(PDECG_ERR_CUDA, "stream sync: %s", cudaGetErrorString(e));
    return PDECG_OK;
}
