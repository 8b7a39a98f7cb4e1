// mate_b200.cu -- host side of the C ABI declared in include/mate_b200.h.
//
// Owns the struct-of-arrays device state of one batch of environments and launches the
// fused step kernel (mate_step.cuh).  No torch types, no CPU simulation fallback.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <memory>
#include <sched.h>
#include <mutex>
#include <string>
#include <vector>

#include "mate_step.cuh"
#include "mate_wrappers.cuh"
#include "mate_agents.cuh"
#include "mate_hostpath.cuh"

using namespace mate;

static thread_local std::string g_error;

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                             \
            return fail(MATE_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));   \
    } while (0)

// ---- supported entity-count shapes (compile-time specialisations of the kernel) ------------
// All 17 presets of the reference (mate/assets/MATE-*.yaml).
#if defined(MATE_DEV_SHAPE_OTHERS)   // development builds: the other BASELINE shapes
#define MATE_SHAPES(X) X(4, 2, 9) X(4, 8, 0) X(8, 8, 9) X(0, 8, 32)
#elif defined(MATE_DEV_SHAPE)   // development builds: a single specialisation (profiles/tools/build_variant.sh)
#define MATE_SHAPES(X) X(4, 8, 9)
#else
#define MATE_SHAPES(X)                                                                        \
    X(1, 1, 0) X(1, 1, 9) X(1, 2, 0) X(1, 2, 9) X(2, 2, 0) X(2, 2, 9) X(2, 4, 0) X(2, 4, 9)   \
    X(4, 2, 0) X(4, 2, 9) X(4, 4, 0) X(4, 4, 9) X(4, 8, 0) X(4, 8, 9) X(8, 8, 0) X(8, 8, 9)   \
    X(0, 8, 32)
#endif

struct KernelInfo {
    void (*launch_wrappers)(const Params&, const ObsOps&, const float*, const float*, int num_envs, cudaStream_t);
    cudaError_t (*prepare_wrappers)();
    void (*launch)(const Params&, int grid, cudaStream_t);
    int envs_per_cta, smem_bytes, dc, dt, epw;
    cudaError_t (*prepare)();
    void (*launch_fov)(const Params&, const int32_t*, const int32_t*, const double*, double*, long long, cudaStream_t);
    void (*launch_soft)(const Params&, const uint8_t*, const uint8_t*, double*, float*, cudaStream_t);
};

// ---- L2 persistence for the fp32 obstacle entries -------------------------------------------------
// The simulation phase of a step reads obs_f4 and the packer reads it again 60 us later, while the observation rows
// stream through the L2.  When the rows of one launch are several times the L2 (MATE-4v8-9 x 65 536: 407 MB against
// 126 MB) the entries are evicted in between and the re-reads go to HBM in the middle of the write phase.  Such a
// simulator sets aside a part of the L2 for persisting lines while it lives (12 MB by default, MATE_B200_L2_KEEP=<MB>,
// 0 = off) and its step launches carry a persisting access window over the array.  Measured (profiles/r2v_summary.md):
// 10 - 16 MB set aside -1.5 to -2.2 % ms/step, 24 MB +4 %, 32 MB +22 % (the write stream needs the rest of the L2);
// no gain for the smaller workloads and a loss when the array does not fit the set-aside (MATE-Navigation), so only
// simulators that meet both conditions take part.  The device limit is reference-counted and restored afterwards.
struct L2Keep {
    std::mutex mu;
    int users[64] = {};
    size_t previous[64] = {};
};
static L2Keep g_l2;
static size_t l2_keep_wanted() {
    size_t mb = 12;
    if (const char* v = getenv("MATE_B200_L2_KEEP")) mb = (size_t)std::max(0, atoi(v));
    return mb << 20;
}
// returns the window a simulator of this size gets (0 = none) and takes a reference on the device's set-aside
static size_t l2_keep_acquire(int dev, size_t array_bytes, size_t row_bytes_per_launch) {
    const size_t want = l2_keep_wanted();
    if (want == 0 || array_bytes == 0 || array_bytes > want || dev < 0 || dev >= 64) return 0;
    int l2 = 0, max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (l2 <= 0 || row_bytes_per_launch < 3 * (size_t)l2) return 0;     // the rows do not flush the L2 between the two reads
    if ((size_t)std::max(max_persist, 0) < want || (size_t)std::max(max_window, 0) < array_bytes) return 0;
    std::lock_guard<std::mutex> lock(g_l2.mu);
    if (g_l2.users[dev] == 0) {
        size_t current = 0;
        cudaDeviceGetLimit(&current, cudaLimitPersistingL2CacheSize);
        g_l2.previous[dev] = current;
        if (current < want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return 0; }
    }
    ++g_l2.users[dev];
    return array_bytes;
}
static void l2_keep_release(int dev) {
    std::lock_guard<std::mutex> lock(g_l2.mu);
    if (dev < 0 || dev >= 64 || g_l2.users[dev] == 0) return;
    if (--g_l2.users[dev] == 0) {
        cudaCtxResetPersistingL2Cache();
        if (g_l2.previous[dev] < l2_keep_wanted()) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, g_l2.previous[dev]);
        cudaGetLastError();
    }
}

template <int NC, int NT, int NO>
static void launch_shape2(const Params& p, int grid, cudaStream_t stream) {
    using S = Shape2<NC, NT, NO>;
    Params q = p;
    q.warp_stride = S::WARP_BYTES + (p.obs_ops.n > 0 ? S::OPS_BYTES : 0);   // scratch of the folded observation wrappers
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(S::WARPS * 32);
    cfg.dynamicSmemBytes = (size_t)S::WARPS * q.warp_stride; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (NO > 0 && q.l2_window_bytes > 0 && q.mode == MODE_STEP) {   // resets and prepared episodes are written once: nothing to keep
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = q.obs_f4;
        attr[0].val.accessPolicyWindow.num_bytes = (size_t)q.l2_window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, mate_step_kernel2<NC, NT, NO>, q);
}
template <int NC, int NT, int NO>
static void launch_fov_shape(const Params& p, const int32_t* env, const int32_t* camera, const double* angle, double* out,
                             long long n, cudaStream_t stream) {
    fov_range_kernel<NC, NO><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p, env, camera, angle, out, n);
}
template <int NC, int NT, int NO>
static void launch_soft_shape(const Params& p, const uint8_t* mask_ct, const uint8_t* done, double* edges, float* out, cudaStream_t stream) {
    const long long warps = (long long)p.num_envs * (NC > 0 ? NC : 1);
    soft_edges_kernel<NC, NO><<<(unsigned)((warps * 2 + 127) / 128), 128, 0, stream>>>(p, done, edges);
    soft_coverage_kernel<NC, NT, NO><<<(unsigned)((warps + 3) / 4), 128, 0, stream>>>(p, mask_ct, done, edges, out);
}
template <int NC, int NT, int NO>
static cudaError_t prepare_shape2() {
    using S = Shape2<NC, NT, NO>;
    return cudaFuncSetAttribute(mate_step_kernel2<NC, NT, NO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                S::SMEM_BYTES + S::WARPS * S::OPS_BYTES);
}

template <int NC, int NT, int NO>
static void launch_wrappers_shape(const Params& p, const ObsOps& ops, const float* cam_affine, const float* tgt_affine,
                                  int num_envs, cudaStream_t stream) {
    using W = WShape<NC, NT, NO>;
    const int grid = (num_envs + W::WARPS - 1) / W::WARPS;
    obs_transform_kernel<NC, NT, NO><<<grid, W::WARPS * 32, W::SMEM_BYTES, stream>>>(p, ops, cam_affine, tgt_affine);
}
template <int NC, int NT, int NO>
static cudaError_t prepare_wrappers_shape() {
    return cudaFuncSetAttribute(obs_transform_kernel<NC, NT, NO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                WShape<NC, NT, NO>::SMEM_BYTES);
}

static bool find_kernel(int nc, int nt, int no, KernelInfo* out) {
#define X(NC, NT, NO)                                                                          \
    if (nc == NC && nt == NT && no == NO) {                                                    \
        using S = Shape2<NC, NT, NO>;                                                          \
        *out = KernelInfo{&launch_wrappers_shape<NC, NT, NO>, &prepare_wrappers_shape<NC, NT, NO>,       \
                          &launch_shape2<NC, NT, NO>, S::ENVS_PER_CTA, S::SMEM_BYTES, S::DC, S::DT, \
                          32, &prepare_shape2<NC, NT, NO>, &launch_fov_shape<NC, NT, NO>,        \
                          &launch_soft_shape<NC, NT, NO>};                                      \
        return true;                                                                           \
    }
    MATE_SHAPES(X)
#undef X
    return false;
}

struct MateSim {
    MateConfig cfg;
    int num_envs = 0, bpad = 0, device = 0;
    long long env_index_base = 0;
    KernelInfo kernel{};
    Params base{};            // state pointers + scalars; per-call fields are filled per launch
    void* state_block = nullptr;
    size_t state_bytes = 0;
    double* d_ranges = nullptr;
    unsigned long long seed = 0;
    long long launches = 0;
    // prepared resets (mate_step.cuh): second state block, parameter block that addresses it, side stream
    void* next_block = nullptr;
    Params next_base{};       // = base with the state pointers of the second block
    Params* d_next = nullptr; // device copy of next_base (Params::next of the step launches)
    cudaStream_t side = nullptr;
    cudaEvent_t side_event = nullptr;
    double* soft_edges = nullptr;   // scratch of mate_b200_soft_coverage, allocated on first use
    int refill_mode = 0;      // 0 = off, 1 = side stream every refill_period steps, 2 = same stream after every step (tests)
    int refill_period = 512;
    int tile_envs = 32;       // environments per warp tile (Params::tile_envs), chosen from the batch size at creation
    long long steps_since_refill = 0;
    // host-path (step_host) resources
    static constexpr int kHostStreams = 4;
    cudaStream_t hstreams[kHostStreams] = {};
    cudaEvent_t hevents[kHostStreams] = {};
    float *h_cam_act = nullptr, *h_tgt_act = nullptr, *h_cam_obs = nullptr, *h_tgt_obs = nullptr, *h_rewards = nullptr;
    uint8_t* h_done = nullptr;
    bool host_ready = false;
    size_t l2_window = 0;                     // bytes of obs_f4 the step launches keep in the L2 (0: none), see l2_keep_acquire
    // compacted device -> host leg (mate_hostpath.cuh): region 0 = camera rows, 1 = target rows
    static constexpr int kMaxHostChunks = 64;
    int compact_mode = 0;                     // 1: rows cross the link compacted and are expanded by host threads
    uint4* d_compact[2] = {nullptr, nullptr}; // compact streams (a launch chunk's stream starts where its dense rows start)
    CompactEntry* d_table[2] = {nullptr, nullptr};
    unsigned int* d_count = nullptr;          // [kMaxHostChunks][2] chunks kept
    uint4* p_compact[2] = {nullptr, nullptr}; // pinned host copies of the above
    CompactEntry* p_table[2] = {nullptr, nullptr};
    // MATE_STEP_HOST_ROWS_KEPT: the rows of the previous call stay on the device (second pair of row buffers, used in turn);
    // only the 64-byte groups that differ from them cross the link and are rewritten in the caller's buffers
    float* h_rows2[2] = {nullptr, nullptr};
    int row_parity = 0;
    const float* kept_rows[2] = {nullptr, nullptr};  // the caller's buffers that hold the previous call's rows
    bool kept_valid = false;
    int last_leg = 0;                                // device -> host leg of the last step_host call (mate_b200_host_leg_info)
    unsigned long long last_link_bytes = 0;
    unsigned int* p_count = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t e_counts[kMaxHostChunks] = {}, e_stream[kMaxHostChunks] = {}, e_tables[kMaxHostChunks] = {};
    unsigned int* p_count_dev = nullptr;      // device address of the pinned p_count (the sizes are stored there by a kernel, not copied)
    std::unique_ptr<ExpandPool> pool;
};

extern "C" const char* mate_b200_last_error(void) { return g_error.c_str(); }
extern "C" int mate_b200_abi_version(void) { return MATE_B200_ABI_VERSION; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int mate_b200_create(const MateConfig* cfg, int32_t num_envs, int32_t device,
                                int64_t env_index_base, MateSim** out) {
    if (!cfg || !out) return fail(MATE_EINVAL, "null argument");
    if (num_envs <= 0) return fail(MATE_EINVAL, "num_envs must be positive");
    const int nc = cfg->num_cameras, nt = cfg->num_targets, no = cfg->num_obstacles;
    KernelInfo kernel;
    if (!find_kernel(nc, nt, no, &kernel)) {
        char buf[256];
        snprintf(buf, sizeof(buf),
                 "no kernel specialisation for (cameras=%d, targets=%d, obstacles=%d); add it to MATE_SHAPES in "
                 "mate_b200/csrc/mate_b200.cu and rebuild", nc, nt, no);
        return fail(MATE_EINVAL, buf);
    }
    if ((nc > 0 && !cfg->camera_location_ranges) || !cfg->target_location_ranges || (no > 0 && !cfg->obstacle_location_ranges))
        return fail(MATE_EINVAL, "missing location ranges");
    if (!(cfg->target_step_size > 0.0) || !(cfg->target_sight_range > 0.0)) return fail(MATE_EINVAL, "target parameters must be positive");
    if (nc > 0 && (!(cfg->camera_min_viewing_angle > 0.0) || cfg->camera_min_viewing_angle > 180.0 || !(cfg->camera_rotation_step > 0.0) ||
                   !(cfg->camera_zooming_step > 0.0) || !(cfg->camera_max_sight_range > 0.0)))
        return fail(MATE_EINVAL, "camera parameters out of range");
    if (cfg->max_episode_steps <= 0) return fail(MATE_EINVAL, "max_episode_steps must be positive");
    if (cfg->num_cargoes_per_target < MATE_NUM_WAREHOUSES) return fail(MATE_EINVAL, "num_cargoes_per_target must be >= 4");
    if (cfg->num_high_capacity_targets < 0 || cfg->num_high_capacity_targets > nt) return fail(MATE_EINVAL, "bad num_high_capacity_targets");
    const double freight_scale = std::ceil(2.0 * kTerrain / cfg->target_step_size);       // environment.py:521
    const double bounty_scale = std::ceil(freight_scale * std::fmax(0.0, cfg->bounty_factor));
    if (2.0 * bounty_scale > 65535.0 || (double)cfg->num_cargoes_per_target * nt > 65535.0)
        return fail(MATE_EINVAL, "bounty/cargo counts exceed the packed 16-bit state fields");

    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(kernel.prepare());
    CUDA_TRY(kernel.prepare_wrappers());
    MateSim* sim = new MateSim();
    sim->cfg = *cfg;
    sim->cfg.camera_location_ranges = sim->cfg.target_location_ranges = sim->cfg.obstacle_location_ranges = nullptr;
    sim->num_envs = num_envs;
    sim->device = device;
    sim->env_index_base = env_index_base;
    sim->kernel = kernel;
    const int bpad = (int)align_up((size_t)num_envs, 256);
    sim->bpad = bpad;

    // one allocation, carved into the SoA arrays (each 256-byte aligned)
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cam_x = carve(sizeof(double) * nc * bpad), o_cam_y = carve(sizeof(double) * nc * bpad);
    const size_t o_cam_phi = carve(sizeof(double) * nc * bpad), o_cam_theta = carve(sizeof(double) * nc * bpad);
    const size_t o_tgt_x = carve(sizeof(double) * nt * bpad), o_tgt_y = carve(sizeof(double) * nt * bpad);
    const size_t o_obs_x = carve(sizeof(double) * no * bpad), o_obs_y = carve(sizeof(double) * no * bpad), o_obs_r = carve(sizeof(double) * no * bpad);
    const size_t o_obs_f4 = carve(sizeof(float4) * no * bpad);
    const size_t o_pack = carve(sizeof(uint32_t) * nt * bpad);
    const size_t o_cargo = carve(sizeof(uint4) * 2 * bpad), o_env_a = carve(sizeof(uint4) * bpad), o_env_b = carve(sizeof(int4) * bpad);
    const size_t o_cc = carve(sizeof(unsigned long long) * bpad);
    const size_t o_stats = carve(sizeof(float) * 16);
    sim->state_bytes = off;
    if (cudaMalloc(&sim->state_block, off) != cudaSuccess) { delete sim; return fail(MATE_ENOMEM, "cudaMalloc(state) failed"); }
    CUDA_TRY(cudaMemset(sim->state_block, 0, off));
    char* b = (char*)sim->state_block;
    Params& p = sim->base;
    p.cam_x = (double*)(b + o_cam_x); p.cam_y = (double*)(b + o_cam_y); p.cam_phi = (double*)(b + o_cam_phi); p.cam_theta = (double*)(b + o_cam_theta);
    p.tgt_x = (double*)(b + o_tgt_x); p.tgt_y = (double*)(b + o_tgt_y);
    p.obs_x = (double*)(b + o_obs_x); p.obs_y = (double*)(b + o_obs_y); p.obs_r = (double*)(b + o_obs_r);
    p.obs_f4 = (float4*)(b + o_obs_f4);
    p.tgt_pack = (uint32_t*)(b + o_pack);
    p.cargo = (uint4*)(b + o_cargo); p.env_a = (uint4*)(b + o_env_a); p.env_b = (int4*)(b + o_env_b);
    p.cc_clear = (unsigned long long*)(b + o_cc);
    p.stats = (float*)(b + o_stats);

    // Warp tiles: 32 environments per warp for the batches that give every SM several of those; smaller batches are cut
    // into tiles of 16 or 8 so that the latency-bound launch still has warps to overlap (measured, MATE-4v8-9: 16 384
    // envs 0.0697 / 0.0622 / 0.0685 ms with 32 / 16 / 8, 4096 envs 0.0610 / 0.0468 / 0.0417; from 32 768 envs on 32 wins:
    // MATE-8v8-9 x 32 768 0.129 / 0.178 / 0.218 -- idle lanes still cost issue slots).
    sim->tile_envs = num_envs >= 24576 ? 32 : (num_envs >= 8192 ? 16 : 8);
    if (const char* v = getenv("MATE_B200_TILE")) { const int t = atoi(v); if (t == 8 || t == 16 || t == 32) sim->tile_envs = t; }
    p.tile_envs = sim->tile_envs;

    // prepared resets: a second state block of the same layout + first-view masks + the ready tags
    sim->refill_mode = 1;
    if (const char* v = getenv("MATE_B200_REFILL")) {
        if (!strcmp(v, "0") || !strcmp(v, "off")) sim->refill_mode = 0;
        else if (!strcmp(v, "sync")) sim->refill_mode = sim->refill_mode ? 2 : 0;
        else if (atoi(v) > 0) sim->refill_period = atoi(v);
    }
    if (sim->refill_mode) {
        const int mask_words = (nc + nt) * (no <= 16 ? 1 : 2);
        const size_t o_masks = carve(sizeof(uint32_t) * mask_words * bpad);
        const size_t o_vals = carve(sizeof(float) * (3 * nt + 5 * nc) * bpad);
        const size_t o_ready = carve(sizeof(uint32_t) * bpad);
        const size_t block2 = off;
        if (cudaMalloc(&sim->next_block, block2) != cudaSuccess) { cudaFree(sim->state_block); delete sim; return fail(MATE_ENOMEM, "cudaMalloc(prepared state) failed"); }
        CUDA_TRY(cudaMemset(sim->next_block, 0, block2));
        char* b2 = (char*)sim->next_block;
        Params& n = sim->next_base;
        n = p;   // scalars are copied again below, once they are set
        n.cam_x = (double*)(b2 + o_cam_x); n.cam_y = (double*)(b2 + o_cam_y); n.cam_phi = (double*)(b2 + o_cam_phi); n.cam_theta = (double*)(b2 + o_cam_theta);
        n.tgt_x = (double*)(b2 + o_tgt_x); n.tgt_y = (double*)(b2 + o_tgt_y);
        n.obs_x = (double*)(b2 + o_obs_x); n.obs_y = (double*)(b2 + o_obs_y); n.obs_r = (double*)(b2 + o_obs_r);
        n.obs_f4 = (float4*)(b2 + o_obs_f4);
        n.tgt_pack = (uint32_t*)(b2 + o_pack);
        n.cargo = (uint4*)(b2 + o_cargo); n.env_a = (uint4*)(b2 + o_env_a); n.env_b = (int4*)(b2 + o_env_b);
        n.cc_clear = (unsigned long long*)(b2 + o_cc);
        n.stats = (float*)(b2 + o_stats);
        n.masks = (uint32_t*)(b2 + o_masks);
        n.vals = (float*)(b2 + o_vals);
        n.ready = (uint32_t*)(b2 + o_ready);
        p.ready = n.ready;
        CUDA_TRY(cudaMalloc(&sim->d_next, sizeof(Params)));
        CUDA_TRY(cudaStreamCreateWithFlags(&sim->side, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&sim->side_event, cudaEventDisableTiming));
    }

    // location ranges on device (reset)
    std::vector<double> ranges((size_t)4 * (nc + nt + no) + 4, 0.0);
    if (nc) memcpy(ranges.data(), cfg->camera_location_ranges, sizeof(double) * 4 * nc);
    memcpy(ranges.data() + 4 * nc, cfg->target_location_ranges, sizeof(double) * 4 * nt);
    if (no) memcpy(ranges.data() + 4 * (nc + nt), cfg->obstacle_location_ranges, sizeof(double) * 4 * no);
    CUDA_TRY(cudaMalloc(&sim->d_ranges, ranges.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpy(sim->d_ranges, ranges.data(), ranges.size() * sizeof(double), cudaMemcpyHostToDevice));
    p.cam_ranges = sim->d_ranges; p.tgt_ranges = sim->d_ranges + 4 * nc; p.obs_ranges = sim->d_ranges + 4 * (nc + nt);

    p.num_envs = num_envs; p.bpad = bpad; p.env_index_base = env_index_base; p.seed = 0;
    sim->l2_window = l2_keep_acquire(device, no > 0 ? ((size_t)(no - 1) * bpad + (size_t)num_envs) * sizeof(float4) : 0,
                                     (size_t)num_envs * ((size_t)nc * sim->kernel.dc + (size_t)nt * sim->kernel.dt) * sizeof(float));
    p.l2_window_bytes = sim->l2_window;
    p.max_episode_steps = cfg->max_episode_steps; p.num_cargoes_per_target = cfg->num_cargoes_per_target;
    p.num_high_capacity = cfg->num_high_capacity_targets; p.start_with_cargoes = cfg->targets_start_with_cargoes != 0;
    p.shuffle = cfg->shuffle_entities != 0; p.reward_sparse = cfg->reward_sparse != 0;
    p.transmittance = std::fmin(std::fmax(cfg->obstacle_transmittance, 0.0), 1.0);   // environment.py:1470-1476
    p.transmittance_is_one = p.transmittance == 1.0;
    p.freight_scale = (int)freight_scale; p.bounty_scale = (int)bounty_scale; p.reward_scale = (int)(freight_scale + bounty_scale);
    p.cam_radius = cfg->camera_radius; p.cam_min_view = cfg->camera_min_viewing_angle; p.cam_rmax = cfg->camera_max_sight_range;
    p.cam_rot_step = cfg->camera_rotation_step; p.cam_zoom_step = cfg->camera_zooming_step;
    p.cam_area_product = cfg->camera_min_viewing_angle * (cfg->camera_max_sight_range * cfg->camera_max_sight_range);  // entities.py:285
    p.tgt_step_size = cfg->target_step_size; p.tgt_sight_range = cfg->target_sight_range;
    p.obs_r_low = cfg->obstacle_radius_low; p.obs_r_high = cfg->obstacle_radius_high;

    if (sim->refill_mode) {
        Params n = p;   // all scalars / range pointers of the live block ...
        const Params& o = sim->next_base;   // ... with the state pointers of the second block
        n.cam_x = o.cam_x; n.cam_y = o.cam_y; n.cam_phi = o.cam_phi; n.cam_theta = o.cam_theta; n.tgt_x = o.tgt_x; n.tgt_y = o.tgt_y;
        n.obs_x = o.obs_x; n.obs_y = o.obs_y; n.obs_r = o.obs_r; n.obs_f4 = o.obs_f4; n.tgt_pack = o.tgt_pack;
        n.cargo = o.cargo; n.env_a = o.env_a; n.env_b = o.env_b; n.cc_clear = o.cc_clear; n.stats = o.stats;
        n.masks = o.masks; n.vals = o.vals; n.ready = o.ready; n.live_env_b = p.env_b; n.next = nullptr;
        sim->next_base = n;
        CUDA_TRY(cudaMemcpy(sim->d_next, &sim->next_base, sizeof(Params), cudaMemcpyHostToDevice));
        p.next = sim->d_next;
    }

    // neutral initial state so that a step before reset/set_state is well defined
    {
        std::vector<double> theta((size_t)nc * bpad, cfg->camera_min_viewing_angle > 0 ? cfg->camera_min_viewing_angle : 90.0);
        if (nc) CUDA_TRY(cudaMemcpy(p.cam_theta, theta.data(), theta.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<uint32_t> pack((size_t)nt * bpad, pack_target(0, -1, 0, 1, 0, 0));
        CUDA_TRY(cudaMemcpy(p.tgt_pack, pack.data(), pack.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    *out = sim;
    return MATE_OK;
}

extern "C" int mate_b200_destroy(MateSim* sim) {
    if (!sim) return MATE_OK;
    cudaSetDevice(sim->device);
    if (sim->host_ready) {
        for (int i = 0; i < MateSim::kHostStreams; ++i) { cudaStreamDestroy(sim->hstreams[i]); cudaEventDestroy(sim->hevents[i]); }
        cudaFree(sim->h_cam_act); cudaFree(sim->h_tgt_act); cudaFree(sim->h_cam_obs); cudaFree(sim->h_tgt_obs);
        cudaFree(sim->h_rewards); cudaFree(sim->h_done);
        sim->pool.reset();
        for (int r = 0; r < 2; ++r) { cudaFree(sim->d_compact[r]); cudaFree(sim->d_table[r]); cudaFreeHost(sim->p_compact[r]); cudaFreeHost(sim->p_table[r]); cudaFree(sim->h_rows2[r]); }
        cudaFree(sim->d_count); cudaFreeHost(sim->p_count);
        if (sim->copy_stream) {
            cudaStreamDestroy(sim->copy_stream);
            for (int i = 0; i < MateSim::kMaxHostChunks; ++i) { cudaEventDestroy(sim->e_counts[i]); cudaEventDestroy(sim->e_stream[i]); cudaEventDestroy(sim->e_tables[i]); }
        }
    }
    if (sim->side) { cudaStreamSynchronize(sim->side); cudaStreamDestroy(sim->side); }
    if (sim->side_event) cudaEventDestroy(sim->side_event);
    cudaFree(sim->soft_edges);
    cudaFree(sim->next_block);
    cudaFree(sim->d_next);
    cudaFree(sim->state_block);
    cudaFree(sim->d_ranges);
    if (sim->l2_window) l2_keep_release(sim->device);
    delete sim;
    return MATE_OK;
}

extern "C" int mate_b200_obs_dims(const MateSim* sim, int32_t* cam_dim, int32_t* tgt_dim) {
    if (!sim) return fail(MATE_EINVAL, "null handle");
    if (cam_dim) *cam_dim = sim->kernel.dc;
    if (tgt_dim) *tgt_dim = sim->kernel.dt;
    return MATE_OK;
}

extern "C" int mate_b200_fov_range(MateSim* sim, const int32_t* env, const int32_t* camera, const double* angle_deg,
                                   double* out, int64_t n, void* stream) {
    if (!sim || !env || !camera || !angle_deg || !out || n < 0) return fail(MATE_EINVAL, "bad argument");
    if (sim->cfg.num_cameras == 0) return fail(MATE_EINVAL, "the configuration has no cameras");
    if (n == 0) return MATE_OK;
    CUDA_TRY(cudaSetDevice(sim->device));
    sim->kernel.launch_fov(sim->base, env, camera, angle_deg, out, n, (cudaStream_t)stream);
    sim->launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("fov_range launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int64_t mate_b200_launch_count(const MateSim* sim) { return sim ? sim->launches : 0; }
extern "C" int mate_b200_host_leg_info(const MateSim* sim, int32_t* leg, uint64_t* row_bytes_on_link) {
    if (!sim) return fail(MATE_EINVAL, "null handle");
    if (leg) *leg = sim->last_leg;
    if (row_bytes_on_link) *row_bytes_on_link = sim->last_link_bytes;
    return MATE_OK;
}

extern "C" int mate_b200_transform_observations(MateSim* sim, float* cam_obs, float* tgt_obs, const int32_t* ops,
                                                int32_t num_ops, const float* cam_affine, const float* tgt_affine,
                                                void* stream) {
    if (!sim || !tgt_obs || (sim->cfg.num_cameras > 0 && !cam_obs) || (num_ops > 0 && !ops)) return fail(MATE_EINVAL, "null argument");
    if (num_ops < 0 || num_ops > MATE_MAX_OBS_OPS) return fail(MATE_EINVAL, "too many observation wrappers (MATE_MAX_OBS_OPS)");
    if (num_ops == 0) return MATE_OK;
    if (((uintptr_t)cam_obs & 15) || ((uintptr_t)tgt_obs & 15)) return fail(MATE_EINVAL, "observation buffers must be 16-byte aligned");
    ObsOps o{};
    o.n = num_ops;
    for (int i = 0; i < num_ops; ++i) {
        if (ops[i] < MATE_OBS_ENHANCED_CAMERA || ops[i] > MATE_OBS_RESCALED) return fail(MATE_EINVAL, "unknown observation wrapper code");
        if (ops[i] == MATE_OBS_RESCALED && (!tgt_affine || (sim->cfg.num_cameras > 0 && !cam_affine)))
            return fail(MATE_EINVAL, "MATE_OBS_RESCALED needs the affine tables");
        o.op[i] = ops[i];
    }
    CUDA_TRY(cudaSetDevice(sim->device));
    Params p = sim->base;
    p.cam_obs = cam_obs; p.tgt_obs = tgt_obs;
    sim->kernel.launch_wrappers(p, o, cam_affine, tgt_affine, sim->num_envs, (cudaStream_t)stream);
    sim->launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("wrapper kernel launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int mate_b200_set_observation_ops(MateSim* sim, const int32_t* ops, int32_t num_ops, const float* cam_affine,
                                             const float* tgt_affine) {
    if (!sim || (num_ops > 0 && !ops)) return fail(MATE_EINVAL, "null argument");
    if (num_ops < 0 || num_ops > MATE_MAX_OBS_OPS) return fail(MATE_EINVAL, "too many observation wrappers (MATE_MAX_OBS_OPS)");
    ObsOps o{};
    o.n = num_ops;
    for (int i = 0; i < num_ops; ++i) {
        if (ops[i] < MATE_OBS_ENHANCED_CAMERA || ops[i] > MATE_OBS_RESCALED) return fail(MATE_EINVAL, "unknown observation wrapper code");
        if (ops[i] == MATE_OBS_RESCALED && (!tgt_affine || (sim->cfg.num_cameras > 0 && !cam_affine)))
            return fail(MATE_EINVAL, "MATE_OBS_RESCALED needs the affine tables");
        o.op[i] = ops[i];
    }
    // canonical stacks -- (Enhanced | Shared)* Relative? Rescaled? -- are applied while the rows are composed (FoldOps)
    FoldOps f{};
    int k = 0;
    while (k < num_ops && ops[k] >= MATE_OBS_ENHANCED_CAMERA && ops[k] <= MATE_OBS_SHARED_TARGET) f.mask_op[f.n_mask++] = ops[k++];
    if (k < num_ops && ops[k] == MATE_OBS_RELATIVE) { f.relative = 1; ++k; }
    if (k < num_ops && ops[k] == MATE_OBS_RESCALED) { f.rescaled = 1; ++k; }
    f.fast = (num_ops > 0 && k == num_ops) ? 1 : 0;
    if (f.fast && f.rescaled) {
        CUDA_TRY(cudaSetDevice(sim->device));
        const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets, no = sim->cfg.num_obstacles;
        const int dc = sim->kernel.dc, dt = sim->kernel.dt;
        std::vector<float> cam((size_t)2 * std::max(dc, 1), 0.f), tgt((size_t)2 * dt, 0.f);
        if (nc) CUDA_TRY(cudaMemcpy(cam.data(), cam_affine, sizeof(float) * 2 * dc, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(tgt.data(), tgt_affine, sizeof(float) * 2 * dt, cudaMemcpyDeviceToHost));
        auto col = [](const std::vector<float>& t, int c) { return make_float2(t[2 * c], t[2 * c + 1]); };
        auto same = [](float2 a, float2 b) { return a.x == b.x && a.y == b.y; };
        bool uniform = true;
        for (int j = 0; j < 13; ++j) { f.pres[j] = col(tgt, j); if (nc) uniform = uniform && same(f.pres[j], col(cam, j)); }
        for (int j = 0; j < 14; ++j) f.tself[j] = col(tgt, 13 + j);
        for (int j = 0; j < 9 && nc; ++j) f.cself[j] = col(cam, 13 + j);
        const int t_tgt = 27 + 7 * nc + 4 * no, t_obs = 27 + 7 * nc, t_cam = 27, c_tgt = 22, c_obs = 22 + 5 * nt, c_cam = 22 + 5 * nt + 4 * no;
        for (int j = 0; j < 5; ++j) f.tgt[j] = col(tgt, t_tgt + j);
        for (int j = 0; j < 4 && no; ++j) f.obs[j] = col(tgt, t_obs + j);
        for (int j = 0; j < 7 && nc; ++j) f.cam[j] = col(tgt, t_cam + j);
        // every entry of a kind has the same bounds in both teams' rows (mate/constants.py:195-254): checked, not assumed
        for (int e = 0; e < nt; ++e) for (int j = 0; j < 5; ++j) {
            uniform = uniform && same(f.tgt[j], col(tgt, t_tgt + 5 * e + j));
            if (nc) uniform = uniform && same(f.tgt[j], col(cam, c_tgt + 5 * e + j));
        }
        for (int e = 0; e < no; ++e) for (int j = 0; j < 4; ++j) {
            uniform = uniform && same(f.obs[j], col(tgt, t_obs + 4 * e + j));
            if (nc) uniform = uniform && same(f.obs[j], col(cam, c_obs + 4 * e + j));
        }
        for (int e = 0; e < nc; ++e) for (int j = 0; j < 7; ++j)
            uniform = uniform && same(f.cam[j], col(tgt, t_cam + 7 * e + j)) && same(f.cam[j], col(cam, c_cam + 7 * e + j));
        if (!uniform) f.fast = 0;   // tables of the caller's own: generic path
    }
    if (const char* v = getenv("MATE_B200_FOLD")) { if (v[0] == '0') f.fast = 0; }   // tests: force the generic shared-memory path
    sim->base.obs_ops = o;
    sim->base.cam_affine = cam_affine;
    sim->base.tgt_affine = tgt_affine;
    sim->base.fold = f;
    return MATE_OK;
}

extern "C" int mate_b200_soft_coverage(MateSim* sim, const uint8_t* mask_ct, const uint8_t* done, float* soft_matrix, void* stream) {
    if (!sim || !mask_ct || !soft_matrix) return fail(MATE_EINVAL, "null argument");
    if (sim->cfg.num_cameras == 0) return fail(MATE_EINVAL, "the configuration has no cameras");
    CUDA_TRY(cudaSetDevice(sim->device));
    if (!sim->soft_edges)   // ranges at the two sector edges of every camera, filled by the pre-pass
        CUDA_TRY(cudaMalloc(&sim->soft_edges, sizeof(double) * 2 * (size_t)sim->num_envs * sim->cfg.num_cameras));
    sim->kernel.launch_soft(sim->base, mask_ct, done, sim->soft_edges, soft_matrix, (cudaStream_t)stream);
    sim->launches += 2;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("soft coverage launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int mate_b200_auxiliary_terms(MateSim* sim, const MateStepAux* aux, const float* rewards, const float* soft_matrix,
                                         float* cam_terms, float* tgt_terms, void* stream) {
    if (!sim || !aux || !rewards || !tgt_terms || (sim->cfg.num_cameras > 0 && !cam_terms)) return fail(MATE_EINVAL, "null argument");
    if (!aux->coverage || !aux->target_dones || !aux->is_colliding || !aux->warehouse_dist || !aux->tgt_goal || !aux->tgt_empty_bits ||
        (sim->cfg.num_cameras > 0 && (!aux->mask_ct || !aux->mask_tc)))
        return fail(MATE_EINVAL, "auxiliary_terms needs mask_ct, mask_tc, coverage, target_dones, is_colliding, warehouse_dist, tgt_goal, tgt_empty_bits");
    CUDA_TRY(cudaSetDevice(sim->device));
    if (((uintptr_t)cam_terms & 15) || ((uintptr_t)tgt_terms & 15) || ((uintptr_t)aux->warehouse_dist & 15))
        return fail(MATE_EINVAL, "term tensors and warehouse_dist must be 16-byte aligned");
    const int threads = 128;
    const long long agents = (long long)sim->num_envs * (sim->cfg.num_cameras + sim->cfg.num_targets);
    aux_terms_kernel<<<(unsigned)((agents + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        *aux, rewards, soft_matrix, cam_terms, tgt_terms, sim->num_envs, sim->cfg.num_cameras, sim->cfg.num_targets);
    sim->launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("auxiliary terms launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int mate_b200_greedy_target_actions(MateSim* sim, double* memory, const uint8_t* reset_mask, double noise_scale,
                                               uint64_t seed, uint64_t serial, const MateAgentReplay* replay, float* tgt_act,
                                               void* stream) {
    if (!sim || !memory || !tgt_act) return fail(MATE_EINVAL, "null argument");
    if (((uintptr_t)tgt_act & 7) || ((uintptr_t)memory & 7)) return fail(MATE_EINVAL, "agent buffers must be 8-byte aligned");
    CUDA_TRY(cudaSetDevice(sim->device));
    MateAgentReplay r{};
    if (replay) r = *replay;
    const int threads = 128;
    greedy_target_kernel<<<(sim->num_envs + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
        sim->base, sim->cfg.num_targets, memory, reset_mask, noise_scale, seed, serial, r, tgt_act);
    sim->launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("greedy target launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int mate_b200_greedy_camera_actions(MateSim* sim, double* memory, const uint8_t* tracked, const uint8_t* reset_mask,
                                               uint64_t seed, uint64_t serial, const MateCameraAgentReplay* replay, float* cam_act,
                                               void* stream) {
    if (!sim || !memory || !tracked || !cam_act) return fail(MATE_EINVAL, "null argument");
    if (sim->cfg.num_cameras == 0) return fail(MATE_EINVAL, "the configuration has no cameras");
    if (((uintptr_t)cam_act & 7) || ((uintptr_t)memory & 7)) return fail(MATE_EINVAL, "agent buffers must be 8-byte aligned");
    CUDA_TRY(cudaSetDevice(sim->device));
    MateCameraAgentReplay r{};
    if (replay) r = *replay;
    // one thread per camera agent; a warp's 32 agent memories are staged in shared memory
    const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets;
    if (32 % nc != 0) return fail(MATE_EINVAL, "the batched GreedyCameraAgent needs a number of cameras that divides 32");
    const long long agents = (long long)sim->num_envs * nc;
    greedy_camera_kernel<<<(unsigned)((agents + kCameraAgentThreads - 1) / kCameraAgentThreads), kCameraAgentThreads, 0, (cudaStream_t)stream>>>(
        sim->base, nc, nt, memory, tracked, reset_mask, seed, serial, r, cam_act);
    sim->launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("greedy camera launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int mate_b200_decode_actions(const int64_t* index, const float* table, int32_t table_size, float* out,
                                        int64_t count, void* stream) {
    if (!index || !table || !out || table_size <= 0 || count < 0) return fail(MATE_EINVAL, "bad argument");
    if (count == 0) return MATE_OK;
    if (((uintptr_t)out & 7) || ((uintptr_t)table & 7)) return fail(MATE_EINVAL, "action buffers must be 8-byte aligned");
    decode_actions_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(index, table, table_size, out, count);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("decode launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

// Launch the fused kernel over envs [begin, begin + count); I/O pointers are for env 0.
static int launch_range(MateSim* sim, Params p, int begin, int count, cudaStream_t stream) {
    const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets;
    const int dc = sim->kernel.dc, dt = sim->kernel.dt;
    // SoA rows are indexed [row * bpad + env]: shifting the base pointers selects the sub-range
    p.cam_x += begin; p.cam_y += begin; p.cam_phi += begin; p.cam_theta += begin;
    p.tgt_x += begin; p.tgt_y += begin; p.obs_x += begin; p.obs_y += begin; p.obs_r += begin; p.obs_f4 += begin;
    p.tgt_pack += begin; p.cargo += begin; p.env_a += begin; p.env_b += begin; p.cc_clear += begin;
    if (p.cam_act) p.cam_act += (size_t)begin * nc * 2;
    if (p.tgt_act) p.tgt_act += (size_t)begin * nt * 2;
    if (p.cam_obs) p.cam_obs += (size_t)begin * nc * dc;
    if (p.tgt_obs) p.tgt_obs += (size_t)begin * nt * dt;
    if (p.rewards) p.rewards += (size_t)begin * 2;
    if (p.done) p.done += begin;
    if (p.env_mask) p.env_mask += begin;
    if (p.replay_transmit) p.replay_transmit += (size_t)begin * nc * nt;
    if (p.replay_choice) p.replay_choice += (size_t)begin * nt;
    if (p.ready) p.ready += begin;
    p.next_offset = begin;   // `next` addresses whole-batch arrays
    p.env_index_base += begin;
    p.num_envs = count;
    const int per_cta = sim->kernel.envs_per_cta / 32 * sim->tile_envs;
    const int grid = (count + per_cta - 1) / per_cta;
    sim->kernel.launch(p, grid, stream);
    sim->launches += 1;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

static int check_obs_alignment(const void* cam_obs, const void* tgt_obs) {
    if (((uintptr_t)cam_obs & 15) || ((uintptr_t)tgt_obs & 15)) return fail(MATE_EINVAL, "observation buffers must be 16-byte aligned");
    return MATE_OK;
}

// the view masks / per-target flags are written with 4-, 8- and 16-byte stores
static int check_aux_alignment(const MateStepAux* aux) {
    if (!aux) return MATE_OK;
    const void* ptrs[] = {aux->mask_ct, aux->mask_cc, aux->mask_co, aux->mask_tc, aux->mask_to, aux->mask_tt, aux->target_dones,
                          aux->is_colliding, aux->warehouse_dist};
    for (const void* ptr : ptrs)
        if ((uintptr_t)ptr & 15) return fail(MATE_EINVAL, "auxiliary output buffers must be 16-byte aligned");
    return MATE_OK;
}

static void fill_aux(Params& p, const MateStepAux* aux, const MateReplay* replay) {
    if (aux) {
        p.aux = *aux; p.has_aux = 1;
        p.has_aux_detail = aux->mask_ct || aux->mask_cc || aux->mask_co || aux->mask_tc || aux->mask_to || aux->mask_tt ||
                           aux->target_dones || aux->is_colliding || aux->warehouse_dist || aux->tgt_goal || aux->tgt_empty_bits;
    } else {
        memset(&p.aux, 0, sizeof(p.aux)); p.has_aux = 0; p.has_aux_detail = 0;
    }
    p.replay_transmit = replay ? replay->transmit : nullptr;
    p.replay_choice = replay ? replay->goal_choice : nullptr;
}

// Launch MODE_PREPARE over the second state block: re-initialises the prepared next episode of every env
// whose tag is stale.  Off the step path: on the side stream, ordered after the work already enqueued on
// `stream` (refill_mode 2: on `stream` itself, used by the tests to make the prepared path deterministic).
static int launch_prepare(MateSim* sim, cudaStream_t stream) {
    if (!sim->refill_mode) return MATE_OK;
    cudaStream_t target = stream;
    if (sim->refill_mode == 1) {
        CUDA_TRY(cudaEventRecord(sim->side_event, stream));
        CUDA_TRY(cudaStreamWaitEvent(sim->side, sim->side_event, 0));
        target = sim->side;
    }
    Params n = sim->next_base;
    n.mode = MODE_PREPARE; n.flags = 0; n.seed = sim->seed; n.l2_window_bytes = 0;
    n.cam_act = n.tgt_act = nullptr; n.cam_obs = n.tgt_obs = n.rewards = nullptr; n.done = nullptr; n.env_mask = nullptr;
    fill_aux(n, nullptr, nullptr);
    const int per_cta = sim->kernel.envs_per_cta / 32 * sim->tile_envs;
    const int grid = (sim->num_envs + per_cta - 1) / per_cta;
    sim->kernel.launch(n, grid, target);
    sim->launches += 1;
    sim->steps_since_refill = 0;
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("prepare launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

// the prepared episodes no longer match the live state (explicit reset with a new seed, set_state)
static int invalidate_prepared(MateSim* sim, cudaStream_t stream) {
    if (!sim->refill_mode) return MATE_OK;
    CUDA_TRY(cudaStreamSynchronize(sim->side));
    CUDA_TRY(cudaMemsetAsync(sim->base.ready, 0, sizeof(uint32_t) * sim->bpad, stream));
    return MATE_OK;
}

__global__ void rewind_episode_kernel(int4* __restrict__ env_b, const uint8_t* __restrict__ env_mask, int num_envs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < num_envs && (env_mask == nullptr || env_mask[e] != 0)) env_b[e].w = 0;
}

extern "C" int mate_b200_seed(MateSim* sim, const uint8_t* env_mask, uint64_t seed, void* stream) {
    if (!sim) return fail(MATE_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(sim->device));
    sim->seed = seed;
    if (int rc = invalidate_prepared(sim, (cudaStream_t)stream)) return rc;
    rewind_episode_kernel<<<(sim->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sim->base.env_b, env_mask, sim->num_envs);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(MATE_ECUDA, std::string("seed launch: ") + cudaGetErrorString(err));
    return MATE_OK;
}

extern "C" int mate_b200_reset(MateSim* sim, const uint8_t* env_mask, uint64_t seed, float* cam_obs,
                               float* tgt_obs, void* stream) {
    if (!sim || !tgt_obs || (sim->cfg.num_cameras > 0 && !cam_obs)) return fail(MATE_EINVAL, "null argument");
    if (int rc = check_obs_alignment(cam_obs, tgt_obs)) return rc;
    CUDA_TRY(cudaSetDevice(sim->device));
    sim->seed = seed;
    Params p = sim->base;
    p.mode = MODE_RESET; p.flags = 0; p.seed = seed;
    p.cam_obs = cam_obs; p.tgt_obs = tgt_obs; p.env_mask = env_mask;
    fill_aux(p, nullptr, nullptr);
    if (int rc = invalidate_prepared(sim, (cudaStream_t)stream)) return rc;
    if (int rc = launch_range(sim, p, 0, sim->num_envs, (cudaStream_t)stream)) return rc;
    return launch_prepare(sim, (cudaStream_t)stream);
}

extern "C" int mate_b200_step(MateSim* sim, const float* cam_act, const float* tgt_act, float* cam_obs,
                              float* tgt_obs, float* rewards, uint8_t* done, const MateStepAux* aux,
                              const MateReplay* replay, uint32_t flags, void* stream) {
    if (!sim || !tgt_act || !tgt_obs || !rewards || !done) return fail(MATE_EINVAL, "null argument");
    if (sim->cfg.num_cameras > 0 && (!cam_act || !cam_obs)) return fail(MATE_EINVAL, "null camera buffers");
    if (int rc = check_obs_alignment(cam_obs, tgt_obs)) return rc;
    if (((uintptr_t)cam_act & 7) || ((uintptr_t)tgt_act & 7) || ((uintptr_t)rewards & 7)) return fail(MATE_EINVAL, "action/reward buffers must be 8-byte aligned");
    if (int rc = check_aux_alignment(aux)) return rc;
    CUDA_TRY(cudaSetDevice(sim->device));
    Params p = sim->base;
    p.mode = MODE_STEP; p.flags = flags; p.seed = sim->seed;
    p.cam_act = cam_act; p.tgt_act = tgt_act; p.cam_obs = cam_obs; p.tgt_obs = tgt_obs; p.rewards = rewards; p.done = done;
    fill_aux(p, aux, replay);
    if (int rc = launch_range(sim, p, 0, sim->num_envs, (cudaStream_t)stream)) return rc;
    if ((flags & MATE_STEP_AUTO_RESET) && sim->refill_mode &&
        (sim->refill_mode == 2 || ++sim->steps_since_refill >= sim->refill_period))
        return launch_prepare(sim, (cudaStream_t)stream);
    return MATE_OK;
}

extern "C" int mate_b200_observe(MateSim* sim, float* cam_obs, float* tgt_obs, const MateStepAux* aux,
                                 const MateReplay* replay, void* stream) {
    if (!sim || !tgt_obs || (sim->cfg.num_cameras > 0 && !cam_obs)) return fail(MATE_EINVAL, "null argument");
    if (int rc = check_obs_alignment(cam_obs, tgt_obs)) return rc;
    if (int rc = check_aux_alignment(aux)) return rc;
    CUDA_TRY(cudaSetDevice(sim->device));
    Params p = sim->base;
    p.mode = MODE_OBSERVE; p.flags = 0; p.seed = sim->seed;
    p.cam_obs = cam_obs; p.tgt_obs = tgt_obs;
    fill_aux(p, aux, replay);
    return launch_range(sim, p, 0, sim->num_envs, (cudaStream_t)stream);
}

// ---- host-buffer path: chunked H2D -> kernel -> D2H over a few streams ------------------------
static int ensure_host_path(MateSim* sim) {
    if (sim->host_ready) return MATE_OK;
    const size_t B = sim->num_envs;
    const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets;
    for (int i = 0; i < MateSim::kHostStreams; ++i) {
        CUDA_TRY(cudaStreamCreateWithFlags(&sim->hstreams[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&sim->hevents[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMalloc(&sim->h_cam_act, sizeof(float) * (B * nc * 2 + 4)));
    CUDA_TRY(cudaMalloc(&sim->h_tgt_act, sizeof(float) * B * nt * 2));
    CUDA_TRY(cudaMalloc(&sim->h_cam_obs, sizeof(float) * (B * nc * sim->kernel.dc + 4)));
    CUDA_TRY(cudaMalloc(&sim->h_tgt_obs, sizeof(float) * B * nt * sim->kernel.dt));
    CUDA_TRY(cudaMalloc(&sim->h_rewards, sizeof(float) * B * 2));
    CUDA_TRY(cudaMalloc(&sim->h_done, B));
    // Compacted device -> host leg (mate_hostpath.cuh): what bounds it is the number of host threads that rebuild the dense
    // rows (measured with 15: 9.3 - 9.6e6 env-steps/s against 8.6 - 8.7e6 for the dense copy, with 10: 8.3e6, with 6: 6.4e6,
    // profiles/r2v_hostpath.md), so it is used when this process has at least 14 of them to itself -- all host threads the
    // process may run on but the caller's, divided among the processes that share the host (one per GPU) -- and the rows
    // are whole 16-byte chunks.  MATE_B200_HOST_COMPACT=0 / 1 forces the choice, MATE_B200_HOST_THREADS the pool size.
    int threads = (int)std::thread::hardware_concurrency();
    {
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) threads = std::min(threads > 0 ? threads : CPU_COUNT(&set), CPU_COUNT(&set));
    }
    int sharers = 1;
    if (const char* v = getenv("LOCAL_WORLD_SIZE")) sharers = std::max(1, atoi(v));
    threads = std::max(1, threads / sharers - 1);
    if (const char* v = getenv("MATE_B200_HOST_THREADS")) threads = std::max(1, atoi(v));
    threads = std::min(threads, 64);
    // 1: every call takes the compacted leg (rebuilding all rows needs >= 14 threads to beat the dense copy); 2: only the calls
    // that patch changes into kept rows do (5 threads are enough for those: 1.02e7 against 8.5e6; 3 when four or more
    // processes share the host, whose memory system then bounds the dense copies: 8 x B200 1.74e7 against 1.49e7), the others
    // copy densely
    sim->compact_mode = threads >= 14 ? 1 : ((threads >= 5 || (threads >= 3 && sharers >= 4)) ? 2 : 0);
    if (const char* v = getenv("MATE_B200_HOST_COMPACT")) sim->compact_mode = v[0] == '1' ? 1 : (v[0] == '2' ? 2 : 0);
    if ((sizeof(float) * nc * sim->kernel.dc) % 16 || (sizeof(float) * nt * sim->kernel.dt) % 16) sim->compact_mode = 0;
    if (sim->compact_mode) {
        const size_t region_bytes[2] = {sizeof(float) * B * nc * sim->kernel.dc, sizeof(float) * B * nt * sim->kernel.dt};
        bool ok = true;
        for (int r = 0; r < 2 && ok; ++r) {
            if (region_bytes[r] == 0) continue;
            const size_t chunks = region_bytes[r] / 16 + 8, blocks = chunks / kCompactBlock + MateSim::kMaxHostChunks + 1;
            ok = cudaMalloc(&sim->d_compact[r], chunks * 16) == cudaSuccess && cudaMalloc(&sim->d_table[r], blocks * sizeof(CompactEntry)) == cudaSuccess &&
                 cudaMallocHost(&sim->p_compact[r], chunks * 16) == cudaSuccess && cudaMallocHost(&sim->p_table[r], blocks * sizeof(CompactEntry)) == cudaSuccess &&
                 cudaMalloc(&sim->h_rows2[r], region_bytes[r] + 16) == cudaSuccess;
        }
        ok = ok && cudaMalloc(&sim->d_count, sizeof(unsigned int) * MateSim::kMaxHostChunks * 2) == cudaSuccess &&
             cudaMallocHost(&sim->p_count, sizeof(unsigned int) * MateSim::kMaxHostChunks * 2) == cudaSuccess &&
             cudaHostGetDevicePointer(reinterpret_cast<void**>(&sim->p_count_dev), sim->p_count, 0) == cudaSuccess &&
             cudaStreamCreateWithFlags(&sim->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < MateSim::kMaxHostChunks && ok; ++i)
            ok = cudaEventCreateWithFlags(&sim->e_tables[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sim->e_counts[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sim->e_stream[i], cudaEventDisableTiming) == cudaSuccess;
        if (ok) sim->pool.reset(new ExpandPool(threads, sim->device));
        else {   // not enough pinned host memory or device memory for the second copy of the rows: the dense leg needs neither
            cudaGetLastError();
            sim->compact_mode = 0;
            for (int r = 0; r < 2; ++r) { cudaFree(sim->h_rows2[r]); sim->h_rows2[r] = nullptr; }
        }
    }
    sim->host_ready = true;
    return MATE_OK;
}

extern "C" int mate_b200_step_host(MateSim* sim, const float* cam_act, const float* tgt_act, float* cam_obs,
                                   float* tgt_obs, float* rewards, uint8_t* done, uint32_t flags) {
    if (!sim || !tgt_act || !tgt_obs || !rewards || !done) return fail(MATE_EINVAL, "null argument");
    const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets;
    if (nc > 0 && (!cam_act || !cam_obs)) return fail(MATE_EINVAL, "null camera buffers");
    CUDA_TRY(cudaSetDevice(sim->device));
    if (int rc = ensure_host_path(sim)) return rc;
    const int dc = sim->kernel.dc, dt = sim->kernel.dt;
    const int B = sim->num_envs;
    // chunk size: a multiple of the CTA tile, ~16 chunks so copies and kernels overlap
    int chunk = (B + 15) / 16;
    const int tile = sim->kernel.envs_per_cta * 8;
    chunk = (int)align_up((size_t)std::max(chunk, 1), (size_t)tile);
    Params p = sim->base;
    p.mode = MODE_STEP; p.flags = flags; p.seed = sim->seed;
    // the rows of this call and of the previous one take turns in two pairs of device buffers (compacted leg only)
    const bool two_buffers = sim->compact_mode && sim->h_rows2[1] != nullptr && (nc == 0 || sim->h_rows2[0] != nullptr);
    float* const rows_now[2] = {two_buffers && sim->row_parity ? sim->h_rows2[0] : sim->h_cam_obs, two_buffers && sim->row_parity ? sim->h_rows2[1] : sim->h_tgt_obs};
    float* const rows_before[2] = {two_buffers && !sim->row_parity ? sim->h_rows2[0] : sim->h_cam_obs, two_buffers && !sim->row_parity ? sim->h_rows2[1] : sim->h_tgt_obs};
    p.cam_act = sim->h_cam_act; p.tgt_act = sim->h_tgt_act; p.cam_obs = rows_now[0]; p.tgt_obs = rows_now[1];
    p.rewards = sim->h_rewards; p.done = sim->h_done;
    fill_aux(p, nullptr, nullptr);
    // work queued earlier on the legacy default stream (torch's default) or on blocking streams is finished first;
    // callers that use non-blocking streams of their own synchronise them before this call (include/mate_b200.h)
    CUDA_TRY(cudaEventRecord(sim->hevents[0], nullptr));
    for (int i = 0; i < MateSim::kHostStreams; ++i) CUDA_TRY(cudaStreamWaitEvent(sim->hstreams[i], sim->hevents[0], 0));
    // ---- compacted device -> host leg (mate_hostpath.cuh): every launch chunk's rows are compacted on the device, the
    //      compact streams cross the link one after the other, host threads expand them into the caller's buffers
    const int num_chunks = (B + chunk - 1) / chunk;
    // MATE_STEP_HOST_ROWS_KEPT: the caller's row buffers still hold the rows of the previous call (same buffers, not
    // modified since) and the device still holds them too: only the 64-byte groups that differ cross the link and are
    // rewritten.  Otherwise the all-zero chunks are dropped and every byte of the caller's buffers is written.
    const bool delta = two_buffers && (flags & MATE_STEP_HOST_ROWS_KEPT) && sim->kept_valid && sim->kept_rows[0] == cam_obs && sim->kept_rows[1] == tgt_obs;
    const bool compact = (sim->compact_mode == 1 || (sim->compact_mode == 2 && delta)) && num_chunks <= MateSim::kMaxHostChunks && B % 4 == 0 &&
                         !((uintptr_t)cam_obs & 3) && !((uintptr_t)tgt_obs & 3);
    sim->kept_valid = false;
    if (compact) {
        static const bool trace = getenv("MATE_B200_HOST_TRACE") != nullptr;
        const auto t_start = std::chrono::steady_clock::now();
        auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
        double t_enq = 0, t_counts = 0, t_last_stream = 0;
        const size_t per_env[2] = {(size_t)nc * dc * 4, (size_t)nt * dt * 4};   // bytes of rows per environment and region
        float* const host_rows[2] = {cam_obs, tgt_obs};
        float* const dev_rows[2] = {rows_now[0], rows_now[1]};
        std::vector<long long> table_base(2 * (num_chunks + 1), 0);
        CUDA_TRY(cudaMemsetAsync(sim->d_count, 0, sizeof(unsigned int) * MateSim::kMaxHostChunks * 2, sim->hstreams[0]));
        CUDA_TRY(cudaEventRecord(sim->hevents[1], sim->hstreams[0]));
        for (int i = 1; i < MateSim::kHostStreams; ++i) CUDA_TRY(cudaStreamWaitEvent(sim->hstreams[i], sim->hevents[1], 0));
        int kc = 0;
        for (int begin = 0; begin < B; begin += chunk, ++kc) {
            const int count = std::min(chunk, B - begin);
            cudaStream_t s = sim->hstreams[kc % MateSim::kHostStreams];
            if (nc) CUDA_TRY(cudaMemcpyAsync(sim->h_cam_act + (size_t)begin * nc * 2, cam_act + (size_t)begin * nc * 2, sizeof(float) * count * nc * 2, cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(sim->h_tgt_act + (size_t)begin * nt * 2, tgt_act + (size_t)begin * nt * 2, sizeof(float) * count * nt * 2, cudaMemcpyHostToDevice, s));
            if (int rc = launch_range(sim, p, begin, count, s)) return rc;
            for (int r = 0; r < 2; ++r) {
                table_base[2 * (kc + 1) + r] = table_base[2 * kc + r];
                if (per_env[r] == 0) continue;
                const size_t first = (size_t)begin * per_env[r] / 16;            // B % 4 == 0 and chunk % 4 == 0: whole chunks
                const long long nchunks = (long long)((size_t)count * per_env[r] / 16);
                const long long nblocks = (nchunks + kCompactBlock - 1) / kCompactBlock;
                CompactEntry* table = sim->d_table[r] + table_base[2 * kc + r];
                table_base[2 * (kc + 1) + r] += nblocks;
                if (delta) compact_changes_kernel<<<1184, 128, 0, s>>>(reinterpret_cast<const uint4*>(dev_rows[r]) + first, reinterpret_cast<const uint4*>(rows_before[r]) + first,
                                                                       nchunks, sim->d_compact[r] + first, table, sim->d_count + 2 * kc + r);
                else compact_chunks_kernel<<<1184, 128, 0, s>>>(reinterpret_cast<const uint4*>(dev_rows[r]) + first, nchunks,
                                                                sim->d_compact[r] + first, table, sim->d_count + 2 * kc + r);
            }
            sim->launches += (nc ? 2 : 1);
            // The sizes of the chunk's compact streams go to the host by a store into pinned memory, not by a copy: a copy
            // would queue behind the compact streams of earlier chunks on the copy engine, and the next stream could only be
            // enqueued after the previous one had crossed the link (a bubble per chunk).
            publish_counts_kernel<<<1, 32, 0, s>>>(sim->d_count + 2 * kc, sim->p_count_dev + 2 * kc);
            CUDA_TRY(cudaEventRecord(sim->e_counts[kc], s));
            for (int r = 0; r < 2; ++r) {
                const long long nblocks = table_base[2 * (kc + 1) + r] - table_base[2 * kc + r];
                if (nblocks > 0) CUDA_TRY(cudaMemcpyAsync(sim->p_table[r] + table_base[2 * kc + r], sim->d_table[r] + table_base[2 * kc + r], sizeof(CompactEntry) * nblocks, cudaMemcpyDeviceToHost, s));
            }
            CUDA_TRY(cudaMemcpyAsync(rewards + (size_t)begin * 2, sim->h_rewards + (size_t)begin * 2, sizeof(float) * count * 2, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaMemcpyAsync(done + begin, sim->h_done + begin, count, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaEventRecord(sim->e_tables[kc], s));
        }
        t_enq = since();
        // the sizes are known once a chunk's kernels have run: its compact streams go on the copy stream in chunk order
        const int pieces = std::max(1, sim->pool->size());
        kc = 0;
        for (int begin = 0; begin < B; begin += chunk, ++kc) {
            CUDA_TRY(cudaEventSynchronize(sim->e_counts[kc]));
            CUDA_TRY(cudaStreamWaitEvent(sim->copy_stream, sim->e_tables[kc], 0));   // a chunk's stream event then covers its tables too
            for (int r = 0; r < 2; ++r) {
                if (per_env[r] == 0 || sim->p_count[2 * kc + r] == 0) continue;
                const size_t first = (size_t)begin * per_env[r] / 16;
                CUDA_TRY(cudaMemcpyAsync(sim->p_compact[r] + first, sim->d_compact[r] + first, (size_t)sim->p_count[2 * kc + r] * 16,
                                         cudaMemcpyDeviceToHost, sim->copy_stream));
            }
            CUDA_TRY(cudaEventRecord(sim->e_stream[kc], sim->copy_stream));
            // expansion: the pool's threads wait for the chunk's stream themselves, so the caller's thread goes on to the next sizes
            const int count = std::min(chunk, B - begin);
            for (int r = 0; r < 2; ++r) {
                if (per_env[r] == 0) continue;
                const size_t first = (size_t)begin * per_env[r] / 16;
                const long long nchunks = (long long)((size_t)count * per_env[r] / 16);
                const long long nblocks = (nchunks + kCompactBlock - 1) / kCompactBlock;
                const long long per_piece = (nblocks + pieces - 1) / pieces;
                __m128i* dst = reinterpret_cast<__m128i*>(host_rows[r]) + first;
                const bool aligned = (((uintptr_t)dst) & 15) == 0;
                for (long long b0 = 0; b0 < nblocks; b0 += per_piece)
                    sim->pool->submit(ExpandPool::Work{sim->p_table[r] + table_base[2 * kc + r], delta, reinterpret_cast<const __m128i*>(sim->p_compact[r] + first),
                                                       dst, nchunks, b0, std::min(nblocks, b0 + per_piece), aligned, sim->e_stream[kc]});
            }
        }
        t_counts = since();
        CUDA_TRY(cudaStreamSynchronize(sim->copy_stream));
        t_last_stream = since();
        sim->pool->wait();
        for (int i = 0; i < MateSim::kHostStreams; ++i) CUDA_TRY(cudaStreamSynchronize(sim->hstreams[i]));
        {
            unsigned long long link = 0;
            for (int i = 0; i < 2 * num_chunks; ++i) link += (unsigned long long)sim->p_count[i] * 16ull;
            link += (unsigned long long)(table_base[2 * num_chunks] + table_base[2 * num_chunks + 1]) * sizeof(CompactEntry);
            sim->last_leg = delta ? 2 : 1; sim->last_link_bytes = link;
        }
        // the caller's buffers and the device now hold this call's rows
        sim->kept_rows[0] = cam_obs; sim->kept_rows[1] = tgt_obs; sim->kept_valid = true;
        if (two_buffers) sim->row_parity ^= 1;
        if (trace) {
            unsigned long long kept = 0;
            for (int i = 0; i < 2 * num_chunks; ++i) kept += sim->p_count[i];
            fprintf(stderr, "step_host compact: enqueued %.2f ms, sizes known + copies enqueued %.2f, last stream arrived %.2f, expanded %.2f; kept %.1f MB of %.1f MB, %d threads, changes only %d\n",
                    t_enq, t_counts, t_last_stream, since(), kept * 16 / 1e6, (per_env[0] + per_env[1]) * (double)B / 1e6, sim->pool->size(), (int)delta);
        }
        if ((flags & MATE_STEP_AUTO_RESET) && sim->refill_mode &&
            (sim->refill_mode == 2 || ++sim->steps_since_refill >= sim->refill_period))
            return launch_prepare(sim, nullptr);
        return MATE_OK;
    }
    int k = 0;
    for (int begin = 0; begin < B; begin += chunk, ++k) {
        const int count = std::min(chunk, B - begin);
        cudaStream_t s = sim->hstreams[k % MateSim::kHostStreams];
        if (nc) CUDA_TRY(cudaMemcpyAsync(sim->h_cam_act + (size_t)begin * nc * 2, cam_act + (size_t)begin * nc * 2, sizeof(float) * count * nc * 2, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(sim->h_tgt_act + (size_t)begin * nt * 2, tgt_act + (size_t)begin * nt * 2, sizeof(float) * count * nt * 2, cudaMemcpyHostToDevice, s));
        if (int rc = launch_range(sim, p, begin, count, s)) return rc;
        if (nc) CUDA_TRY(cudaMemcpyAsync(cam_obs + (size_t)begin * nc * dc, rows_now[0] + (size_t)begin * nc * dc, sizeof(float) * count * nc * dc, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(tgt_obs + (size_t)begin * nt * dt, rows_now[1] + (size_t)begin * nt * dt, sizeof(float) * count * nt * dt, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(rewards + (size_t)begin * 2, sim->h_rewards + (size_t)begin * 2, sizeof(float) * count * 2, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(done + begin, sim->h_done + begin, count, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < MateSim::kHostStreams; ++i) CUDA_TRY(cudaStreamSynchronize(sim->hstreams[i]));
    sim->last_leg = 0; sim->last_link_bytes = (unsigned long long)B * ((size_t)nc * dc + (size_t)nt * dt) * sizeof(float);
    if (two_buffers) {   // the caller's buffers and the device hold this call's rows: the next call may send changes only
        sim->kept_rows[0] = cam_obs; sim->kept_rows[1] = tgt_obs; sim->kept_valid = true; sim->row_parity ^= 1;
    }
    // the prepared episodes are refilled on the side stream like on the device path
    if ((flags & MATE_STEP_AUTO_RESET) && sim->refill_mode &&
        (sim->refill_mode == 2 || ++sim->steps_since_refill >= sim->refill_period))
        return launch_prepare(sim, nullptr);
    return MATE_OK;
}

// ---- state get / set (host arrays, synchronous) --------------------------------------------------
template <typename T>
static int download(std::vector<T>& host, const T* dev, size_t n) {
    host.resize(n);
    CUDA_TRY(cudaMemcpy(host.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
    return MATE_OK;
}
template <typename T>
static int upload(const std::vector<T>& host, T* dev) {
    CUDA_TRY(cudaMemcpy(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MATE_OK;
}

extern "C" int mate_b200_get_state(MateSim* sim, MateStateView* v) {
    if (!sim || !v) return fail(MATE_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(sim->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets, no = sim->cfg.num_obstacles;
    const size_t bp = sim->bpad, B = sim->num_envs;
    const Params& p = sim->base;
    std::vector<double> a, b2, c, d;
    if (nc && (v->cam_xy || v->cam_phi || v->cam_theta)) {
        if (download(a, p.cam_x, nc * bp) || download(b2, p.cam_y, nc * bp) || download(c, p.cam_phi, nc * bp) || download(d, p.cam_theta, nc * bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < nc; ++k) {
                if (v->cam_xy) { v->cam_xy[(e * nc + k) * 2] = a[k * bp + e]; v->cam_xy[(e * nc + k) * 2 + 1] = b2[k * bp + e]; }
                if (v->cam_phi) v->cam_phi[e * nc + k] = c[k * bp + e];
                if (v->cam_theta) v->cam_theta[e * nc + k] = d[k * bp + e];
            }
    }
    if (v->tgt_xy) {
        if (download(a, p.tgt_x, nt * bp) || download(b2, p.tgt_y, nt * bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < nt; ++k) { v->tgt_xy[(e * nt + k) * 2] = a[k * bp + e]; v->tgt_xy[(e * nt + k) * 2 + 1] = b2[k * bp + e]; }
    }
    if (no && v->obs_xyr) {
        if (download(a, p.obs_x, no * bp) || download(b2, p.obs_y, no * bp) || download(c, p.obs_r, no * bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < no; ++k) { v->obs_xyr[(e * no + k) * 3] = a[k * bp + e]; v->obs_xyr[(e * no + k) * 3 + 1] = b2[k * bp + e]; v->obs_xyr[(e * no + k) * 3 + 2] = c[k * bp + e]; }
    }
    {
        std::vector<uint32_t> pack;
        if (download(pack, p.tgt_pack, nt * bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < nt; ++k) {
                const uint32_t q = pack[k * bp + e];
                if (v->tgt_capacity) v->tgt_capacity[e * nt + k] = tp_capacity(q);
                if (v->tgt_goal) v->tgt_goal[e * nt + k] = tp_goal(q);
                if (v->tgt_weight) v->tgt_weight[e * nt + k] = tp_weight(q);
                if (v->tgt_bounty) v->tgt_bounty[e * nt + k] = tp_bounty(q);
                if (v->tgt_empty_bits) v->tgt_empty_bits[e * nt + k] = tp_empty(q);
            }
    }
    {
        std::vector<uint4> cargo, ea;
        std::vector<int4> eb;
        if (download(cargo, p.cargo, 2 * bp) || download(ea, p.env_a, bp) || download(eb, p.env_b, bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e) {
            const uint32_t words[8] = {cargo[e].x, cargo[e].y, cargo[e].z, cargo[e].w, cargo[bp + e].x, cargo[bp + e].y, cargo[bp + e].z, cargo[bp + e].w};
            if (v->remaining) for (int i = 0; i < 16; ++i) v->remaining[e * 16 + i] = (int32_t)((words[i >> 1] >> ((i & 1) * 16)) & 0xFFFF);
            if (v->awaiting) { v->awaiting[e * 4 + 0] = ea[e].x & 0xFFFF; v->awaiting[e * 4 + 1] = ea[e].x >> 16; v->awaiting[e * 4 + 2] = ea[e].y & 0xFFFF; v->awaiting[e * 4 + 3] = ea[e].y >> 16; }
            if (v->episode_step) v->episode_step[e] = (int32_t)ea[e].z;
            if (v->num_delivered) v->num_delivered[e] = (int32_t)ea[e].w;
            if (v->episode_reward) { v->episode_reward[e * 2] = (double)eb[e].x; v->episode_reward[e * 2 + 1] = (double)eb[e].y; }
            if (v->episode_id) v->episode_id[e] = eb[e].w;
        }
    }
    return MATE_OK;
}

extern "C" int mate_b200_set_state(MateSim* sim, const MateStateView* v) {
    if (!sim || !v) return fail(MATE_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(sim->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (sim->refill_mode) CUDA_TRY(cudaMemset(sim->base.ready, 0, sizeof(uint32_t) * sim->bpad));   // prepared episodes may no longer match
    const int nc = sim->cfg.num_cameras, nt = sim->cfg.num_targets, no = sim->cfg.num_obstacles;
    const size_t bp = sim->bpad, B = sim->num_envs;
    const Params& p = sim->base;
    std::vector<double> a, b2, c, d;
    if (nc) {
        if (download(a, p.cam_x, nc * bp) || download(b2, p.cam_y, nc * bp) || download(c, p.cam_phi, nc * bp) || download(d, p.cam_theta, nc * bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < nc; ++k) {
                if (v->cam_xy) { a[k * bp + e] = v->cam_xy[(e * nc + k) * 2]; b2[k * bp + e] = v->cam_xy[(e * nc + k) * 2 + 1]; }
                if (v->cam_phi) c[k * bp + e] = v->cam_phi[e * nc + k];
                if (v->cam_theta) {
                    const double th = v->cam_theta[e * nc + k];
                    if (!(th > 0.0) || th > 180.0) return fail(MATE_EINVAL, "cam_theta out of (0, 180]");
                    d[k * bp + e] = th;
                }
            }
        if (upload(a, p.cam_x) || upload(b2, p.cam_y) || upload(c, p.cam_phi) || upload(d, p.cam_theta)) return MATE_ECUDA;
    }
    if (v->cam_xy || v->obs_xyr)   // geometry changed: invalidate the per-episode camera line-of-sight cache
        CUDA_TRY(cudaMemset(p.cc_clear, 0, sizeof(unsigned long long) * bp));
    if (v->tgt_xy) {
        a.assign(nt * bp, 0.0); b2.assign(nt * bp, 0.0);
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < nt; ++k) { a[k * bp + e] = v->tgt_xy[(e * nt + k) * 2]; b2[k * bp + e] = v->tgt_xy[(e * nt + k) * 2 + 1]; }
        if (upload(a, p.tgt_x) || upload(b2, p.tgt_y)) return MATE_ECUDA;
    }
    if (no && v->obs_xyr) {
        a.assign(no * bp, 0.0); b2.assign(no * bp, 0.0); c.assign(no * bp, 0.0);
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < no; ++k) { a[k * bp + e] = v->obs_xyr[(e * no + k) * 3]; b2[k * bp + e] = v->obs_xyr[(e * no + k) * 3 + 1]; c[k * bp + e] = v->obs_xyr[(e * no + k) * 3 + 2]; }
        if (upload(a, p.obs_x) || upload(b2, p.obs_y) || upload(c, p.obs_r)) return MATE_ECUDA;
        std::vector<float4> f4(no * bp);   // fp32 shadow read by the prefilters and the observation packer
        for (size_t i = 0; i < f4.size(); ++i) f4[i] = make_float4((float)a[i], (float)b2[i], (float)c[i], 0.f);
        if (upload(f4, p.obs_f4)) return MATE_ECUDA;
    }
    {
        std::vector<uint32_t> pack;
        if (download(pack, p.tgt_pack, nt * bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e)
            for (int k = 0; k < nt; ++k) {
                const uint32_t q = pack[k * bp + e];
                int capacity = v->tgt_capacity ? v->tgt_capacity[e * nt + k] : tp_capacity(q);
                int goal = v->tgt_goal ? v->tgt_goal[e * nt + k] : tp_goal(q);
                int weight = v->tgt_weight ? v->tgt_weight[e * nt + k] : tp_weight(q);
                int bounty = v->tgt_bounty ? v->tgt_bounty[e * nt + k] : tp_bounty(q);
                int empty = v->tgt_empty_bits ? v->tgt_empty_bits[e * nt + k] : tp_empty(q);
                if (capacity < 1 || capacity > 2 || goal < -1 || goal > 3 || weight < 0 || weight > 3 || bounty < 0 || bounty > 65535)
                    return fail(MATE_EINVAL, "target integer state out of range");
                pack[k * bp + e] = pack_target(bounty, goal, weight, capacity, empty, tp_colliding(q));
            }
        if (upload(pack, p.tgt_pack)) return MATE_ECUDA;
    }
    {
        std::vector<uint4> cargo, ea;
        std::vector<int4> eb;
        if (download(cargo, p.cargo, 2 * bp) || download(ea, p.env_a, bp) || download(eb, p.env_b, bp)) return MATE_ECUDA;
        for (size_t e = 0; e < B; ++e) {
            if (v->remaining) {
                uint32_t words[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int i = 0; i < 16; ++i) words[i >> 1] |= ((uint32_t)v->remaining[e * 16 + i] & 0xFFFF) << ((i & 1) * 16);
                cargo[e] = make_uint4(words[0], words[1], words[2], words[3]);
                cargo[bp + e] = make_uint4(words[4], words[5], words[6], words[7]);
            }
            if (v->awaiting) {
                ea[e].x = ((uint32_t)v->awaiting[e * 4 + 0] & 0xFFFF) | ((uint32_t)v->awaiting[e * 4 + 1] << 16);
                ea[e].y = ((uint32_t)v->awaiting[e * 4 + 2] & 0xFFFF) | ((uint32_t)v->awaiting[e * 4 + 3] << 16);
            }
            if (v->episode_step) ea[e].z = (uint32_t)v->episode_step[e];
            if (v->num_delivered) ea[e].w = (uint32_t)v->num_delivered[e];
            if (v->episode_reward) { eb[e].x = (int)std::lrint(v->episode_reward[e * 2]); eb[e].y = (int)std::lrint(v->episode_reward[e * 2 + 1]); eb[e].z = 0; }
            if (v->episode_id) eb[e].w = v->episode_id[e];
        }
        if (upload(cargo, p.cargo) || upload(ea, p.env_a) || upload(eb, p.env_b)) return MATE_ECUDA;
    }
    return launch_prepare(sim, nullptr);   // prepare the next episodes of the new state right away
}

extern "C" int mate_b200_episode_stats(MateSim* sim, float* out16, int32_t reset_after, void* stream) {
    if (!sim || !out16) return fail(MATE_EINVAL, "null argument");
    CUDA_TRY(cudaSetDevice(sim->device));
    CUDA_TRY(cudaMemcpyAsync(out16, sim->base.stats, sizeof(float) * 16, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (reset_after) CUDA_TRY(cudaMemsetAsync(sim->base.stats, 0, sizeof(float) * 16, (cudaStream_t)stream));
    return MATE_OK;
}

#ifdef MATE_DEV_TIMELINE   // development builds only (profiles/tools/timeline.py): per-tile phase time stamps of the last step launch
extern "C" int mate_b200_debug_timeline(unsigned long long* out, int32_t count) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, mate::g_timeline, sizeof(unsigned long long) * (size_t)count));
    return MATE_OK;
}
#endif
