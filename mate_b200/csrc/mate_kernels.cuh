// mate_kernels.cuh -- fused per-step kernel of the B200-native MultiAgentTracking simulator.
//
// One launch per env.step: camera kinematics, target motion with disc/boundary collision,
// the five visibility masks (incl. the obstacle-occluded camera field of view), cargo
// pickup/delivery + rewards + done, auto-reset, and packing of all per-agent observation
// rows.  Reference behaviour being restated (paths relative to the reference root):
//   mate/environment.py:590-676 (step), :1271-1388 (_assign_goals/_simulate/_update_view),
//   :908-983 (joint_observation), :679-834 (reset); mate/entities.py:158-184 (obstruct),
//   :347-360 (Camera.simulate), :362-511 (FOV polyline + perceive), :645-668 (Target.simulate).
//
// Mapping (sm_100a, no tensor cores -- nothing here is a contraction):
//   * HBM state is struct-of-arrays [field][env]; decision state (positions, angles) is
//     fp64 so that the visibility predicates reproduce the float64 reference bit-for-bit
//     in practice; I/O (actions, observations, rewards) is fp32.
//   * A "group" of G = pow2 >= max(Nc, Nt) lanes owns one environment, 32/G environments
//     per warp; lane j of a group owns camera j, target j and obstacles j, j+G, ...  A warp
//     load of one SoA field therefore touches fully-used 32-byte sectors.
//   * Entity state is staged per environment in shared memory (fp64) so every lane can
//     loop over all cameras / obstacles with broadcast LDS; each entity-owning lane
//     computes the *column* of each mask (who sees my entity), which is exactly what the
//     observation packer needs, so no mask transposes.
//   * Observation rows are assembled in shared memory in the final [env][agent][feature]
//     layout and leave the SM as one contiguous bulk copy per warp and tensor
//     (cp.async.bulk shared->global, i.e. TMA), 16-byte aligned.
//   * The camera field of view is evaluated on the fly: instead of materialising the
//     reference's ~500-sample (phi, rho) polyline per camera at reset (35 kB/env), the
//     two polyline samples that bracket the query bearing are found analytically and
//     only those two rays are cast against the obstacle discs.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mate_b200.h"

#ifndef MATE_POOL_KB
#define MATE_POOL_KB 112   // shared-memory budget per CTA (KB): 112 -> 2 CTAs/SM, 74 -> 3 CTAs/SM
#endif
#ifndef MATE_MIN_CTAS
#define MATE_MIN_CTAS 2
#endif
#ifndef MATE_UNROLL_SENSE
#define MATE_UNROLL_SENSE 1
#endif
#ifndef MATE_UNROLL_PACK
#define MATE_UNROLL_PACK 1
#endif

#include "mate_common.cuh"

namespace mate {

constexpr int kUnrollSense = MATE_UNROLL_SENSE;   // tuning knobs: the kernel is instruction-fetch sensitive
constexpr int kUnrollPack = MATE_UNROLL_PACK;

// Static assignment of obstacles to the lanes of a group.  A warp executes every slot phase in
// lock step, so what matters is the NUMBER of slots (max obstacles per lane), not the per-lane
// total: plain round-robin, obstacle o -> lane o % G, slot o / G.
template <int NC, int NT, int NO, int G>
struct ObstacleMap {
    int owner[NO > 0 ? NO : 1] = {};
    int slot[NO > 0 ? NO : 1] = {};
    int max_count = 0;
    constexpr ObstacleMap() {
        for (int o = 0; o < NO; ++o) { owner[o] = o % G; slot[o] = o / G; }
        max_count = (NO + G - 1) / G;
    }
};

template <int NC, int NT, int NO>
struct Shape {
    static constexpr int NO_ = NO;
    static constexpr int MAXE = (NC > NT ? NC : NT) > 1 ? (NC > NT ? NC : NT) : 1;
    static constexpr int G = MAXE <= 1 ? 1 : MAXE <= 2 ? 2 : MAXE <= 4 ? 4 : MAXE <= 8 ? 8 : MAXE <= 16 ? 16 : 32;
    static constexpr int EPW = 32 / G;                       // environments per warp
    static constexpr ObstacleMap<NC, NT, NO, G> MAP{};
    static constexpr int OS = MAP.max_count;                 // obstacle slots per lane
    static constexpr int DC = 22 + 5 * NT + 4 * NO + 7 * NC; // mate/constants.py:267-282
    static constexpr int DT = 27 + 7 * NC + 4 * NO + 5 * NT; // mate/constants.py:285-300
    // shared-memory entity block per env (doubles): cams (x, y, phi, theta, rs, rs^2, cos phi, sin phi,
    // cos^2(theta/2)), tgts (x, y), obstacles (x, y, r)
    static constexpr int CAMF = 9;
    static constexpr int E_CAM = 0;
    static constexpr int E_TGT = E_CAM + CAMF * NC;
    static constexpr int E_OBS = E_TGT + 2 * NT;
    static constexpr int E_SCR = ((E_OBS + 3 * NO + 1) / 2) * 2;  // 16 x u32 scratch (reset results)
    // fp32 shadow of the entity block for the prefilters (floats, appended after the scratch):
    // cams {x, y, rs^2, cos phi, sin phi, cos^2(theta/2)}, tgts {x, y}, obstacles {x, y, r, -}
    static constexpr int FCAMF = 6;
    static constexpr int F_CAM = 0;
    static constexpr int F_TGT = F_CAM + FCAMF * NC;
    static constexpr int F_OBS = ((F_TGT + 2 * NT + 3) / 4) * 4;   // float4 aligned
    static constexpr int F_FLOATS = F_OBS + 4 * NO;
    static constexpr int E_F32 = E_SCR + 8;                  // (doubles) start of the fp32 block, 16-byte aligned
    static constexpr int E_RAW = E_F32 + (F_FLOATS + 1) / 2;
    // env stride in doubles: == 2 (mod 4) keeps every block 16-byte aligned (float4 shadow entries) and puts
    // the (up to 8) groups of a warp on different banks for the broadcast LDS.64
    static constexpr int ES = E_RAW + ((6 - E_RAW % 4) % 4);
    static constexpr int CAM_ROW = NC * DC;                  // floats per env in cam_obs
    static constexpr int TGT_ROW = NT * DT;
    static constexpr int STAGE_CAM_FLOATS = ((EPW * CAM_ROW + 3) / 4) * 4;
    static constexpr int STAGE_TGT_FLOATS = ((EPW * TGT_ROW + 3) / 4) * 4;
    static constexpr int STAGE_BYTES = (STAGE_CAM_FLOATS + STAGE_TGT_FLOATS) * 4;
    static constexpr int E_BYTES = ((EPW * ES * 8 + 15) / 16) * 16;
    // Shared memory per CTA: one small entity block per warp plus a POOL of staging buffers for
    // the packed observation rows (a warp borrows one only while it packs and stores), so that
    // occupancy is not limited by the 6 KB/env of staged rows.
    static constexpr int WARPS = 8;                          // warps per CTA
    static constexpr int ENVS_PER_CTA = WARPS * EPW;
    static constexpr int POOL_BUDGET = MATE_POOL_KB * 1024 - WARPS * E_BYTES - 64;
    static constexpr int NBUF = POOL_BUDGET / STAGE_BYTES >= 3 ? 3 : (POOL_BUDGET / STAGE_BYTES >= 2 ? 2 : 1);
    static constexpr int LOCK_OFFSET = WARPS * E_BYTES;
    static constexpr int POOL_OFFSET = LOCK_OFFSET + 64;
    static constexpr int SMEM_BYTES = POOL_OFFSET + NBUF * STAGE_BYTES;
    // a warp's rows start at a multiple of 4 floats in the output tensors => float4 / bulk copies
    static constexpr bool CAM_VEC = (EPW * CAM_ROW) % 4 == 0;
    static constexpr bool TGT_VEC = (EPW * TGT_ROW) % 4 == 0;
};

// aux outputs = the reference's public per-step attributes (environment.py:634-661)
template <int OSN>
struct AuxArgs {
    uint32_t ct_col, tt_col, cc_col, tc_col, co_col[OSN], to_col[OSN];
    int my_obs[OSN];
    float whd[NW];
    float cov_now, cov_real, transport;
    int tdone, colliding, delivered, episode_step, e, j;
};

template <int NC, int NT, int NO, int OSN>
__device__ __noinline__ void write_aux(MateStepAux ax, AuxArgs<OSN> a) {
    const int e = a.e, j = a.j;
    if (j < NT) {
        if (ax.mask_ct) for (int c = 0; c < NC; ++c) ax.mask_ct[((size_t)e * NC + c) * NT + j] = (a.ct_col >> c) & 1;
        if (ax.mask_tt) for (int t = 0; t < NT; ++t) ax.mask_tt[((size_t)e * NT + t) * NT + j] = (a.tt_col >> t) & 1;
        if (ax.target_dones) ax.target_dones[(size_t)e * NT + j] = (uint8_t)a.tdone;
        if (ax.is_colliding) ax.is_colliding[(size_t)e * NT + j] = (uint8_t)a.colliding;
        if (ax.warehouse_dist) for (int w = 0; w < NW; ++w) ax.warehouse_dist[((size_t)e * NT + j) * NW + w] = a.whd[w];
    }
    if (j < NC) {
        if (ax.mask_cc) for (int c = 0; c < NC; ++c) ax.mask_cc[((size_t)e * NC + c) * NC + j] = (a.cc_col >> c) & 1;
        if (ax.mask_tc) for (int t = 0; t < NT; ++t) ax.mask_tc[((size_t)e * NT + t) * NC + j] = (a.tc_col >> t) & 1;
    }
    for (int s = 0; s < OSN; ++s) {
        const int o = a.my_obs[s];
        if (o >= 0 && NO > 0) {
            if (ax.mask_co) for (int c = 0; c < NC; ++c) ax.mask_co[((size_t)e * NC + c) * NO + o] = (a.co_col[s] >> c) & 1;
            if (ax.mask_to) for (int t = 0; t < NT; ++t) ax.mask_to[((size_t)e * NT + t) * NO + o] = (a.to_col[s] >> t) & 1;
        }
    }
    if (j == 0) {
        if (ax.coverage) {   // coverage statistics (environment.py:966-979)
            ax.coverage[(size_t)e * 3 + 0] = a.cov_now;
            ax.coverage[(size_t)e * 3 + 1] = a.cov_real;
            ax.coverage[(size_t)e * 3 + 2] = a.transport;
        }
        if (ax.num_delivered) ax.num_delivered[e] = a.delivered;
        if (ax.episode_step) ax.episode_step[e] = a.episode_step;
    }
}

template <class S, int O>
__device__ __forceinline__ void assign_obstacles(int j, int* my_obs) {
    if constexpr (O < S::NO_) {
        constexpr int owner = S::MAP.owner[O], slot = S::MAP.slot[O];
        if (j == owner) my_obs[slot] = O;
        assign_obstacles<S, O + 1>(j, my_obs);
    }
}

// =============================================================================================
// The fused kernel
// =============================================================================================
template <int NC, int NT, int NO>
__global__ void __launch_bounds__(Shape<NC, NT, NO>::WARPS * 32, MATE_MIN_CTAS)
mate_step_kernel(const Params p) {
    using S = Shape<NC, NT, NO>;
    constexpr int G = S::G, EPW = S::EPW, OS = S::OS, DC = S::DC, DT = S::DT, CF = S::CAMF;
    constexpr int OSN = OS > 0 ? OS : 1;
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t GMASK = G == 32 ? 0xffffffffu : ((1u << G) - 1u);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / G, j = lane % G;
    const int gbase = g * G;                                   // first lane of my group
    double* Ewarp = reinterpret_cast<double*>(smem_raw + (size_t)warp * S::E_BYTES);
    uint32_t* locks = reinterpret_cast<uint32_t*>(smem_raw + S::LOCK_OFFSET);
    double* E = Ewarp + g * S::ES;                             // my env's entity block
    double* Ecam = E + S::E_CAM;
    double* Etgt = E + S::E_TGT;
    double* Eobs = E + S::E_OBS;
    uint32_t* scr = reinterpret_cast<uint32_t*>(E + S::E_SCR);
    float* F = reinterpret_cast<float*>(E + S::E_F32);         // fp32 shadow block for the prefilters
    float* Fcam = F + S::F_CAM;
    float* Ftgt = F + S::F_TGT;
    float* Fobs = F + S::F_OBS;
    constexpr int FC = S::FCAMF;
    if (threadIdx.x < S::NBUF) locks[threadIdx.x] = 0u;
    __syncthreads();

    const int env0 = (blockIdx.x * S::WARPS + warp) * EPW;     // first env of this warp
    const int e = env0 + g;                                    // my env (local index)
    const bool env_ok = e < p.num_envs;
    const int er = env_ok ? e : p.num_envs - 1;                // index used for READS (tail lanes mirror the last env)
    const int bp = p.bpad;
    const int mode = p.mode;

    // obstacles owned by this lane (static, load-balanced map)
    int my_obs[OSN];
#pragma unroll
    for (int s = 0; s < OSN; ++s) my_obs[s] = -1;
    assign_obstacles<S, 0>(j, my_obs);

    // ------------------------------------------------------------------ load state
    uint32_t tpack = 0;
    double tx = 0.0, ty = 0.0;
    if (j < NC) {
        Ecam[j * CF + 0] = p.cam_x[(size_t)j * bp + er];
        Ecam[j * CF + 1] = p.cam_y[(size_t)j * bp + er];
        Ecam[j * CF + 2] = p.cam_phi[(size_t)j * bp + er];
        Ecam[j * CF + 3] = p.cam_theta[(size_t)j * bp + er];
    }
    if (j < NT) {
        tx = p.tgt_x[(size_t)j * bp + er];
        ty = p.tgt_y[(size_t)j * bp + er];
        tpack = p.tgt_pack[(size_t)j * bp + er];
    }
#pragma unroll
    for (int s = 0; s < OS; ++s) {
        const int o = my_obs[s];
        if (o >= 0) {
            const double x = p.obs_x[(size_t)o * bp + er], y = p.obs_y[(size_t)o * bp + er], r = p.obs_r[(size_t)o * bp + er];
            Eobs[3 * o + 0] = x; Eobs[3 * o + 1] = y; Eobs[3 * o + 2] = r;
            reinterpret_cast<float4*>(Fobs)[o] = make_float4((float)x, (float)y, (float)r, 0.f);
        }
    }
    Cargo cargo;
    {
        const uint4 c0 = p.cargo[er], c1 = p.cargo[(size_t)bp + er];
        cargo.rem[0] = c0.x; cargo.rem[1] = c0.y; cargo.rem[2] = c0.z; cargo.rem[3] = c0.w;
        cargo.rem[4] = c1.x; cargo.rem[5] = c1.y; cargo.rem[6] = c1.z; cargo.rem[7] = c1.w;
    }
    const uint4 ea = p.env_a[er];
    const int4 eb = p.env_b[er];
    constexpr bool CC_CACHE = NC >= 2 && NC <= 8 && NO > 0;   // static camera<->camera lines of sight fit one u64
    unsigned long long ccw = CC_CACHE ? p.cc_clear[er] : 0ull;

    cargo.aw[0] = ea.x; cargo.aw[1] = ea.y;
    int episode_step = (int)ea.z, delivered = (int)ea.w;
    int ep_reward = eb.x, delayed_ep_reward = eb.y, episode_id = eb.w;
    float coverage_sum = __int_as_float(eb.z);
    RngKey key{p.seed, (uint32_t)(p.env_index_base + e), (uint32_t)episode_id};

    bool cargo_dirty = false, geometry_dirty = false;
    int tdone = 0;                      // target_dones[j]
    int reward_i = 0, delayed_i = 0;    // this step's rewards (integers)
    int done = 0;
    float whd[NW] = {0.f, 0.f, 0.f, 0.f};

    // ------------------------------------------------------------------ _simulate (environment.py:1326-1354)
    if (j < NC) {
        double* C = Ecam + j * CF;
        if (mode == MODE_STEP) {   // Camera.simulate (entities.py:347-360)
            const float2 a = reinterpret_cast<const float2*>(p.cam_act)[(size_t)er * NC + j];
            const double da = fmin(fmax((double)a.x, -p.cam_rot_step), p.cam_rot_step);
            const double dv = fmin(fmax((double)a.y, -p.cam_zoom_step), p.cam_zoom_step);
            const double phi = normalize_angle(C[2] + da);
            const double theta = fmin(fmax(C[3] + dv, p.cam_min_view), 180.0);
            C[2] = phi; C[3] = theta;
            if (env_ok) { p.cam_phi[(size_t)j * bp + e] = phi; p.cam_theta[(size_t)j * bp + e] = theta; }
        }
        camera_derive(C, p.cam_area_product);
        float* Fc = Fcam + j * FC;
        Fc[0] = (float)C[0]; Fc[1] = (float)C[1]; Fc[2] = (float)C[5]; Fc[3] = (float)C[6]; Fc[4] = (float)C[7]; Fc[5] = (float)C[8];
    }
    __syncwarp();
    if (mode == MODE_STEP && j < NT) {   // Target.simulate (entities.py:645-668), brute force over all discs
        const float2 a = reinterpret_cast<const float2*>(p.tgt_act)[(size_t)er * NT + j];
        const double step_size = p.tgt_step_size / (double)tp_capacity(tpack);
        StepVec s{(double)a.x, (double)a.y, 0.0, 0.0, step_size, false, false};
        const double n2 = s.vx * s.vx + s.vy * s.vy;
        if (n2 > step_size * step_size * (1.0 - 1e-12)) {
            s.n = sqrt(n2); s.has_n = true;
            if (s.n > step_size) {
                // Vector2D.norm setter (utils.py:223-229): the reference re-derives the vector from
                // (step_size, atan2(v)); v * (step_size / |v|) is the same vector to 1 ulp
                const double k = step_size / s.n;
                s.vx *= k; s.vy *= k; s.n = step_size;
            }
            s.bound = s.n * (1.0 + 1e-12);
        }
        const double desx = tx + s.vx, desy = ty + s.vy;
        {
            // fp32 broad phase: a disc farther than bound + R (+ slack for fp32 rounding) cannot touch the step
            const float ftx = (float)tx, fty = (float)ty, fb = (float)s.bound * 1.00001f + 0.01f;
#pragma unroll 1
            for (int o = 0; o < NO; ++o) {
                const float4 ob = reinterpret_cast<const float4*>(Fobs)[o];
                const float dx = ob.x - ftx, dy = ob.y - fty, reach = fb + ob.z;
                if (dx * dx + dy * dy > reach * reach && s.bound <= step_size * 1.000001) continue;
                obstruct_step(s, tx, ty, Eobs[3 * o], Eobs[3 * o + 1], Eobs[3 * o + 2]);
            }
            const float reach_c = fb + (float)p.cam_radius, reach_c2 = reach_c * reach_c;
#pragma unroll 1
            for (int c = 0; c < NC; ++c) {
                const float dx = Fcam[c * FC] - ftx, dy = Fcam[c * FC + 1] - fty;
                if (dx * dx + dy * dy > reach_c2 && s.bound <= step_size * 1.000001) continue;
                obstruct_step(s, tx, ty, Ecam[c * CF], Ecam[c * CF + 1], p.cam_radius);
            }
        }
        const double nx = fmin(fmax(tx + s.vx, -kTerrain), kTerrain);
        const double ny = fmin(fmax(ty + s.vy, -kTerrain), kTerrain);
        const int colliding = !(fabs(nx - desx) <= 1e-6 && fabs(ny - desy) <= 1e-6);
        tx = nx; ty = ny;
        tpack = (tpack & ~(1u << 27)) | ((uint32_t)colliding << 27);
    }

    // column masks of my entities
    uint32_t ct_col = 0, tt_col = 0;     // who sees my target: cameras / targets
    uint32_t cc_col = 0, tc_col = 0;     // who sees my camera: cameras / targets
    uint32_t co_col[OSN], to_col[OSN];
#pragma unroll
    for (int s = 0; s < OSN; ++s) { co_col[s] = 0; to_col[s] = 0; }
    float cov_now = 0.f, cov_real = 0.f;

    auto emit_aux = [&]() {
        if (!p.has_aux || !env_ok) return;
        const float transport = delivered > 0 ? (float)((double)delayed_ep_reward / ((double)p.reward_scale * (double)delivered)) : 0.f;
        if (!p.has_aux_detail) {   // the common case: only the info-dict scalars (environment.py:634-639)
            if (j == 0) {
                if (p.aux.coverage) {
                    p.aux.coverage[(size_t)e * 3 + 0] = cov_now;
                    p.aux.coverage[(size_t)e * 3 + 1] = cov_real;
                    p.aux.coverage[(size_t)e * 3 + 2] = transport;
                }
                if (p.aux.num_delivered) p.aux.num_delivered[e] = delivered;
                if (p.aux.episode_step) p.aux.episode_step[e] = episode_step;
            }
            return;
        }
        AuxArgs<OSN> a;
        a.ct_col = ct_col; a.tt_col = tt_col; a.cc_col = cc_col; a.tc_col = tc_col;
#pragma unroll
        for (int s = 0; s < OSN; ++s) { a.co_col[s] = co_col[s]; a.to_col[s] = to_col[s]; a.my_obs[s] = my_obs[s]; }
#pragma unroll
        for (int w = 0; w < NW; ++w) a.whd[w] = whd[w];
        a.cov_now = cov_now; a.cov_real = cov_real; a.transport = transport;
        a.tdone = tdone; a.colliding = tp_colliding(tpack); a.delivered = delivered; a.episode_step = episode_step;
        a.e = e; a.j = j;
        write_aux<NC, NT, NO, OSN>(p.aux, a);
    };

    bool auto_reset_needed = false;
    int draw_step = (mode == MODE_STEP) ? episode_step + 1 : episode_step;

    // squared thresholds of the omnidirectional sensing tests
    const double sr = p.tgt_sight_range;
    const double sr_lo = sr * sr * (1.0 - 1e-12), sr_hi = sr * sr * (1.0 + 1e-12);
    const double src = sr + p.cam_radius;
    const double src_lo = src * src * (1.0 - 1e-12), src_hi = src * src * (1.0 + 1e-12);

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const bool do_reset = (pass == 0)
            ? ((mode == MODE_RESET) && env_ok && (p.env_mask == nullptr || p.env_mask[e] != 0))
            : auto_reset_needed;
        const bool view_active = (pass == 0) || do_reset;
        // ============================================================== reset (environment.py:679-775)
        if (__any_sync(FULL, do_reset)) {
            if (do_reset && j == 0) {
                ResetCfg rc;
                rc.cam_radius = p.cam_radius; rc.cam_min_view = p.cam_min_view; rc.cam_rot_step = p.cam_rot_step;
                rc.cam_area_product = p.cam_area_product; rc.tgt_step_size = p.tgt_step_size;
                rc.obs_r_low = p.obs_r_low; rc.obs_r_high = p.obs_r_high;
                rc.cam_ranges = p.cam_ranges; rc.tgt_ranges = p.tgt_ranges; rc.obs_ranges = p.obs_ranges;
                rc.shuffle = p.shuffle; rc.num_high_capacity = p.num_high_capacity;
                rc.num_cargoes_per_target = p.num_cargoes_per_target;
                RngKey k2 = key;
                k2.episode = (uint32_t)(episode_id + 1);
                env_reset<NC, NT, NO, CF>(rc, k2, Ecam, Etgt, Eobs, scr);
            }
            __syncwarp();
            if (do_reset) {
#pragma unroll
                for (int k = 0; k < 8; ++k) cargo.rem[k] = scr[k];
                cargo.aw[0] = scr[8]; cargo.aw[1] = scr[9];
                const uint32_t cap2 = scr[10];
                episode_id += 1; key.episode = (uint32_t)episode_id;
                episode_step = 0; delivered = 0; ep_reward = 0; delayed_ep_reward = 0; coverage_sum = 0.f;
                cargo_dirty = true; geometry_dirty = true; ccw = 0ull;
                if (j < NT) {
                    tx = Etgt[2 * j]; ty = Etgt[2 * j + 1];
                    tpack = pack_target(0, -1, 0, ((cap2 >> j) & 1) ? 2 : 1, 0, 0);
                }
                tdone = 0;
                draw_step = 0;
                if (j < NC) {
                    double* C = Ecam + j * CF;
                    camera_derive(C, p.cam_area_product);
                    float* Fc = Fcam + j * FC;
                    Fc[0] = (float)C[0]; Fc[1] = (float)C[1]; Fc[2] = (float)C[5]; Fc[3] = (float)C[6]; Fc[4] = (float)C[7]; Fc[5] = (float)C[8];
                }
#pragma unroll
                for (int s = 0; s < OS; ++s) {
                    const int o = my_obs[s];
                    if (o >= 0) reinterpret_cast<float4*>(Fobs)[o] = make_float4((float)Eobs[3 * o], (float)Eobs[3 * o + 1], (float)Eobs[3 * o + 2], 0.f);
                }
            }
            __syncwarp();
        }
        // publish target positions for the view phase
        if (j < NT) { Etgt[2 * j] = tx; Etgt[2 * j + 1] = ty; Ftgt[2 * j] = (float)tx; Ftgt[2 * j + 1] = (float)ty; }
        __syncwarp();

        // ============================================================== _update_view (environment.py:1356-1388)
        // pending bits: 0..15 camera c vs my target, 16..31 camera c vs my camera
        uint32_t pending = 0, cc_reach = 0;
        bool cc_valid = true;
        const bool is_t = j < NT, is_c = j < NC;
        const double mx = is_c ? Ecam[j * CF] : 0.0, my = is_c ? Ecam[j * CF + 1] : 0.0;
        const float ftx = (float)tx, fty = (float)ty, fmx = (float)mx, fmy = (float)my;
        // ---- omnidirectional sensing by targets, Sensor.perceive (entities.py:229-232) ----
        // fp32 on squares; inside a 4e-6 relative band the fp64 test decides.  Lane t evaluates
        // "target t senses X" for one entity X per iteration; a ballot hands X's owner the whole
        // column (who senses X), which is the form the observation packer needs.
        if (__any_sync(FULL, view_active)) {
            const float fsr2 = (float)(sr * sr), fsrc2 = (float)(src * src);
            uint32_t tt_new = 0, tc_new = 0;
            if (is_t) {   // target <-> target is symmetric: my row is my column
#pragma unroll kUnrollSense
                for (int t = 0; t < NT; ++t) {
                    const float dx = Ftgt[2 * t] - ftx, dy = Ftgt[2 * t + 1] - fty, d2 = dx * dx + dy * dy;
                    bool sees = d2 < fsr2 * (1.0f - 4e-6f);
                    if (!sees && d2 <= fsr2 * (1.0f + 4e-6f)) {
                        const double ex = Etgt[2 * t] - tx, ey = Etgt[2 * t + 1] - ty;
                        sees = dist_le(ex * ex + ey * ey, sr, sr_lo, sr_hi);
                    }
                    tt_new |= (uint32_t)(sees || t == j) << t;
                }
            }
#pragma unroll kUnrollSense
            for (int c = 0; c < NC; ++c) {   // target j senses camera c
                const float dx = Fcam[c * FC] - ftx, dy = Fcam[c * FC + 1] - fty, d2 = dx * dx + dy * dy;
                bool sees = d2 < fsrc2 * (1.0f - 4e-6f);
                if (!sees && d2 <= fsrc2 * (1.0f + 4e-6f)) {
                    const double ex = Ecam[c * CF] - tx, ey = Ecam[c * CF + 1] - ty;
                    sees = dist_le(ex * ex + ey * ey, src, src_lo, src_hi);
                }
                const uint32_t col = (__ballot_sync(FULL, sees && is_t) >> gbase) & GMASK;
                if (j == c) tc_new = col;
            }
            uint32_t co_new[OSN], to_new[OSN];
#pragma unroll
            for (int s = 0; s < OSN; ++s) { co_new[s] = 0; to_new[s] = 0; }
            auto sense_obstacle = [&](const int o, const int owner, const int slot) {
                // target j senses obstacle o; camera j has obstacle o in its set
                const float4 ob = reinterpret_cast<const float4*>(Fobs)[o];
                const float rtf = (float)sr + ob.z, rt2 = rtf * rtf;
                const float dx = ob.x - ftx, dy = ob.y - fty, d2 = dx * dx + dy * dy;
                bool sees = d2 < rt2 * (1.0f - 4e-6f);
                const float rcf = (float)p.cam_rmax + ob.z, rc2 = rcf * rcf;
                const float cxd = ob.x - fmx, cyd = ob.y - fmy, c2 = cxd * cxd + cyd * cyd;
                bool inset = c2 < rc2 * (1.0f - 4e-6f);   // entities.py:363-368 (strict <)
                if ((!sees && d2 <= rt2 * (1.0f + 4e-6f)) || (NC > 0 && !inset && c2 <= rc2 * (1.0f + 4e-6f))) {
                    // inside the fp32 band: the fp64 tests decide
                    const double ex = Eobs[3 * o] - tx, ey = Eobs[3 * o + 1] - ty;
                    const double rt = sr + Eobs[3 * o + 2];
                    sees = dist_le(ex * ex + ey * ey, rt, rt * rt * (1.0 - 1e-12), rt * rt * (1.0 + 1e-12));
                    const double fx = Eobs[3 * o] - mx, fy = Eobs[3 * o + 1] - my;
                    const double rc = p.cam_rmax + Eobs[3 * o + 2];
                    inset = dist_lt(fx * fx + fy * fy, rc, rc * rc * (1.0 - 1e-12), rc * rc * (1.0 + 1e-12));
                }
                const uint32_t tcol = (__ballot_sync(FULL, sees && is_t) >> gbase) & GMASK;
                uint32_t ccol = 0;
                if (NC > 0) ccol = (__ballot_sync(FULL, inset && is_c) >> gbase) & GMASK;
                if (j == owner) {
#pragma unroll
                    for (int s = 0; s < OSN; ++s) if (s == slot) { to_new[s] = tcol; co_new[s] = ccol; }
                }
            };
#pragma unroll kUnrollSense
            for (int o = 0; o < NO; ++o) sense_obstacle(o, o % G, o / G);
            if (view_active) {
                tt_col = tt_new; tc_col = tc_new;
#pragma unroll
                for (int s = 0; s < OSN; ++s) { co_col[s] = co_new[s]; to_col[s] = to_new[s]; }
            }
        }
        if (view_active) {
            ct_col = 0; cc_col = 0;
            // (the sensing masks are computed below, warp-uniformly, with ballots)
            // ---- cameras: range + sector first (Camera.perceive, entities.py:494-501) ----
            // camera->camera occlusion is static within an episode (neither end moves): it is
            // evaluated once after reset / set_state for ALL ordered pairs and cached in `ccw`.
            cc_valid = CC_CACHE && (ccw >> 63) != 0ull;
            cc_reach = 0;
            if (is_c) cc_col = 1u << j;   // environment.py:1383-1384
#pragma unroll 1
            for (int c = 0; c < NC; ++c) {
                const double* C = Ecam + c * CF;
                const float* Fc = Fcam + c * FC;
                if (is_t && fov_reach(Fc, C, ftx, fty, tx, ty)) pending |= 1u << c;
                if (is_c && c != j) {
                    const bool reach = fov_reach(Fc, C, fmx, fmy, mx, my);
                    cc_reach |= (uint32_t)reach << c;
                    if (cc_valid) cc_col |= (uint32_t)(reach && ((ccw >> (8 * j + c)) & 1ull)) << c;
                    else if (CC_CACHE || reach) pending |= 1u << (16 + c);
                }
            }
        }
        // ---- then the stochastic transmittance draw and the occlusion test (entities.py:503-505) ----
        // (warp-uniform loop: lanes without work keep voting)
        uint32_t cc_clear_col = 0;
        while (__any_sync(FULL, pending != 0)) {
            if (pending != 0) {
                const int b = __ffs(pending) - 1;
                pending &= pending - 1;
                const int c = b & 15;
                const bool is_cam = b >= 16;
                const double* C = Ecam + c * CF;
                const double cx = C[0], cy = C[1];
                const double qx = is_cam ? Ecam[j * CF] : tx, qy = is_cam ? Ecam[j * CF + 1] : ty;
                bool sees = false;
                if (!is_cam) {   // camera->camera uses transmittance 0.0: binomial(1, 0) == 0
                    if (p.replay_transmit) sees = p.replay_transmit[((size_t)er * NC + c) * NT + j] != 0;
                    else sees = rng_u01(key, STREAM_TRANSMIT, (uint32_t)draw_step * (uint32_t)(NC * NT) + (uint32_t)(c * NT + j)) < p.transmittance;
                }
                if (!sees) {
                    if (NO == 0 || p.transmittance_is_one) {
                        sees = true;   // polyline is the flat max_sight_range circle; dist <= rs <= Rmax
                    } else {
                        const float* Fc = Fcam + c * FC;
                        const float fqx = is_cam ? Fcam[j * FC] : (float)tx, fqy = is_cam ? Fcam[j * FC + 1] : (float)ty;
                        const int fast = occlusion_fast<NO>(reinterpret_cast<const float4*>(Fobs), 1, Fc[0], Fc[1], fqx - Fc[0], fqy - Fc[1], (float)p.cam_rmax);
                        sees = fast == 1;
                        if (fast == 2) {
                            const double relx = qx - cx, rely = qy - cy;
                            sees = occlusion_exact<NO>(ObsRef{Eobs, Eobs + 1, Eobs + 2, 3}, cx, cy, relx, rely, sqrt(relx * relx + rely * rely), p.cam_rmax);
                        }
                    }
                }
                if (is_cam) cc_clear_col |= (uint32_t)sees << c; else ct_col |= (uint32_t)sees << c;
            }
        }
        const bool cc_fresh = view_active && !cc_valid && NC >= 2;
        if (__any_sync(FULL, cc_fresh)) {
            // fold the freshly evaluated static lines of sight into the mask and the per-episode cache
            unsigned long long w = 0ull;
#pragma unroll
            for (int k = 0; k < NC; ++k) w |= (unsigned long long)(__shfl_sync(FULL, cc_clear_col, gbase + k) & 0xFFu) << (8 * k);
            if (cc_fresh) {
                cc_col |= cc_reach & cc_clear_col;
                if (CC_CACHE) {
                    ccw = w | (1ull << 63);
                    if (env_ok && j == 0) p.cc_clear[e] = ccw;
                }
            }
        }
        const bool tracked = (j < NT) && ct_col != 0;
        const uint32_t tracked_bits = (__ballot_sync(FULL, tracked) >> gbase) & GMASK;

        // ============================================================== _assign_goals (environment.py:1271-1324)
        const bool step_goals = (pass == 0) && (mode == MODE_STEP);
        const bool goals_active = step_goals || do_reset;
        {
            int bounty = tp_bounty(tpack);
            const bool counted = goals_active && tracked && bounty > 0;
            const int ncount = __popc((__ballot_sync(FULL, counted) >> gbase) & GMASK);
            int r = -ncount, delayed = 0;
            int my_wh = -1;
            if (goals_active && j < NT) {
                bounty = max(bounty - (int)tracked, 0);
                // the four warehouses sit at (+-925, +-925): the one this target could be in is
                // given by the signs of its coordinates (constants.py:70-72 order: ++, -+, --, +-)
                const int wq = (ty >= 0.0) ? ((tx >= 0.0) ? 0 : 1) : ((tx >= 0.0) ? 3 : 2);
                const double ax = fabs(tx) - kWarehouseCoord, ay = fabs(ty) - kWarehouseCoord;
                if (fmax(fabs(ax), fabs(ay)) <= kWarehouseRadius) my_wh = wq;
                if (p.has_aux && p.aux.warehouse_dist) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const double wx = (w == 0 || w == 3) ? kWarehouseCoord : -kWarehouseCoord;
                        const double wy = (w < 2) ? kWarehouseCoord : -kWarehouseCoord;
                        whd[w] = (float)norm2(tx - wx, ty - wy);
                    }
                }
                tpack = (tpack & ~0xFFFFu) | (uint32_t)bounty;
            }
            uint32_t in_bits = (__ballot_sync(FULL, my_wh >= 0) >> gbase) & GMASK;
            const int old_goal = tp_goal(tpack);
            // Sequential over the targets standing in a warehouse, ascending index.  The loop runs
            // warp-wide; each group consumes its own bit set and every lane of a group performs the
            // same updates on its replicated copy of the cargo table.
            while (__any_sync(FULL, in_bits != 0)) {
                const bool act = in_bits != 0;
                const int t = act ? (__ffs(in_bits) - 1) : 0;
                in_bits &= in_bits - 1;
                const uint32_t tp_t = __shfl_sync(FULL, tpack, gbase + t);
                const int w = __shfl_sync(FULL, my_wh, gbase + t);
                if (act) {
                    int goal = tp_goal(tp_t), weight = tp_weight(tp_t), bnty = tp_bounty(tp_t);
                    const int capacity = tp_capacity(tp_t);
                    int empty = tp_empty(tp_t);
                    bool proceed = true;
                    if (goal >= 0) {
                        if (goal == w) {
                            const int reward = weight * p.freight_scale + bnty;
                            r += reward;
                            delayed += reward - (weight * p.bounty_scale - bnty);
                            delivered += weight;
                            cargo.awaiting_add(goal, -weight);
                        } else {
                            proceed = false;
                        }
                    }
                    if (proceed) {
                        bnty = 0; weight = 0; goal = -1;
                        if (cargo.row_any(w)) {
                            int new_goal;
                            if (p.replay_choice) {
                                new_goal = p.replay_choice[(size_t)er * NT + t];
                            } else {   // np_random.choice(flatnonzero(remaining[w] > 0))
                                const uint32_t ncand = (uint32_t)((cargo.get(w, 0) > 0) + (cargo.get(w, 1) > 0) + (cargo.get(w, 2) > 0) + (cargo.get(w, 3) > 0));
                                const int pick = (int)rng_below(key, STREAM_CHOICE, (uint32_t)draw_step * (uint32_t)NT + (uint32_t)t, ncand);
                                new_goal = 0;
                                int seen = 0;
#pragma unroll
                                for (int gg = 0; gg < NW; ++gg) {
                                    if (cargo.get(w, gg) > 0) { if (seen == pick) new_goal = gg; ++seen; }
                                }
                            }
                            new_goal = min(max(new_goal, 0), NW - 1);
                            const int rem = cargo.get(w, new_goal);
                            weight = min(capacity, rem);
                            cargo.add(w, new_goal, -weight);
                            bnty = weight * p.bounty_scale;
                            goal = new_goal;
                        }
                        cargo_dirty = true;
                    }
                    // empty_bits for the warehouse the target stands in (environment.py:1317-1318)
                    empty = cargo.row_any(w) ? (empty & ~(1 << w)) : (empty | (1 << w));
                    if (t == j) tpack = pack_target(bnty, goal, weight, capacity, empty, tp_colliding(tp_t));
                }
            }
            if (goals_active && j < NT) tdone = (tp_goal(tpack) != old_goal) && (old_goal >= 0);
            if (step_goals) { reward_i = r; delayed_i = delayed; }
            if (__any_sync(FULL, do_reset)) {
                if (do_reset) { tdone = 0; delivered = 0; }   // environment.py:785-788
                // targets_start_with_cargoes (environment.py:789-812): sequential over targets without a goal
                if (p.start_with_cargoes) {
#pragma unroll 1
                    for (int t = 0; t < NT; ++t) {
                        const uint32_t tp_t = __shfl_sync(FULL, tpack, gbase + t);
                        if (do_reset && tp_goal(tp_t) < 0) {
                            int perm[NW] = {0, 1, 2, 3};   // np_random.permutation(4)
#pragma unroll
                            for (int i = NW - 1; i >= 1; --i) {
                                const int k = (int)rng_below(key, STREAM_INIT_GOAL, (uint32_t)(t * 8 + i), (uint32_t)(i + 1));
                                int vi = perm[0], vk = perm[0];
#pragma unroll
                                for (int q = 1; q < NW; ++q) { vi = (q == i) ? perm[q] : vi; vk = (q == k) ? perm[q] : vk; }
#pragma unroll
                                for (int q = 0; q < NW; ++q) { if (q == i) perm[q] = vk; else if (q == k) perm[q] = vi; }
                            }
                            bool assigned = false;
#pragma unroll
                            for (int k = 0; k < NW; ++k) {
                                const int w = perm[k];
                                if (!assigned && cargo.row_any(w)) {
                                    const uint32_t ncand = (uint32_t)((cargo.get(w, 0) > 0) + (cargo.get(w, 1) > 0) + (cargo.get(w, 2) > 0) + (cargo.get(w, 3) > 0));
                                    const int pick = (int)rng_below(key, STREAM_INIT_GOAL, (uint32_t)(t * 8 + 4), ncand);
                                    int goal = 0, seen = 0;
#pragma unroll
                                    for (int gg = 0; gg < NW; ++gg) {
                                        if (cargo.get(w, gg) > 0) { if (seen == pick) goal = gg; ++seen; }
                                    }
                                    const int capacity = tp_capacity(tp_t);
                                    const int weight = min(capacity, cargo.get(w, goal));
                                    cargo.add(w, goal, -weight);
                                    if (t == j) tpack = pack_target(weight * p.bounty_scale, goal, weight, capacity, tp_empty(tp_t), 0);
                                    assigned = true;
                                }
                            }
                        }
                    }
                }
            }
        }
        // coverage statistics of the current view (environment.py:966-972)
        {
            const bool wb = (j < NT) && tp_bounty(tpack) > 0;
            const uint32_t wb_bits = (__ballot_sync(FULL, wb) >> gbase) & GMASK;
            if (view_active || goals_active) {
                const int nwb = __popc(wb_bits);
                cov_now = (float)__popc(tracked_bits) / (float)NT;
                cov_real = nwb > 0 ? (float)__popc(wb_bits & tracked_bits) / (float)nwb : 0.f;
            }
        }

        if (!step_goals) break;

        // ============================================================== finish step (environment.py:614-632)
        ep_reward += reward_i;
        delayed_ep_reward += delayed_i;
        episode_step += 1;
        coverage_sum += cov_now;
        done = !(episode_step <= p.max_episode_steps && cargo.any_awaiting());
        if (env_ok && j == 0) {
            const int r_out = p.reward_sparse ? delayed_i : reward_i;
            reinterpret_cast<float2*>(p.rewards)[e] = make_float2(-(float)r_out, (float)r_out);
            p.done[e] = (uint8_t)done;
            if (done) {
                atomicAdd(&p.stats[0], 1.0f);
                atomicAdd(&p.stats[1], (float)ep_reward);
                atomicAdd(&p.stats[2], (float)episode_step);
                atomicAdd(&p.stats[3], (float)delivered);
                atomicAdd(&p.stats[4], coverage_sum / (float)episode_step);
            }
        }
        emit_aux();   // aux reflects the step just taken (before any auto-reset)
        auto_reset_needed = env_ok && done && (p.flags & MATE_STEP_AUTO_RESET);
        if (!__any_sync(FULL, auto_reset_needed)) break;
    }
    if (mode != MODE_STEP) emit_aux();
    if (mode == MODE_STEP && lane == 0 && warp == 0) {
        const int first = blockIdx.x * S::ENVS_PER_CTA;
        const int n = min(S::ENVS_PER_CTA, p.num_envs - first);
        if (n > 0) atomicAdd(&p.stats[5], (float)n);
    }

    // ------------------------------------------------------------------ write state back
    if (env_ok) {
        if (j < NT && mode != MODE_OBSERVE) {
            p.tgt_x[(size_t)j * bp + e] = tx;
            p.tgt_y[(size_t)j * bp + e] = ty;
            p.tgt_pack[(size_t)j * bp + e] = tpack;
        }
        if (geometry_dirty) {
            if (j < NC) {
                p.cam_x[(size_t)j * bp + e] = Ecam[j * CF + 0];
                p.cam_y[(size_t)j * bp + e] = Ecam[j * CF + 1];
                p.cam_phi[(size_t)j * bp + e] = Ecam[j * CF + 2];
                p.cam_theta[(size_t)j * bp + e] = Ecam[j * CF + 3];
            }
#pragma unroll
            for (int s = 0; s < OS; ++s) {
                const int o = my_obs[s];
                if (o >= 0) {
                    p.obs_x[(size_t)o * bp + e] = Eobs[3 * o + 0];
                    p.obs_y[(size_t)o * bp + e] = Eobs[3 * o + 1];
                    p.obs_r[(size_t)o * bp + e] = Eobs[3 * o + 2];
                    p.obs_f4[(size_t)o * bp + e] = make_float4((float)Eobs[3 * o], (float)Eobs[3 * o + 1], (float)Eobs[3 * o + 2], 0.f);
                }
            }
        }
        if (j == 0 && mode != MODE_OBSERVE) {
            if (cargo_dirty) {
                p.cargo[e] = make_uint4(cargo.rem[0], cargo.rem[1], cargo.rem[2], cargo.rem[3]);
                p.cargo[(size_t)bp + e] = make_uint4(cargo.rem[4], cargo.rem[5], cargo.rem[6], cargo.rem[7]);
            }
            p.env_a[e] = make_uint4(cargo.aw[0], cargo.aw[1], (uint32_t)episode_step, (uint32_t)delivered);
            p.env_b[e] = make_int4(ep_reward, delayed_ep_reward, __float_as_int(coverage_sum), episode_id);
        }
    }

    // ------------------------------------------------------------------ joint_observation (environment.py:908-983)
    // Each entity-owning lane scatters its public state into the rows of the observers that see
    // it (the staged rows were zero-filled above, masked-out entries stay zero).
    // borrow a staging buffer from the CTA's pool (held only while packing + storing)
    int buf = 0;
    if (lane == 0) {
        buf = warp % S::NBUF;
        while (atomicCAS(&locks[buf], 0u, 1u) != 0u) {
            buf = (buf + 1 == S::NBUF) ? 0 : buf + 1;
            __nanosleep(32);
        }
        __threadfence_block();
    }
    buf = __shfl_sync(FULL, buf, 0);
    float* stage_cam = reinterpret_cast<float*>(smem_raw + S::POOL_OFFSET + (size_t)buf * S::STAGE_BYTES);
    float* stage_tgt = stage_cam + S::STAGE_CAM_FLOATS;
    {   // masked-out entries of an observation are all-zero: clear, then write only what is visible
        float4* z = reinterpret_cast<float4*>(stage_cam);
        constexpr int NZ = S::STAGE_BYTES / 16;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int i = lane; i < NZ; i += 32) z[i] = zero;
    }
    __syncwarp();
    float* srow_cam = stage_cam + g * S::CAM_ROW;
    float* srow_tgt = stage_tgt + g * S::TGT_ROW;
    constexpr int C_SELF = 13, C_TGT = 22, C_OBS = 22 + 5 * NT, C_CAM = 22 + 5 * NT + 4 * NO;
    constexpr int T_SELF = 13, T_CAM = 27, T_OBS = 27 + 7 * NC, T_TGT = 27 + 7 * NC + 4 * NO;
    if (j < NT) {   // target entity j: Target.state (entities.py:631-637)
        const float fx = (float)tx, fy = (float)ty, fsr = (float)p.tgt_sight_range;
        const int goal = tp_goal(tpack), weight = tp_weight(tpack), capacity = tp_capacity(tpack), empty = tp_empty(tpack);
        const float floaded = (goal >= 0 && weight > 0) ? 1.f : 0.f;
        {
            float* q = srow_cam + C_TGT + 5 * j;
#pragma unroll kUnrollPack
            for (int c = 0; c < NC; ++c, q += DC)
                if ((ct_col >> c) & 1) { q[0] = fx; q[1] = fy; q[2] = fsr; q[3] = floaded; q[4] = 1.f; }
        }
        {
            float* q = srow_tgt + T_TGT + 5 * j;
#pragma unroll kUnrollPack
            for (int t = 0; t < NT; ++t, q += DT)
                if ((tt_col >> t) & 1) { q[0] = fx; q[1] = fy; q[2] = fsr; q[3] = floaded; q[4] = 1.f; }
        }
        // my own row: preserved data + private state
        float* q = srow_tgt + j * DT;
        q[0] = (float)NC; q[1] = (float)NT; q[2] = (float)NO; q[3] = (float)j;
        q[4] = 925.f; q[5] = 925.f; q[6] = -925.f; q[7] = 925.f; q[8] = -925.f; q[9] = -925.f; q[10] = 925.f; q[11] = -925.f;
        q[12] = 75.f;
        q += T_SELF;
        q[0] = fx; q[1] = fy; q[2] = fsr; q[3] = floaded;
        q[4] = (float)(p.tgt_step_size / (double)capacity); q[5] = (float)capacity;
#pragma unroll
        for (int w = 0; w < NW; ++w) { q[6 + w] = (goal == w) ? (float)weight : 0.f; q[10 + w] = (float)((empty >> w) & 1); }
    }
    if (j < NC) {   // camera entity j: Camera.state (entities.py:313-324)
        const double* C = Ecam + j * CF;
        const float v0 = (float)C[0], v1 = (float)C[1], v2 = (float)p.cam_radius;
        const float v3 = (float)(C[4] * C[6]), v4 = (float)(C[4] * C[7]), v5 = (float)C[3];
        {
            float* q = srow_cam + C_CAM + 7 * j;
#pragma unroll kUnrollPack
            for (int c = 0; c < NC; ++c, q += DC)
                if ((cc_col >> c) & 1) { q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3; q[4] = v4; q[5] = v5; q[6] = 1.f; }
        }
        {
            float* q = srow_tgt + T_CAM + 7 * j;
#pragma unroll kUnrollPack
            for (int t = 0; t < NT; ++t, q += DT)
                if ((tc_col >> t) & 1) { q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3; q[4] = v4; q[5] = v5; q[6] = 1.f; }
        }
        float* q = srow_cam + j * DC;
        q[0] = (float)NC; q[1] = (float)NT; q[2] = (float)NO; q[3] = (float)j;
        q[4] = 925.f; q[5] = 925.f; q[6] = -925.f; q[7] = 925.f; q[8] = -925.f; q[9] = -925.f; q[10] = 925.f; q[11] = -925.f;
        q[12] = 75.f;
        q += C_SELF;
        q[0] = v0; q[1] = v1; q[2] = v2; q[3] = v3; q[4] = v4; q[5] = v5;
        q[6] = (float)p.cam_rmax; q[7] = (float)p.cam_rot_step; q[8] = (float)p.cam_zoom_step;
    }
#pragma unroll
    for (int s = 0; s < OS; ++s) {   // obstacle entities: Obstacle.state (entities.py:147-148)
        const int o = my_obs[s];
        if (o >= 0) {
            const float v0 = (float)Eobs[3 * o], v1 = (float)Eobs[3 * o + 1], v2 = (float)Eobs[3 * o + 2];
            {
                float* q = srow_cam + C_OBS + 4 * o;
#pragma unroll kUnrollPack
                for (int c = 0; c < NC; ++c, q += DC)
                    if ((co_col[s] >> c) & 1) { q[0] = v0; q[1] = v1; q[2] = v2; q[3] = 1.f; }
            }
            {
                float* q = srow_tgt + T_OBS + 4 * o;
#pragma unroll kUnrollPack
                for (int t = 0; t < NT; ++t, q += DT)
                    if ((to_col[s] >> t) & 1) { q[0] = v0; q[1] = v1; q[2] = v2; q[3] = 1.f; }
            }
        }
    }

    // ------------------------------------------------------------------ staged rows -> HBM
    // The warp's EPW environments are contiguous in both output tensors: one bulk (TMA) copy
    // per tensor, issued by one lane; tail warps fall back to a plain coalesced copy.
    {
        const int nvalid = min(EPW, p.num_envs - env0);
        constexpr bool BULK = (NC == 0 || S::CAM_VEC) && S::TGT_VEC;
        if (BULK && nvalid == EPW) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (NC > 0) {
                    const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage_cam);
                    float* dst = p.cam_obs + (size_t)env0 * S::CAM_ROW;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(dst), "r"(src), "r"((uint32_t)(EPW * S::CAM_ROW * 4)) : "memory");
                }
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(stage_tgt);
                float* dst = p.tgt_obs + (size_t)env0 * S::TGT_ROW;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(dst), "r"(src), "r"((uint32_t)(EPW * S::TGT_ROW * 4)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __threadfence_block();
                atomicExch(&locks[buf], 0u);
            }
        } else {
            __syncwarp();
            if (NC > 0) {
                const int nfl = max(nvalid, 0) * S::CAM_ROW;
                float* dst = p.cam_obs + (size_t)env0 * S::CAM_ROW;
                for (int i = lane; i < nfl; i += 32) dst[i] = stage_cam[i];
            }
            const int nfl = max(nvalid, 0) * S::TGT_ROW;
            float* dst = p.tgt_obs + (size_t)env0 * S::TGT_ROW;
            for (int i = lane; i < nfl; i += 32) dst[i] = stage_tgt[i];
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicExch(&locks[buf], 0u); }
        }
    }
}

}  // namespace mate
