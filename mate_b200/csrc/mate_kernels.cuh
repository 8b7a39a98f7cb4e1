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

namespace mate {

constexpr int kUnrollSense = MATE_UNROLL_SENSE;   // tuning knobs: the kernel is instruction-fetch sensitive
constexpr int kUnrollPack = MATE_UNROLL_PACK;

constexpr int NW = MATE_NUM_WAREHOUSES;
constexpr double kTerrain = 1000.0;          // mate/constants.py:52
constexpr double kWarehouseRadius = 75.0;    // mate/constants.py:67
constexpr double kWarehouseCoord = 925.0;    // mate/constants.py:70-72
constexpr double kRad2Deg = 57.295779513082320876798154814105;
constexpr double kDeg2Rad = 0.017453292519943295769236907684886;
constexpr int kResetRetries = 500;           // mate/environment.py:53

enum Mode : int { MODE_STEP = 0, MODE_OBSERVE = 1, MODE_RESET = 2 };

enum Stream : uint32_t {
    STREAM_SHUFFLE_CAM = 0, STREAM_SHUFFLE_TGT = 1, STREAM_SHUFFLE_OBS = 2, STREAM_CAPACITY = 3,
    STREAM_PLACE = 4, STREAM_CARGO = 5, STREAM_INIT_GOAL = 6, STREAM_TRANSMIT = 7, STREAM_CHOICE = 8
};

// ---- packed per-target integer state (one u32 per target) -------------------------------
// bits 0-15 bounty | 16-18 goal+1 | 19-20 cargo weight | 21-22 capacity | 23-26 empty_bits | 27 colliding
__host__ __device__ inline uint32_t pack_target(int bounty, int goal, int weight, int capacity, int empty, int colliding) {
    return (uint32_t)(bounty & 0xFFFF) | ((uint32_t)(goal + 1) << 16) | ((uint32_t)(weight & 3) << 19) |
           ((uint32_t)(capacity & 3) << 21) | ((uint32_t)(empty & 15) << 23) | ((uint32_t)(colliding & 1) << 27);
}
__host__ __device__ inline int tp_bounty(uint32_t p) { return (int)(p & 0xFFFF); }
__host__ __device__ inline int tp_goal(uint32_t p) { return (int)((p >> 16) & 7) - 1; }
__host__ __device__ inline int tp_weight(uint32_t p) { return (int)((p >> 19) & 3); }
__host__ __device__ inline int tp_capacity(uint32_t p) { return (int)((p >> 21) & 3); }
__host__ __device__ inline int tp_empty(uint32_t p) { return (int)((p >> 23) & 15); }
__host__ __device__ inline int tp_colliding(uint32_t p) { return (int)((p >> 27) & 1); }

struct Params {
    // --- state (device, struct-of-arrays, row stride = bpad) ---
    double* cam_x; double* cam_y; double* cam_phi; double* cam_theta;   // [NC][bpad]
    double* tgt_x; double* tgt_y;                                       // [NT][bpad]
    double* obs_x; double* obs_y; double* obs_r;                        // [NO][bpad]
    uint32_t* tgt_pack;                                                 // [NT][bpad]
    uint4* cargo;      // [2][bpad]  remaining_cargoes as 16 x u16
    uint4* env_a;      // [bpad] x: awaiting0|awaiting1<<16, y: awaiting2|awaiting3<<16, z: episode_step, w: delivered
    int4* env_b;       // [bpad] x: episode reward, y: delayed episode reward, z: coverage_sum (float bits), w: episode_id
    unsigned long long* cc_clear;   // [bpad] per-episode cache: bit 63 valid, bit (8 j + c) = camera c has a clear line of sight to camera j
    float* stats;      // [16] episode statistics accumulators
    // --- per-call I/O (device) ---
    const float* cam_act; const float* tgt_act;
    float* cam_obs; float* tgt_obs; float* rewards; uint8_t* done;
    const uint8_t* env_mask;
    MateStepAux aux; int has_aux; int has_aux_detail;   // detail = anything beyond coverage / num_delivered / episode_step
    const uint8_t* replay_transmit; const int8_t* replay_choice;
    // --- scalars ---
    int num_envs; int bpad; int mode; uint32_t flags;
    long long env_index_base; unsigned long long seed;
    int max_episode_steps; int num_cargoes_per_target; int num_high_capacity; int start_with_cargoes;
    int shuffle; int reward_sparse; int transmittance_is_one;
    int freight_scale; int bounty_scale; int reward_scale;
    double cam_radius, cam_min_view, cam_rmax, cam_rot_step, cam_zoom_step, cam_area_product;
    double tgt_step_size, tgt_sight_range, transmittance;
    double obs_r_low, obs_r_high;
    const double* cam_ranges; const double* tgt_ranges; const double* obs_ranges;  // device [N][4]
};

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

// ---- small math helpers --------------------------------------------------------------------
// mate/utils.py:155-158: (a + 180) % 360 - 180 with Python's float modulo.  Every angle the
// kernel normalises lies in (-540, 540), where fmod reduces to one exact add/subtract, so this
// produces the same bits as the reference expression.
__device__ __forceinline__ double normalize_angle(double a) {
    double x = a + 180.0;
    if (x < 0.0) x += 360.0;
    else if (x >= 360.0) x -= 360.0;
    return x - 180.0;
}
// sqrt(d2) <= t, decided on squares; the exact square root is only taken inside a 1e-12 band
__device__ __noinline__ bool dist_cmp_exact(double d2, double t, bool strict) {
    const double d = sqrt(d2);
    return strict ? d < t : d <= t;
}
// t2lo = t^2 (1 - 1e-12), t2hi = t^2 (1 + 1e-12)
__device__ __forceinline__ bool dist_le(double d2, double t, double t2lo, double t2hi) {
    if (d2 < t2lo) return true;
    if (d2 > t2hi) return false;
    return dist_cmp_exact(d2, t, false);
}
__device__ __forceinline__ bool dist_lt(double d2, double t, double t2lo, double t2hi) {
    if (d2 < t2lo) return true;
    if (d2 > t2hi) return false;
    return dist_cmp_exact(d2, t, true);
}
__device__ __forceinline__ double norm2(double x, double y) { return sqrt(x * x + y * y); }
__device__ __forceinline__ double atan2_deg(double y, double x) { return atan2(y, x) * kRad2Deg; }
__device__ __forceinline__ void sincos_deg(double deg, double* s, double* c) { sincos(deg * kDeg2Rad, s, c); }

// ---- Philox4x32-10, same draw scheme as oracle/mate_oracle.c ------------------------------
struct RngKey { unsigned long long seed; uint32_t env; uint32_t episode; };

__device__ __noinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ inline uint4 rng_words(const RngKey& k, uint32_t stream, uint32_t index) {
    return philox(index, stream, k.env, k.episode, (uint32_t)k.seed, (uint32_t)(k.seed >> 32));
}
__device__ inline double rng_u01(const RngKey& k, uint32_t stream, uint32_t index) {
    uint4 w = rng_words(k, stream, index);
    unsigned long long bits = ((unsigned long long)w.y << 32) | w.x;
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}
__device__ inline uint32_t rng_below(const RngKey& k, uint32_t stream, uint32_t index, uint32_t n) {
    return __umulhi(rng_words(k, stream, index).x, n);
}

// ---- per-env cargo table, replicated in every lane of the group ---------------------------
struct Cargo {
    uint32_t rem[8];   // remaining[w][g] as u16: word (w*4+g)/2, half (w*4+g)&1
    uint32_t aw[2];    // awaiting[4] as u16
    __device__ __forceinline__ int get(int w, int g) const {
        int i = w * 4 + g;
        uint32_t word = rem[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) word = (i >> 1) == k ? rem[k] : word;
        return (int)((word >> ((i & 1) * 16)) & 0xFFFF);
    }
    __device__ __forceinline__ void add(int w, int g, int delta) {
        int i = w * 4 + g;
        uint32_t inc = (uint32_t)delta << ((i & 1) * 16);
#pragma unroll
        for (int k = 0; k < 8; ++k) if ((i >> 1) == k) rem[k] += inc;
    }
    __device__ __forceinline__ bool row_any(int w) const {
        uint32_t a = rem[0] | rem[1];
#pragma unroll
        for (int k = 1; k < 4; ++k) a = (w == k) ? (rem[2 * k] | rem[2 * k + 1]) : a;
        return a != 0;
    }
    __device__ __forceinline__ int awaiting(int g) const { return (int)((aw[g >> 1] >> ((g & 1) * 16)) & 0xFFFF); }
    __device__ __forceinline__ void awaiting_add(int g, int delta) {
        uint32_t inc = (uint32_t)delta << ((g & 1) * 16);
        if (g >> 1) aw[1] += inc; else aw[0] += inc;
    }
    __device__ __forceinline__ bool any_awaiting() const { return (aw[0] | aw[1]) != 0; }
};

// =============================================================================================
// Obstacle.obstruct(ray, keep_tangential=True) for target motion (mate/entities.py:158-184).
// (vx, vy) is the step vector with cached norm n (n < 0 => recompute), cached angle `ang`
// (valid if has_ang), origin (ox, oy); disc centre (px, py), radius R.
// =============================================================================================
struct StepVec { double vx, vy, n, ang, bound; bool has_n, has_ang; };   // bound >= |v| for the cheap reject

__device__ __noinline__ StepVec obstruct_exact(StepVec s, double ox, double oy, double px, double py, double R) {
    const double relx = px - ox, rely = py - oy;
    const double reln = norm2(relx, rely);
    if (!s.has_n) { s.n = norm2(s.vx, s.vy); s.has_n = true; }
    const double norm = s.n;
    if (norm == 0.0 || reln < R) {   // return -ray
        s.vx = -s.vx; s.vy = -s.vy; s.has_n = false; s.has_ang = false;
        return s;
    }
    if (reln >= norm + R) return s;
    const double inner = relx * s.vx + rely * s.vy;
    if (inner >= 0.0) {
        const double c = fmin(1.0, inner / (reln * norm));
        const double perp = reln * sqrt(1.0 - c * c);
        if (R > perp) {
            const double hc = sqrt(R * R - perp * perp);
            const double nn = fmax(0.0, reln * c - hc);
            if (nn < norm) {
                if (!s.has_ang) { s.ang = atan2_deg(s.vy, s.vx); s.has_ang = true; }
                double sn, cs;
                sincos_deg(s.ang, &sn, &cs);
                const double radx = (ox + nn * cs) - px, rady = (oy + nn * sn) - py;
                const double k = (norm - nn) * hc / (R * R);
                s.vx = s.vx + radx * k; s.vy = s.vy + rady * k;
                s.has_n = false; s.has_ang = false;
                s.bound = fabs(s.vx) + fabs(s.vy);
            }
        }
    }
    return s;
}

// cheap conservative reject: with |v| <= bound, `relative.norm >= norm + radius` certainly holds
__device__ __forceinline__ void obstruct_step(StepVec& s, double ox, double oy, double px, double py, double R) {
    const double relx = px - ox, rely = py - oy;
    const double reach = s.bound + R;
    if (relx * relx + rely * rely > reach * reach * (1.0 + 1e-9)) return;
    s = obstruct_exact(s, ox, oy, px, py, R);
}

// =============================================================================================
// On-the-fly field-of-view range: value of the reference's sampled (phi, rho) polyline
// (Camera.add_obstacles + interp1d, mate/entities.py:362-479, 507-511) at bearing `a`
// (degrees, already normalised to [-180, 180)), WITHOUT materialising the polyline.
// The polyline's sample angles are: the 360 integer degrees; per visible obstacle the four
// edge rays L-+0.01, R-+0.01 and the lattice linspace(L, R, n+1).  We find the two samples
// that bracket `a`, cast those two rays against all obstacle discs (sequential shortening ==
// min over discs) and interpolate linearly like np.interp.
// =============================================================================================
struct RaySample { double angle; double n0; int tangent_of; };

template <int NO>
__device__ __forceinline__ double cast_ray(const double* __restrict__ Eobs, double cx, double cy,
                                           double angle, double n0, int tangent_of) {
    double sn, cs;
    sincospi(angle * (1.0 / 180.0), &sn, &cs);
    double n = n0;
#pragma unroll 1
    for (int o = 0; o < NO; ++o) {
        if (o == tangent_of) continue;   // exact tangent ray: never shortened by its own disc
        const double relx = Eobs[3 * o] - cx, rely = Eobs[3 * o + 1] - cy, R = Eobs[3 * o + 2];
        const double proj = relx * cs + rely * sn;
        if (proj < 0.0) continue;
        const double d2 = relx * relx + rely * rely;
        const double perp2 = d2 - proj * proj;
        if (!(R * R > perp2)) continue;
        const double hc = sqrt(R * R - fmax(perp2, 0.0));
        const double nn = fmax(0.0, proj - hc);
        if (nn < n) n = nn;
    }
    return n;
}

__device__ __forceinline__ void consider(RaySample& P, RaySample& S, double a, double s, double n0, int tangent_of) {
    if (s <= a) {
        if (s > P.angle || (s == P.angle && n0 < P.n0)) { P.angle = s; P.n0 = n0; P.tangent_of = tangent_of; }
    } else {
        if (s < S.angle || (s == S.angle && n0 < S.n0)) { S.angle = s; S.n0 = n0; S.tangent_of = tangent_of; }
    }
}

template <int NO>
__device__ __noinline__ double sight_range_at(const double* __restrict__ Eobs, double cx, double cy,
                                              double rmax, double a, double ux, double uy) {
    // (ux, uy): unit vector of bearing a (rel / dist), used only for the conservative prefilter
    const double fl = floor(a);
    RaySample P{fl, rmax, -1}, S{fl + 1.0, rmax, -1};
    const double c1 = 0.99984154; // cos(1.02 deg)
    const double s1 = 0.01780139; // sin(1.02 deg)
#pragma unroll 1
    for (int o = 0; o < NO; ++o) {
        const double relx = Eobs[3 * o] - cx, rely = Eobs[3 * o + 1] - cy, R = Eobs[3 * o + 2];
        const double d2 = relx * relx + rely * rely;
        {   // entities.py:365 (strict <) and :378 (camera inside the disc), on squares
            const double reach = rmax + R, reach2 = reach * reach;
            if (d2 > reach2 * (1.0 + 1e-12)) continue;
            if (d2 > reach2 * (1.0 - 1e-12) && !dist_cmp_exact(d2, reach, true)) continue;
            if (d2 < R * R * (1.0 + 1e-12) && dist_cmp_exact(d2, R, true)) return 0.0;
        }
        // prefilter: can any sample angle of this obstacle fall inside (floor(a), floor(a)+1)?
        // angular distance bearing<->centre must be <= half + 1.02 deg
        const double p = relx * ux + rely * uy + R * s1 * 1.0000001;
        if (p < 0.0) continue;
        if (p * p < (d2 - R * R) * (c1 * c1) * 0.9999999) continue;
        const double d = sqrt(d2);
        const double ang_o = atan2_deg(rely, relx);
        const double half = asin(R / d) * kRad2Deg;
        const double left = ang_o - half, right = ang_o + half;
        consider(P, S, a, normalize_angle(left - 0.01), rmax, -1);
        consider(P, S, a, normalize_angle(left + 0.01), rmax, -1);
        consider(P, S, a, normalize_angle(right - 0.01), rmax, -1);
        consider(P, S, a, normalize_angle(right + 0.01), rmax, -1);
        const int two_half = (int)(2.0 * half);
        const int nlat = two_half > 16 ? two_half : 16;
        const double step = (right - left) / (double)nlat;   // np.linspace: arange(num) * step + start
        const double max_rho = fmin(rmax, d + R);
        if (step > 0.0) {
#pragma unroll 1
            for (int k = -1; k <= 1; ++k) {
                const double ap = a + 360.0 * (double)k;
                if (ap < left - step || ap > right + step) continue;
                const int i0 = (int)floor((ap - left) / step);
#pragma unroll 1
                for (int i = i0 - 1; i <= i0 + 2; ++i) {
                    if (i < 0 || i > nlat) continue;
                    const double raw = (i == nlat) ? right : ((double)i * step + left);
                    consider(P, S, a, normalize_angle(raw), max_rho, (i == 0 || i == nlat) ? o : -1);
                }
            }
        }
    }
    const double rho_p = cast_ray<NO>(Eobs, cx, cy, P.angle, P.n0, P.tangent_of);
    if (a == P.angle) return rho_p;                // np.interp: exact hit on a sample
    // the closing sample (phi0 + 360, rho0) of the polyline is the -180 grid ray (entities.py:470-471)
    const double s_angle = (S.angle >= 180.0) ? -180.0 : S.angle;
    const double rho_s = cast_ray<NO>(Eobs, cx, cy, s_angle, S.n0, S.tangent_of);
    const double slope = (rho_s - rho_p) / (S.angle - P.angle);
    return slope * (a - P.angle) + rho_p;
}

// Camera.perceive (mate/entities.py:491-505) up to the stochastic draw, exact arithmetic of
// the reference: returns 0 = not in range/sector, 1 = reached the draw.
__device__ __noinline__ int fov_reach_exact(double cx, double cy, double phi, double theta, double rs,
                                            double qx, double qy) {
    const double relx = qx - cx, rely = qy - cy;
    const double dist = norm2(relx, rely);
    if (dist > rs) return 0;
    const double ang = atan2_deg(rely, relx);
    double ra = fabs(phi - ang);
    ra = fmin(ra, 360.0 - ra);
    if (ra * 2.0 > theta) return 0;
    return 1;
}

// The same two tests decided in fp32 on squares / dot products (no sqrt, no atan2).  fp32
// coordinates carry <= 6e-5 absolute error, i.e. <= 1e-5 relative on these quantities; only when
// a test falls inside a 4e-5 relative band around its boundary is the exact fp64 expression of
// the reference evaluated (C = fp64 camera block {x, y, phi, theta, rs, ...}).
// F = fp32 camera block {x, y, rs^2, cos phi, sin phi, cos^2(theta/2)}.
__device__ __forceinline__ int fov_reach(const float* __restrict__ F, const double* __restrict__ C,
                                         float fqx, float fqy, double qx, double qy) {
    const float relx = fqx - F[0], rely = fqy - F[1];
    const float d2 = relx * relx + rely * rely;
    const float rs2 = F[2];
    if (d2 > rs2 * (1.0f + 4e-5f)) return 0;
    const float dot = relx * F[3] + rely * F[4];
    const float sq = dot >= 0.0f ? dot * dot : -(dot * dot);
    const float diff = sq - d2 * F[5];           // >= 0  <=>  angle(rel, heading) <= theta / 2
    const float band = 4e-5f * d2 + 1e-3f;
    if (diff < -band) return 0;
    if (diff > band && d2 < rs2 * (1.0f - 4e-5f)) return 1;
    return fov_reach_exact(C[0], C[1], C[2], C[3], C[4], qx, qy);
}

// Conservative occlusion classification of the query point q = cam + rel against all obstacle
// discs, WITHOUT evaluating the sampled polyline.  Both polyline samples that bracket the query
// bearing lie inside a fan of +-1.05 degrees around it (integer-degree grid).  Per disc, with
// `perp` the distance of its centre from the line of sight and `w` the half-width of the fan at
// the farthest range where it can meet the disc, every ray of the fan passes the centre at a
// lateral offset in [pmin, pmax] = [perp - w, perp + w], and a ray with offset p < R is cut at
// range rho(p) = sqrt(d_o^2 - p^2) - sqrt(R^2 - p^2), which grows with p.  Hence
//   * perp - w > R, disc behind the camera or entirely beyond the target: the disc is irrelevant;
//   * pmax < R and |rel| > rho(pmax): every fan ray is cut short of the target  -> occluded (0);
//   * |rel| < rho(pmin): no fan ray is cut before the target                    -> irrelevant;
//   * otherwise (silhouette edge inside the fan, or target at the disc's front surface): exact.
// Returns 1 = certainly visible, 0 = certainly occluded, 2 = evaluate the polyline exactly.
// All margins (>= 0.05 units, 0.04 degrees) are far above fp32 rounding (<= 1e-3 units here), so
// the classification runs in fp32 on the shadow block Fobs = {x, y, r, -} per obstacle.
template <int NO>
__device__ __forceinline__ int occlusion_fast(const float* __restrict__ Fobs, float cx, float cy,
                                              float relx, float rely, float rmax) {
    const float d2 = relx * relx + rely * rely;
    const float inv_dist = rsqrtf(d2);
    const float dist = d2 * inv_dist;
    const float ux = relx * inv_dist, uy = rely * inv_dist;    // unit bearing
    const float tan_fan = 0.018332f;                           // tan(1.05 deg)
    const float d_hi = dist * (1.0f + 1e-4f) + 0.05f, d_lo = dist * (1.0f - 1e-4f) - 0.05f;
    bool all_clear = true;
#pragma unroll 1
    for (int o = 0; o < NO; ++o) {
        const float4 ob = reinterpret_cast<const float4*>(Fobs)[o];
        const float ox = ob.x - cx, oy = ob.y - cy, R = ob.z;
        const float do2 = ox * ox + oy * oy;
        const float reach = rmax + R + 0.05f;
        if (do2 > reach * reach) continue;                     // certainly not in the camera's obstacle set (entities.py:365)
        const float proj = ox * ux + oy * uy;                  // along the line of sight
        if (proj + R < 0.0f || proj - R > d_hi) continue;      // behind the camera / beyond the target
        const float perp = fabsf(ox * uy - oy * ux);           // distance of the centre from the line of sight
        const float w = tan_fan * fminf(d_hi, proj + R) + 0.05f;
        const float pmin = fmaxf(perp - w, 0.0f), pmax = perp + w;
        if (pmin >= R) continue;                               // the fan passes beside the disc
        const float inner = rmax + R - 0.05f;
        const bool in_set = do2 < inner * inner;               // certainly in the camera's obstacle set
        const float R2 = R * R;
        if (pmax < R * 0.9999f && in_set && do2 > R2 + 1.0f) {
            const float rho_max = sqrtf(do2 - pmax * pmax) - sqrtf(R2 - pmax * pmax);
            if (d_lo > rho_max * 1.0001f) return 0;            // every ray of the fan is cut before the target
        }
        const float rho_min = sqrtf(fmaxf(do2 - pmin * pmin, 0.0f)) - sqrtf(R2 - pmin * pmin);
        if (d_hi < rho_min * 0.9999f) continue;                // the target is in front of the disc
        all_clear = false;
    }
    return all_clear ? 1 : 2;
}

// Camera.perceive after the draw (entities.py:505): dist <= sight_range_at(angle) * (1 + 1e-6)
template <int NO>
__device__ __noinline__ bool occlusion_exact(const double* __restrict__ Eobs, double cx, double cy,
                                             double relx, double rely, double dist, double rmax) {
    const double ang = atan2_deg(rely, relx);
    const double range = sight_range_at<NO>(Eobs, cx, cy, rmax, normalize_angle(ang), relx / dist, rely / dist);
    return dist <= range * (1.0 + 1e-6);
}

// =============================================================================================
// Rare paths, kept out of line so that the hot path stays inside the instruction cache
// =============================================================================================

// C = {x, y, phi, theta, rs, rs^2, cos phi, sin phi, cos^2(theta/2)}
__device__ __noinline__ void camera_derive(double* C, double area_product) {
    const double rs = sqrt(area_product / C[3]);      // entities.py:334,360
    C[4] = rs; C[5] = rs * rs;
    double sn, cs;
    sincospi(C[2] * (1.0 / 180.0), &sn, &cs);
    C[6] = cs; C[7] = sn;
    const double ch = cospi(C[3] * (1.0 / 360.0));
    C[8] = ch * ch;
}

struct ResetCfg {
    double cam_radius, cam_min_view, cam_rot_step, cam_area_product, tgt_step_size, obs_r_low, obs_r_high;
    const double* cam_ranges; const double* tgt_ranges; const double* obs_ranges;
    int shuffle, num_high_capacity, num_cargoes_per_target;
};

// MultiAgentTracking.reset (mate/environment.py:679-775) for one environment, executed by ONE
// lane: entity shuffle, capacities, rejection placement (cameras, obstacles, targets) and the
// cargo table, on the counter-based Philox streams.  Results go to the shared-memory entity
// block and to the 16-word scratch: [0..7] remaining cargoes (u16 pairs), [8..9] awaiting
// (u16 pairs), [10] capacity-2 bit set.
template <int NC, int NT, int NO, int CF>
__device__ __noinline__ void env_reset(ResetCfg cfg, RngKey key, double* Ecam, double* Etgt, double* Eobs,
                                       uint32_t* scr) {
    int perm_c[NC > 0 ? NC : 1], perm_t[NT], perm_o[NO > 0 ? NO : 1];
    for (int i = 0; i < NC; ++i) perm_c[i] = i;
    for (int i = 0; i < NT; ++i) perm_t[i] = i;
    for (int i = 0; i < NO; ++i) perm_o[i] = i;
    if (cfg.shuffle) {   // environment.py:707-710 (Fisher-Yates from the top, like RandomState.shuffle)
        for (int i = NC - 1; i >= 1; --i) { int k = (int)rng_below(key, STREAM_SHUFFLE_CAM, i, i + 1); int t = perm_c[i]; perm_c[i] = perm_c[k]; perm_c[k] = t; }
        for (int i = NT - 1; i >= 1; --i) { int k = (int)rng_below(key, STREAM_SHUFFLE_TGT, i, i + 1); int t = perm_t[i]; perm_t[i] = perm_t[k]; perm_t[k] = t; }
        for (int i = NO - 1; i >= 1; --i) { int k = (int)rng_below(key, STREAM_SHUFFLE_OBS, i, i + 1); int t = perm_o[i]; perm_o[i] = perm_o[k]; perm_o[k] = t; }
    }
    uint32_t cap2 = 0;   // bit t set => capacity 2 (environment.py:712-722)
    if (cfg.num_high_capacity > 0) {
        if (cfg.shuffle) {
            int idx[NT];
            for (int i = 0; i < NT; ++i) idx[i] = i;
            for (int i = 0; i < cfg.num_high_capacity; ++i) {
                int k = i + (int)rng_below(key, STREAM_CAPACITY, i, NT - i);
                int t = idx[i]; idx[i] = idx[k]; idx[k] = t;
                cap2 |= 1u << idx[i];
            }
        } else {
            for (int i = 0; i < cfg.num_high_capacity; ++i) cap2 |= 1u << i;
        }
    }
    // rejection placement (environment.py:724-737): cameras, obstacles, targets
    int serial = 0;
    for (int kind = 0; kind < 3; ++kind) {
        const int count = kind == 0 ? NC : (kind == 1 ? NO : NT);
        for (int i = 0; i < count; ++i, ++serial) {
            const double* range = kind == 0 ? cfg.cam_ranges + 4 * perm_c[i]
                                : (kind == 1 ? cfg.obs_ranges + 4 * perm_o[i] : cfg.tgt_ranges + 4 * perm_t[i]);
            const double r0 = range[0], r1 = range[1], r2 = range[2], r3 = range[3];
            const double min_distance = kind == 2 ? 0.0 : cfg.tgt_step_size;
            double x = 0, y = 0, radius = kind == 0 ? cfg.cam_radius : 0.0, phi = 0, theta = 0, rs = 0;
            bool ok = false;
            for (int attempt = 0; attempt < kResetRetries && !ok; ++attempt) {
                const uint32_t base = ((uint32_t)serial * kResetRetries + (uint32_t)attempt) * 8u;
                if (kind == 1)   // Obstacle.reset: radius first (entities.py:150-152)
                    radius = __dadd_rn(cfg.obs_r_low, __dmul_rn(cfg.obs_r_high - cfg.obs_r_low, rng_u01(key, STREAM_PLACE, base + 2)));
                x = __dadd_rn(r0, __dmul_rn(r1 - r0, rng_u01(key, STREAM_PLACE, base + 0)));   // Entity.reset (entities.py:60-65)
                y = __dadd_rn(r2, __dmul_rn(r3 - r2, rng_u01(key, STREAM_PLACE, base + 1)));
                const double lim = __dsub_rn(kTerrain, __dmul_rn(1.2, radius));
                x = fmin(fmax(x, -lim), lim);
                y = fmin(fmax(y, -lim), lim);
                if (kind == 0) {   // Camera.reset (entities.py:326-334)
                    const uint32_t nrot = (uint32_t)(360.0 / cfg.cam_rot_step);
                    phi = normalize_angle(__dmul_rn(cfg.cam_rot_step, (double)rng_below(key, STREAM_PLACE, base + 3, nrot)));
                    theta = __dadd_rn(cfg.cam_min_view, __dmul_rn(180.0 - cfg.cam_min_view, rng_u01(key, STREAM_PLACE, base + 4)));
                    rs = sqrt(cfg.cam_area_product / theta);
                }
                ok = true;
                for (int w = 0; w < NW && ok; ++w) {   // warehouse discs, radius 0.75 * 75 (environment.py:724-727)
                    const double wx = (w == 0 || w == 3) ? kWarehouseCoord : -kWarehouseCoord;
                    const double wy = (w < 2) ? kWarehouseCoord : -kWarehouseCoord;
                    const double dx = x - wx, dy = y - wy;
                    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    if (__dmul_rn(dist, 1.0 + 1e-6) < __dadd_rn(__dadd_rn(radius, 0.75 * kWarehouseRadius), min_distance)) ok = false;
                }
                const int ncam_placed = kind == 0 ? i : NC;
                for (int q = 0; q < ncam_placed && ok; ++q) {
                    const double dx = x - Ecam[q * CF], dy = y - Ecam[q * CF + 1];
                    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    if (__dmul_rn(dist, 1.0 + 1e-6) < __dadd_rn(__dadd_rn(radius, cfg.cam_radius), min_distance)) ok = false;
                    else if (kind == 0 && dist < __dmul_rn(0.1, fmin(rs, Ecam[q * CF + 4]))) ok = false;   // Camera.overlap (entities.py:484-489)
                }
                const int nobs_placed = kind == 0 ? 0 : (kind == 1 ? i : NO);
                for (int q = 0; q < nobs_placed && ok; ++q) {
                    const double dx = x - Eobs[3 * q], dy = y - Eobs[3 * q + 1];
                    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    if (__dmul_rn(dist, 1.0 + 1e-6) < __dadd_rn(__dadd_rn(radius, Eobs[3 * q + 2]), min_distance)) ok = false;
                }
                // already placed targets have radius 0 and targets use min_distance 0:
                // dist * (1 + 1e-6) < 0 never holds (entities.py:96-100)
            }
            if (!ok && kind == 1) radius = 0.0;   // environment.py:734-736
            if (kind == 0) { Ecam[i * CF] = x; Ecam[i * CF + 1] = y; Ecam[i * CF + 2] = phi; Ecam[i * CF + 3] = theta; Ecam[i * CF + 4] = rs; }
            else if (kind == 1) { Eobs[3 * i] = x; Eobs[3 * i + 1] = y; Eobs[3 * i + 2] = radius; }
            else { Etgt[2 * i] = x; Etgt[2 * i + 1] = y; }
        }
    }
    // cargo table (environment.py:768-775)
    Cargo cargo;
    for (int k = 0; k < 8; ++k) cargo.rem[k] = 0;
    uint32_t draw = 0;
    for (;;) {
        bool all_rows = true;
        for (int w = 0; w < NW; ++w) all_rows = all_rows && cargo.row_any(w);
        if (all_rows) break;
        for (int i = 0; i < cfg.num_cargoes_per_target * NT; ++i, ++draw) {
            const uint4 w4 = rng_words(key, STREAM_CARGO, draw);
            const int sender = (int)__umulhi(w4.x, NW);
            int recipient = (int)__umulhi(w4.y, NW - 1);
            if (recipient >= sender) recipient += 1;   // choice(4, size=2, replace=False)
            cargo.add(sender, recipient, 1);
        }
    }
    cargo.aw[0] = cargo.aw[1] = 0;
    for (int gg = 0; gg < NW; ++gg) { int sum = 0; for (int w = 0; w < NW; ++w) sum += cargo.get(w, gg); cargo.awaiting_add(gg, sum); }
    for (int k = 0; k < 8; ++k) scr[k] = cargo.rem[k];
    scr[8] = cargo.aw[0]; scr[9] = cargo.aw[1]; scr[10] = cap2;
}

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
                        const int fast = occlusion_fast<NO>(Fobs, Fc[0], Fc[1], fqx - Fc[0], fqy - Fc[1], (float)p.cam_rmax);
                        sees = fast == 1;
                        if (fast == 2) {
                            const double relx = qx - cx, rely = qy - cy;
                            sees = occlusion_exact<NO>(Eobs, cx, cy, relx, rely, sqrt(relx * relx + rely * rely), p.cam_rmax);
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
