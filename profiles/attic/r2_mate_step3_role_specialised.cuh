// mate_step3.cuh -- MultiAgentTracking.step with ROLE-SPECIALISED warps (third generation of the step kernel).
//
// mate_step_kernel2 gives a warp tile (32 environments, one lane each) to ONE warp: target simulate -> view ->
// goals -> observation rows, a chain of ~23 k dependent warp instructions during whose first half the HBM is idle and
// whose second half (407 MB of rows at 65 536 environments) is write-bandwidth bound (profiles/r2a_summary.md).
// The reference's step has more parallelism than that (paths relative to the reference root):
//   * after Target.simulate (mate/entities.py:645-668) the target side -- Sensor.perceive for the targets
//     (entities.py:229-232), the warehouse / cargo state machine of _assign_goals (mate/environment.py:1282-1324) and
//     the TARGET rows of joint_observation (environment.py:947-983) -- does not depend on what the cameras see: the
//     only place where tracked bits enter is the bounty arithmetic of the reward (environment.py:1275-1280, 1293-1296),
//     and the bounty is in no observation;
//   * the camera side -- Camera.simulate (entities.py:347-360), Camera.perceive with its transmittance draws and the
//     obstacle-occluded field of view (entities.py:491-511), the CAMERA rows -- needs the new target positions and
//     nothing else from the target side (the `loaded` flag of a target's public state changes in _assign_goals).
// So a CTA is two warps that share one tile in shared memory:
//   warp 0, TARGET role: Target.simulate (fast path + near-disc queue) | sensing masks | cargo state machine, done |
//                        target rows -> HBM | rewards with the tracked bits, episode bookkeeping, state write-back;
//   warp 1, CAMERA role: Camera.simulate | camera obstacle sets, camera-camera view | (new target positions) range +
//                        sector, transmittance draws, occlusion queues | camera rows -> HBM.
// They meet at five named barriers (bar.arrive / bar.sync, 64 threads).  The target rows (2/3 of the bytes) leave the
// SM while the camera warp still works on its occlusion queue, the chain of a tile is about half as long, and twice as
// many warps hide each other's latencies.  Episodes that end are rare (B / 10 001 per step): those lanes are left out
// of both roles' rows and handled after the tile by the second pass of mate_step_kernel2's tile function (MODE_LATE:
// adopt the prepared episode or reset in place, first view, initial goals, rows).
//
// Results are bit-identical to mate_step_kernel2 (same arithmetic per predicate, same counter-based draws); the
// parity tests run both (MATE_B200_KERNEL=2 selects the older kernel).
#pragma once

#include "mate_step.cuh"

#ifndef MATE3_MIN_CTAS
#define MATE3_MIN_CTAS 14      // 2-warp CTAs per SM: caps registers at 72 so that 65 536 envs = 2048 tiles (13.8 per SM) are ONE wave
#endif
#ifndef MATE3_TILES_PER_CTA
#define MATE3_TILES_PER_CTA 1  // tiles a CTA walks through (2 with MATE3_MIN_CTAS 8 = 128 registers, same residency)
#endif

#ifndef MATE3_PACE_T
#define MATE3_PACE_T 0         // SM cycles per environment while the target role packs (0 = unpaced)
#endif
#ifndef MATE3_PACE_C
#define MATE3_PACE_C 0         // same for the camera role
#endif
#ifndef MATE3_X
#define MATE3_X 0              // TIMING EXPERIMENTS ONLY (rows missing): 1 no target rows, 2 no camera rows, 4 target rows after the camera masks
#endif

namespace mate {

template <int NC, int NT, int NO>
struct Shape3 : Shape2<NC, NT, NO> {
    using S2 = Shape2<NC, NT, NO>;
    // the tile block has the Shape2 layout (stage | masks | fp32 entries | pair queue | exact queue), so that the
    // MODE_LATE pass can run on it, plus the near-disc queue of the target role and the role exchange words
    static constexpr int OFF_Q3 = S2::OFF_Q2 + S2::QCAP * 2;
    static constexpr int OFF_X = OFF_Q3 + S2::QCAP * 2;
    static constexpr int TILE_BYTES = ((OFF_X + 16 + 15) / 16) * 16;
    static constexpr int TPC = MATE3_TILES_PER_CTA;
};

// Hardware barriers are a per-SM resource (64 on sm_100a; a CTA that uses ids 0..5 limits the SM to 10 CTAs: measured,
// profiles/r2c_summary.md), so the five meeting points of a tile share three ids.  An id is reused only after both warps
// have passed its previous use: the camera role arrives at CAM_MASKS / TILE_DONE after it has passed TGT_READY, which the
// target role signals after it has passed CAM_READY.  GOALS_READY needs its own id: the target role may signal it before
// the camera role has reached TGT_READY.
enum RoleBarrier : int { BAR_CAM_READY = 1, BAR_TGT_READY = 2, BAR_CAM_MASKS = 1, BAR_GOALS_READY = 3, BAR_TILE_DONE = 2 };

// bar.sync waits until both role warps have arrived; bar.arrive only signals.  The fence before an arrive makes the
// shared-memory writes of this warp visible to the warp that waits.
__device__ __forceinline__ void role_sync(const int id) {
    __syncwarp();
    asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory");
}
__device__ __forceinline__ void role_arrive(const int id) {
    __syncwarp();
    __threadfence_block();
    asm volatile("bar.arrive %0, 64;" :: "r"(id) : "memory");
}

// second pass of the tile function of mate_step_kernel2 for the lanes whose episode ended in this step
template <int NC, int NT, int NO>
__device__ __noinline__ void late_resets(const Params& p, unsigned char* tb, const int env0, const uint32_t late_lanes) {
    tile_body<NC, NT, NO>(p, tb, env0, MODE_LATE, late_lanes);
}

// =============================================================================================
// TARGET role
// =============================================================================================
template <int NC, int NT, int NO>
__device__ __noinline__ void target_role(const Params& p, unsigned char* tb, const int env0) {
    using S = Shape3<NC, NT, NO>;
    constexpr int MW = S::MW, CV = S::CV;
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float* stage = reinterpret_cast<float*>(tb + S::OFF_STAGE);
    uint32_t* mk = reinterpret_cast<uint32_t*>(tb + S::OFF_MASK);
    float* val = reinterpret_cast<float*>(tb + S::OFF_VAL);
    uint16_t* queue = reinterpret_cast<uint16_t*>(tb + S::OFF_Q3);
    uint32_t* xch = reinterpret_cast<uint32_t*>(tb + S::OFF_X);
    uint32_t* mymk = mk + lane * S::MSTRIDE;
    float* myval = val + lane * S::VSTRIDE;
    const float* mycam = myval + S::V_C;

    const int e = env0 + lane;
    const bool env_ok = e < p.num_envs;
    const int er = env_ok ? e : p.num_envs - 1;                // tail lanes mirror the last env (reads only)
    const size_t bp = p.bpad;
    const int nvalid = min(32, p.num_envs - env0);
    const uint32_t valid_bits = nvalid >= 32 ? FULL : ((1u << nvalid) - 1u);

    const uint4 ea = p.env_a[er];
    const int4 eb = p.env_b[er];

    // ------------------------------------------------------------------ Target.simulate (entities.py:645-668)
    {
        uint32_t slow = 0;     // targets that may touch a disc: re-simulated exactly below
        // fp32 broad phase on the old locations: a disc farther than step_size + R (+ slack for fp32 rounding)
        // cannot touch the step
        float otx[NT], oty[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) { otx[t] = (float)p.tgt_x[(size_t)t * bp + er]; oty[t] = (float)p.tgt_y[(size_t)t * bp + er]; }
        const float fb = (float)p.tgt_step_size * 1.00001f + 0.01f;
#pragma unroll (NO <= 12 ? 12 : 4)
        for (int o = 0; o < NO; ++o) {
            const float4 ob = p.obs_f4[(size_t)o * bp + er];
            const float reach = fb + ob.z, reach2 = reach * reach;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float dx = ob.x - otx[t], dy = ob.y - oty[t];
                slow |= (uint32_t)(!(dx * dx + dy * dy > reach2)) << t;
            }
        }
        role_sync(BAR_CAM_READY);   // the camera role has written the fp32 camera entries (x, y, theta, heading)
        const float reach_c = fb + (float)p.cam_radius, reach_c2 = reach_c * reach_c;
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
            const float cx = mycam[CV * c], cy = mycam[CV * c + 1];
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float dx = cx - otx[t], dy = cy - oty[t];
                slow |= (uint32_t)(!(dx * dx + dy * dy > reach_c2)) << t;
            }
        }
        if (MATE2_PF_OBS == 1 && NO > 0 && slow != 0) prefetch_discs64<NO>(p, er);
        double ntx_ = p.tgt_x[er], nty_ = p.tgt_y[er];
        uint32_t npk_ = p.tgt_pack[er];
        float2 nta_ = reinterpret_cast<const float2*>(p.tgt_act)[(size_t)er * NT];
#pragma unroll 1
        for (int t = 0; t < NT; ++t) {   // the next target's state is fetched while this one is stepped
            double tx = ntx_, ty = nty_;
            uint32_t tpack = npk_;
            const float2 a = nta_;
            if (t + 1 < NT) {
                const size_t i = (size_t)(t + 1) * bp + er;
                ntx_ = p.tgt_x[i]; nty_ = p.tgt_y[i]; npk_ = p.tgt_pack[i];
                nta_ = reinterpret_cast<const float2*>(p.tgt_act)[(size_t)er * NT + t + 1];
            }
            if (!((slow >> t) & 1)) {
                const int cap = tp_capacity(tpack);   // 1 or 2
                const double step_size = cap == 1 ? p.tgt_step_size : p.tgt_step_size * 0.5;
                double vx = (double)a.x, vy = (double)a.y;
                const double n2 = vx * vx + vy * vy;
                if (n2 > step_size * step_size * (1.0 - 1e-12)) {
                    const double n = sqrt(n2);
                    if (n > step_size) {   // Vector2D.norm setter (utils.py:223-229), see DESIGN.md
                        const double k = step_size / n;
                        vx *= k; vy *= k;
                    }
                }
                const double desx = tx + vx, desy = ty + vy;
                const double nx = fmin(fmax(desx, -kTerrain), kTerrain);
                const double ny = fmin(fmax(desy, -kTerrain), kTerrain);
                const int colliding = !(fabs(nx - desx) <= 1e-6 && fabs(ny - desy) <= 1e-6);
                tx = nx; ty = ny;
                tpack = (tpack & ~(1u << 27)) | ((uint32_t)colliding << 27);
                if (env_ok) { p.tgt_x[(size_t)t * bp + e] = tx; p.tgt_y[(size_t)t * bp + e] = ty; }
            }
            myval[S::V_T + 3 * t + 0] = (float)tx; myval[S::V_T + 3 * t + 1] = (float)ty;
            myval[S::V_T + 3 * t + 2] = __uint_as_float(tpack);
        }
        // exact re-simulation of the targets near a disc: queued and processed 32 at a time, one per lane
        __syncwarp();
        int count = 0;   // warp-uniform
        for (;;) {
            const bool more = __any_sync(FULL, slow != 0);
            if (more) {
                const bool has = slow != 0;
                const int t = has ? (__ffs(slow) - 1) : 0;
                slow &= slow - 1;
                const uint32_t ballot = __ballot_sync(FULL, has);
                if (has) queue[count + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)((lane << 8) | t);
                count += __popc(ballot);
                __syncwarp();
            }
            if (count >= 32 || (!more && count > 0)) {
                const int n = min(count, 32);
                count -= n;
                process_slow_targets<NC, NT, NO, S>(p, env0, val, queue, count, n);
                __syncwarp();
            }
            if (!more && count == 0) break;
        }
    }
    role_arrive(BAR_TGT_READY);   // the new target positions are in shared memory

    // ------------------------------------------------------------------ Sensor.perceive for the targets (entities.py:229-232)
    {
        const float fsr = (float)p.tgt_sight_range;
        const float fsr2 = fsr * fsr, fsrc = fsr + (float)p.cam_radius, fsrc2 = fsrc * fsrc;
        const double sr = p.tgt_sight_range, src = sr + p.cam_radius;
        float ftx[NT], fty[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) { ftx[t] = myval[S::V_T + 3 * t]; fty[t] = myval[S::V_T + 3 * t + 1]; }
        uint32_t trow[NT], trow2[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) { trow[t] = bit_tgt(t); trow2[t] = 0; }   // environment.py:1376-1377
        // fp32 on squares; inside a 4e-6 relative band the fp64 test decides (resolve_band, rare)
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            uint32_t band_t = 0, band_c = 0;
#pragma unroll
            for (int u = t + 1; u < NT; ++u) {   // symmetric
                const float dx = ftx[u] - ftx[t], dy = fty[u] - fty[t], d2 = dx * dx + dy * dy;
                if (d2 < fsr2 * (1.0f - 4e-6f)) { trow[t] |= bit_tgt(u); trow[u] |= bit_tgt(t); }
                else if (d2 <= fsr2 * (1.0f + 4e-6f)) band_t |= 1u << u;
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {       // target t senses camera c
                const float dx = mycam[CV * c] - ftx[t], dy = mycam[CV * c + 1] - fty[t], d2 = dx * dx + dy * dy;
                if (d2 < fsrc2 * (1.0f - 4e-6f)) trow[t] |= bit_cam(c);
                else if (d2 <= fsrc2 * (1.0f + 4e-6f)) band_c |= 1u << c;
            }
            if (band_t | band_c) {
                const double* ax = p.tgt_x + (size_t)t * bp + er;
                const double* ay = p.tgt_y + (size_t)t * bp + er;
                if (band_t) {
                    const uint32_t fix = resolve_band(p, er, ax, ay, band_t, 0, sr, false);
#pragma unroll
                    for (int u = t + 1; u < NT; ++u) if ((fix >> u) & 1) { trow[t] |= bit_tgt(u); trow[u] |= bit_tgt(t); }
                }
                if (band_c) trow[t] |= resolve_band(p, er, ax, ay, band_c, 1, src, false);   // bit_cam(c) == 1 << c
            }
        }
#pragma unroll (NO <= 12 ? 3 : 4)
        for (int o = 0; o < NO; ++o) {
            const float4 ob = p.obs_f4[(size_t)o * bp + er];
            const uint32_t obit = MW == 1 ? (1u << (16 + o)) : (1u << (o & 31));
            const float rtf = fsr + ob.z, rt2 = rtf * rtf;
            uint32_t band_t = 0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {       // target t senses obstacle o
                const float dx = ob.x - ftx[t], dy = ob.y - fty[t], d2 = dx * dx + dy * dy;
                if (d2 < rt2 * (1.0f - 4e-6f)) { if (MW == 1) trow[t] |= obit; else trow2[t] |= obit; }
                else if (d2 <= rt2 * (1.0f + 4e-6f)) band_t |= 1u << t;
            }
            if (band_t) {
                const size_t io = (size_t)o * bp + er;
                const uint32_t fix_t = resolve_band(p, er, p.obs_x + io, p.obs_y + io, band_t, 0, sr + p.obs_r[io], false);
#pragma unroll
                for (int t = 0; t < NT; ++t) if ((fix_t >> t) & 1) { if (MW == 1) trow[t] |= obit; else trow2[t] |= obit; }
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) { mymk[(NC + t) * MW] = trow[t]; if (MW == 2) mymk[(NC + t) * MW + 1] = trow2[t]; }
    }

    // ------------------------------------------------------------------ _assign_goals, the part that does not need the
    // tracked bits (environment.py:1282-1324): deliveries, pick-ups, empty bits.  A target's bounty keeps its old value
    // in the packed word; what the reward needs later is remembered in `events`: bit t = target t delivered, bit 8 + t =
    // its goal / cargo was cleared or replaced (the bounty restarts), bits 16 + 2 t = the weight it delivered.
    Cargo cargo;
    cargo.aw[0] = ea.x; cargo.aw[1] = ea.y;
    int episode_step = (int)ea.z, delivered = (int)ea.w;
    const int episode_id = eb.w;
    bool cargo_loaded = false, cargo_dirty = false;
    uint32_t events = 0, tdone_bits = 0;
    {
        const RngKey key{p.seed, (uint32_t)(p.env_index_base + e), (uint32_t)episode_id};
        const int draw_step = episode_step + 1;
        uint32_t in_bits = 0, whs = 0, old_goals = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const uint32_t tpack = __float_as_uint(myval[S::V_T + 3 * t + 2]);
            old_goals |= (uint32_t)(tp_goal(tpack) + 1) << (3 * t);
            // the four warehouses sit at (+-925, +-925): the one this target could be in is given by
            // the signs of its coordinates (constants.py:70-72 order: ++, -+, --, +-)
            const float fx = myval[S::V_T + 3 * t], fy = myval[S::V_T + 3 * t + 1];
            const float m = fmaxf(fabsf(fabsf(fx) - (float)kWarehouseCoord), fabsf(fabsf(fy) - (float)kWarehouseCoord));
            bool inside = m < (float)kWarehouseRadius - 1e-3f;
            if (!inside && m <= (float)kWarehouseRadius + 1e-3f) {
                const double tx = p.tgt_x[(size_t)t * bp + er], ty = p.tgt_y[(size_t)t * bp + er];
                inside = fmax(fabs(fabs(tx) - kWarehouseCoord), fabs(fabs(ty) - kWarehouseCoord)) <= kWarehouseRadius;
            }
            if (inside) {
                const int wq = (fy >= 0.0f) ? ((fx >= 0.0f) ? 0 : 1) : ((fx >= 0.0f) ? 3 : 2);
                in_bits |= 1u << t;
                whs |= (uint32_t)wq << (2 * t);
            }
        }
        // Sequential over the targets standing in a warehouse, ascending index, one per lane and iteration
        while (__any_sync(FULL, in_bits != 0)) {
            if (in_bits != 0) {
                if (!cargo_loaded) {
                    const uint4 c0 = p.cargo[er], c1 = p.cargo[bp + er];
                    cargo.rem[0] = c0.x; cargo.rem[1] = c0.y; cargo.rem[2] = c0.z; cargo.rem[3] = c0.w;
                    cargo.rem[4] = c1.x; cargo.rem[5] = c1.y; cargo.rem[6] = c1.z; cargo.rem[7] = c1.w;
                    cargo_loaded = true;
                }
                const int t = __ffs(in_bits) - 1;
                in_bits &= in_bits - 1;
                const int w = (whs >> (2 * t)) & 3;
                const uint32_t tp_t = __float_as_uint(myval[S::V_T + 3 * t + 2]);
                int goal = tp_goal(tp_t), weight = tp_weight(tp_t);
                const int capacity = tp_capacity(tp_t);
                int empty = tp_empty(tp_t);
                bool proceed = true;
                if (goal >= 0) {
                    if (goal == w) {
                        events |= (1u << t) | ((uint32_t)weight << (16 + 2 * t));
                        delivered += weight;
                        cargo.awaiting_add(goal, -weight);
                    } else {
                        proceed = false;
                    }
                }
                if (proceed) {
                    events |= 1u << (8 + t);
                    weight = 0; goal = -1;
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
                        weight = min(capacity, cargo.get(w, new_goal));
                        cargo.add(w, new_goal, -weight);
                        goal = new_goal;
                    }
                    cargo_dirty = true;
                }
                // empty_bits for the warehouse the target stands in (environment.py:1317-1318)
                empty = cargo.row_any(w) ? (empty & ~(1 << w)) : (empty | (1 << w));
                myval[S::V_T + 3 * t + 2] = __uint_as_float(pack_target(tp_bounty(tp_t), goal, weight, capacity, empty, tp_colliding(tp_t)));
            }
        }
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int old_goal = (int)((old_goals >> (3 * t)) & 7u) - 1;
            const int goal = tp_goal(__float_as_uint(myval[S::V_T + 3 * t + 2]));
            tdone_bits |= (uint32_t)((goal != old_goal) && (old_goal >= 0)) << t;
        }
    }
    episode_step += 1;
    const int done = !(episode_step <= p.max_episode_steps && cargo.any_awaiting());   // environment.py:628-632
    const bool auto_reset_needed = env_ok && done && (p.flags & MATE_STEP_AUTO_RESET);
    const uint32_t late = __ballot_sync(FULL, auto_reset_needed);
    if (lane == 0) xch[0] = late;
    role_arrive(BAR_GOALS_READY);   // goals / cargo weights of the targets are final, and so is the set of ending episodes

    // ------------------------------------------------------------------ target rows (environment.py:947-983)
    if (!(MATE3_X & 5)) pack_rows<NC, NT, NO, NC, S::R>(p, env0, valid_bits & ~late, stage, mk, val, MATE3_PACE_T);

    // ------------------------------------------------------------------ rewards, episode bookkeeping (environment.py:614-632, 1275-1296)
    role_sync(BAR_CAM_MASKS);       // the camera rows of the mask block are final
    if ((MATE3_X & 4) && !(MATE3_X & 1)) pack_rows<NC, NT, NO, NC, S::R>(p, env0, valid_bits & ~late, stage, mk, val);
    {
        uint32_t tracked_bits = 0;
        {
            uint32_t any_c = 0;
#pragma unroll
            for (int c = 0; c < NC; ++c) any_c |= mymk[c * MW];
            tracked_bits = (any_c >> 8) & 0xFFu;
        }
        int r = 0, delayed = 0;
        uint32_t wb_bits = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const uint32_t tp = __float_as_uint(myval[S::V_T + 3 * t + 2]);
            int bounty = tp_bounty(tp);
            const int tracked = (tracked_bits >> t) & 1;
            if (tracked && bounty > 0) r -= 1;
            bounty = max(bounty - tracked, 0);
            if ((events >> t) & 1u) {
                const int weight = (int)((events >> (16 + 2 * t)) & 3u);
                const int reward = weight * p.freight_scale + bounty;
                r += reward;
                delayed += reward - (weight * p.bounty_scale - bounty);
            }
            if ((events >> (8 + t)) & 1u) bounty = tp_weight(tp) * p.bounty_scale;
            wb_bits |= (uint32_t)(bounty > 0) << t;
            if (env_ok) p.tgt_pack[(size_t)t * bp + e] = (tp & ~0xFFFFu) | (uint32_t)bounty;
        }
        // coverage statistics of the current view (environment.py:966-972)
        const int nwb = __popc(wb_bits);
        const float cov_now = (float)__popc(tracked_bits) / (float)NT;
        const float cov_real = nwb > 0 ? (float)__popc(wb_bits & tracked_bits) / (float)nwb : 0.f;
        const int ep_reward = eb.x + r, delayed_ep_reward = eb.y + delayed;
        const float coverage_sum = __int_as_float(eb.z) + cov_now;
        if (env_ok) {
            const int r_out = p.reward_sparse ? delayed : r;
            reinterpret_cast<float2*>(p.rewards)[e] = make_float2(-(float)r_out, (float)r_out);
            p.done[e] = (uint8_t)done;
            if (done) {
                atomicAdd(&p.stats[0], 1.0f);
                atomicAdd(&p.stats[1], (float)ep_reward);
                atomicAdd(&p.stats[2], (float)episode_step);
                atomicAdd(&p.stats[3], (float)delivered);
                atomicAdd(&p.stats[4], coverage_sum / (float)episode_step);
            }
            if (!done && p.next != nullptr && (p.flags & MATE_STEP_AUTO_RESET) && episode_step == p.max_episode_steps)
                prefetch_prepared<NC, NT, NO, S>(*p.next, e);   // the next step ends this episode (time limit)
            // aux reflects the step just taken (before any auto-reset)
            if (p.has_aux) {
                const float transport = delivered > 0 ? (float)((double)delayed_ep_reward / ((double)p.reward_scale * (double)delivered)) : 0.f;
                if (!p.has_aux_detail) {   // the common case: only the info-dict scalars (environment.py:634-639)
                    if (p.aux.coverage) {
                        p.aux.coverage[(size_t)e * 3 + 0] = cov_now;
                        p.aux.coverage[(size_t)e * 3 + 1] = cov_real;
                        p.aux.coverage[(size_t)e * 3 + 2] = transport;
                    }
                    if (p.aux.num_delivered) p.aux.num_delivered[e] = delivered;
                    if (p.aux.episode_step) p.aux.episode_step[e] = episode_step;
                } else {
                    write_aux_env<NC, NT, NO, S>(p, e, mymk, myval, tdone_bits, cov_now, cov_real, transport, delivered, episode_step);
                }
            }
            // write state back
            if (cargo_dirty) {
                p.cargo[e] = make_uint4(cargo.rem[0], cargo.rem[1], cargo.rem[2], cargo.rem[3]);
                p.cargo[bp + e] = make_uint4(cargo.rem[4], cargo.rem[5], cargo.rem[6], cargo.rem[7]);
            }
            p.env_a[e] = make_uint4(cargo.aw[0], cargo.aw[1], (uint32_t)episode_step, (uint32_t)delivered);
            p.env_b[e] = make_int4(ep_reward, delayed_ep_reward, __float_as_int(coverage_sum), episode_id);
        }
        if (lane == 0) atomicAdd(&p.stats[5], (float)nvalid);
    }

    // ------------------------------------------------------------------ episodes that ended (rare)
    role_sync(BAR_TILE_DONE);       // the camera role is done with the tile
    if (late != 0u) late_resets<NC, NT, NO>(p, tb, env0, late);
}

// =============================================================================================
// CAMERA role
// =============================================================================================
template <int NC, int NT, int NO>
__device__ __noinline__ void camera_role(const Params& p, unsigned char* tb, const int env0) {
    using S = Shape3<NC, NT, NO>;
    constexpr int MW = S::MW, CV = S::CV;
    constexpr int NCX = NC > 0 ? NC : 1;
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float* stage = reinterpret_cast<float*>(tb + S::OFF_STAGE);
    uint32_t* mk = reinterpret_cast<uint32_t*>(tb + S::OFF_MASK);
    float* val = reinterpret_cast<float*>(tb + S::OFF_VAL);
    uint16_t* queue = reinterpret_cast<uint16_t*>(tb + S::OFF_Q);
    uint16_t* queue2 = reinterpret_cast<uint16_t*>(tb + S::OFF_Q2);
    const uint32_t* xch = reinterpret_cast<const uint32_t*>(tb + S::OFF_X);
    uint32_t* mymk = mk + lane * S::MSTRIDE;
    float* myval = val + lane * S::VSTRIDE;
    float* mycam = myval + S::V_C;

    const int e = env0 + lane;
    const bool env_ok = e < p.num_envs;
    const int er = env_ok ? e : p.num_envs - 1;
    const size_t bp = p.bpad;
    const int nvalid = min(32, p.num_envs - env0);
    const uint32_t valid_bits = nvalid >= 32 ? FULL : ((1u << nvalid) - 1u);

    // key of this environment's transmittance draws in this step
    const int draw_step = (int)p.env_a[er].z + 1;
    const int episode_id = p.env_b[er].w;
    unsigned long long ccw = NC >= 2 ? p.cc_clear[er] : 0ull;   // static camera <-> camera lines of sight (per episode)

    // ------------------------------------------------------------------ Camera.simulate (entities.py:347-360)
    {
        double nx_ = p.cam_x[er], ny_ = p.cam_y[er], nphi_ = p.cam_phi[er], nth_ = p.cam_theta[er];
        float2 na_ = reinterpret_cast<const float2*>(p.cam_act)[(size_t)er * NC];
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {   // the next camera's state is fetched while this one is derived
            const double x = nx_, y = ny_;
            double phi = nphi_, theta = nth_;
            const float2 a = na_;
            if (c + 1 < NC) {
                const size_t i = (size_t)(c + 1) * bp + er;
                nx_ = p.cam_x[i]; ny_ = p.cam_y[i]; nphi_ = p.cam_phi[i]; nth_ = p.cam_theta[i];
                na_ = reinterpret_cast<const float2*>(p.cam_act)[(size_t)er * NC + c + 1];
            }
            const double da = fmin(fmax((double)a.x, -p.cam_rot_step), p.cam_rot_step);
            const double dv = fmin(fmax((double)a.y, -p.cam_zoom_step), p.cam_zoom_step);
            phi = normalize_angle(phi + da);
            theta = fmin(fmax(theta + dv, p.cam_min_view), 180.0);
            if (env_ok) { p.cam_phi[(size_t)c * bp + e] = phi; p.cam_theta[(size_t)c * bp + e] = theta; }
            store_camera(mycam + CV * c, x, y, phi, theta, p.cam_area_product);
        }
    }
    role_arrive(BAR_CAM_READY);

    // ------------------------------------------------------------------ the part of _update_view that needs no target
    float fcx[NCX], fcy[NCX];
#pragma unroll
    for (int c = 0; c < NC; ++c) { fcx[c] = mycam[CV * c]; fcy[c] = mycam[CV * c + 1]; }
    uint32_t crow[NCX], crow2[NCX];
#pragma unroll
    for (int c = 0; c < NC; ++c) { crow[c] = bit_cam(c); crow2[c] = 0; }   // environment.py:1383-1384
#pragma unroll (NO <= 12 ? 3 : 4)
    for (int o = 0; o < NO; ++o) {       // camera c has obstacle o in its set (entities.py:363-368, strict <)
        const float4 ob = p.obs_f4[(size_t)o * bp + er];
        const uint32_t obit = MW == 1 ? (1u << (16 + o)) : (1u << (o & 31));
        const float rcf = (float)p.cam_rmax + ob.z, rc2 = rcf * rcf;
        uint32_t band_c = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float dx = ob.x - fcx[c], dy = ob.y - fcy[c], d2 = dx * dx + dy * dy;
            if (d2 < rc2 * (1.0f - 4e-6f)) { if (MW == 1) crow[c] |= obit; else crow2[c] |= obit; }
            else if (d2 <= rc2 * (1.0f + 4e-6f)) band_c |= 1u << c;
        }
        if (band_c) {
            const size_t io = (size_t)o * bp + er;
            const uint32_t fix_c = resolve_band(p, er, p.obs_x + io, p.obs_y + io, band_c, 1, p.cam_rmax + p.obs_r[io], true);
#pragma unroll
            for (int c = 0; c < NC; ++c) if ((fix_c >> c) & 1) { if (MW == 1) crow[c] |= obit; else crow2[c] |= obit; }
        }
    }
    // camera -> camera: range + sector now, the occlusion part is static within an episode and cached in `ccw`
    if (NC >= 2 && (ccw >> 63) == 0ull) {
        ccw = build_cc_cache<NC, NO>(p, er);
        if (env_ok) p.cc_clear[e] = ccw;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const float* cv = mycam + CV * c;
        const float cs = cv[3], sn = cv[4], rs2 = cs * cs + sn * sn;   // heading scaled by Rs
        const float ch = cospif(cv[2] * (1.0f / 360.0f)), ch2 = ch * ch;
        uint32_t reach_c = 0, band_c = 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            if (j == c) continue;
            const int reach = fov_reach32(fcx[c], fcy[c], rs2, cs, sn, ch2, fcx[j], fcy[j]);
            reach_c |= (uint32_t)(reach == 1) << j;
            band_c |= (uint32_t)(reach == 2) << j;
        }
        if (band_c) reach_c |= resolve_fov_band(p, er, c, band_c, 1);
        // bits 8 j + c of ccw, j = 0..NC-1: camera c has a clear line of sight to camera j
        uint32_t clear_c = 0;
#pragma unroll
        for (int j = 0; j < NC; ++j) clear_c |= (uint32_t)((ccw >> (8 * j + c)) & 1ull) << j;
        crow[c] |= reach_c & clear_c;   // bit_cam(j) == 1 << j
    }

    // ------------------------------------------------------------------ Camera.perceive for the targets (entities.py:491-505)
    role_sync(BAR_TGT_READY);
    unsigned long long pend = 0ull;     // bit c * NT + t: camera c reaches target t (range + sector)
    {
        float ftx[NT], fty[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) { ftx[t] = myval[S::V_T + 3 * t]; fty[t] = myval[S::V_T + 3 * t + 1]; }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float* cv = mycam + CV * c;
            const float cs = cv[3], sn = cv[4], rs2 = cs * cs + sn * sn;
            const float ch = cospif(cv[2] * (1.0f / 360.0f)), ch2 = ch * ch;
            uint32_t reach_t = 0, band_t = 0;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int reach = fov_reach32(fcx[c], fcy[c], rs2, cs, sn, ch2, ftx[t], fty[t]);
                reach_t |= (uint32_t)(reach == 1) << t;
                band_t |= (uint32_t)(reach == 2) << t;
            }
            if (band_t) reach_t |= resolve_fov_band(p, er, c, band_t, 0);
            pend |= (unsigned long long)reach_t << (c * NT);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) { mymk[c * MW] = crow[c]; if (MW == 2) mymk[c * MW + 1] = crow2[c]; }
    __syncwarp();
    // then the stochastic transmittance draw and the occlusion test (entities.py:503-505): all pending (camera,
    // target) pairs of the tile go through a queue in shared memory and are evaluated 32 at a time, one pair per
    // lane; pairs the conservative occlusion classification cannot decide go through a second queue to the exact
    // polyline.
    {
        int count = 0, count2 = 0;   // warp-uniform
        for (;;) {
            const bool more = __any_sync(FULL, pend != 0ull);
            if (more) {
                const bool has = pend != 0ull;
                const int b = has ? (__ffsll((long long)pend) - 1) : 0;
                pend &= pend - 1ull;
                const uint32_t ballot = __ballot_sync(FULL, has);
                const int pos = count + __popc(ballot & ((1u << lane) - 1u));
                if (has) queue[pos] = (uint16_t)((lane << 8) | b);
                count += __popc(ballot);
                __syncwarp();
            }
            if (count >= 32 || (!more && count > 0)) {
                const int n = min(count, 32);
                count -= n;
                const bool has = lane < n;
                const uint32_t item = has ? queue[count + lane] : 0u;
                const int src = item >> 8, b = item & 0xFF;
                const int c = b / NT, t = b - c * NT;
                const uint32_t src_episode = __shfl_sync(FULL, (uint32_t)episode_id, src);
                const int src_draw = __shfl_sync(FULL, draw_step, src);
                bool need_exact = false;
                if (has) {
                    const int env = env0 + src;
                    const int envr = min(env, p.num_envs - 1);
                    bool sees;
                    if (p.replay_transmit) {
                        sees = p.replay_transmit[((size_t)envr * NC + c) * NT + t] != 0;
                    } else {
                        const RngKey k{p.seed, (uint32_t)(p.env_index_base + env), src_episode};
                        sees = rng_u01(k, STREAM_TRANSMIT, (uint32_t)src_draw * (uint32_t)(NC * NT) + (uint32_t)(c * NT + t)) < p.transmittance;
                    }
                    if (!sees) {
                        if (NO == 0 || p.transmittance_is_one) {
                            sees = true;   // polyline is the flat max_sight_range circle; dist <= rs <= Rmax
                        } else {
                            const float* v = val + src * S::VSTRIDE;
                            const float cx = v[S::V_C + CV * c], cy = v[S::V_C + CV * c + 1];
                            const int fast = occlusion_fast<NO>(p.obs_f4 + envr, bp, cx, cy, v[S::V_T + 3 * t] - cx, v[S::V_T + 3 * t + 1] - cy, (float)p.cam_rmax);
                            sees = fast == 1;
                            need_exact = fast == 2;
                            if (MATE2_PF_OBS == 1 && need_exact) prefetch_discs64<NO>(p, envr);
                        }
                    }
                    if (sees) atomicOr(&mk[src * S::MSTRIDE + c * MW], bit_tgt(t));
                }
                const uint32_t ballot2 = __ballot_sync(FULL, need_exact);
                if (need_exact) queue2[count2 + __popc(ballot2 & ((1u << lane) - 1u))] = (uint16_t)item;
                count2 += __popc(ballot2);
                __syncwarp();
            }
            if (count2 >= 32 || (!more && count == 0 && count2 > 0)) {
                const int n = min(count2, 32);
                count2 -= n;
                process_exact<NC, NT, NO, S>(p, env0, mk, queue2, count2, n);
                __syncwarp();
            }
            if (!more && count == 0 && count2 == 0) break;
        }
    }
    role_arrive(BAR_CAM_MASKS);     // the camera rows of the mask block are final

    // ------------------------------------------------------------------ camera rows (environment.py:936-945, 952-983)
    role_sync(BAR_GOALS_READY);     // the `loaded` flags of the targets' public states are final
    const uint32_t late = xch[0];
    if (!(MATE3_X & 2)) pack_rows<NC, NT, NO, 0, NC>(p, env0, valid_bits & ~late, stage, mk, val, MATE3_PACE_C);
    role_arrive(BAR_TILE_DONE);
}

// =============================================================================================
// The role-specialised step kernel (MODE_STEP only; reset / observe / prepare stay with mate_step_kernel2)
// =============================================================================================
template <int NC, int NT, int NO>
__global__ void __launch_bounds__(64, MATE3_MIN_CTAS)
mate_step_kernel3(const __grid_constant__ Params p) {
    using S = Shape3<NC, NT, NO>;
    static_assert(NC >= 1 && NC <= 8 && NT <= 8 && NO <= 32, "role kernel: at least one camera; mask layout as in Shape2");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int role = threadIdx.x >> 5;   // warp-uniform
#pragma unroll 1
    for (int k = 0; k < S::TPC; ++k) {
        const int env0 = (blockIdx.x * S::TPC + k) * 32;
        if (env0 >= p.num_envs) break;   // uniform over the CTA
        if (role == 0) target_role<NC, NT, NO>(p, smem_raw, env0);
        else camera_role<NC, NT, NO>(p, smem_raw, env0);
        if (k + 1 < S::TPC) __syncthreads();   // the tile block is reused
    }
}

}  // namespace mate
