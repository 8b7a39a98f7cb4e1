// mate_agents.cuh -- batched rule-based opponents for the single-team wrappers.
//
// GreedyTargetAgent (mate/agents/greedy.py:235-365), driven like MultiCamera drives its opponents
// (mate/wrappers/single_team.py:79-92, 261-279: observe -> communicate -> act), for all targets of all
// environments in one launch.  One thread = one environment: the team's message exchange (broadcast of the
// remembered non-empty warehouses, greedy.py:338-365, routed to every teammate, environment.py:1249-1269) is an
// AND over at most 8 four-bit sets and stays in registers.  The agents read their private state from the
// simulator's state arrays (not from the observation tensors, which observation wrappers may have transformed).
#pragma once

#include "mate_common.cuh"

namespace mate {

constexpr uint32_t STREAM_AGENT_BINOMIAL = 16, STREAM_AGENT_SAMPLE = 17, STREAM_AGENT_CHOICE = 18, STREAM_AGENT_RESET = 19;
constexpr int kAgentMemory = 6;   // per target: goal, non-empty warehouse set, previous x, y, previous noise x, y

__global__ void greedy_target_kernel(const Params p, const int nt, double* __restrict__ memory, const uint8_t* __restrict__ reset_mask,
                                     const double noise_scale, const unsigned long long seed, const unsigned long long serial,
                                     const MateAgentReplay replay, float* __restrict__ tgt_act) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.num_envs) return;
    const size_t bp = p.bpad;
    const bool reset = reset_mask != nullptr && reset_mask[e] != 0;
    const RngKey key{seed, (uint32_t)(p.env_index_base + e), 0x41474E54u /* 'AGNT' */};
    const uint32_t draw = (uint32_t)serial * 8u;
    // observe -> process_messages (greedy.py:330-336), and the sets that are broadcast
    uint32_t sent_and = 15u;
    for (int t = 0; t < nt; ++t) {
        double* m = memory + ((size_t)e * nt + t) * kAgentMemory;
        const uint32_t tp = p.tgt_pack[(size_t)t * bp + e];
        if (reset) {   // reset(observation), greedy.py:265-277
            const double x = p.tgt_x[(size_t)t * bp + e], y = p.tgt_y[(size_t)t * bp + e];
            const double step_size = p.tgt_step_size / (double)tp_capacity(tp);
            double sx, sy;
            if (replay.reset_sample) { sx = replay.reset_sample[((size_t)e * nt + t) * 2]; sy = replay.reset_sample[((size_t)e * nt + t) * 2 + 1]; }
            else {
                sx = (2.0 * rng_u01(key, STREAM_AGENT_RESET, draw + (uint32_t)t * 2u) - 1.0) * step_size;
                sy = (2.0 * rng_u01(key, STREAM_AGENT_RESET, draw + (uint32_t)t * 2u + 1u) - 1.0) * step_size;
            }
            m[0] = (tp_goal(tp) >= 0 && tp_weight(tp) > 0) ? (double)tp_goal(tp) : -1.0;
            m[1] = 15.0; m[2] = x; m[3] = y; m[4] = 0.5 * sx; m[5] = 0.5 * sy;
        }
        uint32_t non_empty = (uint32_t)m[1];
        const uint32_t seen = (uint32_t)tp_empty(tp);
        if (seen & non_empty) {
            non_empty &= ~seen;
            m[1] = (double)non_empty;
            sent_and &= non_empty;
        }
    }
    // receive_responses (greedy.py:355-365) + act (greedy.py:289-328)
    for (int t = 0; t < nt; ++t) {
        double* m = memory + ((size_t)e * nt + t) * kAgentMemory;
        const uint32_t tp = p.tgt_pack[(size_t)t * bp + e];
        const double x = p.tgt_x[(size_t)t * bp + e], y = p.tgt_y[(size_t)t * bp + e];
        const double step_size = p.tgt_step_size / (double)tp_capacity(tp);
        const bool state_has_goal = tp_goal(tp) >= 0 && tp_weight(tp) > 0;
        const uint32_t non_empty = (uint32_t)m[1] & sent_and;
        int goal = (int)m[0];
        if (state_has_goal) goal = tp_goal(tp);
        if (goal < 0 || (!state_has_goal && !((non_empty >> goal) & 1u))) {
            goal = -1;
            if (non_empty != 0u) {   // np_random.choice(list(non_empty_warehouses)): uniform over the set, ascending order
                if (replay.choice) goal = replay.choice[(size_t)e * nt + t];
                else {
                    int pick = (int)rng_below(key, STREAM_AGENT_CHOICE, draw + (uint32_t)t, (uint32_t)__popc(non_empty));
                    for (int w = 0; w < NW; ++w) if ((non_empty >> w) & 1u) { if (pick == 0) { goal = w; break; } --pick; }
                }
                goal = min(max(goal, 0), NW - 1);
            }
        }
        const double pdx = x - m[2], pdy = y - m[3];
        double ax = 0.0, ay = 0.0;
        if (goal >= 0) {
            ax = ((goal == 0 || goal == 3) ? kWarehouseCoord : -kWarehouseCoord) - x;
            ay = ((goal < 2) ? kWarehouseCoord : -kWarehouseCoord) - y;
        }
        const double norm = sqrt(ax * ax + ay * ay);
        if (norm > step_size) { const double k = step_size / norm; ax *= k; ay *= k; }
        const double prob = sqrt(pdx * pdx + pdy * pdy) > 0.2 * step_size ? 0.05 : 0.75;
        bool redraw;
        if (replay.binomial) redraw = replay.binomial[(size_t)e * nt + t] != 0;
        else redraw = rng_u01(key, STREAM_AGENT_BINOMIAL, draw + (uint32_t)t) < prob;
        double nx = m[4], ny = m[5];
        if (redraw) {   // noise_scale * action_space.sample(): uniform over the agent's own action box
            double sx, sy;
            if (replay.sample) { sx = replay.sample[((size_t)e * nt + t) * 2]; sy = replay.sample[((size_t)e * nt + t) * 2 + 1]; }
            else {
                sx = (2.0 * rng_u01(key, STREAM_AGENT_SAMPLE, draw + (uint32_t)t * 2u) - 1.0) * step_size;
                sy = (2.0 * rng_u01(key, STREAM_AGENT_SAMPLE, draw + (uint32_t)t * 2u + 1u) - 1.0) * step_size;
            }
            nx = noise_scale * sx; ny = noise_scale * sy;
        }
        ax = fmin(fmax(ax + nx, -step_size), step_size);
        ay = fmin(fmax(ay + ny, -step_size), step_size);
        reinterpret_cast<float2*>(tgt_act)[(size_t)e * nt + t] = make_float2((float)ax, (float)ay);
        m[0] = (double)goal; m[1] = (double)non_empty; m[2] = x; m[3] = y; m[4] = nx; m[5] = ny;
    }
}

}  // namespace mate
