#!/usr/bin/env python
"""Golden fixtures for the batched opponent agents -- TEST INFRASTRUCTURE ONLY.

Runs the unmodified reference (``/root/reference`` through ``oracle/gymshim``): greedy cameras against
``GreedyTargetAgent`` targets (mate/agents/greedy.py:235-365) driven exactly like ``MultiCamera`` drives its
opponents (mate/wrappers/single_team.py:79-92, 261-279: observe -> communicate -> act).  For every step it records
the simulator state the agents acted on, the memory of every target agent BEFORE the step (goal, remembered
non-empty warehouses, previous location, previous noise), the agent's stochastic draws (binomial, uniform sample,
choice) and the joint action plus the memory AFTER the step.

    python oracle/gen_agent_golden.py        # writes tests/golden/agents_*.npz
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  pylint: disable=wrong-import-position

REPO = os.path.dirname(HERE)

CONFIGS = [
    # name, config, seed, steps
    ('agents_4v8-9', 'MATE-4v8-9.yaml', 41, 700),
    ('agents_Navigation', 'MATE-Navigation.yaml', 42, 700),
    ('agents_4v2-0', 'MATE-4v2-0.yaml', 43, 400),
]


class DrawLog:
    """Delegating proxy around an agent's RandomState that logs binomial() and choice() results."""

    def __init__(self, rng):
        self._rng = rng
        self.binomial_out = None
        self.choice_out = None
        self.randint_out = None

    def __getattr__(self, name):
        return getattr(self._rng, name)

    def binomial(self, n, p, *args, **kwargs):
        out = self._rng.binomial(n, p, *args, **kwargs)
        self.binomial_out = int(out)
        return out

    def choice(self, a, *args, **kwargs):
        out = self._rng.choice(a, *args, **kwargs)
        self.choice_out = int(out)
        return out

    def randint(self, *args, **kwargs):
        out = self._rng.randint(*args, **kwargs)
        if self.randint_out is not None:
            self.randint_out.append(int(out))
        return out


def instrument(agent):
    log = DrawLog(agent.np_random)
    agent._np_random = log  # pylint: disable=protected-access
    samples = []
    original = agent.action_space.sample

    def sample():
        out = original()
        samples.append(np.asarray(out, dtype=np.float64))
        return out

    agent.action_space.sample = sample
    return log, samples


def memory(agent):
    goal = agent.goal
    non_empty = sum(1 << w for w in agent.non_empty_warehouses)
    return (np.int64(-1 if goal is None else goal), np.int64(non_empty), np.asarray(agent.prev_state.location, dtype=np.float64),
            np.asarray(agent.prev_noise, dtype=np.float64))


def camera_memory(agent, nc, nt):
    """GreedyCameraAgent's memory as arrays (mate/agents/greedy.py:21-66)."""
    mem = np.array([[ts.state[0], ts.state[1], ts.state[2], ts.state[3]] for ts in agent.memory], dtype=np.float64).reshape(nt, 4)
    neighbors = sum(1 << c for c in agent.neighboring_teammate_states)
    return {
        'cam_mem': mem, 'cam_time2forget': np.asarray(agent.time2forget, dtype=np.int64), 'cam_never_loaded': np.asarray(agent.never_loaded, dtype=np.uint8),
        'cam_prev_action': np.asarray(agent.prev_action, dtype=np.float64).reshape(2),
        'cam_delay': np.asarray(agent.communication_delay, dtype=np.int64).reshape(nc), 'cam_neighbors': np.int64(neighbors),
        'cam_has_state_message': np.uint8('state' in agent.message2send),
    }


def run(mate, name, config, seed, num_steps, out_dir):
    from mate.wrappers.single_team import group_reset, group_step  # pylint: disable=import-outside-toplevel

    env = mate.make('MultiAgentTracking-v0', config=config)
    u = env.unwrapped
    env.seed(seed)
    nc, nt = u.num_cameras, u.num_targets
    cam_obs, tgt_obs = env.reset()
    camera_agents = mate.GreedyCameraAgent(seed=seed).spawn(nc) if nc else []
    target_agents = mate.GreedyTargetAgent(seed=seed + 1, noise_scale=0.5).spawn(nt)
    group_reset(camera_agents, cam_obs)
    group_reset(target_agents, tgt_obs)
    logs = [instrument(agent) for agent in target_agents]   # after reset: action_space exists now
    cam_logs = [instrument(agent) for agent in camera_agents]
    reset_noise = np.array([agent.prev_noise for agent in target_agents], dtype=np.float64)

    rows = {}

    def push(key, value):
        rows.setdefault(key, []).append(np.asarray(value))

    cam_infos = tgt_infos = None
    done, step = False, 0
    while not done and step < num_steps:
        for k, v in gg.dump_state(u).items():
            push(k, v)
        before = [memory(agent) for agent in target_agents]
        for log, samples in logs:
            log.binomial_out = log.choice_out = None
            samples.clear()
        if nc:
            push('mask_ct', np.array(u.camera_target_view_mask, dtype=np.uint8))
            cam_before = [camera_memory(agent, nc, nt) for agent in camera_agents]
            for log, samples in cam_logs:
                log.binomial_out = None
                log.randint_out = []
                samples.clear()
        cam_act = np.asarray(group_step(env, camera_agents, cam_obs, cam_infos), dtype=np.float64) if nc else np.zeros((0, 2))
        if nc:
            cam_after = [camera_memory(agent, nc, nt) for agent in camera_agents]
            for key in cam_before[0]:
                push(key + '_before', [m[key] for m in cam_before])
                push(key + '_after', [m[key] for m in cam_after])
            push('cam_draw_binomial', [-1 if log.binomial_out is None else log.binomial_out for log, _ in cam_logs])
            push('cam_draw_sample', [samples[0] if samples else np.zeros(2) for _, samples in cam_logs])
            # one randint per message sent, in recipient order: recover the recipients from the delays that changed
            delays = np.full((nc, nc), -1, dtype=np.int64)
            for c, ((log, _), before_c, after_c) in enumerate(zip(cam_logs, cam_before, cam_after)):
                sent_to = [k for k in range(nc) if k != c and max(before_c['cam_delay'][k] - 1, 0) == 0 and after_c['cam_delay'][k] > 0]
                assert len(sent_to) == len(log.randint_out), (sent_to, log.randint_out)
                for k, value in zip(sent_to, log.randint_out):
                    assert after_c['cam_delay'][k] == value
                    delays[c, k] = value
            push('cam_draw_delay', delays)
            push('cam_act', cam_act.reshape(nc, 2))
        tgt_act = np.asarray(group_step(env, target_agents, tgt_obs, tgt_infos), dtype=np.float64)
        after = [memory(agent) for agent in target_agents]
        push('agent_goal_before', [m[0] for m in before])
        push('agent_non_empty_before', [m[1] for m in before])
        push('agent_prev_xy_before', [m[2] for m in before])
        push('agent_prev_noise_before', [m[3] for m in before])
        push('agent_goal_after', [m[0] for m in after])
        push('agent_non_empty_after', [m[1] for m in after])
        push('agent_prev_noise_after', [m[3] for m in after])
        push('draw_binomial', [log.binomial_out for log, _ in logs])
        push('draw_choice', [-1 if log.choice_out is None else log.choice_out for log, _ in logs])
        push('draw_sample', [samples[0] if samples else np.zeros(2) for _, samples in logs])
        assert all(len(samples) <= 1 for _, samples in logs)
        push('tgt_act', tgt_act.reshape(nt, 2))
        (cam_obs, tgt_obs), _, done, (cam_infos, tgt_infos) = env.step((gg.f32(cam_act).reshape(nc, 2), tgt_act.reshape(nt, 2)))
        step += 1

    out = {'config_name': np.array(config), 'seed': np.int64(seed), 'count': np.int64(step), 'noise_scale': np.float64(0.5),
           'reset_noise': reset_noise}
    out.update(gg.config_scalars(u))
    for k, v in rows.items():
        out['g_' + k] = np.stack(v)
    path = os.path.join(out_dir, name + '.npz')
    np.savez_compressed(path, **out)
    picks = int((out['g_draw_choice'] >= 0).sum())
    print(f'{name}: steps={step} delivered={u.num_delivered_cargoes} goal picks={picks} noise redraws={int(out["g_draw_binomial"].sum())} '
          f'size={os.path.getsize(path) / 1e6:.2f}MB')


def main():
    mate = gg._import_reference()  # pylint: disable=protected-access
    out_dir = os.path.join(REPO, 'tests', 'golden')
    for name, config, seed, steps in CONFIGS:
        run(mate, name, config, seed, steps, out_dir)


if __name__ == '__main__':
    main()
