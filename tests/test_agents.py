"""Batched opponents against recorded reference agents: ``tests/golden/agents_*.npz`` hold,
for every step of greedy-vs-greedy episodes, the simulator state the reference's ``GreedyTargetAgent`` team acted on,
every agent's memory before and after, its random draws and the joint action (``oracle/gen_agent_golden.py``)."""

import glob
import os

import numpy as np
import pytest

import golden_util as gu
from oracle.greedy_target_ref import team_step

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(gu.GOLDEN_DIR, 'agents_*.npz')))


def _inputs(g, i):
    nt = int(g['cfg_counts'][1])
    goal_bits = np.zeros((nt, 4))
    for t in range(nt):
        if g['g_tgt_goal'][i][t] >= 0:
            goal_bits[t, g['g_tgt_goal'][i][t]] = g['g_tgt_goal_weight'][i][t]
    memory = {'goal': g['g_agent_goal_before'][i], 'non_empty': g['g_agent_non_empty_before'][i],
              'prev_xy': g['g_agent_prev_xy_before'][i], 'prev_noise': g['g_agent_prev_noise_before'][i]}
    draws = {'binomial': g['g_draw_binomial'][i], 'sample': g['g_draw_sample'][i], 'choice': g['g_draw_choice'][i]}
    return goal_bits, memory, draws


@pytest.mark.parametrize('name', NAMES)
def test_restatement_matches_the_reference_agents(name):
    """CPU: the NumPy restatement (the checker of the CUDA kernel) reproduces actions and memory of the recorded
    reference agents at every step."""
    g = gu.load(name)
    step = float(g['cfg_target'][0])
    for i in range(int(g['count'])):
        goal_bits, memory, draws = _inputs(g, i)
        act, after = team_step(g['g_tgt_xy'][i], step / g['g_tgt_capacity'][i], goal_bits, g['g_tgt_empty_bits'][i].astype(bool),
                               memory, draws, float(g['noise_scale']))
        np.testing.assert_allclose(act, g['g_tgt_act'][i], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(after['goal'], g['g_agent_goal_after'][i])
        np.testing.assert_array_equal(after['non_empty'], g['g_agent_non_empty_after'][i])
        np.testing.assert_allclose(after['prev_noise'], g['g_agent_prev_noise_after'][i], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_greedy_target_kernel_matches_the_reference_agents(name):
    """GPU: every recorded step is one environment of a batch: simulator state and agent memory are injected, the
    recorded draws are replayed, and the joint action and the memory after the step are compared."""
    import torch

    import mate_b200
    from mate_b200.sim import BatchedSim

    g = gu.load(name)
    n = int(g['count'])
    sim = BatchedSim(gu.flat_config(g), n, device=0)
    sim.set_state(gu.stack_states([gu.state_arrays(g, 'g_', i) for i in range(n)]))
    agent = mate_b200.GreedyTargetAgent(seed=1, noise_scale=float(g['noise_scale']))
    agent.bind(sim)
    memory = np.concatenate([g['g_agent_goal_before'][..., None], g['g_agent_non_empty_before'][..., None],
                             g['g_agent_prev_xy_before'], g['g_agent_prev_noise_before']], axis=-1).astype(np.float64)
    agent.memory.copy_(torch.from_numpy(memory))
    actions = agent.act(replay={'binomial': g['g_draw_binomial'], 'sample': g['g_draw_sample'], 'choice': g['g_draw_choice']})
    torch.cuda.synchronize()
    np.testing.assert_allclose(actions.cpu().numpy(), g['g_tgt_act'], rtol=1e-6, atol=1e-5)
    after = agent.memory.cpu().numpy()
    np.testing.assert_array_equal(after[..., 0], g['g_agent_goal_after'])
    np.testing.assert_array_equal(after[..., 1], g['g_agent_non_empty_after'])
    np.testing.assert_allclose(after[..., 2:4], g['g_tgt_xy'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(after[..., 4:6], g['g_agent_prev_noise_after'], rtol=0, atol=1e-12)
    # reset(observation) (greedy.py:265-277) with the recorded sample: previous noise = 0.5 * sample
    agent.act(reset_mask=True, replay={'binomial': np.zeros_like(g['g_draw_binomial']), 'reset_sample': g['g_draw_sample']})
    torch.cuda.synchronize()
    after = agent.memory.cpu().numpy()
    np.testing.assert_allclose(after[..., 4:6], 0.5 * g['g_draw_sample'], rtol=0, atol=1e-12)
    expected_goal = np.where((g['g_tgt_goal'] >= 0) & (g['g_tgt_goal_weight'] > 0), g['g_tgt_goal'], -1)
    keeps = expected_goal >= 0
    np.testing.assert_array_equal(after[..., 0][keeps], expected_goal[keeps])
    sim.close()


CAMERA_NAMES = [n for n in NAMES if 'Navigation' not in n]


def _camera_inputs(g, i):
    nt = int(g['cfg_counts'][1])
    loaded = (g['g_tgt_goal'][i] >= 0) & (g['g_tgt_goal_weight'][i] > 0)
    tgt_state = np.c_[g['g_tgt_xy'][i], np.full(nt, g['cfg_target'][1]), loaded.astype(float)]
    mem = {'memory': g['g_cam_mem_before'][i], 'time2forget': g['g_cam_time2forget_before'][i],
           'never_loaded': g['g_cam_never_loaded_before'][i], 'prev_action': g['g_cam_prev_action_before'][i],
           'delay': g['g_cam_delay_before'][i], 'neighbors': g['g_cam_neighbors_before'][i],
           'has_state': g['g_cam_has_state_message_before'][i]}
    draws = {'binomial': g['g_cam_draw_binomial'][i], 'sample': g['g_cam_draw_sample'][i], 'delay': g['g_cam_draw_delay'][i]}
    return tgt_state, mem, draws


def _pack_camera_memory(g, suffix):
    """[n, Nc, 6 Nt + Nc + 4] in the layout of include/mate_b200.h."""
    n, nc, nt = g['g_cam_mem' + suffix].shape[:3]
    return np.concatenate([
        g['g_cam_mem' + suffix].reshape(n, nc, 4 * nt), g['g_cam_time2forget' + suffix], g['g_cam_never_loaded' + suffix],
        g['g_cam_prev_action' + suffix], g['g_cam_delay' + suffix], g['g_cam_neighbors' + suffix][..., None],
        g['g_cam_has_state_message' + suffix][..., None]], axis=-1).astype(np.float64)


@pytest.mark.parametrize('name', CAMERA_NAMES)
def test_camera_restatement_matches_the_reference_agents(name):
    """CPU: the NumPy restatement of GreedyCameraAgent reproduces actions, memory, delays and known teammates of the
    recorded reference agents at every step."""
    from oracle.greedy_camera_ref import team_step as camera_team_step

    g = gu.load(name)
    cam = g['cfg_camera']
    for i in range(int(g['count'])):
        tgt_state, mem, draws = _camera_inputs(g, i)
        act, after = camera_team_step(g['g_cam_xy'][i], g['g_cam_phi'][i], g['g_cam_theta'][i], tgt_state, g['g_mask_ct'][i].astype(bool),
                                      mem, draws, (cam[1], cam[2], cam[3], cam[4]))
        np.testing.assert_allclose(act, g['g_cam_act'][i], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(after['time2forget'], g['g_cam_time2forget_after'][i])
        np.testing.assert_array_equal(after['never_loaded'], g['g_cam_never_loaded_after'][i])
        np.testing.assert_array_equal(after['delay'], g['g_cam_delay_after'][i])
        np.testing.assert_array_equal(after['neighbors'], g['g_cam_neighbors_after'][i])
        np.testing.assert_allclose(after['memory'], g['g_cam_mem_after'][i], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('name', CAMERA_NAMES)
def test_greedy_camera_kernel_matches_the_reference_agents(name):
    """GPU: every recorded step is one environment of a batch (state and agent memory injected, draws replayed)."""
    import torch

    import mate_b200
    from mate_b200.sim import BatchedSim

    g = gu.load(name)
    n = int(g['count'])
    sim = BatchedSim(gu.flat_config(g), n, device=0)
    sim.set_state(gu.stack_states([gu.state_arrays(g, 'g_', i) for i in range(n)]))
    agent = mate_b200.GreedyCameraAgent(seed=1)
    agent.bind(sim)
    agent.memory.copy_(torch.from_numpy(_pack_camera_memory(g, '_before')))
    actions = agent.act(g['g_mask_ct'], replay={'binomial': g['g_cam_draw_binomial'], 'sample': g['g_cam_draw_sample'],
                                                'delay': g['g_cam_draw_delay']})
    torch.cuda.synchronize()
    np.testing.assert_allclose(actions.cpu().numpy(), g['g_cam_act'], rtol=1e-6, atol=1e-5)
    want = _pack_camera_memory(g, '_after')
    got = agent.memory.cpu().numpy()
    nt = int(g['cfg_counts'][1])
    np.testing.assert_allclose(got[..., :6 * nt], want[..., :6 * nt], rtol=0, atol=1e-12)          # memory, time2forget, never_loaded
    np.testing.assert_allclose(got[..., 6 * nt:6 * nt + 2], g['g_cam_act'], rtol=0, atol=1e-9)      # previous action
    np.testing.assert_array_equal(got[..., 6 * nt + 2:], want[..., 6 * nt + 2:])                   # delays, teammates, pending state
    # reset(observation), greedy.py:44-66
    agent.act(g['g_mask_ct'], reset_mask=True, replay={'binomial': np.zeros_like(g['g_cam_draw_binomial']), 'sample': g['g_cam_draw_sample'],
                                                      'delay': np.full_like(g['g_cam_draw_delay'], 7)})
    torch.cuda.synchronize()
    got = agent.memory.cpu().numpy()
    seen = g['g_mask_ct'].astype(bool)
    np.testing.assert_array_equal(got[..., 4 * nt:5 * nt], np.where(seen, 25.0, 0.0))
    assert (got[..., 6 * nt + 3 + seen.shape[1]] == 0).all()       # the state message went out in the first step
    assert (got[..., 6 * nt + 2 + seen.shape[1]] == (2 ** seen.shape[1] - 1) - 2 ** np.arange(seen.shape[1])).all()   # everyone knows everyone else
    sim.close()


@pytest.mark.gpu
def test_multi_target_with_live_greedy_cameras():
    """MultiTarget with the batched GreedyCameraAgent on its own Philox draws: target-side API shapes, camera actions
    inside the action box, the greedy cameras track (coverage well above what idle cameras get), reproducible."""
    import torch

    import mate_b200

    def run(agent_seed, idle=False):
        env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=64, wrappers=[
            lambda e: mate_b200.MultiTarget(e, camera_agent=mate_b200.GreedyCameraAgent(seed=agent_seed))])
        obs = env.reset(seed=3)
        assert obs.shape == (64, 8, env.unwrapped.sim.dt)
        rng = np.random.RandomState(0)
        coverage = 0.0
        for _ in range(300):
            act = torch.from_numpy((rng.uniform(-1, 1, (64, 8, 2)) * 20.0).astype(np.float32)).cuda()
            if idle:
                env.opponent_agent.act = lambda tracked, reset_mask=None, replay=None: torch.zeros((64, 4, 2), device='cuda')
            obs, reward, done, infos = env.step(act)
            cam_act = env.opponent_agent.actions
            assert float(cam_act[..., 0].abs().max()) <= 5.0 + 1e-5 and float(cam_act[..., 1].abs().max()) <= 2.5 + 1e-5
            coverage += float(infos['coverage_rate'].mean())
        assert reward.shape == (64,) and done.shape == (64,)
        env.unwrapped.close()
        return coverage / 300

    tracked_a, tracked_b, idle = run(5), run(5), run(5, idle=True)
    assert tracked_a == tracked_b
    assert tracked_a > idle + 0.05, (tracked_a, idle)


@pytest.mark.gpu
def test_multi_camera_with_live_greedy_targets():
    """MultiCamera with the batched GreedyTargetAgent on its own Philox draws: the camera-side API shapes, actions
    inside each target's action box, cargo gets delivered, and the run is reproducible from the seed."""
    import torch

    import mate_b200

    def run(seed):
        env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=64, wrappers=[
            mate_b200.RepeatedRewardIndividualDone,
            lambda e: mate_b200.MultiCamera(e, target_agent=mate_b200.GreedyTargetAgent(seed=seed))])
        obs = env.reset(seed=3)
        assert obs.shape == (64, 4, env.unwrapped.sim.dc)
        total = 0.0
        zeros = torch.zeros((64, 4, 2), device='cuda')
        for _ in range(400):
            obs, reward, done, infos = env.step(zeros)
            act = env.opponent_agent.actions
            limit = (env.unwrapped.target_step_size / env.unwrapped.sim.get_state()['tgt_capacity']) if _ == 0 else None
            if limit is not None:
                assert (act.abs().cpu().numpy() <= limit[..., None] + 1e-5).all()
            total += float(reward.sum())
        assert obs.shape == (64, 4, env.unwrapped.sim.dc) and reward.shape == (64, 4) and done.shape == (64, 4)
        delivered = float(env.unwrapped.num_delivered_cargoes.float().mean())
        env.unwrapped.close()
        return total, delivered

    total_a, delivered_a = run(5)
    total_b, delivered_b = run(5)
    assert total_a == total_b and delivered_a == delivered_b
    assert delivered_a > 1.0     # greedy targets haul cargo (the reference's agents deliver ~43 in 700 steps)


@pytest.mark.gpu
def test_single_team_wrappers_in_reference_compatible_mode():
    """num_envs=None: MultiCamera / MultiTarget keep the reference's single-env types."""
    import mate_b200

    env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v2-9.yaml',
                         wrappers=[lambda e: mate_b200.MultiCamera(e, target_agent=mate_b200.GreedyTargetAgent(seed=0))])
    obs = env.reset(seed=1)
    assert isinstance(obs, np.ndarray) and obs.shape == (4, 96)
    for _ in range(20):
        obs, reward, done, infos = env.step(np.zeros((4, 2)))
    assert isinstance(reward, float) and isinstance(done, bool) and len(infos) == 4 and obs.shape == (4, 96)
    env.unwrapped.close()
    env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v2-9.yaml',
                         wrappers=[lambda e: mate_b200.MultiTarget(e, camera_agent=mate_b200.GreedyCameraAgent(seed=0))])
    obs = env.reset(seed=1)
    assert isinstance(obs, np.ndarray) and obs.shape == (2, 101)
    for _ in range(20):
        obs, reward, done, infos = env.step(np.zeros((2, 2)))
    assert isinstance(reward, float) and isinstance(done, bool) and len(infos) == 2
    env.unwrapped.close()


@pytest.mark.gpu
def test_agent_draws_do_not_depend_on_the_batch_split():
    """Multi-GPU sharding: the agents' Philox streams are keyed on the GLOBAL environment index,
    so one batch of 128 environments and two shards of 64 (env_index_base 0 / 64) produce the same joint actions."""
    import torch

    import mate_b200
    from mate_b200.config import flatten_config, read_config
    from mate_b200.sim import BatchedSim

    cfg = flatten_config(read_config('MATE-4v8-9.yaml'))

    def run(num_envs, base):
        sim = BatchedSim(cfg, num_envs, device=0, env_index_base=base)
        sim.reset(seed=9)
        sim.observe(aux=True)
        targets = mate_b200.GreedyTargetAgent(seed=5)
        cameras = mate_b200.GreedyCameraAgent(seed=6)
        targets.bind(sim)
        cameras.bind(sim)
        out = []
        reset = True
        for _ in range(40):
            tracked = (sim.cam_obs[..., 26:22 + 5 * 8:5] > 0.5).to(torch.uint8)
            tgt_act = targets.act(reset_mask=reset).clone()
            cam_act = cameras.act(tracked, reset_mask=reset).clone()
            reset = None
            sim.step(cam_act, tgt_act, auto_reset=True, aux=True)
            out.append((cam_act.cpu().numpy(), tgt_act.cpu().numpy()))
        sim.close()
        return out

    whole, lo, hi = run(128, 0), run(64, 0), run(64, 64)
    for (cam_w, tgt_w), (cam_l, tgt_l), (cam_h, tgt_h) in zip(whole, lo, hi):
        np.testing.assert_array_equal(cam_w, np.concatenate([cam_l, cam_h]))
        np.testing.assert_array_equal(tgt_w, np.concatenate([tgt_l, tgt_h]))


@pytest.mark.gpu
def test_greedy_target_noise_draws_are_distinct_across_steps_and_targets():
    """The noise sample of target t at step s must be its own Philox draw: with 8 targets the per-step counter stride has
    to cover 2 * Nt indices (a stride of 8 made target t >= 4 at step s re-read the draw of target t - 4 at step s + 1).
    The remembered noise (memory[..., 4:6]) equals noise_scale * U(-1, 1) * step_size of the step that drew it."""
    import mate_b200
    from mate_b200.config import flatten_config, read_config
    from mate_b200.sim import BatchedSim

    cfg = flatten_config(read_config('MATE-4v8-9.yaml'))
    sim = BatchedSim(cfg, 32, device=0)
    sim.reset(seed=2)
    agent = mate_b200.GreedyTargetAgent(seed=11, noise_scale=1.0)
    agent.bind(sim)
    step_size = cfg['target_step_size'] / sim.get_state()['tgt_capacity']          # [B, Nt]
    zeros = np.zeros((32, 4, 2), dtype=np.float32)
    seen = {}
    prev = None
    reset = True
    for s in range(60):
        act = agent.act(reset_mask=reset).clone()
        reset = None
        noise = agent.memory[..., 4:6].cpu().numpy() / step_size[..., None]          # u in (-1, 1)
        for e in range(32):
            for t in range(8):
                if prev is not None and (noise[e, t] == prev[e, t]).all():
                    continue   # not redrawn in this step
                key = (e, round(float(noise[e, t, 0]), 12), round(float(noise[e, t, 1]), 12))
                assert key not in seen, f'env {e}: target {t} at step {s} repeats the draw of {seen[key]}'
                seen[key] = (t, s)
        prev = noise
        sim.step(zeros, act, auto_reset=False)
    assert len(seen) > 32 * 8 * 3   # the noise is redrawn often enough for the check to mean something
    sim.close()
