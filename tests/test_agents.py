"""Batched opponents against recorded reference agents: ``tests/golden/agents_*.npz`` hold,
for every step of greedy-vs-greedy episodes, the simulator state the reference's ``GreedyTargetAgent`` team acted on,
every agent's memory before and after, its random draws and the joint action (``oracle/gen_agent_golden.py``)."""

import glob
import os

import numpy as np
import pytest

import golden_util as gu
from greedy_target_ref import team_step

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
