"""The reference-facing Python interface (mate_b200.make / MultiAgentTracking) on the GPU."""

import numpy as np
import pytest

import golden_util as gu

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def test_make_batched_env():
    import mate_b200 as mate

    env = mate.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=512)
    assert isinstance(env, mate.MultiAgentTracking)
    assert (env.num_cameras, env.num_targets, env.num_obstacles) == (4, 8, 9)
    assert str(env) == '<MultiAgentTracking<MultiAgentTracking-v0>>(4 cameras, 8 targets, 9 obstacles, 512 envs)'
    cam_obs, tgt_obs = env.reset(seed=3)
    assert cam_obs.shape == (512, 4, 126) and tgt_obs.shape == (512, 8, 131) and cam_obs.is_cuda
    assert env.camera_observation_space.shape == (126,) and env.target_observation_space.shape == (131,)
    assert env.state_space.shape == (13 + 9 * 4 + 14 * 8 + 3 * 9 + 8 + 8 + 16,)
    cam_act = torch.zeros((512, 4, 2), device='cuda')
    tgt_act = torch.ones((512, 8, 2), device='cuda')
    (cam_obs, tgt_obs), (cam_r, tgt_r), done, (cam_info, tgt_info) = env.step((cam_act, tgt_act))
    assert cam_r.shape == (512,) and done.dtype == torch.bool and torch.equal(cam_r, -tgt_r)
    assert set(cam_info) >= {'raw_reward', 'normalized_raw_reward', 'coverage_rate', 'real_coverage_rate',
                             'mean_transport_rate', 'num_delivered_cargoes'}
    assert env.camera_target_view_mask.shape == (512, 4, 8)
    assert env.state().shape == (512, env.state_space.shape[0])
    # mask attributes agree with the flags inside the observations
    flags = cam_obs[:, :, 22:22 + 40].reshape(512, 4, 8, 5)[..., 4] > 0
    assert torch.equal(flags, env.camera_target_view_mask)
    stats = env.episode_statistics()
    assert stats['env_steps'] == 512
    with pytest.raises(ValueError):
        mate.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=4, reward_type='bogus')
    env.close()


def test_registered_ids_and_navigation():
    import mate_b200 as mate

    env = mate.make('MATE-Navigation-v0', num_envs=64)
    cam_obs, tgt_obs = env.reset(seed=0)
    assert cam_obs.shape == (64, 0, 190) and tgt_obs.shape == (64, 8, 195)
    (cam_obs, tgt_obs), rewards, done, infos = env.step((None, torch.zeros((64, 8, 2), device='cuda')))
    assert tgt_obs.shape == (64, 8, 195)
    env.load_config('MATE-4v2-9.yaml')
    assert (env.num_cameras, env.num_targets) == (4, 2)
    env.close()


def test_reference_compatible_single_env_mode_replays_a_reference_trace():
    """num_envs=None: NumPy float64 / float / bool / list-of-dict types of the reference, no
    auto-reset; fed with the reference's initial state and actions it returns the reference's
    observations (deterministic part of the trace: steps before the first stochastic draw)."""
    import mate_b200 as mate

    g = gu.load('4v8-9_greedy')
    env = mate.make('MultiAgentTracking-v0', config=str(g['config_name']))
    env.set_state(gu.state_arrays(g))
    n = 0
    for k in range(int(g['num_steps'])):
        if g['step_reached'][k].any():
            break
        obs, rew, done, infos = env.step((g['step_cam_act'][k], g['step_tgt_act'][k]))
        assert isinstance(rew[0], float) and isinstance(done, bool) and isinstance(infos[0], list)
        assert obs[0].dtype == np.float64 and obs[0].shape == (4, 126)
        assert rew[1] == g['step_reward'][k, 1]
        assert (env.camera_target_view_mask == g['step_mask_ct'][k]).all()
        n += 1
    assert n >= 1
    with pytest.raises(AssertionError):
        env.step((np.full((4, 2), np.nan), np.zeros((8, 2))))
    env.close()


def test_reset_with_an_explicit_seed_is_reproducible():
    """reset(seed=s) reseeds (mate/environment.py:700-701): the same seed gives the same episode on a long-lived
    environment; reset() without a seed continues with independent episodes."""
    import mate_b200

    env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=96)
    cam_a, tgt_a = (x.clone() for x in env.reset(seed=21))
    zeros = (torch.zeros((96, 4, 2), device='cuda'), torch.zeros((96, 8, 2), device='cuda'))
    for _ in range(3):
        env.step(zeros)
    cam_n, tgt_n = (x.clone() for x in env.reset())
    assert not torch.equal(tgt_n, tgt_a)
    cam_b, tgt_b = (x.clone() for x in env.reset(seed=21))
    assert torch.equal(cam_a, cam_b) and torch.equal(tgt_a, tgt_b)
    cam_c, tgt_c = (x.clone() for x in env.reset())       # ... and so is the episode that follows it
    assert torch.equal(cam_n, cam_c) and torch.equal(tgt_n, tgt_c)
    env.seed(22)
    cam_d, tgt_d = (x.clone() for x in env.reset())
    assert not torch.equal(tgt_d, tgt_a)
    env.close()


def test_wrappers_survive_load_config():
    """Wrapper.load_config re-runs every wrapper's constructor on the re-initialised environment (like the
    reference's wrappers): observation transformations, action tables, spaces and opponents follow the new config."""
    import mate_b200

    stack = [mate_b200.EnhancedObservation, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation,
             mate_b200.RepeatedRewardIndividualDone, lambda e: mate_b200.DiscreteCamera(e, levels=7),
             lambda e: mate_b200.MultiCamera(e, target_agent=mate_b200.GreedyTargetAgent(seed=4))]
    env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=48, wrappers=stack)
    env.reset(seed=1)
    env.load_config('MATE-4v2-9.yaml')
    fresh = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v2-9.yaml', num_envs=48, wrappers=stack)
    assert env.unwrapped.num_targets == 2 and env.levels == 7
    assert env.unwrapped.sim._obs_ops == fresh.unwrapped.sim._obs_ops and len(env.unwrapped.sim._obs_ops) == 4
    assert env.observation_space.spaces[0].shape == fresh.observation_space.spaces[0].shape
    obs_a, obs_b = env.reset(seed=8), fresh.reset(seed=8)
    assert obs_a.shape == (48, 4, env.unwrapped.sim.dc) and torch.equal(obs_a, obs_b)
    action = torch.randint(0, 49, (48, 4), device='cuda')
    for _ in range(5):
        out_a, out_b = env.step(action), fresh.step(action)
        assert torch.equal(out_a[0], out_b[0]) and torch.equal(out_a[1], out_b[1])
    env.close()
    fresh.close()


def test_message_api_in_single_env_mode():
    """send_messages / receive_messages / route_messages and the 'messages' entry of the step infos
    (mate/environment.py:641-669, 836-892, 1249-1269); the reference-side comparison is tests/test_reference_wrappers.py."""
    import mate_b200 as mate

    env = mate.make('MATE-4v8-9-v0')
    env.reset(seed=1)
    env.send_messages(mate.Message(sender=2, recipient=None, content={'k': 1}, team=mate.Team.TARGET))
    env.send_messages([mate.Message(sender=0, recipient=3, content='x', team=mate.Team.CAMERA)])
    assert [m.recipient for m in env.receive_messages(agent_id=(mate.Team.CAMERA, 3))] == [3]
    assert env.receive_messages(agent_id=(mate.Team.CAMERA, 3)) == []
    assert env.target_communication_edges[2].tolist() == [1] * 8
    _, _, _, (cam_infos, tgt_infos) = env.step((np.zeros((4, 2)), np.zeros((8, 2))))
    assert [len(info['messages']) for info in cam_infos] == [0, 0, 0, 1]
    assert all(len(info['messages']) == 1 and info['messages'][0].broadcasting and info['messages'][0]['k'] == 1 for info in tgt_infos)
    assert tgt_infos[2]['out_communication_edges'] == 8 and tgt_infos[5]['in_communication_edges'] == 1
    assert env.receive_messages() == ([[], [], [], []], [[] for _ in range(8)])
    assert env.target_total_communication_edges.sum() == 8 and env.target_communication_edges.sum() == 0
    _, _, _, (cam_infos, _) = env.step((np.zeros((4, 2)), np.zeros((8, 2))))
    assert cam_infos[3]['messages'] == []
    env.close()
    batched = mate.make('MATE-4v8-9-v0', num_envs=64)
    with pytest.raises(NotImplementedError):
        batched.send_messages(mate.Message(0, 1, {}, mate.Team.CAMERA))
    batched.close()
