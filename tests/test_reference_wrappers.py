"""The reference's OWN wrapper classes (``mate.wrappers``, unmodified, imported from ``/root/reference`` through the
stand-in ``gym`` of ``oracle/gymshim``) stacked on the reference-compatible single-environment mode of
``mate_b200.MultiAgentTracking`` -- BASELINE.json: "the mate.wrappers keep working unchanged".

The reference and a GPU never meet (``/root/reference`` exists in the build container only, the B200 box has no
reference), so the simulator behind the environment class is substituted by ``RecordedSim``: it serves states,
observations, masks and step outcomes recorded from the unmodified reference (``tests/golden/wrappers_*.npz``,
``tests/golden/aux_*.npz``) through the interface of ``mate_b200.sim.BatchedSim``.  What is under test is the HOST
layer a reference wrapper touches: the class identity checks (``mate/wrappers/typing.py:58-87``), the attribute
surface with the reference's NumPy types (``mate/environment.py:471-519``), the spaces, the entity views and the
tuple layouts.  The CUDA path itself is compared with the same fixtures in ``tests/test_wrappers.py`` and
``tests/test_aux_wrappers.py`` (``-m gpu``).  Skipped where the reference is not mounted."""

import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get('MATE_REFERENCE', '/root/reference')
GOLDEN = os.path.join(HERE, 'golden')

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'mate')), reason='the reference is not mounted')

torch = pytest.importorskip('torch')


def _flat_config(g):
    import golden_util as gu

    return gu.flat_config(g)


class RecordedSim:
    """Stand-in for ``mate_b200.sim.BatchedSim`` (one environment) that replays recorded samples."""

    fixture = None     # set by the tests before the environment is constructed: (npz, prefix)

    def __init__(self, flat_config, num_envs, device=0, env_index_base=0):
        assert num_envs == 1
        self.cfg = flat_config
        self.device = torch.device('cpu')
        self.num_envs = 1
        nc, nt, no = flat_config['num_cameras'], flat_config['num_targets'], flat_config['num_obstacles']
        self.nc, self.nt, self.no = nc, nt, no
        self.dc = 22 + 5 * nt + 4 * no + 7 * nc
        self.dt = 27 + 7 * nc + 4 * no + 5 * nt
        self.g, self.prefix = type(self).fixture
        self.index = 0
        self.actions = []
        self._aux = None
        self.launch_count = 0

    # ---- the part of the BatchedSim interface the environment class uses
    def alloc_aux(self):
        nc, nt, no = self.nc, self.nt, self.no
        u8 = torch.uint8
        self._aux = {
            'mask_ct': torch.zeros((1, nc, nt), dtype=u8), 'mask_cc': torch.zeros((1, nc, nc), dtype=u8),
            'mask_co': torch.zeros((1, nc, no), dtype=u8), 'mask_tc': torch.zeros((1, nt, nc), dtype=u8),
            'mask_to': torch.zeros((1, nt, no), dtype=u8), 'mask_tt': torch.zeros((1, nt, nt), dtype=u8),
            'target_dones': torch.zeros((1, nt), dtype=u8), 'is_colliding': torch.zeros((1, nt), dtype=u8),
            'warehouse_dist': torch.zeros((1, nt, 4)), 'coverage': torch.zeros((1, 3)),
            'num_delivered': torch.zeros(1, dtype=torch.int32), 'episode_step': torch.zeros(1, dtype=torch.int32),
            'tgt_goal': torch.zeros((1, nt), dtype=torch.int32), 'tgt_empty_bits': torch.zeros((1, nt), dtype=u8),
        }
        return self._aux

    def seed(self, seed):
        pass

    def close(self):
        pass

    def load(self, index):
        self.index = index

    def set_state(self, arrays):
        pass

    def _get(self, key):
        return self.g[self.prefix + key][self.index]

    def get_state(self):
        empty = self._get('tgt_empty_bits')
        return {
            'cam_xy': self._get('cam_xy')[None], 'cam_phi': self._get('cam_phi')[None], 'cam_theta': self._get('cam_theta')[None],
            'tgt_xy': self._get('tgt_xy')[None], 'obs_xyr': self._get('obs_xyr')[None],
            'tgt_capacity': self._get('tgt_capacity')[None].astype(np.int32), 'tgt_goal': self._get('tgt_goal')[None].astype(np.int32),
            'tgt_weight': self._get('tgt_goal_weight')[None].astype(np.int32),
            'tgt_bounty': self._get('bounties')[None].astype(np.int32),
            'tgt_empty_bits': (empty.astype(np.int32) << np.arange(4)).sum(axis=-1)[None].astype(np.int32),
            'remaining': self._get('remaining')[None].astype(np.int32), 'awaiting': self._get('awaiting')[None].astype(np.int32),
            'num_delivered': np.array([self._get('num_delivered')], dtype=np.int32),
            'episode_step': np.array([self._get('episode_step')], dtype=np.int32),
            'episode_reward': np.array([[self._get('ep_reward'), self._get('delayed_ep_reward')]], dtype=np.float64),
        }

    def _masks_from_observations(self, cam, tgt):
        """The view masks are the flag entries of the recorded raw observations."""
        nc, nt, no = self.nc, self.nt, self.no
        a = self._aux
        if nc:
            a['mask_ct'][0] = torch.from_numpy(cam[:, 22:22 + 5 * nt].reshape(nc, nt, 5)[..., 4].astype(np.uint8))
            a['mask_co'][0] = torch.from_numpy(cam[:, 22 + 5 * nt:22 + 5 * nt + 4 * no].reshape(nc, no, 4)[..., 3].astype(np.uint8))
            a['mask_cc'][0] = torch.from_numpy(cam[:, 22 + 5 * nt + 4 * no:].reshape(nc, nc, 7)[..., 6].astype(np.uint8))
            a['mask_tc'][0] = torch.from_numpy(tgt[:, 27:27 + 7 * nc].reshape(nt, nc, 7)[..., 6].astype(np.uint8))
        a['mask_to'][0] = torch.from_numpy(tgt[:, 27 + 7 * nc:27 + 7 * nc + 4 * no].reshape(nt, no, 4)[..., 3].astype(np.uint8))
        a['mask_tt'][0] = torch.from_numpy(tgt[:, 27 + 7 * nc + 4 * no:].reshape(nt, nt, 5)[..., 4].astype(np.uint8))

    def observe(self, replay=None, aux=False):
        cam, tgt = self._get('cam_obs'), self._get('tgt_obs')
        self._masks_from_observations(cam, tgt)
        self._aux['is_colliding'][0] = torch.from_numpy(self._get('tgt_colliding').astype(np.uint8))
        return torch.from_numpy(cam.astype(np.float32))[None], torch.from_numpy(tgt.astype(np.float32))[None]

    def reset(self, seed=None, env_mask=None):
        return self.observe()

    def step(self, cam_act, tgt_act, auto_reset=True, replay=None, aux=False):
        """Serves the recorded outcome of the step of sample `index` (``a_out_*``)."""
        self.actions.append((None if cam_act is None else cam_act.numpy().copy(), tgt_act.numpy().copy()))
        a = self._aux
        for key in ('mask_ct', 'mask_cc', 'mask_co', 'mask_tc', 'mask_to', 'mask_tt'):
            if a[key].numel():
                a[key][0] = torch.from_numpy(self._get('out_' + key).astype(np.uint8))
        a['target_dones'][0] = torch.from_numpy(self._get('out_tgt_individual_done').astype(np.uint8))
        a['is_colliding'][0] = torch.from_numpy(self._get('out_tgt_is_colliding').astype(np.uint8))
        xy = self._get('out_tgt_xy')
        wh = 925.0 * np.array([[1.0, 1.0], [-1.0, 1.0], [-1.0, -1.0], [1.0, -1.0]])
        a['warehouse_dist'][0] = torch.from_numpy(np.linalg.norm(xy[:, None] - wh[None], axis=-1).astype(np.float32))
        a['tgt_goal'][0] = torch.from_numpy(self._get('out_tgt_goal').astype(np.int32))
        a['episode_step'][0] = int(self._get('out_episode_step'))
        rewards = torch.from_numpy(self._get('out_team_reward').astype(np.float32))[None]
        cam_obs = torch.zeros((1, self.nc, self.dc))
        tgt_obs = torch.zeros((1, self.nt, self.dt))
        return (cam_obs, tgt_obs), rewards, torch.zeros(1, dtype=torch.uint8)


class StepOutcomeSim(RecordedSim):
    """For the step test the state the wrappers read AFTER the step is the recorded outcome."""

    def get_state(self):
        s = super().get_state()
        s['tgt_goal'] = self._get('out_tgt_goal')[None].astype(np.int32)
        s['tgt_xy'] = self._get('out_tgt_xy')[None]
        return s


@pytest.fixture(scope='module')
def mate():
    sys.path.insert(0, os.path.join(REPO, 'oracle', 'gymshim'))
    sys.path.insert(0, REFERENCE)
    import mate as reference  # pylint: disable=import-outside-toplevel

    return reference


def _drop_in(mate, monkeypatch, sim_class, fixture):
    """What a maintainer of the reference adds: the reference's class name bound to a subclass of
    the B200 environment that is also a ``gym.Env``."""
    import gym  # the stand-in  # pylint: disable=import-outside-toplevel

    import mate_b200.environment as b200  # pylint: disable=import-outside-toplevel

    meta = type('DropInMeta', (type(mate.MultiAgentTracking), b200.EnvMeta), {})
    sim_class.fixture = fixture
    drop_in = meta('MultiAgentTracking', (b200.MultiAgentTracking, gym.Env), {'_sim_class': sim_class})
    import mate.environment  # pylint: disable=import-outside-toplevel
    import mate.wrappers.typing  # pylint: disable=import-outside-toplevel

    for module in (mate, mate.environment, mate.wrappers.typing):
        monkeypatch.setattr(module, 'MultiAgentTracking', drop_in)
    return drop_in


STACKS = {   # oracle/gen_wrapper_golden.py
    'enhanced_both': [('EnhancedObservation', {'team': 'both'})],
    'enhanced_camera': [('EnhancedObservation', {'team': 'camera'})],
    'enhanced_target': [('EnhancedObservation', {'team': 'target'})],
    'shared_both': [('SharedFieldOfView', {'team': 'both'})],
    'shared_camera': [('SharedFieldOfView', {'team': 'camera'})],
    'shared_target': [('SharedFieldOfView', {'team': 'target'})],
    'relative': [('RelativeCoordinates', {})],
    'rescaled': [('RescaledObservation', {})],
    'shared_relative_rescaled': [('SharedFieldOfView', {'team': 'both'}), ('RelativeCoordinates', {}), ('RescaledObservation', {})],
    'enhanced_relative_rescaled': [('EnhancedObservation', {'team': 'both'}), ('RelativeCoordinates', {}), ('RescaledObservation', {})],
    'enhcam_sharedtgt_relative': [('EnhancedObservation', {'team': 'camera'}), ('SharedFieldOfView', {'team': 'target'}), ('RelativeCoordinates', {})],
}


def _apply_stack(wrapped, base, observation):
    chain = []
    env = wrapped
    while env is not base:
        chain.append(env)
        env = env.env
    obs = tuple(np.array(o, dtype=np.float64, copy=True) for o in observation)
    for wrapper in reversed(chain):
        obs = wrapper.observation(obs)
        obs = tuple(np.array(o, dtype=np.float64, copy=True) for o in obs)
    return obs


@pytest.mark.parametrize('fixture_name', ['wrappers_4v8-9', 'wrappers_8v8-9', 'wrappers_Navigation', 'wrappers_4v2-0'])
def test_reference_observation_wrappers_on_the_b200_environment(mate, monkeypatch, fixture_name):
    g = np.load(os.path.join(GOLDEN, fixture_name + '.npz'))
    drop_in = _drop_in(mate, monkeypatch, RecordedSim, (g, 'w_'))
    env = drop_in(config=str(g['config_name']))
    assert isinstance(env, mate.MultiAgentTracking)
    nc = env.num_cameras
    stacks = {}
    for name, spec in STACKS.items():
        wrapped = env
        for cls, kwargs in spec:
            wrapped = getattr(mate, cls)(wrapped, **kwargs)      # the reference's own asserts run here
        assert isinstance(wrapped, mate.MultiAgentTracking)       # ... and the see-through instance check
        stacks[name] = wrapped
    for i in range(int(g['count'])):
        env.sim.load(i)
        env.set_state({})                                          # new snapshot
        raw = env.joint_observation()
        assert raw[0].dtype == np.float64 and raw[1].shape == (env.num_targets, env.target_observation_dim)
        for name, wrapped in stacks.items():
            cam, tgt = _apply_stack(wrapped, env, raw)
            # the raw observations went through the float32 I/O of the C ABI: coordinates around 1000 carry 6e-5
            atol = 2e-6 if 'rescaled' in name else 3e-4
            if nc:
                np.testing.assert_allclose(cam, g[f'w_{name}_cam_obs'][i], rtol=1e-5, atol=atol, err_msg=f'{name} sample {i}')
            np.testing.assert_allclose(tgt, g[f'w_{name}_tgt_obs'][i], rtol=1e-5, atol=atol, err_msg=f'{name} sample {i}')
    # entity views as the reference's wrappers and agents read them
    env.sim.load(0)
    env.set_state({})
    for t, target in enumerate(env.targets):
        np.testing.assert_allclose(target.state(private=True)[:6], raw_private(env, g, 0, t)[:6], rtol=1e-12)
        assert isinstance(target.is_colliding, bool)
    if env.num_obstacles:
        assert env.obstacle_states_flagged.shape == (env.num_obstacles, 4) and (env.obstacle_states_flagged[:, 3] == 1).all()
        assert env.obstacles[0].distance(env.targets[0]) == pytest.approx(
            float(np.linalg.norm(g['w_obs_xyr'][0][0, :2] - g['w_tgt_xy'][0][0])))


def raw_private(env, g, i, t):
    """Private state of target t as the reference's own observation recorded it."""
    return g['w_tgt_obs'][i][t, 13:27]


@pytest.mark.parametrize('fixture_name', ['aux_4v8-9', 'aux_8v8-9'])
def test_readme_wrapper_stack_steps_on_the_b200_environment(mate, monkeypatch, fixture_name):
    """README stack of the reference: EnhancedObservation -> MoreTrainingInformation -> DiscreteCamera ->
    RepeatedRewardIndividualDone; the infos the reference's wrappers assemble from the environment's attributes
    against what they assembled from the reference environment (``a_out_*``)."""
    g = np.load(os.path.join(GOLDEN, fixture_name + '.npz'))
    w = np.load(os.path.join(GOLDEN, fixture_name.replace('aux_', 'wrappers_') + '.npz'))
    drop_in = _drop_in(mate, monkeypatch, StepOutcomeSim, (g, 'a_'))
    base = drop_in(config=str(g['config_name']))
    env = mate.EnhancedObservation(base, team='both')
    env = mate.MoreTrainingInformation(env)
    env = mate.DiscreteCamera(env, levels=5)
    env = mate.RepeatedRewardIndividualDone(env)
    nc, nt = base.num_cameras, base.num_targets
    base._needs_reset = False   # pylint: disable=protected-access
    rng = np.random.RandomState(0)
    for i in list(range(0, int(g['count']), 7))[:60]:
        base.sim.load(i)
        idx = rng.randint(0, 25, size=nc)
        tgt_act = g['a_tgt_act'][i]
        (cam_obs, tgt_obs), (cam_rew, tgt_rew), (cam_done, tgt_done), (cam_infos, tgt_infos) = env.step((idx, tgt_act))
        # DiscreteCamera decoded the indices through the reference's table before they reached the simulator
        np.testing.assert_allclose(base.sim.actions[-1][0][0], w['discrete_camera_5'][idx], rtol=1e-6)
        assert cam_obs.shape == (nc, base.camera_observation_dim) and tgt_obs.shape == (nt, base.target_observation_dim)
        assert list(cam_rew) == [float(g['a_out_team_reward'][i][0])] * nc and list(tgt_rew) == [float(g['a_out_team_reward'][i][1])] * nt
        assert len(cam_done) == nc and len(tgt_done) == nt
        for c, info in enumerate(cam_infos):
            assert info['num_tracked'] == g['a_out_cam_num_tracked'][i][c]
            assert bool(info['is_sensed']) == bool(g['a_out_cam_is_sensed'][i][c])
            assert (info['camera_target_view_mask'] == g['a_out_mask_ct'][i].astype(bool)).all()
        for t, info in enumerate(tgt_infos):
            assert info['goal'] == g['a_out_tgt_goal'][i][t]
            assert info['goal_distance'] == pytest.approx(g['a_out_tgt_goal_distance'][i][t], rel=1e-5, abs=2e-4)
            np.testing.assert_allclose(info['warehouse_distances'], g['a_out_tgt_warehouse_distances'][i][t], rtol=1e-5, atol=2e-4)
            assert bool(info['individual_done']) == bool(g['a_out_tgt_individual_done'][i][t])
            assert bool(info['is_tracked']) == bool(g['a_out_tgt_is_tracked'][i][t])
            assert bool(info['is_colliding']) == bool(g['a_out_tgt_is_colliding'][i][t])
            assert info['state'].shape == base.state_space.shape


# ---------------------------------------------------------------------------------------------------------------------
# intra-team messages (mate/environment.py:836-892, 1249-1269): the same traffic through the reference environment
# and through the B200 environment class, with the reference's own agents and communication wrappers on top
# ---------------------------------------------------------------------------------------------------------------------
class ObservationsOnlySim(RecordedSim):
    """Serves the recorded observations of sample `index` as the outcome of every step (``wrappers_*.npz`` holds states
    and observations, no step outcomes): enough for everything that happens ABOVE the simulator."""

    def step(self, cam_act, tgt_act, auto_reset=True, replay=None, aux=False):
        self.actions.append((None if cam_act is None else cam_act.numpy().copy(), tgt_act.numpy().copy()))
        return self.observe(), torch.zeros((1, 2)), torch.zeros(1, dtype=torch.uint8)


def _message_view(messages):
    return [(m.sender, m.recipient, m.team.value, bool(m.broadcasting), repr(m.content)) for m in messages]


def test_message_api_matches_the_reference_environment(mate, monkeypatch):
    g = np.load(os.path.join(GOLDEN, 'wrappers_4v8-9.npz'))
    reference_env = mate.MultiAgentTracking(config=str(g['config_name']))   # the unmodified reference class (CPU)
    reference_env.seed(0)
    reference_env.reset()
    drop_in = _drop_in(mate, monkeypatch, ObservationsOnlySim, (g, 'w_'))
    env = drop_in(config=str(g['config_name']))
    env.reset()
    Message, Team = mate.utils.Message, mate.utils.Team
    rng = np.random.RandomState(3)
    for _ in range(4):
        for e in (reference_env, env):
            state = rng.get_state()
            for _ in range(6):
                team = Team.CAMERA if rng.rand() < 0.5 else Team.TARGET
                n = e.num_cameras if team is Team.CAMERA else e.num_targets
                sender = int(rng.randint(n))
                recipient = None if rng.rand() < 0.3 else int(rng.randint(n))
                e.send_messages(Message(sender=sender, recipient=recipient, content={'k': int(rng.randint(100))}, team=team))
            e.send_messages([Message(sender=0, recipient=1, content='a', team=Team.TARGET),
                             Message(sender=1, recipient=None, content='b', team=Team.TARGET)])
            if e is reference_env:
                rng.set_state(state)
        # one agent picks its messages up early; the infos of the next step still carry everything that was sent
        assert _message_view(env.receive_messages(agent_id=(Team.TARGET, 1))) == \
            _message_view(reference_env.receive_messages(agent_id=(Team.TARGET, 1)))
        assert env.receive_messages(agent_id=(Team.TARGET, 1)) == []
        for a, b in zip(env.communication_edges, reference_env.communication_edges):
            assert (a == b).all()
        action = (np.zeros((env.num_cameras, 2)), np.zeros((env.num_targets, 2)))
        _, _, _, (cam_infos, tgt_infos) = env.step(action)
        _, _, _, (ref_cam_infos, ref_tgt_infos) = reference_env.step(action)
        for mine, theirs in zip(cam_infos + tgt_infos, ref_cam_infos + ref_tgt_infos):
            assert _message_view(mine['messages']) == _message_view(theirs['messages'])
            assert mine['out_communication_edges'] == theirs['out_communication_edges']
            assert mine['in_communication_edges'] == theirs['in_communication_edges']
        for a, b in zip(env.total_communication_edges, (reference_env.camera_total_communication_edges,
                                                        reference_env.target_total_communication_edges)):
            assert (a == b).all()
        assert env.receive_messages() == ([[] for _ in range(env.num_cameras)], [[] for _ in range(env.num_targets)])
    with pytest.raises(AssertionError):
        env.send_messages([Message(0, 1, 'x', Team.CAMERA), Message(0, 1, 'y', Team.TARGET)])


def test_reference_agents_and_communication_wrappers_on_the_b200_environment(mate, monkeypatch):
    """The reference's single-team wrapper drives the reference's own greedy camera agents (observe -> send -> receive ->
    act through the environment's message queues) on the B200 environment class, below the reference's communication
    wrappers."""
    g = np.load(os.path.join(GOLDEN, 'wrappers_4v8-9.npz'))
    drop_in = _drop_in(mate, monkeypatch, ObservationsOnlySim, (g, 'w_'))
    base = drop_in(config=str(g['config_name']))
    env = mate.RestrictedCommunicationRange(base, range_limit=1500.0)
    env = mate.RandomMessageDropout(env, dropout_rate=0.25)
    # (ExtraCommunicationDelays is left out: its heap compares Message objects on ties, which the reference's own
    #  dataclass does not support -- it fails on the reference environment in the same way)
    env = mate.MultiTarget(env, camera_agent=mate.GreedyCameraAgent(seed=0))
    assert isinstance(env, mate.MultiAgentTracking)
    base.sim.load(0)
    tgt_obs = env.reset()
    assert tgt_obs.shape == (base.num_targets, base.target_observation_dim)
    for i in range(1, 12):
        base.sim.load(i)
        tgt_obs, tgt_reward, done, tgt_infos = env.step(np.zeros((base.num_targets, 2)))
        assert tgt_obs.shape == (base.num_targets, base.target_observation_dim) and len(tgt_infos) == base.num_targets
        cam_act = base.sim.actions[-1][0][0]
        assert cam_act.shape == (base.num_cameras, 2) and np.isfinite(cam_act).all()
        assert (np.abs(cam_act) <= np.array([base.camera_rotation_step, base.camera_zooming_step]) + 1e-6).all()
    # the greedy cameras tell their teammates where they are (first step) and which targets they track, through the
    # environment's message queues and the communication wrappers
    assert int(base.camera_total_communication_edges.sum()) > 0
    plain = mate.NoCommunication(drop_in(config=str(g['config_name'])), team='both')
    plain.reset()
    plain.send_messages(mate.utils.Message(sender=0, recipient=None, content={}, team=mate.utils.Team.TARGET))
    assert plain.receive_messages() == ([[] for _ in range(base.num_cameras)], [[] for _ in range(base.num_targets)])
