"""Batched restatements of the reference's observation / action wrappers.

Same class names, constructor arguments, ordering rules and attribute surface as ``mate/wrappers/*.py``;
the difference is where the work happens.  The reference applies every wrapper as NumPy slicing on the joint
observation of ONE environment (``np.hstack``, ``.any(axis=0)``, per-row loops).  Here an observation
wrapper only *registers* its transformation with the environment: all registered transformations are
applied in the reference's order by one CUDA kernel over the ``[B, N, D]`` observation tensors right after
the step kernel (``mate_b200_transform_observations``, ``mate_b200/csrc/mate_wrappers.cuh``), so a stack of
wrappers costs one extra read + write of the observations, not one per wrapper.

    env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=65536,
                         wrappers=[mate_b200.SharedFieldOfView, mate_b200.RelativeCoordinates,
                                   mate_b200.RescaledObservation, mate_b200.DiscreteCamera])
"""

import functools

import numpy as np
import torch

from mate_b200 import _abi, spaces

TEAMS = ('both', 'camera', 'target', 'none')


class Wrapper:
    """gym.Wrapper-like forwarding: unknown public attributes resolve on the wrapped environment."""

    def __init__(self, env):
        self.env = env

    def __init_subclass__(cls, **kwargs):
        # remember the constructor arguments of the outermost __init__ call: load_config re-runs it
        super().__init_subclass__(**kwargs)
        init = cls.__dict__.get('__init__')
        if init is None:
            return

        @functools.wraps(init)
        def remembering_init(self, env, *args, **kw):
            if type(self) is cls:
                self.__dict__['_ctor_args'] = (args, kw)
            init(self, env, *args, **kw)

        cls.__init__ = remembering_init

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(f"attempted to get missing private attribute '{name}'")
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    def joint_observation(self):
        return self.env.joint_observation()

    def load_config(self, config=None):
        """Like every wrapper of the reference (``self.env.load_config(config); self.__init__(self.env, ...)``):
        the wrapped environment is re-initialised first, then this wrapper re-runs its own constructor on it, which
        re-registers observation transformations, rebuilds action tables and spaces and re-binds opponent agents
        to the new simulator."""
        self.env.load_config(config=config)
        args, kw = self.__dict__.get('_ctor_args', ((), {}))
        self.__init__(self.env, *args, **kw)  # pylint: disable=unnecessary-dunder-call

    def close(self):
        return self.env.close()

    def __str__(self):
        return f'<{type(self).__name__}{self.env}>'


def _has_wrapper(env, cls):
    while isinstance(env, Wrapper):
        if isinstance(env, cls):
            return True
        env = env.env
    return False


class _ObservationWrapper(Wrapper):
    """Registers its op codes with the base environment (applied on the GPU after every step)."""

    def _register(self, ops):
        base = self.unwrapped
        base.sim.set_observation_wrappers(list(base.sim._obs_ops) + list(ops))  # pylint: disable=protected-access


class EnhancedObservation(_ObservationWrapper):
    """mate/wrappers/enhanced_observation.py: all observation masks True; targets see the empty status of all
    warehouses."""

    def __init__(self, env, team='both'):
        assert team in TEAMS, f'Invalid argument team {team!r}. Expect one of {TEAMS}.'
        assert not _has_wrapper(env, RelativeCoordinates), f'You should use wrapper `{type(self)}` before `RelativeCoordinates`.'
        assert not _has_wrapper(env, RescaledObservation), f'You should use wrapper `{type(self)}` before `RescaledObservation`.'
        super().__init__(env)
        self.team = team
        self.enhanced_camera = team in ('camera', 'both')
        self.enhanced_target = team in ('target', 'both')
        ops = []
        if self.enhanced_camera and env.num_cameras > 0:
            ops.append(_abi.OBS_ENHANCED_CAMERA)
        if self.enhanced_target:
            ops.append(_abi.OBS_ENHANCED_TARGET)
        self._register(ops)

    def __str__(self):
        return f'<{type(self).__name__}(team={self.team}){self.env}>'


class SharedFieldOfView(_ObservationWrapper):
    """mate/wrappers/shared_field_of_view.py: "or" of the view masks of the agents of a team."""

    def __init__(self, env, team='both'):
        assert team in TEAMS, f'Invalid argument team {team!r}. Expect one of {TEAMS}.'
        assert not _has_wrapper(env, RelativeCoordinates), f'You should use wrapper `{type(self)}` before `RelativeCoordinates`.'
        assert not _has_wrapper(env, RescaledObservation), f'You should use wrapper `{type(self)}` before `RescaledObservation`.'
        super().__init__(env)
        self.team = team
        self.shared_camera = team in ('camera', 'both')
        self.shared_target = team in ('target', 'both')
        ops = []
        if self.shared_camera and env.num_cameras > 0:
            ops.append(_abi.OBS_SHARED_CAMERA)
        if self.shared_target:
            ops.append(_abi.OBS_SHARED_TARGET)
        self._register(ops)

    def __str__(self):
        return f'<{type(self).__name__}(team={self.team}){self.env}>'


class RelativeCoordinates(_ObservationWrapper):
    """mate/wrappers/relative_coordinates.py: locations of the other (visible) entities and of the warehouses
    relative to the observing agent."""

    def __init__(self, env):
        assert not _has_wrapper(env, RelativeCoordinates), f'You should not use wrapper `{type(self)}` more than once.'
        super().__init__(env)
        self._register([_abi.OBS_RELATIVE])


class RescaledObservation(_ObservationWrapper):
    """mate/wrappers/rescaled_observation.py: all entity states rescaled to [-1, +1]."""

    def __init__(self, env):
        assert not _has_wrapper(env, RescaledObservation), f'You should not use wrapper `{type(self)}` more than once.'
        super().__init__(env)
        base = self.unwrapped
        from mate_b200.sim import rescale_tables  # pylint: disable=import-outside-toplevel

        cam_t, tgt_t = rescale_tables(base.num_cameras, base.num_targets, base.num_obstacles)

        def rescaled(space, table):
            t = table.astype(np.float64)
            low = np.where(np.isfinite(space.low), space.low * t[:, 0] + t[:, 1], space.low)
            high = np.where(np.isfinite(space.high), space.high * t[:, 0] + t[:, 1], space.high)
            return spaces.Box(low=low, high=high)

        self.camera_observation_space = rescaled(base.camera_observation_space, cam_t)
        self.target_observation_space = rescaled(base.target_observation_space, tgt_t)
        self.camera_joint_observation_space = spaces.Tuple((self.camera_observation_space,) * base.num_cameras)
        self.target_joint_observation_space = spaces.Tuple((self.target_observation_space,) * base.num_targets)
        self.observation_space = spaces.Tuple((self.camera_joint_observation_space, self.target_joint_observation_space))
        self._register([_abi.OBS_RESCALED])


def camera_action_grid(levels):
    """mate/wrappers/discrete_action_spaces.py:98-117: action_grid[i + levels * j] = (x_i, y_j) in [-1, 1]^2."""
    ticks = np.linspace(start=-1.0, stop=+1.0, num=levels, endpoint=True)
    return np.stack(np.meshgrid(ticks, ticks), axis=-1).reshape(-1, 2)


def target_action_grid(levels):
    """mate/wrappers/discrete_action_spaces.py:204-228: the square grid mapped into the unit disc."""
    grid = camera_action_grid(levels)
    angle = np.arctan2(grid[..., -1], grid[..., 0])
    bound = 1.0 / np.cos(np.pi * ((angle / np.pi + 0.25) % 0.5 - 0.25))
    return grid / bound[..., np.newaxis]


class _DiscreteActions(Wrapper):
    def _table(self, grid):
        return torch.from_numpy(np.ascontiguousarray(grid, dtype=np.float32)).to(self.unwrapped.device)

    def _decode(self, index, table):
        base = self.unwrapped
        index = torch.as_tensor(np.asarray(index) if not torch.is_tensor(index) else index)
        return base.sim.decode_actions(index.to(base.device), table)


class DiscreteCamera(_DiscreteActions):
    """mate/wrappers/discrete_action_spaces.py:21-117: cameras use a levels x levels grid of discrete actions."""

    def __init__(self, env, levels=5):
        assert not _has_wrapper(env, DiscreteCamera), f'You should not use wrapper `{type(self)}` more than once.'
        assert levels >= 3 and levels % 2 == 1, f'The discrete level must be an odd number that not less than 3. Got levels = {levels}.'
        assert env.num_cameras > 0, 'There must be at least one camera in the environment.'
        super().__init__(env)
        self.levels = levels
        self.camera_action_space = spaces.Discrete(levels * levels)
        self.camera_joint_action_space = spaces.Tuple((self.camera_action_space,) * env.num_cameras)
        self.action_space = spaces.Tuple((self.camera_joint_action_space, env.target_joint_action_space))
        self.action_high = np.asarray([env.camera_rotation_step, env.camera_zooming_step], dtype=np.float64)
        self.normalized_action_grid = camera_action_grid(levels)
        self.action_grid = self.action_high * self.normalized_action_grid
        self._device_table = self._table(self.action_grid)

    def action(self, action):
        camera_joint_action, target_joint_action = action
        return self._decode(camera_joint_action, self._device_table), target_joint_action

    def step(self, action):
        return self.env.step(self.action(action))


class DiscreteTarget(_DiscreteActions):
    """mate/wrappers/discrete_action_spaces.py:127-228: targets use a levels x levels grid of discrete actions."""

    def __init__(self, env, levels=5):
        assert not _has_wrapper(env, DiscreteTarget), f'You should not use wrapper `{type(self)}` more than once.'
        assert levels >= 3 and levels % 2 == 1, f'The discrete level must be an odd number that not less than 3. Got levels = {levels}.'
        super().__init__(env)
        self.levels = levels
        self.target_action_space = spaces.Discrete(levels * levels)
        self.target_joint_action_space = spaces.Tuple((self.target_action_space,) * env.num_targets)
        self.action_space = spaces.Tuple((env.camera_joint_action_space, self.target_joint_action_space))
        self.action_high = np.asarray([env.target_step_size, env.target_step_size], dtype=np.float64)
        self.normalized_action_grid = target_action_grid(levels)
        self.action_grid = self.action_high * self.normalized_action_grid
        self._device_table = self._table(self.action_grid)

    def action(self, action):
        camera_joint_action, target_joint_action = action
        return camera_joint_action, self._decode(target_joint_action, self._device_table)

    def step(self, action):
        return self.env.step(self.action(action))


class RepeatedRewardIndividualDone(Wrapper):
    """mate/wrappers/repeated_reward_individual_done.py: the team rewards repeated per agent and a done flag per
    agent (targets: ``target_dones`` if ``target_done_at_destination``)."""

    def __init__(self, env, target_done_at_destination=False):
        assert not _has_wrapper(env, RepeatedRewardIndividualDone), f'You should not use wrapper `{type(self)}` more than once.'
        super().__init__(env)
        self.target_done_at_destination = target_done_at_destination

    def step(self, action):
        observation, (camera_team_reward, target_team_reward), done, info = self.env.step(action)
        base = self.unwrapped
        nc, nt = base.num_cameras, base.num_targets
        if base.batched:
            done_t = torch.as_tensor(done)
            if self.target_done_at_destination:
                target_dones = base.target_dones.to(done_t.dtype)
            else:
                target_dones = done_t.unsqueeze(-1).expand(-1, nt)
            reward = (camera_team_reward.unsqueeze(-1).expand(-1, nc), target_team_reward.unsqueeze(-1).expand(-1, nt))
            return observation, reward, (done_t.unsqueeze(-1).expand(-1, nc), target_dones), info
        target_dones = [bool(x) for x in np.asarray(base.target_dones).reshape(-1)] if self.target_done_at_destination else [done] * nt
        reward = ([camera_team_reward] * nc, [target_team_reward] * nt)
        return observation, reward, ([done] * nc, target_dones), info


# ------------------------------------------------------------------------------------------------------------
# MoreTrainingInformation and the auxiliary-reward wrappers.  The per-agent terms come
# from ONE kernel per step over the auxiliary outputs of the step kernel (``mate_b200_auxiliary_terms``); the
# wrappers only combine columns of that tensor.  Batched: rewards / infos are ``[B, N]`` CUDA tensors (infos: two
# dicts of tensors); reference-compatible single env: Python floats and the reference's list-of-dict infos.
# ------------------------------------------------------------------------------------------------------------

CAMERA_REWARD_KEYS = ('raw_reward', 'coverage_rate', 'real_coverage_rate', 'mean_transport_rate',
                      'soft_coverage_score', 'num_tracked', 'baseline')
TARGET_REWARD_KEYS = ('raw_reward', 'coverage_rate', 'real_coverage_rate', 'mean_transport_rate',
                      'normalized_goal_distance', 'sparse_delivery', 'soft_coverage_score', 'is_tracked',
                      'is_colliding', 'baseline')
_REDUCERS = {
    'mean': lambda x: x.mean(dim=-1, keepdim=True), 'sum': lambda x: x.sum(dim=-1, keepdim=True),
    'max': lambda x: x.max(dim=-1, keepdim=True).values, 'min': lambda x: x.min(dim=-1, keepdim=True).values,
}


def _update_infos(base, infos, per_agent, shared=None):
    """Merge ``{key: [B, N] tensor}`` into the infos of one team (dict of tensors, or the reference's list of dicts)."""
    if base.batched:
        infos.update(per_agent)
        if shared:
            infos.update(shared)
        return
    host = {k: v[0].tolist() for k, v in per_agent.items()}
    for i, info in enumerate(infos):
        for k, values in host.items():
            info[k] = values[i]
        if shared:
            info.update(shared)


class MoreTrainingInformation(Wrapper):
    """mate/wrappers/more_training_information.py: more entries in the info dicts.  Cameras: ``num_tracked``,
    ``is_sensed``; targets: ``goal``, ``goal_distance``, ``warehouse_distances``, ``individual_done``, ``is_tracked``,
    ``is_colliding``; both: the private states of all agents and the view masks.  The entries that need the cargo
    tables / the global state on the host (``state``, ``remaining_cargoes``, ``remaining_cargo_counts``,
    ``awaiting_cargo_counts``, ``obstacle_states``) cost a device synchronisation and are added only with
    ``full_observability=True`` (the default for the reference-compatible single env)."""

    def __init__(self, env, full_observability=None):
        assert not _has_wrapper(env, MoreTrainingInformation), f'You should not use wrapper `{type(self)}` more than once.'
        super().__init__(env)
        base = self.unwrapped
        self.full_observability = (not base.batched) if full_observability is None else bool(full_observability)

    def step(self, action):
        results = self.env.step(action)
        (camera_joint_observation, target_joint_observation), _, _, (camera_infos, target_infos) = results
        base = self.unwrapped
        cam_terms, tgt_terms = base.auxiliary_terms()
        aux = base._aux  # pylint: disable=protected-access
        if base.num_cameras:
            _update_infos(base, camera_infos, {'num_tracked': cam_terms[..., 5].long(), 'is_sensed': cam_terms[..., 7] != 0})
        _update_infos(base, target_infos, {
            'goal': tgt_terms[..., 10].long(), 'goal_distance': tgt_terms[..., 11], 'warehouse_distances': tgt_terms[..., 12:16],
            'individual_done': tgt_terms[..., 5] != 0, 'is_tracked': tgt_terms[..., 7] != 0, 'is_colliding': tgt_terms[..., 8] != 0,
        })
        # full observability (more_training_information.py:84-98)
        offset = 13   # PRESERVED_DIM
        shared = {
            'camera_target_view_mask': aux['mask_ct'].bool(), 'camera_obstacle_view_mask': aux['mask_co'].bool(),
            'target_camera_view_mask': aux['mask_tc'].bool(), 'target_obstacle_view_mask': aux['mask_to'].bool(),
            'target_target_view_mask': aux['mask_tt'].bool(),
        }
        if base.batched:
            shared['camera_states'] = camera_joint_observation[..., offset:offset + 9]
            shared['target_states'] = target_joint_observation[..., offset:offset + 14]
        else:
            shared = {k: v[0].cpu().numpy() for k, v in shared.items()}
            shared['camera_states'] = np.array(camera_joint_observation[..., offset:offset + 9])
            shared['target_states'] = np.array(target_joint_observation[..., offset:offset + 14])
        if self.full_observability:
            remaining = np.asarray(base.remaining_cargoes)
            shared.update(state=base.state(), obstacle_states=np.asarray(base.obstacle_states), remaining_cargoes=remaining,
                          remaining_cargo_counts=remaining.sum(axis=-1), awaiting_cargo_counts=np.asarray(base.awaiting_cargo_counts))
        if base.batched:
            camera_infos.update(shared)
            target_infos.update(shared)
        else:
            for info in list(camera_infos) + list(target_infos):
                info.update({k: (v.copy() if hasattr(v, 'copy') else v) for k, v in shared.items()})
        return results


class _AuxiliaryRewards(Wrapper):
    """Weighted sum of per-agent reward terms (shared implementation of the two wrappers below)."""

    ACCEPTABLE_KEYS = ()
    TEAM = 0   # index into the (camera, target) tuples

    def __init__(self, env, coefficients, reduction='none'):
        cls = type(self)
        assert _has_wrapper(env, RepeatedRewardIndividualDone), (
            f'You should use wrapper `{cls}` with wrapper `RepeatedRewardIndividualDone`. '
            f'Please wrap the environment with wrapper `RepeatedRewardIndividualDone` first. Got env = {env}.')
        assert not _has_wrapper(env, cls), f'You should not use wrapper `{cls}` more than once. Got env = {env}.'
        assert reduction in ('mean', 'sum', 'max', 'min', 'none'), f'Invalid reduction method {reduction}.'
        assert set(self.ACCEPTABLE_KEYS).issuperset(coefficients.keys()), (
            f'The coefficient mapping only accepts keys in {self.ACCEPTABLE_KEYS}. '
            f'Got list(coefficients.keys()) = {list(coefficients.keys())}.')
        self.coefficients = {}
        for key, coefficient in coefficients.items():
            assert callable(coefficient) or isinstance(coefficient, (float, int)), (
                f'The argument `coefficient` should be a callable function or a float number. '
                f'Got coefficients[{key!r}] = {coefficient!r}.')
            self.coefficients[key] = coefficient if not isinstance(coefficient, int) else float(coefficient)
        super().__init__(env)
        self.episode_id = -1
        self.reduction = reduction
        if 'soft_coverage_score' in self.coefficients:
            assert self.unwrapped.num_cameras > 0, "'soft_coverage_score' needs at least one camera."
            self.unwrapped.want_soft_coverage = True   # one more kernel per step (mate_b200_soft_coverage)

    def reset(self, **kwargs):
        self.episode_id += 1
        return self.env.reset(**kwargs)

    def _terms(self, base):
        raise NotImplementedError

    def step(self, action):
        observations, rewards, dones, infos = self.env.step(action)
        base = self.unwrapped
        terms = self._terms(base)                              # [B, N, K] in ACCEPTABLE_KEYS order
        num_agents = terms.shape[1]
        team_infos = infos[self.TEAM]
        if num_agents == 0:
            return observations, rewards, dones, infos
        raw_reward = terms[..., 0]
        reward = torch.zeros_like(raw_reward)
        per_agent = {}
        episode_step = base._aux['episode_step']  # pylint: disable=protected-access
        for key, coefficient in self.coefficients.items():
            value = terms[..., self.ACCEPTABLE_KEYS.index(key)]
            if callable(coefficient):
                if base.batched:   # vectorised call: agent ids [1, N], episode steps [B, 1], values [B, N]
                    agent_ids = torch.arange(num_agents, device=value.device).unsqueeze(0)
                    coefficient = coefficient(agent_ids, self.episode_id, episode_step.unsqueeze(-1), raw_reward, value)
                else:              # the reference's call, one agent at a time
                    coefficient = torch.tensor([[coefficient(i, self.episode_id, int(episode_step[0]), float(raw_reward[0, i]), float(value[0, i]))
                                                 for i in range(num_agents)]], dtype=value.dtype, device=value.device)
            reward = reward + coefficient * value
            per_agent.setdefault(key, value)
            per_agent[f'auxiliary_reward_{key}'] = value
            per_agent[f'reward_coefficient_{key}'] = coefficient if torch.is_tensor(coefficient) else torch.full_like(value, float(coefficient))
        per_agent['reward'] = reward
        if self.reduction in _REDUCERS:
            reward = _REDUCERS[self.reduction](reward).expand(-1, num_agents)
            per_agent['shared_reward'] = reward
        if base.batched:
            for key in self.coefficients:   # info.setdefault(key, ...): keep what an inner wrapper / the env already put there
                if key in team_infos:
                    per_agent.pop(key, None)
            team_infos.update(per_agent)
            team_reward = reward
        else:
            host = {k: v[0].tolist() for k, v in per_agent.items()}
            for i, info in enumerate(team_infos):
                for k, values in host.items():
                    if k in self.coefficients:
                        info.setdefault(k, values[i])
                    else:
                        info[k] = values[i]
            team_reward = reward[0].tolist()
        rewards = (team_reward, rewards[1]) if self.TEAM == 0 else (rewards[0], team_reward)
        return observations, rewards, dones, infos


class AuxiliaryCameraRewards(_AuxiliaryRewards):
    """mate/wrappers/auxiliary_camera_rewards.py: weighted sum of ``raw_reward``, ``coverage_rate``,
    ``real_coverage_rate``, ``mean_transport_rate``, ``soft_coverage_score``, ``num_tracked`` and ``baseline`` per
    camera."""

    ACCEPTABLE_KEYS = CAMERA_REWARD_KEYS
    TEAM = 0

    def _terms(self, base):
        return base.auxiliary_terms()[0][..., :len(CAMERA_REWARD_KEYS)]


class AuxiliaryTargetRewards(_AuxiliaryRewards):
    """mate/wrappers/auxiliary_target_rewards.py: weighted sum of ``raw_reward``, ``coverage_rate``,
    ``real_coverage_rate``, ``mean_transport_rate``, ``normalized_goal_distance``, ``sparse_delivery``,
    ``soft_coverage_score``, ``is_tracked``, ``is_colliding`` and ``baseline`` per target."""

    ACCEPTABLE_KEYS = TARGET_REWARD_KEYS
    TEAM = 1

    def _terms(self, base):
        return base.auxiliary_terms()[1][..., :len(TARGET_REWARD_KEYS)]


# ------------------------------------------------------------------------------------------------------------
# single-team wrapper with a batched built-in opponent.
# ------------------------------------------------------------------------------------------------------------

class MultiCamera(Wrapper):
    """mate/wrappers/single_team.py:281-292: a single-team multi-agent environment for the camera team; the targets
    are played by ``target_agent`` (``mate_b200.agents.GreedyTargetAgent``), whose whole team acts in one kernel per
    step on the GPU.  ``reset`` returns the camera joint observation, ``step(camera_joint_action)`` returns
    ``(camera_joint_observation, camera_reward, done, camera_infos)``."""

    def __init__(self, env, target_agent):
        from mate_b200 import agents  # pylint: disable=import-outside-toplevel

        assert isinstance(target_agent, agents.TargetAgentBase), (
            f'You should provide an instance of target agent. Got target_agent = {target_agent!r}.')
        assert not _has_wrapper(env, MultiCamera), f'You should not use wrapper `{type(self)}` more than once.'
        assert env.num_cameras > 0, 'There must be at least one camera in the environment.'
        super().__init__(env)
        self.opponent_agent = target_agent
        self.num_teammates, self.num_opponents = env.num_cameras, env.num_targets
        self.action_space = env.action_space.spaces[0]
        self.observation_space = env.observation_space.spaces[0]
        self.repeated_reward_individual_done = _has_wrapper(env, RepeatedRewardIndividualDone)
        self.opponent_joint_observation = None
        self.opponent_infos = None
        self._reset_mask = True
        target_agent.bind(self.unwrapped.sim)

    def reset(self, **kwargs):
        joint_observation, self.opponent_joint_observation = self.env.reset(**kwargs)
        self.opponent_infos = None
        self._reset_mask = True          # group_reset(opponent_agents, ...), single_team.py:209-219
        return joint_observation

    def step(self, action):
        base = self.unwrapped
        opponent_joint_action = self.opponent_agent.act(reset_mask=self._reset_mask)
        if not base.batched:
            opponent_joint_action = opponent_joint_action[0].double().cpu().numpy()
        (joint_observation, self.opponent_joint_observation), (reward, _), done, (infos, self.opponent_infos) = \
            self.env.step((action, opponent_joint_action))
        # finished episodes were auto-reset inside the step: their opponents start over on the new state
        self._reset_mask = base.sim.done if base.batched else None
        if self.repeated_reward_individual_done:
            done = done[0]
        return joint_observation, reward, done, infos

    def __str__(self):
        return f'<{type(self).__name__}(opponent={type(self.opponent_agent).__module__}.{type(self.opponent_agent).__name__}){self.env}>'


class MultiTarget(Wrapper):
    """mate/wrappers/single_team.py:295-306: a single-team multi-agent environment for the target team; the cameras
    are played by ``camera_agent`` (``mate_b200.agents.GreedyCameraAgent``), whose whole team acts in one kernel per
    step on the GPU.  ``reset`` returns the target joint observation, ``step(target_joint_action)`` returns
    ``(target_joint_observation, target_reward, done, target_infos)``."""

    def __init__(self, env, camera_agent):
        from mate_b200 import agents  # pylint: disable=import-outside-toplevel

        assert isinstance(camera_agent, agents.CameraAgentBase), (
            f'You should provide an instance of camera agent. Got camera_agent = {camera_agent!r}.')
        assert not _has_wrapper(env, MultiTarget) and not _has_wrapper(env, MultiCamera), (
            f'You should not use wrapper `{type(self)}` with another single-team wrapper.')
        assert env.num_cameras > 0, 'There must be at least one camera in the environment.'
        super().__init__(env)
        self.opponent_agent = camera_agent
        self.num_teammates, self.num_opponents = env.num_targets, env.num_cameras
        self.action_space = env.action_space.spaces[1]
        self.observation_space = env.observation_space.spaces[1]
        self.repeated_reward_individual_done = _has_wrapper(env, RepeatedRewardIndividualDone)
        self.opponent_joint_observation = None
        self.opponent_infos = None
        self._reset_mask = True
        camera_agent.bind(self.unwrapped.sim)

    def reset(self, **kwargs):
        self.opponent_joint_observation, joint_observation = self.env.reset(**kwargs)
        self.opponent_infos = None
        self._reset_mask = True
        return joint_observation

    def _tracked_flags(self):
        """Target flags of the cameras' current observations ``[B, Nc, Nt]`` (also right for environments that were
        auto-reset in the last step, whose auxiliary masks still describe the finished step)."""
        base = self.unwrapped
        nt = base.num_targets
        flags = base.sim.cam_obs[..., 22 + 4:22 + 5 * nt:5]   # camera observation layout, mate/constants.py:267-282
        return (flags > 0.5).to(torch.uint8)

    def step(self, action):
        base = self.unwrapped
        opponent_joint_action = self.opponent_agent.act(self._tracked_flags(), reset_mask=self._reset_mask)
        if not base.batched:
            opponent_joint_action = opponent_joint_action[0].double().cpu().numpy()
        (self.opponent_joint_observation, joint_observation), (_, reward), done, (self.opponent_infos, infos) = \
            self.env.step((opponent_joint_action, action))
        self._reset_mask = base.sim.done if base.batched else None
        if self.repeated_reward_individual_done:
            done = done[1]
        return joint_observation, reward, done, infos

    def __str__(self):
        return f'<{type(self).__name__}(opponent={type(self.opponent_agent).__module__}.{type(self.opponent_agent).__name__}){self.env}>'
