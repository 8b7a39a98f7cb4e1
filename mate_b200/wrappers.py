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

import numpy as np
import torch

from mate_b200 import _abi, spaces

TEAMS = ('both', 'camera', 'target', 'none')


class Wrapper:
    """gym.Wrapper-like forwarding: unknown public attributes resolve on the wrapped environment."""

    def __init__(self, env):
        self.env = env

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
        self.env.load_config(config=config)

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
