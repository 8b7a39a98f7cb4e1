"""Read-only entity views over the simulator state.

The reference keeps one Python object per camera / target / obstacle (``mate/entities.py``) and its wrappers and
agents read their attributes (``env.cameras[c].state()``, ``env.targets[t].is_colliding``, ...).  Here the entities
live in struct-of-arrays device memory; ``env.cameras`` / ``env.targets`` / ``env.obstacles`` return lightweight
views over ONE host snapshot of that state (taken once per step, on first access) with the reference's attribute
names and conventions.  In the reference-compatible single-environment mode every attribute has the reference's type
(``np.ndarray`` of shape ``[2]``, Python ``float`` / ``bool``); in batched mode the same attributes carry a leading
``[B]`` dimension.  The views cannot be written to: the simulation is advanced by ``env.step`` only.
"""

import numpy as np

from mate_b200 import constants as consts

__all__ = ['EntityView', 'ObstacleView', 'CameraView', 'TargetView', 'fov_sample_angles']


def _normalize_angle(angle):
    """mate/utils.py:155-158."""
    return (np.asarray(angle, dtype=np.float64) + 180.0) % 360.0 - 180.0


def fov_sample_angles(camera_xy, obstacles_xyr, max_sight_range):
    """Sample angles (degrees, ascending, unique) of a camera's field-of-view polyline as ``Camera.add_obstacles``
    builds it (mate/entities.py:362-479): the integer-degree grid and, for every obstacle of the camera's set
    (centre closer than ``max_sight_range + radius``), the four rays 0.01 degrees beside its tangents and the
    ``linspace`` lattice over its angular extent.  Returns ``None`` if the camera sits inside an obstacle (the
    reference then collapses the polyline to range 0)."""
    angles = [np.linspace(-180.0, +180.0, num=360, endpoint=False)]
    cx, cy = float(camera_xy[0]), float(camera_xy[1])
    for ox, oy, radius in np.asarray(obstacles_xyr, dtype=np.float64).reshape(-1, 3):
        dx, dy = ox - cx, oy - cy
        dist = float(np.hypot(dx, dy))
        if not dist < max_sight_range + radius:
            continue
        if radius > dist:
            return None
        bearing = float(np.degrees(np.arctan2(dy, dx)))
        half = float(np.degrees(np.arcsin(radius / dist)))
        left, right = bearing - half, bearing + half
        angles.append(np.array([left - 0.01, left + 0.01, right - 0.01, right + 0.01]))
        angles.append(np.linspace(left, right, num=max(16, int(2 * half)) + 1, endpoint=True))
    return np.unique(_normalize_angle(np.concatenate(angles)))


class EntityView:
    """Common part of the three views (mate/entities.py:27-100)."""

    def __init__(self, env, index):
        self._env = env
        self.index = index

    def _pick(self, array):
        """Entity `index` of a ``[B, N, ...]`` state array; the batch dimension is dropped in single-env mode."""
        out = array[:, self.index]
        return out if self._env.batched else out[0]

    def _scalar(self, array, cast=float):
        out = self._pick(array)
        return out if self._env.batched else cast(out)

    @property
    def x(self):
        return self.location[..., 0] if self._env.batched else float(self.location[0])

    @property
    def y(self):
        return self.location[..., 1] if self._env.batched else float(self.location[1])

    def distance(self, other):
        """Euclidean distance to another entity view or to a point (mate/entities.py:90-93)."""
        other_location = other.location if isinstance(other, EntityView) else np.asarray(other, dtype=np.float64)
        out = np.linalg.norm(self.location - other_location, axis=-1)
        return out if self._env.batched else float(out)

    def overlap(self, other, min_distance=0.0):
        """mate/entities.py:95-99."""
        return self.distance(other) * (1 + 1e-6) < self.radius + other.radius + min_distance

    def __setattr__(self, name, value):
        if not name.startswith('_') and name != 'index':
            raise AttributeError(f'{type(self).__name__} is a read-only view of the simulator state; use env.step / env.set_state')
        object.__setattr__(self, name, value)


class ObstacleView(EntityView):
    """mate/entities.py:102-155."""

    @property
    def location(self):
        return self._pick(self._env._snapshot()['obs_xyr'])[..., :2].copy()

    @property
    def radius(self):
        r = self._pick(self._env._snapshot()['obs_xyr'])[..., 2]
        return r.copy() if self._env.batched else float(r)

    @property
    def transmittance(self):
        return self._env.obstacle_transmittance

    def state(self, private=False):  # pylint: disable=unused-argument
        return self._pick(self._env._snapshot()['obs_xyr']).astype(np.float64)


class CameraView(EntityView):
    """mate/entities.py:235-543 (attributes, ``state``, ``sight_range_at``, ``boundary_between``)."""

    @property
    def location(self):
        return self._pick(self._env._snapshot()['cam_xy']).copy()

    @property
    def radius(self):
        return self._env.flat_config['camera_radius']

    @property
    def orientation(self):
        return self._scalar(self._env._snapshot()['cam_phi'])

    @property
    def viewing_angle(self):
        return self._scalar(self._env._snapshot()['cam_theta'])

    @property
    def sight_range(self):
        """sqrt(area_product / viewing_angle), mate/entities.py:285, 334."""
        cfg = self._env.flat_config
        area_product = cfg['camera_min_viewing_angle'] * cfg['camera_max_sight_range'] ** 2
        out = np.sqrt(area_product / self._pick(self._env._snapshot()['cam_theta']))
        return out if self._env.batched else float(out)

    max_sight_range = property(lambda self: self._env.flat_config['camera_max_sight_range'])
    min_viewing_angle = property(lambda self: self._env.flat_config['camera_min_viewing_angle'])
    rotation_step = property(lambda self: self._env.flat_config['camera_rotation_step'])
    zooming_step = property(lambda self: self._env.flat_config['camera_zooming_step'])
    action_space = property(lambda self: self._env.camera_action_space)

    @property
    def obstacles(self):
        """The obstacles of this camera's set (mate/entities.py:363-369), single-env mode."""
        mask = np.asarray(self._env.camera_obstacle_view_mask)
        if self._env.batched:
            raise NotImplementedError('per-camera obstacle sets as objects exist in single-environment mode only; use camera_obstacle_view_mask')
        return [o for i, o in enumerate(self._env.obstacles) if mask[self.index, i]]

    def state(self, private=False):
        """mate/entities.py:313-324."""
        rs, phi = self.sight_range, np.deg2rad(self.orientation)
        loc = self.location
        parts = [loc[..., 0], loc[..., 1], np.broadcast_to(self.radius, np.shape(rs)), rs * np.cos(phi), rs * np.sin(phi),
                 np.broadcast_to(self.viewing_angle, np.shape(rs))]
        if private:
            parts += [np.broadcast_to(v, np.shape(rs)) for v in (self.max_sight_range, self.rotation_step, self.zooming_step)]
        return np.stack([np.asarray(p, dtype=np.float64) for p in parts], axis=-1)

    def sight_range_at(self, angle, outer=False):
        """Range of the obstacle-occluded field of view at a bearing in degrees (mate/entities.py:507-511), evaluated
        by the CUDA polyline routine of the step kernel (``mate_b200_fov_range``); single-env mode."""
        if outer:
            raise NotImplementedError('the outer boundary is only evaluated inside mate_b200_soft_coverage (soft_coverage_score)')
        if self._env.batched:
            raise NotImplementedError('use env.sim.fov_range(env_index, camera_index, angle) in batched mode')
        angle = np.atleast_1d(np.asarray(angle, dtype=np.float64))
        out = self._env.sim.fov_range(np.zeros(len(angle), dtype=np.int32), np.full(len(angle), self.index, dtype=np.int32), angle).cpu().numpy()
        return out if out.size > 1 else float(out[0])

    def _polyline(self):
        """(phis, rhos) of the sampled polyline, closed like the reference's interp1d table (mate/entities.py:455-478)."""
        env = self._env
        snap = env._snapshot()
        phis = fov_sample_angles(snap['cam_xy'][0, self.index], snap['obs_xyr'][0] if env.obstacle_transmittance != 1.0 else np.zeros((0, 3)),
                                 self.max_sight_range)
        if phis is None:
            phis = np.arange(-180.0, 180.0, 90.0)
        rhos = np.atleast_1d(self.sight_range_at(phis))
        return np.append(phis, phis[0] + 360.0), np.append(rhos, rhos[0])

    def boundary_between(self, angle_left, angle_right, outer=False):
        """Vertices of the field-of-view polyline between two bearings (mate/entities.py:513-543); single-env mode."""
        assert 0.0 < angle_right - angle_left <= 360.0
        if outer:
            raise NotImplementedError('the outer boundary is only evaluated inside mate_b200_soft_coverage (soft_coverage_score)')
        left = float(_normalize_angle(angle_left))
        angle_left, angle_right = left, left + (angle_right - angle_left)
        phis_all, rhos_all = self._polyline()
        if angle_right <= +180.0:
            mask = np.logical_and(angle_left < phis_all, phis_all < angle_right)
            phis, rhos = phis_all[mask], rhos_all[mask]
        else:
            mask1 = np.logical_and(angle_left < phis_all, phis_all <= +180.0)
            mask2 = np.logical_and(phis_all > -180.0, phis_all < angle_right - 360.0)
            phis = np.concatenate([phis_all[mask1], phis_all[mask2]])
            rhos = np.concatenate([rhos_all[mask1], rhos_all[mask2]])
        phis = np.concatenate([[angle_left], phis, [angle_right]])
        rhos = np.concatenate([[self.sight_range_at(angle_left)], rhos, [self.sight_range_at(angle_right)]])
        return phis.astype(np.float64), rhos.astype(np.float64)


class TargetView(EntityView):
    """mate/entities.py:546-668 (attributes and ``state``)."""

    radius = consts.TARGET_RADIUS

    @property
    def location(self):
        return self._pick(self._env._snapshot()['tgt_xy']).copy()

    sight_range = property(lambda self: self._env.flat_config['target_sight_range'])
    transport_product = property(lambda self: self._env.flat_config['target_step_size'])
    action_space = property(lambda self: self._env.target_action_space)

    @property
    def capacity(self):
        return self._scalar(self._env._snapshot()['tgt_capacity'], int)

    @property
    def step_size(self):
        """transport_product / capacity (mate/entities.py:612-620)."""
        out = self.transport_product / self._pick(self._env._snapshot()['tgt_capacity'])
        return out if self._env.batched else float(out)

    @property
    def goal_bits(self):
        """The cargo weight at the index of the goal warehouse (mate/environment.py:1305-1309)."""
        snap = self._env._snapshot()
        goal, weight = self._pick(snap['tgt_goal']), self._pick(snap['tgt_weight'])
        bits = (np.arange(consts.NUM_WAREHOUSES) == np.asarray(goal)[..., None]) * np.asarray(weight)[..., None]
        return bits.astype(np.int64)

    @property
    def empty_bits(self):
        packed = np.asarray(self._pick(self._env._snapshot()['tgt_empty_bits']))
        return ((packed[..., None] >> np.arange(consts.NUM_WAREHOUSES)) & 1).astype(bool)

    @property
    def is_loaded(self):
        out = self.goal_bits.any(axis=-1)
        return out if self._env.batched else bool(out)

    @property
    def is_colliding(self):
        out = self._pick(np.asarray(self._env._aux_numpy('is_colliding'))).astype(bool)
        return out if self._env.batched else bool(out)

    def state(self, private=False):
        """mate/entities.py:631-637."""
        loc = self.location
        shape = loc.shape[:-1]
        parts = [loc[..., 0], loc[..., 1], np.broadcast_to(self.sight_range, shape), np.asarray(self.is_loaded, dtype=np.float64)]
        state = np.stack([np.asarray(p, dtype=np.float64) for p in parts], axis=-1)
        if private:
            extra = np.stack([np.asarray(self.step_size, dtype=np.float64), np.asarray(self.capacity, dtype=np.float64)], axis=-1)
            state = np.concatenate([state, extra, self.goal_bits.astype(np.float64), self.empty_bits.astype(np.float64)], axis=-1)
        return state.astype(np.float64)
