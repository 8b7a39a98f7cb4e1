"""Layout constants of the observation / state vectors (the reference's
``mate/constants.py:52-191, 267-368``), restated for the batched simulator."""

import functools

import numpy as np

from mate_b200 import spaces

TERRAIN_SIZE = 1000.0
TERRAIN_WIDTH = 2.0 * TERRAIN_SIZE
WAREHOUSE_RADIUS = 0.075 * TERRAIN_SIZE
WAREHOUSES = (TERRAIN_SIZE - WAREHOUSE_RADIUS) * np.array([[+1.0, +1.0], [-1.0, +1.0], [-1.0, -1.0], [+1.0, -1.0]])
NUM_WAREHOUSES = len(WAREHOUSES)
MAX_CAMERA_VIEWING_ANGLE = 180.0
TARGET_RADIUS = 0.0

PRESERVED_DIM = 3 + 1 + 2 * NUM_WAREHOUSES + 1
OBSERVATION_OFFSET = PRESERVED_DIM
CAMERA_STATE_DIM_PUBLIC = 6
CAMERA_STATE_DIM_PRIVATE = 9
TARGET_STATE_DIM_PUBLIC = 4
TARGET_STATE_DIM_PRIVATE = 6 + 2 * NUM_WAREHOUSES
OBSTACLE_STATE_DIM = 3
CAMERA_ACTION_DIM = 2
TARGET_ACTION_DIM = 2

_INF = np.inf
_T2 = 2.0 * TERRAIN_SIZE
PRESERVED_LOW = np.array([0.0] * 4 + [-_T2] * (2 * NUM_WAREHOUSES) + [0.0])
PRESERVED_HIGH = np.array([_INF] * 4 + [_T2] * (2 * NUM_WAREHOUSES) + [TERRAIN_SIZE])
CAMERA_PUBLIC_LOW = np.array([-_T2, -_T2, 0.0, -TERRAIN_WIDTH, -TERRAIN_WIDTH, 0.0])
CAMERA_PUBLIC_HIGH = np.array([_T2, _T2, TERRAIN_SIZE, TERRAIN_WIDTH, TERRAIN_WIDTH, MAX_CAMERA_VIEWING_ANGLE])
CAMERA_PRIVATE_LOW = np.append(CAMERA_PUBLIC_LOW, [0.0, 0.0, 0.0])
CAMERA_PRIVATE_HIGH = np.append(CAMERA_PUBLIC_HIGH, [TERRAIN_WIDTH, MAX_CAMERA_VIEWING_ANGLE, MAX_CAMERA_VIEWING_ANGLE])
TARGET_PUBLIC_LOW = np.array([-_T2, -_T2, 0.0, -1.0])
TARGET_PUBLIC_HIGH = np.array([_T2, _T2, TERRAIN_WIDTH, 1.0])
TARGET_PRIVATE_LOW = np.concatenate([TARGET_PUBLIC_LOW, [0.0, 1.0], [0.0] * NUM_WAREHOUSES, [-1.0] * NUM_WAREHOUSES])
TARGET_PRIVATE_HIGH = np.concatenate([TARGET_PUBLIC_HIGH, [TERRAIN_WIDTH, 2.0], [_INF] * NUM_WAREHOUSES, [1.0] * NUM_WAREHOUSES])
OBSTACLE_LOW = np.array([-_T2, -_T2, 0.0])
OBSTACLE_HIGH = np.array([_T2, _T2, TERRAIN_SIZE])


def _flagged(low, high, reps):
    return np.tile(np.append(low, -1.0), reps), np.tile(np.append(high, 1.0), reps)


@functools.lru_cache(maxsize=None)
def camera_observation_space_of(num_cameras, num_targets, num_obstacles):
    parts = [(PRESERVED_LOW, PRESERVED_HIGH), (CAMERA_PRIVATE_LOW, CAMERA_PRIVATE_HIGH),
             _flagged(TARGET_PUBLIC_LOW, TARGET_PUBLIC_HIGH, num_targets),
             _flagged(OBSTACLE_LOW, OBSTACLE_HIGH, num_obstacles),
             _flagged(CAMERA_PUBLIC_LOW, CAMERA_PUBLIC_HIGH, num_cameras)]
    return spaces.Box(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))


@functools.lru_cache(maxsize=None)
def target_observation_space_of(num_cameras, num_targets, num_obstacles):
    parts = [(PRESERVED_LOW, PRESERVED_HIGH), (TARGET_PRIVATE_LOW, TARGET_PRIVATE_HIGH),
             _flagged(CAMERA_PUBLIC_LOW, CAMERA_PUBLIC_HIGH, num_cameras),
             _flagged(OBSTACLE_LOW, OBSTACLE_HIGH, num_obstacles),
             _flagged(TARGET_PUBLIC_LOW, TARGET_PUBLIC_HIGH, num_targets)]
    return spaces.Box(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))


@functools.lru_cache(maxsize=None)
def camera_observation_indices_of(num_cameras, num_targets, num_obstacles):
    return np.cumsum([0, PRESERVED_DIM, CAMERA_STATE_DIM_PRIVATE, num_targets * (TARGET_STATE_DIM_PUBLIC + 1),
                      num_obstacles * (OBSTACLE_STATE_DIM + 1), num_cameras * (CAMERA_STATE_DIM_PUBLIC + 1)])


@functools.lru_cache(maxsize=None)
def target_observation_indices_of(num_cameras, num_targets, num_obstacles):
    return np.cumsum([0, PRESERVED_DIM, TARGET_STATE_DIM_PRIVATE, num_cameras * (CAMERA_STATE_DIM_PUBLIC + 1),
                      num_obstacles * (OBSTACLE_STATE_DIM + 1), num_targets * (TARGET_STATE_DIM_PUBLIC + 1)])


def _slices(indices, opponent_dim, teammate_dim):
    return {
        'preserved_data': slice(indices[0], indices[1]),
        'self_state': slice(indices[1], indices[2]),
        'opponent_states_with_mask': slice(indices[2], indices[3]),
        'opponent_mask': slice(indices[2] + opponent_dim, indices[3], opponent_dim + 1),
        'obstacle_states_with_mask': slice(indices[3], indices[4]),
        'obstacle_mask': slice(indices[3] + OBSTACLE_STATE_DIM, indices[4], OBSTACLE_STATE_DIM + 1),
        'teammate_states_with_mask': slice(indices[4], indices[5]),
        'teammate_mask': slice(indices[4] + teammate_dim, indices[5], teammate_dim + 1),
    }


@functools.lru_cache(maxsize=None)
def camera_observation_slices_of(num_cameras, num_targets, num_obstacles):
    return _slices(camera_observation_indices_of(num_cameras, num_targets, num_obstacles),
                   TARGET_STATE_DIM_PUBLIC, CAMERA_STATE_DIM_PUBLIC)


@functools.lru_cache(maxsize=None)
def target_observation_slices_of(num_cameras, num_targets, num_obstacles):
    return _slices(target_observation_indices_of(num_cameras, num_targets, num_obstacles),
                   CAMERA_STATE_DIM_PUBLIC, TARGET_STATE_DIM_PUBLIC)


CAMERA_DEFAULT_ACTION = np.asarray([0.0, 0.0], dtype=np.float64)
TARGET_DEFAULT_ACTION = np.asarray([0.0, 0.0], dtype=np.float64)


def _team_index(team):
    """0 = camera team, 1 = target team; accepts the reference's ``Team`` enum (``.value``), an int or a name."""
    if isinstance(team, str):
        return {'camera': 0, 'target': 1}[team.lower()]
    return int(getattr(team, 'value', team))


def observation_space_of(team, num_cameras, num_targets, num_obstacles):
    """mate/constants.py:257-265."""
    return (camera_observation_space_of, target_observation_space_of)[_team_index(team)](num_cameras, num_targets, num_obstacles)


def observation_indices_of(team, num_cameras, num_targets, num_obstacles):
    """mate/constants.py:304-312."""
    return (camera_observation_indices_of, target_observation_indices_of)[_team_index(team)](num_cameras, num_targets, num_obstacles)


def observation_slices_of(team, num_cameras, num_targets, num_obstacles):
    """mate/constants.py:361-369."""
    return (camera_observation_slices_of, target_observation_slices_of)[_team_index(team)](num_cameras, num_targets, num_obstacles)


def _coordinate_mask(private_dim, first_dim, first_count, num_obstacles, last_dim, last_count):
    preserved = np.zeros(PRESERVED_DIM, dtype=bool)
    preserved[-1 - 2 * NUM_WAREHOUSES:-1] = True           # the warehouse locations
    own = np.zeros(private_dim, dtype=bool)                # the agent's own state is not a relative coordinate

    def flagged(dim, count):                               # first two entries (x, y) of every flagged entity block
        block = np.zeros(dim + 1, dtype=bool)
        block[:2] = True
        return np.tile(block, count)

    return np.concatenate([preserved, own, flagged(first_dim, first_count), flagged(OBSTACLE_STATE_DIM, num_obstacles),
                           flagged(last_dim, last_count)])


@functools.lru_cache(maxsize=None)
def camera_coordinate_mask_of(num_cameras, num_targets, num_obstacles):
    """True where an entry of a camera's observation is a coordinate of another entity or of a warehouse
    (mate/constants.py:372-398)."""
    return _coordinate_mask(CAMERA_STATE_DIM_PRIVATE, TARGET_STATE_DIM_PUBLIC, num_targets, num_obstacles,
                            CAMERA_STATE_DIM_PUBLIC, num_cameras)


@functools.lru_cache(maxsize=None)
def target_coordinate_mask_of(num_cameras, num_targets, num_obstacles):
    """True where an entry of a target's observation is a coordinate of another entity or of a warehouse
    (mate/constants.py:401-427)."""
    return _coordinate_mask(TARGET_STATE_DIM_PRIVATE, CAMERA_STATE_DIM_PUBLIC, num_cameras, num_obstacles,
                            TARGET_STATE_DIM_PUBLIC, num_targets)


def coordinate_mask_of(team, num_cameras, num_targets, num_obstacles):
    """mate/constants.py:430-440."""
    return (camera_coordinate_mask_of, target_coordinate_mask_of)[_team_index(team)](num_cameras, num_targets, num_obstacles)
