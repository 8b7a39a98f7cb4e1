"""Configuration handling of the batched simulator.

Mirrors the semantics of the reference's ``read_config`` / ``validate_config``
(mate/environment.py:113-269): a configuration is a dict, or a path to a JSON/YAML file
looked up in the working directory and then among the bundled presets; keyword overrides
are deep-merged; missing keys get the reference's defaults and the same validation errors
(``ValueError``) are raised.  The 17 preset layouts of the reference
(mate/assets/MATE-*.yaml) are generated programmatically in :data:`PRESETS` -- a preset is
addressed by the reference's file name (``'MATE-4v8-9.yaml'``).

:func:`flatten_config` turns a validated configuration into the flat POD the C ABI takes
(``MateConfig`` in ``include/mate_b200.h``).
"""

import copy
import difflib
import json
import os
import warnings
from typing import Any, Dict, Mapping, Optional, Union

import numpy as np

__all__ = ['PRESETS', 'DEFAULT_CONFIG_FILE', 'preset', 'read_config', 'validate_config', 'flatten_config']

NUM_WAREHOUSES = 4
DEFAULT_CONFIG_FILE = 'MATE-4v8-9.yaml'   # mate/environment.py:38

# entity defaults (mate/entities.py:16-22, 248-254, 563-566)
CAMERA_DEFAULTS = {
    'radius': 40.0, 'min_viewing_angle': 90.0, 'max_sight_range': 500.0,
    'rotation_step': 5.0, 'zooming_step': 2.5,
}
TARGET_DEFAULTS = {'sight_range': 500.0, 'step_size': 10.0}
OBSTACLE_DEFAULT_TRANSMITTANCE = 0.0


# ----------------------------------------------------------------------------- presets
def _quadrants(lo, hi):
    """Boxes [x_lo, x_hi, y_lo, y_hi] in the four quadrants, anticlockwise from (+, +) via (+, -)."""
    return [[lo, hi, lo, hi], [lo, hi, -hi, -lo], [-hi, -lo, -hi, -lo], [-hi, -lo, lo, hi]]


def _edge_strips(coord, half):
    return [[coord, coord, -half, half], [-half, half, coord, coord],
            [-coord, -coord, -half, half], [-half, half, -coord, -coord]]


_CENTER = [-200, 200, -200, 200]


def _camera_block(num_cameras):
    common = {'min_viewing_angle': 30.0, 'max_sight_range': 1500.0, 'rotation_step': 5.0,
              'zooming_step': 2.5, 'radius': 40.0}
    if num_cameras == 1:
        return {'location': [[0, 0]], **common}
    if num_cameras == 2:
        return {'location': [[-300, -300], [300, 300]], **common}
    if num_cameras == 4:
        return {'location_random_range': _quadrants(500, 800), **common}
    if num_cameras == 8:
        ring = [[500, 600, -100, 100], [-100, 100, 500, 600], [-600, -500, -100, 100], [-100, 100, -600, -500]]
        return {'location_random_range': _quadrants(700, 850) + ring, **dict(common, max_sight_range=1000.0)}
    raise ValueError(num_cameras)


def _obstacle_block(num_obstacles):
    if num_obstacles == 9:
        ranges = _quadrants(200, 800) + _edge_strips(900, 500) + [list(_CENTER)]
    elif num_obstacles == 32:
        ranges = (_quadrants(200, 800) * 2 + _edge_strips(900, 500) * 2
                  + [list(_CENTER) for _ in range(8)] + [[-900, 900, -900, 900] for _ in range(8)])
    else:
        raise ValueError(num_obstacles)
    return {'location_random_range': ranges, 'radius_random_range': [25.0, 100.0], 'transmittance': 0.1}


def _make_preset(num_cameras, num_targets, num_obstacles):
    cfg = {
        'name': f'MultiAgentTracking({num_cameras}v{num_targets}, {num_obstacles})',
        'max_episode_steps': 10000,
        'num_cargoes_per_target': 8,
        'high_capacity_target_split': 0.5,
        'targets_start_with_cargoes': True,
        'bounty_factor': 1.0,
        'shuffle_entities': True,
        'reward_type': 'dense',
    }
    if num_cameras:
        cfg['camera'] = _camera_block(num_cameras)
    cfg['target'] = {'location_random_range': [list(_CENTER) for _ in range(num_targets)],
                     'step_size': 20.0, 'sight_range': 500.0}
    if num_obstacles:
        cfg['obstacle'] = _obstacle_block(num_obstacles)
    return cfg


def _build_presets():
    presets = {}
    for nc, nt in ((1, 1), (1, 2), (2, 2), (2, 4), (4, 2), (4, 4), (4, 8), (8, 8)):
        for no in (0, 9):
            cfg = _make_preset(nc, nt, no)
            if (nc, nt) == (1, 1):   # the reference's 1v1 files leave these to the defaults
                del cfg['high_capacity_target_split']
                if no == 0:
                    del cfg['shuffle_entities']
            presets[f'MATE-{nc}v{nt}-{no}.yaml'] = cfg
    navigation = _make_preset(0, 8, 32)
    navigation.update(targets_start_with_cargoes=False, reward_type='sparse')
    del navigation['bounty_factor']
    presets['MATE-Navigation.yaml'] = navigation
    presets['MATE.yaml'] = presets['MATE-4v8-9.yaml']   # symlink in the reference's assets
    return presets


PRESETS = _build_presets()


def preset(name: str) -> Dict[str, Any]:
    """A deep copy of the named preset (``'MATE-4v8-9.yaml'`` or ``'MATE-4v8-9'``)."""
    key = name if name in PRESETS else name + '.yaml'
    if key not in PRESETS:
        raise ValueError(f'Unknown preset "{name}". Did you mean: "{_did_you_mean(name)}"?')
    return copy.deepcopy(PRESETS[key])


def _did_you_mean(path: str) -> str:
    names = list(PRESETS)
    for ext in ('*.yaml', '*.yml', '*.json'):
        import glob  # pylint: disable=import-outside-toplevel

        names.extend(os.path.basename(p) for p in glob.glob(os.path.join(os.getcwd(), ext)))
    close = difflib.get_close_matches(os.path.basename(str(path)), names, n=1, cutoff=0.0)
    return close[0] if close else names[0]


def _warn(msg):
    warnings.warn(msg, stacklevel=3)


def _deep_update(base: Dict[str, Any], override: Mapping[str, Any]) -> Dict[str, Any]:
    out = copy.deepcopy(dict(base))
    for key, value in override.items():
        if isinstance(out.get(key), dict) and isinstance(value, Mapping):
            out[key] = _deep_update(out[key], value)
        else:
            out[key] = copy.deepcopy(value)
    return out


def _load_file(path: str) -> Dict[str, Any]:
    ext = os.path.splitext(path)[1].lower()
    if ext not in ('.json', '.yaml', '.yml'):
        raise ValueError(
            'The configuration should be a dictionary mapping or a path to a readable JSON/YAML file. '
            f'Got {path!r}.'
        )
    with open(path, encoding='UTF-8') as file:
        if ext == '.json':
            return json.load(file)
        import yaml  # pylint: disable=import-outside-toplevel

        return yaml.load(file, yaml.SafeLoader)


def read_config(config_or_path: Optional[Union[Mapping[str, Any], str, os.PathLike]] = None,
                **kwargs) -> Dict[str, Any]:
    """Load a configuration from a dict, a JSON/YAML file or a preset name, apply keyword
    overrides, fill defaults and validate (mate/environment.py:113-193)."""
    if config_or_path is None:
        config = {}
    elif isinstance(config_or_path, Mapping):
        config = copy.deepcopy(dict(config_or_path))
    else:
        path = os.fspath(config_or_path)
        if os.path.exists(path):
            config = _load_file(path)
        elif os.path.isfile(os.path.join(os.getcwd(), path)):
            config = _load_file(os.path.join(os.getcwd(), path))
        elif os.path.basename(path) in PRESETS:
            config = preset(os.path.basename(path))
        else:
            raise ValueError(
                f'Cannot found the configuration file "{path}". Did you mean: "{_did_you_mean(path)}"?'
            )
        if not isinstance(config, Mapping):
            raise ValueError(f'The configuration file {path!r} does not hold a mapping.')
        config = dict(config)
    config = _deep_update(config, kwargs)
    validate_config(config)
    for entity in ('camera', 'obstacle', 'target'):
        config.setdefault(entity, {})
    return config


def _num_entities(sub: Mapping[str, Any]) -> int:
    return len(sub.get('location', [])) + len(sub.get('location_random_range', []))


def validate_config(config: Dict[str, Any]) -> None:
    """Fill defaults and raise ``ValueError`` like the reference (mate/environment.py:196-269)."""
    config.setdefault('max_episode_steps', 10000)
    if config['max_episode_steps'] <= 0:
        raise ValueError('`max_episode_steps` must be a positive integer.')
    config.setdefault('reward_type', 'dense')
    if config['reward_type'] not in ('dense', 'sparse'):
        raise ValueError(f'Invalid reward type {config["reward_type"]}. Expect one of {("dense", "sparse")}')
    if 'target' not in config:
        raise ValueError('Missing key "target". There must be at least one target in the environment.')
    if _num_entities(config['target']) == 0:
        raise ValueError('There must be at least one target in the environment.')
    if 'num_cargoes_per_target' not in config:
        raise ValueError('Missing key "num_cargoes_per_target".')
    if config['num_cargoes_per_target'] < NUM_WAREHOUSES:
        raise ValueError(
            f'`num_cargoes_per_target` should be no less than {NUM_WAREHOUSES}. '
            f'Got {config["num_cargoes_per_target"]}.'
        )
    config.setdefault('high_capacity_target_split', 0.5)
    if not 0.0 <= config['high_capacity_target_split'] <= 1.0:
        raise ValueError(
            f'`high_capacity_target_split` must be between 0 and 1. Got {config["high_capacity_target_split"]}.'
        )
    config['targets_start_with_cargoes'] = bool(config.get('targets_start_with_cargoes', True))
    config.setdefault('bounty_factor', 1.0)
    if not config['bounty_factor'] >= 0.0:
        raise ValueError(f'`bounty_factor` must be a non-negative number. Got {config["bounty_factor"]}.')
    config['shuffle_entities'] = bool(config.get('shuffle_entities', True))
    for entity, defaults in (('camera', CAMERA_DEFAULTS), ('target', TARGET_DEFAULTS)):
        if entity in config:
            for key, default in defaults.items():
                config[entity].setdefault(key, default)
                if not config[entity][key] > 0.0:
                    raise ValueError(f'`{entity}/{key}` must be a positive number. Got {config[entity][key]}.')


def _ranges(sub: Mapping[str, Any]) -> np.ndarray:
    """[N, 4] = (x_low, x_high, y_low, y_high); fixed locations first, like make_from_config
    (mate/environment.py:380-390)."""
    rows = []
    for loc in sub.get('location', []):
        x, y = (float(v) for v in np.asarray(loc, dtype=np.float64).ravel()[:2])
        rows.append([x, x, y, y])
    for rng in sub.get('location_random_range', []):
        if isinstance(rng, Mapping):
            low, high = np.asarray(rng['low'], dtype=np.float64), np.asarray(rng['high'], dtype=np.float64)
            rows.append([low[0], high[0], low[1], high[1]])
        elif hasattr(rng, 'low') and hasattr(rng, 'high'):
            rows.append([float(rng.low[0]), float(rng.high[0]), float(rng.low[1]), float(rng.high[1])])
        else:
            flat = [float(v) for v in rng]
            rows.append([flat[0], flat[1], flat[2], flat[3]])   # low = [0::2], high = [1::2]
    return np.asarray(rows, dtype=np.float64).reshape(len(rows), 4)


def flatten_config(config: Mapping[str, Any]) -> Dict[str, Any]:
    """The flat, validated POD form of a configuration (fields of ``MateConfig``)."""
    camera = config.get('camera', {}) or {}
    target = config['target']
    obstacle = config.get('obstacle', {}) or {}
    cam_ranges, tgt_ranges, obs_ranges = _ranges(camera), _ranges(target), _ranges(obstacle)
    nc, nt, no = len(cam_ranges), len(tgt_ranges), len(obs_ranges)
    if 'radius_random_range' in obstacle:
        rr = obstacle['radius_random_range']
        if isinstance(rr, Mapping):
            r_low, r_high = float(np.ravel(rr['low'])[0]), float(np.ravel(rr['high'])[0])
        elif hasattr(rr, 'low'):
            r_low, r_high = float(np.ravel(rr.low)[0]), float(np.ravel(rr.high)[0])
        else:
            r_low, r_high = float(rr[0]), float(rr[1])
    elif 'radius' in obstacle:
        r_low = r_high = float(obstacle['radius'])
    else:
        r_low = r_high = 0.0
        if no:
            raise ValueError('You should specify either a fixed radius or a random range for the obstacle radius.')
    split = min(max(0.0, float(config.get('high_capacity_target_split', 0.5))), 1.0)
    transmittance = float(obstacle.get('transmittance', OBSTACLE_DEFAULT_TRANSMITTANCE))
    if not 0.0 <= transmittance <= 1.0:
        raise ValueError(f'The argument `transmittance` within the range of [0.0, 1.0]. Got transmittance = {transmittance}.')

    def cam(key):
        return float(camera.get(key, CAMERA_DEFAULTS[key]))

    if nc and not 0.0 < cam('min_viewing_angle') <= 180.0:
        raise ValueError(f'`camera/min_viewing_angle` must be within (0, 180]. Got {cam("min_viewing_angle")}.')
    return {
        'num_cameras': nc,
        'num_targets': nt,
        'num_obstacles': no,
        'max_episode_steps': int(config['max_episode_steps']),
        'num_cargoes_per_target': int(config['num_cargoes_per_target']),
        'num_high_capacity_targets': int(nt * split),                      # environment.py:1527-1535
        'targets_start_with_cargoes': int(bool(config.get('targets_start_with_cargoes', True))),
        'shuffle_entities': int(bool(config.get('shuffle_entities', True))),
        'reward_sparse': int(config.get('reward_type', 'dense') == 'sparse'),
        'bounty_factor': max(0.0, float(config.get('bounty_factor', 1.0))),
        'camera_radius': cam('radius'),
        'camera_min_viewing_angle': cam('min_viewing_angle'),
        'camera_max_sight_range': cam('max_sight_range'),
        'camera_rotation_step': cam('rotation_step'),
        'camera_zooming_step': cam('zooming_step'),
        'target_step_size': float(target.get('step_size', TARGET_DEFAULTS['step_size'])),
        'target_sight_range': float(target.get('sight_range', TARGET_DEFAULTS['sight_range'])),
        'obstacle_transmittance': transmittance,
        'obstacle_radius_low': r_low,
        'obstacle_radius_high': r_high,
        'camera_location_ranges': cam_ranges,
        'target_location_ranges': tgt_ranges,
        'obstacle_location_ranges': obs_ranges,
    }
