"""Host-side configuration logic (CPU)."""

import os

import numpy as np
import pytest

import golden_util as gu
from mate_b200 import config as mc

REFERENCE_ASSETS = '/root/reference/mate/assets'


@pytest.mark.parametrize('name', gu.trace_names())
def test_presets_flatten_like_the_reference(name):
    """flatten_config(preset) equals what the reference derived from its own YAML file
    (recorded in the golden fixture by oracle/gen_golden.py)."""
    g = gu.load(name)
    ref = gu.flat_config(g)
    mine = mc.flatten_config(mc.read_config(str(g['config_name'])))
    assert set(ref) == set(mine)
    for key, value in ref.items():
        if isinstance(value, np.ndarray):
            # the reference shuffles nothing at construction: *_ordered lists keep the file order
            np.testing.assert_array_equal(mine[key], value, err_msg=key)
        else:
            assert mine[key] == value, key


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ASSETS), reason='reference tree not mounted')
@pytest.mark.parametrize('name', sorted(mc.PRESETS))
def test_presets_equal_reference_yaml(name):
    import yaml

    with open(os.path.join(REFERENCE_ASSETS, name), encoding='UTF-8') as file:
        ref = yaml.load(file, yaml.SafeLoader)
    assert mc.preset(name) == ref


def test_overrides_and_defaults():
    cfg = mc.read_config('MATE-4v8-9.yaml', max_episode_steps=50, camera={'max_sight_range': 900.0})
    assert cfg['max_episode_steps'] == 50
    assert cfg['camera']['max_sight_range'] == 900.0
    assert cfg['camera']['rotation_step'] == 5.0
    flat = mc.flatten_config(cfg)
    assert flat['num_high_capacity_targets'] == 4 and flat['camera_max_sight_range'] == 900.0
    nav = mc.flatten_config(mc.read_config('MATE-Navigation.yaml'))
    assert nav['num_cameras'] == 0 and nav['reward_sparse'] == 1 and nav['bounty_factor'] == 1.0
    assert nav['targets_start_with_cargoes'] == 0


def test_validation_errors():
    with pytest.raises(ValueError):
        mc.read_config('MATE-4v8-9.yaml', max_episode_steps=0)
    with pytest.raises(ValueError):
        mc.read_config('MATE-4v8-9.yaml', reward_type='shaped')
    with pytest.raises(ValueError):
        mc.read_config({'num_cargoes_per_target': 8})
    with pytest.raises(ValueError):
        mc.read_config('MATE-4v8-9.yaml', num_cargoes_per_target=2)
    with pytest.raises(ValueError):
        mc.read_config('MATE-4v9-9.yaml')
    with pytest.raises(ValueError):
        mc.read_config('MATE-4v8-9.yaml', camera={'rotation_step': -1.0})


def test_dict_and_file_configs(tmp_path):
    import json

    cfg = mc.preset('MATE-2v4-9')
    path = tmp_path / 'custom.json'
    path.write_text(json.dumps(cfg))
    a = mc.flatten_config(mc.read_config(str(path)))
    b = mc.flatten_config(mc.read_config(cfg))
    for key in a:
        np.testing.assert_array_equal(a[key], b[key])
    # fixed camera locations become degenerate ranges
    assert (a['camera_location_ranges'] == [[-300, -300, -300, -300], [300, 300, 300, 300]]).all()
