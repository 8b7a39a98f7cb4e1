"""Helpers shared by the parity tests: load the reference-generated fixtures in
tests/golden/ (written by oracle/gen_golden.py) into the flat config / state-array form
the oracle and the CUDA simulator take."""

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def trace_names():
    return sorted(
        os.path.basename(p)[:-4]
        for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))
        if not p.endswith('_resets.npz') and not p.endswith('_resetsamples.npz')
        and not os.path.basename(p).startswith(('wrappers_', 'aux_', 'agents_'))
    )


def reset_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*_resets.npz')))


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))


def flat_config(g):
    nc, nt, no = (int(x) for x in g['cfg_counts'])
    cam = g['cfg_camera']
    tgt = g['cfg_target']
    return {
        'num_cameras': nc,
        'num_targets': nt,
        'num_obstacles': no,
        'max_episode_steps': int(g['cfg_max_episode_steps']),
        'num_cargoes_per_target': int(g['cfg_num_cargoes_per_target']),
        'num_high_capacity_targets': int(g['cfg_num_high_capacity_targets']),
        'targets_start_with_cargoes': int(g['cfg_targets_start_with_cargoes']),
        'shuffle_entities': int(g['cfg_shuffle_entities']),
        'reward_sparse': int(g['cfg_reward_sparse']),
        'bounty_factor': float(g['cfg_bounty_factor']),
        'camera_radius': float(cam[0]),
        'camera_min_viewing_angle': float(cam[1]),
        'camera_max_sight_range': float(cam[2]),
        'camera_rotation_step': float(cam[3]),
        'camera_zooming_step': float(cam[4]),
        'target_step_size': float(tgt[0]),
        'target_sight_range': float(tgt[1]),
        'obstacle_transmittance': float(g['cfg_transmittance']),
        'obstacle_radius_low': float(g['cfg_obstacle_radius_range'][0]),
        'obstacle_radius_high': float(g['cfg_obstacle_radius_range'][1]),
        'camera_location_ranges': np.asarray(g['cfg_camera_ranges'], dtype=np.float64).reshape(nc, 4),
        'target_location_ranges': np.asarray(g['cfg_target_ranges'], dtype=np.float64).reshape(nt, 4),
        'obstacle_location_ranges': np.asarray(g['cfg_obstacle_ranges'], dtype=np.float64).reshape(no, 4),
    }


def empty_bits_to_int(bits):
    bits = np.asarray(bits)
    return (bits.astype(np.int32) << np.arange(4, dtype=np.int32)).sum(axis=-1).astype(np.int32)


def state_arrays(g, prefix='init_', index=None):
    """MateStateView arrays (batch of 1) from a golden record."""

    def get(key):
        value = g[prefix + key]
        return value if index is None else value[index]

    nc = int(g['cfg_counts'][0])
    nt = int(g['cfg_counts'][1])
    no = int(g['cfg_counts'][2])
    return {
        'cam_xy': np.asarray(get('cam_xy'), dtype=np.float64).reshape(1, nc, 2),
        'cam_phi': np.asarray(get('cam_phi'), dtype=np.float64).reshape(1, nc),
        'cam_theta': np.asarray(get('cam_theta'), dtype=np.float64).reshape(1, nc),
        'tgt_xy': np.asarray(get('tgt_xy'), dtype=np.float64).reshape(1, nt, 2),
        'obs_xyr': np.asarray(get('obs_xyr'), dtype=np.float64).reshape(1, no, 3),
        'tgt_capacity': np.asarray(get('tgt_capacity'), dtype=np.int32).reshape(1, nt),
        'tgt_goal': np.asarray(get('tgt_goal'), dtype=np.int32).reshape(1, nt),
        'tgt_weight': np.asarray(get('tgt_goal_weight'), dtype=np.int32).reshape(1, nt),
        'tgt_bounty': np.asarray(get('bounties'), dtype=np.int32).reshape(1, nt),
        'tgt_empty_bits': empty_bits_to_int(get('tgt_empty_bits')).reshape(1, nt),
        'remaining': np.asarray(get('remaining'), dtype=np.int32).reshape(1, 4, 4),
        'awaiting': np.asarray(get('awaiting'), dtype=np.int32).reshape(1, 4),
        'num_delivered': np.asarray(get('num_delivered'), dtype=np.int32).reshape(1),
        'episode_step': np.asarray(get('episode_step'), dtype=np.int32).reshape(1),
        'episode_id': np.zeros(1, dtype=np.int32),
        'episode_reward': np.array(
            [[float(get('ep_reward')), float(get('delayed_ep_reward'))]], dtype=np.float64
        ),
    }


def stack_states(list_of_arrays):
    keys = list_of_arrays[0].keys()
    return {k: np.concatenate([a[k] for a in list_of_arrays], axis=0) for k in keys}


def norm_angle(a):
    return (a + 180.0) % 360.0 - 180.0


def in_tangent_sliver(cam_xy, tgt_xy, obs_xyr, rmax, width=0.0101):
    """True if the bearing camera->target lies within `width` degrees of a ray that is exactly
    tangent to a visible obstacle.  There the reference's FOV polyline sample is decided by
    rounding noise (see DESIGN.md "tangent rays"), so a camera->target mask bit may legitimately
    differ from the recorded reference run."""
    rel = np.asarray(tgt_xy) - np.asarray(cam_xy)
    bearing = np.rad2deg(np.arctan2(rel[1], rel[0]))
    for x, y, r in obs_xyr:
        orel = np.array([x, y]) - np.asarray(cam_xy)
        d = np.sqrt(orel @ orel)
        if not d < rmax + r or r > d:
            continue
        ang = np.rad2deg(np.arctan2(orel[1], orel[0]))
        half = np.rad2deg(np.arcsin(r / d))
        for a in (ang - half, ang + half):
            diff = abs(norm_angle(bearing - a))
            if diff <= width:
                return True
    return False


def step_state_arrays(g, k):
    """MateStateView arrays (batch of 1) of the reference AFTER step k of a trace."""
    nc, nt, no = (int(x) for x in g['cfg_counts'])
    return {
        'cam_xy': np.asarray(g['init_cam_xy'], dtype=np.float64).reshape(1, nc, 2),
        'cam_phi': g['step_cam_phi'][k].reshape(1, nc).astype(np.float64),
        'cam_theta': g['step_cam_theta'][k].reshape(1, nc).astype(np.float64),
        'tgt_xy': g['step_tgt_xy'][k].reshape(1, nt, 2).astype(np.float64),
        'obs_xyr': np.asarray(g['init_obs_xyr'], dtype=np.float64).reshape(1, no, 3),
        'tgt_capacity': np.asarray(g['init_tgt_capacity'], dtype=np.int32).reshape(1, nt),
        'tgt_goal': g['step_tgt_goal'][k].astype(np.int32).reshape(1, nt),
        'tgt_weight': g['step_tgt_goal_weight'][k].astype(np.int32).reshape(1, nt),
        'tgt_bounty': g['step_bounties'][k].astype(np.int32).reshape(1, nt),
        'tgt_empty_bits': empty_bits_to_int(g['step_tgt_empty_bits'][k]).reshape(1, nt),
        'remaining': g['step_remaining'][k].astype(np.int32).reshape(1, 4, 4),
        'awaiting': g['step_awaiting'][k].astype(np.int32).reshape(1, 4),
        'num_delivered': np.asarray(g['step_num_delivered'][k], dtype=np.int32).reshape(1),
        'episode_step': np.asarray(g['step_episode_step'][k], dtype=np.int32).reshape(1),
        'episode_id': np.zeros(1, dtype=np.int32),
        'episode_reward': np.array(
            [[float(g['step_ep_reward'][k]), float(g['step_delayed_ep_reward'][k])]], dtype=np.float64
        ),
    }
