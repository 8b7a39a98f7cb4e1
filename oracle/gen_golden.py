#!/usr/bin/env python
"""Generate golden fixtures by RUNNING the unmodified reference -- TEST INFRASTRUCTURE ONLY.

Run in the build container (where ``/root/reference`` is mounted)::

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference (XuehaiPan/mate, pure Python) is imported unmodified from
``/root/reference`` through the stand-in ``gym`` package in ``oracle/gymshim``.  Nothing
here is imported by the product package; the fixtures it writes are what pins the C
oracle (``oracle/mate_oracle.c``) and, through it, the CUDA path.

What is recorded per trace (one ``.npz`` each):

* the full simulator state after ``reset`` (everything SURVEY.md section 8a lists as
  "state carried between steps") plus each camera's field-of-view table
  (``Camera.sight_range_func.x/.y``, reference ``mate/entities.py:457-478``);
* the float32-representable joint actions that were fed to ``env.step``;
* the two stochastic step-path draws, made dense: ``transmit[T, Nc, Nt]`` (outcome of
  ``Camera.perceive``'s ``binomial(1, transmittance)``, ``mate/entities.py:503``, for the
  pairs that reached it; ``reached[T, Nc, Nt]`` says which) and ``goal_choice[T, Nt]``
  (outcome of ``np_random.choice`` in ``_assign_goals``, ``mate/environment.py:1303``);
* per step: rewards, done, the five view masks, entity kinematic state, cargo tables,
  coverage statistics; joint observations at a stride.
"""

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get('MATE_REFERENCE', '/root/reference')


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, 'gymshim'))
    sys.path.insert(0, REFERENCE)
    import mate  # noqa: E402  pylint: disable=import-outside-toplevel

    return mate


class RecordingRNG:
    """Delegating proxy that logs binomial() outcomes (RandomState attributes are read-only)."""

    def __init__(self, rng, log):
        self._rng = rng
        self._log = log

    def __getattr__(self, name):
        return getattr(self._rng, name)

    def binomial(self, n, p, *args, **kwargs):
        out = self._rng.binomial(n, p, *args, **kwargs)
        self._log.append(int(out))
        return out


class Instrument:
    """Attach draw recorders to an UNMODIFIED reference env (instance-level only)."""

    def __init__(self, env):
        self.env = env.unwrapped
        self.records = []  # (camera_obj, other_obj, outcome)
        self.install()

    def install(self):
        u = self.env
        self.logs = {}
        for camera in u.cameras_ordered:
            log = []
            self.logs[id(camera)] = log
            space = camera.location_random_range
            rng = space.np_random
            if isinstance(rng, RecordingRNG):
                rng = rng._rng
            space._np_random = RecordingRNG(rng, log)
            if not getattr(camera, '_golden_wrapped', False):
                camera.perceive = self._wrap(camera, type(camera).perceive)
                camera._golden_wrapped = True

    def _wrap(self, camera, unbound):
        def perceive(other, transmittance=0.0):
            log = self.logs[id(camera)]
            before = len(log)
            result = unbound(camera, other, transmittance)
            if len(log) > before:
                self.records.append((camera, other, log[-1]))
            return result

        return perceive

    def pop_dense(self):
        """Dense [Nc, Nt] transmit outcome / reached arrays for the draws since last pop."""
        u = self.env
        nc, nt = u.num_cameras, u.num_targets
        transmit = np.zeros((nc, nt), dtype=np.uint8)
        reached = np.zeros((nc, nt), dtype=np.uint8)
        tindex = {id(t): i for i, t in enumerate(u.targets)}
        cindex = {id(c): i for i, c in enumerate(u.cameras)}
        for camera, other, outcome in self.records:
            if id(other) in tindex:
                c, t = cindex[id(camera)], tindex[id(other)]
                assert not reached[c, t]
                reached[c, t] = 1
                transmit[c, t] = outcome
            else:
                assert outcome == 0  # camera->camera draws use transmittance 0.0
        self.records.clear()
        return transmit, reached


def dump_state(u):
    """Everything the step path carries between steps."""
    nc, nt, no = u.num_cameras, u.num_targets, u.num_obstacles
    goal_weight = np.zeros(nt, dtype=np.int64)
    for t in range(nt):
        if u.target_goals[t] >= 0:
            goal_weight[t] = u.target_goal_bits[t, u.target_goals[t]]
        assert (u.targets[t].goal_bits == u.target_goal_bits[t]).all()
    return {
        'cam_xy': np.array([c.location for c in u.cameras], dtype=np.float64).reshape(nc, 2),
        'cam_phi': np.array([c.orientation for c in u.cameras], dtype=np.float64),
        'cam_theta': np.array([c.viewing_angle for c in u.cameras], dtype=np.float64),
        'cam_sight_range': np.array([c.sight_range for c in u.cameras], dtype=np.float64),
        'tgt_xy': np.array([t.location for t in u.targets], dtype=np.float64).reshape(nt, 2),
        'tgt_capacity': np.array(u.target_capacities, dtype=np.int64),
        'tgt_goal': np.array(u.target_goals, dtype=np.int64),
        'tgt_goal_weight': goal_weight,
        'tgt_empty_bits': np.array([t.empty_bits for t in u.targets], dtype=np.uint8).reshape(nt, 4),
        'tgt_colliding': np.array([t.is_colliding for t in u.targets], dtype=np.uint8),
        'tgt_orientation': np.array(u.target_orientations, dtype=np.float64),
        'freights': np.array(u.freights, dtype=np.int64),
        'bounties': np.array(u.bounties, dtype=np.int64),
        'target_steps': np.array(u.target_steps, dtype=np.int64),
        'tracked_steps': np.array(u.tracked_steps, dtype=np.int64),
        'obs_xyr': np.array([o.state() for o in u.obstacles], dtype=np.float64).reshape(no, 3),
        'remaining': np.array(u.remaining_cargoes, dtype=np.int64),
        'awaiting': np.array(u.awaiting_cargo_counts, dtype=np.int64),
        'num_delivered': np.int64(u.num_delivered_cargoes),
        'ep_reward': np.float64(u.target_team_episode_reward),
        'delayed_ep_reward': np.float64(u.delayed_target_team_episode_reward),
        'episode_step': np.int64(u.episode_step),
    }


def dump_masks(u):
    return {
        'mask_ct': np.array(u.camera_target_view_mask, dtype=np.uint8),
        'mask_cc': np.array(u.camera_camera_view_mask, dtype=np.uint8),
        'mask_co': np.array(u.camera_obstacle_view_mask, dtype=np.uint8),
        'mask_tc': np.array(u.target_camera_view_mask, dtype=np.uint8),
        'mask_to': np.array(u.target_obstacle_view_mask, dtype=np.uint8),
        'mask_tt': np.array(u.target_target_view_mask, dtype=np.uint8),
    }


def config_scalars(u):
    cfg = u.config
    cam = cfg.get('camera', {})
    nc = u.num_cameras

    def ranges(entities):
        out = []
        for e in entities:
            box = e.location_random_range
            out.append([box.low[0], box.high[0], box.low[1], box.high[1]])
        return np.array(out, dtype=np.float64).reshape(len(entities), 4)

    radius_range = np.zeros(2)
    if u.num_obstacles > 0:
        box = u.obstacles_ordered[0].radius_random_range
        radius_range = np.array([float(np.ravel(box.low)[0]), float(np.ravel(box.high)[0])])
    return {
        'cfg_counts': np.array([nc, u.num_targets, u.num_obstacles], dtype=np.int64),
        'cfg_max_episode_steps': np.int64(u.max_episode_steps),
        'cfg_num_cargoes_per_target': np.int64(u.num_cargoes_per_target),
        'cfg_num_high_capacity_targets': np.int64(u.num_high_capacity_targets),
        'cfg_targets_start_with_cargoes': np.int64(u.targets_start_with_cargoes),
        'cfg_shuffle_entities': np.int64(u.shuffle_entities),
        'cfg_reward_sparse': np.int64(cfg['reward_type'] == 'sparse'),
        'cfg_bounty_factor': np.float64(u.bounty_factor),
        'cfg_camera': np.array(
            [
                cam.get('radius', 40.0) if nc else 40.0,
                cam.get('min_viewing_angle', 90.0) if nc else 90.0,
                cam.get('max_sight_range', 500.0) if nc else 500.0,
                cam.get('rotation_step', 5.0) if nc else 5.0,
                cam.get('zooming_step', 2.5) if nc else 2.5,
            ],
            dtype=np.float64,
        ),
        'cfg_target': np.array([u.target_step_size, u.target_sight_range], dtype=np.float64),
        'cfg_transmittance': np.float64(u.obstacle_transmittance),
        'cfg_camera_ranges': ranges(u.cameras_ordered),
        'cfg_target_ranges': ranges(u.targets_ordered),
        'cfg_obstacle_ranges': ranges(u.obstacles_ordered),
        'cfg_obstacle_radius_range': radius_range,
        'cfg_scales': np.array(
            [u.freight_scale, u.bounty_scale, u.reward_scale, u.max_target_team_episode_reward],
            dtype=np.float64,
        ),
    }


def fov_tables(u):
    xs, ys, offs = [], [], [0]
    for c in u.cameras:
        xs.append(np.asarray(c.sight_range_func.x, dtype=np.float64))
        ys.append(np.asarray(c.sight_range_func.y, dtype=np.float64))
        offs.append(offs[-1] + len(xs[-1]))
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)  # noqa: E731
    return {'fov_phi': cat(xs), 'fov_rho': cat(ys), 'fov_off': np.array(offs, dtype=np.int64)}


def f32(a):
    """Round to float32-representable float64 (the CUDA ABI takes fp32 actions)."""
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def run_trace(mate, config, seed, num_steps, policy, obs_stride, out_path):
    from mate.wrappers.single_team import group_reset, group_step  # pylint: disable=import-outside-toplevel

    env = mate.make('MultiAgentTracking-v0', config=config)
    u = env.unwrapped
    env.seed(seed)
    inst = Instrument(env)
    cam_obs, tgt_obs = env.reset()
    transmit0, reached0 = inst.pop_dense()

    nc, nt = u.num_cameras, u.num_targets
    out = {'config_name': np.array(config), 'seed': np.int64(seed), 'policy': np.array(policy)}
    out.update(config_scalars(u))
    out.update({'init_' + k: v for k, v in dump_state(u).items()})
    out.update({'init_' + k: v for k, v in dump_masks(u).items()})
    out.update(fov_tables(u))
    out['init_cam_obs'] = cam_obs
    out['init_tgt_obs'] = tgt_obs
    out['init_transmit'] = transmit0
    out['init_reached'] = reached0

    rng = np.random.RandomState(seed + 12345)
    if policy == 'greedy':
        camera_agents = mate.GreedyCameraAgent(seed=seed).spawn(nc) if nc else []
        target_agents = mate.GreedyTargetAgent(seed=seed).spawn(nt)
        group_reset(camera_agents, cam_obs)
        group_reset(target_agents, tgt_obs)
    cam_infos = tgt_infos = None

    per_step = {}

    def push(key, value):
        per_step.setdefault(key, []).append(np.array(value))

    obs_steps = []
    step_count = 0
    done = False
    while not done and step_count < num_steps:
        if policy == 'random':
            cam_act = rng.uniform(-1.0, 1.0, size=(nc, 2)) * np.array(
                [u.camera_rotation_step, u.camera_zooming_step] if nc else [0.0, 0.0]
            )
            # a bit beyond the per-target limit so that the norm clamp is exercised
            tgt_act = rng.uniform(-1.0, 1.0, size=(nt, 2)) * u.target_step_size
        else:
            cam_act = (
                np.asarray(group_step(env, camera_agents, cam_obs, cam_infos), dtype=np.float64)
                if nc
                else np.zeros((0, 2))
            )
            tgt_act = np.asarray(group_step(env, target_agents, tgt_obs, tgt_infos), dtype=np.float64)
        cam_act = f32(cam_act).reshape(nc, 2)
        tgt_act = f32(tgt_act).reshape(nt, 2)

        goals_before = u.target_goals.copy()
        (cam_obs, tgt_obs), (cam_r, tgt_r), done, (cam_infos, tgt_infos) = env.step((cam_act, tgt_act))
        transmit, reached = inst.pop_dense()
        goals_after = u.target_goals
        picked = np.logical_and(goals_after != goals_before, goals_after >= 0)
        goal_choice = np.where(picked, goals_after, -1).astype(np.int8)

        push('cam_act', cam_act)
        push('tgt_act', tgt_act)
        push('transmit', transmit)
        push('reached', reached)
        push('goal_choice', goal_choice)
        push('reward', np.array([cam_r, tgt_r], dtype=np.float64))
        push('done', np.uint8(done))
        push('target_dones', np.array(u.target_dones, dtype=np.uint8))
        push('coverage', np.array([u.coverage_rate, u.real_coverage_rate, u.mean_transport_rate]))
        push('warehouse_dist', np.array(u.target_warehouse_distances, dtype=np.float64))
        state = dump_state(u)
        for key in (
            'cam_phi', 'cam_theta', 'cam_sight_range', 'tgt_xy', 'tgt_goal', 'tgt_goal_weight',
            'tgt_empty_bits', 'tgt_colliding', 'tgt_orientation', 'freights', 'bounties',
            'remaining', 'awaiting', 'num_delivered', 'ep_reward', 'delayed_ep_reward',
            'episode_step', 'target_steps', 'tracked_steps',
        ):
            push(key, state[key])
        for key, value in dump_masks(u).items():
            if key != 'mask_co':
                push(key, value)
        if step_count % obs_stride == 0 or done:
            obs_steps.append(step_count)
            push('cam_obs', cam_obs)
            push('tgt_obs', tgt_obs)
        step_count += 1

    for key, values in per_step.items():
        out['step_' + key] = np.stack(values)
    out['obs_steps'] = np.array(obs_steps, dtype=np.int64)
    out['num_steps'] = np.int64(step_count)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **out)
    total_reached = int(out['step_reached'].sum())
    print(
        f'{os.path.basename(out_path)}: steps={step_count} done={bool(done)} '
        f'return={u.target_team_episode_reward} delivered={u.num_delivered_cargoes} '
        f'transmit_draws={total_reached} pickups={int((out["step_goal_choice"] >= 0).sum())} '
        f'fov_sizes={np.diff(out["fov_off"]).tolist()} size={os.path.getsize(out_path) / 1e6:.2f}MB'
    )


def run_resets(mate, config, seed, count, out_path):
    """Post-reset states of `count` episodes: pins FOV-table build + reset-obs parity and
    provides reference samples for the distributional reset checks."""
    env = mate.make('MultiAgentTracking-v0', config=config)
    u = env.unwrapped
    env.seed(seed)
    inst = Instrument(env)
    rows = {}
    for _ in range(count):
        cam_obs, tgt_obs = env.reset()
        transmit, reached = inst.pop_dense()
        rec = dump_state(u)
        rec.update(dump_masks(u))
        rec.update(fov_tables(u))
        rec['cam_obs'] = cam_obs
        rec['tgt_obs'] = tgt_obs
        rec['transmit'] = transmit
        rec['reached'] = reached
        for k, v in rec.items():
            rows.setdefault(k, []).append(np.asarray(v))
    out = {'config_name': np.array(config), 'seed': np.int64(seed), 'count': np.int64(count)}
    out.update(config_scalars(u))
    for k, v in rows.items():
        if k in ('fov_phi', 'fov_rho'):
            out['reset_' + k] = np.concatenate(v)
            out['reset_' + k + '_len'] = np.array([len(x) for x in v], dtype=np.int64)
        else:
            out['reset_' + k] = np.stack(v)
    np.savez_compressed(out_path, **out)
    print(f'{os.path.basename(out_path)}: resets={count} size={os.path.getsize(out_path) / 1e6:.2f}MB')


def run_reset_samples(mate, config, seed, count, out_path):
    """Many post-reset states (state only): reference samples for the distributional checks of
    the Philox-driven reset (tests/test_reset_distribution.py)."""
    env = mate.make('MultiAgentTracking-v0', config=config)
    u = env.unwrapped
    env.seed(seed)
    keys = ('cam_xy', 'cam_phi', 'cam_theta', 'tgt_xy', 'tgt_capacity', 'tgt_goal', 'tgt_goal_weight',
            'bounties', 'obs_xyr', 'remaining', 'awaiting')
    rows = {k: [] for k in keys}
    for _ in range(count):
        env.reset()
        rec = dump_state(u)
        for k in keys:
            rows[k].append(np.asarray(rec[k]))
    out = {'config_name': np.array(config), 'seed': np.int64(seed), 'count': np.int64(count)}
    out.update(config_scalars(u))
    for k, v in rows.items():
        out['sample_' + k] = np.stack(v).astype(np.float32 if v[0].dtype.kind == 'f' else np.int16)
    np.savez_compressed(out_path, **out)
    print(f'{os.path.basename(out_path)}: reset samples={count} size={os.path.getsize(out_path) / 1e6:.2f}MB')


TRACES = [
    # name, config, seed, steps, policy, obs_stride
    ('4v2-9_random', 'MATE-4v2-9.yaml', 0, 10050, 'random', 64),
    ('4v2-9_greedy', 'MATE-4v2-9.yaml', 1, 4000, 'greedy', 16),
    ('4v8-9_random', 'MATE-4v8-9.yaml', 0, 600, 'random', 4),
    ('4v8-9_greedy', 'MATE-4v8-9.yaml', 0, 4000, 'greedy', 16),
    ('8v8-9_random', 'MATE-8v8-9.yaml', 2, 300, 'random', 4),
    ('8v8-9_greedy', 'MATE-8v8-9.yaml', 3, 4000, 'greedy', 32),
    ('4v8-0_random', 'MATE-4v8-0.yaml', 4, 400, 'random', 4),
    ('4v8-0_greedy', 'MATE-4v8-0.yaml', 5, 4000, 'greedy', 32),
    ('Navigation_random', 'MATE-Navigation.yaml', 6, 400, 'random', 4),
    ('Navigation_greedy', 'MATE-Navigation.yaml', 7, 4000, 'greedy', 32),
    ('2v4-9_greedy', 'MATE-2v4-9.yaml', 8, 3000, 'greedy', 32),
    ('1v1-9_random', 'MATE-1v1-9.yaml', 9, 300, 'random', 8),
]

RESETS = [
    ('4v8-9_resets', 'MATE-4v8-9.yaml', 100, 12),
    ('8v8-9_resets', 'MATE-8v8-9.yaml', 101, 6),
    ('Navigation_resets', 'MATE-Navigation.yaml', 102, 6),
    ('4v2-9_resets', 'MATE-4v2-9.yaml', 103, 6),
]


RESET_SAMPLES = [
    ('4v8-9_resetsamples', 'MATE-4v8-9.yaml', 200, 250),
    ('Navigation_resetsamples', 'MATE-Navigation.yaml', 201, 150),
]


def main():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument('--out', default=os.path.join(REPO, 'tests', 'golden'))
    parser.add_argument('--only', default=None, help='substring filter on trace names')
    args = parser.parse_args()
    mate = _import_reference()
    for name, config, seed, steps, policy, stride in TRACES:
        if args.only and args.only not in name:
            continue
        run_trace(mate, config, seed, steps, policy, stride, os.path.join(args.out, name + '.npz'))
    for name, config, seed, count in RESETS:
        if args.only and args.only not in name:
            continue
        run_resets(mate, config, seed, count, os.path.join(args.out, name + '.npz'))
    for name, config, seed, count in RESET_SAMPLES:
        if args.only and args.only not in name:
            continue
        run_reset_samples(mate, config, seed, count, os.path.join(args.out, name + '.npz'))


if __name__ == '__main__':
    main()
