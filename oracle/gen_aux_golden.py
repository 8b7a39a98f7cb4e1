#!/usr/bin/env python
"""Golden fixtures for the auxiliary-reward / training-information wrappers -- TEST INFRASTRUCTURE ONLY.

Runs the unmodified reference (``/root/reference`` through ``oracle/gymshim``) with greedy agents under the
wrapper stack

    MoreTrainingInformation -> RepeatedRewardIndividualDone -> AuxiliaryCameraRewards -> AuxiliaryTargetRewards

(``mate/wrappers/more_training_information.py``, ``auxiliary_camera_rewards.py``, ``auxiliary_target_rewards.py``)
and records, for sampled steps, everything needed to repeat exactly that step elsewhere -- the simulator state
BEFORE the step, the joint action, the recorded stochastic draws -- together with what the reference's wrappers
put into the rewards and info dicts AFTER the step.

    python oracle/gen_aux_golden.py        # writes tests/golden/aux_*.npz
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  pylint: disable=wrong-import-position

REPO = os.path.dirname(HERE)

CAMERA_KEYS = ('raw_reward', 'coverage_rate', 'real_coverage_rate', 'mean_transport_rate', 'soft_coverage_score',
               'num_tracked', 'baseline')
TARGET_KEYS = ('raw_reward', 'coverage_rate', 'real_coverage_rate', 'mean_transport_rate', 'normalized_goal_distance',
               'sparse_delivery', 'soft_coverage_score', 'is_tracked', 'is_colliding', 'baseline')
# one distinct coefficient per key so that a swapped column shows up in the weighted sum
CAMERA_COEF = {k: 0.5 + 0.25 * i for i, k in enumerate(CAMERA_KEYS)}
TARGET_COEF = {k: -0.75 + 0.5 * i for i, k in enumerate(TARGET_KEYS)}

CONFIGS = [
    # name, config, seed, steps, stride
    ('aux_4v8-9', 'MATE-4v8-9.yaml', 31, 1500, 25),
    ('aux_8v8-9', 'MATE-8v8-9.yaml', 32, 500, 25),
    ('aux_Navigation', 'MATE-Navigation.yaml', 33, 1200, 40),
    ('aux_4v2-0', 'MATE-4v2-0.yaml', 34, 500, 25),
]


def run(mate, name, config, seed, num_steps, stride, out_dir):
    from mate.wrappers.single_team import group_reset, group_step  # pylint: disable=import-outside-toplevel

    base = mate.make('MultiAgentTracking-v0', config=config)
    u = base.unwrapped
    nc, nt = u.num_cameras, u.num_targets
    env = mate.MoreTrainingInformation(base)
    env = mate.RepeatedRewardIndividualDone(env)
    if nc:
        env = mate.AuxiliaryCameraRewards(env, coefficients=dict(CAMERA_COEF), reduction='none')
    # without cameras the reference's soft coverage score of a target is a max over an empty array: leave it out
    target_coef = {k: v for k, v in TARGET_COEF.items() if nc or k != 'soft_coverage_score'}
    env = mate.AuxiliaryTargetRewards(env, coefficients=target_coef, reduction='none')
    env.seed(seed)
    inst = gg.Instrument(base)

    rows = {}

    def push(key, value):
        rows.setdefault(key, []).append(np.asarray(value))

    cam_obs, tgt_obs = env.reset()
    inst.pop_dense()
    camera_agents = mate.GreedyCameraAgent(seed=seed).spawn(nc) if nc else []
    target_agents = mate.GreedyTargetAgent(seed=seed).spawn(nt)
    group_reset(camera_agents, cam_obs)
    group_reset(target_agents, tgt_obs)
    cam_infos = tgt_infos = None
    done, step = False, 0
    while not done and step < num_steps:
        cam_act = np.asarray(group_step(env, camera_agents, cam_obs, cam_infos), dtype=np.float64) if nc else np.zeros((0, 2))
        tgt_act = np.asarray(group_step(env, target_agents, tgt_obs, tgt_infos), dtype=np.float64)
        cam_act, tgt_act = gg.f32(cam_act).reshape(nc, 2), gg.f32(tgt_act).reshape(nt, 2)
        before = gg.dump_state(u)
        goals_before = u.target_goals.copy()
        (cam_obs, tgt_obs), (cam_rew, tgt_rew), (cam_done, tgt_done), (cam_infos, tgt_infos) = env.step((cam_act, tgt_act))
        done = all(cam_done) and all(tgt_done) if nc else all(tgt_done)
        transmit, reached = inst.pop_dense()
        step += 1
        picked = np.logical_and(u.target_goals != goals_before, u.target_goals >= 0)
        goal_choice = np.where(picked, u.target_goals, -1).astype(np.int8)
        interesting = bool(np.any(u.target_dones)) or bool(any(t.is_colliding for t in u.targets))
        if not (step % stride == 0 or interesting or done):
            continue
        for k, v in before.items():
            push(k, v)
        push('cam_act', cam_act)
        push('tgt_act', tgt_act)
        push('transmit', transmit)
        push('reached', reached)
        push('goal_choice', goal_choice)
        # after the step: what the reference's wrappers report
        push('out_cam_reward', np.asarray(cam_rew, dtype=np.float64).reshape(nc))
        push('out_tgt_reward', np.asarray(tgt_rew, dtype=np.float64).reshape(nt))
        push('out_cam_terms', np.array([[float(info[f'auxiliary_reward_{k}']) for k in CAMERA_KEYS] for info in cam_infos],
                                       dtype=np.float64).reshape(nc, len(CAMERA_KEYS)))
        push('out_tgt_terms', np.array([[float(info.get(f'auxiliary_reward_{k}', 0.0)) for k in TARGET_KEYS] for info in tgt_infos],
                                       dtype=np.float64).reshape(nt, len(TARGET_KEYS)))
        push('out_cam_num_tracked', np.array([info['num_tracked'] for info in cam_infos], dtype=np.int64).reshape(nc))
        push('out_cam_is_sensed', np.array([info['is_sensed'] for info in cam_infos], dtype=np.uint8).reshape(nc))
        push('out_tgt_goal', np.array([info['goal'] for info in tgt_infos], dtype=np.int64))
        push('out_tgt_goal_distance', np.array([info['goal_distance'] for info in tgt_infos], dtype=np.float64))
        push('out_tgt_warehouse_distances', np.array([info['warehouse_distances'] for info in tgt_infos], dtype=np.float64))
        push('out_tgt_individual_done', np.array([info['individual_done'] for info in tgt_infos], dtype=np.uint8))
        push('out_tgt_is_tracked', np.array([info['is_tracked'] for info in tgt_infos], dtype=np.uint8))
        push('out_tgt_is_colliding', np.array([info['is_colliding'] for info in tgt_infos], dtype=np.uint8))
        push('out_remaining_cargo_counts', np.asarray(tgt_infos[0]['remaining_cargo_counts'], dtype=np.int64))
        for k, v in gg.dump_masks(u).items():
            push('out_' + k, v)
        after = gg.dump_state(u)
        push('out_tgt_xy', after['tgt_xy'])
        push('out_tgt_empty_bits', after['tgt_empty_bits'])
        push('out_team_reward', np.array([cam_infos[0]['raw_reward'] if nc else -tgt_infos[0]['raw_reward'], tgt_infos[0]['raw_reward']], dtype=np.float64))
        push('out_episode_step', np.int64(u.episode_step))
        if nc:
            push('out_soft_matrix', np.asarray(mate.AuxiliaryCameraRewards.compute_soft_coverage_scores(u), dtype=np.float64))

    out = {'config_name': np.array(config), 'seed': np.int64(seed), 'count': np.int64(len(rows['cam_act'])),
           'camera_keys': np.array(CAMERA_KEYS), 'target_keys': np.array(TARGET_KEYS),
           'camera_coef': np.array([CAMERA_COEF[k] for k in CAMERA_KEYS]), 'target_coef': np.array([TARGET_COEF[k] for k in TARGET_KEYS])}
    out.update(gg.config_scalars(u))
    for k, v in rows.items():
        out['a_' + k] = np.stack(v)
    path = os.path.join(out_dir, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: samples={int(out["count"])} steps={step} delivered={u.num_delivered_cargoes} '
          f'deliveries_sampled={int(out["a_out_tgt_individual_done"].sum())} size={os.path.getsize(path) / 1e6:.2f}MB')


def main():
    mate = gg._import_reference()  # pylint: disable=protected-access
    out_dir = os.path.join(REPO, 'tests', 'golden')
    for name, config, seed, steps, stride in CONFIGS:
        run(mate, name, config, seed, steps, stride, out_dir)


if __name__ == '__main__':
    main()
