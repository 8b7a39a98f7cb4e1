#!/usr/bin/env python
"""Golden fixtures for the observation / action wrappers -- TEST INFRASTRUCTURE ONLY.

Runs the unmodified reference (``/root/reference`` through ``oracle/gymshim``) with greedy agents,
and at sampled steps records the full simulator state, the recorded transmittance draws, the raw joint
observation AND the output of the reference's wrappers applied to it (``mate/wrappers/*.py``), for the
wrapper stacks listed in ``STACKS``.  Also records the discrete -> continuous action tables of
``DiscreteCamera`` / ``DiscreteTarget``.

    python oracle/gen_wrapper_golden.py        # writes tests/golden/wrappers_*.npz
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402  pylint: disable=wrong-import-position

REPO = os.path.dirname(HERE)

# name -> list of (wrapper class name, kwargs), innermost first
STACKS = {
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

CONFIGS = [
    # name, config, seed, steps, stride
    ('wrappers_4v8-9', 'MATE-4v8-9.yaml', 21, 1400, 56),
    ('wrappers_8v8-9', 'MATE-8v8-9.yaml', 22, 600, 60),
    ('wrappers_Navigation', 'MATE-Navigation.yaml', 23, 1200, 80),
    ('wrappers_4v2-0', 'MATE-4v2-0.yaml', 24, 600, 60),
]


def apply_stack(wrapped, base, observation):
    """observation() of every ObservationWrapper between `base` and `wrapped`, innermost first."""
    chain = []
    env = wrapped
    while env is not base:
        chain.append(env)
        env = env.env
    cam, tgt = observation
    obs = (np.array(cam, dtype=np.float64, copy=True), np.array(tgt, dtype=np.float64, copy=True))
    for wrapper in reversed(chain):
        obs = wrapper.observation(obs)
        obs = (np.array(obs[0], dtype=np.float64, copy=True), np.array(obs[1], dtype=np.float64, copy=True))
    return obs


def run(mate, name, config, seed, num_steps, stride, out_dir):
    from mate.wrappers.single_team import group_reset, group_step  # pylint: disable=import-outside-toplevel

    env = mate.make('MultiAgentTracking-v0', config=config)
    u = env.unwrapped
    env.seed(seed)
    inst = gg.Instrument(env)
    nc, nt = u.num_cameras, u.num_targets
    stacks = {}
    for sname, spec in STACKS.items():
        wrapped = env
        for cls, kwargs in spec:
            wrapped = getattr(mate, cls)(wrapped, **kwargs)
        stacks[sname] = wrapped

    rows = {}

    def push(key, value):
        rows.setdefault(key, []).append(np.asarray(value))

    def sample(cam_obs, tgt_obs, transmit, reached):
        for k, v in gg.dump_state(u).items():
            push(k, v)
        push('transmit', transmit)
        push('reached', reached)
        push('cam_obs', cam_obs)
        push('tgt_obs', tgt_obs)
        for sname, wrapped in stacks.items():
            cam, tgt = apply_stack(wrapped, env, (cam_obs, tgt_obs))
            push(sname + '_cam_obs', cam)
            push(sname + '_tgt_obs', tgt)

    cam_obs, tgt_obs = env.reset()
    transmit, reached = inst.pop_dense()
    sample(cam_obs, tgt_obs, transmit, reached)
    camera_agents = mate.GreedyCameraAgent(seed=seed).spawn(nc) if nc else []
    target_agents = mate.GreedyTargetAgent(seed=seed).spawn(nt)
    group_reset(camera_agents, cam_obs)
    group_reset(target_agents, tgt_obs)
    cam_infos = tgt_infos = None
    done, step = False, 0
    while not done and step < num_steps:
        cam_act = np.asarray(group_step(env, camera_agents, cam_obs, cam_infos), dtype=np.float64) if nc else np.zeros((0, 2))
        tgt_act = np.asarray(group_step(env, target_agents, tgt_obs, tgt_infos), dtype=np.float64)
        (cam_obs, tgt_obs), _, done, (cam_infos, tgt_infos) = env.step((gg.f32(cam_act).reshape(nc, 2), gg.f32(tgt_act).reshape(nt, 2)))
        transmit, reached = inst.pop_dense()
        step += 1
        if step % stride == 0 or done:
            sample(cam_obs, tgt_obs, transmit, reached)

    out = {'config_name': np.array(config), 'seed': np.int64(seed), 'count': np.int64(len(rows['cam_obs'])),
           'stack_names': np.array(sorted(STACKS))}
    out.update(gg.config_scalars(u))
    for k, v in rows.items():
        out['w_' + k] = np.stack(v)
    # discrete action tables (mate/wrappers/discrete_action_spaces.py): index -> continuous action
    for levels in (3, 5):
        if nc:
            dc = mate.DiscreteCamera(env, levels=levels)
            idx = np.arange(levels * levels)
            out[f'discrete_camera_{levels}'] = np.stack([dc.action((np.full(nc, i), np.zeros((nt, 2))))[0][0] for i in idx])
        dt = mate.DiscreteTarget(env, levels=levels)
        idx = np.arange(levels * levels)
        out[f'discrete_target_{levels}'] = np.stack([dt.action((np.zeros((nc, 2)), np.full(nt, i)))[1][0] for i in idx])
    path = os.path.join(out_dir, name + '.npz')
    np.savez_compressed(path, **out)
    print(f'{name}: samples={int(out["count"])} steps={step} delivered={u.num_delivered_cargoes} size={os.path.getsize(path) / 1e6:.2f}MB')


def main():
    mate = gg._import_reference()  # pylint: disable=protected-access
    out_dir = os.path.join(REPO, 'tests', 'golden')
    for name, config, seed, steps, stride in CONFIGS:
        run(mate, name, config, seed, steps, stride, out_dir)


if __name__ == '__main__':
    main()
