"""Observation / action wrappers against the reference's own wrapper classes:
``tests/golden/wrappers_*.npz`` hold, for sampled states of greedy-agent episodes, the raw joint observation
and the output of the reference wrappers stacked on it (written by ``oracle/gen_wrapper_golden.py``)."""

import glob
import os

import numpy as np
import pytest

import golden_util as gu

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(gu.GOLDEN_DIR, 'wrappers_*.npz')))

# stack name -> wrapper factories, innermost first (the same stacks as oracle/gen_wrapper_golden.py)
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


@pytest.mark.parametrize('name', NAMES)
def test_rescale_tables_and_action_grids_match_the_reference(name):
    """CPU: the per-column (scale, shift) tables reproduce normalize_observation, the discrete action tables
    reproduce DiscreteCamera / DiscreteTarget."""
    from mate_b200 import wrappers
    from mate_b200.sim import rescale_tables

    g = gu.load(name)
    nc, nt, no = (int(x) for x in g['cfg_counts'])
    cam_t, tgt_t = rescale_tables(nc, nt, no)
    if nc:
        np.testing.assert_allclose(g['w_cam_obs'] * cam_t[:, 0] + cam_t[:, 1], g['w_rescaled_cam_obs'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(g['w_tgt_obs'] * tgt_t[:, 0] + tgt_t[:, 1], g['w_rescaled_tgt_obs'], rtol=1e-6, atol=1e-6)
    step = float(g['cfg_target'][0])
    for levels in (3, 5):
        np.testing.assert_allclose(wrappers.target_action_grid(levels) * step, g[f'discrete_target_{levels}'], rtol=0, atol=1e-12)
        if nc:
            high = np.array([g['cfg_camera'][3], g['cfg_camera'][4]])
            np.testing.assert_allclose(wrappers.camera_action_grid(levels) * high, g[f'discrete_camera_{levels}'], rtol=0, atol=1e-12)


def test_wrapper_order_rules():
    """The reference's ordering assertions (enhanced_observation.py:33-47, shared_field_of_view.py:35-49)."""
    from mate_b200 import wrappers

    class Fake:   # enough surface for the constructors that do not touch the simulator
        num_cameras = 4
        unwrapped = None

    fake = Fake()
    fake.unwrapped = fake
    inner = wrappers.Wrapper(fake)
    rel = wrappers.RelativeCoordinates.__new__(wrappers.RelativeCoordinates)
    wrappers.Wrapper.__init__(rel, inner)
    with pytest.raises(AssertionError):
        wrappers.EnhancedObservation(rel)
    with pytest.raises(AssertionError):
        wrappers.SharedFieldOfView(rel)
    with pytest.raises(AssertionError):
        wrappers.RelativeCoordinates(rel)
    with pytest.raises(AssertionError):
        wrappers.DiscreteCamera(inner, levels=4)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_observation_wrappers_match_the_reference(name):
    """GPU: sampled reference states are injected, re-observed with the recorded transmittance draws, and every
    wrapper stack is applied by the transform kernel; compared with the reference wrappers' output."""
    import torch

    import mate_b200

    g = gu.load(name)
    nc = int(g['cfg_counts'][0])
    count = int(g['count'])
    config = str(g['config_name'])
    states = gu.stack_states([gu.state_arrays(g, 'w_', i) for i in range(count)])
    for stack in [None] + sorted(STACKS):
        spec = [] if stack is None else STACKS[stack]
        if any(cls == 'EnhancedObservation' and kw['team'] == 'camera' for cls, kw in spec) and False:
            continue
        wrappers = [(lambda env, c=cls, k=kw: getattr(mate_b200, c)(env, **k)) for cls, kw in spec]
        env = mate_b200.make('MultiAgentTracking-v0', config=config, num_envs=count, wrappers=wrappers)
        base = env.unwrapped
        base.sim.set_state(states)
        cam, tgt = base.sim.observe(replay=(g['w_transmit'], None))
        torch.cuda.synchronize()
        prefix = 'w_' if stack is None else f'w_{stack}_'
        if nc:
            np.testing.assert_allclose(cam.cpu().numpy(), g[prefix + 'cam_obs'], rtol=1e-5, atol=2e-5, err_msg=f'{name} {stack} cameras')
        np.testing.assert_allclose(tgt.cpu().numpy(), g[prefix + 'tgt_obs'], rtol=1e-5, atol=2e-5, err_msg=f'{name} {stack} targets')
        base.close()


@pytest.mark.gpu
def test_discrete_actions_and_repeated_rewards():
    """DiscreteCamera / DiscreteTarget decode on the GPU to the same continuous actions as the tables;
    RepeatedRewardIndividualDone repeats the team reward per agent."""
    import torch

    import mate_b200

    B = 64
    ref = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=B)
    env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=B,
                         wrappers=[mate_b200.DiscreteCamera, mate_b200.DiscreteTarget, mate_b200.RepeatedRewardIndividualDone])
    ref.reset(seed=3)
    env.reset(seed=3)
    rng = np.random.RandomState(0)
    dc = mate_b200.wrappers.camera_action_grid(5) * np.array([ref.camera_rotation_step, ref.camera_zooming_step])
    dt = mate_b200.wrappers.target_action_grid(5) * ref.target_step_size
    for _ in range(5):
        ci, ti = rng.randint(0, 25, size=(B, 4)), rng.randint(0, 25, size=(B, 8))
        (cam_a, tgt_a), (cr, tr), (cd, td), _ = env.step((ci, ti))
        (cam_b, tgt_b), (cr_b, tr_b), done_b, _ = ref.step((torch.from_numpy(dc[ci].astype(np.float32)).cuda(),
                                                           torch.from_numpy(dt[ti].astype(np.float32)).cuda()))
        assert torch.equal(cam_a, cam_b) and torch.equal(tgt_a, tgt_b)
        assert cr.shape == (B, 4) and tr.shape == (B, 8) and cd.shape == (B, 4) and td.shape == (B, 8)
        assert torch.equal(cr[:, 0], cr_b) and torch.equal(tr[:, 3], tr_b) and torch.equal(td[:, 0], done_b)


@pytest.mark.gpu
@pytest.mark.parametrize('preset', ['MATE-1v1-0.yaml', 'MATE-2v2-9.yaml', 'MATE-1v2-9.yaml'])
def test_wrapper_stack_on_odd_shapes_runs_and_keeps_invariants(preset):
    """Shapes whose observation blocks are not 16-byte multiples (scalar load/store path of the kernel):
    shared + relative + rescaled keeps flags in {-1 .. 1} -> {0, 1} untouched and maps to [-1, 1]."""
    import torch

    import mate_b200

    B = 37
    raw = mate_b200.make('MultiAgentTracking-v0', config=preset, num_envs=B)
    env = mate_b200.make('MultiAgentTracking-v0', config=preset, num_envs=B,
                         wrappers=[mate_b200.SharedFieldOfView, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation])
    (cam_r, tgt_r), (cam_w, tgt_w) = raw.reset(seed=5), env.reset(seed=5)
    from mate_b200.sim import rescale_tables

    nc, nt, no = raw.num_cameras, raw.num_targets, raw.num_obstacles
    cam_t, tgt_t = rescale_tables(nc, nt, no)
    # the preserved counts / index columns are only shifted by their (zero) lower bound
    assert torch.equal(cam_w[..., :4], cam_r[..., :4]) and torch.equal(tgt_w[..., :4], tgt_r[..., :4])
    # own location is not a relative coordinate: it is just rescaled
    own = tgt_r[..., 13:15].cpu().numpy() * tgt_t[13:15, 0] + tgt_t[13:15, 1]
    np.testing.assert_allclose(tgt_w[..., 13:15].cpu().numpy(), own, rtol=1e-6, atol=1e-6)
    assert torch.isfinite(cam_w).all() and torch.isfinite(tgt_w).all()
    # teammates are always shared: every teammate flag is set (1 stays 1 under the flag columns' [-1, 1] bounds)
    assert bool((tgt_w[..., raw.sim.dt - 5 * nt + 4::5] == 1.0).all())


@pytest.mark.gpu
@pytest.mark.parametrize('fold', ['fast', 'generic'])
@pytest.mark.parametrize('preset', ['MATE-4v8-9.yaml', 'MATE-Navigation.yaml', 'MATE-2v2-9.yaml', 'MATE-8v8-9.yaml'])
def test_wrappers_folded_into_the_step_equal_the_separate_pass(preset, fold, monkeypatch):
    """The wrappers registered with the simulator (applied by the packer of the step kernel while the rows are in
    shared memory) give bit-identical rows to ``mate_b200_transform_observations`` run as a separate pass over the
    raw rows, through reset, steps with auto-reset and observe."""
    import torch

    import mate_b200
    from mate_b200 import _abi

    # 'fast': the wrappers are applied while the packer composes the rows (FoldOps); 'generic': on the staged rows
    monkeypatch.setenv('MATE_B200_FOLD', '1' if fold == 'fast' else '0')
    B = 1000
    ops = [_abi.OBS_SHARED_CAMERA, _abi.OBS_ENHANCED_TARGET, _abi.OBS_RELATIVE, _abi.OBS_RESCALED]
    folded = mate_b200.make('MultiAgentTracking-v0', config=preset, num_envs=B, max_episode_steps=6).unwrapped
    plain = mate_b200.make('MultiAgentTracking-v0', config=preset, num_envs=B, max_episode_steps=6).unwrapped
    if folded.num_cameras == 0:
        ops = ops[1:]
    folded.sim.set_observation_wrappers(ops)

    def separate_pass(cam, tgt):
        plain.sim.set_observation_wrappers(ops)          # registers ...
        plain.sim.transform_observations()               # ... and transforms the raw rows in place, as a pass of its own
        out = cam.clone(), tgt.clone()
        plain.sim.set_observation_wrappers([])
        return out

    a = folded.sim.reset(seed=3)
    b = separate_pass(*plain.sim.reset(seed=3))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1)
    nc, nt = folded.num_cameras, folded.num_targets
    for k in range(15):
        cam_act = (torch.rand((B, max(nc, 1), 2), device='cuda', generator=gen) * 2 - 1)[:, :nc] * 2.0
        tgt_act = (torch.rand((B, nt, 2), device='cuda', generator=gen) * 2 - 1) * folded.target_step_size
        (ca, ta), ra, da = folded.sim.step(cam_act, tgt_act, auto_reset=True)
        (cb, tb), rb, db = plain.sim.step(cam_act, tgt_act, auto_reset=True)
        cb, tb = separate_pass(cb, tb)
        assert torch.equal(ca, cb) and torch.equal(ta, tb), k
        assert torch.equal(ra, rb) and torch.equal(da, db), k
    a = folded.sim.observe()
    b = separate_pass(*plain.sim.observe())
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
