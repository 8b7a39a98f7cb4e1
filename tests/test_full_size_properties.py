"""Size-independent properties at BASELINE.json's full sizes (65 536 envs on one B200)."""

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

B_FULL = 65536


def _make(preset, B, base=0, **kw):
    from mate_b200.config import flatten_config, read_config
    from mate_b200.sim import BatchedSim

    cfg = flatten_config(read_config(preset, **kw))
    return cfg, BatchedSim(cfg, B, device=0, env_index_base=base)


def _actions(cfg, B, k, device='cuda'):
    gen = torch.Generator(device=device)
    gen.manual_seed(100 + k)
    scale = torch.tensor([cfg['camera_rotation_step'], cfg['camera_zooming_step']], device=device)
    cam = (torch.rand((B, max(cfg['num_cameras'], 1), 2), device=device, generator=gen) * 2 - 1) * scale
    tgt = (torch.rand((B, cfg['num_targets'], 2), device=device, generator=gen) * 2 - 1) * cfg['target_step_size']
    return cam[:, :cfg['num_cameras']], tgt


@pytest.mark.parametrize('preset', ['MATE-4v8-9.yaml', 'MATE-4v8-0.yaml', 'MATE-Navigation.yaml'])
def test_determinism_sharding_and_conservation(preset):
    """(i) same seed + actions => identical results; (ii) two half batches with
    env_index_base offsets == the full batch (RNG keyed on the global env index);
    (iii) cargo is conserved and observations are self-consistent."""
    cfg, full = _make(preset, B_FULL, max_episode_steps=25)
    _, lo = _make(preset, B_FULL // 2, base=0, max_episode_steps=25)
    _, hi = _make(preset, B_FULL // 2, base=B_FULL // 2, max_episode_steps=25)
    nc, nt, no = cfg['num_cameras'], cfg['num_targets'], cfg['num_obstacles']
    half = B_FULL // 2
    for sim in (full, lo, hi):
        sim.reset(seed=5)
    total = cfg['num_cargoes_per_target'] * nt
    n_done = 0
    for k in range(30):
        cam, tgt = _actions(cfg, B_FULL, k)
        (co, to), rew, done = full.step(cam, tgt)
        (co_l, to_l), rew_l, done_l = lo.step(cam[:half], tgt[:half])
        (co_h, to_h), rew_h, done_h = hi.step(cam[half:], tgt[half:])
        assert torch.equal(to[:half], to_l) and torch.equal(to[half:], to_h)
        assert torch.equal(co[:half], co_l) and torch.equal(co[half:], co_h)
        assert torch.equal(rew[:half], rew_l) and torch.equal(rew[half:], rew_h)
        assert torch.equal(done[:half], done_l) and torch.equal(done[half:], done_h)
        n_done += int(done.sum())
        # rewards are integer valued, camera reward = -target reward (environment.py:621)
        assert torch.equal(rew, rew.round()) and torch.equal(rew[:, 0], -rew[:, 1])
        # preserved data of every row (environment.py:499-501)
        assert (to[:, :, 0] == nc).all() and (to[:, :, 1] == nt).all() and (to[:, :, 2] == no).all()
        assert torch.equal(to[:, :, 3], torch.arange(nt, device='cuda', dtype=torch.float32).expand(B_FULL, nt))
        assert (to[:, :, 12] == 75.0).all()
        # positions stay on the terrain
        assert (to[:, :, 13:15].abs() <= 1000.0).all()
        # masked blocks are all-zero, visible ones carry the flag: x == 0 and y == 0 wherever flag == 0
        tgt_blocks = to[:, :, 27 + 7 * nc + 4 * no:].reshape(B_FULL, nt, nt, 5)
        hidden = tgt_blocks[..., 4] == 0
        assert (tgt_blocks[hidden] == 0).all()
        assert (tgt_blocks[..., 4].diagonal(dim1=1, dim2=2) == 1).all()   # a target always sees itself
    assert n_done >= B_FULL   # time limit + auto-reset exercised at full size
    s = full.get_state()
    carried = np.where(s['tgt_goal'] >= 0, s['tgt_weight'], 0).sum(axis=1)
    conserved = s['remaining'].reshape(B_FULL, -1).sum(axis=1) + carried + s['num_delivered']
    assert (conserved % total == 0).all() and (conserved > 0).all()
    awaiting = s['remaining'].sum(axis=1)
    for w in range(4):
        awaiting[:, w] += np.where(s['tgt_goal'] == w, s['tgt_weight'], 0).sum(axis=1)
    assert (awaiting == s['awaiting']).all()


def test_time_limit_done_at_max_plus_one():
    """done fires at episode_step == max_episode_steps + 1 (environment.py:630-632)."""
    cfg, sim = _make('MATE-4v8-9.yaml', 4096, max_episode_steps=7)
    sim.reset(seed=1)
    for k in range(8):
        cam, tgt = _actions(cfg, 4096, k)
        _, _, done = sim.step(cam, tgt, auto_reset=False)
        assert bool(done.all()) == (k == 7), k


def test_observe_is_idempotent_and_set_get_state_roundtrip():
    cfg, sim = _make('MATE-8v8-9.yaml', 2048)
    sim.reset(seed=9)
    cam, tgt = _actions(cfg, 2048, 0)
    sim.step(cam, tgt)
    a = [t.clone() for t in sim.observe()]
    b = [t.clone() for t in sim.observe()]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    state = sim.get_state()
    _, other = _make('MATE-8v8-9.yaml', 2048)
    other.set_state(state)
    again = other.get_state()
    for key in state:
        np.testing.assert_array_equal(state[key], again[key], err_msg=key)
    c = other.observe()
    # preserved data + private state do not depend on the (seed-keyed) transmittance draws
    assert torch.equal(a[1][:, :, :27], c[1][:, :, :27])


def _rim_grazing(state, env, cfg):
    """Does a target of environment `env` sit on the rim of a disc (obstacle or camera) to within 1e-9?

    ``Obstacle.obstruct`` (mate/entities.py:158-184) bends a step that enters a disc so that it ends ON the rim when the
    step only grazes the disc with its last 1e-6 units; in the next step the test ``relative.norm < self.radius``
    (entities.py:161) -- which REVERSES the step -- is then decided by the last bit of ``sqrt(dx^2 + dy^2)``, in the
    reference as well as here.  The oracle reproduces NumPy's rounding of this container; the CUDA path's positions differ
    from the oracle's by one ulp now and then (fused multiply-adds), so about one such decision per
    10^7 target-steps comes out differently.  Like the tangent rays it is a documented class, checked here explicitly."""
    tgt = state['tgt_xy'][env]
    discs = [state['obs_xyr'][env]] if cfg['num_obstacles'] else []
    if cfg['num_cameras']:
        cam = state['cam_xy'][env]
        discs.append(np.concatenate([cam, np.full((len(cam), 1), cfg['camera_radius'])], axis=1))
    if not discs:
        return False
    discs = np.concatenate(discs)
    d = np.linalg.norm(tgt[:, None, :] - discs[None, :, :2], axis=-1) - discs[None, :, 2]
    return bool((np.abs(d) < 1e-9).any())


@pytest.mark.parametrize('preset,B', [('MATE-4v8-9.yaml', B_FULL), ('MATE-4v8-0.yaml', B_FULL), ('MATE-Navigation.yaml', B_FULL),
                                      ('MATE-8v8-9.yaml', B_FULL // 2)])
def test_values_against_the_oracle_at_full_size(preset, B, monkeypatch):
    """CUDA against the float64 C oracle at BASELINE.json's sizes, in the regime bench.py times: episode clocks
    staggered over one episode (time-limit resets in every step), prepared episodes refilled asynchronously on the side
    stream (a short period, so that auto-resets both adopt prepared episodes and run in place).  Masks, flags, rewards,
    done, cargo counts bit-exact; observations to 1e-5; positions to 1e-8.  The only tolerated difference is the
    rim-grazing class (see `_rim_grazing`): every environment that differs must belong to it, at most 4 of the 2-4
    million environment-steps of a run may, and the batch is re-synchronised from the oracle afterwards."""
    from mate_b200.config import flatten_config, read_config
    from mate_b200.sim import BatchedSim
    from oracle.oracle import Oracle

    steps, horizon = 64, 40
    cfg = flatten_config(read_config(preset, max_episode_steps=horizon))
    nc, nt = cfg['num_cameras'], cfg['num_targets']
    monkeypatch.setenv('MATE_B200_REFILL', '8')
    sim = BatchedSim(cfg, B, device=0)
    ref = Oracle(cfg, B, num_threads=16)
    seed = 77
    sim.reset(seed=seed)
    ref.reset(seed=seed)
    stagger = np.random.RandomState(3).randint(0, horizon + 1, size=B).astype(np.int32)
    sim.set_state({'episode_step': stagger})
    ref.set_state({'episode_step': stagger})
    aux, raux = sim.alloc_aux(), ref.alloc_aux()
    rng = np.random.RandomState(8)
    wh = 925.0 * np.array([[1.0, 1.0], [-1.0, 1.0], [-1.0, -1.0], [1.0, -1.0]])

    def differing(got, want, exact):   # environments in which a device tensor differs from the oracle's array
        want = torch.from_numpy(np.ascontiguousarray(want)).cuda()
        bad = (got != want) if exact else ((got - want).abs() > 1e-5 + 1e-5 * want.abs())
        return torch.nonzero(bad.reshape(B, -1).any(dim=1)).flatten()

    n_done = grazing_events = 0
    for k in range(steps):
        before = ref.get_state()
        cam_act = (rng.uniform(-1, 1, (B, nc, 2)) * [cfg['camera_rotation_step'], cfg['camera_zooming_step']]).astype(np.float32)
        tgt_act = (rng.uniform(-1, 1, (B, nt, 2)) * cfg['target_step_size']).astype(np.float32)
        if k % 2 == 0:   # every other step the targets head for their goal warehouse, so that cargo moves
            goal = before['tgt_goal']
            direction = wh[np.where(goal >= 0, goal, 0)] - before['tgt_xy']
            direction /= np.maximum(np.linalg.norm(direction, axis=-1, keepdims=True), 1e-9)
            tgt_act = np.where((goal >= 0)[..., None], cfg['target_step_size'] * direction + 0.3 * tgt_act, tgt_act).astype(np.float32)
        (cam, tgt), rew, done = sim.step(torch.from_numpy(cam_act).cuda(), torch.from_numpy(tgt_act).cuda(), auto_reset=True, aux=True)
        (rcam, rtgt), rrew, rdone = ref.step(cam_act, tgt_act, seed=seed, auto_reset=True, aux=raux)
        bad = [differing(aux[key], raux[key], True) for key in ('mask_ct', 'mask_cc', 'mask_co', 'mask_tc', 'mask_to', 'mask_tt',
                                                                'target_dones', 'is_colliding', 'num_delivered', 'episode_step')]
        bad += [differing(rew, rrew, True), differing(done, rdone, True), differing(tgt, rtgt, False)]
        if nc:
            bad.append(differing(cam, rcam, False))
        bad = torch.unique(torch.cat(bad)).cpu().tolist()
        if bad:
            assert len(bad) <= 2, (preset, k, bad[:10])
            for env in bad:
                assert _rim_grazing(before, env, cfg), (preset, k, env, 'differs from the oracle and is not a rim-grazing case')
            grazing_events += len(bad)
            sim.set_state(ref.get_state())   # re-synchronise (the diverged environment would differ from now on)
        n_done += int(rdone.sum())
    assert grazing_events <= 4, grazing_events
    assert n_done >= B   # every environment ended an episode at least once
    s_cuda, s_ref = sim.get_state(), ref.get_state()
    for key in s_ref:
        if s_ref[key].dtype.kind == 'f':
            np.testing.assert_allclose(s_cuda[key], s_ref[key], rtol=0, atol=1e-8, err_msg=key)
        else:
            assert (s_cuda[key] == s_ref[key]).all(), key
    stats = sim.episode_stats().cpu().numpy()
    assert stats[6] > 0 and stats[6] + stats[7] == stats[0], stats[:8]
    print(f'{preset}: {B * steps} environment-steps, {grazing_events} rim-grazing divergences, {int(stats[6])} adopted / {int(stats[7])} in-place resets')
    sim.close()
