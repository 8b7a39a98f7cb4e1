"""Multi-process sharding contract on CPU (gloo, world_size 2).

Environments are independent and the RNG streams are keyed on the GLOBAL environment index,
so a batch split across ranks (``env_index_base = rank * B_local``) must reproduce the
unsharded batch bit for bit, and the only collective -- the all-reduce of the 16-float
episode-statistics vector -- must sum to the unsharded statistics.  The oracle stands in
for the device here (this is a test of the host-side contract, no GPU)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mate_b200.config import flatten_config, read_config

B, STEPS, SEED = 64, 45, 11


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _actions(cfg, k):
    rng = np.random.RandomState(1000 + k)
    cam = rng.uniform(-1, 1, (B, cfg['num_cameras'], 2)) * [cfg['camera_rotation_step'], cfg['camera_zooming_step']]
    tgt = rng.uniform(-1, 1, (B, cfg['num_targets'], 2)) * cfg['target_step_size']
    return cam.astype(np.float32), tgt.astype(np.float32)


def _rollout(cfg, num_envs, base, sl):
    from oracle.oracle import Oracle

    sim = Oracle(cfg, num_envs, env_index_base=base)
    cam0, tgt0 = sim.reset(seed=SEED)
    outs = [(cam0.copy(), tgt0.copy())]
    dones = []
    for k in range(STEPS):
        cam, tgt = _actions(cfg, k)
        (co, to), rew, done = sim.step(cam[sl], tgt[sl], seed=SEED, auto_reset=True)
        outs.append((co.copy(), to.copy(), rew.copy()))
        dones.append(done.copy())
    return outs, np.stack(dones), sim.episode_stats()


def _worker(rank, world, port, result_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg = flatten_config(read_config('MATE-4v8-9.yaml', max_episode_steps=20))
    local = B // world
    sl = slice(rank * local, (rank + 1) * local)
    outs, dones, stats = _rollout(cfg, local, rank * local, sl)
    t = torch.from_numpy(stats.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    np.savez(f'{result_path}.{rank}.npz', last_tgt=outs[-1][1], last_rew=outs[-1][2], first_cam=outs[0][0],
             dones=dones, reduced=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path):
    cfg = flatten_config(read_config('MATE-4v8-9.yaml', max_episode_steps=20))
    outs, dones, stats = _rollout(cfg, B, 0, slice(0, B))
    assert dones.sum() >= B   # the time limit fired: auto-reset was exercised on every rank
    port = _free_port()
    path = str(tmp_path / 'shard')
    mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
    parts = [np.load(f'{path}.{r}.npz') for r in range(2)]
    np.testing.assert_array_equal(np.concatenate([p['first_cam'] for p in parts]), outs[0][0])
    np.testing.assert_array_equal(np.concatenate([p['last_tgt'] for p in parts]), outs[-1][1])
    np.testing.assert_array_equal(np.concatenate([p['last_rew'] for p in parts]), outs[-1][2])
    np.testing.assert_array_equal(np.concatenate([p['dones'] for p in parts], axis=1), dones)
    for p in parts:   # the all-reduced statistics equal the unsharded ones on every rank
        np.testing.assert_allclose(p['reduced'], stats, rtol=1e-12)
