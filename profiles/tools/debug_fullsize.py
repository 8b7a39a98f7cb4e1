"""Find the first CUDA-vs-oracle mismatch of tests/test_full_size_properties.py::test_values_against_the_oracle_at_full_size and dump it."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ['MATE_B200_REFILL'] = os.environ.get('MATE_B200_REFILL', '8')
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from oracle.oracle import Oracle
preset, B = sys.argv[1], int(sys.argv[2])
steps, horizon = 64, 40
cfg = flatten_config(read_config(preset, max_episode_steps=horizon))
nc, nt = cfg['num_cameras'], cfg['num_targets']
sim = BatchedSim(cfg, B, device=0); ref = Oracle(cfg, B, num_threads=16)
seed = 77
sim.reset(seed=seed); ref.reset(seed=seed)
stagger = np.random.RandomState(3).randint(0, horizon + 1, size=B).astype(np.int32)
sim.set_state({'episode_step': stagger}); ref.set_state({'episode_step': stagger})
aux, raux = sim.alloc_aux(), ref.alloc_aux()
rng = np.random.RandomState(8)
wh = 925.0 * np.array([[1.0, 1.0], [-1.0, 1.0], [-1.0, -1.0], [1.0, -1.0]])
np.set_printoptions(precision=17, linewidth=200)
found = 0
for k in range(steps):
    cam_act = (rng.uniform(-1, 1, (B, nc, 2)) * [cfg['camera_rotation_step'], cfg['camera_zooming_step']]).astype(np.float32)
    tgt_act = (rng.uniform(-1, 1, (B, nt, 2)) * cfg['target_step_size']).astype(np.float32)
    if k % 2 == 0:
        st = ref.get_state(); goal = st['tgt_goal']
        direction = wh[np.where(goal >= 0, goal, 0)] - st['tgt_xy']
        direction /= np.maximum(np.linalg.norm(direction, axis=-1, keepdims=True), 1e-9)
        tgt_act = np.where((goal >= 0)[..., None], cfg['target_step_size'] * direction + 0.3 * tgt_act, tgt_act).astype(np.float32)
    pre_c, pre_r = sim.get_state(), ref.get_state()
    sim.step(torch.from_numpy(cam_act).cuda(), torch.from_numpy(tgt_act).cuda(), auto_reset=True, aux=True)
    ref.step(cam_act, tgt_act, seed=seed, auto_reset=True, aux=raux)
    for key in ('mask_ct', 'mask_cc', 'mask_co', 'mask_tc', 'mask_to', 'mask_tt', 'target_dones', 'is_colliding', 'num_delivered', 'episode_step'):
        a, b = aux[key].cpu().numpy(), raux[key]
        if not (a == b).all():
            idx = np.argwhere(a != b)
            print('step', k, key, 'mismatches', len(idx), idx[:5].tolist())
            e = int(idx[0][0])
            post_c, post_r = sim.get_state(), ref.get_state()
            print(' env', e, 'episode_step before', pre_r['episode_step'][e], 'cuda', a[e].tolist(), 'oracle', b[e].tolist())
            for kk in ('cam_xy', 'cam_phi', 'cam_theta', 'tgt_xy', 'obs_xyr', 'tgt_capacity'):
                if kk in pre_r:
                    print('  pre ', kk, 'equal' if (pre_c[kk][e] == pre_r[kk][e]).all() else 'DIFF', pre_r[kk][e].tolist())
            for kk in ('cam_phi', 'cam_theta', 'tgt_xy'):
                print('  post', kk, 'equal' if (post_c[kk][e] == post_r[kk][e]).all() else 'DIFF', 'cuda', post_c[kk][e].tolist(), 'oracle', post_r[kk][e].tolist())
            print('  tgt_act', tgt_act[e].tolist()); print('  cam_act', cam_act[e].tolist())
            found += 1
    if found >= 2: break
