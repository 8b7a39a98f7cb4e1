"""History of one target before a CUDA-vs-oracle divergence (see debug_fullsize.py)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ['MATE_B200_REFILL'] = '8'
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from oracle.oracle import Oracle
preset, B, E, T, O, last = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
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
def d_of(st):
    o = st['obs_xyr'][E, O]; p = st['tgt_xy'][E, T]
    dx, dy = o[0] - p[0], o[1] - p[1]
    return np.sqrt(dx * dx + dy * dy) - o[2]
for k in range(last + 1):
    cam_act = (rng.uniform(-1, 1, (B, nc, 2)) * [cfg['camera_rotation_step'], cfg['camera_zooming_step']]).astype(np.float32)
    tgt_act = (rng.uniform(-1, 1, (B, nt, 2)) * cfg['target_step_size']).astype(np.float32)
    if k % 2 == 0:
        st = ref.get_state(); goal = st['tgt_goal']
        direction = wh[np.where(goal >= 0, goal, 0)] - st['tgt_xy']
        direction /= np.maximum(np.linalg.norm(direction, axis=-1, keepdims=True), 1e-9)
        tgt_act = np.where((goal >= 0)[..., None], cfg['target_step_size'] * direction + 0.3 * tgt_act, tgt_act).astype(np.float32)
    if k >= last - 10:
        pc, pr = sim.get_state(), ref.get_state()
        print('step', k, 'ep_step', pr['episode_step'][E], 'pos cuda', pc['tgt_xy'][E, T].tolist(), 'oracle', pr['tgt_xy'][E, T].tolist(),
              'd-r cuda %.3e oracle %.3e' % (d_of(pc), d_of(pr)), 'act', tgt_act[E, T].tolist(), 'goal', pr['tgt_goal'][E, T], 'colliding', raux['is_colliding'][E, T])
    sim.step(torch.from_numpy(cam_act).cuda(), torch.from_numpy(tgt_act).cuda(), auto_reset=True, aux=True)
    ref.step(cam_act, tgt_act, seed=seed, auto_reset=True, aux=raux)
