"""ms/step in consecutive windows of 20 steps from a fresh (staggered) state: is the start of a run heavier than its steady state?"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from mate_b200 import _abi
cfg = flatten_config(read_config('MATE-4v8-9.yaml'))
B = 65536
sim = BatchedSim(cfg, B, device=0)
sim.reset(seed=0)
steps0 = np.random.RandomState(1234).randint(0, cfg['max_episode_steps'] + 1, size=B).astype(np.int32)
sim.set_state({'episode_step': steps0})
g = torch.Generator(device='cuda'); g.manual_seed(0)
cams = [(torch.rand((B, 4, 2), device='cuda', generator=g) * 2 - 1) * torch.tensor([cfg['camera_rotation_step'], cfg['camera_zooming_step']], device='cuda') for _ in range(8)]
tgts = [(torch.rand((B, 8, 2), device='cuda', generator=g) * 2 - 1) * cfg['target_step_size'] for _ in range(8)]
sim.alloc_aux()
for name, ctype, _, _ in _abi.AUX_FIELDS:
    if name not in ('coverage', 'num_delivered'):
        setattr(sim._aux_struct, name, ctype())
torch.cuda.synchronize()
W = 20
res = []
k = 0
for w in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(W):
        (cam, tgt), rew, done = sim.step(cams[k % 8], tgts[k % 8], auto_reset=True, aux=True); k += 1
    b.record(); torch.cuda.synchronize()
    cov = float(sim._aux['coverage'][:, 0].mean()) if hasattr(sim, '_aux') and sim._aux else -1
    nz = float((tgt != 0).float().mean())
    res.append((k, a.elapsed_time(b) / W, cov, nz))
for r in res:
    print('steps..%5d  %.5f ms/step  coverage %.4f  nonzero tgt obs %.4f' % r)
