import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from bench import make_actions
cfg = flatten_config(read_config('MATE-4v8-9.yaml'))
B = 16384
sim = BatchedSim(cfg, B, device=0); sim.reset(seed=0)
cams, tgts = make_actions(cfg, B, torch.device('cuda', 0), 8, seed=0)
for k in range(300):
    (cam, tgt), _, _ = sim.step(cams[k % 8], tgts[k % 8], auto_reset=True)
flat = torch.cat([cam.reshape(B, -1), tgt.reshape(B, -1)], dim=1)
print('floats nonzero', float((flat != 0).float().mean()))
for w in (4, 8, 16):
    n = flat.shape[1] // w * w
    nz = (flat[:, :n].reshape(B, -1, w) != 0).any(-1).float().mean()
    print(f'{w * 4}-byte chunks nonzero', float(nz))
