"""mate_b200_step_host against the device-resident step for the other BASELINE shapes, every device -> host leg."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim

os.environ['MATE_B200_REFILL'] = 'sync'
for config in ('MATE-Navigation.yaml', 'MATE-8v8-9.yaml', 'MATE-4v2-9.yaml', 'MATE-4v8-0.yaml'):
    for mode in ('0', '1', '2'):
        os.environ['MATE_B200_HOST_COMPACT'] = mode
        cfg = flatten_config(read_config(config, max_episode_steps=5))
        nc, nt = cfg['num_cameras'], cfg['num_targets']
        B = 1536
        a, b = BatchedSim(cfg, B, device=0), BatchedSim(cfg, B, device=0)
        a.reset(seed=3); b.reset(seed=3)
        out = (torch.zeros((B, max(nc, 1), a.dc)).pin_memory(), torch.zeros((B, nt, a.dt)).pin_memory(), torch.zeros((B, 2)).pin_memory(), torch.zeros(B, dtype=torch.uint8).pin_memory())
        rng = np.random.RandomState(0)
        legs = set()
        for k in range(8):
            ca = torch.from_numpy((rng.uniform(-1, 1, (B, max(nc, 1), 2)) * [5.0, 2.5]).astype(np.float32)).pin_memory()
            ta = torch.from_numpy((rng.uniform(-1, 1, (B, nt, 2)) * 20.0).astype(np.float32)).pin_memory()
            (cam, tgt), rew, done = a.step(ca.cuda(), ta.cuda(), auto_reset=True)
            b.step_host(ca, ta, out, auto_reset=True, rows_kept=(k != 3))
            torch.cuda.synchronize()
            legs.add(b.host_leg_info()[0])
            if nc: assert torch.equal(cam.cpu(), out[0]), (config, mode, k)
            assert torch.equal(tgt.cpu(), out[1]) and torch.equal(rew.cpu(), out[2]) and torch.equal(done.cpu(), out[3]), (config, mode, k)
        print(config, 'mode', mode, 'ok, legs used', sorted(legs))
        a.close(); b.close()
