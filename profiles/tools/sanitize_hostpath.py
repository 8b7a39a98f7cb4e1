"""Small invocations of what the end of round 2 added, for compute-sanitizer: mate_b200_step_host through both compacted
legs (compact_chunks_kernel, compact_changes_kernel, publish_counts_kernel; ragged batch sizes, auto-resets), the step
launch with the persisting L2 window forced on, the folded wrappers after their trim."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim

os.environ['MATE_B200_REFILL'] = 'sync'
for mode, B in (('1', 1028), ('2', 516), ('1', 4)):
    os.environ['MATE_B200_HOST_COMPACT'] = mode
    cfg = flatten_config(read_config('MATE-4v8-9.yaml', max_episode_steps=4))
    a, b = BatchedSim(cfg, B, device=0), BatchedSim(cfg, B, device=0)
    a.reset(seed=3); b.reset(seed=3)
    out = (torch.zeros((B, 4, a.dc)).pin_memory(), torch.zeros((B, 8, a.dt)).pin_memory(), torch.zeros((B, 2)).pin_memory(), torch.zeros(B, dtype=torch.uint8).pin_memory())
    rng = np.random.RandomState(0)
    for k in range(9):
        ca = torch.from_numpy((rng.uniform(-1, 1, (B, 4, 2)) * [5.0, 2.5]).astype(np.float32)).pin_memory()
        ta = torch.from_numpy((rng.uniform(-1, 1, (B, 8, 2)) * 20.0).astype(np.float32)).pin_memory()
        (cam, tgt), rew, done = a.step(ca.cuda(), ta.cuda(), auto_reset=True)
        b.step_host(ca, ta, out, auto_reset=True, rows_kept=(k != 4))
        torch.cuda.synchronize()
        assert torch.equal(cam.cpu(), out[0]) and torch.equal(tgt.cpu(), out[1]), (mode, B, k)
    print('step_host mode', mode, 'B', B, 'ok, last leg', b.host_leg_info())
    a.close(); b.close()
print('host path ok')
