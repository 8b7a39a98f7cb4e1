"""Experiment: the batch as G independent groups on G CUDA streams (no join between groups): does the packer of one
group overlap the simulation phase of the other?  usage: python scratch/pipeline_test.py [G] [offset_us]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from mate_b200 import _abi
from bench import make_actions

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
offset_us = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
B = 65536 // G
cfg = flatten_config(read_config('MATE-4v8-9.yaml'))
dev = torch.device('cuda', 0)
sims, streams, acts = [], [], []
for g in range(G):
    sim = BatchedSim(cfg, B, device=0, env_index_base=g * B)
    sim.reset(seed=0)
    steps0 = np.random.RandomState(1234 + g).randint(0, cfg['max_episode_steps'] + 1, size=B).astype(np.int32)
    sim.set_state({'episode_step': steps0})
    sim.alloc_aux()
    for name, ctype, _, _ in _abi.AUX_FIELDS:
        if name not in ('coverage', 'num_delivered'):
            setattr(sim._aux_struct, name, ctype())
    sims.append(sim); streams.append(torch.cuda.Stream(dev)); acts.append(make_actions(cfg, B, dev, 8, seed=g))
torch.cuda.synchronize()

def run(steps):
    for k in range(steps):
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                if k == 0 and g > 0 and offset_us > 0:
                    torch.cuda._sleep(int(offset_us * g * 1965))   # cycles
                sims[g].step(acts[g][0][k % 8], acts[g][1][k % 8], auto_reset=True, aux=True)

run(20); torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(G)]
K = 1000
for g in range(G):
    evs[g][0].record(streams[g])
run(K)
for g in range(G):
    evs[g][1].record(streams[g])
torch.cuda.synchronize()
ms = max(e[0].elapsed_time(e[1]) for e in evs) / K
print(f'groups={G} offset_us={offset_us} ms_per_full_batch_step={ms:.5f} env_steps_per_s={65536 / ms * 1e3:.4g} frac={6857 * 65536 / (ms * 1e-3) / 1e9 / 6553:.4f}')
