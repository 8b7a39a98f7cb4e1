"""Per-tile phase timeline of one step launch (development build with -DMATE_DEV_TIMELINE).
usage: MATE_B200_LIB=scratch/variants/libmate_tl.so [MATE_B200_KERNEL=2] python profiles/tools/timeline.py [out.npy]"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from mate_b200 import _abi

CONFIG = os.environ.get('TL_CONFIG', 'MATE-4v8-9.yaml')
cfg = flatten_config(read_config(CONFIG))
B = int(os.environ.get('TL_ENVS', '65536'))
NC, NT = cfg['num_cameras'], cfg['num_targets']
sim = BatchedSim(cfg, B, device=0)
sim.reset(seed=0)
steps0 = np.random.RandomState(1234).randint(0, cfg['max_episode_steps'] + 1, size=B).astype(np.int32)
sim.set_state({'episode_step': steps0})
g = torch.Generator(device='cuda'); g.manual_seed(0)
cams = [(torch.rand((B, max(NC, 1), 2), device='cuda', generator=g) * 2 - 1) * torch.tensor([cfg['camera_rotation_step'], cfg['camera_zooming_step']], device='cuda') for _ in range(4)]
tgts = [(torch.rand((B, NT, 2), device='cuda', generator=g) * 2 - 1) * cfg['target_step_size'] for _ in range(4)]
sim.alloc_aux()
for name, ctype, _, _ in _abi.AUX_FIELDS:
    if name not in ('coverage', 'num_delivered'):
        setattr(sim._aux_struct, name, ctype())
lib = ctypes.CDLL(os.environ['MATE_B200_LIB'])
lib.mate_b200_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int32]
lib.mate_b200_debug_timeline.restype = ctypes.c_int
all_t = []
for rep in range(3):
    for k in range(50):
        sim.step(cams[k % 4], tgts[k % 4], auto_reset=True, aux=True)
    torch.cuda.synchronize()
    out = np.zeros(4096 * 8, dtype=np.uint64)
    assert lib.mate_b200_debug_timeline(out.ctypes.data_as(ctypes.c_void_p), out.size) == 0
    t = out.reshape(4096, 8)[:(B + 31) // 32, :7].astype(np.int64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1000.0   # us
    all_t.append(t)
    names = ['start', 'tgt fast done', 'slow tgts done', 'dense view done', 'queues done', 'pack start', 'pack end']
    print('--- launch %d: kernel span %.1f us' % (rep, t[:, 6].max()))
    for k, n in enumerate(names):
        c = t[:, k]
        print('  %-16s min %6.1f  p10 %6.1f  median %6.1f  p90 %6.1f  p99 %6.1f  max %6.1f' % (n, c.min(), np.percentile(c, 10), np.median(c), np.percentile(c, 90), np.percentile(c, 99), c.max()))
    d = np.diff(t, axis=1)
    for k in range(6):
        c = d[:, k]
        print('  %-28s median %6.1f  p90 %6.1f  max %6.1f' % (names[k] + ' -> ' + names[k + 1].split()[0], np.median(c), np.percentile(c, 90), c.max()))
if len(sys.argv) > 1:
    np.save(sys.argv[1], np.stack(all_t))
