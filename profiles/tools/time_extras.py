"""Per-kernel cost of the N3 / N4 rows at the bench size (MATE-4v8-9 x 65536 envs): CUDA events around 100 calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mate_b200

B = 65536
env = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=B)
base = env.unwrapped
env.reset(seed=0)
sim = base.sim
ca = torch.zeros((B, 4, 2), device='cuda'); ta = (torch.rand((B, 8, 2), device='cuda') * 2 - 1) * 20
for _ in range(50):
    sim.step(ca, ta, aux=True)

def timed(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

print('step + all aux outputs      %.4f ms' % timed(lambda: sim.step(ca, ta, aux=True)))
print('aux terms (N3)              %.4f ms' % timed(lambda: sim.auxiliary_terms()))
print('soft coverage (N3)          %.4f ms' % timed(lambda: sim.soft_coverage()))
tgt_agent = mate_b200.GreedyTargetAgent(seed=0); tgt_agent.bind(sim); tgt_agent.act(reset_mask=True)
cam_agent = mate_b200.GreedyCameraAgent(seed=0); cam_agent.bind(sim)
tracked = base._aux['mask_ct']
cam_agent.act(tracked, reset_mask=True)
print('GreedyTargetAgent team (N4) %.4f ms' % timed(lambda: tgt_agent.act()))
print('GreedyCameraAgent team (N4) %.4f ms' % timed(lambda: cam_agent.act(tracked)))
q = torch.rand(B * 4, device='cuda', dtype=torch.float64) * 360 - 180
e = torch.arange(B * 4, device='cuda', dtype=torch.int32) // 4; c = torch.arange(B * 4, device='cuda', dtype=torch.int32) % 4
print('fov_range, %d queries     %.4f ms' % (B * 4, timed(lambda: sim.fov_range(e, c, q), 20)))
menv = mate_b200.make('MultiAgentTracking-v0', config='MATE-4v8-9.yaml', num_envs=B, wrappers=[
    mate_b200.RepeatedRewardIndividualDone, lambda e_: mate_b200.MultiCamera(e_, target_agent=mate_b200.GreedyTargetAgent(seed=1))])
menv.reset(seed=0)
print('MultiCamera.step end to end %.4f ms' % timed(lambda: menv.step(ca), 50))
