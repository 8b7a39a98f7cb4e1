"""One wrapped configuration of the step for an ncu capture: python wrapped_step_once.py [shared|enhanced|rescaled] [steps]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mate_b200
which = sys.argv[1] if len(sys.argv) > 1 else 'shared'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
wr = {'shared': [mate_b200.SharedFieldOfView, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation],
      'enhanced': [mate_b200.EnhancedObservation, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation],
      'relative': [mate_b200.RelativeCoordinates], 'sharedonly': [mate_b200.SharedFieldOfView],
      'rescaled': [mate_b200.RescaledObservation], 'none': []}[which]
env = mate_b200.make("MultiAgentTracking-v0", config="MATE-4v8-9.yaml", num_envs=65536, wrappers=wr)
env.reset(seed=0)
base = env.unwrapped
g = torch.Generator(device='cuda').manual_seed(0)
ca = (torch.rand((65536, 4, 2), device="cuda", generator=g) * 2 - 1) * torch.tensor([5.0, 2.5], device='cuda')
ta = (torch.rand((65536, 8, 2), device="cuda", generator=g) * 2 - 1) * 20.0
for _ in range(20): base.sim.step(ca, ta)
torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps): base.sim.step(ca, ta)
b.record(); torch.cuda.synchronize()
print(which, a.elapsed_time(b) / steps, "ms/step")
