import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mate_b200
for wr in ([], [mate_b200.RescaledObservation], [mate_b200.SharedFieldOfView, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation],
           [mate_b200.EnhancedObservation, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation]):
    env = mate_b200.make("MultiAgentTracking-v0", config="MATE-4v8-9.yaml", num_envs=65536, wrappers=wr)
    env.reset(seed=0)
    base = env.unwrapped
    ca = torch.zeros((65536,4,2), device="cuda"); ta = torch.zeros((65536,8,2), device="cuda")
    for _ in range(20): base.sim.step(ca, ta)
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): base.sim.step(ca, ta)
    b.record(); torch.cuda.synchronize()
    print([w.__name__ for w in wr], a.elapsed_time(b)/200, "ms/step")
    base.close()
