"""Small invocations of every kernel path that changed in the second half of round 2, for compute-sanitizer:
skewed / 8-byte copy-out (MATE-4v2-9), grouped obstacle rounds (4v2-9, 8v8-9, Navigation), rolled camera loop (8v8-9),
warp tiles of 8 / 16 / 32, folded wrappers, soft coverage, auxiliary terms, both greedy teams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mate_b200

for config, B, tile in (('MATE-4v2-9.yaml', 301, '8'), ('MATE-8v8-9.yaml', 130, '16'), ('MATE-Navigation.yaml', 97, '32'), ('MATE-4v8-9.yaml', 203, '32')):
    os.environ['MATE_B200_TILE'] = tile
    env = mate_b200.make('MultiAgentTracking-v0', config=config, num_envs=B, max_episode_steps=5, wrappers=[
        mate_b200.SharedFieldOfView, mate_b200.RelativeCoordinates, mate_b200.RescaledObservation] if '4v8' in config else [])
    base = env.unwrapped
    env.reset(seed=1)
    nc, nt = base.num_cameras, base.num_targets
    for k in range(14):
        ca = (torch.rand((B, nc, 2), device='cuda') * 2 - 1) * 5
        ta = (torch.rand((B, nt, 2), device='cuda') * 2 - 1) * 20
        env.step((ca, ta))
    torch.cuda.synchronize()
    print(config, 'ok', base.episode_statistics()['episodes'])
    base.close()
os.environ.pop('MATE_B200_TILE')
B = 128
env = mate_b200.make('MATE-4v8-9-v0', num_envs=B, max_episode_steps=6, wrappers=[
    mate_b200.MoreTrainingInformation, mate_b200.RepeatedRewardIndividualDone,
    lambda e: mate_b200.AuxiliaryCameraRewards(e, coefficients={'raw_reward': 1.0, 'soft_coverage_score': 1.0}),
    lambda e: mate_b200.MultiCamera(e, target_agent=mate_b200.GreedyTargetAgent(seed=0))])
env.reset(seed=0)
for k in range(10):
    env.step((torch.rand((B, 4, 2), device='cuda') * 2 - 1) * 5)
env = mate_b200.make('MATE-4v8-9-v0', num_envs=B, max_episode_steps=6, wrappers=[
    mate_b200.RepeatedRewardIndividualDone, lambda e: mate_b200.MultiTarget(e, camera_agent=mate_b200.GreedyCameraAgent(seed=0))])
env.reset(seed=0)
for k in range(10):
    env.step((torch.rand((B, 8, 2), device='cuda') * 2 - 1) * 20)
torch.cuda.synchronize()
print('wrappers + agents ok')
