import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, mate_b200
B = 128
env = mate_b200.make('MultiAgentTracking-v0', config='MATE-8v8-9.yaml', num_envs=B, wrappers=[
    lambda e: mate_b200.MoreTrainingInformation(e, full_observability=True), mate_b200.SharedFieldOfView, mate_b200.RescaledObservation,
    mate_b200.DiscreteCamera, mate_b200.RepeatedRewardIndividualDone,
    lambda e: mate_b200.AuxiliaryCameraRewards(e, coefficients={'soft_coverage_score': 1.0, 'coverage_rate': 1.0}, reduction='sum'),
    lambda e: mate_b200.MultiCamera(e, target_agent=mate_b200.GreedyTargetAgent(seed=3))])
print(env)
obs = env.reset(seed=1)
for k in range(30):
    obs, reward, done, infos = env.step(torch.randint(0, 25, (B, 8), device="cuda"))
print(obs.shape, reward.shape, done.shape, sorted(infos)[:8], float(reward.mean()))
assert infos['remaining_cargoes'].shape == (B, 4, 4) and infos['state'].shape[0] == B
env.unwrapped.close()
env = mate_b200.make('MATE-Navigation-v0', num_envs=B, wrappers=[mate_b200.MoreTrainingInformation, mate_b200.RepeatedRewardIndividualDone,
    lambda e: mate_b200.AuxiliaryTargetRewards(e, coefficients={'normalized_goal_distance': -1.0, 'is_colliding': -0.1})])
env.reset(seed=2)
for k in range(30):
    obs, reward, done, infos = env.step((torch.zeros((B, 0, 2), device='cuda'), (torch.rand((B, 8, 2), device='cuda') - 0.5) * 40))
print(obs[1].shape, reward[1].shape, float(reward[1].mean()), float(infos[1]['is_colliding'].float().mean()))
try:
    mate_b200.AuxiliaryTargetRewards(mate_b200.RepeatedRewardIndividualDone(env.unwrapped), coefficients={'soft_coverage_score': 1.0})
    print('ERROR: expected an assertion')
except AssertionError as ex:
    print('ok:', ex)
env.unwrapped.close()
# MultiTarget on an 8v8 config with wrappers below
env = mate_b200.make('MATE-8v8-9-v0', num_envs=B, wrappers=[mate_b200.EnhancedObservation, mate_b200.RelativeCoordinates,
    lambda e: mate_b200.MultiTarget(e, camera_agent=mate_b200.GreedyCameraAgent(seed=4))])
obs = env.reset(seed=3)
cov = 0
for k in range(100):
    obs, reward, done, infos = env.step((torch.rand((B, 8, 2), device='cuda') - 0.5) * 40)
    cov += float(infos['coverage_rate'].mean())
print(obs.shape, reward.shape, cov / 100)
env.unwrapped.close()
print('extras ok')
