import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from mate_b200.config import flatten_config, read_config
from mate_b200.sim import BatchedSim
from oracle.oracle import Oracle
cfg = flatten_config(read_config('MATE-2v4-9.yaml', max_episode_steps=51))
B=130; seed=1234
sim=BatchedSim(cfg,B); ref=Oracle(cfg,B,num_threads=8)
sim.reset(seed=seed); ref.reset(seed=seed)
rng=np.random.RandomState(99)
nc,nt=2,4
for k in range(53):
    cam_act=(rng.uniform(-1,1,(B,nc,2))*[5,2.5]).astype(np.float32); tgt_act=(rng.uniform(-1,1,(B,nt,2))*20).astype(np.float32)
    (cam,tgt),rew,done=sim.step(torch.from_numpy(cam_act).cuda(),torch.from_numpy(tgt_act).cuda(),auto_reset=True)
    (rcam,rtgt),rrew,rdone=ref.step(cam_act,tgt_act,seed=seed,auto_reset=True)
    c=cam.cpu().numpy(); bad=np.argwhere(~np.isclose(c,rcam,rtol=1e-5,atol=1e-5))
    t=tgt.cpu().numpy(); badt=np.argwhere(~np.isclose(t,rtgt,rtol=1e-5,atol=1e-5))
    if len(bad) or len(badt):
        print('step',k,'cam bad',len(bad),'tgt bad',len(badt), 'done', int(rdone.sum()))
        for b in bad[:25]: print('  cam',b, c[tuple(b)], rcam[tuple(b)])
        for b in badt[:10]: print('  tgt',b, t[tuple(b)], rtgt[tuple(b)])
        s1,s2=sim.get_state(),ref.get_state()
        for key in s2:
            d=np.abs(s1[key].astype(float)-s2[key].astype(float)).max()
            if d>1e-9: print('  state diff',key,d)
        break
