"""Small end-to-end workload for compute-sanitizer: all K1 tiers (heavy fixtures), K2, K3, K4, K5."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bg_b200
dev = torch.device("cuda:0")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
d = np.load(os.path.join(G, "adversarial.npz"))
sel = np.concatenate([np.argsort(-d["counts"])[:40], np.arange(0, 200)])
c, o, a = bg_b200.legal_moves(torch.as_tensor(d["boards"][sel]).to(dev), torch.as_tensor(d["players"][sel]).to(dev),
                              torch.as_tensor(d["dice"][sel]).to(dev))
assert c.cpu().tolist() == d["counts"][sel].tolist()
env = bg_b200.B200BackgammonVecEnv(num_envs=512, device=dev, seed=3, check_every=0)
env.reset()
for t in range(12):
    env.step(env.random_actions(5, t))
env.encode_resident(True, True)
net = bg_b200.ValueNet.random_init(dev)
acts, vals = bg_b200.greedy_actions(env, net)
s = bg_b200.TwoPlySearch(net, max_afterstates_per_chunk=256)
best, scores, offsets, A = s.search(env.boards52[:48].clone(), env.players[:48].clone(), env.dice[:48].clone())
env.check_status()
torch.cuda.synchronize()
print("sanitize workload ok", int(best.shape[0]), int(A.shape[0]), s.leaves_evaluated)
# N1 / N2: policy kernel, PPO rollout + GAE + ManualUpdate (bg_ppo_loss_grad), host step buffers
from bg_b200.ppo import PPOConfig, PPOTrainer
pnet = bg_b200.PolicyValueNet.random_init(dev, seed=0)
tr = PPOTrainer(env, pnet, PPOConfig(t_horizon=6, num_epochs=2), seed=1)
tr.train(2, log=None)
host = bg_b200.HostStepBuffers(env)
h_acts = torch.zeros(512, dtype=torch.int32).pin_memory()
for t in range(4):
    env.step(h_acts, with_features=True, host=host); host.wait()
    h_acts.numpy()[:] = np.minimum(host.legal_counts.numpy() - 1, 3).clip(0)
env.check_status()
torch.cuda.synchronize()
print("sanitize workload 2 ok", tr.learner.last)
