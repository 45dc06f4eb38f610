"""PPO golden vectors from the UNMODIFIED reference learner (src/agent/ppo_agent.py) run in this container:

  * a short rollout through the reference's VectorizedBackgammonEnv + BackgammonPPOAgent.select_action
    (observations, prefix masks, sampled actions, their log-probs and values, rewards, dones; step-major memory
    exactly as train.py:51-66 fills it),
  * compute_returns on that memory (ppo_agent.py:206-216),
  * update() with the learning rate set to 0, so that the four epochs see the same weights and the means the agent
    keeps (last_policy_loss / last_value_loss / last_entropy_loss / last_total_loss) are the losses of ONE epoch,
  * update() with the reference's lr for ONE... (all four) epochs: the weights after the update.

    python tests/golden/make_golden_ppo.py      -> tests/golden/ppo.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402


def main():
    ref_loader.load()
    torch.set_num_threads(4)
    torch.manual_seed(0)
    np.random.seed(7)
    from src.agent.ppo_agent import BackgammonPPOAgent
    from src.environment.vec_bg_env import VectorizedBackgammonEnv
    dev = torch.device("cpu")
    N, T = 4, 130
    envs = VectorizedBackgammonEnv(num_envs=N, match_length=15, max_legal_moves=500, device=dev)
    agent = BackgammonPPOAgent(action_size=500, device=dev)
    sd0 = {k: v.detach().clone().numpy() for k, v in agent.policy_network.state_dict().items()}
    obs = envs.reset()
    boards, players = [], []
    for t in range(T):
        masks = envs.get_action_masks()
        boards.append(np.stack([e.board.tensor.numpy().copy() for e in envs.envs]))
        players.append(np.array([int(e.current_player) for e in envs.envs], np.int8))
        actions = agent.select_action(obs, masks)
        obs, rewards, dones, infos = envs.step(actions)
        for i in range(N):                                           # train.py:64-66
            agent.memory[-N + i]["reward"] = rewards[i].unsqueeze(0)
            agent.memory[-N + i]["done"] = dones[i].unsqueeze(0)
    mem = agent.memory
    M = len(mem)
    g = dict(
        N=N, T=T,
        obs=np.concatenate([m["observation"].numpy() for m in mem]).astype(np.float32),
        counts=np.array([int(m["action_mask"].sum().item()) for m in mem], np.int32),
        mask_is_prefix=np.array([bool((m["action_mask"][0, :int(m["action_mask"].sum().item())] == 1).all()) for m in mem]),
        actions=np.array([int(m["action"].item()) for m in mem], np.int32),
        logp=np.array([float(m["action_log_prob"].item()) for m in mem], np.float32),
        values=np.array([float(m["state_value"].item()) for m in mem], np.float32),
        rewards=np.array([float(m["reward"].item()) for m in mem], np.float32),
        dones=np.array([bool(m["done"].item()) for m in mem], np.uint8),
        boards=np.concatenate(boards).astype(np.int8), players=np.concatenate(players).astype(np.int8),
    )
    assert g["obs"].shape == (M, 198) and M == N * T
    rew = torch.tensor(g["rewards"]); dn = torch.tensor(g["dones"]).float()
    g["returns_interleaved"] = np.array(agent.compute_returns(rew, dn), np.float32)
    for k, v in sd0.items():
        g["w0." + k] = v
    # (a) lr = 0: losses of one epoch on the initial weights
    saved = list(agent.memory)
    for pg in agent.optimizer.param_groups:
        pg["lr"] = 0.0
    agent.update()
    g["loss_lr0"] = np.array([agent.last_policy_loss, agent.last_value_loss, agent.last_entropy_loss, agent.last_total_loss], np.float64)
    g["entropy_coef"] = np.float64(0.15)
    # (b) the real update (lr 1e-3, 4 epochs, Adam + GradScaler under CPU bf16 autocast)
    torch.manual_seed(0)
    agent2 = BackgammonPPOAgent(action_size=500, device=dev)
    agent2.policy_network.load_state_dict({k: torch.tensor(v) for k, v in sd0.items()})
    agent2.memory = saved
    agent2.update()
    for k, v in agent2.policy_network.state_dict().items():
        g["w1." + k] = v.detach().numpy()
    g["loss_lr1e3"] = np.array([agent2.last_policy_loss, agent2.last_value_loss, agent2.last_entropy_loss, agent2.last_total_loss], np.float64)
    np.savez_compressed(os.path.join(HERE, "ppo.npz"), **g)
    print("ppo.npz:", M, "samples; losses lr0", g["loss_lr0"], "lr1e-3", g["loss_lr1e3"], "dones", int(g["dones"].sum()),
          "passes", int((g["counts"] == 0).sum()), "prefix masks", bool(g["mask_is_prefix"].all()))


if __name__ == "__main__":
    main()
