#!/usr/bin/env python
"""Data-parallel policy gradient on a self-contained Pong ("ping-pong") — the RL toy of the reference's experimental tree.

Reference: experimental/atari_pingpong/{pingpong,base_agent,mlp_agent,cnn_agent,diff_trainer}.py + run.sh: every worker plays
its own episodes, the per-episode policy gradients are summed across workers through the framework and applied identically
everywhere; an `--init` run writes the initial checkpoint all workers load. gym/ALE are not in this image, so the court is
simulated here (80x80 binary frames like the reference's pre-processed Atari frames; frame differences as the observation).

    python experimental/pingpong/pingpong.py --init --checkpoint /tmp/pp.npz
    kungfu-run -np 4 python experimental/pingpong/pingpong.py --checkpoint /tmp/pp.npz --batch-size 2 --episodes 40
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import kungfu_b200.torch as kf  # noqa: E402


class Court:
    """A one-paddle Pong: the ball bounces off three walls, the agent's paddle guards the right edge.
    Reward +1 for a return, -1 for a miss; an episode is `points` balls."""
    H = W = 80
    PADDLE = 12

    def __init__(self, seed: int, points: int = 5):
        self.rng = np.random.RandomState(seed)
        self.points = points
        self.reset()

    def _serve(self):
        self.bx, self.by = 10.0, float(self.rng.randint(10, self.H - 10))
        ang = self.rng.uniform(-0.9, 0.9)
        self.vx, self.vy = 2.0 * np.cos(ang), 2.0 * np.sin(ang)

    def reset(self):
        self.py = self.H // 2
        self.left = self.points
        self._serve()
        return self.frame()

    def frame(self) -> np.ndarray:
        f = np.zeros((self.H, self.W), np.float32)
        f[int(np.clip(self.by, 0, self.H - 1)), int(np.clip(self.bx, 0, self.W - 1))] = 1.0
        lo = int(np.clip(self.py - self.PADDLE // 2, 0, self.H - self.PADDLE))
        f[lo:lo + self.PADDLE, self.W - 2] = 1.0
        return f

    def step(self, action: int):
        """action 0 = up, 1 = down. Returns (frame, reward, done)."""
        self.py = int(np.clip(self.py + (-3 if action == 0 else 3), self.PADDLE // 2, self.H - self.PADDLE // 2))
        self.bx += self.vx
        self.by += self.vy
        if self.by < 0 or self.by > self.H - 1:
            self.vy = -self.vy
            self.by = float(np.clip(self.by, 0, self.H - 1))
        if self.bx < 0:
            self.vx = -self.vx
            self.bx = 0.0
        reward = 0.0
        if self.bx >= self.W - 2:
            hit = abs(self.by - self.py) <= self.PADDLE / 2 + 1
            reward = 1.0 if hit else -1.0
            self.left -= 1
            self._serve()
        return self.frame(), reward, self.left == 0


class MLPAgent(torch.nn.Module):
    """Karpathy-style 2-layer policy: P(up | frame difference)."""

    def __init__(self, inputs: int = 80 * 80, hidden: int = 64):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(inputs, hidden), torch.nn.ReLU(), torch.nn.Linear(hidden, 1))

    def forward(self, x):
        return self.net(x.flatten(1)).squeeze(-1)


class CNNAgent(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Conv2d(1, 8, 8, stride=4), torch.nn.ReLU(), torch.nn.Conv2d(8, 16, 4, stride=2), torch.nn.ReLU(),
                                       torch.nn.Flatten(), torch.nn.Linear(16 * 8 * 8, 1))

    def forward(self, x):
        return self.net(x.view(-1, 1, 80, 80)).squeeze(-1)


def discount(rewards, gamma=0.99):
    out, run = np.zeros(len(rewards), np.float32), 0.0
    for i in reversed(range(len(rewards))):
        if rewards[i] != 0:
            run = 0.0  # a point was scored: game boundary
        run = run * gamma + rewards[i]
        out[i] = run
    return out


def play_episode(env: Court, agent: torch.nn.Module, rng: np.random.RandomState):
    obs, acts, rews = [], [], []
    prev = np.zeros((Court.H, Court.W), np.float32)
    cur, done = env.reset(), False
    while not done:
        x = cur - prev
        prev = cur
        with torch.no_grad():
            p_up = torch.sigmoid(agent(torch.from_numpy(x)[None])).item()
        a = 0 if rng.uniform() < p_up else 1
        cur, r, done = env.step(a)
        obs.append(x)
        acts.append(a)
        rews.append(r)
    return np.stack(obs), np.array(acts), np.array(rews, np.float32)


def make_optimizer(name, params, lr):
    if name == "adam":
        return torch.optim.Adam(params, lr=lr)
    if name == "rmsp":
        return torch.optim.RMSprop(params, lr=lr, alpha=0.99)
    return torch.optim.SGD(params, lr=lr)


def main():
    ap = argparse.ArgumentParser(description="RL example.")
    ap.add_argument("--agent", default="mlp", choices=["mlp", "cnn"])
    ap.add_argument("--optimizer", default="rmsp", help="sgd | adam | rmsp")
    ap.add_argument("--batch-size", type=int, default=10, help="episodes per worker per update")
    ap.add_argument("--episodes", type=int, default=200, help="max number of episodes per worker")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--checkpoint", default="", help="path to checkpoint file")
    ap.add_argument("--init", action="store_true", help="generate the init checkpoint and exit")
    a = ap.parse_args()
    torch.manual_seed(0)
    agent = MLPAgent() if a.agent == "mlp" else CNNAgent()
    if a.init:
        assert a.checkpoint, "--init needs --checkpoint"
        np.savez(a.checkpoint, **{k: v.numpy() for k, v in agent.state_dict().items()})
        print("wrote", a.checkpoint)
        return
    if a.checkpoint and os.path.exists(a.checkpoint):
        with np.load(a.checkpoint) as z:
            agent.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
    rank, size = kf.current_rank(), kf.current_cluster_size()
    opt = kf.optimizers.SynchronousSGDOptimizer(make_optimizer(a.optimizer, agent.parameters(), a.lr), named_parameters=agent.named_parameters())
    kf.broadcast_parameters(agent.state_dict())
    env, rng = Court(seed=1000 + rank), np.random.RandomState(2000 + rank)
    running = None
    for ep in range(0, a.episodes, a.batch_size):
        opt.zero_grad()
        total = 0.0
        for _ in range(a.batch_size):
            obs, acts, rews = play_episode(env, agent, rng)
            adv = discount(rews)
            adv = (adv - adv.mean()) / (adv.std() + 1e-6)
            logits = agent(torch.from_numpy(obs))
            target_up = torch.from_numpy((acts == 0).astype(np.float32))
            nll = torch.nn.functional.binary_cross_entropy_with_logits(logits, target_up, reduction="none")
            (nll * torch.from_numpy(adv)).sum().div(a.batch_size).backward()
            total += float(rews.sum())
        opt.step()  # gradients averaged over all workers' episodes
        mean = kf.ops.all_reduce_fn(torch.tensor([total / a.batch_size])).item() / size
        running = mean if running is None else 0.9 * running + 0.1 * mean
        if rank == 0:
            print(f"episode {(ep + a.batch_size) * size:6d} (all workers)  mean reward {mean:+.2f}  running {running:+.2f}", flush=True)
    digest = torch.cat([p.detach().reshape(-1) for p in agent.parameters()]).double().sum().item()
    same = kf.ops.all_reduce_fn(torch.tensor([digest], dtype=torch.float64), op="max").item() == kf.ops.all_reduce_fn(
        torch.tensor([digest], dtype=torch.float64), op="min").item()
    if rank == 0:
        print("RESULT: replicas identical =", same, flush=True)
    if a.checkpoint and rank == 0:
        np.savez(a.checkpoint, **{k: v.numpy() for k, v in agent.state_dict().items()})


if __name__ == "__main__":
    main()
