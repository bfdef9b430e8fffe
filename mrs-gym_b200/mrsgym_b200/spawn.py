"""Start-state sampling (the step before the hot path).

DefaultSpawn restates MRS.default_spawn_dist (/root/reference/mrsgym/MRS.py:69-78): z ~ U[1, 3],
xy ~ N(0, 1) pulled onto the unit disc when outside it (Util.SphereTransform with within=True,
/root/reference/mrsgym/Util.py:172-195).  sample_start_pos is the batched form of the rejection
loop of MRS.generate_start_pos (MRS.py:127-154): agents closer than 2*AGENT_RADIUS to another
agent of their env are re-drawn until no env has a collision.
"""
from __future__ import annotations

import torch


class DefaultSpawn:
    def __init__(self, n_agents, z_low=1.0, z_high=3.0, xy_radius=1.0):
        self.n_agents, self.z_low, self.z_high, self.xy_radius = n_agents, z_low, z_high, xy_radius

    def sample(self, sample_shape=()):
        shape = tuple(sample_shape) + (self.n_agents,)
        xy = torch.randn(shape + (2,)) * self.xy_radius
        mag = xy.norm(dim=-1, keepdim=True).clamp_min(self.xy_radius)
        xy = xy / mag * self.xy_radius
        z = self.z_low + (self.z_high - self.z_low) * torch.rand(shape + (1,))
        return torch.cat([xy, z], dim=-1)


def _draw(dist, E, N):
    """[E, N, 3] from a distribution whose sample() gives (N,3) or (3,)."""
    try:
        s = dist.sample((E,))
        if s.shape == (E, N, 3):
            return s.to(torch.float32)
    except Exception:
        pass
    rows = []
    for _ in range(E):
        s = dist.sample()
        if s.dim() == 1:
            s = torch.stack([dist.sample() for _ in range(N)], dim=0)
        rows.append(s)
    return torch.stack(rows, dim=0).to(torch.float32)


def sample_start_pos(dist, E, N, agent_radius, max_rounds=10000):
    pos = _draw(dist, E, N)
    eye = torch.eye(N, dtype=torch.bool)
    for _ in range(max_rounds):
        d = (pos.unsqueeze(2) - pos.unsqueeze(1)).norm(dim=-1)
        d = d.masked_fill(eye, float('inf'))
        hit = d < 2 * agent_radius
        # re-draw the higher-indexed agent of every colliding pair (keeps at least one of them)
        bad = torch.tril(hit, diagonal=-1).any(dim=-1)
        if not bool(bad.any()):
            return pos
        fresh = _draw(dist, E, N)
        pos = torch.where(bad.unsqueeze(-1), fresh, pos)
    raise RuntimeError('start-position rejection sampling did not converge (N_AGENTS too dense for START_POS)')
