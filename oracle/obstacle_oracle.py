"""CPU restatement (torch) of the kinematic obstacle advance of the "dynamic_env" environment.

TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else); product path: csrc/obstacle_core.cuh behind
agx_obstacle_step.  PARITY UNPINNED: the reference sets the obstacles' twist from env_actions before every physics step
(env_manager/obstacle_manager.py:40-44) and lets PhysX move them; PhysX is not observable here and the reference holds no
test or golden vector for it.  The spec is ours -- the robot integrator's semi-implicit form (hp1_oracle.rigid_body_integrate
without forces): v *= max(0, 1 - dt c_lin), w likewise, x += dt v, q = normalize(dq(w dt) (x) q)."""
import torch

from . import hp1_oracle as O


def obstacle_step(state, twist, dt, substeps, linear_damping=0.1, angular_damping=0.1):
    """state [N,A,13] (x, q xyzw, v, w), twist [N,A,6] or None.  Returns the new state (input untouched)."""
    s = state.clone().reshape(-1, 13)
    tw = None if twist is None else twist.reshape(-1, 6)
    kl, ka = max(0.0, 1.0 - dt * linear_damping), max(0.0, 1.0 - dt * angular_damping)
    x, q, v, w = s[:, 0:3].clone(), s[:, 3:7].clone(), s[:, 7:10].clone(), s[:, 10:13].clone()
    for _ in range(substeps):
        if tw is not None:
            v, w = tw[:, 0:3].clone(), tw[:, 3:6].clone()
        v, w = v * kl, w * ka
        x = x + v * dt
        wn = torch.norm(w, dim=1, keepdim=True)
        half = 0.5 * dt * wn
        s_over = torch.where(wn > 0, torch.sin(half) / torch.where(wn > 0, wn, torch.ones_like(wn)), torch.zeros_like(wn))
        qn = O.quat_mul(torch.cat([w * s_over, torch.cos(half)], dim=1), q)
        q = qn / torch.norm(qn, dim=1, keepdim=True)
    return torch.cat([x, q, v, w], dim=1).reshape(state.shape)
