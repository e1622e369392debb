"""Synthetic batches with the distribution the benchmark is quoted on (SURVEY 8d / BASELINE.md):

  theta ~ U(-2.5, 2.5)^(B x 3)            slightly wider than the +-2.094395 limits so both JL branches fire
  target = FK(theta'), theta' ~ U(-2.094395, 2.094395)   reachable, like kinematics/dataset.py:61-68
  obstacles (B,4,4): x, y ~ U(-1.2, 1.2) (dataset.py:101), r = 0.4, 4th column 0 (dataset.py:80-85); no rejection sampling

`device="cpu"` gives host tensors (tests feed the same values to the oracle); on CUDA the targets come from
this package's own FK kernel.
"""
import torch

LIMIT = 2.094395
SHIPPED_LAMBDA_OBS4_DYN = [4.060915, 0.0, 0.0, 0.280665, 0.0, 0.0, 1.449509, 0.0, 0.0, 1.679491, 0.0, 0.0, 1.564413,
                           0.0, 0.0, 1.505343]   # values of the reference's shipped ik_threelink_obs4_dyn multipliers


def make_batch(B, seed=0, device="cpu", num_links=3):
  """Row-major tensors as the reference's DataLoader would hand them: theta (B,J), q_ee_desired (B,3), obstacles (B,4,4)."""
  dev = torch.device(device)
  g = torch.Generator(device=dev).manual_seed(seed)
  theta = (torch.rand(B, num_links, generator=g, device=dev) * 2 - 1) * 2.5
  theta_t = (torch.rand(B, num_links, generator=g, device=dev) * 2 - 1) * LIMIT
  obst = torch.zeros(B, 4, 4, device=dev)
  obst[:, :, :2] = (torch.rand(B, 4, 2, generator=g, device=dev) * 2 - 1) * 1.2
  obst[:, :, 2] = 0.4
  if dev.type == "cuda":
    from . import ops
    q_des = ops.fk_forward_kinematics(theta_t)[0].contiguous()
  else:
    phi = torch.cumsum(theta_t, dim=1)
    q_des = torch.stack([torch.cos(phi).sum(1), torch.sin(phi).sum(1), phi[:, -1]], dim=1)
  return theta, q_des, obst


def to_planes(theta, q_des, obst, num_obstacles=4):
  """Re-lay a row-major batch as structure-of-arrays planes; returns VIEWS with the reference's logical shapes
  ((B,J), (B,2), (B,n,3)) whose strides describe the planes, so they can be passed to the same API."""
  B = theta.shape[0]
  th_p = theta.t().contiguous()                        # (J, B)
  tg_p = q_des[:, :2].t().contiguous()                 # (2, B)
  ob_p = obst[:, :num_obstacles, :3].permute(1, 2, 0).contiguous()   # (n, 3, B)
  return th_p.t(), tg_p.t(), ob_p.permute(2, 0, 1)
