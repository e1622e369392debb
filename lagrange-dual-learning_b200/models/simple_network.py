"""The IK network (reference: models/simple_network.py:5-31).  Stays on PyTorch / cuBLAS: it is a stack of small dense
layers, not part of the FK/penalty path.  Parameter names (input, hidden.<i>.0, output) match the reference so
state_dicts are interchangeable."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _hidden_block(width, dropout):
  return nn.Sequential(nn.Linear(width, width, bias=True), nn.ReLU(inplace=True), nn.Dropout(p=dropout))


class SimpleNetwork(nn.Module):
  """input_dim -> hidden_units -> (depth x [Linear, ReLU, Dropout]) -> output_dim."""

  def __init__(self, input_dim, output_dim, hidden_units=40, depth=8, dropout=0.05):
    super(SimpleNetwork, self).__init__()
    self.input = nn.Linear(input_dim, hidden_units, bias=True)
    self.hidden = nn.ModuleList(_hidden_block(hidden_units, dropout) for _ in range(depth))
    self.output = nn.Linear(hidden_units, output_dim, bias=True)

  def features(self, x):
    h = F.relu(self.input(x))
    for block in self.hidden:
      h = block(h)
    return h

  def forward(self, x):
    return self.output(self.features(x))

  def forward_planes(self, x):
    """Same numbers as forward(), laid out for the fused IK kernels: the last layer is evaluated transposed
    (W h^T + b, one (output_dim, B) GEMM) and returned as a (B, output_dim) VIEW of that buffer, i.e. one contiguous
    plane per joint.  Not in the reference; autograd flows through it as usual."""
    h = self.features(x)
    planes = torch.addmm(self.output.bias.unsqueeze(1), self.output.weight, h.t())
    return planes.t()
