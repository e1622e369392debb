"""The IK network (reference: models/simple_network.py:5-31).  Stays on PyTorch / cuBLAS: it is a stack of small dense
layers, not part of the FK/penalty path.  Parameter names (input, hidden.<i>.0, output) match the reference so
state_dicts are interchangeable."""
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

  def forward(self, x):
    h = F.relu(self.input(x))
    for block in self.hidden:
      h = block(h)
    return self.output(h)
