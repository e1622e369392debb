"""INTEGRATION.md section 3 is executable documentation: run its ctypes stub against the built library."""
import os
import re

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_integration_md_ctypes_stub_runs(cuda_device):
  text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
  section = text[text.index("## 3. Raw C ABI from Python"):]
  code = re.search(r"```python\n(.*?)```", section, re.S).group(1)
  code = code.replace('"lagrange-dual-learning_b200/lib/libikl.so"', repr(os.path.join(ROOT, "lagrange-dual-learning_b200", "lib", "libikl.so")))
  scope = {}
  exec(compile(code, "INTEGRATION.md#3", "exec"), scope)
  torch.cuda.synchronize()
  out, dtheta = scope["out"], scope["dtheta"]
  assert torch.isfinite(out).all() and torch.isfinite(dtheta).all() and float(dtheta.abs().max()) > 0
  # same numbers as the maintained binding
  from ikl import ConstraintSpec, standard_config, fused
  spec = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_4obs_dynamic"), 3, 0.005, 1.0)
  viol, loss, d2 = fused.lagrangian_forward_backward(spec, scope["theta"], scope["q_des"], scope["obst"], [1.0] * 16)
  assert torch.equal(viol, out[:16]) and torch.equal(d2, dtheta)
