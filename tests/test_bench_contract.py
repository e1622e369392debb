"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm runs on the CPU and prints one
line with the agreed keys; the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches")


HAVE_REF_TREE = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "scripts", "trainer.py"))


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_prints_one_contract_line(kind):
  """kind "reference": the unmodified reference from baseline/_ref (tools/install_ref.py); "port": the oracle restatement,
  which is what the arm falls back to when that tree is absent (forced here with --force-port)."""
  if kind == "reference" and not HAVE_REF_TREE:
    pytest.skip("baseline/_ref not installed")
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-batch-log2", "12"] + (["--force-port"] if kind == "port" else []),
                       capture_output=True, text=True, timeout=600)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for k in REQUIRED:
    assert k in d, k
  assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
  assert d["metric"] == "fk_penalty_fwd_bwd_samples_per_s" and d["unit"] == "samples/s" and d["dtype"] == "f32" and d["vs_baseline"] is None
  assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert "workload" in d["config"] and d["gpu_launches"] == 0 and d["value"] > 0
  # the config is the ours arm's, word for word (same function); what this arm sampled of it is stated beside it
  sys.path.insert(0, ROOT)
  import argparse
  import bench
  want = bench.workload_config(argparse.Namespace(batch_log2=24, layout="planes"), 1, "epoch")
  assert d["config"] == want and d["reference_sample"]["batch_per_step"] == 1 << 12 and d["reference_sample"]["device"] == "cpu"


def test_reference_arm_other_ranks_exit_quietly():
  env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="this is the no-GPU behaviour")
def test_product_arm_refuses_to_run_without_cuda():
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
  assert out.returncode != 0 and "no CPU fallback" in out.stderr
