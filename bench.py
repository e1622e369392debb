#!/usr/bin/env python
"""Benchmark of the IK-Lagrangian hot path (BASELINE.json metric: FK+penalty fwd+bwd samples/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one fused forward+backward pass of the hot path over one synthetic batch: per sample the 3-link FK,
the end-effector error, 3 joint-limit penalties and 12 tip-vs-circle penalties (4 per-sample dynamic obstacles),
the 16 batch sums, the Lagrangian and d(Lagrangian)/d(theta)  (config 5 of BASELINE.json:
three_link_arm + cfg_joint_limits_and_4obs_dynamic, batch 2^24 per GPU).  With N > 1 every rank owns its own 2^24
shard (weak scaling) and each step ends with an NCCL all-reduce of the 16 violation sums so the multiplier update
sees identical numbers on every rank.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM); `e2e` is the same
metric through the host-buffer API (pinned host tensors in, gradient + sums back on the host, PCIe copies inside the
timed region); `roofline` is the fused kernel against the measured HBM copy bandwidth; `cpu_baseline` is the oracle
port of the reference's PyTorch-eager CPU path timed on this box's host cores.

--impl reference times that CPU path alone (the reference has no GPU-specific code; its arithmetic is PyTorch eager).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "lagrange-dual-learning_b200")
for p in (ROOT, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)          # (--impl reference with baseline/_ref takes PKG off the path again: ref_harness)

METRIC = "fk_penalty_fwd_bwd_samples_per_s"
UNIT = "samples/s"
ALGO_BYTES_PER_SAMPLE = 80          # read theta 12 + target xy 8 + 4x(x,y,r) 48, write dtheta 12
CONFIG_NAME = "cfg_joint_limits_and_4obs_dynamic"
GOOD_SLOPE, BAD_SLOPE = 0.005, 1.0   # scripts/experiments/ik_threelink_obs4_dyn.sh:15-16
MULTIPLIER_LR = 1e-4                 # scripts/options.py:17 (the reference's default dual step size)


def bind_to_gpu_numa_node(index):
  """Run this process (and therefore its pinned host buffers, by first touch) on the CPU socket the GPU hangs off:
  host->device copies from the far socket cross the inter-socket link and stop scaling beyond a few GPUs."""
  try:
    out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                         capture_output=True, text=True, timeout=20).stdout.strip()
    bus = out.lower()
    if bus.startswith("00000000:"):
      bus = bus[4:]
    node_file = "/sys/bus/pci/devices/%s/numa_node" % bus
    node = int(open(node_file).read().strip())
    if node < 0:
      return None
    cpus = []
    for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
      lo, _, hi = part.partition("-")
      cpus.extend(range(int(lo), int(hi or lo) + 1))
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
    if allowed:
      os.sched_setaffinity(0, allowed)
      return {"numa_node": node, "cpus": len(allowed)}
  except Exception:
    return None
  return None


def measured_peak():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  try:
    with open(path) as f:
      return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
  except Exception:
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons while the GPU is busy (B200_PROFILING.md recipe)."""
  FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.lines = []
    self.proc = None
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                    "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._pump, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.lines.append((time.time(), line.strip()))

  def stop(self, t0, t1):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, smax, reasons, power = [], [], set(), []
    names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    for ts, line in self.lines:
      if ts < t0 or ts > t1 + 0.15:
        continue
      parts = [x.strip() for x in line.split(",")]
      if len(parts) < 7:
        continue
      try:
        sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
      except ValueError:
        continue
      for n, v in zip(names, parts[3:7]):
        if v.lower().startswith("active"):
          reasons.add(n)
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
            "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_oracle_samples_per_s(batch_log2, steps, warmup):
  """The reference's CPU path (PyTorch eager fwd + autograd bwd) restated in oracle/ik_oracle.py, all host threads."""
  import torch
  from oracle import ik_oracle as O
  from ikl import synthetic, standard_config
  torch.set_num_threads(os.cpu_count() or 1)
  B = 1 << batch_log2
  ospec = O.spec_from_json(standard_config(CONFIG_NAME), 3, GOOD_SLOPE, BAD_SLOPE)
  theta, q_des, obst = synthetic.make_batch(B, seed=0)
  lam = torch.tensor(synthetic.SHIPPED_LAMBDA_OBS4_DYN)
  for _ in range(warmup):
    O.lagrangian_with_grad(theta, q_des, lam, ospec, obst)
  t0 = time.perf_counter()
  for _ in range(steps):
    O.lagrangian_with_grad(theta, q_des, lam, ospec, obst)
  dt = (time.perf_counter() - t0) / steps
  return B / dt, dt, torch.get_num_threads(), B


def cpu_c_oracle_samples_per_s(batch_log2):
  """The plain-C / OpenMP restatement (oracle/ik_oracle.c, float arithmetic, closed-form gradient): what a tuned CPU port
  of the same maths reaches on this host -- reported as context next to the reference-style PyTorch baseline."""
  import numpy as np
  from oracle import build_oracle as C, ik_oracle as O
  from ikl import synthetic, standard_config
  B = 1 << batch_log2
  ospec = O.spec_from_json(standard_config(CONFIG_NAME), 3, GOOD_SLOPE, BAD_SLOPE)
  theta, q_des, obst = (t.numpy() for t in synthetic.make_batch(B, seed=0))
  lam = np.asarray(synthetic.SHIPPED_LAMBDA_OBS4_DYN, dtype=np.float32)
  C.lagrangian_f32(ospec, theta, q_des, obst, lam)
  best = 1e9
  for _ in range(5):
    t0 = time.perf_counter()
    C.lagrangian_f32(ospec, theta, q_des, obst, lam)
    best = min(best, time.perf_counter() - t0)
  return {"value": B / best, "unit": UNIT, "cores": C.max_threads(), "kind": "port",
          "sample": "oracle/ik_oracle.c (C + OpenMP, fp32 maths, closed-form gradient) on a 2^%d-sample batch, best of 5" % batch_log2}


SHIPPED_LAMBDA_OBS4_DYN = [4.060915, 0.0, 0.0, 0.280665, 0.0, 0.0, 1.449509, 0.0, 0.0, 1.679491, 0.0, 0.0, 1.564413,
                           0.0, 0.0, 1.505343]   # resources/pretrained_models/ik_threelink_obs4_dyn/models/weights_650/multipliers.pth


def reference_tree_samples_per_s(batch_log2, steps, warmup):
  """The UNMODIFIED reference (baseline/_ref, installed by tools/install_ref.py) on the host cores: its own
  IkLagrangeDualTrainer.process_batch (scripts/trainer.py:159-249) over its own kinematics/*.py, then
  lagrange_loss.backward() (trainer.py:264), with theta as a leaf in place of the network (the MLP is not part of the
  hot path).  Nothing of this repo's product tree is importable in this process (ref_harness removes it from sys.path)."""
  import tempfile
  import torch
  from baseline import ref_harness as H
  trainer_mod, options_mod = H.load_reference_scripts(dropin=False)
  H.assert_resolution(trainer_mod, dropin=False)
  assert trainer_mod.device == "cpu"
  torch.set_num_threads(os.cpu_count() or 1)
  B = 1 << batch_log2
  tmp = tempfile.mkdtemp(prefix="ikl_ref_arm_")
  opt = H.reference_options(options_mod, [
      "--model_name", "bench", "--config_file", os.path.join(H.REF_TREE, "resources", CONFIG_NAME + ".json"), "--log_dir", tmp,
      "--num_workers", "0", "--train_dataset_size", "8", "--val_dataset_size", "8", "--penalty_good_slope", str(GOOD_SLOPE),
      "--penalty_bad_slope", str(BAD_SLOPE)])
  import contextlib
  with H.git_worktree(tmp), contextlib.redirect_stdout(sys.stderr):     # the trainer's banner must not share stdout with the JSON line
    tr = trainer_mod.IkLagrangeDualTrainer(opt)
  # the synthetic batch of ikl.synthetic.make_batch, restated here because that package is out of reach
  g = torch.Generator().manual_seed(0)
  theta = (torch.rand(B, 3, generator=g) * 2 - 1) * 2.5
  phi = torch.cumsum((torch.rand(B, 3, generator=g) * 2 - 1) * 2.094395, dim=1)
  obst = torch.zeros(B, 4, 4)
  obst[:, :, :2] = (torch.rand(B, 4, 2, generator=g) * 2 - 1) * 1.2
  obst[:, :, 2] = 0.4
  q_des = torch.stack([torch.cos(phi).sum(1), torch.sin(phi).sum(1), phi[:, -1]], dim=1)
  lam = torch.tensor(SHIPPED_LAMBDA_OBS4_DYN)

  class ThetaLeaf(torch.nn.Module):
    def __init__(self):
      super().__init__()
      self.theta = torch.nn.Parameter(theta)

    def forward(self, x):
      return self.theta
  tr.model = ThetaLeaf()
  x = torch.zeros(B, 20)

  def step():
    out = tr.process_batch({"input_tensor": x, "q_ee_desired": q_des, "dynamic_obstacles": obst}, lam)
    tr.model.zero_grad()
    out["lagrange_loss"].backward()
  for _ in range(warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  dt = (time.perf_counter() - t0) / steps
  return B / dt, dt, torch.get_num_threads(), B


def parity_summary():
  """Measured parity of the benchmarked kernels against the float64 oracle at 2^24 samples (tools/parity_report.py, run on the
  B200 and committed as profiles/r02/parity_report.json) -- numbers, not a pass/fail flag; the tests enforce the tolerances."""
  path = os.path.join(ROOT, "profiles", "r02", "parity_report.json")
  try:
    with open(path) as f:
      rep = json.load(f)
    c5 = rep["configs"][CONFIG_NAME]["planes"]
    worst_sum = max(rep["configs"][n][l]["sums_max_rel_to_sum_abs_terms"] for n in rep["configs"] for l in ("planes", "rows"))
    return {"source": "profiles/r02/parity_report.json (tools/parity_report.py, float64 oracle, 2^24 samples, all five configs)",
            "tolerance": "1e-5 relative (north_star)",
            "positions_max_rel_to_reach": rep["positions"]["max_rel_to_reach_3"],
            "penalties_max_rel": max(rep["penalties"]["joint_limit_max_rel_to_max_ref"], rep["penalties"]["circle_max_rel_to_max_ref"]),
            "dtheta_p99_99_rel": c5["dtheta_p99_99_rel_away_from_kinks"], "dtheta_max_rel_away_from_kinks": c5["dtheta_max_rel_away_from_kinks"],
            "dtheta_fraction_beyond_1e-5": c5["fraction_above_1e-5_S_away_from_kinks"],
            "dtheta_fraction_beyond_conditioning_bound": c5["fraction_above_conditioning_bound"],
            "kink_fraction_masked": rep["configs"][CONFIG_NAME]["kink_fraction"],
            "sums_max_rel_to_sum_abs_terms_all_configs": worst_sum,
            "lambda_after_12_dual_steps_rel": rep["lambda_after_12_dual_steps"]["max_rel_err"]}
  except Exception as e:
    return {"unavailable": str(e)[:200]}


def cpu_baseline_leg(args):
  """cpu_baseline of the ours arm: the reference's own code from baseline/_ref in a fresh process (so that its `kinematics`
  package, not this repo's drop-in, is the one imported) on a bounded 2^cpu_batch_log2 sample; the oracle port in-process
  when the reference tree is not installed."""
  if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "scripts", "trainer.py")):
    try:
      r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5", "--warmup", "1",
                          "--cpu-batch-log2", str(args.cpu_batch_log2)], capture_output=True, text=True, timeout=600,
                         env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
      ref_line = json.loads(r.stdout.strip().splitlines()[-1])
      return ref_line["cpu_baseline"]
    except Exception as e:
      sys.stderr.write("cpu_baseline: reference subprocess failed (%s); timing the oracle port instead\n" % (e,))
  sps, dtc, cores, Bc = cpu_oracle_samples_per_s(args.cpu_batch_log2, 5, 1)
  return {"value": sps, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": "oracle port of the reference PyTorch-eager CPU path (fwd + autograd bwd, fp32) on a 2^%d-sample "
                    "batch of the same workload, mean of 5 after 1 warm-up (%.0f ms each)" % (args.cpu_batch_log2, dtc * 1e3)}


def workload_config(args, world, exchange_mode):
  """The `config` of the bench line -- the workload, identical in both arms (`--impl reference` times a bounded sample of it on
  the host cores and says which in cpu_baseline.sample).  exchange_mode: "epoch" | "peer" | "nccl" (N > 1 only)."""
  B = 1 << args.batch_log2
  epoch = exchange_mode == "epoch"
  return {"workload": "three_link_arm + %s, K=16: fused FK+penalty+Lagrangian fwd+bwd, batch 2^%d per GPU" % (CONFIG_NAME, args.batch_log2),
          "batch_per_gpu": B, "layout": args.layout, "slopes": [GOOD_SLOPE, BAD_SLOPE], "weights": "by value (host lambda)",
          "l2": "no flush: each step streams %.2f GB of inputs (> 126 MB L2)" % (B * (100 if args.layout == "rows" else 68) / 1e9),
          "step": ("one batch of a dual step (scripts/trainer.py:267-286): fused fwd+bwd launch, 16 sums added to the rank's float64 running "
                   "sum; the region's LAST launch closes the epoch in the same kernel (exchange of the running sums + lambda update)"
                   if epoch else "fused fwd+bwd launch with the 16 sums all-reduced every step"),
          "parallelism": ("single GPU" if world == 1 else
                          "batch shard per GPU; running sums exchanged ONCE per dual step, inside the closing launch, over NVLink peer "
                          "memory (ikl_lagrangian_forward_backward_dual_step); no other collective in the timed region" if epoch else
                          "batch shard per GPU; the 16 violation sums all-reduced inside the fused kernel over NVLink peer memory, every step"
                          if exchange_mode == "peer" else
                          "batch shard per GPU; asynchronous NCCL all-reduce of the 16 violation sums per step (overlaps the next step's kernel)")}


def run_reference(args):
  """--impl reference: the reference's CPU path (PyTorch eager forward + autograd backward, fp32, all host threads),
  exactly K timed steps after W warm-ups; the per-step sample is sized so that the run stays within minutes.  With
  baseline/_ref present (tools/install_ref.py) it is the reference's own code (`kind: "reference"`); otherwise the oracle
  port of it (`kind: "port"`, bit-identical outputs: tests/test_oracle_pinned.py)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  os.environ["CUDA_VISIBLE_DEVICES"] = ""        # this arm is the host-core path; the reference would otherwise pick the GPU
  steps, warmup = max(args.steps, 1), max(args.warmup, 1)
  batch_log2 = args.cpu_batch_log2 if steps <= 50 else (args.cpu_batch_log2 - 2 if steps <= 500 else args.cpu_batch_log2 - 4)
  have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "scripts", "trainer.py")) and not args.force_port
  if have_ref:
    sps, dt, cores, B = reference_tree_samples_per_s(batch_log2, steps, warmup)
    kind = "reference"
    sample = ("the unmodified reference from baseline/_ref: IkLagrangeDualTrainer.process_batch (scripts/trainer.py:159-249) + "
              "lagrange_loss.backward(), theta as a leaf, fp32, batch 2^%d per step, %d steps" % (batch_log2, steps))
  else:
    sps, dt, cores, B = cpu_oracle_samples_per_s(batch_log2, steps, warmup)
    kind = "port"
    sample = "oracle port of the reference PyTorch-eager CPU path (fwd + autograd bwd, fp32), batch 2^%d per step, %d steps" % (
        batch_log2, steps)
  line = {
      "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
      "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
      "data": "synthetic",
      # the workload is the ours arm's, word for word; what this arm actually ran of it is in reference_sample / cpu_baseline.sample
      "config": workload_config(args, int(os.environ.get("WORLD_SIZE", "1")), args.exchange),
      "reference_sample": {"batch_per_step": B, "device": "cpu", "layout": "rows (the reference's own tensors)",
                           "note": "bounded sample of the workload in `config`: same generator, same config file, same slopes and multipliers"},
      "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
      "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }
  print(json.dumps(line))


def reference_trainer_over_dropins(theta, q_des, obst, lam_dev, dev, batch_log2=22):
  """process_batch (scripts/trainer.py:159-249) + lagrange_loss.backward() of the reference's own trainer class, unmodified,
  on the GPU over the function-level drop-ins (one libikl.so launch per FK / penalty call, forward and backward), theta as
  a leaf.  The same harness as tests/test_gpu_reference_trainer.py."""
  import tempfile
  import torch
  from ikl import _cabi
  from baseline import ref_harness as H
  if not H.have_reference_tree():
    return {"unavailable": "baseline/_ref is not installed"}
  trainer_mod, options_mod = H.load_reference_scripts()
  H.assert_resolution(trainer_mod)
  tmp = tempfile.mkdtemp(prefix="ikl_bench_dropin_")
  cfg_path = os.path.join(H.REF_TREE, "resources", CONFIG_NAME + ".json")
  opt = H.reference_options(options_mod, ["--model_name", "bench", "--config_file", cfg_path, "--log_dir", tmp, "--num_workers", "0",
                                          "--train_dataset_size", "16", "--val_dataset_size", "16", "--batch_size", "16",
                                          "--penalty_good_slope", str(GOOD_SLOPE), "--penalty_bad_slope", str(BAD_SLOPE), "--hidden_units", "16"])
  import contextlib
  with H.git_worktree(tmp), contextlib.redirect_stdout(sys.stderr):
    tr = trainer_mod.IkLagrangeDualTrainer(opt)
  Bd = 1 << batch_log2

  class FixedTheta(torch.nn.Module):
    def __init__(self, t):
      super().__init__()
      self.theta = torch.nn.Parameter(t.clone())

    def forward(self, x):
      return self.theta
  tr.model = FixedTheta(theta[:Bd]).to(dev)
  inputs = {"input_tensor": torch.zeros(Bd, 20, device=dev), "q_ee_desired": q_des[:Bd].clone(), "joint_theta": theta[:Bd].clone(),
            "dynamic_obstacles": obst[:Bd].clone()}

  def step():
    out = tr.process_batch(dict(inputs), lam_dev)
    tr.model.zero_grad()
    out["lagrange_loss"].backward()
  step()
  torch.cuda.synchronize()
  n0 = _cabi.launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(3):
    step()
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 3
  res = {"samples_per_s": Bd / (ms * 1e-3), "ms_per_step": ms, "batch": Bd, "libikl_launches_per_step": (_cabi.launch_count() - n0) // 3,
         "note": "the reference's own IkLagrangeDualTrainer.process_batch + backward(), file baseline/_ref/scripts/trainer.py unmodified, FK and "
                 "every penalty call one CUDA drop-in launch each way; the rest is the trainer's own torch glue (slice-assign, sums)"}
  del tr, inputs
  torch.cuda.empty_cache()
  return res


def multi_gpu_check(args, spec, dev, world, rank, lam, exchange):
  """Correctness of the multi-GPU path, checked in the same run as the numbers (all ranks take part):
    (1) the all-reduced sums are bit-identical on every rank;
    (2) they equal the float64 sum of the ranks' local sums to <= 1e-5 of sum_r |partial_r| (north_star tolerance);
    (3) three dual epochs (Adam steps + dual-ascent update) of the real trainer on device-resident data leave lambda and
        the network weights bit-identical on every rank."""
  import torch
  import torch.distributed as dist
  from ikl import fused, synthetic
  K = spec.num_constraints
  Bc = 1 << 20
  theta, q_des, obst = synthetic.make_batch(Bc, seed=500 + rank, device=str(dev))
  th, tg, ob = synthetic.to_planes(theta, q_des, obst)
  local, _, _ = fused.lagrangian_forward_backward(spec, th, tg, ob, lam)
  if exchange is not None:
    glob, gloss, _ = fused.lagrangian_forward_backward(spec, th, tg, ob, lam, exchange=exchange)
    glob = glob.clone()
  else:
    glob = local.clone()
    dist.all_reduce(glob)
  gathered = [torch.zeros_like(glob) for _ in range(world)]
  dist.all_gather(gathered, glob)
  sums_identical = all(torch.equal(g, gathered[0]) for g in gathered)
  parts = [torch.zeros(K, dtype=torch.float64, device=dev) for _ in range(world)]
  dist.all_gather(parts, local.double())
  want = torch.stack(parts).sum(0)
  scale = torch.stack(parts).abs().sum(0)
  rel = float(((glob.double() - want).abs() / scale.clamp(min=1e-30)).max())
  res = {"exchange": "in-kernel peer memory" if exchange is not None else "nccl all_reduce",
         "sums_bit_identical_across_ranks": bool(sums_identical), "sums_max_rel_err_vs_f64_sum_of_partials": rel,
         "sums_ok": bool(sums_identical and rel <= 1e-5)}
  # (3) the trainer under torch.distributed
  try:
    import tempfile
    sys.path.insert(0, os.path.join(PKG, "fused_overlay"))
    from ikl.configs import write_config
    from scripts.options import Options
    from scripts.trainer import IkLagrangeDualTrainer
    import contextlib
    tmp = tempfile.mkdtemp(prefix="ikl_bench_multi_%d_" % rank)
    cfg_path = write_config(__import__("ikl").standard_config(CONFIG_NAME), os.path.join(tmp, "cfg.json"))
    opt = Options().parse(["--model_name", "check", "--config_file", cfg_path, "--log_dir", tmp, "--device_data", "--train_dataset_size", "65536",
                           "--val_dataset_size", "16384", "--batch_size", "16384", "--lagrange_iters", "3", "--train_iters", "2",
                           "--hidden_units", "32", "--initial_lambda", "1", "--penalty_good_slope", str(GOOD_SLOPE), "--model_save_hz", "1000"])
    with contextlib.redirect_stdout(sys.stderr):
      tr = IkLagrangeDualTrainer(opt)
      for _ in range(3):
        tr.train_with_relaxation(tr.lambda_multipliers)
        tr.lambda_multipliers = tr.update_multipliers(tr.lambda_multipliers)
    flat = torch.cat([tr.lambda_multipliers.reshape(-1)] + [p.detach().reshape(-1) for p in tr.model.parameters()])
    flats = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(flats, flat)
    res["trainer_lambda_and_weights_bit_identical_across_ranks"] = bool(all(torch.equal(f, flats[0]) for f in flats))
    res["trainer_lambda_moved"] = bool((tr.lambda_multipliers != 1.0).any().item())
  except Exception as e:
    res["trainer_check_error"] = str(e)[:300]
  res["ok"] = bool(res["sums_ok"] and res.get("trainer_lambda_and_weights_bit_identical_across_ranks", False))
  return res


def strong_scaling_metric(args, spec, dev, world, rank, lam, exchange, kernel_ms_full):
  """north_star config 5 as written: ONE batch of 2^24 samples sharded over the N ranks (2^24 / N each), fwd+bwd + the
  all-reduce of the 16 sums every step.  At 2^21 samples per rank the kernel lasts ~27 us -- the order of a kernel launch
  from Python and of a small NCCL all-reduce -- so the steps are captured in a CUDA graph (no host in the loop) and the
  exchange happens inside the kernel over peer memory.  Device time, max over ranks."""
  import torch
  import torch.distributed as dist
  from ikl import fused, synthetic
  total = 1 << args.strong_total_log2
  Bs = total // world
  K = spec.num_constraints
  theta, q_des, obst = synthetic.make_batch(Bs, seed=7000 + rank, device=str(dev))
  th, tg, ob = synthetic.to_planes(theta, q_des, obst)
  del theta, q_des, obst
  dth = torch.empty_strided(th.shape, th.stride(), dtype=torch.float32, device=dev)
  out = torch.empty(K + 1, dtype=torch.float32, device=dev)
  S = 50                                   # steps per graph replay

  def timed_graph(body, reps=10, last_body=None):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for k in range(3):
        (last_body if (last_body is not None and k == 2) else body)()
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g, stream=side):
        for k in range(S):
          (last_body if (last_body is not None and k == S - 1) else body)()
    torch.cuda.current_stream(dev).wait_stream(side)
    g.replay()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (reps * S)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  local_ms = timed_graph(lambda: fused.lagrangian_forward_backward(spec, th, tg, ob, lam, dtheta=dth, out=out))
  res = {"total_batch": total, "batch_per_gpu": Bs, "steps_per_graph_replay": S, "kernel_only_ms": local_ms}
  if world == 1:
    step_ms = local_ms
    res["exchange"] = "none (single GPU)"
  elif exchange is not None:
    step_ms = timed_graph(lambda: fused.lagrangian_forward_backward(spec, th, tg, ob, lam, dtheta=dth, out=out, exchange=exchange))
    res["exchange"] = "inside the fused kernel: NVLink peer stores of the 16 fp64 sums + flag, last block pulls and adds"
    if args.exchange == "epoch":
      # the bench line's own form: a graph replay is one dual step of S batches; the sums stay local (float64 running sum) until
      # the S-th launch, which exchanges the running sums and moves lambda in the same kernel
      res["per_step_exchange_ms_per_step"] = step_ms
      acc = torch.zeros(K, dtype=torch.float64, device=dev)
      lam_dev = torch.tensor(lam, device=dev)
      closer = fused.DualStep(lam_dev, MULTIPLIER_LR, out=torch.empty_like(lam_dev), exchange=exchange)
      step_ms = timed_graph(lambda: fused.lagrangian_forward_backward(spec, th, tg, ob, lam, viol_accum=acc, dtheta=dth, out=out),
                            last_body=lambda: fused.lagrangian_forward_backward(spec, th, tg, ob, lam, viol_accum=acc, dtheta=dth, out=out,
                                                                                close_epoch=closer))
      res["exchange"] = ("once per dual step of %d batches: the closing launch exchanges the ranks' float64 running sums over NVLink peer memory "
                         "and moves lambda in the same kernel; `per_step_exchange_ms_per_step` = the 16 sums all-reduced inside EVERY launch" % S)
  else:
    def body():
      fused.lagrangian_forward_backward(spec, th, tg, ob, lam, dtheta=dth, out=out)
      dist.all_reduce(out)
    step_ms = timed_graph(body)
    res["exchange"] = "NCCL all_reduce of 17 floats after each kernel (captured in the graph)"
  if world > 1 and exchange is not None and args.compare_nccl:
    def body_nccl():
      fused.lagrangian_forward_backward(spec, th, tg, ob, lam, dtheta=dth, out=out)
      dist.all_reduce(out)
    res["nccl_variant_ms_per_step"] = timed_graph(body_nccl)
  ideal = kernel_ms_full * (total / float(1 << args.batch_log2)) / world
  res.update({"ms_per_step": step_ms, "samples_per_s": total / (step_ms * 1e-3), "exchange_overhead_us": (step_ms - local_ms) * 1e3,
              "ideal_ms": ideal, "efficiency_vs_ideal": ideal / step_ms,
              "frac_of_hbm_peak": ALGO_BYTES_PER_SAMPLE * Bs / (step_ms * 1e-3) / 1e9 / measured_peak()[0],
              "fixed_launch_cost_us": (local_ms - ideal) * 1e3})
  fixed_us, ex_us = res["fixed_launch_cost_us"], res["exchange_overhead_us"]
  res["limiter"] = ("%s: the part of a launch that does not shrink with the shard -- graph-node launch gap, pipeline ramp-up, tail and the "
                    "final reduction -- costs %.1f us per step, the in-step exchange %.1f us, against %.1f us of streaming at the HBM rate; "
                    "no host and no NCCL launch is in the loop" % (
                        "fixed per-launch cost" if fixed_us >= ex_us else "exchange latency + rank skew", fixed_us, ex_us, ideal * 1e3))
  return res


def grad_allreduce_metric(args, dev, world, peer):
  """The one exchange data-parallel training adds per step: the sum of the replicas' parameter gradients (83 203 floats for
  the H=100 network + the 16 sums).  Timed as 50 back-to-back calls inside one CUDA graph (device time, max over ranks):
  the peer-memory kernel of this repo (ikl_allreduce_sum_f32) against NCCL's all_reduce on the same buffer."""
  import torch
  import torch.distributed as dist
  n = 83203 + 16
  S = 50
  buf = torch.zeros(n, device=dev)

  def timed_graph(body, reps=10):
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      for _ in range(3):
        body()
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g, stream=side):
        for _ in range(S):
          body()
    torch.cuda.current_stream(dev).wait_stream(side)
    g.replay()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / (reps * S)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del g                                   # a captured NCCL collective must be gone before the process group is destroyed
    return float(t.item()) * 1e3
  res = {"elements": n, "calls_per_graph_replay": S}
  if peer is not None:
    res["peer_memory_kernel_us"] = timed_graph(lambda: peer.all_reduce_(buf))
  res["nccl_all_reduce_us"] = timed_graph(lambda: dist.all_reduce(buf))
  return res


def train_step_metric(args, spec, dev, world, lam, peer=None):
  """Secondary metric of BASELINE.json on the configuration SURVEY 8(d) names: full dual-learning train steps/s with 2^24
  samples per step over ALL ranks (2^24 / N per rank: strong scaling), micro-batched.  Per micro-batch: SimpleNetwork (H=100,
  depth 8, dropout 0.05, 83 203 parameters; cuBLAS) forward with the last layer transposed so theta comes out as planes,
  the fused TMA-staged Lagrangian kernel (sums + d/dtheta in one pass, no gradient rescaling pass), backward through the
  network accumulating into .grad; per step ONE flat all-reduce of [gradients | 16 violation sums] when N > 1, then Adam
  (lr 5e-5).  Synthetic on-device data.  The step is bounded by the network, which north_star keeps on cuBLAS:
  `fused_kernel_share` says how little of it is the hot path of this repo."""
  import torch
  import torch.distributed as dist
  from ikl import dist as ikl_dist, fused, synthetic
  from models.simple_network import SimpleNetwork
  rank = int(os.environ.get("RANK", "0"))
  total = 1 << args.train_total_log2
  per_rank = total // world
  micro = min(1 << args.train_micro_log2, per_rank)
  n_micro = per_rank // micro
  torch.manual_seed(1234)
  model = SimpleNetwork(20, 3, hidden_units=100, depth=8, dropout=0.05).to(dev)
  ikl_dist.broadcast_parameters(model)
  opt = torch.optim.Adam(model.parameters(), lr=5e-5, betas=(0.9, 0.999))
  theta0, q_des, obst = synthetic.make_batch(per_rank, seed=1000 + rank, device=str(dev))
  x = torch.zeros(per_rank, 20, device=dev)                 # kinematics/dataset.py:124-153 input template
  x[:, 0:2] = q_des[:, :2]
  x[:, 2], x[:, 3] = -synthetic.LIMIT, synthetic.LIMIT
  x[:, 4:] = obst.reshape(per_rank, 16)
  _, q_pl, ob_pl = synthetic.to_planes(theta0, q_des, obst)
  del theta0, q_des, obst
  lam_dev = torch.tensor(lam, device=dev)
  viol_sum = torch.zeros(spec.num_constraints, device=dev)

  def step():
    model.zero_grad()
    viol_sum.zero_()
    for m in range(n_micro):
      sl = slice(m * micro, (m + 1) * micro)
      theta = model.forward_planes(x[sl])
      viol, loss = fused.FusedIkLagrangian.apply(theta, q_pl[sl], ob_pl[sl], lam_dev, spec, lam, True)
      loss.backward()                                   # gradients of the sums add up over the micro-batches
      viol_sum.add_(viol.detach())
    ikl_dist.all_reduce_gradients(model, extra=viol_sum, peer=peer)
    opt.step()

  for _ in range(2):
    step()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n = max(2, min(args.steps, 4))
  e0.record()
  for _ in range(n):
    step()
  e1.record()
  torch.cuda.synchronize()
  t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms = float(t.item())
  # the fused kernel alone on one micro-batch of the same tensors
  with torch.no_grad():
    th = model.forward_planes(x[:micro]).detach()
  dth = torch.empty_strided(th.shape, th.stride(), dtype=torch.float32, device=dev)
  for _ in range(3):
    fused.lagrangian_forward_backward(spec, th, q_pl[:micro], ob_pl[:micro], lam, dtheta=dth)
  k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  k0.record()
  for _ in range(20):
    fused.lagrangian_forward_backward(spec, th, q_pl[:micro], ob_pl[:micro], lam, dtheta=dth)
  k1.record()
  torch.cuda.synchronize()
  kernel_ms = k0.elapsed_time(k1) / 20
  return {"steps_per_s": 1e3 / ms, "ms_per_step": ms, "samples_per_s": total / (ms * 1e-3), "samples_per_step": total,
          "batch_per_gpu": per_rank, "micro_batch": micro, "micro_batches_per_step": n_micro, "steps": n, "scaling": "strong",
          "fused_kernel_ms_per_micro_batch": kernel_ms, "fused_kernel_share": kernel_ms * n_micro / ms,
          "model": "SimpleNetwork(20->100x8->3, dropout 0.05, 83203 params) on cuBLAS + fused IK Lagrangian (planes, TMA) + Adam",
          "note": "bounded by the fp32 cuBLAS network (out of scope per north_star); the hot path of this repo is fused_kernel_share of the step",
          "exchange": ("one flat all-reduce of [83203 gradients | 16 violation sums] per step, %s" % (
              "one kernel over NVLink peer memory (ikl_allreduce_sum_f32)" if peer is not None else "NCCL")) if world > 1 else "none"}


def small_batch_train_metric(args, spec, dev, lam):
  """The reference's own regime (batch 8, scripts/options.py:15): train steps/s eager and as one CUDA graph per step,
  beside the reference-style CPU step (torch CPU network + oracle Lagrangian + autograd + Adam) on the host."""
  import torch
  from ikl import fused, synthetic
  from ikl.graph_step import GraphedTrainStep
  from models.simple_network import SimpleNetwork
  from oracle import ik_oracle as O
  from ikl import standard_config
  B = 8
  torch.manual_seed(7)
  theta0, q_des, obst = synthetic.make_batch(B, seed=77, device=str(dev))
  x = torch.zeros(B, 20, device=dev)
  x[:, 0:2] = q_des[:, :2]
  x[:, 2], x[:, 3] = -synthetic.LIMIT, synthetic.LIMIT
  x[:, 4:] = obst.reshape(B, 16)
  lam_dev = torch.tensor(lam, device=dev)
  inputs = {"input_tensor": x, "q_ee_desired": q_des, "dynamic_obstacles": obst}
  out = {"batch": B}

  def timed(fn, n):
    for _ in range(20):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
      fn()
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3)

  model = SimpleNetwork(20, 3, hidden_units=100, depth=8, dropout=0.05).to(dev)
  opt = torch.optim.Adam(model.parameters(), lr=5e-5)

  def eager():
    theta = model(x)
    viol, loss = fused.FusedIkLagrangian.apply(theta, q_des, obst, lam_dev, spec, lam)
    model.zero_grad()
    loss.backward()
    opt.step()
  out["eager_steps_per_s"] = timed(eager, 300)
  gmodel = SimpleNetwork(20, 3, hidden_units=100, depth=8, dropout=0.05).to(dev)
  gopt = torch.optim.Adam(gmodel.parameters(), lr=5e-5, capturable=True)
  g = GraphedTrainStep(gmodel, gopt, spec, B, dev)
  g.set_multipliers(lam_dev)
  out["graph_steps_per_s"] = timed(lambda: g.step(inputs), 2000)
  # the reference's way, on the host cores
  cmodel = SimpleNetwork(20, 3, hidden_units=100, depth=8, dropout=0.05)
  copt = torch.optim.Adam(cmodel.parameters(), lr=5e-5)
  ospec = O.spec_from_json(standard_config(CONFIG_NAME), 3, GOOD_SLOPE, BAD_SLOPE)
  cx, cq, cob, clam = x.cpu(), q_des.cpu(), obst.cpu(), torch.tensor(lam)

  def cpu_step():
    o = O.lagrangian(cmodel(cx), cq, clam, ospec, cob)
    cmodel.zero_grad()
    o["lagrange_loss"].backward()
    copt.step()
  for _ in range(10):
    cpu_step()
  t0 = time.perf_counter()
  for _ in range(100):
    cpu_step()
  out["cpu_reference_steps_per_s"] = 100 / (time.perf_counter() - t0)
  out["note"] = "batch 8 (the reference's default), SimpleNetwork H=100 depth 8, Adam; graph = one CUDA-graph replay per step"
  return out


def run_ours(args):
  import torch
  import torch.distributed as dist
  from ikl import ConstraintSpec, standard_config, fused, synthetic, _cabi
  from ikl.host_pipeline import HostLagrangianPipeline

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise RuntimeError("bench.py (impl ours) needs a CUDA device: the product path has no CPU fallback")
  numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else None
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
      os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line (NCCL prints its version banner there)
    opts = None
    try:
      opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:
      opts = None
    dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    # optionally keep SMs out of the persistent grid for the concurrent NCCL kernel (measured: not needed on B200, the
    # 64-byte all-reduce overlaps the next step's kernel either way: profiles/r01/bench_n2_*.json)
    _cabi.set_reserved_sms(args.reserve_sms)

  # The K sums are all-reduced INSIDE the fused kernel over NVLink peer memory (ikl.exchange.PeerExchange, include/ikl.h
  # IklExchange); --exchange nccl keeps the round-1 form (one asynchronous 64-byte NCCL all-reduce per step) for comparison.
  exchange, exchange_note = None, None
  if world > 1 and args.exchange in ("epoch", "peer"):
    ok = torch.ones(1, device=dev)
    try:
      from ikl.exchange import PeerExchange
      exchange = PeerExchange(device=dev)
    except Exception as e:
      ok.zero_()
      exchange_note = "peer exchange unavailable on this rank (%s)" % (str(e)[:160],)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) == 0.0:
      exchange, exchange_note = None, exchange_note or "peer exchange unavailable on another rank"

  spec = ConstraintSpec.from_config(standard_config(CONFIG_NAME), 3, GOOD_SLOPE, BAD_SLOPE)
  K = spec.num_constraints
  B = 1 << args.batch_log2
  lam = list(synthetic.SHIPPED_LAMBDA_OBS4_DYN)
  theta, q_des, obst = synthetic.make_batch(B, seed=rank, device=str(dev))
  if args.layout == "planes":
    th_in, tg_in, ob_in = synthetic.to_planes(theta, q_des, obst)
  else:
    th_in, tg_in, ob_in = theta, q_des, obst
  dtheta = torch.empty_strided(th_in.shape, th_in.stride(), dtype=torch.float32, device=dev)
  accum = torch.zeros(K, dtype=torch.float64, device=dev)
  lam_dev = torch.tensor(lam, device=dev)

  lam_next = torch.empty_like(lam_dev)
  # Default (--exchange epoch): the timed steps are the batches of ONE dual step (scripts/trainer.py:267-286 sums the K-vectors
  # of every batch and moves lambda once).  Every step adds its 16 sums into the rank's float64 running sum; the LAST step of
  # the region is launched through ikl_lagrangian_forward_backward_dual_step: that same kernel exchanges the running sums over
  # NVLink peer memory -- once per epoch, which is when the path needs global sums -- and writes the new lambda.  At N = 1 the
  # closing launch only fuses the lambda update.  --exchange peer: a per-step all-reduce inside every launch instead.
  epoch_mode = args.exchange == "epoch" and (world == 1 or exchange is not None)
  closer = fused.DualStep(lam_dev, MULTIPLIER_LR, out=lam_next, exchange=exchange) if epoch_mode else None

  pending = []

  def step(last=False):
    if epoch_mode:
      viol, loss, _ = fused.lagrangian_forward_backward(spec, th_in, tg_in, ob_in, lam, viol_accum=accum, dtheta=dtheta,
                                                        close_epoch=closer if last else None)
      return viol, loss
    if exchange is not None:          # ONE kernel: compute + all-reduce of the 16 sums over peer memory; viol is global
      viol, loss, _ = fused.lagrangian_forward_backward(spec, th_in, tg_in, ob_in, lam, viol_accum=accum, dtheta=dtheta, exchange=exchange)
      return viol, loss
    viol, loss, _ = fused.lagrangian_forward_backward(spec, th_in, tg_in, ob_in, lam, viol_accum=accum, dtheta=dtheta)
    if world > 1:
      # 16 floats: every rank sees the global sums that drive the multiplier update.  Asynchronous: NCCL's stream
      # waits for this step's kernel, the compute stream does not wait for NCCL until drain().
      pending.append((dist.all_reduce(viol, async_op=True), viol))
      if len(pending) > 64:
        pending.pop(0)[0].wait()
    return viol, loss

  def drain():
    while pending:
      pending.pop(0)[0].wait()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  sampler = ClockSampler(local_rank) if rank == 0 else None
  n_warm = max(args.warmup, 3)
  for i in range(n_warm):
    step(last=(i == n_warm - 1))   # the warm-up is an epoch of its own: the closing launch is warm too
  drain()
  accum.zero_()
  barrier()
  launches0 = _cabi.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t_wall0 = time.time()
  ev0.record()
  for i in range(args.steps):
    viol, loss = step(last=(i == args.steps - 1))
  drain()                          # the timed region ends when every all-reduce has completed
  ev1.record()
  barrier()
  t_wall1 = time.time()
  launches = _cabi.launch_count() - launches0
  epoch_check = None
  if epoch_mode:
    # what the timed region computed: the GLOBAL epoch sums in `accum` and the new lambda, identical on every rank
    local_epoch = torch.zeros(K, dtype=torch.float64, device=dev)
    one, _, _ = fused.lagrangian_forward_backward(spec, th_in, tg_in, ob_in, lam, viol_accum=local_epoch, dtheta=dtheta)
    parts = [local_epoch.clone()]
    if world > 1:
      parts = [torch.zeros_like(local_epoch) for _ in range(world)]
      dist.all_gather(parts, local_epoch)
    want = args.steps * torch.stack(parts).sum(0)
    rel = float(((accum - want).abs() / (args.steps * torch.stack(parts).abs().sum(0)).clamp(min=1e-30)).max())
    lam_want = fused.dual_update(lam_dev, accum, MULTIPLIER_LR)
    lams = [lam_next.clone()]
    if world > 1:
      lams = [torch.zeros_like(lam_next) for _ in range(world)]
      dist.all_gather(lams, lam_next)
    epoch_check = {"epoch_sums_rel_err_vs_steps_x_f64_sum_of_rank_sums": rel,
                   "lambda_equals_ikl_dual_update_of_the_epoch_sums": bool(torch.equal(lam_next, lam_want)),
                   "lambda_bit_identical_across_ranks": bool(all(torch.equal(l, lams[0]) for l in lams)),
                   "lambda_moved": bool((lam_next != lam_dev).any().item()), "multiplier_lr": MULTIPLIER_LR}
    epoch_check["ok"] = bool(rel <= 1e-9 and epoch_check["lambda_equals_ikl_dual_update_of_the_epoch_sums"] and
                             epoch_check["lambda_bit_identical_across_ranks"] and epoch_check["lambda_moved"])
  kernel_name = _cabi.last_kernel_name()
  ms_total = ev0.elapsed_time(ev1)
  t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms_total = float(t.item())
  ms_step = ms_total / args.steps
  value = world * B / (ms_step * 1e-3)

  # kernel-only duration for the roofline: the same launches back to back with nothing else on the stream
  kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n_k = max(args.steps, 20)
  kev0.record()
  for _ in range(n_k):
    fused.lagrangian_forward_backward(spec, th_in, tg_in, ob_in, lam, dtheta=dtheta)
  kev1.record()
  torch.cuda.synchronize()
  kernel_ms = kev0.elapsed_time(kev1) / n_k
  if world == 1:
    kernel_ms = ms_step            # at N=1 the timed region IS K back-to-back launches of this one kernel
  # keep the GPU under the same load long enough for nvidia-smi to see it (a 20-step region lasts milliseconds)
  t_load0 = time.time()
  while rank == 0 and time.time() - t_load0 < 1.2:
    for _ in range(200):
      fused.lagrangian_forward_backward(spec, th_in, tg_in, ob_in, lam, dtheta=dtheta)
    torch.cuda.synchronize()
  t_load1 = time.time()
  clocks = sampler.stop(t_wall0, t_load1) if sampler else None

  line = None
  if rank == 0:
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_SAMPLE * B / (kernel_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
      try:
        with open(prof) as f:
          traffic = json.load(f).get(args.layout, {}).get("dram_bytes_per_launch")
      except Exception:
        traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, world, "epoch" if epoch_mode else ("peer" if exchange is not None else "nccl")),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel": kernel_name, "kernel_ms": kernel_ms, "algorithmic_bytes_per_sample": ALGO_BYTES_PER_SAMPLE,
                     "peak_source": peak_src},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "native_lib": _cabi.loaded_path(),
    }
    if epoch_check is not None:
      line["dual_step_check"] = epoch_check

  # ---- context numbers (rank 0, N = 1): the reference's own tensor layout, and the reference's eager path on this GPU
  extras = {}
  if rank == 0 and world == 1 and not args.skip_context:
    dth_rows = torch.empty_like(theta)
    for _ in range(3):
      fused.lagrangian_forward_backward(spec, theta, q_des, obst, lam, dtheta=dth_rows)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(20):
      fused.lagrangian_forward_backward(spec, theta, q_des, obst, lam, dtheta=dth_rows)
    r1.record()
    torch.cuda.synchronize()
    rows_ms = r0.elapsed_time(r1) / 20
    extras["rows_layout"] = {"kernel": _cabi.last_kernel_name(), "kernel_ms": rows_ms, "samples_per_s": B / (rows_ms * 1e-3),
                             "frac_of_hbm_peak_algorithmic": ALGO_BYTES_PER_SAMPLE * B / (rows_ms * 1e-3) / 1e9 / measured_peak()[0],
                             "note": "theta (B,3), q_ee_desired (B,3), dynamic_obstacles (B,4,4) exactly as the reference's DataLoader "
                                     "hands them: 100 physical bytes per sample"}
    del dth_rows
    # the launch of a TRAINING step (gradient only: scripts/trainer.py:262-265 consumes nothing but lagrange_loss.backward()),
    # on the benchmark config and on the 4-static-obstacle config (32 B/sample: bound by the FP32 / ALU pipes, not by HBM)
    try:
      def timed20(fn):
        for _ in range(3):
          fn()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(20):
          fn()
        g1.record()
        torch.cuda.synchronize()
        return g0.elapsed_time(g1) / 20
      go = {}
      ms = timed20(lambda: fused.lagrangian_backward(spec, th_in, tg_in, ob_in, lam, dtheta=dtheta))
      go["4obs_dynamic"] = {"kernel": _cabi.last_kernel_name(), "kernel_ms": ms, "with_sums_kernel_ms": kernel_ms,
                            "frac_of_hbm_peak": ALGO_BYTES_PER_SAMPLE * B / (ms * 1e-3) / 1e9 / measured_peak()[0]}
      spec4 = ConstraintSpec.from_config(standard_config("cfg_joint_limits_and_4obs_static"), 3, GOOD_SLOPE, BAD_SLOPE)
      ms_full = timed20(lambda: fused.lagrangian_forward_backward(spec4, th_in, tg_in, None, lam, dtheta=dtheta))
      ms = timed20(lambda: fused.lagrangian_backward(spec4, th_in, tg_in, None, lam, dtheta=dtheta))
      go["4obs_static"] = {"kernel": _cabi.last_kernel_name(), "kernel_ms": ms, "with_sums_kernel_ms": ms_full,
                           "frac_of_hbm_peak": 32 * B / (ms * 1e-3) / 1e9 / measured_peak()[0], "algorithmic_bytes_per_sample": 32}
      extras["gradient_only_launch"] = go
    except Exception as e:
      extras["gradient_only_launch"] = {"unavailable": str(e)[:200]}
    try:
      from oracle import ik_oracle as O
      Bg = 1 << 22
      ospec = O.spec_from_json(standard_config(CONFIG_NAME), 3, GOOD_SLOPE, BAD_SLOPE)
      th_g, qd_g, ob_g = theta[:Bg].clone(), q_des[:Bg].clone(), obst[:Bg].clone()
      O.lagrangian_with_grad(th_g, qd_g, lam_dev, ospec, ob_g)
      torch.cuda.synchronize()
      g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      g0.record()
      for _ in range(3):
        O.lagrangian_with_grad(th_g, qd_g, lam_dev, ospec, ob_g)
      g1.record()
      torch.cuda.synchronize()
      eager_ms = g0.elapsed_time(g1) / 3
      extras["reference_eager_cuda"] = {"samples_per_s": Bg / (eager_ms * 1e-3), "ms_per_step": eager_ms, "batch": Bg,
                                        "note": "the reference's op-by-op PyTorch path (oracle port, ~960 eager kernels fwd+bwd) on this same GPU"}
      del th_g, qd_g, ob_g
      torch.cuda.empty_cache()
    except Exception as e:
      extras["reference_eager_cuda"] = {"unavailable": str(e)[:200]}

    # level (i) of the boundary, timed: the reference's UNMODIFIED scripts/trainer.py (baseline/_ref, loaded by file path) with
    # kinematics.* resolving to this repo's function-level CUDA drop-ins -- what a user gets with zero code changes
    try:
      extras["reference_trainer_over_dropins"] = reference_trainer_over_dropins(theta, q_des, obst, lam_dev, dev)
    except Exception as e:
      extras["reference_trainer_over_dropins"] = {"unavailable": str(e)[:200]}

  # ---- end to end through the host-buffer API (rank-local; every rank does the same work) -------------------
  e2e = None
  if not args.skip_e2e:
    del th_in, tg_in, ob_in, dtheta
    Be = 1 << args.e2e_batch_log2
    h_theta = theta[:Be].cpu().pin_memory()
    h_q = q_des[:Be].cpu().pin_memory()
    h_ob = obst[:Be].cpu().pin_memory()
    h_dth = torch.empty(Be, 3).pin_memory()
    del theta, q_des, obst
    pipe = HostLagrangianPipeline(spec, chunk=1 << args.e2e_chunk_log2, device=dev)
    for _ in range(2):
      pipe.run(h_theta, h_q, h_ob, lam, h_dth)
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e = max(3, min(args.steps, 10))
    ee0.record()
    for _ in range(n_e):
      viol_h, loss_h = pipe.run(h_theta, h_q, h_ob, lam, h_dth)       # returns once dtheta and the sums are on the host
    ee1.record()                                                       # the current stream has waited on the pipeline's streams
    torch.cuda.synchronize()
    dt = ee0.elapsed_time(ee1) * 1e-3 / n_e
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    e2e = {"value": world * Be / dt, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
           "ms_per_step": dt * 1e3, "batch_per_gpu": Be, "steps": n_e,
           "pcie_gbs": {"h2d": pipe.h2d_bytes / dt / 1e9, "d2h": pipe.d2h_bytes / dt / 1e9,
                        "note": "both directions overlap; the host-to-device link is the bound of this arm, not the kernels"},
           "api": "ikl.host_pipeline.HostLagrangianPipeline.run (pinned host rows-layout tensors in, host dtheta + sums out)",
           "numa_bind": numa}
    # the same call on COMPACT host tensors -- q_ee_desired (B,2), obstacles (B,4,3): only the 68 bytes per sample the path
    # reads cross PCIe instead of the reference's 88 (unused phi column and pad column)
    try:
      h_q2 = h_q[:, :2].contiguous().pin_memory()
      h_ob3 = h_ob[:, :, :3].contiguous().pin_memory()
      for _ in range(2):
        pipe.run(h_theta, h_q2, h_ob3, lam, h_dth)
      barrier()
      c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      c0.record()
      for _ in range(n_e):
        pipe.run(h_theta, h_q2, h_ob3, lam, h_dth)
      c1.record()
      torch.cuda.synchronize()
      dtc = c0.elapsed_time(c1) * 1e-3 / n_e
      tc = torch.tensor([dtc], dtype=torch.float64, device=dev)
      if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
      dtc = float(tc.item())
      e2e["compact_inputs"] = {"value": world * Be / dtc, "ms_per_step": dtc * 1e3, "h2d_bytes_per_step": int(pipe.h2d_bytes),
                               "d2h_bytes_per_step": int(pipe.d2h_bytes), "pcie_h2d_gbs": pipe.h2d_bytes / dtc / 1e9,
                               "note": "q_ee_desired (B,2) and obstacles (B,4,3) on the host: 68 B/sample in instead of 88"}
      del h_q2, h_ob3
    except Exception as ex_c:
      e2e["compact_inputs"] = {"unavailable": str(ex_c)[:200]}

  multi_check = None
  if world > 1 and not args.skip_multi_check:
    multi_check = multi_gpu_check(args, spec, dev, world, rank, lam, exchange)
    if exchange_note:
      multi_check["exchange_note"] = exchange_note
  strong = None
  if not args.skip_strong:
    strong = strong_scaling_metric(args, spec, dev, world, rank, lam, exchange, kernel_ms_full=kernel_ms)

  # gradient all-reduce of the data-parallel train step: one kernel over NVLink peer memory (--exchange nccl: NCCL)
  peer_ar, grad_ar = None, None
  if world > 1 and args.exchange in ("epoch", "peer") and exchange is not None:
    ok = torch.ones(1, device=dev)
    try:
      from ikl.exchange import PeerAllReduce
      peer_ar = PeerAllReduce(83203 + 16, device=dev)
    except Exception as e:
      ok.zero_()
      exchange_note = (exchange_note or "") + " peer all-reduce unavailable (%s)" % (str(e)[:120],)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok.item()) == 0.0:
      peer_ar = None
  if world > 1 and not args.skip_strong:
    grad_ar = grad_allreduce_metric(args, dev, world, peer_ar)

  train = None
  if not args.skip_train_step:
    train = train_step_metric(args, spec, dev, world, lam, peer=peer_ar)

  small = None
  if not args.skip_train_step and rank == 0 and world == 1:
    small = small_batch_train_metric(args, spec, dev, lam)

  if rank == 0:
    line["e2e"] = e2e
    line["parity"] = parity_summary()
    line["train_step"] = train
    line["multi_gpu_check"] = multi_check
    line["strong_scaling"] = strong
    line["grad_allreduce"] = grad_ar
    line["train_step_batch8"] = small
    line["context"] = extras
    if not args.skip_cpu and world == 1:
      line["cpu_baseline"] = cpu_baseline_leg(args)
      try:
        line["cpu_baseline_c"] = cpu_c_oracle_samples_per_s(args.cpu_batch_log2 + 1)
        if e2e is not None:         # what the headline e2e ratio would be against a tuned CPU port of the same maths
          e2e["e2e_vs_cpu_baseline_c"] = e2e["value"] / line["cpu_baseline_c"]["value"]
      except Exception as e:          # no C compiler on the box: the PyTorch port above is the baseline of record
        line["cpu_baseline_c"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=10)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--batch-log2", type=int, default=24)
  ap.add_argument("--layout", default="planes", choices=["planes", "rows"])
  ap.add_argument("--e2e-batch-log2", type=int, default=24)
  ap.add_argument("--e2e-chunk-log2", type=int, default=21)
  ap.add_argument("--cpu-batch-log2", type=int, default=21)
  ap.add_argument("--skip-train-step", action="store_true")
  ap.add_argument("--reserve-sms", type=int, default=0, help="N > 1 only: SMs left out of the persistent grid for the concurrent NCCL kernel")
  ap.add_argument("--no-numa-bind", action="store_true")
  ap.add_argument("--skip-context", action="store_true")
  ap.add_argument("--skip-e2e", action="store_true")
  ap.add_argument("--skip-cpu", action="store_true")
  ap.add_argument("--skip-multi-check", action="store_true")
  ap.add_argument("--skip-strong", action="store_true")
  ap.add_argument("--exchange", default="epoch", choices=["epoch", "peer", "nccl"],
                  help="when the 16 sums become global: epoch = once per dual step, inside the closing launch, over NVLink peer memory "
                       "(what update_multipliers needs); peer = all-reduced inside EVERY launch; nccl = asynchronous NCCL all-reduce per step")
  ap.add_argument("--strong-total-log2", type=int, default=24, help="total batch of the strong-scaling measurement (sharded over the ranks)")
  ap.add_argument("--compare-nccl", action="store_true", help="strong scaling: also time kernel + NCCL all-reduce in a CUDA graph")
  ap.add_argument("--train-total-log2", type=int, default=24, help="samples per train step over ALL ranks (SURVEY 8d), micro-batched")
  ap.add_argument("--train-micro-log2", type=int, default=20, help="micro-batch per GPU inside a train step")
  ap.add_argument("--force-port", action="store_true", help="--impl reference: time the oracle port even when baseline/_ref is installed")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
