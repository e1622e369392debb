"""Checkpoint / device helpers with the reference's names and on-disk formats (utils/training_utils.py:5-77).

Layout on disk (unchanged): <folder>/weights_<epoch>/{model.pth, adam.pth, multipliers.pth}; model.pth and adam.pth are
state_dicts, multipliers.pth is the raw (K,) float32 tensor.  Additions over the reference: load_multipliers() (the
reference never reloads lambda, so a resumed run restarts at initial_lambda) and map_location on every load so
checkpoints written on a GPU open on a CPU-only host.
"""
import os

import torch


def _weights_dir(folder, epoch, create=False):
  path = os.path.join(folder, "weights_{}".format(epoch))
  if create:
    os.makedirs(path, exist_ok=True)
  return path


def save_model(model, adam, folder, epoch):
  """Write model (and, if given, optimizer) state; returns (model_path, adam_path or None)."""
  d = _weights_dir(folder, epoch, create=True)
  model_path = os.path.join(d, "model.pth")
  torch.save(model.state_dict(), model_path)
  adam_path = None
  if adam is not None:
    adam_path = os.path.join(d, "adam.pth")
    torch.save(adam.state_dict(), adam_path)
  return model_path, adam_path


def load_model(model, adam, model_path, adam_path, map_location=None):
  """Load whatever keys of the checkpoint the current model knows (unknown keys are dropped, as in the reference)."""
  current = model.state_dict()
  stored = torch.load(model_path, map_location=map_location)
  current.update({k: v for k, v in stored.items() if k in current})
  model.load_state_dict(current)
  if adam is not None:
    state = adam.state_dict()
    state.update(torch.load(adam_path, map_location=map_location))
    adam.load_state_dict(state)
  return model, adam


def save_multipliers(lambda_multipliers, folder, epoch):
  """torch.save the (K,) multiplier tensor as weights_<epoch>/multipliers.pth; returns the path."""
  path = os.path.join(_weights_dir(folder, epoch, create=True), "multipliers.pth")
  torch.save(lambda_multipliers, path)
  return path


def load_multipliers(folder, epoch=None, map_location="cpu"):
  """Read multipliers.pth from `folder` (a weights_<epoch> directory, or its parent plus `epoch`)."""
  d = folder if epoch is None else _weights_dir(folder, epoch)
  return torch.load(os.path.join(d, "multipliers.pth"), map_location=map_location)


def count_parameters(model):
  """Number of trainable parameters."""
  return sum(p.numel() for p in model.parameters() if p.requires_grad)


def get_best_system_device():
  """"cuda" if present, else "mps", else "cpu" (the IK kernels themselves need "cuda")."""
  if torch.cuda.is_available():
    return "cuda"
  if torch.backends.mps.is_available():
    return "mps"
  return "cpu"
