"""Repository-relative folders (reference: utils/paths.py:4-15)."""
import os

_TOP = os.path.abspath(os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))


def top_folder(rel=''):
  """Absolute path of `rel` under the top of this tree."""
  return os.path.join(_TOP, rel)


def data_folder(rel=''):
  return os.path.join(top_folder('data'), rel)


def models_folder(rel=''):
  return os.path.join(top_folder('models'), rel)
