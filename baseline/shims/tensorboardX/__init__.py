"""Test shim for `tensorboardX` (the reference's logging dependency, Pipfile; not installed in this image).

scripts/trainer.py:18,116 only needs `SummaryWriter(logdir=...)` with add_scalar / add_scalars / close.  The shim forwards to
torch's bundled writer when tensorboard is importable and swallows the calls otherwise.  It is on sys.path ONLY in the tests
that execute the reference's own trainer (tests/ref_harness.py); a real installation has the real package."""


class SummaryWriter(object):
  def __init__(self, logdir=None, **kwargs):
    self._w = None
    try:
      from torch.utils.tensorboard import SummaryWriter as _TorchWriter
      self._w = _TorchWriter(log_dir=logdir)
    except Exception:
      self._w = None

  def add_scalar(self, *a, **k):
    if self._w is not None:
      self._w.add_scalar(*a, **k)

  def add_scalars(self, *a, **k):
    if self._w is not None:
      self._w.add_scalars(*a, **k)

  def close(self):
    if self._w is not None:
      self._w.close()
