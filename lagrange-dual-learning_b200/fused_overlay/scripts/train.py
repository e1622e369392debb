"""python lagrange-dual-learning_b200/fused_overlay/scripts/train.py --model_name ... --config_file ...   (reference: scripts/train.py:6-9)

Under torchrun (WORLD_SIZE > 1) every process binds to its LOCAL_RANK GPU and joins an NCCL group before the trainer is
built; the trainer then shards each batch over the ranks (ikl/dist.py)."""
import os
import sys

_OVERLAY = os.path.abspath(os.path.join(os.path.dirname(os.path.realpath(__file__)), ".."))
for _p in (os.path.dirname(_OVERLAY), _OVERLAY):        # package root (ikl, kinematics, utils, models), then this overlay (scripts)
  if _p not in sys.path:
    sys.path.insert(0, _p)

from scripts.options import Options  # noqa: E402


def main(argv=None):
  opt = Options().parse(argv)
  if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
  from scripts.trainer import IkLagrangeDualTrainer
  IkLagrangeDualTrainer(opt).main()


if __name__ == "__main__":
  main()
