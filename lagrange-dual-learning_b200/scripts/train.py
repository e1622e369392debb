"""python scripts/train.py --model_name ... --config_file ...   (reference: scripts/train.py:6-9)"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.realpath(__file__)), "..")))

from scripts.options import Options  # noqa: E402
from scripts.trainer import IkLagrangeDualTrainer  # noqa: E402

if __name__ == "__main__":
  IkLagrangeDualTrainer(Options().parse()).main()
