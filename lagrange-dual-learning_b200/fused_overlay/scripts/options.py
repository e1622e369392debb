"""Command-line options of the trainer: the reference's 23 flags with the same names and defaults
(scripts/options.py:9-35), plus --distributed / --unfused for this implementation."""
import argparse

from utils.paths import top_folder

_FLAGS = [
    # name, type, default, help
    ("model_name", str, None, "The name of this experiment/model"),
    ("log_dir", str, top_folder("training_logs"), "Where to save models and tensorboard events"),
    ("lagrange_iters", int, 1000, "Number of lagrange dual epochs"),
    ("train_iters", int, 500, "Number of training iterations for each relaxation"),
    ("batch_size", int, 8, "Use a small batch size i.e 8"),
    ("train_dataset_size", int, 32000, "Number of training examples to generate"),
    ("val_dataset_size", int, 4000, "Number of validation examples to generate"),
    ("multiplier_lr", float, 1e-4, "Update LR for the multipliers"),
    ("optimizer_lr", float, 1e-4, "Learning rate for Adam"),
    ("initial_lambda", float, 40, "Initial value for the Lagrange multipliers"),
    ("config_file", str, None, "Path to the configuration file that specifies constraints"),
    ("num_workers", int, 4, "Number of data workers to fetch training samples"),
    ("model_save_hz", int, 200, "Save the model after this many epochs"),
    ("load_weights_folder", str, None, "Path containing a model.pth"),
    ("hidden_units", int, 40, "The number of hidden units for each network layer"),
    ("cache_save_path", str, None, "If given, load/save the dataset to the specified path for fast loading in the future"),
    ("penalty_bad_slope", float, 1.0, "Slope for piecewise linear constraint penalties."),
    ("penalty_good_slope", float, 0.1, "Slope for piecewise linear constraint penalties."),
]


class Options(object):
  def __init__(self):
    self.parser = argparse.ArgumentParser(description="Options for the IkLagrangeDualTrainer")
    for name, typ, default, text in _FLAGS:
      self.parser.add_argument("--" + name, type=typ, default=default, help=text)
    self.parser.add_argument("--num_links", type=int, default=3, choices=[3, 8], help="Number of links in the robotic arm")
    self.parser.add_argument("--load_adam", action="store_true", default=False, help="Should the adam state be loaded?")
    # Visualization / evaluation arguments (accepted for command-line compatibility).
    self.parser.add_argument("--show_groundtruth_theta", action="store_true", default=False,
                             help="If True, visualize the joint angles from the dataset instead of the network prediction")
    self.parser.add_argument("--no_plot", action="store_true", default=False, help="If True, skipping plotting during evaluation")
    # Additions of the B200 implementation.
    self.parser.add_argument("--unfused", action="store_true", default=False,
                             help="process_batch goes op by op through the function-level CUDA drop-ins instead of the fused kernel")
    self.parser.add_argument("--device_data", action="store_true", default=False,
                             help="Generate the datasets on the GPU and batch them there (for large batch sizes)")
    self.parser.add_argument("--cuda_graph", action="store_true", default=False,
                             help="Replay each train step (network, fused Lagrangian, backward, Adam) as one CUDA graph")
    self.parser.add_argument("--val_batch_size", type=int, default=0,
                             help="Batch size of the validation loader (dual step and validation pass); 0 = --batch_size as in "
                                  "the reference.  The sums are additive, so a few large batches give the same dual step "
                                  "(up to the dropout draws of the reference's train-mode validation) in far fewer launches")
    self.parser.add_argument("--train_sums", action="store_true", default=False,
                             help="Also form the constraint sums and the Lagrangian's value in every TRAINING step.  The reference's "
                                  "training loop consumes only lagrange_loss.backward() (trainer.py:262-265), so by default those "
                                  "launches are the gradient-only kernel; dual steps and validation always compute the sums")
    self.parser.add_argument("--load_multipliers", action="store_true", default=False,
                             help="Also reload multipliers.pth from --load_weights_folder (the reference restarts lambda)")

  def parse(self, argv=None):
    self.options = self.parser.parse_args(argv)
    return self.options

  def parse_default(self, config_file=None):
    """A small smoke configuration (the reference hard-codes its author's home directory here)."""
    self.options = self.parser.parse_args(
        ["--model_name", "default", "--config_file", config_file or "", "--train_dataset_size", "1000",
         "--val_dataset_size", "100", "--train_iters", "10"])
    return self.options
