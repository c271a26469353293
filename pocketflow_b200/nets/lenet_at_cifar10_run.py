"""Execution script: `python -m pocketflow_b200.nets.lenet_at_cifar10_run --learner uniform --uql_weight_bits 8 ...`
(/root/reference/nets/lenet_at_cifar10_run.py)."""
import sys

from .lenet_at_cifar10 import ModelHelper
from .run_utils import run

if __name__ == '__main__':
    sys.exit(run(ModelHelper))
