"""Execution script: `python -m pocketflow_b200.nets.resnet_at_ilsvrc12_run --learner uniform --uql_weight_bits 8 ...`
(/root/reference/nets/resnet_at_ilsvrc12_run.py)."""
import sys

from .resnet_at_ilsvrc12 import ModelHelper
from .run_utils import run

if __name__ == '__main__':
    sys.exit(run(ModelHelper))
