"""Execution script: `python -m pocketflow_b200.nets.mobilenet_at_ilsvrc12_run --learner uniform --uql_weight_bits 8 ...`
(/root/reference/nets/mobilenet_at_ilsvrc12_run.py)."""
import sys

from .mobilenet_at_ilsvrc12 import ModelHelper
from .run_utils import run

if __name__ == '__main__':
    sys.exit(run(ModelHelper))
