"""Two eager training steps of a bench workload (no CUDA graph) — the target of the ncu launch list
(tools/gpu_launchlist.sh).  usage: python tools/one_step.py [workload] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.cuda.set_device(0)
    lrn = bench.build_learner(workload, 1)
    lrn.iterator_train.prefill()
    for _ in range(steps):
        lrn.train_step()
    torch.cuda.synchronize()
    print('losses', lrn.sess_train.fetch_losses())


if __name__ == '__main__':
    main()
