"""GPU checks that round 1 could only write, first run on a B200 in round 2 (profiles/r2_gpu_validate_unverified.txt):
* the RL bit search (reference: learners/uniform_quantization/bit_optimizer.py:137-327) and the pruning-ratio search
  (learners/weight_sparsification/pr_optimizer.py:411-611) through the REAL learners' compiled step;
* --enbl_device_preprocess end to end (utils/external/imagenet_preprocessing.py:225-260 on the device): the image
  placeholder equals the host pipeline's batch bit for bit.
The tool scripts hold the checks (they are also runnable stand-alone under gpurun); a failed assert fails the test."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'tools', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rl_bit_search_through_the_real_learner():
    _tool('rl_smoke').main()


def test_pruning_ratio_search_through_the_real_learner():
    _tool('rl_smoke').ws_main()


def test_device_preprocess_feeds_the_same_batches_as_the_host_pipeline():
    pytest.importorskip('PIL')
    _tool('preproc_e2e_check').main()
