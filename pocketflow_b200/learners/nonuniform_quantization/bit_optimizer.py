"""Bit allocation for the NonUniformQuantLearner (/root/reference/learners/nonuniform_quantization/bit_optimizer.py).

Without the RL agent every layer gets the flag values (:135-142).  With `--nuql_enbl_rl_agent` the search is the uniform
learner's roll-out loop on the `nuql_*` flags (the two reference files differ in flag names, `tune_global_steps` and
one initialisation op): the DDPG agent proposes per-layer bit-widths under the `nuql_equivalent_bits` budget, every
roll-out restores the pre-trained weights, sets the bit-widths — a new bit-width changes the SIZE of a layer's
codebook, so the codebooks are re-fitted to the restored weights by the quantile initialisation, as the reference's
`cluster_init` does after its restore (:200-206) — fine-tunes and scores on the evaluation split."""
from ...flags import DEFINE_boolean, DEFINE_integer, DEFINE_string
from ..uniform_quantization.bit_optimizer import BitOptimizer as UniformBitOptimizer

DEFINE_integer('nuql_equivalent_bits', 4, 'average number of bits per weight the RL search may spend')
DEFINE_integer('nuql_nb_rlouts', 200, 'number of roll-outs of the RL search')
DEFINE_integer('nuql_w_bit_min', 2, 'smallest bit-width a layer may get')
DEFINE_integer('nuql_w_bit_max', 8, 'largest bit-width a layer may get')
DEFINE_integer('nuql_tune_layerwise_steps', 100, 'layer-wise fine-tuning steps inside a roll-out')
DEFINE_integer('nuql_tune_global_steps', 2101, 'global fine-tuning steps inside a roll-out')
DEFINE_string('nuql_tune_save_path', './rl_tune_models/model.ckpt', 'where roll-outs save the tuned model')
DEFINE_integer('nuql_tune_disp_steps', 300, 'progress-line interval inside a roll-out')
DEFINE_boolean('nuql_enbl_random_layers', True, 'visit the layers in a random order in every roll-out')
DEFINE_boolean('nuql_enbl_rl_agent', False, 'search the per-layer bit-widths with the RL agent')
DEFINE_boolean('nuql_enbl_rl_global_tune', True, 'fine-tune all layers before a roll-out is scored')
DEFINE_boolean('nuql_enbl_rl_layerwise_tune', False, 'fine-tune layer by layer before a roll-out is scored')


class BitOptimizer(UniformBitOptimizer):
    PREFIX = 'nuql'
