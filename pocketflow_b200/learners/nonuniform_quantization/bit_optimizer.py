"""Bit allocation for the NonUniformQuantLearner (/root/reference/learners/nonuniform_quantization/bit_optimizer.py).

Without the RL agent every layer gets the flag values (:135-142) — that is what this build provides.  The roll-out
search is the uniform learner's loop with `nuql_*` flags upstream; it is not wired here because a new bit-width changes
the SIZE of a layer's codebook, so every roll-out would have to re-run the quantile initialisation on the restored
weights before fine-tuning (a device-side step that needs a GPU to validate).  The flags are declared so that the
reference's command lines parse."""
from ...flags import FLAGS, DEFINE_boolean, DEFINE_integer, DEFINE_string

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


class BitOptimizer(object):
    def __init__(self, nb_matmuls, nb_activations):
        self.nb_matmuls, self.nb_activations = nb_matmuls, nb_activations

    def run(self):
        if FLAGS.nuql_enbl_rl_agent:
            raise NotImplementedError('--nuql_enbl_rl_agent: the RL bit search is built for the uniform learner only '
                                      '(codebooks would have to be re-initialised per roll-out)')
        return [FLAGS.nuql_weight_bits] * self.nb_matmuls, [FLAGS.nuql_activation_bits] * self.nb_activations
