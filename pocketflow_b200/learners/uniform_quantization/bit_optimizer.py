"""Bit allocation for the UniformQuantLearner (/root/reference/learners/uniform_quantization/bit_optimizer.py).
Without the RL agent the reference returns the flag values for every layer (:128-135); the DDPG
roll-out search around the step is a "next" row."""
from ...flags import FLAGS, DEFINE_boolean, DEFINE_float

DEFINE_boolean('uql_enbl_rl_agent', False, 'enable the RL agent for bit allocation')
DEFINE_boolean('uql_enbl_rl_layerwise_tune', False, 'layerwise fine-tuning inside RL roll-outs')
DEFINE_float('uql_equivalent_bits', 4, 'equivalent # of bits for the RL agent')


class BitOptimizer(object):
    def __init__(self, dataset_name, weights, statistics, *unused):
        self.nb_matmuls = statistics['nb_matmuls']
        self.nb_activations = statistics['nb_activations']

    def run(self):
        if FLAGS.uql_enbl_rl_agent:
            raise NotImplementedError('RL bit allocation (DDPG roll-outs) is not built yet; run without '
                                      '--uql_enbl_rl_agent')
        return [FLAGS.uql_weight_bits] * self.nb_matmuls, [FLAGS.uql_activation_bits] * self.nb_activations
