"""Roll-out bookkeeping of the bit-allocation search (behaviour of
/root/reference/learners/uniform_quantization/rl_helper.py:26-122): the state vector the agent sees for every layer
and the projection of its raw action onto the bit-widths the remaining budget still allows.

State of layer i (length L + 6 for L layers): one-hot(i) | kernel shape as 4 numbers (a dense [in, out] kernel is
read as [1, 1, in, out]) | #weights(i) / max #weights | #weights of the layers after i / total #weights.
Every number is float64 numpy, as in the reference (the pinned tests compare exactly)."""
import random

import numpy as np

from ...flags import FLAGS


class RLHelper(object):
    # pylint: disable=too-many-instance-attributes
    def __init__(self, total_bits, num_weights, var_shapes, random_layers=False, flag_prefix='uql'):
        """total_bits: budget = sum over layers of bits x #weights; num_weights: per layer; var_shapes: kernel shapes
        (rank 2 or 4) in layer order; random_layers: visit the layers in a fresh random order every roll-out;
        flag_prefix: 'uql' / 'nuql' — the non-uniform learner's helper is the same class on its own flags
        (learners/nonuniform_quantization/rl_helper.py is the uniform one with the flags renamed)."""
        self.flag_prefix = flag_prefix
        self.num_weights = num_weights
        self.nb_vars = len(num_weights)
        self.total_bits = total_bits
        self.total_num_weights = sum(num_weights)
        self.random_layers = random_layers
        self.layer_idxs = list(range(self.nb_vars))
        self.s_dims = self.nb_vars + 6
        self.var_shapes = [self._as_conv_shape(s) for s in var_shapes]
        later = [np.sum(num_weights[i + 1:]) / self.total_num_weights for i in range(self.nb_vars)]
        largest = np.max(num_weights)
        self.states = np.zeros((self.nb_vars, self.s_dims))
        self.states[:, :self.nb_vars] = np.eye(self.nb_vars)
        for i in range(self.nb_vars):
            self.states[i, self.nb_vars:] = np.hstack((self.var_shapes[i], [num_weights[i] / largest, later[i]]))
        self.reset_budget()

    @staticmethod
    def _as_conv_shape(shape):
        assert len(shape) in [2, 4], 'Unknown weight shape. Must be a 2 (fc) or 4 (conv) dimensional.'
        shape = np.asarray(shape, np.float64)
        return shape if shape.size == 4 else np.hstack((np.ones(2), shape))

    def reset_budget(self):
        self.w_bits_used = 0
        self.quantized_layers = 0
        self.num_weights_to_quantize = self.total_num_weights

    def reset(self):
        """Start of a roll-out: nothing spent yet; optionally a new visiting order."""
        self.reset_budget()
        if self.random_layers:
            random.shuffle(self.layer_idxs)

    def calc_state(self, idx):
        return self.states[idx:idx + 1].copy()

    def calc_reward(self, accuracy):
        return accuracy * np.ones((1, 1))

    def calc_w(self, action, idx):
        """Bit-width (array of shape (1, 1)) for layer `idx` given the actor's output in [0, w_bit_max - w_bit_min].
        All but the last visited layer: round, add the minimum, and cap at what leaves every unvisited layer its
        minimum.  The last visited layer takes the whole remainder.  Both are capped at the maximum."""
        lo, hi = getattr(FLAGS, self.flag_prefix + '_w_bit_min'), getattr(FLAGS, self.flag_prefix + '_w_bit_max')
        n_here = self.num_weights[idx]
        spare = self.total_bits - self.w_bits_used - self.num_weights_to_quantize * lo
        assert spare >= 0, 'Not enough budget for layer {}'.format(idx)
        is_last = self.quantized_layers == self.nb_vars - 1
        if is_last:
            bits = np.floor((self.total_bits - self.w_bits_used) / n_here) * np.ones((1, 1))
        else:
            bits = np.minimum(np.round(action) + lo, lo + np.floor(spare * 1.0 / n_here))
        bits = np.minimum(bits, hi)
        self.quantized_layers += 1
        self.num_weights_to_quantize -= n_here
        self.w_bits_used += bits[0][0] * n_here
        return bits
