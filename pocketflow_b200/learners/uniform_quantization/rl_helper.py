"""Roll-out bookkeeping for the bit-allocation search (/root/reference/learners/uniform_quantization/rl_helper.py:26-122):
the state vector of every layer and the projection of the actor's raw action onto the bit-widths the remaining budget
still allows."""
import random

import numpy as np

from ...flags import FLAGS


class RLHelper(object):
    # pylint: disable=too-many-instance-attributes
    def __init__(self, total_bits, num_weights, var_shapes, random_layers=False):
        """total_bits: the budget (sum over layers of bits x #weights); num_weights: #weights per layer;
        var_shapes: the kernels' shapes (rank 2 = dense, rank 4 = conv), in layer order."""
        self.nb_vars = len(num_weights)
        self.num_weights = num_weights
        self.total_num_weights = sum(num_weights)
        self.s_dims = self.nb_vars + 6             # one-hot layer id, 4 shape entries, 2 size ratios
        self.total_bits = total_bits
        self.w_bits_used = 0
        self.random_layers = random_layers
        self.layer_idxs = list(range(self.nb_vars))
        self.num_weights_to_quantize = self.total_num_weights
        self.quantized_layers = 0
        self.var_shapes = []
        for shape in var_shapes:
            assert len(shape) in [2, 4], 'Unknown weight shape. Must be a 2 (fc) or 4 (conv) dimensional.'
            shape = np.asarray(shape, np.float64)
            self.var_shapes.append(np.hstack((np.ones(2), shape)) if len(shape) == 2 else shape)
        self.states = np.zeros((self.nb_vars, self.s_dims))
        for idx in range(self.nb_vars):
            state = self.states[idx]
            state[idx] = 1.0
            state[self.nb_vars:self.nb_vars + 4] = self.var_shapes[idx]
            state[self.nb_vars + 4] = self.num_weights[idx] / np.max(self.num_weights)
            state[self.nb_vars + 5] = np.sum(self.num_weights[idx + 1:]) / self.total_num_weights

    def calc_state(self, idx):
        return np.copy(self.states[idx])[None, :]

    def calc_reward(self, accuracy):
        return accuracy * np.ones((1, 1))

    def reset(self):
        """Before each roll-out."""
        self.w_bits_used = 0
        self.quantized_layers = 0
        if self.random_layers:
            random.shuffle(self.layer_idxs)
        self.num_weights_to_quantize = self.total_num_weights

    def calc_w(self, action, idx):
        """Bit-width for layer `idx` from the actor's output `action` (shape (1, 1), in [0, w_bit_max - w_bit_min]):
        rounded, shifted by the minimum, capped so that every layer still to come can get the minimum; the last
        layer of the roll-out takes whatever the budget has left (at most the maximum)."""
        duty = self.total_bits - self.w_bits_used - self.num_weights_to_quantize * FLAGS.uql_w_bit_min
        assert duty >= 0, 'Not enough budget for layer {}'.format(idx)
        if self.quantized_layers != self.nb_vars - 1:
            action = np.round(action) + FLAGS.uql_w_bit_min
            action = np.minimum(action, FLAGS.uql_w_bit_min + np.floor(duty * 1.0 / self.num_weights[idx]))
        else:
            action = np.floor((self.total_bits - self.w_bits_used) / self.num_weights[idx]) * np.ones((1, 1))
        action = np.minimum(action, FLAGS.uql_w_bit_max)
        self.w_bits_used += action[0][0] * self.num_weights[idx]
        self.num_weights_to_quantize -= self.num_weights[idx]
        self.quantized_layers += 1
        return action
