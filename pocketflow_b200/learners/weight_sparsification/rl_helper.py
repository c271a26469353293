"""Roll-out bookkeeping for the pruning-ratio search (/root/reference/learners/weight_sparsification/rl_helper.py:24-161):
per-layer state vectors (normalised by their column maxima) and the map from the actor's action in [0, 1] to a pruning
ratio that keeps the overall target reachable."""
import numpy as np

from ...flags import FLAGS


class RLHelper(object):
    def __init__(self, var_shapes, skip_head_n_tail):
        """var_shapes: shapes of the maskable kernels in layer order (rank 2 or 4); skip_head_n_tail: never prune the
        first and the last layer (the reference does this on CIFAR-10)."""
        nb_vars = len(var_shapes)
        shapes = []
        self.prune_ratios = np.zeros(nb_vars)
        self.nb_params_full = np.zeros(nb_vars)
        for idx, shape in enumerate(var_shapes):
            shape = np.asarray(shape, np.float64)
            assert shape.size in [2, 4], '# of variable dimensions is %d (invalid)' % shape.size
            shape = np.hstack((np.ones(2), shape)) if shape.size == 2 else shape
            shapes.append(shape)
            self.nb_params_full[idx] = np.prod(shape)
        self.s_dims = nb_vars + 4 + 3              # one-hot id, shape, #params of this / earlier (kept) / later layers
        self.states = np.zeros((nb_vars, self.s_dims))
        for idx in range(nb_vars):
            state = self.states[idx]
            state[idx] = 1.0
            state[nb_vars:nb_vars + 4] = shapes[idx]
            state[nb_vars + 4] = self.nb_params_full[idx]
            state[nb_vars + 6] = np.sum(self.nb_params_full[idx + 1:])
        # column nb_vars + 5 (parameters kept in the earlier layers) is filled in per call; it shares the last
        # column's normaliser
        self.state_normalizer = np.max(self.states, axis=0)
        self.state_normalizer[-2] = self.state_normalizer[-1]
        keep = 1.0 - FLAGS.ws_prune_ratio
        self.prune_ratios_min = max(0.0, 1.0 - keep * 3.0) * np.ones(nb_vars)
        self.prune_ratios_max = (1.0 - keep / 3.0) * np.ones(nb_vars)
        if skip_head_n_tail:
            for arr in (self.prune_ratios_min, self.prune_ratios_max):
                arr[0] = 0.0
                arr[-1] = 0.0

    def calc_state(self, idx):
        state = np.copy(self.states[idx])
        state[-2] = np.sum(self.nb_params_full[:idx] * (1.0 - self.prune_ratios[:idx]))
        state /= self.state_normalizer
        return state[None, :]

    def calc_reward(self, accuracy):
        if FLAGS.ws_reward_type == 'single-obj':
            return accuracy
        if FLAGS.ws_reward_type == 'multi-obj':
            return accuracy * np.log(1.0 + self.calc_overall_prune_ratio())
        raise ValueError('unrecognized reward type: ' + FLAGS.ws_reward_type)

    def cvt_action_to_prune_ratio(self, idx, action):
        """action 0.5 -> the target ratio; 0 -> this layer's minimum, 1 -> its maximum, linear in between; clipped."""
        pr_min, pr_max = self.__calc_prune_ratio_min_max(idx)
        if action > 0.5:
            prune_ratio = pr_max - (1.0 - action) / 0.5 * (pr_max - FLAGS.ws_prune_ratio)
        else:
            prune_ratio = pr_min + (action - 0.0) / 0.5 * (FLAGS.ws_prune_ratio - pr_min)
        self.prune_ratios[idx] = max(pr_min, min(pr_max, prune_ratio))
        return self.prune_ratios[idx]

    def calc_overall_prune_ratio(self):
        return np.sum(self.nb_params_full * self.prune_ratios) / np.sum(self.nb_params_full)

    def __calc_prune_ratio_min_max(self, idx):
        """With the single-objective reward the overall target is a hard constraint: the minimum for layer idx is
        raised to what is still needed if every later layer were pruned at its maximum."""
        pr_min, pr_max = self.prune_ratios_min[idx], self.prune_ratios_max[idx]
        if FLAGS.ws_reward_type == 'single-obj':
            pruned_at_most = np.sum(self.nb_params_full[:idx] * self.prune_ratios[:idx]) \
                + np.sum(self.nb_params_full[idx + 1:] * self.prune_ratios_max[idx + 1:])
            pruned_needed = np.sum(self.nb_params_full) * FLAGS.ws_prune_ratio
            pr_req = (pruned_needed - pruned_at_most) / self.nb_params_full[idx]
            assert pr_req < pr_max + 1e-4, 'cannot reach the required pruning ratio: %f vs. %f' % (pr_req, pr_max)
            pr_min = max(pr_min, pr_req)
        return pr_min, pr_max
