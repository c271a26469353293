"""Roll-out bookkeeping of the pruning-ratio search (behaviour of
/root/reference/learners/weight_sparsification/rl_helper.py:24-161).

State of layer i (length L + 7 for L maskable layers), divided column-wise by the largest value the column can take:
one-hot(i) | kernel shape as 4 numbers | #params(i) | #params kept so far in layers < i (depends on the ratios already
chosen in this roll-out) | #params in layers > i.  An action a in [0, 1] maps linearly to a ratio: a = 0.5 is the
overall target, 0 the layer's floor, 1 its ceiling; with the single-objective reward the floor is raised so that the
overall target stays reachable even if every later layer is pruned at its ceiling.  float64 numpy throughout."""
import numpy as np

from ...flags import FLAGS


def _conv_shape(shape):
    shape = np.asarray(shape, np.float64)
    assert shape.size in [2, 4], '# of variable dimensions is %d (invalid)' % shape.size
    return shape if shape.size == 4 else np.hstack((np.ones(2), shape))


class RLHelper(object):
    def __init__(self, var_shapes, skip_head_n_tail):
        """var_shapes: kernel shapes of the maskable variables in layer order; skip_head_n_tail: pin the first and the
        last layer at ratio 0 (what the reference does on CIFAR-10)."""
        shapes = [_conv_shape(s) for s in var_shapes]
        L = len(shapes)
        self.nb_params_full = np.array([np.prod(s) for s in shapes]) if L else np.zeros(0)
        self.prune_ratios = np.zeros(L)
        self.s_dims = L + 7
        after = np.array([np.sum(self.nb_params_full[i + 1:]) for i in range(L)])
        self.states = np.zeros((L, self.s_dims))
        self.states[:, :L] = np.eye(L)
        self.states[:, L:L + 4] = np.array(shapes).reshape(L, 4)
        self.states[:, L + 4] = self.nb_params_full
        self.states[:, L + 6] = after                       # column L + 5 is filled in by calc_state
        self.state_normalizer = self.states.max(axis=0)
        self.state_normalizer[L + 5] = self.state_normalizer[L + 6]
        keep = 1.0 - FLAGS.ws_prune_ratio
        floor, ceiling = max(0.0, 1.0 - 3.0 * keep), 1.0 - keep / 3.0
        self.prune_ratios_min = np.full(L, floor)
        self.prune_ratios_max = np.full(L, ceiling)
        if skip_head_n_tail:
            self.prune_ratios_min[[0, -1]] = 0.0
            self.prune_ratios_max[[0, -1]] = 0.0

    def calc_state(self, idx):
        state = self.states[idx].copy()
        state[-2] = np.sum(self.nb_params_full[:idx] * (1.0 - self.prune_ratios[:idx]))
        return (state / self.state_normalizer)[None, :]

    def calc_overall_prune_ratio(self):
        return np.sum(self.nb_params_full * self.prune_ratios) / np.sum(self.nb_params_full)

    def calc_reward(self, accuracy):
        kind = FLAGS.ws_reward_type
        if kind == 'multi-obj':
            return accuracy * np.log(1.0 + self.calc_overall_prune_ratio())
        if kind != 'single-obj':
            raise ValueError('unrecognized reward type: ' + kind)
        return accuracy

    def _bounds(self, idx):
        lo, hi = self.prune_ratios_min[idx], self.prune_ratios_max[idx]
        if FLAGS.ws_reward_type == 'single-obj':
            n = self.nb_params_full
            best_case = np.sum(n[:idx] * self.prune_ratios[:idx]) + np.sum(n[idx + 1:] * self.prune_ratios_max[idx + 1:])
            needed = (np.sum(n) * FLAGS.ws_prune_ratio - best_case) / n[idx]
            assert needed < hi + 1e-4, 'cannot reach the required pruning ratio: %f vs. %f' % (needed, hi)
            lo = max(lo, needed)
        return lo, hi

    def cvt_action_to_prune_ratio(self, idx, action):
        lo, hi = self._bounds(idx)
        target = FLAGS.ws_prune_ratio
        if action > 0.5:
            ratio = hi - (1.0 - action) / 0.5 * (hi - target)
        else:
            ratio = lo + (action - 0.0) / 0.5 * (target - lo)
        self.prune_ratios[idx] = max(lo, min(hi, ratio))
        return self.prune_ratios[idx]
