"""Pruning-ratio protocols of the WeightSparseLearner
(/root/reference/learners/weight_sparsification/pr_optimizer.py:385-409).  'uniform' and 'heurist'
are host-side formulas; 'optimal' (DDPG roll-outs around the step) is a "next" row."""
import numpy as np

from ...flags import FLAGS


class PROptimizer(object):
    def __init__(self, maskable_vars):
        self.maskable_vars = maskable_vars

    def run(self):
        if FLAGS.ws_prune_ratio_prtl == 'uniform':
            return self.__calc_uniform_prune_ratios()
        elif FLAGS.ws_prune_ratio_prtl == 'heurist':
            return self.__calc_heurist_prune_ratios()
        elif FLAGS.ws_prune_ratio_prtl == 'optimal':
            raise NotImplementedError('the RL-based protocol is not built yet; use --ws_prune_ratio_prtl uniform|heurist')
        raise ValueError('unrecognized pruning ratio protocol: ' + FLAGS.ws_prune_ratio_prtl)

    def __calc_uniform_prune_ratios(self):
        return [(var.name, FLAGS.ws_prune_ratio) for var in self.maskable_vars]

    def __calc_heurist_prune_ratios(self):
        """ratio_i = alpha * log(n_i), alpha = s * sum(n_i) / sum(n_i * log(n_i))  (:394-409)."""
        nb_params = np.array([var.numel for var in self.maskable_vars], dtype=np.float64)
        alpha = FLAGS.ws_prune_ratio * np.sum(nb_params) / np.sum(nb_params * np.log(nb_params))
        return [(var.name, float(alpha * np.log(n))) for var, n in zip(self.maskable_vars, nb_params)]
