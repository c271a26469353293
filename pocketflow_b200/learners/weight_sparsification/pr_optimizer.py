"""Pruning-ratio protocols of the WeightSparseLearner
(/root/reference/learners/weight_sparsification/pr_optimizer.py:83-611).

'uniform' and 'heurist' are host-side formulas (:385-409).  'optimal' is a DDPG search (:411-470): each roll-out walks
the maskable layers in order, maps the actor's action to a pruning ratio that keeps the overall target reachable
(RLHelper), prunes the pre-trained model at those ratios, retrains briefly and takes the validation accuracy as the
reward of every transition.  The device half is reached through a `tuner` object (the WeightSparseLearner):
`pr_prune(ratios)` (restore the full model, mask it), `pr_retrain(nb_iters_rg, nb_iters_ft)`,
`pr_evaluate() -> (loss, {metric: value})`.  The learner implements them on its own training step, with the layer-wise
regression stage (:283-314, :542-548: every pruned layer's output regressed onto the full model's, Adam, masked
gradients, inference-mode BN) on a forward-only copy of the full model; see the deviations of the global fine-tuning
listed there.  The ratios and the reward travel
between ranks by broadcast instead of the reference's ./ws.prune.ratios and ./ws.reward files."""
import math

import numpy as np

from ...flags import FLAGS
from ...utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from ..uniform_quantization.bit_optimizer import broadcast_list, is_primary_worker
from .rl_helper import RLHelper


class PROptimizer(object):
    def __init__(self, maskable_vars, dataset_name=None, tuner=None, seed=None):
        """tuner (only for 'optimal'): .device, pr_prune(prune_ratios) (restore the full model, mask it at these
        ratios, fresh optimizers), pr_retrain(nb_iters_rg, nb_iters_ft), pr_evaluate() -> (loss, {metric: value})."""
        self.maskable_vars = maskable_vars
        self.dataset_name = dataset_name
        self.tuner = tuner
        self.seed = seed
        self.rewards = []
        if FLAGS.ws_prune_ratio_prtl not in ('uniform', 'heurist', 'optimal'):
            raise ValueError('unrecognzed WS pruning ratio protocol: ' + FLAGS.ws_prune_ratio_prtl)

    def run(self):
        if FLAGS.ws_prune_ratio_prtl == 'uniform':
            out = self.__calc_uniform_prune_ratios()
        elif FLAGS.ws_prune_ratio_prtl == 'heurist':
            out = self.__calc_heurist_prune_ratios()
        elif FLAGS.ws_prune_ratio_prtl == 'optimal':
            out = self.__calc_optimal_prune_ratios()
        else:
            raise ValueError('unrecognized pruning ratio protocol: ' + FLAGS.ws_prune_ratio_prtl)
        return out

    def __calc_uniform_prune_ratios(self):
        return [(var.name, FLAGS.ws_prune_ratio) for var in self.maskable_vars]

    def __calc_heurist_prune_ratios(self):
        """ratio_i = alpha * log(n_i), alpha = s * sum(n_i) / sum(n_i * log(n_i))  (:394-409)."""
        nb_params = np.array([var.numel for var in self.maskable_vars], dtype=np.float64)
        alpha = FLAGS.ws_prune_ratio * np.sum(nb_params) / np.sum(nb_params * np.log(nb_params))
        return [(var.name, float(alpha * np.log(n))) for var, n in zip(self.maskable_vars, nb_params)]

    # ------------------------------------------------------------------ 'optimal'
    def __calc_optimal_prune_ratios(self):
        if self.tuner is None:
            raise ValueError("--ws_prune_ratio_prtl optimal needs the learner (tuner=) to prune, retrain and evaluate "
                             "the roll-outs")
        from ...rl_agents.ddpg.agent import Agent as DdpgAgent
        nb_vars = len(self.maskable_vars)
        primary = is_primary_worker()
        if primary:
            skip_head_n_tail = (self.dataset_name == 'cifar_10')            # skip head & tail layers on CIFAR-10
            self.rl_helper = RLHelper([tuple(v.shape) for v in self.maskable_vars], skip_head_n_tail)
            self.agent = DdpgAgent(self.rl_helper.s_dims, 1, FLAGS.ws_nb_rlouts, nb_vars * FLAGS.ws_nb_rlouts_min,
                                   0.0, 1.0, seed=self.seed)
            self.agent.init()
        reward_best, prune_ratios_best = -np.inf, None
        for idx_rlout in range(FLAGS.ws_nb_rlouts):
            prune_ratios, states_n_actions = None, None
            if primary:
                print('starting %d-th roll-out' % idx_rlout)
                prune_ratios, states_n_actions = self.__calc_rlout_actions()
            prune_ratios = np.array(broadcast_list(prune_ratios, nb_vars, self.tuner.device))
            reward = self.__calc_rlout_reward(prune_ratios)
            reward = broadcast_list([reward] if primary else None, 1, self.tuner.device)[0]
            if primary:
                self.rewards.append(reward)
                self.agent.finalize_rlout(reward * np.ones(nb_vars))
                self.__record_rlout_transitions(states_n_actions, reward)
            if reward_best < reward:
                if primary:
                    print('best reward updated: %.4f -> %.4f' % (reward_best, reward))
                    print('optimal pruning ratios: ' + ' '.join(['%.2f' % pr for pr in prune_ratios]))
                reward_best, prune_ratios_best = reward, np.copy(prune_ratios)
        return [(var.name, float(prune_ratios_best[idx])) for idx, var in enumerate(self.maskable_vars)]

    def __calc_rlout_actions(self):
        """One pass of the noisy actor over the layers; the agent takes one training step per layer (:472-493)."""
        self.agent.init_rlout()
        prune_ratios, states_n_actions = [], []
        for idx in range(len(self.maskable_vars)):
            state = self.rl_helper.calc_state(idx)
            action = self.agent.actions_noisy(state)
            prune_ratios.append(self.rl_helper.cvt_action_to_prune_ratio(idx, action[0][0]))
            states_n_actions.append((state, action))
            actor_loss, critic_loss, noise_std = self.agent.train()
        print('a-loss = %.2e | c-loss = %.2e | noise std. = %.2e' % (actor_loss, critic_loss, noise_std))
        return prune_ratios, states_n_actions

    def __calc_rlout_reward(self, prune_ratios):
        """prune -> (evaluate) -> retrain -> evaluate; the reward is computed on the primary worker (:495-541)."""
        tuner = self.tuner
        tuner.pr_prune(prune_ratios)
        primary = is_primary_worker()
        if primary:
            loss_pre, metrics_pre = tuner.pr_evaluate()
        nb_workers = mgw.size() if FLAGS.enbl_multi_gpu else 1
        tuner.pr_retrain(int(math.ceil(FLAGS.ws_nb_iters_rg / nb_workers)), int(math.ceil(FLAGS.ws_nb_iters_ft / nb_workers)))
        if not primary:
            return None
        loss_post, metrics_post = tuner.pr_evaluate()
        key = 'accuracy' if 'accuracy' in metrics_post else 'acc_top5'
        assert key in metrics_post and key in metrics_pre, 'either <accuracy> or <acc_top5> must be evaluated and returned'
        reward_pre = self.rl_helper.calc_reward(metrics_pre[key])
        reward = self.rl_helper.calc_reward(metrics_post[key])
        metrics_diff = ' | '.join(['%s: %.4f -> %.4f' % (k, metrics_pre[k], metrics_post[k]) for k in metrics_post])
        print('loss: %.4e -> %.4e | %s | reward: %.4f -> %.4f | prune_ratio = %.4f'
              % (loss_pre, loss_post, metrics_diff, reward_pre, reward, self.rl_helper.calc_overall_prune_ratio()))
        return float(reward)

    def __record_rlout_transitions(self, states_n_actions, reward):
        for idx, (state, action) in enumerate(states_n_actions):
            last = idx == len(states_n_actions) - 1
            state_next = np.zeros_like(state) if last else states_n_actions[idx + 1][0]
            self.agent.record(state, action, reward * np.ones((1, 1)), np.ones((1, 1)) * float(last), state_next)
