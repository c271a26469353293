"""Bit allocation for the UniformQuantLearner (/root/reference/learners/uniform_quantization/bit_optimizer.py:57-366).

Without the RL agent every layer gets the flag value (:128-135).  With `--uql_enbl_rl_agent` a DDPG agent searches the
per-layer weight bit-widths under the budget `uql_equivalent_bits` x #weights: each roll-out walks the layers (in
random order), turns the actor's action into a bit-width that keeps the budget feasible (RLHelper.calc_w), restores
the pre-trained weights, fine-tunes for a short while with those bit-widths, and takes the validation accuracy as the
reward of every transition of the roll-out; the best allocation seen is returned.

The reference drives TensorFlow sessions and savers directly and passes the chosen bits between ranks through text
files in the working directory (:343-366); here the training-side work goes through four methods of the learner
(`rl_restore`, `rl_set_bits`, `rl_finetune`, `rl_evaluate`) and the bits travel by one broadcast (SURVEY §8f-3:
file-free).  The search itself — agent, replay, rewards — runs on the primary worker only, as in the reference."""
import numpy as np
import torch

from ...flags import FLAGS, DEFINE_boolean, DEFINE_float, DEFINE_integer, DEFINE_string
from ...rl_agents.ddpg.agent import Agent as DdpgAgent
from ...utils.multi_gpu_wrapper import MultiGpuWrapper as mgw
from .rl_helper import RLHelper

DEFINE_float('uql_equivalent_bits', 4, 'equivalent compression bits for non-rl quantization')
DEFINE_integer('uql_nb_rlouts', 200, 'total number of rlouts for rl training')
DEFINE_integer('uql_w_bit_min', 2, 'minimum number of bits for weights')
DEFINE_integer('uql_w_bit_max', 8, 'maximum number of bits for weights')
DEFINE_integer('uql_tune_layerwise_steps', 100, 'fine tuning steps for each layer')
DEFINE_integer('uql_tune_global_steps', 2000, 'fine tuning steps for each layer')
DEFINE_string('uql_tune_save_path', './rl_tune_models/model.ckpt', 'dir to save tuned models during rl trianing')
DEFINE_integer('uql_tune_disp_steps', 300, 'interval steps to show tuning details')
DEFINE_boolean('uql_enbl_random_layers', True, 'enable random permutation of layers for the rl agent')
DEFINE_boolean('uql_enbl_rl_agent', False, 'enable rl agent for uniform quantization')
DEFINE_boolean('uql_enbl_rl_global_tune', True, 'Tune the weights globally before get reward or not')
DEFINE_boolean('uql_enbl_rl_layerwise_tune', False, 'Tune the weights layerwisely before get reward or not')


def is_primary_worker():
    return not FLAGS.enbl_multi_gpu or mgw.rank() == 0


def broadcast_list(values, length, device='cpu'):
    """Rank 0's list of numbers to every rank (replaces the reference's arranged_layer_bits.txt round trip)."""
    if not FLAGS.enbl_multi_gpu or mgw.size() == 1:
        return [float(v) for v in values]
    t = torch.zeros(length, dtype=torch.float64, device=device)
    if mgw.rank() == 0:
        t.copy_(torch.as_tensor(np.asarray(values, np.float64)))
    mgw.broadcast_global_variables([t])
    return t.cpu().tolist()


class BitOptimizer(object):  # pylint: disable=too-many-instance-attributes
    """Currently only weight bits are inferred via RL; activations stay at 32 bits during the search."""
    PREFIX = 'uql'       # the non-uniform learner's optimizer is this class on the nuql_* flags

    def _f(self, name):
        return getattr(FLAGS, '%s_%s' % (self.PREFIX, name))

    def __init__(self, dataset_name, weights, statistics, tuner=None, barrier_fn=None, seed=None):
        """weights: the kernels to quantize (objects with .shape); statistics: the learner's dict ('num_weights',
        'nb_matmuls', 'nb_activations'); tuner: the learner (only needed with the RL agent)."""
        self.dataset_name = dataset_name
        self.weights = weights
        self.statistics = statistics
        self.tuner = tuner
        self.auto_barrier = barrier_fn or (lambda: None)
        self.nb_matmuls = statistics['nb_matmuls']
        self.nb_activations = statistics['nb_activations']
        if not self._f('enbl_rl_agent'):
            return
        if tuner is None:
            raise ValueError('the RL bit search needs the learner to fine-tune and evaluate roll-outs')
        if self._f('enbl_rl_layerwise_tune'):
            # get_layerwise_tune_op (utils.py:136-161) minimises mean((conv(x, Q(v)) - conv(x, v))^2) over v ALONE: both terms
            # carry v through an identity (the STE), so its gradient is zero up to the rounding of the STE chain — Adam
            # then takes lr-sized steps along the sign of that rounding noise.  Off by default in the reference; not built.
            raise NotImplementedError('layer-wise fine-tuning inside roll-outs (--%s_enbl_rl_layerwise_tune, off by '
                                      'default in the reference) is not built; use the global fine-tuning' % self.PREFIX)
        self.total_num_weights = sum(statistics['num_weights'])
        self.total_bits = self.total_num_weights * self._f('equivalent_bits')
        self.w_rl_helper = RLHelper(self.total_bits, statistics['num_weights'], [tuple(w.shape) for w in weights],
                                    random_layers=self._f('enbl_random_layers'), flag_prefix=self.PREFIX)
        self.mgw_size = int(mgw.size()) if FLAGS.enbl_multi_gpu else 1
        self.tune_global_steps = int(self._f('tune_global_steps') / self.mgw_size)
        self.tune_global_disp_steps = int(self._f('tune_disp_steps') / self.mgw_size)
        self.s_dims = self.w_rl_helper.s_dims
        self.a_dims = 1
        buff_size = len(weights) * int(self._f('nb_rlouts') // 4)
        self.agent = DdpgAgent(self.s_dims, self.a_dims, self._f('nb_rlouts'), buff_size, a_min=0.,
                               a_max=self._f('w_bit_max') - self._f('w_bit_min'), seed=seed)
        self.reward_list = []

    def run(self):
        """The bit allocation, with the RL search or without."""
        if self._f('enbl_rl_agent'):
            return self.__calc_optimal_bits()
        return [self._f('weight_bits')] * self.nb_matmuls, [self._f('activation_bits')] * self.nb_activations

    # ------------------------------------------------------------------ search
    def __calc_optimal_bits(self):
        fp_a_bit_list = [32] * self.nb_activations
        optimal_reward, optimal_bits = -np.inf, None
        if is_primary_worker():
            self.agent.init()
        for idx_rlout in range(self._f('nb_rlouts')):
            arranged, states_n_actions = None, None
            if is_primary_worker():
                print('starting %d-th roll-out:' % idx_rlout)
                arranged, states_n_actions = self.__calc_rollout_actions(idx_rlout)
            self.auto_barrier()
            arranged = [int(round(b)) for b in broadcast_list(arranged, self.nb_matmuls, self.tuner.device)]
            reward = self.__calc_rollout_reward(arranged, fp_a_bit_list)
            self.auto_barrier()
            if is_primary_worker():
                self.reward_list.append(reward[0][0])
                self.agent.finalize_rlout(reward)
                self.__record_rollout_transitions(states_n_actions, reward)
                self.__train_rl_agent(idx_rlout)
                if optimal_reward < reward[0][0]:
                    optimal_reward, optimal_bits = reward[0][0], arranged
            self.auto_barrier()
        if is_primary_worker():
            print('Finished RL training')
            print('Optimal reward: {0}, Optimal w_bit_list: {1}'.format(optimal_reward, optimal_bits))
        optimal_bits = [int(round(b)) for b in broadcast_list(optimal_bits, self.nb_matmuls, self.tuner.device)]
        return optimal_bits, fp_a_bit_list

    def __calc_rollout_actions(self, idx_rlout):
        """One pass of the noisy actor over the layers -> (bits in layer order, [(state, action)] in layer order)."""
        self.agent.init_rlout()
        self.w_rl_helper.reset()
        states_n_actions = [(None, None)] * self.nb_matmuls
        arranged = [-1] * self.nb_matmuls
        for idx in self.w_rl_helper.layer_idxs:
            state = self.w_rl_helper.calc_state(idx)
            action = self.w_rl_helper.calc_w(self.agent.actions_noisy(state), idx)
            assert 1 <= action[0][0] <= 32, 'the quantization bits must be in [1, 32]'
            assert np.shape(action) == (1, 1), '"action" must be in shape (1,1)'
            states_n_actions[idx] = (state, action)
            arranged[idx] = action[0][0]
        assert -1 not in arranged, 'Some layers are not assigned with proper bits'
        print('Un-allocated bit percentage: %.3f' % self.check_bits(arranged))
        print('#_rlout: {0}, layer_bits: {1}'.format(idx_rlout, arranged))
        return arranged, states_n_actions

    def check_bits(self, bit_list):
        """Fraction of the budget left unused; an allocation over the budget is an error (:316-326)."""
        used_bits = sum(v * p for v, p in zip(bit_list, self.statistics['num_weights']))
        if self.total_bits < used_bits:
            raise ValueError('The average bit is out of constraint')
        return (self.total_bits - used_bits) / self.total_bits

    def __calc_rollout_reward(self, layer_bits, a_bits):
        """Restore -> fine-tune with these bits -> validation accuracy (top-1 on CIFAR-10, top-5 on ILSVRC-12)."""
        tuner = self.tuner
        if self._f('enbl_rl_global_tune'):
            tuner.rl_restore()
        tuner.rl_set_bits(layer_bits, a_bits)
        self.auto_barrier()
        if self._f('enbl_rl_global_tune'):
            tuner.rl_finetune(self.tune_global_steps, self.tune_global_disp_steps)
        if not is_primary_worker():
            return None
        _, acc_top1, acc_top5 = tuner.rl_evaluate()
        if self.dataset_name == 'cifar_10':
            reward = self.w_rl_helper.calc_reward(acc_top1)
        elif self.dataset_name == 'ilsvrc_12':
            reward = self.w_rl_helper.calc_reward(acc_top5)
        else:
            raise ValueError('Unknown dataset name')
        print('acc_top1 = %.4f | acc_top5 = %.4f | reward = %.4f' % (acc_top1, acc_top5, reward[0][0]))
        return reward

    def __record_rollout_transitions(self, states_n_actions, reward):
        """Layer n -> layer n + 1 in LAYER order (not visiting order), every transition carrying the roll-out's
        reward; the last layer is terminal with an all-zero next state (:291-303)."""
        for n in range(self.nb_matmuls):
            state, action = states_n_actions[n]
            if n != self.nb_matmuls - 1:
                terminal, state_next = np.zeros((1, 1)), states_n_actions[n + 1][0]
            else:
                terminal, state_next = np.ones((1, 1)), np.zeros((1, self.s_dims))
            self.agent.record(state, action, reward, terminal, state_next)

    def __train_rl_agent(self, idx_rlout):
        for _ in range(self.nb_matmuls):
            actor_loss, critic_loss, param_noise_std = self.agent.train()
        print('roll-out #%d: a-loss = %.2e | c-loss = %.2e | noise std. = %.2e'
              % (idx_rlout, actor_loss, critic_loss, param_noise_std))
