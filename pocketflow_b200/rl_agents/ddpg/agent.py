"""DDPG (deep deterministic policy gradient) agent (/root/reference/rl_agents/ddpg/agent.py:110-418), torch-native.

The reference builds the agent as a TensorFlow graph driven through `sess.run`; here the same pieces are methods:
`actions_noisy(states)` / `actions_clean(states)` replace running the `actions_noisy` / `actions_clean` tensors, the rest
of the interface (init, init_rlout, finalize_rlout, record, train) is unchanged.  Main / target / parameter-noise
copies of the actor, main / target critic, a replay ring, a reward baseline (exponential moving average subtracted
from the sampled rewards) and the two noise protocols.  State / return normalisation is hard-wired off in the reference
(agent.py:249-257) and is not carried over.  Deviation (flagged): torch.optim.Adam places epsilon after the bias
correction, TensorFlow's Adam before — irrelevant for a stochastic search, so not reproduced here (the training step's
optimizer kernel does reproduce TensorFlow's form)."""
import copy

import numpy as np
import torch

from ...flags import FLAGS, DEFINE_float, DEFINE_integer, DEFINE_boolean
from .actor_critic import Actor, Critic
from .noise import AdaptiveNoiseSpec, TimeDecayNoiseSpec
from .replay_buffer import ReplayBuffer

DEFINE_float('ddpg_tau', 0.01, 'DDPG: target networks\' update coefficient')
DEFINE_float('ddpg_gamma', 0.9, 'DDPG: reward discounting factor')
DEFINE_float('ddpg_lrn_rate', 1e-3, 'DDPG: actor & critic networks\' learning rate')
DEFINE_float('ddpg_loss_w_dcy', 0.0, 'DDPG: weight decaying coefficient')
DEFINE_integer('ddpg_record_step', 1, 'DDPG: recording step size')
DEFINE_integer('ddpg_batch_size', 64, 'DDPG: batch size')
DEFINE_boolean('ddpg_enbl_bsln_func', True, 'DDPG: enable baseline function')
DEFINE_float('ddpg_bsln_decy_rate', 0.95, 'DDPG: baseline function\'s decaying rate')


def _l2(params):
    return sum(0.5 * (p ** 2).sum() for p in params)


def _t(a):
    return torch.as_tensor(np.asarray(a, np.float32))


class Agent(object):  # pylint: disable=too-many-instance-attributes
    def __init__(self, s_dims, a_dims, nb_rlouts, buf_size, a_min=0.0, a_max=1.0, seed=None):
        self.s_dims, self.a_dims, self.a_min, self.a_max = s_dims, a_dims, float(a_min), float(a_max)
        self.reward_ema = None
        self.in_explore = True
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)
        self.memory = ReplayBuffer(s_dims, a_dims, buf_size, seed)
        if FLAGS.ddpg_noise_prtl == 'adapt':
            self.noise_spec = AdaptiveNoiseSpec()
        elif FLAGS.ddpg_noise_prtl == 'tdecy':
            self.noise_spec = TimeDecayNoiseSpec(nb_rlouts)
        else:
            raise ValueError('unrecognized noise adjustment protocol: ' + FLAGS.ddpg_noise_prtl)
        if FLAGS.ddpg_noise_type not in ('action', 'param'):
            raise ValueError('unrecognized noise type: ' + FLAGS.ddpg_noise_type)
        self.action_noise_std = 0.0
        self.seed = seed
        self.init()

    # ------------------------------------------------------------------ life cycle
    def init(self):
        """Before all roll-outs: fresh networks and optimizers, targets = mains, empty replay ring."""
        if self.seed is not None:
            torch.manual_seed(self.seed)
        self.actor = Actor(self.s_dims, self.a_dims, self.a_min, self.a_max)
        self.critic = Critic(self.s_dims, self.a_dims)
        self.actor_tr, self.critic_tr = copy.deepcopy(self.actor), copy.deepcopy(self.critic)
        self.actor_np, self.actor_ns = copy.deepcopy(self.actor), copy.deepcopy(self.actor)
        for net in (self.actor_tr, self.critic_tr, self.actor_np, self.actor_ns):
            for p in net.parameters():
                p.requires_grad_(False)
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=FLAGS.ddpg_lrn_rate, eps=1e-8)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=FLAGS.ddpg_lrn_rate, eps=1e-8)
        self.memory.reset()
        self.noise_spec.reset()
        self.in_explore = True
        self.reward_ema = None

    def init_rlout(self):
        """Before each roll-out: advance the time-decay schedule (once learning has begun) and redraw the noise."""
        if FLAGS.ddpg_noise_prtl == 'tdecy' and not self.in_explore:
            self.noise_spec.adapt()
        if FLAGS.ddpg_noise_type == 'action':
            self.action_noise_std = self.noise_spec.stdev_curr
        else:
            self._perturb(self.actor_np, self.noise_spec.stdev_curr)

    def finalize_rlout(self, rewards):
        """After each roll-out: update the baseline (moving average of the roll-outs' mean reward)."""
        if not FLAGS.ddpg_enbl_bsln_func:
            return
        mean = float(np.mean(rewards))
        if self.reward_ema is None:
            self.reward_ema = mean
        else:
            self.reward_ema = FLAGS.ddpg_bsln_decy_rate * self.reward_ema + (1.0 - FLAGS.ddpg_bsln_decy_rate) * mean

    # ------------------------------------------------------------------ acting
    def actions_clean(self, states):
        """The deterministic policy (deployment)."""
        with torch.no_grad():
            return self.actor(_t(states)).numpy()

    def actions_noisy(self, states):
        """The exploring policy: the parameter-perturbed actor, or the clean one plus clipped Gaussian action noise."""
        with torch.no_grad():
            if FLAGS.ddpg_noise_type == 'param':
                return self.actor_np(_t(states)).numpy()
            a = self.actor(_t(states))
            a = a + torch.randn(a.shape, generator=self.gen) * self.action_noise_std
            return a.clamp(self.a_min, self.a_max).numpy()

    def _perturb(self, noisy, std):
        """noisy <- actor (+ N(0, std) on every perturbable parameter) (agent.py:89-108)."""
        perturbable = {id(p) for p in self.actor.perturbable_params}
        with torch.no_grad():
            for p, q in zip(self.actor.parameters(), noisy.parameters()):
                q.copy_(p)
                if id(p) in perturbable:
                    q.add_(torch.randn(p.shape, generator=self.gen) * float(std))

    # ------------------------------------------------------------------ learning
    def record(self, states, actions, rewards, terminals, states_next):
        """Append transitions (arrays with one row per transition) to the replay ring."""
        step = FLAGS.ddpg_record_step
        n = np.asarray(states).shape[0]
        pick = slice(None, None, step)
        self.memory.append(np.asarray(states)[pick], np.asarray(actions)[pick],
                           np.asarray(rewards, np.float32).reshape(n, 1)[pick],
                           np.asarray(terminals, np.float32).reshape(n, 1)[pick], np.asarray(states_next)[pick])

    def train(self):
        """One actor + critic update from a sampled mini-batch, then the soft target update.
        Returns (actor loss, critic loss, current noise stdev); a no-op until the replay ring is full."""
        if not self.memory.is_ready():
            return 0.0, 0.0, self.noise_spec.stdev_curr
        self.in_explore = False
        if FLAGS.ddpg_noise_prtl == 'adapt':
            mb = self.memory.sample(FLAGS.ddpg_batch_size)
            self._perturb(self.actor_ns, self.noise_spec.stdev_curr)
            with torch.no_grad():
                s = _t(mb['states'])
                self.noise_spec.adapt(float((self.actor(s) - self.actor_ns(s)).abs().mean()))
        mb = self.memory.sample(FLAGS.ddpg_batch_size)
        rewards = mb['rewards'] - (self.reward_ema if FLAGS.ddpg_enbl_bsln_func and self.reward_ema is not None else 0.0)
        s, a, r = _t(mb['states']), _t(mb['actions']), _t(rewards)
        term, s_next = _t(mb['terminals']), _t(mb['states_next'])
        with torch.no_grad():
            target_q = r + (1.0 - term) * FLAGS.ddpg_gamma * self.critic_tr(s_next, self.actor_tr(s_next))
        # both losses are taken at the pre-update parameters, as one sess.run of [actor_updt, critic_updt] does
        critic_loss = 0.5 * ((self.critic(s, a) - target_q) ** 2).sum()
        actor_loss = -self.critic(s, self.actor(s)).mean()
        if FLAGS.ddpg_loss_w_dcy:
            critic_loss = critic_loss + FLAGS.ddpg_loss_w_dcy * _l2(self.critic.parameters())
            actor_loss = actor_loss + FLAGS.ddpg_loss_w_dcy * _l2(self.actor.parameters())
        self.actor_opt.zero_grad()
        self.critic_opt.zero_grad()
        actor_grads = torch.autograd.grad(actor_loss, list(self.actor.parameters()))
        critic_grads = torch.autograd.grad(critic_loss, list(self.critic.parameters()))
        for p, g in zip(self.actor.parameters(), actor_grads):
            p.grad = g
        for p, g in zip(self.critic.parameters(), critic_grads):
            p.grad = g
        self.actor_opt.step()
        self.critic_opt.step()
        tau = FLAGS.ddpg_tau
        with torch.no_grad():
            for net, net_tr in ((self.actor, self.actor_tr), (self.critic, self.critic_tr)):
                for p, p_tr in zip(net.parameters(), net_tr.parameters()):
                    p_tr.mul_(1.0 - tau).add_(p, alpha=tau)
        return actor_loss.item(), critic_loss.item(), self.noise_spec.stdev_curr
