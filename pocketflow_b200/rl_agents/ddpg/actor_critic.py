"""Actor and critic networks (/root/reference/rl_agents/ddpg/actor_critic.py:22-154) as torch modules: blocks of
dense -> layer norm -> ReLU; the actor ends in a sigmoid scaled to [a_min, a_max], the critic joins the action after its
first block.  Tiny (64-wide) MLPs evaluated on one state at a time: they live on the host."""
import torch
from torch import nn

from ...flags import FLAGS, DEFINE_integer

DEFINE_integer('ddpg_actor_depth', 2, 'DDPG: actor network\'s depth')
DEFINE_integer('ddpg_actor_width', 64, 'DDPG: actor network\'s width')
DEFINE_integer('ddpg_critic_depth', 2, 'DDPG: critic network\'s depth')
DEFINE_integer('ddpg_critic_width', 64, 'DDPG: critic network\'s width')

ENBL_LAYER_NORM = True
LAYER_NORM_EPS = 1e-12          # tf.contrib.layers.layer_norm's variance_epsilon


def dense(n_in, n_out):
    """tf.layers.dense defaults: Glorot-uniform kernel, zero bias."""
    layer = nn.Linear(n_in, n_out)
    nn.init.xavier_uniform_(layer.weight)
    nn.init.zeros_(layer.bias)
    return layer


def dense_block(n_in, units):
    layers = [dense(n_in, units)]
    if ENBL_LAYER_NORM:
        layers.append(nn.LayerNorm(units, eps=LAYER_NORM_EPS))
    layers.append(nn.ReLU())
    return nn.Sequential(*layers)


class Model(nn.Module):
    @property
    def perturbable_params(self):
        """Everything trainable except the layer-norm gains / offsets (actor_critic.py:73-77)."""
        skip = {id(p) for m in self.modules() if isinstance(m, nn.LayerNorm) for p in m.parameters()}
        return [p for p in self.parameters() if id(p) not in skip]


class Actor(Model):
    def __init__(self, s_dims, a_dims, a_min, a_max):
        super(Actor, self).__init__()
        self.a_min, self.a_max = float(a_min), float(a_max)
        blocks, n_in = [], s_dims
        for _ in range(FLAGS.ddpg_actor_depth):
            blocks.append(dense_block(n_in, FLAGS.ddpg_actor_width))
            n_in = FLAGS.ddpg_actor_width
        self.body = nn.Sequential(*blocks)
        self.head = dense(n_in, a_dims)

    def forward(self, states):
        return torch.sigmoid(self.head(self.body(states))) * (self.a_max - self.a_min) + self.a_min


class Critic(Model):
    def __init__(self, s_dims, a_dims):
        super(Critic, self).__init__()
        width = FLAGS.ddpg_critic_width
        self.stem = dense_block(s_dims, width)
        blocks, n_in = [], width + a_dims
        for _ in range(FLAGS.ddpg_critic_depth):
            blocks.append(dense_block(n_in, width))
            n_in = width
        self.body = nn.Sequential(*blocks)
        self.head = dense(n_in, 1)

    def forward(self, states, actions):
        return self.head(self.body(torch.cat([self.stem(states), actions], dim=1)))
