"""Replay buffer of (state, action, reward, terminal, next state) transitions
(/root/reference/rl_agents/ddpg/replay_buffer.py:21-121): a fixed-size ring that only becomes sampleable once it is
completely full, sampled uniformly with replacement."""
import numpy as np

KEYS = ('states', 'actions', 'rewards', 'terminals', 'states_next')


class ReplayBuffer(object):
    def __init__(self, s_dims, a_dims, buf_size, seed=None):
        self.s_dims, self.a_dims, self.buf_size = s_dims, a_dims, int(buf_size)
        widths = dict(states=s_dims, actions=a_dims, rewards=1, terminals=1, states_next=s_dims)
        self.buffers = {k: np.zeros((self.buf_size, widths[k]), np.float32) for k in KEYS}
        self.idx_smpl = 0            # next slot to write
        self.nb_smpls = 0            # valid transitions
        self.rng = np.random.RandomState(seed)

    def reset(self):
        self.idx_smpl, self.nb_smpls = 0, 0

    def is_ready(self):
        """Sampling starts only when every slot holds a transition (replay_buffer.py:66-73)."""
        return self.nb_smpls == self.buf_size

    def append(self, states, actions, rewards, terminals, states_next):
        """Write n transitions at the cursor, wrapping to the head of the ring."""
        batch = dict(zip(KEYS, (states, actions, rewards, terminals, states_next)))
        n = np.asarray(states).shape[0]
        if n > self.buf_size:
            raise ValueError('%d transitions do not fit a buffer of %d' % (n, self.buf_size))
        slots = (self.idx_smpl + np.arange(n)) % self.buf_size
        for k in KEYS:
            self.buffers[k][slots] = np.asarray(batch[k], np.float32).reshape(n, -1)
        self.idx_smpl = (self.idx_smpl + n) % self.buf_size
        self.nb_smpls = min(self.nb_smpls + n, self.buf_size)

    def sample(self, batch_size):
        idxs = self.rng.randint(0, self.nb_smpls, batch_size)
        return {k: self.buffers[k][idxs] for k in KEYS}
