"""Exploration-noise schedules (/root/reference/rl_agents/ddpg/noise.py:23-86): the standard deviation of the
parameter / action noise either adapts to a target action distance or decays geometrically over the roll-outs."""
from ...flags import FLAGS, DEFINE_string, DEFINE_float

DEFINE_string('ddpg_noise_type', 'param', 'DDPG: noise type (\'action\' OR \'param\')')
DEFINE_string('ddpg_noise_prtl', 'tdecy', 'DDPG: noise adjustment protocol (\'adapt\' OR \'tdecy\')')
DEFINE_float('ddpg_noise_std_init', 1e+0, 'DDPG: parameter / action noise\'s initial stdev.')
DEFINE_float('ddpg_noise_dst_finl', 1e-2, 'DDPG: action noise\'s final distance')
DEFINE_float('ddpg_noise_adpt_rat', 1.03, 'DDPG: parameter noise\'s adaption rate')
DEFINE_float('ddpg_noise_std_finl', 1e-5, 'DDPG: parameter / action noise\'s final stdev.')


class AdaptiveNoiseSpec(object):
    """<ddpg_noise_type> 'param' + <ddpg_noise_prtl> 'adapt': shrink the stdev while the perturbed actor's actions are
    further than ddpg_noise_dst_finl from the clean ones, grow it otherwise."""

    def __init__(self):
        self.stdev_curr = FLAGS.ddpg_noise_std_init

    def reset(self):
        self.stdev_curr = FLAGS.ddpg_noise_std_init

    def adapt(self, dst_curr):
        if dst_curr > FLAGS.ddpg_noise_dst_finl:
            self.stdev_curr /= FLAGS.ddpg_noise_adpt_rat
        else:
            self.stdev_curr *= FLAGS.ddpg_noise_adpt_rat


class TimeDecayNoiseSpec(object):
    """<ddpg_noise_prtl> 'tdecy': stdev_init -> stdev_finl in nb_rlouts equal geometric steps."""

    def __init__(self, nb_rlouts):
        self.stdev_curr = FLAGS.ddpg_noise_std_init
        self.decy_rat = (FLAGS.ddpg_noise_std_finl / FLAGS.ddpg_noise_std_init) ** (1.0 / nb_rlouts)

    def reset(self):
        self.stdev_curr = FLAGS.ddpg_noise_std_init

    def adapt(self):
        self.stdev_curr *= self.decy_rat
