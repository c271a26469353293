"""Exploration-noise schedules of the DDPG agent (behaviour of /root/reference/rl_agents/ddpg/noise.py:23-86).

Both schedules expose the current standard deviation as `stdev_curr`, restart from `--ddpg_noise_std_init` on
`reset()` and move on `adapt(...)`:
* TimeDecayNoiseSpec — geometric decay that reaches `--ddpg_noise_std_finl` after `nb_rlouts` calls;
* AdaptiveNoiseSpec — multiplicative feedback on the measured distance between clean and perturbed actions
  (parameter-space noise): too far -> shrink, close enough -> grow, by the factor `--ddpg_noise_adpt_rat`."""
from ...flags import FLAGS, DEFINE_string, DEFINE_float

DEFINE_string('ddpg_noise_type', 'param', "DDPG exploration noise lives in 'param' (actor weights) or 'action' space")
DEFINE_string('ddpg_noise_prtl', 'tdecy', "DDPG noise schedule: 'tdecy' (geometric decay) or 'adapt' (distance feedback)")
DEFINE_float('ddpg_noise_std_init', 1e+0, 'DDPG noise: standard deviation at the first roll-out')
DEFINE_float('ddpg_noise_dst_finl', 1e-2, "DDPG noise ('adapt'): target distance between clean and noisy actions")
DEFINE_float('ddpg_noise_adpt_rat', 1.03, "DDPG noise ('adapt'): multiplicative step of the feedback")
DEFINE_float('ddpg_noise_std_finl', 1e-5, "DDPG noise ('tdecy'): standard deviation after the last roll-out")


class _NoiseSchedule(object):
    def __init__(self):
        self.stdev_curr = None
        self.reset()

    def reset(self):
        self.stdev_curr = FLAGS.ddpg_noise_std_init


class AdaptiveNoiseSpec(_NoiseSchedule):
    def adapt(self, dst_curr):
        too_far = dst_curr > FLAGS.ddpg_noise_dst_finl
        if too_far:
            self.stdev_curr /= FLAGS.ddpg_noise_adpt_rat
        else:
            self.stdev_curr *= FLAGS.ddpg_noise_adpt_rat


class TimeDecayNoiseSpec(_NoiseSchedule):
    def __init__(self, nb_rlouts):
        super(TimeDecayNoiseSpec, self).__init__()
        span = FLAGS.ddpg_noise_std_finl / FLAGS.ddpg_noise_std_init
        self.decy_rat = span ** (1.0 / nb_rlouts)

    def adapt(self):
        self.stdev_curr *= self.decy_rat
