"""tf.app.flags shim: the same flag names and defaults the reference declares, no TensorFlow.

The reference defines flags at import time in whichever module is imported, e.g.
`tf.app.flags.DEFINE_integer('uql_weight_bits', 4, ...)` (learners/uniform_quantization/learner.py:36).
`FLAGS` here is a process-global namespace with the same semantics; re-defining a flag with a new
default (each net module re-declares lrn_rate_init, loss_w_dcy, ...) overrides the default unless
the value was set explicitly."""
import sys


class _Flags(object):
    def __init__(self):
        object.__setattr__(self, '_defaults', {})
        object.__setattr__(self, '_values', {})
        object.__setattr__(self, '_help', {})

    def _define(self, name, default, helpstr=''):
        self._defaults[name] = default
        self._help[name] = helpstr

    def __getattr__(self, name):
        if name in self._values:
            return self._values[name]
        if name in self._defaults:
            return self._defaults[name]
        raise AttributeError('unknown flag: %s' % name)

    def __setattr__(self, name, value):
        self._values[name] = value

    def __contains__(self, name):
        return name in self._defaults or name in self._values

    def reset(self):
        self._values.clear()

    def parse(self, argv=None):
        """--name value | --name=value | --flag / --noflag for booleans."""
        argv = list(sys.argv[1:] if argv is None else argv)
        i = 0
        while i < len(argv):
            a = argv[i]
            if not a.startswith('--'):
                raise ValueError('unexpected argument: ' + a)
            a = a[2:]
            if '=' in a:
                k, v = a.split('=', 1)
            elif a in self._defaults and isinstance(self._defaults[a], bool):
                k, v = a, 'true'
            elif a.startswith('no') and a[2:] in self._defaults and isinstance(self._defaults[a[2:]], bool):
                k, v = a[2:], 'false'
            else:
                k, v = a, argv[i + 1]
                i += 1
            if k not in self._defaults:
                raise ValueError('unknown flag: --' + k)
            d = self._defaults[k]
            if isinstance(d, bool):
                v = v.lower() in ('1', 'true', 'yes')
            elif isinstance(d, int):
                v = int(v)
            elif isinstance(d, float):
                v = float(v)
            self._values[k] = v
            i += 1


FLAGS = _Flags()


def DEFINE_string(name, default, helpstr=''):
    FLAGS._define(name, default, helpstr)


def DEFINE_integer(name, default, helpstr=''):
    FLAGS._define(name, default, helpstr)


def DEFINE_float(name, default, helpstr=''):
    FLAGS._define(name, float(default) if default is not None else None, helpstr)


def DEFINE_boolean(name, default, helpstr=''):
    FLAGS._define(name, bool(default), helpstr)


# flags every run script declares (nets/resnet_at_cifar10_run.py:27-31)
DEFINE_string('log_dir', './logs', 'logging directory')
DEFINE_boolean('enbl_multi_gpu', False, 'enable multi-GPU training')
DEFINE_string('learner', 'full-prec', 'learner\'s name')
DEFINE_string('exec_mode', 'train', 'execution mode: train / eval')
DEFINE_boolean('debug', False, 'debugging information')
