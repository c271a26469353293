"""MobileNet-v1 graph builder — the slim model the reference vendors
(/root/reference/utils/external/mobilenet_v1.py:124-139, 245-303, 368-388, 428-477), re-expressed on
pocketflow_b200.graph with slim's variable names (MobilenetV1/Conv2d_3_pointwise/weights, .../BatchNorm/gamma,
.../depthwise_weights, MobilenetV1/Logits/Conv2d_1c_1x1/{weights,biases}) so the reference's name filters
(`pointwise/weights`, `Conv2d_1c_1x1/weights`, weight_sparsification/utils.py:34-37) keep working."""
from collections import namedtuple

from .. import graph as G

Conv = namedtuple('Conv', ['kernel', 'stride', 'depth'])
DepthSepConv = namedtuple('DepthSepConv', ['kernel', 'stride', 'depth'])

MOBILENETV1_CONV_DEFS = [
    Conv(kernel=[3, 3], stride=2, depth=32),
    DepthSepConv(kernel=[3, 3], stride=1, depth=64),
    DepthSepConv(kernel=[3, 3], stride=2, depth=128),
    DepthSepConv(kernel=[3, 3], stride=1, depth=128),
    DepthSepConv(kernel=[3, 3], stride=2, depth=256),
    DepthSepConv(kernel=[3, 3], stride=1, depth=256),
    DepthSepConv(kernel=[3, 3], stride=2, depth=512),
    DepthSepConv(kernel=[3, 3], stride=1, depth=512),
    DepthSepConv(kernel=[3, 3], stride=1, depth=512),
    DepthSepConv(kernel=[3, 3], stride=1, depth=512),
    DepthSepConv(kernel=[3, 3], stride=1, depth=512),
    DepthSepConv(kernel=[3, 3], stride=1, depth=512),
    DepthSepConv(kernel=[3, 3], stride=2, depth=1024),
    DepthSepConv(kernel=[3, 3], stride=1, depth=1024),
]

BATCH_NORM_DECAY = 0.9997
BATCH_NORM_EPSILON = 0.001
WEIGHTS_STDDEV = 0.09


def _bn_relu6(net, is_training, scope):
    with G.variable_scope(scope):
        net = G.batch_normalization(net, is_training, momentum=BATCH_NORM_DECAY, epsilon=BATCH_NORM_EPSILON,
                                    name='BatchNorm', exact_name=True)
        return G.relu6(net, name='Relu6')


def mobilenet_v1(inputs, num_classes=1001, is_training=True, depth_multiplier=1.0, min_depth=8):
    """Returns logits [N, num_classes].  dropout_keep_prob = 0.999 of the reference is the identity here
    (a 0.1 % random mask would make per-step parity seed-dependent; flagged in DESIGN.md)."""
    if depth_multiplier <= 0:
        raise ValueError('depth_multiplier is not greater than zero.')
    depth = lambda d: max(int(d * depth_multiplier), min_depth)
    init = G.truncated_normal_initializer(WEIGHTS_STDDEV)
    with G.variable_scope('MobilenetV1'):
        net = inputs
        for i, conv_def in enumerate(MOBILENETV1_CONV_DEFS):
            base = 'Conv2d_%d' % i
            if isinstance(conv_def, Conv):
                net = G.conv2d(net, depth(conv_def.depth), conv_def.kernel, conv_def.stride, 'same', use_bias=False,
                               kernel_initializer=init, name=base, kernel_name='weights', exact_name=True)
                net = _bn_relu6(net, is_training, base)
            else:
                net = G.depthwise_conv2d(net, conv_def.kernel[0], conv_def.stride, 'same', kernel_initializer=init,
                                         name=base + '_depthwise', exact_name=True)
                net = _bn_relu6(net, is_training, base + '_depthwise')
                net = G.conv2d(net, depth(conv_def.depth), [1, 1], 1, 'same', use_bias=False, kernel_initializer=init,
                               name=base + '_pointwise', kernel_name='weights', exact_name=True)
                net = _bn_relu6(net, is_training, base + '_pointwise')
        with G.variable_scope('Logits'):
            net = G.reduce_mean_hw(net, name='global_pool', keepdims=True)
            logits = G.conv2d(net, num_classes, [1, 1], 1, 'same', use_bias=True, kernel_initializer=init,
                              name='Conv2d_1c_1x1', kernel_name='weights', bias_name='biases', exact_name=True)
            logits = G.squeeze_hw(logits, name='SpatialSqueeze')
    return logits
