"""ResNet v1/v2 graph builder — the architecture of the TF-official model the reference vendors
(/root/reference/utils/external/resnet_model.py:55-554), re-expressed on pocketflow_b200.graph.
NHWC only (the quantizers re-create convs without data_format)."""
from .. import graph as G

_BATCH_NORM_DECAY = 0.997
_BATCH_NORM_EPSILON = 1e-5
DEFAULT_VERSION = 2


def batch_norm(inputs, training):
    return G.batch_normalization(inputs, training, momentum=_BATCH_NORM_DECAY, epsilon=_BATCH_NORM_EPSILON)


def conv2d_fixed_padding(inputs, filters, kernel_size, strides):
    """strides > 1: explicit symmetric-ish padding then VALID; else SAME (resnet_model.py:92-103)."""
    if strides > 1:
        pad_total = kernel_size - 1
        pad_beg = pad_total // 2
        padding = ((pad_beg, pad_total - pad_beg), (pad_beg, pad_total - pad_beg))
    else:
        padding = 'same'
    return G.conv2d(inputs, filters, kernel_size, strides, padding, use_bias=False,
                    kernel_initializer=G.variance_scaling_initializer())


def _building_block_v2(inputs, filters, training, projection_shortcut, strides):
    shortcut = inputs
    inputs = batch_norm(inputs, training)
    inputs = G.relu(inputs)
    if projection_shortcut is not None:
        shortcut = projection_shortcut(inputs)
    inputs = conv2d_fixed_padding(inputs, filters, 3, strides)
    inputs = batch_norm(inputs, training)
    inputs = G.relu(inputs)
    inputs = conv2d_fixed_padding(inputs, filters, 3, 1)
    return inputs + shortcut


def _bottleneck_block_v2(inputs, filters, training, projection_shortcut, strides):
    shortcut = inputs
    inputs = batch_norm(inputs, training)
    inputs = G.relu(inputs)
    if projection_shortcut is not None:
        shortcut = projection_shortcut(inputs)
    inputs = conv2d_fixed_padding(inputs, filters, 1, 1)
    inputs = batch_norm(inputs, training)
    inputs = G.relu(inputs)
    inputs = conv2d_fixed_padding(inputs, filters, 3, strides)
    inputs = batch_norm(inputs, training)
    inputs = G.relu(inputs)
    inputs = conv2d_fixed_padding(inputs, 4 * filters, 1, 1)
    return inputs + shortcut


def block_layer(inputs, filters, bottleneck, block_fn, blocks, strides, training, name):
    filters_out = filters * 4 if bottleneck else filters

    def projection_shortcut(x):
        return conv2d_fixed_padding(x, filters_out, 1, strides)

    inputs = block_fn(inputs, filters, training, projection_shortcut, strides)
    for _ in range(1, blocks):
        inputs = block_fn(inputs, filters, training, None, 1)
    return G.identity(inputs, name)


class Model(object):
    def __init__(self, resnet_size, bottleneck, num_classes, num_filters, kernel_size, conv_stride,
                 first_pool_size, first_pool_stride, block_sizes, block_strides,
                 resnet_version=DEFAULT_VERSION, data_format=None):
        if resnet_version != 2:
            raise ValueError('only ResNet v2 (the reference default, resnet_model.py:46) is implemented')
        if data_format not in (None, 'channels_last'):
            raise ValueError('NHWC (channels_last) only')
        self.resnet_size = resnet_size
        self.bottleneck = bottleneck
        self.block_fn = _bottleneck_block_v2 if bottleneck else _building_block_v2
        self.num_classes, self.num_filters = num_classes, num_filters
        self.kernel_size, self.conv_stride = kernel_size, conv_stride
        self.first_pool_size, self.first_pool_stride = first_pool_size, first_pool_stride
        self.block_sizes, self.block_strides = block_sizes, block_strides

    def __call__(self, inputs, training):
        with G.variable_scope('resnet_model'):
            inputs = conv2d_fixed_padding(inputs, self.num_filters, self.kernel_size, self.conv_stride)
            inputs = G.identity(inputs, 'initial_conv')
            if self.first_pool_size:
                inputs = G.max_pooling2d(inputs, self.first_pool_size, self.first_pool_stride, padding='same')
                inputs = G.identity(inputs, 'initial_max_pool')
            for i, num_blocks in enumerate(self.block_sizes):
                num_filters = self.num_filters * (2 ** i)
                inputs = block_layer(inputs, num_filters, self.bottleneck, self.block_fn, num_blocks,
                                     self.block_strides[i], training, 'block_layer{}'.format(i + 1))
            inputs = batch_norm(inputs, training)
            inputs = G.relu(inputs)
            inputs = G.reduce_mean_hw(inputs)
            inputs = G.identity(inputs, 'final_reduce_mean')
            inputs = G.dense(inputs, self.num_classes)
            inputs = G.identity(inputs, 'final_dense')
            return inputs
