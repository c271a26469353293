"""(type, name) op lists of the training graphs this repo's ModelHelpers emit (teacher under 'distilled_model'), the
input of the reference's op-search functions in make_golden_from_reference.py and of the test that compares."""
import importlib


def build_graph(net, flags, dst=True):
    from pocketflow_b200 import graph as G
    from pocketflow_b200.flags import FLAGS
    FLAGS.reset()
    mod = importlib.reload(importlib.import_module('pocketflow_b200.nets.' + net))   # re-DEFINE this net's flag defaults
    for k, v in flags.items():
        setattr(FLAGS, k, v)
    mh = mod.ModelHelper()
    g = G.Graph()
    with g.as_default():
        with G.variable_scope('data'):
            it = mh.build_dataset_train()
            im, lab = it.get_next()
        if dst:
            with G.variable_scope('distilled_model'):
                mh.forward_eval(im)
        with G.variable_scope('model'):
            mh.forward_train(im)
    return g


GRAPHS = {
    'resnet20_cifar10_dst': ('resnet_at_cifar10', dict(resnet_size=20, batch_size=2), True),
    'mobilenet_v1_ilsvrc12': ('mobilenet_at_ilsvrc12', dict(batch_size=2, nb_classes=1001), False),
    'lenet_cifar10': ('lenet_at_cifar10', dict(batch_size=2), False),
}


def op_lists():
    return {name: [(op.type, op.name) for op in build_graph(net, flags, dst).ops] for name, (net, flags, dst) in GRAPHS.items()}
