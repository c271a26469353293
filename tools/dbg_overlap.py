import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import test_step_gpu as T
res = {}
for name, overlap, graph, prefetch in (("serial", "0", False, "1"), ("ovl", "1", False, "1"), ("ovl_graph", "1", True, "1")):
                                       ('ovl', '1', False, '1'), ('ovl_nopf', '1', False, '0'), ('ovl_graph', '1', True, '1')):
    os.environ['PF_OVERLAP'] = overlap; os.environ['PF_CONV_PATH'] = 'tc'; os.environ['PF_INPUT_PREFETCH'] = prefetch
    lrn = T.make_uq_learner(resnet_size=20, batch=32, dst=True)
    ex = lrn.sess_train
    lrn.iterator_train.prefill()
    if graph:
        P0, O0 = ex.store.P.clone(), ex.store.O.clone()
        lrn.feed(ex, lrn.iterator_train)
        ex.capture()
        ex.store.P.copy_(P0); ex.store.O.copy_(O0); ex.S1.zero_(); ex.S2.zero_()
        lrn.iterator_train.cursor = 0
        lrn.iterator_train._staging = None
    losses = []
    for _ in range(3):
        lrn.train_step()
        losses.append(float(ex.fetch_losses()['loss']))
    torch.cuda.synchronize()
    res[name] = (ex.store.P.clone(), losses)
    print(name, losses)
base = res['serial'][0]
for k, (P, l) in res.items():
    print(k, 'equal' if torch.equal(P, base) else 'DIFF max %.3e' % (P - base).abs().max().item())
