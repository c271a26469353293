"""CPU oracle of one full compression-aware training step, op by op and un-fused, in PyTorch-CPU
fp32 — the stand-in for the reference's TF-CPU learner (TensorFlow 1.x is not importable here;
SURVEY.md §8c/§8d).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product path.  PARITY UNPINNED (see oracle/pf_oracle.py).

It interprets the same layer graph the ModelHelper builds, with the reference's op decomposition:
separate reduce_max / reduce_min / sub / div / mul / round / div / mul / add for every fake-quant
(learners/uniform_quantization/utils.py:163-245), F.conv2d for tf.nn.conv2d, batch-norm with batch
statistics, per-variable l2_loss terms, tf.losses.softmax_cross_entropy, the distillation term
(learners/distillation_helper.py:98-100), autograd for compute_gradients with the STE overrides, and
per-variable optimizer updates (oracle/pf_oracle.py: adam_step / momentum_step).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pf_oracle as O

F32 = np.float32


class _RoundSTE(torch.autograd.Function):
    """tf.round under gradient_override_map({'Round': 'Identity'}) (utils.py:185-186)."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)          # half-to-even, like tf.round

    @staticmethod
    def backward(ctx, g):
        return g


def uq_k(bits):
    return torch.tensor(float(O.uq_k(bits)), dtype=torch.float32)


def fake_quant(x, bits, axis=None):
    """__uniform_quantize + __scale + __inv_scale, op by op; min/max under stop_gradient."""
    with torch.no_grad():
        if axis is None:
            w_max, w_min = x.max(), x.min()
        else:
            w_max, w_min = x.max(dim=axis).values, x.min(dim=axis).values
    eps = torch.tensor(1e-10, dtype=torch.float32)
    alpha = w_max - w_min + eps
    beta = w_min
    xn = (x - beta) / alpha
    k = uq_k(bits)
    q = _RoundSTE.apply(xn * k) / k
    return alpha * q + beta


def weight_fake_quant(w, bits, use_buckets, bucket_type, bucket_size):
    if not use_buckets:
        return fake_quant(w, bits)
    shape = w.shape
    if bucket_type == 'channel':
        return fake_quant(w.reshape(-1, shape[-1]), bits, axis=0).reshape(shape)
    flat = w.reshape(-1)
    n = flat.shape[0]
    multiple, rest = divmod(n, bucket_size)
    if rest:
        flat = torch.cat([flat, torch.ones(bucket_size - rest) * flat[-1]])
    q = fake_quant(flat.reshape(bucket_size, -1), bits, axis=0).reshape(-1)
    return q[:n].reshape(shape)


def codebook_quant(w, clusters):
    """__nonuni_quantize (nonuniform_quantization/utils.py:168-194, 284-307), 'weights' mode."""
    with torch.no_grad():
        w_max, w_min = w.max(), w.min()
    alpha = w_max - w_min + torch.tensor(1e-10)
    beta = w_min
    xn = (w - beta) / alpha
    c = clusters if torch.is_tensor(clusters) else torch.as_tensor(clusters, dtype=torch.float32)
    with torch.no_grad():
        idx = torch.argmin(torch.abs(xn.unsqueeze(-1) - c), dim=-1)
        sgn = torch.sign(xn + 1e-6)
    # gradient_override_map {'Mul': 'Add', 'Sign': 'Identity'} (utils.py:303-306): the upstream gradient goes unchanged
    # to BOTH factors — to tf.gather(c, min_index) (a segment sum into the codebook) and, through Sign-as-Identity, to x_n
    q = c[idx] * sgn + (xn - xn.detach())
    return alpha * q + beta


def _conv(x, w, attrs):
    (sh, sw), (pt, pl) = attrs['strides'], attrs['pad']
    kh, kw = attrs['ksize']
    n, h, wd, c = x.shape
    xt = x.permute(0, 3, 1, 2)
    wt = w.permute(3, 2, 0, 1)
    # trailing pads implied by the output size
    return xt, wt, (sh, sw), (pt, pl), (kh, kw)


def conv2d_nhwc(x, w, attrs, out_shape):
    xt, wt, (sh, sw), (pt, pl), (kh, kw) = _conv(x, w, attrs)
    p, q = out_shape[1], out_shape[2]
    pb = max((p - 1) * sh + kh - x.shape[1] - pt, 0)
    pr = max((q - 1) * sw + kw - x.shape[2] - pl, 0)
    xt = F.pad(xt, (pl, pr, pt, pb))
    y = F.conv2d(xt, wt, stride=(sh, sw))
    return y[:, :, :p, :q].permute(0, 2, 3, 1).contiguous()


class StepOracle:
    """Interprets `ops` (a topologically ordered op list of pocketflow_b200.graph) on the CPU."""

    def __init__(self, ops, logits_t, images_t, labels_t=None, loss=None, weight_quant=None,
                 act_quant=None, teacher=None, threads=None):
        self.ops, self.logits_t, self.images_t, self.labels_t = ops, logits_t, images_t, labels_t
        self.loss, self.teacher = loss, teacher
        self.wq = weight_quant or {}
        self.aq = act_quant or {}
        self.wq_bits = dict(zip([o.name for o in self.wq.get('ops', [])], self.wq.get('bits', [])))
        self.aq_bits = dict(zip([o.name for o in self.aq.get('ops', [])], self.aq.get('bits', [])))
        self.clusters = {}
        if threads:
            torch.set_num_threads(threads)

    def forward(self, params, images, training=True, stats_out=None, force=None, local_out=None):
        """params: name -> torch tensor.  Returns dict tensor-name -> value (NHWC).
        force / local_out (layer-local parity tests): after an op has been evaluated its result is recorded in
        local_out[name] and then REPLACED by force[name] when given — every op is thus applied to inputs produced by
        the implementation under test, so a difference cannot be amplified by the quantizers downstream."""
        val = {self.images_t.name: images}
        for op in self.ops:
            ty = op.type
            if ty == 'Placeholder':
                continue
            x = val[op.inputs[0].name] if op.inputs else None
            if ty in ('Conv2D', 'MatMul'):
                w = params[op.vars['kernel'].name]
                if op.name in self.wq_bits:
                    if self.wq.get('kind', 'uniform') == 'uniform':
                        w = weight_fake_quant(w, self.wq_bits[op.name], self.wq.get('use_buckets', False),
                                              self.wq.get('bucket_type', 'channel'), self.wq.get('bucket_size', 256))
                    else:
                        # the codebook: the op's `clusters` variable when the graph carries one, else set by the test
                        cv = op.vars.get('clusters')
                        w = codebook_quant(w, params[cv.name] if cv is not None and cv.name in params
                                           else self.clusters[op.name])
                y = conv2d_nhwc(x, w, op.attrs, op.output.shape) if ty == 'Conv2D' else x @ w
                if 'bias' in op.vars:
                    y = y + params[op.vars['bias'].name]
            elif ty == 'DepthwiseConv2dNative':
                w = params[op.vars['kernel'].name]
                if op.name in self.wq_bits:
                    w = weight_fake_quant(w, self.wq_bits[op.name], self.wq.get('use_buckets', False),
                                          self.wq.get('bucket_type', 'channel'), self.wq.get('bucket_size', 256))
                (sh, sw), (pt, pl), (kh, kw) = op.attrs['strides'], op.attrs['pad'], op.attrs['ksize']
                p, q = op.output.shape[1], op.output.shape[2]
                pb = max((p - 1) * sh + kh - x.shape[1] - pt, 0)
                pr = max((q - 1) * sw + kw - x.shape[2] - pl, 0)
                xt = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
                y = F.conv2d(xt, w.permute(2, 3, 0, 1), stride=(sh, sw), groups=x.shape[-1])
                y = y[:, :, :p, :q].permute(0, 2, 3, 1).contiguous()
            elif ty == 'FusedBatchNorm':
                ga, be = params[op.vars['gamma'].name], params[op.vars['beta'].name]
                mm, mv = params[op.vars['moving_mean'].name], params[op.vars['moving_variance'].name]
                eps, mom = op.attrs['epsilon'], op.attrs['momentum']
                if op.attrs['training'] and training:
                    red = tuple(range(x.dim() - 1))
                    mean = x.mean(dim=red)
                    var = ((x - mean) ** 2).mean(dim=red)
                    y = (x - mean) * torch.rsqrt(var + eps) * ga + be
                    if stats_out is not None:
                        m = x.numel() // x.shape[-1]
                        with torch.no_grad():
                            stats_out[op.vars['moving_mean'].name] = mm * mom + mean * (1 - mom)
                            stats_out[op.vars['moving_variance'].name] = mv * mom + var * (m / max(m - 1, 1)) * (1 - mom)
                else:
                    y = (x - mm) * torch.rsqrt(mv + eps) * ga + be
            elif ty in ('Relu', 'Relu6'):
                y = torch.relu(x) if ty == 'Relu' else torch.clamp(x, 0.0, 6.0)
                if op.name in self.aq_bits:
                    y = fake_quant(y, self.aq_bits[op.name])
            elif ty == 'MaxPool':
                (kh, kw), (sh, sw), (pt, pl) = op.attrs['ksize'], op.attrs['strides'], op.attrs['pad']
                p, q = op.output.shape[1], op.output.shape[2]
                pb = max((p - 1) * sh + kh - x.shape[1] - pt, 0)
                pr = max((q - 1) * sw + kw - x.shape[2] - pl, 0)
                xt = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float('-inf'))
                y = F.max_pool2d(xt, (kh, kw), (sh, sw)).permute(0, 2, 3, 1).contiguous()
            elif ty == 'Mean':
                y = x.mean(dim=(1, 2), keepdim=len(op.output.shape) == 4)
            elif ty == 'Reshape':
                y = x.reshape(op.output.shape)
            elif ty == 'Identity':
                y = x
            elif ty == 'Add':
                y = x + val[op.inputs[1].name]
            elif ty == 'Softmax':
                y = torch.softmax(x, dim=-1)
            else:
                raise NotImplementedError(ty)
            if local_out is not None:
                local_out[op.output.name] = y
            if force is not None and op.output.name in force:
                y = force[op.output.name]
            val[op.output.name] = y
        return val

    def step(self, state, images, labels, optimizer, lr, teacher_state=None, masks=None, grad_scale=1.0,
             beta_powers=(0.9, 0.999), frozen=()):
        """One training step from `state` (name -> np array).  Returns (losses dict, new_state, grads).
        frozen: names of trainable variables the optimizer leaves alone (var_list of compute_gradients,
        nonuniform_quantization/learner.py:252-270); they still count in the l2 loss."""
        params = {}
        for k, v in state.items():
            t = torch.from_numpy(np.array(v, dtype=F32, copy=True))
            params[k] = t
        train_names = [v.name for v in self._trainable()]
        for n in train_names:
            params[n].requires_grad_(True)
        x = torch.from_numpy(np.asarray(images, F32))
        lab = torch.from_numpy(np.asarray(labels, F32))
        stats = {}
        val = self.forward(params, x, True, stats)
        logits = val[self.loss.ce[1].name]
        # tf.losses.softmax_cross_entropy: batch mean of -sum(labels * log_softmax)
        hard = (-(lab * torch.log_softmax(logits, dim=-1)).sum(-1)).mean()
        l2 = torch.zeros(())
        for v, coef in self.loss.l2.items():
            l2 = l2 + coef * (params[v.name] ** 2).sum() / 2
        total = hard + l2
        dst = torch.zeros(())
        if self.loss.dst is not None:
            tparams = {k: torch.from_numpy(np.array(v, dtype=F32, copy=True)) for k, v in teacher_state.items()}
            with torch.no_grad():
                tl = self.teacher.forward(tparams, x, False)[self.teacher.logits_t.name]
            w, T = self.loss.dst[2], self.loss.dst[3]
            soft = torch.softmax(tl / T, dim=-1)
            dst = w * (-(soft * torch.log_softmax(val[self.loss.dst[0].name] / T, dim=-1)).sum(-1)).mean()
            total = total + dst
        # the l2 gradient is added analytically below, exactly like the fused optimizer (g*scale + wd*w)
        data_loss = hard + dst
        grads = torch.autograd.grad(data_loss, [params[n] for n in train_names], allow_unused=True)
        new_state = {k: np.array(v, dtype=F32, copy=True) for k, v in state.items()}
        gout = {}
        wd_of = {v.name: c for v, c in self.loss.l2.items()}
        for n, g in zip(train_names, grads):
            g = np.zeros_like(state[n]) if g is None else g.numpy()
            gout[n] = g
            if n in frozen:
                continue
            wd = wd_of.get(n, 0.0)
            if optimizer['kind'] == 'adam':
                m0 = optimizer['slots'].setdefault(n + '/m', np.zeros_like(state[n]))
                v0 = optimizer['slots'].setdefault(n + '/v', np.zeros_like(state[n]))
                w1, m1, v1 = O.adam_step(state[n], m0, v0, g, lr, F32(beta_powers[0]), F32(beta_powers[1]),
                                         optimizer.get('beta1', 0.9), optimizer.get('beta2', 0.999),
                                         optimizer.get('eps', 1e-8), wd, grad_scale)
                optimizer['slots'][n + '/m'], optimizer['slots'][n + '/v'] = m1, v1
            else:
                a0 = optimizer['slots'].setdefault(n + '/acc', np.zeros_like(state[n]))
                mk = masks.get(n) if masks else None
                w1, a1 = O.momentum_step(state[n], a0, g, lr, optimizer.get('momentum', 0.9), mk, wd, grad_scale)
                optimizer['slots'][n + '/acc'] = a1
            new_state[n] = w1
        for k, v in stats.items():
            new_state[k] = v.numpy().astype(F32)
        with torch.no_grad():
            top1 = (logits.argmax(-1) == lab.argmax(-1)).float().mean()
        losses = dict(ce=F32(hard.item()), l2=F32(l2.item()), dst_loss=F32(dst.item()),
                      model_loss=F32(hard.item() + l2.item()), loss=F32(total.item()), acc_top1=F32(top1.item()))
        return losses, new_state, gout

    def _trainable(self):
        out = []
        for op in self.ops:
            for v in op.vars.values():
                if v.trainable and v not in out:
                    out.append(v)
        return out


def cpg_layer_regression(orc_full, orc_prnd, state_full, state_prnd, images, conv_full, conv_prnd, training=True):
    """reg_loss_i = tf.nn.l2_loss(conv_i(full model) - conv_i(pruned model)) on one mini-batch and its gradient w.r.t.
    the pruned model's kernel of that layer (/root/reference/learners/channel_pruning_gpu/learner.py:339-354, :370);
    both models in training mode (forward_train), only the pruned model's BN moving statistics are updated (:283-286).
    conv_full / conv_prnd: the two Conv2D ops (pocketflow_b200.graph).  Returns (loss, grad, new pruned-model stats).
    training=False: both networks in inference mode — the layer-wise regression of the weight-sparsification learner's
    pruning-ratio search (learners/weight_sparsification/pr_optimizer.py:166-181 forward_eval, :298-299)."""
    x = torch.from_numpy(np.asarray(images, F32))
    pf = {k: torch.from_numpy(np.array(v, dtype=F32, copy=True)) for k, v in state_full.items()}
    pp = {k: torch.from_numpy(np.array(v, dtype=F32, copy=True)) for k, v in state_prnd.items()}
    kname = conv_prnd.vars['kernel'].name
    pp[kname].requires_grad_(True)
    with torch.no_grad():
        out_f = orc_full.forward(pf, x, training)[conv_full.output.name]
    stats = {}
    out_p = orc_prnd.forward(pp, x, training, stats)[conv_prnd.output.name]
    loss = ((out_f - out_p) ** 2).sum() / 2
    grad, = torch.autograd.grad(loss, [pp[kname]])
    return F32(loss.item()), grad.numpy().astype(F32), {k: v.numpy().astype(F32) for k, v in stats.items()}
