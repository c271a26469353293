"""CPU oracle for the PocketFlow compression-aware training hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; the product path (``pocketflow_b200``) never does and fails loudly when
``libpf_b200.so`` is missing.

PARITY UNPINNED in the strict sense — but the STRUCTURE of uniform_quantize, nonuniform_quantize, ws_build_mask,
ws_prune_ratio_dyn, distillation_loss and the schedules IS pinned bit-for-bit against the reference's own Python, executed
from /root/reference on numpy-backed stub tensors (tests/golden/make_golden_from_reference.py): the reference (Tencent/PocketFlow @53b82cb) ships no golden
vectors, known-answer tests or fixtures for this path, and its arithmetic lives
in TensorFlow 1.x, which cannot be imported in this image.  This file restates
the reference's op chains literally, in the reference's op ORDER, in numpy
float32 (numpy never contracts mul+add into FMA), and pins itself with
hand-derivable KATs (tests/test_oracle_kat.py).  TF-internal semantics that the
restatement assumes (round-half-even, percentile 'nearest' index rule, argmin
first-index ties, TF-Adam epsilon placement) are each documented at the
function that encodes them.

All citations are file:line into /root/reference.
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# uniform quantization  (learners/uniform_quantization/utils.py:163-289)
# ----------------------------------------------------------------------------
def uq_k(bits):
    """k = tf.cast(2 ** mbits - 1, tf.float32), mbits int64 (utils.py:184).

    bits=8 -> 255.0 ; bits=32 -> 4294967295 rounds to 4294967296.0f."""
    return F32(np.int64(2) ** np.int64(bits) - np.int64(1))


def split_bucket(w, bucket_size):
    """utils.py:247-274.  Flatten, pad with copies of the LAST element up to a
    multiple of bucket_size, reshape([bucket_size, -1]).  Bucket j (column j)
    therefore holds flat[i*multiple + j]: a STRIDED group, not a chunk."""
    flat = np.ascontiguousarray(w, dtype=F32).reshape(-1)
    num_w = flat.shape[0]
    multiple, rest = divmod(num_w, bucket_size)
    if rest != 0:
        flat = np.concatenate([flat, np.ones(bucket_size - rest, F32) * flat[-1]])
        multiple += 1
    padded = (bucket_size - rest) if rest != 0 else 0
    return flat.reshape(bucket_size, -1), multiple, padded


def channel_bucket(w):
    """utils.py:276-289: reshape([-1, cout]); one bucket per output channel."""
    cout = w.shape[-1]
    return np.ascontiguousarray(w, dtype=F32).reshape(-1, cout), cout, 0


def uq_scale(w, axis):
    """utils.py:201-231.  alpha = max - min + 1e-10 ; beta = min ;
    (w - beta) / alpha with true fp32 division."""
    w_max = np.max(w, axis=axis)
    w_min = np.min(w, axis=axis)
    eps = F32(1e-10)
    alpha = (w_max - w_min).astype(F32) + eps
    beta = w_min.astype(F32)
    return ((w - beta) / alpha).astype(F32), alpha.astype(F32), beta


def uq_inv_scale(q, alpha, beta):
    """utils.py:233-245: alpha * w + beta as a separate Mul and Add."""
    return ((alpha * q).astype(F32) + beta).astype(F32)


def uniform_quantize(x, bits, mode='weight', use_buckets=False, bucket_type='channel',
                     bucket_size=256, return_scales=False):
    """UniformQuantization.__uniform_quantize (utils.py:163-199).

    tf.round is round-half-to-even == np.rint.  Returns the fake-quantized
    tensor with x's shape."""
    x = np.ascontiguousarray(x, dtype=F32)
    orig_shape = x.shape
    padded = 0
    if use_buckets and mode == 'weight':
        if bucket_type == 'split':
            xb, _, padded = split_bucket(x, bucket_size)
        elif bucket_type == 'channel':
            xb, _, padded = channel_bucket(x)
        else:
            raise ValueError("Unrecognized bucket type, must be 'weight' or 'channel'.")
        axis = 0
    else:
        if mode not in ('weight', 'activation'):
            raise ValueError('Unknown mode for scalling')
        xb, axis = x, None
    xn, alpha, beta = uq_scale(xb, axis)
    k = uq_k(bits)
    q = (np.rint((xn * k).astype(F32)) / k).astype(F32)
    qw = uq_inv_scale(q, alpha, beta)
    if use_buckets and mode == 'weight':
        qw = qw.reshape(-1)
        if padded:
            qw = qw[:-padded]
    qw = qw.reshape(orig_shape)
    if return_scales:
        return qw, alpha, beta
    return qw


def uq_ste_grad(g, alpha, bits):
    """STE backward (utils.py:185-186, 224-225): Round->Identity, min/max under
    stop_gradient, so d qw/d w is the chain of the surviving linear ops:
    g*alpha (Mul grad) -> /k (RealDiv grad) -> *k (Mul grad) -> /alpha (RealDiv grad)."""
    k = uq_k(bits)
    g = np.asarray(g, F32)
    t = (g * alpha).astype(F32)
    t = (t / k).astype(F32)
    t = (t * k).astype(F32)
    return (t / alpha).astype(F32)


def bucket_storage_bits(bucket_num):
    """utils.py:299-306: alpha and beta, 32 bits each, per bucket."""
    return int(bucket_num) * 32 * 2


# ----------------------------------------------------------------------------
# percentile (tf.contrib.distributions.percentile, interpolation='nearest')
# ----------------------------------------------------------------------------
def percentile_index(n, q):
    """Index into the DESCENDING sort that TF-1.12 percentile('nearest') gathers
    (recollection of contrib/distributions/python/ops/sample_stats.py):
        q = to_double(q); frac = 1 - q/100; d = to_double(n)
        idx = clip(int32(round_half_even((d-1)*frac)), 0, n-1)"""
    qd = float(q)
    frac = 1.0 - qd / 100.0
    idx = int(np.rint((float(n) - 1.0) * frac))
    return min(max(idx, 0), n - 1)


def percentile_nearest(x, q, axis=None):
    """percentile of x at q (0..100), 'nearest' interpolation, via a full
    descending sort (= nn.top_k(k=n))."""
    x = np.asarray(x)
    if axis is None:
        y = x.reshape(-1)
        idx = percentile_index(y.shape[0], q)
        return np.sort(y)[::-1][idx]
    assert axis == 0
    n = x.shape[0]
    idx = percentile_index(n, q)
    return np.sort(x, axis=0)[::-1][idx]


# ----------------------------------------------------------------------------
# channel pruning, GPU variant  (learners/channel_pruning_gpu/learner.py:250-260, 356-402, 445-518)
# ----------------------------------------------------------------------------
def cpg_group_norms(w):
    """var_norm = sqrt(reduce_sum(square(var), axis=[0, 1, 3], keepdims=True)) (learner.py:255, :379) of a kernel
    [R,S,Cin,Cout] — one norm per INPUT channel, float32.  The order of TF's reduction is not specified: this is numpy's
    float32 sum (what the golden generator's stub evaluates the reference's op with); device results are compared with a
    tolerance, the SET of zeroed channels exactly."""
    w = np.asarray(w, F32)
    return np.sqrt(np.sum(w * w, axis=(0, 1, 3), dtype=F32)).astype(F32)


def cpg_prox_step(w, g, lrn_rate_pgd, prune_perctl):
    """ops['prune'] of one layer (learner.py:375-383):
        var_new = var - lr * grad ; var_norm = ||var_new||_{[0,1,3]} ; threshold = percentile(var_norm, prune_perctl)
        shrk_vec = maximum(1 - threshold / var_norm, 0) ; var <- var_new * shrk_vec
    Returns (var, var_norm, threshold).  (0/0 gives NaN in TF's shrk_vec; this restatement returns 0 there.)"""
    w, g = np.asarray(w, F32), np.asarray(g, F32)
    var_new = (w - F32(lrn_rate_pgd) * g).astype(F32)
    var_norm = cpg_group_norms(var_new)
    threshold = F32(percentile_nearest(var_norm, F32(prune_perctl)))     # fed through a float32 placeholder (:365)
    with np.errstate(divide='ignore', invalid='ignore'):
        shrk = np.maximum(F32(1.0) - threshold / var_norm, F32(0.0)).astype(F32)
    shrk = np.where(np.isnan(shrk), F32(0.0), shrk).astype(F32)
    return (var_new * shrk[None, None, :, None]).astype(F32), var_norm, threshold


def cpg_channel_mask(w):
    """mask_updt_ops (learner.py:255-259): tile(cast(var_norm > 0, float32)) over the kernel's shape."""
    w = np.asarray(w, F32)
    keep = (cpg_group_norms(w) > 0).astype(F32)
    return np.broadcast_to(keep[None, None, :, None], w.shape).astype(F32).copy()


def cpg_selection_schedule(reg_losses, prune_ratio, nb_iters_layer, lr_init=1e-10, incr=1.4, decr=0.7):
    """The host loop of __choose_channels (learner.py:476-497) for one layer, given the regression losses it observed:
    [(lrn_rate_pgd, prune_perctl)] fed at every iteration.  lr grows when the loss fell (loss < loss_prev, loss_prev
    starting at 0.0), shrinks otherwise; the percentile ramps linearly to prune_ratio * 100."""
    out, lr, prev = [], lr_init, 0.0
    for it in range(nb_iters_layer):
        out.append((lr, prune_ratio * 100.0 * (it + 1) / nb_iters_layer))
        lr = lr * incr if reg_losses[it] < prev else lr * decr
        prev = reg_losses[it]
    return out


# ----------------------------------------------------------------------------
# weight sparsification  (learners/weight_sparsification/learner.py:260-332)
# ----------------------------------------------------------------------------
def ws_prune_ratio_dyn(global_step, nb_iters_train, prune_ratio_fnl,
                       iter_ratio_beg=0.1, iter_ratio_end=0.5, exponent=3.0):
    """__calc_prune_ratio_dyn (learner.py:296-312), float32 graph arithmetic."""
    idx_beg = int(nb_iters_train * iter_ratio_beg)
    idx_end = int(nb_iters_train * iter_ratio_end)
    base = F32(F32(int(global_step) - idx_beg) / F32(idx_end - idx_beg))
    base = F32(min(F32(1.0), max(F32(0.0), base)))
    one = F32(1.0)
    return F32(F32(prune_ratio_fnl) * F32(one - F32(np.power(F32(one - base), F32(exponent)))))


def ws_mask_rank(n, prune_ratio):
    """Descending-sort index of the threshold for a tensor of n elements:
    q = float32(prune_ratio) * 100 in float32, then percentile_index
    (learner.py:284)."""
    q = F32(F32(prune_ratio) * F32(100.0))
    return percentile_index(n, q)


def ws_build_mask(var, var_bkup, mask, prune_ratio):
    """One maskable variable of __build_masks (learner.py:281-288):
        bkup = where(mask > 0.5, var, bkup)
        thr  = percentile(|bkup|, prune_ratio*100)
        mask = float(|bkup| > thr)          (strict: ties at thr are pruned)
        var  = bkup * mask
    Returns (var, bkup, mask, thr)."""
    var = np.asarray(var, F32)
    var_bkup = np.where(np.asarray(mask, F32) > F32(0.5), var, np.asarray(var_bkup, F32)).astype(F32)
    q = F32(F32(prune_ratio) * F32(100.0))
    thr = percentile_nearest(np.abs(var_bkup), q)
    new_mask = (np.abs(var_bkup) > thr).astype(F32)
    new_var = (var_bkup * new_mask).astype(F32)
    return new_var, var_bkup, new_mask, F32(thr)


def ws_heurist_ratios(nb_params, prune_ratio):
    """PROptimizer 'heurist' protocol (pr_optimizer.py:394-409):
    ratio_i = alpha * log(n_i), alpha = s * sum(n_i) / sum(n_i log n_i)."""
    n = np.asarray(nb_params, dtype=np.float64)
    alpha = prune_ratio * np.sum(n) / np.sum(n * np.log(n))
    return alpha * np.log(n)


def calc_prune_ratio(vars_list):
    """learner.py:51-65: 1 - nnz/size over a list of variables, float32."""
    nnz = sum(int(np.count_nonzero(v)) for v in vars_list)
    tot = sum(int(np.size(v)) for v in vars_list)
    return F32(F32(1.0) - F32(nnz) / F32(tot))


def momentum_step(w, acc, g, lr, momentum, mask=None, wd=0.0, grad_scale=1.0):
    """Gradient assembly + tf.train.MomentumOptimizer.apply_gradients
    (learner.py:201-212, 314-332):
        g_tot = g*grad_scale + wd*w      (allreduce average; l2_loss gradient)
        g_tot = g_tot * mask             (__calc_grads_pruned)
        acc   = acc*momentum + g_tot ; w = w - lr*acc
    Separate fp32 mul/add everywhere."""
    w = np.asarray(w, F32)
    g = (np.asarray(g, F32) * F32(grad_scale)).astype(F32)
    if wd != 0.0:
        g = (g + (F32(wd) * w).astype(F32)).astype(F32)
    if mask is not None:
        g = (g * np.asarray(mask, F32)).astype(F32)
    acc = ((np.asarray(acc, F32) * F32(momentum)).astype(F32) + g).astype(F32)
    w = (w - (acc * F32(lr)).astype(F32)).astype(F32)
    return w, acc


def adam_alpha(lr, beta1_power, beta2_power):
    """alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t), all float32 (ApplyAdam)."""
    return F32(F32(F32(lr) * F32(np.sqrt(F32(F32(1.0) - F32(beta2_power))))) / F32(F32(1.0) - F32(beta1_power)))


def adam_step(w, m, v, g, lr, beta1_power, beta2_power, beta1=0.9, beta2=0.999,
              eps=1e-8, wd=0.0, grad_scale=1.0):
    """tf.train.AdamOptimizer as TF-1.x ApplyAdam computes it (recollection of
    core/kernels/training_ops.cc; uniform_quantization/learner.py:244):
        alpha = lr*sqrt(1-b2^t)/(1-b1^t)
        m += (g - m)*(1-b1) ; v += (g*g - v)*(1-b2)
        w -= (m*alpha) / (sqrt(v) + eps)       (eps OUTSIDE the bias correction)
    beta*_power are the running products (b^t) held as float32 variables."""
    w = np.asarray(w, F32)
    g = (np.asarray(g, F32) * F32(grad_scale)).astype(F32)
    if wd != 0.0:
        g = (g + (F32(wd) * w).astype(F32)).astype(F32)
    alpha = adam_alpha(lr, beta1_power, beta2_power)
    m = np.asarray(m, F32)
    v = np.asarray(v, F32)
    m = (m + ((g - m).astype(F32) * F32(F32(1.0) - F32(beta1))).astype(F32)).astype(F32)
    v = (v + (((g * g).astype(F32) - v).astype(F32) * F32(F32(1.0) - F32(beta2))).astype(F32)).astype(F32)
    w = (w - ((m * alpha).astype(F32) / (np.sqrt(v).astype(F32) + F32(eps)).astype(F32)).astype(F32)).astype(F32)
    return w, m, v


# ----------------------------------------------------------------------------
# losses  (learners/distillation_helper.py:86-103, nets/*_at_*.py calc_loss)
# ----------------------------------------------------------------------------
def softmax(x):
    x = np.asarray(x, F32)
    e = np.exp((x - np.max(x, axis=-1, keepdims=True)).astype(F32)).astype(F32)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32)


def softmax_xent_rows(labels, logits):
    """softmax_cross_entropy_with_logits_v2 per row, TF's xent functor order:
    z = x - max; lse = log(sum exp z); loss = sum(labels * (lse - z));
    backprop = exp(z)/sum - labels."""
    x = np.asarray(logits, F32)
    z = (x - np.max(x, axis=-1, keepdims=True)).astype(F32)
    e = np.exp(z).astype(F32)
    s = np.sum(e, axis=-1, keepdims=True, dtype=F32)
    lse = np.log(s).astype(F32)
    loss = np.sum((np.asarray(labels, F32) * (lse - z).astype(F32)).astype(F32), axis=-1, dtype=F32)
    backprop = ((e / s).astype(F32) - np.asarray(labels, F32)).astype(F32)
    return loss, backprop


def softmax_cross_entropy(labels, logits):
    """tf.losses.softmax_cross_entropy: batch MEAN of the row losses.
    Returns (loss, dloss/dlogits)."""
    rows, bp = softmax_xent_rows(labels, logits)
    n = rows.shape[0]
    return F32(np.sum(rows, dtype=F32) / F32(n)), (bp / F32(n)).astype(F32)


def distillation_loss(logits_pri, logits_dst, loss_w_dst=4.0, tempr_dst=4.0):
    """DistillationHelper.calc_loss (distillation_helper.py:86-103):
    w * softmax_cross_entropy(softmax(t/T), s/T).  Soft-label CROSS-ENTROPY (not
    KL), no T^2 factor.  Returns (loss, dloss/dlogits_pri)."""
    T = F32(tempr_dst)
    logits_soft = (np.asarray(logits_pri, F32) / T).astype(F32)
    labels_soft = softmax((np.asarray(logits_dst, F32) / T).astype(F32))
    loss, g = softmax_cross_entropy(labels_soft, logits_soft)
    return F32(F32(loss_w_dst) * loss), ((g * F32(loss_w_dst)).astype(F32) / T).astype(F32)


def l2_loss(v):
    """tf.nn.l2_loss: sum(v**2)/2."""
    v = np.asarray(v, F32)
    return F32(np.sum((v * v).astype(F32), dtype=F32) / F32(2.0))


def accuracy(labels_onehot, outputs):
    return F32(np.mean((np.argmax(labels_onehot, 1) == np.argmax(outputs, 1)).astype(F32)))


# ----------------------------------------------------------------------------
# non-uniform quantization  (learners/nonuniform_quantization/utils.py:168-366)
# ----------------------------------------------------------------------------
def nuq_quantile_init(x_normalized, nb_clusters, axis=None):
    """__quantile_init (utils.py:349-366): centroid j = percentile(x_n,
    (j+1)*100/(k+1)), 'nearest'.  q is an int64/int64 true division = float64."""
    cs = [percentile_nearest(x_normalized, (j + 1) * 100 / (nb_clusters + 1), axis=axis)
          for j in range(nb_clusters)]
    return np.asarray(cs, dtype=F32)


def nuq_assign(x_normalized, clusters):
    """utils.py:297-307: idx = argmin_j |x_n - c_j| (first index on ties);
    q = c[idx] * sign(x_n + 1e-6)."""
    xn = np.asarray(x_normalized, F32)
    c = np.asarray(clusters, F32)
    d = np.abs((xn[..., None] - c).astype(F32))
    idx = np.argmin(d, axis=-1)
    q = (c[idx] * np.sign((xn + F32(1e-6)).astype(F32))).astype(F32)
    return q, idx


def nonuniform_quantize(x, bits, clusters=None):
    """__nonuni_quantize (utils.py:168-194), no buckets, 'weight' mode: per-layer
    min/max normalise, quantile-initialised codebook (unless given), nearest
    centroid, inverse scale.  Returns (qx, clusters, idx)."""
    x = np.ascontiguousarray(x, dtype=F32)
    xn, alpha, beta = uq_scale(x, None)
    k = int(2 ** bits)
    if clusters is None:
        clusters = nuq_quantile_init(xn, k)
    q, idx = nuq_assign(xn, clusters)
    return uq_inv_scale(q, alpha, beta), clusters, idx


def nuq_grads(g, idx, nb_clusters, alpha):
    """STE of the codebook quantizer (utils.py:304-307, Mul->Add, Sign->Identity):
    d/dx_n = g_q (then /alpha through the normalisation), d/dc_j = sum_{idx=j} g_q,
    where g_q = g*alpha is the gradient reaching q through the inverse scale."""
    gq = (np.asarray(g, F32) * alpha).astype(F32)
    gx = (gq / alpha).astype(F32)
    gc = np.zeros(nb_clusters, dtype=np.float64)
    np.add.at(gc, np.asarray(idx).reshape(-1), gq.reshape(-1).astype(np.float64))
    return gx, gc.astype(F32)


# ----------------------------------------------------------------------------
# schedules  (uniform_quantization/learner.py:50-70, utils/lrn_rate_utils.py:23-46)
# ----------------------------------------------------------------------------
def piecewise_constant(step, bnds, vals):
    """tf.train.piecewise_constant: vals[0] for step <= bnds[0], vals[i] for
    bnds[i-1] < step <= bnds[i], vals[-1] beyond."""
    for b, v in zip(bnds, vals):
        if step <= b:
            return v
    return vals[-1]


def uq_bnds_decay_rates(model_name, dataset_name, nb_smpls_train, batch_size, world,
                        lrn_rate_init, batch_size_norm, quant_epochs=60, enbl_warm_start=False,
                        enbl_multi_gpu=False):
    """setup_bnds_decay_rates (uniform_quantization/learner.py:50-70).  For
    lenet@cifar_10 the reference leaves bnds unbound (UnboundLocalError, SURVEY
    A.6-1); the ResNet/CIFAR schedule is adopted there and flagged."""
    bs = batch_size * world if enbl_multi_gpu else batch_size
    nb_batches_per_epoch = int(nb_smpls_train / bs)
    init_lr = lrn_rate_init * batch_size * world / batch_size_norm if enbl_multi_gpu else lrn_rate_init
    if dataset_name == 'cifar_10':
        bnds = [nb_batches_per_epoch * 15, nb_batches_per_epoch * 40]
        decay_rates = [1e-3, 1e-4, 1e-5]
    elif dataset_name == 'ilsvrc_12':
        if model_name.startswith('resnet'):
            bnds = [nb_batches_per_epoch * 5, nb_batches_per_epoch * 20]
        else:
            bnds = [nb_batches_per_epoch * 5, nb_batches_per_epoch * 30]
        decay_rates = [1e-4, 1e-5, 1e-6]
    else:
        raise ValueError('unrecognized dataset')
    finetune_steps = nb_batches_per_epoch * quant_epochs
    init_lr = init_lr if enbl_warm_start else lrn_rate_init
    return init_lr, bnds, decay_rates, finetune_steps


def lrn_rate_piecewise_constant(step, batch_size, idxs_epoch, decay_rates, lrn_rate_init,
                                batch_size_norm, nb_smpls_train, nb_epochs_rat=1.0):
    """setup_lrn_rate_piecewise_constant (utils/lrn_rate_utils.py:23-46)."""
    idxs_epoch = [e * nb_epochs_rat for e in idxs_epoch]
    init = lrn_rate_init * batch_size / batch_size_norm
    nbpe = float(nb_smpls_train) / batch_size
    bnds = [int(nbpe * e) for e in idxs_epoch]
    vals = [init * d for d in decay_rates]
    return piecewise_constant(step, bnds, vals)
