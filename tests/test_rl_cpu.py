"""The RL hyper-parameter search around the training step: roll-out bookkeeping pinned
against the reference's own code (tests/golden/ref_executed_v1.json, sections uq_rl_helper / ddpg_* / uq_bit_optimizer),
the torch DDPG agent on the reference's move-to-target toy, and the BitOptimizer loop driven with a stand-in learner."""
import json
import os
import random

import numpy as np
import pytest

from pocketflow_b200.flags import FLAGS

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'ref_executed_v1.json')))


@pytest.fixture(autouse=True)
def _fresh_flags():
    import pocketflow_b200.learners.uniform_quantization.bit_optimizer  # noqa: F401  (defines the uql_* flags)
    import pocketflow_b200.rl_agents.ddpg.agent  # noqa: F401
    FLAGS.reset()
    yield
    FLAGS.reset()


def test_rl_helper_matches_the_executed_reference():
    from pocketflow_b200.learners.uniform_quantization.rl_helper import RLHelper
    assert len(GOLD['uq_rl_helper']) == 60
    for g in GOLD['uq_rl_helper']:
        FLAGS.uql_w_bit_min, FLAGS.uql_w_bit_max = g['w_bit_min'], g['w_bit_max']
        shapes = [tuple(s) for s in g['shapes']]
        nums = [int(np.prod(s)) for s in shapes]
        h = RLHelper(sum(nums) * g['equivalent_bits'], nums, shapes, random_layers=g['random_layers'])
        assert h.s_dims == g['s_dims']
        for i, want in enumerate(g['states']):
            assert h.calc_state(i).shape == (1, h.s_dims)
            np.testing.assert_array_equal(h.calc_state(i)[0], np.asarray(want))
        random.seed(g['py_seed'])
        for ro in g['rollouts']:
            h.reset()
            assert list(h.layer_idxs) == ro['order']
            bits = []
            for raw, idx in zip(ro['raw'], ro['order']):
                a = h.calc_w(np.array([[raw]]), idx)
                assert a.shape == (1, 1)
                bits.append(float(a[0][0]))
            assert bits == ro['bits'] and float(h.w_bits_used) == ro['used']
            assert h.w_bits_used <= h.total_bits and all(g['w_bit_min'] <= b <= g['w_bit_max'] for b in bits)
        np.testing.assert_array_equal(h.calc_reward(0.625), np.asarray(g['reward']))


def test_replay_buffer_and_noise_match_the_executed_reference():
    from pocketflow_b200.rl_agents.ddpg.replay_buffer import ReplayBuffer
    from pocketflow_b200.rl_agents.ddpg import noise
    for g in GOLD['ddpg_replay_buffer']:
        rng = np.random.RandomState(g['seed'])
        buf = ReplayBuffer(3, 2, g['buf_size'])
        for n, t in zip(g['chunks'], g['trace']):
            buf.append(rng.randn(n, 3), rng.randn(n, 2), rng.randn(n, 1), (rng.rand(n, 1) < 0.3).astype(float), rng.randn(n, 3))
            assert (buf.idx_smpl, buf.nb_smpls, buf.is_ready()) == (t['idx_smpl'], t['nb_smpls'], t['ready'])
            np.testing.assert_array_equal(buf.buffers['states'], np.asarray(t['states'], np.float32))
            np.testing.assert_array_equal(buf.buffers['rewards'], np.asarray(t['rewards'], np.float32))
        mb = buf.sample(16)
        assert mb['states'].shape == (16, 3) and mb['actions'].shape == (16, 2) and mb['terminals'].shape == (16, 1)
    for g in GOLD['ddpg_noise']:
        FLAGS.ddpg_noise_std_init, FLAGS.ddpg_noise_std_finl = g['std_init'], g['std_finl']
        td = noise.TimeDecayNoiseSpec(g['nb_rlouts'])
        seq = []
        for _ in range(5):
            td.adapt()
            seq.append(td.stdev_curr)
        assert seq == g['tdecy']
        ad = noise.AdaptiveNoiseSpec()
        seq = []
        for dst in (0.5, 0.02, 0.001, 0.0, 0.3):
            ad.adapt(dst)
            seq.append(ad.stdev_curr)
        assert seq == g['adapt']
        td.reset()
        assert td.stdev_curr == g['std_init']


def _move_to_target(agent, policy, train, nb_dims, rlout_len, rng):
    """rl_agents/unit_tests/move_to_target.py: reward = progress towards the origin minus the distance moved (<= 0)."""
    x = rng.uniform(-10, 10, (1, nb_dims))
    rewards = []
    for i in range(rlout_len):
        a = policy(x)
        xn = x + a
        r = np.linalg.norm(x) - np.linalg.norm(xn) - np.linalg.norm(x - xn)
        if train:
            agent.record(x, a, r * np.ones((1, 1)), np.ones((1, 1)) * (i == rlout_len - 1), xn)
            agent.train()
        x = xn
        rewards.append(r)
    return float(np.mean(rewards))


@pytest.mark.parametrize('noise_type,noise_prtl', [('param', 'tdecy'), ('action', 'tdecy'), ('param', 'adapt')])
def test_ddpg_agent_learns_the_move_to_target_problem(noise_type, noise_prtl):
    import torch
    from pocketflow_b200.rl_agents.ddpg.agent import Agent
    FLAGS.ddpg_noise_type, FLAGS.ddpg_noise_prtl = noise_type, noise_prtl
    torch.manual_seed(0)
    nb_dims, nb_rlouts, rlout_len = 2, 30, 50
    rng = np.random.RandomState(0)
    agent = Agent(nb_dims, nb_dims, nb_rlouts, int(rlout_len * nb_rlouts * 0.25), -1.0, 1.0, seed=0)
    before = np.mean([_move_to_target(agent, agent.actions_clean, False, nb_dims, rlout_len, rng) for _ in range(20)])
    assert agent.train() == (0.0, 0.0, agent.noise_spec.stdev_curr)            # nothing to learn from yet
    std0 = agent.noise_spec.stdev_curr
    for _ in range(nb_rlouts):
        agent.init_rlout()
        noisy = agent.actions_noisy(np.zeros((1, nb_dims)))
        assert noisy.shape == (1, nb_dims) and (-1.0 <= noisy).all() and (noisy <= 1.0).all()
        agent.finalize_rlout([_move_to_target(agent, agent.actions_noisy, True, nb_dims, rlout_len, rng)])
    after = np.mean([_move_to_target(agent, agent.actions_clean, False, nb_dims, rlout_len, rng) for _ in range(20)])
    assert agent.memory.is_ready() and not agent.in_explore and agent.reward_ema is not None
    assert agent.noise_spec.stdev_curr != std0
    assert after > 0.35 * before and after > -0.6, (before, after)             # optimum 0; untrained about -1.3
    # targets track the mains, the parameter-noise copy differs from the clean actor only where it may
    for p, q in zip(agent.actor.parameters(), agent.actor_tr.parameters()):
        assert not q.requires_grad and (p - q).abs().max() < 1.0
    agent.init()
    assert agent.memory.nb_smpls == 0 and agent.in_explore and agent.noise_spec.stdev_curr == FLAGS.ddpg_noise_std_init


def test_parameter_noise_leaves_layer_norm_untouched():
    import torch
    from pocketflow_b200.rl_agents.ddpg.agent import Agent
    agent = Agent(5, 1, 10, 8, 0.0, 6.0, seed=1)
    agent.init_rlout()
    perturbable = {id(p) for p in agent.actor.perturbable_params}
    n_ln = 0
    for (name, p), q in zip(agent.actor.named_parameters(), agent.actor_np.parameters()):
        if id(p) in perturbable:
            assert not torch.equal(p, q), name
        else:
            n_ln += 1
            assert torch.equal(p, q), name
    assert n_ln == 2 * FLAGS.ddpg_actor_depth                                   # gain + offset of every layer norm
    a = agent.actions_clean(np.random.RandomState(0).randn(7, 5))
    assert a.shape == (7, 1) and (0.0 <= a).all() and (a <= 6.0).all()


class _Tuner(object):
    """Stands in for the learner: the 'accuracy' after fine-tuning is a known function of the bit allocation (more
    bits on the layers with large `gain` are worth more), so the search has something real to find."""
    device = 'cpu'

    def __init__(self, gains):
        self.gains = np.asarray(gains, float)
        self.calls = []
        self.bits = None

    def rl_restore(self):
        self.calls.append('restore')

    def rl_set_bits(self, w_bits, a_bits):
        self.calls.append('set_bits')
        self.bits, self.a_bits = list(w_bits), list(a_bits)

    def rl_finetune(self, nb_steps, disp_steps):
        self.calls.append(('finetune', nb_steps, disp_steps))

    def rl_evaluate(self):
        self.calls.append('evaluate')
        acc = float(np.sum(self.gains * (1.0 - 2.0 ** (-np.asarray(self.bits, float) / 2.0))) / np.sum(self.gains))
        return 1.0 - acc, acc, min(1.0, acc + 0.1)


def test_bit_optimizer_without_the_agent_returns_the_flag_values():
    from pocketflow_b200.learners.uniform_quantization.bit_optimizer import BitOptimizer
    FLAGS.uql_weight_bits, FLAGS.uql_activation_bits = 5, 7
    bo = BitOptimizer('cifar_10', [], dict(nb_matmuls=3, nb_activations=2, num_weights=[1, 2, 3]))
    assert bo.run() == ([5, 5, 5], [7, 7])


def test_bit_optimizer_search_respects_the_budget_and_keeps_the_best_rollout(capsys):
    from types import SimpleNamespace
    from pocketflow_b200.learners.uniform_quantization.bit_optimizer import BitOptimizer
    shapes = [(3, 3, 8, 16), (3, 3, 16, 16), (1, 1, 16, 64), (64, 10)]
    nums = [int(np.prod(s)) for s in shapes]
    FLAGS.uql_enbl_rl_agent, FLAGS.uql_nb_rlouts, FLAGS.uql_equivalent_bits = True, 24, 4
    FLAGS.uql_tune_global_steps, FLAGS.uql_tune_disp_steps = 40, 10
    random.seed(0)
    tuner = _Tuner([1.0, 0.2, 3.0, 0.5])
    stats = dict(nb_matmuls=4, nb_activations=3, num_weights=nums)
    bo = BitOptimizer('cifar_10', [SimpleNamespace(shape=s) for s in shapes], stats, tuner=tuner, seed=0)
    recorded = []
    real_record = bo.agent.record
    bo.agent.record = lambda *a: (recorded.append([np.asarray(x).copy() for x in a]), real_record(*a))[1]
    w_bits, a_bits = bo.run()
    out = capsys.readouterr().out
    assert a_bits == [32, 32, 32] and len(w_bits) == 4 and all(isinstance(b, int) and 2 <= b <= 8 for b in w_bits)
    assert sum(b * n for b, n in zip(w_bits, nums)) <= 4 * sum(nums)                       # the budget
    # every roll-out: restore -> set bits -> fine-tune (global steps, display interval) -> evaluate
    per = [tuner.calls[i:i + 4] for i in range(0, len(tuner.calls), 4)]
    assert len(per) == 24 and all(c == ['restore', 'set_bits', ('finetune', 40, 10), 'evaluate'] for c in per)
    assert len(bo.reward_list) == 24 and 'Finished RL training' in out and out.count('starting') == 24
    # the returned allocation is the best roll-out's, and its reward is what the stand-in assigns to it
    tuner.bits = w_bits
    assert abs(tuner.rl_evaluate()[1] - max(bo.reward_list)) < 1e-12
    # transitions: 4 per roll-out in LAYER order, the roll-out's reward on each, the last one terminal with a zero
    # next state, next states chained
    assert len(recorded) == 24 * 4
    for r in range(24):
        tr = recorded[4 * r:4 * r + 4]
        for n, (s, a, rew, term, s_next) in enumerate(tr):
            assert s.shape == (1, bo.s_dims) and s[0, n] == 1.0 and a.shape == (1, 1)
            assert float(rew[0][0]) == bo.reward_list[r] and float(term[0][0]) == (1.0 if n == 3 else 0.0)
            if n < 3:
                np.testing.assert_array_equal(s_next, tr[n + 1][0])
            else:
                assert not s_next.any()
    assert bo.agent.memory.is_ready() and not bo.agent.in_explore                          # 96 transitions fill 4 * 6 slots
    # budget helper: the same numbers the reference's __check_bits gives
    g = GOLD['uq_bit_optimizer']
    small = BitOptimizer('cifar_10', [SimpleNamespace(shape=(2, 5)), SimpleNamespace(shape=(4, 5)), SimpleNamespace(shape=(1, 5))],
                         dict(nb_matmuls=3, nb_activations=0, num_weights=[10, 20, 5]), tuner=tuner, seed=0)
    assert small.check_bits([4, 3, 8]) == g['check_ok']
    with pytest.raises(ValueError, match=g['check_over']):
        small.check_bits([8, 8, 8])
    # the reference's transition layout for 3 layers, from its own __record_rollout_transitions
    want = g['transitions']
    assert [np.asarray(t[3]).tolist() for t in want] == [[[0.0]], [[0.0]], [[1.0]]] and not np.asarray(want[2][4]).any()
    assert np.asarray(want[0][4]).tolist() == np.asarray(want[1][0]).tolist()


def test_bit_optimizer_refuses_what_it_cannot_do():
    from pocketflow_b200.learners.uniform_quantization.bit_optimizer import BitOptimizer
    stats = dict(nb_matmuls=1, nb_activations=0, num_weights=[10])
    FLAGS.uql_enbl_rl_agent = True
    with pytest.raises(ValueError, match='needs the learner'):
        BitOptimizer('cifar_10', [], stats)
    FLAGS.uql_enbl_rl_layerwise_tune = True
    with pytest.raises(NotImplementedError):
        BitOptimizer('cifar_10', [], stats, tuner=_Tuner([1.0]))


def test_nuq_bit_optimizer_is_the_same_search_on_the_nuql_flags():
    """learners/nonuniform_quantization/bit_optimizer.py + rl_helper.py are the uniform learner's files with the flags
    renamed: same loop, its own budget / range / roll-out count."""
    from types import SimpleNamespace
    import pocketflow_b200.learners.nonuniform_quantization.learner  # noqa: F401  (declares nuql_weight_bits, ...)
    from pocketflow_b200.learners.nonuniform_quantization.bit_optimizer import BitOptimizer
    FLAGS.nuql_weight_bits, FLAGS.nuql_activation_bits = 3, 32
    stats = dict(nb_matmuls=2, nb_activations=1, num_weights=[10, 20])
    assert BitOptimizer('cifar_10', [], stats).run() == ([3, 3], [32])
    shapes = [(3, 3, 8, 16), (3, 3, 16, 16), (64, 10)]
    nums = [int(np.prod(s)) for s in shapes]
    FLAGS.nuql_enbl_rl_agent, FLAGS.nuql_nb_rlouts, FLAGS.nuql_equivalent_bits = True, 6, 3
    FLAGS.nuql_w_bit_min, FLAGS.nuql_w_bit_max = 2, 5
    FLAGS.nuql_tune_global_steps, FLAGS.nuql_tune_disp_steps = 8, 4
    FLAGS.uql_nb_rlouts, FLAGS.uql_w_bit_max = 1, 8                       # must not be read
    random.seed(1)
    tuner = _Tuner([1.0, 0.3, 2.0])
    bo = BitOptimizer('cifar_10', [SimpleNamespace(shape=s) for s in shapes],
                      dict(nb_matmuls=3, nb_activations=2, num_weights=nums), tuner=tuner, seed=0)
    w_bits, a_bits = bo.run()
    assert a_bits == [32, 32] and all(2 <= b <= 5 for b in w_bits)
    assert sum(b * n for b, n in zip(w_bits, nums)) <= 3 * sum(nums)
    per = [tuner.calls[i:i + 4] for i in range(0, len(tuner.calls), 4)]
    assert len(per) == 6 and all(c == ['restore', 'set_bits', ('finetune', 8, 4), 'evaluate'] for c in per)
    FLAGS.nuql_enbl_rl_layerwise_tune = True
    with pytest.raises(NotImplementedError, match='nuql_enbl_rl_layerwise_tune'):
        BitOptimizer('cifar_10', [], stats, tuner=tuner)


def test_ws_rl_helper_matches_the_executed_reference():
    import pocketflow_b200.learners.weight_sparsification.learner  # noqa: F401  (defines the ws_* flags)
    from pocketflow_b200.learners.weight_sparsification.rl_helper import RLHelper
    assert len(GOLD['ws_rl_helper']) == 36
    n_err = 0
    for g in GOLD['ws_rl_helper']:
        FLAGS.ws_prune_ratio, FLAGS.ws_reward_type = g['ws_prune_ratio'], g['reward_type']
        h = RLHelper([tuple(s) for s in g['shapes']], g['skip_head_n_tail'])
        assert h.s_dims == g['s_dims']
        for ro in g['rollouts']:
            states, ratios, error = [], [], None
            try:
                for idx, a in enumerate(ro['actions']):
                    states.append(h.calc_state(idx)[0].tolist())
                    ratios.append(float(h.cvt_action_to_prune_ratio(idx, a)))
            except AssertionError as e:
                error = str(e)
                n_err += 1
            assert error == ro['error']
            assert states == ro['states'] and ratios == ro['ratios']
            assert float(h.calc_overall_prune_ratio()) == ro['overall'] and float(h.calc_reward(0.8)) == ro['reward']
            if error is None and g['reward_type'] == 'single-obj':
                assert ro['overall'] >= g['ws_prune_ratio'] - 1e-9                 # the target is a hard constraint
    assert n_err == 27


class _PruneTuner(object):
    """Stand-in for the device half of the 'optimal' protocol: accuracy falls with the pruning of 'sensitive' layers and
    recovers a little with retraining."""
    device = 'cpu'

    def __init__(self, sens):
        self.sens = np.asarray(sens, float)
        self.calls, self.retrained = [], False

    def pr_prune(self, prune_ratios):
        self.calls.append('prune')
        self.ratios, self.retrained = np.asarray(prune_ratios, float), False

    def pr_retrain(self, nb_iters_rg, nb_iters_ft):
        self.calls.append(('retrain', nb_iters_rg, nb_iters_ft))
        self.retrained = True

    def pr_evaluate(self):
        self.calls.append('evaluate')
        acc = 1.0 - float(np.sum(self.sens * self.ratios ** 2) / np.sum(self.sens)) * (0.6 if self.retrained else 1.0)
        return 1.0 - acc, {'accuracy': acc, 'acc_top5': min(1.0, acc + 0.05)}


def test_pr_optimizer_protocols_and_the_optimal_search(capsys):
    from types import SimpleNamespace
    import pocketflow_b200.learners.weight_sparsification.learner  # noqa: F401
    from pocketflow_b200.learners.weight_sparsification.pr_optimizer import PROptimizer
    shapes = [(3, 3, 3, 16), (3, 3, 16, 32), (3, 3, 32, 32), (1, 1, 32, 64), (64, 10)]
    mvars = [SimpleNamespace(name='model/v%d:0' % i, shape=s, numel=int(np.prod(s))) for i, s in enumerate(shapes)]
    FLAGS.ws_prune_ratio, FLAGS.ws_prune_ratio_prtl = 0.6, 'uniform'
    assert PROptimizer(mvars).run() == [(v.name, 0.6) for v in mvars]
    FLAGS.ws_prune_ratio_prtl = 'heurist'
    heur = PROptimizer(mvars).run()
    n = np.array([v.numel for v in mvars], float)
    assert abs(sum(r * k for (_, r), k in zip(heur, n)) / n.sum() - 0.6) < 1e-12          # overall ratio = the target
    FLAGS.ws_prune_ratio_prtl = 'optimal'
    with pytest.raises(ValueError, match='needs the learner'):
        PROptimizer(mvars, 'cifar_10').run()
    FLAGS.ws_prune_ratio_prtl = 'random'
    with pytest.raises(ValueError):
        PROptimizer(mvars)
    FLAGS.ws_prune_ratio_prtl = 'optimal'
    FLAGS.ws_nb_rlouts, FLAGS.ws_nb_rlouts_min, FLAGS.ws_nb_iters_rg, FLAGS.ws_nb_iters_ft = 20, 4, 20, 400
    tuner = _PruneTuner([5.0, 1.0, 0.2, 0.2, 3.0])
    opt = PROptimizer(mvars, 'ilsvrc_12', tuner=tuner, seed=0)
    out = opt.run()
    text = capsys.readouterr().out
    assert [name for name, _ in out] == [v.name for v in mvars]
    ratios = np.array([r for _, r in out])
    assert (ratios >= 0).all() and (ratios <= 1.0 - 0.4 / 3.0 + 1e-12).all()
    assert float(np.sum(ratios * n) / n.sum()) >= 0.6 - 1e-9                              # single-obj: target reached
    per = [tuner.calls[i:i + 4] for i in range(0, len(tuner.calls), 4)]
    assert len(per) == 20 and all(c == ['prune', 'evaluate', ('retrain', 20, 400), 'evaluate'] for c in per)
    assert len(opt.rewards) == 20 and text.count('starting') == 20 and 'best reward updated' in text
    tuner.pr_prune(ratios)
    tuner.pr_retrain(0, 0)
    assert abs(tuner.pr_evaluate()[1]['accuracy'] - max(opt.rewards)) < 1e-12             # the best roll-out is returned
    assert opt.agent.memory.is_ready() and opt.agent.memory.buf_size == 5 * 4
    # CIFAR-10: head and tail layers are never pruned
    FLAGS.ws_nb_rlouts = 3
    out = PROptimizer(mvars, 'cifar_10', tuner=_PruneTuner([1.0] * 5), seed=0).run()
    assert out[0][1] == 0.0 and out[-1][1] == 0.0 and all(r > 0 for _, r in out[1:-1])


def test_uq_learner_rl_hooks_compose_the_executor_calls(tmp_path):
    """The learner side of the bit search (rl_restore / rl_set_bits / rl_finetune / rl_evaluate), run against a
    recording stand-in for the executor: the learner itself only constructs on a GPU."""
    from types import SimpleNamespace
    from pocketflow_b200.learners.uniform_quantization.learner import UniformQuantLearner as L
    from pocketflow_b200.learners.abstract_learner import save_checkpoint
    log = []
    state = {'model/w:0': np.arange(4, dtype=np.float32)}
    store = SimpleNamespace(state_dict=lambda: {k: v.copy() for k, v in state.items()},
                            load_state_dict=lambda d, strict=True: log.append(('load', sorted(d), strict, d['model/w:0'].tolist())),
                            P='P', O='O')
    ex = SimpleNamespace(store=store, step_count=37,
                         reset_optimizer_state=lambda: log.append('reset_opt'),
                         set_quant_bits=lambda w, a: log.append(('bits', list(w), list(a))),
                         forward_eval_loss=lambda: log.append('eval_fwd'),
                         fetch_losses=lambda: dict(loss=2.0, acc_top1=0.25, acc_top5=0.5))
    me = SimpleNamespace(sess_train=ex, _rl_initial_state=None, iterator_train='it',
                         train_step=lambda: log.append('step'),
                         feed=lambda e, it: log.append(('feed', it)), eval_iterator=lambda: 'eval_it',
                         _UniformQuantLearner__monitor_progress=lambda r, t, i: (log.append(('monitor', i)), t)[1],
                         _UniformQuantLearner__eval_batch_size=lambda: FLAGS.batch_size_eval)
    FLAGS.save_path = str(tmp_path / 'models' / 'model.ckpt')
    # no checkpoint on disk: the state the learner was built with is the restore point, taken once
    L.rl_restore(me)
    state['model/w:0'] += 100.0                                  # later training must not leak into the restore point
    L.rl_restore(me)
    assert log == [('load', ['model/w:0'], False, [0.0, 1.0, 2.0, 3.0]), 'reset_opt'] * 2
    # a pre-trained checkpoint under --save_path wins
    del log[:]
    me._rl_initial_state = None
    save_checkpoint(FLAGS.save_path, {'model/w:0': np.full(4, 7.0, np.float32)}, 5)
    L.rl_restore(me)
    assert log[0] == ('load', ['model/w:0'], False, [7.0] * 4)
    del log[:]
    L.rl_set_bits(me, [2, 8], [32])
    L.rl_finetune(me, 5, 2)
    assert log == [('bits', [2, 8], [32]), 'step', 'step', ('monitor', 1), 'step', 'step', ('monitor', 3), 'step']
    assert ex.step_count == 0                                    # reset_ft_step
    del log[:]
    FLAGS.nb_smpls_eval, FLAGS.batch_size_eval = 300, 100
    assert L.rl_evaluate(me) == (2.0, 0.25, 0.5)
    assert log == [('feed', 'eval_it'), 'eval_fwd'] * 3


def test_ws_learner_search_hooks_compose_the_executor_calls(tmp_path):
    """pr_reset / pr_prune / pr_retrain / pr_evaluate of the WeightSparseLearner against a recording stand-in."""
    from types import SimpleNamespace
    from pocketflow_b200.learners.weight_sparsification.learner import WeightSparseLearner as L
    log = []
    state = {'model/w:0': np.arange(3, dtype=np.float32)}
    store = SimpleNamespace(state_dict=lambda: {k: v.copy() for k, v in state.items()},
                            load_state_dict=lambda d, strict=True: log.append(('load', d['model/w:0'].tolist(), strict)),
                            P='P', O='O')
    ex = SimpleNamespace(store=store, MASK=SimpleNamespace(fill_=lambda v: log.append(('mask_fill', v))),
                         mask_builder=SimpleNamespace(build=lambda r: log.append(('build', r))),
                         reset_optimizer_state=lambda: log.append('reset_opt'),
                         run_step=lambda lr, ar: log.append(('step', lr, ar)),
                         forward_eval_loss=lambda: log.append('eval_fwd'),
                         fetch_losses=lambda: dict(loss=1.5, acc_top1=0.5, acc_top5=0.75))
    me = SimpleNamespace(sess_train=ex, _pr_full_state=None, iterator_train='it', dataset_name='cifar_10',
                         feed=lambda e, it: log.append(('feed', it)), eval_iterator=lambda: 'eval_it',
                         grad_allreduce=lambda: None)
    me.pr_reset = lambda: L.pr_reset(me)
    FLAGS.save_path = str(tmp_path / 'none' / 'model.ckpt')
    L.pr_prune(me, np.array([0.0, 0.5, 0.25]))
    state['model/w:0'] += 9.0
    L.pr_prune(me, [0.1, 0.1, 0.1])
    restore = [('load', [0.0, 1.0, 2.0], False), ('mask_fill', 1.0), 'reset_opt']
    assert log == restore + [('build', [0.0, 0.5, 0.25])] + restore + [('build', [0.1, 0.1, 0.1])]
    del log[:]
    FLAGS.ws_lrn_rate_ft = 3e-4
    me.pr_regress_layers = lambda n: log.append(('regress', n))
    L.pr_retrain(me, 20, 3)
    assert log == [('regress', 20)] + [('feed', 'it'), ('step', 3e-4, None)] * 3   # layer-wise regression, then fine-tuning
    del log[:]
    L.pr_retrain(me, 0, 1)
    assert log == [('feed', 'it'), ('step', 3e-4, None)]
    del log[:]
    FLAGS.ws_nb_iters_feval = 2
    assert L.pr_evaluate(me) == (1.5, {'accuracy': 0.5})
    assert log == [('feed', 'eval_it'), 'eval_fwd'] * 2
    me.dataset_name = 'ilsvrc_12'
    FLAGS.ws_nb_iters_feval, FLAGS.nb_smpls_eval, FLAGS.batch_size_eval = 0, 300, 100
    del log[:]
    assert L.pr_evaluate(me) == (1.5, {'acc_top1': 0.5, 'acc_top5': 0.75}) and len(log) == 6
