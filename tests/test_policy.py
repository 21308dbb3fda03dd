"""Policy replay path (``SURVEY.md 8f.1``): GRU sequence kernels, mixture-of-Gaussians head, optimiser step, PPO iteration.

The reference's tests hold no vectors for this path (parity unpinned by the reference).  The CPU tests pin this repo's fp32
restatements (``*_reference``) against independent library implementations of the same published definitions
(``torch.nn.GRUCell``, ``torch.distributions``, ``torch.optim.Adam``); the GPU tests compare the CUDA kernels with those
restatements.  Tolerances: the tensor-core path rounds GEMM operands to bf16 (fp32 accumulation, fp32 state), so sequence
outputs are compared at 3e-2 of the fp32 result's scale; the fp32 elementwise kernels (mixture head, optimiser) at 1e-5.
"""

import math

import numpy as np
import pytest
import torch

from ksim_b200 import optim, policy, replay

SPEC = policy.MixtureSpec(num_joints=5, num_mixtures=4, min_std=0.01, max_std=1.0, var_scale=0.5, mean_bias=(0.1, -0.2, 0.0, 0.3, 0.05),
                          clip_lo=(-0.5, -1.0, -0.3, -2.0, -0.1), clip_hi=(0.5, 1.0, 0.3, 2.0, 0.1))


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-12))


# ------------------------------------------------------------------------------------------------------------- CPU
def test_gru_cell_reference_matches_torch_grucell():
    """eqx.nn.GRUCell = torch.nn.GRUCell with b_ih = bias and b_hh = [0, 0, bias_n] (same gate order r, z, n)."""
    torch.manual_seed(0)
    hd, b = 16, 7
    cell = torch.nn.GRUCell(hd, hd).double()
    w_ih, w_hh, bias = cell.weight_ih.detach(), cell.weight_hh.detach(), cell.bias_ih.detach().clone()
    bias_n = cell.bias_hh.detach()[2 * hd :].clone()
    with torch.no_grad():
        cell.bias_hh[: 2 * hd] = 0.0
    x, h = torch.randn(b, hd, dtype=torch.float64), torch.randn(b, hd, dtype=torch.float64)
    torch.testing.assert_close(replay.gru_cell_reference(x, h, w_ih, w_hh, bias, bias_n), cell(x, h), rtol=1e-12, atol=1e-12)


def test_gru_stack_reference_restarts_the_carry_after_a_terminal_step():
    torch.manual_seed(1)
    stack = replay.GRUStack(2)
    x = torch.randn(3, 2, replay.HIDDEN)
    keep = torch.tensor([[1.0, 0.0], [1.0, 1.0], [1.0, 1.0]])
    y, carry = replay.gru_stack_reference(stack, x, None, keep)
    # env 1 ended at step 0: its steps 1.. must equal a fresh sequence started at step 1
    y2, carry2 = replay.gru_stack_reference(stack, x[1:, 1:], None, None)
    torch.testing.assert_close(y[1:, 1], y2[:, 0])
    torch.testing.assert_close(carry[:, 1], carry2[:, 0])


def test_mixture_reference_matches_torch_distributions():
    torch.manual_seed(2)
    pred = torch.randn(11, 3 * SPEC.num_joints * SPEC.num_mixtures, dtype=torch.float64)
    x = torch.randn(11, SPEC.num_joints, dtype=torch.float64) * 0.4
    mean, std, logw = policy.mixture_params_reference(pred, SPEC)
    assert float(std.max()) <= SPEC.max_std and float(std.min()) >= SPEC.min_std * SPEC.var_scale
    lo, hi = torch.tensor(SPEC.clip_lo, dtype=torch.float64)[:, None], torch.tensor(SPEC.clip_hi, dtype=torch.float64)[:, None]
    assert bool((mean >= lo).all()) and bool((mean <= hi).all())
    dist = torch.distributions.MixtureSameFamily(torch.distributions.Categorical(logits=logw), torch.distributions.Normal(mean, std))
    torch.testing.assert_close(policy.mixture_log_prob_reference(pred, x, SPEC), dist.log_prob(x), rtol=1e-10, atol=1e-10)


def test_optimizer_reference_matches_torch_adam():
    """The optax chain against torch.optim.Adam (L2 weight decay = add_decayed_weights before scale_by_adam) driven with the
    schedule's learning rate and a hand-clipped gradient."""
    rs = np.random.RandomState(3)
    cfg = optim.OptimizerConfig(learning_rate=1e-2, warmup_steps=4, adam_weight_decay=1e-3, grad_clip=0.5)
    p = rs.randn(37)
    tp = torch.nn.Parameter(torch.tensor(p))
    opt = torch.optim.Adam([tp], lr=1.0, betas=(cfg.b1, cfg.b2), eps=cfg.eps, weight_decay=cfg.adam_weight_decay)
    mu, nu = np.zeros_like(p), np.zeros_like(p)
    for count in range(7):
        g = rs.randn(37) * (0.02 if count % 2 else 1.0)   # alternately below and above the clip norm
        p, mu, nu, norm, lr = optim.optimizer_step_reference(p, g, mu, nu, count, cfg)
        init = cfg.learning_rate * cfg.init_scale
        assert lr == pytest.approx(init + (cfg.learning_rate - init) * min(count / cfg.warmup_steps, 1.0))
        gc = g if norm < cfg.grad_clip else g / norm * cfg.grad_clip
        for grp in opt.param_groups:
            grp["lr"] = lr
        tp.grad = torch.tensor(gc)
        opt.step()
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-9, atol=1e-12)
    # zero_nans
    g = rs.randn(37)
    g[5] = np.nan
    p2, *_ = optim.optimizer_step_reference(p, g, mu, nu, 7, cfg)
    g[5] = 0.0
    p3, *_ = optim.optimizer_step_reference(p, g, mu, nu, 7, cfg)
    np.testing.assert_array_equal(p2, p3)


def test_flat_params_views_and_grad_accumulation():
    torch.manual_seed(4)
    m = policy.MLP(6, 3, 8, 2, "elu")
    before = [p.detach().clone() for p in m.parameters()]
    fp = optim.FlatParams(m)
    for p, b in zip(m.parameters(), before):
        torch.testing.assert_close(p.detach(), b)
        assert p.data_ptr() >= fp.flat.data_ptr() and p.data_ptr() % 16 == 0
        torch.testing.assert_close(p.shadow.float(), b.to(torch.bfloat16).float())
    x = torch.randn(5, 6)
    m.reference(x).sum().backward()
    assert float(fp.grad.abs().sum()) > 0      # autograd accumulated into the flat buffer, not into fresh tensors
    g1 = fp.grad.clone()
    m.reference(x).sum().backward()
    torch.testing.assert_close(fp.grad, 2 * g1)
    fp.zero_grad()
    assert float(fp.grad.abs().sum()) == 0
    assert all(p.grad is not None for p in m.parameters())
    assert fp.allreduce_grads() == 1.0          # single process: no collective


def test_get_ppo_variables_reference_shapes_and_done_semantics():
    torch.manual_seed(5)
    m = policy.Model(num_actor_inputs=9, num_critic_inputs=13, num_joints=SPEC.num_joints, num_mixtures=SPEC.num_mixtures,
                     mean_bias=SPEC.mean_bias, clip_lo=SPEC.clip_lo, clip_hi=SPEC.clip_hi)
    t, b = 4, 3
    oa, oc, act = torch.randn(t, b, 9), torch.randn(t, b, 13), torch.randn(t, b, SPEC.num_joints) * 0.3
    done = torch.zeros(t, b, dtype=torch.bool)
    done[1, 2] = True
    lp, v, carry = policy.get_ppo_variables_reference(m, oa, oc, act, done, {})
    assert lp.shape == (t, b, SPEC.num_joints) and v.shape == (t, b) and carry["actor"].shape == (2, b, replay.HIDDEN)
    lp2, v2, _ = policy.get_ppo_variables_reference(m, oa[2:, 2:], oc[2:, 2:], act[2:, 2:], done[2:, 2:], {})
    torch.testing.assert_close(lp[2:, 2], lp2[:, 0])
    torch.testing.assert_close(v[2:, 2], v2[:, 0])


def test_kernel_paths_refuse_cpu_tensors():
    from ksim_b200 import binding

    with pytest.raises(binding.B200SimError):
        policy.MixtureOfGaussians(torch.zeros(2, 3 * SPEC.num_joints * SPEC.num_mixtures), SPEC).log_prob(torch.zeros(2, SPEC.num_joints))
    with pytest.raises(binding.B200SimError):
        replay.GRUStack(1)(torch.zeros(1, 2, replay.HIDDEN), None, None)
    with pytest.raises(binding.B200SimError):
        optim.FlatParams(policy.MLP(4, 2, 4, 1, "elu")).step()


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("t,b,depth,stacks", [(3, 128, 1, 1), (5, 200, 2, 2), (24, 512, 2, 2)])
def test_gru_kernels_match_fp32_restatement(t, b, depth, stacks):
    torch.manual_seed(10 + t)
    dev = torch.device("cuda:0")
    hd = replay.HIDDEN
    nets = [replay.GRUStack(depth).to(dev) for _ in range(stacks)]
    xs = [torch.randn(t, b, hd, device=dev).mul_(0.7).requires_grad_(True) for _ in nets]
    carries = [torch.randn(depth, b, hd, device=dev).mul_(0.5) for _ in nets]
    keep = (torch.rand(t, b, device=dev) > 0.1).float()
    gy = [torch.randn(t, b, hd, device=dev) for _ in nets]
    outs = replay.gru_stacks_forward(nets, xs, carries, keep)
    sum((y.float() * g).sum() for (y, _), g in zip(outs, gy)).backward()
    torch.cuda.synchronize()
    assert replay.last_status(stacks, t, b, False, dev) == 0 and replay.last_status(stacks, t, b, True, dev) == 0
    got = [[x.grad.clone()] + [p.grad.clone() for p in s.parameters()] for x, s in zip(xs, nets)]
    for x, s in zip(xs, nets):
        x.grad = None
        s.zero_grad()
    torch.backends.cuda.matmul.allow_tf32 = False
    refs = [replay.gru_stack_reference(s, x, c, keep) for s, x, c in zip(nets, xs, carries)]
    sum((y * g).sum() for (y, _), g in zip(refs, gy)).backward()
    for i in range(stacks):
        assert _rel(outs[i][0], refs[i][0]) < 3e-2
        assert _rel(outs[i][1], refs[i][1]) < 3e-2
        want = [xs[i].grad] + [p.grad for p in nets[i].parameters()]
        for g, w in zip(got[i], want):
            assert _rel(g, w) < 3e-2


@pytest.mark.gpu
def test_mixture_kernels_match_reference():
    torch.manual_seed(20)
    dev = torch.device("cuda:0")
    rows = 1000
    pred = (torch.randn(rows, 3 * SPEC.num_joints * SPEC.num_mixtures, device=dev) * 1.5).requires_grad_(True)
    x = torch.randn(rows, SPEC.num_joints, device=dev) * 0.5
    g = torch.randn(rows, SPEC.num_joints, device=dev)
    lp = policy.MixtureOfGaussians(pred, SPEC).log_prob(x)
    (lp * g).sum().backward()
    got = pred.grad.clone()
    pred.grad = None
    ref = policy.mixture_log_prob_reference(pred.double(), x.double(), SPEC)
    (ref * g.double()).sum().backward()
    torch.testing.assert_close(lp.double(), ref.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(got.double(), pred.grad.double(), rtol=1e-4, atol=1e-5)
    # one gradient per row, broadcast over the joints (what the PPO loss hands back)
    pred.grad = None
    grow = torch.randn(rows, device=dev)
    (policy.MixtureOfGaussians(pred, SPEC).log_prob(x) * grow[:, None]).sum().backward()
    got_row = pred.grad.clone()
    pred.grad = None
    policy.MixtureOfGaussians(pred, SPEC).log_prob(x).backward(grow[:, None].expand(-1, SPEC.num_joints))
    torch.testing.assert_close(pred.grad, got_row, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_mixture_sampling():
    torch.manual_seed(21)
    dev = torch.device("cuda:0")
    spec = policy.MixtureSpec(num_joints=2, num_mixtures=3)
    one = torch.tensor([0.5, -0.5, 0.0, 1.0, 0.2, -1.0,      # means  [J][M]
                        -1.0, 0.0, -2.0, -3.0, 0.5, -1.0,    # raw stds
                        1.0, 0.0, -1.0, -2.0, 0.0, 2.0])     # logits
    rows = 200_000
    pred = one.to(dev).repeat(rows, 1)
    dist = policy.MixtureOfGaussians(pred, spec)
    key = torch.tensor([7, 9], dtype=torch.int32, device=dev)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    a, lp = dist.sample(key, counter, with_log_prob=True)
    torch.testing.assert_close(lp, dist.log_prob(a), rtol=1e-6, atol=1e-6)
    mean, std, logw = policy.mixture_params_reference(one.double(), spec)
    w = logw.exp()
    want_mean = (w * mean).sum(-1)
    want_var = (w * (std**2 + mean**2)).sum(-1) - want_mean**2
    assert torch.allclose(a.double().mean(0).cpu(), want_mean, atol=4 * float(want_var.max().sqrt()) / math.sqrt(rows) + 1e-3)
    assert torch.allclose(a.double().var(0).cpu(), want_var, rtol=3e-2)
    counter += 1
    a2 = dist.sample(key, counter)
    assert float((a2 != a).float().mean()) > 0.99      # the device-side counter advances the stream
    counter -= 1
    torch.testing.assert_close(dist.sample(key, counter), a)   # and the same counter reproduces it
    mode = dist.mode()
    best = logw.argmax(-1)
    torch.testing.assert_close(mode[0].double().cpu(), mean.gather(-1, best[:, None])[:, 0], rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_optimizer_step_matches_reference():
    torch.manual_seed(22)
    dev = torch.device("cuda:0")
    cfg = optim.OptimizerConfig(learning_rate=1e-2, warmup_steps=3, adam_weight_decay=1e-3, grad_clip=2.0)
    m = policy.MLP(64, 33, 128, 2, "gelu").to(dev)
    fp = optim.FlatParams(m, cfg)
    p = fp.flat.double().cpu().numpy()
    mu, nu = np.zeros_like(p), np.zeros_like(p)
    for count in range(5):
        g = torch.randn(fp.numel, device=dev) * (0.001 if count == 1 else 0.05)
        if count == 3:
            g[17] = float("nan")
        fp.grad.copy_(g)
        scale = 0.5 if count == 4 else 1.0
        fp.step(scale)
        p, mu, nu, norm, lr = optim.optimizer_step_reference(p, g.double().cpu().numpy(), mu, nu, count, cfg, scale)
        torch.cuda.synchronize()
        assert float(fp.stats[0]) == pytest.approx(norm, rel=1e-5)
        assert float(fp.stats[1]) == pytest.approx(lr, rel=1e-6)
        np.testing.assert_allclose(fp.flat.cpu().numpy(), p, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(fp.mu.cpu().numpy(), mu, rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(fp.nu.cpu().numpy(), nu, rtol=1e-5, atol=1e-10)
        torch.testing.assert_close(fp.bf16.float(), fp.flat.to(torch.bfloat16).float(), rtol=0, atol=0)
    assert int(fp.step_count) == 5
    for q in m.parameters():   # still views of the flat buffers
        assert fp.flat.data_ptr() <= q.data_ptr() < fp.flat.data_ptr() + 4 * fp.numel


@pytest.mark.gpu
def test_get_ppo_variables_matches_reference_and_trains():
    torch.manual_seed(23)
    dev = torch.device("cuda:0")
    m = policy.Model(num_actor_inputs=43, num_critic_inputs=340, num_joints=17).to(dev)
    t, b = 6, 160
    oa, oc = torch.randn(t, b, 43, device=dev), torch.randn(t, b, 340, device=dev)
    act = torch.randn(t, b, 17, device=dev) * 0.3
    done = torch.rand(t, b, device=dev) < 0.1
    carry = {"actor": torch.randn(2, b, 512, device=dev) * 0.3, "critic": torch.randn(2, b, 512, device=dev) * 0.3}
    lp, v, c = policy.get_ppo_variables(m, oa, oc, act, done, carry)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        rlp, rv, rc = policy.get_ppo_variables_reference(m, oa, oc, act, done, carry)
    assert _rel(v, rv) < 3e-2 and _rel(c["actor"], rc["actor"]) < 3e-2 and _rel(c["critic"], rc["critic"]) < 3e-2
    assert float((lp - rlp).abs().max()) < 0.15 and float((lp - rlp).abs().mean()) < 0.02   # log-densities: absolute
    # gradients reach every parameter through the flat buffer, and a step changes the parameters
    fp = optim.FlatParams(m)
    lp, v, _ = policy.get_ppo_variables(m, oa, oc, act, done, carry)
    fp.zero_grad()
    (lp.mean() + v.mean()).backward()
    for name, q in m.named_parameters():
        assert float(q.grad.abs().sum()) > 0, name
    before = fp.flat.clone()
    fp.step()
    assert float((fp.flat - before).abs().max()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("use_graphs", [False, True])
def test_ppo_iteration_runs(walking_spec, walking_randomizers, use_graphs):
    from ksim_b200.engine import B200Engine
    from ksim_b200.examples import walking
    from ksim_b200.trainer import PPOConfig, PPOTrainer

    torch.manual_seed(24)
    dev = torch.device("cuda:0")
    eng = B200Engine(walking_spec, 256, device=dev, randomizers=walking_randomizers, seed=0)
    model = walking.make_model(walking_spec).to(dev)
    tr = PPOTrainer(eng, model, walking.network_columns(eng.program, walking.ACTOR_INPUTS),
                    walking.network_columns(eng.program, walking.CRITIC_INPUTS),
                    PPOConfig(batch_size=128, num_passes=2, rollout_length_frames=6), use_graphs=use_graphs)
    before = tr.flat.flat.clone()
    for _ in range(2):
        times = tr.iteration()
    assert all(math.isfinite(x) and x > 0 for x in times.values())
    assert int(tr.flat.step_count) == 2 * 2 * 2
    assert bool(torch.isfinite(tr.flat.flat).all()) and float((tr.flat.flat - before).abs().max()) > 0
    assert math.isfinite(float(tr.loss))
    assert bool(torch.isfinite(tr.rec).all())
    assert float(tr.actions.abs().sum()) > 0
    # the clipped means keep sampled actions near the joint ranges (std <= 1)
    assert float(tr.data["logp_on"].max()) < 10.0


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(1536, 12288), (512, 24 * 512), (7, 1000), (3, 13)])
def test_row_sums_match_torch(rows, cols):
    """``b200sim_row_sums`` (the GRU replay's bias gradients): fp32 sums of bf16 rows, against torch in float64; deterministic."""
    from ksim_b200.replay import _row_sums

    torch.manual_seed(5)
    x = (torch.randn(rows, cols, device="cuda:0") * 0.3).to(torch.bfloat16)
    got = _row_sums(x)
    want = x.double().sum(dim=1)
    scale = x.double().abs().sum(dim=1)
    assert float(((got.double() - want).abs() / scale).max()) < 2e-6
    assert torch.equal(got, _row_sums(x))
