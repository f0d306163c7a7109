"""Parity of the B200-native modules (CUDA kernels through the C-ABI) with the reference, on the golden
fixtures dumped from the unmodified reference modules (fp64 run stored as fp32) and against the fp64 oracle.

Tolerances: fp32 — outputs <= 1e-4 relative (we assert 2e-5), module-level gradients <= 1e-4
relative except the documented fp32 noise floor of the BatchNorm-weight gradients of the max-pool branch
(3e-5 in the reference itself) -> 5e-4; end-to-end gradients through 10 layers are graded against the
reference's own fp32-vs-fp64 error (App. D: 2e-3), asserted <= 2e-2, with identical top-1.
bf16 activations (stated separately): outputs <= 3e-2 relative, input gradient <= 1e-1, the whole parameter-gradient
vector <= 5e-2 relative; single cancellation-heavy tensors (BatchNorm weights of the max-pool branch, alpha: fp32
noise floor already 3e-5) are only bounded at 0.5.
"""
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

MODULE_CASES = [n for n, c in H.CASES.items() if not c['kind'].endswith('_model')]
MODEL_CASES = [n for n, c in H.CASES.items() if c['kind'].endswith('_model')]


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')


# Gradient tensors whose fp32 error sits above 1e-4 IN THE REFERENCE ITSELF (reference fp32 vs fp64, SURVEY.md App. D:
# noise floor 2e-7 ... 7e-5, worst on cancellation-dominated sums): BatchNorm weights behind the max-pool / in front of
# a second BatchNorm, alpha (a sum of ~1e6 signed terms) and the conv1/conv2 relation weights behind tanh'.
NOISE_FLOOR = ('.bn.weight', '.1.weight', '.4.weight', 'alpha', '.conv1.weight', '.conv2.weight', '.conv1.bias', '.conv2.bias',
               'bn.bias', '.1.bias', '.4.bias')


@pytest.mark.parametrize('name', MODULE_CASES)
def test_module_fp32_matches_reference(name):
    """Module level, fp32: outputs <= 2e-5, input gradient and parameter gradients <= 1e-4 relative to the fp64
    fixtures, except the documented noise-floor tensors (<= 5e-4)."""
    _cuda()
    res = H.run_case(name, 'cuda')
    rep = []
    fx = H.load_fixture(name)
    allow = {k: 5e-4 for k in fx['grads'] if k.endswith(NOISE_FLOOR)}
    fails = H.compare(name, res, fx, tol_y=2e-5, tol_dx=1e-4, tol_g=1e-4, tol_buf=1e-5, report=rep, tol_gk=allow,
                      tol_gall=1e-4)
    print(rep[0])
    assert not fails, '\n'.join(fails)


def _yardstick():
    import json
    import os
    with open(os.path.join(H.GOLDEN, 'bf16_yardstick.json')) as f:
        return json.load(f)


@pytest.mark.parametrize('name', MODULE_CASES)
def test_module_bf16_matches_reference(name):
    """bf16 activations: graded against the reference math under torch.autocast(bfloat16)
    (tests/golden/bf16_yardstick.json, made by oracle/make_bf16_yardstick.py): outputs, input gradient and the
    whole parameter-gradient vector no worse than 2x that yard-stick (floor 1e-2); single cancellation-heavy
    tensors (BN weights of the max-pool branch, alpha) only bounded at max(0.6, 3x their own yard-stick error) - the
    yard-stick itself exceeds 1.0 on them (bf16 noise larger than the true gradient)."""
    _cuda()
    ys = _yardstick()[name]
    res = H.run_case(name, 'cuda', torch.bfloat16)
    rep = []
    lim = lambda v: max(2.0 * v, 1e-2)
    fails = H.compare(name, res, H.load_fixture(name), tol_y=lim(ys['y']), tol_dx=lim(ys['dx']), tol_g=0.6,
                      tol_buf=2e-2, report=rep, tol_gall=lim(ys['gall']),
                      tol_gk={k: 3.0 * v for k, v in ys.get('g', {}).items()})
    print('bf16 ' + rep[0] + '  (yard-stick y %.1e dx %.1e gall %.1e)' % (ys['y'], ys['dx'], ys['gall']))
    assert not fails, '\n'.join(fails)


@pytest.mark.parametrize('name', MODEL_CASES)
def test_model_fp32_matches_reference(name):
    _cuda()
    res = H.run_case(name, 'cuda')
    fx = H.load_fixture(name)
    rep = []
    # end-to-end gradients sit on the fp32 noise floor (reference fp32 vs fp64: ~2e-3 on dx); with
    # atomics-ordered sums an occasional ReLU / max-pool gate flip moves one small tensor further: bound 5e-2, typical 7e-3
    fails = H.compare(name, res, fx, tol_y=1e-4, tol_dx=2e-2, tol_g=5e-2, tol_buf=1e-4, report=rep)
    print(rep[0])
    assert (res['y'].argmax(1) == fx['y'].argmax(1)).all(), 'top-1 differs'
    assert not fails, '\n'.join(fails)


@pytest.mark.parametrize('name', MODEL_CASES)
def test_model_bf16_close_to_reference(name):
    """Whole models with bf16 activations on the seeded golden state: logits <= 6e-2 (the state is a chaotic regime:
    the reference under autocast is at 1.1e-2 ... 1.5e-2), identical top-1 away from near-ties, and GRADIENTS no worse
    than 2x the reference-under-autocast yard-stick (tests/golden/bf16_yardstick.json) for the input gradient and the
    kept parameter gradients.  The benchmarked regime (fresh init, batch 64) is graded in test_parity_large_gpu.py."""
    _cuda()
    res = H.run_case(name, 'cuda', torch.bfloat16)
    fx = H.load_fixture(name)
    ys = _yardstick()[name]
    e = H.O.rel_err(res['y'], fx['y'])
    e_dx = H.O.rel_err(res['dx'], fx['dx'])
    ks = [k for k in fx['grads'] if k in res['grads']]
    ga = torch.cat([res['grads'][k].reshape(-1) for k in ks])
    gb = torch.cat([fx['grads'][k].reshape(-1) for k in ks])
    e_g = H.O.rel_err(ga, gb)
    print('bf16 %s logits %.3e dx %.3e grads %.3e  (yard-stick y %.1e dx %.1e gall %.1e)' % (name, e, e_dx, e_g, ys['y'], ys['dx'], ys['gall']))
    assert e < 6e-2
    assert e_dx <= max(2.0 * ys['dx'], 5e-2) and e_g <= max(2.0 * ys['gall'], 5e-2)
    # top-1 identical wherever the reference margin is not a near-tie
    top2 = fx['y'].topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.05 * fx['y'].abs().max()
    assert (res['y'].argmax(1)[clear] == fx['y'].argmax(1)[clear]).all()


def test_state_dict_contract():
    """Keys and shapes of App. B (892 entries / 1,693,260 parameters for CTR-GCN on NW-UCLA); round trip."""
    import tam_gcn_b200.ctrgcn as C
    m = C.Model(num_class=10, num_point=20, num_person=1, graph='graph.ucla.Graph',
                graph_args=dict(labeling_mode='spatial'))
    sd = m.state_dict()
    assert len(sd) == 892 and sum(p.numel() for p in m.parameters()) == 1693260
    case = H.CASES['ctrgcn_ucla_train']
    state = H.build_case(case)['state']
    assert set(sd) == set(state)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(state[k].shape), k


def test_eval_mode_and_frozen_parameters():
    """extract_feature with a frozen GCN left in train() mode (models/resnet_gcn_attention.py:24-26,82): forward
    only, batch statistics, running stats still updated; and eval-mode forward of the whole model."""
    _cuda()
    import tam_gcn_b200.ctrgcn as C
    from oracle import gcn_oracle as O
    case = H.CASES['ctrgcn_ucla_train']
    built = H.build_case(case)
    m = H.load_state(H.our_module(case), case, built['state']).cuda()
    for p in m.parameters():
        p.requires_grad_(False)
    m.train()
    x = built['x'].cuda()
    f, f2 = m.extract_feature(x)
    p64 = O.clone_state(built['state'], torch.float64)
    ref = O.ctrgcn_extract_feature(built['x'].double(), p64, 20, train=True)
    assert f.shape == ref.shape == (4, 256, 13, 20, 1)
    assert O.rel_err(f, ref) < 1e-4
    assert int(m.data_bn.num_batches_tracked) == 1 and int(m.l3.gcn1.bn.num_batches_tracked) == 1
    assert O.rel_err(m.l7.tcn1.branches[2][4].running_var, p64['l7.tcn1.branches.2.4.running_var']) < 1e-5
    # eval mode needs calibrated running statistics (with the seeded random ones activations reach
    # 1e7 and the reference's own fp32-vs-fp64 error is 3e-3): calibrate with 20 train-mode oracle forwards,
    # load that state into our model, compare eval logits.
    for _ in range(20):
        O.ctrgcn_forward(built['x'].double(), p64, 20, train=True)
    m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in p64.items()}, strict=True)
    m.eval()
    with torch.no_grad():
        y = m(x)
    yref = O.ctrgcn_forward(built['x'].double(), p64, 20, train=False)
    assert O.rel_err(y, yref) < 1e-4
    assert (y.argmax(1).cpu() == yref.argmax(1)).all()


def test_input_gradient_in_eval_mode_stgcn():
    """Saliency pass of tools/train_stgcn_group.py:292-309: d(score)/d(input) through ST-GCN."""
    _cuda()
    from oracle import gcn_oracle as O
    case = H.CASES['stgcn_ntu_train']
    built = H.build_case(case)
    p64 = O.clone_state(built['state'], torch.float64)
    for _ in range(20):        # calibrate running statistics (see test_eval_mode_and_frozen_parameters)
        O.stgcn_forward(built['x'].double(), p64, 25, train=True)
    m = H.our_module(case)
    m.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in p64.items()}, strict=True)
    m = m.cuda().eval()
    x = built['x'].cuda().requires_grad_(True)
    label = torch.arange(x.shape[0], device='cuda') % 60
    out = m(x)
    torch.gather(out, 1, label.unsqueeze(1)).squeeze().sum().backward()
    x64 = built['x'].double().requires_grad_(True)
    o64 = O.stgcn_forward(x64, p64, 25, train=False)
    torch.gather(o64, 1, label.cpu().unsqueeze(1)).squeeze().sum().backward()
    assert O.rel_err(out, o64) < 1e-4
    # fp32 round-off through 10 layers with ReLU gates: the reference's own fp32-vs-fp64 input gradient differs by
    # ~2e-3 relative, a single flipped gate moves a whole neighbourhood
    assert O.rel_err(x.grad, x64.grad) < 1e-2
    sal, sal_ref = x.grad.abs().sum(dim=(1, 2, 4)).cpu(), x64.grad.abs().sum(dim=(1, 2, 4))
    assert O.rel_err(sal, sal_ref) < 5e-3


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
def test_st_gcn_dropout_training_path(dt):
    """st_gcn(dropout > 0) in training mode (models/stgcn.py:81): the mask sits between BatchNorm and the residual sum.
    Checked against the same modules composed in torch with the same RNG state, and for gradient flow; in eval mode the
    block must equal the dropout-free one."""
    _cuda()
    import tam_gcn_b200
    from tam_gcn_b200 import stgcn as S
    tam_gcn_b200.set_act_dtype(dt)
    try:
        torch.manual_seed(0)
        blk = S.st_gcn(16, 32, (9, 3), stride=2, dropout=0.4).cuda()
        ref = S.st_gcn(16, 32, (9, 3), stride=2, dropout=0).cuda()
        ref.load_state_dict(blk.state_dict())
        x = torch.randn(3, 16, 20, 25, device='cuda').to(dt)
        A = torch.rand(3, 25, 25, device='cuda') * 0.2
        blk.eval(); ref.eval()
        ye, _ = blk(x, A)
        yr, _ = ref(x, A)
        assert torch.equal(ye, yr)
        blk.train()
        xg = x.clone().requires_grad_(True)
        torch.manual_seed(7)
        y, _ = blk(xg, A)
        assert y.shape == (3, 32, 10, 25) and y.dtype == dt and torch.isfinite(y.float()).all()
        y.float().sum().backward()
        assert torch.isfinite(xg.grad.float()).all() and float(xg.grad.float().abs().sum()) > 0
        for n_, p_ in blk.named_parameters():
            assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_
        # the dropout really drops: an undropped evaluation of the same (training-mode) block differs
        blk.tcn[4].p = 0.0
        y0, _ = blk(x, A)
        blk.tcn[4].p = 0.4
        assert float((y.detach().float() - y0.detach().float()).abs().max()) > 1e-3
    finally:
        tam_gcn_b200.set_act_dtype(torch.float32)


def test_three_dim_input_and_python_alpha():
    _cuda()
    import tam_gcn_b200.ctrgcn as C
    from oracle import gcn_oracle as O
    case = H.CASES['ctrgc_64_64']
    built = H.build_case(case)
    m = H.load_state(H.our_module(case), case, built['state']).cuda()
    x = built['x'].cuda()
    y = m(x)                       # A=None, alpha=1 (python scalar)
    p64 = O.clone_state(built['state'], torch.float64)
    assert O.rel_err(y, O.ctrgc(built['x'].double(), p64, 'm', None, 1)) < 2e-5
    # (N, T, V*C) input of Model.forward (models/ctrgcn.py:325-327)
    mcase = H.CASES['ctrgcn_ucla_train']
    mb = H.build_case(mcase)
    mm = H.load_state(H.our_module(mcase), mcase, mb['state']).cuda().eval()
    x5 = mb['x'].cuda()
    x3 = x5[..., 0].permute(0, 2, 3, 1).reshape(x5.shape[0], x5.shape[2], -1)
    with torch.no_grad():
        assert torch.allclose(mm(x3), mm(x5), atol=1e-5, rtol=1e-5)


def test_patch_reference_rebinds_module_globals():
    """`tam_gcn_b200.patch_reference` on a module whose network class looks its layers up as module globals at
    construction time (what models/ctrgcn.py does; tests/ref_standin.py stands in for it on the GPU box, the real
    reference is exercised by tests/test_reference_dropin_cpu.py): the foreign Model, with its own ATen prologue / head,
    runs on the native layers and agrees with tam_gcn_b200.ctrgcn.Model and with the golden fixture."""
    _cuda()
    import ref_standin
    import tam_gcn_b200
    tam_gcn_b200.patch_reference(ctrgcn_module=ref_standin)
    import tam_gcn_b200.ctrgcn as C
    assert ref_standin.TCN_GCN_unit is C.TCN_GCN_unit and ref_standin.CTRGC is C.CTRGC
    case = H.CASES['ctrgcn_ucla_train']
    built = H.build_case(case)
    m = ref_standin.Model(case['A'])
    m.load_state_dict({k: v.clone() for k, v in built['state'].items()}, strict=True)
    m = m.cuda().train()
    x = built['x'].cuda().requires_grad_(True)
    y = m(x)
    y.backward(built['cot'].cuda())
    fx = H.load_fixture('ctrgcn_ucla_train')
    e_y, e_dx = H.O.rel_err(y, fx['y']), H.O.rel_err(x.grad, fx['dx'])
    print('patched stand-in Model: y %.2e dx %.2e' % (e_y, e_dx))
    assert e_y < 1e-4 and e_dx < 2e-2 and (y.argmax(1).cpu() == fx['y'].argmax(1)).all()
