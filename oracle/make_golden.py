"""Pin the oracle against the UNMODIFIED reference and write golden fixtures.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every case it (1) builds a perturbed, reference-named state with the oracle's own seeded
generators, (2) loads it with strict=True into the reference module imported from
/root/reference (this checks the App. B state_dict contract), (3) runs reference and oracle
in fp64 on the same input/cotangent, asserts they agree to round-off (<=1e-11 rel), and
(4) stores input, cotangent, reference outputs / gradients / running stats as float32 in
tests/golden/<case>.pt.  States are NOT stored (they are regenerated from the seed; a
checksum guards against RNG drift).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(1, '/root/reference')
sys.dont_write_bytecode = True

from oracle import gcn_oracle as O          # noqa: E402
from oracle.cases import CASES, build_case  # noqa: E402

import models.ctrgcn as RC                   # noqa: E402  (reference)
import models.stgcn as RS                    # noqa: E402  (reference)

OUT = os.path.join(ROOT, 'tests', 'golden')
D = torch.float64


def strip(p, pre):
    n = len(pre) + 1
    return {k[n:]: v for k, v in p.items() if k.startswith(pre + '.')}


def ref_module(case, state):
    kind, a = case['kind'], case['args']
    A = case['A']
    if kind == 'ctrgc':
        m = RC.CTRGC(a['cin'], a['cout'])
    elif kind == 'unit_gcn':
        m = RC.unit_gcn(a['cin'], a['cout'], A, residual=a.get('residual', True))
    elif kind == 'ms_tcn':
        m = RC.MultiScale_TemporalConv(a['cin'], a['cout'], kernel_size=a['kernel_size'], stride=a['stride'],
                                       dilations=list(a['dilations']), residual=a['residual'],
                                       residual_kernel_size=a.get('residual_kernel_size', 1))
    elif kind == 'unit_tcn':
        m = RC.unit_tcn(a['cin'], a['cout'], kernel_size=a['kernel_size'], stride=a['stride'])
    elif kind == 'tcn_gcn_unit':
        m = RC.TCN_GCN_unit(a['cin'], a['cout'], A, stride=a['stride'], residual=a['residual'])
    elif kind == 'ctrgcn_model':
        m = RC.Model(num_class=a['num_class'], num_point=a['num_point'], num_person=a['num_person'],
                     graph=a['graph'], graph_args=dict(labeling_mode='spatial'))
    elif kind == 'ctg':
        m = RS.ConvTemporalGraphical(a['cin'], a['cout'], a['K'])
    elif kind == 'st_gcn':
        m = RS.st_gcn(a['cin'], a['cout'], (9, a['K']), a['stride'], residual=a['residual'])
    elif kind == 'stgcn_model':
        m = RS.Model(in_channels=3, num_class=a['num_class'], num_point=a['num_point'], num_person=1,
                     graph=a['graph'], graph_args=dict(labeling_mode='spatial'))
    elif kind == 'fusion_model':
        # the reference asks torch.hub for ImageNet weights (models/resnet_gcn_attention.py:32): no network here, and the
        # backbone is outside the hot path — build it unpretrained and replace its stages by identities so that the
        # UNMODIFIED forward (:82-120) runs the head on the f_rgb the case supplies as `x_rgb`
        import models.resnet as RR
        import models.resnet_gcn_attention as RF
        real = RR.resnet50
        RF.resnet50 = lambda pretrained=True: real(pretrained=False)
        try:
            m = RF.ResNet_GCN_Attention(num_class=a['num_class'], num_point=a['num_point'], num_person=a['num_person'],
                                        graph=a['graph'], graph_args=dict(labeling_mode='spatial'), in_channels_rgb=3)
        finally:
            RF.resnet50 = real
        r = torch.nn.Module()
        for nm in ('conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4'):
            setattr(r, nm, torch.nn.Identity())
        m.resnet = r
    else:
        raise KeyError(kind)
    m = m.double()
    sd = strip(state, 'm') if not kind.endswith('_model') else state
    sd = {k: v for k, v in sd.items() if not k.startswith('__')}
    missing, unexpected = m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    assert not missing and not unexpected
    return m


def ref_forward(case, m, x, extra):
    kind = case['kind']
    if kind == 'ctrgc':
        return m(x, extra['A'], extra['alpha'])
    if kind == 'ctg':
        return m(x, extra['A'])[0]
    if kind == 'st_gcn':
        return m(x, extra['A'])[0]
    if kind == 'fusion_model':
        return m(x, extra['f_rgb'])
    return m(x)


def main():
    os.makedirs(OUT, exist_ok=True)
    worst = 0.0
    only = [a for a in sys.argv[1:] if not a.startswith('-')]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        built = build_case(case)
        state64 = O.clone_state(built['state'], D)
        m = ref_module(case, state64)
        m.train(case['train'])
        x = built['x'].to(D).requires_grad_(True)
        extra = {k: v.to(D).requires_grad_(True) for k, v in built['extra'].items()}
        if case['kind'] == 'ctrgc':
            # alpha / A are arguments of CTRGC.forward, not module state
            pass
        y = ref_forward(case, m, x, extra)
        ct = built['cot'].to(D)
        y.backward(ct)
        ref = {'y': y.detach(), 'dx': x.grad.detach()}
        grads = {k: p.grad.detach() for k, p in m.named_parameters() if p.grad is not None}
        for k, v in extra.items():
            grads['__' + k] = v.grad.detach()
        bufs = {k: b.detach().clone() for k, b in m.named_buffers()}

        # --- oracle on the same data ---
        from oracle.cases import oracle_forward
        po = O.clone_state(built['state'], D, requires_grad=True)
        xo = built['x'].to(D).requires_grad_(True)
        eo = {k: v.to(D).requires_grad_(True) for k, v in built['extra'].items()}
        yo = oracle_forward(case, xo, po, eo)
        yo.backward(ct)
        e_y = O.rel_err(yo, ref['y'])
        e_dx = O.rel_err(xo.grad, ref['dx'])
        pre = '' if case['kind'].endswith('_model') else 'm.'
        e_g = 0.0
        for k, gref in grads.items():
            go = eo[k[2:]].grad if k.startswith('__') else po[pre + k].grad
            assert go is not None, k
            if gref.norm() > 1e-9 * max(1.0, float(ct.norm())):
                e_g = max(e_g, O.rel_err(go, gref))
        e_b = 0.0
        for k, b in bufs.items():
            if b.is_floating_point():
                e_b = max(e_b, O.rel_err(po[pre + k], b))
            else:
                assert int(po[pre + k]) == int(b), k
        print(f'{name:28s} y {e_y:.2e} dx {e_dx:.2e} dW {e_g:.2e} buf {e_b:.2e}')
        assert max(e_y, e_dx, e_b) < 1e-11 and e_g < 1e-9, name
        worst = max(worst, e_y, e_dx, e_g, e_b)

        fx = {'case': name,
              'state_checksum': float(sum(v.double().abs().sum() for v in built['state'].values()
                                          if torch.is_tensor(v) and v.is_floating_point())),
              'x': built['x'], 'cot': built['cot'],
              # the 1.6 MB backbone feature map of the fusion case is regenerated from the seed, not stored
              'extra': {} if case['kind'] == 'fusion_model' else built['extra'],
              'y': ref['y'].float(), 'dx': ref['dx'].float()}
        if case.get('store_grads', True):
            fx['grads'] = {k: v.float() for k, v in grads.items()}
        else:
            fx['grad_norms'] = {k: float(v.norm()) for k, v in grads.items()}
            keep = case.get('keep_grads', ())
            fx['grads'] = {k: v.float() for k, v in grads.items() if k in keep}
        fx['buffers'] = {k: (v.float() if v.is_floating_point() else v) for k, v in bufs.items()}
        torch.save(fx, os.path.join(OUT, name + '.pt'))
    print('worst oracle-vs-reference rel err', worst)


if __name__ == '__main__':
    main()
