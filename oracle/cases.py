"""Golden-case table shared by oracle/make_golden.py and the tests  (test infrastructure).

A case = module kind + ctor args + input shape + seed.  `build_case` regenerates the seeded
state / input / cotangent; `oracle_forward` runs the oracle restatement on it.
"""
import numpy as np
import torch

from oracle import gcn_oracle as O
from tam_gcn_b200.graph import ucla, ntu_rgb_d

GRAPHS = {'graph.ucla.Graph': ucla.Graph, 'graph.ntu_rgb_d.Graph': ntu_rgb_d.Graph}


def _A(name):
    return GRAPHS[name]().A


def case(kind, graph='graph.ucla.Graph', N=2, T=8, seed=0, train=True, **args):
    return dict(kind=kind, graph=graph, N=N, T=T, seed=seed, train=train, args=args)


CASES = {
    # CTRGC (models/ctrgcn.py:150-177)
    'ctrgc_64_64': case('ctrgc', cin=64, cout=64, seed=1),
    'ctrgc_3_64': case('ctrgc', cin=3, cout=64, T=10, seed=2),
    'ctrgc_128_256_ntu': case('ctrgc', graph='graph.ntu_rgb_d.Graph', cin=128, cout=256, T=4, seed=3),
    # unit_gcn (:196-263)
    'unit_gcn_64_64': case('unit_gcn', cin=64, cout=64, seed=4),
    'unit_gcn_3_64': case('unit_gcn', cin=3, cout=64, T=12, seed=5),
    'unit_gcn_64_128_ntu': case('unit_gcn', graph='graph.ntu_rgb_d.Graph', cin=64, cout=128, T=6, seed=6),
    'unit_gcn_64_64_eval': case('unit_gcn', cin=64, cout=64, seed=7, train=False),
    # MultiScale_TemporalConv (:72-147)
    'ms_tcn_64_s1': case('ms_tcn', cin=64, cout=64, kernel_size=5, stride=1, dilations=(1, 2), residual=False, T=12, seed=8),
    'ms_tcn_128_s2': case('ms_tcn', cin=128, cout=128, kernel_size=5, stride=2, dilations=(1, 2), residual=False, T=12, seed=9),
    'ms_tcn_general': case('ms_tcn', cin=48, cout=48, kernel_size=3, stride=1, dilations=(1, 2, 3, 4), residual=True, T=12, seed=10),
    'ms_tcn_general_s2_resconv': case('ms_tcn', cin=32, cout=96, kernel_size=[3, 5, 3, 5], stride=2,
                                      dilations=(1, 2, 3, 4), residual=True, T=13, seed=11),
    'ms_tcn_64_eval': case('ms_tcn', cin=64, cout=64, kernel_size=5, stride=1, dilations=(1, 2), residual=False, T=12,
                           seed=12, train=False),
    # unit_tcn (:179-193)
    'unit_tcn_k1_s2': case('unit_tcn', cin=64, cout=128, kernel_size=1, stride=2, T=12, seed=13),
    'unit_tcn_k9': case('unit_tcn', cin=32, cout=32, kernel_size=9, stride=1, T=16, seed=14),
    # TCN_GCN_unit (:266-284)
    'tcn_gcn_64_64': case('tcn_gcn_unit', cin=64, cout=64, stride=1, residual=True, T=12, seed=15),
    'tcn_gcn_64_128_s2': case('tcn_gcn_unit', cin=64, cout=128, stride=2, residual=True, T=12, seed=16),
    'tcn_gcn_3_64_nores': case('tcn_gcn_unit', cin=3, cout=64, stride=1, residual=False, T=12, seed=17),
    'tcn_gcn_256_ntu': case('tcn_gcn_unit', graph='graph.ntu_rgb_d.Graph', cin=256, cout=256, stride=1, residual=True,
                            T=8, seed=18),
    # whole CTR-GCN (:287-374)
    'ctrgcn_ucla_train': dict(case('ctrgcn_model', N=4, T=52, seed=19, num_class=10, num_point=20, num_person=1,
                                   graph='graph.ucla.Graph'), store_grads=False,
                              keep_grads=('l1.gcn1.alpha', 'l5.gcn1.PA', 'l10.gcn1.convs.1.conv4.weight', 'fc.weight',
                                          'l3.tcn1.branches.0.3.conv.weight', 'l8.residual.conv.weight',
                                          'l2.gcn1.offset_conv.0.weight', 'data_bn.weight')),
    'ctrgcn_ntu_m2_train': dict(case('ctrgcn_model', graph='graph.ntu_rgb_d.Graph', N=2, T=16, seed=20, num_class=60,
                                     num_point=25, num_person=2), store_grads=False,
                                keep_grads=('l1.gcn1.alpha', 'fc.weight', 'l6.gcn1.convs.2.conv1.weight')),
    # ST-GCN (models/stgcn.py)
    'ctg_64_64_ntu': case('ctg', graph='graph.ntu_rgb_d.Graph', cin=64, cout=64, K=3, T=10, seed=21),
    'st_gcn_64_64_ntu': case('st_gcn', graph='graph.ntu_rgb_d.Graph', cin=64, cout=64, K=3, stride=1, residual=True, T=20, seed=22),
    'st_gcn_64_128_s2_ntu': case('st_gcn', graph='graph.ntu_rgb_d.Graph', cin=64, cout=128, K=3, stride=2, residual=True, T=20, seed=23),
    'st_gcn_3_64_nores': case('st_gcn', cin=3, cout=64, K=3, stride=1, residual=False, T=20, seed=24),
    'stgcn_ntu_train': dict(case('stgcn_model', graph='graph.ntu_rgb_d.Graph', N=2, T=32, seed=25, num_class=60, num_point=25),
                            store_grads=False, keep_grads=('edge_importance.3', 'fcn.weight', 'st_gcn_networks.4.residual.0.weight')),
    # cross-modal fusion head (models/resnet_gcn_attention.py): frozen CTR-GCN -> attention gate -> gated pool -> classifier;
    # the backbone output f_rgb is an input of the case (the ResNet itself is outside the hot path)
    'fusion_ucla_train': dict(case('fusion_model', N=4, T=52, seed=26, num_class=10, num_point=20, num_person=1,
                                   graph='graph.ucla.Graph'), store_grads=False,
                              keep_grads=('attention_transform.0.bias', 'attention_transform.1.weight', 'classifier.weight',
                                          'attention_transform.3.bias')),
}
for _c in CASES.values():
    _c['args'].setdefault('graph', _c['graph'])
    _c['A'] = _A(_c['graph'])


def build_case(c):
    """-> dict(state fp32-valued, x, cot, extra) — all float32 tensors, deterministic from the seed."""
    kind, a, A, seed = c['kind'], c['args'], c['A'], c['seed']
    g = torch.Generator().manual_seed(1000 + seed)
    V = A.shape[1]
    p = {}
    extra = {}
    N, T = c['N'], c['T']
    if kind == 'ctrgc':
        O.add_ctrgc(p, 'm', a['cin'], a['cout'], g)
        extra['A'] = torch.as_tensor(A[1], dtype=torch.float32) + 0.02 * torch.randn(V, V, generator=g)
        extra['alpha'] = torch.full((1,), 0.7)
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], T, V)
    elif kind == 'unit_gcn':
        O.add_unit_gcn(p, 'm', a['cin'], a['cout'], A, g)
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], T, V)
    elif kind == 'ms_tcn':
        resconv = a['residual'] and not (a['cin'] == a['cout'] and a['stride'] == 1)
        O.add_ms_tcn(p, 'm', a['cin'], a['cout'], a['kernel_size'], a['dilations'], g, resconv,
                     a.get('residual_kernel_size', 1))
        To = (T - 1) // a['stride'] + 1
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], To, V)
    elif kind == 'unit_tcn':
        O._add_conv(p, 'm.conv', a['cout'], a['cin'], a['kernel_size'], g)
        O._add_bn(p, 'm.bn', a['cout'], g)
        To = (T - 1) // a['stride'] + 1
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], To, V)
    elif kind == 'tcn_gcn_unit':
        O.add_tcn_gcn_unit(p, 'm', a['cin'], a['cout'], A, a['stride'], a['residual'], g)
        To = (T - 1) // a['stride'] + 1
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], To, V)
    elif kind == 'ctrgcn_model':
        p = O.make_ctrgcn_state(A, a['num_class'], a['num_person'], seed=seed, dtype=torch.float64)
        xs, ys = (N, 3, T, V, a['num_person']), (N, a['num_class'])
    elif kind == 'ctg':
        O._add_conv(p, 'm.conv', a['cout'] * a['K'], a['cin'], 1, g)
        extra['A'] = torch.as_tensor(A, dtype=torch.float32) * (1 + 0.1 * torch.randn(A.shape, generator=g))
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], T, V)
    elif kind == 'st_gcn':
        K = a['K']
        O._add_conv(p, 'm.gcn.conv', a['cout'] * K, a['cin'], 1, g)
        O._add_bn(p, 'm.tcn.0', a['cout'], g)
        O._add_conv(p, 'm.tcn.2', a['cout'], a['cout'], 9, g)
        O._add_bn(p, 'm.tcn.3', a['cout'], g)
        if a['residual'] and not (a['cin'] == a['cout'] and a['stride'] == 1):
            O._add_conv(p, 'm.residual.0', a['cout'], a['cin'], 1, g)
            O._add_bn(p, 'm.residual.1', a['cout'], g)
        extra['A'] = torch.as_tensor(A, dtype=torch.float32) * (1 + 0.1 * torch.randn(A.shape, generator=g))
        To = (T - 1) // a['stride'] + 1
        xs, ys = (N, a['cin'], T, V), (N, a['cout'], To, V)
    elif kind == 'stgcn_model':
        p = O.make_stgcn_state(A, a['num_class'], seed=seed, dtype=torch.float64)
        xs, ys = (N, 3, T, V, 2), (N, a['num_class'])
    elif kind == 'fusion_model':
        pg = O.make_ctrgcn_state(A, a['num_class'], a['num_person'], seed=seed, dtype=torch.float64)
        p = {'gcn.' + k: v for k, v in pg.items()}
        O.add_fusion_head(p, g, a['num_class'])
        extra['f_rgb'] = torch.relu(torch.randn(N, 2048, 7, 7, generator=g))
        xs, ys = (N, 3, T, V, a['num_person']), (N, a['num_class'])
    else:
        raise KeyError(kind)
    state = O.cast_state(p, torch.float32)          # fp32-representable values, shared by all precisions
    if len(xs) == 5:
        x = O.synthetic_skeletons(*[xs[i] for i in (0, 2, 3, 4)], C=xs[1], seed=seed)
    else:
        x = torch.randn(xs, generator=g)
    cot = torch.randn(ys, generator=g)
    return dict(state=state, x=x, cot=cot, extra=extra)


def oracle_forward(c, x, p, extra):
    kind, a, train = c['kind'], c['args'], c['train']
    V = c['A'].shape[1]
    if kind == 'ctrgc':
        return O.ctrgc(x, p, 'm', extra['A'], extra['alpha'])
    if kind == 'unit_gcn':
        return O.unit_gcn(x, p, 'm', train)
    if kind == 'ms_tcn':
        return O.ms_tcn(x, p, 'm', a['kernel_size'], a['stride'], a['dilations'], a['residual'],
                        a.get('residual_kernel_size', 1), train)
    if kind == 'unit_tcn':
        return O.unit_tcn(x, p, 'm', a['kernel_size'], a['stride'], train)
    if kind == 'tcn_gcn_unit':
        return O.tcn_gcn_unit(x, p, 'm', a['stride'], a['residual'], train=train)
    if kind == 'ctrgcn_model':
        return O.ctrgcn_forward(x, p, V, train)
    if kind == 'ctg':
        return O.conv_temporal_graphical(x, extra['A'], p, 'm', a['K'])
    if kind == 'st_gcn':
        return O.st_gcn_block(x, extra['A'], p, 'm', a['stride'], a['residual'], train=train)
    if kind == 'stgcn_model':
        return O.stgcn_forward(x, p, V, train)
    if kind == 'fusion_model':
        return O.fusion_forward(x, extra['f_rgb'], p, V, train)
    raise KeyError(kind)
