"""B200-native CTR-GCN modules: drop-in replacements for the classes of the reference's models/ctrgcn.py.

Same class names, constructor signatures, forward signatures, attribute names and `state_dict`
keys/shapes as the reference (models/ctrgcn.py:52-374), so

  * `--model tam_gcn_b200.ctrgcn.Model` works through the reference's `torchlight.import_class`
    (torchlight/torchlight/io.py:51-55), and
  * `tam_gcn_b200.patch_reference()` can rebind `models.ctrgcn.{CTRGC, unit_gcn, ...}` so the
    reference's own `models.ctrgcn.Model` runs unchanged on these layers.

The modules own ordinary `nn.Conv2d` / `nn.BatchNorm2d` children purely as PARAMETER CONTAINERS (that
is what fixes the state_dict layout); their `forward` is never called.  All computation goes through
the autograd functions of `tam_gcn_b200.functional`, i.e. hand-written sm_100a kernels.  CUDA only.
"""
import importlib
import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn


def import_class(name):
    """Resolve a dotted class path.  `graph.ucla.Graph` / `graph.ntu_rgb_d.Graph` (the reference's own
    graph modules, reference graph/ucla.py, graph/ntu_rgb_d.py) resolve to this package's builders when
    the reference tree is not importable."""
    mod_name, _, cls = name.rpartition('.')
    try:
        return getattr(importlib.import_module(mod_name), cls)
    except ImportError:
        if mod_name.startswith('graph.'):
            return getattr(importlib.import_module('tam_gcn_b200.' + mod_name), cls)
        raise


def _kaiming_conv(conv):
    nn.init.kaiming_normal_(conv.weight, mode='fan_out')
    if conv.bias is not None:
        nn.init.zeros_(conv.bias)


def _const_bn(bn, scale):
    nn.init.constant_(bn.weight, scale)
    nn.init.zeros_(bn.bias)


def _ms_tcn_reinit(m):
    # reference weights_init (models/ctrgcn.py:38-49), applied by MultiScale_TemporalConv to itself
    if isinstance(m, nn.Conv2d):
        _kaiming_conv(m)
    elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.zero_()


class TemporalConv(nn.Module):
    """(k x 1) dilated, strided temporal convolution followed by BatchNorm (models/ctrgcn.py:52-69)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1):
        super().__init__()
        pad = (kernel_size + (kernel_size - 1) * (dilation - 1) - 1) // 2
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=(kernel_size, 1), padding=(pad, 0),
                              stride=(stride, 1), dilation=(dilation, 1))
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        return Fn.ConvBnFn.apply(x, self.conv, self.bn, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias)


class MultiScale_TemporalConv(nn.Module):
    """Multi-branch temporal block (models/ctrgcn.py:72-147): len(dilations) x [1x1 -> BN -> ReLU -> (k x 1)
    dilated conv -> BN], a [1x1 -> BN -> ReLU -> MaxPool(3x1) -> BN] branch and a strided [1x1 -> BN] branch,
    concatenated over channels, plus an optional residual."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilations=[1, 2, 3, 4], residual=True,
                 residual_kernel_size=1):
        super().__init__()
        assert out_channels % (len(dilations) + 2) == 0, '# out channels should be multiples of # branches'
        self.num_branches = len(dilations) + 2
        self.num_dil = len(dilations)
        self.stride = stride
        bc = out_channels // self.num_branches
        self.branch_channels = bc
        if type(kernel_size) == list:
            assert len(kernel_size) == len(dilations)
        else:
            kernel_size = [kernel_size] * len(dilations)
        branches = []
        for ks, dil in zip(kernel_size, dilations):
            branches.append(nn.Sequential(nn.Conv2d(in_channels, bc, kernel_size=1, padding=0), nn.BatchNorm2d(bc),
                                          nn.ReLU(inplace=True),
                                          TemporalConv(bc, bc, kernel_size=ks, stride=stride, dilation=dil)))
        branches.append(nn.Sequential(nn.Conv2d(in_channels, bc, kernel_size=1, padding=0), nn.BatchNorm2d(bc),
                                      nn.ReLU(inplace=True),
                                      nn.MaxPool2d(kernel_size=(3, 1), stride=(stride, 1), padding=(1, 0)),
                                      nn.BatchNorm2d(bc)))
        branches.append(nn.Sequential(nn.Conv2d(in_channels, bc, kernel_size=1, padding=0, stride=(stride, 1)),
                                      nn.BatchNorm2d(bc)))
        self.branches = nn.ModuleList(branches)
        if not residual:
            self.res_kind = 'none'
            self.residual = lambda x: 0
        elif in_channels == out_channels and stride == 1:
            self.res_kind = 'identity'
            self.residual = lambda x: x
        else:
            self.res_kind = 'conv'
            self.residual = TemporalConv(in_channels, out_channels, kernel_size=residual_kernel_size, stride=stride)
        self.apply(_ms_tcn_reinit)

    def _run(self, x, r_in, res_kind, res_mod, relu):
        params = Fn.ms_tcn_params(self)
        if res_kind == 'conv':
            params = params + Fn.res_conv_params(res_mod)
        return Fn.MsTcnFn.apply(x, r_in, self, res_kind, res_mod, relu, *params)

    def forward(self, x):
        return self._run(x, None, self.res_kind, self.residual if self.res_kind == 'conv' else None, False)


class CTRGC(nn.Module):
    """Channel-wise topology refinement graph convolution (models/ctrgcn.py:150-177)."""

    def __init__(self, in_channels, out_channels, rel_reduction=8, mid_reduction=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        if in_channels == 3 or in_channels == 9:
            self.rel_channels = 8
            self.mid_channels = 16
        else:
            self.rel_channels = in_channels // rel_reduction
            self.mid_channels = in_channels // mid_reduction
        self.conv1 = nn.Conv2d(in_channels, self.rel_channels, kernel_size=1)
        self.conv2 = nn.Conv2d(in_channels, self.rel_channels, kernel_size=1)
        self.conv3 = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.conv4 = nn.Conv2d(self.rel_channels, out_channels, kernel_size=1)
        self.tanh = nn.Tanh()
        for c in (self.conv1, self.conv2, self.conv3, self.conv4):
            _kaiming_conv(c)

    def forward(self, x, A=None, alpha=1):
        V = x.shape[-1]
        if A is None:
            A = torch.zeros(V, V, device=x.device, dtype=torch.float32)
        if not torch.is_tensor(alpha):
            alpha = torch.full((1,), float(alpha), device=x.device, dtype=torch.float32)
        return Fn.CtrgcFn.apply(x, A, alpha, self, *Fn.ctrgc_params(self))


class unit_tcn(nn.Module):
    """(k x 1) conv + BatchNorm; the ReLU it constructs is never applied (models/ctrgcn.py:179-193)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, stride=1):
        super().__init__()
        pad = int((kernel_size - 1) / 2)
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=(kernel_size, 1), padding=(pad, 0),
                              stride=(stride, 1))
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        _kaiming_conv(self.conv)
        _const_bn(self.bn, 1)

    def forward(self, x):
        return Fn.ConvBnFn.apply(x, self.conv, self.bn, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias)


class unit_gcn(nn.Module):
    """Three CTRGC subsets + BN + the reference's offset branch + residual + ReLU (models/ctrgcn.py:196-263)."""

    def __init__(self, in_channels, out_channels, A, coff_embedding=4, adaptive=True, residual=True):
        super().__init__()
        self.inter_c = out_channels // coff_embedding
        self.out_c = out_channels
        self.in_c = in_channels
        self.adaptive = adaptive
        self.num_subset = A.shape[0]
        self.convs = nn.ModuleList([CTRGC(in_channels, out_channels) for _ in range(self.num_subset)])
        self.has_down = bool(residual and in_channels != out_channels)
        self.residual_identity = bool(residual and in_channels == out_channels)
        if self.has_down:
            self.down = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        elif residual:
            self.down = lambda x: x
        else:
            self.down = lambda x: 0
        self.offset_conv = nn.Sequential(nn.Conv2d(out_channels, out_channels, 1), nn.BatchNorm2d(out_channels),
                                         nn.Tanh())
        A32 = torch.from_numpy(np.asarray(A).astype(np.float32))
        if adaptive:
            self.PA = nn.Parameter(A32)
        else:
            self.register_buffer('A', A32, persistent=False)    # the reference keeps a plain tensor (no state_dict key)
        self.alpha = nn.Parameter(torch.zeros(1))
        self.bn = nn.BatchNorm2d(out_channels)
        self.soft = nn.Softmax(-2)
        self.relu = nn.ReLU(inplace=True)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                _kaiming_conv(m)
            elif isinstance(m, nn.BatchNorm2d):
                _const_bn(m, 1)
        _const_bn(self.bn, 1e-6)
        nn.init.zeros_(self.offset_conv[0].weight)
        nn.init.zeros_(self.offset_conv[0].bias)

    def forward(self, x):
        return Fn.UnitGcnFn.apply(x, self, *Fn.unit_gcn_params(self))


class TCN_GCN_unit(nn.Module):
    """relu( tcn1(gcn1(x)) + residual(x) )  (models/ctrgcn.py:266-284); the residual add and the ReLU are fused
    into the epilogue of tcn1."""

    def __init__(self, in_channels, out_channels, A, stride=1, residual=True, adaptive=True, kernel_size=5,
                 dilations=[1, 2]):
        super().__init__()
        self.gcn1 = unit_gcn(in_channels, out_channels, A, adaptive=adaptive)
        self.tcn1 = MultiScale_TemporalConv(out_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                            dilations=dilations, residual=False)
        self.relu = nn.ReLU(inplace=True)
        if not residual:
            self.res_kind = 'none'
            self.residual = lambda x: 0
        elif in_channels == out_channels and stride == 1:
            self.res_kind = 'identity'
            self.residual = lambda x: x
        else:
            self.res_kind = 'conv'
            self.residual = unit_tcn(in_channels, out_channels, kernel_size=1, stride=stride)

    def forward(self, x):
        y = self.gcn1(x)
        if self.tcn1.res_kind != 'none':     # not constructible through this class; keep exact semantics anyway
            y = self.tcn1(y)
            r = self.residual(x)
            return torch.relu(y + r)
        # the residual's cotangent w.r.t. x is handed to gcn1's backward (same input x) instead of an autograd add pass
        self.gcn1.__dict__.pop('_tamgcn_res_cot', None)            # nothing may be left over from an aborted backward
        self.tcn1.__dict__['_tamgcn_res_sink'] = self.gcn1 if (self.res_kind != 'none' and x.requires_grad) else None
        return self.tcn1._run(y, x if self.res_kind != 'none' else None, self.res_kind,
                              self.residual if self.res_kind == 'conv' else None, True)


class Model(nn.Module):
    """CTR-GCN network (models/ctrgcn.py:287-374): data_bn, ten TCN_GCN_units (64,64,64,64,128,128,128,256,256,256;
    stride 2 at l5 and l8), global average pooling over (T,V) and persons, linear classifier."""

    def __init__(self, num_class=60, num_point=25, num_person=2, graph=None, graph_args=dict(), in_channels=3,
                 drop_out=0, adaptive=True):
        super().__init__()
        if graph is None:
            raise ValueError()
        Graph = import_class(graph) if isinstance(graph, str) else graph
        self.graph = Graph(**graph_args)
        A = self.graph.A
        self.num_class = num_class
        self.num_point = num_point
        self.data_bn = nn.BatchNorm1d(num_person * in_channels * num_point)
        c = 64
        spec = [(in_channels, c, 1, False), (c, c, 1, True), (c, c, 1, True), (c, c, 1, True), (c, 2 * c, 2, True),
                (2 * c, 2 * c, 1, True), (2 * c, 2 * c, 1, True), (2 * c, 4 * c, 2, True), (4 * c, 4 * c, 1, True),
                (4 * c, 4 * c, 1, True)]
        for i, (ci, co, s, res) in enumerate(spec):
            setattr(self, 'l%d' % (i + 1), TCN_GCN_unit(ci, co, A, stride=s, residual=res, adaptive=adaptive))
        self.fc = nn.Linear(4 * c, num_class)
        nn.init.normal_(self.fc.weight, 0, math.sqrt(2. / num_class))
        _const_bn(self.data_bn, 1)
        self.drop_out = nn.Dropout(drop_out) if drop_out else (lambda x: x)
        self.act_dtype = None      # None -> follow tam_gcn_b200.get_act_dtype()

    def _trunk(self, x):
        from . import get_act_dtype
        # permute + data_bn + permute + cast: one kernel (csrc/head.cu), the input is read in place through its strides
        M = x.shape[4] if x.dim() == 5 else 1
        x = Fn.DataBnFn.apply(x, self.data_bn, self.num_point, False, self.act_dtype or get_act_dtype(),
                              self.data_bn.weight, self.data_bn.bias)
        N = x.shape[0] // M
        for i in range(1, 11):
            x = getattr(self, 'l%d' % i)(x)
        return x, N, M

    def forward(self, x):
        x, N, M = self._trunk(x)
        if isinstance(self.drop_out, nn.Dropout) and self.training and self.drop_out.p > 0:
            # dropout between pooling and classifier (models/ctrgcn.py:346): pool with the kernel, ATen dropout + linear
            return self.fc(self.drop_out(Fn.PoolFcFn.apply(x, M, None, None)))
        return Fn.PoolFcFn.apply(x, M, self.fc.weight, self.fc.bias)

    def extract_feature(self, x):
        x, N, M = self._trunk(x)
        NM, C, T, V = x.size()
        x = x.float().view(N, M, C, T, V).permute(0, 2, 3, 4, 1).contiguous()
        return x, x
