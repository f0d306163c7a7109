"""B200-native ST-GCN graph layers: drop-in replacements for the classes of the reference's models/stgcn.py.

Same class names, constructor / forward signatures and state_dict keys as models/stgcn.py:37-252.
`ConvTemporalGraphical` and `st_gcn` run on the fused kernels of `tam_gcn_b200.functional`; `Model`
is the thin wrapper (data_bn, edge-importance weighting, pooling, 1x1 classifier).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .ctrgcn import import_class


class ConvTemporalGraphical(nn.Module):
    """(t_k x 1) conv Cin -> K*Cout followed by einsum('nkctv,kvw->nctw') with the K-partition adjacency
    (models/stgcn.py:37-63)."""

    def __init__(self, in_channels, out_channels, kernel_size, t_kernel_size=1, t_stride=1, t_padding=0,
                 t_dilation=1, bias=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv = nn.Conv2d(in_channels, out_channels * kernel_size, kernel_size=(t_kernel_size, 1),
                              padding=(t_padding, 0), stride=(t_stride, 1), dilation=(t_dilation, 1), bias=bias)

    def forward(self, x, A):
        assert A.size(0) == self.kernel_size
        return Fn.CtgFn.apply(x, A, self, self.conv.weight, self.conv.bias), A


class st_gcn(nn.Module):
    """Graph conv -> BN -> ReLU -> (k_t x 1) conv (stride) -> BN -> Dropout, + residual, ReLU
    (models/stgcn.py:66-99), fused into a handful of kernels."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dropout=0, residual=True):
        super().__init__()
        assert len(kernel_size) == 2
        assert kernel_size[0] % 2 == 1
        padding = ((kernel_size[0] - 1) // 2, 0)
        self.gcn = ConvTemporalGraphical(in_channels, out_channels, kernel_size[1])
        self.tcn = nn.Sequential(nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
                                 nn.Conv2d(out_channels, out_channels, (kernel_size[0], 1), (stride, 1), padding),
                                 nn.BatchNorm2d(out_channels), nn.Dropout(dropout, inplace=True))
        self.dropout_p = dropout
        if not residual:
            self.res_kind = 'none'
            self.residual = lambda x: 0
        elif in_channels == out_channels and stride == 1:
            self.res_kind = 'identity'
            self.residual = lambda x: x
        else:
            self.res_kind = 'conv'
            self.residual = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=(stride, 1)),
                                          nn.BatchNorm2d(out_channels))
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x, A):
        assert A.size(0) == self.gcn.kernel_size
        if self.dropout_p and self.training:
            # Rare path (the reference's own configs use dropout 0 here, models/stgcn.py:81): the graph convolution runs on
            # the native kernels; the TCN tail, whose dropout mask sits between BatchNorm and the residual sum, goes
            # through the same torch modules on the GPU (cuDNN), under autocast when activations are bf16.
            g, _ = self.gcn(x, A)
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=(g.dtype == torch.bfloat16)):
                out = self.relu(self.tcn(g) + self.residual(x))
            return out.to(g.dtype), A
        return Fn.StGcnFn.apply(x, A, self, *Fn.st_gcn_params(self)), A


class Model(nn.Module):
    """ST-GCN with learnable per-layer edge importance (models/stgcn.py:102-252)."""

    def __init__(self, in_channels=3, num_class=4, num_point=20, num_person=1, graph=None, graph_args=dict(),
                 edge_importance_weighting=True, dropout=0, **kwargs):
        super().__init__()
        if graph is None:
            raise ValueError("Graph class must be specified")
        Graph = import_class(graph) if isinstance(graph, str) else graph
        self.graph = Graph(**graph_args)
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)
        ks = (9, A.size(0))
        self.num_point = num_point
        self.data_bn = nn.BatchNorm1d(num_person * in_channels * num_point)
        chans = [(in_channels, 64, 1), (64, 64, 1), (64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 1),
                 (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 256, 1)]
        self.st_gcn_networks = nn.ModuleList(
            [st_gcn(ci, co, ks, s, residual=(i > 0), **kwargs) for i, (ci, co, s) in enumerate(chans)])
        if edge_importance_weighting:
            self.edge_importance = nn.ParameterList([nn.Parameter(torch.ones(self.A.size()))
                                                     for _ in self.st_gcn_networks])
        else:
            self.edge_importance = [1] * len(self.st_gcn_networks)
        self.fcn = nn.Conv2d(256, num_class, kernel_size=1)
        self.drop_out = nn.Dropout(dropout) if dropout else (lambda x: x)
        self.act_dtype = None

    def _trunk(self, x):
        from . import get_act_dtype
        # data_bn over (v, c) channels with the persons folded into the batch (models/stgcn.py:174-181): one kernel
        M = x.shape[4] if x.dim() == 5 else 1
        x = Fn.DataBnFn.apply(x, self.data_bn, self.num_point, True, self.act_dtype or get_act_dtype(),
                              self.data_bn.weight, self.data_bn.bias)
        N = x.shape[0] // M
        for gcn, importance in zip(self.st_gcn_networks, self.edge_importance):
            x, _ = gcn(x, self.A * importance)
        return x, N, M

    def forward(self, x):
        x, N, M = self._trunk(x)
        if isinstance(self.drop_out, nn.Dropout) and self.training and self.drop_out.p > 0:
            x = self.drop_out(Fn.PoolFcFn.apply(x, M, None, None))
            return F.linear(x, self.fcn.weight.view(self.fcn.out_channels, -1), self.fcn.bias)
        # global pooling + the 1x1 classifier on the pooled (N,256,1,1) feature == pooled linear layer: one kernel
        return Fn.PoolFcFn.apply(x, M, self.fcn.weight, self.fcn.bias)

    def extract_feature(self, x):
        x, N, M = self._trunk(x)
        x = x.float()
        _, c, t, v = x.size()
        feature = x.view(N, M, c, t, v).permute(0, 2, 3, 4, 1)
        out = torch.einsum('oc,nctv->notv', self.fcn.weight.view(self.fcn.out_channels, -1), x)
        if self.fcn.bias is not None:
            out = out + self.fcn.bias.view(1, -1, 1, 1)
        output = out.view(N, M, -1, t, v).permute(0, 2, 3, 4, 1)
        return output, feature

    def get_edge_importance_per_joint(self):
        """Per-joint score: sum of |incoming| + |outgoing| edge-importance mass over all layers and partitions,
        normalised by its maximum (models/stgcn.py:227-252)."""
        V = self.A.size(1)
        score = np.zeros(V)
        for imp in self.edge_importance:
            w = imp.detach().cpu().numpy()
            score += w.sum(axis=(0, 1)) + w.sum(axis=(0, 2))
        return score / score.max()
