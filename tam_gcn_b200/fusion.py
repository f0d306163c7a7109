"""Cross-modal head of the reference's fusion model (models/resnet_gcn_attention.py:6-122) on the B200-native kernels.

    f_gcn, _ = gcn.extract_feature(x_gcn)                 # (N, 256, T/4, V, M)          :82
    f_gcn = f_gcn.mean(dim=(2, 3, 4))                     # (N, 256)                     :85
    att = Sigmoid(Linear(ReLU(BatchNorm1d(Linear(f_gcn)))))          # (N, 2048)          :59-65, :89
    f_rgb = resnet.layer4(...)                            # (N, 2048, 7, 7)              :96-105
    out = classifier(global_pool(f_rgb * att[..., None, None]))                          :111-120

What runs where:
  * the CTR-GCN trunk: `tam_gcn_b200.ctrgcn.Model` (hand-written kernels).  The permuted fp32 feature tensor that
    `extract_feature` returns is never materialised: the mean over (T, V, M) is taken straight from the trunk output
    by the pooling kernel (csrc/head.cu).
  * the attention MLP: the two Linear layers are (1 x 1) convolutions over a "(1, C, 1, N)" tensor — the batch is the
    position axis — so they run on the library's convolution kernels with everything those already fuse: BatchNorm1d
    statistics in the producer's epilogue, BatchNorm-apply + ReLU as the consumer's lazy operand, the hand-written
    BatchNorm / ReLU backward.  A transpose kernel with a fused sigmoid is the only glue.  fp32 throughout (a 64 x 2048
    problem: latency, not bandwidth).
  * the gate, the global average pool and the classifier are ONE kernel: mean_hw(f_rgb * att) == att * mean_hw(f_rgb).
  * the ResNet-50 backbone is not part of the skeleton-GCN hot path (out of scope); any module with the
    torchvision ResNet attribute layout can be plugged in and runs in stock PyTorch.

Parameter names follow the reference (`gcn.*`, `attention_transform.{0,1,3}.*`, `classifier.*`, `resnet.*`), so a
reference checkpoint loads with `load_state_dict` (processor/recognition_cross_modal.py:101-113 loads into `model.gcn`).
"""
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .ctrgcn import Model as CTRGCN
from .functional import _BnBwd, _BnCoef, _GradOut, _bn_backward, _bn_forward, _bwd_opnd, _full, _require_cuda, _zeros
from .ops import Opnd


class AttentionMlpFn(torch.autograd.Function):
    """att = sigmoid(W2 relu(BN(W1 f + b1)) + b2);  f: (N, Cin) fp32 -> att: (N, Cout) fp32."""

    @staticmethod
    def forward(ctx, f, lin1, bn, lin2, *params):
        _require_cuda(f)
        f = f.contiguous().float()
        N, Cin = f.shape
        Ch, Co = lin1.weight.shape[0], lin2.weight.shape[0]
        if lin1.weight.shape[1] != Cin or lin2.weight.shape[1] != Ch or lin1.weight.dtype != torch.float32:
            raise ValueError('attention MLP shape / dtype mismatch')
        train = bn.training
        dev = f.device
        fT = torch.empty((1, Cin, 1, N), device=dev, dtype=torch.float32)
        ops.transpose_act(f, fT.view(Cin, N), 0)
        h = torch.empty((1, Ch, 1, N), device=dev, dtype=torch.float32)
        st = _zeros((2, Ch), f, torch.float64) if train else None
        b1 = lin1.bias if lin1.bias is not None else torch.zeros(Ch, device=dev)
        b2 = lin2.bias if lin2.bias is not None else torch.zeros(Co, device=dev)
        ops.conv_fwd(fT, lin1.weight, b1, h, stats=st)
        cf = _BnCoef(Ch, f)
        _bn_forward([bn], [_full(Ch)], cf, st, N, train)
        hop = Opnd(h, a=cf.scale, c=cf.shift, relu=True)
        z = torch.empty((1, Co, 1, N), device=dev, dtype=torch.float32)
        ops.conv_fwd(hop, lin2.weight, b2, z)
        att = torch.empty((N, Co), device=dev, dtype=torch.float32)
        ops.transpose_act(z.view(Co, N), att, 1)
        ctx.mods, ctx.train, ctx.cf = (lin1, bn, lin2), train, cf
        ctx.save_for_backward(fT, h, att)
        return att

    @staticmethod
    def backward(ctx, g):
        lin1, bn, lin2 = ctx.mods
        fT, h, att = ctx.saved_tensors
        cf, train = ctx.cf, ctx.train
        N, Co = att.shape
        Ch, Cin = lin1.weight.shape
        go = _GradOut(att)
        dz = torch.empty((1, Co, 1, N), device=att.device, dtype=torch.float32)
        ops.transpose_act(g.contiguous().float(), dz.view(Co, N), 2, aux=att)
        hop = Opnd(h, a=cf.scale, c=cf.shift, relu=True)
        dW2 = go.buf(lin2.weight)
        db2 = go.buf(lin2.bias) if lin2.bias is not None else _zeros((Co,), att, torch.float32)
        ops.conv_wgrad(dz, hop, dW2, db2)
        DH = torch.empty_like(h)
        sh = _zeros((2, Ch), att, torch.float64)
        ops.conv_dgrad(dz, lin2.weight, DH, mask=Opnd(h, a=cf.scale, c=cf.shift), stats=(sh[0], sh[1]))
        bw = _BnBwd(Ch, att)
        _bn_backward([bn], [_full(Ch)], cf, bw, sh[0], sh[1], N, train, go)
        dh = _bwd_opnd(DH, h, bw, train)
        dW1 = go.buf(lin1.weight)
        db1 = go.buf(lin1.bias) if lin1.bias is not None else _zeros((Ch,), att, torch.float32)
        ops.conv_wgrad(dh, fT, dW1, db1)
        df = None
        if ctx.needs_input_grad[0]:
            dfT = torch.empty_like(fT)
            ops.conv_dgrad(dh, lin1.weight, dfT)
            df = torch.empty((N, Cin), device=att.device, dtype=torch.float32)
            ops.transpose_act(dfT.view(Cin, N), df, 0)
        return (df, None, None, None, go.ret(dW1, lin1.weight), go.ret(db1, lin1.bias), go.ret(bw.dgamma, bn.weight),
                go.ret(bw.dbeta, bn.bias), go.ret(dW2, lin2.weight), go.ret(db2, lin2.bias))


def attention_mlp(f, seq):
    """`seq`: the reference's attention_transform Sequential (Linear, BatchNorm1d, ReLU, Linear, Sigmoid)."""
    lin1, bn, lin2 = seq[0], seq[1], seq[3]
    return AttentionMlpFn.apply(f, lin1, bn, lin2, lin1.weight, lin1.bias, bn.weight, bn.bias, lin2.weight, lin2.bias)


class GatedPoolFcFn(torch.autograd.Function):
    """logits = classifier(mean_hw(x * gate[..., None, None]));  x: (N, C, H, W) fp32 / bf16, gate: (N, C) fp32."""

    @staticmethod
    def forward(ctx, x, gate, weight, bias):
        _require_cuda(x)
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        gate = gate.contiguous().float()
        N, C, H, W_ = x.shape
        pooled = torch.empty((N, C), device=x.device, dtype=torch.float32)
        logits = torch.empty((N, weight.shape[0]), device=x.device, dtype=torch.float32)
        ops.pool_fc_fwd(x, 1, weight, bias, pooled, logits, gate=gate)
        ctx.xmeta = (x.shape, x.dtype)
        ctx.save_for_backward(pooled, gate, weight, bias)
        return logits

    @staticmethod
    def backward(ctx, dl):
        pooled, gate, weight, bias = ctx.saved_tensors
        shape, dtype = ctx.xmeta
        go = _GradOut(pooled)
        dl = dl.contiguous().float()
        g = torch.empty(shape, device=pooled.device, dtype=dtype) if ctx.needs_input_grad[0] else None
        dgate = torch.empty_like(gate) if ctx.needs_input_grad[1] else None
        dW = go.buf(weight)
        db = go.buf(bias) if bias is not None else None
        ops.pool_fc_bwd(dl, pooled, weight, 1, g, dW, db, gate=gate, dgate=dgate)
        return g, dgate, go.ret(dW, weight), go.ret(db, bias)


def _attention_transform(cin, cout):
    # models/resnet_gcn_attention.py:59-65
    return nn.Sequential(nn.Linear(cin, cout // 2), nn.BatchNorm1d(cout // 2), nn.ReLU(inplace=True),
                         nn.Linear(cout // 2, cout), nn.Sigmoid())


class GcnAttentionBranch(nn.Module):
    """The GCN branch of the fusion model (BASELINE.json configs[4]): skeleton -> CTR-GCN trunk -> mean over (T, V, M)
    -> attention MLP -> (N, 2048) channel gate.  models/resnet_gcn_attention.py:9-28,59-65,82-89."""

    def __init__(self, num_class=10, num_point=20, num_person=1, graph=None, graph_args=dict(), in_channels_gcn=3,
                 drop_out=0, adaptive=True, freeze_gcn=True, resnet_feature_dim=2048):
        super().__init__()
        if graph is None:
            raise ValueError()
        self.gcn = CTRGCN(num_class=num_class, num_point=num_point, num_person=num_person, graph=graph,
                          graph_args=graph_args, in_channels=in_channels_gcn, drop_out=drop_out, adaptive=adaptive)
        if freeze_gcn:
            for p in self.gcn.parameters():
                p.requires_grad = False
        self.attention_transform = _attention_transform(256, resnet_feature_dim)

    def gcn_feature(self, x_gcn):
        """(N, 256): extract_feature(...)[0].mean((2, 3, 4)) without materialising the permuted feature tensor."""
        x, N, M = self.gcn._trunk(x_gcn)
        return Fn.PoolFcFn.apply(x, M, None, None)

    def forward(self, x_gcn):
        return attention_mlp(self.gcn_feature(x_gcn), self.attention_transform)


class ResNet_GCN_Attention(GcnAttentionBranch):
    """Drop-in for models/resnet_gcn_attention.ResNet_GCN_Attention (same constructor arguments, forward signature and
    parameter names).  `resnet`: a backbone with the torchvision ResNet layout (conv1, bn1, relu, maxpool, layer1..4);
    None builds the reference's own `models.resnet.resnet50` when that tree is importable, torchvision's otherwise.
    `pretrained` defaults to False (there is no network here; the reference asks for ImageNet weights, :32)."""

    def __init__(self, num_class=10, num_point=20, num_person=1, graph=None, graph_args=dict(), in_channels_gcn=3,
                 in_channels_rgb=15, drop_out=0, adaptive=True, freeze_gcn=True, resnet=None, pretrained=False):
        super().__init__(num_class, num_point, num_person, graph, graph_args, in_channels_gcn, drop_out, adaptive,
                         freeze_gcn, 2048)
        if resnet is None:
            try:
                from models.resnet import resnet50
            except ImportError:
                from torchvision.models import resnet50 as tv50
                resnet50 = lambda pretrained=False: tv50(weights='IMAGENET1K_V1' if pretrained else None)   # noqa: E731
            resnet = resnet50(pretrained=pretrained)
        self.resnet = resnet
        if in_channels_rgb != 3:                     # inflate conv1 as the reference does (:37-52)
            c1 = self.resnet.conv1
            new = nn.Conv2d(in_channels_rgb, c1.out_channels, kernel_size=c1.kernel_size, stride=c1.stride,
                            padding=c1.padding, bias=False)
            with torch.no_grad():
                new.weight[:] = c1.weight.repeat(1, in_channels_rgb // 3, 1, 1) / (in_channels_rgb // 3)
            self.resnet.conv1 = new
        self.resnet.fc = nn.Identity()
        self.resnet.avgpool = nn.Identity()
        self.global_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.classifier = nn.Linear(2048, num_class)

    def forward(self, x_gcn, x_rgb):
        att = GcnAttentionBranch.forward(self, x_gcn)
        r = self.resnet
        f = r.maxpool(r.relu(r.bn1(r.conv1(x_rgb))))
        f = r.layer4(r.layer3(r.layer2(r.layer1(f))))
        return GatedPoolFcFn.apply(f, att, self.classifier.weight, self.classifier.bias)
