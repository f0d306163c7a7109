"""Stand-in for the reference's `models/ctrgcn.py` module on machines without the reference tree (the GPU box) —
TEST INFRASTRUCTURE.  It reproduces the one property `tam_gcn_b200.patch_reference` relies on: the network class
looks its layer classes up as MODULE GLOBALS at construction time (models/ctrgcn.py:305-314), and wraps them with plain
ATen ops for the prologue and the head (models/ctrgcn.py:324-348).  The layer names below are placeholders that
`patch_reference(ctrgcn_module=this_module)` rebinds; nothing here is a layer implementation."""
import torch.nn as nn

TemporalConv = MultiScale_TemporalConv = CTRGC = unit_tcn = unit_gcn = TCN_GCN_unit = None


class Model(nn.Module):
    def __init__(self, A, num_class=10, num_point=20, num_person=1, in_channels=3):
        super().__init__()
        self.num_point = num_point
        self.data_bn = nn.BatchNorm1d(num_person * in_channels * num_point)
        plan = [(in_channels, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 64, 1, True), (64, 128, 2, True),
                (128, 128, 1, True), (128, 128, 1, True), (128, 256, 2, True), (256, 256, 1, True), (256, 256, 1, True)]
        for i, (ci, co, s, res) in enumerate(plan, 1):
            setattr(self, 'l%d' % i, TCN_GCN_unit(ci, co, A, stride=s, residual=res))      # module-global lookup
        self.fc = nn.Linear(256, num_class)

    def forward(self, x):
        N, C, T, V, M = x.shape
        h = self.data_bn(x.permute(0, 4, 3, 1, 2).reshape(N, M * V * C, T))
        h = h.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).reshape(N * M, C, T, V)
        for i in range(1, 11):
            h = getattr(self, 'l%d' % i)(h)
        return self.fc(h.view(N, M, h.shape[1], -1).mean(3).mean(1))
