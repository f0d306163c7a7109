"""Training / inference step engine for the B200-native models.

`Trainer` is the call a user makes per batch (the reference's hot loop is processor/recognition_rgb.py:48-66:
`output = model(data); loss = CE(output, label); zero_grad(); backward(); step()`).  It runs that step as ONE
CUDA graph: forward, hand-written backward, (for world_size > 1) the flat-bucket NCCL gradient all-reduce and
the fused SGD update are captured once and replayed, so the ~600 kernel launches of a step cost no host time.

One process per GPU.  Data parallelism shards the batch across ranks; BatchNorm statistics stay per rank
(the reference's nn.DataParallel semantics, processor/io.py:85-87); gradients are averaged over ranks.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F


class Trainer:
    def __init__(self, model, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, use_graph=True,
                 process_group=None, fused=True):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        # the reference's optimiser (processor/recognition_rgb.py:21-28): SGD, momentum 0.9, nesterov, weight decay
        self.opt = torch.optim.SGD(self.params, lr=lr, momentum=momentum, nesterov=nesterov,
                                   weight_decay=weight_decay, fused=fused)
        self.use_graph = use_graph
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.pg = process_group
        self.graph = None
        self.static_x = self.static_y = self.static_loss = None
        self.flat = None
        self.captured_launches = 0
        self._warm = 0
        if self.world > 1:
            self.broadcast_parameters()

    # ---- data parallel plumbing ------------------------------------------------------------------
    def broadcast_parameters(self):
        """One-time broadcast of rank 0's parameters and buffers."""
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t, 0, group=self.pg)

    def _allreduce_grads(self):
        """Average gradients over ranks through ONE flat fp32 bucket (6.8 MB for CTR-GCN/NW-UCLA)."""
        grads = [p.grad for p in self.params]
        n = sum(g.numel() for g in grads)
        if self.flat is None or self.flat.numel() != n:
            self.flat = torch.empty(n, device=grads[0].device, dtype=torch.float32)
        torch.cat([g.reshape(-1) for g in grads], out=self.flat)
        self.flat.mul_(1.0 / self.world)
        dist.all_reduce(self.flat, group=self.pg)
        off = 0
        for p in self.params:
            k = p.numel()
            p.grad = self.flat[off:off + k].view_as(p)
            off += k

    # ---- one optimisation step -------------------------------------------------------------------
    def _step_body(self, x, y):
        from . import arena
        self.opt.zero_grad(set_to_none=True)
        ar = arena.arena(x.device)
        ar.begin_step()                              # every zero-initialised accumulator of the step: one memset
        try:
            out = self.model(x)
            loss = F.cross_entropy(out, y)
            loss.backward()
            if self.world > 1:
                self._allreduce_grads()
            self.opt.step()
        finally:
            ar.end_step()
        return loss.detach()

    def _capture(self, x, y):
        from . import _C
        self.static_x = x.clone()
        self.static_y = y.clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):                       # allocator / optimiser-state / smem-attribute warm-up
                self._step_body(self.static_x, self.static_y)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _C.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._step_body(self.static_x, self.static_y)
        self.captured_launches = _C.launch_count() - n0

    def step(self, x, y):
        """x: (N, C, T, V, M) float32 CUDA tensor, y: (N,) int64 CUDA tensor -> loss (0-dim CUDA tensor)."""
        if not self.use_graph:
            return self._step_body(x, y)
        if self.graph is None:
            self._capture(x, y)
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.static_loss

    def step_from_host(self, x_host, y_host):
        """End-to-end step from pinned host buffers: H2D copy of the batch, the step, D2H read of the loss."""
        if self.graph is None and self.use_graph:
            self._capture(x_host.cuda(non_blocking=True), y_host.cuda(non_blocking=True))
        if self.use_graph:
            self.static_x.copy_(x_host, non_blocking=True)
            self.static_y.copy_(y_host, non_blocking=True)
            self.graph.replay()
            return float(self.static_loss)           # D2H + sync
        return float(self._step_body(x_host.cuda(non_blocking=True), y_host.cuda(non_blocking=True)))


class Predictor:
    """Inference (eval-mode forward) as a CUDA graph, for the large-batch throughput sweep."""

    def __init__(self, model, use_graph=True):
        self.model = model.eval()
        self.use_graph = use_graph
        self.graph = None
        self.static_x = self.static_out = None
        self.captured_launches = 0

    @torch.no_grad()
    def __call__(self, x):
        from . import _C
        if not self.use_graph:
            return self.model(x)
        if self.graph is None or self.static_x.shape != x.shape:
            self.static_x = x.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.model(self.static_x)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _C.launch_count()
            with torch.cuda.graph(self.graph):
                self.static_out = self.model(self.static_x)
            self.captured_launches = _C.launch_count() - n0
        self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
