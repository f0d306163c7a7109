"""Training / inference step engine for the B200-native models.

`Trainer` is the call a user makes per batch (the reference's hot loop is processor/recognition_rgb.py:48-66:
`output = model(data); loss = CE(output, label); zero_grad(); backward(); step()`).  It runs that step as ONE
CUDA graph made of this library's kernels only: data_bn prologue, the ten CTR-GCN blocks forward, pooled classifier
and cross-entropy, the hand-written backward, (world_size > 1) the NCCL gradient all-reduce, and one fused SGD kernel.

Host-side runtime pieces (all per Trainer, nothing global):

  * `params.ParamStore`: every parameter lives in one flat fp32 buffer P, every gradient in a mirror buffer G that the
    backward kernels accumulate into directly; momentum is a third flat buffer.  The optimiser is one kernel over
    (P, G, momentum) with the learning rate in device memory, so `set_lr` (the reference's per-epoch
    `adjust_learning_rate`, processor/recognition_rgb.py:43-46) reaches a captured graph.
  * `arena.ZeroArena`: the ~270 small zero-initialised accumulators of a step (fp64 BatchNorm sums ...) are slices
    of one buffer cleared by one memset; each Trainer owns its arena, and a buffer a graph has captured is never freed.
  * weight-gradient kernels are off the critical path of backward (nothing downstream reads them before the
    optimiser), so they are launched on a side stream and joined before the all-reduce / optimiser.
  * data parallelism (one process per GPU, batch sharded, per-rank BatchNorm statistics — the reference's
    nn.DataParallel semantics, processor/io.py:85-87): G is all-reduced in place, in two slices — the upper layers'
    slice is launched from inside backward as soon as their gradients are final and overlaps the lower layers'
    backward; the 1/world_size factor is folded into the optimiser kernel.
"""
import os

import torch
import torch.distributed as dist

from . import arena as _arena
from . import functional as Fn
from . import ops
from .params import ParamStore


class Trainer:
    def __init__(self, model, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, use_graph=True,
                 process_group=None, side_stream=True, overlap_allreduce=True, loss_fn=None, wgrad_sm_share=None):
        self.model = model
        self.store = ParamStore(model)
        self.device = self.store.device
        self.cuda = self.device.type == 'cuda'
        self.mom = torch.zeros_like(self.store.P)
        self.lr = torch.full((1,), float(lr), device=self.device, dtype=torch.float32)
        self._one = torch.ones((), device=self.device, dtype=torch.float32)     # seed of backward (no fill kernel per step)
        self.momentum, self.nesterov, self.weight_decay = float(momentum), bool(nesterov), float(weight_decay)
        self.loss_fn = loss_fn or Fn.cross_entropy
        self.use_graph = use_graph and self.cuda
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.pg = process_group
        self.arena = _arena.ZeroArena(self.device)
        self.packs = ops.PackCache(model) if self.cuda else None
        self._arena_keep = []                       # buffers baked into captured graphs: never freed
        # stream priorities (TAMGCN_STREAM_PRIO=0 turns them off): the step graph is captured on a high-priority stream (the data-gradient
        # chain), the weight gradients run at low priority, so a freed SM goes to the critical path first
        self.prio = os.environ.get('TAMGCN_STREAM_PRIO', '1') == '1'
        hi = -1 if self.prio else 0
        self.side = torch.cuda.Stream(device=self.device, priority=0) if (self.cuda and side_stream) else None
        self.wgrad_sm_share = int(wgrad_sm_share or os.environ.get('TAMGCN_WGRAD_SM_SHARE', 75))
        self.bwd_main_sm_share = int(os.environ.get('TAMGCN_BWD_MAIN_SM_SHARE', 100)) if self.side is not None else 100
        nbr = int(os.environ.get('TAMGCN_BRANCH_STREAMS', 2))
        self.branch_streams = [torch.cuda.Stream(device=self.device, priority=hi) for _ in range(nbr)] if (self.cuda and side_stream and nbr > 0) else None
        self.capture_stream = torch.cuda.Stream(device=self.device, priority=hi) if (self.cuda and self.prio) else None
        self.overlap = bool(overlap_allreduce) and self.world > 1 and self.cuda
        self.graph = None
        self.static_x = self.static_y = self.static_loss = None
        self.captured_launches = 0
        self.store.attach_grads()
        self._ranges = self.store.trainable_ranges()
        # all-reduce split: parameters of the upper half of the layers (by offset in G) are reduced from inside backward
        self._split = None
        self._hook_handle = None
        self._pending = None
        if self.overlap:
            self._install_overlap_hook()
        if self.world > 1:
            self.broadcast_parameters()

    # ---- learning rate --------------------------------------------------------------------------------------
    def set_lr(self, lr):
        """Change the learning rate (also of an already captured graph: the kernel reads it from device memory)."""
        self.lr.fill_(float(lr))

    def get_lr(self):
        return float(self.lr)

    # ---- data parallel plumbing ------------------------------------------------------------------------------
    def broadcast_parameters(self):
        """One-time broadcast of rank 0's parameters (one flat buffer) and buffers."""
        with torch.no_grad():
            dist.broadcast(self.store.P, 0, group=self.pg)
            for t in self.model.buffers():
                dist.broadcast(t, 0, group=self.pg)

    def _install_overlap_hook(self):
        """Pick the module in the middle of the network (by parameter offset) and hook the gradient of its input:
        when that gradient exists, every layer above has launched all of its backward kernels."""
        kids = [m for m in self.model.children() if any(p.requires_grad for p in m.parameters())]
        if len(kids) < 4:
            return
        mid = kids[len(kids) // 2]
        first = min(self.store.offset_of(p) for p in mid.parameters())
        if not any(lo <= first < hi for lo, hi in self._ranges):
            return
        self._split = first

        def pre_hook(mod, inputs):
            x = inputs[0]
            if torch.is_tensor(x) and x.requires_grad and self._armed:
                x.register_hook(self._upper_grads_ready)

        self._armed = False
        self._hook_handle = mid.register_forward_pre_hook(pre_hook)

    def _upper_grads_ready(self, grad):
        if self._pending is None and self._armed:
            hi = self._ranges[-1][1]
            self._pending = self._allreduce_async(self._split, hi)
        return None

    def _allreduce_async(self, lo, hi):
        """Launch the all-reduce of G[lo:hi] so that it waits for everything launched so far on the main AND the side
        stream, without making the main stream wait for it."""
        main = torch.cuda.current_stream()
        if self.side is not None:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                return dist.all_reduce(self.store.G[lo:hi], group=self.pg, async_op=True)
        return dist.all_reduce(self.store.G[lo:hi], group=self.pg, async_op=True)

    def _allreduce_grads(self):
        """Sum gradients over ranks in place in G (the mean's 1/world is applied by the optimiser kernel)."""
        G = self.store.G
        lo0 = self._ranges[0][0]
        hi0 = self._ranges[-1][1]
        if self._pending is not None:
            dist.all_reduce(G[lo0:self._split], group=self.pg)
            self._pending.wait()
            self._pending = None
        else:
            dist.all_reduce(G[lo0:hi0], group=self.pg)

    # ---- one optimisation step -----------------------------------------------------------------------------
    def _step_body(self, x, y):
        st = self.store
        st.G.zero_()                                  # ONE memset for all parameter gradients
        self.arena.begin_step()                       # ONE memset for every other zero-initialised accumulator
        self._pending = None
        self._armed = self.overlap and self._split is not None
        try:
            with _arena.use(self.arena), st.direct_grads(), ops.side_stream(self.side, self.wgrad_sm_share), \
                    ops.branches(self.branch_streams), Fn.pack_cache(self.packs):
                out = self.model(x)
                loss = self.loss_fn(out, y)
                with ops.main_sm_share(self.bwd_main_sm_share):
                    loss.backward(self._one if loss.dtype == torch.float32 and loss.dim() == 0 else None)
            ops.join_side_stream(self.side)
            if self.world > 1:
                self._allreduce_grads()
            for lo, hi in self._ranges:
                ops.sgd_step(st.P[lo:hi], st.G[lo:hi], self.mom[lo:hi], self.lr, self.momentum, self.weight_decay,
                             self.nesterov, 1.0 / self.world)
        finally:
            self._armed = False
            self.arena.end_step()
        return loss.detach()

    def _snapshot(self):
        bufs = [b for b in self.model.buffers()]
        return self.store.P.clone(), self.mom.clone(), [b.clone() for b in bufs], bufs

    def _restore(self, snap):
        P, mom, vals, bufs = snap
        with torch.no_grad():
            self.store.P.copy_(P)
            self.mom.copy_(mom)
            for b, v in zip(bufs, vals):
                b.copy_(v)

    def _capture(self, x, y):
        from . import _C
        if not self.store.valid():
            raise RuntimeError('Trainer: the model parameters were moved (.to() / .data = ...) after the Trainer was '
                               'built; build the Trainer after placing the model on its device')
        self.static_x = x.clone()
        self.static_y = y.clone()
        # warm-up (allocator, arena sizing, shared-memory opt-ins) must not change the training state: the first
        # captured replay is optimisation step #1, exactly as in the reference loop
        snap = self._snapshot()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._step_body(self.static_x, self.static_y)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        self.arena.freeze()                          # no re-allocation from here on
        self._arena_keep.append(self.arena.buf)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _C.launch_count()
        with torch.cuda.graph(self.graph, stream=self.capture_stream):
            self.static_loss = self._step_body(self.static_x, self.static_y)
        self.captured_launches = _C.launch_count() - n0

    def step(self, x, y):
        """x: (N, C, T, V, M) float32 tensor, y: (N,) int64 tensor -> loss (0-dim tensor, a fresh copy)."""
        if not self.use_graph:
            if not self.store.valid():
                raise RuntimeError('Trainer: model parameters no longer alias the flat parameter buffer')
            return self._step_body(x, y)
        if self.graph is None:
            self._capture(x, y)
        if x.shape != self.static_x.shape or y.shape != self.static_y.shape:
            return self._step_body(x, y)             # e.g. a smaller last batch: run it eagerly
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.static_loss.clone()

    def step_from_host(self, x_host, y_host):
        """End-to-end step from pinned host buffers: H2D copy of the batch, the step, D2H read of the loss."""
        if self.use_graph:
            if self.graph is None:
                self._capture(x_host.to(self.device, non_blocking=True), y_host.to(self.device, non_blocking=True))
            if x_host.shape == self.static_x.shape:
                self.static_x.copy_(x_host, non_blocking=True)
                self.static_y.copy_(y_host, non_blocking=True)
                self.graph.replay()
                return float(self.static_loss)       # D2H + sync
        return float(self._step_body(x_host.to(self.device, non_blocking=True), y_host.to(self.device, non_blocking=True)))


class Predictor:
    """Inference (eval-mode forward) as a CUDA graph, for the large-batch throughput sweep."""

    def __init__(self, model, use_graph=True, eval_mode=True):
        """eval_mode=False keeps the model in train() — the frozen-but-training GCN branch of the cross-modal fusion
        model (models/resnet_gcn_attention.py:24-26): forward only, batch statistics, running stats updated."""
        self.model = model.eval() if eval_mode else model.train()
        self.packs = ops.PackCache(model)
        self.use_graph = use_graph
        self.graph = None
        # independent branches of a block (MS-TCN branches, the strided 1x1 next to the heads) on two extra streams inside
        # the captured forward (TAMGCN_INFER_BRANCHES=0 turns it off)
        dev = next(model.parameters()).device
        self.branch_streams = ([torch.cuda.Stream(device=dev) for _ in range(2)]
                               if (dev.type == 'cuda' and use_graph and os.environ.get('TAMGCN_INFER_BRANCHES', '1') == '1') else None)
        self.static_x = self.static_out = None
        self.captured_launches = 0

    @torch.no_grad()
    def __call__(self, x):
        from . import _C
        if not self.use_graph:
            with Fn.pack_cache(self.packs):
                return self.model(x)
        if self.graph is None or self.static_x.shape != x.shape:
            self.static_x = x.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    with Fn.pack_cache(self.packs), ops.branches(self.branch_streams):
                        self.model(self.static_x)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _C.launch_count()
            with torch.cuda.graph(self.graph), Fn.pack_cache(self.packs), ops.branches(self.branch_streams):
                self.static_out = self.model(self.static_x)           # one batched weight-tile refresh per replay
            self.captured_launches = _C.launch_count() - n0
        self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def from_host(self, x_host):
        """End to end: pinned host batch -> H2D -> forward -> D2H of the logits."""
        if self.graph is None or self.static_x.shape != x_host.shape:
            self(x_host.cuda(non_blocking=True))
        self.static_x.copy_(x_host, non_blocking=True)
        self.graph.replay()
        return self.static_out.cpu()
