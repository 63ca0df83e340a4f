"""BatchNorm1d of the sparse backbones (the training-mode half).

The reference builds its backbones with ``norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)``
(pcdet/models/backbones_3d/spconv_backbone.py:75, :193) and ``SparseSequential`` applies it to ``.features``
(pcdet/ops/spconv/modules.py).  This subclass keeps the constructor, the parameters / buffers and therefore the
``state_dict`` keys of ``nn.BatchNorm1d``; in training mode on CUDA fp32 row matrices the batch statistics, the
normalisation and the backward run in fv2p_batchnorm_train_fwd / _bwd (csrc/batchnorm.cu: two launches each way, no
host synchronisation).  In eval mode the fused engine folds the running statistics into the conv epilogue
(engine.fold_bn); the op-for-op module graph in eval mode evaluates the same affine map with torch's elementwise
``F.batch_norm`` exactly like the reference.
"""
import torch
from torch import nn

from . import _lib


class _BatchNormTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, factor, ws):
        lib = _lib.load()
        x = x.contiguous()
        n, c = x.shape
        y = torch.empty_like(x)
        save_mean = torch.empty((c,), dtype=torch.float32, device=x.device)
        save_invstd = torch.empty((c,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.fv2p_batchnorm_train_fwd(_lib.ptr(x), n, c, _lib.ptr(weight), _lib.ptr(bias), float(eps),
                                                    float(factor), _lib.ptr(running_mean), _lib.ptr(running_var), 0,
                                                    _lib.ptr(y), _lib.ptr(save_mean), _lib.ptr(save_invstd),
                                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr(x.device)),
                       "batchnorm_train_fwd")
        ctx.save_for_backward(x, weight, save_mean, save_invstd)
        ctx.ws = ws
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(*[t for t in (running_mean, running_var) if t is not None])
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _lib.load()
        x, weight, save_mean, save_invstd = ctx.saved_tensors
        n, c = x.shape
        grad_out = grad_out.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty((c,), dtype=torch.float32, device=x.device)
        db = torch.empty((c,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.fv2p_batchnorm_train_bwd(_lib.ptr(x), _lib.ptr(grad_out), n, c, _lib.ptr(weight),
                                                    _lib.ptr(save_mean), _lib.ptr(save_invstd), _lib.ptr(dx),
                                                    _lib.ptr(dw), _lib.ptr(db), _lib.ptr(ctx.ws), ctx.ws.numel(),
                                                    _lib.stream_ptr(x.device)), "batchnorm_train_bwd")
        return dx, (dw if weight is not None else None), (db if ctx.has_bias else None), None, None, None, None, None


class BatchNorm1d(nn.BatchNorm1d):
    """``nn.BatchNorm1d`` whose training-mode forward / backward on CUDA fp32 ``[N, C]`` inputs are the library's."""

    def _workspace(self, device):
        ws = getattr(self, "_fv2p_ws", None)
        if ws is None or ws.device != device:
            nbytes = _lib.load().fv2p_batchnorm_workspace_bytes(self.num_features)
            ws = torch.zeros(((nbytes + 7) // 8,), dtype=torch.int64, device=device).view(torch.uint8)
            self._fv2p_ws = ws
        return ws

    def forward(self, input):
        native = (self.training or not self.track_running_stats) and input.is_cuda and input.dim() == 2 and \
            input.dtype == torch.float32 and input.shape[0] >= 2 and self.num_features <= 1024
        if not native:
            # eval mode (affine map of the running statistics), 3-D inputs, other dtypes: torch, as in the reference
            return super().forward(input)
        self._check_input_dim(input)
        # the factor torch applies (torch/nn/modules/batchnorm.py:_BatchNorm.forward)
        factor = 0.0 if self.momentum is None else self.momentum
        running_mean = running_var = None
        if self.training and self.track_running_stats:
            running_mean, running_var = self.running_mean, self.running_var
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
                if self.momentum is None:
                    factor = 1.0 / float(self.num_batches_tracked)
        return _BatchNormTrainFn.apply(input, self.weight, self.bias, running_mean, running_var, self.eps, factor,
                                       self._workspace(input.device))
