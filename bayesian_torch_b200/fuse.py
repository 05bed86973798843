"""Inference-time graph fusion for converted torchvision ResNets.

In eval mode the modules that follow every Bayesian convolution in torchvision's BasicBlock / Bottleneck are a
per-channel affine (BatchNorm with running statistics), a residual add and a ReLU.  On the MC-batched
activations ([S*B, C, H, W]) each of them is a full HBM round trip in PyTorch; `fuse_inference(model)` moves them
into the epilogue of the fused conv kernel (BtEpilogue in include/btb200.h):

    conv -> bn -> relu          =>  conv[scale, shift, relu]
    conv -> bn -> (+id) -> relu =>  conv[scale, shift, residual, relu]
    downsample: conv -> bn      =>  conv[scale, shift]

The result is numerically the same function (the affine is applied to the fp32 accumulator instead of the
rounded conv output).  Only modules whose structure is recognised are touched; everything else keeps running
through stock PyTorch.  The model must stay in eval mode afterwards.
"""
import torch
import torch.nn as nn

from ._core import BayesConvBase


def _bn_ok(bn):
    return isinstance(bn, nn.BatchNorm2d) and bn.running_mean is not None and bn.running_var is not None


def _bn_affine(bn):
    if not _bn_ok(bn):
        return None
    w = bn.weight.detach().float() if bn.affine else torch.ones_like(bn.running_mean, dtype=torch.float32)
    b = bn.bias.detach().float() if bn.affine else torch.zeros_like(bn.running_mean, dtype=torch.float32)
    scale = w / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = b - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def _bn_stamp(bn):
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    return tuple((-1, 0) if t is None else (t._version, t.data_ptr()) for t in ts)


def _attach(conv, bn, relu):
    """Fold `bn` (and optionally a ReLU) into conv's epilogue.  The folded scale / shift follow the BatchNorm tensors:
    refresh_epilogue() recomputes them when load_state_dict() or an in-place update changed the statistics."""
    aff = _bn_affine(bn)
    if aff is None or not isinstance(conv, BayesConvBase) or conv._nd != 2:
        return False
    conv._bt_ep_scale, conv._bt_ep_shift = aff
    conv._bt_ep_relu = bool(relu)
    conv._bt_ep_bn = (bn, _bn_stamp(bn))
    return True


def refresh_epilogues(model):
    """Recompute every folded BatchNorm affine whose source statistics changed (version stamp); called by
    mc_predict before it looks up / captures a CUDA graph, and usable directly after load_state_dict()."""
    changed = False
    for m in model.modules():
        src = getattr(m, "_bt_ep_bn", None)
        if src is not None and _bn_stamp(src[0]) != src[1]:
            dev = m._bt_ep_scale.device
            sc, sh = _bn_affine(src[0])
            m._bt_ep_scale, m._bt_ep_shift = sc.to(dev), sh.to(dev)
            m._bt_ep_bn = (src[0], _bn_stamp(src[0]))
            changed = True
    return changed


class FusedBasicBlock(nn.Module):
    """torchvision.models.resnet.BasicBlock.forward with BN / ReLU / residual folded into the conv epilogues."""

    def __init__(self, block):
        super().__init__()
        self.conv1, self.conv2 = block.conv1, block.conv2
        self.bn1, self.bn2 = block.bn1, block.bn2            # kept (unused in forward) so state_dict is unchanged
        self.downsample = block.downsample
        self.stride = block.stride
        ok1 = _attach(self.conv1, block.bn1, relu=True)
        ok2 = _attach(self.conv2, block.bn2, relu=True)       # relu after the residual add
        assert ok1 and ok2, "fuse_inference: _block_ok() admitted a block whose BatchNorm cannot be folded"
        self._ds_fused = False
        if self.downsample is not None and len(self.downsample) == 2:
            self._ds_fused = _attach(self.downsample[0], self.downsample[1], relu=False)

    def forward(self, x):
        if self.downsample is None:
            identity = x
        elif self._ds_fused:
            identity = self.downsample[0](x)
        else:
            identity = self.downsample(x)
        out = self.conv1(x)
        return self.conv2._forward_impl(out, False, residual=identity)


class FusedBottleneck(nn.Module):
    def __init__(self, block):
        super().__init__()
        self.conv1, self.conv2, self.conv3 = block.conv1, block.conv2, block.conv3
        self.bn1, self.bn2, self.bn3 = block.bn1, block.bn2, block.bn3
        self.downsample = block.downsample
        self.stride = block.stride
        oks = [_attach(self.conv1, block.bn1, relu=True), _attach(self.conv2, block.bn2, relu=True),
               _attach(self.conv3, block.bn3, relu=True)]
        assert all(oks), "fuse_inference: _block_ok() admitted a block whose BatchNorm cannot be folded"
        self._ds_fused = False
        if self.downsample is not None and len(self.downsample) == 2:
            self._ds_fused = _attach(self.downsample[0], self.downsample[1], relu=False)

    def forward(self, x):
        if self.downsample is None:
            identity = x
        elif self._ds_fused:
            identity = self.downsample[0](x)
        else:
            identity = self.downsample(x)
        out = self.conv2(self.conv1(x))
        return self.conv3._forward_impl(out, False, residual=identity)


class _Folded(nn.Module):
    """A module whose arithmetic was folded into the preceding conv's epilogue: forward is the identity, the original
    module stays registered so that its parameters / buffers keep their state_dict keys (load_state_dict followed by
    refresh_epilogues() re-folds them)."""

    def __init__(self, inner):
        super().__init__()
        self._bt_inner = [inner]          # not a registered child: the keys must stay 'bn1.weight', not 'bn1.inner.weight'
        for n, p in inner.named_parameters(recurse=False):
            self.register_parameter(n, p)
        for n, b in inner.named_buffers(recurse=False):
            self.register_buffer(n, b)

    def forward(self, x):
        return x


class FusedMaxPool2d(nn.Module):
    """nn.MaxPool2d (floor mode, dilation 1) on channels-last CUDA activations through bt_maxpool2d_nhwc;
    anything else is handed to the original module."""

    def __init__(self, pool):
        super().__init__()
        self.pool = pool
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        self.k, self.s, self.p = pair(pool.kernel_size), pair(pool.stride or pool.kernel_size), pair(pool.padding)
        self.ok = pair(pool.dilation) == (1, 1) and not pool.ceil_mode and not pool.return_indices

    def forward(self, x):
        from . import _native
        if getattr(x, "_bt_pooled", False):      # the producing conv kernel pooled inside its epilogue (_core._launch)
            return x
        ve = 8 if x.dtype == torch.bfloat16 else 4
        if not (self.ok and x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float32)
                and x.shape[1] % ve == 0):
            return self.pool(x)
        xp = x.permute(0, 2, 3, 1)
        if not xp.is_contiguous():
            xp = xp.contiguous()
        return _native.maxpool2d_nhwc(xp, self.k, self.s, self.p).permute(0, 3, 1, 2)


class FusedAvgPool(nn.Module):
    """nn.AdaptiveAvgPool2d((1, 1)) that passes a [N, C, 1, 1] input through: the mean of one element is that element
    (bit-exact), and at CIFAR resolution the ResNet head is exactly that case -- one ATen reduction launch (7-9 us, 2% of
    a step at 8 MC samples per rank) less."""

    def __init__(self, pool):
        super().__init__()
        self.pool = pool

    def forward(self, x):
        if x.dim() == 4 and x.shape[2] == 1 and x.shape[3] == 1:
            return x
        return self.pool(x)


def _block_ok(block, convs_bns):
    """every conv is a Bayesian 2-D conv converted by dnn_to_bnn AND every norm is a BatchNorm2d with running statistics
    (norm_layer=GroupNorm, track_running_stats=False or an Identity left by an earlier fold keep the stock block)"""
    if not isinstance(getattr(block, "relu", None), nn.ReLU):
        return False
    for cname, bname in convs_bns:
        c, bn = getattr(block, cname, None), getattr(block, bname, None)
        if not isinstance(c, BayesConvBase) or not c.dnn_to_bnn_flag or c._nd != 2 or not _bn_ok(bn):
            return False
        if bn.num_features != c.out_channels:
            return False
    return True


def fuse_inference(model):
    """In-place; returns the model.  Requires model.eval() (BatchNorm running statistics are baked in)."""
    if model.training:
        raise RuntimeError("fuse_inference folds BatchNorm running statistics: call model.eval() first")
    try:
        from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet
    except Exception:                                         # torchvision absent: nothing to recognise
        return model

    def walk(parent):
        for name, child in list(parent._modules.items()):
            if child is None:
                continue
            if type(child) is BasicBlock and _block_ok(child, (("conv1", "bn1"), ("conv2", "bn2"))):
                setattr(parent, name, FusedBasicBlock(child))
            elif type(child) is Bottleneck and _block_ok(child, (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"))):
                setattr(parent, name, FusedBottleneck(child))
            else:
                walk(child)

    walk(model)
    if isinstance(model, ResNet) and isinstance(model.conv1, BayesConvBase) and isinstance(model.relu, nn.ReLU):
        if _attach(model.conv1, model.bn1, relu=True):
            # bn1 stays registered (state_dict keys unchanged) but ResNet._forward_impl calls self.bn1 / self.relu:
            # route them through pass-through wrappers that keep the original module as a child
            model.bn1 = _Folded(model.bn1)
            model.relu = nn.Identity()
    if isinstance(model, ResNet) and type(model.maxpool) is nn.MaxPool2d:
        model.maxpool = FusedMaxPool2d(model.maxpool)
        mp = model.maxpool
        # conv1 -> bn1 -> relu -> maxpool(3, 2, 1): offer the pool to conv1's kernel (taken when its tiling allows)
        if isinstance(model.bn1, _Folded) and mp.ok and (mp.k, mp.s, mp.p) == ((3, 3), (2, 2), (1, 1)):
            model.conv1._bt_ep_pool = True
    if isinstance(model, ResNet) and type(model.avgpool) is nn.AdaptiveAvgPool2d and \
            tuple(model.avgpool.output_size if isinstance(model.avgpool.output_size, (tuple, list))
                  else (model.avgpool.output_size,) * 2) == (1, 1):
        model.avgpool = FusedAvgPool(model.avgpool)
    model._bt_fused_inference = True
    return model
