"""Training path: autograd.Function wrappers around the fused forward and the KL kernel.

The reference layers are ordinary differentiable PyTorch modules (layers/variational_layers/linear_variational.py:157-201
under autograd; training step examples/main_bayesian_cifar_dnn2bnn.py:404-420: loss = CE + KL / batch; loss.backward()).
Here the FORWARD stays the single fused sm_100a launch; nothing weight-sized is saved for the backward -- eps and the
Flipout signs are regenerated from the Philox key (seed, layer_key, sample index) with bt_rng_export, exactly the draws
the forward used:

    sigma = softplus(rho),  dsigma/drho = sigmoid(rho)
    reparam:  W = mu + sigma*eps                     flipout:  D = sigma*eps
      dW   = wgrad(x, dy)                              dmu  = wgrad(x, dy)
      dmu  = dW,  drho = dW * eps * sigmoid(rho)       dD   = wgrad(x*s_in, dy*s_out),  drho = dD * eps * sigmoid(rho)
      dx   = igrad(dy, W)                              dx   = igrad(dy, mu) + igrad(dy*s_out, D) * s_in
    KL (mean over n elements, prior N(pm, ps)):
      dKL/dmu = (mu - pm) / (ps^2 n),   dKL/drho = (sigma/ps^2 - 1/sigma) * sigmoid(rho) / n

(oracle/bt_oracle_grad.py states the same formulas on the CPU and is pinned on gradients minted from the reference's
own autograd; tests/test_gpu_grad.py holds this module to it.)  The data-gradient / weight-gradient GEMMs and
convolutions of the backward run through ATen's convolution_backward (cuDNN / cuBLAS): the backward is a "next" row of
the scope table, not the hot path -- the forward never touches ATen.
"""
import torch

from . import _native


def _logical_eps_from_debug(layer, e_phys):
    """debug eps is handed to the kernel in PHYSICAL [Cout, taps, Cin/g] order; the backward wants the parameter's shape"""
    mu_w = layer._mu_rho()[0]
    if mu_w.dim() == 2:
        return e_phys.view_as(mu_w)
    nd = mu_w.dim() - 2
    phys = e_phys.view(mu_w.shape[0], *mu_w.shape[2:], mu_w.shape[1])
    return phys.permute(0, nd + 1, *range(1, nd + 1))


def _conv_args(layer):
    from ._core import _tuple
    nd = layer._nd
    return (_tuple(layer.stride, nd), _tuple(layer.padding, nd), _tuple(layer.dilation, nd), layer.groups)


def _backward_products(layer, x, w, gy, need_x, need_w):
    """(dx, dw, db_sum) of out = conv(x, w) [+ b] for this layer's geometry, through ATen."""
    if layer._nd == 0:
        g2 = gy.reshape(-1, gy.shape[-1])
        dx = (g2 @ w).view_as(x) if need_x else None
        dw = g2.t() @ x.reshape(-1, x.shape[-1]) if need_w else None
        return dx, dw
    st, pd, dl, groups = _conv_args(layer)
    transposed = bool(getattr(layer, "_transposed", False))
    out_pad = list(getattr(layer, "_out_pad", (0,) * layer._nd))
    dx, dw, _ = torch.ops.aten.convolution_backward(gy, x, w, None, list(st), list(pd), list(dl), transposed, out_pad,
                                                    groups, [need_x, need_w, False])
    return dx, dw


class _BayesForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, debug, x, mu_w, rho_w, mu_b, rho_b):
        out = layer._launch(x, False, debug)
        ctx.layer, ctx.debug, ctx.last = layer, debug, dict(layer._bt_last)
        ctx.save_for_backward(x, mu_w, rho_w, mu_b, rho_b)
        ctx.out_shape = tuple(out.shape)
        return out

    @staticmethod
    def backward(ctx, gy):
        layer, dbg = ctx.layer, (ctx.debug or {})
        x, mu_w, rho_w, mu_b, rho_b = ctx.saved_tensors
        need = ctx.needs_input_grad          # (layer, debug, x, mu_w, rho_w, mu_b, rho_b)
        flip = layer._family == "flipout"
        cdt = x.dtype                        # compute dtype of the backward products
        gy = gy.contiguous() if not gy.is_contiguous() and not gy.is_contiguous(memory_format=torch.channels_last) else gy
        # ---- the draws of the forward, regenerated from the Philox counters (or the injected debug tensors)
        keep = layer._bt_last
        layer._bt_last = ctx.last
        try:
            if dbg.get("eps_w_in") is not None:
                eps_w = _logical_eps_from_debug(layer, dbg["eps_w_in"])
                eps_b = dbg.get("eps_b_in")
            else:
                eps_w, eps_b = layer.materialize_eps(0)
            s_in = s_out = None
            if flip:
                if dbg.get("sign_in") is not None:
                    nd = layer._nd
                    inv = (0, nd + 1, *range(1, nd + 1))
                    s_in = dbg["sign_in"].view(x.shape[0], *x.shape[2:], x.shape[1]).permute(inv) if nd else dbg["sign_in"].view_as(x)
                    s_out = dbg["sign_out"].view(ctx.out_shape[0], *ctx.out_shape[2:], ctx.out_shape[1]).permute(inv) if nd \
                        else dbg["sign_out"].view(ctx.out_shape)
                else:
                    s_in, s_out = layer.materialize_signs(tuple(x.shape), ctx.out_shape, 0)
        finally:
            layer._bt_last = keep
        rho32 = rho_w.detach().float()
        sig = torch.nn.functional.softplus(rho32)
        dsig = torch.sigmoid(rho32)
        eps_w = eps_w.float()
        need_p = need[3] or need[4]
        gx = gmu = grho = gmub = grhob = None
        bdims = tuple(i for i in range(gy.dim()) if i != (gy.dim() - 1 if layer._nd == 0 else 1))
        if not flip:
            w = (mu_w.detach().float() + sig * eps_w).to(cdt)
            gx, dw = _backward_products(layer, x, w, gy, need[2], need_p)
            if need_p:
                dw = dw.float()
                gmu = dw.to(mu_w.dtype) if need[3] else None
                grho = (dw * eps_w * dsig).to(rho_w.dtype) if need[4] else None
            if mu_b is not None and (need[5] or need[6]):
                db = gy.float().sum(bdims)
                gmub = db.to(mu_b.dtype) if need[5] else None
                grhob = (db * eps_b.float() * torch.sigmoid(rho_b.detach().float())).to(rho_b.dtype) if need[6] else None
        else:
            s_in, s_out = s_in.to(cdt), s_out.to(cdt)
            gys = gy * s_out
            d = (sig * eps_w).to(cdt)
            gx1, dmu = _backward_products(layer, x, mu_w.detach().to(cdt), gy, need[2], need[3])
            gx2, dd = _backward_products(layer, x * s_in, d, gys, need[2], need[4])
            if need[2]:
                gx = gx1 + gx2 * s_in
            gmu = dmu.to(mu_w.dtype) if need[3] else None
            grho = (dd.float() * eps_w * dsig).to(rho_w.dtype) if need[4] else None
            if mu_b is not None and (need[5] or need[6]):
                gmub = gy.float().sum(bdims).to(mu_b.dtype) if need[5] else None
                grhob = (gys.float().sum(bdims) * eps_b.float() * torch.sigmoid(rho_b.detach().float())).to(rho_b.dtype) \
                    if need[6] else None
        return None, None, gx, gmu, grho, gmub, grhob


class _KLLoss(torch.autograd.Function):
    """kl_loss() = mean-KL(weight) + mean-KL(bias) (linear_variational.py:144-155): forward = the fused KL kernel,
    backward = the closed form above (elementwise; tensor priors edited after init are honoured)."""

    @staticmethod
    def forward(ctx, layer, mu_w, rho_w, mu_b, rho_b):
        ctx.layer = layer
        ctx.save_for_backward(mu_w, rho_w, mu_b, rho_b)
        return layer._kl_launch()

    @staticmethod
    def backward(ctx, g):
        layer = ctx.layer
        mu_w, rho_w, mu_b, rho_b = ctx.saved_tensors
        need = ctx.needs_input_grad
        uniform = layer._priors_uniform()

        def grads(mu, rho, pm, ps):
            n = mu.numel()
            mu32, rho32 = mu.detach().float(), rho.detach().float()
            sig = torch.nn.functional.softplus(rho32)
            ps2 = ps * ps
            gm = (mu32 - pm) / (ps2 * n)
            gr = (sig / ps2 - 1.0 / sig) * torch.sigmoid(rho32) / n
            return (gm * g).to(mu.dtype), (gr * g).to(rho.dtype)

        pm_w = float(layer.prior_mean) if uniform else layer.prior_weight_mu.float()
        ps_w = float(layer.prior_variance) if uniform else layer.prior_weight_sigma.float()
        gmu, grho = grads(mu_w, rho_w, pm_w, ps_w)
        gmub = grhob = None
        if mu_b is not None:
            pm_b = float(layer.prior_mean) if uniform else layer.prior_bias_mu.float()
            ps_b = float(layer.prior_variance) if uniform else layer.prior_bias_sigma.float()
            gmub, grhob = grads(mu_b, rho_b, pm_b, ps_b)
        return (None, gmu if need[1] else None, grho if need[2] else None,
                gmub if need[3] else None, grhob if need[4] else None)


def forward_with_grad(layer, x, return_kl, debug, residual):
    from ._core import _mc
    if _mc.active and _mc.n_samples > 1:
        raise RuntimeError("bayesian_torch_b200: the MC-sample context (several weight samples per launch) is an "
                           "inference feature; wrap it in torch.no_grad() -- training draws one sample per forward")
    if residual is not None or layer._bt_ep_scale is not None or layer._bt_ep_relu:
        raise RuntimeError("bayesian_torch_b200: the fused inference epilogue (fuse_inference) is not differentiable; "
                           "use it under torch.no_grad() / model.eval() only")
    mu_w, rho_w = layer._phys_params()
    out = _BayesForward.apply(layer, debug, x, mu_w, rho_w, layer.mu_bias, layer.rho_bias)
    if return_kl:
        return out, kl_with_grad(layer)
    return out


def kl_with_grad(layer):
    mu_w, rho_w = layer._phys_params()
    return _KLLoss.apply(layer, mu_w, rho_w, layer.mu_bias, layer.rho_bias).to(mu_w.dtype)
