"""Host side of the B200 Bayesian layers: parameter registration, RNG bookkeeping, the MC-sample
context and the dispatch to the fused kernels.  The public classes in ``layers/`` are thin
subclasses that only fix (family, nd) and mirror the reference constructors.

Reference behaviour mirrored here (paths relative to /root/reference/bayesian_torch/):
  * parameter / buffer names, shapes, N(mu_init, 0.1) / N(rho_init, 0.1) init and its draw order
    (layers/variational_layers/linear_variational.py:87-142, conv_variational.py:272-346,
     flipout_layers/linear_flipout.py:83-135, conv_flipout.py:297-360);
  * forward(x, return_kl=True) -> (out, kl) | out, dnn_to_bnn_flag forcing return_kl=False
    (linear_variational.py:157-201), sampling in train() and eval() alike;
  * kl_loss() = mean-KL(weight) + mean-KL(bias) with prior_variance used as sigma_p
    (linear_variational.py:131-155).
"""
import contextlib
import itertools
import os
import threading

import torch
import torch.nn as nn
from torch.nn import Parameter

from . import _native
from ._base import BaseVariationalLayer_

# ----------------------------------------------------------------------------- RNG bookkeeping
_rng_lock = threading.Lock()
_state = {"seed": None, "epoch": 0}
_layer_counter = itertools.count(1)


def manual_seed(seed):
    """Seed of the on-chip Philox streams (key).  Also restarts every layer's draw counter."""
    with _rng_lock:
        _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
        _state["epoch"] += 1


def current_seed():
    if _state["seed"] is None:
        manual_seed(torch.initial_seed())
    return _state["seed"]


def assign_layer_keys(model):
    """Deterministic Philox layer keys = index in model.modules() (identical on every rank)."""
    for idx, m in enumerate(model.modules()):
        if isinstance(m, BayesLayerBase):
            m._bt_layer_key = idx + 1
    return model


class _MC(threading.local):
    def __init__(self):
        self.active = False
        self.n_samples = 1
        self.batch = None
        self.sample0 = 0
        self.sample_word = None      # int32 CUDA tensor [1]: run-time addend of the sample index (CUDA-graph replays)
        self.sample_word_value = 0   # what the host last wrote into it


_mc = _MC()


@contextlib.contextmanager
def mc_sample_context(n_samples, batch, sample0, sample_word=None, sample_word_value=0):
    """Inside this context every Bayesian layer evaluates `n_samples` independent weight samples
    in ONE launch: activations carry the samples stacked along the batch dimension
    ([n_samples * batch, ...]); an input whose batch dimension is `batch` is shared by all samples
    (first layer).  Sample s uses the global Philox sample index sample0 + s (+ *sample_word when given: a device
    word the kernels read at run time, so a captured CUDA graph draws fresh eps on every replay).  The context
    disables autograd (MC stacking is an inference feature; training draws one sample per forward)."""
    prev = (_mc.active, _mc.n_samples, _mc.batch, _mc.sample0, _mc.sample_word, _mc.sample_word_value)
    _mc.active, _mc.n_samples, _mc.batch, _mc.sample0 = True, int(n_samples), int(batch), int(sample0)
    _mc.sample_word, _mc.sample_word_value = sample_word, int(sample_word_value)
    try:
        with torch.no_grad():      # an inference context by construction: several weight samples per launch
            yield
    finally:
        _mc.active, _mc.n_samples, _mc.batch, _mc.sample0, _mc.sample_word, _mc.sample_word_value = prev


def _SIGMA_CACHE_ENABLED():
    return os.environ.get("BT_DISABLE_SIGMA_CACHE") is None     # A/B switch (benchmarks)


def _tuple(v, n):
    if isinstance(v, (tuple, list)):
        if len(v) != n:
            raise ValueError(f"expected {n} values, got {tuple(v)}")
        return tuple(int(a) for a in v)
    return (int(v),) * n


def _is_channels_last(t):
    """True if `t` ([N, C, *spatial]) is dense with C innermost (physical [N, *spatial, C])."""
    perm = (0, *range(2, t.dim()), 1)
    return t.permute(perm).is_contiguous()


def _to_channels_last(t):
    perm = (0, *range(2, t.dim()), 1)
    inv = (0, t.dim() - 1, *range(1, t.dim() - 1))
    return t.permute(perm).contiguous().permute(inv)


class BayesLayerBase(BaseVariationalLayer_):
    """Common machinery.  `_family` in {"reparam", "flipout"}; `_nd` = 0 (linear) or 1/2/3."""

    _family = "reparam"
    _nd = 0

    def __init__(self):
        super().__init__()
        self._bt_layer_key = next(_layer_counter)
        self._bt_calls = 0
        self._bt_epoch = -1
        self._bt_last = None  # (seed, layer_key, sample0, n_samples, x_rows_per_sample, out_rows_per_sample)
        self._bt_prior_versions = None
        self._bt_prior_uniform = True
        # fused inference epilogue (set by bayesian_torch_b200.fuse.fuse_inference): eval-mode BatchNorm folded
        # to a per-channel affine and / or ReLU applied inside the conv kernel's epilogue
        self._bt_ep_scale = None
        self._bt_ep_shift = None
        self._bt_ep_relu = False
        self._bt_ep_pool = False    # fuse_inference: a 3x3 / stride-2 / pad-1 max-pool follows the folded bn + relu
        # the repacked / transformed parameter caches follow Tensor._version; load_state_dict() copies in place (seen),
        # but `.data` writes are not -- invalidate_caches() is the explicit hook for those
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_caches())

    def invalidate_caches(self):
        """Drop every cached transform of the parameters (sigma = softplus(rho), channel-padded / im2col repacks, prior
        uniformity).  Needed after writing parameters through `.data` (e.g. MOPED-style `rho.data.copy_(...)`,
        utils/util.py:102,115), which does not bump Tensor._version."""
        self._bt_sigma_cache = None
        self._bt_prior_versions = None
        if hasattr(self, "_bt_pad_cache"):
            self._bt_pad_cache = None
        if hasattr(self, "_bt_tr_cache"):
            self._bt_tr_cache = None

    # ---- registration helpers
    def _register(self, wname, wshape, out_features, bias, mu_init, rho_init):
        self._wname = wname
        self.register_parameter(f"mu_{wname}", Parameter(torch.empty(*wshape)))
        self.register_parameter(f"rho_{wname}", Parameter(torch.empty(*wshape)))
        self.register_buffer(f"eps_{wname}", torch.empty(*wshape), persistent=False)
        self.register_buffer("prior_weight_mu", torch.empty(*wshape), persistent=False)
        self.register_buffer("prior_weight_sigma", torch.empty(*wshape), persistent=False)
        if bias:
            self.mu_bias = Parameter(torch.empty(out_features))
            self.rho_bias = Parameter(torch.empty(out_features))
            self.register_buffer("eps_bias", torch.empty(out_features), persistent=False)
            self.register_buffer("prior_bias_mu", torch.empty(out_features), persistent=False)
            self.register_buffer("prior_bias_sigma", torch.empty(out_features), persistent=False)
        else:
            self.register_parameter("mu_bias", None)
            self.register_parameter("rho_bias", None)
            self.register_buffer("eps_bias", None, persistent=False)
            self.register_buffer("prior_bias_mu", None, persistent=False)
            self.register_buffer("prior_bias_sigma", None, persistent=False)
        self._mu_init, self._rho_init = float(mu_init), float(rho_init)
        self.init_parameters()

    def init_parameters(self):
        """Same fills and the same RNG draw order as the reference init
        (reparam: linear_variational.py:131-142; flipout conv draws the bias params before the
        prior fills, which consumes no RNG: conv_flipout.py:346-360)."""
        mu_w, rho_w = self._mu_rho()
        self.prior_weight_mu.fill_(self.prior_mean)
        self.prior_weight_sigma.fill_(self.prior_variance)
        mu_w.data.normal_(mean=self._mu_init, std=0.1)
        rho_w.data.normal_(mean=self._rho_init, std=0.1)
        if self.mu_bias is not None:
            self.prior_bias_mu.fill_(self.prior_mean)
            self.prior_bias_sigma.fill_(self.prior_variance)
            self.mu_bias.data.normal_(mean=self._mu_init, std=0.1)
            self.rho_bias.data.normal_(mean=self._rho_init, std=0.1)
        self._bt_prior_versions = None

    def _mu_rho(self):
        return getattr(self, f"mu_{self._wname}"), getattr(self, f"rho_{self._wname}")

    # ---- priors: constant fills are passed as scalars (no HBM traffic); tensors that were edited
    # after init (e.g. MOPED-style priors, utils/util.py:102,115) are passed as tensors.
    def _priors_uniform(self):
        bufs = [self.prior_weight_mu, self.prior_weight_sigma, self.prior_bias_mu, self.prior_bias_sigma]
        versions = tuple(-1 if b is None else b._version for b in bufs)
        if versions != self._bt_prior_versions:
            uni = True
            for b, v in ((bufs[0], self.prior_mean), (bufs[1], self.prior_variance),
                         (bufs[2], self.prior_mean), (bufs[3], self.prior_variance)):
                if b is not None and not bool((b == v).all()):
                    uni = False
                    break
            self._bt_prior_uniform = uni
            self._bt_prior_versions = versions
        return self._bt_prior_uniform

    # ---- physical layout: weights are kept dense with the input channel innermost
    # ([Cout, *k, Cin/g] in memory, logical shape unchanged -> state_dict compatible)
    def _phys_params(self):
        mu_w, rho_w = self._mu_rho()
        if self._nd > 0:
            for prm in (mu_w, rho_w):
                if not _is_channels_last(prm.data):
                    prm.data = _to_channels_last(prm.data)
        else:
            for prm in (mu_w, rho_w):
                if not prm.data.is_contiguous():
                    prm.data = prm.data.contiguous()
        return mu_w, rho_w

    def _check_param(self, t, name):
        _native.require_cuda(t, name)

    def _sigma_of(self, rho_k, pmode):
        """softplus(rho) in rho's dtype and physical layout, cached per (tensor, version).  Like the repacked parameter
        caches it follows `Tensor._version`, i.e. in-place updates made through `.data` are NOT seen -- MC inference
        (the only user) runs on frozen parameters."""
        rho_w = self._mu_rho()[1]
        key = (rho_k.data_ptr(), rho_w._version, rho_k.dtype, rho_k.device, pmode, getattr(self, "_bt_kalign", 8))
        c = getattr(self, "_bt_sigma_cache", None)
        if c is None or c[0] != key:
            sig = torch.nn.functional.softplus(rho_k.float()).to(rho_k.dtype)
            if sig.stride() != rho_k.stride():                     # keep the kernel's physical layout
                sig = torch.empty_like(rho_k).copy_(sig)
            self._bt_sigma_cache = c = (key, sig)
        return c[1]

    def _padded_cin(self):
        """Input channels as the kernel sees them (convs with Cin % 8 != 0 are zero-padded, see BayesConvBase)."""
        return None

    def _kernel_params(self, pmode=None):
        mu_w, rho_w = self._phys_params()
        return mu_w.data, rho_w.data

    # ---- KL
    def kl_loss(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters(recurse=False)):
            from ._autograd import kl_with_grad
            return kl_with_grad(self)
        return self._kl_launch()

    def _kl_launch(self):
        mu_w, rho_w = self._phys_params()
        self._check_param(mu_w, f"mu_{self._wname}")
        mu_b, rho_b = self.mu_bias, self.rho_bias
        if self._priors_uniform():
            kl = _native.kl_gaussian(mu_w.data, rho_w.data, None, None,
                                     None if mu_b is None else mu_b.data, None if rho_b is None else rho_b.data,
                                     None, None, self.prior_mean, self.prior_variance)
        else:
            pm, ps = self.prior_weight_mu, self.prior_weight_sigma
            if self._nd > 0:
                if not _is_channels_last(pm):
                    pm = _to_channels_last(pm)
                if not _is_channels_last(ps):
                    ps = _to_channels_last(ps)
            kl = _native.kl_gaussian(mu_w.data, rho_w.data, pm, ps,
                                     None if mu_b is None else mu_b.data, None if rho_b is None else rho_b.data,
                                     self.prior_bias_mu, self.prior_bias_sigma, self.prior_mean, self.prior_variance)
        return kl.to(mu_w.dtype)

    # ---- sample bookkeeping
    def _next_sample(self):
        if _mc.active:
            return _mc.sample0, _mc.n_samples
        if self._bt_epoch != _state["epoch"]:
            self._bt_epoch = _state["epoch"]
            self._bt_calls = 0
        idx = self._bt_calls
        self._bt_calls += 1
        return idx, 1

    # ---- geometry
    def _geometry(self, x, n_samples):
        raise NotImplementedError

    def _forward_impl(self, x, return_kl, debug=None, residual=None):
        """forward(x, return_kl): inference / no-grad calls go straight to the fused kernel; when autograd is recording
        and the input or a parameter requires grad, the same launch is wrapped in an autograd.Function whose backward
        regenerates eps / signs from the Philox key (nothing weight-sized is stored) -- see _autograd.py."""
        if self.dnn_to_bnn_flag:
            return_kl = False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters(recurse=False))):
            from ._autograd import forward_with_grad
            return forward_with_grad(self, x, return_kl, debug, residual)
        return self._launch(x, return_kl, debug, residual)

    def _launch(self, x, return_kl, debug=None, residual=None):
        _native.require_cuda(x, "input")
        mu_w, rho_w = self._phys_params()
        self._check_param(mu_w, f"mu_{self._wname}")
        if x.device != mu_w.device:
            raise RuntimeError(f"input on {x.device} but parameters on {mu_w.device}")
        seed = current_seed()
        sample0, n_samples = self._next_sample()
        x_phys, geom, out_shape_phys, to_logical = self._geometry(x, n_samples)
        pmode = getattr(self, "_bt_pmode", None)       # None | "pad" (zero channels) | "im2col" (materialised stem)
        mu_k, rho_k = self._kernel_params(pmode)
        padded = pmode is not None
        # MC inference evaluates many weight samples of frozen parameters: sigma = softplus(rho) is the same for all of
        # them, so it is computed once per parameter version and the kernels skip 2 of their 4 MUFU ops per sampled
        # weight (geom.rho_is_sigma).  Only inside mc_sample_context and never with the KL side output / debug hooks.
        geom.rho_is_sigma = 0
        word_value = 0
        if _mc.active and _mc.sample_word is not None:
            geom.sample_offset = _mc.sample_word.data_ptr()
            word_value = _mc.sample_word_value
        if _mc.active and not return_kl and not debug and _SIGMA_CACHE_ENABLED():
            rho_k = self._sigma_of(rho_k, pmode)
            geom.rho_is_sigma = 1
        # fuse_inference: torchvision's stem max-pool (3x3, stride 2, padding 1) right behind this conv's bn + relu runs
        # inside the kernel when it can keep whole output rows in a tile (BtForwardPlan.pool_fused); the tensor handed
        # back is marked so that FusedMaxPool2d passes it through.
        pooled = False
        if self._bt_ep_pool and self._nd == 2 and residual is None and not debug and not return_kl \
                and not torch.is_grad_enabled():
            pooled = self._plan_pool(geom, x_phys.dtype, mu_k.dtype)
            if pooled:
                oh, ow = self._bt_outsp
                out_shape_phys = (out_shape_phys[0], oh // 2, ow // 2, out_shape_phys[-1])
        out = torch.empty(out_shape_phys, dtype=x_phys.dtype, device=x.device)
        kl = None
        kl_via_kernel = return_kl and self._priors_uniform() and not padded
        if kl_via_kernel:
            kl = torch.empty((), dtype=torch.float32, device=x.device)
        dbg = dict(debug or {})
        if padded and dbg:
            dbg = self._pad_debug(dbg, pmode)
        if self._bt_ep_scale is not None:
            if self._bt_ep_scale.device != x.device:
                self._bt_ep_scale = self._bt_ep_scale.to(x.device)
                self._bt_ep_shift = self._bt_ep_shift.to(x.device)
            dbg.update(ep_scale=self._bt_ep_scale, ep_shift=self._bt_ep_shift)
        if self._bt_ep_relu:
            dbg["ep_relu"] = True
        if residual is not None:
            dbg["ep_residual"] = self._residual_phys(residual, out_shape_phys)
        if _native.timing_hook is not None:      # bench.py roofline pass: logical sizes of this launch
            dbg["info"] = self._logical_info(x, geom)
        _native.layer_forward(
            _native.MODE_FLIPOUT if self._family == "flipout" else _native.MODE_REPARAM, geom, x_phys,
            mu_k, rho_k, None if self.mu_bias is None else self.mu_bias.data,
            None if self.rho_bias is None else self.rho_bias.data, out,
            kl_out=kl, prior_mu=self.prior_mean, prior_sigma=self.prior_variance,
            seed=seed, layer_key=self._bt_layer_key, sample0=sample0, **dbg)
        self._bt_last = dict(seed=seed, layer_key=self._bt_layer_key, sample0=(sample0 + word_value) & 0xFFFFFFFF,
                             n_samples=n_samples, geom=geom, pmode=pmode)
        result = to_logical(out)
        if pooled:
            result._bt_pooled = True
        if return_kl:
            if kl is None:
                kl = self.kl_loss()
            return result, kl.to(mu_w.dtype)
        return result

    def _plan_pool(self, geom, x_dtype, p_dtype):
        """set geom.pool_hw when the kernel selected for this launch pools inside its epilogue (asked once per shape)"""
        oh, ow = self._bt_outsp
        if oh % 2 or ow % 2:
            return False
        geom.pool_hw[0], geom.pool_hw[1] = oh, ow
        key = (geom.n_samples, geom.x_shared, geom.batch, geom.c_in, geom.c_out, tuple(geom.out_dhw), oh, ow, x_dtype, p_dtype,
               geom.rho_is_sigma)
        cache = self.__dict__.setdefault("_bt_pool_plans", {})
        fused = cache.get(key)
        if fused is None:
            mode = _native.MODE_FLIPOUT if self._family == "flipout" else _native.MODE_REPARAM
            plan = _native.plan_forward(mode, geom, x_dtype, p_dtype, sm_count=_native.sm_count())
            fused = cache[key] = bool(plan["pool_fused"])
        if not fused:
            geom.pool_hw[0] = geom.pool_hw[1] = 0
        return fused

    def forward(self, input, return_kl=True):
        return self._forward_impl(input, return_kl)

    def _logical_info(self, x, geom):
        """logical element counts of one launch for roofline accounting: the layer's own input (not a materialised
        im2col / channel-padded copy), weights, and K with / without the filter taps that only ever see zero padding"""
        mu_w = self._mu_rho()[0]
        taps_all, taps_used = 1, 1
        for i in range(3):
            k, d, p_, s_, n_in, n_out = geom.k_dhw[i], geom.dil[i], geom.pad[i], geom.stride[i], geom.in_dhw[i], geom.out_dhw[i]
            taps_all *= k
            if geom.transposed or getattr(self, "_bt_pmode", None) == "im2col":
                taps_used *= k
            else:
                taps_used *= sum(1 for kk in range(k) if any(0 <= o * s_ - p_ + kk * d < n_in for o in range(n_out)))
        cin_g = mu_w.numel() // mu_w.shape[0] // max(1, (mu_w[0, 0].numel() if mu_w.dim() > 2 else 1))
        if getattr(self, "_transposed", False):
            cin_g = self.in_channels // self.groups
        ktaps = mu_w[0, 0].numel() if mu_w.dim() > 2 else 1
        per_sample_x = x.numel() // (1 if geom.x_shared else geom.n_samples)
        return {"x_logical_numel": per_sample_x, "w_numel": mu_w.numel(), "b_numel": 0 if self.mu_bias is None else self.mu_bias.numel(),
                "k_logical": ktaps * cin_g, "k_used": (taps_used if ktaps == taps_all else ktaps) * cin_g,
                "flipout": self._family == "flipout", "layer": type(self).__name__}

    def _residual_phys(self, residual, out_shape_phys):
        """Residual (logical N C *sp) -> dense channels-last memory matching the kernel's output."""
        nd = residual.dim() - 2
        r = residual.permute(0, *range(2, nd + 2), 1) if nd > 0 else residual
        if not r.is_contiguous():
            r = r.contiguous()
        if tuple(r.shape) != tuple(out_shape_phys):
            raise RuntimeError(f"residual shape {tuple(residual.shape)} does not match the layer output")
        return r

    # ---- the reference's eps buffers, materialised on demand from the Philox counters
    def materialize_eps(self, sample=0):
        """Fill eps_<weight|kernel> / eps_bias with the draw used for MC sample `sample` of the
        LAST forward (the reference keeps them after every forward, linear_variational.py:161,173)."""
        if self._bt_last is None:
            raise RuntimeError("materialize_eps() needs a previous forward()")
        last = self._bt_last
        mu_w, _ = self._mu_rho()
        cout = mu_w.shape[0]
        taps = 1
        for s in mu_w.shape[2:]:
            taps *= s
        # channels per tap as the kernel counted them (zero-channel padding widens the tap, im2col does not)
        cin_k = (self._padded_cin() if last.get("pmode") == "pad" else None) or mu_w.shape[1]
        kk = cin_k * taps
        eps_w = torch.empty((cout, cin_k, *mu_w.shape[2:]), dtype=torch.float32, device=mu_w.device)
        _native.rng_export(0, eps_w, cout, kk, taps, kk, last["seed"], last["layer_key"], last["sample0"] + sample)
        if cin_k != mu_w.shape[1]:
            eps_w = eps_w[:, :mu_w.shape[1]].contiguous()
        setattr(self, f"eps_{self._wname}", eps_w.to(mu_w.dtype))
        eps_b = None
        if self.mu_bias is not None:
            eps_b = torch.empty(cout, dtype=torch.float32, device=mu_w.device)
            _native.rng_export(1, eps_b, cout, 1, 1, 1, last["seed"], last["layer_key"], last["sample0"] + sample)
            self.eps_bias = eps_b.to(mu_w.dtype)
        return eps_w, eps_b   # fp32 (exactly what the kernel used); the buffers hold them in the parameter dtype

    def materialize_signs(self, x_shape, out_shape, sample=0):
        """Flipout only: the +-1 tensors (logical NC... layout) of MC sample `sample` of the last forward."""
        last = self._bt_last
        dev = self._mu_rho()[0].device

        def gen(what, shape, cpg):
            rows = shape[0]
            for s in shape[2:]:
                rows *= s
            ch = shape[1]
            t = torch.empty((rows, ch), dtype=torch.float32, device=dev)
            _native.rng_export(what, t, rows, ch, 1, cpg, last["seed"], last["layer_key"], last["sample0"] + sample)
            phys = t.view(shape[0], *shape[2:], ch)
            return phys.permute(0, phys.dim() - 1, *range(1, phys.dim() - 1))

        groups = getattr(self, "groups", 1)
        return gen(2, tuple(x_shape), x_shape[1]), gen(3, tuple(out_shape), out_shape[1] // groups)


class BayesLinearBase(BayesLayerBase):
    _nd = 0

    def _init_linear(self, in_features, out_features, prior_mean, prior_variance, posterior_mu_init,
                     posterior_rho_init, bias):
        self.in_features = in_features
        self.out_features = out_features
        self.prior_mean = prior_mean
        self.prior_variance = prior_variance
        self.bias = bias
        self._register("weight", (out_features, in_features), out_features, bias, posterior_mu_init,
                       posterior_rho_init)

    def _geometry(self, x, n_samples):
        if x.shape[-1] != self.in_features:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} x "
                               f"{self.in_features}x{self.out_features})")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows = x2.shape[0]
        shared = 0
        if n_samples > 1:
            if _mc.batch is not None and lead[0] == _mc.batch:
                shared = 1
                rows_per_sample = rows
            else:
                if rows % n_samples:
                    raise RuntimeError(f"MC context: {rows} rows not divisible by {n_samples} samples")
                rows_per_sample = rows // n_samples
        else:
            rows_per_sample = rows
        g = _native.BtLayerGeom()
        g.n_samples, g.x_shared, g.batch = n_samples, shared, rows_per_sample
        g.c_in, g.c_out, g.groups = self.in_features, self.out_features, 1
        for i in range(3):
            g.in_dhw[i] = g.out_dhw[i] = g.k_dhw[i] = g.stride[i] = g.dil[i] = 1
            g.pad[i] = 0
        out_rows = rows_per_sample * n_samples
        out_lead = tuple(lead) if not shared else (lead[0] * n_samples, *lead[1:])

        def to_logical(o):
            return o.view(*out_lead, self.out_features)

        return x2, g, (out_rows, self.out_features), to_logical


class BayesConvBase(BayesLayerBase):
    def _init_conv(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                   prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias, validate):
        nd = self._nd
        if validate:  # conv_variational.py:98-101 (the Flipout classes do not validate)
            if in_channels % groups != 0:
                raise ValueError('invalid in_channels size')
            if out_channels % groups != 0:
                raise ValueError('invalid in_channels size')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups
        self.prior_mean = prior_mean
        self.prior_variance = prior_variance
        self.bias = bias
        ks = _tuple(kernel_size, nd)  # accepts int or tuple (superset of the reference's Conv1d)
        self._bt_pad_cache = None
        self._register("kernel", (out_channels, in_channels // groups, *ks), out_channels, bias,
                       posterior_mu_init, posterior_rho_init)

    # ---- channel padding: an ungrouped conv whose Cin is not a multiple of 8 (e.g. the RGB stem) would need the
    # kernel's scalar gather; instead the activations get zero channels up to the next multiple of 8 and the kernel
    # reads a cached repack of (mu, rho) with mu = 0 / rho = -100 (sigma = 0 -> W = 0 exactly) in the pad channels.
    # eps counters then live in the padded (tap, channel) space; materialize_eps() slices the real channels out.
    def _padded_cin(self):
        cin = self.in_channels
        if self.groups != 1 or cin % 8 == 0:
            return None
        return (cin + 7) // 8 * 8

    def _kernel_params(self, pmode=None):
        mu_w, rho_w = self._phys_params()
        if pmode is None:
            return mu_w.data, rho_w.data
        cp = self._padded_cin()
        kal = getattr(self, "_bt_kalign", 8)
        key = (pmode, kal, mu_w._version, rho_w._version, mu_w.data_ptr(), rho_w.data_ptr(), mu_w.dtype, mu_w.device)
        if self._bt_pad_cache is None or self._bt_pad_cache[0] != key:
            nd = self._nd
            perm = (0, *range(2, nd + 2), 1)

            def pad(t, fill):
                phys = t.data.permute(perm)
                if pmode == "im2col":      # [Cout, taps * Cin] rows, zero-extended to a multiple of 8 / 64 columns
                    flat = phys.reshape(phys.shape[0], -1)
                    kpad = (flat.shape[1] + kal - 1) // kal * kal
                    out = flat.new_full((flat.shape[0], kpad), fill)
                    out[:, : flat.shape[1]] = flat
                    return out
                out = phys.new_full((*phys.shape[:-1], cp), fill)
                out[..., : phys.shape[-1]] = phys
                return out

            self._bt_pad_cache = (key, pad(mu_w, 0.0), pad(rho_w, -100.0))
        return self._bt_pad_cache[1], self._bt_pad_cache[2]

    def _pad_debug(self, dbg, pmode="pad"):
        cp, cin = self._padded_cin(), self.in_channels
        out = dict(dbg)
        e = dbg.get("eps_w_in")
        if e is not None and pmode == "im2col":   # [Cout, taps * Cin] -> zero-extend the columns
            kal = getattr(self, "_bt_kalign", 8)
            kpad = (e.shape[1] + kal - 1) // kal * kal
            out["eps_w_in"] = torch.nn.functional.pad(e, (0, kpad - e.shape[1])).contiguous()
            return out
        if e is not None:            # [Cout, taps * Cin] -> [Cout, taps * Cin_pad]
            e3 = e.view(e.shape[0], -1, cin)
            out["eps_w_in"] = torch.nn.functional.pad(e3, (0, cp - cin)).reshape(e.shape[0], -1).contiguous()
        si = dbg.get("sign_in")
        if si is not None:           # [..., Cin] -> [..., Cin_pad]
            out["sign_in"] = torch.nn.functional.pad(si, (0, cp - cin), value=1.0).contiguous()
        return out

    def _geometry(self, x, n_samples):
        nd = self._nd
        if x.dim() != nd + 2:
            raise RuntimeError(f"Expected {nd + 2}D input to conv{nd}d, but got input of size: {list(x.shape)}")
        if x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input to have {self.in_channels} channels, but got {x.shape[1]}")
        if isinstance(self.padding, str):
            raise ValueError("string padding modes are not supported by the B200 conv kernels")
        ks = tuple(self._mu_rho()[0].shape[2:])
        st, pd, dl = _tuple(self.stride, nd), _tuple(self.padding, nd), _tuple(self.dilation, nd)
        perm = (0, *range(2, nd + 2), 1)
        cp = self._padded_cin()
        nb = x.shape[0]
        insp0 = tuple(x.shape[2:])
        outsp0 = tuple((insp0[i] + 2 * pd[i] - dl[i] * (ks[i] - 1) - 1) // st[i] + 1 for i in range(nd))
        # Few-channel 2-D reparameterization convs (the RGB stem): materialise im2col(x) once -- [B*OH*OW, taps*Cin]
        # rows are contiguous, so the kernel runs it as a linear layer with fully coalesced 128-byte row reads and
        # taps*Cin (not taps*8) columns; in MC inference x is shared by all samples, so this is done once per step.
        # Falls back to zero-channel padding when the im2col matrix would be large.
        self._bt_pmode = None
        if cp is not None:
            self._bt_pmode = "pad"
            taps = 1
            for k in ks:
                taps *= k
            # bf16 activations: pad the materialised rows to whole 64-column (128-byte) slabs, which makes the layer
            # eligible for the direct kernel (bt_direct.cuh); the extra columns are zero in x and W (mu 0, sigma 0)
            self._bt_kalign = kal = 64 if x.dtype == torch.bfloat16 else 8
            kpad = (taps * self.in_channels + kal - 1) // kal * kal
            rows = nb
            for o in outsp0:
                rows *= max(o, 0)
            if nd == 2 and self._family == "reparam" and all(o >= 1 for o in outsp0) and \
                    rows * kpad * x.element_size() <= (256 << 20):
                self._bt_pmode = "im2col"
        if self._bt_pmode == "im2col":
            # ONE launch (csrc/bt_im2col.cu): rows [B*OH*OW, kpad], column = (kh, kw, c), zero-extended to kpad
            ktrue = ks[0] * ks[1] * self.in_channels
            kpad = (ktrue + kal - 1) // kal * kal
            xp = _native.im2col2d(x, ks, st, pd, dl, kpad)
        else:
            xp = x.permute(perm)
            if not xp.is_contiguous():
                xp = xp.contiguous()
            if cp is not None:
                xp = torch.nn.functional.pad(xp, (0, cp - self.in_channels))
        shared = 0
        if n_samples > 1:
            if _mc.batch is not None and nb == _mc.batch:
                shared = 1
                batch = nb
            else:
                if nb % n_samples:
                    raise RuntimeError(f"MC context: batch {nb} not divisible by {n_samples} samples")
                batch = nb // n_samples
        else:
            batch = nb
        insp = tuple(x.shape[2:])
        outsp = tuple((insp[i] + 2 * pd[i] - dl[i] * (ks[i] - 1) - 1) // st[i] + 1 for i in range(nd))
        if any(o < 1 for o in outsp):
            raise RuntimeError(f"Calculated padded input size per channel: {insp}. Kernel size: {ks}. "
                               "Kernel size can't be greater than actual input size")
        g = _native.BtLayerGeom()
        g.n_samples, g.x_shared, g.batch = n_samples, shared, batch
        g.c_in, g.c_out, g.groups = (cp or self.in_channels), self.out_channels, self.groups
        off = 3 - nd
        for i in range(3):
            g.in_dhw[i] = g.out_dhw[i] = g.k_dhw[i] = g.stride[i] = g.dil[i] = 1
            g.pad[i] = 0
        if self._bt_pmode == "im2col":          # a linear layer over the materialised rows
            rows_per_img = 1
            for o in outsp:
                rows_per_img *= o
            g.batch = batch * rows_per_img
            g.c_in = xp.shape[1]
        for i in range(nd if self._bt_pmode != "im2col" else 0):
            g.in_dhw[off + i], g.out_dhw[off + i], g.k_dhw[off + i] = insp[i], outsp[i], ks[i]
            g.stride[off + i], g.pad[off + i], g.dil[off + i] = st[i], pd[i], dl[i]
        out_shape = (batch * n_samples, *outsp, self.out_channels)
        self._bt_outsp = outsp
        inv = (0, nd + 1, *range(1, nd + 1))

        def to_logical(o):
            return o.permute(inv)

        return xp, g, out_shape, to_logical


class BayesConvTransposeBase(BayesConvBase):
    """ConvTranspose{1,2,3}d (reference conv_variational.py:577-1094, conv_flipout.py:640-1228).

    Parameters keep the reference's shape [C_in, C_out/groups, *k] (state_dict compatible).  The kernel runs the
    transposed convolution as an implicit GEMM with FRACTIONALLY-STRIDED gather (out[o] += x[i] * W[., k] for
    o = i*stride - pad + k*dil, i.e. i = (o + pad - k*dil) / stride when that division is exact -- BtLayerGeom.
    transposed) on a cached repack of (mu, rho) to the kernel's weight matrix [C_out, *k, C_in/groups].  eps counters
    live in that matrix; materialize_eps() permutes them back to the parameter's shape.  The prior-sigma buffer has the
    parameter's shape (the reference's Flipout variants allocate it with the wrong shape, conv_flipout.py:706-709, which
    only works because it is filled with a constant)."""

    _transposed = True

    def _init_conv_transpose(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, output_padding,
                             prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias, validate):
        nd = self._nd
        if validate:
            if in_channels % groups != 0:
                raise ValueError('invalid in_channels size')
            if out_channels % groups != 0:
                raise ValueError('invalid in_channels size')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.output_padding = output_padding
        self.dilation = dilation
        self.groups = groups
        self.prior_mean = prior_mean
        self.prior_variance = prior_variance
        self.bias = bias
        ks = _tuple(kernel_size, nd)
        self._bt_pad_cache = None
        self._bt_tr_cache = None
        self._register("kernel", (in_channels, out_channels // groups, *ks), out_channels, bias,
                       posterior_mu_init, posterior_rho_init)

    @property
    def _out_pad(self):
        return _tuple(self.output_padding, self._nd)

    def _padded_cin(self):
        return None

    def _phys_params(self):                      # the parameters stay as they are; the kernel reads a repack
        return self._mu_rho()

    def _to_kernel_matrix(self, t):
        """[C_in, C_out/g, *k] -> [C_out, *k, C_in/g] (dense): Wk[g*N + n, tap, ci] = W[g*Cin_g + ci, n, tap]"""
        G, nd = self.groups, self._nd
        cin_g, n = self.in_channels // G, self.out_channels // G
        v = t.reshape(G, cin_g, n, *t.shape[2:])
        return v.permute(0, 2, *range(3, 3 + nd), 1).reshape(self.out_channels, *t.shape[2:], cin_g).contiguous()

    def _from_kernel_matrix(self, m):
        """inverse of _to_kernel_matrix for a [C_out, C_in/g, *k] tensor in the export's logical order"""
        G, nd = self.groups, self._nd
        cin_g, n = self.in_channels // G, self.out_channels // G
        v = m.reshape(G, n, cin_g, *m.shape[2:])
        return v.permute(0, 2, 1, *range(3, 3 + nd)).reshape(self.in_channels, n, *m.shape[2:]).contiguous()

    def _kernel_params(self, pmode=None):
        mu_w, rho_w = self._mu_rho()
        key = (mu_w._version, rho_w._version, mu_w.data_ptr(), rho_w.data_ptr(), mu_w.dtype, mu_w.device)
        if self._bt_tr_cache is None or self._bt_tr_cache[0] != key:
            self._bt_tr_cache = (key, self._to_kernel_matrix(mu_w.data), self._to_kernel_matrix(rho_w.data))
        return self._bt_tr_cache[1], self._bt_tr_cache[2]

    def _sigma_of(self, rho_k, pmode):
        key = (rho_k.data_ptr(), self._mu_rho()[1]._version, rho_k.dtype, rho_k.device)
        c = getattr(self, "_bt_sigma_cache", None)
        if c is None or c[0] != key:
            self._bt_sigma_cache = c = (key, torch.nn.functional.softplus(rho_k.float()).to(rho_k.dtype).contiguous())
        return c[1]

    def _geometry(self, x, n_samples):
        nd = self._nd
        if x.dim() != nd + 2:
            raise RuntimeError(f"Expected {nd + 2}D input to conv_transpose{nd}d, but got input of size: {list(x.shape)}")
        if x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input to have {self.in_channels} channels, but got {x.shape[1]}")
        ks = tuple(self._mu_rho()[0].shape[2:])
        st, pd, dl, op = _tuple(self.stride, nd), _tuple(self.padding, nd), _tuple(self.dilation, nd), self._out_pad
        for i in range(nd):
            if op[i] >= max(st[i], dl[i]):
                raise RuntimeError("output padding must be smaller than either stride or dilation")
        self._bt_pmode = None
        perm = (0, *range(2, nd + 2), 1)
        xp = x.permute(perm)
        if not xp.is_contiguous():
            xp = xp.contiguous()
        nb = x.shape[0]
        shared = 0
        if n_samples > 1:
            if _mc.batch is not None and nb == _mc.batch:
                shared, batch = 1, nb
            else:
                if nb % n_samples:
                    raise RuntimeError(f"MC context: batch {nb} not divisible by {n_samples} samples")
                batch = nb // n_samples
        else:
            batch = nb
        insp = tuple(x.shape[2:])
        outsp = tuple((insp[i] - 1) * st[i] - 2 * pd[i] + dl[i] * (ks[i] - 1) + op[i] + 1 for i in range(nd))
        if any(o < 1 for o in outsp):
            raise RuntimeError(f"conv_transpose{nd}d: computed output size {outsp} is too small")
        g = _native.BtLayerGeom()
        g.n_samples, g.x_shared, g.batch = n_samples, shared, batch
        g.c_in, g.c_out, g.groups = self.in_channels, self.out_channels, self.groups
        g.transposed = 1
        off = 3 - nd
        for i in range(3):
            g.in_dhw[i] = g.out_dhw[i] = g.k_dhw[i] = g.stride[i] = g.dil[i] = 1
            g.pad[i] = 0
        for i in range(nd):
            g.in_dhw[off + i], g.out_dhw[off + i], g.k_dhw[off + i] = insp[i], outsp[i], ks[i]
            g.stride[off + i], g.pad[off + i], g.dil[off + i] = st[i], pd[i], dl[i]
        out_shape = (batch * n_samples, *outsp, self.out_channels)
        self._bt_outsp = outsp
        inv = (0, nd + 1, *range(1, nd + 1))
        return xp, g, out_shape, (lambda o: o.permute(inv))

    def materialize_eps(self, sample=0):
        if self._bt_last is None:
            raise RuntimeError("materialize_eps() needs a previous forward()")
        last = self._bt_last
        mu_w, _ = self._mu_rho()
        G = self.groups
        cin_g = self.in_channels // G
        ks = tuple(mu_w.shape[2:])
        taps = 1
        for s_ in ks:
            taps *= s_
        kk = cin_g * taps
        e = torch.empty((self.out_channels, cin_g, *ks), dtype=torch.float32, device=mu_w.device)
        _native.rng_export(0, e, self.out_channels, kk, taps, kk, last["seed"], last["layer_key"], last["sample0"] + sample)
        eps_w = self._from_kernel_matrix(e)
        self.eps_kernel = eps_w.to(mu_w.dtype)
        eps_b = None
        if self.mu_bias is not None:
            eps_b = torch.empty(self.out_channels, dtype=torch.float32, device=mu_w.device)
            _native.rng_export(1, eps_b, self.out_channels, 1, 1, 1, last["seed"], last["layer_key"], last["sample0"] + sample)
            self.eps_bias = eps_b.to(mu_w.dtype)
        return eps_w, eps_b

    def kernel_eps_layout(self, eps_ref):
        """reference-shaped eps [C_in, C_out/g, *k] -> the PHYSICAL [C_out, taps * C_in/g] matrix the debug hook takes"""
        return self._to_kernel_matrix(eps_ref).reshape(self.out_channels, -1).contiguous()
