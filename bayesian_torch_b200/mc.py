"""Monte-Carlo inference driver: N weight samples per input batch, sharded over GPUs.

Replaces the sequential loop of the reference's evaluate()
(/root/reference/bayesian_torch/examples/main_bayesian_cifar_dnn2bnn.py:541-557:
`for mc_run in range(num_monte_carlo): model.forward(data)` -> stack -> softmax -> mean):

  * the MC-sample index is a GRID DIMENSION of every fused layer kernel: a chunk of S samples is
    evaluated in one pass through the model with the samples stacked along the batch dimension
    (eval-mode BatchNorm / ReLU / pooling are per-image, so stacking is exact);
  * softmax + running sums of p and p^2 are one kernel (csrc/bt_mc.cu), no per-sample D2H copy;
  * multi-GPU: one process per GPU, rank r evaluates a contiguous block of the N global sample
    indices (Philox counters use the GLOBAL index, so the result does not depend on the number of
    ranks up to fp32 summation order) and ONE all-reduce of the [2, B, C] fp32 moment buffer
    crosses NVLink.  Nothing else is communicated.
"""
import torch
import torch.distributed as dist

from . import _native
from ._core import BayesLayerBase, mc_sample_context


def shard_samples(n_samples, world_size, rank):
    """Contiguous block [start, start+count) of the global MC sample indices owned by `rank`."""
    if n_samples < 1 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad sharding request n={n_samples} world={world_size} rank={rank}")
    base, rem = divmod(n_samples, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def all_reduce_moments(sums, group=None):
    """The ONE collective of the path: sum the fp32 moment buffer ([2, B, C], plus [B] per-sample entropy sums when
    uncertainties are requested -- one flat tensor) over ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def _forward_moments(model, x, start, count, chunk, sample_offset, with_entropy, sample_word=None):
    """This rank's share of the work: `count` samples from global index sample_offset + start, `chunk` per pass ->
    the flat moment buffer (and its [2,B,C] / [B] views).  With `sample_word` (int32 CUDA tensor [1]) the offset is NOT
    baked into the launches: the kernels add *sample_word at run time (CUDA-graph replays with fresh draws)."""
    batch = x.shape[0]
    sums = ent = buf = None
    done = 0
    base = start if sample_word is not None else sample_offset + start
    while done < count:
        s = min(chunk, count - done)
        with mc_sample_context(s, batch, base + done, sample_word, sample_offset if sample_word is not None else 0):
            logits = model(x)
        if logits.dim() != 2 or logits.shape[0] != s * batch:
            raise RuntimeError(f"mc_predict expects logits [S*B, C]; got {tuple(logits.shape)} for S={s}, B={batch} "
                               "(is the first layer of the model a bayesian_torch_b200 layer?)")
        if not logits.is_contiguous():
            logits = logits.contiguous()
        if sums is None:
            buf, sums, ent = _moment_buffer(batch, logits.shape[1], x.device, with_entropy, zero=False)
        _native.mc_accumulate(logits, s, batch, sums, accumulate=done > 0, entropy_sum=ent)
        done += s
    if sums is None:  # a rank with no samples (n_samples < world): contributes zeros
        with mc_sample_context(1, batch, 0):
            n_classes = model(x).shape[1]
        buf, sums, ent = _moment_buffer(batch, n_classes, x.device, with_entropy, zero=True)
    return buf, sums, ent


class _MCGraph:
    """The whole per-rank MC pass (every layer launch, pooling, softmax/moments) captured ONCE as a CUDA graph and
    replayed per input batch: ~70 kernel launches through python + ctypes cost ~1.8 ms of host time per step, which
    is as long as the GPU needs at N = 1 and 2-3x longer than a rank needs at N = 4-8 GPUs (tools/small_s_check.py).
    The Philox seed and the rank's sample range are baked in, like the tensor shapes; the SAMPLE OFFSET is not: every
    layer kernel adds a device word (BtLayerGeom.sample_offset) that run() rewrites before the replay, so each replay
    can draw fresh eps -- the reference draws new eps on every forward (conv_variational.py:362).  The collective and the
    finalize stay outside the graph."""

    def __init__(self, model, x, start, count, chunk, sample_offset, with_entropy):
        self.static_x = x.clone()
        self.sample_word = torch.full((1,), int(sample_offset) & 0x7FFFFFFF, dtype=torch.int32, device=x.device)
        self.param_versions = _param_versions(model)
        args = (model, self.static_x, start, count, chunk, int(sample_offset) & 0x7FFFFFFF, with_entropy, self.sample_word)
        side = torch.cuda.Stream(device=x.device)           # warm-up off the capture stream (fills every host cache)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            _forward_moments(*args)
        torch.cuda.current_stream(x.device).wait_stream(side)
        l0 = _native.launch_count
        self.graph = torch.cuda.CUDAGraph()
        prev = _native.set_pointer_checks(False)             # (pointer-attribute queries are illegal during capture;
        try:                                                 #  the identical arguments were just checked eagerly)
            # thread-local capture mode: only THIS thread is held to the capture rules, so the NCCL watchdog thread
            # of a multi-GPU job may keep polling its events while the pass is being captured
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.buf, self.sums, self.ent = _forward_moments(*args)
        finally:
            _native.set_pointer_checks(prev)
        self.n_launches = _native.launch_count - l0          # libbtb200 kernels per replay
        # python does not run at replay time: keep the layers' "last draw" bookkeeping (materialize_eps) in step with
        # the offset each replay uses
        self.capture_offset = int(sample_offset) & 0x7FFFFFFF
        self.layers = [m for m in model.modules() if isinstance(m, BayesLayerBase) and m._bt_last is not None]
        self.layer_base = [(m._bt_last["sample0"] - self.capture_offset) & 0xFFFFFFFF for m in self.layers]

    def run(self, x, sample_offset):
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x)
        self.sample_word.fill_(int(sample_offset) & 0x7FFFFFFF)
        self.graph.replay()
        for m, base in zip(self.layers, self.layer_base):
            m._bt_last["sample0"] = (base + (int(sample_offset) & 0x7FFFFFFF)) & 0xFFFFFFFF
        _native.launch_count += self.n_launches
        return self.buf, self.sums, self.ent


_graphs = {}


def _param_versions(model):
    return [p._version for p in model.parameters()]


def _graph_key(model, x, start, count, chunk, with_entropy):
    from ._core import current_seed
    return (id(model), tuple(x.shape), tuple(x.stride()), x.dtype, x.device, start, count, chunk, with_entropy,
            current_seed())


def drop_graphs(model=None):
    """forget the captured CUDA graphs (of `model`, or all): they bake parameter pointers and cached transforms in"""
    for k in list(_graphs):
        if model is None or k[0] == id(model):
            _graphs.pop(k)


@torch.no_grad()
def mc_predict(model, x, n_samples, chunk=None, group=None, sample_offset=0, return_var=True,
               return_uncertainty=False, use_graph=False, fresh=False):
    """Predictive mean (and variance) of softmax(model(x)) over `n_samples` weight samples.

    x: [B, ...] CUDA tensor; returns (mean [B, C], var [B, C] or None), fp32, identical on all ranks.
    chunk: samples evaluated per pass (default: all samples of this rank in one pass).
    return_uncertainty: also return (predictive_entropy [B], mutual_information [B]) of the MC ensemble -- the
    reference's utils/util.py:45-60 on the stacked per-sample probabilities -- computed on device from running sums
    that ride the same single all-reduce; result = (mean, var, pred_entropy, mutual_info).
    use_graph: capture this rank's pass as a CUDA graph on first use (per model / shape / seed / sample range /
    parameter version) and replay it afterwards; the results are the same tensors' worth of numbers, the host cost per
    call drops from ~1.8 ms to a replay.  `sample_offset` is NOT part of the graph: it is written to a device word the
    kernels read, so replays with different offsets draw different eps and each equals the eager run with that offset.
    fresh: draw NEW weight samples on every call, like the reference's evaluate() (every forward draws new eps,
    conv_variational.py:362): the call uses sample_offset + (number of samples this model has drawn so far) and
    advances that counter by n_samples (identically on every rank).
    """
    if model.training:
        raise RuntimeError("mc_predict stacks MC samples along the batch dimension; call model.eval() first "
                           "(train-mode BatchNorm would mix statistics across samples)")
    _native.require_cuda(x, "input")
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    start, count = shard_samples(n_samples, world, rank)
    batch = x.shape[0]
    if chunk is None or chunk > count:
        chunk = max(count, 1)
    if fresh:
        drawn = getattr(model, "_bt_mc_drawn", 0)
        sample_offset = (sample_offset + drawn) & 0x7FFFFFFF
        model._bt_mc_drawn = drawn + n_samples
    if getattr(model, "_bt_fused_inference", False):
        from .fuse import refresh_epilogues
        refresh_epilogues(model)              # folded BatchNorm statistics changed (load_state_dict)? re-fold
    if use_graph:
        key = _graph_key(model, x, start, count, chunk, return_uncertainty)
        g = _graphs.get(key)
        if g is not None and g is not False and g.param_versions != _param_versions(model):
            g = None                                         # a parameter was updated in place: re-capture
        if g is None:
            if len(_graphs) >= 8:                            # bounded: graphs pin their activation pools
                _graphs.pop(next(iter(_graphs)))
            try:
                g = _MCGraph(model, x, start, count, chunk, sample_offset, return_uncertainty)
            except RuntimeError as e:                        # capture refused (e.g. an op that synchronises): stay eager
                import warnings
                warnings.warn(f"mc_predict: CUDA-graph capture failed, running eagerly ({e})", RuntimeWarning)
                g = False
            _graphs[key] = g
        if g is False:
            buf, sums, ent = _forward_moments(model, x, start, count, chunk, sample_offset, return_uncertainty)
        else:
            buf, sums, ent = g.run(x, sample_offset)
    else:
        buf, sums, ent = _forward_moments(model, x, start, count, chunk, sample_offset, return_uncertainty)
    all_reduce_moments(buf, group)
    mean = torch.empty(sums.shape[1:], dtype=torch.float32, device=x.device)
    var = torch.empty_like(mean) if return_var else None
    _native.mc_finalize(sums, n_samples, mean, var)
    if not return_uncertainty:
        return mean, var
    pred_entropy = torch.empty(batch, dtype=torch.float32, device=x.device)
    mutual_info = torch.empty_like(pred_entropy)
    _native.mc_uncertainty(sums, ent, n_samples, pred_entropy, mutual_info)
    return mean, var, pred_entropy, mutual_info


def _moment_buffer(batch, n_classes, device, with_entropy, zero):
    """One flat fp32 tensor: [2, B, C] sums of p and p^2, then (optionally) [B] sums of the per-sample entropies."""
    n = 2 * batch * n_classes
    make = torch.zeros if zero else torch.empty
    buf = make(n + (batch if with_entropy else 0), dtype=torch.float32, device=device)
    return buf, buf[:n].view(2, batch, n_classes), (buf[n:] if with_entropy else None)


def count_bayes_layers(model):
    return sum(1 for m in model.modules() if isinstance(m, BayesLayerBase))
