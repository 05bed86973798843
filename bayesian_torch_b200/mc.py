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


@torch.no_grad()
def mc_predict(model, x, n_samples, chunk=None, group=None, sample_offset=0, return_var=True,
               return_uncertainty=False):
    """Predictive mean (and variance) of softmax(model(x)) over `n_samples` weight samples.

    x: [B, ...] CUDA tensor; returns (mean [B, C], var [B, C] or None), fp32, identical on all ranks.
    chunk: samples evaluated per pass (default: all samples of this rank in one pass).
    return_uncertainty: also return (predictive_entropy [B], mutual_information [B]) of the MC ensemble -- the
    reference's utils/util.py:45-60 on the stacked per-sample probabilities -- computed on device from running sums
    that ride the same single all-reduce; result = (mean, var, pred_entropy, mutual_info).
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
    sums = ent = buf = None
    done = 0
    while done < count:
        s = min(chunk, count - done)
        with mc_sample_context(s, batch, sample_offset + start + done):
            logits = model(x)
        if logits.dim() != 2 or logits.shape[0] != s * batch:
            raise RuntimeError(f"mc_predict expects logits [S*B, C]; got {tuple(logits.shape)} for S={s}, B={batch} "
                               "(is the first layer of the model a bayesian_torch_b200 layer?)")
        if not logits.is_contiguous():
            logits = logits.contiguous()
        if sums is None:
            buf, sums, ent = _moment_buffer(batch, logits.shape[1], x.device, return_uncertainty, zero=False)
        _native.mc_accumulate(logits, s, batch, sums, accumulate=done > 0, entropy_sum=ent)
        done += s
    if sums is None:  # a rank with no samples (n_samples < world): contributes zeros
        with mc_sample_context(1, batch, 0):
            n_classes = model(x).shape[1]
        buf, sums, ent = _moment_buffer(batch, n_classes, x.device, return_uncertainty, zero=True)
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
