"""Multi-GPU plumbing of the imputation path: shard by cell range, one process per GPU, no data-path collective.

Reference: Higashi.impute_no_nbr / impute_with_nbr (Higashi_wrapper.py:1457-1475, 1578-1599) split the cells with
`np.array_split(arange(n_cells), gpu_num - 1)`, run `python Impute.py ... cell_start cell_end ... gpu_id` per shard and
merge the part files with `linkhdf5` (utils.py:217-253), which also divides every cell vector by its mean.
Here the split is over ALL `world` GPUs (no idle GPU), the workers are `python -m higashi_b200.Impute` processes (or the
ranks of a torchrun job), and `link_parts` is the merge."""
import os
import subprocess
import sys

import numpy as np

from . import sink as S


def cell_shards(n_cells: int, world: int):
    """Contiguous cell ranges [(start, end)] as np.array_split would produce them."""
    parts = np.array_split(np.arange(n_cells), world)
    return [(int(p[0]), int(p[-1]) + 1) if len(p) else (0, 0) for p in parts]


def max_over_ranks(value: float) -> float:
    """Device-time reduction used by bench.py: MAX over ranks (gloo on CPU tests, nccl on GPUs)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def link_parts(name, shards, temp_dir, impute_list, normalise=True):
    """linkhdf5 (utils.py:217-253): merge `{chrom}_{name}_part_{i}` containers into `{chrom}_{name}`; every cell
    vector is divided by its mean (utils.py:234) unless normalise=False; part files are removed."""
    for chrom in impute_list:
        out = S.ChromSink(temp_dir, chrom, name)
        for i, (c0, c1) in enumerate(shards):
            # one dataset of the part file in memory at a time (the reference reads one `cell_%d` at a time too)
            seen = set()
            for key, v in S.iter_container(temp_dir, chrom, "%s_part_%d" % (name, i)):
                if key == "coordinates":
                    if i == 0:
                        out.create_dataset("coordinates", v)
                    continue
                cell = int(key[5:])
                if not (c0 <= cell < c1):
                    continue
                if normalise:
                    v = v / np.mean(v)
                out.create_dataset(key, v.astype(np.float32), compress=True)
                seen.add(cell)
            if len(seen) != c1 - c0:
                raise RuntimeError("part %d of %s misses cells: %s" % (i, chrom, sorted(set(range(c0, c1)) - seen)[:8]))
        out.close()
        for i in range(len(shards)):
            for ext in (".hdf5", ".npz"):
                p = os.path.join(temp_dir, "%s_%s_part_%d%s" % (chrom, name, i, ext))
                if os.path.exists(p):
                    os.remove(p)


def impute_multi_gpu(config_path, model_path, name, mode, n_cells, sparse_path, weighted_info=None, gpus=None,
                     impute_list=None, temp_dir=None, normalise=True):
    """One `python -m higashi_b200.Impute` worker per GPU over contiguous cell shards, then link_parts."""
    import torch
    gpus = list(range(torch.cuda.device_count())) if gpus is None else list(gpus)
    shards = cell_shards(n_cells, len(gpus))
    procs = []
    for i, ((c0, c1), g) in enumerate(zip(shards, gpus)):
        cmd = [sys.executable, "-m", "higashi_b200.Impute", config_path, model_path, "%s_part_%d" % (name, i), mode, str(c0), str(c1),
               sparse_path, weighted_info if weighted_info is not None else "None", str(g)]
        procs.append(subprocess.Popen(cmd))
    rcs = [p.wait() for p in procs]
    if any(rcs):
        raise RuntimeError("imputation worker failed: return codes %s" % rcs)
    if impute_list is not None and temp_dir is not None:
        link_parts(name, shards, temp_dir, impute_list, normalise)
    return shards


# ----------------------------------------------------------------------------------------------------- training (DP)
def reduce_gradients(flat_grads, presence):
    """The single exchange step of data-parallel training.

    flat_grads: 1-D tensor (fixed layout, zeros where a rank produced no gradient); presence: 1-D tensor with one entry
    per parameter tensor (1 = this rank produced a gradient).  Every rank gets the MEAN gradient over the ranks and the
    OR of the presence flags -- parameters no rank touched keep `grad = None`, exactly like the reference's single-GPU
    run where e.g. only the batch chromosome's projection receives a gradient (torch Adam skips grad=None).

    ONE collective: when `presence` is the tail of the storage `flat_grads` starts (TrainEngine allocates them that way) the
    two are summed in a single all_reduce; otherwise they are packed into one buffer first."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat_grads, presence
    world = dist.get_world_size()
    n, m = flat_grads.numel(), presence.numel()
    contiguous_pair = (flat_grads.dtype == presence.dtype and flat_grads.is_contiguous() and presence.is_contiguous() and
                       presence.data_ptr() == flat_grads.data_ptr() + n * flat_grads.element_size() and
                       flat_grads.untyped_storage().data_ptr() == presence.untyped_storage().data_ptr())
    if contiguous_pair:
        both = torch.as_strided(flat_grads, (n + m,), (1,))
        dist.all_reduce(both, op=dist.ReduceOp.SUM)              # NCCL over NVLink on the GPUs, gloo in the CPU tests
    else:
        both = torch.cat([flat_grads, presence.to(flat_grads.dtype)])
        dist.all_reduce(both, op=dist.ReduceOp.SUM)
        flat_grads.copy_(both[:n])
        presence.copy_(both[n:])
    flat_grads.div_(world)
    presence.clamp_(max=1.0)                                     # sum of 0/1 flags -> OR
    return flat_grads, presence


class DataParallelTrainer:
    """train_epoch body (Higashi_wrapper.py:916-1010) for stage 2/3 on the CUDA training kernels, data-parallel:
    every rank draws its own minibatch, forward + backward locally, ONE allreduce of the flat gradient, identical Adam
    step on every rank."""

    def __init__(self, engine, lr=1e-3):
        import torch
        self.eng = engine
        self.names = list(engine.layout.keys())
        self.params = [torch.nn.Parameter(v) for v in engine.named_params().values()]
        self.grad_views = list(engine.named_grads().values())
        # one fused update kernel over all parameter tensors instead of torch's default foreach implementation (7 launches / step)
        self.opt = torch.optim.Adam(self.params, lr=lr, fused=bool(self.params and self.params[0].is_cuda))
        # presence flags live right behind the flat gradient (same storage): gradients + flags travel in ONE all_reduce
        self.presence = engine.presence if getattr(engine, "presence", None) is not None else \
            torch.zeros(len(self.names), dtype=torch.float32, device=engine.grads.device)
        self._patterns = {}
        self.loss_mode_cached = 0
        self.world = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.world = dist.get_world_size()
        except Exception:
            pass

    def _presence_pattern(self, chrom, stage):
        """Which parameter tensors receive a gradient for a batch of chromosome `chrom` (host list, cached)."""
        key = (chrom, stage, self.loss_mode_cached)
        if key not in self._patterns:
            from .train_engine import grad_presence
            pat = grad_presence(self.names, chrom, stage, self.loss_mode_cached == 2, self.eng.cell_on_hook)
            import torch
            self._patterns[key] = (pat, torch.tensor([1.0 if v else 0.0 for v in pat], dtype=torch.float32, device=self.presence.device))
        return self._patterns[key]

    def step(self, x, chrom, to_neighs, y, w, stage=2, loss_mode=0, train_mode=True, alpha=1.0, beta=1e-2,
             median_sparsity=0.0, contractive_weight=0.0, want_loss=True, sampled=None, uniform_batches=False):
        """One update.  stage 1 reproduces train_epoch's three backward passes (Higashi_wrapper.py:970-1006):
        gradient of the main loss, gradient of the reconstruction loss, ratio1 = max(beta*|g_main|/|g_recon|,
        100*median_sparsity - 3), total = alpha*main + ratio1*recon + contractive.
        uniform_batches=True: the caller guarantees that every rank runs the same (chrom, stage, loss_mode) this step (e.g. a
        shared chromosome schedule), so the set of parameters with a gradient is known locally and the step needs no
        device-to-host read of the reduced presence flags."""
        self.n_steps = getattr(self, "n_steps", 0) + 1
        self.eng.set_dropout(train_mode, seed=0x5EED0000 + self.n_steps)
        self.eng.zero_grad()
        _, loss = self.eng.forward(x, chrom, to_neighs, y=y, w=w, stage=stage, loss_mode=loss_mode, want_outputs=want_loss,
                                   want_mean=False, sampled=sampled)
        self.eng.backward()
        if alpha != 1.0:
            self.eng.grads.mul_(alpha)
        if stage == 1 and self.eng.cell_on_hook:
            pfx = self.eng.embed_prefix
            w0 = "%s.wstack.0.weight_list.0" % pfx
            main_norm = float(self.eng.named_grads()[w0].norm(2))
            self.recon_loss, recon_norm = self.eng.recon_loss()
            ratio1 = max(beta * main_norm / max(recon_norm, 1e-30), 100.0 * median_sparsity - 3.0)
            self.eng.recon_apply(ratio1)
            if contractive_weight > 0:
                for name in (w0, "%s.wstack.0.reverse_weight_list.0" % pfx):
                    self.eng.named_grads()[name].add_(self.eng.named_params()[name], alpha=2.0 * contractive_weight)
        self.loss_mode_cached = loss_mode
        pat, pat_dev = self._presence_pattern(chrom, stage)
        if self.world > 1:
            self.presence.copy_(pat_dev)
            reduce_gradients(self.eng.grads, self.presence)
            if not uniform_batches:
                pat = (self.presence > 0).tolist()        # OR over the ranks (they may have drawn different chromosomes)
        for p, g, on in zip(self.params, self.grad_views, pat):
            p.grad = g if on else None
        self.opt.step()
        return loss
