"""Data-parallel training check (run under torchrun on >= 2 GPUs):
rank-seeded minibatches, forward + backward on the CUDA training kernels, ONE NCCL allreduce of the flat gradient per step,
torch Adam on the flat views.  Verifies that (1) the averaged gradient equals the mean of the per-rank gradients, (2) the
parameters stay bit-identical across ranks, (3) the loss goes down; prints hyperedges/s."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from higashi_b200 import batches, parallel, synthetic as S  # noqa: E402
from higashi_b200.train_engine import TrainEngine  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ds = S.make_dataset("C1", n_cells=200, bins=[250, 120], seed=3)        # identical data + weights on every rank
    sd, feats = S.random_state_dict(ds, 64, seed=1)
    eng = TrainEngine(sd, ds.num, feats, ds.cell_feats, device=local)
    tr = parallel.DataParallelTrainer(eng, lr=1e-3)
    rng = np.random.default_rng(100 + rank)                                # rank-seeded sampler
    fixed = batches.synthetic_batch(ds.csr, 0, ds.num_list, 256, 5, np.random.default_rng(7))
    fixed_nb = batches.neighbor_csr(ds.csr, fixed[0], 0, ds.num_list)
    _, loss0 = eng.forward(fixed[0], 0, fixed_nb, y=fixed[1], w=fixed[2])
    # (1) gradient averaging
    x, y, w = batches.synthetic_batch(ds.csr, 0, ds.num_list, 128, 5, rng)
    nb = batches.neighbor_csr(ds.csr, x, 0, ds.num_list)
    eng.zero_grad(); eng.forward(x, 0, nb, y=y, w=w); eng.backward()
    local_g = eng.grads.clone()
    if world > 1:
        gathered = [torch.empty_like(local_g) for _ in range(world)]
        dist.all_gather(gathered, local_g)
        eng.allreduce_grads()
        ref = torch.stack(gathered).mean(0)
        assert torch.allclose(eng.grads, ref, rtol=1e-5, atol=1e-8), float((eng.grads - ref).abs().max())
    # (2) + (3) training steps
    steps, n_h = 60, 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for it in range(steps):
        chrom = it % 2
        x, y, w = batches.synthetic_batch(ds.csr, chrom, ds.num_list, 256, 5, rng)
        nb = batches.neighbor_csr(ds.csr, x, chrom, ds.num_list, training=True, rng=rng)
        tr.step(x, chrom, nb, y, w)
        n_h += len(x)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if world > 1:
        chk = eng.params.double().sum().reshape(1)
        both = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        assert all(float(b) == float(both[0]) for b in both), [float(b) for b in both]
    _, loss1 = eng.forward(fixed[0], 0, fixed_nb, y=fixed[1], w=fixed[2])
    assert loss1 < loss0, (loss0, loss1)
    if rank == 0:
        print("DP_TRAIN_OK world=%d loss %.4f -> %.4f  %.0f hyperedges/s (incl. host batch producer)" % (world, loss0, loss1, world * n_h / dt))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
