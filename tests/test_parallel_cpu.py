"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo ranks shard the cells, write part containers,
reduce the max time, and the merge reproduces linkhdf5 (utils.py:217-253)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from higashi_b200 import parallel, sink


def test_cell_shards_match_array_split():
    for n, w in [(620, 8), (7, 2), (6000, 3), (5, 8)]:
        sh = parallel.cell_shards(n, w)
        ref = np.array_split(np.arange(n), w)
        assert len(sh) == w
        for (a, b), r in zip(sh, ref):
            assert b - a == len(r)
            if len(r):
                assert a == r[0] and b == r[-1] + 1
        assert sum(b - a for a, b in sh) == n


def _worker(rank, world, tmp, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_cells, P = 11, 7
    c0, c1 = parallel.cell_shards(n_cells, world)[rank]
    sk = sink.ChromSink(tmp, "chr1", "emb_nbr_0_impute_part_%d" % rank)
    sk.create_dataset("coordinates", np.arange(2 * P, dtype=np.int64).reshape(P, 2))
    for cell in range(c0, c1):
        sk.create_dataset("cell_%d" % cell, (np.arange(P) + 1 + cell).astype(np.float32), compress=True)
    sk.close()
    t = parallel.max_over_ranks(10.0 + rank)            # rank 1 is the slowest
    assert t == 10.0 + world - 1
    dist.barrier()
    if rank == 0:
        parallel.link_parts("emb_nbr_0_impute", parallel.cell_shards(n_cells, world), tmp, ["chr1"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_link(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, str(tmp_path), port), nprocs=world, join=True)
    f = sink.read_container(str(tmp_path), "chr1", "emb_nbr_0_impute")
    assert f["coordinates"].shape == (7, 2)
    for cell in range(11):
        v = (np.arange(7) + 1 + cell).astype(np.float32)
        np.testing.assert_allclose(f["cell_%d" % cell], v / v.mean(), rtol=1e-6)      # utils.py:234
    assert not [p for p in os.listdir(tmp_path) if "_part_" in p]


# ----------------------------------------------------------------------------------------- training exchange step (gloo)
def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 3 parameter tensors of 4 values; rank 0 touches tensors 0 and 1, rank 1 touches 0 and 2 -> nobody skips 0,
    flat = torch.zeros(12)
    pres = torch.zeros(3)
    if rank == 0:
        flat[0:4] = 1.0; flat[4:8] = 2.0; pres[0] = 1; pres[1] = 1
    else:
        flat[0:4] = 3.0; pres[0] = 1
    parallel.reduce_gradients(flat, pres)
    # the layout TrainEngine uses: presence flags right behind the gradients in ONE storage -> a single all_reduce
    buf = torch.zeros(12 + 3)
    flat2, pres2 = buf[:12], buf[12:]
    if rank == 0:
        flat2[0:4] = 1.0; flat2[4:8] = 2.0; pres2[0] = 1; pres2[1] = 1
    else:
        flat2[0:4] = 3.0; pres2[0] = 1
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
    parallel.reduce_gradients(flat2, pres2)
    dist.all_reduce = real
    if rank == 0:
        ret["flat"] = flat.tolist(); ret["pres"] = pres.tolist()
        ret["flat2"] = flat2.tolist(); ret["pres2"] = pres2.tolist(); ret["calls"] = calls
    dist.destroy_process_group()


def test_two_rank_gradient_exchange():
    world, port = 2, 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["flat"] == [2.0] * 4 + [1.0] * 4 + [0.0] * 4         # mean over ranks, zeros where absent
    assert ret["pres"] == [1.0, 1.0, 0.0]                            # tensor 2: no rank had a gradient -> grad stays None
    assert ret["flat2"] == ret["flat"] and ret["pres2"] == ret["pres"]
    assert ret["calls"] == [15]                                      # gradients + flags in ONE collective
