"""Host-side ceiling of concurrent pinned D2H / H2D streams on one box (VERDICT r1 weak #7c).

    python scripts/pcie_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_ceiling.py

Every rank streams a 900 MB block (the size of one fp32 score block of bench.py, 1000 cells x 225 023 pairs) device -> pinned host
back to back for ~2 s, all ranks at the same time (barrier on both sides, wall time = max over ranks).  Prints ONE JSON line on
rank 0: per-rank and aggregate GB/s for D2H alone, H2D alone and D2H with a concurrent H2D of a 75 MB block (one step's own-cell
contacts).  This is the number the 8-GPU end-to-end line of bench.py cannot exceed: 8 x 3.9e9 triplets/s x 4 B = 125 GB/s of
fp32 scores, half that with the fp16 score block."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 900 * 1000 * 1000 // 4
    m = 75 * 1000 * 1000 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    h2 = torch.empty(m, dtype=torch.float32).pin_memory()
    d2 = torch.empty(m, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        fn(); barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / reps

    def d2h():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s2):
            d.copy_(h, non_blocking=True)

    def duplex():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)

    reps = 40
    out = {"n_gpus": world, "block_mb": 900}
    for name, fn, nbytes in (("d2h", d2h, n * 4), ("h2d", h2d, n * 4), ("d2h_with_75mb_h2d", duplex, n * 4)):
        dt = timed(fn, reps)
        out[name + "_gbs_per_rank"] = nbytes / dt / 1e9
        out[name + "_gbs_aggregate"] = world * nbytes / dt / 1e9
    numa = None
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        numa = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
    except Exception:
        pass
    out["numa_node_rank0"] = numa
    out["host_cpus"] = os.cpu_count()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
