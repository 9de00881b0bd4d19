#!/usr/bin/env python
"""bench.py -- imputed (cell, bin_i, bin_j) triplets / s on synthetic scHi-C data (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

Workload (N=1): BASELINE.json configs[1] -- synthetic 4DN sci-Hi-C Kim-shaped data, 6000 cells, hg19 @ 1 Mb
genome-wide (2897 bins, P = 225 023 bin pairs per cell), impute_with_nbr (k = 4), d_model = 64.
One "step" = one pass of the hot path (K1 gather + projection, K2 fused scorer, K3 scatter) over one batch
of `--cells-per-step` cells.  Inputs are larger than L2 across steps (every step touches a new cell range
and writes a fresh [cells, P] fp32 block far above 126 MB).

`value`  : whole-job triplets/s with inputs resident in HBM, device-timed (CUDA events, max over ranks).
`e2e`    : same metric through the C ABI with HOST buffers: every step uploads the stored contacts of its own cells from
           pinned host memory (hsg_refresh_contacts) and reads the scores back into pinned host memory.
`--impl reference` : the UNMODIFIED reference's own Impute.impute_process (baseline/_ref, a pip --target install of the reference
           that is git-ignored and travels to the GPU box; import shims in oracle/ref_shims.py) on all host threads, one cell
           per step of a 64-cell sample of the same shape; the oracle/ port only if that import fails (`cpu_baseline.kind`).

Multi-GPU (torchrun): imputation shards by cell range, one rank per GPU, no data-path collective; every
rank imputes its own synthetic shard (weak scaling); torch.distributed is only the barrier + max-time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torch warnings) may write to file descriptor 1
# behind Python's back, so fd 1 is pointed at stderr for the whole run and the result line goes to a private copy of the
# original stdout.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    _claim_stdout()
    _RESULT_OUT.write(json.dumps(obj) + "\n")
    _RESULT_OUT.flush()


FLOP_PER_TRIPLET_TENSOR = 110592.0      # fc1 + pff_n1 + classifier layer 1 at d = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--cells-per-step", type=int, default=1000)
    ap.add_argument("--mode", default="impute", choices=["impute", "train"],
                    help="impute (default, the BASELINE headline) or train (hyperedges/s of the data-parallel training step)")
    ap.add_argument("--precision", default=os.environ.get("HSG_PRECISION", "auto"))
    ap.add_argument("--n-cells", type=int, default=0, help="shrink the dataset (debug only; marks the run invalid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training line of the default invocation")
    ap.add_argument("--d-model", type=int, default=0, help="override d_model of the workload (64 or 128)")
    ap.add_argument("--no-c3", action="store_true", help="skip the secondary strong-scaling C3 job of the default invocation")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index, self.rows, self.proc, self.period_ms = index, [], None, int(period_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local):
    """Multi-rank runs: keep this rank's threads (and therefore its first-touch pinned staging buffers) on the NUMA node the GPU
    hangs off, so that eight ranks streaming scores to the host do not all cross the socket interconnect.  Best effort: any
    missing sysfs entry or an empty intersection with the allowed CPU set leaves the process untouched."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def build_dataset(args, rank):
    from higashi_b200 import synthetic as S
    kw = {}
    if args.n_cells:
        kw["n_cells"] = args.n_cells
    ds = S.make_dataset(args.workload, shard=rank, **kw)
    if getattr(args, "d_model", 0):
        ds.d = int(args.d_model)
    sd, feats = S.random_state_dict(ds, ds.d, seed=0)
    return ds, sd, feats


def knn_lists(ds, k, seed):
    """kNN cell graph for the synthetic shard (oracle-free: random embeddings -> exact top-k by numpy)."""
    if k <= 0:
        return None, None
    rng = np.random.Generator(np.random.PCG64(seed))
    emb = rng.standard_normal((ds.n_cells, 16)).astype(np.float32)
    emb[:, :3] += rng.integers(0, 6, size=(ds.n_cells, 1)) * 2.0
    x = emb.astype(np.float64)
    xx = (x * x).sum(-1)
    d2 = np.maximum(xx[:, None] + xx[None, :] - 2 * x @ x.T, 0)
    np.fill_diagonal(d2, 0)
    dist = np.sqrt(d2).astype(np.float32)
    part = np.argpartition(dist, k + 1, axis=-1)[:, :k + 1]
    pd = np.take_along_axis(dist, part, axis=-1)
    order = np.argsort(pd, axis=-1, kind="stable")
    nbr = np.take_along_axis(part, order, axis=-1)
    dsel = np.take_along_axis(pd, order, axis=-1)
    nbr[:, 0] = np.arange(ds.n_cells)          # self first (distance 0)
    scale = np.quantile(dsel[:, 1:].reshape(-1), 0.25)
    w = np.exp(-dsel / scale)
    w = (w / w.sum(-1, keepdims=True)).astype(np.float32)
    return (nbr + 1).astype(np.int64), w


# ------------------------------------------------------------------------------------------ reference arm
def _reference_imputer(args, n_small):
    """The reference's own modules on a shrunken shard of the workload (oracle/ref_runner.py), or None when the reference is not
    importable on this box (no baseline/_ref install): the callers then time the oracle port and say so."""
    if os.environ.get("HSG_REFERENCE_ARM", "") == "port":
        return None, None, "forced by HSG_REFERENCE_ARM=port"
    try:
        from oracle import ref_runner as RR
        if not RR.available():
            return None, None, "baseline/_ref/higashi not found"
        from higashi_b200 import synthetic as S
        ds = S.make_dataset(args.workload, shard=0, n_cells=n_small)
        nbr, w = knn_lists(ds, ds.k, 99)
        return RR.ReferenceImputer(ds, nbr, w), ds, ""
    except Exception as e:                                   # missing dependency of the reference on this box
        return None, None, "reference import failed: %r" % (e,)


def workload_string(name, n_cells, ds, P):
    """One spelling of the workload for both arms (`--impl ours` and `--impl reference`)."""
    return ("%s: synthetic %d cells x %d bins (%d chromosomes), P=%d pairs/cell, k=%d, d_model=%d, impute_with_nbr"
            % (name, n_cells, sum(ds.bins), len(ds.bins), P, ds.k, ds.d))


def _full_cells(name):
    from higashi_b200 import synthetic as S
    return int(S.CONFIGS[name]["n_cells"])


def run_reference(args):
    """`--impl reference`: the UNMODIFIED reference (Impute.impute_process from baseline/_ref through oracle/ref_shims.py) on the
    host cores; the oracle port only when the reference cannot be imported."""
    os.environ["CUDA_VISIBLE_DEVICES"] = ""          # the reference moves tensors to cuda when it sees one: this arm is its CPU path
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_small = 64
    cells_per_step = 1
    R, ds, why = _reference_imputer(args, n_small)
    one_thread = None
    if R is not None:
        kind = "reference"
        P = None
        if not args.no_cpu_baseline:
            dt, n = R.impute(0, 1)                   # Impute.py:14 pins torch to ONE thread at import: the reference's default
            one_thread = n / dt
        R.set_threads(cores)
        times = []
        for it in range(args.warmup + args.steps):
            c0 = (it * cells_per_step) % (n_small - cells_per_step + 1)
            dt, n = R.impute(c0, c0 + cells_per_step)
            P = n // cells_per_step
            if it >= args.warmup:
                times.append(dt)
        sample = ("%d cell(s) per step x %d steps of %s: the reference's Impute.impute_process (prep_one + fix_cell2 + "
                  "predict(batch 1e5) + one cell_%%d dataset per chromosome; dict-backed h5py, no gzip) imported from %s, "
                  "torch.set_num_threads(%d) after the import" % (cells_per_step, args.steps, args.workload,
                                                                   os.path.relpath(R.shims.REFERENCE_ROOT, ROOT), cores))
    else:
        from oracle import impute as OI, scorer as OS
        kind = "port"
        torch.set_num_threads(cores)
        # the oracle needs only the cells it imputes (+ neighbours): a shrunken shard of the same shape
        from higashi_b200 import synthetic as S
        ds = S.make_dataset(args.workload, shard=0, n_cells=n_small)
        sd, feats = S.random_state_dict(ds, ds.d, seed=0)
        nbr, w = knn_lists(ds, ds.k, 99)
        tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
        chroms = list(range(len(ds.bins)))
        with torch.no_grad():
            tables = OS.embed_tables(tsd, feats)
        mx = int(1e5) if ds.max_bin < 0 else ds.max_bin
        big, _, _, _ = OI.enumerate_pairs(ds.num_list, chroms, ds.min_bin, mx)
        P = len(big)
        mats = lambda c, cell: ds.csr.matrix(c, cell - 1)
        times = []
        for it in range(args.warmup + args.steps):
            cells = [(it * cells_per_step + q) % n_small for q in range(cells_per_step)]
            t0 = time.perf_counter()
            OI.impute_cells(tsd, tables, ds.cell_feats, ds.num_list, mats, cells, chroms, big, 0, nbr, w, "sigmoid")
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        sample = ("%d cell(s) per step x %d steps of %s (oracle/ torch-CPU fp32 restatement of Impute.impute_process: prep_one + "
                  "fix_cell2 + predict(batch 1e5)); the reference itself was not used: %s" % (cells_per_step, args.steps, args.workload, why))
    total = sum(times)
    val = args.steps * cells_per_step * P / total
    cb = {"value": val, "unit": "triplets/s", "cores": cores, "kind": kind, "sample": sample}
    if one_thread is not None:
        cb["value_1_thread"] = one_thread
    line = {"impl": "reference", "metric": "imputed (cell,bin_i,bin_j) triplets/s", "value": val, "unit": "triplets/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the SAME workload string as the GPU arm prints (the job is the full configuration); what was actually timed -- a bounded
            # sample of it, per-triplet cost being independent of the cell count -- is spelled out next to it and in cpu_baseline.sample
            "config": {"workload": workload_string(args.workload, _full_cells(args.workload), ds, P),
                       "cells_per_step": cells_per_step, "triplets_per_step": cells_per_step * P,
                       "sample": "%d-cell sample of the same synthetic shape, %d cell(s) per step" % (n_small, cells_per_step)},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_reference_train(args):
    """Reference arm of --mode train: the oracle's torch-CPU restatement of the stage-2 training step (forward_stage2 + BCE +
    autograd + Adam; Higashi_wrapper.py:728-759, 1006-1010), all host threads, on batches of the same shape."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import scorer as OS
    from higashi_b200 import batches, synthetic as S
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ds = S.make_dataset("C4", n_cells=args.n_cells or 300, seed=4000, density=0.02)
    sd, feats = S.random_state_dict(ds, ds.d, seed=0)
    tsd = {k: torch.from_numpy(np.array(v, dtype=np.float32, copy=True)) for k, v in sd.items()}
    with torch.no_grad():
        cell_table = OS.encoder_branch(torch.as_tensor(feats[0]), tsd, "encode1.static_nn", 0)
    tables = [cell_table] + [None] * len(ds.bins)
    params = [t.requires_grad_(True) for k, t in tsd.items() if t.dtype == torch.float32 and "attribute_dict" not in k]
    opt = torch.optim.Adam(params, lr=1e-3)
    rng = np.random.default_rng(10)
    pool = []
    for i in range(4):
        chrom = int(rng.integers(0, len(ds.bins)))
        x, y, w = batches.synthetic_batch(ds.csr, chrom, ds.num_list, 512, 5, rng)
        indptr, cols, vals = batches.neighbor_csr(ds.csr, x, chrom, ds.num_list, training=True, rng=rng)
        rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
        base = int(ds.num_list[chrom]) + 1
        nb = (np.stack([rows, cols]).astype(np.int64), vals, np.arange(base, base + ds.bins[chrom], dtype=np.int64))
        pool.append((x, nb, torch.from_numpy(y.astype(np.float32)), torch.from_numpy(w.astype(np.float32))))
    times = []
    for it in range(args.warmup + args.steps):
        x, nb, y, w = pool[it % len(pool)]
        t0 = time.perf_counter()
        opt.zero_grad()
        out = OS.forward_stage2(x, nb, ds.num_list, tsd, feats, ds.cell_feats, tables)
        mean = out[0] if isinstance(out, (tuple, list)) else out
        loss = torch.nn.functional.binary_cross_entropy_with_logits(mean.reshape(-1), y.reshape(-1), weight=w.reshape(-1))
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    B = float(np.mean([len(pool[it % len(pool)][0]) for it in range(args.warmup, args.warmup + args.steps)]))
    val = args.steps * B / total
    emit(({"impl": "reference", "metric": "train hyperedges/s", "value": val, "unit": "hyperedges/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "C4-shaped: %d cells x %d bins, stage-2 training step, %d hyperedges per step, d_model=%d"
                                             % (ds.n_cells, sum(ds.bins), B, ds.d)},
                      "cpu_baseline": {"value": val, "unit": "hyperedges/s", "cores": cores, "kind": "port",
                                       "sample": "%d steps of %d hyperedges (oracle/ forward_stage2 + torch autograd + Adam, no dropout)"
                                                 % (args.steps, B)},
                      "e2e": {"value": val, "unit": "hyperedges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def cpu_baseline(args, ds_shape):
    """Bounded CPU sample for the `cpu_baseline` object of the main line: the reference arm of this script in a child process
    (the reference's CPU path must not see the GPU this process holds), 3 steps of one cell; the in-process oracle port when the
    child fails."""
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                              "--steps", "3", "--warmup", "1"], capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
        lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
        cb = json.loads(lines[-1])["cpu_baseline"]
        if cb.get("value", 0) > 0:
            return cb
    except Exception:
        pass
    return cpu_baseline_port(args, ds_shape)


def cpu_baseline_port(args, ds_shape):
    """Bounded CPU sample (oracle port)."""
    import torch
    from higashi_b200 import synthetic as S
    from oracle import impute as OI, scorer as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_small = 48
    ds = S.make_dataset(args.workload, shard=0, n_cells=n_small)
    sd, feats = S.random_state_dict(ds, ds.d, seed=0)
    nbr, w = knn_lists(ds, ds.k, 99)
    tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    chroms = list(range(len(ds.bins)))
    with torch.no_grad():
        tables = OS.embed_tables(tsd, feats)
    mx = int(1e5) if ds.max_bin < 0 else ds.max_bin
    big, _, _, _ = OI.enumerate_pairs(ds.num_list, chroms, ds.min_bin, mx)
    mats = lambda c, cell: ds.csr.matrix(c, cell - 1)
    OI.impute_cells(tsd, tables, ds.cell_feats, ds.num_list, mats, [0], chroms, big[: 20000], 0, nbr, w, "sigmoid")  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        OI.impute_cells(tsd, tables, ds.cell_feats, ds.num_list, mats, [done % n_small], chroms, big, 0, nbr, w, "sigmoid")
        done += 1
        el = time.perf_counter() - t0
        if el > 12.0 or done >= 16:
            break
    return {"value": done * len(big) / el, "unit": "triplets/s", "cores": cores, "kind": "port",
            "sample": "%d cells of %s (P=%d) in %.1f s; oracle/ torch-CPU fp32 port of Impute.impute_process inner loop"
                      % (done, args.workload, len(big), el)}


# ------------------------------------------------------------------------------------------ our arm
TRAIN_FLOP_PER_HYPEREDGE = 1.1e6        # forward + backward of one hyperedge at d_model = 64 (3 tokens; DESIGN.md "training")
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # nominal CUDA-core fp32: 148 SMs x 128 lanes x 2 x 1.965 GHz (no measured figure in MEASURED_PEAKS.json)


def train_bench(args, rank, world, local, steps=None, want_cpu=True):
    """Secondary metric of BASELINE.json: training hyperedges/s (positives + negatives through forward + backward + ONE NCCL
    gradient allreduce + Adam), stage-2 model on C4-shaped synthetic data: mm10 @ 500 kb (4935 bins), the FULL 8 % contact density
    of C4, a 250-cell slice per rank (2000 cells / 8: the step cost does not depend on the number of cells and generating all 2000
    takes ~100 s of host time per rank), 3072 hyperedges per rank and step (512 positives + 5 negatives each, Higashi_wrapper.py:728,
    759).  Negatives, labels, weights and the neighbour lists of every batch are produced ON THE DEVICE inside the timed region;
    the positives of every step are uploaded from host memory inside it (that is the step's H2D).  Two schedules are timed:
    `shared` (every rank trains the same chromosome in a given step: no device-to-host read in the step) and `per_rank` (the
    reference semantics: every rank draws its own chromosome, DataGenerator.next_iter, Modules.py:1204-1234)."""
    import torch
    import torch.distributed as dist
    from higashi_b200 import batches, parallel, synthetic as S
    from higashi_b200.train_engine import TrainEngine
    ds = S.make_dataset("C4", n_cells=args.n_cells or 250, seed=4000 + rank)
    sd, feats = S.random_state_dict(ds, ds.d, seed=0)
    eng = TrainEngine(sd, ds.num, feats, ds.cell_feats, device=local)
    tr = parallel.DataParallelTrainer(eng, lr=1e-3)
    rng = np.random.default_rng(10 + rank)
    n_pos, neg = 512, 5
    nbr, nbr_w = knn_lists(ds, ds.k, 99 + rank)
    eng.set_graph(ds.csr, nbr, nbr_w)
    max_bin = int(max(ds.bins))
    steps = steps or args.steps

    def make_pool(chrom_rng, n):
        pool = []
        for i in range(n):
            chrom = int(chrom_rng.integers(0, len(ds.bins)))
            x, _, _ = batches.synthetic_batch(ds.csr, chrom, ds.num_list, 2 * n_pos, 0, rng)
            x = x[(x[:, 2] - x[:, 1] > 1)][:n_pos]
            pool.append((np.ascontiguousarray(x), chrom, rng.uniform(0.5, 2.0, size=len(x)).astype(np.float32)))
        return pool

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {}
    for mode in ("shared", "per_rank"):
        uniform = mode == "shared"
        pool = make_pool(np.random.default_rng(10) if uniform else np.random.default_rng(1000 + rank), 8)
        sizes, h2d = [], [0]

        def begin(i):
            x, chrom, pw = pool[i % len(pool)]
            eng.sample_batch_begin(x, chrom, neg, 0.5, max_bin, pos_w=pw, training=True, classification=True, seed=1000 * rank + i)
            h2d[0] += x.nbytes + pw.nbytes

        def step(i, want_loss):
            # batch i was started one step ago; batch i + 1 is produced (own stream) while the step on batch i runs
            B, nnz = eng.sample_batch_end()
            begin(i + 1)
            sizes.append(B)
            return tr.step(None, pool[i % len(pool)][1], None, None, None, stage=2, loss_mode=0, train_mode=True, want_loss=want_loss,
                           sampled=(B, nnz), uniform_batches=uniform)
        begin(0)
        warm = max(args.warmup, len(pool) + 2)          # every pooled batch shape once: workspaces grow to their final size
        for i in range(warm):
            step(i, False)
        barrier()
        n_steps = max(steps, 1)
        # device-timed leg: >= 2 s of steps, loss read back every 10th step
        l0 = eng.ctx.launch_count()
        # nvidia-smi polling takes the driver lock the launches need: at -lms 100 the launch-bound step ran 1.4x - 2.2x slower under
        # it (measured: 0.83 -> 1.19 / 1.9 ms per step), so this leg is sampled once a second (>= 2 samples in its >= 2.5 s)
        sampler = ClockSampler(local, period_ms=1000) if uniform else None
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        done = 0
        while True:
            for i in range(n_steps):
                step(warm + done + i, (done + i) % 10 == 0)
            done += n_steps
            torch.cuda.synchronize()
            enough = torch.tensor([1.0 if time.perf_counter() - t0 >= 2.5 else 0.0], device="cuda")
            if world > 1:
                dist.all_reduce(enough, op=dist.ReduceOp.MIN)
            if enough.item() > 0 or done >= 200 * n_steps:
                break
        ev1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        B = float(np.mean(sizes[-done:]))
        val = world * done * B / (float(ms.item()) * 1e-3)
        launches = int(eng.ctx.launch_count() - l0)
        # e2e leg: wall clock, the loss of EVERY step read back to the host (D2H), positives uploaded every step (H2D)
        h2d[0] = 0
        barrier()
        t0 = time.perf_counter()
        for i in range(n_steps):
            step(warm + done + i, True)
        barrier()
        dt = parallel.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n_steps * B / dt, "unit": "hyperedges/s", "h2d_bytes_per_step": int(h2d[0] // n_steps), "d2h_bytes_per_step": 4}
        eng.sample_batch_end()                          # drain the batch the last step started: the next schedule begins afresh
        torch.cuda.synchronize()
        rec = {"value": val, "ms_per_step": float(ms.item()) / done, "steps_timed": done, "e2e": e2e, "gpu_launches_per_step": launches / done}
        if uniform:
            rec["clocks"] = clocks
            ach = TRAIN_FLOP_PER_HYPEREDGE * done * B / (float(ms.item()) * 1e-3) / 1e12
            rec["roofline"] = {"bound": "fp32 CUDA cores", "achieved": ach, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS,
                               "peak_source": "nominal 148 SMs x 128 fp32 lanes x 2 x 1.965 GHz; MEASURED_PEAKS.json holds no fp32 figure",
                               "note": "%.2g FLOP per hyperedge (forward + backward, d_model 64); the step is ~80 small launches, bound by launch "
                                       "latency and tall-skinny fp32 GEMMs, not by a pipe" % TRAIN_FLOP_PER_HYPEREDGE}
        out[mode] = rec
    line = {"metric": "train hyperedges/s", "value": out["shared"]["value"], "unit": "hyperedges/s", "n_gpus": world,
            "ms_per_step": out["shared"]["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "C4-shaped: %d-cell slice x %d bins (mm10 @ 500 kb, 8 %% contact density), stage-2 (GraphSAGE) training step, "
                                   "%d hyperedges per rank per step (512 positives + 5 negatives each; negatives and neighbour lists produced on "
                                   "the device inside the timed region), d_model=%d, dropout on, Adam, 1 NCCL allreduce/step"
                                   % (ds.n_cells, sum(ds.bins), int(B), ds.d)},
            "schedules": out, "clocks": out["shared"].get("clocks"), "roofline": out["shared"].get("roofline"), "e2e": out["shared"]["e2e"],
            "gpu_launches": int(out["shared"]["gpu_launches_per_step"] * out["shared"]["steps_timed"])}
    if want_cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_train(ds, sd, feats)
    eng.close()
    return line


def cpu_baseline_train(ds, sd, feats):
    """Bounded CPU sample of the same step: oracle/ forward_stage2 + torch autograd + Adam on all host threads."""
    import torch
    from higashi_b200 import batches
    from oracle import scorer as OS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tsd = {k: torch.from_numpy(np.array(v, dtype=np.float32, copy=True)) for k, v in sd.items()}
    with torch.no_grad():
        cell_table = OS.encoder_branch(torch.as_tensor(feats[0]), tsd, "encode1.static_nn", 0)
    tables = [cell_table] + [None] * len(ds.bins)
    params = [t.requires_grad_(True) for k, t in tsd.items() if t.dtype == torch.float32 and "attribute_dict" not in k]
    opt = torch.optim.Adam(params, lr=1e-3)
    rng = np.random.default_rng(10)
    chrom = int(rng.integers(0, len(ds.bins)))
    x, y, w = batches.synthetic_batch(ds.csr, chrom, ds.num_list, 512, 5, rng)
    indptr, cols, vals = batches.neighbor_csr(ds.csr, x, chrom, ds.num_list, training=True, rng=rng)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    base = int(ds.num_list[chrom]) + 1
    nb = (np.stack([rows, cols]).astype(np.int64), vals, np.arange(base, base + ds.bins[chrom], dtype=np.int64))
    yt, wt = torch.from_numpy(y.astype(np.float32)), torch.from_numpy(w.astype(np.float32))
    done, t0 = 0, None
    while True:
        opt.zero_grad()
        out = OS.forward_stage2(x, nb, ds.num_list, tsd, feats, ds.cell_feats, tables)
        mean = out[0] if isinstance(out, (tuple, list)) else out
        loss = torch.nn.functional.binary_cross_entropy_with_logits(mean.reshape(-1), yt.reshape(-1), weight=wt.reshape(-1))
        loss.backward()
        opt.step()
        if t0 is None:
            t0 = time.perf_counter()            # first step = warm-up
            continue
        done += 1
        el = time.perf_counter() - t0
        if el > 8.0 or done >= 30:
            break
    return {"value": done * len(x) / el, "unit": "hyperedges/s", "cores": cores, "kind": "port",
            "sample": "%d steps of %d hyperedges in %.1f s (oracle/ forward_stage2 + torch autograd + Adam, no dropout)" % (done, len(x), el)}


def run_train(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = train_bench(args, rank, world, local)
    line.update({"steps": args.steps, "warmup": args.warmup})
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def c3_strong(args, rank, world, local):
    """Strong scaling on the configuration the north star names (BASELINE.json configs[2]): ONE synthetic 4200-cell, 500 kb,
    genome-wide dataset (5776 bins, P = 898 321 pairs per cell, k = 4), its cells split over the ranks the way the reference splits
    them over its worker processes (np.array_split, Higashi_wrapper.py:1457-1475); a rank holds its own cells plus the kNN halo
    (the neighbour cells its cells aggregate, replicated), no collective.  One "job" = every rank imputes all of its own cells;
    `value` = 4200 * P / (max over ranks of the device time of the job); `e2e` = the same job through host buffers (contacts of
    own + halo cells uploaded from pinned memory once per job, every score block read back into pinned memory)."""
    import types
    import torch
    import torch.distributed as dist
    from higashi_b200 import engine as E, parallel, synthetic as S
    def setup():
        n_glob, k = S.CONFIGS["C3"]["n_cells"], S.CONFIGS["C3"]["k"]
        nbr_g, w_g = knn_lists(types.SimpleNamespace(n_cells=n_glob), k, 4242)          # the same graph on every rank
        c0, c1 = parallel.cell_shards(n_glob, world)[rank]
        own = np.arange(c0, c1)
        local_cells = np.unique(np.concatenate([own, (nbr_g[own] - 1).reshape(-1)]))
        t_gen = time.perf_counter()
        ds = S.make_dataset_cells("C3", local_cells, threads=max(2, min(16, (os.cpu_count() or 8) // max(world, 1))))
        t_gen = time.perf_counter() - t_gen
        g2l = np.full(n_glob, -1, dtype=np.int64)
        g2l[local_cells] = np.arange(len(local_cells))
        nbr_l = np.repeat(np.arange(1, len(local_cells) + 1, dtype=np.int64)[:, None], k + 1, axis=1)     # halo cells: themselves only
        w_l = np.zeros((len(local_cells), k + 1), dtype=np.float32)
        w_l[:, 0] = 1.0
        nbr_l[g2l[own]] = g2l[nbr_g[own] - 1] + 1
        w_l[g2l[own]] = w_g[own]
        l0 = int(g2l[c0])
        assert np.array_equal(g2l[own], np.arange(l0, l0 + len(own)))                    # own cells stay contiguous
        sd, feats = S.random_state_dict(ds, ds.d, seed=0)
        eng = E.ScorerEngine(local)
        eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
        P = eng.set_pairs(list(range(len(ds.bins))), ds.min_bin, ds.max_bin)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_rowptr, h_col, h_val = pin(ds.csr.row_ptr), pin(ds.csr.col), pin(ds.csr.val)
        eng.ctx.set_contacts(h_rowptr.numpy(), h_col.numpy(), h_val.numpy())
        eng.set_neighbors(nbr_l, w_l)
        eng.prepare(getattr(E, "DEFAULT_PRECISION", E.PREC_TC16), 0, E.ACT_SIGMOID)
        cps = 250
        out = torch.empty((cps, P), dtype=torch.float32, device="cuda")
        stream = torch.cuda.current_stream()
        chunks = [(a, min(a + cps, l0 + len(own))) for a in range(l0, l0 + len(own), cps)]
        return n_glob, k, own, local_cells, t_gen, ds, l0, eng, P, h_rowptr, h_col, h_val, cps, out, stream, chunks

    # a rank that fails while it builds its shard must not leave the others waiting in a barrier: every rank reports, all agree
    err, st = None, None
    try:
        st = setup()
    except Exception as e:                                                            # noqa: BLE001
        err = repr(e)
    if world > 1:
        flag = torch.tensor([0.0 if err else 1.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() < 1.0:
            if st is not None:
                st[7].close()
            return {"error": err or "another rank failed while building its C3 shard"}
    elif err:
        return {"error": err}
    n_glob, k, own, local_cells, t_gen, ds, l0, eng, P, h_rowptr, h_col, h_val, cps, out, stream, chunks = st

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job():
        for a, b in chunks:
            eng.impute_to_device(a, b, out[: b - a], stream.cuda_stream)
    job()                                                                            # warm-up pass
    barrier()
    reps = 1 if world == 1 else 3
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(reps):
        job()
    ev1.record(stream)
    barrier()
    ms = parallel.max_over_ranks(ev0.elapsed_time(ev1) / reps)
    res = {"metric": "imputed (cell,bin_i,bin_j) triplets/s", "unit": "triplets/s", "scaling": "strong", "n_gpus": world,
           "value": n_glob * P / (ms * 1e-3), "ms_per_job": ms, "jobs_timed": reps,
           "config": {"workload": "C3: ONE synthetic dataset of %d cells x %d bins, P=%d pairs/cell, k=%d, d_model=%d, all-bin-pair "
                                  "imputation sharded by cell range (np.array_split), kNN halo replicated, no collective"
                                  % (n_glob, sum(ds.bins), P, k, ds.d),
                      "cells_this_rank": int(len(own)), "halo_cells_this_rank": int(len(local_cells) - len(own)),
                      "cells_per_call": cps, "host_seconds_generating_the_shard": round(t_gen, 1)}}

    def e2e_leg(half):
        h_out = torch.empty((cps, P), dtype=torch.float16 if half else torch.float32).pin_memory()
        nnz = int(ds.csr.row_ptr[-1])

        def e2e_job():
            eng.ctx.refresh_contacts_ptr(0, len(local_cells), h_col.data_ptr(), h_val.data_ptr(), nnz)      # H2D, once per job
            for a, b in chunks:
                eng.ctx.impute_cells_host_ptr(a, b, h_out.data_ptr(), half=half)
        e2e_job()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_job()
        barrier()
        dt = parallel.max_over_ranks((time.perf_counter() - t0) / reps)
        return {"value": n_glob * P / dt, "unit": "triplets/s", "h2d_bytes_per_job": nnz * 8,
                "d2h_bytes_per_job": int(len(own)) * P * (2 if half else 4)}
    if not args.no_e2e:
        res["e2e"] = e2e_leg(False)
        res["e2e_f16"] = e2e_leg(True)
    eng.close()
    del out
    torch.cuda.empty_cache()
    return res


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        (run_reference_train if args.mode == "train" else run_reference)(args)
        return
    if args.mode == "train":
        run_train(args)
        return
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from higashi_b200 import engine as E

    ds, sd, feats = build_dataset(args, rank)
    nbr, w = knn_lists(ds, ds.k, 99 + rank)
    eng = E.ScorerEngine(local)
    eng.load_model(sd, ds.num, feats=feats, cell_feats=ds.cell_feats)
    chroms = list(range(len(ds.bins)))
    P = eng.set_pairs(chroms, ds.min_bin, ds.max_bin)
    precision = {"fp32": E.PREC_FP32, "bf16": E.PREC_BF16, "tc16": E.PREC_TC16, "hi": E.PREC_TC16_HI,
                 "auto": getattr(E, "DEFAULT_PRECISION", E.PREC_FP32)}[args.precision]
    # pinned host copies of the inputs (for the e2e leg) -- the engine keeps its own device copy
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_rowptr, h_col, h_val = pin(ds.csr.row_ptr), pin(ds.csr.col), pin(ds.csr.val)
    eng.ctx.set_contacts(h_rowptr.numpy(), h_col.numpy(), h_val.numpy())
    eng.set_neighbors(nbr, w)
    eng.prepare(precision, 0, E.ACT_SIGMOID)

    cps = min(args.cells_per_step, ds.n_cells)
    n_ranges = max(1, ds.n_cells // cps)
    out = torch.empty((cps, P), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def step(i):
        c0 = (i % n_ranges) * cps
        eng.impute_to_device(c0, c0 + cps, out, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    l0 = eng.ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.ctx.launch_count() - l0
    tmax = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all = float(tmax.item())
    triplets_per_rank = args.steps * cps * P
    value = world * triplets_per_rank / (ms_all * 1e-3)

    # ---- per-kernel device times (separate pass: the profiling mode synchronises after each batch)
    eng.ctx.set_profiling(True)
    stage = np.zeros(4)
    for i in range(min(args.steps, 4)):
        step(i)
        torch.cuda.synchronize()
        stage += np.array(eng.ctx.last_timing())
    nprof = min(args.steps, 4)
    eng.ctx.set_profiling(False)
    k2_ms_per_step = stage[2] / nprof
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    # fc1 (128 -> d), W0, W1 (d -> d), C0 (d -> d / 2), three tokens per triplet: 110 592 FLOP at d = 64, 344 064 at d = 128
    flop_per_triplet = 3.0 * (256.0 * ds.d + 5.0 * ds.d * ds.d)
    assert ds.d != 64 or flop_per_triplet == FLOP_PER_TRIPLET_TENSOR
    achieved_tf = flop_per_triplet * cps * P / (k2_ms_per_step * 1e-3) / 1e12 if k2_ms_per_step > 0 else 0.0
    traffic = None
    try:        # DRAM bytes per launch of the dominant kernel from the committed ncu capture, scaled to this run's launch size
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k2_traffic.json")))
        per_triplet = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["triplets_per_launch"]
        traffic = per_triplet * min(cps, 40) * P
    except Exception:
        pass
    roofline = {"kernel": "K2 fused scorer", "bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "traffic_note": "DRAM bytes per K2 launch (one batch of <= 40 cells) from profiles/r01_k2_traffic.json (ncu); algorithmic bytes = 4 B/triplet",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                "stage_ms_per_step": {"gather_K1a": stage[0] / nprof, "project_K1b": stage[1] / nprof, "scorer_K2": k2_ms_per_step},
                "k2_share_of_step": float(stage[2] / max(stage[3], 1e-9))}

    # secondary roofline: the gather (K1a) against the measured HBM copy bandwidth.  Algorithmic bytes per cell:
    # the CSR rows of the cell and its k neighbours (8 B per stored entry + the row extents) + the [n_bins, d] fp32 rows it writes.
    try:
        rp = ds.csr.row_ptr
        rows_per_cell = int(sum(ds.bins))
        nnz_cell = np.diff(rp[::rows_per_cell]).astype(np.float64)            # stored entries per cell
        k1 = 1 if nbr is None else int(np.asarray(nbr).shape[1])
        cells_prof = np.concatenate([np.arange((i % n_ranges) * cps, (i % n_ranges) * cps + cps) for i in range(nprof)])
        src = cells_prof[:, None] if nbr is None else (np.asarray(nbr)[cells_prof] - 1)
        bytes_k1a = float((8.0 * nnz_cell[src].sum() + 16.0 * rows_per_cell * k1 * len(cells_prof) + 4.0 * ds.d * rows_per_cell * len(cells_prof)))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        gbs = bytes_k1a / max(stage[0] * 1e-3, 1e-12) / 1e9
        roofline["gather_K1a"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                  "note": "algorithmic bytes (k+1 CSR rows per token: 8 B per entry + extents, + the fp32 row written) / K1a device "
                                          "time; the kernel is bound by its per-candidate instruction count, not by HBM (DESIGN.md section 5)"}
    except Exception as e:                                                     # never lose the headline line over the secondary figure
        roofline["gather_K1a"] = {"error": str(e)}

    # ---- e2e: host buffers in, host buffers out, through the C ABI
    # Every step uploads ITS inputs from pinned host memory -- the stored contacts of the step's own cells (hsg_refresh_contacts:
    # the per-cell `.to(device)` of fix_cell2, Modules.py:1432-1469, for a block of cells; the kNN lists and the rest of the matrix
    # are per-job inputs, uploaded once like np.load(sparse_nondiag_adj_nbr_1.npy) at Impute.py:154) -- and reads its scores back
    # into pinned host memory (D2H double-buffered against compute inside hsg_impute_cells_host).  `e2e` moves fp32 scores;
    # `e2e_f16` is the same step with the reduced-precision score block of the sink (fp16 on the wire, half the D2H bytes).
    e2e = e2e_f16 = None
    if not args.no_e2e:
        rp, nb_rows = ds.csr.row_ptr, int(sum(ds.bins))

        def e2e_leg(half):
            h_out = torch.empty((cps, P), dtype=torch.float16 if half else torch.float32).pin_memory()
            h2d_total = [0]

            def e2e_step(i):
                c0 = (i % n_ranges) * cps
                lo, hi = int(rp[c0 * nb_rows]), int(rp[(c0 + cps) * nb_rows])
                eng.ctx.refresh_contacts_ptr(c0, c0 + cps, h_col.data_ptr() + 4 * lo, h_val.data_ptr() + 4 * lo, hi - lo)   # H2D, async
                eng.ctx.impute_cells_host_ptr(c0, c0 + cps, h_out.data_ptr(), half=half)                                  # compute + D2H
                h2d_total[0] += (hi - lo) * 8
            for i in range(min(args.warmup, 2)):
                e2e_step(i)
            barrier()
            h2d_total[0] = 0
            t0 = time.perf_counter()
            for i in range(args.steps):
                e2e_step(args.warmup + i)
            barrier()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return {"value": world * triplets_per_rank / float(tt.item()), "unit": "triplets/s",
                    "h2d_bytes_per_step": int(h2d_total[0] // max(args.steps, 1)), "d2h_bytes_per_step": int(cps * P * (2 if half else 4))}
        e2e = e2e_leg(False)
        if precision != E.PREC_FP32 and eng.scorer_variant() == E.PREC_TC16:
            e2e_f16 = e2e_leg(True)
            e2e_f16["note"] = "scores leave the scorer as fp16 (hsg_impute_cells_host_f16): <= 2.44e-4 extra on sigmoid outputs"

    eng.close()
    del out
    torch.cuda.empty_cache()
    # ---- secondary metric of BASELINE.json: train hyperedges/s (same process group, same GPUs)
    train_line = None
    if not args.no_train and not args.n_cells and not args.d_model:
        try:
            train_line = train_bench(args, rank, world, local, steps=20)
        except Exception as e:                                       # never lose the headline line over the secondary figure
            train_line = {"error": repr(e)}
    c3_line = None
    if not args.no_c3 and not args.n_cells and args.workload == "C2" and not args.d_model:
        try:
            c3_line = c3_strong(args, rank, world, local)
        except Exception as e:
            c3_line = {"error": repr(e)}
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(args, ds)
        line = {"metric": "imputed (cell,bin_i,bin_j) triplets/s", "value": value, "unit": "triplets/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if precision == E.PREC_FP32 else "f16 operands / f32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
                "config": {"workload": workload_string(args.workload, ds.n_cells, ds, P),
                           "cells_per_step": cps, "triplets_per_step": cps * P, "sharding": "cell range per rank, no collective",
                           "numa_node": numa_node,
                           "l2_policy": "each step reads a new cell range and writes a fresh %d MB score block (> 126 MB L2)"
                                        % (cps * P * 4 // (1 << 20))},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        if e2e_f16 is not None:
            line["e2e_f16"] = e2e_f16
        if train_line is not None:
            line["secondary"] = {"train": train_line}
        if c3_line is not None:
            line.setdefault("secondary", {})["c3_strong"] = c3_line
        if args.n_cells:
            line["invalid"] = "dataset shrunk with --n-cells (debug run)"
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
