#!/usr/bin/env python
"""bench.py -- pseudoalign reads/s on synthetic 150 bp reads against the salmonella_10 index (BASELINE.json configs[1]:
full-intersection, 10 M reads per GPU), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through libfulgor_gpu.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation on the host cores

A "step" is one pass of the hot path (k-mer lookup -> color-set ids -> full intersection -> CSR color lists) over one
batch of synthetic reads. `value` = reads/s with the batch already resident in HBM (CUDA events on the library's
launch stream); `e2e` = the same through the host-buffer C-ABI call (pinned host memory, H2D + kernels + D2H inside
the timed region). Multi-GPU: one process per GPU (torchrun), the index image is broadcast once with NCCL, every rank
pseudoaligns its own shard of reads, no data-path collective ("scaling": "weak").
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

INDEX = "salmonella_10.fur"
READ_LEN = 150
METRIC = "pseudoalign_reads_per_sec_150bp_full_intersection"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (0 = 10 M on salmonella_10, 1 M on many-color indexes)")
    ap.add_argument("--min-len", type=int, default=READ_LEN)
    ap.add_argument("--max-len", type=int, default=READ_LEN)
    ap.add_argument("--algo", default="fi", choices=["fi", "tu"])
    ap.add_argument("--threshold", type=float, default=0.8)
    ap.add_argument("--index", default=INDEX)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.rows = []
        self.proc = None
        self.gpu = gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        """statistics over the samples that fell inside the timed windows [(t0, t1), ...] (perf_counter seconds)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if any(a <= t <= b for a, b in windows)]
        for r in inside:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(image, reads, cid_csr, res_csr, k):
    """bytes(read) = L + 160*v + sum_{c in hit sets}(8 + ceil(bits(c)/8)) + 8 + 4*|R|, summed over the batch.
    v = valid k-mers: the synthetic reads are pure ACGT, so v = L - k + 1."""
    import fulgor_b200.imageview as iv

    bits = iv.color_set_bits(image)
    bases, off = reads
    L = np.diff(off.astype(np.int64))
    v = np.maximum(L - k + 1, 0)
    cid_off, cids = cid_csr
    set_bytes = 8 + (bits[cids] + 7) // 8
    total = int(L.sum()) + 160 * int(v.sum()) + int(set_bytes.sum()) + 8 * (len(off) - 1) + 4 * int(res_csr[1].size)
    return total


def workload_of(args, ck):
    """reads per GPU per step, genome pack to draw reads from, and the label of the workload"""
    small = args.index.startswith("salmonella_10")
    n = args.reads or (10_000_000 if small else 1_000_000)
    genomes = args.index.split(".")[0]
    lens = f"{args.min_len} bp" if args.min_len == args.max_len else f"{args.min_len}-{args.max_len} bp"
    which = {("salmonella_10.fur", "fi", 150, 150): "BASELINE.json configs[1]", ("salmonella_10.fur", "tu", 150, 150): "BASELINE.json configs[2]"}.get(
        (args.index, args.algo, args.min_len, args.max_len), "")
    if args.index.startswith("synth_4546"):
        which = ("stand-in for BASELINE.json configs[3]/[4]: 4,546 SYNTHETIC genomes (tools/make_standin_4546.sh), the real salmonella_4546 "
                 "collection cannot be downloaded here")
    return n, genomes, lens, which


def run_reference(args, rank, world):
    """the reference's own CPU implementation of the path (oracle/_ref, unmodified reference sources; else the oracle port)
    on all host threads, each step a bounded sample of the workload"""
    if rank != 0:
        return
    import _checkers as ck

    cores = os.cpu_count() or 1
    algo = 0 if args.algo == "fi" else 1
    path = ck.index_path(args.index)
    n, genomes, lens, which = workload_of(args, ck)
    if ck.reference_available():
        impl, kind, threads = ck.Reference(path), "reference", cores
    else:
        impl, kind, threads = ck.Oracle(path), "port", 1
    per_thread = (100_000 if kind == "reference" else 30_000) if impl.num_colors <= 32 else 4_000
    sample = args.cpu_sample or min(n, per_thread * threads)
    reads = ck.gen_reads(sample, args.min_len, args.max_len, seed=42, threads=min(cores, 32), genomes=genomes)
    call = (lambda: impl.pseudoalign(reads, algo, args.threshold, threads=threads)) if kind == "reference" else (
        lambda: impl.pseudoalign(reads, algo, args.threshold))
    for _ in range(args.warmup):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args), "value": value,
        "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{args.index}, {'full-intersection' if algo == 0 else f'threshold-union tau={args.threshold}'}, synthetic {lens} reads "
                               f"({which}; our arm runs {n} reads per GPU per step; this arm: a bounded sample of {sample} reads per step on the host CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sample} reads x {args.steps} steps, library-level fetch_color_set_ids + pseudoalign, {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def metric_name(args):
    m = METRIC if args.algo == "fi" else METRIC.replace("full_intersection", "threshold_union")
    if not (args.min_len == args.max_len == READ_LEN):
        m = m.replace("150bp", f"{args.min_len}_{args.max_len}bp")
    return m


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import _checkers as ck  # read generator (+ the cpu_baseline leg); never on the measured GPU path
    import fulgor_b200 as fg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one process per GPU: keep this rank's pinned buffers and copy threads on the GPU's NUMA node (the cpu_baseline leg gets all cores back)
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = fg.bind_host_thread(local_rank)

    # ---- index: rank 0 parses + flattens the .fur; one NCCL broadcast replicates the image; no other collective on the data path
    algo = fg.FULL_INTERSECTION if args.algo == "fi" else fg.THRESHOLD_UNION
    from fulgor_b200 import replicate

    if world > 1:
        idx, d_image = replicate.open_replica(ck.index_path(args.index), local_rank)
        image = d_image.cpu().numpy()
    else:
        image = fg.build_image(ck.index_path(args.index))
        idx = fg.Index.from_image(image, local_rank)

    # ---- this rank's shard of synthetic reads (weak scaling: n per GPU), in pinned host memory
    n, genomes, lens, which = workload_of(args, ck)
    cores = os.cpu_count() or 1
    gen_threads = max(1, min(32, cores // max(1, world)))
    bases_np, off_np = ck.gen_reads(n, args.min_len, args.max_len, seed=42, first=rank * n, threads=gen_threads, genomes=genomes)
    nbases = int(off_np[n])
    # output capacity: exact bound when it is small, else measured on a sample of the batch (+25 %); a step that still overflows fails loudly
    if idx.num_colors <= 32:
        cap = n * idx.num_colors
    else:
        ns = min(n, 20_000)
        so, _ = idx.pseudoalign((bases_np[: int(off_np[ns])], off_np[: ns + 1]), algo, args.threshold)
        cap = int(int(so[ns]) / ns * n * 1.25) + (1 << 20)
    pin_bases = fg.PinnedBuffer(nbases + 64)
    pin_off = fg.PinnedBuffer(8 * (n + 1))
    pin_coff = fg.PinnedBuffer(8 * (n + 1))
    pin_colors = fg.PinnedBuffer(4 * cap)
    pin_bases.view(np.uint8, nbases)[:] = bases_np
    pin_off.view(np.uint64, n + 1)[:] = off_np
    del bases_np

    # device-resident copy for the kernel-only number
    d_bases = torch.from_numpy(pin_bases.view(np.uint8, nbases)).to(dev)
    d_off = torch.from_numpy(pin_off.view(np.int64, n + 1)).to(dev)
    d_coff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_colors = torch.empty(cap, dtype=torch.int32, device=dev)

    def step_device():
        return idx.pseudoalign_device(algo, args.threshold, d_bases.data_ptr(), d_off.data_ptr(), n, 0, d_coff.data_ptr(), d_colors.data_ptr(), cap)

    def step_host():
        rc = idx.pseudoalign_raw(algo, args.threshold, pin_bases.ptr, pin_off.ptr, n, pin_coff.ptr, pin_colors.ptr, cap)
        if rc != 0:
            raise RuntimeError(f"fulgor_gpu_pseudoalign rc={rc}: {fg.lib().fulgor_gpu_last_error().decode()}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()  # nvidia-smi needs a moment to start: it runs through the warm-up, only samples inside the timed windows count

    # ---- kernel-only: inputs resident in HBM. Device time from CUDA events recorded by the library on its launch stream.
    for _ in range(args.warmup):
        total = step_device()
    barrier()
    dev_ms, k_ms, launches = 0.0, [0.0, 0.0, 0.0], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        total = step_device()
        l, ms = idx.last_kernel_times()
        dev_ms += sum(ms)
        k_ms = [a + b for a, b in zip(k_ms, ms)]
        launches += l
    barrier()
    t1 = time.perf_counter()
    wall_ms = (t1 - t0) * 1e3
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    ms_per_step = dev_ms / args.steps
    value = world * n / (ms_per_step / 1e3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory; H2D + kernels + D2H inside the timed region)
    for _ in range(args.warmup):
        step_host()
    barrier()
    t2 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    t3 = time.perf_counter()
    clocks = sampler.stop([(t0, t1), (t2, t3)])
    e2e_s = max_over_ranks(t3 - t2)
    total_colors = int(pin_coff.view(np.uint64, n + 1)[n])
    e2e = {"value": world * n * args.steps / e2e_s, "unit": UNIT, "ms_per_step": e2e_s / args.steps * 1e3,
           "h2d_bytes_per_step": nbases + 8 * (n + 1), "d2h_bytes_per_step": 8 * (n + 1) + 4 * total_colors + 24 * ((n + (1 << 20) - 1) >> 20)}

    # the device-resident and the host-buffer path must agree (same CSR)
    same = bool(np.array_equal(d_coff.cpu().numpy().view(np.uint64), pin_coff.view(np.uint64, n + 1)) and
                np.array_equal(d_colors[:total_colors].cpu().numpy().view(np.uint32), pin_colors.view(np.uint32, total_colors)))
    if not same:
        raise SystemExit("bench.py: device-resident and host-buffer results differ")

    if rank == 0:
        # ---- roofline of the dominant kernel: algorithmic bytes / its mean launch time
        fused = idx.num_colors <= 32
        res_csr = (pin_coff.view(np.uint64, n + 1), pin_colors.view(np.uint32, total_colors))
        sub = min(n, 1_000_000 if fused else 100_000)  # per-read algorithmic bytes are measured on the first `sub` reads and scaled (i.i.d. synthetic reads)
        reads_sub = (pin_bases.view(np.uint8, int(off_np[sub])), off_np[: sub + 1])
        cid_csr = idx.fetch_color_set_ids(reads_sub)
        sub_res = (res_csr[0][: sub + 1], res_csr[1][: int(res_csr[0][sub])])
        bytes_per_read = algorithmic_bytes(image, reads_sub, cid_csr, sub_res, idx.k) / sub
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        top = 0 if (fused or k_ms[0] >= k_ms[1]) else 1
        table_off = os.environ.get("FULGOR_GPU_TABLE_MAX_MB", "") == "0"  # the decoded color-set table is built unless disabled / over budget
        kname = ("k_pseudoalign_small" if fused else "k_fetch_color_sets") if top == 0 else ("k_color_sets_general" if table_off else "k_color_sets_table")
        top_ms_per_launch = k_ms[top] / args.steps
        path_ms_per_step = (k_ms[0] + k_ms[1]) / args.steps
        achieved = bytes_per_read * n / (path_ms_per_step / 1e3) / 1e9
        traffic, ncu_view = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                per_read = tj.get(f"{kname}_dram_bytes_per_read@{args.index}")
                traffic = per_read * n if per_read else None  # ncu DRAM bytes per read of the same kernel x reads per launch
                ncu_view = tj.get(f"{kname}_ncu@{args.index}")  # what actually bounds the kernel (pipe utilisation from the committed capture)
            except (ValueError, OSError):
                pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "kernel": kname, "kernel_ms_per_launch": top_ms_per_launch,
                    "kernel_share_of_step": (k_ms[top] / sum(k_ms)) if world == 1 else None,
                    "lookup_ms": k_ms[0] / args.steps, "color_sets_ms": k_ms[1] / args.steps, "scan_emit_ms": k_ms[2] / args.steps,
                    "algorithmic_bytes_per_read": bytes_per_read, "peak_source": peak_src, "ncu": ncu_view,
                    "note": "algorithmic bytes = SURVEY.md 8(d) (independent lookups: 160 B per valid k-mer + hit color sets + output) over the "
                            "lookup + color-set kernels' time. The kernels use SEED-AND-EXTEND (one MPHF lookup and ~1.2 string comparisons per run "
                            "of ~5 k-mers), i.e. do less work than independent lookups, so frac may exceed 1; the index is L2-resident and the "
                            "kernels are bound by the integer pipes, not by DRAM (see profiles/)"}

        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            os.sched_setaffinity(0, all_cpus)  # the reference gets every core of the box
            if ck.reference_available():
                ref, kind, threads = ck.Reference(ck.index_path(args.index)), "reference", cores
            else:
                ref, kind, threads = ck.Oracle(ck.index_path(args.index)), "port", 1
            per_thread = (250_000 if kind == "reference" else 200_000) if fused else 8_000
            sample = args.cpu_sample or min(n, per_thread * threads)
            sreads = (pin_bases.view(np.uint8, int(off_np[sample])), off_np[: sample + 1])
            tc = time.perf_counter()
            cpu_out = ref.pseudoalign(sreads, algo, args.threshold, threads=threads) if kind == "reference" else ref.pseudoalign(sreads, algo, args.threshold)
            cdt = time.perf_counter() - tc
            ok = bool(np.array_equal(cpu_out[0], res_csr[0][: sample + 1]) and np.array_equal(cpu_out[1], res_csr[1][: int(res_csr[0][sample])]))
            cpu_baseline = {"value": sample / cdt, "unit": UNIT, "cores": threads, "kind": kind,
                            "sample": f"first {sample} reads of the GPU batch, one pass, library-level (no parsing/formatting)",
                            "matches_gpu_output": ok}

        print(json.dumps({
            "metric": metric_name(args), "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{args.index} (k={idx.k}, m={idx.m}, {idx.num_colors} colors), "
                                   f"{'full-intersection' if algo == 0 else f'threshold-union tau={args.threshold}'}, {n} synthetic {lens} reads per GPU"
                                   + (f" ({which})" if which else ""),
                       "reads_per_gpu": n, "read_len": [args.min_len, args.max_len], "index": args.index,
                       "l2": f"inputs ({nbases / 1e9:.2f} GB of reads per step) are larger than L2; the {image.size / 1e6:.0f} MB index image is L2-resident "
                             "by nature of the workload",
                       "parallelism": f"reads sharded over {world} GPU(s), index replicated by one NCCL broadcast, no data-path collective",
                       "host_numa_binding": f"{numa_cpus} CPUs next to the GPU" if numa_cpus else "none (single node or unknown topology)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "wall_ms_per_step": wall_ms / args.steps, "results_total_colors": total_colors,
        }))
    barrier()
    idx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
