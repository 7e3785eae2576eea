#!/usr/bin/env python
"""bench.py — Gbases/s of fastplong's per-read hot loop (adapter-trim + Q-filter + stats) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 1M ONT-like reads, mean 15 kb, adapters as the reference's evaluator
auto-detects them on this generator's reads, --cut_front --cut_tail -W 10, default Q/length filters, pre+post
Stats.  A seeded host tile of 16,384 reads is replicated 64x in HBM (the full set is 30 GB of payload).
One step = one pass of processSingleEnd over the whole per-GPU batch (+ the Stats/FilterResult all-reduce when N>1).
`value` = device-resident throughput; `e2e` = the same metric through fpl_process_host with pinned HOST buffers
(H2D of every byte + D2H of the per-read records inside the timed region).

--impl reference times the reference's own CPU implementation (oracle/_ref/fastplong_ref, built from the
unmodified sources) on a bounded sample of the same workload with all host threads it can use.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TILE_READS = 16384
REPLICAS = 64
MEAN_LEN = 15000
SEED = 20260924
ALG_BYTES_PER_BASE = 2      # 1 sequence byte + 1 quality byte, each read once
ALG_BYTES_PER_READ = 64     # offset, length, result record
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
# CPU arm: a sample large enough that the binary's fixed costs (start-up, adapter detection pre-pass, report writing)
# do not dominate: 40k reads x 15 kb = 0.6 Gbases, a few seconds of reference CPU time per run with 16 workers
REF_SAMPLE_READS = int(os.environ.get("FPL_BENCH_REF_READS", "40000"))   # the override exists for the CPU test of this arm


WORKLOADS = {
    # name: (reads per tile, replicas, mean length, description)
    "c2": (TILE_READS, REPLICAS, MEAN_LEN, "configs[1]: 1M ONT reads mean 15 kb, auto-detected adapters + --cut_front --cut_tail "
                                           "-W 10, default filters, pre+post stats"),
    "c3": (8192, 8, 20000, "configs[2] shape at 1/76 scale: 65k HiFi-like reads mean 20 kb, 64-entry adapter FASTA + polyX"),
    "c5": (512, 8, 500000, "configs[4] shape: 4k ultra-long reads mean 500 kb, adapter trim only (-Q -L)"),
}


def workload_options(name="c2"):
    from fastplong_b200 import Options, synth
    if name == "c3":
        import numpy as np
        rng = np.random.default_rng(64)
        fasta = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(rng.integers(20, 45)))) for _ in range(64)]
        return Options(start_adapter=synth.ADAPTER_START, end_adapter=synth.ADAPTER_END, adapter_fasta=sorted(fasta),
                       trim_poly_x=True)
    if name == "c5":
        return Options(start_adapter=synth.ADAPTER_START, end_adapter=synth.ADAPTER_END, disable_quality_filtering=True,
                       disable_length_filtering=True)
    # -s/-e left at "auto" on the CLI; the strings below are what Evaluator::evalAdapterAndReadNum detects on this
    # generator's reads (verified with oracle/_ref/fastplong_ref) — the pre-pass itself is
    # outside the hot path.
    return Options(start_adapter=synth.ADAPTER_START, end_adapter=synth.ADAPTER_END, cut_front=True, cut_tail=True,
                   cut_window_size=10)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_device_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fastplong_b200 import synth
    from fastplong_b200.binding import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    opt = workload_options(args.workload)
    opt.device = local
    wl_reads, wl_reps, mean_len, wl_desc = WORKLOADS[args.workload]
    tile_reads = args.tile_reads or wl_reads
    replicas = args.replicas or wl_reps
    # ---- synthetic workload: seeded host tile (per rank: independent shards, weak scaling) ----
    t0 = time.time()
    kw = {}
    if args.workload == "c3":
        kw = dict(q_mean=33.0, q_sd=6.0, q_clip=60, p_polya=0.05, planted=opt.adapter_fasta[:4], p_planted=0.2)
    tile = synth.ont_like(tile_reads, mean_len, SEED + rank, **kw)
    gen_s = time.time() - t0
    tile_bytes = tile.n_bytes - 256            # drop the tail pad: replicas are laid back to back (multiple of 128)
    n_reads = tile_reads * replicas
    n_bases = tile.n_bases * replicas
    d_seq = torch.empty(tile_bytes * replicas + 256, dtype=torch.uint8, device=dev)
    d_qual = torch.empty_like(d_seq)
    h_seq = torch.from_numpy(tile.seq[:tile_bytes]).pin_memory()
    h_qual = torch.from_numpy(tile.qual[:tile_bytes]).pin_memory()
    t_seq = h_seq.to(dev, non_blocking=True)
    t_qual = h_qual.to(dev, non_blocking=True)
    for k in range(replicas):
        d_seq[k * tile_bytes:(k + 1) * tile_bytes].copy_(t_seq)
        d_qual[k * tile_bytes:(k + 1) * tile_bytes].copy_(t_qual)
    d_seq[tile_bytes * replicas:].zero_()
    d_qual[tile_bytes * replicas:].zero_()
    offs = (torch.from_numpy(tile.offsets).to(dev)[None, :] +
            (torch.arange(replicas, device=dev, dtype=torch.int64) * tile_bytes)[:, None]).reshape(-1).contiguous()
    lens = torch.from_numpy(tile.lens).to(dev).repeat(replicas).contiguous()
    del t_seq, t_qual
    torch.cuda.synchronize()

    eng = Engine(opt)
    ext = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    cur = torch.cuda.current_stream()

    def merge_stats():
        # the multi-GPU replacement of Stats::merge / FilterResult::merge: one NCCL all-reduce per block
        cur.wait_stream(ext)
        for blk in (eng.stats_device(0), eng.stats_device(1), eng.counters_device()):
            dist.all_reduce(torch.as_tensor(blk, device=dev))
        ext.wait_stream(cur)

    if world > 1:
        # all ranks pad their Stats blocks to the same number of cycles before the first all-reduce
        c = torch.tensor([int(tile.lens.max())], device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.MAX)
        eng.reserve_cycles(int(c.item()))

    def step():
        eng.reset()
        eng.process_device(d_seq.data_ptr(), d_qual.data_ptr(), offs.data_ptr(), lens.data_ptr(), n_reads,
                           d_seq.numel())
        if world > 1:
            merge_stats()

    for _ in range(max(args.warmup, 3)):
        step()
    eng.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # ---- timed region: K steps, device time by CUDA events on the library's stream ----
    eng.set_timing(True)
    l0 = eng.launch_count
    sampler = ClockSampler(physical_device_index(local))
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    for _ in range(args.steps):
        step()
    ev1.record(ext)
    eng.sync()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    ktimes = eng.kernel_times()
    eng.set_timing(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tb = torch.tensor([float(n_bases)], device=dev, dtype=torch.float64)
        dist.all_reduce(tb)
        total_bases = float(tb.item())
    else:
        total_bases = float(n_bases)
    ms_per_step = ms / args.steps
    value = total_bases / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D of the payload + D2H of the records ----
    from fastplong_b200 import PackedBatch
    from fastplong_b200.abi import RESULT_DTYPE
    h_off = torch.from_numpy(tile.offsets).pin_memory()
    h_len = torch.from_numpy(tile.lens).pin_memory()
    h_res = torch.empty(tile_reads * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    hb = PackedBatch(h_seq.numpy(), h_qual.numpy(), h_off.numpy(), h_len.numpy())
    res_view = h_res.numpy().view(RESULT_DTYPE)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        eng.reset()
        for _ in range(replicas):          # the per-GPU batch arrives as `replicas` host submissions
            eng.process(hb, out=res_view)
        if world > 1:
            merge_stats()
            eng.sync()

    e2e_value = None
    if not args.no_e2e:
        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_value = total_bases / e2e_s / 1e9
    h2d = replicas * (2 * tile_bytes + tile_reads * 12)
    d2h = replicas * tile_reads * RESULT_DTYPE.itemsize

    # ---- roofline of the dominant kernel (largest share of the step's device time) ----
    peak, peak_src = peaks()
    alg_bytes_step = ALG_BYTES_PER_BASE * n_bases + ALG_BYTES_PER_READ * n_reads   # per GPU, per step
    kern = {}
    tot_kernel_ms = sum(v[0] for v in ktimes.values()) or 1.0
    for name, (kms, cnt) in ktimes.items():
        if cnt == 0:
            continue
        per_step_ms = kms / args.steps
        kern[name] = {"ms_per_step": round(per_step_ms, 4), "share": round(kms / tot_kernel_ms, 4),
                      "launches_per_step": cnt // args.steps,
                      "achieved_gbs": round(alg_bytes_step / (per_step_ms * 1e-3) / 1e9, 1) if per_step_ms > 0 else None}
    dom = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
    traffic = None
    secondary = None      # the integer-pipe roofline, reported where it binds (from the committed ncu capture)
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if dom in tj:   # DRAM bytes per algorithmic byte from the committed ncu --set full capture
                nl = max(1, kern[dom]["launches_per_step"])
                traffic = tj[dom]["dram_bytes_per_alg_byte"] * alg_bytes_step / nl
                if "alu_pipe_active_pct" in tj[dom]:
                    secondary = {"bound": "integer logic pipe (LOP3/SHF/PRMT issue slots)",
                                 "frac": round(tj[dom]["alu_pipe_active_pct"] / 100.0, 4),
                                 "source": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active, ncu --set full capture "
                                           "summarised in profiles/r01_ncu_full_summary.txt (not measured live)"}
        except Exception:
            traffic = None
            secondary = None
    roofline = None
    if dom:
        nl = max(1, kern[dom]["launches_per_step"])
        achieved = kern[dom]["achieved_gbs"]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": alg_bytes_step / nl, "avg_launch_ms": round(kern[dom]["ms_per_step"] / nl, 5),
                    "launches_per_step": nl, "secondary": secondary}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = reference_cpu_run(sample_reads=REF_SAMPLE_READS, repeats=1)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "Gbases/s processed (adapter-trim + Q-filter + pre/post stats)", "value": round(value, 3),
        "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": wl_desc,
                   "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "mean_len": mean_len,
                   "tile": f"{tile_reads} seeded reads x {replicas} replicas in HBM",
                   "adapters": "as auto-detected by the reference evaluator on this generator (30 bp start + revcomp end)",
                   "l2": "inputs (2 x %.1f GB) larger than L2; no flush needed" % (d_seq.numel() / 1e9),
                   "parallelism": f"reads sharded over {world} GPU(s), NCCL all-reduce of Stats/FilterResult blocks"
                   if world > 1 else "single GPU", "read_tiling": (os.environ.get("FPL_TILE_MBASES") + " Mbases") if os.environ.get("FPL_TILE_MBASES") else "none (every kernel streams the whole batch)",
                   "host_tile_gen_s": round(gen_s, 1)},
        "clocks": clocks,
        "e2e": None if e2e_value is None else {"value": round(e2e_value, 3), "unit": "Gbases/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "how": "fpl_process_host on pinned host buffers, %d submissions/step; inside a call the upload runs in 16 MiB pieces "
                       "on a copy stream and the kernels of a piece start when it has arrived" % replicas},
        "gpu_launches": int(launches),
        "roofline": roofline, "kernels": kern, "cpu_baseline": cpu_baseline,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
_SAMPLE_CACHE = {}


def reference_cpu_run(sample_reads, repeats):
    """fastplong_ref (the unmodified reference, oracle/_ref) on a bounded sample of the same workload, plain FASTQ
    on tmpfs, all the worker threads it accepts (-w min(16, nproc)).  Returns the cpu_baseline object."""
    from fastplong_b200 import synth
    binary = os.path.join(ROOT, "oracle", "_ref", "fastplong_ref")
    if not os.path.exists(binary):
        return {"value": None, "unit": "Gbases/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/fastplong_ref not built"}
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    fq = os.path.join(tmp, f"fpl_bench_{os.getpid()}_{sample_reads}.fq")
    if fq in _SAMPLE_CACHE:
        n_bases = _SAMPLE_CACHE[fq]
    else:
        batch = synth.ont_like(sample_reads, MEAN_LEN, SEED)
        synth.to_fastq(batch, fq)
        n_bases = _SAMPLE_CACHE[fq] = batch.n_bases
        import atexit
        atexit.register(lambda: os.path.exists(fq) and os.remove(fq))
    cores = min(16, os.cpu_count() or 1)
    opt = workload_options()
    opt.start_adapter = opt.end_adapter = "auto"   # the binary runs its own detection pre-pass
    best = None
    try:
        for _ in range(repeats):
            cmd = [binary, "-i", fq, "-o", fq + ".out", "-j", fq + ".json", "-h", fq + ".html", "-w", str(cores), "-V"]
            cmd += opt.cli_flags()
            # -V makes the reference log "start to load data" (src/seprocessor.cpp:334) and "start to generate reports"
            # (:105-106); its own timestamps have 1 s resolution, so the lines are stamped here as they arrive
            t0 = time.perf_counter()
            pr = subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, bufsize=1)
            t_load = t_rep = None
            tail = []
            for ln in pr.stderr:
                now = time.perf_counter()
                tail.append(ln)
                if t_load is None and "start to load data" in ln:
                    t_load = now
                elif "start to generate reports" in ln:
                    t_rep = now
            pr.wait()
            wall = time.perf_counter() - t0
            if pr.returncode != 0:
                return {"value": None, "unit": "Gbases/s", "cores": cores, "kind": "reference",
                        "sample": "fastplong_ref failed: " + "".join(tail)[-200:]}
            phase = (t_rep - t_load) if (t_load is not None and t_rep is not None and t_rep > t_load) else wall
            if best is None or phase < best[0]:
                best = (phase, wall)
    finally:
        for suffix in (".out", ".json", ".html"):
            try:
                os.remove(fq + suffix)
            except OSError:
                pass
    phase, wall = best
    return {"value": round(n_bases / phase / 1e9, 5), "unit": "Gbases/s", "cores": cores, "kind": "reference",
            "sample": f"{sample_reads} reads / {n_bases} bases of the same generator, fastplong_ref -w {cores}, plain FASTQ on "
                      f"tmpfs; value = bases / processing phase ({phase:.2f} s between the -V lines 'start to load data' and "
                      f"'start to generate reports': reader + workers + writer, the path itself); the whole binary took "
                      f"{wall:.2f} s (adapter detection pre-pass, Stats allocation and JSON/HTML reports included)",
            "phase_s": round(phase, 3), "wall_s": round(wall, 3), "whole_binary_gbases_s": round(n_bases / wall / 1e9, 5),
            "bases": n_bases}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_reads = REF_SAMPLE_READS
    for _ in range(max(args.warmup, 1)):
        reference_cpu_run(sample_reads, 1)
    walls, base = [], None
    for _ in range(args.steps):
        base = reference_cpu_run(sample_reads, 1)
        if base["value"] is None:
            emit({"impl": "reference", "unavailable": base["sample"]})
            return
        walls.append(base["phase_s"])
    ms = 1e3 * sum(walls) / len(walls)      # a step = the processing phase of one run over the sample
    value = base["bases"] / (ms * 1e-3) / 1e9
    cb = dict(base)
    cb["value"] = round(value, 5)
    line = {"impl": "reference", "metric": "Gbases/s processed (adapter-trim + Q-filter + pre/post stats)",
            "value": round(value, 5), "unit": "Gbases/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": round(ms, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1] shape, bounded sample: %d ONT reads mean 15 kb (%.1f Gbases) per step, auto-detected "
                                   "adapters + --cut_front --cut_tail -W 10 (reference CPU build, host cores only; a step is timed "
                                   "over the binary's processing phase)" % (sample_reads, base["bases"] / 1e9)},
            "cpu_baseline": cb,
            "e2e": {"value": round(value, 5), "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner with a plain
    printf when NCCL_DEBUG=VERSION): keep a private handle on the real stdout for the line and point file descriptor 1
    at stderr for everybody else."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="c2 = BASELINE configs[1] (the bench line); c3/c5 = other BASELINE config shapes, informational")
    ap.add_argument("--tile-reads", type=int, default=0)
    ap.add_argument("--replicas", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling aid: skip the end-to-end leg (keeps an ncu launch list short)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
