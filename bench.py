#!/usr/bin/env python
"""bench.py — Gbases/s SHIMMER-indexed (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[1]): sequence_to_shmmrs on 1000 synthetic 5 Mb contigs (uniform ACGT), spec
w=80 k=56 r=4 min_span=64, per GPU (weak scaling: every rank indexes its own 1000 contigs; no data-path collective).
A step = one pass of the hot path over that batch.

  value : device-resident throughput (inputs already in HBM), CUDA events on the launch stream, max over ranks
  e2e   : the same work through the reference-facing C ABI call pgr_b200_shmmrs_batch with HOST (pinned) buffers:
          H2D of the 5 GB of bases and D2H of the MM128 result inside the timed region
  roofline : dominant kernel (l0_minimizers), algorithmic bytes (1 B/base + 16 B/shimmer) / its CUDA-event time
  cpu_baseline : the C++ oracle (restatement of the reference's rayon path) on the box's host cores, bounded sample

  e2e_pageable : the same call with plain pageable host memory (what a Rust `&Vec<u8>` hands over)
  assembly_like : the same batch decorated like a real assembly (N gaps, soft-masked lower case, microsatellites, inverted
          repeats; bench_synth.decorate_assembly_like) — the paths uniform ACGT never takes — with its own oracle check
  index_build : BASELINE.json configs[2], the full ShmmrFragMap build on 94 x 50 Mb haplotypes, STRONG scaling over the N
          GPUs with the NCCL all-to-all inside libpgr_b200, host-to-CSR and HBM-resident, parity flags (bench_index.py)

`--impl reference` times the oracle alone (all host threads), same metric/config, on a bounded sample per step.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

N_CONTIGS = 1000
CONTIG_LEN = 5_000_000
SPEC = (80, 56, 4, 64)
SLACK = 16384
ALGO_BYTES_PER_BASE_IN = 1.0
ALGO_BYTES_PER_SHMMR = 16.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--contigs", type=int, default=N_CONTIGS, help="contigs per GPU (default = the BASELINE config)")
    ap.add_argument("--contig-len", type=int, default=CONTIG_LEN)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-index", action="store_true", help="skip the config-3 index build (index_build object)")
    ap.add_argument("--no-assembly", action="store_true", help="skip the assembly-like variant of the batch")
    ap.add_argument("--no-sketch", action="store_true", help="skip the sketch-mode leg")
    ap.add_argument("--index-haps", type=int, default=94)
    ap.add_argument("--index-hap-len", type=int, default=50_000_000)
    return ap.parse_args()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_to_gpu_cpus(local_rank):
    """Pin this rank (and the pinned host buffers it allocates afterwards) to the CPUs NVML reports as local to its GPU,
    so that on a two-socket box the H2D copies of the e2e leg do not cross the socket interconnect.  Returns the CPU list
    or None when NVML, the mask or the cgroup leave nothing to bind to.  PGR_B200_NO_BIND=1 disables it."""
    if os.environ.get("PGR_B200_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        index = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            ids = [int(t) for t in vis.split(",")]
            if local_rank < len(ids):
                index = ids[local_rank]
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        use = cpus & os.sched_getaffinity(0)
        if not use or use == os.sched_getaffinity(0):
            return None if not use else sorted(use)
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None


def synth_contig(seed, length):
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=length, dtype=np.uint8)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def stop(self):
        t_end = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_begin = getattr(self, "t_begin", 0.0)
        # the sampler is started before the warm-up so that it is already streaming; keep the samples that arrived during
        # the timed region (plus one polling interval of slack)
        for ts, ln in self.lines:
            if ts < t_begin or ts > t_end + 0.05:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    """the reference arm: the C++ oracle (port of the reference's rayon path) on all host cores, bounded sample"""
    if rank != 0:
        return
    import orc
    cores = host_cores()
    # bounded sample of the workload: enough contigs for ~1-2 s of wall time per step on this box
    n_sample = min(args.contigs, max(2 * cores, 16))
    seqs = [synth_contig(1000 + i, args.contig_len) for i in range(n_sample)]
    spec = orc.mkspec(*SPEC)
    rids = list(range(n_sample))
    bases = n_sample * args.contig_len
    for _ in range(args.warmup):
        orc.shmmrs_batch(rids[:cores], seqs[:cores], spec, False, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.shmmrs_batch(rids, seqs, spec, False, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = bases / dt / 1e9
    sample = "%d of %d contigs x %d bases per step, one sequence per thread (seq_db.rs:461)" % (n_sample, args.contigs, args.contig_len)
    line = {
        "impl": "reference", "metric": "Gbases/sec SHIMMER-indexed", "value": val, "unit": "Gbases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "sequence_to_shmmrs on %d synthetic %d-base contigs per GPU, w=80 k=56 r=4 min_span=64" % (args.contigs, args.contig_len),
                   "note": "C++ restatement of the reference's rayon CPU path (the Rust reference cannot be built in this image)"},
        "cpu_baseline": {"value": val, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    bound = bind_to_gpu_cpus(local_rank)   # before torch / CUDA create their threads and pinned buffers

    import torch
    import torch.distributed as dist

    import pgr_tk_b200 as pg

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if pg.device_count() <= local_rank:
        raise SystemExit("no CUDA device for rank %d: this benchmark has no CPU fallback" % rank)
    pg.set_default_device(local_rank)

    n_contigs, clen = args.contigs, args.contig_len
    assert clen % 32 == 0
    bases = n_contigs * clen
    spec = pg.ShmmrSpec(*SPEC)

    # ---- synthetic workload, generated on the device (seeded per rank), laid out as the library's sequence store ----
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    store = torch.zeros(SLACK + bases + SLACK, dtype=torch.uint8, device=dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    chunk = 250_000_000
    for o in range(0, bases, chunk):
        m = min(chunk, bases - o)
        idx = torch.randint(0, 4, (m,), generator=g, device=dev, dtype=torch.uint8)
        store[SLACK + o: SLACK + o + m] = lut[idx.to(torch.int64)]
        del idx
    torch.cuda.synchronize()
    offs = (SLACK + np.arange(n_contigs, dtype=np.uint64) * np.uint64(clen)).astype(np.uint64)
    lens = np.full(n_contigs, clen, dtype=np.uint64)

    ctx = pg.Ctx(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_device_seqs(store.data_ptr(), offs, lens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------------------
    n_shmmrs = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        n_shmmrs = ctx.shmmrs(spec)
    barrier()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0_ms, stage_ms, launches = [], {}, 0
    ev0.record(stream)
    for _ in range(args.steps):
        n_shmmrs = ctx.shmmrs(spec)
        launches += ctx.counters()[0]
        for nm, ms in ctx.timings():
            stage_ms.setdefault(nm, []).append(ms)
            if nm == "l0_minimizers":
                l0_ms.append(ms)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * bases / (dev_ms_max * 1e-3) / 1e9

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    l0 = statistics.mean(l0_ms) if l0_ms else float("nan")
    algo_bytes = ALGO_BYTES_PER_BASE_IN * bases + ALGO_BYTES_PER_SHMMR * n_shmmrs
    achieved = algo_bytes / (l0 * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "l0_kernel<80,56> (l0_minimizers)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": l0,
                "algorithmic_bytes_per_launch": algo_bytes,
                "note": "integer-issue bound, not HBM bound (SURVEY §8d): ~94 int32 instructions per base, ALU pipe 73 % busy; see int_issue, DESIGN.md and profiles/"}
    traffic_file = os.path.join(ROOT, "profiles", "l0_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            roofline["traffic"] = tj.get("dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source")
            if tj.get("thread_instr_per_base") and clocks.get("sm_mhz"):
                # the ceiling that actually binds this kernel: INT32 instruction issue, 128 lanes per SM per clock
                tipb = float(tj["thread_instr_per_base"])
                n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
                peak_i = n_sm * 128 * float(clocks["sm_mhz"]) * 1e6
                ach_i = tipb * bases / (l0 * 1e-3)
                roofline["int_issue"] = {"thread_instr_per_base": tipb, "achieved_instr_per_s": ach_i, "peak_instr_per_s": peak_i,
                                         "frac": ach_i / peak_i, "alu_pipe_busy": tj.get("alu_pipe_busy"),
                                         "source": tj.get("instr_source"), "peak_source": "%d SMs x 128 lanes x %.0f MHz (sampled)" % (n_sm, clocks["sm_mhz"])}
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N=1 only); doubles as a full-size parity spot check of the resident result ------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import orc
        cores = host_cores()
        n_sample = min(n_contigs, max(2 * cores, 16))
        sample = store[SLACK: SLACK + n_sample * clen].cpu().numpy()
        seqs = [sample[i * clen:(i + 1) * clen] for i in range(n_sample)]
        ospec = orc.mkspec(*SPEC)
        orc.shmmrs_batch(list(range(min(cores, n_sample))), seqs[:cores], ospec, False, nthreads=cores)  # warm
        t0 = time.perf_counter()
        cpu_out, cpu_off = orc.shmmrs_batch(list(range(n_sample)), seqs, ospec, False, nthreads=cores)
        dt = time.perf_counter() - t0
        got, goff = ctx.shmmrs_download()
        k = int(goff[n_sample])
        assert k == int(cpu_off[-1]) and np.array_equal(got[:k], cpu_out), "GPU result differs from the oracle on the CPU sample"
        cpu = {"value": n_sample * clen / dt / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
               "sample": "first %d of %d contigs x %d bases, one sequence per thread (seq_db.rs:461); parity of these contigs checked" % (n_sample, n_contigs, clen)}
        del sample, seqs, got

    # ---- sketch mode of the same call (ShmmrSpec.sketch, shmmrutils.rs:558-630) on the same resident batch, rank 0 / N = 1 ----
    sketch = None
    if world == 1 and not args.no_sketch:
        sspec = pg.ShmmrSpec(*SPEC, True)
        for _ in range(2):
            n_sk = ctx.shmmrs(sspec)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = min(5, args.steps)
        k0.record(stream)
        for _ in range(reps):
            n_sk = ctx.shmmrs(sspec)
        k1.record(stream)
        torch.cuda.synchronize()
        s_ms = k0.elapsed_time(k1) / reps
        sk_stages = dict(ctx.timings())
        # the dominant kernel reads 1 B per base and writes 8 B of masks per 32 bases
        sketch = {"value": bases / (s_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": s_ms, "reps": reps, "shmmrs": int(n_sk),
                  "stages_ms": sk_stages, "spec": "k=56 r=4 min_span=64 sketch=true (threshold 2^(60-r): 1 position in 256 before the span filter)",
                  "mask_kernel_algorithmic_gbs": (bases * (1.0 + 8.0 / 32.0)) / (sk_stages.get("sketch_masks", s_ms) * 1e-3) / 1e9}
        if not args.no_cpu:
            import orc
            n_chk = min(n_contigs, 8)
            sample = store[SLACK: SLACK + n_chk * clen].cpu().numpy()
            cpu_out, cpu_off = orc.shmmrs_batch(list(range(n_chk)), [sample[i * clen:(i + 1) * clen] for i in range(n_chk)],
                                                orc.mkspec(*SPEC, True), False, nthreads=host_cores())
            got, goff = ctx.shmmrs_download()
            kk = int(goff[n_chk])
            sketch["parity_first_%d_contigs_vs_oracle" % n_chk] = bool(kk == int(cpu_off[-1]) and np.array_equal(got[:kk], cpu_out))
            assert sketch["parity_first_%d_contigs_vs_oracle" % n_chk], "sketch mode: GPU result differs from the oracle"
            del sample
        ctx.shmmrs(spec)   # leave the minimizer result resident, as the legs below expect

    # ---- end to end through the C ABI with host buffers ----------------------------------------------------------------
    e2e = e2e_direct = e2e_pageable = assembly = None
    if not args.no_e2e:
        L = pg.lib()
        hb = pg.host_alloc(bases)
        hb_t = torch.from_numpy(hb.array)
        hb_t.copy_(store[SLACK: SLACK + bases])   # same bytes as the device-resident run
        torch.cuda.synchronize()
        ptrs = (C.c_void_p * n_contigs)(*[hb.ptr + i * clen for i in range(n_contigs)])
        clens = (C.c_size_t * n_contigs)(*([clen] * n_contigs))
        rids = np.arange(n_contigs, dtype=np.uint32)
        offs_out = np.zeros(n_contigs + 1, dtype=np.uint64)
        out = C.c_void_p()
        def time_e2e(n_iter, skip):
            ts, d2h_ = [], 0
            for it in range(skip + n_iter):
                barrier()
                t0 = time.perf_counter()
                rc = L.pgr_b200_shmmrs_batch(n_contigs, rids.ctypes.data, ptrs, clens, C.byref(spec), 0, C.byref(out), offs_out.ctypes.data)
                t1 = time.perf_counter()
                if rc != 0:
                    raise SystemExit("pgr_b200_shmmrs_batch failed: %s" % L.pgr_b200_last_error().decode())
                assert int(offs_out[-1]) == n_shmmrs, (int(offs_out[-1]), n_shmmrs)
                d2h_ = int(offs_out[-1]) * 16 + (n_contigs + 1) * 8
                L.pgr_b200_free(out)
                if it >= skip:
                    ts.append(t1 - t0)
            ms = statistics.mean(ts) * 1e3
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), d2h_

        # default transport: bases packed to 3 bits on the host cores, expanded on the device (pack_upload.cuh)
        tb0 = pg.lib().pgr_b200_transport_bytes()
        e_ms, d2h = time_e2e(args.steps, 3)
        took_packed = pg.lib().pgr_b200_last_transport() == pg.TRANSPORT_PACKED
        # what really crossed PCIe per step: bit planes of the packed slots (8 B per 32 bases when a slot holds bases only, 12
        # otherwise) + the plain bytes of the slots the hybrid feeding copied as they are; counted by the library
        packed_bytes = (pg.lib().pgr_b200_transport_bytes() - tb0) // (args.steps + 3)
        e2e = {"value": world * bases / (e_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": packed_bytes if took_packed else bases, "d2h_bytes_per_step": d2h, "host_input_bytes_per_step": bases,
               "transport": ("packed: %d host threads (%s) turn the caller's bytes into 3 bit planes per 32-base block, 12 B per 32 bases cross PCIe, "
                             "unpack_kernel restores canonical ASCII in the device store" % (pg.lib().pgr_b200_pool_threads(), pg.pack_isa())) if took_packed
                            else "direct copy of the page-locked bytes (the library's choice with %d host threads for this rank)" % pg.lib().pgr_b200_pool_threads(),
               "api": "pgr_b200_shmmrs_batch (host pointers in pinned memory -> host MM128 array)"}
        # A/B: the caller's bytes copied as they are (round-1 path; asynchronous because the buffer is page-locked)
        pg.set_transport(pg.TRANSPORT_DIRECT)
        d_ms, _ = time_e2e(min(5, args.steps), 1)
        pg.set_transport(pg.TRANSPORT_PACKED)
        e2e_direct = {"value": world * bases / (d_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": d_ms, "h2d_bytes_per_step": bases,
                      "api": "pgr_b200_shmmrs_batch, PGR_TRANSPORT_DIRECT (pinned host pointers, 1 B per base over PCIe)"}
        # ---- the same call on pageable memory: what a drop-in `sequence_to_shmmrs(&Vec<u8>)` caller hands over -----------
        pg_arr = np.empty(bases, dtype=np.uint8)
        np.copyto(pg_arr, hb.array)
        pptrs = (C.c_void_p * n_contigs)(*[pg_arr.ctypes.data + i * clen for i in range(n_contigs)])
        ptimes = []
        for it in range(1 + min(3, args.steps)):
            barrier()
            t0 = time.perf_counter()
            rc = L.pgr_b200_shmmrs_batch(n_contigs, rids.ctypes.data, pptrs, clens, C.byref(spec), 0, C.byref(out), offs_out.ctypes.data)
            t1 = time.perf_counter()
            if rc != 0:
                raise SystemExit("pgr_b200_shmmrs_batch (pageable) failed: %s" % L.pgr_b200_last_error().decode())
            assert int(offs_out[-1]) == n_shmmrs
            L.pgr_b200_free(out)
            if it >= 1:
                ptimes.append(t1 - t0)
        p_ms = statistics.mean(ptimes) * 1e3
        tp = torch.tensor([p_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        p_ms = float(tp.item())
        e2e_pageable = {"value": world * bases / (p_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": p_ms, "reps": len(ptimes),
                        "api": "pgr_b200_shmmrs_batch (host pointers in PAGEABLE memory -> host MM128 array)"}
        del pg_arr, pptrs

        # ---- assembly-like decoration of the same batch (rank 0, N = 1): gaps, soft masking, microsatellites, inverted repeats
        if not args.no_assembly and world == 1:
            import concurrent.futures as cf
            import bench_synth as S
            with cf.ThreadPoolExecutor(max_workers=min(16, host_cores())) as ex:
                list(ex.map(lambda i: S.decorate_assembly_like(hb.array[i * clen:(i + 1) * clen], 5000 + i), range(n_contigs)))
            store[SLACK: SLACK + bases].copy_(hb_t)
            torch.cuda.synchronize()
            ctx.set_device_seqs(store.data_ptr(), offs, lens)
            for _ in range(2):
                n_asm = ctx.shmmrs(spec)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = min(5, args.steps)
            a0.record(stream)
            for _ in range(reps):
                n_asm = ctx.shmmrs(spec)
            a1.record(stream)
            torch.cuda.synchronize()
            a_ms = a0.elapsed_time(a1) / reps
            cnt = ctx.counters()
            asm_stages = dict(ctx.timings())
            assembly = {"value": bases / (a_ms * 1e-3) / 1e9, "unit": "Gbases/s", "ms_per_step": a_ms, "reps": reps, "shmmrs": n_asm,
                        "slowdown_vs_uniform_acgt": a_ms / dev_ms_max,
                        "n_fraction": float((hb.array[:50_000_000] == ord("N")).mean()), "lower_case_fraction": float(((hb.array[:50_000_000] & 0x20) != 0).mean()),
                        "whole_sequence_replays": int(cnt[2]), "local_patches": int(cnt[4]), "kernel_launches_per_step": int(cnt[0]),
                        "stages_ms": asm_stages,
                        "input": "config-2 batch + bench_synth.decorate_assembly_like: N runs 10 kb-5 Mb, soft-masked runs, (AT)n/(CG)n/(ACGT)n and inverted repeats every ~50 kb"}
            if not args.no_cpu:
                import orc
                n_chk = min(n_contigs, max(host_cores(), 8))
                seqs_chk = [hb.array[i * clen:(i + 1) * clen] for i in range(n_chk)]
                cpu_out, cpu_off = orc.shmmrs_batch(list(range(n_chk)), seqs_chk, orc.mkspec(*SPEC), False, nthreads=host_cores())
                got, goff = ctx.shmmrs_download()
                k = int(goff[n_chk])
                assembly["parity_first_%d_contigs_vs_oracle" % n_chk] = bool(k == int(cpu_off[-1]) and np.array_equal(got[:k], cpu_out))
                assert assembly["parity_first_%d_contigs_vs_oracle" % n_chk], "assembly-like batch: GPU result differs from the oracle"
        hb.free()

    ctx.close()
    del store
    torch.cuda.empty_cache()

    # ---- BASELINE configs[2]: the multi-GPU ShmmrFragMap build, strong scaling, NCCL all-to-all inside the library -----------
    index_build = None
    if not args.no_index:
        import bench_index
        from pgr_tk_b200 import distributed as PD
        comm = PD.init_comm(local_rank) if world > 1 else pg.Comm(pg.comm_unique_id(), 0, 1, local_rank)
        index_build = bench_index.run(pg, torch, dist, rank, world, local_rank, comm, reps=3, n_hap=args.index_haps, hap_len=args.index_hap_len,
                                      oracle_parity=not args.no_cpu, cores=host_cores())
        comm.close()

    if rank == 0:
        line = {
            "metric": "Gbases/sec SHIMMER-indexed", "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "sequence_to_shmmrs on %d synthetic %d-base contigs per GPU, w=80 k=56 r=4 min_span=64" % (n_contigs, clen),
                       "bases_per_gpu": bases, "shmmrs_per_gpu": n_shmmrs, "l2": "inputs (%.1f GB) larger than L2" % (bases / 1e9),
                       "parallelism": "sequences sharded over %d GPU(s), no data-path collective" % world,
                       "cpu_binding": ("rank 0 bound to %d GPU-local CPUs (NVML affinity)" % len(bound)) if bound else "none"},
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
            "stages_ms": {k: statistics.mean(v) for k, v in stage_ms.items()},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if e2e_direct is not None:
            line["e2e_direct"] = e2e_direct
        if e2e_pageable is not None:
            line["e2e_pageable"] = e2e_pageable
        if assembly is not None:
            line["assembly_like"] = assembly
        if sketch is not None:
            line["sketch_mode"] = sketch
        if index_build is not None:
            line["index_build"] = index_build
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
