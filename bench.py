#!/usr/bin/env python
"""bench.py -- Spectre spectral-mix forward throughput on B200 (BASELINE.json metric, configs[4] workload).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N (SURVEY 8e / BASELINE.json configs[4]): a GLOBAL batch of 8192 rows of (4096 tokens x 768 channels)
fp32, batch-sharded -- rank r of W owns rows shard_rows(8192, r, W) and streams them through the fused kernel in
micro-batches of <= 148 rows (one launch each: 148 rows = 96 full waves of the 148 persistent CTAs) into a ring of reused
output buffers.  A "step" is one pass over the whole global batch, so per-rank work shrinks as N grows: STRONG scaling.
Every (batch row, channel) column is an independent transform (spectre.py:506, :712-713): no data-path collective
exists; NCCL carries the barriers and the max / sum reductions of the measurement.  The rank's shard of V and of the
gate is resident in HBM when the timed region starts (103 GB + 6.5 GB at N = 1); only the outputs are recycled.

Rank 0 prints ONE JSON line:
  value      whole-job tokens/s over the timed K steps (a seconds-long region at N = 1: a SUSTAINED figure, clocks and
             throttle reasons in `clocks`)
  roofline   the kernel's algorithmic bytes (6336.1 B/token) / CUDA-event time against the measured copy
             bandwidth (MEASURED_PEAKS.json); `roofline.burst` = the same kernel timed as 20 launches after a cool-down
             (the round-1 headline figure), so both fractions are visible
  parity_check   sampled batch rows of the timed output compared with the CPU oracle (rel-L2 <= 1e-5, else exit 1)
  e2e        the same metric through the host-buffer C-ABI entry (spectre_mix_fwd_host: pinned host memory -> H2D ->
             kernel -> D2H inside the timed region), per rank and aggregated, next to the box's pure-copy ceiling
  cpu_baseline   the reference's CPU path (torch.fft head loop restated in oracle/) on this box's host cores, and the
             stock reference module (baseline/_ref/spectre.py, when present) beside it
`--impl reference` times only the CPU path.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEQ, D_MODEL, HEADS, GROUPS = 4096, 768, 12, 4
D_G = D_MODEL // HEADS // GROUPS          # 16 channels per gate group
NG = HEADS * GROUPS                       # 48 gate rows per batch row
F_HALF = SEQ // 2 + 1
METRIC = "Spectre-block fwd tokens/sec at seq=4096 d=768; achieved HBM GB/s vs peak"
UNIT = "tokens/s"
REF_COPY = os.path.join(ROOT, "baseline", "_ref", "spectre.py")


def algorithmic_bytes(B: int, es: int = 4) -> int:
    """V in + out + gate, per call."""
    return B * SEQ * D_MODEL * es * 2 + B * NG * F_HALF * 8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic_per_token():
    """dram bytes per token from the committed ncu capture (profiles/roofline_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_token"])
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML while a timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        self.period, self.samples, self.power, self.reasons, self.max_mhz = period_s, [], [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_min_mhz": min(self.samples), "sm_max_mhz": self.max_mhz,
                "power_w_max": round(max(self.power), 1) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- CPU legs
def _cpu_inputs(torch, batch):
    gen = torch.Generator().manual_seed(0)
    V = torch.randn(batch, SEQ, D_MODEL, generator=gen)
    gate = torch.randn(batch, NG, F_HALF, dtype=torch.cfloat, generator=gen)
    return V, gate


def cpu_reference_path(torch, batch: int, min_seconds: float, max_reps: int):
    """The reference's CPU path (spectre.py:506, :542-553 looped over heads as :712-713) on host cores."""
    from oracle import spectre_mix_oracle as oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    V, gate = _cpu_inputs(torch, batch)
    with torch.no_grad():
        oracle.mix_head_loop(V, gate, SEQ, HEADS)          # warm-up (MKL plan creation)
        times = []
        t_all = time.perf_counter()
        while len(times) < max_reps and (time.perf_counter() - t_all < min_seconds or len(times) < 3):
            t0 = time.perf_counter()
            oracle.mix_head_loop(V, gate, SEQ, HEADS)
            times.append(time.perf_counter() - t0)
    return {"tokens": batch * SEQ, "times": times, "threads": torch.get_num_threads()}


def stock_reference_timings(torch, batch: int = 4, reps: int = 3):
    """The UNMODIFIED reference module (baseline/_ref/spectre.py, a git-ignored copy made by __graft_entry__.build())
    on host cores, BASELINE.md section 5 step 2b: the full SpectreBlock.forward (spectre.py:967-982) and, inside the same
    forward, the time between entering torch.fft.rfft and leaving torch.fft.irfft of every head (spectre.py:506-551:
    the hot path plus the gate generator that sits between the two calls).  None when the copy is absent."""
    if not os.path.exists(REF_COPY):
        return None
    import importlib.util
    import warnings
    spec = importlib.util.spec_from_file_location("_ref_spectre", REF_COPY)
    ref = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(ref)
    torch.manual_seed(0)
    blk = ref.SpectreBlock(D_MODEL, HEADS, SEQ, pooling_type="mean", wavelet_on_rate=0.0, num_groups=GROUPS).eval()
    x = torch.randn(batch, SEQ, D_MODEL)
    # time spent between rfft entry and irfft exit, accumulated over the 12 heads of a forward (timers around the library
    # calls only: the reference file itself is untouched)
    acc = {"t": 0.0, "t0": 0.0}
    rfft0, irfft0 = torch.fft.rfft, torch.fft.irfft

    def rfft_timed(*a, **k):
        acc["t0"] = time.perf_counter()
        return rfft0(*a, **k)

    def irfft_timed(*a, **k):
        r = irfft0(*a, **k)
        acc["t"] += time.perf_counter() - acc["t0"]
        return r

    full, hot = [], []
    with torch.no_grad():
        blk(x)
        torch.fft.rfft, torch.fft.irfft = rfft_timed, irfft_timed
        try:
            for _ in range(reps):
                acc["t"] = 0.0
                t0 = time.perf_counter()
                blk(x)
                full.append(time.perf_counter() - t0)
                hot.append(acc["t"])
        finally:
            torch.fft.rfft, torch.fft.irfft = rfft0, irfft0
    tok = batch * SEQ
    return {"kind": "reference", "source": "baseline/_ref/spectre.py (unmodified copy of the reference)",
            "sample": f"SpectreBlock({D_MODEL}, {HEADS}, {SEQ}, pooling_type='mean', wavelet_on_rate=0.0) fp32, B={batch}, "
                      f"{reps} reps after 1 warm-up, median",
            "block_forward_tokens_per_s": tok / statistics.median(full),
            "rfft_to_irfft_tokens_per_s": tok / statistics.median(hot), "cores": torch.get_num_threads()}


def run_reference(args):
    """--impl reference: the CPU implementation of the path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import spectre_mix_oracle as oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    batch = 8
    V, gate = _cpu_inputs(torch, batch)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            oracle.mix_head_loop(V, gate, SEQ, HEADS)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle.mix_head_loop(V, gate, SEQ, HEADS)
        dt = time.perf_counter() - t0
    value = batch * SEQ * args.steps / dt
    sample = (f"B={batch} rows of seq={SEQ} d={D_MODEL} fp32 per step, a bounded sample of the {args.global_batch}-row workload "
              f"(CPU throughput is batch-insensitive, SURVEY section 6)")
    stock = None
    try:
        stock = stock_reference_timings(torch, batch=2, reps=2)
    except Exception as e:  # noqa: BLE001 -- the stock module is a secondary record
        stock = {"error": repr(e)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"reference CPU path: torch.fft rfft -> gate -> irfft looped over {HEADS} heads "
                               f"(oracle port of spectre.py:506,542-553,712-718), seq={SEQ} d={D_MODEL}, sample of the "
                               f"batch={args.global_batch} workload",
                   "seq_len": SEQ, "d_model": D_MODEL, "heads": HEADS, "gate_groups": NG, "batch_per_step": batch,
                   "global_batch": args.global_batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                         "stock_reference_module": stock},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def gpu_local_cpus(torch, index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None: pinned host buffers allocated from there keep the
    host-buffer leg's PCIe copies off the inter-socket link."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else None
    except Exception:  # noqa: BLE001 -- no sysfs / no NUMA information: leave the affinity alone
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--global-batch", type=int, default=8192, help="rows of the global batch (BASELINE.json configs[4])")
    ap.add_argument("--micro-batch", type=int, default=148,
                    help="rows per kernel launch: 148 rows = 14208 tiles = 96 full waves of the 148 persistent CTAs")
    ap.add_argument("--burst-steps", type=int, default=20, help="launches of the cool-start burst sub-record")
    ap.add_argument("--e2e-batch", type=int, default=148, help="batch rows per GPU per step (host-buffer leg)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (kernel experiments only)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import fft_b200
    from fft_b200 import _lib
    from fft_b200.dist import micro_batches, reduce_measurement, shard_rows
    lib = _lib.load()  # fail loudly when the CUDA library is missing: there is no fallback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's shard of the global batch, resident in HBM (generated on the device from seed + rank)
    row0, row1 = shard_rows(args.global_batch, rank, world)
    rows = row1 - row0
    MB = args.micro_batch
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    V = torch.empty(rows, SEQ, D_MODEL, device=dev)
    gate = torch.empty(rows, NG, F_HALF, dtype=torch.cfloat, device=dev)
    for r0 in range(0, rows, 256):
        V[r0:r0 + 256].normal_(generator=gen)
        torch.view_as_real(gate[r0:r0 + 256]).normal_(0.0, 0.7071067811865476, generator=gen)   # = randn(cfloat)
    nring = 2
    outs = [torch.empty(min(MB, rows), SEQ, D_MODEL, device=dev) for _ in range(nring)]
    micro = micro_batches(rows, MB)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def launch_rows(a, b, o):
        Vm, gm = V[a:b], gate[a:b]
        rc = lib.spectre_mix_fwd(Vm.data_ptr(), 0, Vm.stride(0), Vm.stride(1), gm.data_ptr(), None, 0, o.data_ptr(), 0,
                                 o.stride(0), o.stride(1), b - a, SEQ, SEQ, D_MODEL, D_G, sp)
        _lib.check(rc, "spectre_mix_fwd")

    def one_step():
        for i, (a, b) in enumerate(micro):
            launch_rows(a, b, outs[i % nring])

    # ---- burst sub-record first (cool start, the round-1 headline measurement): 20 launches of one micro-batch each
    burst = None
    nb = min(MB, rows)
    if args.burst_steps > 0 and rows >= 2 * nb:
        for i in range(3):
            launch_rows((i % 2) * nb, (i % 2 + 1) * nb, outs[i % nring])
        barrier()
        time.sleep(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank, 0.002) as bclk:
            torch.cuda.synchronize()
            e0.record(stream)
            for i in range(args.burst_steps):
                launch_rows((i % 2) * nb, (i % 2 + 1) * nb, outs[i % nring])
            e1.record(stream)
            torch.cuda.synchronize()
        bms = e0.elapsed_time(e1) / args.burst_steps
        burst = {"ms_per_launch": bms, "launches": args.burst_steps, "rows_per_launch": nb, "clocks": bclk.summary()}
        time.sleep(1.0)

    # ---- the timed region: K passes over the global batch
    for _ in range(args.warmup):
        one_step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            one_step()
        ev1.record(stream)
        barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    checksum = float(sum(o.double().sum() for o in outs))
    # whole-job view: max time over ranks, total tokens, summed output checksum (no data-path collective exists)
    elapsed_ms, _, checksum = reduce_measurement(elapsed_ms, rows * SEQ * args.steps, checksum, device=dev)
    tokens_per_step = args.global_batch * SEQ
    value = tokens_per_step * args.steps / (elapsed_ms * 1e-3)
    ms_per_step = elapsed_ms / args.steps

    # ---- parity of what was just timed: rows of the last two micro-batches (still in the output ring) against the oracle
    parity = None
    if rank == 0:
        from oracle import spectre_mix_oracle as oracle
        picks = []
        for back in (1, 2):
            if len(micro) >= back:
                i = len(micro) - back
                a, b = micro[i]
                n = b - a
                picks += [(i, a, j) for j in sorted({0, n // 3, (2 * n) // 3, n - 1})]
        picks = picks[:8] if len(picks) >= 8 else picks
        worst = 0.0
        for (i, a, j) in picks:
            want = oracle.mix_head_loop(V[a + j:a + j + 1].cpu(), gate[a + j:a + j + 1].cpu(), SEQ, HEADS)
            got = outs[i % nring][j:j + 1].cpu()
            worst = max(worst, float((got - want).norm() / want.norm()))
        parity = {"rows_checked": len(picks), "max_rel_l2": worst, "tolerance": 1e-5, "ok": bool(worst <= 1e-5),
                  "against": "oracle.mix_head_loop (torch.fft head loop, spectre.py:506,542-553,712-718) on the same rows",
                  "where": "rows of the last two micro-batches of the timed region, read from the output ring afterwards"}

    # roofline of the (only) kernel: every launch of the step is this kernel
    peak, peak_src = measured_peak()
    alg_step = algorithmic_bytes(rows)                    # this rank's bytes per step
    achieved = alg_step / (ms_per_step * 1e-3) / 1e9
    tpt = ncu_traffic_per_token()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if tpt is None else tpt * min(MB, rows) * SEQ, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algorithmic_bytes(min(MB, rows)), "launches_per_step_per_gpu": len(micro),
                "avg_launch_ms": ms_per_step / len(micro), "kernel": "spx::spectre_mix_kernel",
                "timed_region_s": elapsed_ms * 1e-3,
                "note": "frac is over the whole timed region (seconds of back-to-back launches: sustained, power-capped when "
                        "clocks.reasons says so); burst = the same kernel, 20 launches after a cool-down"}
    if burst is not None:
        b_ach = algorithmic_bytes(nb) / (burst["ms_per_launch"] * 1e-3) / 1e9
        roofline["burst"] = {"achieved": b_ach, "frac": b_ach / peak, **burst}

    # ---- e2e: host buffers through the C-ABI host entry (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        Be = args.e2e_batch
        # allocate (first-touch) and drive the host buffers from the GPU's own NUMA node when sysfs tells which one that is
        saved_affinity = os.sched_getaffinity(0)
        near = gpu_local_cpus(torch, local_rank)
        if near is not None:
            os.sched_setaffinity(0, near[1])
        hV = torch.randn(Be, SEQ, D_MODEL, generator=torch.Generator().manual_seed(100 + rank)).pin_memory()
        hg = torch.randn(Be, NG, F_HALF, dtype=torch.cfloat, generator=torch.Generator().manual_seed(200 + rank)).pin_memory()
        ho = torch.empty(Be, SEQ, D_MODEL).pin_memory()
        fft_b200.spectral_mix_host(hV, hg, n_fft=SEQ, group_width=D_G, out=ho)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fft_b200.spectral_mix_host(hV, hg, n_fft=SEQ, group_width=D_G, out=ho)
            _ = float(ho[0, 0, 0])  # the step's result is read on the host
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        # the box's ceiling for this leg: the same bytes as plain pinned copies, both directions at once, no kernel
        dVe, dOe = torch.empty(Be, SEQ, D_MODEL, device=dev), torch.empty(Be, SEQ, D_MODEL, device=dev)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        barrier()
        t1 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(s_up):
                dVe.copy_(hV, non_blocking=True)
            with torch.cuda.stream(s_dn):
                ho.copy_(dOe, non_blocking=True)
        torch.cuda.synchronize()
        copy_s = (time.perf_counter() - t1) / 2
        del dVe, dOe
        te = torch.tensor([e2e_s, copy_s], device=dev, dtype=torch.float64)
        per_rank = [te.clone() for _ in range(world)]
        if dist is not None:
            dist.all_gather(per_rank, te)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = Be * SEQ * world * args.e2e_steps / float(te[0].item())
        os.sched_setaffinity(0, saved_affinity)
        h2d, d2h = int(hV.numel() * 4 + hg.numel() * 8), int(ho.numel() * 4)
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "batch_per_gpu": Be, "steps": args.e2e_steps,
               "entry": "spectre_mix_fwd_host (C ABI, pinned host buffers, ramped chunks of H2D/kernel/D2H on 4 streams)",
               "per_rank_tokens_per_s": [Be * SEQ * args.e2e_steps / float(t[0].item()) for t in per_rank],
               "copy_ceiling": {"what": "the same V and out bytes as plain pinned cudaMemcpyAsync, H2D and D2H at once on two "
                                        "streams, all ranks together, no kernel",
                                "tokens_per_s": Be * SEQ * world / float(te[1].item()),
                                "per_rank_GBps_each_way": [hV.numel() * 4 / float(t[1].item()) / 1e9 for t in per_rank]},
               "host_numa_node": None if near is None else near[0]}
        del hV, hg, ho

    # ---- CPU baseline beside it (rank 0 at N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_reference_path(torch, batch=8, min_seconds=args.cpu_seconds, max_reps=200)
        best = min(r["times"])
        cpu = {"value": r["tokens"] / statistics.median(r["times"]), "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"B=8 rows of seq={SEQ} d={D_MODEL} fp32, {len(r['times'])} reps after 1 warm-up, median "
                         f"(best {r['tokens'] / best:.3e}); oracle/spectre_mix_oracle.py mix_head_loop = spectre.py:506,"
                         f"542-553 looped over {HEADS} heads as :712-718"}
        try:
            cpu["stock_reference_module"] = stock_reference_timings(torch)
        except Exception as e:  # noqa: BLE001 -- secondary record
            cpu["stock_reference_module"] = {"error": repr(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[4]: batch={args.global_batch} seq={SEQ} d={D_MODEL} fp32 batch-sharded over "
                                   f"{world} GPU(s); fused rFFT->gate->irFFT kernel, all {HEADS} heads / {NG} gate groups per launch, "
                                   f"micro-batches of {MB} rows streamed through {nring} reused output buffers, no spectral memory",
                       "global_batch": args.global_batch, "rows_per_gpu": rows, "micro_batch": MB, "seq_len": SEQ,
                       "d_model": D_MODEL, "heads": HEADS, "gate_groups": NG,
                       "parallelism": f"batch-shard x{world} (shard_rows; no data-path collective)",
                       "l2": f"inputs larger than L2: the rank's shard of V is {rows * SEQ * D_MODEL * 4 / 1e9:.1f} GB resident in HBM, "
                             f"every micro-batch reads {min(MB, rows) * SEQ * D_MODEL * 4 / 1e6:.0f} MB of it once per step",
                       "launch": "programmatic dependent launch: each micro-batch's set-up overlaps the previous launch's tail "
                                 "(griddepcontrol.wait before any tensor access)",
                       "plan": fft_b200.plan_info(min(MB, rows), SEQ, SEQ, D_MODEL, D_G)},
            "roofline": roofline, "parity_check": parity, "e2e": e2e, "cpu_baseline": cpu, "clocks": clk.summary(),
            "gpu_launches": args.steps * len(micro) * world, "checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        print(f"bench.py: parity check FAILED: rel-L2 {parity['max_rel_l2']:.3e} > 1e-5", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
