#!/usr/bin/env python
"""Benchmark of the Pauli-frame sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = one pass of the sampler over 2^24 shots (per GPU) of the d=25 r=25 p=1e-3 rotated surface
code memory-Z circuit (BASELINE.json configs[2], fixture tests/golden/circuits/c3_surface_z_d25_r25.stim),
producing bit-packed (b8) detection events + observables.

  value  device-resident throughput: results land in a preallocated HBM buffer (no PCIe in the timed region)
  e2e    same metric through the public API with HOST output buffers, D2H inside the timed region: the headline number is
         the page-locked caller buffer (direct DMA); the default pageable numpy path (sample(bit_packed=True)) and the
         bare D2H copy ceiling of the same bytes are reported beside it
  --impl reference   the unmodified reference CLI (oracle/_ref/stim detect, built from /root/reference by
                     oracle/Makefile) timed on the host cores, one process per core, 2^17 shots per process and step.

On N > 1 GPUs every rank samples its own shot range (no data-path collective); after the timed region the path's one
collective is exercised on hardware: per-detector flip counts of BASELINE config 4 are summed with NCCL and compared with a
single-sampler run of the same global shot range (stim_b200.sharding.check_shard_invariance).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CIRCUIT = os.path.join(ROOT, "tests", "golden", "circuits", "c3_surface_z_d25_r25.stim")
REF_STIM = os.path.join(ROOT, "oracle", "_ref", "stim")
WORKLOAD = "surface_code:rotated_memory_z d=25 rounds=25 p=1e-3 (all four noise knobs), b8 detection events + observables"
METRIC = "detector shots/s, rotated surface code d=25 r=25 p=1e-3"
ALG_BYTES_PER_SHOT = 1951       # ceil((15600 detectors + 1 observable) / 8): result bytes written once
ALG_LOP3_PER_SHOT = 168026 / 32  # XOR word-ops of the frame algorithm per shot
LOP3_LANES_PER_CLK_PER_SM = 64   # fallback only (B300_MICROARCH.md: alu pipe rt_SMSP = 2 -> 16 lanes/clk/SMSP); bench measures it
C4_CIRCUIT = os.path.join(ROOT, "tests", "golden", "circuits", "c4_color_d15_r15.stim")
# the other BASELINE.json configurations (--config): parity-test cases, benchable on request
CONFIGS = {
    "c1": ("c1_rep_d3_r10.stim", "repetition_code:memory d=3 rounds=10 p=1e-3"),
    "c2": ("c2_surface_x_d5_r5.stim", "surface_code:rotated_memory_x d=5 rounds=5 p=1e-3"),
    "c3": ("c3_surface_z_d25_r25.stim", None),
    "c4": ("c4_color_d15_r15.stim", "color_code:memory_xyz d=15 rounds=15 p=1e-3"),
    "c4v": ("c4v_color_d15_r15_mpp_dense.stim", "color_code:memory_xyz d=15 rounds=15 with MPP and dense noise"),
    "c5": ("c5_surface_x_d51_r51.stim", "surface_code:rotated_memory_x d=51 rounds=51 p=1e-3 (5201 qubits)"),
}


def ncu_capture(engine):
    """The committed `ncu --set full` summary of the dominant kernel of `engine` (profiles/r2_events_full.json for the event
    engine, the interpreter's otherwise). Returns (dict or None, file name)."""
    names = ("r2_events_full.json",) if engine == "events" else ("r2_interp_full.json", "r1_interp_full.json")
    for name in names:
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f), name
        except Exception:
            continue
    return None, None


def bind_to_gpu_numa_node(torch, index):
    """One process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned result buffer (first touch) and the
    PCIe traffic stay on the socket the GPU hangs off. Best effort; returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region: NVML polled every ~2 ms from a thread of this process
    (the timed region of a multi-GPU run is ~90 ms, shorter than one `nvidia-smi -lms` period when eight of them start at
    once); `nvidia-smi` is the fallback when pynvml is missing."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            self.index = int(vis.split(",")[index]) if vis else index
        except (ValueError, IndexError):
            self.index = index
        self.samples = []   # nvidia-smi lines
        self.nvml = []      # (sm_mhz, reason bits)
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.max_mhz = None
        self.source = None

    def _poll_nvml(self, pynvml, handle):
        while not self.stop_flag:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                try:
                    bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.nvml.append((float(sm), int(bits)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.source == "nvml":
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(v for v, _ in self.nvml)
            bits = 0
            for _, b in self.nvml:
                bits |= b
            reasons = sorted(name for mask, name in self.NVML_REASONS if bits & mask)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


def run_reference(shots_per_proc, nproc, seed0):
    """One bounded sample of the reference CPU sampler: nproc processes x shots_per_proc shots. Returns seconds."""
    t0 = time.perf_counter()
    procs = [
        subprocess.Popen([REF_STIM, "detect", "--shots", str(shots_per_proc), "--in", CIRCUIT, "--out_format", "b8",
                          "--append_observables", "--out", "/dev/null", "--seed", str(seed0 + i)])
        for i in range(nproc)
    ]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference stim detect failed")
    return time.perf_counter() - t0


def reference_arm(args, rank, world):
    if rank != 0:
        return
    nproc = os.cpu_count() or 1
    shots_per_proc = 1 << 17  # >= 2^17 shots per process, so process start + circuit parsing (< 30 ms) are amortised
    if not os.path.exists(REF_STIM):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/stim is not built (run make -C oracle ref)"}))
        return
    for w in range(args.warmup):
        run_reference(shots_per_proc, nproc, 1000 * w)
    t = 0.0
    for k in range(args.steps):
        t += run_reference(shots_per_proc, nproc, 12345 + 1000 * k)
    total = shots_per_proc * nproc * args.steps
    value = total / t
    sample = f"{nproc} processes x {shots_per_proc} shots per step, stim detect --out_format b8 --append_observables (W=256 AVX2 build)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bit-sliced (AVX2 256-bit words)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "shots_per_step": shots_per_proc * nproc},
        "cpu_baseline": {"value": value, "unit": "shots/s", "cores": nproc, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shots-log2", type=int, default=24)
    ap.add_argument("--e2e-shots-log2", type=int, default=24)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the headline c3)")
    ap.add_argument("--engine", default="auto", choices=["auto", "interp", "events"],
                    help="sampling engine (include/gstim.h); auto = the library's own choice (the event engine for this circuit)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    global CIRCUIT, WORKLOAD, METRIC, ALG_BYTES_PER_SHOT, ALG_LOP3_PER_SHOT
    if args.config != "c3":
        fname, wl = CONFIGS[args.config]
        CIRCUIT = os.path.join(ROOT, "tests", "golden", "circuits", fname)
        WORKLOAD = wl + ", b8 detection events + observables"
        METRIC = "detector shots/s, " + wl
        ALG_BYTES_PER_SHOT = None  # from the circuit below
        ALG_LOP3_PER_SHOT = float("nan")
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch

    import stim_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: stim_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with open(CIRCUIT) as f:
        text = f.read()
    circuit = stim_b200.Circuit(text)
    sampler = circuit.compile_detector_sampler(seed=12345, device=local_rank, engine=args.engine)
    sampler.shot_offset = rank << 44  # disjoint Philox counter ranges per GPU; no inter-GPU traffic
    D, L = circuit.num_detectors, circuit.num_observables
    nbytes = (D + L + 7) // 8
    if ALG_BYTES_PER_SHOT is None:
        ALG_BYTES_PER_SHOT = nbytes
    assert nbytes == ALG_BYTES_PER_SHOT
    shots = 1 << args.shots_log2
    while shots * nbytes > (40 << 30):  # (c5: 16.6 KB per shot)
        shots >>= 1

    # ---- device-resident arm -----------------------------------------------------------------
    out = torch.empty((shots, nbytes), dtype=torch.uint8, device="cuda")
    for _ in range(args.warmup):
        sampler.sample_device(shots, out.data_ptr(), append_observables=True)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    dev_ms = interp_ms = transpose_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        sampler.sample_device(shots, out.data_ptr(), append_observables=True)
        dev_ms += sampler.last_call_ms()
        a, b = sampler.last_kernel_ms()
        interp_ms += a
        transpose_ms += b
        launches += sampler.last_launch_count()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clk = clocks.stop()
    info = sampler.engine_info()
    engine = info["last_engine"]
    # interpreter: one interpreter + one transposer launch per chunk; event engine: one launch per call
    interp_launches = (launches // 2 if engine == "interp" else launches) or 1
    ms_per_step = max_over_ranks(dev_ms / args.steps)
    wall_ms_per_step = max_over_ranks(wall_ms / args.steps)
    value = world * shots / (ms_per_step * 1e-3)
    # a cheap size-independent sanity property on the full-size output: detection fraction ~1.8 %
    frac = float(torch.count_nonzero(out[:4096]).item()) / (4096 * nbytes)
    del out
    torch.cuda.empty_cache()

    # ---- end-to-end arm (host buffers, D2H inside the timed region) ----------------------------------
    # Headline: the caller's page-locked buffer at the same 2^24 shots per step when the host can pin 32.7 GB per rank
    # (else the largest power of two it can; every rank uses the same size, so the decision is taken together).
    e2e_log2 = min(args.e2e_shots_log2, shots.bit_length() - 1)
    try:  # never pin more than 40 % of the host memory that is available, all ranks together
        with open("/proc/meminfo") as f:
            avail = next(int(ln.split()[1]) * 1024 for ln in f if ln.startswith("MemAvailable"))
        while e2e_log2 > 16 and world * (nbytes << e2e_log2) > 0.4 * avail:
            e2e_log2 -= 1
    except (OSError, StopIteration, ValueError):
        pass
    if world > 1:
        t = torch.tensor([float(e2e_log2)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        e2e_log2 = int(t.item())
    host = None
    while host is None:
        try:
            host = torch.empty((1 << e2e_log2, nbytes), dtype=torch.uint8, pin_memory=True)
            ok = 1.0
        except RuntimeError:
            ok = 0.0
        if world > 1:
            t = torch.tensor([ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = float(t.item())
        if ok == 0.0:
            host = None
            torch.cuda.empty_cache()
            e2e_log2 -= 1
            if e2e_log2 < 16:
                raise RuntimeError("cannot allocate a pinned host buffer for the end-to-end arm")
    e2e_shots = 1 << e2e_log2
    host_np = host.numpy()
    sampler.sample(e2e_shots, bit_packed=True, append_observables=True, dets_out=host_np)  # warm-up (allocations)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sampler.sample(e2e_shots, bit_packed=True, append_observables=True, dets_out=host_np)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * e2e_shots * e2e_steps / e2e_s
    # the bare D2H copy of the same bytes into the same buffer: the ceiling the end-to-end number can reach
    dev_src = torch.empty((min(e2e_shots, 1 << 22), nbytes), dtype=torch.uint8, device="cuda")
    n_cp = dev_src.shape[0]
    host[:n_cp].copy_(dev_src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        host[:n_cp].copy_(dev_src, non_blocking=True)
    barrier()
    d2h_s = max_over_ranks(time.perf_counter() - t0)
    d2h_gbs = world * 3 * n_cp * nbytes / d2h_s / 1e9
    del dev_src
    # the default API path: sample(bit_packed=True) returns a fresh pageable numpy array (pinned staging + threaded copy)
    page_shots = min(e2e_shots, 1 << 22)
    sampler.sample(page_shots, bit_packed=True, append_observables=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        sampler.sample(page_shots, bit_packed=True, append_observables=True)
    barrier()
    page_s = max_over_ranks(time.perf_counter() - t0)
    page_value = world * page_shots * 2 / page_s
    del host, host_np

    # ---- the path's one collective, on hardware (N > 1): NCCL sum of per-detector flip counts, shard invariance --------
    collective = None
    if dist is not None:
        from stim_b200 import sharding

        with open(C4_CIRCUIT) as f:
            c4 = stim_b200.Circuit(f.read())
        ok = sharding.check_shard_invariance(c4, seed=2026, shots_per_rank=1 << 18, device=local_rank)
        collective = {"op": "all_reduce(SUM) of uint64[D+L] flip counts and adjacent-pair counts, NCCL", "ranks": world,
                      "circuit": "tests/golden/circuits/c4_color_d15_r15.stim", "shots_per_rank": 1 << 18,
                      "sum_over_ranks_equals_single_sampler": bool(ok)}

    # the other engine on the same workload, 2^22 shots (kept beside the headline for continuity with round 1)
    other = None
    try:
        other_engine = "interp" if engine == "events" else "events"
        s2 = circuit.compile_detector_sampler(seed=999, device=local_rank, engine=other_engine)
        n2 = min(shots, 1 << 22)
        out2 = torch.empty((n2, nbytes), dtype=torch.uint8, device="cuda")
        best = None
        for _ in range(3):
            s2.sample_device(n2, out2.data_ptr(), append_observables=True)
            best = s2.last_call_ms() if best is None else min(best, s2.last_call_ms())
        other = {"engine": other_engine, "value": n2 / (best * 1e-3), "unit": "shots/s (one GPU, device-resident)", "shots": n2}
        del out2, s2
    except ValueError:
        pass

    lop3 = stim_b200.measure_lop3_peak(local_rank)

    peak, peak_src = measured_peak_gbs()
    cap, traffic_file = ncu_capture(engine)
    traffic_per_shot = None if cap is None else (float(cap["dram_bytes_read"]) + float(cap["dram_bytes_write"])) / float(cap["shots"])
    interp_s = interp_ms * 1e-3
    achieved = ALG_BYTES_PER_SHOT * shots * args.steps / interp_s / 1e9
    sm_mhz = clk.get("sm_mhz") or 1965.0
    lanes = lop3["lane_ops_per_clk_per_sm"] or LOP3_LANES_PER_CLK_PER_SM
    lop3_peak = 148 * lanes * sm_mhz * 1e6
    per_gpu_rate_interp = shots * args.steps / interp_s
    line = {
        "metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (bitwise XOR of output words; integer fixed point in the noise gaps)", "data": "synthetic",
        "config": {
            "workload": WORKLOAD, "shots_per_gpu_per_step": shots, "output": f"b8 dets+obs, {nbytes} B/shot, resident in HBM",
            "circuit": os.path.relpath(CIRCUIT, ROOT),
            "l2": f"each step writes {shots * nbytes / 1e9:.1f} GB of fresh output (>> 126 MB L2); nothing is reused across steps",
            "engine": engine,
            "engine_info": {k: info[k] for k in ("tile_shots", "num_sites", "num_entries", "num_slices", "events_per_shot", "flips_per_shot")},
            "threads": 1024 if engine == "events" else int(sampler.stats.threads),
            "columns_per_block": sampler.last_block_columns(),
            "nonzero_byte_fraction_check": frac,
        },
        "gpu_launches": launches,
        "kernel_ms_per_step": ({"events": interp_ms / args.steps} if engine == "events" else
                               {"interp": interp_ms / args.steps, "transpose": transpose_ms / args.steps}),
        "other_engine": other,
        "clocks": clk,
        "e2e": {
            "value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": e2e_shots * nbytes,
            "shots_per_step": e2e_shots, "steps": e2e_steps, "host_buffer": "pinned (caller's page-locked array, direct DMA)",
            "numa_node": numa, "result_gbs": e2e_value * nbytes / 1e9,
            "d2h_copy_ceiling_gbs": d2h_gbs,
            "d2h_copy_ceiling_note": "bare cudaMemcpyAsync of the same bytes into the same pinned buffer on every rank at once: "
                                     "what PCIe / host DRAM allow with no sampling at all",
            "pageable": {"value": page_value, "unit": "shots/s", "shots_per_step": page_shots,
                         "note": "default sample(bit_packed=True): fresh pageable numpy array (pinned staging + threaded copy)"},
        },
        "collective": collective,
        "roofline": {
            "kernel": "gstim_sparse_kernel" if engine == "events" else "gstim_interp_kernel", "bound": "hbm", "achieved": achieved,
            "peak": peak, "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": (None if traffic_per_shot is None else traffic_per_shot * shots * args.steps / interp_launches),
            "traffic_source": f"dram__bytes_read+write of profiles/{traffic_file} scaled to the shots of one launch",
            "peak_source": peak_src,
            "launches": interp_launches, "avg_launch_ms": interp_ms / interp_launches,
            "alg_bytes_per_launch": ALG_BYTES_PER_SHOT * shots * args.steps / interp_launches,
            "alu_bound": {"lop3_per_shot": ALG_LOP3_PER_SHOT, "achieved_lop3_per_s": per_gpu_rate_interp * ALG_LOP3_PER_SHOT,
                          "peak_lop3_per_s": lop3_peak, "frac": per_gpu_rate_interp * ALG_LOP3_PER_SHOT / lop3_peak,
                          "note": "word-ops of the reference's frame algorithm; the event engine does not execute them (it is "
                                  "latency-bound at 32 warps per SM, see issue_bound and profiles/r2_notes.md)" if engine == "events" else "",
                          "peak_basis": f"measured: LOP3 microbenchmark in this run = {lanes:.2f} lanes/clk/SM "
                                        f"({lop3['lane_ops_per_sec'] / 1e12:.2f} T lane-ops/s at {lop3['sm_mhz']:.0f} MHz) x 148 SMs x "
                                        f"{sm_mhz:.0f} MHz sampled during the timed region",
                          "lop3_probe": lop3},
            # the event engine's instruction stream: warp instructions per shot (ncu capture) against 4 issue slots per clock and SM
            "issue_bound": (None if cap is None or engine != "events" else {
                "warp_instructions_per_shot": float(cap["warp_instructions"]) / float(cap["shots"]),
                "achieved_warp_inst_per_s": per_gpu_rate_interp * float(cap["warp_instructions"]) / float(cap["shots"]),
                "peak_warp_inst_per_s": 148 * 4 * sm_mhz * 1e6,
                "frac": per_gpu_rate_interp * float(cap["warp_instructions"]) / float(cap["shots"]) / (148 * 4 * sm_mhz * 1e6),
                "source": f"smsp__inst_executed.sum of profiles/{traffic_file}"}),
        },
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_STIM):
        nproc = os.cpu_count() or 1
        sp = 1 << 17
        t1 = run_reference(sp, 1, 777)
        tn = run_reference(sp, nproc, 888)
        line["cpu_baseline"] = {
            "value": sp * nproc / tn, "unit": "shots/s", "cores": nproc, "kind": "reference",
            "sample": f"oracle/_ref/stim detect b8 --append_observables: {nproc} processes x {sp} shots (one per core); "
                      f"single process: {sp / t1:.0f} shots/s",
            "single_thread_value": sp / t1,
        }
    elif rank == 0 and world == 1:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
