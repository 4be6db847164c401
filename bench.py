#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

A step = one pass of the hot path over one batch: compress the 1 GiB synthetic mixed-compressibility
buffer (16384 x 64 KiB fragments, BASELINE config 2) into one Snappy stream, then uncompress it.
  value : uncompressed GB / (t_compress + t_uncompress), inputs resident in HBM, CUDA-event timed,
          max over ranks; compress_gbps / uncompress_gbps are the two directions separately.
  e2e   : the same through the reference-facing host-buffer C ABI (snappy_b200_compress /
          snappy_b200_uncompress) from pinned host memory, H2D + D2H inside the timed region.
  roofline : the dominant kernel's (N + C) algorithmic bytes / its CUDA-event duration vs measured HBM.
  cpu_baseline : the oracle (C restatement of Snappy.jl) on this box's host cores.
`--impl reference` times the oracle alone (the reference is Julia and cannot run in this image).
N > 1 (torchrun): weak scaling -- N streams of 1 GiB, every stream sharded over the N ranks along
whole-fragment boundaries, sizes exchanged by NCCL all-gather, segments assembled over NVLink.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GIB = 1 << 30
FRAGMENT = 65536
METRIC = "compress+uncompress GB/s (1 GiB synth mixed buffer; uncompressed bytes / (t_compress + t_uncompress))"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fragments", type=int, default=16384, help="fragments per GPU (16384 = 1 GiB)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_local_cpus(dev_index):
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs), or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        return cpus or None
    except Exception:
        return None


class gpu_local_node:
    """Pinned host buffers of the e2e leg are allocated while this thread is confined to the GPU-local CPUs, so
    that first touch puts them on the GPU's NUMA node (a remote node costs ~30 % of PCIe throughput: 69 vs 52 ms
    per step, bimodal from run to run).  The affinity is restored on exit: the CPU baseline uses every core."""

    def __init__(self, dev_index):
        self.dev, self.old, self.local = dev_index, None, False

    def __enter__(self):
        try:
            cpus = gpu_local_cpus(self.dev)
            old = os.sched_getaffinity(0)
            if cpus and (cpus & old) and (cpus & old) != old:
                os.sched_setaffinity(0, cpus & old)
                self.old = old
            self.local = bool(cpus and (cpus & old))
        except Exception:
            self.old = None
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except Exception:
                pass
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
        return out


def make_input(nfrag, seed):
    from snappy_jl_b200 import synth
    return synth.mix(nfrag, seed=seed)


# ------------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline of the b200 arm, and the whole --impl reference arm
# ------------------------------------------------------------------------------------------
def oracle_roundtrip_threads(raw, threads):
    """T host threads over T independent buffers (the Threads.@threads analogue of SURVEY 8(d)):
    each compresses and uncompresses its own slice as an independent stream.  Returns seconds
    (compress, uncompress).  ctypes releases the GIL, so the threads run in parallel."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    pyoracle.build()
    nfrag = (raw.size + FRAGMENT - 1) // FRAGMENT
    per = (nfrag + threads - 1) // threads
    slices = [raw[i * per * FRAGMENT: min((i + 1) * per * FRAGMENT, raw.size)] for i in range(threads)]
    slices = [s for s in slices if s.size]
    comp = [None] * len(slices)
    back = [None] * len(slices)

    def run(fn):
        ts = [threading.Thread(target=fn, args=(i,)) for i in range(len(slices))]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    def c(i):
        comp[i] = pyoracle.compress_np(slices[i])

    def u(i):
        back[i] = pyoracle.uncompress_np(comp[i])

    tc = run(c)
    tu = run(u)
    assert all(np.array_equal(b, s) for b, s in zip(back, slices)), "oracle round trip failed"
    return tc, tu, sum(int(x.size) for x in comp)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # bounded sample: 256 MiB of the same workload per step (the full 1 GiB x (W+K) steps would take minutes)
    nfrag = min(args.fragments, 4096)
    raw = make_input(nfrag, args.seed)
    for _ in range(args.warmup):
        oracle_roundtrip_threads(raw, cores)
    tc = tu = 0.0
    for _ in range(args.steps):
        a, b, csize = oracle_roundtrip_threads(raw, cores)
        tc += a
        tu += b
    n = raw.size * args.steps
    value = n / (tc + tu) / 1e9
    sample = "%d MiB prefix of the mix (seed %d) per step, %d independent buffers on %d threads" % (
        raw.size >> 20, args.seed, cores, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": (tc + tu) / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "compress_gbps": n / tc / 1e9, "uncompress_gbps": n / tu / 1e9,
        "config": {"workload": "1 GiB synthetic mixed-compressibility buffer (16384 x 64 KiB fragments), "
                               "compress+uncompress; reference arm runs a bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Julia (no Julia in this image): this is oracle/snappy_oracle.c, the C "
                "restatement of Snappy.jl, on host cores",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import snappy_jl_b200 as Snappy
    from snappy_jl_b200 import device, multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1

    nfrag = args.fragments
    n = nfrag * FRAGMENT

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs ---------------------------------------------------------------------------
    if world == 1:
        raw = make_input(nfrag, args.seed)
        with gpu_local_node(local) as numa:
            host_in = torch.from_numpy(raw).pin_memory()
        d_in = host_in.to(dev, non_blocking=True)
        shards = None
    else:
        # stream s = concat over ranks r of mix(nfrag/world, seed(s, r)); this rank holds run r=rank of every s
        per = nfrag // world
        shards, totals = [], []
        for s in range(world):
            part = make_input(per, args.seed + 1000 * s + rank)
            shards.append(torch.from_numpy(part).to(dev))
            totals.append(per * world * FRAGMENT)
        raw = None
    torch.cuda.synchronize()

    cap = Snappy.maxlength_compressed(n)
    codec = multi.CudaCodec()
    launches = 0

    if world == 1:
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_back = torch.empty(n, dtype=torch.uint8, device=dev)

        def step_compress():
            return device.compress_device(d_in, out=d_out, want_index=True)

        def step_uncompress(stream, index):
            return device.uncompress_device(stream, out=d_back, index=index, claimed=n)
    else:
        def step_compress():
            return multi.compress_streams(shards, totals, codec)

        def step_uncompress(stream, index):
            return multi.uncompress_streams(stream, index, totals[rank], codec)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(args.warmup):
        stream, index = step_compress()
        step_uncompress(stream, index)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("SNAPPY_BENCH_NO_SAMPLER"):
        sampler.start()
    tc = tu = 0.0
    kc = ku = 0.0
    csize = 0
    for _ in range(args.steps):
        barrier()
        ev[0].record()
        stream, index = step_compress()
        ev[1].record()
        kc += device.last_kernel_ms(0)
        launches += device.last_launch_count(0) * (world if world > 1 else 1)
        back = step_uncompress(stream, index)
        ev[2].record()
        ku += device.last_kernel_ms(1)
        launches += device.last_launch_count(1) * (world if world > 1 else 1)
        barrier()
        a, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        if world > 1:
            t = torch.tensor([a, b], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            a, b = float(t[0]), float(t[1])
        tc += a
        tu += b
        csize = int(stream.numel())
    clocks = sampler.stop() if rank == 0 else None
    # correctness of what was timed
    if world == 1:
        assert torch.equal(back, d_in), "round trip mismatch in the timed configuration"
    else:
        for s in range(world):
            assert torch.equal(back[s], shards[s]), "round trip mismatch (stream %d)" % s

    total_bytes = n * (world if world > 1 else 1)  # uncompressed bytes all ranks processed per step
    t_step_ms = (tc + tu) / args.steps
    value = total_bytes * args.steps / ((tc + tu) / 1e3) / 1e9
    comp_gbps = total_bytes * args.steps / (tc / 1e3) / 1e9
    unc_gbps = total_bytes * args.steps / (tu / 1e3) / 1e9

    # ---- roofline of the dominant kernel (per launch; N + C algorithmic bytes) ----------------
    peak, peak_src = measured_peak()
    kc_ms, ku_ms = kc / args.steps, ku / args.steps
    if world > 1:
        # per-rank kernels each cover 1/world of a stream; report the compress shard kernel of the last call
        alg_c = (shards[0].numel() + csize / world)
        alg_u = alg_c
    else:
        alg_c = alg_u = n + csize
    dominant = "compress" if kc_ms >= ku_ms else "uncompress"
    k_ms = kc_ms if dominant == "compress" else ku_ms
    achieved = (alg_c / (k_ms / 1e3) / 1e9) if k_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak if peak else None, "traffic": None,
        "kernel": "k_compress_window" if dominant == "compress" else "k_decode_fragments",
        "kernel_ms": k_ms, "algorithmic_bytes": alg_c, "peak_source": peak_src,
        "other": {"compress_kernel_ms": kc_ms, "uncompress_kernel_ms": ku_ms,
                  "uncompress_achieved": (alg_u / (ku_ms / 1e3) / 1e9) if ku_ms > 0 else None},
    }
    # DRAM traffic per 1 GiB launch from the committed ncu --set full captures (profiles/traffic.json).
    # k_compress_window runs as TWO concurrent kernels that share the fragments (shared-memory tables /
    # global tables); ncu serialises kernels, so each was captured doing the whole 1 GiB alone and the
    # launch's traffic is their mix by share of fragments -- reported as an estimate, next to the captures.
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
        ent = tr.get(roofline["kernel"])
        if isinstance(ent, dict):
            roofline["traffic"] = ent.get("per_launch_estimate")
            roofline["traffic_captures"] = ent
        else:
            roofline["traffic"] = ent
    except Exception:
        pass

    # ---- e2e through the host-buffer C ABI (rank-local stream at N > 1 is not defined: N = 1 only) ----
    e2e = None
    if world == 1 and not args.no_e2e:
        with gpu_local_node(local) as numa:
            h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
            h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
        import ctypes
        lib = Snappy._abi.lib()

        def e2e_step():
            ol = ctypes.c_size_t(cap)
            rc = lib.snappy_b200_compress(host_in.data_ptr(), n, h_out.data_ptr(), ctypes.byref(ol))
            assert rc == 0, rc
            bl = ctypes.c_size_t(n)
            rc = lib.snappy_b200_uncompress(h_out.data_ptr(), ol.value, h_back.data_ptr(), ctypes.byref(bl))
            assert rc == 0 and bl.value == n, rc
            return ol.value

        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 3))
        for _ in range(reps):
            c_len = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        assert torch.equal(h_back, host_in)
        e2e = {"value": n / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": n + c_len,
               "d2h_bytes_per_step": c_len + n, "ms_per_step": dt * 1e3,
               "api": "snappy_b200_compress + snappy_b200_uncompress on pinned host buffers",
               "pinned_on_gpu_numa_node": numa.local}
        launches_e2e = device.last_launch_count(0) + device.last_launch_count(1)
    elif world > 1:
        e2e = None

    # ---- cpu_baseline: the oracle on host cores, rank 0 at N = 1 only -------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_frag = min(nfrag, 4096)
        sample = raw[: sample_frag * FRAGMENT]
        a1, b1, _ = oracle_roundtrip_threads(sample, 1)
        aT, bT, _ = oracle_roundtrip_threads(sample, cores)
        cpu = {"value": sample.size / (aT + bT) / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": "first %d MiB of the same buffer, %d independent buffers on %d threads" % (
                   sample.size >> 20, cores, cores),
               "single_thread": {"value": sample.size / (a1 + b1) / 1e9,
                                 "compress_gbps": sample.size / a1 / 1e9,
                                 "uncompress_gbps": sample.size / b1 / 1e9},
               "compress_gbps": sample.size / aT / 1e9, "uncompress_gbps": sample.size / bT / 1e9}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "compress_gbps": comp_gbps, "uncompress_gbps": unc_gbps,
            "compressed_ratio": csize / (n if world == 1 else totals[rank]),
            "config": {"workload": "1 GiB synthetic mixed-compressibility buffer (%d x 64 KiB fragments) per GPU, "
                                   "compress then uncompress, bit-exact vs Snappy.jl restatement" % nfrag,
                       "bytes_per_gpu": n, "seed": args.seed,
                       "l2": "inputs (1 GiB) larger than L2 (126 MB); no explicit flush",
                       "sharding": "none" if world == 1 else
                       "%d streams x %d ranks, whole-fragment runs, NCCL size all-gather + all-to-all assembly" % (world, world)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
