#!/usr/bin/env python
"""bench.py — multichannel subband frames/s through the GSC pipe (BASELINE.json metric).

  python bench.py --gpus 1 --steps 10 --warmup 3                      # our arm, 1 GPU
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # our arm, N GPUs (utterance shards, weak scaling)
  python bench.py --impl reference --gpus 1 --steps K --warmup W      # the reference's own CPU path on the host cores

One "step" = one pass of the hot path (OverSampledDFT analysis -> GSC with NLMS sidelobe canceller -> OverSampledDFT
synthesis) over one batch of synthetic utterances.  Workload at every N: BASELINE.json configs[1] per GPU — 8 mics,
M = 512 subbands (m = 4, r = 1, delay-compensation type 2), 256 synthetic 5 s utterances (SURVEY.md §8d generator),
81 152 frames per GPU per step.  One frame = one beamformed output frame of one utterance.

`value`    device-timed (CUDA events on the pipeline stream), inputs already resident in HBM.
`e2e`      same metric through the public C-ABI with HOST buffers: pinned H2D of the step's samples + delays, the three
           kernels, D2H of the resynthesised signal and statistics inside the timed region.
`roofline` dominant kernel (largest share of the step): algorithmic bytes per frame (SURVEY.md §8d / DESIGN.md §5) x
           frames / its CUDA-event time, against the measured HBM copy peak in MEASURED_PEAKS.json.
`cpu_baseline` the reference's own C++ (oracle/_ref, compiled unmodified; NLMS restated in C++) on a bounded sample of
           the same workload on the host cores.
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

FS = 16000.0
CFG = dict(C=8, M=512, m=4, r=1, U=256, n=80000)
LMS = dict()  # unit_test/confs/gsclms.json defaults (min_frames 128, slowdown_after 4096, ...)
METRIC = "multichannel subband frames/sec through GSC pipe"


def frames_per_utt(n, M, m, r):
    D = M >> r
    R = 1 << r
    return -(-n // D) + m * R // 2


def load_proto(M):
    p = np.load(os.path.join(ROOT, "tests", "golden", "prototype_M%d_m4_r1.npz" % M))
    return p["h"], p["g"]


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.stop_flag = False; self.rows = []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    first, count, C, n, M, m, r = args
    from oracle import ref
    from distant_speech_recognition_b200 import synthetic
    h, g = load_proto(M)
    xs = [synthetic.make_utterance(first + i, C, n, pcm16=True)[:2] for i in range(count)]
    t0 = time.perf_counter()
    frames = 0
    for x, d in xs:
        res = ref.beamform(x, h, g, d, M, m, r, bf_kind=ref.BF_GSC_LMS, do_synthesis=True, want_subband=False)
        frames += int(res["stats"][1])
    return frames, time.perf_counter() - t0


def cpu_reference(utts_per_core, cores, seed0=0):
    """The reference's own CPU implementation of the path on `cores` host cores (one utterance range per process,
    the reference is single-threaded).  Returns (frames/s, frames, seconds, cores)."""
    import multiprocessing as mp
    c = CFG
    jobs = [(seed0 + i * utts_per_core, utts_per_core, c["C"], c["n"], c["M"], c["m"], c["r"]) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
        wall = res[0][1]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
        wall = max(r[1] for r in res)  # processing time only (input synthesis excluded), slowest worker
    frames = sum(r[0] for r in res)
    return frames / wall, frames, wall, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    upc = max(1, CFG["U"] // cores)  # utterances per core per step: one whole configs[1] batch per step (~15 s of CPU work on 16 cores)
    v0, _, _, _ = cpu_reference(1, cores)  # untimed warm-up pass (page-in of the reference library, fork pool), also sizes the sample:
    T0 = frames_per_utt(CFG["n"], CFG["M"], CFG["m"], CFG["r"])
    while upc > 1 and args.steps * (upc * cores * T0 / v0) > 150.0:  # keep the K timed steps within ~2.5 minutes of processing
        upc //= 2
    vals, fr, secs = [], 0, 0.0
    for k in range(args.steps):
        v, f, s, _ = cpu_reference(upc, cores, seed0=k * upc * cores)
        vals.append(v); fr += f; secs += s
    value = fr / secs
    c = CFG
    T = frames_per_utt(c["n"], c["M"], c["m"], c["r"])
    sample = "%d utterances (%d per core x %d cores) of configs[1] per step, %d steps" % (upc * cores, upc, cores, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "xrt": value * (c["n"] / FS) / T,
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    c = CFG
    return {"workload": "configs[1]: 8-mic SubbandGSC (NLMS sidelobe canceller), 512 subbands, batch of 256 synthetic 5 s utterances per GPU",
            "channels": c["C"], "subbands": c["M"], "m": c["m"], "r": c["r"], "utterances_per_gpu": c["U"], "samples_per_utterance": c["n"],
            "frames_per_utterance": frames_per_utt(c["n"], c["M"], c["m"], c["r"]), "global_utterances": c["U"] * n_gpus,
            "parallelism": "utterance shards x%d (no data-path collective)" % n_gpus,
            "l2_policy": "inputs (655 MB samples, 1.33 GB snapshots per step) exceed the 126 MB L2; no explicit flush",
            "resident_input": "float32 [U][C][n] like SampleFeature (16-bit PCM with --i16-input); the e2e arm uploads 16-bit PCM, which the analysis kernel reads directly",
            "kernel_variants": {k: ("packed 2 x fp32" if os.environ.get(k, "1") not in ("", "0") else "scalar")
                                for k in ("BTKB_ANALYSIS_PACKED", "BTKB_PERBIN_PACKED", "BTKB_SYNTHESIS_PACKED")}}


# ----------------------------------------------------------------------------- our arm
def bind_to_gpu_numa_node(props):
    """N > 1 only: run this rank (and the pinned buffers it first-touches) on the CPUs NVML names as local to its GPU, so that the
    ranks' host<->device copies do not all cross the same socket interconnect (at N = 2 the unbound e2e scaled 1.5x while the device
    part scaled 1.99x, profiles/r01f_bench_n2.json).  Best effort: any failure leaves the affinity untouched.  Returns a description
    for the JSON line, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        if not after:
            raise RuntimeError("empty affinity")
        return {"gpu": bus, "cpus": "%d-%d (%d of %d)" % (after[0], after[-1], len(after), before)}
    except Exception:  # noqa: BLE001
        return None


def make_inputs(rank):
    from distant_speech_recognition_b200 import synthetic
    import multiprocessing as mp
    c = CFG
    first = rank * c["U"]
    procs = min(os.cpu_count() or 1, 8)
    chunks = [(first + i * (c["U"] // procs), c["U"] // procs) for i in range(procs)] if c["U"] % procs == 0 else [(first, c["U"])]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        parts = pool.starmap(_gen_chunk, [(a, b, c["C"], c["n"]) for a, b in chunks])
    x = np.concatenate([p[0] for p in parts]); d = np.concatenate([p[1] for p in parts])
    return x, d


def _gen_chunk(first, count, C, n):
    from distant_speech_recognition_b200 import synthetic
    return synthetic.make_batch(count, C, n, first=first, pcm16=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from distant_speech_recognition_b200 import _capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _capi.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = None
    if world > 1:
        # stdout carries the single JSON line: NCCL prints its version banner to stdout when the first communicator is created
        # (whatever NCCL_DEBUG says in this image), so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        numa = bind_to_gpu_numa_node(torch.cuda.get_device_properties(local))
    c = CFG
    U, C, n, M, m, r = c["U"], c["C"], c["n"], c["M"], c["m"], c["r"]
    T = frames_per_utt(n, M, m, r)
    frames_step = U * T
    h, g = load_proto(M)

    x_np, delays = make_inputs(rank)
    x_pin = torch.from_numpy(x_np).pin_memory()
    del x_np
    pipe = _capi.Pipeline(C, M, m, r, beamformer=_capi.BF_GSC_LMS, lms=LMS, max_utterances=U, max_samples=n, device=local)
    pipe.set_prototypes(h, g)
    pipe.set_delays(delays)
    nb = (T - m * (1 << r) // 2) * (M >> r)
    out_pin = torch.empty((U, nb), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs in HBM before the timed region, as float32 — the type SampleFeature holds its samples in
    # (feature/feature.h:193).  --i16-input keeps them as 16-bit PCM instead (what a wav file holds; the analysis kernel reads it directly,
    # same snapshots bit for bit, same kernel time, fewer algorithmic bytes).
    x16_pin = x_pin.to(torch.int16).pin_memory()   # exact: the synthetic samples sit on the int16 grid
    if args.i16_input:
        pipe.submit_i16_pointer(x16_pin.data_ptr(), U, n)
    else:
        pipe.submit_pointer(x_pin.data_ptr(), U, n)
    pipe.synchronize()
    for _ in range(max(args.warmup, 3)):
        pipe.run(True)
    pipe.synchronize()
    sampler = ClockSampler(local); sampler.start()
    ks = {"analysis_ms": 0.0, "perbin_ms": 0.0, "synthesis_ms": 0.0}
    launches = 0
    barrier()
    tot_ms = 0.0
    for _ in range(args.steps):
        pipe.run(True)
        t = pipe.last_timing()  # CUDA events recorded on the pipeline stream around the step and around each kernel
        tot_ms += t["total_ms"]; launches += t["launches"]
        for k in ks:
            ks[k] += t[k]
    barrier()
    # ---- end-to-end arm through the package's multi-GPU front end (btk20.batch.ShardedBatchBeamformer -> C-ABI) with HOST buffers:
    # 16-bit PCM samples in pinned memory (what the reference's SampleFeature reads from wav files), NP sub-batches on NP pipeline
    # handles so that the H2D copy of one sub-batch overlaps the kernels / D2H of the previous one.  Every step uploads all samples
    # + delays and downloads the resynthesised signal + statistics.
    from distant_speech_recognition_b200.btk20.batch import ShardedBatchBeamformer
    NP = 8 if U % 8 == 0 else (4 if U % 4 == 0 else 1)   # 8 sub-batches: 6.3 ms/step, 4: 7.7 ms (tools/dbg/e2e_probe.py; PCIe alone: 5.9 ms)
    sbb = ShardedBatchBeamformer(C, h, g, M, m, r, FS, beamformer=dict(LMS, type="gsclms"), utterances=U, max_samples=n, sub_batches=NP, device=local,
                                 world=world, rank=rank)
    row_bytes = out_pin.shape[1] * 4

    def e2e_step():
        sbb.step(x16_pin.data_ptr(), delays, out_pin.data_ptr(), row_bytes)

    for _ in range(2):
        e2e_step()
    sbb.drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sbb.drain()   # every step's results are on the host before the clock stops
    barrier()
    e2e_s = time.perf_counter() - t0
    launches_e2e = sbb.launches()
    stats = sbb.stats.copy()
    # ---- copy-only ceiling of the same step on the same buffers: what the host <-> device links of this box deliver when every rank
    # moves its int16 samples up and its float signal down at once, with no kernel in between (N > 1: all ranks at the same time)
    xdev = torch.empty((U // NP, C, n), dtype=torch.int16, device="cuda")
    ydev = torch.zeros((U // NP, out_pin.shape[1]), dtype=torch.float32, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    Usb = U // NP

    def copy_step():
        for i in range(NP):
            with torch.cuda.stream(s_up):
                xdev.copy_(x16_pin[i * Usb:(i + 1) * Usb], non_blocking=True)
            with torch.cuda.stream(s_dn):
                out_scratch[i * Usb:(i + 1) * Usb].copy_(ydev, non_blocking=True)

    out_scratch = torch.empty_like(out_pin).pin_memory()
    copy_step(); barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        copy_step()
    barrier()
    copy_s = time.perf_counter() - t0
    del xdev, ydev, out_scratch
    sampler.stop_flag = True; sampler.join(timeout=2)

    # ---- parity at benchmark size (rank 0): four utterances of the 256-batch, chosen by a fixed seed, against the fp64 restatement
    # of the reference (oracle/restate.py) — the device-resident arm's subband output and time signal, and the e2e arm's host buffer
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import restate
        pick = sorted(np.random.default_rng(20261017).choice(U, size=4, replace=False).tolist())
        Ydev = pipe.fetch_subband(); ydev_t = pipe.fetch_time()
        xs = x_pin.numpy()
        e_sub = e_time = e_e2e = 0.0
        for u in pick:
            Xo = np.stack([restate.analysis(xs[u, c], h, M, m, r) for c in range(C)], axis=1)
            Yo, _, _ = restate.gsc_lms(Xo, FS, delays[u], **LMS)
            yo = restate.synthesis(Yo, g, M, m, r)
            Kb = M // 2 + 1
            e_sub = max(e_sub, float(np.linalg.norm(Ydev[u] - Yo[:, :Kb]) / np.linalg.norm(Yo[:, :Kb])))
            e_time = max(e_time, float(np.linalg.norm(ydev_t[u] - yo) / np.linalg.norm(yo)))
            e_e2e = max(e_e2e, float(np.linalg.norm(out_pin[u].numpy() - yo) / np.linalg.norm(yo)))
        parity = {"utterances": pick, "rel_l2_subband": e_sub, "rel_l2_time": e_time, "rel_l2_time_e2e_buffer": e_e2e, "tolerance": 1.0e-4,
                  "oracle": "oracle/restate.py (fp64 restatement of modulated.cc:363-469,551-612 and pybeamformer.py:659-734)",
                  "pass": bool(max(e_sub, e_time, e_e2e) < 1.0e-4)}

    dev_s = tot_ms / 1000.0
    if world > 1:
        tt = torch.tensor([dev_s, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(tt[0]), float(tt[1])
        tc = torch.tensor([copy_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        copy_s = float(tc[0])
        # the single end-of-run exchange of the path: per-utterance statistics gathered to every rank
        stats_all = sbb.gather_stats()
    else:
        stats_all = stats
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * frames_step * args.steps / dev_s
    e2e = world * frames_step * args.steps / e2e_s
    peak, peak_src = hbm_peak()
    K = M // 2 + 1; D = M >> r
    alg = {"analysis_ms": C * D * (2 if args.i16_input else 4) + C * K * 8, "perbin_ms": (C + 1) * K * 8, "synthesis_ms": K * 8 + D * 4}
    dom = max(ks, key=lambda k: ks[k])
    kname = {"analysis_ms": "k_analysis", "perbin_ms": "k_perbin<8,LMS>", "synthesis_ms": "k_synthesis"}[dom]
    achieved = alg[dom] * frames_step * args.steps / (ks[dom] / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)); traffic = tj.get(kname)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1000.0 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "xrt": value * (n / FS) / T, "config": workload_config(world),
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(U * C * n * 2 + delays.nbytes), "d2h_bytes_per_step": int(out_pin.numel() * 4 + stats.nbytes),
                "ms_per_step": 1000.0 * e2e_s / args.steps, "input": "int16 PCM, pinned", "output": "resynthesised float32 signal + per-utterance statistics (the subband spectra stay on the device, as in the reference arm's want_subband=False)",
                "sub_batches": NP, "cpu_binding": numa, "front_end": "btk20.batch.ShardedBatchBeamformer",
                "copy_only_ms_per_step": 1000.0 * copy_s / args.steps, "frac_of_copy_ceiling": copy_s / e2e_s,
                "copy_only_note": "same pinned buffers, same sub-batch chunking, H2D int16 + D2H float on two streams, no kernels; max over ranks", "pipelining": "results of sub-batch i are fetched just before its handle is re-submitted (uploads of the other sub-batches stay queued); all results on the host before the clock stops"},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {k: v / args.steps for k, v in ks.items()},
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": "profiles/traffic.json (tools/ncu_traffic.py from the round's ncu --set full capture of this kernel; not measurable inside an unprofiled run)",
                     "peak_source": peak_src, "algorithmic_bytes_per_frame": alg[dom],
                     "note": {"analysis_ms": "k_analysis is bound by the shared-memory data pipe, not by HBM (ncu, profiles/r02m_ncu_k_analysis_details.txt: L1/TEX data-pipe wavefronts 72 % of peak with the byte count at the decomposition's minimum, issue slots 50 %, DRAM 47 %); the HBM fraction is reported as the contract asks",
                              "perbin_ms": "HBM-bound: ncu DRAM traffic equals the algorithmic bytes (profiles/traffic.json)",
                              "synthesis_ms": "bound by the shared-memory data pipe (ncu: L1/TEX 70 %)"}[dom],
                     "all_kernels_frac": {k: alg[k] * frames_step * args.steps / (ks[k] / 1000.0) / 1e9 / peak for k in ks if ks[k] > 0}},
        "clocks": sampler.summary(),
        "parity_check": parity,
        "stats_check": {"utterances": int(stats_all.shape[0]), "frames": float(stats_all[:, 1].sum()), "nlms_updates": float(stats_all[:, 2].sum())},
    }
    if not args.no_cpu_baseline and world == 1:
        ncores = os.cpu_count() or 1
        upc = max(1, U // ncores)   # the whole batch of the step spread over the host cores: ~15 s of CPU work on a 16-core box
        v, f, s, cores = cpu_reference(upc, ncores)
        v1, f1, s1, _ = cpu_reference(8, 1)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference",
                                "sample": "%d utterances of configs[1] (%d per core, %d processes; reference C++ via oracle/_ref, NLMS restated in C++), %.1f s wall = %.0f s of CPU work"
                                          % (upc * cores, upc, cores, s, s * cores),
                                "single_core_value": v1, "xrt": v * (n / FS) / T}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--i16-input", action="store_true", help="device-resident arm: 16-bit PCM samples in HBM instead of float32")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of four utterances of the batch")
    ap.add_argument("--utterances", type=int, default=None, help="override utterances per GPU (default 256 = configs[1])")
    args = ap.parse_args()
    if args.utterances:
        CFG["U"] = args.utterances
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
