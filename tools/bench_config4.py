#!/usr/bin/env python
"""BASELINE.json configs[4] at its stated size: 8-mic SubbandGSC (NLMS) behind multi-channel WPE dereverberation (confs/wpe.json: 33 lags,
2 iterations, load -18 dB, fp64 normal equations like the reference), 1 024 subbands, an 8 k-utterance stream of 5 s utterances sharded
over the GPUs of one box (1 024 utterances per GPU), analysis -> WPE -> GSC-NLMS -> synthesis.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_config4.py
  python tools/bench_config4.py --utterances-per-gpu 64           # one GPU, a sample

Every rank processes its shard in sub-batches of --sub-batch utterances (device-resident 16-bit PCM in, time signal left on the device;
this measures the device pipeline, the PCIe path is bench.py's e2e arm).  Time = barrier-bracketed wall clock, max over ranks.  One JSON line
from rank 0, with a parity check of the first sub-batch against oracle/restate.py when --parity is given, and — with --cpu-sample — the
reference's own C++ WPE (oracle/_ref) timed on rank 0's host cores over a bounded sample: the first `nbins` bins (and their mirrors) of one
utterance, scaled to the M bins the reference estimates (the estimation is independent per bin, dereverberation.cc:665-690)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))

WPE = dict(lower_num=0, upper_num=32, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances-per-gpu", type=int, default=1024)
    ap.add_argument("--sub-batch", type=int, default=64)
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic utterances, tiled to the sub-batch")
    ap.add_argument("--fp32-normal-equations", action="store_true")
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="bins of one utterance to time the reference's C++ WPE on (0 = skip)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from distant_speech_recognition_b200 import _capi
    from bench_configs import proto, tiled_batch
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        sys.stdout.flush(); saved_fd = os.dup(1); os.dup2(2, 1)   # NCCL's version banner goes to stderr, stdout carries the JSON line
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved_fd, 1); os.close(saved_fd)
    C, M, n = 8, 1024, 80000
    Ub, Ug = args.sub_batch, args.utterances_per_gpu
    assert Ug % Ub == 0
    h, g = proto(M)
    x, d = tiled_batch(Ub, C, n, args.distinct)
    x16 = np.ascontiguousarray(x.astype(np.int16))
    wpe = dict(WPE, fp32_normal_equations=1 if args.fp32_normal_equations else 0)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=Ub, max_samples=n, wpe=wpe, device=local)
    p.set_prototypes(h, g); p.set_delays(d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def sub_batch():
        p.submit_i16(x16); p.run(True); p.synchronize()

    sub_batch()                                        # warm-up (also pages the library in)
    T = p.num_frames
    wpe_ms = p.last_timing_wpe(); tim = p.last_timing(); form = ("lag-domain (L x L)", "frame-domain (S x S)")[p.last_wpe_form()]
    parity = None
    if args.parity and rank == 0:
        from oracle import restate
        Y = p.fetch_subband()[0]; y = p.fetch_time()[0]
        X = np.stack([restate.analysis(x[0, c], h, M, 4, 1) for c in range(C)], axis=1)
        Xd = restate.wpe(X, **WPE)[0]
        Yo, _, _ = restate.gsc_lms(Xd, 16000.0, d[0])
        yo = restate.synthesis(Yo, g, M, 4, 1)
        parity = {"rel_l2_subband": float(np.linalg.norm(Y - Yo[:, :M // 2 + 1]) / np.linalg.norm(Yo[:, :M // 2 + 1])),
                  "rel_l2_time": float(np.linalg.norm(y - yo) / np.linalg.norm(yo)), "oracle": "oracle/restate.py (fp64), utterance 0 of the sub-batch"}
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ug // Ub):
        sub_batch()
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
    if rank == 0:
        frames = world * Ug * T
        line = {"workload": "configs[4]: 8-mic SubbandGSC (NLMS) + multi-channel WPE (33 lags, 2 iterations, %s normal equations), 1024 subbands, %d utterances of 5 s on %d GPU(s) (%d per GPU, sub-batches of %d)"
                            % ("fp32" if args.fp32_normal_equations else "fp64", world * Ug, world, Ug, Ub),
                "n_gpus": world, "utterances": world * Ug, "frames": frames, "seconds": dt, "frames_per_s": frames / dt, "xrt": world * Ug * (n / 16000.0) / dt,
                "ms_per_utterance_per_gpu": 1e3 * dt / Ug, "wpe_ms_per_sub_batch": wpe_ms, "wpe_normal_equations": form, "kernel_ms_per_sub_batch": tim, "parity_check": parity,
                "input": "16-bit PCM resident on the host, uploaded per sub-batch (pageable); time signal left on the device"}
        if args.cpu_sample > 0:
            from oracle import ref, restate
            nb = args.cpu_sample
            X = np.stack([restate.analysis(x[0, c], h, M, 4, 1) for c in range(C)], axis=1)
            bw = 16000.0 / 2.0 * nb / (M // 2)           # set_band_width_: bins <= bw / (fs/2) * M/2 and their mirrors are processed
            t1 = time.perf_counter(); ref.wpe(X, band_width=bw, **{k: v for k, v in WPE.items() if k != "band_width"}); cpu = time.perf_counter() - t1
            per_utt = cpu * M / (2 * nb + 1)               # the reference estimates all M bins, mirror half included (estimate_Gn_ loops subbandX < size(), :665-672)
            line["cpu_reference_wpe"] = {"kind": "reference", "cores": 1, "sample": "the reference's C++ MultiChannelWPEDereverberation (oracle/_ref) restricted by band_width to bins 0..%d and their mirrors (%d of the %d bins it estimates) of one utterance: %.1f s; scaled to all %d bins: %.0f s per utterance per core (WPE only, without filter banks and beamformer)" % (nb, 2 * nb + 1, M, cpu, M, per_utt),
                                         "s_per_utterance_per_core": per_utt, "frames_per_s_per_core": T / per_utt}
        print(json.dumps(line))
    p.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
