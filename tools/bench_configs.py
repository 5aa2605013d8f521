#!/usr/bin/env python
"""Device-side timing of the BASELINE.json configs other than the bench.py headline (configs[1]) — parity-test shapes run at
full size for the record (profiles/), not bench lines.  python tools/bench_configs.py > gpurun_out/configs.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distant_speech_recognition_b200 import _capi, synthetic

FS = 16000.0


def proto(M):
    p = np.load(os.path.join(ROOT, "tests", "golden", "prototype_M%d_m4_r1.npz" % M)); return p["h"], p["g"]


def tiled_batch(U, C, n, distinct):
    x, d = synthetic.make_batch(distinct, C, n, pcm16=True)
    reps = -(-U // distinct)
    return np.tile(x, (reps, 1, 1))[:U], np.tile(d, (reps, 1))[:U]


def timed(fn, steps=5, warm=2):
    for _ in range(warm): fn()
    t0 = time.perf_counter()
    for _ in range(steps): fn()
    return (time.perf_counter() - t0) / steps


def _rel(a, b):
    return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300))


def parity_and_cpu(p, x, d, M, kind, picks, labels=None):
    """Parity at benchmark size: the utterances `picks` of the batch the pipeline has just processed against the fp64 restatement
    (oracle/restate.py), and the reference's own C++ (oracle/_ref) timed on ONE host core for the first of them (cpu_baseline)."""
    import time as _t
    from oracle import restate, ref
    h, g = proto(M)
    C = x.shape[1]; K = M // 2 + 1; FS = 16000.0
    Y = p.fetch_subband(); y = p.fetch_time()
    es, et = 0.0, 0.0
    for u in picks:
        X = np.stack([restate.analysis(x[u, c], h, M, 4, 1) for c in range(C)], axis=1)
        wq = restate.calc_mainlobe(M, C, FS, d[u])
        if kind == "ds":
            Yo = restate.subband_ds(X, wq)
        elif kind == "smimvdr_zelinski":
            R, _ = restate.smi_covariance(X, FS, M // 2, (tuple(labels[u]),), 10.0)
            w = restate.calc_mvdr_weights(R + float(np.float32(1e-4)) * np.eye(C), wq, single=False)
            Yo, _ = restate.zelinski_postfilter(restate.subband_mvdr(X, w), X, wq, 0.7, 2, 0)
        else:
            Yo, _, _ = restate.gsc_lms(X, FS, d[u])
        yo = restate.synthesis(Yo, g, M, 4, 1)
        T = X.shape[0]
        es = max(es, _rel(Y[u, :T], Yo[:, :K])); et = max(et, _rel(y[u, :len(yo)], yo))
    u = picks[0]
    bf = {"ds": ref.BF_DS, "smimvdr_zelinski": ref.BF_SMI_MVDR, "gsclms": ref.BF_GSC_LMS}[kind]
    kw = dict(pf=dict(kind="zelinski", alpha=0.7, type=2), smi_label=tuple(labels[u]), mvdr_mu=1e-4) if kind == "smimvdr_zelinski" else {}
    t0 = _t.perf_counter(); r = ref.beamform(x[u], h, g, d[u], M, 4, 1, bf_kind=bf, do_synthesis=True, want_subband=False, **kw); cpu = _t.perf_counter() - t0
    return {"parity_check": {"utterances": list(picks), "rel_l2_subband": es, "rel_l2_time": et, "tolerance": 1e-4, "oracle": "oracle/restate.py (fp64)", "pass": bool(max(es, et) < 1e-4)},
            "cpu_baseline": {"kind": "reference", "cores": 1, "value": float(r["stats"][1]) / cpu, "unit": "frames/s",
                             "sample": "utterance %d of the batch through the compiled reference (oracle/_ref), %.1f s on one host core" % (u, cpu)}}


def main():
    out = {}
    # configs[0]: 2-mic D&S, M=256, one 10 s utterance
    C, M, U, n = 2, 256, 1, 160000
    h, g = proto(M); x, d = tiled_batch(U, C, n, 1)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_DS, max_utterances=U, max_samples=n); p.set_prototypes(h, g); p.set_delays(d); p.submit(x)
    def f0(): p.run(True); p.synchronize()
    s = timed(f0); T = p.num_frames
    out["configs[0] 2-mic SubbandDS M=256 1x10s"] = dict(frames=U * T, ms=1e3 * s, frames_per_s=U * T / s, kernels=p.last_timing(), **parity_and_cpu(p, x, d, M, "ds", [0]))
    p.close()
    # configs[2]: 8-mic SMI-MVDR + Zelinski, M=512, 1000 utterances
    C, M, U, n = 8, 512, 1000, 80000
    h, g = proto(M); x, d = tiled_batch(U, C, n, 16)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_MVDR, postfilter=_capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.synchronize()
    labels = np.tile(np.array([[1.0, 5.0]]), (U, 1))
    def f2():
        p.run_analysis(); p.accumulate_covariance(labels, 10.0); p.calc_mvdr_weights(1e-4); p.run_beamformer(True); p.synchronize()
    s = timed(f2, steps=3, warm=1); T = p.num_frames
    out["configs[2] 8-mic SMI-MVDR+Zelinski M=512 1000x5s"] = dict(frames=U * T, ms=1e3 * s, frames_per_s=U * T / s, kernels_last_call=p.last_timing(),
                                                                   **parity_and_cpu(p, x, d, M, "smimvdr_zelinski", [3, 522, 999], labels))
    p.close(); del x
    # configs[3]: 64-mic GSC-NLMS, M=512, 256 utterances
    C, M, U, n = 64, 512, 256, 80000
    h, g = proto(M); x, d = tiled_batch(U, C, n, 4)
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.synchronize()
    def f3(): p.run(True); p.synchronize()
    s = timed(f3, steps=3, warm=1); T = p.num_frames; t = p.last_timing()
    bytes_perbin = (C + 1) * (M // 2 + 1) * 8 * U * T
    out["configs[3] 64-mic GSC-NLMS M=512 256x5s"] = dict(frames=U * T, ms=1e3 * s, frames_per_s=U * T / s, kernels=t,
                                                          perbin_hbm_frac=bytes_perbin / (t["perbin_ms"] * 1e-3) / 1e9 / 6530.3, **parity_and_cpu(p, x, d, M, "gsclms", [1, 254]))
    p.close()
    # configs[3] with the covariance pass: 64-mic SMI-MVDR (k_covariance_wide + k_mvdr_solve_wide + static apply), same batch
    p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_MVDR, max_utterances=U, max_samples=n)
    p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.synchronize()
    labels = np.tile(np.array([[1.0, 5.0]]), (U, 1))
    p.run_analysis(); p.synchronize()
    def f3c(): p.accumulate_covariance(labels, 10.0); p.synchronize()
    s_cov = timed(f3c, steps=2, warm=1)
    def f3s(): p.calc_mvdr_weights(1e-4); p.synchronize()
    s_sol = timed(f3s, steps=2, warm=1)
    def f3a(): p.run_beamformer(True); p.synchronize()
    s_app = timed(f3a, steps=2, warm=1); T = p.num_frames
    cov_flops = 8.0 * C * C * (M // 2 + 1) * U * T   # upper bound: every frame accumulated
    out["configs[3]+covariance 64-mic SMI-MVDR M=512 256x5s"] = dict(frames=U * T, covariance_ms=1e3 * s_cov, solve_ms=1e3 * s_sol, apply_synthesis_ms=1e3 * s_app,
                                                                       covariance_hbm_frac=C * (M // 2 + 1) * 8 * U * T / s_cov / 1e9 / 6566.7,
                                                                       covariance_tflops_if_all_frames=cov_flops / s_cov / 1e12)
    p.close(); del x
    # configs[4]: 8-mic GSC-NLMS behind multi-channel WPE (confs/wpe.json: 33 lags, 2 iterations), M=1024; a bounded sample of the
    # 1 024-utterance-per-GPU shard (WPE is compute-bound: fp64 normal equations, L = 264 unknowns per channel and bin)
    C, M, U, n = 8, 1024, 8, 80000
    h, g = proto(M); x, d = tiled_batch(U, C, n, 4)
    for tag, fp32 in (("fp64 normal equations (reference arithmetic)", 0), ("fp32 normal equations", 1)):
        wpe = dict(lower_num=0, upper_num=32, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4, fp32_normal_equations=fp32)
        p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n, wpe=wpe)
        p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.synchronize()
        def f4(): p.run(True); p.synchronize()
        s = timed(f4, steps=1, warm=1); T = p.num_frames; t = p.last_timing()
        out["configs[4] 8-mic WPE+GSC-NLMS M=1024 sample of %d x 5s, %s" % (U, tag)] = dict(
            frames=U * T, ms=1e3 * s, frames_per_s=U * T / s, wpe_ms=p.last_timing_wpe(), kernels=t,
            extrapolated_s_per_1024_utterance_shard=s * 1024 / U)
        p.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
