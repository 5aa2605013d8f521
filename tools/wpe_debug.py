"""GPU debugging aid for the WPE kernels: per-config status and errors vs the fp64 restatement (not a test)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from distant_speech_recognition_b200 import _capi as capi
from oracle import restate
from test_oracle import WPE_A, WPE_B, WPE_C, WPE_8

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run(name, C, M, cfgs):
    g = np.load(os.path.join(G, "golden_%s.npz" % name)); pr = np.load(os.path.join(G, "prototype_M%d_m4_r1.npz" % M))
    x = g["x"]; K = M // 2 + 1
    X = np.stack([restate.analysis(x[c], pr["h"], M, 4, 1) for c in range(C)], axis=1)
    for tag, kw0 in cfgs:
        for iters, f32 in ((1, 0), (kw0["iterations_num"], 0), (kw0["iterations_num"], 1)):
            kw = dict(kw0); kw["iterations_num"] = iters; kw["fp32_normal_equations"] = f32
            start, end = kw.pop("start_frame_no", 0), kw.pop("end_frame_no", -1)
            p = capi.Pipeline(C, M, 4, 1, max_utterances=1, max_samples=x.shape[1], beamformer=capi.BF_DS, wpe=kw)
            p.set_prototypes(pr["h"], pr["g"]); p.submit(x[None]); p.run_analysis()
            kwo = dict(kw); kwo.pop("fp32_normal_equations")
            Xo, Go, used = restate.wpe(X, samplerate=16000.0, start_frame_no=start, end_frame_no=end, **kwo)
            try:
                p.run_wpe(start, end)
                Xd = p.fetch_snapshots()[0]; Gd = p.get_wpe_filter()[0]
                Gor = np.transpose(Go, (1, 0, 2))
                print(name, tag, "iters", iters, "fp32" if f32 else "fp64", "X' err", rel(Xd, Xo[:, :, :K]), "G err", rel(Gd, Gor), "max|G|", float(np.abs(Gor).max()),
                      "nan", int(np.isnan(Xd).sum()), "ms", p.last_timing_wpe(), flush=True)
            except capi.BtkbError as e:
                Gd = np.empty((1, K, C, C * (kw["upper_num"] - kw["lower_num"] + 1)), np.complex64)
                print(name, tag, "iters", iters, "fp32" if f32 else "fp64", "FAILED:", str(e).split("\n")[0], flush=True)
            p.close()


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "c4"):
    run("wpe_c4_m256", 4, 256, (("a", WPE_A), ("b", WPE_B), ("c", WPE_C)))
if which in ("all", "c8"):
    run("wpe_c8_m512", 8, 512, (("8", WPE_8),))
if which == "b":
    run("wpe_c4_m256", 4, 256, (("b", WPE_B),))
