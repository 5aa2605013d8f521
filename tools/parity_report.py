#!/usr/bin/env python
"""Print the parity numbers (relative L2 of the CUDA path vs the reference goldens and vs the fp64 oracle) as JSON.
Run on a GPU box: python tools/parity_report.py > gpurun_out/parity.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2
from distant_speech_recognition_b200 import _capi as capi
from oracle import restate

FS = 16000.0
def protos(M):
    p = np.load(os.path.join(ROOT, "tests", "golden", "prototype_M%d_m4_r1.npz" % M)); return p["h"], p["g"]
def pipe(C, M, n, **kw):
    h, g = protos(M); p = capi.Pipeline(C, M, 4, 1, max_utterances=1, max_samples=n, **kw); p.set_prototypes(h, g); return p
out = {}
g = load_golden("ds_c2_m256"); p = pipe(2, 256, g["x"].shape[1], beamformer=capi.BF_DS); p.set_delays(g["delays"][None]); p.submit(g["x"][None]); p.run(True)
out["ds_c2_m256"] = dict(X=rel_l2(p.fetch_snapshots()[0][:, 0], g["X0"]), Y=rel_l2(p.fetch_subband()[0], g["Y"]), time=rel_l2(p.fetch_time()[0], g["time"]))
g = load_golden("gsclms_c8_m512"); p = pipe(8, 512, g["x"].shape[1], beamformer=capi.BF_GSC_LMS, lms=dict(min_frames=int(g["min_frames"]))); p.set_delays(g["delays"][None]); p.submit(g["x"][None]); p.run(True)
out["gsclms_c8_m512"] = dict(Y=rel_l2(p.fetch_subband()[0], g["Y"]), time=rel_l2(p.fetch_time()[0], g["time"]), waH=rel_l2(p.get_active_weights()[0], g["waH"]))
g = load_golden("gsc_zelinski_c8_m512"); p = pipe(8, 512, g["x"].shape[1], beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2); p.set_delays(g["delays"][None]); p.set_active_weights(g["wa"][None]); p.submit(g["x"][None]); p.run(True)
out["gsc_zelinski_c8_m512"] = dict(Y=rel_l2(p.fetch_subband()[0], g["Y"]), time=rel_l2(p.fetch_time()[0], g["time"]))
g = load_golden("smimvdr_zelinski_c8_m512"); x = g["x"]; M = 512; h, gg = protos(M)
p = pipe(8, 512, x.shape[1], beamformer=capi.BF_MVDR, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2); p.set_delays(g["delays"][None]); p.submit(x[None]); p.run_analysis()
p.accumulate_covariance(labels=g["label"][None], energy_threshold=10.0); p.calc_mvdr_weights(float(g["mu"])); p.run_beamformer(True)
X = np.stack([restate.analysis(x[c], h, M, 4, 1) for c in range(8)], axis=1)
R, nf = restate.smi_covariance(X, FS, 256, ((0.25, 0.75),), 10.0); wq = restate.calc_mainlobe(M, 8, FS, g["delays"])
w64 = restate.calc_mvdr_weights(R + float(np.float32(g["mu"])) * np.eye(8), wq, single=False)
Y64, _ = restate.zelinski_postfilter(restate.subband_mvdr(X, w64), X, wq, 0.7, 2, 0)
Y = p.fetch_subband()[0]; y = p.fetch_time()[0]
ev = np.linalg.eigvalsh(R[100] + 1e-4 * np.eye(8))
out["smimvdr_zelinski_c8_m512"] = dict(cov_vs_ref=rel_l2(p.get_covariance()[0], g["cov"]), w_vs_ref=rel_l2(p.get_weights()[0], g["w"]), w_vs_fp64=rel_l2(p.get_weights()[0], w64[:257]),
    ref_w_vs_fp64=rel_l2(g["w"], w64[:257]), Y_vs_ref=rel_l2(Y, g["Y"]), Y_vs_fp64=rel_l2(Y, Y64[:, :257]), refY_vs_fp64=rel_l2(g["Y"], Y64[:, :257]),
    time_vs_ref=rel_l2(y, g["time"]), time_vs_fp64=rel_l2(y, restate.synthesis(Y64, gg, M, 4, 1)), cond_R_bin100=float(ev[-1] / ev[0]))
g = load_golden("mvdrsd_zelinski1_c4_m256"); p = pipe(4, 256, g["x"].shape[1], beamformer=capi.BF_MVDR, postfilter=capi.PF_ZELINSKI, pf_alpha=0.6, pf_type=1, pf_min_frames=5)
p.set_delays(g["delays"][None]); p.set_diffuse_noise_model(1, g["mpos"]); p.calc_mvdr_weights(float(g["mu"])); p.submit(g["x"][None]); p.run(True)
out["mvdrsd_zelinski1_c4_m256"] = dict(w=rel_l2(p.get_weights()[0], g["w"]), Y=rel_l2(p.fetch_subband()[0], g["Y"]), time=rel_l2(p.fetch_time()[0], g["time"]))
# ---- rows added after the first report: RLS, WPE, SOS (blind MVDR / GEV), 64-mic tensor-core covariance
g = load_golden("gscrls_c8_m512"); p = pipe(8, 512, g["x"].shape[1], beamformer=capi.BF_GSC_RLS, rls=dict(min_frames=int(g["min_frames"]))); p.set_delays(g["delays"][None]); p.submit(g["x"][None]); p.run(True)
out["gscrls_c8_m512 (reference Python's own output)"] = dict(Y=rel_l2(p.fetch_subband()[0], g["Y"]), time=rel_l2(p.fetch_time()[0][: len(g["time"])], g["time"]))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_oracle import WPE_8
g = load_golden("wpe_c8_m512"); p = pipe(8, 512, g["x"].shape[1], beamformer=capi.BF_DS, wpe=dict(WPE_8)); p.submit(g["x"][None]); p.run_analysis(); p.run_wpe()
out["wpe_c8_m512 (compiled reference)"] = dict(X_dereverberated=rel_l2(p.fetch_snapshots()[0], g["Xa"]))
for name, C, M, kind in (("bmvdr_vad_c8_m512", 8, 512, capi.SOS_BMVDR), ("bmvdr_tfmask_c4_m256", 4, 256, capi.SOS_BMVDR), ("gev_vad_c8_m512", 8, 512, capi.SOS_GEV), ("gev_tfmask_c4_m256", 4, 256, capi.SOS_GEV)):
    g = load_golden(name); p = pipe(C, M, g["x"].shape[1], beamformer=capi.BF_DS); p.submit(g["x"][None]); p.run_analysis()
    if "mask_t" in g.files: p.sos_accumulate_from_tfmask(g["mask_t"], g["mask_j"], float(g["energy_threshold"]))
    else: p.sos_accumulate_from_label(g["labels"], float(g["energy_threshold"]))
    p.sos_calc_weights(kind, gamma=float(g["gamma"]), ref_micx=int(g["ref_micx"]), offset=float(g["offset"]))
    w = p.get_weights()[0]; sgn = float(np.sign(np.real(np.vdot(w[0], g["w"][0])))) if kind == capi.SOS_GEV else 1.0
    if sgn < 0: p.set_weights((sgn * w)[None])
    p.run_beamformer(True)
    out[name + " (reference Python's own output)"] = dict(w=rel_l2(sgn * w, g["w"]), Y=rel_l2(p.fetch_subband()[0], g["Y"]), time=rel_l2(p.fetch_time()[0], g["time"]), global_sign=sgn)
from distant_speech_recognition_b200 import synthetic
xx, dd = synthetic.make_batch(1, 64, 20000, first=77); h, gg = protos(512)
p = pipe(64, 512, 20000, beamformer=capi.BF_MVDR); p.set_delays(dd); p.submit(xx); p.run_analysis(); p.accumulate_covariance(labels=np.array([[0.4, 0.8]]), energy_threshold=10.0)
X = np.stack([restate.analysis(xx[0, c], h, 512, 4, 1) for c in range(64)], axis=1); R, nf = restate.smi_covariance(X, FS, 256, ((0.4, 0.8),), 10.0)
out["64-mic covariance, tcgen05 3xTF32 (vs fp64 restatement)"] = dict(R=rel_l2(p.get_covariance()[0], R), noise_frames=int(nf))
print(json.dumps(out, indent=1))
