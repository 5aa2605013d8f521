#!/usr/bin/env python
"""Parity numbers (relative L2 of the CUDA path, through the C-ABI) on the reference's own test fixtures and parameter files:
golden_online_kinect_c4_m256 (test_online_beamforming.py x confs/*.json), golden_sos_kinect[_vad]_c4_m256
(test_sos_batch_beamforming.py) and golden_wpe_kinect_c4_m256 (test_subband_dereverberator.py).  One row per configuration; a row
that raises is reported as {"error": ...} instead of stopping the report.
Run on a GPU box: python tools/parity_report_kinect.py > gpurun_out/parity_kinect.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, rel_l2, GOLDEN
from distant_speech_recognition_b200 import _capi as capi

FS, M, C = 16000.0, 256, 4
pr = np.load(os.path.join(GOLDEN, "prototype_shipped_M256_m4_r1.npz"))
g = load_golden("online_kinect_c4_m256")
x16 = np.ascontiguousarray(g["x16"]); s0, n = int(g["s0_static"]), int(g["n_static"]); xs = np.ascontiguousarray(x16[:, s0:s0 + n])
d, mpos = g["delays"], g["mpos"]
out = {}


def row(name, fn):
    try:
        out[name] = fn()
    except Exception as e:  # noqa: BLE001 (a report, not a test)
        out[name] = {"error": "%s: %s" % (type(e).__name__, e)}


def online(name, x, setup=None, **kw):
    def f():
        p = capi.Pipeline(C, M, 4, 1, max_utterances=1, max_samples=x.shape[1], **kw)
        p.set_prototypes(pr["h"], pr["g"]); p.set_delays(d[None])
        if setup:
            setup(p)
        p.submit_i16(x[None]); p.run(True)
        r = dict(Y=rel_l2(p.fetch_subband()[0], g["Y_" + name]), time=rel_l2(p.fetch_time()[0], g["time_" + name]),
                 total_energy_ratio=float(p.fetch_stats()[0, 0] / float(g["energy_" + name])))
        if "n_updates_" + name in g.files:
            r["updates"] = [int(p.fetch_stats()[0, 2]), int(g["n_updates_" + name])]
        p.close()
        return r
    row("online/" + name, f)


def sd(p):
    p.set_diffuse_noise_model(1, mpos); p.calc_mvdr_weights(0.01)


def coh(load):
    def f(p):
        sd(p); p.pf_set_diffuse_noise_model(mpos, FS); p.pf_set_diagonal_loading(load)
    return f


online("ds", xs, beamformer=capi.BF_GSC)
online("ds_and_zelinski", xs, beamformer=capi.BF_GSC, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2)
online("sd", xs, sd, beamformer=capi.BF_MVDR)
online("sd_and_zelinski", xs, sd, beamformer=capi.BF_MVDR, postfilter=capi.PF_ZELINSKI, pf_alpha=0.7, pf_type=2)
online("sd_and_mccowan", xs, coh(0.01), beamformer=capi.BF_MVDR, postfilter=capi.PF_MCCOWAN, pf_alpha=0.7, pf_type=2)
online("sd_and_lefkimmiatis", xs, coh(0.1), beamformer=capi.BF_MVDR, postfilter=capi.PF_LEFKIMMIATIS, pf_alpha=0.8, pf_type=2, pf_min_sv=1e-4, pf_fbin1=100)
online("gsclms", x16, beamformer=capi.BF_GSC_LMS)
online("gscrls", x16, beamformer=capi.BF_GSC_RLS)

gv = load_golden("sos_kinect_vad_c4_m256"); f0, f1 = [int(v) for v in gv["frames"]]
for name, kind in (("bmvdr_vad", capi.SOS_BMVDR), ("gev_vad", capi.SOS_GEV)):
    def f(name=name, kind=kind):
        p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=x16.shape[1]); p.set_prototypes(pr["h"], pr["g"])
        p.submit_i16(x16[None]); p.run_analysis(); p.sos_accumulate_from_label(gv["labels"], 10.0)
        p.sos_calc_weights(kind, gamma=1e-6, ref_micx=0, offset=0.0)
        w = p.get_weights()[0]; sgn = 1.0 if kind == capi.SOS_BMVDR else float(np.sign(np.real(np.vdot(w[0], gv["w_" + name][0]))))
        if sgn < 0:
            p.set_weights((sgn * w)[None])
        p.run_beamformer(True)
        r = dict(w=rel_l2(sgn * w, gv["w_" + name]), Y=rel_l2(p.fetch_subband()[0][f0:f1], gv["Y_" + name]), time=rel_l2(p.fetch_time()[0], gv["time_" + name]), global_sign=sgn)
        p.close()
        return r
    row("sos/" + name, f)


def smi():
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_MVDR, max_utterances=1, max_samples=x16.shape[1]); p.set_prototypes(pr["h"], pr["g"])
    p.set_delays(gv["delays"][None]); p.submit_i16(x16[None]); p.run_analysis(); p.accumulate_covariance(labels=gv["labels"][:1], energy_threshold=10.0)
    r = dict(cov=rel_l2(p.get_covariance()[0], gv["cov_smimvdr"]))
    p.calc_mvdr_weights(float(gv["mu_smimvdr"])); p.run_beamformer(True)
    r.update(w=rel_l2(p.get_weights()[0][1:], gv["w_smimvdr"][1:]), Y=rel_l2(p.fetch_subband()[0][f0:f1], gv["Y_smimvdr"]), time=rel_l2(p.fetch_time()[0], gv["time_smimvdr"]))
    p.close()
    return r


row("sos/smimvdr", smi)


def tfmask():
    gt = load_golden("sos_kinect_c4_m256"); xt = gt["x16"]
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=xt.shape[1]); p.set_prototypes(pr["h"], pr["g"])
    p.submit_i16(xt[None]); p.run_analysis(); p.sos_accumulate_from_tfmask(gt["mask_t"].astype(np.float32), gt["mask_j"].astype(np.float32), 10.0)
    p.sos_calc_weights(capi.SOS_BMVDR, gamma=1e-6, ref_micx=0, offset=0.0); w = p.get_weights()[0]; p.run_beamformer(True)
    r = dict(w=rel_l2(w, gt["w_bmvdr"]), Y=rel_l2(p.fetch_subband()[0], gt["Y_bmvdr"]), time=rel_l2(p.fetch_time()[0], gt["time_bmvdr"]))
    p.close()
    return r


row("sos/bmvdr_tfmask", tfmask)


def wpe():
    gw = load_golden("wpe_kinect_c4_m256"); conf = json.loads(str(gw["conf"])); a, b = [int(v) for v in gw["frames"]]
    p = capi.Pipeline(C, M, 4, 1, beamformer=capi.BF_DS, max_utterances=1, max_samples=x16.shape[1], wpe=conf); p.set_prototypes(pr["h"], pr["g"])
    p.submit_i16(x16[None]); p.run_analysis(); p.run_wpe(); Xd = p.fetch_snapshots()
    r = dict(X_multi=rel_l2(Xd[0][a:b], gw["X_multi"]), wpe_ms=p.last_timing_wpe())
    q = capi.Pipeline(1, M, 4, 1, beamformer=capi.BF_DS, max_utterances=4, max_samples=x16.shape[1]); q.set_prototypes(pr["h"], pr["g"])
    q.set_subband(np.ascontiguousarray(np.transpose(Xd[0], (1, 0, 2)))); q.run_synthesis(); r["time_multi"] = rel_l2(q.fetch_time(), gw["time_multi"])
    q.close(); p.close()
    return r


row("wpe/multi_channel (confs/wpe.json)", wpe)
print(json.dumps(out, indent=1))
