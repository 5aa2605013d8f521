#!/usr/bin/env python
"""configs[4] sample (8 mics, M = 1024, confs/wpe.json: 33 lags, 2 iterations): device time of the WPE pass for U utterances of 5 s, in
both forms of the normal equations (BTKB_WPE_FORM, btkb_wpe.cu) and both precisions; the forms are compared on the filters and on the
dereverberated snapshots."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distant_speech_recognition_b200 import _capi
from bench_configs import proto, tiled_batch, timed
C, M, U, n = 8, 1024, int(os.environ.get("WPE_U", "8")), int(os.environ.get("WPE_N", "80000"))
h, g = proto(M); x, d = tiled_batch(U, C, n, 4)
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
ref = {}
for form in os.environ.get("WPE_FORMS", "lag,frame").split(","):
    os.environ["BTKB_WPE_FORM"] = form
    for tag, fp32 in [t for t in (("fp64", 0), ("fp32", 1)) if t[0] in os.environ.get("WPE_PREC", "fp64,fp32")]:
        wpe = dict(lower_num=0, upper_num=32, iterations_num=2, load_db=-18.0, band_width=0.0, diagonal_bias=1e-4, fp32_normal_equations=fp32)
        p = _capi.Pipeline(C, M, 4, 1, beamformer=_capi.BF_GSC_LMS, max_utterances=U, max_samples=n, wpe=wpe)
        p.set_prototypes(h, g); p.set_delays(d); p.submit(x); p.synchronize()
        s = timed(lambda: (p.run(True), p.synchronize()), steps=1, warm=1)
        G = p.get_wpe_filter()[:2]; X = p.fetch_snapshots()[:2]
        line = dict(ms=1e3 * s, wpe_ms=p.last_timing_wpe(), form=("lag", "frame")[p.last_wpe_form()], knobs={k: v for k, v in os.environ.items() if k.startswith("BTKB_WPE_")},
                    ms_per_utterance=p.last_timing_wpe() / U, s_per_1024_utt=s * 1024 / U)
        if tag == "fp64" and not ref:
            ref = dict(G=G, X=X, form=form)
        elif ref:
            line["vs_%s_fp64" % ref["form"]] = dict(filters_rel_l2=rel(G, ref["G"]), snapshots_rel_l2=rel(X, ref["X"]))
        print(json.dumps({"wpe chain %s, %s-domain, %d utterances" % (tag, form, U): line}), flush=True)
        p.close()
