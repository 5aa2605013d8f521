"""btk20.batch — the batch front end: the same pipe as the stream classes, for many utterances per submission (the form
bench.py measures).  Not part of the reference's surface (btk2.0 processes one utterance, one frame at a time); parameter
names follow the reference's scripts and JSON configs (unit_test/confs/*.json)."""
import numpy as np

from .. import _capi

_KINDS = {"delay_and_sum": _capi.BF_DS, "ds": _capi.BF_DS, "gsc": _capi.BF_GSC, "gsclms": _capi.BF_GSC_LMS,
          "mvdr": _capi.BF_MVDR, "sd": _capi.BF_MVDR, "smimvdr": _capi.BF_MVDR}


class BatchBeamformer:
    """ap_conf-style construction: BatchBeamformer(chan_num, h_fb, g_fb, M, m, r, beamformer={"type": "gsclms", ...},
    postfilter={"type": "zelinski", "subtype": 2, "alpha": 0.7})."""

    def __init__(self, chan_num, h_fb, g_fb, M=512, m=4, r=1, samplerate=16000, beamformer=None, postfilter=None,
                 max_utterances=256, max_samples=80000, device=0):
        bf = dict(beamformer or {"type": "delay_and_sum"})
        self.type = bf.pop("type")
        if self.type not in _KINDS:
            raise ValueError("unsupported beamformer type %r" % self.type)
        pf = dict(postfilter or {})
        if pf and pf.get("type") != "zelinski":
            raise ValueError("unsupported post-filter %r" % pf.get("type"))
        lms = {k: v for k, v in bf.items() if k in ("beta", "gamma", "init_diagonal_load", "regularization_param", "energy_floor",
                                                      "sil_thresh", "max_wa_l2norm", "min_frames", "slowdown_after")}
        self.mu = bf.get("mu", 1e-4 if self.type == "smimvdr" else 0.01)
        self.energy_threshold = bf.get("energy_threshold", 10)
        self.samplerate = samplerate
        self.pipe = _capi.Pipeline(chan_num, M, m, r, 2, samplerate, _KINDS[self.type], _capi.PF_ZELINSKI if pf else _capi.PF_NONE,
                                   pf.get("alpha", 0.6), pf.get("subtype", 2), pf.get("min_frames", 0), lms or None,
                                   max_utterances, max_samples, device)
        self.pipe.set_prototypes(h_fb, g_fb)

    def process(self, samples, delays, lengths=None, vad_labels=None, mpos=None, sspeed=343740.0, synthesis=True):
        """samples float32 [U][C][n], delays [U][C] -> (time [U][n_out] float32, subband [U][T][K] complex64, stats [U][3])."""
        p = self.pipe
        U = samples.shape[0]
        p.set_delays(delays)
        p.submit(samples, lengths)
        if self.type == "smimvdr":
            p.run_analysis()
            p.accumulate_covariance(vad_labels, self.energy_threshold)
            p.calc_mvdr_weights(self.mu)
            p.run_beamformer(synthesis)
        elif self.type in ("mvdr", "sd"):
            p.set_diffuse_noise_model(U, np.asarray(mpos, np.float64), sspeed)
            p.calc_mvdr_weights(self.mu)
            p.run(synthesis)
        else:
            p.run(synthesis)
        return (p.fetch_time() if synthesis else None), p.fetch_subband(), p.fetch_stats()
