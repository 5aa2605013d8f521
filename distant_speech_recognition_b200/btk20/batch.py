"""btk20.batch — the batch front end: the same pipe as the stream classes, for many utterances per submission (the form
bench.py measures).  Not part of the reference's surface (btk2.0 processes one utterance, one frame at a time); parameter
names follow the reference's scripts and JSON configs (unit_test/confs/*.json)."""
import numpy as np

from .. import _capi

_KINDS = {"delay_and_sum": _capi.BF_DS, "ds": _capi.BF_DS, "gsc": _capi.BF_GSC, "lcmv": _capi.BF_GSC, "gsclms": _capi.BF_GSC_LMS,
          "gscrls": _capi.BF_GSC_RLS, "mvdr": _capi.BF_MVDR, "sd": _capi.BF_MVDR, "super_directive": _capi.BF_MVDR, "smimvdr": _capi.BF_MVDR,
          "bmvdr": _capi.BF_DS, "gev": _capi.BF_DS}
_PFS = {"zelinski": _capi.PF_ZELINSKI, "mccowan": _capi.PF_MCCOWAN, "lefkimmiatis": _capi.PF_LEFKIMMIATIS}
_LMS_KEYS = ("beta", "gamma", "init_diagonal_load", "regularization_param", "energy_floor", "sil_thresh", "max_wa_l2norm", "min_frames", "slowdown_after")
_RLS_KEYS = ("beta", "gamma", "mu", "init_diagonal_load", "regularization_param", "sil_thresh", "constraint_option", "alpha2", "max_wa_l2norm", "min_frames")
_WPE_KEYS = ("lower_num", "upper_num", "iterations_num", "load_db", "band_width", "diagonal_bias", "fp32_normal_equations")


class BatchBeamformer:
    """ap_conf-style construction:

        BatchBeamformer(chan_num, h_fb, g_fb, M, m, r, beamformer={"type": "gsclms", ...},
                        postfilter={"type": "zelinski", "subtype": 2, "alpha": 0.7}, wpe={"lower_num": 0, "upper_num": 32, ...})

    beamformer types: delay_and_sum, gsc, lcmv, gsclms, gscrls, super_directive (sd / mvdr), smimvdr, bmvdr, gev
    (unit_test/confs/{ds,sd,lcmv_and_zelinski,gsclms,gscrls,smimvdr,bmvdr_vad,bmvdr_tfmask,gev_vad,gev_tfmask}.json);
    post-filters: zelinski, mccowan, lefkimmiatis (confs/*_and_{zelinski,mccowan,lefkimmiatis}.json); wpe: confs/wpe.json."""

    def __init__(self, chan_num, h_fb, g_fb, M=512, m=4, r=1, samplerate=16000, beamformer=None, postfilter=None, wpe=None,
                 max_utterances=256, max_samples=80000, device=0):
        bf = dict(beamformer or {"type": "delay_and_sum"})
        self.type = bf.pop("type")
        if self.type not in _KINDS:
            raise ValueError("unsupported beamformer type %r" % self.type)
        pf = dict(postfilter or {})
        if pf and pf.get("type") not in _PFS:
            raise ValueError("unsupported post-filter %r" % pf.get("type"))
        self.bf_conf, self.pf_conf = bf, pf
        lms = {k: v for k, v in bf.items() if k in _LMS_KEYS} if self.type == "gsclms" else None
        rls = {k: v for k, v in bf.items() if k in _RLS_KEYS} if self.type == "gscrls" else None
        wpe_kw = {k: v for k, v in dict(wpe).items() if k in _WPE_KEYS} if wpe is not None else None
        self.mu = bf.get("mu", bf.get("diagonal_load", 1e-4 if self.type == "smimvdr" else 0.01))
        self.energy_threshold = bf.get("energy_threshold", 10)
        self.samplerate = samplerate
        pf_default_alpha = 0.8 if pf.get("type") == "lefkimmiatis" else 0.6
        self.pipe = _capi.Pipeline(chan_num, M, m, r, 2, samplerate, _KINDS[self.type], _PFS[pf["type"]] if pf else _capi.PF_NONE,
                                   pf.get("alpha", pf_default_alpha), pf.get("subtype", 2), pf.get("min_frames", 0), lms or None,
                                   max_utterances, max_samples, device, pf_min_sv=pf.get("min_sv", 1e-8), pf_fbin1=pf.get("fbin_no1", 128 if pf.get("type") == "lefkimmiatis" else 0),
                                   rls=rls or None, wpe=wpe_kw)
        self.pipe.set_prototypes(h_fb, g_fb)

    def process(self, samples, delays=None, lengths=None, vad_labels=None, tfmasks=None, mpos=None, delays_jammers=None, sspeed=343740.0,
                synthesis=True):
        """samples float32 [U][C][n], delays [U][C] -> (time [U][n_out] float32, subband [U][T][K] complex64, stats [U][3]).
        vad_labels: [U][2] (smimvdr) or [U][NL][2] (bmvdr / gev); tfmasks: (mask_t, mask_j) [U][T][K] (bmvdr / gev);
        delays_jammers [U][Nc-1][C] (lcmv); mpos [C][3] (super_directive, mccowan, lefkimmiatis)."""
        p = self.pipe
        U = samples.shape[0]
        sos = self.type in ("bmvdr", "gev")
        if not sos:
            if self.type == "lcmv":
                p.set_delays_lcmv(delays, delays_jammers)
            else:
                p.set_delays(delays)
        if self.pf_conf.get("type") in ("mccowan", "lefkimmiatis"):
            p.pf_set_diffuse_noise_model(np.asarray(mpos, np.float64), self.samplerate, sspeed)
            p.pf_set_diagonal_loading(self.bf_conf.get("diagonal_load", 0.01 if self.pf_conf["type"] == "mccowan" else 0.1))
        p.submit(samples, lengths)
        wpe_on = bool(p.cfg.wpe.enabled)

        def front():   # analysis (+ dereverberation)
            p.run_analysis()
            if wpe_on:
                p.run_wpe()

        if self.type == "smimvdr":
            front()
            p.accumulate_covariance(vad_labels, self.energy_threshold)
            p.calc_mvdr_weights(self.mu)
            p.run_beamformer(synthesis)
        elif sos:
            front()
            p.sos_reset_stats()
            if tfmasks is not None:
                p.sos_accumulate_from_tfmask(tfmasks[0], tfmasks[1], self.energy_threshold)
            else:
                p.sos_accumulate_from_label(vad_labels, self.energy_threshold)
            p.sos_calc_weights(_capi.SOS_BMVDR if self.type == "bmvdr" else _capi.SOS_GEV, self.bf_conf.get("gamma", 1e-6),
                               self.bf_conf.get("ref_micx", 0), self.bf_conf.get("offset", 0.0))
            p.run_beamformer(synthesis)
        elif self.type in ("mvdr", "sd", "super_directive"):
            p.set_diffuse_noise_model(U, np.asarray(mpos, np.float64), sspeed)
            p.calc_mvdr_weights(self.mu)
            p.run(synthesis)
        else:
            p.run(synthesis)
        return (p.fetch_time() if synthesis else None), p.fetch_subband(), p.fetch_stats()


class ShardedBatchBeamformer:
    """The multi-GPU front end: one process per GPU (torchrun), rank g owns the contiguous utterance range
    `sharding.shard_range(total, world, rank)`; utterances are independent units (reset_stats per utterance,
    lib/pybeamformer.py:745-762), so there is NO data-path collective — the only exchange is one all-gather of the per-utterance
    statistics [sum y^2, frames, updates] at the end of a run (`gather_stats`).

    Within a rank the shard is cut into `sub_batches` pipeline handles, each with its own CUDA stream, and the host loop is
    software-pipelined over sub-batches AND steps: a sub-batch's results are collected right before its handle is re-submitted, so
    the H2D engine always has the other sub-batches' uploads queued while this one's kernels run and its D2H drains.  Input is 16-bit
    PCM in pinned host memory (what SampleFeature reads from wav files, feature/feature.cc:256-305), output the resynthesised float
    signal into a caller-supplied pinned buffer."""

    def __init__(self, chan_num, h_fb, g_fb, M=512, m=4, r=1, samplerate=16000, beamformer=None, utterances=256, max_samples=80000,
                 sub_batches=8, device=0, world=None, rank=None):
        from .. import sharding
        self.sharding = sharding
        bf = dict(beamformer or {"type": "gsclms"})
        kind = bf.pop("type")
        if kind not in ("gsclms", "gscrls", "delay_and_sum", "ds", "gsc"):
            raise ValueError("ShardedBatchBeamformer streams one-pass beamformers (delay_and_sum, gsc, gsclms, gscrls); got %r" % kind)
        self.world, self.rank = self._dist(world, rank)
        self.U = int(utterances)
        NP = max(1, min(int(sub_batches), self.U))
        while self.U % NP:
            NP -= 1
        self.NP, self.Us, self.n, self.C = NP, self.U // NP, int(max_samples), chan_num
        lms = {k: v for k, v in bf.items() if k in _LMS_KEYS} if kind == "gsclms" else None
        rls = {k: v for k, v in bf.items() if k in _RLS_KEYS} if kind == "gscrls" else None
        self.pipes = []
        for _ in range(NP):
            q = _capi.Pipeline(chan_num, M, m, r, 2, samplerate, _KINDS[kind], max_utterances=self.Us, max_samples=self.n, device=device,
                               lms=lms or None, rls=rls or None)
            q.set_prototypes(h_fb, g_fb)
            self.pipes.append(q)
        self.D = M >> r
        self.pending = [None] * NP
        self.stats = np.zeros((self.U, 3))

    @staticmethod
    def _dist(world, rank):
        if world is not None:
            return int(world), int(rank or 0)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                return dist.get_world_size(), dist.get_rank()
        except Exception:  # noqa: BLE001
            pass
        return 1, 0

    def shard(self, total):
        """[start, stop) of this rank in a global batch of `total` utterances."""
        return self.sharding.shard_range(total, self.world, self.rank)

    def _collect(self, i):
        out_ptr, itemsize_row = self.pending[i]
        q = self.pipes[i]
        q.fetch_time_into(out_ptr + i * self.Us * itemsize_row)
        self.stats[i * self.Us:(i + 1) * self.Us] = q.fetch_stats()
        self.pending[i] = None

    def step(self, x16_ptr, delays, out_ptr, out_row_bytes):
        """One pass over the shard: x16_ptr -> pinned int16 [U][C][n], delays [U][C], out_ptr -> pinned float32 rows of out_row_bytes.
        Returns immediately after queueing; results of this step are complete after the next step() or drain()."""
        Us, n, C = self.Us, self.n, self.C
        for i, q in enumerate(self.pipes):
            if self.pending[i] is not None:
                self._collect(i)
            q.submit_i16_pointer(x16_ptr + i * Us * C * n * 2, Us, n)
            q.set_delays(delays[i * Us:(i + 1) * Us])
            q.run(True)
            self.pending[i] = (out_ptr, out_row_bytes)

    def drain(self):
        for i in range(self.NP):
            if self.pending[i] is not None:
                self._collect(i)

    def launches(self):
        return sum(int(q.last_timing()["launches"]) for q in self.pipes)

    def gather_stats(self):
        """The single end-of-run exchange: every rank's per-utterance statistics, in rank order -> [U_total][3]."""
        return self.sharding.gather_stats(self.stats)

    def close(self):
        for q in self.pipes:
            q.close()
