"""TEST INFRASTRUCTURE (oracle) — never imported by the product path (stim_b200/).

CPU restatement of the reference's measurement -> detection-event conversion,
measurements_to_detection_events_helper (/root/reference/src/stim/simulators/measurements_to_detection_events.inl:30-131):
walk the noiseless circuit; a detector row = XOR of the recorded measurement rows it names (:84-90), inverted when the
reference sample's parity over the same measurements is 1 (:91-93), XORed with the detector-flip row of a frame simulation
that only sees the sweep bits, frame randomisation off (:60-63, :79-80); observables alike (:99-119, :126-131).

Parity status: PINNED by tests/golden/m2d_cases.json (outputs of the reference CLI `stim m2d`, tools/gen_m2d_golden.py)."""
import numpy as np

from . import frame_oracle as fo


class _SweepOnlyOracle(fo.FrameOracle):
    """The frame simulator of m2d: no noise, no collapse randomisation, a sweep table; records which measurements every
    detector / observable names."""

    def __init__(self, text, n_shots, sweep_bits):
        K = max(1, (n_shots + 127) // 128)
        super().__init__(text, 0, K, 1, 0)
        self.sweep_rows = {}
        if sweep_bits is not None:
            for k in range(sweep_bits.shape[1]):
                col = np.zeros(self.W * 32, dtype=np.uint8)
                col[:n_shots] = sweep_bits[:, k]
                self.sweep_rows[k] = np.packbits(col, bitorder="little").view(np.uint32).copy()
        self.det_recs, self.obs_recs = [], {}

    randomize = False  # guarantee_anticommutation_via_frame_randomization = false (:63)

    def run_sites(self, clocks, lam, group, on_event):
        return  # aliased_noiseless_circuit()

    def do_op(self, name, args, targets):
        if name == "DETECTOR":
            self.det_recs.append([len(self.rec) - (t & fo.T_VAL) for t in targets])
        elif name == "OBSERVABLE_INCLUDE":
            self.obs_recs.setdefault(int(args[0]), []).extend(len(self.rec) - (t & fo.T_VAL) for t in targets if t & fo.T_REC)
        super().do_op(name, args, targets)


def convert(text, measurements, sweep_bits=None, reference_sample=None, append_observables=False):
    """measurements uint8/bool [shots, M] (one byte per bit), sweep_bits [shots, S] or None, reference_sample [M] or None
    (None = all zero, i.e. skip_reference_sample). Returns (dets [shots, D], obs [shots, L]) as uint8, or one array with
    the observables appended."""
    measurements = np.asarray(measurements).astype(np.uint8)
    shots, M = measurements.shape
    o = _SweepOnlyOracle(text, shots, None if sweep_bits is None else np.asarray(sweep_bits).astype(np.uint8)).run()
    assert len(o.rec) == M, (len(o.rec), M)
    ref = np.zeros(M, dtype=np.uint8) if reference_sample is None else np.asarray(reference_sample).astype(np.uint8)
    dets = o.detectors()[:shots].copy()
    obs = o.observables()[:shots].copy()
    n_obs = (max(o.obs_recs) + 1) if o.obs_recs else 0
    if obs.shape[1] < n_obs:
        obs = np.concatenate([obs, np.zeros((shots, n_obs - obs.shape[1]), dtype=np.uint8)], axis=1)
    for d, recs in enumerate(o.det_recs):
        for m in recs:
            dets[:, d] ^= measurements[:, m] ^ ref[m]
    for l, recs in o.obs_recs.items():
        for m in recs:
            obs[:, l] ^= measurements[:, m] ^ ref[m]
    if append_observables:
        return np.concatenate([dets, obs], axis=1)
    return dets, obs
