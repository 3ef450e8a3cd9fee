"""A sinter-shaped consumer of the sampler (SURVEY §8 f rank 1): the sample -> decode -> count loop of
sinter's _CompiledStimThenDecodeSampler.sample (/root/reference/glue/sample/src/sinter/_decoding/_stim_then_decode_sampler.py:
162-223) with the shots kept on the GPU. The reference loop pulls bit-packed detection events and observables to the host,
counts detection events, drops post-selected shots, calls the decoder and classifies errors in numpy; here the detection
events and observables are sampled into device memory (CompiledDetectorSampler.sample_torch: no PCIe drain) and the same
counting / post-selection / classification runs on them as torch bit operations, so only the decoder's verdict per shot has
to be produced — by a decoder that lives on the GPU, or by a host decoder fed from the device rows.

torch is used for the reductions only (popcounts, masks); the sampling is the library's CUDA path."""
import collections
import time
from typing import Callable, Optional

import numpy as np


class AnonTaskStats(dict):
    """shots / errors / discards / seconds / custom_counts, like sinter.AnonTaskStats."""

    __getattr__ = dict.__getitem__


class StimThenDecodeSampler:
    def __init__(self, circuit, decode_shots_bit_packed: Callable, *, count_detection_events: bool = False,
                 postselection_mask: Optional[np.ndarray] = None, postselected_observables_mask: Optional[np.ndarray] = None,
                 seed=None, device: int = 0):
        """decode_shots_bit_packed(dets) takes the bit-packed detection events as a CUDA uint8 tensor [shots, ceil(D / 8)] and
        returns bit-packed observable predictions [shots, ceil(L / 8)] (CUDA tensor or numpy array)."""
        self.sampler = circuit.compile_detector_sampler(seed=seed, device=device)
        self.decode = decode_shots_bit_packed
        self.count_detection_events = count_detection_events
        self.num_det, self.num_obs = circuit.num_detectors, circuit.num_observables
        self.postselection_mask = postselection_mask
        self.postselected_observables_mask = postselected_observables_mask

    def sample(self, max_shots: int) -> AnonTaskStats:
        import torch

        t0 = time.monotonic()
        dets, actual_obs = self.sampler.sample_torch(max_shots, separate_observables=True)  # device-resident, bit-packed
        num_shots = dets.shape[0]
        custom_counts = collections.Counter()
        if self.count_detection_events:
            custom_counts["detectors_checked"] += self.num_det * num_shots
            n = 0
            for b in range(8):  # (the reference's loop, _stim_then_decode_sampler.py:171-172, on the device)
                n += int(torch.count_nonzero(dets & (1 << b)).item())
            custom_counts["detection_events"] += n
        num_discards_1 = 0
        if self.postselection_mask is not None:
            mask = torch.as_tensor(np.asarray(self.postselection_mask, dtype=np.uint8), device=dets.device)
            discarded = torch.any((dets & mask) != 0, dim=1)
            num_discards_1 = int(torch.count_nonzero(discarded).item())
            if num_discards_1:
                dets, actual_obs = dets[~discarded], actual_obs[~discarded]
        predictions = self.decode(dets)
        if not isinstance(predictions, torch.Tensor):
            predictions = torch.as_tensor(np.asarray(predictions), device=dets.device)
        if predictions.dtype != torch.uint8 or predictions.ndim != 2:
            raise ValueError("predictions must be a 2d uint8 array")
        if predictions.shape[0] != num_shots - num_discards_1 or predictions.shape[1] < actual_obs.shape[1]:
            raise ValueError("predictions has the wrong shape")
        # classify_discards_and_errors (sinter/_decoding/_stim_then_decode_sampler.py:35-83), on the device
        wrong = predictions[:, : actual_obs.shape[1]] ^ actual_obs
        num_discards_2 = 0
        if self.postselected_observables_mask is not None:
            pm = torch.as_tensor(np.asarray(self.postselected_observables_mask, dtype=np.uint8), device=dets.device)
            dropped = torch.any((wrong & pm) != 0, dim=1)
            num_discards_2 = int(torch.count_nonzero(dropped).item())
            wrong = wrong[~dropped]
        num_errors = int(torch.count_nonzero(torch.any(wrong != 0, dim=1)).item())
        return AnonTaskStats(shots=num_shots, errors=num_errors, discards=num_discards_1 + num_discards_2,
                             seconds=time.monotonic() - t0, custom_counts=custom_counts)
