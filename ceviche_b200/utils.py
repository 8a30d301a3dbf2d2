"""Caller helpers of the FDTD path with the reference's signatures (ceviche/utils.py:316-332)."""
import numpy as np
import torch


def measure_fields(F, source, steps, probes, component='Ez'):
    """ Returns a time series of the measured `component` fields from FDTD `F`
        driven by `source` and measured at `probes` (ceviche/utils.py:316-332).

        `source` is either the reference's callable t -> dense J array (then the loop is the
        reference's own: one forward() per step), or a tuple (profile, waveform[steps]) which runs
        the fused device loop with in-kernel injection and sampling.  The source drives Jz, as in
        the reference.  Returns a (steps, n_probes) numpy array. """
    F.initialize_fields()
    if not isinstance(probes, list):
        probes = [probes]
    if isinstance(source, tuple):
        profile, waveform = source
        series = F.run(steps, [("z", profile, np.asarray(waveform)[:steps])], [(component, p) for p in probes])
        return series.detach().cpu().numpy()
    masks = [torch.as_tensor(np.asarray(p, dtype=np.float64)).to(F.device).reshape(F.grid_shape) for p in probes]
    measured = np.zeros((steps, len(probes)))
    for t_index in range(steps):
        fields = F.forward(Jz=source(t_index))
        for probe_index, mask in enumerate(masks):
            measured[t_index, probe_index] = float(torch.sum(fields[component].double() * mask))
    return measured
