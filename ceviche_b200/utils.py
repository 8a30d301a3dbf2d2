"""Caller helpers of the FDTD path with the reference's signatures (ceviche/utils.py:316-332)."""
import numpy as np
import torch


def measure_fields(F, source, steps, probes, component='Ez', verbose=False):
    """ Returns a time series of the measured `component` fields from FDTD `F`
        driven by `source` and measured at `probes` (ceviche/utils.py:316-332).

        `source` is either the reference's callable t -> dense J array (then the loop is the
        reference's own: one forward() per step), or a tuple (profile, waveform[steps]) which runs
        the fused device loop with in-kernel injection and sampling.  The source drives Jz, as in
        the reference.  Returns a (steps, n_probes) numpy array. """
    F.initialize_fields()
    if not isinstance(probes, list):
        probes = [probes]
    if verbose and steps >= 20:       # the reference prints a progress line every steps // 20 time steps (utils.py:326-327)
        print('{:.2f} % done'.format(0.0))
    if isinstance(source, tuple):
        profile, waveform = source
        series = F.run(steps, [("z", profile, np.asarray(waveform)[:steps])], [(component, p) for p in probes])
        return series.detach().cpu().numpy()
    masks = [torch.as_tensor(np.asarray(p, dtype=np.float64)).to(F.device).reshape(F.grid_shape) for p in probes]
    measured = np.zeros((steps, len(probes)))
    for t_index in range(steps):
        if verbose and steps >= 20 and t_index and t_index % (steps // 20) == 0:
            print('{:.2f} % done'.format(float(t_index) / steps * 100.0))
        fields = F.forward(Jz=source(t_index))
        for probe_index, mask in enumerate(masks):
            measured[t_index, probe_index] = float(torch.sum(fields[component].double() * mask))
    return measured


def aniplot(F, source, steps, component='Ez', num_panels=10, show=True):
    """ Animate an FDTD (F) with `source` for `steps` time steps: the `component` field at `num_panels` equally
        spaced times (ceviche/utils.py:279-313; same loop, same panel schedule t % (steps // num_panels) == 0, colour
        limits +-max|field| / 2, z-slice 0).  Returns the list of (t_index, 2-D numpy panel); the figure is drawn only
        where matplotlib is importable (a soft dependency, as HIPS autograd is), and shown if `show`. """
    F.initialize_fields()
    panels = []
    every = max(1, steps // num_panels)
    for t_index in range(steps):
        fields = F.forward(Jz=source(t_index))
        if t_index % every == 0 and len(panels) < num_panels:
            panels.append((t_index, fields[component][:, :, 0].detach().cpu().numpy().copy()))
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return panels
    f, ax_list = plt.subplots(1, num_panels, figsize=(20 * num_panels, 20))
    for ax, (t_index, arr) in zip(np.atleast_1d(ax_list), panels):
        max_E = np.abs(arr).max()
        im = ax.pcolormesh(arr.T, cmap='RdBu')
        im.set_clim([-max_E / 2.0, max_E / 2.0])
        ax.set_title('time = {} seconds'.format(F.dt * t_index))
    if show:
        plt.show()
    return panels


# ---- spectra of probe series (ceviche/utils.py:350-403).  torch.fft is differentiable in both modes,
# so the reference's hand-registered `my_fft` primitive (utils.py:350-370) needs no counterpart.
def my_fft(x):
    """FFT along axis 0... of a 1-D series, like np.fft.fft (utils.py:350-356)."""
    return torch.fft.fft(x)


def get_spectrum(series, dt):
    """ Get FFT of series: Hamming-windowed, along time (utils.py:373-388).  Returns (freqs, signal_f)
    with signal_f of shape (steps, n).  Accepts numpy or torch; stays on the input's device.
    Deviation, on purpose: the reference reshapes to (steps, -1) and then calls np.fft.fft, which
    transforms the LAST axis (length 1 for a single series, i.e. a no-op); the transform here is along
    time, which is what the function documents. """
    if not torch.is_tensor(series):
        series = torch.as_tensor(np.asarray(series))
    steps = series.shape[0]
    series = series.reshape((steps, -1))
    window = torch.as_tensor(np.hamming(steps).reshape((steps, 1)), dtype=series.real.dtype, device=series.device)
    signal_f = torch.fft.fft(window * series, dim=0)
    freqs = torch.as_tensor(np.fft.fftfreq(steps, d=dt), device=series.device)
    return freqs, signal_f


def get_max_power_freq(series, dt):
    """utils.py:391-394 (argmax over the flattened spectrum's real part ordering, as numpy does for complex)."""
    freqs, signal_f = get_spectrum(series, dt)
    flat = signal_f.reshape(-1)
    # numpy's argmax on complex arrays orders lexicographically by (real, imag)
    key = flat.real.double()
    idx = int(torch.argmax(key))
    return freqs[idx % freqs.shape[0]] if signal_f.shape[1] == 1 else freqs[idx // signal_f.shape[1]]


def get_spectral_power(series, dt):
    """utils.py:397-400."""
    freqs, signal_f = get_spectrum(series, dt)
    return freqs, torch.square(torch.abs(signal_f))
