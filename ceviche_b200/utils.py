"""Caller helpers of the FDTD path with the reference's signatures: the caller loops (ceviche/utils.py:279-332), the
Yee-grid averaging the `eps_r` setter is built on (utils.py:153-198), the shape / value helpers and finite-difference
checkers its tests use (utils.py:108-150, 200-236), and the spectrum helpers (utils.py:335-411).  Everything accepts numpy
arrays or torch tensors (any device) and answers in kind."""
import numpy as np
import torch

from .fdtd import reshape_to_ND          # noqa: F401  (utils.py:206-214; lives next to the object that uses it)


def _roll(q, shift, axis):
    return torch.roll(q, shift, axis) if torch.is_tensor(q) else np.roll(q, shift, axis)


# ---- Yee-grid averaging (ceviche/utils.py:153-198)
def grid_center_to_xyz(Q_mid, averaging=True):
    """ The cell-centred quantity `Q_mid` (Nx, Ny, Nz) as felt at the Ex, Ey, Ez positions of the Yee lattice: the mean with
        the previous cell along x, y, z (periodic); `averaging=False` returns three copies (utils.py:153-176).  This is how
        `fdtd.eps_r` becomes eps_xx / eps_yy / eps_zz. """
    if not averaging:
        dup = (lambda q: q.clone()) if torch.is_tensor(Q_mid) else (lambda q: np.array(q, copy=True))
        return dup(Q_mid), dup(Q_mid), dup(Q_mid)
    return tuple((Q_mid + _roll(Q_mid, 1, axis)) / 2 for axis in range(3))


def grid_xyz_to_center(Q_xx, Q_yy, Q_zz):
    """ The same one-cell average applied to three component arrays, each along its own axis (utils.py:179-189; as in the
        reference the shift is +1 again -- it is the same operator, not its inverse). """
    as_float = lambda q: q.to(torch.float64) if torch.is_tensor(q) else np.asarray(q).astype('float')
    return tuple((as_float(q) + _roll(q, 1, axis)) / 2 for axis, q in enumerate((Q_xx, Q_yy, Q_zz)))


def vec_zz_to_xy(info_dict, vec_zz, grid_averaging=True):
    """ Grid averaging of a flattened 2-D zz quantity onto the x and y positions (utils.py:191-196). """
    arr_zz = vec_zz.reshape(tuple(info_dict['shape']))[:, :, None]
    arr_xx, arr_yy, _ = grid_center_to_xyz(arr_zz, averaging=grid_averaging)
    return arr_xx.flatten(), arr_yy.flatten()


# ---- shape / value helpers (ceviche/utils.py:200-236)
def float_2_array(x):
    """ A scalar becomes a 1-element array; arrays and tensors pass through (utils.py:200-204). """
    return x if isinstance(x, np.ndarray) or torch.is_tensor(x) else np.array([x])


def get_value(x):
    """ The plain value of a traced quantity (utils.py:220-224 unboxes a HIPS ArrayBox): here a tensor's detached value
        (its primal under forward-mode AD); anything else passes through. """
    if torch.is_tensor(x):
        import torch.autograd.forward_ad as fwAD
        x = x.detach()
        return fwAD.unpack_dual(x).primal if fwAD._current_level >= 0 else x
    return x


def get_shape(x):
    """ The shape of x, even if it is not an array (utils.py:229-236). """
    if isinstance(x, (float, int)):
        return (1,)
    if isinstance(x, (tuple, list)):
        return (len(x),)
    return tuple(x.shape)


# ---- finite-difference checkers (ceviche/utils.py:108-150)
def _copy(a):
    return a.clone() if torch.is_tensor(a) else np.array(a, copy=True)


def der_num(fn, arg, index, delta):
    """ Central difference of fn(arg) w.r.t. arg[index] with step `delta` (utils.py:108-115). """
    fwd, back = _copy(arg), _copy(arg)
    fwd[index] += delta / 2
    back[index] -= delta / 2
    return (fn(fwd) - fn(back)) / delta


def grad_num(fn, arg, delta=1e-6):
    """ Numerical gradient of a scalar function, one central difference per entry of `arg` (utils.py:117-125).  For a
        complex `arg` the derivative along the imaginary axis is added as the reference does (real part + imaginary part of
        the step); for real arguments -- every FDTD use -- it is the plain real gradient, returned as complex128 like the
        reference's. """
    n = int(arg.numel() if torch.is_tensor(arg) else np.size(arg))
    is_complex = arg.is_complex() if torch.is_tensor(arg) else np.iscomplexobj(arg)
    grad = np.zeros((n,), dtype=np.complex128)
    for i in range(n):
        g = der_num(fn, arg, i, delta)
        if is_complex:
            g = g + der_num(fn, arg, i, 1j * delta)
        grad[i] = complex(g)
    return grad


def jac_num(fn, arg, step_size=1e-7):
    """ DEPRECATED in the reference in favour of jacobian(mode='numerical'); kept for its callers (utils.py:127-150):
        one-sided differences, returned as (n_in, n_out) -- the TRANSPOSE of jacobian()'s layout, as in the reference. """
    to_np = lambda v: v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v, dtype=np.float64)
    in_array = float_2_array(arg).flatten()
    out_array = to_np(float_2_array(fn(arg))).flatten()
    jac = np.zeros((int(in_array.shape[0]), out_array.size))
    for i in range(jac.shape[0]):
        input_i = _copy(in_array)
        input_i[i] += step_size
        output_i = to_np(float_2_array(fn(input_i.reshape(in_array.shape)))).flatten()
        jac[i, :] = (output_i - out_array) / step_size
    return jac


def imarr(arr):
    """ Array -> the orientation matplotlib's imshow wants: z-slice 0 of a 3-D array, transposed, flipped (utils.py:335-341). """
    a = get_value(arr)
    a = a.cpu().numpy() if torch.is_tensor(a) else np.array(a, copy=True)
    if a.ndim == 3:
        a = a[:, :, 0]
    return np.flipud(a.T)


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


def plot_spectral_power(series, dt, f_top=2e14, show=True):
    """ Plot |FFT|^2 of a series up to `f_top` (utils.py:402-411); returns (freqs, power) of the plotted half.  matplotlib
        is a soft dependency: without it nothing is drawn. """
    steps = len(series)
    freqs, power = get_spectral_power(series, dt)
    freqs, power = freqs[:steps // 2].cpu().numpy(), power[:steps // 2].cpu().numpy()
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return freqs, power
    plt.plot(freqs, power)
    plt.xlim([0, f_top])
    plt.xlabel('frequency (Hz)')
    plt.ylabel('power (|signal|^2)')
    if show:
        plt.show()
    return freqs, power
