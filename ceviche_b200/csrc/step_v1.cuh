// Baseline half-step kernels: one thread per cell, neighbours read straight from global
// memory (L1/L2 absorb the reuse).  Kept as the simple, obviously-correct formulation:
// the tuned marching kernels in step_v2.cuh are tested against it bit for bit, and it
// serves grids whose contiguous extent cannot be vectorised.
#pragma once
#include "common.cuh"

namespace cev {

constexpr int V1_TZ = 64;
constexpr int V1_TY = 4;

// H half-step, fdtd.py:80-97 with curl_E of derivatives.py:16-22 (forward differences,
// periodic).  E is formed on the fly as E = mE*D (fdtd.py:135-137) at the cell and at its
// +1 neighbours, so E never has to live in HBM.
template <typename T, typename AT>
__global__ void __launch_bounds__(V1_TZ* V1_TY) k_step_H_v1(const StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int i = a.x0 + rest / a.nty;
    const int k = tz * blockDim.x + threadIdx.x;     // block = 64 x 4 threads, or 256 x 1 for single-row planes
    const int j = ty * blockDim.y + threadIdx.y;
    if (k >= a.Nz || j >= a.Ny) return;

    const int64_t plane = (int64_t)a.Ny * a.Nz;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const int kp = (k + 1 == a.Nz) ? 0 : k + 1;
    const int64_t o_in = (int64_t)j * a.Nz + k;      // offset inside a plane
    const int64_t o = i * plane + o_in;
    const int64_t o_jp = i * plane + (int64_t)jp * a.Nz + k;
    const int64_t o_kp = i * plane + (int64_t)j * a.Nz + kp;
    const bool last = (i + 1 == a.Nx);

    // Every load is issued first, predicated on the (kernel-uniform) component mask and tangent flag, and the
    // arithmetic follows: branches around dependent load/multiply pairs would serialise the memory latency.
    struct ERaw { AT m, d, dm, dp; };
    auto e_load = [&](int c, const T* M, const T* D, const T* dM, const T* Dp, int64_t off) {
        const bool on = (a.on >> c) & 1u, tan = on && a.dmE[c];
        ERaw r;
        r.m = on ? (AT)M[off] : AT(0);
        r.d = on ? (AT)D[off] : AT(0);
        r.dm = tan ? (AT)dM[off] : AT(0);
        r.dp = tan ? (AT)Dp[off] : AT(0);
        return r;
    };
    auto e_val = [&](int c, const ERaw& r) {      // E = mE*D (+ dmE*D_primal for a tangent state)
        const AT e = mul_rn(r.m, r.d);
        return a.dmE[c] ? add_rn(e, mul_rn(r.dm, r.dp)) : e;
    };
#define E_HERE(c, off) e_load(c, a.mE[c], a.Din[c], a.dmE[c], a.Dp[c], off)
    const ERaw rx0 = E_HERE(0, o), ry0 = E_HERE(1, o), rz0 = E_HERE(2, o);
    const ERaw rxj = E_HERE(0, o_jp), rzj = E_HERE(2, o_jp);
    const ERaw rxk = E_HERE(0, o_kp), ryk = E_HERE(1, o_kp);
#undef E_HERE
    const int64_t o_ip = last ? o_in : o + plane;     // x+1 plane: inside the array, or the halo / wrap plane
    const ERaw ryi = e_load(1, last ? a.mEhi[1] : a.mE[1], last ? a.Dhi[1] : a.Din[1], last ? a.dmEhi[1] : a.dmE[1],
                            last ? a.Dphi[1] : a.Dp[1], o_ip);
    const ERaw rzi = e_load(2, last ? a.mEhi[2] : a.mE[2], last ? a.Dhi[2] : a.Din[2], last ? a.dmEhi[2] : a.dmE[2],
                            last ? a.Dphi[2] : a.Dp[2], o_ip);
    const AT Ex = e_val(0, rx0), Ey = e_val(1, ry0), Ez = e_val(2, rz0);
    const AT Ex_jp = e_val(0, rxj), Ez_jp = e_val(2, rzj);
    const AT Ex_kp = e_val(0, rxk), Ey_kp = e_val(1, ryk);
    const AT Ey_ip = e_val(1, ryi), Ez_ip = e_val(2, rzi);
    const AT inv = a.inv_dL;
    const AT CEx = curl2<AT>(Ez_jp, Ez, Ey_kp, Ey, inv);
    const AT CEy = curl2<AT>(Ex_kp, Ex, Ez_ip, Ez, inv);
    const AT CEz = curl2<AT>(Ey_ip, Ey, Ex_jp, Ex, inv);

    const AT ux = a.uH[0][i], uy = a.uH[1][j], uz = a.uH[2][k];
    const AT rx = a.rH[0][i], ry = a.rH[1][j], rz = a.rH[2][k];
    const int mx = a.mapH[0][i], my = a.mapH[1][j], mz = a.mapH[2][k];
    const AT s = -a.cdt;

    // A masked-out component keeps its loads, integral updates and store predicated off (straight-line code: the
    // three updates overlap their memory latency).
    const bool hx_on = a.on & 8u, hy_on = a.on & 16u, hz_on = a.on & 32u;
    const AT Hx0 = hx_on ? (AT)a.Hin[0][o] : AT(0), Hy0 = hy_on ? (AT)a.Hin[1][o] : AT(0), Hz0 = hz_on ? (AT)a.Hin[2][o] : AT(0);
    // x component: (a,b) = (y,z), own = x.   ICE_x (nHx,Ny,Nz), IH_x (Nx,nHy,nHz)
    {
        const int64_t ic = (mx >= 0 && hx_on) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0 && hx_on) ? ((int64_t)i * a.nH[1] + my) * a.nH[2] + mz : -1;
        const AT v = update_component<T, AT>(Hx0, CEx, uy, ry, uz, rz, ux, s, a.ICE[0], ic, a.IH[0], is);
        if (hx_on) a.Hout[0][o] = (T)v;
    }
    // y component: (a,b) = (x,z), own = y.   ICE_y (Nx,nHy,Nz), IH_y (nHx,Ny,nHz)
    {
        const int64_t ic = (my >= 0 && hy_on) ? ((int64_t)i * a.nH[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0 && hy_on) ? ((int64_t)mx * a.Ny + j) * a.nH[2] + mz : -1;
        const AT v = update_component<T, AT>(Hy0, CEy, ux, rx, uz, rz, uy, s, a.ICE[1], ic, a.IH[1], is);
        if (hy_on) a.Hout[1][o] = (T)v;
    }
    // z component: (a,b) = (x,y), own = z.   ICE_z (Nx,Ny,nHz), IH_z (nHx,nHy,Nz)
    {
        const int64_t ic = (mz >= 0 && hz_on) ? ((int64_t)i * a.Ny + j) * a.nH[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0 && hz_on) ? ((int64_t)mx * a.nH[1] + my) * a.Nz + k : -1;
        const AT v = update_component<T, AT>(Hz0, CEz, ux, rx, uy, ry, uz, s, a.ICE[2], ic, a.IH[2], is);
        if (hz_on) a.Hout[2][o] = (T)v;
    }
}

// D/E half-step, fdtd.py:105-137 with curl_H of derivatives.py:24-30 (backward differences,
// periodic), dense J added after the update (fdtd.py:125-127), optional E output.
template <typename T, typename AT>
__global__ void __launch_bounds__(V1_TZ* V1_TY) k_step_D_v1(const StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int i = a.x0 + rest / a.nty;
    const int k = tz * blockDim.x + threadIdx.x;     // block = 64 x 4 threads, or 256 x 1 for single-row planes
    const int j = ty * blockDim.y + threadIdx.y;
    if (k >= a.Nz || j >= a.Ny) return;

    const int64_t plane = (int64_t)a.Ny * a.Nz;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const int km = (k == 0) ? a.Nz - 1 : k - 1;
    const int64_t o_in = (int64_t)j * a.Nz + k;
    const int64_t o = i * plane + o_in;
    const int64_t o_jm = i * plane + (int64_t)jm * a.Nz + k;
    const int64_t o_km = i * plane + (int64_t)j * a.Nz + km;

#define H_AT(c, ptr, off) (((a.on >> (3 + (c))) & 1u) ? (AT)(ptr)[off] : AT(0))
    const AT Hx = H_AT(0, a.Hin[0], o), Hy = H_AT(1, a.Hin[1], o), Hz = H_AT(2, a.Hin[2], o);
    const AT Hx_jm = H_AT(0, a.Hin[0], o_jm), Hz_jm = H_AT(2, a.Hin[2], o_jm);
    const AT Hx_km = H_AT(0, a.Hin[0], o_km), Hy_km = H_AT(1, a.Hin[1], o_km);
    AT Hy_im, Hz_im;
    if (i > 0) {
        Hy_im = H_AT(1, a.Hin[1], o - plane);
        Hz_im = H_AT(2, a.Hin[2], o - plane);
    } else {
        Hy_im = H_AT(1, a.Hlo[1], o_in);
        Hz_im = H_AT(2, a.Hlo[2], o_in);
    }
#undef H_AT
    const AT inv = a.inv_dL;
    const AT CHx = curl2<AT>(Hz, Hz_jm, Hy, Hy_km, inv);
    const AT CHy = curl2<AT>(Hx, Hx_km, Hz, Hz_im, inv);
    const AT CHz = curl2<AT>(Hy, Hy_im, Hx, Hx_jm, inv);

    const AT ux = a.uD[0][i], uy = a.uD[1][j], uz = a.uD[2][k];
    const AT rx = a.rD[0][i], ry = a.rD[1][j], rz = a.rD[2][k];
    const int mx = a.mapD[0][i], my = a.mapD[1][j], mz = a.mapD[2][k];
    const AT s = a.cdt;
    // masked-out components: loads and integral updates predicated off, straight-line code (see k_step_H_v1)
    const bool dx_on = a.on & 1u, dy_on = a.on & 2u, dz_on = a.on & 4u;
    const AT Dx0 = dx_on ? (AT)a.Din[0][o] : AT(0), Dy0 = dy_on ? (AT)a.Din[1][o] : AT(0), Dz0 = dz_on ? (AT)a.Din[2][o] : AT(0);
    AT Dn[3];
    {
        const int64_t ic = (mx >= 0 && dx_on) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0 && dx_on) ? ((int64_t)i * a.nD[1] + my) * a.nD[2] + mz : -1;
        Dn[0] = update_component<T, AT>(Dx0, CHx, uy, ry, uz, rz, ux, s, a.ICH[0], ic, a.ID[0], is);
    }
    {
        const int64_t ic = (my >= 0 && dy_on) ? ((int64_t)i * a.nD[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0 && dy_on) ? ((int64_t)mx * a.Ny + j) * a.nD[2] + mz : -1;
        Dn[1] = update_component<T, AT>(Dy0, CHy, ux, rx, uz, rz, uy, s, a.ICH[1], ic, a.ID[1], is);
    }
    {
        const int64_t ic = (mz >= 0 && dz_on) ? ((int64_t)i * a.Ny + j) * a.nD[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0 && dz_on) ? ((int64_t)mx * a.nD[1] + my) * a.Nz + k : -1;
        Dn[2] = update_component<T, AT>(Dz0, CHz, ux, rx, uy, ry, uz, s, a.ICH[2], ic, a.ID[2], is);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (!((a.on >> c) & 1u)) continue;      // identically-zero component: nothing to read or write
        AT d = Dn[c];
        if (a.J[c]) {
            const AT sc = a.Jwave[c] ? (AT)(*a.Jwave[c]) : a.Jscale[c];
            d = add_rn(d, mul_rn((AT)a.J[c][o], sc));
        }
        a.Dout[c][o] = (T)d;
        if (a.Eout[c]) a.Eout[c][o] = (T)mul_rn((AT)a.mE[c][o], (AT)(T)d);  // E from the STORED D, as fdtd.py:135
    }
}

// E = mE * D  (fdtd.py:135-137), plain streaming kernel.
template <typename T, typename AT>
__global__ void k_compute_E(const T* __restrict__ mE, const T* __restrict__ D, const T* __restrict__ dmE,
                            const T* __restrict__ Dp, T* __restrict__ E, int64_t n) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        AT e = mul_rn((AT)mE[q], (AT)D[q]);
        if (dmE) e = add_rn(e, mul_rn((AT)dmE[q], (AT)Dp[q]));   // tangent: dE = mE dD + dmE D
        E[q] = (T)e;
    }
}

// Sparse J injection: D[field][cell] += weight * waveform[src]   (fdtd.py:125-127 for
// J = profile * scalar(t), the only form the reference's callers use, utils.py:328).
struct SourceTable {
    int64_t        n;            // total points
    const int32_t* comp;         // internal component per point
    const int32_t* src;          // source id per point (column of the waveform row)
    const int64_t* cell;
    const double*  weight;
};

template <typename T, typename AT>
__global__ void k_inject(SourceTable s, T* D0, T* D1, T* D2, const double* __restrict__ wave_row, int64_t cell_lo,
                         int64_t cell_hi) {
    // (the plan keeps its points sorted by (component, cell): see inject_points)
    inject_points<T, AT, int64_t>(0, s.n, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, s.comp, s.src, s.cell,
                                  s.weight, wave_row, D0, D1, D2, cell_lo, cell_hi);
}

// Stand-alone probe sampling (after the last step of a run).
template <typename T, typename AT>
__global__ void k_probe_only(const StepArgs<T, AT> a) {
    probe_block<T, AT>(a, a.aux_slot0 + blockIdx.x);
}

// Running-DFT monitors: acc[q, f] += field(point q) * phasor[f] once per time step (the frequency-domain field
// of a region without storing its time series; the natural producer of mode-overlap objectives,
// ceviche/utils.py:373-400 computes the same sums from stored series with an FFT).  One thread per point; the
// accumulators are complex fp64 (re, im interleaved).
struct MonitorTable {
    int64_t        n;        // total points of all monitors
    int            nfreq;
    const int32_t* field;    // per point: CEV_FIELD_* + internal component
    const int64_t* cell;
};
template <typename T, typename AT>
__global__ void k_monitor(const StepArgs<T, AT> a, MonitorTable m, const double* __restrict__ phasor_row,
                          double* __restrict__ acc) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m.n) return;
    const double v = (double)probe_value<T, AT>(a, m.field[q], m.cell[q]);
    double* o = acc + q * m.nfreq * 2;
    for (int f = 0; f < m.nfreq; ++f) {
        o[2 * f] += v * phasor_row[2 * f];
        o[2 * f + 1] += v * phasor_row[2 * f + 1];
    }
}

}  // namespace cev
