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
    const int k = tz * V1_TZ + threadIdx.x;
    const int j = ty * V1_TY + threadIdx.y;
    if (k >= a.Nz || j >= a.Ny) return;

    const int64_t plane = (int64_t)a.Ny * a.Nz;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const int kp = (k + 1 == a.Nz) ? 0 : k + 1;
    const int64_t o_in = (int64_t)j * a.Nz + k;      // offset inside a plane
    const int64_t o = i * plane + o_in;
    const int64_t o_jp = i * plane + (int64_t)jp * a.Nz + k;
    const int64_t o_kp = i * plane + (int64_t)j * a.Nz + kp;
    const bool last = (i + 1 == a.Nx);

#define E_AT(c, off)                                                                        \
    (a.dmE[c] ? add_rn(mul_rn((AT)a.mE[c][off], (AT)a.Din[c][off]), mul_rn((AT)a.dmE[c][off], (AT)a.Dp[c][off])) \
              : mul_rn((AT)a.mE[c][off], (AT)a.Din[c][off]))
    const AT Ex = E_AT(0, o), Ey = E_AT(1, o), Ez = E_AT(2, o);
    const AT Ex_jp = E_AT(0, o_jp), Ez_jp = E_AT(2, o_jp);
    const AT Ex_kp = E_AT(0, o_kp), Ey_kp = E_AT(1, o_kp);
    AT Ey_ip, Ez_ip;
    if (!last) {
        Ey_ip = E_AT(1, o + plane);
        Ez_ip = E_AT(2, o + plane);
    } else {
        Ey_ip = mul_rn((AT)a.mEhi[1][o_in], (AT)a.Dhi[1][o_in]);
        Ez_ip = mul_rn((AT)a.mEhi[2][o_in], (AT)a.Dhi[2][o_in]);
        if (a.dmE[1]) Ey_ip = add_rn(Ey_ip, mul_rn((AT)a.dmEhi[1][o_in], (AT)a.Dphi[1][o_in]));
        if (a.dmE[2]) Ez_ip = add_rn(Ez_ip, mul_rn((AT)a.dmEhi[2][o_in], (AT)a.Dphi[2][o_in]));
    }
#undef E_AT
    const AT inv = a.inv_dL;
    const AT CEx = curl2<AT>(Ez_jp, Ez, Ey_kp, Ey, inv);
    const AT CEy = curl2<AT>(Ex_kp, Ex, Ez_ip, Ez, inv);
    const AT CEz = curl2<AT>(Ey_ip, Ey, Ex_jp, Ex, inv);

    const AT ux = a.uH[0][i], uy = a.uH[1][j], uz = a.uH[2][k];
    const AT rx = a.rH[0][i], ry = a.rH[1][j], rz = a.rH[2][k];
    const int mx = a.mapH[0][i], my = a.mapH[1][j], mz = a.mapH[2][k];
    const AT s = -a.cdt;

    // x component: (a,b) = (y,z), own = x.   ICE_x (nHx,Ny,Nz), IH_x (Nx,nHy,nHz)
    {
        const int64_t ic = (mx >= 0) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0) ? ((int64_t)i * a.nH[1] + my) * a.nH[2] + mz : -1;
        a.Hout[0][o] = (T)update_component<T, AT>((AT)a.Hin[0][o], CEx, uy, ry, uz, rz, ux, s, a.ICE[0], ic, a.IH[0], is);
    }
    // y component: (a,b) = (x,z), own = y.   ICE_y (Nx,nHy,Nz), IH_y (nHx,Ny,nHz)
    {
        const int64_t ic = (my >= 0) ? ((int64_t)i * a.nH[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0) ? ((int64_t)mx * a.Ny + j) * a.nH[2] + mz : -1;
        a.Hout[1][o] = (T)update_component<T, AT>((AT)a.Hin[1][o], CEy, ux, rx, uz, rz, uy, s, a.ICE[1], ic, a.IH[1], is);
    }
    // z component: (a,b) = (x,y), own = z.   ICE_z (Nx,Ny,nHz), IH_z (nHx,nHy,Nz)
    {
        const int64_t ic = (mz >= 0) ? ((int64_t)i * a.Ny + j) * a.nH[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0) ? ((int64_t)mx * a.nH[1] + my) * a.Nz + k : -1;
        a.Hout[2][o] = (T)update_component<T, AT>((AT)a.Hin[2][o], CEz, ux, rx, uy, ry, uz, s, a.ICE[2], ic, a.IH[2], is);
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
    const int k = tz * V1_TZ + threadIdx.x;
    const int j = ty * V1_TY + threadIdx.y;
    if (k >= a.Nz || j >= a.Ny) return;

    const int64_t plane = (int64_t)a.Ny * a.Nz;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const int km = (k == 0) ? a.Nz - 1 : k - 1;
    const int64_t o_in = (int64_t)j * a.Nz + k;
    const int64_t o = i * plane + o_in;
    const int64_t o_jm = i * plane + (int64_t)jm * a.Nz + k;
    const int64_t o_km = i * plane + (int64_t)j * a.Nz + km;

    const AT Hx = (AT)a.Hin[0][o], Hy = (AT)a.Hin[1][o], Hz = (AT)a.Hin[2][o];
    const AT Hx_jm = (AT)a.Hin[0][o_jm], Hz_jm = (AT)a.Hin[2][o_jm];
    const AT Hx_km = (AT)a.Hin[0][o_km], Hy_km = (AT)a.Hin[1][o_km];
    AT Hy_im, Hz_im;
    if (i > 0) {
        Hy_im = (AT)a.Hin[1][o - plane];
        Hz_im = (AT)a.Hin[2][o - plane];
    } else {
        Hy_im = (AT)a.Hlo[1][o_in];
        Hz_im = (AT)a.Hlo[2][o_in];
    }
    const AT inv = a.inv_dL;
    const AT CHx = curl2<AT>(Hz, Hz_jm, Hy, Hy_km, inv);
    const AT CHy = curl2<AT>(Hx, Hx_km, Hz, Hz_im, inv);
    const AT CHz = curl2<AT>(Hy, Hy_im, Hx, Hx_jm, inv);

    const AT ux = a.uD[0][i], uy = a.uD[1][j], uz = a.uD[2][k];
    const AT rx = a.rD[0][i], ry = a.rD[1][j], rz = a.rD[2][k];
    const int mx = a.mapD[0][i], my = a.mapD[1][j], mz = a.mapD[2][k];
    const AT s = a.cdt;
    AT Dn[3];
    {
        const int64_t ic = (mx >= 0) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0) ? ((int64_t)i * a.nD[1] + my) * a.nD[2] + mz : -1;
        Dn[0] = update_component<T, AT>((AT)a.Din[0][o], CHx, uy, ry, uz, rz, ux, s, a.ICH[0], ic, a.ID[0], is);
    }
    {
        const int64_t ic = (my >= 0) ? ((int64_t)i * a.nD[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0) ? ((int64_t)mx * a.Ny + j) * a.nD[2] + mz : -1;
        Dn[1] = update_component<T, AT>((AT)a.Din[1][o], CHy, ux, rx, uz, rz, uy, s, a.ICH[1], ic, a.ID[1], is);
    }
    {
        const int64_t ic = (mz >= 0) ? ((int64_t)i * a.Ny + j) * a.nD[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0) ? ((int64_t)mx * a.nD[1] + my) * a.Nz + k : -1;
        Dn[2] = update_component<T, AT>((AT)a.Din[2][o], CHz, ux, rx, uy, ry, uz, s, a.ICH[2], ic, a.ID[2], is);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
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
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= s.n) return;
    if (s.cell[q] < cell_lo || s.cell[q] >= cell_hi) return;   // only the x-planes this launch updated
    T* D = s.comp[q] == 0 ? D0 : (s.comp[q] == 1 ? D1 : D2);
    const T add = (T)(s.weight[q] * wave_row[s.src[q]]);
    atomicAdd(&D[s.cell[q]], add);
}

// Stand-alone probe sampling (after the last step of a run).
template <typename T, typename AT>
__global__ void k_probe_only(const StepArgs<T, AT> a) {
    probe_block<T, AT>(a, a.aux_slot0 + blockIdx.x);
}

}  // namespace cev
