// Reverse-mode (VJP) kernels: the time-reversed adjoint FDTD step.
//
// The reference differentiates ceviche/fdtd.py:74-144 by taping every numpy op with HIPS autograd
// (ceviche/jacobians.py:29-35).  Here the transposed step is written out.  With periodic
// boundaries curl_E^T = curl_H and curl_H^T = curl_E (forward difference transposed = minus the
// backward difference), so the adjoint sweeps reuse the two stencils with their roles swapped.
// For step n, given the cotangents (lH, lD, lICE, lIH, lICH, lID) of the state after step n:
//
//   adj_D (cell-local)   gD = lD ; gICH = lICH + m3D gD ; gID = lID + m4D gD
//                        gC = m2D gD + gICH ; lD <- m1D gD + gID ; lICH <- gICH ; lID <- gID
//   adj_H (stencil)      gH = lH + curl_E(gC) ; gICE = lICE + m3H gH ; gIH = lIH + m4H gH
//                        gC2 = m2H gH + gICE ; lH <- m1H gH + gIH ; lICE <- gICE ; lIH <- gIH
//   adj_E (stencil)      lE = curl_H(gC2) ; lD += mE lE ; G_mE += lE * D_{n-1}
//
// after which (lH, lD, ...) are the cotangents of the state after step n-1 and G_mE has gained
// step n's contribution to dL/d(1/eps).  eps_r enters the step only through mE = 1/eps_yee
// (fdtd.py:67, 314-316), so G_mE is the whole gradient; the chain to eps_r is done by the host.
#pragma once
#include "common.cuh"

namespace cev {

template <typename T, typename AT>
struct AdjArgs {
    int Nx, Ny, Nz;
    T* lH[3];
    T* lD[3];
    T* lICE[3];
    T* lIH[3];
    T* lICH[3];
    T* lID[3];
    T* gC[3];        // scratch: cotangent of curl_H(H_n)
    T* gC2[3];       // scratch: cotangent of curl_E(E_{n-1})
    double* G[3];    // dL/d(mE), accumulated in fp64 whatever the storage type
    const T* mE[3];
    const T* Dprev[3];   // forward D after step n-1
    const int* mapH[3];
    const int* mapD[3];
    int nH[3], nD[3];
    const AT* uH[3];
    const AT* rH[3];
    const AT* uD[3];
    const AT* rD[3];
    AT cdt, inv_dL;
};

// transposed component update, cell-local.  Returns gC; updates l, lIcurl, lIself in place.
template <typename T, typename AT>
__device__ __forceinline__ AT adj_component(AT g, AT ua, AT ra, AT ub, AT rb, AT uc, AT scdt, T* l, int64_t o,
                                            T* lIcurl, int64_t icurl, T* lIself, int64_t iself) {
    AT m1, m2;
    coef12<AT>(ua, ra, ub, rb, scdt, m1, m2);
    AT gIc = AT(0), gIs = AT(0);
    if (icurl >= 0) {
        const AT m3 = mul_rn(mul_rn(scdt, uc + uc), mul_rn(ra, rb));
        gIc = (AT)lIcurl[icurl] + m3 * g;
        lIcurl[icurl] = (T)gIc;
    }
    if (iself >= 0) {
        const AT m4 = mul_rn(mul_rn(mul_rn(AT(-4), ua), ub), mul_rn(ra, rb));
        gIs = (AT)lIself[iself] + m4 * g;
        lIself[iself] = (T)gIs;
    }
    l[o] = (T)(m1 * g + gIs);
    return m2 * g + gIc;
}

#define CEV_CELL_INDEX()                                             \
    const int k = blockIdx.x * blockDim.x + threadIdx.x;             \
    const int j = blockIdx.y * blockDim.y + threadIdx.y;             \
    const int i = blockIdx.z;                                        \
    if (k >= a.Nz || j >= a.Ny) return;                              \
    const int64_t plane = (int64_t)a.Ny * a.Nz;                      \
    const int64_t o = i * plane + (int64_t)j * a.Nz + k;

template <typename T, typename AT>
__global__ void k_adj_D(const AdjArgs<T, AT> a) {
    CEV_CELL_INDEX();
    const AT ux = a.uD[0][i], uy = a.uD[1][j], uz = a.uD[2][k];
    const AT rx = a.rD[0][i], ry = a.rD[1][j], rz = a.rD[2][k];
    const int mx = a.mapD[0][i], my = a.mapD[1][j], mz = a.mapD[2][k];
    const AT s = a.cdt;
    {
        const int64_t ic = (mx >= 0) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0) ? ((int64_t)i * a.nD[1] + my) * a.nD[2] + mz : -1;
        a.gC[0][o] = (T)adj_component<T, AT>((AT)a.lD[0][o], uy, ry, uz, rz, ux, s, a.lD[0], o, a.lICH[0], ic, a.lID[0], is);
    }
    {
        const int64_t ic = (my >= 0) ? ((int64_t)i * a.nD[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0) ? ((int64_t)mx * a.Ny + j) * a.nD[2] + mz : -1;
        a.gC[1][o] = (T)adj_component<T, AT>((AT)a.lD[1][o], ux, rx, uz, rz, uy, s, a.lD[1], o, a.lICH[1], ic, a.lID[1], is);
    }
    {
        const int64_t ic = (mz >= 0) ? ((int64_t)i * a.Ny + j) * a.nD[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0) ? ((int64_t)mx * a.nD[1] + my) * a.Nz + k : -1;
        a.gC[2][o] = (T)adj_component<T, AT>((AT)a.lD[2][o], ux, rx, uy, ry, uz, s, a.lD[2], o, a.lICH[2], ic, a.lID[2], is);
    }
}

template <typename T, typename AT>
__global__ void k_adj_H(const AdjArgs<T, AT> a) {
    CEV_CELL_INDEX();
    const int ip = (i + 1 == a.Nx) ? 0 : i + 1;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const int kp = (k + 1 == a.Nz) ? 0 : k + 1;
    const int64_t o_ip = ip * plane + (int64_t)j * a.Nz + k;
    const int64_t o_jp = i * plane + (int64_t)jp * a.Nz + k;
    const int64_t o_kp = i * plane + (int64_t)j * a.Nz + kp;
    const AT inv = a.inv_dL;
    // curl_E (forward differences, derivatives.py:16-22) of gC
    const AT cx = ((AT)a.gC[2][o_jp] - (AT)a.gC[2][o]) * inv - ((AT)a.gC[1][o_kp] - (AT)a.gC[1][o]) * inv;
    const AT cy = ((AT)a.gC[0][o_kp] - (AT)a.gC[0][o]) * inv - ((AT)a.gC[2][o_ip] - (AT)a.gC[2][o]) * inv;
    const AT cz = ((AT)a.gC[1][o_ip] - (AT)a.gC[1][o]) * inv - ((AT)a.gC[0][o_jp] - (AT)a.gC[0][o]) * inv;
    const AT ux = a.uH[0][i], uy = a.uH[1][j], uz = a.uH[2][k];
    const AT rx = a.rH[0][i], ry = a.rH[1][j], rz = a.rH[2][k];
    const int mx = a.mapH[0][i], my = a.mapH[1][j], mz = a.mapH[2][k];
    const AT s = -a.cdt;
    {
        const int64_t ic = (mx >= 0) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0) ? ((int64_t)i * a.nH[1] + my) * a.nH[2] + mz : -1;
        a.gC2[0][o] = (T)adj_component<T, AT>((AT)a.lH[0][o] + cx, uy, ry, uz, rz, ux, s, a.lH[0], o, a.lICE[0], ic, a.lIH[0], is);
    }
    {
        const int64_t ic = (my >= 0) ? ((int64_t)i * a.nH[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0) ? ((int64_t)mx * a.Ny + j) * a.nH[2] + mz : -1;
        a.gC2[1][o] = (T)adj_component<T, AT>((AT)a.lH[1][o] + cy, ux, rx, uz, rz, uy, s, a.lH[1], o, a.lICE[1], ic, a.lIH[1], is);
    }
    {
        const int64_t ic = (mz >= 0) ? ((int64_t)i * a.Ny + j) * a.nH[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0) ? ((int64_t)mx * a.nH[1] + my) * a.Nz + k : -1;
        a.gC2[2][o] = (T)adj_component<T, AT>((AT)a.lH[2][o] + cz, ux, rx, uy, ry, uz, s, a.lH[2], o, a.lICE[2], ic, a.lIH[2], is);
    }
}

template <typename T, typename AT>
__global__ void k_adj_E(const AdjArgs<T, AT> a) {
    CEV_CELL_INDEX();
    const int im = (i == 0) ? a.Nx - 1 : i - 1;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const int km = (k == 0) ? a.Nz - 1 : k - 1;
    const int64_t o_im = im * plane + (int64_t)j * a.Nz + k;
    const int64_t o_jm = i * plane + (int64_t)jm * a.Nz + k;
    const int64_t o_km = i * plane + (int64_t)j * a.Nz + km;
    const AT inv = a.inv_dL;
    // curl_H (backward differences, derivatives.py:24-30) of gC2
    AT lE[3];
    lE[0] = ((AT)a.gC2[2][o] - (AT)a.gC2[2][o_jm]) * inv - ((AT)a.gC2[1][o] - (AT)a.gC2[1][o_km]) * inv;
    lE[1] = ((AT)a.gC2[0][o] - (AT)a.gC2[0][o_km]) * inv - ((AT)a.gC2[2][o] - (AT)a.gC2[2][o_im]) * inv;
    lE[2] = ((AT)a.gC2[1][o] - (AT)a.gC2[1][o_im]) * inv - ((AT)a.gC2[0][o] - (AT)a.gC2[0][o_jm]) * inv;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a.lD[c][o] = (T)((AT)a.lD[c][o] + (AT)a.mE[c][o] * lE[c]);
        if (a.G[c]) a.G[c][o] += (double)lE[c] * (double)a.Dprev[c][o];
    }
}

// Seeds of a probe-series objective: for step n and probe p with cotangent g = gbar[n, p],
//   E-probe: lD += mE g w ; G_mE += g w D_n      D-probe: lD += g w      H-probe: lH += g w
// One thread per point (atomics: several probes may share a cell).
template <typename T, typename AT>
__global__ void k_adj_seed(const AdjArgs<T, AT> a, ProbeTable pr, const int32_t* __restrict__ slot_owner,
                           const double* __restrict__ gbar_row, const T* D0, const T* D1, const T* D2) {
    const int slot = blockIdx.x;
    const int field = pr.slot_field[slot];
    const int c = field % 3;
    const int64_t n = pr.slot_n[slot], wb = pr.slot_wbegin[slot], ib = pr.slot_ibegin[slot], c0 = pr.slot_cell0[slot];
    const double g = gbar_row[slot_owner[slot]];
    if (g == 0.0) return;
    const T* Dn = c == 0 ? D0 : (c == 1 ? D1 : D2);
    for (int64_t q = threadIdx.x; q < n; q += blockDim.x) {
        const int64_t cell = (ib < 0) ? (c0 + q) : pr.idx[ib + q];
        const double gw = g * pr.weight[wb + q];
        if (field < 3) {
            atomicAdd(&a.lD[c][cell], (T)((double)a.mE[c][cell] * gw));
            if (a.G[c]) atomicAdd(&a.G[c][cell], gw * (double)Dn[cell]);
        } else if (field < 6) {
            atomicAdd(&a.lD[c][cell], (T)gw);
        } else {
            atomicAdd(&a.lH[c][cell], (T)gw);
        }
    }
}

}  // namespace cev
