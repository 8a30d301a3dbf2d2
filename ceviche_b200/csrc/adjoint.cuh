// Reverse-mode (VJP) kernels: the time-reversed adjoint FDTD step.
//
// The reference differentiates ceviche/fdtd.py:74-144 by taping every numpy op with HIPS autograd
// (ceviche/jacobians.py:29-35).  Here the transposed step is written out.  With periodic
// boundaries curl_E^T = curl_H and curl_H^T = curl_E (forward difference transposed = minus the
// backward difference), so the adjoint sweeps reuse the two stencils with their roles swapped.
// For step n, given the cotangents (lH, lD, lICE, lIH, lICH, lID) of the state after step n:
//
//   (D, cell-local)      gD = lD ; gICH = lICH + m3D gD ; gID = lID + m4D gD
//                        gC = m2D gD + gICH ; lD <- m1D gD + gID ; lICH <- gICH ; lID <- gID
//   (H, stencil)         gH = lH + curl_E(gC) ; gICE = lICE + m3H gH ; gIH = lIH + m4H gH
//                        gC2 = m2H gH + gICE ; lH <- m1H gH + gIH ; lICE <- gICE ; lIH <- gIH
//   (E, stencil)         lE = curl_H(gC2) ; lD += mE lE ; G_mE += lE * D_{n-1}
//
// after which (lH, lD, ...) are the cotangents of the state after step n-1 and G_mE has gained
// step n's contribution to dL/d(1/eps).  eps_r enters the step only through mE = 1/eps_yee
// (fdtd.py:67, 314-316), so G_mE is the whole gradient; the chain to eps_r is done by the host.
//
// Two kernels per step, one scratch vector field (round 1: three kernels, two scratch fields, 42 words per cell):
//   k_adj_H   evaluates gC ON THE FLY at the cell and its +1 neighbours from the OLD lD / lICH (off the PML it is
//             just C0 dt lD), applies the H part, writes lH and gC2: lD, lH in; lH, gC2 out = 12 words per cell;
//   k_adj_ED  takes lE from gC2, applies the DEFERRED cell-local D part (lD, lICH, lID) and adds mE lE:
//             gC2, lD, mE in; lD out = 12 words, plus D_{n-1} in and the fp64 G_mE read-modify-write INSIDE the
//             design box only (g_box): gradients are rarely wanted outside the region being optimised.
// 24 words per cell outside the box, 33 inside (fp64; SURVEY 8(d) counts 27 with G everywhere and no scratch).
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
    T* gC[3];        // cev_fdtd_adjoint_run's tensor-map kernels only (adjoint_v5.cuh): cotangent of curl_H(H_n), carried
    T* gC2[3];       // scratch: cotangent of curl_E(E_{n-1})
    double* G[3];    // dL/d(mE), accumulated in fp64 whatever the storage type
    int gb[6];       // design box (internal axes: x0, x1, y0, y1, z0, z1): G is only accumulated inside
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
                                            T* lIcurl, int64_t icurl, T* lIself, int64_t iself, AT extra = AT(0)) {
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
    l[o] = (T)(m1 * g + gIs + extra);
    return m2 * g + gIc;
}

#define CEV_CELL_INDEX()                                             \
    const int k = blockIdx.x * blockDim.x + threadIdx.x;             \
    const int j = blockIdx.y * blockDim.y + threadIdx.y;             \
    const int i = blockIdx.z;                                        \
    if (k >= a.Nz || j >= a.Ny) return;                              \
    const int64_t plane = (int64_t)a.Ny * a.Nz;                      \
    const int64_t o = i * plane + (int64_t)j * a.Nz + k;

// gC of component c at cell (i, j, k) = flat offset o, from the OLD cotangents (nothing is written): off the D-side
// PML it is C0 dt lD exactly (m2 = s r_a r_b with r = 1, no integral).  mx / my / mz: the cell's compact indices.
template <typename T, typename AT>
__device__ __forceinline__ AT adj_gC(const AdjArgs<T, AT>& a, int c, int i, int j, int k, int mx, int my, int mz, int64_t o) {
    const AT s = a.cdt;
    const AT g = (AT)a.lD[c][o];
    if ((mx & my & mz) < 0) return s * g;                // all three are -1: not in the PML of any axis
    AT ra, rb, uc;
    int64_t ic;
    if (c == 0) {
        ra = a.rD[1][j]; rb = a.rD[2][k]; uc = a.uD[0][i];
        ic = (mx >= 0) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
    } else if (c == 1) {
        ra = a.rD[0][i]; rb = a.rD[2][k]; uc = a.uD[1][j];
        ic = (my >= 0) ? ((int64_t)i * a.nD[1] + my) * a.Nz + k : -1;
    } else {
        ra = a.rD[0][i]; rb = a.rD[1][j]; uc = a.uD[2][k];
        ic = (mz >= 0) ? ((int64_t)i * a.Ny + j) * a.nD[2] + mz : -1;
    }
    const AT rr = mul_rn(ra, rb);
    AT v = mul_rn(s, rr) * g;
    if (ic >= 0) v += (AT)a.lICH[c][ic] + mul_rn(mul_rn(s, uc + uc), rr) * g;
    return v;
}

template <typename T, typename AT>
__global__ void __launch_bounds__(256) k_adj_H(const AdjArgs<T, AT> a) {
    CEV_CELL_INDEX();
    const int ip = (i + 1 == a.Nx) ? 0 : i + 1;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const int kp = (k + 1 == a.Nz) ? 0 : k + 1;
    const int64_t o_ip = ip * plane + (int64_t)j * a.Nz + k;
    const int64_t o_jp = i * plane + (int64_t)jp * a.Nz + k;
    const int64_t o_kp = i * plane + (int64_t)j * a.Nz + kp;
    const AT inv = a.inv_dL;
    const int dx = a.mapD[0][i], dy = a.mapD[1][j], dz = a.mapD[2][k];
    const int dxp = a.mapD[0][ip], dyp = a.mapD[1][jp], dzp = a.mapD[2][kp];
    // curl_E (forward differences, derivatives.py:16-22) of gC
    const AT gx = adj_gC<T, AT>(a, 0, i, j, k, dx, dy, dz, o);
    const AT gy = adj_gC<T, AT>(a, 1, i, j, k, dx, dy, dz, o);
    const AT gz = adj_gC<T, AT>(a, 2, i, j, k, dx, dy, dz, o);
    const AT gz_jp = adj_gC<T, AT>(a, 2, i, jp, k, dx, dyp, dz, o_jp), gx_jp = adj_gC<T, AT>(a, 0, i, jp, k, dx, dyp, dz, o_jp);
    const AT gy_kp = adj_gC<T, AT>(a, 1, i, j, kp, dx, dy, dzp, o_kp), gx_kp = adj_gC<T, AT>(a, 0, i, j, kp, dx, dy, dzp, o_kp);
    const AT gz_ip = adj_gC<T, AT>(a, 2, ip, j, k, dxp, dy, dz, o_ip), gy_ip = adj_gC<T, AT>(a, 1, ip, j, k, dxp, dy, dz, o_ip);
    const AT cx = (gz_jp - gz) * inv - (gy_kp - gy) * inv;
    const AT cy = (gx_kp - gx) * inv - (gz_ip - gz) * inv;
    const AT cz = (gy_ip - gy) * inv - (gx_jp - gx) * inv;
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
__global__ void __launch_bounds__(256) k_adj_ED(const AdjArgs<T, AT> a) {
    CEV_CELL_INDEX();
    const int im = (i == 0) ? a.Nx - 1 : i - 1;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const int km = (k == 0) ? a.Nz - 1 : k - 1;
    const int64_t o_im = im * plane + (int64_t)j * a.Nz + k;
    const int64_t o_jm = i * plane + (int64_t)jm * a.Nz + k;
    const int64_t o_km = i * plane + (int64_t)j * a.Nz + km;
    const AT inv = a.inv_dL;
    // curl_H (backward differences, derivatives.py:24-30) of gC2
    const AT c0 = (AT)a.gC2[0][o], c1 = (AT)a.gC2[1][o], c2 = (AT)a.gC2[2][o];
    AT lE[3];
    lE[0] = (c2 - (AT)a.gC2[2][o_jm]) * inv - (c1 - (AT)a.gC2[1][o_km]) * inv;
    lE[1] = (c0 - (AT)a.gC2[0][o_km]) * inv - (c2 - (AT)a.gC2[2][o_im]) * inv;
    lE[2] = (c1 - (AT)a.gC2[1][o_im]) * inv - (c0 - (AT)a.gC2[0][o_jm]) * inv;
    // the deferred cell-local D part, then lD += mE lE
    const int mx = a.mapD[0][i], my = a.mapD[1][j], mz = a.mapD[2][k];
    const bool in_box = i >= a.gb[0] && i < a.gb[1] && j >= a.gb[2] && j < a.gb[3] && k >= a.gb[4] && k < a.gb[5];
    if ((mx & my & mz) < 0) {                                // off the PML: m1 = 1, no integrals
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a.lD[c][o] = (T)((AT)a.lD[c][o] + (AT)a.mE[c][o] * lE[c]);
            if (in_box && a.G[c]) a.G[c][o] += (double)lE[c] * (double)a.Dprev[c][o];
        }
        return;
    }
    const AT ux = a.uD[0][i], uy = a.uD[1][j], uz = a.uD[2][k];
    const AT rx = a.rD[0][i], ry = a.rD[1][j], rz = a.rD[2][k];
    const AT s = a.cdt;
    {
        const int64_t ic = (mx >= 0) ? ((int64_t)mx * a.Ny + j) * a.Nz + k : -1;
        const int64_t is = (my >= 0 && mz >= 0) ? ((int64_t)i * a.nD[1] + my) * a.nD[2] + mz : -1;
        adj_component<T, AT>((AT)a.lD[0][o], uy, ry, uz, rz, ux, s, a.lD[0], o, a.lICH[0], ic, a.lID[0], is, (AT)a.mE[0][o] * lE[0]);
    }
    {
        const int64_t ic = (my >= 0) ? ((int64_t)i * a.nD[1] + my) * a.Nz + k : -1;
        const int64_t is = (mx >= 0 && mz >= 0) ? ((int64_t)mx * a.Ny + j) * a.nD[2] + mz : -1;
        adj_component<T, AT>((AT)a.lD[1][o], ux, rx, uz, rz, uy, s, a.lD[1], o, a.lICH[1], ic, a.lID[1], is, (AT)a.mE[1][o] * lE[1]);
    }
    {
        const int64_t ic = (mz >= 0) ? ((int64_t)i * a.Ny + j) * a.nD[2] + mz : -1;
        const int64_t is = (mx >= 0 && my >= 0) ? ((int64_t)mx * a.nD[1] + my) * a.Nz + k : -1;
        adj_component<T, AT>((AT)a.lD[2][o], ux, rx, uy, ry, uz, s, a.lD[2], o, a.lICH[2], ic, a.lID[2], is, (AT)a.mE[2][o] * lE[2]);
    }
    if (in_box)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (a.G[c]) a.G[c][o] += (double)lE[c] * (double)a.Dprev[c][o];
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
            if (a.G[c]) {
                const int64_t pl = (int64_t)a.Ny * a.Nz;
                const int ci = (int)(cell / pl), cj = (int)((cell % pl) / a.Nz), ck = (int)(cell % a.Nz);
                if (ci >= a.gb[0] && ci < a.gb[1] && cj >= a.gb[2] && cj < a.gb[3] && ck >= a.gb[4] && ck < a.gb[5])
                    atomicAdd(&a.G[c][cell], gw * (double)Dn[cell]);
            }
        } else if (field < 6) {
            atomicAdd(&a.lD[c][cell], (T)gw);
        } else {
            atomicAdd(&a.lH[c][cell], (T)gw);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-map path of cev_fdtd_adjoint_run (adjoint_v5.cuh), start of a segment: the cell-local D part of the segment's last step applied to the TRUE cotangent lD
// (-> lDp in place, gC, lICH, lID).  One thread per cell; once per checkpoint segment.
template <typename T, typename AT>
__global__ void __launch_bounds__(256) k_adj_Dlocal(const AdjArgs<T, AT> a) {
    CEV_CELL_INDEX();
    const int mx = a.mapD[0][i], my = a.mapD[1][j], mz = a.mapD[2][k];
    const AT s = a.cdt;
    if ((mx & my & mz) < 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) a.gC[c][o] = (T)(s * (AT)a.lD[c][o]);
        return;
    }
    const AT ux = a.uD[0][i], uy = a.uD[1][j], uz = a.uD[2][k];
    const AT rx = a.rD[0][i], ry = a.rD[1][j], rz = a.rD[2][k];
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

// Probe-series seeds of step k in the eager form: a seed v on the true lD of cell o (E-probe: v = mE g w, D-probe:
// v = g w) is, after the D part of step k, (m1 + m4) v on lDp, (m2 + m3) v on gC, m3 v on lICH, m4 v on lID (the D
// part is linear).  H-probe seeds go to lH unchanged.  One thread per point, atomics (probes may share cells).
template <typename T, typename AT>
__global__ void k_adj_seed_eager(const AdjArgs<T, AT> a, ProbeTable pr, const int32_t* __restrict__ slot_owner,
                                 const double* __restrict__ gbar_row, const T* D0, const T* D1, const T* D2, int boxed) {
    const int slot = blockIdx.x;
    const int field = pr.slot_field[slot];
    const int c = field % 3;
    const int64_t n = pr.slot_n[slot], wb = pr.slot_wbegin[slot], ib = pr.slot_ibegin[slot], c0 = pr.slot_cell0[slot];
    const double gb = gbar_row[slot_owner[slot]];
    if (gb == 0.0) return;
    const T* Dn = c == 0 ? D0 : (c == 1 ? D1 : D2);
    const int64_t pl = (int64_t)a.Ny * a.Nz;
    for (int64_t q = threadIdx.x; q < n; q += blockDim.x) {
        const int64_t cell = (ib < 0) ? (c0 + q) : pr.idx[ib + q];
        const double gw = gb * pr.weight[wb + q];
        if (field >= 6) {
            atomicAdd(&a.lH[c][cell], (T)gw);
            continue;
        }
        const int i = (int)(cell / pl), j = (int)((cell % pl) / a.Nz), k = (int)(cell % a.Nz);
        double v = gw;
        if (field < 3) {
            v = (double)a.mE[c][cell] * gw;
            if (a.G[c] && i >= a.gb[0] && i < a.gb[1] && j >= a.gb[2] && j < a.gb[3] && k >= a.gb[4] && k < a.gb[5]) {
                // (boxed: D_k is the forward run's record of the design box only, C-order)
                const int64_t dn = boxed ? ((int64_t)(i - a.gb[0]) * (a.gb[3] - a.gb[2]) + (j - a.gb[2])) * (a.gb[5] - a.gb[4]) + (k - a.gb[4]) : cell;
                atomicAdd(&a.G[c][cell], gw * (double)Dn[dn]);
            }
        }
        const int m[3] = {a.mapD[0][i], a.mapD[1][j], a.mapD[2][k]};
        const int ijk[3] = {i, j, k};
        const int A = (c + 1) % 3, B = (c + 2) % 3;
        const double ua = (double)a.uD[A][ijk[A]], ub = (double)a.uD[B][ijk[B]], uc = (double)a.uD[c][ijk[c]];
        const double rr = (double)a.rD[A][ijk[A]] * (double)a.rD[B][ijk[B]];
        const double s = (double)a.cdt;
        const double m1 = 2 * rr - 1, m2 = s * rr, m3 = s * 2 * uc * rr, m4 = -4 * ua * ub * rr;
        double dl = m1 * v, dg = m2 * v;
        if (m[c] >= 0) {
            int64_t ic;
            if (c == 0) ic = ((int64_t)m[0] * a.Ny + j) * a.Nz + k;
            else if (c == 1) ic = ((int64_t)i * a.nD[1] + m[1]) * a.Nz + k;
            else ic = ((int64_t)i * a.Ny + j) * a.nD[2] + m[2];
            atomicAdd(&a.lICH[c][ic], (T)(m3 * v));
            dg += m3 * v;
        }
        if (m[A] >= 0 && m[B] >= 0) {
            int64_t is;
            if (c == 0) is = ((int64_t)i * a.nD[1] + m[1]) * a.nD[2] + m[2];
            else if (c == 1) is = ((int64_t)m[0] * a.Ny + j) * a.nD[2] + m[2];
            else is = ((int64_t)m[0] * a.nD[1] + m[1]) * a.Nz + k;
            atomicAdd(&a.lID[c][is], (T)(m4 * v));
            dl += m4 * v;
        }
        atomicAdd(&a.lD[c][cell], (T)dl);
        atomicAdd(&a.gC[c][cell], (T)dg);
    }
}

// ---- the transposed step in three parts with STORED stencil inputs and x-halo planes: the form the x-slab
// decomposition drives (one process per GPU, slab.py): k_adj_Dlocal (above; cell-local D part: lD -> m1 lD + gID,
// gC), then the two stencil parts below.  A slab's +x / -x neighbour planes of gC / gC2 (components y, z: the only
// ones differenced along x) come from the neighbouring rank; NULL = periodic wrap inside the array.
template <typename T, typename AT>
__global__ void __launch_bounds__(256) k_adj_H_stored(const AdjArgs<T, AT> a, const T* gCy_hi, const T* gCz_hi) {
    CEV_CELL_INDEX();
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const int kp = (k + 1 == a.Nz) ? 0 : k + 1;
    const int64_t row = (int64_t)j * a.Nz + k;
    const int64_t o_jp = i * plane + (int64_t)jp * a.Nz + k;
    const int64_t o_kp = i * plane + (int64_t)j * a.Nz + kp;
    const AT inv = a.inv_dL;
    const AT gx = (AT)a.gC[0][o], gy = (AT)a.gC[1][o], gz = (AT)a.gC[2][o];
    AT gy_ip, gz_ip;
    if (i + 1 < a.Nx) {
        gy_ip = (AT)a.gC[1][o + plane];
        gz_ip = (AT)a.gC[2][o + plane];
    } else {
        gy_ip = gCy_hi ? (AT)gCy_hi[row] : (AT)a.gC[1][row];
        gz_ip = gCz_hi ? (AT)gCz_hi[row] : (AT)a.gC[2][row];
    }
    // curl_E (forward differences, derivatives.py:16-22) of gC
    const AT cx = ((AT)a.gC[2][o_jp] - gz) * inv - ((AT)a.gC[1][o_kp] - gy) * inv;
    const AT cy = ((AT)a.gC[0][o_kp] - gx) * inv - (gz_ip - gz) * inv;
    const AT cz = (gy_ip - gy) * inv - ((AT)a.gC[0][o_jp] - gx) * inv;
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

// lE = curl_H(gC2) (backward differences, derivatives.py:24-30); lD += mE lE (lD already holds the cell-local D
// part); G_mE += lE D_{n-1} inside the design box
template <typename T, typename AT>
__global__ void __launch_bounds__(256) k_adj_E_stored(const AdjArgs<T, AT> a, const T* gC2y_lo, const T* gC2z_lo) {
    CEV_CELL_INDEX();
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const int km = (k == 0) ? a.Nz - 1 : k - 1;
    const int64_t row = (int64_t)j * a.Nz + k;
    const int64_t o_jm = i * plane + (int64_t)jm * a.Nz + k;
    const int64_t o_km = i * plane + (int64_t)j * a.Nz + km;
    const AT inv = a.inv_dL;
    const AT c0 = (AT)a.gC2[0][o], c1 = (AT)a.gC2[1][o], c2 = (AT)a.gC2[2][o];
    AT c1_im, c2_im;
    if (i > 0) {
        c1_im = (AT)a.gC2[1][o - plane];
        c2_im = (AT)a.gC2[2][o - plane];
    } else {
        const int64_t last = (int64_t)(a.Nx - 1) * plane + row;
        c1_im = gC2y_lo ? (AT)gC2y_lo[row] : (AT)a.gC2[1][last];
        c2_im = gC2z_lo ? (AT)gC2z_lo[row] : (AT)a.gC2[2][last];
    }
    AT lE[3];
    lE[0] = (c2 - (AT)a.gC2[2][o_jm]) * inv - (c1 - (AT)a.gC2[1][o_km]) * inv;
    lE[1] = (c0 - (AT)a.gC2[0][o_km]) * inv - (c2 - c2_im) * inv;
    lE[2] = (c1 - c1_im) * inv - (c0 - (AT)a.gC2[0][o_jm]) * inv;
    const bool in_box = i >= a.gb[0] && i < a.gb[1] && j >= a.gb[2] && j < a.gb[3] && k >= a.gb[4] && k < a.gb[5];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a.lD[c][o] = (T)((AT)a.lD[c][o] + (AT)a.mE[c][o] * lE[c]);
        if (in_box && a.G[c]) a.G[c][o] += (double)lE[c] * (double)a.Dprev[c][o];
    }
}

}  // namespace cev
