// Tuned half-step kernels for sm_100a: x-marching, z-vectorised.
//
// Thread (lane, ty) of a CTA owns V consecutive z-cells (one 16-byte vector: float4 / double2)
// of row j and marches over a chunk of x-planes.  The x-neighbour plane (i+1 for the H half-step,
// i-1 for the D half-step) is carried in registers from one iteration to the next, so every
// plane of D / 1/eps / H is fetched from L2 once per chunk; the y-neighbour row is a second,
// L1-resident vector load; the z-neighbour comes from the adjacent lane by warp shuffle (one
// scalar load on the last / first lane of a warp).  All 9 (H) / 8 (D) vector loads of an
// iteration are independent, which is what keeps enough bytes in flight to cover HBM latency.
// PML: coefficients come from the per-axis tables; the cell-invariant parts are hoisted out of
// the marching loop and the general formula only runs on x-PML planes (a CTA-uniform branch).
// Results are bit-identical to step_v1.cuh (same rounding sequence), which the tests assert.
#pragma once
#include "common.cuh"

namespace cev {

template <typename T, int V>
struct alignas(sizeof(T) * V) Vec {
    T v[V];
};

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ldv(const T* p) {
    return *reinterpret_cast<const Vec<T, V>*>(p);
}
template <typename T, int V>
__device__ __forceinline__ void stv(T* p, const Vec<T, V>& x) {
    *reinterpret_cast<Vec<T, V>*>(p) = x;
}

constexpr int V2_BY = 4;        // rows per CTA (one warp per row)
constexpr int V2_MIN_CTAS = 4;  // register cap: 65536 / (4 * 128) = 128 per thread

template <typename T, typename AT, int V>
__global__ void __launch_bounds__(32 * V2_BY, V2_MIN_CTAS) k_step_H_v2(const StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    const int lane = threadIdx.x;
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int xc = rest / a.nty;
    const int j = ty * V2_BY + threadIdx.y;
    if (j >= a.Ny) return;                       // warp-uniform
    const int k0raw = (tz * 32 + lane) * V;
    const bool active = k0raw < a.Nz;
    const int k0 = active ? k0raw : 0;           // inactive lanes shadow cell 0 (loads only)
    const int xs = a.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, a.x1);

    const int plane = a.Ny * a.Nz;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const bool z_edge = (lane == 31) || (k0 + V >= a.Nz);   // +1 neighbour not in lane+1
    const int kp = (k0 + V >= a.Nz) ? 0 : k0 + V;
    const int orow = j * a.Nz + k0;
    const int orow_jp = jp * a.Nz + k0;
    const int okp = j * a.Nz + kp;

    // cell-invariant PML data
    const AT uy = a.uH[1][j], ry = a.rH[1][j];
    const int my = a.mapH[1][j];
    AT uz[V], rz[V];
    int mz[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        uz[e] = a.uH[2][k0 + e];
        rz[e] = a.rH[2][k0 + e];
        mz[e] = a.mapH[2][k0 + e];
    }
    const AT s = -a.cdt;
    const AT zero = AT(0), one = AT(1);
    // coefficients off the x-PML (ux = 0): hoisted
    // m1, m2 off the x-PML (ux = 0, rx = 1): hoisted out of the marching loop
    AT m1x[V], m2x[V], m1y0[V], m2y0[V], m1z0, m2z0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        coef12<AT>(uy, ry, uz[e], rz[e], s, m1x[e], m2x[e]);
        coef12<AT>(zero, one, uz[e], rz[e], s, m1y0[e], m2y0[e]);
    }
    coef12<AT>(zero, one, uy, ry, s, m1z0, m2z0);

    // E = mE*D of the current plane (own cells)
    AT Ecur[3][V];
    {
        const int o = xs * plane + orow;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Vec<T, V> d = ldv<T, V>(a.Din[c] + o), m = ldv<T, V>(a.mE[c] + o);
#pragma unroll
            for (int e = 0; e < V; ++e) Ecur[c][e] = mul_rn((AT)m.v[e], (AT)d.v[e]);
        }
    }

    for (int i = xs; i < xe; ++i) {
        const int pbase = i * plane;
        const bool last = (i + 1 == a.Nx);
        // ---- issue every load of this iteration up front
        Vec<T, V> dn[3], mn[3], h[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T* Dn = last ? a.Dhi[c] : a.Din[c] + pbase + plane;
            const T* Mn = last ? a.mEhi[c] : a.mE[c] + pbase + plane;
            dn[c] = ldv<T, V>(Dn + orow);
            mn[c] = ldv<T, V>(Mn + orow);
            h[c] = ldv<T, V>(a.Hin[c] + pbase + orow);
        }
        const Vec<T, V> dxj = ldv<T, V>(a.Din[0] + pbase + orow_jp), mxj = ldv<T, V>(a.mE[0] + pbase + orow_jp);
        const Vec<T, V> dzj = ldv<T, V>(a.Din[2] + pbase + orow_jp), mzj = ldv<T, V>(a.mE[2] + pbase + orow_jp);
        AT ex_kp = __shfl_down_sync(0xffffffffu, Ecur[0][0], 1);
        AT ey_kp = __shfl_down_sync(0xffffffffu, Ecur[1][0], 1);
        if (z_edge) {
            ex_kp = mul_rn((AT)a.mE[0][pbase + okp], (AT)a.Din[0][pbase + okp]);
            ey_kp = mul_rn((AT)a.mE[1][pbase + okp], (AT)a.Din[1][pbase + okp]);
        }
        const AT ux = a.uH[0][i], rx = a.rH[0][i];
        const int mx = a.mapH[0][i];

        AT Enext[3][V];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int e = 0; e < V; ++e) Enext[c][e] = mul_rn((AT)mn[c].v[e], (AT)dn[c].v[e]);

        Vec<T, V> out[3];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Ex = Ecur[0][e], Ey = Ecur[1][e], Ez = Ecur[2][e];
            const AT Ex_jp = mul_rn((AT)mxj.v[e], (AT)dxj.v[e]);
            const AT Ez_jp = mul_rn((AT)mzj.v[e], (AT)dzj.v[e]);
            const AT Ex_kp = (e + 1 < V) ? Ecur[0][(e + 1) % V] : ex_kp;
            const AT Ey_kp = (e + 1 < V) ? Ecur[1][(e + 1) % V] : ey_kp;
            const AT CEx = curl2<AT>(Ez_jp, Ez, Ey_kp, Ey, a.inv_dL);
            const AT CEy = curl2<AT>(Ex_kp, Ex, Enext[2][e], Ez, a.inv_dL);
            const AT CEz = curl2<AT>(Enext[1][e], Ey, Ex_jp, Ex, a.inv_dL);
            const int k = k0 + e;
            AT m1y = m1y0[e], m2y = m2y0[e], m1z = m1z0, m2z = m2z0;
            int ic0 = -1, is0 = -1, ic1 = -1, is1 = -1, ic2 = -1, is2 = -1;
            if (mx >= 0) {   // x-PML plane: CTA-uniform
                coef12<AT>(ux, rx, uz[e], rz[e], s, m1y, m2y);
                coef12<AT>(ux, rx, uy, ry, s, m1z, m2z);
                ic0 = (mx * a.Ny + j) * a.Nz + k;
                if (mz[e] >= 0) is1 = (mx * a.Ny + j) * a.nH[2] + mz[e];
                if (my >= 0) is2 = (mx * a.nH[1] + my) * a.Nz + k;
            }
            if (my >= 0) {
                ic1 = (i * a.nH[1] + my) * a.Nz + k;
                if (mz[e] >= 0) is0 = (i * a.nH[1] + my) * a.nH[2] + mz[e];
            }
            if (mz[e] >= 0) ic2 = (i * a.Ny + j) * a.nH[2] + mz[e];
            if (active) {
                out[0].v[e] = (T)update_cell<T, AT>((AT)h[0].v[e], CEx, m1x[e], m2x[e], uy, ry, uz[e], rz[e], ux, s,
                                                    a.ICE[0], ic0, a.IH[0], is0);
                out[1].v[e] = (T)update_cell<T, AT>((AT)h[1].v[e], CEy, m1y, m2y, ux, rx, uz[e], rz[e], uy, s,
                                                    a.ICE[1], ic1, a.IH[1], is1);
                out[2].v[e] = (T)update_cell<T, AT>((AT)h[2].v[e], CEz, m1z, m2z, ux, rx, uy, ry, uz[e], s,
                                                    a.ICE[2], ic2, a.IH[2], is2);
            }
        }
        if (active) {
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Hout[c] + pbase + orow, out[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int e = 0; e < V; ++e) Ecur[c][e] = Enext[c][e];
    }
}

// EXTRAS = dense J input and/or E output (the per-step forward() API); the fused run() path
// instantiates EXTRAS = false and carries neither.
template <typename T, typename AT, int V, bool EXTRAS>
__global__ void __launch_bounds__(32 * V2_BY, V2_MIN_CTAS) k_step_D_v2(const StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    const int lane = threadIdx.x;
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int xc = rest / a.nty;
    const int j = ty * V2_BY + threadIdx.y;
    if (j >= a.Ny) return;
    const int k0raw = (tz * 32 + lane) * V;
    const bool active = k0raw < a.Nz;
    const int k0 = active ? k0raw : 0;
    const int xs = a.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, a.x1);

    const int plane = a.Ny * a.Nz;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const bool z_edge = (lane == 0) || (k0 == 0);
    const int km = (k0 == 0) ? a.Nz - 1 : k0 - 1;
    const int orow = j * a.Nz + k0;
    const int orow_jm = jm * a.Nz + k0;
    const int okm = j * a.Nz + km;

    const AT uy = a.uD[1][j], ry = a.rD[1][j];
    const int my = a.mapD[1][j];
    AT uz[V], rz[V];
    int mz[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        uz[e] = a.uD[2][k0 + e];
        rz[e] = a.rD[2][k0 + e];
        mz[e] = a.mapD[2][k0 + e];
    }
    const AT s = a.cdt;
    const AT zero = AT(0), one = AT(1);
    // m1, m2 off the x-PML (ux = 0, rx = 1): hoisted out of the marching loop
    AT m1x[V], m2x[V], m1y0[V], m2y0[V], m1z0, m2z0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        coef12<AT>(uy, ry, uz[e], rz[e], s, m1x[e], m2x[e]);
        coef12<AT>(zero, one, uz[e], rz[e], s, m1y0[e], m2y0[e]);
    }
    coef12<AT>(zero, one, uy, ry, s, m1z0, m2z0);

    // H of the previous plane (own cells), y and z components
    AT Hprev[2][V];
    {
        const T* P1 = (xs > 0) ? a.Hin[1] + (xs - 1) * plane : a.Hlo[1];
        const T* P2 = (xs > 0) ? a.Hin[2] + (xs - 1) * plane : a.Hlo[2];
        const Vec<T, V> p1 = ldv<T, V>(P1 + orow), p2 = ldv<T, V>(P2 + orow);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            Hprev[0][e] = (AT)p1.v[e];
            Hprev[1][e] = (AT)p2.v[e];
        }
    }

    for (int i = xs; i < xe; ++i) {
        const int pbase = i * plane;
        Vec<T, V> h[3], d[3], jv[3], mev[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            h[c] = ldv<T, V>(a.Hin[c] + pbase + orow);
            d[c] = ldv<T, V>(a.Din[c] + pbase + orow);
            if (EXTRAS && a.J[c]) jv[c] = ldv<T, V>(a.J[c] + pbase + orow);
            if (EXTRAS && a.Eout[c]) mev[c] = ldv<T, V>(a.mE[c] + pbase + orow);
        }
        const Vec<T, V> hxj = ldv<T, V>(a.Hin[0] + pbase + orow_jm), hzj = ldv<T, V>(a.Hin[2] + pbase + orow_jm);
        AT hx_km = __shfl_up_sync(0xffffffffu, (AT)h[0].v[V - 1], 1);
        AT hy_km = __shfl_up_sync(0xffffffffu, (AT)h[1].v[V - 1], 1);
        if (z_edge) {
            hx_km = (AT)a.Hin[0][pbase + okm];
            hy_km = (AT)a.Hin[1][pbase + okm];
        }
        const AT ux = a.uD[0][i], rx = a.rD[0][i];
        const int mx = a.mapD[0][i];

        Vec<T, V> out[3], eout[3];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Hx = (AT)h[0].v[e], Hy = (AT)h[1].v[e], Hz = (AT)h[2].v[e];
            const AT Hx_km = (e > 0) ? (AT)h[0].v[(e + V - 1) % V] : hx_km;
            const AT Hy_km = (e > 0) ? (AT)h[1].v[(e + V - 1) % V] : hy_km;
            const AT CHx = curl2<AT>(Hz, (AT)hzj.v[e], Hy, Hy_km, a.inv_dL);
            const AT CHy = curl2<AT>(Hx, Hx_km, Hz, Hprev[1][e], a.inv_dL);
            const AT CHz = curl2<AT>(Hy, Hprev[0][e], Hx, (AT)hxj.v[e], a.inv_dL);
            const int k = k0 + e;
            AT m1y = m1y0[e], m2y = m2y0[e], m1z = m1z0, m2z = m2z0;
            int ic0 = -1, is0 = -1, ic1 = -1, is1 = -1, ic2 = -1, is2 = -1;
            if (mx >= 0) {   // x-PML plane: CTA-uniform
                coef12<AT>(ux, rx, uz[e], rz[e], s, m1y, m2y);
                coef12<AT>(ux, rx, uy, ry, s, m1z, m2z);
                ic0 = (mx * a.Ny + j) * a.Nz + k;
                if (mz[e] >= 0) is1 = (mx * a.Ny + j) * a.nD[2] + mz[e];
                if (my >= 0) is2 = (mx * a.nD[1] + my) * a.Nz + k;
            }
            if (my >= 0) {
                ic1 = (i * a.nD[1] + my) * a.Nz + k;
                if (mz[e] >= 0) is0 = (i * a.nD[1] + my) * a.nD[2] + mz[e];
            }
            if (mz[e] >= 0) ic2 = (i * a.Ny + j) * a.nD[2] + mz[e];
            if (active) {
                AT dn[3];
                dn[0] = update_cell<T, AT>((AT)d[0].v[e], CHx, m1x[e], m2x[e], uy, ry, uz[e], rz[e], ux, s,
                                           a.ICH[0], ic0, a.ID[0], is0);
                dn[1] = update_cell<T, AT>((AT)d[1].v[e], CHy, m1y, m2y, ux, rx, uz[e], rz[e], uy, s,
                                           a.ICH[1], ic1, a.ID[1], is1);
                dn[2] = update_cell<T, AT>((AT)d[2].v[e], CHz, m1z, m2z, ux, rx, uy, ry, uz[e], s,
                                           a.ICH[2], ic2, a.ID[2], is2);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (EXTRAS && a.J[c]) {
                        const AT sc = a.Jwave[c] ? (AT)(*a.Jwave[c]) : a.Jscale[c];
                        dn[c] = add_rn(dn[c], mul_rn((AT)jv[c].v[e], sc));
                    }
                    out[c].v[e] = (T)dn[c];
                    if (EXTRAS && a.Eout[c]) eout[c].v[e] = (T)mul_rn((AT)mev[c].v[e], (AT)out[c].v[e]);
                }
            }
            Hprev[0][e] = Hy;
            Hprev[1][e] = Hz;
        }
        if (active) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                stv<T, V>(a.Dout[c] + pbase + orow, out[c]);
                if (EXTRAS && a.Eout[c]) stv<T, V>(a.Eout[c] + pbase + orow, eout[c]);
            }
        }
    }
}

}  // namespace cev
