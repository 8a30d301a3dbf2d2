// Tuned half-step kernels for sm_100a: x-marching, z-vectorised.
//
// Thread (lane, ty) of a CTA owns V consecutive z-cells (one 16-byte vector: float4 / double2)
// of row j and marches over a chunk of x-planes.  The x-neighbour plane (i+1 for the H half-step,
// i-1 for the D half-step) is carried in registers from one iteration to the next, so every
// plane of D / 1/eps / H is fetched from L2 once per chunk; the y-neighbour row is a second,
// L1/L2-resident vector load; the z-neighbour comes from the adjacent lane by warp shuffle (one
// scalar load on the last / first lane of a warp).  All vector loads of an iteration are
// independent and issued up front.
//
// Register diet (occupancy is what buys HBM bandwidth here -- scripts/microbench/streams.cu shows
// the same 12-stream marching pattern reaching ~7 TB/s when the kernel is lean): nothing PML-related
// is carried across iterations.  A thread knows three flags (its row is in the y-PML, any of its
// cells is in the z-PML, the current plane is in the x-PML); when none is set -- the bulk of the
// grid -- the update is the vacuum one (m1 = 1, m2 = -+C0 dt); otherwise the general path re-reads
// the tiny per-axis tables (L1-resident) for that cell.
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

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

constexpr int V2_BY = 4;        // warps per CTA; a warp covers LZ lanes along z times 32/LZ rows
// register caps = 65536 / (MIN_CTAS * 128) per thread; tuned on B200 (scripts/tune.py): the H kernel
// carries more state and prefers fewer, fatter CTAs; the D kernel prefers occupancy
#ifndef V2_H_MIN_CTAS
#define V2_H_MIN_CTAS 4
#endif
#ifndef V2_D_MIN_CTAS
#define V2_D_MIN_CTAS 8
#endif

// General (PML) update of the three components of one cell.  IS_H selects the H or the D tables at
// compile time (no address of the parameter struct is taken: it stays in the constant bank).
template <typename T, typename AT, bool IS_H>
__device__ __forceinline__ void pml_cell(const StepArgs<T, AT>& a, int i, int j, int k, AT s, const AT* old,
                                         const AT* curl, AT* out) {
#define TAB(name, ax) (IS_H ? a.name##H[ax] : a.name##D[ax])
    const AT ux = TAB(u, 0)[i], uy = TAB(u, 1)[j], uz = TAB(u, 2)[k];
    const AT rx = TAB(r, 0)[i], ry = TAB(r, 1)[j], rz = TAB(r, 2)[k];
    const int mx = TAB(map, 0)[i], my = TAB(map, 1)[j], mz = TAB(map, 2)[k];
    const int n1 = IS_H ? a.nH[1] : a.nD[1], n2 = IS_H ? a.nH[2] : a.nD[2];
#undef TAB
    T* const Ic0 = IS_H ? a.ICE[0] : a.ICH[0];
    T* const Ic1 = IS_H ? a.ICE[1] : a.ICH[1];
    T* const Ic2 = IS_H ? a.ICE[2] : a.ICH[2];
    T* const Is0 = IS_H ? a.IH[0] : a.ID[0];
    T* const Is1 = IS_H ? a.IH[1] : a.ID[1];
    T* const Is2 = IS_H ? a.IH[2] : a.ID[2];
    // x: (a,b) = (y,z), own x.   Icurl_x (nCx,Ny,Nz), Iself_x (Nx,nCy,nCz)
    out[0] = update_component<T, AT>(old[0], curl[0], uy, ry, uz, rz, ux, s, Ic0,
                                     mx >= 0 ? (int64_t)((mx * a.Ny + j) * a.Nz + k) : -1, Is0,
                                     (my >= 0 && mz >= 0) ? (int64_t)((i * n1 + my) * n2 + mz) : -1);
    // y: (a,b) = (x,z), own y.   Icurl_y (Nx,nCy,Nz), Iself_y (nCx,Ny,nCz)
    out[1] = update_component<T, AT>(old[1], curl[1], ux, rx, uz, rz, uy, s, Ic1,
                                     my >= 0 ? (int64_t)((i * n1 + my) * a.Nz + k) : -1, Is1,
                                     (mx >= 0 && mz >= 0) ? (int64_t)((mx * a.Ny + j) * n2 + mz) : -1);
    // z: (a,b) = (x,y), own z.   Icurl_z (Nx,Ny,nCz), Iself_z (nCx,nCy,Nz)
    out[2] = update_component<T, AT>(old[2], curl[2], ux, rx, uy, ry, uz, s, Ic2,
                                     mz >= 0 ? (int64_t)((i * a.Ny + j) * n2 + mz) : -1, Is2,
                                     (mx >= 0 && my >= 0) ? (int64_t)((mx * n1 + my) * a.Nz + k) : -1);
}

template <typename T, typename AT, int V, int LZ>
__global__ void __launch_bounds__(32 * V2_BY, V2_H_MIN_CTAS) k_step_H_v2(const StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    constexpr int RW = 32 / LZ;                  // rows per warp
    const int lane = threadIdx.x;
    const int lz = lane % LZ, ly = lane / LZ;
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int xc = rest / a.nty;
    const int jraw = (ty * V2_BY + threadIdx.y) * RW + ly;
    if (jraw - ly >= a.Ny) return;               // warp-uniform: the whole warp is below the grid
    const int k0raw = (tz * LZ + lz) * V;
    const bool active = k0raw < a.Nz && jraw < a.Ny;
    const int j = jraw < a.Ny ? jraw : a.Ny - 1; // inactive lanes shadow a valid cell (loads only)
    const int k0 = k0raw < a.Nz ? k0raw : 0;
    const int xs = a.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, a.x1);

    const int plane = a.Ny * a.Nz;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const bool z_edge = (lz == LZ - 1) || (k0 + V >= a.Nz);   // +1 neighbour not in lane+1
    const int kp = (k0 + V >= a.Nz) ? 0 : k0 + V;
    const int orow = j * a.Nz + k0;
    const int orow_jp = jp * a.Nz + k0;
    const int okp = j * a.Nz + kp;

    bool yz_pml = a.mapH[1][j] >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) yz_pml |= a.mapH[2][k0 + e] >= 0;
    const AT s = -a.cdt;
    const AT inv = a.inv_dL;

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
        // ---- every load of this iteration, up front
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
        if (a.pf_dist > 0 && (lz & 7) == 0 && i + a.pf_dist < xe) {   // one lane per 128-byte line
            const int po = (i + a.pf_dist) * plane + orow;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                prefetch_l2(a.Hin[c] + po);
                if (i + a.pf_dist + 1 < a.Nx) {
                    prefetch_l2(a.Din[c] + po + plane);
                    prefetch_l2(a.mE[c] + po + plane);
                }
            }
        }
        const bool pml = yz_pml || (a.mapH[0][i] >= 0);

        // ---- curls (consume the neighbour loads); E of the next plane becomes current
        AT CE[3][V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Ex = Ecur[0][e], Ey = Ecur[1][e], Ez = Ecur[2][e];
            const AT Ex_jp = mul_rn((AT)mxj.v[e], (AT)dxj.v[e]);
            const AT Ez_jp = mul_rn((AT)mzj.v[e], (AT)dzj.v[e]);
            const AT Ex_kp = (e + 1 < V) ? Ecur[0][(e + 1) % V] : ex_kp;
            const AT Ey_kp = (e + 1 < V) ? Ecur[1][(e + 1) % V] : ey_kp;
            const AT Ey_ip = mul_rn((AT)mn[1].v[e], (AT)dn[1].v[e]);
            const AT Ez_ip = mul_rn((AT)mn[2].v[e], (AT)dn[2].v[e]);
            CE[0][e] = curl2<AT>(Ez_jp, Ez, Ey_kp, Ey, inv);
            CE[1][e] = curl2<AT>(Ex_kp, Ex, Ez_ip, Ez, inv);
            CE[2][e] = curl2<AT>(Ey_ip, Ey, Ex_jp, Ex, inv);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            Ecur[0][e] = mul_rn((AT)mn[0].v[e], (AT)dn[0].v[e]);
            Ecur[1][e] = mul_rn((AT)mn[1].v[e], (AT)dn[1].v[e]);
            Ecur[2][e] = mul_rn((AT)mn[2].v[e], (AT)dn[2].v[e]);
        }

        // ---- update
        if (active) {
            Vec<T, V> out[3];
            if (!pml) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e)   // m1 = 1, m2 = -C0 dt: exactly what the general formula gives off the PML
                        out[c].v[e] = (T)add_rn((AT)h[c].v[e], mul_rn(s, CE[c][e]));
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    AT old[3] = {(AT)h[0].v[e], (AT)h[1].v[e], (AT)h[2].v[e]};
                    AT curl[3] = {CE[0][e], CE[1][e], CE[2][e]};
                    AT o3[3];
                    pml_cell<T, AT, true>(a, i, j, k0 + e, s, old, curl, o3);
                    out[0].v[e] = (T)o3[0];
                    out[1].v[e] = (T)o3[1];
                    out[2].v[e] = (T)o3[2];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Hout[c] + pbase + orow, out[c]);
        }
    }
}

// EXTRAS = dense J input and/or E output (the per-step forward() API); the fused run() path
// instantiates EXTRAS = false and carries neither.
template <typename T, typename AT, int V, int LZ, bool EXTRAS>
__global__ void __launch_bounds__(32 * V2_BY, V2_D_MIN_CTAS) k_step_D_v2(const StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    constexpr int RW = 32 / LZ;                  // rows per warp
    const int lane = threadIdx.x;
    const int lz = lane % LZ, ly = lane / LZ;
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int xc = rest / a.nty;
    const int jraw = (ty * V2_BY + threadIdx.y) * RW + ly;
    if (jraw - ly >= a.Ny) return;               // warp-uniform: the whole warp is below the grid
    const int k0raw = (tz * LZ + lz) * V;
    const bool active = k0raw < a.Nz && jraw < a.Ny;
    const int j = jraw < a.Ny ? jraw : a.Ny - 1; // inactive lanes shadow a valid cell (loads only)
    const int k0 = k0raw < a.Nz ? k0raw : 0;
    const int xs = a.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, a.x1);

    const int plane = a.Ny * a.Nz;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const bool z_edge = (lz == 0) || (k0 == 0);
    const int km = (k0 == 0) ? a.Nz - 1 : k0 - 1;
    const int orow = j * a.Nz + k0;
    const int orow_jm = jm * a.Nz + k0;
    const int okm = j * a.Nz + km;

    bool yz_pml = a.mapD[1][j] >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) yz_pml |= a.mapD[2][k0 + e] >= 0;
    const AT s = a.cdt;
    const AT inv = a.inv_dL;

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
        if (a.pf_dist > 0 && (lz & 7) == 0 && i + a.pf_dist < xe) {
            const int po = (i + a.pf_dist) * plane + orow;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                prefetch_l2(a.Hin[c] + po);
                prefetch_l2(a.Din[c] + po);
            }
        }
        const bool pml = yz_pml || (a.mapD[0][i] >= 0);

        AT CH[3][V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Hx = (AT)h[0].v[e], Hy = (AT)h[1].v[e], Hz = (AT)h[2].v[e];
            const AT Hx_km = (e > 0) ? (AT)h[0].v[(e + V - 1) % V] : hx_km;
            const AT Hy_km = (e > 0) ? (AT)h[1].v[(e + V - 1) % V] : hy_km;
            CH[0][e] = curl2<AT>(Hz, (AT)hzj.v[e], Hy, Hy_km, inv);
            CH[1][e] = curl2<AT>(Hx, Hx_km, Hz, Hprev[1][e], inv);
            CH[2][e] = curl2<AT>(Hy, Hprev[0][e], Hx, (AT)hxj.v[e], inv);
            Hprev[0][e] = Hy;
            Hprev[1][e] = Hz;
        }

        if (active) {
            Vec<T, V> out[3];
            if (!pml) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e) out[c].v[e] = (T)add_rn((AT)d[c].v[e], mul_rn(s, CH[c][e]));
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    AT old[3] = {(AT)d[0].v[e], (AT)d[1].v[e], (AT)d[2].v[e]};
                    AT curl[3] = {CH[0][e], CH[1][e], CH[2][e]};
                    AT o3[3];
                    pml_cell<T, AT, false>(a, i, j, k0 + e, s, old, curl, o3);
                    out[0].v[e] = (T)o3[0];
                    out[1].v[e] = (T)o3[1];
                    out[2].v[e] = (T)o3[2];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (EXTRAS) {
                    Vec<T, V> eo;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (a.J[c]) {
                            const AT sc = a.Jwave[c] ? (AT)(*a.Jwave[c]) : a.Jscale[c];
                            out[c].v[e] = (T)add_rn((AT)out[c].v[e], mul_rn((AT)jv[c].v[e], sc));
                        }
                        if (a.Eout[c]) eo.v[e] = (T)mul_rn((AT)mev[c].v[e], (AT)out[c].v[e]);
                    }
                    if (a.Eout[c]) stv<T, V>(a.Eout[c] + pbase + orow, eo);
                }
                stv<T, V>(a.Dout[c] + pbase + orow, out[c]);
            }
        }
    }
}

}  // namespace cev
