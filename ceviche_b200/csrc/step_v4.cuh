// Fused full-step kernel for sm_100a: the H half-step and the D half-step of one time step
// (fdtd.py:80-97 and :105-127) in ONE pass over the grid.
//
// Why: the two-sweep leap-frog moves 21 words per cell-update (H sweep 12, D sweep 9).  Marching along x,
// a CTA that has just produced H_new on plane i holds everything the D update of plane i needs except
// H_new of the row below (j-1) and of the cell before (k-1).  If the CTA also computes H_new on a one-cell
// halo (one extra row, one extra vector of lanes) those come from shared memory / the neighbouring lane,
// and the state is read once and written once per step: D, 1/eps, H in; H, D out = 15 words per cell.
//
// Consequences of the halo recomputation (all handled here):
//   * the state is ping-ponged (H_in -> H_out, D_in -> D_out): a neighbouring CTA still needs the OLD H and D
//     of the cells this CTA owns.  The H-side PML integrals (ICE, IH) are ping-ponged for the same reason
//     (a halo cell in the PML needs the old integral); the D-side ones (ICH, ID) are owner-only, in place.
//   * a chunk of x-planes [xs, xe) starts with a "pre-roll" iteration that rebuilds H_new of plane xs-1
//     (needed by the x-difference of curl_H) without storing anything.
//   * every value is computed by the same operation sequence as in step_v1/v2/v3.cuh, so results are
//     bit-identical to them (tests/test_gpu_variants.py).
//
// Thread layout: a warp covers LZ lanes along z (one 16-byte vector each) x 32/LZ rows; a CTA has BY warps,
// i.e. R = BY*32/LZ rows.  Row 0 and lane 0 of every row are the halo; a CTA owns (R-1) rows x (LZ-1) vectors.
#pragma once
#include "common.cuh"
#include "step_v2.cuh"

namespace cev {

// Two schedules of the same kernel (bit-identical): V4_PIPELINE 1 issues the loads of plane i+1 before the barrier and
// the D phase of plane i (needs ~255 registers: 2 CTAs of 128 threads per SM); 0 issues them at the top of their own
// iteration (fits 168 registers: 3 CTAs per SM).  Measured on B200 (scripts/tune.py, TUNE_RUN): see DESIGN.md.
#ifndef V4_PIPELINE
#define V4_PIPELINE 0
#endif
#ifndef V4_MIN_CTAS
#define V4_MIN_CTAS (V4_PIPELINE ? 2 : 3)
#endif

// H-side PML work with separate input / output integral arrays and a store switch (halo cells and the
// pre-roll plane compute the update but must not write).  Arithmetic identical to PmlCtx<.., true>.
template <typename T, typename AT, int V>
struct PmlCtxH4 {
    Vec<T, V> I0, I1;
    T I2[V];
    int ic0, ic1;

    __device__ __forceinline__ void load(const StepArgs<T, AT>& a, int i, int j, int k0, int mx, int my, const int* mz) {
        ic0 = mx >= 0 ? (mx * a.Ny + j) * a.Nz + k0 : -1;          // ICE_x (nHx,Ny,Nz)
        ic1 = my >= 0 ? (i * a.nH[1] + my) * a.Nz + k0 : -1;       // ICE_y (Nx,nHy,Nz)
        if (ic0 >= 0) I0 = ldv<T, V>(a.ICE[0] + ic0);
        if (ic1 >= 0) I1 = ldv<T, V>(a.ICE[1] + ic1);
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (mz[e] >= 0) I2[e] = a.ICE[2][(i * a.Ny + j) * a.nH[2] + mz[e]];   // ICE_z (Nx,Ny,nHz)
    }

    __device__ __forceinline__ void apply(const StepArgs<T, AT>& a, int i, int j, int k0, int mx, int my, const int* mz,
                                          AT s, const Vec<T, V>* old, const AT (*curl)[V], Vec<T, V>* out, bool store) {
        const int n1 = a.nH[1], n2 = a.nH[2];
        const AT ux = a.uH[0][i], rx = a.rH[0][i];
        const AT uy = a.uH[1][j], ry = a.rH[1][j];
        AT uz[V], rz[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            uz[e] = a.uH[2][k0 + e];
            rz[e] = a.rH[2][k0 + e];
        }
        Vec<T, V> n0, n1v;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int k = k0 + e;
            {   // x: (a,b) = (y,z), own x.   IH_x (Nx,nHy,nHz)
                AT m1, m2;
                coef12<AT>(uy, ry, uz[e], rz[e], s, m1, m2);
                AT v = muladd(m1, (AT)old[0].v[e], mul_rn(m2, curl[0][e]));
                if (ic0 >= 0) {
                    const AT I = (AT)I0.v[e] + curl[0][e];
                    n0.v[e] = (T)I;
                    v = muladd(mul_rn(mul_rn(s, ux + ux), mul_rn(ry, rz[e])), I, v);
                }
                if (my >= 0 && mz[e] >= 0) {
                    const int q = (i * n1 + my) * n2 + mz[e];
                    const AT I = (AT)a.IH[0][q] + (AT)old[0].v[e];
                    if (store) a.IHout[0][q] = (T)I;
                    v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), uy), uz[e]), mul_rn(ry, rz[e])), I, v);
                }
                out[0].v[e] = (T)v;
            }
            {   // y: (a,b) = (x,z), own y.   IH_y (nHx,Ny,nHz)
                AT m1, m2;
                coef12<AT>(ux, rx, uz[e], rz[e], s, m1, m2);
                AT v = muladd(m1, (AT)old[1].v[e], mul_rn(m2, curl[1][e]));
                if (ic1 >= 0) {
                    const AT I = (AT)I1.v[e] + curl[1][e];
                    n1v.v[e] = (T)I;
                    v = muladd(mul_rn(mul_rn(s, uy + uy), mul_rn(rx, rz[e])), I, v);
                }
                if (mx >= 0 && mz[e] >= 0) {
                    const int q = (mx * a.Ny + j) * n2 + mz[e];
                    const AT I = (AT)a.IH[1][q] + (AT)old[1].v[e];
                    if (store) a.IHout[1][q] = (T)I;
                    v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), ux), uz[e]), mul_rn(rx, rz[e])), I, v);
                }
                out[1].v[e] = (T)v;
            }
            {   // z: (a,b) = (x,y), own z.   IH_z (nHx,nHy,Nz)
                AT m1, m2;
                coef12<AT>(ux, rx, uy, ry, s, m1, m2);
                AT v = muladd(m1, (AT)old[2].v[e], mul_rn(m2, curl[2][e]));
                if (mz[e] >= 0) {
                    const AT I = (AT)I2[e] + curl[2][e];
                    if (store) a.ICEout[2][(i * a.Ny + j) * n2 + mz[e]] = (T)I;
                    v = muladd(mul_rn(mul_rn(s, uz[e] + uz[e]), mul_rn(rx, ry)), I, v);
                }
                if (mx >= 0 && my >= 0) {
                    const int q = (mx * n1 + my) * a.Nz + k;
                    const AT I = (AT)a.IH[2][q] + (AT)old[2].v[e];
                    if (store) a.IHout[2][q] = (T)I;
                    v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), ux), uy), mul_rn(rx, ry)), I, v);
                }
                out[2].v[e] = (T)v;
            }
        }
        if (store && ic0 >= 0) stv<T, V>(a.ICEout[0] + ic0, n0);
        if (store && ic1 >= 0) stv<T, V>(a.ICEout[1] + ic1, n1v);
    }
};

// V4_PML_NOINLINE 1 puts the PML work of the fused kernel behind a call boundary: inlined, the two general PML updates
// push the kernel to 240+ registers for EVERY thread, while the PML-free kernel needs 128 (4 CTAs per SM) and beats the
// two-kernel path by 20-25 % (profiles/r1_tune_fused_nopml.log).  Out of line the main path keeps that budget, but the
// PML iterations then pay the call, the stack traffic and integral loads issued late, and with a barrier per plane the
// slow warps hold the CTA: measured 14.5 instead of 20.9 Gcell/s at 256^3 fp64 (profiles/r1_tune_fused_pml_noinline.log).
// Kept as an experiment switch, off.
#ifndef V4_PML_NOINLINE
#define V4_PML_NOINLINE 0
#endif
template <typename T, typename AT, int V>
__device__ __noinline__ void v4_pml_H(const StepArgs<T, AT>& a, int i, int j, int k0, int mx, int my, const int* mz, AT s,
                                      const Vec<T, V>* h, const AT (*CE)[V], Vec<T, V>* out, bool store) {
    PmlCtxH4<T, AT, V> ctx;
    ctx.load(a, i, j, k0, mx, my, mz);
    ctx.apply(a, i, j, k0, mx, my, mz, s, h, CE, out, store);
}
template <typename T, typename AT, int V>
__device__ __noinline__ void v4_pml_D(const StepArgs<T, AT>& a, int i, int j, int k0, int mx, int my, const int* mz, AT s,
                                      const Vec<T, V>* d, const AT (*CH)[V], Vec<T, V>* out) {
    PmlCtx<T, AT, V, false> ctx;
    ctx.load(a, i, j, k0, mx, my, mz, 7u);
    ctx.apply(a, i, j, k0, mx, my, mz, s, d, CH, out, 7u);
}

// NOPML = the plan has no PML on any axis: every PML branch is compiled out (the lean kernel: what the fused
// formulation can do when the register budget is not spent on the PML paths).
template <typename T, typename AT, int V, int LZ, int BY, bool NOPML = false>
__global__ void __launch_bounds__(32 * BY, (BY >= 8 ? 1 : ((NOPML || V4_PML_NOINLINE) ? 4 : V4_MIN_CTAS))) k_step_fused(const __grid_constant__ StepArgs<T, AT> a) {
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    constexpr int RW = 32 / LZ;                  // rows per warp
    constexpr int R = BY * RW;                   // rows per CTA (row 0 = halo)
    constexpr int OY = R - 1, OZ = LZ - 1;       // owned rows / owned vectors per CTA
    __shared__ Vec<T, V> sh[2][2][R][LZ];        // [ping-pong][Hx | Hz][row][lane]: H_new of the current plane

    const int lane = threadIdx.x;
    const int lz = lane % LZ;
    const int r = threadIdx.y * RW + lane / LZ;
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int xc = rest / a.nty;

    // the launch covers box[0] (the whole grid, or -- hybrid path -- the PML-free interior); tiles are laid out from
    // the box origin, halo rows / lanes may lie outside the box (they are only read)
    const Box& B = a.box[0];
    const int Nzv = a.Nz / V, bz0 = B.z0 / V, bz1 = B.z1 / V;
    int zv = bz0 + tz * OZ - 1 + lz;
    const bool own_z = lz >= 1 && zv < min(bz0 + (tz + 1) * OZ, bz1);
    if (zv < 0) zv = Nzv - 1;                    // halo lane of the first tile: periodic wrap
    if (zv >= Nzv) zv = Nzv - 1;                 // lanes past the row shadow a valid vector (no stores)
    int j = B.y0 + ty * OY - 1 + r;
    const bool own_y = r >= 1 && j < min(B.y0 + (ty + 1) * OY, B.y1);
    if (j < 0) j = a.Ny - 1;
    if (j >= a.Ny) j = a.Ny - 1;
    const bool own = own_y && own_z;
    const int k0 = zv * V;
    const int xs = B.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, B.x1);

    const int plane = a.Ny * a.Nz;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const bool z_edge = (lz == LZ - 1) || (k0 + V >= a.Nz);   // +1 neighbour is not in lane+1
    const int kp = (k0 + V >= a.Nz) ? 0 : k0 + V;
    const int orow = j * a.Nz + k0;
    const int orow_jp = jp * a.Nz + k0;
    const int okp = j * a.Nz + kp;

    const int myH = a.mapH[1][j], myD = a.mapD[1][j];
    int mzH[V], mzD[V];
    bool yz_pmlH = myH >= 0, yz_pmlD = myD >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mzH[e] = a.mapH[2][k0 + e];
        mzD[e] = a.mapD[2][k0 + e];
        yz_pmlH |= mzH[e] >= 0;
        yz_pmlD |= mzD[e] >= 0;
    }
    const AT sH = -a.cdt, sD = a.cdt;
    const AT inv = a.inv_dL;

    // E = mE*D of the pre-roll plane xs-1 (own cells)
    AT Ecur[3][V];
    {
        const int o = ((xs == 0) ? a.Nx - 1 : xs - 1) * plane + orow;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Vec<T, V> d = ldv<T, V>(a.Din[c] + o), m = ldv<T, V>(a.mE[c] + o);
#pragma unroll
            for (int e = 0; e < V; ++e) Ecur[c][e] = mul_rn((AT)m.v[e], (AT)d.v[e]);
        }
    }
    Vec<T, V> dcur[3];            // D_old of the current plane (loaded one iteration earlier as "next")
    AT Hprev[2][V];               // H_new(y, z) of the previous plane, as stored (rounded to T)
#pragma unroll
    for (int e = 0; e < V; ++e) {
        Hprev[0][e] = AT(0);
        Hprev[1][e] = AT(0);
#pragma unroll
        for (int c = 0; c < 3; ++c) dcur[c].v[e] = T(0);
    }

    // Loads of one iteration.  They are issued one iteration AHEAD, right after the H phase has consumed the
    // previous set and before the barrier + D phase, so their latency overlaps the D half-step of the plane
    // before (the compiler does not move memory operations across the barrier).
    Vec<T, V> dn[3], mn[3], h[3], dxj, mxj, dzj, mzj;
    T sx_m = T(0), sx_d = T(0), sy_m = T(0), sy_d = T(0);     // z-edge lanes: 1/eps and D of the cell after the vector
    PmlCtxH4<T, AT, V> ctxH;
    auto issue = [&](int ii) {
        const int i = (ii < 0) ? a.Nx - 1 : ii;
        const int in = (ii + 1 == a.Nx) ? 0 : ii + 1;
        const int pbase = i * plane, nbase = in * plane;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dn[c] = ldv<T, V>(a.Din[c] + nbase + orow);
            mn[c] = ldv<T, V>(a.mE[c] + nbase + orow);
            h[c] = ldv<T, V>(a.Hin[c] + pbase + orow);
        }
        dxj = ldv<T, V>(a.Din[0] + pbase + orow_jp);
        mxj = ldv<T, V>(a.mE[0] + pbase + orow_jp);
        dzj = ldv<T, V>(a.Din[2] + pbase + orow_jp);
        mzj = ldv<T, V>(a.mE[2] + pbase + orow_jp);
        if (z_edge) {
            sx_m = a.mE[0][pbase + okp];
            sx_d = a.Din[0][pbase + okp];
            sy_m = a.mE[1][pbase + okp];
            sy_d = a.Din[1][pbase + okp];
        }
        if (a.pf_dist > 0 && ii + 1 + a.pf_dist < a.Nx && (lz & 7) == 0) {   // a later plane into L2
            const int po = (ii + 1 + a.pf_dist) * plane + orow;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                prefetch_l2(a.Din[c] + po);
                prefetch_l2(a.mE[c] + po);
                prefetch_l2(a.Hin[c] + po - plane);
            }
        }
        const int mxH = a.mapH[0][i];
        if (!NOPML && !V4_PML_NOINLINE && (yz_pmlH || mxH >= 0)) ctxH.load(a, i, j, k0, mxH, myH, mzH);
    };

    int buf = 0;
#if V4_PIPELINE
    issue(xs - 1);
#endif
    for (int ii = xs - 1; ii < xe; ++ii) {
#if !V4_PIPELINE
        issue(ii);
#endif
        const bool pre = ii < xs;                              // pre-roll: H_new only, nothing stored (CTA-uniform)
        const int i = (ii < 0) ? a.Nx - 1 : ii;
        const int pbase = i * plane;
        const bool st_on = own && !pre;
        const int mxH = a.mapH[0][i], mxD = a.mapD[0][i];
        const bool pmlH = !NOPML && (yz_pmlH || mxH >= 0);
        const bool pmlD = !NOPML && st_on && (yz_pmlD || mxD >= 0);
        PmlCtx<T, AT, V, false> ctxD;
        if (pmlD && !V4_PML_NOINLINE) ctxD.load(a, i, j, k0, mxD, myD, mzD, 7u);   // consumed after the barrier

        // ---- curl_E and H_new of plane i (own cells + halo row / halo lane)
        AT ex_kp = __shfl_down_sync(0xffffffffu, Ecur[0][0], 1);
        AT ey_kp = __shfl_down_sync(0xffffffffu, Ecur[1][0], 1);
        if (z_edge) {
            ex_kp = mul_rn((AT)sx_m, (AT)sx_d);
            ey_kp = mul_rn((AT)sy_m, (AT)sy_d);
        }
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
        Vec<T, V> hn[3];
        if (!pmlH) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int e = 0; e < V; ++e) hn[c].v[e] = (T)muladd(sH, CE[c][e], (AT)h[c].v[e]);
        } else {
#if V4_PML_NOINLINE
            Vec<T, V> h2[3], o2[3];
            AT ce2[3][V];
            int mz2[V];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                h2[c] = h[c];
#pragma unroll
                for (int e = 0; e < V; ++e) ce2[c][e] = CE[c][e];
            }
#pragma unroll
            for (int e = 0; e < V; ++e) mz2[e] = mzH[e];
            v4_pml_H<T, AT, V>(a, i, j, k0, mxH, myH, mz2, sH, h2, ce2, o2, st_on);
#pragma unroll
            for (int c = 0; c < 3; ++c) hn[c] = o2[c];
#else
            ctxH.apply(a, i, j, k0, mxH, myH, mzH, sH, h, CE, hn, st_on);
#endif
        }
        if (st_on) {
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Hout[c] + pbase + orow, hn[c]);
        }
        Vec<T, V> dnext[3];       // D_old of plane i+1: the next iteration's D phase needs it
#pragma unroll
        for (int c = 0; c < 3; ++c) dnext[c] = dn[c];
#if V4_PIPELINE
        if (ii + 1 < xe) issue(ii + 1);
#endif

        // ---- D half-step of plane i from H_new: row j-1 through shared memory, cell k-1 from the lane before
        if (!pre) {
            sh[buf][0][r][lz] = hn[0];
            sh[buf][1][r][lz] = hn[2];
        }
        __syncthreads();
        if (!pre) {
            const int rm = r > 0 ? r - 1 : 0;
            const Vec<T, V> hxj = sh[buf][0][rm][lz], hzj = sh[buf][1][rm][lz];
            const AT hx_km = __shfl_up_sync(0xffffffffu, (AT)hn[0].v[V - 1], 1);
            const AT hy_km = __shfl_up_sync(0xffffffffu, (AT)hn[1].v[V - 1], 1);
            if (own) {
                AT CH[3][V];
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const AT Hx = (AT)hn[0].v[e], Hy = (AT)hn[1].v[e], Hz = (AT)hn[2].v[e];
                    const AT Hx_km = (e > 0) ? (AT)hn[0].v[(e + V - 1) % V] : hx_km;
                    const AT Hy_km = (e > 0) ? (AT)hn[1].v[(e + V - 1) % V] : hy_km;
                    CH[0][e] = curl2<AT>(Hz, (AT)hzj.v[e], Hy, Hy_km, inv);
                    CH[1][e] = curl2<AT>(Hx, Hx_km, Hz, Hprev[1][e], inv);
                    CH[2][e] = curl2<AT>(Hy, Hprev[0][e], Hx, (AT)hxj.v[e], inv);
                }
                Vec<T, V> out[3];
                if (!pmlD) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int e = 0; e < V; ++e) out[c].v[e] = (T)muladd(sD, CH[c][e], (AT)dcur[c].v[e]);
                } else {
#if V4_PML_NOINLINE
                    Vec<T, V> d2[3], o2[3];
                    AT ch2[3][V];
                    int mz2[V];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        d2[c] = dcur[c];
#pragma unroll
                        for (int e = 0; e < V; ++e) ch2[c][e] = CH[c][e];
                    }
#pragma unroll
                    for (int e = 0; e < V; ++e) mz2[e] = mzD[e];
                    v4_pml_D<T, AT, V>(a, i, j, k0, mxD, myD, mz2, sD, d2, ch2, o2);
#pragma unroll
                    for (int c = 0; c < 3; ++c) out[c] = o2[c];
#else
                    ctxD.apply(a, i, j, k0, mxD, myD, mzD, sD, dcur, CH, out, 7u);
#endif
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) stv<T, V>(a.Dout[c] + pbase + orow, out[c]);
            }
            buf ^= 1;
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            Hprev[0][e] = (AT)hn[1].v[e];
            Hprev[1][e] = (AT)hn[2].v[e];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) dcur[c] = dnext[c];
    }

    // ---- in-kernel source injection: D += J after the update (fdtd.py:125-127)
    if (a.src_wave) {
        __syncthreads();
        inject_points<T, AT, int32_t>(a.src_begin[bid], a.src_begin[bid + 1], threadIdx.y * 32 + threadIdx.x, 32 * BY, a.src_comp,
                                      a.src_id, a.src_cell, a.src_w, a.src_wave, a.Dout[0], a.Dout[1], a.Dout[2]);
    }
}

}  // namespace cev
