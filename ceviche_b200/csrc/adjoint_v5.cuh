// Tensor-map TMA kernels of the reverse (adjoint) sweep for sm_100a: the transposed FDTD step as TWO marching
// kernels with the data movement of the forward half-step kernels (step_v5.cuh).
//
// The reference differentiates ceviche/fdtd.py:74-144 by taping every numpy op (ceviche/jacobians.py:29-35); the
// transposed step is written out in adjoint.cuh.  Here it is re-phased so that BOTH stencils read a stored array and
// every PML operation is local to the thread's own cells ("eager" form).  Between two steps the sweep carries
//     lH            cotangent of H after step k (true),
//     lDp           cotangent of D after step k with the cell-local D part of step k ALREADY applied: m1D g + gID,
//     gC            cotangent of curl_H(H) of step k: m2D g + gICH            (g = the true cotangent of D_k),
//     lICE, lIH     true; lICH, lID already advanced by step k's D part,
// and one step back in time is
//     k_adj_H_v5    gH = lH + curl_E(gC)  [forward differences, derivatives.py:16-22, of a STORED array];
//                   H part (fdtd.py:95-97 transposed): lICE += m3H gH, lIH += m4H gH, lH <- m1H gH + lIH,
//                   gC2 <- m2H gH + lICE.                      gC[3], lH[3] in; lH[3], gC2[3] out = 12 words / cell
//     k_adj_ED_v5   lE = curl_H(gC2)      [backward differences, derivatives.py:24-30];  g = lDp + mE lE is the true
//                   cotangent of D_{k-1};  G_mE += lE D_{k-1} inside the design box;  then the D part of step k-1 at
//                   once: lICH += m3D g, lID += m4D g, lDp <- m1D g + lID, gC <- m2D g + lICH.
//                   gC2[3], lDp[3], mE[3] in; lDp[3], gC[3] out = 15 words / cell (+ 9 inside the design box)
// 27 words per cell and step, none of them a re-read; off the PML the local parts are the identity / one scaling.
// Probe-series seeds enter between the two kernels of consecutive steps in the same eager form (k_adj_seed_eager).
// A segment starts with k_adj_Dlocal (true lD -> eager form) and its last k_adj_ED_v5 launch leaves the true lD
// (eager = 0), so the cotangent state handed between cev_fdtd_adjoint_run calls is the plain one.
//
// The adjoint has a tolerance to meet (1e-10 against the autograd oracle), not the reference's rounding sequence:
// plain arithmetic, FMA contraction allowed.
#pragma once
#include "adjoint.cuh"
#include "step_v5.cuh"

namespace cev {

// what the two kernels need beyond StepArgs.  StepArgs fields are re-used as follows.
//   k_adj_H_v5 : Din = gC, Hin = Hout = lH, Eout = gC2, ICE = lICE, IH = lIH
//   k_adj_ED_v5: Hin = gC2, Din = Dout = lDp, mE = 1/eps, Eout = gC, ICH = lICH, ID = lID
template <typename T>
struct AdjV5Extra {
    const T* Dprev[3];     // forward D after step k-1 (read inside the design box only)
    double* G[3];          // dL/d(1/eps), fp64, nullable per component
    int gb[6];             // design box x0, x1, y0, y1, z0, z1 (internal axes)
    int eager;             // 1: apply the D part of step k-1 (gC, lDp, integrals); 0: leave the true lD in Dout
    int boxed;             // 1: Dprev[c] holds the design box only, C-order (gb extents): the forward run's D-box record
};

struct V5MapsAdjH {
    CUtensorMap C[3];      // main boxes of gC
    CUtensorMap L[3];      // pln boxes of lH
    CUtensorMap Crow[2];   // row boxes of gC: components x, z
};
struct V5MapsAdjED {
    CUtensorMap C[3];      // main boxes of gC2
    CUtensorMap L[3];      // pln boxes of lDp
    CUtensorMap M[3];      // pln boxes of 1/eps
    CUtensorMap Crow[2];   // row boxes of gC2: components x, z
};

// minimum resident CTAs per SM asked of the compiler (register cap = 65536 / (threads * this)).  Measured on B200
// (profiles/r2_tune_adjoint_min_blocks.log, config 4): fp32 storage gains 14 % with 3 CTAs (168 registers instead of
// 214-254: the four-cell vectors of a thread), fp64 is indifferent, 4 CTAs spill and lose.
#ifndef CEV_ADJ_MINB_F32
#define CEV_ADJ_MINB_F32 3
#endif
#ifndef CEV_ADJ_MINB_F64
#define CEV_ADJ_MINB_F64 0
#endif
#define CEV_ADJ_MINB(T) (sizeof(T) == 4 ? CEV_ADJ_MINB_F32 : CEV_ADJ_MINB_F64)

template <typename T, int V, int BY>
struct AdjV5Layout {
    using L = V5Layout<T, V, BY>;
    static constexpr int H_STAGE = 3 * L::BLK + 3 * L::PLN;
    static constexpr int ED_STAGE = 3 * L::BLK + 6 * L::PLN;
    static constexpr size_t h_bytes(int ns) { return (size_t)ns * H_STAGE * sizeof(T) + (size_t)ns * sizeof(uint64_t); }
    static constexpr size_t ed_bytes(int ns) { return (size_t)ns * ED_STAGE * sizeof(T) + (size_t)ns * sizeof(uint64_t); }
};

// ---------------------------------------------------------------------------------------------------------
// H part of the transposed step.  Same tiling, ring and wrap handling as k_step_H_v5.
template <typename T, typename AT, int V, int BY, int NS>
__global__ void __launch_bounds__(32 * BY, CEV_ADJ_MINB(T)) k_adj_H_v5(const StepArgs<T, AT> a, const __grid_constant__ V5MapsAdjH maps) {
    using L = V5Layout<T, V, BY>;
    constexpr int BZ = L::BZ, ROWP = L::ROWP, BLK = L::BLK, LOFF = 3 * L::BLK;
    constexpr int STAGE = AdjV5Layout<T, V, BY>::H_STAGE;
    extern __shared__ __align__(128) unsigned char v5_smem[];
    const int bid = blockIdx.x;
    T* const stage0 = reinterpret_cast<T*>(v5_smem);
    uint64_t* const full = reinterpret_cast<uint64_t*>(v5_smem + (size_t)NS * STAGE * sizeof(T));
    const int lane = threadIdx.x, w = threadIdx.y;
    const V5Tile t = v5_locate<V, BY>(a, bid, lane, w);
    const int xs = t.xs, xe = t.xe, y0 = t.y0, z0 = t.z0;
    const int j = y0 + t.row;
    const bool active = z0 + t.vec * V < a.Nz;
    const int k0 = active ? z0 + t.vec * V : z0;
    const int plane = a.Ny * a.Nz;

    __shared__ V5XTab<AT> xt;
    v5_fill_xtab<AT>(xt, a.mapH[0], a.uH[0], a.rH[0], xs, xe, a.x1, w * 32 + lane, 32 * BY);
    if (lane == 0 && w == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], BY);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // plane p -> stage (p - xs) % NS.  p < xe: 8 boxes (gC[3] with halos, lH[3], the halo rows of gC_x and gC_z);
    // p == xe: the x+1 neighbour of the chunk's last plane, own rows of gC_y, gC_z (plane 0 when xe == Nx: periodic)
    const int yrow = (y0 + BY >= a.Ny) ? 0 : y0 + BY;
    auto box_cur = [&](int q, T* st, int p, uint64_t* bar) {
        if (q < 6) {
            const int c = q / 2;
            if (q % 2 == 0) tma_box_3d(st + c * BLK, &maps.C[c], z0, y0, p, bar);
            else tma_box_3d(st + LOFF + c * L::PLN, &maps.L[c], z0, y0, p, bar);
        } else {
            const int r = q - 6;                           // 0: component x, 1: component z
            tma_box_3d(st + 2 * r * BLK + BY * ROWP, &maps.Crow[r], z0, yrow, p, bar);
        }
    };
    constexpr uint32_t BM = L::BOX_MAIN, BR = L::BOX_ROW, BP = L::BOX_PLN;
    auto issue = [&](int p, int s) {                       // called by every warp, converged
        if (!elect_one()) return;
        uint64_t* bar = &full[s];
        T* st = stage0 + (size_t)s * STAGE;
        if (p < xe) {
            uint32_t bytes = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q % BY == w) bytes += q >= 6 ? BR : (q % 2 == 1 ? BP : BM);
            mbar_arrive_tx(bar, bytes);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q % BY == w) box_cur(q, st, p, bar);
        } else {
            mbar_arrive_tx(bar, w < 2 ? BM : 0u);
            if (w < 2) tma_box_3d(st + (1 + w) * BLK, &maps.C[1 + w], z0, y0, p >= a.Nx ? 0 : p, bar);
        }
    };
#pragma unroll
    for (int d = 0; d < NS - 1; ++d)
        if (xs + d <= xe) issue(xs + d, d);

    const AT s = -a.cdt;
    const AT inv = a.inv_dL;
    PmlLean<T, AT, V, true> pml;
    pml.init(a, j, k0, s);
    const int warp_yz = v5_warp_flags(pml.fy, pml.fz_any());
    const int orow = j * a.Nz + k0;
    const bool zedge = k0 + V >= a.Nz;               // the +1 z-neighbour of this lane's last cell is k = 0
    const int col = t.vec * V;
    const int r0 = t.row * ROWP + col, r1 = r0 + ROWP;
    const int lrow = LOFF + t.row * BZ + col;

    int sc = 0;
    uint32_t ph = 0;
    int sp = NS - 1;
    T gn0 = T(0), gn1 = T(0);
    if (zedge) {
        const int okp = xs * plane + j * a.Nz;
        gn0 = a.Din[0][okp]; gn1 = a.Din[1][okp];
    }
    for (int i = xs; i < xe; ++i) {
        if (i + NS - 1 <= xe) issue(i + NS - 1, sp);
        sp = (sp + 1 == NS) ? 0 : sp + 1;
        const int sn = (sc + 1 == NS) ? 0 : sc + 1;
        const uint32_t phn = (sn == 0) ? ph ^ 1u : ph;
        const int q = i - xs;
        const int mx = xt.mx[q];
        const bool in_pml = pml.yz || mx >= 0;
        if (in_pml && active) pml.load(a, i, mx);
        {
            const int mxp = xt.mx[q + V5_PF];
            if (active && (pml.yz || mxp >= 0) && i + V5_PF < a.x1) pml.prefetch(a, i + V5_PF, mxp, (lane & 7) == 0);
        }
        const int pbase = i * plane;
        const T gx0 = gn0, gx1 = gn1;
        if (zedge && i + 1 < xe) {
            const int okp = pbase + plane + j * a.Nz;
            gn0 = a.Din[0][okp]; gn1 = a.Din[1][okp];
        }
        mbar_wait(&full[sc], ph);
        mbar_wait(&full[sn], phn);
        const T* cur = stage0 + (size_t)sc * STAGE;
        const T* nxt = stage0 + (size_t)sn * STAGE;

        Vec<T, V> c[3], h[3];
#pragma unroll
        for (int q3 = 0; q3 < 3; ++q3) {
            c[q3] = ldv<T, V>(cur + q3 * BLK + r0);
            h[q3] = ldv<T, V>(cur + lrow + q3 * L::PLN);
        }
        const Vec<T, V> cxj = ldv<T, V>(cur + 0 * BLK + r1), czj = ldv<T, V>(cur + 2 * BLK + r1);
        const Vec<T, V> cyn = ldv<T, V>(nxt + 1 * BLK + r0), czn = ldv<T, V>(nxt + 2 * BLK + r0);
        AT cx_kp, cy_kp;
        if (zedge) {
            cx_kp = (AT)gx0;
            cy_kp = (AT)gx1;
        } else {
            cx_kp = (AT)cur[0 * BLK + r0 + V];
            cy_kp = (AT)cur[1 * BLK + r0 + V];
        }
        AT g[3][V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Cx = (AT)c[0].v[e], Cy = (AT)c[1].v[e], Cz = (AT)c[2].v[e];
            const AT Cx_kp = (e + 1 < V) ? (AT)c[0].v[(e + 1) % V] : cx_kp;
            const AT Cy_kp = (e + 1 < V) ? (AT)c[1].v[(e + 1) % V] : cy_kp;
            g[0][e] = (AT)h[0].v[e] + (((AT)czj.v[e] - Cz) - (Cy_kp - Cy)) * inv;
            g[1][e] = (AT)h[1].v[e] + ((Cx_kp - Cx) - ((AT)czn.v[e] - Cz)) * inv;
            g[2][e] = (AT)h[2].v[e] + (((AT)cyn.v[e] - Cy) - ((AT)cxj.v[e] - Cx)) * inv;
        }
        if (active) {
            Vec<T, V> lout[3], gc[3];
            const int path = (mx >= 0 ? 1 : 0) | warp_yz;
            if (path == 0) {
#pragma unroll
                for (int q3 = 0; q3 < 3; ++q3)
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        lout[q3].v[e] = (T)g[q3][e];
                        gc[q3].v[e] = (T)(s * g[q3][e]);
                    }
            } else {
                pml.apply_adj(a, i, mx, xt.u[q], xt.r[q], s, g, lout, gc);
            }
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3) {
                stv<T, V>(a.Hout[q3] + pbase + orow, lout[q3]);
                stv<T, V>(a.Eout[q3] + pbase + orow, gc[q3]);
            }
        }
        __syncthreads();
        sc = sn;
        ph = phn;
    }
}

// ---------------------------------------------------------------------------------------------------------
// E and D parts.  Same tiling, ring and wrap handling as k_step_D_v5 (boxes of the stencil input start one vector
// before the tile along z; the halo row j-1 is row BY of each block).
template <typename T, typename AT, int V, int BY, int NS>
__global__ void __launch_bounds__(32 * BY, CEV_ADJ_MINB(T)) k_adj_ED_v5(const StepArgs<T, AT> a, const __grid_constant__ V5MapsAdjED maps,
                                                        const AdjV5Extra<T> x) {
    using L = V5Layout<T, V, BY>;
    constexpr int BZ = L::BZ, ROWP = L::ROWP, BLK = L::BLK, LOFF = 3 * L::BLK, MOFF = 3 * L::BLK + 3 * L::PLN;
    constexpr int STAGE = AdjV5Layout<T, V, BY>::ED_STAGE;
    extern __shared__ __align__(128) unsigned char v5_smem[];
    const int bid = blockIdx.x;
    T* const stage0 = reinterpret_cast<T*>(v5_smem);
    uint64_t* const full = reinterpret_cast<uint64_t*>(v5_smem + (size_t)NS * STAGE * sizeof(T));
    const int lane = threadIdx.x, w = threadIdx.y;
    const V5Tile t = v5_locate<V, BY>(a, bid, lane, w);
    const int xs = t.xs, xe = t.xe, y0 = t.y0, z0 = t.z0;
    const int j = y0 + t.row;
    const bool active = z0 + t.vec * V < a.Nz;
    const int k0 = active ? z0 + t.vec * V : z0;
    const int plane = a.Ny * a.Nz;

    __shared__ V5XTab<AT> xt;
    v5_fill_xtab<AT>(xt, a.mapD[0], a.uD[0], a.rD[0], xs, xe, a.x1, w * 32 + lane, 32 * BY);
    if (lane == 0 && w == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], BY);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // plane p (p >= xs-1) -> stage (p - xs + 1) % NS.  p >= xs: 11 boxes (gC2[3] with halos, lDp[3], 1/eps[3], the halo
    // rows of gC2_x and gC2_z); p == xs-1: own rows of gC2_y, gC2_z (plane Nx-1 when xs == 0: periodic)
    const int yrow = (y0 == 0) ? a.Ny - 1 : y0 - 1;
    auto box_cur = [&](int q, T* st, int p, uint64_t* bar) {
        if (q < 9) {
            const int c = q / 3, kind = q % 3;
            if (kind == 0) tma_box_3d(st + c * BLK, &maps.C[c], z0 - V, y0, p, bar);
            else if (kind == 1) tma_box_3d(st + LOFF + c * L::PLN, &maps.L[c], z0, y0, p, bar);
            else tma_box_3d(st + MOFF + c * L::PLN, &maps.M[c], z0, y0, p, bar);
        } else {
            const int r = q - 9;                           // 0: component x, 1: component z
            tma_box_3d(st + 2 * r * BLK + BY * ROWP, &maps.Crow[r], z0 - V, yrow, p, bar);
        }
    };
    constexpr uint32_t BM = L::BOX_MAIN, BR = L::BOX_ROW, BP = L::BOX_PLN;
    auto issue = [&](int p, int s) {
        if (!elect_one()) return;
        uint64_t* bar = &full[s];
        T* st = stage0 + (size_t)s * STAGE;
        if (p >= xs) {
            uint32_t bytes = 0;
#pragma unroll
            for (int q = 0; q < 11; ++q)
                if (q % BY == w) bytes += q >= 9 ? BR : (q % 3 == 0 ? BM : BP);
            mbar_arrive_tx(bar, bytes);
#pragma unroll
            for (int q = 0; q < 11; ++q)
                if (q % BY == w) box_cur(q, st, p, bar);
        } else {                                     // p == xs - 1
            mbar_arrive_tx(bar, w < 2 ? BM : 0u);
            if (w < 2) tma_box_3d(st + (1 + w) * BLK, &maps.C[1 + w], z0 - V, y0, p < 0 ? a.Nx - 1 : p, bar);
        }
    };
#pragma unroll
    for (int d = 0; d < NS - 1; ++d)
        if (xs - 1 + d < xe) issue(xs - 1 + d, d);

    const AT s = a.cdt;
    const AT inv = a.inv_dL;
    PmlLean<T, AT, V, false> pml;
    pml.init(a, j, k0, s);
    const int warp_yz = v5_warp_flags(pml.fy, pml.fz_any());
    const int orow = j * a.Nz + k0;
    const bool zedge = k0 == 0;                      // the -1 z-neighbour of this lane's first cell is k = Nz-1
    const int col = V + t.vec * V;
    const int r0 = t.row * ROWP + col, rm = (t.row == 0 ? BY : t.row - 1) * ROWP + col;
    const int lrow = LOFF + t.row * BZ + t.vec * V, mrow = MOFF + t.row * BZ + t.vec * V;
    // design box: which of this thread's cells accumulate G (rows / columns are loop-invariant)
    bool gbox_e[V], gbox_any = false;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        gbox_e[e] = active && j >= x.gb[2] && j < x.gb[3] && k0 + e >= x.gb[4] && k0 + e < x.gb[5];
        gbox_any |= gbox_e[e];
    }
    // where the forward D of this thread's cells lives: the full-grid array, or the box record
    const int bzx = x.gb[5] - x.gb[4], bpl = (x.gb[3] - x.gb[2]) * bzx;
    const int drow = x.boxed ? (j - x.gb[2]) * bzx + (k0 - x.gb[4]) : orow;

    int sc = 0;                                      // stage of plane i-1
    uint32_t ph = 0;
    int sp = NS - 1;
    T gn0 = T(0), gn1 = T(0);
    if (zedge) {
        const int okm = xs * plane + j * a.Nz + a.Nz - 1;
        gn0 = a.Hin[0][okm]; gn1 = a.Hin[1][okm];
    }
    for (int i = xs; i < xe; ++i) {
        if (i + NS - 2 < xe) issue(i + NS - 2, sp);
        sp = (sp + 1 == NS) ? 0 : sp + 1;
        const int sn = (sc + 1 == NS) ? 0 : sc + 1;
        const uint32_t phn = (sn == 0) ? ph ^ 1u : ph;
        const int q = i - xs;
        const int mx = xt.mx[q];
        const bool in_pml = x.eager && (pml.yz || mx >= 0);
        if (in_pml && active) pml.load(a, i, mx);
        if (x.eager) {
            const int mxp = xt.mx[q + V5_PF];
            if (active && (pml.yz || mxp >= 0) && i + V5_PF < a.x1) pml.prefetch(a, i + V5_PF, mxp, (lane & 7) == 0);
        }
        const int pbase = i * plane;
        // design box: forward D and the fp64 accumulators of this plane (loaded before the barrier wait)
        const bool gacc = gbox_any && i >= x.gb[0] && i < x.gb[1];
        T dprev[3][V];
        double gold[3][V];
        if (gacc) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if (gbox_e[e] && x.G[c]) {
                        dprev[c][e] = x.Dprev[c][(x.boxed ? (i - x.gb[0]) * bpl : pbase) + drow + e];
                        gold[c][e] = x.G[c][(int64_t)pbase + orow + e];
                    }
        }
        const T gx0 = gn0, gx1 = gn1;
        if (zedge && i + 1 < xe) {
            const int okm = pbase + plane + j * a.Nz + a.Nz - 1;
            gn0 = a.Hin[0][okm]; gn1 = a.Hin[1][okm];
        }
        mbar_wait(&full[sc], ph);
        mbar_wait(&full[sn], phn);
        const T* prv = stage0 + (size_t)sc * STAGE;
        const T* cur = stage0 + (size_t)sn * STAGE;

        Vec<T, V> c[3], d[3], m[3];
#pragma unroll
        for (int q3 = 0; q3 < 3; ++q3) {
            c[q3] = ldv<T, V>(cur + q3 * BLK + r0);
            d[q3] = ldv<T, V>(cur + lrow + q3 * L::PLN);
            m[q3] = ldv<T, V>(cur + mrow + q3 * L::PLN);
        }
        const Vec<T, V> cxj = ldv<T, V>(cur + 0 * BLK + rm), czj = ldv<T, V>(cur + 2 * BLK + rm);
        const Vec<T, V> cyp = ldv<T, V>(prv + 1 * BLK + r0), czp = ldv<T, V>(prv + 2 * BLK + r0);
        AT cx_km, cy_km;
        if (zedge) {
            cx_km = (AT)gx0;
            cy_km = (AT)gx1;
        } else {
            cx_km = (AT)cur[0 * BLK + r0 - 1];
            cy_km = (AT)cur[1 * BLK + r0 - 1];
        }
        AT lE[3][V], g[3][V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Cx = (AT)c[0].v[e], Cy = (AT)c[1].v[e], Cz = (AT)c[2].v[e];
            const AT Cx_km = (e > 0) ? (AT)c[0].v[(e + V - 1) % V] : cx_km;
            const AT Cy_km = (e > 0) ? (AT)c[1].v[(e + V - 1) % V] : cy_km;
            lE[0][e] = ((Cz - (AT)czj.v[e]) - (Cy - Cy_km)) * inv;
            lE[1][e] = ((Cx - Cx_km) - (Cz - (AT)czp.v[e])) * inv;
            lE[2][e] = ((Cy - (AT)cyp.v[e]) - (Cx - (AT)cxj.v[e])) * inv;
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3) g[q3][e] = (AT)d[q3].v[e] + (AT)m[q3].v[e] * lE[q3][e];
        }
        if (gacc) {
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3)
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if (gbox_e[e] && x.G[q3])
                        x.G[q3][(int64_t)pbase + orow + e] = gold[q3][e] + (double)lE[q3][e] * (double)dprev[q3][e];
        }
        if (active) {
            Vec<T, V> lout[3], gc[3];
            const int path = x.eager ? ((mx >= 0 ? 1 : 0) | warp_yz) : -1;
            if (path <= 0) {
#pragma unroll
                for (int q3 = 0; q3 < 3; ++q3)
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        lout[q3].v[e] = (T)g[q3][e];
                        gc[q3].v[e] = (T)(s * g[q3][e]);
                    }
            } else {
                pml.apply_adj(a, i, mx, xt.u[q], xt.r[q], s, g, lout, gc);
            }
#pragma unroll
            for (int q3 = 0; q3 < 3; ++q3) {
                stv<T, V>(a.Dout[q3] + pbase + orow, lout[q3]);
                if (x.eager) stv<T, V>(a.Eout[q3] + pbase + orow, gc[q3]);
            }
        }
        __syncthreads();
        sc = sn;
        ph = phn;
    }
}

}  // namespace cev
