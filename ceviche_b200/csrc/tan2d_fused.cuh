// Fused full time step of the TANGENT states of a forward-mode sweep on 2-D TM grids, all B states in one launch.
//
// The batched tangent launches of step_v2.cuh move, per tangent cell and time step, 8 words in the H half-step and 4 in
// the D half-step (ncu: DRAM-bound at 5.6-5.9 TB/s, profiles/r2_ncu_config5_batched_tangents.json).  Here the two
// half-steps are ONE pass: dD, d(1/eps), dH (2) in; dH (2), dD out = 7 words (1/eps and the primal D are shared by the B
// states and come from L2).  Marching along x, the thread that has just produced dH' on row i can do the D half-step of
// row i with dH'_z of row i-1 carried in registers and dH'_x of the cell to the left from the neighbouring lane; the first
// lane of a warp and the first row of an x-chunk RECOMPUTE that one halo value from the inputs (one extra cell / row).
// A neighbouring CTA may still need the old values of cells this CTA has already updated, so the step is out of place:
// dD, dH and the H-side PML integrals are ping-ponged between the caller's arrays and plan-owned shadows (the D-side
// corner integral is touched by its own cell only and stays in place).
//
// 2-D TM in the plan's internal axes (x, z_logical = 1 plane, y_logical contiguous): E / D component 1, H components 0
// and 2 (MASK_TM); fdtd.py:74-144 with curl_E / curl_H of derivatives.py:16-30 where every term of a dead component is
// the literal zero the masked kernels pass, and every product and sum in the order of common.cuh's update_cell: results
// are bit-identical to the two-kernel path (asserted by the tests).
#pragma once
#include "step_v2.cuh"

namespace cev {

// update_cell (common.cuh) with the integrals read from `in` and written to `out` arrays; store = false: value only
// (halo recomputation)
template <typename T, typename AT>
__device__ __forceinline__ AT update_pp(AT old, AT curl, AT ua, AT ra, AT ub, AT rb, AT uc, AT scdt, const T* Icin, T* Icout,
                                        int64_t ic, const T* Isin, T* Isout, int64_t is, bool store, bool pml = true) {
    // off the PML of all three axes (u = 0, r = 1: m1 = 1, m2 = s, no integrals) the general formula below gives exactly
    // old + s curl: the same bits with two instructions (the vacuum path of the marching kernels)
    if (!pml) return muladd(scdt, curl, old);
    AT m1, m2;
    coef12<AT>(ua, ra, ub, rb, scdt, m1, m2);
    AT v = muladd(m1, old, mul_rn(m2, curl));
    if (ic >= 0) {
        const AT I = (AT)Icin[ic] + curl;
        if (store) Icout[ic] = (T)I;
        v = muladd(mul_rn(mul_rn(scdt, uc + uc), mul_rn(ra, rb)), I, v);
    }
    if (is >= 0) {
        const AT I = (AT)Isin[is] + old;
        if (store) Isout[is] = (T)I;
        v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), ua), ub), mul_rn(ra, rb)), I, v);
    }
    return v;
}

#ifndef CEV_TAN2D_MINB
#define CEV_TAN2D_MINB 4      // CTAs per SM the register budget is capped for (tuned: profiles/r2_tune_fused_tangent_occupancy.log)
#endif

template <typename T, typename AT, int V>
__global__ void __launch_bounds__(32 * V2_BY, CEV_TAN2D_MINB) k_tan2d_fused_batch(const StepArgs<T, AT>* table, int B, int64_t t_probe) {
    const StepArgs<T, AT>& a = batch_args<T, AT>(table, blockIdx.x % B, t_probe);
    const int bid = blockIdx.x / B;
    if (bid >= a.n_tiles) {       // probes of the previous step (E / D and H families) on the input state
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x, w = threadIdx.y;
    const int tz = bid % a.ntz, xc = bid / a.ntz;
    const int Nx = a.Nx, Nz = a.Nz;
    const int xs = xc * a.xchunk, xe = min(xs + a.xchunk, Nx);
    const int k0raw = ((tz * V2_BY + w) * 32 + lane) * V;
    const bool active = k0raw < Nz;
    const int k0 = active ? k0raw : 0;                         // inactive lanes shadow valid cells (loads only)
    const bool zlast = lane == 31 || k0 + V >= Nz;             // the +1 neighbour of the last cell is not in lane + 1
    const bool zfirst = lane == 0 || k0 == 0;                  // the -1 neighbour of the first cell is not in lane - 1
    const int kp = (k0 + V >= Nz) ? 0 : k0 + V;
    const int km = (k0 == 0) ? Nz - 1 : k0 - 1;

    const T* __restrict__ tD = a.Din[1];
    const T* __restrict__ tH0 = a.Hin[0];
    const T* __restrict__ tH2 = a.Hin[2];
    const T* __restrict__ mE = a.mE[1];
    const T* __restrict__ dmE = a.dmE[1];
    const T* __restrict__ Dp = a.Dp[1];
    const AT sH = -a.cdt, sD = a.cdt, inv = a.inv_dL;
    const int nH2 = a.nH[2], nD2 = a.nD[2];
    // axis 1 has extent 1: no PML there (u = 0, r = 1 from the tables)
    const AT uH1 = a.uH[1][0], rH1 = a.rH[1][0], uD1 = a.uD[1][0];

    // per-cell z tables (loop-invariant)
    AT uH2[V], rH2[V], uD2[V], rD2[V];
    int mH2[V], mD2[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
        uH2[e] = a.uH[2][k0 + e]; rH2[e] = a.rH[2][k0 + e];
        uD2[e] = a.uD[2][k0 + e]; rD2[e] = a.rD[2][k0 + e];
        mH2[e] = a.mapH[2][k0 + e]; mD2[e] = a.mapD[2][k0 + e];
    }
    const AT uH2m = a.uH[2][km], rH2m = a.rH[2][km];

    auto e_row = [&](int row, int k) {     // tangent E = mE dD + dmE D_primal at (row, k .. k+V-1)
        const int o = row * Nz + k;
        const Vec<T, V> d = ldv<T, V>(tD + o), m = ldv<T, V>(mE + o), dm = ldv<T, V>(dmE + o), dp = ldv<T, V>(Dp + o);
        struct R { AT e[V]; Vec<T, V> d; } r;
#pragma unroll
        for (int e = 0; e < V; ++e) r.e[e] = e_of<true, T, AT, V>(m, d, dm, dp, e);
        r.d = d;
        return r;
    };
    auto e_cell = [&](int row, int k) {    // ... at one cell
        const int o = row * Nz + k;
        return add_rn(mul_rn((AT)mE[o], (AT)tD[o]), mul_rn((AT)dmE[o], (AT)Dp[o]));
    };

    // ---- pre-roll: E of row xs (with its +1 halo cell) and dH'_z of row xs - 1 (recomputed from the inputs)
    AT Ecur[V + 1];
    Vec<T, V> dcur;
    {
        const auto r = e_row(xs, k0);
#pragma unroll
        for (int e = 0; e < V; ++e) Ecur[e] = r.e[e];
        dcur = r.d;
        const AT nb = __shfl_down_sync(FULL, Ecur[0], 1);
        Ecur[V] = zlast ? e_cell(xs, kp) : nb;
    }
    AT H2prev[V];
    {
        const int im = (xs == 0) ? Nx - 1 : xs - 1;
        const auto r = e_row(im, k0);
        const Vec<T, V> h2 = ldv<T, V>(tH2 + im * Nz + k0);
        const AT u0 = a.uH[0][im], r0 = a.rH[0][im];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT ce2 = curl2<AT>(Ecur[e], r.e[e], AT(0), AT(0), inv);
            const int64_t ic = mH2[e] >= 0 ? (int64_t)im * nH2 + mH2[e] : -1;
            H2prev[e] = (AT)(T)update_pp<T, AT>((AT)h2.v[e], ce2, u0, r0, uH1, rH1, uH2[e], sH, a.ICE[2], nullptr, ic, nullptr, nullptr, -1, false);
        }
    }

    for (int i = xs; i < xe; ++i) {
        const int ip = (i + 1 == Nx) ? 0 : i + 1;
        const int pf = a.pf_dist + 1;      // rows ahead (prefetch_planes = 1, the default: two)
        if ((lane & (128 / (int)(sizeof(T) * V) - 1)) == 0 && pf > 1 && i + pf < Nx) {      // one lane per 128-byte line into L2
            const int o2 = (i + pf) * Nz + k0;
            prefetch_l2(tD + o2); prefetch_l2(dmE + o2); prefetch_l2(mE + o2); prefetch_l2(Dp + o2);
            prefetch_l2(tH0 + o2 - Nz); prefetch_l2(tH2 + o2 - Nz);
        }
        // all loads of the row first
        const auto rn = e_row(ip, k0);
        const Vec<T, V> h0 = ldv<T, V>(tH0 + i * Nz + k0), h2 = ldv<T, V>(tH2 + i * Nz + k0);
        AT En[V + 1];
#pragma unroll
        for (int e = 0; e < V; ++e) En[e] = rn.e[e];
        {
            const AT nb = __shfl_down_sync(FULL, En[0], 1);
            En[V] = zlast ? e_cell(ip, kp) : nb;
        }
        const AT uH0 = a.uH[0][i], rH0 = a.rH[0][i], uD0 = a.uD[0][i], rD0 = a.rD[0][i];
        const int mxH = a.mapH[0][i], mxD = a.mapD[0][i];
        // H half-step (fdtd.py:80-97): dH'_x from -(dE/dz), dH'_z from +(dE/dx)
        AT H0n[V], H2n[V];
        Vec<T, V> o0, o2;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT ce0 = curl2<AT>(AT(0), AT(0), Ecur[e + 1], Ecur[e], inv);
            const AT ce2 = curl2<AT>(En[e], Ecur[e], AT(0), AT(0), inv);
            const int64_t ic0 = mxH >= 0 ? (int64_t)mxH * Nz + k0 + e : -1;
            const int64_t ic2 = mH2[e] >= 0 ? (int64_t)i * nH2 + mH2[e] : -1;
            const bool pH = mxH >= 0 || mH2[e] >= 0;
            H0n[e] = update_pp<T, AT>((AT)h0.v[e], ce0, uH1, rH1, uH2[e], rH2[e], uH0, sH, a.ICE[0], a.ICEout[0], ic0, nullptr, nullptr, -1, active, pH);
            H2n[e] = update_pp<T, AT>((AT)h2.v[e], ce2, uH0, rH0, uH1, rH1, uH2[e], sH, a.ICE[2], a.ICEout[2], ic2, nullptr, nullptr, -1, active, pH);
            o0.v[e] = (T)H0n[e];
            o2.v[e] = (T)H2n[e];
        }
        // dH'_x of the cell to the left: the neighbouring lane's, or recomputed from the inputs
        AT H0km = __shfl_up_sync(FULL, (AT)o0.v[V - 1], 1);
        if (zfirst) {
            const int o = i * Nz + km;
            const AT ekm = e_cell(i, km);
            const AT ce0 = curl2<AT>(AT(0), AT(0), Ecur[0], ekm, inv);
            const int64_t ic0 = mxH >= 0 ? (int64_t)mxH * Nz + km : -1;
            H0km = (AT)(T)update_pp<T, AT>((AT)tH0[o], ce0, uH1, rH1, uH2m, rH2m, uH0, sH, a.ICE[0], nullptr, ic0, nullptr, nullptr, -1, false);
        }
        // D half-step (fdtd.py:105-122) from the STORED (rounded) dH'
        Vec<T, V> od;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT hx = (AT)o0.v[e], hxm = e > 0 ? (AT)o0.v[(e + V - 1) % V] : H0km;
            const AT ch1 = curl2<AT>(hx, hxm, (AT)o2.v[e], H2prev[e], inv);
            const int64_t is = (mxD >= 0 && mD2[e] >= 0) ? (int64_t)mxD * nD2 + mD2[e] : -1;
            od.v[e] = (T)update_pp<T, AT>((AT)dcur.v[e], ch1, uD0, rD0, uD2[e], rD2[e], uD1, sD, nullptr, nullptr, -1, a.ID[1], a.ID[1], is, active,
                                          mxD >= 0 || mD2[e] >= 0);
        }
        if (active) {
            stv<T, V>(a.Hout[0] + i * Nz + k0, o0);
            stv<T, V>(a.Hout[2] + i * Nz + k0, o2);
            stv<T, V>(a.Dout[1] + i * Nz + k0, od);
        }
#pragma unroll
        for (int e = 0; e <= V; ++e) Ecur[e] = En[e];
#pragma unroll
        for (int e = 0; e < V; ++e) H2prev[e] = (AT)o2.v[e];
        dcur = rn.d;
    }
}

}  // namespace cev
