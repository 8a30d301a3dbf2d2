// Shared device-side definitions for the FDTD kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cev {

// Explicitly rounded arithmetic: no FMA contraction, so that (i) every kernel variant produces
// bit-identical results and (ii) the rounding sequence is the reference's numpy one (each product
// and each sum rounded separately, fdtd.py:95-97 / derivatives.py:18).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
// a*b + c: two roundings (the reference's numpy sequence) by default; -DCEV_FMA fuses them (experiment:
// fewer FP64 instructions / less power, results still within the parity tolerance but not the reference's
// rounding sequence)
__device__ __forceinline__ double muladd(double a, double b, double c) {
#ifdef CEV_FMA
    return __fma_rn(a, b, c);
#else
    return __dadd_rn(__dmul_rn(a, b), c);
#endif
}
__device__ __forceinline__ float muladd(float a, float b, float c) {
#ifdef CEV_FMA
    return __fmaf_rn(a, b, c);
#else
    return __fadd_rn(__fmul_rn(a, b), c);
#endif
}

// Sampling of probe point sets: one CTA per "slot" (a chunk of one probe's points) does a
// fixed-order reduction and writes ONE partial sum -> deterministic series, no atomics.
struct ProbeTable {
    int            n_slots;
    const int32_t* slot_field;   // CEV_FIELD_* + component
    const int64_t* slot_wbegin;  // offset into weight[]
    const int64_t* slot_ibegin;  // offset into idx[], <0 => dense
    const int64_t* slot_cell0;   // dense: first cell
    const int64_t* slot_n;       // points in the slot
    const int64_t* idx;
    const double*  weight;
};

// A launch of the marching kernels covers up to 6 boxes of cells (the whole grid; or the PML-free
// interior; or the six slabs of the PML shell), each tiled from its own origin.
struct Box {
    int x0, x1, y0, y1, z0, z1;   // cell ranges (z0 a multiple of the vector width)
    int cta0, ntz, nty;           // first CTA of the box, tiles along z and y
};
constexpr int MAX_BOXES = 6;

// Outside-in order of the x-chunks of a launch (tensor-map kernels): rank 0, 1, 2, 3 ... -> chunk first, last, second,
// last but one ...; the chunks in the x-PML (the slowest CTAs) start first, cheap interior chunks fill the tail.
__host__ __device__ __forceinline__ int v5_chunk_of_rank(int r, int nchunks) {
    return (r & 1) ? nchunks - 1 - (r >> 1) : (r >> 1);
}
__host__ __device__ __forceinline__ int v5_rank_of_chunk(int c, int nchunks) {
    return (2 * c <= nchunks - 1) ? 2 * c : 2 * (nchunks - 1 - c) + 1;
}

// Everything a half-step kernel needs.  Axes/components are in the plan's INTERNAL order
// (a cyclic relabelling of x,y,z chosen so that the last internal axis is the contiguous
// one with extent > 1; see cev_fdtd.cu).  T = storage type, AT = arithmetic type.
template <typename T, typename AT>
struct StepArgs {
    int Nx, Ny, Nz;
    int x0, x1;
    const T* Hin[3];
    T*       Hout[3];
    const T* Din[3];
    T*       Dout[3];
    const T* mE[3];
    T*       Eout[3];
    const T* Dhi[3];    // plane (Ny*Nz) standing for i = Nx   (wrap: plane 0; slab: halo buffer)
    const T* mEhi[3];
    const T* Hlo[3];    // plane standing for i = -1            (wrap: plane Nx-1; slab: halo buffer)
    T* ICE[3];
    T* IH[3];
    T* ICH[3];
    T* ID[3];
    T* ICEout[3];       // fused full-step kernel only (step_v4.cuh): the H-side integrals are ping-ponged
    T* IHout[3];
    const int* mapH[3];
    const int* mapD[3];
    int nH[3], nD[3];
    const AT* uH[3];    // u = sigma*dt/(2 eps0) per axis (H sampling / D sampling)
    const AT* rH[3];    // r = 1/(1+u)
    const AT* uD[3];
    const AT* rD[3];
    AT cdt;             // C_0 * dt
    AT inv_dL;
    // forward-mode tangent of E = mE*D w.r.t. eps_r: E = mE*D + dmE*Dp (D is then the TANGENT state,
    // Dp the primal D); all NULL for an ordinary step
    const T* dmE[3];
    const T* Dp[3];
    const T* dmEhi[3];
    const T* Dphi[3];
    const T* J[3];      // dense source per component (nullable)
    AT       Jscale[3];
    const double* Jwave[3];   // nullable device scalar overriding Jscale (waveform entry)
    // sparse sources J(t) = sum_s profile_s * waveform[t, s] (fdtd.py:125-127 for the callers' J = profile *
    // scalar(t), utils.py:328), injected by the D half-step itself: the points are pre-sorted by the CTA
    // that owns their cell; src_begin[bid] .. src_begin[bid+1] is CTA bid's run.  src_wave == NULL: none.
    const int*     src_begin;
    const int32_t* src_comp;
    const int32_t* src_id;
    const int32_t* src_cell;
    const double*  src_w;
    const double*  src_wave;
    // tiling + auxiliary probe CTAs appended to the grid
    // components known to be identically zero are skipped (2-D TM / TE runs): bits 0-2 = E/D component active,
    // bits 3-5 = H component active (internal order).  The caller guarantees the inactive ones are and stay zero.
    unsigned on;
    int n_tiles, ntz, nty, xchunk;
    int wz;                   // marching kernels: 1 = the warps of a CTA tile z (grids with a single row per plane)
    int xorder;               // tensor-map kernels: 1 = x-chunks dealt outside-in (v5_chunk_of_rank)
    // x-slab halo exchange through peer-mapped memory (tensor-map kernels; all NULL / 0 on a periodic grid):
    // the CTAs that produce the slab's boundary plane also store its y, z components into the neighbour's halo
    // buffer and bump the neighbour's arrival counter; the CTAs that read this slab's own halo plane first wait
    // until its counter has reached own_target (see step_v5.cuh)
    T* peer_out[2];
    unsigned long long* peer_flag;
    const unsigned long long* own_flag;
    unsigned long long own_target;
    int* halo_err;
    int n_boxes;
    Box box[MAX_BOXES];
    int pf_dist;              // L2 prefetch distance in x-planes (0 = off)
    ProbeTable pr;
    int        aux_slot0;     // first slot handled by the aux CTAs of this launch
    int64_t    t_probe;       // row of partials[] they write
    double*    partials;
};

template <typename T, typename AT>
__device__ __forceinline__ AT probe_value(const StepArgs<T, AT>& a, int field, int64_t cell) {
    const int c = field % 3;
    if (field < 3) {
        AT e = mul_rn((AT)a.mE[c][cell], (AT)a.Din[c][cell]);
        if (a.dmE[c]) e = add_rn(e, mul_rn((AT)a.dmE[c][cell], (AT)a.Dp[c][cell]));
        return e;
    }
    if (field < 6) return (AT)a.Din[c][cell];
    return (AT)a.Hin[c][cell];
}

// One CTA reduces one slot with its first PROBE_THREADS threads, whatever the CTA shape of the
// kernel it rides on: the summation order (hence the series, bit for bit) is the same everywhere.
constexpr int PROBE_THREADS = 128;
template <typename T, typename AT>
__device__ void probe_block(const StepArgs<T, AT>& a, int slot) {
    __shared__ double red[PROBE_THREADS];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid >= PROBE_THREADS) return;       // CTAs have >= 128 threads; the rest sit out (no later barrier for them)
    const int nth = PROBE_THREADS;
    const int field = a.pr.slot_field[slot];
    const int64_t n = a.pr.slot_n[slot];
    const int64_t wb = a.pr.slot_wbegin[slot];
    const int64_t ib = a.pr.slot_ibegin[slot];
    const int64_t c0 = a.pr.slot_cell0[slot];
    double acc = 0.0;
    for (int64_t q = tid; q < n; q += nth) {
        const int64_t cell = (ib < 0) ? (c0 + q) : a.pr.idx[ib + q];
        acc += (double)probe_value<T, AT>(a, field, cell) * a.pr.weight[wb + q];
    }
    red[tid] = acc;
    // named barrier over exactly the participating threads
    asm volatile("bar.sync 1, %0;" ::"n"(PROBE_THREADS));
    for (int s = nth >> 1; s > 0; s >>= 1) {
        if (tid < s) red[tid] += red[tid + s];
        asm volatile("bar.sync 1, %0;" ::"n"(PROBE_THREADS));
    }
    if (tid == 0) a.partials[a.t_probe * a.pr.n_slots + slot] = red[0];
}

// Sparse J injection, D += J with J = sum_s profile_s * waveform[t, s] (fdtd.py:125-127 for the callers'
// J = profile * scalar(t), utils.py:328).  The points q0 <= q < q1 are sorted by (component, cell) with the source
// order kept inside a run of equal keys (the host sorts once, stably): the thread that meets the first point of a run
// sums the run's terms in that order and adds the sum to the cell ONCE -- the reference's own order (the J arrays
// are summed, then D += J), no atomics: overlapping sources give the same bits whatever the launch geometry.
template <typename T, typename AT, typename CELL>
__device__ __forceinline__ void inject_points(int64_t q0, int64_t q1, int tid, int nthreads, const int32_t* comp,
                                              const int32_t* id, const CELL* cell, const double* w, const double* wave,
                                              T* D0, T* D1, T* D2, int64_t cell_lo = 0, int64_t cell_hi = INT64_MAX) {
    for (int64_t q = q0 + tid; q < q1; q += nthreads) {
        const int c = comp[q];
        const CELL o = cell[q];
        if (q > q0 && comp[q - 1] == c && cell[q - 1] == o) continue;       // not the first point of its run
        if ((int64_t)o < cell_lo || (int64_t)o >= cell_hi) continue;        // (baseline path: only the planes this launch updated)
        double j = __dmul_rn(w[q], wave[id[q]]);
        for (int64_t r = q + 1; r < q1 && comp[r] == c && cell[r] == o; ++r) j = __dadd_rn(j, __dmul_rn(w[r], wave[id[r]]));
        T* D = c == 0 ? D0 : (c == 1 ? D1 : D2);
        D[o] = (T)add_rn((AT)D[o], (AT)j);
    }
}

// (a1-a0)/dL - (b1-b0)/dL, both quotients rounded before the subtraction (derivatives.py:16-30)
template <typename AT>
__device__ __forceinline__ AT curl2(AT a1, AT a0, AT b1, AT b0, AT inv) {
    return muladd(a1 - a0, inv, -mul_rn(b1 - b0, inv));
}

// Coefficients of fdtd.py:272-311 rewritten division-free from the per-axis tables
//   u = sigma*dt/(2 eps0), r = 1/(1+u):   m0*dt = (1+ua)(1+ub),  1/m0 = dt*ra*rb
//   m1 = (1-ua-ub-ua*ub) ra rb = (2 - (1+ua)(1+ub)) ra rb = 2 ra rb - 1
//   m2 = s*C0*dt ra rb,  m3 = s*C0*dt*2uc ra rb,  m4 = -4 ua ub ra rb
// (s = -1 for H, +1 for D; (ua,ub) = the two other axes, uc = the component's own axis).
// Off the PML (u = 0, r = 1) this gives m1 = 1 and m2 = s*C0*dt exactly.
template <typename AT>
__device__ __forceinline__ void coef12(AT ua, AT ra, AT ub, AT rb, AT scdt, AT& m1, AT& m2) {
    const AT rr = mul_rn(ra, rb);
    m1 = add_rn(rr + rr, AT(-1));
    m2 = mul_rn(scdt, rr);
    (void)ua;
    (void)ub;
}

// One field component of one half-step (fdtd.py:85-97 for H, :110-122 for D):
//   I_curl += curl ; I_self += old ; new = m1*old + m2*curl + m3*I_curl + m4*I_self
// The integral arrays exist only where their coefficient is non-zero (index < 0: skip); m3 / m4
// are only evaluated there.
template <typename T, typename AT>
__device__ __forceinline__ AT update_cell(AT old, AT curl, AT m1, AT m2, AT ua, AT ra, AT ub, AT rb, AT uc, AT scdt,
                                          T* Icurl, int64_t icurl, T* Iself, int64_t iself) {
    AT v = muladd(m1, old, mul_rn(m2, curl));
    if (icurl >= 0) {
        const AT I = (AT)Icurl[icurl] + curl;
        Icurl[icurl] = (T)I;
        v = muladd(mul_rn(mul_rn(scdt, uc + uc), mul_rn(ra, rb)), I, v);
    }
    if (iself >= 0) {
        const AT I = (AT)Iself[iself] + old;
        Iself[iself] = (T)I;
        v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), ua), ub), mul_rn(ra, rb)), I, v);
    }
    return v;
}

template <typename T, typename AT>
__device__ __forceinline__ AT update_component(AT old, AT curl, AT ua, AT ra, AT ub, AT rb, AT uc, AT scdt,
                                               T* Icurl, int64_t icurl, T* Iself, int64_t iself) {
    AT m1, m2;
    coef12<AT>(ua, ra, ub, rb, scdt, m1, m2);
    return update_cell<T, AT>(old, curl, m1, m2, ua, ra, ub, rb, uc, scdt, Icurl, icurl, Iself, iself);
}

}  // namespace cev
