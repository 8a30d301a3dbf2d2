// Tensor-map TMA half-step kernels for sm_100a (kernel_variant = 6; the default for large 3-D grids).
//
// Same arithmetic as every other variant (bit-identical results, asserted by the tests), but the planes are
// moved by `cp.async.bulk.tensor.3d` box copies (SASS UTMALDG) described by CUtensorMap descriptors: ONE
// instruction per array per x-plane brings a CTA's tile of BY rows x BZ cells -- with its +-1 halo vector along z
// inside the same box and the +-1 halo row along y as a second one-row box -- into a ring of shared-memory
// stages guarded by mbarriers.  Out-of-range box coordinates are zero-filled by the TMA unit, so the periodic
// wrap needs no special-case copy: the halo ROW is its own box at row (y0 + BY) mod Ny, the wrapped halo CELL
// along z is one scalar load by the edge lane, and the wrapped PLANE along x is the same descriptor at x = 0
// (or the descriptor of the neighbour's halo buffer on an x-slab).
//
// What this buys over the register-marching kernels (step_v2.cuh): no address arithmetic and no registers for
// the 13 loads per plane (the v2 interior loop spends ~50 % of its instructions on 64-bit addressing and spills
// at its occupancy cap), and NS-2 planes in flight per CTA whatever the arithmetic of the current plane costs,
// which is what the PML shell needs.  The PML update itself is restated with everything that does not depend on
// the x-plane hoisted out of the marching loop (PmlLean below): same rounding sequence, ~3x fewer instructions.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "step_v2.cuh"
#include "step_v3.cuh"

namespace cev {

template <typename T, int V, int BY>
struct V5Layout {
    static constexpr int BZ = 32 * V;                 // cells of a tile along z: one 16-byte vector per lane
    static constexpr int ROWP = BZ + 2 * V;           // smem row pitch = box extent along z (544 bytes for both dtypes):
                                                      // BY * ROWP * sizeof(T) is a multiple of 128, so the halo row
                                                      // that follows the BY own rows is a legal TMA destination
    static constexpr int BLK = ((BY + 1) * ROWP * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    static constexpr int PLN = BY * BZ;               // arrays staged without halos
    static constexpr int H_STAGE = 6 * BLK + 3 * PLN; // H half-step: D[3], mE[3] with halos; H[3] without
    static constexpr int D_STAGE = 3 * BLK + 3 * PLN; // D half-step: H[3] with halos; D[3] without
    static constexpr uint32_t BOX_MAIN = BY * ROWP * sizeof(T), BOX_ROW = ROWP * sizeof(T), BOX_PLN = PLN * sizeof(T);
    static constexpr size_t h_bytes(int ns) { return (size_t)ns * H_STAGE * sizeof(T) + (size_t)ns * sizeof(uint64_t); }
    static constexpr size_t d_bytes(int ns) { return (size_t)ns * D_STAGE * sizeof(T) + (size_t)ns * sizeof(uint64_t); }
};

// Descriptors of one launch.  "main" boxes are (ROWP, BY, 1) cells, "row" boxes (ROWP, 1, 1), "pln" boxes
// (BZ, BY, 1); the hi / lo descriptors stand for the plane beyond the last / before the first x-plane of the
// array (the array itself on a periodic grid, the halo buffer on an x-slab) and are addressed at x = x_hi / x_lo.
struct V5MapsH {
    CUtensorMap D[3], M[3];      // main boxes of D and 1/eps
    CUtensorMap H[3];            // pln boxes of H
    CUtensorMap Drow[2], Mrow[2];   // row boxes: components x, z (the ones differenced along y)
    CUtensorMap Dhi[2], Mhi[2];     // main boxes of the x+1 plane beyond the array: components y, z
    int x_hi;
};
struct V5MapsD {
    CUtensorMap H[3];            // main boxes of H
    CUtensorMap D[3];            // pln boxes of D
    CUtensorMap Hrow[2];         // row boxes: components x, z
    CUtensorMap Hlo[2];          // main boxes of the x-1 plane before the array: components y, z
    int x_lo;
};

__device__ __forceinline__ void tma_box_3d(void* dst, const CUtensorMap* map, int cz, int cy, int cx, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(cz), "r"(cy), "r"(cx), "r"(smem_u32(bar))
        : "memory");
}

// ---- x-slab halos over peer-mapped memory -----------------------------------------------------------------
// The neighbour GPU writes this slab's halo plane straight into its memory (NVLink stores from inside the
// neighbour's half-step kernel) and then bumps an arrival counter there with a system-scope release.  A CTA about
// to fetch the halo plane spins on the counter (acquire) and then orders the TMA unit's reads after it.  The wait
// is bounded: a neighbour that never arrives sets *err instead of hanging the GPU.
__device__ __forceinline__ void halo_wait(const unsigned long long* flag, unsigned long long target, int* err) {
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= target) break;
        if (clock64() - t0 > (1LL << 33)) {          // ~4 s at 1.9 GHz
            *err = 1;
            break;
        }
        __nanosleep(100);
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}
// after every thread of the CTA has stored its part of the boundary plane to the neighbour (and met at a barrier)
__device__ __forceinline__ void halo_signal(unsigned long long* peer_flag) {
    __threadfence_system();
    atomicAdd_system(peer_flag, 1ULL);
}

template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ldv_cg(const T* p) {        // 16-byte load that bypasses L1
    const int4 raw = __ldcg(reinterpret_cast<const int4*>(p));
    Vec<T, V> out;
    static_assert(sizeof(out) == sizeof(raw), "Vec is one 16-byte vector");
    memcpy(&out, &raw, sizeof(raw));
    return out;
}

// bit 1: some lane of the warp is in the y-PML, bit 2: some lane has a cell in the z-PML (loop-invariant votes)
__device__ __forceinline__ int v5_warp_flags(bool fy, bool fz) {
    return (__any_sync(0xffffffffu, fy) ? 2 : 0) | (__any_sync(0xffffffffu, fz) ? 4 : 0);
}

// one lane of a converged warp (SASS ELECT): the TMA instructions take uniform operands, and a branch the
// compiler knows to be taken by a single lane lets it issue them without a per-lane fallback loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------------------
// PML update of one thread (V cells of one row) for one x-plane, fdtd.py:85-97 / :110-122 with the
// coefficients of fdtd.py:272-311 in the division-free form of common.cuh (coef12 / update_cell): the SAME
// products and sums in the same order, so results are bit-identical to step_v1.cuh.  Everything that depends
// only on (j, k) -- table entries, r_y r_z, the corner coefficient of the x component, compact offsets -- is
// computed once before the marching loop.
template <typename T, typename AT, int V, bool IS_H>
struct PmlLean {
    int my, mz[V];
    bool fy, fz[V], yz;
    AT uy, ry, su2y;
    AT uz[V], rz[V], su2z[V], rrx[V], m4x[V];
    AT m1y, m2y, m1zc[V], m2zc[V];             // (2 r - 1, s r) of the single-axis paths below
    int o_ic1, o_ic2, o_is0, o_is1, o_is2;     // plane-independent parts of the compact offsets
    int n1, n2, Ny, Nz, orow;
    // old integrals of the current plane (loaded before the thread blocks on the TMA barrier)
    Vec<T, V> I0, I1, S2;
    T I2[V], S0[V], S1[V];

#define CEV_TAB(a, name, ax) (IS_H ? (a).name##H[ax] : (a).name##D[ax])
    __device__ __forceinline__ void init(const StepArgs<T, AT>& a, int j, int k0, AT s) {
        n1 = IS_H ? a.nH[1] : a.nD[1];
        n2 = IS_H ? a.nH[2] : a.nD[2];
        Ny = a.Ny;
        Nz = a.Nz;
        orow = j * a.Nz + k0;
        my = CEV_TAB(a, map, 1)[j];
        uy = CEV_TAB(a, u, 1)[j];
        ry = CEV_TAB(a, r, 1)[j];
        fy = my >= 0;
        yz = fy;
        su2y = mul_rn(s, uy + uy);
        m1y = add_rn(ry + ry, AT(-1));
        m2y = mul_rn(s, ry);
        const AT n4uy = mul_rn(AT(-4), uy);
        int mz0 = 0;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            mz[e] = CEV_TAB(a, map, 2)[k0 + e];
            uz[e] = CEV_TAB(a, u, 2)[k0 + e];
            rz[e] = CEV_TAB(a, r, 2)[k0 + e];
            fz[e] = mz[e] >= 0;
            yz |= fz[e];
            su2z[e] = mul_rn(s, uz[e] + uz[e]);
            rrx[e] = mul_rn(ry, rz[e]);
            m4x[e] = mul_rn(mul_rn(n4uy, uz[e]), rrx[e]);
            m1zc[e] = add_rn(rz[e] + rz[e], AT(-1));
            m2zc[e] = mul_rn(s, rz[e]);
            if (e == 0) mz0 = mz[0];
        }
        (void)mz0;
        o_ic1 = (fy ? my : 0) * a.Nz + k0;          // Icurl_y (Nx, n1, Nz): + i * n1 * Nz
        o_ic2 = j * n2;                             // Icurl_z (Nx, Ny, n2): + i * Ny * n2 + mz[e]
        o_is0 = (fy ? my : 0) * n2;                 // Iself_x (Nx, n1, n2): + i * n1 * n2 + mz[e]
        o_is1 = j * n2;                             // Iself_y (nx', Ny, n2): + mx * Ny * n2 + mz[e]
        o_is2 = (fy ? my : 0) * a.Nz + k0;          // Iself_z (nx', n1, Nz): + mx * n1 * Nz
    }

    __device__ __forceinline__ void load(const StepArgs<T, AT>& a, int i, int mx) {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        T* const* Is = IS_H ? a.IH : a.ID;
        const bool fx = mx >= 0;
        if (fx) I0 = ldv<T, V>(Ic[0] + mx * (Ny * Nz) + orow);                 // Icurl_x (nx', Ny, Nz)
        if (fy) I1 = ldv<T, V>(Ic[1] + i * (n1 * Nz) + o_ic1);
        if (fx && fy) S2 = ldv<T, V>(Is[2] + mx * (n1 * Nz) + o_is2);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (fz[e]) {
                I2[e] = Ic[2][i * (Ny * n2) + o_ic2 + mz[e]];
                if (fy) S0[e] = Is[0][i * (n1 * n2) + o_is0 + mz[e]];
                if (fx) S1[e] = Is[1][mx * (Ny * n2) + o_is1 + mz[e]];
            }
        }
    }

    // old[c].v[e], curl[c][e] -> out[c].v[e]; the new integrals go back to global memory
    __device__ __forceinline__ void apply(const StepArgs<T, AT>& a, int i, int mx, AT ux, AT rx, AT s, const Vec<T, V>* old,
                                          const AT (*curl)[V], Vec<T, V>* out) {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        T* const* Is = IS_H ? a.IH : a.ID;
        const bool fx = mx >= 0;
        const AT su2x = mul_rn(s, ux + ux);
        const AT n4ux = mul_rn(AT(-4), ux);
        // z component: (a, b) = (x, y): the same coefficients for the V cells
        const AT rrz = mul_rn(rx, ry);
        const AT m1z = add_rn(rrz + rrz, AT(-1)), m2z = mul_rn(s, rrz);
        const AT m4z = mul_rn(mul_rn(n4ux, uy), rrz);
        Vec<T, V> n0, n1v, s2v;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            // x: (a, b) = (y, z), own axis x.  Icurl_x on the x-slab, Iself_x where the y- and z-PML overlap
            {
                const AT rr = rrx[e];
                const AT m1 = add_rn(rr + rr, AT(-1)), m2 = mul_rn(s, rr);
                AT v = muladd(m1, (AT)old[0].v[e], mul_rn(m2, curl[0][e]));
                if (fx) {
                    const AT I = (AT)I0.v[e] + curl[0][e];
                    n0.v[e] = (T)I;
                    v = muladd(mul_rn(su2x, rr), I, v);
                }
                if (fy && fz[e]) {
                    const AT I = (AT)S0[e] + (AT)old[0].v[e];
                    Is[0][i * (n1 * n2) + o_is0 + mz[e]] = (T)I;
                    v = muladd(m4x[e], I, v);
                }
                out[0].v[e] = (T)v;
            }
            // y: (a, b) = (x, z), own axis y.  Icurl_y on the y-slab, Iself_y where the x- and z-PML overlap
            {
                const AT rr = mul_rn(rx, rz[e]);
                const AT m1 = add_rn(rr + rr, AT(-1)), m2 = mul_rn(s, rr);
                AT v = muladd(m1, (AT)old[1].v[e], mul_rn(m2, curl[1][e]));
                if (fy) {
                    const AT I = (AT)I1.v[e] + curl[1][e];
                    n1v.v[e] = (T)I;
                    v = muladd(mul_rn(su2y, rr), I, v);
                }
                if (fx && fz[e]) {
                    const AT I = (AT)S1[e] + (AT)old[1].v[e];
                    Is[1][mx * (Ny * n2) + o_is1 + mz[e]] = (T)I;
                    v = muladd(mul_rn(mul_rn(n4ux, uz[e]), rr), I, v);
                }
                out[1].v[e] = (T)v;
            }
            // z: (a, b) = (x, y), own axis z.  Icurl_z on the z-slab, Iself_z where the x- and y-PML overlap
            {
                AT v = muladd(m1z, (AT)old[2].v[e], mul_rn(m2z, curl[2][e]));
                if (fz[e]) {
                    const AT I = (AT)I2[e] + curl[2][e];
                    Ic[2][i * (Ny * n2) + o_ic2 + mz[e]] = (T)I;
                    v = muladd(mul_rn(su2z[e], rrz), I, v);
                }
                if (fx && fy) {
                    const AT I = (AT)S2.v[e] + (AT)old[2].v[e];
                    s2v.v[e] = (T)I;
                    v = muladd(m4z, I, v);
                }
                out[2].v[e] = (T)v;
            }
        }
        if (fx) stv<T, V>(Ic[0] + mx * (Ny * Nz) + orow, n0);
        if (fy) stv<T, V>(Ic[1] + i * (n1 * Nz) + o_ic1, n1v);
        if (fx && fy) stv<T, V>(Is[2] + mx * (n1 * Nz) + o_is2, s2v);
    }

    // Transposed update of the reverse sweep (adjoint_v5.cuh).  g[c][e] = cotangent of the component's NEW value ->
    // lout = cotangent of its OLD value (m1 g + gIself), gc = cotangent of its curl (m2 g + gIcurl); the cotangents of
    // the integrals (the same compact arrays, their old values fetched by load()) gain m3 g / m4 g.  Plain arithmetic:
    // the adjoint has a tolerance to meet, not the reference's rounding sequence, so FMA contraction is welcome.
    __device__ __forceinline__ void apply_adj(const StepArgs<T, AT>& a, int i, int mx, AT ux, AT rx, AT s, const AT (*g)[V],
                                              Vec<T, V>* lout, Vec<T, V>* gc) {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        T* const* Is = IS_H ? a.IH : a.ID;
        const bool fx = mx >= 0;
        const AT su2x = s * (ux + ux);
        const AT n4ux = AT(-4) * ux;
        const AT rrz = rx * ry;
        const AT m1z = rrz + rrz - AT(1), m2z = s * rrz, m4z = n4ux * uy * rrz;
        Vec<T, V> n0, n1v, s2v;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            {   // x: (a, b) = (y, z)
                const AT rr = rrx[e], gg = g[0][e];
                AT l = (rr + rr - AT(1)) * gg, c = s * rr * gg;
                if (fx) {
                    const AT I = (AT)I0.v[e] + su2x * rr * gg;
                    n0.v[e] = (T)I;
                    c += I;
                }
                if (fy && fz[e]) {
                    const AT I = (AT)S0[e] + m4x[e] * gg;
                    Is[0][i * (n1 * n2) + o_is0 + mz[e]] = (T)I;
                    l += I;
                }
                lout[0].v[e] = (T)l;
                gc[0].v[e] = (T)c;
            }
            {   // y: (a, b) = (x, z)
                const AT rr = rx * rz[e], gg = g[1][e];
                AT l = (rr + rr - AT(1)) * gg, c = s * rr * gg;
                if (fy) {
                    const AT I = (AT)I1.v[e] + su2y * rr * gg;
                    n1v.v[e] = (T)I;
                    c += I;
                }
                if (fx && fz[e]) {
                    const AT I = (AT)S1[e] + n4ux * uz[e] * rr * gg;
                    Is[1][mx * (Ny * n2) + o_is1 + mz[e]] = (T)I;
                    l += I;
                }
                lout[1].v[e] = (T)l;
                gc[1].v[e] = (T)c;
            }
            {   // z: (a, b) = (x, y)
                const AT gg = g[2][e];
                AT l = m1z * gg, c = m2z * gg;
                if (fz[e]) {
                    const AT I = (AT)I2[e] + su2z[e] * rrz * gg;
                    Ic[2][i * (Ny * n2) + o_ic2 + mz[e]] = (T)I;
                    c += I;
                }
                if (fx && fy) {
                    const AT I = (AT)S2.v[e] + m4z * gg;
                    s2v.v[e] = (T)I;
                    l += I;
                }
                lout[2].v[e] = (T)l;
                gc[2].v[e] = (T)c;
            }
        }
        if (fx) stv<T, V>(Ic[0] + mx * (Ny * Nz) + orow, n0);
        if (fy) stv<T, V>(Ic[1] + i * (n1 * Nz) + o_ic1, n1v);
        if (fx && fy) stv<T, V>(Is[2] + mx * (n1 * Nz) + o_is2, s2v);
    }

    __device__ __forceinline__ bool fz_any() const {
        bool f = false;
#pragma unroll
        for (int e = 0; e < V; ++e) f |= fz[e];
        return f;
    }

    // ---- single-axis paths.  Most PML cells lie in the PML of ONE axis (the faces of the shell), where the other two
    // axes have u = 0, r = 1 exactly and the general update above collapses: products with 1 and sums with 0 are
    // exact, so the expressions below are the general ones with those factors dropped -- bit for bit the same
    // values, a third of the instructions.  The caller picks a path per WARP (votes over the lanes' flags); lanes
    // of the warp that are off the PML take it with u = 0, r = 1 and get the vacuum update.
    // x only: every lane has the same coefficients (fx holds for the whole plane)
    __device__ __forceinline__ void apply_x(const StepArgs<T, AT>& a, int mx, AT ux, AT rx, AT s, const Vec<T, V>* old,
                                            const AT (*curl)[V], Vec<T, V>* out) {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        const AT su2x = mul_rn(s, ux + ux);
        const AT m1 = add_rn(rx + rx, AT(-1)), m2 = mul_rn(s, rx);
        Vec<T, V> n0;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT I = (AT)I0.v[e] + curl[0][e];
            n0.v[e] = (T)I;
            out[0].v[e] = (T)muladd(su2x, I, muladd(s, curl[0][e], (AT)old[0].v[e]));
            out[1].v[e] = (T)muladd(m1, (AT)old[1].v[e], mul_rn(m2, curl[1][e]));
            out[2].v[e] = (T)muladd(m1, (AT)old[2].v[e], mul_rn(m2, curl[2][e]));
        }
        stv<T, V>(Ic[0] + mx * (Ny * Nz) + orow, n0);
    }
    // y only (this thread's row may or may not be in the y-PML: fy)
    __device__ __forceinline__ void apply_y(const StepArgs<T, AT>& a, int i, AT s, const Vec<T, V>* old, const AT (*curl)[V],
                                            Vec<T, V>* out) {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        Vec<T, V> n1v;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            out[0].v[e] = (T)muladd(m1y, (AT)old[0].v[e], mul_rn(m2y, curl[0][e]));
            AT v = muladd(s, curl[1][e], (AT)old[1].v[e]);
            if (fy) {
                const AT I = (AT)I1.v[e] + curl[1][e];
                n1v.v[e] = (T)I;
                v = muladd(su2y, I, v);
            }
            out[1].v[e] = (T)v;
            out[2].v[e] = (T)muladd(m1y, (AT)old[2].v[e], mul_rn(m2y, curl[2][e]));
        }
        if (fy) stv<T, V>(Ic[1] + i * (n1 * Nz) + o_ic1, n1v);
    }
    // z only (each of this thread's cells may or may not be in the z-PML: fz[e])
    __device__ __forceinline__ void apply_z(const StepArgs<T, AT>& a, int i, AT s, const Vec<T, V>* old, const AT (*curl)[V],
                                            Vec<T, V>* out) {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            out[0].v[e] = (T)muladd(m1zc[e], (AT)old[0].v[e], mul_rn(m2zc[e], curl[0][e]));
            out[1].v[e] = (T)muladd(m1zc[e], (AT)old[1].v[e], mul_rn(m2zc[e], curl[1][e]));
            AT v = muladd(s, curl[2][e], (AT)old[2].v[e]);
            if (fz[e]) {
                const AT I = (AT)I2[e] + curl[2][e];
                Ic[2][i * (Ny * n2) + o_ic2 + mz[e]] = (T)I;
                v = muladd(su2z[e], I, v);
            }
            out[2].v[e] = (T)v;
        }
    }

    // pull the curl integrals of a later plane of this thread into L2 (mxp: that plane's compact x index)
    __device__ __forceinline__ void prefetch(const StepArgs<T, AT>& a, int ip, int mxp, bool line_lane) const {
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        if (mxp >= 0 && line_lane) prefetch_l2(Ic[0] + mxp * (Ny * Nz) + orow);
        if (fy && line_lane) prefetch_l2(Ic[1] + ip * (n1 * Nz) + o_ic1);
        if (fz[0]) prefetch_l2(Ic[2] + ip * (Ny * n2) + o_ic2 + mz[0]);
        else if (fz[V - 1]) prefetch_l2(Ic[2] + ip * (Ny * n2) + o_ic2 + mz[V - 1]);
    }
#undef CEV_TAB
};

// Per-plane PML table entries of a CTA's x-chunk (compact index, u, r of the x axis) staged in shared memory by
// the prologue: the marching loop then reads them with LDS latency instead of one dependent global load per plane.
constexpr int V5_MAXCH = 64;      // longest x-chunk (the host clamps)
constexpr int V5_PF = 3;          // L2 prefetch distance of the PML integrals, in planes
template <typename AT>
struct V5XTab {
    int mx[V5_MAXCH + V5_PF + 1];
    AT u[V5_MAXCH], r[V5_MAXCH];
};
template <typename AT>
__device__ __forceinline__ void v5_fill_xtab(V5XTab<AT>& t, const int* map, const AT* u, const AT* r, int xs, int xe, int x1,
                                             int tid, int nthreads) {
    for (int q = tid; q < xe - xs + V5_PF + 1; q += nthreads) {
        t.mx[q] = xs + q < x1 ? map[xs + q] : -1;
        if (xs + q < xe) {
            t.u[q] = u[xs + q];
            t.r[q] = r[xs + q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Which tile / x-chunk a CTA serves and which cells of it a thread owns.  A warp covers 8 vectors along z times
// 4 rows (not one whole row): a warp takes the general PML path if ANY of its lanes is in the PML, and the z-PML
// (20 cells at each end of a row at the BASELINE configs) then costs a quarter of the warps of the two end tiles
// instead of all of them (with one row per warp EVERY warp of a 256-cell fp32 row holds some z-PML lanes).
// x-chunks are dealt outside-in (first, last, second, ...): the chunks in the x-PML, whose CTAs run longest, start
// first and the cheap interior chunks fill the tail of the launch.  The host's source tiling uses the same order.
struct V5Tile {
    int xs, xe, y0, z0, row, vec;
};
template <int V, int BY, typename ARGS>
__device__ __forceinline__ V5Tile v5_locate(const ARGS& a, int bid, int lane, int w) {
    V5Tile t;
    const int tz = bid % a.ntz;
    const int rest = bid / a.ntz;
    const int ty = rest % a.nty;
    const int nchunks = (a.x1 - a.x0 + a.xchunk - 1) / a.xchunk;
    const int xc = a.xorder ? v5_chunk_of_rank(rest / a.nty, nchunks) : rest / a.nty;
    t.xs = a.x0 + xc * a.xchunk;
    t.xe = min(t.xs + a.xchunk, a.x1);
    t.y0 = ty * BY;
    t.z0 = tz * 32 * V;
    t.row = (w >> 2) * 4 + (lane >> 3);              // 4 z-segments of 8 vectors per row group of 4 rows
    t.vec = (w & 3) * 8 + (lane & 7);
    return t;
}

// ---------------------------------------------------------------------------------------------------------
// H half-step.  CTA = BY warps on a tile of BY rows x BZ cells (thread -> cells: v5_locate); one launch box
// (the whole y-z plane of x-planes [a.x0, a.x1)), x-chunks of a.xchunk planes.
template <typename T, typename AT, int V, int BY, int NS>
__global__ void __launch_bounds__(32 * BY) k_step_H_v5(const StepArgs<T, AT> a, const __grid_constant__ V5MapsH maps) {
    using L = V5Layout<T, V, BY>;
    constexpr int BZ = L::BZ, ROWP = L::ROWP, BLK = L::BLK, HOFF = 6 * L::BLK;
    extern __shared__ __align__(128) unsigned char v5_smem[];
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    T* const stage0 = reinterpret_cast<T*>(v5_smem);
    uint64_t* const full = reinterpret_cast<uint64_t*>(v5_smem + (size_t)NS * L::H_STAGE * sizeof(T));
    const int lane = threadIdx.x, w = threadIdx.y;
    const V5Tile t = v5_locate<V, BY>(a, bid, lane, w);
    const int xs = t.xs, xe = t.xe, y0 = t.y0, z0 = t.z0;
    const int j = y0 + t.row;                        // (Ny is a multiple of BY: the host checks)
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

    // plane p -> stage (p - xs) % NS.  p < xe: a "current" plane (13 boxes); p == xe: only the x+1 neighbour of the
    // chunk's last plane (own rows of D_y, D_z, mE_y, mE_z: 4 boxes).  The boxes of a plane are dealt round-robin to
    // the BY warps (box q -> warp q % BY) so that no warp carries the whole issue cost; the elected lane of every
    // warp arms the stage's barrier with the bytes of its own boxes.
    const int yrow = (y0 + BY >= a.Ny) ? 0 : y0 + BY;
    auto box_cur = [&](int q, T* st, int p, uint64_t* bar) {
        if (q < 9) {
            const int c = q / 3, t = q % 3;
            if (t == 0) tma_box_3d(st + c * BLK, &maps.D[c], z0, y0, p, bar);
            else if (t == 1) tma_box_3d(st + (3 + c) * BLK, &maps.M[c], z0, y0, p, bar);
            else tma_box_3d(st + HOFF + c * L::PLN, &maps.H[c], z0, y0, p, bar);
        } else {
            const int r = (q - 9) / 2;                     // 0: component x, 1: component z
            const int blk = (q - 9) % 2 == 0 ? 2 * r : 3 + 2 * r;
            tma_box_3d(st + blk * BLK + BY * ROWP, (q - 9) % 2 == 0 ? &maps.Drow[r] : &maps.Mrow[r], z0, yrow, p, bar);
        }
    };
    auto box_nxt = [&](int q, T* st, int p, uint64_t* bar) {
        const bool hi = p >= a.Nx;
        if (hi && a.own_flag) halo_wait(a.own_flag, a.own_target, a.halo_err);    // x-slab: the right neighbour's plane 0
        const int px = hi ? maps.x_hi : p;
        const int c = 1 + q % 2;                           // q = 0..3: D_y, D_z, mE_y, mE_z
        if (q < 2) tma_box_3d(st + c * BLK, hi ? &maps.Dhi[c - 1] : &maps.D[c], z0, y0, px, bar);
        else tma_box_3d(st + (3 + c) * BLK, hi ? &maps.Mhi[c - 1] : &maps.M[c], z0, y0, px, bar);
    };
    constexpr uint32_t BM = L::BOX_MAIN, BR = L::BOX_ROW, BP = L::BOX_PLN;
    auto issue = [&](int p, int s) {                       // called by every warp, converged
        if (!elect_one()) return;
        uint64_t* bar = &full[s];
        T* st = stage0 + (size_t)s * L::H_STAGE;
        if (p < xe) {
            // bytes of the boxes q = w, w + BY, ... < 13: q % 3 == 2 is a halo-free H box, q >= 9 a row box
            uint32_t bytes = 0;
#pragma unroll
            for (int q = 0; q < 13; ++q)
                if (q % BY == w) bytes += q >= 9 ? BR : (q % 3 == 2 ? BP : BM);
            mbar_arrive_tx(bar, bytes);
            if (BY == 4) {
                switch (w) {
                    case 0: box_cur(0, st, p, bar); box_cur(4, st, p, bar); box_cur(8, st, p, bar); box_cur(12, st, p, bar); break;
                    case 1: box_cur(1, st, p, bar); box_cur(5, st, p, bar); box_cur(9, st, p, bar); break;
                    case 2: box_cur(2, st, p, bar); box_cur(6, st, p, bar); box_cur(10, st, p, bar); break;
                    default: box_cur(3, st, p, bar); box_cur(7, st, p, bar); box_cur(11, st, p, bar); break;
                }
            } else {
                switch (w) {
                    case 0: box_cur(0, st, p, bar); box_cur(8, st, p, bar); break;
                    case 1: box_cur(1, st, p, bar); box_cur(9, st, p, bar); break;
                    case 2: box_cur(2, st, p, bar); box_cur(10, st, p, bar); break;
                    case 3: box_cur(3, st, p, bar); box_cur(11, st, p, bar); break;
                    case 4: box_cur(4, st, p, bar); box_cur(12, st, p, bar); break;
                    case 5: box_cur(5, st, p, bar); break;
                    case 6: box_cur(6, st, p, bar); break;
                    default: box_cur(7, st, p, bar); break;
                }
            }
        } else {
            mbar_arrive_tx(bar, w < 4 ? BM : 0u);
            if (w < 4) box_nxt(w, st, p, bar);
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
    const int r0 = t.row * ROWP + col, r1 = r0 + ROWP;   // own row / row j+1 inside a block
    const int hrow = HOFF + t.row * BZ + col;

    int sc = 0;                                      // stage of the current plane, parity of its barrier
    uint32_t ph = 0;
    int sp = NS - 1;                                 // stage to refill at this iteration (plane i + NS - 1)
    T gn0 = T(0), gn1 = T(0), gn2 = T(0), gn3 = T(0);
    if (zedge) {
        const int okp = xs * plane + j * a.Nz;
        gn0 = a.mE[0][okp]; gn1 = a.Din[0][okp]; gn2 = a.mE[1][okp]; gn3 = a.Din[1][okp];
    }
    for (int i = xs; i < xe; ++i) {
        if (i + NS - 1 <= xe) issue(i + NS - 1, sp);
        sp = (sp + 1 == NS) ? 0 : sp + 1;
        const int sn = (sc + 1 == NS) ? 0 : sc + 1;
        const uint32_t phn = (sn == 0) ? ph ^ 1u : ph;
        const int q = i - xs;
        const int mx = xt.mx[q];
        const bool in_pml = pml.yz || mx >= 0;
        // the old PML integrals go through the LSU: issued (with the L2 prefetch of a later plane's) before the thread
        // blocks on the TMA barriers, so that their latency overlaps the wait and the shared-memory loads
        if (in_pml && active) pml.load(a, i, mx);
        {
            const int mxp = xt.mx[q + V5_PF];
            if (active && (pml.yz || mxp >= 0) && i + V5_PF < a.x1) pml.prefetch(a, i + V5_PF, mxp, (lane & 7) == 0);
        }
        const int pbase = i * plane;
        // periodic wrap along z: the box is zero-filled beyond the row, so the edge lane takes the cell k = 0 of its row
        // straight from global memory -- fetched ONE PLANE AHEAD (these loads miss every cache level: issued at the
        // point of use they put a DRAM round trip on the warp's critical path, 10 % of the kernel when measured)
        const T gx0 = gn0, gx1 = gn1, gx2 = gn2, gx3 = gn3;
        if (zedge && i + 1 < xe) {
            const int okp = pbase + plane + j * a.Nz;
            gn0 = a.mE[0][okp]; gn1 = a.Din[0][okp]; gn2 = a.mE[1][okp]; gn3 = a.Din[1][okp];
        }
        mbar_wait(&full[sc], ph);
        mbar_wait(&full[sn], phn);
        const T* cur = stage0 + (size_t)sc * L::H_STAGE;
        const T* nxt = stage0 + (size_t)sn * L::H_STAGE;

        AT E[3][V], CE[3][V];
        Vec<T, V> h[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Vec<T, V> d = ldv<T, V>(cur + c * BLK + r0);
            const Vec<T, V> m = ldv<T, V>(cur + (3 + c) * BLK + r0);
            h[c] = ldv<T, V>(cur + hrow + c * L::PLN);
#pragma unroll
            for (int e = 0; e < V; ++e) E[c][e] = mul_rn((AT)m.v[e], (AT)d.v[e]);
        }
        const Vec<T, V> dxj = ldv<T, V>(cur + 0 * BLK + r1), mxj = ldv<T, V>(cur + 3 * BLK + r1);
        const Vec<T, V> dzj = ldv<T, V>(cur + 2 * BLK + r1), mzj = ldv<T, V>(cur + 5 * BLK + r1);
        const Vec<T, V> dyn = ldv<T, V>(nxt + 1 * BLK + r0), myn = ldv<T, V>(nxt + 4 * BLK + r0);
        const Vec<T, V> dzn = ldv<T, V>(nxt + 2 * BLK + r0), mzn = ldv<T, V>(nxt + 5 * BLK + r0);
        AT ex_kp, ey_kp;
        if (zedge) {
            ex_kp = mul_rn((AT)gx0, (AT)gx1);
            ey_kp = mul_rn((AT)gx2, (AT)gx3);
        } else {
            ex_kp = mul_rn((AT)cur[3 * BLK + r0 + V], (AT)cur[0 * BLK + r0 + V]);
            ey_kp = mul_rn((AT)cur[4 * BLK + r0 + V], (AT)cur[1 * BLK + r0 + V]);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Ex = E[0][e], Ey = E[1][e], Ez = E[2][e];
            const AT Ex_jp = mul_rn((AT)mxj.v[e], (AT)dxj.v[e]);
            const AT Ez_jp = mul_rn((AT)mzj.v[e], (AT)dzj.v[e]);
            const AT Ex_kp = (e + 1 < V) ? E[0][(e + 1) % V] : ex_kp;
            const AT Ey_kp = (e + 1 < V) ? E[1][(e + 1) % V] : ey_kp;
            const AT Ey_ip = mul_rn((AT)myn.v[e], (AT)dyn.v[e]);
            const AT Ez_ip = mul_rn((AT)mzn.v[e], (AT)dzn.v[e]);
            CE[0][e] = curl2<AT>(Ez_jp, Ez, Ey_kp, Ey, inv);
            CE[1][e] = curl2<AT>(Ex_kp, Ex, Ez_ip, Ez, inv);
            CE[2][e] = curl2<AT>(Ey_ip, Ey, Ex_jp, Ex, inv);
        }
        if (active) {
            Vec<T, V> out[3];
            // path of the WARP: none / one axis / general (see PmlLean)
            const int path = (mx >= 0 ? 1 : 0) | warp_yz;
            if (path == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e) out[c].v[e] = (T)muladd(s, CE[c][e], (AT)h[c].v[e]);
            } else if (path == 1) {
                pml.apply_x(a, mx, xt.u[q], xt.r[q], s, h, CE, out);
            } else if (path == 2) {
                pml.apply_y(a, i, s, h, CE, out);
            } else if (path == 4) {
                pml.apply_z(a, i, s, h, CE, out);
            } else {
                pml.apply(a, i, mx, xt.u[q], xt.r[q], s, h, CE, out);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Hout[c] + pbase + orow, out[c]);
            if (a.peer_out[0] && i + 1 == a.Nx) {    // x-slab: H_y, H_z of the last plane are the right neighbour's i = -1
                stv<T, V>(a.peer_out[0] + orow, out[1]);
                stv<T, V>(a.peer_out[1] + orow, out[2]);
            }
        }
        __syncthreads();                             // every warp is done with the oldest stage: it may be refilled
        sc = sn;
        ph = phn;
    }
    if (a.peer_out[0] && xe == a.Nx && lane == 0 && w == 0) halo_signal(a.peer_flag);   // (after the loop's last barrier)
}

// ---------------------------------------------------------------------------------------------------------
// D half-step.  Plane p (p >= xs-1) -> stage (p - xs + 1) % NS; the H boxes start one vector BEFORE the tile
// along z (own cells at column V, the -1 neighbour of lane 0 at column V-1) and the halo row j-1 is row BY of
// each block (row 0 .. BY-1 are the own rows).
// Register budget of the D half-step: by default the compiler's own choice under __launch_bounds__(threads) (158 registers in
// fp64 = 3 CTAs per SM, 122 in fp32 = 4).  -DCEV_V5_D_MINB=n caps it for n CTAs per SM (4: 128 registers in fp64 without
// spills) -- an A/B knob; note that n = 1 is NOT the default: it lets the compiler take 196 registers.
#ifdef CEV_V5_D_MINB
#define CEV_V5_D_BOUNDS __launch_bounds__(32 * BY, CEV_V5_D_MINB)
#else
#define CEV_V5_D_BOUNDS __launch_bounds__(32 * BY)
#endif

template <typename T, typename AT, int V, int BY, int NS>
__global__ void CEV_V5_D_BOUNDS k_step_D_v5(const StepArgs<T, AT> a, const __grid_constant__ V5MapsD maps) {
    using L = V5Layout<T, V, BY>;
    constexpr int BZ = L::BZ, ROWP = L::ROWP, BLK = L::BLK, DOFF = 3 * L::BLK;
    extern __shared__ __align__(128) unsigned char v5_smem[];
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    T* const stage0 = reinterpret_cast<T*>(v5_smem);
    uint64_t* const full = reinterpret_cast<uint64_t*>(v5_smem + (size_t)NS * L::D_STAGE * sizeof(T));
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

    // 8 boxes per current plane (H[3] with halos, D[3], the halo rows of H_x and H_z), dealt round-robin to the warps;
    // p == xs - 1 (the x-1 neighbour of the chunk's first plane): own rows of H_y, H_z only
    const int yrow = (y0 == 0) ? a.Ny - 1 : y0 - 1;
    auto box_cur = [&](int q, T* st, int p, uint64_t* bar) {
        if (q < 6) {
            const int c = q / 2;
            if (q % 2 == 0) tma_box_3d(st + c * BLK, &maps.H[c], z0 - V, y0, p, bar);
            else tma_box_3d(st + DOFF + c * L::PLN, &maps.D[c], z0, y0, p, bar);
        } else {
            const int r = q - 6;                           // 0: component x, 1: component z
            tma_box_3d(st + 2 * r * BLK + BY * ROWP, &maps.Hrow[r], z0 - V, yrow, p, bar);
        }
    };
    constexpr uint32_t BM = L::BOX_MAIN, BR = L::BOX_ROW, BP = L::BOX_PLN;
    auto issue = [&](int p, int s) {                       // called by every warp, converged
        if (!elect_one()) return;
        uint64_t* bar = &full[s];
        T* st = stage0 + (size_t)s * L::D_STAGE;
        if (p >= xs) {
            uint32_t bytes = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q % BY == w) bytes += q >= 6 ? BR : (q % 2 == 1 ? BP : BM);
            mbar_arrive_tx(bar, bytes);
            if (BY == 4) {
                switch (w) {
                    case 0: box_cur(0, st, p, bar); box_cur(4, st, p, bar); break;
                    case 1: box_cur(1, st, p, bar); box_cur(5, st, p, bar); break;
                    case 2: box_cur(2, st, p, bar); box_cur(6, st, p, bar); break;
                    default: box_cur(3, st, p, bar); box_cur(7, st, p, bar); break;
                }
            } else {
                switch (w) {
                    case 0: box_cur(0, st, p, bar); break;
                    case 1: box_cur(1, st, p, bar); break;
                    case 2: box_cur(2, st, p, bar); break;
                    case 3: box_cur(3, st, p, bar); break;
                    case 4: box_cur(4, st, p, bar); break;
                    case 5: box_cur(5, st, p, bar); break;
                    case 6: box_cur(6, st, p, bar); break;
                    default: box_cur(7, st, p, bar); break;
                }
            }
        } else {                                     // p == xs - 1
            mbar_arrive_tx(bar, w < 2 ? BM : 0u);
            if (w < 2) {
                const bool lo = p < 0;
                if (lo && a.own_flag) halo_wait(a.own_flag, a.own_target, a.halo_err);    // x-slab: the left neighbour's last plane
                const int px = lo ? maps.x_lo : p;
                tma_box_3d(st + (1 + w) * BLK, lo ? &maps.Hlo[w] : &maps.H[1 + w], z0 - V, y0, px, bar);
            }
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
    const int r0 = t.row * ROWP + col, rm = (t.row == 0 ? BY : t.row - 1) * ROWP + col;   // own row / row j-1 inside a block
    const int drow = DOFF + t.row * BZ + t.vec * V;

    int sc = 0;                                      // stage of plane i-1
    uint32_t ph = 0;
    int sp = NS - 1;                                 // stage to refill at this iteration (plane i + NS - 2)
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
        const bool in_pml = pml.yz || mx >= 0;
        if (in_pml && active) pml.load(a, i, mx);
        {
            const int mxp = xt.mx[q + V5_PF];
            if (active && (pml.yz || mxp >= 0) && i + V5_PF < a.x1) pml.prefetch(a, i + V5_PF, mxp, (lane & 7) == 0);
        }
        const int pbase = i * plane;
        // periodic wrap along z: the cell k = Nz-1 of the row, straight from global memory, one plane ahead (see k_step_H_v5)
        const T gx0 = gn0, gx1 = gn1;
        if (zedge && i + 1 < xe) {
            const int okm = pbase + plane + j * a.Nz + a.Nz - 1;
            gn0 = a.Hin[0][okm]; gn1 = a.Hin[1][okm];
        }
        mbar_wait(&full[sc], ph);
        mbar_wait(&full[sn], phn);
        const T* prv = stage0 + (size_t)sc * L::D_STAGE;
        const T* cur = stage0 + (size_t)sn * L::D_STAGE;

        Vec<T, V> h[3], d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            h[c] = ldv<T, V>(cur + c * BLK + r0);
            d[c] = ldv<T, V>(cur + drow + c * L::PLN);
        }
        const Vec<T, V> hxj = ldv<T, V>(cur + 0 * BLK + rm), hzj = ldv<T, V>(cur + 2 * BLK + rm);
        const Vec<T, V> hyp = ldv<T, V>(prv + 1 * BLK + r0), hzp = ldv<T, V>(prv + 2 * BLK + r0);
        AT hx_km, hy_km;
        if (zedge) {
            hx_km = (AT)gx0;
            hy_km = (AT)gx1;
        } else {
            hx_km = (AT)cur[0 * BLK + r0 - 1];
            hy_km = (AT)cur[1 * BLK + r0 - 1];
        }
        AT CH[3][V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Hx = (AT)h[0].v[e], Hy = (AT)h[1].v[e], Hz = (AT)h[2].v[e];
            const AT Hx_km = (e > 0) ? (AT)h[0].v[(e + V - 1) % V] : hx_km;
            const AT Hy_km = (e > 0) ? (AT)h[1].v[(e + V - 1) % V] : hy_km;
            CH[0][e] = curl2<AT>(Hz, (AT)hzj.v[e], Hy, Hy_km, inv);
            CH[1][e] = curl2<AT>(Hx, Hx_km, Hz, (AT)hzp.v[e], inv);
            CH[2][e] = curl2<AT>(Hy, (AT)hyp.v[e], Hx, (AT)hxj.v[e], inv);
        }
        if (active) {
            Vec<T, V> out[3];
            const int path = (mx >= 0 ? 1 : 0) | warp_yz;
            if (path == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e) out[c].v[e] = (T)muladd(s, CH[c][e], (AT)d[c].v[e]);
            } else if (path == 1) {
                pml.apply_x(a, mx, xt.u[q], xt.r[q], s, d, CH, out);
            } else if (path == 2) {
                pml.apply_y(a, i, s, d, CH, out);
            } else if (path == 4) {
                pml.apply_z(a, i, s, d, CH, out);
            } else {
                pml.apply(a, i, mx, xt.u[q], xt.r[q], s, d, CH, out);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Dout[c] + pbase + orow, out[c]);
        }
        __syncthreads();
        sc = sn;
        ph = phn;
    }

    // ---- in-kernel source injection: D += J after the update (fdtd.py:125-127)
    if (a.src_wave) {
        inject_points<T, AT, int32_t>(a.src_begin[bid], a.src_begin[bid + 1], threadIdx.y * 32 + threadIdx.x, 32 * BY, a.src_comp,
                                      a.src_id, a.src_cell, a.src_w, a.src_wave, a.Dout[0], a.Dout[1], a.Dout[2]);
    }
    // ---- x-slab: D_y, D_z of plane 0 (sources included) are the left neighbour's i = nx plane
    if (a.peer_out[0] && xs == 0) {
        __syncthreads();                             // the injections above have been issued
        if (active) {
            const Vec<T, V> dy = ldv_cg<T, V>(a.Dout[1] + orow);     // (L2, where the injections landed)
            const Vec<T, V> dz = ldv_cg<T, V>(a.Dout[2] + orow);
            stv<T, V>(a.peer_out[0] + orow, dy);
            stv<T, V>(a.peer_out[1] + orow, dz);
        }
        __syncthreads();
        if (lane == 0 && w == 0) halo_signal(a.peer_flag);
    }
}

}  // namespace cev
