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
// load unless the component is known to be identically zero (kernel-uniform condition)
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> ldv_if(bool on, const T* p) {
    Vec<T, V> z;
#pragma unroll
    for (int e = 0; e < V; ++e) z.v[e] = T(0);
    return on ? ldv<T, V>(p) : z;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

constexpr int V2_BY = 4;        // warps per CTA; a warp covers LZ lanes along z times 32/LZ rows
// register caps = 65536 / (MIN_CTAS * 128) per thread; tuned on B200 (scripts/tune.py, 256^3 with and
// without PML): the H kernel carries more state and prefers fewer, fatter CTAs (3 for fp64, 4 for fp32);
// the D kernel prefers occupancy
#ifndef V2_H_MIN_CTAS
#define V2_H_MIN_CTAS 3
#endif
#ifndef V2_D_MIN_CTAS
#define V2_D_MIN_CTAS 5
#endif
#ifndef V2_INTERIOR_MIN_CTAS
#define V2_INTERIOR_MIN_CTAS 6
#endif

// PML work of one thread for one x-plane.  load() is called right after the main loads of the
// iteration are issued: it fetches the old values of the curl integrals (vector loads for the x- and
// y-slab arrays, whose z-runs are contiguous) and the table entries, so that their latency overlaps
// the field loads instead of adding a second memory round trip.  apply() then does the general
// update of fdtd.py:85-97 / :110-122 for the V cells.  IS_H selects the H or D tables at compile
// time (no address of the parameter struct is taken: it stays in the constant bank).
template <typename T, typename AT, int V, bool IS_H>
struct PmlCtx {
    Vec<T, V> I0, I1;     // old Icurl_x / Icurl_y runs
    T I2[V];              // old Icurl_z cells
    int ic0, ic1;         // run offsets (or -1)

#define CEV_TAB(a, name, ax) (IS_H ? (a).name##H[ax] : (a).name##D[ax])
    __device__ __forceinline__ void load(const StepArgs<T, AT>& a, int i, int j, int k0, int mx, int my, const int* mz,
                                         unsigned on3) {
        const int n1 = IS_H ? a.nH[1] : a.nD[1], n2 = IS_H ? a.nH[2] : a.nD[2];
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        ic0 = (mx >= 0 && (on3 & 1)) ? (mx * a.Ny + j) * a.Nz + k0 : -1;      // Icurl_x (nCx,Ny,Nz)
        ic1 = (my >= 0 && (on3 & 2)) ? (i * n1 + my) * a.Nz + k0 : -1;        // Icurl_y (Nx,nCy,Nz)
#ifdef CEV_EXP_NO_RMW
        ic0 = ic1 = -1;
#endif
        if (ic0 >= 0) I0 = ldv<T, V>(Ic[0] + ic0);
        if (ic1 >= 0) I1 = ldv<T, V>(Ic[1] + ic1);
#pragma unroll
        for (int e = 0; e < V; ++e)
            if (mz[e] >= 0 && (on3 & 4)) I2[e] = Ic[2][(i * a.Ny + j) * n2 + mz[e]];   // Icurl_z (Nx,Ny,nCz)
    }

    // pull the curl integrals of plane ip (a later iteration of this thread) into L2: without it they
    // are the only loads of a PML iteration that still pay the full DRAM latency
    static __device__ __forceinline__ void prefetch(const StepArgs<T, AT>& a, int ip, int j, int k0, int my, const int* mz,
                                                    bool line_lane) {
        const int n1 = IS_H ? a.nH[1] : a.nD[1], n2 = IS_H ? a.nH[2] : a.nD[2];
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        const int mxp = CEV_TAB(a, map, 0)[ip];
        if (line_lane) {
            if (mxp >= 0) prefetch_l2(Ic[0] + (mxp * a.Ny + j) * a.Nz + k0);
            if (my >= 0) prefetch_l2(Ic[1] + (ip * n1 + my) * a.Nz + k0);
        }
        if (mz[0] >= 0) prefetch_l2(Ic[2] + (ip * a.Ny + j) * n2 + mz[0]);
        else if (mz[V - 1] >= 0) prefetch_l2(Ic[2] + (ip * a.Ny + j) * n2 + mz[V - 1]);
    }

    // old[c][e], curl[c][e] -> out[c].v[e]
    __device__ __forceinline__ void apply(const StepArgs<T, AT>& a, int i, int j, int k0, int mx, int my, const int* mz,
                                          AT s, const Vec<T, V>* old, const AT (*curl)[V], Vec<T, V>* out, unsigned on3) {
        const int n1 = IS_H ? a.nH[1] : a.nD[1], n2 = IS_H ? a.nH[2] : a.nD[2];
        T* const* Ic = IS_H ? a.ICE : a.ICH;
        T* const* Is = IS_H ? a.IH : a.ID;
        // the tiny per-axis tables are L1-resident: fetched here, after the long-latency loads were consumed,
        // so they do not hold registers while the thread waits for HBM
        const AT ux = CEV_TAB(a, u, 0)[i], rx = CEV_TAB(a, r, 0)[i];
        const AT uy = CEV_TAB(a, u, 1)[j], ry = CEV_TAB(a, r, 1)[j];
        AT uz[V], rz[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            uz[e] = CEV_TAB(a, u, 2)[k0 + e];
            rz[e] = CEV_TAB(a, r, 2)[k0 + e];
        }
        Vec<T, V> n0, n1v;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int k = k0 + e;
            // x: (a,b) = (y,z), own x.   Iself_x (Nx,nCy,nCz)
            if (on3 & 1) {
                AT m1, m2;
                coef12<AT>(uy, ry, uz[e], rz[e], s, m1, m2);
                AT v = muladd(m1, (AT)old[0].v[e], mul_rn(m2, curl[0][e]));
                if (ic0 >= 0) {
                    const AT I = (AT)I0.v[e] + curl[0][e];
                    n0.v[e] = (T)I;
                    v = muladd(mul_rn(mul_rn(s, ux + ux), mul_rn(ry, rz[e])), I, v);
                }
                if (my >= 0 && mz[e] >= 0) {
                    T* q = Is[0] + (i * n1 + my) * n2 + mz[e];
                    const AT I = (AT)*q + (AT)old[0].v[e];
                    *q = (T)I;
                    v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), uy), uz[e]), mul_rn(ry, rz[e])), I, v);
                }
                out[0].v[e] = (T)v;
            }
            // y: (a,b) = (x,z), own y.   Iself_y (nCx,Ny,nCz)
            if (on3 & 2) {
                AT m1, m2;
                coef12<AT>(ux, rx, uz[e], rz[e], s, m1, m2);
                AT v = muladd(m1, (AT)old[1].v[e], mul_rn(m2, curl[1][e]));
                if (ic1 >= 0) {
                    const AT I = (AT)I1.v[e] + curl[1][e];
                    n1v.v[e] = (T)I;
                    v = muladd(mul_rn(mul_rn(s, uy + uy), mul_rn(rx, rz[e])), I, v);
                }
                if (mx >= 0 && mz[e] >= 0) {
                    T* q = Is[1] + (mx * a.Ny + j) * n2 + mz[e];
                    const AT I = (AT)*q + (AT)old[1].v[e];
                    *q = (T)I;
                    v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), ux), uz[e]), mul_rn(rx, rz[e])), I, v);
                }
                out[1].v[e] = (T)v;
            }
            // z: (a,b) = (x,y), own z.   Iself_z (nCx,nCy,Nz)
            if (on3 & 4) {
                AT m1, m2;
                coef12<AT>(ux, rx, uy, ry, s, m1, m2);
                AT v = muladd(m1, (AT)old[2].v[e], mul_rn(m2, curl[2][e]));
                if (mz[e] >= 0) {   // (on3 & 4 holds here)
                    const AT I = (AT)I2[e] + curl[2][e];
                    Ic[2][(i * a.Ny + j) * n2 + mz[e]] = (T)I;
                    v = muladd(mul_rn(mul_rn(s, uz[e] + uz[e]), mul_rn(rx, ry)), I, v);
                }
                if (mx >= 0 && my >= 0) {
                    T* q = Is[2] + (mx * n1 + my) * a.Nz + k;
                    const AT I = (AT)*q + (AT)old[2].v[e];
                    *q = (T)I;
                    v = muladd(mul_rn(mul_rn(mul_rn(AT(-4), ux), uy), mul_rn(rx, ry)), I, v);
                }
                out[2].v[e] = (T)v;
            }
        }
        if (ic0 >= 0) stv<T, V>(Ic[0] + ic0, n0);
        if (ic1 >= 0) stv<T, V>(Ic[1] + ic1, n1v);
    }
#undef CEV_TAB
};

// INTERIOR = the launch covers only cells off every PML: all PML code is compiled out (few registers,
// high occupancy); otherwise the general kernel.
// MASK = which components are live (bits 0-2 E/D, 3-5 H, internal order): 63 = all six, a compile-time constant with
// which every mask test folds away and the code is exactly the unmasked kernel (runtime tests cost the full-vector
// kernels 20-50 % on B200: measured, profiles/); MASK_TM / MASK_TE = the two polarisations of a 2-D grid, also
// compile-time (lean kernels: the dead components cost neither registers nor instructions); -1 = read a.on at run time.
constexpr int MASK_TM = 0b101010;   // 2-D grids are relabelled (x, z, y): logical {Dz, Hx, Hy} = internal {D_y, H_x, H_z}
constexpr int MASK_TE = 0b010101;   //                                   logical {Dx, Dy, Hz} = internal {D_x, D_z, H_y}
// A batch of launches of the same kernel on B independent states (the tangent states of a forward-mode sweep) as ONE
// launch: CTA blockIdx.x serves state blockIdx.x % B (neighbouring CTAs work on the same tile of different states, so
// the arrays the states share -- 1/eps and the primal D of a tangent H half-step -- are read from HBM once and hit in L2
// for the others).  The per-state StepArgs live in a device table; the CTA copies its entry to shared memory.
template <typename T, typename AT>
__device__ __forceinline__ const StepArgs<T, AT>& batch_args(const StepArgs<T, AT>* table, int b, int64_t t_probe) {
    __shared__ __align__(16) unsigned char raw[sizeof(StepArgs<T, AT>)];
    static_assert(sizeof(StepArgs<T, AT>) % 4 == 0, "StepArgs is copied word by word");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(table + b);
    uint32_t* dst = reinterpret_cast<uint32_t*>(raw);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nth = blockDim.x * blockDim.y;
    for (int q = tid; q < (int)(sizeof(StepArgs<T, AT>) / 4); q += nth) dst[q] = src[q];
    __syncthreads();
    StepArgs<T, AT>& a = *reinterpret_cast<StepArgs<T, AT>*>(raw);
    if (tid == 0) a.t_probe = t_probe;
    __syncthreads();
    return a;
}

// TAN = forward-mode tangent step: the state is a TANGENT state and E = mE*D + dmE*D_primal (product rule on
// fdtd.py:135-137), same rounding sequence as the baseline kernel; with TAN = false the extra loads fold away.
template <bool TAN, typename T, typename AT, int V>
__device__ __forceinline__ AT e_of(const Vec<T, V>& m, const Vec<T, V>& d, const Vec<T, V>& dm, const Vec<T, V>& dp, int e) {
    const AT x = mul_rn((AT)m.v[e], (AT)d.v[e]);
    return TAN ? add_rn(x, mul_rn((AT)dm.v[e], (AT)dp.v[e])) : x;
}

template <typename T, typename AT, int V, int LZ, bool INTERIOR, int MASK = 63, bool TAN = false>
__device__ __forceinline__ void step_H_v2_body(const StepArgs<T, AT>& a, const int bid) {
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    constexpr int RW = 32 / LZ;                  // rows per warp
    const int lane = threadIdx.x;
    const int lz = lane % LZ, ly = lane / LZ;
    int bx = 0;                                  // which box of this launch (CTA-uniform)
#pragma unroll
    for (int q = 1; q < MAX_BOXES; ++q)
        if (q < a.n_boxes && bid >= a.box[q].cta0) bx = q;
    const Box& B = a.box[bx];
    const int lid = bid - B.cta0;
    const int tz = lid % B.ntz;
    const int rest = lid / B.ntz;
    const int ty = rest % B.nty;
    const int xc = rest / B.nty;
    // warps of the CTA stacked along y (default) or, for single-row planes, along z
    const int jraw = B.y0 + (a.wz ? ty : ty * V2_BY + threadIdx.y) * RW + ly;
    if (jraw - ly >= B.y1) return;               // warp-uniform: the whole warp is outside the box
    const int k0raw = B.z0 + ((a.wz ? tz * V2_BY + threadIdx.y : tz) * LZ + lz) * V;
    const bool active = k0raw < B.z1 && jraw < B.y1;
    const int j = jraw < B.y1 ? jraw : B.y1 - 1; // inactive lanes shadow a valid cell (loads only)
    const int k0 = k0raw < B.z1 ? k0raw : B.z0;
    const int xs = B.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, B.x1);

    const int plane = a.Ny * a.Nz;
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    const bool z_edge = (lz == LZ - 1) || (k0 + V >= B.z1);   // +1 neighbour not in lane+1 (tile / box edge)
    const int kp = (k0 + V >= a.Nz) ? 0 : k0 + V;
    const int orow = j * a.Nz + k0;
    const int orow_jp = jp * a.Nz + k0;
    const int okp = j * a.Nz + kp;

    const int my = INTERIOR ? -1 : a.mapH[1][j];
    int mz[V];
    bool yz_pml = my >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mz[e] = INTERIOR ? -1 : a.mapH[2][k0 + e];
        yz_pml |= mz[e] >= 0;
    }
    const AT s = -a.cdt;
    const AT inv = a.inv_dL;
    const unsigned on_all = MASK < 0 ? a.on : (unsigned)MASK;
    const unsigned onE = on_all & 7u, onH = (on_all >> 3) & 7u;

    // E = mE*D of the current plane (own cells)
    AT Ecur[3][V];
    {
        const int o = xs * plane + orow;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const bool oc = (onE >> c) & 1u;
            const Vec<T, V> d = ldv_if<T, V>(oc, a.Din[c] + o), m = ldv_if<T, V>(oc, a.mE[c] + o);
            Vec<T, V> td, tm;
            if (TAN) {
                td = ldv_if<T, V>(oc, a.Dp[c] + o);
                tm = ldv_if<T, V>(oc, a.dmE[c] + o);
            }
#pragma unroll
            for (int e = 0; e < V; ++e) Ecur[c][e] = e_of<TAN, T, AT, V>(m, d, tm, td, e);
        }
    }

    for (int i = xs; i < xe; ++i) {
        const int pbase = i * plane;
        const bool last = (i + 1 == a.Nx);
        // ---- every load of this iteration, up front
        Vec<T, V> dn[3], mn[3], h[3];
        Vec<T, V> tdn[3], tmn[3], tdxj, tmxj, tdzj, tmzj;     // tangent steps only: D_primal, d(1/eps)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const T* Dn = last ? a.Dhi[c] : a.Din[c] + pbase + plane;
            const T* Mn = last ? a.mEhi[c] : a.mE[c] + pbase + plane;
            dn[c] = ldv_if<T, V>((onE >> c) & 1u, Dn + orow);
            mn[c] = ldv_if<T, V>((onE >> c) & 1u, Mn + orow);
            h[c] = ldv_if<T, V>((onH >> c) & 1u, a.Hin[c] + pbase + orow);
            if (TAN) {
                tdn[c] = ldv_if<T, V>((onE >> c) & 1u, (last ? a.Dphi[c] : a.Dp[c] + pbase + plane) + orow);
                tmn[c] = ldv_if<T, V>((onE >> c) & 1u, (last ? a.dmEhi[c] : a.dmE[c] + pbase + plane) + orow);
            }
        }
        const Vec<T, V> dxj = ldv_if<T, V>(onE & 1u, a.Din[0] + pbase + orow_jp), mxj = ldv_if<T, V>(onE & 1u, a.mE[0] + pbase + orow_jp);
        const Vec<T, V> dzj = ldv_if<T, V>(onE & 4u, a.Din[2] + pbase + orow_jp), mzj = ldv_if<T, V>(onE & 4u, a.mE[2] + pbase + orow_jp);
        if (TAN) {
            tdxj = ldv_if<T, V>(onE & 1u, a.Dp[0] + pbase + orow_jp);
            tmxj = ldv_if<T, V>(onE & 1u, a.dmE[0] + pbase + orow_jp);
            tdzj = ldv_if<T, V>(onE & 4u, a.Dp[2] + pbase + orow_jp);
            tmzj = ldv_if<T, V>(onE & 4u, a.dmE[2] + pbase + orow_jp);
        }
        AT ex_kp = __shfl_down_sync(0xffffffffu, Ecur[0][0], 1);
        AT ey_kp = __shfl_down_sync(0xffffffffu, Ecur[1][0], 1);
        if (z_edge) {
            if (onE & 1u) {
                ex_kp = mul_rn((AT)a.mE[0][pbase + okp], (AT)a.Din[0][pbase + okp]);
                if (TAN) ex_kp = add_rn(ex_kp, mul_rn((AT)a.dmE[0][pbase + okp], (AT)a.Dp[0][pbase + okp]));
            }
            if (onE & 2u) {
                ey_kp = mul_rn((AT)a.mE[1][pbase + okp], (AT)a.Din[1][pbase + okp]);
                if (TAN) ey_kp = add_rn(ey_kp, mul_rn((AT)a.dmE[1][pbase + okp], (AT)a.Dp[1][pbase + okp]));
            }
        }
        if (a.pf_dist > 0 && i + a.pf_dist < a.x1) {   // next plane(s) of every stream into L2, also across the chunk end
            const int ip = i + a.pf_dist;
            if ((lz & 7) == 0) {                       // one lane per 128-byte line
                const int po = ip * plane + orow;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if ((onH >> c) & 1u) prefetch_l2(a.Hin[c] + po);
                    if (ip + 1 < a.Nx && ((onE >> c) & 1u)) {
                        prefetch_l2(a.Din[c] + po + plane);
                        prefetch_l2(a.mE[c] + po + plane);
                        if (TAN) {
                            prefetch_l2(a.Dp[c] + po + plane);
                            prefetch_l2(a.dmE[c] + po + plane);
                        }
                    }
                }
            }
            if (!INTERIOR) PmlCtx<T, AT, V, true>::prefetch(a, ip, j, k0, my, mz, (lz & 7) == 0);
        }
        const int mx = INTERIOR ? -1 : a.mapH[0][i];
#ifdef CEV_EXP_NO_PML
        const bool pml = false;
#else
        const bool pml = !INTERIOR && (yz_pml || mx >= 0);
#endif
        PmlCtx<T, AT, V, true> ctx;
        if (pml) ctx.load(a, i, j, k0, mx, my, mz, onH);

        // ---- curls (consume the neighbour loads); E of the next plane becomes current
        AT CE[3][V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const AT Ex = Ecur[0][e], Ey = Ecur[1][e], Ez = Ecur[2][e];
            const AT Ex_jp = e_of<TAN, T, AT, V>(mxj, dxj, tmxj, tdxj, e);
            const AT Ez_jp = e_of<TAN, T, AT, V>(mzj, dzj, tmzj, tdzj, e);
            const AT Ex_kp = (e + 1 < V) ? Ecur[0][(e + 1) % V] : ex_kp;
            const AT Ey_kp = (e + 1 < V) ? Ecur[1][(e + 1) % V] : ey_kp;
            const AT Ey_ip = e_of<TAN, T, AT, V>(mn[1], dn[1], tmn[1], tdn[1], e);
            const AT Ez_ip = e_of<TAN, T, AT, V>(mn[2], dn[2], tmn[2], tdn[2], e);
            CE[0][e] = curl2<AT>(Ez_jp, Ez, Ey_kp, Ey, inv);
            CE[1][e] = curl2<AT>(Ex_kp, Ex, Ez_ip, Ez, inv);
            CE[2][e] = curl2<AT>(Ey_ip, Ey, Ex_jp, Ex, inv);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            Ecur[0][e] = e_of<TAN, T, AT, V>(mn[0], dn[0], tmn[0], tdn[0], e);
            Ecur[1][e] = e_of<TAN, T, AT, V>(mn[1], dn[1], tmn[1], tdn[1], e);
            Ecur[2][e] = e_of<TAN, T, AT, V>(mn[2], dn[2], tmn[2], tdn[2], e);
        }

        // ---- update
        if (active) {
            Vec<T, V> out[3];
            if (!pml) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e)   // m1 = 1, m2 = -C0 dt: exactly what the general formula gives off the PML
                        out[c].v[e] = (T)muladd(s, CE[c][e], (AT)h[c].v[e]);
            } else {
                ctx.apply(a, i, j, k0, mx, my, mz, s, h, CE, out, onH);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if ((onH >> c) & 1u) stv<T, V>(a.Hout[c] + pbase + orow, out[c]);
        }
    }
}
template <typename T, typename AT, int V, int LZ, bool INTERIOR, int MASK = 63, bool TAN = false>
__global__ void __launch_bounds__(32 * V2_BY, (INTERIOR ? V2_INTERIOR_MIN_CTAS : ((TAN && MASK != MASK_TM && MASK != MASK_TE) ? 2 : (sizeof(T) == 8 ? V2_H_MIN_CTAS : V2_H_MIN_CTAS + 1))))
k_step_H_v2(const StepArgs<T, AT> a) {
    step_H_v2_body<T, AT, V, LZ, INTERIOR, MASK, TAN>(a, blockIdx.x);
}

template <typename T, typename AT, int V, int LZ, bool INTERIOR, int MASK = 63, bool TAN = false>
__global__ void __launch_bounds__(32 * V2_BY, (INTERIOR ? V2_INTERIOR_MIN_CTAS : ((TAN && MASK != MASK_TM && MASK != MASK_TE) ? 2 : (sizeof(T) == 8 ? V2_H_MIN_CTAS : V2_H_MIN_CTAS + 1))))
k_step_H_v2_batch(const StepArgs<T, AT>* table, int B, int64_t t_probe) {
    const StepArgs<T, AT>& a = batch_args<T, AT>(table, blockIdx.x % B, t_probe);
    step_H_v2_body<T, AT, V, LZ, INTERIOR, MASK, TAN>(a, blockIdx.x / B);
}


// EXTRAS = dense J input and/or E output (the per-step forward() API); the fused run() path
// instantiates EXTRAS = false and carries neither.
template <typename T, typename AT, int V, int LZ, bool EXTRAS, bool INTERIOR, int MASK = 63>
__device__ __forceinline__ void step_D_v2_body(const StepArgs<T, AT>& a, const int bid) {
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    constexpr int RW = 32 / LZ;                  // rows per warp
    const int lane = threadIdx.x;
    const int lz = lane % LZ, ly = lane / LZ;
    int bx = 0;                                  // which box of this launch (CTA-uniform)
#pragma unroll
    for (int q = 1; q < MAX_BOXES; ++q)
        if (q < a.n_boxes && bid >= a.box[q].cta0) bx = q;
    const Box& B = a.box[bx];
    const int lid = bid - B.cta0;
    const int tz = lid % B.ntz;
    const int rest = lid / B.ntz;
    const int ty = rest % B.nty;
    const int xc = rest / B.nty;
    const int jraw = B.y0 + (a.wz ? ty : ty * V2_BY + threadIdx.y) * RW + ly;
    const bool warp_on = jraw - ly < B.y1;       // warp-uniform; no early return: the CTA meets at a barrier below
    const int k0raw = B.z0 + ((a.wz ? tz * V2_BY + threadIdx.y : tz) * LZ + lz) * V;
    const bool active = k0raw < B.z1 && jraw < B.y1;
    const int j = jraw < B.y1 ? jraw : B.y1 - 1; // inactive lanes shadow a valid cell (loads only)
    const int k0 = k0raw < B.z1 ? k0raw : B.z0;
    const int xs = B.x0 + xc * a.xchunk;
    const int xe = min(xs + a.xchunk, B.x1);

    const int plane = a.Ny * a.Nz;
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    const bool z_edge = (lz == 0) || (k0 == B.z0);            // -1 neighbour not in lane-1 (tile / box edge)
    const int km = (k0 == 0) ? a.Nz - 1 : k0 - 1;
    const int orow = j * a.Nz + k0;
    const int orow_jm = jm * a.Nz + k0;
    const int okm = j * a.Nz + km;

    const int my = INTERIOR ? -1 : a.mapD[1][j];
    int mz[V];
    bool yz_pml = my >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mz[e] = INTERIOR ? -1 : a.mapD[2][k0 + e];
        yz_pml |= mz[e] >= 0;
    }
    const AT s = a.cdt;
    const AT inv = a.inv_dL;
    const unsigned on_all = MASK < 0 ? a.on : (unsigned)MASK;
    const unsigned onE = on_all & 7u, onH = (on_all >> 3) & 7u;

    // H of the previous plane (own cells), y and z components
    AT Hprev[2][V];
    {
        const T* P1 = (xs > 0) ? a.Hin[1] + (xs - 1) * plane : a.Hlo[1];
        const T* P2 = (xs > 0) ? a.Hin[2] + (xs - 1) * plane : a.Hlo[2];
        const Vec<T, V> p1 = ldv_if<T, V>(onH & 2u, P1 + orow), p2 = ldv_if<T, V>(onH & 4u, P2 + orow);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            Hprev[0][e] = (AT)p1.v[e];
            Hprev[1][e] = (AT)p2.v[e];
        }
    }

    for (int i = xs; i < (warp_on ? xe : xs); ++i) {
        const int pbase = i * plane;
        Vec<T, V> h[3], d[3], jv[3], mev[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            h[c] = ldv_if<T, V>((onH >> c) & 1u, a.Hin[c] + pbase + orow);
            d[c] = ldv_if<T, V>((onE >> c) & 1u, a.Din[c] + pbase + orow);
            if (EXTRAS && a.J[c]) jv[c] = ldv<T, V>(a.J[c] + pbase + orow);
            if (EXTRAS && a.Eout[c]) mev[c] = ldv<T, V>(a.mE[c] + pbase + orow);
        }
        const Vec<T, V> hxj = ldv_if<T, V>(onH & 1u, a.Hin[0] + pbase + orow_jm), hzj = ldv_if<T, V>(onH & 4u, a.Hin[2] + pbase + orow_jm);
        AT hx_km = __shfl_up_sync(0xffffffffu, (AT)h[0].v[V - 1], 1);
        AT hy_km = __shfl_up_sync(0xffffffffu, (AT)h[1].v[V - 1], 1);
        if (z_edge) {
            if (onH & 1u) hx_km = (AT)a.Hin[0][pbase + okm];
            if (onH & 2u) hy_km = (AT)a.Hin[1][pbase + okm];
        }
        if (a.pf_dist > 0 && i + a.pf_dist < a.x1) {
            const int ip = i + a.pf_dist;
            if ((lz & 7) == 0) {
                const int po = ip * plane + orow;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if ((onH >> c) & 1u) prefetch_l2(a.Hin[c] + po);
                    if ((onE >> c) & 1u) prefetch_l2(a.Din[c] + po);
                }
            }
            if (!INTERIOR) PmlCtx<T, AT, V, false>::prefetch(a, ip, j, k0, my, mz, (lz & 7) == 0);
        }
        const int mx = INTERIOR ? -1 : a.mapD[0][i];
#ifdef CEV_EXP_NO_PML
        const bool pml = false;
#else
        const bool pml = !INTERIOR && (yz_pml || mx >= 0);
#endif
        PmlCtx<T, AT, V, false> ctx;
        if (pml) ctx.load(a, i, j, k0, mx, my, mz, onE);

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
                    for (int e = 0; e < V; ++e) out[c].v[e] = (T)muladd(s, CH[c][e], (AT)d[c].v[e]);
            } else {
                ctx.apply(a, i, j, k0, mx, my, mz, s, d, CH, out, onE);
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
                if ((onE >> c) & 1u) stv<T, V>(a.Dout[c] + pbase + orow, out[c]);
            }
        }
    }

    // ---- in-kernel source injection: D += J after the update (fdtd.py:125-127)
    if (a.src_wave) {
        __syncthreads();                         // every D store of this CTA has been issued and is visible CTA-wide
        inject_points<T, AT, int32_t>(a.src_begin[bid], a.src_begin[bid + 1], threadIdx.y * 32 + threadIdx.x, 32 * V2_BY, a.src_comp,
                                      a.src_id, a.src_cell, a.src_w, a.src_wave, a.Dout[0], a.Dout[1], a.Dout[2]);
    }
}
template <typename T, typename AT, int V, int LZ, bool EXTRAS, bool INTERIOR, int MASK = 63>
__global__ void __launch_bounds__(32 * V2_BY, (INTERIOR ? V2_INTERIOR_MIN_CTAS : V2_D_MIN_CTAS)) k_step_D_v2(const StepArgs<T, AT> a) {
    step_D_v2_body<T, AT, V, LZ, EXTRAS, INTERIOR, MASK>(a, blockIdx.x);
}

template <typename T, typename AT, int V, int LZ, bool EXTRAS, bool INTERIOR, int MASK = 63>
__global__ void __launch_bounds__(32 * V2_BY, (INTERIOR ? V2_INTERIOR_MIN_CTAS : V2_D_MIN_CTAS))
k_step_D_v2_batch(const StepArgs<T, AT>* table, int B, int64_t t_probe) {
    const StepArgs<T, AT>& a = batch_args<T, AT>(table, blockIdx.x % B, t_probe);
    step_D_v2_body<T, AT, V, LZ, EXTRAS, INTERIOR, MASK>(a, blockIdx.x / B);
}


}  // namespace cev
