// TMA-staged marching half-step kernels for sm_100a (kernel_variant = 3).
//
// Same tiling and the same arithmetic as step_v2.cuh (bit-identical results), but the field planes
// are moved by the TMA engine instead of by the threads: a CTA of V3_BY warps owns V3_BY rows x 32
// vectors and marches along x; for every x-plane one elected lane per warp issues
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` copies (SASS UBLKCP) of its row of
// every stream -- with the +1/-1 halo vector along z and the halo row along y -- into a ring of
// shared-memory stages, each guarded by an mbarrier that counts the expected bytes.  Loads for plane
// i+2 are in flight while the warps compute plane i from shared memory, so HBM latency is decoupled
// from the (longer, register-hungry) PML arithmetic, and no registers are held by loads in flight.
// A __syncthreads() per plane releases the oldest stage for refilling.
#pragma once
#include "common.cuh"
#include "step_v2.cuh"

namespace cev {

#ifndef V3_BY_ROWS
#define V3_BY_ROWS 4
#endif
#ifndef V3_H_NSTAGES
#define V3_H_NSTAGES 3
#endif
#ifndef V3_D_NSTAGES
#define V3_D_NSTAGES 4
#endif
constexpr int V3_BY = V3_BY_ROWS;        // warps = rows per CTA
// planes resident per CTA: the two in use (current + x-neighbour) and the ones being loaded.  Tuned on B200:
// the D kernel (14 KB per stage) gains from a 4-deep ring, the H kernel (22 KB per stage) loses occupancy.
constexpr int V3_H_STAGES = V3_H_NSTAGES;
constexpr int V3_D_STAGES = V3_D_NSTAGES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(b)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

template <typename T, int V>
struct V3Layout {
    static constexpr int BY = V3_BY, BZ = 32 * V, ROWP = BZ + V, NSH = V3_H_STAGES, NSD = V3_D_STAGES;
    // H kernel stage: D[3], mE[3] as (BY+1) rows x ROWP (own cells then the +1 halo vector), H[3] as BY x BZ
    static constexpr int H_DM = 6 * (BY + 1) * ROWP, H_STAGE = H_DM + 3 * BY * BZ;
    // D kernel stage: H[3] as (BY+1) rows x ROWP (row 0 = halo row, -1 halo vector then own cells), D[3] as BY x BZ
    static constexpr int D_HH = 3 * (BY + 1) * ROWP, D_STAGE = D_HH + 3 * BY * BZ;
    static constexpr size_t h_bytes() { return (size_t)NSH * H_STAGE * sizeof(T) + NSH * sizeof(uint64_t); }
    static constexpr size_t d_bytes() { return (size_t)NSD * D_STAGE * sizeof(T) + NSD * sizeof(uint64_t); }
};

// common CTA prologue: which box / tile / chunk; returns false for CTAs with nothing to do
struct V3Tile {
    int jraw, j, k0t, ncell, xs, xe, y1, bx;
    bool row_on;
};
template <typename T, typename AT, int V>
__device__ __forceinline__ V3Tile v3_locate(const StepArgs<T, AT>& a, int bid, int w) {
    constexpr int BZ = 32 * V;
    V3Tile t;
    int bx = 0;
#pragma unroll
    for (int q = 1; q < MAX_BOXES; ++q)
        if (q < a.n_boxes && bid >= a.box[q].cta0) bx = q;
    const Box& B = a.box[bx];
    const int lid = bid - B.cta0;
    const int tz = lid % B.ntz;
    const int rest = lid / B.ntz;
    const int ty = rest % B.nty;
    const int xc = rest / B.nty;
    t.bx = bx;
    t.jraw = B.y0 + ty * V3_BY + w;
    t.row_on = t.jraw < B.y1;
    t.j = t.row_on ? t.jraw : B.y1 - 1;
    t.y1 = B.y1;
    t.k0t = B.z0 + tz * BZ;
    t.ncell = min(BZ, B.z1 - t.k0t);
    t.xs = B.x0 + xc * a.xchunk;
    t.xe = min(t.xs + a.xchunk, B.x1);
    return t;
}

template <typename T, typename AT, int V>
__global__ void __launch_bounds__(32 * V3_BY) k_step_H_v3(const StepArgs<T, AT> a) {
    using L = V3Layout<T, V>;
    constexpr int BY = L::BY, BZ = L::BZ, ROWP = L::ROWP, NS = L::NSH;
    extern __shared__ __align__(128) unsigned char v3_smem[];
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    T* const stage0 = reinterpret_cast<T*>(v3_smem);
    uint64_t* const full = reinterpret_cast<uint64_t*>(v3_smem + (size_t)NS * L::H_STAGE * sizeof(T));
    const int lane = threadIdx.x, w = threadIdx.y;
    const V3Tile t = v3_locate<T, AT, V>(a, bid, w);
    const int j = t.j, xs = t.xs, xe = t.xe;
    const int plane = a.Ny * a.Nz;
    const bool active = t.row_on && lane * V < t.ncell;
    const int k0 = t.k0t + (lane * V < t.ncell ? lane * V : 0);

    if (lane == 0 && w == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], BY);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nb = (uint32_t)(t.ncell * sizeof(T));
    const bool zwrap = t.k0t + t.ncell >= a.Nz;          // the +1 z-neighbour of the tile's last cell is k = 0
    const bool need_jp = (w == BY - 1) || (t.jraw + 1 >= t.y1);
    const int jp = (j + 1 == a.Ny) ? 0 : j + 1;
    // plane p into stage (p - xs) % NS.  cur_role: the plane will be a "current" plane (all components, halos, H);
    // otherwise it only serves as the x+1 neighbour of the chunk's last plane (E_y, E_z of the own cells).
    auto issue = [&](int p, bool cur_role) {
        if (lane != 0) return;
        const int s = (p - xs) % NS;
        uint64_t* bar = &full[s];
        if (!t.row_on) {
            mbar_arrive_tx(bar, 0);
            return;
        }
        T* st = stage0 + (size_t)s * L::H_STAGE;
        const int rowoff = j * a.Nz + t.k0t;
        uint32_t bytes = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!cur_role && c == 0) continue;
            const T* Dp = (p < a.Nx) ? a.Din[c] + (size_t)p * plane : a.Dhi[c];
            const T* Mp = (p < a.Nx) ? a.mE[c] + (size_t)p * plane : a.mEhi[c];
            T* dD = st + (c * (BY + 1) + w) * ROWP;
            T* dM = st + ((3 + c) * (BY + 1) + w) * ROWP;
            if (cur_role && !zwrap) {
                bulk_g2s(dD, Dp + rowoff, nb + 16, bar);
                bulk_g2s(dM, Mp + rowoff, nb + 16, bar);
                bytes += 2 * (nb + 16);
            } else {
                bulk_g2s(dD, Dp + rowoff, nb, bar);
                bulk_g2s(dM, Mp + rowoff, nb, bar);
                bytes += 2 * nb;
                if (cur_role) {
                    bulk_g2s(dD + t.ncell, Dp + j * a.Nz, 16, bar);
                    bulk_g2s(dM + t.ncell, Mp + j * a.Nz, 16, bar);
                    bytes += 32;
                }
            }
            if (cur_role && need_jp && c != 1) {           // halo row j+1 (E_x and E_z only)
                const int off = jp * a.Nz + t.k0t;
                bulk_g2s(dD + ROWP, Dp + off, nb, bar);
                bulk_g2s(dM + ROWP, Mp + off, nb, bar);
                bytes += 2 * nb;
            }
            if (cur_role) {
                bulk_g2s(st + L::H_DM + (c * BY + w) * BZ, a.Hin[c] + (size_t)p * plane + rowoff, nb, bar);
                bytes += nb;
            }
        }
        mbar_arrive_tx(bar, bytes);
    };

#pragma unroll
    for (int d = 0; d < NS - 1; ++d)
        if (xs + d <= xe) issue(xs + d, xs + d < xe);

    const int my = a.mapH[1][j];
    int mz[V];
    bool yz_pml = my >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mz[e] = a.mapH[2][k0 + e];
        yz_pml |= mz[e] >= 0;
    }
    const AT s = -a.cdt;
    const AT inv = a.inv_dL;
    const int orow = j * a.Nz + k0;

    for (int i = xs; i < xe; ++i) {
        if (i + NS - 1 <= xe) issue(i + NS - 1, i + NS - 1 < xe);
        const int q = i - xs;
        // PML integrals go through the LSU: issue them (and the L2 prefetch of the next plane's) before
        // blocking on the TMA barriers so that their latency overlaps the wait
        const int mx = a.mapH[0][i];
        const bool pml = yz_pml || mx >= 0;
        PmlCtx<T, AT, V, true> ctx;
        if (pml && active) ctx.load(a, i, j, k0, mx, my, mz, 7u);
        if (active && i + 1 < a.x1) PmlCtx<T, AT, V, true>::prefetch(a, i + 1, j, k0, my, mz, (lane & 7) == 0);
        mbar_wait(&full[q % NS], (q / NS) & 1);
        mbar_wait(&full[(q + 1) % NS], ((q + 1) / NS) & 1);
        const T* cur = stage0 + (size_t)(q % NS) * L::H_STAGE;
        const T* nxt = stage0 + (size_t)((q + 1) % NS) * L::H_STAGE;
        const int col = lane * V;
        const int pbase = i * plane;

        AT E[3][V], CE[3][V];
        Vec<T, V> h[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Vec<T, V> d = ldv<T, V>(cur + (c * (BY + 1) + w) * ROWP + col);
            const Vec<T, V> m = ldv<T, V>(cur + ((3 + c) * (BY + 1) + w) * ROWP + col);
            h[c] = ldv<T, V>(cur + L::H_DM + (c * BY + w) * BZ + col);
#pragma unroll
            for (int e = 0; e < V; ++e) E[c][e] = mul_rn((AT)m.v[e], (AT)d.v[e]);
        }
        const Vec<T, V> dxj = ldv<T, V>(cur + (0 * (BY + 1) + w + 1) * ROWP + col);
        const Vec<T, V> mxj = ldv<T, V>(cur + (3 * (BY + 1) + w + 1) * ROWP + col);
        const Vec<T, V> dzj = ldv<T, V>(cur + (2 * (BY + 1) + w + 1) * ROWP + col);
        const Vec<T, V> mzj = ldv<T, V>(cur + (5 * (BY + 1) + w + 1) * ROWP + col);
        const Vec<T, V> dyn = ldv<T, V>(nxt + (1 * (BY + 1) + w) * ROWP + col);
        const Vec<T, V> myn = ldv<T, V>(nxt + (4 * (BY + 1) + w) * ROWP + col);
        const Vec<T, V> dzn = ldv<T, V>(nxt + (2 * (BY + 1) + w) * ROWP + col);
        const Vec<T, V> mzn = ldv<T, V>(nxt + (5 * (BY + 1) + w) * ROWP + col);
        const AT ex_kp = mul_rn((AT)cur[(3 * (BY + 1) + w) * ROWP + col + V], (AT)cur[(0 * (BY + 1) + w) * ROWP + col + V]);
        const AT ey_kp = mul_rn((AT)cur[(4 * (BY + 1) + w) * ROWP + col + V], (AT)cur[(1 * (BY + 1) + w) * ROWP + col + V]);
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
            if (!pml) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e) out[c].v[e] = (T)muladd(s, CE[c][e], (AT)h[c].v[e]);
            } else {
                ctx.apply(a, i, j, k0, mx, my, mz, s, h, CE, out, 7u);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Hout[c] + pbase + orow, out[c]);
        }
        __syncthreads();     // every warp is done with the oldest stage: it may be refilled
    }
}

template <typename T, typename AT, int V>
__global__ void __launch_bounds__(32 * V3_BY) k_step_D_v3(const StepArgs<T, AT> a) {
    using L = V3Layout<T, V>;
    constexpr int BY = L::BY, BZ = L::BZ, ROWP = L::ROWP, NS = L::NSD;
    extern __shared__ __align__(128) unsigned char v3_smem[];
    const int bid = blockIdx.x;
    if (bid >= a.n_tiles) {
        probe_block<T, AT>(a, a.aux_slot0 + (bid - a.n_tiles));
        return;
    }
    T* const stage0 = reinterpret_cast<T*>(v3_smem);
    uint64_t* const full = reinterpret_cast<uint64_t*>(v3_smem + (size_t)NS * L::D_STAGE * sizeof(T));
    const int lane = threadIdx.x, w = threadIdx.y;
    const V3Tile t = v3_locate<T, AT, V>(a, bid, w);
    const int j = t.j, xs = t.xs, xe = t.xe;
    const int plane = a.Ny * a.Nz;
    const bool active = t.row_on && lane * V < t.ncell;
    const int k0 = t.k0t + (lane * V < t.ncell ? lane * V : 0);

    if (lane == 0 && w == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], BY);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nb = (uint32_t)(t.ncell * sizeof(T));
    const int jm = (j == 0) ? a.Ny - 1 : j - 1;
    // plane p (p >= xs-1) into stage (p - xs + 1) % NS.  cur_role false: only H_y, H_z of the own cells (x-1 neighbour)
    auto issue = [&](int p, bool cur_role) {
        if (lane != 0) return;
        const int sidx = (p - xs + 1) % NS;
        uint64_t* bar = &full[sidx];
        if (!t.row_on) {
            mbar_arrive_tx(bar, 0);
            return;
        }
        T* st = stage0 + (size_t)sidx * L::D_STAGE;
        const int rowoff = j * a.Nz + t.k0t;
        uint32_t bytes = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!cur_role && c == 0) continue;
            const T* Hp = (p >= 0) ? a.Hin[c] + (size_t)p * plane : a.Hlo[c];
            T* dH = st + (c * (BY + 1) + w + 1) * ROWP;      // own row lives at smem row w+1; columns: [-1 vector | own]
            if (cur_role && t.k0t > 0) {
                bulk_g2s(dH, Hp + rowoff - V, nb + 16, bar);
                bytes += nb + 16;
            } else {
                bulk_g2s(dH + V, Hp + rowoff, nb, bar);
                bytes += nb;
                if (cur_role) {                              // k = -1 wraps to the end of the row
                    bulk_g2s(dH, Hp + j * a.Nz + a.Nz - V, 16, bar);
                    bytes += 16;
                }
            }
            if (cur_role && w == 0 && c != 1) {              // halo row j-1 (H_x and H_z only)
                bulk_g2s(st + (c * (BY + 1)) * ROWP + V, Hp + jm * a.Nz + t.k0t, nb, bar);
                bytes += nb;
            }
            if (cur_role) {
                bulk_g2s(st + L::D_HH + (c * BY + w) * BZ, a.Din[c] + (size_t)p * plane + rowoff, nb, bar);
                bytes += nb;
            }
        }
        mbar_arrive_tx(bar, bytes);
    };

    issue(xs - 1, false);
#pragma unroll
    for (int d = 0; d < NS - 2; ++d)
        if (xs + d < xe) issue(xs + d, true);

    const int my = a.mapD[1][j];
    int mz[V];
    bool yz_pml = my >= 0;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        mz[e] = a.mapD[2][k0 + e];
        yz_pml |= mz[e] >= 0;
    }
    const AT s = a.cdt;
    const AT inv = a.inv_dL;
    const int orow = j * a.Nz + k0;

    for (int i = xs; i < xe; ++i) {
        if (i + NS - 2 < xe) issue(i + NS - 2, true);
        const int q = i - xs;                      // prev plane is use #q, current plane use #(q+1) of the ring
        const int mx = a.mapD[0][i];
        const bool pml = yz_pml || mx >= 0;
        PmlCtx<T, AT, V, false> ctx;
        if (pml && active) ctx.load(a, i, j, k0, mx, my, mz, 7u);
        if (active && i + 1 < a.x1) PmlCtx<T, AT, V, false>::prefetch(a, i + 1, j, k0, my, mz, (lane & 7) == 0);
        mbar_wait(&full[q % NS], (q / NS) & 1);
        mbar_wait(&full[(q + 1) % NS], ((q + 1) / NS) & 1);
        const T* prv = stage0 + (size_t)(q % NS) * L::D_STAGE;
        const T* cur = stage0 + (size_t)((q + 1) % NS) * L::D_STAGE;
        const int col = V + lane * V;
        const int pbase = i * plane;

        Vec<T, V> h[3], d[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            h[c] = ldv<T, V>(cur + (c * (BY + 1) + w + 1) * ROWP + col);
            d[c] = ldv<T, V>(cur + L::D_HH + (c * BY + w) * BZ + lane * V);
        }
        const Vec<T, V> hxj = ldv<T, V>(cur + (0 * (BY + 1) + w) * ROWP + col);
        const Vec<T, V> hzj = ldv<T, V>(cur + (2 * (BY + 1) + w) * ROWP + col);
        const Vec<T, V> hyp = ldv<T, V>(prv + (1 * (BY + 1) + w + 1) * ROWP + col);
        const Vec<T, V> hzp = ldv<T, V>(prv + (2 * (BY + 1) + w + 1) * ROWP + col);
        const AT hx_km = (AT)cur[(0 * (BY + 1) + w + 1) * ROWP + col - 1];
        const AT hy_km = (AT)cur[(1 * (BY + 1) + w + 1) * ROWP + col - 1];

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
            if (!pml) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int e = 0; e < V; ++e) out[c].v[e] = (T)muladd(s, CH[c][e], (AT)d[c].v[e]);
            } else {
                ctx.apply(a, i, j, k0, mx, my, mz, s, d, CH, out, 7u);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) stv<T, V>(a.Dout[c] + pbase + orow, out[c]);
        }
        __syncthreads();
    }

    // ---- in-kernel source injection: D += J after the update (fdtd.py:125-127)
    if (a.src_wave) {
        inject_points<T, AT, int32_t>(a.src_begin[bid], a.src_begin[bid + 1], threadIdx.y * 32 + threadIdx.x, 32 * V3_BY, a.src_comp,
                                      a.src_id, a.src_cell, a.src_w, a.src_wave, a.Dout[0], a.Dout[1], a.Dout[2]);
    }
}

}  // namespace cev
