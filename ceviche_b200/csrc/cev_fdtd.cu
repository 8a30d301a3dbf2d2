// ceviche_b200: host side of the C ABI (include/ceviche_b200.h) + kernel launches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/ceviche_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "common.cuh"
#include "step_v1.cuh"
#include "step_v2.cuh"
#include "step_v3.cuh"
#include "step_v4.cuh"
#include "step_v5.h"
#include "adjoint_v5.h"
#include "tan2d_fused.cuh"
#include "adjoint.cuh"

namespace {

// ceviche/constants.py:7-9 -- the reference's (non-SI) values; they set dt, sigma and every coefficient.
constexpr double EPSILON_0 = 8.85418782e-12;
constexpr double MU_0 = 1.25663706e-6;

thread_local std::string g_err;

int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}

#define CUDA_TRY(expr)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (expr);                                                        \
        if (e_ != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

constexpr int PROBE_CHUNK = 8192;   // points per probe slot (one CTA each)

struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DeviceBuf() = default;
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
    ~DeviceBuf() { release(); }          // (owners are destroyed with their device current: cev_fdtd_destroy / create)
    int alloc(size_t n) {
        release();
        if (n == 0) return 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) return fail("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
        bytes = n;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

}  // namespace

struct cev_fdtd {
    int device = 0, dtype = CEV_F64, arith64 = 1;
    // Internal axis A holds logical axis perm[A] (inv = the inverse map).  Identity for 3-D grids; a 2-D grid
    // (Nz = 1) is relabelled (x, z, y) so that x stays the marching axis and y becomes the vectorised one -- a swap
    // of two axes, under which the curl changes sign (parity = -1: C0*dt enters every kernel negated, which is
    // exact); a 1-D grid (Ny = Nz = 1) is relabelled cyclically (y, z, x).  Memory is untouched in both cases.
    int perm[3] = {0, 1, 2}, inv[3] = {0, 1, 2};
    int parity = 1;
    int64_t Nl[3] = {0, 0, 0};   // logical extents
    int N[3] = {0, 0, 0};        // internal extents
    double dL = 0, dt = 0, cdt = 0;
    int nH[3] = {0, 0, 0}, nD[3] = {0, 0, 0};   // internal compact counts
    int variant = 0;             // 0 auto, 1 baseline kernels, 2 marching kernels, 3 TMA-staged marching kernels,
                                 // 4 fused full-step kernel wherever it applies (cev_fdtd_run_fused), 5 hybrid,
                                 // 6 tensor-map TMA kernels (step_v5.cuh) wherever they apply
    int tma_rows = 4, tma_stages_H = 3, tma_stages_D = 4;   // tile rows / ring depths of the tensor-map kernels
    int auto_v5 = 1;                    // kernel_variant 0 (auto) picks them on large 3-D grids
    int adjoint_variant = 0;            // reverse sweep of cev_fdtd_adjoint_run: 0 auto, 1 simple kernels (adjoint.cuh),
                                        // 2 tensor-map kernels (adjoint_v5.cuh) wherever they apply
    int tma_stages_adjH = 3, tma_stages_adjED = 3;     // (profiles/r2_tune_adjoint_ring_depths.log)
    // x-slab halo exchange through peer-mapped memory (cev_fdtd_halo_attach): this slab's exchange block and the
    // neighbours' (device pointers valid on this device), and how many H / D half-steps have used them
    struct Halo {
        unsigned char* own = nullptr;
        unsigned char* left = nullptr;
        unsigned char* right = nullptr;
        uint64_t nH = 0, nD = 0;
        bool paused = false;        // option "halo_pause": launches ignore the blocks (periodic wrap inside the slab)
        bool on() const { return own && !paused; }
    } halo;
    cev::V5MapCache* v5 = nullptr;      // CUtensorMap descriptors of this plan's arrays
    // D-box recorder (cev_fdtd_set_recorder): cev_fdtd_run stores D of the box after every step into slot count++
    struct Recorder {
        void* buf = nullptr;
        int64_t capacity = 0, count = 0;
        int box[6] = {0, 0, 0, 0, 0, 0};     // internal axes
    } rec;
    int fused_shape = 0;         // tile shape of the fused kernel: 0 auto, else LZ*100 + BY
    int xchunk = 0;              // 0 auto
    int pf_dist = 1;             // L2 prefetch distance of the marching kernels (planes)
    int lz = 8;                  // lanes of a warp along z in the marching kernels (8, 16 or 32)
    unsigned on = 63u;           // internal component mask (bits 0-2 E/D, 3-5 H): see StepArgs::on
    bool smem_attr_H = false, smem_attr_D = false;   // dynamic-smem opt-in done for this plan's device / dtype
    int split = 0;               // 1: separate launches for the PML-free interior box and the PML shell
    int in_lo[3] = {0, 0, 0}, in_hi[3] = {0, 0, 0};   // per internal axis: longest index run off the PML (H and D sampling)
    DeviceBuf tables;            // u/r (f32 + f64) and maps for 3 axes x {H, D}
    const void* uH[3][2];        // [axis][0: f32, 1: f64]
    const void* rH[3][2];
    const void* uD[3][2];
    const void* rD[3][2];
    const int* mapH[3];
    const int* mapD[3];
    // sources
    int nsrc = 0;
    int64_t n_src_pts = 0;
    DeviceBuf src_comp, src_id, src_cell, src_weight;
    std::vector<int32_t> h_src_comp, h_src_id;      // host copies: re-sorted per kernel tiling
    std::vector<int64_t> h_src_cell;
    std::vector<double> h_src_w;
    struct SrcTiling {                              // source points sorted by owning CTA of one launch geometry
        int x0, x1, xchunk, lz, vec, part, rows, xorder;
        DeviceBuf begin, comp, id, cell, w;
    };
    std::vector<std::unique_ptr<SrcTiling>> src_tilings;
    // probes
    int nprobe = 0, n_slots = 0, n_slots_ED = 0;
    std::vector<int32_t> slot_probe;
    DeviceBuf pr_field, pr_wbegin, pr_ibegin, pr_cell0, pr_n, pr_idx, pr_weight, pr_owner;
    // running-DFT monitors
    int64_t n_mon_pts = 0;
    int mon_nfreq = 0;
    DeviceBuf mon_field, mon_cell;
    const double* mon_phasors = nullptr;   // bound per run: [nsteps, nfreq, 2]
    double* mon_acc = nullptr;             // [n_mon_pts, nfreq, 2]

    // CUDA graphs of the caller loop for launch-bound (small) grids: blocks of GRAPH_K time steps captured once
    // per (state pointers, sources/probes/options epoch) and replayed; waveform rows and probe partial sums go
    // through plan-owned staging buffers so that the captured launches are completely static.
    struct RunGraph {
        cudaGraphExec_t exec = nullptr;
        cev_state st;
        uint64_t epoch = 0;
    };
    std::vector<RunGraph> graphs;
    DeviceBuf stage_w, stage_p;
    // One checkpoint segment of the reverse sweep (recomputation + transposed steps) as a CUDA graph, for grids whose
    // kernels are shorter than a launch: captured the second time the same segment (same arrays, same length) is asked
    // for and replayed from then on; waveform / cotangent rows go through the staging buffers.
    struct AdjGraph {
        cudaGraphExec_t exec = nullptr;
        std::vector<unsigned char> key, seen;       // key of the captured graph / of the last plain segment
        uint64_t epoch = 0;
        int64_t replays = 0;
    } adj_graph;
    DeviceBuf adj_stage_w, adj_stage_g;
    cudaStream_t cap_stream = nullptr;
    std::vector<cudaStream_t> side;        // forward mode: one side stream (+ event) per tangent state
    std::vector<cudaEvent_t> side_ev;
    cudaEvent_t main_ev = nullptr;
    int jvp_batch = -1;          // -1 auto / 1: the B tangent half-steps of a forward-mode sweep as ONE launch each
                                 // (k_step_{H,D}_v2_batch) wherever the marching kernels serve them; 0: one launch per tangent
    DeviceBuf batch_tab_H, batch_tab_D;     // per-tangent StepArgs of those launches
    int jvp_fused = -1;          // -1 auto / 1: 2-D TM tangent states advance by the fused full-step kernel (tan2d_fused.cuh),
                                 // ping-ponged between the caller's arrays and tan_shadow; 0: the batched two-kernel path
    DeviceBuf batch_tab_F0, batch_tab_F1, tan_shadow;
    int jvp_streams = -1;        // -1 auto (2-D grids <= 2^23 cells, any grid <= 2^20 cells), 0 never, 1 always
    uint64_t epoch = 1;          // bumped whenever sources, probes or options change
    int use_graph = -1;          // -1 auto (small grids), 0 never, 1 whenever possible
    void drop_graphs() {
        for (auto& g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
        if (adj_graph.exec) cudaGraphExecDestroy(adj_graph.exec);
        adj_graph = AdjGraph();
    }

    int to_internal(int logical_axis) const { return inv[logical_axis]; }
    int to_logical(int internal_axis) const { return perm[internal_axis]; }
    bool x_is_x() const { return perm[0] == 0; }     // logical x-ranges / x-halo planes are internal ones
};

namespace {

using namespace cev;

void halo_layout(const cev_fdtd* p, cev_halo_layout* lay) {
    const size_t es = p->dtype == CEV_F64 ? 8 : 4;
    const size_t pb = ((size_t)p->Nl[1] * p->Nl[2] * es + 255) / 256 * 256;
    for (int c = 0; c < 2; ++c) {
        lay->D_hi[c] = (0 + c) * pb;
        lay->inv_eps_hi[c] = (2 + c) * pb;
        lay->H_lo[c] = (4 + c) * pb;
    }
    lay->flag_D = 6 * pb;
    lay->flag_H = 6 * pb + 128;
    lay->err = 6 * pb + 256;
    lay->plane_bytes = pb;
    lay->bytes = 6 * pb + 512;
}

void fill_probe_table(const cev_fdtd* p, ProbeTable& pr) {
    pr.n_slots = p->n_slots;
    pr.slot_field = (const int32_t*)p->pr_field.p;
    pr.slot_wbegin = (const int64_t*)p->pr_wbegin.p;
    pr.slot_ibegin = (const int64_t*)p->pr_ibegin.p;
    pr.slot_cell0 = (const int64_t*)p->pr_cell0.p;
    pr.slot_n = (const int64_t*)p->pr_n.p;
    pr.idx = (const int64_t*)p->pr_idx.p;
    pr.weight = (const double*)p->pr_weight.p;
}

template <typename T, typename AT>
int fill_args(const cev_fdtd* p, const cev_state* st, StepArgs<T, AT>& a, const cev_tangent* tan = nullptr) {
    memset(&a, 0, sizeof a);
    a.Nx = p->N[0];
    a.Ny = p->N[1];
    a.Nz = p->N[2];
    const int64_t plane = (int64_t)a.Ny * a.Nz;
    constexpr int w = sizeof(AT) == 8 ? 1 : 0;
    for (int A = 0; A < 3; ++A) {
        const int L = p->to_logical(A);
        a.Hin[A] = (const T*)st->H[L];
        a.Hout[A] = (T*)st->H[L];
        a.Din[A] = (const T*)st->D[L];
        a.Dout[A] = (T*)st->D[L];
        a.mE[A] = (const T*)st->inv_eps[L];
        if (!a.Hin[A] || !a.Din[A] || !a.mE[A]) return fail("cev_state: H, D and inv_eps must be non-NULL");
        a.ICE[A] = (T*)st->ICE[L];
        a.IH[A] = (T*)st->IH[L];
        a.ICH[A] = (T*)st->ICH[L];
        a.ID[A] = (T*)st->ID[L];
        a.Dhi[A] = st->D_xhi[L] ? (const T*)st->D_xhi[L] : a.Din[A];
        a.mEhi[A] = st->inv_eps_xhi[L] ? (const T*)st->inv_eps_xhi[L] : a.mE[A];
        a.Hlo[A] = st->H_xlo[L] ? (const T*)st->H_xlo[L] : a.Hin[A] + (int64_t)(a.Nx - 1) * plane;
        a.mapH[A] = p->mapH[A];
        a.mapD[A] = p->mapD[A];
        a.nH[A] = p->nH[A];
        a.nD[A] = p->nD[A];
        a.uH[A] = (const AT*)p->uH[A][w];
        a.rH[A] = (const AT*)p->rH[A][w];
        a.uD[A] = (const AT*)p->uD[A][w];
        a.rD[A] = (const AT*)p->rD[A][w];
        a.Jscale[A] = AT(1);
        if (tan) {
            a.dmE[A] = (const T*)tan->d_inv_eps[L];
            a.Dp[A] = (const T*)tan->D_primal[L];
            if (!a.dmE[A] || !a.Dp[A]) return fail("cev_tangent: d_inv_eps and D_primal must be non-NULL");
            a.dmEhi[A] = a.dmE[A];     // tangents run on whole (periodic) grids only
            a.Dphi[A] = a.Dp[A];
        }
    }
    if (p->halo.on()) {      // attached exchange block: the halo planes this slab reads live there
        if (tan) return fail("tangent steps do not support x-slab halos");
        cev_halo_layout lay;
        halo_layout(p, &lay);
        for (int c = 1; c < 3; ++c) {
            a.Dhi[c] = (const T*)(p->halo.own + lay.D_hi[c - 1]);
            a.mEhi[c] = (const T*)(p->halo.own + lay.inv_eps_hi[c - 1]);
            a.Hlo[c] = (const T*)(p->halo.own + lay.H_lo[c - 1]);
        }
    }
    if (tan && (st->D_xhi[1] || st->D_xhi[2])) return fail("tangent steps do not support x-halo planes");
    if (!p->x_is_x() && (st->D_xhi[1] || st->D_xhi[2] || st->H_xlo[1] || st->H_xlo[2]))
        return fail("x-halo planes need Ny > 1 or Nz > 1 (no slab decomposition of a 1-D grid)");
    // PML integral arrays must exist wherever the kernels will touch them
    for (int A = 0; A < 3; ++A) {
        const int B = (A + 1) % 3, C = (A + 2) % 3;
        if (p->nH[A] > 0 && !a.ICE[A]) return fail("cev_state: ICE missing for a PML axis");
        if (p->nD[A] > 0 && !a.ICH[A]) return fail("cev_state: ICH missing for a PML axis");
        if (p->nH[B] > 0 && p->nH[C] > 0 && !a.IH[A]) return fail("cev_state: IH missing for a PML corner");
        if (p->nD[B] > 0 && p->nD[C] > 0 && !a.ID[A]) return fail("cev_state: ID missing for a PML corner");
    }
    a.cdt = (AT)(p->parity * p->cdt);
    a.inv_dL = (AT)(1.0 / p->dL);
    a.on = p->on;
    fill_probe_table(p, a.pr);
    a.t_probe = -1;
    return 0;
}

// block shape of the baseline kernels: 64 x 4 cells, or one row of 256 for single-row planes (2-D grids)
template <typename T, typename AT>
dim3 v1_block(const StepArgs<T, AT>& a) {
    return a.Ny == 1 ? dim3(V1_TZ * V1_TY, 1) : dim3(V1_TZ, V1_TY);
}

template <typename T, typename AT>
void set_tiles_v1(StepArgs<T, AT>& a, int64_t x0, int64_t x1) {
    a.x0 = (int)x0;
    a.x1 = (int)x1;
    a.xchunk = 1;
    const dim3 blk = v1_block(a);
    a.ntz = (a.Nz + blk.x - 1) / blk.x;
    a.nty = (a.Ny + blk.y - 1) / blk.y;
    a.n_tiles = a.ntz * a.nty * (int)(x1 - x0);
}

template <typename T>
constexpr int vec_width() {
    return 16 / (int)sizeof(T);
}

// The marching kernels need 16-byte vectors along z: Nz a multiple of the vector width and every
// array 16-byte aligned (torch allocations are; odd Nz falls back to the baseline kernels).
template <typename T, typename AT>
bool can_march(const cev_fdtd* p, const StepArgs<T, AT>& a, bool isH) {
    if (p->variant == 1) return false;
    constexpr int V = vec_width<T>();
    if (a.Nz % V != 0) return false;
    auto ok = [](const void* q) { return q == nullptr || ((uintptr_t)q % 16) == 0; };
    for (int c = 0; c < 3; ++c) {
        if (!ok(a.Hin[c]) || !ok(a.Hout[c]) || !ok(a.Din[c]) || !ok(a.Dout[c]) || !ok(a.mE[c]) || !ok(a.Eout[c]) ||
            !ok(a.Dhi[c]) || !ok(a.mEhi[c]) || !ok(a.Hlo[c]) || !ok(a.J[c]) || !ok(a.dmE[c]) || !ok(a.Dp[c]) ||
            !ok(a.dmEhi[c]) || !ok(a.Dphi[c]))
            return false;
    }
    (void)isH;
    return p->variant >= 2 || a.Nz >= 2 * V;
}

// Tiling of one marching launch over a list of boxes.  part: 0 = the whole range [x0,x1) x Ny x Nz in one
// box (general kernel); 1 = the PML-free interior box clipped to [x0,x1); 2 = the shell = the rest, as up to six slabs.
template <typename T, typename AT>
void set_tiles_v2(const cev_fdtd* p, StepArgs<T, AT>& a, int64_t x0, int64_t x1, int part, int lz_override = 0,
                  int lz_force = 0, const int* inner = nullptr /* {x0,x1,y0,y1,z0,z1}: overrides the interior box */) {
    constexpr int V = vec_width<T>();
    // planes of a single row (2-D grids: internal Ny = 1): one row per warp and the warps side by side along z
    const bool wz = a.Ny == 1 && part == 0 && !lz_override;
    const int LZ = lz_override ? lz_override : (wz ? 32 : (lz_force ? lz_force : p->lz));
    const int rows = wz ? 1 : V2_BY * (32 / LZ), zc = (wz ? V2_BY : 1) * LZ * V;   // cells of a CTA tile
    a.wz = wz ? 1 : 0;
    a.x0 = (int)x0;
    a.x1 = (int)x1;
    int chunk = p->xchunk;
    if (chunk <= 0) {
        // short chunks keep the concurrently-active working set (CTAs x streams x planes) inside L2
        // and give the scheduler many CTAs to balance; tuned on B200 (scripts/tune.py)
        const int cols = ((a.Nz + zc - 1) / zc) * ((a.Ny + rows - 1) / rows);
        chunk = cols >= 512 ? 8 : 4;
        if (lz_override) chunk = 16;      // TMA-staged kernels: longer chunks amortise the pipeline fill
        if (wz) {                         // a plane is one row: longer chunks amortise the carried-in plane
            chunk = 32;
            while (chunk > 1 && (int64_t)cols * ((x1 - x0 + chunk - 1) / chunk) < 148 * 4) chunk /= 2;   // tuned: scripts/tune2d.py
        }
    }
    a.xchunk = chunk;
    a.pf_dist = p->pf_dist;
    a.n_boxes = 0;
    int cta = 0;
    auto add = [&](int bx0, int bx1, int by0, int by1, int bz0, int bz1) {
        if (bx1 <= bx0 || by1 <= by0 || bz1 <= bz0) return;
        Box& B = a.box[a.n_boxes++];
        B.x0 = bx0; B.x1 = bx1; B.y0 = by0; B.y1 = by1; B.z0 = bz0; B.z1 = bz1;
        B.ntz = (bz1 - bz0 + zc - 1) / zc;
        B.nty = (by1 - by0 + rows - 1) / rows;
        B.cta0 = cta;
        cta += B.ntz * B.nty * ((bx1 - bx0 + chunk - 1) / chunk);
    };
    const int X0 = (int)x0, X1 = (int)x1;
    // interior box (z limits rounded inwards to the vector width)
    const int ix0 = std::max(X0, inner ? inner[0] : p->in_lo[0]), ix1 = std::min(X1, inner ? inner[1] : p->in_hi[0]);
    const int iy0 = inner ? inner[2] : p->in_lo[1], iy1 = inner ? inner[3] : p->in_hi[1];
    const int iz0 = inner ? inner[4] : (p->in_lo[2] + V - 1) / V * V, iz1 = inner ? inner[5] : p->in_hi[2] / V * V;
    if (part == 0) {
        add(X0, X1, 0, a.Ny, 0, a.Nz);
    } else if (part == 1) {
        add(ix0, ix1, iy0, iy1, iz0, iz1);
    } else {
        add(X0, std::min(ix0, X1), 0, a.Ny, 0, a.Nz);          // x-low slab (or everything if no interior x here)
        add(std::max(ix1, std::min(ix0, X1)), X1, 0, a.Ny, 0, a.Nz);   // x-high slab
        add(ix0, ix1, 0, iy0, 0, a.Nz);                        // y-low
        add(ix0, ix1, iy1, a.Ny, 0, a.Nz);                     // y-high
        add(ix0, ix1, iy0, iy1, 0, iz0);                       // z-low
        add(ix0, ix1, iy0, iy1, iz1, a.Nz);                    // z-high
    }
    a.n_tiles = cta;
    a.ntz = a.n_boxes ? a.box[0].ntz : 1;
    a.nty = a.n_boxes ? a.box[0].nty : 1;
}

// TMA-staged kernels: rows of 32 vectors, V3_BY rows per CTA
template <typename T, typename AT>
void set_tiles_v3(const cev_fdtd* p, StepArgs<T, AT>& a, int64_t x0, int64_t x1) {
    constexpr int V = vec_width<T>();
    set_tiles_v2(p, a, x0, x1, 0, 32);
    // redo the CTA counts with V3_BY rows per CTA
    int cta = 0;
    for (int b = 0; b < a.n_boxes; ++b) {
        Box& B = a.box[b];
        B.ntz = (B.z1 - B.z0 + 32 * V - 1) / (32 * V);
        B.nty = (B.y1 - B.y0 + V3_BY - 1) / V3_BY;
        B.cta0 = cta;
        cta += B.ntz * B.nty * ((B.x1 - B.x0 + a.xchunk - 1) / a.xchunk);
    }
    a.n_tiles = cta;
}

// Is it worth (and possible) to split [x0,x1) into an interior launch and a shell launch?
template <typename T>
bool want_split(const cev_fdtd* p, int64_t x0, int64_t x1) {
    if (!p->split || p->N[1] == 1) return false;
    constexpr int V = vec_width<T>();
    const int64_t ix = std::min<int64_t>(x1, p->in_hi[0]) - std::max<int64_t>(x0, p->in_lo[0]);
    const int64_t iy = p->in_hi[1] - p->in_lo[1];
    const int64_t iz = p->in_hi[2] / V * V - (p->in_lo[2] + V - 1) / V * V;
    if (ix <= 0 || iy <= 0 || iz <= 0) return false;
    const int64_t all = (x1 - x0) * p->N[1] * p->N[2];
    const int64_t inner = ix * iy * iz;
    return inner != all && inner * 8 >= all && inner >= 32768;   // some PML, a worthwhile interior
}

// Source points of the x-planes [a.x0, a.x1), sorted by the CTA of the marching D kernel that owns
// their cell (built once per launch geometry, cached in the plan).
template <typename T, typename AT>
int attach_sources_v2(cev_fdtd* p, StepArgs<T, AT>& a, const double* wave_row, int part, int lz_override = 0,
                      int rows_override = 0) {
    // lz_override < 0: fused kernel, a CTA owns rows_override rows x (-lz_override) vectors
    constexpr int V = vec_width<T>();
    const int LZ = lz_override ? lz_override : p->lz;
    cev_fdtd::SrcTiling* hit = nullptr;
    const int rows = rows_override ? rows_override : V2_BY * (32 / LZ);
    for (auto& t : p->src_tilings)
        if (t->x0 == a.x0 && t->x1 == a.x1 && t->xchunk == a.xchunk && t->lz == LZ && t->vec == V && t->part == part &&
            t->rows == rows && t->xorder == a.xorder)
            hit = t.get();
    if (!hit) {
        const int zcells = (LZ < 0 ? -LZ : LZ) * V;
        const int64_t plane = (int64_t)a.Ny * a.Nz;
        std::vector<int> owner;
        std::vector<int64_t> pick;
        for (int64_t q = 0; q < p->n_src_pts; ++q) {
            const int64_t cell = p->h_src_cell[q];
            const int i = (int)(cell / plane), j = (int)((cell % plane) / a.Nz), k = (int)(cell % a.Nz);
            for (int b = 0; b < a.n_boxes; ++b) {
                const Box& B = a.box[b];
                if (i < B.x0 || i >= B.x1 || j < B.y0 || j >= B.y1 || k < B.z0 || k >= B.z1) continue;
                int xc = (i - B.x0) / a.xchunk;
                if (a.xorder) xc = v5_rank_of_chunk(xc, (B.x1 - B.x0 + a.xchunk - 1) / a.xchunk);
                owner.push_back(B.cta0 + (xc * B.nty + (j - B.y0) / rows) * B.ntz + (k - B.z0) / zcells);
                pick.push_back(q);
                break;
            }
        }
        std::vector<int64_t> order(pick.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return owner[x] < owner[y]; });
        const int64_t m = (int64_t)pick.size();
        std::vector<int> begin(a.n_tiles + 1, 0);
        std::vector<int32_t> comp(m), id(m), cell(m);
        std::vector<double> w(m);
        for (int64_t r = 0; r < m; ++r) {
            const int64_t q = pick[order[r]];
            begin[owner[order[r]] + 1]++;
            comp[r] = p->h_src_comp[q];
            id[r] = p->h_src_id[q];
            cell[r] = (int32_t)p->h_src_cell[q];
            w[r] = p->h_src_w[q];
        }
        for (int b = 0; b < a.n_tiles; ++b) begin[b + 1] += begin[b];
        std::unique_ptr<cev_fdtd::SrcTiling> t(new cev_fdtd::SrcTiling());
        t->x0 = a.x0; t->x1 = a.x1; t->xchunk = a.xchunk; t->lz = LZ; t->vec = V; t->part = part; t->rows = rows;
        t->xorder = a.xorder;
        const size_t mm = (size_t)(m > 0 ? m : 1);
        if (t->begin.alloc(begin.size() * 4) || t->comp.alloc(mm * 4) || t->id.alloc(mm * 4) || t->cell.alloc(mm * 4) ||
            t->w.alloc(mm * 8))
            return -1;
        CUDA_TRY(cudaMemcpy(t->begin.p, begin.data(), begin.size() * 4, cudaMemcpyHostToDevice));
        if (m) {
            CUDA_TRY(cudaMemcpy(t->comp.p, comp.data(), m * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(t->id.p, id.data(), m * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(t->cell.p, cell.data(), m * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaMemcpy(t->w.p, w.data(), m * 8, cudaMemcpyHostToDevice));
        }
        if (p->src_tilings.size() >= 16) p->src_tilings.erase(p->src_tilings.begin());
        p->src_tilings.push_back(std::move(t));
        hit = p->src_tilings.back().get();
    }
    a.src_begin = (const int*)hit->begin.p;
    a.src_comp = (const int32_t*)hit->comp.p;
    a.src_id = (const int32_t*)hit->id.p;
    a.src_cell = (const int32_t*)hit->cell.p;
    a.src_w = (const double*)hit->w.p;
    a.src_wave = wave_row;
    return 0;
}

// which: 0 = E/D-family slots, 1 = H-family slots
template <typename T, typename AT>
int attach_probes(const cev_fdtd* p, StepArgs<T, AT>& a, int which, int64_t t, double* partials) {
    if (t < 0 || !partials || p->n_slots == 0) return 0;
    a.aux_slot0 = which == 0 ? 0 : p->n_slots_ED;
    a.t_probe = t;
    a.partials = partials;
    return which == 0 ? p->n_slots_ED : p->n_slots - p->n_slots_ED;
}

// The tensor-map TMA kernels (step_v5.cuh) serve whole y-z planes of 3-D grids with all six components live.
template <typename T, typename AT>
bool want_v5(cev_fdtd* p, const StepArgs<T, AT>& a, int64_t x0, int64_t x1, bool isH) {
    const bool forced = p->variant == 6 || p->halo.on();
    if (!forced && (p->variant != 0 || !p->auto_v5)) return false;
    if (x1 <= x0 || !v5_eligible<T, AT>(a, p->tma_rows)) return false;
    if (!forced && (int64_t)a.Ny * a.Nz < (1 << 14)) return false;     // small planes: too few CTAs per chunk
    // measured on B200 (profiles/r2_tune_tensor_map.log): the fp32 H half-step of very large grids is the one case
    // where the register-marching kernel is still ahead (6.29 vs 5.94 TB/s at 512^3)
    if (!forced && isH && sizeof(T) == 4 && (int64_t)a.Nx * a.Ny * a.Nz >= ((int64_t)1 << 26)) return false;
    if (!p->v5) p->v5 = v5_cache_create();
    return true;
}

template <typename T, typename AT>
int launch_H(cev_fdtd* p, const cev_state* st, const cev_tangent* tan, void* const H_out[3], int64_t x0, int64_t x1,
             int64_t probe_t, double* partials, cudaStream_t s) {
    StepArgs<T, AT> a;
    if (fill_args(p, st, a, tan)) return -1;
    if (H_out) {
        a.on = 63u;              // out-of-place: every output array is written
        for (int A = 0; A < 3; ++A) {
            a.Hout[A] = (T*)H_out[p->to_logical(A)];
            if (!a.Hout[A]) return fail("H_out entries must be non-NULL");
        }
    }
    const bool march = can_march(p, a, true);
    if (!march) {
        set_tiles_v1(a, x0, x1);
        const int aux = attach_probes(p, a, 0, probe_t, partials);
        if (a.n_tiles + aux == 0) return 0;
        k_step_H_v1<T, AT><<<a.n_tiles + aux, v1_block(a), 0, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    constexpr int V = vec_width<T>();
    const dim3 blk(32, V2_BY);
    if (tan) {      // tangent step E = mE*dD + dmE*D: the marching kernel with 32 lanes along z
        set_tiles_v2(p, a, x0, x1, 0, 0, 32);
        const int aux = attach_probes(p, a, 0, probe_t, partials);
        const int g = a.n_tiles + aux;
        if (g == 0) return 0;
        if (a.on == 63u) k_step_H_v2<T, AT, V, 32, false, 63, true><<<g, blk, 0, s>>>(a);
        else if (a.wz && a.on == (unsigned)MASK_TM) k_step_H_v2<T, AT, V, 32, false, MASK_TM, true><<<g, blk, 0, s>>>(a);
        else if (a.wz && a.on == (unsigned)MASK_TE) k_step_H_v2<T, AT, V, 32, false, MASK_TE, true><<<g, blk, 0, s>>>(a);
        else k_step_H_v2<T, AT, V, 32, false, -1, true><<<g, blk, 0, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    if (p->halo.on() && (!want_v5(p, a, x0, x1, true) || H_out || x0 != 0 || x1 != a.Nx))
        return fail("x-slab halos over peer memory need in-place whole-slab launches of the tensor-map kernels "
                    "(3-D grid, Ny a multiple of %d, Nz a multiple of the 16-byte vector and >= 32 vectors)", p->tma_rows);
    if (want_v5(p, a, x0, x1, true)) {          // tensor-map TMA kernel: box copies described by CUtensorMaps
        v5_set_tiles(a, x0, x1, p->tma_rows, p->xchunk);
        if (p->halo.on()) {      // this launch reads the right neighbour's D plane 0 and feeds its H plane -1
            cev_halo_layout lay;
            halo_layout(p, &lay);
            a.peer_out[0] = (T*)(p->halo.right + lay.H_lo[0]);
            a.peer_out[1] = (T*)(p->halo.right + lay.H_lo[1]);
            a.peer_flag = (unsigned long long*)(p->halo.right + lay.flag_H);
            a.own_flag = (const unsigned long long*)(p->halo.own + lay.flag_D);
            a.own_target = p->halo.nH * (uint64_t)(a.ntz * a.nty);     // the neighbour's D half-steps so far, all CTAs of its plane 0
            a.halo_err = (int*)(p->halo.own + lay.err);
            p->halo.nH++;
        }
        const int aux = attach_probes(p, a, 0, probe_t, partials);
        if (v5_launch_H<T, AT>(p->v5, a, p->tma_rows, p->tma_stages_H, aux, s)) return fail("%s", v5_last_error());
        return 0;
    }
    if (p->variant == 3 && a.on == 63u) {       // TMA-staged kernel: one launch, rows of 32 vectors
        set_tiles_v3(p, a, x0, x1);
        const int aux = attach_probes(p, a, 0, probe_t, partials);
        if (a.n_tiles + aux == 0) return 0;
        const size_t smem = V3Layout<T, V>::h_bytes();
        if (!p->smem_attr_H) {      // per plan: a plan has one device and one (T, AT)
            CUDA_TRY(cudaFuncSetAttribute(k_step_H_v3<T, AT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            p->smem_attr_H = true;
        }
        k_step_H_v3<T, AT, V><<<a.n_tiles + aux, dim3(32, V3_BY), smem, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    const bool masked = a.on != 63u;       // some components are identically zero: the mask-aware instantiation
    const bool split = !masked && want_split<T>(p, x0, x1);
    bool probes_done = false;
    for (int part = split ? 1 : 0; part <= (split ? 2 : 0); ++part) {
        set_tiles_v2(p, a, x0, x1, part);
        int aux = 0;
        if (!probes_done) {
            aux = attach_probes(p, a, 0, probe_t, partials);
            probes_done = true;
        } else {
            a.t_probe = -1;
        }
        const int g = a.n_tiles + aux;
        if (g == 0) continue;
        const int lz = a.wz ? 32 : p->lz;
        if (masked) {
            if (lz == 8) k_step_H_v2<T, AT, V, 8, false, -1><<<g, blk, 0, s>>>(a);
            else if (lz == 16) k_step_H_v2<T, AT, V, 16, false, -1><<<g, blk, 0, s>>>(a);
            else if (a.on == (unsigned)MASK_TM) k_step_H_v2<T, AT, V, 32, false, MASK_TM><<<g, blk, 0, s>>>(a);
            else if (a.on == (unsigned)MASK_TE) k_step_H_v2<T, AT, V, 32, false, MASK_TE><<<g, blk, 0, s>>>(a);
            else k_step_H_v2<T, AT, V, 32, false, -1><<<g, blk, 0, s>>>(a);
        } else if (part == 1) {
            if (lz == 8) k_step_H_v2<T, AT, V, 8, true><<<g, blk, 0, s>>>(a);
            else if (lz == 16) k_step_H_v2<T, AT, V, 16, true><<<g, blk, 0, s>>>(a);
            else k_step_H_v2<T, AT, V, 32, true><<<g, blk, 0, s>>>(a);
        } else {
            if (lz == 8) k_step_H_v2<T, AT, V, 8, false><<<g, blk, 0, s>>>(a);
            else if (lz == 16) k_step_H_v2<T, AT, V, 16, false><<<g, blk, 0, s>>>(a);
            else k_step_H_v2<T, AT, V, 32, false><<<g, blk, 0, s>>>(a);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T, typename AT>
int launch_inject(cev_fdtd* p, const cev_state* st, void* const D_out[3], const double* wave_row, int64_t x0, int64_t x1,
                  cudaStream_t s);

// wave_row != NULL: the plan's sparse sources of planes [x0, x1) are injected (fused into the marching
// kernel; a small follow-up kernel after the baseline one).
template <typename T, typename AT>
int launch_D(cev_fdtd* p, const cev_state* st, void* const D_out[3], void* const E_out[3], const void* const J[3],
             const double J_scale[3], const double* const J_wave[3], const double* wave_row, int64_t x0, int64_t x1,
             int64_t probe_t, double* partials, cudaStream_t s) {
    StepArgs<T, AT> a;
    if (fill_args(p, st, a)) return -1;
    for (int A = 0; A < 3; ++A) {
        const int L = p->to_logical(A);
        if (D_out) {
            a.Dout[A] = (T*)D_out[L];
            if (!a.Dout[A]) return fail("D_out entries must be non-NULL");
        }
        if (E_out) a.Eout[A] = (T*)E_out[L];
        if (J) a.J[A] = (const T*)J[L];
        if (J_scale) a.Jscale[A] = (AT)J_scale[L];
        if (J_wave) a.Jwave[A] = J_wave[L];
    }
    if (D_out) a.on = 63u;       // out-of-place: every output array is written
    const bool march = can_march(p, a, false);
    const bool inject = wave_row && p->n_src_pts > 0 && x1 > x0;
    if (!march) {
        set_tiles_v1(a, x0, x1);
        const int aux = attach_probes(p, a, 1, probe_t, partials);
        if (a.n_tiles + aux == 0) return 0;
        k_step_D_v1<T, AT><<<a.n_tiles + aux, v1_block(a), 0, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        if (inject) return launch_inject<T, AT>(p, st, D_out, wave_row, x0, x1, s);
        return 0;
    }
    constexpr int V = vec_width<T>();
    const dim3 blk(32, V2_BY);
    const bool extras = a.J[0] || a.J[1] || a.J[2] || a.Eout[0] || a.Eout[1] || a.Eout[2];
    if (extras) a.on = 63u;
    if (p->halo.on() && (extras || !want_v5(p, a, x0, x1, false) || D_out || x0 != 0 || x1 != a.Nx))
        return fail("x-slab halos over peer memory need in-place whole-slab launches of the tensor-map kernels "
                    "(3-D grid, Ny a multiple of %d, Nz a multiple of the 16-byte vector and >= 32 vectors, no dense J)", p->tma_rows);
    if (!extras && want_v5(p, a, x0, x1, false)) {     // tensor-map TMA kernel
        v5_set_tiles(a, x0, x1, p->tma_rows, p->xchunk);
        if (p->halo.on()) {      // this launch reads the left neighbour's last H plane and feeds its D plane nx
            cev_halo_layout lay;
            halo_layout(p, &lay);
            a.peer_out[0] = (T*)(p->halo.left + lay.D_hi[0]);
            a.peer_out[1] = (T*)(p->halo.left + lay.D_hi[1]);
            a.peer_flag = (unsigned long long*)(p->halo.left + lay.flag_D);
            a.own_flag = (const unsigned long long*)(p->halo.own + lay.flag_H);
            a.own_target = (p->halo.nD + 1) * (uint64_t)(a.ntz * a.nty);   // the neighbour's H half-step of THIS time step
            a.halo_err = (int*)(p->halo.own + lay.err);
            p->halo.nD++;
        }
        if (inject && attach_sources_v2(p, a, wave_row, 0, 32, p->tma_rows)) return -1;
        const int aux = attach_probes(p, a, 1, probe_t, partials);
        if (v5_launch_D<T, AT>(p->v5, a, p->tma_rows, p->tma_stages_D, aux, s)) return fail("%s", v5_last_error());
        return 0;
    }
    // auto: the TMA-staged D kernel wins in fp64 (measured, scripts/tune.py); fp32 and the H half-step stay on the
    // register-marching kernels
    if ((p->variant == 3 || ((p->variant == 0 || p->variant >= 4) && sizeof(T) == 8 && x1 - x0 >= 4 && a.Ny >= V3_BY)) && !extras && a.on == 63u) {
        set_tiles_v3(p, a, x0, x1);
        if (inject && attach_sources_v2(p, a, wave_row, 0, 32, V3_BY)) return -1;
        const int aux = attach_probes(p, a, 1, probe_t, partials);
        if (a.n_tiles + aux == 0) return 0;
        const size_t smem = V3Layout<T, V>::d_bytes();
        if (!p->smem_attr_D) {
            CUDA_TRY(cudaFuncSetAttribute(k_step_D_v3<T, AT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            p->smem_attr_D = true;
        }
        k_step_D_v3<T, AT, V><<<a.n_tiles + aux, dim3(32, V3_BY), smem, s>>>(a);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    const bool masked = a.on != 63u;       // (never with extras: those force the full mask above)
    const bool split = !extras && !masked && want_split<T>(p, x0, x1);   // the per-step forward() API keeps one launch
    bool probes_done = false;
    for (int part = split ? 1 : 0; part <= (split ? 2 : 0); ++part) {
        set_tiles_v2(p, a, x0, x1, part);
        if (inject && (a.wz ? attach_sources_v2(p, a, wave_row, part, -(V2_BY * 32), 1) : attach_sources_v2(p, a, wave_row, part)))
            return -1;
        int aux = 0;
        if (!probes_done) {
            aux = attach_probes(p, a, 1, probe_t, partials);
            probes_done = true;
        } else {
            a.t_probe = -1;
        }
        const int g = a.n_tiles + aux;
        if (g == 0) continue;
        const int lz = a.wz ? 32 : p->lz;
        if (masked) {
            if (lz == 8) k_step_D_v2<T, AT, V, 8, false, false, -1><<<g, blk, 0, s>>>(a);
            else if (lz == 16) k_step_D_v2<T, AT, V, 16, false, false, -1><<<g, blk, 0, s>>>(a);
            else if (a.on == (unsigned)MASK_TM) k_step_D_v2<T, AT, V, 32, false, false, MASK_TM><<<g, blk, 0, s>>>(a);
            else if (a.on == (unsigned)MASK_TE) k_step_D_v2<T, AT, V, 32, false, false, MASK_TE><<<g, blk, 0, s>>>(a);
            else k_step_D_v2<T, AT, V, 32, false, false, -1><<<g, blk, 0, s>>>(a);
        } else if (extras) {
            lz == 8 ? k_step_D_v2<T, AT, V, 8, true, false><<<g, blk, 0, s>>>(a)
                    : (lz == 16 ? k_step_D_v2<T, AT, V, 16, true, false><<<g, blk, 0, s>>>(a)
                                : k_step_D_v2<T, AT, V, 32, true, false><<<g, blk, 0, s>>>(a));
        } else if (part == 1) {
            lz == 8 ? k_step_D_v2<T, AT, V, 8, false, true><<<g, blk, 0, s>>>(a)
                    : (lz == 16 ? k_step_D_v2<T, AT, V, 16, false, true><<<g, blk, 0, s>>>(a)
                                : k_step_D_v2<T, AT, V, 32, false, true><<<g, blk, 0, s>>>(a));
        } else {
            lz == 8 ? k_step_D_v2<T, AT, V, 8, false, false><<<g, blk, 0, s>>>(a)
                    : (lz == 16 ? k_step_D_v2<T, AT, V, 16, false, false><<<g, blk, 0, s>>>(a)
                                : k_step_D_v2<T, AT, V, 32, false, false><<<g, blk, 0, s>>>(a));
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T, typename AT>
int launch_inject(cev_fdtd* p, const cev_state* st, void* const D_out[3], const double* wave_row, int64_t x0, int64_t x1,
                  cudaStream_t s) {
    if (p->n_src_pts == 0) return 0;
    SourceTable t;
    t.n = p->n_src_pts;
    t.comp = (const int32_t*)p->src_comp.p;
    t.src = (const int32_t*)p->src_id.p;
    t.cell = (const int64_t*)p->src_cell.p;
    t.weight = (const double*)p->src_weight.p;
    T* D[3];
    for (int A = 0; A < 3; ++A) D[A] = (T*)(D_out ? D_out[p->to_logical(A)] : st->D[p->to_logical(A)]);
    const int64_t plane = (int64_t)p->N[1] * p->N[2];
    const int bs = 128;
    k_inject<T, AT><<<(unsigned)((t.n + bs - 1) / bs), bs, 0, s>>>(t, D[0], D[1], D[2], wave_row, x0 * plane, x1 * plane);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T, typename AT>
int launch_probe_only(cev_fdtd* p, const cev_state* st, const cev_tangent* tan, int which, int64_t t, double* partials,
                      cudaStream_t s) {
    StepArgs<T, AT> a;
    if (fill_args(p, st, a, tan)) return -1;
    const int aux = attach_probes(p, a, which, t, partials);
    if (aux == 0) return 0;
    k_probe_only<T, AT><<<aux, dim3(V1_TZ, V1_TY), 0, s>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T, typename AT>
int launch_compute_E(cev_fdtd* p, const cev_state* st, const cev_tangent* tan, void* const E_out[3], cudaStream_t s) {
    const int64_t n = p->Nl[0] * p->Nl[1] * p->Nl[2];
    if (n == 0) return 0;
    for (int c = 0; c < 3; ++c) {
        if (!E_out[c]) continue;
        const int bs = 256;
        const int64_t want = (n + bs - 1) / bs;
        const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
        k_compute_E<T, AT><<<grid, bs, 0, s>>>((const T*)st->inv_eps[c], (const T*)st->D[c],
                                               tan ? (const T*)tan->d_inv_eps[c] : nullptr,
                                               tan ? (const T*)tan->D_primal[c] : nullptr, (T*)E_out[c], n);
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

// Monitors see the state after step n (H_n, D_n both final after the D launch).
template <typename T, typename AT>
int launch_monitors(cev_fdtd* p, const cev_state* st, int64_t n, cudaStream_t s) {
    if (p->n_mon_pts == 0 || !p->mon_acc) return 0;
    StepArgs<T, AT> a;
    if (fill_args(p, st, a)) return -1;
    MonitorTable m;
    m.n = p->n_mon_pts;
    m.nfreq = p->mon_nfreq;
    m.field = (const int32_t*)p->mon_field.p;
    m.cell = (const int64_t*)p->mon_cell.p;
    const int bs = 128;
    k_monitor<T, AT><<<(unsigned)((m.n + bs - 1) / bs), bs, 0, s>>>(a, m, p->mon_phasors + n * p->mon_nfreq * 2, p->mon_acc);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

constexpr int GRAPH_K = 50;              // time steps per captured block
constexpr int64_t GRAPH_MAX_CELLS = 1 << 18;   // auto: grids whose half-step kernels are shorter than a launch

// The graph of GRAPH_K steps for this state (captured on first use).  Row q of stage_p stands for step base-1+q.
template <typename T, typename AT>
int get_run_graph(cev_fdtd* p, const cev_state* st, cudaGraphExec_t* out) {
    if (!p->graphs.empty() && p->graphs[0].epoch != p->epoch) p->drop_graphs();
    for (auto& g : p->graphs)
        if (!memcmp(&g.st, st, sizeof(cev_state))) {
            *out = g.exec;
            return 0;
        }
    if (p->graphs.empty()) {
        if (p->stage_w.alloc((size_t)GRAPH_K * std::max(1, p->nsrc) * 8) ||
            p->stage_p.alloc((size_t)(GRAPH_K + 1) * std::max(1, p->n_slots) * 8))
            return -1;
    }
    if (!p->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
    const int64_t Nx = p->N[0];
    double* sp = (double*)p->stage_p.p;
    const double* sw = p->n_src_pts > 0 ? (const double*)p->stage_w.p : nullptr;
    CUDA_TRY(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeRelaxed));
    int rc = 0;
    for (int r = 0; r < GRAPH_K && !rc; ++r) {
        rc = launch_H<T, AT>(p, st, nullptr, nullptr, 0, Nx, r, sp, p->cap_stream);
        if (!rc)
            rc = launch_D<T, AT>(p, st, nullptr, nullptr, nullptr, nullptr, nullptr, sw ? sw + (int64_t)r * p->nsrc : nullptr, 0,
                                 Nx, r + 1, sp, p->cap_stream);
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(p->cap_stream, &graph);
    if (rc || e != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        return rc ? rc : fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    }
    cev_fdtd::RunGraph g;
    const cudaError_t e2 = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
    g.st = *st;
    g.epoch = p->epoch;
    if (p->graphs.size() >= 4) {
        cudaGraphExecDestroy(p->graphs[0].exec);
        p->graphs.erase(p->graphs.begin());
    }
    p->graphs.push_back(g);
    *out = g.exec;
    return 0;
}

// D of the recorder's box after a step -> one slot [3][bx][by][bz] of the record (storage type)
template <typename T>
__global__ void k_record_box(const T* D0, const T* D1, const T* D2, T* out, int Ny, int Nz, int x0, int y0, int z0, int bx,
                             int by, int bz) {
    const int64_t n = (int64_t)bx * by * bz;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int k = (int)(q % bz), j = (int)((q / bz) % by), i = (int)(q / ((int64_t)bz * by));
    const int64_t o = ((int64_t)(x0 + i) * Ny + (y0 + j)) * Nz + (z0 + k);
    out[q] = D0[o];
    out[n + q] = D1[o];
    out[2 * n + q] = D2[o];
}

template <typename T, typename AT>
int record_box(cev_fdtd* p, const cev_state* st, cudaStream_t s) {
    auto& r = p->rec;
    if (r.count >= r.capacity) return fail("D-box recorder is full (%lld slots)", (long long)r.capacity);
    const int bx = r.box[1] - r.box[0], by = r.box[3] - r.box[2], bz = r.box[5] - r.box[4];
    const int64_t n = (int64_t)bx * by * bz;
    const T* D[3];
    for (int A = 0; A < 3; ++A) D[A] = (const T*)st->D[p->to_logical(A)];
    k_record_box<T><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(D[0], D[1], D[2], (T*)r.buf + r.count * 3 * n, p->N[1], p->N[2],
                                                                 r.box[0], r.box[2], r.box[4], bx, by, bz);
    CUDA_TRY(cudaGetLastError());
    r.count++;
    return 0;
}

template <typename T, typename AT>
int run_loop(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* waveform, double* partials,
             cudaStream_t s) {
    const int64_t Nx = p->N[0];
    int64_t n0 = 0;
    const int64_t cells = (int64_t)p->N[0] * p->N[1] * p->N[2];
    const bool monitors = p->n_mon_pts > 0 && p->mon_acc;     // (their phasor row changes every step: no graph replay)
    const bool recording = p->rec.buf != nullptr;
    const bool graphs = !monitors && !recording && !p->halo.on() &&      // (the halo targets change with every step: no static graph)
                        (p->use_graph == 1 || (p->use_graph < 0 && cells <= GRAPH_MAX_CELLS));
    if (graphs && nsteps >= 2 * GRAPH_K + 1) {
        // step 0 the ordinary way (it also builds the source tilings and sets kernel attributes), then whole blocks
        if (launch_H<T, AT>(p, st, nullptr, nullptr, 0, Nx, -1, partials, s)) return -1;
        if (launch_D<T, AT>(p, st, nullptr, nullptr, nullptr, nullptr, nullptr, waveform, 0, Nx, 0, partials, s)) return -1;
        cudaGraphExec_t exec = nullptr;
        if (get_run_graph<T, AT>(p, st, &exec)) return -1;
        const int ns = p->n_slots, nED = p->n_slots_ED;
        for (n0 = 1; n0 + GRAPH_K <= nsteps; n0 += GRAPH_K) {
            if (waveform)
                CUDA_TRY(cudaMemcpyAsync(p->stage_w.p, waveform + n0 * p->nsrc, (size_t)GRAPH_K * p->nsrc * 8,
                                         cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaGraphLaunch(exec, s));
            const double* sp = (const double*)p->stage_p.p;
            if (nED > 0)       // E/D-family slots: staging rows 0..K-1 are steps n0-1 .. n0+K-2
                CUDA_TRY(cudaMemcpy2DAsync(partials + (n0 - 1) * ns, (size_t)ns * 8, sp, (size_t)ns * 8, (size_t)nED * 8,
                                           GRAPH_K, cudaMemcpyDeviceToDevice, s));
            if (ns > nED)      // H-family slots: staging rows 1..K are steps n0 .. n0+K-1
                CUDA_TRY(cudaMemcpy2DAsync(partials + n0 * ns + nED, (size_t)ns * 8, sp + ns + nED, (size_t)ns * 8,
                                           (size_t)(ns - nED) * 8, GRAPH_K, cudaMemcpyDeviceToDevice, s));
        }
    }
    for (int64_t n = n0; n < nsteps; ++n) {
        // E/D probes of step n-1 ride on the H launch of step n (D is read-only there);
        // H probes of step n ride on its D launch (H is read-only there).
        if (launch_H<T, AT>(p, st, nullptr, nullptr, 0, Nx, n - 1, partials, s)) return -1;
        if (launch_D<T, AT>(p, st, nullptr, nullptr, nullptr, nullptr, nullptr, waveform ? waveform + n * p->nsrc : nullptr, 0,
                            Nx, n, partials, s))
            return -1;
        if (monitors && launch_monitors<T, AT>(p, st, n, s)) return -1;
        if (recording && record_box<T, AT>(p, st, s)) return -1;
    }
    if (nsteps > 0 && launch_probe_only<T, AT>(p, st, nullptr, 0, nsteps - 1, partials, s)) return -1;
    return 0;
}

// ---- batched tangent half-steps: the B launches of one half-step as ONE (step_v2.cuh, k_step_*_v2_batch) -----------
// Per-tangent StepArgs exactly as launch_H / launch_D would build them, uploaded once per sweep (the probe row is a
// kernel parameter).  which = 0: tangent H half-steps (E = mE dD + dmE D), 1: tangent D half-steps on masked
// single-row-plane grids (2-D TM / TE).  Returns 1 if the batch was built, 0 if this plan / state is not served
// (callers fall back to one launch per tangent), -1 on error.
template <typename T, typename AT>
int build_tangent_batch(cev_fdtd* p, int which, int B, const cev_state* tst, const cev_tangent* tan, double* tpartials,
                        int64_t stride, int64_t Nx, int* n_tiles, int* aux, unsigned* mask, cudaStream_t s) {
    if (p->halo.on() || p->variant == 1) return 0;
    std::vector<StepArgs<T, AT>> tab((size_t)B);
    const int saved_chunk = p->xchunk;
    for (int b = 0; b < B; ++b) {
        StepArgs<T, AT>& a = tab[b];
        if (fill_args(p, &tst[b], a, which == 0 ? &tan[b] : nullptr)) return -1;
        if (!can_march(p, a, which == 0)) return 0;
        if (which == 1 && !(a.Ny == 1 && a.on != 63u)) return 0;       // D: only the masked 2-D path is batched
        if (a.Ny == 1 && p->xchunk <= 0) p->xchunk = 32;                // B x as many CTAs: long chunks still fill the GPU
        if (which == 0) set_tiles_v2(p, a, 0, Nx, 0, 0, 32);
        else set_tiles_v2(p, a, 0, Nx, 0);
        p->xchunk = saved_chunk;
        if (which == 1 && !a.wz) return 0;
        a.aux_slot0 = which == 0 ? 0 : p->n_slots_ED;
        a.partials = tpartials ? tpartials + b * stride : nullptr;
        a.t_probe = -1;
        if (b > 0 && (a.n_tiles != tab[0].n_tiles || a.on != tab[0].on)) return 0;
    }
    *n_tiles = tab[0].n_tiles;
    *aux = (tpartials && p->n_slots > 0) ? (which == 0 ? p->n_slots_ED : p->n_slots - p->n_slots_ED) : 0;
    *mask = tab[0].on;
    DeviceBuf& buf = which == 0 ? p->batch_tab_H : p->batch_tab_D;
    const size_t bytes = (size_t)B * sizeof(StepArgs<T, AT>);
    if (buf.bytes < bytes) {
        buf.release();
        if (buf.alloc(bytes)) return -1;
    }
    CUDA_TRY(cudaMemcpyAsync(buf.p, tab.data(), bytes, cudaMemcpyHostToDevice, s));
    return 1;
}

template <typename T, typename AT>
int launch_tangent_batch(cev_fdtd* p, int which, int B, int n_tiles, int aux, unsigned mask, bool wz, int64_t probe_t, cudaStream_t s) {
    constexpr int V = vec_width<T>();
    const dim3 blk(32, V2_BY);
    const int per = n_tiles + (probe_t >= 0 ? aux : 0);
    if (per == 0) return 0;
    const unsigned g = (unsigned)per * (unsigned)B;
    if (which == 0) {
        const StepArgs<T, AT>* tab = (const StepArgs<T, AT>*)p->batch_tab_H.p;
        if (mask == 63u) k_step_H_v2_batch<T, AT, V, 32, false, 63, true><<<g, blk, 0, s>>>(tab, B, probe_t);
        else if (wz && mask == (unsigned)MASK_TM) k_step_H_v2_batch<T, AT, V, 32, false, MASK_TM, true><<<g, blk, 0, s>>>(tab, B, probe_t);
        else if (wz && mask == (unsigned)MASK_TE) k_step_H_v2_batch<T, AT, V, 32, false, MASK_TE, true><<<g, blk, 0, s>>>(tab, B, probe_t);
        else k_step_H_v2_batch<T, AT, V, 32, false, -1, true><<<g, blk, 0, s>>>(tab, B, probe_t);
    } else {
        const StepArgs<T, AT>* tab = (const StepArgs<T, AT>*)p->batch_tab_D.p;
        if (mask == (unsigned)MASK_TM) k_step_D_v2_batch<T, AT, V, 32, false, false, MASK_TM><<<g, blk, 0, s>>>(tab, B, probe_t);
        else if (mask == (unsigned)MASK_TE) k_step_D_v2_batch<T, AT, V, 32, false, false, MASK_TE><<<g, blk, 0, s>>>(tab, B, probe_t);
        else k_step_D_v2_batch<T, AT, V, 32, false, false, -1><<<g, blk, 0, s>>>(tab, B, probe_t);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- forward-mode sweep on 2-D TM grids with the fused tangent step (tan2d_fused.cuh) -------------------------------
// Per time step: ONE launch advances all B tangent states by a whole step (it reads the primal D of the previous step, so
// it goes first), then the primal H and D half-steps.  The tangent states ping-pong between the caller's arrays (even
// steps read them) and plan-owned shadows; after an odd number of steps the shadows are copied back.
// Returns 1 if it ran, 0 if this plan / state is not served, -1 on error.
template <typename T, typename AT>
int jvp_loop_fused2d(cev_fdtd* p, const cev_state* st, int B, const cev_state* tst, const cev_tangent* tan, int64_t nsteps,
                     const double* waveform, double* partials, double* tpartials, cudaStream_t s) {
    constexpr int V = vec_width<T>();
    const int64_t Nx = p->N[0];
    if (!p->jvp_fused || p->halo.on() || p->N[1] != 1 || p->on != (unsigned)MASK_TM || nsteps < 1) return 0;
    const int64_t stride = nsteps * p->n_slots;
    const int Nz = p->N[2];
    const size_t es = sizeof(T);
    const size_t nfield = (size_t)Nx * Nz * es, nI0 = (size_t)std::max(p->nH[0], 0) * Nz * es, nI2 = (size_t)Nx * std::max(p->nH[2], 0) * es;
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t per = 3 * up(nfield) + up(nI0) + up(nI2);
    std::vector<StepArgs<T, AT>> tabA((size_t)B), tabB((size_t)B);
    if (p->tan_shadow.bytes < per * (size_t)B) {
        p->tan_shadow.release();
        if (p->tan_shadow.alloc(per * (size_t)B)) return -1;
    }
    for (int b = 0; b < B; ++b) {
        StepArgs<T, AT>& a = tabA[b];
        if (fill_args(p, &tst[b], a, &tan[b])) return -1;
        if (!can_march(p, a, true)) return 0;
        unsigned char* base = (unsigned char*)p->tan_shadow.p + per * (size_t)b;
        T* sD = (T*)base;
        T* sH0 = (T*)(base + up(nfield));
        T* sH2 = (T*)(base + 2 * up(nfield));
        T* sI0 = (T*)(base + 3 * up(nfield));
        T* sI2 = (T*)(base + 3 * up(nfield) + up(nI0));
        a.ntz = (Nz + V2_BY * 32 * V - 1) / (V2_BY * 32 * V);
        // measured on B200 (profiles/r2_tune_fused_tangent_step.log): fp32 (four cells per thread, half as many CTAs) wants
        // shorter chunks than fp64
        a.xchunk = p->xchunk > 0 ? p->xchunk : (sizeof(T) == 4 ? 16 : 32);
        a.n_tiles = a.ntz * (int)((Nx + a.xchunk - 1) / a.xchunk);
        a.pf_dist = p->pf_dist;
        a.aux_slot0 = 0;
        a.partials = tpartials ? tpartials + b * stride : nullptr;
        a.t_probe = -1;
        StepArgs<T, AT>& bb = tabB[b];
        bb = a;
        // A: caller -> shadow
        a.Hout[0] = sH0; a.Hout[2] = sH2; a.Dout[1] = sD; a.ICEout[0] = sI0; a.ICEout[2] = sI2;
        // B: shadow -> caller
        bb.Hin[0] = sH0; bb.Hin[2] = sH2; bb.Din[1] = sD; bb.ICE[0] = sI0; bb.ICE[2] = sI2;
        bb.ICEout[0] = a.ICE[0]; bb.ICEout[2] = a.ICE[2];
        if (b > 0 && a.n_tiles != tabA[0].n_tiles) return 0;
    }
    const size_t tbytes = (size_t)B * sizeof(StepArgs<T, AT>);
    for (DeviceBuf* buf : {&p->batch_tab_F0, &p->batch_tab_F1})
        if (buf->bytes < tbytes) {
            buf->release();
            if (buf->alloc(tbytes)) return -1;
        }
    CUDA_TRY(cudaMemcpyAsync(p->batch_tab_F0.p, tabA.data(), tbytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(p->batch_tab_F1.p, tabB.data(), tbytes, cudaMemcpyHostToDevice, s));
    const int n_tiles = tabA[0].n_tiles;
    const int aux = (tpartials && p->n_slots > 0) ? p->n_slots : 0;
    const dim3 blk(32, V2_BY);
    for (int64_t n = 0; n < nsteps; ++n) {
        const StepArgs<T, AT>* tab = (const StepArgs<T, AT>*)((n & 1) ? p->batch_tab_F1.p : p->batch_tab_F0.p);
        const unsigned g = (unsigned)(n_tiles + (n > 0 ? aux : 0)) * (unsigned)B;
        k_tan2d_fused_batch<T, AT, V><<<g, blk, 0, s>>>(tab, B, n - 1);
        CUDA_TRY(cudaGetLastError());
        if (launch_H<T, AT>(p, st, nullptr, nullptr, 0, Nx, n - 1, partials, s)) return -1;
        if (launch_D<T, AT>(p, st, nullptr, nullptr, nullptr, nullptr, nullptr, waveform ? waveform + n * p->nsrc : nullptr, 0, Nx,
                            n, partials, s))
            return -1;
    }
    if (nsteps & 1)      // the final tangent states live in the shadows
        for (int b = 0; b < B; ++b) {
            const StepArgs<T, AT>& a = tabA[b];
            CUDA_TRY(cudaMemcpyAsync((void*)a.Hin[0], a.Hout[0], nfield, cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync((void*)a.Hin[2], a.Hout[2], nfield, cudaMemcpyDeviceToDevice, s));
            CUDA_TRY(cudaMemcpyAsync((void*)a.Din[1], a.Dout[1], nfield, cudaMemcpyDeviceToDevice, s));
            if (nI0) CUDA_TRY(cudaMemcpyAsync(a.ICE[0], a.ICEout[0], nI0, cudaMemcpyDeviceToDevice, s));
            if (nI2) CUDA_TRY(cudaMemcpyAsync(a.ICE[2], a.ICEout[2], nI2, cudaMemcpyDeviceToDevice, s));
        }
    if (launch_probe_only<T, AT>(p, st, nullptr, 0, nsteps - 1, partials, s)) return -1;
    for (int b = 0; b < B; ++b)
        for (int which = 0; which < 2; ++which)
            if (launch_probe_only<T, AT>(p, &tst[b], &tan[b], which, nsteps - 1, tpartials ? tpartials + b * stride : nullptr, s))
                return -1;
    return 1;
}

// Primal + B tangents in one sweep.  Per step: tangent H half-steps first (they need the primal D of
// the previous step), then the primal step, then the tangent D half-steps (same linear update, J = 0).
template <typename T, typename AT>
int jvp_loop(cev_fdtd* p, const cev_state* st, int B, const cev_state* tst, const cev_tangent* tan, int64_t nsteps,
             const double* waveform, double* partials, double* tpartials, cudaStream_t s) {
    const int64_t Nx = p->N[0];
    const int64_t stride = nsteps * p->n_slots;
    {
        const int rc = jvp_loop_fused2d<T, AT>(p, st, B, tst, tan, nsteps, waveform, partials, tpartials, s);
        if (rc != 0) return rc < 0 ? -1 : 0;
    }
    // The B tangent states are independent of each other; on grids whose single launches cannot fill the GPU they
    // run on B side streams (fork / join with events) so that their kernels overlap.  Dependencies: a tangent H
    // half-step reads the primal D of the previous step (so it waits for the primal D launch, and the next primal D
    // launch waits for it); a tangent D half-step only touches its own state.
    const int64_t cells = (int64_t)p->N[0] * p->N[1] * p->N[2];
    // auto: where it was measured to pay (scripts/tune2d.py) -- 2-D grids up to 2^23 cells, and small grids of any shape
    const bool small = cells <= ((int64_t)1 << 20) || (p->N[1] == 1 && cells <= ((int64_t)1 << 23));
    // One launch per half-step for all B tangents where the marching kernels serve them (the H half-step everywhere, the D
    // half-step on 2-D polarised grids); what is not batched runs as one launch per tangent, on side streams if `fork`.
    int bH = 0, bD = 0, ntH = 0, auxH = 0, ntD = 0, auxD = 0;
    unsigned maskH = 63u, maskD = 63u;
    if (B >= 2 && p->jvp_batch != 0 && nsteps > 0) {
        bH = build_tangent_batch<T, AT>(p, 0, B, tst, tan, tpartials, stride, Nx, &ntH, &auxH, &maskH, s);
        if (bH < 0) return -1;
        bD = build_tangent_batch<T, AT>(p, 1, B, tst, tan, tpartials, stride, Nx, &ntD, &auxD, &maskD, s);
        if (bD < 0) return -1;
    }
    const bool wz_grid = p->N[1] == 1;
    const bool fork = B >= 2 && !(bH && bD) && (p->jvp_streams == 1 || (p->jvp_streams < 0 && small));
    if (fork) {
        while ((int)p->side.size() < B) {
            cudaStream_t q;
            cudaEvent_t e;
            CUDA_TRY(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            p->side.push_back(q);
            p->side_ev.push_back(e);
        }
        if (!p->main_ev) CUDA_TRY(cudaEventCreateWithFlags(&p->main_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(p->main_ev, s));            // everything queued on the caller's stream so far
    }
    auto ts = [&](int b) { return fork ? p->side[b] : s; };
    for (int64_t n = 0; n < nsteps; ++n) {
        if (bH) {
            if (fork)      // (the tangent D half-steps of step n-1 ran on the side streams)
                for (int b = 0; b < B; ++b) {
                    CUDA_TRY(cudaEventRecord(p->side_ev[b], ts(b)));
                    CUDA_TRY(cudaStreamWaitEvent(s, p->side_ev[b], 0));
                }
            if (launch_tangent_batch<T, AT>(p, 0, B, ntH, auxH, maskH, wz_grid, n - 1, s)) return -1;
        } else {
            for (int b = 0; b < B; ++b) {
                if (fork) CUDA_TRY(cudaStreamWaitEvent(ts(b), p->main_ev, 0));     // primal D of step n-1 is final
                if (launch_H<T, AT>(p, &tst[b], &tan[b], nullptr, 0, Nx, n - 1, tpartials ? tpartials + b * stride : nullptr, ts(b))) return -1;
                if (fork) CUDA_TRY(cudaEventRecord(p->side_ev[b], ts(b)));
            }
        }
        if (launch_H<T, AT>(p, st, nullptr, nullptr, 0, Nx, n - 1, partials, s)) return -1;
        if (fork && !bH)
            for (int b = 0; b < B; ++b) CUDA_TRY(cudaStreamWaitEvent(s, p->side_ev[b], 0));   // they have read the primal D
        if (launch_D<T, AT>(p, st, nullptr, nullptr, nullptr, nullptr, nullptr, waveform ? waveform + n * p->nsrc : nullptr, 0,
                            Nx, n, partials, s))
            return -1;
        if (fork) CUDA_TRY(cudaEventRecord(p->main_ev, s));
        if (bD) {
            if (launch_tangent_batch<T, AT>(p, 1, B, ntD, auxD, maskD, wz_grid, n, s)) return -1;
        } else {
            for (int b = 0; b < B; ++b) {
                if (fork && bH) CUDA_TRY(cudaStreamWaitEvent(ts(b), p->main_ev, 0));   // the batched H half-step ran on the caller's stream
                if (launch_D<T, AT>(p, &tst[b], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, Nx, n,
                                    tpartials ? tpartials + b * stride : nullptr, ts(b)))
                    return -1;
            }
        }
    }
    if (nsteps > 0) {
        if (launch_probe_only<T, AT>(p, st, nullptr, 0, nsteps - 1, partials, s)) return -1;
        for (int b = 0; b < B; ++b) {
            if (fork) CUDA_TRY(cudaStreamWaitEvent(ts(b), p->main_ev, 0));
            if (launch_probe_only<T, AT>(p, &tst[b], &tan[b], 0, nsteps - 1, tpartials ? tpartials + b * stride : nullptr, ts(b))) return -1;
        }
    }
    if (fork)          // join: the caller's stream continues after every side stream
        for (int b = 0; b < B; ++b) {
            CUDA_TRY(cudaEventRecord(p->side_ev[b], ts(b)));
            CUDA_TRY(cudaStreamWaitEvent(s, p->side_ev[b], 0));
        }
    return 0;
}


// ---- fused full-step kernel (step_v4.cuh): one launch per time step, state ping-ponged between `in` and `out`
template <typename T, typename AT>
bool can_fuse(const cev_fdtd* p, const StepArgs<T, AT>& a) {
    constexpr int V = vec_width<T>();
    if (p->on != 63u || a.dmE[0]) return false;
    if (a.Nz % V != 0 || a.Nz / V < 2) return false;
    auto ok = [](const void* q) { return q == nullptr || ((uintptr_t)q % 16) == 0; };
    for (int c = 0; c < 3; ++c)
        if (!ok(a.Hin[c]) || !ok(a.Hout[c]) || !ok(a.Din[c]) || !ok(a.Dout[c]) || !ok(a.mE[c]) || !ok(a.ICE[c]) ||
            !ok(a.ICEout[c]) || !ok(a.ICH[c]))
            return false;
    return true;
}

template <typename T, typename AT, int LZ, int BY>
int launch_fused_shape(cev_fdtd* p, StepArgs<T, AT>& a, const double* wave_row, int n_aux_slots, int64_t probe_t,
                       double* partials, cudaStream_t s, const int* box = nullptr /* PML-free {x0,x1,y0,y1,z0,z1} */) {
    constexpr int V = vec_width<T>();
    constexpr int OY = BY * (32 / LZ) - 1, OZ = LZ - 1;
    a.x0 = 0;
    a.x1 = a.Nx;
    int chunk = p->xchunk > 0 ? p->xchunk : 16;   // the pre-roll plane costs 1/chunk extra H work
    a.xchunk = chunk;
    a.pf_dist = p->pf_dist;
    a.n_boxes = 1;
    Box& B = a.box[0];
    B.x0 = 0; B.x1 = a.Nx; B.y0 = 0; B.y1 = a.Ny; B.z0 = 0; B.z1 = a.Nz;
    if (box) { B.x0 = box[0]; B.x1 = box[1]; B.y0 = box[2]; B.y1 = box[3]; B.z0 = box[4]; B.z1 = box[5]; }
    a.ntz = ((B.z1 - B.z0) / V + OZ - 1) / OZ;
    a.nty = (B.y1 - B.y0 + OY - 1) / OY;
    a.n_tiles = a.ntz * a.nty * ((B.x1 - B.x0 + chunk - 1) / chunk);
    B.cta0 = 0; B.ntz = a.ntz; B.nty = a.nty;
    if (wave_row && p->n_src_pts > 0 && attach_sources_v2(p, a, wave_row, box ? 6 : 4, -OZ, OY)) return -1;
    int aux = 0;
    if (probe_t >= 0 && partials && n_aux_slots > 0) {
        a.aux_slot0 = 0;
        a.t_probe = probe_t;
        a.partials = partials;
        aux = n_aux_slots;
    }
    const bool nopml = box || p->nH[0] + p->nH[1] + p->nH[2] + p->nD[0] + p->nD[1] + p->nD[2] == 0;
    if (a.n_tiles + aux == 0) return 0;
    if (nopml && BY == 4 && LZ == 16) k_step_fused<T, AT, V, 16, 4, true><<<a.n_tiles + aux, dim3(32, 4), 0, s>>>(a);
    else k_step_fused<T, AT, V, LZ, BY><<<a.n_tiles + aux, dim3(32, BY), 0, s>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// One fused step in -> out.  n_aux_slots: how many probe slots (from slot 0: the E/D family comes first) are sampled
// on the INPUT state, into row probe_t of partials.
template <typename T, typename AT>
int launch_fused(cev_fdtd* p, const cev_state* in, const cev_state* out, const double* wave_row, int n_aux_slots,
                 int64_t probe_t, double* partials, cudaStream_t s, bool* done) {
    StepArgs<T, AT> a;
    *done = false;
    if (fill_args(p, in, a)) return -1;
    for (int A = 0; A < 3; ++A) {
        const int L = p->to_logical(A);
        a.Hout[A] = (T*)out->H[L];
        a.Dout[A] = (T*)out->D[L];
        a.ICEout[A] = (T*)out->ICE[L];
        a.IHout[A] = (T*)out->IH[L];
        if (!a.Hout[A] || !a.Dout[A]) return fail("fused step: the shadow state needs H and D");
        if ((a.ICE[A] && !a.ICEout[A]) || (a.IH[A] && !a.IHout[A])) return fail("fused step: the shadow state needs ICE / IH wherever the state has them");
    }
    if (!can_fuse(p, a)) return 0;
    int shape = p->fused_shape;
    if (shape == 0) shape = 1604;
    switch (shape) {
        case 1604: if (launch_fused_shape<T, AT, 16, 4>(p, a, wave_row, n_aux_slots, probe_t, partials, s)) return -1; break;
        case 1608: if (launch_fused_shape<T, AT, 16, 8>(p, a, wave_row, n_aux_slots, probe_t, partials, s)) return -1; break;
        case 3204: if (launch_fused_shape<T, AT, 32, 4>(p, a, wave_row, n_aux_slots, probe_t, partials, s)) return -1; break;
        case 3208: if (launch_fused_shape<T, AT, 32, 8>(p, a, wave_row, n_aux_slots, probe_t, partials, s)) return -1; break;
        case 804:  if (launch_fused_shape<T, AT, 8, 4>(p, a, wave_row, n_aux_slots, probe_t, partials, s)) return -1; break;
        default: return fail("fused_shape must be one of 804, 1604, 1608, 3204, 3208");
    }
    *done = true;
    return 0;
}


// ---- hybrid step: the lean (PML-free) fused kernel on the interior box, the two general half-step kernels out of
// place on the six slabs of the PML shell.  The fused box is the PML-free interior shrunk by one cell on the low side of
// every PML axis, so that the halo cells it recomputes are PML-free too.  Ping-pong of H and D only: the PML integrals
// are touched by the shell kernels alone, in place.
template <typename T>
bool hybrid_box(const cev_fdtd* p, int box[6]) {
    constexpr int V = vec_width<T>();
    bool any_pml = false;
    for (int A = 0; A < 3; ++A) {
        const bool pml = p->nH[A] + p->nD[A] > 0;
        any_pml |= pml;
        box[2 * A] = pml ? p->in_lo[A] + 1 : 0;
        box[2 * A + 1] = pml ? p->in_hi[A] : p->N[A];
    }
    box[4] = (box[4] + V - 1) / V * V;
    box[5] = box[5] / V * V;
    if (!any_pml) return false;                       // (no PML: the fused kernel serves the whole grid)
    for (int A = 0; A < 3; ++A)
        if (box[2 * A + 1] - box[2 * A] < 16) return false;
    return true;
}

template <typename T, typename AT>
int launch_hybrid_step(cev_fdtd* p, const cev_state* in, const cev_state* out, const int box[6], const double* wave_row,
                       int n_aux_slots, int64_t probe_t, double* partials, cudaStream_t s) {
    constexpr int V = vec_width<T>();
    const dim3 blk(32, V2_BY);
    // 1. interior: fused H + D, in -> out
    {
        StepArgs<T, AT> a;
        if (fill_args(p, in, a)) return -1;
        for (int A = 0; A < 3; ++A) {
            const int L = p->to_logical(A);
            a.Hout[A] = (T*)out->H[L];
            a.Dout[A] = (T*)out->D[L];
        }
        if (launch_fused_shape<T, AT, 16, 4>(p, a, wave_row, n_aux_slots, probe_t, partials, s, box)) return -1;
    }
    // 2. shell, H half-step: reads in.D / in.H, writes out.H
    {
        StepArgs<T, AT> a;
        if (fill_args(p, in, a)) return -1;
        for (int A = 0; A < 3; ++A) a.Hout[A] = (T*)out->H[p->to_logical(A)];
        set_tiles_v2(p, a, 0, a.Nx, 2, 0, 0, box);
        a.t_probe = -1;
        if (a.n_tiles > 0) {
            if (p->lz == 8) k_step_H_v2<T, AT, V, 8, false><<<a.n_tiles, blk, 0, s>>>(a);
            else if (p->lz == 16) k_step_H_v2<T, AT, V, 16, false><<<a.n_tiles, blk, 0, s>>>(a);
            else k_step_H_v2<T, AT, V, 32, false><<<a.n_tiles, blk, 0, s>>>(a);
        }
    }
    // 3. shell, D half-step: reads out.H (shell cells and their interior neighbours) and in.D, writes out.D
    {
        cev_state mid = *in;
        for (int c = 0; c < 3; ++c) mid.H[c] = out->H[c];
        StepArgs<T, AT> a;
        if (fill_args(p, &mid, a)) return -1;
        for (int A = 0; A < 3; ++A) a.Dout[A] = (T*)out->D[p->to_logical(A)];
        set_tiles_v2(p, a, 0, a.Nx, 2, 0, 0, box);
        a.t_probe = -1;
        if (wave_row && p->n_src_pts > 0 && attach_sources_v2(p, a, wave_row, 5)) return -1;
        if (a.n_tiles > 0) {
            if (p->lz == 8) k_step_D_v2<T, AT, V, 8, false, false><<<a.n_tiles, blk, 0, s>>>(a);
            else if (p->lz == 16) k_step_D_v2<T, AT, V, 16, false, false><<<a.n_tiles, blk, 0, s>>>(a);
            else k_step_D_v2<T, AT, V, 32, false, false><<<a.n_tiles, blk, 0, s>>>(a);
        }
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// nsteps time steps; the result ends up in `st` (an odd step count starts with one two-kernel step).
template <typename T, typename AT>
int run_loop_fused(cev_fdtd* p, const cev_state* st, const cev_state* shadow, int64_t nsteps, const double* waveform,
                   double* partials, cudaStream_t s) {
    const int64_t Nx = p->N[0];
    cev_state B = *st;            // same 1/eps and D-side integrals, shadow H / D / ICE / IH
    for (int c = 0; c < 3; ++c) {
        B.H[c] = shadow->H[c];
        B.D[c] = shadow->D[c];
        B.ICE[c] = shadow->ICE[c];
        B.IH[c] = shadow->IH[c];
    }
    const cev_state* cur = st;
    const cev_state* nxt = &B;
    int64_t n = 0;
    bool h_probes_pending = false;       // H-family probes of step n-1 not sampled yet
    auto two_kernel_step = [&](int64_t t) -> int {
        if (h_probes_pending && launch_probe_only<T, AT>(p, cur, nullptr, 1, t - 1, partials, s)) return -1;
        if (launch_H<T, AT>(p, cur, nullptr, nullptr, 0, Nx, t - 1, partials, s)) return -1;
        if (launch_D<T, AT>(p, cur, nullptr, nullptr, nullptr, nullptr, nullptr, waveform ? waveform + t * p->nsrc : nullptr, 0,
                            Nx, t, partials, s))
            return -1;
        h_probes_pending = false;
        return 0;
    };
    if (nsteps & 1) {
        if (two_kernel_step(0)) return -1;
        n = 1;
    }
    int hbox[6];
    bool hybrid = p->variant == 5 && hybrid_box<T>(p, hbox);
    if (hybrid) {          // same applicability rules as the fused kernel
        StepArgs<T, AT> probe_args;
        if (fill_args(p, st, probe_args)) return -1;
        for (int A = 0; A < 3; ++A) {
            probe_args.Hout[A] = (T*)B.H[p->to_logical(A)];
            probe_args.Dout[A] = (T*)B.D[p->to_logical(A)];
            probe_args.ICEout[A] = probe_args.ICE[A];
        }
        hybrid = can_fuse(p, probe_args) && can_march(p, probe_args, true);
    }
    if (hybrid)            // only H and D are ping-ponged: the shell kernels update the PML integrals in place
        for (int c = 0; c < 3; ++c) {
            B.ICE[c] = st->ICE[c];
            B.IH[c] = st->IH[c];
        }
    for (; n < nsteps; ++n) {
        bool done = false;
        const int slots = h_probes_pending ? p->n_slots : p->n_slots_ED;   // E/D probes of step n-1 always ride here
        if (hybrid) {
            if (launch_hybrid_step<T, AT>(p, cur, nxt, hbox, waveform ? waveform + n * p->nsrc : nullptr, slots, n - 1, partials, s)) return -1;
            h_probes_pending = true;
            std::swap(cur, nxt);
            continue;
        }
        if (launch_fused<T, AT>(p, cur, nxt, waveform ? waveform + n * p->nsrc : nullptr, slots, n - 1, partials, s, &done)) return -1;
        if (!done) {              // not fusable (geometry / alignment): finish with the two-kernel path, in `st`
            if (cur != st) return fail("internal: fused fallback on the shadow state");
            for (; n < nsteps; ++n)
                if (two_kernel_step(n)) return -1;
            break;
        }
        h_probes_pending = true;
        std::swap(cur, nxt);
    }
    if (cur != st) return fail("internal: fused run ended on the shadow state");
    if (nsteps > 0) {
        if (launch_probe_only<T, AT>(p, st, nullptr, 0, nsteps - 1, partials, s)) return -1;
        if (h_probes_pending && launch_probe_only<T, AT>(p, st, nullptr, 1, nsteps - 1, partials, s)) return -1;
    }
    return 0;
}

template <typename T, typename AT>
int fill_adj(const cev_fdtd* p, const cev_state* fwd, const cev_adjoint* adj, AdjArgs<T, AT>& a) {
    memset(&a, 0, sizeof a);
    a.Nx = p->N[0];
    a.Ny = p->N[1];
    a.Nz = p->N[2];
    constexpr int w = sizeof(AT) == 8 ? 1 : 0;
    for (int A = 0; A < 3; ++A) {
        const int L = p->to_logical(A);
        a.lH[A] = (T*)adj->lH[L];
        a.lD[A] = (T*)adj->lD[L];
        a.lICE[A] = (T*)adj->lICE[L];
        a.lIH[A] = (T*)adj->lIH[L];
        a.lICH[A] = (T*)adj->lICH[L];
        a.lID[A] = (T*)adj->lID[L];
        a.gC[A] = (T*)adj->gC[L];
        a.gC2[A] = (T*)adj->gC2[L];
        a.G[A] = adj->G_mE[L];
        a.mE[A] = (const T*)fwd->inv_eps[L];
        a.Dprev[A] = (const T*)fwd->D[L];
        if (!a.lH[A] || !a.lD[A] || !a.gC2[A] || !a.mE[A] || !a.Dprev[A])
            return fail("cev_adjoint: lH, lD, gC2 and the forward inv_eps / D must be non-NULL");
        const int Bx = (A + 1) % 3, C = (A + 2) % 3;
        if (p->nH[A] > 0 && !a.lICE[A]) return fail("cev_adjoint: lICE missing for a PML axis");
        if (p->nD[A] > 0 && !a.lICH[A]) return fail("cev_adjoint: lICH missing for a PML axis");
        if (p->nH[Bx] > 0 && p->nH[C] > 0 && !a.lIH[A]) return fail("cev_adjoint: lIH missing for a PML corner");
        if (p->nD[Bx] > 0 && p->nD[C] > 0 && !a.lID[A]) return fail("cev_adjoint: lID missing for a PML corner");
        a.mapH[A] = p->mapH[A];
        a.mapD[A] = p->mapD[A];
        a.nH[A] = p->nH[A];
        a.nD[A] = p->nD[A];
        a.uH[A] = (const AT*)p->uH[A][w];
        a.rH[A] = (const AT*)p->rH[A][w];
        a.uD[A] = (const AT*)p->uD[A][w];
        a.rD[A] = (const AT*)p->rD[A][w];
    }
    a.cdt = (AT)(p->parity * p->cdt);
    a.inv_dL = (AT)(1.0 / p->dL);
    // design box (logical x0, x1, y0, y1, z0, z1; all zero = the whole grid) in internal axis order
    bool whole = true;
    for (int q = 0; q < 6; ++q) whole = whole && adj->g_box[q] == 0;
    for (int A = 0; A < 3; ++A) {
        const int L = p->to_logical(A);
        const int64_t lo = whole ? 0 : adj->g_box[2 * L], hi = whole ? p->Nl[L] : adj->g_box[2 * L + 1];
        if (lo < 0 || hi > p->Nl[L] || lo > hi) return fail("cev_adjoint: g_box outside the grid");
        a.gb[2 * A] = (int)lo;
        a.gb[2 * A + 1] = (int)hi;
    }
    return 0;
}

template <typename T, typename AT>
int launch_adjoint_step(cev_fdtd* p, const cev_state* fwd, const cev_adjoint* adj, cudaStream_t s) {
    AdjArgs<T, AT> a;
    if (fill_adj(p, fwd, adj, a)) return -1;
    const dim3 blk(64, 4);
    const dim3 grd((a.Nz + 63) / 64, (a.Ny + 3) / 4, a.Nx);
    if (grd.y > 65535 || grd.z > 65535) return fail("adjoint kernels: grid extent too large");
    k_adj_H<T, AT><<<grd, blk, 0, s>>>(a);
    k_adj_ED<T, AT><<<grd, blk, 0, s>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

template <typename T, typename AT>
int launch_adjoint_seed(cev_fdtd* p, const cev_state* fwd, const cev_adjoint* adj, const double* gbar_row, cudaStream_t s) {
    if (p->n_slots == 0) return 0;
    AdjArgs<T, AT> a;
    if (fill_adj(p, fwd, adj, a)) return -1;
    ProbeTable pr;
    fill_probe_table(p, pr);
    k_adj_seed<T, AT><<<p->n_slots, 128, 0, s>>>(a, pr, (const int32_t*)p->pr_owner.p, gbar_row, a.Dprev[0], a.Dprev[1], a.Dprev[2]);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// The tensor-map kernels of the reverse sweep (adjoint_v5.cuh) serve whole 3-D grids with the forward kernels' tile
// constraints; they carry the cotangents in the "eager" form and need the third scratch vector field adj->gC.
template <typename T, typename AT>
bool adjoint_v5_ok(cev_fdtd* p, const AdjArgs<T, AT>& a) {
    constexpr int V = vec_width<T>();
    if (p->adjoint_variant == 1) return false;
    if (p->perm[0] != 0 || p->perm[1] != 1 || p->perm[2] != 2 || p->on != 63u) return false;
    if (a.Ny < p->tma_rows || a.Ny % p->tma_rows != 0 || a.Nz % V != 0 || a.Nz < 32 * V) return false;
    if (p->adjoint_variant != 2 && (int64_t)a.Ny * a.Nz < (1 << 14)) return false;     // small planes: too few CTAs per chunk
    auto al = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
    for (int c = 0; c < 3; ++c) {
        if (!a.gC[c]) return false;
        if (!al(a.lH[c]) || !al(a.lD[c]) || !al(a.gC[c]) || !al(a.gC2[c]) || !al(a.mE[c]) || !al(a.lICE[c]) || !al(a.lIH[c]) ||
            !al(a.lICH[c]) || !al(a.lID[c]))
            return false;
    }
    return true;
}

// StepArgs of the two tensor-map adjoint kernels (field re-use documented at AdjV5Extra)
template <typename T, typename AT>
void adjoint_v5_args(cev_fdtd* p, const AdjArgs<T, AT>& a, StepArgs<T, AT>& aH, StepArgs<T, AT>& aED) {
    memset(&aH, 0, sizeof aH);
    aH.Nx = a.Nx; aH.Ny = a.Ny; aH.Nz = a.Nz;
    for (int A = 0; A < 3; ++A) {
        aH.mapH[A] = a.mapH[A]; aH.mapD[A] = a.mapD[A];
        aH.nH[A] = a.nH[A]; aH.nD[A] = a.nD[A];
        aH.uH[A] = a.uH[A]; aH.rH[A] = a.rH[A]; aH.uD[A] = a.uD[A]; aH.rD[A] = a.rD[A];
        aH.mE[A] = a.mE[A];
    }
    aH.cdt = a.cdt;
    aH.inv_dL = a.inv_dL;
    aH.on = 63u;
    aH.t_probe = -1;
    aED = aH;
    for (int A = 0; A < 3; ++A) {
        aH.Din[A] = a.gC[A]; aH.Hin[A] = a.lH[A]; aH.Hout[A] = a.lH[A]; aH.Eout[A] = a.gC2[A];
        aH.ICE[A] = a.lICE[A]; aH.IH[A] = a.lIH[A];
        aED.Hin[A] = a.gC2[A]; aED.Din[A] = a.lD[A]; aED.Dout[A] = a.lD[A]; aED.Eout[A] = a.gC[A];
        aED.ICH[A] = a.lICH[A]; aED.ID[A] = a.lID[A];
    }
    v5_set_tiles(aH, 0, a.Nx, p->tma_rows, p->xchunk);
    v5_set_tiles(aED, 0, a.Nx, p->tma_rows, p->xchunk);
    if (!p->v5) p->v5 = v5_cache_create();
}

// One part of the transposed step with stored stencil inputs and caller-supplied x-halo planes (x-slabs).
template <typename T, typename AT>
int launch_adjoint_part(cev_fdtd* p, int part, const cev_state* fwd, const cev_adjoint* adj, const void* const halo[3], cudaStream_t s) {
    AdjArgs<T, AT> a;
    if (fill_adj(p, fwd, adj, a)) return -1;
    for (int c = 0; c < 3; ++c)
        if (!a.gC[c]) return fail("cev_fdtd_adjoint_part needs adj->gC");
    const dim3 blk(64, 4);
    const dim3 grd((a.Nz + 63) / 64, (a.Ny + 3) / 4, a.Nx);
    if (grd.y > 65535 || grd.z > 65535) return fail("adjoint kernels: grid extent too large");
    const T* h1 = halo ? (const T*)halo[p->to_logical(1)] : nullptr;
    const T* h2 = halo ? (const T*)halo[p->to_logical(2)] : nullptr;
    if ((h1 == nullptr) != (h2 == nullptr)) return fail("cev_fdtd_adjoint_part: give both halo planes or none");
    if (part == 0) k_adj_Dlocal<T, AT><<<grd, blk, 0, s>>>(a);
    else if (part == 1) k_adj_H_stored<T, AT><<<grd, blk, 0, s>>>(a, h1, h2);
    else k_adj_E_stored<T, AT><<<grd, blk, 0, s>>>(a, h1, h2);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// The reverse sweep of nsteps steps with the tensor-map kernels (adjoint_v5.cuh), cotangents in the eager form between
// the steps.  D[k * 3 + A] = forward D (internal component A) after k steps: full-grid arrays, or (boxed) the
// design-box record.
template <typename T, typename AT>
int adjoint_v5_sweep(cev_fdtd* p, const AdjArgs<T, AT>& a, int64_t nsteps, const double* gbar, const void* const* D, int boxed,
                     cudaStream_t s) {
    StepArgs<T, AT> aH, aED;
    adjoint_v5_args(p, a, aH, aED);
    ProbeTable pr;
    fill_probe_table(p, pr);
    const dim3 blk(64, 4);
    const dim3 grd((a.Nz + 63) / 64, (a.Ny + 3) / 4, a.Nx);
    if (grd.y > 65535 || grd.z > 65535) return fail("adjoint kernels: grid extent too large");
    k_adj_Dlocal<T, AT><<<grd, blk, 0, s>>>(a);
    CUDA_TRY(cudaGetLastError());
    for (int64_t k = nsteps; k >= 1; --k) {
        if (gbar && p->n_slots > 0) {
            k_adj_seed_eager<T, AT><<<p->n_slots, 128, 0, s>>>(a, pr, (const int32_t*)p->pr_owner.p, gbar + (k - 1) * p->nprobe,
                                                               (const T*)D[k * 3], (const T*)D[k * 3 + 1], (const T*)D[k * 3 + 2], boxed);
            CUDA_TRY(cudaGetLastError());
        }
        if (v5_launch_adj_H<T, AT>(p->v5, aH, p->tma_rows, p->tma_stages_adjH, s)) return fail("%s", v5_last_error());
        if (v5_launch_adj_ED<T, AT>(p->v5, aED, D + (k - 1) * 3, a.G, a.gb, k > 1 ? 1 : 0, boxed, p->tma_rows, p->tma_stages_adjED, s))
            return fail("%s", v5_last_error());
    }
    return 0;
}

// One checkpoint segment of the reverse sweep, entirely on the device queue: recompute the forward steps from the
// segment's start state keeping D after every step (hist[k] = D after k steps; hist[0] is the start D), then the
// transposed steps in reverse order, each preceded by the probe-series seeds of its time step.
template <typename T, typename AT>
int adjoint_run_enqueue(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* waveform, const double* gbar,
                        void* const (*hist)[3], const cev_adjoint* adj, cudaStream_t s) {
    const int64_t Nx = p->N[0];
    cev_state cur = *st;
    for (int64_t k = 1; k <= nsteps; ++k) {
        for (int c = 0; c < 3; ++c) cur.D[c] = hist[k - 1][c];
        if (launch_H<T, AT>(p, &cur, nullptr, nullptr, 0, Nx, -1, nullptr, s)) return -1;
        if (launch_D<T, AT>(p, &cur, hist[k], nullptr, nullptr, nullptr, nullptr,
                            waveform ? waveform + (k - 1) * p->nsrc : nullptr, 0, Nx, -1, nullptr, s))
            return -1;
    }
    cev_state fwd;
    memset(&fwd, 0, sizeof fwd);
    for (int c = 0; c < 3; ++c) fwd.inv_eps[c] = st->inv_eps[c];
    for (int c = 0; c < 3; ++c) fwd.D[c] = hist[nsteps][c];
    AdjArgs<T, AT> a;
    if (fill_adj(p, &fwd, adj, a)) return -1;
    const bool seeds = gbar && p->n_slots > 0;
    if (adjoint_v5_ok(p, a)) {
        std::vector<const void*> slots((size_t)(nsteps + 1) * 3);
        for (int64_t k = 0; k <= nsteps; ++k)
            for (int A = 0; A < 3; ++A) slots[k * 3 + A] = hist[k][p->to_logical(A)];
        return adjoint_v5_sweep<T, AT>(p, a, nsteps, seeds ? gbar : nullptr, slots.data(), 0, s);
    }
    for (int64_t k = nsteps; k >= 1; --k) {
        if (seeds) {
            for (int c = 0; c < 3; ++c) fwd.D[c] = hist[k][c];
            if (launch_adjoint_seed<T, AT>(p, &fwd, adj, gbar + (k - 1) * p->nprobe, s)) return -1;
        }
        for (int c = 0; c < 3; ++c) fwd.D[c] = hist[k - 1][c];
        if (launch_adjoint_step<T, AT>(p, &fwd, adj, s)) return -1;
    }
    return 0;
}

template <typename T, typename AT>
int adjoint_run(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* waveform, const double* gbar,
                void* const (*hist)[3], const cev_adjoint* adj, cudaStream_t s) {
    const int64_t cells = (int64_t)p->N[0] * p->N[1] * p->N[2];
    const bool want = !p->halo.on() && !p->rec.buf && nsteps >= 4 &&
                      (p->use_graph == 1 || (p->use_graph < 0 && cells <= GRAPH_MAX_CELLS));
    if (!want) return adjoint_run_enqueue<T, AT>(p, st, nsteps, waveform, gbar, hist, adj, s);
    // everything the captured launches depend on
    std::vector<unsigned char> key;
    auto put = [&](const void* q, size_t n) { key.insert(key.end(), (const unsigned char*)q, (const unsigned char*)q + n); };
    put(st, sizeof *st);
    put(adj, sizeof *adj);
    put(&nsteps, sizeof nsteps);
    put(hist, (size_t)(nsteps + 1) * sizeof hist[0]);
    const int flags = (waveform ? 1 : 0) | (gbar ? 2 : 0);
    put(&flags, sizeof flags);
    auto& G = p->adj_graph;
    if (G.epoch != p->epoch) {
        if (G.exec) cudaGraphExecDestroy(G.exec);
        G = cev_fdtd::AdjGraph();
        G.epoch = p->epoch;
    }
    const size_t wbytes = (size_t)nsteps * std::max(1, p->nsrc) * 8, gbytes = (size_t)nsteps * std::max(1, p->nprobe) * 8;
    if (!(G.exec && G.key == key)) {
        if (G.seen != key) {          // first time: plain launches (they also build tilings / set kernel attributes)
            G.seen = key;
            return adjoint_run_enqueue<T, AT>(p, st, nsteps, waveform, gbar, hist, adj, s);
        }
        // second time: capture
        if (p->adj_stage_w.bytes < wbytes) { p->adj_stage_w.release(); if (p->adj_stage_w.alloc(wbytes)) return -1; }
        if (p->adj_stage_g.bytes < gbytes) { p->adj_stage_g.release(); if (p->adj_stage_g.alloc(gbytes)) return -1; }
        if (!p->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
        if (G.exec) cudaGraphExecDestroy(G.exec);
        G.exec = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(p->cap_stream, cudaStreamCaptureModeRelaxed));
        const int rc = adjoint_run_enqueue<T, AT>(p, st, nsteps, waveform ? (const double*)p->adj_stage_w.p : nullptr,
                                                  gbar ? (const double*)p->adj_stage_g.p : nullptr, hist, adj, p->cap_stream);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(p->cap_stream, &graph);
        if (rc || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc ? rc : fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
        }
        const cudaError_t e2 = cudaGraphInstantiate(&G.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            G.exec = nullptr;
            return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
        }
        G.key = key;
    }
    if (waveform) CUDA_TRY(cudaMemcpyAsync(p->adj_stage_w.p, waveform, (size_t)nsteps * p->nsrc * 8, cudaMemcpyDeviceToDevice, s));
    if (gbar) CUDA_TRY(cudaMemcpyAsync(p->adj_stage_g.p, gbar, (size_t)nsteps * p->nprobe * 8, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaGraphLaunch(G.exec, s));
    G.replays++;
    return 0;
}

// The reverse sweep WITHOUT recomputation: the transposed step is linear in the cotangents and needs the forward
// solution only for dL/d(1/eps) += lE D, which is only wanted inside the design box -- so the forward run records D of
// that box after every step (cev_fdtd_set_recorder) and the sweep reads the record.
template <typename T, typename AT>
int adjoint_run_boxed(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* gbar, const void* rec, const cev_adjoint* adj,
                      cudaStream_t s) {
    cev_state fwd;
    memset(&fwd, 0, sizeof fwd);
    for (int c = 0; c < 3; ++c) {
        fwd.inv_eps[c] = st->inv_eps[c];
        fwd.D[c] = const_cast<void*>(st->inv_eps[c]);       // (fill_adj wants a non-NULL forward D; the sweep reads the record instead)
    }
    AdjArgs<T, AT> a;
    if (fill_adj(p, &fwd, adj, a)) return -1;
    if (!adjoint_v5_ok(p, a)) return fail("cev_fdtd_adjoint_run_boxed needs a grid the tensor-map adjoint kernels serve (cev_fdtd_adjoint_boxed_supported) and adj->gC");
    const int64_t n = (int64_t)(a.gb[1] - a.gb[0]) * (a.gb[3] - a.gb[2]) * (a.gb[5] - a.gb[4]);
    if (n <= 0) return fail("cev_fdtd_adjoint_run_boxed: empty g_box");
    std::vector<const void*> slots((size_t)(nsteps + 1) * 3);
    for (int64_t k = 0; k <= nsteps; ++k)
        for (int A = 0; A < 3; ++A) slots[k * 3 + A] = (const T*)rec + (k * 3 + A) * n;
    return adjoint_v5_sweep<T, AT>(p, a, nsteps, p->n_slots > 0 ? gbar : nullptr, slots.data(), 1, s);
}

#define DISPATCH(plan, fn, ...)                                                       \
    ((plan)->dtype == CEV_F64 ? fn<double, double>(__VA_ARGS__)                       \
                              : ((plan)->arith64 ? fn<float, double>(__VA_ARGS__)     \
                                                 : fn<float, float>(__VA_ARGS__)))

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_range(const cev_fdtd* p, int64_t x0, int64_t x1) {
    // ranges are over the LOGICAL x axis; only plans whose internal x is the logical x may sub-range
    const int64_t nx = p->Nl[0];
    if (x0 < 0 || x1 > nx || x0 > x1) return fail("x-range [%lld,%lld) outside [0,%lld)", (long long)x0, (long long)x1, (long long)nx);
    if (!p->x_is_x() && !(x0 == 0 && x1 == nx)) return fail("partial x-ranges need Ny > 1 or Nz > 1");
    return 0;
}

}  // namespace

extern "C" {

const char* cev_last_error(void) { return g_err.c_str(); }
int cev_abi_version(void) { return CEV_ABI_VERSION; }

int cev_fdtd_create(cev_fdtd** out, int device, int dtype, int arith_f64, int64_t nx, int64_t Ny, int64_t Nz, double dL,
                    double dt, const double* sH[3], const double* sD[3]) {
    if (!out) return fail("plan out-pointer is NULL");
    *out = nullptr;
    if (dtype != CEV_F32 && dtype != CEV_F64) return fail("dtype must be 0 (fp32) or 1 (fp64)");
    if (nx < 1 || Ny < 1 || Nz < 1) return fail("grid extents must be >= 1");
    if (nx * Ny * Nz >= (int64_t)1 << 31) return fail("grid too large for 32-bit cell indexing inside a slab");
    if (!(dL > 0) || !(dt > 0)) return fail("dL and dt must be positive");
    if (!sH || !sD) return fail("sigma profiles are NULL");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d visible)", device, ndev);
    DeviceGuard guard(device);

    cev_fdtd* p = new cev_fdtd();
    p->device = device;
    p->dtype = dtype;
    p->arith64 = (dtype == CEV_F64) ? 1 : (arith_f64 ? 1 : 0);
    p->tma_stages_D = dtype == CEV_F64 ? 4 : 3;      // tuned on B200 (3 CTAs x 4 stages vs 4 CTAs x 3 stages per SM)
    p->Nl[0] = nx;
    p->Nl[1] = Ny;
    p->Nl[2] = Nz;
    p->dL = dL;
    p->dt = dt;
    p->cdt = (1.0 / std::sqrt(EPSILON_0 * MU_0)) * dt;
    // make the last internal axis the contiguous one with extent > 1 (cyclic relabelling keeps the curl)
    if (Nz == 1 && Ny > 1) {              // 2-D: (x, z, y), a swap
        p->perm[1] = 2; p->perm[2] = 1;
        p->parity = -1;
    } else if (Nz == 1 && Ny == 1 && nx > 1) {   // 1-D: (y, z, x), cyclic
        p->perm[0] = 1; p->perm[1] = 2; p->perm[2] = 0;
    }
    for (int A = 0; A < 3; ++A) p->inv[p->perm[A]] = A;
    for (int A = 0; A < 3; ++A) p->N[A] = (int)p->Nl[p->to_logical(A)];

    // tables: per internal axis, for H and D sampling: u (f32,f64), r (f32,f64), map (int)
    size_t total = 0;
    for (int A = 0; A < 3; ++A) total += (size_t)p->N[A] * 2 * (4 + 8 + 4 + 8 + 4) + 2 * 64;   // + alignment slack
    std::vector<unsigned char> host(total);
    if (p->tables.alloc(total)) {
        delete p;
        return -1;
    }
    size_t off = 0;
    auto put = [&](const void* src, size_t bytes) -> const void* {
        memcpy(host.data() + off, src, bytes);
        const void* dev = (const unsigned char*)p->tables.p + off;
        off += bytes;
        return dev;
    };
    for (int A = 0; A < 3; ++A) {
        const int L = p->to_logical(A);
        const int n = p->N[A];
        for (int kind = 0; kind < 2; ++kind) {
            const double* sig = kind == 0 ? sH[L] : sD[L];
            if (!sig) {
                delete p;
                return fail("sigma profile for axis %d is NULL", L);
            }
            std::vector<double> u(n), r(n);
            std::vector<float> uf(n), rf(n);
            std::vector<int> map(n);
            int cnt = 0;
            for (int q = 0; q < n; ++q) {
                if (!(sig[q] >= 0) || !std::isfinite(sig[q])) {
                    delete p;
                    return fail("sigma profile must be finite and >= 0");
                }
                u[q] = sig[q] * dt / (2 * EPSILON_0);
                r[q] = 1.0 / (1.0 + u[q]);
                uf[q] = (float)u[q];
                rf[q] = (float)r[q];
                map[q] = (sig[q] != 0.0) ? cnt++ : -1;
            }
            const void* duf = put(uf.data(), n * 4);
            // keep 8-byte alignment for the doubles
            if (off % 8) off += 8 - off % 8;
            const void* dud = put(u.data(), n * 8);
            const void* drf = put(rf.data(), n * 4);
            if (off % 8) off += 8 - off % 8;
            const void* drd = put(r.data(), n * 8);
            const void* dmap = put(map.data(), n * 4);
            if (off % 8) off += 8 - off % 8;
            if (kind == 0) {
                p->uH[A][0] = duf; p->uH[A][1] = dud; p->rH[A][0] = drf; p->rH[A][1] = drd;
                p->mapH[A] = (const int*)dmap; p->nH[A] = cnt;
            } else {
                p->uD[A][0] = duf; p->uD[A][1] = dud; p->rD[A][0] = drf; p->rD[A][1] = drd;
                p->mapD[A] = (const int*)dmap; p->nD[A] = cnt;
            }
        }
    }
    for (int A = 0; A < 3; ++A) {   // longest run of indices where neither the H nor the D profile is in the PML
        const int L = p->to_logical(A);
        int best_lo = 0, best_hi = 0, run_lo = 0;
        for (int q = 0; q <= p->N[A]; ++q) {
            const bool zero = q < p->N[A] && sH[L][q] == 0.0 && sD[L][q] == 0.0;
            if (!zero) {
                if (q - run_lo > best_hi - best_lo) { best_lo = run_lo; best_hi = q; }
                run_lo = q + 1;
            }
        }
        p->in_lo[A] = best_lo;
        p->in_hi[A] = best_hi;
    }
    if (off > total) {
        delete p;
        return fail("internal: table overflow");
    }
    cudaError_t e = cudaMemcpy(p->tables.p, host.data(), off, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        p->tables.release();
        delete p;
        return fail("table upload failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return 0;
}

int cev_fdtd_destroy(cev_fdtd* p) {
    if (!p) return 0;
    DeviceGuard guard(p->device);
    p->drop_graphs();
    p->stage_w.release(); p->stage_p.release();
    p->adj_stage_w.release(); p->adj_stage_g.release();
    p->batch_tab_H.release(); p->batch_tab_D.release();
    p->batch_tab_F0.release(); p->batch_tab_F1.release(); p->tan_shadow.release();
    if (p->cap_stream) cudaStreamDestroy(p->cap_stream);
    for (auto q : p->side) cudaStreamDestroy(q);
    for (auto e : p->side_ev) cudaEventDestroy(e);
    if (p->main_ev) cudaEventDestroy(p->main_ev);
    if (p->v5) v5_cache_destroy(p->v5);
    p->tables.release();
    p->src_comp.release(); p->src_id.release(); p->src_cell.release(); p->src_weight.release();
    for (auto& t : p->src_tilings) {
        t->begin.release(); t->comp.release(); t->id.release(); t->cell.release(); t->w.release();
    }
    p->pr_field.release(); p->pr_wbegin.release(); p->pr_ibegin.release(); p->pr_cell0.release();
    p->pr_n.release(); p->pr_idx.release(); p->pr_weight.release(); p->pr_owner.release();
    p->mon_field.release(); p->mon_cell.release();
    delete p;
    return 0;
}

int cev_fdtd_set_option(cev_fdtd* p, const char* name, int64_t value) {
    if (!p || !name) return fail("NULL argument");
    p->epoch++;
    if (!strcmp(name, "jvp_fused")) {
        if (value < -1 || value > 1) return fail("jvp_fused must be -1 (auto), 0 or 1");
        p->jvp_fused = (int)value;
    } else if (!strcmp(name, "jvp_batch")) {
        if (value < -1 || value > 1) return fail("jvp_batch must be -1 (auto), 0 or 1");
        p->jvp_batch = (int)value;
    } else if (!strcmp(name, "jvp_streams")) {
        if (value < -1 || value > 1) return fail("jvp_streams must be -1 (auto), 0 or 1");
        p->jvp_streams = (int)value;
    } else if (!strcmp(name, "use_graph")) {
        if (value < -1 || value > 1) return fail("use_graph must be -1 (auto: small grids), 0 or 1");
        p->use_graph = (int)value;
    } else if (!strcmp(name, "kernel_variant")) {
        if (value < 0 || value > 6)
            return fail("kernel_variant must be 0 (auto), 1 (baseline), 2 (marching), 3 (TMA-staged), 4 (fused), 5 (hybrid) or 6 (tensor-map TMA)");
        p->variant = (int)value;
    } else if (!strcmp(name, "tma_rows")) {
        if (!v5_supported_shape((int)value, 3)) return fail("tma_rows must be 4 or 8");
        p->tma_rows = (int)value;
    } else if (!strcmp(name, "tma_stages") || !strcmp(name, "tma_stages_H") || !strcmp(name, "tma_stages_D")) {
        if (!v5_supported_shape(p->tma_rows, (int)value)) return fail("%s must be 3 or 4", name);
        const char which = name[10] ? name[11] : 0;             // "tma_stages" sets both
        if (which != 'D') p->tma_stages_H = (int)value;
        if (which != 'H') p->tma_stages_D = (int)value;
    } else if (!strcmp(name, "adjoint_variant")) {
        if (value < 0 || value > 2) return fail("adjoint_variant must be 0 (auto), 1 (simple kernels) or 2 (tensor-map kernels)");
        p->adjoint_variant = (int)value;
    } else if (!strcmp(name, "tma_stages_adjH") || !strcmp(name, "tma_stages_adjED")) {
        if (!v5_supported_shape(p->tma_rows, (int)value)) return fail("%s must be 3 or 4", name);
        (name[14] == 'H' ? p->tma_stages_adjH : p->tma_stages_adjED) = (int)value;
    } else if (!strcmp(name, "halo_pause")) {
        if (value != 0 && value != 1) return fail("halo_pause must be 0 or 1");
        p->halo.paused = value != 0;
    } else if (!strcmp(name, "auto_tensor_map")) {
        if (value != 0 && value != 1) return fail("auto_tensor_map must be 0 or 1");
        p->auto_v5 = (int)value;
    } else if (!strcmp(name, "fused_shape")) {
        if (value != 0 && value != 804 && value != 1604 && value != 1608 && value != 3204 && value != 3208)
            return fail("fused_shape must be 0 (auto) or one of 804, 1604, 1608, 3204, 3208 (lanes_z*100 + warps)");
        p->fused_shape = (int)value;
    } else if (!strcmp(name, "prefetch_planes")) {
        if (value < 0 || value > 64) return fail("prefetch_planes must be in [0, 64]");
        p->pf_dist = (int)value;
    } else if (!strcmp(name, "active_components")) {
        // logical bits 0-2: D/E x,y,z may be non-zero; bits 3-5: H x,y,z.  The caller guarantees the others are and stay 0.
        if (value < 0 || value > 63) return fail("active_components is a 6-bit mask");
        unsigned m = 0;
        for (int c = 0; c < 3; ++c) {
            if (value & (1 << c)) m |= 1u << p->to_internal(c);
            if (value & (8 << c)) m |= 8u << p->to_internal(c);
        }
        p->on = m;
    } else if (!strcmp(name, "split_launch")) {
        if (value != 0 && value != 1) return fail("split_launch must be 0 or 1");
        p->split = (int)value;
    } else if (!strcmp(name, "lanes_z")) {
        if (value != 8 && value != 16 && value != 32) return fail("lanes_z must be 8, 16 or 32");
        p->lz = (int)value;
    } else if (!strcmp(name, "xchunk")) {
        if (value < 0) return fail("xchunk must be >= 0");
        p->xchunk = (int)value;
    } else {
        return fail("unknown option '%s'", name);
    }
    return 0;
}

int cev_fdtd_pml_shapes(const cev_fdtd* p, int64_t shapes[12][3]) {
    if (!p || !shapes) return fail("NULL argument");
    for (int fam = 0; fam < 4; ++fam) {
        const bool isH = fam < 2;
        const bool self = (fam == 1 || fam == 3);   // IH / ID: compact on the two OTHER axes
        for (int c = 0; c < 3; ++c) {
            for (int ax = 0; ax < 3; ++ax) {
                const int cnt = isH ? p->nH[p->to_internal(ax)] : p->nD[p->to_internal(ax)];
                const bool compact = self ? (ax != c) : (ax == c);
                shapes[fam * 3 + c][ax] = compact ? cnt : p->Nl[ax];
            }
        }
    }
    return 0;
}

int cev_fdtd_step_H(cev_fdtd* p, const cev_state* st, void* const H_out[3], int64_t x0, int64_t x1, void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (check_range(p, x0, x1)) return -1;
    DeviceGuard guard(p->device);
    const int64_t a0 = p->x_is_x() ? x0 : 0, a1 = p->x_is_x() ? x1 : p->N[0];
    return DISPATCH(p, launch_H, p, st, (const cev_tangent*)nullptr, H_out, a0, a1, (int64_t)-1, (double*)nullptr, (cudaStream_t)stream);
}

int cev_fdtd_step_H_ex(cev_fdtd* p, const cev_state* st, const cev_tangent* tan, void* const H_out[3], int64_t x0,
                       int64_t x1, int64_t probe_t, double* partials, void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (check_range(p, x0, x1)) return -1;
    DeviceGuard guard(p->device);
    const int64_t a0 = p->x_is_x() ? x0 : 0, a1 = p->x_is_x() ? x1 : p->N[0];
    return DISPATCH(p, launch_H, p, st, tan, H_out, a0, a1, probe_t, partials, (cudaStream_t)stream);
}

int cev_fdtd_step_D_ex(cev_fdtd* p, const cev_state* st, void* const D_out[3], void* const E_out[3], const void* const J[3],
                       const double J_scale[3], const double* waveform_row, int64_t x0, int64_t x1, int64_t probe_t,
                       double* partials, void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (check_range(p, x0, x1)) return -1;
    DeviceGuard guard(p->device);
    const int64_t a0 = p->x_is_x() ? x0 : 0, a1 = p->x_is_x() ? x1 : p->N[0];
    return DISPATCH(p, launch_D, p, st, D_out, E_out, J, J_scale, (const double* const*)nullptr, waveform_row, a0, a1,
                    probe_t, partials, (cudaStream_t)stream);
}

int cev_fdtd_sample_probes(cev_fdtd* p, const cev_state* st, const cev_tangent* tan, int which, int64_t t, double* partials,
                           void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (which != 0 && which != 1) return fail("which must be 0 (E/D probes) or 1 (H probes)");
    DeviceGuard guard(p->device);
    return DISPATCH(p, launch_probe_only, p, st, tan, which, t, partials, (cudaStream_t)stream);
}

int cev_fdtd_jvp_run(cev_fdtd* p, const cev_state* st, int B, const cev_state* tangents, const cev_tangent* tans,
                     int64_t nsteps, const double* waveform, double* partials, double* tangent_partials, void* stream) {
    if (!p || !st || B < 0 || (B > 0 && (!tangents || !tans))) return fail("bad jvp arguments");
    if (nsteps < 0) return fail("nsteps must be >= 0");
    if (nsteps == 0) return 0;
    if (p->n_src_pts > 0 && !waveform) return fail("plan has sources but waveform is NULL");
    if (p->n_slots > 0 && (!partials || (B > 0 && !tangent_partials))) return fail("plan has probes but partials is NULL");
    DeviceGuard guard(p->device);
    return DISPATCH(p, jvp_loop, p, st, B, tangents, tans, nsteps, p->n_src_pts > 0 ? waveform : nullptr, partials,
                    tangent_partials, (cudaStream_t)stream);
}

int cev_fdtd_adjoint_step(cev_fdtd* p, const cev_state* fwd, const cev_adjoint* adj, void* stream) {
    if (!p || !fwd || !adj) return fail("NULL argument");
    DeviceGuard guard(p->device);
    return DISPATCH(p, launch_adjoint_step, p, fwd, adj, (cudaStream_t)stream);
}

int cev_fdtd_adjoint_run(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* waveform, const double* gbar,
                         void* const (*D_hist)[3], const cev_adjoint* adj, void* stream) {
    if (!p || !st || !D_hist || !adj) return fail("NULL argument");
    if (nsteps < 0) return fail("nsteps must be >= 0");
    if (nsteps == 0) return 0;
    if (p->n_src_pts > 0 && !waveform) return fail("plan has sources but waveform is NULL");
    if (st->D_xhi[1] || st->D_xhi[2] || st->H_xlo[1] || st->H_xlo[2] || p->halo.on())
        return fail("cev_fdtd_adjoint_run steps a whole (periodic) grid");
    for (int64_t k = 0; k <= nsteps; ++k)
        for (int c = 0; c < 3; ++c)
            if (!D_hist[k][c]) return fail("D_hist needs nsteps + 1 slots of three arrays");
    DeviceGuard guard(p->device);
    return DISPATCH(p, adjoint_run, p, st, nsteps, p->n_src_pts > 0 ? waveform : nullptr, gbar, D_hist, adj, (cudaStream_t)stream);
}

// series[t, p] = sum of the partial sums of probe p's slots, in slot order: a fixed summation order whatever the
// number of rows (a GEMM with the 0/1 fold matrix picks its tiling, hence its order, by shape)
__global__ void k_fold_slots(const double* __restrict__ partials, const int32_t* __restrict__ owner, int n_slots, int n_probes,
                             int64_t rows, double* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= rows * n_probes) return;
    const int64_t t = q / n_probes;
    const int pr = (int)(q % n_probes);
    double acc = 0.0;
    for (int sl = 0; sl < n_slots; ++sl)
        if (owner[sl] == pr) acc += partials[t * n_slots + sl];
    out[q] = acc;
}

int cev_fdtd_fold_probes(cev_fdtd* p, const double* partials, int64_t rows, double* series, void* stream) {
    if (!p || !series) return fail("NULL argument");
    if (rows < 0) return fail("rows must be >= 0");
    if (rows == 0 || p->nprobe == 0) return 0;
    if (!partials && p->n_slots > 0) return fail("partials is NULL");
    DeviceGuard guard(p->device);
    const int64_t n = rows * p->nprobe;
    k_fold_slots<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(partials, (const int32_t*)p->pr_owner.p, p->n_slots,
                                                                                p->nprobe, rows, series);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int64_t cev_fdtd_adjoint_graph_replays(const cev_fdtd* p) { return p ? p->adj_graph.replays : -1; }

int cev_fdtd_adjoint_boxed_supported(const cev_fdtd* p) {
    if (!p) return 0;
    const int V = p->dtype == CEV_F64 ? 2 : 4;
    if (p->adjoint_variant == 1 || p->perm[0] != 0 || p->perm[1] != 1 || p->perm[2] != 2 || p->on != 63u) return 0;
    if (p->N[1] < p->tma_rows || p->N[1] % p->tma_rows != 0 || p->N[2] % V != 0 || p->N[2] < 32 * V) return 0;
    if (p->adjoint_variant != 2 && (int64_t)p->N[1] * p->N[2] < (1 << 14)) return 0;
    return 1;
}

int cev_fdtd_set_recorder(cev_fdtd* p, const int64_t box[6], void* buf, int64_t capacity) {
    if (!p) return fail("NULL argument");
    if (!buf || !box) {
        p->rec = cev_fdtd::Recorder();
        return 0;
    }
    if (p->perm[0] != 0 || p->perm[1] != 1 || p->perm[2] != 2) return fail("the D-box recorder needs a 3-D grid");
    if (capacity <= 0) return fail("recorder capacity must be positive");
    for (int A = 0; A < 3; ++A)
        if (box[2 * A] < 0 || box[2 * A + 1] > p->Nl[A] || box[2 * A] >= box[2 * A + 1]) return fail("recorder box outside the grid");
    p->rec.buf = buf;
    p->rec.capacity = capacity;
    p->rec.count = 0;
    for (int q = 0; q < 6; ++q) p->rec.box[q] = (int)box[q];
    return 0;
}

int cev_fdtd_adjoint_run_boxed(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* gbar, const void* D_box_record,
                               const cev_adjoint* adj, void* stream) {
    if (!p || !st || !D_box_record || !adj) return fail("NULL argument");
    if (nsteps < 0) return fail("nsteps must be >= 0");
    if (nsteps == 0) return 0;
    if (p->halo.on()) return fail("cev_fdtd_adjoint_run_boxed steps a whole (periodic) grid");
    DeviceGuard guard(p->device);
    return DISPATCH(p, adjoint_run_boxed, p, st, nsteps, gbar, D_box_record, adj, (cudaStream_t)stream);
}

int cev_fdtd_adjoint_part(cev_fdtd* p, int part, const cev_state* fwd, const cev_adjoint* adj, const void* const halo[3],
                          void* stream) {
    if (!p || !fwd || !adj) return fail("NULL argument");
    if (part < 0 || part > 2) return fail("part must be 0 (cell-local D part), 1 (H part) or 2 (E part)");
    if (halo && !p->x_is_x()) return fail("x-halo planes need Ny > 1 or Nz > 1 (no slab decomposition of a 1-D grid)");
    DeviceGuard guard(p->device);
    return DISPATCH(p, launch_adjoint_part, p, part, fwd, adj, halo, (cudaStream_t)stream);
}

int cev_fdtd_adjoint_seed(cev_fdtd* p, const cev_state* fwd, const cev_adjoint* adj, const double* gbar_row, void* stream) {
    if (!p || !fwd || !adj || !gbar_row) return fail("NULL argument");
    DeviceGuard guard(p->device);
    return DISPATCH(p, launch_adjoint_seed, p, fwd, adj, gbar_row, (cudaStream_t)stream);
}

int cev_fdtd_step_D(cev_fdtd* p, const cev_state* st, void* const D_out[3], void* const E_out[3], const void* const J[3],
                    const double J_scale[3], int64_t x0, int64_t x1, void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (check_range(p, x0, x1)) return -1;
    DeviceGuard guard(p->device);
    const int64_t a0 = p->x_is_x() ? x0 : 0, a1 = p->x_is_x() ? x1 : p->N[0];
    return DISPATCH(p, launch_D, p, st, D_out, E_out, J, J_scale, (const double* const*)nullptr, (const double*)nullptr, a0,
                    a1, (int64_t)-1, (double*)nullptr, (cudaStream_t)stream);
}

int cev_fdtd_compute_E(cev_fdtd* p, const cev_state* st, const cev_tangent* tan, void* const E_out[3], void* stream) {
    if (!p || !st || !E_out) return fail("NULL argument");
    DeviceGuard guard(p->device);
    return DISPATCH(p, launch_compute_E, p, st, tan, E_out, (cudaStream_t)stream);
}

int cev_fdtd_set_sources(cev_fdtd* p, int nsrc, const cev_points* src) {
    if (!p || nsrc < 0 || (nsrc > 0 && !src)) return fail("bad source arguments");
    DeviceGuard guard(p->device);
    const int64_t ncell = p->Nl[0] * p->Nl[1] * p->Nl[2];
    int64_t total = 0;
    for (int s = 0; s < nsrc; ++s) {
        if (src[s].field < CEV_FIELD_D || src[s].field >= CEV_FIELD_D + 3) return fail("source %d: field must be a D component (3..5)", s);
        if (src[s].n < 0 || (src[s].n > 0 && (!src[s].idx || !src[s].weight))) return fail("source %d: needs idx and weight arrays", s);
        if (src[s].n > ncell) return fail("source %d: more points than cells", s);
        total += src[s].n;
    }
    // everything is fetched and validated on the host BEFORE the plan is touched: a failing call leaves the
    // previous sources (and their tilings) in place
    std::vector<int32_t> comp(total), id(total);
    std::vector<int64_t> cell(total);
    std::vector<double> w(total);
    int64_t off = 0;
    for (int s = 0; s < nsrc; ++s) {
        for (int64_t q = 0; q < src[s].n; ++q) {
            comp[off + q] = p->to_internal(src[s].field - CEV_FIELD_D);
            id[off + q] = s;
        }
        if (src[s].n) {
            CUDA_TRY(cudaMemcpy(cell.data() + off, src[s].idx, src[s].n * 8, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(w.data() + off, src[s].weight, src[s].n * 8, cudaMemcpyDeviceToHost));
        }
        off += src[s].n;
    }
    for (int64_t q = 0; q < total; ++q)
        if (cell[q] < 0 || cell[q] >= ncell) return fail("source point outside the grid");
    {   // sorted by (component, cell), source order kept inside equal keys: what inject_points (common.cuh) expects
        std::vector<int64_t> order(total);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
            return comp[x] != comp[y] ? comp[x] < comp[y] : cell[x] < cell[y];
        });
        std::vector<int32_t> comp2(total), id2(total);
        std::vector<int64_t> cell2(total);
        std::vector<double> w2(total);
        for (int64_t r = 0; r < total; ++r) {
            comp2[r] = comp[order[r]]; id2[r] = id[order[r]]; cell2[r] = cell[order[r]]; w2[r] = w[order[r]];
        }
        comp.swap(comp2); id.swap(id2); cell.swap(cell2); w.swap(w2);
    }
    DeviceBuf d_comp, d_id, d_cell, d_w;
    if (d_comp.alloc(total * 4) || d_id.alloc(total * 4) || d_cell.alloc(total * 8) || d_w.alloc(total * 8)) return -1;
    if (total) {
        CUDA_TRY(cudaMemcpy(d_comp.p, comp.data(), total * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_id.p, id.data(), total * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_cell.p, cell.data(), total * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d_w.p, w.data(), total * 8, cudaMemcpyHostToDevice));
    }
    // commit
    p->epoch++;
    p->nsrc = nsrc;
    p->n_src_pts = total;
    std::swap(p->src_comp.p, d_comp.p); std::swap(p->src_comp.bytes, d_comp.bytes);
    std::swap(p->src_id.p, d_id.p); std::swap(p->src_id.bytes, d_id.bytes);
    std::swap(p->src_cell.p, d_cell.p); std::swap(p->src_cell.bytes, d_cell.bytes);
    std::swap(p->src_weight.p, d_w.p); std::swap(p->src_weight.bytes, d_w.bytes);
    p->h_src_comp.swap(comp);
    p->h_src_id.swap(id);
    p->h_src_cell.swap(cell);
    p->h_src_w.swap(w);
    p->src_tilings.clear();          // (DeviceBuf destructors free the device copies)
    return 0;
}

int cev_fdtd_set_probes(cev_fdtd* p, int nprobe, const cev_points* probe, int64_t* n_slots_out) {
    if (!p || nprobe < 0 || (nprobe > 0 && !probe)) return fail("bad probe arguments");
    DeviceGuard guard(p->device);
    const int64_t ncell = p->Nl[0] * p->Nl[1] * p->Nl[2];
    std::vector<int32_t> field, owner;
    std::vector<int64_t> wbegin, ibegin, cell0, cnt;
    int64_t wtotal = 0, itotal = 0;
    std::vector<int64_t> woff(nprobe), ioff(nprobe);
    for (int q = 0; q < nprobe; ++q) {
        const cev_points& P = probe[q];
        if (P.field < 0 || P.field > 8) return fail("probe %d: field code must be 0..8", q);
        if (P.n < 0 || (P.n > 0 && !P.weight)) return fail("probe %d: needs a weight array", q);
        if (!P.idx && (P.cell0 < 0 || P.cell0 + P.n > ncell)) return fail("probe %d: dense range outside the grid", q);
        if (P.idx && P.n > 0) {      // sparse point sets are read and scattered to (adjoint seeds) unchecked by the kernels
            std::vector<int64_t> cells(P.n);
            CUDA_TRY(cudaMemcpy(cells.data(), P.idx, P.n * 8, cudaMemcpyDeviceToHost));
            for (int64_t r = 0; r < P.n; ++r)
                if (cells[r] < 0 || cells[r] >= ncell) return fail("probe %d: point outside the grid", q);
        }
        woff[q] = wtotal;
        ioff[q] = P.idx ? itotal : -1;
        wtotal += P.n;
        if (P.idx) itotal += P.n;
    }
    // E/D-family slots first, then H-family: each family is sampled by a different launch
    int nED = 0;
    for (int fam = 0; fam < 2; ++fam) {
        for (int q = 0; q < nprobe; ++q) {
            const cev_points& P = probe[q];
            const bool isH = P.field >= CEV_FIELD_H;
            if ((fam == 0) == isH) continue;
            const int fcode = (P.field / 3) * 3 + p->to_internal(P.field % 3);
            int64_t done = 0;
            do {   // at least one slot per probe so that every series column is written
                const int64_t m = (P.n - done < PROBE_CHUNK) ? (P.n - done) : PROBE_CHUNK;
                field.push_back(fcode);
                owner.push_back(q);
                wbegin.push_back(woff[q] + done);
                ibegin.push_back(P.idx ? ioff[q] + done : -1);
                cell0.push_back(P.idx ? 0 : P.cell0 + done);
                cnt.push_back(m);
                done += m;
            } while (done < P.n);
        }
        if (fam == 0) nED = (int)field.size();
    }
    const int ns = (int)field.size();
    if (p->pr_field.alloc(ns * 4) || p->pr_wbegin.alloc(ns * 8) || p->pr_ibegin.alloc(ns * 8) || p->pr_cell0.alloc(ns * 8) ||
        p->pr_n.alloc(ns * 8) || p->pr_idx.alloc(itotal * 8) || p->pr_weight.alloc(wtotal * 8) || p->pr_owner.alloc(ns * 4))
        return -1;
    if (ns) {
        CUDA_TRY(cudaMemcpy(p->pr_field.p, field.data(), ns * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->pr_wbegin.p, wbegin.data(), ns * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->pr_ibegin.p, ibegin.data(), ns * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->pr_cell0.p, cell0.data(), ns * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->pr_n.p, cnt.data(), ns * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->pr_owner.p, owner.data(), ns * 4, cudaMemcpyHostToDevice));
    }
    for (int q = 0; q < nprobe; ++q) {
        const cev_points& P = probe[q];
        if (P.n == 0) continue;
        CUDA_TRY(cudaMemcpy((double*)p->pr_weight.p + woff[q], P.weight, P.n * 8, cudaMemcpyDeviceToDevice));
        if (P.idx) CUDA_TRY(cudaMemcpy((int64_t*)p->pr_idx.p + ioff[q], P.idx, P.n * 8, cudaMemcpyDeviceToDevice));
    }
    p->epoch++;
    p->nprobe = nprobe;
    p->n_slots = ns;
    p->n_slots_ED = nED;
    p->slot_probe = owner;
    if (n_slots_out) *n_slots_out = ns;
    return 0;
}

int cev_fdtd_set_monitors(cev_fdtd* p, int nmon, const cev_points* mon, int nfreq, int64_t* n_points) {
    if (!p || nmon < 0 || (nmon > 0 && !mon) || nfreq < 0) return fail("bad monitor arguments");
    DeviceGuard guard(p->device);
    const int64_t ncell = p->Nl[0] * p->Nl[1] * p->Nl[2];
    int64_t total = 0;
    for (int m = 0; m < nmon; ++m) {
        if (mon[m].field < 0 || mon[m].field > 8) return fail("monitor %d: field code must be 0..8", m);
        if (mon[m].n < 0 || (mon[m].n > 0 && !mon[m].idx)) return fail("monitor %d: needs an index array", m);
        total += mon[m].n;
    }
    p->epoch++;
    p->mon_phasors = nullptr;
    p->mon_acc = nullptr;
    p->n_mon_pts = total;
    p->mon_nfreq = nfreq;
    if (p->mon_field.alloc(total * 4) || p->mon_cell.alloc(total * 8)) return -1;
    int64_t off = 0;
    for (int m = 0; m < nmon; ++m) {
        if (mon[m].n == 0) continue;
        std::vector<int32_t> f(mon[m].n, (mon[m].field / 3) * 3 + p->to_internal(mon[m].field % 3));
        std::vector<int64_t> cells(mon[m].n);
        CUDA_TRY(cudaMemcpy(cells.data(), mon[m].idx, mon[m].n * 8, cudaMemcpyDeviceToHost));
        for (int64_t q = 0; q < mon[m].n; ++q)
            if (cells[q] < 0 || cells[q] >= ncell) return fail("monitor %d: point outside the grid", m);
        CUDA_TRY(cudaMemcpy((int32_t*)p->mon_field.p + off, f.data(), mon[m].n * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy((int64_t*)p->mon_cell.p + off, mon[m].idx, mon[m].n * 8, cudaMemcpyDeviceToDevice));
        off += mon[m].n;
    }
    if (n_points) *n_points = total;
    return 0;
}

int cev_fdtd_bind_monitors(cev_fdtd* p, const double* phasors, double* acc) {
    if (!p) return fail("NULL argument");
    if ((phasors == nullptr) != (acc == nullptr)) return fail("phasors and acc must both be given or both be NULL");
    p->mon_phasors = phasors;
    p->mon_acc = acc;
    return 0;
}

int cev_fdtd_probe_slots(const cev_fdtd* p, int32_t* slot_probe) {
    if (!p || !slot_probe) return fail("NULL argument");
    for (int s = 0; s < p->n_slots; ++s) slot_probe[s] = p->slot_probe[s];
    return 0;
}

int cev_fdtd_run(cev_fdtd* p, const cev_state* st, int64_t nsteps, const double* waveform, double* partials, void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (nsteps < 0) return fail("nsteps must be >= 0");
    if (nsteps == 0) return 0;
    if (p->n_src_pts > 0 && !waveform) return fail("plan has sources but waveform is NULL");
    if (p->n_slots > 0 && !partials) return fail("plan has probes but partials is NULL");
    if (st->D_xhi[1] || st->D_xhi[2] || st->H_xlo[1] || st->H_xlo[2])
        return fail("cev_fdtd_run steps a whole (periodic) grid, or an x-slab whose halos are attached with cev_fdtd_halo_attach; "
                    "slabs with caller-owned halo planes are driven per half-step");
    DeviceGuard guard(p->device);
    return DISPATCH(p, run_loop, p, st, nsteps, p->n_src_pts > 0 ? waveform : nullptr, partials, (cudaStream_t)stream);
}

int cev_fdtd_run_fused(cev_fdtd* p, const cev_state* st, const cev_state* shadow, int64_t nsteps, const double* waveform,
                       double* partials, void* stream) {
    if (!p || !st || !shadow) return fail("NULL argument");
    if (nsteps < 0) return fail("nsteps must be >= 0");
    if (nsteps == 0) return 0;
    if (p->n_src_pts > 0 && !waveform) return fail("plan has sources but waveform is NULL");
    if (p->n_slots > 0 && !partials) return fail("plan has probes but partials is NULL");
    if (st->D_xhi[1] || st->D_xhi[2] || st->H_xlo[1] || st->H_xlo[2]) return fail("cev_fdtd_run_fused steps a whole (periodic) grid; slabs are driven per half-step");
    if (p->rec.buf) return fail("the D-box recorder works with cev_fdtd_run, not cev_fdtd_run_fused");
    DeviceGuard guard(p->device);
    return DISPATCH(p, run_loop_fused, p, st, shadow, nsteps, p->n_src_pts > 0 ? waveform : nullptr, partials,
                    (cudaStream_t)stream);
}

// ---- x-slab halo exchange through peer-mapped memory ------------------------------------------------------
int cev_fdtd_halo_layout(const cev_fdtd* p, cev_halo_layout* lay) {
    if (!p || !lay) return fail("NULL argument");
    halo_layout(p, lay);
    return 0;
}

int cev_halo_alloc(int device, size_t bytes, void** block, unsigned char ipc_handle[CEV_IPC_HANDLE_BYTES]) {
    if (!block || bytes == 0) return fail("bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == CEV_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(device);
    void* q = nullptr;
    CUDA_TRY(cudaMalloc(&q, bytes));
    if (cudaMemset(q, 0, bytes) != cudaSuccess) {
        cudaFree(q);
        return fail("cudaMemset of the halo block failed");
    }
    if (ipc_handle) {
        cudaIpcMemHandle_t h;
        const cudaError_t e = cudaIpcGetMemHandle(&h, q);
        if (e != cudaSuccess) {
            cudaFree(q);
            return fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        }
        memcpy(ipc_handle, &h, sizeof h);
    }
    *block = q;
    return 0;
}

int cev_halo_open(int device, const unsigned char ipc_handle[CEV_IPC_HANDLE_BYTES], void** block) {
    if (!ipc_handle || !block) return fail("NULL argument");
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    CUDA_TRY(cudaIpcOpenMemHandle(block, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int cev_halo_close(int device, void* block) {
    if (!block) return 0;
    DeviceGuard guard(device);
    CUDA_TRY(cudaIpcCloseMemHandle(block));
    return 0;
}

int cev_halo_free(int device, void* block) {
    if (!block) return 0;
    DeviceGuard guard(device);
    CUDA_TRY(cudaFree(block));
    return 0;
}

int cev_fdtd_halo_attach(cev_fdtd* p, void* own, void* left, void* right) {
    if (!p) return fail("NULL argument");
    if (!own && !left && !right) {
        p->halo = cev_fdtd::Halo();
        return 0;
    }
    if (!own || !left || !right) return fail("own, left and right exchange blocks must all be given (or all NULL to detach)");
    if (!p->x_is_x() || p->Nl[1] == 1 || p->Nl[2] == 1) return fail("x-slab halos need a 3-D grid");
    p->halo.own = (unsigned char*)own;
    p->halo.left = (unsigned char*)left;
    p->halo.right = (unsigned char*)right;
    p->halo.nH = p->halo.nD = 0;
    return 0;
}

int cev_fdtd_halo_push_static(cev_fdtd* p, const cev_state* st, void* stream) {
    if (!p || !st) return fail("NULL argument");
    if (!p->halo.own) return fail("no exchange block attached");
    DeviceGuard guard(p->device);
    cev_halo_layout lay;
    halo_layout(p, &lay);
    const size_t es = p->dtype == CEV_F64 ? 8 : 4;
    const size_t nb = (size_t)p->Nl[1] * p->Nl[2] * es;
    for (int c = 1; c < 3; ++c) {       // my plane 0 of 1/eps_y, 1/eps_z is the LEFT neighbour's plane nx
        if (!st->inv_eps[c]) return fail("cev_state: inv_eps missing");
        CUDA_TRY(cudaMemcpyAsync(p->halo.left + lay.inv_eps_hi[c - 1], st->inv_eps[c], nb, cudaMemcpyDefault, (cudaStream_t)stream));
    }
    return 0;
}

int cev_fdtd_halo_reset(cev_fdtd* p, void* stream) {
    if (!p) return fail("NULL argument");
    if (!p->halo.own) return fail("no exchange block attached");
    DeviceGuard guard(p->device);
    cev_halo_layout lay;
    halo_layout(p, &lay);
    for (int c = 0; c < 2; ++c) {
        CUDA_TRY(cudaMemsetAsync(p->halo.own + lay.D_hi[c], 0, lay.plane_bytes, (cudaStream_t)stream));
        CUDA_TRY(cudaMemsetAsync(p->halo.own + lay.H_lo[c], 0, lay.plane_bytes, (cudaStream_t)stream));
    }
    return 0;
}

int cev_fdtd_halo_error(cev_fdtd* p, int* err) {
    if (!p || !err) return fail("NULL argument");
    if (!p->halo.own) return fail("no exchange block attached");
    DeviceGuard guard(p->device);
    cev_halo_layout lay;
    halo_layout(p, &lay);
    CUDA_TRY(cudaMemcpy(err, p->halo.own + lay.err, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
