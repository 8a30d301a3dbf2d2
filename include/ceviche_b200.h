/* ceviche_b200 -- C ABI of the B200-native FDTD time-stepping engine.
 *
 * Drop-in boundary for ONE path of fancompute/ceviche: the body of
 * ceviche.fdtd.forward() (reference ceviche/fdtd.py:74-144: curl_E / curl_H of
 * ceviche/derivatives.py:16-30, the sigma-PML update of fdtd.py:85-122, the J
 * injection of fdtd.py:125-127 and E = D/eps of fdtd.py:135-137), its caller loop
 * (ceviche/utils.py:316-332) and its derivatives with respect to eps_r.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - every function returns 0 on success, <0 on error; cev_last_error() gives the
 *    message of the last failure on the calling thread.  No exception crosses the ABI.
 *  - all FIELD pointers are DEVICE pointers borrowed from the caller (torch tensors in
 *    the Python host layer); the library never allocates, frees or retains field
 *    memory.  The opaque plan owns only small tables (1-D PML profiles, index maps,
 *    source / probe point sets).
 *  - every compute call is asynchronous and ordered on the caller's cudaStream_t
 *    (passed as void*); a plan is bound to one device and is not thread-safe.
 *  - arrays are C-order (Nx, Ny, Nz), z contiguous, exactly like the reference's
 *    numpy arrays; vector arguments are indexed [0]=x, [1]=y, [2]=z.
 *  - dtype: 0 = fp32 storage, 1 = fp64 storage (the reference is fp64).
 */
#ifndef CEVICHE_B200_H
#define CEVICHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEV_ABI_VERSION 3
#define CEV_F32 0
#define CEV_F64 1

/* field codes used by probes / seeds: E=0..2, D=3..5, H=6..8 (x,y,z) */
#define CEV_FIELD_E 0
#define CEV_FIELD_D 3
#define CEV_FIELD_H 6

typedef struct cev_fdtd cev_fdtd;

/* State of one simulation (or one x-slab of it).  Replaces the attribute soup of
 * fdtd.initialize_fields() (fdtd.py:147-211): H*, D*, and the four PML integral families
 * ICE*, IH*, ICH*, ID*.  E is NOT state: E = inv_eps * D is formed on the fly by the H
 * half-step (bit-identical to fdtd.py:135-137) and only materialised on request.
 * The PML integrals are stored compactly, only where their coefficient is non-zero
 * (shapes from cev_fdtd_pml_shapes).  *_xhi / *_xlo are the x-halo planes (Ny*Nz
 * contiguous) a slab reads beyond its last / before its first x-plane; NULL means
 * "wrap inside this array" (single-slab np.roll semantics, derivatives.py:16-30). */
typedef struct cev_state {
    void*       H[3];
    void*       D[3];
    const void* inv_eps[3];      /* mE{x,y,z}1 = 1/eps_{xx,yy,zz}, fdtd.py:314-316 */
    void*       ICE[3];
    void*       IH[3];
    void*       ICH[3];
    void*       ID[3];
    const void* D_xhi[3];        /* D planes at local i = nx (from the right neighbour); [0] unused */
    const void* inv_eps_xhi[3];
    const void* H_xlo[3];        /* H planes at local i = -1 (from the left neighbour); [0] unused */
} cev_state;

/* Sparse (or dense) point sets for sources, probes and adjoint seeds.  All arrays are
 * device pointers.  A set with idx == NULL is dense: point q is cell (cell0 + q). */
typedef struct cev_points {
    int32_t        field;        /* CEV_FIELD_* + component (sources: D component 3..5) */
    int64_t        n;
    const int64_t* idx;          /* flat C-order cell index, or NULL */
    int64_t        cell0;
    const double*  weight;       /* n weights (profile / mask values) */
} cev_points;

/* Forward-mode tangent inputs: with eps_r perturbed along v, d(inv_eps) = -d(eps_yee)/eps_yee^2 and the
 * tangent state obeys the SAME step with J = 0, except E = inv_eps*dD + d_inv_eps*D_primal (product rule
 * on fdtd.py:135-137).  Both arrays are full grids, borrowed device pointers. */
typedef struct cev_tangent {
    const void* d_inv_eps[3];
    const void* D_primal[3];
} cev_tangent;

/* Cotangent state of the reverse sweep (same layouts as cev_state: compact PML arrays), one scratch
 * vector field, and the fp64 accumulators of dL/d(inv_eps).  G_mE entries may be NULL (not wanted).
 * g_box = {x0, x1, y0, y1, z0, z1}: G_mE is only accumulated for cells inside this box (the design region; eps_r
 * of cell (i,j,k) needs G_mE at (i,j,k), (i+1,j,k), (i,j+1,k), (i,j,k+1): utils.py:167-174); all zeros = whole grid. */
typedef struct cev_adjoint {
    void*   lH[3];
    void*   lD[3];
    void*   lICE[3];
    void*   lIH[3];
    void*   lICH[3];
    void*   lID[3];
    void*   gC2[3];
    double* G_mE[3];
    int64_t g_box[6];
    void*   gC[3];      /* second scratch vector field: optional (all NULL = the simple kernels); with it
                         * cev_fdtd_adjoint_run may use the tensor-map kernels (csrc/adjoint_v5.cuh) */
} cev_adjoint;

const char* cev_last_error(void);
int         cev_abi_version(void);

/* Plan = geometry + PML tables.  Replaces fdtd._set_time_step (fdtd.py:213-222, dt is
 * computed by the host layer and passed in), fdtd._compute_sigmas (fdtd.py:224-263, the
 * six 1-D profiles sH{x,y,z}, sD{x,y,z} on the HOST, lengths nx, Ny, Nz) and
 * fdtd._compute_update_parameters (fdtd.py:265-316): the 30 coefficient arrays are never
 * built; the kernels evaluate them from the profiles.  nx is the number of local x-planes. */
int cev_fdtd_create(cev_fdtd** plan, int device, int dtype, int arith_f64,
                    int64_t nx, int64_t Ny, int64_t Nz, double dL, double dt,
                    const double* sH[3], const double* sD[3]);
int cev_fdtd_destroy(cev_fdtd* plan);

/* Tuning / test knobs: "kernel_variant" 0 auto | 1 baseline (one thread per cell) | 2 marching | 3 TMA-staged |
 * 4 fused full-step kernel in cev_fdtd_run_fused | 5 hybrid there (lean fused kernel on the PML-free interior, the
 * half-step kernels on the PML shell); 6 tensor-map TMA kernels (cp.async.bulk.tensor box copies; "tma_rows" 4|8 tile rows, "tma_stages[_H|_D]" 3|4 ring
 * depth, "auto_tensor_map" 0|1: let variant 0 pick them on large 3-D grids); "fused_shape" 0 auto | lanes_z*100 + warps; "use_graph" / "jvp_streams" -1 auto | 0 | 1;
 * "xchunk" x-planes per CTA of the marching kernels (0 = auto); "lanes_z" 8|16|32 lanes of a warp along z;
 * "prefetch_planes" L2 prefetch distance; "split_launch" 0|1 separate launches for the PML-free interior and
 * the PML shell; "jvp_batch" 0|1 the tangent half-steps of a forward-mode sweep as one launch per half-step (1, default)
 * or one per tangent state; "jvp_fused" -1 auto | 0 | 1 both tangent half-steps of all states in ONE launch on 2-D TM grids
 * (tan2d_fused.cuh); "adjoint_variant" 0 auto | 1 one-thread-per-cell | 2 tensor-map kernels; "tma_stages_adjED" ring depth
 * of the adjoint E/D part; "halo_pause" 0|1 suspend the peer-store halo exchange (x-slab recomputation legs).
 * Results do not depend on them (bit-identical; the adjoint variants agree to 1e-12).
 * "active_components": 6-bit mask (bits 0-2: D/E x,y,z; bits 3-5: H x,y,z) of the components that may be non-zero;
 * the kernels neither read nor write the others (2-D TM / TE runs move 10-11 instead of 21 words per cell).  The
 * caller guarantees that the masked-out components are identically zero and are not driven. */
int cev_fdtd_set_option(cev_fdtd* plan, const char* name, int64_t value);

/* Logical shapes of the 12 compact PML integral arrays, order ICE[3], IH[3], ICH[3], ID[3]. */
int cev_fdtd_pml_shapes(const cev_fdtd* plan, int64_t shapes[12][3]);

/* H half-step, fdtd.py:80-97, on local x-planes [x0, x1).  H_out == NULL: in place. */
int cev_fdtd_step_H(cev_fdtd* plan, const cev_state* st, void* const H_out[3],
                    int64_t x0, int64_t x1, void* stream);
/* D/E half-step, fdtd.py:105-137.  D_out == NULL: in place.  E_out nullable (each entry).
 * J[c] nullable dense source added after the update, scaled by J_scale[c] (fdtd.py:125-127). */
int cev_fdtd_step_D(cev_fdtd* plan, const cev_state* st, void* const D_out[3], void* const E_out[3],
                    const void* const J[3], const double J_scale[3],
                    int64_t x0, int64_t x1, void* stream);
/* E = inv_eps * D (fdtd.py:135-137) into E_out (tan != NULL: the tangent dE). */
int cev_fdtd_compute_E(cev_fdtd* plan, const cev_state* st, const cev_tangent* tan, void* const E_out[3], void* stream);

/* Extended half-steps used by the slab driver and the derivative sweeps: optional tangent inputs,
 * probe sampling riding on the launch (probe_t >= 0: E/D probes of the PREVIOUS step on an H launch,
 * H probes of THIS step on a D launch, written to row probe_t of partials), and in-kernel injection of
 * the plan's sources lying in planes [x0, x1), scaled by waveform_row[n_sources] (device), after the D
 * update (fdtd.py:125-127). */
int cev_fdtd_step_H_ex(cev_fdtd* plan, const cev_state* st, const cev_tangent* tan, void* const H_out[3],
                       int64_t x0, int64_t x1, int64_t probe_t, double* partials, void* stream);
int cev_fdtd_step_D_ex(cev_fdtd* plan, const cev_state* st, void* const D_out[3], void* const E_out[3],
                       const void* const J[3], const double J_scale[3], const double* waveform_row,
                       int64_t x0, int64_t x1, int64_t probe_t, double* partials, void* stream);
/* Stand-alone probe sampling: which = 0 (E/D probes) or 1 (H probes) into row t of partials. */
int cev_fdtd_sample_probes(cev_fdtd* plan, const cev_state* st, const cev_tangent* tan, int which, int64_t t,
                           double* partials, void* stream);

/* Sources and probes of the caller loop (utils.py:316-332): J(t) = sum_s profile_s * waveform[t, s],
 * series[t, p] = sum(field_p * mask_p).  Point sets are copied into the plan. */
int cev_fdtd_set_sources(cev_fdtd* plan, int nsrc, const cev_points* src);
int cev_fdtd_set_probes(cev_fdtd* plan, int nprobe, const cev_points* probe, int64_t* n_slots);
/* slot -> probe map (host array of n_slots ints) so the caller can fold partial sums. */
int cev_fdtd_probe_slots(const cev_fdtd* plan, int32_t* slot_probe);
/* Probe series from the slot partial sums: series[t, p] = sum over probe p's slots of partials[t, slot], in slot
 * order (deterministic, independent of `rows`).  partials [rows, n_slots], series [rows, n_probes], device fp64. */
int cev_fdtd_fold_probes(cev_fdtd* plan, const double* partials, int64_t rows, double* series, void* stream);

/* Running-DFT monitors (frequency-domain fields of a region without storing its time series; replaces
 * storing field snapshots and transforming them, ceviche/utils.py:316-332 + 373-400).  Each monitor is a
 * point set (field code + idx; weights unused).  cev_fdtd_bind_monitors attaches, for the following
 * cev_fdtd_run calls, phasors = device double [nsteps, nfreq, 2] (re, im of exp(-i w_f t_n)) and
 * acc = device double [n_points, nfreq, 2], points in monitor order: after step n of a run,
 * acc[q, f] += field_q * phasors[n, f].  Bind (NULL, NULL) to detach. */
int cev_fdtd_set_monitors(cev_fdtd* plan, int nmon, const cev_points* mon, int nfreq, int64_t* n_points);
int cev_fdtd_bind_monitors(cev_fdtd* plan, const double* phasors, double* acc);

/* nsteps fused leap-frog steps with in-kernel source injection and probe sampling.
 * waveform: device double [nsteps, nsrc]; partials: device double [nsteps, n_slots]
 * (series[t, p] = sum of the slots of p, in slot order: deterministic). */
int cev_fdtd_run(cev_fdtd* plan, const cev_state* st, int64_t nsteps,
                 const double* waveform, double* partials, void* stream);

/* Same contract and bit-identical results, with ONE kernel per time step: the H and the D half-step of
 * fdtd.py:80-127 are fused, so the state is read once and written once per step (15 instead of 21 words per
 * cell).  That needs ping-pong buffers: `shadow` carries caller-owned scratch arrays of the same shapes in its
 * H, D, ICE and IH members (contents irrelevant, other members ignored).  The result always ends up in `st`.
 * Grids the fused kernel does not serve (contiguous extent not a multiple of the 16-byte vector, masked
 * components) silently take the two-kernel path of cev_fdtd_run. */
int cev_fdtd_run_fused(cev_fdtd* plan, const cev_state* st, const cev_state* shadow, int64_t nsteps,
                       const double* waveform, double* partials, void* stream);

/* Forward mode (replaces one traced re-run per direction, ceviche/jacobians.py:38-51): the primal and
 * B tangent states advance together; tangent_partials is [B, nsteps, n_slots]. */
int cev_fdtd_jvp_run(cev_fdtd* plan, const cev_state* st, int B, const cev_state* tangents, const cev_tangent* tans,
                     int64_t nsteps, const double* waveform, double* partials, double* tangent_partials, void* stream);

/* Reverse mode (replaces autograd's tape, ceviche/jacobians.py:29-35): one transposed time step.
 * fwd->D must hold the forward D after step n-1 and fwd->inv_eps the material; on return adj holds the
 * cotangents of the state after step n-1 and G_mE has gained step n's term.  cev_fdtd_adjoint_seed adds
 * the probe-series cotangents gbar_row[n_probes] (device) of one step; there fwd->D is D after THAT step. */
int cev_fdtd_adjoint_step(cev_fdtd* plan, const cev_state* fwd, const cev_adjoint* adj, void* stream);
/* One checkpoint segment of the reverse sweep of a run, on the device queue (cev_fdtd_adjoint_run).
 * st = the forward state at the START of the segment (H and the PML integrals are advanced in place by the
 * recomputation: pass scratch copies; st->D is ignored).  D_hist = nsteps + 1 caller-owned slots of three full-grid
 * arrays: slot 0 must hold D at the start of the segment, slot k receives D after k steps.  waveform / gbar: the
 * segment's nsteps rows ([nsteps, n_sources] / [nsteps, n_probes], device; gbar nullable).  On return adj holds the
 * cotangents of the state at the start of the segment and G_mE has gained the segment's terms. */
int cev_fdtd_adjoint_run(cev_fdtd* plan, const cev_state* st, int64_t nsteps, const double* waveform, const double* gbar,
                         void* const (*D_hist)[3], const cev_adjoint* adj, void* stream);
int cev_fdtd_adjoint_seed(cev_fdtd* plan, const cev_state* fwd, const cev_adjoint* adj, const double* gbar_row,
                          void* stream);
/* On grids whose kernels are shorter than a launch (option "use_graph": auto = up to 2^18 cells) cev_fdtd_adjoint_run
 * captures a segment -- recomputation and transposed steps -- into a CUDA graph the second time it is called with the
 * same arrays and length, and replays it from then on (the waveform / gbar rows travel through staging buffers).
 * Callers that want replays keep st, D_hist and adj on the same buffers from segment to segment.  Returns how many
 * segments of this plan were graph replays so far (diagnostics / tests). */
int64_t cev_fdtd_adjoint_graph_replays(const cev_fdtd* plan);
/* The transposed step in three parts with stored stencil inputs, for x-slabs (one process per GPU): the caller
 * exchanges one plane pair between the parts, as the forward half-steps do.
 *   part 0  cell-local D part: lD <- m1 lD + gID, gC <- m2 lD + gICH (lICH, lID advanced)
 *   part 1  H part: gH = lH + curl_E(gC); halo = the +x neighbour plane (local i = nx) of gC: entries of the two
 *           components differenced along x (y, z) = plane 0 of the right neighbour's gC; lH, gC2 written
 *   part 2  E part: lE = curl_H(gC2); halo = the -x neighbour plane (local i = -1) of gC2 = the left neighbour's last
 *           plane; lD += inv_eps lE, G_mE += lE D (fwd->D = D before the step) inside g_box
 * halo = NULL: periodic wrap inside the array (then parts 0, 1, 2 in a row equal cev_fdtd_adjoint_step).
 * Needs adj->gC. */
int cev_fdtd_adjoint_part(cev_fdtd* plan, int part, const cev_state* fwd, const cev_adjoint* adj,
                          const void* const halo[3], void* stream);
/* Reverse sweep WITHOUT recomputation, for gradients wanted inside a design box only.  The transposed step is linear
 * in the cotangents; the forward solution enters only dL/d(1/eps) += lE * D, so it suffices to keep D of the box:
 *   cev_fdtd_set_recorder(plan, box, buf, capacity): from now on every time step of cev_fdtd_run stores D (three
 *     components, the plan's storage type, C-order [3][bx][by][bz]) of box = {x0, x1, y0, y1, z0, z1} after the step
 *     into the next of `capacity` slots of the caller's device buffer `buf`; buf = NULL switches it off.
 *   cev_fdtd_adjoint_run_boxed(plan, st, nsteps, gbar, record, adj, stream): the nsteps transposed steps in reverse
 *     order with the probe-series seeds gbar [nsteps, n_probes] (device, nullable); only st->inv_eps is read.
 *     record = nsteps + 1 slots: slot 0 holds D of the box BEFORE the first step, slot k after step k;
 *     adj->g_box must be the recorded box and adj->gC must be set.
 * Served by the tensor-map kernels only (csrc/adjoint_v5.cuh): cev_fdtd_adjoint_boxed_supported(plan) tells. */
int cev_fdtd_adjoint_boxed_supported(const cev_fdtd* plan);
int cev_fdtd_set_recorder(cev_fdtd* plan, const int64_t box[6], void* buf, int64_t capacity);
int cev_fdtd_adjoint_run_boxed(cev_fdtd* plan, const cev_state* st, int64_t nsteps, const double* gbar,
                               const void* D_box_record, const cev_adjoint* adj, void* stream);

/* ---- x-slab decomposition over the GPUs of one box: halo planes through peer-mapped memory ----
 * The reference has no parallel path (single-thread numpy); this is the multi-GPU boundary SURVEY 8(b) sketched as
 * cev_fdtd_halo_ptrs.  The grid is cut into contiguous x-slabs, one plan per slab (nx = local planes).  A slab reads
 * two halo planes per half-step: D_y, D_z (and the static 1/eps_y, 1/eps_z) of its RIGHT neighbour's first plane in
 * the H half-step (curl_E, derivatives.py:16-22: forward differences) and H_y, H_z of its LEFT neighbour's last
 * plane in the D half-step (curl_H, derivatives.py:24-30: backward differences); np.roll wraps, so the slabs form
 * a ring.  Each slab owns an EXCHANGE BLOCK (cev_halo_layout) holding the planes it reads and two arrival counters;
 * the neighbours store into it directly from inside their half-step kernels (NVLink peer stores by the CTAs that
 * produce the boundary plane, followed by a system-scope release increment of the counter), and the CTAs that read
 * a halo plane first wait for its counter: no extra launch, no collective.  The boundary x-chunks run first in each
 * launch, so a halo has a whole half-step to arrive.  One process per GPU: the block is shared by CUDA IPC.
 *   cev_halo_alloc / cev_halo_free     this slab's block (zeroed) + its 64-byte IPC handle
 *   cev_halo_open / cev_halo_close     map a neighbour's block into this process (peer access is enabled lazily)
 *   cev_fdtd_halo_attach(plan, own, left, right)   from now on cev_fdtd_run / the in-place whole-slab half-steps of
 *        this plan use the blocks (cev_state's *_xhi / *_xlo members are ignored); (NULL, NULL, NULL) detaches.
 *        Every slab must issue the same sequence of half-steps.  A 2-slab ring passes the same block as left and right.
 *   cev_fdtd_halo_push_static   copy this slab's first plane of 1/eps_y, 1/eps_z to the left neighbour (once per eps_r)
 *   cev_fdtd_halo_reset         zero the field halo planes of this slab's own block (with initialize_fields)
 *   cev_fdtd_halo_error         1 if a halo wait timed out (a neighbour never arrived) since the block was allocated */
#define CEV_IPC_HANDLE_BYTES 64
typedef struct cev_halo_layout {
    size_t bytes;               /* size of an exchange block */
    size_t plane_bytes;
    size_t D_hi[2];             /* byte offsets: D_y, D_z of the plane at local i = nx */
    size_t inv_eps_hi[2];
    size_t H_lo[2];             /* H_y, H_z of the plane at local i = -1 */
    size_t flag_D, flag_H;      /* uint64 arrival counters of the D / H halo planes */
    size_t err;                 /* int32 */
} cev_halo_layout;
int cev_fdtd_halo_layout(const cev_fdtd* plan, cev_halo_layout* layout);
int cev_halo_alloc(int device, size_t bytes, void** block, unsigned char ipc_handle[CEV_IPC_HANDLE_BYTES]);
int cev_halo_open(int device, const unsigned char ipc_handle[CEV_IPC_HANDLE_BYTES], void** block);
int cev_halo_close(int device, void* block);
int cev_halo_free(int device, void* block);
int cev_fdtd_halo_attach(cev_fdtd* plan, void* own_block, void* left_block, void* right_block);
int cev_fdtd_halo_push_static(cev_fdtd* plan, const cev_state* st, void* stream);
int cev_fdtd_halo_reset(cev_fdtd* plan, void* stream);
int cev_fdtd_halo_error(cev_fdtd* plan, int* err);

#ifdef __cplusplus
}
#endif
#endif /* CEVICHE_B200_H */
