// Host side of the tensor-map TMA kernels of the reverse sweep (adjoint_v5.cuh): descriptors + launches.
#include "adjoint_v5.h"

#include "adjoint_v5.cuh"
#include "step_v5_maps.h"

namespace cev {

namespace {

template <typename T, typename AT, int BY, int NS>
int launch_adj_H_shape(const StepArgs<T, AT>& a, const V5MapsAdjH& m, cudaStream_t s) {
    constexpr int V = 16 / (int)sizeof(T);
    const size_t smem = AdjV5Layout<T, V, BY>::h_bytes(NS);
    static int done[64] = {0};
    if (v5_set_smem_attr((const void*)k_adj_H_v5<T, AT, V, BY, NS>, smem, done)) return -1;
    k_adj_H_v5<T, AT, V, BY, NS><<<a.n_tiles, dim3(32, BY), smem, s>>>(a, m);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : v5_fail_msg("k_adj_H_v5 launch failed: ", cudaGetErrorString(e));
}

template <typename T, typename AT, int BY, int NS>
int launch_adj_ED_shape(const StepArgs<T, AT>& a, const V5MapsAdjED& m, const AdjV5Extra<T>& x, cudaStream_t s) {
    constexpr int V = 16 / (int)sizeof(T);
    const size_t smem = AdjV5Layout<T, V, BY>::ed_bytes(NS);
    static int done[64] = {0};
    if (v5_set_smem_attr((const void*)k_adj_ED_v5<T, AT, V, BY, NS>, smem, done)) return -1;
    k_adj_ED_v5<T, AT, V, BY, NS><<<a.n_tiles, dim3(32, BY), smem, s>>>(a, m, x);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : v5_fail_msg("k_adj_ED_v5 launch failed: ", cudaGetErrorString(e));
}

}  // namespace

template <typename T, typename AT>
int v5_launch_adj_H(V5MapCache* c, const StepArgs<T, AT>& a, int rows, int stages, cudaStream_t s) {
    constexpr int es = (int)sizeof(T);
    V5MapsAdjH m;
    for (int q = 0; q < 3; ++q) {
        if (v5_get_map(c, a.Din[q], a.Nx, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.C[q])) return -1;
        if (v5_get_map(c, a.Hin[q], a.Nx, a.Ny, a.Nz, es, BOX_PLN, rows, &m.L[q])) return -1;
    }
    for (int r = 0; r < 2; ++r)
        if (v5_get_map(c, a.Din[r == 0 ? 0 : 2], a.Nx, a.Ny, a.Nz, es, BOX_ROW, rows, &m.Crow[r])) return -1;
#define CEV_ADJ_H(BY, NS) return launch_adj_H_shape<T, AT, BY, NS>(a, m, s)
    if (rows == 4 && stages == 3) CEV_ADJ_H(4, 3);
    if (rows == 4 && stages == 4) CEV_ADJ_H(4, 4);
    if (rows == 8 && stages == 3) CEV_ADJ_H(8, 3);
    if (rows == 8 && stages == 4) CEV_ADJ_H(8, 4);
#undef CEV_ADJ_H
    return v5_fail_msg("unsupported tile shape of the tensor-map adjoint kernels", "");
}

template <typename T, typename AT>
int v5_launch_adj_ED(V5MapCache* c, const StepArgs<T, AT>& a, const void* const Dprev[3], double* const G[3], const int gb[6],
                     int eager, int boxed, int rows, int stages, cudaStream_t s) {
    constexpr int es = (int)sizeof(T);
    V5MapsAdjED m;
    AdjV5Extra<T> x;
    for (int q = 0; q < 3; ++q) {
        if (v5_get_map(c, a.Hin[q], a.Nx, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.C[q])) return -1;
        if (v5_get_map(c, a.Din[q], a.Nx, a.Ny, a.Nz, es, BOX_PLN, rows, &m.L[q])) return -1;
        if (v5_get_map(c, a.mE[q], a.Nx, a.Ny, a.Nz, es, BOX_PLN, rows, &m.M[q])) return -1;
        x.Dprev[q] = (const T*)Dprev[q];
        x.G[q] = G[q];
    }
    for (int r = 0; r < 2; ++r)
        if (v5_get_map(c, a.Hin[r == 0 ? 0 : 2], a.Nx, a.Ny, a.Nz, es, BOX_ROW, rows, &m.Crow[r])) return -1;
    for (int q = 0; q < 6; ++q) x.gb[q] = gb[q];
    x.eager = eager;
    x.boxed = boxed;
#define CEV_ADJ_ED(BY, NS) return launch_adj_ED_shape<T, AT, BY, NS>(a, m, x, s)
    if (rows == 4 && stages == 3) CEV_ADJ_ED(4, 3);
    if (rows == 4 && stages == 4) CEV_ADJ_ED(4, 4);
    if (rows == 8 && stages == 3) CEV_ADJ_ED(8, 3);
    if (rows == 8 && stages == 4) CEV_ADJ_ED(8, 4);
#undef CEV_ADJ_ED
    return v5_fail_msg("unsupported tile shape of the tensor-map adjoint kernels", "");
}

#define CEV_ADJ_V5_INSTANTIATE(T, AT)                                                                              \
    template int v5_launch_adj_H<T, AT>(V5MapCache*, const StepArgs<T, AT>&, int, int, cudaStream_t);              \
    template int v5_launch_adj_ED<T, AT>(V5MapCache*, const StepArgs<T, AT>&, const void* const[3], double* const[3], \
                                         const int[6], int, int, int, int, cudaStream_t);
CEV_ADJ_V5_INSTANTIATE(double, double)
CEV_ADJ_V5_INSTANTIATE(float, double)
CEV_ADJ_V5_INSTANTIATE(float, float)

}  // namespace cev
