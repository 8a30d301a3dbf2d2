// Host side of the tensor-map TMA kernels: descriptor cache + launches.
#include "step_v5.h"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <string>
#include <unordered_map>

#include "step_v5.cuh"
#include "step_v5_maps.h"

namespace cev {

namespace {

thread_local std::string g_v5_err;

int v5_fail(const char* what, const char* detail = "") {
    g_v5_err = std::string(what) + detail;
    return -1;
}

// cuTensorMapEncodeTiled lives in the driver library; resolve it through the runtime so that the
// shared library does not link against libcuda.
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}


struct Key {
    const void* ptr;
    int64_t nx;
    int tag;      // kind | rows << 4 | esize << 8
    bool operator==(const Key& o) const { return ptr == o.ptr && nx == o.nx && tag == o.tag; }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        return std::hash<const void*>()(k.ptr) ^ (std::hash<int64_t>()(k.nx) * 1000003u) ^ ((size_t)k.tag * 7919u);
    }
};

}  // namespace

struct V5MapCache {
    std::unordered_map<Key, CUtensorMap, KeyHash> maps;
    int Ny = 0, Nz = 0;
};

V5MapCache* v5_cache_create() { return new V5MapCache(); }
void v5_cache_destroy(V5MapCache* c) { delete c; }
const char* v5_last_error() { return g_v5_err.c_str(); }
bool v5_supported_shape(int rows, int stages) { return (rows == 4 || rows == 8) && (stages == 3 || stages == 4); }

// descriptor of `nx` planes of (Ny, Nz) cells of `esize` bytes at `ptr`, box kind / rows as given (also used by
// adjoint_v5.cu: declared in step_v5_maps.h)
int v5_get_map(V5MapCache* c, const void* ptr, int64_t nx, int Ny, int Nz, int esize, int kind, int rows, CUtensorMap* out) {
    if (c->Ny != Ny || c->Nz != Nz) {       // (a cache belongs to one plan: one plane shape)
        c->maps.clear();
        c->Ny = Ny;
        c->Nz = Nz;
    }
    const Key key{ptr, nx, kind | (rows << 4) | (esize << 8)};
    auto it = c->maps.find(key);
    if (it != c->maps.end()) {
        *out = it->second;
        return 0;
    }
    auto enc = encode_fn();
    if (!enc) return v5_fail("cuTensorMapEncodeTiled is not available from this driver");
    const int V = 16 / esize;
    const cuuint64_t gdim[3] = {(cuuint64_t)Nz, (cuuint64_t)Ny, (cuuint64_t)nx};
    const cuuint64_t gstr[2] = {(cuuint64_t)Nz * esize, (cuuint64_t)Ny * Nz * esize};
    const cuuint32_t bz = kind == BOX_PLN ? 32 * V : 32 * V + 2 * V;
    const cuuint32_t by = kind == BOX_ROW ? 1 : rows;
    const cuuint32_t box[3] = {bz, by, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                           const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[96];
        snprintf(buf, sizeof buf, " (CUresult %d, Nz=%d Ny=%d nx=%lld box=%ux%u)", (int)r, Nz, Ny, (long long)nx, bz, by);
        return v5_fail("cuTensorMapEncodeTiled failed", buf);
    }
    if (c->maps.size() >= 8192) c->maps.clear();
    c->maps.emplace(key, m);
    *out = m;
    return 0;
}

int v5_set_smem_attr(const void* kernel, size_t bytes, int* done) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && done[dev]) return 0;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return v5_fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ", cudaGetErrorString(e));
    if (dev < 64) done[dev] = 1;
    return 0;
}
int v5_fail_msg(const char* what, const char* detail) { return v5_fail(what, detail); }

namespace {

// dynamic shared memory opt-in, once per device and kernel instantiation (`done` is a static of the CALLER, which is
// a distinct function per instantiation; the kernel pointer TYPE is shared by all of them)
template <typename K>
int set_smem(K kernel, size_t bytes, int* done) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && done[dev]) return 0;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return v5_fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ", cudaGetErrorString(e));
    if (dev < 64) done[dev] = 1;
    return 0;
}

template <typename T, typename AT, int BY, int NS>
int launch_H_shape(const StepArgs<T, AT>& a, const V5MapsH& m, int n_aux, cudaStream_t s) {
    constexpr int V = 16 / (int)sizeof(T);
    const size_t smem = V5Layout<T, V, BY>::h_bytes(NS);
    static int done[64] = {0};
    if (set_smem(k_step_H_v5<T, AT, V, BY, NS>, smem, done)) return -1;
    k_step_H_v5<T, AT, V, BY, NS><<<a.n_tiles + n_aux, dim3(32, BY), smem, s>>>(a, m);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : v5_fail("k_step_H_v5 launch failed: ", cudaGetErrorString(e));
}

template <typename T, typename AT, int BY, int NS>
int launch_D_shape(const StepArgs<T, AT>& a, const V5MapsD& m, int n_aux, cudaStream_t s) {
    constexpr int V = 16 / (int)sizeof(T);
    const size_t smem = V5Layout<T, V, BY>::d_bytes(NS);
    static int done[64] = {0};
    if (set_smem(k_step_D_v5<T, AT, V, BY, NS>, smem, done)) return -1;
    k_step_D_v5<T, AT, V, BY, NS><<<a.n_tiles + n_aux, dim3(32, BY), smem, s>>>(a, m);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : v5_fail("k_step_D_v5 launch failed: ", cudaGetErrorString(e));
}

bool aligned16(const void* q) { return q == nullptr || ((uintptr_t)q % 16) == 0; }

}  // namespace

template <typename T, typename AT>
bool v5_eligible(const StepArgs<T, AT>& a, int rows) {
    constexpr int V = 16 / (int)sizeof(T);
    if (a.on != 63u || a.dmE[0] || a.dmE[1] || a.dmE[2]) return false;
    if (a.Ny < rows || a.Ny % rows != 0) return false;
    if (a.Nz % V != 0 || a.Nz < 32 * V) return false;
    for (int c = 0; c < 3; ++c) {
        if (a.J[c] || a.Eout[c]) return false;
        if (!aligned16(a.Hin[c]) || !aligned16(a.Hout[c]) || !aligned16(a.Din[c]) || !aligned16(a.Dout[c]) ||
            !aligned16(a.mE[c]) || !aligned16(a.Dhi[c]) || !aligned16(a.mEhi[c]) || !aligned16(a.Hlo[c]) ||
            !aligned16(a.ICE[c]) || !aligned16(a.IH[c]) || !aligned16(a.ICH[c]) || !aligned16(a.ID[c]))
            return false;
    }
    return true;
}

template <typename T, typename AT>
void v5_set_tiles(StepArgs<T, AT>& a, int64_t x0, int64_t x1, int rows, int xchunk) {
    constexpr int V = 16 / (int)sizeof(T);
    a.x0 = (int)x0;
    a.x1 = (int)x1;
    a.wz = 0;
    a.xorder = 1;
    a.ntz = (a.Nz + 32 * V - 1) / (32 * V);
    a.nty = a.Ny / rows;
    if (xchunk > 0) {
        a.xchunk = xchunk < V5_MAXCH ? xchunk : V5_MAXCH;
    } else {       // 16 planes amortise the pipeline fill; shorter chunks where that leaves fewer than ~9 waves of CTAs
        a.xchunk = 16;
        while (a.xchunk > 4 && (int64_t)a.ntz * a.nty * ((x1 - x0 + a.xchunk - 1) / a.xchunk) < 4096) a.xchunk /= 2;
    }
    const int nchunks = (int)((x1 - x0 + a.xchunk - 1) / a.xchunk);
    a.n_tiles = a.ntz * a.nty * nchunks;
    a.n_boxes = 1;
    Box& B = a.box[0];
    B.x0 = (int)x0; B.x1 = (int)x1; B.y0 = 0; B.y1 = a.Ny; B.z0 = 0; B.z1 = a.Nz;
    B.cta0 = 0; B.ntz = a.ntz; B.nty = a.nty;
}

template <typename T, typename AT>
int v5_launch_H(V5MapCache* c, const StepArgs<T, AT>& a, int rows, int stages, int n_aux, cudaStream_t s) {
    if (a.n_tiles + n_aux == 0) return 0;
    constexpr int es = (int)sizeof(T);
    V5MapsH m;
    for (int q = 0; q < 3; ++q) {
        if (v5_get_map(c, a.Din[q], a.Nx, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.D[q])) return -1;
        if (v5_get_map(c, a.mE[q], a.Nx, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.M[q])) return -1;
        if (v5_get_map(c, a.Hin[q], a.Nx, a.Ny, a.Nz, es, BOX_PLN, rows, &m.H[q])) return -1;
    }
    for (int r = 0; r < 2; ++r) {
        const int cr = r == 0 ? 0 : 2, ch = r == 0 ? 1 : 2;
        if (v5_get_map(c, a.Din[cr], a.Nx, a.Ny, a.Nz, es, BOX_ROW, rows, &m.Drow[r])) return -1;
        if (v5_get_map(c, a.mE[cr], a.Nx, a.Ny, a.Nz, es, BOX_ROW, rows, &m.Mrow[r])) return -1;
        // the plane standing for i = Nx: ONE plane at a.Dhi (plane 0 of the array itself on a periodic grid)
        if (v5_get_map(c, a.Dhi[ch], 1, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.Dhi[r])) return -1;
        if (v5_get_map(c, a.mEhi[ch], 1, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.Mhi[r])) return -1;
    }
    m.x_hi = 0;
#define CEV_V5_H(BY, NS) return launch_H_shape<T, AT, BY, NS>(a, m, n_aux, s)
    if (rows == 4 && stages == 3) CEV_V5_H(4, 3);
    if (rows == 4 && stages == 4) CEV_V5_H(4, 4);
    if (rows == 8 && stages == 3) CEV_V5_H(8, 3);
    if (rows == 8 && stages == 4) CEV_V5_H(8, 4);
#undef CEV_V5_H
    return v5_fail("unsupported tile shape of the tensor-map kernels");
}

template <typename T, typename AT>
int v5_launch_D(V5MapCache* c, const StepArgs<T, AT>& a, int rows, int stages, int n_aux, cudaStream_t s) {
    if (a.n_tiles + n_aux == 0) return 0;
    constexpr int es = (int)sizeof(T);
    V5MapsD m;
    for (int q = 0; q < 3; ++q) {
        if (v5_get_map(c, a.Hin[q], a.Nx, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.H[q])) return -1;
        if (v5_get_map(c, a.Din[q], a.Nx, a.Ny, a.Nz, es, BOX_PLN, rows, &m.D[q])) return -1;
    }
    for (int r = 0; r < 2; ++r) {
        const int cr = r == 0 ? 0 : 2, cl = r == 0 ? 1 : 2;
        if (v5_get_map(c, a.Hin[cr], a.Nx, a.Ny, a.Nz, es, BOX_ROW, rows, &m.Hrow[r])) return -1;
        // the plane standing for i = -1: ONE plane at a.Hlo (the array's last plane on a periodic grid)
        if (v5_get_map(c, a.Hlo[cl], 1, a.Ny, a.Nz, es, BOX_MAIN, rows, &m.Hlo[r])) return -1;
    }
    m.x_lo = 0;
#define CEV_V5_D(BY, NS) return launch_D_shape<T, AT, BY, NS>(a, m, n_aux, s)
    if (rows == 4 && stages == 3) CEV_V5_D(4, 3);
    if (rows == 4 && stages == 4) CEV_V5_D(4, 4);
    if (rows == 8 && stages == 3) CEV_V5_D(8, 3);
    if (rows == 8 && stages == 4) CEV_V5_D(8, 4);
#undef CEV_V5_D
    return v5_fail("unsupported tile shape of the tensor-map kernels");
}

#define CEV_V5_INSTANTIATE(T, AT)                                                                             \
    template bool v5_eligible<T, AT>(const StepArgs<T, AT>&, int);                                            \
    template void v5_set_tiles<T, AT>(StepArgs<T, AT>&, int64_t, int64_t, int, int);                          \
    template int v5_launch_H<T, AT>(V5MapCache*, const StepArgs<T, AT>&, int, int, int, cudaStream_t);       \
    template int v5_launch_D<T, AT>(V5MapCache*, const StepArgs<T, AT>&, int, int, int, cudaStream_t);
CEV_V5_INSTANTIATE(double, double)
CEV_V5_INSTANTIATE(float, double)
CEV_V5_INSTANTIATE(float, float)

}  // namespace cev
