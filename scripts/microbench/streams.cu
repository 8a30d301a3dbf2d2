// Microbenchmarks that bound what the half-step kernels can reach on this B200:
//  A. copy (1 read + 1 write stream), grid-stride 16B vectors
//  B. 9 read + 3 write streams, elementwise, grid-stride (the H half-step's stream count, ideal pattern)
//  C. same 12 streams, but with the marching access pattern (CTA = 4 rows x 512 B, walks x-planes)
//  D. C + the extra j+1 row loads
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o streams streams.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
typedef double2 V;

__global__ void k_copy(const V* __restrict__ a, V* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
struct P12 { const V* r[9]; V* w[3]; };
__global__ void k_12(P12 p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        V x[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) x[c] = p.r[c][i];
#pragma unroll
        for (int c = 0; c < 3; ++c) { V o; o.x = x[c].x * x[c + 3].x + x[c + 6].x; o.y = x[c].y * x[c + 3].y + x[c + 6].y; p.w[c][i] = o; }
    }
}
// marching: CTA (32,BY) owns rows j0..j0+BY-1, z-segment of 32 vectors; loops over planes [xs,xe)
template <int BY, bool JP>
__global__ void k_march(P12 p, int Nx, int Ny, int Nzv, int xchunk) {
    const int ntz = Nzv / 32, nty = Ny / BY;
    const int bid = blockIdx.x;
    const int tz = bid % ntz, ty = (bid / ntz) % nty, xc = bid / (ntz * nty);
    const int j = ty * BY + threadIdx.y, k = tz * 32 + threadIdx.x;
    const int jp = (j + 1 == Ny) ? 0 : j + 1;
    const size_t plane = (size_t)Ny * Nzv;
    const int xs = xc * xchunk, xe = min(xs + xchunk, Nx);
    for (int i = xs; i < xe; ++i) {
        const size_t o = i * plane + (size_t)j * Nzv + k;
        V x[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) x[c] = p.r[c][o];
        V y[4];
        if (JP) {
            const size_t oj = i * plane + (size_t)jp * Nzv + k;
            y[0] = p.r[0][oj]; y[1] = p.r[3][oj]; y[2] = p.r[2][oj]; y[3] = p.r[5][oj];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            V o2; o2.x = x[c].x * x[c + 3].x + x[c + 6].x; o2.y = x[c].y * x[c + 3].y + x[c + 6].y;
            if (JP) { o2.x += y[c].x * y[3].x; o2.y += y[c].y * y[3].y; }
            p.w[c][o] = o2;
        }
    }
}
template <typename F> float timeit(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 256;
    const size_t n = (size_t)N * N * N / 2;   // double2 vectors
    std::vector<V*> buf(12);
    for (auto& b : buf) { cudaMalloc(&b, n * sizeof(V)); cudaMemset(b, 0, n * sizeof(V)); }
    P12 p; for (int c = 0; c < 9; ++c) p.r[c] = buf[c]; for (int c = 0; c < 3; ++c) p.w[c] = buf[c + 6];   // in place like H
    const double gb12 = 12.0 * n * sizeof(V) / 1e9, gb2 = 2.0 * n * sizeof(V) / 1e9;
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        float t = timeit([&] { k_copy<<<g, 256>>>(buf[0], buf[1], n); });
        printf("A copy      grid %5d x256: %.4f ms %7.0f GB/s\n", g, t, gb2 / t * 1e3);
    }
    for (int bs : {128, 256}) for (int g : {148 * 4, 148 * 8, 148 * 16}) {
        float t = timeit([&] { k_12<<<g, bs>>>(p, n); });
        printf("B 12-stream grid %5d x%d: %.4f ms %7.0f GB/s\n", g, bs, t, gb12 / t * 1e3);
    }
    const int Nzv = N / 2;
    for (int xchunk : {8, 16, 64, N}) {
        const int nchunks = (N + xchunk - 1) / xchunk;
        { const int grid = (Nzv / 32) * (N / 4) * nchunks;
          float t = timeit([&] { k_march<4, false><<<grid, dim3(32, 4)>>>(p, N, N, Nzv, xchunk); });
          printf("C march BY=4 xchunk %4d grid %6d: %.4f ms %7.0f GB/s\n", xchunk, grid, t, gb12 / t * 1e3);
          t = timeit([&] { k_march<4, true><<<grid, dim3(32, 4)>>>(p, N, N, Nzv, xchunk); });
          printf("D march+jp BY=4 xchunk %4d          : %.4f ms %7.0f GB/s (12-stream bytes)\n", xchunk, t, gb12 / t * 1e3); }
        { const int grid = (Nzv / 32) * (N / 8) * nchunks;
          float t = timeit([&] { k_march<8, false><<<grid, dim3(32, 8)>>>(p, N, N, Nzv, xchunk); });
          printf("C march BY=8 xchunk %4d grid %6d: %.4f ms %7.0f GB/s\n", xchunk, grid, t, gb12 / t * 1e3);
          t = timeit([&] { k_march<8, true><<<grid, dim3(32, 8)>>>(p, N, N, Nzv, xchunk); });
          printf("D march+jp BY=8 xchunk %4d          : %.4f ms %7.0f GB/s (12-stream bytes)\n", xchunk, t, gb12 / t * 1e3); }
    }
    return 0;
}
