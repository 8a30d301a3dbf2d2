// Host interface of the tensor-map TMA half-step kernels (step_v5.cuh), compiled in their own
// translation unit (step_v5.cu).  cev_fdtd.cu fills StepArgs as for every other variant and calls these.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace cev {

// Cache of CUtensorMap descriptors keyed by (array pointer, planes, box kind): encoding one costs about a
// microsecond on the host, a simulation reuses the same ~30 arrays for its whole life.
struct V5MapCache;
V5MapCache* v5_cache_create();
void v5_cache_destroy(V5MapCache* c);
const char* v5_last_error();

// rows per CTA: 4 or 8; stages of the shared-memory ring: 3 or 4
bool v5_supported_shape(int rows, int stages);

// Does this launch fit the kernels?  (3-D plane of >= 1 full tile row, Ny a multiple of `rows`, all six components
// live, no tangent inputs, no dense J / E output, 16-byte aligned arrays.)
template <typename T, typename AT>
bool v5_eligible(const StepArgs<T, AT>& a, int rows);

// Tiling of planes [x0, x1): one box = the whole y-z plane; fills a.x0, a.x1, a.xchunk, a.ntz, a.nty, a.n_tiles, a.box[0].
template <typename T, typename AT>
void v5_set_tiles(StepArgs<T, AT>& a, int64_t x0, int64_t x1, int rows, int xchunk);

// n_aux probe CTAs are appended to the grid (a.aux_slot0 / a.t_probe / a.partials already set).
template <typename T, typename AT>
int v5_launch_H(V5MapCache* c, const StepArgs<T, AT>& a, int rows, int stages, int n_aux, cudaStream_t s);
template <typename T, typename AT>
int v5_launch_D(V5MapCache* c, const StepArgs<T, AT>& a, int rows, int stages, int n_aux, cudaStream_t s);

}  // namespace cev
