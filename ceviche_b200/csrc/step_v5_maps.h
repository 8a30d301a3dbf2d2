// Internal to the tensor-map translation units (step_v5.cu, adjoint_v5.cu): the descriptor cache's encoder.
#pragma once
#include <cuda.h>

#include "step_v5.h"

namespace cev {

enum BoxKind { BOX_MAIN = 0, BOX_ROW = 1, BOX_PLN = 2 };   // (BZ + 2V, rows) | (BZ + 2V, 1) | (BZ, rows) cells of one plane

// descriptor of `nx` planes of (Ny, Nz) cells of `esize` bytes at `ptr`; cached per (ptr, nx, kind, rows, esize)
int v5_get_map(V5MapCache* c, const void* ptr, int64_t nx, int Ny, int Nz, int esize, int kind, int rows, CUtensorMap* out);
// dynamic shared memory opt-in, once per device and kernel instantiation (`done`: 64 ints, static of the caller)
int v5_set_smem_attr(const void* kernel, size_t bytes, int* done);
int v5_fail_msg(const char* what, const char* detail);

}  // namespace cev
