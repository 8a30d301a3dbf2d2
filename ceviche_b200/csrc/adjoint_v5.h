// Host interface of the tensor-map TMA kernels of the reverse sweep (adjoint_v5.cuh), compiled in their own
// translation unit (adjoint_v5.cu).  The caller fills StepArgs as documented in adjoint_v5.cuh (AdjV5Extra) and
// tiles them with v5_set_tiles.
#pragma once
#include "step_v5.h"

namespace cev {

// H part: a.Din = gC, a.Hin = a.Hout = lH, a.Eout = gC2, a.ICE / a.IH = lICE / lIH
template <typename T, typename AT>
int v5_launch_adj_H(V5MapCache* c, const StepArgs<T, AT>& a, int rows, int stages, cudaStream_t s);
// E + D parts: a.Hin = gC2, a.Din = a.Dout = lDp, a.mE = 1/eps, a.Eout = gC, a.ICH / a.ID = lICH / lID;
// Dprev = forward D after step k-1, G = fp64 accumulators (entries nullable), gb = design box (internal axes),
// eager = 0 for the last launch of a segment (leaves the true cotangent of D); boxed = 1: Dprev[c] holds the design
// box only (C-order, the extents of gb): the D-box record of the forward run (cev_fdtd_set_recorder)
template <typename T, typename AT>
int v5_launch_adj_ED(V5MapCache* c, const StepArgs<T, AT>& a, const void* const Dprev[3], double* const G[3], const int gb[6],
                     int eager, int boxed, int rows, int stages, cudaStream_t s);

}  // namespace cev
