/* C restatement of one FDTD time step of the reference (TEST INFRASTRUCTURE -- never linked into the product).
 *
 * Follows ceviche/fdtd.py:74-144 with curl_E / curl_H of ceviche/derivatives.py:16-30 operation for operation, in
 * the evaluation order numpy uses for the reference's expressions, so that results are BIT-IDENTICAL to the reference
 * and to oracle/fdtd_numpy.py (tests/test_oracle_c.py); compile with -ffp-contract=off (no FMA contraction).
 * The 24 coefficient arrays of fdtd.py:272-311 and the three 1/eps arrays of :314-316 are inputs (built by the numpy
 * oracle from the reference formulas), exactly as the reference streams them every step.
 * Unlike the reference (single-threaded numpy, ~102 array passes per step) this is ONE fused pass per half-step,
 * parallel over x-planes with OpenMP: it is the "best plain CPU" baseline next to the faithful numpy one.
 *
 * Arrays are C-order (Nx, Ny, Nz) doubles; v[c] = component c (x, y, z).  */
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* n > 0: use n OpenMP threads from now on (launchers such as torchrun export OMP_NUM_THREADS=1).  Returns the thread
 * count parallel regions will use. */
int oracle_omp_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define IDX(i, j, k) (((size_t)(i) * Ny + (j)) * Nz + (k))

/* component c of the forward-difference curl at (i,j,k): (F_v[u+1] - F_v)/dL - (F_u[v+1] - F_u)/dL, (u,v) = (c+1,c+2) */
static inline double curl_fwd(int c, double* const F[3], int Nx, int Ny, int Nz, int i, int j, int k, double dL) {
    const int ip = (i + 1 == Nx) ? 0 : i + 1, jp = (j + 1 == Ny) ? 0 : j + 1, kp = (k + 1 == Nz) ? 0 : k + 1;
    const size_t o = IDX(i, j, k);
    const size_t nb[3] = {IDX(ip, j, k), IDX(i, jp, k), IDX(i, j, kp)};
    const int u = (c + 1) % 3, v = (c + 2) % 3;
    return (F[v][nb[u]] - F[v][o]) / dL - (F[u][nb[v]] - F[u][o]) / dL;
}

static inline double curl_bwd(int c, double* const F[3], int Nx, int Ny, int Nz, int i, int j, int k, double dL) {
    const int im = (i == 0) ? Nx - 1 : i - 1, jm = (j == 0) ? Ny - 1 : j - 1, km = (k == 0) ? Nz - 1 : k - 1;
    const size_t o = IDX(i, j, k);
    const size_t nb[3] = {IDX(im, j, k), IDX(i, jm, k), IDX(i, j, km)};
    const int u = (c + 1) % 3, v = (c + 2) % 3;
    return (F[v][o] - F[v][nb[u]]) / dL - (F[u][o] - F[u][nb[v]]) / dL;
}

/* mH / mD: [component][4] coefficient arrays (m1..m4); J entries may be NULL. */
void oracle_fdtd_step(int Nx, int Ny, int Nz, double dL, double* const H[3], double* const D[3], double* const E[3],
                      double* const ICE[3], double* const IH[3], double* const ICH[3], double* const ID[3],
                      const double* const mH[12], const double* const mD[12], const double* const mE[3],
                      const double* const J[3]) {
    /* fdtd.py:80-97: CE = curl_E(E); ICE += CE; IH += H; H = m1*H + m2*CE + m3*ICE + m4*IH */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j)
            for (int k = 0; k < Nz; ++k) {
                const size_t o = IDX(i, j, k);
                for (int c = 0; c < 3; ++c) {
                    const double ce = curl_fwd(c, E, Nx, Ny, Nz, i, j, k, dL);
                    const double ice = ICE[c][o] + ce;
                    const double ih = IH[c][o] + H[c][o];
                    ICE[c][o] = ice;
                    IH[c][o] = ih;
                    H[c][o] = ((mH[4 * c][o] * H[c][o] + mH[4 * c + 1][o] * ce) + mH[4 * c + 2][o] * ice) + mH[4 * c + 3][o] * ih;
                }
            }
    /* fdtd.py:105-137: CH = curl_H(H); ICH += CH; ID += D; D = m1*D + m2*CH + m3*ICH + m4*ID; D += J; E = mE*D */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < Nx; ++i)
        for (int j = 0; j < Ny; ++j)
            for (int k = 0; k < Nz; ++k) {
                const size_t o = IDX(i, j, k);
                for (int c = 0; c < 3; ++c) {
                    const double ch = curl_bwd(c, H, Nx, Ny, Nz, i, j, k, dL);
                    const double ich = ICH[c][o] + ch;
                    const double id = ID[c][o] + D[c][o];
                    ICH[c][o] = ich;
                    ID[c][o] = id;
                    double d = ((mD[4 * c][o] * D[c][o] + mD[4 * c + 1][o] * ch) + mD[4 * c + 2][o] * ich) + mD[4 * c + 3][o] * id;
                    if (J[c]) d += J[c][o];
                    D[c][o] = d;
                    E[c][o] = mE[c][o] * d;
                }
            }
}

int oracle_fdtd_abi(void) { return 1; }
