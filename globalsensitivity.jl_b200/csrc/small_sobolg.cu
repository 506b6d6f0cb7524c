// small_sobolg.cu -- fused design-swap + Sobol-G + estimator kernel with second order, d <= 12 (BASELINE config 3).
#include "small_launch.cuh"

namespace gsa {

template <int D, bool SECOND>
static int go(gsa_handle* h, int est, const double* A, const double* B, const double2* coef, const TileGeom& g, int t0, int t1, int nstat,
              double* tile_rec, int* occ) {
    SmallArgs<SobolGSource<D, SECOND>> a;
    a.src.A = A; a.src.B = B; a.src.coef = coef; a.src.buf_col0 = g.buf_col0;
    a.tile_rec = tile_rec; a.g = g; a.nout = 1; a.nstat = nstat; a.est = est; a.t_begin = t0; a.t_end = t1;
    if (est == GSA_EI_JANSEN1999) return launch_small<SobolGSource<D, SECOND>, D, SECOND, true>(h, a, occ);
    return launch_small<SobolGSource<D, SECOND>, D, SECOND, false>(h, a, occ);
}

int launch_small_sobolg(gsa_handle* h, int d, bool second, int est, const double* A, const double* B, const double2* coef,
                        const TileGeom& g, int t0, int t1, int nstat, double* tile_rec, int* occ) {
    if (second && d > SMALL_DMAX2) return fail(GSA_ERR_UNSUPPORTED, "small second-order kernel supports d <= %d", SMALL_DMAX2);
    switch (d) {
#define CASE(DD) \
    case DD: return second ? go<(DD <= SMALL_DMAX2 ? DD : 1), true>(h, est, A, B, coef, g, t0, t1, nstat, tile_rec, occ) : go<DD, false>(h, est, A, B, coef, g, t0, t1, nstat, tile_rec, occ);
        GSA_SMALL_D_CASES(CASE)
#undef CASE
        default: return fail(GSA_ERR_UNSUPPORTED, "small Sobol-G kernel supports d <= %d", SMALL_DMAX);
    }
}

}  // namespace gsa
