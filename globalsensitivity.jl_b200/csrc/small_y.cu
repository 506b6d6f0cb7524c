// small_y.cu -- instantiations of the thread-per-sample estimator kernel on a precomputed Y (K3, d <= 12).
#include "small_launch.cuh"

namespace gsa {

template <int D, bool SECOND>
static int go(gsa_handle* h, int est, const double* Y, int64_t n, int nout, int step, const TileGeom& g, int nstat, double* tile_rec, int* occ) {
    SmallArgs<YSource<D, SECOND>> a{};
    a.src.Y = Y; a.src.n = n; a.src.nout = nout; a.src.step = step;
    a.tile_rec = tile_rec; a.g = g; a.nout = nout; a.nstat = nstat; a.est = est; a.t_begin = 0; a.t_end = g.ntiles;
    if (est == GSA_EI_JANSEN1999) return launch_small<YSource<D, SECOND>, D, SECOND, true>(h, a, occ);
    return launch_small<YSource<D, SECOND>, D, SECOND, false>(h, a, occ);
}

int launch_small_y(gsa_handle* h, int d, bool second, int est, const double* Y, int64_t n, int nout, int step, const TileGeom& g,
                   int nstat, double* tile_rec, int* occ) {
    if (second && d > SMALL_DMAX2) return fail(GSA_ERR_UNSUPPORTED, "small second-order kernel supports d <= %d", SMALL_DMAX2);
    switch (d) {
#define CASE(DD) \
    case DD: return second ? go<(DD <= SMALL_DMAX2 ? DD : 1), true>(h, est, Y, n, nout, step, g, nstat, tile_rec, occ) : go<DD, false>(h, est, Y, n, nout, step, g, nstat, tile_rec, occ);
        GSA_SMALL_D_CASES(CASE)
#undef CASE
#define CASE(DD) \
    case DD: return second ? fail(GSA_ERR_UNSUPPORTED, "small second-order kernel supports d <= %d", SMALL_DMAX2) : go<DD, false>(h, est, Y, n, nout, step, g, nstat, tile_rec, occ);
        GSA_SMALL_D_CASES_Y(CASE)
#undef CASE
        default: return fail(GSA_ERR_UNSUPPORTED, "small estimator kernel supports d <= %d", SMALL_DMAX_Y);
    }
}

}  // namespace gsa
