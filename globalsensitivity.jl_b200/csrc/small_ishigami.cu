// small_ishigami.cu -- fused design-swap + Ishigami + estimator kernel, d = 3 or 4 (BASELINE configs 1 and 2).
#include "small_launch.cuh"

namespace gsa {

template <int D, bool SECOND>
static int go(gsa_handle* h, int est, const double* A, const double* B, double pa, double pb, const TileGeom& g, int t0, int t1,
              int nstat, double* tile_rec, int* occ) {
    SmallArgs<IshigamiSource<D, SECOND>> a;
    a.src.A = A; a.src.B = B; a.src.buf_col0 = g.buf_col0; a.src.pa = pa; a.src.pb = pb;
    a.tile_rec = tile_rec; a.g = g; a.nout = 1; a.nstat = nstat; a.est = est; a.t_begin = t0; a.t_end = t1;
    if (est == GSA_EI_JANSEN1999) return launch_small<IshigamiSource<D, SECOND>, D, SECOND, true>(h, a, occ);
    return launch_small<IshigamiSource<D, SECOND>, D, SECOND, false>(h, a, occ);
}

int launch_small_ishigami(gsa_handle* h, int d, bool second, int est, const double* A, const double* B, double pa, double pb,
                          const TileGeom& g, int t0, int t1, int nstat, double* tile_rec, int* occ) {
    if (d == 3) return second ? go<3, true>(h, est, A, B, pa, pb, g, t0, t1, nstat, tile_rec, occ) : go<3, false>(h, est, A, B, pa, pb, g, t0, t1, nstat, tile_rec, occ);
    if (d == 4) return second ? go<4, true>(h, est, A, B, pa, pb, g, t0, t1, nstat, tile_rec, occ) : go<4, false>(h, est, A, B, pa, pb, g, t0, t1, nstat, tile_rec, occ);
    return fail(GSA_ERR_UNSUPPORTED, "small Ishigami kernel supports d = 3, 4");
}

}  // namespace gsa
