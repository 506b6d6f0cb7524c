// models.cuh -- built-in device models: the registry that replaces the Julia closure `f` of
// gsa(f, ::Sobol, A, B; batch=true) (reference src/sobol_sensitivity.jl:143).  Shapes follow the
// reference's own test models (test/sobol_method.jl:4-24, :146-158) and BASELINE.json's configs.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

// x^4 the way Julia's literal power lowers it (Base.literal_pow -> x*x*x*x)
__device__ __forceinline__ double pow4(double x) { return ((x * x) * x) * x; }

// Scalar/multi-output models with the full-vector ("generic") interface.
//   eval(x, d, p, np, y, nout): x = one sample (d values), p = np parameters, y = nout outputs.
struct IshigamiModel {
    static constexpr int kId = GSA_MODEL_ISHIGAMI;
    __host__ __device__ static int nout(int, int) { return 1; }
    template <class X>
    __device__ __forceinline__ static void eval(const X& x, int, const double* p, int, double* y, int) {
        double s1 = sin(x[0]), s2 = sin(x[1]);
        y[0] = s1 + p[0] * (s2 * s2) + p[1] * pow4(x[2]) * s1;
    }
};

struct SobolGModel {
    static constexpr int kId = GSA_MODEL_SOBOL_G;
    __host__ __device__ static int nout(int, int) { return 1; }
    // one factor g_k(x) -- also used by the product-form fast path
    __device__ __forceinline__ static double factor(double x, double a) { return (fabs(4.0 * x - 2.0) + a) / (1.0 + a); }
    template <class X>
    __device__ __forceinline__ static void eval(const X& x, int d, const double* p, int, double* y, int) {
        double prod = 1.0;
        for (int i = 0; i < d; ++i) prod *= factor(x[i], p[i]);
        y[0] = prod;
    }
};

struct LinearModel {  // y = W x, W nout x d column-major
    static constexpr int kId = GSA_MODEL_LINEAR;
    __host__ __device__ static int nout(int np, int d) { return np / d; }
    template <class X>
    __device__ __forceinline__ static void eval(const X& x, int d, const double* p, int np, double* y, int) {
        int no = np / d;
        for (int o = 0; o < no; ++o) y[o] = 0.0;
        for (int k = 0; k < d; ++k)
            for (int o = 0; o < no; ++o) y[o] = fma(p[o + no * k], x[k], y[o]);
    }
};

struct IshiLinearModel {  // reference test/sobol_method.jl:146-158
    static constexpr int kId = GSA_MODEL_ISHI_LINEAR;
    __host__ __device__ static int nout(int, int) { return 2; }
    template <class X>
    __device__ __forceinline__ static void eval(const X& x, int, const double*, int, double* y, int) {
        double s1 = sin(x[0]), s2 = sin(x[1]);
        y[0] = s1 + 7.0 * (s2 * s2) + 0.1 * pow4(x[2]) * s1;
        y[1] = 7.0 * x[0] + 0.1 * x[1];
    }
};

// A model supplied by the user as relocatable device code and linked into the generic kernel at run time
// (gsa_model_register_fatbin).  The symbol is resolved by nvJitLink.
extern "C" __device__ void gsa_user_model(const double* x, int d, const double* p, int np, double* y, int nout);
struct UserModel {
    template <class X>
    __device__ __forceinline__ static void eval(const X& x, int d, const double* p, int np, double* y, int nout) {
        gsa_user_model(&x[0], d, p, np, y, nout);
    }
};

}  // namespace gsa
