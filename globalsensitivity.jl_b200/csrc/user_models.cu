// user_models.cu -- gsa_model_register_fatbin: link a user's `__device__` model into the generic fused kernel.
//
// The kernel body with an unresolved `gsa_user_model` symbol is embedded in this library as a relocatable sm_100a cubin
// (user_blob.S).  Registration links it against the user's image (fatbin / cubin / PTX / LTO-IR / object, auto-detected)
// with nvJitLink -- loaded with dlopen so that the library itself has no link-time dependency on it -- and loads the
// result with cudaLibraryLoadData.  If the user's symbol has another name, a six-line PTX trampoline is added to the link.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gsa_internal.h"

extern "C" const unsigned char gsa_user_kernel_blob[];
extern "C" const unsigned char gsa_user_kernel_blob_end[];
extern "C" const unsigned char gsa_user_product_blob[];
extern "C" const unsigned char gsa_user_product_blob_end[];
extern "C" const unsigned char gsa_user_additive_blob[];
extern "C" const unsigned char gsa_user_additive_blob_end[];
extern "C" const unsigned char gsa_user_kernel_lto_blob[];
extern "C" const unsigned char gsa_user_kernel_lto_blob_end[];
extern "C" const unsigned char gsa_user_product_lto_blob[];
extern "C" const unsigned char gsa_user_product_lto_blob_end[];
extern "C" const unsigned char gsa_user_additive_lto_blob[];
extern "C" const unsigned char gsa_user_additive_lto_blob_end[];

namespace gsa {

namespace {

// the slice of the nvJitLink API we need (nvJitLink.h), resolved at run time
typedef void* nvJitLinkHandle;
enum { NVJL_INPUT_CUBIN = 1, NVJL_INPUT_PTX = 2, NVJL_INPUT_LTOIR = 3, NVJL_INPUT_ANY = 10 };
struct JitLinkApi {
    void* so = nullptr;
    int (*Create)(nvJitLinkHandle*, uint32_t, const char**) = nullptr;
    int (*Destroy)(nvJitLinkHandle*) = nullptr;
    int (*AddData)(nvJitLinkHandle, int, const void*, size_t, const char*) = nullptr;
    int (*Complete)(nvJitLinkHandle) = nullptr;
    int (*GetLinkedCubinSize)(nvJitLinkHandle, size_t*) = nullptr;
    int (*GetLinkedCubin)(nvJitLinkHandle, void*) = nullptr;
    int (*GetErrorLogSize)(nvJitLinkHandle, size_t*) = nullptr;
    int (*GetErrorLog)(nvJitLinkHandle, char*) = nullptr;
};

void* sym(void* so, const char* base) {
    // exported names are versioned: __nvJitLinkCreate_12_9 ... _12_0
    char name[96];
    for (int minor = 9; minor >= 0; --minor) {
        snprintf(name, sizeof(name), "__%s_12_%d", base, minor);
        if (void* p = dlsym(so, name)) return p;
    }
    return dlsym(so, base);
}

int load_api(JitLinkApi& a) {
    if (a.so) return GSA_OK;
    // The linker must be at least as new as the nvcc that built the embedded kernel (CUDA 12.9): prefer the toolkit's copy
    // over whatever libnvJitLink.so.12 the host process may already have loaded (PyTorch bundles an older one, which
    // rejects a 12.9 cubin as "bad input").  A path with a slash is opened as its own object even if the soname matches.
    std::vector<std::string> names;
    if (const char* e = getenv("GSA_NVJITLINK")) names.push_back(e);
    for (const char* env : {"CUDA_HOME", "CUDA_PATH"})
        if (const char* e = getenv(env)) names.push_back(std::string(e) + "/lib64/libnvJitLink.so.12");
    names.push_back("/usr/local/cuda/lib64/libnvJitLink.so.12");
    names.push_back("libnvJitLink.so.12");
    names.push_back("libnvJitLink.so");
    for (const std::string& n : names)
        if ((a.so = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL))) break;
    if (!a.so) return fail(GSA_ERR_UNSUPPORTED, "cannot load libnvJitLink.so.12 (needed to link user device models): %s", dlerror());
    a.Create = (decltype(a.Create))sym(a.so, "nvJitLinkCreate");
    a.Destroy = (decltype(a.Destroy))sym(a.so, "nvJitLinkDestroy");
    a.AddData = (decltype(a.AddData))sym(a.so, "nvJitLinkAddData");
    a.Complete = (decltype(a.Complete))sym(a.so, "nvJitLinkComplete");
    a.GetLinkedCubinSize = (decltype(a.GetLinkedCubinSize))sym(a.so, "nvJitLinkGetLinkedCubinSize");
    a.GetLinkedCubin = (decltype(a.GetLinkedCubin))sym(a.so, "nvJitLinkGetLinkedCubin");
    a.GetErrorLogSize = (decltype(a.GetErrorLogSize))sym(a.so, "nvJitLinkGetErrorLogSize");
    a.GetErrorLog = (decltype(a.GetErrorLog))sym(a.so, "nvJitLinkGetErrorLog");
    if (!a.Create || !a.Destroy || !a.AddData || !a.Complete || !a.GetLinkedCubinSize || !a.GetLinkedCubin)
        return fail(GSA_ERR_UNSUPPORTED, "libnvJitLink does not export the expected nvJitLink* entry points");
    return GSA_OK;
}

std::string trampoline_ptx(const char* symbol) {
    const std::string s(symbol);
    return std::string(".version 8.8\n.target sm_100a\n.address_size 64\n") +
           ".extern .func " + s + "(.param .b64 a0, .param .b32 a1, .param .b64 a2, .param .b32 a3, .param .b64 a4, .param .b32 a5);\n" +
           ".visible .func gsa_user_model(.param .b64 b0, .param .b32 b1, .param .b64 b2, .param .b32 b3, .param .b64 b4, .param .b32 b5)\n{\n"
           "  .reg .b64 %rd<3>;\n  .reg .b32 %r<3>;\n"
           "  ld.param.u64 %rd0, [b0];\n  ld.param.u32 %r0, [b1];\n  ld.param.u64 %rd1, [b2];\n  ld.param.u32 %r1, [b3];\n"
           "  ld.param.u64 %rd2, [b4];\n  ld.param.u32 %r2, [b5];\n"
           "  {\n  .param .b64 c0;\n  .param .b32 c1;\n  .param .b64 c2;\n  .param .b32 c3;\n  .param .b64 c4;\n  .param .b32 c5;\n"
           "  st.param.b64 [c0], %rd0;\n  st.param.b32 [c1], %r0;\n  st.param.b64 [c2], %rd1;\n  st.param.b32 [c3], %r1;\n"
           "  st.param.b64 [c4], %rd2;\n  st.param.b32 [c5], %r2;\n"
           "  call.uni " + s + ", (c0, c1, c2, c3, c4, c5);\n  }\n  ret;\n}\n";
}

// the same for a per-coordinate function  double <symbol>(int k, double x, const double* p, int np)
std::string trampoline_ptx_separable(const char* symbol, const char* target) {
    const std::string s(symbol), t(target);
    return std::string(".version 8.8\n.target sm_100a\n.address_size 64\n") +
           ".extern .func (.param .b64 r0) " + s + "(.param .b32 a0, .param .b64 a1, .param .b64 a2, .param .b32 a3);\n" +
           ".visible .func (.param .b64 rr) " + t + "(.param .b32 b0, .param .b64 b1, .param .b64 b2, .param .b32 b3)\n{\n"
           "  .reg .b64 %rd<4>;\n  .reg .b32 %r<2>;\n"
           "  ld.param.u32 %r0, [b0];\n  ld.param.u64 %rd0, [b1];\n  ld.param.u64 %rd1, [b2];\n  ld.param.u32 %r1, [b3];\n"
           "  {\n  .param .b32 c0;\n  .param .b64 c1;\n  .param .b64 c2;\n  .param .b32 c3;\n  .param .b64 cr;\n"
           "  st.param.b32 [c0], %r0;\n  st.param.b64 [c1], %rd0;\n  st.param.b64 [c2], %rd1;\n  st.param.b32 [c3], %r1;\n"
           "  call.uni (cr), " + s + ", (c0, c1, c2, c3);\n  ld.param.b64 %rd2, [cr];\n  }\n"
           "  st.param.b64 [rr], %rd2;\n  ret;\n}\n";
}

JitLinkApi g_api;
std::mutex g_api_mu;

// linked cubins, keyed by (kind, symbol, image): a model registered on several devices (gsa_multi) or handles is linked once
struct LinkedCubin {
    int kind;
    std::string symbol;
    uint64_t hash;
    size_t len;
    std::vector<char> cubin;
};
std::vector<LinkedCubin> g_linked;
std::mutex g_linked_mu;

uint64_t fnv1a(const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

}  // namespace

static int link_user_image(const void* image, size_t len, const char* symbol, int kind, std::vector<char>* cubin_out) {
    {
        std::lock_guard<std::mutex> lk(g_api_mu);
        int rc = load_api(g_api);
        if (rc) return rc;
    }
    const JitLinkApi& a = g_api;
    nvJitLinkHandle jl = nullptr;
    const char* opts[] = {"-arch=sm_100a"};
    int e = a.Create(&jl, 1, opts);
    if (e) return fail(GSA_ERR_CUDA, "nvJitLinkCreate failed (%d)", e);
    auto log_and_fail = [&](const char* what, int code) {
        std::string log;
        size_t n = 0;
        if (a.GetErrorLogSize && a.GetErrorLog && a.GetErrorLogSize(jl, &n) == 0 && n > 1) {
            log.resize(n);
            a.GetErrorLog(jl, &log[0]);
        }
        a.Destroy(&jl);
        return fail(GSA_ERR_INVALID, "%s failed (%d): %.380s", what, code, log.c_str());
    };
    const unsigned char *blob = gsa_user_kernel_blob, *blob_end = gsa_user_kernel_blob_end;
    const char* target = "gsa_user_model";
    if (kind == GSA_USER_PRODUCT) { blob = gsa_user_product_blob; blob_end = gsa_user_product_blob_end; target = "gsa_user_factor"; }
    if (kind == GSA_USER_ADDITIVE) { blob = gsa_user_additive_blob; blob_end = gsa_user_additive_blob_end; target = "gsa_user_term"; }
    if (strcmp(symbol, target) != 0) {
        // A trampoline maps `symbol` to the name the kernels call.  If the image defines that name ITSELF the link would see it
        // twice -- and nvJitLink 12.9 segfaults on that "Multiple definition" error when an -lto link has run earlier in the
        // process (reproduced with the plain API, no GPU involved).  So look first: kernels + image alone link cleanly exactly when
        // the image already provides the name.
        nvJitLinkHandle probe = nullptr;
        if (a.Create(&probe, 1, opts) == 0) {
            const bool defines = a.AddData(probe, NVJL_INPUT_CUBIN, blob, (size_t)(blob_end - blob), "gsa_user_kernel") == 0 &&
                                 a.AddData(probe, NVJL_INPUT_ANY, image, len, "user_model") == 0 && a.Complete(probe) == 0;
            a.Destroy(&probe);
            if (defines) {
                a.Destroy(&jl);
                return fail(GSA_ERR_INVALID, "nvJitLink: the image already defines '%s'; registering it as '%s' would define that name twice "
                                             "(register it under '%s', or rename the function)", target, symbol, target);
            }
        }
    }
    if ((e = a.AddData(jl, NVJL_INPUT_CUBIN, blob, (size_t)(blob_end - blob), "gsa_user_kernel"))) return log_and_fail("nvJitLinkAddData(kernel)", e);
    if ((e = a.AddData(jl, NVJL_INPUT_ANY, image, len, "user_model"))) return log_and_fail("nvJitLinkAddData(user image)", e);
    std::string ptx;
    if (strcmp(symbol, target) != 0) {
        ptx = kind == GSA_USER_GENERIC ? trampoline_ptx(symbol) : trampoline_ptx_separable(symbol, target);
        if ((e = a.AddData(jl, NVJL_INPUT_PTX, ptx.c_str(), ptx.size() + 1, "gsa_trampoline"))) return log_and_fail("nvJitLinkAddData(trampoline)", e);
    }
    if ((e = a.Complete(jl))) return log_and_fail("nvJitLinkComplete (is the model built with -rdc=true for sm_100a, with the expected signature?)", e);
    {   // nvJitLinkComplete can report success while the log carries an error (seen with a doubly defined symbol, and the
        // resulting image faults at run time): refuse any link whose error log is not clean
        size_t n = 0;
        if (a.GetErrorLogSize && a.GetErrorLog && a.GetErrorLogSize(jl, &n) == 0 && n > 1) {
            std::string log(n, '\0');
            a.GetErrorLog(jl, &log[0]);
            if (log.find("error") != std::string::npos) return log_and_fail("nvJitLinkComplete (reported errors)", 0);
        }
    }
    size_t cs = 0;
    if ((e = a.GetLinkedCubinSize(jl, &cs))) return log_and_fail("nvJitLinkGetLinkedCubinSize", e);
    cubin_out->resize(cs);
    if ((e = a.GetLinkedCubin(jl, cubin_out->data()))) return log_and_fail("nvJitLinkGetLinkedCubin", e);
    a.Destroy(&jl);
    return GSA_OK;
}

int register_user_model(gsa_handle* h, const void* image, size_t len, const char* symbol, int kind, int nout, int* model_id) {
    if (!image || len == 0 || !symbol || !*symbol || !model_id) return fail(GSA_ERR_INVALID, "image, symbol or model_id is NULL/empty");
    if (kind != GSA_USER_GENERIC && kind != GSA_USER_PRODUCT && kind != GSA_USER_ADDITIVE) return fail(GSA_ERR_INVALID, "unknown user model kind %d", kind);
    if (nout < 1 || nout > 4) return fail(GSA_ERR_UNSUPPORTED, "user models support 1..4 outputs (got %d)", nout);
    if (kind != GSA_USER_GENERIC && nout != 1) return fail(GSA_ERR_INVALID, "separable user models have one output (got %d)", nout);
    for (const char* c = symbol; *c; ++c)
        if (!((*c >= 'a' && *c <= 'z') || (*c >= 'A' && *c <= 'Z') || (*c >= '0' && *c <= '9') || *c == '_' || *c == '$'))
            return fail(GSA_ERR_INVALID, "symbol '%s' is not a plain C identifier (declare the model extern \"C\")", symbol);
    // link once per (kind, symbol, image); every handle / device loads the cached cubin
    const uint64_t hash = fnv1a(image, len);
    const std::vector<char>* cubin = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_linked_mu);
        for (const LinkedCubin& c : g_linked)
            if (c.kind == kind && c.hash == hash && c.len == len && c.symbol == symbol) cubin = &c.cubin;
        if (!cubin) {
            LinkedCubin c;
            c.kind = kind; c.symbol = symbol; c.hash = hash; c.len = len;
            int rc = link_user_image(image, len, symbol, kind, &c.cubin);
            if (rc) return rc;
            g_linked.push_back(std::move(c));
            cubin = &g_linked.back().cubin;
        }
        UserModelEntry um;
        um.nout = nout;
        um.kind = kind;
        // LTO needs the default symbol name (a PTX trampoline cannot be inlined) and an image that carries LTO-IR; whether it
        // does is found out at the first launch (user_separable_kernel)
        const char* env = getenv("GSA_USER_LTO");
        um.try_lto = !(env && atoi(env) == 0) &&
                     strcmp(symbol, kind == GSA_USER_PRODUCT ? "gsa_user_factor" : kind == GSA_USER_ADDITIVE ? "gsa_user_term" : "gsa_user_model") == 0;
        if (um.try_lto) um.image.assign(static_cast<const char*>(image), static_cast<const char*>(image) + len);
        GSA_CUDA(cudaLibraryLoadData(&um.lib, cubin->data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        cudaError_t ce = cudaSuccess;
        static const char* const names[2][2] = {{"gsa_user_fused_first_s", "gsa_user_fused_first_l"}, {"gsa_user_fused_second_s", "gsa_user_fused_second_l"}};
        for (int s2 = 0; s2 < 2 && ce == cudaSuccess; ++s2)
            for (int lg = 0; lg < 2 && ce == cudaSuccess; ++lg) ce = cudaLibraryGetKernel(&um.k_generic[s2][lg], um.lib, names[s2][lg]);
        if (ce != cudaSuccess) {
            cudaLibraryUnload(um.lib);
            return cuda_fail(ce, "cudaLibraryGetKernel");
        }
        h->user_models.push_back(um);
    }
    *model_id = GSA_MODEL_USER_BASE + (int)h->user_models.size() - 1;
    return GSA_OK;
}

// LTO link of ONE separable kernel against the user's image: the kernels' LTO-IR + the image (its LTO-IR member is picked under
// -lto), everything but `kernel_name` dropped before code generation (-kernels-used).  Fails harmlessly (returns false) when the
// image has no LTO-IR or the linker is too old: the caller then uses the call-based kernel.
static bool link_separable_lto(const UserModelEntry& um, const char* kernel_name, std::vector<char>* cubin_out) {
    {
        std::lock_guard<std::mutex> lk(g_api_mu);
        if (load_api(g_api) != GSA_OK) return false;
    }
    const JitLinkApi& a = g_api;
    nvJitLinkHandle jl = nullptr;
    const std::string used = std::string("-kernels-used=") + kernel_name;
    const char* opts[] = {"-arch=sm_100a", "-lto", used.c_str()};
    if (a.Create(&jl, 3, opts)) return false;
    const unsigned char* blob = um.kind == GSA_USER_PRODUCT ? gsa_user_product_lto_blob : um.kind == GSA_USER_ADDITIVE ? gsa_user_additive_lto_blob : gsa_user_kernel_lto_blob;
    const unsigned char* blob_end = um.kind == GSA_USER_PRODUCT ? gsa_user_product_lto_blob_end : um.kind == GSA_USER_ADDITIVE ? gsa_user_additive_lto_blob_end : gsa_user_kernel_lto_blob_end;
    bool ok = a.AddData(jl, NVJL_INPUT_LTOIR, blob, (size_t)(blob_end - blob), "gsa_separable_ltoir") == 0 &&
              a.AddData(jl, NVJL_INPUT_ANY, um.image.data(), um.image.size(), "user_model") == 0 && a.Complete(jl) == 0;
    if (ok) {
        size_t n = 0;
        if (a.GetErrorLogSize && a.GetErrorLog && a.GetErrorLogSize(jl, &n) == 0 && n > 1) {
            std::string log(n, '\0');
            a.GetErrorLog(jl, &log[0]);
            if (log.find("error") != std::string::npos) ok = false;
        }
    }
    size_t cs = 0;
    if (ok) ok = a.GetLinkedCubinSize(jl, &cs) == 0 && cs > 0;
    if (ok) {
        cubin_out->resize(cs);
        ok = a.GetLinkedCubin(jl, cubin_out->data()) == 0;
    }
    a.Destroy(&jl);
    return ok;
}

// kernel `name` of user model `um`, LTO-linked against the user's image on first use (the model inlined); nullptr when the image
// carries no LTO-IR / the symbol is not the default one / the link fails: the caller then takes the call-based kernel
static cudaKernel_t user_lto_kernel(UserModelEntry& um, const char* name) {
    for (const auto& kv : um.sep_kernels)
        if (kv.first == name) return kv.second;
    cudaKernel_t k = nullptr;
    if (um.try_lto) {
        // linked once per (image, kernel) in the process, loaded once per handle
        const uint64_t hash = fnv1a(um.image.data(), um.image.size());
        const std::string key = std::string("lto:") + name;
        const std::vector<char>* cubin = nullptr;
        std::lock_guard<std::mutex> lk(g_linked_mu);
        for (const LinkedCubin& c : g_linked)
            if (c.kind == um.kind && c.hash == hash && c.len == um.image.size() && c.symbol == key) cubin = &c.cubin;
        if (!cubin) {
            LinkedCubin c;
            c.kind = um.kind; c.symbol = key; c.hash = hash; c.len = um.image.size();
            if (link_separable_lto(um, name, &c.cubin)) {
                g_linked.push_back(std::move(c));
                cubin = &g_linked.back().cubin;
            }
        }
        if (cubin) {
            cudaLibrary_t lib = nullptr;
            if (cudaLibraryLoadData(&lib, cubin->data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
                cudaLibraryGetKernel(&k, lib, name) == cudaSuccess) {
                um.lto_libs.push_back(lib);
            } else {
                cudaGetLastError();
                if (lib) cudaLibraryUnload(lib);
                k = nullptr;
            }
        }
        if (!k) um.try_lto = false;   // no LTO-IR in the image (or an old linker): the call-based kernels from here on
        if (getenv("GSA_DEBUG")) fprintf(stderr, "[gsa] user kernel %s: %s\n", name, k ? "LTO (model inlined)" : "call-based");
    }
    if (k) um.sep_kernels.emplace_back(name, k);
    return k;
}

// the generic kernel of a user model: [second order][large instantiation]
const void* user_generic_kernel(gsa_handle* h, int model_id, bool second, bool large) {
    const int idx = model_id - GSA_MODEL_USER_BASE;
    if (idx < 0 || idx >= (int)h->user_models.size()) return nullptr;
    UserModelEntry& um = h->user_models[idx];
    static const char* const names[2][2] = {{"gsa_user_fused_first_s", "gsa_user_fused_first_l"}, {"gsa_user_fused_second_s", "gsa_user_fused_second_l"}};
    if (um.kind == GSA_USER_GENERIC)
        if (cudaKernel_t k = user_lto_kernel(um, names[second ? 1 : 0][large ? 1 : 0])) return (const void*)k;
    return (const void*)um.k_generic[second ? 1 : 0][large ? 1 : 0];
}

// the O(d) kernel of a separable user model for (lanes per sample, coordinates per lane, Jansen | other one-term estimators)
int user_separable_kernel(gsa_handle* h, int model_id, int T, int cpl, bool jansen, const void** kernel) {
    const int idx = model_id - GSA_MODEL_USER_BASE;
    if (idx < 0 || idx >= (int)h->user_models.size()) return fail(GSA_ERR_INVALID, "unknown user model id %d for this handle", model_id);
    UserModelEntry& um = h->user_models[idx];
    if (um.kind == GSA_USER_GENERIC) return fail(GSA_ERR_INVALID, "user model %d is not separable", model_id);
    char name[64];
    snprintf(name, sizeof(name), "gsa_user_sep_T%d_C%d_J%d", T, cpl, jansen ? 1 : 0);
    cudaKernel_t k = user_lto_kernel(um, name);
    if (!k) {
        for (const auto& kv : um.sep_kernels)
            if (kv.first == name) k = kv.second;
        if (!k) {
            GSA_CUDA(cudaLibraryGetKernel(&k, um.lib, name));
            um.sep_kernels.emplace_back(name, k);
        }
    }
    *kernel = (const void*)k;
    return GSA_OK;
}

}  // namespace gsa
