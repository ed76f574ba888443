// user_kinds.cu — NVRTC build of the K3 kernels for the *_USER kinds.
// The user's μ, σ, g, f arrive as CUDA-C expressions (hdp_user_exprs); they are spliced
// into the same k_mlp.cuh the static library is built from (embedded at build time),
// compiled for sm_100a and loaded with cudaLibraryLoadData.  libnvrtc is dlopen'ed.
#include <dlfcn.h>
#include <string.h>
#include <string>
#include <vector>
#include "user_kinds.h"
#include "hdp_embedded.inc"

namespace hdp {

namespace {
typedef struct _nvrtcProgram* nvrtcProgram;
struct NvrtcApi {
    void* lib = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    int (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};

NvrtcApi* nvrtc_api()
{
    static NvrtcApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> names = {"libnvrtc.so.12", "libnvrtc.so"};
        const char* roots[3] = {getenv("CUDA_HOME"), getenv("CUDA_PATH"), "/usr/local/cuda"};
        for (const char* root : roots) {
            if (!root) continue;
            names.push_back(std::string(root) + "/lib64/libnvrtc.so.12");
            names.push_back(std::string(root) + "/lib64/libnvrtc.so");
        }
        for (auto& n : names) {
            api.lib = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.err = "libnvrtc.so.12 not found (needed for *_USER kinds)"; return; }
#define LOAD(field, sym)                                         \
    *(void**)(&api.field) = dlsym(api.lib, sym);                 \
    if (!api.field) { api.err = std::string("libnvrtc lacks ") + sym; return; }
        LOAD(CreateProgram, "nvrtcCreateProgram")
        LOAD(DestroyProgram, "nvrtcDestroyProgram")
        LOAD(CompileProgram, "nvrtcCompileProgram")
        LOAD(GetProgramLogSize, "nvrtcGetProgramLogSize")
        LOAD(GetProgramLog, "nvrtcGetProgramLog")
        LOAD(GetCUBINSize, "nvrtcGetCUBINSize")
        LOAD(GetCUBIN, "nvrtcGetCUBIN")
        LOAD(AddNameExpression, "nvrtcAddNameExpression")
        LOAD(GetLoweredName, "nvrtcGetLoweredName")
        LOAD(GetErrorString, "nvrtcGetErrorString")
#undef LOAD
    });
    return &api;
}

bool bad_expr(const char* e)
{
    // expressions become macro bodies: keep them single-statement C expressions
    if (!e || !*e) return true;
    for (const char* p = e; *p; ++p)
        if (*p == '#' || *p == ';' || *p == '{' || *p == '}' || *p == '\n' || *p == '\\') return true;
    return strlen(e) > 4000;
}
} // namespace

int user_needed(const hdp_problem* p)
{
    return p->mu_kind == HDP_COEF_USER || p->sigma_kind == HDP_COEF_USER || p->g_kind == HDP_G_USER ||
           p->f_kind == HDP_F_USER || p->f_kind == HDP_F_NL_USER;
}

std::string user_key(const hdp_problem* p, const hdp_user_exprs* u, bool f32, int G)
{
    std::string k = f32 ? "f32/" : "f64/";
    k += std::to_string(G) + "/";
    k += (p->mu_kind == HDP_COEF_USER && u->mu) ? u->mu : "-";
    k += "\x1f";
    k += (p->sigma_kind == HDP_COEF_USER && u->sigma) ? u->sigma : "-";
    k += "\x1f";
    k += (p->g_kind == HDP_G_USER && u->g) ? u->g : "-";
    k += "\x1f";
    k += ((p->f_kind == HDP_F_USER || p->f_kind == HDP_F_NL_USER) && u->f) ? u->f : "-";
    return k;
}

int user_compile(const hdp_problem* p, const hdp_user_exprs* u, bool f32, int G, UserBuild* out)
{
    out->cubin.clear();
    out->log.clear();
    if (!u) { out->log = "problem uses a *_USER kind but cfg.user is NULL"; return HDP_ERR_BAD_ARG; }
    struct Need { bool on; const char* e; const char* what; } needs[4] = {
        {p->mu_kind == HDP_COEF_USER, u->mu, "mu"}, {p->sigma_kind == HDP_COEF_USER, u->sigma, "sigma"},
        {p->g_kind == HDP_G_USER, u->g, "g"}, {p->f_kind == HDP_F_USER || p->f_kind == HDP_F_NL_USER, u->f, "f"}};
    for (auto& n : needs)
        if (n.on && bad_expr(n.e)) {
            out->log = std::string("user expression for ") + n.what +
                       " is missing or not a plain C expression (no ';', braces, '#', newlines)";
            return HDP_ERR_BAD_ARG;
        }
    NvrtcApi* a = nvrtc_api();
    if (!a->err.empty()) { out->log = a->err; return HDP_ERR_NVRTC; }
    const std::string T = f32 ? "float" : "double";
    std::string src;
    src += "typedef signed char int8_t; typedef unsigned char uint8_t; typedef int int32_t; typedef unsigned int uint32_t;\n"
           "typedef long long int64_t; typedef unsigned long long uint64_t; typedef unsigned long uintptr_t;\n";
    const char* macro[4] = {"HDP_USER_MU_EXPR", "HDP_USER_SIGMA_EXPR", "HDP_USER_G_EXPR", "HDP_USER_F_EXPR"};
    for (int i = 0; i < 4; ++i)
        if (needs[i].on) src += std::string("#define ") + macro[i] + " (" + needs[i].e + ")\n";
    src += "#include \"k_mlp.cuh\"\n";
    const std::string n_units = "hdp::mlp_units_kernel<" + T + ", " + std::to_string(G) + ", 1>";
    const std::string n_expand = "hdp::mlp_expand_kernel<" + T + ", " + std::to_string(G) + ">";
    const std::string n_combine = "hdp::mlp_combine_kernel<" + T + ">";
    nvrtcProgram prog = nullptr;
    int rc = a->CreateProgram(&prog, src.c_str(), "hdp_user_kinds.cu", kEmbCount, kEmbSrc, kEmbNames);
    if (rc != 0) { out->log = std::string("nvrtcCreateProgram: ") + a->GetErrorString(rc); return HDP_ERR_NVRTC; }
    a->AddNameExpression(prog, n_units.c_str());
    a->AddNameExpression(prog, n_expand.c_str());
    a->AddNameExpression(prog, n_combine.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-default-device",
                          "-DHDP_MLP_MINBLOCKS=6"};
    rc = a->CompileProgram(prog, 5, opts);
    size_t ls = 0;
    a->GetProgramLogSize(prog, &ls);
    if (ls > 1) {
        out->log.resize(ls);
        a->GetProgramLog(prog, &out->log[0]);
    }
    if (rc != 0) {
        out->log = std::string("NVRTC compilation of the user kinds failed: ") + a->GetErrorString(rc) + "\n" + out->log;
        a->DestroyProgram(&prog);
        return HDP_ERR_BAD_ARG; // the expressions are the caller's
    }
    const char* low = nullptr;
    a->GetLoweredName(prog, n_units.c_str(), &low);   out->units_name = low ? low : "";
    a->GetLoweredName(prog, n_expand.c_str(), &low);  out->expand_name = low ? low : "";
    a->GetLoweredName(prog, n_combine.c_str(), &low); out->combine_name = low ? low : "";
    size_t cs = 0;
    a->GetCUBINSize(prog, &cs);
    out->cubin.resize(cs);
    if (cs) a->GetCUBIN(prog, out->cubin.data());
    a->DestroyProgram(&prog);
    if (!cs || out->units_name.empty()) { out->log += "\nNVRTC produced no cubin"; return HDP_ERR_NVRTC; }
    return HDP_OK;
}

int user_compile_sde(const hdp_problem* p, const hdp_user_exprs* u, bool f32, SdeUserBuild* out)
{
    out->cubin.clear();
    out->log.clear();
    out->names.clear();
    if (!u) { out->log = "the problem uses HDP_COEF_USER but no expressions were set (hdp_set_user_exprs)"; return HDP_ERR_BAD_ARG; }
    struct Need { bool on; const char* e; const char* what; const char* macro; } needs[2] = {
        {p->mu_kind == HDP_COEF_USER, u->mu, "mu", "HDP_USER_MU_EXPR"},
        {p->sigma_kind == HDP_COEF_USER, u->sigma, "sigma", "HDP_USER_SIGMA_EXPR"}};
    for (auto& n : needs)
        if (n.on && bad_expr(n.e)) {
            out->log = std::string("user expression for ") + n.what +
                       " is missing or not a plain C expression (no ';', braces, '#', newlines)";
            return HDP_ERR_BAD_ARG;
        }
    NvrtcApi* a = nvrtc_api();
    if (!a->err.empty()) { out->log = a->err; return HDP_ERR_NVRTC; }
    const std::string T = f32 ? "float" : "double";
    std::string src;
    src += "typedef signed char int8_t; typedef unsigned char uint8_t; typedef int int32_t; typedef unsigned int uint32_t;\n"
           "typedef long long int64_t; typedef unsigned long long uint64_t; typedef unsigned long uintptr_t;\n";
    for (auto& n : needs)
        if (n.on) src += std::string("#define ") + n.macro + " (" + n.e + ")\n";
    src += "#define HDP_FIXED_MU_KIND " + std::to_string(p->mu_kind) + "\n#define HDP_FIXED_SIGMA_KIND " +
           std::to_string(p->sigma_kind) + "\n";
    src += "#include \"k_sde.cuh\"\n";
    const char* tf[2] = {"false", "true"};
    std::vector<std::string> exprs;
    for (int al = 0; al < 2; ++al)
        for (int inj = 0; inj < 2; ++inj)
            exprs.push_back("hdp::sde_loop_kernel<" + T + ", " + tf[al] + ", " + tf[inj] + ", hdp::K1_GENERIC>");
    for (int al = 0; al < 2; ++al)
        for (int inj = 0; inj < 2; ++inj)
            for (int tr = 0; tr < 2; ++tr)
                exprs.push_back("hdp::em_ensemble_kernel<" + T + ", " + tf[al] + ", " + tf[inj] + ", " + tf[tr] + ">");
    nvrtcProgram prog = nullptr;
    int rc = a->CreateProgram(&prog, src.c_str(), "hdp_user_sde.cu", kEmbCount, kEmbSrc, kEmbNames);
    if (rc != 0) { out->log = std::string("nvrtcCreateProgram: ") + a->GetErrorString(rc); return HDP_ERR_NVRTC; }
    for (auto& e : exprs) a->AddNameExpression(prog, e.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-default-device"};
    rc = a->CompileProgram(prog, 4, opts);
    size_t ls = 0;
    a->GetProgramLogSize(prog, &ls);
    if (ls > 1) {
        out->log.resize(ls);
        a->GetProgramLog(prog, &out->log[0]);
    }
    if (rc != 0) {
        out->log = std::string("NVRTC compilation of the user coefficients failed: ") + a->GetErrorString(rc) + "\n" + out->log;
        a->DestroyProgram(&prog);
        return HDP_ERR_BAD_ARG; // the expressions are the caller's
    }
    for (auto& e : exprs) {
        const char* low = nullptr;
        a->GetLoweredName(prog, e.c_str(), &low);
        out->names.push_back(low ? low : "");
    }
    size_t cs = 0;
    a->GetCUBINSize(prog, &cs);
    out->cubin.resize(cs);
    if (cs) a->GetCUBIN(prog, out->cubin.data());
    a->DestroyProgram(&prog);
    if (!cs || out->names[0].empty()) { out->log += "\nNVRTC produced no cubin"; return HDP_ERR_NVRTC; }
    return HDP_OK;
}

} // namespace hdp

extern "C" int hdp_user_kinds_check(const hdp_problem* prob, const hdp_user_exprs* user, int is_f32, char* log,
                                    size_t log_cap)
{
    if (!prob) return HDP_ERR_BAD_ARG;
    if (log && log_cap) log[0] = 0;
    if (!hdp::user_needed(prob)) return HDP_OK;
    hdp::UserBuild b;
    int rc = hdp::user_compile(prob, user, is_f32 != 0, 4, &b);
    if (log && log_cap) {
        strncpy(log, b.log.c_str(), log_cap - 1);
        log[log_cap - 1] = 0;
    }
    return rc;
}

extern "C" int hdp_user_coefs_check(const hdp_problem* prob, const hdp_user_exprs* user, int is_f32, char* log,
                                    size_t log_cap)
{
    if (!prob) return HDP_ERR_BAD_ARG;
    if (log && log_cap) log[0] = 0;
    if (prob->mu_kind != HDP_COEF_USER && prob->sigma_kind != HDP_COEF_USER) return HDP_OK;
    hdp::SdeUserBuild b;
    int rc = hdp::user_compile_sde(prob, user, is_f32 != 0, &b);
    if (log && log_cap) {
        strncpy(log, b.log.c_str(), log_cap - 1);
        log[log_cap - 1] = 0;
    }
    return rc;
}
