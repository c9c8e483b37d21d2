// sbm_api.cu - host side of libsbm_b200.so (C ABI declared in include/sbm_b200.h).
//
//  * compiles the emitted model device functions + the embedded integrator templates with NVRTC for
//    sm_100a, caches the cubin on disk, loads it through the driver API (libcuda is dlopen'ed so the
//    library itself loads on a machine without a GPU - compute calls then fail loudly);
//  * owns the device copy of the flat project plan;
//  * launches the integrators and the statically compiled loss-epilogue kernels (sbm_epilogue.cuh).
//
// No CPU fallback exists anywhere in this file.
#include <cuda_runtime.h>
#include <cuda.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <mutex>

#include "../../include/sbm_b200.h"
#include "sbm_kernel_args.h"
#include "sbm_epilogue.cuh"
#include "sbm_lm.cuh"
#include "sbm_sampler.cuh"

// generated at build time from sbm_kernel_args.h + kernels/sbm_ode_kernels.cuh
extern const char* sbm_embedded_args_header;
extern const char* sbm_embedded_kernels;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

extern "C" const char* sbm_last_error(void) { return g_err.c_str(); }
extern "C" int sbm_version(void) { return 100; }

#define RT(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(SBM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// driver API through dlopen
// ------------------------------------------------------------------------------------------------
struct Driver {
    void* lib = nullptr;
    CUresult (*cuInit_)(unsigned) = nullptr;
    CUresult (*cuModuleLoadData_)(CUmodule*, const void*) = nullptr;
    CUresult (*cuModuleUnload_)(CUmodule) = nullptr;
    CUresult (*cuModuleGetFunction_)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*cuFuncGetAttribute_)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*cuFuncSetAttribute_)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*cuLaunchKernel_)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*cuGetErrorString_)(CUresult, const char**) = nullptr;
    CUresult (*cuOccupancy_)(int*, CUfunction, int, size_t) = nullptr;
    bool ok = false;
    std::string why;
};

static Driver& driver()
{
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!d.lib) { d.why = "libcuda.so.1 not found (no NVIDIA driver on this machine)"; return; }
#define LOAD(field, sym)                                                                           \
        *(void**)(&d.field) = dlsym(d.lib, sym);                                                   \
        if (!d.field) { d.why = std::string("missing driver symbol ") + sym; return; }
        LOAD(cuInit_, "cuInit")
        LOAD(cuModuleLoadData_, "cuModuleLoadData")
        LOAD(cuModuleUnload_, "cuModuleUnload")
        LOAD(cuModuleGetFunction_, "cuModuleGetFunction")
        LOAD(cuFuncGetAttribute_, "cuFuncGetAttribute")
        LOAD(cuFuncSetAttribute_, "cuFuncSetAttribute")
        LOAD(cuLaunchKernel_, "cuLaunchKernel")
        LOAD(cuGetErrorString_, "cuGetErrorString")
        LOAD(cuOccupancy_, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
#undef LOAD
        d.ok = true;
    });
    return d;
}

static std::string cu_err(CUresult r)
{
    const char* s = nullptr;
    if (driver().cuGetErrorString_) driver().cuGetErrorString_(r, &s);
    return s ? s : "unknown CUDA driver error";
}

#define DRV(call)                                                                                  \
    do {                                                                                           \
        CUresult r__ = (call);                                                                     \
        if (r__ != CUDA_SUCCESS) return fail(SBM_E_CUDA, std::string(#call) + ": " + cu_err(r__)); \
    } while (0)

extern "C" int sbm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int require_device()
{
    if (!driver().ok) return fail(SBM_E_CUDA, "CUDA driver unavailable: " + driver().why);
    if (sbm_device_count() <= 0) return fail(SBM_E_CUDA, "no CUDA device visible (this engine has no CPU fallback)");
    RT(cudaFree(0));   // make sure the primary context exists and is current
    return 0;
}

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct KernelSlot {
    CUfunction fn = nullptr;
    int regs = -1, local = -1, smem = -1;
};

struct sbm_model {
    sbm_model_desc_t desc{};
    std::string arch;
    int G = 1, C = 1, block = 128, block_plain = 128, block_push = 128;
    int dyn_smem = 0, dyn_smem_plain = 0, dyn_smem_push = 0;
    // CTA-transposed sensitivity kernel (kernels/sbm_dp5_ct.cuh): W sensitivity warps x C columns + 1 state warp
    int ctW = 0, ctC = 0, ctCS = 0, ctXW = 0, ct_tmem_cols = 0, ct_dyn_smem = 0, ct_ctas_per_sm = 0, sm_count = 0;
    bool ct_ok = false;
    bool has_rodas = true;
    bool has_rodas_sens = false; // RODAS4 with sensitivities: autonomous model, the emitter provided sbm_dJG
    int layout = 0;              // 0 auto, 1 lane-per-column groups, 2 CTA-transposed
    std::string source;          // full NVRTC translation unit (kept for diagnostics)
    std::string log;
    std::vector<char> cubin;
    bool from_cache = false;
    bool loaded = false;
    int device = -1;             // device whose context holds `mod` (one module per model; see ensure_loaded)
    CUmodule mod = nullptr;
    KernelSlot dp5_plain, dp5_plain_dense, dp5_push, dp5_sens, dp5_ct, rodas4, rodas4_dense, rodas4_sens;
    std::mutex mu;
};

static unsigned long long fnv1a(const std::string& s, unsigned long long seed)
{
    unsigned long long h = 1469598103934665603ull ^ seed;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}

// Work decomposition of the sensitivity kernel (see kernels/sbm_ode_kernels.cuh): a group of G lanes per
// trajectory, C columns of the augmented system per lane.  x (C*N doubles) stays in registers, the six
// stage vectors in shared memory (6*C*N*8 B per lane).  A smaller group evaluates the model's primal part
// for more trajectories per warp instruction, so take the smallest G whose per-lane state fits: C*N <= 24.
static void pick_config(int N, int ncol, int& G, int& C)
{
    for (int g = 1; g <= 32; g *= 2) {
        int c = (ncol + g - 1) / g;
        if (c * N <= 24) { G = g; C = c; return; }
    }
    G = 32;
    C = (ncol + 31) / 32;
}

// CTA size: as large as possible (<= 128) while the stage vectors of one CTA stay below 48 KB, so that
// several CTAs share an SM and no opt-in is needed; one warp per CTA for the largest systems.
static int pick_block(int per_thread_bytes)
{
    for (int blk = 128; blk >= 64; blk /= 2)
        if ((long long)per_thread_bytes * blk <= 48 * 1024) return blk;
    return 32;
}

extern "C" int sbm_model_compile(const char* model_src, const sbm_model_desc_t* desc, sbm_model** out)
{
    if (!model_src || !desc || !out) return fail(SBM_E_INVALID, "sbm_model_compile: null argument");
    if (desc->n_vars < 1 || desc->n_vars > 64) return fail(SBM_E_INVALID, "n_vars must be in [1, 64]");
    if (desc->n_params < 0 || desc->n_free < 0 || desc->n_free > desc->n_params)
        return fail(SBM_E_INVALID, "bad parameter counts");
    sbm_model* m = new sbm_model();
    m->desc = *desc;
    m->arch = desc->arch ? desc->arch : "sm_100a";
    pick_config(desc->n_vars, desc->n_free + 1, m->G, m->C);
    if (desc->group_lanes > 0 && desc->cols_per_lane > 0) {
        m->G = desc->group_lanes;
        m->C = desc->cols_per_lane;
    }
    // per-thread shared memory: stage vectors + this thread's share of the group's parameter row
    const int npp8 = (desc->n_params | 1) * 8;
    m->block = pick_block(6 * m->C * desc->n_vars * 8 + (npp8 + m->G - 1) / m->G);
    m->block_plain = pick_block(6 * desc->n_vars * 8 + npp8);
    // the push-form kernel keeps the state in registers up to SBM_PUSH_YREG_MAX states (sbm_ode_kernels.cuh), beyond that in a
    // fifth shared-memory vector
    int push_yreg_max = 25;
    if (const char* e = getenv("SBM_PUSH_YREG_MAX")) push_yreg_max = atoi(e);
    const int push_nv = desc->n_vars <= push_yreg_max ? 4 : 5;
    m->block_push = pick_block(push_nv * desc->n_vars * 8 + npp8);
    if (desc->block_threads > 0) m->block = desc->block_threads;
    if (m->G < 1 || m->G > 32 || (m->G & (m->G - 1)) || m->G * m->C < desc->n_free + 1 || (m->block % 32) || m->block > 1024) {
        delete m;
        return fail(SBM_E_INVALID, "invalid group_lanes / cols_per_lane / block_threads");
    }
    m->dyn_smem = 6 * m->C * desc->n_vars * m->block * 8;
    m->dyn_smem_plain = 6 * desc->n_vars * m->block_plain * 8;
    m->dyn_smem_push = push_nv * desc->n_vars * m->block_push * 8;
    if (m->dyn_smem > 200 * 1024 || m->dyn_smem_plain > 200 * 1024) {
        delete m;
        return fail(SBM_E_INVALID, "stage vectors do not fit in shared memory for this group shape");
    }

    // CTA-transposed layout (kernels/sbm_dp5_ct.cuh): the emitter chose W warps x C column slots and split the
    // model's primal part into W shares (SBM_CT_W / SBM_CT_C in the model source; 0 = does not fit)
    m->layout = desc->sens_layout > 0 ? desc->sens_layout : 0;
    {
        const int N = desc->n_vars, NF = desc->n_free;
        int NJ = 1, NG = 1, W = 0, C = 0, XW = 0, TM = 0, RV = 0, AS = 0, NB = 2, NJP = 0, JTM = 0;
        const char* q;
        if ((q = strstr(model_src, "#define SBM_NJ "))) NJ = atoi(q + 15);
        if ((q = strstr(model_src, "#define SBM_NG "))) NG = atoi(q + 15);
        if ((q = strstr(model_src, "#define SBM_CT_W "))) W = atoi(q + 17);
        if ((q = strstr(model_src, "#define SBM_CT_C "))) C = atoi(q + 17);
        if ((q = strstr(model_src, "#define SBM_CT_XW "))) XW = atoi(q + 18);
        if ((q = strstr(model_src, "#define SBM_CT_TMEM "))) TM = atoi(q + 20);
        if ((q = strstr(model_src, "#define SBM_CT_REGV "))) RV = atoi(q + 20);
        if ((q = strstr(model_src, "#define SBM_CT_ASYNC "))) AS = atoi(q + 21);
        if ((q = strstr(model_src, "#define SBM_NJU "))) NJ = atoi(q + 16);       // J entries equal up to sign share a slot
        if ((q = strstr(model_src, "#define SBM_NJP "))) NJP = atoi(q + 16);
        if ((q = strstr(model_src, "#define SBM_CT_JTM "))) JTM = atoi(q + 19);
        if (AS) NB = 3;
        if (JTM) NJ = NJP;                        // J buffers in shared memory: row pairs padded for the tensor-memory copies
        (void)NG;
        if (W >= 1 && C >= 1 && NF >= 1) {
            const int CL = NF - (W - 1) * C;
            const long long bytes = 8ll * (NB * NJ * 32 + NB * NF * (N + (N & 1)) * 32 + 16 + 7 * N * 32 + 2 * N * 32 + 2 * N * 32 + desc->n_params * 32 +
                                           (W + 3) * 32 + 32 + 12 * 32 + 4 * 32 + 32) +
                                    4ll * (12 * 32 + 3 * 32) + 8ll * 2 * W * 32 +
                                    8ll * (2 * 4 * 32 + 32) + 4ll * (32 + (desc->n_params + (desc->n_params & 1)) * 32) +
                                    8ll * (2 * 4 * 32 + 32) + 4ll * 32;   // rows staged ahead by the model warp: offsets, values, next time, index   // staged rows (SBM_CT_EM = 4), row base, project columns
            const int tm_need = ((W + 3) / 4) * 4 * C * (N + (N & 1)) + (JTM ? 4 * NJP : 0);
            if (CL >= 1 && CL <= C && CL * N >= 2 && W + XW <= 32 && bytes <= 227 * 1024 && ((TM == 1 && tm_need <= 512) || RV == 1)) {
                m->ctW = W; m->ctC = C; m->ctCS = CL; m->ctXW = XW; m->ct_dyn_smem = (int)bytes; m->ct_ok = true;
                // tensor-memory columns one CTA allocates (power of two >= 32): bounds the CTAs that can run on an SM
                m->ct_tmem_cols = RV ? 0 : tm_need <= 32 ? 32 : tm_need <= 64 ? 64 : tm_need <= 128 ? 128 : tm_need <= 256 ? 256 : 512;
            }
        }
        if (m->layout == 2 && !m->ct_ok) { delete m; return fail(SBM_E_INVALID, "CTA-transposed layout does not fit this model"); }
    }

    std::ostringstream src;
    src << sbm_embedded_args_header << "\n" << model_src << "\n" << sbm_embedded_kernels << "\n";
    if (m->ct_ok) src << "SBM_DEFINE_DP5_CT(sbm_dp5_ct, " << m->ctW << ", " << m->ctC << ", " << m->ctXW << ")\n";
    src << "SBM_DEFINE_DP5(sbm_dp5_plain, 1, 1, 0, " << m->block_plain << ")\n";
    src << "SBM_DEFINE_DP5_DENSE(sbm_dp5_plain_dense, " << m->block_plain << ")\n";
    src << "SBM_DEFINE_DP5_PUSH(sbm_dp5_push, " << m->block_push << ")\n";
    src << "SBM_DEFINE_DP5(sbm_dp5_sens, " << m->G << ", " << m->C << ", 1, " << m->block << ")\n";
    // the Rosenbrock kernel keeps the dense Jacobian and its LU in registers: fine up to ~12 states, hopeless beyond
    m->has_rodas = desc->n_vars <= 12;
    if (m->has_rodas) src << "SBM_DEFINE_RODAS4(sbm_rodas4_plain, " << m->block_plain << ")\n";
    if (m->has_rodas) src << "SBM_DEFINE_RODAS4_DENSE(sbm_rodas4_plain_dense, " << m->block_plain << ")\n";
    m->has_rodas_sens = m->has_rodas && desc->n_free >= 1 && strstr(model_src, "#define SBM_HAS_DJG 1") != nullptr &&
                        strstr(model_src, "#define SBM_AUTONOMOUS 1") != nullptr;
    if (m->has_rodas_sens) src << "SBM_DEFINE_RODAS4_SENS(sbm_rodas4_sens, " << m->block_plain << ")\n";
    m->source = src.str();

    const std::string archopt = "--gpu-architecture=" + m->arch;
    const std::string yregopt = "-DSBM_PUSH_YREG_MAX=" + std::to_string(push_yreg_max);
    std::vector<const char*> opts = {archopt.c_str(), "--std=c++17", "-lineinfo", "--fmad=true",
                                     "--ptxas-options=-v", "-DSBM_NVRTC=1", yregopt.c_str()};
    // extra -D options for experiments (e.g. SBM_KERNEL_DEFINES="-DSBM_CT_PROF=1 -DSBM_CT_Y_SMEM=1")
    std::vector<std::string> extra_defs;
    if (const char* ed = getenv("SBM_KERNEL_DEFINES")) {
        std::istringstream is(ed);
        std::string tok;
        while (is >> tok) extra_defs.push_back(tok);
    }
    for (auto& d : extra_defs) opts.push_back(d.c_str());
    std::string key_material = m->source;
    for (auto o : opts) { key_material += o; key_material += "\n"; }
    char key[64];
    snprintf(key, sizeof key, "%016llx%016llx", fnv1a(key_material, 0), fnv1a(key_material, 0x9e3779b97f4a7c15ull));

    std::string cache_path;
    if (desc->cache_dir && desc->cache_dir[0]) {
        mkdir(desc->cache_dir, 0755);
        cache_path = std::string(desc->cache_dir) + "/sbm_" + key + ".cubin";
        std::ifstream f(cache_path, std::ios::binary);
        if (f) {
            m->cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            if (!m->cubin.empty()) {
                m->from_cache = true;
                std::ifstream lf(cache_path + ".log");
                if (lf) m->log.assign(std::istreambuf_iterator<char>(lf), std::istreambuf_iterator<char>());
            }
        }
    }

    if (m->cubin.empty()) {
        nvrtcProgram prog;
        nvrtcResult r = nvrtcCreateProgram(&prog, m->source.c_str(), "sbm_model.cu", 0, nullptr, nullptr);
        if (r != NVRTC_SUCCESS) { delete m; return fail(SBM_E_NVRTC, std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r)); }
        r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
        size_t logsz = 0;
        nvrtcGetProgramLogSize(prog, &logsz);
        if (logsz > 1) { m->log.resize(logsz); nvrtcGetProgramLog(prog, &m->log[0]); }
        if (r != NVRTC_SUCCESS) {
            std::string msg = std::string("NVRTC compile failed: ") + nvrtcGetErrorString(r) + "\n" + m->log;
            nvrtcDestroyProgram(&prog);
            delete m;
            return fail(SBM_E_NVRTC, msg);
        }
        size_t sz = 0;
        r = nvrtcGetCUBINSize(prog, &sz);
        if (r != NVRTC_SUCCESS || sz == 0) { nvrtcDestroyProgram(&prog); delete m; return fail(SBM_E_NVRTC, "no cubin produced"); }
        m->cubin.resize(sz);
        nvrtcGetCUBIN(prog, m->cubin.data());
        nvrtcDestroyProgram(&prog);
        if (!cache_path.empty()) {
            std::string tmp = cache_path + ".tmp" + std::to_string((long)getpid());
            { std::ofstream f(tmp, std::ios::binary); f.write(m->cubin.data(), (std::streamsize)m->cubin.size()); }
            rename(tmp.c_str(), cache_path.c_str());
            std::ofstream lf(cache_path + ".log");
            lf << m->log;
        }
    }
    *out = m;
    return 0;
}

static int slot_load(sbm_model* m, KernelSlot& s, const char* name)
{
    Driver& d = driver();
    DRV(d.cuModuleGetFunction_(&s.fn, m->mod, name));
    d.cuFuncGetAttribute_(&s.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, s.fn);
    d.cuFuncGetAttribute_(&s.local, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, s.fn);
    d.cuFuncGetAttribute_(&s.smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, s.fn);
    return 0;
}

static int ensure_loaded(sbm_model* m)
{
    std::lock_guard<std::mutex> lk(m->mu);
    int rc = require_device();
    if (rc) return rc;
    int dev = 0;
    RT(cudaGetDevice(&dev));
    if (m->loaded) {
        // the CUmodule lives in the context of the device that was current at first use (one process per GPU is the
        // deployment model): launching it from another device's context would use a foreign handle
        if (dev != m->device)
            return fail(SBM_E_INVALID, "model was loaded on CUDA device " + std::to_string(m->device) + " but device " +
                        std::to_string(dev) + " is current (compile one model per device / process)");
        return 0;
    }
    DRV(driver().cuModuleLoadData_(&m->mod, m->cubin.data()));
    rc = [&]() -> int {
        int r;
        if ((r = slot_load(m, m->dp5_plain, "sbm_dp5_plain"))) return r;
        if ((r = slot_load(m, m->dp5_plain_dense, "sbm_dp5_plain_dense"))) return r;
        if ((r = slot_load(m, m->dp5_push, "sbm_dp5_push"))) return r;
        DRV(driver().cuFuncSetAttribute_(m->dp5_push.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->dyn_smem_push));
        if ((r = slot_load(m, m->dp5_sens, "sbm_dp5_sens"))) return r;
        if (m->has_rodas && (r = slot_load(m, m->rodas4, "sbm_rodas4_plain"))) return r;
        if (m->has_rodas && (r = slot_load(m, m->rodas4_dense, "sbm_rodas4_plain_dense"))) return r;
        if (m->has_rodas_sens && (r = slot_load(m, m->rodas4_sens, "sbm_rodas4_sens"))) return r;
        DRV(driver().cuFuncSetAttribute_(m->dp5_sens.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->dyn_smem));
        DRV(driver().cuFuncSetAttribute_(m->dp5_plain.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->dyn_smem_plain));
        DRV(driver().cuFuncSetAttribute_(m->dp5_plain_dense.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->dyn_smem_plain));
        RT(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev));
        if (m->ct_ok) {
            if ((r = slot_load(m, m->dp5_ct, "sbm_dp5_ct"))) return r;
            DRV(driver().cuFuncSetAttribute_(m->dp5_ct.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->ct_dyn_smem));
            DRV(driver().cuOccupancy_(&m->ct_ctas_per_sm, m->dp5_ct.fn, (m->ctW + m->ctXW) * 32, (size_t)m->ct_dyn_smem));
            if (m->ct_tmem_cols > 0 && m->ct_ctas_per_sm > 512 / m->ct_tmem_cols) m->ct_ctas_per_sm = 512 / m->ct_tmem_cols;
            if (m->ct_ctas_per_sm < 1) m->ct_ok = false;
        }
        return 0;
    }();
    if (rc) {                               // do not leak the module on a failed load (it is loaded again on retry)
        driver().cuModuleUnload_(m->mod);
        m->mod = nullptr;
        return rc;
    }
    m->device = dev;
    m->loaded = true;
    return 0;
}

extern "C" int sbm_model_info(sbm_model* m, sbm_model_info_t* o)
{
    if (!m || !o) return fail(SBM_E_INVALID, "sbm_model_info: null argument");
    if (driver().ok && sbm_device_count() > 0) {
        int rc = ensure_loaded(m);
        if (rc) return rc;
    }
    o->n_vars = m->desc.n_vars; o->n_params = m->desc.n_params; o->n_free = m->desc.n_free;
    o->group_lanes = m->G; o->cols_per_lane = m->C; o->block_threads = m->block;
    o->stage_smem = 1; o->dyn_smem_bytes = m->dyn_smem;
    o->regs_dp5_plain = m->dp5_plain.regs; o->regs_dp5_sens = m->dp5_sens.regs; o->regs_rodas4 = m->rodas4.regs;
    o->local_dp5_plain = m->dp5_plain.local; o->local_dp5_sens = m->dp5_sens.local; o->local_rodas4 = m->rodas4.local;
    o->smem_dp5_plain = m->dp5_plain.smem; o->smem_dp5_sens = m->dp5_sens.smem; o->smem_rodas4 = m->rodas4.smem;
    o->cubin_bytes = (int)m->cubin.size();
    o->from_cache = m->from_cache ? 1 : 0;
    o->ct_warps = m->ct_ok ? m->ctW + m->ctXW : 0; o->ct_cols = m->ct_ok ? m->ctC : 0; o->ct_cols_state = m->ct_ok ? m->ctCS : 0;
    o->ct_dyn_smem_bytes = m->ct_dyn_smem; o->ct_ctas_per_sm = m->ct_ctas_per_sm;
    o->regs_dp5_ct = m->dp5_ct.regs; o->local_dp5_ct = m->dp5_ct.local; o->smem_dp5_ct = m->dp5_ct.smem;
    o->sens_layout = (m->layout == 1 || !m->ct_ok) ? 1 : 2;
    return 0;
}

extern "C" const char* sbm_model_log(const sbm_model* m) { return m ? m->log.c_str() : ""; }

extern "C" void sbm_model_destroy(sbm_model* m)
{
    if (!m) return;
    if (m->loaded && m->mod) driver().cuModuleUnload_(m->mod);
    delete m;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t n)
    {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, n) != cudaSuccess) { cudaGetLastError(); return SBM_E_NOMEM; }
        bytes = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct HostSlot {
    DevBuf theta, sims, jac, resid, ssr, sf, sfgrad, grad, status, counters;
    cudaEvent_t ev_loss = nullptr, ev_copied = nullptr;
    bool pending = false;
    void release()
    {
        theta.release(); sims.release(); jac.release(); resid.release(); ssr.release(); sf.release(); sfgrad.release();
        grad.release(); status.release(); counters.release();
        if (ev_loss) cudaEventDestroy(ev_loss);
        if (ev_copied) cudaEventDestroy(ev_copied);
        ev_loss = ev_copied = nullptr;
    }
};

struct sbm_plan {
    sbm_model* model = nullptr;
    sbm_plan_desc_t d{};
    // device copies
    int *pmap = nullptr, *tout_off = nullptr, *orow_off = nullptr, *row_sim = nullptr, *row_jac = nullptr;
    int *row_group = nullptr, *row_kind = nullptr, *row_aux = nullptr, *grp_off = nullptr, *grp_rows = nullptr;
    double *pfix = nullptr, *y0 = nullptr, *tout = nullptr, *data = nullptr, *sigma = nullptr;
    unsigned long long* row_mask = nullptr;
    std::vector<void*> owned;
    bool jac_full = false;        // the integrator writes EVERY entry of jac (no memset needed before the launch)
    int n_grp_rows = 0;           // length of grp_rows
    int sm_count = 0;
    std::mutex mu;                // the *_host entry points share the workspaces below: one call at a time per plan
    cudaStream_t copy_stream = nullptr;      // copy-out of the *_host entry points
    HostSlot* slots = nullptr;        // [2] grow-only workspaces of the *_host entry points
};

template <typename T>
static int upload(sbm_plan* p, T*& dst, const T* src, size_t n)
{
    dst = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(T);
    void* q = nullptr;
    if (cudaMalloc(&q, bytes) != cudaSuccess) { cudaGetLastError(); return fail(SBM_E_NOMEM, "cudaMalloc failed for plan table"); }
    p->owned.push_back(q);
    if (n && src) RT(cudaMemcpy(q, src, n * sizeof(T), cudaMemcpyHostToDevice));
    dst = (T*)q;
    return 0;
}

static void plan_release_slots(sbm_plan* p);

extern "C" void sbm_plan_destroy(sbm_plan* p)
{
    if (!p) return;
    for (void* q : p->owned) cudaFree(q);
    plan_release_slots(p);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    delete p;
}

extern "C" int sbm_plan_create(sbm_model* m, const sbm_plan_desc_t* d, sbm_plan** out)
{
    if (!d || !out) return fail(SBM_E_INVALID, "sbm_plan_create: null argument");
    if ((m && d->E < 1) || d->n_rows < 0 || d->n_proj < 0) return fail(SBM_E_INVALID, "sbm_plan_create: bad sizes");
    // m == NULL gives a loss-only plan (row tables for sbm_loss, no integrator tables)
    int rc = m ? ensure_loaded(m) : require_device();
    if (rc) return rc;
    sbm_plan* p = new sbm_plan();
    p->model = m;
    p->d = *d;
    p->slots = new HostSlot[2];
#define UP(field, n) if ((rc = upload(p, p->field, d->field, (size_t)(n)))) { sbm_plan_destroy(p); return rc; }
    if (m) {
        const size_t NP = (size_t)m->desc.n_params, N = (size_t)m->desc.n_vars;
        UP(pmap, d->E * NP) UP(pfix, d->E * NP) UP(y0, d->E * N)
        UP(tout_off, d->E + 1) UP(tout, d->n_tout) UP(orow_off, d->n_tout + 1)
        UP(row_sim, d->n_orow) UP(row_jac, d->n_orow) UP(row_mask, d->n_orow)
    }
    UP(data, d->n_rows) UP(sigma, d->n_rows) UP(row_group, d->n_rows) UP(row_kind, d->n_rows) UP(row_aux, d->n_rows)
    UP(grp_off, d->n_groups + 1)
    {
        // grp_off is a host array: read its last entry for the size of grp_rows
        int n_gr = d->n_groups > 0 ? d->grp_off[d->n_groups] : 0;
        UP(grp_rows, n_gr)
        p->n_grp_rows = n_gr;
    }
#undef UP
    {
        int dev = 0;
        RT(cudaGetDevice(&dev));
        RT(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    // Does the integrator write every entry of the Jacobian?  True when every row is an observable row of some output
    // time and every project column is fed by a free parameter of the experiment that owns the row (C4: one experiment,
    // 22 global parameters).  Then sbm_solve skips the memset of [B][n_rows][n_proj] (461 MB per call on C4).
    if (m && !d->jac_atomic && d->n_rows > 0 && d->n_proj > 0 && (long long)d->n_rows * d->n_proj <= (1ll << 24)) {
        const int NP = m->desc.n_params;
        std::vector<char> covered((size_t)d->n_rows * d->n_proj, 0);
        for (int e = 0; e < d->E; ++e)
            for (int kt = d->tout_off[e]; kt < d->tout_off[e + 1]; ++kt)
                for (int r = d->orow_off[kt]; r < d->orow_off[kt + 1]; ++r)
                    for (int q = 0; q < NP; ++q) {
                        const int c = d->pmap[(size_t)e * NP + q];
                        if (c >= 0 && c < d->n_proj && d->row_jac[r] >= 0 && d->row_jac[r] < d->n_rows)
                            covered[(size_t)d->row_jac[r] * d->n_proj + c] = 1;
                    }
        bool all = true;
        for (char c : covered) if (!c) { all = false; break; }
        p->jac_full = all;
    }
    *out = p;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// solve
// ------------------------------------------------------------------------------------------------
extern "C" int sbm_solve(sbm_plan* p, const sbm_solver_opts_t* o, long long B,
                         const double* theta, long long theta_sb, long long theta_sp,
                         double* sims, double* jac, int* status, unsigned long long* counters, void* stream)
{
    if (!p || !o || !theta || !sims || !status || !counters) return fail(SBM_E_INVALID, "sbm_solve: null argument");
    if (B <= 0) return fail(SBM_E_INVALID, "sbm_solve: B must be positive");
    sbm_model* m = p->model;
    if (!m) return fail(SBM_E_INVALID, "sbm_solve: this is a loss-only plan (created without a model)");
    int rc = ensure_loaded(m);
    if (rc) return rc;
    if (o->solver == SBM_SOLVER_RODAS4 && jac && !m->has_rodas_sens)
        return fail(SBM_E_INVALID, "RODAS4 with sensitivities needs an autonomous model with at most 12 states and a free parameter");
    if (o->solver != SBM_SOLVER_DP5 && o->solver != SBM_SOLVER_RODAS4) return fail(SBM_E_INVALID, "unknown solver");
    if (o->solver == SBM_SOLVER_RODAS4 && !m->has_rodas)
        return fail(SBM_E_INVALID, "RODAS4 kernel (in-register dense LU) supports at most 12 states");
    cudaStream_t st = (cudaStream_t)stream;

    SbmSolveArgs a;
    memset(&a, 0, sizeof a);
    a.pmap = p->pmap; a.pfix = p->pfix; a.y0 = p->y0; a.tout_off = p->tout_off; a.tout = p->tout;
    a.orow_off = p->orow_off; a.row_sim = p->row_sim; a.row_jac = p->row_jac; a.row_mask = p->row_mask;
    a.theta = theta; a.theta_sb = theta_sb; a.theta_sp = theta_sp;
    a.sims = sims; a.jac = jac; a.status = status; a.counters = counters;
    a.B = B; a.E = p->d.E; a.n_rows = p->d.n_rows; a.n_proj = p->d.n_proj;
    a.max_steps = o->max_steps > 0 ? o->max_steps : 500000;
    a.jac_atomic = p->d.jac_atomic;
    a.jac_raw = o->jac_raw;
    if (o->jac_raw && jac && p->d.n_proj != m->desc.n_free) return fail(SBM_E_INVALID, "jac_raw needs a plan with n_proj == n_free");
    a.rtol = o->rtol; a.atol = o->atol; a.h0 = o->h0;
    a.n_tout = p->d.n_tout; a.n_orow = p->d.n_orow;
    a.dense = (o->dense_output == 1 || (o->dense_output == 0 && (long long)p->d.n_tout >= 256ll * p->d.E)) ? 1 : 0;

    // (pmap of a fixed model parameter is -1, so a column is "covered" only through a free parameter; with jac_raw the
    //  columns are the model's free parameters themselves)
    if (jac && !(p->jac_full && !o->jac_raw))
        RT(cudaMemsetAsync(jac, 0, (size_t)B * p->d.n_rows * p->d.n_proj * sizeof(double), st));

    const long long ntraj = B * p->d.E;
    KernelSlot* ks;
    long long threads;
    int block = m->block_plain, dyn = 0;
    long long blocks = -1;
    if (o->solver == SBM_SOLVER_RODAS4 && jac) { ks = &m->rodas4_sens; threads = ntraj; }
    else if (o->solver == SBM_SOLVER_RODAS4) { ks = a.dense ? &m->rodas4_dense : &m->rodas4; threads = ntraj; }
    else if (jac && m->ct_ok && m->layout != 1) {
        // persistent CTAs, 32 trajectories in flight each, refilled from a global queue
        ks = &m->dp5_ct; block = (m->ctW + m->ctXW) * 32; dyn = m->ct_dyn_smem; threads = 0;
        blocks = (ntraj + 31) / 32;
        const long long resident = (long long)m->sm_count * m->ct_ctas_per_sm;
        if (blocks > resident) blocks = resident;
        // experiment / determinism tests: a smaller persistent grid changes which trajectories share a CTA and when
        if (const char* mc = getenv("SBM_CT_MAX_CTAS")) { long long v = atoll(mc); if (v >= 1 && v < blocks) blocks = v; }
        RT(cudaMemsetAsync(counters + SBM_CNT_QUEUE, 0, sizeof(unsigned long long), st));
    }
    else if (jac) { ks = &m->dp5_sens; threads = ntraj * m->G; block = m->block; dyn = m->dyn_smem; }
    else if (a.dense) { ks = &m->dp5_plain_dense; threads = ntraj; dyn = m->dyn_smem_plain; }
    else {
        // states only: the push-form kernel (4 stage accumulators in shared memory instead of 6 stage derivatives; measured
        // against the flat stage loop of sbm_dp5_plain: C5 51.1 vs 68.5 ms, C4 31.0 vs 31.8 ms, the 2-state MM model 1.12 vs
        // 1.65 ms).  SBM_PLAIN_PUSH=0 selects the flat loop (experiments, cross-checks)
        bool push = true;
        if (const char* e = getenv("SBM_PLAIN_PUSH")) push = atoi(e) != 0;
        if (push) {
            // persistent grid, trajectories handed out through the queue counter (as for the CTA-transposed kernel)
            ks = &m->dp5_push; threads = ntraj; block = m->block_push; dyn = m->dyn_smem_push;
            int per_sm = 0;
            DRV(driver().cuOccupancy_(&per_sm, ks->fn, block, (size_t)dyn));
            if (per_sm < 1) per_sm = 1;
            blocks = (ntraj + block - 1) / block;
            const long long resident = (long long)m->sm_count * per_sm;
            if (blocks > resident) blocks = resident;
            RT(cudaMemsetAsync(counters + SBM_CNT_QUEUE, 0, sizeof(unsigned long long), st));
        }
        else { ks = &m->dp5_plain; threads = ntraj; dyn = m->dyn_smem_plain; }
    }
    if (blocks < 0) blocks = (threads + block - 1) / block;
    if (blocks > 2147483647ll) return fail(SBM_E_INVALID, "batch too large for one launch");
    void* params[] = {&a};
    DRV(driver().cuLaunchKernel_(ks->fn, (unsigned)blocks, 1, 1, (unsigned)block, 1, 1, (unsigned)dyn, (CUstream)st, params, nullptr));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// loss epilogue
// ------------------------------------------------------------------------------------------------
static int loss_common(sbm_plan* p, int loss_kind, int flags, int phase, long long B,
                       const double* theta, long long theta_sb, long long theta_sp,
                       double* sims, double* jac, double* sums, double* resid, double* ssr, double* sf, double* sfgrad,
                       double* grad, void* stream, const char* who)
{
    if (!p || !theta || !sims) return fail(SBM_E_INVALID, std::string(who) + ": null argument");
    if (phase != SBM_LOSS_PHASE_SUMS) {
        if (!resid || !ssr) return fail(SBM_E_INVALID, std::string(who) + ": null argument");
        if (p->d.n_groups > 0 && !sf) return fail(SBM_E_INVALID, std::string(who) + ": sf buffer required when scale-factor groups exist");
        if (jac && p->d.n_groups > 0 && !sfgrad) return fail(SBM_E_INVALID, std::string(who) + ": sfgrad workspace required with jac");
    }
    if (phase != SBM_LOSS_PHASE_ALL && p->d.n_groups > 0 && !sums) return fail(SBM_E_INVALID, std::string(who) + ": sums buffer required");
    if (loss_kind < 0 || loss_kind > 2) return fail(SBM_E_INVALID, std::string(who) + ": unknown loss");
    if (B <= 0) return fail(SBM_E_INVALID, std::string(who) + ": B must be positive");
    int rc = require_device();
    if (rc) return rc;
    SbmLossArgs a;
    memset(&a, 0, sizeof a);
    a.data = p->data; a.sigma = p->sigma; a.row_group = p->row_group; a.row_kind = p->row_kind; a.row_aux = p->row_aux;
    a.grp_off = p->grp_off; a.grp_rows = p->grp_rows;
    a.theta = theta; a.theta_sb = theta_sb; a.theta_sp = theta_sp;
    a.sims = sims; a.jac = jac; a.resid = resid; a.ssr = ssr; a.sf = sf; a.sfgrad = sfgrad; a.grad = grad; a.sums = sums;
    a.B = B; a.n_rows = p->d.n_rows; a.n_proj = p->d.n_proj; a.n_groups = p->d.n_groups;
    a.loss_kind = loss_kind; a.flags = flags; a.phase = phase;
    RT(sbm_launch_loss(a, (cudaStream_t)stream, p->sm_count, p->n_grp_rows));
    return 0;
}

extern "C" int sbm_loss(sbm_plan* p, int loss_kind, int flags, long long B,
                        const double* theta, long long theta_sb, long long theta_sp,
                        double* sims, double* jac, double* resid, double* ssr, double* sf, double* sfgrad,
                        double* grad, void* stream)
{
    return loss_common(p, loss_kind, flags, SBM_LOSS_PHASE_ALL, B, theta, theta_sb, theta_sp, sims, jac, nullptr, resid, ssr, sf,
                       sfgrad, grad, stream, "sbm_loss");
}

extern "C" int sbm_loss_sums(sbm_plan* p, int loss_kind, int flags, long long B,
                             const double* theta, long long theta_sb, long long theta_sp,
                             double* sims, double* jac, double* sums, void* stream)
{
    return loss_common(p, loss_kind, flags, SBM_LOSS_PHASE_SUMS, B, theta, theta_sb, theta_sp, sims, jac, sums, nullptr, nullptr,
                       nullptr, nullptr, nullptr, stream, "sbm_loss_sums");
}

extern "C" int sbm_loss_finish(sbm_plan* p, int loss_kind, int flags, long long B,
                               const double* theta, long long theta_sb, long long theta_sp,
                               double* sims, double* jac, const double* sums, double* resid, double* ssr, double* sf,
                               double* sfgrad, double* grad, void* stream)
{
    return loss_common(p, loss_kind, flags, SBM_LOSS_PHASE_FINISH, B, theta, theta_sb, theta_sp, sims, jac,
                       const_cast<double*>(sums), resid, ssr, sf, sfgrad, grad, stream, "sbm_loss_finish");
}

extern "C" int sbm_normal_equations(long long B, int n_rows, int n_proj, const double* jac, const double* resid, double* out,
                                    void* stream)
{
    if (!jac || !resid || !out) return fail(SBM_E_INVALID, "sbm_normal_equations: null argument");
    if (B <= 0 || n_rows <= 0 || n_proj <= 0) return fail(SBM_E_INVALID, "sbm_normal_equations: B, n_rows, n_proj must be positive");
    if (n_proj > 32) return fail(SBM_E_INVALID, "sbm_normal_equations: n_proj > 32 is not supported");
    int rc = require_device();
    if (rc) return rc;
    int dev = 0, sms = 0;
    RT(cudaGetDevice(&dev));
    RT(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SbmGramArgs a;
    a.jac = jac; a.resid = resid; a.out = out; a.B = B; a.n_rows = n_rows; a.n_proj = n_proj;
    RT(sbm_launch_gram(a, (cudaStream_t)stream, sms));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched Levenberg-Marquardt step
// ------------------------------------------------------------------------------------------------
extern "C" int sbm_lm_step(long long B, int n_rows, int n_proj, const double* jac, const double* resid, const double* lambda,
                           double* dscale, double* delta, double* pred, double* gmax, int* info, void* stream)
{
    if (!jac || !resid || !lambda || !delta || !pred || !gmax || !info) return fail(SBM_E_INVALID, "sbm_lm_step: null argument");
    if (B <= 0 || n_rows <= 0 || n_proj <= 0) return fail(SBM_E_INVALID, "sbm_lm_step: B, n_rows, n_proj must be positive");
    if (n_proj > 32) return fail(SBM_E_INVALID, "sbm_lm_step: n_proj > 32 is not supported");
    int rc = require_device();
    if (rc) return rc;
    SbmLmArgs a;
    a.jac = jac; a.resid = resid; a.lambda = lambda; a.dscale = dscale; a.delta = delta; a.pred = pred; a.gmax = gmax; a.info = info;
    a.B = B; a.n_rows = n_rows; a.n_proj = n_proj;
    if (sbm_launch_lm_step(a, (cudaStream_t)stream)) return fail(SBM_E_INVALID, "sbm_lm_step: unsupported shape");
    RT(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched Metropolis sampler: proposal and accept / reject on the device
// ------------------------------------------------------------------------------------------------
extern "C" int sbm_mc_propose(long long W, int P, const double* samp, const double* cur, double* nxt, const double* normals,
                              unsigned long long seed, unsigned step, void* stream)
{
    if (!samp || !cur || !nxt) return fail(SBM_E_INVALID, "sbm_mc_propose: null argument");
    if (W <= 0 || P <= 0 || P > 4096) return fail(SBM_E_INVALID, "sbm_mc_propose: bad sizes");
    int rc = require_device();
    if (rc) return rc;
    int dev = 0, sms = 0;
    RT(cudaGetDevice(&dev));
    RT(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SbmMcArgs a;
    memset(&a, 0, sizeof a);
    a.W = W; a.P = P; a.samp = samp; a.cur = const_cast<double*>(cur); a.nxt = nxt; a.normals = normals; a.seed = seed; a.step = step;
    RT(sbm_launch_mc_propose(a, (cudaStream_t)stream, sms));
    return 0;
}

extern "C" int sbm_mc_accept(sbm_plan* p, long long W, int init, double temperature, double* cur, const double* nxt, double* curF,
                             const double* ssr_new, const double* sims_new, const double* sf_mu, const double* sf_sig,
                             const double* uniforms, unsigned long long seed, unsigned step, double* accepted,
                             double* ens_out, double* fes_out, void* stream)
{
    if (!p || !cur || !curF || !ssr_new || !sims_new || !accepted) return fail(SBM_E_INVALID, "sbm_mc_accept: null argument");
    if (!init && !nxt) return fail(SBM_E_INVALID, "sbm_mc_accept: null argument");
    if (W <= 0 || !(temperature > 0.0)) return fail(SBM_E_INVALID, "sbm_mc_accept: bad sizes / temperature");
    if ((ens_out == nullptr) != (fes_out == nullptr)) return fail(SBM_E_INVALID, "sbm_mc_accept: ens_out and fes_out go together");
    int rc = require_device();
    if (rc) return rc;
    SbmMcArgs a;
    memset(&a, 0, sizeof a);
    a.W = W; a.P = p->d.n_proj; a.n_rows = p->d.n_rows; a.n_groups = (sf_mu && sf_sig) ? p->d.n_groups : 0;
    a.cur = cur; a.nxt = const_cast<double*>(nxt); a.curF = curF; a.ssr_new = ssr_new; a.sims_new = sims_new;
    a.temperature = temperature; a.data = p->data; a.sigma = p->sigma; a.grp_off = p->grp_off; a.grp_rows = p->grp_rows;
    a.sf_mu = sf_mu; a.sf_sig = sf_sig; a.uniforms = uniforms; a.seed = seed; a.step = step; a.accepted = accepted;
    a.ens_out = ens_out; a.fes_out = fes_out; a.init = init;
    RT(sbm_launch_mc_accept(a, (cudaStream_t)stream, p->sm_count));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fused host-buffer entry points
// ------------------------------------------------------------------------------------------------
// Host evaluations.  A call copies theta in, integrates, runs the loss epilogue and copies the results out.  On C4 the
// copy-out is 494 MB of scaled Jacobians per 65,536 sets (~9 ms over PCIe against ~400 ms of integration).  It cannot hide
// inside ONE call: the persistent integrator fills every SM (all registers, 150 KB of shared memory), so a second launch
// only starts in the tail of the first, and the loss kernel of an early chunk then queues behind the next chunk's CTAs -
// measured on C4: 2 chunks (80 % / 20 %, two streams) 155 k solves/s end to end, 8 equal chunks 122 k, serial 158 k.
// What does hide it is the NEXT evaluation: sbm_jacobian_host_begin enqueues an evaluation on the caller's stream and its
// copy-out on the plan's copy stream (two workspace slots), sbm_host_end waits for a slot.  An optimiser that evaluates
// trial points back to back (or a sampler) keeps two in flight.
static int host_eval(sbm_plan* p, const sbm_solver_opts_t* o, int loss_kind, int flags, long long B, bool with_jac,
                     const double* theta_host, double* jac_host, double* model_jac_host, double* resid_host,
                     double* ssr_host, double* grad_host, double* sf_host, double* sims_host, int* status_host,
                     unsigned long long* counters_host, void* stream, int slot, bool wait)
{
    if (!p || !o || !theta_host) return fail(SBM_E_INVALID, "null argument");
    if (B <= 0) return fail(SBM_E_INVALID, "B must be positive");
    if (slot < 0 || slot > 1) return fail(SBM_E_INVALID, "slot must be 0 or 1");
    int rc = require_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(p->mu);          // the workspaces below belong to the plan: one host call at a time
    HostSlot& w = p->slots[slot];
    if (w.pending) return fail(SBM_E_INVALID, "this slot still has an evaluation in flight (call sbm_host_end first)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t R = (size_t)p->d.n_rows, P = (size_t)p->d.n_proj, Gn = (size_t)p->d.n_groups, E = (size_t)p->d.E;
    const size_t nB = (size_t)B;
    if (w.theta.ensure(nB * P * 8) || w.sims.ensure(nB * R * 8) || w.resid.ensure(nB * R * 8) ||
        w.ssr.ensure(nB * 8) || w.sf.ensure((nB * Gn + 1) * 8) || w.status.ensure(nB * E * 4) ||
        w.counters.ensure(SBM_N_COUNTERS * 8))
        return fail(SBM_E_NOMEM, "device workspace allocation failed");
    if (with_jac && (w.jac.ensure(nB * R * P * 8) || w.sfgrad.ensure((nB * Gn * P + 1) * 8) || w.grad.ensure(nB * P * 8)))
        return fail(SBM_E_NOMEM, "device workspace allocation failed (Jacobian)");
    if (!p->copy_stream) RT(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    if (!w.ev_loss) {
        RT(cudaEventCreateWithFlags(&w.ev_loss, cudaEventDisableTiming));
        RT(cudaEventCreateWithFlags(&w.ev_copied, cudaEventDisableTiming));
    }
    double* d_theta = (double*)w.theta.p;
    double* d_sims = (double*)w.sims.p;
    double* d_jac = with_jac ? (double*)w.jac.p : nullptr;
    RT(cudaMemcpyAsync(d_theta, theta_host, nB * P * 8, cudaMemcpyHostToDevice, st));
    RT(cudaMemsetAsync(w.counters.p, 0, SBM_N_COUNTERS * 8, st));
    RT(cudaMemsetAsync(d_sims, 0, nB * R * 8, st));
    rc = sbm_solve(p, o, B, d_theta, (long long)P, 1, d_sims, d_jac, (int*)w.status.p, (unsigned long long*)w.counters.p, stream);
    if (rc) return rc;
    if (with_jac && model_jac_host) RT(cudaMemcpyAsync(model_jac_host, d_jac, nB * R * P * 8, cudaMemcpyDeviceToHost, st));
    rc = sbm_loss(p, loss_kind, flags, B, d_theta, (long long)P, 1, d_sims, d_jac, (double*)w.resid.p, (double*)w.ssr.p,
                  (double*)w.sf.p, with_jac ? (double*)w.sfgrad.p : nullptr, with_jac ? (double*)w.grad.p : nullptr, stream);
    if (rc) return rc;
    // copy-out on the plan's copy stream: the caller's stream is free for the next evaluation (other slot)
    RT(cudaEventRecord(w.ev_loss, st));
    cudaStream_t sc = p->copy_stream;
    RT(cudaStreamWaitEvent(sc, w.ev_loss, 0));
    if (with_jac && jac_host) RT(cudaMemcpyAsync(jac_host, d_jac, nB * R * P * 8, cudaMemcpyDeviceToHost, sc));
    if (resid_host) RT(cudaMemcpyAsync(resid_host, w.resid.p, nB * R * 8, cudaMemcpyDeviceToHost, sc));
    if (sims_host) RT(cudaMemcpyAsync(sims_host, d_sims, nB * R * 8, cudaMemcpyDeviceToHost, sc));
    if (ssr_host) RT(cudaMemcpyAsync(ssr_host, w.ssr.p, nB * 8, cudaMemcpyDeviceToHost, sc));
    if (sf_host && Gn) RT(cudaMemcpyAsync(sf_host, w.sf.p, nB * Gn * 8, cudaMemcpyDeviceToHost, sc));
    if (status_host) RT(cudaMemcpyAsync(status_host, w.status.p, nB * E * 4, cudaMemcpyDeviceToHost, sc));
    if (with_jac && grad_host) RT(cudaMemcpyAsync(grad_host, w.grad.p, nB * P * 8, cudaMemcpyDeviceToHost, sc));
    if (counters_host) RT(cudaMemcpyAsync(counters_host, w.counters.p, SBM_N_COUNTERS * 8, cudaMemcpyDeviceToHost, sc));
    RT(cudaEventRecord(w.ev_copied, sc));
    w.pending = true;
    if (wait) {
        RT(cudaEventSynchronize(w.ev_copied));
        w.pending = false;
    }
    return 0;
}

extern "C" int sbm_host_end(sbm_plan* p, int slot)
{
    if (!p || slot < 0 || slot > 1) return fail(SBM_E_INVALID, "sbm_host_end: bad argument");
    std::lock_guard<std::mutex> lk(p->mu);
    HostSlot& w = p->slots[slot];
    if (!w.pending) return 0;
    RT(cudaEventSynchronize(w.ev_copied));
    w.pending = false;
    return 0;
}

static void plan_release_slots(sbm_plan* p)
{
    if (!p->slots) return;
    for (int i = 0; i < 2; ++i) p->slots[i].release();
    delete[] p->slots;
    p->slots = nullptr;
}

extern "C" int sbm_residuals_host(sbm_plan* p, const sbm_solver_opts_t* opts, int loss_kind, long long B,
                                  const double* theta_host, double* resid_host, double* ssr_host, double* sf_host,
                                  double* sims_host, int* status_host, unsigned long long* counters_host, void* stream)
{
    return host_eval(p, opts, loss_kind, 0, B, false, theta_host, nullptr, nullptr, resid_host, ssr_host, nullptr, sf_host,
                     sims_host, status_host, counters_host, stream, 0, true);
}

extern "C" int sbm_jacobian_host_begin(sbm_plan* p, const sbm_solver_opts_t* opts, int loss_kind, int flags, long long B,
                                       const double* theta_host, double* jac_host, double* resid_host, double* ssr_host,
                                       double* grad_host, double* sf_host, double* sims_host, int* status_host,
                                       unsigned long long* counters_host, void* stream, int slot)
{
    return host_eval(p, opts, loss_kind, flags, B, true, theta_host, jac_host, nullptr, resid_host, ssr_host, grad_host,
                     sf_host, sims_host, status_host, counters_host, stream, slot, false);
}

extern "C" int sbm_jacobian_host(sbm_plan* p, const sbm_solver_opts_t* opts, int loss_kind, int flags, long long B,
                                 const double* theta_host, double* jac_host, double* model_jac_host, double* resid_host,
                                 double* ssr_host, double* grad_host, double* sf_host, double* sims_host, int* status_host,
                                 unsigned long long* counters_host, void* stream)
{
    return host_eval(p, opts, loss_kind, flags, B, true, theta_host, jac_host, model_jac_host, resid_host, ssr_host, grad_host,
                     sf_host, sims_host, status_host, counters_host, stream, 0, true);
}

// ------------------------------------------------------------------------------------------------
// FP64 roofline denominator
// ------------------------------------------------------------------------------------------------
extern "C" int sbm_fp64_peak(int iters, double* tflops_out, double* ms_out)
{
    int rc = require_device();
    if (rc) return rc;
    if (iters <= 0) iters = 4096;
    double tf = 0.0, ms = 0.0;
    cudaError_t e = sbm_run_fp64_peak(iters, &tf, &ms);
    if (e != cudaSuccess) return fail(SBM_E_CUDA, std::string("fp64 peak kernel: ") + cudaGetErrorString(e));
    if (tflops_out) *tflops_out = tf;
    if (ms_out) *ms_out = ms;
    return 0;
}
