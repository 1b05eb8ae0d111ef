// libhydrob200.so — host side of the C-ABI declared in include/hydrob200.h.
//
// Responsibilities: splice a host-generated CUDA C body into the hand-written kernel templates
// (templates/*.cu, embedded at build time), compile with NVRTC for sm_100a, cache cubins on
// disk, load them lazily through the driver API (entry points fetched with
// cudaGetDriverEntryPoint, so the library itself loads on a box without libcuda) and launch.
// No CPU execution path exists: without a device every run call returns HB_ERR_NO_DEVICE.
#include "../../include/hydrob200.h"

#include <cuda.h>
#include <cuda_runtime_api.h>
#include <immintrin.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <sched.h>
#include <string>
#include <thread>
#include <vector>

#include "templates_embedded.h"   // generated: kStepperTemplate, kRouteTemplate

namespace {

thread_local std::string g_err = "";
int g_device = -1;
std::string g_cache_dir;
std::mutex g_mu;

int fail(int code, const char *fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define HB_CUDA(expr)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) return fail(HB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// ---- driver entry points ------------------------------------------------------------------------
struct Driver {
    bool ok = false;
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
    decltype(&cuFuncGetAttribute) FuncGetAttribute = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
    decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) Occupancy = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    decltype(&cuModuleGetGlobal) ModuleGetGlobal = nullptr;
    decltype(&cuMemcpyDtoDAsync) MemcpyDtoDAsync = nullptr;
} g_drv;

template <typename F>
bool load_sym(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
        return false;
    fn = reinterpret_cast<F>(p);
    return true;
}

int ensure_device() {
    std::lock_guard<std::mutex> lk(g_mu);
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(HB_ERR_NO_DEVICE, "no CUDA device available (%s); hydrob200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (g_device < 0) g_device = 0;
    HB_CUDA(cudaSetDevice(g_device));
    HB_CUDA(cudaFree(nullptr));   // create / bind the primary context
    if (!g_drv.ok) {
        bool ok = load_sym("cuModuleLoadData", g_drv.ModuleLoadData) && load_sym("cuModuleUnload", g_drv.ModuleUnload) &&
                  load_sym("cuModuleGetFunction", g_drv.ModuleGetFunction) &&
                  load_sym("cuFuncSetAttribute", g_drv.FuncSetAttribute) &&
                  load_sym("cuFuncGetAttribute", g_drv.FuncGetAttribute) && load_sym("cuLaunchKernel", g_drv.LaunchKernel) &&
                  load_sym("cuOccupancyMaxActiveBlocksPerMultiprocessor", g_drv.Occupancy) &&
                  load_sym("cuGetErrorString", g_drv.GetErrorString) && load_sym("cuModuleGetGlobal", g_drv.ModuleGetGlobal) &&
                  load_sym("cuMemcpyDtoDAsync", g_drv.MemcpyDtoDAsync);
        if (!ok) return fail(HB_ERR_CUDA, "could not resolve CUDA driver entry points");
        g_drv.ok = true;
    }
    return HB_OK;
}

int drv_fail(CUresult r, const char *what) {
    const char *s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    return fail(HB_ERR_CUDA, "%s failed: %s", what, s ? s : "unknown driver error");
}
#define HB_DRV(expr)                                  \
    do {                                              \
        CUresult r_ = (expr);                         \
        if (r_ != CUDA_SUCCESS) return drv_fail(r_, #expr); \
    } while (0)

// ---- small utilities -----------------------------------------------------------------------------
uint64_t fnv1a(const std::string &s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    return h;
}

std::string lib_dir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<void *>(&hb_version), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        return k == std::string::npos ? "." : p.substr(0, k);
    }
    return ".";
}

std::string cache_dir() {
    if (!g_cache_dir.empty()) return g_cache_dir;
    const char *env = getenv("HB_CACHE_DIR");
    std::string d = env && *env ? env : lib_dir() + "/cache";
    mkdir(d.c_str(), 0755);
    return d;
}

bool read_file(const std::string &path, std::vector<char> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    bool ok = n > 0 && fread(out.data(), 1, n, f) == (size_t)n;
    fclose(f);
    return ok;
}

void write_file_atomic(const std::string &path, const void *data, size_t n) {
    std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    bool ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    if (ok) rename(tmp.c_str(), path.c_str());
    else unlink(tmp.c_str());
}

// Replace each `//@@HB:NAME` marker line of the template by the body's section of that name.
// Body format: lines `//@@HB:NAME` open a section that runs to the next marker.
std::string splice(const char *tmpl, const char *body) {
    struct Sec { std::string name, text; };
    std::vector<Sec> secs;
    {
        std::string b = body ? body : "";
        size_t pos = 0;
        Sec *cur = nullptr;
        while (pos < b.size()) {
            size_t eol = b.find('\n', pos);
            if (eol == std::string::npos) eol = b.size();
            std::string line = b.substr(pos, eol - pos);
            if (line.rfind("//@@HB:", 0) == 0) {
                std::string nm = line.substr(7);
                while (!nm.empty() && (nm.back() == '\r' || nm.back() == ' ')) nm.pop_back();
                secs.push_back({nm, ""});
                cur = &secs.back();
            } else if (cur) {
                cur->text += line;
                cur->text += '\n';
            }
            pos = eol + 1;
        }
    }
    std::string t = tmpl, out;
    size_t pos = 0;
    while (pos < t.size()) {
        size_t eol = t.find('\n', pos);
        if (eol == std::string::npos) eol = t.size();
        std::string line = t.substr(pos, eol - pos);
        size_t k = line.find_first_not_of(" \t");
        if (k != std::string::npos && line.compare(k, 7, "//@@HB:") == 0) {
            std::string nm = line.substr(k + 7);
            while (!nm.empty() && (nm.back() == '\r' || nm.back() == ' ')) nm.pop_back();
            out += "// ---- begin generated section " + nm + " ----\n";
            for (auto &s : secs)
                if (s.name == nm) out += s.text;
            out += "// ---- end generated section " + nm + " ----\n";
        } else {
            out += line;
            out += '\n';
        }
        pos = eol + 1;
    }
    return out;
}

}  // namespace

// ---- model object ----------------------------------------------------------------------------------
struct hb_model_s {
    int kind = 0;   // 0 stepper, 1 route, 2 fused grid, 3 adjoint
    hb_adjoint_spec aspec{};
    hb_stepper_spec spec{};
    hb_route_spec rspec{};
    hb_grid_spec gspec{};
    std::string source;
    std::string kernel_name;
    std::vector<char> cubin;
    bool from_cache = false;
    // device side (lazy)
    CUmodule module = nullptr;
    CUfunction func = nullptr;
    int dyn_smem = 0;
    int block = 128;
    int uhw_total = 0, uhh_total = 0;
    int wave_tc = 0;            // WAVE grid kernel: steps per launch after clamping
    int wave_qslots = 2;        // WAVE grid kernel: outflow words per cell
    bool wave_err_init = false;
    CUdeviceptr nn_const = 0;   // module-scope __constant__ hb_nn_const[]
    // workspace for hb_run (host-pointer entry point)
    struct Buf { void *p = nullptr; size_t cap = 0; };
    Buf ws[16];
    Buf pin_in, pin_out;        // pinned staging rings of hb_run for pageable host arrays
    cudaStream_t stream = nullptr, stream_h2d = nullptr, stream_d2h = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
};

namespace {

int nvrtc_compile(const std::string &src, const std::vector<std::string> &opts, std::vector<char> &cubin, bool &from_cache) {
    int maj = 0, min = 0;
    nvrtcVersion(&maj, &min);
    std::string keysrc = src;
    for (auto &o : opts) keysrc += "\n//opt:" + o;
    keysrc += "\n//nvrtc:" + std::to_string(maj) + "." + std::to_string(min);
    char name[64];
    snprintf(name, sizeof name, "%016llx_%zu.cubin", (unsigned long long)fnv1a(keysrc), keysrc.size());
    std::string path = cache_dir() + "/" + name;
    if (!getenv("HB_NO_CACHE") && read_file(path, cubin)) {
        from_cache = true;
        return HB_OK;
    }
    from_cache = false;
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "hb_kernel.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(HB_ERR_COMPILE, "nvrtcCreateProgram failed");
    std::vector<const char *> copts;
    for (auto &o : opts) copts.push_back(o.c_str());
    nvrtcResult r = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
    if (r != NVRTC_SUCCESS) {
        if (const char *dump = getenv("HB_DUMP_FAILED_SOURCE")) {   // debugging aid: keep the rejected source
            if (FILE *f = fopen(dump, "w")) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
        }
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        if (log.size() > 3500) log.resize(3500);
        return fail(HB_ERR_COMPILE, "NVRTC: %s\n%s", nvrtcGetErrorString(r), log.c_str());
    }
    size_t n = 0;
    nvrtcGetCUBINSize(prog, &n);
    cubin.resize(n);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    write_file_atomic(path, cubin.data(), cubin.size());
    return HB_OK;
}

int stepper_smem_bytes(const hb_stepper_spec &s, int block, int stages) {
    auto up128 = [](int x) { return (x + 127) / 128 * 128; };
    int warps = block / 32;
    const int npw = s.nodes_per_warp == 16 ? 16 : 32;
    // per-thread MLP path: parameters in the module's constant bank; tensor path: staged in shared memory + scratch
    int nn = s.nn_tensor ? (s.precision ? up128(s.nn_tc_bytes) : up128(s.nn_words * 8) + warps * 1024) : 0;
    const int es = s.precision ? 4 : 8;
    int in_b = up128(stages * npw * s.n_inputs * es) + up128(stages * npw * s.n_aux_inputs * es);
    int out_b = s.output_mode == HB_OUT_ROWS ? up128(2 * npw * s.n_rows * es) : 0;
    return nn + warps * (in_b + out_b) + warps * stages * 8;
}

int load_model(hb_model_s *m) {
    if (m->func) return HB_OK;
    int rc = ensure_device();
    if (rc) return rc;
    HB_DRV(g_drv.ModuleLoadData(&m->module, m->cubin.data()));
    HB_DRV(g_drv.ModuleGetFunction(&m->func, m->module, m->kernel_name.c_str()));
    // the 48 KB default limit counts static + dynamic shared memory (the templates hold a 2 KB exp table and, for bodies with
    // powers, a 2 KB log table statically): opt in
    // whenever the sum comes near it, not only when the dynamic part alone exceeds it
    if (m->dyn_smem + 6144 > 48 * 1024)
        HB_DRV(g_drv.FuncSetAttribute(m->func, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, m->dyn_smem));
    if (m->kind == 3 && m->aspec.nn_words > 0) {
        size_t bytes = 0;
        HB_DRV(g_drv.ModuleGetGlobal(&m->nn_const, &bytes, m->module, "hb_nn_const"));
        if (bytes != (size_t)m->aspec.nn_words * 8) return fail(HB_ERR_CUDA, "hb_nn_const has %zu bytes, expected %d", bytes, m->aspec.nn_words * 8);
    }
    if (m->kind == 0 && m->spec.nn_words > 0) {
        size_t bytes = 0;
        HB_DRV(g_drv.ModuleGetGlobal(&m->nn_const, &bytes, m->module, "hb_nn_const"));
        const size_t want = (size_t)m->spec.nn_words * (m->spec.precision ? 4 : 8);
        if (bytes != want) return fail(HB_ERR_CUDA, "hb_nn_const has %zu bytes, expected %zu", bytes, want);
    }
    return HB_OK;
}

int ws_reserve(hb_model_s::Buf &b, size_t bytes) {
    if (bytes <= b.cap) return HB_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = bytes;
    return HB_OK;
}

int pin_reserve(hb_model_s::Buf &b, size_t bytes) {
    if (bytes <= b.cap) return HB_OK;
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMallocHost(&b.p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(HB_ERR_NOMEM, "cudaMallocHost(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = bytes;
    return HB_OK;
}

// true when the DMA engines cannot address `p` directly (ordinary malloc'd / Julia GC / numpy memory)
bool host_is_pageable(const void *p) {
    if (!p) return false;
    if (getenv("HB_NO_STAGING")) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Large host copy with streaming (non-temporal) stores: the destination is not read back by this thread, so skipping the
// read-for-ownership of a plain memcpy saves a third of the DRAM traffic (glibc only switches to NT stores above a
// threshold of tens of MB per call, larger than the pieces the pool hands out).
#if defined(__x86_64__)
__attribute__((target("avx2"))) void stream_copy_avx2(char *dst, const char *src, size_t n) {
    size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
    if (head > n) head = n;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t blocks = n / 128;
    for (size_t i = 0; i < blocks; ++i) {
        __m256i a = _mm256_loadu_si256((const __m256i *)(src + 0)), b = _mm256_loadu_si256((const __m256i *)(src + 32));
        __m256i c = _mm256_loadu_si256((const __m256i *)(src + 64)), d = _mm256_loadu_si256((const __m256i *)(src + 96));
        _mm256_stream_si256((__m256i *)(dst + 0), a);
        _mm256_stream_si256((__m256i *)(dst + 32), b);
        _mm256_stream_si256((__m256i *)(dst + 64), c);
        _mm256_stream_si256((__m256i *)(dst + 96), d);
        src += 128; dst += 128;
    }
    _mm_sfence();
    memcpy(dst, src, n - blocks * 128);
}
#endif
void stream_copy(char *dst, const char *src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("HB_NO_NT_COPY");
    if (avx2 && n >= 4096) { stream_copy_avx2(dst, src, n); return; }
#endif
    memcpy(dst, src, n);
}

// Host threads that move chunks between pageable caller arrays and the pinned staging ring.  One pool per
// process; size = HB_HOST_THREADS, default min(16, cores this process may run on).
class HostPool {
  public:
    void copy(char *dst, const char *src, size_t bytes) {
        const size_t piece = (size_t)4 << 20;
        if (bytes <= piece || n_ <= 1) { stream_copy(dst, src, bytes); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t o = 0; o < bytes; o += piece) q_.push_back({dst + o, src + o, bytes - o < piece ? bytes - o : piece});
            pending_ += (bytes + piece - 1) / piece;
        }
        cv_.notify_all();
        work(false);                       // the caller copies too
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
    explicit HostPool(int n) : n_(n) {
        for (int i = 1; i < n_; ++i) th_.emplace_back([this] { work(true); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int threads() const { return n_; }

  private:
    struct Job { char *dst; const char *src; size_t n; };
    void work(bool forever) {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (forever) cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                j = q_.front();
                q_.pop_front();
            }
            stream_copy(j.dst, j.src, j.n);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    size_t pending_ = 0;
    bool stop_ = false;
};

HostPool &host_pool() {
    static HostPool *pool = [] {
        int n = 0;
        if (const char *e = getenv("HB_HOST_THREADS")) n = atoi(e);
        if (n <= 0) {
            cpu_set_t set;
            CPU_ZERO(&set);
            n = sched_getaffinity(0, sizeof set, &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            if (n > 16) n = 16;
        }
        if (n < 1) n = 1;
        if (n > 64) n = 64;
        return new HostPool(n);            // lives for the process (worker threads are detached from static destruction order)
    }();
    return *pool;
}

int check_desc(const hb_model_s *m, const hb_run_desc *d) {
    const hb_stepper_spec &s = m->spec;
    if (!d || d->struct_size != (int32_t)sizeof(hb_run_desc)) return fail(HB_ERR_INVALID, "hb_run_desc.struct_size mismatch");
    if (d->n_nodes <= 0 || d->n_steps < 0) return fail(HB_ERR_INVALID, "n_nodes must be > 0 and n_steps >= 0");
    if (d->min_value != d->min_value) return fail(HB_ERR_INVALID, "min_value is NaN");   // any real floor: hydrosolve takes what the config carries (src/solve.jl:40); HydroConfig's own > 0 check is the host mirror's
    if (s.n_inputs > 0 && d->n_steps > 0 && !d->input) return fail(HB_ERR_INVALID, "input is NULL");
    if (s.n_params > 0 && (!d->params || !d->param_offset || !d->param_mode)) return fail(HB_ERR_INVALID, "params / param tables are NULL");
    for (int j = 0; j < s.n_params; ++j) {
        int mode = d->param_mode[j];
        int64_t off = d->param_offset[j];
        int64_t span = mode == HB_PARAM_SCALAR ? 1 : (mode == HB_PARAM_PER_NODE ? d->n_nodes : 1);
        if (mode < 0 || mode - HB_PARAM_GATHER >= d->n_htype_sets) return fail(HB_ERR_INVALID, "param_mode[%d]=%d refers to a missing htypes set", j, mode);
        if (mode >= HB_PARAM_GATHER && !d->htypes) return fail(HB_ERR_INVALID, "htypes is NULL but parameter %d gathers through it", j);
        if (off < 0 || off + span > d->n_param_values) return fail(HB_ERR_INVALID, "parameter %d addresses outside params (offset %lld)", j, (long long)off);
    }
    if (m->uhw_total > 0 && !d->uh_weights) return fail(HB_ERR_INVALID, "uh_weights is NULL");
    if (d->n_compact_rows < 0 || d->n_compact_rows > 4 || (d->n_compact_rows > 0 && s.output_mode != HB_OUT_ROWS))
        return fail(HB_ERR_INVALID, "n_compact_rows must be 0..4 (rows mode)");
    for (int k = 0; k < d->n_compact_rows; ++k)
        if (d->compact_rows[k] < 0 || d->compact_rows[k] >= s.n_rows) return fail(HB_ERR_INVALID, "compact row outside the record");
    if (s.nn_words > 0 && !d->nn_params) return fail(HB_ERR_INVALID, "nn_params is NULL");
    if (s.output_mode == HB_OUT_ROWS) {
        if (s.n_rows > 0 && d->n_steps > 0 && !d->out) return fail(HB_ERR_INVALID, "out is NULL");
    } else {
        if (!d->objective || !d->target) return fail(HB_ERR_INVALID, "objective mode needs target and objective buffers");
        if (d->warm_up < 1 || d->warm_up > d->n_steps) return fail(HB_ERR_INVALID, "warm_up must be in 1..n_steps");
        if (s.n_tangents > 0 && !d->gradient) return fail(HB_ERR_INVALID, "gradient run needs the gradient buffer");
    }
    return HB_OK;
}

// Launch with DEVICE pointers in `d` (param tables are host pointers).
int launch_stepper(hb_model_s *m, const hb_run_desc *d, cudaStream_t stream) {
    const hb_stepper_spec &s = m->spec;
    // HbArgs mirror (templates/stepper.cu): 13 pointers, 2 i64, 1 double, 4 + 5 int, then int tables
    unsigned char buf[188 + 8 * HB_MAX_PARAMS + 16];
    memset(buf, 0, sizeof buf);
    size_t o = 0;
    auto put = [&](const void *p, size_t n) { memcpy(buf + o, p, n); o += n; };
    // gradient runs have no record stream: the `out` slot carries the [D][N] gradient buffer
    const void *ptrs[13] = {d->input, d->params, d->htypes, d->initstates, d->uh_weights, d->uh_hist_in,
                            d->nn_params, d->target, s.n_tangents > 0 ? (void *)d->gradient : d->out, d->objective, d->final_states,
                            d->uh_hist_out, d->aux_input};
    put(ptrs, sizeof ptrs);
    int64_t N = d->n_nodes, T = d->n_steps;
    put(&N, 8);
    put(&T, 8);
    double minv = d->min_value;
    put(&minv, 8);
    int warm = d->warm_up, tpn = d->target_per_node;
    // TMA bulk copies need 16-byte aligned global addresses: base pointer and per-step stride
    const int es = s.precision ? 4 : 8;
    int tma_in = ((reinterpret_cast<uintptr_t>(d->input) & 15) == 0) && ((N * s.n_inputs * es) % 16 == 0);
    if (s.n_aux_inputs > 0 && d->aux_input && (((reinterpret_cast<uintptr_t>(d->aux_input) & 15) != 0) || ((N * s.n_aux_inputs * es) % 16 != 0))) tma_in = 0;
    int tma_out = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && ((N * s.n_rows * es) % 16 == 0);
    if (getenv("HB_DISABLE_TMA")) tma_in = tma_out = 0;
    put(&warm, 4);
    put(&tpn, 4);
    if (d->n_compact_rows > 0) tma_out = 0;
    put(&tma_in, 4);
    put(&tma_out, 4);
    int cn = d->n_compact_rows, crow[4] = {d->compact_rows[0], d->compact_rows[1], d->compact_rows[2], d->compact_rows[3]};
    put(&cn, 4);
    put(crow, 16);
    int np = s.n_params > 0 ? s.n_params : 1;
    std::vector<int> tmp(np, 0);
    for (int j = 0; j < s.n_params; ++j) tmp[j] = d->param_offset[j];
    put(tmp.data(), 4 * np);
    for (int j = 0; j < s.n_params; ++j) tmp[j] = d->param_mode[j];
    put(tmp.data(), 4 * np);
    void *params[1] = {buf};
    const int64_t nodes_per_block = (int64_t)(m->block / 32) * (s.nodes_per_warp == 16 ? 16 : 32);
    unsigned grid = (unsigned)((N + nodes_per_block - 1) / nodes_per_block);
    if (T == 0 && s.output_mode == HB_OUT_ROWS && !d->final_states) return HB_OK;
    if (s.nn_words > 0) {   // MLP parameters -> constant bank (stream-ordered before the launch), in the run's arithmetic type
        if (s.precision) {
            int rc = ws_reserve(m->ws[11], (size_t)s.nn_words * 4);
            if (rc) return rc;
            if ((rc = hb_convert_f64_f32_device(d->nn_params, (float *)m->ws[11].p, s.nn_words, stream))) return fail(rc, "hb_convert_f64_f32_device failed");
            HB_DRV(g_drv.MemcpyDtoDAsync(m->nn_const, (CUdeviceptr)(uintptr_t)m->ws[11].p, (size_t)s.nn_words * 4, (CUstream)stream));
        } else {
            HB_DRV(g_drv.MemcpyDtoDAsync(m->nn_const, (CUdeviceptr)(uintptr_t)d->nn_params, (size_t)s.nn_words * 8, (CUstream)stream));
        }
    }
    HB_DRV(g_drv.LaunchKernel(m->func, grid, 1, 1, m->block, 1, 1, m->dyn_smem, (CUstream)stream, params, nullptr));
    return HB_OK;
}

int launch_route_step(hb_model_s *m, const hb_route_desc *d, int64_t t, const double *s_prev, double *s_next,
                      const double *q_cur, double *q_next, cudaStream_t st) {
    const hb_route_spec &s = m->rspec;
    const int64_t N = d->n_nodes, T = d->n_steps;
    const int nv = s.n_inputs > 0 ? s.n_inputs : 1, np = s.n_params > 0 ? s.n_params : 1;
    // HbRouteArgs mirror (templates/route.cu): 11 pointers, 3 i64, 1 double, 3 int, int tables
    unsigned char buf[11 * 8 + 3 * 8 + 8 + 3 * 4 + 4 * (16 + 2 * HB_MAX_PARAMS) + 16];
    memset(buf, 0, sizeof buf);
    size_t o = 0;
    auto put = [&](const void *p, size_t n) { memcpy(buf + o, p, n); o += n; };
    const void *ptrs[11] = {d->records, d->params, d->htypes, s_prev, s_next, q_cur, q_next, d->up_ptr, d->up_idx, d->route_inputs, d->route_out};
    put(ptrs, sizeof ptrs);
    put(&N, 8);
    put(&t, 8);
    put(&T, 8);
    double minv = d->min_value;
    put(&minv, 8);
    int R = d->n_rows, sr = d->s_row, orow = d->o_row;
    put(&R, 4);
    put(&sr, 4);
    put(&orow, 4);
    std::vector<int> tmp(nv > np ? nv : np, 0);
    for (int v = 0; v < s.n_inputs; ++v) tmp[v] = d->in_row[v];
    put(tmp.data(), 4 * nv);
    for (int j = 0; j < np; ++j) tmp[j] = j < s.n_params ? d->param_offset[j] : 0;
    put(tmp.data(), 4 * np);
    for (int j = 0; j < np; ++j) tmp[j] = j < s.n_params ? d->param_mode[j] : 0;
    put(tmp.data(), 4 * np);
    void *params[1] = {buf};
    unsigned grid = (unsigned)((N + m->block - 1) / m->block);
    HB_DRV(g_drv.LaunchKernel(m->func, grid, 1, 1, m->block, 1, 1, 0, (CUstream)st, params, nullptr));
    return HB_OK;
}

int check_route_desc(const hb_model_s *m, const hb_route_desc *d) {
    if (!m || m->kind != 1) return fail(HB_ERR_INVALID, "not a route model");
    if (!d || d->struct_size != (int32_t)sizeof(hb_route_desc)) return fail(HB_ERR_INVALID, "hb_route_desc.struct_size mismatch");
    const hb_route_spec &s = m->rspec;
    const bool planes = d->route_inputs || d->route_out;
    if (planes && (!d->route_inputs || !d->route_out)) return fail(HB_ERR_INVALID, "planes mode needs both route_inputs and route_out");
    if (d->n_nodes <= 0 || d->n_steps < 0 || (!d->records && !planes) || !d->up_ptr || !d->up_idx) return fail(HB_ERR_INVALID, "bad route descriptor");
    if (planes) return HB_OK;
    if (d->min_value != d->min_value) return fail(HB_ERR_INVALID, "min_value is NaN");   // any real floor: hydrosolve takes what the config carries (src/solve.jl:40); HydroConfig's own > 0 check is the host mirror's
    if (d->s_row < 0 || d->s_row >= d->n_rows || d->o_row < 0 || d->o_row >= d->n_rows) return fail(HB_ERR_INVALID, "route rows outside the record");
    for (int v = 0; v < s.n_inputs; ++v)
        if (d->in_row[v] < 0 || d->in_row[v] >= d->n_rows) return fail(HB_ERR_INVALID, "route input row outside the record");
    return HB_OK;
}

}  // namespace

// ================================ C-ABI ==============================================================
extern "C" {

void hb_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }   // for the other translation units of the library

int hb_version(void) { return HB_ABI_VERSION; }
const char *hb_last_error(void) { return g_err.c_str(); }

int hb_device_count(int *count) {
    if (!count) return fail(HB_ERR_INVALID, "count is NULL");
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *count = c;
    return HB_OK;
}

int hb_init(int device) {
    int cnt = 0;
    hb_device_count(&cnt);
    if (cnt == 0) return fail(HB_ERR_NO_DEVICE, "no CUDA device available; hydrob200 has no CPU fallback");
    if (device < 0 || device >= cnt) return fail(HB_ERR_INVALID, "device %d out of range (0..%d)", device, cnt - 1);
    g_device = device;
    return ensure_device();
}

int hb_set_cache_dir(const char *path) {
    if (!path) return fail(HB_ERR_INVALID, "path is NULL");
    g_cache_dir = path;
    mkdir(path, 0755);
    return HB_OK;
}

int hb_compile_stepper(const hb_stepper_spec *spec, const char *body, hb_model_t *out) {
    if (!spec || !body || !out) return fail(HB_ERR_INVALID, "NULL argument");
    if (spec->struct_size != (int32_t)sizeof(hb_stepper_spec)) return fail(HB_ERR_INVALID, "hb_stepper_spec.struct_size mismatch");
    if (spec->n_inputs < 0 || spec->n_rows < 0 || spec->n_states < 0 || spec->n_params < 0 || spec->n_params > HB_MAX_PARAMS ||
        spec->n_uh < 0 || spec->n_uh > HB_MAX_UH || spec->nn_words < 0 || spec->n_aux_inputs < 0 || spec->n_aux_inputs > 16 ||
        (spec->n_aux_inputs > 0 && (spec->shared_forcing || spec->output_mode != HB_OUT_ROWS)))
        return fail(HB_ERR_INVALID, "hb_stepper_spec out of range");
    if (spec->output_mode != HB_OUT_ROWS && spec->output_mode != HB_OUT_OBJECTIVE) return fail(HB_ERR_INVALID, "bad output_mode");
    if (spec->precision != 0 && spec->precision != 1) return fail(HB_ERR_INVALID, "precision must be 0 (Float64) or 1 (Float32)");
    // Float32 + nn_tensor = the tcgen05 path (templates/hb_tc.cuh): the CTA's 128 threads are the M dimension of the MMA
    if (spec->precision && spec->nn_tensor && (spec->nn_tc_bytes <= 0 || (spec->block_threads != 0 && spec->block_threads != 128) ||
                                                 spec->nodes_per_warp == 16))
        return fail(HB_ERR_INVALID, "the Float32 tensor path needs nn_tc_bytes > 0, 128 threads per block and 32 nodes per warp");
    if (!spec->precision && spec->nn_tc_bytes != 0) return fail(HB_ERR_INVALID, "nn_tc_bytes is for the Float32 tensor path");
    if (spec->nodes_per_warp != 0 && spec->nodes_per_warp != 32 && spec->nodes_per_warp != 16)
        return fail(HB_ERR_INVALID, "nodes_per_warp must be 0, 16 or 32");
    if (spec->nodes_per_warp == 16 && !spec->nn_tensor) return fail(HB_ERR_INVALID, "nodes_per_warp = 16 is meant for tensor-path MLP bodies");
    if (spec->n_tangents < 0 || spec->n_tangents > HB_MAX_TANGENTS) return fail(HB_ERR_INVALID, "n_tangents must be in 0..%d", HB_MAX_TANGENTS);
    if (spec->n_tangents > 0 && (spec->output_mode != HB_OUT_OBJECTIVE || spec->precision || spec->nn_words > 0))
        return fail(HB_ERR_INVALID, "gradient runs need objective mode, Float64 and a body without MLPs");
    if (spec->metric < 0 || spec->metric > HB_METRIC_SSE) return fail(HB_ERR_INVALID, "unknown metric %d", spec->metric);
    hb_model_s *m = new hb_model_s();
    m->spec = *spec;
    // defaults (tools/r2_tune2.sh, profiles/round2_block_stages_sweep.txt): record-writing Float64 bodies without MLPs are HBM-bound and
    // run best as many small blocks with a deep forcing ring — 64 threads x 8 steps: ExpHydro 0.947, GR4J + UH 0.93 (128 x 4: 0.87), HBV
    // 0.93 (0.85); MLP bodies (tensor tiles of 128 basins per CTA) and objective runs keep 128 x 4
    const bool plain_rows = spec->output_mode != HB_OUT_OBJECTIVE && spec->nn_words == 0 && !spec->precision;
    m->block = spec->block_threads > 0 ? spec->block_threads : (plain_rows ? 64 : 128);
    if (m->block % 32 || m->block > 1024) { delete m; return fail(HB_ERR_INVALID, "block_threads must be a multiple of 32 <= 1024"); }
    int stages = spec->stages > 0 ? spec->stages : (plain_rows ? 8 : 4);
    for (int u = 0; u < spec->n_uh; ++u) {
        if (spec->uh_kmax[u] < 1) { delete m; return fail(HB_ERR_INVALID, "uh_kmax must be >= 1"); }
        m->uhw_total += spec->uh_kmax[u];
        m->uhh_total += spec->uh_kmax[u] - 1;
    }
    m->dyn_smem = stepper_smem_bytes(*spec, m->block, stages);
    // static shared memory of the templates: the 256-entry exp table (2 KB), the log table of bodies with powers (2 KB) + a few words
    const int kStaticSmem = 2048 + 64 + (strstr(body, "#define HB_LOG_TABLE 1") ? 2048 : 0);
    if (m->dyn_smem + kStaticSmem > 227 * 1024) {
        const int need = m->dyn_smem + kStaticSmem;
        delete m;
        return fail(HB_ERR_INVALID, "kernel needs %d bytes of shared memory (> 227 KB)", need);
    }
    // resident blocks per SM the shared-memory budget allows (register limit is the compiler's)
    int minblocks = (227 * 1024) / (m->dyn_smem + 1024);
    int cap = 2048 / m->block;
    if (minblocks > cap) minblocks = cap;
    if (minblocks < 1) minblocks = 1;
    if (spec->max_registers > 0) {      // register budget per thread -> resident blocks the compiler must allow
        int byregs = 65536 / (spec->max_registers * m->block);
        if (byregs < 1) byregs = 1;
        if (byregs < minblocks) minblocks = byregs;
    } else if (minblocks > 4) {
        minblocks = 4;                  // default: 128 regs/thread at 128 threads x 4 blocks; FP64-heavy bodies need them
    }
    char defs[1024];
    snprintf(defs, sizeof defs,
             "//@@HB:DEFINES\n#define HB_V %d\n#define HB_R %d\n#define HB_S %d\n#define HB_NP %d\n#define HB_UHW_TOTAL %d\n"
             "#define HB_UHH_TOTAL %d\n#define HB_NN_WORDS %d\n#define HB_MODE %d\n#define HB_METRIC %d\n#define HB_SHARED_FORCING %d\n"
             "#define HB_BLOCK %d\n#define HB_STAGES %d\n#define HB_MINBLOCKS %d\n#define HB_NN_MMA %d\n#define HB_F32 %d\n#define HB_NPW %d\n"
             "#define HB_TANGENTS %d\n#define HB_V2 %d\n",
             spec->n_inputs, spec->output_mode == HB_OUT_ROWS ? spec->n_rows : 0, spec->n_states, spec->n_params, m->uhw_total,
             m->uhh_total, spec->nn_words, spec->output_mode, spec->metric, spec->shared_forcing ? 1 : 0, m->block, stages, minblocks,
             (spec->nn_tensor && !spec->precision) ? 1 : 0, spec->precision ? 1 : 0, spec->nodes_per_warp == 16 ? 16 : 32, spec->n_tangents,
             spec->n_aux_inputs);
    std::string full_body = std::string(defs) + body + "\n//@@HB:MATH\n" + kMathHeader + "\n";   // body may append to DEFINES (e.g. HB_NC)
    if (spec->nn_tensor && spec->precision) full_body += std::string("//@@HB:TC\n") + kTcHeader + "\n";
    if (const char *x = getenv("HB_EXTRA_DEFINES")) full_body = std::string("//@@HB:DEFINES\n") + x + "\n" + full_body;   // tuning experiments
    m->source = splice(kStepperTemplate, full_body.c_str());
    m->kernel_name = "hb_stepper";
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    int rc = nvrtc_compile(m->source, opts, m->cubin, m->from_cache);
    if (rc) { delete m; return rc; }
    *out = m;
    return HB_OK;
}

int hb_model_source(hb_model_t m, const char **src, int64_t *len) {
    if (!m || !src) return fail(HB_ERR_INVALID, "NULL argument");
    *src = m->source.c_str();
    if (len) *len = (int64_t)m->source.size();
    return HB_OK;
}

int hb_model_cubin(hb_model_t m, const void **cubin, int64_t *len) {
    if (!m || !cubin) return fail(HB_ERR_INVALID, "NULL argument");
    *cubin = m->cubin.data();
    if (len) *len = (int64_t)m->cubin.size();
    return HB_OK;
}

int hb_model_info(hb_model_t m, hb_kernel_info *info) {
    if (!m || !info) return fail(HB_ERR_INVALID, "NULL argument");
    int rc = load_model(m);
    if (rc) return rc;
    memset(info, 0, sizeof *info);
    int v = 0;
    HB_DRV(g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, m->func));
    info->registers_per_thread = v;
    HB_DRV(g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, m->func));
    info->static_smem_bytes = v;
    HB_DRV(g_drv.FuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, m->func));
    info->local_bytes_per_thread = v;
    info->dynamic_smem_bytes = m->dyn_smem;
    info->block_threads = m->block;
    HB_DRV(g_drv.Occupancy(&v, m->func, m->block, m->dyn_smem));
    info->max_blocks_per_sm = v;
    info->from_cache = m->from_cache ? 1 : 0;
    info->cubin_bytes = (int32_t)m->cubin.size();
    return HB_OK;
}

int hb_free_model(hb_model_t m) {
    if (!m) return HB_OK;
    for (auto &b : m->ws)
        if (b.p) cudaFree(b.p);
    if (m->pin_in.p) cudaFreeHost(m->pin_in.p);
    if (m->pin_out.p) cudaFreeHost(m->pin_out.p);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->stream_h2d) cudaStreamDestroy(m->stream_h2d);
    if (m->stream_d2h) cudaStreamDestroy(m->stream_d2h);
    for (auto e : m->chunk_ev) cudaEventDestroy(e);
    if (m->module && g_drv.ModuleUnload) g_drv.ModuleUnload(m->module);
    delete m;
    return HB_OK;
}

int hb_run_device(hb_model_t m, const hb_run_desc *d, void *stream) {
    if (!m) return fail(HB_ERR_INVALID, "model is NULL");
    int rc = check_desc(m, d);
    if (rc) return rc;
    rc = load_model(m);
    if (rc) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    cudaStream_t st = (cudaStream_t)stream;
    if (d->elapsed_ms) {
        if (!m->ev0) { HB_CUDA(cudaEventCreate(&m->ev0)); HB_CUDA(cudaEventCreate(&m->ev1)); }
        HB_CUDA(cudaEventRecord(m->ev0, st));
    }
    rc = launch_stepper(m, d, st);
    if (rc) return rc;
    if (d->elapsed_ms) {
        HB_CUDA(cudaEventRecord(m->ev1, st));
        HB_CUDA(cudaEventSynchronize(m->ev1));
        HB_CUDA(cudaEventElapsedTime(d->elapsed_ms, m->ev0, m->ev1));
    }
    return HB_OK;
}

int hb_run(hb_model_t m, const hb_run_desc *d) {
    if (!m) return fail(HB_ERR_INVALID, "model is NULL");
    if (m->kind != 0) return fail(HB_ERR_INVALID, "not a stepper model");
    if (m->spec.n_aux_inputs > 0 || d->n_compact_rows > 0)
        return fail(HB_ERR_INVALID, "auxiliary inputs / compact rows are device-resident features (hb_run_device)");
    int rc = check_desc(m, d);
    if (rc) return rc;
    rc = load_model(m);
    if (rc) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    if (!m->stream) {
        HB_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        HB_CUDA(cudaStreamCreateWithFlags(&m->stream_h2d, cudaStreamNonBlocking));
        HB_CUDA(cudaStreamCreateWithFlags(&m->stream_d2h, cudaStreamNonBlocking));
        HB_CUDA(cudaEventCreate(&m->ev0));
        HB_CUDA(cudaEventCreate(&m->ev1));
    }
    const hb_stepper_spec &s = m->spec;
    const int64_t N = d->n_nodes, T = d->n_steps;
    const int64_t Nin = s.shared_forcing ? 1 : N;
    const bool rows_mode = s.output_mode == HB_OUT_ROWS;
    hb_run_desc dd = *d;
    // ---- small, time-independent inputs: one H2D each on the compute stream ----
    struct In { int slot; const void *host; size_t bytes; const void **dev; };
    const In ins[] = {
        {1, d->params, (size_t)d->n_param_values * 8, (const void **)&dd.params},
        {2, d->htypes, (size_t)d->n_htype_sets * N * 4, (const void **)&dd.htypes},
        {3, d->initstates, (size_t)s.n_states * N * 8, (const void **)&dd.initstates},
        {4, d->uh_weights, (size_t)m->uhw_total * N * 8, (const void **)&dd.uh_weights},
        {5, d->uh_hist_in, (size_t)m->uhh_total * N * 8, (const void **)&dd.uh_hist_in},
        {6, d->nn_params, (size_t)s.nn_words * 8, (const void **)&dd.nn_params},
        {7, d->target, !rows_mode ? (size_t)T * (d->target_per_node ? N : 1) * 8 : 0, (const void **)&dd.target},
    };
    for (const In &in : ins) {
        *in.dev = nullptr;
        if (!in.host || in.bytes == 0) continue;
        if ((rc = ws_reserve(m->ws[in.slot], in.bytes))) return rc;
        HB_CUDA(cudaMemcpyAsync(m->ws[in.slot].p, in.host, in.bytes, cudaMemcpyHostToDevice, m->stream));
        *in.dev = m->ws[in.slot].p;
    }
    const size_t es = s.precision ? 4 : 8;
    const size_t in_step = (size_t)Nin * s.n_inputs * es;         // forcing bytes per time step
    const size_t out_step = rows_mode ? (size_t)N * s.n_rows * es : 0;
    // ---- time chunks: H2D(c+1) | kernel(c) | D2H(c-1) overlap on three streams; the bucket states
    //      and UH history are carried between chunk launches in device ping-pong buffers (a chunked
    //      run is bit-identical to a one-shot run: states are clamped values, carried subexpressions
    //      are functions of the state).  Objective mode reduces over the whole series: one chunk.
    //      The device side is a RING of `nslots` chunk-sized forcing / record slots (not the whole
    //      series), so a run larger than HBM streams through, and the first call allocates ~1 GB
    //      instead of the full forcing + result.
    //      Pageable host arrays (a plain Julia Array / numpy array) cannot be DMA'd directly: their
    //      chunks go through a pinned staging ring, copied by the library's host threads while the
    //      DMA engines work on the neighbouring chunks.
    const bool in_pageable = rows_mode && in_step && T > 0 && host_is_pageable(d->input);
    const bool out_pageable = rows_mode && out_step && T > 0 && host_is_pageable(d->out);
    int64_t nchunk = 1;
    if (rows_mode && T > 1) {
        size_t per_chunk = (in_pageable || out_pageable) ? (size_t)64 << 20 : (size_t)256 << 20;   // bytes of results per chunk
        if (const char *e = getenv("HB_CHUNK_BYTES")) { long long v = atoll(e); if (v > 0) per_chunk = (size_t)v; }
        nchunk = (int64_t)((out_step * T + per_chunk - 1) / per_chunk);
        if (nchunk > T) nchunk = T;
        if (nchunk < 1) nchunk = 1;
        if (getenv("HB_NO_PIPELINE")) nchunk = 1;
    }
    int nslots = 3;
    if (const char *e = getenv("HB_RING_SLOTS")) { int v = atoi(e); if (v >= 2 && v <= 16) nslots = v; }
    if (nchunk < nslots) nslots = (int)nchunk;
    const int64_t cs_max = (T + nchunk - 1) / nchunk;             // longest chunk (chunks are balanced: T*c/nchunk)
    const size_t in_slot = in_step * (size_t)cs_max, out_slot = out_step * (size_t)cs_max;
    const bool stage_in = in_pageable && nchunk > 1, stage_out = out_pageable && nchunk > 1;
    if (in_step && (rc = ws_reserve(m->ws[0], in_slot * nslots))) return rc;
    if (rows_mode && (rc = ws_reserve(m->ws[8], out_slot * nslots))) return rc;
    if (stage_in && (rc = pin_reserve(m->pin_in, in_slot * nslots))) return rc;
    if (stage_out && (rc = pin_reserve(m->pin_out, out_slot * nslots))) return rc;
    if (!rows_mode && (rc = ws_reserve(m->ws[9], (size_t)N * 8))) return rc;
    const size_t grad_bytes = (size_t)s.n_tangents * N * 8;
    if (grad_bytes && (rc = ws_reserve(m->ws[10], grad_bytes))) return rc;
    const size_t st_bytes = (size_t)s.n_states * N * 8, uh_bytes = (size_t)m->uhh_total * N * 8;
    if (nchunk > 1) {
        if (st_bytes && ((rc = ws_reserve(m->ws[12], st_bytes)) || (rc = ws_reserve(m->ws[13], st_bytes)))) return rc;
        if (uh_bytes && ((rc = ws_reserve(m->ws[14], uh_bytes)) || (rc = ws_reserve(m->ws[15], uh_bytes)))) return rc;
    } else {
        if (d->final_states && st_bytes && (rc = ws_reserve(m->ws[12], st_bytes))) return rc;
        if (d->uh_hist_out && uh_bytes && (rc = ws_reserve(m->ws[14], uh_bytes))) return rc;
    }
    // per ring slot: forcing landed / kernel done (slot's forcing consumed, records complete) / records copied out
    while ((int)m->chunk_ev.size() < 3 * nslots) {
        cudaEvent_t e;
        HB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->chunk_ev.push_back(e);
    }
    auto ev_h2d = [&](int sl) { return m->chunk_ev[3 * sl]; };
    auto ev_krn = [&](int sl) { return m->chunk_ev[3 * sl + 1]; };
    auto ev_d2h = [&](int sl) { return m->chunk_ev[3 * sl + 2]; };
    HB_CUDA(cudaEventRecord(m->ev0, m->stream));
    // the copy streams must not start before the workspace of a previous call is idle
    HB_CUDA(cudaStreamWaitEvent(m->stream_h2d, m->ev0, 0));
    HB_CUDA(cudaStreamWaitEvent(m->stream_d2h, m->ev0, 0));
    const void *state_in = dd.initstates, *uhh_in = dd.uh_hist_in;
    int last_state_slot = -1, last_uh_slot = -1;
    auto chunk_t0 = [&](int64_t c) { return T * c / nchunk; };
    // drain chunk c of a staged (pageable) result: wait for its D2H into the pinned slot, copy it to the caller's array
    auto drain = [&](int64_t c) -> int {
        const int sl = (int)(c % nslots);
        const int64_t t0 = chunk_t0(c), t1 = chunk_t0(c + 1);
        if (t1 <= t0) return HB_OK;
        HB_CUDA(cudaEventSynchronize(ev_d2h(sl)));
        host_pool().copy((char *)d->out + out_step * t0, (const char *)m->pin_out.p + out_slot * sl, out_step * (t1 - t0));
        return HB_OK;
    };
    for (int64_t c = 0; c < nchunk; ++c) {
        const int64_t t0 = chunk_t0(c), t1 = chunk_t0(c + 1);
        const int sl = (int)(c % nslots);
        const bool reuse = c >= nslots;
        char *dev_in = (char *)m->ws[0].p + in_slot * sl;
        char *dev_out = rows_mode ? (char *)m->ws[8].p + out_slot * sl : nullptr;
        if (t1 > t0 && in_step) {
            const char *src = (const char *)d->input + in_step * t0;
            if (stage_in) {
                if (reuse) HB_CUDA(cudaEventSynchronize(ev_h2d(sl)));      // the pinned slot's previous chunk is on the device
                char *pin = (char *)m->pin_in.p + in_slot * sl;
                host_pool().copy(pin, src, in_step * (t1 - t0));
                src = pin;
            }
            if (reuse) HB_CUDA(cudaStreamWaitEvent(m->stream_h2d, ev_krn(sl), 0));   // the device slot's previous chunk is consumed
            HB_CUDA(cudaMemcpyAsync(dev_in, src, in_step * (t1 - t0), cudaMemcpyHostToDevice, m->stream_h2d));
        }
        HB_CUDA(cudaEventRecord(ev_h2d(sl), m->stream_h2d));
        HB_CUDA(cudaStreamWaitEvent(m->stream, ev_h2d(sl), 0));
        if (reuse && rows_mode) HB_CUDA(cudaStreamWaitEvent(m->stream, ev_d2h(sl), 0)); // the record slot's previous chunk left the device
        hb_run_desc dc = dd;
        dc.elapsed_ms = nullptr;
        dc.n_steps = t1 - t0;
        dc.input = dev_in;
        dc.out = dev_out;
        dc.objective = rows_mode ? nullptr : (double *)m->ws[9].p;
        dc.gradient = grad_bytes ? (double *)m->ws[10].p : nullptr;
        dc.initstates = (const double *)state_in;
        dc.uh_hist_in = (const double *)uhh_in;
        const bool last = (c == nchunk - 1);
        dc.final_states = nullptr;
        dc.uh_hist_out = nullptr;
        if (st_bytes && (!last || d->final_states)) {
            last_state_slot = 12 + (int)(c & 1);
            if (nchunk == 1) last_state_slot = 12;
            dc.final_states = (double *)m->ws[last_state_slot].p;
        }
        if (uh_bytes && (!last || d->uh_hist_out)) {
            last_uh_slot = 14 + (int)(c & 1);
            if (nchunk == 1) last_uh_slot = 14;
            dc.uh_hist_out = (double *)m->ws[last_uh_slot].p;
        }
        if ((rc = launch_stepper(m, &dc, m->stream))) return rc;
        state_in = dc.final_states;
        uhh_in = dc.uh_hist_out;
        HB_CUDA(cudaEventRecord(ev_krn(sl), m->stream));
        if (rows_mode) {
            HB_CUDA(cudaStreamWaitEvent(m->stream_d2h, ev_krn(sl), 0));
            if (t1 > t0) {
                char *dst = stage_out ? (char *)m->pin_out.p + out_slot * sl : (char *)d->out + out_step * t0;
                HB_CUDA(cudaMemcpyAsync(dst, dev_out, out_step * (t1 - t0), cudaMemcpyDeviceToHost, m->stream_d2h));
            }
            HB_CUDA(cudaEventRecord(ev_d2h(sl), m->stream_d2h));
            // staged result: the chunk nslots-1 behind has to leave its pinned slot before the next iteration reuses it
            if (stage_out && c >= nslots - 1 && (rc = drain(c - (nslots - 1)))) return rc;
        }
    }
    HB_CUDA(cudaEventRecord(m->ev1, m->stream));
    if (!rows_mode) HB_CUDA(cudaMemcpyAsync(d->objective, m->ws[9].p, (size_t)N * 8, cudaMemcpyDeviceToHost, m->stream));
    if (grad_bytes) HB_CUDA(cudaMemcpyAsync(d->gradient, m->ws[10].p, grad_bytes, cudaMemcpyDeviceToHost, m->stream));
    if (d->final_states && st_bytes && last_state_slot >= 0)
        HB_CUDA(cudaMemcpyAsync(d->final_states, m->ws[last_state_slot].p, st_bytes, cudaMemcpyDeviceToHost, m->stream));
    if (d->uh_hist_out && uh_bytes && last_uh_slot >= 0)
        HB_CUDA(cudaMemcpyAsync(d->uh_hist_out, m->ws[last_uh_slot].p, uh_bytes, cudaMemcpyDeviceToHost, m->stream));
    if (stage_out)
        for (int64_t c = nchunk - (nslots - 1) > 0 ? nchunk - (nslots - 1) : 0; c < nchunk; ++c)
            if ((rc = drain(c))) return rc;
    HB_CUDA(cudaStreamSynchronize(m->stream));
    HB_CUDA(cudaStreamSynchronize(m->stream_d2h));
    HB_CUDA(cudaStreamSynchronize(m->stream_h2d));
    // device time of the compute stream from first to last kernel (includes waits for forcing chunks)
    if (d->elapsed_ms) HB_CUDA(cudaEventElapsedTime(d->elapsed_ms, m->ev0, m->ev1));
    return HB_OK;
}

// ---- routing ---------------------------------------------------------------------------------------------
int hb_compile_route(const hb_route_spec *spec, const char *body, hb_model_t *out) {
    if (!spec || !body || !out) return fail(HB_ERR_INVALID, "NULL argument");
    if (spec->struct_size != (int32_t)sizeof(hb_route_spec)) return fail(HB_ERR_INVALID, "hb_route_spec.struct_size mismatch");
    if (spec->n_inputs < 0 || spec->n_inputs > 16 || spec->n_params < 0 || spec->n_params > HB_MAX_PARAMS)
        return fail(HB_ERR_INVALID, "hb_route_spec out of range");
    hb_model_s *m = new hb_model_s();
    m->kind = 1;
    m->rspec = *spec;
    m->block = spec->block_threads > 0 ? spec->block_threads : 256;
    m->dyn_smem = 0;
    char defs[256];
    snprintf(defs, sizeof defs, "//@@HB:DEFINES\n#define HB_RV %d\n#define HB_NP %d\n#define HB_BLOCK %d\n", spec->n_inputs,
             spec->n_params, m->block);
    std::string full_body = std::string(defs) + body + "\n//@@HB:MATH\n" + kMathHeader + "\n";
    m->source = splice(kRouteTemplate, full_body.c_str());
    m->kernel_name = "hb_route_step";
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    int rc = nvrtc_compile(m->source, opts, m->cubin, m->from_cache);
    if (rc) { delete m; return rc; }
    *out = m;
    return HB_OK;
}

int hb_run_route_device(hb_model_t m, const hb_route_desc *d, void *stream) {
    int rc0 = check_route_desc(m, d);
    if (rc0) return rc0;
    int rc = load_model(m);
    if (rc) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = d->n_nodes, T = d->n_steps;
    // ping-pong state / outflow buffers (library workspace)
    for (int k = 0; k < 4; ++k)
        if ((rc = ws_reserve(m->ws[k], (size_t)N * 8))) return rc;
    double *sbuf[2] = {(double *)m->ws[0].p, (double *)m->ws[1].p};
    double *qbuf[2] = {(double *)m->ws[2].p, (double *)m->ws[3].p};
    if (d->initstates) HB_CUDA(cudaMemcpyAsync(sbuf[0], d->initstates, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
    else HB_CUDA(cudaMemsetAsync(sbuf[0], 0, (size_t)N * 8, st));
    if (d->elapsed_ms) {
        if (!m->ev0) { HB_CUDA(cudaEventCreate(&m->ev0)); HB_CUDA(cudaEventCreate(&m->ev1)); }
        HB_CUDA(cudaEventRecord(m->ev0, st));
    }
    for (int64_t t = -1; t < T; ++t) {
        // step t >= 0 reads s/q half (t & 1), written by launch t-1, and writes the other half;
        // the prologue launch (t = -1) reads the initial state and writes the outflows of step 0
        const int cur = t < 0 ? 0 : (int)(t & 1);
        if (t < 0) rc = launch_route_step(m, d, t, sbuf[0], sbuf[1], qbuf[1], qbuf[0], st);
        else rc = launch_route_step(m, d, t, sbuf[cur], sbuf[cur ^ 1], qbuf[cur], qbuf[cur ^ 1], st);
        if (rc) return rc;
    }
    // launch t = T-1 wrote the state half ((T-1) & 1) ^ 1 == T & 1
    if (d->final_states)
        HB_CUDA(cudaMemcpyAsync(d->final_states, sbuf[(int)(T & 1)], (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
    if (d->elapsed_ms) {
        HB_CUDA(cudaEventRecord(m->ev1, st));
        HB_CUDA(cudaEventSynchronize(m->ev1));
        HB_CUDA(cudaEventElapsedTime(d->elapsed_ms, m->ev0, m->ev1));
    }
    return HB_OK;
}

int hb_route_step_device(hb_model_t m, const hb_route_desc *d, int64_t t, const double *s_prev, double *s_next,
                         const double *q_cur, double *q_next, void *stream) {
    int rc = check_route_desc(m, d);
    if (rc) return rc;
    if (t < -1 || t >= d->n_steps || !s_prev || !s_next || !q_cur || !q_next) return fail(HB_ERR_INVALID, "bad route step arguments");
    if ((rc = load_model(m))) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    return launch_route_step(m, d, t, s_prev, s_next, q_cur, q_next, (cudaStream_t)stream);
}

// ---- reverse-mode gradient (templates/adjoint.cu) ----------------------------------------------------------------
int hb_compile_adjoint(const hb_adjoint_spec *spec, const char *body, hb_model_t *out) {
    if (!spec || !body || !out) return fail(HB_ERR_INVALID, "NULL argument");
    if (spec->struct_size != (int32_t)sizeof(hb_adjoint_spec)) return fail(HB_ERR_INVALID, "hb_adjoint_spec.struct_size mismatch");
    if (spec->n_inputs < 0 || spec->n_states < 0 || spec->n_params < 0 || spec->n_params > HB_MAX_PARAMS || spec->nn_words < 0 ||
        spec->acc_doubles_per_warp < 0)
        return fail(HB_ERR_INVALID, "hb_adjoint_spec out of range");
    hb_model_s *m = new hb_model_s();
    m->kind = 3;
    m->aspec = *spec;
    m->block = spec->block_threads > 0 ? spec->block_threads : 128;
    if (m->block % 32 || m->block > 1024) { delete m; return fail(HB_ERR_INVALID, "block_threads must be a multiple of 32 <= 1024"); }
    // per warp: transposition scratch (2 x 32 x 20 doubles) + the accumulator fragments of every dense layer
    m->dyn_smem = (m->block / 32) * (2 * 32 * 17 + spec->acc_doubles_per_warp) * 8;      // scratch pitch: HB_OP in templates/adjoint.cu
    if (m->dyn_smem + 4096 > 227 * 1024) { delete m; return fail(HB_ERR_INVALID, "adjoint kernel needs too much shared memory; use a smaller block"); }
    // resident blocks per SM the shared memory allows (2 KB exp table, 2 KB log table, 1 KB system per block), at most 3: the
    // kernel is latency-bound, so three blocks of 168 registers beat two of 255 (M50 50 k nodes: one wave instead of 1.32)
    int minblocks = (227 * 1024) / (m->dyn_smem + 5 * 1024);
    if (minblocks > 3) minblocks = 3;
    if (minblocks < 1) minblocks = 1;
    if (const char *e = getenv("HB_ADJ_MINBLOCKS")) { int v = atoi(e); if (v >= 1 && v <= 4) minblocks = v; }
    char defs[512];
    snprintf(defs, sizeof defs, "//@@HB:DEFINES\n#define HB_V %d\n#define HB_S %d\n#define HB_NP %d\n#define HB_NN_WORDS %d\n#define HB_BLOCK %d\n"
             "#define HB_UHW_TOTAL 0\n#define HB_UHH_TOTAL 0\n#define HB_ADJ_MINBLOCKS %d\n",
             spec->n_inputs, spec->n_states, spec->n_params, spec->nn_words, m->block, minblocks);
    std::string full_body = std::string(defs) + body + "\n//@@HB:MATH\n" + kMathHeader + "\n";
    m->source = splice(kAdjointTemplate, full_body.c_str());
    m->kernel_name = "hb_adjoint";
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    int rc = nvrtc_compile(m->source, opts, m->cubin, m->from_cache);
    if (rc) { delete m; return rc; }
    *out = m;
    return HB_OK;
}

int hb_run_adjoint_device(hb_model_t m, const hb_adjoint_desc *d, void *stream) {
    if (!m || m->kind != 3) return fail(HB_ERR_INVALID, "not an adjoint model");
    if (!d || d->struct_size != (int32_t)sizeof(hb_adjoint_desc)) return fail(HB_ERR_INVALID, "hb_adjoint_desc.struct_size mismatch");
    const hb_adjoint_spec &s = m->aspec;
    if (d->n_nodes <= 0 || d->n_steps < 0 || !d->scale || !d->grad_params || !d->grad_init || !d->loss ||
        (s.n_inputs > 0 && !d->input) || (s.n_states > 0 && d->n_steps > 1 && !d->trajectory) || (s.nn_words > 0 && (!d->nn_params || !d->grad_nn)))
        return fail(HB_ERR_INVALID, "bad adjoint descriptor");
    if (d->warm_up < 1) return fail(HB_ERR_INVALID, "warm_up must be >= 1");
    if (s.n_params > 0 && (!d->params || !d->param_offset || !d->param_mode)) return fail(HB_ERR_INVALID, "params / param tables are NULL");
    int rc = load_model(m);
    if (rc) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    cudaStream_t st = (cudaStream_t)stream;
    // HbAdjArgs mirror (templates/adjoint.cu): 11 pointers, 2 i64, 1 double, 2 int, int tables
    unsigned char buf[11 * 8 + 16 + 8 + 8 + 8 * HB_MAX_PARAMS + 16];
    memset(buf, 0, sizeof buf);
    size_t o = 0;
    auto put = [&](const void *p, size_t n) { memcpy(buf + o, p, n); o += n; };
    const void *ptrs[11] = {d->input, d->trajectory, d->params, d->htypes, d->initstates, d->target, d->scale,
                            d->grad_params, d->grad_init, d->grad_nn, d->loss};
    put(ptrs, sizeof ptrs);
    int64_t N = d->n_nodes, T = d->n_steps;
    put(&N, 8);
    put(&T, 8);
    double minv = d->min_value;
    put(&minv, 8);
    int warm = d->warm_up, tpn = d->target_per_node;
    put(&warm, 4);
    put(&tpn, 4);
    const int np = s.n_params > 0 ? s.n_params : 1;
    std::vector<int> tmp(np, 0);
    for (int j = 0; j < s.n_params; ++j) tmp[j] = d->param_offset[j];
    put(tmp.data(), 4 * np);
    for (int j = 0; j < s.n_params; ++j) tmp[j] = d->param_mode[j];
    put(tmp.data(), 4 * np);
    void *params[1] = {buf};
    if (d->elapsed_ms) {
        if (!m->ev0) { HB_CUDA(cudaEventCreate(&m->ev0)); HB_CUDA(cudaEventCreate(&m->ev1)); }
        HB_CUDA(cudaEventRecord(m->ev0, st));
    }
    if (s.nn_words > 0) {
        HB_CUDA(cudaMemsetAsync(d->grad_nn, 0, (size_t)s.nn_words * 8, st));
        HB_DRV(g_drv.MemcpyDtoDAsync(m->nn_const, (CUdeviceptr)(uintptr_t)d->nn_params, (size_t)s.nn_words * 8, (CUstream)st));
    }
    const unsigned grid = (unsigned)((N + m->block - 1) / m->block);
    HB_DRV(g_drv.LaunchKernel(m->func, grid, 1, 1, m->block, 1, 1, m->dyn_smem, (CUstream)st, params, nullptr));
    if (d->elapsed_ms) {
        HB_CUDA(cudaEventRecord(m->ev1, st));
        HB_CUDA(cudaEventSynchronize(m->ev1));
        HB_CUDA(cudaEventElapsedTime(d->elapsed_ms, m->ev0, m->ev1));
    }
    return HB_OK;
}

// ---- fused dense-grid model -------------------------------------------------------------------------------
int hb_compile_grid(const hb_grid_spec *spec, const char *body, hb_model_t *out) {
    if (!spec || !body || !out) return fail(HB_ERR_INVALID, "NULL argument");
    if (spec->struct_size != (int32_t)sizeof(hb_grid_spec)) return fail(HB_ERR_INVALID, "hb_grid_spec.struct_size mismatch");
    if (spec->n_params < 0 || spec->n_params > HB_MAX_PARAMS || spec->n_params_stepper > spec->n_params || spec->n_route_inputs < 1 ||
        spec->n_route_inputs > 16 || spec->n_rows < 2 || spec->s_row < 0 || spec->s_row >= spec->n_rows || spec->o_row < 0 ||
        spec->o_row >= spec->n_rows)
        return fail(HB_ERR_INVALID, "hb_grid_spec out of range");
    const int tc = spec->steps_per_chunk, bw = spec->block_w, bh = spec->block_h;
    hb_model_s *m = new hb_model_s();
    m->kind = 2;
    m->gspec = *spec;
    char defs[1024];
    if (spec->variant == HB_GRID_WAVE) {
        m->block = bw > 0 ? bw : 128;
        const int stages = spec->stages > 0 ? spec->stages : 3;
        if (tc < 0 || tc > 4096 || m->block % 32 || m->block > 1024 || stages > 16 || spec->n_inputs < 1) {
            delete m;
            return fail(HB_ERR_INVALID, "bad wave configuration (steps_per_chunk %d, block %d, stages %d)", tc, bw, stages);
        }
        auto up128 = [](int x) { return (x + 127) / 128 * 128; };
        const int warps = m->block / 32;
        const int obuf = getenv("HB_WAVE_OBUF") ? atoi(getenv("HB_WAVE_OBUF")) : 3;     // tuning: record tiles per warp (2 or 3)
        m->wave_qslots = getenv("HB_WAVE_QSLOTS") ? atoi(getenv("HB_WAVE_QSLOTS")) : 2; // tuning: outflow words per cell (2..8)
        if (m->wave_qslots < 2 || m->wave_qslots > 8) { delete m; return fail(HB_ERR_INVALID, "HB_WAVE_QSLOTS must be 2..8"); }
        if (obuf < 2 || obuf > 4) { delete m; return fail(HB_ERR_INVALID, "HB_WAVE_OBUF must be 2..4"); }
        m->dyn_smem = warps * (up128(stages * 32 * spec->n_inputs * 8) + up128(obuf * 32 * spec->n_rows * 8)) + warps * stages * 8;
        const int static_smem = 2048 + 64 + 64 + (strstr(body, "#define HB_LOG_TABLE 1") ? 2048 : 0);
        if (m->dyn_smem + static_smem > 227 * 1024) {
            const int need = m->dyn_smem + static_smem;
            delete m;
            return fail(HB_ERR_INVALID, "wave kernel needs %d bytes of shared memory", need);
        }
        int minblocks = (227 * 1024) / (m->dyn_smem + 1024 + 1024);
        const int cap = 2048 / m->block;
        if (minblocks > cap) minblocks = cap;
        if (minblocks < 1) minblocks = 1;
        if (spec->max_registers > 0) {
            int byregs = 65536 / (spec->max_registers * m->block);
            if (byregs < 1) byregs = 1;
            if (byregs < minblocks) minblocks = byregs;
        } else if (minblocks > 4) {
            minblocks = 4;
        }
        snprintf(defs, sizeof defs,
                 "//@@HB:DEFINES\n#define HB_V %d\n#define HB_R %d\n#define HB_S %d\n#define HB_NP %d\n#define HB_NPS %d\n#define HB_RV %d\n"
                 "#define HB_SROW %d\n#define HB_OROW %d\n#define HB_BLOCK %d\n#define HB_STAGES %d\n#define HB_MINBLOCKS %d\n"
                 "#define HB_UHW_TOTAL 0\n#define HB_UHH_TOTAL 0\n#define HB_NN_WORDS 0\n#define HB_OBUF %d\n#define HB_QSLOTS %d\n",
                 spec->n_inputs, spec->n_rows, spec->n_states, spec->n_params, spec->n_params_stepper, spec->n_route_inputs, spec->s_row,
                 spec->o_row, m->block, stages, minblocks, obuf, m->wave_qslots);
        m->kernel_name = "hb_grid_wave";
    } else if (spec->variant == HB_GRID_TILE) {
        if (tc < 1 || tc > 8 || bw % 32 || bw < 32 || bh < 1 || bw * bh > 1024 || bw - 2 * tc < 1 || bh - 2 * tc < 1) {
            delete m;
            return fail(HB_ERR_INVALID, "bad grid tiling (steps_per_chunk %d, block %d x %d)", tc, bw, bh);
        }
        m->block = bw * bh;
        m->dyn_smem = m->block * 8 + (m->block / 32) * 32 * spec->n_rows * 8 + m->block + 64;
        snprintf(defs, sizeof defs,
                 "//@@HB:DEFINES\n#define HB_V %d\n#define HB_R %d\n#define HB_S %d\n#define HB_NP %d\n#define HB_NPS %d\n#define HB_RV %d\n"
                 "#define HB_SROW %d\n#define HB_OROW %d\n#define HB_TC %d\n#define HB_BW %d\n#define HB_BH %d\n#define HB_UHW_TOTAL 0\n"
                 "#define HB_UHH_TOTAL 0\n#define HB_NN_WORDS 0\n",
                 spec->n_inputs, spec->n_rows, spec->n_states, spec->n_params, spec->n_params_stepper, spec->n_route_inputs, spec->s_row,
                 spec->o_row, tc, bw, bh);
        m->kernel_name = "hb_grid_chunk";
    } else {
        delete m;
        return fail(HB_ERR_INVALID, "unknown grid kernel variant %d", spec->variant);
    }
    std::string full_body = std::string(defs) + body + "\n//@@HB:MATH\n" + kMathHeader + "\n";
    if (const char *x = getenv("HB_EXTRA_DEFINES")) full_body = std::string("//@@HB:DEFINES\n") + x + "\n" + full_body;   // tuning experiments
    m->source = splice(spec->variant == HB_GRID_WAVE ? kGridWaveTemplate : kGridTemplate, full_body.c_str());
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=true"};
    int rc = nvrtc_compile(m->source, opts, m->cubin, m->from_cache);
    if (rc) { delete m; return rc; }
    *out = m;
    return HB_OK;
}

namespace {

// WAVE variant: T_c is limited by forward progress — a warp waits on blocks up to `band` logical indices
// ahead, recursively one step shallower each hop, so T_c * band blocks must be co-resident at once.
int wave_steps_per_launch(hb_model_s *m, int cols, int *out_tc) {
    int per_sm = 0, sms = 0;
    HB_DRV(g_drv.Occupancy(&per_sm, m->func, m->block, m->dyn_smem));
    HB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device));
    const int resident = per_sm * sms;
    const int band = (cols + m->block - 1) / m->block + 2;
    int tc = m->gspec.steps_per_chunk;
    const int fit = (resident / 2) / band;
    if (fit < 1) return fail(HB_ERR_INVALID, "grid too wide for the wave kernel (%d columns, %d co-resident blocks)", cols, resident);
    if (tc <= 0) {
        // automatic: the rows of the last T_c steps behind the dispatch frontier cannot run freely (each waits for the row
        // below to be dispatched), so T_c rows should be a small share of the co-resident rows — about a quarter measured
        // best (2048 columns: 8 steps per launch, 1024: 18, 512 and narrower: 32), while few steps per launch pay the
        // per-launch prologue more often
        const long long resident_cells = (long long)resident * m->block;
        tc = (int)(resident_cells / (4LL * cols));
        if (tc < 4) tc = 4;
        if (tc > 32) tc = 32;
    }
    if (tc > fit) tc = fit;
    *out_tc = tc;
    return HB_OK;
}

}  // namespace

int hb_run_grid_device(hb_model_t m, const hb_grid_desc *d, void *stream) {
    if (!m || m->kind != 2) return fail(HB_ERR_INVALID, "not a grid model");
    if (!d || d->struct_size != (int32_t)sizeof(hb_grid_desc)) return fail(HB_ERR_INVALID, "hb_grid_desc.struct_size mismatch");
    const hb_grid_spec &s = m->gspec;
    if (d->rows <= 0 || d->cols <= 0 || d->n_steps < 0 || !d->input || !d->out || !d->flwdir) return fail(HB_ERR_INVALID, "bad grid descriptor");
    const bool window = d->out_rows > 0 || d->out_cols > 0;
    const int o_r0 = d->out_rows > 0 ? d->out_row0 : 0, o_nr = d->out_rows > 0 ? d->out_rows : d->rows;
    const int o_c0 = d->out_cols > 0 ? d->out_col0 : 0, o_nc = d->out_cols > 0 ? d->out_cols : d->cols;
    if (window && (s.variant != HB_GRID_WAVE || o_r0 < 0 || o_c0 < 0 || o_r0 + o_nr > d->rows || o_c0 + o_nc > d->cols))
        return fail(HB_ERR_INVALID, "an output window needs the wave kernel and must lie inside the grid");
    if (d->min_value != d->min_value) return fail(HB_ERR_INVALID, "min_value is NaN");   // any real floor: hydrosolve takes what the config carries (src/solve.jl:40); HydroConfig's own > 0 check is the host mirror's
    if (s.n_params > 0 && (!d->params || !d->param_offset || !d->param_mode)) return fail(HB_ERR_INVALID, "params / param tables are NULL");
    int rc = load_model(m);
    if (rc) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = (int64_t)d->rows * d->cols, T = d->n_steps;
    const size_t sb = (size_t)(s.n_states > 0 ? s.n_states : 1) * N * 8;
    for (int k = 0; k < 2; ++k) {
        if ((rc = ws_reserve(m->ws[k], sb))) return rc;
        if ((rc = ws_reserve(m->ws[2 + k], (size_t)N * 8))) return rc;
    }
    const bool wave = s.variant == HB_GRID_WAVE;
    int tc = s.steps_per_chunk;
    if (wave) {
        if ((rc = wave_steps_per_launch(m, d->cols, &tc))) return rc;
        m->wave_tc = tc;
        const size_t qslots = (size_t)m->wave_qslots;
        if ((rc = ws_reserve(m->ws[4], qslots * N * 16))) return rc;          // {outflow, step tag} words, ring over the step index
        if ((rc = ws_reserve(m->ws[6], 256))) return rc;
        HB_CUDA(cudaMemsetAsync(m->ws[4].p, 0, qslots * N * 16, st));         // tags = 0: nothing published yet
        HB_CUDA(cudaMemsetAsync(m->ws[6].p, 0, 4, st));                       // ticket (the error word [1] is sticky)
        if (!m->wave_err_init) { HB_CUDA(cudaMemsetAsync((char *)m->ws[6].p + 4, 0, 4, st)); m->wave_err_init = true; }
    }
    if (d->elapsed_ms) {
        if (!m->ev0) { HB_CUDA(cudaEventCreate(&m->ev0)); HB_CUDA(cudaEventCreate(&m->ev1)); }
        HB_CUDA(cudaEventRecord(m->ev0, st));
    }
    unsigned gx, gy;
    if (wave) {
        gx = (unsigned)((N + m->block - 1) / m->block);
        gy = 1;
    } else {
        gx = (unsigned)((d->cols + (s.block_w - 2 * tc) - 1) / (s.block_w - 2 * tc));
        gy = (unsigned)((d->rows + (s.block_h - 2 * tc) - 1) / (s.block_h - 2 * tc));
    }
    const void *su_in = d->init_bucket, *rs_in = d->init_route;
    int half = 0;
    const int np = s.n_params > 0 ? s.n_params : 1;
    uint32_t ticket_base = 0;
    for (int64_t t0 = 0; t0 < T; t0 += tc) {
        // HbGridArgs / HbWaveArgs mirrors (templates/grid.cu, templates/gridwave.cu)
        unsigned char buf[12 * 8 + 4 * 8 + 32 + 8 + 32 + 8 * HB_MAX_PARAMS + 16];
        memset(buf, 0, sizeof buf);
        size_t o = 0;
        auto put = [&](const void *p, size_t n) { memcpy(buf + o, p, n); o += n; };
        void *su_out = m->ws[half].p, *rs_out = m->ws[2 + half].p;
        const void *ptrs[9] = {d->input, d->out, d->params, d->htypes, su_in, su_out, rs_in, rs_out, d->flwdir};
        put(ptrs, sizeof ptrs);
        int ns = (int)((T - t0) < tc ? (T - t0) : tc), rows = d->rows, cols = d->cols, pad = 0;
        double minv = d->min_value;
        if (wave) {
            const void *wp[3] = {m->ws[4].p, m->ws[5].p, m->ws[6].p};
            put(wp, sizeof wp);
            put(&N, 8);
            put(&t0, 8);
            put(&o_r0, 4);
            put(&o_nr, 4);
            put(&o_c0, 4);
            put(&o_nc, 4);
            // TMA bulk copies need 16-byte aligned global addresses: base pointers and per-step strides
            const int64_t in_ts = d->in_t_stride > 0 ? d->in_t_stride : N, out_ts = d->out_t_stride > 0 ? d->out_t_stride : (int64_t)o_nr * o_nc;
            const int in_pitch = d->in_pitch > 0 ? d->in_pitch : d->cols, in_col0 = d->in_pitch > 0 ? d->in_col0 : 0;
            const int out_pitch = d->out_pitch > 0 ? d->out_pitch : o_nc, out_base = d->out_pitch > 0 ? d->out_base : 0;
            int tma_in = ((reinterpret_cast<uintptr_t>(d->input) & 15) == 0) && ((in_ts * s.n_inputs * 8) % 16 == 0);
            int tma_out = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && ((out_ts * s.n_rows * 8) % 16 == 0);
            if (getenv("HB_DISABLE_TMA")) tma_in = tma_out = 0;
            if (getenv("HB_DISABLE_TMA_IN")) tma_in = 0;
            if (getenv("HB_DISABLE_TMA_OUT")) tma_out = 0;
            put(&ticket_base, 4);
            put(&ns, 4);
            put(&rows, 4);
            put(&cols, 4);
            put(&tma_in, 4);
            put(&tma_out, 4);
            put(&minv, 8);
            put(&in_ts, 8);
            put(&out_ts, 8);
            put(&in_pitch, 4);
            put(&in_col0, 4);
            put(&out_pitch, 4);
            put(&out_base, 4);
            ticket_base += gx;
        } else {
            put(&N, 8);
            put(&t0, 8);
            put(&ns, 4);
            put(&rows, 4);
            put(&cols, 4);
            put(&pad, 4);
            put(&minv, 8);
        }
        std::vector<int> tmp(np, 0);
        for (int j = 0; j < s.n_params; ++j) tmp[j] = d->param_offset[j];
        put(tmp.data(), 4 * np);
        for (int j = 0; j < s.n_params; ++j) tmp[j] = d->param_mode[j];
        put(tmp.data(), 4 * np);
        void *params[1] = {buf};
        HB_DRV(g_drv.LaunchKernel(m->func, gx, gy, 1, m->block, 1, 1, m->dyn_smem, (CUstream)st, params, nullptr));
        su_in = su_out;
        rs_in = rs_out;
        half ^= 1;
    }
    if (T > 0) {
        if (d->final_bucket && s.n_states > 0) HB_CUDA(cudaMemcpyAsync(d->final_bucket, su_in, (size_t)s.n_states * N * 8, cudaMemcpyDeviceToDevice, st));
        if (d->final_route) HB_CUDA(cudaMemcpyAsync(d->final_route, rs_in, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (d->elapsed_ms) {
        HB_CUDA(cudaEventRecord(m->ev1, st));
        HB_CUDA(cudaEventSynchronize(m->ev1));
        HB_CUDA(cudaEventElapsedTime(d->elapsed_ms, m->ev0, m->ev1));
        if (wave) return hb_grid_status(m, stream, nullptr);     // a timed run is synchronous anyway: report an expired poll here
    }
    return HB_OK;
}

int hb_grid_status(hb_model_t m, void *stream, int32_t *steps_per_launch) {
    if (!m || m->kind != 2) return fail(HB_ERR_INVALID, "not a grid model");
    if (steps_per_launch) *steps_per_launch = m->gspec.variant == HB_GRID_WAVE ? m->wave_tc : m->gspec.steps_per_chunk;
    if (m->gspec.variant != HB_GRID_WAVE || !m->ws[6].p) return HB_OK;
    HB_CUDA(cudaSetDevice(g_device));
    HB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    uint32_t err = 0;
    HB_CUDA(cudaMemcpy(&err, (char *)m->ws[6].p + 4, 4, cudaMemcpyDeviceToHost));
    if (err) {
        cudaMemset((char *)m->ws[6].p + 4, 0, 4);
        return fail(HB_ERR_CUDA, "hb_grid_wave: a warp gave up waiting for a neighbouring warp's outflows (forward-progress failure); results are invalid");
    }
    return HB_OK;
}


// ---- host <-> device link probe: the roofline of the end-to-end (host buffer) path ---------------------------
// Plain cudaMemcpyAsync of `bytes` between pinned host memory and the device: H2D alone, D2H alone, and both
// directions at once on two streams (what hb_run's pipeline asks of the link).  gbs[0..3] = H2D, D2H, H2D while
// D2H runs, D2H while H2D runs (GB/s, mean over `reps` copies).  bench.py runs it on all ranks at the same time.
int hb_probe_pcie_gbs(int64_t bytes, int reps, double *gbs) {
    if (!gbs || bytes <= 0 || reps <= 0) return fail(HB_ERR_INVALID, "bad argument");
    int rc = ensure_device();
    if (rc) return rc;
    void *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t a0 = nullptr, a1 = nullptr, b0 = nullptr, b1 = nullptr;
    auto cleanup = [&] {
        if (h0) cudaFreeHost(h0);
        if (h1) cudaFreeHost(h1);
        if (d0) cudaFree(d0);
        if (d1) cudaFree(d1);
        if (s0) cudaStreamDestroy(s0);
        if (s1) cudaStreamDestroy(s1);
        for (cudaEvent_t e : {a0, a1, b0, b1}) if (e) cudaEventDestroy(e);
    };
#define HB_P(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cleanup(); return fail(HB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } } while (0)
    HB_P(cudaMallocHost(&h0, (size_t)bytes));
    HB_P(cudaMallocHost(&h1, (size_t)bytes));
    HB_P(cudaMalloc(&d0, (size_t)bytes));
    HB_P(cudaMalloc(&d1, (size_t)bytes));
    memset(h0, 1, (size_t)bytes);
    memset(h1, 2, (size_t)bytes);
    HB_P(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    HB_P(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    HB_P(cudaEventCreate(&a0)); HB_P(cudaEventCreate(&a1)); HB_P(cudaEventCreate(&b0)); HB_P(cudaEventCreate(&b1));
    HB_P(cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0));      // warm-up
    HB_P(cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1));
    HB_P(cudaStreamSynchronize(s0)); HB_P(cudaStreamSynchronize(s1));
    float ms = 0;
    const double gb = (double)bytes * reps / 1e9;
    HB_P(cudaEventRecord(a0, s0));
    for (int i = 0; i < reps; ++i) HB_P(cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0));
    HB_P(cudaEventRecord(a1, s0));
    HB_P(cudaEventSynchronize(a1));
    HB_P(cudaEventElapsedTime(&ms, a0, a1));
    gbs[0] = gb / (ms * 1e-3);
    HB_P(cudaEventRecord(b0, s1));
    for (int i = 0; i < reps; ++i) HB_P(cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1));
    HB_P(cudaEventRecord(b1, s1));
    HB_P(cudaEventSynchronize(b1));
    HB_P(cudaEventElapsedTime(&ms, b0, b1));
    gbs[1] = gb / (ms * 1e-3);
    HB_P(cudaEventRecord(a0, s0));
    HB_P(cudaEventRecord(b0, s1));
    for (int i = 0; i < reps; ++i) {
        HB_P(cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0));
        HB_P(cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1));
    }
    HB_P(cudaEventRecord(a1, s0));
    HB_P(cudaEventRecord(b1, s1));
    HB_P(cudaEventSynchronize(a1));
    HB_P(cudaEventSynchronize(b1));
    HB_P(cudaEventElapsedTime(&ms, a0, a1));
    gbs[2] = gb / (ms * 1e-3);
    HB_P(cudaEventElapsedTime(&ms, b0, b1));
    gbs[3] = gb / (ms * 1e-3);
#undef HB_P
    cleanup();
    return HB_OK;
}

// ---- raster ingest: HOST rasters -> the stepper's [T][N][V] forcing on the device, streamed ------------------------
// (/root/reference/ext/HydroModelsRastersExt/netcdf_loader.jl:29-62, :119-136 build the same array
// on the host with a Julia loop per position.)  The rasters never sit on the device as a whole: time chunks of every
// variable go host -> (pinned staging slot, filled by the library's host threads when the source is pageable) -> device
// scratch slot -> pack kernel -> out, with two slots so that the copy of chunk c+1 runs under the pack of chunk c.
namespace {
struct IngestState {
    std::mutex mu;
    hb_model_s::Buf dev[2], pin[2], ptrs[2];
    cudaStream_t copy = nullptr;
    cudaEvent_t landed[2] = {nullptr, nullptr}, packed[2] = {nullptr, nullptr};
};
IngestState g_ingest;
}  // namespace

extern "C" int hb_pack_forcing_device(const double *const *, int, int64_t, int64_t, int64_t, int64_t, const int32_t *, const int32_t *, int64_t,
                                      double *, void *);
extern "C" int hb_pack_forcing_dense_device(const double *const *, int, int64_t, int64_t, int64_t, int64_t, int, int, double *, void *);

int hb_ingest_forcing(const double *const *vars_host, int32_t n_vars, int64_t n_steps, int64_t t_stride, int64_t row_stride,
                      int64_t col_stride, int32_t rows, int32_t cols, const int32_t *node_row, const int32_t *node_col, int64_t n_nodes,
                      double *out, void *stream) {
    if (!vars_host || n_vars <= 0 || n_vars > 64 || n_steps < 0 || rows <= 0 || cols <= 0 || !out) return fail(HB_ERR_INVALID, "bad ingest arguments");
    const int64_t plane = (int64_t)rows * cols;
    if (t_stride != plane) return fail(HB_ERR_INVALID, "ingest needs rasters whose time axis is the slowest (t_stride == rows * cols)");
    const bool dense = !node_row && !node_col;
    if (!dense && (!node_row || !node_col || n_nodes <= 0)) return fail(HB_ERR_INVALID, "node_row / node_col / n_nodes");
    if (dense && n_vars > 5) return fail(HB_ERR_INVALID, "the dense gather takes at most 5 variables");
    for (int v = 0; v < n_vars; ++v)
        if (!vars_host[v]) return fail(HB_ERR_INVALID, "raster %d is NULL", v);
    if (n_steps == 0) return HB_OK;
    int rc = ensure_device();
    if (rc) return rc;
    HB_CUDA(cudaSetDevice(g_device));
    std::lock_guard<std::mutex> lk(g_ingest.mu);
    IngestState &g = g_ingest;
    cudaStream_t st = (cudaStream_t)stream;
    if (!g.copy) {
        HB_CUDA(cudaStreamCreateWithFlags(&g.copy, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            HB_CUDA(cudaEventCreateWithFlags(&g.landed[k], cudaEventDisableTiming));
            HB_CUDA(cudaEventCreateWithFlags(&g.packed[k], cudaEventDisableTiming));
        }
    }
    // steps per chunk: ~64 MiB of raster data per slot (all variables), at most the pack kernels' 65535 steps per launch
    size_t per_slot = (size_t)64 << 20;
    if (const char *e = getenv("HB_INGEST_CHUNK_BYTES")) { long long v = atoll(e); if (v > 0) per_slot = (size_t)v; }
    const size_t step_bytes = (size_t)plane * 8 * n_vars;
    int64_t cs = (int64_t)(per_slot / step_bytes);
    if (cs < 1) cs = 1;
    if (cs > 32768) cs = 32768;
    if (cs > n_steps) cs = n_steps;
    const int64_t nchunk = (n_steps + cs - 1) / cs;
    const int nslots = nchunk > 1 ? 2 : 1;
    const size_t slot_bytes = step_bytes * (size_t)cs;
    bool pageable = false;
    for (int v = 0; v < n_vars; ++v) pageable = pageable || host_is_pageable(vars_host[v]);
    for (int k = 0; k < nslots; ++k) {
        if ((rc = ws_reserve(g.dev[k], slot_bytes)) || (rc = ws_reserve(g.ptrs[k], 64 * sizeof(void *)))) return rc;
        if (pageable && (rc = pin_reserve(g.pin[k], slot_bytes))) return rc;
        // the slot's pointer table: variable v of the chunk starts at dev + v * cs * plane doubles
        const double *tab[64];
        for (int v = 0; v < n_vars; ++v) tab[v] = (const double *)g.dev[k].p + (size_t)v * cs * plane;
        HB_CUDA(cudaMemcpyAsync(g.ptrs[k].p, tab, n_vars * sizeof(void *), cudaMemcpyHostToDevice, st));
    }
    HB_CUDA(cudaStreamSynchronize(st));         // tables are in place (and `tab` may go out of scope)
    // the copy stream must not overwrite a slot the caller's stream may still be reading from a previous call
    HB_CUDA(cudaEventRecord(g.packed[0], st));
    HB_CUDA(cudaStreamWaitEvent(g.copy, g.packed[0], 0));
    for (int64_t c = 0; c < nchunk; ++c) {
        const int sl = (int)(c % nslots);
        const int64_t t0 = c * cs, nt = (n_steps - t0) < cs ? (n_steps - t0) : cs;
        const size_t var_bytes = (size_t)nt * plane * 8;
        if (c >= nslots) {
            HB_CUDA(cudaStreamWaitEvent(g.copy, g.packed[sl], 0));              // device slot consumed by its pack kernel
            if (pageable) HB_CUDA(cudaEventSynchronize(g.landed[sl]));          // pinned slot's previous chunk is on the device
        }
        for (int v = 0; v < n_vars; ++v) {
            const char *src = (const char *)(vars_host[v] + t0 * plane);
            char *dst = (char *)g.dev[sl].p + (size_t)v * cs * plane * 8;
            if (pageable) {
                char *pin = (char *)g.pin[sl].p + (size_t)v * cs * plane * 8;
                host_pool().copy(pin, src, var_bytes);
                src = pin;
            }
            HB_CUDA(cudaMemcpyAsync(dst, src, var_bytes, cudaMemcpyHostToDevice, g.copy));
        }
        HB_CUDA(cudaEventRecord(g.landed[sl], g.copy));
        HB_CUDA(cudaStreamWaitEvent(st, g.landed[sl], 0));
        double *dst_out = out + (size_t)t0 * (dense ? plane : n_nodes) * n_vars;
        rc = dense ? hb_pack_forcing_dense_device((const double *const *)g.ptrs[sl].p, n_vars, nt, plane, row_stride, col_stride, rows, cols, dst_out, st)
                   : hb_pack_forcing_device((const double *const *)g.ptrs[sl].p, n_vars, nt, plane, row_stride, col_stride, node_row, node_col,
                                            n_nodes, dst_out, st);
        if (rc) return fail(rc, "pack kernel launch failed in hb_ingest_forcing");
        HB_CUDA(cudaEventRecord(g.packed[sl], st));
    }
    HB_CUDA(cudaStreamSynchronize(st));          // the caller's rasters and the staging slots are free again
    return HB_OK;
}

int hb_host_copy(void *dst, const void *src, int64_t bytes) {
    if ((!dst || !src) && bytes > 0) return fail(HB_ERR_INVALID, "NULL argument");
    if (bytes < 0) return fail(HB_ERR_INVALID, "negative size");
    host_pool().copy((char *)dst, (const char *)src, (size_t)bytes);
    return HB_OK;
}

int hb_host_threads(int *n) {
    if (!n) return fail(HB_ERR_INVALID, "n is NULL");
    *n = host_pool().threads();
    return HB_OK;
}

// ---- memory helpers ----------------------------------------------------------------------------------
int hb_malloc_device(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(HB_ERR_INVALID, "bad argument");
    int rc = ensure_device();
    if (rc) return rc;
    cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(HB_ERR_NOMEM, "cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e)); }
    return HB_OK;
}
int hb_free_device(void *ptr) { if (ptr) HB_CUDA(cudaFree(ptr)); return HB_OK; }
int hb_malloc_host(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(HB_ERR_INVALID, "bad argument");
    int rc = ensure_device();
    if (rc) return rc;
    cudaError_t e = cudaMallocHost(ptr, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(HB_ERR_NOMEM, "cudaMallocHost(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e)); }
    return HB_OK;
}
int hb_free_host(void *ptr) { if (ptr) HB_CUDA(cudaFreeHost(ptr)); return HB_OK; }
int hb_memcpy_h2d(void *dst, const void *src, int64_t bytes, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return HB_OK;
}
int hb_memcpy_d2h(void *dst, const void *src, int64_t bytes, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return HB_OK;
}
int hb_memset_device(void *dst, int value, int64_t bytes, void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    HB_CUDA(cudaMemsetAsync(dst, value, (size_t)bytes, (cudaStream_t)stream));
    return HB_OK;
}
int hb_stream_create(void **stream) {
    if (!stream) return fail(HB_ERR_INVALID, "stream is NULL");
    int rc = ensure_device();
    if (rc) return rc;
    cudaStream_t s;
    HB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return HB_OK;
}
int hb_stream_destroy(void *stream) {
    if (stream) HB_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return HB_OK;
}
int hb_stream_synchronize(void *stream) {
    int rc = ensure_device();
    if (rc) return rc;
    HB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return HB_OK;
}

}  // extern "C"
