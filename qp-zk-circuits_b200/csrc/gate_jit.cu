// Gate programs compiled at run time: a qpzk_gate_program (the circuit's gate constraints as a straight-line register
// program, see quotient.cu / include/qpzk.h) is turned into CUDA source -- one statement per program step, the program's
// registers as local variables, immediates as literals -- compiled for sm_100a with NVRTC and launched in place of the
// interpreter.  Replaces what upstream's `evaluate_gate_constraints` does with code monomorphised by rustc
// (qp-plonky2 =1.5.5 plonk/vanishing_poly.rs).
//
// Why: the interpreter spends 52 instructions per program step, half of them on fetch/dispatch and on a register file
// that lives in local memory (profiles/r1_summary.md: long-scoreboard 9.1 per issue).  Compiled, every program register
// is a hardware register (68 for upstream's PoseidonGate, 2904 steps; no local memory, no stack) and the field arithmetic
// is the same text as the library's own (goldilocks.cuh between its JIT markers), so the results are bit-identical by
// construction and the tests compare both paths.  Fully inlined the kernel is 71 k instructions = 1.1 MB and runs at the
// instruction-fetch rate of an SM (520 us at 2^13 rows against 864 us interpreted); with the products as calls of two
// shared functions 20 k instructions stream and the product bodies stay in the instruction cache: 400 us.
//
// NVRTC is loaded with dlopen (no link-time dependency); when it is missing, when compilation fails, for short
// programs, or with QPZK_GATE_JIT=0, the interpreter runs -- on the device, like everything else.  Compiled programs are
// cached per process by a hash of (code, immediates); a cudaLibrary_t serves every device and context.
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ctx.cuh"

namespace {

const char* kPrelude =
#include "gl_jit_src.inc"
    ;

// ---- the six NVRTC entry points we use ----
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    int (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*destroyProgram)(nvrtcProgram*) = nullptr;
    int (*compileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*getCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*getCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*getProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*getProgramLog)(nvrtcProgram, char*) = nullptr;
    bool ok = false;
};

Nvrtc& nvrtc() {
    static Nvrtc n = [] {
        Nvrtc r;
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names)
            if ((r.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!r.h) return r;
        auto sym = [&](const char* s) { return dlsym(r.h, s); };
        r.createProgram = reinterpret_cast<decltype(r.createProgram)>(sym("nvrtcCreateProgram"));
        r.destroyProgram = reinterpret_cast<decltype(r.destroyProgram)>(sym("nvrtcDestroyProgram"));
        r.compileProgram = reinterpret_cast<decltype(r.compileProgram)>(sym("nvrtcCompileProgram"));
        r.getCUBINSize = reinterpret_cast<decltype(r.getCUBINSize)>(sym("nvrtcGetCUBINSize"));
        r.getCUBIN = reinterpret_cast<decltype(r.getCUBIN)>(sym("nvrtcGetCUBIN"));
        r.getProgramLogSize = reinterpret_cast<decltype(r.getProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
        r.getProgramLog = reinterpret_cast<decltype(r.getProgramLog)>(sym("nvrtcGetProgramLog"));
        r.ok = r.createProgram && r.destroyProgram && r.compileProgram && r.getCUBINSize && r.getCUBIN && r.getProgramLogSize && r.getProgramLog;
        return r;
    }();
    return n;
}

uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

struct Compiled { cudaLibrary_t lib = nullptr; cudaKernel_t kernel = nullptr; bool failed = false; };
std::mutex g_mu;
std::map<std::pair<uint64_t, uint64_t>, Compiled> g_cache;   // (hash, num_ops) -> compiled program
std::string g_last_log;
std::atomic<int> g_mode{[] { const char* e = getenv("QPZK_GATE_JIT"); return e ? atoi(e) : 1; }()};

// Code generation options.  block: threads per block.  sync_every: a block barrier after every so many program steps.
// The compiled PoseidonGate is 1.1 MB of straight-line code; left alone, the warps of an SM drift apart, every warp
// streams the whole program through the instruction cache on its own and the kernel is bound by instruction fetch
// (measured: 0.24 instructions per cycle per scheduler).  With a barrier every few dozen steps the warps of a block stay
// inside a window that fits the instruction cache and one fetch serves all of them.
struct JitOptions { uint32_t block, sync_every, calls; };
JitOptions jit_options() {
    static const JitOptions o = [] {
        JitOptions r{128, 48, 1};
        // calls = 1: products are calls of two small functions instead of inline code.  The straight-line kernel is bound
        // by instruction fetch (1.1 MB for PoseidonGate); with the 24-instruction product bodies shared, 323 KB stream
        // through the instruction cache and the bodies stay in it (the compiler keeps every program register in a
        // hardware register across the calls: 64 registers, no stack).
        if (const char* e = getenv("QPZK_GATE_JIT_CALLS")) r.calls = atoi(e) ? 1u : 0u;
        if (const char* e = getenv("QPZK_GATE_JIT_BLOCK")) { const int v = atoi(e); if (v >= 32 && v <= 1024 && v % 32 == 0) r.block = (uint32_t)v; }
        if (const char* e = getenv("QPZK_GATE_JIT_SYNC")) { const int v = atoi(e); r.sync_every = v > 0 ? (uint32_t)v : 0; }
        return r;
    }();
    return o;
}

// Compiled programs are also kept on disk -- $QPZK_JIT_CACHE_DIR, else jit_cache/ beside libqpzk.so -- as
// <hash>_<steps>.cubin, so that one compilation (27 s for PoseidonGate) serves every later process; the build step
// warms it for the gate program it exports (__graft_entry__.build()).
std::string cache_dir() {
    if (const char* e = getenv("QPZK_JIT_CACHE_DIR")) return e;
    Dl_info info;
    if (dladdr(reinterpret_cast<const void*>(&fnv1a), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        const size_t k = p.rfind('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/jit_cache";
    }
    return "jit_cache";
}
std::string cache_path(uint64_t h, uint32_t num_ops) {
    char name[64];
    snprintf(name, sizeof name, "/%016llx_%u.cubin", (unsigned long long)h, num_ops);
    return cache_dir() + name;
}
bool read_file(const std::string& path, std::vector<char>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    const bool ok = n > 0 && fread(out->data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}
void write_file_atomic(const std::string& path, const std::vector<char>& data) {
    const std::string dir = cache_dir();
    mkdir(dir.c_str(), 0755);
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (ok) rename(tmp.c_str(), path.c_str()); else remove(tmp.c_str());
}
uint64_t program_hash(const uint32_t* code, uint32_t num_ops, const uint64_t* imm, uint32_t num_imm, uint32_t num_regs) {
    uint64_t h = fnv1a(code, 16ull * num_ops, 0xcbf29ce484222325ull);
    h = fnv1a(imm, 8ull * num_imm, h);
    h = fnv1a(&num_regs, 4, h);
    const JitOptions o = jit_options();
    h = fnv1a(&o, sizeof o, h);
    return fnv1a(kPrelude, strlen(kPrelude), h);    // a change of the arithmetic text invalidates cached cubins
}

void append_u64(std::string& s, uint64_t v) {
    char buf[32];
    snprintf(buf, sizeof buf, "0x%llxull", (unsigned long long)v);
    s += buf;
}

}  // namespace

namespace qpzk {

// The CUDA source of a (validated) program.  Statement forms mirror gate_program_kernel case by case.
std::string gate_jit_source(const uint32_t* code, uint32_t num_ops, const uint64_t* imm, uint32_t num_regs) {
    std::string s;
    s.reserve((size_t)num_ops * 48 + 16384);
    s += "typedef unsigned long long uint64_t;\ntypedef unsigned int uint32_t;\ntypedef unsigned short uint16_t;\n";
    s += kPrelude;
    const bool calls = jit_options().calls != 0;
    s += "\n}  // namespace qpzk\nusing namespace qpzk;\n";
    if (calls)
        s += "__device__ __noinline__ uint64_t qj_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }\n"
             "__device__ __noinline__ uint64_t qj_ma(uint64_t a, uint64_t b, uint64_t c) { return gl_canon(mul_addp(a, b, c)); }\n";
    s += "extern \"C\" __global__ void __launch_bounds__(" + std::to_string(jit_options().block) + ") qpzk_gate_jit(const uint64_t* __restrict__ w_leaves, uint32_t w_width,\n"
         "    const uint64_t* __restrict__ cs_leaves, uint32_t cs_width, uint64_t* __restrict__ terms, uint32_t num_terms, uint64_t num_points) {\n"
         "  const uint64_t j0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;\n"
         "  const bool live = j0 < num_points;\n  const uint64_t j = live ? j0 : num_points - 1;   // every thread reaches every barrier\n"
         "  const uint64_t* __restrict__ w = w_leaves + j * w_width;\n  const uint64_t* __restrict__ c = cs_leaves + j * cs_width;\n"
         "  uint64_t* __restrict__ o = terms + j * num_terms;\n  uint64_t ";
    for (uint32_t r = 0; r < num_regs; r++) { s += "r" + std::to_string(r); s += r + 1 < num_regs ? "," : ";\n"; }
    uint32_t emitted = 0;
    const uint32_t sync_every = jit_options().sync_every;
    auto R = [](uint32_t r) { return "r" + std::to_string(r); };
    for (uint32_t pc = 0; pc < num_ops; pc++) {
        const uint32_t w0 = code[4 * pc], op = w0 & 0xffu, cc = (w0 >> 8) & 0xffu, d = code[4 * pc + 1], a = code[4 * pc + 2], b = code[4 * pc + 3];
        s += "  ";
        switch (op) {
            case 0: s += R(d) + "=__ldg(w+" + std::to_string(a) + ");"; break;
            case 1: s += R(d) + "=__ldg(c+" + std::to_string(a) + ");"; break;
            case 2: s += R(d) + "="; append_u64(s, imm[a]); s += ";"; break;
            case 3: s += R(d) + "=gl_add(" + R(a) + "," + R(b) + ");"; break;
            case 4: s += R(d) + "=gl_sub(" + R(a) + "," + R(b) + ");"; break;
            case 5: s += R(d) + (calls ? "=qj_mul(" : "=gl_mul(") + R(a) + "," + R(b) + ");"; break;
            case 7: s += R(d) + (calls ? "=qj_ma(" : "=gl_canon(mul_addp(") + R(a) + "," + R(b) + "," + R(cc) + (calls ? ");" : "));"); break;
            case 8: s += R(d) + (calls ? "=qj_mul(" : "=gl_mul(") + R(a) + ","; append_u64(s, imm[b]); s += ");"; break;
            case 9: s += R(d) + "=gl_add(" + R(a) + ","; append_u64(s, imm[b]); s += ");"; break;
            case 10: s += R(d) + (calls ? "=qj_ma(" : "=gl_canon(mul_addp(") + R(a) + ","; append_u64(s, imm[b]); s += "," + R(cc) + (calls ? ");" : "));"); break;
            default: s += "if(live)o[" + std::to_string(emitted++) + "]=" + R(a) + ";"; break;
        }
        s += "\n";
        if (sync_every && (pc + 1) % sync_every == 0 && pc + 1 < num_ops) s += "  __syncthreads();\n";
    }
    s += "}\n";
    return s;
}

// source -> cubin for sm_100a; empty on failure (log in *log)
static std::vector<char> compile_cubin(const std::string& src, std::string* log) {
    std::vector<char> cubin;
    Nvrtc& n = nvrtc();
    nvrtcProgram prog = nullptr;
    if (!n.ok || n.createProgram(&prog, src.c_str(), "qpzk_gate_jit.cu", 0, nullptr, nullptr) != 0) return cubin;
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17"};
    const int rc = n.compileProgram(prog, 2, opts);
    size_t ls = 0;
    if (log && n.getProgramLogSize(prog, &ls) == 0 && ls > 1) { log->resize(ls); n.getProgramLog(prog, &(*log)[0]); }
    size_t cs = 0;
    if (rc == 0 && n.getCUBINSize(prog, &cs) == 0 && cs) {
        cubin.resize(cs);
        if (n.getCUBIN(prog, cubin.data()) != 0) cubin.clear();
    }
    n.destroyProgram(&prog);
    return cubin;
}

// cubin of the program from the disk cache or the compiler (and then into the cache); empty on failure
std::vector<char> gate_jit_cubin(const uint32_t* code, uint32_t num_ops, const uint64_t* imm, uint32_t num_imm, uint32_t num_regs,
                                 std::string* log, bool* from_cache) {
    const uint64_t h = program_hash(code, num_ops, imm, num_imm, num_regs);
    const std::string path = cache_path(h, num_ops);
    std::vector<char> cubin;
    if (from_cache) *from_cache = false;
    if (read_file(path, &cubin) && cubin.size() > 64 && memcmp(cubin.data(), "\177ELF", 4) == 0) {
        if (from_cache) *from_cache = true;
        return cubin;
    }
    cubin = compile_cubin(gate_jit_source(code, num_ops, imm, num_regs), log);
    if (!cubin.empty()) write_file_atomic(path, cubin);
    return cubin;
}

// Returns the compiled kernel of the program, compiling it on first use; nullptr = use the interpreter.
cudaKernel_t gate_jit_get(const uint32_t* code, uint32_t num_ops, const uint64_t* imm, uint32_t num_imm, uint32_t num_regs) {
    const int mode = g_mode.load();
    static const uint32_t min_ops = [] { const char* e = getenv("QPZK_GATE_JIT_MIN_OPS"); return e ? (uint32_t)atoi(e) : 512u; }();
    if (!mode || num_ops < min_ops) return nullptr;
    const uint64_t h = program_hash(code, num_ops, imm, num_imm, num_regs);
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(h, (uint64_t)num_ops);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second.failed ? nullptr : it->second.kernel;
    Compiled c;
    c.failed = true;
    const std::vector<char> cubin = gate_jit_cubin(code, num_ops, imm, num_imm, num_regs, &g_last_log, nullptr);
    if (!cubin.empty()) {
        if (cudaLibraryLoadData(&c.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
            cudaLibraryGetKernel(&c.kernel, c.lib, "qpzk_gate_jit") == cudaSuccess)
            c.failed = false;
        else
            cudaGetLastError();
    }
    g_cache[key] = c;
    return c.failed ? nullptr : c.kernel;
}

cudaError_t launch_gate_jit(cudaKernel_t k, const GateProgramArgs& a, cudaStream_t st) {
    if (!a.num_points) return cudaSuccess;
    const uint32_t bs = jit_options().block;
    const unsigned blocks = (unsigned)((a.num_points + bs - 1) / bs);
    const uint64_t* w = a.w_leaves; const uint64_t* cs = a.cs_leaves; uint64_t* terms = a.terms;
    uint32_t ww = a.w_width, cw = a.cs_width, nt = a.num_terms; uint64_t np = a.num_points;
    void* args[] = {&w, &ww, &cs, &cw, &terms, &nt, &np};
    return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(blocks), dim3(bs), args, 0, st);
}

}  // namespace qpzk

// Generates and compiles the program for sm_100a without loading it (needs no GPU) and leaves the cubin in the disk
// cache: 0 and the cubin size on success.
// A build / CI check of the run-time compiler path; the log of a failed compilation goes to the thread's error string.
extern "C" int qpzk_gate_jit_compile_check(const qpzk_gate_program* prog, uint64_t* cubin_bytes) {
    if (!prog || !prog->code || !cubin_bytes) return fail(QPZK_EINVAL, "null argument");
    *cubin_bytes = 0;
    Nvrtc& n = nvrtc();
    if (!n.ok) return fail(QPZK_EINVAL, "NVRTC (libnvrtc.so.12) could not be loaded");
    for (uint32_t pc = 0; pc < prog->num_ops; pc++) {   // the same bounds the upload enforces, as far as the generator needs them
        const uint32_t op = prog->code[4 * pc] & 0xffu, a = prog->code[4 * pc + 2], b = prog->code[4 * pc + 3];
        if (op > 10) return fail(QPZK_EINVAL, "gate program: unknown opcode");
        if ((op == 2 && a >= prog->num_imm) || ((op == 8 || op == 9 || op == 10) && b >= prog->num_imm))
            return fail(QPZK_EINVAL, "gate program: immediate out of range");
    }
    std::string log;
    bool cached = false;
    const std::vector<char> cubin = gate_jit_cubin(prog->code, prog->num_ops, prog->imm, prog->num_imm, prog->num_regs, &log, &cached);
    const size_t cs = cubin.size();
    if (!cs) return fail(QPZK_EINVAL, "gate program did not compile: " + log.substr(0, 2000));
    *cubin_bytes = cs;
    return 0;
}

// 0: interpret every program from now on; 1: compile programs of QPZK_GATE_JIT_MIN_OPS (512) steps or more (default,
// or $QPZK_GATE_JIT).  Process wide; programs already uploaded keep the path they were uploaded with.
extern "C" void qpzk_gate_jit_set_mode(int mode) { g_mode.store(mode ? 1 : 0); }
// 1 = NVRTC is available in this process (compiled gate programs possible), 0 = interpreter only
extern "C" int qpzk_gate_jit_available(void) { return nvrtc().ok ? 1 : 0; }
// Number of gate programs compiled so far in this process (a test hook: shows which path ran).
extern "C" uint64_t qpzk_gate_jit_compiled(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t n = 0;
    for (auto& kv : g_cache) n += kv.second.failed ? 0 : 1;
    return n;
}
