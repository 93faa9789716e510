// jit.hpp — run-time side of the circuit-specialised kernels (see jit_gen.hpp): compiles
// "jit_spec.h + jit_kernels.cuh" with NVRTC for sm_100a, loads the cubin through the
// driver API and launches it on the caller's stream.  NVRTC and the driver are dlopen'ed
// lazily, so the library itself keeps loading on a machine without them (the CPU test
// container); a missing NVRTC is an ERROR for a caller that asked for a specialised
// kernel (ahk_set_jit(e, 1)) and a silent "use the interpreted kernel" for the automatic
// policy.  The kernel sources are embedded in the library at build time (build.sh ->
// jit_sources.inc), nothing is read from the file system at run time.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <cstring>
#include <string>
#include <vector>

namespace ahk {
namespace jit {

struct Source { const char* name; const char* text; };
// name / text of every header the specialised translation unit includes
extern const Source g_sources[];
extern const int g_nsources;

struct Api {
  void* nvrtc = nullptr;
  void* cuda = nullptr;
  bool tried = false, ok = false;
  std::string why;
  // NVRTC
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  // driver
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorStringCu)(CUresult, const char**) = nullptr;
};

inline Api& api() { static Api a; return a; }

inline void* open_first(const std::vector<std::string>& names) {
  for (auto& n : names) {
    if (n.empty()) continue;
    if (void* h = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL)) return h;
  }
  return nullptr;
}

// Loads NVRTC (compiler) and, unless compile_only, the driver API.  Thread-safe.
inline bool load(bool compile_only, std::string& err) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  Api& a = api();
  if (!a.nvrtc) {
    const char* env = getenv("AHK_NVRTC");
    a.nvrtc = open_first({env ? env : "", "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                          "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"});
    if (!a.nvrtc) { err = "NVRTC (libnvrtc.so.12) not found; set AHK_NVRTC to its path"; return false; }
#define SYM(field, name) *(void**)(&a.field) = dlsym(a.nvrtc, name); if (!a.field) { err = std::string("NVRTC lacks ") + name; return false; }
    SYM(CreateProgram, "nvrtcCreateProgram") SYM(DestroyProgram, "nvrtcDestroyProgram")
    SYM(CompileProgram, "nvrtcCompileProgram") SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    SYM(GetCUBIN, "nvrtcGetCUBIN") SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    SYM(GetProgramLog, "nvrtcGetProgramLog") SYM(GetErrorString, "nvrtcGetErrorString")
#undef SYM
  }
  if (compile_only) return true;
  if (!a.cuda) {
    a.cuda = open_first({"libcuda.so.1", "libcuda.so"});
    if (!a.cuda) { err = "CUDA driver library (libcuda.so.1) not found"; return false; }
#define SYM(field, name) *(void**)(&a.field) = dlsym(a.cuda, name); if (!a.field) { err = std::string("libcuda lacks ") + name; return false; }
    SYM(ModuleLoadData, "cuModuleLoadData") SYM(ModuleUnload, "cuModuleUnload")
    SYM(ModuleGetFunction, "cuModuleGetFunction") SYM(FuncSetAttribute, "cuFuncSetAttribute")
    SYM(LaunchKernel, "cuLaunchKernel") SYM(GetErrorStringCu, "cuGetErrorString")
#undef SYM
  }
  return true;
}

// ---- on-disk cache of compiled kernels ----------------------------------------------------
// $AHK_JIT_CACHE (default ~/.cache/ahkab_b200), one <hash>.cubin per (generated constants,
// engine sources): a second process / run with the same circuit skips NVRTC.  AHK_JIT_CACHE=off
// disables it.  Files are written to a temporary name and renamed (atomic).
inline unsigned long long fnv1a(const char* p, size_t n, unsigned long long h = 1469598103934665603ull) {
  for (size_t i = 0; i < n; i++) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
  return h;
}
inline std::string cache_path(const std::string& spec) {
  const char* env = getenv("AHK_JIT_CACHE");
  if (env && std::string(env) == "off") return "";
  std::string dir;
  if (env && *env) dir = env;
  else if (const char* home = getenv("HOME")) dir = std::string(home) + "/.cache/ahkab_b200";
  else return "";
  unsigned long long h = fnv1a(spec.data(), spec.size());
  for (int i = 0; i < g_nsources; i++) h = fnv1a(g_sources[i].text, strlen(g_sources[i].text), h);
  h = fnv1a("sm_100a nvrtc12 lineinfo v1", 27, h);
  char name[64];
  snprintf(name, sizeof name, "/%016llx.cubin", h);
  return dir + name;
}
inline bool cache_read(const std::string& path, std::vector<char>& out) {
  if (path.empty()) return false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  bool ok = n > 64;
  if (ok) { out.resize((size_t)n); ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n; }
  fclose(f);
  if (ok && memcmp(out.data(), "\x7f" "ELF", 4) != 0) ok = false;   // a cubin is an ELF image
  if (!ok) out.clear();
  return ok;
}
inline void cache_write(const std::string& path, const std::vector<char>& cubin) {
  if (path.empty() || cubin.empty()) return;
  const size_t slash = path.rfind('/');
  const std::string dir = path.substr(0, slash);
  for (size_t i = 1; i <= dir.size(); i++)   // mkdir -p
    if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);
  const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  fclose(f);
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

// spec: text of jit_spec.h.  Returns the sm_100a cubin (empty + err on failure); *from_cache
// tells whether NVRTC ran.
inline std::vector<char> compile(const std::string& spec, std::string& err, bool* from_cache = nullptr) {
  std::vector<char> cubin;
  const std::string cpath = cache_path(spec);
  if (from_cache) *from_cache = false;
  if (cache_read(cpath, cubin)) { if (from_cache) *from_cache = true; return cubin; }
  if (!load(true, err)) return cubin;
  Api& a = api();
  std::vector<const char*> names, texts;
  names.push_back("jit_spec.h"); texts.push_back(spec.c_str());
  for (int i = 0; i < g_nsources; i++) { names.push_back(g_sources[i].name); texts.push_back(g_sources[i].text); }
  nvrtcProgram prog;
  const char* tu = "#include \"jit_spec.h\"\n#include \"jit_kernels.cuh\"\n";
  nvrtcResult r = a.CreateProgram(&prog, tu, "ahk_jit.cu", (int)names.size(), texts.data(), names.data());
  if (r != NVRTC_SUCCESS) { err = std::string("nvrtcCreateProgram: ") + a.GetErrorString(r); return cubin; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--extra-device-vectorization"};
  int nopts = 3;
  r = a.CompileProgram(prog, nopts, opts);
  if (r != NVRTC_SUCCESS) {
    size_t ls = 0;
    a.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) a.GetProgramLog(prog, &log[0]);
    err = std::string("NVRTC: ") + a.GetErrorString(r) + "\n" + log;
    a.DestroyProgram(&prog);
    return cubin;
  }
  size_t sz = 0;
  a.GetCUBINSize(prog, &sz);
  cubin.resize(sz);
  if (sz) a.GetCUBIN(prog, cubin.data());
  a.DestroyProgram(&prog);
  if (!sz) err = "NVRTC produced no cubin";
  else cache_write(cpath, cubin);
  return cubin;
}

// A self-contained translation unit (no engine headers): biglin3_gen.hpp's compiled schedules.
// Cached on disk by the hash of its text.
inline std::vector<char> compile_text(const std::string& tu, std::string& err, bool* from_cache = nullptr) {
  std::vector<char> cubin;
  std::string cpath;
  {
    const char* env = getenv("AHK_JIT_CACHE");
    if (!(env && std::string(env) == "off")) {
      std::string dir;
      if (env && *env) dir = env;
      else if (const char* home = getenv("HOME")) dir = std::string(home) + "/.cache/ahkab_b200";
      if (!dir.empty()) {
        unsigned long long h = fnv1a(tu.data(), tu.size());
        h = fnv1a("sm_100a nvrtc12 text v1", 23, h);
        char name[64];
        snprintf(name, sizeof name, "/%016llx.cubin", h);
        cpath = dir + name;
      }
    }
  }
  if (from_cache) *from_cache = false;
  if (cache_read(cpath, cubin)) { if (from_cache) *from_cache = true; return cubin; }
  if (!load(true, err)) return cubin;
  Api& a = api();
  nvrtcProgram prog;
  nvrtcResult r = a.CreateProgram(&prog, tu.c_str(), "ahk_gen.cu", 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) { err = std::string("nvrtcCreateProgram: ") + a.GetErrorString(r); return cubin; }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  r = a.CompileProgram(prog, 3, opts);
  if (r != NVRTC_SUCCESS) {
    size_t ls = 0;
    a.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) a.GetProgramLog(prog, &log[0]);
    err = std::string("NVRTC: ") + a.GetErrorString(r) + "\n" + log.substr(0, 4000);
    a.DestroyProgram(&prog);
    return cubin;
  }
  size_t sz = 0;
  a.GetCUBINSize(prog, &sz);
  cubin.resize(sz);
  if (sz) a.GetCUBIN(prog, cubin.data());
  a.DestroyProgram(&prog);
  if (!sz) err = "NVRTC produced no cubin";
  else cache_write(cpath, cubin);
  return cubin;
}

struct Kernel {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  bool failed = false;     // compilation / load failed once: do not retry on every call
  std::string why;
  bool slu = false;        // built with a static elimination schedule (jit_gen.hpp: slu_codegen)
  bool calibrated = false; // ... whose pivot sequence a calibration launch has checked (engine.cu)
};

inline std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (api().GetErrorStringCu) api().GetErrorStringCu(r, &s);
  return s ? s : "CUDA driver error " + std::to_string((int)r);
}

// Compiles + loads `entry` of the specialised translation unit (current CUDA context =
// the runtime's primary context of the current device).
inline bool build(const std::string& spec, const char* entry, size_t smem, Kernel& k, std::string& err,
                  bool* from_cache = nullptr) {
  if (!load(false, err)) return false;
  std::vector<char> cubin = compile(spec, err, from_cache);
  if (cubin.empty()) return false;
  Api& a = api();
  CUresult r = a.ModuleLoadData(&k.mod, cubin.data());
  if (r != CUDA_SUCCESS) { err = "cuModuleLoadData: " + cu_err(r); return false; }
  r = a.ModuleGetFunction(&k.fn, k.mod, entry);
  if (r != CUDA_SUCCESS) { err = std::string("cuModuleGetFunction(") + entry + "): " + cu_err(r); return false; }
  if (smem > 48 * 1024) {
    r = a.FuncSetAttribute(k.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
    if (r != CUDA_SUCCESS) { err = "cuFuncSetAttribute(smem): " + cu_err(r); return false; }
  }
  return true;
}

// Compiles + loads a self-contained translation unit and resolves two entry points of it
// (k2 shares k1's module: unload k1 only).
inline bool build_text2(const std::string& tu, const char* entry1, const char* entry2, Kernel& k1, Kernel& k2,
                        std::string& err, bool* from_cache = nullptr) {
  if (!load(false, err)) return false;
  std::vector<char> cubin = compile_text(tu, err, from_cache);
  if (cubin.empty()) return false;
  Api& a = api();
  CUresult r = a.ModuleLoadData(&k1.mod, cubin.data());
  if (r != CUDA_SUCCESS) { err = "cuModuleLoadData: " + cu_err(r); return false; }
  r = a.ModuleGetFunction(&k1.fn, k1.mod, entry1);
  if (r != CUDA_SUCCESS) { err = std::string("cuModuleGetFunction(") + entry1 + "): " + cu_err(r); return false; }
  r = a.ModuleGetFunction(&k2.fn, k1.mod, entry2);
  if (r != CUDA_SUCCESS) { err = std::string("cuModuleGetFunction(") + entry2 + "): " + cu_err(r); return false; }
  k2.mod = nullptr;
  return true;
}

inline bool launch(const Kernel& k, int blocks, int threads, size_t smem, void* stream, void* params,
                   std::string& err) {
  void* args[] = {params};
  CUresult r = api().LaunchKernel(k.fn, (unsigned)blocks, 1, 1, (unsigned)threads, 1, 1, (unsigned)smem,
                                  (CUstream)stream, args, nullptr);
  if (r != CUDA_SUCCESS) { err = "cuLaunchKernel: " + cu_err(r); return false; }
  return true;
}

inline void unload(Kernel& k) {
  if (k.mod && api().ModuleUnload) api().ModuleUnload(k.mod);
  k.mod = nullptr; k.fn = nullptr;
}

}  // namespace jit
}  // namespace ahk
