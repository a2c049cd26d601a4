// C-ABI runtime of the B200 evaluator (include/exa_b200.h).
//
// Replaces, behind one handle, what ext/ExaModelsKernelAbstractions.jl does in Julia:
//   build_extension (ext:33-191)      -> exb_create: plan, generate + nvcc + load the model's
//                                        kernel module, re-lay-out iterator data as SoA columns,
//                                        build the sorted (target, slot) lists for grad! / cons!
//   the callbacks (ext:212-547)       -> exb_obj / exb_cons / exb_grad / exb_jac / exb_hess /
//                                        exb_*_structure*: ONE generated launch per callback
//                                        (all patterns), plus a fixed deterministic reduction
//                                        where the reference has one.
// Host-only C++ (g++); generated kernels are launched through driver entry points obtained
// from the CUDA runtime, fixed kernels live in exb_fixed.cu.  There is no CPU fallback.
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is dlopen'ed (see Nccl below)
#include <fcntl.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/exa_b200.h"
#include "exb_device.cuh"
#include "exb_fixed.h"
#include "exb_plan.hpp"

extern const char* exb_device_header_text;  // exb_embed.cpp

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

const char* NVCC_FLAGS_CLEAN = "-cubin -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v";   // -v: registers / spills per kernel into <hash>.cubin.log

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ULL) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

std::string lib_dir() {
  Dl_info info;
  if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
  }
  return ".";
}

std::string cache_dir() {
  const char* e = getenv("EXB_CACHE_DIR");
  std::string d = e && *e ? e : lib_dir() + "/../_kcache";
  mkdir(d.c_str(), 0777);
  return d;
}

std::string nvcc_path() {
  const char* e = getenv("EXB_NVCC");
  if (e && *e) return e;
  if (access("/usr/local/cuda/bin/nvcc", X_OK) == 0) return "/usr/local/cuda/bin/nvcc";
  return "nvcc";
}

// development knob: EXB_TUNE_DEFS="A,B" prepends `#define A 1` ... to the generated module (part of its hash)
std::string extra_defs() {
  std::string out;
  const char* e = getenv("EXB_TUNE_DEFS");
  if (!e) return out;
  std::stringstream ss(e);
  std::string tok;
  while (std::getline(ss, tok, ',')) if (!tok.empty()) out += "#define " + tok + " 1\n";
  return out;
}

// ---- driver entry points (no link-time dependency on libcuda) ---------------------------
struct Drv {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
  bool ok = false;
};
Drv g_drv;
std::once_flag g_drv_once;

template <class F>
bool entry(const char* name, F& f) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  f = (F)p;
  return true;
}
bool load_driver() {
  std::call_once(g_drv_once, [] {
    g_drv.ok = entry("cuModuleLoadData", g_drv.ModuleLoadData) && entry("cuModuleUnload", g_drv.ModuleUnload) &&
               entry("cuModuleGetFunction", g_drv.ModuleGetFunction) && entry("cuModuleGetGlobal", g_drv.ModuleGetGlobal) && entry("cuLaunchKernel", g_drv.LaunchKernel) &&
               entry("cuFuncGetAttribute", g_drv.FuncGetAttribute) && entry("cuFuncSetAttribute", g_drv.FuncSetAttribute) && entry("cuGetErrorString", g_drv.GetErrorString) &&
               entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", g_drv.OccupancyMaxActiveBlocks);
  });
  return g_drv.ok;
}
std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return s ? s : "CUDA driver error " + std::to_string((int)r);
}

// ---- NCCL entry points, dlopen'ed at the first exb_comm_* call: libexa_b200.so has no link-time dependency on NCCL, and a
// process that already loaded a copy (torch's bundled libnccl.so.2) shares it instead of loading a second one --------------
struct Nccl {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false; std::string why;
};
Nccl g_nccl;
std::once_flag g_nccl_once;
bool load_nccl() {
  std::call_once(g_nccl_once, [] {
    void* h = nullptr;
    std::vector<std::string> names;
    if (const char* e = getenv("EXB_NCCL_LIB")) names.push_back(e);
    names.push_back("libnccl.so.2"); names.push_back("libnccl.so");
    for (auto& n : names) if ((h = dlopen(n.c_str(), RTLD_NOW | RTLD_NOLOAD))) break;    // a copy already in the process wins
    if (!h) for (auto& n : names) if ((h = dlopen(n.c_str(), RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) { g_nccl.why = "libnccl.so.2 not found (set EXB_NCCL_LIB)"; return; }
    auto sym = [&](const char* n, auto& f) { f = (std::remove_reference_t<decltype(f)>)dlsym(h, n); return f != nullptr; };
    g_nccl.ok = sym("ncclGetUniqueId", g_nccl.GetUniqueId) && sym("ncclCommInitRank", g_nccl.CommInitRank) &&
                sym("ncclCommDestroy", g_nccl.CommDestroy) && sym("ncclCommCount", g_nccl.CommCount) &&
                sym("ncclCommUserRank", g_nccl.CommUserRank) && sym("ncclAllReduce", g_nccl.AllReduce) &&
                sym("ncclBroadcast", g_nccl.Broadcast) && sym("ncclAllGather", g_nccl.AllGather) &&
                sym("ncclGroupStart", g_nccl.GroupStart) && sym("ncclGroupEnd", g_nccl.GroupEnd) &&
                sym("ncclGetErrorString", g_nccl.GetErrorString);
    if (!g_nccl.ok) g_nccl.why = "NCCL library lacks a required entry point";
  });
  return g_nccl.ok;
}

enum { KN_HESS, KN_JAC, KN_SGRAD, KN_GGRAD, KN_CONS, KN_OBJ, KN_JSTRUCT64, KN_JSTRUCT32, KN_GSTRUCT64, KN_HSTRUCT64, KN_HSTRUCT32, KN_AUGROW, KN_JPROD, KN_JTPROD, KN_HPROD, KN_HESSC, KN_EVAL, KN_GRADT, KN_EVAL1, KN_EVAL0, KN_COUNT };
const char* KNAME[KN_COUNT] = {"exb_hess_g0", "exb_jac_g0", "exb_sgrad_g0", "exb_ggrad_g0", "exb_cons_g0", "exb_obj_g0", "exb_jstruct64_g0",
                               "exb_jstruct32_g0", "exb_gstruct64_g0", "exb_hstruct64_g0", "exb_hstruct32_g0", "exb_augrow_g0",
                               "exb_jprod_g0", "exb_jtprod_g0", "exb_hprod_g0", "exb_hessc_g0", "exb_eval_g0", "exb_gradt_g0", "exb_eval1_g0", "exb_eval0_g0"};

}  // namespace

// One kernel module is compiled per LAUNCH-SHAPE VARIANT (same generated code, different
// __launch_bounds__ min-blocks => different register budget: 32 / 40 / natural).  Which variant is
// fastest depends on the pattern (LV: 32 registers, +25%; AC-OPF: 40; Goddard rocket: natural), so the
// runtime measures each kernel once on the device at its first call and remembers the winner
// (exb_<hash>.tune next to the modules).
struct Variant { int minb; std::string source, hash, cubin_path, cu_path; };
struct exb_plan {
  exb::Plan pl;
  std::vector<Variant> var;
  std::string full_source;            // variant 0, for exb_plan_source
  std::string hash;                   // variant-independent hash (tune file name)
  std::string cubin_path, tune_path;  // variant 0 module path
  bool from_cache = false;
  const std::vector<int>& list(int kn) const {
    switch (kn) {
      case KN_HESS: return pl.k_hess_l;
      case KN_HSTRUCT64: case KN_HSTRUCT32: case KN_HPROD: case KN_HESSC: return pl.k_hess;
      case KN_JAC: case KN_JSTRUCT64: case KN_JSTRUCT32: case KN_JPROD: case KN_JTPROD: return pl.k_jac;
      case KN_SGRAD: case KN_GSTRUCT64: return pl.k_sgrad;
      case KN_GGRAD: return pl.k_ggrad;
      case KN_CONS: return pl.k_cons;
      case KN_OBJ: return pl.k_obj;
      case KN_EVAL: case KN_EVAL1: case KN_EVAL0: return pl.k_eval;
      case KN_GRADT: return pl.k_tgrad;
      default: return pl.k_aug;
    }
  }
};

namespace {

struct Launch {          // one generated kernel, ready to launch
  CUfunction fn = nullptr;            // the variant in use
  std::vector<CUfunction> cand;       // one per compiled variant
  int best = -1;                      // index into cand once tuned
  ExbGroup g{};          // device pointers
  unsigned nblocks = 0;
  unsigned smem = 0;
  // persistent form of the kernel (exb_hessp_g0, only when every pattern has an x window): one more candidate per variant
  std::vector<CUfunction> pcand; std::vector<unsigned> pgrid;
  int use_p = -1;                     // >= 0: index into pcand of the persistent kernel in use
  unsigned psmem = 0; int pw[4] = {0, 0, 0, 0};
  ExbGroup gp{};                      // the persistent kernel's own block -> pattern table
  std::vector<ExbChunk> hchunk;       // host copy of the chunk table (windowed launches of the pipelined host shims)
  std::vector<int> ppt, ns;           // per listed pattern: points per thread, slots per point
  bool is_tile = false; ExbTile tile{};   // column-tile kernel (exb_tile_body): third kernel parameter
};

int make_plan(const void* ir, size_t bytes, exb_plan** out, const void* const* host_data = nullptr, int n_data = 0) {
  if (!ir || !out) return fail(EXB_ERR_ARG, "null argument");
  exb_plan* p = new exb_plan();
  if (!exb::build_plan(p->pl, ir, bytes, host_data, n_data)) {
    std::string e = p->pl.error;
    delete p;
    return fail(EXB_ERR_IR, e);
  }
  std::vector<int> minbs = {16, 12, 1};
  if (getenv("EXB_TUNE_MINB")) minbs = {p->pl.minb};
  std::string d = cache_dir();
  char buf[32];
  snprintf(buf, sizeof buf, "%016llx", (unsigned long long)fnv1a(NVCC_FLAGS_CLEAN, fnv1a(std::string(exb_device_header_text) + extra_defs() + p->pl.source + std::to_string(p->pl.block))));
  p->hash = buf;
  for (int mb : minbs) {
    Variant v; v.minb = mb;
    v.source = "#define EXB_BLOCK " + std::to_string(p->pl.block) + "\n#define EXB_MINB " + std::to_string(mb) + "\n" + extra_defs() +
               (p->pl.idx32 ? "#define EXB_IDX32 1\n" : "") +
               ((int)p->pl.pats.size() <= exb::EXB_CPAT_MAX && !p->pl.pats.empty() ? "#define EXB_NPAT " + std::to_string(p->pl.pats.size()) + "\n" : std::string()) +
               std::string(exb_device_header_text) + "\n" + p->pl.source;
    snprintf(buf, sizeof buf, "%016llx", (unsigned long long)fnv1a(NVCC_FLAGS_CLEAN, fnv1a(v.source)));
    v.hash = buf;
    v.cubin_path = d + "/exb_" + v.hash + ".cubin";
    v.cu_path = d + "/exb_" + v.hash + ".cu";
    p->var.push_back(v);
  }
  p->full_source = p->var[0].source;
  p->cubin_path = p->var[0].cubin_path;
  p->tune_path = d + "/exb_" + p->hash + ".tune";
  *out = p;
  return EXB_OK;
}

std::string shq(const std::string& p) {   // single-quote a path for the shell
  std::string o = "'";
  for (char c : p) { if (c == '\'') o += "'\\''"; else o += c; }
  return o + "'";
}

bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && st.st_size > 0; }

// Compiled modules are kept gzip-compressed in the cache (`<hash>.cubin.gz`): ~80 % of a cubin built with -lineinfo is the
// embedded debug PTX text, which compresses 6x, and the cache travels with every snapshot of the tree.  EXB_KEEP_CUBIN=1
// keeps the raw file next to it (cuobjdump / nvdisasm).
bool module_cached(const std::string& cubin) { return file_exists(cubin + ".gz") || file_exists(cubin); }
bool gz_compress_file(const std::string& in, const std::string& out) {
  std::ifstream f(in, std::ios::binary);
  std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (buf.empty()) return false;
  const std::string tmp = out + ".tmp" + std::to_string((long)getpid());
  gzFile g = gzopen(tmp.c_str(), "wb6");
  if (!g) return false;
  size_t off = 0; bool ok = true;
  while (off < buf.size() && ok) {
    const unsigned n = (unsigned)std::min<size_t>(buf.size() - off, 1u << 24);
    ok = gzwrite(g, buf.data() + off, n) == (int)n;
    off += n;
  }
  ok = (gzclose(g) == Z_OK) && ok;
  if (!ok || rename(tmp.c_str(), out.c_str()) != 0) { unlink(tmp.c_str()); return false; }
  return true;
}
bool read_module(const std::string& cubin, std::vector<char>& image) {
  image.clear();
  if (file_exists(cubin)) {   // a raw module (EXB_KEEP_CUBIN, or a cache written by an older build) wins
    std::ifstream f(cubin, std::ios::binary);
    image.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    if (!image.empty()) return true;   // (else: another process compressed + removed it meanwhile: read the .gz)
  }
  gzFile g = gzopen((cubin + ".gz").c_str(), "rb");
  if (!g) return false;
  std::vector<char> chunk(1u << 22);
  int n;
  while ((n = gzread(g, chunk.data(), (unsigned)chunk.size())) > 0) image.insert(image.end(), chunk.begin(), chunk.begin() + n);
  const bool ok = n == 0;
  gzclose(g);
  if (!ok) image.clear();
  return ok && !image.empty();
}

int compile_plan(exb_plan* p, bool allow_compile) {
  bool all = true;
  for (auto& v : p->var) all = all && module_cached(v.cubin_path);
  if (all) { p->from_cache = true; return EXB_OK; }
  if (!allow_compile) return fail(EXB_ERR_COMPILE, "kernel module " + p->cubin_path + " is not cached and EXB_FLAG_NO_COMPILE is set");
  // serialise concurrent builders of the same model (ranks of one job, parallel tests)
  std::string lock = p->cubin_path + ".lock";
  int fd = open(lock.c_str(), O_CREAT | O_RDWR, 0666);
  if (fd >= 0) flock(fd, LOCK_EX);
  int rc = EXB_OK;
  std::string cmd;   // the variants compile concurrently: one shell, background jobs, wait
  std::vector<const Variant*> todo;
  for (auto& v : p->var) {
    if (module_cached(v.cubin_path)) continue;
    { std::ofstream f(v.cu_path); f << v.source; }
    std::string tmp = v.cubin_path + ".tmp" + std::to_string((long)getpid());
    cmd += "( " + shq(nvcc_path()) + " " + NVCC_FLAGS_CLEAN + " -o " + shq(tmp) + " " + shq(v.cu_path) + " > " + shq(v.cubin_path + ".log") +
           " 2>&1 && mv " + shq(tmp) + " " + shq(v.cubin_path) + " ) & ";
    todo.push_back(&v);
  }
  if (!todo.empty()) {
    cmd += "wait";
    int st = system(cmd.c_str());
    (void)st;
    for (const Variant* v : todo) {
      if (file_exists(v->cubin_path)) {   // compiled: keep it compressed
        if (gz_compress_file(v->cubin_path, v->cubin_path + ".gz") && !getenv("EXB_KEEP_CUBIN")) unlink(v->cubin_path.c_str());
        continue;
      }
      std::ifstream lf(v->cubin_path + ".log");
      std::stringstream ss; ss << lf.rdbuf();
      std::string msg = ss.str();
      if (msg.size() > 4000) msg = msg.substr(0, 4000);
      rc = fail(EXB_ERR_COMPILE, "nvcc failed for " + v->cu_path + ": " + msg);
      break;
    }
  } else {
    p->from_cache = true;
  }
  if (fd >= 0) { flock(fd, LOCK_UN); close(fd); }
  return rc;
}

}  // namespace

struct exb_model {
  exb_plan* plan = nullptr;
  int device = 0, rank = 0, world = 1;
  bool sorted_products = false;   // EXB_FLAG_SORTED_PRODUCTS: the reference's sorted-structure SpMV instead of the fused kernels
  std::vector<CUmodule> mods;
  Launch k[KN_COUNT];
  std::vector<void*> dev;      // everything cudaMalloc'ed by the handle
  size_t dev_bytes = 0;
  double* d_theta = nullptr;
  double* d_objpart = nullptr; double* d_obj = nullptr;
  double* d_objpart_e = nullptr; long long n_objpart_e = 0;   // objective partials of the fused evaluation kernel
  double* d_objpart_e1 = nullptr; double* d_objpart_e0 = nullptr;   // ... of its first-order / value-only forms (one partial per block)
  double* d_gradbuf = nullptr; double* d_conbuf = nullptr;
  // sorted (target, slot) lists: grad (ext:39-46) and constraint augmentation (ext:48-53)
  void *g_slot = nullptr, *g_target = nullptr, *g_ptr = nullptr; long long g_runs = 0; int g_i32 = 0, g_dense = 0;
  void *g_long = nullptr, *a_long = nullptr;   // runs of thousands of slots on one target (exb_fx_long_runs): summed by chunks
  std::vector<void*> long_lists;              // every such list, for exb_destroy
  void *a_slot = nullptr, *a_target = nullptr, *a_ptr = nullptr; long long a_runs = 0; int a_i32 = 0, a_dense = 0;
  // sorted structure for the matrix-free products (ext:56-175) and the duplicate-free COO (utils.jl:425-579)
  struct Sorted { void *slot = nullptr, *other = nullptr, *target = nullptr, *ptr = nullptr, *lng = nullptr; long long runs = 0, nslots = 0; int i32 = 0, dense = 0;
                  long long *urows = nullptr, *ucols = nullptr; };
  Sorted jrow, jcol, hrow, hcol, jcmp, hcmp;
  bool prod_ready = false, cmp_ready = false, cmpj_ready = false;
  bool hess_tile = false;   // duplicate-free Hessian straight from the column-tile kernel (exb_hessc_g0): no sorted list, no raw buffer
  double *d_jacbuf = nullptr, *d_hessbuf = nullptr;
  std::vector<long long> lo, hi;   // local point range per pattern
  long long x_lo = 0, x_hi = 0;    // [x_lo, x_hi): the part of x this handle's points can read (the host shims upload only that)
  long long v_lo = 0, v_hi = 0;    // [v_lo, v_hi): the variables this handle owns (0-based; all of them unless sharded)
  long long last_h2d = 0, last_d2h = 0;   // bytes moved by the last exb_host_* call
  // host shims
  cudaStream_t hstream = nullptr, hstream2 = nullptr;
  std::vector<cudaEvent_t> hev;
  double *hx = nullptr, *hy = nullptr, *hout = nullptr; size_t hx_n = 0, hy_n = 0, hout_n = 0;
  double *dx = nullptr, *dy = nullptr, *dout = nullptr; size_t dx_n = 0, dy_n = 0, dout_n = 0;
  long long launches = 0, last_launches = 0;
  std::string dev_name;            // device name without blanks (part of the tuning key)
  double tune_s = 0, create_s = 0, nvcc_s = 0, load_s = 0, plan_s = 0;   // where model-build time went (exb_build_info)
  // multi-GPU: a communicator over the `world` handles of one sharded model (exb_comm_*); with it the reducing callbacks
  // complete themselves on the caller's stream
  ncclComm_t comm = nullptr; bool comm_owned = false; int comm_mode = EXB_COMM_REPLICATE;
  cudaStream_t cstream = nullptr; cudaEvent_t cev1 = nullptr, cev2 = nullptr;   // side stream of the early objective all-reduce (exb_eval)
  long long collectives = 0, last_collectives = 0;
  // per-callback device timing (the TimedNLPModel role, src/utils.jl:271-408): CUDA events around each callback
  bool timing = false;
  struct Pending { int cb; cudaEvent_t e0, e1; };
  std::vector<Pending> pending;
  double cb_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long cb_calls[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

#define CU_TRY(m, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(EXB_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct DeviceGuard {
  int prev = -1; bool active = false;
  explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); active = true; } }
  ~DeviceGuard() { if (active) cudaSetDevice(prev); }
};

int dmalloc(exb_model* m, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 8;
  CU_TRY(m, cudaMalloc(p, bytes));
  m->dev.push_back(*p);
  m->dev_bytes += bytes;
  return EXB_OK;
}

int launch_fn(exb_model* m, int kn, CUfunction fn, const ExbCall& c, cudaStream_t st) {
  Launch& L = m->k[kn];
  ExbGroup g = L.g;
  ExbCall cc = c;
  ExbTile tt = L.tile;
  void* params[3] = {&g, &cc, &tt};   // the third parameter exists in the column-tile kernels only
  CUresult r = g_drv.LaunchKernel(fn, L.nblocks, 1, 1, (unsigned)m->plan->pl.block, 1, 1, L.smem, (CUstream)st, params, nullptr);
  if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("launch of ") + KNAME[kn] + ": " + cu_err(r));
  m->launches++; m->last_launches++;
  return EXB_OK;
}

int launch_persistent(exb_model* m, int kn, size_t pi, const ExbCall& c, cudaStream_t st) {
  Launch& L = m->k[kn];
  ExbGroup g = L.gp;
  ExbCall cc = c;
  for (int q = 0; q < 4; q++) cc.pw[q] = L.pw[q];
  void* params[2] = {&g, &cc};
  CUresult r = g_drv.LaunchKernel(L.pcand[pi], L.pgrid[pi], 1, 1, (unsigned)m->plan->pl.block, 1, 1, L.psmem, (CUstream)st, params, nullptr);
  if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("launch of the persistent ") + KNAME[kn] + ": " + cu_err(r));
  m->launches++; m->last_launches++;
  return EXB_OK;
}

// Tuning of a kernel's launch-shape variant: run every variant on the given buffers (each one fully defines the output, so
// the result is valid whichever ran last), time them with events -- MEDIAN of 5 runs after a warm-up (9 for launches under
// 50 us, whose timing is noisier) -- and keep the fastest.  Happens at the kernel's first call (that one call synchronises the
// stream; later calls do not), or for every kernel at exb_create / exb_tune (EXB_FLAG_TUNE_AT_CREATE), so that no solver
// callback ever synchronises.  The verdict is remembered per (kernel, device, grid-size class) in exb_<hash>.tune.
std::string tune_key(exb_model* m, int kn) {
  const Launch& L = m->k[kn];
  int lg = 0; for (unsigned v = L.nblocks; v > 1; v >>= 1) lg++;
  return std::string(KNAME[kn]) + "@" + m->dev_name + "@" + std::to_string(lg / 2);   // size class: a factor of 4 in blocks
}
void tune_remember(exb_model* m, const std::string& line) {   // one O_APPEND write per verdict: atomic for concurrent writers
  int fd = open(m->plan->tune_path.c_str(), O_WRONLY | O_APPEND | O_CREAT, 0666);
  if (fd < 0) return;
  ssize_t w = write(fd, line.data(), line.size());
  (void)w;
  close(fd);
}
int tune(exb_model* m, int kn, const ExbCall& c, cudaStream_t st) {
  Launch& L = m->k[kn];
  cudaEvent_t e0, e1;
  CU_TRY(m, cudaEventCreate(&e0)); CU_TRY(m, cudaEventCreate(&e1));
  int rc = EXB_OK; float best_ms = 0, best_classic_ms = 0; int best = 0, best_classic = 0;
  const size_t nc = L.cand.size(), np = L.pcand.size();
  const auto t_begin = std::chrono::steady_clock::now();
  for (size_t v = 0; v < nc + np && !rc; v++) {
    auto run = [&]() { return v < nc ? launch_fn(m, kn, L.cand[v], c, st) : launch_persistent(m, kn, v - nc, c, st); };
    rc = run();                                                    // warm-up (module load, caches)
    std::vector<float> ts;
    int reps = 5;
    for (int rep = 0; rep < reps && !rc; rep++) {
      cudaEventRecord(e0, st);
      rc = run();
      cudaEventRecord(e1, st);
      if (cudaEventSynchronize(e1) != cudaSuccess) rc = fail(EXB_ERR_CUDA, "kernel failed while tuning");
      float t = 0; cudaEventElapsedTime(&t, e0, e1);
      ts.push_back(t);
      if (rep == 4 && reps == 5 && t < 0.05f) reps = 9;
    }
    if (rc) break;
    std::sort(ts.begin(), ts.end());
    const float ms = ts[ts.size() / 2];
    if (v == 0 || ms < best_ms) { best_ms = ms; best = (int)v; }
    if (v < nc && (v == 0 || ms < best_classic_ms)) { best_classic_ms = ms; best_classic = (int)v; }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  m->tune_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  if (rc) return rc;
  // the classic form stays available (windowed launches of the host shims, graph capture); the persistent one is used
  // for whole-grid launches when it measured faster
  L.best = best_classic; L.fn = L.cand[(size_t)best_classic];
  L.use_p = best >= (int)nc ? best - (int)nc : -1;
  std::string line = tune_key(m, kn) + " " + std::to_string(m->plan->var[(size_t)best_classic].minb) + "\n";
  if (np > 0) line += tune_key(m, kn) + "+persistent " + std::to_string(L.use_p >= 0 ? m->plan->var[(size_t)L.use_p].minb : -1) + "\n";
  tune_remember(m, line);
  return EXB_OK;
}

int launch(exb_model* m, int kn, const ExbCall& c, cudaStream_t st) {
  Launch& L = m->k[kn];
  if (!L.fn || L.nblocks == 0) return EXB_OK;
  if (kn == KN_JPROD || kn == KN_JTPROD || kn == KN_HPROD) {   // not idempotent (atomics): never tuned themselves; they use the
    const Launch& V = m->k[kn == KN_HPROD ? KN_HESS : KN_JAC];  // launch-shape variant their value kernel was tuned to
    const size_t vi = V.best >= 0 && (size_t)V.best < L.cand.size() ? (size_t)V.best : 0;
    return launch_fn(m, kn, L.cand[vi], c, st);
  }
  if (L.best < 0) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
    if (cs == cudaStreamCaptureStatusNone) return tune(m, kn, c, st);   // timing needs a stream sync: not while capturing a graph
  }
  if (L.use_p >= 0) return launch_persistent(m, kn, (size_t)L.use_p, c, st);
  return launch_fn(m, kn, L.fn, c, st);
}

// Re-lay-out one AoS field as a device column (int fields narrowed to int32 when they fit).
int upload_column(exb_model* m, const unsigned char* base, long long n, long long stride, const exb::Field& f, const void** col, bool* is32) {
  *is32 = false;
  const bool is_int = f.type == exb::FT_I64 || f.type == exb::FT_I32;
  if (is_int) {
    std::vector<long long> v((size_t)n);
    bool fits = true;
    for (long long k = 0; k < n; k++) {
      const unsigned char* q = base + (size_t)k * (size_t)stride + f.off;
      long long x;
      if (f.type == exb::FT_I64) { int64_t t; memcpy(&t, q, 8); x = t; } else { int32_t t; memcpy(&t, q, 4); x = t; }
      v[(size_t)k] = x;
      if (x > 2147483647LL || x < -2147483648LL) fits = false;
    }
    void* d = nullptr;
    if (fits) {
      std::vector<int> w((size_t)n);
      for (long long k = 0; k < n; k++) w[(size_t)k] = (int)v[(size_t)k];
      int rc = dmalloc(m, &d, (size_t)n * 4); if (rc) return rc;
      CU_TRY(m, cudaMemcpy(d, w.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
      *is32 = true;
    } else {
      int rc = dmalloc(m, &d, (size_t)n * 8); if (rc) return rc;
      CU_TRY(m, cudaMemcpy(d, v.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    *col = d;
  } else {
    std::vector<double> v((size_t)n);
    for (long long k = 0; k < n; k++) {
      const unsigned char* q = base + (size_t)k * (size_t)stride + f.off;
      if (f.type == exb::FT_F64) memcpy(&v[(size_t)k], q, 8); else { float t; memcpy(&t, q, 4); v[(size_t)k] = t; }
    }
    void* d = nullptr;
    int rc = dmalloc(m, &d, (size_t)n * 8); if (rc) return rc;
    CU_TRY(m, cudaMemcpy(d, v.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    *col = d;
  }
  return EXB_OK;
}

// Block -> pattern table of one launch: nb[q] blocks of pattern q, handed out in chunks of (1 << shift) blocks (see ExbGroup)
void make_chunks(const std::vector<long long>& nb, int& shift, std::vector<ExbChunk>& chunks) {
  long long tot = 0;
  for (long long v : nb) tot += v;
  shift = 0;
  while ((tot >> shift) > 4096 && shift < 9) shift++;
  const long long csz = 1LL << shift;
  chunks.clear();
  if (nb.size() <= 4) {   // few patterns: interleave them so shared parts of x are streamed from HBM once
    // proportional merge: always take the pattern that is least far through its own range, so patterns with
    // different points-per-block still walk x side by side
    std::vector<long long> nch(nb.size()), at(nb.size(), 0);
    long long left = 0;
    for (size_t q = 0; q < nb.size(); q++) { nch[q] = (nb[q] + csz - 1) / csz; left += nch[q]; }
    while (left-- > 0) {
      size_t best = 0; double bf = 2.0;
      for (size_t q = 0; q < nb.size(); q++) {
        if (at[q] >= nch[q]) continue;
        const double f = ((double)at[q] + 0.5) / (double)nch[q];
        if (f < bf) { bf = f; best = q; }
      }
      chunks.push_back(ExbChunk{(int)best, (int)(at[best] * csz)});
      at[best]++;
    }
  } else {                 // many patterns: one after the other, so an SM runs one pattern's code at a time
    for (size_t q = 0; q < nb.size(); q++)   // (interleaving 32 patterns thrashes the instruction cache: 3x slower)
      for (long long r = 0; r * csz < nb[q]; r++) chunks.push_back(ExbChunk{(int)q, (int)(r * csz)});
  }
}

int build_model(exb_model* m, const void* const* host_data, int n_data) {
  exb_plan* P = m->plan;
  const exb::Plan& pl = P->pl;
  if (pl.m.ndatabufs > n_data || (pl.m.ndatabufs > 0 && !host_data))
    return fail(EXB_ERR_ARG, "IR references more data buffers than were passed");
  CU_TRY(m, cudaFree(0));
  if (!load_driver()) return fail(EXB_ERR_CUDA, "CUDA driver entry points unavailable");
  {  // sm_100 only
    cudaDeviceProp prop;
    CU_TRY(m, cudaGetDeviceProperties(&prop, m->device));
    m->dev_name = prop.name;
    for (char& ch : m->dev_name) if (ch == ' ' || ch == '@') ch = '_';
    if (prop.major != 10) return fail(EXB_ERR_CUDA, "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; this evaluator targets sm_100a (B200) only");
  }
  // module
  CUresult r = CUDA_SUCCESS;
  for (auto& v : P->var) {
    std::vector<char> image;
    if (!read_module(v.cubin_path, image)) return fail(EXB_ERR_COMPILE, "cannot read " + v.cubin_path + "[.gz]");
    CUmodule mod = nullptr;
    r = g_drv.ModuleLoadData(&mod, image.data());
    if (r != CUDA_SUCCESS) return fail(EXB_ERR_COMPILE, "cuModuleLoadData(" + v.cubin_path + "): " + cu_err(r));
    m->mods.push_back(mod);
  }
  // a previous run's tuning results for this model's kernels: "<kernel>@<device>@<size class> <minb>" lines (last entry wins);
  // they are matched once the grids are known (after the kernel loop below)
  std::vector<std::pair<std::string, int>> tune_lines;
  {
    std::ifstream tf(P->tune_path);
    std::string name; int mb;
    while (tf >> name >> mb) tune_lines.push_back({name, mb});
  }

  // per-pattern arguments
  const size_t np = pl.pats.size();
  std::vector<ExbPatArgs> pa(np);
  m->lo.resize(np); m->hi.resize(np);
  for (size_t k = 0; k < np; k++) {
    const exb::PatternPlan& p = pl.pats[k];
    ExbPatArgs& a = pa[k];
    memset(&a, 0, sizeof a);
    const long long n = p.ir.nitr;
    m->lo[k] = n * m->rank / m->world; m->hi[k] = n * (m->rank + 1) / m->world;
    a.n = m->hi[k] - m->lo[k]; a.k0 = m->lo[k]; a.nfull = n;
    a.start = p.ir.range_start;
    a.o0 = p.o0; a.o1 = p.o1; a.o2 = p.o2; a.aux = p.oa;
    for (size_t d = 0; d < p.ir.dims.size() && d < EXB_MAXD; d++) a.dim[d] = p.ir.dims[d];
    if (p.ir.itr_kind == exb::ITR_AOS) {
      const unsigned char* base = (const unsigned char*)host_data[p.ir.databuf];
      if (!base && n > 0) return fail(EXB_ERR_ARG, "null data buffer");
      for (size_t f = 0; f < p.ir.fields.size(); f++) {
        bool is32 = false;
        if (f < p.ir.iota.size() && p.ir.iota[f]) continue;   // iota column: a function of the point number, never loaded
        int rc = upload_column(m, base, n, p.ir.stride, p.ir.fields[f], &a.col[f], &is32);
        if (rc) return rc;
        if (is32) a.i32mask |= (1LL << f);
      }
    }
  }
  {  // fused evaluation kernel (exb_eval_g0): an objective pattern's blocks leave their partial sums at aux + block number
    long long base = 0;
    for (size_t k = 0; k < np; k++) {
      const exb::PatternPlan& p = pl.pats[k];
      if (p.ir.kind != exb::KIND_OBJ) continue;
      pa[k].aux = base;
      base += (pa[k].n + (long long)pl.block * p.ppte - 1) / ((long long)pl.block * p.ppte);
    }
    m->n_objpart_e = base;
  }
  {  // the part of x the local points can read: shifts of range values and fixed indices are known to the plan
    long long xl = pl.m.nvar, xh = 0; bool all = false;
    for (size_t k = 0; k < np && !all; k++) {
      const exb::PatternPlan& p = pl.pats[k];
      if (pa[k].n == 0) continue;
      if (!p.xr_ok) { all = true; break; }
      if (p.xr_shift) {
        xl = std::min<long long>(xl, p.ir.range_start + pa[k].k0 + p.rlo - 1);
        xh = std::max<long long>(xh, p.ir.range_start + pa[k].k0 + pa[k].n - 1 + p.rhi);
      }
      if (p.xr_fixed) { xl = std::min<long long>(xl, p.flo - 1); xh = std::max<long long>(xh, p.fhi); }
    }
    // variables are partitioned too: rank r OWNS [nvar r / W, nvar (r + 1) / W).  The owner-computes gradient kernel
    // (exb_ggrad_body) writes exactly the owned range, evaluating whichever points touch it (x is replicated), so the
    // gradient of a shift-indexed objective needs NO exchange between ranks at all
    m->v_lo = pl.m.nvar * m->rank / m->world; m->v_hi = pl.m.nvar * (m->rank + 1) / m->world;
    for (size_t k = 0; k < np && !all; k++) {
      const exb::PatternPlan& p = pl.pats[k];
      if (!(p.gather1 || p.tgrad || (pl.tile_ok && p.o2step > 0)) || m->world == 1 || m->v_hi <= m->v_lo) continue;
      if (!p.xr_ok) { all = true; break; }
      const long long w = p.xr_shift ? p.rhi - p.rlo : 0;   // a variable's points read x within this distance of it
      xl = std::min<long long>(xl, m->v_lo - w); xh = std::max<long long>(xh, m->v_hi + w);
    }
    if (all) { xl = 0; xh = pl.m.nvar; }
    xl = std::max<long long>(xl, 0); xh = std::min<long long>(xh, pl.m.nvar);
    if (xh < xl) { xl = 0; xh = 0; }
    m->x_lo = xl; m->x_hi = xh;
  }
  // per-pattern arguments into each module's constant bank (see EXB_PAT in exb_device.cuh)
  if ((int)np <= exb::EXB_CPAT_MAX && np > 0) {
    for (CUmodule mod : m->mods) {
      CUdeviceptr dp = 0; size_t bytes = 0;
      r = g_drv.ModuleGetGlobal(&dp, &bytes, mod, "exb_cpat");
      if (r != CUDA_SUCCESS || bytes != np * sizeof(ExbPatArgs)) return fail(EXB_ERR_COMPILE, "exb_cpat missing from module or of the wrong size");
      CU_TRY(m, cudaMemcpy((void*)dp, pa.data(), bytes, cudaMemcpyHostToDevice));
    }
  }
  // kernels
  for (int kn = 0; kn < KN_COUNT; kn++) {
    const std::vector<int>& lst = P->list(kn);
    if (lst.empty()) continue;
    if (kn == KN_HESSC && !pl.tile_ok) continue;
    Launch& L = m->k[kn];
    for (CUmodule mod : m->mods) {
      CUfunction fn = nullptr;
      r = g_drv.ModuleGetFunction(&fn, mod, KNAME[kn]);
      if (r != CUDA_SUCCESS) return fail(EXB_ERR_COMPILE, std::string("kernel ") + KNAME[kn] + " missing from module: " + cu_err(r));
      L.cand.push_back(fn);
    }
    const bool tunable = kn == KN_HESS || kn == KN_JAC || kn == KN_SGRAD || kn == KN_GGRAD || kn == KN_CONS || kn == KN_OBJ || kn == KN_HESSC || kn == KN_EVAL || kn == KN_GRADT || kn == KN_EVAL1 || kn == KN_EVAL0;
    L.best = (!tunable || L.cand.size() == 1) ? 0 : -1;
    L.fn = L.cand[0];
    const long long BLK = P->pl.block;
    std::vector<ExbPatArgs> args(lst.size());
    std::vector<long long> nb(lst.size());
    long long tot = 0, maxnb = 0; int maxns = 1;
    for (size_t q = 0; q < lst.size(); q++) {
      args[q] = pa[(size_t)lst[q]];
      const exb::PatternPlan& p = pl.pats[(size_t)lst[q]];
      const bool k2 = kn == KN_HESS || kn == KN_HPROD, k1 = kn == KN_JAC || kn == KN_SGRAD || kn == KN_JPROD || kn == KN_JTPROD, k0 = kn == KN_CONS || kn == KN_OBJ;
      const int ppt = kn == KN_EVAL ? p.ppte : kn == KN_EVAL1 ? p.ppte1 : kn == KN_EVAL0 ? p.ppt0 : k2 ? p.ppt2 : k1 ? p.ppt1 : k0 ? p.ppt0 : 1;
      nb[q] = (args[q].n + BLK * ppt - 1) / (BLK * ppt);
      tot += nb[q]; if (nb[q] > maxnb) maxnb = nb[q];
      int ns = kn == KN_EVAL ? p.o1step + p.o2step : kn == KN_EVAL1 ? p.o1step : k2 ? p.o2step : k1 ? p.o1step : 1;
      if (kn == KN_HESS) {   // a split entry stages only its window of slots (rows padded to an odd word count)
        const int nw = pl.k_hess_w[q].second - pl.k_hess_w[q].first;
        if (nw != p.o2step) ns = nw | 1;
      }
      if (ns <= EXB_TILE_MAX_NS && ns * ppt > maxns) maxns = ns * ppt;   // tile words per thread
      if ((kn == KN_EVAL || kn == KN_EVAL1) && p.egrad) {   // staging rows of the in-sweep gradient: (block's points + halo) x odd stride
        const long long words = (BLK * ppt + (p.g_cbmax - p.g_cbmin)) * (long long)(p.o1step | 1);
        const int per_thread = (int)((words + BLK - 1) / BLK);
        if (per_thread > maxns) maxns = per_thread;
      }
    }
    if (kn == KN_HESSC || kn == KN_GRADT) {   // one block per tile of T consecutive COLUMNS of the owned range (exb_tile_body), no block -> pattern map
      void* d_args = nullptr;
      int rc = dmalloc(m, &d_args, args.size() * sizeof(ExbPatArgs)); if (rc) return rc;
      CU_TRY(m, cudaMemcpy(d_args, args.data(), args.size() * sizeof(ExbPatArgs), cudaMemcpyHostToDevice));
      L.g.pat = (const ExbPatArgs*)d_args; L.g.chunk = nullptr; L.g.np = (int)lst.size(); L.g.shift = 0;
      ExbTile& t = L.tile; memset(&t, 0, sizeof t);
      L.is_tile = true;
      t.c_lo = m->v_lo + 1; t.c_hi = m->v_hi + 1;
      const bool hc = kn == KN_HESSC;
      t.T = (int)BLK * (hc ? pl.tile_ppt : pl.tgrad_ppt) - (hc ? pl.tile_halo : pl.tgrad_halo);
      t.D = hc ? (int)pl.hd.size() : 1;
      if (t.T < 32) { if (hc) continue; return fail(EXB_ERR_INTERNAL, "tile gradient: halo too wide for the block shape"); }
      size_t words = (size_t)t.T * (size_t)t.D;
      if (hc) for (int r = 0; r < t.D; r++) { t.lo[r] = pl.h_lo[(size_t)r]; t.len[r] = pl.h_len[(size_t)r]; t.dist[r] = pl.hd[(size_t)r]; }
      for (int pi : lst) {
        const exb::PatternPlan& p = pl.pats[(size_t)pi];
        if (hc) words = std::max(words, (size_t)(t.T + (p.t_cbmax - p.t_cbmin)) * (size_t)(p.o2step | 1));
        else words = std::max(words, (size_t)(t.T + (p.g_cbmax - p.g_cbmin)) * (size_t)(p.o1step | 1));
      }
      L.nblocks = (unsigned)((t.c_hi - t.c_lo + t.T - 1) / t.T);
      words = (words + 1) & ~(size_t)1;
      // two staging buffers (one barrier per pattern instead of two) when there are several patterns AND the doubled footprint
      // still leaves >= 8 blocks per SM resident: on LV (2 patterns, 21 KB tiles) halving the occupancy cost 0.143 -> 0.18 ms
      t.half = (lst.size() > 1 && 16 * words <= 26 * 1024) ? (int)words : 0;
      L.smem = (unsigned)(8 * (words + (size_t)t.half));
      if (L.smem > 48u * 1024u)
        for (CUfunction fn : L.cand) {
          r = g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)L.smem);
          if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("cuFuncSetAttribute(max dynamic smem): ") + cu_err(r));
        }
      if (hc) m->hess_tile = true;
      continue;
    }
    if (kn == KN_GGRAD) {   // one thread per VARIABLE (exb_ggrad_body), no block -> pattern map; runs even when this shard has no points (g = 0)
      void* d_args = nullptr;
      int rc = dmalloc(m, &d_args, args.size() * sizeof(ExbPatArgs)); if (rc) return rc;
      CU_TRY(m, cudaMemcpy(d_args, args.data(), args.size() * sizeof(ExbPatArgs), cudaMemcpyHostToDevice));
      L.g.pat = (const ExbPatArgs*)d_args; L.g.chunk = nullptr; L.g.np = (int)lst.size(); L.g.shift = 0;
      L.nblocks = (unsigned)((m->v_hi - m->v_lo + BLK * EXB_GVPT - 1) / (BLK * EXB_GVPT));
      L.smem = 0;
      continue;
    }
    if (tot == 0) continue;   // nothing local to evaluate (e.g. a shard with no points)
    // chunked round-robin interleave of the patterns' block ranges (see ExbGroup)
    int shift = 0;
    std::vector<ExbChunk> chunks;
    make_chunks(nb, shift, chunks);
    const long long csz = 1LL << shift;
    if ((long long)chunks.size() * csz > 2147483647LL) return fail(EXB_ERR_ARG, "too many blocks");
    void *d_args = nullptr, *d_chunk = nullptr;
    int rc = dmalloc(m, &d_args, args.size() * sizeof(ExbPatArgs)); if (rc) return rc;
    rc = dmalloc(m, &d_chunk, chunks.size() * sizeof(ExbChunk)); if (rc) return rc;
    CU_TRY(m, cudaMemcpy(d_args, args.data(), args.size() * sizeof(ExbPatArgs), cudaMemcpyHostToDevice));
    CU_TRY(m, cudaMemcpy(d_chunk, chunks.data(), chunks.size() * sizeof(ExbChunk), cudaMemcpyHostToDevice));
    L.g.pat = (const ExbPatArgs*)d_args; L.g.chunk = (const ExbChunk*)d_chunk; L.g.np = (int)lst.size(); L.g.shift = shift;
    L.nblocks = (unsigned)(chunks.size() * csz);
    L.hchunk = chunks;
    for (size_t q = 0; q < lst.size(); q++) {
      const exb::PatternPlan& p = pl.pats[(size_t)lst[q]];
      L.ppt.push_back(kn == KN_HESS ? p.ppt2 : (kn == KN_JAC || kn == KN_SGRAD) ? p.ppt1 : (kn == KN_CONS || kn == KN_OBJ) ? p.ppt0 : 1);
      L.ns.push_back(kn == KN_HESS ? p.o2step : (kn == KN_JAC || kn == KN_SGRAD) ? p.o1step : 1);
    }
    L.smem = (kn == KN_HESS || kn == KN_JAC || kn == KN_SGRAD || kn == KN_EVAL || kn == KN_EVAL1) ? (maxns > 1 ? (unsigned)(BLK * maxns * 8) : 16u) : kn == KN_EVAL0 ? 16u : 0u;
    if (L.smem > 48u * 1024u)   // tiles of patterns with many slots per point: opt in to large dynamic shared memory
      for (CUfunction fn : L.cand) {
        r = g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)L.smem);
        if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("cuFuncSetAttribute(max dynamic smem): ") + cu_err(r));
      }
    if (kn == KN_HESS && pl.hess_windowed) {   // persistent form: ONE point per thread per tile (small tiles keep 14+ blocks per SM
      // resident next to the double-buffered staging tile and x / y windows), its own block -> pattern table
      int xw = 2, yw = 0, stw = 2;
      std::vector<long long> nbp(lst.size());
      for (size_t q = 0; q < lst.size(); q++) {
        const exb::PatternPlan& p = pl.pats[(size_t)lst[q]];
        nbp[q] = (args[q].n + BLK - 1) / BLK;
        xw = std::max(xw, (int)BLK + (int)(p.xhi - p.xlo));
        if (p.ir.kind == exb::KIND_CON) yw = std::max(yw, (int)BLK);
        if (p.o2step > 1 && p.o2step <= EXB_TILE_MAX_NS) stw = std::max(stw, (int)BLK * p.o2step);
      }
      xw = (xw + 1) & ~1; yw = (yw + 1) & ~1; stw = (stw + 1) & ~1;
      int pshift = 0;
      std::vector<ExbChunk> pch;
      make_chunks(nbp, pshift, pch);
      if ((long long)pch.size() << pshift > 2147483647LL) return fail(EXB_ERR_ARG, "too many blocks");
      void* d_pch = nullptr;
      int prc = dmalloc(m, &d_pch, pch.size() * sizeof(ExbChunk)); if (prc) return prc;
      CU_TRY(m, cudaMemcpy(d_pch, pch.data(), pch.size() * sizeof(ExbChunk), cudaMemcpyHostToDevice));
      L.gp = L.g; L.gp.chunk = (const ExbChunk*)d_pch; L.gp.shift = pshift;
      L.pw[0] = stw; L.pw[1] = xw; L.pw[2] = yw; L.pw[3] = (int)(pch.size() << pshift);
      L.psmem = (unsigned)(8 * (2 * stw + 2 * (xw + yw)));
      int nsm = 0;
      CU_TRY(m, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, m->device));
      for (CUmodule mod : m->mods) {
        CUfunction fn = nullptr;
        r = g_drv.ModuleGetFunction(&fn, mod, "exb_hessp_g0");
        if (r != CUDA_SUCCESS) return fail(EXB_ERR_COMPILE, std::string("kernel exb_hessp_g0 missing from module: ") + cu_err(r));
        if (L.psmem > 48u * 1024u) {
          r = g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)L.psmem);
          if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("cuFuncSetAttribute(max dynamic smem): ") + cu_err(r));
        }
        int per_sm = 0;
        r = g_drv.OccupancyMaxActiveBlocks(&per_sm, fn, (int)BLK, (size_t)L.psmem);
        if (r != CUDA_SUCCESS || per_sm < 1) per_sm = 1;
        L.pcand.push_back(fn);
        L.pgrid.push_back((unsigned)std::min<long long>((long long)L.pw[3], (long long)per_sm * nsm));
      }
      L.best = -1;        // verdicts (classic and persistent) are matched below
    }
  }
  // remembered tuning verdicts, now that every kernel's grid (hence its size class) is known
  for (int kn = 0; kn < KN_COUNT; kn++) {
    Launch& L = m->k[kn];
    if (L.cand.size() < 2 || L.best >= 0) continue;
    const std::string key = tune_key(m, kn);
    int cls = -1, per = -2;
    for (auto& ln : tune_lines) {
      if (ln.first == key) { cls = -1; for (size_t vi = 0; vi < P->var.size(); vi++) if (P->var[vi].minb == ln.second) cls = (int)vi; }
      if (ln.first == key + "+persistent") { per = -1; for (size_t vi = 0; vi < P->var.size(); vi++) if (P->var[vi].minb == ln.second) per = (int)vi; }
    }
    const bool need_p = !L.pcand.empty();
    if (cls >= 0 && (!need_p || per >= -1)) {
      L.best = cls; L.fn = L.cand[(size_t)cls];
      if (need_p) L.use_p = per;
    }
    if (kn == KN_HESS && need_p) {
      if (const char* e = getenv("EXB_TUNE_FORCE_PERSISTENT")) {   // test / development knob: 1 = always, 0 = never
        if (L.best < 0) { L.best = 0; L.fn = L.cand[0]; }
        L.use_p = atoi(e) ? 0 : -1;
      }
    }
  }
  // scratch owned by the handle (ext:21-31,180-190)
  int rc;
  rc = dmalloc(m, (void**)&m->d_theta, (size_t)pl.m.npar * 8); if (rc) return rc;
  CU_TRY(m, cudaMemset(m->d_theta, 0, (size_t)(pl.m.npar ? pl.m.npar : 1) * 8));
  rc = dmalloc(m, (void**)&m->d_objpart, (size_t)(m->k[KN_OBJ].nblocks + 1) * 8); if (rc) return rc;
  CU_TRY(m, cudaMemset(m->d_objpart, 0, (size_t)(m->k[KN_OBJ].nblocks + 1) * 8));   // padding blocks never write
  rc = dmalloc(m, (void**)&m->d_obj, 8); if (rc) return rc;
  rc = dmalloc(m, (void**)&m->d_objpart_e, (size_t)(m->n_objpart_e + 1) * 8); if (rc) return rc;
  CU_TRY(m, cudaMemset(m->d_objpart_e, 0, (size_t)(m->n_objpart_e + 1) * 8));
  for (int lv = 0; lv < 2; lv++) {   // padding blocks and constraint blocks never write their partial: zeroed once
    double** dp = lv ? &m->d_objpart_e1 : &m->d_objpart_e0;
    const size_t nb_ = (size_t)m->k[lv ? KN_EVAL1 : KN_EVAL0].nblocks + 1;
    rc = dmalloc(m, (void**)dp, nb_ * 8); if (rc) return rc;
    CU_TRY(m, cudaMemset(*dp, 0, nb_ * 8));
  }
  rc = dmalloc(m, (void**)&m->d_gradbuf, (size_t)pl.nnzg * 8); if (rc) return rc;
  rc = dmalloc(m, (void**)&m->d_conbuf, (size_t)pl.nconaug * 8); if (rc) return rc;
  // gradient sparsity: (var, slot) sorted by var (ext:39-46)
  if (pl.nnzg > 0 && m->k[KN_GSTRUCT64].nblocks > 0) {
    long long* keys = nullptr;
    CU_TRY(m, cudaMalloc((void**)&keys, (size_t)pl.nnzg * 8));
    cudaError_t e = exb_fx_fill(keys, pl.nnzg, EXB_FX_SENTINEL, 0);
    ExbCall c{}; c.cols = keys; c.rows = nullptr;
    int lrc = e == cudaSuccess ? launch(m, KN_GSTRUCT64, c, 0) : fail(EXB_ERR_CUDA, cudaGetErrorString(e));
    if (!lrc) {
      long long ns = 0;
      e = exb_fx_sort_runs(keys, pl.nnzg, (long long**)&m->g_slot, (long long**)&m->g_target, (long long**)&m->g_ptr, &m->g_runs, &ns, 0);
      if (e == cudaSuccess) e = exb_fx_pack_runs(&m->g_slot, &m->g_target, &m->g_ptr, ns, m->g_runs, std::max(pl.nnzg, pl.m.nvar), &m->g_i32, &m->g_dense, 0);
      if (e == cudaSuccess) e = exb_fx_long_runs(m->g_ptr, m->g_target, m->g_i32, m->g_runs, ns, &m->g_long, 0);
      if (m->g_long) m->long_lists.push_back(m->g_long);
      if (e != cudaSuccess) lrc = fail(EXB_ERR_CUDA, std::string("gradient sparsity sort: ") + cudaGetErrorString(e));
    }
    cudaFree(keys);
    if (lrc) return lrc;
    for (void* q : {(void*)m->g_slot, (void*)m->g_target, (void*)m->g_ptr}) if (q) m->dev.push_back(q);
  }
  // constraint-augmentation sparsity: (row, slot) sorted by row (ext:48-53, kers ext:199-202)
  if (pl.nconaug > 0 && m->k[KN_AUGROW].nblocks > 0) {
    long long* keys = nullptr;
    CU_TRY(m, cudaMalloc((void**)&keys, (size_t)pl.nconaug * 8));
    cudaError_t e = exb_fx_fill(keys, pl.nconaug, EXB_FX_SENTINEL, 0);
    ExbCall c{}; c.rows = keys;
    int lrc = e == cudaSuccess ? launch(m, KN_AUGROW, c, 0) : fail(EXB_ERR_CUDA, cudaGetErrorString(e));
    if (!lrc) {
      long long ns = 0;
      e = exb_fx_sort_runs(keys, pl.nconaug, (long long**)&m->a_slot, (long long**)&m->a_target, (long long**)&m->a_ptr, &m->a_runs, &ns, 0);
      if (e == cudaSuccess) e = exb_fx_pack_runs(&m->a_slot, &m->a_target, &m->a_ptr, ns, m->a_runs, std::max(pl.nconaug, pl.ncon), &m->a_i32, &m->a_dense, 0);
      if (e == cudaSuccess) e = exb_fx_long_runs(m->a_ptr, m->a_target, m->a_i32, m->a_runs, ns, &m->a_long, 0);
      if (m->a_long) m->long_lists.push_back(m->a_long);
      if (e != cudaSuccess) lrc = fail(EXB_ERR_CUDA, std::string("augmentation sparsity sort: ") + cudaGetErrorString(e));
    }
    cudaFree(keys);
    if (lrc) return lrc;
    for (void* q : {(void*)m->a_slot, (void*)m->a_target, (void*)m->a_ptr}) if (q) m->dev.push_back(q);
  }
  CU_TRY(m, cudaDeviceSynchronize());
  m->launches = 0; m->last_launches = 0;
  return EXB_OK;
}

void free_model(exb_model* m) {
  if (!m) return;
  DeviceGuard dg(m->device);
  if (m->comm && m->comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  if (m->cstream) cudaStreamDestroy(m->cstream);
  if (m->cev1) cudaEventDestroy(m->cev1);
  if (m->cev2) cudaEventDestroy(m->cev2);
  for (void* p : m->dev) cudaFree(p);
  for (void* q : m->long_lists) exb_fx_long_free(q);
  for (CUmodule mod : m->mods) if (mod && g_drv.ModuleUnload) g_drv.ModuleUnload(mod);
  if (m->hx) cudaFreeHost(m->hx);
  if (m->hy) cudaFreeHost(m->hy);
  if (m->hout) cudaFreeHost(m->hout);
  if (m->dx) cudaFree(m->dx);
  if (m->dy) cudaFree(m->dy);
  if (m->dout) cudaFree(m->dout);
  if (m->hstream) cudaStreamDestroy(m->hstream);
  if (m->hstream2) cudaStreamDestroy(m->hstream2);
  for (cudaEvent_t e : m->hev) cudaEventDestroy(e);
  for (auto& p : m->pending) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
  delete m->plan;
  delete m;
}

int ensure_host(exb_model* m, double** h, size_t* hn, double** d, size_t* dn, size_t nh, size_t n) {
  if (n == 0) n = 1;
  if (nh > 0 && *hn < nh) {
    n = n > nh ? n : nh;
    if (*h) cudaFreeHost(*h);
    *h = nullptr; *hn = 0;
    CU_TRY(m, cudaMallocHost((void**)h, nh * 8));
    *hn = nh;
  }
  if (*dn < n) {
    if (*d) cudaFree(*d);
    *d = nullptr; *dn = 0;
    CU_TRY(m, cudaMalloc((void**)d, n * 8));
    *dn = n;
  }
  return EXB_OK;
}


// ---- collectives of a sharded model, issued on the caller's stream through the handle's communicator ------
#define NC_TRY(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(EXB_ERR_CUDA, std::string(#x) + ": " + g_nccl.GetErrorString(r_)); } while (0)
bool comm_on(const exb_model* m) { return m->comm != nullptr && m->world > 1; }
int comm_allreduce(exb_model* m, double* buf, long long n, cudaStream_t st) {
  if (n <= 0) return EXB_OK;
  NC_TRY(g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, m->comm, st));
  m->collectives++; m->last_collectives++;
  return EXB_OK;
}
struct Seg { int rank; long long lo, hi; };
// in-place all-gather of ragged contiguous segments (rank seg.rank holds buf[lo, hi)): one NCCL group of broadcasts
int comm_allgatherv(exb_model* m, double* buf, const std::vector<Seg>& segs, cudaStream_t st) {
  bool any = false;
  for (auto& g : segs) any = any || g.hi > g.lo;
  if (!any) return EXB_OK;
  NC_TRY(g_nccl.GroupStart());
  for (auto& g : segs)
    if (g.hi > g.lo) {
      ncclResult_t r = g_nccl.Broadcast(buf + g.lo, buf + g.lo, (size_t)(g.hi - g.lo), ncclDouble, g.rank, m->comm, st);
      if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(EXB_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(r)); }
    }
  NC_TRY(g_nccl.GroupEnd());
  m->collectives++; m->last_collectives++;
  return EXB_OK;
}
// variables are owned in equal contiguous ranges [nvar r / W, nvar (r + 1) / W)
int comm_allgather_vars(exb_model* m, double* g, cudaStream_t st) {
  const long long nvar = m->plan->pl.m.nvar; const int W = m->world;
  if (nvar % W == 0) {
    const long long cnt = nvar / W;
    NC_TRY(g_nccl.AllGather(g + cnt * m->rank, g, (size_t)cnt, ncclDouble, m->comm, st));
    m->collectives++; m->last_collectives++;
    return EXB_OK;
  }
  std::vector<Seg> segs;
  for (int r = 0; r < W; r++) segs.push_back({r, nvar * r / W, nvar * (r + 1) / W});
  return comm_allgatherv(m, g, segs, st);
}
// rows of the base constraints are owned with their points: pattern k, rank r -> [o0 + n r / W, o0 + n (r + 1) / W)
void row_segments(const exb_model* m, std::vector<Seg>& segs) {
  const exb::Plan& pl = m->plan->pl;
  for (auto& p : pl.pats)
    if (p.ir.kind == exb::KIND_CON)
      for (int r = 0; r < m->world; r++) segs.push_back({r, p.o0 + p.ir.nitr * r / m->world, p.o0 + p.ir.nitr * (r + 1) / m->world});
}
// c / Jv of a sharded handle: own base rows + own augmentation terms, zero elsewhere
int comm_finish_rows(exb_model* m, double* c, cudaStream_t st) {
  if (!comm_on(m)) return EXB_OK;
  const exb::Plan& pl = m->plan->pl;
  if (pl.nconaug > 0) return comm_allreduce(m, c, pl.ncon, st);   // augmentation terms land in arbitrary rows: sum the shards
  if (m->comm_mode == EXB_COMM_OWNER) return EXB_OK;              // every rank already holds the rows of its own points
  std::vector<Seg> segs;
  row_segments(m, segs);
  if (segs.size() > 512) return comm_allreduce(m, c, pl.ncon, st);
  return comm_allgatherv(m, c, segs, st);
}

}  // namespace

enum { CB_OBJ = 0, CB_GRAD, CB_CONS, CB_JAC, CB_HESS, CB_JPROD, CB_JTPROD, CB_HPROD };
struct TimeScope {   // records an event pair on the caller's stream when timing is on (never synchronises)
  exb_model* m; int cb; cudaStream_t st; cudaEvent_t e0 = nullptr;
  TimeScope(exb_model* m_, int cb_, void* st_) : m(m_), cb(cb_), st((cudaStream_t)st_) {
    if (m && m->timing && cudaEventCreate(&e0) == cudaSuccess) cudaEventRecord(e0, st); else e0 = nullptr;
  }
  ~TimeScope() {
    if (!e0) return;
    cudaEvent_t e1 = nullptr;
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return; }
    cudaEventRecord(e1, st);
    m->pending.push_back({cb, e0, e1});
  }
};
#define EXB_GUARD(m) if (!(m) || !(m)->plan) return fail(EXB_ERR_HANDLE, "invalid handle"); DeviceGuard dg_((m)->device); (m)->last_launches = 0; (m)->last_collectives = 0
#define EXB_BEGIN try {
#define EXB_END } catch (const std::exception& e) { return fail(EXB_ERR_INTERNAL, e.what()); } catch (...) { return fail(EXB_ERR_INTERNAL, "unknown exception"); }

extern "C" {

const char* exb_last_error(void) { return g_err.c_str(); }
int exb_abi_version(void) { return EXB_ABI_VERSION; }

int exb_plan_create(const void* ir, size_t ir_bytes, const exb_options* opt, exb_plan** out) {
  EXB_BEGIN
  (void)opt;
  return make_plan(ir, ir_bytes, out);
  EXB_END
}
// same, with the iterator data at hand: lets the plan recognise iota columns (exb_ir.hpp detect_iota), as exb_create does
int exb_plan_create_data(const void* ir, size_t ir_bytes, const void* const* host_data, int n_data, const exb_options* opt, exb_plan** out) {
  EXB_BEGIN
  (void)opt;
  return make_plan(ir, ir_bytes, out, host_data, n_data);
  EXB_END
}
int exb_plan_destroy(exb_plan* p) { delete p; return EXB_OK; }
int exb_plan_dims(const exb_plan* p, int64_t* o) {
  if (!p || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  const exb::Plan& pl = p->pl;
  o[0] = pl.m.nvar; o[1] = pl.ncon; o[2] = pl.nnzj; o[3] = pl.nnzh; o[4] = pl.nobj; o[5] = pl.nnzg; o[6] = pl.nconaug; o[7] = pl.m.npar;
  return EXB_OK;
}
int exb_plan_npatterns(const exb_plan* p) { return p ? (int)p->pl.pats.size() : -1; }
int exb_plan_pattern(const exb_plan* p, int k, int64_t* o) {
  if (!p || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  if (k < 0 || (size_t)k >= p->pl.pats.size()) return fail(EXB_ERR_ARG, "pattern index out of range");
  const exb::PatternPlan& q = p->pl.pats[(size_t)k];
  o[0] = q.ir.kind; o[1] = q.ir.nitr; o[2] = q.o0; o[3] = q.o1; o[4] = q.o2; o[5] = q.o1step; o[6] = q.o2step;
  o[7] = (int64_t)q.comp1.size(); o[8] = (int64_t)q.comp2.size();
  return EXB_OK;
}
int exb_plan_comp(const exb_plan* p, int k, int which, int64_t* o) {
  if (!p || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  if (k < 0 || (size_t)k >= p->pl.pats.size()) return fail(EXB_ERR_ARG, "pattern index out of range");
  const std::vector<int>& c = which == 1 ? p->pl.pats[(size_t)k].comp1 : p->pl.pats[(size_t)k].comp2;
  for (size_t q = 0; q < c.size(); q++) o[q] = c[q];
  return EXB_OK;
}
int exb_plan_tile(const exb_plan* p, int64_t* o) {
  if (!p || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  const exb::Plan& pl = p->pl;
  o[0] = pl.tile_ok ? 1 : 0; o[1] = 0; o[2] = (int64_t)pl.hd.size(); o[3] = pl.tile_halo;
  for (long long v : pl.h_len) o[1] += v;
  return EXB_OK;
}
int exb_plan_source(const exb_plan* p, const char** src, size_t* len) {
  if (!p || !src) return fail(EXB_ERR_HANDLE, "invalid handle");
  *src = p->full_source.c_str();
  if (len) *len = p->full_source.size();
  return EXB_OK;
}
int exb_plan_module_path(const exb_plan* p, char* buf, size_t buflen) {
  if (!p || !buf) return fail(EXB_ERR_HANDLE, "invalid handle");
  if (p->cubin_path.size() + 1 > buflen) return fail(EXB_ERR_ARG, "buffer too small");
  memcpy(buf, p->cubin_path.c_str(), p->cubin_path.size() + 1);
  return EXB_OK;
}
int exb_plan_compile(exb_plan* p) {
  EXB_BEGIN
  if (!p) return fail(EXB_ERR_HANDLE, "invalid handle");
  return compile_plan(p, true);
  EXB_END
}

static int tune_all(exb_model* m, const double* x, const double* y, bool x_on_host, void* stream);

int exb_create(const void* ir, size_t ir_bytes, const void* const* host_data, int n_data, const exb_options* opt, exb_model** out) {
  EXB_BEGIN
  if (!out) return fail(EXB_ERR_ARG, "null argument");
  *out = nullptr;
  exb_options o{-1, 0, 1, 0, 0, nullptr};
  if (opt) o = *opt;
  if (o.world < 1 || o.rank < 0 || o.rank >= o.world) return fail(EXB_ERR_ARG, "bad rank / world");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(EXB_ERR_CUDA, "no CUDA device: this evaluator has no CPU fallback");
  int dev = o.device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return fail(EXB_ERR_CUDA, "cudaGetDevice failed"); }
  if (dev >= ndev) return fail(EXB_ERR_ARG, "device ordinal out of range");
  exb_plan* P = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = make_plan(ir, ir_bytes, &P, host_data, n_data);
  if (rc) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  rc = compile_plan(P, !(o.flags & EXB_FLAG_NO_COMPILE));
  if (rc) { delete P; return rc; }
  const auto t2 = std::chrono::steady_clock::now();
  exb_model* m = new exb_model();
  m->plan = P; m->device = dev; m->rank = o.rank; m->world = o.world;
  m->sorted_products = (o.flags & EXB_FLAG_SORTED_PRODUCTS) != 0;
  DeviceGuard dg(dev);
  rc = build_model(m, host_data, n_data);
  if (rc) { std::string keep = g_err; free_model(m); g_err = keep; return rc; }
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  m->plan_s = secs(t0, t1); m->nvcc_s = P->from_cache ? 0.0 : secs(t1, t2); m->load_s = secs(t2, std::chrono::steady_clock::now());
  if (o.flags & EXB_FLAG_TUNE_AT_CREATE) {   // so that no solver callback ever synchronises (see tune())
    rc = tune_all(m, o.tune_x0, nullptr, true, nullptr);
    if (rc) { std::string keep = g_err; free_model(m); g_err = keep; return rc; }
  }
  m->create_s = secs(t0, std::chrono::steady_clock::now());
  *out = m;
  return EXB_OK;
  EXB_END
}

int exb_destroy(exb_model* m) { free_model(m); return EXB_OK; }

int exb_dims(const exb_model* m, int64_t* o) {
  if (!m) return fail(EXB_ERR_HANDLE, "invalid handle");
  return exb_plan_dims(m->plan, o);
}

int exb_set_params(exb_model* m, const double* theta, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  const long long n = m->plan->pl.m.npar;
  if (n > 0) CU_TRY(m, cudaMemcpyAsync(m->d_theta, theta, (size_t)n * 8, cudaMemcpyDefault, (cudaStream_t)stream));
  return EXB_OK;
  EXB_END
}

int exb_obj_async(exb_model* m, const double* x, double* out_dev, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_OBJ, stream);
  cudaStream_t st = (cudaStream_t)stream;
  ExbCall c{}; c.x = x; c.th = m->d_theta; c.out2 = m->d_objpart;
  int rc = launch(m, KN_OBJ, c, st); if (rc) return rc;
  CU_TRY(m, exb_fx_sum(m->d_objpart, m->k[KN_OBJ].nblocks, out_dev, st));
  m->launches++; m->last_launches++;
  if (comm_on(m)) return comm_allreduce(m, out_dev, 1, st);   // partial sums of the shards (ext:253-271 has one device); 8 bytes
  return EXB_OK;
  EXB_END
}
int exb_obj(exb_model* m, const double* x, double* out_host, void* stream) {
  EXB_BEGIN
  if (!m) return fail(EXB_ERR_HANDLE, "invalid handle");
  int rc = exb_obj_async(m, x, m->d_obj, stream); if (rc) return rc;
  DeviceGuard dg(m->device);
  CU_TRY(m, cudaMemcpyAsync(out_host, m->d_obj, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU_TRY(m, cudaStreamSynchronize((cudaStream_t)stream));
  return EXB_OK;
  EXB_END
}

// grad! = [fill] + owner-computed patterns (exb_ggrad_g0) + slot patterns (exb_sgrad_g0 -> gradbuffer -> segmented sum) + the
// sharded model's collective.  slots_ready: the gradient buffer was already filled by the fused evaluation kernel.
static int grad_impl(exb_model* m, const double* x, double* g, cudaStream_t st, bool slots_ready) {
  const exb::Plan& pl = m->plan->pl;
  ExbCall c{}; c.x = x; c.th = m->d_theta; c.out = m->d_gradbuf;
  const bool tiled = !pl.k_tgrad.empty();
  const bool gathered = !pl.k_ggrad.empty() || tiled;
  // fill!(g, 0) (ext:317): needed when some variable has no objective term -- and on a sharded handle, whose g must be zero
  // outside what it computes so that the shards add up
  // (not when the shards never need adding: a communicator is attached and every objective pattern is owner-computed)
  const bool owner_only = comm_on(m) && pl.k_sgrad.empty() && gathered;
  if ((m->world > 1 && !owner_only) || (m->world == 1 && !gathered && !(m->g_dense && m->g_runs == pl.m.nvar)))
    CU_TRY(m, cudaMemsetAsync(g, 0, (size_t)pl.m.nvar * 8, st));
  if (tiled) {      // shift-indexed objective patterns: a block per tile of OWNED variables evaluates the points around it once and
    ExbCall ct{}; ct.x = x; ct.th = m->d_theta; ct.out = g;   // gathers their slots (exb_gradt_g0); g[v] assigned, 0 where untouched
    int rc = launch(m, KN_GRADT, ct, st); if (rc) return rc;
  }
  if (!pl.k_ggrad.empty()) {   // per-variable form (patterns the tile kernel did not take): one thread per OWNED variable
    ExbCall cg{}; cg.x = x; cg.th = m->d_theta; cg.out = g; cg.v0 = m->v_lo; cg.nout = m->v_hi - m->v_lo; cg.sigma = tiled ? 1.0 : 0.0;
    if (tiled && m->k[KN_GGRAD].best < 0 && m->k[KN_GGRAD].nblocks > 0) {
      // on top of the tile kernel's result the launch ACCUMULATES, so it is not idempotent and must not be repeated by the
      // tuner: rank its variants once in assign mode on a scratch vector (skipped while a graph is being captured)
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusNone; }
      if (cs == cudaStreamCaptureStatusNone) {
        double* tmp = nullptr;
        CU_TRY(m, cudaMalloc((void**)&tmp, (size_t)(pl.m.nvar > 0 ? pl.m.nvar : 1) * 8));
        ExbCall ct = cg; ct.out = tmp; ct.sigma = 0.0;
        int trc = tune(m, KN_GGRAD, ct, st);
        cudaFree(tmp);
        if (trc) return trc;
      } else {
        m->k[KN_GGRAD].best = 0;
      }
    }
    int rc = launch(m, KN_GGRAD, cg, st); if (rc) return rc;
  }
  if (!slots_ready) { int rc = launch(m, KN_SGRAD, c, st); if (rc) return rc; }  // kerg, ext:669-679
  CU_TRY(m, exb_fx_compress(m->d_gradbuf, m->g_ptr, m->g_slot, m->g_target, m->g_i32, m->g_runs, g, (gathered || m->world > 1) ? 1 : 0, m->g_long, st));   // ext:691-697
  if (m->g_runs > 0) { const int nl = m->g_long ? 3 : 1; m->launches += nl; m->last_launches += nl; }
  if (comm_on(m)) {
    // slot-kernel patterns leave partial sums over this shard's points anywhere in g: sum the shards.  A model whose
    // objective patterns are all owner-computed has every g[v] exact on the rank that owns v: nothing to reduce, and
    // nothing to send at all unless the caller wants g replicated
    if (!pl.k_sgrad.empty()) return comm_allreduce(m, g, pl.m.nvar, st);
    if (m->comm_mode == EXB_COMM_REPLICATE) return comm_allgather_vars(m, g, st);
  }
  return EXB_OK;
}
int exb_grad(exb_model* m, const double* x, double* g, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_GRAD, stream);
  return grad_impl(m, x, g, (cudaStream_t)stream, false);
  EXB_END
}

// cons! = [fill] + value kernel (base rows assigned, augmentation terms into conbuffer) + segmented sum of the augmentation
// terms + the sharded model's collective
static int cons_prepare(exb_model* m, double* cvals, cudaStream_t st) {
  const exb::Plan& pl = m->plan->pl;
  // a sharded handle owns only part of the base rows: zero the rest so that ranks can be summed (not needed when a
  // communicator is attached and there is no augmentation: rows are then owned, or all-gathered, never added)
  if (m->world > 1 && !(comm_on(m) && pl.nconaug == 0)) CU_TRY(m, cudaMemsetAsync(cvals, 0, (size_t)pl.ncon * 8, st));
  return EXB_OK;
}
static int cons_finish(exb_model* m, double* cvals, cudaStream_t st) {
  CU_TRY(m, exb_fx_compress(m->d_conbuf, m->a_ptr, m->a_slot, m->a_target, m->a_i32, m->a_runs, cvals, 1, m->a_long, st));   // ext:691-697
  if (m->a_runs > 0) { const int nl = m->a_long ? 3 : 1; m->launches += nl; m->last_launches += nl; }
  return comm_finish_rows(m, cvals, st);
}
int exb_cons(exb_model* m, const double* x, double* cvals, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_CONS, stream);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = cons_prepare(m, cvals, st); if (rc) return rc;
  ExbCall c{}; c.x = x; c.th = m->d_theta; c.out = cvals; c.out2 = m->d_conbuf;
  rc = launch(m, KN_CONS, c, st); if (rc) return rc;                             // kerf + kerf2, ext:681-688
  return cons_finish(m, cvals, st);
  EXB_END
}

int exb_jac(exb_model* m, const double* x, double* vals, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_JAC, stream);
  ExbCall c{}; c.x = x; c.th = m->d_theta; c.out = vals;
  return launch(m, KN_JAC, c, (cudaStream_t)stream);
  EXB_END
}

int exb_hess(exb_model* m, const double* x, const double* y, double obj_weight, double* vals, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_HESS, stream);
  ExbCall c{}; c.x = x; c.y = y; c.th = m->d_theta; c.sigma = obj_weight; c.out = vals;
  return launch(m, KN_HESS, c, (cudaStream_t)stream);
  EXB_END
}

// Sharded handles with a communicator: the objective's all-reduce is a latency (one double, ~20 us over 8 GPUs), not a
// bandwidth cost.  After the sweep and the fixed-order sum of the partials it is started on a SIDE stream, the gradient and
// constraint finishing steps run meanwhile on the caller's stream, and `join` makes the caller's stream wait for the reduced
// objective at the end.  (Tried and measured worse: computing the objective FIRST with the separate value kernel so that the
// all-reduce hides behind the sweep itself -- the extra pass over the objective patterns costs more than the latency it hides:
// 2 GPUs, config 5 full evaluation 0.325 -> 0.356 ms, LV weak 0.257 -> 0.267 ms.)
static int obj_fork(exb_model* m, double* obj_dev, cudaStream_t st) {
  if (!m->cstream) {
    CU_TRY(m, cudaStreamCreateWithFlags(&m->cstream, cudaStreamNonBlocking));
    CU_TRY(m, cudaEventCreateWithFlags(&m->cev1, cudaEventDisableTiming));
    CU_TRY(m, cudaEventCreateWithFlags(&m->cev2, cudaEventDisableTiming));
  }
  CU_TRY(m, cudaEventRecord(m->cev1, st));
  CU_TRY(m, cudaStreamWaitEvent(m->cstream, m->cev1, 0));
  int rc = comm_allreduce(m, obj_dev, 1, m->cstream); if (rc) return rc;
  CU_TRY(m, cudaEventRecord(m->cev2, m->cstream));
  return EXB_OK;
}
static int obj_join(exb_model* m, cudaStream_t st) {
  CU_TRY(m, cudaStreamWaitEvent(st, m->cev2, 0));
  return EXB_OK;
}

// Gradient written by the fused evaluation kernels themselves (exb_eval_block, P::EGRAD): an unsharded handle whose model has ONE
// objective pattern with gradient slots, shift-indexed.  Returns the pointer the sweep writes g through (nullptr: separate
// gradient launch as before) after zeroing the variables no point of the pattern reaches.
static int egrad_prepare(exb_model* m, double* g, cudaStream_t st, double** e_g) {
  *e_g = nullptr;
  const exb::Plan& pl = m->plan->pl;
  static const bool off = getenv("EXB_NO_EGRAD") != nullptr;
  if (off || m->world != 1 || pl.egrad_pat < 0 || !g) return EXB_OK;
  const exb::PatternPlan& p = pl.pats[(size_t)pl.egrad_pat];
  const long long lo = p.ir.range_start + p.g_cbmin, hi = p.ir.range_start + p.ir.nitr - 1 + p.g_cbmax;   // 1-based variables reached
  if (p.ir.nitr <= 0 || lo < 1 || hi > pl.m.nvar) return EXB_OK;
  if (lo > 1 && hi < pl.m.nvar && pl.m.nvar <= (1 << 20)) {   // small model, two ranges: one fill of everything is one launch less
    CU_TRY(m, cudaMemsetAsync(g, 0, (size_t)pl.m.nvar * 8, st));
  } else {
    if (lo > 1) CU_TRY(m, cudaMemsetAsync(g, 0, (size_t)(lo - 1) * 8, st));
    if (hi < pl.m.nvar) CU_TRY(m, cudaMemsetAsync(g + hi, 0, (size_t)(pl.m.nvar - hi) * 8, st));
  }
  *e_g = g;
  return EXB_OK;
}

// One sweep for several callbacks at the same x (the composition of src/nlp.jl:1827-1940).  With every bit of `mask` set each
// data point is evaluated ONCE by exb_eval_g0 (value + first-order + second-order slots; csrc/exb_device.cuh exb_eval_block):
// c / conbuffer, the objective's block partials, jac, the gradient slots and hess are written by that one launch; what
// follows are the same small finishing steps the separate callbacks have (fixed-order sum of the partials, owner-computed
// gradient, segmented sums, the sharded model's collectives).  Any other mask runs the requested callbacks one by one.
// Outputs are the same words the separate callbacks write (same generated code for every slot).
int exb_eval(exb_model* m, unsigned mask, const double* x, const double* y, double obj_weight, double* obj_dev, double* g, double* cvals,
             double* jac, double* hess, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  cudaStream_t st = (cudaStream_t)stream;
  const exb::Plan& pl = m->plan->pl;
  if ((mask & EXB_EVAL_OBJ) && !obj_dev) return fail(EXB_ERR_ARG, "exb_eval: obj requested but obj_dev is NULL");
  if (((mask & EXB_EVAL_GRAD) && !g) || ((mask & EXB_EVAL_CONS) && !cvals && pl.ncon > 0) || ((mask & EXB_EVAL_JAC) && !jac && pl.nnzj > 0) ||
      ((mask & EXB_EVAL_HESS) && !hess && pl.nnzh > 0))
    return fail(EXB_ERR_ARG, "exb_eval: a requested output is NULL");
  static const bool no_fused = getenv("EXB_NO_FUSED_EVAL") != nullptr;
  const bool fused = mask == EXB_EVAL_ALL && !no_fused && m->k[KN_EVAL].fn && m->k[KN_EVAL].nblocks > 0;
  const int level = mask == EXB_EVAL_FIRST ? 1 : mask == EXB_EVAL_VALUES ? 0 : -1;
  const int knl = level == 1 ? KN_EVAL1 : KN_EVAL0;
  if (!fused && level >= 0 && !no_fused && m->k[knl].fn && m->k[knl].nblocks > 0) {
    // first-order evaluation (obj + grad! + cons! + jac_coord!) or values only (obj + cons!) from one sweep
    int rc = cons_prepare(m, cvals, st); if (rc) return rc;
    ExbCall c{}; c.x = x; c.th = m->d_theta;
    c.e_jac = jac; c.e_c = cvals; c.e_gb = m->d_gradbuf; c.e_cb = m->d_conbuf; c.e_obj = level == 1 ? m->d_objpart_e1 : m->d_objpart_e0;
    if (level == 1) { rc = egrad_prepare(m, g, st, &c.e_g); if (rc) return rc; }
    rc = launch(m, knl, c, st); if (rc) return rc;
    CU_TRY(m, exb_fx_sum(c.e_obj, m->k[knl].nblocks, obj_dev, st));
    m->launches++; m->last_launches++;
    const bool side = comm_on(m);
    if (side) { rc = obj_fork(m, obj_dev, st); if (rc) return rc; }
    if (level == 1 && !c.e_g) { rc = grad_impl(m, x, g, st, true); if (rc) return rc; }
    rc = cons_finish(m, cvals, st); if (rc) return rc;
    return side ? obj_join(m, st) : EXB_OK;
  }
  if (!fused) {
    int rc = EXB_OK;
    long long nl = 0, nc = 0;   // the callbacks reset the per-call counters: keep the totals of this call
    auto acc = [&]() { nl += m->last_launches; nc += m->last_collectives; };
    if (!rc && (mask & EXB_EVAL_OBJ)) { rc = exb_obj_async(m, x, obj_dev, stream); acc(); }
    if (!rc && (mask & EXB_EVAL_GRAD)) { rc = exb_grad(m, x, g, stream); acc(); }
    if (!rc && (mask & EXB_EVAL_CONS)) { rc = exb_cons(m, x, cvals, stream); acc(); }
    if (!rc && (mask & EXB_EVAL_JAC)) { rc = exb_jac(m, x, jac, stream); acc(); }
    if (!rc && (mask & EXB_EVAL_HESS)) { rc = exb_hess(m, x, y, obj_weight, hess, stream); acc(); }
    m->last_launches = nl; m->last_collectives = nc;
    return rc;
  }
  int rc = cons_prepare(m, cvals, st); if (rc) return rc;
  ExbCall c{}; c.x = x; c.y = y; c.th = m->d_theta; c.sigma = obj_weight; c.out = hess;
  c.e_jac = jac; c.e_c = cvals; c.e_gb = m->d_gradbuf; c.e_cb = m->d_conbuf; c.e_obj = m->d_objpart_e;
  rc = egrad_prepare(m, g, st, &c.e_g); if (rc) return rc;
  rc = launch(m, KN_EVAL, c, st); if (rc) return rc;
  CU_TRY(m, exb_fx_sum(m->d_objpart_e, m->n_objpart_e, obj_dev, st));
  m->launches++; m->last_launches++;
  const bool side = comm_on(m);
  if (side) { rc = obj_fork(m, obj_dev, st); if (rc) return rc; }
  if (!c.e_g) { rc = grad_impl(m, x, g, st, true); if (rc) return rc; }
  rc = cons_finish(m, cvals, st); if (rc) return rc;
  return side ? obj_join(m, st) : EXB_OK;
  EXB_END
}

static int structure(exb_model* m, int kn, void* rows, void* cols, void* stream) {
  ExbCall c{}; c.rows = rows; c.cols = cols;
  const exb::Plan& pl = m->plan->pl;
  const long long n = (kn == KN_JSTRUCT64 || kn == KN_JSTRUCT32) ? pl.nnzj : pl.nnzh;
  if (n == 0) return EXB_OK;   // nothing to write (e.g. a model without constraints): null outputs are fine
  if (!rows || !cols) return fail(EXB_ERR_ARG, "null rows / cols");
  return launch(m, kn, c, (cudaStream_t)stream);
}
int exb_jac_structure64(exb_model* m, int64_t* rows, int64_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); return structure(m, KN_JSTRUCT64, rows, cols, stream); EXB_END
}
int exb_jac_structure32(exb_model* m, int32_t* rows, int32_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); return structure(m, KN_JSTRUCT32, rows, cols, stream); EXB_END
}
int exb_hess_structure64(exb_model* m, int64_t* rows, int64_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); return structure(m, KN_HSTRUCT64, rows, cols, stream); EXB_END
}
int exb_hess_structure32(exb_model* m, int32_t* rows, int32_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); return structure(m, KN_HSTRUCT32, rows, cols, stream); EXB_END
}

}  // extern "C" (callbacks)

// ---- sorted-structure features -----------------------------------------------------------------------
namespace {

int track(exb_model* m, void* p) { if (p) m->dev.push_back(p); return 0; }

// keys (1-based, device, length n) -> Sorted list; `other_src` (optional): per-slot "other" index gathered into sorted order
int build_sorted(exb_model* m, const long long* keys, long long n, long long max_index, const long long* other_src,
                 bool keep_unique_keys, long long key_mult, exb_model::Sorted& S) {
  long long *slot = nullptr, *target = nullptr, *ptr = nullptr;
  cudaError_t e = exb_fx_sort_runs(keys, n, &slot, &target, &ptr, &S.runs, &S.nslots, 0);
  if (e != cudaSuccess) return fail(EXB_ERR_CUDA, std::string("structure sort: ") + cudaGetErrorString(e));
  if (keep_unique_keys && S.runs > 0) {   // decode the unique (col, row) keys into coordinate arrays
    CU_TRY(m, cudaMalloc((void**)&S.urows, (size_t)S.runs * 8)); CU_TRY(m, cudaMalloc((void**)&S.ucols, (size_t)S.runs * 8));
    CU_TRY(m, exb_fx_decode_keys(target, key_mult, S.ucols, S.urows, S.runs, 0));
    CU_TRY(m, cudaStreamSynchronize(0));
    track(m, S.urows); track(m, S.ucols);
    cudaFree(target); target = nullptr;   // values land at position t: dense targets
  }
  S.slot = slot; S.target = target; S.ptr = ptr;
  e = exb_fx_pack_runs(&S.slot, &S.target, &S.ptr, S.nslots, S.runs, std::max(max_index, n), &S.i32, &S.dense, 0);
  if (e != cudaSuccess) return fail(EXB_ERR_CUDA, std::string("structure pack: ") + cudaGetErrorString(e));
  e = exb_fx_long_runs(S.ptr, S.target, S.i32, S.runs, S.nslots, &S.lng, 0);
  if (e != cudaSuccess) return fail(EXB_ERR_CUDA, std::string("long-run list: ") + cudaGetErrorString(e));
  if (S.lng) m->long_lists.push_back(S.lng);
  if (other_src && S.nslots > 0) {
    CU_TRY(m, cudaMalloc(&S.other, (size_t)S.nslots * (S.i32 ? 4 : 8)));
    CU_TRY(m, exb_fx_gather(other_src, S.slot, S.i32, S.other, S.nslots, 0));
    CU_TRY(m, cudaStreamSynchronize(0));
  }
  track(m, S.slot); track(m, S.target); track(m, S.ptr); track(m, S.other);
  return EXB_OK;
}

int structure_raw(exb_model* m, int kn, void* rows, void* cols, cudaStream_t st);   // defined below

int ensure_buffers(exb_model* m, int passes = 3) {
  const exb::Plan& pl = m->plan->pl;
  if ((passes & 1) && !m->d_jacbuf) { int rc = dmalloc(m, (void**)&m->d_jacbuf, (size_t)pl.nnzj * 8); if (rc) return rc; }
  if ((passes & 2) && !m->d_hessbuf) { int rc = dmalloc(m, (void**)&m->d_hessbuf, (size_t)pl.nnzh * 8); if (rc) return rc; }
  return EXB_OK;
}

// which: 1 = products (row- and column-sorted), 2 = compressed (sorted by (col, row)); passes: bit 0 Jacobian, bit 1 Hessian
int ensure_sorted(exb_model* m, int which, int passes = 3) {
  // products: a sharded handle sorts the slots of ITS points only (everything else carries the sentinel key and is dropped), so its
  // deterministic SpMV yields the shard's partial product.  The duplicate-free COO through the sorted list would need the
  // duplicates of different shards summed: single handle only (shift-indexed models use the tile kernel, which shards).
  if (m->world != 1 && which == 2) return fail(EXB_ERR_ARG, "the duplicate-free COO of a model that is not shift-indexed is built from a sorted "
                                                           "list of ALL slots: not available on sharded handles");
  if (which == 1) { if (m->prod_ready) return EXB_OK; passes = 3; }
  if (which == 2) { if (m->cmpj_ready) passes &= ~1; if (m->cmp_ready) passes &= ~2; if (!passes) return EXB_OK; }
  const exb::Plan& pl = m->plan->pl;
  int rc = ensure_buffers(m, passes); if (rc) return rc;
  for (int pass = 0; pass < 2; pass++) {   // 0: Jacobian, 1: Hessian
    if (!(passes & (1 << pass))) continue;
    const long long n = pass == 0 ? pl.nnzj : pl.nnzh;
    const long long nrow = pass == 0 ? pl.ncon : pl.m.nvar, ncol = pl.m.nvar;
    if (n == 0) continue;
    long long *rows = nullptr, *cols = nullptr, *keys = nullptr;
    CU_TRY(m, cudaMalloc((void**)&rows, (size_t)n * 8));
    cudaError_t e1 = cudaMalloc((void**)&cols, (size_t)n * 8), e2 = cudaMalloc((void**)&keys, (size_t)n * 8);
    rc = (e1 != cudaSuccess || e2 != cudaSuccess) ? fail(EXB_ERR_CUDA, "out of device memory building the sorted structure") : EXB_OK;
    if (!rc && m->world > 1) {   // slots of other shards: sentinel keys (the structure kernels write this handle's points only)
      cudaError_t ef = exb_fx_fill(rows, n, EXB_FX_SENTINEL, 0);
      if (ef == cudaSuccess) ef = exb_fx_fill(cols, n, EXB_FX_SENTINEL, 0);
      if (ef != cudaSuccess) rc = fail(EXB_ERR_CUDA, cudaGetErrorString(ef));
    }
    if (!rc) rc = structure_raw(m, pass == 0 ? KN_JSTRUCT64 : KN_HSTRUCT64, rows, cols, 0);
    if (!rc && cudaStreamSynchronize(0) != cudaSuccess) rc = fail(EXB_ERR_CUDA, "structure kernel failed");
    if (!rc && which == 1) {
      rc = build_sorted(m, rows, n, std::max(nrow, ncol), cols, false, 0, pass == 0 ? m->jrow : m->hrow);
      if (!rc) rc = build_sorted(m, cols, n, std::max(nrow, ncol), rows, false, 0, pass == 0 ? m->jcol : m->hcol);
    }
    if (!rc && which == 2) {
      const long long mult = nrow + 1;   // key = col * (nrow + 1) + row: sorted by (col, row) like the reference's ((j, i), k) tuples
      if ((double)(ncol + 1) * (double)mult > 9.0e18) rc = fail(EXB_ERR_ARG, "model too large for 64-bit (col, row) keys");
      if (!rc) {
        cudaError_t e = exb_fx_make_keys(cols, rows, mult, keys, n, 0);
        if (e != cudaSuccess) rc = fail(EXB_ERR_CUDA, cudaGetErrorString(e));
      }
      if (!rc) rc = build_sorted(m, keys, n, n, nullptr, true, mult, pass == 0 ? m->jcmp : m->hcmp);
    }
    cudaFree(rows); cudaFree(cols); cudaFree(keys);
    if (rc) return rc;
  }
  if (which == 1) m->prod_ready = true;
  else { if (passes & 1) m->cmpj_ready = true; if (passes & 2) m->cmp_ready = true; }
  return EXB_OK;
}

int spmv(exb_model* m, const exb_model::Sorted& S, const double* buf, const double* v, double* y, int acc, int skipdiag, cudaStream_t st) {
  CU_TRY(m, exb_fx_spmv(buf, S.ptr, S.slot, S.other, S.target, S.i32, S.runs, v, y, acc, skipdiag, S.lng, st));
  if (S.runs > 0) { const int nl = S.lng ? 3 : 1; m->launches += nl; m->last_launches += nl; }   // + the two chunked long-run kernels
  return EXB_OK;
}

}  // namespace

extern "C" {

// Matrix-free products.  Default: fused into the derivative sweep (csrc/exb_device.cuh exb_jprod_block / exb_jtprod_block /
// exb_hprod_block): no COO values are written or re-read, no sorted structure is built, sharded handles return partial sums.
// EXB_FLAG_SORTED_PRODUCTS: the reference's device scheme (COO into scratch, then SpMV over a pre-sorted structure, ext:353-511).
int exb_jprod(exb_model* m, const double* x, const double* v, double* Jv, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_JPROD, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const exb::Plan& pl = m->plan->pl;
  if (!m->sorted_products) {
    if (!(comm_on(m) && pl.nconaug == 0)) CU_TRY(m, cudaMemsetAsync(Jv, 0, (size_t)pl.ncon * 8, st));
    if (pl.nconaug > 0) CU_TRY(m, cudaMemsetAsync(m->d_conbuf, 0, (size_t)pl.nconaug * 8, st));
    ExbCall c{}; c.x = x; c.v = v; c.th = m->d_theta; c.out = Jv; c.out2 = m->d_conbuf;
    int rc = launch(m, KN_JPROD, c, st); if (rc) return rc;
    CU_TRY(m, exb_fx_compress(m->d_conbuf, m->a_ptr, m->a_slot, m->a_target, m->a_i32, m->a_runs, Jv, 1, m->a_long, st));
    if (m->a_runs > 0) { m->launches++; m->last_launches++; }
    return comm_finish_rows(m, Jv, st);
  }
  int rc = ensure_sorted(m, 1); if (rc) return rc;
  rc = exb_jac(m, x, m->d_jacbuf, stream); if (rc) return rc;
  CU_TRY(m, cudaMemsetAsync(Jv, 0, (size_t)pl.ncon * 8, st));
  rc = spmv(m, m->jrow, m->d_jacbuf, v, Jv, 0, 0, st); if (rc) return rc;   // kerspmv, ext:482-488
  return comm_on(m) ? comm_allreduce(m, Jv, pl.ncon, st) : EXB_OK;
  EXB_END
}
int exb_jtprod(exb_model* m, const double* x, const double* v, double* Jtv, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_JTPROD, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const exb::Plan& pl = m->plan->pl;
  if (!m->sorted_products) {
    CU_TRY(m, cudaMemsetAsync(Jtv, 0, (size_t)pl.m.nvar * 8, st));
    ExbCall c{}; c.x = x; c.v = v; c.th = m->d_theta; c.out = Jtv;
    int rc = launch(m, KN_JTPROD, c, st); if (rc) return rc;
    return comm_on(m) ? comm_allreduce(m, Jtv, pl.m.nvar, st) : EXB_OK;   // partial products of the shards
  }
  int rc = ensure_sorted(m, 1); if (rc) return rc;
  rc = exb_jac(m, x, m->d_jacbuf, stream); if (rc) return rc;
  CU_TRY(m, cudaMemsetAsync(Jtv, 0, (size_t)pl.m.nvar * 8, st));
  rc = spmv(m, m->jcol, m->d_jacbuf, v, Jtv, 0, 0, st); if (rc) return rc;   // kerspmv2, ext:489-495
  return comm_on(m) ? comm_allreduce(m, Jtv, pl.m.nvar, st) : EXB_OK;
  EXB_END
}
int exb_hprod(exb_model* m, const double* x, const double* y, const double* v, double obj_weight, double* Hv, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  TimeScope ts_(m, CB_HPROD, stream);
  cudaStream_t st = (cudaStream_t)stream;
  const exb::Plan& pl = m->plan->pl;
  if (!m->sorted_products) {
    CU_TRY(m, cudaMemsetAsync(Hv, 0, (size_t)pl.m.nvar * 8, st));
    ExbCall c{}; c.x = x; c.y = y; c.v = v; c.th = m->d_theta; c.sigma = obj_weight; c.out = Hv;
    int rc = launch(m, KN_HPROD, c, st); if (rc) return rc;
    return comm_on(m) ? comm_allreduce(m, Hv, pl.m.nvar, st) : EXB_OK;
  }
  int rc = ensure_sorted(m, 1); if (rc) return rc;
  rc = exb_hess(m, x, y, obj_weight, m->d_hessbuf, stream); if (rc) return rc;
  CU_TRY(m, cudaMemsetAsync(Hv, 0, (size_t)pl.m.nvar * 8, st));
  rc = spmv(m, m->hrow, m->d_hessbuf, v, Hv, 0, 0, st); if (rc) return rc;   // lower triangle incl. diagonal (kersyspmv, ext:496-503)
  rc = spmv(m, m->hcol, m->d_hessbuf, v, Hv, 1, 1, st); if (rc) return rc;   // transpose of the strict lower part (kersyspmv2, ext:504-511)
  return comm_on(m) ? comm_allreduce(m, Hv, pl.m.nvar, st) : EXB_OK;
  EXB_END
}

// Duplicate-free COO.  Hessian of a shift-indexed model (Plan::tile_ok): ONE generated launch (exb_hessc_g0) writes the unique
// entries with their duplicates summed -- no raw values, no sorted list, and the structure / count are closed forms of the
// pattern shifts; available on sharded handles (a rank writes the contiguous range of entries of the columns it owns).
// Everything else: the reference's scheme -- raw COO values, then a segmented sum through the (col, row)-sorted list.
static long long tile_unique(const exb_model* m) { long long n = 0; for (long long v : m->plan->pl.h_len) n += v; return n; }
static long long tile_before(const ExbTile& t, long long c) {
  long long p = 0;
  for (int r = 0; r < t.D; r++) { long long v = c - t.lo[r]; v = v < 0 ? 0 : v; p += v > t.len[r] ? t.len[r] : v; }
  return p;
}
int exb_compressed_dims(exb_model* m, int64_t* nj, int64_t* nh) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (nh) {
    if (m->hess_tile) *nh = tile_unique(m);
    else { int rc = ensure_sorted(m, 2, 2); if (rc) return rc; *nh = m->hcmp.runs; }
  }
  if (nj) { int rc = ensure_sorted(m, 2, 1); if (rc) return rc; *nj = m->jcmp.runs; }
  return EXB_OK;
  EXB_END
}
// out[0..1] = 0-based half-open range of the duplicate-free Hessian values a (sharded) handle writes; out[2] = 1 when the
// Hessian comes straight from the column-tile kernel (one launch, no sorted list)
int exb_compressed_shard(exb_model* m, int64_t* out3) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (!out3) return fail(EXB_ERR_ARG, "null argument");
  if (m->hess_tile) {
    const ExbTile& t = m->k[KN_HESSC].tile;
    out3[0] = tile_before(t, t.c_lo); out3[1] = tile_before(t, t.c_hi); out3[2] = 1;
    return EXB_OK;
  }
  int rc = ensure_sorted(m, 2, 2); if (rc) return rc;
  out3[0] = 0; out3[1] = m->hcmp.runs; out3[2] = 0;
  return EXB_OK;
  EXB_END
}
static int cmp_structure(exb_model* m, const exb_model::Sorted& S, void* rows, void* cols, int idx32, void* stream) {
  if (S.runs == 0) return EXB_OK;
  if (!rows || !cols) return fail(EXB_ERR_ARG, "null rows / cols");
  if (idx32) {
    CU_TRY(m, exb_fx_narrow(S.urows, (int*)rows, S.runs, (cudaStream_t)stream));
    CU_TRY(m, exb_fx_narrow(S.ucols, (int*)cols, S.runs, (cudaStream_t)stream));
  } else {
    CU_TRY(m, cudaMemcpyAsync(rows, S.urows, (size_t)S.runs * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    CU_TRY(m, cudaMemcpyAsync(cols, S.ucols, (size_t)S.runs * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return EXB_OK;
}
static int hess_structure_cmp(exb_model* m, void* rows, void* cols, int idx32, void* stream) {
  if (m->hess_tile) {   // closed form of the pattern shifts; the whole structure, also on a sharded handle
    if (!rows || !cols) return fail(EXB_ERR_ARG, "null rows / cols");
    ExbTile t = m->k[KN_HESSC].tile;
    t.c_lo = 1; t.c_hi = m->plan->pl.m.nvar + 1;
    CU_TRY(m, exb_fx_tile_structure(&t, rows, cols, idx32, (cudaStream_t)stream));
    m->launches++; m->last_launches++;
    return EXB_OK;
  }
  int rc = ensure_sorted(m, 2, 2); if (rc) return rc;
  return cmp_structure(m, m->hcmp, rows, cols, idx32, stream);
}
int exb_jac_structure_compressed64(exb_model* m, int64_t* rows, int64_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); int rc = ensure_sorted(m, 2, 1); if (rc) return rc; return cmp_structure(m, m->jcmp, rows, cols, 0, stream); EXB_END
}
int exb_jac_structure_compressed32(exb_model* m, int32_t* rows, int32_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); int rc = ensure_sorted(m, 2, 1); if (rc) return rc; return cmp_structure(m, m->jcmp, rows, cols, 1, stream); EXB_END
}
int exb_hess_structure_compressed64(exb_model* m, int64_t* rows, int64_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); return hess_structure_cmp(m, rows, cols, 0, stream); EXB_END
}
int exb_hess_structure_compressed32(exb_model* m, int32_t* rows, int32_t* cols, void* stream) {
  EXB_BEGIN EXB_GUARD(m); return hess_structure_cmp(m, rows, cols, 1, stream); EXB_END
}
int exb_jac_compressed(exb_model* m, const double* x, double* vals, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  int rc = ensure_sorted(m, 2, 1); if (rc) return rc;
  rc = exb_jac(m, x, m->d_jacbuf, stream); if (rc) return rc;
  const exb_model::Sorted& S = m->jcmp;
  CU_TRY(m, exb_fx_compress(m->d_jacbuf, S.ptr, S.slot, S.target, S.i32, S.runs, vals, 0, S.lng, (cudaStream_t)stream));   // ker_compress!, ext:1295-1303
  if (S.runs > 0) { m->launches++; m->last_launches++; }
  return EXB_OK;
  EXB_END
}
int exb_hess_compressed(exb_model* m, const double* x, const double* y, double obj_weight, double* vals, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (m->hess_tile) {
    TimeScope ts_(m, CB_HESS, stream);
    ExbCall c{}; c.x = x; c.y = y; c.th = m->d_theta; c.sigma = obj_weight; c.out = vals;
    return launch(m, KN_HESSC, c, (cudaStream_t)stream);
  }
  int rc = ensure_sorted(m, 2, 2); if (rc) return rc;
  rc = exb_hess(m, x, y, obj_weight, m->d_hessbuf, stream); if (rc) return rc;
  const exb_model::Sorted& S = m->hcmp;
  CU_TRY(m, exb_fx_compress(m->d_hessbuf, S.ptr, S.slot, S.target, S.i32, S.runs, vals, 0, S.lng, (cudaStream_t)stream));
  if (S.runs > 0) { m->launches++; m->last_launches++; }
  return EXB_OK;
  EXB_END
}

}  // extern "C"

namespace {
int structure_raw(exb_model* m, int kn, void* rows, void* cols, cudaStream_t st) {
  ExbCall c{}; c.rows = rows; c.cols = cols;
  return launch(m, kn, c, st);
}
}  // namespace

extern "C" {

// ---- host-buffer shims (the WrapperNLPModel role, src/utils.jl:16-267) -----------------------
static int host_stream(exb_model* m) {
  if (!m->hstream) CU_TRY(m, cudaStreamCreateWithFlags(&m->hstream, cudaStreamNonBlocking));
  return EXB_OK;
}
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}
// H2D of x (and y): straight from the caller's buffer when it is page-locked, else through pinned staging
static int host_in(exb_model* m, const double* x, const double* y) {
  const exb::Plan& pl = m->plan->pl;
  int rc = host_stream(m); if (rc) return rc;
  const bool px = is_pinned(x);
  rc = ensure_host(m, &m->hx, &m->hx_n, &m->dx, &m->dx_n, px ? 0 : (size_t)pl.m.nvar, (size_t)pl.m.nvar); if (rc) return rc;
  const size_t xn = (size_t)(m->x_hi - m->x_lo);   // only the part of x this handle's points can read (all of it unless sharded)
  if (!px && xn) memcpy(m->hx + m->x_lo, x + m->x_lo, xn * 8);
  if (xn) CU_TRY(m, cudaMemcpyAsync(m->dx + m->x_lo, (px ? x : m->hx) + m->x_lo, xn * 8, cudaMemcpyHostToDevice, m->hstream));
  m->last_h2d = (long long)xn * 8 + (y ? pl.ncon * 8 : 0); m->last_d2h = 0;
  if (y) {
    const bool py = is_pinned(y);
    rc = ensure_host(m, &m->hy, &m->hy_n, &m->dy, &m->dy_n, py ? 0 : (size_t)pl.ncon, (size_t)pl.ncon); if (rc) return rc;
    if (!py) memcpy(m->hy, y, (size_t)pl.ncon * 8);
    CU_TRY(m, cudaMemcpyAsync(m->dy, py ? y : m->hy, (size_t)pl.ncon * 8, cudaMemcpyHostToDevice, m->hstream));
  }
  return EXB_OK;
}
// D2H of out[lo, hi) slices (a sharded handle returns only what it wrote)
static int host_out(exb_model* m, double* out, const std::vector<std::pair<long long, long long>>& sl) {
  bool staged = false;
  for (auto& s : sl) {
    const size_t n = (size_t)(s.second - s.first);
    if (n == 0) continue;
    m->last_d2h += (long long)n * 8;
    if (is_pinned(out + s.first)) {
      CU_TRY(m, cudaMemcpyAsync(out + s.first, m->dout + s.first, n * 8, cudaMemcpyDeviceToHost, m->hstream));
    } else {
      CU_TRY(m, cudaMemcpyAsync(m->hout + s.first, m->dout + s.first, n * 8, cudaMemcpyDeviceToHost, m->hstream));
      staged = true;
    }
  }
  CU_TRY(m, cudaStreamSynchronize(m->hstream));
  if (staged)
    for (auto& s : sl)
      if (s.second > s.first && !is_pinned(out + s.first)) memcpy(out + s.first, m->hout + s.first, (size_t)(s.second - s.first) * 8);
  return EXB_OK;
}
static std::vector<std::pair<long long, long long>> whole(long long n) { return {{0LL, n}}; }
static std::vector<std::pair<long long, long long>> slices(const exb_model* m, int which) {   // 1: jac, 2: hess
  const exb::Plan& pl = m->plan->pl;
  if (m->world == 1) return whole(which == 1 ? pl.nnzj : pl.nnzh);
  std::vector<std::pair<long long, long long>> v;
  for (size_t k = 0; k < pl.pats.size(); k++) {
    const exb::PatternPlan& p = pl.pats[k];
    if (which == 1 && p.ir.kind == exb::KIND_OBJ) continue;
    const long long o = which == 1 ? p.o1 : p.o2, st = which == 1 ? p.o1step : p.o2step;
    v.push_back({o + m->lo[k] * st, o + m->hi[k] * st});
  }
  return v;
}
// Pipelined form of exb_host_jac / exb_host_hess for page-locked caller buffers: the COO values are 8-9x the size of
// x and y, so the D2H copy is the step's critical path (PCIe).  The launch is cut into windows of consecutive chunks
// of the block -> pattern table; a window covers one contiguous point range of each pattern, hence one contiguous
// output slice per pattern, which is copied back on a second stream while the next window's multipliers go up and
// its kernel runs.  x is needed by every window (indices are data), so it goes first and whole.
static int host_coo_pipelined(exb_model* m, int kn, const double* x, const double* y, double sigma, double* out, bool* done) {
  *done = false;
  Launch& L = m->k[kn];
  const exb::Plan& pl = m->plan->pl;
  static const int wenv = getenv("EXB_HOST_WINDOWS") ? atoi(getenv("EXB_HOST_WINDOWS")) : 8;
  const long long nch = (long long)L.hchunk.size();
  const long long total = kn == KN_HESS ? pl.nnzh : pl.nnzj;
  if (wenv < 2 || L.best < 0 || !L.fn || nch < 2 || total < (1LL << 20)) return EXB_OK;   // untuned kernel / tiny output: plain path
  if (!is_pinned(x) || (y && !is_pinned(y))) return EXB_OK;
  for (auto& s_ : slices(m, kn == KN_HESS ? 2 : 1)) if (s_.second > s_.first && !is_pinned(out + s_.first)) return EXB_OK;
  const std::vector<int>& lst = m->plan->list(kn);
  int rc = host_stream(m); if (rc) return rc;
  if (!m->hstream2) CU_TRY(m, cudaStreamCreateWithFlags(&m->hstream2, cudaStreamNonBlocking));
  const int W = (int)std::min<long long>(wenv, nch);
  while ((int)m->hev.size() < W) { cudaEvent_t e; CU_TRY(m, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); m->hev.push_back(e); }
  rc = ensure_host(m, &m->hx, &m->hx_n, &m->dx, &m->dx_n, 0, (size_t)pl.m.nvar); if (rc) return rc;
  if (y) { rc = ensure_host(m, &m->hy, &m->hy_n, &m->dy, &m->dy_n, 0, (size_t)pl.ncon); if (rc) return rc; }
  rc = ensure_host(m, &m->hout, &m->hout_n, &m->dout, &m->dout_n, 0, (size_t)total); if (rc) return rc;
  // x: when every listed pattern is shift-indexed (variables = range value + const, no fixed-index variable) a window of points
  // reads a window of x, so x goes up window by window too (only the first window's piece is not overlapped with a download);
  // otherwise indices are data and x goes first and whole
  bool xwin = true;
  for (int pi : lst) { const exb::PatternPlan& p = pl.pats[(size_t)pi]; if (!p.xr_ok || p.xr_fixed || !p.xr_shift) xwin = false; }
  long long x_up = m->x_lo;   // x[x_lo, x_up) is on the device
  m->last_h2d = 0; m->last_d2h = 0;
  if (!xwin && m->x_hi > m->x_lo) {
    CU_TRY(m, cudaMemcpyAsync(m->dx + m->x_lo, x + m->x_lo, (size_t)(m->x_hi - m->x_lo) * 8, cudaMemcpyHostToDevice, m->hstream));
    m->last_h2d = (m->x_hi - m->x_lo) * 8; x_up = m->x_hi;
  }
  bool ywin = y != nullptr;   // multipliers can follow the windows only when every row is `o0 + k` (no augmentation in the list)
  for (int pi : lst) if (pl.pats[(size_t)pi].ir.kind == exb::KIND_AUG) ywin = false;
  if (y && !ywin) { CU_TRY(m, cudaMemcpyAsync(m->dy, y, (size_t)pl.ncon * 8, cudaMemcpyHostToDevice, m->hstream)); m->last_h2d += pl.ncon * 8; }
  const long long csz = 1LL << L.g.shift, BLK = pl.block;
  std::vector<long long> bmin(lst.size()), bmax(lst.size());
  for (int w = 0; w < W; w++) {
    const long long c0 = nch * w / W, c1 = nch * (w + 1) / W;
    if (c1 <= c0) continue;
    for (size_t q = 0; q < lst.size(); q++) { bmin[q] = -1; bmax[q] = -1; }
    for (long long c = c0; c < c1; c++) {
      const int q = L.hchunk[(size_t)c].pat;
      if (q < 0) continue;
      const long long b0 = L.hchunk[(size_t)c].b0;
      if (bmin[(size_t)q] < 0 || b0 < bmin[(size_t)q]) bmin[(size_t)q] = b0;
      if (b0 + csz > bmax[(size_t)q]) bmax[(size_t)q] = b0 + csz;
    }
    std::vector<std::pair<long long, long long>> outs;
    if (xwin) {   // the part of x this window's points read (windows advance monotonically through every pattern)
      long long need = x_up;
      for (size_t q = 0; q < lst.size(); q++) {
        if (bmin[q] < 0) continue;
        const size_t pi = (size_t)lst[q];
        const exb::PatternPlan& p = pl.pats[pi];
        const long long n = m->hi[pi] - m->lo[pi], per = BLK * L.ppt[q];
        const long long p1 = std::min(n, bmax[q] * per);
        need = std::max(need, p.ir.range_start + m->lo[pi] + p1 - 1 + p.rhi);   // exclusive 0-based bound of the last variable read
      }
      if (w == W - 1) need = m->x_hi;
      need = std::min(need, m->x_hi);
      if (need > x_up) {
        CU_TRY(m, cudaMemcpyAsync(m->dx + x_up, x + x_up, (size_t)(need - x_up) * 8, cudaMemcpyHostToDevice, m->hstream));
        m->last_h2d += (need - x_up) * 8; x_up = need;
      }
    }
    for (size_t q = 0; q < lst.size(); q++) {
      if (bmin[q] < 0) continue;
      const size_t pi = (size_t)lst[q];
      const exb::PatternPlan& p = pl.pats[pi];
      const long long n = m->hi[pi] - m->lo[pi], per = BLK * L.ppt[q];
      const long long p0 = std::min(n, bmin[q] * per), p1 = std::min(n, bmax[q] * per);
      if (p1 <= p0) continue;
      if (ywin && p.ir.kind == exb::KIND_CON) {
        const long long r0 = p.o0 + m->lo[pi] + p0;
        CU_TRY(m, cudaMemcpyAsync(m->dy + r0, y + r0, (size_t)(p1 - p0) * 8, cudaMemcpyHostToDevice, m->hstream));
        m->last_h2d += (p1 - p0) * 8;
      }
      const long long o = kn == KN_HESS ? p.o2 : p.o1;
      outs.push_back({o + (m->lo[pi] + p0) * L.ns[q], o + (m->lo[pi] + p1) * L.ns[q]});
    }
    {   // this window's blocks: the same kernel over a sub-range of the chunk table
      ExbGroup g = L.g; g.chunk = L.g.chunk + c0;
      ExbCall cc{}; cc.x = m->dx; cc.y = y ? m->dy : nullptr; cc.th = m->d_theta; cc.sigma = sigma; cc.out = m->dout;
      void* params[2] = {&g, &cc};
      CUresult r = g_drv.LaunchKernel(L.fn, (unsigned)((c1 - c0) * csz), 1, 1, (unsigned)pl.block, 1, 1, L.smem, (CUstream)m->hstream, params, nullptr);
      if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("windowed launch of ") + KNAME[kn] + ": " + cu_err(r));
      m->launches++; m->last_launches++;
    }
    CU_TRY(m, cudaEventRecord(m->hev[(size_t)w], m->hstream));
    CU_TRY(m, cudaStreamWaitEvent(m->hstream2, m->hev[(size_t)w], 0));
    for (auto& sl : outs) {
      CU_TRY(m, cudaMemcpyAsync(out + sl.first, m->dout + sl.first, (size_t)(sl.second - sl.first) * 8, cudaMemcpyDeviceToHost, m->hstream2));
      m->last_d2h += (sl.second - sl.first) * 8;
    }
  }
  CU_TRY(m, cudaStreamSynchronize(m->hstream2));
  CU_TRY(m, cudaStreamSynchronize(m->hstream));
  *done = true;
  return EXB_OK;
}

int exb_host_obj(exb_model* m, const double* x, double* out) {
  EXB_BEGIN
  EXB_GUARD(m);
  int rc = host_in(m, x, nullptr); if (rc) return rc;
  return exb_obj(m, m->dx, out, m->hstream);
  EXB_END
}
#define EXB_HOST_VEC(N, SLICES, CALL)                                                                 \
  EXB_BEGIN                                                                                           \
  EXB_GUARD(m);                                                                                       \
  const exb::Plan& pl = m->plan->pl; (void)pl;                                                        \
  int rc = host_in(m, x, yy); if (rc) return rc;                                                      \
  const std::vector<std::pair<long long, long long>> sl = SLICES;                                     \
  bool all_pinned = true;                                                                             \
  for (auto& s_ : sl) if (s_.second > s_.first && !is_pinned(out + s_.first)) all_pinned = false;     \
  rc = ensure_host(m, &m->hout, &m->hout_n, &m->dout, &m->dout_n, all_pinned ? 0 : (size_t)(N), (size_t)(N)); if (rc) return rc; \
  rc = CALL; if (rc) return rc;                                                                       \
  return host_out(m, out, sl);                                                                        \
  EXB_END
int exb_host_grad(exb_model* m, const double* x, double* out) {
  const double* yy = nullptr;
  EXB_HOST_VEC(pl.m.nvar, whole(pl.m.nvar), exb_grad(m, m->dx, m->dout, m->hstream))
}
int exb_host_cons(exb_model* m, const double* x, double* out) {
  const double* yy = nullptr;
  EXB_HOST_VEC(pl.ncon, whole(pl.ncon), exb_cons(m, m->dx, m->dout, m->hstream))
}
int exb_host_jac(exb_model* m, const double* x, double* out) {
  const double* yy = nullptr;
  { EXB_BEGIN EXB_GUARD(m); bool done = false; int prc = host_coo_pipelined(m, KN_JAC, x, nullptr, 0.0, out, &done); if (prc || done) return prc; EXB_END }
  EXB_HOST_VEC(pl.nnzj, slices(m, 1), exb_jac(m, m->dx, m->dout, m->hstream))
}
int exb_host_hess(exb_model* m, const double* x, const double* y, double obj_weight, double* out) {
  const double* yy = y;
  { EXB_BEGIN EXB_GUARD(m); bool done = false; int prc = host_coo_pipelined(m, KN_HESS, x, y, obj_weight, out, &done); if (prc || done) return prc; EXB_END }
  EXB_HOST_VEC(pl.nnzh, slices(m, 2), exb_hess(m, m->dx, y ? m->dy : nullptr, obj_weight, m->dout, m->hstream))
}
// Pipelined form of exb_host_hess_compressed for page-locked caller buffers.  In a shift-indexed model a window of COLUMNS
// needs only the matching windows of x and y (plus a halo of a few entries), so everything streams: while window w's values
// travel to the host (second stream), window w + 1's x / y pieces go up and its tiles are evaluated -- H2D and D2H overlap
// (PCIe is full duplex) instead of following each other.
static int host_hessc_pipelined(exb_model* m, const double* x, const double* y, double sigma, double* out, bool* done) {
  *done = false;
  Launch& L = m->k[KN_HESSC];
  const exb::Plan& pl = m->plan->pl;
  static const int wenv = getenv("EXB_HOST_WINDOWS") ? atoi(getenv("EXB_HOST_WINDOWS")) : 8;
  if (!m->hess_tile || wenv < 2 || L.best < 0 || !L.fn || L.nblocks < 16) return EXB_OK;
  const ExbTile& T0 = L.tile;
  const long long o_lo = tile_before(T0, T0.c_lo), o_hi = tile_before(T0, T0.c_hi), total = tile_unique(m);
  if (o_hi - o_lo < (1LL << 18)) return EXB_OK;
  if (!is_pinned(x) || (y && !is_pinned(y)) || !is_pinned(out + o_lo)) return EXB_OK;
  const std::vector<int>& lst = m->plan->list(KN_HESSC);
  long long H = 0;   // how far from a window of columns the variables read by its tiles can lie
  for (int pi : lst) {
    const exb::PatternPlan& p = pl.pats[(size_t)pi];
    if (!p.xr_ok || p.xr_fixed) return EXB_OK;
    H = std::max<long long>(H, (p.t_cbmax - p.t_cbmin) + (p.rhi - p.rlo) + 1);
  }
  int rc = host_stream(m); if (rc) return rc;
  if (!m->hstream2) CU_TRY(m, cudaStreamCreateWithFlags(&m->hstream2, cudaStreamNonBlocking));
  const int W = (int)std::min<long long>(wenv, L.nblocks);
  while ((int)m->hev.size() < W) { cudaEvent_t e; CU_TRY(m, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); m->hev.push_back(e); }
  rc = ensure_host(m, &m->hx, &m->hx_n, &m->dx, &m->dx_n, 0, (size_t)pl.m.nvar); if (rc) return rc;
  if (y) { rc = ensure_host(m, &m->hy, &m->hy_n, &m->dy, &m->dy_n, 0, (size_t)pl.ncon); if (rc) return rc; }
  rc = ensure_host(m, &m->hout, &m->hout_n, &m->dout, &m->dout_n, 0, (size_t)total); if (rc) return rc;
  m->last_h2d = 0; m->last_d2h = 0;
  long long x_up = std::max<long long>(0, T0.c_lo - 1 - H);           // x[x_up_lo, x_up) is on the device so far (0-based)
  const long long x_first = x_up;
  std::vector<long long> y_up(lst.size(), -1);                        // per constraint pattern: rows [.., y_up) uploaded (0-based row numbers)
  (void)x_first;
  for (int w = 0; w < W; w++) {
    const long long b0 = (long long)L.nblocks * w / W, b1 = (long long)L.nblocks * (w + 1) / W;
    if (b1 <= b0) continue;
    const long long cw_lo = T0.c_lo + b0 * T0.T, cw_hi = std::min<long long>(T0.c_hi, T0.c_lo + b1 * T0.T);   // 1-based columns [cw_lo, cw_hi)
    const long long x_need = std::min<long long>(pl.m.nvar, cw_hi - 1 + H);
    if (x_need > x_up) {
      CU_TRY(m, cudaMemcpyAsync(m->dx + x_up, x + x_up, (size_t)(x_need - x_up) * 8, cudaMemcpyHostToDevice, m->hstream));
      m->last_h2d += (x_need - x_up) * 8; x_up = x_need;
    }
    if (y)
      for (size_t q = 0; q < lst.size(); q++) {
        const exb::PatternPlan& p = pl.pats[(size_t)lst[q]];
        if (p.ir.kind != exb::KIND_CON || p.ir.nitr == 0) continue;
        // points staged by the window's tiles: range values [cw_lo - cbmax, cw_hi - 1 - cbmin], point k = value - start, row o0 + k
        long long k0 = cw_lo - p.t_cbmax - p.ir.range_start, k1 = cw_hi - p.t_cbmin - p.ir.range_start;   // [k0, k1)
        k0 = std::max<long long>(0, k0); k1 = std::min<long long>(p.ir.nitr, k1);
        if (k1 <= k0) continue;
        long long r0 = p.o0 + k0, r1 = p.o0 + k1;
        if (y_up[q] > r0) r0 = y_up[q];
        if (r1 > r0) {
          CU_TRY(m, cudaMemcpyAsync(m->dy + r0, y + r0, (size_t)(r1 - r0) * 8, cudaMemcpyHostToDevice, m->hstream));
          m->last_h2d += (r1 - r0) * 8; y_up[q] = r1;
        }
      }
    {
      ExbGroup g = L.g;
      ExbCall cc{}; cc.x = m->dx; cc.y = y ? m->dy : nullptr; cc.th = m->d_theta; cc.sigma = sigma; cc.out = m->dout;
      ExbTile tt = T0; tt.c_lo = cw_lo; tt.c_hi = cw_hi;
      void* params[3] = {&g, &cc, &tt};
      CUresult r = g_drv.LaunchKernel(L.fn, (unsigned)(b1 - b0), 1, 1, (unsigned)pl.block, 1, 1, L.smem, (CUstream)m->hstream, params, nullptr);
      if (r != CUDA_SUCCESS) return fail(EXB_ERR_CUDA, std::string("windowed launch of ") + KNAME[KN_HESSC] + ": " + cu_err(r));
      m->launches++; m->last_launches++;
    }
    CU_TRY(m, cudaEventRecord(m->hev[(size_t)w], m->hstream));
    CU_TRY(m, cudaStreamWaitEvent(m->hstream2, m->hev[(size_t)w], 0));
    const long long q0 = tile_before(T0, cw_lo), q1 = tile_before(T0, cw_hi);
    if (q1 > q0) {
      CU_TRY(m, cudaMemcpyAsync(out + q0, m->dout + q0, (size_t)(q1 - q0) * 8, cudaMemcpyDeviceToHost, m->hstream2));
      m->last_d2h += (q1 - q0) * 8;
    }
  }
  CU_TRY(m, cudaStreamSynchronize(m->hstream2));
  CU_TRY(m, cudaStreamSynchronize(m->hstream));
  *done = true;
  return EXB_OK;
}

// duplicate-free forms with host buffers: the D2H copy carries the unique entries only (LV: 2N - 1 instead of 9N - 15 doubles)
int exb_host_jac_compressed(exb_model* m, const double* x, double* out) {
  const double* yy = nullptr;
  long long nj = 0;
  { EXB_BEGIN EXB_GUARD(m); int rc0 = ensure_sorted(m, 2, 1); if (rc0) return rc0; nj = m->jcmp.runs; EXB_END }
  EXB_HOST_VEC(nj, whole(nj), exb_jac_compressed(m, m->dx, m->dout, m->hstream))
}
int exb_host_hess_compressed(exb_model* m, const double* x, const double* y, double obj_weight, double* out) {
  const double* yy = y;
  { EXB_BEGIN EXB_GUARD(m); bool done = false; int prc = host_hessc_pipelined(m, x, y, obj_weight, out, &done); if (prc || done) return prc; EXB_END }
  int64_t sh[3] = {0, 0, 0};
  { int rc0 = exb_compressed_shard(m, sh); if (rc0) return rc0; }
  long long nh = 0;
  { int64_t t = 0; int rc0 = exb_compressed_dims(m, nullptr, &t); if (rc0) return rc0; nh = t; }
  std::vector<std::pair<long long, long long>> mine = {{(long long)sh[0], (long long)sh[1]}};
  EXB_HOST_VEC(nh, mine, exb_hess_compressed(m, m->dx, y ? m->dy : nullptr, obj_weight, m->dout, m->hstream))
}
// matrix-free products with host buffers: x (and y) go up through host_in, the multiplied vector through its own staging
static int host_vec_in(exb_model* m, const double* v, long long n, double** dv) {
  *dv = nullptr;
  CU_TRY(m, cudaMalloc((void**)dv, (size_t)(n > 0 ? n : 1) * 8));
  cudaError_t e = cudaMemcpyAsync(*dv, v, (size_t)n * 8, cudaMemcpyHostToDevice, m->hstream);
  if (e != cudaSuccess) { cudaFree(*dv); *dv = nullptr; return fail(EXB_ERR_CUDA, cudaGetErrorString(e)); }
  m->last_h2d += n * 8;
  return EXB_OK;
}
int exb_host_jprod(exb_model* m, const double* x, const double* v, double* out) {
  const double* yy = nullptr; double* dv = nullptr;
  { EXB_BEGIN EXB_GUARD(m); int r0 = host_stream(m); if (r0) return r0; r0 = host_vec_in(m, v, m->plan->pl.m.nvar, &dv); if (r0) return r0; EXB_END }
  struct Free { double* p; ~Free() { if (p) cudaFree(p); } } fr{dv};
  EXB_HOST_VEC(pl.ncon, whole(pl.ncon), (m->last_h2d += 0, exb_jprod(m, m->dx, dv, m->dout, m->hstream)))
}
int exb_host_jtprod(exb_model* m, const double* x, const double* v, double* out) {
  const double* yy = nullptr; double* dv = nullptr;
  { EXB_BEGIN EXB_GUARD(m); int r0 = host_stream(m); if (r0) return r0; r0 = host_vec_in(m, v, m->plan->pl.ncon, &dv); if (r0) return r0; EXB_END }
  struct Free { double* p; ~Free() { if (p) cudaFree(p); } } fr{dv};
  EXB_HOST_VEC(pl.m.nvar, whole(pl.m.nvar), exb_jtprod(m, m->dx, dv, m->dout, m->hstream))
}
int exb_host_hprod(exb_model* m, const double* x, const double* y, const double* v, double obj_weight, double* out) {
  const double* yy = y; double* dv = nullptr;
  { EXB_BEGIN EXB_GUARD(m); int r0 = host_stream(m); if (r0) return r0; r0 = host_vec_in(m, v, m->plan->pl.m.nvar, &dv); if (r0) return r0; EXB_END }
  struct Free { double* p; ~Free() { if (p) cudaFree(p); } } fr{dv};
  EXB_HOST_VEC(pl.m.nvar, whole(pl.m.nvar), exb_hprod(m, m->dx, y ? m->dy : nullptr, dv, obj_weight, m->dout, m->hstream))
}
}  // extern "C"
template <typename I>
static int host_structure_t(exb_model* m, int kn, long long n, I* rows, I* cols) {
  int rc = host_stream(m); if (rc) return rc;
  I *dr = nullptr, *dc = nullptr;
  CU_TRY(m, cudaMalloc((void**)&dr, (size_t)(n ? n : 1) * sizeof(I)));
  cudaError_t e = cudaMalloc((void**)&dc, (size_t)(n ? n : 1) * sizeof(I));
  if (e != cudaSuccess) { cudaFree(dr); return fail(EXB_ERR_CUDA, cudaGetErrorString(e)); }
  rc = structure(m, kn, dr, dc, m->hstream);
  if (!rc) {
    cudaMemcpyAsync(rows, dr, (size_t)n * sizeof(I), cudaMemcpyDeviceToHost, m->hstream);
    cudaMemcpyAsync(cols, dc, (size_t)n * sizeof(I), cudaMemcpyDeviceToHost, m->hstream);
    e = cudaStreamSynchronize(m->hstream);
    if (e != cudaSuccess) rc = fail(EXB_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(dr); cudaFree(dc);
  return rc;
}
extern "C" {
int exb_host_jac_structure32(exb_model* m, int32_t* rows, int32_t* cols) {
  EXB_BEGIN EXB_GUARD(m); return host_structure_t<int32_t>(m, KN_JSTRUCT32, m->plan->pl.nnzj, rows, cols); EXB_END
}
int exb_host_hess_structure32(exb_model* m, int32_t* rows, int32_t* cols) {
  EXB_BEGIN EXB_GUARD(m); return host_structure_t<int32_t>(m, KN_HSTRUCT32, m->plan->pl.nnzh, rows, cols); EXB_END
}
static int host_structure(exb_model* m, int kn, long long n, int64_t* rows, int64_t* cols) {
  int rc = host_stream(m); if (rc) return rc;
  long long *dr = nullptr, *dc = nullptr;
  CU_TRY(m, cudaMalloc((void**)&dr, (size_t)(n ? n : 1) * 8));
  cudaError_t e = cudaMalloc((void**)&dc, (size_t)(n ? n : 1) * 8);
  if (e != cudaSuccess) { cudaFree(dr); return fail(EXB_ERR_CUDA, cudaGetErrorString(e)); }
  rc = structure(m, kn, dr, dc, m->hstream);
  if (!rc) {
    cudaMemcpyAsync(rows, dr, (size_t)n * 8, cudaMemcpyDeviceToHost, m->hstream);
    cudaMemcpyAsync(cols, dc, (size_t)n * 8, cudaMemcpyDeviceToHost, m->hstream);
    e = cudaStreamSynchronize(m->hstream);
    if (e != cudaSuccess) rc = fail(EXB_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(dr); cudaFree(dc);
  return rc;
}
int exb_host_jac_structure64(exb_model* m, int64_t* rows, int64_t* cols) {
  EXB_BEGIN EXB_GUARD(m); return host_structure(m, KN_JSTRUCT64, m->plan->pl.nnzj, rows, cols); EXB_END
}
int exb_host_hess_structure64(exb_model* m, int64_t* rows, int64_t* cols) {
  EXB_BEGIN EXB_GUARD(m); return host_structure(m, KN_HSTRUCT64, m->plan->pl.nnzh, rows, cols); EXB_END
}

// Tune every kernel now, on scratch outputs: each value callback is called once (its first call is what ranks the variants).
// x: nvar doubles (device, or host when x_on_host; NULL: all ones); y: ncon device doubles (NULL: all ones).
static int tune_all(exb_model* m, const double* x, const double* y, bool x_on_host, void* stream) {
  const exb::Plan& pl = m->plan->pl;
  cudaStream_t st = (cudaStream_t)stream;
  double *dx = nullptr, *dy = nullptr, *g = nullptr, *c = nullptr, *jac = nullptr, *hess = nullptr, *obj = nullptr;
  auto n1 = [](long long n) { return (size_t)(n > 0 ? n : 1) * 8; };
  static const long long ONE = 0x3ff0000000000000LL;
  int rc = EXB_OK;
  cudaError_t e = cudaSuccess;
  auto al = [&](double** p, long long n) { if (e == cudaSuccess) e = cudaMalloc((void**)p, n1(n)); };
  al(&dx, pl.m.nvar); al(&dy, pl.ncon); al(&g, pl.m.nvar); al(&c, pl.ncon); al(&jac, pl.nnzj); al(&hess, pl.nnzh); al(&obj, 1);
  if (e == cudaSuccess) {
    if (x) e = cudaMemcpyAsync(dx, x, (size_t)pl.m.nvar * 8, x_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st);
    else e = exb_fx_fill((long long*)dx, pl.m.nvar, ONE, st);
  }
  if (e == cudaSuccess) {
    if (y) e = cudaMemcpyAsync(dy, y, (size_t)pl.ncon * 8, cudaMemcpyDeviceToDevice, st);
    else e = exb_fx_fill((long long*)dy, pl.ncon, ONE, st);
  }
  if (e != cudaSuccess) rc = fail(EXB_ERR_CUDA, std::string("exb_tune: ") + cudaGetErrorString(e));
  if (!rc) rc = exb_eval(m, EXB_EVAL_ALL, dx, dy, 1.0, obj, g, c, jac, hess, stream);
  if (!rc) rc = exb_obj_async(m, dx, obj, stream);
  if (!rc) rc = exb_grad(m, dx, g, stream);
  if (!rc) rc = exb_cons(m, dx, c, stream);
  if (!rc) rc = exb_jac(m, dx, jac, stream);
  if (!rc) rc = exb_hess(m, dx, dy, 1.0, hess, stream);
  if (!rc && m->hess_tile) rc = exb_hess_compressed(m, dx, dy, 1.0, hess, stream);
  if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = fail(EXB_ERR_CUDA, "exb_tune: kernel failed");
  for (double* p : {dx, dy, g, c, jac, hess, obj}) if (p) cudaFree(p);
  m->launches = 0; m->last_launches = 0;
  return rc;
}
int exb_tune(exb_model* m, const double* x, const double* y, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  return tune_all(m, x, y, false, stream);
  EXB_END
}
// out[0..4] = seconds spent in: planning + code generation, nvcc (0 when the module came from the cache), module load + data
// upload + build-time sorts, tuning so far, exb_create in total
int exb_build_info(const exb_model* m, double* out5) {
  if (!m || !out5) return fail(EXB_ERR_HANDLE, "invalid handle");
  out5[0] = m->plan_s; out5[1] = m->nvcc_s; out5[2] = m->load_s; out5[3] = m->tune_s; out5[4] = m->create_s;
  return EXB_OK;
}

// ---- multi-GPU: the communicator of a sharded model (include/exa_b200.h) ----------------------------------------------------
int exb_comm_unique_id(void* id128) {
  EXB_BEGIN
  if (!id128) return fail(EXB_ERR_ARG, "null argument");
  if (!load_nccl()) return fail(EXB_ERR_CUDA, "NCCL unavailable: " + g_nccl.why);
  ncclUniqueId id;
  NC_TRY(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(id) == EXB_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(id128, &id, sizeof id);
  return EXB_OK;
  EXB_END
}
int exb_comm_init(exb_model* m, const void* id128) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (!id128) return fail(EXB_ERR_ARG, "null argument");
  if (m->comm) return fail(EXB_ERR_ARG, "the handle already has a communicator");
  if (!load_nccl()) return fail(EXB_ERR_CUDA, "NCCL unavailable: " + g_nccl.why);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NC_TRY(g_nccl.CommInitRank(&m->comm, m->world, id, m->rank));
  m->comm_owned = true;
  return EXB_OK;
  EXB_END
}
int exb_comm_attach(exb_model* m, void* nccl_comm) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (!nccl_comm) return fail(EXB_ERR_ARG, "null communicator");
  if (m->comm) return fail(EXB_ERR_ARG, "the handle already has a communicator");
  if (!load_nccl()) return fail(EXB_ERR_CUDA, "NCCL unavailable: " + g_nccl.why);
  int n = 0, r = -1;
  NC_TRY(g_nccl.CommCount((ncclComm_t)nccl_comm, &n));
  NC_TRY(g_nccl.CommUserRank((ncclComm_t)nccl_comm, &r));
  if (n != m->world || r != m->rank) return fail(EXB_ERR_ARG, "communicator rank / size differ from the handle's shard");
  m->comm = (ncclComm_t)nccl_comm; m->comm_owned = false;
  return EXB_OK;
  EXB_END
}
int exb_comm_destroy(exb_model* m) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (m->comm && m->comm_owned) g_nccl.CommDestroy(m->comm);
  m->comm = nullptr; m->comm_owned = false;
  return EXB_OK;
  EXB_END
}
int exb_comm_set_mode(exb_model* m, int mode) {
  if (!m) return fail(EXB_ERR_HANDLE, "invalid handle");
  if (mode != EXB_COMM_REPLICATE && mode != EXB_COMM_OWNER) return fail(EXB_ERR_ARG, "unknown mode");
  m->comm_mode = mode;
  return EXB_OK;
}
// replicate the sharded COO values (which = 1: jac, 2: hess): every rank broadcasts the slices it wrote
int exb_comm_gather_coo(exb_model* m, int which, double* vals, void* stream) {
  EXB_BEGIN
  EXB_GUARD(m);
  if (which != 1 && which != 2) return fail(EXB_ERR_ARG, "which must be 1 (jac) or 2 (hess)");
  if (m->world == 1) return EXB_OK;
  if (!m->comm) return fail(EXB_ERR_ARG, "no communicator: call exb_comm_init first");
  const exb::Plan& pl = m->plan->pl;
  std::vector<Seg> segs;
  for (auto& p : pl.pats) {
    if (which == 1 && p.ir.kind == exb::KIND_OBJ) continue;
    const long long o = which == 1 ? p.o1 : p.o2, stp = which == 1 ? p.o1step : p.o2step;
    for (int r = 0; r < m->world; r++) segs.push_back({r, o + stp * (p.ir.nitr * r / m->world), o + stp * (p.ir.nitr * (r + 1) / m->world)});
  }
  return comm_allgatherv(m, vals, segs, (cudaStream_t)stream);
  EXB_END
}
int exb_owned(const exb_model* m, int64_t* o) {
  if (!m || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  o[0] = m->v_lo; o[1] = m->v_hi;
  return EXB_OK;
}

int exb_shard(const exb_model* m, int k, int64_t* o) {
  if (!m || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  if (k < 0 || (size_t)k >= m->plan->pl.pats.size()) return fail(EXB_ERR_ARG, "pattern index out of range");
  const exb::PatternPlan& p = m->plan->pl.pats[(size_t)k];
  o[0] = m->lo[(size_t)k]; o[1] = m->hi[(size_t)k];
  const bool in_jac = p.ir.kind != exb::KIND_OBJ;
  o[2] = in_jac ? p.o1 + o[0] * p.o1step : 0; o[3] = in_jac ? p.o1 + o[1] * p.o1step : 0;
  o[4] = p.o2 + o[0] * p.o2step; o[5] = p.o2 + o[1] * p.o2step;
  return EXB_OK;
}
int exb_set_timing(exb_model* m, int on) {
  if (!m) return fail(EXB_ERR_HANDLE, "invalid handle");
  m->timing = on != 0;
  return EXB_OK;
}
// ms8 / calls8: accumulated device milliseconds and call counts for obj grad cons jac hess jprod jtprod hprod
int exb_timings(exb_model* m, double* ms8, int64_t* calls8, int reset) {
  EXB_BEGIN
  if (!m) return fail(EXB_ERR_HANDLE, "invalid handle");
  DeviceGuard dg(m->device);
  for (auto& p : m->pending) {
    float t = 0;
    if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&t, p.e0, p.e1) == cudaSuccess) {
      m->cb_ms[p.cb] += t; m->cb_calls[p.cb]++;
    } else cudaGetLastError();
    cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
  }
  m->pending.clear();
  for (int k = 0; k < 8; k++) {
    if (ms8) ms8[k] = m->cb_ms[k];
    if (calls8) calls8[k] = m->cb_calls[k];
    if (reset) { m->cb_ms[k] = 0; m->cb_calls[k] = 0; }
  }
  return EXB_OK;
  EXB_END
}
int exb_host_bytes(const exb_model* m, int64_t* o) {
  if (!m || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  o[0] = m->last_h2d; o[1] = m->last_d2h;
  return EXB_OK;
}
int exb_kernel_choice(const exb_model* m, int callback, int64_t* o) {
  if (!m || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  static const int map[5] = {KN_OBJ, KN_GGRAD, KN_CONS, KN_JAC, KN_HESS};
  if (callback < 0 || callback > 4) return fail(EXB_ERR_ARG, "callback must be 0 (obj) .. 4 (hess)");
  int kn = map[callback];
  if (kn == KN_GGRAD && m->k[KN_GRADT].nblocks > 0) kn = KN_GRADT;
  if (kn == KN_GGRAD && m->k[kn].nblocks == 0) kn = KN_SGRAD;
  const Launch& L = m->k[kn];
  o[0] = L.best >= 0 && (size_t)L.best < m->plan->var.size() ? m->plan->var[(size_t)L.best].minb : -1;
  o[1] = L.use_p >= 0 ? 1 : 0;
  o[2] = L.use_p >= 0 ? (int64_t)L.pgrid[(size_t)L.use_p] : (int64_t)L.nblocks;
  o[3] = kn == KN_GGRAD ? 1 : kn == KN_GRADT ? 2 : 0;
  return EXB_OK;
}
int exb_comm_stats(const exb_model* m, int64_t* o) {
  if (!m || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  o[0] = m->collectives; o[1] = m->last_collectives; o[2] = m->comm ? 1 : 0; o[3] = m->comm_mode;
  return EXB_OK;
}
int exb_stats(const exb_model* m, int64_t* o) {
  if (!m || !o) return fail(EXB_ERR_HANDLE, "invalid handle");
  o[0] = m->launches; o[1] = m->last_launches; o[2] = (int64_t)m->dev_bytes; o[3] = m->plan->from_cache ? 1 : 0;
  return EXB_OK;
}

}  // extern "C"
