// tinyfusers_b200 — process-lifetime state of the C-ABI library: error string, device probe,
// TMA descriptor encoding through the driver entry point, launch bookkeeping.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "tf_common.cuh"
#include "tinyfusers_b200.h"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int g_sms = 0;

void tf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* tf_last_error(void) { return g_err; }

extern "C" int tf_version(void) { return 100; }

int tf_launch_count_add(int n) {
  g_launches += n;
  return 0;
}
extern "C" long long tf_launch_count(void) { return g_launches.load(); }
extern "C" void tf_launch_count_reset(void) { g_launches = 0; }

static int g_pdl = -1;
int tf_pdl_enabled() {
  if (g_pdl < 0) {
    // Programmatic dependent launch between this library's kernels: every kernel triggers its dependents on entry and
    // waits (griddepcontrol.wait) before it touches a producer's output, so the next kernel's launch latency, barrier /
    // TMEM setup and parameter prefetches overlap the previous kernel. Measured on B200 inside the captured UNet step:
    // neutral early in round 1 (long kernels), 237.2 -> 244.9 steps/s on the final round-1 build (416 short launches).
    // TINYFUSERS_B200_PDL=0 turns it off.
    const char* e = getenv("TINYFUSERS_B200_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl;
}
extern "C" int tf_set_pdl(int enable) {
  g_pdl = enable ? 1 : 0;
  return TF_OK;
}

int tf_num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

extern "C" int tf_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    tf_set_error("tf_init: no CUDA device visible (%s); this library has no CPU fallback",
                 cudaGetErrorString(e));
    return TF_ERR_DEVICE;
  }
  TF_CHECK_ARG(device >= 0 && device < count, "tf_init: device %d out of range (%d visible)", device, count);
  // One process per GPU: function attributes (opt-in shared memory sizes), the SM count and the tuning
  // table are set once per process, for the device of the first tf_init. A second device in the same process is refused
  // loudly instead of failing later with "invalid argument" on the first > 48 KB launch.
  static int g_device = -1;
  if (g_device >= 0 && g_device != device) {
    tf_set_error("tf_init: this process already serves device %d; tinyfusers_b200 runs one process per GPU (got device %d)",
                 g_device, device);
    return TF_ERR_DEVICE;
  }
  TF_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0;
  TF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  TF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    tf_set_error("tf_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                 major, minor);
    return TF_ERR_DEVICE;
  }
  g_sms = 0;
  tf_num_sms();
  g_device = device;
  return TF_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int tf_encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base,
                   const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                   const uint32_t* elem_strides, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    tf_set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old or no GPU)");
    return TF_ERR_DEVICE;
  }
  cuuint64_t d[5] = {0, 0, 0, 0, 0}, s[4] = {0, 0, 0, 0};
  cuuint32_t b[5] = {0, 0, 0, 0, 0}, e[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tf_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
                 (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
                 (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
                 rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
    return TF_ERR_ARG;
  }
  return TF_OK;
}


// ---- CUDA-graph capture without a host framework ----------------------------------------------------------------------
// The package captures the denoising step with torch.cuda.CUDAGraph (torch is its container library). A host that binds this
// library from something else (the reference's own CuPy / ctypes world: tinyfusers/native/cudart/ops.py) gets the same
// capture / replay through these four calls: every tf_* launch is graph-capturable (no hidden synchronisation, programmatic
// dependent launch attributes are recorded as graph edges).
extern "C" int tf_graph_begin_capture(void* stream) {
  TF_CUDA(cudaStreamBeginCapture(reinterpret_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal));
  return TF_OK;
}

extern "C" int tf_graph_end_capture(void* stream, void** graph_exec_out) {
  TF_CHECK_ARG(graph_exec_out != nullptr, "tf_graph_end_capture: null output");
  cudaGraph_t graph = nullptr;
  TF_CUDA(cudaStreamEndCapture(reinterpret_cast<cudaStream_t>(stream), &graph));
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    tf_set_error("tf_graph_end_capture: cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  *graph_exec_out = exec;
  return TF_OK;
}

extern "C" int tf_graph_launch(void* graph_exec, void* stream) {
  TF_CHECK_ARG(graph_exec != nullptr, "tf_graph_launch: null graph");
  TF_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), reinterpret_cast<cudaStream_t>(stream)));
  return TF_OK;
}

extern "C" int tf_graph_destroy(void* graph_exec) {
  if (graph_exec) TF_CUDA(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
  return TF_OK;
}
