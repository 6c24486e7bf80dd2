// context.cu — context lifetime, error reporting and defaults of the C ABI (include/hslam_b200.h).
#include "common.cuh"

static thread_local std::string g_last_error;

namespace hs {
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace hs

extern "C" {

const char* hslam_last_error(void) { return g_last_error.c_str(); }
int hslam_abi_version(void) { return HSLAM_ABI_VERSION; }

void hslam_params_default(hslam_params* p) {
  if (!p) return;
  p->huberTH = 9.0f;                 // Src/Settings.cpp:68
  p->coarseCutoffTH = 20.0f;         // :77
  p->outlierTHSumComponent = 2500;   // :64
  p->affineOptModeA = 1e12f;         // :109
  p->affineOptModeB = 1e8f;          // :110
  p->scale_xi_rot = 1.0f;            // Include/GlobalTypes.h:34-41
  p->scale_xi_trans = 0.5f;
  p->scale_a = 10.0f;
  p->scale_b = 1000.0f;
  p->scale_f = 50.0f;
  p->scale_c = 50.0f;
  p->scale_idepth = 1.0f;
  const int it[5] = {10, 20, 50, 50, 50};  // Src/CoarseTracker.cpp:518
  for (int i = 0; i < 5; ++i) p->max_tracker_iterations[i] = it[i];
  p->frameEnergyTHN = 0.7f;            // Src/Settings.cpp:73
  p->frameEnergyTHFacMedian = 1.5f;    // :74
  p->frameEnergyTHConstWeight = 0.5f;  // :75
  p->overallEnergyTHWeight = 1.0f;     // :66
}

int hslam_create(const hslam_params* params, int device, hslam_ctx** out) {
  HS_ARG(out, "hslam_create: null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    // No CPU fallback by design: the caller must see this.
    hs::set_error(std::string("hslam_create: no usable CUDA device (") + cudaGetErrorString(e) + ")");
    return HSLAM_ERR_CUDA;
  }
  HS_ARG(device >= 0 && device < ndev, "hslam_create: device index out of range");
  HS_CUDA(cudaSetDevice(device));
  hslam_ctx* c = new hslam_ctx();
  c->device = device;
  if (params) c->prm = *params; else hslam_params_default(&c->prm);
  cudaDeviceProp prop;
  HS_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (!prop.cooperativeLaunch) {
    delete c;
    hs::set_error("hslam_create: device lacks cooperative launch");
    return HSLAM_ERR_CUDA;
  }
  HS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // the side stream carries result export + D2H copies that should overlap the main stream's kernels: highest priority
    int lo = 0, hi = 0;
    HS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    HS_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
  }
  HS_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  HS_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  *out = c;
  return 0;
}

int hslam_destroy(hslam_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : c->prof_pool) cudaEventDestroy(e);
  for (auto& b : c->host_blocks) cudaFreeHost(b.p);
  if (c->trace_dev) cudaFree(c->trace_dev);
  if (c->trace_host) cudaFreeHost(c->trace_host);
  cudaStreamSynchronize(c->side);
  cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
  cudaStreamDestroy(c->side);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int hslam_synchronize(hslam_ctx* c) {
  HS_ARG(c, "hslam_synchronize: null");
  HS_CUDA(cudaSetDevice(c->device));
  HS_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int hslam_get_stream(hslam_ctx* c, void** s) {
  HS_ARG(c && s, "hslam_get_stream: null");
  *s = (void*)c->stream;
  return 0;
}

int64_t hslam_launch_count(hslam_ctx* c) { return c ? c->launches : -1; }

int hslam_host_alloc(hslam_ctx* c, size_t bytes, void** out) {
  HS_ARG(c && out && bytes > 0, "hslam_host_alloc: bad argument");
  HS_CUDA(cudaSetDevice(c->device));
  char* p = nullptr;
  HS_CUDA(cudaMallocHost((void**)&p, bytes));
  c->host_blocks.push_back({p, bytes});
  *out = p;
  return 0;
}

int hslam_host_free(hslam_ctx* c, void* p) {
  HS_ARG(c, "hslam_host_free: null context");
  if (!p) return 0;
  for (size_t i = 0; i < c->host_blocks.size(); ++i)
    if (c->host_blocks[i].p == (char*)p) {
      cudaStreamSynchronize(c->stream);
      cudaFreeHost(p);
      c->host_blocks.erase(c->host_blocks.begin() + i);
      return 0;
    }
  hs::set_error("hslam_host_free: pointer was not returned by hslam_host_alloc");
  return HSLAM_ERR_ARG;
}

static const char* k_names[HSLAM_K_COUNT] = {"pyramid_level", "trk_eval", "trk_track", "ba_linearize", "ba_apply",
                                             "ba_top", "ba_fold_top", "ba_point_sum", "ba_schur", "ba_fold_sc",
                                             "ba_stitch", "ba_resub", "ba_fixlin", "unpack_level", "ba_energy", "ba_layout", "trace", "coarse_depth", "ba_energy_th"};
const char* hslam_kernel_name(int k) { return (k >= 0 && k < HSLAM_K_COUNT) ? k_names[k] : ""; }

int hslam_profile_enable(hslam_ctx* c, int enable) {
  HS_ARG(c, "hslam_profile_enable: null");
  c->profile = enable != 0;
  return 0;
}

int hslam_profile_read(hslam_ctx* c, float ms_out[HSLAM_K_COUNT], int count_out[HSLAM_K_COUNT]) {
  HS_ARG(c && ms_out && count_out, "hslam_profile_read: null");
  HS_CUDA(cudaSetDevice(c->device));
  HS_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < HSLAM_K_COUNT; ++k) { ms_out[k] = 0.f; count_out[k] = 0; }
  for (auto& r : c->prof_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    ms_out[r.k] += ms; count_out[r.k]++;
    c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
  }
  c->prof_recs.clear();
  return 0;
}

}  // extern "C"
