// pyramid.cu — Frame::CreateDirPyrs (reference Src/Frame.cpp:104-181) on the device.
//
// One launch per level.  Level 0 reads the fp32 image; level l>0 recomputes the 2x2 box means of
// level l-1 for the five pixels a gradient needs (the re-reads hit L1/L2), so there is no
// intermediate intensity plane and no second pass.  Output: float4 {I, dx, dy, absSquaredGrad} per
// pixel — a texel is ONE 16-byte load in the gather kernels instead of three 4-byte loads of the
// reference's 12-byte Vec3f.
// HBM-bound streaming: algorithmic bytes per level-0 pixel = 4 (read) + 16 (write).
#include "common.cuh"
#include <cstring>

namespace hs {

template <bool LEVEL0>
__device__ __forceinline__ float intensity_at(const float* __restrict__ img, const float4* __restrict__ prev,
                                              int idx, int wl, int wprev) {
  if (LEVEL0) return __ldg(img + idx);
  int x = idx % wl, y = idx / wl;
  const float4* p = prev + 2 * x + 2 * y * wprev;
  // 0.25f * (((a + b) + c) + d), Frame.cpp:137-140
  float s = xadd(xadd(xadd(__ldg(p).x, __ldg(p + 1).x), __ldg(p + wprev).x), __ldg(p + wprev + 1).x);
  return xmul(0.25f, s);
}

template <bool LEVEL0>
__global__ void __launch_bounds__(256) pyramid_level_kernel(const float* __restrict__ img,
                                                            const float4* __restrict__ prev,
                                                            float4* __restrict__ out, int wl, int hl, int wprev,
                                                            const float* __restrict__ gamma) {
  int n = wl * hl;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    float I = intensity_at<LEVEL0>(img, prev, idx, wl, wprev);
    float dx = 0.f, dy = 0.f, ag = 0.f;
    if (idx >= wl && idx < wl * (hl - 1)) {  // Frame.cpp:143-165: flat index, row ends wrap
      float ip = intensity_at<LEVEL0>(img, prev, idx + 1, wl, wprev);
      float im = intensity_at<LEVEL0>(img, prev, idx - 1, wl, wprev);
      float jp = intensity_at<LEVEL0>(img, prev, idx + wl, wl, wprev);
      float jm = intensity_at<LEVEL0>(img, prev, idx - wl, wl, wprev);
      dx = xmul(0.5f, xsub(ip, im));
      dy = xmul(0.5f, xsub(jp, jm));
      if (!isfinite(dx)) dx = 0.f;
      if (!isfinite(dy)) dy = 0.f;
      ag = xadd(xmul(dx, dx), xmul(dy, dy));
      if (gamma) {  // PhotometricUndistorter::getBGradOnly (Include/photometricUndistorter.h:24-31)
        int c = (int)xadd(I, 0.5f);
        c = c < 5 ? 5 : (c > 250 ? 250 : c);
        float gw = __ldg(gamma + c);
        ag = xmul(ag, xmul(gw, gw));
      }
    }
    out[idx] = make_float4(I, dx, dy, ag);
  }
}

// level-0 pair image: per pixel {I,dx,dy}(x) | {I,dx,dy}(x+1) | 2 pads  (the last column repeats itself; never sampled)
__global__ void pair_image_kernel(const float4* __restrict__ in, float4* __restrict__ out, int w, int h) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * h; i += gridDim.x * blockDim.x) {
    const float4 a = in[i];
    const float4 b = ((i % w) + 1 < w) ? in[i + 1] : a;
    out[2 * (size_t)i] = make_float4(a.x, a.y, a.z, b.x);
    out[2 * (size_t)i + 1] = make_float4(b.y, b.z, 0.f, 0.f);
  }
}

int build_pyramid(hslam_frame* f, const float* img_dev) {
  hslam_ctx* ctx = f->ctx;
  for (int l = 0; l < f->levels; ++l) {
    int wl = f->w >> l, hl = f->h >> l, n = wl * hl;
    int blocks = (n + 255) / 256;
    int cap = ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (l == 0)
      HS_LAUNCH(ctx, HSLAM_K_PYRAMID, (pyramid_level_kernel<true><<<blocks, 256, 0, ctx->stream>>>(
                                          img_dev, nullptr, f->dI[0], wl, hl, 0, f->gamma)));
    else
      HS_LAUNCH(ctx, HSLAM_K_PYRAMID, (pyramid_level_kernel<false><<<blocks, 256, 0, ctx->stream>>>(
                                          nullptr, f->dI[l - 1], f->dI[l], wl, hl, f->w >> (l - 1), f->gamma)));
  }
  if (f->pair0)   // a BA window uses this frame: keep its pair image in step with the pyramid
    HS_LAUNCH(ctx, HSLAM_K_PYRAMID, (pair_image_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(f->dI[0], f->pair0, f->w, f->h)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

__global__ void unpack_level_kernel(const float4* __restrict__ in, float* __restrict__ dI3, float* __restrict__ ag,
                                    int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 v = in[i];
    if (dI3) { dI3[3 * i] = v.x; dI3[3 * i + 1] = v.y; dI3[3 * i + 2] = v.z; }
    if (ag) ag[i] = v.w;
  }
}

}  // namespace hs

using namespace hs;

extern "C" {

static int frame_alloc(hslam_ctx* ctx, int w, int h, int levels, const float* gamma_host, hslam_frame** out) {
  HS_ARG(ctx && out, "hslam_frame_create: null argument");
  HS_ARG(w >= 16 && h >= 16 && levels >= 1 && levels <= HSLAM_MAX_PYR_LEVELS, "hslam_frame_create: bad size/levels");
  HS_ARG(((w >> (levels - 1)) >= 4) && ((h >> (levels - 1)) >= 4), "hslam_frame_create: too many levels for this size");
  HS_CUDA(cudaSetDevice(ctx->device));
  hslam_frame* f = new hslam_frame();
  f->ctx = ctx; f->w = w; f->h = h; f->levels = levels;
  size_t tot = 0;
  for (int l = 0; l < levels; ++l) tot += (size_t)(w >> l) * (h >> l);
  HS_CUDA(cudaMalloc(&f->base, tot * sizeof(float4)));
  size_t off = 0;
  for (int l = 0; l < levels; ++l) { f->dI[l] = f->base + off; off += (size_t)(w >> l) * (h >> l); }
  HS_CUDA(cudaMalloc(&f->staging, (size_t)w * h * sizeof(float)));
  HS_CUDA(cudaMallocHost(&f->pinned, (size_t)w * h * sizeof(float)));
  HS_CUDA(cudaEventCreateWithFlags(&f->h2d_done, cudaEventDisableTiming));
  if (gamma_host) {
    HS_CUDA(cudaMalloc(&f->gamma, 256 * sizeof(float)));
    HS_CUDA(cudaMemcpyAsync(f->gamma, gamma_host, 256 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  }
  *out = f;
  return 0;
}

int hslam_frame_update(hslam_frame* f, const float* img, int img_is_device) {
  HS_ARG(f && img, "hslam_frame_update: null argument");
  hslam_ctx* ctx = f->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const float* src = img;
  if (!img_is_device) {
    // pageable -> pinned -> device keeps the copy asynchronous w.r.t. the host
    HS_CUDA(cudaEventSynchronize(f->h2d_done));  // previous copy out of `pinned` finished
    memcpy(f->pinned, img, (size_t)f->w * f->h * sizeof(float));
    HS_CUDA(cudaMemcpyAsync(f->staging, f->pinned, (size_t)f->w * f->h * sizeof(float), cudaMemcpyHostToDevice,
                            ctx->stream));
    HS_CUDA(cudaEventRecord(f->h2d_done, ctx->stream));
    src = f->staging;
  }
  return build_pyramid(f, src);
}

int hslam_frame_create(hslam_ctx* ctx, const float* img_host, int w, int h, int levels, const float* gamma,
                       hslam_frame** out) {
  HS_ARG(img_host, "hslam_frame_create: null image");
  int rc = frame_alloc(ctx, w, h, levels, gamma, out);
  if (rc) return rc;
  return hslam_frame_update(*out, img_host, 0);
}

int hslam_frame_create_from_device(hslam_ctx* ctx, const float* img_dev, int w, int h, int levels, const float* gamma,
                                   hslam_frame** out) {
  HS_ARG(img_dev, "hslam_frame_create_from_device: null image");
  int rc = frame_alloc(ctx, w, h, levels, gamma, out);
  if (rc) return rc;
  return hslam_frame_update(*out, img_dev, 1);
}

}  // extern "C"
// Internal: make sure the frame has its level-0 pair image (allocated and built on the frame's context stream).
int hs_frame_ensure_pair(hslam_frame* f) {
  if (f->pair0) return 0;
  hslam_ctx* ctx = f->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_CUDA(cudaMalloc(&f->pair0, (size_t)f->w * f->h * 2 * sizeof(float4)));
  HS_LAUNCH(ctx, HSLAM_K_PYRAMID, (pair_image_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(f->dI[0], f->pair0, f->w, f->h)));
  HS_CUDA(cudaGetLastError());
  return 0;
}
extern "C" {

int hslam_frame_level_size(hslam_frame* f, int lvl, int* w, int* h) {
  HS_ARG(f && lvl >= 0 && lvl < f->levels, "hslam_frame_level_size: bad level");
  if (w) *w = f->w >> lvl;
  if (h) *h = f->h >> lvl;
  return 0;
}

int hslam_frame_download(hslam_frame* f, int lvl, float* dI_xyz_host, float* absSqGrad_host) {
  HS_ARG(f && lvl >= 0 && lvl < f->levels, "hslam_frame_download: bad level");
  hslam_ctx* ctx = f->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  int n = (f->w >> lvl) * (f->h >> lvl);
  float *d3 = nullptr, *da = nullptr;
  if (dI_xyz_host) HS_CUDA(cudaMalloc(&d3, (size_t)n * 3 * sizeof(float)));
  if (absSqGrad_host) HS_CUDA(cudaMalloc(&da, (size_t)n * sizeof(float)));
  HS_LAUNCH(ctx, HSLAM_K_UNPACK, (unpack_level_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(f->dI[lvl], d3, da, n)));
  if (d3) HS_CUDA(cudaMemcpyAsync(dI_xyz_host, d3, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (da) HS_CUDA(cudaMemcpyAsync(absSqGrad_host, da, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  if (d3) cudaFree(d3);
  if (da) cudaFree(da);
  return 0;
}

int hslam_frame_destroy(hslam_frame* f) {
  if (!f) return 0;
  cudaSetDevice(f->ctx->device);
  cudaStreamSynchronize(f->ctx->stream);
  cudaFree(f->base);
  if (f->pair0) cudaFree(f->pair0);
  cudaFree(f->staging);
  cudaFree(f->gamma);
  cudaFreeHost(f->pinned);
  if (f->h2d_done) cudaEventDestroy(f->h2d_done);
  delete f;
  return 0;
}

}  // extern "C"
