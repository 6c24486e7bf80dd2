// common.cuh — shared pieces of the sm_100a hot-path library (product code).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/hslam_b200.h"

namespace hs {

void set_error(const std::string& msg);

#define HS_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      hs::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +  \
                    std::to_string(__LINE__) + ")");                                               \
      return HSLAM_ERR_CUDA;                                                                       \
    }                                                                                              \
  } while (0)

#define HS_ARG(cond, msg)                 \
  do {                                    \
    if (!(cond)) {                        \
      hs::set_error(std::string(msg));    \
      return HSLAM_ERR_ARG;               \
    }                                     \
  } while (0)

// ---- exactly-rounded fp32 primitives -------------------------------------------------
// Everything that decides a mask (projection, bounds gates, bilinear weights, residual,
// Huber weight, energy) is written with these so that ptxas can never contract a*b+c into an
// FMA: the CPU oracle is compiled with -ffp-contract=off and the two must agree bit for bit.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
// a0 + (a1 + a2): Eigen 3.3 redux order for a 3-term coefficient product (see oracle_math.hpp)
__device__ __forceinline__ float xdot3(float m0, float m1, float m2, float x, float y, float z) {
  return xadd(xmul(m0, x), xadd(xmul(m1, y), xmul(m2, z)));
}

// getInterpolatedElement33 (Include/GlobalTypes.h:377-388) on the float4 {I,dx,dy,absgrad} image.
// Four 16-byte texel loads through the read-only path.
struct Hit3 { float I, dx, dy; };
// L2 "evict_last" policy for the pyramid texels: the images are re-read by every residual of the window
// (and by every LM iteration of the tracker) while the Jacobian streams are written once, so the images
// are the lines worth keeping in L2.
__device__ __forceinline__ uint64_t l2_keep_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_keep(const float4* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// 256-bit load (sm_100: LDG.E.256) of one record of the level-0 PAIR image: {I,dx,dy}(x,y) | {I,dx,dy}(x+1,y) | 2 pads.
struct Pair8 { float4 a, b; };
__device__ __forceinline__ Pair8 ldg_pair_keep(const float4* p, uint64_t pol) {
  Pair8 v;
  asm("ld.global.nc.L2::cache_hint.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
      : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
      : "l"(p), "l"(pol));
  return v;
}
// 256-bit store (sm_100: STG.E.256): one 32-byte record = one full sector in one instruction
__device__ __forceinline__ void stg256(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
#ifndef HS_TEXEL_L2_KEEP
#define HS_TEXEL_L2_KEEP 1
#endif
template <bool KEEP = false>
__device__ __forceinline__ Hit3 interp33(const float4* __restrict__ img, float x, float y, int width, uint64_t pol = 0) {
  int ix = (int)x, iy = (int)y;
  float dx = xsub(x, (float)ix), dy = xsub(y, (float)iy);
  float dxdy = xmul(dx, dy);
  const float4* bp = img + ix + iy * width;
#if HS_TEXEL_L2_KEEP
  float4 tl, tr, bl, br;
  if (KEEP) { tl = ldg_keep(bp, pol); tr = ldg_keep(bp + 1, pol); bl = ldg_keep(bp + width, pol); br = ldg_keep(bp + width + 1, pol); }
  else { tl = __ldg(bp); tr = __ldg(bp + 1); bl = __ldg(bp + width); br = __ldg(bp + width + 1); }
#else
  float4 tl = __ldg(bp), tr = __ldg(bp + 1), bl = __ldg(bp + width), br = __ldg(bp + width + 1);
#endif
  float wBR = dxdy, wBL = xsub(dy, dxdy), wTR = xsub(dx, dxdy);
  float wTL = xadd(xsub(xsub(1.0f, dx), dy), dxdy);
  Hit3 h;
  h.I = xadd(xadd(xadd(xmul(wBR, br.x), xmul(wBL, bl.x)), xmul(wTR, tr.x)), xmul(wTL, tl.x));
  h.dx = xadd(xadd(xadd(xmul(wBR, br.y), xmul(wBL, bl.y)), xmul(wTR, tr.y)), xmul(wTL, tl.y));
  h.dy = xadd(xadd(xadd(xmul(wBR, br.z), xmul(wBL, bl.z)), xmul(wTR, tr.z)), xmul(wTL, tl.z));
  return h;
}

// The same interpolation on the PAIR image (32 bytes per pixel, see hslam_frame::pair0): the two texels of a row arrive
// in ONE 256-bit load, so a bilinear tap is two loads instead of four (half the load instructions and L1 wavefronts of the
// gather, every load exactly one 32-byte sector).  Same arithmetic, bit-identical result.
__device__ __forceinline__ Hit3 interp33_pair(const float4* __restrict__ pimg, float x, float y, int width, uint64_t pol) {
  int ix = (int)x, iy = (int)y;
  float dx = xsub(x, (float)ix), dy = xsub(y, (float)iy);
  float dxdy = xmul(dx, dy);
  const float4* bp = pimg + 2 * ((size_t)ix + (size_t)iy * width);
  const Pair8 top = ldg_pair_keep(bp, pol), bot = ldg_pair_keep(bp + 2 * (size_t)width, pol);
  float wBR = dxdy, wBL = xsub(dy, dxdy), wTR = xsub(dx, dxdy);
  float wTL = xadd(xsub(xsub(1.0f, dx), dy), dxdy);
  Hit3 h;
  h.I = xadd(xadd(xadd(xmul(wBR, bot.a.w), xmul(wBL, bot.a.x)), xmul(wTR, top.a.w)), xmul(wTL, top.a.x));
  h.dx = xadd(xadd(xadd(xmul(wBR, bot.b.x), xmul(wBL, bot.a.y)), xmul(wTR, top.b.x)), xmul(wTL, top.a.y));
  h.dy = xadd(xadd(xadd(xmul(wBR, bot.b.y), xmul(wBL, bot.a.z)), xmul(wTR, top.b.y)), xmul(wTL, top.a.z));
  return h;
}

// Fixed-shape warp tree: lane 0 ends with the sum; the association is the same on every run.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hs

struct hslam_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;        // copy/export stream that overlaps D2H with the accumulate kernels
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  hslam_params prm;
  int sm_count = 148;
  int64_t launches = 0;
  // cudaFuncAttributeMaxDynamicSharedMemorySize is per device: remembered per context, not per process
  bool top_attr_done = false;
  size_t schur_smem_configured = 48 * 1024;
  // optional per-kernel event timing (hslam_profile_*)
  bool profile = false;
  struct ProfRec { int k; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  // staging of hslam_trace_immature
  char* trace_dev = nullptr; char* trace_host = nullptr; size_t trace_cap = 0;
  // pinned host blocks handed out by hslam_host_alloc (copies to/from them are direct DMA)
  struct HostBlock { char* p; size_t bytes; };
  std::vector<HostBlock> host_blocks;
  bool is_pinned(const void* q, size_t bytes) const {
    const char* c = (const char*)q;
    for (const auto& b : host_blocks) if (c >= b.p && c + bytes <= b.p + b.bytes) return true;
    return false;
  }
};

namespace hs {
struct ProfScope {   // brackets one launch with events when profiling is on; always counts the launch
  hslam_ctx* c; int k; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(hslam_ctx* ctx, int kid) : c(ctx), k(kid) {
    c->launches++;
    if (!c->profile) return;
    auto get = [&]() { cudaEvent_t e; if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); } else cudaEventCreate(&e); return e; };
    a = get(); b = get();
    cudaEventRecord(a, c->stream);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, c->stream);
    c->prof_recs.push_back({k, a, b});
  }
};
}  // namespace hs
#define HS_LAUNCH(ctx, kid, ...) do { hs::ProfScope ps__(ctx, kid); __VA_ARGS__; } while (0)

struct hslam_frame {
  hslam_ctx* ctx = nullptr;
  int w = 0, h = 0, levels = 0;
  float4* dI[HSLAM_MAX_PYR_LEVELS] = {nullptr};  // {I,dx,dy,absSquaredGrad} per pixel
  float4* base = nullptr;                        // one allocation for all levels
  float4* pair0 = nullptr;                       // level-0 PAIR image (2 float4 per pixel), built on first use by a BA window
  float* staging = nullptr;                      // device copy of the input image (H2D target)
  float* gamma = nullptr;                        // 256-entry BGrad table or null
  float* pinned = nullptr;                       // pinned host staging for async H2D
  cudaEvent_t h2d_done = nullptr;                // guards re-use of `pinned`
};
