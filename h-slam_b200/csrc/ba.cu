// ba.cu — host side of the windowed-BA hot path: device-resident mirror of EnergyFunctional's
// points/residuals, re-packing on topology change, launch orchestration and the C ABI
// (include/hslam_b200.h).  Device code: ba_kernels.cuh.
//
// Layout in HBM.  Residuals are re-ordered once per topology change by (host,target) pair, so a
// block of the linearise kernel shares one target image and one precalc record, and a block of the
// top-Hessian kernel reduces into ONE relative 13x13 accumulator.  Everything per residual is SoA
// with a padded stride (component-major): every load/store of a warp is one coalesced 128-byte line.
// Points are ordered by host so a Schur block owns the n^2 8x8 blocks of one host.
//
// Host<->device traffic goes through one pinned arena per handle; copies are asynchronous and a call
// synchronises at most once (not at all between hslam_ba_begin_async and hslam_ba_wait).
#include "ba_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <dlfcn.h>
#include <cstdlib>

using namespace hs;

int hs_frame_ensure_pair(hslam_frame* f);   // pyramid.cu

template <class T> struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = n + n / 8 + 64;
    cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
    if (e != cudaSuccess) { hs::set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); cap = 0; return HSLAM_ERR_NOMEM; }
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct hslam_ba {
  hslam_ctx* ctx = nullptr;
  int n = 0, P = 0, R = 0, rs = 0, LD = 0, N = 0;
  bool have_window = false, have_topology = false;
  int add_priors = 1;
  bool tmap_ok = true;                   // the Jacobian tensor maps of the top kernel could be encoded
  // window tables
  DevBuf<float> d_precalc, d_adHTd;
  DevBuf<double> d_adH, d_adT, d_fprior, d_fdprior;
  std::vector<double> h_fprior, h_fdprior;   // host copies for calcLEnergy's prior terms
  // points
  DevBuf<float4> d_uvid, d_color, d_weights, d_Hcd[2];
  DevBuf<float> d_priorF, d_deltaF, d_HdiF, d_bdSumF, d_idh, d_Hdd[2], d_bd[2], d_step;
  DevBuf<int> d_phost, d_res_begin, d_res_index, d_slotpos;
  int pstr = 12;                         // floats per point in the record arrays (sc_pstr(maxres))
  std::vector<int> p_sorted2orig, p_orig2sorted;
  std::vector<float4> h_uvid;
  // residuals
  DevBuf<int> d_rpoint, d_rpair;
  DevBuf<int8_t> d_state, d_newState;
  DevBuf<uint8_t> d_isLin, d_isActive;
  DevBuf<float> d_senergy, d_newEnergy, d_newEnergyWO, d_J, d_JpJdF, d_rtz, d_cpt, d_pto, d_prec;
  std::vector<int> r_sorted2orig, r_orig2sorted;
  // host mirrors of the sorted residual order (for re-grouping when residuals become linearised)
  std::vector<int> h_rpair, h_rpoint, h_rb, h_ri;   // per sorted residual: pair, sorted point; CSR of sorted points over sorted residual positions
  std::vector<uint8_t> h_lin;                        // per sorted residual: isLinearized
  std::vector<int> h_segBegin;                       // [2*n*n+1]: segment (pair, isLinearized) -> first sorted position
  DevBuf<int> d_perm;
  DevBuf<float> d_permF;                             // scratch of the re-grouping gather
  DevBuf<int> d_r_s2o, d_p_s2o;          // device copies of the sorted->caller permutations
  DevBuf<unsigned char> d_marg;          // marginalisation flags per sorted point
  DevBuf<char> d_scratch;                // staging for layout conversion kernels
  size_t scratch_used = 0;
  int n_linearized = 0;
  // chunks / partials
  DevBuf<TopChunk> d_topchunks;
  DevBuf<int> d_pair_chunk_begin;
  DevBuf<float> d_toppartial;
  DevBuf<unsigned int> d_pair_ticket;
  int n_topblocks[3] = {0, 0, 0};   // top-kernel BLOCKS per tile schedule (A, L, marginalisation)
  DevBuf<TopTile> d_toptiles;
  DevBuf<int> d_top_block_begin;
  DevBuf<LinTile> d_lintiles;
  int sched_tile0[3] = {0, 0, 0}, sched_bb0[3] = {0, 0, 0};   // offsets of every schedule in d_toptiles / d_top_block_begin
  int n_lintiles = 0;
  int jmode = 0;                 // 0 = full Jacobians stored, 1 = GN-loop mode (see hslam_ba_set_jacobian_mode)
  DevBuf<PtChunk> d_ptchunks;
  DevBuf<int> d_host_chunk_begin;
  DevBuf<float> d_scpartial;
  DevBuf<unsigned int> d_host_ticket;
  int n_ptchunks = 0;
  int maxres = 1;
  int target_blocks = 0;                 // blocks per stage this window should fill (0 = 2 x SM count)
  std::vector<int> h_hostBegin;   // kept for re-chunking (hslam_ba_set_parallelism)
  DevBuf<double> d_accTop, d_G, d_Hb, d_epart, d_energy;
  DevBuf<int> d_nres;          // [2]
  DevBuf<unsigned int> d_ticket;
  DevBuf<float> d_xc, d_xAd;
  DevBuf<double> d_x;
  DevBuf<int> d_ids;
  DevBuf<float> d_th;            // result of hslam_ba_new_frame_energy_th
  // descriptor
  BADesc hdesc;
  DevBuf<BADesc> d_desc;
  uint64_t desc_version = 1, desc_uploaded = 0;
  // pinned arena + deferred host-side unpacking
  char* arena = nullptr;
  size_t arena_cap = 0, arena_used = 0;
  std::vector<std::function<void()>> pending;
  bool async_mode = false;
  // CUDA graph of the single-window GN iteration (hslam_ba_iteration_device): re-captured when a launch parameter changes
  cudaGraphExec_t g_iter = nullptr;
  long long g_key[12] = {0};
  int g_kernels = 0;
};

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

// ---------------------------------------------------------------- pinned arena
static int ba_flush(hslam_ba* ba) {
  HS_CUDA(cudaStreamSynchronize(ba->ctx->stream));
  for (auto& f : ba->pending) f();
  ba->pending.clear();
  ba->arena_used = 0;
  ba->scratch_used = 0;
  return 0;
}
static int arena_alloc(hslam_ba* ba, size_t bytes, void** out) {
  bytes = (bytes + 63) & ~(size_t)63;
  if (ba->arena_used + bytes > ba->arena_cap) {
    TRY(ba_flush(ba));   // nothing in flight any more
    if (bytes > ba->arena_cap) {
      if (ba->arena) cudaFreeHost(ba->arena);
      ba->arena = nullptr;
      ba->arena_cap = std::max(bytes * 2, (size_t)8 << 20);
      HS_CUDA(cudaMallocHost((void**)&ba->arena, ba->arena_cap));
    }
  }
  *out = ba->arena + ba->arena_used;
  ba->arena_used += bytes;
  return 0;
}
static int end_call(hslam_ba* ba) { return (ba->async_mode || ba->pending.empty()) ? 0 : ba_flush(ba); }

template <class T> static int upload(hslam_ba* ba, T* dst, const T* src, size_t n) {
  if (!n) return 0;
  void* s;
  TRY(arena_alloc(ba, n * sizeof(T), &s));
  memcpy(s, src, n * sizeof(T));
  HS_CUDA(cudaMemcpyAsync(dst, s, n * sizeof(T), cudaMemcpyHostToDevice, ba->ctx->stream));
  return 0;
}
// enqueue a D2H copy; `unpack(staging)` runs after the next synchronisation
template <class T> static int download_then(hslam_ba* ba, const T* src_dev, size_t n, std::function<void(const T*)> unpack) {
  if (!n) return 0;
  void* s;
  TRY(arena_alloc(ba, n * sizeof(T), &s));
  HS_CUDA(cudaMemcpyAsync(s, src_dev, n * sizeof(T), cudaMemcpyDeviceToHost, ba->ctx->stream));
  const T* sp = (const T*)s;
  ba->pending.push_back([unpack, sp]() { unpack(sp); });
  return 0;
}
template <class T> static int download(hslam_ba* ba, T* dst_host, const T* src_dev, size_t n) {
  if (!dst_host) return 0;
  return download_then<T>(ba, src_dev, n, [dst_host, n](const T* s) { memcpy(dst_host, s, n * sizeof(T)); });
}
// device scratch for layout conversions; lives until the next flush (bump allocation)
static int scratch_alloc(hslam_ba* ba, size_t bytes, void** out) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (ba->scratch_used + bytes > ba->d_scratch.cap) {
    TRY(ba_flush(ba));
    if (bytes > ba->d_scratch.cap) TRY(ba->d_scratch.ensure(std::max(bytes * 2, (size_t)16 << 20)));
  }
  *out = ba->d_scratch.p + ba->scratch_used;
  ba->scratch_used += bytes;
  return 0;
}
// per-residual SoA block [comps][rs] -> caller order [R][comps]: permuted on the device, then ONE
// contiguous D2H and a plain memcpy on the host (the host-side scatter cost ~0.3 ms per 84k residuals)
template <class T> static int download_res(hslam_ba* ba, const T* dev, int comps, T* out) {
  if (!out || ba->R == 0) return 0;
  void* tmp;
  TRY(scratch_alloc(ba, (size_t)ba->R * comps * sizeof(T), &tmp));
  HS_LAUNCH(ba->ctx, HSLAM_K_BA_EXPORT, (ba_export_kernel<T><<<(ba->R + 255) / 256, 256, 0, ba->ctx->stream>>>(
                                            dev, comps, ba->rs, ba->R, ba->d_r_s2o.p, (T*)tmp)));
  return download(ba, out, (const T*)tmp, (size_t)ba->R * comps);
}
template <class T> static int download_pts(hslam_ba* ba, const T* dev, int comps, T* out) {
  if (!out || ba->P == 0) return 0;
  void* tmp;
  TRY(scratch_alloc(ba, (size_t)ba->P * comps * sizeof(T), &tmp));
  // point arrays are [P] or [P][4] point-major: treat as comps interleaved -> stride 1 per component is
  // not SoA, so export component by component with stride = comps
  if (comps == 1) {
    HS_LAUNCH(ba->ctx, HSLAM_K_BA_EXPORT, (ba_export_kernel<T><<<(ba->P + 255) / 256, 256, 0, ba->ctx->stream>>>(
                                              dev, 1, ba->P, ba->P, ba->d_p_s2o.p, (T*)tmp)));
  } else {
    // [P][comps] AoS: export as `comps`-wide records
    HS_LAUNCH(ba->ctx, HSLAM_K_BA_EXPORT, (ba_export_aos_kernel<T><<<(ba->P + 255) / 256, 256, 0, ba->ctx->stream>>>(
                                              dev, comps, ba->P, ba->d_p_s2o.p, (T*)tmp)));
  }
  return download(ba, out, (const T*)tmp, (size_t)ba->P * comps);
}

// ---------------------------------------------------------------- descriptor
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*hs_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static hs_encode_tiled_fn hs_encode_tiled() {
  static hs_encode_tiled_fn fn = []() -> hs_encode_tiled_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (hs_encode_tiled_fn)p;
  }();
  return fn;
}
// A component-major per-residual array [nrows][rs] as a 2-D tensor (inner = residual); box = TOP_ROW residuals x `rows` rows
static bool encode_rows_map(CUtensorMap* out, float* base, int rs, int nrows, int rows) {
  hs_encode_tiled_fn fn = hs_encode_tiled();
  if (!fn || !base) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)rs, (cuuint64_t)nrows};
  const cuuint64_t gstride[1] = {(cuuint64_t)rs * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TOP_ROW, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void fill_desc(hslam_ba* ba) {
  BADesc& D = ba->hdesc;
  ba->tmap_ok = ba->R == 0 || (encode_rows_map(&D.tmapJ[0], ba->d_J.p, ba->rs, JS, 22) && encode_rows_map(&D.tmapJ[1], ba->d_J.p, ba->rs, JS, 17) &&
                               encode_rows_map(&D.tmapJ[2], ba->d_J.p, ba->rs, JS, TOP_NJ_L) && encode_rows_map(&D.tmapJ[3], ba->d_rtz.p, ba->rs, 8, 8));
  BARes& r = D.r;
  r.R = ba->R; r.rs = ba->rs;
  r.point = ba->d_rpoint.p; r.pair = ba->d_rpair.p;
  r.state_state = ba->d_state.p; r.newState = ba->d_newState.p; r.isLin = ba->d_isLin.p; r.isActive = ba->d_isActive.p;
  r.state_energy = ba->d_senergy.p; r.newEnergy = ba->d_newEnergy.p; r.newEnergyWO = ba->d_newEnergyWO.p;
  r.J = ba->d_J.p; r.JpJdF = ba->d_JpJdF.p; r.rtz = ba->d_rtz.p; r.cpt = ba->d_cpt.p; r.pto = ba->d_pto.p;
  r.prec = ba->d_prec.p; r.slotpos = ba->d_slotpos.p; r.pstr = ba->pstr;
  BAPts& p = D.p;
  p.P = ba->P; p.uvid = ba->d_uvid.p; p.color = ba->d_color.p; p.weights = ba->d_weights.p;
  p.priorF = ba->d_priorF.p; p.deltaF = ba->d_deltaF.p; p.host = ba->d_phost.p;
  p.res_begin = ba->d_res_begin.p; p.res_index = ba->d_res_index.p;
  p.HdiF = ba->d_HdiF.p; p.bdSumF = ba->d_bdSumF.p; p.idepth_hessian = ba->d_idh.p;
  for (int m = 0; m < 2; ++m) { p.Hdd[m] = ba->d_Hdd[m].p; p.bd[m] = ba->d_bd[m].p; p.Hcd[m] = ba->d_Hcd[m].p; }
  p.step = ba->d_step.p;
  D.W.precalc = ba->d_precalc.p; D.W.adHTdeltaF = ba->d_adHTd.p;
  D.epart = ba->d_epart.p; D.ticket = ba->d_ticket.p; D.energy = ba->d_energy.p;
  D.lin_warps = ba->n_lintiles * (LIN_THREADS / 32);
  D.lintiles = ba->d_lintiles.p; D.jmode = ba->jmode;
  D.topchunks = ba->d_topchunks.p; D.toppartial = ba->d_toppartial.p; D.pair_ticket = ba->d_pair_ticket.p;
  D.pair_chunk_begin = ba->d_pair_chunk_begin.p; D.accTop = ba->d_accTop.p;
  for (int m = 0; m < 3; ++m) {
    D.sched[m].tiles = ba->d_toptiles.p + ba->sched_tile0[m];
    D.sched[m].block_begin = ba->d_top_block_begin.p + ba->sched_bb0[m];
    D.sched[m].blocks = ba->n_topblocks[m]; D.sched[m].pad = 0;
  }
  D.ptchunks = ba->d_ptchunks.p; D.scpartial = ba->d_scpartial.p; D.host_ticket = ba->d_host_ticket.p;
  D.host_chunk_begin = ba->d_host_chunk_begin.p; D.G = ba->d_G.p; D.LD = ba->LD; D.maxres = ba->maxres;
  StitchArgs& S = D.S;
  S.n = ba->n; S.LD = ba->LD;
  S.adHost = ba->d_adH.p; S.adTarget = ba->d_adT.p; S.accTop = ba->d_accTop.p; S.G = ba->d_G.p; S.Hb = ba->d_Hb.p;
  S.nres = ba->d_nres.p; S.framePrior = ba->d_fprior.p; S.frameDeltaPrior = ba->d_fdprior.p;
  S.add_priors = ba->add_priors; S.has_L = ba->n_linearized != 0; S.L_set = 1;
  for (int i = 0; i < 4; ++i) S.cDeltaF[i] = D.W.cDeltaF[i];
  D.marg = nullptr; D.shift_prior = 1;
  D.r_s2o = ba->d_r_s2o.p; D.p_s2o = ba->d_p_s2o.p;
  D.x = ba->d_x.p; D.xc = ba->d_xc.p; D.xAd = ba->d_xAd.p;
  ba->desc_version++;
}
static int sync_desc(hslam_ba* ba) {
  if (ba->desc_uploaded == ba->desc_version) return 0;
  TRY(ba->d_desc.ensure(1));
  TRY(upload(ba, ba->d_desc.p, &ba->hdesc, 1));
  ba->desc_uploaded = ba->desc_version;
  return 0;
}

// Chunk tables of the top-Hessian and Schur kernels.  A chunk is the unit one block reduces into ONE fp32
// partial; the partials of a pair / host are folded in chunk order into fp64.  The target is ~2 waves of
// blocks over the SMs for a window that runs alone; a batch of windows shares the SMs, so
// hslam_ba_batch_create asks every member for proportionally fewer, larger chunks (fewer partials to
// write and fold, fewer block-level reductions).
static int build_chunks(hslam_ba* ba) {
  hslam_ctx* ctx = ba->ctx;
  const int n = ba->n, nn = n * n;
  const std::vector<int>& hostBegin = ba->h_hostBegin;
  const int target_blocks = ba->target_blocks > 0 ? ba->target_blocks : 2 * ctx->sm_count;
  {
    // Three tile schedules over the (pair, isLinearized) segments of the sorted residual array: A = the non-linearised
    // half of every pair, L = the linearised half, marginalisation = the whole pair.  Tiles of <= TOP_THREADS residuals
    // that never cross a pair; equal runs of tiles per block; chunks = maximal same-pair runs inside a block (see
    // ba_top_kernel).  Chunk indices are global: schedule m's chunks follow schedule m-1's.
    const std::vector<int>& seg = ba->h_segBegin;
    struct HT { int start, count, pair; };
    std::vector<TopTile> tiles;
    std::vector<TopChunk> ch;
    std::vector<int> bb;
    std::vector<int> pcb(3 * nn + 1, 0);
    for (int m = 0; m < 3; ++m) {
      std::vector<HT> ht;
      for (int k = 0; k < nn; ++k) {
        const int lo = m == 1 ? seg[2 * k + 1] : seg[2 * k], hi = m == 0 ? seg[2 * k + 1] : seg[2 * k + 2];
        for (int s0 = lo; s0 < hi; s0 += TOP_THREADS) ht.push_back({s0, std::min(TOP_THREADS, hi - s0), k});
      }
      const int NT = (int)ht.size();
      const int per = std::max(1, (NT + target_blocks - 1) / target_blocks);
      const int T = (NT + per - 1) / per;
      ba->sched_tile0[m] = (int)tiles.size();
      ba->sched_bb0[m] = (int)bb.size();
      ba->n_topblocks[m] = T;
      const int c0 = (int)ch.size();
#ifndef TOP_ITEM_TILES
#define TOP_ITEM_TILES 12
#endif
      // ITEMS: a pair's tiles are cut into pieces of <= TOP_ITEM_TILES tiles (= one chunk each: one fp32 partial, folded
      // in chunk order).  Chunk ids follow the pair order (the fold walks pcb[pair] .. pcb[pair + 1]).  The items are then
      // dealt to the blocks in the order (host, piece, target), each to the block with the least tiles so far: the
      // blocks of a window — resident together, progressing in step — work on the SAME stretch of a host's points
      // towards its different targets at the same time, so the 32-byte records that they store into the point-major
      // record array (stride pstr per point, one record per target) meet in L2 and leave as whole lines.  With equal
      // runs of tiles per block the six records of a point were written a whole kernel apart: 45 us of 390 went into
      // partial-line write-backs (measured: records stored in residual order instead).
      struct Item { int t0, nt, chunk, host, piece, target; };
      std::vector<Item> items;
      {
        const int lim = std::max(1, std::min(TOP_ITEM_TILES, per));
        int t = 0;
        while (t < NT) {
          int e = t;
          while (e < NT && ht[e].pair == ht[t].pair) ++e;
          const int nt = e - t, pieces = (nt + lim - 1) / lim, sz = (nt + pieces - 1) / pieces;
          for (int q = 0; q < pieces; ++q) {
            const int a = t + q * sz, b = std::min(e, a + sz);
            if (a >= b) break;
            int cnt = 0;
            for (int x = a; x < b; ++x) cnt += ht[x].count;
            ch.push_back(TopChunk{ht[t].pair, ht[a].start, cnt, m});
            items.push_back(Item{a, b - a, (int)ch.size() - 1, ht[t].pair % n, q, ht[t].pair / n});
          }
          t = e;
        }
      }
      std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) {
        if (x.host != y.host) return x.host < y.host;
        if (x.piece != y.piece) return x.piece < y.piece;
        return x.target < y.target;
      });
      std::vector<std::vector<int>> mine(T);
      std::vector<int> load(T, 0);
      for (size_t it = 0; it < items.size(); ++it) {
        int best = 0;
        for (int b = 1; b < T; ++b) if (load[b] < load[best]) best = b;
        mine[best].push_back((int)it); load[best] += items[it].nt;
      }
      int off = 0;
      for (int b = 0; b < T; ++b) {
        bb.push_back(off);
        for (int it : mine[b])
          for (int x = items[it].t0; x < items[it].t0 + items[it].nt; ++x, ++off)
            tiles.push_back(TopTile{ht[x].start, ht[x].count, items[it].chunk, ht[x].pair});
      }
      bb.push_back(NT);
      int c = c0;
      for (int k = 0; k < nn; ++k) {
        pcb[m * nn + k] = c;
        while (c < (int)ch.size() && ch[c].pair == k) ++c;
      }
    }
    pcb[3 * nn] = (int)ch.size();
    TRY(ba->d_topchunks.ensure(ch.size() + 1)); TRY(ba->d_pair_chunk_begin.ensure(pcb.size()));
    TRY(ba->d_toptiles.ensure(tiles.size() + 1)); TRY(ba->d_top_block_begin.ensure(bb.size() + 1));
    TRY(ba->d_toppartial.ensure((ch.size() + 1) * TOP_PART));
    TRY(ba->d_pair_ticket.ensure(3 * nn + 1));
    HS_CUDA(cudaMemsetAsync(ba->d_pair_ticket.p, 0, (3 * nn + 1) * sizeof(unsigned int), ctx->stream));
    TRY(upload(ba, ba->d_topchunks.p, ch.data(), ch.size())); TRY(upload(ba, ba->d_pair_chunk_begin.p, pcb.data(), pcb.size()));
    TRY(upload(ba, ba->d_toptiles.p, tiles.data(), tiles.size())); TRY(upload(ba, ba->d_top_block_begin.p, bb.data(), bb.size()));
    // pairs without any chunk in a schedule are never folded: their accumulators must read as zero
    HS_CUDA(cudaMemsetAsync(ba->d_accTop.p, 0, 3 * (size_t)nn * TOP_DENSE * sizeof(double), ctx->stream));
    // lineariser: tiles of <= LIN_THREADS non-linearised residuals of one pair
    // (Order: pair-major, i.e. the sorted order.  A host-major order — the j-th tiles of all targets of a host next to each
    // other, so that the host's point records are re-read from L2 — was measured: 636 vs 635 us, all-active 1.446 vs 1.408 ms.)
    std::vector<LinTile> lt;
    for (int k = 0; k < nn; ++k)
      for (int s0 = seg[2 * k]; s0 < seg[2 * k + 1]; s0 += LIN_THREADS)
        lt.push_back(LinTile{s0, std::min(LIN_THREADS, seg[2 * k + 1] - s0), k, 0});
    ba->n_lintiles = (int)lt.size();
    TRY(ba->d_lintiles.ensure(lt.size() + 1));
    TRY(upload(ba, ba->d_lintiles.p, lt.data(), lt.size()));
    TRY(ba->d_epart.ensure((size_t)(lt.size() + 2) * (LIN_THREADS / 32)));   // one energy partial per warp
  }
  {
    // Chunks are cut in whole
    // sub-tiles of SC_PTS points, as evenly as possible within a host, and a little below the target so that
    // every block of a launch is resident at once (no second, nearly empty wave).
    int total_sub = 0;
    for (int h = 0; h < n; ++h) total_sub += (hostBegin[h + 1] - hostBegin[h] + SC_PTS - 1) / SC_PTS;
#ifndef SC_BLOCKS_NUM
#define SC_BLOCKS_NUM 3
#define SC_BLOCKS_DEN 2
#endif
    int sc_blocks = SC_BLOCKS_NUM * target_blocks / SC_BLOCKS_DEN;
    if (sc_blocks > 4 * n) sc_blocks -= n;
    // batched windows: measured (64 x config-2 windows) the Schur launch keeps gaining down to ~6 chunk waves for
    // the whole batch (finer chunks balance the 3-blocks-per-SM kernel; below that the fold cost grows)
    // An explicit target (set_parallelism, i.e. a member of a batch) asks for 3 x as many Schur chunks as top blocks:
    // ~6 waves of the 2-blocks-per-SM 8x8-tile kernel for the whole batch (measured at 64 config-2 windows: 14 chunks
    // per window 218 us, 28: 211, 35: 212, 42: 215, 56: 223 — and the fold grows with the chunk count: 30 .. 45 us).
    if (ba->target_blocks > 0) sc_blocks = std::min(3 * ba->target_blocks, std::max(n, total_sub / 2));
    if (const char* e = getenv("HSLAM_SC_BLOCKS")) sc_blocks = atoi(e);   // tuning override (per window)
    if (sc_blocks < 1) sc_blocks = 1;
    const int q = std::max(1, (total_sub + sc_blocks - 1) / sc_blocks);   // sub-tiles per chunk
    std::vector<PtChunk> ch;
    std::vector<int> hcb(n + 1, 0);
    for (int h = 0; h < n; ++h) {
      hcb[h] = (int)ch.size();
      const int len = hostBegin[h + 1] - hostBegin[h];
      if (len == 0) continue;
      const int nsub = (len + SC_PTS - 1) / SC_PTS, parts = (nsub + q - 1) / q;
      int s0 = hostBegin[h];
      for (int c = 0; c < parts; ++c) {
        const int subs = nsub / parts + (c < nsub % parts ? 1 : 0);
        const int cnt = std::min(subs * SC_PTS, hostBegin[h + 1] - s0);
        ch.push_back(PtChunk{h, s0, cnt, 0});
        s0 += cnt;
      }
    }
    hcb[n] = (int)ch.size();
    ba->n_ptchunks = (int)ch.size();
    TRY(ba->d_ptchunks.ensure(ch.size() + 1)); TRY(ba->d_host_chunk_begin.ensure(hcb.size()));
    TRY(ba->d_scpartial.ensure((ch.size() + 1) * (size_t)ba->LD * ba->LD));
    // tiles of the host's own 8-block are never written by the Schur kernel: they must read as zero
    HS_CUDA(cudaMemsetAsync(ba->d_scpartial.p, 0, (ch.size() + 1) * (size_t)ba->LD * ba->LD * sizeof(float), ctx->stream));
    TRY(ba->d_host_ticket.ensure(n + 1));
    HS_CUDA(cudaMemsetAsync(ba->d_host_ticket.p, 0, (n + 1) * sizeof(unsigned int), ctx->stream));
    TRY(upload(ba, ba->d_ptchunks.p, ch.data(), ch.size())); TRY(upload(ba, ba->d_host_chunk_begin.p, hcb.data(), hcb.size()));
    HS_CUDA(cudaMemsetAsync(ba->d_G.p, 0, (size_t)n * ba->LD * ba->LD * sizeof(double), ctx->stream));
  }
  return 0;
}

// slotpos of every sorted residual: the float offset of its 32-byte records in the point-major record arrays
// (BARes::JpJdF / prec), from the host mirrors of the points' residual lists.
static int upload_slotpos(hslam_ba* ba) {
  const int P = ba->P, R = ba->R;
  std::vector<int> sp((size_t)R + 1, 0);
  for (int p = 0; p < P; ++p)
    for (int i = ba->h_rb[p], j = 0; i < ba->h_rb[p + 1]; ++i, ++j) sp[ba->h_ri[i]] = p * ba->pstr + 8 * j;
  return upload(ba, ba->d_slotpos.p, sp.data(), (size_t)R);
}

// Residuals have changed their isLinearized flag (hslam_ba_fix_linearization): restore the sorted order's invariant —
// inside every (host,target) pair the non-linearised residuals first, then the linearised ones, each group in its old
// (stable) order — by permuting every per-residual device array once, and rebuild the tables that hold sorted positions.
static int regroup(hslam_ba* ba) {
  hslam_ctx* ctx = ba->ctx;
  const int R = ba->R, n = ba->n, nn = n * n;
  TRY(ba_flush(ba));
  std::vector<int> seg(2 * nn + 1, 0);
  for (int k = 0; k < R; ++k) seg[2 * ba->h_rpair[k] + (ba->h_lin[k] ? 1 : 0) + 1]++;
  for (int k = 0; k < 2 * nn; ++k) seg[k + 1] += seg[k];
  std::vector<int> newpos(R), perm(R);
  bool identity = true;
  {
    std::vector<int> cur(seg.begin(), seg.end() - 1);
    for (int k = 0; k < R; ++k) {
      const int q = cur[2 * ba->h_rpair[k] + (ba->h_lin[k] ? 1 : 0)]++;
      newpos[k] = q; perm[q] = k;
      identity = identity && q == k;
    }
  }
  ba->h_segBegin = seg;
  if (!identity) {
    std::vector<int> s2o(R), rpair(R), rpoint(R);
    std::vector<uint8_t> lin(R);
    for (int q = 0; q < R; ++q) {
      const int k = perm[q];
      s2o[q] = ba->r_sorted2orig[k]; rpair[q] = ba->h_rpair[k]; rpoint[q] = ba->h_rpoint[k]; lin[q] = ba->h_lin[k];
    }
    ba->r_sorted2orig.swap(s2o); ba->h_rpair.swap(rpair); ba->h_rpoint.swap(rpoint); ba->h_lin.swap(lin);
    for (int q = 0; q < R; ++q) ba->r_orig2sorted[ba->r_sorted2orig[q]] = q;
    for (int& v : ba->h_ri) v = newpos[v];
    // device: every per-residual array through one scratch buffer
    const size_t rsz = ba->rs;
    TRY(ba->d_perm.ensure(R));
    TRY(ba->d_permF.ensure((size_t)JS * rsz));
    TRY(upload(ba, ba->d_perm.p, perm.data(), R));
    const unsigned gx = (unsigned)((R + 255) / 256);
    auto permF = [&](float* arr, int comps) -> int {
      HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_permute_kernel<float><<<dim3(gx, comps), 256, 0, ctx->stream>>>(arr, ba->d_permF.p, rsz, R, ba->d_perm.p)));
      HS_CUDA(cudaMemcpyAsync(arr, ba->d_permF.p, (size_t)comps * rsz * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
      return 0;
    };
    auto permB = [&](void* arr) -> int {
      HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_permute_kernel<unsigned char><<<dim3(gx, 1), 256, 0, ctx->stream>>>(
                                            (const unsigned char*)arr, (unsigned char*)ba->d_permF.p, rsz, R, ba->d_perm.p)));
      HS_CUDA(cudaMemcpyAsync(arr, ba->d_permF.p, (size_t)R, cudaMemcpyDeviceToDevice, ctx->stream));
      return 0;
    };
    TRY(permF(ba->d_J.p, JS)); TRY(permF(ba->d_rtz.p, 8)); TRY(permF(ba->d_cpt.p, 3)); TRY(permF(ba->d_pto.p, 16));
    TRY(permF(ba->d_senergy.p, 1)); TRY(permF(ba->d_newEnergy.p, 1)); TRY(permF(ba->d_newEnergyWO.p, 1));
    TRY(permB(ba->d_state.p)); TRY(permB(ba->d_newState.p)); TRY(permB(ba->d_isLin.p)); TRY(permB(ba->d_isActive.p));
    // (the JpJdF / point-term records are indexed by (point, slot): they do not move)
    HS_CUDA(cudaGetLastError());
    TRY(upload(ba, ba->d_rpoint.p, ba->h_rpoint.data(), R)); TRY(upload(ba, ba->d_rpair.p, ba->h_rpair.data(), R));
    TRY(upload(ba, ba->d_r_s2o.p, ba->r_sorted2orig.data(), R));
    TRY(upload(ba, ba->d_res_index.p, ba->h_ri.data(), R));
    TRY(upload_slotpos(ba));
  }
  TRY(build_chunks(ba));
  fill_desc(ba);
  return ba_flush(ba);
}

extern "C" {

int hslam_ba_create(hslam_ctx* ctx, hslam_ba** out) {
  HS_ARG(ctx && out, "hslam_ba_create: null argument");
  HS_CUDA(cudaSetDevice(ctx->device));
  hslam_ba* ba = new hslam_ba();
  ba->ctx = ctx;
  memset(&ba->hdesc, 0, sizeof(ba->hdesc));
  TRY(ba->d_ticket.ensure(4));
  TRY(ba->d_nres.ensure(4));
  TRY(ba->d_energy.ensure(2));
  HS_CUDA(cudaMemsetAsync(ba->d_ticket.p, 0, 4 * sizeof(unsigned int), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_nres.p, 0, 4 * sizeof(int), ctx->stream));
  *out = ba;
  return 0;
}

int hslam_ba_destroy(hslam_ba* ba) {
  if (!ba) return 0;
  cudaSetDevice(ba->ctx->device);
  cudaStreamSynchronize(ba->ctx->stream);
  ba->pending.clear();
  ba->d_precalc.release(); ba->d_adHTd.release(); ba->d_adH.release(); ba->d_adT.release();
  ba->d_fprior.release(); ba->d_fdprior.release();
  ba->d_uvid.release(); ba->d_color.release(); ba->d_weights.release();
  ba->d_priorF.release(); ba->d_deltaF.release(); ba->d_HdiF.release(); ba->d_bdSumF.release(); ba->d_idh.release();
  for (int m = 0; m < 2; ++m) { ba->d_Hdd[m].release(); ba->d_bd[m].release(); ba->d_Hcd[m].release(); }
  ba->d_prec.release(); ba->d_slotpos.release();
  ba->d_accTop.release();
  ba->d_step.release(); ba->d_phost.release(); ba->d_res_begin.release(); ba->d_res_index.release();
  ba->d_rpoint.release(); ba->d_rpair.release(); ba->d_state.release(); ba->d_newState.release();
  ba->d_isLin.release(); ba->d_isActive.release(); ba->d_senergy.release(); ba->d_newEnergy.release();
  ba->d_newEnergyWO.release(); ba->d_J.release(); ba->d_JpJdF.release(); ba->d_rtz.release(); ba->d_cpt.release();
  ba->d_pto.release(); ba->d_topchunks.release(); ba->d_pair_chunk_begin.release(); ba->d_toppartial.release();
  ba->d_pair_ticket.release(); ba->d_host_ticket.release(); ba->d_toptiles.release(); ba->d_top_block_begin.release();
  ba->d_ptchunks.release(); ba->d_host_chunk_begin.release(); ba->d_lintiles.release(); ba->d_perm.release(); ba->d_permF.release(); ba->d_scpartial.release(); ba->d_G.release();
  ba->d_Hb.release(); ba->d_epart.release(); ba->d_energy.release(); ba->d_nres.release(); ba->d_ticket.release();
  ba->d_xc.release(); ba->d_xAd.release(); ba->d_x.release(); ba->d_ids.release(); ba->d_th.release(); ba->d_desc.release();
  ba->d_r_s2o.release(); ba->d_p_s2o.release(); ba->d_scratch.release(); ba->d_marg.release();
  if (ba->g_iter) cudaGraphExecDestroy(ba->g_iter);
  if (ba->arena) cudaFreeHost(ba->arena);
  delete ba;
  return 0;
}

int hslam_ba_begin_async(hslam_ba* ba) {
  HS_ARG(ba, "hslam_ba_begin_async: null");
  ba->async_mode = true;
  return 0;
}
int hslam_ba_wait(hslam_ba* ba) {
  HS_ARG(ba, "hslam_ba_wait: null");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  ba->async_mode = false;
  return ba_flush(ba);
}

int hslam_ba_set_parallelism(hslam_ba* ba, int target_blocks) {
  HS_ARG(ba && target_blocks >= 0, "hslam_ba_set_parallelism: bad argument");
  if (ba->target_blocks == target_blocks) return 0;
  ba->target_blocks = target_blocks;
  if (!ba->have_topology) return 0;
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(ba_flush(ba));
  TRY(build_chunks(ba));
  fill_desc(ba);
  return ba_flush(ba);
}

int hslam_ba_get_parallelism(hslam_ba* ba, int* target_blocks_out) {
  HS_ARG(ba && target_blocks_out, "hslam_ba_get_parallelism: null");
  *target_blocks_out = ba->target_blocks;
  return 0;
}

int hslam_ba_set_add_priors(hslam_ba* ba, int add) {
  HS_ARG(ba, "hslam_ba_set_add_priors: null");
  ba->add_priors = add ? 1 : 0;
  ba->hdesc.S.add_priors = ba->add_priors;
  ba->desc_version++;
  return 0;
}

int hslam_ba_set_window(hslam_ba* ba, const hslam_window* w) {
  HS_ARG(ba && w, "hslam_ba_set_window: null argument");
  HS_ARG(w->n_frames >= 1 && w->n_frames <= HSLAM_MAX_FRAMES, "hslam_ba_set_window: n_frames out of range");
  HS_ARG(w->frames && w->frameEnergyTH && w->precalc && w->adHost && w->adTarget && w->adHTdeltaF && w->framePrior &&
             w->frameDeltaPrior, "hslam_ba_set_window: null table");
  HS_ARG(!ba->have_topology || w->n_frames == ba->n, "hslam_ba_set_window: n_frames changed, call set_topology again");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const int n = w->n_frames;
  ba->n = n; ba->N = 8 * n + 4; ba->LD = ((8 * n + 5) + 3) & ~3;
  BAWin& W = ba->hdesc.W;
  W.n = n; W.w = w->w; W.h = w->h;
  W.n_rcp = (65536 + n - 1) / n;
  W.fx = w->fx; W.fy = w->fy; W.cx = w->cx; W.cy = w->cy; W.fxi = w->fxi; W.fyi = w->fyi;
  W.huber = ctx->prm.huberTH; W.outlierTH = ctx->prm.outlierTHSumComponent; W.scale_idepth = ctx->prm.scale_idepth;
  W.scale_f = ctx->prm.scale_f; W.scale_c = ctx->prm.scale_c; W.affA = ctx->prm.affineOptModeA; W.affB = ctx->prm.affineOptModeB;
  W.thN = ctx->prm.frameEnergyTHN; W.thFac = ctx->prm.frameEnergyTHFacMedian; W.thConst = ctx->prm.frameEnergyTHConstWeight;
  W.thOverall = ctx->prm.overallEnergyTHWeight;
  for (int i = 0; i < 4; ++i) { W.cDeltaF[i] = w->cDeltaF[i]; ba->hdesc.S.cPrior[i] = w->cPrior[i]; }
  for (int f = 0; f < n; ++f) {
    HS_ARG(w->frames[f] && w->frames[f]->ctx == ctx, "hslam_ba_set_window: frame missing or from another context");
    HS_ARG(w->frames[f]->w == w->w && w->frames[f]->h == w->h, "hslam_ba_set_window: frame size mismatch");
    TRY(hs_frame_ensure_pair(const_cast<hslam_frame*>(w->frames[f])));   // lazily built cache of the frame
    W.img[f] = w->frames[f]->dI[0];
    W.pimg[f] = w->frames[f]->pair0;
    W.frameEnergyTH[f] = w->frameEnergyTH[f];
  }
  size_t nn = (size_t)n * n;
  TRY(ba->d_precalc.ensure(nn * HSLAM_PRECALC_STRIDE)); TRY(ba->d_adHTd.ensure(nn * 8));
  TRY(ba->d_adH.ensure(nn * 64)); TRY(ba->d_adT.ensure(nn * 64));
  TRY(ba->d_fprior.ensure(n * 8)); TRY(ba->d_fdprior.ensure(n * 8));
  TRY(upload(ba, ba->d_adH.p, w->adHost, nn * 64));
  TRY(upload(ba, ba->d_adT.p, w->adTarget, nn * 64));
  TRY(upload(ba, ba->d_fprior.p, w->framePrior, (size_t)n * 8));
  TRY(upload(ba, ba->d_fdprior.p, w->frameDeltaPrior, (size_t)n * 8));
  ba->h_fprior.assign(w->framePrior, w->framePrior + (size_t)n * 8);
  ba->h_fdprior.assign(w->frameDeltaPrior, w->frameDeltaPrior + (size_t)n * 8);
  TRY(upload(ba, ba->d_precalc.p, w->precalc, nn * HSLAM_PRECALC_STRIDE));
  TRY(upload(ba, ba->d_adHTd.p, w->adHTdeltaF, nn * 8));
  TRY(ba->d_x.ensure(ba->N)); TRY(ba->d_xc.ensure(4)); TRY(ba->d_xAd.ensure(nn * 8));
  TRY(ba->d_Hb.ensure(3 * ((size_t)ba->N * ba->N + ba->N) + 4));   // + the all-reduce tail (energy, resInA, resInL)
  if (ba->d_accTop.cap < 3 * nn * TOP_DENSE) {
    TRY(ba->d_accTop.ensure(3 * nn * TOP_DENSE));
    HS_CUDA(cudaMemsetAsync(ba->d_accTop.p, 0, 3 * nn * TOP_DENSE * sizeof(double), ctx->stream));
  }
  if (ba->d_G.cap < (size_t)n * ba->LD * ba->LD) {
    TRY(ba->d_G.ensure((size_t)n * ba->LD * ba->LD));
    HS_CUDA(cudaMemsetAsync(ba->d_G.p, 0, (size_t)n * ba->LD * ba->LD * sizeof(double), ctx->stream));
  }
  ba->have_window = true;
  fill_desc(ba);
  return 0;
}

int hslam_ba_set_topology(hslam_ba* ba, const hslam_point_soa* pt, const hslam_residual_soa* rs) {
  HS_ARG(ba && pt && rs, "hslam_ba_set_topology: null argument");
  HS_ARG(ba->have_window, "hslam_ba_set_topology: call hslam_ba_set_window first");
  const int P = pt->n_points, R = rs->n_residuals, n = ba->n;
  HS_ARG(P >= 0 && R >= 0, "hslam_ba_set_topology: negative count");
  HS_ARG(P == 0 || (pt->u && pt->v && pt->idepth && pt->idepth_zero && pt->color && pt->weights && pt->priorF &&
                    pt->deltaF && pt->host_idx && pt->res_begin), "hslam_ba_set_topology: null point array");
  HS_ARG(R == 0 || (pt->res_index && rs->point_idx && rs->target_idx && rs->state_state && rs->is_linearized &&
                    rs->is_active), "hslam_ba_set_topology: null residual array");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  HS_ARG(P == 0 || pt->res_begin[P] == R, "hslam_ba_set_topology: CSR does not cover every residual exactly once");
  ba->maxres = 1;
  for (int p = 0; p < P; ++p) ba->maxres = std::max(ba->maxres, pt->res_begin[p + 1] - pt->res_begin[p]);
  HS_ARG(ba->maxres <= HSLAM_MAX_FRAMES, "hslam_ba_set_topology: a point has more residuals than HSLAM_MAX_FRAMES");
  HS_ARG(R < (1 << 26), "hslam_ba_set_topology: more than 2^26 residuals in one window");
  TRY(ba_flush(ba));
  // ---- points sorted by host (stable counting sort)
  std::vector<int> hostCount(n + 1, 0);
  for (int p = 0; p < P; ++p) {
    HS_ARG(pt->host_idx[p] >= 0 && pt->host_idx[p] < n, "hslam_ba_set_topology: host_idx out of range");
    hostCount[pt->host_idx[p] + 1]++;
  }
  for (int h = 0; h < n; ++h) hostCount[h + 1] += hostCount[h];
  std::vector<int> hostBegin(hostCount);
  ba->p_sorted2orig.assign(P, 0); ba->p_orig2sorted.assign(P, 0);
  {
    std::vector<int> cur(hostCount.begin(), hostCount.end() - 1);
    for (int p = 0; p < P; ++p) { int sidx = cur[pt->host_idx[p]]++; ba->p_sorted2orig[sidx] = p; ba->p_orig2sorted[p] = sidx; }
  }
  // ---- residuals: walk points in sorted order, then stable counting sort by pair
  std::vector<int> walk; walk.reserve(R);
  std::vector<char> seen(R, 0);
  for (int sp = 0; sp < P; ++sp) {
    int p = ba->p_sorted2orig[sp];
    for (int i = pt->res_begin[p]; i < pt->res_begin[p + 1]; ++i) {
      int r = pt->res_index[i];
      HS_ARG(r >= 0 && r < R && !seen[r], "hslam_ba_set_topology: residual index invalid or listed twice");
      HS_ARG(rs->point_idx[r] == p, "hslam_ba_set_topology: residual.point_idx disagrees with the point's list");
      HS_ARG(rs->target_idx[r] >= 0 && rs->target_idx[r] < n, "hslam_ba_set_topology: target_idx out of range");
      seen[r] = 1;
      walk.push_back(r);
    }
  }
  HS_ARG((int)walk.size() == R, "hslam_ba_set_topology: some residuals belong to no point");
  const int nn = n * n;
  // segments: (pair, isLinearized) -> the non-linearised residuals of a pair first, then its linearised ones
  std::vector<int> segBegin(2 * nn + 1, 0);
  auto pair_of = [&](int r) { return pt->host_idx[rs->point_idx[r]] + rs->target_idx[r] * n; };
  auto seg_of = [&](int r) { return 2 * pair_of(r) + (rs->is_linearized[r] ? 1 : 0); };
  for (int r : walk) segBegin[seg_of(r) + 1]++;
  for (int k = 0; k < 2 * nn; ++k) segBegin[k + 1] += segBegin[k];
  ba->r_sorted2orig.assign(R, 0); ba->r_orig2sorted.assign(R, 0);
  {
    std::vector<int> cur(segBegin.begin(), segBegin.end() - 1);
    for (int r : walk) { int k = cur[seg_of(r)]++; ba->r_sorted2orig[k] = r; ba->r_orig2sorted[r] = k; }
  }
  ba->h_segBegin = segBegin;
  ba->P = P; ba->R = R; ba->rs = (R + 31) & ~31;
  if (ba->rs == 0) ba->rs = 32;
  const size_t rsz = ba->rs;
  TRY(ba->d_r_s2o.ensure(R + 1)); TRY(ba->d_p_s2o.ensure(P + 1));
  TRY(upload(ba, ba->d_r_s2o.p, ba->r_sorted2orig.data(), R)); TRY(upload(ba, ba->d_p_s2o.p, ba->p_sorted2orig.data(), P));
  // ---- allocate
  TRY(ba->d_uvid.ensure(P + 1)); TRY(ba->d_color.ensure(2 * (size_t)P + 2)); TRY(ba->d_weights.ensure(2 * (size_t)P + 2));
  TRY(ba->d_priorF.ensure(P + 1)); TRY(ba->d_deltaF.ensure(P + 1)); TRY(ba->d_phost.ensure(P + 1));
  TRY(ba->d_res_begin.ensure(P + 2)); TRY(ba->d_res_index.ensure(R + 1));
  TRY(ba->d_HdiF.ensure(P + 1)); TRY(ba->d_bdSumF.ensure(P + 1)); TRY(ba->d_idh.ensure(P + 1)); TRY(ba->d_step.ensure(P + 1));
  for (int m = 0; m < 2; ++m) { TRY(ba->d_Hdd[m].ensure(P + 1)); TRY(ba->d_bd[m].ensure(P + 1)); TRY(ba->d_Hcd[m].ensure(P + 1)); }
  ba->pstr = sc_pstr(ba->maxres);
  const size_t recsz = (size_t)(P + 1) * ba->pstr;
  TRY(ba->d_prec.ensure(recsz)); TRY(ba->d_slotpos.ensure(rsz + TOP_ROW)); /* + slack: the top kernel stages whole rows */
  TRY(ba->d_rpoint.ensure(rsz)); TRY(ba->d_rpair.ensure(rsz)); TRY(ba->d_state.ensure(rsz)); TRY(ba->d_newState.ensure(rsz));
  TRY(ba->d_isLin.ensure(rsz + 256)); TRY(ba->d_isActive.ensure(rsz + 256)); /* + slack: the top kernel copies whole 160-byte flag rows */ TRY(ba->d_senergy.ensure(rsz)); TRY(ba->d_newEnergy.ensure(rsz));
  TRY(ba->d_newEnergyWO.ensure(rsz)); TRY(ba->d_J.ensure((size_t)JS * rsz + 2 * TOP_ROW)); /* + slack: the top kernel stages whole rows */ TRY(ba->d_JpJdF.ensure(recsz));
  TRY(ba->d_rtz.ensure(8 * rsz)); TRY(ba->d_cpt.ensure(3 * rsz)); TRY(ba->d_pto.ensure(16 * rsz));
  // ---- pack + upload points
  {
    std::vector<float4> uvid(P), col(2 * (size_t)P), wgt(2 * (size_t)P);
    std::vector<float> prior(P), delta(P);
    std::vector<int> host(P), rb(P + 1), ri(R);
    int pos = 0;
    for (int sp = 0; sp < P; ++sp) {
      int p = ba->p_sorted2orig[sp];
      uvid[sp] = make_float4(pt->u[p], pt->v[p], pt->idepth[p], pt->idepth_zero[p]);
      const float* c = pt->color + 8 * (size_t)p; const float* w = pt->weights + 8 * (size_t)p;
      col[2 * sp] = make_float4(c[0], c[1], c[2], c[3]); col[2 * sp + 1] = make_float4(c[4], c[5], c[6], c[7]);
      wgt[2 * sp] = make_float4(w[0], w[1], w[2], w[3]); wgt[2 * sp + 1] = make_float4(w[4], w[5], w[6], w[7]);
      prior[sp] = pt->priorF[p]; delta[sp] = pt->deltaF[p]; host[sp] = pt->host_idx[p];
      rb[sp] = pos;
      for (int i = pt->res_begin[p]; i < pt->res_begin[p + 1]; ++i) ri[pos++] = ba->r_orig2sorted[pt->res_index[i]];
    }
    rb[P] = pos;
    ba->h_rb = rb; ba->h_ri = ri;
    ba->h_uvid = uvid;
    TRY(upload(ba, ba->d_uvid.p, uvid.data(), P)); TRY(upload(ba, ba->d_color.p, col.data(), 2 * (size_t)P));
    TRY(upload(ba, ba->d_weights.p, wgt.data(), 2 * (size_t)P)); TRY(upload(ba, ba->d_priorF.p, prior.data(), P));
    TRY(upload(ba, ba->d_deltaF.p, delta.data(), P)); TRY(upload(ba, ba->d_phost.p, host.data(), P));
    TRY(upload(ba, ba->d_res_begin.p, rb.data(), P + 1)); TRY(upload(ba, ba->d_res_index.p, ri.data(), R));
  }
  // ---- pack + upload residuals
  {
    std::vector<int> rp(R), rpair(R);
    std::vector<int8_t> st(R);
    std::vector<uint8_t> lin(R), act(R);
    std::vector<float> se(R);
    ba->n_linearized = 0;
    for (int k = 0; k < R; ++k) {
      int r = ba->r_sorted2orig[k];
      rp[k] = ba->p_orig2sorted[rs->point_idx[r]];
      rpair[k] = pair_of(r);
      st[k] = rs->state_state[r]; lin[k] = rs->is_linearized[r] ? 1 : 0; act[k] = rs->is_active[r] ? 1 : 0;
      se[k] = rs->state_energy ? (float)rs->state_energy[r] : 0.f;
      ba->n_linearized += lin[k];
    }
    ba->h_rpair = rpair; ba->h_rpoint = rp; ba->h_lin = lin;
    TRY(upload(ba, ba->d_rpoint.p, rp.data(), R)); TRY(upload(ba, ba->d_rpair.p, rpair.data(), R));
    TRY(upload(ba, ba->d_state.p, st.data(), R)); TRY(upload(ba, ba->d_isLin.p, lin.data(), R));
    TRY(upload(ba, ba->d_isActive.p, act.data(), R)); TRY(upload(ba, ba->d_senergy.p, se.data(), R));
  }
  HS_CUDA(cudaMemsetAsync(ba->d_J.p, 0, (size_t)JS * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_JpJdF.p, 0, recsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_slotpos.p, 0, (rsz + TOP_ROW) * sizeof(int), ctx->stream));
  TRY(upload_slotpos(ba));
  HS_CUDA(cudaMemsetAsync(ba->d_rtz.p, 0, 8 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_cpt.p, 0, 3 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_pto.p, 0, 16 * rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_newState.p, HSLAM_RES_OUT, rsz, ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_newEnergy.p, 0, rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_newEnergyWO.p, 0, rsz * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_prec.p, 0, recsz * sizeof(float), ctx->stream));   // flag 0: no active residual in any slot yet
  for (int m = 0; m < 2; ++m) {
    HS_CUDA(cudaMemsetAsync(ba->d_Hdd[m].p, 0, (P + 1) * sizeof(float), ctx->stream));
    HS_CUDA(cudaMemsetAsync(ba->d_bd[m].p, 0, (P + 1) * sizeof(float), ctx->stream));
    HS_CUDA(cudaMemsetAsync(ba->d_Hcd[m].p, 0, (P + 1) * sizeof(float4), ctx->stream));
  }
  HS_CUDA(cudaMemsetAsync(ba->d_HdiF.p, 0, (P + 1) * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_bdSumF.p, 0, (P + 1) * sizeof(float), ctx->stream));
  HS_CUDA(cudaMemsetAsync(ba->d_idh.p, 0, (P + 1) * sizeof(float), ctx->stream));
  ba->h_hostBegin = hostBegin;
  TRY(build_chunks(ba));
  ba->have_topology = true;
  fill_desc(ba);
  TRY(ba_flush(ba));
  return 0;
}

int hslam_ba_update_points(hslam_ba* ba, const float* idepth, const float* deltaF, const float* idepth_zero) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_update_points: no topology");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const int P = ba->P;
  if (P == 0) return 0;
  // caller-order arrays go up as they are; a tiny kernel scatters them into the sorted point records
  const float* srcs[3] = {idepth, idepth_zero, deltaF};
  for (int i = 0; i < 3; ++i) {
    if (!srcs[i]) continue;
    void* tmp;
    TRY(scratch_alloc(ba, (size_t)P * sizeof(float), &tmp));
    TRY(upload(ba, (float*)tmp, srcs[i], P));
    if (i < 2)
      HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_import_kernel<float><<<(P + 255) / 256, 256, 0, ctx->stream>>>(
                                            (const float*)tmp, P, ba->d_p_s2o.p, (float*)ba->d_uvid.p, 4, i == 0 ? 2 : 3)));
    else
      HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_import_kernel<float><<<(P + 255) / 256, 256, 0, ctx->stream>>>(
                                            (const float*)tmp, P, ba->d_p_s2o.p, ba->d_deltaF.p, 1, 0)));
  }
  HS_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------- launches (single window)
template <class... A>
static void lin_launch(bool apply, bool fullj, int blocks, cudaStream_t st, A... args) {
  if (apply && fullj) ba_linearize_kernel<true, true><<<blocks, LIN_THREADS, 0, st>>>(args...);
  else if (apply) ba_linearize_kernel<true, false><<<blocks, LIN_THREADS, 0, st>>>(args...);
  else if (fullj) ba_linearize_kernel<false, true><<<blocks, LIN_THREADS, 0, st>>>(args...);
  else ba_linearize_kernel<false, false><<<blocks, LIN_THREADS, 0, st>>>(args...);
}
static int launch_linearize(hslam_ba* ba, bool apply) {
  hslam_ctx* ctx = ba->ctx;
  TRY(sync_desc(ba));
  if (ba->n_lintiles > 0)
    HS_LAUNCH(ctx, HSLAM_K_BA_LINEARIZE, lin_launch(apply, ba->jmode == 0, ba->n_lintiles, ctx->stream, (const BADesc*)ba->d_desc.p,
                                                    (const int*)nullptr, 1, 0, (const int*)nullptr, 0));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static const size_t kTopSmem = std::max(TOP_STAGES * sizeof(TopStage), TOP_STAGES_L * sizeof(TopStageL)) + 128;   // + alignment slack (TMA destinations)
static int top_attr(hslam_ctx* ctx) {
  bool& done = ctx->top_attr_done;
  if (!done) {
    HS_CUDA(cudaFuncSetAttribute(ba_top_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopSmem));
    HS_CUDA(cudaFuncSetAttribute(ba_top_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopSmem));
    HS_CUDA(cudaFuncSetAttribute(ba_top_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTopSmem));
    done = true;
  }
  return 0;
}
// modes: bit 0 = A, bit 1 = L
static int launch_top(hslam_ba* ba, int modes, bool fold = true) {
  hslam_ctx* ctx = ba->ctx;
  TRY(sync_desc(ba));
  TRY(top_attr(ctx));
  HS_ARG(ba->tmap_ok, "hslam_ba: cuTensorMapEncodeTiled failed for the Jacobian tensor maps (driver too old?)");
  if (ba->n_topblocks[0] > 0 && (modes & 1))
    HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<0><<<ba->n_topblocks[0], 2 * TOP_THREADS, kTopSmem, ctx->stream>>>(ba->d_desc.p, nullptr, 1)));
  if (ba->n_topblocks[1] > 0 && (modes & 2))
    HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<1><<<ba->n_topblocks[1], 2 * TOP_THREADS, kTopSmem, ctx->stream>>>(ba->d_desc.p, nullptr, 1)));
  if (ba->n_topblocks[2] > 0 && (modes & 4))   // addPoint<2> over the points flagged in desc.marg
    HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<2><<<ba->n_topblocks[2], 2 * TOP_THREADS, kTopSmem, ctx->stream>>>(ba->d_desc.p, nullptr, 1)));
  if (fold)   // chunk partials -> dense fp64 accumulators (the fused paths let the Schur fold launch do it)
    HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_SC, (ba_fold_sc_kernel<<<fold_top_blocks(ba->n, modes), 256, 0, ctx->stream>>>(ba->d_desc.p, nullptr, 1, modes, 0)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

// Schur kernel variant of a window: 1 = 8x8 tiles, 8 point slices (n <= 7); 2 = 8x8 tiles, 4 slices (n <= 10);
// 5 = the 4x4-tile kernel with up to five tiles per thread (n <= 16).  HSLAM_SCHUR_4X4=1 (A/B measurements) selects the
// 4x4-tile kernel for every n: 11 / 12 / 15 = one / two / five tiles per thread.
static int schur_maxt(int LD) {
  static const bool old4 = getenv("HSLAM_SCHUR_4X4") != nullptr;
  const int n = LD / 8 - 1;
  if (!old4 && n <= 10) return n <= 7 ? 1 : 2;
  int NTR = LD / 4 - 2, t = NTR * (NTR + 1) / 2;
  return old4 ? (t <= 128 ? 11 : (t <= 256 ? 12 : 15)) : 5;
}
// 4x4 kernel: rows + staged records; the final exchange of the two point halves needs MAXT * 128 * 16 floats
static size_t schur_smem(int LD, int maxres) {
  const int k = schur_maxt(LD);
  if (k <= 2) return sc8_smem_floats(LD, maxres) * sizeof(float);
  size_t fl = std::max(sc_smem_floats(LD, maxres), (size_t)(k % 10) * 128 * 16 + 4);
  return fl * sizeof(float);
}
template <class F> static int schur_dispatch(int maxt, F f) {
  if (maxt == 1) return f(ba_schur8_kernel<8>);
  if (maxt == 2) return f(ba_schur8_kernel<4>);
  if (maxt == 11) return f(ba_schur_kernel<1>);
  if (maxt == 12) return f(ba_schur_kernel<2>);
  return f(ba_schur_kernel<5>);
}
static int schur_attr(hslam_ctx* ctx, size_t smem) {
  size_t& configured = ctx->schur_smem_configured;
  if (smem > configured) {
    for (int maxt : {1, 2, 5, 11, 12})
      TRY(schur_dispatch(maxt, [&](auto k) -> int {
        HS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
      }));
    configured = smem;
  }
  return 0;
}
static void schur_launch(int maxt, int blocks, size_t smem, cudaStream_t st, const BADesc* d, const int* pf, int n) {
  schur_dispatch(maxt, [&](auto k) -> int { k<<<blocks, SC_THREADS, smem, st>>>(d, pf, n); return 0; });
}

static int launch_schur(hslam_ba* ba, int top_modes = 0) {
  hslam_ctx* ctx = ba->ctx;
  TRY(sync_desc(ba));
  size_t smem = schur_smem(ba->LD, ba->maxres);
  TRY(schur_attr(ctx, smem));
  if (ba->n_ptchunks > 0)
    HS_LAUNCH(ctx, HSLAM_K_BA_SCHUR, schur_launch(schur_maxt(ba->LD), ba->n_ptchunks, smem, ctx->stream, ba->d_desc.p, nullptr, 1));
  HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_SC, (ba_fold_sc_kernel<<<ba->n * fold_sc_blocks(ba->LD) + (top_modes ? fold_top_blocks(ba->n, top_modes) : 0), 256, 0, ctx->stream>>>(ba->d_desc.p, nullptr, 1, top_modes, 1)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static int launch_stitch(hslam_ba* ba, int sys_begin, int sys_end, int with_energy = 0) {
  hslam_ctx* ctx = ba->ctx;
  TRY(sync_desc(ba));
  int blocks = 3 * (ba->n + 1) * (ba->n + 1);
  HS_LAUNCH(ctx, HSLAM_K_BA_STITCH, (ba_stitch_kernel<8><<<blocks, 256, 0, ctx->stream>>>(ba->d_desc.p, nullptr, 1, sys_begin, sys_end, with_energy)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static int launch_point_sum(hslam_ba* ba, int modes) {
  hslam_ctx* ctx = ba->ctx;
  int pblocks = (ba->P + 127) / 128;
  if (pblocks > 0)
    HS_LAUNCH(ctx, HSLAM_K_BA_POINT_SUM,
              (ba_point_sum_kernel<<<pblocks, 128, 0, ctx->stream>>>(ba->hdesc.r, ba->hdesc.p, modes & 1, (modes >> 1) & 1, ba->maxres)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

static int queue_Hb(hslam_ba* ba, int sys, double* H, double* b) {
  const size_t NN = (size_t)ba->N * ba->N, per = NN + ba->N;
  const int N = ba->N;
  if (!H && !b) return 0;
  return download_then<double>(ba, ba->d_Hb.p + per * sys, per, [H, b, NN, N](const double* s) {
    if (H) memcpy(H, s, NN * sizeof(double));
    if (b) memcpy(b, s + NN, N * sizeof(double));
  });
}

static int enqueue_accumulate_all(hslam_ba* ba, int with_energy = 0) {
  const int modes = 1 | (ba->n_linearized != 0 ? 2 : 0);
  TRY(launch_top(ba, modes, false));
  TRY(launch_schur(ba, 3));   // the fold launch also folds the top accumulators of A and L (L reads as zero without chunks)
  TRY(launch_stitch(ba, 0, 3, with_energy));
  return 0;
}

extern "C" {

int hslam_ba_linearize(hslam_ba* ba, int fixLinearization, double* energy_sum, int8_t* newState, float* newEnergy,
                       float* newEnergyWithOutlier, float* centerProjectedTo, float* projectedTo) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_linearize: window/topology not set");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(launch_linearize(ba, fixLinearization != 0));
  if (energy_sum) {
    HS_LAUNCH(ba->ctx, HSLAM_K_BA_ENERGY, (ba_energy_kernel<<<1, 256, 0, ba->ctx->stream>>>(ba->d_desc.p)));
    HS_CUDA(cudaGetLastError());
  }
  TRY(download(ba, energy_sum, ba->d_energy.p, 1));
  TRY(download_res(ba, ba->d_newState.p, 1, newState));
  TRY(download_res(ba, ba->d_newEnergy.p, 1, newEnergy));
  TRY(download_res(ba, ba->d_newEnergyWO.p, 1, newEnergyWithOutlier));
  TRY(download_res(ba, ba->d_cpt.p, 3, centerProjectedTo));
  TRY(download_res(ba, ba->d_pto.p, 16, projectedTo));
  return end_call(ba);
}

int hslam_ba_apply_res(hslam_ba* ba) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_apply_res: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  int blocks = (ba->R + 255) / 256;
  if (blocks < 1) blocks = 1;
  HS_LAUNCH(ba->ctx, HSLAM_K_BA_APPLY, (ba_apply_kernel<<<blocks, 256, 0, ba->ctx->stream>>>(ba->hdesc.r)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

int hslam_ba_new_frame_energy_th(hslam_ba* ba, int newest, float* th_out) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_new_frame_energy_th: window/topology not set");
  HS_ARG(newest >= 0 && newest < ba->n, "hslam_ba_new_frame_energy_th: frame index out of range");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  TRY(sync_desc(ba));
  TRY(ba->d_th.ensure(1));
  HS_LAUNCH(ctx, HSLAM_K_BA_ENERGY_TH, (ba_energy_th_kernel<<<1, TH_THREADS, 0, ctx->stream>>>(ba->d_desc.p, newest, ba->d_th.p, nullptr, nullptr)));
  HS_CUDA(cudaGetLastError());
  // the device descriptor already holds the new threshold; the host mirror follows once the value is known
  TRY(download_then<float>(ba, ba->d_th.p, 1, [ba, newest, th_out](const float* s) {
    ba->hdesc.W.frameEnergyTH[newest] = *s;
    ba->desc_version++;
    if (th_out) *th_out = *s;
  }));
  return end_call(ba);
}

int hslam_ba_fix_linearization(hslam_ba* ba, const int32_t* ids, int count) {
  HS_ARG(ba && ba->have_topology && (count == 0 || ids), "hslam_ba_fix_linearization: bad argument");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  if (count == 0) return 0;
  void* s;
  TRY(ba->d_ids.ensure(count));
  TRY(arena_alloc(ba, (size_t)count * sizeof(int), &s));
  int* sorted = (int*)s;
  for (int i = 0; i < count; ++i) {
    HS_ARG(ids[i] >= 0 && ids[i] < ba->R, "hslam_ba_fix_linearization: id out of range");
    sorted[i] = ba->r_orig2sorted[ids[i]];
  }
  HS_CUDA(cudaMemcpyAsync(ba->d_ids.p, sorted, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (ba->jmode != 0) {
    // GN-loop mode: the per-pixel rows of the Jacobian were not stored by the last pass; re-form them for these
    // residuals at the window's current state (the reference re-linearises right before fixLinearizationF, Mapping.cpp:286-293)
    TRY(sync_desc(ba));
    HS_LAUNCH(ctx, HSLAM_K_BA_LINEARIZE, lin_launch(false, true, (count + LIN_THREADS - 1) / LIN_THREADS, ctx->stream, (const BADesc*)ba->d_desc.p,
                                                    (const int*)nullptr, 1, 0, (const int*)ba->d_ids.p, count));
  }
  HS_LAUNCH(ctx, HSLAM_K_BA_FIXLIN, (ba_fixlin_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(
                                        ba->hdesc.W, ba->hdesc.r, ba->hdesc.p, ba->d_ids.p, count)));
  HS_CUDA(cudaGetLastError());
  int fresh = 0;
  for (int i = 0; i < count; ++i) { uint8_t& f = ba->h_lin[sorted[i]]; fresh += f ? 0 : 1; f = 1; }
  ba->n_linearized += fresh;
  // keep the sorted order's invariant (linearised residuals behind the others of their pair) and re-cut the schedules
  return regroup(ba);
}

int hslam_ba_set_jacobian_mode(hslam_ba* ba, int mode) {
  HS_ARG(ba && (mode == 0 || mode == 1), "hslam_ba_set_jacobian_mode: mode must be 0 (full) or 1 (GN loop)");
  if (ba->jmode == mode) return 0;
  ba->jmode = mode;
  ba->hdesc.jmode = mode;
  ba->desc_version++;
  return 0;
}

int hslam_ba_accumulate(hslam_ba* ba, int which, double* H, double* b, int* nres) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_accumulate: window/topology not set");
  HS_ARG(which >= 0 && which <= 2, "hslam_ba_accumulate: which must be HSLAM_ACC_A/L/SC");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  if (which == HSLAM_ACC_A || which == HSLAM_ACC_L) {
    if (which == HSLAM_ACC_A || ba->n_linearized != 0) TRY(launch_top(ba, 1 << which));
    TRY(launch_point_sum(ba, 1 << which));   // side effects of addPoint<mode> on the points
    TRY(launch_stitch(ba, which, which + 1));
    TRY(queue_Hb(ba, which, H, b));
    TRY(download(ba, nres, ba->d_nres.p + which, 1));
  } else {
    TRY(launch_schur(ba));
    TRY(launch_stitch(ba, 2, 3));
    TRY(queue_Hb(ba, 2, H, b));
  }
  return end_call(ba);
}

int hslam_ba_accumulate_all_device(hslam_ba* ba, double** Hb_dev_out, int* count_out) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_accumulate_all: window/topology not set");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(enqueue_accumulate_all(ba));
  if (Hb_dev_out) *Hb_dev_out = ba->d_Hb.p;
  if (count_out) *count_out = 3 * (ba->N * ba->N + ba->N);
  return 0;
}

// ---- NCCL, resolved at run time (no link-time dependency): the library named by HSLAM_NCCL_LIB, else whatever libnccl
// the process has already loaded (torch's), else libnccl.so.2.  Communicator creation goes through the SAME handle, so a
// comm made by hslam_nccl_comm_init is used by the library version that made it.
struct NcclApi {
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ struct Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
struct Id128 { char b[128]; };
static NcclApi* nccl_api() {
  static NcclApi api = []() {
    NcclApi a;
    void* h = nullptr;
    if (const char* path = getenv("HSLAM_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    auto sym = [&](const char* name) -> void* {
      void* p = h ? dlsym(h, name) : dlsym(RTLD_DEFAULT, name);
      return p;
    };
    if (!h && !dlsym(RTLD_DEFAULT, "ncclAllReduce")) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    a.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, Id128, int))sym("ncclCommInitRank");
    a.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
    a.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
    return a;
  }();
  return &api;
}
static int nccl_fail(const char* what, int rc) {
  NcclApi* a = nccl_api();
  hs::set_error(std::string(what) + ": NCCL error " + std::to_string(rc) + (a->GetErrorString ? std::string(" (") + a->GetErrorString(rc) + ")" : ""));
  return HSLAM_ERR_NCCL;
}
#define HS_NCCL_READY(what) do { if (!nccl_api()->ok) { hs::set_error(std::string(what) + ": libnccl not found (set HSLAM_NCCL_LIB)"); return HSLAM_ERR_NCCL; } } while (0)

int hslam_nccl_get_unique_id(char id_out[128]) {
  HS_ARG(id_out, "hslam_nccl_get_unique_id: null");
  HS_NCCL_READY("hslam_nccl_get_unique_id");
  int rc = nccl_api()->GetUniqueId(id_out);
  return rc ? nccl_fail("ncclGetUniqueId", rc) : 0;
}
int hslam_nccl_comm_init(hslam_ctx* ctx, const char id[128], int n_ranks, int rank, void** comm_out) {
  HS_ARG(ctx && id && comm_out && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "hslam_nccl_comm_init: bad argument");
  HS_NCCL_READY("hslam_nccl_comm_init");
  HS_CUDA(cudaSetDevice(ctx->device));
  Id128 uid;
  memcpy(uid.b, id, 128);
  int rc = nccl_api()->CommInitRank(comm_out, n_ranks, uid, rank);
  return rc ? nccl_fail("ncclCommInitRank", rc) : 0;
}
int hslam_nccl_comm_destroy(void* comm) {
  if (!comm) return 0;
  HS_NCCL_READY("hslam_nccl_comm_destroy");
  int rc = nccl_api()->CommDestroy(comm);
  return rc ? nccl_fail("ncclCommDestroy", rc) : 0;
}

// the all-reduce itself: the stitched systems plus three doubles behind them — energy (stats[0] of linearizeAll) and
// resInA / resInL, which the stitch / energy kernels write there
static int enqueue_allreduce(hslam_ba* ba, void* nccl_comm) {
  const size_t count = 3 * ((size_t)ba->N * ba->N + ba->N) + 3;
  // ncclDouble = 8, ncclSum = 0 (nccl.h)
  int rc = nccl_api()->AllReduce(ba->d_Hb.p, ba->d_Hb.p, count, 8, 0, nccl_comm, ba->ctx->stream);
  return rc ? nccl_fail("ncclAllReduce", rc) : 0;
}

int hslam_ba_allreduce(hslam_ba* ba, void* nccl_comm) {
  HS_ARG(ba && nccl_comm && ba->have_window, "hslam_ba_allreduce: bad argument");
  HS_NCCL_READY("hslam_ba_allreduce");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  return enqueue_allreduce(ba, nccl_comm);
}

static int iteration_impl(hslam_ba* ba, void* nccl_comm) {
  hslam_ctx* ctx = ba->ctx;
  static const bool no_graph = getenv("HSLAM_NO_GRAPH") != nullptr;
  if (ctx->profile || no_graph) {
    TRY(launch_linearize(ba, true));
    TRY(enqueue_accumulate_all(ba, 1));   // the stitch also sums the energy partials of the linearise pass
    return nccl_comm ? enqueue_allreduce(ba, nccl_comm) : 0;
  }
  // The launch chain (linearise, top A [+ L], Schur, fold, stitch) as ONE graph launch: a window that runs
  // alone is latency-bound, the gaps between the dependent launches cost more than the kernels.
  TRY(sync_desc(ba));   // descriptor upload stays outside the graph
  const long long key[12] = {ba->n_lintiles, ba->n_topblocks[0], ba->n_topblocks[1], ba->n_ptchunks, ba->n, ba->LD,
                             ba->maxres * 4 + ba->jmode * 2 + (ba->n_linearized != 0), 0,
                             (long long)(size_t)ba->d_Hb.p, (long long)(size_t)ba->d_desc.p, ba->P, ba->R};
  if (!ba->g_iter || memcmp(key, ba->g_key, sizeof(key)) != 0) {
    TRY(top_attr(ctx));
    TRY(schur_attr(ctx, schur_smem(ba->LD, ba->maxres)));
    if (ba->g_iter) { cudaGraphExecDestroy(ba->g_iter); ba->g_iter = nullptr; }
    const int64_t l0 = ctx->launches;
    HS_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = launch_linearize(ba, true);
    if (!rc) rc = enqueue_accumulate_all(ba, 1);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    HS_CUDA(e);
    ba->g_kernels = (int)(ctx->launches - l0);
    ctx->launches = l0;
    HS_CUDA(cudaGraphInstantiate(&ba->g_iter, g, 0));
    cudaGraphDestroy(g);
    memcpy(ba->g_key, key, sizeof(key));
  }
  HS_CUDA(cudaGraphLaunch(ba->g_iter, ctx->stream));
  ctx->launches += ba->g_kernels;
  // the all-reduce follows the graph on the same stream (NCCL is kept out of the capture: its first collective on a
  // communicator sets up connections, which is not capturable)
  return nccl_comm ? enqueue_allreduce(ba, nccl_comm) : 0;
}

int hslam_ba_iteration_device(hslam_ba* ba, double** Hb_dev_out, int* count_out) {
  HS_ARG(ba && ba->have_topology && ba->have_window, "hslam_ba_iteration: window/topology not set");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  if (Hb_dev_out) *Hb_dev_out = ba->d_Hb.p;
  if (count_out) *count_out = 3 * (ba->N * ba->N + ba->N);
  return iteration_impl(ba, nullptr);
}

int hslam_ba_iteration_allreduce(hslam_ba* ba, void* nccl_comm) {
  HS_ARG(ba && nccl_comm && ba->have_topology && ba->have_window, "hslam_ba_iteration_allreduce: bad argument");
  HS_NCCL_READY("hslam_ba_iteration_allreduce");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  return iteration_impl(ba, nccl_comm);
}

int hslam_ba_download_Hb(hslam_ba* ba, double* out) {
  HS_ARG(ba && out && ba->have_window, "hslam_ba_download_Hb: bad argument");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(download(ba, out, ba->d_Hb.p, 3 * ((size_t)ba->N * ba->N + ba->N)));
  return end_call(ba);
}

int hslam_ba_download_stats(hslam_ba* ba, double* energy, int* nres2) {
  HS_ARG(ba && ba->have_window, "hslam_ba_download_stats: bad argument");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  // the three doubles behind the systems: this rank's values, or the window's totals after hslam_ba_allreduce
  const size_t count = 3 * ((size_t)ba->N * ba->N + ba->N);
  TRY(download_then<double>(ba, ba->d_Hb.p + count, 3, [energy, nres2](const double* s) {
    if (energy) *energy = s[0];
    if (nres2) { nres2[0] = (int)(s[1] + 0.5); nres2[1] = (int)(s[2] + 0.5); }
  }));
  return end_call(ba);
}

int hslam_ba_accumulate_all(hslam_ba* ba, double* Hb_out_host, int* nres2) {
  HS_ARG(Hb_out_host, "hslam_ba_accumulate_all: null output");
  TRY(hslam_ba_accumulate_all_device(ba, nullptr, nullptr));
  TRY(download(ba, Hb_out_host, ba->d_Hb.p, 3 * ((size_t)ba->N * ba->N + ba->N)));
  TRY(download(ba, nres2, ba->d_nres.p, 2));
  return end_call(ba);
}

int hslam_ba_point_results(hslam_ba* ba, float* HdiF, float* bdSumF, float* idepth_hessian, float* Hdd_accAF,
                           float* bd_accAF, float* Hcd_accAF, float* Hdd_accLF, float* bd_accLF, float* Hcd_accLF) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_point_results: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(download_pts(ba, ba->d_HdiF.p, 1, HdiF)); TRY(download_pts(ba, ba->d_bdSumF.p, 1, bdSumF));
  TRY(download_pts(ba, ba->d_idh.p, 1, idepth_hessian));
  TRY(download_pts(ba, ba->d_Hdd[0].p, 1, Hdd_accAF)); TRY(download_pts(ba, ba->d_bd[0].p, 1, bd_accAF));
  TRY(download_pts(ba, (const float*)ba->d_Hcd[0].p, 4, Hcd_accAF));
  TRY(download_pts(ba, ba->d_Hdd[1].p, 1, Hdd_accLF)); TRY(download_pts(ba, ba->d_bd[1].p, 1, bd_accLF));
  TRY(download_pts(ba, (const float*)ba->d_Hcd[1].p, 4, Hcd_accLF));
  return end_call(ba);
}

int hslam_ba_residual_results(hslam_ba* ba, int8_t* state_state, uint8_t* is_active, float* JpJdF, float* J74,
                              float* res_toZeroF) {
  HS_ARG(ba && ba->have_topology, "hslam_ba_residual_results: no topology");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(download_res(ba, ba->d_state.p, 1, state_state));
  TRY(download_res(ba, ba->d_isActive.p, 1, is_active));
  if (JpJdF && ba->R > 0) {   // [R][8] records in sorted order -> caller order
    void* tmp;
    TRY(scratch_alloc(ba, (size_t)ba->R * 8 * sizeof(float), &tmp));
    HS_LAUNCH(ba->ctx, HSLAM_K_BA_EXPORT, (ba_export_jpjd_kernel<<<(ba->R + 255) / 256, 256, 0, ba->ctx->stream>>>(
                                              ba->hdesc.r, ba->d_r_s2o.p, (float*)tmp)));
    TRY(download(ba, JpJdF, (const float*)tmp, (size_t)ba->R * 8));
  }
  TRY(download_res(ba, ba->d_rtz.p, 8, res_toZeroF));
  TRY(download_res(ba, ba->d_J.p, 74, J74));
  return end_call(ba);
}

int hslam_ba_marginalize_points(hslam_ba* ba, const int32_t* point_ids, int count, float prior_scale, double* M, double* Mb,
                                double* Msc, double* Mbsc, int* nres) {
  HS_ARG(ba && ba->have_topology && ba->have_window && (count == 0 || point_ids), "hslam_ba_marginalize_points: bad argument");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const int P = ba->P;
  std::vector<unsigned char> flags((size_t)P + 1, 0);
  for (int i = 0; i < count; ++i) {
    HS_ARG(point_ids[i] >= 0 && point_ids[i] < P, "hslam_ba_marginalize_points: point id out of range");
    flags[ba->p_orig2sorted[point_ids[i]]] = 1;
  }
  TRY(ba->d_marg.ensure((size_t)P + 1));
  TRY(upload(ba, ba->d_marg.p, flags.data(), (size_t)P + 1));
  if (P > 0)
    HS_LAUNCH(ctx, HSLAM_K_BA_FIXLIN, (ba_scale_prior_kernel<<<(P + 255) / 256, 256, 0, ctx->stream>>>(ba->hdesc.p, ba->d_marg.p, prior_scale)));
  // the pass runs on a descriptor with the marginalisation switches set: flagged points only, no prior shift in the
  // Schur scalars, no priors in the stitch (stitchDouble(..., usePrior=false), EnergyFunctional.cpp:578-579)
  const BADesc saved = ba->hdesc;
  ba->hdesc.marg = ba->d_marg.p; ba->hdesc.shift_prior = 0;
  ba->hdesc.S.add_priors = 0; ba->hdesc.S.has_L = 1; ba->hdesc.S.L_set = 2;
  ba->desc_version++;
  TRY(launch_top(ba, 4, false));
  TRY(launch_schur(ba, 4));
  TRY(launch_stitch(ba, 1, 3));
  TRY(queue_Hb(ba, 1, M, Mb));
  TRY(queue_Hb(ba, 2, Msc, Mbsc));
  TRY(download(ba, nres, ba->d_nres.p + 1, 1));
  TRY(ba_flush(ba));   // the descriptor is restored only after the pass has run
  ba->hdesc = saved;
  ba->desc_version++;
  return 0;
}

int hslam_ba_calc_l_energy(hslam_ba* ba, double* energy_out) {
  HS_ARG(ba && energy_out && ba->have_topology && ba->have_window, "hslam_ba_calc_l_energy: bad argument");
  hslam_ctx* ctx = ba->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  TRY(sync_desc(ba));
  const int blocks = std::max(1, (ba->P + 127) / 128);
  void* tmp;
  TRY(scratch_alloc(ba, (size_t)blocks * sizeof(double), &tmp));
  HS_LAUNCH(ctx, HSLAM_K_BA_ENERGY, (ba_lenergy_kernel<<<blocks, 128, 0, ctx->stream>>>(ba->d_desc.p, (double*)tmp)));
  HS_CUDA(cudaGetLastError());
  std::vector<double> part(blocks);
  TRY(download(ba, part.data(), (const double*)tmp, (size_t)blocks));
  TRY(ba_flush(ba));
  // frame and calibration priors (EnergyFunctional.cpp:355-358), then the point partials in block order
  double E = 0.0;
  for (int f = 0; f < ba->n; ++f)
    for (int i = 0; i < 8; ++i) E += ba->h_fdprior[8 * f + i] * ba->h_fprior[8 * f + i] * ba->h_fdprior[8 * f + i];
  for (int i = 0; i < 4; ++i) E += (double)(ba->hdesc.W.cDeltaF[i] * (float)ba->hdesc.S.cPrior[i] * ba->hdesc.W.cDeltaF[i]);
  for (double v : part) E += v;
  *energy_out = E;
  return 0;
}

int hslam_ba_resubstitute(hslam_ba* ba, const double* x, float* step_out) {
  HS_ARG(ba && x && ba->have_topology && ba->have_window, "hslam_ba_resubstitute: bad argument");
  HS_CUDA(cudaSetDevice(ba->ctx->device));
  TRY(sync_desc(ba));
  TRY(upload(ba, ba->d_x.p, x, (size_t)ba->N));
  // xAd = x_h^T adHostF + x_t^T adTargetF (EnergyFunctional.cpp:228-238) is formed on the device
  HS_LAUNCH(ba->ctx, HSLAM_K_BA_RESUB, (ba_xad_kernel<<<1, 256, 0, ba->ctx->stream>>>(ba->d_desc.p)));
  int blocks = (ba->P + 127) / 128;
  if (blocks > 0)
    HS_LAUNCH(ba->ctx, HSLAM_K_BA_RESUB, (ba_resub_kernel<<<blocks, 128, 0, ba->ctx->stream>>>(ba->d_desc.p, nullptr, 1, nullptr, nullptr, 0)));
  HS_CUDA(cudaGetLastError());
  TRY(download_pts(ba, ba->d_step.p, 1, step_out));
  return end_call(ba);
}

}  // extern "C"

// ---------------------------------------------------------------- device-resident inputs of makeCoarseDepthL0
// For the listed residuals (caller ids; the residual of each contributing point into the newest keyframe) gather
// centerProjectedTo (3 floats) and the HdiF of the residual's point into device arrays: the tracker builds its reference
// clouds from them without a host round trip (tracker.cu, hslam_tracker_make_coarse_depth_from_ba).
namespace hs {
__global__ void ba_gather_coarse_kernel(BARes r, BAPts pts, const int* __restrict__ sorted_ids, int count, float* __restrict__ cpt3,
                                        float* __restrict__ hdif) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int k = sorted_ids[i];
  cpt3[3 * i] = r.cpt[k]; cpt3[3 * i + 1] = r.cpt[r.rs + k]; cpt3[3 * i + 2] = r.cpt[2 * r.rs + k];
  hdif[i] = pts.HdiF[r.point[k]];
}
}  // namespace hs
int hs_ba_gather_coarse_inputs(hslam_ba* ba, hslam_ctx* ctx, const int32_t* residual_ids, int count, float* d_cpt3, float* d_hdif) {
  HS_ARG(ba && ba->have_topology && ba->ctx == ctx, "make_coarse_depth_from_ba: window handle missing or from another context");
  if (count == 0) return 0;
  void* s;
  TRY(ba->d_ids.ensure(count));
  TRY(arena_alloc(ba, (size_t)count * sizeof(int), &s));
  int* sorted = (int*)s;
  for (int i = 0; i < count; ++i) {
    HS_ARG(residual_ids[i] >= 0 && residual_ids[i] < ba->R, "make_coarse_depth_from_ba: residual id out of range");
    sorted[i] = ba->r_orig2sorted[residual_ids[i]];
  }
  HS_CUDA(cudaMemcpyAsync(ba->d_ids.p, sorted, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  HS_LAUNCH(ctx, HSLAM_K_COARSE_DEPTH, (hs::ba_gather_coarse_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(
                                           ba->hdesc.r, ba->hdesc.p, ba->d_ids.p, count, d_cpt3, d_hdif)));
  HS_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- batches of independent windows
struct hslam_ba_batch {
  hslam_ctx* ctx = nullptr;
  std::vector<hslam_ba*> items;
  std::vector<uint64_t> versions;
  DevBuf<BADesc> d_descs;
  DevBuf<int> d_prefix;          // NPF tables of (count+1): linearize, top A, schur, stitch, schur fold, resubstitute, top L
  BADesc* h_descs = nullptr;     // pinned
  int* h_prefix = nullptr;       // pinned
  cudaEvent_t copied = nullptr;
  cudaEvent_t ev_grp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // step_host: per window-group fork points
  int blocks[7] = {0, 0, 0, 0, 0, 0, 0};
  int uniform[7] = {0, 0, 0, 0, 0, 0, 0};
  bool fullJ = true;             // some member stores full Jacobians: the batch launch does for all   // blocks per window of a stage when every window has the same count, else 0
  size_t smem = 0;
  int maxt = 1, ni = 2;
  bool any_L = false;
  // host API staging: one pinned block up, one down, their device twins, and the per-window offset table
  char* h_up = nullptr; char* h_dn = nullptr;
  size_t h_up_cap = 0, h_dn_cap = 0;
  DevBuf<char> d_up, d_dn;
  DevBuf<BAIO> d_io;
  std::vector<BAIO> io;
  // CUDA graph of the launch chain (re-captured when batch_refresh changes a launch parameter)
  cudaGraphExec_t graph_iter = nullptr, graph_acc = nullptr;
  uint64_t refresh_version = 1, graph_iter_version = 0, graph_acc_version = 0;
  int graph_iter_kernels = 0, graph_acc_kernels = 0;
  size_t up_tab_begin = 0, up_dF_begin = 0, up_idz_begin = 0, up_end = 0;   // H2D block regions: [0,tab) offset table + x slots, [tab,idz) window chunks, [idz,end) idepth_zero slots
  std::vector<size_t> up_win_begin;   // start of window i's chunk [tables | idepth | deltaF]; [n] = up_idz_begin
  cudaStream_t up_stream = nullptr;   // step_host: uploads + import of group g+1 run here while group g is linearised
  cudaEvent_t ev_up[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_begin = nullptr;
  int th_newest = -1;                 // >= 0: hslam_ba_batch_iteration also runs setNewFrameEnergyTH for this frame of every window
  cudaEvent_t ev_th[2] = {nullptr, nullptr};
  size_t dn_state_end = 0, dn_res_end = 0, dn_sys_end = 0, dn_end = 0;   // D2H block regions: [0,res) per-residual arrays, [res,sys) systems, [sys,end) steps
};
static const int NPF = 7;

static int batch_refresh(hslam_ba_batch* B) {
  const int n = (int)B->items.size();
  bool dirty = false;
  for (int i = 0; i < n; ++i) {
    hslam_ba* ba = B->items[i];
    HS_ARG(ba->have_window && ba->have_topology, "hslam_ba_batch: a window has no window/topology set");
    HS_ARG(ba->tmap_ok, "hslam_ba_batch: cuTensorMapEncodeTiled failed for the Jacobian tensor maps (driver too old?)");
    if (B->versions[i] != ba->desc_version) dirty = true;
  }
  if (!dirty) return 0;
  HS_CUDA(cudaEventSynchronize(B->copied));
  int* pf[NPF];
  for (int k = 0; k < NPF; ++k) { pf[k] = B->h_prefix + k * (n + 1); pf[k][0] = 0; }
  const size_t smem0 = B->smem; const int maxt0 = B->maxt, ni0 = B->ni; const bool anyL0 = B->any_L;
  B->smem = 0; B->maxt = 0; B->ni = 2; B->any_L = false;
  bool fullJ = false;
  for (int i = 0; i < n; ++i) {
    hslam_ba* ba = B->items[i];
    B->h_descs[i] = ba->hdesc;
    B->versions[i] = ba->desc_version;
    int lin = ba->n_lintiles;
    int top = ba->n_topblocks[0];
    B->any_L = B->any_L || ba->n_linearized != 0;
    fullJ = fullJ || ba->jmode == 0;
    pf[6][i + 1] = pf[6][i] + ba->n_topblocks[1];
    pf[0][i + 1] = pf[0][i] + lin;
    pf[1][i + 1] = pf[1][i] + top;
    pf[2][i + 1] = pf[2][i] + ba->n_ptchunks;
    pf[3][i + 1] = pf[3][i] + 3 * (ba->n + 1) * (ba->n + 1);
    pf[4][i + 1] = pf[4][i] + ba->n * fold_sc_blocks(ba->LD) + fold_top_blocks(ba->n, 3);   // Schur fold + top folds (A, L)
    pf[5][i + 1] = pf[5][i] + (ba->P + 127) / 128;
    B->smem = std::max(B->smem, schur_smem(ba->LD, ba->maxres));
    B->maxt = std::max(B->maxt, schur_maxt(ba->LD));
  }
  // launch parameters of the captured graphs: a change of descriptor CONTENT alone (new thresholds, new camera scalars)
  // does not invalidate them
  int blocks0[NPF], uniform0[NPF];
  for (int k = 0; k < NPF; ++k) { blocks0[k] = B->blocks[k]; uniform0[k] = B->uniform[k]; }
  for (int k = 0; k < NPF; ++k) {
    B->blocks[k] = pf[k][n];
    B->uniform[k] = pf[k][1];
    for (int i = 1; i < n; ++i) if (pf[k][i + 1] - pf[k][i] != pf[k][1]) B->uniform[k] = 0;
    if (n < 2 || pf[k][1] == 0) B->uniform[k] = 0;
  }
  bool launch_changed = smem0 != B->smem || maxt0 != B->maxt || ni0 != B->ni || anyL0 != B->any_L || fullJ != B->fullJ;
  B->fullJ = fullJ;
  for (int k = 0; k < NPF; ++k) launch_changed |= blocks0[k] != B->blocks[k] || uniform0[k] != B->uniform[k];
  HS_CUDA(cudaMemcpyAsync(B->d_descs.p, B->h_descs, n * sizeof(BADesc), cudaMemcpyHostToDevice, B->ctx->stream));
  HS_CUDA(cudaMemcpyAsync(B->d_prefix.p, B->h_prefix, NPF * (n + 1) * sizeof(int), cudaMemcpyHostToDevice, B->ctx->stream));
  HS_CUDA(cudaEventRecord(B->copied, B->ctx->stream));
  if (launch_changed) B->refresh_version++;
  return 0;
}

extern "C" {

int hslam_ba_batch_create(hslam_ctx* ctx, hslam_ba* const* items, int count, hslam_ba_batch** out) {
  HS_ARG(ctx && items && out && count >= 1, "hslam_ba_batch_create: bad argument");
  HS_CUDA(cudaSetDevice(ctx->device));
  hslam_ba_batch* B = new hslam_ba_batch();
  B->ctx = ctx;
  for (int i = 0; i < count; ++i) {
    HS_ARG(items[i] && items[i]->ctx == ctx, "hslam_ba_batch_create: window missing or from another context");
    B->items.push_back(items[i]);
  }
  // the members share the SMs: fewer, larger chunks per window (~3 waves of blocks for the whole batch)
  for (int i = 0; i < count; ++i)
    if (items[i]->target_blocks == 0) {
      int tb = std::max(4, 4 * ctx->sm_count / count);   // ~two full waves of the 2-blocks-per-SM kernels for the whole batch
      if (const char* e = getenv("HSLAM_BATCH_BLOCKS")) tb = std::max(1, atoi(e));   // tuning override
      TRY(hslam_ba_set_parallelism(items[i], tb));
    }
  B->versions.assign(count, 0);
  TRY(B->d_descs.ensure(count));
  TRY(B->d_prefix.ensure(NPF * (size_t)(count + 1)));
  TRY(B->d_io.ensure(count));
  B->io.resize(count);
  HS_CUDA(cudaMallocHost((void**)&B->h_descs, count * sizeof(BADesc)));
  HS_CUDA(cudaMallocHost((void**)&B->h_prefix, NPF * (size_t)(count + 1) * sizeof(int)));
  HS_CUDA(cudaEventCreateWithFlags(&B->copied, cudaEventDisableTiming));
  for (int g = 0; g < 8; ++g) HS_CUDA(cudaEventCreateWithFlags(&B->ev_grp[g], cudaEventDisableTiming));
  for (int g = 0; g < 2; ++g) HS_CUDA(cudaEventCreateWithFlags(&B->ev_th[g], cudaEventDisableTiming));
  for (int g = 0; g < 8; ++g) HS_CUDA(cudaEventCreateWithFlags(&B->ev_up[g], cudaEventDisableTiming));
  HS_CUDA(cudaEventCreateWithFlags(&B->ev_begin, cudaEventDisableTiming));
  HS_CUDA(cudaStreamCreateWithFlags(&B->up_stream, cudaStreamNonBlocking));
  *out = B;
  return 0;
}

int hslam_ba_batch_destroy(hslam_ba_batch* B) {
  if (!B) return 0;
  cudaSetDevice(B->ctx->device);
  cudaStreamSynchronize(B->ctx->stream);
  B->d_descs.release(); B->d_prefix.release(); B->d_up.release(); B->d_dn.release(); B->d_io.release();
  cudaFreeHost(B->h_descs); cudaFreeHost(B->h_prefix);
  if (B->h_up) cudaFreeHost(B->h_up);
  if (B->h_dn) cudaFreeHost(B->h_dn);
  cudaEventDestroy(B->copied);
  for (int g = 0; g < 8; ++g) if (B->ev_grp[g]) cudaEventDestroy(B->ev_grp[g]);
  for (int g = 0; g < 2; ++g) if (B->ev_th[g]) cudaEventDestroy(B->ev_th[g]);
  for (int g = 0; g < 8; ++g) if (B->ev_up[g]) cudaEventDestroy(B->ev_up[g]);
  if (B->ev_begin) cudaEventDestroy(B->ev_begin);
  if (B->up_stream) { cudaStreamSynchronize(B->up_stream); cudaStreamDestroy(B->up_stream); }
  if (B->graph_iter) cudaGraphExecDestroy(B->graph_iter);
  if (B->graph_acc) cudaGraphExecDestroy(B->graph_acc);
  delete B;
  return 0;
}

// One GN iteration's hot path for EVERY window of the batch: five launches in total.
// Windows [g0, g1) of the batch (the whole batch by default).
static int batch_enqueue_linearize(hslam_ba_batch* B, int g0 = 0, int g1 = -1) {
  hslam_ctx* ctx = B->ctx;
  const int n = (int)B->items.size();
  if (g1 < 0) g1 = n;
  const int b0 = B->h_prefix[g0], cnt = B->h_prefix[g1] - b0;   // table 0 = linearise blocks
  if (cnt <= 0) return 0;
  HS_LAUNCH(ctx, HSLAM_K_BA_LINEARIZE, lin_launch(true, B->fullJ, cnt, ctx->stream, (const BADesc*)B->d_descs.p, (const int*)B->d_prefix.p,
                                                  B->uniform[0] ? -B->uniform[0] : n, b0, (const int*)nullptr, 0));
  HS_CUDA(cudaGetLastError());
  return 0;
}
// System::setNewFrameEnergyTH inside the device-resident iteration (FullSystemOptimize.cpp:124: part of linearizeAll):
// one block per window on the side stream, forked in front of the stitch and joined behind it — the accumulate kernels
// do not read the threshold, the next linearisation does.  (Forked right behind the lineariser it ran beside the top
// kernel, whose two 128-register blocks per SM then no longer fit on 64 of the SMs: +23 us per step; the stitch's small
// blocks do not care.)
static int batch_fork_th(hslam_ba_batch* B) {
  hslam_ctx* ctx = B->ctx;
  HS_CUDA(cudaEventRecord(B->ev_th[0], ctx->stream));
  HS_CUDA(cudaStreamWaitEvent(ctx->side, B->ev_th[0], 0));
  ctx->launches++;   // side stream: counted, not event-timed
  ba_energy_th_kernel<<<(unsigned)B->items.size(), TH_THREADS, 0, ctx->side>>>(B->d_descs.p, B->th_newest, nullptr, nullptr, nullptr);
  HS_CUDA(cudaGetLastError());
  HS_CUDA(cudaEventRecord(B->ev_th[1], ctx->side));
  return 0;
}
static int batch_join_th(hslam_ba_batch* B) { HS_CUDA(cudaStreamWaitEvent(B->ctx->stream, B->ev_th[1], 0)); return 0; }

static int batch_enqueue_top(hslam_ba_batch* B) {
  hslam_ctx* ctx = B->ctx;
  const int n = (int)B->items.size();
  const int* pf = B->d_prefix.p;
  TRY(schur_attr(ctx, B->smem));
  TRY(top_attr(ctx));
  // modes A and L touch disjoint residuals, accumulators and records: outside the per-kernel event profiling they run
  // side by side (L on the side stream), so the two ~2-wave launches share ONE wave tail
  static const bool serial_top = getenv("HSLAM_TOP_SERIAL") != nullptr;
  const bool has_L = B->any_L && B->blocks[6] > 0;
  const bool par = has_L && B->blocks[1] > 0 && !ctx->profile && !serial_top;
  if (par) {
    HS_CUDA(cudaEventRecord(B->ev_th[0], ctx->stream));
    HS_CUDA(cudaStreamWaitEvent(ctx->side, B->ev_th[0], 0));
  }
  if (B->blocks[1] > 0) {
    HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<0><<<B->blocks[1], 2 * TOP_THREADS, kTopSmem, ctx->stream>>>(B->d_descs.p, pf + (n + 1), B->uniform[1] ? -B->uniform[1] : n)));
  }
  if (has_L) {
    cudaStream_t ls = par ? ctx->side : ctx->stream;
    if (par) {
      ctx->launches++;
      ba_top_kernel<1><<<B->blocks[6], 2 * TOP_THREADS, kTopSmem, ls>>>(B->d_descs.p, pf + 6 * (n + 1), B->uniform[6] ? -B->uniform[6] : n);
      HS_CUDA(cudaEventRecord(B->ev_th[1], ctx->side));
      HS_CUDA(cudaStreamWaitEvent(ctx->stream, B->ev_th[1], 0));
    } else {
      HS_LAUNCH(ctx, HSLAM_K_BA_TOP, (ba_top_kernel<1><<<B->blocks[6], 2 * TOP_THREADS, kTopSmem, ls>>>(B->d_descs.p, pf + 6 * (n + 1), B->uniform[6] ? -B->uniform[6] : n)));
    }
  }
  HS_CUDA(cudaGetLastError());
  return 0;
}
static int batch_enqueue_accumulate(hslam_ba_batch* B, bool with_th = false) {
  hslam_ctx* ctx = B->ctx;
  const int n = (int)B->items.size();
  const int* pf = B->d_prefix.p;
  TRY(batch_enqueue_top(B));
  if (B->blocks[2] > 0)
    HS_LAUNCH(ctx, HSLAM_K_BA_SCHUR, schur_launch(B->maxt, B->blocks[2], B->smem, ctx->stream, B->d_descs.p, pf + 2 * (n + 1), B->uniform[2] ? -B->uniform[2] : n));
  HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_SC, (ba_fold_sc_kernel<<<B->blocks[4], 256, 0, ctx->stream>>>(B->d_descs.p, pf + 4 * (n + 1), B->uniform[4] ? -B->uniform[4] : n, 3, 1)));
  if (with_th) TRY(batch_fork_th(B));
  HS_LAUNCH(ctx, HSLAM_K_BA_STITCH, (ba_stitch_kernel<4><<<B->blocks[3], 128, 0, ctx->stream>>>(B->d_descs.p, pf + 3 * (n + 1), B->uniform[3] ? -B->uniform[3] : n, 0, 3, 1)));
  HS_CUDA(cudaGetLastError());
  if (with_th) TRY(batch_join_th(B));
  return 0;
}

// The launch chain as a CUDA graph: one graph launch per iteration instead of 4-6 kernel launches (shorter gaps between
// the kernels on the device, one driver call on the host).  Bypassed while per-kernel event profiling is on.
static int batch_graph_launch(hslam_ba_batch* B, bool with_linearize) {
  hslam_ctx* ctx = B->ctx;
  cudaGraphExec_t& exec = with_linearize ? B->graph_iter : B->graph_acc;
  uint64_t& ver = with_linearize ? B->graph_iter_version : B->graph_acc_version;
  int& nk = with_linearize ? B->graph_iter_kernels : B->graph_acc_kernels;
  static const bool disabled = getenv("HSLAM_NO_GRAPH") != nullptr;
  const bool with_th = with_linearize && B->th_newest >= 0;
  if (ctx->profile || disabled) {
    if (with_linearize) TRY(batch_enqueue_linearize(B));
    return batch_enqueue_accumulate(B, with_th);
  }
  if (!exec || ver != B->refresh_version) {
    TRY(schur_attr(ctx, B->smem));
    TRY(top_attr(ctx));
    if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
    const int64_t l0 = ctx->launches;
    HS_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (with_linearize) rc = batch_enqueue_linearize(B);
    if (!rc) rc = batch_enqueue_accumulate(B, with_th);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    HS_CUDA(e);
    nk = (int)(ctx->launches - l0);
    ctx->launches = l0;
    HS_CUDA(cudaGraphInstantiate(&exec, g, 0));
    cudaGraphDestroy(g);
    ver = B->refresh_version;
  }
  HS_CUDA(cudaGraphLaunch(exec, ctx->stream));
  ctx->launches += nk;
  return 0;
}

int hslam_ba_batch_set_newest_frame(hslam_ba_batch* B, int newest_frame) {
  HS_ARG(B, "hslam_ba_batch_set_newest_frame: null");
  for (hslam_ba* ba : B->items) HS_ARG(newest_frame < ba->n, "hslam_ba_batch_set_newest_frame: frame index out of range for a window");
  if (newest_frame < 0) newest_frame = -1;
  if (newest_frame != B->th_newest) { B->th_newest = newest_frame; B->refresh_version++; }   // the graph is re-captured
  return 0;
}

int hslam_ba_batch_iteration(hslam_ba_batch* B) {
  HS_ARG(B, "hslam_ba_batch_iteration: null");
  HS_CUDA(cudaSetDevice(B->ctx->device));
  TRY(batch_refresh(B));
  return batch_graph_launch(B, true);
}

// ---- host-buffer API of a batch: one staging block up, one down per call ------------------------------
static int pinned_ensure(char** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = bytes + bytes / 4 + 4096;
  HS_CUDA(cudaMallocHost((void**)p, *cap));
  return 0;
}
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// Scalar part of hslam_ba_set_window (the tables travel through the staging block).
static int window_scalars(hslam_ba* ba, const hslam_window* w) {
  HS_ARG(w->n_frames == ba->n, "hslam_ba_batch_step_host: n_frames differs from the window's topology");
  HS_ARG(w->frames && w->frameEnergyTH && w->precalc && w->adHost && w->adTarget && w->adHTdeltaF && w->framePrior &&
             w->frameDeltaPrior, "hslam_ba_batch_step_host: null table");
  hslam_ctx* ctx = ba->ctx;
  BAWin W = ba->hdesc.W;
  W.w = w->w; W.h = w->h;
  W.fx = w->fx; W.fy = w->fy; W.cx = w->cx; W.cy = w->cy; W.fxi = w->fxi; W.fyi = w->fyi;
  for (int i = 0; i < 4; ++i) W.cDeltaF[i] = w->cDeltaF[i];
  for (int f = 0; f < ba->n; ++f) {
    HS_ARG(w->frames[f] && w->frames[f]->ctx == ctx, "hslam_ba_batch_step_host: frame missing or from another context");
    HS_ARG(w->frames[f]->w == w->w && w->frames[f]->h == w->h, "hslam_ba_batch_step_host: frame size mismatch");
    TRY(hs_frame_ensure_pair(const_cast<hslam_frame*>(w->frames[f])));
    W.img[f] = w->frames[f]->dI[0];
    W.pimg[f] = w->frames[f]->pair0;
    W.frameEnergyTH[f] = w->frameEnergyTH[f];
  }
  StitchArgs S = ba->hdesc.S;
  for (int i = 0; i < 4; ++i) { S.cPrior[i] = w->cPrior[i]; S.cDeltaF[i] = w->cDeltaF[i]; }
  // host copies of the frame priors: calcLEnergy's prior term reads them (the tables themselves travel in the staging block)
  ba->h_fprior.assign(w->framePrior, w->framePrior + (size_t)ba->n * 8);
  ba->h_fdprior.assign(w->frameDeltaPrior, w->frameDeltaPrior + (size_t)ba->n * 8);
  if (memcmp(&W, &ba->hdesc.W, sizeof(BAWin)) != 0 || memcmp(&S, &ba->hdesc.S, sizeof(StitchArgs)) != 0) {
    ba->hdesc.W = W; ba->hdesc.S = S;
    ba->desc_version++;
  }
  return 0;
}

// Fixed layout of the D2H block: every output of every window has its slot, whether or not a call asks for it
// (hslam_ba_batch_io_bind hands these slots out as the caller's buffers).
static int batch_layout_down(hslam_ba_batch* B) {
  const int n = (int)B->items.size();
  size_t dn = 0;
  auto take = [&](size_t bytes) { long long off = (long long)dn; dn = al256(dn + bytes); return off; };
  // window-major within each kind: the states / the energies of a GROUP of windows are one contiguous span each
  for (int i = 0; i < n; ++i) B->io[i].dn_newState = take(B->items[i]->R);
  B->dn_state_end = dn;
  for (int i = 0; i < n; ++i) {
    const size_t R = B->items[i]->R;
    B->io[i].dn_newEnergy = take(R * 4); B->io[i].dn_newEnergyWO = take(R * 4);
  }
  B->dn_res_end = dn;
  for (int i = 0; i < n; ++i) {
    const size_t N = B->items[i]->N;
    B->io[i].dn_energy = take(8); B->io[i].dn_nres = take(8); B->io[i].dn_th = take(4); B->io[i].dn_Hb = take(3 * (N * N + N) * 8);
    B->io[i].newest = 0; B->io[i].pad = 0;
  }
  B->dn_sys_end = dn;
  for (int i = 0; i < n; ++i) B->io[i].dn_step = take((size_t)B->items[i]->P * 4);
  B->dn_end = dn;
  TRY(B->d_dn.ensure(std::max(dn, (size_t)256)));
  TRY(pinned_ensure(&B->h_dn, &B->h_dn_cap, std::max(dn, (size_t)256)));
  return 0;
}

// Fixed layout of the H2D block: [BAIO table | x of every window | per window: tables, idepth, deltaF | idepth_zero of
// every window].  Window-major chunks: the inputs of a GROUP of windows are one contiguous span, so step_host can
// upload group g+1 while group g is being linearised.
static int batch_layout_up(hslam_ba_batch* B) {
  const int n = (int)B->items.size();
  size_t up = al256(n * sizeof(BAIO));
  auto take = [&](size_t bytes) { long long off = (long long)up; up = al256(up + bytes); return off; };
  for (int i = 0; i < n; ++i) B->io[i].up_x = take((size_t)B->items[i]->N * 8);   // solution vectors (resubstitute)
  B->up_tab_begin = up;
  B->up_win_begin.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    hslam_ba* ba = B->items[i];
    const size_t nn = (size_t)ba->n * ba->n;
    BAIO& L = B->io[i];
    B->up_win_begin[i] = up;
    L.up_precalc = take(nn * HSLAM_PRECALC_STRIDE * 4); L.up_adHTd = take(nn * 8 * 4);
    L.up_adH = take(nn * 64 * 8); L.up_adT = take(nn * 64 * 8);
    L.up_fprior = take((size_t)ba->n * 8 * 8); L.up_fdprior = take((size_t)ba->n * 8 * 8);
    L.up_idepth = take((size_t)ba->P * 4);
  }
  // deltaF and idepth_zero have regions of their own: a caller that lets the device form deltaF (derive_deltaF) or leaves
  // idepth_zero unchanged does not pay for their bytes
  B->up_win_begin[n] = B->up_dF_begin = up;
  for (int i = 0; i < n; ++i) B->io[i].up_deltaF = take((size_t)B->items[i]->P * 4);
  B->up_idz_begin = up;
  for (int i = 0; i < n; ++i) B->io[i].up_idz = take((size_t)B->items[i]->P * 4);
  B->up_end = up;
  TRY(pinned_ensure(&B->h_up, &B->h_up_cap, up));
  TRY(B->d_up.ensure(up));
  return 0;
}

int hslam_ba_batch_io_bind(hslam_ba_batch* B, hslam_ba_step_io* io, float** step_out) {
  HS_ARG(B && io, "hslam_ba_batch_io_bind: null argument");
  HS_CUDA(cudaSetDevice(B->ctx->device));
  HS_CUDA(cudaStreamSynchronize(B->ctx->stream));
  TRY(batch_layout_down(B));
  TRY(batch_layout_up(B));
  const int n = (int)B->items.size();
  for (int i = 0; i < n; ++i) {
    const BAIO& L = B->io[i];
    // per-point INPUT slots: the caller writes idepth / deltaF / idepth_zero in place (no staging copy)
    io[i].idepth = (const float*)(B->h_up + L.up_idepth); io[i].deltaF = (const float*)(B->h_up + L.up_deltaF);
    io[i].idepth_zero = nullptr; (void)L.up_idz;
    io[i].energy = (double*)(B->h_dn + L.dn_energy); io[i].nres2 = (int32_t*)(B->h_dn + L.dn_nres);
    io[i].Hb = (double*)(B->h_dn + L.dn_Hb);
    io[i].newEnergy = (float*)(B->h_dn + L.dn_newEnergy); io[i].newEnergyWithOutlier = (float*)(B->h_dn + L.dn_newEnergyWO);
    io[i].newState = (int8_t*)(B->h_dn + L.dn_newState);
    io[i].frameEnergyTH_out = (float*)(B->h_dn + L.dn_th);
    if (step_out) step_out[i] = (float*)(B->h_dn + L.dn_step);
  }
  return 0;
}

int hslam_ba_batch_window_bind(hslam_ba_batch* B, int i, hslam_window* w) {
  HS_ARG(B && w && i >= 0 && i < (int)B->items.size(), "hslam_ba_batch_window_bind: bad argument");
  HS_CUDA(cudaSetDevice(B->ctx->device));
  HS_CUDA(cudaStreamSynchronize(B->ctx->stream));
  TRY(batch_layout_down(B));
  TRY(batch_layout_up(B));
  const BAIO& L = B->io[i];
  w->precalc = (const float*)(B->h_up + L.up_precalc); w->adHTdeltaF = (const float*)(B->h_up + L.up_adHTd);
  w->adHost = (const double*)(B->h_up + L.up_adH); w->adTarget = (const double*)(B->h_up + L.up_adT);
  w->framePrior = (const double*)(B->h_up + L.up_fprior); w->frameDeltaPrior = (const double*)(B->h_up + L.up_fdprior);
  return 0;
}

int hslam_ba_batch_step_host(hslam_ba_batch* B, const hslam_ba_step_io* io) {
  HS_ARG(B && io, "hslam_ba_batch_step_host: null argument");
  hslam_ctx* ctx = B->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const int n = (int)B->items.size();
  // ---- layout of the two staging blocks (both fixed: see batch_layout_up / batch_layout_down)
  TRY(batch_layout_down(B));
  TRY(batch_layout_up(B));
  std::vector<BAIO> eff(B->io);   // offsets of what is absent in this call become -1 for the import kernel
  // pass 1: scalars, argument checks and the offset table of every window (no bulk data yet)
  for (int i = 0; i < n; ++i) {
    hslam_ba* ba = B->items[i];
    HS_ARG(ba->have_window && ba->have_topology, "hslam_ba_batch_step_host: a window has no window/topology set");
    const hslam_ba_step_io& o = io[i];
    BAIO& E = eff[i];
    if (o.window) TRY(window_scalars(ba, o.window));
    else E.up_precalc = E.up_adHTd = E.up_adH = E.up_adT = E.up_fprior = E.up_fdprior = -1;
    if (!o.idepth) E.up_idepth = -1;
    if (!o.deltaF) E.up_deltaF = -1;
    E.pad = (o.derive_deltaF && o.idepth && !o.deltaF) ? 1 : 0;   // the import kernel forms deltaF = idepth - idepth_zero
    if (!o.idepth_zero) E.up_idz = -1;
    E.up_x = -1;
    // outputs the caller did not ask for are neither exported nor copied
    if (!o.newEnergy) E.dn_newEnergy = -1;
    if (!o.newEnergyWithOutlier) E.dn_newEnergyWO = -1;
    if (!o.newState) E.dn_newState = -1;
    if (o.frameEnergyTH_out) {
      HS_ARG(o.newest_frame >= 0 && o.newest_frame < ba->n, "hslam_ba_batch_step_host: newest_frame out of range");
      E.newest = o.newest_frame;
    } else {
      E.dn_th = -1;
    }
  }
  memcpy(B->h_up, eff.data(), n * sizeof(BAIO));   // the offset table leads the block
  TRY(batch_refresh(B));   // descriptors (only if a scalar changed) and block-prefix tables
  const BAIO* d_io = reinterpret_cast<const BAIO*>(B->d_up.p);
  bool any_res = false, any_sys = false, any_th = false;
  for (int i = 0; i < n; ++i) {
    any_res |= io[i].newEnergy || io[i].newEnergyWithOutlier || io[i].newState;
    any_th |= io[i].frameEnergyTH_out != nullptr;
    any_sys |= io[i].Hb || io[i].nres2 || io[i].energy || io[i].frameEnergyTH_out;
  }
  // The batch is processed in up to 8 window groups.  Upload stream: pack + H2D + import of group g+1 while the
  // context stream linearises group g; side stream (high priority): export + D2H of the per-residual results of
  // group g while group g+1 is linearised, the last group's under the accumulate kernels.  Per-window results do
  // not depend on the grouping.
  int G = n >= 8 ? 4 : 1;
  if (const char* e = getenv("HSLAM_STEP_GROUPS")) G = std::max(1, std::min(8, std::min(n, atoi(e))));
  HS_CUDA(cudaEventRecord(B->ev_begin, ctx->stream));          // uploads overwrite state the stream's earlier work may read
  HS_CUDA(cudaStreamWaitEvent(B->up_stream, B->ev_begin, 0));
  HS_CUDA(cudaMemcpyAsync(B->d_up.p, B->h_up, al256(n * sizeof(BAIO)), cudaMemcpyHostToDevice, B->up_stream));
  // ---- results.  Three kinds of caller buffer: a slot of the batch's own pinned block (hslam_ba_batch_io_bind:
  // nothing to do, the two bulk copies below fill it), other pinned memory (hslam_host_alloc: direct DMA per
  // array), pageable memory (memcpy from the pinned block after the synchronisation).  The per-residual arrays
  // are final after the linearise kernel: they are exported and copied on the side stream WHILE the accumulate
  // kernels run.
  struct Out { void* dst; long long off; size_t bytes; };
  std::vector<Out> staged;
  bool bulk_res = false, bulk_sys = false;
  auto emit = [&](cudaStream_t st, bool& bulk, void* dst, long long off, size_t bytes) -> int {
    if (!dst || bytes == 0) return 0;
    if (dst == (void*)(B->h_dn + off)) { bulk = true; return 0; }
    if (ctx->is_pinned(dst, bytes)) {
      HS_CUDA(cudaMemcpyAsync(dst, B->d_dn.p + off, bytes, cudaMemcpyDeviceToHost, st));
    } else {
      bulk = true;
      staged.push_back({dst, off, bytes});
    }
    return 0;
  };
  for (int g = 0; g < G; ++g) {
    const int g0 = (int)((long long)n * g / G), g1 = (int)((long long)n * (g + 1) / G);
    // ---- inputs of the group: pack what is not already in its pinned slot, one H2D of the group's span, import
    bool any_in = false, any_idz = false, any_dF = false;
    int gP = 1, gR = 1;
    for (int i = g0; i < g1; ++i) {
      hslam_ba* ba = B->items[i];
      const hslam_ba_step_io& o = io[i];
      const BAIO& L = B->io[i];
      const size_t nn = (size_t)ba->n * ba->n, P = ba->P;
      if (o.window) {
        const hslam_window* w = o.window;
        // a table that already lives in its slot of the pinned upload block (hslam_ba_batch_window_bind) needs no copy
        auto tab = [&](const void* src, long long off, size_t bytes) { if (src != (const void*)(B->h_up + off)) memcpy(B->h_up + off, src, bytes); };
        tab(w->precalc, L.up_precalc, nn * HSLAM_PRECALC_STRIDE * 4);
        tab(w->adHTdeltaF, L.up_adHTd, nn * 8 * 4);
        tab(w->adHost, L.up_adH, nn * 64 * 8);
        tab(w->adTarget, L.up_adT, nn * 64 * 8);
        tab(w->framePrior, L.up_fprior, (size_t)ba->n * 8 * 8);
        tab(w->frameDeltaPrior, L.up_fdprior, (size_t)ba->n * 8 * 8);
        any_in = true;
      }
      // a per-point input that already lives in its slot of the pinned block (hslam_ba_batch_io_bind) needs no copy
      auto put = [&](const float* src, long long off, bool& flag) {
        if (!src) return;
        if ((const char*)src != B->h_up + off) memcpy(B->h_up + off, src, P * 4);
        flag = true;
      };
      put(o.idepth, L.up_idepth, any_in);
      put(o.deltaF, L.up_deltaF, any_dF);
      put(o.idepth_zero, L.up_idz, any_idz);
      gP = std::max(gP, ba->P); gR = std::max(gR, ba->R);
    }
    if (any_in) {
      const size_t lo = B->up_win_begin[g0], hi = B->up_win_begin[g1];
      HS_CUDA(cudaMemcpyAsync(B->d_up.p + lo, B->h_up + lo, hi - lo, cudaMemcpyHostToDevice, B->up_stream));
    }
    if (any_dF) {
      const size_t lo = (size_t)B->io[g0].up_deltaF, hi = g1 < n ? (size_t)B->io[g1].up_deltaF : B->up_idz_begin;
      HS_CUDA(cudaMemcpyAsync(B->d_up.p + lo, B->h_up + lo, hi - lo, cudaMemcpyHostToDevice, B->up_stream));
    }
    if (any_idz) {
      const size_t lo = (size_t)B->io[g0].up_idz, hi = g1 < n ? (size_t)B->io[g1].up_idz : B->up_end;
      HS_CUDA(cudaMemcpyAsync(B->d_up.p + lo, B->h_up + lo, hi - lo, cudaMemcpyHostToDevice, B->up_stream));
    }
    if (any_in || any_idz || any_dF) {
      dim3 grid((unsigned)std::max(1, std::min(64, (gP + 255) / 256)), (unsigned)(g1 - g0));
      ctx->launches++;   // upload stream: counted, not event-timed
      ba_batch_import_kernel<<<grid, 256, 0, B->up_stream>>>(B->d_descs.p + g0, d_io + g0, B->d_up.p);
      HS_CUDA(cudaGetLastError());
    }
    HS_CUDA(cudaEventRecord(B->ev_up[g], B->up_stream));
    HS_CUDA(cudaStreamWaitEvent(ctx->stream, B->ev_up[g], 0));
    // ---- linearise the group
    TRY(batch_enqueue_linearize(B, g0, g1));
    if (!any_res && !any_th) continue;
    HS_CUDA(cudaEventRecord(B->ev_grp[g], ctx->stream));
    HS_CUDA(cudaStreamWaitEvent(ctx->side, B->ev_grp[g], 0));
    if (any_th) {
      // System::setNewFrameEnergyTH of the group's windows (one block each), on the side stream under the next group's
      // linearisation: writes the threshold into the window's descriptor and into its slot of the D2H block
      ctx->launches++;
      ba_energy_th_kernel<<<g1 - g0, TH_THREADS, 0, ctx->side>>>(B->d_descs.p + g0, -1, nullptr, d_io + g0, B->d_dn.p);
      HS_CUDA(cudaGetLastError());
    }
    if (!any_res) continue;
    // ---- per-residual results of the group
    dim3 grid((unsigned)std::max(1, std::min(128, (gR + 255) / 256)), (unsigned)(g1 - g0));
    ctx->launches++;   // side stream: counted, not event-timed
    ba_batch_export_kernel<<<grid, 256, 0, ctx->side>>>(B->d_descs.p + g0, d_io + g0, B->d_dn.p, 1);
    HS_CUDA(cudaGetLastError());
    bool bulk_st = false, bulk_en = false;
    for (int i = g0; i < g1; ++i) {
      const size_t R = B->items[i]->R;
      TRY(emit(ctx->side, bulk_en, io[i].newEnergy, B->io[i].dn_newEnergy, R * 4));
      TRY(emit(ctx->side, bulk_en, io[i].newEnergyWithOutlier, B->io[i].dn_newEnergyWO, R * 4));
      TRY(emit(ctx->side, bulk_st, io[i].newState, B->io[i].dn_newState, R));
    }
    // the group's slices of the state region and of the energy region are contiguous (batch_layout_down)
    if (bulk_st) {
      const size_t lo = (size_t)B->io[g0].dn_newState, hi = g1 < n ? (size_t)B->io[g1].dn_newState : B->dn_state_end;
      HS_CUDA(cudaMemcpyAsync(B->h_dn + lo, B->d_dn.p + lo, hi - lo, cudaMemcpyDeviceToHost, ctx->side));
    }
    if (bulk_en) {
      const size_t lo = (size_t)B->io[g0].dn_newEnergy, hi = g1 < n ? (size_t)B->io[g1].dn_newEnergy : B->dn_res_end;
      HS_CUDA(cudaMemcpyAsync(B->h_dn + lo, B->d_dn.p + lo, hi - lo, cudaMemcpyDeviceToHost, ctx->side));
    }
    bulk_res |= bulk_st || bulk_en;
  }
  if (any_res || any_th) HS_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
  // ---- accumulate.  SPLIT form (a batch of same-shape windows whose system outputs are the bound slots of the batch's
  // result block): the A and L systems, the energy and the counts are final once the top accumulators are folded — they
  // are stitched, exported and copied to the host on a second stream WHILE the Schur fold runs; only the Schur-complement
  // system (a third of the bytes) is copied behind the last kernel.
  bool split = any_sys && n >= 8 && !ctx->profile && B->uniform[2] && B->uniform[3] && B->uniform[4] && getenv("HSLAM_NO_SPLIT") == nullptr;
  for (int i = 0; i < n && split; ++i) {
    const hslam_ba* ba = B->items[i];
    split = !io[i].newEnergy && !io[i].newEnergyWithOutlier && ba->n == B->items[0]->n && ba->LD == B->items[0]->LD && io[i].Hb == (double*)(B->h_dn + B->io[i].dn_Hb) &&
            (!io[i].energy || io[i].energy == (double*)(B->h_dn + B->io[i].dn_energy)) &&
            (!io[i].nres2 || (void*)io[i].nres2 == (void*)(B->h_dn + B->io[i].dn_nres)) &&
            (!io[i].frameEnergyTH_out || (void*)io[i].frameEnergyTH_out == (void*)(B->h_dn + B->io[i].dn_th)) &&
            (i == 0 || B->io[i].dn_energy - B->io[i - 1].dn_energy == B->io[1].dn_energy - B->io[0].dn_energy);
  }
  if (split) {
    const int nf = B->items[0]->n, LD = B->items[0]->LD;
    const size_t N = B->items[0]->N, per = (N * N + N) * 8, pitch = (size_t)(B->io[1].dn_energy - B->io[0].dn_energy);
    const int* pf = B->d_prefix.p;
    TRY(batch_enqueue_top(B));
    const int ftb = fold_top_blocks(nf, 3), fsb = nf * fold_sc_blocks(LD);
    HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_SC, (ba_fold_sc_kernel<<<n * ftb, 256, 0, ctx->stream>>>(B->d_descs.p, nullptr, -ftb, 3, 0)));
    // (the high-priority side stream: its blocks go ahead of the Schur kernel's, and the threshold kernels, whose floats
    // ride in this copy, precede in stream order.  Measured alternatives: the same on the normal-priority upload stream
    // 1.57 ms per step_host instead of 1.48 — the stitch then trickles in behind the Schur blocks; with the per-residual
    // energy arrays requested (43 MB on the side stream) the split form loses 0.3 ms, so it is not used then.)
    cudaStream_t bs = ctx->side;
    HS_CUDA(cudaEventRecord(B->ev_th[0], ctx->stream));
    HS_CUDA(cudaStreamWaitEvent(bs, B->ev_th[0], 0));
    ctx->launches += 2;   // counted, not event-timed
    ba_stitch_kernel<4><<<B->blocks[3], 128, 0, bs>>>(B->d_descs.p, pf + 3 * (n + 1), -B->uniform[3], 0, 2, 1);
    ba_batch_export_kernel<<<dim3(2, (unsigned)n), 256, 0, bs>>>(B->d_descs.p, d_io, B->d_dn.p, 4);
    HS_CUDA(cudaGetLastError());
    {   // per window [energy | counts | threshold | H_A b_A H_L b_L]: one strided copy
      const size_t lo = (size_t)B->io[0].dn_energy, width = (size_t)(B->io[0].dn_Hb - B->io[0].dn_energy) + 2 * per;
      HS_CUDA(cudaMemcpy2DAsync(B->h_dn + lo, pitch, B->d_dn.p + lo, pitch, width, n, cudaMemcpyDeviceToHost, bs));
    }
    HS_CUDA(cudaEventRecord(B->ev_th[1], bs));
    HS_LAUNCH(ctx, HSLAM_K_BA_SCHUR, schur_launch(B->maxt, B->blocks[2], B->smem, ctx->stream, B->d_descs.p, pf + 2 * (n + 1), -B->uniform[2]));
    HS_LAUNCH(ctx, HSLAM_K_BA_FOLD_SC, (ba_fold_sc_kernel<<<n * fsb, 256, 0, ctx->stream>>>(B->d_descs.p, nullptr, -fsb, 0, 1)));
    HS_LAUNCH(ctx, HSLAM_K_BA_STITCH, (ba_stitch_kernel<4><<<B->blocks[3], 128, 0, ctx->stream>>>(B->d_descs.p, pf + 3 * (n + 1), -B->uniform[3], 2, 3, 0)));
    HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_batch_export_kernel<<<dim3(2, (unsigned)n), 256, 0, ctx->stream>>>(B->d_descs.p, d_io, B->d_dn.p, 8)));
    HS_CUDA(cudaGetLastError());
    {
      const size_t lo = (size_t)B->io[0].dn_Hb + 2 * per;
      HS_CUDA(cudaMemcpy2DAsync(B->h_dn + lo, pitch, B->d_dn.p + lo, pitch, per, n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    HS_CUDA(cudaStreamWaitEvent(ctx->stream, B->ev_th[1], 0));
  } else {
  TRY(batch_graph_launch(B, false));
  if (any_th) HS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));   // the thresholds ride in the systems region
  }
  if (any_sys && !split) {
    HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_batch_export_kernel<<<dim3(2, (unsigned)n), 256, 0, ctx->stream>>>(B->d_descs.p, d_io, B->d_dn.p, 2)));
    HS_CUDA(cudaGetLastError());
    for (int i = 0; i < n; ++i) {
      const size_t N = B->items[i]->N;
      TRY(emit(ctx->stream, bulk_sys, io[i].energy, B->io[i].dn_energy, 8)); TRY(emit(ctx->stream, bulk_sys, io[i].nres2, B->io[i].dn_nres, 8));
      TRY(emit(ctx->stream, bulk_sys, io[i].Hb, B->io[i].dn_Hb, 3 * (N * N + N) * 8));
      if (io[i].frameEnergyTH_out) bulk_sys = true;   // always through the block: the host mirror below reads it there
    }
    if (bulk_sys)
      HS_CUDA(cudaMemcpyAsync(B->h_dn + B->dn_res_end, B->d_dn.p + B->dn_res_end, B->dn_sys_end - B->dn_res_end,
                              cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (any_res) HS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));   // the context stream stays the one to wait on
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (const Out& s : staged) memcpy(s.dst, B->h_dn + s.off, s.bytes);
  if (any_th)
    for (int i = 0; i < n; ++i) {
      if (!io[i].frameEnergyTH_out) continue;
      // the device descriptor of the batch already holds the new threshold: bring the host mirrors in line without
      // marking the batch dirty (a caller that hands the same value back in its next hslam_window changes nothing)
      hslam_ba* ba = B->items[i];
      const float th = *reinterpret_cast<const float*>(B->h_dn + B->io[i].dn_th);
      const bool in_sync = B->versions[i] == ba->desc_version;
      ba->hdesc.W.frameEnergyTH[io[i].newest_frame] = th;
      B->h_descs[i].W.frameEnergyTH[io[i].newest_frame] = th;
      ba->desc_version++;
      if (in_sync) B->versions[i] = ba->desc_version;
      if (io[i].frameEnergyTH_out != (float*)(B->h_dn + B->io[i].dn_th)) *io[i].frameEnergyTH_out = th;
    }
  return 0;
}

int hslam_ba_batch_resubstitute_host(hslam_ba_batch* B, const double* const* x, float* const* step_out) {
  HS_ARG(B && x && step_out, "hslam_ba_batch_resubstitute_host: null argument");
  hslam_ctx* ctx = B->ctx;
  HS_CUDA(cudaSetDevice(ctx->device));
  const int n = (int)B->items.size();
  TRY(batch_layout_down(B));
  TRY(batch_layout_up(B));
  std::vector<BAIO> eff(B->io);
  for (int i = 0; i < n; ++i) {
    hslam_ba* ba = B->items[i];
    HS_ARG(ba->have_window && ba->have_topology && x[i], "hslam_ba_batch_resubstitute_host: window not set or null x");
    BAIO& E = eff[i];
    E.up_precalc = E.up_adHTd = E.up_adH = E.up_adT = E.up_fprior = E.up_fdprior = E.up_idepth = E.up_deltaF = E.up_idz = -1;
    memcpy(B->h_up + E.up_x, x[i], (size_t)ba->N * 8);
  }
  memcpy(B->h_up, eff.data(), n * sizeof(BAIO));
  TRY(batch_refresh(B));
  HS_CUDA(cudaMemcpyAsync(B->d_up.p, B->h_up, B->up_tab_begin, cudaMemcpyHostToDevice, ctx->stream));   // offset table + x slots
  const BAIO* d_io = reinterpret_cast<const BAIO*>(B->d_up.p);
  HS_LAUNCH(ctx, HSLAM_K_BA_EXPORT, (ba_batch_import_kernel<<<dim3(1, (unsigned)n), 256, 0, ctx->stream>>>(B->d_descs.p, d_io, B->d_up.p)));
  HS_LAUNCH(ctx, HSLAM_K_BA_RESUB, (ba_xad_kernel<<<n, 256, 0, ctx->stream>>>(B->d_descs.p)));
  // window groups: the steps of group g travel on the side stream while group g+1 is back-substituted
  int G = n >= 8 ? 4 : 1;
  if (const char* e = getenv("HSLAM_STEP_GROUPS")) G = std::max(1, std::min(8, std::min(n, atoi(e))));
  const int* hp = B->h_prefix + 5 * (n + 1);
  std::vector<int> staged;
  for (int g = 0; g < G; ++g) {
    const int g0 = (int)((long long)n * g / G), g1 = (int)((long long)n * (g + 1) / G);
    const int b0 = hp[g0], cnt = hp[g1] - b0;
    if (cnt > 0)
      HS_LAUNCH(ctx, HSLAM_K_BA_RESUB, (ba_resub_kernel<<<cnt, 128, 0, ctx->stream>>>(
                                           B->d_descs.p, B->d_prefix.p + 5 * (n + 1), B->uniform[5] ? -B->uniform[5] : n, d_io, B->d_dn.p, b0)));
    HS_CUDA(cudaGetLastError());
    cudaStream_t cs = ctx->stream;
    if (G > 1) {
      HS_CUDA(cudaEventRecord(B->ev_grp[g], ctx->stream));
      HS_CUDA(cudaStreamWaitEvent(ctx->side, B->ev_grp[g], 0));
      cs = ctx->side;
    }
    bool bulk = false;
    for (int i = g0; i < g1; ++i) {
      if (!step_out[i]) continue;
      const size_t bytes = (size_t)B->items[i]->P * 4;
      if ((void*)step_out[i] == (void*)(B->h_dn + B->io[i].dn_step)) bulk = true;
      else if (ctx->is_pinned(step_out[i], bytes))
        HS_CUDA(cudaMemcpyAsync(step_out[i], B->d_dn.p + B->io[i].dn_step, bytes, cudaMemcpyDeviceToHost, cs));
      else { bulk = true; staged.push_back(i); }
    }
    if (bulk) {   // the group's step slots are contiguous (batch_layout_down)
      const size_t lo = (size_t)B->io[g0].dn_step, hi = g1 < n ? (size_t)B->io[g1].dn_step : B->dn_end;
      HS_CUDA(cudaMemcpyAsync(B->h_dn + lo, B->d_dn.p + lo, hi - lo, cudaMemcpyDeviceToHost, cs));
    }
  }
  if (G > 1) {
    HS_CUDA(cudaEventRecord(ctx->ev_join, ctx->side));
    HS_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  }
  HS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i : staged) memcpy(step_out[i], B->h_dn + B->io[i].dn_step, (size_t)B->items[i]->P * 4);
  return 0;
}

}  // extern "C"
