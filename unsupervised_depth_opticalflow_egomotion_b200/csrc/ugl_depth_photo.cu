// Fused reprojection-photometric term of the depth / geom modes, both directions and all pyramid levels in one launch:
//   reconstruction (model_geometry.py:80-103 -> inverse_warp2, structures/inverse_warp.py:263-303): backproject the
//   centre disparity with K_s^-1, transform + project with P = K_s [R|t], normalise, out-of-range -> 2, bilinear sample of
//   the area-resized source frame, reprojection valid mask;
//   compute_texture_mask (model_geometry.py:134-140);
//   mask fusion (model_depth.py:262-269: valid * texture; model_geometry.py:854-855: flow-branch mask * texture);
//   compute_photometric_loss (model_geometry.py:143-153) summed over directions and levels.
// Forward = per-(sample, level, chunk) partial sums + finalize; backward recomputes the per-pixel quantities (nothing
// per-pixel is saved) and writes d loss / d disparity densely and d loss / d P through the deterministic two-stage reduction.
#include "ugl_common.cuh"
#include "ugl_geometry.cuh"
#include "ugl_reduce.cuh"
#include "ugl_flow_grad.cuh"   // accumulator layout + closing scales of the depth-mode single-pass kernel

#ifndef UGL_DP_FWD_MINB
#define UGL_DP_FWD_MINB 3
#endif
#ifndef UGL_DP_BWD_MINB
#define UGL_DP_BWD_MINB 2
#endif

namespace ugl {

struct DepthPhotoLevel {
  int h, w;
  const float* img;            // (B,3,h,w) centre frame, bilinear pyramid
  const float* src_area[2];    // (B,3,h,w) source frame (0: left / bwd pose, 1: right / fwd pose), area pyramid
  const float* src_bil[2];     // (B,3,h,w) source frame, bilinear pyramid (texture mask)
  const float* disp;           // (B,1,h,w)
  const float* Kinv;           // (B,3,3)
  const float* P[2];           // (B,3,4)
  const float* ext_mask[2];    // (B,1,h,w) or null: mask from the flow branch (geom mode) instead of the reprojection valid
  const unsigned char* ext_bytes;  // (B,h,w) or null: packed masks of ugl_geom_flow_forward_grad; base = all ext_need[dir] bits set
  unsigned ext_need[2];
  float* valid_out[2];         // optional (B,1,h,w)
  float* tex_out[2];           // optional (B,1,h,w)
  float* grad_disp;            // (B,1,h,w)   [backward]
};

struct DepthPhotoParams {
  int B, scales, chunks;
  DepthPhotoLevel lv[kMaxLevels];
  float* partials;             // fwd: [B][scales][chunks][4]; bwd: [B][scales][chunks][24]
  float* den;                  // [B][scales][2]  mean(mask) + 1e-12, kept for backward
  float* loss;                 // (B,)
  const float* gloss;          // (B,)            [backward]
  float* grad_P[2][kMaxLevels];// (B,3,4)         [backward]
  // single-pass mode (ugl_depth_photo_forward_grad / _combine)
  float* basis[kMaxLevels];    // (B,2,h,w) un-normalised d loss / d disparity through direction 0 / 1
  float* psum;                 // [B][scales][2][12] un-normalised d loss / d P sums
};

struct DepthPixel {
  float I[3], rec[3];
  float mask;                  // fused {0,1} mask of this direction
  float tex;                   // texture mask
  Projected pr;
  NormCoord nc;
  Tap tap;
};

// the coalesced loads of one pixel and direction: issued one loop iteration ahead of their use (the kernels are bound by the latency
// of their loads, not by bytes)
struct DepthDirect {
  float disp, I[3], bil[3], base;
};

__device__ __forceinline__ DepthDirect depth_direct(const DepthPhotoLevel& L, int b, int dir, long p) {
  const long plane = (long)L.h * L.w;
  DepthDirect d;
  d.disp = L.disp[(long)b * plane + p];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long o3 = ((long)b * 3 + c) * plane;
    d.I[c] = L.img[o3 + p];
    d.bil[c] = L.src_bil[dir][o3 + p];
  }
  d.base = -1.f;                                 // < 0: the reprojection valid mask
  if (L.ext_bytes) d.base = ((unsigned)L.ext_bytes[(long)b * plane + p] & L.ext_need[dir]) == L.ext_need[dir] ? 1.f : 0.f;
  else if (L.ext_mask[dir]) d.base = L.ext_mask[dir][(long)b * plane + p];
  return d;
}

// The pixels of a (sample, level, direction) are split into gridDim.x contiguous chunks (a multiple of the CTA size each); a CTA walks
// its chunk front to back, so consecutive iterations gather from the same rows of the source frame (L1 reuse).
struct ChunkRange { long begin, end; };
__device__ __forceinline__ ChunkRange chunk_range(long plane) {
  const long per = (((plane + gridDim.x - 1) / gridDim.x) + kRedThreads - 1) / kRedThreads * kRedThreads;
  ChunkRange r;
  r.begin = blockIdx.x * per;
  r.end = r.begin + per < plane ? r.begin + per : plane;
  return r;
}

// everything the forward and the backward need for one pixel and one direction
template <bool kGrad>
__device__ __forceinline__ void depth_pixel(const DepthPhotoLevel& L, const float* sK, const float* sP, int b, int dir, int i, int j,
                                            long p, const WarpGeom& g, const DepthDirect& dl, DepthPixel& o, float* gix, float* giy, float k_scale) {
  const long plane = (long)L.h * L.w;
  o.pr = project_pixel(sK, sP, dl.disp, j, i);
  o.nc = normalise(o.pr, g);
  o.tap = make_tap(unnormalize(o.nc.gx, L.w), unnormalize(o.nc.gy, L.h), L.w, L.h);
  const float valid = (fabsf(o.nc.gx) <= 1.0f && fabsf(o.nc.gy) <= 1.0f) ? 1.f : 0.f;
  float a = 0.f, s = 0.f;
  Corners cs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const long o3 = ((long)b * 3 + c) * plane;
    o.I[c] = dl.I[c];
    cs[c] = tap_fetch(L.src_area[dir] + o3, L.w, o.tap);
    o.rec[c] = corners_value(cs[c], o.tap);
    a = add_rn(a, fabsf(sub_rn(o.I[c], o.rec[c])));
    s = add_rn(s, fabsf(sub_rn(o.I[c], dl.bil[c])));
  }
  const float r3 = 1.0f / 3.0f;
  const float tex = div_c(a, 3.0f, r3) < div_c(s, 3.0f, r3) ? 1.f : 0.f;
  const float base = dl.base < 0.f ? valid : dl.base;
  o.mask = mul_rn(base, tex);
  o.tex = tex;
  if (!kGrad) {
    if (L.valid_out[dir]) L.valid_out[dir][(long)b * plane + p] = valid;
    if (L.tex_out[dir]) L.tex_out[dir][(long)b * plane + p] = tex;
  } else {
    float gx = 0.f, gy = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gr = k_scale * o.mask * sgnf(o.rec[c] - o.I[c]);      // d loss / d rec_c
      gx += gr * corners_ddx(cs[c], o.tap);
      gy += gr * corners_ddy(cs[c], o.tap);
    }
    *gix = gx; *giy = gy;
  }
}

// grid (chunks, B, 2 * levels): the two directions of a level run in separate CTAs (half the live state per thread, twice the
// CTAs in flight: the kernel is bound by the latency of its gathers, not by bytes)
__global__ void __launch_bounds__(kRedThreads, UGL_DP_FWD_MINB) depth_photo_fwd_kernel(const __grid_constant__ DepthPhotoParams p) {
  __shared__ float sK[9], sP[12];
  __shared__ float red[(kRedThreads / 32) * 2];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const DepthPhotoLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const WarpGeom g = make_warp_geom(L.w, L.h);
  const long plane = (long)L.h * L.w;
  float acc[2] = {0.f, 0.f};
  const ChunkRange cr = chunk_range(plane);
  long px = cr.begin + threadIdx.x;
  DepthDirect nxt = {};
  if (px < cr.end) nxt = depth_direct(L, b, dir, px);
#pragma unroll 1
  for (; px < cr.end; px += kRedThreads) {
    const DepthDirect dl = nxt;
    if (px + kRedThreads < cr.end) nxt = depth_direct(L, b, dir, px + kRedThreads);
    const int i = (int)(px / L.w), j = (int)(px % L.w);
    DepthPixel o;
    depth_pixel<false>(L, sK, sP, b, dir, i, j, px, g, dl, o, nullptr, nullptr, 0.f);
    acc[0] += (fabsf(o.I[0] - o.rec[0]) + fabsf(o.I[1] - o.rec[1]) + fabsf(o.I[2] - o.rec[2])) * o.mask;
    acc[1] += o.mask;
  }
  const float v = block_reduce_n<kRedThreads, 2>(acc, red);
  if (threadIdx.x < 2) p.partials[(((long)b * p.scales + l) * p.chunks + blockIdx.x) * 4 + 2 * dir + threadIdx.x] = v;
}

// one warp per sample: per level fixed-order fp64 sum of the chunk partials, closing formula, sum over levels
__global__ void depth_photo_finalize_kernel(const __grid_constant__ DepthPhotoParams p) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= p.B) return;
  float total = 0.f;
  for (int l = 0; l < p.scales; ++l) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = lane; c < p.chunks; c += 32) {
      const float* r = p.partials + (((long)b * p.scales + l) * p.chunks + c) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += (double)r[k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
    }
    if (lane == 0) {
      const float hw = (float)p.lv[l].h * (float)p.lv[l].w;
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float den = (float)(s[2 * dir + 1] / hw) + 1e-12f;
        p.den[((long)b * p.scales + l) * 2 + dir] = den;
        total += (float)(s[2 * dir] / (3.0 * hw)) / den;
      }
    }
  }
  if (lane == 0) p.loss[b] = total;
}

// grid (chunks, B, 2 * levels), one direction per CTA.  grad_disp receives exactly two contributions per pixel (one per
// direction) by atomicAdd onto a zeroed map: fl(0 + a + b) == fl(0 + b + a), so the result does not depend on their order.
__global__ void __launch_bounds__(kRedThreads, UGL_DP_BWD_MINB) depth_photo_bwd_kernel(const __grid_constant__ DepthPhotoParams p) {
  __shared__ float sK[9], sP[12];
  __shared__ float red[(kRedThreads / 32) * 12];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const DepthPhotoLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const WarpGeom g = make_warp_geom(L.w, L.h);
  const long plane = (long)L.h * L.w;
  const float hw = (float)L.h * (float)L.w;
  const float ks = p.gloss[b] / (3.0f * hw) / p.den[((long)b * p.scales + l) * 2 + dir];
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  const ChunkRange cr = chunk_range(plane);
  long px = cr.begin + threadIdx.x;
  DepthDirect nxt = {};
  if (px < cr.end) nxt = depth_direct(L, b, dir, px);
#pragma unroll 1
  for (; px < cr.end; px += kRedThreads) {
    const DepthDirect dl = nxt;
    if (px + kRedThreads < cr.end) nxt = depth_direct(L, b, dir, px + kRedThreads);
    const int i = (int)(px / L.w), j = (int)(px % L.w);
    DepthPixel o;
    float gix, giy;
    depth_pixel<true>(L, sK, sP, b, dir, i, j, px, g, dl, o, &gix, &giy, ks);
    const float g_u = o.nc.ox ? 0.f : gix * g.sx;
    const float g_v = o.nc.oy ? 0.f : giy * g.sy;
    const float gD = project_backward(o.pr, sP, g_u, g_v, 0.f, acc);
    atomicAdd(&L.grad_disp[(long)b * plane + px], gD);
  }
  const float v = block_reduce_n<kRedThreads, 12>(acc, red);
  if (threadIdx.x < 12) p.partials[(((long)b * p.scales + l) * p.chunks + blockIdx.x) * 24 + 12 * dir + threadIdx.x] = v;
}

// ---- single-pass variant ---------------------------------------------------------------------------------------------
// The backward kernel above repeats the whole forward (projection, gathers, masks) to get at the per-pixel derivatives.  Like the
// flow kernel, this variant computes them in the forward launch while the taps are in registers: it writes the UN-normalised
// d loss / d disparity of each direction (one float per pixel and direction) and reduces the un-normalised d loss / d P sums; the
// per-sample normaliser 1 / mean(mask) and the upstream gradient are applied afterwards by an element-wise combine.  One gather
// kernel per step instead of two.  Partial rows: [B][scales][chunks][2][14] = {sum |I - rec| mask, sum mask, 12 P sums}.
constexpr int kDpAcc = 14;
__global__ void __launch_bounds__(kRedThreads, UGL_DP_BWD_MINB) depth_photo_fwdgrad_kernel(const __grid_constant__ DepthPhotoParams p) {
  __shared__ float sK[9], sP[12];
  __shared__ float red[(kRedThreads / 32) * kDpAcc];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const DepthPhotoLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const WarpGeom g = make_warp_geom(L.w, L.h);
  const long plane = (long)L.h * L.w;
  float* basis = p.basis[l] + ((long)b * 2 + dir) * plane;
  float acc[kDpAcc];
#pragma unroll
  for (int k = 0; k < kDpAcc; ++k) acc[k] = 0.f;
  const ChunkRange cr = chunk_range(plane);
  long px = cr.begin + threadIdx.x;
  DepthDirect nxt = {};
  if (px < cr.end) nxt = depth_direct(L, b, dir, px);
#pragma unroll 1
  for (; px < cr.end; px += kRedThreads) {
    const DepthDirect dl = nxt;
    if (px + kRedThreads < cr.end) nxt = depth_direct(L, b, dir, px + kRedThreads);
    const int i = (int)(px / L.w), j = (int)(px % L.w);
    DepthPixel o;
    float gix, giy;
    depth_pixel<true>(L, sK, sP, b, dir, i, j, px, g, dl, o, &gix, &giy, 1.0f);
    if (L.valid_out[dir]) L.valid_out[dir][(long)b * plane + px] = (fabsf(o.nc.gx) <= 1.0f && fabsf(o.nc.gy) <= 1.0f) ? 1.f : 0.f;
    if (L.tex_out[dir]) L.tex_out[dir][(long)b * plane + px] = o.tex;
    acc[12] += (fabsf(o.I[0] - o.rec[0]) + fabsf(o.I[1] - o.rec[1]) + fabsf(o.I[2] - o.rec[2])) * o.mask;
    acc[13] += o.mask;
    const float g_u = o.nc.ox ? 0.f : gix * g.sx;
    const float g_v = o.nc.oy ? 0.f : giy * g.sy;
    basis[px] = project_backward(o.pr, sP, g_u, g_v, 0.f, acc);      // adds the 12 P partials into acc[0..11]
  }
  const float v = block_reduce_n<kRedThreads, kDpAcc>(acc, red);
  if (threadIdx.x < kDpAcc) p.partials[((((long)b * p.scales + l) * p.chunks + blockIdx.x) * 2 + dir) * kDpAcc + threadIdx.x] = v;
}

// one CTA per sample: fixed-order fp64 sums of the chunk partials -> den, loss and the un-normalised P sums.  The 28 (direction,
// accumulator) columns of a level are summed by 9 thread groups over interleaved chunks, then across the groups in group order.
constexpr int kDpFinGroups = 9, kDpFinThreads = 256;
__global__ void __launch_bounds__(kDpFinThreads) depth_photo_fwdgrad_finalize_kernel(const __grid_constant__ DepthPhotoParams p) {
  __shared__ double part[kDpFinGroups][2 * kDpAcc];
  __shared__ double tot[2 * kDpAcc];
  const int b = blockIdx.x, t = threadIdx.x;
  const int grp = t / (2 * kDpAcc), col = t % (2 * kDpAcc);
  float total = 0.f;
  for (int l = 0; l < p.scales; ++l) {
    if (grp < kDpFinGroups) {
      const int dir = col / kDpAcc, k = col % kDpAcc;
      double s = 0.0;
      for (int c = grp; c < p.chunks; c += kDpFinGroups) s += (double)p.partials[((((long)b * p.scales + l) * p.chunks + c) * 2 + dir) * kDpAcc + k];
      part[grp][col] = s;
    }
    __syncthreads();
    if (t < 2 * kDpAcc) {
      double s = 0.0;
#pragma unroll
      for (int g = 0; g < kDpFinGroups; ++g) s += part[g][t];
      tot[t] = s;
      const int dir = t / kDpAcc, k = t % kDpAcc;
      if (k < 12) p.psum[(((long)b * p.scales + l) * 2 + dir) * 12 + k] = (float)s;
    }
    __syncthreads();
    if (t == 0) {
      const float hw = (float)p.lv[l].h * (float)p.lv[l].w;
#pragma unroll
      for (int dir = 0; dir < 2; ++dir) {
        const float den = (float)(tot[dir * kDpAcc + 13] / hw) + 1e-12f;
        p.den[((long)b * p.scales + l) * 2 + dir] = den;
        total += (float)(tot[dir * kDpAcc + 12] / (3.0 * hw)) / den;
      }
    }
  }
  if (t == 0) p.loss[b] = total;
}

// element-wise backward of the single-pass variant: grid (chunks, B, levels)
__global__ void __launch_bounds__(kRedThreads) depth_photo_combine_kernel(const __grid_constant__ DepthPhotoParams p) {
  const int b = blockIdx.y, l = blockIdx.z;
  const DepthPhotoLevel& L = p.lv[l];
  const long plane = (long)L.h * L.w;
  const float hw = (float)L.h * (float)L.w;
  const float k0 = p.gloss[b] / (3.0f * hw) / p.den[((long)b * p.scales + l) * 2 + 0];
  const float k1 = p.gloss[b] / (3.0f * hw) / p.den[((long)b * p.scales + l) * 2 + 1];
  const float* b0 = p.basis[l] + (long)b * 2 * plane;
  const float* b1 = b0 + plane;
  float* gd = L.grad_disp + (long)b * plane;
  if ((plane & 3) == 0) {
    const long q = plane >> 2;
    for (long i = blockIdx.x * (long)kRedThreads + threadIdx.x; i < q; i += (long)gridDim.x * kRedThreads) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(b0) + i), y = __ldcs(reinterpret_cast<const float4*>(b1) + i);
      reinterpret_cast<float4*>(gd)[i] = make_float4(k0 * x.x + k1 * y.x, k0 * x.y + k1 * y.y, k0 * x.z + k1 * y.z, k0 * x.w + k1 * y.w);
    }
  } else {
    for (long i = blockIdx.x * (long)kRedThreads + threadIdx.x; i < plane; i += (long)gridDim.x * kRedThreads) gd[i] = k0 * b0[i] + k1 * b1[i];
  }
  if (blockIdx.x == 0 && threadIdx.x < 24) {
    const int dir = threadIdx.x / 12, k = threadIdx.x % 12;
    p.grad_P[dir][l][b * 12 + k] = (dir == 0 ? k0 : k1) * p.psum[(((long)b * p.scales + l) * 2 + dir) * 12 + k];
  }
}

// backward of the depth-mode single-pass kernel (ugl_depth_ssim_forward_grad): the saved basis holds the un-normalised
// d loss / d (u, v) of the L1 and SSIM terms per direction; scale, chain through the projection -> grad_disp (two atomic
// contributions per pixel onto a zeroed map, order-free) and the grad_P partial sums.  grid (chunks, B, 2 * levels).
struct DepthSsimCombine {
  const float* basis[kMaxLevels];   // (B,8,h,w)
  const float* stats;               // (B,scales,GA_COUNT)
  const float* gloss;               // (2,B)
};
__global__ void __launch_bounds__(kRedThreads, 3) depth_ssim_combine_kernel(const __grid_constant__ DepthPhotoParams p,
                                                                            const __grid_constant__ DepthSsimCombine q) {
  __shared__ float sK[9], sP[12];
  __shared__ float red[(kRedThreads / 32) * 12];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const DepthPhotoLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const long plane = (long)L.h * L.w;
  float k_pix, k_ssim;
  depth_combine_scales(q.stats + ((long)b * p.scales + l) * GA_COUNT, L.h, L.w, q.gloss, p.B, b, dir, k_pix, k_ssim);
  const float* bs = q.basis[l] + ((long)b * 8 + 4 * dir) * plane;
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  // contiguous chunk per CTA, the five coalesced loads of the next iteration in flight under the current one (as the forward kernels)
  struct Direct { float b0, b1, b2, b3, D; };
  auto direct = [&](long q) {
    Direct d;
    d.b0 = __ldcs(bs + q); d.b1 = __ldcs(bs + plane + q); d.b2 = __ldcs(bs + 2 * plane + q); d.b3 = __ldcs(bs + 3 * plane + q);
    d.D = L.disp[(long)b * plane + q];
    return d;
  };
  const ChunkRange cr = chunk_range(plane);
  long px = cr.begin + threadIdx.x;
  Direct nxt = {};
  if (px < cr.end) nxt = direct(px);
#pragma unroll 1
  for (; px < cr.end; px += kRedThreads) {
    const Direct cur = nxt;
    if (px + kRedThreads < cr.end) nxt = direct(px + kRedThreads);
    const int i = (int)(px / L.w), j = (int)(px % L.w);
    const float gu = k_pix * cur.b0 + k_ssim * cur.b2;
    const float gv = k_pix * cur.b1 + k_ssim * cur.b3;
    const Projected r = project_pixel(sK, sP, cur.D, j, i);
    const float gD = project_backward(r, sP, gu, gv, 0.f, acc);
    atomicAdd(&L.grad_disp[(long)b * plane + px], gD);
  }
  const float v = block_reduce_n<kRedThreads, 12>(acc, red);
  if (threadIdx.x < 12) p.partials[(((long)b * p.scales + l) * p.chunks + blockIdx.x) * 24 + 12 * dir + threadIdx.x] = v;
}

// one warp per (sample, level): grad_P[dir][level][b] = fixed-order sum of the chunk partials
__global__ void depth_photo_bwd_finalize_kernel(const __grid_constant__ DepthPhotoParams p) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= p.B * p.scales) return;
  const int b = w / p.scales, l = w % p.scales;
  if (lane < 24) {
    double s = 0.0;
    for (int c = 0; c < p.chunks; ++c) s += (double)p.partials[(((long)b * p.scales + l) * p.chunks + c) * 24 + lane];
    p.grad_P[lane / 12][l][b * 12 + lane % 12] = (float)s;
  }
}

}  // namespace ugl

using namespace ugl;

static int depth_photo_fill(const UglDepthPhotoArgs* a, bool backward, DepthPhotoParams& p) {
  if (!a) return fail(UGL_EINVAL, "depth_photo: null args");
  if (a->batch <= 0 || a->batch > 65535 || a->scales <= 0 || a->scales > UGL_MAX_LEVELS)
    return fail(UGL_EINVAL, "depth_photo: bad batch/scales (%d/%d)", a->batch, a->scales);
  p.B = a->batch; p.scales = a->scales;
  long max_plane = 0;
  for (int l = 0; l < a->scales; ++l) {
    DepthPhotoLevel& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l];
    if (L.h < 2 || L.w < 2) return fail(UGL_EUNSUPPORTED, "depth_photo: level %d is %dx%d", l, L.h, L.w);
    L.img = a->img[l]; L.disp = a->disp[l]; L.Kinv = a->Kinv[l];
    L.ext_bytes = a->ext_bytes[l]; L.ext_need[0] = (unsigned)a->ext_need[0]; L.ext_need[1] = (unsigned)a->ext_need[1];
    if (!L.img || !L.disp || !L.Kinv) return fail(UGL_EINVAL, "depth_photo: null input at level %d", l);
    for (int d = 0; d < 2; ++d) {
      L.src_area[d] = a->src_area[d][l]; L.src_bil[d] = a->src_bil[d][l]; L.P[d] = a->P[d][l];
      L.ext_mask[d] = a->ext_mask[d][l]; L.valid_out[d] = a->valid_out[d][l]; L.tex_out[d] = a->tex_out[d][l];
      if (!L.src_area[d] || !L.src_bil[d] || !L.P[d]) return fail(UGL_EINVAL, "depth_photo: null input at level %d", l);
      p.grad_P[d][l] = backward ? a->grad_P[d][l] : nullptr;
      if (backward && !p.grad_P[d][l]) return fail(UGL_EINVAL, "depth_photo: null grad_P at level %d", l);
    }
    L.grad_disp = backward ? a->grad_disp[l] : nullptr;
    if (backward && !L.grad_disp) return fail(UGL_EINVAL, "depth_photo: null grad_disp at level %d", l);
    const long pl = (long)L.h * L.w;
    max_plane = pl > max_plane ? pl : max_plane;
  }
  p.chunks = reduce_chunks(max_plane);
  if (!a->den) return fail(UGL_EINVAL, "depth_photo: null den");
  p.den = a->den; p.loss = a->loss; p.gloss = a->grad_loss;
  p.partials = static_cast<float*>(a->workspace);
  const uint64_t need = (uint64_t)p.B * p.scales * p.chunks * 24 * sizeof(float);
  if (!a->workspace || a->workspace_bytes < need) return fail(UGL_EWORKSPACE, "depth_photo: workspace too small (%llu < %llu)",
                                                              (unsigned long long)a->workspace_bytes, (unsigned long long)need);
  return UGL_OK;
}

extern "C" uint64_t ugl_depth_photo_workspace_bytes(const UglDepthPhotoArgs* a) {
  if (!a) return 0;
  long max_plane = 0;
  for (int l = 0; l < a->scales && l < UGL_MAX_LEVELS; ++l) {
    const long pl = (long)a->height[l] * a->width[l];
    max_plane = pl > max_plane ? pl : max_plane;
  }
  return (uint64_t)a->batch * a->scales * reduce_chunks(max_plane) * 24 * sizeof(float);
}

extern "C" int ugl_depth_photo_forward(const UglDepthPhotoArgs* a) {
  DepthPhotoParams p;
  int rc = depth_photo_fill(a, false, p);
  if (rc) return rc;
  if (!a->loss) return fail(UGL_EINVAL, "depth_photo_forward: null loss");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  depth_photo_fwd_kernel<<<dim3(p.chunks, p.B, 2 * p.scales), kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_photo_fwd_kernel"))) return rc;
  depth_photo_finalize_kernel<<<(p.B + 3) / 4, 128, 0, st>>>(p);
  return check_launch("depth_photo_finalize_kernel");
}

extern "C" int ugl_depth_photo_backward(const UglDepthPhotoArgs* a) {
  DepthPhotoParams p;
  int rc = depth_photo_fill(a, true, p);
  if (rc) return rc;
  if (!a->grad_loss) return fail(UGL_EINVAL, "depth_photo_backward: null grad_loss");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  for (int l = 0; l < p.scales; ++l) {     // the two directions accumulate into grad_disp
    const cudaError_t e = cudaMemsetAsync(p.lv[l].grad_disp, 0, sizeof(float) * (size_t)p.B * p.lv[l].h * p.lv[l].w, st);
    if (e != cudaSuccess) return fail((int)e, "depth_photo_backward: memset: %s", cudaGetErrorString(e));
  }
  depth_photo_bwd_kernel<<<dim3(p.chunks, p.B, 2 * p.scales), kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_photo_bwd_kernel"))) return rc;
  depth_photo_bwd_finalize_kernel<<<(p.B * p.scales + 3) / 4, 128, 0, st>>>(p);
  return check_launch("depth_photo_bwd_finalize_kernel");
}

extern "C" int ugl_depth_ssim_combine(const UglDepthSsimArgs* g) {
  if (!g) return fail(UGL_EINVAL, "depth_ssim_combine: null args");
  const UglDepthPhotoArgs* a = &g->photo;
  DepthPhotoParams p;
  // the combine reads disp, Kinv, P (and the level sizes) of the photo args; images are not needed
  if (a->batch <= 0 || a->batch > 65535 || a->scales <= 0 || a->scales > UGL_MAX_LEVELS)
    return fail(UGL_EINVAL, "depth_ssim_combine: bad batch/scales (%d/%d)", a->batch, a->scales);
  if (!g->grad_loss4 || !g->stats) return fail(UGL_EINVAL, "depth_ssim_combine: null grad_loss / stats");
  p.B = a->batch; p.scales = a->scales;
  DepthSsimCombine q;
  q.stats = g->stats; q.gloss = g->grad_loss4;
  long max_plane = 0;
  for (int l = 0; l < a->scales; ++l) {
    DepthPhotoLevel& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l];
    L.disp = a->disp[l]; L.Kinv = a->Kinv[l]; L.grad_disp = a->grad_disp[l];
    if (!L.disp || !L.Kinv || !L.grad_disp || !g->basis[l]) return fail(UGL_EINVAL, "depth_ssim_combine: null pointer at level %d", l);
    q.basis[l] = g->basis[l];
    for (int d = 0; d < 2; ++d) {
      L.P[d] = a->P[d][l]; p.grad_P[d][l] = a->grad_P[d][l];
      if (!L.P[d] || !p.grad_P[d][l]) return fail(UGL_EINVAL, "depth_ssim_combine: null P / grad_P at level %d", l);
    }
    const long pl = (long)L.h * L.w;
    max_plane = pl > max_plane ? pl : max_plane;
  }
  p.chunks = reduce_chunks(max_plane);
  const uint64_t need = (uint64_t)p.B * p.scales * p.chunks * 24 * sizeof(float);
  if (!a->workspace || a->workspace_bytes < need)
    return fail(UGL_EWORKSPACE, "depth_ssim_combine: workspace too small (%llu < %llu)", (unsigned long long)a->workspace_bytes, (unsigned long long)need);
  p.partials = static_cast<float*>(a->workspace);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  for (int l = 0; l < p.scales; ++l) {
    const cudaError_t e = cudaMemsetAsync(p.lv[l].grad_disp, 0, sizeof(float) * (size_t)p.B * p.lv[l].h * p.lv[l].w, st);
    if (e != cudaSuccess) return fail((int)e, "depth_ssim_combine: memset: %s", cudaGetErrorString(e));
  }
  depth_ssim_combine_kernel<<<dim3(p.chunks, p.B, 2 * p.scales), kRedThreads, 0, st>>>(p, q);
  int rc = check_launch("depth_ssim_combine_kernel");
  if (rc) return rc;
  depth_photo_bwd_finalize_kernel<<<(p.B * p.scales + 3) / 4, 128, 0, st>>>(p);
  return check_launch("depth_photo_bwd_finalize_kernel");
}

// ---- single-pass variant: forward_grad + combine --------------------------------------------------------------------
static uint64_t depth_photo_grad_need(int B, int scales, int chunks) { return (uint64_t)B * scales * chunks * 2 * ugl::kDpAcc * sizeof(float); }

extern "C" uint64_t ugl_depth_photo_grad_workspace_bytes(const UglDepthPhotoGradArgs* g) {
  if (!g) return 0;
  const UglDepthPhotoArgs* a = &g->photo;
  long max_plane = 0;
  for (int l = 0; l < a->scales && l < UGL_MAX_LEVELS; ++l) {
    const long pl = (long)a->height[l] * a->width[l];
    max_plane = pl > max_plane ? pl : max_plane;
  }
  const uint64_t base = ugl_depth_photo_workspace_bytes(a), need = depth_photo_grad_need(a->batch, a->scales, reduce_chunks(max_plane));
  return need > base ? need : base;
}

extern "C" int ugl_depth_photo_forward_grad(const UglDepthPhotoGradArgs* g) {
  if (!g) return fail(UGL_EINVAL, "depth_photo_forward_grad: null args");
  const UglDepthPhotoArgs* a = &g->photo;
  DepthPhotoParams p;
  int rc = depth_photo_fill(a, false, p);
  if (rc) return rc;
  if (!a->loss || !g->psum) return fail(UGL_EINVAL, "depth_photo_forward_grad: null loss / psum");
  if (a->workspace_bytes < depth_photo_grad_need(p.B, p.scales, p.chunks))
    return fail(UGL_EWORKSPACE, "depth_photo_forward_grad: workspace too small (%llu bytes)", (unsigned long long)a->workspace_bytes);
  for (int l = 0; l < p.scales; ++l) {
    if (!g->basis[l]) return fail(UGL_EINVAL, "depth_photo_forward_grad: null basis at level %d", l);
    p.basis[l] = g->basis[l];
  }
  p.psum = g->psum;
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  depth_photo_fwdgrad_kernel<<<dim3(p.chunks, p.B, 2 * p.scales), kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_photo_fwdgrad_kernel"))) return rc;
  depth_photo_fwdgrad_finalize_kernel<<<p.B, kDpFinThreads, 0, st>>>(p);
  return check_launch("depth_photo_fwdgrad_finalize_kernel");
}

extern "C" int ugl_depth_photo_combine(const UglDepthPhotoGradArgs* g) {
  if (!g) return fail(UGL_EINVAL, "depth_photo_combine: null args");
  const UglDepthPhotoArgs* a = &g->photo;
  if (a->batch <= 0 || a->batch > 65535 || a->scales <= 0 || a->scales > UGL_MAX_LEVELS)
    return fail(UGL_EINVAL, "depth_photo_combine: bad batch/scales (%d/%d)", a->batch, a->scales);
  if (!a->grad_loss || !a->den || !g->psum) return fail(UGL_EINVAL, "depth_photo_combine: null grad_loss / den / psum");
  DepthPhotoParams p;
  p.B = a->batch; p.scales = a->scales; p.den = a->den; p.gloss = a->grad_loss; p.psum = g->psum;
  long max_plane = 0;
  for (int l = 0; l < a->scales; ++l) {
    DepthPhotoLevel& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l]; L.grad_disp = a->grad_disp[l];
    p.basis[l] = g->basis[l];
    if (L.h < 2 || L.w < 2 || !L.grad_disp || !p.basis[l]) return fail(UGL_EINVAL, "depth_photo_combine: bad level %d", l);
    if ((reinterpret_cast<uintptr_t>(L.grad_disp) | reinterpret_cast<uintptr_t>(p.basis[l])) & 15u)
      return fail(UGL_EALIGN, "depth_photo_combine: level %d not 16-byte aligned", l);
    for (int d = 0; d < 2; ++d) {
      p.grad_P[d][l] = a->grad_P[d][l];
      if (!p.grad_P[d][l]) return fail(UGL_EINVAL, "depth_photo_combine: null grad_P at level %d", l);
    }
    const long pl = (long)L.h * L.w;
    max_plane = pl > max_plane ? pl : max_plane;
  }
  p.chunks = reduce_chunks(max_plane);
  depth_photo_combine_kernel<<<dim3(p.chunks, p.B, p.scales), kRedThreads, 0, static_cast<cudaStream_t>(a->stream)>>>(p);
  return check_launch("depth_photo_combine_kernel");
}
