// Depth-consistency term (compute_consis_loss, unmasked: model_depth.py:154-163; enabled in model_depth_texture.py:308-309,
// disabled in the live Model_depth :333-335) for both source frames and all levels in one forward and one backward pass:
//   per pixel: project the centre disparity with P = K_s [R|t] (inverse_warp2, structures/inverse_warp.py:263-303), computed depth
//   = clamped Z, projected depth = bilinear sample of the source frame's disparity (zeros padding) clamped at 1e-3,
//   d = clamp(|comp - proj| / |comp + proj|, 0, 1); loss = sum over levels and frames of mean(d).
// Backward: dense grad of the centre disparity (two order-free atomic contributions per pixel), grad_P by the fixed-order
// two-stage reduction, and the gradient of the SOURCE disparities — a scatter through the bilinear taps — accumulated in 64-bit
// fixed point (ugl_scatter.cuh), i.e. bit-reproducible.  Replaces 84 per-method launches per depth step with 8.
#include "ugl_common.cuh"
#include "ugl_geometry.cuh"
#include "ugl_reduce.cuh"
#include "ugl_scatter.cuh"

namespace ugl {

struct ConsisLevel {
  int h, w;
  const float* disp;          // (B,1,h,w)
  const float* ref[2];        // (B,1,h,w) source-frame disparities
  const float* Kinv;          // (B,3,3)
  const float* P[2];          // (B,3,4)
  float* gdisp;               // (B,1,h,w)
  float* gref[2];             // (B,1,h,w)
  float* coef[2];             // scratch (B,h,w): d loss / d projected depth
  unsigned long long* fix[2]; // scratch (B,h,w): fixed-point accumulators
};
struct ConsisParams {
  int B, scales, chunks;
  ConsisLevel lv[kMaxLevels];
  float* partials;            // fwd [B][scales][chunks][2]; bwd [B][scales][chunks][24]
  unsigned* maxbits;          // [scales][2]
  float* loss;                // (B,)
  const float* gloss;         // (B,)
  float* gP[2][kMaxLevels];   // (B,3,4)
};

struct ConsisPixel { Projected pr; NormCoord nc; Tap tap; Corners cs; float sampled, proj, comp; };
__device__ __forceinline__ ConsisPixel consis_pixel(const ConsisLevel& L, const float* sK, const float* sP, int b, int dir, int i, int j,
                                                    long px, const WarpGeom& g) {
  ConsisPixel o;
  const long plane = (long)L.h * L.w;
  o.pr = project_pixel(sK, sP, L.disp[(long)b * plane + px], j, i);
  o.nc = normalise(o.pr, g);
  o.tap = make_tap(unnormalize(o.nc.gx, L.w), unnormalize(o.nc.gy, L.h), L.w, L.h);
  o.cs = tap_fetch(L.ref[dir] + (long)b * plane, L.w, o.tap);
  o.sampled = corners_value(o.cs, o.tap);
  o.proj = o.sampled < kDepthMin ? kDepthMin : o.sampled;
  o.comp = o.pr.Z;
  return o;
}

// grid (chunks, B, 2 * levels)
__global__ void __launch_bounds__(kRedThreads, 3) depth_consis_fwd_kernel(const __grid_constant__ ConsisParams p) {
  __shared__ float sK[9], sP[12];
  __shared__ float red[kRedThreads / 32];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const ConsisLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const WarpGeom g = make_warp_geom(L.w, L.h);
  const long plane = (long)L.h * L.w;
  float acc[1] = {0.f};
  for (long px = blockIdx.x * (long)kRedThreads + threadIdx.x; px < plane; px += (long)gridDim.x * kRedThreads) {
    const ConsisPixel o = consis_pixel(L, sK, sP, b, dir, (int)(px / L.w), (int)(px % L.w), px, g);
    const float v = div_rn(fabsf(sub_rn(o.comp, o.proj)), fabsf(add_rn(o.comp, o.proj)));
    acc[0] += v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
  }
  const float v = block_reduce_n<kRedThreads, 1>(acc, red);
  if (threadIdx.x == 0) p.partials[(((long)b * p.scales + l) * p.chunks + blockIdx.x) * 2 + dir] = v;
}

// one warp per sample: fixed-order fp64 sums, loss[b] = sum over frames (left first) and levels of mean(d)
__global__ void depth_consis_finalize_kernel(const __grid_constant__ ConsisParams p) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= p.B) return;
  float total = 0.f;
  for (int dir = 0; dir < 2; ++dir)
    for (int l = 0; l < p.scales; ++l) {
      double s = 0.0;
      for (int c = lane; c < p.chunks; c += 32) s += (double)p.partials[(((long)b * p.scales + l) * p.chunks + c) * 2 + dir];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      total += (float)(s / ((double)p.lv[l].h * (double)p.lv[l].w));
    }
  if (lane == 0) p.loss[b] = total;
}

__global__ void __launch_bounds__(kRedThreads, 2) depth_consis_bwd_kernel(const __grid_constant__ ConsisParams p) {
  __shared__ float sK[9], sP[12];
  __shared__ float red[(kRedThreads / 32) * 12];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const ConsisLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const WarpGeom g = make_warp_geom(L.w, L.h);
  const long plane = (long)L.h * L.w;
  const float go = p.gloss[b] / ((float)L.h * (float)L.w);
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  float cmax = 0.f;
  for (long px = blockIdx.x * (long)kRedThreads + threadIdx.x; px < plane; px += (long)gridDim.x * kRedThreads) {
    const ConsisPixel o = consis_pixel(L, sK, sP, b, dir, (int)(px / L.w), (int)(px % L.w), px, g);
    const float n = o.comp - o.proj, d = o.comp + o.proj;
    const float an = fabsf(n), ad = fabsf(d);
    const float v = an / ad;
    const float gv = (v >= 0.f && v <= 1.f) ? go : 0.f;
    const float dn = gv * sgnf(n) / ad;               // d/d n
    const float dd = -gv * an / (ad * ad) * sgnf(d);  // d/d d
    const float g_comp = dn + dd;
    const float g_proj = (o.sampled >= kDepthMin) ? (-dn + dd) : 0.f;     // clamp(min=1e-3) of the sampled depth
    const float gix = g_proj * corners_ddx(o.cs, o.tap), giy = g_proj * corners_ddy(o.cs, o.tap);
    const float g_u = o.nc.ox ? 0.f : gix * g.sx;
    const float g_v = o.nc.oy ? 0.f : giy * g.sy;
    const float gD = project_backward(o.pr, sP, g_u, g_v, g_comp, acc);
    atomicAdd(&L.gdisp[(long)b * plane + px], gD);
    L.coef[dir][(long)b * plane + px] = g_proj;
    const float a = fabsf(g_proj);
    cmax = (a > cmax && a <= 3.0e38f) ? a : cmax;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  if ((threadIdx.x & 31) == 0) atomicMax(p.maxbits + l * 2 + dir, __float_as_uint(cmax));
  const float v = block_reduce_n<kRedThreads, 12>(acc, red);
  if (threadIdx.x < 12) p.partials[(((long)b * p.scales + l) * p.chunks + blockIdx.x) * 24 + 12 * dir + threadIdx.x] = v;
}

__global__ void __launch_bounds__(kRedThreads, 3) depth_consis_scatter_kernel(const __grid_constant__ ConsisParams p) {
  __shared__ float sK[9], sP[12];
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const ConsisLevel& L = p.lv[l];
  if (threadIdx.x < 9) sK[threadIdx.x] = L.Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = L.P[dir][b * 12 + threadIdx.x];
  __syncthreads();
  const WarpGeom g = make_warp_geom(L.w, L.h);
  const long plane = (long)L.h * L.w;
  const int e = fixed_point_exponent(__uint_as_float(p.maxbits[l * 2 + dir]), plane);
  for (long px = blockIdx.x * (long)kRedThreads + threadIdx.x; px < plane; px += (long)gridDim.x * kRedThreads) {
    const float c = L.coef[dir][(long)b * plane + px];
    if (c == 0.f) continue;
    const Projected r = project_pixel(sK, sP, L.disp[(long)b * plane + px], (int)(px % L.w), (int)(px / L.w));
    const NormCoord nc = normalise(r, g);
    const Tap t = make_tap(unnormalize(nc.gx, L.w), unnormalize(nc.gy, L.h), L.w, L.h);
    if (t.inb == 0u) continue;
    scatter_tap(L.fix[dir] + (long)b * plane, L.w, t, c, e);
  }
}

__global__ void __launch_bounds__(kRedThreads) depth_consis_convert_kernel(const __grid_constant__ ConsisParams p) {
  const int b = blockIdx.y, l = blockIdx.z >> 1, dir = blockIdx.z & 1;
  const ConsisLevel& L = p.lv[l];
  const long plane = (long)L.h * L.w;
  const int e = fixed_point_exponent(__uint_as_float(p.maxbits[l * 2 + dir]), plane);
  for (long px = blockIdx.x * (long)kRedThreads + threadIdx.x; px < plane; px += (long)gridDim.x * kRedThreads)
    L.gref[dir][(long)b * plane + px] = (float)ldexp((double)(long long)L.fix[dir][(long)b * plane + px], -e);
}

__global__ void depth_consis_bwd_finalize_kernel(const __grid_constant__ ConsisParams p) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= p.B * p.scales) return;
  const int b = w / p.scales, l = w % p.scales;
  if (lane < 24) {
    double s = 0.0;
    for (int c = 0; c < p.chunks; ++c) s += (double)p.partials[(((long)b * p.scales + l) * p.chunks + c) * 24 + lane];
    p.gP[lane / 12][l][b * 12 + lane % 12] = (float)s;
  }
}

static uint64_t consis_pixels(const UglDepthConsisArgs* a, long* max_plane) {
  uint64_t n = 0;
  long mp = 0;
  for (int l = 0; l < a->scales && l < UGL_MAX_LEVELS; ++l) {
    const long pl = (long)a->height[l] * a->width[l];
    n += (uint64_t)pl * a->batch;
    mp = pl > mp ? pl : mp;
  }
  if (max_plane) *max_plane = mp;
  return n;
}

static int consis_fill(const UglDepthConsisArgs* a, bool backward, ConsisParams& p) {
  if (!a) return fail(UGL_EINVAL, "depth_consis: null args");
  if (a->batch <= 0 || a->batch > 65535 || a->scales <= 0 || a->scales > UGL_MAX_LEVELS)
    return fail(UGL_EINVAL, "depth_consis: bad batch/scales (%d/%d)", a->batch, a->scales);
  p.B = a->batch; p.scales = a->scales;
  long max_plane = 0;
  const uint64_t npix = consis_pixels(a, &max_plane);
  p.chunks = reduce_chunks(max_plane);
  if (!a->workspace || a->workspace_bytes < ugl_depth_consis_workspace_bytes(a))
    return fail(UGL_EWORKSPACE, "depth_consis: workspace too small (%llu bytes given)", (unsigned long long)a->workspace_bytes);
  if (reinterpret_cast<uintptr_t>(a->workspace) & 7u) return fail(UGL_EALIGN, "depth_consis: workspace not 8-byte aligned");
  // layout: [fixed-point planes 2 * npix * 8][maxbits 256][coef 2 * npix * 4][partials]
  char* cur = static_cast<char*>(a->workspace);
  unsigned long long* fix = reinterpret_cast<unsigned long long*>(cur); cur += 2 * npix * 8;
  p.maxbits = reinterpret_cast<unsigned*>(cur); cur += 256;
  float* coef = reinterpret_cast<float*>(cur); cur += 2 * npix * 4;
  p.partials = reinterpret_cast<float*>(cur);
  uint64_t off = 0;
  for (int l = 0; l < a->scales; ++l) {
    ConsisLevel& L = p.lv[l];
    L.h = a->height[l]; L.w = a->width[l];
    if (L.h < 2 || L.w < 2) return fail(UGL_EUNSUPPORTED, "depth_consis: level %d is %dx%d", l, L.h, L.w);
    L.disp = a->disp[l]; L.Kinv = a->Kinv[l];
    if (!L.disp || !L.Kinv) return fail(UGL_EINVAL, "depth_consis: null input at level %d", l);
    L.gdisp = backward ? a->grad_disp[l] : nullptr;
    if (backward && !L.gdisp) return fail(UGL_EINVAL, "depth_consis: null grad_disp at level %d", l);
    const uint64_t n = (uint64_t)L.h * L.w * a->batch;
    for (int d = 0; d < 2; ++d) {
      L.ref[d] = a->ref_disp[d][l]; L.P[d] = a->P[d][l];
      if (!L.ref[d] || !L.P[d]) return fail(UGL_EINVAL, "depth_consis: null ref_disp / P at level %d", l);
      L.gref[d] = backward ? a->grad_ref[d][l] : nullptr;
      p.gP[d][l] = backward ? a->grad_P[d][l] : nullptr;
      if (backward && (!L.gref[d] || !p.gP[d][l])) return fail(UGL_EINVAL, "depth_consis: null grad_ref / grad_P at level %d", l);
      L.coef[d] = coef + off; L.fix[d] = fix + off;
      off += n;
    }
  }
  p.loss = a->loss; p.gloss = a->grad_loss;
  return UGL_OK;
}

}  // namespace ugl

using namespace ugl;

extern "C" uint64_t ugl_depth_consis_workspace_bytes(const UglDepthConsisArgs* a) {
  if (!a) return 0;
  long max_plane = 0;
  const uint64_t npix = consis_pixels(a, &max_plane);
  return 2 * npix * 8 + 256 + 2 * npix * 4 + (uint64_t)a->batch * a->scales * reduce_chunks(max_plane) * 24 * sizeof(float);
}

extern "C" int ugl_depth_consis_forward(const UglDepthConsisArgs* a) {
  ConsisParams p;
  int rc = consis_fill(a, false, p);
  if (rc) return rc;
  if (!a->loss) return fail(UGL_EINVAL, "depth_consis_forward: null loss");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  depth_consis_fwd_kernel<<<dim3(p.chunks, p.B, 2 * p.scales), kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_consis_fwd_kernel"))) return rc;
  depth_consis_finalize_kernel<<<(p.B + 3) / 4, 128, 0, st>>>(p);
  return check_launch("depth_consis_finalize_kernel");
}

extern "C" int ugl_depth_consis_backward(const UglDepthConsisArgs* a) {
  ConsisParams p;
  int rc = consis_fill(a, true, p);
  if (rc) return rc;
  if (!a->grad_loss) return fail(UGL_EINVAL, "depth_consis_backward: null grad_loss");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  // zero the fixed-point planes and the max trackers (one contiguous region), and the dense disparity gradients
  cudaError_t e = cudaMemsetAsync(a->workspace, 0, (size_t)(reinterpret_cast<char*>(p.maxbits) - static_cast<char*>(a->workspace)) + 256, st);
  for (int l = 0; l < p.scales && e == cudaSuccess; ++l)
    e = cudaMemsetAsync(p.lv[l].gdisp, 0, sizeof(float) * (size_t)p.B * p.lv[l].h * p.lv[l].w, st);
  if (e != cudaSuccess) return fail((int)e, "depth_consis_backward: memset: %s", cudaGetErrorString(e));
  const dim3 grid(p.chunks, p.B, 2 * p.scales);
  depth_consis_bwd_kernel<<<grid, kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_consis_bwd_kernel"))) return rc;
  depth_consis_scatter_kernel<<<grid, kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_consis_scatter_kernel"))) return rc;
  depth_consis_convert_kernel<<<grid, kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("depth_consis_convert_kernel"))) return rc;
  depth_consis_bwd_finalize_kernel<<<(p.B * p.scales + 3) / 4, 128, 0, st>>>(p);
  return check_launch("depth_consis_bwd_finalize_kernel");
}
