// Depth + pose reprojection (inverse_warp2), rigid flow (calculate_rigid_flow) and the epipolar
// distance map (compute_epipolar_map).  Reference: core/networks/structures/inverse_warp.py:227-342,
// core/networks/model_geometry.py:355-403.
//
// The 3x3 / 3x4 per-sample matrices (K^-1, P = K [R|t], F = K^-T [t]x R K^-1) are inputs: the host side
// builds them with the very ops the reference uses (torch.inverse, bmm, euler2mat) under autograd, and
// these kernels return d loss / d P and d loss / d F (12 / 9 numbers per sample, deterministic two-stage
// reduction) so the chain to the 6-DoF pose stays in the reference's own arithmetic.
#include "ugl_common.cuh"
#include "ugl_geometry.cuh"
#include "ugl_reduce.cuh"
#include "ugl_scatter.cuh"

namespace ugl {

__device__ __forceinline__ void load_mats(const float* Kinv, const float* P, int b, float* sK, float* sP) {
  if (threadIdx.x < 9) sK[threadIdx.x] = Kinv[b * 9 + threadIdx.x];
  if (threadIdx.x < 12) sP[threadIdx.x] = P[b * 12 + threadIdx.x];
  __syncthreads();
}

// ---- inverse_warp2 forward ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
reproject_fwd_kernel(const float* __restrict__ img, const float* __restrict__ depth, const float* __restrict__ ref_depth,
                     const float* __restrict__ Kinv, const float* __restrict__ P, int C, int H, int W, float* __restrict__ out,
                     float* __restrict__ valid, float* __restrict__ proj_depth, float* __restrict__ comp_depth) {
  __shared__ float sK[9], sP[12];
  const int b = blockIdx.y;
  load_mats(Kinv, P, b, sK, sP);
  const WarpGeom g = make_warp_geom(W, H);
  const long plane = (long)H * W;
  for (long p = blockIdx.x * (long)kRedThreads + threadIdx.x; p < plane; p += (long)gridDim.x * kRedThreads) {
    const int i = (int)(p / W), j = (int)(p % W);
    const Projected r = project_pixel(sK, sP, depth[(long)b * plane + p], j, i);
    const NormCoord n = normalise(r, g);
    const Tap t = make_tap(unnormalize(n.gx, W), unnormalize(n.gy, H), W, H);
    for (int c = 0; c < C; ++c) {
      const Corners k = tap_fetch(img + ((long)b * C + c) * plane, W, t);
      out[((long)b * C + c) * plane + p] = corners_value(k, t);
    }
    if (valid) valid[(long)b * plane + p] = (fabsf(n.gx) <= 1.0f && fabsf(n.gy) <= 1.0f) ? 1.f : 0.f;
    if (proj_depth) {
      const float d = corners_value(tap_fetch(ref_depth + (long)b * plane, W, t), t);
      proj_depth[(long)b * plane + p] = d < kDepthMin ? kDepthMin : d;
    }
    if (comp_depth) comp_depth[(long)b * plane + p] = r.Z;
  }
}

struct ReprojectBwdPixel {
  const float *img, *depth, *ref_depth, *Kinv, *P, *go_img, *go_proj, *go_comp;
  float* g_depth;
  int C, H, W;
  __device__ void operator()(int b, long p, float* acc) const {
    const long plane = (long)H * W;
    const int i = (int)(p / W), j = (int)(p % W);
    const WarpGeom g = make_warp_geom(W, H);
    const float* Kb = Kinv + b * 9;
    const float* Pb = P + b * 12;
    const Projected r = project_pixel(Kb, Pb, depth[(long)b * plane + p], j, i);
    const NormCoord n = normalise(r, g);
    const Tap t = make_tap(unnormalize(n.gx, W), unnormalize(n.gy, H), W, H);
    float gix = 0.f, giy = 0.f;
    if (go_img) {
      for (int c = 0; c < C; ++c) {
        const Corners k = tap_fetch(img + ((long)b * C + c) * plane, W, t);
        const float go = go_img[((long)b * C + c) * plane + p];
        gix += go * corners_ddx(k, t);
        giy += go * corners_ddy(k, t);
      }
    }
    if (go_proj) {
      const Corners k = tap_fetch(ref_depth + (long)b * plane, W, t);
      if (corners_value(k, t) >= kDepthMin) {
        const float go = go_proj[(long)b * plane + p];
        gix += go * corners_ddx(k, t);
        giy += go * corners_ddy(k, t);
      }
    }
    // ix = (gx+1) W/2 - 1/2 ; gx = 2 u/(W-1) - 1 (constant 2 where it was overwritten)
    const float g_u = n.ox ? 0.f : gix * g.sx;
    const float g_v = n.oy ? 0.f : giy * g.sy;
    const float g_Z = go_comp ? go_comp[(long)b * plane + p] : 0.f;
    const float gD = project_backward(r, Pb, g_u, g_v, g_Z, acc);
    if (g_depth) g_depth[(long)b * plane + p] = gD;
  }
};

struct MatFinal {   // per-sample n-vector
  float* out;
  int n;
  __device__ void operator()(int b, const double* S) const {
    for (int k = 0; k < n; ++k) out[b * n + k] = (float)S[k];
  }
};

// scatter part of the backward: d loss / d img and d loss / d ref_depth (fixed point, deterministic)
__global__ void __launch_bounds__(256)
reproject_scatter_kernel(const float* __restrict__ depth, const float* __restrict__ ref_depth, const float* __restrict__ Kinv,
                         const float* __restrict__ P, const float* __restrict__ go_img, const float* __restrict__ go_proj, int B,
                         int C, int H, int W, const unsigned* __restrict__ maxbits, unsigned long long* __restrict__ acc_img,
                         unsigned long long* __restrict__ acc_ref) {
  const long plane = (long)H * W, n = (long)B * plane;
  const WarpGeom g = make_warp_geom(W, H);
  const int e_img = fixed_point_exponent(__uint_as_float(maxbits[0]), plane);
  const int e_ref = fixed_point_exponent(__uint_as_float(maxbits[1]), plane);
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < n; idx += (long)gridDim.x * blockDim.x) {
    const long p = idx % plane;
    const int b = (int)(idx / plane);
    const int i = (int)(p / W), j = (int)(p % W);
    const Projected r = project_pixel(Kinv + b * 9, P + b * 12, depth[idx], j, i);
    const NormCoord nc = normalise(r, g);
    const Tap t = make_tap(unnormalize(nc.gx, W), unnormalize(nc.gy, H), W, H);
    if (t.inb == 0u) continue;
    if (acc_img)
      for (int c = 0; c < C; ++c) scatter_tap(acc_img + ((long)b * C + c) * plane, W, t, go_img[((long)b * C + c) * plane + p], e_img);
    if (acc_ref) {
      const float d = corners_value(tap_fetch(ref_depth + (long)b * plane, W, t), t);
      if (d >= kDepthMin) scatter_tap(acc_ref + (long)b * plane, W, t, go_proj[idx], e_ref);
    }
  }
}

// ---- calculate_rigid_flow ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
rigid_flow_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ Kinv, const float* __restrict__ P, int H, int W,
                      float* __restrict__ out) {
  __shared__ float sK[9], sP[12];
  const int b = blockIdx.y;
  load_mats(Kinv, P, b, sK, sP);
  const long plane = (long)H * W;
  for (long p = blockIdx.x * (long)kRedThreads + threadIdx.x; p < plane; p += (long)gridDim.x * kRedThreads) {
    const int i = (int)(p / W), j = (int)(p % W);
    const Projected r = project_pixel(sK, sP, depth[(long)b * plane + p], j, i);
    out[((long)b * 2) * plane + p] = sub_rn(r.u, (float)j);
    out[((long)b * 2 + 1) * plane + p] = sub_rn(r.v, (float)i);
  }
}

struct RigidFlowBwdPixel {
  const float *depth, *Kinv, *P, *go;
  float* g_depth;
  int H, W;
  __device__ void operator()(int b, long p, float* acc) const {
    const long plane = (long)H * W;
    const int i = (int)(p / W), j = (int)(p % W);
    const float* Pb = P + b * 12;
    const Projected r = project_pixel(Kinv + b * 9, Pb, depth[(long)b * plane + p], j, i);
    const float gD = project_backward(r, Pb, go[((long)b * 2) * plane + p], go[((long)b * 2 + 1) * plane + p], 0.f, acc);
    if (g_depth) g_depth[(long)b * plane + p] = gD;
  }
};

// ---- epipolar distance map ---------------------------------------------------------------------------------
struct EpipolarFwd {
  const float *flow, *F;
  float* out;
  int H, W;
  __device__ void operator()(long idx) const {
    const long plane = (long)H * W;
    const long p = idx % plane;
    const int b = (int)(idx / plane);
    const Epi e = epipolar_pixel(F + b * 9, flow[((long)b * 2) * plane + p], flow[((long)b * 2 + 1) * plane + p], (int)(p % W), (int)(p / W));
    out[idx] = e.dist;
  }
};

struct EpipolarBwdPixel {
  const float *flow, *F, *go;
  float* g_flow;
  int H, W;
  __device__ void operator()(int b, long p, float* acc) const {
    const long plane = (long)H * W;
    const int i = (int)(p / W), j = (int)(p % W);
    const Epi e = epipolar_pixel(F + b * 9, flow[((long)b * 2) * plane + p], flow[((long)b * 2 + 1) * plane + p], j, i);
    const float g = go[(long)b * plane + p];
    const float s = sgnf(e.n) * g / e.d;           // d dist / d n * g
    if (g_flow) {
      g_flow[((long)b * 2) * plane + p] = s * e.l[0];
      g_flow[((long)b * 2 + 1) * plane + p] = s * e.l[1];
    }
    // d dist / d l_k = sign(n) p2_k / d  -  |n| / d^2 * l_k / r  (k < 2; 0 where r == 0)
    const float t = e.r > 0.f ? -g * fabsf(e.n) / (e.d * e.d) / e.r : 0.f;
    const float gl[3] = {s * e.p2[0] + t * e.l[0], s * e.p2[1] + t * e.l[1], s * e.p2[2]};
    const float p1[3] = {(float)j, (float)i, 1.0f};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[k * 3 + c] += gl[k] * p1[c];
  }
};

}  // namespace ugl

using namespace ugl;

#define UGL_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return fail(code, __VA_ARGS__); \
  } while (0)

extern "C" int ugl_reproject_forward(const float* img, const float* depth, const float* ref_depth, const float* Kinv, const float* P,
                                     int32_t B, int32_t C, int32_t H, int32_t W, float* out, float* valid, float* proj_depth,
                                     float* comp_depth, void* stream) {
  UGL_REQUIRE(img && depth && Kinv && P && out, UGL_EINVAL, "reproject_forward: null pointer");
  UGL_REQUIRE(!proj_depth || ref_depth, UGL_EINVAL, "reproject_forward: proj_depth requested without ref_depth");
  UGL_REQUIRE(B > 0 && C > 0 && H > 1 && W > 1, UGL_EINVAL, "reproject_forward: bad shape");
  reproject_fwd_kernel<<<dim3(reduce_chunks((long)H * W), B), kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      img, depth, ref_depth, Kinv, P, C, H, W, out, valid, proj_depth, comp_depth);
  return check_launch("reproject_fwd_kernel");
}

extern "C" uint64_t ugl_reproject_backward_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W, int32_t need_grad_img,
                                                           int32_t need_grad_ref) {
  uint64_t n = reduce_workspace_bytes(B, (long)H * W, 12) + 256;
  if (need_grad_img) n += (uint64_t)B * C * H * W * 8;
  if (need_grad_ref) n += (uint64_t)B * H * W * 8;
  return n;
}

extern "C" int ugl_reproject_backward(const float* img, const float* depth, const float* ref_depth, const float* Kinv, const float* P,
                                      const float* go_img, const float* go_proj, const float* go_comp, int32_t B, int32_t C, int32_t H,
                                      int32_t W, float* grad_depth, float* grad_P, float* grad_img, float* grad_ref_depth,
                                      void* ws, uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(img && depth && Kinv && P && grad_P, UGL_EINVAL, "reproject_backward: null pointer");
  UGL_REQUIRE(!go_proj || ref_depth, UGL_EINVAL, "reproject_backward: go_proj without ref_depth");
  UGL_REQUIRE(ws && ws_bytes >= ugl_reproject_backward_workspace_bytes(B, C, H, W, grad_img != nullptr, grad_ref_depth != nullptr),
              UGL_EWORKSPACE, "reproject_backward: workspace too small");
  UGL_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7u) == 0, UGL_EALIGN, "reproject_backward: workspace not 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long plane = (long)H * W;
  // layout: [fixed-point planes][2 x maxbits, padded to 256 B][reduction partials]
  unsigned long long* acc_img = nullptr;
  unsigned long long* acc_ref = nullptr;
  char* cur = static_cast<char*>(ws);
  if (grad_img) { acc_img = reinterpret_cast<unsigned long long*>(cur); cur += (uint64_t)B * C * plane * 8; }
  if (grad_ref_depth) { acc_ref = reinterpret_cast<unsigned long long*>(cur); cur += (uint64_t)B * plane * 8; }
  unsigned* maxbits = reinterpret_cast<unsigned*>(cur);
  cur += 256;
  ReprojectBwdPixel px{img, depth, ref_depth, Kinv, P, go_img, go_proj, go_comp, grad_depth, C, H, W};
  MatFinal fin{grad_P, 12};
  int rc = launch_sample_reduce<12>(px, fin, B, plane, cur, ws_bytes - (uint64_t)(cur - static_cast<char*>(ws)), st, "reproject_backward");
  if (rc) return rc;
  const bool s_img = grad_img && go_img, s_ref = grad_ref_depth && go_proj;
  if (grad_img || grad_ref_depth) {
    cudaError_t e = cudaMemsetAsync(ws, 0, (uint64_t)(reinterpret_cast<char*>(maxbits) - static_cast<char*>(ws)) + 256, st);
    if (e != cudaSuccess) return fail((int)e, "reproject_backward: memset: %s", cudaGetErrorString(e));
    if (s_img) absmax_kernel<<<scatter_grid((long)B * C * plane), 256, 0, st>>>(go_img, (long)B * C * plane, maxbits);
    if (s_ref) absmax_kernel<<<scatter_grid((long)B * plane), 256, 0, st>>>(go_proj, (long)B * plane, maxbits + 1);
    if (s_img || s_ref) {
      reproject_scatter_kernel<<<scatter_grid((long)B * plane), 256, 0, st>>>(depth, ref_depth, Kinv, P, go_img, go_proj, B, C, H, W,
                                                                           maxbits, s_img ? acc_img : nullptr, s_ref ? acc_ref : nullptr);
      if ((rc = check_launch("reproject_scatter_kernel"))) return rc;
    }
    if (grad_img) fixed_to_float_kernel<<<scatter_grid((long)B * C * plane), 256, 0, st>>>(acc_img, (long)B * C * plane, plane, maxbits, grad_img);
    if (grad_ref_depth) fixed_to_float_kernel<<<scatter_grid((long)B * plane), 256, 0, st>>>(acc_ref, (long)B * plane, plane, maxbits + 1, grad_ref_depth);
    if ((rc = check_launch("reproject_backward(scatter)"))) return rc;
  }
  return UGL_OK;
}

extern "C" int ugl_rigid_flow_forward(const float* depth, const float* Kinv, const float* P, int32_t B, int32_t H, int32_t W, float* out,
                                      void* stream) {
  UGL_REQUIRE(depth && Kinv && P && out, UGL_EINVAL, "rigid_flow_forward: null pointer");
  rigid_flow_fwd_kernel<<<dim3(reduce_chunks((long)H * W), B), kRedThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, Kinv, P, H, W, out);
  return check_launch("rigid_flow_fwd_kernel");
}

extern "C" int ugl_rigid_flow_backward(const float* depth, const float* Kinv, const float* P, const float* grad_out, int32_t B, int32_t H,
                                       int32_t W, float* grad_depth, float* grad_P, void* ws, uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(depth && Kinv && P && grad_out && grad_P, UGL_EINVAL, "rigid_flow_backward: null pointer");
  RigidFlowBwdPixel px{depth, Kinv, P, grad_out, grad_depth, H, W};
  MatFinal fin{grad_P, 12};
  return launch_sample_reduce<12>(px, fin, B, (long)H * W, ws, ws_bytes, static_cast<cudaStream_t>(stream), "rigid_flow_backward");
}

extern "C" int ugl_epipolar_forward(const float* flow, const float* F, int32_t B, int32_t H, int32_t W, float* out, void* stream) {
  UGL_REQUIRE(flow && F && out, UGL_EINVAL, "epipolar_forward: null pointer");
  EpipolarFwd f{flow, F, out, H, W};
  return launch_pointwise(f, (long)B * H * W, static_cast<cudaStream_t>(stream), "epipolar_forward");
}

extern "C" int ugl_epipolar_backward(const float* flow, const float* F, const float* grad_out, int32_t B, int32_t H, int32_t W,
                                     float* grad_flow, float* grad_F, void* ws, uint64_t ws_bytes, void* stream) {
  UGL_REQUIRE(flow && F && grad_out && grad_F, UGL_EINVAL, "epipolar_backward: null pointer");
  EpipolarBwdPixel px{flow, F, grad_out, grad_flow, H, W};
  MatFinal fin{grad_F, 9};
  return launch_sample_reduce<9>(px, fin, B, (long)H * W, ws, ws_bytes, static_cast<cudaStream_t>(stream), "epipolar_backward");
}
