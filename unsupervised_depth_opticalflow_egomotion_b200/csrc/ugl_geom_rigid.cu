// Level-0 rigid-consistency terms of the geom mode in one forward and one backward kernel:
//   loss_depth_flow_consis = P(|rigid_flow(disp, pose) - flow| (2 ch), valid*occ*dyn)  for both directions
//                            calculate_rigid_flow + compute_dynamic_mask's flow_diff + compute_depth_flow_consis_loss,
//                            model_geometry.py:685-732, 921-926
//   loss_epipolar          = mean(dist_bwd) + mean(dist_fwd)   compute_epipolar_map + compute_epipolar_loss (:355-418; the
//                            masked value is overwritten by the plain mean, :415-416)
// The masks come packed from ugl_geom_flow_forward_grad.  Replaces ~40 launches per step (2 rigid flows, 2 dynamic-mask maps,
// 4 mask unpacks, 4 masked means with their finalizes and element-wise backwards, 2 epipolar maps, 2 + 2 matrix-gradient
// reductions) with 4.  Gradients: both level-0 flows, the centre disparity, P (both poses), F (both poses).
#include "ugl_common.cuh"
#include "ugl_geometry.cuh"
#include "ugl_reduce.cuh"

namespace ugl {

struct RigidTermsParams {
  int B, H, W, chunks;
  const float* flow[2];          // (B,2,H,W): 0 = bwd (centre->left), 1 = fwd
  const float* disp;             // (B,1,H,W)
  const unsigned char* mask;     // (B,H,W)
  unsigned need[2];
  const float *Kinv, *P[2], *F[2];
  float* partials;               // fwd [B][chunks][6]; bwd [B][chunks][42]
  float *loss_dfc, *loss_epi, *den;          // (B,), (B,), (B,2)
  const float *g_dfc, *g_epi;    // (B,)
  float* gflow[2];               // (B,2,H,W)
  float* gdisp;                  // (B,1,H,W)
  float *gP[2], *gF[2];          // (B,3,4), (B,3,3)
};

__device__ __forceinline__ void rigid_load_mats(const RigidTermsParams& p, int b, float* sm /* 9 + 24 + 18 */) {
  const int t = threadIdx.x;
  if (t < 9) sm[t] = p.Kinv[b * 9 + t];
  else if (t < 21) sm[t] = p.P[0][b * 12 + t - 9];
  else if (t < 33) sm[t] = p.P[1][b * 12 + t - 21];
  else if (t < 42) sm[t] = p.F[0][b * 9 + t - 33];
  else if (t < 51) sm[t] = p.F[1][b * 9 + t - 42];
  __syncthreads();
}

// the coalesced loads of one pixel and direction, issued one loop iteration ahead of their use
struct RigidDirect { float D, u, v; unsigned bits; };
__device__ __forceinline__ RigidDirect rigid_direct(const RigidTermsParams& p, int b, int d, long plane, long px) {
  RigidDirect r;
  r.D = p.disp[(long)b * plane + px];
  r.bits = p.mask[(long)b * plane + px];
  r.u = p.flow[d][((long)b * 2) * plane + px];
  r.v = p.flow[d][((long)b * 2 + 1) * plane + px];
  return r;
}
// contiguous chunk of the sample's pixels for this CTA (a multiple of the CTA size), walked front to back
__device__ __forceinline__ void rigid_chunk(long plane, long& begin, long& end) {
  const long per = (((plane + gridDim.x - 1) / gridDim.x) + kRedThreads - 1) / kRedThreads * kRedThreads;
  begin = blockIdx.x * per;
  end = begin + per < plane ? begin + per : plane;
}

// grid (chunks, B, 2): one direction per CTA (the kernels are latency-bound: half the live state, twice the CTAs in flight)
__global__ void __launch_bounds__(kRedThreads, 3) rigid_terms_fwd_kernel(const __grid_constant__ RigidTermsParams p) {
  __shared__ float sm[51];
  __shared__ float red[(kRedThreads / 32) * 3];
  const int b = blockIdx.y, d = blockIdx.z;
  rigid_load_mats(p, b, sm);
  const long plane = (long)p.H * p.W;
  float acc[3] = {0.f, 0.f, 0.f};
  long px, end;
  rigid_chunk(plane, px, end);
  px += threadIdx.x;
  RigidDirect nxt = {};
  if (px < end) nxt = rigid_direct(p, b, d, plane, px);
#pragma unroll 1
  for (; px < end; px += kRedThreads) {
    const RigidDirect cur = nxt;
    if (px + kRedThreads < end) nxt = rigid_direct(p, b, d, plane, px + kRedThreads);
    const int i = (int)(px / p.W), j = (int)(px % p.W);
    const float D = cur.D, u = cur.u, v = cur.v;
    const unsigned bits = cur.bits;
    const Projected r = project_pixel(sm, sm + 9 + 12 * d, D, j, i);
    const float du = fabsf(sub_rn(sub_rn(r.u, (float)j), u)), dv = fabsf(sub_rn(sub_rn(r.v, (float)i), v));
    const float m = (bits & p.need[d]) == p.need[d] ? 1.f : 0.f;
    acc[0] += (du + dv) * m;
    acc[1] += m;
    acc[2] += epipolar_pixel(sm + 33 + 9 * d, u, v, j, i).dist;
  }
  const float v = block_reduce_n<kRedThreads, 3>(acc, red);
  // partial row layout [dfc_b, m_b, dfc_f, m_f, epi_b, epi_f]
  if (threadIdx.x < 3) p.partials[((long)b * gridDim.x + blockIdx.x) * 6 + (threadIdx.x < 2 ? 2 * d + threadIdx.x : 4 + d)] = v;
}

struct RigidTermsFinal {
  float *dfc, *epi, *den;
  float hw;
  __device__ void operator()(int b, const double* S) const {
    const float db = (float)(S[1] / hw) + 1e-12f, df = (float)(S[3] / hw) + 1e-12f;
    den[b * 2] = db; den[b * 2 + 1] = df;
    dfc[b] = (float)(S[0] / (2.0 * hw)) / db + (float)(S[2] / (2.0 * hw)) / df;
    epi[b] = (float)(S[4] / hw) + (float)(S[5] / hw);
  }
};

// grid (chunks, B, 2), one direction per CTA; grad_disp gets its two contributions by atomicAdd onto a zeroed map (order-free: two addends)
__global__ void __launch_bounds__(kRedThreads, 2) rigid_terms_bwd_kernel(const __grid_constant__ RigidTermsParams p) {
  __shared__ float sm[51];
  __shared__ float red[(kRedThreads / 32) * 21];
  const int b = blockIdx.y, d = blockIdx.z;
  rigid_load_mats(p, b, sm);
  const long plane = (long)p.H * p.W;
  const float hw = (float)p.H * (float)p.W;
  const float gd = p.g_dfc ? p.g_dfc[b] : 0.f, ge = (p.g_epi ? p.g_epi[b] : 0.f) / hw;
  const float kd = gd / (2.0f * hw) / p.den[b * 2 + d];
  const float* Pd = sm + 9 + 12 * d;
  float acc[21];                                    // P (12), F (9) of this direction
#pragma unroll
  for (int k = 0; k < 21; ++k) acc[k] = 0.f;
  long px, end;
  rigid_chunk(plane, px, end);
  px += threadIdx.x;
  RigidDirect nxt = {};
  if (px < end) nxt = rigid_direct(p, b, d, plane, px);
#pragma unroll 1
  for (; px < end; px += kRedThreads) {
    const RigidDirect cur = nxt;
    if (px + kRedThreads < end) nxt = rigid_direct(p, b, d, plane, px + kRedThreads);
    const int i = (int)(px / p.W), j = (int)(px % p.W);
    const float D = cur.D, u = cur.u, v = cur.v;
    const unsigned bits = cur.bits;
    const Projected r = project_pixel(sm, Pd, D, j, i);
    const float m = (bits & p.need[d]) == p.need[d] ? 1.f : 0.f;
    const float k = kd * m;
    const float su = sgnf(sub_rn(sub_rn(r.u, (float)j), u)) * k, sv = sgnf(sub_rn(sub_rn(r.v, (float)i), v)) * k;   // d/d rigid flow
    const float gD = project_backward(r, Pd, su, sv, 0.f, acc);
    atomicAdd(&p.gdisp[(long)b * plane + px], gD);
    // epipolar distance
    const Epi e = epipolar_pixel(sm + 33 + 9 * d, u, v, j, i);
    const float s = sgnf(e.n) * ge / e.d;
    p.gflow[d][((long)b * 2) * plane + px] = s * e.l[0] - su;
    p.gflow[d][((long)b * 2 + 1) * plane + px] = s * e.l[1] - sv;
    const float t = e.r > 0.f ? -ge * fabsf(e.n) / (e.d * e.d) / e.r : 0.f;
    const float gl[3] = {s * e.p2[0] + t * e.l[0], s * e.p2[1] + t * e.l[1], s * e.p2[2]};
    const float p1[3] = {(float)j, (float)i, 1.0f};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[12 + a * 3 + c] += gl[a] * p1[c];
  }
  const float v = block_reduce_n<kRedThreads, 21>(acc, red);
  // partial row layout [P_bwd (12), P_fwd (12), F_bwd (9), F_fwd (9)]
  if (threadIdx.x < 21) p.partials[((long)b * gridDim.x + blockIdx.x) * 42 + (threadIdx.x < 12 ? 12 * d + threadIdx.x : 24 + 9 * d + threadIdx.x - 12)] = v;
}

struct RigidTermsBwdFinal {
  float *gP0, *gP1, *gF0, *gF1;
  __device__ void operator()(int b, const double* S) const {
    for (int k = 0; k < 12; ++k) { gP0[b * 12 + k] = (float)S[k]; gP1[b * 12 + k] = (float)S[12 + k]; }
    for (int k = 0; k < 9; ++k) { gF0[b * 9 + k] = (float)S[24 + k]; gF1[b * 9 + k] = (float)S[33 + k]; }
  }
};

static int rigid_fill(const UglGeomRigidArgs* a, bool backward, RigidTermsParams& p) {
  if (!a) return fail(UGL_EINVAL, "geom_rigid: null args");
  if (a->batch <= 0 || a->batch > 65535 || a->height < 1 || a->width < 1)
    return fail(UGL_EINVAL, "geom_rigid: bad batch/size (%d, %dx%d)", a->batch, a->height, a->width);
  p.B = a->batch; p.H = a->height; p.W = a->width;
  p.chunks = reduce_chunks((long)p.H * p.W);
  const void* in[10] = {a->flow_bwd, a->flow_fwd, a->disp, a->mask_bytes, a->Kinv, a->P_bwd, a->P_fwd, a->F_bwd, a->F_fwd, a->den};
  for (int k = 0; k < 10; ++k)
    if (!in[k]) return fail(UGL_EINVAL, "geom_rigid: null input pointer (%d)", k);
  p.flow[0] = a->flow_bwd; p.flow[1] = a->flow_fwd; p.disp = a->disp; p.mask = a->mask_bytes;
  p.need[0] = (unsigned)a->need[0]; p.need[1] = (unsigned)a->need[1];
  p.Kinv = a->Kinv; p.P[0] = a->P_bwd; p.P[1] = a->P_fwd; p.F[0] = a->F_bwd; p.F[1] = a->F_fwd;
  p.den = a->den;
  const uint64_t need = (uint64_t)p.B * p.chunks * (backward ? 42 : 6) * sizeof(float);
  if (!a->workspace || a->workspace_bytes < need)
    return fail(UGL_EWORKSPACE, "geom_rigid: workspace too small (%llu < %llu)", (unsigned long long)a->workspace_bytes, (unsigned long long)need);
  p.partials = static_cast<float*>(a->workspace);
  return UGL_OK;
}

}  // namespace ugl

using namespace ugl;

extern "C" uint64_t ugl_geom_rigid_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  return (uint64_t)B * reduce_chunks((long)H * W) * 42 * sizeof(float);
}

extern "C" int ugl_geom_rigid_forward(const UglGeomRigidArgs* a) {
  RigidTermsParams p;
  int rc = rigid_fill(a, false, p);
  if (rc) return rc;
  if (!a->loss_dfc || !a->loss_epi) return fail(UGL_EINVAL, "geom_rigid_forward: null output");
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  rigid_terms_fwd_kernel<<<dim3(p.chunks, p.B, 2), kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("rigid_terms_fwd_kernel"))) return rc;
  RigidTermsFinal fin{a->loss_dfc, a->loss_epi, a->den, (float)p.H * (float)p.W};
  sample_finalize_kernel<6><<<(p.B + 3) / 4, 128, 0, st>>>(p.partials, p.chunks, p.B, fin);
  return check_launch("rigid_terms finalize");
}

extern "C" int ugl_geom_rigid_backward(const UglGeomRigidArgs* a) {
  RigidTermsParams p;
  int rc = rigid_fill(a, true, p);
  if (rc) return rc;
  void* out[7] = {a->grad_flow_bwd, a->grad_flow_fwd, a->grad_disp, a->grad_P_bwd, a->grad_P_fwd, a->grad_F_bwd, a->grad_F_fwd};
  for (int k = 0; k < 7; ++k)
    if (!out[k]) return fail(UGL_EINVAL, "geom_rigid_backward: null gradient pointer (%d)", k);
  p.g_dfc = a->grad_dfc; p.g_epi = a->grad_epi;     // either may be null (= zero upstream gradient)
  p.gflow[0] = a->grad_flow_bwd; p.gflow[1] = a->grad_flow_fwd; p.gdisp = a->grad_disp;
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const cudaError_t e = cudaMemsetAsync(p.gdisp, 0, sizeof(float) * (size_t)p.B * p.H * p.W, st);
  if (e != cudaSuccess) return fail((int)e, "geom_rigid_backward: memset: %s", cudaGetErrorString(e));
  rigid_terms_bwd_kernel<<<dim3(p.chunks, p.B, 2), kRedThreads, 0, st>>>(p);
  if ((rc = check_launch("rigid_terms_bwd_kernel"))) return rc;
  RigidTermsBwdFinal fin{a->grad_P_bwd, a->grad_P_fwd, a->grad_F_bwd, a->grad_F_fwd};
  sample_finalize_kernel<42><<<(p.B + 3) / 4, 128, 0, st>>>(p.partials, p.chunks, p.B, fin);
  return check_launch("rigid_terms bwd finalize");
}
