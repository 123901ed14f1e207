// Per-step 3x3 / 3x4 set-up of the depth and geom modes in one launch (one thread per sample and pose):
//   K_s      rows 0-1 of K divided by the level's down-scale             model_geometry.py:92-93
//   K_s^-1   adjugate / determinant                                       inverse_warp.py:284 (intrinsics.inverse())
//   [R|t]    euler2mat (R = Rx Ry Rz) + translation                       inverse_warp.py:110-145, 172-187
//   P_s      K_s [R|t]                                                    inverse_warp.py:289
//   F        K^-T [t]x R K^-1                                             inverse_warp.py:354-364, model_geometry.py:355-370
// and the analytic backward  (grad P_s, grad F) -> grad pose.  The reference runs these as ~100 tiny torch ops per step
// (and as many again in autograd's backward); the values are the same, formed with the same operation order
// (products and sums individually rounded, matrix products accumulated over k = 0,1,2).
#include "ugl_common.cuh"
#include "ugl_host.cuh"

namespace ugl {

struct PoseSetupParams {
  const float *pose, *K, *Kinv_full;   // (B,n,6), (B,3,3), (B,3,3)
  int B, n, S;
  float down[kMaxLevels];
  float* Kinv_out[kMaxLevels];          // (B,3,3) per level
  float* P_out[2 * kMaxLevels];         // (B,3,4), index k*S + s
  float* F_out[2];                      // (B,3,3) per pose or null
  const float* gP[2 * kMaxLevels];      // backward: may be null (= zero)
  const float* gF[2];
  float* gpose;                         // (B,n,6)
};

// C (3 x N) = A (3x3) B (3 x N), accumulated over k in order (the order of a plain GEMM inner loop)
template <int N>
__device__ __forceinline__ void mm3(const float* A, const float* Bm, float* Cm) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c)
      Cm[r * N + c] = fma_rn(A[r * 3 + 2], Bm[2 * N + c], fma_rn(A[r * 3 + 1], Bm[1 * N + c], mul_rn(A[r * 3 + 0], Bm[0 * N + c])));
}
// C = A^T B
template <int N>
__device__ __forceinline__ void mm3_tn(const float* A, const float* Bm, float* Cm) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c)
      Cm[r * N + c] = fma_rn(A[2 * 3 + r], Bm[2 * N + c], fma_rn(A[1 * 3 + r], Bm[1 * N + c], mul_rn(A[0 * 3 + r], Bm[0 * N + c])));
}
// C = A B^T (3x3)
__device__ __forceinline__ void mm3_nt(const float* A, const float* Bm, float* Cm) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      Cm[r * 3 + c] = fma_rn(A[r * 3 + 2], Bm[c * 3 + 2], fma_rn(A[r * 3 + 1], Bm[c * 3 + 1], mul_rn(A[r * 3 + 0], Bm[c * 3 + 0])));
}

struct Rot { float Rx[9], Ry[9], Rz[9], Ryz[9], R[9], cx, sx, cy, sy, cz, sz; };
__device__ __forceinline__ void make_rotation(const float* v /* rx, ry, rz */, Rot& o) {
  o.cx = cosf(v[0]); o.sx = sinf(v[0]); o.cy = cosf(v[1]); o.sy = sinf(v[1]); o.cz = cosf(v[2]); o.sz = sinf(v[2]);
  const float Rz[9] = {o.cz, -o.sz, 0.f, o.sz, o.cz, 0.f, 0.f, 0.f, 1.f};
  const float Ry[9] = {o.cy, 0.f, o.sy, 0.f, 1.f, 0.f, -o.sy, 0.f, o.cy};
  const float Rx[9] = {1.f, 0.f, 0.f, 0.f, o.cx, -o.sx, 0.f, o.sx, o.cx};
#pragma unroll
  for (int k = 0; k < 9; ++k) { o.Rx[k] = Rx[k]; o.Ry[k] = Ry[k]; o.Rz[k] = Rz[k]; }
  float Rxy[9];
  mm3<3>(o.Rx, o.Ry, Rxy);          // (Rx @ Ry) @ Rz, left to right like the reference expression
  mm3<3>(Rxy, o.Rz, o.R);
  mm3<3>(o.Ry, o.Rz, o.Ryz);        // backward only
}

__device__ __forceinline__ void scaled_K(const float* K, float down, float* Ks) {
#pragma unroll
  for (int k = 0; k < 6; ++k) Ks[k] = div_rn(K[k], down);
#pragma unroll
  for (int k = 6; k < 9; ++k) Ks[k] = K[k];
}

__device__ __forceinline__ void inverse3(const float* M, float* out) {
  const float a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
  const float A = sub_rn(mul_rn(e, i), mul_rn(f, h)), Bc = sub_rn(mul_rn(f, g), mul_rn(d, i)), Cc = sub_rn(mul_rn(d, h), mul_rn(e, g));
  const float det = add_rn(add_rn(mul_rn(a, A), mul_rn(b, Bc)), mul_rn(c, Cc));
  const float adj[9] = {A, sub_rn(mul_rn(c, h), mul_rn(b, i)), sub_rn(mul_rn(b, f), mul_rn(c, e)),
                        Bc, sub_rn(mul_rn(a, i), mul_rn(c, g)), sub_rn(mul_rn(c, d), mul_rn(a, f)),
                        Cc, sub_rn(mul_rn(b, g), mul_rn(a, h)), sub_rn(mul_rn(a, e), mul_rn(b, d))};
#pragma unroll
  for (int k = 0; k < 9; ++k) out[k] = div_rn(adj[k], det);
}

__global__ void pose_setup_fwd_kernel(const __grid_constant__ PoseSetupParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.B * p.n) return;
  const int b = t / p.n, k = t - b * p.n;
  const float* v = p.pose + (long)t * 6;
  const float* K = p.K + b * 9;
  Rot r;
  make_rotation(v + 3, r);
  const float Rt[12] = {r.R[0], r.R[1], r.R[2], v[0], r.R[3], r.R[4], r.R[5], v[1], r.R[6], r.R[7], r.R[8], v[2]};
  for (int s = 0; s < p.S; ++s) {
    float Ks[9], Ps[12];
    scaled_K(K, p.down[s], Ks);
    if (k == 0) {
      float Ki[9];
      inverse3(Ks, Ki);
      for (int q = 0; q < 9; ++q) p.Kinv_out[s][b * 9 + q] = Ki[q];
    }
    mm3<4>(Ks, Rt, Ps);
    for (int q = 0; q < 12; ++q) p.P_out[k * p.S + s][b * 12 + q] = Ps[q];
  }
  if (p.F_out[k]) {
    const float* Ki = p.Kinv_full + b * 9;
    const float T[9] = {0.f, -v[2], v[1], v[2], 0.f, -v[0], -v[1], v[0], 0.f};
    float E[9], EK[9], F[9];
    mm3<3>(T, r.R, E);
    mm3<3>(E, Ki, EK);
    mm3_tn<3>(Ki, EK, F);
    for (int q = 0; q < 9; ++q) p.F_out[k][b * 9 + q] = F[q];
  }
}

__global__ void pose_setup_bwd_kernel(const __grid_constant__ PoseSetupParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.B * p.n) return;
  const int b = t / p.n, k = t - b * p.n;
  const float* v = p.pose + (long)t * 6;
  const float* K = p.K + b * 9;
  Rot r;
  make_rotation(v + 3, r);
  float gR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gt[3] = {0.f, 0.f, 0.f};
  for (int s = 0; s < p.S; ++s) {
    const float* g = p.gP[k * p.S + s];
    if (!g) continue;
    float Ks[9], gRt[12];
    scaled_K(K, p.down[s], Ks);
    mm3_tn<4>(Ks, g + b * 12, gRt);            // d/d[R|t] of K_s [R|t]
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gR[rr * 3 + c] += gRt[rr * 4 + c];
      gt[rr] += gRt[rr * 4 + 3];
    }
  }
  if (p.gF[k]) {
    const float* Ki = p.Kinv_full + b * 9;
    const float T[9] = {0.f, -v[2], v[1], v[2], 0.f, -v[0], -v[1], v[0], 0.f};
    float tmp[9], gE[9], gT[9], gR2[9];
    mm3<3>(Ki, p.gF[k] + b * 9, tmp);          // F = Ki^T E Ki  ->  gE = Ki gF Ki^T
    mm3_nt(tmp, Ki, gE);
    mm3_nt(gE, r.R, gT);                        // E = T R        ->  gT = gE R^T, gR += T^T gE
    mm3_tn<3>(T, gE, gR2);
#pragma unroll
    for (int q = 0; q < 9; ++q) gR[q] += gR2[q];
    gt[0] += gT[7] - gT[5];
    gt[1] += gT[2] - gT[6];
    gt[2] += gT[3] - gT[1];
  }
  // R = Rx (Ry Rz)
  float gRx[9], gRyz[9], gRy[9], gRz[9];
  mm3_nt(gR, r.Ryz, gRx);
  mm3_tn<3>(r.Rx, gR, gRyz);
  mm3_nt(gRyz, r.Rz, gRy);
  mm3_tn<3>(r.Ry, gRyz, gRz);
  const float gcx = gRx[4] + gRx[8], gsx = gRx[7] - gRx[5];
  const float gcy = gRy[0] + gRy[8], gsy = gRy[2] - gRy[6];
  const float gcz = gRz[0] + gRz[4], gsz = gRz[3] - gRz[1];
  float* o = p.gpose + (long)t * 6;
  o[0] = gt[0]; o[1] = gt[1]; o[2] = gt[2];
  o[3] = r.cx * gsx - r.sx * gcx;
  o[4] = r.cy * gsy - r.sy * gcy;
  o[5] = r.cz * gsz - r.sz * gcz;
}

static int pose_fill(const float* pose, const float* K, const float* Kinv, const float* down, int B, int n, int S, PoseSetupParams& p) {
  if (!pose || !K || !down) return fail(UGL_EINVAL, "pose_setup: null pointer");
  if (B <= 0 || n < 1 || n > 2 || S < 1 || S > UGL_MAX_LEVELS) return fail(UGL_EINVAL, "pose_setup: bad batch/poses/levels (%d/%d/%d)", B, n, S);
  p.pose = pose; p.K = K; p.Kinv_full = Kinv; p.B = B; p.n = n; p.S = S;
  for (int s = 0; s < S; ++s) {
    if (!(down[s] > 0.f)) return fail(UGL_EINVAL, "pose_setup: downscale %d must be positive", s);
    p.down[s] = down[s];
  }
  return UGL_OK;
}

}  // namespace ugl

using namespace ugl;

extern "C" int ugl_pose_setup_forward(const float* pose, const float* K, const float* K_inv, const float* downscales, int32_t B, int32_t n,
                                      int32_t S, float* const* Kinv_out, float* const* P_out, float* const* F_out, void* stream) {
  PoseSetupParams p;
  int rc = pose_fill(pose, K, K_inv, downscales, B, n, S, p);
  if (rc) return rc;
  if (!Kinv_out || !P_out) return fail(UGL_EINVAL, "pose_setup_forward: null output arrays");
  for (int s = 0; s < S; ++s) {
    if (!Kinv_out[s]) return fail(UGL_EINVAL, "pose_setup_forward: null Kinv_out[%d]", s);
    p.Kinv_out[s] = Kinv_out[s];
    for (int k = 0; k < n; ++k) {
      if (!P_out[k * S + s]) return fail(UGL_EINVAL, "pose_setup_forward: null P_out[%d]", k * S + s);
      p.P_out[k * S + s] = P_out[k * S + s];
    }
  }
  for (int k = 0; k < 2; ++k) p.F_out[k] = (F_out && k < n) ? F_out[k] : nullptr;
  if ((p.F_out[0] || p.F_out[1]) && !K_inv) return fail(UGL_EINVAL, "pose_setup_forward: F needs K_inv");
  const int threads = 64, total = B * n;
  pose_setup_fwd_kernel<<<(total + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("pose_setup_fwd_kernel");
}

extern "C" int ugl_pose_setup_backward(const float* pose, const float* K, const float* K_inv, const float* downscales, int32_t B, int32_t n,
                                       int32_t S, const float* const* grad_P, const float* const* grad_F, float* grad_pose, void* stream) {
  PoseSetupParams p;
  int rc = pose_fill(pose, K, K_inv, downscales, B, n, S, p);
  if (rc) return rc;
  if (!grad_pose) return fail(UGL_EINVAL, "pose_setup_backward: null grad_pose");
  for (int q = 0; q < n * S; ++q) p.gP[q] = grad_P ? grad_P[q] : nullptr;
  for (int k = 0; k < 2; ++k) p.gF[k] = (grad_F && k < n) ? grad_F[k] : nullptr;
  if ((p.gF[0] || p.gF[1]) && !K_inv) return fail(UGL_EINVAL, "pose_setup_backward: grad_F needs K_inv");
  p.gpose = grad_pose;
  const int threads = 64, total = B * n;
  pose_setup_bwd_kernel<<<(total + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("pose_setup_bwd_kernel");
}
