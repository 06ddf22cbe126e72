// Per-edge geometry of the bundle adjustment, on the repo's own SE3 operators (lie.cuh).
//
// An edge (i, j, k) says: patch k lives in frame i and is observed in frame j.  With G_ij = G_j * G_i^-1 the
// centre pixel (u, v, inverse depth d) back-projects to the homogeneous point p = ((u-cx)/fx, (v-cy)/fy, 1, d),
// moves to P = R_ij p_xyz + d t_ij and projects to (fx X/Z + cx, fy Y/Z + cy).  Semantics that a parity-correct
// replacement of cuda_ba must keep (dpvo/fastba/ba_cuda.cu:265-333): intrinsics of frame 0 for
// every edge; Jacobians use 1/Z only when Z >= 0.2 (else 0) while the projection divides by the raw Z; an edge
// counts when its residual is below 128 px, Z > 0.2 and the projection lies within 64 px of the image (whose
// size is taken as 2cx x 2cy); the frame-i Jacobian is -Adj(G_ij)^T applied to the frame-j one.
//
// Derivation of the frame-j rows (left perturbation xi = (tau, phi), P -> P + tau*W + phi x P, W = d):
//   a = f/Z (pixel per unit X),  b = -f X / Z^2 (pixel per unit Z)
//   d(px)/d(tau) = (a W, 0, b W)        d(px)/d(phi) = (b Y, a Z - b X, -a Y)
//   d(py)/d(tau) = (0, a' W, b' W)      d(py)/d(phi) = (-a' Z + b' Y, -b' X, a' X)
// and a Z = f (the reference writes f (1 + X^2/Z^2) for a Z - b X, kept in that form so that gated edges with
// Z < 0.2, whose 1/Z is replaced by 0, produce the same values).
#pragma once
#include "lie.cuh"

namespace dpvo {

struct EdgeLin {
  float w[2], r[2], Jz[2];
  float Ji[2][6];   // Adj(G_ij)^T Jj; enters the normal equations with a minus sign
  float Jj[2][6];
};

struct EdgeCam { float fx, fy, cx, cy; };

// relative pose of an edge from the raw pose table ([tx ty tz qx qy qz qw] rows; fastba does not renormalise)
__device__ __forceinline__ lie::SE3<float> edge_relative_pose(const float* __restrict__ poses, int64_t ix, int64_t jx, bool normalise) {
  const float* pi = poses + ix * 7;
  const float* pj = poses + jx * 7;
  lie::SE3<float> Gi, Gj;
  Gi.t = lie::v3<float>(pi[0], pi[1], pi[2]);
  Gj.t = lie::v3<float>(pj[0], pj[1], pj[2]);
  if (normalise) {                                  // lietorch loads group elements normalised (so3.h:35-37)
    Gi.q = lie::q_load<float>(pi + 3);
    Gj.q = lie::q_load<float>(pj + 3);
  } else {
    Gi.q.x = pi[3]; Gi.q.y = pi[4]; Gi.q.z = pi[5]; Gi.q.w = pi[6];
    Gj.q.x = pj[3]; Gj.q.y = pj[4]; Gj.q.z = pj[5]; Gj.q.w = pj[6];
  }
  // G_j * G_i^-1 without the renormalisations of se3_mul / se3_inv: conj(q_i) is exact for unit quaternions and
  // fastba composes raw products (ba_cuda.cu:74-85)
  lie::Quat<float> qic; qic.x = -Gi.q.x; qic.y = -Gi.q.y; qic.z = -Gi.q.z; qic.w = Gi.q.w;
  lie::SE3<float> G;
  G.q.x = Gj.q.w * qic.x + Gj.q.x * qic.w + Gj.q.y * qic.z - Gj.q.z * qic.y;
  G.q.y = Gj.q.w * qic.y + Gj.q.y * qic.w + Gj.q.z * qic.x - Gj.q.x * qic.z;
  G.q.z = Gj.q.w * qic.z + Gj.q.z * qic.w + Gj.q.x * qic.y - Gj.q.y * qic.x;
  G.q.w = Gj.q.w * qic.w - Gj.q.x * qic.x - Gj.q.y * qic.y - Gj.q.z * qic.z;
  G.t = Gj.t - lie::q_rot(G.q, Gi.t);
  return G;
}

__device__ __forceinline__ void linearize_edge_at(const float* __restrict__ poses, const float* __restrict__ patches, int P,
                                                  const float* __restrict__ target, const float* __restrict__ weight,
                                                  int64_t e, int64_t ix, int64_t jx, int64_t kx, const EdgeCam& K, EdgeLin& L) {
  const int c = P / 2;
  const float* pk = patches + kx * 3 * P * P + c * P + c;
  const lie::SE3<float> G = edge_relative_pose(poses, ix, jx, false);
  const float W = pk[2 * P * P];
  const lie::V3<float> p = lie::v3<float>((pk[0] - K.cx) / K.fx, (pk[P * P] - K.cy) / K.fy, 1.0f);
  const lie::V3<float> Pj = lie::q_rot(G.q, p) + W * G.t;
  const float X = Pj.x, Y = Pj.y, Z = Pj.z;
  const float d = ((double)Z >= 0.2) ? 1.0f / Z : 0.0f;
  const float d2 = d * d;
  const float x1 = K.fx * (X / Z) + K.cx;
  const float y1 = K.fy * (Y / Z) + K.cy;
  const float rx = target[e * 2 + 0] - x1;
  const float ry = target[e * 2 + 1] - y1;
  const bool counted = (sqrtf(rx * rx + ry * ry) < 128.0f) && ((double)Z > 0.2) && (x1 > -64.0f) && (y1 > -64.0f) &&
                       (x1 < 2.0f * K.cx + 64.0f) && (y1 < 2.0f * K.cy + 64.0f);
  const float gate = counted ? 1.0f : 0.0f;
  L.r[0] = rx; L.r[1] = ry;
  L.w[0] = gate * weight[e * 2 + 0];
  L.w[1] = gate * weight[e * 2 + 1];
  const float ax = K.fx * d, ay = K.fy * d;               // pixel per unit X / Y
  const float bx = -K.fx * X * d2, by = -K.fy * Y * d2;   // pixel per unit Z
  L.Jz[0] = ax * G.t.x + bx * G.t.z;
  L.Jz[1] = ay * G.t.y + by * G.t.z;
  L.Jj[0][0] = ax * W; L.Jj[0][1] = 0.0f;  L.Jj[0][2] = bx * W;
  L.Jj[0][3] = bx * Y; L.Jj[0][4] = K.fx * (1.0f + X * X * d2); L.Jj[0][5] = -ax * Y;
  L.Jj[1][0] = 0.0f;   L.Jj[1][1] = ay * W; L.Jj[1][2] = by * W;
  L.Jj[1][3] = K.fy * (-1.0f - Y * Y * d2); L.Jj[1][4] = -by * X; L.Jj[1][5] = ay * X;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lie::Tan6<float> a;
    a.tau = lie::v3<float>(L.Jj[r][0], L.Jj[r][1], L.Jj[r][2]);
    a.phi = lie::v3<float>(L.Jj[r][3], L.Jj[r][4], L.Jj[r][5]);
    const lie::Tan6<float> b = lie::se3_adjT(G, a);
    L.Ji[r][0] = b.tau.x; L.Ji[r][1] = b.tau.y; L.Ji[r][2] = b.tau.z;
    L.Ji[r][3] = b.phi.x; L.Ji[r][4] = b.phi.y; L.Ji[r][5] = b.phi.z;
  }
}

// pose <- Exp(xi) * pose, xi = (tau, phi)  (left update, ba_cuda.cu:156-206)
__device__ __forceinline__ void retract_pose(float* __restrict__ pose, const float* __restrict__ xi) {
  lie::SE3<float> X;
  X.t = lie::v3<float>(pose[0], pose[1], pose[2]);
  X.q.x = pose[3]; X.q.y = pose[4]; X.q.z = pose[5]; X.q.w = pose[6];
  lie::Tan6<float> a;
  a.tau = lie::v3<float>(xi[0], xi[1], xi[2]);
  a.phi = lie::v3<float>(xi[3], xi[4], xi[5]);
  const lie::SE3<float> Y = lie::se3_mul(lie::se3_exp(a), X);
  lie::se3_store(Y, pose);
}

}  // namespace dpvo
