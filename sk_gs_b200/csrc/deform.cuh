// deform.cuh - per-Gaussian device code shared by the stand-alone skinning kernels (fk_lbs.cu) and the fused
// deform + preprocess kernel (raster_fwd.cu): K nearest joints, skinning weights (4 modes), linear blend of the joint
// transforms, output assembly.  Both translation units are built with -fmad=false, so the two paths produce IDENTICAL
// bits (tests/test_gpu_fused_path.py asserts it) - the fused kernel is an execution plan, not a second implementation.
// Semantics: SURVEY.md App. A.2 - A.3 (reference networks/sk_gs.py:751-774, 1147-1149, 1192, 1202-1203; activations
// networks/gaussian_splatting.py:155-160).
#pragma once
#include "common.cuh"

namespace skgs {

constexpr int MAXK = 8;
constexpr int JT_FLOATS = 24;  // per joint: pos 3 | t 3 | R 9 (row-major) | d_rot 4 | d_scale 3 | aux 2

// The joint table every per-Gaussian loop reads.  Built ONCE per call by fk_table_kernel (fk_lbs.cu) into global memory
// as six consecutive arrays (SoA), copied by every CTA into shared memory as it is.
struct JointTable {
  const float* pos;   // [M][3] joint positions
  const float* t;     // [M][3] translation of the global joint transform
  const float* R;     // [M][9] rotation of the global joint transform, row-major
  const float* dq;    // [M][4] sk_d_rot
  const float* ds;    // [M][3] sk_d_scale
  const float* aux;   // [M][2] kernel modes: 1/(2 r^2), sigmoid(weight)
};

__host__ __device__ inline JointTable joint_table_view(const float* base, int M) {
  JointTable jt;
  jt.pos = base;
  jt.t = jt.pos + 3 * M;
  jt.R = jt.t + 3 * M;
  jt.dq = jt.R + 9 * M;
  jt.ds = jt.dq + 4 * M;
  jt.aux = jt.ds + 3 * M;
  return jt;
}

// cooperative copy of the table into shared memory (all threads of the CTA; caller synchronises)
__device__ __forceinline__ void load_joint_table(float* smem, const float* __restrict__ table, int M) {
  for (int k = threadIdx.x; k < JT_FLOATS * M; k += blockDim.x) smem[k] = table[k];
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int KT>
struct LbsOut {
  float dx, dy, dz;      // d_xyz
  float r0, r1, r2, r3;  // d_rot
  float s0, s1, s2;      // d_scale
  float w[KT];           // normalised skinning weights, KNN order
  int idx[KT];           // K nearest joints, ascending squared distance, ties -> lower index
};

// K nearest joints of p (brute force over the shared-memory table, K-list by insertion in registers), the skinning
// weights of the selected mode and the blend  d_xyz = sum_k w_k (R_k p + t_k) - p,  d_rot = sum w dq,  d_scale = sum w ds.
// LARGEST (sp-stage warp method 'largest', networks/sk_gs.py:850-851,805-806,811-812): the mean follows the ONE
// transform with the largest weight (first maximum, as torch.argmax), rotation / scale residuals stay blended.
template <int KT, bool LARGEST = false>
__device__ __forceinline__ void lbs_gaussian(const JointTable& jt, int M, int mode, float temperature,
                                             const float* __restrict__ sp_W_row, float px, float py, float pz,
                                             LbsOut<KT>& o) {
  float bd[KT];
  int bi[KT];
#pragma unroll
  for (int k = 0; k < KT; k++) {
    bd[k] = __int_as_float(0x7f800000);
    bi[k] = 0;
  }
  for (int a = 0; a < M; a++) {
    const float dx = px - jt.pos[3 * a], dy = py - jt.pos[3 * a + 1], dz = pz - jt.pos[3 * a + 2];
    const float d2 = (dx * dx + dy * dy) + dz * dz;
    // insertion into the ascending K-list (strict '<': ties keep the lower joint index first)
#pragma unroll
    for (int k = KT - 1; k >= 0; k--) {
      if (d2 < bd[k]) {
        if (k + 1 < KT) {
          bd[k + 1] = bd[k];
          bi[k + 1] = bi[k];
        }
        bd[k] = d2;
        bi[k] = a;
      }
    }
  }
  float w[KT];
  float wsum = 0.f;
  if (mode == SKGS_LBS_W) {
    float mx = -__int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < KT; k++) {
      w[k] = sp_W_row[bi[k]];
      mx = fmaxf(mx, w[k]);
    }
#pragma unroll
    for (int k = 0; k < KT; k++) {
      w[k] = expf(w[k] - mx);
      wsum += w[k];
    }
  } else if (mode == SKGS_LBS_DIST) {
    float mx = -__int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < KT; k++) {
      w[k] = -bd[k] / temperature;
      mx = fmaxf(mx, w[k]);
    }
#pragma unroll
    for (int k = 0; k < KT; k++) {
      w[k] = expf(w[k] - mx);
      wsum += w[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < KT; k++) {
      w[k] = expf(-bd[k] * jt.aux[2 * bi[k]]) * jt.aux[2 * bi[k] + 1] + 1e-7f;
      wsum += w[k];
    }
  }
  const float winv = 1.0f / wsum;
  float ox = 0.f, oy = 0.f, oz = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  int kmax = 0;
  if constexpr (LARGEST) {
    float wmax = w[0] * winv;
#pragma unroll
    for (int k = 1; k < KT; k++)
      if (w[k] * winv > wmax) {
        wmax = w[k] * winv;
        kmax = k;
      }
  }
#pragma unroll
  for (int k = 0; k < KT; k++) {
    const float wk = w[k] * winv;
    o.w[k] = wk;
    o.idx[k] = bi[k];
    const int a = bi[k];
    const float* R = jt.R + 9 * a;
    const float yx = R[0] * px + R[1] * py + R[2] * pz + jt.t[3 * a];
    const float yy = R[3] * px + R[4] * py + R[5] * pz + jt.t[3 * a + 1];
    const float yz = R[6] * px + R[7] * py + R[8] * pz + jt.t[3 * a + 2];
    if constexpr (LARGEST) {
      if (k == kmax) { ox = yx; oy = yy; oz = yz; }
    } else {
      ox += wk * yx; oy += wk * yy; oz += wk * yz;
    }
    r0 += wk * jt.dq[4 * a]; r1 += wk * jt.dq[4 * a + 1]; r2 += wk * jt.dq[4 * a + 2]; r3 += wk * jt.dq[4 * a + 3];
    s0 += wk * jt.ds[3 * a]; s1 += wk * jt.ds[3 * a + 1]; s2 += wk * jt.ds[3 * a + 2];
  }
  o.dx = ox - px; o.dy = oy - py; o.dz = oz - pz;
  o.r0 = r0; o.r1 = r1; o.r2 = r2; o.r3 = r3;
  o.s0 = s0; o.s1 = s1; o.s2 = s2;
}

// Output assembly of one Gaussian (networks/sk_gs.py:1192,1202-1203):
//   point = _xyz + d_xyz, scale = exp(_scaling) + d_scale, rotation = normalize(_rotation + d_rot) (F.normalize, eps
//   1e-12), opacity = sigmoid(_opacity)
struct Assembled {
  float px, py, pz, sx, sy, sz, qx, qy, qz, qw, opacity;
};

__device__ __forceinline__ Assembled assemble_gaussian(float x, float y, float z, float ls0, float ls1, float ls2,
                                                       float4 rot, float opacity_logit, float dx, float dy, float dz,
                                                       float4 d_rot, float ds0, float ds1, float ds2) {
  Assembled a;
  a.px = x + dx; a.py = y + dy; a.pz = z + dz;
  a.sx = expf(ls0) + ds0; a.sy = expf(ls1) + ds1; a.sz = expf(ls2) + ds2;
  float4 r = rot;
  r.x += d_rot.x; r.y += d_rot.y; r.z += d_rot.z; r.w += d_rot.w;
  const float n = fmaxf(sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w), 1e-12f);
  a.qx = r.x / n; a.qy = r.y / n; a.qz = r.z / n; a.qw = r.w / n;
  a.opacity = sigmoidf(opacity_logit);
  return a;
}

}  // namespace skgs
