// fk_lbs.cu - skeleton forward kinematics + linear blend skinning (+ output assembly) for sm_100a.
//   fk_table_kernel   : ONE CTA per call: the M joint transforms (local transform, L pointer-jumping rounds over the
//                       binary-lifting table, global transform) in shared memory -> sk_T [M][7] and the joint table
//                       (pos | t | R | d_rot | d_scale | aux, 24 floats per joint) the per-Gaussian kernels copy into
//                       shared memory
//   lbs_fwd_kernel    : streams Gaussians: brute-force K-nearest joints in registers, skinning weights (4 modes), blend
//                       of mean / rotation / scale (deform.cuh, shared with the fused kernel of raster_fwd.cu)
//   lbs_bwd_jm_kernel : per-Gaussian gradients staged per chunk, joint-major accumulation in registers (M <= 256), the
//                       FK backward (level-synchronous sweep through the kinematic chain) in the last CTA to finish;
//                       lbs_bwd_kernel (shared-memory atomics) + fk_bwd_kernel for more joints
//   sp_table_kernel / sp_bwd_kernel : the sp-stage twin - the table comes from per-superpoint SE3 predicted by the
//                       deformation network (networks/sk_gs.py:776-856)
//   assemble_*        : the element-wise activations of networks/sk_gs.py:1192,1202-1203
// Semantics: SURVEY.md App. A.1-A.3 (reference networks/sk_gs.py:193-206,751-774,1069-1150; lietorch algebra as in
// my_ext/_C/include/lie.h:45-64,142-159,228-249).  Quaternions are (x,y,z,w).
#include "deform.cuh"

namespace skgs {

constexpr int FK_THREADS = 256;
// per-joint accumulator layout of the backward pass
constexpr int NJ = 19;  // dt[3] dq[4] d(sk_d_rot)[4] d(sk_d_scale)[3] dj(d2)[3] d(radius)[1] d(weight)[1]

struct Quat {
  float x, y, z, w;
};
struct Vec3 {
  float x, y, z;
};

__device__ __forceinline__ Quat q_normalize(Quat q) {
  const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  const float inv = 1.0f / n;
  return {q.x * inv, q.y * inv, q.z * inv, q.w * inv};
}
__device__ __forceinline__ Quat q_mul(Quat a, Quat b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
          a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__device__ __forceinline__ Quat q_conj(Quat a) { return {-a.x, -a.y, -a.z, a.w}; }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// p + w*2(v x p) + v x 2(v x p)   (lie.h:59-64)
__device__ __forceinline__ Vec3 q_rotate(Quat q, Vec3 p) {
  const Vec3 v = {q.x, q.y, q.z};
  Vec3 uv = cross(v, p);
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  const Vec3 c = cross(v, uv);
  return {p.x + q.w * uv.x + c.x, p.y + q.w * uv.y + c.y, p.z + q.w * uv.z + c.z};
}
// gradient of G . rotate(q, p) w.r.t. the (unit) quaternion entries, polynomial form
__device__ __forceinline__ Quat q_rotate_grad_q(Quat q, Vec3 p, Vec3 G) {
  const Vec3 v = {q.x, q.y, q.z};
  const Vec3 pxG = cross(p, G);
  const Vec3 vxp = cross(v, p);
  const float Gv = G.x * v.x + G.y * v.y + G.z * v.z;
  const float vp = v.x * p.x + v.y * p.y + v.z * p.z;
  const float Gp = G.x * p.x + G.y * p.y + G.z * p.z;
  Quat g;
  g.x = 2.f * q.w * pxG.x + 2.f * Gv * p.x + 2.f * vp * G.x - 4.f * Gp * v.x;
  g.y = 2.f * q.w * pxG.y + 2.f * Gv * p.y + 2.f * vp * G.y - 4.f * Gp * v.y;
  g.z = 2.f * q.w * pxG.z + 2.f * Gv * p.z + 2.f * vp * G.z - 4.f * Gp * v.z;
  g.w = 2.f * (G.x * vxp.x + G.y * vxp.y + G.z * vxp.z);
  return g;
}
// rotate by the inverse (conjugate) of a unit quaternion
__device__ __forceinline__ Vec3 q_rotate_inv(Quat q, Vec3 p) { return q_rotate(q_conj(q), p); }
// Jacobian-transpose of normalisation at u (|u| = n): (g - (g.uhat) uhat) / n
__device__ __forceinline__ Quat q_normalize_bwd(Quat uhat, float n, Quat g) {
  const float d = g.x * uhat.x + g.y * uhat.y + g.z * uhat.z + g.w * uhat.w;
  const float inv = 1.0f / n;
  return {(g.x - d * uhat.x) * inv, (g.y - d * uhat.y) * inv, (g.z - d * uhat.z) * inv, (g.w - d * uhat.w) * inv};
}
__device__ __forceinline__ Quat so3_exp(Vec3 phi) {  // lie.h:142-159
  const float theta2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
  const float theta = sqrtf(theta2);
  float imag, real;
  if (theta < 1e-6f) {
    const float theta4 = theta2 * theta2;
    imag = 0.5f - (1.0f / 48.0f) * theta2 + (1.0f / 3840.0f) * theta4;
    real = 1.0f - (1.0f / 8.0f) * theta2 + (1.0f / 384.0f) * theta4;
  } else {
    imag = sinf(0.5f * theta) / theta;
    real = cosf(0.5f * theta);
  }
  return q_normalize({imag * phi.x, imag * phi.y, imag * phi.z, real});
}

__device__ __forceinline__ Quat load_q(const float* p) { return {p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ Vec3 load_v(const float* p) { return {p[0], p[1], p[2]}; }

// local rotation of joint a: normalize(sk_r) [left-multiplied by the repose delta]
__device__ __forceinline__ Quat local_rotation(const skgs_skeleton& sk, int a) {
  Quat r = q_normalize(load_q(sk.sk_r + 4 * a));
  if (sk.sk_r_delta != nullptr) {
    const Quat d = sk.sk_r_delta_dim == 3 ? so3_exp(load_v(sk.sk_r_delta + 3 * a))
                                          : q_normalize(load_q(sk.sk_r_delta + 4 * a));
    r = q_normalize(q_mul(d, r));
  }
  return r;
}

// Build sk_T for all joints in shared memory: se[2][M][7] ping-pong.  Returns index of the buffer holding the result.
__device__ int fk_build(const skgs_skeleton& sk, float* se0, float* se1) {
  const int M = sk.M;
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    float* o = se0 + 7 * a;
    if (a == sk.root) {
      o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
      o[6] = 1.f;
    } else {
      const Quat r = local_rotation(sk, a);
      const Vec3 j = load_v(sk.joints + 3 * a);
      const Vec3 rj = q_rotate(r, {-j.x, -j.y, -j.z});
      o[0] = j.x + rj.x; o[1] = j.y + rj.y; o[2] = j.z + rj.z;
      o[3] = r.x; o[4] = r.y; o[5] = r.z; o[6] = r.w;
    }
  }
  __syncthreads();
  float* cur = se0;
  float* nxt = se1;
  for (int l = 0; l < sk.L; l++) {  // out = out[parents[:, l]] o out   (networks/sk_gs.py:199-200)
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
      const int p = sk.parents[a * sk.L + l];
      const float* A = cur + 7 * p;
      const float* B = cur + 7 * a;
      const Quat qa = q_normalize(load_q(A + 3)), qb = q_normalize(load_q(B + 3));
      const Vec3 tb = q_rotate(qa, load_v(B));
      const Quat qo = q_normalize(q_mul(qa, qb));
      float* o = nxt + 7 * a;
      o[0] = A[0] + tb.x; o[1] = A[1] + tb.y; o[2] = A[2] + tb.z;
      o[3] = qo.x; o[4] = qo.y; o[5] = qo.z; o[6] = qo.w;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  if (sk.g_tr != nullptr) {  // global_T * out  (:201-206)
    const Quat qg = q_normalize(load_q(sk.g_tr + 3));
    const Vec3 tg = load_v(sk.g_tr);
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
      const float* B = cur + 7 * a;
      const Quat qb = q_normalize(load_q(B + 3));
      const Vec3 tb = q_rotate(qg, load_v(B));
      const Quat qo = q_normalize(q_mul(qg, qb));
      float* o = nxt + 7 * a;
      o[0] = tg.x + tb.x; o[1] = tg.y + tb.y; o[2] = tg.z + tb.z;
      o[3] = qo.x; o[4] = qo.y; o[5] = qo.z; o[6] = qo.w;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  return cur == se0 ? 0 : 1;
}

__device__ __forceinline__ void quat_to_rows(Quat q, float* R) {
  const float x = q.x, y = q.y, z = q.z, w = q.w;
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
}

size_t fk_table_smem_bytes(int M) { return (size_t)M * (7 + 7) * sizeof(float); }

// Forward kinematics ONCE per call (one CTA; M <= 1024): local transforms, L pointer-jumping rounds over the
// binary-lifting table, global transform - in shared memory - then sk_T [M][7] and the joint table the per-Gaussian
// kernels read (deform.cuh: pos | t | R | d_rot | d_scale | aux, SoA) go to global memory.
__global__ void __launch_bounds__(1024)
fk_table_kernel(skgs_skeleton sk, float* __restrict__ sk_T, float* __restrict__ table) {
  extern __shared__ float fsm[];
  const int M = sk.M;
  float* se0 = fsm;
  float* se1 = se0 + 7 * M;
  pdl_wait();
  pdl_trigger();
  const int res = fk_build(sk, se0, se1);
  const float* T = res == 0 ? se0 : se1;
  float* pos = table;
  float* tt = pos + 3 * M;
  float* R = tt + 3 * M;
  float* dq = R + 9 * M;
  float* ds = dq + 4 * M;
  float* aux = ds + 3 * M;
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    const float* Ta = T + 7 * a;
    if (sk_T != nullptr)
      for (int c = 0; c < 7; c++) sk_T[7 * a + c] = Ta[c];
    for (int c = 0; c < 3; c++) pos[3 * a + c] = sk.joints[3 * a + c];
    for (int c = 0; c < 3; c++) tt[3 * a + c] = Ta[c];
    quat_to_rows(q_normalize(load_q(Ta + 3)), R + 9 * a);
    for (int c = 0; c < 4; c++) dq[4 * a + c] = sk.sk_d_rot[4 * a + c];
    for (int c = 0; c < 3; c++) ds[3 * a + c] = sk.sk_d_scale[3 * a + c];
    float a0 = 0.f, a1 = 1.f;
    if (sk.mode == SKGS_LBS_KERNEL || sk.mode == SKGS_LBS_WEIGHTED_KERNEL) {
      const float r = expf(sk.sp_radius[a]);
      a0 = 1.0f / (2.0f * r * r);
      a1 = sk.mode == SKGS_LBS_WEIGHTED_KERNEL ? sigmoidf(sk.sp_weight[a]) : 1.0f;
    }
    aux[2 * a] = a0;
    aux[2 * a + 1] = a1;
  }
}

template <int KT, bool LARGEST = false>  // KT = K as a compile-time constant (1..8): the K-list lives in registers
__global__ void __launch_bounds__(FK_THREADS)
lbs_fwd_kernel(int M, int mode, float temperature, const float* __restrict__ table, const float* __restrict__ sp_W,
               int P, const float* __restrict__ xyz, float* __restrict__ d_xyz, float* __restrict__ d_rot,
               float* __restrict__ d_scale, float* __restrict__ weights, int64_t* __restrict__ indices) {
  extern __shared__ float fsm[];
  pdl_wait();
  pdl_trigger();
  load_joint_table(fsm, table, M);
  __syncthreads();
  const JointTable jt = joint_table_view(fsm, M);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const float px = xyz[3 * (size_t)i], py = xyz[3 * (size_t)i + 1], pz = xyz[3 * (size_t)i + 2];
    LbsOut<KT> o;
    lbs_gaussian<KT, LARGEST>(jt, M, mode, temperature, sp_W ? sp_W + (size_t)i * M : nullptr, px, py, pz, o);
    d_xyz[3 * (size_t)i] = o.dx; d_xyz[3 * (size_t)i + 1] = o.dy; d_xyz[3 * (size_t)i + 2] = o.dz;
    *reinterpret_cast<float4*>(d_rot + 4 * (size_t)i) = make_float4(o.r0, o.r1, o.r2, o.r3);
    d_scale[3 * (size_t)i] = o.s0; d_scale[3 * (size_t)i + 1] = o.s1; d_scale[3 * (size_t)i + 2] = o.s2;
#pragma unroll
    for (int k = 0; k < KT; k++) {
      weights[(size_t)i * KT + k] = o.w[k];
      indices[(size_t)i * KT + k] = (int64_t)o.idx[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LBS backward: per-Gaussian gradients + per-joint accumulation
// ------------------------------------------------------------------------------------------------------------------
size_t lbs_bwd_smem_bytes(int M) { return (size_t)M * (3 + 7 + 4 + 3 + 2 + 1 + NJ) * sizeof(float); }

__global__ void __launch_bounds__(FK_THREADS)
lbs_bwd_kernel(skgs_skeleton sk, int P, const float* __restrict__ xyz, const float* __restrict__ sk_T,
               const float* __restrict__ weights, const int64_t* __restrict__ indices,
               const float* __restrict__ g_dxyz, const float* __restrict__ g_drot, const float* __restrict__ g_dscale,
               const float* __restrict__ g_w, float* __restrict__ dL_dsp_W, float* __restrict__ dL_dsp_W_knn,
               float* __restrict__ jacc /*[M][NJ]*/, int largest) {
  extern __shared__ float bsm[];
  const int M = sk.M, K = sk.K;
  float* s_pos = bsm;             // [M][3]
  float* s_T = s_pos + 3 * M;     // [M][7] (t, unit q)
  float* s_dq = s_T + 7 * M;      // [M][4]
  float* s_ds = s_dq + 4 * M;     // [M][3]
  float* s_aux = s_ds + 3 * M;    // [M][2]
  float* s_sig = s_aux + 2 * M;   // [M] sigmoid(weight)
  float* s_acc = s_sig + M;       // [M][NJ]
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    for (int c = 0; c < 3; c++) s_pos[3 * a + c] = sk.joints[3 * a + c];
    for (int c = 0; c < 3; c++) s_T[7 * a + c] = sk_T[7 * a + c];
    const Quat q = q_normalize(load_q(sk_T + 7 * a + 3));
    s_T[7 * a + 3] = q.x; s_T[7 * a + 4] = q.y; s_T[7 * a + 5] = q.z; s_T[7 * a + 6] = q.w;
    for (int c = 0; c < 4; c++) s_dq[4 * a + c] = sk.sk_d_rot[4 * a + c];
    for (int c = 0; c < 3; c++) s_ds[3 * a + c] = sk.sk_d_scale[3 * a + c];
    s_aux[2 * a] = s_aux[2 * a + 1] = 0.f;
    s_sig[a] = 1.f;
    if (sk.mode == SKGS_LBS_KERNEL || sk.mode == SKGS_LBS_WEIGHTED_KERNEL) {
      const float r = expf(sk.sp_radius[a]);
      s_aux[2 * a] = 1.0f / (2.0f * r * r);
      s_aux[2 * a + 1] = 1.0f / (r * r);
      if (sk.mode == SKGS_LBS_WEIGHTED_KERNEL) s_sig[a] = sigmoidf(sk.sp_weight[a]);
    }
  }
  for (int k = threadIdx.x; k < M * NJ; k += blockDim.x) s_acc[k] = 0.f;
  __syncthreads();

  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const Vec3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const Vec3 G = g_dxyz ? Vec3{g_dxyz[3 * i], g_dxyz[3 * i + 1], g_dxyz[3 * i + 2]} : Vec3{0.f, 0.f, 0.f};
    const float4 gr = g_drot ? *reinterpret_cast<const float4*>(g_drot + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const Vec3 gs = g_dscale ? Vec3{g_dscale[3 * i], g_dscale[3 * i + 1], g_dscale[3 * i + 2]} : Vec3{0.f, 0.f, 0.f};
    float w[MAXK], dw[MAXK];
    int idx[MAXK];
    float wdw = 0.f;
    int kmax = 0;  // warp method 'largest': the mean follows the transform of the largest weight only
#pragma unroll
    for (int k = 0; k < MAXK; k++)
      if (k < K) {
        w[k] = weights[(size_t)i * K + k];
        if (w[k] > w[kmax]) kmax = k;
      }
#pragma unroll
    for (int k = 0; k < MAXK; k++)
      if (k < K) {
        const int a = idx[k] = (int)indices[(size_t)i * K + k];
        const float* Ta = s_T + 7 * a;
        const Quat q = {Ta[3], Ta[4], Ta[5], Ta[6]};
        const Vec3 y = q_rotate(q, p);
        float d = largest ? 0.f : G.x * (y.x + Ta[0]) + G.y * (y.y + Ta[1]) + G.z * (y.z + Ta[2]);
        d += gr.x * s_dq[4 * a] + gr.y * s_dq[4 * a + 1] + gr.z * s_dq[4 * a + 2] + gr.w * s_dq[4 * a + 3];
        d += gs.x * s_ds[3 * a] + gs.y * s_ds[3 * a + 1] + gs.z * s_ds[3 * a + 2];
        if (g_w) d += g_w[(size_t)i * K + k];
        dw[k] = d;
        wdw += w[k] * d;
        // per-joint sums: dt, dq (polynomial gradient, projected later in fk_bwd), d(sk_d_rot), d(sk_d_scale)
        float* acc = s_acc + NJ * a;
        const float wa = largest ? (k == kmax ? 1.f : 0.f) : w[k];
        const Vec3 wG = {wa * G.x, wa * G.y, wa * G.z};
        const Quat gq = q_rotate_grad_q(q, p, wG);
        atomicAdd(acc + 0, wG.x); atomicAdd(acc + 1, wG.y); atomicAdd(acc + 2, wG.z);
        atomicAdd(acc + 3, gq.x); atomicAdd(acc + 4, gq.y); atomicAdd(acc + 5, gq.z); atomicAdd(acc + 6, gq.w);
        atomicAdd(acc + 7, w[k] * gr.x); atomicAdd(acc + 8, w[k] * gr.y); atomicAdd(acc + 9, w[k] * gr.z);
        atomicAdd(acc + 10, w[k] * gr.w);
        atomicAdd(acc + 11, w[k] * gs.x); atomicAdd(acc + 12, w[k] * gs.y); atomicAdd(acc + 13, w[k] * gs.z);
      }
    // ---- through the weight function
    if (sk.mode == SKGS_LBS_W) {
      if (dL_dsp_W != nullptr) {
        float* row = dL_dsp_W + (size_t)i * M;
        for (int a = 0; a < M; a++) row[a] = 0.f;
#pragma unroll
        for (int k = 0; k < MAXK; k++)
          if (k < K) row[idx[k]] = w[k] * (dw[k] - wdw);
      }
      if (dL_dsp_W_knn != nullptr) {
#pragma unroll
        for (int k = 0; k < MAXK; k++)
          if (k < K) dL_dsp_W_knn[(size_t)i * K + k] = w[k] * (dw[k] - wdw);
      }
    } else {
      // u_k = e_k * s_k + 1e-7, w = u / S.  Recover S from the largest weight to avoid cancellation.
      float d2[MAXK], e[MAXK];
      float S = 0.f;
#pragma unroll
      for (int k = 0; k < MAXK; k++)
        if (k < K) {
          const int a = idx[k];
          const float dx = p.x - s_pos[3 * a], dy = p.y - s_pos[3 * a + 1], dz = p.z - s_pos[3 * a + 2];
          d2[k] = dx * dx + dy * dy + dz * dz;
          if (sk.mode == SKGS_LBS_DIST) {
            e[k] = 0.f;
          } else {
            e[k] = expf(-d2[k] * s_aux[2 * a]);
            S += e[k] * s_sig[a] + 1e-7f;
          }
        }
#pragma unroll
      for (int k = 0; k < MAXK; k++)
        if (k < K) {
          const int a = idx[k];
          float* acc = s_acc + NJ * a;
          float dd2;  // dL / d(d2_k)
          if (sk.mode == SKGS_LBS_DIST) {
            dd2 = -(w[k] * (dw[k] - wdw)) / sk.temperature;
          } else {
            const float du = (dw[k] - wdw) / S;
            dd2 = -du * s_sig[a] * e[k] * s_aux[2 * a];
            atomicAdd(acc + 17, du * s_sig[a] * e[k] * d2[k] * s_aux[2 * a + 1]);  // d / d(log radius)
            if (sk.mode == SKGS_LBS_WEIGHTED_KERNEL)
              atomicAdd(acc + 18, du * e[k] * s_sig[a] * (1.0f - s_sig[a]));       // d / d(weight logit)
          }
          atomicAdd(acc + 14, -2.0f * dd2 * (p.x - s_pos[3 * a]));
          atomicAdd(acc + 15, -2.0f * dd2 * (p.y - s_pos[3 * a + 1]));
          atomicAdd(acc + 16, -2.0f * dd2 * (p.z - s_pos[3 * a + 2]));
        }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < M * NJ; k += blockDim.x) {
    const float v = s_acc[k];
    if (v != 0.f) atomicAdd(jacc + k, v);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// LBS backward, joint-major variant (M <= 256).  The first version added 14 values per (Gaussian, joint) pair with
// shared-memory atomics (66 us at P = 100K: the ATOMS pipe is the limit).  Here a CTA stages a chunk of 256 Gaussians
// in shared memory (phase 1, thread = Gaussian: gradient w.r.t. the K weights, weight-function backward, dL/dsp_W rows
// written coalesced), then re-maps its threads to (joint, slice) pairs (phase 2): each thread walks the Gaussians of its
// slice, and accumulates the contributions to ITS joint in 19 registers that live across all chunks of the persistent
// CTA.  No atomics until the final cross-slice / cross-CTA reduction.
// ------------------------------------------------------------------------------------------------------------------
struct FkBwdOut {   // outputs of the FK backward (any may be NULL)
  const float* dL_dsk_T_direct;
  float *dL_djoints, *dL_dsk_r, *dL_dsk_d_rot, *dL_dsk_d_scale, *dL_dg_tr, *dL_dsp_radius, *dL_dsp_weight;
};
__device__ __forceinline__ void fk_bwd_body(const skgs_skeleton& sk, const float* __restrict__ jacc,
                                            const float* __restrict__ dL_dsk_T_direct, float* __restrict__ dL_djoints,
                                            float* __restrict__ dL_dsk_r, float* __restrict__ dL_dsk_d_rot,
                                            float* __restrict__ dL_dsk_d_scale, float* __restrict__ dL_dg_tr,
                                            float* __restrict__ dL_dsp_radius, float* __restrict__ dL_dsp_weight,
                                            float* ksm);

constexpr int JM_CHUNK = 256;
constexpr int JM_MAX_M = 256;

struct JmChunk {          // SoA records of one chunk
  float p[JM_CHUNK][3];
  float G[JM_CHUNK][3];
  float gr[JM_CHUNK][4];
  float gs[JM_CHUNK][3];
  float w[JM_CHUNK][MAXK];
  float dl[JM_CHUNK][MAXK];   // mode W: dL/dlogit_k ; other modes: dL/d(d2_k)
  float e1[JM_CHUNK][MAXK];   // kernel modes: contribution to d/d(log radius) of joint idx_k
  float e2[JM_CHUNK][MAXK];   // weighted_kernel: contribution to d/d(weight logit)
  int idx[JM_CHUNK][MAXK];
  uint16_t list[JM_CHUNK * MAXK];  // (g << 3 | k) pairs grouped by joint
  uint32_t cnt[JM_MAX_M];          // pairs per joint in this chunk
  uint32_t off[JM_MAX_M];
};

size_t lbs_bwd_jm_smem_bytes(int M) { return sizeof(JmChunk) + (size_t)M * (3 + 7 + 4 + 3 + 2 + 1 + NJ) * sizeof(float); }

__global__ void __launch_bounds__(FK_THREADS)
lbs_bwd_jm_kernel(skgs_skeleton sk, int P, const float* __restrict__ xyz, const float* __restrict__ sk_T,
                  const float* __restrict__ weights, const int64_t* __restrict__ indices,
                  const float* __restrict__ g_dxyz, const float* __restrict__ g_drot,
                  const float* __restrict__ g_dscale, const float* __restrict__ g_w, float* __restrict__ dL_dsp_W,
                  float* __restrict__ dL_dsp_W_knn, float* __restrict__ jacc /*[M][NJ]*/, uint32_t* __restrict__ done,
                  FkBwdOut fk) {
  extern __shared__ __align__(16) unsigned char jm_raw[];
  pdl_wait();
  pdl_trigger();
  JmChunk& C = *reinterpret_cast<JmChunk*>(jm_raw);
  float* tab = reinterpret_cast<float*>(jm_raw + sizeof(JmChunk));
  const int M = sk.M, K = sk.K;
  float* s_pos = tab;             // [M][3]
  float* s_T = s_pos + 3 * M;     // [M][7]
  float* s_dq = s_T + 7 * M;      // [M][4]
  float* s_ds = s_dq + 4 * M;     // [M][3]
  float* s_aux = s_ds + 3 * M;    // [M][2]
  float* s_sig = s_aux + 2 * M;   // [M]
  float* s_acc = s_sig + M;       // [M][NJ]
  const int tid = threadIdx.x;
  for (int a = tid; a < M; a += blockDim.x) {
    for (int c = 0; c < 3; c++) s_pos[3 * a + c] = sk.joints[3 * a + c];
    for (int c = 0; c < 3; c++) s_T[7 * a + c] = sk_T[7 * a + c];
    const Quat q = q_normalize(load_q(sk_T + 7 * a + 3));
    s_T[7 * a + 3] = q.x; s_T[7 * a + 4] = q.y; s_T[7 * a + 5] = q.z; s_T[7 * a + 6] = q.w;
    for (int c = 0; c < 4; c++) s_dq[4 * a + c] = sk.sk_d_rot[4 * a + c];
    for (int c = 0; c < 3; c++) s_ds[3 * a + c] = sk.sk_d_scale[3 * a + c];
    s_aux[2 * a] = s_aux[2 * a + 1] = 0.f;
    s_sig[a] = 1.f;
    if (sk.mode == SKGS_LBS_KERNEL || sk.mode == SKGS_LBS_WEIGHTED_KERNEL) {
      const float r = expf(sk.sp_radius[a]);
      s_aux[2 * a] = 1.0f / (2.0f * r * r);
      s_aux[2 * a + 1] = 1.0f / (r * r);
      if (sk.mode == SKGS_LBS_WEIGHTED_KERNEL) s_sig[a] = sigmoidf(sk.sp_weight[a]);
    }
  }
  for (int k = tid; k < M * NJ; k += blockDim.x) s_acc[k] = 0.f;
  __syncthreads();
  // phase-2 role of this thread: joint `ja`, slice `js` of the chunk
  const int Mp = ((M + 31) / 32) * 32;          // joints padded to whole warps
  const int nslice = FK_THREADS / Mp;           // >= 1 because M <= 256
  const int ja = tid % Mp, js = tid / Mp;
  const bool jactive = ja < M && js < nslice;
  Quat qa = {0.f, 0.f, 0.f, 1.f};
  Vec3 pa = {0.f, 0.f, 0.f};
  if (ja < M) {
    qa = {s_T[7 * ja + 3], s_T[7 * ja + 4], s_T[7 * ja + 5], s_T[7 * ja + 6]};
    pa = {s_pos[3 * ja], s_pos[3 * ja + 1], s_pos[3 * ja + 2]};
  }
  float acc[NJ];
#pragma unroll
  for (int c = 0; c < NJ; c++) acc[c] = 0.f;

  for (int base = blockIdx.x * JM_CHUNK; base < P; base += gridDim.x * JM_CHUNK) {
    const int n = min(JM_CHUNK, P - base);
    for (int a = tid; a < M; a += blockDim.x) C.cnt[a] = 0;
    __syncthreads();
    // ---------------------------------------------------------------- phase 1: thread = Gaussian
    uint32_t rank[MAXK];
    int myidx[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; k++) { rank[k] = 0; myidx[k] = -1; }
    if (tid < n) {
      const int i = base + tid;
      const Vec3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
      const Vec3 G = g_dxyz ? Vec3{g_dxyz[3 * i], g_dxyz[3 * i + 1], g_dxyz[3 * i + 2]} : Vec3{0.f, 0.f, 0.f};
      const float4 gr = g_drot ? *reinterpret_cast<const float4*>(g_drot + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      const Vec3 gs = g_dscale ? Vec3{g_dscale[3 * i], g_dscale[3 * i + 1], g_dscale[3 * i + 2]} : Vec3{0.f, 0.f, 0.f};
      C.p[tid][0] = p.x; C.p[tid][1] = p.y; C.p[tid][2] = p.z;
      C.G[tid][0] = G.x; C.G[tid][1] = G.y; C.G[tid][2] = G.z;
      C.gr[tid][0] = gr.x; C.gr[tid][1] = gr.y; C.gr[tid][2] = gr.z; C.gr[tid][3] = gr.w;
      C.gs[tid][0] = gs.x; C.gs[tid][1] = gs.y; C.gs[tid][2] = gs.z;
      float w[MAXK], dw[MAXK];
      int idx[MAXK];
      float wdw = 0.f;
#pragma unroll
      for (int k = 0; k < MAXK; k++)
        if (k < K) {
          w[k] = weights[(size_t)i * K + k];
          const int a = idx[k] = (int)indices[(size_t)i * K + k];
          const float* Ta = s_T + 7 * a;
          const Vec3 y = q_rotate({Ta[3], Ta[4], Ta[5], Ta[6]}, p);
          float d = G.x * (y.x + Ta[0]) + G.y * (y.y + Ta[1]) + G.z * (y.z + Ta[2]);
          d += gr.x * s_dq[4 * a] + gr.y * s_dq[4 * a + 1] + gr.z * s_dq[4 * a + 2] + gr.w * s_dq[4 * a + 3];
          d += gs.x * s_ds[3 * a] + gs.y * s_ds[3 * a + 1] + gs.z * s_ds[3 * a + 2];
          if (g_w) d += g_w[(size_t)i * K + k];
          dw[k] = d;
          wdw += w[k] * d;
          C.w[tid][k] = w[k];
          C.idx[tid][k] = a;
          myidx[k] = a;
          rank[k] = atomicAdd(&C.cnt[a], 1u);
        }
      if (sk.mode == SKGS_LBS_W) {
        // dL/dsp_W is dense [P][M] with K non-zeros per row: the rows were zero-filled by a memset node, only the K
        // logit gradients are scattered here
        if (dL_dsp_W != nullptr) {
          float* row = dL_dsp_W + (size_t)i * M;
#pragma unroll
          for (int k = 0; k < MAXK; k++)
            if (k < K) row[idx[k]] = w[k] * (dw[k] - wdw);
        }
        if (dL_dsp_W_knn != nullptr) {  // compact form: the K logit gradients in KNN order
#pragma unroll
          for (int k = 0; k < MAXK; k++)
            if (k < K) dL_dsp_W_knn[(size_t)i * K + k] = w[k] * (dw[k] - wdw);
        }
      } else {
        float d2[MAXK], e[MAXK];
        float S = 0.f;
#pragma unroll
        for (int k = 0; k < MAXK; k++)
          if (k < K) {
            const int a = idx[k];
            const float dx = p.x - s_pos[3 * a], dy = p.y - s_pos[3 * a + 1], dz = p.z - s_pos[3 * a + 2];
            d2[k] = dx * dx + dy * dy + dz * dz;
            e[k] = sk.mode == SKGS_LBS_DIST ? 0.f : expf(-d2[k] * s_aux[2 * a]);
            S += e[k] * s_sig[a] + 1e-7f;
          }
#pragma unroll
        for (int k = 0; k < MAXK; k++)
          if (k < K) {
            const int a = idx[k];
            float dd2, r1 = 0.f, r2 = 0.f;
            if (sk.mode == SKGS_LBS_DIST) {
              dd2 = -(w[k] * (dw[k] - wdw)) / sk.temperature;
            } else {
              const float du = (dw[k] - wdw) / S;
              dd2 = -du * s_sig[a] * e[k] * s_aux[2 * a];
              r1 = du * s_sig[a] * e[k] * d2[k] * s_aux[2 * a + 1];
              r2 = du * e[k] * s_sig[a] * (1.0f - s_sig[a]);
            }
            C.dl[tid][k] = dd2;
            C.e1[tid][k] = r1;
            C.e2[tid][k] = r2;
          }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- group the (Gaussian, k) pairs by joint
    if (tid < 32) {  // exclusive scan of the per-joint counts by one warp (M <= 256: 8 per lane)
      uint32_t run = 0;
      for (int a0 = 0; a0 < M; a0 += 32) {
        const int a = a0 + tid;
        const uint32_t c = a < M ? C.cnt[a] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += t;
        }
        if (a < M) C.off[a] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXK; k++)
      if (myidx[k] >= 0) C.list[C.off[myidx[k]] + rank[k]] = (uint16_t)((tid << 3) | k);
    __syncthreads();
    // ---------------------------------------------------------------- phase 2: thread = (joint, slice)
    if (jactive) {  // thread (ja, js) takes every nslice-th pair of joint ja's list: all lanes do useful work
      const uint32_t beg = C.off[ja], cntj = C.cnt[ja];
      for (uint32_t e = js; e < cntj; e += nslice) {
        const uint32_t code = C.list[beg + e];
        const int g = (int)(code >> 3), mk = (int)(code & 7u);
        const float wk = C.w[g][mk];
        const Vec3 p = {C.p[g][0], C.p[g][1], C.p[g][2]};
        const Vec3 wG = {wk * C.G[g][0], wk * C.G[g][1], wk * C.G[g][2]};
        const Quat gq = q_rotate_grad_q(qa, p, wG);
        acc[0] += wG.x; acc[1] += wG.y; acc[2] += wG.z;
        acc[3] += gq.x; acc[4] += gq.y; acc[5] += gq.z; acc[6] += gq.w;
        acc[7] += wk * C.gr[g][0]; acc[8] += wk * C.gr[g][1]; acc[9] += wk * C.gr[g][2]; acc[10] += wk * C.gr[g][3];
        acc[11] += wk * C.gs[g][0]; acc[12] += wk * C.gs[g][1]; acc[13] += wk * C.gs[g][2];
        if (sk.mode != SKGS_LBS_W) {
          const float dd2 = C.dl[g][mk];
          acc[14] += -2.0f * dd2 * (p.x - pa.x);
          acc[15] += -2.0f * dd2 * (p.y - pa.y);
          acc[16] += -2.0f * dd2 * (p.z - pa.z);
          acc[17] += C.e1[g][mk];
          acc[18] += C.e2[g][mk];
        }
      }
    }
    __syncthreads();
  }
  if (jactive) {
#pragma unroll
    for (int c = 0; c < NJ; c++)
      if (acc[c] != 0.f) atomicAdd(&s_acc[NJ * ja + c], acc[c]);
  }
  __syncthreads();
  for (int k = tid; k < M * NJ; k += blockDim.x) {
    const float v = s_acc[k];
    if (v != 0.f) atomicAdd(jacc + k, v);
  }
  // ---- FK backward as the tail of this kernel: the LAST CTA to get here sees every CTA's per-joint sums and runs the
  //      level-synchronous sweep through the kinematic chain itself (no second launch, no single-CTA kernel)
  if (done == nullptr) return;  // sp-stage caller: its own per-superpoint tail kernel follows
  __shared__ uint32_t s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(done, 1u) == gridDim.x - 1u) ? 1u : 0u;
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  fk_bwd_body(sk, jacc, fk.dL_dsk_T_direct, fk.dL_djoints, fk.dL_dsk_r, fk.dL_dsk_d_rot, fk.dL_dsk_d_scale, fk.dL_dg_tr,
              fk.dL_dsp_radius, fk.dL_dsp_weight, reinterpret_cast<float*>(jm_raw));  // the chunk staging area is free
}

// ------------------------------------------------------------------------------------------------------------------
// FK backward: one CTA.  Level-synchronous re-evaluation T_root = g, T_a = T_parent o L_a and its reverse sweep.
// ------------------------------------------------------------------------------------------------------------------
size_t fk_bwd_smem_bytes(int M) { return (size_t)M * (7 + 7 + 7 + 7 + 1) * sizeof(float) + 16; }  // + s_maxdepth

// body of the FK backward, run by ONE CTA (any block size): as the kernel below, or as the tail of lbs_bwd_jm_kernel by
// the last CTA that finishes (no separate launch)
__device__ __forceinline__ void fk_bwd_body(const skgs_skeleton& sk, const float* __restrict__ jacc,
                                            const float* __restrict__ dL_dsk_T_direct, float* __restrict__ dL_djoints,
                                            float* __restrict__ dL_dsk_r, float* __restrict__ dL_dsk_d_rot,
                                            float* __restrict__ dL_dsk_d_scale, float* __restrict__ dL_dg_tr,
                                            float* __restrict__ dL_dsp_radius, float* __restrict__ dL_dsp_weight,
                                            float* ksm) {
  const int M = sk.M, L = sk.L;
  float* s_L = ksm;              // local transforms [M][7]
  float* s_T = s_L + 7 * M;      // global transforms [M][7]
  float* s_gT = s_T + 7 * M;     // gradient w.r.t. global transforms [M][7]
  float* s_gL = s_gT + 7 * M;    // gradient w.r.t. local transforms [M][7]
  int* s_depth = reinterpret_cast<int*>(s_gL + 7 * M);
  int& s_maxdepth = s_depth[M];
  if (threadIdx.x == 0) s_maxdepth = 0;
  __syncthreads();
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    // depth by walking parents[:, 0]
    int d = 0, f = a;
    while (f != sk.root && d <= M) {
      f = sk.parents[f * L];
      d++;
    }
    s_depth[a] = d;
    atomicMax(&s_maxdepth, d);
    const Quat r = local_rotation(sk, a);
    const Vec3 j = load_v(sk.joints + 3 * a);
    const Vec3 rj = q_rotate(r, {-j.x, -j.y, -j.z});
    float* o = s_L + 7 * a;
    o[0] = j.x + rj.x; o[1] = j.y + rj.y; o[2] = j.z + rj.z; o[3] = r.x; o[4] = r.y; o[5] = r.z; o[6] = r.w;
    for (int c = 0; c < 7; c++) {
      s_gT[7 * a + c] = __ldcg(jacc + NJ * a + c) + (dL_dsk_T_direct ? dL_dsk_T_direct[7 * a + c] : 0.f);
      s_gL[7 * a + c] = 0.f;
    }
  }
  __syncthreads();
  const int maxdepth = s_maxdepth;
  Quat qg = {0.f, 0.f, 0.f, 1.f};
  Vec3 tg = {0.f, 0.f, 0.f};
  float ng = 1.f;
  if (sk.g_tr != nullptr) {
    const Quat raw = load_q(sk.g_tr + 3);
    ng = sqrtf(raw.x * raw.x + raw.y * raw.y + raw.z * raw.z + raw.w * raw.w);
    qg = q_normalize(raw);
    tg = load_v(sk.g_tr);
  }
  for (int lev = 0; lev <= maxdepth; lev++) {
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
      if (s_depth[a] != lev) continue;
      float* o = s_T + 7 * a;
      if (lev == 0) {
        o[0] = tg.x; o[1] = tg.y; o[2] = tg.z; o[3] = qg.x; o[4] = qg.y; o[5] = qg.z; o[6] = qg.w;
      } else {
        const int p = sk.parents[a * L];
        const float* A = s_T + 7 * p;
        const float* B = s_L + 7 * a;
        const Quat qa = load_q(A + 3), qb = load_q(B + 3);
        const Vec3 tb = q_rotate(qa, load_v(B));
        const Quat qo = q_normalize(q_mul(qa, qb));
        o[0] = A[0] + tb.x; o[1] = A[1] + tb.y; o[2] = A[2] + tb.z; o[3] = qo.x; o[4] = qo.y; o[5] = qo.z; o[6] = qo.w;
      }
    }
    __syncthreads();
  }
  for (int lev = maxdepth; lev >= 1; lev--) {
    for (int a = threadIdx.x; a < M; a += blockDim.x) {
      if (s_depth[a] != lev) continue;
      const int p = sk.parents[a * L];
      const float* A = s_T + 7 * p;
      const float* B = s_L + 7 * a;
      const Quat qa = load_q(A + 3), qb = load_q(B + 3), qo = load_q(s_T + 7 * a + 3);
      const Vec3 gt = load_v(s_gT + 7 * a);
      // quaternion gradient arrives w.r.t. the stored unit quaternion: project onto its tangent space
      const Quat gm = q_normalize_bwd(qo, 1.0f, load_q(s_gT + 7 * a + 3));
      const Quat gqa_prod = q_mul(gm, q_conj(qb));   // d(qa qb)/dqa ^T gm
      const Quat gqb = q_mul(q_conj(qa), gm);        // d(qa qb)/dqb ^T gm
      const Quat gqa_rot = q_rotate_grad_q(qa, load_v(B), gt);
      const Vec3 glt = q_rotate_inv(qa, gt);
      float* gp = s_gT + 7 * p;
      atomicAdd(gp + 0, gt.x); atomicAdd(gp + 1, gt.y); atomicAdd(gp + 2, gt.z);
      atomicAdd(gp + 3, gqa_prod.x + gqa_rot.x); atomicAdd(gp + 4, gqa_prod.y + gqa_rot.y);
      atomicAdd(gp + 5, gqa_prod.z + gqa_rot.z); atomicAdd(gp + 6, gqa_prod.w + gqa_rot.w);
      float* gl = s_gL + 7 * a;
      gl[0] = glt.x; gl[1] = glt.y; gl[2] = glt.z; gl[3] = gqb.x; gl[4] = gqb.y; gl[5] = gqb.z; gl[6] = gqb.w;
    }
    __syncthreads();
  }
  // root -> global transform
  if (threadIdx.x == 0 && dL_dg_tr != nullptr) {
    if (sk.g_tr != nullptr) {
      const float* g = s_gT + 7 * sk.root;
      const Quat gq = q_normalize_bwd(qg, ng, load_q(g + 3));
      dL_dg_tr[0] = g[0]; dL_dg_tr[1] = g[1]; dL_dg_tr[2] = g[2];
      dL_dg_tr[3] = gq.x; dL_dg_tr[4] = gq.y; dL_dg_tr[5] = gq.z; dL_dg_tr[6] = gq.w;
    } else {
      for (int c = 0; c < 7; c++) dL_dg_tr[c] = 0.f;
    }
  }
  // local transform -> joints, sk_r ; plus the direct per-joint sums
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    const float* gl = s_gL + 7 * a;
    Vec3 gj = {__ldcg(jacc + NJ * a + 14), __ldcg(jacc + NJ * a + 15), __ldcg(jacc + NJ * a + 16)};
    Quat gr = {0.f, 0.f, 0.f, 0.f};
    if (a != sk.root) {
      const Quat r = load_q(s_L + 7 * a + 3);
      const Vec3 j = load_v(sk.joints + 3 * a);
      const Vec3 glt = load_v(gl);
      // t = j + rotate(r, -j)
      const Vec3 back = q_rotate_inv(r, glt);
      gj.x += glt.x - back.x; gj.y += glt.y - back.y; gj.z += glt.z - back.z;
      const Quat g1 = q_rotate_grad_q(r, {-j.x, -j.y, -j.z}, glt);
      Quat g = {g1.x + gl[3], g1.y + gl[4], g1.z + gl[5], g1.w + gl[6]};
      const Quat raw = load_q(sk.sk_r + 4 * a);
      const float n0 = sqrtf(raw.x * raw.x + raw.y * raw.y + raw.z * raw.z + raw.w * raw.w);
      const Quat r0 = q_normalize(raw);
      if (sk.sk_r_delta != nullptr) {
        const Quat d = sk.sk_r_delta_dim == 3 ? so3_exp(load_v(sk.sk_r_delta + 3 * a))
                                              : q_normalize(load_q(sk.sk_r_delta + 4 * a));
        g = q_normalize_bwd(r, 1.0f, g);   // through normalize(d r0), |d r0| = 1
        g = q_mul(q_conj(d), g);
      }
      gr = q_normalize_bwd(r0, n0, g);
    }
    if (dL_djoints) { dL_djoints[3 * a] = gj.x; dL_djoints[3 * a + 1] = gj.y; dL_djoints[3 * a + 2] = gj.z; }
    if (dL_dsk_r) { dL_dsk_r[4 * a] = gr.x; dL_dsk_r[4 * a + 1] = gr.y; dL_dsk_r[4 * a + 2] = gr.z; dL_dsk_r[4 * a + 3] = gr.w; }
    if (dL_dsk_d_rot)
      for (int c = 0; c < 4; c++) dL_dsk_d_rot[4 * a + c] = __ldcg(jacc + NJ * a + 7 + c);
    if (dL_dsk_d_scale)
      for (int c = 0; c < 3; c++) dL_dsk_d_scale[3 * a + c] = __ldcg(jacc + NJ * a + 11 + c);
    if (dL_dsp_radius) dL_dsp_radius[a] = __ldcg(jacc + NJ * a + 17);
    if (dL_dsp_weight) dL_dsp_weight[a] = __ldcg(jacc + NJ * a + 18);
  }
}

__global__ void __launch_bounds__(1024)
fk_bwd_kernel(skgs_skeleton sk, const float* __restrict__ jacc, const float* __restrict__ dL_dsk_T_direct,
              float* __restrict__ dL_djoints, float* __restrict__ dL_dsk_r, float* __restrict__ dL_dsk_d_rot,
              float* __restrict__ dL_dsk_d_scale, float* __restrict__ dL_dg_tr, float* __restrict__ dL_dsp_radius,
              float* __restrict__ dL_dsp_weight) {
  extern __shared__ float ksm[];
  pdl_wait();
  pdl_trigger();
  fk_bwd_body(sk, jacc, dL_dsk_T_direct, dL_djoints, dL_dsk_r, dL_dsk_d_rot, dL_dsk_d_scale, dL_dg_tr, dL_dsp_radius,
              dL_dsp_weight, ksm);
}

// ------------------------------------------------------------------------------------------------------------------
// output assembly (networks/sk_gs.py:1192,1202-1203; activations networks/gaussian_splatting.py:155-160)
// ------------------------------------------------------------------------------------------------------------------
__global__ void assemble_fwd_kernel(int P, const float* __restrict__ xyz, const float* __restrict__ scaling,
                                    const float* __restrict__ rotation, const float* __restrict__ opacity,
                                    const float* __restrict__ d_xyz, const float* __restrict__ d_rot,
                                    const float* __restrict__ d_scale, float* __restrict__ points,
                                    float* __restrict__ scales, float* __restrict__ rotations,
                                    float* __restrict__ opacities) {
  pdl_wait();
  pdl_trigger();
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) {
    const size_t i3 = 3 * (size_t)i;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const Assembled a = assemble_gaussian(
        xyz[i3], xyz[i3 + 1], xyz[i3 + 2], scaling[i3], scaling[i3 + 1], scaling[i3 + 2],
        *reinterpret_cast<const float4*>(rotation + 4 * (size_t)i), opacity[i], d_xyz ? d_xyz[i3] : 0.f,
        d_xyz ? d_xyz[i3 + 1] : 0.f, d_xyz ? d_xyz[i3 + 2] : 0.f,
        d_rot ? *reinterpret_cast<const float4*>(d_rot + 4 * (size_t)i) : z4, d_scale ? d_scale[i3] : 0.f,
        d_scale ? d_scale[i3 + 1] : 0.f, d_scale ? d_scale[i3 + 2] : 0.f);
    points[i3] = a.px; points[i3 + 1] = a.py; points[i3 + 2] = a.pz;
    scales[i3] = a.sx; scales[i3 + 1] = a.sy; scales[i3 + 2] = a.sz;
    *reinterpret_cast<float4*>(rotations + 4 * (size_t)i) = make_float4(a.qx, a.qy, a.qz, a.qw);
    opacities[i] = a.opacity;
  }
}

__global__ void assemble_bwd_kernel(int P, const float* __restrict__ scaling, const float* __restrict__ rotation,
                                    const float* __restrict__ opacity, const float* __restrict__ d_rot,
                                    const float* __restrict__ gp, const float* __restrict__ gs,
                                    const float* __restrict__ gr, const float* __restrict__ go,
                                    float* __restrict__ dxyz, float* __restrict__ dscaling,
                                    float* __restrict__ drotation, float* __restrict__ dopacity,
                                    float* __restrict__ dd_xyz, float* __restrict__ dd_rot,
                                    float* __restrict__ dd_scale) {
  const int stride = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int e = t0; e < 3 * P; e += stride) {
    const float a = gp ? gp[e] : 0.f, b = gs ? gs[e] : 0.f;
    if (dxyz) dxyz[e] = a;
    if (dd_xyz) dd_xyz[e] = a;
    if (dscaling) dscaling[e] = b * expf(scaling[e]);
    if (dd_scale) dd_scale[e] = b;
  }
  for (int i = t0; i < P; i += stride) {
    float4 r = *reinterpret_cast<const float4*>(rotation + 4 * i);
    if (d_rot) {
      const float4 d = *reinterpret_cast<const float4*>(d_rot + 4 * i);
      r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
    }
    const float4 g = gr ? *reinterpret_cast<const float4*>(gr + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float nn = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
    float4 o;
    if (nn > 1e-12f) {
      const float inv = 1.0f / nn;
      const float ux = r.x * inv, uy = r.y * inv, uz = r.z * inv, uw = r.w * inv;
      const float d = g.x * ux + g.y * uy + g.z * uz + g.w * uw;
      o = make_float4((g.x - d * ux) * inv, (g.y - d * uy) * inv, (g.z - d * uz) * inv, (g.w - d * uw) * inv);
    } else {
      o = make_float4(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f, g.w * 1e12f);
    }
    if (drotation) *reinterpret_cast<float4*>(drotation + 4 * i) = o;
    if (dd_rot) *reinterpret_cast<float4*>(dd_rot + 4 * i) = o;
    if (dopacity) {
      const float s = sigmoidf(opacity[i]);
      dopacity[i] = (go ? go[i] : 0.f) * s * (1.0f - s);
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// sp-stage: the same per-Gaussian kernels with the transform table coming from the per-superpoint SE3
// the deformation network predicts instead of from forward kinematics.  Reference: `warp` networks/sk_gs.py:776-828,
// `calc_LBS_weight` :751-774, called from `sp_stage` :830-856.  lietorch (un-vendored) supplies SE3.act = R(q) p + t
// (my_ext/_C/include/lie.h:246) and, in backward, gradients w.r.t. the 7-vector projected onto the tangent space of
// R^3 x S^3 at (t, q) - reproduced here by differentiating through q / |q|.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sp_table_kernel(skgs_superpoints sp, float* __restrict__ spT, float* __restrict__ table) {
  const int M = sp.M;
  pdl_wait();
  pdl_trigger();
  float* pos = table;
  float* tt = pos + 3 * M;
  float* R = tt + 3 * M;
  float* dq = R + 9 * M;
  float* ds = dq + 4 * M;
  float* aux = ds + 3 * M;
  for (int a = threadIdx.x; a < M; a += blockDim.x) {
    const Vec3 c = load_v(sp.sp_points + 3 * a);
    const Quat raw = load_q(sp.sp_r + 4 * a);
    const Quat q = q_normalize(raw);
    Vec3 t = load_v(sp.sp_t + 3 * a);
    if (sp.method == SKGS_WARP_LBS_C) {  // rotation about the superpoint: t + c + R(q)(-c)   (:797-798)
      const Vec3 rc = q_rotate(q, {-c.x, -c.y, -c.z});
      t = {t.x + c.x + rc.x, t.y + c.y + rc.y, t.z + c.z + rc.z};
    }
    if (spT != nullptr) {
      spT[7 * a] = t.x; spT[7 * a + 1] = t.y; spT[7 * a + 2] = t.z;
      spT[7 * a + 3] = raw.x; spT[7 * a + 4] = raw.y; spT[7 * a + 5] = raw.z; spT[7 * a + 6] = raw.w;
    }
    pos[3 * a] = c.x; pos[3 * a + 1] = c.y; pos[3 * a + 2] = c.z;
    tt[3 * a] = t.x; tt[3 * a + 1] = t.y; tt[3 * a + 2] = t.z;
    quat_to_rows(q, R + 9 * a);
    const float* rot = sp.sp_rot ? sp.sp_rot : sp.sp_r;  // :818-821
    for (int k = 0; k < 4; k++) dq[4 * a + k] = rot[4 * a + k];
    for (int k = 0; k < 3; k++) ds[3 * a + k] = sp.sp_scale ? sp.sp_scale[3 * a + k] : 0.f;
    float a0 = 0.f, a1 = 1.f;
    if (sp.mode == SKGS_LBS_KERNEL || sp.mode == SKGS_LBS_WEIGHTED_KERNEL) {
      const float r = expf(sp.sp_radius[a]);
      a0 = 1.0f / (2.0f * r * r);
      a1 = sp.mode == SKGS_LBS_WEIGHTED_KERNEL ? sigmoidf(sp.sp_weight[a]) : 1.0f;
    }
    aux[2 * a] = a0;
    aux[2 * a + 1] = a1;
  }
}

struct SpBwdOut {
  const float* dL_dspT;  // direct gradient on the returned [M][7] or NULL
  float *dL_dsp_points, *dL_dsp_t, *dL_dsp_r, *dL_dsp_rot, *dL_dsp_scale, *dL_dsp_radius, *dL_dsp_weight;
};

// per-superpoint tail of the backward: the per-joint sums of the LBS backward -> gradients of warp's inputs
__global__ void __launch_bounds__(256)
sp_bwd_kernel(skgs_superpoints sp, const float* __restrict__ jacc, SpBwdOut o) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= sp.M) return;
  const float* acc = jacc + NJ * a;
  Vec3 gt = {acc[0], acc[1], acc[2]};
  Quat gq = {acc[3], acc[4], acc[5], acc[6]};
  if (o.dL_dspT != nullptr) {
    gt.x += o.dL_dspT[7 * a]; gt.y += o.dL_dspT[7 * a + 1]; gt.z += o.dL_dspT[7 * a + 2];
    gq.x += o.dL_dspT[7 * a + 3]; gq.y += o.dL_dspT[7 * a + 4]; gq.z += o.dL_dspT[7 * a + 5];
    gq.w += o.dL_dspT[7 * a + 6];
  }
  const Quat raw = load_q(sp.sp_r + 4 * a);
  const float n = sqrtf(raw.x * raw.x + raw.y * raw.y + raw.z * raw.z + raw.w * raw.w);
  const Quat q = q_normalize(raw);
  Vec3 gc = {acc[14], acc[15], acc[16]};  // through the squared distances of the weight function
  if (sp.method == SKGS_WARP_LBS_C) {
    const Vec3 c = load_v(sp.sp_points + 3 * a);
    const Vec3 back = q_rotate_inv(q, gt);
    gc.x += gt.x - back.x; gc.y += gt.y - back.y; gc.z += gt.z - back.z;
    const Quat g1 = q_rotate_grad_q(q, {-c.x, -c.y, -c.z}, gt);
    gq.x += g1.x; gq.y += g1.y; gq.z += g1.z; gq.w += g1.w;
  }
  Quat gr = q_normalize_bwd(q, n, gq);  // tangent-space projection (lietorch FromVec backward)
  if (sp.sp_rot == nullptr) {            // d_rotation blends sp_r itself (:820-821): plain gradient
    gr.x += acc[7]; gr.y += acc[8]; gr.z += acc[9]; gr.w += acc[10];
  }
  if (o.dL_dsp_points) { o.dL_dsp_points[3 * a] = gc.x; o.dL_dsp_points[3 * a + 1] = gc.y; o.dL_dsp_points[3 * a + 2] = gc.z; }
  if (o.dL_dsp_t) { o.dL_dsp_t[3 * a] = gt.x; o.dL_dsp_t[3 * a + 1] = gt.y; o.dL_dsp_t[3 * a + 2] = gt.z; }
  if (o.dL_dsp_r) { o.dL_dsp_r[4 * a] = gr.x; o.dL_dsp_r[4 * a + 1] = gr.y; o.dL_dsp_r[4 * a + 2] = gr.z; o.dL_dsp_r[4 * a + 3] = gr.w; }
  if (o.dL_dsp_rot)
    for (int k = 0; k < 4; k++) o.dL_dsp_rot[4 * a + k] = sp.sp_rot ? acc[7 + k] : 0.f;
  if (o.dL_dsp_scale)
    for (int k = 0; k < 3; k++) o.dL_dsp_scale[3 * a + k] = sp.sp_scale ? acc[11 + k] : 0.f;
  if (o.dL_dsp_radius) o.dL_dsp_radius[a] = acc[17];
  if (o.dL_dsp_weight) o.dL_dsp_weight[a] = acc[18];
}

static int fk_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int check_skeleton(const skgs_skeleton* sk, int P) {
  SKGS_CHECK_ARG(sk != nullptr, "skeleton is NULL");
  SKGS_CHECK_ARG(sk->M >= 1 && sk->M <= 1024, "M=%d out of range [1,1024] (reference guard sp_gs_joint.cu:114)", sk->M);
  SKGS_CHECK_ARG(sk->L >= 0 && sk->L <= 16, "L=%d out of range", sk->L);
  SKGS_CHECK_ARG(sk->root >= 0 && sk->root < sk->M, "root=%d out of range", sk->root);
  SKGS_CHECK_ARG(sk->K >= 1 && sk->K <= MAXK && sk->K <= sk->M, "K=%d must be in [1,%d] and <= M", sk->K, MAXK);
  SKGS_CHECK_ARG(sk->mode >= 0 && sk->mode <= 3, "unknown LBS mode %d", sk->mode);
  SKGS_CHECK_ARG(sk->joints && sk->sk_r && sk->sk_d_rot && sk->sk_d_scale, "joints/sk_r/sk_d_rot/sk_d_scale required");
  SKGS_CHECK_ARG(sk->L == 0 || sk->parents, "parents required when L > 0");
  SKGS_CHECK_ARG(sk->mode != SKGS_LBS_W || sk->sp_W || P == 0, "mode W needs sp_W");
  SKGS_CHECK_ARG((sk->mode != SKGS_LBS_KERNEL && sk->mode != SKGS_LBS_WEIGHTED_KERNEL) || sk->sp_radius,
                 "kernel modes need sp_radius");
  SKGS_CHECK_ARG(sk->mode != SKGS_LBS_WEIGHTED_KERNEL || sk->sp_weight, "weighted_kernel needs sp_weight");
  SKGS_CHECK_ARG(sk->mode != SKGS_LBS_DIST || sk->temperature > 0.f, "dist mode needs temperature > 0");
  SKGS_CHECK_ARG(sk->sk_r_delta == nullptr || sk->sk_r_delta_dim == 3 || sk->sk_r_delta_dim == 4,
                 "sk_r_delta_dim must be 3 or 4");
  return SKGS_OK;
}

// sk_T [M][7] (may be NULL) and the joint table (JT_FLOATS * M floats) of one skeleton pose: one CTA
int launch_fk_table(const skgs_skeleton* sk, float* sk_T, float* table, cudaStream_t st) {
  int rc = check_skeleton(sk, 0);
  if (rc) return rc;
  const size_t smem = fk_table_smem_bytes(sk->M);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    SKGS_CUDA(cudaFuncSetAttribute(fk_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int threads = ((sk->M + 31) / 32) * 32;
  threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
  {
    ProfScope prof_("fk_table_kernel", st);
    SKGS_CUDA(launch_pdl(fk_table_kernel, dim3(1), dim3(threads), smem, st, *sk, sk_T, table));
    SKGS_CHECK_LAUNCH("fk_table_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs

using namespace skgs;

extern "C" {

int skgs_fk_lbs_forward(const skgs_skeleton* sk, int32_t P, const float* xyz, float* d_xyz, float* d_rot,
                        float* d_scale, float* sk_T, float* weights, int64_t* indices, void* workspace, void* stream) {
  int rc = check_skeleton(sk, P);
  if (rc) return rc;
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  SKGS_CHECK_ARG(P == 0 || (xyz && d_xyz && d_rot && d_scale && weights && indices), "NULL per-Gaussian buffer");
  SKGS_CHECK_ARG(workspace != nullptr, "workspace (skgs_fk_lbs_workspace_bytes) is required");
  cudaStream_t st = (cudaStream_t)stream;
  float* table = reinterpret_cast<float*>(workspace);
  rc = launch_fk_table(sk, sk_T, table, st);
  if (rc || P == 0) return rc;
  const size_t smem = (size_t)sk->M * JT_FLOATS * sizeof(float);
  int grid = (P + FK_THREADS - 1) / FK_THREADS;
  const int cap = fk_num_sms() * 8;
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  {
    ProfScope prof_("lbs_fwd_kernel", st);
#define SKGS_FK_CASE(KK)                                                                                            \
  case KK: {                                                                                                        \
    static size_t smem_set = 0;                                                                                     \
    if (smem > 48 * 1024 && smem > smem_set) {                                                                      \
      SKGS_CUDA(cudaFuncSetAttribute(lbs_fwd_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                              \
    }                                                                                                               \
    SKGS_CUDA(launch_pdl(lbs_fwd_kernel<KK>, dim3(grid), dim3(FK_THREADS), smem, st, sk->M, sk->mode,              \
                         sk->temperature, (const float*)table, sk->sp_W, P, xyz, d_xyz, d_rot, d_scale, weights,   \
                         indices));                                                                                 \
  } break;
    switch (sk->K) {
      SKGS_FK_CASE(1) SKGS_FK_CASE(2) SKGS_FK_CASE(3) SKGS_FK_CASE(4) SKGS_FK_CASE(5) SKGS_FK_CASE(6) SKGS_FK_CASE(7)
      SKGS_FK_CASE(8)
      default: break;
    }
#undef SKGS_FK_CASE
    SKGS_CHECK_LAUNCH("lbs_fwd_kernel");
  }
  return SKGS_OK;
}

// forward: the joint table (JT_FLOATS per joint); backward: the per-joint accumulators (NJ per joint) + one ticket word
size_t skgs_fk_lbs_workspace_bytes(int32_t M) {
  return (size_t)(M > 0 ? M : 1) * (JT_FLOATS > NJ ? JT_FLOATS : NJ) * sizeof(float) + 64;
}

int skgs_fk_lbs_backward(const skgs_skeleton* sk, int32_t P, const float* xyz, const float* sk_T, const float* weights,
                         const int64_t* indices, const float* dL_dd_xyz, const float* dL_dd_rot,
                         const float* dL_dd_scale, const float* dL_dsk_T, const float* dL_dweights, float* dL_djoints,
                         float* dL_dsk_r, float* dL_dsk_d_rot, float* dL_dsk_d_scale, float* dL_dg_tr, float* dL_dsp_W,
                         float* dL_dsp_W_knn, float* dL_dsp_radius, float* dL_dsp_weight, void* workspace,
                         void* stream) {
  int rc = check_skeleton(sk, P);
  if (rc) return rc;
  SKGS_CHECK_ARG(workspace != nullptr && sk_T != nullptr, "workspace and sk_T are required");
  SKGS_CHECK_ARG(P == 0 || (xyz && weights && indices), "NULL per-Gaussian buffer");
  cudaStream_t st = (cudaStream_t)stream;
  float* jacc = reinterpret_cast<float*>(workspace);
  SKGS_CUDA(cudaMemsetAsync(jacc, 0, skgs_fk_lbs_workspace_bytes(sk->M), st));
  if (P > 0 && sk->M <= JM_MAX_M) {
    const size_t smem = lbs_bwd_jm_smem_bytes(sk->M);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      SKGS_CUDA(cudaFuncSetAttribute(lbs_bwd_jm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    int grid = (P + JM_CHUNK - 1) / JM_CHUNK;
    const int cap = fk_num_sms() * 2;
    grid = grid > cap ? cap : grid;
    if (sk->mode == SKGS_LBS_W && dL_dsp_W != nullptr)
      SKGS_CUDA(cudaMemsetAsync(dL_dsp_W, 0, (size_t)P * sk->M * sizeof(float), st));
    {
      ProfScope prof_("lbs_bwd_kernel", st);
      FkBwdOut fk{dL_dsk_T, dL_djoints, dL_dsk_r, dL_dsk_d_rot, dL_dsk_d_scale, dL_dg_tr, dL_dsp_radius, dL_dsp_weight};
      uint32_t* done = reinterpret_cast<uint32_t*>(jacc + (size_t)sk->M * NJ);  // zeroed with the accumulators above
      SKGS_CUDA(launch_pdl(lbs_bwd_jm_kernel, dim3(grid), dim3(FK_THREADS), smem, st, *sk, P, xyz, sk_T, weights,
                           indices, dL_dd_xyz, dL_dd_rot, dL_dd_scale, dL_dweights, dL_dsp_W, dL_dsp_W_knn, jacc, done,
                           fk));
      SKGS_CHECK_LAUNCH("lbs_bwd_jm_kernel");
    }
    return SKGS_OK;  // the FK backward ran as the tail of the kernel
  } else if (P > 0) {
    const size_t smem = lbs_bwd_smem_bytes(sk->M);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      SKGS_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    int grid = (P + FK_THREADS - 1) / FK_THREADS;
    const int cap = fk_num_sms() * 4;
    grid = grid > cap ? cap : grid;
    {
      ProfScope prof_("lbs_bwd_kernel", st);
      lbs_bwd_kernel<<<grid, FK_THREADS, smem, st>>>(*sk, P, xyz, sk_T, weights, indices, dL_dd_xyz, dL_dd_rot,
                                                   dL_dd_scale, dL_dweights, dL_dsp_W, dL_dsp_W_knn, jacc, 0);
      SKGS_CHECK_LAUNCH("lbs_bwd_kernel");
    }
  }
  {
    const size_t smem = fk_bwd_smem_bytes(sk->M);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      SKGS_CUDA(cudaFuncSetAttribute(fk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    int threads = ((sk->M + 31) / 32) * 32;
    threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
    {
      ProfScope prof_("fk_bwd_kernel", st);
      fk_bwd_kernel<<<1, threads, smem, st>>>(*sk, jacc, dL_dsk_T, dL_djoints, dL_dsk_r, dL_dsk_d_rot, dL_dsk_d_scale,
                                            dL_dg_tr, dL_dsp_radius, dL_dsp_weight);
    SKGS_CHECK_LAUNCH("fk_bwd_kernel");
    }
  }
  return SKGS_OK;
}

int skgs_assemble_forward(int32_t P, const float* xyz, const float* scaling, const float* rotation,
                          const float* opacity, const float* d_xyz, const float* d_rot, const float* d_scale,
                          float* points, float* scales, float* rotations, float* opacities, void* stream) {
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  if (P == 0) return SKGS_OK;
  SKGS_CHECK_ARG(xyz && scaling && rotation && opacity && points && scales && rotations && opacities, "NULL buffer");
  int grid = (P + 255) / 256;
  const int cap = fk_num_sms() * 8;
  grid = grid > cap ? cap : grid;
  {
    ProfScope prof_("assemble_fwd_kernel", (cudaStream_t)stream);
    SKGS_CUDA(launch_pdl(assemble_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, P, xyz, scaling, rotation,
                         opacity, d_xyz, d_rot, d_scale, points, scales, rotations, opacities));
    SKGS_CHECK_LAUNCH("assemble_fwd_kernel");
  }
  return SKGS_OK;
}

int skgs_assemble_backward(int32_t P, const float* scaling, const float* rotation, const float* opacity,
                           const float* d_rot, const float* dL_dpoints, const float* dL_dscales,
                           const float* dL_drotations, const float* dL_dopacities, float* dL_dxyz, float* dL_dscaling,
                           float* dL_drotation, float* dL_dopacity, float* dL_dd_xyz, float* dL_dd_rot,
                           float* dL_dd_scale, void* stream) {
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  if (P == 0) return SKGS_OK;
  SKGS_CHECK_ARG(scaling && rotation && opacity, "NULL buffer");
  int grid = (3 * P + 255) / 256;
  const int cap = fk_num_sms() * 8;
  grid = grid > cap ? cap : grid;
  {
    ProfScope prof_("assemble_bwd_kernel", (cudaStream_t)stream);
    assemble_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, scaling, rotation, opacity, d_rot, dL_dpoints,
                                                             dL_dscales, dL_drotations, dL_dopacities, dL_dxyz,
                                                             dL_dscaling, dL_drotation, dL_dopacity, dL_dd_xyz,
                                                             dL_dd_rot, dL_dd_scale);
  SKGS_CHECK_LAUNCH("assemble_bwd_kernel");
  }
  return SKGS_OK;
}


// ---- sp-stage LBS (f-4)
static int check_superpoints(const skgs_superpoints* sp, int P) {
  SKGS_CHECK_ARG(sp != nullptr, "superpoints are NULL");
  SKGS_CHECK_ARG(sp->M >= 1 && sp->M <= 1024, "M=%d out of range [1,1024]", sp->M);
  SKGS_CHECK_ARG(sp->K >= 1 && sp->K <= MAXK && sp->K <= sp->M, "K=%d must be in [1,%d] and <= M", sp->K, MAXK);
  SKGS_CHECK_ARG(sp->mode >= 0 && sp->mode <= 3, "unknown LBS mode %d", sp->mode);
  SKGS_CHECK_ARG(sp->method >= 0 && sp->method <= 2, "unknown warp method %d", sp->method);
  SKGS_CHECK_ARG(sp->sp_points && sp->sp_t && sp->sp_r, "sp_points / sp_t / sp_r required");
  SKGS_CHECK_ARG(sp->mode != SKGS_LBS_W || sp->sp_W || P == 0, "mode W needs sp_W");
  SKGS_CHECK_ARG((sp->mode != SKGS_LBS_KERNEL && sp->mode != SKGS_LBS_WEIGHTED_KERNEL) || sp->sp_radius,
                 "kernel modes need sp_radius");
  SKGS_CHECK_ARG(sp->mode != SKGS_LBS_WEIGHTED_KERNEL || sp->sp_weight, "weighted_kernel needs sp_weight");
  SKGS_CHECK_ARG(sp->mode != SKGS_LBS_DIST || sp->temperature > 0.f, "dist mode needs temperature > 0");
  return SKGS_OK;
}

int skgs_sp_lbs_forward(const skgs_superpoints* sp, int32_t P, const float* points, float* d_points,
                        float* d_rotation, float* d_scales, float* spT, float* weights, int64_t* indices,
                        void* workspace, void* stream) {
  int rc = check_superpoints(sp, P);
  if (rc) return rc;
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  SKGS_CHECK_ARG(P == 0 || (points && d_points && d_rotation && d_scales && weights && indices),
                 "NULL per-Gaussian buffer");
  SKGS_CHECK_ARG(workspace != nullptr, "workspace (skgs_fk_lbs_workspace_bytes) is required");
  cudaStream_t st = (cudaStream_t)stream;
  float* table = reinterpret_cast<float*>(workspace);
  {
    int threads = ((sp->M + 31) / 32) * 32;
    threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
    ProfScope prof_("sp_table_kernel", st);
    SKGS_CUDA(launch_pdl(sp_table_kernel, dim3(1), dim3(threads), 0, st, *sp, spT, table));
    SKGS_CHECK_LAUNCH("sp_table_kernel");
  }
  if (P == 0) return SKGS_OK;
  const size_t smem = (size_t)sp->M * JT_FLOATS * sizeof(float);
  int grid = (P + FK_THREADS - 1) / FK_THREADS;
  const int cap = fk_num_sms() * 8;
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  {
    ProfScope prof_("lbs_fwd_kernel", st);
#define SKGS_SP_LAUNCH(KK, LG)                                                                                         \
  {                                                                                                                    \
    static size_t smem_set = 0;                                                                                        \
    if (smem > 48 * 1024 && smem > smem_set) {                                                                         \
      SKGS_CUDA(cudaFuncSetAttribute(lbs_fwd_kernel<KK, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      smem_set = smem;                                                                                                 \
    }                                                                                                                  \
    SKGS_CUDA(launch_pdl(lbs_fwd_kernel<KK, LG>, dim3(grid), dim3(FK_THREADS), smem, st, sp->M, sp->mode,            \
                         sp->temperature, (const float*)table, sp->sp_W, P, points, d_points, d_rotation, d_scales,   \
                         weights, indices));                                                                           \
  }
#define SKGS_SP_CASE(KK)                                   \
  case KK:                                                 \
    if (sp->method == SKGS_WARP_LARGEST) SKGS_SP_LAUNCH(KK, true) else SKGS_SP_LAUNCH(KK, false) break;
    switch (sp->K) {
      SKGS_SP_CASE(1) SKGS_SP_CASE(2) SKGS_SP_CASE(3) SKGS_SP_CASE(4) SKGS_SP_CASE(5) SKGS_SP_CASE(6) SKGS_SP_CASE(7)
      SKGS_SP_CASE(8)
      default: break;
    }
#undef SKGS_SP_CASE
#undef SKGS_SP_LAUNCH
    SKGS_CHECK_LAUNCH("lbs_fwd_kernel");
  }
  return SKGS_OK;
}

int skgs_sp_lbs_backward(const skgs_superpoints* sp, int32_t P, const float* points, const float* spT,
                         const float* weights, const int64_t* indices, const float* dL_dd_points,
                         const float* dL_dd_rotation, const float* dL_dd_scales, const float* dL_dspT,
                         const float* dL_dweights, float* dL_dsp_points, float* dL_dsp_t, float* dL_dsp_r,
                         float* dL_dsp_rot, float* dL_dsp_scale, float* dL_dsp_W, float* dL_dsp_W_knn,
                         float* dL_dsp_radius, float* dL_dsp_weight, void* workspace, void* stream) {
  int rc = check_superpoints(sp, P);
  if (rc) return rc;
  SKGS_CHECK_ARG(workspace != nullptr && spT != nullptr, "workspace and spT are required");
  SKGS_CHECK_ARG(P == 0 || (points && weights && indices), "NULL per-Gaussian buffer");
  cudaStream_t st = (cudaStream_t)stream;
  float* jacc = reinterpret_cast<float*>(workspace);
  float* zeros = jacc + (size_t)sp->M * NJ + 16;  // 3 M zero floats standing in for an absent sp_scale
  SKGS_CUDA(cudaMemsetAsync(jacc, 0, skgs_sp_lbs_workspace_bytes(sp->M), st));
  // the per-Gaussian backward kernels see the superpoints as a skeleton without a kinematic chain
  skgs_skeleton sk{};
  sk.M = sp->M; sk.L = 0; sk.root = 0; sk.K = sp->K; sk.mode = sp->mode; sk.temperature = sp->temperature;
  sk.joints = sp->sp_points; sk.sk_r = sp->sp_r; sk.sk_d_rot = sp->sp_rot ? sp->sp_rot : sp->sp_r;
  sk.sk_d_scale = sp->sp_scale ? sp->sp_scale : zeros;
  sk.sp_W = sp->sp_W; sk.sp_radius = sp->sp_radius; sk.sp_weight = sp->sp_weight;
  if (P > 0 && sp->mode == SKGS_LBS_W && dL_dsp_W != nullptr)
    SKGS_CUDA(cudaMemsetAsync(dL_dsp_W, 0, (size_t)P * sp->M * sizeof(float), st));
  if (P > 0 && sp->M <= JM_MAX_M && sp->method != SKGS_WARP_LARGEST) {
    const size_t smem = lbs_bwd_jm_smem_bytes(sp->M);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      SKGS_CUDA(cudaFuncSetAttribute(lbs_bwd_jm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    int grid = (P + JM_CHUNK - 1) / JM_CHUNK;
    const int cap = fk_num_sms() * 2;
    grid = grid > cap ? cap : grid;
    ProfScope prof_("lbs_bwd_kernel", st);
    FkBwdOut none{};
    SKGS_CUDA(launch_pdl(lbs_bwd_jm_kernel, dim3(grid), dim3(FK_THREADS), smem, st, sk, P, points, spT, weights,
                         indices, dL_dd_points, dL_dd_rotation, dL_dd_scales, dL_dweights, dL_dsp_W, dL_dsp_W_knn, jacc,
                         (uint32_t*)nullptr, none));
    SKGS_CHECK_LAUNCH("lbs_bwd_jm_kernel");
  } else if (P > 0) {
    const size_t smem = lbs_bwd_smem_bytes(sp->M);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
      SKGS_CUDA(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    int grid = (P + FK_THREADS - 1) / FK_THREADS;
    const int cap = fk_num_sms() * 2;
    grid = grid > cap ? cap : grid;
    ProfScope prof_("lbs_bwd_kernel", st);
    lbs_bwd_kernel<<<grid, FK_THREADS, smem, st>>>(sk, P, points, spT, weights, indices, dL_dd_points, dL_dd_rotation,
                                                 dL_dd_scales, dL_dweights, dL_dsp_W, dL_dsp_W_knn, jacc,
                                                 sp->method == SKGS_WARP_LARGEST ? 1 : 0);
    SKGS_CHECK_LAUNCH("lbs_bwd_kernel");
  }
  {
    ProfScope prof_("sp_bwd_kernel", st);
    SpBwdOut o{dL_dspT, dL_dsp_points, dL_dsp_t, dL_dsp_r, dL_dsp_rot, dL_dsp_scale, dL_dsp_radius, dL_dsp_weight};
    sp_bwd_kernel<<<(sp->M + 255) / 256, 256, 0, st>>>(*sp, jacc, o);
    SKGS_CHECK_LAUNCH("sp_bwd_kernel");
  }
  return SKGS_OK;
}

size_t skgs_sp_lbs_workspace_bytes(int32_t M) {
  const size_t m = (size_t)(M > 0 ? M : 1);
  const size_t fwd = m * JT_FLOATS * sizeof(float), bwd = (m * (NJ + 3) + 16) * sizeof(float);
  return (fwd > bwd ? fwd : bwd) + 64;
}

}  // extern "C"
