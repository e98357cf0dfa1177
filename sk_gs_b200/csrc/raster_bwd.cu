// raster_bwd.cu - backward half of the tile rasterizer for sm_100a:
//   (K6 composite_bwd_kernel lives in composite.cu)
//   K7 preprocess_bwd_kernel : conic -> cov2D -> cov3D -> (scale, rotation); mean2D / depth / SH colour -> mean3D; SH
// Semantics: SURVEY.md App. A.7-A.8 (reference gaussian_render.cu:182-341, gaussian_preprocess_colmap.cu:240-481,
// gaussian_rasterizer_backwrad.cu:26-127).  The contributing-pair tests (power, alpha) use the same contraction-proof
// helpers as the forward kernel so that exactly the same pairs are visited.
#include "common.cuh"

namespace skgs {

__device__ __constant__ float b_SH_C0 = 0.28209479177387814f;
__device__ __constant__ float b_SH_C1 = 0.4886025119029199f;
__device__ __constant__ float b_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                            -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float b_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                            0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                            -0.5900435899266435f};

constexpr int NGRAD = 12;  // packed per-Gaussian accumulators (composite.cu): mx my ca cb | cc op z - | r g b -

// ------------------------------------------------------------------------------------------------------------------
// K7: preprocess backward
// ------------------------------------------------------------------------------------------------------------------
#ifndef SKGS_PB_THREADS
#define SKGS_PB_THREADS 256
#endif
#ifndef SKGS_PB_MINBLOCKS
#define SKGS_PB_MINBLOCKS 2   // 128 registers: two CTAs per SM (132 registers without the cap = one)
#endif
constexpr int PB_THREADS = SKGS_PB_THREADS;

// Optional fusion of the assembly backward (networks/sk_gs.py:1192,1202-1203 + the activations of
// networks/gaussian_splatting.py:155-160) into this kernel: with `scaling != NULL` the gradients of the assembled
// Gaussians are chained on the spot to the canonical parameters and to the LBS outputs, and the intermediate
// dL/dscales, dL/drotations, dL/dopacity never touch HBM (skgs_assemble_backward as a second kernel otherwise).
struct AssembleBwd {
  const float *scaling, *rotation, *opacity, *d_rot;  // _scaling (log), _rotation (raw), _opacity (logit), LBS d_rot or NULL
  float *dscaling, *drotation, *dopacity;             // gradients of the canonical parameters
  float* dd_scale;                                    // gradient of the LBS output d_scale (= dL/dscales)
  float *dd_xyz, *dd_rot;                             // optional private copies of dL_dmeans3D / drotation (= dL/d_xyz,
                                                      // dL/d_rot) for a caller that hands those two buffers to an
                                                      // all-reduce while the LBS backward still reads them
};

__global__ void __launch_bounds__(PB_THREADS, SKGS_PB_MINBLOCKS)
preprocess_bwd_kernel(RasterParams rp, const float* __restrict__ means3D, const float* __restrict__ shs,
                      const float* __restrict__ scales, const float* __restrict__ rotations,
                      const float* __restrict__ cov3D_in, const int32_t* __restrict__ radii,
                      const uint8_t* __restrict__ clamped, float* __restrict__ ggrad,
                      uint32_t* __restrict__ bwd_ticket, float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dcolors, float* __restrict__ dL_dopacity, float* __restrict__ dL_dscales,
                      float* __restrict__ dL_drotations, float* __restrict__ dL_dcov3D, AssembleBwd ab) {
  __shared__ float s_V[16], s_P[16], s_cam[3];
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (tid < 16) {
    s_V[tid] = rp.view[tid];
    s_P[tid] = rp.proj[tid];
  }
  if (tid < 3) s_cam[tid] = rp.campos[tid];
  // restore the invariants composite_bwd relies on, so that a second backward on the same state needs no memset:
  // its work ticket is 0 and the per-Gaussian accumulators are zero outside the composite_bwd -> here window
  if (blockIdx.x == 0 && tid == 0) *bwd_ticket = 0u;
  __syncthreads();
  const int i = blockIdx.x * PB_THREADS + tid;
  if (i >= rp.P) return;
  const float* V = s_V;
  const float* Pm = s_P;
  const int M3 = rp.M * 3;
  const bool vis = radii[i] > 0;
  float g[NGRAD];
  {
    float4* gp = reinterpret_cast<float4*>(ggrad + (size_t)i * NGRAD);
    const float4 a = gp[0], b = gp[1], c = gp[2];
    gp[0] = gp[1] = gp[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
    g[8] = c.x; g[9] = c.y; g[10] = c.z; g[11] = c.w;
  }
  if (dL_dmeans2D) {
    dL_dmeans2D[3 * i] = vis ? g[0] : 0.f;
    dL_dmeans2D[3 * i + 1] = vis ? g[1] : 0.f;
    dL_dmeans2D[3 * i + 2] = 0.f;
  }
  if (dL_dopacity) dL_dopacity[i] = vis ? g[5] : 0.f;
  if (dL_dcolors) {
    dL_dcolors[3 * i] = vis ? g[8] : 0.f;
    dL_dcolors[3 * i + 1] = vis ? g[9] : 0.f;
    dL_dcolors[3 * i + 2] = vis ? g[10] : 0.f;
  }
  if (!vis) {
    dL_dmeans3D[3 * i] = dL_dmeans3D[3 * i + 1] = dL_dmeans3D[3 * i + 2] = 0.f;
    if (ab.dd_xyz != nullptr) ab.dd_xyz[3 * i] = ab.dd_xyz[3 * i + 1] = ab.dd_xyz[3 * i + 2] = 0.f;
    if (dL_dcov3D)
      for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
    if (dL_dsh) {
      float4* d4 = reinterpret_cast<float4*>(dL_dsh + (size_t)i * M3);
      if (M3 % 4 == 0)
        for (int k = 0; k < M3 / 4; k++) d4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      else
        for (int k = 0; k < M3; k++) dL_dsh[(size_t)i * M3 + k] = 0.f;
    }
    if (dL_dscales) dL_dscales[3 * i] = dL_dscales[3 * i + 1] = dL_dscales[3 * i + 2] = 0.f;
    if (dL_drotations)
      dL_drotations[4 * i] = dL_drotations[4 * i + 1] = dL_drotations[4 * i + 2] = dL_drotations[4 * i + 3] = 0.f;
    if (ab.scaling != nullptr) {
      for (int k = 0; k < 3; k++) ab.dscaling[3 * i + k] = ab.dd_scale[3 * i + k] = 0.f;
      *reinterpret_cast<float4*>(ab.drotation + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ab.dd_rot != nullptr) *reinterpret_cast<float4*>(ab.dd_rot + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
      ab.dopacity[i] = 0.f;
    }
    return;
  }
  const float mx = means3D[3 * i], my = means3D[3 * i + 1], mz = means3D[3 * i + 2];
  float c6[6];
#pragma unroll
  for (int k = 0; k < 6; k++) c6[k] = cov3D_in[6 * i + k];
  // ---- conic -> cov2D -> cov3D, T -> J -> t -> mean   (gaussian_preprocess_colmap.cu:240-354)
  const float pvx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
  const float pvy = V[1] * mx + V[5] * my + V[9] * mz + V[13];
  const float pvz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
  const float limx = 1.3f * rp.tanfovx, limy = 1.3f * rp.tanfovy;
  const float txtz = pvx / pvz, tytz = pvy / pvz;
  const float tx = fminf(limx, fmaxf(-limx, txtz)) * pvz;
  const float ty = fminf(limy, fmaxf(-limy, tytz)) * pvz;
  const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
  const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
  const float j00 = rp.fx / pvz, j02 = -(rp.fx * tx) / (pvz * pvz);
  const float j11 = rp.fy / pvz, j12 = -(rp.fy * ty) / (pvz * pvz);
  const float a0[3] = {j00 * V[0] + j02 * V[2], j00 * V[4] + j02 * V[6], j00 * V[8] + j02 * V[10]};
  const float a1[3] = {j11 * V[1] + j12 * V[2], j11 * V[5] + j12 * V[6], j11 * V[9] + j12 * V[10]};
  const float u0[3] = {c6[0] * a0[0] + c6[1] * a0[1] + c6[2] * a0[2], c6[1] * a0[0] + c6[3] * a0[1] + c6[4] * a0[2],
                       c6[2] * a0[0] + c6[4] * a0[1] + c6[5] * a0[2]};
  const float u1[3] = {c6[0] * a1[0] + c6[1] * a1[1] + c6[2] * a1[2], c6[1] * a1[0] + c6[3] * a1[1] + c6[4] * a1[2],
                       c6[2] * a1[0] + c6[4] * a1[1] + c6[5] * a1[2]};
  const float a = (a0[0] * u0[0] + a0[1] * u0[1] + a0[2] * u0[2]) + 0.3f;
  const float b = a0[0] * u1[0] + a0[1] * u1[1] + a0[2] * u1[2];
  const float c = (a1[0] * u1[0] + a1[1] * u1[1] + a1[2] * u1[2]) + 0.3f;
  const float denom = a * c - b * b;
  const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
  const float dcx = g[2], dcy = g[3], dcz = g[4];
  float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
  float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (denom2inv != 0.f) {
    dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
    dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
    dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
    dcov[0] = a0[0] * a0[0] * dL_da + a0[0] * a1[0] * dL_db + a1[0] * a1[0] * dL_dc;
    dcov[3] = a0[1] * a0[1] * dL_da + a0[1] * a1[1] * dL_db + a1[1] * a1[1] * dL_dc;
    dcov[5] = a0[2] * a0[2] * dL_da + a0[2] * a1[2] * dL_db + a1[2] * a1[2] * dL_dc;
    dcov[1] = 2 * a0[0] * a0[1] * dL_da + (a0[0] * a1[1] + a0[1] * a1[0]) * dL_db + 2 * a1[0] * a1[1] * dL_dc;
    dcov[2] = 2 * a0[0] * a0[2] * dL_da + (a0[0] * a1[2] + a0[2] * a1[0]) * dL_db + 2 * a1[0] * a1[2] * dL_dc;
    dcov[4] = 2 * a0[2] * a0[1] * dL_da + (a0[1] * a1[2] + a0[2] * a1[1]) * dL_db + 2 * a1[1] * a1[2] * dL_dc;
  }
  if (dL_dcov3D)
#pragma unroll
    for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = dcov[k];
  float dT0[3], dT1[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    dT0[k] = 2 * u0[k] * dL_da + u1[k] * dL_db;
    dT1[k] = 2 * u1[k] * dL_dc + u0[k] * dL_db;
  }
  const float dL_dJ00 = V[0] * dT0[0] + V[4] * dT0[1] + V[8] * dT0[2];
  const float dL_dJ02 = V[2] * dT0[0] + V[6] * dT0[1] + V[10] * dT0[2];
  const float dL_dJ11 = V[1] * dT1[0] + V[5] * dT1[1] + V[9] * dT1[2];
  const float dL_dJ12 = V[2] * dT1[0] + V[6] * dT1[1] + V[10] * dT1[2];
  const float tzi = 1.f / pvz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
  const float dL_dtx = x_grad_mul * -rp.fx * tz2 * dL_dJ02;
  const float dL_dty = y_grad_mul * -rp.fy * tz2 * dL_dJ12;
  const float dL_dtz = -rp.fx * tz2 * dL_dJ00 - rp.fy * tz2 * dL_dJ11 + (2 * rp.fx * tx) * tz3 * dL_dJ02 +
                       (2 * rp.fy * ty) * tz3 * dL_dJ12;
  float dmx = V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz;
  float dmy = V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz;
  float dmz = V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz;
  // ---- mean2D -> mean3D (:432-452)
  {
    const float hw = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
    const float m_w = 1.0f / (hw + 0.0000001f);
    const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
    const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
    dmx += (Pm[0] * m_w - Pm[3] * mul1) * g[0] + (Pm[1] * m_w - Pm[3] * mul2) * g[1];
    dmy += (Pm[4] * m_w - Pm[7] * mul1) * g[0] + (Pm[5] * m_w - Pm[7] * mul2) * g[1];
    dmz += (Pm[8] * m_w - Pm[11] * mul1) * g[0] + (Pm[9] * m_w - Pm[11] * mul2) * g[1];
  }
  // ---- depth output: z_view = third row of the view rotation . mean
  dmx += V[2] * g[6];
  dmy += V[6] * g[6];
  dmz += V[10] * g[6];
  // ---- SH backward (gaussian_rasterizer_backwrad.cu:26-127)
  if (shs != nullptr && dL_dsh != nullptr) {
    const float dox = mx - s_cam[0], doy = my - s_cam[1], doz = mz - s_cam[2];
    const float len = sqrtf(dox * dox + doy * doy + doz * doz);
    const float x = dox / len, y = doy / len, z = doz / len;
    const uint8_t cl = clamped[i];
    float dRGB[3] = {(cl & 1) ? 0.f : g[8], (cl & 2) ? 0.f : g[9], (cl & 4) ? 0.f : g[10]};
    float sh[48], dsh[48];
    const float* sf = shs + (size_t)i * M3;
    if (M3 % 4 == 0) {
      const float4* sp = reinterpret_cast<const float4*>(sf);
#pragma unroll
      for (int k = 0; k < 12; k++)
        if (k < M3 / 4) {
          const float4 v = __ldg(sp + k);
          sh[4 * k] = v.x; sh[4 * k + 1] = v.y; sh[4 * k + 2] = v.z; sh[4 * k + 3] = v.w;
        }
    } else {
#pragma unroll
      for (int k = 0; k < 48; k++)
        if (k < M3) sh[k] = __ldg(sf + k);
    }
#pragma unroll
    for (int k = 0; k < 48; k++) dsh[k] = 0.f;
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
    const int D = rp.D;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      const float dc = dRGB[ch];
#define SH(k) sh[(k) * 3 + ch]
#define DSH(k) dsh[(k) * 3 + ch]
      float rx = 0.f, ry = 0.f, rz = 0.f;
      DSH(0) = b_SH_C0 * dc;
      if (D > 0) {
        DSH(1) = (-b_SH_C1 * y) * dc;
        DSH(2) = (b_SH_C1 * z) * dc;
        DSH(3) = (-b_SH_C1 * x) * dc;
        rx = -b_SH_C1 * SH(3);
        ry = -b_SH_C1 * SH(1);
        rz = b_SH_C1 * SH(2);
        if (D > 1) {
          DSH(4) = (b_SH_C2[0] * xy) * dc;
          DSH(5) = (b_SH_C2[1] * yz) * dc;
          DSH(6) = (b_SH_C2[2] * (2.f * zz - xx - yy)) * dc;
          DSH(7) = (b_SH_C2[3] * xz) * dc;
          DSH(8) = (b_SH_C2[4] * (xx - yy)) * dc;
          rx += b_SH_C2[0] * y * SH(4) + b_SH_C2[2] * 2.f * -x * SH(6) + b_SH_C2[3] * z * SH(7) +
                b_SH_C2[4] * 2.f * x * SH(8);
          ry += b_SH_C2[0] * x * SH(4) + b_SH_C2[1] * z * SH(5) + b_SH_C2[2] * 2.f * -y * SH(6) +
                b_SH_C2[4] * 2.f * -y * SH(8);
          rz += b_SH_C2[1] * y * SH(5) + b_SH_C2[2] * 2.f * 2.f * z * SH(6) + b_SH_C2[3] * x * SH(7);
          if (D > 2) {
            DSH(9) = (b_SH_C3[0] * y * (3.f * xx - yy)) * dc;
            DSH(10) = (b_SH_C3[1] * xy * z) * dc;
            DSH(11) = (b_SH_C3[2] * y * (4.f * zz - xx - yy)) * dc;
            DSH(12) = (b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dc;
            DSH(13) = (b_SH_C3[4] * x * (4.f * zz - xx - yy)) * dc;
            DSH(14) = (b_SH_C3[5] * z * (xx - yy)) * dc;
            DSH(15) = (b_SH_C3[6] * x * (xx - 3.f * yy)) * dc;
            rx += (b_SH_C3[0] * SH(9) * 3.f * 2.f * xy + b_SH_C3[1] * SH(10) * yz + b_SH_C3[2] * SH(11) * -2.f * xy +
                   b_SH_C3[3] * SH(12) * -3.f * 2.f * xz + b_SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) +
                   b_SH_C3[5] * SH(14) * 2.f * xz + b_SH_C3[6] * SH(15) * 3.f * (xx - yy));
            ry += (b_SH_C3[0] * SH(9) * 3.f * (xx - yy) + b_SH_C3[1] * SH(10) * xz +
                   b_SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + b_SH_C3[3] * SH(12) * -3.f * 2.f * yz +
                   b_SH_C3[4] * SH(13) * -2.f * xy + b_SH_C3[5] * SH(14) * -2.f * yz +
                   b_SH_C3[6] * SH(15) * -3.f * 2.f * xy);
            rz += (b_SH_C3[1] * SH(10) * xy + b_SH_C3[2] * SH(11) * 4.f * 2.f * yz +
                   b_SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + b_SH_C3[4] * SH(13) * 4.f * 2.f * xz +
                   b_SH_C3[5] * SH(14) * (xx - yy));
          }
        }
      }
#undef SH
#undef DSH
      ddx += rx * dc;
      ddy += ry * dc;
      ddz += rz * dc;
    }
    if (M3 % 4 == 0) {
      float4* d4 = reinterpret_cast<float4*>(dL_dsh + (size_t)i * M3);
#pragma unroll
      for (int k = 0; k < 12; k++)
        if (k < M3 / 4) d4[k] = make_float4(dsh[4 * k], dsh[4 * k + 1], dsh[4 * k + 2], dsh[4 * k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < 48; k++)
        if (k < M3) dL_dsh[(size_t)i * M3 + k] = dsh[k];
    }
    const float sum2 = dox * dox + doy * doy + doz * doz;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmx += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
    dmy += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
    dmz += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
  }
  dL_dmeans3D[3 * i] = dmx;
  dL_dmeans3D[3 * i + 1] = dmy;
  dL_dmeans3D[3 * i + 2] = dmz;
  if (ab.dd_xyz != nullptr) {
    ab.dd_xyz[3 * i] = dmx;
    ab.dd_xyz[3 * i + 1] = dmy;
    ab.dd_xyz[3 * i + 2] = dmz;
  }
  // ---- cov3D -> scale, rotation (:357-420)
  if (scales != nullptr && (dL_dscales != nullptr || ab.scaling != nullptr)) {
    float dsc[3];
    float qr, qx, qy, qz;
    const float4 q = *reinterpret_cast<const float4*>(rotations + 4 * i);
    if (rp.quat_wxyz) {
      qr = q.x; qx = q.y; qy = q.z; qz = q.w;
    } else {
      qx = q.x; qy = q.y; qz = q.z; qr = q.w;
    }
    float R[3][3];
    R[0][0] = 1.f - 2.f * (qy * qy + qz * qz);
    R[0][1] = 2.f * (qx * qy - qr * qz);
    R[0][2] = 2.f * (qx * qz + qr * qy);
    R[1][0] = 2.f * (qx * qy + qr * qz);
    R[1][1] = 1.f - 2.f * (qx * qx + qz * qz);
    R[1][2] = 2.f * (qy * qz - qr * qx);
    R[2][0] = 2.f * (qx * qz - qr * qy);
    R[2][1] = 2.f * (qy * qz + qr * qx);
    R[2][2] = 1.f - 2.f * (qx * qx + qy * qy);
    const float s[3] = {rp.mod * scales[3 * i], rp.mod * scales[3 * i + 1], rp.mod * scales[3 * i + 2]};
    const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                            {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                            {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
    float dR[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float v[3];
#pragma unroll
      for (int r = 0; r < 3; r++) v[r] = dS[r][0] * R[0][k] + dS[r][1] * R[1][k] + dS[r][2] * R[2][k];
      const float dot = R[0][k] * v[0] + R[1][k] * v[1] + R[2][k] * v[2];
      dsc[k] = 2.0f * s[k] * dot;
#pragma unroll
      for (int r = 0; r < 3; r++) dR[r][k] = 2.0f * s[k] * s[k] * v[r];
    }
#define MT(k, a) dR[a][k]
    const float dqx = 2 * qy * (MT(1, 0) + MT(0, 1)) + 2 * qz * (MT(2, 0) + MT(0, 2)) +
                      2 * qr * (MT(1, 2) - MT(2, 1)) - 4 * qx * (MT(2, 2) + MT(1, 1));
    const float dqy = 2 * qx * (MT(1, 0) + MT(0, 1)) + 2 * qr * (MT(2, 0) - MT(0, 2)) +
                      2 * qz * (MT(1, 2) + MT(2, 1)) - 4 * qy * (MT(2, 2) + MT(0, 0));
    const float dqz = 2 * qr * (MT(0, 1) - MT(1, 0)) + 2 * qx * (MT(2, 0) + MT(0, 2)) +
                      2 * qy * (MT(1, 2) + MT(2, 1)) - 4 * qz * (MT(1, 1) + MT(0, 0));
    const float dqr = 2 * qz * (MT(0, 1) - MT(1, 0)) + 2 * qy * (MT(2, 0) - MT(0, 2)) + 2 * qx * (MT(1, 2) - MT(2, 1));
#undef MT
    float4 o;
    if (rp.quat_wxyz)
      o = make_float4(dqr, dqx, dqy, dqz);
    else
      o = make_float4(dqx, dqy, dqz, dqr);
    if (dL_dscales != nullptr) {
      dL_dscales[3 * i] = dsc[0]; dL_dscales[3 * i + 1] = dsc[1]; dL_dscales[3 * i + 2] = dsc[2];
    }
    if (dL_drotations != nullptr) *reinterpret_cast<float4*>(dL_drotations + 4 * i) = o;
    if (ab.scaling != nullptr) {  // fused assembly backward (same formulas as assemble_bwd_kernel, fk_lbs.cu)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        ab.dscaling[3 * i + k] = dsc[k] * expf(ab.scaling[3 * i + k]);
        ab.dd_scale[3 * i + k] = dsc[k];
      }
      float4 r = *reinterpret_cast<const float4*>(ab.rotation + 4 * i);
      if (ab.d_rot != nullptr) {
        const float4 d = *reinterpret_cast<const float4*>(ab.d_rot + 4 * i);
        r.x += d.x; r.y += d.y; r.z += d.z; r.w += d.w;
      }
      const float nn = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w);
      float4 go;
      if (nn > 1e-12f) {
        const float inv = 1.0f / nn;
        const float ux = r.x * inv, uy = r.y * inv, uz = r.z * inv, uw = r.w * inv;
        const float d = o.x * ux + o.y * uy + o.z * uz + o.w * uw;
        go = make_float4((o.x - d * ux) * inv, (o.y - d * uy) * inv, (o.z - d * uz) * inv, (o.w - d * uw) * inv);
      } else {
        go = make_float4(o.x * 1e12f, o.y * 1e12f, o.z * 1e12f, o.w * 1e12f);
      }
      *reinterpret_cast<float4*>(ab.drotation + 4 * i) = go;
      if (ab.dd_rot != nullptr) *reinterpret_cast<float4*>(ab.dd_rot + 4 * i) = go;
      const float sg = 1.0f / (1.0f + expf(-ab.opacity[i]));
      ab.dopacity[i] = g[5] * sg * (1.0f - sg);
    }
  }
}

int launch_preprocess_bwd(const RasterParams& rp, const float* means3D, const float* shs, const float* colors_precomp,
                          const float* scales, const float* rotations, const float* cov3D_precomp,
                          const int32_t* radii, char* geom, const skgs_raster_layout& lay, uint32_t* bwd_ticket,
                          float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors, float* dL_dopacity,
                          float* dL_dscales, float* dL_drotations, float* dL_dcov3D, const float* const* assemble_in,
                          float* const* assemble_out, cudaStream_t st) {
  if (rp.P == 0) return SKGS_OK;
  AssembleBwd ab = {};
  if (assemble_in != nullptr) {
    ab.scaling = assemble_in[0]; ab.rotation = assemble_in[1]; ab.opacity = assemble_in[2]; ab.d_rot = assemble_in[3];
    ab.dscaling = assemble_out[0]; ab.drotation = assemble_out[1]; ab.dopacity = assemble_out[2];
    ab.dd_scale = assemble_out[3];
    ab.dd_xyz = assemble_out[4];
    ab.dd_rot = assemble_out[5];
  }
  (void)colors_precomp;
  const float* cov = cov3D_precomp ? cov3D_precomp : reinterpret_cast<const float*>(geom + lay.cov3D);
  {
    ProfScope prof_("preprocess_bwd_kernel", st);
    SKGS_CUDA(launch_pdl(preprocess_bwd_kernel, dim3((rp.P + PB_THREADS - 1) / PB_THREADS), dim3(PB_THREADS), 0, st, rp,
                         means3D, shs, cov3D_precomp ? nullptr : scales, rotations, cov, radii,
                         reinterpret_cast<const uint8_t*>(geom + lay.clamped),
                         reinterpret_cast<float*>(geom + lay.geom_grads), bwd_ticket, dL_dmeans3D, dL_dmeans2D, dL_dsh,
                         dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, ab));
    SKGS_CHECK_LAUNCH("preprocess_bwd_kernel");
  }
  return SKGS_OK;
}

}  // namespace skgs
