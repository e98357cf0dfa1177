/*
 * oracle/raster_oracle.c - CPU restatement of the SK_GS / 3DGS tile rasterizer (forward + backward).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in sk_gs_b200/ (the product) may import, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg do (as the checker /
 * the CPU baseline, never as the thing shipped).
 *
 * Parity status: "parity pinned where the reference allows it" - the reference holds NO golden vectors for this
 * path and its rasterizer has no CPU implementation.  This restatement is pinned by
 *   (1) tests/golden/ npz files generated from the reference's own pure-torch helpers (SH eval, cov3D, cov2D, quaternion->R,
 *       skeleton_warp, find_root) imported from /root/reference by tests/golden/make_golden.py, and
 *   (2) on the GPU box, the reference's own CUDA extension compiled from /root/reference into oracle/_ref/
 *       (tests/test_gpu_reference_ext.py) - tolerance-level, because nvcc's FMA contraction of the reference source
 *       is not reproducible on a CPU.
 *
 * Each function cites the reference lines (relative to /root/reference/my_ext/_C/) it restates:
 *   preprocess fwd : src/nerf/gaussian_preprocess_colmap.cu:26-224, include/gaussian_render.h:42-47,
 *                    SH colour src/nerf/gaussian_rasterizer_forward.cu:97-137
 *   binning        : src/nerf/gaussian_rasterizer_forward.cu:30-94,203-241
 *   composite fwd  : src/nerf/gaussian_render.cu:16-112 (+ upstream bg / depth / alpha outputs)
 *   composite bwd  : src/nerf/gaussian_render.cu:182-341 (+ bg, depth, alpha terms)
 *   preprocess bwd : src/nerf/gaussian_preprocess_colmap.cu:240-481, src/nerf/gaussian_rasterizer_backwrad.cu:26-127
 *
 * Arithmetic contract (what makes radii / keys / ranges / images reproducible bit-for-bit on a GPU):
 *   - compiled with -ffp-contract=off: every fp32 '*' and '+' below is a separately rounded IEEE operation,
 *     evaluated left-to-right exactly as written; fused multiply-adds appear ONLY as explicit fmaf().
 *   - division and sqrtf are IEEE correctly rounded; no rsqrt, no fast-math.
 *   - ndc2Pix is evaluated in double like the reference (its literals are double, :26).
 *   - exp() inside compositing is the fully specified orc_exp() below (range reduction + degree-6 polynomial),
 *     because libm/CUDA expf are not bit-reproducible across CPU and GPU and a 1-ulp difference can flip the
 *     alpha < 1/255 or T < 1e-4 tests.
 *   - backward sums over pixels are accumulated in double (order-insensitive reference values).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16
#define ORC_API __attribute__((visibility("default")))
/* the file is compiled with -mfma (oracle/Makefile): fmaf() below is one vfmadd instruction */
#define ORC_HOT

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- fully specified exp for x <= 0 (see header).  2^t, t = x*log2(e); n = rint(t) by the 1.5*2^23 trick;
 * f = t - n exact; 2^f by a degree-6 minimax polynomial (Horner, explicit fmaf); scale by adding n to the exponent. */
static inline float orc_exp(float x) {
  float t = x * 1.4426950408889634f;
  t = fmaxf(t, -120.0f);
  float r = t + 12582912.0f;
  float nf = r - 12582912.0f;
  float f = t - nf;
  float p = 0.00015345810970757157f;
  p = fmaf(p, f, 0.0013399930903688073f);
  p = fmaf(p, f, 0.009618489071726799f);
  p = fmaf(p, f, 0.05550328642129898f);
  p = fmaf(p, f, 0.24022646248340607f);
  p = fmaf(p, f, 0.6931471824645996f);
  p = fmaf(p, f, 1.0f);
  uint32_t rb, pb;
  memcpy(&rb, &r, 4);
  memcpy(&pb, &p, 4);
  pb += rb << 23; /* low 9 bits of bits(1.5*2^23) are zero, so (rb<<23) == (n<<23) mod 2^32 */
  memcpy(&p, &pb, 4);
  return p;
}
ORC_API float orc_exp_scalar(float x) { return orc_exp(x); }

/* reference: gaussian_rasterizer_forward.cu:30-42 */
ORC_API uint32_t orc_higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb)
      msb += step;
    else
      msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

/* reference: include/gaussian_render.h:42-47 (integer truncation toward zero, clamp to the tile grid) */
static inline void get_rect(float px, float py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
  int a;
  a = (int)((px - (float)max_radius) / (float)TILE);
  *x0 = a < 0 ? 0 : (a > gx ? gx : a);
  a = (int)((py - (float)max_radius) / (float)TILE);
  *y0 = a < 0 ? 0 : (a > gy ? gy : a);
  a = (int)((px + (float)max_radius + (float)(TILE - 1)) / (float)TILE);
  *x1 = a < 0 ? 0 : (a > gx ? gx : a);
  a = (int)((py + (float)max_radius + (float)(TILE - 1)) / (float)TILE);
  *y1 = a < 0 ? 0 : (a > gy ? gy : a);
}

/* rotation matrix rows from an (unnormalised, as in the reference :129) quaternion; R[r][c] standard (row, col) */
static inline void quat_to_R(const float* q, int wxyz, float R[3][3]) {
  float r, x, y, z;
  if (wxyz) {
    r = q[0]; x = q[1]; y = q[2]; z = q[3];
  } else {
    x = q[0]; y = q[1]; z = q[2]; r = q[3];
  }
  R[0][0] = 1.f - 2.f * (y * y + z * z);
  R[0][1] = 2.f * (x * y - r * z);
  R[0][2] = 2.f * (x * z + r * y);
  R[1][0] = 2.f * (x * y + r * z);
  R[1][1] = 1.f - 2.f * (x * x + z * z);
  R[1][2] = 2.f * (y * z - r * x);
  R[2][0] = 2.f * (x * z - r * y);
  R[2][1] = 2.f * (y * z + r * x);
  R[2][2] = 1.f - 2.f * (x * x + y * y);
}

/* Sigma = R diag(s^2) R^T = sum_k (s_k r_k)(s_k r_k)^T, r_k = k-th column of R.  reference :121-152 */
static inline void cov3d_from_scale_rot(const float* scale, float mod, const float* q, int wxyz, float* cov6) {
  float R[3][3];
  quat_to_R(q, wxyz, R);
  float s[3] = {mod * scale[0], mod * scale[1], mod * scale[2]};
  float Mm[3][3]; /* Mm[k][a] = s_k * R[a][k] */
  for (int k = 0; k < 3; k++)
    for (int a = 0; a < 3; a++) Mm[k][a] = s[k] * R[a][k];
  cov6[0] = Mm[0][0] * Mm[0][0] + Mm[1][0] * Mm[1][0] + Mm[2][0] * Mm[2][0];
  cov6[1] = Mm[0][0] * Mm[0][1] + Mm[1][0] * Mm[1][1] + Mm[2][0] * Mm[2][1];
  cov6[2] = Mm[0][0] * Mm[0][2] + Mm[1][0] * Mm[1][2] + Mm[2][0] * Mm[2][2];
  cov6[3] = Mm[0][1] * Mm[0][1] + Mm[1][1] * Mm[1][1] + Mm[2][1] * Mm[2][1];
  cov6[4] = Mm[0][1] * Mm[0][2] + Mm[1][1] * Mm[1][2] + Mm[2][1] * Mm[2][2];
  cov6[5] = Mm[0][2] * Mm[0][2] + Mm[1][2] * Mm[1][2] + Mm[2][2] * Mm[2][2];
}

/* rows a0, a1 of A = J * Rv (2x3), with the 1.3*tanfov clamp of t.xy.  reference :85-105 */
static inline void ewa_rows(float tx, float ty, float tz, float fx, float fy, float tanfovx, float tanfovy,
                            const float* V, float a0[3], float a1[3], float* txc, float* tyc) {
  const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
  const float txtz = tx / tz, tytz = ty / tz;
  tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
  ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
  const float j00 = fx / tz, j02 = -(fx * tx) / (tz * tz);
  const float j11 = fy / tz, j12 = -(fy * ty) / (tz * tz);
  /* Rv row r = (V[r], V[4+r], V[8+r]) */
  a0[0] = j00 * V[0] + j02 * V[2];
  a0[1] = j00 * V[4] + j02 * V[6];
  a0[2] = j00 * V[8] + j02 * V[10];
  a1[0] = j11 * V[1] + j12 * V[2];
  a1[1] = j11 * V[5] + j12 * V[6];
  a1[2] = j11 * V[9] + j12 * V[10];
  *txc = tx;
  *tyc = ty;
}

static inline void sym3_mul(const float* c, const float* a, float* u) {
  u[0] = c[0] * a[0] + c[1] * a[1] + c[2] * a[2];
  u[1] = c[1] * a[0] + c[3] * a[1] + c[4] * a[2];
  u[2] = c[2] * a[0] + c[4] * a[1] + c[5] * a[2];
}

/* SH -> RGB.  reference gaussian_rasterizer_forward.cu:97-137.  sh: [M][3] for this Gaussian */
static inline void sh_to_rgb(int deg, const float* sh, float dirx, float diry, float dirz, float rgb[3],
                             uint8_t clamped[3]) {
  float len = sqrtf(dirx * dirx + diry * diry + dirz * dirz);
  float x = dirx / len, y = diry / len, z = dirz / len;
  for (int c = 0; c < 3; c++) {
    float res = SH_C0 * sh[0 * 3 + c];
    if (deg > 0) {
      res = res - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        res = res + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
              SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
              SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
        if (deg > 2) {
          res = res + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
                SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] + SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] +
                SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
        }
      }
    }
    res += 0.5f;
    clamped[c] = res < 0.f;
    rgb[c] = res < 0.f ? 0.f : res;
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * preprocess forward.  reference: preprocessCUDA_colmap, gaussian_preprocess_colmap.cu:155-224
 * quat_wxyz: 1 = upstream diff_gaussian_rasterization convention (r = rot[0]); 0 = in-tree my_ext (r = rot[3], :133)
 * cov3D_precomp / colors_precomp / shs may be NULL ("absent").
 * ---------------------------------------------------------------------------------------------------------- */
ORC_API ORC_HOT void orc_preprocess_fwd(int P, int D, int M, const float* means, const float* scales, float mod,
                                        const float* rots, int quat_wxyz, const float* opac, const float* shs,
                                        const float* cov3D_precomp, const float* colors_precomp, const float* V,
                                        const float* Pm, const float* campos, int W, int H, float tanfovx,
                                        float tanfovy, int32_t* radii, float* means2D, float* depths, float* cov3Ds,
                                        float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
  const float fy = H / (2.0f * tanfovy), fx = W / (2.0f * tanfovx); /* gaussian_rasterizer_forward.cu:163-164 */
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    radii[i] = 0;
    tiles_touched[i] = 0;
    const float x = means[3 * i], y = means[3 * i + 1], z = means[3 * i + 2];
    /* :28-35 (column-major indexing of the transposed torch matrix) */
    const float pvx = V[0] * x + V[4] * y + V[8] * z + V[12];
    const float pvy = V[1] * x + V[5] * y + V[9] * z + V[13];
    const float pvz = V[2] * x + V[6] * y + V[10] * z + V[14];
    if (pvz <= 0.2f) continue; /* :73 */
    const float hx = Pm[0] * x + Pm[4] * y + Pm[8] * z + Pm[12];
    const float hy = Pm[1] * x + Pm[5] * y + Pm[9] * z + Pm[13];
    const float hw = Pm[3] * x + Pm[7] * y + Pm[11] * z + Pm[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ppx = hx * pw, ppy = hy * pw;
    const float* cov3D;
    if (cov3D_precomp) {
      cov3D = cov3D_precomp + 6 * i;
    } else {
      cov3d_from_scale_rot(scales + 3 * i, mod, rots + 4 * i, quat_wxyz, cov3Ds + 6 * i);
      cov3D = cov3Ds + 6 * i;
    }
    float a0[3], a1[3], u0[3], u1[3], txc, tyc;
    ewa_rows(pvx, pvy, pvz, fx, fy, tanfovx, tanfovy, V, a0, a1, &txc, &tyc);
    sym3_mul(cov3D, a0, u0);
    sym3_mul(cov3D, a1, u1);
    const float c00 = (a0[0] * u0[0] + a0[1] * u0[1] + a0[2] * u0[2]) + 0.3f; /* :113-114 */
    const float c01 = a0[0] * u1[0] + a0[1] * u1[1] + a0[2] * u1[2];
    const float c11 = (a1[0] * u1[0] + a1[1] * u1[1] + a1[2] * u1[2]) + 0.3f;
    const float det = c00 * c11 - c01 * c01;
    if (det == 0.0f) continue; /* :193 */
    const float det_inv = 1.f / det;
    const float con_x = c11 * det_inv, con_y = -c01 * det_inv, con_z = c00 * det_inv;
    const float mid = 0.5f * (c00 + c11);
    const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda1 = mid + disc, lambda2 = mid - disc;
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    /* ndc2Pix in double, :26 */
    const float pix_x = (float)((((double)ppx + 1.0) * (double)W - 1.0) * 0.5);
    const float pix_y = (float)((((double)ppy + 1.0) * (double)H - 1.0) * 0.5);
    int x0, y0, x1, y1;
    int irad = my_radius > 2.0e9f ? 2000000000 : (int)my_radius;
    get_rect(pix_x, pix_y, irad, gx, gy, &x0, &y0, &x1, &y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    if (colors_precomp == NULL) {
      sh_to_rgb(D, shs + (size_t)i * M * 3, x - campos[0], y - campos[1], z - campos[2], rgb + 3 * i,
                clamped + 3 * i);
    }
    depths[i] = pvz;
    radii[i] = irad;
    means2D[2 * i] = pix_x;
    means2D[2 * i + 1] = pix_y;
    conic_opacity[4 * i] = con_x;
    conic_opacity[4 * i + 1] = con_y;
    conic_opacity[4 * i + 2] = con_z;
    conic_opacity[4 * i + 3] = opac[i];
    tiles_touched[i] = (uint32_t)((y1 - y0) * (x1 - x0));
  }
}

/* ---- thin single-formula entry points (used by tests/test_oracle_golden.py against the reference's torch helpers) ---- */
ORC_API void orc_sh_to_rgb(int n, int deg, int M, const float* shs, const float* dirs, float* rgb, uint8_t* clamped) {
  for (int i = 0; i < n; i++)
    sh_to_rgb(deg, shs + (size_t)i * M * 3, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], rgb + 3 * i, clamped + 3 * i);
}
ORC_API void orc_cov3d(int n, const float* scales, float mod, const float* rots, int quat_wxyz, float* cov6) {
  for (int i = 0; i < n; i++) cov3d_from_scale_rot(scales + 3 * i, mod, rots + 4 * i, quat_wxyz, cov6 + 6 * i);
}
/* cov2D = (J Rv) Sigma (J Rv)^T + 0.3 I, returns (c00, c01, c11) per point */
ORC_API void orc_cov2d(int n, const float* means, const float* cov6, const float* V, float fx, float fy, float tanfovx,
                       float tanfovy, float* out3) {
  for (int i = 0; i < n; i++) {
    const float x = means[3 * i], y = means[3 * i + 1], z = means[3 * i + 2];
    const float pvx = V[0] * x + V[4] * y + V[8] * z + V[12];
    const float pvy = V[1] * x + V[5] * y + V[9] * z + V[13];
    const float pvz = V[2] * x + V[6] * y + V[10] * z + V[14];
    float a0[3], a1[3], u0[3], u1[3], txc, tyc;
    ewa_rows(pvx, pvy, pvz, fx, fy, tanfovx, tanfovy, V, a0, a1, &txc, &tyc);
    sym3_mul(cov6 + 6 * i, a0, u0);
    sym3_mul(cov6 + 6 * i, a1, u1);
    out3[3 * i] = (a0[0] * u0[0] + a0[1] * u0[1] + a0[2] * u0[2]) + 0.3f;
    out3[3 * i + 1] = a0[0] * u1[0] + a0[1] * u1[1] + a0[2] * u1[2];
    out3[3 * i + 2] = (a1[0] * u1[0] + a1[1] * u1[1] + a1[2] * u1[2]) + 0.3f;
  }
}

/* inclusive scan of tiles_touched; returns R.  reference gaussian_rasterizer_forward.cu:203-209 */
ORC_API uint64_t orc_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets) {
  uint64_t s = 0;
  for (int i = 0; i < P; i++) {
    s += tiles_touched[i];
    offsets[i] = (uint32_t)s;
  }
  return s;
}

/* stable LSD radix sort of (key,val) pairs on key bits [0,end_bit), 16-bit digits */
static void radix_sort_pairs(uint64_t n, uint64_t* k_in, uint32_t* v_in, uint64_t* k_out, uint32_t* v_out,
                             int end_bit) {
  uint64_t *ka = k_in, *kb = k_out;
  uint32_t *va = v_in, *vb = v_out;
  uint64_t* hist = (uint64_t*)malloc(65536 * sizeof(uint64_t));
  int passes = 0;
  for (int shift = 0; shift < end_bit; shift += 16, passes++) {
    int bits = end_bit - shift < 16 ? end_bit - shift : 16;
    uint64_t mask = ((uint64_t)1 << bits) - 1;
    memset(hist, 0, 65536 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) hist[(ka[i] >> shift) & mask]++;
    uint64_t s = 0;
    for (uint64_t d = 0; d <= mask; d++) {
      uint64_t c = hist[d];
      hist[d] = s;
      s += c;
    }
    for (uint64_t i = 0; i < n; i++) {
      uint64_t p = hist[(ka[i] >> shift) & mask]++;
      kb[p] = ka[i];
      vb[p] = va[i];
    }
    uint64_t* tk = ka; ka = kb; kb = tk;
    uint32_t* tv = va; va = vb; vb = tv;
  }
  free(hist);
  if (ka != k_out) {
    memcpy(k_out, ka, n * sizeof(uint64_t));
    memcpy(v_out, va, n * sizeof(uint32_t));
  }
  (void)passes;
}

/* duplicateWithKeys + SortPairs + identifyTileRanges.  reference gaussian_rasterizer_forward.cu:45-94,219-241.
 * keys_unsorted/vals_unsorted are scratch+output (the emission order), keys/vals the sorted lists, ranges [tiles][2]. */
ORC_API void orc_binning(int P, uint64_t R, const float* means2D, const float* depths, const int32_t* radii,
                         const uint32_t* offsets, int W, int H, uint64_t* keys_unsorted, uint32_t* vals_unsorted,
                         uint64_t* keys, uint32_t* vals, uint32_t* ranges) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  for (int i = 0; i < P; i++) {
    if (radii[i] > 0) {
      uint64_t off = (i == 0) ? 0 : offsets[i - 1];
      int x0, y0, x1, y1;
      get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
      uint32_t dbits;
      memcpy(&dbits, &depths[i], 4);
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          uint64_t key = (uint64_t)(y * gx + x);
          key <<= 32;
          key |= dbits;
          keys_unsorted[off] = key;
          vals_unsorted[off] = (uint32_t)i;
          off++;
        }
    }
  }
  memset(ranges, 0, (size_t)gx * gy * 2 * sizeof(uint32_t));
  if (R == 0) return;
  /* sort a copy so that the unsorted emission order stays available to the tests */
  uint64_t* ktmp = (uint64_t*)malloc(R * sizeof(uint64_t));
  uint32_t* vtmp = (uint32_t*)malloc(R * sizeof(uint32_t));
  memcpy(ktmp, keys_unsorted, R * sizeof(uint64_t));
  memcpy(vtmp, vals_unsorted, R * sizeof(uint32_t));
  int bit = (int)orc_higher_msb((uint32_t)(gx * gy));
  radix_sort_pairs(R, ktmp, vtmp, keys, vals, 32 + bit);
  free(ktmp);
  free(vtmp);
  for (uint64_t idx = 0; idx < R; idx++) {
    uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0)
      ranges[2 * cur] = 0;
    else {
      uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
      if (cur != prev) {
        ranges[2 * prev + 1] = (uint32_t)idx;
        ranges[2 * cur] = (uint32_t)idx;
      }
    }
    if (idx == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
  }
}

/* the fully specified per-(pixel,Gaussian) falloff exponent:
 *   power = -1/2 (A dx^2 + C dy^2) - B dx dy  evaluated as  fma(dx, fma(A', dx, B'*dy), (C'*dy)*dy)
 * with A' = -0.5*A, B' = -B, C' = -0.5*C (all exact).  reference gaussian_render.cu:75-78 */
static inline float pair_power(float gx, float gy, float A, float B, float C, float px, float py, float* dx_,
                               float* dy_) {
  const float dx = gx - px, dy = gy - py;
  const float Ap = -0.5f * A, Bp = -B, Cp = -0.5f * C;
  *dx_ = dx;
  *dy_ = dy;
  return fmaf(dx, fmaf(Ap, dx, Bp * dy), (Cp * dy) * dy);
}

/* ------------------------------------------------------------------------------------------------------------
 * composite forward.  reference renderCUDA_forward gaussian_render.cu:16-112 with the upstream-contract deltas of
 * out_color = C + T*bg, out_depth = sum z a T, out_alpha = 1 - T, final_T kept exactly.
 * ---------------------------------------------------------------------------------------------------------- */
ORC_API ORC_HOT void orc_composite_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                                       const float* means2D, const float* conic_opacity, const float* colors,
                                       const float* depths, const float* bg, float* out_color, float* out_depth,
                                       float* out_alpha, uint32_t* n_contrib, float* final_T) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
  for (int tile = 0; tile < gx * gy; tile++) {
    const int tx = tile % gx, ty = tile / gx;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    for (int ly = 0; ly < TILE; ly++)
      for (int lx = 0; lx < TILE; lx++) {
        const int px = tx * TILE + lx, py = ty * TILE + ly;
        if (px >= W || py >= H) continue;
        const float pxf = (float)px, pyf = (float)py;
        float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dp = 0.f;
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = r0; k < r1; k++) {
          contributor++;
          const uint32_t g = point_list[k];
          float dx, dy;
          const float power = pair_power(means2D[2 * g], means2D[2 * g + 1], conic_opacity[4 * g],
                                         conic_opacity[4 * g + 1], conic_opacity[4 * g + 2], pxf, pyf, &dx, &dy);
          if (power > 0.0f) continue;
          const float alpha = fminf(0.99f, conic_opacity[4 * g + 3] * orc_exp(power));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = T * (1.0f - alpha);
          if (test_T < 0.0001f) break;
          const float w = alpha * T;
          C0 = fmaf(colors[3 * g], w, C0);
          C1 = fmaf(colors[3 * g + 1], w, C1);
          C2 = fmaf(colors[3 * g + 2], w, C2);
          Dp = fmaf(depths[g], w, Dp);
          T = test_T;
          last = contributor;
        }
        const size_t pid = (size_t)py * W + px;
        out_color[pid] = fmaf(T, bg[0], C0);
        out_color[(size_t)H * W + pid] = fmaf(T, bg[1], C1);
        out_color[2 * (size_t)H * W + pid] = fmaf(T, bg[2], C2);
        out_depth[pid] = Dp;
        out_alpha[pid] = 1.0f - T;
        n_contrib[pid] = last;
        final_T[pid] = T;
      }
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * composite backward.  reference renderCUDA_backward gaussian_render.cu:182-341 + bg / depth / alpha terms.
 * Outputs (accumulated in double, written as float): dL_dmean2D [P][3] (z = 0), dL_dconic [P][4] (slots x,y,w used,
 * :333-335), dL_dopacity [P], dL_dcolors [P][3], dL_dz [P] (gradient of the depth output w.r.t. per-Gaussian depth).
 * ---------------------------------------------------------------------------------------------------------- */
ORC_API ORC_HOT void orc_composite_bwd(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                                       const float* means2D, const float* conic_opacity, const float* colors,
                                       const float* depths, const float* bg, const uint32_t* n_contrib,
                                       const float* final_T, const float* dL_dpix, const float* dL_ddepth,
                                       const float* dL_dalpha_map, float* dL_dmean2D, float* dL_dconic,
                                       float* dL_dopacity, float* dL_dcolors, float* dL_dz) {
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int NG = 10;
  double* acc = (double*)calloc((size_t)P * NG, sizeof(double));
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel
  {
    double* loc = NULL;
    size_t loc_cap = 0;
#pragma omp for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
      const int tx = tile % gx, ty = tile / gx;
      const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
      const size_t n = r1 - r0;
      if (n == 0) continue;
      if (n * NG > loc_cap) {
        free(loc);
        loc_cap = n * NG;
        loc = (double*)malloc(loc_cap * sizeof(double));
      }
      memset(loc, 0, n * NG * sizeof(double));
      for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
          const int px = tx * TILE + lx, py = ty * TILE + ly;
          if (px >= W || py >= H) continue;
          const size_t pid = (size_t)py * W + px;
          const float pxf = (float)px, pyf = (float)py;
          const float T_final = final_T[pid];
          float T = T_final;
          const float dpix[3] = {dL_dpix[pid], dL_dpix[(size_t)H * W + pid], dL_dpix[2 * (size_t)H * W + pid]};
          const float dD = dL_ddepth ? dL_ddepth[pid] : 0.f;
          const float dA = dL_dalpha_map ? dL_dalpha_map[pid] : 0.f;
          const float bg_dot = bg[0] * dpix[0] + bg[1] * dpix[1] + bg[2] * dpix[2];
          const float tail = bg_dot - dA; /* dL/dT_final */
          const uint32_t last = n_contrib[pid];
          float accum[3] = {0, 0, 0}, accum_d = 0.f, last_alpha = 0.f, last_color[3] = {0, 0, 0}, last_depth = 0.f;
          for (uint32_t kk = last; kk-- > 0;) {
            const uint32_t g = point_list[r0 + kk];
            float dx, dy;
            const float power = pair_power(means2D[2 * g], means2D[2 * g + 1], conic_opacity[4 * g],
                                           conic_opacity[4 * g + 1], conic_opacity[4 * g + 2], pxf, pyf, &dx, &dy);
            if (power > 0.0f) continue;
            const float o = conic_opacity[4 * g + 3];
            const float G = orc_exp(power);
            const float alpha = fminf(0.99f, o * G);
            if (alpha < 1.0f / 255.0f) continue;
            T = T / (1.f - alpha);
            const float dchannel_dcolor = alpha * T;
            float dL_dalpha = 0.f;
            double* l = loc + (size_t)kk * NG;
            for (int ch = 0; ch < 3; ch++) {
              const float c = colors[3 * g + ch];
              accum[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum[ch];
              last_color[ch] = c;
              dL_dalpha += (c - accum[ch]) * dpix[ch];
              l[6 + ch] += (double)(dchannel_dcolor * dpix[ch]);
            }
            {
              const float zc = depths[g];
              accum_d = last_alpha * last_depth + (1.f - last_alpha) * accum_d;
              last_depth = zc;
              dL_dalpha += (zc - accum_d) * dD;
              l[9] += (double)(dchannel_dcolor * dD);
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * tail;
            const float dL_dG = o * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float A = conic_opacity[4 * g], B = conic_opacity[4 * g + 1], Cc = conic_opacity[4 * g + 2];
            const float dG_ddelx = -gdx * A - gdy * B;
            const float dG_ddely = -gdy * Cc - gdx * B;
            l[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
            l[1] += (double)(dL_dG * dG_ddely * ddely_dy);
            l[2] += (double)(-0.5f * gdx * dx * dL_dG);
            l[3] += (double)(-0.5f * gdx * dy * dL_dG);
            l[4] += (double)(-0.5f * gdy * dy * dL_dG);
            l[5] += (double)(G * dL_dalpha);
          }
        }
#pragma omp critical
      {
        for (size_t k = 0; k < n; k++) {
          const uint32_t g = point_list[r0 + k];
          for (int c = 0; c < NG; c++) acc[(size_t)g * NG + c] += loc[k * NG + c];
        }
      }
    }
    free(loc);
  }
  for (int i = 0; i < P; i++) {
    const double* a = acc + (size_t)i * NG;
    dL_dmean2D[3 * i] = (float)a[0];
    dL_dmean2D[3 * i + 1] = (float)a[1];
    dL_dmean2D[3 * i + 2] = 0.f;
    dL_dconic[4 * i] = (float)a[2];
    dL_dconic[4 * i + 1] = (float)a[3];
    dL_dconic[4 * i + 2] = 0.f;
    dL_dconic[4 * i + 3] = (float)a[4];
    dL_dopacity[i] = (float)a[5];
    dL_dcolors[3 * i] = (float)a[6];
    dL_dcolors[3 * i + 1] = (float)a[7];
    dL_dcolors[3 * i + 2] = (float)a[8];
    dL_dz[i] = (float)a[9];
  }
  free(acc);
}

/* ------------------------------------------------------------------------------------------------------------
 * preprocess backward.  reference computeCov2DCUDA_colmap :240-354, preprocessCUDA_backward_colmap :424-461,
 * computeCov3D_colmap bwd :357-420, SH bwd gaussian_rasterizer_backwrad.cu:26-127, dnormvdv gaussian_render.h:56-65.
 * Inputs: per-Gaussian grads from composite bwd.  Outputs zero for radii <= 0 (:244,430).
 * dL_dcov3D [P][6] is always written (it is the returned gradient when cov3D_precomp is given).
 * ---------------------------------------------------------------------------------------------------------- */
ORC_API void orc_preprocess_bwd(int P, int D, int M, const float* means, const int32_t* radii, const float* shs,
                                const uint8_t* clamped, const float* scales, const float* rots, int quat_wxyz,
                                float mod, const float* cov3Ds, const float* V, const float* Pm, const float* campos,
                                int W, int H, float tanfovx, float tanfovy, const float* dL_dmean2D,
                                const float* dL_dconic, const float* dL_dcolors, const float* dL_dz,
                                float* dL_dmeans, float* dL_dcov3D, float* dL_dsh, float* dL_dscales,
                                float* dL_drots) {
  const float fy = H / (2.0f * tanfovy), fx = W / (2.0f * tanfovx);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < P; i++) {
    for (int k = 0; k < 3; k++) dL_dmeans[3 * i + k] = 0.f;
    for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.f;
    if (dL_dsh)
      for (int k = 0; k < 3 * M; k++) dL_dsh[(size_t)i * 3 * M + k] = 0.f;
    if (dL_dscales)
      for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = 0.f;
    if (dL_drots)
      for (int k = 0; k < 4; k++) dL_drots[4 * i + k] = 0.f;
    if (!(radii[i] > 0)) continue;
    const float mx = means[3 * i], my = means[3 * i + 1], mz = means[3 * i + 2];
    const float* cov3D = cov3Ds + 6 * i;
    /* ---- cov2D backward (:240-354) ---- */
    const float pvx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
    const float pvy = V[1] * mx + V[5] * my + V[9] * mz + V[13];
    const float pvz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = pvx / pvz, tytz = pvy / pvz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    float a0[3], a1[3], u0[3], u1[3], tx, ty;
    ewa_rows(pvx, pvy, pvz, fx, fy, tanfovx, tanfovy, V, a0, a1, &tx, &ty);
    const float tz = pvz;
    sym3_mul(cov3D, a0, u0);
    sym3_mul(cov3D, a1, u1);
    const float a = (a0[0] * u0[0] + a0[1] * u0[1] + a0[2] * u0[2]) + 0.3f;
    const float b = a0[0] * u1[0] + a0[1] * u1[1] + a0[2] * u1[2];
    const float c = (a1[0] * u1[0] + a1[1] * u1[1] + a1[2] * u1[2]) + 0.3f;
    const float denom = a * c - b * b;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    const float dcx = dL_dconic[4 * i], dcy = dL_dconic[4 * i + 1], dcz = dL_dconic[4 * i + 3];
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    if (denom2inv != 0) {
      dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
      dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
      dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
      /* T[0][k] = a0[k], T[1][k] = a1[k] in the reference's glm indexing (:296-309) */
      dL_dcov3D[6 * i + 0] = a0[0] * a0[0] * dL_da + a0[0] * a1[0] * dL_db + a1[0] * a1[0] * dL_dc;
      dL_dcov3D[6 * i + 3] = a0[1] * a0[1] * dL_da + a0[1] * a1[1] * dL_db + a1[1] * a1[1] * dL_dc;
      dL_dcov3D[6 * i + 5] = a0[2] * a0[2] * dL_da + a0[2] * a1[2] * dL_db + a1[2] * a1[2] * dL_dc;
      dL_dcov3D[6 * i + 1] =
          2 * a0[0] * a0[1] * dL_da + (a0[0] * a1[1] + a0[1] * a1[0]) * dL_db + 2 * a1[0] * a1[1] * dL_dc;
      dL_dcov3D[6 * i + 2] =
          2 * a0[0] * a0[2] * dL_da + (a0[0] * a1[2] + a0[2] * a1[0]) * dL_db + 2 * a1[0] * a1[2] * dL_dc;
      dL_dcov3D[6 * i + 4] =
          2 * a0[2] * a0[1] * dL_da + (a0[1] * a1[2] + a0[2] * a1[1]) * dL_db + 2 * a1[1] * a1[2] * dL_dc;
    }
    /* dL/dT rows (:316-327): dL_dT0 = 2*(Sigma a0) dL_da + (Sigma a1) dL_db ; dL_dT1 = 2*(Sigma a1) dL_dc + (Sigma a0) dL_db */
    float dT0[3], dT1[3];
    for (int k = 0; k < 3; k++) {
      dT0[k] = 2 * u0[k] * dL_da + u1[k] * dL_db;
      dT1[k] = 2 * u1[k] * dL_dc + u0[k] * dL_db;
    }
    /* W[k][:] in glm = column k of W_glm = row k of Rv = (V[k], V[4+k], V[8+k]) (:331-334) */
    const float dL_dJ00 = V[0] * dT0[0] + V[4] * dT0[1] + V[8] * dT0[2];
    const float dL_dJ02 = V[2] * dT0[0] + V[6] * dT0[1] + V[10] * dT0[2];
    const float dL_dJ11 = V[1] * dT1[0] + V[5] * dT1[1] + V[9] * dT1[2];
    const float dL_dJ12 = V[2] * dT1[0] + V[6] * dT1[1] + V[10] * dT1[2];
    const float tzi = 1.f / tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
    const float dL_dtx = x_grad_mul * -fx * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -fy * tz2 * dL_dJ12;
    const float dL_dtz = -fx * tz2 * dL_dJ00 - fy * tz2 * dL_dJ11 + (2 * fx * tx) * tz3 * dL_dJ02 +
                         (2 * fy * ty) * tz3 * dL_dJ12;
    /* transformVec4x3Transpose (:54-61) */
    float dmx = V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz;
    float dmy = V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz;
    float dmz = V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz;
    /* ---- mean2D -> mean3D (:432-452) ---- */
    {
      const float hw = Pm[3] * mx + Pm[7] * my + Pm[11] * mz + Pm[15];
      const float m_w = 1.0f / (hw + 0.0000001f);
      const float mul1 = (Pm[0] * mx + Pm[4] * my + Pm[8] * mz + Pm[12]) * m_w * m_w;
      const float mul2 = (Pm[1] * mx + Pm[5] * my + Pm[9] * mz + Pm[13]) * m_w * m_w;
      const float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
      dmx += (Pm[0] * m_w - Pm[3] * mul1) * g2x + (Pm[1] * m_w - Pm[3] * mul2) * g2y;
      dmy += (Pm[4] * m_w - Pm[7] * mul1) * g2x + (Pm[5] * m_w - Pm[7] * mul2) * g2y;
      dmz += (Pm[8] * m_w - Pm[11] * mul1) * g2x + (Pm[9] * m_w - Pm[11] * mul2) * g2y;
    }
    /* ---- depth output: z = p_view.z, d z / d mean = third row of Rv (upstream depth forks) ---- */
    if (dL_dz) {
      dmx += V[2] * dL_dz[i];
      dmy += V[6] * dL_dz[i];
      dmz += V[10] * dL_dz[i];
    }
    /* ---- SH backward (gaussian_rasterizer_backwrad.cu:26-127) ---- */
    if (shs) {
      const float* sh = shs + (size_t)i * M * 3;
      float* dsh = dL_dsh + (size_t)i * M * 3;
      const float dox = mx - campos[0], doy = my - campos[1], doz = mz - campos[2];
      const float len = sqrtf(dox * dox + doy * doy + doz * doz);
      const float x = dox / len, y = doy / len, z = doz / len;
      float dRGB[3];
      for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolors[3 * i + ch] * (clamped[3 * i + ch] ? 0.f : 1.f);
      float dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
      for (int ch = 0; ch < 3; ch++) {
#define SH(k) sh[(k) * 3 + ch]
#define DSH(k) dsh[(k) * 3 + ch]
        DSH(0) = SH_C0 * dRGB[ch];
        if (D > 0) {
          DSH(1) = (-SH_C1 * y) * dRGB[ch];
          DSH(2) = (SH_C1 * z) * dRGB[ch];
          DSH(3) = (-SH_C1 * x) * dRGB[ch];
          dRGBdx[ch] = -SH_C1 * SH(3);
          dRGBdy[ch] = -SH_C1 * SH(1);
          dRGBdz[ch] = SH_C1 * SH(2);
          if (D > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4) = (SH_C2[0] * xy) * dRGB[ch];
            DSH(5) = (SH_C2[1] * yz) * dRGB[ch];
            DSH(6) = (SH_C2[2] * (2.f * zz - xx - yy)) * dRGB[ch];
            DSH(7) = (SH_C2[3] * xz) * dRGB[ch];
            DSH(8) = (SH_C2[4] * (xx - yy)) * dRGB[ch];
            dRGBdx[ch] += SH_C2[0] * y * SH(4) + SH_C2[2] * 2.f * -x * SH(6) + SH_C2[3] * z * SH(7) +
                          SH_C2[4] * 2.f * x * SH(8);
            dRGBdy[ch] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * 2.f * -y * SH(6) +
                          SH_C2[4] * 2.f * -y * SH(8);
            dRGBdz[ch] += SH_C2[1] * y * SH(5) + SH_C2[2] * 2.f * 2.f * z * SH(6) + SH_C2[3] * x * SH(7);
            if (D > 2) {
              DSH(9) = (SH_C3[0] * y * (3.f * xx - yy)) * dRGB[ch];
              DSH(10) = (SH_C3[1] * xy * z) * dRGB[ch];
              DSH(11) = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dRGB[ch];
              DSH(12) = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB[ch];
              DSH(13) = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dRGB[ch];
              DSH(14) = (SH_C3[5] * z * (xx - yy)) * dRGB[ch];
              DSH(15) = (SH_C3[6] * x * (xx - 3.f * yy)) * dRGB[ch];
              dRGBdx[ch] += (SH_C3[0] * SH(9) * 3.f * 2.f * xy + SH_C3[1] * SH(10) * yz +
                             SH_C3[2] * SH(11) * -2.f * xy + SH_C3[3] * SH(12) * -3.f * 2.f * xz +
                             SH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * SH(14) * 2.f * xz +
                             SH_C3[6] * SH(15) * 3.f * (xx - yy));
              dRGBdy[ch] += (SH_C3[0] * SH(9) * 3.f * (xx - yy) + SH_C3[1] * SH(10) * xz +
                             SH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) + SH_C3[3] * SH(12) * -3.f * 2.f * yz +
                             SH_C3[4] * SH(13) * -2.f * xy + SH_C3[5] * SH(14) * -2.f * yz +
                             SH_C3[6] * SH(15) * -3.f * 2.f * xy);
              dRGBdz[ch] += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * 4.f * 2.f * yz +
                             SH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) + SH_C3[4] * SH(13) * 4.f * 2.f * xz +
                             SH_C3[5] * SH(14) * (xx - yy));
            }
          }
        }
#undef SH
#undef DSH
      }
      const float ddx = dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1] + dRGBdx[2] * dRGB[2];
      const float ddy = dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1] + dRGBdy[2] * dRGB[2];
      const float ddz = dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1] + dRGBdz[2] * dRGB[2];
      /* dnormvdv (gaussian_render.h:56-65) */
      const float sum2 = dox * dox + doy * doy + doz * doz;
      const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
      dmx += ((+sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
      dmy += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
      dmz += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
    }
    dL_dmeans[3 * i] = dmx;
    dL_dmeans[3 * i + 1] = dmy;
    dL_dmeans[3 * i + 2] = dmz;
    /* ---- cov3D -> scale, rotation (:357-420) ---- */
    if (scales && dL_dscales) {
      float R[3][3];
      const float* q = rots + 4 * i;
      quat_to_R(q, quat_wxyz, R);
      float r, x, y, z;
      if (quat_wxyz) {
        r = q[0]; x = q[1]; y = q[2]; z = q[3];
      } else {
        x = q[0]; y = q[1]; z = q[2]; r = q[3];
      }
      const float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
      const float* g = dL_dcov3D + 6 * i;
      /* symmetric dL/dSigma with halved off-diagonals */
      const float dS[3][3] = {{g[0], 0.5f * g[1], 0.5f * g[2]}, {0.5f * g[1], g[3], 0.5f * g[4]},
                              {0.5f * g[2], 0.5f * g[4], g[5]}};
      /* Sigma = sum_k s_k^2 r_k r_k^T (r_k = column k of R):
       *   dL/ds_k    = 2 s_k r_k^T dS r_k         (reference: dot(Rt[k], dL_dMt[k]) with dL_dM = 2 M dS)
       *   dL/dR[a][k] = 2 s_k^2 (dS r_k)[a]       (reference: dL_dMt[k] *= s_k)                         */
      float dR[3][3];
      for (int k = 0; k < 3; k++) {
        float v[3];
        for (int a2 = 0; a2 < 3; a2++) v[a2] = dS[a2][0] * R[0][k] + dS[a2][1] * R[1][k] + dS[a2][2] * R[2][k];
        const float dot = R[0][k] * v[0] + R[1][k] * v[1] + R[2][k] * v[2];
        dL_dscales[3 * i + k] = 2.0f * s[k] * dot;
        for (int a2 = 0; a2 < 3; a2++) dR[a2][k] = 2.0f * s[k] * s[k] * v[a2];
      }
      /* reference's dL_dMt[k][a] (after *= s_k) equals dR[a][k] here; quaternion gradient :406-414 */
#define MT(k, a) dR[a][k]
      const float dqx = 2 * y * (MT(1, 0) + MT(0, 1)) + 2 * z * (MT(2, 0) + MT(0, 2)) + 2 * r * (MT(1, 2) - MT(2, 1)) -
                        4 * x * (MT(2, 2) + MT(1, 1));
      const float dqy = 2 * x * (MT(1, 0) + MT(0, 1)) + 2 * r * (MT(2, 0) - MT(0, 2)) + 2 * z * (MT(1, 2) + MT(2, 1)) -
                        4 * y * (MT(2, 2) + MT(0, 0));
      const float dqz = 2 * r * (MT(0, 1) - MT(1, 0)) + 2 * x * (MT(2, 0) + MT(0, 2)) + 2 * y * (MT(1, 2) + MT(2, 1)) -
                        4 * z * (MT(1, 1) + MT(0, 0));
      const float dqr = 2 * z * (MT(0, 1) - MT(1, 0)) + 2 * y * (MT(2, 0) - MT(0, 2)) + 2 * x * (MT(1, 2) - MT(2, 1));
#undef MT
      if (quat_wxyz) {
        dL_drots[4 * i] = dqr; dL_drots[4 * i + 1] = dqx; dL_drots[4 * i + 2] = dqy; dL_drots[4 * i + 3] = dqz;
      } else {
        dL_drots[4 * i] = dqx; dL_drots[4 * i + 1] = dqy; dL_drots[4 * i + 2] = dqz; dL_drots[4 * i + 3] = dqr;
      }
      /* NB the reference returns dL/d(mod*scale) as dL/dscale, i.e. it omits the mod factor (:397-399); kept. */
    }
  }
}
