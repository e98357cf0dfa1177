// api.cu - C ABI glue of libskgs_b200.so: error reporting, arena layout, rasterizer entry points.
// See include/skgs_b200.h for the contract and the reference interfaces each entry point replaces.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace skgs {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SKGS_PDL");
    on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return on == 1;
}

int tile_cell_stride() {  // ints between two cells of the tile difference grid (SKGS_CELL_STRIDE; tuning)
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("SKGS_CELL_STRIDE");
    v = e ? atoi(e) : 32;  // one 128-byte line per cell: same-line atomics serialise in L2 (A/B: 68.8 -> 40.9 us)
    if (v < 1) v = 1;
  }
  return v;
}

// ---- optional per-kernel timing -------------------------------------------------------------------------------
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
bool prof_enabled() { return g_prof_on; }
void prof_begin(const char* name, cudaStream_t st) {
  ProfRec r{name, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, st);
}

int sort_passes(int gx, int gy);

constexpr size_t ARENA_ALIGN = 256;

static int fill_params(const skgs_raster_settings* s, int P, int M, RasterParams& rp) {
  SKGS_CHECK_ARG(s != nullptr, "settings is NULL");
  SKGS_CHECK_ARG(s->image_width > 0 && s->image_height > 0, "image size must be positive (got %d x %d)",
                 s->image_width, s->image_height);
  SKGS_CHECK_ARG(s->viewmatrix && s->projmatrix && s->campos, "viewmatrix / projmatrix / campos must be device pointers");
  SKGS_CHECK_ARG(s->sh_degree >= 0 && s->sh_degree <= 3, "sh_degree %d not in 0..3", s->sh_degree);
  SKGS_CHECK_ARG(P >= 0, "P < 0");
  rp.P = P;
  rp.M = M;
  rp.D = s->sh_degree;
  rp.W = s->image_width;
  rp.H = s->image_height;
  rp.gx = (rp.W + TILE - 1) / TILE;
  rp.gy = (rp.H + TILE - 1) / TILE;
  rp.tanfovx = s->tanfovx;
  rp.tanfovy = s->tanfovy;
  rp.fy = rp.H / (2.0f * s->tanfovy);  // gaussian_rasterizer_forward.cu:163-164
  rp.fx = rp.W / (2.0f * s->tanfovx);
  rp.mod = s->scale_modifier;
  rp.quat_wxyz = s->quat_wxyz;
  rp.view = s->viewmatrix;
  rp.proj = s->projmatrix;
  rp.campos = s->campos;
  rp.bg = s->bg;
  return SKGS_OK;
}

static int check_inputs(const skgs_raster_settings* s, int P, int M, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        const float* rotations, const float* cov3D_precomp) {
  if (P == 0) return SKGS_OK;
  SKGS_CHECK_ARG(means3D != nullptr, "means3D is NULL");
  SKGS_CHECK_ARG(opacities != nullptr, "opacities is NULL");
  // same rule as networks/renderer/gaussian_render.py:250-255
  SKGS_CHECK_ARG((shs == nullptr) != (colors_precomp == nullptr),
                 "Please provide excatly one of either SHs or precomputed colors!");
  SKGS_CHECK_ARG(((scales == nullptr && rotations == nullptr) && cov3D_precomp != nullptr) ||
                     ((scales != nullptr && rotations != nullptr) && cov3D_precomp == nullptr),
                 "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
  if (shs) {
    SKGS_CHECK_ARG(M >= (s->sh_degree + 1) * (s->sh_degree + 1) && M <= 16,
                   "sh has %d coefficients, degree %d needs %d (max 16)", M, s->sh_degree,
                   (s->sh_degree + 1) * (s->sh_degree + 1));
  }
  return SKGS_OK;
}

}  // namespace skgs

using namespace skgs;

extern "C" {

const char* skgs_last_error(void) { return g_err; }
int skgs_abi_version(void) { return 4; }
int skgs_built_for_sm(void) { return 100; }
uint64_t skgs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void skgs_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on = on != 0;
}

int skgs_profile_collect(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<int, double>> agg;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    auto it = agg.find(r.name);
    if (it == agg.end()) {
      order.push_back(r.name);
      agg[r.name] = {1, (double)ms};
    } else {
      it->second.first++;
      it->second.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto& n : order) {
    snprintf(line, sizeof(line), "%s %d %.6f\n", n.c_str(), agg[n].first, agg[n].second * 1000.0);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int)out.size();
}

int skgs_raster_layout_query(int32_t P, int32_t W, int32_t H, int64_t R_cap, skgs_raster_layout* out) {
  SKGS_CHECK_ARG(out != nullptr, "layout is NULL");
  SKGS_CHECK_ARG(P >= 0 && W > 0 && H > 0 && R_cap >= 0, "bad sizes P=%d W=%d H=%d R_cap=%lld", P, W, H,
                 (long long)R_cap);
  SKGS_CHECK_ARG(R_cap < (1ll << 31), "R_cap=%lld exceeds the 2^31 entries 32-bit list offsets can address",
                 (long long)R_cap);
  memset(out, 0, sizeof(*out));
  const size_t Pz = (size_t)P;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o = align_up(o + bytes, ARENA_ALIGN);
    return at;
  };
  // ---- geom (header + scan_state first: they are reset by one memset per forward)
  out->header = take(sizeof(skgs_raster_header));
  out->scan_state = take(((Pz + 255) / 256 + 1) * sizeof(uint64_t));
  out->means2D = take(Pz * 8);
  out->depths = take(Pz * 4);
  out->cov3D = take(Pz * 24);
  out->conic_opacity = take(Pz * 16);
  out->rgbd = take(Pz * 16);
  out->cull = take(Pz * 16);
  out->clamped = take(Pz);
  out->tiles_touched = take(Pz * 4);
  out->point_offsets = take(Pz * 4);
  out->geom_grads = take(Pz * 48);
  out->geom_bytes = o;
  // ---- binning: keys / values in emission order, the per-tile segments they are scattered into, and - after the
  //      per-tile sort - the sorted lists (in keys / vals again); the tile grid and the cursors are reset by one memset
  o = 0;
  const size_t Rz = (size_t)R_cap;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  out->keys = take(Rz * 8);
  out->vals = take(Rz * 4);
  out->tile_pairs = take(Rz * 8);
  out->tile_grid = take((size_t)(gx + 1) * (gy + 1) * sizeof(int32_t) * tile_cell_stride());
  out->tile_cursors = take((size_t)gx * gy * sizeof(uint32_t));
  out->binning_bytes = o;
  // ---- img
  o = 0;
  out->ranges = take((size_t)gx * gy * 8);
  out->n_contrib = take((size_t)W * H * 4);
  out->final_T = take((size_t)W * H * 4);
  out->tile_order = take((size_t)gx * gy * 16);
  out->work_counters = take(8 * sizeof(uint32_t));
  out->img_bytes = o;
  return SKGS_OK;
}

int skgs_raster_forward_geometry(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                 const float* shs, const float* colors_precomp, const float* opacities,
                                 const float* scales, const float* rotations, const float* cov3D_precomp, void* geom,
                                 int32_t* radii, void* binning, int64_t R_cap, void* img, uint32_t* num_rendered_host,
                                 void* stream) {
  RasterParams rp;
  int rc = fill_params(s, P, M, rp);
  if (rc) return rc;
  rc = check_inputs(s, P, M, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp);
  if (rc) return rc;
  SKGS_CHECK_ARG(geom != nullptr && (P == 0 || radii != nullptr), "geom arena / radii is NULL");
  SKGS_CHECK_ARG(binning == nullptr || (img != nullptr && R_cap > 0), "key emission needs the img arena and R_cap > 0");
  skgs_raster_layout lay;
  rc = skgs_raster_layout_query(P, rp.W, rp.H, binning ? R_cap : 0, &lay);
  if (rc) return rc;
  return launch_preprocess_scan(rp, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                (char*)geom, lay, radii, num_rendered_host, (char*)binning, (char*)img,
                                binning ? R_cap : 0, (cudaStream_t)stream);
}

int skgs_deform_forward_geometry(const skgs_skeleton* sk, const skgs_raster_settings* s, int32_t P, int32_t M_sh,
                                 const float* xyz, const float* scaling, const float* rotation,
                                 const float* opacity_logit, const float* shs, float* points, float* scales,
                                 float* rotations, float* opacities, float* d_rot, float* weights, int64_t* indices,
                                 float* sk_T, void* lbs_workspace, void* geom, int32_t* radii, void* binning,
                                 int64_t R_cap, void* img, uint32_t* num_rendered_host, void* stream) {
  RasterParams rp;
  int rc = fill_params(s, P, M_sh, rp);
  if (rc) return rc;
  SKGS_CHECK_ARG(sk != nullptr, "skeleton is NULL");
  SKGS_CHECK_ARG(P > 0, "the fused forward needs P > 0");
  SKGS_CHECK_ARG(xyz && scaling && rotation && opacity_logit && shs, "NULL canonical parameter");
  SKGS_CHECK_ARG(points && scales && rotations && opacities && d_rot && weights && indices, "NULL output");
  SKGS_CHECK_ARG(lbs_workspace && geom && radii && binning && img && R_cap > 0, "NULL workspace / arena");
  SKGS_CHECK_ARG(s->quat_wxyz == 0, "the fused forward assembles (x,y,z,w) rotations: quat_wxyz must be 0");
  SKGS_CHECK_ARG(M_sh >= (s->sh_degree + 1) * (s->sh_degree + 1) && M_sh <= 16,
                 "sh has %d coefficients, degree %d needs %d (max 16)", M_sh, s->sh_degree,
                 (s->sh_degree + 1) * (s->sh_degree + 1));
  skgs_raster_layout lay;
  rc = skgs_raster_layout_query(P, rp.W, rp.H, R_cap, &lay);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_fk_table(sk, sk_T, (float*)lbs_workspace, st);
  if (rc) return rc;
  rc = launch_deform_preprocess(rp, sk, (const float*)lbs_workspace, xyz, scaling, rotation, opacity_logit, shs, points,
                                scales, rotations, opacities, d_rot, weights, indices, (char*)geom, lay, radii,
                                (char*)binning, (char*)img, R_cap, st);
  if (rc) return rc;
  if (num_rendered_host) {
    auto* hdr = reinterpret_cast<skgs_raster_header*>((char*)geom + lay.header);
    SKGS_CUDA(cudaMemcpyAsync(num_rendered_host, hdr, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  }
  return SKGS_OK;
}

// binning (optionally with key emission from the stored geometry) + tile order + compositing
static int render_stage(const skgs_raster_settings* s, const RasterParams& rp, void* geom, void* binning, int64_t R_cap,
                        int64_t R_hint, void* img, const int32_t* radii, bool emit, float* out_color, float* out_depth,
                        float* out_alpha, uint32_t* num_rendered_host, cudaStream_t st) {
  SKGS_CHECK_ARG(geom && img && out_color && out_depth && out_alpha, "NULL arena / output");
  SKGS_CHECK_ARG(R_cap == 0 || binning != nullptr, "binning arena is NULL");
  skgs_raster_layout lay;
  int rc = skgs_raster_layout_query(rp.P, rp.W, rp.H, R_cap, &lay);
  if (rc) return rc;
  rc = launch_binning(rp, (char*)geom, (char*)binning, (char*)img, lay, radii, R_cap, R_hint, emit, num_rendered_host,
                      st);
  if (rc) return rc;

  if (s->debug & 2) return SKGS_OK;  // test hook: stop after binning
  return launch_composite_fwd(rp, (char*)geom, (char*)binning, (char*)img, lay, out_color, out_depth, out_alpha, st);
}

int skgs_raster_forward_render(const skgs_raster_settings* s, int32_t P, void* geom, void* binning, int64_t R_cap,
                               int64_t R_hint, void* img, const int32_t* radii, int32_t keys_emitted, float* out_color,
                               float* out_depth, float* out_alpha, uint32_t* num_rendered_host, void* stream) {
  RasterParams rp;
  int rc = fill_params(s, P, 0, rp);
  if (rc) return rc;
  return render_stage(s, rp, geom, binning, R_cap, R_hint, img, radii, keys_emitted == 0, out_color, out_depth,
                      out_alpha, num_rendered_host, (cudaStream_t)stream);
}

int skgs_raster_forward(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales,
                        const float* rotations, const float* cov3D_precomp, void* geom, void* binning, int64_t R_cap,
                        void* img, float* out_color, float* out_depth, float* out_alpha, int32_t* radii,
                        uint32_t* num_rendered_host, void* stream) {
  RasterParams rp;
  int rc = fill_params(s, P, M, rp);
  if (rc) return rc;
  rc = check_inputs(s, P, M, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp);
  if (rc) return rc;
  SKGS_CHECK_ARG(geom != nullptr && img != nullptr && (P == 0 || radii != nullptr), "geom / img arena / radii is NULL");
  SKGS_CHECK_ARG(R_cap == 0 || binning != nullptr, "binning arena is NULL");
  skgs_raster_layout lay;
  rc = skgs_raster_layout_query(P, rp.W, rp.H, R_cap, &lay);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // one kernel: preprocess + scan + key emission (the capacity is known up front)
  rc = launch_preprocess_scan(rp, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                              (char*)geom, lay, radii, nullptr, (char*)binning, (char*)img, R_cap, st);
  if (rc) return rc;
  return render_stage(s, rp, geom, binning, R_cap, 0, img, radii, /*emit=*/false, out_color, out_depth, out_alpha,
                      num_rendered_host, st);
}

static int raster_backward_impl(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                const float* shs, const float* colors_precomp, const float* scales,
                                const float* rotations, const float* cov3D_precomp, const int32_t* radii, void* geom,
                                const void* binning, int64_t R_cap, const void* img, const float* dL_dcolor,
                                const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D,
                                float* dL_dsh, float* dL_dcolors, float* dL_dopacity, float* dL_dscales,
                                float* dL_drotations, float* dL_dcov3D, const float* const* assemble_in,
                                float* const* assemble_out, void* stream) {
  RasterParams rp;
  int rc = fill_params(s, P, M, rp);
  if (rc) return rc;
  if (P == 0) return SKGS_OK;
  SKGS_CHECK_ARG(geom && img && dL_dcolor && radii && means3D, "NULL arena / dL_dcolor / radii / means3D");
  SKGS_CHECK_ARG(R_cap == 0 || binning != nullptr, "binning arena is NULL");
  SKGS_CHECK_ARG(dL_dmeans3D != nullptr, "dL_dmeans3D is required");
  SKGS_CHECK_ARG(shs == nullptr || dL_dsh != nullptr, "dL_dsh is required when shs is given");
  SKGS_CHECK_ARG(assemble_in != nullptr || scales == nullptr ||
                     (dL_dscales != nullptr && dL_drotations != nullptr && rotations != nullptr),
                 "dL_dscales / dL_drotations are required when scales/rotations are given");
  SKGS_CHECK_ARG(cov3D_precomp == nullptr || dL_dcov3D != nullptr, "dL_dcov3D is required when cov3D_precomp is given");
  skgs_raster_layout lay;
  rc = skgs_raster_layout_query(P, rp.W, rp.H, R_cap, &lay);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_composite_bwd(rp, (char*)geom, (const char*)binning, (const char*)img, lay, dL_dcolor, dL_ddepth,
                            dL_dalpha, (s->debug & 4) ? 1 : 0, st);
  if (rc) return rc;
  uint32_t* bwd_ticket = reinterpret_cast<uint32_t*>((char*)const_cast<void*>(img) + lay.work_counters) + 1;
  return launch_preprocess_bwd(rp, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, radii, (char*)geom,
                               lay, bwd_ticket, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacity, dL_dscales,
                               dL_drotations, dL_dcov3D, assemble_in, assemble_out, st);
}

int skgs_raster_backward(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D, const float* shs,
                         const float* colors_precomp, const float* scales, const float* rotations,
                         const float* cov3D_precomp, const int32_t* radii, void* geom, const void* binning,
                         int64_t R_cap, const void* img, const float* dL_dcolor, const float* dL_ddepth,
                         const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh,
                         float* dL_dcolors, float* dL_dopacity, float* dL_dscales, float* dL_drotations,
                         float* dL_dcov3D, void* stream) {
  return raster_backward_impl(s, P, M, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, radii, geom,
                              binning, R_cap, img, dL_dcolor, dL_ddepth, dL_dalpha, dL_dmeans3D, dL_dmeans2D, dL_dsh,
                              dL_dcolors, dL_dopacity, dL_dscales, dL_drotations, dL_dcov3D, nullptr, nullptr, stream);
}

int skgs_raster_assemble_backward(const skgs_raster_settings* s, int32_t P, int32_t M, const float* means3D,
                                  const float* shs, const float* scales, const float* rotations, const int32_t* radii,
                                  void* geom, const void* binning, int64_t R_cap, const void* img,
                                  const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                                  const float* scaling, const float* rotation, const float* opacity_logit,
                                  const float* d_rot, float* dL_dxyz, float* dL_dmeans2D, float* dL_dsh,
                                  float* dL_dscaling, float* dL_drotation, float* dL_dopacity, float* dL_dd_scale,
                                  float* dL_dd_xyz, float* dL_dd_rot, void* stream) {
  SKGS_CHECK_ARG(s != nullptr && s->quat_wxyz == 0, "the fused backward works on (x,y,z,w) rotations: quat_wxyz must be 0");
  SKGS_CHECK_ARG(shs && scales && rotations && scaling && rotation && opacity_logit, "NULL input");
  SKGS_CHECK_ARG(dL_dxyz && dL_dsh && dL_dscaling && dL_drotation && dL_dopacity && dL_dd_scale, "NULL output");
  const float* in[4] = {scaling, rotation, opacity_logit, d_rot};
  float* out[6] = {dL_dscaling, dL_drotation, dL_dopacity, dL_dd_scale, dL_dd_xyz, dL_dd_rot};
  return raster_backward_impl(s, P, M, means3D, shs, nullptr, scales, rotations, nullptr, radii, geom, binning, R_cap,
                              img, dL_dcolor, dL_ddepth, dL_dalpha, dL_dxyz, dL_dmeans2D, dL_dsh, nullptr, nullptr,
                              nullptr, nullptr, nullptr, in, out, stream);
}

}  // extern "C"
