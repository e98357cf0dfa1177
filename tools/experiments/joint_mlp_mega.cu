// joint_mlp_mega.cu - NOT part of the product build.  joint_mlp.cu with every stage of one direction recorded and run by
// ONE cooperative kernel (grid barriers between stages, L2-coherent operand loads) instead of 11 launches.  Correct
// (all 27 joint-network / train-loop tests pass) but SLOWER on B200: forward + backward at M = 32 167 us as one graph
// against 124 us for the 22 launches (full iteration 0.757 vs 0.704 ms).  Per stage the time is the product itself
// (~3 us: two to three dependent L2 round trips for 40 loads per thread, then a 256-long FMA chain) plus the barrier
// (~1.5 us with 32-296 CTAs polling one word), not the launch.  What would help is keeping the activations on chip
// between layers (thread-block cluster + distributed shared memory) and prefetching the next layer's weights behind the
// barrier - not a regrouping of the same global-memory products.
// joint_mlp.cu - the joint-rotation network of the `sk` stage, forward and backward: the step BEFORE
// forward kinematics.  joints [M,3], time t  ->  per-joint rotation quaternion sk_r [M,4], d_rot [M,4], d_scale [M,3].
//
// Reference: SimpleDeformationNetwork networks/sk_gs.py:134-164 (frequency encoders of position and time, concatenation,
// MLP), MLP_with_skips my_ext/blocks/mlp.py:44-85 (ReLU after every hidden layer; a skip layer concatenates the encoded
// input AFTER its ReLU; one Linear per head), frequency encoder my_ext/_C/src/nerf/freqencoder.cu:7-31 / :36-62, the
// head sk_r = normalize(out + (0,0,0,1)) networks/sk_gs.py:1075-1076; configuration exps/default.yaml:48-55.
// The reference runs this as ~30 torch/cuBLAS launches forward and ~60 backward on M <= 64 rows - pure launch latency.
//
// Round-1 form: M is tiny, so every matrix product is one launch of a small strided fp32 GEMM (`small_gemm_kernel`,
// exact FFMA chains, no tensor cores: results must match the reference's fp32 Linear layers) with bias, ReLU, ReLU-mask,
// skip-concatenation and bias-gradient fused in, and the independent products of one backward layer (dW, dZ_prev, dx0)
// share a launch: 11 launches forward, 11 backward, all capturable in the step's CUDA graph.
//
// Round-2 form (default): the same products, recorded instead of launched, run as the STAGES of one cooperative kernel
// per direction (`joint_mlp_mega_kernel`): a stage's CTA tiles are spread over the grid, a grid-wide barrier (one atomic
// + acquire spin per CTA) separates dependent stages, every inter-stage operand is read with L2-coherent loads.  One
// launch instead of eleven: the chain of dependent launches (each ~9 us inside the graph for ~3 us of work) was the
// cost.  SKGS_MLP_MEGA=0 selects the one-launch-per-product form (same arithmetic, bit-identical results).
#include <cstring>

#include "common.cuh"

namespace skgs {
namespace {

constexpr int GM_TI = 32;                    // rows of C per CTA: one per lane
#ifndef SKGS_GM_WARPS
#define SKGS_GM_WARPS 4
#endif
#ifndef SKGS_GM_JW
#define SKGS_GM_JW 2
#endif
constexpr int GM_WARPS = SKGS_GM_WARPS;
constexpr int GM_JW = SKGS_GM_JW;            // columns of C per warp
constexpr int GM_TJ = GM_WARPS * GM_JW;      // columns of C per CTA
constexpr int GM_RC = 128;                   // reduction chunk staged in shared memory
constexpr int GM_LD = GM_RC + 4;             // row stride: 16-byte aligned rows, conflict-free 128-bit loads
constexpr int GM_THREADS = GM_WARPS * 32;
constexpr int MAX_X0_USERS = 8;
static_assert((GM_TI * GM_RC) % GM_THREADS == 0 && (GM_TJ * GM_RC) % GM_THREADS == 0, "whole loads per thread");

// C(i,j) = epilogue( sum_r A(i,r) * B(j,r) ),  i < I, j < J, r < R; every operand is addressed through element strides.
struct GemmOp {
  int I, J, R;
  const float* A;  int a_si, a_sr;                       // (every operand here is far below 2^31 elements)
  const float* A2; int a2_si, a2_sr; int r_split;         // r >= r_split reads A2(i, r - r_split)  (skip concatenation)
  const float* B;  int b_sj, b_sr;
  const float* B2; int b2_sj, b2_sr; int j_split;         // j >= j_split reads B2(j - j_split, r)
  int ones_col;                                           // B(ones_col, r) = 1: column sums of A (bias gradient)
  float* C; int c_si, c_sj;
  float* C_ones;                                          // where column `ones_col` of C goes (contiguous in i)
  const float* bias;                                      // + bias[j]
  int relu;                                               // max(., 0)
  const float* mask; int m_si, m_sj;                      // * (mask(i,j) > 0): ReLU backward
};

// Operand access is split in two so that the loads of a chunk can all be in flight at once: `*_addr` yields an address
// that is always safe to read (out-of-range elements read the operand's first element), `*_value` applies the padding
// rule to the loaded value when it is written to shared memory.  (A load whose result is selected right away, or that
// sits behind a data-dependent branch, makes the in-order warp wait for it: 40 serialised L2 round trips per chunk.)
// All of this works on a register-resident copy of the descriptor: the launch batches several products, so the
// descriptor is indexed dynamically in parameter space and every field access would otherwise be an LDC round trip.
__device__ __forceinline__ const float* gemm_a_addr(const GemmOp& op, int i, int r) {
  const int off = r < op.r_split ? i * op.a_si + r * op.a_sr : i * op.a2_si + (r - op.r_split) * op.a2_sr;
  const float* base = r < op.r_split ? op.A : op.A2;
  return (i < op.I && r < op.R) ? base + off : op.A;
}
__device__ __forceinline__ float gemm_a_value(const GemmOp& op, int i, int r, float loaded) {
  return (i < op.I && r < op.R) ? loaded : 0.f;
}
__device__ __forceinline__ const float* gemm_b_addr(const GemmOp& op, int j, int r) {
  const int off = j < op.j_split ? j * op.b_sj + r * op.b_sr : (j - op.j_split) * op.b2_sj + r * op.b2_sr;
  const float* base = j < op.j_split ? op.B : op.B2;
  return (j < op.J && r < op.R && j != op.ones_col) ? base + off : op.B;
}
__device__ __forceinline__ float gemm_b_value(const GemmOp& op, int j, int r, float loaded) {
  return (j < op.J && r < op.R) ? (j == op.ones_col ? 1.f : loaded) : 0.f;
}

// Up to GM_BATCH independent products per launch (e.g. dW, dZ_prev and dx0 of one layer): dependent launches cost more
// than these kernels run, so everything that may run side by side shares a grid.
constexpr int GM_BATCH = 3;
struct GemmBatch {
  GemmOp op[GM_BATCH];
  int block_start[GM_BATCH + 1];
  int count;
};

// COHERENT: operands may have been written by other CTAs of the SAME launch (mega kernel): bypass L1 / the read-only path
template <bool COHERENT>
__device__ __forceinline__ float gm_load(const float* p) {
  return COHERENT ? __ldcg(p) : __ldg(p);
}

// one CTA tile (32 rows x GM_TJ columns) of one product of `batch`; `block` counts the tiles of the whole batch
template <bool COHERENT>
__device__ __forceinline__ void gemm_tile(const GemmBatch& batch, int block, float (*As)[GM_LD], float (*Bs)[GM_LD]) {
  constexpr int A_PER = GM_TI * GM_RC / GM_THREADS, B_PER = GM_TJ * GM_RC / GM_THREADS;
  int which = 0;
  while (which + 1 < batch.count && block >= batch.block_start[which + 1]) ++which;
  const GemmOp op = batch.op[which];  // one copy into registers (see above)
  const int local = block - batch.block_start[which];
  const int blocks_i = (op.I + GM_TI - 1) / GM_TI;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i0 = (local % blocks_i) * GM_TI, j0 = (local / blocks_i) * GM_TJ;
  const bool a_fast_r = op.a_sr == 1, b_fast_r = op.b_sr == 1;  // walk the contiguous direction with consecutive threads
  float acc[GM_JW];
#pragma unroll
  for (int jj = 0; jj < GM_JW; ++jj) acc[jj] = 0.f;

  // The operands are tiny and live in L2: the kernel is pure load latency, so every thread issues ALL its loads of a
  // chunk back to back into registers, and the next chunk's loads are in flight while the current one is multiplied.
  float ra[A_PER], rb[B_PER];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int u = 0; u < A_PER; ++u) {
      const int e = tid + u * GM_THREADS;
      const int ii = a_fast_r ? e / GM_RC : e % GM_TI, rr = a_fast_r ? e % GM_RC : e / GM_TI;
      ra[u] = gm_load<COHERENT>(gemm_a_addr(op, i0 + ii, r0 + rr));
    }
#pragma unroll
    for (int u = 0; u < B_PER; ++u) {
      const int e = tid + u * GM_THREADS;
      const int jj = b_fast_r ? e / GM_RC : e % GM_TJ, rr = b_fast_r ? e % GM_RC : e / GM_TJ;
      rb[u] = gm_load<COHERENT>(gemm_b_addr(op, j0 + jj, r0 + rr));
    }
  };
  fetch(0);
  for (int r0 = 0; r0 < op.R; r0 += GM_RC) {
#pragma unroll
    for (int u = 0; u < A_PER; ++u) {
      const int e = tid + u * GM_THREADS;
      const int ii = a_fast_r ? e / GM_RC : e % GM_TI, rr = a_fast_r ? e % GM_RC : e / GM_TI;
      As[ii][rr] = gemm_a_value(op, i0 + ii, r0 + rr, ra[u]);
    }
#pragma unroll
    for (int u = 0; u < B_PER; ++u) {
      const int e = tid + u * GM_THREADS;
      const int jj = b_fast_r ? e / GM_RC : e % GM_TJ, rr = b_fast_r ? e % GM_RC : e / GM_TJ;
      Bs[jj][rr] = gemm_b_value(op, j0 + jj, r0 + rr, rb[u]);
    }
    __syncthreads();
    if (r0 + GM_RC < op.R) fetch(r0 + GM_RC);
#pragma unroll 4
    for (int rr = 0; rr < GM_RC; rr += 4) {
      const float4 a = *reinterpret_cast<const float4*>(&As[lane][rr]);
#pragma unroll
      for (int jj = 0; jj < GM_JW; ++jj) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[warp * GM_JW + jj][rr]);
        acc[jj] = fmaf(a.x, b.x, acc[jj]);
        acc[jj] = fmaf(a.y, b.y, acc[jj]);
        acc[jj] = fmaf(a.z, b.z, acc[jj]);
        acc[jj] = fmaf(a.w, b.w, acc[jj]);
      }
    }
    __syncthreads();
  }

  const int i = i0 + lane;
  if (i >= op.I) return;
#pragma unroll
  for (int jj = 0; jj < GM_JW; ++jj) {
    const int j = j0 + warp * GM_JW + jj;
    if (j >= op.J) continue;
    float v = acc[jj];
    if (j == op.ones_col) {
      op.C_ones[i] = v;
      continue;
    }
    if (op.bias) v += gm_load<COHERENT>(op.bias + j);
    if (op.relu) v = fmaxf(v, 0.f);
    if (op.mask) v = gm_load<COHERENT>(op.mask + i * op.m_si + j * op.m_sj) > 0.f ? v : 0.f;
    op.C[i * op.c_si + j * op.c_sj] = v;
  }
}

__global__ void __launch_bounds__(GM_THREADS) small_gemm_kernel(const __grid_constant__ GemmBatch batch) {
  __shared__ __align__(16) float As[GM_TI][GM_LD];
  __shared__ __align__(16) float Bs[GM_TJ][GM_LD];
  gemm_tile<false>(batch, (int)blockIdx.x, As, Bs);
}

GemmOp gemm_op(int I, int J, int R) {
  GemmOp op;
  memset(&op, 0, sizeof(op));
  op.I = I;
  op.J = J;
  op.R = R;
  op.r_split = R;
  op.j_split = J;
  op.ones_col = -1;
  return op;
}

// ---------------------------------------------------------------------------------------------------------------
// stages: what one direction of the network consists of.  Either every stage is a launch of its own, or - default - the
// stages are recorded and run by ONE cooperative kernel with grid barriers in between (joint_mlp_mega_kernel below).
// ---------------------------------------------------------------------------------------------------------------
enum StageKind { ST_GEMM = 0, ST_ENCODE = 1, ST_HEAD = 2, ST_HEAD_BWD = 3, ST_ENCODE_BWD = 4 };
struct EwArgs {            // arguments of the element-wise stages (meaning per kind: see ew_stage)
  int M, i0, i1, i2;
  const float *p0, *p1, *p2, *p3;
  float *q0, *q1, *q2;
  long long stride;
};
struct Stage {
  int kind, blocks;
  union {
    GemmBatch gemm;
    EwArgs ew;
  };
};
constexpr int MEGA_MAX_STAGES = 36;  // depth <= 32 hidden layers + heads + encoder + head epilogue
struct MegaArgs {
  Stage stage[MEGA_MAX_STAGES];
  int count;
  unsigned int* bar;  // grid-barrier counter, zero at launch
};
static_assert(sizeof(MegaArgs) <= 32 * 1024 - 64, "kernel parameters are limited to 32 KB");

struct Recorder {
  MegaArgs a;
  int max_blocks;
};
thread_local Recorder* g_rec = nullptr;

bool mega_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SKGS_MLP_MEGA");
    on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return on == 1;
}

int record_stage(const Stage& st) {
  SKGS_CHECK_ARG(g_rec->a.count < MEGA_MAX_STAGES, "joint_mlp: more than %d stages", MEGA_MAX_STAGES);
  g_rec->a.stage[g_rec->a.count++] = st;
  if (st.blocks > g_rec->max_blocks) g_rec->max_blocks = st.blocks;
  return SKGS_OK;
}

int launch_gemms(const GemmOp* ops, int n, const char* name, cudaStream_t st) {
  GemmBatch b;
  memset(&b, 0, sizeof(b));
  for (int k = 0; k < n; ++k) {
    if (ops[k].I <= 0 || ops[k].J <= 0) continue;
    b.op[b.count] = ops[k];
    b.block_start[b.count + 1] =
        b.block_start[b.count] + ((ops[k].I + GM_TI - 1) / GM_TI) * ((ops[k].J + GM_TJ - 1) / GM_TJ);
    ++b.count;
  }
  if (b.count == 0) return SKGS_OK;
  if (g_rec != nullptr) {
    Stage s;
    memset(&s, 0, sizeof(s));
    s.kind = ST_GEMM;
    s.blocks = b.block_start[b.count];
    s.gemm = b;
    return record_stage(s);
  }
  {
    ProfScope prof_(name, st);
    small_gemm_kernel<<<b.block_start[b.count], GM_THREADS, 0, st>>>(b);
  }
  SKGS_CHECK_LAUNCH(name);
  return SKGS_OK;
}

int launch_gemm(const GemmOp& op, const char* name, cudaStream_t st) { return launch_gemms(&op, 1, name, st); }

// ---------------------------------------------------------------------------------------------------------------
// frequency encoders (freqencoder.cu:7-31): x0[m] = [enc_p(joints[m]) | enc_t(t)],
// enc(x) = (x, sin(2^0 x), sin(2^0 x + pi/2), sin(2^1 x), ...) in blocks of D.  The argument arithmetic is the
// reference's (scalbnf, one fp32 addition of fl32(pi/2)); the sine itself is the accurate sinf, not __sinf.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float freq_value(const float* x, int D, int c) {
  if (c < D) return x[c];
  const int col = c / D - 1, d = c % D, freq = col >> 1;
  const float phase = (col & 1) ? 1.5707963267948966f : 0.f;
  return sinf(__fadd_rn(scalbnf(x[d], freq), phase));
}

// The four element-wise stages, one CTA of GM_THREADS threads = `block`-th slice of the stage's index space.
//   ST_ENCODE     M, i0 = deg_p, i1 = deg_t; p0 = joints, p1 = t; q0 = x0
//   ST_HEAD       M, i0 = n_out, i1 = rotation_head; p0 = out; q0 = sk_r, q1 = d_rot, q2 = d_scale
//                 heads: out[m] = (q[4], d_rot[4], d_scale[3]); sk_r = normalize(q + (0,0,0,1)), F.normalize's eps 1e-12
//   ST_HEAD_BWD   M, i0 = n_out, i1 = rotation_head; p0 = out, p1 = g_sk_r, p2 = g_d_rot, p3 = g_d_scale; q0 = d_out
//   ST_ENCODE_BWD M, i0 = deg_p, i1 = enc, i2 = n_users; p0 = x0, p1 = dx0; stride = user stride; q0 = dL_djoints
//                 freqencoder.cu:36-62: dx[d] = g[d] + sum_f 2^f (g_sin * out_cos - g_cos * out_sin), g = sum of the
//                 gradients that reached the encoded input (layer 0 and every skip layer); one warp per (joint, coord)
template <bool COHERENT>
__device__ __forceinline__ void ew_stage(int kind, const EwArgs& a, int block) {
  const int gt = block * GM_THREADS + (int)threadIdx.x;
  if (kind == ST_ENCODE) {
    const int Cp = 3 * (1 + 2 * a.i0), Ct = 1 + 2 * a.i1, enc = Cp + Ct;
    if (gt >= a.M * enc) return;
    const int m = gt / enc, c = gt - m * enc;
    a.q0[gt] = c < Cp ? freq_value(a.p0 + 3 * m, 3, c) : freq_value(a.p1, 1, c - Cp);
  } else if (kind == ST_HEAD) {
    const int m = gt, n_out = a.i0;
    if (m >= a.M) return;
    const float* o = a.p0 + (size_t)m * n_out;
    float q[4] = {gm_load<COHERENT>(o), gm_load<COHERENT>(o + 1), gm_load<COHERENT>(o + 2), gm_load<COHERENT>(o + 3)};
    if (a.i1) {
      q[3] += 1.0f;
      const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
      for (int k = 0; k < 4; ++k) q[k] /= n;
    }
    for (int k = 0; k < 4; ++k) a.q0[4 * m + k] = q[k];
    for (int k = 0; k < 4; ++k) a.q1[4 * m + k] = gm_load<COHERENT>(o + 4 + k);
    for (int k = 0; k < 3; ++k) a.q2[3 * m + k] = gm_load<COHERENT>(o + 8 + k);
  } else if (kind == ST_HEAD_BWD) {
    const int m = gt, n_out = a.i0;
    if (m >= a.M) return;
    const float* o = a.p0 + (size_t)m * n_out;
    float* d = a.q0 + (size_t)m * n_out;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.p1)
      for (int k = 0; k < 4; ++k) g[k] = a.p1[4 * m + k];
    if (a.i1) {
      const float q[4] = {o[0], o[1], o[2], o[3] + 1.0f};
      const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      if (nrm > 1e-12f) {  // d normalize: (g - q^ (q^ . g)) / |q|
        const float inv = 1.0f / nrm;
        float dot = 0.f;
        for (int k = 0; k < 4; ++k) dot += q[k] * inv * g[k];
        for (int k = 0; k < 4; ++k) g[k] = (g[k] - q[k] * inv * dot) * inv;
      } else {
        for (int k = 0; k < 4; ++k) g[k] *= 1e12f;
      }
    }
    for (int k = 0; k < 4; ++k) d[k] = g[k];
    for (int k = 0; k < 4; ++k) d[4 + k] = a.p2 ? a.p2[4 * m + k] : 0.f;
    for (int k = 0; k < 3; ++k) d[8 + k] = a.p3 ? a.p3[3 * m + k] : 0.f;
  } else if (kind == ST_ENCODE_BWD) {
    // one warp per (joint, coordinate): lane 0 takes the pass-through term, lane f + 1 frequency f (deg_p <= 16)
    const int w = gt >> 5, lane = threadIdx.x & 31;
    if (w >= a.M * 3) return;
    const int deg_p = a.i0, enc = a.i1, n_users = a.i2;
    const int m = w / 3, dd = w - m * 3;
    const float* o = a.p0 + (size_t)m * enc;
    const float* g = a.p1 + (size_t)m * enc;
    float r = 0.f;
    if (lane == 0) {
      for (int u = 0; u < n_users; ++u) r += gm_load<COHERENT>(g + u * a.stride + dd);
    } else if (lane <= deg_p) {
      const int f = lane - 1, cs = 3 + 6 * f + dd, cc = cs + 3;
      float gs = 0.f, gc = 0.f;
      for (int u = 0; u < n_users; ++u) {
        gs += gm_load<COHERENT>(g + u * a.stride + cs);
        gc += gm_load<COHERENT>(g + u * a.stride + cc);
      }
      r = scalbnf(1.0f, f) * (gs * o[cc] - gc * o[cs]);
    }
    r = warp_sum(r);
    if (lane == 0) a.q0[w] = r;
  }
}

__global__ void __launch_bounds__(GM_THREADS) joint_ew_kernel(int kind, EwArgs a) { ew_stage<false>(kind, a, blockIdx.x); }

int launch_ew(int kind, const EwArgs& a, int blocks, const char* name, cudaStream_t st) {
  if (blocks <= 0) return SKGS_OK;
  if (g_rec != nullptr) {
    Stage s;
    memset(&s, 0, sizeof(s));
    s.kind = kind;
    s.blocks = blocks;
    s.ew = a;
    return record_stage(s);
  }
  {
    ProfScope prof_(name, st);
    joint_ew_kernel<<<blocks, GM_THREADS, 0, st>>>(kind, a);
  }
  SKGS_CHECK_LAUNCH(name);
  return SKGS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// all stages of one direction in ONE cooperative launch.  A stage's tiles are spread over the grid; stages are
// separated by a grid barrier: every CTA publishes its writes (fence), arrives with one atomic and thread 0 spins with
// acquire loads until the whole grid has arrived.  Everything a later stage reads is loaded through L2 (gm_load<true>).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(GM_THREADS) joint_mlp_mega_kernel(const __grid_constant__ MegaArgs a) {
  __shared__ __align__(16) float As[GM_TI][GM_LD];
  __shared__ __align__(16) float Bs[GM_TJ][GM_LD];
  unsigned int target = 0;
  for (int s = 0; s < a.count; ++s) {
    const int kind = a.stage[s].kind, blocks = a.stage[s].blocks;
    for (int b = blockIdx.x; b < blocks; b += gridDim.x) {
      if (kind == ST_GEMM)
        gemm_tile<true>(a.stage[s].gemm, b, As, Bs);
      else
        ew_stage<true>(kind, a.stage[s].ew, b);
      __syncthreads();  // the operand tiles in shared memory are reused by the next tile
    }
    if (s + 1 < a.count) {
      target += gridDim.x;
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.bar, 1u);
        while (ld_acquire_u32(a.bar) < target) {
        }
      }
      __syncthreads();
    }
  }
}

// run the recorded stages: memset of the barrier word + one cooperative launch (all CTAs must be co-resident)
int launch_mega(Recorder& rec, unsigned int* bar, const char* name, cudaStream_t st) {
  if (rec.a.count == 0) return SKGS_OK;
  static int resident = 0;
  if (resident == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    SKGS_CUDA(cudaGetDevice(&dev));
    SKGS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SKGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, joint_mlp_mega_kernel, GM_THREADS, 0));
    resident = sms * (per_sm < 1 ? 1 : per_sm);
  }
  rec.a.bar = bar;
  int grid = rec.max_blocks < resident ? rec.max_blocks : resident;
  if (grid < 1) grid = 1;
  SKGS_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned int), st));
  void* args[1] = {(void*)&rec.a};
  {
    ProfScope prof_(name, st);
    SKGS_CUDA(cudaLaunchCooperativeKernel((const void*)joint_mlp_mega_kernel, dim3(grid), dim3(GM_THREADS), args, 0, st));
  }
  SKGS_CHECK_LAUNCH(name);
  return SKGS_OK;
}

// arms the recorder for the lifetime of one API call (unless SKGS_MLP_MEGA=0), disarms it on every exit path
struct RecordScope {
  Recorder& rec;
  bool armed;
  explicit RecordScope(Recorder& r) : rec(r), armed(mega_enabled()) {
    if (armed) {
      memset(&rec, 0, sizeof(rec));
      g_rec = &rec;
    }
  }
  ~RecordScope() { g_rec = nullptr; }
  int finish(unsigned int* bar, const char* name, cudaStream_t st) {
    if (!armed) return SKGS_OK;
    g_rec = nullptr;
    return launch_mega(rec, bar, name, st);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// host-side description of the network
// ---------------------------------------------------------------------------------------------------------------
struct Net {
  int M, enc, width, depth, n_out;
  int in_dim[33];          // input width of hidden layer i (i < depth) and of the heads (i == depth)
  bool skip_in[33];        // that input is [previous activation | x0]
  long long w_off[33], b_off[33], total;
  int x0_users, user_of_layer[33];  // which dx0 buffer receives the gradient of layer i's x0 part
  // workspace offsets (floats)
  long long o_x0, o_act, o_out, o_dza, o_dzb, o_dzh, o_dx0, ws_floats;
};

int describe(const skgs_joint_mlp* n, Net& N) {
  SKGS_CHECK_ARG(n != nullptr, "joint_mlp: null descriptor");
  SKGS_CHECK_ARG(n->M >= 0 && n->degree_p >= 0 && n->degree_t >= 0 && n->degree_p <= 16 && n->degree_t <= 16,
                 "joint_mlp: invalid M %d / encoder degrees %d, %d", n->M, n->degree_p, n->degree_t);
  SKGS_CHECK_ARG(n->width > 0 && n->depth >= 1 && n->depth <= 32, "joint_mlp: invalid width %d / depth %d", n->width,
                 n->depth);
  SKGS_CHECK_ARG(n->n_out == 11, "joint_mlp: the heads are (4, 4, 3) = 11 outputs (sk_gs.py:519), got %d", n->n_out);
  N.M = n->M;
  N.enc = 3 * (1 + 2 * n->degree_p) + (1 + 2 * n->degree_t);
  N.width = n->width;
  N.depth = n->depth;
  N.n_out = n->n_out;
  long long off = 0;
  N.x0_users = 0;
  for (int i = 0; i <= N.depth; ++i) {
    N.skip_in[i] = i > 0 && ((n->skip_mask >> (i - 1)) & 1);
    N.in_dim[i] = i == 0 ? N.enc : N.width + (N.skip_in[i] ? N.enc : 0);
    N.user_of_layer[i] = (i == 0 || N.skip_in[i]) ? N.x0_users++ : -1;
    const int out = i < N.depth ? N.width : N.n_out;
    N.w_off[i] = off;
    off += (long long)out * N.in_dim[i];
    N.b_off[i] = off;
    off += out;
  }
  SKGS_CHECK_ARG(N.x0_users <= MAX_X0_USERS, "joint_mlp: at most %d skip connections", MAX_X0_USERS - 1);
  N.total = off;
  long long w = 0;
  auto take = [&](long long count) {
    const long long at = w;
    w += (count + 63) / 64 * 64;
    return at;
  };
  N.o_x0 = take((long long)N.M * N.enc);
  N.o_act = take((long long)N.depth * N.M * N.width);
  N.o_out = take((long long)N.M * N.n_out);
  N.o_dza = take((long long)N.M * N.width);
  N.o_dzb = take((long long)N.M * N.width);
  N.o_dzh = take((long long)N.M * N.n_out);
  N.o_dx0 = take((long long)N.x0_users * N.M * N.enc);
  N.ws_floats = w;
  return SKGS_OK;
}

}  // namespace
}  // namespace skgs

using namespace skgs;

extern "C" {

int skgs_joint_mlp_layout(const skgs_joint_mlp* net, int64_t* weight_offsets, int64_t* bias_offsets,
                          int32_t* in_dims, int64_t* total) {
  Net N;
  if (int rc = describe(net, N)) return rc;
  for (int i = 0; i <= N.depth; ++i) {
    if (weight_offsets) weight_offsets[i] = N.w_off[i];
    if (bias_offsets) bias_offsets[i] = N.b_off[i];
    if (in_dims) in_dims[i] = N.in_dim[i];
  }
  if (total) *total = N.total;
  return SKGS_OK;
}

size_t skgs_joint_mlp_workspace_bytes(const skgs_joint_mlp* net) {
  Net N;
  if (describe(net, N)) return 0;
  return (size_t)N.ws_floats * sizeof(float) + 256;
}

int skgs_joint_mlp_forward(const skgs_joint_mlp* net, const float* joints, const float* t, float* sk_r, float* d_rot,
                           float* d_scale, void* workspace, void* stream) {
  Net N;
  if (int rc = describe(net, N)) return rc;
  if (N.M == 0) return SKGS_OK;
  SKGS_CHECK_ARG(net->theta && joints && t && sk_r && d_rot && d_scale && workspace, "joint_mlp_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  float* x0 = ws + N.o_x0;
  const float* theta = net->theta;
  Recorder rec;
  RecordScope scope(rec);  // default: the stages below are recorded and run by one cooperative launch at the end
  {
    EwArgs e = {};
    e.M = N.M; e.i0 = net->degree_p; e.i1 = net->degree_t;
    e.p0 = joints; e.p1 = t; e.q0 = x0;
    if (int rc = launch_ew(ST_ENCODE, e, (N.M * N.enc + GM_THREADS - 1) / GM_THREADS, "joint_encode_kernel", st)) return rc;
  }
  for (int i = 0; i <= N.depth; ++i) {
    const bool head = i == N.depth;
    const int out = head ? N.n_out : N.width, in = N.in_dim[i];
    GemmOp op = gemm_op(N.M, out, in);
    if (i == 0) {
      op.A = x0;
      op.a_si = N.enc;
      op.a_sr = 1;
    } else {
      op.A = ws + N.o_act + (long long)(i - 1) * N.M * N.width;
      op.a_si = N.width;
      op.a_sr = 1;
      if (N.skip_in[i]) {
        op.r_split = N.width;
        op.A2 = x0;
        op.a2_si = N.enc;
        op.a2_sr = 1;
      }
    }
    op.B = theta + N.w_off[i];
    op.b_sj = in;
    op.b_sr = 1;
    op.bias = theta + N.b_off[i];
    op.relu = head ? 0 : 1;
    op.C = head ? ws + N.o_out : ws + N.o_act + (long long)i * N.M * N.width;
    op.c_si = out;
    op.c_sj = 1;
    if (int rc = launch_gemm(op, head ? "joint_mlp_head_gemm" : "joint_mlp_layer_gemm", st)) return rc;
  }
  {
    EwArgs e = {};
    e.M = N.M; e.i0 = N.n_out; e.i1 = net->rotation_head;
    e.p0 = ws + N.o_out; e.q0 = sk_r; e.q1 = d_rot; e.q2 = d_scale;
    if (int rc = launch_ew(ST_HEAD, e, (N.M + GM_THREADS - 1) / GM_THREADS, "joint_head_kernel", st)) return rc;
  }
  return scope.finish(reinterpret_cast<unsigned int*>(ws + N.ws_floats), "joint_mlp_fwd_kernel", st);
}

int skgs_joint_mlp_backward(const skgs_joint_mlp* net, const float* dL_dsk_r, const float* dL_dd_rot,
                            const float* dL_dd_scale, float* dL_dtheta, float* dL_djoints, void* workspace,
                            void* stream) {
  Net N;
  if (int rc = describe(net, N)) return rc;
  if (N.M == 0) return SKGS_OK;
  SKGS_CHECK_ARG(net->theta && dL_dtheta && workspace, "joint_mlp_backward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const float* x0 = ws + N.o_x0;
  const float* theta = net->theta;
  float* dzh = ws + N.o_dzh;
  Recorder rec;
  RecordScope scope(rec);
  {
    EwArgs e = {};
    e.M = N.M; e.i0 = N.n_out; e.i1 = net->rotation_head;
    e.p0 = ws + N.o_out; e.p1 = dL_dsk_r; e.p2 = dL_dd_rot; e.p3 = dL_dd_scale; e.q0 = dzh;
    if (int rc = launch_ew(ST_HEAD_BWD, e, (N.M + GM_THREADS - 1) / GM_THREADS, "joint_head_bwd_kernel", st)) return rc;
  }
  // dZ of layer i (pre-activation gradient, [M, out_i]); the heads first, then hidden layers depth-1 .. 0
  const float* dz = dzh;
  for (int i = N.depth; i >= 0; --i) {
    const bool head = i == N.depth;
    const int out = head ? N.n_out : N.width, in = N.in_dim[i];
    const float* prev = i == 0 ? x0 : ws + N.o_act + (long long)(i - 1) * N.M * N.width;
    const int prev_w = i == 0 ? N.enc : N.width;
    GemmOp ops[GM_BATCH];
    int n_ops = 0;
    // dW[n][k] = sum_m dZ[m][n] * input[m][k], plus the bias gradient as the "ones" column k == in
    {
      GemmOp op = gemm_op(out, in + 1, N.M);
      op.A = dz;
      op.a_si = 1;
      op.a_sr = out;
      op.B = prev;
      op.b_sj = 1;
      op.b_sr = prev_w;
      if (i > 0 && N.skip_in[i]) {
        op.j_split = N.width;
        op.B2 = x0;
        op.b2_sj = 1;
        op.b2_sr = N.enc;
      }
      op.ones_col = in;
      op.C = dL_dtheta + N.w_off[i];
      op.c_si = in;
      op.c_sj = 1;
      op.C_ones = dL_dtheta + N.b_off[i];
      ops[n_ops++] = op;
    }
    // gradient w.r.t. the encoded input, where this layer reads it
    if (N.user_of_layer[i] >= 0 && dL_djoints) {
      GemmOp op = gemm_op(N.M, N.enc, out);
      op.A = dz;
      op.a_si = out;
      op.a_sr = 1;
      op.B = theta + N.w_off[i] + (i == 0 ? 0 : N.width);
      op.b_sj = 1;
      op.b_sr = in;
      op.C = ws + N.o_dx0 + (long long)N.user_of_layer[i] * N.M * N.enc;
      op.c_si = N.enc;
      op.c_sj = 1;
      ops[n_ops++] = op;
    }
    // dZ of the previous hidden layer: (dZ W)[:, :width] masked by that layer's ReLU
    const float* dz_next = dz;
    if (i > 0) {
      float* dz_prev = ws + (((N.depth - i) & 1) ? N.o_dzb : N.o_dza);
      GemmOp op = gemm_op(N.M, N.width, out);
      op.A = dz;
      op.a_si = out;
      op.a_sr = 1;
      op.B = theta + N.w_off[i];
      op.b_sj = 1;
      op.b_sr = in;
      op.mask = prev;
      op.m_si = N.width;
      op.m_sj = 1;
      op.C = dz_prev;
      op.c_si = N.width;
      op.c_sj = 1;
      ops[n_ops++] = op;
      dz_next = dz_prev;
    }
    // the three products only read dZ of this layer: one launch
    if (int rc = launch_gemms(ops, n_ops, "joint_mlp_bwd_gemms", st)) return rc;
    dz = dz_next;
  }
  if (dL_djoints) {
    EwArgs e = {};
    e.M = N.M; e.i0 = net->degree_p; e.i1 = N.enc; e.i2 = N.x0_users;
    e.p0 = x0; e.p1 = ws + N.o_dx0; e.stride = (long long)N.M * N.enc; e.q0 = dL_djoints;
    if (int rc = launch_ew(ST_ENCODE_BWD, e, (N.M * 3 * 32 + GM_THREADS - 1) / GM_THREADS, "joint_encode_bwd_kernel", st))
      return rc;
  }
  return scope.finish(reinterpret_cast<unsigned int*>(ws + N.ws_floats), "joint_mlp_bwd_kernel", st);
}

}  // extern "C"
