// kernels.cuh — sm_100a device code of the STOMP rollout engine.
//
// Data layout in HBM (all fp64 unless noted; the LAST index is time and is contiguous):
//   theta[B][D][N]   pad_start[B][D]   pad_goal[B][D]
//   noise / params / control_costs / cumulative [B][R][D][N]      state_costs[B][R][N]
//   voxels[nx][ny][nz] (u8 / u16 squared cell distance, or f32 metres)
//
// Kernel <-> reference map (paths relative to stomp_motion_planner/ in the reference):
//   k_select_reuse   PolicyImprovement::generateRollouts, reuse part   src/policy_improvement.cpp:178-225
//   k_generate       generateRollouts (new part) + computeProjectedNoise + computeControlCosts
//                    src/policy_improvement.cpp:228-236,473-489; src/covariant_trajectory_policy.cpp:228-255;
//                    include/stomp_motion_planner/multivariate_gaussian.h:88-94
//   k_cost           StompOptimizer::execute / handleJointLimits / performForwardKinematics
//                    src/stomp_optimizer.cpp:562-709,1063-1165; stomp_collision_space.h:193-228
//   k_cumulative     computeRolloutCumulativeCosts + Rollout::getCost   src/policy_improvement.cpp:149-156,301-320
//   k_update         computeRolloutProbabilities + computeParameterUpdates + updateParameters
//                    src/policy_improvement.cpp:322-383; src/covariant_trajectory_policy.cpp:306-323
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/stomp_b200.h"

// build-time tuning knobs (A/B tested on B200, see profiles/README.md)
#ifndef STOMP_COST_PARAMS_CG
#define STOMP_COST_PARAMS_CG 1     // k_cost: stage rollouts with 16-byte cp.async.cg (L1 bypass) instead of 8-byte .ca (A/B: 0.336 -> 0.332 ms)
#endif
#ifndef STOMP_COST_VOX_EVICT_LAST
#define STOMP_COST_VOX_EVICT_LAST 0   // k_cost: L1 evict_last hint on the voxel gathers (A/B: no effect)
#endif
#ifndef STOMP_GEN_MIN_BLOCKS
#define STOMP_GEN_MIN_BLOCKS 4     // k_generate: resident CTAs per SM targeted by the register allocation
#endif

namespace stomp_dev {

constexpr int kPad = STOMP_DIFF_RULE_LENGTH - 1;  // 6 fixed points each side
constexpr int kMaxHb = 6;                         // half bandwidth of R (jerk stencil: 5, +1 spare)
constexpr int kMaxSlots = 4;
constexpr int kTileSteps = 29;                    // productive timesteps per warp (lanes 1..29; halo -1, +1, +2)

// ---------------------------------------------------------------------------------------------
// device tables
// ---------------------------------------------------------------------------------------------
template <typename Real>
struct alignas(16) DevNode {
  Real A0[9], A1[9], A2[9], p[3], ax[3];
  int parent, type, q_index, save_slot, load_slot, sphere_begin, sphere_end, pad_;
  int cluster_begin, cluster_end, pad2_[2];
};

template <typename Real>
struct alignas(16) DevSphere {
  Real pos[3], radius, clearance, inv_clearance, weight;
  int original_index;
};

// Broad phase of the collision cost (exact, see k_cost): a cluster is a run of consecutive spheres of one node with a centre
// c (node frame) and a threshold thr.  If the coarse lower bound of the distance field at the cluster centre is >= thr and the
// centre is `margin` fine cells inside the grid, every sphere of the cluster provably has zero potential and is not in
// collision, and a warp whose lanes all see that skips the cluster's spheres.  thr = +inf disables the test for a cluster.
template <typename Real>
struct alignas(16) DevCluster {
  Real c[3], thr, margin, pad_;
  int begin, end;       // sphere range
};
struct CullField {      // conservative coarse distance field: g[coarse cell] <= distance (m) of every fine cell of the 4^3 block
  const float* g;
  int nx, ny, nz;       // coarse dims
  int enabled;
};

// OrientationConstraintEvaluator (src/constraint_evaluator.cpp:50-114) folded for the device: the constrained segment's
// rotation is F(node).R * rel, the nominal orientation is pre-inverted, weights are 0 when the tolerance is >= pi.
constexpr int kMaxConstraints = 8;
template <typename Real>
struct alignas(16) DevConstraint {
  Real rel[9], nominal_inv[9], tol[3], w[3], weight;
  int node, body_fixed;
};

struct Band {           // banded Cholesky factor of R = C C^T and the projection scaling
  const double* fw;         // [N][8]  {1/C(i,i), C(i,i-1)/C(i,i), ..., C(i,i-6)/C(i,i), 0}
  const double* bw;         // [N][8]  {1/C(i,i), C(i+1,i)/C(i,i), ..., C(i+6,i)/C(i,i), 0}
  const double* proj_scale; // [N]     s_p = 1/(N * max column p of R^-1)
};

struct Stencil {        // control-cost stencils of CovariantTrajectoryPolicy
  double coef[3][7];    // rule_k[j] / dt^(k+1)
  double weight[3];     // derivative_costs
};

struct Sdf {
  const void* vox;       // [nx][ny][nz], z fastest: the layout of the C ABI (set_sdf / get_sdf), of the rebuild and of the trilinear mode
  int nx, ny, nz, dtype;
  double origin[3], res, inv_res;
  // the same cells in bricks of 4 x 4 x 2 (one 32-byte sector of u8 cells per brick): what k_cost's nearest-cell lookup gathers
  // from.  A warp's lanes are consecutive timesteps of one sphere, i.e. positions a cell or two apart in an arbitrary
  // direction; in the z-fastest layout a sector is a 32-cell column, so motion across x or y costs one sector per lane.
  const void* brick;     // nullptr: gather from `vox`
  int nby, nbz;          // bricks along y and z
};
#ifndef STOMP_SDF_BRICKS
#define STOMP_SDF_BRICKS 1
#endif
__host__ __device__ __forceinline__ int brick_index(int cx, int cy, int cz, int nby, int nbz) {
  return ((((cx >> 2) * nby + (cy >> 2)) * nbz + (cz >> 1)) << 5) | ((cx & 3) << 3) | ((cy & 3) << 1) | (cz & 1);
}
template <typename V>
__global__ void k_brick_relayout(int nx, int ny, int nz, int nby, int nbz, const V* __restrict__ src, V* __restrict__ dst) {
  const size_t cells = size_t(nx) * ny * nz;
  for (size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < cells; v += size_t(gridDim.x) * blockDim.x) {
    const int z = int(v % nz), y = int((v / nz) % ny), x = int(v / (size_t(nz) * ny));
    dst[brick_index(x, y, z, nby, nbz)] = src[v];
  }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: one stream per (problem, global rollout, dimension), counter = sample pair
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t stream, uint32_t iteration, uint32_t pair,
                                            double& z0, double& z1) {
  uint32_t c[4] = {pair, iteration, uint32_t(stream), uint32_t(stream >> 32)};
  philox4x32_10(c, uint32_t(seed), uint32_t(seed >> 32));
  // Box-Muller in fp32 (the exploration noise needs the distribution, not 53-bit normals): u1 in (0,1] from 32 bits
  // (|z| <= 6.66), the angle from 32 bits; evaluating log / sincospi in fp64 was ~20 % of k_generate's instructions
  const float u1 = (float(c[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float u2 = float(c[2] >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, co;
  sincospif(2.0f * u2, &sn, &co);
  z0 = double(r * co);
  z1 = double(r * sn);
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier completion -------------------------------------------------
// One elected thread arms the CTA's mbarrier with the byte count and issues one bulk copy per contiguous block (robot
// tables, rollout rows); the copy engine moves the bytes into shared memory while the threads go on, and everybody waits
// on the barrier's phase.  Used by k_cost (tables, rollout rows: replaces per-thread 16-byte cp.async loops) and by
// k_generate_mma (matrix rows).  Sizes and addresses are multiples of 16 bytes.
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra MBAR_WAIT_%=;\n"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// k_select_reuse: per problem, rank (cost, index) pairs ascending like std::sort on std::pair<double,int>;
// the extra (noise-less) rollout has index -1 and therefore wins ties.  out: src[b][j] for j < R_reuse
// (>=0: rollout slot of the previous iteration, -1: the extra rollout).
// ---------------------------------------------------------------------------------------------
__global__ void k_select_reuse(const double* __restrict__ totals, int R, int R_reuse, int use_extra,
                               int* __restrict__ reuse_src) {
  int b = blockIdx.x;
  const double* tot = totals + size_t(b) * (R + 1);
  int n = R + (use_extra ? 1 : 0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double ci = tot[i];
    int ii = (i == R) ? -1 : i;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      double cj = tot[j];
      int jj = (j == R) ? -1 : j;
      rank += (cj < ci) || (cj == ci && jj < ii);
    }
    if (rank < R_reuse) reuse_src[size_t(b) * R_reuse + rank] = ii;
  }
}

// k_select_gather (small batches): the reuse selection of k_select_reuse and, in the same launch, the reused slots filled from
// the candidate pass — k_generate ran over ALL previous rollouts (+ the noise-less one) as if each were reused while the
// noise-less rollout was still being costed, so nothing but this copy is left between that cost and the update.
// Grid (B, R_reuse): CTA (b, j) finds the candidate of rank j and copies its noise / M noise / control costs / parameters /
// state costs into slot R_gen + j.
struct SelectGatherArgs {
  int R, R_reuse, D, N, use_extra;
  const double* totals;        // [B][R + 1]
  int* reuse_src;              // [B][R_reuse]
  const double* cand_noise;    // [B][R + 1][D][N]   candidate R = the noise-less rollout
  const double* cand_control;  // [B][R + 1][D][N]
  const double* cand_y;        // [B][R + 1][D][N] or nullptr
  const double* params_prev;   // [B][R][D][N]
  const double* theta;         // [B][D][N]
  const double* state_prev;    // [B][R][N]
  const double* extra_state;   // [B][N]
  const double* extra_control; // [B][D][N] or nullptr.  Given: the noise-less rollout's total (candidate R) is summed here, exactly
                               // as k_extra_total sums it, and totals[R] is not read — k_extra_total then runs beside this kernel
  double* noise;               // [B][R][D][N]
  double* control;
  double* params;
  double* noise_projected;     // optional tap
  double* state;               // [B][R][N]
};

// (512 threads: the copies are a chain of dependent round trips per thread, so more threads = fewer trips; the sum of the
// noise-less rollout's total keeps k_extra_total's shape — 128 threads, stride 128, warp trees, four warps in order — with the
// loads of a thread issued together)
constexpr int kSelectGatherThreads = 512;
__global__ void __launch_bounds__(kSelectGatherThreads) k_select_gather(SelectGatherArgs a) {
  __shared__ int s_src;
  __shared__ double s_extra, sred[4], sred2[4];
  const int R = a.R, b = blockIdx.x / a.R_reuse, j = blockIdx.x - b * a.R_reuse;
  const double* tot = a.totals + size_t(b) * (R + 1);
  const int n = R + (a.use_extra ? 1 : 0);
  const bool own_extra = a.use_extra && a.extra_control != nullptr;
  if (own_extra) {     // k_extra_total's sum, operation for operation (128 threads): same bits as totals[R] will hold
    if (threadIdx.x < 128) {
      double acc = 0.0, accs = 0.0;
      const double* xs = a.extra_state + size_t(b) * a.N;
      const double* xc = a.extra_control + size_t(b) * a.D * a.N;
      const int DNx = a.D * a.N;
      if (DNx <= 8 * 128 && a.N <= 8 * 128) {
        double v[8], w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = threadIdx.x + q * 128;
          v[q] = i < DNx ? xc[i] : 0.0;
          w[q] = i < a.N ? xs[i] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = threadIdx.x + q * 128;
          if (i < a.N) accs += w[q];
          if (i < DNx) acc += v[q];
        }
      } else {
        for (int t = threadIdx.x; t < a.N; t += 128) accs += xs[t];
        for (int i = threadIdx.x; i < DNx; i += 128) acc += xc[i];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, off);
        accs += __shfl_xor_sync(0xffffffffu, accs, off);
      }
      if ((threadIdx.x & 31) == 0) { sred[threadIdx.x >> 5] = acc; sred2[threadIdx.x >> 5] = accs; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s1 = 0.0, s2 = 0.0;
      for (int w = 0; w < 4; ++w) s1 += sred[w], s2 += sred2[w];
      s_extra = s2 + s1;
    }
    __syncthreads();
  }
  auto total_of = [&](int i) -> double { return (own_extra && i == R) ? s_extra : tot[i]; };
  for (int i = threadIdx.x; i < n; i += blockDim.x) {     // the ranking of k_select_reuse
    const double ci = total_of(i);
    const int ii = (i == R) ? -1 : i;
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const double ck = total_of(k);
      const int kk = (k == R) ? -1 : k;
      rank += (ck < ci) || (ck == ci && kk < ii);
    }
    if (rank == j) s_src = ii;
  }
  __syncthreads();
  const int src = s_src;
  if (threadIdx.x == 0) a.reuse_src[size_t(b) * a.R_reuse + j] = src;
  const int DN = a.D * a.N, slot = R - a.R_reuse + j;
  const size_t cand = (size_t(b) * (R + 1) + (src >= 0 ? src : R)) * DN;
  const size_t dst = (size_t(b) * R + slot) * DN;
  const double* psrc = src >= 0 ? a.params_prev + (size_t(b) * R + src) * DN : a.theta + size_t(b) * DN;
  for (int k = threadIdx.x; k < DN; k += blockDim.x) {
    a.noise[dst + k] = a.cand_noise[cand + k];
    a.control[dst + k] = a.cand_control[cand + k];
    a.params[dst + k] = psrc[k];
    if (a.noise_projected) a.noise_projected[dst + k] = a.cand_y[cand + k];
  }
  const double* ssrc = src >= 0 ? a.state_prev + (size_t(b) * R + src) * a.N : a.extra_state + size_t(b) * a.N;
  double* sdst = a.state + (size_t(b) * R + slot) * a.N;
  for (int t = threadIdx.x; t < a.N; t += blockDim.x) sdst[t] = ssrc[t];
}

// k_advance_iteration (CUDA-graph replay of stomp_engine_run): what the host does between two iterations, on the device —
// the Philox generation counter advances and the noise scales sigma_d decay_d^(it-1) of the next iteration are taken from a
// table the host computed for the whole run (so they are bit-identical to the values the per-iteration upload would carry).
__global__ void k_advance_iteration(uint32_t* __restrict__ generation, int* __restrict__ table_index, const double* __restrict__ table,
                                    int D, double* __restrict__ noise_scale) {
  const int idx = *table_index;
  for (int d = threadIdx.x; d < D; d += blockDim.x) noise_scale[d] = table[size_t(idx) * D + d];
  __syncthreads();
  if (threadIdx.x == 0) {
    *generation += 1u;
    *table_index = idx + 1;
  }
}

constexpr int kMaxScaleArgs = 32;
struct GenArgs {
  int B, R, D, N;
  int R_gen;                 // slots < R_gen are new, slots >= R_gen are reused (gathered)
  int r_begin, r_count;      // rollout slots processed by this launch: [r_begin, r_begin + r_count)
  int mode_generate;         // 1: produce noise/params (new: sample or injected; reused: gather)
  int mode_project;          // 1: y = M * noise is added before the stencil; 0: stencil on params + eps_in
  int mode_control;          // 1: compute control costs
  int injected;              // 1: noise for new rollouts comes from eps_in
  int extra;                 // 1: vectors are (b, d) of the extra rollout: params = theta, noise = 0
  int pre;                   // 1: look-ahead pass for the NEXT iteration's new rollouts: only what does not depend on theta —
                             // noise (Philox) and noise_projected = M * noise — is produced, into compact [B][r_count][D][N]
                             // buffers (R = r_count here); k_finish_rollouts adds theta and the control costs later
  double control_weight;     // 0.5 * control_cost_weight
  uint64_t seed;
  uint32_t iteration;
  const uint32_t* iteration_ptr;   // non-null (CUDA-graph replay): the Philox generation counter lives on the device
  int64_t rollout_id_offset; // global rollout id of local slot 0 (rollout sharding)
  int64_t rollouts_global;   // global number of rollouts per problem (stream id stride)
  const double* theta;       // [B][D][N]
  const double* pad_start;   // [B][D]
  const double* pad_goal;    // [B][D]
  const double* noise_scale; // [D] sigma_d * decay_d^(it-1)
  int scale_by_value;        // 1: the scales travel in scale_v (D <= kMaxScaleArgs) — no upload on the launching stream
  double scale_v[32];
  const double* eps_in;      // [B][R][D][N] injected noise (or caller noise when !mode_project)
  const double* params_prev; // [B][prev_stride][D][N] previous-iteration parameters (reuse gather)
  // k_generate_seg: the band solves split over `segs` time segments per vector (spike tables, engine.cu)
  int segs;
  const double* seg_hf;      // [N][6] forward solve:  x_i = local_i + sum_j hf[i][j] x_{start-1-j}
  const double* seg_hb;      // [N][6] backward solve: x_i = local_i + sum_j hb[i][j] x_{end+j}
  int prev_stride;           // rollouts per problem in params_prev (= R except in the candidate pass, where R counts candidates)
  const int* reuse_src;      // [B][R_reuse]
  double* noise;             // [B][R][D][N]
  double* params;            // [B][R][D][N]  (extra: unused)
  double* noise_projected;   // optional tap
  double* control;           // [B][R][D][N]  (extra: [B][D][N])
  double* csum;              // optional [B][R][D]: sum over t of the vector's control costs, in t order (k_generate only; feeds
                             // k_totals, which then replaces k_cumulative's pass over the control costs)
  double* scratch;           // time-major [N][scratch_stride] work buffer
  size_t scratch_stride;     // >= number of vectors, multiple of the CTA size
  Band band;
  Stencil st;
};

// the noise scale of dimension d: from the launch arguments (constant bank) or from the device array
__device__ __forceinline__ double gen_noise_scale(const GenArgs& a, int d) { return a.scale_by_value ? a.scale_v[d] : a.noise_scale[d]; }


// Banded triangular solves with R = C C^T.  The tables hold, per row, the inverse diagonal and the six
// sub-diagonal coefficients already multiplied by it (zero beyond the bandwidth), so one step is
//   x_i = fma(-c1, x_{i-1}, fma(-c2, x_{i-2}, ... inv_i * b_i))
// with the last outputs kept in registers: the dependent chain is ONE DFMA per step.
//   fw[i] = {1/C(i,i), C(i,i-1)/C(i,i), ..., C(i,i-6)/C(i,i), 0}      (forward:  C   x = b)
//   bw[i] = {1/C(i,i), C(i+1,i)/C(i,i), ..., C(i+6,i)/C(i,i), 0}      (backward: C^T x = b)
// (built once per request on the host, engine.cu)
__device__ __forceinline__ void band_forward(double* x, const double* fw, int N) {
  double w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0, w6 = 0;
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    const double2* c = reinterpret_cast<const double2*>(fw + i * 8);
    double2 c01 = c[0], c23 = c[1], c45 = c[2], c67 = c[3];
    double s = c01.x * x[i];
    s = fma(-c67.x, w6, s);
    s = fma(-c45.y, w5, s);
    s = fma(-c45.x, w4, s);
    s = fma(-c23.y, w3, s);
    s = fma(-c23.x, w2, s);
    s = fma(-c01.y, w1, s);
    x[i] = s;
    w6 = w5; w5 = w4; w4 = w3; w3 = w2; w2 = w1; w1 = s;
  }
}

__device__ __forceinline__ void band_backward(double* x, const double* bw, int N) {
  double w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0, w6 = 0;
#pragma unroll 4
  for (int i = N - 1; i >= 0; --i) {
    const double2* c = reinterpret_cast<const double2*>(bw + i * 8);
    double2 c01 = c[0], c23 = c[1], c45 = c[2], c67 = c[3];
    double s = c01.x * x[i];
    s = fma(-c67.x, w6, s);
    s = fma(-c45.y, w5, s);
    s = fma(-c45.x, w4, s);
    s = fma(-c23.y, w3, s);
    s = fma(-c23.x, w2, s);
    s = fma(-c01.y, w1, s);
    x[i] = s;
    w6 = w5; w5 = w4; w4 = w3; w3 = w2; w2 = w1; w1 = s;
  }
}

// Control cost of one padded row from its seven taps w[j] = x_all[p - 3 + j] (dropped taps = 0):
// sum_k 0.5 w_control w_k (sum_j rule_k[j] x_all[p - 3 + j])^2   (covariant_trajectory_policy.cpp:228-255).
// One definition with explicit fused multiply-adds, shared by the generation kernels and k_finish_rollouts: whichever of them
// produced a rollout's control costs, the bits are the same.
__device__ __forceinline__ double stencil_row_cost(const Stencil& st, double control_weight, const double* w) {
  double cost = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (st.weight[k] == 0.0) continue;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) acc = fma(st.coef[k][j], w[j], acc);
    cost = fma(control_weight * st.weight[k], acc * acc, cost);
  }
  return cost;
}

// ---------------------------------------------------------------------------------------------
// k_generate: streaming, one thread per vector v = (b, r, d) (or (b, d) for the extra rollout).
//
// Every step of the band solves / stencils is sequential in time, so a thread walks its own series.  To keep
// all global traffic coalesced without parking the whole series in shared memory (which capped occupancy at
// ~8 warps/SM), the kernel works in sweeps:
//   A (backward in t)  noise: eps = sigma C^-T z from the Philox stream | injected | parameters - theta (reused);
//                      writes noise and parameters
//   B (forward)        w = C^-1 (s .* eps)                         -> scratch, TIME-MAJOR [t][vector]
//   C (backward)       y = C^-T w = M eps ; x = parameters + y     -> scratch (in place)
//   D (forward)        7-tap control-cost stencils over [pads, x, pads] with a sliding register window; writes
//                      control costs
// Row-major arrays ([vector][t]) are read / written through per-warp 32 x 16 transposition tiles in shared
// memory (128-byte row segments); the time-major scratch is accessed directly (lane = vector -> coalesced).
// ---------------------------------------------------------------------------------------------
#ifndef STOMP_GEN_CHUNK
#define STOMP_GEN_CHUNK 8   // timesteps per transposition tile (A/B on B200: 8 -> 0.169 ms, 16 -> 0.178 ms)
#endif
constexpr int kChunk = STOMP_GEN_CHUNK;
constexpr int kTileLd = kChunk + 1;

__device__ __forceinline__ const double* shfl_ptr(const double* p, int src_lane) {
  unsigned long long v = reinterpret_cast<unsigned long long>(p);
  v = __shfl_sync(0xffffffffu, v, src_lane);
  return reinterpret_cast<const double*>(v);
}

// tile[row][col] = row_ptr(row)[c0 + col] for the 32 rows whose base pointers the lanes hold (nullptr -> 0).
// The copies are asynchronous (cp.async, LDGSTS): all 16 row segments of a tile - and of a second tile issued
// right after - are in flight together, so a chunk pays one memory latency instead of one per row pair.
__device__ __forceinline__ void warp_tile_load_async(double* tile, const double* my_row, int c0, int len, int lane) {
#pragma unroll
  for (int e = lane; e < 32 * kChunk; e += 32) {
    const int row = e / kChunk, col = e % kChunk;
    const double* rp = shfl_ptr(my_row, row);
    double* dst = tile + row * kTileLd + col;
    if (rp != nullptr && col < len) {
      const unsigned saddr = static_cast<unsigned>(__cvta_generic_to_shared(dst));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(rp + c0 + col) : "memory");
    } else {
      *dst = 0.0;
    }
  }
}
__device__ __forceinline__ void warp_tile_wait(void) {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void warp_tile_load(double* tile, const double* my_row, int c0, int len, int lane) {
  __syncwarp();
  warp_tile_load_async(tile, my_row, c0, len, lane);
  warp_tile_wait();
}
__device__ __forceinline__ void warp_tile_load2(double* t0, const double* row0, double* t1, const double* row1, int c0, int len,
                                                int lane) {
  __syncwarp();
  warp_tile_load_async(t0, row0, c0, len, lane);
  warp_tile_load_async(t1, row1, c0, len, lane);
  warp_tile_wait();
}

__device__ __forceinline__ void warp_tile_store(const double* tile, double* my_row, int c0, int len, int lane) {
  __syncwarp();
#pragma unroll
  for (int e = lane; e < 32 * kChunk; e += 32) {
    const int row = e / kChunk, col = e % kChunk;
    double* rp = const_cast<double*>(shfl_ptr(my_row, row));
    if (rp != nullptr && col < len) rp[c0 + col] = tile[row * kTileLd + col];
  }
  __syncwarp();
}

struct BandWindow {
  double w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0, w6 = 0;
  // one step of a banded triangular solve: row = {1/diag, c1/diag, ..., c6/diag, 0}; returns x_i
  __device__ __forceinline__ double step(const double* row, double rhs) {
    const double2* c = reinterpret_cast<const double2*>(row);
    const double2 c01 = c[0], c23 = c[1], c45 = c[2], c67 = c[3];
    double s = c01.x * rhs;
    s = fma(-c67.x, w6, s);
    s = fma(-c45.y, w5, s);
    s = fma(-c45.x, w4, s);
    s = fma(-c23.y, w3, s);
    s = fma(-c23.x, w2, s);
    s = fma(-c01.y, w1, s);
    w6 = w5; w5 = w4; w4 = w3; w3 = w2; w2 = w1; w1 = s;
    return s;
  }
};

template <bool kPre>
__global__ void __launch_bounds__(128, STOMP_GEN_MIN_BLOCKS) k_generate(GenArgs a) {
  extern __shared__ double smem[];
  const int N = a.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sfw = smem;                        // [N][8]
  double* sbw = sfw + N * 8;                 // [N][8]
  double* sscale = sbw + N * 8;              // [N]
  double* tE = sscale + N + size_t(warp) * 2 * 32 * kTileLd;   // per-warp tiles
  double* tT = tE + 32 * kTileLd;
  for (int k = threadIdx.x; k < N * 8; k += blockDim.x) sfw[k] = a.band.fw[k], sbw[k] = a.band.bw[k];
  for (int i = threadIdx.x; i < N; i += blockDim.x) sscale[i] = a.band.proj_scale[i];
  __syncthreads();

  const int per_problem = (a.extra ? 1 : a.r_count) * a.D;
  const long long nvec = (long long)a.B * per_problem;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = v < nvec;
  int b = 0, r = 0, d = 0;
  if (active) {
    b = int(v / per_problem);
    int rem = int(v - (long long)b * per_problem);
    r = rem / a.D;
    d = rem - r * a.D;
    r += a.extra ? 0 : a.r_begin;
  }
  const size_t row_off = a.extra ? (size_t(b) * a.D + d) * N : ((size_t(b) * a.R + r) * a.D + d) * N;
  const double* th_row = (active && !kPre) ? a.theta + (size_t(b) * a.D + d) * N : nullptr;
  double* wb = a.scratch + v;                 // time-major scratch: element i at wb[i * sstride]
  const size_t sstride = a.scratch_stride;
  const int last_c0 = ((N - 1) / kChunk) * kChunk;
  double* my_tE = tE + lane * kTileLd;
  double* my_tT = tT + lane * kTileLd;

  // ---- sweep A: noise and parameters --------------------------------------------------------------------
  if (a.mode_generate && !a.extra) {
    const bool is_new = r < a.R_gen;
    const bool philox = active && is_new && !a.injected;
    const double* src = nullptr;
    if (active) {
      if (!is_new) {
        int sidx = a.reuse_src[size_t(b) * (a.R - a.R_gen) + (r - a.R_gen)];
        src = sidx >= 0 ? a.params_prev + ((size_t(b) * a.prev_stride + sidx) * a.D + d) * N : th_row;
      } else if (a.injected) {
        src = a.eps_in + row_off;
      }
    }
    double* out_noise = active ? a.noise + row_off : nullptr;
    double* out_params = active ? a.params + row_off : nullptr;
    const double sg = active ? gen_noise_scale(a, d) : 0.0;
    const uint64_t stream = (uint64_t(b) * uint64_t(a.rollouts_global) + uint64_t(a.rollout_id_offset + r)) * uint64_t(a.D) + d;
    const uint32_t gen_iteration = a.iteration_ptr ? *a.iteration_ptr : a.iteration;
    BandWindow bw;
    for (int c0 = last_c0; c0 >= 0; c0 -= kChunk) {
      const int len = min(kChunk, N - c0);
      warp_tile_load2(tE, src, tT, th_row, c0, len, lane);
      if (philox) {
        // eps = sigma * C^-T z  ~ N(0, sigma^2 R^-1)  with R = C C^T.
        // 1. the chunk's standard normals: independent Philox / Box-Muller evaluations (ILP), parked in the tile row
        {
          const int p_lo = c0 >> 1, p_hi = (c0 + len - 1) >> 1;
          for (int pr = p_lo; pr <= p_hi; ++pr) {
            double z0, z1;
            normal_pair(a.seed, stream, gen_iteration, uint32_t(pr), z0, z1);
            const int i0 = 2 * pr - c0, i1 = i0 + 1;
            if (i0 >= 0 && i0 < len) my_tE[i0] = z0;
            if (i1 >= 0 && i1 < len) my_tE[i1] = z1;
          }
        }
        // 2. the sequential back-substitution (one DFMA on the critical path per step)
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
          const int k = len - 1 - kk;
          if (k >= 0) {
            const double e = sg * bw.step(sbw + (c0 + k) * 8, my_tE[k]);
            my_tT[k] += e;          // parameters = theta + noise
            my_tE[k] = e;
          }
        }
      } else if (active && is_new) {   // injected noise
        for (int k = 0; k < len; ++k) my_tT[k] += my_tE[k];
      } else if (active) {             // reused rollout: noise = parameters - theta (policy_improvement.cpp:222)
        for (int k = 0; k < len; ++k) {
          const double pv = my_tE[k];
          my_tE[k] = pv - my_tT[k];
          my_tT[k] = pv;
        }
      }
      warp_tile_store(tE, out_noise, c0, len, lane);
      if (!kPre) warp_tile_store(tT, out_params, c0, len, lane);
    }
  }
  if (!a.mode_control) return;

  // ---- sweeps B + C: x = parameters + M * noise into the time-major scratch -----------------------------------
  if (a.mode_project && !a.extra) {
    const double* nrow = active ? a.noise + row_off : nullptr;
    {
      BandWindow fwd;
      double* wp = wb;
      for (int c0 = 0; c0 < N; c0 += kChunk) {
        const int len = min(kChunk, N - c0);
        warp_tile_load(tE, nrow, c0, len, lane);
        if (active) {
#pragma unroll
          for (int k = 0; k < kChunk; ++k)
            if (k < len) {
              *wp = fwd.step(sfw + (c0 + k) * 8, sscale[c0 + k] * my_tE[k]);
              wp += sstride;
            }
        }
      }
    }
    {
      const double* prow = (active && !kPre) ? a.params + row_off : nullptr;
      double* ytap = (active && a.noise_projected) ? a.noise_projected + row_off : nullptr;
      BandWindow bwd;
      for (int c0 = last_c0; c0 >= 0; c0 -= kChunk) {
        const int len = min(kChunk, N - c0);
        if (!kPre) warp_tile_load(tT, prow, c0, len, lane);
        if (active) {
          double* wp = wb + size_t(c0) * sstride;
          double wv[kChunk];
#pragma unroll
          for (int k = 0; k < kChunk; ++k)       // independent loads first: one memory latency per chunk
            if (k < len) wv[k] = wp[size_t(k) * sstride];
#pragma unroll
          for (int kk = 0; kk < kChunk; ++kk) {
            const int k = kChunk - 1 - kk;
            if (k < len) {
              const double y = bwd.step(sbw + (c0 + k) * 8, wv[k]);
              my_tE[k] = y;
              if (!kPre) wv[k] = my_tT[k] + y;
            }
          }
          if (!kPre) {
#pragma unroll
            for (int k = 0; k < kChunk; ++k)
              if (k < len) wp[size_t(k) * sstride] = wv[k];
          }
        }
        if (a.noise_projected) warp_tile_store(tE, ytap, c0, len, lane);
      }
    }
    if (kPre) return;
  } else {
    // no projection: x = theta (extra rollout) or x = parameters + caller noise (Policy::computeControlCosts)
    const double* prow = active ? (a.extra ? th_row : a.params + row_off) : nullptr;
    const double* erow = (active && !a.extra) ? a.eps_in + row_off : nullptr;
    for (int c0 = 0; c0 < N; c0 += kChunk) {
      const int len = min(kChunk, N - c0);
      warp_tile_load2(tT, prow, tE, erow, c0, len, lane);
      if (active) {
        double* wp = wb + size_t(c0) * sstride;
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) wp[size_t(k) * sstride] = my_tT[k] + my_tE[k];
      }
    }
  }

  // ---- sweep D: control-cost stencils over the padded series (covariant_trajectory_policy.cpp:228-255) -------
  {
    const int Nall = N + 2 * kPad;
    double xs = 0.0, xg = 0.0;
    if (active) { xs = a.pad_start[size_t(b) * a.D + d]; xg = a.pad_goal[size_t(b) * a.D + d]; }
    auto xall = [&](int idx) -> double {
      if (idx < 0 || idx >= Nall) return 0.0;   // dropped taps of the differentiation matrices
      return idx < kPad ? xs : (idx >= kPad + N ? xg : wb[size_t(idx - kPad) * sstride]);
    };
    auto stencil_cost = [&](const double* w) -> double { return stencil_row_cost(a.st, a.control_weight, w); };
    double hc[kPad], tc[kPad], w[7];
    if (active) {
      // the six trailing padded rows only see the last three free values and the goal padding
#pragma unroll
      for (int q = 0; q < kPad; ++q) {
        const int p = kPad + N + q;
        double ww[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) ww[j] = xall(p + j - 3);
        tc[q] = stencil_cost(ww);
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) w[j] = xall(j - 3);
#pragma unroll
      for (int p = 0; p < kPad; ++p) {   // leading padded rows
        hc[p] = stencil_cost(w);
        const double next = xall(p + 4);
#pragma unroll
        for (int j = 0; j < 6; ++j) w[j] = w[j + 1];
        w[6] = next;
      }
    }
    double* crow = active ? a.control + row_off : nullptr;
    double csum = 0.0;
    for (int c0 = 0; c0 < N; c0 += kChunk) {
      const int len = min(kChunk, N - c0);
      __syncwarp();
      if (active) {
        // look-ahead values x_all[t + kPad + 4] of the chunk: independent loads, one memory latency per chunk
        double nx[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k) {
          const int f = c0 + k + 4;           // free index of the look-ahead element
          nx[k] = 0.0;
          if (k < len) nx[k] = f < N ? wb[size_t(f) * sstride] : (f < N + kPad ? xg : 0.0);
        }
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) {
            const int t = c0 + k;
            double cost = stencil_cost(w);
            if (t == 0) {
#pragma unroll
              for (int q = 0; q < kPad; ++q) cost += hc[q];
            }
            if (t == N - 1) {
#pragma unroll
              for (int q = 0; q < kPad; ++q) cost += tc[kPad - 1 - q];
            }
            my_tE[k] = cost;
            csum += cost;
#pragma unroll
            for (int j = 0; j < 6; ++j) w[j] = w[j + 1];
            w[6] = nx[k];
          }
      }
      warp_tile_store(tE, crow, c0, len, lane);
    }
    if (active && a.csum && !a.extra) a.csum[(size_t(b) * a.R + r) * a.D + d] = csum;
  }
}

// ---------------------------------------------------------------------------------------------
// k_generate_seg: k_generate with every band solve split over P time segments (P warps of the CTA; lane = vector, warp =
// segment): an experiment against k_generate's one-thread-per-vector chains of ~4 N dependent steps (C2: 8 warps per SM, 11 %
// of the warp slots).
//
// A banded triangular solve restricted to a segment needs only the six values next to the segment from its neighbour.  Each
// segment first solves with those taken as zero (the LOCAL solution), the true first / last six values of the segments are
// then fixed in a chain of P - 1 steps of 36 FMAs, and every row is corrected independently:
//   x_i = local_i + sum_j H[i][j] X_j,   X = the neighbour's six final boundary values,
// where H (the "spikes", host-computed in fp64 from the same band tables) is the response of the segment's rows to unit
// boundary values.  The correction of one solve is folded into the pass that runs the next one:
//   A1  backward local solve of C^T e = z (Philox)                                -> scratch (time-major)
//   A2  boundary chain (backward)
//   A3+B1  forward over the segment: e final, noise = sigma e, parameters; local forward solve of C w = s .* noise -> scratch
//   B2  boundary chain (forward)
//   B3+C1  backward: w final; local backward solve of C^T y = w                   -> scratch
//   C2  boundary chain (backward)
//   C3  y final, x = parameters + y                                              -> scratch;  CTA barrier
//   D   control-cost stencils over [pads, x, pads] (neighbouring segments' x from the scratch)
// Five passes over N / P steps instead of four over N.  Same linear maps as k_generate, evaluated in another order: the
// results agree to rounding (test_dense_generation_kernels_match_the_band_solves), not bit for bit.
// Measured slower than k_generate at every batch shape tried (engine.cu, launch_generate_range): opt-in, STOMP_GENERATE=seg.
// ---------------------------------------------------------------------------------------------
constexpr int kSegBand = 6;
#ifndef STOMP_SEG_MIN_BLOCKS
#define STOMP_SEG_MIN_BLOCKS 2
#endif
// segment boundaries on multiples of the transposition chunk (aligned tiles, Philox pairs never split); the last one ends at N
__host__ __device__ inline int seg_start(int seg, int N, int P) {
  if (seg >= P) return N;
  return int(((long long)seg * N / P + kChunk / 2) / kChunk) * kChunk;
}

template <bool kBackward>
__device__ __forceinline__ void seg_chain(double* bnd, int seg, int P, int lane, const double* loc, const double* H, int row0, int row_step,
                                          double* neighbour) {
  // loc[m]: this segment's local boundary values (m = 0 nearest its own interior end ... see callers); bnd[seg][m][lane]
  // receives the final ones.  kBackward: segment seg depends on seg + 1; else on seg - 1.  H row of boundary value m:
  // H + (row0 + m * row_step) * 6.
  for (int step = 0; step < P; ++step) {
    const int turn = kBackward ? P - 1 - step : step;
    if (seg == turn) {
      const bool has = kBackward ? seg < P - 1 : seg > 0;
      double fin[kSegBand];
#pragma unroll
      for (int m = 0; m < kSegBand; ++m) fin[m] = loc[m];
      if (has) {
        const double* nb = bnd + size_t(kBackward ? seg + 1 : seg - 1) * kSegBand * 32 + lane;
#pragma unroll
        for (int j = 0; j < kSegBand; ++j) neighbour[j] = nb[j * 32];
#pragma unroll
        for (int m = 0; m < kSegBand; ++m) {
          const double* h = H + size_t(row0 + m * row_step) * kSegBand;
#pragma unroll
          for (int j = 0; j < kSegBand; ++j) fin[m] = fma(h[j], neighbour[j], fin[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < kSegBand; ++m) bnd[(size_t(seg) * kSegBand + m) * 32 + lane] = fin[m];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256, STOMP_SEG_MIN_BLOCKS) k_generate_seg(GenArgs a) {
  extern __shared__ double smem[];
  const int N = a.N;
  const int lane = threadIdx.x & 31, P = a.segs, nwarps = blockDim.x >> 5, G = nwarps / P;   // G groups of 32 vectors per CTA
  const int seg = (threadIdx.x >> 5) % P, grp = (threadIdx.x >> 5) / P;
  double* sfw = smem;                        // [N][8]
  double* sbw = sfw + N * 8;                 // [N][8]
  double* shf = sbw + N * 8;                 // [N][6]
  double* shb = shf + N * kSegBand;          // [N][6]
  double* sscale = shb + N * kSegBand;       // [N]
  double* bnd = sscale + N + size_t(grp) * P * kSegBand * 32;   // [G][P][6][32]: the boundary exchange of this vector group
  double* tE = sscale + N + size_t(G) * P * kSegBand * 32 + size_t(threadIdx.x >> 5) * 2 * 32 * kTileLd;   // per-warp tiles
  double* tT = tE + 32 * kTileLd;
  for (int k = threadIdx.x; k < N * 8; k += blockDim.x) sfw[k] = a.band.fw[k], sbw[k] = a.band.bw[k];
  for (int k = threadIdx.x; k < N * kSegBand; k += blockDim.x) shf[k] = a.seg_hf[k], shb[k] = a.seg_hb[k];
  for (int i = threadIdx.x; i < N; i += blockDim.x) sscale[i] = a.band.proj_scale[i];
  __syncthreads();

  const int per_problem = a.r_count * a.D;
  const long long nvec = (long long)a.B * per_problem;
  const long long v = ((long long)blockIdx.x * G + grp) * 32 + lane;
  const bool active = v < nvec;
  int b = 0, r = 0, d = 0;
  if (active) {
    b = int(v / per_problem);
    const int rem = int(v - (long long)b * per_problem);
    r = rem / a.D;
    d = rem - r * a.D;
    r += a.r_begin;
  }
  const size_t row_off = ((size_t(b) * a.R + r) * a.D + d) * N;
  const double* th_row = active ? a.theta + (size_t(b) * a.D + d) * N : nullptr;
  double* wb = a.scratch + v;                 // time-major scratch: element i at wb[i * sstride]
  const size_t sstride = a.scratch_stride;
  const int s0 = seg_start(seg, N, P), s1 = seg_start(seg + 1, N, P);   // this warp's time segment [s0, s1)
  const int seg_last_c0 = s0 + ((s1 - s0 - 1) / kChunk) * kChunk;
  double* my_tE = tE + lane * kTileLd;
  double* my_tT = tT + lane * kTileLd;
  const bool is_new = r < a.R_gen;
  const bool philox = active && is_new && !a.injected;
  const bool has_next = seg < P - 1, has_prev = seg > 0;
  double nxt[kSegBand], prv[kSegBand], loc[kSegBand];
#pragma unroll
  for (int j = 0; j < kSegBand; ++j) nxt[j] = prv[j] = 0.0;

  // ---- A1: local backward solve of C^T e = z over the segment ------------------------------------------------------------
  {
    const uint64_t stream = (uint64_t(b) * uint64_t(a.rollouts_global) + uint64_t(a.rollout_id_offset + r)) * uint64_t(a.D) + d;
    const uint32_t gen_iteration = a.iteration_ptr ? *a.iteration_ptr : a.iteration;
    BandWindow bw;
    if (philox) {
      for (int c0 = seg_last_c0; c0 >= s0; c0 -= kChunk) {
        const int len = min(kChunk, s1 - c0);
        double z[kChunk];
        {
          const int p_lo = c0 >> 1, p_hi = (c0 + len - 1) >> 1;
          for (int pr = p_lo; pr <= p_hi; ++pr) {
            double z0, z1;
            normal_pair(a.seed, stream, gen_iteration, uint32_t(pr), z0, z1);
            const int i0 = 2 * pr - c0, i1 = i0 + 1;
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
              if (k == i0) z[k] = z0;
              if (k == i1) z[k] = z1;
            }
          }
        }
        double* wp = wb + size_t(c0) * sstride;
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
          const int k = kChunk - 1 - kk;
          if (k < len) wp[size_t(k) * sstride] = bw.step(sbw + (c0 + k) * 8, z[k]);
        }
      }
    }
    loc[0] = bw.w1; loc[1] = bw.w2; loc[2] = bw.w3; loc[3] = bw.w4; loc[4] = bw.w5; loc[5] = bw.w6;   // e_local[s0 + m]
  }
  // ---- A2: the segments' true first six values, last segment first ---------------------------------------------------------
  seg_chain<true>(bnd, seg, P, lane, loc, shb, s0, 1, nxt);
  if (has_next) {
#pragma unroll
    for (int j = 0; j < kSegBand; ++j) nxt[j] = bnd[(size_t(seg + 1) * kSegBand + j) * 32 + lane];
  }
  __syncthreads();   // the exchange array is reused by the next chain

  // ---- A3 + B1: noise and parameters; local forward solve of C w = s .* noise ------------------------------------------------
  {
    const double* src = nullptr;
    if (active) {
      if (!is_new) {
        const int sidx = a.reuse_src[size_t(b) * (a.R - a.R_gen) + (r - a.R_gen)];
        src = sidx >= 0 ? a.params_prev + ((size_t(b) * a.prev_stride + sidx) * a.D + d) * N : th_row;
      } else if (a.injected) {
        src = a.eps_in + row_off;
      }
    }
    double* out_noise = active ? a.noise + row_off : nullptr;
    double* out_params = active ? a.params + row_off : nullptr;
    const double sg = active ? gen_noise_scale(a, d) : 0.0;
    BandWindow fwd;
    for (int c0 = s0; c0 < s1; c0 += kChunk) {
      const int len = min(kChunk, s1 - c0);
      warp_tile_load2(tE, src, tT, th_row, c0, len, lane);
      double* wp = wb + size_t(c0) * sstride;
      double ev[kChunk];
      if (philox) {
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) ev[k] = wp[size_t(k) * sstride];
      }
      if (active) {
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) {
            const int t = c0 + k;
            double eps, par;
            if (philox) {
              double e = ev[k];
              if (has_next) {
                const double* h = shb + t * kSegBand;
#pragma unroll
                for (int j = 0; j < kSegBand; ++j) e = fma(h[j], nxt[j], e);
              }
              eps = sg * e;
              par = my_tT[k] + eps;
            } else if (is_new) {      // injected noise
              eps = my_tE[k];
              par = my_tT[k] + eps;
            } else {                  // reused rollout: noise = parameters - theta (policy_improvement.cpp:222)
              par = my_tE[k];
              eps = par - my_tT[k];
            }
            my_tE[k] = eps;
            my_tT[k] = par;
            if (a.mode_control) ev[k] = fwd.step(sfw + t * 8, sscale[t] * eps);
          }
        if (a.mode_control) {
#pragma unroll
          for (int k = 0; k < kChunk; ++k)
            if (k < len) wp[size_t(k) * sstride] = ev[k];
        }
      }
      warp_tile_store(tE, out_noise, c0, len, lane);
      warp_tile_store(tT, out_params, c0, len, lane);
    }
    loc[0] = fwd.w1; loc[1] = fwd.w2; loc[2] = fwd.w3; loc[3] = fwd.w4; loc[4] = fwd.w5; loc[5] = fwd.w6;     // w_local[s1 - 1 - m]
  }
  if (!a.mode_control) return;
  // ---- B2: the segments' true last six values, first segment first --------------------------------------------------------
  seg_chain<false>(bnd, seg, P, lane, loc, shf, s1 - 1, -1, prv);
  if (has_prev) {
#pragma unroll
    for (int j = 0; j < kSegBand; ++j) prv[j] = bnd[(size_t(seg - 1) * kSegBand + j) * 32 + lane];
  }
  __syncthreads();

  // ---- B3 + C1: w final; local backward solve of C^T y = w --------------------------------------------------------------
  {
    BandWindow bwd;
    if (active) {
      for (int c0 = seg_last_c0; c0 >= s0; c0 -= kChunk) {
        const int len = min(kChunk, s1 - c0);
        double* wp = wb + size_t(c0) * sstride;
        double wv[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) wv[k] = wp[size_t(k) * sstride];
#pragma unroll
        for (int kk = 0; kk < kChunk; ++kk) {
          const int k = kChunk - 1 - kk;
          if (k < len) {
            const int t = c0 + k;
            double w = wv[k];
            if (has_prev) {
              const double* h = shf + t * kSegBand;
#pragma unroll
              for (int j = 0; j < kSegBand; ++j) w = fma(h[j], prv[j], w);
            }
            wv[k] = bwd.step(sbw + t * 8, w);
          }
        }
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) wp[size_t(k) * sstride] = wv[k];
      }
    }
    loc[0] = bwd.w1; loc[1] = bwd.w2; loc[2] = bwd.w3; loc[3] = bwd.w4; loc[4] = bwd.w5; loc[5] = bwd.w6;     // y_local[s0 + m]
  }
  // ---- C2 ------------------------------------------------------------------------------------------------------------------
  seg_chain<true>(bnd, seg, P, lane, loc, shb, s0, 1, nxt);
  if (has_next) {
#pragma unroll
    for (int j = 0; j < kSegBand; ++j) nxt[j] = bnd[(size_t(seg + 1) * kSegBand + j) * 32 + lane];
  }

  // ---- C3: y final, x = parameters + y into the scratch ----------------------------------------------------------------------
  {
    const double* prow = active ? a.params + row_off : nullptr;
    double* ytap = (active && a.noise_projected) ? a.noise_projected + row_off : nullptr;
    for (int c0 = s0; c0 < s1; c0 += kChunk) {
      const int len = min(kChunk, s1 - c0);
      warp_tile_load(tT, prow, c0, len, lane);
      if (active) {
        double* wp = wb + size_t(c0) * sstride;
        double yv[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) yv[k] = wp[size_t(k) * sstride];
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) {
            double y = yv[k];
            if (has_next) {
              const double* h = shb + (c0 + k) * kSegBand;
#pragma unroll
              for (int j = 0; j < kSegBand; ++j) y = fma(h[j], nxt[j], y);
            }
            my_tE[k] = y;
            wp[size_t(k) * sstride] = my_tT[k] + y;
          }
      }
      if (a.noise_projected) warp_tile_store(tE, ytap, c0, len, lane);
    }
  }
  __syncthreads();   // the stencils read the neighbouring segments' x

  // ---- D: control-cost stencils over the padded series (covariant_trajectory_policy.cpp:228-255) -----------------------------
  {
    const int Nall = N + 2 * kPad;
    double xs = 0.0, xg = 0.0;
    if (active) { xs = a.pad_start[size_t(b) * a.D + d]; xg = a.pad_goal[size_t(b) * a.D + d]; }
    auto xall = [&](int idx) -> double {
      if (idx < 0 || idx >= Nall) return 0.0;   // dropped taps of the differentiation matrices
      return idx < kPad ? xs : (idx >= kPad + N ? xg : wb[size_t(idx - kPad) * sstride]);
    };
    auto stencil_cost = [&](const double* w) -> double { return stencil_row_cost(a.st, a.control_weight, w); };
    double w[7];
    if (active) {
      if (seg == P - 1) {       // the six trailing padded rows, added to the last free row in the order k_generate adds them
        double tc[kPad];
#pragma unroll
        for (int q = 0; q < kPad; ++q) {
          double ww[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) ww[j] = xall(kPad + N + q + j - 3);
          tc[q] = stencil_cost(ww);
        }
#pragma unroll
        for (int q = 0; q < kPad; ++q) loc[q] = tc[q];      // (loc / prv are free again: tail / head row costs)
      }
      if (seg == 0) {           // the six leading padded rows
#pragma unroll
        for (int q = 0; q < kPad; ++q) {
          double ww[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) ww[j] = xall(q + j - 3);
          prv[q] = stencil_cost(ww);
        }
      }
#pragma unroll
      for (int j = 0; j < 7; ++j) w[j] = xall(s0 + kPad + j - 3);
    }
    double* crow = active ? a.control + row_off : nullptr;
    for (int c0 = s0; c0 < s1; c0 += kChunk) {
      const int len = min(kChunk, s1 - c0);
      __syncwarp();
      if (active) {
        double nx[kChunk];
#pragma unroll
        for (int k = 0; k < kChunk; ++k) {
          const int f = c0 + k + 4;           // free index of the look-ahead element
          nx[k] = 0.0;
          if (k < len) nx[k] = f < N ? wb[size_t(f) * sstride] : (f < N + kPad ? xg : 0.0);
        }
#pragma unroll
        for (int k = 0; k < kChunk; ++k)
          if (k < len) {
            const int t = c0 + k;
            double cost = stencil_cost(w);
            if (t == 0) {
#pragma unroll
              for (int q = 0; q < kPad; ++q) cost += prv[q];
            }
            if (t == N - 1) {
#pragma unroll
              for (int q = 0; q < kPad; ++q) cost += loc[kPad - 1 - q];
            }
            my_tE[k] = cost;
#pragma unroll
            for (int j = 0; j < 6; ++j) w[j] = w[j + 1];
            w[6] = nx[k];
          }
      }
      warp_tile_store(tE, crow, c0, len, lane);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_generate_dense: the same noise / parameters / M*noise / control costs as k_generate for SMALL batches, where k_generate's
// one-thread-per-vector band solves are a serial chain of ~4 N steps with nothing to hide it behind (one planning problem:
// 35 vectors).  One CTA per vector, one thread per timestep, and the two triangular solves written as the dense products
// they stand for — exactly the reference's formulation (MultivariateGaussian: L z; projection: M eps):
//   eps_t = sigma * sum_{j >= t} Cinv[j][t] z_j          (eps = sigma C^-T z, the same linear map and the same Philox z
//                                                          as the band path: realisations agree to rounding)
//   y_t   = sum_j Ms[j][t] eps_j,  Ms[j][t] = R^-1[j][t] s_j  (R^-1 symmetric: rows read with t across the lanes)
// followed by the control-cost stencils over [pads, parameters + y, pads].  N^2 flops per vector instead of ~20 N, but a
// few microseconds of latency instead of ~N microseconds.  engine.cu picks the kernel from the batch size.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_generate_dense(GenArgs a, const double* __restrict__ cinv, const double* __restrict__ ms) {
  extern __shared__ double dsm[];
  const int N = a.N, Nall = N + 2 * kPad;
  double* z = dsm;            // [N] standard normals, then reused for nothing else
  double* e = z + N;          // [N] noise
  double* xs = e + N;         // [Nall] padded x = parameters + M noise
  const int per_problem = a.r_count * a.D;
  const int v = blockIdx.x;
  const int b = v / per_problem, rem = v - b * per_problem;
  const int r = a.r_begin + rem / a.D, d = rem % a.D;
  const size_t row_off = ((size_t(b) * a.R + r) * a.D + d) * N;
  const double* th = a.theta + (size_t(b) * a.D + d) * N;
  const bool is_new = r < a.R_gen;
  const bool philox = is_new && !a.injected;
  const double* src = nullptr;
  if (!is_new) {
    const int sidx = a.reuse_src[size_t(b) * (a.R - a.R_gen) + (r - a.R_gen)];
    src = sidx >= 0 ? a.params_prev + ((size_t(b) * a.prev_stride + sidx) * a.D + d) * N : th;
  }
  if (philox) {
    const uint64_t stream = (uint64_t(b) * uint64_t(a.rollouts_global) + uint64_t(a.rollout_id_offset + r)) * uint64_t(a.D) + d;
    const uint32_t gen_iteration = a.iteration_ptr ? *a.iteration_ptr : a.iteration;
    for (int pr = threadIdx.x; 2 * pr < N; pr += blockDim.x) {
      double z0, z1;
      normal_pair(a.seed, stream, gen_iteration, uint32_t(pr), z0, z1);
      z[2 * pr] = z0;
      if (2 * pr + 1 < N) z[2 * pr + 1] = z1;
    }
    __syncthreads();
    const double sg = gen_noise_scale(a, d);
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      double a0 = 0.0, a1 = 0.0;
      int j = t;
      for (; j + 1 < N; j += 2) {
        a0 = fma(cinv[size_t(j) * N + t], z[j], a0);
        a1 = fma(cinv[size_t(j + 1) * N + t], z[j + 1], a1);
      }
      if (j < N) a0 = fma(cinv[size_t(j) * N + t], z[j], a0);
      e[t] = sg * (a0 + a1);
    }
  } else if (is_new) {
    for (int t = threadIdx.x; t < N; t += blockDim.x) e[t] = a.eps_in[row_off + t];
  } else {
    for (int t = threadIdx.x; t < N; t += blockDim.x) e[t] = src[t] - th[t];   // policy_improvement.cpp:222
  }
  __syncthreads();
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    const double ev = e[t];
    a.noise[row_off + t] = ev;
    if (a.pre) {            // look-ahead pass: M * noise only (theta is not known yet)
      double a0 = 0.0, a1 = 0.0;
      int j = 0;
      for (; j + 1 < N; j += 2) {
        a0 = fma(ms[size_t(j) * N + t], e[j], a0);
        a1 = fma(ms[size_t(j + 1) * N + t], e[j + 1], a1);
      }
      if (j < N) a0 = fma(ms[size_t(j) * N + t], e[j], a0);
      a.noise_projected[row_off + t] = a0 + a1;
      continue;
    }
    const double pv = is_new ? th[t] + ev : src[t];
    a.params[row_off + t] = pv;
    if (!a.mode_control) continue;
    double a0 = 0.0, a1 = 0.0;
    int j = 0;
    for (; j + 1 < N; j += 2) {
      a0 = fma(ms[size_t(j) * N + t], e[j], a0);
      a1 = fma(ms[size_t(j + 1) * N + t], e[j + 1], a1);
    }
    if (j < N) a0 = fma(ms[size_t(j) * N + t], e[j], a0);
    const double y = a0 + a1;
    if (a.noise_projected) a.noise_projected[row_off + t] = y;
    xs[kPad + t] = pv + y;
  }
  if (a.pre) return;
  for (int q = threadIdx.x; q < kPad; q += blockDim.x) {
    xs[q] = a.pad_start[size_t(b) * a.D + d];
    xs[kPad + N + q] = a.pad_goal[size_t(b) * a.D + d];
  }
  if (!a.mode_control) return;
  __syncthreads();
  auto row_cost = [&](int p) -> double {     // padded row p; taps outside [0, Nall) are the dropped ones
    double w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int idx = p + j - 3;
      w[j] = (idx < 0 || idx >= Nall) ? 0.0 : xs[idx];
    }
    return stencil_row_cost(a.st, a.control_weight, w);
  };
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    double cost = row_cost(kPad + t);
    if (t == 0)
      for (int q = 0; q < kPad; ++q) cost += row_cost(q);
    if (t == N - 1)
      for (int q = 0; q < kPad; ++q) cost += row_cost(kPad + N + (kPad - 1 - q));
    a.control[row_off + t] = cost;
  }
}

// ---------------------------------------------------------------------------------------------
// k_finish_rollouts: second half of the pipelined generation.  The look-ahead pass (GenArgs::pre) left the new rollouts'
// noise and M * noise in compact buffers while the previous iteration was still running; once theta is updated this kernel
// does what is left, one thread per (problem, new rollout, dimension, timestep):
//   parameters = theta + noise;  x = parameters + M noise;  control cost = stencils over [pads, x, pads]
// — the same additions in the same order and the same stencil_row_cost as the one-pass kernels: bit-identical outputs.
// ---------------------------------------------------------------------------------------------
struct FinishArgs {
  int B, R, G, D, N;            // G new rollout slots [0, G) of R
  const double* theta;          // [B][D][N]
  const double* pre_noise;      // [B][G][D][N]
  const double* pre_y;          // [B][G][D][N]
  const double* pad_start;      // [B][D]
  const double* pad_goal;       // [B][D]
  double* noise;                // [B][R][D][N]
  double* params;               // [B][R][D][N]
  double* control;              // [B][R][D][N]
  double* noise_projected;      // optional tap
  double control_weight;
  Stencil st;
};

__global__ void __launch_bounds__(256) k_finish_rollouts(FinishArgs a) {
  const int N = a.N, Nall = N + 2 * kPad;
  const long long total = (long long)a.B * a.G * a.D * N;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long vec = idx / N;
  const int t = int(idx - vec * N);
  const int d = int(vec % a.D);
  const long long br = vec / a.D;
  const int r = int(br % a.G), b = int(br / a.G);
  const double* th = a.theta + (size_t(b) * a.D + d) * N;
  const double* en = a.pre_noise + size_t(vec) * N;
  const double* yn = a.pre_y + size_t(vec) * N;
  const size_t dst = ((size_t(b) * a.R + r) * a.D + d) * N + t;
  const double xs = a.pad_start[size_t(b) * a.D + d], xg = a.pad_goal[size_t(b) * a.D + d];
  auto xall = [&](int p) -> double {          // padded series; taps outside [0, Nall) are the dropped ones
    if (p < 0 || p >= Nall) return 0.0;
    if (p < kPad) return xs;
    if (p >= kPad + N) return xg;
    const int f = p - kPad;
    return (th[f] + en[f]) + yn[f];
  };
  auto row_cost = [&](int p) -> double {
    double w[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) w[j] = xall(p + j - 3);
    return stencil_row_cost(a.st, a.control_weight, w);
  };
  const double ev = en[t];
  a.noise[dst] = ev;
  a.params[dst] = th[t] + ev;
  if (a.noise_projected) a.noise_projected[dst] = yn[t];
  double cost = row_cost(kPad + t);
  if (t == 0)
    for (int q = 0; q < kPad; ++q) cost += row_cost(q);
  if (t == N - 1)
    for (int q = 0; q < kPad; ++q) cost += row_cost(kPad + N + (kPad - 1 - q));
  a.control[dst] = cost;
}

// ---------------------------------------------------------------------------------------------
// k_generate_mma: the same four outputs (noise, parameters, M*noise, control costs) with both linear maps evaluated as fp64
// tensor-core GEMMs (mma.sync.aligned.m8n8k4.f64, SASS DMMA) over a tile of kMmaV vectors per CTA:
//   E[t][v] = sigma_v * sum_{j >= t} A1[t][j] Z[j][v],   A1[t][j] = C^-1[j][t]          (eps = sigma C^-T z, upper triangular)
//   Y[t][v] = sum_j A2[t][j] E[j][v],                    A2[t][j] = R^-1[j][t] s_j       (y = M eps)
// M = timesteps (8-row tiles, four per warp), N = the tile's vectors (two 8-column tiles), K = j in steps of 4.
// The matrices are stored zero-padded as Mp[pass][j][t] (Np = N rounded up to 8 rows of `strideA` doubles per 128-timestep pass)
// and streamed through shared memory in chunks of kMmaChunk rows by TMA bulk copies (one cp.async.bulk per chunk, two buffers,
// one mbarrier each): the copy of chunk
// c + 1 is in flight while the warps issue the DMMAs of chunk c, and no thread computes an address or a predicate for an A
// operand (the first version loaded fragments with __ldg: 24 instructions per load, 21 % of the kernel, 230 us).  Staged rows
// have a stride of 4 (mod 16) doubles and the vector tiles a stride of 20, which makes every fragment load conflict free.
// Z is filled from the same Philox streams as the other generation kernels, so the three agree to rounding
// (test_dense_generation_kernels_match_the_band_solves).
// ---------------------------------------------------------------------------------------------
constexpr int kMmaV = 16;        // vectors per CTA
constexpr int kMmaS = 20;        // row stride of the vector tiles in doubles (= 4 mod 16: conflict-free fragment loads)
#ifndef STOMP_MMA_CHUNK
#define STOMP_MMA_CHUNK 8
#endif
constexpr int kMmaChunk = STOMP_MMA_CHUNK;     // matrix rows (j) per staged chunk (8 = two k-steps)
constexpr int kMmaGroup = 128;   // timesteps per pass (16 m-tiles: four per warp)

__host__ __device__ inline int mma_a_stride(int cols) {   // >= cols, = 4 (mod 16)
  int st = cols;
  while ((st & 15) != 4) ++st;
  return st;
}
__host__ __device__ inline size_t mma_smem_bytes(int N) {
  const int Np = (N + 7) & ~7, rows = Np > N + 2 * kPad ? Np : N + 2 * kPad;
  return (size_t(2) * rows * kMmaS + size_t(2) * kMmaChunk * mma_a_stride(Np < kMmaGroup ? Np : kMmaGroup)) * 8 + 16;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// acc[i][nt][:] = sum_j Mp[j][t] * Bs[j][v] for the m-tiles mt = tc0 / 8 + warp + 4 i of the pass starting at timestep tc0.
// kTriangular: Mp[j][t] = 0 for j < t, so chunks below the pass and k-steps below an m-tile are skipped, and only the columns
// t <= j of a row are copied.  All threads of the CTA call this together; `phase` = the parities of the two mbarriers.
template <bool kTriangular>
__device__ __forceinline__ void mma_gemm(const double* __restrict__ Mp, const double* __restrict__ Bs, double* __restrict__ As,
                                         unsigned bar0, unsigned (&phase)[2], int Np, int tc0, int tcols, double (&acc)[4][2][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, frow = lane >> 2, fk = lane & 3;
  const int strideA = mma_a_stride(Np < kMmaGroup ? Np : kMmaGroup);
  const int mtiles_here = tcols >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
  const int nchunks = Np / kMmaChunk;
  const int c_begin = kTriangular ? tc0 / kMmaChunk : 0;
  // the matrix is stored pass by pass with the staged row stride ([pass][Np][strideA], engine.cu), so a chunk of rows is ONE
  // contiguous bulk copy (per-row copies of 832 bytes made the TMA request rate the bottleneck: 0.33 ms instead of 0.16)
  const double* Mpass = Mp + size_t(tc0 / kMmaGroup) * Np * strideA;
  auto issue = [&](int c, int buf) {     // one thread: arm the barrier, one bulk copy for the chunk's rows
    const unsigned bytes = unsigned(kMmaChunk * strideA * 8);
    mbar_expect_tx(bar0 + 8u * buf, bytes);
    bulk_g2s(smem_u32(As + size_t(buf) * kMmaChunk * strideA), Mpass + size_t(c) * kMmaChunk * strideA, bytes, bar0 + 8u * buf);
  };
  if (threadIdx.x == 0) issue(c_begin, 0);
  for (int c = c_begin; c < nchunks; ++c) {
    const int buf = (c - c_begin) & 1;
    if (threadIdx.x == 0 && c + 1 < nchunks) issue(c + 1, buf ^ 1);   // the other buffer was released by the barrier below
    mbar_wait(bar0 + 8u * buf, phase[buf]);
    phase[buf] ^= 1u;
    const double* Ab = As + size_t(buf) * kMmaChunk * strideA + fk * strideA + frow;
    // all fragment loads of the chunk first (independent LDS: one shared-memory latency per chunk), then the DMMAs; an m-tile
    // that does not exist or lies above the diagonal (warp-uniform conditions) is skipped
    double bf[kMmaChunk / 4][2], af[kMmaChunk / 4][4];
#pragma unroll
    for (int ks = 0; ks < kMmaChunk / 4; ++ks) {
      const int j = c * kMmaChunk + ks * 4 + fk;
      bf[ks][0] = Bs[j * kMmaS + frow];
      bf[ks][1] = Bs[j * kMmaS + 8 + frow];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[ks][i] = Ab[ks * 4 * strideA + min(warp + 4 * i, mtiles_here - 1) * 8];
    }
#pragma unroll
    for (int ks = 0; ks < kMmaChunk / 4; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ml = warp + 4 * i;                                  // m-tile within the pass
        if (ml < mtiles_here && (!kTriangular || c * (kMmaChunk / 4) + ks >= 2 * (tc0 / 8 + ml))) {
          dmma(acc[i][0][0], acc[i][0][1], af[ks][i], bf[ks][0]);
          dmma(acc[i][1][0], acc[i][1][1], af[ks][i], bf[ks][1]);
        }
      }
    __syncthreads();      // everyone is done with `buf` before it is refilled two chunks later
  }
}

__global__ void __launch_bounds__(128) k_generate_mma(GenArgs a, const double* __restrict__ a1p, const double* __restrict__ a2p) {
  extern __shared__ __align__(16) double msm[];
  const int N = a.N, Np = (N + 7) & ~7, Nall = N + 2 * kPad;
  const int rows = Np > Nall ? Np : Nall;
  double* T0 = msm;                              // [rows][kMmaS]  standard normals Z, later the padded x = parameters + M noise
  double* T1 = T0 + size_t(rows) * kMmaS;        // [rows][kMmaS]  noise E, later the control costs
  double* As = T1 + size_t(rows) * kMmaS;        // [2][kMmaChunk][strideA] staged matrix rows
  const unsigned bar0 = smem_u32(As + size_t(2) * kMmaChunk * mma_a_stride(Np < kMmaGroup ? Np : kMmaGroup));
  __shared__ size_t s_row[kMmaV], s_th[kMmaV];
  __shared__ const double* s_src[kMmaV];
  __shared__ double s_sg[kMmaV], s_ps[kMmaV], s_pg[kMmaV];
  __shared__ int s_flags[kMmaV];                 // bit 0 valid, bit 1 new rollout, bit 2 Philox noise
  __shared__ uint64_t s_stream[kMmaV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, frow = lane >> 2, fk = lane & 3;
  const int per_problem = a.r_count * a.D;
  const long long nvec = (long long)a.B * per_problem;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8u, 1); }
  if (threadIdx.x < kMmaV) {
    const int vl = threadIdx.x;
    const long long v = (long long)blockIdx.x * kMmaV + vl;
    int flags = 0;
    if (v < nvec) {
      const int b = int(v / per_problem), rem = int(v - (long long)b * per_problem);
      const int r = a.r_begin + rem / a.D, d = rem % a.D;
      const bool is_new = r < a.R_gen;
      flags = 1 | (is_new ? 2 : 0) | ((is_new && !a.injected) ? 4 : 0);
      s_row[vl] = ((size_t(b) * a.R + r) * a.D + d) * N;
      s_th[vl] = (size_t(b) * a.D + d) * N;
      const double* src = nullptr;
      if (!is_new) {
        const int sidx = a.reuse_src[size_t(b) * (a.R - a.R_gen) + (r - a.R_gen)];
        src = sidx >= 0 ? a.params_prev + ((size_t(b) * a.prev_stride + sidx) * a.D + d) * N : a.theta + (size_t(b) * a.D + d) * N;
      }
      s_src[vl] = src;
      s_sg[vl] = gen_noise_scale(a, d);
      s_ps[vl] = a.pad_start[size_t(b) * a.D + d];
      s_pg[vl] = a.pad_goal[size_t(b) * a.D + d];
      s_stream[vl] = (uint64_t(b) * uint64_t(a.rollouts_global) + uint64_t(a.rollout_id_offset + r)) * uint64_t(a.D) + d;
    }
    s_flags[vl] = flags;
  }
  __syncthreads();
  int any_philox = 0;
#pragma unroll
  for (int vl = 0; vl < kMmaV; ++vl) any_philox |= s_flags[vl] & 4;
  unsigned phase[2] = {0u, 0u};
  // ---- Z: standard normals of the Philox vectors (zeros elsewhere and in the padded rows) -------------------------------
  if (any_philox) {
    const uint32_t gen_iteration = a.iteration_ptr ? *a.iteration_ptr : a.iteration;
    const int npairs = (Np + 1) >> 1;
    for (int idx = threadIdx.x; idx < kMmaV * npairs; idx += blockDim.x) {
      const int vl = idx % kMmaV, pr = idx / kMmaV;
      double z0 = 0.0, z1 = 0.0;
      if ((s_flags[vl] & 4) && 2 * pr < N) {
        normal_pair(a.seed, s_stream[vl], gen_iteration, uint32_t(pr), z0, z1);
        if (2 * pr + 1 >= N) z1 = 0.0;
      }
      T0[(2 * pr) * kMmaS + vl] = z0;
      if (2 * pr + 1 < Np) T0[(2 * pr + 1) * kMmaS + vl] = z1;
    }
    __syncthreads();
  }
  double acc[4][2][2];
  // ---- GEMM 1 + epilogue: noise and parameters, one pass per 128 timesteps ------------------------------------------------------
  for (int tc0 = 0; tc0 < Np; tc0 += kMmaGroup) {
    const int tcols = min(kMmaGroup, Np - tc0);
    if (any_philox) {
      mma_gemm<true>(a1p, T0, As, bar0, phase, Np, tc0, tcols, acc);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[i][nt][0] = acc[i][nt][1] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ml = warp + 4 * i;
      if (ml * 8 >= tcols) continue;
      const int t = tc0 + ml * 8 + frow;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int vl = nt * 8 + 2 * fk + h;
          T1[t * kMmaS + vl] = (s_flags[vl] & 4) ? s_sg[vl] * acc[i][nt][h] : 0.0;     // eps of a Philox vector
        }
    }
  }
  __syncthreads();          // Philox eps complete; Z is dead: T0 becomes the padded x
  // ---- noise and parameters out, coalesced along time (32 consecutive t of one vector per warp); parameters into T0 ----------
  for (int idx = threadIdx.x; idx < kMmaV * Np; idx += blockDim.x) {
    const int vl = idx / Np, t = idx - vl * Np;
    const int fl = s_flags[vl];
    if (!(fl & 1) || t >= N) continue;
    const double th = a.theta[s_th[vl] + t];
    double e, pv;
    if (fl & 4) { e = T1[t * kMmaS + vl]; pv = th + e; }
    else if (fl & 2) { e = a.eps_in[s_row[vl] + t]; pv = th + e; }
    else { pv = s_src[vl][t]; e = pv - th; }                     // policy_improvement.cpp:222
    a.noise[s_row[vl] + t] = e;
    a.params[s_row[vl] + t] = pv;
    T1[t * kMmaS + vl] = e;
    T0[(kPad + t) * kMmaS + vl] = pv;
  }
  if (!a.mode_control) return;
  __syncthreads();          // E and the parameters complete
  // ---- GEMM 2 + epilogue: y = M eps, x = parameters + y ------------------------------------------------------------------
  for (int tc0 = 0; tc0 < Np; tc0 += kMmaGroup) {
    const int tcols = min(kMmaGroup, Np - tc0);
    mma_gemm<false>(a2p, T1, As, bar0, phase, Np, tc0, tcols, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ml = warp + 4 * i;
      if (ml * 8 >= tcols) continue;
      const int t = tc0 + ml * 8 + frow;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int vl = nt * 8 + 2 * fk + h;
          const int fl = s_flags[vl];
          if ((fl & 1) && t < N) {
            const double y = acc[i][nt][h];
            if (a.noise_projected) a.noise_projected[s_row[vl] + t] = y;
            T0[(kPad + t) * kMmaS + vl] += y;          // x = parameters + M eps
          }
        }
    }
  }
  for (int idx = threadIdx.x; idx < kMmaV * kPad; idx += blockDim.x) {
    const int vl = idx % kMmaV, q = idx / kMmaV;
    T0[q * kMmaS + vl] = s_ps[vl];
    T0[(kPad + N + q) * kMmaS + vl] = s_pg[vl];
  }
  __syncthreads();          // x complete; E is dead: T1 becomes the control costs
  // ---- control-cost stencils over [pads, x, pads] (covariant_trajectory_policy.cpp:228-255) -----------------------------------
  const double* Xs = T0;
  double* Cs = T1;
  auto row_cost = [&](int vl, int p, bool interior) -> double {     // padded row p; taps outside [0, Nall) are the dropped ones
    double cost = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (a.st.weight[k] == 0.0) continue;
      double s_ = 0.0;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int idx = p + j - 3;
        if (!interior && (idx < 0 || idx >= Nall)) continue;
        s_ += a.st.coef[k][j] * Xs[idx * kMmaS + vl];
      }
      cost += a.control_weight * a.st.weight[k] * (s_ * s_);
    }
    return cost;
  };
  for (int idx = threadIdx.x; idx < kMmaV * N; idx += blockDim.x) {
    const int vl = idx % kMmaV, t = idx / kMmaV;
    if (!(s_flags[vl] & 1)) continue;
    double cost = row_cost(vl, kPad + t, true);
    if (t == 0)
      for (int q = 0; q < kPad; ++q) cost += row_cost(vl, q, false);
    if (t == N - 1)
      for (int q = 0; q < kPad; ++q) cost += row_cost(vl, kPad + N + (kPad - 1 - q), false);
    Cs[t * kMmaS + vl] = cost;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kMmaV * N; idx += blockDim.x) {
    const int vl = idx / N, t = idx - vl * N;
    if (s_flags[vl] & 1) a.control[s_row[vl] + t] = Cs[t * kMmaS + vl];
  }
}

// ---------------------------------------------------------------------------------------------
// k_cost: the cost plugin.  One CTA per rollout; each warp owns a tile of 29 free timesteps
// (lane l <-> timestep tile*29 - 1 + l), so the finite-difference velocity taps (-1, 0, +1, +2) of the
// productive lanes 1..29 come from neighbouring lanes by warp shuffle and no sphere position ever leaves
// the register file.
// ---------------------------------------------------------------------------------------------
template <typename Real>
struct CostArgs {
  int n_rollouts;            // rollouts per problem processed by this launch
  int total_rollouts;        // problems * n_rollouts
  int pack;                  // rollouts whose timelines one CTA concatenates (see k_cost)
  int tiles_per_job;         // ceil((pack * (N + 3) - 3) / 29)
  int D, N, K, num_nodes;
  int include_pads;          // 1: start/goal padding points count towards collision_free (iteration_ == 0)
  size_t params_problem_stride, params_rollout_stride;  // in doubles
  size_t cost_problem_stride;                           // in doubles
  int flag_problem_stride, flag_offset;
  const double* params;      // rollout parameters
  const double* pad_start;   // [B][D]
  const double* pad_goal;    // [B][D]
  const double* qinv_t;      // [N][N] transposed scaled quad_cost_inv_ (row fv = column fv)
  const int* has_limits;     // [D]
  const double* limit_min;   // [D]
  const double* limit_max;   // [D]
  const DevNode<Real>* nodes;
  const DevSphere<Real>* spheres;
  const Real* sqrt_table;    // [256] distance of a squared cell distance (u8 grids)
  Sdf sdf;
  double inv_time;           // 1/discretization
  double obstacle_weight;
  double constraint_weight;  // constraint_cost_weight
  int num_constraints;
  const DevConstraint<Real>* constraints;
  int* constraints_satisfied; // same indexing as collision_free (may be NULL)
  double* costs;             // [..][N]
  int* collision_free;       // [..]
  double* clipped;           // optional tap [same layout as params]
  stomp_sphere_debug* debug; // optional tap [N+3][K] (rollout 0 of problem 0)
  int params_16B;            // rollout rows are 16-byte aligned and D*N is even: staged with 16-byte copies
  const DevCluster<Real>* clusters;   // sphere clusters (always a partition of the sphere table)
  int num_clusters;
  CullField cull;
  // lookup constants in the kernel's arithmetic type, computed once on the host: they are read as constant-bank operands
  // instead of occupying ~20 registers per thread (grid origin, resolution, -origin/res, finite-difference rule / dt)
  Real g_ox, g_oy, g_oz, g_res, g_inv_res, g_nox, g_noy, g_noz;
  Real c_m1, c_0, c_p1, c_p2;
  unsigned lim_x, lim_y, lim_z;     // n - 2 per axis
};

template <typename Real> struct Math;
template <> struct Math<double> {
  static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
  static __device__ __forceinline__ double round_(double x) { return round(x); }
  static __device__ __forceinline__ double fabs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double asin_(double x) { return asin(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __device__ __forceinline__ double cos_(double x) { return cos(x); }
  static __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
};
template <> struct Math<float> {
  static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float round_(float x) { return roundf(x); }
  static __device__ __forceinline__ float fabs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float asin_(float x) { return asinf(x); }
  static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
  static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
  static __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
};

// cost of one orientation constraint for the segment rotation Rs = Fr * rel (Fr: rotation of the carrying node's frame, or
// nullptr for a static segment); bullet's btMatrix3x3::getEulerYPR (solution 1) restated.  Returns "satisfied".
template <typename Real>
__device__ __forceinline__ bool constraint_cost(const DevConstraint<Real>& c, const Real* Fr, Real& cost) {
  Real Rs[9], res[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Rs[i * 3 + j] = Fr ? Fr[i * 3] * c.rel[j] + Fr[i * 3 + 1] * c.rel[3 + j] + Fr[i * 3 + 2] * c.rel[6 + j] : c.rel[i * 3 + j];
  const Real* a = c.body_fixed ? c.nominal_inv : Rs;
  const Real* b = c.body_fixed ? Rs : c.nominal_inv;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) res[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
  const Real kHalfPi = Real(1.5707963267948966);
  Real roll, pitch, yaw;
  if (Math<Real>::fabs_(res[6]) >= Real(1)) {
    yaw = Real(0);
    roll = Math<Real>::atan2_(res[7], res[8]);
    pitch = res[6] < Real(0) ? kHalfPi : -kHalfPi;
  } else {
    pitch = -Math<Real>::asin_(res[6]);
    const Real cp = Math<Real>::cos_(pitch);
    roll = Math<Real>::atan2_(res[7] / cp, res[8] / cp);
    yaw = Math<Real>::atan2_(res[3] / cp, res[0] / cp);
  }
  roll = Math<Real>::fabs_(roll); pitch = Math<Real>::fabs_(pitch); yaw = Math<Real>::fabs_(yaw);
  cost = c.weight * (c.w[0] * roll + c.w[1] * pitch + c.w[2] * yaw);
  return !(roll > c.tol[0] || pitch > c.tol[1] || yaw > c.tol[2]);
}

__device__ __forceinline__ double shfl_rel(double v, int delta) {
  // value of lane (lane + delta), clamped at the warp edges (edge lanes are halo lanes, never productive)
  return delta < 0 ? __shfl_up_sync(0xffffffffu, v, unsigned(-delta)) : __shfl_down_sync(0xffffffffu, v, unsigned(delta));
}
__device__ __forceinline__ float shfl_rel(float v, int delta) {
  return delta < 0 ? __shfl_up_sync(0xffffffffu, v, unsigned(-delta)) : __shfl_down_sync(0xffffffffu, v, unsigned(delta));
}

// Explicit 32-bit shared-memory addressing for the per-sphere hot loop: the tables live at run-time offsets of the
// dynamic shared segment, and with generic pointers ptxas re-derived the shared window base (S2UR/ULEA) at almost every
// access (14 % of k_cost's issued instructions, ncu source page); a 32-bit .shared address needs none of that.
__device__ __forceinline__ void lds2(unsigned addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ double lds1(unsigned addr, double) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds1(unsigned addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts1(unsigned addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void sts1(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// one collision sphere from the shared table (layout of DevSphere<Real>)
struct SphereRegsD { double s0, s1, s2, radius, clearance, inv_clearance; };
struct SphereRegsF { float s0, s1, s2, radius, clearance, inv_clearance; };
__device__ __forceinline__ void load_sphere(unsigned addr, SphereRegsD& r) {
  lds2(addr, r.s0, r.s1);
  lds2(addr + 16, r.s2, r.radius);
  lds2(addr + 32, r.clearance, r.inv_clearance);
}
__device__ __forceinline__ void load_sphere(unsigned addr, SphereRegsF& r) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.s0), "=f"(r.s1), "=f"(r.s2), "=f"(r.radius) : "r"(addr));
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.clearance), "=f"(r.inv_clearance) : "r"(addr + 16));
}
template <typename Real> struct SphereRegsOf;
template <> struct SphereRegsOf<double> { typedef SphereRegsD type; };
template <> struct SphereRegsOf<float> { typedef SphereRegsF type; };

// distance_field VoxelGrid::getCellFromLocation: int(round((loc - origin) / resolution)), and the 1-cell margin
// test of getDistanceGradient.  Fast path (fp64): t = pos/res - origin/res as one DFMA, then t + 1.5*2^36 puts
// t as Q16.16 fixed point into the low mantissa word: cell = (lo + 0x8000) >> 16 and the distance to the
// rounding boundary are integer-pipe work, and the high word tells whether 0 <= t < 65536.  Whenever any of the
// three coordinates is within 2^-15 cell of a rounding boundary (or out of that range) the exact division +
// round-half-away form decides (one rarely taken branch per sphere), so the indices are identical to the
// reference form for every input.
struct GridD {   // per-kernel constants of the lookup
  double ox, oy, oz, res, inv_res, nox, noy, noz;   // no* = -origin/res
  int nx1, ny1, nz1, sny, snz, nby, nbz;
  unsigned lx, ly, lz;                              // n - 2: a cell is interior when unsigned(c - 1) < l
};
struct GridF {
  float ox, oy, oz, res, inv_res, nox, noy, noz;
  int nx1, ny1, nz1, sny, snz, nby, nbz;
  unsigned lx, ly, lz;
};
template <typename Real> struct GridOf;
template <> struct GridOf<double> { typedef GridD type; };
template <> struct GridOf<float> { typedef GridF type; };

__device__ __forceinline__ int exact_cell(double pos, double origin, double res) {
  const double tq = (pos - origin) / res;
  return fabs(tq) < 2.0e9 ? int(round(tq)) : -1;
}
__device__ __forceinline__ int exact_cell(float pos, float origin, float res) {
  const float tq = (pos - origin) / res;
  return fabsf(tq) < 2.0e9f ? int(roundf(tq)) : -1;
}

// returns true when (cx,cy,cz) is inside the grid with the 1-cell margin (1 <= cell <= n-2 on every axis)
__device__ __forceinline__ bool voxel_cells(const GridD& g, double px, double py, double pz, int& cx, int& cy, int& cz) {
  const double magic = 103079215104.0;                    // 1.5 * 2^36
  const double mx = fma(px, g.inv_res, g.nox) + magic;
  const double my = fma(py, g.inv_res, g.noy) + magic;
  const double mz = fma(pz, g.inv_res, g.noz) + magic;
  const unsigned lx = unsigned(__double2loint(mx)), ly = unsigned(__double2loint(my)), lz = unsigned(__double2loint(mz));
  const unsigned fx = lx + 0x8000u, fy = ly + 0x8000u, fz = lz + 0x8000u;
  cx = int(fx >> 16); cy = int(fy >> 16); cz = int(fz >> 16);
  // near a rounding boundary (fraction within 2 units of 0 or 65536), wrapped, or outside [0, 65536)
  const unsigned bad = unsigned(((fx & 0xffffu) - 2u) > 0xfffbu) | unsigned(((fy & 0xffffu) - 2u) > 0xfffbu) |
                       unsigned(((fz & 0xffffu) - 2u) > 0xfffbu) | unsigned(fx < lx) | unsigned(fy < ly) | unsigned(fz < lz) |
                       unsigned(__double2hiint(mx) != 0x42380000) | unsigned(__double2hiint(my) != 0x42380000) |
                       unsigned(__double2hiint(mz) != 0x42380000);
  if (bad) {
    cx = exact_cell(px, g.ox, g.res);
    cy = exact_cell(py, g.oy, g.res);
    cz = exact_cell(pz, g.oz, g.res);
  }
  return (unsigned(cx - 1) < g.lx) & (unsigned(cy - 1) < g.ly) & (unsigned(cz - 1) < g.lz);
}
__device__ __forceinline__ bool voxel_cells(const GridF& g, float px, float py, float pz, int& cx, int& cy, int& cz) {
  const float tx = fmaf(px, g.inv_res, g.nox), ty = fmaf(py, g.inv_res, g.noy), tz = fmaf(pz, g.inv_res, g.noz);
  cx = __float2int_rn(tx); cy = __float2int_rn(ty); cz = __float2int_rn(tz);
  const bool bad = (fabsf(tx - float(cx)) > 0.5f - 1e-3f) | (fabsf(ty - float(cy)) > 0.5f - 1e-3f) |
                   (fabsf(tz - float(cz)) > 0.5f - 1e-3f) | !(fabsf(tx) < 1.0e9f) | !(fabsf(ty) < 1.0e9f) | !(fabsf(tz) < 1.0e9f);
  if (bad) {
    cx = exact_cell(px, g.ox, g.res);
    cy = exact_cell(py, g.oy, g.res);
    cz = exact_cell(pz, g.oz, g.res);
  }
  return (unsigned(cx - 1) < g.lx) & (unsigned(cy - 1) < g.ly) & (unsigned(cz - 1) < g.lz);
}

// distance of one voxel: PropagationDistanceField::getDistance = sqrt_table[d^2] (u8 / u16 grids) or metres (f32)
template <typename Real, int kVox>
__device__ __forceinline__ Real voxel_distance(const void* vox, int idx, unsigned sqrt_tab_addr, Real res) {
  if (kVox == STOMP_VOXEL_U8_SQ) {
#if STOMP_COST_VOX_EVICT_LAST
    unsigned v;
    asm volatile("ld.global.nc.L1::evict_last.u8 %0, [%1];" : "=r"(v) : "l"(static_cast<const uint8_t*>(vox) + idx));
    return lds1(sqrt_tab_addr + v * unsigned(sizeof(Real)), Real(0));
#else
    return lds1(sqrt_tab_addr + unsigned(__ldg(static_cast<const uint8_t*>(vox) + idx)) * unsigned(sizeof(Real)), Real(0));
#endif
  }
  if (kVox == STOMP_VOXEL_U16_SQ) return Math<Real>::sqrt_(Real(__ldg(static_cast<const uint16_t*>(vox) + idx))) * res;
  return Real(__ldg(static_cast<const float*>(vox) + idx));
}

// STOMP_SDF_TRILINEAR (engine extension; the reference's lookup is the nearest-cell one above): trilinear interpolation of the
// eight cell distances around the point.  Every corner follows the nearest-cell rule (0 on / outside the outermost layer of the
// grid), so this is the continuous extension of the reference's piecewise-constant field.  The two corners that differ only in
// z are adjacent bytes / words, so a lane's eight gathers touch at most four sectors.  (cx, cy, cz) = the lower corner.
template <typename Real, int kVox, typename Grid>
__device__ __forceinline__ Real corner_distance(const Grid& g, const void* vox, int cx, int cy, int cz, unsigned sqrt_tab_addr) {
  const bool in = (unsigned(cx - 1) < unsigned(g.nx1 - 1)) & (unsigned(cy - 1) < unsigned(g.ny1 - 1)) & (unsigned(cz - 1) < unsigned(g.nz1 - 1));
  return in ? voxel_distance<Real, kVox>(vox, (cx * g.sny + cy) * g.snz + cz, sqrt_tab_addr, g.res) : Real(0);
}
template <typename Real, int kVox, typename Grid>
__device__ __forceinline__ Real trilinear_distance(const Grid& g, const void* vox, Real px, Real py, Real pz, unsigned sqrt_tab_addr,
                                                   int& cx, int& cy, int& cz) {
  const Real tx = (px - g.ox) / g.res, ty = (py - g.oy) / g.res, tz = (pz - g.oz) / g.res;
  const Real flx = floor(tx), fly = floor(ty), flz = floor(tz);
  cx = Math<Real>::fabs_(flx) < Real(2.0e9) ? int(flx) : -1;
  cy = Math<Real>::fabs_(fly) < Real(2.0e9) ? int(fly) : -1;
  cz = Math<Real>::fabs_(flz) < Real(2.0e9) ? int(flz) : -1;
  const Real fx = tx - flx, fy = ty - fly, fz = tz - flz;
  Real acc = Real(0);
#pragma unroll
  for (int dx = 0; dx < 2; ++dx)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dz = 0; dz < 2; ++dz) {
        const Real w = (dx ? fx : Real(1) - fx) * (dy ? fy : Real(1) - fy) * (dz ? fz : Real(1) - fz);
        acc += w * corner_distance<Real, kVox>(g, vox, cx + dx, cy + dy, cz + dz, sqrt_tab_addr);
      }
  return acc;
}

// Lane packing: a warp tile is a window of 32 consecutive points of the CTA's *concatenated* timeline
//   [rollout 0: t = -1 .. N+1][rollout 1: t = -1 .. N+1] ...          (N + 3 points per rollout, `pack` rollouts)
// advanced by 29 points per tile; lanes 1..29 of a window are productive when their point is a free timestep, and their
// velocity taps (-1, +1, +2) are always points of the same rollout inside the same window.  With pack = 1 this is the
// plain "29 timesteps per warp" tiling; packing 2 rollouts of N = 100 needs 7 warps instead of 8.
// One CTA per job (the hardware block scheduler balances rollouts whose joint-limit projection takes longer); the job loop
// only matters when a caller launches fewer CTAs than jobs.
#ifndef STOMP_COST_MIN_BLOCKS
#define STOMP_COST_MIN_BLOCKS 4
#endif
constexpr int kCostMaxThreads = 224;   // 7 warps; 72 registers -> 4 CTAs (28 warps) per SM

// kSplit (small batches: a handful of rollouts cannot fill the machine, so k_cost's time is ONE warp's latency — C1: 36 us for
// five rollouts): kSplitGroups warps share a 29-timestep tile.  Each runs the whole FK (redundant, but off nobody's critical
// path) and evaluates every kSplitGroups-th sphere cluster, parking each sphere's weighted contribution in shared memory;
// after a barrier one warp per tile adds the K contributions in sphere order — the same additions in the same order as the
// sequential loop (a sphere that contributed nothing adds +0), so the costs are bit-identical to the unsplit kernel.
constexpr int kSplitGroups = 4;
constexpr int kSplitMaxThreads = 512;

template <typename Real, bool kDebug, int kVox, bool kCons, bool kTri = false, bool kCull = false, bool kSplit = false>
__global__ void __launch_bounds__(kSplit ? kSplitMaxThreads : kCostMaxThreads, kSplit ? 1 : STOMP_COST_MIN_BLOCKS) k_cost(CostArgs<Real> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int D = a.D, N = a.N, K = a.K, P = a.pack;
  constexpr int S = kSplit ? kSplitGroups : 1;
  const int DN = D * N;
  // shared layout: rollouts [P][D][N] | start/goal padding [P][2][D] | nodes | spheres | sqrt table | constraints | clusters |
  // mbarrier
  double* q = reinterpret_cast<double*>(smem_raw);                        // joint-limit-projected trajectories
  double* pads = q + size_t(P) * DN;
  DevNode<Real>* nodes = reinterpret_cast<DevNode<Real>*>(smem_raw + ((size_t(P) * (DN + 2 * D) * 8 + 15) & ~size_t(15)));
  DevSphere<Real>* spheres = reinterpret_cast<DevSphere<Real>*>(nodes + a.num_nodes);
  Real* sqrt_tab = reinterpret_cast<Real*>(spheres + K);                  // [256]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned sph_addr = smem_u32(spheres), tab_addr = smem_u32(sqrt_tab);
  DevConstraint<Real>* cons = reinterpret_cast<DevConstraint<Real>*>(sqrt_tab + 256);
  const int num_cons = kCons ? a.num_constraints : 0;
  DevCluster<Real>* clusters = reinterpret_cast<DevCluster<Real>*>(cons + num_cons);
  const unsigned bar = smem_u32(clusters + a.num_clusters);
  Real* contrib = reinterpret_cast<Real*>(reinterpret_cast<unsigned char*>(clusters + a.num_clusters) + 16);   // kSplit: [tiles][K][32]
  const unsigned table_bytes = unsigned(sizeof(DevNode<Real>)) * a.num_nodes + unsigned(sizeof(DevSphere<Real>)) * K +
                               256u * unsigned(sizeof(Real)) + unsigned(sizeof(DevConstraint<Real>)) * num_cons +
                               unsigned(sizeof(DevCluster<Real>)) * a.num_clusters;
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  typename GridOf<Real>::type g;
  g.ox = a.g_ox; g.oy = a.g_oy; g.oz = a.g_oz; g.res = a.g_res; g.inv_res = a.g_inv_res;
  g.nox = a.g_nox; g.noy = a.g_noy; g.noz = a.g_noz;
  g.nx1 = a.sdf.nx - 1; g.ny1 = a.sdf.ny - 1; g.nz1 = a.sdf.nz - 1; g.sny = a.sdf.ny; g.snz = a.sdf.nz;
  g.nby = a.sdf.nby; g.nbz = a.sdf.nbz;
  g.lx = a.lim_x; g.ly = a.lim_y; g.lz = a.lim_z;
  const Real c_m1 = a.c_m1, c_0 = a.c_0, c_p1 = a.c_p1, c_p2 = a.c_p2;
  const int ntiles = a.tiles_per_job;
  const int seg = N + 3;                       // timeline points per rollout: t = -1 .. N+1
  constexpr bool bricks = STOMP_SDF_BRICKS && !kTri;      // the engine always builds the bricked copy
  const void* vox = bricks ? a.sdf.brick : a.sdf.vox;
  const int njobs = (a.total_rollouts + P - 1) / P;

  unsigned phase = 0;
  for (int job = blockIdx.x; job < njobs; job += gridDim.x, phase ^= 1u) {
    const int first = job * P, count = min(P, a.total_rollouts - first);
    if (job != int(blockIdx.x)) __syncthreads();   // previous job fully consumed
    if (threadIdx.x == 0) {
      // robot tables (first job only) and the rollout rows of this job: one bulk copy each, completion on the mbarrier
      const bool tables = job == int(blockIdx.x);
      const unsigned row_bytes = a.params_16B ? unsigned(DN) * 8u : 0u;
      mbar_expect_tx(bar, (tables ? table_bytes : 0u) + row_bytes * unsigned(count));
      if (tables) {
        bulk_g2s(smem_u32(nodes), a.nodes, unsigned(sizeof(DevNode<Real>)) * a.num_nodes, bar);
        if (K > 0) bulk_g2s(sph_addr, a.spheres, unsigned(sizeof(DevSphere<Real>)) * K, bar);
        bulk_g2s(tab_addr, a.sqrt_table, 256u * unsigned(sizeof(Real)), bar);
        if (num_cons > 0) bulk_g2s(smem_u32(cons), a.constraints, unsigned(sizeof(DevConstraint<Real>)) * num_cons, bar);
        if (a.num_clusters > 0) bulk_g2s(smem_u32(clusters), a.clusters, unsigned(sizeof(DevCluster<Real>)) * a.num_clusters, bar);
      }
      if (a.params_16B)
        for (int p = 0; p < count; ++p) {
          const int ro = first + p;
          const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
          bulk_g2s(smem_u32(q + size_t(p) * DN), a.params + size_t(b) * a.params_problem_stride + size_t(r) * a.params_rollout_stride,
                   row_bytes, bar);
        }
    }
    if (!a.params_16B)   // rows that are not 16-byte aligned (odd D*N): plain loads
      for (int i = threadIdx.x; i < count * DN; i += blockDim.x) {
        const int p = i / DN, k = i - p * DN, ro = first + p;
        const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
        q[i] = a.params[size_t(b) * a.params_problem_stride + size_t(r) * a.params_rollout_stride + k];
      }
    if (int(threadIdx.x) < count * 2 * D) {   // start / goal padding values of the job's problems
      const int p = threadIdx.x / (2 * D), k = threadIdx.x - p * 2 * D;
      const int b = (first + p) / a.n_rollouts;
      pads[threadIdx.x] = k < D ? a.pad_start[size_t(b) * D + k] : a.pad_goal[size_t(b) * D + k - D];
    }
    if (int(threadIdx.x) < count) {   // flags start at "free / satisfied"; ordered before the "= 0" stores by the barriers below
      const int ro = first + threadIdx.x;
      const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
      if (a.collision_free) a.collision_free[size_t(b) * a.flag_problem_stride + a.flag_offset + r] = 1;
      if (a.constraints_satisfied) a.constraints_satisfied[size_t(b) * a.flag_problem_stride + a.flag_offset + r] = 1;
    }
    mbar_wait(bar, phase);
    __syncthreads();

    // ---- handleJointLimits: warp per joint column, <= 11 passes of (arg max violation, rank-1 correction) ----------
    for (int col = warp; col < count * D; col += nwarps) {
      const int d = col % D;
      if (!a.has_limits[d]) continue;
      const double jmax = a.limit_max[d], jmin = a.limit_min[d];
      double* qd = q + size_t(col) * N;
      {
        // most columns violate nothing (C2: 80 %): find that out with compares only, before the arg-max machinery
        bool any = false;
        for (int i = lane; i < N; i += 32) {
          const double v = qd[i];
          any |= (v > jmax && jmax - v < -1e-6) || (v < jmin && jmin - v > 1e-6);
        }
        if (!__any_sync(0xffffffffu, any)) continue;
      }
      for (int pass = 0; pass < 11; ++pass) {
        double best_abs = 1e-6, best_amount = 0.0;
        int best_idx = -1;
        for (int i = lane; i < N; i += 32) {
          double v = qd[i], amount = 0.0;
          if (v > jmax) amount = jmax - v;
          else if (v < jmin) amount = jmin - v;
          double aa = fabs(amount);
          if (aa > best_abs) { best_abs = aa; best_amount = amount; best_idx = i; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          double oa = __shfl_xor_sync(0xffffffffu, best_abs, off);
          double om = __shfl_xor_sync(0xffffffffu, best_amount, off);
          int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
          // strict '>' in index order keeps the first maximum: larger violation wins, ties -> smaller index
          bool take = (oi >= 0) && (best_idx < 0 || oa > best_abs || (oa == best_abs && oi < best_idx));
          if (take) { best_abs = oa; best_amount = om; best_idx = oi; }
        }
        if (best_idx < 0) break;
        const double* colv = a.qinv_t + size_t(best_idx) * N;
        double multiplier = best_amount / colv[best_idx];
        for (int i = lane; i < N; i += 32) qd[i] += multiplier * colv[i];
        __syncwarp();
      }
    }
    __syncthreads();
    if (a.clipped) {
      for (int i = threadIdx.x; i < count * DN; i += blockDim.x) {
        const int p = i / DN, k = i - p * DN, ro = first + p;
        const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
        a.clipped[size_t(b) * a.params_problem_stride + size_t(r) * a.params_rollout_stride + k] = q[i];
      }
    }

    // ---- FK + spheres + SDF + velocity + cost -------------------------------------------------------
    if (kSplit) {
      for (int i = threadIdx.x; i < ntiles * K * 32; i += blockDim.x) contrib[i] = Real(0);
      __syncthreads();
    }
    for (int tw = warp; tw < ntiles * S; tw += nwarps) {
      const int tile = tw / S, sgroup = tw - tile * S;      // kSplit: this warp's share of the tile's sphere clusters
      // this lane's point of the concatenated timeline
      const int gp = tile * kTileSteps + lane;
      const int p = min(gp / seg, count - 1);                 // rollout within the pack (clamped: surplus lanes idle on the last one)
      const int t = gp - (gp / seg) * seg - 1 + (gp / seg - p) * seg;   // -1 .. N+1 for real points, beyond for surplus lanes
      const bool real_point = gp / seg < count;
      const int ro = first + p;
      const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
      const double* qp = q + size_t(p) * DN + min(max(t, 0), N - 1);   // this lane's column of the rollout
      const double* padp = pads + size_t(p) * 2 * D + (t < 0 ? 0 : D);
      const bool in_pad = t < 0 || t >= N;
      const bool productive = real_point && lane >= 1 && lane <= kTileSteps && t >= 0 && t < N;
      const bool counts = productive || (a.include_pads && real_point && (t == -1 || t == N));
      int collided = 0, violated = 0;
      Real cost = Real(0), ccost = Real(0);
      Real F[12], saved[kMaxSlots][12];
#pragma unroll
      for (int i = 0; i < 12; ++i) F[i] = Real(0);
      for (int ci = 0; ci < num_cons; ++ci)   // constraints on segments no group joint moves
        if (cons[ci].node < 0) {
          Real cc;
          if (!constraint_cost<Real>(cons[ci], nullptr, cc)) violated |= int(productive);
          ccost += cc;
        }

      for (int n = 0; n < a.num_nodes; ++n) {
        const DevNode<Real>& nd = nodes[n];
        Real Pm[12];
        {
          Real qv = Real(0);
          if (nd.q_index >= 0) qv = Real(in_pad ? padp[nd.q_index] : qp[size_t(nd.q_index) * N]);
          if (nd.type == STOMP_JOINT_REVOLUTE) {
            Real sn, cs;
            Math<Real>::sincos_(qv, &sn, &cs);
#pragma unroll
            for (int i = 0; i < 9; ++i) Pm[i] = nd.A0[i] + cs * nd.A1[i] + sn * nd.A2[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) Pm[9 + i] = nd.p[i];
          } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) Pm[i] = nd.A0[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) Pm[9 + i] = nd.p[i] + qv * nd.ax[i];
          }
        }
        if (nd.parent < 0) {
#pragma unroll
          for (int i = 0; i < 12; ++i) F[i] = Pm[i];
        } else {
          if (nd.load_slot >= 0) {   // branching trees only: the parent is not the previous node
#pragma unroll
            for (int i = 0; i < 12; ++i) F[i] = saved[nd.load_slot][i];
          }
          // F <- F * Pm, row by row in place (row i of the product only needs row i of the parent frame)
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const Real f0 = F[i * 3], f1 = F[i * 3 + 1], f2 = F[i * 3 + 2];
#pragma unroll
            for (int j = 0; j < 3; ++j) F[i * 3 + j] = f0 * Pm[j] + f1 * Pm[3 + j] + f2 * Pm[6 + j];
            F[9 + i] = f0 * Pm[9] + f1 * Pm[10] + f2 * Pm[11] + F[9 + i];
          }
        }
        if (nd.save_slot >= 0) {
#pragma unroll
          for (int i = 0; i < 12; ++i) saved[nd.save_slot][i] = F[i];
        }
        for (int ci = 0; ci < num_cons; ++ci)
          if (cons[ci].node == n) {
            Real cc;
            if (!constraint_cost<Real>(cons[ci], F, cc)) violated |= int(productive);
            ccost += cc;
          }
        const int sph_begin = nd.sphere_begin, sph_end = nd.sphere_end;
        if (sph_end > sph_begin) {
          for (int ci = nd.cluster_begin; ci < nd.cluster_end; ++ci) {
          if (kSplit && (ci % S) != sgroup) continue;
          const DevCluster<Real>& cl = clusters[ci];
          if (kCull) {
            // broad phase: coarse lower bound of the distance at the cluster centre against the cluster's threshold
            const Real qx = F[0] * cl.c[0] + F[1] * cl.c[1] + F[2] * cl.c[2] + F[9];
            const Real qy = F[3] * cl.c[0] + F[4] * cl.c[1] + F[5] * cl.c[2] + F[10];
            const Real qz = F[6] * cl.c[0] + F[7] * cl.c[1] + F[8] * cl.c[2] + F[11];
            // the coarse lookup in fp32: a centre that lands in a neighbouring coarse cell through rounding is covered by the
            // one-cell slack the threshold already carries (engine.cu, upload_cluster_thresholds), and the margin test keeps
            // every fine cell concerned interior
            const float ux = fmaf(float(qx), float(g.inv_res), float(g.nox)) + 0.5f;      // fine cell = floor(u)
            const float uy = fmaf(float(qy), float(g.inv_res), float(g.noy)) + 0.5f;
            const float uz = fmaf(float(qz), float(g.inv_res), float(g.noz)) + 0.5f;
            const float m = float(cl.margin);
            const bool in = (ux >= m) & (ux <= float(g.nx1 + 1) - m) & (uy >= m) & (uy <= float(g.ny1 + 1) - m) & (uz >= m) &
                            (uz <= float(g.nz1 + 1) - m);
            float lb = -1.0f;
            if (in) lb = __ldg(a.cull.g + (int(ux * 0.25f) * a.cull.ny + int(uy * 0.25f)) * a.cull.nz + int(uz * 0.25f));
            const bool clear = Real(lb) >= cl.thr;
            if (__all_sync(0xffffffffu, clear || !(productive || counts))) continue;
          }
          const int jb = cl.begin, je = cl.end;     // locals: the loop must not re-read the bounds from shared memory
          for (int j = jb; j < je; ++j) {
            const unsigned sa = sph_addr + unsigned(j) * unsigned(sizeof(DevSphere<Real>));
            typename SphereRegsOf<Real>::type sp;
            load_sphere(sa, sp);
            const Real s0 = sp.s0, s1 = sp.s1, s2 = sp.s2;
            const Real px = F[0] * s0 + F[1] * s1 + F[2] * s2 + F[9];
            const Real py = F[3] * s0 + F[4] * s1 + F[5] * s2 + F[10];
            const Real pz = F[6] * s0 + F[7] * s1 + F[8] * s2 + F[11];
            int cx, cy, cz;
            Real dist;
            if (kTri) {
              dist = trilinear_distance<Real, kVox>(g, vox, px, py, pz, tab_addr, cx, cy, cz);
            } else {
              const bool inside = voxel_cells(g, px, py, pz, cx, cy, cz);
              // outside the grid (or within one cell of its faces) the reference returns distance 0
              dist = inside ? voxel_distance<Real, kVox>(vox, bricks ? brick_index(cx, cy, cz, g.nby, g.nbz) : (cx * g.sny + cy) * g.snz + cz,
                                                         tab_addr, g.res)
                            : Real(0);
            }
            // three-piece potential (stomp_collision_space.h:209-226), branch-free
            const Real radius = sp.radius, clearance = sp.clearance;
            const Real dd = dist - radius, diff = dd - clearance;
            const Real pot_mid = Real(0.5) * (diff * sp.inv_clearance) * diff;
            const Real pot_neg = -dd + Real(0.5) * clearance;
            const Real pot = dd >= clearance ? Real(0) : (dd >= Real(0) ? pot_mid : pot_neg);
            const bool hit = dist <= radius;
            collided |= int(hit & counts);
            Real vm = Real(0);
            if (__any_sync(0xffffffffu, kDebug || pot != Real(0))) {
              // finite-difference velocity of the sphere centre from the neighbouring lanes' positions (t-1, t+1, t+2): exactly
              // the reference's formulation (positions differentiated, src/stomp_optimizer.cpp:672-690); no velocity frame
              const Real vx = c_m1 * shfl_rel(px, -1) + c_0 * px + c_p1 * shfl_rel(px, 1) + c_p2 * shfl_rel(px, 2);
              const Real vy = c_m1 * shfl_rel(py, -1) + c_0 * py + c_p1 * shfl_rel(py, 1) + c_p2 * shfl_rel(py, 2);
              const Real vz = c_m1 * shfl_rel(pz, -1) + c_0 * pz + c_p1 * shfl_rel(pz, 1) + c_p2 * shfl_rel(pz, 2);
              vm = Math<Real>::sqrt_(vx * vx + vy * vy + vz * vz);
              // product and sum rounded separately (no FMA contraction): the split kernel parks the product in shared memory
              // before it is added, and both kernels must produce the same bits; it is also how the reference's SSE2 build rounds
              const Real contribution = Math<Real>::mul_rn(lds1(sa + 6 * unsigned(sizeof(Real)), Real(0)), pot * vm);   // DevSphere::weight
              if (kSplit) contrib[(tile * K + j) * 32 + lane] = contribution;
              else cost = Math<Real>::add_rn(cost, contribution);
            }
            if (kDebug) {
              if (t >= -1 && t <= N + 1 && (lane >= 1 && lane <= kTileSteps || (tile == 0 && lane == 0) ||
                                             (tile == ntiles - 1 && lane > kTileSteps))) {
                stomp_sphere_debug& rec = a.debug[size_t(t + 1) * K + spheres[j].original_index];
                rec.voxel[0] = cx; rec.voxel[1] = cy; rec.voxel[2] = cz;
                rec.in_collision = hit;
                rec.position[0] = px; rec.position[1] = py; rec.position[2] = pz;
                rec.potential = pot;
                rec.vel_mag = (t >= 0 && t < N && lane >= 1 && lane <= kTileSteps) ? double(vm) : 0.0;
              }
            }
          }
          }
        }
      }
      if (productive && !kSplit) {
        double* out = a.costs + size_t(b) * a.cost_problem_stride + size_t(r) * N;
        out[t] = kCons ? a.obstacle_weight * double(cost) + a.constraint_weight * double(ccost) : a.obstacle_weight * double(cost);
      }
      if (collided && a.collision_free) a.collision_free[size_t(b) * a.flag_problem_stride + a.flag_offset + r] = 0;
      if (kCons && violated && a.constraints_satisfied)
        a.constraints_satisfied[size_t(b) * a.flag_problem_stride + a.flag_offset + r] = 0;
    }
    if (kSplit) {      // the contributions of a timestep, added in sphere order (kCons is never split)
      __syncthreads();
      for (int tile = warp; tile < ntiles; tile += nwarps) {
        const int gp = tile * kTileSteps + lane;
        const int p = min(gp / seg, count - 1);
        const int t = gp - (gp / seg) * seg - 1 + (gp / seg - p) * seg;
        const int ro = first + p;
        const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
        if (gp / seg < count && lane >= 1 && lane <= kTileSteps && t >= 0 && t < N) {
          Real cost = Real(0);
          for (int j = 0; j < K; ++j) cost = Math<Real>::add_rn(cost, contrib[(tile * K + j) * 32 + lane]);
          a.costs[size_t(b) * a.cost_problem_stride + size_t(r) * N + t] = a.obstacle_weight * double(cost);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_torque: the inverse-dynamics term of StompOptimizer::execute (src/stomp_optimizer.cpp:1117-1142, getTorques :1034-1060):
//   costs[t] += torque_cost_weight * sum_j |tau_j(t)|
// One thread per (rollout, free timestep).  q = the joint-limit-projected group trajectory k_cost wrote (`clipped` tap), qd /
// qdd = the 7-tap rules over the padded trajectory (stomp_trajectory.h:286-310), tau = KDL::ChainIdSolver_RNE restated:
// recursive Newton-Euler in segment-tip coordinates, spatial vectors as (linear, angular) pairs like KDL::Twist / Wrench,
// gravity entering as the base acceleration -g.  Off unless torque_cost_weight > 1e-9 (never in the shipped configurations), so
// it is written for clarity, not speed: the per-segment state lives in local memory.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxChain = 24;
struct DevChainLink {
  int type, group;            // stomp_joint_type, group joint index (-1: fixed)
  double rot[9], pos[3], axis[3];
  double m, h[3], I[9];       // KDL::RigidBodyInertia in the segment frame: mass, m * cog, inertia about the frame origin
};
struct TorqueArgs {
  int D, N, n_rollouts, total_rollouts, ns;
  size_t q_problem_stride, q_rollout_stride, cost_problem_stride;   // in doubles
  const double* q;           // joint-limit-projected rollouts, same layout as the rollout parameters
  const double* pad_start;   // [B][D]
  const double* pad_goal;    // [B][D]
  const DevChainLink* chain; // [ns]
  double inv_time, inv_time2, g[3], weight;
  double* costs;             // [..][N], accumulated into
  double* torques;           // optional tap [total_rollouts][N][D]
};

__device__ __forceinline__ void t_cross(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void t_rot_t(const double* R, const double* v, double* o) {   // R^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2];
}
__device__ __forceinline__ void t_rot_n(const double* R, const double* v, double* o) {   // R v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = R[i * 3] * v[0] + R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2];
}
// KDL::Frame::Inverse(Twist): lin' = R^T (lin - p x ang), ang' = R^T ang     (t = {lin[3], ang[3]})
__device__ __forceinline__ void t_to_child(const double* R, const double* p, const double* t, double* o) {
  double c[3], d[3];
  t_cross(p, t + 3, c);
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = t[i] - c[i];
  t_rot_t(R, d, o);
  t_rot_t(R, t + 3, o + 3);
}
// KDL::RigidBodyInertia * Twist -> Wrench {force, torque}
__device__ __forceinline__ void t_inertia(const DevChainLink& L, const double* t, double* w) {
  double hw[3], hv[3], Iw[3];
  t_cross(L.h, t + 3, hw);
  t_cross(L.h, t, hv);
  t_rot_n(L.I, t + 3, Iw);
#pragma unroll
  for (int i = 0; i < 3; ++i) { w[i] = L.m * t[i] - hw[i]; w[3 + i] = Iw[i] + hv[i]; }
}

__global__ void __launch_bounds__(128) k_torque(TorqueArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.total_rollouts * a.N) return;
  const int ro = int(idx / a.N), t = int(idx - (long long)ro * a.N);
  const int b = ro / a.n_rollouts, r = ro - b * a.n_rollouts;
  const double* q = a.q + size_t(b) * a.q_problem_stride + size_t(r) * a.q_rollout_stride;
  const double* ps = a.pad_start + size_t(b) * a.D;
  const double* pg = a.pad_goal + size_t(b) * a.D;
  const double rule_v[7] = {0, 0, -2 / 6.0, -3 / 6.0, 6 / 6.0, -1 / 6.0, 0};
  const double rule_a[7] = {0, -1 / 12.0, 16 / 12.0, -30 / 12.0, 16 / 12.0, -1 / 12.0, 0};

  double XR[kMaxChain][9], Xp[kMaxChain][3], S[kMaxChain][6], f[kMaxChain][6];
  double v[6] = {0, 0, 0, 0, 0, 0}, acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < a.ns; ++i) {
    const DevChainLink& L = a.chain[i];
    double q_ = 0.0, qd_ = 0.0, qdd_ = 0.0;
    if (L.group >= 0) {
      const double* qj = q + size_t(L.group) * a.N;
      for (int k = -3; k <= 3; ++k) {
        const int tt = t + k;
        const double val = tt < 0 ? ps[L.group] : (tt >= a.N ? pg[L.group] : qj[tt]);
        if (k == 0) q_ = val;
        qd_ += (a.inv_time * rule_v[k + 3]) * val;
        qdd_ += (a.inv_time2 * rule_a[k + 3]) * val;
      }
    }
    // segment.pose(q): Frame(Rot(axis, q) * rot, pos) | Frame(rot, pos + q * axis) | Frame(rot, pos)
    double unit[6] = {0, 0, 0, 0, 0, 0};
    if (L.type == STOMP_JOINT_REVOLUTE) {
      double sn, cs;
      sincos(q_, &sn, &cs);
      const double vt = 1.0 - cs, x = L.axis[0], y = L.axis[1], z = L.axis[2];
      const double Rq[9] = {cs + vt * x * x, vt * x * y - sn * z, vt * x * z + sn * y,
                            vt * x * y + sn * z, cs + vt * y * y, vt * y * z - sn * x,
                            vt * x * z - sn * y, vt * y * z + sn * x, cs + vt * z * z};
#pragma unroll
      for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int n = 0; n < 3; ++n) XR[i][m * 3 + n] = Rq[m * 3] * L.rot[n] + Rq[m * 3 + 1] * L.rot[3 + n] + Rq[m * 3 + 2] * L.rot[6 + n];
#pragma unroll
      for (int m = 0; m < 3; ++m) { Xp[i][m] = L.pos[m]; unit[3 + m] = L.group >= 0 ? L.axis[m] : 0.0; }
    } else {
#pragma unroll
      for (int m = 0; m < 9; ++m) XR[i][m] = L.rot[m];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        Xp[i][m] = L.pos[m] + (L.type == STOMP_JOINT_PRISMATIC ? q_ * L.axis[m] : 0.0);
        unit[m] = (L.type == STOMP_JOINT_PRISMATIC && L.group >= 0) ? L.axis[m] : 0.0;
      }
    }
    // S = X.M^-1 * unit twist (the segment tip sits on the joint origin, so no reference-point shift); vj = S * qd
    t_rot_t(XR[i], unit, S[i]);
    t_rot_t(XR[i], unit + 3, S[i] + 3);
    double vj[6], vp[6], ap[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) vj[m] = S[i][m] * qd_;
    if (i == 0) {
      const double ag[6] = {-a.g[0], -a.g[1], -a.g[2], 0, 0, 0};
#pragma unroll
      for (int m = 0; m < 6; ++m) vp[m] = 0.0;
      t_to_child(XR[i], Xp[i], ag, ap);
    } else {
      t_to_child(XR[i], Xp[i], v, vp);
      t_to_child(XR[i], Xp[i], acc, ap);
    }
#pragma unroll
    for (int m = 0; m < 6; ++m) v[m] = vp[m] + vj[m];
    double c1[3], c2[3], c3[3];
    t_cross(v + 3, vj, c1);       // Twist * Twist: (ang x lin' + lin x ang', ang x ang')
    t_cross(v, vj + 3, c2);
    t_cross(v + 3, vj + 3, c3);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      acc[m] = ap[m] + S[i][m] * qdd_ + c1[m] + c2[m];
      acc[3 + m] = ap[3 + m] + S[i][3 + m] * qdd_ + c3[m];
    }
    double Ia[6], Iv[6], d1[3], d2[3], d3[3];
    t_inertia(L, acc, Ia);
    t_inertia(L, v, Iv);
    t_cross(v + 3, Iv, d1);       // Twist * Wrench: (ang x force, ang x torque + lin x force)
    t_cross(v + 3, Iv + 3, d2);
    t_cross(v, Iv, d3);
#pragma unroll
    for (int m = 0; m < 3; ++m) { f[i][m] = Ia[m] + d1[m]; f[i][3 + m] = Ia[3 + m] + d2[m] + d3[m]; }
  }
  double sum = 0.0;
  for (int i = a.ns - 1; i >= 0; --i) {
    const int j = a.chain[i].group;
    if (j >= 0) {
      double tau = 0.0;
#pragma unroll
      for (int m = 0; m < 6; ++m) tau += S[i][m] * f[i][m];
      sum += fabs(tau);
      if (a.torques) a.torques[(size_t(ro) * a.N + t) * a.D + j] = tau;
    }
    if (i != 0) {                 // Frame * Wrench into the parent's coordinates
      double F[3], T[3], pF[3];
      t_rot_n(XR[i], f[i], F);
      t_rot_n(XR[i], f[i] + 3, T);
      t_cross(Xp[i], F, pF);
#pragma unroll
      for (int m = 0; m < 3; ++m) { f[i - 1][m] += F[m]; f[i - 1][3 + m] += T[m] + pF[m]; }
    }
  }
  a.costs[size_t(b) * a.cost_problem_stride + size_t(r) * a.N + t] += a.weight * sum;
}

// ---------------------------------------------------------------------------------------------
// block-wide reverse inclusive scan over threads (warp shuffles + one shared array)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_suffix_scan(double v, double* swarp /* [32] */, double carry_in, double* carry_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    double o = __shfl_down_sync(0xffffffffu, v, off);
    if (lane + off < 32) v += o;
  }
  if (lane == 0) swarp[warp] = v;  // total of the warp
  __syncthreads();
  double add = carry_in;
  for (int w = warp + 1; w < nwarps; ++w) add += swarp[w];
  double total = carry_in;
  for (int w = 0; w < nwarps; ++w) total += swarp[w];
  __syncthreads();
  *carry_out = total;
  return v + add;
}

// k_cumulative: CTA per (b, r).  cumulative[b][r][d][t] = S + C[d] (suffix-summed over t when enabled) and
// totals[b][r] = sum_t S + sum_d sum_t C[d]  (Rollout::getCost).
__global__ void k_cumulative(int R, int r_begin, int r_count, int D, int N, int use_cumulative,
                             const double* __restrict__ state, const double* __restrict__ control,
                             double* __restrict__ cumulative, double* __restrict__ totals) {
  __shared__ double swarp[32];
  __shared__ double sred[32];
  const int b = blockIdx.x / r_count, r = r_begin + (blockIdx.x - b * r_count);
  const double* S = state + (size_t(b) * R + r) * N;
  const double* C = control + (size_t(b) * R + r) * D * N;
  double* out = cumulative + (size_t(b) * R + r) * D * N;
  double acc = 0.0;
  for (int t = threadIdx.x; t < N; t += blockDim.x) acc += S[t];
  const int nchunks = (N + blockDim.x - 1) / blockDim.x;
  for (int d = 0; d < D; ++d) {
    double carry = 0.0;
    for (int chunk = nchunks - 1; chunk >= 0; --chunk) {
      int t = chunk * blockDim.x + threadIdx.x;
      double c = 0.0, s = 0.0;
      if (t < N) { c = C[size_t(d) * N + t]; s = S[t]; acc += c; }
      double v = s + c;
      if (use_cumulative) {
        double next;
        v = block_suffix_scan(t < N ? v : 0.0, swarp, carry, &next);
        carry = next;
      }
      if (t < N && cumulative) out[size_t(d) * N + t] = v;   // nullptr: only the totals are wanted
    }
  }
  // block reduce acc
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) sred[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && totals) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += sred[w];
    totals[size_t(b) * (R + 1) + r] = s;
  }
}

// k_totals: Rollout::getCost() when k_generate has already summed the control costs per vector (GenArgs::csum) and k_update
// adds S + C itself: totals[b][r] = sum_t S[t] + sum_d csum[d], one warp per rollout — 1/8 of the bytes k_cumulative reads.
__global__ void __launch_bounds__(128) k_totals(int R, int r_begin, int r_count, int D, int N, long long num, const double* __restrict__ state,
                                                const double* __restrict__ csum, double* __restrict__ totals) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= num) return;
  const int b = int(w / r_count), r = r_begin + int(w - (long long)b * r_count);
  const double* S = state + (size_t(b) * R + r) * N;
  double acc = 0.0;
  for (int t = lane; t < N; t += 32) acc += S[t];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    const double* c = csum + (size_t(b) * R + r) * D;
    for (int d = 0; d < D; ++d) acc += c[d];
    totals[size_t(b) * (R + 1) + r] = acc;
  }
}

// k_extra_total: totals[b][R] of the extra rollout = sum S_extra + sum control_extra
__global__ void k_extra_total(int R, int D, int N, const double* __restrict__ state_extra,
                              const double* __restrict__ control_extra, double* __restrict__ totals,
                              double* __restrict__ noiseless_cost) {
  __shared__ double sred[32];
  const int b = blockIdx.x;
  double acc = 0.0, accs = 0.0;
  for (int t = threadIdx.x; t < N; t += blockDim.x) accs += state_extra[size_t(b) * N + t];
  for (int i = threadIdx.x; i < D * N; i += blockDim.x) acc += control_extra[size_t(b) * D * N + i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    accs += __shfl_xor_sync(0xffffffffu, accs, off);
  }
  __shared__ double sred2[32];
  if (lane == 0) { sred[warp] = acc; sred2[warp] = accs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, s2 = 0.0;
    for (int w = 0; w < nwarps; ++w) s += sred[w], s2 += sred2[w];
    totals[size_t(b) * (R + 1) + R] = s2 + s;
    if (noiseless_cost) noiseless_cost[b] = s2;
  }
}

// ---------------------------------------------------------------------------------------------
// k_update: CTA per (b, d), thread per timestep.  min / max / exp-normalise over rollouts, probability
// weighted noise, projection through M (banded solve), theta += update.
// ---------------------------------------------------------------------------------------------
struct UpdateArgs {
  int R, D, N, apply;
  int dims_per_cta;          // a CTA handles dims [g*dims_per_cta, ...) of one problem
  const double* dense_ms;    // [N][N] rows j of R^-1 diag(s) (engine.cu), or nullptr: banded solves
  int use_dmma;              // dense projection on the tensor pipe (mma.sync m8n8k4 f64) instead of the scalar DFMA loop
  const double* cumulative;  // [B][R][D][N]
  const double* state;       // [B][R][N] or nullptr.  Given (use_cumulative_costs == 0): the cost of (r, d, t) is formed here
                             // as state[r][t] + cumulative[r][d][t] with `cumulative` pointing at the CONTROL costs — the very
                             // addition k_cumulative does, so k_update need not wait for that kernel
  const double* noise;       // [B][R][D][N]
  double* probabilities;     // optional tap
  double* updates;           // [B][D][N]
  double* theta;             // [B][D][N]
  Band band;
  // optional fusion of addExtraRollouts' control cost: control cost of the UPDATED theta with zero noise
  double* extra_control;     // [B][D][N] or nullptr
  const double* pad_start;   // [B][D]
  const double* pad_goal;    // [B][D]
  double control_weight;     // 0.5 * control_cost_weight
  Stencil st;
};

// exp(x) for the PI^2 weights, x = -10 (c - min) / (max - min) in [-10, 0] (valid for |x| < 700).  The library exp is ~100
// instructions with its range checks; ncu showed k_update issue bound on them (39 M warp instructions, 1 400 per element
// for ten weights).  Cody-Waite reduction x = k ln2 + t, |t| <= 0.35, Taylor polynomial of degree 13 (remainder 6e-18),
// 2^k added to the exponent field: ~22 instructions, within 2 ulp of the correctly rounded value.
__device__ __forceinline__ double exp_weight(double x) {
  const double kf = rint(x * 1.4426950408889634074);
  double t = fma(kf, -6.93147180369123816490e-01, x);
  t = fma(kf, -1.90821492927058770002e-10, t);
  double p = 1.6059043836821613e-10;            // 1/13!
  p = fma(p, t, 2.0876756987868100e-09);        // 1/12!
  p = fma(p, t, 2.5052108385441720e-08);        // 1/11!
  p = fma(p, t, 2.7557319223985893e-07);        // 1/10!
  p = fma(p, t, 2.7557319223985888e-06);        // 1/9!
  p = fma(p, t, 2.4801587301587302e-05);        // 1/8!
  p = fma(p, t, 1.9841269841269841e-04);        // 1/7!
  p = fma(p, t, 1.3888888888888889e-03);        // 1/6!
  p = fma(p, t, 8.3333333333333332e-03);        // 1/5!
  p = fma(p, t, 4.1666666666666664e-02);        // 1/4!
  p = fma(p, t, 1.6666666666666666e-01);        // 1/3!
  p = fma(p, t, 0.5);
  p = fma(p, t, 1.0);
  p = fma(p, t, 1.0);
  return __longlong_as_double(__double_as_longlong(p) + (static_cast<long long>(__double2int_rn(kf)) << 52));
}

// One (problem, dimension, timestep) element of the PI^2 update (policy_improvement.cpp:283-340): min / max of the cumulative
// costs over rollouts, exp(-10 (c - min) / (max - min)) weights, probability-weighted noise.  Rollouts are visited in groups
// of kGroup with all loads of a group issued before the first use (the loop count is a run-time value, so the compiler does
// not software-pipeline it; ncu showed 45 % of the stalls on these loads).  kCache: R <= kGroup, every value is loaded once,
// up front, and both passes run from registers — one memory round trip per element.  Same operations in the same order in
// every instantiation: results do not depend on which one ran.
template <int kGroup, bool kCache, bool kSum>
__device__ __forceinline__ double weighted_noise(const double* __restrict__ c, const double* __restrict__ s, const double* __restrict__ e,
                                                 size_t rstride, size_t sstride, int R, double* __restrict__ prob_out) {
  // cost of rollout r: the cumulative cost, or (kSum) state cost + control cost = what k_cumulative stores without the scan
  auto cost = [&](int r) -> double { return kSum ? s[r * sstride] + c[r * rstride] : c[r * rstride]; };
  double cc[kCache ? kGroup : 1], ec[kCache ? kGroup : 1];
  if (kCache) {
    double sv[kSum ? kGroup : 1];
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      const bool in = j < R;
      cc[j] = in ? c[j * rstride] : 0.0;
      if (kSum) sv[j] = in ? s[j * sstride] : 0.0;
      ec[j] = in ? e[j * rstride] : 0.0;
    }
    if (kSum) {
#pragma unroll
      for (int j = 0; j < kGroup; ++j) cc[j] = sv[j] + cc[j];
    }
  }
  double mn = kCache ? cc[0] : cost(0), mx = mn;
  for (int r0 = 0; r0 < R; r0 += kGroup) {
    double cv[kGroup];
#pragma unroll
    for (int j = 0; j < kGroup; ++j) cv[j] = r0 + j < R ? (kCache ? cc[j] : cost(r0 + j)) : mn;
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      if (cv[j] < mn) mn = cv[j];
      if (cv[j] > mx) mx = cv[j];
    }
  }
  double denom = mx - mn;
  if (denom < 1e-8) denom = 1e-8;
  const double h = -10.0 / denom;
  double p_sum = 0.0, acc = 0.0;
  for (int r0 = 0; r0 < R; r0 += kGroup) {
    double cv[kGroup], ev[kGroup];
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      const bool in = r0 + j < R;
      cv[j] = in ? (kCache ? cc[j] : cost(r0 + j)) : 0.0;
      ev[j] = in ? (kCache ? ec[j] : e[(r0 + j) * rstride]) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kGroup; ++j)
      if (r0 + j < R) {
        const double w = exp_weight(h * (cv[j] - mn));
        p_sum += w;
        acc += ev[j] * w;
      }
  }
  const double inv = 1.0 / p_sum;
  if (prob_out)
    for (int r = 0; r < R; ++r) prob_out[r * rstride] = exp_weight(h * (cost(r) - mn)) * inv;
  return acc * inv;
}

// CTA per (problem, group of dimensions).  Phase 1 (all threads, thread per timestep): min / max over
// rollouts, exp-normalised weights, probability-weighted noise.  Phase 2: the projection through M = R^-1 diag(s) — as the
// dense product it stands for (all threads, a.dense_ms = [j][t] rows of R^-1 diag(s); ncu: the two banded solves ran on one
// lane per dimension for a fifth of the kernel's time while the other warps waited at the barrier), or, for long series, as
// the two banded triangular solves.  Phase 3: theta += update, and the control cost of the updated trajectory from a padded
// copy in shared memory.  Seven CTAs per SM keep a 1024-problem batch in one wave (at 96 registers it took two: 0.088 ms
// instead of 0.061).
// kThreads / kMinBlocks: 128 x 7 for batches that fill the machine; 512 x 1 for small batches, where a CTA is alone on its SM and
// the projection's 8-row tiles (13 for N = 99) each get a warp of their own instead of queueing four deep on four warps.
// kDirect: the costs are formed here as S + C[d] (UpdateArgs::state); a separate instantiation so that the other one keeps its
// registers (as a run-time branch it cost the 7-CTA kernel 32 bytes of spills and 8 us at C2).
template <int kThreads, int kMinBlocks, bool kDirect>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_update(UpdateArgs a) {
  extern __shared__ double smem[];
  const int N = a.N, R = a.R, D = a.D, G = a.dims_per_cta;
  const int Nall = N + 2 * kPad;
  const int stride = Nall | 1;
  const bool dense = a.dense_ms != nullptr;
  double* u = smem;                           // [G][stride]  weighted noise; later the padded updated trajectory
  double* y = u + size_t(G) * stride;         // [G][stride]  dense: projected update
  double* fold = y + size_t(G) * stride;      // [G][2 kPad]  control costs of the padded rows
  double* sfw = fold + size_t(G) * 2 * kPad;  // [N][8]       band path only
  double* sbw = sfw + N * 8;                  // [N][8]
  const int groups = (D + G - 1) / G;
  const int b = blockIdx.x / groups, d0 = (blockIdx.x - b * groups) * G;
  const int nd = min(G, D - d0);
  if (!dense)
    for (int k = threadIdx.x; k < N * 8; k += blockDim.x) sfw[k] = a.band.fw[k], sbw[k] = a.band.bw[k];
  const size_t rstride = size_t(D) * N;
  for (int k = threadIdx.x; k < nd * N; k += blockDim.x) {
    const int dl = k / N, t = k - dl * N, d = d0 + dl;
    const size_t base = (size_t(b) * R * D + d) * N + t;
    double* prob = a.probabilities ? a.probabilities + base : nullptr;
    double v;
    if (kDirect) {
      const double* sp = a.state + size_t(b) * R * N + t;
      v = R <= 10 ? weighted_noise<10, true, true>(a.cumulative + base, sp, a.noise + base, rstride, size_t(N), R, prob)
                  : weighted_noise<5, false, true>(a.cumulative + base, sp, a.noise + base, rstride, size_t(N), R, prob);
    } else {
      v = R <= 10 ? weighted_noise<10, true, false>(a.cumulative + base, nullptr, a.noise + base, rstride, 0, R, prob)
                  : weighted_noise<5, false, false>(a.cumulative + base, nullptr, a.noise + base, rstride, 0, R, prob);
    }
    if (!dense) v *= a.band.proj_scale[t];      // the dense matrix carries the scaling
    u[dl * stride + t] = v;
  }
  __syncthreads();
  if (dense && a.use_dmma) {
    // The batched projection as a tensor-core GEMM (north_star: "on tensor cores when batched across many problems"):
    //   Y[t][q] = sum_j A[t][j] U[j][q],  A[t][j] = R^-1[j][t] s_j (dense_ms[j][t]),  U[j][q] = weighted noise of dimension q
    // with mma.sync.aligned.m8n8k4.f64 (SASS DMMA): M = timesteps (8-row tiles over the warps), N = the CTA's dimensions
    // (8-column tiles), K = j in steps of 4.  One DMMA replaces 8 x 8 x 4 scalar DFMAs of the loop below; A fragments come
    // from the L1/L2-resident matrix (for one k, lanes with consecutive rows read consecutive t), B fragments from `u`.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int frow = lane >> 2, fk = lane & 3;           // fragment coordinates of this lane
    const int mtiles = (N + 7) >> 3, ksteps = (N + 3) >> 2;
    for (int q0 = 0; q0 < nd; q0 += 8) {
      const int qb = q0 + frow;                            // B fragment column (dimension) of this lane
      const double* ub = u + size_t(min(qb, nd - 1)) * stride;
      const bool qv = qb < nd;
      for (int mt = warp; mt < mtiles; mt += nwarps) {
        const int t = mt * 8 + frow;                       // A fragment row (timestep) of this lane
        const bool tv = t < N;
        const double* ap = a.dense_ms + min(t, N - 1);
        double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks) {
          const int j = ks * 4 + fk;
          const bool jv = j < N;
          const double av = (tv && jv) ? __ldg(ap + size_t(j) * N) : 0.0;
          const double bv = (qv && jv) ? ub[j] : 0.0;
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
        }
        // accumulator fragment: row = lane >> 2, columns 2 * (lane & 3) + {0, 1}
        const int qc = q0 + 2 * fk;
        if (tv && qc < nd) y[qc * stride + t] = c0;
        if (tv && qc + 1 < nd) y[(qc + 1) * stride + t] = c1;
      }
    }
  } else if (dense) {
    constexpr int kDims = 8;         // dimensions accumulated per pass over the matrix column
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
      for (int q0 = 0; q0 < nd; q0 += kDims) {
        double acc[kDims];
#pragma unroll
        for (int q = 0; q < kDims; ++q) acc[q] = 0.0;
        const double* mp = a.dense_ms + t;
#pragma unroll 4
        for (int j = 0; j < N; ++j, mp += N) {
          const double m = __ldg(mp);
#pragma unroll
          for (int q = 0; q < kDims; ++q)
            if (q0 + q < nd) acc[q] = fma(m, u[(q0 + q) * stride + j], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < kDims; ++q)
          if (q0 + q < nd) y[(q0 + q) * stride + t] = acc[q];
      }
    }
  } else if (threadIdx.x < nd) {
    double* x = u + threadIdx.x * stride;
    band_forward(x, sfw, N);
    band_backward(x, sbw, N);
  }
  __syncthreads();
  const double* upd = dense ? y : u;
  const size_t off = (size_t(b) * D + d0) * N;
  const bool control = a.apply && a.extra_control;
  for (int k = threadIdx.x; k < nd * N; k += blockDim.x) {
    const int dl = k / N, t = k - dl * N;
    const double v = upd[dl * stride + t];
    a.updates[off + k] = v;
    if (a.apply) {
      const double th = a.theta[off + k] + v;
      a.theta[off + k] = th;
      if (control) (dense ? u : y)[dl * stride + kPad + t] = th;   // padded copy for the stencils (the array phase 2 is done with)
    }
  }
  if (!control) return;
  // control cost of the noise-less (updated) trajectory: 7-tap stencils over [pads, theta, pads]
  // (covariant_trajectory_policy.cpp:228-255 with zero noise; what addExtraRollouts computes for the extra rollout)
  double* xp = dense ? u : y;
  for (int k = threadIdx.x; k < nd * 2 * kPad; k += blockDim.x) {
    const int dl = k / (2 * kPad), q = k - dl * 2 * kPad, d = d0 + dl;
    xp[dl * stride + (q < kPad ? q : N + q)] = q < kPad ? a.pad_start[size_t(b) * D + d] : a.pad_goal[size_t(b) * D + d];
  }
  __syncthreads();
  auto cost_at = [&](const double* row, int p, bool interior) -> double {   // padded row p; interior: every tap is in range
    double cost = 0.0;
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      if (a.st.weight[kk] == 0.0) continue;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int idx = p + j - 3;
        if (!interior && (idx < 0 || idx >= Nall)) continue;   // dropped taps of the differentiation matrices
        acc += a.st.coef[kk][j] * row[idx];
      }
      cost += a.control_weight * a.st.weight[kk] * (acc * acc);
    }
    return cost;
  };
  // the padded rows' costs belong to the first / last free row; one thread each instead of twelve in a row on two threads
  for (int k = threadIdx.x; k < nd * 2 * kPad; k += blockDim.x) {
    const int dl = k / (2 * kPad), q = k - dl * 2 * kPad;
    fold[k] = cost_at(xp + dl * stride, q < kPad ? q : Nall - 1 - (q - kPad), false);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nd * N; k += blockDim.x) {
    const int dl = k / N, t = k - dl * N;
    double c = cost_at(xp + dl * stride, t + kPad, true);
    if (t == 0)
      for (int i = 0; i < kPad; ++i) c += fold[dl * 2 * kPad + i];
    if (t == N - 1)
      for (int i = 0; i < kPad; ++i) c += fold[dl * 2 * kPad + kPad + i];
    a.extra_control[off + k] = c;
  }
}

// ---------------------------------------------------------------------------------------------
// k_track_best: the per-iteration bookkeeping of StompOptimizer::optimize (src/stomp_optimizer.cpp:296-347), CTA per
// problem.  state[b] = {collision_free_iteration_, success_iteration, collision_success_iteration,
// last_improvement_iteration_, iterations, done}; best_cost[b]; best[b][D][N] <- clipped noise-less trajectory.
// ---------------------------------------------------------------------------------------------
struct TrackState {
  int collision_free_iteration, success_iteration, collision_success_iteration, last_improvement_iteration, iterations, done;
};

__global__ void k_track_best(int iteration, int max_after_collision_free, int DN, int flag_stride, int flag_offset,
                             const double* __restrict__ noiseless_cost, const int* __restrict__ collision_free,
                             const int* __restrict__ constraints_satisfied, const double* __restrict__ trajectory, TrackState* __restrict__ state, double* __restrict__ best_cost,
                             double* __restrict__ best, double* __restrict__ cost_log, int B, int* __restrict__ num_done) {
  const int b = blockIdx.x;
  __shared__ int s_copy;
  if (threadIdx.x == 0) {
    TrackState st = state[b];
    s_copy = 0;
    if (!st.done) {
      const double cost = noiseless_cost[b];
      const bool cf = collision_free[size_t(b) * flag_stride + flag_offset] != 0;
      const bool cs = constraints_satisfied[size_t(b) * flag_stride + flag_offset] != 0;
      st.collision_free_iteration = (cf && cs) ? st.collision_free_iteration + 1 : 0;
      if (cf && st.collision_success_iteration == -1) st.collision_success_iteration = iteration;
      if (cf && cs && st.success_iteration == -1) st.success_iteration = iteration;
      if (cost_log) cost_log[size_t(iteration) * B + b] = cost;
      if (iteration == 0) {
        best_cost[b] = cost;
        s_copy = 1;
      } else if (cost < best_cost[b] && cf && cs) {
        best_cost[b] = cost;
        st.last_improvement_iteration = iteration;
        s_copy = 1;
      }
      st.iterations = iteration + 1;
      if (st.collision_free_iteration >= max_after_collision_free) {
        st.done = 1;
        atomicAdd(num_done, 1);
      }
      state[b] = st;
    }
  }
  __syncthreads();
  if (s_copy)
    for (int i = threadIdx.x; i < DN; i += blockDim.x) best[size_t(b) * DN + i] = trajectory[size_t(b) * DN + i];
}

// ---------------------------------------------------------------------------------------------
// Distance-field construction (StompCollisionSpace::setStartState, src/stomp_collision_space.cpp:154-297).
//   k_sdf_mark  : one thread per lattice point of a collision object -> world point -> cell -> occupancy
//   k_edt_pass  : separable exact squared Euclidean distance transform with the reference's cap, one axis per pass
// ---------------------------------------------------------------------------------------------
struct SdfShape {      // one box or cylinder, lattice coordinates pre-accumulated on the host like the reference's loops
  double position[3];
  double R[9];         // KDL::Rotation::Quaternion(x, y, z, w)
  double radius;       // > 0: cylinder test sqrt(xdist^2 + ydist^2) <= radius
  int nx, ny, nz;      // lattice counts
  int x_off, y_off, z_off;  // offsets into the lattice coordinate array
  long long first;     // index of the shape's first lattice point
};

__global__ void k_sdf_mark(int num_shapes, long long total, const SdfShape* __restrict__ shapes, const double* __restrict__ lattice,
                           double ox, double oy, double oz, double res, int nx, int ny, int nz, uint8_t* __restrict__ occ) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = num_shapes - 1;
    while (lo < hi) {   // last shape whose first point is <= i
      int mid = (lo + hi + 1) >> 1;
      if (shapes[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const SdfShape& s = shapes[lo];
    long long k = i - s.first;
    const int iz = int(k % s.nz); k /= s.nz;
    const int iy = int(k % s.ny);
    const int ix = int(k / s.ny);
    const double x = lattice[s.x_off + ix], y = lattice[s.y_off + iy], z = lattice[s.z_off + iz];
    if (s.radius > 0.0) {
      const double xdist = fabs(s.position[0] - x), ydist = fabs(s.position[1] - y);
      if (!(sqrt(xdist * xdist + ydist * ydist) <= s.radius)) continue;
    }
    const double px = s.position[0] - x, py = s.position[1] - y, pz = s.position[2] - z;
    const double wx = s.R[0] * px + s.R[1] * py + s.R[2] * pz + s.position[0];
    const double wy = s.R[3] * px + s.R[4] * py + s.R[5] * pz + s.position[1];
    const double wz = s.R[6] * px + s.R[7] * py + s.R[8] * pz + s.position[2];
    const double tx = (wx - ox) / res, ty = (wy - oy) / res, tz = (wz - oz) / res;
    if (!(fabs(tx) < 1e9 && fabs(ty) < 1e9 && fabs(tz) < 1e9)) continue;
    const int cx = int(round(tx)), cy = int(round(ty)), cz = int(round(tz));
    if (cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz) continue;
    occ[(size_t(cx) * ny + cy) * nz + cz] = 1;
  }
}

// collision-map points (the "points" namespace of the environment, src/stomp_collision_space.cpp:205-213): each point occupies
// the cell it falls into
__global__ void k_sdf_mark_points(long long total, const double* __restrict__ pts, double ox, double oy, double oz, double res, int nx,
                                  int ny, int nz, uint8_t* __restrict__ occ) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const double tx = (pts[3 * i] - ox) / res, ty = (pts[3 * i + 1] - oy) / res, tz = (pts[3 * i + 2] - oz) / res;
    if (!(fabs(tx) < 1e9 && fabs(ty) < 1e9 && fabs(tz) < 1e9)) continue;
    const int cx = int(round(tx)), cy = int(round(ty)), cz = int(round(tz));
    if (cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz) continue;
    occ[(size_t(cx) * ny + cy) * nz + cz] = 1;
  }
}

// robot bodies / primitive collision bodies (StompCollisionSpace::getVoxelsInBody, src/stomp_collision_space.cpp:590-650): one
// thread per lattice point  centre + k * res  of the body's bounding cube; kept when strictly inside the scaled + padded shape
struct SdfBody {
  double c[3];         // bounding-sphere centre = body origin (world)
  double Rt[9];        // world -> body rotation
  double p[3];         // sphere: {r, -, -}; box: half extents; cylinder: {r, half length, -}  (scaled, then padded)
  int type;
  int gmin[3], gn[3];  // first lattice index and count per axis
  long long first;
};

__global__ void k_sdf_mark_bodies(int num_bodies, long long total, const SdfBody* __restrict__ bodies, double ox, double oy, double oz,
                                  double res, int nx, int ny, int nz, uint8_t* __restrict__ occ) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = num_bodies - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (bodies[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const SdfBody& b = bodies[lo];
    long long k = i - b.first;
    const int iz = int(k % b.gn[2]); k /= b.gn[2];
    const int iy = int(k % b.gn[1]);
    const int ix = int(k / b.gn[1]);
    // gridToWorld(centre, g): g * resolution + centre (include/stomp_motion_planner/stomp_collision_space.h:237-241)
    const double xw = (b.gmin[0] + ix) * res + b.c[0], yw = (b.gmin[1] + iy) * res + b.c[1], zw = (b.gmin[2] + iz) * res + b.c[2];
    const double dx = xw - b.c[0], dy = yw - b.c[1], dz = zw - b.c[2];
    const double lx = b.Rt[0] * dx + b.Rt[1] * dy + b.Rt[2] * dz, ly = b.Rt[3] * dx + b.Rt[4] * dy + b.Rt[5] * dz,
                 lz = b.Rt[6] * dx + b.Rt[7] * dy + b.Rt[8] * dz;
    bool inside;
    if (b.type == STOMP_BODY_SPHERE) inside = lx * lx + ly * ly + lz * lz < b.p[0] * b.p[0];
    else if (b.type == STOMP_BODY_BOX) inside = fabs(lx) < b.p[0] && fabs(ly) < b.p[1] && fabs(lz) < b.p[2];
    else inside = fabs(lz) < b.p[1] && lx * lx + ly * ly < b.p[0] * b.p[0];
    if (!inside) continue;
    const double tx = (xw - ox) / res, ty = (yw - oy) / res, tz = (zw - oz) / res;
    if (!(fabs(tx) < 1e9 && fabs(ty) < 1e9 && fabs(tz) < 1e9)) continue;
    const int cx = int(round(tx)), cy = int(round(ty)), cz = int(round(tz));
    if (cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz) continue;
    occ[(size_t(cx) * ny + cy) * nz + cz] = 1;
  }
}

// k_sdf_mark_meshes: getVoxelsInBody for convex-hull mesh bodies (src/stomp_collision_space.cpp:625-646): a lattice point is
// kept when the +z ray from it crosses the (scaled, padded, posed) hull's triangles an odd number of times.  A crossing is
// decided in the xy projection with the rasteriser's half-open edge rule (a point exactly on a shared edge belongs to one of
// the two triangles), then by the sign of the barycentric height above the point.  Every product and sum is individually
// rounded (no FMA contraction): the NumPy restatement the tests hold this to does the same arithmetic, and the cells agree exactly.
struct SdfMesh {
  double c[3];         // bounding-sphere centre (world)
  int gmin[3], gn[3];  // first lattice index and count per axis
  long long first;     // first lattice point of this mesh in the launch
  int tri_first, tri_count;   // its triangles in the [*][9] world-frame vertex array
};

__device__ __forceinline__ bool edge_owns(double dx, double dy) { return dy > 0.0 || (dy == 0.0 && dx < 0.0); }

__global__ void k_sdf_mark_meshes(int num_meshes, long long total, const SdfMesh* __restrict__ meshes, const double* __restrict__ tris,
                                  double ox, double oy, double oz, double res, int nx, int ny, int nz, uint8_t* __restrict__ occ) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = num_meshes - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (meshes[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const SdfMesh& m = meshes[lo];
    long long k = i - m.first;
    const int iz = int(k % m.gn[2]); k /= m.gn[2];
    const int iy = int(k % m.gn[1]);
    const int ix = int(k / m.gn[1]);
    const double xw = __dadd_rn(__dmul_rn(double(m.gmin[0] + ix), res), m.c[0]), yw = __dadd_rn(__dmul_rn(double(m.gmin[1] + iy), res), m.c[1]),
                 zw = __dadd_rn(__dmul_rn(double(m.gmin[2] + iz), res), m.c[2]);
    int crossings = 0;
    for (int t = 0; t < m.tri_count; ++t) {
      const double* T = tris + size_t(m.tri_first + t) * 9;
      double ax = __dsub_rn(T[0], xw), ay = __dsub_rn(T[1], yw), az = __dsub_rn(T[2], zw);
      double bx = __dsub_rn(T[3], xw), by = __dsub_rn(T[4], yw), bz = __dsub_rn(T[5], zw);
      double cx = __dsub_rn(T[6], xw), cy = __dsub_rn(T[7], yw), cz = __dsub_rn(T[8], zw);
      const double area = __dsub_rn(__dmul_rn(__dsub_rn(bx, ax), __dsub_rn(cy, ay)), __dmul_rn(__dsub_rn(by, ay), __dsub_rn(cx, ax)));
      if (area == 0.0) continue;             // seen edge-on from above: the ray runs inside the triangle's plane
      if (area < 0.0) {                      // counter-clockwise in the xy projection
        double s;
        s = bx; bx = cx; cx = s;
        s = by; by = cy; cy = s;
        s = bz; bz = cz; cz = s;
      }
      const double wa = __dsub_rn(__dmul_rn(bx, cy), __dmul_rn(by, cx));   // edge b -> c
      const double wb = __dsub_rn(__dmul_rn(cx, ay), __dmul_rn(cy, ax));   // edge c -> a
      const double wc = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));   // edge a -> b
      const bool in = (wa > 0.0 || (wa == 0.0 && edge_owns(__dsub_rn(cx, bx), __dsub_rn(cy, by)))) &&
                      (wb > 0.0 || (wb == 0.0 && edge_owns(__dsub_rn(ax, cx), __dsub_rn(ay, cy)))) &&
                      (wc > 0.0 || (wc == 0.0 && edge_owns(__dsub_rn(bx, ax), __dsub_rn(by, ay))));
      if (!in) continue;
      const double h = __dadd_rn(__dadd_rn(__dmul_rn(wa, az), __dmul_rn(wb, bz)), __dmul_rn(wc, cz));   // height above the point x area
      crossings += h > 0.0;
    }
    if (!(crossings & 1)) continue;
    const double tx = (xw - ox) / res, ty = (yw - oy) / res, tz = (zw - oz) / res;
    if (!(fabs(tx) < 1e9 && fabs(ty) < 1e9 && fabs(tz) < 1e9)) continue;
    const int cx = int(round(tx)), cy = int(round(ty)), cz = int(round(tz));
    if (cx < 0 || cy < 0 || cz < 0 || cx >= nx || cy >= ny || cz >= nz) continue;
    occ[(size_t(cx) * ny + cy) * nz + cz] = 1;
  }
}

// axis: 0 = x (input: occupancy u8 -> d^2 along x), 1 = y, 2 = z (inputs: u16 partial squared distances).
// out[v] = min over |k| <= cap of in[v + k along axis] + k^2, saturated at cap^2 (kInf marks "nothing within cap").
template <int kAxis, typename Out>
__global__ void k_edt_pass(int nx, int ny, int nz, int cap, const void* __restrict__ in_, Out* __restrict__ out) {
  const unsigned kInf = 0xffffu;
  const size_t cells = size_t(nx) * ny * nz;
  const unsigned cap2 = unsigned(cap) * unsigned(cap);
  for (size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < cells; v += size_t(gridDim.x) * blockDim.x) {
    const int z = int(v % nz), y = int((v / nz) % ny), x = int(v / (size_t(nz) * ny));
    const int c = kAxis == 0 ? x : (kAxis == 1 ? y : z);
    const int n = kAxis == 0 ? nx : (kAxis == 1 ? ny : nz);
    const size_t stride = kAxis == 0 ? size_t(ny) * nz : (kAxis == 1 ? size_t(nz) : 1);
    unsigned best = kInf;
    const int k0 = max(-cap, -c), k1 = min(cap, n - 1 - c);
    for (int k = k0; k <= k1; ++k) {
      const size_t u = v + (long long)k * (long long)stride;
      unsigned val;
      if (kAxis == 0) val = static_cast<const uint8_t*>(in_)[u] ? 0u : kInf;
      else val = static_cast<const uint16_t*>(in_)[u];
      if (val != kInf) best = min(best, val + unsigned(k * k));
    }
    if (kAxis == 2) out[v] = Out(min(best, cap2));
    else out[v] = Out(best > cap2 ? kInf : best);
  }
}

// ---- broad-phase field (k_cost culling) --------------------------------------------------------------------------
// occupancy of a squared-distance grid = its zero set
template <typename V>
__global__ void k_cull_occupancy(size_t cells, const V* __restrict__ vox, uint8_t* __restrict__ occ, unsigned* __restrict__ max_value) {
  unsigned mx = 0;
  for (size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < cells; v += size_t(gridDim.x) * blockDim.x) {
    const unsigned val = vox[v];
    occ[v] = val == 0 ? 1 : 0;
    mx = max(mx, val);
  }
  for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0) atomicMax(max_value, mx);
}
// the grid must BE the capped squared distance transform of its zero set (then the uncapped transform bounds it everywhere)
template <typename V>
__global__ void k_cull_verify(size_t cells, const V* __restrict__ vox, const uint16_t* __restrict__ e2, const unsigned* __restrict__ cap2_given,
                              int* __restrict__ mismatch) {
  const unsigned c2 = *cap2_given;
  for (size_t v = size_t(blockIdx.x) * blockDim.x + threadIdx.x; v < cells; v += size_t(gridDim.x) * blockDim.x)
    if (min(unsigned(e2[v]), c2) != unsigned(vox[v])) *mismatch = 1;
}
// g[coarse] = res * sqrt(min e2 over the 4^3 block), rounded down: a lower bound of the distance of every fine cell in the block
__global__ void k_cull_pool(int nx, int ny, int nz, int cnx, int cny, int cnz, double res, const uint16_t* __restrict__ e2, float* __restrict__ g) {
  const size_t total = size_t(cnx) * cny * cnz;
  for (size_t c = size_t(blockIdx.x) * blockDim.x + threadIdx.x; c < total; c += size_t(gridDim.x) * blockDim.x) {
    const int cz = int(c % cnz), cy = int((c / cnz) % cny), cx = int(c / (size_t(cnz) * cny));
    unsigned best = 0xffffu;
    for (int x = cx * 4; x < min(nx, cx * 4 + 4); ++x)
      for (int y = cy * 4; y < min(ny, cy * 4 + 4); ++y)
        for (int z = cz * 4; z < min(nz, cz * 4 + 4); ++z) best = min(best, unsigned(e2[(size_t(x) * ny + y) * nz + z]));
    float d = float(res * sqrt(double(best)));
    if (double(d) > res * sqrt(double(best))) d = nextafterf(d, 0.0f);
    g[c] = d > 0.0f ? nextafterf(d, 0.0f) : 0.0f;
  }
}

// theta += updates  (Policy::updateParameters for caller-supplied updates)
__global__ void k_axpy(size_t n, const double* __restrict__ x, double* __restrict__ y) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// ---------------------------------------------------------------------------------------------
// k_shard_stats: one launch per statistics phase of the rollout-sharded / huge-R path (config C3), grid = (column blocks of
// 128 (d,t) elements) x (rollout chunks):
//   1. every CTA reduces its chunk of rollouts for its 128 columns: phase MAX {max c, max -c}, phase SUM {sum e, sum e*eps}
//   2. the LAST chunk-CTA of a column block (atomic ticket) reduces that block over the chunks, in chunk order
//   3. the LAST column block to finish does the cross-GPU exchange over NVLink peer memory, in place:
//        store the local 2*D*N vector into slot [phase][my rank] of EVERY peer's exchange buffer (P2P stores), fence at
//        system scope, raise flag [phase][my rank] = epoch on every peer, wait for the `world` flags of the own buffer, reduce
//        the world slots in rank order (the same order on every rank: all ranks hold bit-identical statistics)
//   4. (phase SUM, fused finalize) u = (sum e*eps / sum e) .* s, the projection M u as two banded triangular solves (one lane
//      per dimension), theta += update.
// Exchange buffer of a rank: data[2 phases][world][n] doubles | flags[2 phases][world] u64, mapped into every peer (CUDA IPC).
// Epochs increase monotonically (two exchanges per iteration, alternating phase buffers), so a rank can only overwrite its slot
// of phase p after every peer has raised a later flag, i.e. finished reading the previous use of that slot.  The wait is
// bounded (~2 s of SM clocks); on expiry the sticky error flag is set, the reduction is skipped, and no later finalize applies
// an update (the host sees the flag in stomp_engine_shard_status / get_parameters / the iteration statistics).
// Round 1 ran k_*_partial, k_pair_reduce, k_peer_allreduce and k_finalize as separate launches (7 per iteration).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct ShardStatsArgs {
  int R, D, N, rollouts_per_chunk;
  int is_max;                 // 1: phase MAX, 0: phase SUM
  int do_exchange;            // peer-memory all-reduce of the result (world > 1, peers mapped)
  int do_finalize, apply;     // phase SUM: projection + update in the same launch
  const double* cumulative;   // [R][D*N]
  const double* state;        // [R][N] or nullptr.  Given (no cumulative costs): cost = state[r][t] + cumulative[r][d][t] with
                              // `cumulative` pointing at the CONTROL costs — k_cumulative's own addition, so that kernel is not needed
  const double* noise;        // [R][D*N]
  const double* minmax;       // [2][D*N] (phase SUM input)
  double* part;               // [chunks][2][D*N]
  double* out;                // [2][D*N]: minmax (phase MAX) or sums (phase SUM)
  int* counters;              // [column blocks + 1], zero between launches
  // exchange
  int rank, world;
  unsigned long long epoch;
  unsigned char* peers[kMaxPeers];
  int* err;
  // finalize
  double* updates;            // [D*N]
  double* theta;              // [D*N]
  Band band;
};

__device__ __forceinline__ void peer_exchange(const ShardStatsArgs& a, double* local, int n) {
  const int W = a.world, ph = a.is_max ? 0 : 1;
  const size_t data_bytes = size_t(2) * W * n * sizeof(double);
  for (int p = 0; p < W; ++p) {
    double* dst = reinterpret_cast<double*>(a.peers[p]) + (size_t(ph) * W + a.rank) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcg(local + i);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < W; p += blockDim.x) {
    volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers[p] + data_bytes) + ph * W + a.rank;
    *f = a.epoch;
  }
  for (int p = threadIdx.x; p < W; p += blockDim.x) {
    volatile unsigned long long* f = reinterpret_cast<unsigned long long*>(a.peers[a.rank] + data_bytes) + ph * W + p;
    const long long t0 = clock64();
    while (*f < a.epoch) {
      if (clock64() - t0 > 4000000000ll) { *a.err = 1; s_timeout = 1; break; }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (s_timeout) return;        // some rank never arrived: leave the local partial, the sticky flag stops the update
  const volatile double* src = reinterpret_cast<const double*>(a.peers[a.rank]) + size_t(ph) * W * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = src[i];
    for (int r = 1; r < W; ++r) {
      const double v = src[size_t(r) * n + i];
      acc = a.is_max ? fmax(acc, v) : acc + v;
    }
    local[i] = acc;
  }
  __threadfence();
  __syncthreads();
}

__global__ void __launch_bounds__(128) k_shard_stats(ShardStatsArgs a) {
  extern __shared__ double sm_fin[];   // finalize only: band tables [N][16] + u [D][N|1]
  __shared__ int s_last;
  const int DN = a.D * a.N, chunk = blockIdx.y, nchunks = gridDim.y;
  const int r0 = chunk * a.rollouts_per_chunk, r1 = min(a.R, r0 + a.rollouts_per_chunk);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* srow = a.state ? a.state + (i < DN ? i % a.N : 0) : nullptr;
  auto cost_of = [&](int r) -> double {
    const double c = a.cumulative[size_t(r) * DN + i];
    return srow ? srow[size_t(r) * a.N] + c : c;
  };
  if (i < DN) {
    if (a.is_max) {
      double mx = -1.0e300, mn = 1.0e300;
      for (int r = r0; r < r1; ++r) {
        const double v = cost_of(r);
        mx = fmax(mx, v);
        mn = fmin(mn, v);
      }
      a.part[(size_t(chunk) * 2 + 0) * DN + i] = mx;
      a.part[(size_t(chunk) * 2 + 1) * DN + i] = -mn;
    } else {
      const double mx = a.minmax[i], mn = -a.minmax[DN + i];
      double denom = mx - mn;
      if (denom < 1e-8) denom = 1e-8;
      const double h = -10.0 / denom;
      double se = 0.0, see = 0.0;
      for (int r = r0; r < r1; ++r) {
        const double e = exp_weight(h * (cost_of(r) - mn));
        se += e;
        see += e * a.noise[size_t(r) * DN + i];
      }
      a.part[(size_t(chunk) * 2 + 0) * DN + i] = se;
      a.part[(size_t(chunk) * 2 + 1) * DN + i] = see;
    }
  }
  // ---- last chunk-CTA of this column block: reduce over the chunks (chunk order) --------------------------------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.counters + 1 + blockIdx.x, 1) == nchunks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (i < DN) {
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      double acc = a.is_max ? -1.0e300 : 0.0;
      for (int c = 0; c < nchunks; ++c) {
        const double v = __ldcg(a.part + (size_t(c) * 2 + which) * DN + i);
        acc = a.is_max ? fmax(acc, v) : acc + v;
      }
      a.out[size_t(which) * DN + i] = acc;
    }
  }
  // ---- last column block: exchange across the GPUs, then (phase SUM) the update -----------------------------------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    a.counters[1 + blockIdx.x] = 0;
    s_last = atomicAdd(a.counters, 1) == int(gridDim.x) - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) a.counters[0] = 0;
  __threadfence();
  if (a.do_exchange) peer_exchange(a, a.out, 2 * DN);
  if (!a.do_finalize) return;
  if (a.err != nullptr && *reinterpret_cast<volatile int*>(a.err) != 0) return;   // a timed-out exchange never updates theta
  const int N = a.N, D = a.D, stride = N | 1;
  double* sfw = sm_fin;
  double* sbw = sfw + N * 8;
  double* u = sbw + N * 8;
  for (int k = threadIdx.x; k < N * 8; k += blockDim.x) sfw[k] = a.band.fw[k], sbw[k] = a.band.bw[k];
  for (int k = threadIdx.x; k < DN; k += blockDim.x) {
    const int d = k / N, t = k - d * N;
    u[d * stride + t] = (__ldcg(a.out + DN + k) / __ldcg(a.out + k)) * a.band.proj_scale[t];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    band_forward(u + d * stride, sfw, N);
    band_backward(u + d * stride, sbw, N);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < DN; k += blockDim.x) {
    const int d = k / N, t = k - d * N;
    const double v = u[d * stride + t];
    a.updates[k] = v;
    if (a.apply) a.theta[k] += v;
  }
}

__global__ void k_probabilities(int R, int DN, const double* __restrict__ cumulative, const double* __restrict__ minmax,
                                const double* __restrict__ sums, double* __restrict__ prob) {
  size_t n = size_t(R) * DN;
  for (size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += size_t(gridDim.x) * blockDim.x) {
    int i = int(k % DN);
    double mx = minmax[i], mn = -minmax[DN + i];
    double denom = mx - mn;
    if (denom < 1e-8) denom = 1e-8;
    prob[k] = exp_weight(-10.0 / denom * (cumulative[k] - mn)) / sums[i];
  }
}

__global__ void k_finalize(int D, int N, int apply, const double* __restrict__ sums, double* __restrict__ updates,
                           double* __restrict__ theta, Band band, const int* __restrict__ err) {
  extern __shared__ double smem[];
  if (err != nullptr && *err != 0) return;   // a timed-out peer exchange never updates theta
  const int d = blockIdx.x, DN = D * N;
  double* sfw = smem;
  double* sbw = sfw + N * 8;
  double* u = sbw + N * 8;
  for (int k = threadIdx.x; k < N * 8; k += blockDim.x) sfw[k] = band.fw[k], sbw[k] = band.bw[k];
  for (int t = threadIdx.x; t < N; t += blockDim.x)
    u[t] = (sums[DN + d * N + t] / sums[d * N + t]) * band.proj_scale[t];
  __syncthreads();
  if (threadIdx.x == 0) {
    band_forward(u, sfw, N);
    band_backward(u, sbw, N);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    updates[d * N + t] = u[t];
    if (apply) theta[d * N + t] += u[t];
  }
}

}  // namespace stomp_dev
