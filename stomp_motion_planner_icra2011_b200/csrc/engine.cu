// engine.cu — host side of the B200 STOMP rollout engine and its C ABI (include/stomp_b200.h).
//
// One handle = one device = one CUDA stream.  The handle owns every device buffer; the host only
// orchestrates kernel launches (kernels.cuh) and the once-per-request fp64 setup (host_math.hpp).
// There is no CPU fallback: every entry point that computes launches CUDA kernels, and a missing
// device is an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "host_math.hpp"
#include "kernels.cuh"

using namespace stomp_dev;
namespace sh = stomp_host;

namespace {

thread_local std::string g_error;

int fail(const std::string& msg) {
  g_error = msg;
  return 1;
}

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (void)cudaGetLastError();                                                                 \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(e_));                          \
    }                                                                                           \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) { p = nullptr; return e; }
    n = count;   // only a successful allocation has a size: `buf.n >= need` checks never pass on a null buffer
    e = cudaMemset(p, 0, count * sizeof(T));
    // the handle's stream is non-blocking: make sure the legacy-stream memset cannot overtake later copies
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

struct Engine {
  stomp_engine_desc desc;
  int B = 0, D = 0, N = 0, R = 0, Rre = 0, K = 0;
  int device = 0, num_sms = 148;
  cudaStream_t stream = nullptr, copy_stream = nullptr, tail_stream = nullptr;
  cudaStream_t ws = nullptr;   // stream the launch helpers currently enqueue on (stream or tail_stream)
  cudaEvent_t ev_tail = nullptr, ev_upd = nullptr, ev_selected = nullptr, ev_cost = nullptr;
  // asynchronous result read-back (two requests in flight)
  cudaStream_t result_stream = nullptr;
  cudaEvent_t ev_snap_main[2] = {nullptr, nullptr}, ev_snap_tail[2] = {nullptr, nullptr}, ev_results[2] = {nullptr, nullptr};
  int next_ticket = 0;
  bool ticket_used[2] = {false, false};
  bool overlap = true, tail_dirty = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_copy_done[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  int inject_write = 0, inject_pending_buf = 0, inject_pending_n = 0;
  int64_t launches = 0;
  bool f32 = false;

  sh::PolicyMatrices pm;
  sh::FoldedRobot robot;
  bool have_robot = false, have_sdf = false, have_problems = false;
  // broad phase of k_cost: sphere clusters + conservative coarse distance field (built lazily from robot + distance field)
  struct HostCluster { int begin, end; double c[3], rho_max, need, cap_need; bool clearance_positive; };
  std::vector<HostCluster> clusters;
  DevBuf<unsigned char> dclusters;
  DevBuf<float> cull_g;
  int cull_dims[3] = {0, 0, 0};
  bool cull_dirty = true, cull_on = false, cull_allowed = true;
  double cull_clear_fraction = 0.0;
  // inverse-dynamics (torque) cost term, stomp_engine_set_dynamics
  std::vector<stomp_segment> raw_segments;
  DevBuf<unsigned char> chain;
  DevBuf<double> torque_q, torque_tap;
  int chain_len = 0;
  double torque_weight = 0.0, gravity[3] = {0.0, 0.0, -9.8};
  std::vector<double> noise_stddev, noise_decay;
  uint64_t seed = 0x57012011ull;

  // PolicyImprovement state
  int cur = 0;  // ping-pong index of params / state costs
  bool reused_next = false, extra_added = false;
  int num_gen = 0;
  uint32_t generation = 0;  // rollout generations since set_problems: the Philox "iteration" counter
  bool injected_pending = false;
  double control_cost_weight = 0.0;

  // device buffers
  DevBuf<double> theta, pad_start, pad_goal;
  DevBuf<double> params[2], state[2];
  DevBuf<double> noise, control[2], cumulative, totals;   // control[cur]: ping-pong like params / state (k_cumulative of iteration i
                                                          // reads its control costs while iteration i + 1 already writes new ones)
  DevBuf<double> noise_projected, probabilities, clipped;      // taps (keep_intermediates)
  DevBuf<double> extra_state, extra_control, updates, noiseless_sum;
  DevBuf<double> eps_in2[2];   // injected noise, double buffered (async uploads overlap the previous iteration)
  DevBuf<int> reuse_src, collision_free, constraints_ok, scratch_cflags;
  DevBuf<unsigned char> dconstraints;
  std::vector<stomp_orientation_constraint> constraints;
  double constraint_cost_weight = 0.0;
  DevBuf<double> band_fw, band_bw, proj_scale, qinv_t, noise_scale;
  DevBuf<double> dense_cinv, dense_ms;   // [N][N] C^-1 and R^-1 diag(s) for k_generate_dense (small batches)
  DevBuf<double> mma_a1, mma_a2;         // the same two, zero-padded to [Np][Np] (Np = N rounded up to 8): k_generate_mma's A operands
  bool dense_update = true;   // A/B switch (STOMP_NO_DENSE_UPDATE=1): k_update projects with the banded solves
  bool split_cost = true;     // A/B switch (STOMP_NO_SPLIT_COST=1): small batches take the one-warp-per-tile k_cost like large ones
  // look-ahead generation: the next iteration's noise and M * noise are produced on pre_stream while this iteration runs
  bool lookahead = true;      // A/B switch (STOMP_NO_LOOKAHEAD=1)
  cudaStream_t pre_stream = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_fin = nullptr;
  DevBuf<double> pre_noise, pre_y, pre_scale, gen_scratch3;
  bool pre_valid = false, pre_dirty = false;
  uint32_t pre_generation = 0;
  int pre_iteration = 0, pre_num_gen = 0;
  uint64_t pre_epoch = 0;
  // candidate pass (small batches): every previous rollout prepared as a reused slot on cand_stream, then k_select_gather
  cudaStream_t cand_stream = nullptr;
  cudaEvent_t ev_cand = nullptr;
  DevBuf<double> cand_noise, cand_control, cand_params, cand_y, gen_scratch4;
  DevBuf<int> cand_map;
  std::vector<double> scale_host, pre_scale_host;      // noise scales of this / the look-ahead iteration (launch arguments)
  std::map<const void*, size_t> smem_optin;
  int update_dpc = 0;              // A/B (STOMP_UPDATE_DPC)
  bool scales_in_args = false;     // A/B switch (STOMP_SCALES_IN_ARGS=1): measured slower, see set_noise_scale
  bool cand_valid = false, cand_dirty = false;
  // small batches: k_extra_total (and optimize's k_track_best) run on cand_stream beside the tail stream's chain
  cudaEvent_t ev_nl = nullptr, ev_extra = nullptr;
  bool extra_dirty = false;        // ev_extra was recorded: the noise-less rollout's totals / bookkeeping may still be in flight
  bool extra_offchain = false;     // ... in the iteration just enqueued
  bool offchain_extra = true;      // A/B switch (STOMP_NO_OFFCHAIN_EXTRA=1)
  uint32_t cand_generation = 0;
  uint64_t cand_epoch = 0;
  DevBuf<double> seg_hf, seg_hb;   // spike tables of k_generate_seg for seg_P segments
  int seg_P = 0;
  bool wide_update = true;    // A/B switch (STOMP_NO_WIDE_UPDATE=1): k_update with 512-thread CTAs when there are few of them
  bool cumulative_stale = false;   // the last k_cumulative skipped the cumulative-cost array
  bool totals_stale = false;       // ... or was not launched at all (huge path without reuse): Rollout::getCost() on demand
  DevBuf<double> csum;             // [B][R][D] per-vector control-cost sums from k_generate (GenArgs::csum)
  bool csum_new = false, csum_reused = false;   // ... valid for this iteration's new / reused slots
  bool use_totals_kernel = true;   // A/B switch (STOMP_NO_TOTALS_KERNEL=1)
  int cum_placement = 0;      // A/B (STOMP_CUM_PLACEMENT=chain|late): where a large batch's k_cumulative runs when k_update is direct
  bool direct_update = true;  // A/B switch (STOMP_NO_DIRECT_UPDATE=1): k_update always reads k_cumulative's output
  bool dmma_update = true;    // A/B switch (STOMP_NO_DMMA=1): the dense projection runs as scalar DFMAs instead of DMMA tiles
  int gen_mode = 0;   // 0: pick by batch shape; 1 k_generate, 2 k_generate_dense, 3 k_generate_mma, 4 k_generate_seg: always that one (A/B)
  DevBuf<double> limit_min, limit_max;
  DevBuf<int> has_limits;
  DevBuf<unsigned char> nodes, spheres, sqrt_table, vox, vox_brick;
  DevBuf<double> scratch_params, scratch_noise, scratch_costs;  // execute / compute_control_costs staging
  DevBuf<int> scratch_flags;
  DevBuf<stomp_sphere_debug> debug;
  DevBuf<double> part, minmax, sums;  // sharded / huge-R statistics
  // peer-memory exchange of the sharded statistics (k_shard_stats): this rank's buffer, the peers' mappings of theirs
  DevBuf<unsigned char> xchg;
  DevBuf<int> xchg_err, shard_counters;
  unsigned char* peer_base[kMaxPeers] = {nullptr};
  bool peers_open = false;
  unsigned long long xchg_epoch = 0;
  DevBuf<double> snap_theta[2], snap_cost[2];
  DevBuf<int> snap_flag[2];
  DevBuf<double> gen_scratch, gen_scratch2;   // time-major work buffers of k_generate (main / tail stream)
  DevBuf<double> extra_clipped, best_traj, best_cost, cost_log;   // optimize() bookkeeping
  DevBuf<unsigned char> track_state;
  DevBuf<int> num_done;
  int num_nodes = 0;
  Sdf sdf;
  size_t scratch_n = 0;

  // CUDA-graph replay of stomp_engine_run (steady-state iterations, engine noise): kGraphIters iterations of the two-stream
  // schedule captured once; the per-iteration host work (Philox generation counter, noise scales) moves to k_advance_iteration
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};   // one per parity of the params / state-cost ping-pong at entry
  uint64_t graph_key[2] = {0, 0}, config_epoch = 1;   // config_epoch: bumped by every setter that changes what the launches carry
  int64_t graph_launches = 0;                  // kernel launches inside one replay
  int graph_mode = 0;                          // 0 off (default: replay measured slower, profiles/README.md), 1 on (STOMP_GRAPH=1 / stomp_engine_set_graph_mode)
  bool capturing = false;
  int64_t steady_iterations = 0;               // iterations run in the steady state since set_problems (buffers are sized)
  DevBuf<uint32_t> dev_generation;
  DevBuf<int> dev_table_index;
  DevBuf<double> scale_table;
  // optional per-kernel timing with CUDA events on the launching stream (bench.py roofline)
  // STOMP_CHAIN_PROBE=1: two timing events per overlapped iteration (end of the main stream's k_cumulative, end of the tail
  // stream's chain) and, at destroy, the mean of "tail end - main end" on stderr: which chain k_update really waits for,
  // with far less perturbation than the per-launch events of the timeline
  bool chain_probe = false;
  cudaEvent_t probe_main = nullptr, probe_tail = nullptr, probe_upd = nullptr;
  double probe_sum_us = 0.0, probe_upd_sum_us = 0.0;
  long long probe_n = 0;
  bool prof_on = false, prof_timeline = false;   // timeline: keep the two-stream schedule while recording
  std::vector<int> prof_stream;                  // 0 main, 1 tail, 2 other
  std::vector<cudaEvent_t> prof_a, prof_b;
  std::vector<std::string> prof_name;
  size_t prof_used = 0;

  Band band_view() const { return Band{band_fw.p, band_bw.p, proj_scale.p}; }
  Stencil stencil() const {
    Stencil st;
    for (int k = 0; k < 3; ++k) {
      st.weight[k] = desc.derivative_costs[k];
      for (int j = 0; j < 7; ++j) st.coef[k][j] = pm.stencil_scale[k] * sh::kDiffRules[k][j];
    }
    return st;
  }
  // statistics over rollouts: one CTA per (problem, dims) looping over R (small R, many problems), or the
  // two-stage partial reductions over rollout chunks (rollout sharding, or one problem with many rollouts)
  // k_update forms S + C itself and k_cumulative (the totals of the next reuse selection, the cumulative-cost tap) leaves the
  // critical path: C1 0.071 -> 0.065 ms.  Where it runs instead depends on the batch: beside k_update for small batches; for a
  // machine-filling batch that costs more than it saves (C2 0.465 -> 0.500 ms: 10 240 small CTAs crowd k_update's single wave),
  // so there it follows k_update on the tail stream, under the next iteration's k_generate (8 warps / SM)
  bool direct_now() const { return direct_update && !desc.use_cumulative_costs; }
  // Small batches run a latency schedule (direct k_update, look-ahead generation, candidate pass + k_select_gather): their
  // iteration is a chain of short launches and the machine is mostly idle.  With a machine-filling batch the same work only
  // competes for the SMs (C2: look-ahead 0.476 -> 0.577 ms), so the throughput schedule stays.
  // (C1-shaped problems on a B200, us per iteration, throughput -> latency schedule: B = 1: 65.5 -> 52.2, 4: 67.1 -> 50.5,
  //  16: 70.6 -> 68.2, 48: 148 -> 109, 128: 179 -> 198, 512: 305 -> 352; the limit sits between 48 and 128 problems)
  bool small_batch() const { return (long long)B * R * D * N <= small_max; }
  long long small_max = 1 << 19;
  bool huge_path() const { return desc.rollout_shard_world > 1 || R > 4096 || (B == 1 && R >= 128); }
};

void begin_launch(Engine& e) {
  if (!e.prof_on) return;
  if (e.prof_used == e.prof_a.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    e.prof_a.push_back(a);
    e.prof_b.push_back(b);
    e.prof_name.push_back("");
    e.prof_stream.push_back(0);
  }
  cudaEventRecord(e.prof_a[e.prof_used], e.ws);
}

int check_launch(Engine& e, const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(err));
  e.launches++;
  if (e.prof_on) {
    cudaEventRecord(e.prof_b[e.prof_used], e.ws);
    e.prof_name[e.prof_used] = what;
    e.prof_stream[e.prof_used] = e.ws == e.stream ? 0 : (e.ws == e.tail_stream ? 1 : 2);
    e.prof_used++;
  }
  return 0;
}

// main stream waits for everything enqueued on the tail stream (device-side, no host sync)
int join_streams(Engine& e) {
  e.ws = e.stream;
  if (e.pre_dirty) {
    CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_pre, 0));
    e.pre_dirty = false;
  }
  if (e.cand_dirty) {
    CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_cand, 0));
    e.cand_dirty = false;
  }
  if (e.extra_dirty) {
    CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_extra, 0));
    e.extra_dirty = false;
  }
  if (!e.tail_dirty) return 0;
  CUDA_TRY(cudaEventRecord(e.ev_tail, e.tail_stream));
  CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_tail, 0));
  e.tail_dirty = false;
  return 0;
}

template <typename T>
int upload(Engine& e, DevBuf<T>& buf, const T* host, size_t count) {
  if (buf.n != count) CUDA_TRY(buf.alloc(count));
  if (count) CUDA_TRY(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, e.stream));
  return 0;
}

template <typename Real>
int upload_robot_tables(Engine& e) {
  std::vector<DevNode<Real>> nodes(e.robot.nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) {
    const sh::HostNode& h = e.robot.nodes[i];
    DevNode<Real>& n = nodes[i];
    std::memset(&n, 0, sizeof(n));
    for (int k = 0; k < 9; ++k) n.A0[k] = Real(h.A0[k]), n.A1[k] = Real(h.A1[k]), n.A2[k] = Real(h.A2[k]);
    for (int k = 0; k < 3; ++k) n.p[k] = Real(h.p[k]), n.ax[k] = Real(h.ax[k]);
    n.parent = h.parent; n.type = h.type; n.q_index = h.q_index; n.save_slot = h.save_slot; n.load_slot = h.load_slot;
    n.sphere_begin = h.sphere_begin; n.sphere_end = h.sphere_end;
  }
  // sphere clusters: runs of <= 6 consecutive spheres of a node (consecutive spheres of a link lie next to each other)
  constexpr int kClusterSize = 6;   // A/B on B200 (C2): 4 -> 0.309, 6 -> 0.299, 9 / 14 / 30 -> 0.311 ms
  e.clusters.clear();
  for (size_t i = 0; i < nodes.size(); ++i) {
    nodes[i].cluster_begin = int(e.clusters.size());
    for (int b0 = e.robot.nodes[i].sphere_begin; b0 < e.robot.nodes[i].sphere_end; b0 += kClusterSize) {
      Engine::HostCluster cl;
      cl.begin = b0;
      cl.end = std::min(b0 + kClusterSize, e.robot.nodes[i].sphere_end);
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int j = cl.begin; j < cl.end; ++j)
        for (int k = 0; k < 3; ++k) lo[k] = std::min(lo[k], e.robot.spheres[j].pos[k]), hi[k] = std::max(hi[k], e.robot.spheres[j].pos[k]);
      for (int k = 0; k < 3; ++k) cl.c[k] = 0.5 * (lo[k] + hi[k]);
      cl.rho_max = cl.need = cl.cap_need = 0.0;
      cl.clearance_positive = true;
      for (int j = cl.begin; j < cl.end; ++j) {
        const sh::HostSphere& hs = e.robot.spheres[j];
        const double rho = std::sqrt((hs.pos[0] - cl.c[0]) * (hs.pos[0] - cl.c[0]) + (hs.pos[1] - cl.c[1]) * (hs.pos[1] - cl.c[1]) +
                                     (hs.pos[2] - cl.c[2]) * (hs.pos[2] - cl.c[2]));
        cl.rho_max = std::max(cl.rho_max, rho);
        cl.need = std::max(cl.need, rho + hs.radius + hs.clearance);
        cl.cap_need = std::max(cl.cap_need, hs.radius + hs.clearance);
        if (!(hs.clearance > 0.0)) cl.clearance_positive = false;
      }
      e.clusters.push_back(cl);
    }
    nodes[i].cluster_end = int(e.clusters.size());
  }
  e.cull_dirty = true;
  e.cull_on = false;
  {
    std::vector<DevCluster<Real>> dc(std::max<size_t>(1, e.clusters.size()));
    std::memset(dc.data(), 0, dc.size() * sizeof(DevCluster<Real>));
    for (size_t i = 0; i < e.clusters.size(); ++i) {
      for (int k = 0; k < 3; ++k) dc[i].c[k] = Real(e.clusters[i].c[k]);
      dc[i].thr = std::numeric_limits<Real>::infinity();
      dc[i].margin = Real(0);
      dc[i].begin = e.clusters[i].begin; dc[i].end = e.clusters[i].end;
    }
    if (upload(e, e.dclusters, reinterpret_cast<const unsigned char*>(dc.data()), dc.size() * sizeof(DevCluster<Real>))) return 1;
  }
  std::vector<DevSphere<Real>> sph(e.robot.spheres.size());
  for (size_t i = 0; i < sph.size(); ++i) {
    const sh::HostSphere& h = e.robot.spheres[i];
    std::memset(&sph[i], 0, sizeof(sph[i]));
    for (int k = 0; k < 3; ++k) sph[i].pos[k] = Real(h.pos[k]);
    sph[i].radius = Real(h.radius); sph[i].clearance = Real(h.clearance); sph[i].inv_clearance = Real(h.inv_clearance);
    sph[i].weight = Real(h.weight); sph[i].original_index = h.original_index;
  }
  if (upload(e, e.nodes, reinterpret_cast<const unsigned char*>(nodes.data()), nodes.size() * sizeof(DevNode<Real>))) return 1;
  if (upload(e, e.spheres, reinterpret_cast<const unsigned char*>(sph.data()), sph.size() * sizeof(DevSphere<Real>))) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

template <typename Real>
int upload_sqrt_table(Engine& e) {
  std::vector<Real> tab(256);
  for (int i = 0; i < 256; ++i) tab[i] = Real(std::sqrt(double(i)) * e.sdf.res);  // PropagationDistanceField sqrt_table_
  if (upload(e, e.sqrt_table, reinterpret_cast<const unsigned char*>(tab.data()), tab.size() * sizeof(Real))) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

// ---- broad phase of the collision cost --------------------------------------------------------------------------------
// Valid only when the distance field IS the capped squared distance transform of its zero set (checked on the device): then
// the uncapped transform E of that zero set bounds every lookup.  For a sphere j of a cluster with centre c:
//   cell(x_j) is within |p_j - c| + sqrt(3) res of cell(x_c) (rigid link, two nearest-cell roundings), the coarse value
//   looked up is the minimum of E over a 4^3 block that contains cell(x_c) or a neighbour of it (one more sqrt(3) res), and
//   E is 1-Lipschitz between cell centres, so   E(cell(x_j)) res >= G - |p_j - c| - 2 sqrt(3) res.
// G >= thr = max_j(|p_j - c| + r_j + clearance_j) + (2 sqrt(3) + 1) res therefore gives distance - r_j >= clearance_j for
// every sphere: potential exactly 0, not in collision — provided the field's own cap does not bite first
// (sqrt(cap^2) res > r_j + clearance_j, else the cluster is never culled) and all those cells are interior (margin).
template <typename Real>
int upload_cluster_thresholds(Engine& e, double cap_distance) {
  std::vector<DevCluster<Real>> dc(std::max<size_t>(1, e.clusters.size()));
  std::memset(dc.data(), 0, dc.size() * sizeof(DevCluster<Real>));
  const double res = e.sdf.res, slack = (2.0 * std::sqrt(3.0) + 1.0) * res + 1e-9;
  for (size_t i = 0; i < e.clusters.size(); ++i) {
    const Engine::HostCluster& cl = e.clusters[i];
    for (int k = 0; k < 3; ++k) dc[i].c[k] = Real(cl.c[k]);
    const bool usable = e.cull_on && cl.clearance_positive && cap_distance - cl.cap_need > 1e-9;
    double thr = usable ? cl.need + slack : std::numeric_limits<double>::infinity();
    Real t = Real(thr);
    if (double(t) < thr) t = std::nextafter(t, std::numeric_limits<Real>::infinity());
    dc[i].thr = t;
    dc[i].margin = Real(std::ceil(cl.rho_max / res) + 4.0);
    dc[i].begin = cl.begin; dc[i].end = cl.end;
  }
  return upload(e, e.dclusters, reinterpret_cast<const unsigned char*>(dc.data()), dc.size() * sizeof(DevCluster<Real>));
}

template <typename V>
int build_cull_field(Engine& e, int capcull, bool* ok, double* cap_distance) {
  const int nx = e.sdf.nx, ny = e.sdf.ny, nz = e.sdf.nz;
  const size_t cells = size_t(nx) * ny * nz;
  DevBuf<uint8_t> occ;
  DevBuf<uint16_t> g1, g2;
  DevBuf<unsigned> mx;
  DevBuf<int> mismatch;
  CUDA_TRY(occ.alloc(cells)); CUDA_TRY(g1.alloc(cells)); CUDA_TRY(g2.alloc(cells)); CUDA_TRY(mx.alloc(1)); CUDA_TRY(mismatch.alloc(1));
  const unsigned egrid = unsigned(std::min<size_t>((cells + 255) / 256, size_t(148) * 64));
  const V* vox = static_cast<const V*>(e.sdf.vox);
  k_cull_occupancy<V><<<egrid, 256, 0, e.stream>>>(cells, vox, occ.p, mx.p);
  unsigned cap2_given = 0;
  CUDA_TRY(cudaMemcpyAsync(&cap2_given, mx.p, sizeof(unsigned), cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  capcull = std::max(capcull, int(std::ceil(std::sqrt(double(cap2_given)))) + 1);
  if (capcull > 255) { *ok = false; return 0; }
  k_edt_pass<0, uint16_t><<<egrid, 256, 0, e.stream>>>(nx, ny, nz, capcull, occ.p, g1.p);
  k_edt_pass<1, uint16_t><<<egrid, 256, 0, e.stream>>>(nx, ny, nz, capcull, g1.p, g2.p);
  k_edt_pass<2, uint16_t><<<egrid, 256, 0, e.stream>>>(nx, ny, nz, capcull, g2.p, g1.p);
  k_cull_verify<V><<<egrid, 256, 0, e.stream>>>(cells, vox, g1.p, mx.p, mismatch.p);
  const int cnx = (nx + 3) / 4, cny = (ny + 3) / 4, cnz = (nz + 3) / 4;
  CUDA_TRY(e.cull_g.alloc(size_t(cnx) * cny * cnz));
  k_cull_pool<<<unsigned((size_t(cnx) * cny * cnz + 127) / 128), 128, 0, e.stream>>>(nx, ny, nz, cnx, cny, cnz, e.sdf.res, g1.p, e.cull_g.p);
  int bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&bad, mismatch.p, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  CUDA_TRY(cudaGetLastError());
  e.launches += 6;
  e.cull_dims[0] = cnx; e.cull_dims[1] = cny; e.cull_dims[2] = cnz;
  *ok = bad == 0 && cap2_given > 0;
  *cap_distance = std::sqrt(double(cap2_given)) * e.sdf.res;
  return 0;
}

int ensure_cull(Engine& e) {
  if (!e.cull_dirty) return 0;
  e.cull_dirty = false;
  e.cull_on = false;
  double cap_distance = 0.0;
  const bool candidate = e.cull_allowed && e.have_robot && e.have_sdf && e.desc.sdf_mode == STOMP_SDF_NEAREST &&
                         (e.sdf.dtype == STOMP_VOXEL_U8_SQ || e.sdf.dtype == STOMP_VOXEL_U16_SQ) && !e.clusters.empty();
  if (candidate) {
    CUDA_TRY(cudaStreamSynchronize(e.stream));
    CUDA_TRY(cudaStreamSynchronize(e.tail_stream));
    double need = 0.0;
    for (const Engine::HostCluster& cl : e.clusters) need = std::max(need, cl.need);
    const int capcull = int(std::ceil((need + (2.0 * std::sqrt(3.0) + 1.0) * e.sdf.res) / e.sdf.res)) + 2;
    bool ok = false;
    if (e.sdf.dtype == STOMP_VOXEL_U8_SQ ? build_cull_field<uint8_t>(e, capcull, &ok, &cap_distance)
                                         : build_cull_field<uint16_t>(e, capcull, &ok, &cap_distance))
      return 1;
    e.cull_on = ok;
    if (ok) {
      // the broad phase only pays where most of the workspace is clear of obstacles (C2: 80 % of the coarse cells clear for the
      // median cluster, k_cost -5 %; dense clutter C4: 27 %, k_cost +6 % with it): decide from the field itself
      std::vector<double> thr;
      const double slack = (2.0 * std::sqrt(3.0) + 1.0) * e.sdf.res + 1e-9;
      for (const Engine::HostCluster& cl : e.clusters)
        if (cl.clearance_positive && cap_distance - cl.cap_need > 1e-9) thr.push_back(cl.need + slack);
      if (thr.empty()) {
        e.cull_on = false;
      } else {
        std::nth_element(thr.begin(), thr.begin() + thr.size() / 2, thr.end());
        const double median = thr[thr.size() / 2];
        std::vector<float> g(e.cull_g.n);
        CUDA_TRY(cudaMemcpy(g.data(), e.cull_g.p, g.size() * sizeof(float), cudaMemcpyDeviceToHost));
        size_t clear = 0;
        for (float v : g) clear += double(v) >= median;
        e.cull_clear_fraction = g.empty() ? 0.0 : double(clear) / double(g.size());
        e.cull_on = e.cull_clear_fraction >= 0.5;
      }
    }
  }
  if (e.have_robot) {
    if (e.f32 ? upload_cluster_thresholds<float>(e, cap_distance) : upload_cluster_thresholds<double>(e, cap_distance)) return 1;
    CUDA_TRY(cudaStreamSynchronize(e.stream));
  }
  return 0;
}

// ---- kernel launch helpers ----------------------------------------------------------------------

// dynamic shared memory opt-in of a kernel, once per size (cudaFuncSetAttribute costs a microsecond per call)
template <typename K>
int ensure_smem(Engine& e, K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = e.smem_optin.find(key);
  if (it != e.smem_optin.end() && it->second >= bytes) return 0;
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
  e.smem_optin[key] = bytes;
  return 0;
}

int launch_generate(Engine& e, GenArgs a) {
  const int N = e.N, tpb = 128;
  if (a.r_count == 0 && !a.extra) { a.r_begin = 0; a.r_count = a.R; }
  const long long nvec = (long long)a.B * (a.extra ? 1 : a.r_count) * a.D;
  if (nvec == 0) return 0;
  const unsigned grid = unsigned((nvec + tpb - 1) / tpb);
  const size_t stride = size_t(grid) * tpb;
  DevBuf<double>& scratch = e.ws == e.tail_stream ? e.gen_scratch2
                            : (e.ws == e.pre_stream ? e.gen_scratch3 : (e.ws == e.cand_stream ? e.gen_scratch4 : e.gen_scratch));
  if (scratch.n < stride * N) CUDA_TRY(scratch.alloc(stride * N));
  a.scratch = scratch.p;
  a.scratch_stride = stride;
  const size_t smem = (size_t(N) * 17 + size_t(tpb / 32) * 2 * 32 * kTileLd) * 8;
  if (smem > 220 * 1024) return fail("num_time_steps too large for the band tables of k_generate");
  if (smem > 48 * 1024) {
    CUDA_TRY(cudaFuncSetAttribute(k_generate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    CUDA_TRY(cudaFuncSetAttribute(k_generate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  }
  begin_launch(e);
  if (a.pre) k_generate<true><<<grid, tpb, smem, e.ws>>>(a);
  else k_generate<false><<<grid, tpb, smem, e.ws>>>(a);
  return check_launch(e, a.pre ? "k_generate_ahead" : "k_generate");
}

GenArgs base_gen_args(Engine& e) {
  GenArgs a;
  std::memset(&a, 0, sizeof(a));
  a.B = e.B; a.R = e.R; a.D = e.D; a.N = e.N;
  a.R_gen = e.num_gen;
  a.seed = e.seed;
  a.rollout_id_offset = int64_t(e.desc.rollout_shard_rank) * e.R;
  a.rollouts_global = int64_t(e.desc.rollout_shard_world) * e.R;
  a.theta = e.theta.p; a.pad_start = e.pad_start.p; a.pad_goal = e.pad_goal.p;
  a.noise_scale = e.noise_scale.p;
  if (e.scales_in_args && e.D <= kMaxScaleArgs && !e.capturing && int(e.scale_host.size()) == e.D) {
    a.scale_by_value = 1;
    for (int d = 0; d < e.D; ++d) a.scale_v[d] = e.scale_host[d];
  }
  a.eps_in = e.eps_in2[e.inject_pending_buf].p;
  a.params_prev = e.params[1 - e.cur].p;
  a.prev_stride = e.R;
  a.reuse_src = e.reuse_src.p;
  a.noise = e.noise.p;
  a.params = e.params[e.cur].p;
  a.noise_projected = e.noise_projected.p;
  a.control = e.control[e.cur].p;
  a.band = e.band_view();
  a.st = e.stencil();
  a.iteration_ptr = e.capturing ? e.dev_generation.p : nullptr;
  return a;
}

template <typename Real, bool kDebug, int kVox, bool kCons, bool kTri, bool kCull>
int launch_cost_k(Engine& e, CostArgs<Real>& a, int num_problems) {
  a.total_rollouts = num_problems * a.n_rollouts;
  a.params_16B = ((e.D * e.N) % 2 == 0) && (reinterpret_cast<uintptr_t>(a.params) % 16 == 0) && (a.params_problem_stride % 2 == 0) &&
                 (a.params_rollout_stride % 2 == 0);
  // lane packing: concatenate the timelines of `pack` rollouts per CTA when that needs fewer 29-step warp tiles per rollout
  const int seg = e.N + 3;
  auto tiles_for = [&](int p) { return (p * seg - 3 + kTileSteps - 1) / kTileSteps; };
  auto smem_for = [&](int p, int warps) {
    return ((size_t(p) * (size_t(e.D) * e.N + 2 * e.D) * 8 + 15) & ~size_t(15)) + size_t(e.num_nodes) * sizeof(DevNode<Real>) +
           size_t(e.K) * sizeof(DevSphere<Real>) + 256 * sizeof(Real) +
           e.constraints.size() * sizeof(DevConstraint<Real>) + e.clusters.size() * sizeof(DevCluster<Real>) + 16 /* mbarrier */;
  };
  const int max_warps = kCostMaxThreads / 32;
  int pack = 1;
  if (!kDebug)
    for (int p = 2; p <= 4 && p <= a.total_rollouts; ++p)
      if (tiles_for(p) * pack < tiles_for(pack) * p && smem_for(p, std::min(tiles_for(p), max_warps)) <= 56 * 1024) pack = p;
  // small batches: kSplitGroups warps per tile (kernels.cuh, kSplit) when a few rollouts cannot fill the machine anyway
  if (!kDebug && !kCons && !kTri && e.split_cost && a.total_rollouts <= 2 * e.num_sms && e.K >= 2 * kSplitGroups) {
    const int t1 = tiles_for(1);
    const size_t smem1 = smem_for(1, t1) + size_t(t1) * e.K * 32 * sizeof(Real);
    if (t1 * kSplitGroups * 32 <= kSplitMaxThreads && smem1 <= 200 * 1024) {
      auto ks = k_cost<Real, false, kVox, false, false, kCull, true>;
      if (ensure_smem(e, ks, smem1)) return 1;
      a.pack = 1;
      a.tiles_per_job = t1;
      begin_launch(e);
      ks<<<a.total_rollouts, t1 * kSplitGroups * 32, smem1, e.ws>>>(a);
      return check_launch(e, "k_cost");
    }
  }
  const int tiles = tiles_for(pack);
  int warps = std::min(tiles, max_warps);
  warps = std::max(warps, std::min(pack * e.D, 4));  // joint-limit pass likes a few warps
  const size_t smem = smem_for(pack, warps);
  if (smem > 220 * 1024) return fail("trajectory + robot tables exceed shared memory");
  auto kern = k_cost<Real, kDebug, kVox, kCons, kTri, kCull>;
  if (ensure_smem(e, kern, smem)) return 1;
  a.pack = pack;
  a.tiles_per_job = tiles;
  const int njobs = (a.total_rollouts + pack - 1) / pack;
  // one CTA per job: the hardware block scheduler balances rollouts whose joint-limit projection takes longer (a persistent
  // grid-stride schedule was 15 % slower, profiles/README.md)
  begin_launch(e);
  kern<<<njobs, warps * 32, smem, e.ws>>>(a);
  return check_launch(e, "k_cost");
}

// nearest-cell lookup (the reference's) or the trilinear extension, chosen at create time
template <typename Real, bool kDebug, int kVox, bool kCons>
int launch_cost_c(Engine& e, CostArgs<Real>& a, int num_problems) {
  if (e.desc.sdf_mode == STOMP_SDF_TRILINEAR) return launch_cost_k<Real, kDebug, kVox, kCons, true, false>(e, a, num_problems);
  if (!kDebug && kVox != STOMP_VOXEL_F32 && e.cull_on) return launch_cost_k<Real, kDebug, kVox, kCons, false, true>(e, a, num_problems);
  return launch_cost_k<Real, kDebug, kVox, kCons, false, false>(e, a, num_problems);
}

// the constraint evaluators are compiled out of the common (no path constraints) instantiation; the debug tap always
// uses the general one
template <typename Real, bool kDebug, int kVox>
int launch_cost_v(Engine& e, CostArgs<Real>& a, int num_problems) {
  if (kDebug || !e.constraints.empty()) return launch_cost_c<Real, kDebug, kVox, true>(e, a, num_problems);
  return launch_cost_c<Real, kDebug, kVox, kDebug>(e, a, num_problems);
}

template <typename Real, bool kDebug>
int launch_cost_t(Engine& e, CostArgs<Real>& a, int num_problems) {
  switch (e.sdf.dtype) {
    case STOMP_VOXEL_U8_SQ: return launch_cost_v<Real, kDebug, STOMP_VOXEL_U8_SQ>(e, a, num_problems);
    case STOMP_VOXEL_U16_SQ: return launch_cost_v<Real, kDebug, STOMP_VOXEL_U16_SQ>(e, a, num_problems);
    default: return launch_cost_v<Real, kDebug, STOMP_VOXEL_F32>(e, a, num_problems);
  }
}

template <typename Real>
CostArgs<Real> base_cost_args(Engine& e) {
  CostArgs<Real> a;
  std::memset(&a, 0, sizeof(a));
  a.D = e.D; a.N = e.N; a.K = e.K; a.num_nodes = e.num_nodes;
  a.pad_start = e.pad_start.p; a.pad_goal = e.pad_goal.p;
  a.qinv_t = e.qinv_t.p;
  a.has_limits = e.has_limits.p; a.limit_min = e.limit_min.p; a.limit_max = e.limit_max.p;
  a.nodes = reinterpret_cast<const DevNode<Real>*>(e.nodes.p);
  a.spheres = reinterpret_cast<const DevSphere<Real>*>(e.spheres.p);
  a.sqrt_table = reinterpret_cast<const Real*>(e.sqrt_table.p);
  a.sdf = e.sdf;
  a.inv_time = 1.0 / e.desc.discretization;
  a.obstacle_weight = e.desc.obstacle_cost_weight;
  a.constraint_weight = e.constraint_cost_weight;
  a.num_constraints = int(e.constraints.size());
  a.constraints = reinterpret_cast<const DevConstraint<Real>*>(e.dconstraints.p);
  a.clusters = reinterpret_cast<const DevCluster<Real>*>(e.dclusters.p);
  a.num_clusters = int(e.clusters.size());
  a.cull.g = e.cull_g.p; a.cull.nx = e.cull_dims[0]; a.cull.ny = e.cull_dims[1]; a.cull.nz = e.cull_dims[2]; a.cull.enabled = e.cull_on ? 1 : 0;
  a.g_ox = Real(e.sdf.origin[0]); a.g_oy = Real(e.sdf.origin[1]); a.g_oz = Real(e.sdf.origin[2]);
  a.g_res = Real(e.sdf.res); a.g_inv_res = Real(e.sdf.inv_res);
  a.g_nox = Real(-e.sdf.origin[0] * e.sdf.inv_res); a.g_noy = Real(-e.sdf.origin[1] * e.sdf.inv_res);
  a.g_noz = Real(-e.sdf.origin[2] * e.sdf.inv_res);
  // finite-difference velocity rule {-2, -3, 6, -1} / 6 over t-1 .. t+2, divided by dt (stomp_utils.h:49-56)
  a.c_m1 = Real(a.inv_time * (-2.0 / 6.0)); a.c_0 = Real(a.inv_time * (-3.0 / 6.0)); a.c_p1 = Real(a.inv_time * (6.0 / 6.0));
  a.c_p2 = Real(a.inv_time * (-1.0 / 6.0));
  a.lim_x = unsigned(e.sdf.nx - 2); a.lim_y = unsigned(e.sdf.ny - 2); a.lim_z = unsigned(e.sdf.nz - 2);
  return a;
}

// torque term (weight > 1e-9 only): costs += w * sum_j |tau_j| from the joint-limit-projected rollouts k_cost just wrote
int launch_torque(Engine& e, const double* clipped, size_t pstride, int n_rollouts, int num_problems, double* costs, size_t cstride) {
  TorqueArgs a;
  std::memset(&a, 0, sizeof(a));
  a.D = e.D; a.N = e.N; a.n_rollouts = n_rollouts; a.total_rollouts = num_problems * n_rollouts; a.ns = e.chain_len;
  a.q_problem_stride = pstride; a.q_rollout_stride = size_t(e.D) * e.N; a.cost_problem_stride = cstride;
  a.q = clipped; a.pad_start = e.pad_start.p; a.pad_goal = e.pad_goal.p;
  a.chain = reinterpret_cast<const DevChainLink*>(e.chain.p);
  a.inv_time = 1.0 / e.desc.discretization;
  a.inv_time2 = 1.0 / (e.desc.discretization * e.desc.discretization);
  for (int k = 0; k < 3; ++k) a.g[k] = e.gravity[k];
  a.weight = e.torque_weight;
  a.costs = costs;
  a.torques = e.torque_tap.n >= size_t(a.total_rollouts) * e.N * e.D ? e.torque_tap.p : nullptr;
  const long long total = (long long)a.total_rollouts * e.N;
  begin_launch(e);
  k_torque<<<unsigned((total + 127) / 128), 128, 0, e.ws>>>(a);
  return check_launch(e, "k_torque");
}

int launch_cost_only(Engine& e, const double* params, size_t pstride, int n_rollouts, int num_problems, int include_pads,
                     double* costs, size_t cstride, int* flags, int flag_stride, int flag_offset, double* clipped,
                     stomp_sphere_debug* debug, int* cflags);

// cost plugin over rollouts stored as params[b*pstride + r*D*N], writing costs[b*cstride + r*N]
int launch_cost(Engine& e, const double* params, size_t pstride, int n_rollouts, int num_problems, int include_pads,
                double* costs, size_t cstride, int* flags, int flag_stride, int flag_offset, double* clipped,
                stomp_sphere_debug* debug, int* cflags = nullptr) {
  if (e.cull_dirty && ensure_cull(e)) return 1;
  const bool torque = e.torque_weight > 1e-9 && e.chain_len > 0;   // the reference's test, src/stomp_optimizer.cpp:1120
  if (torque && !clipped) {
    const size_t need = pstride * size_t(num_problems);
    if (e.torque_q.n < need) {
      CUDA_TRY(cudaStreamSynchronize(e.stream));
      CUDA_TRY(cudaStreamSynchronize(e.tail_stream));
      CUDA_TRY(e.torque_q.alloc(need));
    }
    clipped = e.torque_q.p;
  }
  if (launch_cost_only(e, params, pstride, n_rollouts, num_problems, include_pads, costs, cstride, flags, flag_stride, flag_offset,
                       clipped, debug, cflags))
    return 1;
  return torque ? launch_torque(e, clipped, pstride, n_rollouts, num_problems, costs, cstride) : 0;
}

int launch_cost_only(Engine& e, const double* params, size_t pstride, int n_rollouts, int num_problems, int include_pads,
                     double* costs, size_t cstride, int* flags, int flag_stride, int flag_offset, double* clipped,
                     stomp_sphere_debug* debug, int* cflags) {
  if (!e.have_robot || !e.have_sdf) return fail("set_robot and set_sdf must be called before the cost plugin runs");
  if (!e.have_problems) return fail("set_problems must be called before the cost plugin runs");
  if (e.f32) {
    CostArgs<float> a = base_cost_args<float>(e);
    a.n_rollouts = n_rollouts; a.include_pads = include_pads;
    a.params = params; a.params_problem_stride = pstride; a.params_rollout_stride = size_t(e.D) * e.N;
    a.costs = costs; a.cost_problem_stride = cstride;
    a.collision_free = flags; a.flag_problem_stride = flag_stride; a.flag_offset = flag_offset;
    a.constraints_satisfied = cflags;
    a.clipped = clipped; a.debug = debug;
    return debug ? launch_cost_t<float, true>(e, a, num_problems) : launch_cost_t<float, false>(e, a, num_problems);
  }
  CostArgs<double> a = base_cost_args<double>(e);
  a.n_rollouts = n_rollouts; a.include_pads = include_pads;
  a.params = params; a.params_problem_stride = pstride; a.params_rollout_stride = size_t(e.D) * e.N;
  a.costs = costs; a.cost_problem_stride = cstride;
  a.collision_free = flags; a.flag_problem_stride = flag_stride; a.flag_offset = flag_offset;
  a.constraints_satisfied = cflags;
  a.clipped = clipped; a.debug = debug;
  return debug ? launch_cost_t<double, true>(e, a, num_problems) : launch_cost_t<double, false>(e, a, num_problems);
}

int block_for(int N) { return std::min(1024, ((N + 31) / 32) * 32); }

int launch_cumulative(Engine& e, int r_begin, int r_count, bool totals_only, bool keep_totals);
// Huge / sharded path when the statistics kernels add S + C themselves and no tap wants the cumulative costs: without rollout
// reuse nothing ranks the totals either, so nothing is launched (stomp_engine_get fills both on demand); with reuse only the
// totals are produced.
int launch_cumulative_huge(Engine& e, int r_begin, int r_count) {
  const bool lean = e.direct_now() && !e.desc.keep_intermediates;
  if (lean && e.Rre == 0) {
    e.cumulative_stale = e.totals_stale = true;
    return 0;
  }
  return launch_cumulative(e, r_begin, r_count, lean, false);
}

// totals_only: k_update reads S and C itself (direct_now) and nobody asked for the cumulative-cost tap — the D N doubles per
// rollout are not written (stomp_engine_get computes them on demand)
int launch_cumulative(Engine& e, int r_begin = 0, int r_count = -1, bool totals_only = false, bool keep_totals = false) {
  if (r_count < 0) r_count = e.R - r_begin;
  if (r_count == 0) return 0;
  if (!keep_totals && r_begin == 0 && r_count == e.R) e.totals_stale = false;
  // the control costs of every slot in the range were summed by k_generate: only the state costs are left to add
  const bool have_csum = (r_begin >= e.num_gen || e.csum_new) && (r_begin + r_count <= e.num_gen || e.csum_reused);
  if (totals_only && e.use_totals_kernel && have_csum && !keep_totals) {
    const long long num = (long long)e.B * r_count;
    begin_launch(e);
    k_totals<<<unsigned((num * 32 + 127) / 128), 128, 0, e.ws>>>(e.R, r_begin, r_count, e.D, e.N, num, e.state[e.cur].p, e.csum.p, e.totals.p);
    e.cumulative_stale = true;
    return check_launch(e, "k_totals");
  }
  begin_launch(e);
  k_cumulative<<<unsigned(e.B) * r_count, block_for(e.N), 0, e.ws>>>(e.R, r_begin, r_count, e.D, e.N, e.desc.use_cumulative_costs,
                                                                       e.state[e.cur].p, e.control[e.cur].p,
                                                                       totals_only ? nullptr : e.cumulative.p,
                                                                       keep_totals ? nullptr : e.totals.p);
  e.cumulative_stale = totals_only;
  return check_launch(e, "k_cumulative");
}

size_t band_smem(const Engine& e, int rows) { return (size_t(e.N) * 16 + size_t(rows) * (e.N | 1)) * 8; }

int launch_update(Engine& e, int apply, bool fuse_extra_control) {
  UpdateArgs a;
  std::memset(&a, 0, sizeof(a));
  a.R = e.R; a.D = e.D; a.N = e.N; a.apply = apply;
  if (fuse_extra_control) {
    a.extra_control = e.extra_control.p;
    a.pad_start = e.pad_start.p; a.pad_goal = e.pad_goal.p;
    a.control_weight = 0.5 * e.control_cost_weight;
    a.st = e.stencil();
  }
  a.cumulative = e.cumulative.p; a.noise = e.noise.p; a.probabilities = e.probabilities.p;
  // without the suffix sums a cumulative cost is just S + C[d]: k_update forms it itself (same addition, same bits) and does
  // not depend on k_cumulative, which the two-stream schedule then takes off the critical path
  if (e.direct_now()) { a.state = e.state[e.cur].p; a.cumulative = e.control[e.cur].p; }
  a.updates = e.updates.p; a.theta = e.theta.p; a.band = e.band_view();
  // the projection as a dense product (kernels.cuh) while the N x N matrix is a few hundred KB of L2-resident reads per CTA
  a.dense_ms = (e.dense_update && e.N <= 512) ? e.dense_ms.p : nullptr;
  a.use_dmma = e.dmma_update ? 1 : 0;
  auto update_smem = [&](int rows) {
    const size_t stride = size_t(e.N + 2 * kPad) | 1;
    return (2 * rows * stride + size_t(rows) * 2 * kPad + (a.dense_ms ? 0 : size_t(e.N) * 16)) * 8;
  };
  // enough CTAs to fill the machine twice when the batch allows it; otherwise one dimension per CTA
  int dpc = int(std::min<long long>(std::min(e.D, 32), std::max<long long>(1, (long long)e.B * e.D / 296)));
  if (e.update_dpc > 0) dpc = std::max(1, std::min(e.D, e.update_dpc));   // A/B (STOMP_UPDATE_DPC): dimensions per CTA
  size_t smem = update_smem(dpc);
  while (smem > 200 * 1024 && dpc > 1) smem = update_smem(--dpc);
  a.dims_per_cta = dpc;
  const int groups = (e.D + dpc - 1) / dpc;
  // few CTAs (less than one per SM): wide CTAs, the latency of one CTA is the latency of the kernel
  const bool wide = e.wide_update && (long long)e.B * groups <= e.num_sms && a.dense_ms != nullptr;
  const bool direct = a.state != nullptr;
  auto launch = [&](auto kernel, int threads) -> int {
    if (ensure_smem(e, kernel, 220 * 1024)) return 1;
    begin_launch(e);
    kernel<<<unsigned(e.B) * groups, threads, smem, e.ws>>>(a);
    return check_launch(e, "k_update");
  };
  if (wide) return direct ? launch(k_update<512, 1, true>, 512) : launch(k_update<512, 1, false>, 512);
  const int tpb_max = 128;   // A/B on B200 (C2): 64 threads 0.163, 128: 0.097, 256: 0.150, 512: 0.244 ms
  const int threads = std::min(tpb_max, ((dpc * e.N + 31) / 32) * 32);
  return direct ? launch(k_update<128, 7, true>, threads) : launch(k_update<128, 7, false>, threads);
}

// huge-R / sharded statistics (B == 1): one k_shard_stats launch per phase (partials, reduction over the chunks, optional
// peer-memory exchange, optional fused projection + update)
constexpr int kChunks = 128;
int launch_finalize(Engine& e, int apply);

int launch_shard_stats(Engine& e, bool is_max, bool exchange, bool finalize, int apply) {
  // at least 32 rollouts per chunk: a chunk-CTA should have more to do than its share of the cross-chunk reduction
  const int DN = e.D * e.N, rpc = std::max((e.R + kChunks - 1) / kChunks, std::min(e.R, 32)), nch = (e.R + rpc - 1) / rpc;
  const int colblocks = (DN + 127) / 128;
  // the fused finalize keeps the band tables and all D rows in the last CTA's shared memory — which every CTA of the launch then
  // reserves: only worth it while that is small (C3: 18 KB; C5's 110 KB cut the partial reductions to 2 CTAs / SM, 0.70 -> 0.88 ms
  // per iteration).  The probability tap needs the global sums before the update.  Both fall back to a separate k_finalize launch.
  size_t smem = (size_t(e.N) * 16 + size_t(e.D) * (e.N | 1)) * 8;
  const bool fuse = finalize && !e.probabilities.p && smem <= 24 * 1024;
  if (!fuse) smem = 0;
  ShardStatsArgs a;
  std::memset(&a, 0, sizeof(a));
  a.R = e.R; a.D = e.D; a.N = e.N; a.rollouts_per_chunk = rpc;
  a.is_max = is_max ? 1 : 0;
  a.do_exchange = exchange ? 1 : 0;
  a.do_finalize = fuse ? 1 : 0;
  a.apply = apply;
  a.cumulative = e.cumulative.p; a.noise = e.noise.p; a.minmax = e.minmax.p; a.part = e.part.p;
  if (e.direct_now() && !e.probabilities.p) { a.state = e.state[e.cur].p; a.cumulative = e.control[e.cur].p; }   // S + C formed in the kernel
  a.out = is_max ? e.minmax.p : e.sums.p;
  a.counters = e.shard_counters.p;
  a.rank = e.desc.rollout_shard_rank; a.world = e.desc.rollout_shard_world;
  if (exchange) {
    a.epoch = ++e.xchg_epoch;
    for (int r = 0; r < a.world; ++r) a.peers[r] = e.peer_base[r];
  }
  a.err = e.xchg_err.p;
  a.updates = e.updates.p; a.theta = e.theta.p; a.band = e.band_view();
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_shard_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  begin_launch(e);
  k_shard_stats<<<dim3(colblocks, nch), 128, smem, e.ws>>>(a);
  if (check_launch(e, is_max ? "k_shard_stats_max" : "k_shard_stats_sum")) return 1;
  if (finalize && !fuse) return launch_finalize(e, apply);
  return 0;
}
int launch_finalize(Engine& e, int apply) {
  if (e.probabilities.p) {
    begin_launch(e);
    k_probabilities<<<1184, 256, 0, e.ws>>>(e.R, e.D * e.N, e.cumulative.p, e.minmax.p, e.sums.p, e.probabilities.p);
    if (check_launch(e, "k_probabilities")) return 1;
  }
  begin_launch(e);
  k_finalize<<<e.D, block_for(e.N), band_smem(e, 1), e.ws>>>(e.D, e.N, apply, e.sums.p, e.updates.p, e.theta.p, e.band_view(), e.xchg_err.p);
  return check_launch(e, "k_finalize");
}

// ---- PolicyImprovement steps ----------------------------------------------------------------------

// the iteration's noise scales: uploaded on the launching stream.  Handing them to the kernels in the launch arguments instead
// (STOMP_SCALES_IN_ARGS=1: no copy-engine hop on the chain) is SLOWER — C2 0.450 -> 0.481 ms, C1 47 -> 54 us per iteration, as in
// round 1: the hop after k_update happens to pace the main stream against the tail stream favourably
int set_noise_scale(Engine& e, const double* scale) {
  e.scale_host.assign(scale, scale + e.D);
  if (e.scales_in_args && e.D <= kMaxScaleArgs && !e.capturing) return 0;
  return upload(e, e.noise_scale, scale, size_t(e.D));
}

// generateRollouts bookkeeping (src/policy_improvement.cpp:166-176)
struct RolloutPlan {
  bool reuse = false, injected = false;
};

void plan_rollouts(Engine& e, RolloutPlan& p) {
  e.cur = 1 - e.cur;
  e.num_gen = e.R - e.Rre;
  p.reuse = false;
  if (!e.reused_next) {
    e.num_gen = e.R;
    if (e.Rre > 0) e.reused_next = true;
  } else {
    p.reuse = true;
  }
  p.injected = e.injected_pending;
  e.injected_pending = false;
  e.csum_new = e.csum_reused = false;
  ++e.generation;   // the Philox "iteration" counter
}

int launch_select(Engine& e) {
  begin_launch(e);
  k_select_reuse<<<e.B, 128, 0, e.ws>>>(e.totals.p, e.R, e.Rre, e.extra_added ? 1 : 0, e.reuse_src.p);
  if (check_launch(e, "k_select_reuse")) return 1;
  e.extra_added = false;
  return 0;
}

// noise / parameters / projected noise / control costs of rollout slots [r_begin, r_begin + r_count)
// which generation kernel a launch over r_count slots takes: 1 band solves, 2 dense (one CTA per vector), 3 DMMA
int generate_kind(const Engine& e, int r_count) {
  const double nvec = double(e.B) * r_count * e.D;
  const double est_band = 0.95e-6 * e.N;
  const double est_dense = 5.0e-6 + nvec * 12.0 * double(e.N) * e.N / 4.0e12;
  int kind = e.gen_mode == 4 ? 1 : e.gen_mode;   // 4 (seg) = the band formulation, always through k_generate_seg
  const bool mma_ok = mma_smem_bytes(e.N) <= 200 * 1024 && e.N <= 1024 && e.mma_a1.p != nullptr;
  if (kind == 0) kind = (e.N <= 1024 && est_dense < 0.5 * est_band) ? 2 : 1;
  if (kind == 3 && !mma_ok) kind = 1;
  return kind;
}

int launch_generate_range(Engine& e, const RolloutPlan& p, int r_begin, int r_count, bool with_control, bool pre = false,
                          bool candidates = false) {
  if (r_count <= 0) return 0;
  GenArgs a = base_gen_args(e);
  a.mode_generate = 1;
  a.mode_project = 1;
  a.mode_control = with_control ? 1 : 0;
  a.injected = p.injected ? 1 : 0;
  a.iteration = e.generation;
  a.control_weight = 0.5 * e.control_cost_weight;
  a.r_begin = r_begin;
  a.r_count = r_count;
  if (pre) {
    // look-ahead pass for the next iteration's new slots [0, r_count): compact outputs, the next Philox generation, no theta
    a.pre = 1;
    a.R = r_count; a.R_gen = r_count;
    a.iteration = e.generation + 1;
    a.theta = nullptr; a.params = nullptr; a.control = nullptr; a.params_prev = nullptr; a.reuse_src = nullptr; a.eps_in = nullptr;
    a.noise = e.pre_noise.p; a.noise_projected = e.pre_y.p;
    a.noise_scale = e.pre_scale.p;
    if (e.D <= kMaxScaleArgs && int(e.pre_scale_host.size()) == e.D) {     // (off the chain: no pacing effect, one API call less)
      a.scale_by_value = 1;
      for (int d = 0; d < e.D; ++d) a.scale_v[d] = e.pre_scale_host[d];
    } else {
      a.scale_by_value = 0;
    }
    a.iteration_ptr = nullptr;
  }
  if (candidates) {
    // candidate pass for the NEXT iteration's reused slots: every rollout of the current iteration and the noise-less one
    // (candidate R) treated as reused — noise = parameters - theta, M noise, control costs — into [B][R + 1][D][N] buffers
    a.R = e.R + 1; a.R_gen = 0; a.r_begin = 0; a.r_count = e.R + 1;
    a.params_prev = e.params[e.cur].p; a.prev_stride = e.R;
    a.reuse_src = e.cand_map.p;
    a.noise = e.cand_noise.p; a.params = e.cand_params.p; a.control = e.cand_control.p;
    a.noise_projected = e.noise_projected.p ? e.cand_y.p : nullptr;
    a.injected = 0;
    a.iteration_ptr = nullptr;
    r_begin = 0; r_count = e.R + 1;
  }
  const bool uses_injection = !candidates && p.injected && r_begin < e.num_gen;
  if (uses_injection && e.inject_pending_n < e.num_gen)
    return fail("injected noise holds " + std::to_string(e.inject_pending_n) + " rollouts per problem but this iteration generates " +
                std::to_string(e.num_gen));
  if (uses_injection) CUDA_TRY(cudaStreamWaitEvent(e.ws, e.ev_copy_done[e.inject_pending_buf], 0));
  // Two kernels produce the same outputs (the same linear maps on the same Philox normals; they agree to rounding):
  //  * k_generate_dense — one CTA per vector, a few microseconds of latency: small batches (one planning problem = 35 vectors)
  //  * k_generate       — serial band solves, ~0.95 us per timestep whatever the batch: everything else
  // (measured on a B200: C1 0.149 -> 0.097 ms per iteration with the dense kernel, but C5's 15 360 vectors of N = 300
  // 0.72 -> 2.44 ms, since every CTA streams 1.5 N^2 matrix elements from L2; a register-tiled variant sharing the matrices
  // between 32 vectors was bound by shared-memory operand bandwidth and lost everywhere: profiles/README.md)
  const double nvec = double(e.B) * r_count * e.D;
  // k_generate_mma — both linear maps as DMMA GEMMs over tiles of 16 vectors: large batches whose three shared tiles fit
  const size_t mma_smem = mma_smem_bytes(e.N);
  // measured on B200 (profiles/README.md, round 2): the DMMA kernel loses to the band solves (k_generate per iteration, both
  // launches: C2 0.266 vs 0.158 ms, C5 0.66 vs 0.37 ms; new-slot launch under ncu 149 vs 107 us with the fp64 tensor path at
  // 35 % of its peak).  The dense formulation does 1.5 N^2 flops per vector where the band solves do ~20 N, and B200's fp64
  // tensor rate equals its DFMA rate, so it is only taken on request (STOMP_GENERATE=mma)
  // (the candidate pass takes the kernel the reused slots themselves would: the same bits whichever schedule ran)
  const int kind = generate_kind(e, candidates ? e.Rre : r_count);
  if (kind == 3) {
    if (mma_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_generate_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, int(mma_smem)));
    begin_launch(e);
    k_generate_mma<<<unsigned((nvec + kMmaV - 1) / kMmaV), 128, mma_smem, e.ws>>>(a, e.mma_a1.p, e.mma_a2.p);
    if (check_launch(e, "k_generate")) return 1;
  } else if (kind == 2) {
    const size_t smem = (size_t(3) * e.N + 2 * kPad) * 8;
    begin_launch(e);
    k_generate_dense<<<unsigned(nvec), 128, smem, e.ws>>>(a, e.dense_cinv.p, e.dense_ms.p);
    if (check_launch(e, pre ? "k_generate_ahead" : "k_generate")) return 1;
  } else if (e.gen_mode == 4 && e.seg_P > 0 && !pre) {
    // band solves split over seg_P time segments per vector: P x G warps per CTA, G groups of 32 vectors.  Measured on B200
    // (C2, k_generate per iteration, both launches): 0.213 ms (P = 2, G = 2; 0.227 at P = 4) against 0.158 ms for one thread per
    // vector.  ncu of the new-slot launch: 118.6 vs 107.2 us — the spike corrections, the extra scratch passes and the boundary
    // chains cost 52 % more warp instructions (55.7 M vs 36.7 M) and the kernel is more issue- than latency-bound (issue slots
    // 45 % busy at 16 warps / SM); at 80 registers / 3 CTAs per SM it spills (0.38 ms).  Only taken on request
    // (STOMP_GENERATE=seg)
    int G = std::max(1, 8 / e.seg_P);
    if (const char* sg = getenv("STOMP_SEG_G")) G = std::max(1, std::min(8 / e.seg_P, atoi(sg)));
    const int P = e.seg_P;
    const long long nv = (long long)nvec;
    const unsigned grid = unsigned((nv + 32 * G - 1) / (32 * G));
    const size_t stride = size_t(grid) * 32 * G;
    DevBuf<double>& scratch = e.ws == e.tail_stream ? e.gen_scratch2
                              : (e.ws == e.pre_stream ? e.gen_scratch3 : (e.ws == e.cand_stream ? e.gen_scratch4 : e.gen_scratch));
    if (scratch.n < stride * e.N) CUDA_TRY(scratch.alloc(stride * e.N));
    a.scratch = scratch.p;
    a.scratch_stride = stride;
    a.segs = P; a.seg_hf = e.seg_hf.p; a.seg_hb = e.seg_hb.p;
    const size_t smem = (size_t(e.N) * (16 + 12 + 1) + size_t(G) * P * 6 * 32 + size_t(P) * G * 2 * 32 * kTileLd) * 8;
    if (smem > 220 * 1024) return fail("num_time_steps too large for the band tables of k_generate_seg");
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_generate_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    begin_launch(e);
    k_generate_seg<<<grid, 32 * P * G, smem, e.ws>>>(a);
    if (check_launch(e, "k_generate")) return 1;
  } else {
    const bool sums = !pre && !candidates && with_control && e.csum.p != nullptr;
    if (sums) a.csum = e.csum.p;
    if (launch_generate(e, a)) return 1;
    if (sums) {      // which slots' control-cost sums this launch left behind (launch_cumulative)
      if (r_begin < e.num_gen && r_begin == 0 && r_begin + r_count >= e.num_gen) e.csum_new = true;
      if (r_begin + r_count == e.R && r_begin <= e.num_gen && e.num_gen < e.R) e.csum_reused = true;
    }
  }
  if (uses_injection) CUDA_TRY(cudaEventRecord(e.ev_consumed[e.inject_pending_buf], e.ws));
  return 0;
}

// candidate pass for the next iteration's reused slots, on cand_stream after this iteration's update (ev_upd)
bool candidates_possible(const Engine& e) {
  return e.lookahead && e.small_batch() && e.direct_now() && e.overlap && !e.capturing && !e.huge_path() && e.Rre > 0 &&
         e.reused_next && e.R <= 64 && e.desc.rollout_shard_world == 1 && generate_kind(e, e.Rre) != 3;
}

int launch_candidates(Engine& e) {
  const size_t n = size_t(e.B) * (e.R + 1) * e.D * e.N;
  if (e.cand_noise.n < n) {
    CUDA_TRY(cudaStreamSynchronize(e.cand_stream));
    CUDA_TRY(e.cand_noise.alloc(n));
    CUDA_TRY(e.cand_control.alloc(n));
    CUDA_TRY(e.cand_params.alloc(n));
    if (e.noise_projected.p) CUDA_TRY(e.cand_y.alloc(n));
    std::vector<int> map(size_t(e.B) * (e.R + 1));
    for (int b = 0; b < e.B; ++b)
      for (int r = 0; r <= e.R; ++r) map[size_t(b) * (e.R + 1) + r] = r < e.R ? r : -1;
    CUDA_TRY(e.cand_map.alloc(map.size()));
    CUDA_TRY(cudaMemcpy(e.cand_map.p, map.data(), map.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaStreamWaitEvent(e.cand_stream, e.ev_upd, 0));
  RolloutPlan np;
  e.ws = e.cand_stream;
  const int rc = launch_generate_range(e, np, 0, e.R + 1, true, false, true);
  e.ws = e.stream;
  if (rc) return 1;
  CUDA_TRY(cudaEventRecord(e.ev_cand, e.cand_stream));
  e.cand_valid = true;
  e.cand_dirty = true;
  e.cand_generation = e.generation;
  e.cand_epoch = e.config_epoch;
  return 0;
}

int launch_select_gather(Engine& e) {
  SelectGatherArgs a;
  std::memset(&a, 0, sizeof(a));
  a.R = e.R; a.R_reuse = e.Rre; a.D = e.D; a.N = e.N; a.use_extra = e.extra_added ? 1 : 0;
  a.totals = e.totals.p; a.reuse_src = e.reuse_src.p;
  a.cand_noise = e.cand_noise.p; a.cand_control = e.cand_control.p; a.cand_y = e.noise_projected.p ? e.cand_y.p : nullptr;
  a.params_prev = e.params[1 - e.cur].p; a.theta = e.theta.p;
  a.state_prev = e.state[1 - e.cur].p; a.extra_state = e.extra_state.p;
  a.extra_control = e.extra_control.p;
  a.noise = e.noise.p; a.control = e.control[e.cur].p; a.params = e.params[e.cur].p; a.noise_projected = e.noise_projected.p;
  a.state = e.state[e.cur].p;
  begin_launch(e);
  k_select_gather<<<unsigned(e.B) * e.Rre, kSelectGatherThreads, 0, e.ws>>>(a);
  if (check_launch(e, "k_select_gather")) return 1;
  e.extra_added = false;
  return 0;
}

// second half of the pipelined generation (k_finish_rollouts): parameters, noise, control costs of the new slots from the
// look-ahead buffers and the updated theta
int launch_finish_rollouts(Engine& e) {
  FinishArgs a;
  std::memset(&a, 0, sizeof(a));
  a.B = e.B; a.R = e.R; a.G = e.num_gen; a.D = e.D; a.N = e.N;
  a.theta = e.theta.p; a.pre_noise = e.pre_noise.p; a.pre_y = e.pre_y.p;
  a.pad_start = e.pad_start.p; a.pad_goal = e.pad_goal.p;
  a.noise = e.noise.p; a.params = e.params[e.cur].p; a.control = e.control[e.cur].p; a.noise_projected = e.noise_projected.p;
  a.control_weight = 0.5 * e.control_cost_weight;
  a.st = e.stencil();
  const long long total = (long long)e.B * e.num_gen * e.D * e.N;
  begin_launch(e);
  k_finish_rollouts<<<unsigned((total + 255) / 256), 256, 0, e.ws>>>(a);
  return check_launch(e, "k_finish_rollouts");
}

// May the noise of iteration `next_iteration` be generated now, before the current iteration's update?  Only what cannot
// change in between is assumed: engine noise (not injected), the steady-state slot layout, the band / dense kernels.
bool lookahead_possible(const Engine& e, const RolloutPlan& p) {
  return e.lookahead && e.small_batch() && e.overlap && !e.capturing && !e.huge_path() && !p.injected &&
         e.desc.rollout_shard_world == 1 && (e.reused_next || e.Rre == 0) && e.R - e.Rre > 0 &&
         generate_kind(e, e.R - e.Rre) != 3;
}

// enqueue the look-ahead pass of iteration next_iteration on pre_stream; it starts once k_finish_rollouts / k_generate of the
// current iteration (ev_fin) no longer read the look-ahead buffers
int launch_lookahead(Engine& e, int next_iteration) {
  const int G = e.R - e.Rre;
  const size_t n = size_t(e.B) * G * e.D * e.N;
  if (e.pre_noise.n < n) {
    CUDA_TRY(cudaStreamSynchronize(e.pre_stream));
    CUDA_TRY(e.pre_noise.alloc(n));
    CUDA_TRY(e.pre_y.alloc(n));
  }
  if (e.pre_scale.n < size_t(e.D)) CUDA_TRY(e.pre_scale.alloc(size_t(e.D)));
  std::vector<double> scale(e.D);
  for (int d = 0; d < e.D; ++d) scale[d] = e.noise_stddev[d] * std::pow(e.noise_decay[d], next_iteration - 1);
  CUDA_TRY(cudaStreamWaitEvent(e.pre_stream, e.ev_fin, 0));
  e.pre_scale_host = scale;
  if (e.D > kMaxScaleArgs)
    CUDA_TRY(cudaMemcpyAsync(e.pre_scale.p, scale.data(), size_t(e.D) * 8, cudaMemcpyHostToDevice, e.pre_stream));
  RolloutPlan np;
  e.ws = e.pre_stream;
  const int rc = launch_generate_range(e, np, 0, G, true, true);
  e.ws = e.stream;
  if (rc) return 1;
  CUDA_TRY(cudaEventRecord(e.ev_pre, e.pre_stream));
  e.pre_valid = true;
  e.pre_dirty = true;
  e.pre_generation = e.generation + 1;
  e.pre_iteration = next_iteration;
  e.pre_num_gen = G;
  e.pre_epoch = e.config_epoch;
  return 0;
}

// serial form: reuse selection + all rollout slots on the current stream (step-by-step API, sharded phases)
int step_get_rollouts(Engine& e, int iteration_number, bool with_control) {
  (void)iteration_number;
  RolloutPlan p;
  plan_rollouts(e, p);
  if (p.reuse && launch_select(e)) return 1;
  return launch_generate_range(e, p, 0, e.R, with_control);
}

__global__ void k_gather_state(int R, int R_gen, int N, const int* __restrict__ reuse_src, const double* __restrict__ prev,
                               const double* __restrict__ extra, double* __restrict__ cur) {
  const int b = blockIdx.x / (R - R_gen), j = blockIdx.x - b * (R - R_gen);
  const int src = reuse_src[size_t(b) * (R - R_gen) + j];
  const double* s = src >= 0 ? prev + (size_t(b) * R + src) * N : extra + size_t(b) * N;
  double* d = cur + (size_t(b) * R + R_gen + j) * N;
  for (int t = threadIdx.x; t < N; t += blockDim.x) d[t] = s[t];
}

int gather_reused_state(Engine& e) {
  if (e.num_gen == e.R) return 0;
  begin_launch(e);
  k_gather_state<<<unsigned(e.B) * (e.R - e.num_gen), 128, 0, e.ws>>>(e.R, e.num_gen, e.N, e.reuse_src.p,
                                                                          e.state[1 - e.cur].p, e.extra_state.p, e.state[e.cur].p);
  return check_launch(e, "k_gather_state");
}

int step_control_costs(Engine& e) {  // computeRolloutControlCosts for all R rollouts with the current weight
  GenArgs a = base_gen_args(e);
  a.mode_generate = 0; a.mode_project = 1; a.mode_control = 1;
  a.control_weight = 0.5 * e.control_cost_weight;
  return launch_generate(e, a);
}

int step_improve(Engine& e, int apply) {
  if (launch_cumulative(e)) return 1;
  if (e.huge_path()) {
    if (e.B != 1) return fail("the sharded / huge-rollout statistics path supports num_problems == 1 only");
    return launch_shard_stats(e, true, false, false, 0) || launch_shard_stats(e, false, false, true, apply);
  }
  return launch_update(e, apply, false);
}

int step_extra(Engine& e, bool run_cost, int iteration_number, bool have_control = false, bool offchain = false) {
  e.extra_offchain = false;
  if (run_cost) {
    // the previous noise-less rollout's k_extra_total / k_track_best (cand_stream) read what this one overwrites
    if (e.extra_dirty) CUDA_TRY(cudaStreamWaitEvent(e.ws, e.ev_extra, 0));
    if (launch_cost(e, e.theta.p, size_t(e.D) * e.N, 1, e.B, iteration_number == 1, e.extra_state.p, size_t(e.N),
                    e.collision_free.p, e.R + 1, e.R, e.extra_clipped.p, nullptr, e.constraints_ok.p))
      return 1;
  }
  if (!have_control) {
    GenArgs a = base_gen_args(e);
    a.extra = 1; a.mode_generate = 0; a.mode_project = 0; a.mode_control = 1;
    a.control = e.extra_control.p;
    a.control_weight = 0.5 * e.control_cost_weight;
    if (launch_generate(e, a)) return 1;
  }
  cudaStream_t back = e.ws;
  if (offchain) {      // the next k_select_gather sums the noise-less rollout's total itself: nothing on this stream waits for it
    CUDA_TRY(cudaEventRecord(e.ev_nl, e.ws));
    CUDA_TRY(cudaStreamWaitEvent(e.cand_stream, e.ev_nl, 0));
    e.ws = e.cand_stream;
  }
  begin_launch(e);
  k_extra_total<<<e.B, 128, 0, e.ws>>>(e.R, e.D, e.N, e.extra_state.p, e.extra_control.p, e.totals.p, e.noiseless_sum.p);
  const int rc = check_launch(e, "k_extra_total");
  if (offchain && !rc) {
    CUDA_TRY(cudaEventRecord(e.ev_extra, e.cand_stream));
    e.extra_dirty = e.extra_offchain = true;
  }
  e.ws = back;
  if (rc) return 1;
  e.extra_added = true;
  return 0;
}

int iterate_front(Engine& e, int iteration_number) {  // up to and including k_cumulative
  std::vector<double> scale(e.D);
  for (int d = 0; d < e.D; ++d) scale[d] = e.noise_stddev[d] * std::pow(e.noise_decay[d], iteration_number - 1);
  if (set_noise_scale(e, scale.data())) return 1;
  e.control_cost_weight = e.desc.smoothness_cost_weight;
  if (step_get_rollouts(e, iteration_number, true)) return 1;
  if (gather_reused_state(e)) return 1;
  if (launch_cost(e, e.params[e.cur].p, size_t(e.R) * e.D * e.N, e.num_gen, e.B, iteration_number == 1, e.state[e.cur].p,
                  size_t(e.R) * e.N, e.collision_free.p, e.R + 1, 0, e.clipped.p, nullptr, e.constraints_ok.p))
    return 1;
  // like the two-stream schedule: only the totals when k_update adds S + C itself and nobody taps the cumulative costs
  if (e.huge_path()) return launch_cumulative_huge(e, 0, e.R);
  return launch_cumulative(e, 0, -1, e.direct_now() && !e.desc.keep_intermediates);
}

int iterate_serial(Engine& e, int iteration_number) {
  if (join_streams(e)) return 1;
  if (iterate_front(e, iteration_number)) return 1;
  if (e.huge_path()) {
    if (launch_shard_stats(e, true, false, false, 0) || launch_shard_stats(e, false, false, true, 1)) return 1;
  } else {
    if (launch_update(e, 1, true)) return 1;
    return step_extra(e, true, iteration_number, true);
  }
  return step_extra(e, true, iteration_number);
}

// One iteration on two streams.  The noise-less rollout of iteration i, the reuse selection it feeds and everything about the
// REUSED rollout slots of iteration i+1 (gather, M*eps, control costs) only matter to iteration i+1's statistics, not to its
// new rollouts: they run on the tail stream while the main stream already samples and costs the new rollouts of
// iteration i+1.  (A caller that reads the statistics after every iteration synchronises both and sees no overlap.)
//   main: [k_generate(new) -> k_cost(new) -> k_totals(new)] -> wait(tail) -> k_update                  -> record(upd)
//   tail: [k_select_reuse -> k_generate(reused) -> k_gather_state -> k_totals(reused)] record(tail)
//         ... wait(upd) -> k_cost(noise-less) -> k_extra_total
// (k_cumulative in place of k_totals with cumulative costs, taps, or a generation kernel that leaves no per-vector sums.)
// Small batches run the latency schedule on four streams instead:
//   main: wait(pre) -> k_finish_rollouts(new) -> k_cost(new) -> wait(tail) -> k_update                 -> record(upd)
//   tail: wait(upd) -> k_cost(noise-less) -> wait(cand) -> k_select_gather -> record(tail); wait(cost) -> k_cumulative
//   pre : wait(finish) -> k_generate_ahead(new slots of the next iteration)                           -> record(pre)
//   cand: wait(upd) -> k_generate(candidates) -> record(cand); wait(noise-less cost) -> k_extra_total [-> k_track_best]
// Every speculative piece (look-ahead pass, candidate pass) is keyed by generation / iteration number / configuration epoch
// and falls back to the one-pass kernels when the next call is not the one it was prepared for.
int iterate_once(Engine& e, int iteration_number) {
  if (e.huge_path()) {
    if (e.B != 1) return fail("the sharded / huge-rollout statistics path supports num_problems == 1 only");
    if (e.desc.rollout_shard_world > 1)
      return fail("rollout-sharded engines iterate through stomp_engine_iterate_sharded_phase");
  }
  if (!e.overlap || (e.prof_on && !e.prof_timeline)) return iterate_serial(e, iteration_number);
  e.ws = e.stream;
  if (e.capturing) {
    begin_launch(e);
    k_advance_iteration<<<1, 32, 0, e.stream>>>(e.dev_generation.p, e.dev_table_index.p, e.scale_table.p, e.D, e.noise_scale.p);
    if (check_launch(e, "k_advance_iteration")) return 1;
  } else {
    std::vector<double> scale(e.D);
    for (int d = 0; d < e.D; ++d) scale[d] = e.noise_stddev[d] * std::pow(e.noise_decay[d], iteration_number - 1);
    if (set_noise_scale(e, scale.data())) return 1;
  }
  e.control_cost_weight = e.desc.smoothness_cost_weight;
  RolloutPlan p;
  plan_rollouts(e, p);
  // direct k_update (it adds S + C itself): the totals either stay where k_cumulative was (lean) or follow k_update on the tail
  const bool direct = !e.huge_path() && e.direct_now();
  // (measured, same box: C2 old schedule 0.4652 ms, lean k_cumulative on the chain 0.4611, after the update 0.4672; C4 2.107 /
  //  2.093 / 2.007 — with 420 spheres and 200 timesteps per rollout the cost kernels dwarf the generation it then overlaps)
  const bool heavy_cost = (long long)e.K * e.N >= 20000;
  const bool late_cumulative = direct && (e.small_batch() || e.cum_placement == 2 || (e.cum_placement == 0 && heavy_cost));
  const bool totals_only = direct && !e.desc.keep_intermediates;
  const bool use_cand = p.reuse && late_cumulative && e.cand_valid && !e.capturing && e.cand_generation + 1 == e.generation &&
                        e.cand_epoch == e.config_epoch;
  e.cand_valid = false;
  if (use_cand) {
    // the reused slots were prepared by the candidate pass: rank and copy
    e.ws = e.tail_stream;
    CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_cand, 0));
    e.cand_dirty = false;
    if (launch_select_gather(e)) { e.ws = e.stream; return 1; }
  } else if (p.reuse) {
    e.ws = e.tail_stream;
    if (e.cand_dirty) { CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_cand, 0)); e.cand_dirty = false; }   // a discarded pass
    if (e.extra_dirty) CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_extra, 0));   // k_select_reuse reads totals[R]
    if (launch_select(e)) { e.ws = e.stream; return 1; }
    // k_select_reuse ranks the PREVIOUS iteration's Rollout::getCost() values; this iteration's k_cumulative on the main
    // stream overwrites the new slots' entries of that array and must not overtake it.  (It never did while the main
    // stream's path to it was ~330 us long, but nothing ordered the two: found when an experiment shortened that path.)
    CUDA_TRY(cudaEventRecord(e.ev_selected, e.tail_stream));
    if (launch_generate_range(e, p, e.num_gen, e.R - e.num_gen, true) || gather_reused_state(e) ||
        (!late_cumulative && (e.huge_path() ? launch_cumulative_huge(e, e.num_gen, e.R - e.num_gen)
                                            : launch_cumulative(e, e.num_gen, e.R - e.num_gen, totals_only)))) {
      e.ws = e.stream;
      return 1;
    }
  }
  CUDA_TRY(cudaEventRecord(e.ev_tail, e.tail_stream));   // also covers the previous iteration's noise-less rollout
  if (e.chain_probe && p.reuse) CUDA_TRY(cudaEventRecord(e.probe_tail, e.tail_stream));
  e.ws = e.stream;
  // new rollouts: finish what the look-ahead pass prepared during the previous iteration, or generate them in one pass
  const bool use_pre = e.pre_valid && !p.injected && !e.capturing && e.pre_generation == e.generation &&
                       e.pre_iteration == iteration_number && e.pre_num_gen == e.num_gen && e.pre_epoch == e.config_epoch;
  e.pre_valid = false;
  if (use_pre) {
    CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_pre, 0));
    e.pre_dirty = false;
    if (launch_finish_rollouts(e)) return 1;
  } else {
    if (e.pre_dirty) { CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_pre, 0)); e.pre_dirty = false; }   // a discarded pass
    if (launch_generate_range(e, p, 0, e.num_gen, true)) return 1;
  }
  const bool ahead = lookahead_possible(e, p);
  if (ahead) CUDA_TRY(cudaEventRecord(e.ev_fin, e.stream));
  if (launch_cost(e, e.params[e.cur].p, size_t(e.R) * e.D * e.N, e.num_gen, e.B, iteration_number == 1, e.state[e.cur].p,
                  size_t(e.R) * e.N, e.collision_free.p, e.R + 1, 0, e.clipped.p, nullptr, e.constraints_ok.p))
    return 1;
  if (ahead && launch_lookahead(e, iteration_number + 1)) return 1;
  if (late_cumulative) {
    CUDA_TRY(cudaEventRecord(e.ev_cost, e.stream));
  } else {
    if (p.reuse) CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_selected, 0));
    // new slots; the reused slots' were done on the tail stream
    if (e.huge_path() ? launch_cumulative_huge(e, 0, e.num_gen) : launch_cumulative(e, 0, e.num_gen, totals_only)) return 1;
  }
  if (e.chain_probe && p.reuse) CUDA_TRY(cudaEventRecord(e.probe_main, e.stream));
  CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_tail, 0));
  const bool huge = e.huge_path();
  if (huge ? (launch_shard_stats(e, true, false, false, 0) || launch_shard_stats(e, false, false, true, 1)) : launch_update(e, 1, true)) return 1;
  CUDA_TRY(cudaEventRecord(e.ev_upd, e.stream));
  if (late_cumulative && candidates_possible(e) && launch_candidates(e)) return 1;
  if (late_cumulative) {
    // Rollout::getCost() of every slot for the next reuse selection (and the cumulative-cost tap): beside k_update (small
    // batches) or after it (the control costs are double-buffered, so the next iteration's k_generate need not wait for it)
    e.ws = e.tail_stream;
    CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.small_batch() ? e.ev_cost : e.ev_upd, 0));
    if (launch_cumulative(e, 0, -1, totals_only)) { e.ws = e.stream; return 1; }
    e.ws = e.stream;
  }
  if (e.chain_probe && p.reuse && (iteration_number % 16) == 0) {      // sampled: the read-back synchronises
    CUDA_TRY(cudaEventRecord(e.probe_upd, e.stream));
    CUDA_TRY(cudaEventSynchronize(e.probe_upd));
    CUDA_TRY(cudaEventSynchronize(e.probe_tail));
    float d = 0.f, u = 0.f;
    if (cudaEventElapsedTime(&d, e.probe_main, e.probe_tail) == cudaSuccess &&
        cudaEventElapsedTime(&u, e.probe_main, e.probe_upd) == cudaSuccess) {
      e.probe_sum_us += 1e3 * d;
      e.probe_upd_sum_us += 1e3 * u;
      ++e.probe_n;
    }
  }
  e.ws = e.tail_stream;
  CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_upd, 0));
  const int rc = step_extra(e, true, iteration_number, !huge, e.offchain_extra && late_cumulative && candidates_possible(e));
  e.tail_dirty = true;
  e.ws = e.stream;
  if (p.reuse || e.Rre == 0) ++e.steady_iterations;
  return rc;
}

// ---- CUDA-graph replay ------------------------------------------------------------------------------------------------------
constexpr int kGraphIters = 8;   // even: the params / state-cost ping-pong returns to where it started

bool graph_eligible(const Engine& e, int iteration_number) {
  return e.graph_mode != 0 && e.overlap && !e.prof_on && !e.chain_probe && !e.injected_pending && !e.huge_path() &&
         e.desc.rollout_shard_world == 1 && iteration_number > 1 && (e.reused_next || e.Rre == 0) && e.steady_iterations >= 2 &&
         !e.cull_dirty && e.have_robot && e.have_sdf && e.have_problems;
}

// captures kGraphIters steady-state iterations (main + tail stream) into e.graph_exec; host bookkeeping is restored afterwards
int capture_iterations(Engine& e) {
  cudaGraphExec_t& exec = e.graph_exec[e.cur];
  if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; }
  if (join_streams(e)) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  const int cur0 = e.cur, num_gen0 = e.num_gen;
  const uint32_t gen0 = e.generation;
  const bool extra0 = e.extra_added, reused0 = e.reused_next;
  const int64_t launches0 = e.launches, steady0 = e.steady_iterations;
  CUDA_TRY(cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal));
  e.capturing = true;
  int rc = 0;
  // fork: the tail stream joins the capture
  if (cudaEventRecord(e.ev_upd, e.stream) != cudaSuccess || cudaStreamWaitEvent(e.tail_stream, e.ev_upd, 0) != cudaSuccess) rc = 1;
  e.tail_dirty = false;
  for (int k = 0; k < kGraphIters && !rc; ++k) rc = iterate_once(e, 2);   // iteration_number only matters through "== 1"
  // join: everything on the tail stream is part of the graph
  if (!rc && (cudaEventRecord(e.ev_tail, e.tail_stream) != cudaSuccess || cudaStreamWaitEvent(e.stream, e.ev_tail, 0) != cudaSuccess)) rc = 1;
  e.capturing = false;
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(e.stream, &graph);
  e.graph_launches = e.launches - launches0;
  e.cur = cur0; e.num_gen = num_gen0; e.generation = gen0; e.extra_added = extra0; e.reused_next = reused0;
  e.launches = launches0; e.steady_iterations = steady0;
  e.tail_dirty = false;
  e.ws = e.stream;
  if (rc || ce != cudaSuccess || !graph) {
    (void)cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    return rc ? 1 : fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
  }
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) { exec = nullptr; (void)cudaGetLastError(); return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
  e.graph_key[e.cur] = e.config_epoch;
  return 0;
}

// runs `groups` x kGraphIters iterations starting at first_iteration by replaying the captured graph
int replay_iterations(Engine& e, int first_iteration, int groups) {
  if (!e.graph_exec[e.cur] || e.graph_key[e.cur] != e.config_epoch) {
    if (capture_iterations(e)) return 1;
  }
  const int total = groups * kGraphIters;
  std::vector<double> table(size_t(total) * e.D);
  for (int i = 0; i < total; ++i)
    for (int d = 0; d < e.D; ++d)
      table[size_t(i) * e.D + d] = e.noise_stddev[d] * std::pow(e.noise_decay[d], first_iteration + i - 1);
  if (join_streams(e)) return 1;
  if (e.scale_table.n < table.size()) {
    CUDA_TRY(cudaStreamSynchronize(e.stream));
    CUDA_TRY(e.scale_table.alloc(std::max<size_t>(table.size(), size_t(1024) * e.D)));
    // the graphs hold the table's address: a new allocation needs new captures
    e.graph_key[0] = e.graph_key[1] = 0;
    if (capture_iterations(e)) return 1;
  }
  const int zero = 0;
  CUDA_TRY(cudaMemcpyAsync(e.scale_table.p, table.data(), table.size() * 8, cudaMemcpyHostToDevice, e.stream));
  CUDA_TRY(cudaMemcpyAsync(e.dev_generation.p, &e.generation, sizeof(uint32_t), cudaMemcpyHostToDevice, e.stream));
  CUDA_TRY(cudaMemcpyAsync(e.dev_table_index.p, &zero, sizeof(int), cudaMemcpyHostToDevice, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));   // the uploads read host stack memory
  for (int g = 0; g < groups; ++g) CUDA_TRY(cudaGraphLaunch(e.graph_exec[e.cur], e.stream));
  // the graphs run on the main stream only: whatever the tail stream is given next (the reuse selection of a following plain
  // iteration) must wait for them
  CUDA_TRY(cudaEventRecord(e.ev_upd, e.stream));
  CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_upd, 0));
  // the host mirror of what the replayed iterations did (PolicyImprovement bookkeeping of iterate_once)
  e.generation += uint32_t(total);
  e.num_gen = e.R - e.Rre;
  e.extra_added = true;
  e.steady_iterations += total;
  e.launches += int64_t(groups) * e.graph_launches;
  e.tail_dirty = false;   // the graph ends with the tail stream joined
  return 0;
}

// sticky flag of the peer-memory exchange (k_shard_stats): a rank that never arrived leaves theta without the update
int check_shard_error(Engine& e) {
  if (e.desc.rollout_shard_world <= 1 || !e.peers_open || !e.xchg_err.p) return 0;
  int err = 0;
  CUDA_TRY(cudaMemcpyAsync(&err, e.xchg_err.p, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return err ? fail("a peer exchange timed out: some rank did not reach the same iteration; the policy was not updated") : 0;
}

int fill_stats(Engine& e, stomp_iter_stats* stats) {
  if (!stats) return 0;
  if (join_streams(e)) return 1;
  if (check_shard_error(e)) return 1;
  stats->num_generated_rollouts = e.num_gen;
  if (stats->noiseless_cost)
    CUDA_TRY(cudaMemcpyAsync(stats->noiseless_cost, e.noiseless_sum.p, size_t(e.B) * 8, cudaMemcpyDeviceToHost, e.stream));
  if (stats->noiseless_collision_free) {
    CUDA_TRY(cudaMemcpy2DAsync(stats->noiseless_collision_free, 4, e.collision_free.p + e.R, size_t(e.R + 1) * 4, 4, e.B,
                               cudaMemcpyDeviceToHost, e.stream));
  }
  if (stats->noiseless_constraints_satisfied) {
    CUDA_TRY(cudaMemcpy2DAsync(stats->noiseless_constraints_satisfied, 4, e.constraints_ok.p + e.R, size_t(e.R + 1) * 4, 4, e.B,
                               cudaMemcpyDeviceToHost, e.stream));
  }
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int ensure_scratch(Engine& e, size_t n_rollouts_total) {
  if (e.scratch_n >= n_rollouts_total) return 0;
  CUDA_TRY(e.scratch_params.alloc(n_rollouts_total * e.D * e.N));
  CUDA_TRY(e.scratch_noise.alloc(n_rollouts_total * e.D * e.N));
  CUDA_TRY(e.scratch_costs.alloc(n_rollouts_total * e.D * e.N));
  CUDA_TRY(e.scratch_flags.alloc(n_rollouts_total));
  CUDA_TRY(e.scratch_cflags.alloc(n_rollouts_total));
  e.scratch_n = n_rollouts_total;
  return 0;
}

template <typename Real>
int upload_constraints(Engine& e) {
  std::vector<DevConstraint<Real>> dc(e.constraints.size());
  for (size_t i = 0; i < dc.size(); ++i) {
    const stomp_orientation_constraint& c = e.constraints[i];
    std::memset(&dc[i], 0, sizeof(dc[i]));
    const double* q = c.orientation;   // btMatrix3x3::setRotation(btQuaternion(x, y, z, w))
    const double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], s2 = 2.0 / d;
    const double xs = q[0] * s2, ys = q[1] * s2, zs = q[2] * s2, wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    const double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    const double nom[9] = {1.0 - (yy + zz), xy - wz, xz + wy, xy + wz, 1.0 - (xx + zz), yz - wx, xz - wy, yz + wx, 1.0 - (xx + yy)};
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) dc[i].nominal_inv[r * 3 + k] = Real(nom[k * 3 + r]);
    for (int k = 0; k < 9; ++k) dc[i].rel[k] = Real(e.robot.seg_rel[c.segment].R[k]);
    dc[i].node = e.robot.seg_node[c.segment];
    dc[i].body_fixed = c.body_fixed ? 1 : 0;
    const double tol[3] = {c.absolute_roll_tolerance, c.absolute_pitch_tolerance, c.absolute_yaw_tolerance};
    for (int k = 0; k < 3; ++k) {
      dc[i].tol[k] = Real(tol[k]);
      dc[i].w[k] = Real(tol[k] >= M_PI ? 0.0 : 1.0);   // src/constraint_evaluator.cpp:68-74
    }
    dc[i].weight = Real(c.weight);
  }
  static const unsigned char zeros[16] = {0};
  if (dc.empty() ? upload(e, e.dconstraints, zeros, sizeof(zeros))
                 : upload(e, e.dconstraints, reinterpret_cast<const unsigned char*>(dc.data()), dc.size() * sizeof(DevConstraint<Real>)))
    return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}
Engine* E(void* h) { return static_cast<Engine*>(h); }

// the bricked copy of the distance field k_cost gathers from (kernels.cuh, struct Sdf)
int build_bricks(Engine& e) {
  e.sdf.brick = nullptr;
  e.sdf.nby = (e.sdf.ny + 3) / 4;
  e.sdf.nbz = (e.sdf.nz + 1) / 2;
  const size_t esz = e.sdf.dtype == STOMP_VOXEL_U8_SQ ? 1 : e.sdf.dtype == STOMP_VOXEL_U16_SQ ? 2 : 4;
  const size_t cells = size_t((e.sdf.nx + 3) / 4) * e.sdf.nby * e.sdf.nbz * 32;
  if (cells >= (size_t(1) << 31)) return fail("distance field too large (padded to bricks of 4 x 4 x 2, cells must fit a 32-bit index)");
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.tail_stream));
  if (e.vox_brick.n != cells * esz) CUDA_TRY(e.vox_brick.alloc(cells * esz));
  const size_t n = size_t(e.sdf.nx) * e.sdf.ny * e.sdf.nz;
  const unsigned grid = unsigned(std::min<size_t>((n + 255) / 256, size_t(148) * 64));
  begin_launch(e);
  if (esz == 1) k_brick_relayout<uint8_t><<<grid, 256, 0, e.ws>>>(e.sdf.nx, e.sdf.ny, e.sdf.nz, e.sdf.nby, e.sdf.nbz, static_cast<const uint8_t*>(e.sdf.vox), reinterpret_cast<uint8_t*>(e.vox_brick.p));
  else if (esz == 2) k_brick_relayout<uint16_t><<<grid, 256, 0, e.ws>>>(e.sdf.nx, e.sdf.ny, e.sdf.nz, e.sdf.nby, e.sdf.nbz, static_cast<const uint16_t*>(e.sdf.vox), reinterpret_cast<uint16_t*>(e.vox_brick.p));
  else k_brick_relayout<float><<<grid, 256, 0, e.ws>>>(e.sdf.nx, e.sdf.ny, e.sdf.nz, e.sdf.nby, e.sdf.nbz, static_cast<const float*>(e.sdf.vox), reinterpret_cast<float*>(e.vox_brick.p));
  if (check_launch(e, "k_brick_relayout")) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.sdf.brick = e.vox_brick.p;
  return 0;
}


#define ENGINE_NOJOIN(h)                           \
  if (!(h)) return fail("null engine handle");     \
  Engine& e = *E(h);                               \
  CUDA_TRY(cudaSetDevice(e.device));

#define ENGINE_OR_FAIL(h)                          \
  ENGINE_NOJOIN(h)                                 \
  if (join_streams(e)) return 1;

}  // namespace

extern "C" {

const char* stomp_engine_last_error(void) { return g_error.c_str(); }
int stomp_engine_abi_version(void) { return 1; }
const char* stomp_engine_build_info(void) { return "stomp_b200 CUDA engine, sm_100a, " __DATE__ " " __TIME__; }

int stomp_engine_create(const stomp_engine_desc* desc, void** out_engine) {
  if (!desc || !out_engine) return fail("null argument");
  *out_engine = nullptr;
  if (desc->num_dimensions < 1 || desc->num_time_steps < 2 || desc->num_rollouts < 1 || desc->num_problems < 1)
    return fail("num_dimensions, num_rollouts, num_problems must be >= 1 and num_time_steps >= 2");
  if (desc->num_reused_rollouts < 0 || desc->num_reused_rollouts >= desc->num_rollouts)
    return fail("Number of reused rollouts must be strictly less than number of rollouts.");  // policy_improvement.cpp:102-106
  if (desc->dtype != STOMP_F64 && desc->dtype != STOMP_F32) return fail("unknown dtype");
  if (desc->sdf_mode != STOMP_SDF_NEAREST && desc->sdf_mode != STOMP_SDF_TRILINEAR) return fail("unknown sdf_mode");
  if (desc->rollout_shard_world < 1 || desc->rollout_shard_rank < 0 || desc->rollout_shard_rank >= desc->rollout_shard_world)
    return fail("bad rollout shard rank / world");
  if (desc->rollout_shard_world > 1 && (desc->num_reused_rollouts != 0 || desc->num_problems != 1))
    return fail("rollout sharding requires num_reused_rollouts == 0 and num_problems == 1");
  if (desc->discretization <= 0.0 || desc->movement_duration <= 0.0) return fail("durations must be positive");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return fail("no CUDA device available: the STOMP B200 engine has no CPU fallback");
  }
  if (desc->device < 0 || desc->device >= ndev) return fail("CUDA device ordinal out of range");
  Engine* ep = new Engine();
  Engine& e = *ep;
  e.desc = *desc;
  e.B = desc->num_problems; e.D = desc->num_dimensions; e.N = desc->num_time_steps; e.R = desc->num_rollouts;
  e.Rre = desc->num_reused_rollouts;
  e.device = desc->device;
  e.f32 = desc->dtype == STOMP_F32;
  std::string err;
  if (!sh::build_policy_matrices(*desc, e.pm, err)) { stomp_engine_destroy(ep); return fail(err); }
  if (e.pm.chol.hb > kMaxHb) { stomp_engine_destroy(ep); return fail("control cost bandwidth too large"); }
  auto bail = [&](cudaError_t c, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(c);
    (void)cudaGetLastError();
    stomp_engine_destroy(ep);
    return fail(m);
  };
  cudaError_t c;
  if ((c = cudaSetDevice(e.device)) != cudaSuccess) return bail(c, "cudaSetDevice");
  if ((c = cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
  if ((c = cudaDeviceGetAttribute(&e.num_sms, cudaDevAttrMultiProcessorCount, e.device)) != cudaSuccess) return bail(c, "cudaDeviceGetAttribute");
  if ((c = cudaEventCreate(&e.ev0)) != cudaSuccess || (c = cudaEventCreate(&e.ev1)) != cudaSuccess) return bail(c, "cudaEventCreate");
  if ((c = cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
  // equal priorities: with the tail stream at the greatest priority the timeline looks better under event recording
  // (k_extra_total no longer waits 157 us for a CTA slot) but the unrecorded loop is slower, 0.522 vs 0.510 ms (profiles/README.md)
  if ((c = cudaStreamCreateWithFlags(&e.tail_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
  if ((c = cudaStreamCreateWithFlags(&e.pre_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
  if ((c = cudaStreamCreateWithFlags(&e.cand_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
  if ((c = cudaEventCreateWithFlags(&e.ev_cand, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_nl, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_extra, cudaEventDisableTiming)) != cudaSuccess)
    return bail(c, "cudaEventCreate");
  if ((c = cudaEventCreateWithFlags(&e.ev_pre, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_fin, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_cost, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_tail, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_upd, cudaEventDisableTiming)) != cudaSuccess ||
      (c = cudaEventCreateWithFlags(&e.ev_selected, cudaEventDisableTiming)) != cudaSuccess)
    return bail(c, "cudaEventCreate");
  if ((c = cudaStreamCreateWithFlags(&e.result_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(c, "cudaStreamCreate");
  for (int i = 0; i < 2; ++i)
    if ((c = cudaEventCreateWithFlags(&e.ev_snap_main[i], cudaEventDisableTiming)) != cudaSuccess ||
        (c = cudaEventCreateWithFlags(&e.ev_snap_tail[i], cudaEventDisableTiming)) != cudaSuccess ||
        (c = cudaEventCreateWithFlags(&e.ev_results[i], cudaEventDisableTiming)) != cudaSuccess)
      return bail(c, "cudaEventCreate");
  e.ws = e.stream;
  e.overlap = !(getenv("STOMP_NO_OVERLAP") && atoi(getenv("STOMP_NO_OVERLAP")) != 0);
  e.cull_allowed = !(getenv("STOMP_NO_CULL") && atoi(getenv("STOMP_NO_CULL")) != 0);   // A/B switch of the k_cost broad phase
  for (int i = 0; i < 2; ++i)
    if ((c = cudaEventCreateWithFlags(&e.ev_copy_done[i], cudaEventDisableTiming)) != cudaSuccess ||
        (c = cudaEventCreateWithFlags(&e.ev_consumed[i], cudaEventDisableTiming)) != cudaSuccess)
      return bail(c, "cudaEventCreate");
  const size_t BDN = size_t(e.B) * e.D * e.N, BRDN = BDN * e.R, BRN = size_t(e.B) * e.R * e.N;
#define ALLOC(buf, n) if ((c = (buf).alloc(n)) != cudaSuccess) return bail(c, "cudaMalloc " #buf)
  ALLOC(e.theta, BDN); ALLOC(e.pad_start, size_t(e.B) * e.D); ALLOC(e.pad_goal, size_t(e.B) * e.D);
  ALLOC(e.params[0], BRDN); ALLOC(e.params[1], BRDN); ALLOC(e.state[0], BRN); ALLOC(e.state[1], BRN);
  ALLOC(e.csum, size_t(e.B) * e.R * e.D);
  ALLOC(e.noise, BRDN); ALLOC(e.control[0], BRDN); ALLOC(e.control[1], BRDN); ALLOC(e.cumulative, BRDN); ALLOC(e.totals, size_t(e.B) * (e.R + 1));
  if (desc->keep_intermediates) { ALLOC(e.noise_projected, BRDN); ALLOC(e.probabilities, BRDN); ALLOC(e.clipped, BRDN); }
  ALLOC(e.extra_state, size_t(e.B) * e.N); ALLOC(e.extra_control, BDN); ALLOC(e.updates, BDN); ALLOC(e.noiseless_sum, size_t(e.B));
  ALLOC(e.reuse_src, size_t(e.B) * std::max(1, e.Rre)); ALLOC(e.collision_free, size_t(e.B) * (e.R + 1));
  ALLOC(e.constraints_ok, size_t(e.B) * (e.R + 1));
  ALLOC(e.noise_scale, size_t(e.D)); ALLOC(e.dev_generation, 1); ALLOC(e.dev_table_index, 1);
  ALLOC(e.extra_clipped, BDN); ALLOC(e.best_traj, BDN); ALLOC(e.best_cost, size_t(e.B));
  ALLOC(e.track_state, size_t(e.B) * sizeof(TrackState)); ALLOC(e.num_done, 1);
  ALLOC(e.part, size_t(kChunks) * 2 * e.D * e.N); ALLOC(e.minmax, size_t(2) * e.D * e.N); ALLOC(e.sums, size_t(2) * e.D * e.N);
  ALLOC(e.shard_counters, size_t(2) + size_t(e.D) * e.N / 128 + 1); ALLOC(e.xchg_err, 1);
#undef ALLOC
  // matrices
  std::vector<double> qt(size_t(e.N) * e.N);
  for (int i = 0; i < e.N; ++i)
    for (int j = 0; j < e.N; ++j) qt[size_t(j) * e.N + i] = e.pm.Qinv(i, j);
  std::vector<double> fw(size_t(e.N) * 8, 0.0), bw(size_t(e.N) * 8, 0.0);
  {
    const int hb = e.pm.chol.hb;
    for (int i = 0; i < e.N; ++i) {
      const double inv = e.pm.chol.inv_diag[i];
      fw[size_t(i) * 8] = bw[size_t(i) * 8] = inv;
      for (int j = 1; j <= hb; ++j) {
        if (i - j >= 0) fw[size_t(i) * 8 + j] = e.pm.chol.band[size_t(i) * (hb + 1) + j] * inv;
        if (i + j < e.N) bw[size_t(i) * 8 + j] = e.pm.chol.band[size_t(i + j) * (hb + 1) + j] * inv;
      }
    }
  }
  if (upload(e, e.band_fw, fw.data(), fw.size()) || upload(e, e.band_bw, bw.data(), bw.size()) ||
      upload(e, e.proj_scale, e.pm.proj_scale.data(), size_t(e.N)) || upload(e, e.qinv_t, qt.data(), qt.size())) {
    stomp_engine_destroy(ep);
    return 1;
  }
  {
    // spike tables of k_generate_seg (kernels.cuh): response of a segment's rows to unit values of the neighbouring segment's
    // six boundary entries, by the solves' own recurrences
    const int N = e.N;
    e.seg_P = std::max(2, std::min(8, N / 24));
    if (const char* sp = getenv("STOMP_SEG_P")) e.seg_P = std::max(2, std::min(8, atoi(sp)));
    if (N / e.seg_P >= 16) {
      const int P = e.seg_P;
      std::vector<double> hf(size_t(N) * 6, 0.0), hbk(size_t(N) * 6, 0.0);
      for (int p = 0; p < P; ++p) {
        const int s0 = seg_start(p, N, P), s1 = seg_start(p + 1, N, P);
        if (p > 0)
          for (int i = s0; i < s1; ++i)
            for (int j = 0; j < 6; ++j) {
              double acc = 0.0;
              for (int k = 1; k <= 6; ++k) {
                const int rr = i - k;
                if (rr < 0) continue;
                const double tv = rr >= s0 ? hf[size_t(rr) * 6 + j] : (s0 - 1 - rr == j ? 1.0 : 0.0);
                acc -= fw[size_t(i) * 8 + k] * tv;
              }
              hf[size_t(i) * 6 + j] = acc;
            }
        if (p < P - 1)
          for (int i = s1 - 1; i >= s0; --i)
            for (int j = 0; j < 6; ++j) {
              double acc = 0.0;
              for (int k = 1; k <= 6; ++k) {
                const int rr = i + k;
                if (rr >= N) continue;
                const double tv = rr < s1 ? hbk[size_t(rr) * 6 + j] : (rr - s1 == j ? 1.0 : 0.0);
                acc -= bw[size_t(i) * 8 + k] * tv;
              }
              hbk[size_t(i) * 6 + j] = acc;
            }
      }
      if (upload(e, e.seg_hf, hf.data(), hf.size()) || upload(e, e.seg_hb, hbk.data(), hbk.size())) {
        stomp_engine_destroy(ep);
        return 1;
      }
    } else {
      e.seg_P = 0;
    }
  }
  {
    // dense forms of the two solves for the small-batch kernel: C^-1 (column k = forward substitution of e_k) and R^-1 diag(s)
    const int N = e.N, hb = e.pm.chol.hb;
    std::vector<double> cinv(size_t(N) * N, 0.0), msd(size_t(N) * N, 0.0), x(N);
    for (int k = 0; k < N; ++k) {
      std::fill(x.begin(), x.end(), 0.0);
      for (int i = k; i < N; ++i) {
        double s = i == k ? 1.0 : 0.0;
        for (int j = 1; j <= hb && i - j >= k; ++j) s -= e.pm.chol.band[size_t(i) * (hb + 1) + j] * x[i - j];
        x[i] = s * e.pm.chol.inv_diag[i];
        cinv[size_t(i) * N + k] = x[i];
      }
    }
    for (int j = 0; j < N; ++j)
      for (int t = 0; t < N; ++t) msd[size_t(j) * N + t] = e.pm.Rinv(j, t) * e.pm.proj_scale[j];
    if (N <= 1024) {
      // [pass][Np][strideA]: pass = t / 128, rows j, columns t % 128 with the staged row stride (kernels.cuh, mma_gemm)
      const int Np = (N + 7) & ~7, strideA = mma_a_stride(std::min(Np, kMmaGroup)), passes = (Np + kMmaGroup - 1) / kMmaGroup;
      std::vector<double> p1(size_t(passes) * Np * strideA, 0.0), p2(size_t(passes) * Np * strideA, 0.0);
      for (int j = 0; j < N; ++j)
        for (int t = 0; t < N; ++t) {
          const size_t at = (size_t(t / kMmaGroup) * Np + j) * strideA + (t % kMmaGroup);
          p1[at] = cinv[size_t(j) * N + t];   // zero above the diagonal of C^-1: j < t
          p2[at] = msd[size_t(j) * N + t];
        }
      if (upload(e, e.mma_a1, p1.data(), p1.size()) || upload(e, e.mma_a2, p2.data(), p2.size())) {
        stomp_engine_destroy(ep);
        return 1;
      }
    }
    if (upload(e, e.dense_cinv, cinv.data(), cinv.size()) || upload(e, e.dense_ms, msd.data(), msd.size())) {
      stomp_engine_destroy(ep);
      return 1;
    }
  }
  // A/B switches of the generation kernels: STOMP_GENERATE=band|dense forces one, STOMP_NO_DENSE=1 is "band"
  if (const char* g = getenv("STOMP_GENERATE")) e.gen_mode = !strcmp(g, "band") ? 1 : !strcmp(g, "dense") ? 2 : !strcmp(g, "mma") ? 3 : !strcmp(g, "seg") ? 4 : 0;
  if (getenv("STOMP_NO_DENSE") && atoi(getenv("STOMP_NO_DENSE")) != 0) e.gen_mode = 1;
  if (getenv("STOMP_GRAPH")) e.graph_mode = atoi(getenv("STOMP_GRAPH")) != 0 ? 1 : 0;
  e.chain_probe = getenv("STOMP_CHAIN_PROBE") && atoi(getenv("STOMP_CHAIN_PROBE")) != 0;
  if (e.chain_probe && (cudaEventCreate(&e.probe_main) != cudaSuccess || cudaEventCreate(&e.probe_tail) != cudaSuccess ||
                        cudaEventCreate(&e.probe_upd) != cudaSuccess))
    e.chain_probe = false;
  e.dense_update = !(getenv("STOMP_NO_DENSE_UPDATE") && atoi(getenv("STOMP_NO_DENSE_UPDATE")) != 0);
  e.dmma_update = !(getenv("STOMP_NO_DMMA") && atoi(getenv("STOMP_NO_DMMA")) != 0);
  e.direct_update = !(getenv("STOMP_NO_DIRECT_UPDATE") && atoi(getenv("STOMP_NO_DIRECT_UPDATE")) != 0);
  e.lookahead = !(getenv("STOMP_NO_LOOKAHEAD") && atoi(getenv("STOMP_NO_LOOKAHEAD")) != 0);
  e.scales_in_args = getenv("STOMP_SCALES_IN_ARGS") && atoi(getenv("STOMP_SCALES_IN_ARGS")) != 0;
  if (const char* ov = getenv("STOMP_UPDATE_DPC")) e.update_dpc = atoi(ov);
  e.offchain_extra = !(getenv("STOMP_NO_OFFCHAIN_EXTRA") && atoi(getenv("STOMP_NO_OFFCHAIN_EXTRA")) != 0);
  e.use_totals_kernel = !(getenv("STOMP_NO_TOTALS_KERNEL") && atoi(getenv("STOMP_NO_TOTALS_KERNEL")) != 0);
  if (const char* cp = getenv("STOMP_CUM_PLACEMENT")) e.cum_placement = !strcmp(cp, "chain") ? 1 : !strcmp(cp, "late") ? 2 : 0;
  e.wide_update = !(getenv("STOMP_NO_WIDE_UPDATE") && atoi(getenv("STOMP_NO_WIDE_UPDATE")) != 0);
  if (const char* dm = getenv("STOMP_SMALL_BATCH_MAX")) e.small_max = atoll(dm);   // elements B R D N; 0: throughput schedule always
  e.split_cost = !(getenv("STOMP_NO_SPLIT_COST") && atoi(getenv("STOMP_NO_SPLIT_COST")) != 0);
  e.noise_stddev.assign(e.D, 1.0);
  e.noise_decay.assign(e.D, 1.0);
  std::vector<int> hl(e.D, 0);
  std::vector<double> zeros(e.D, 0.0);
  if (upload(e, e.has_limits, hl.data(), size_t(e.D)) || upload(e, e.limit_min, zeros.data(), size_t(e.D)) ||
      upload(e, e.limit_max, zeros.data(), size_t(e.D))) {
    stomp_engine_destroy(ep);
    return 1;
  }
  if ((c = cudaStreamSynchronize(e.stream)) != cudaSuccess) return bail(c, "cudaStreamSynchronize");
  e.control_cost_weight = desc->smoothness_cost_weight;
  *out_engine = ep;
  return 0;
}

int stomp_engine_destroy(void* h) {
  if (!h) return 0;
  Engine* e = E(h);
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  if (e->tail_stream) { cudaStreamSynchronize(e->tail_stream); cudaStreamDestroy(e->tail_stream); }
  if (e->pre_stream) { cudaStreamSynchronize(e->pre_stream); cudaStreamDestroy(e->pre_stream); }
  if (e->cand_stream) { cudaStreamSynchronize(e->cand_stream); cudaStreamDestroy(e->cand_stream); }
  if (e->ev_cand) cudaEventDestroy(e->ev_cand);
  if (e->ev_nl) cudaEventDestroy(e->ev_nl);
  if (e->ev_extra) cudaEventDestroy(e->ev_extra);
  if (e->ev_pre) cudaEventDestroy(e->ev_pre);
  if (e->ev_fin) cudaEventDestroy(e->ev_fin);
  if (e->result_stream) { cudaStreamSynchronize(e->result_stream); cudaStreamDestroy(e->result_stream); }
  for (int i = 0; i < 2; ++i) {
    if (e->ev_snap_main[i]) cudaEventDestroy(e->ev_snap_main[i]);
    if (e->ev_snap_tail[i]) cudaEventDestroy(e->ev_snap_tail[i]);
    if (e->ev_results[i]) cudaEventDestroy(e->ev_results[i]);
  }
  for (int i = 0; i < 2; ++i)
    if (e->graph_exec[i]) cudaGraphExecDestroy(e->graph_exec[i]);
  if (e->ev_tail) cudaEventDestroy(e->ev_tail);
  if (e->ev_cost) cudaEventDestroy(e->ev_cost);
  if (e->peers_open)
    for (int r = 0; r < e->desc.rollout_shard_world; ++r)
      if (r != e->desc.rollout_shard_rank && e->peer_base[r]) cudaIpcCloseMemHandle(e->peer_base[r]);
  if (e->ev_upd) cudaEventDestroy(e->ev_upd);
  if (e->ev_selected) cudaEventDestroy(e->ev_selected);
  if (e->chain_probe) {
    if (e->probe_n)
      std::fprintf(stderr, "stomp_b200 chain probe: tail chain ends %+.1f us after the main stream's k_cumulative, k_update ends "
                           "%.1f us after it (mean of %lld sampled iterations)\n",
                   e->probe_sum_us / e->probe_n, e->probe_upd_sum_us / e->probe_n, e->probe_n);
    cudaEventDestroy(e->probe_main); cudaEventDestroy(e->probe_tail); cudaEventDestroy(e->probe_upd);
  }
  for (int i = 0; i < 2; ++i) {
    if (e->ev_copy_done[i]) cudaEventDestroy(e->ev_copy_done[i]);
    if (e->ev_consumed[i]) cudaEventDestroy(e->ev_consumed[i]);
  }
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (cudaEvent_t ev : e->prof_a) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->prof_b) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return 0;
}

int stomp_engine_set_robot(void* h, const stomp_segment* segments, int32_t num_segments, int32_t reference_segment,
                           const stomp_sphere* spheres, int32_t num_spheres, const stomp_joint_limit* limits) {
  ENGINE_OR_FAIL(h);
  if (!segments || num_segments < 1 || (!spheres && num_spheres > 0)) return fail("null robot tables");
  std::string err;
  if (!sh::fold_robot(segments, num_segments, reference_segment, spheres, num_spheres, e.D, e.robot, err)) return fail(err);
  if (e.robot.num_slots > kMaxSlots) return fail("kinematic tree branches too deeply for the FK kernel (frame slots)");
  e.K = num_spheres;
  e.num_nodes = int(e.robot.nodes.size());
  e.raw_segments.assign(segments, segments + num_segments);
  e.chain_len = 0;          // segment numbering may have changed: the dynamics have to be set again
  e.torque_weight = 0.0;
  if (e.f32 ? upload_robot_tables<float>(e) : upload_robot_tables<double>(e)) return 1;
  std::vector<int> hl(e.D, 0);
  std::vector<double> lo(e.D, 0.0), hi(e.D, 0.0);
  if (limits)
    for (int d = 0; d < e.D; ++d) hl[d] = limits[d].has_limits, lo[d] = limits[d].min, hi[d] = limits[d].max;
  if (upload(e, e.has_limits, hl.data(), size_t(e.D)) || upload(e, e.limit_min, lo.data(), size_t(e.D)) ||
      upload(e, e.limit_max, hi.data(), size_t(e.D)))
    return 1;
  e.constraints.clear();   // segment numbering may have changed
  CUDA_TRY(e.debug.alloc(size_t(e.N + 3) * std::max(1, e.K)));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.have_robot = true;
  ++e.config_epoch;
  return 0;
}

int stomp_engine_set_sdf(void* h, const void* voxels, int32_t nx, int32_t ny, int32_t nz, const double origin[3],
                         double resolution, int32_t voxel_dtype) {
  ENGINE_OR_FAIL(h);
  if (!voxels || nx < 3 || ny < 3 || nz < 3 || resolution <= 0.0) return fail("bad distance field");
  size_t esz = voxel_dtype == STOMP_VOXEL_U8_SQ ? 1 : voxel_dtype == STOMP_VOXEL_U16_SQ ? 2 : voxel_dtype == STOMP_VOXEL_F32 ? 4 : 0;
  if (!esz) return fail("unknown voxel dtype");
  if (size_t(nx) * ny * nz >= (size_t(1) << 31)) return fail("distance field too large (cells must fit a 32-bit index)");
  size_t bytes = size_t(nx) * ny * nz * esz;
  if (upload(e, e.vox, static_cast<const unsigned char*>(voxels), bytes)) return 1;
  e.sdf.vox = e.vox.p;
  e.sdf.nx = nx; e.sdf.ny = ny; e.sdf.nz = nz; e.sdf.dtype = voxel_dtype;
  for (int i = 0; i < 3; ++i) e.sdf.origin[i] = origin[i];
  e.sdf.res = resolution;
  e.sdf.inv_res = 1.0 / resolution;
  if (e.f32 ? upload_sqrt_table<float>(e) : upload_sqrt_table<double>(e)) return 1;
  if (build_bricks(e)) return 1;
  e.have_sdf = true;
  e.cull_dirty = true;
  ++e.config_epoch;
  return 0;
}

namespace {
// lattice of one axis exactly like the reference's `for (double x = low; x <= low + extent + res; x += res)`
void append_lattice(std::vector<double>& lat, double low, double extent, double res, int& offset, int& count) {
  offset = int(lat.size());
  for (double x = low; x <= low + extent + res; x += res) lat.push_back(x);
  count = int(lat.size()) - offset;
}
void quaternion_to_rotation(const double q[4], double R[9]) {   // KDL::Rotation::Quaternion(x, y, z, w)
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * x * y - 2 * w * z; R[2] = 2 * x * z + 2 * w * y;
  R[3] = 2 * x * y + 2 * w * z; R[4] = w2 - x2 + y2 - z2; R[5] = 2 * y * z - 2 * w * x;
  R[6] = 2 * x * z - 2 * w * y; R[7] = 2 * y * z + 2 * w * x; R[8] = w2 - x2 - y2 + z2;
}
// Convex hull of n points (incremental, O(n x faces)): triangles with outward normals as vertex index triples.  What
// bodies::ConvexMesh gets from qhull; coplanar input points may be triangulated differently, which changes no interior.
struct HullTri { int a, b, c; };
bool convex_hull_3d(const double* v, int n, std::vector<HullTri>& out) {
  auto P = [&](int i) { return v + size_t(i) * 3; };
  auto sub = [](const double* p, const double* q, double r[3]) { r[0] = p[0] - q[0]; r[1] = p[1] - q[1]; r[2] = p[2] - q[2]; };
  auto cross = [](const double* p, const double* q, double r[3]) {
    r[0] = p[1] * q[2] - p[2] * q[1]; r[1] = p[2] * q[0] - p[0] * q[2]; r[2] = p[0] * q[1] - p[1] * q[0];
  };
  auto dot = [](const double* p, const double* q) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]; };
  if (n < 4) return false;
  double lo[3] = {P(0)[0], P(0)[1], P(0)[2]}, hi[3] = {lo[0], lo[1], lo[2]};
  for (int i = 1; i < n; ++i)
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], P(i)[k]); hi[k] = std::max(hi[k], P(i)[k]); }
  const double extent = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
  if (!(extent > 0.0)) return false;
  const double eps = 1e-10 * extent;
  // starting tetrahedron: extreme point, farthest from it, farthest from that line, farthest from that plane
  int i0 = 0;
  for (int i = 1; i < n; ++i) if (P(i)[0] < P(i0)[0]) i0 = i;
  int i1 = -1; double best = 0.0;
  for (int i = 0; i < n; ++i) { double d[3]; sub(P(i), P(i0), d); const double l = dot(d, d); if (l > best) { best = l; i1 = i; } }
  if (i1 < 0) return false;
  double e01[3]; sub(P(i1), P(i0), e01);
  int i2 = -1; best = 0.0;
  for (int i = 0; i < n; ++i) { double d[3], c[3]; sub(P(i), P(i0), d); cross(e01, d, c); const double l = dot(c, c); if (l > best) { best = l; i2 = i; } }
  if (i2 < 0 || std::sqrt(best) <= eps * std::sqrt(dot(e01, e01))) return false;
  double e02[3], nrm[3]; sub(P(i2), P(i0), e02); cross(e01, e02, nrm);
  const double nl = std::sqrt(dot(nrm, nrm));
  int i3 = -1; best = 0.0;
  for (int i = 0; i < n; ++i) { double d[3]; sub(P(i), P(i0), d); const double l = std::fabs(dot(nrm, d)) / nl; if (l > best) { best = l; i3 = i; } }
  if (i3 < 0 || best <= eps) return false;   // flat mesh: no volume to voxelise
  double inner[3];
  for (int k = 0; k < 3; ++k) inner[k] = 0.25 * (P(i0)[k] + P(i1)[k] + P(i2)[k] + P(i3)[k]);
  struct Face { int a, b, c; double n[3]; double off; bool alive; };
  std::vector<Face> faces;
  auto add_face = [&](int a, int b, int c) {
    Face f; f.a = a; f.b = b; f.c = c; f.alive = true;
    double u[3], w[3]; sub(P(b), P(a), u); sub(P(c), P(a), w); cross(u, w, f.n);
    const double l = std::sqrt(dot(f.n, f.n));
    if (l > 0.0) { f.n[0] /= l; f.n[1] /= l; f.n[2] /= l; }
    f.off = dot(f.n, P(a));
    if (dot(f.n, inner) - f.off > 0.0) {   // orient outward
      std::swap(f.b, f.c);
      f.n[0] = -f.n[0]; f.n[1] = -f.n[1]; f.n[2] = -f.n[2]; f.off = -f.off;
    }
    faces.push_back(f);
  };
  add_face(i0, i1, i2); add_face(i0, i1, i3); add_face(i0, i2, i3); add_face(i1, i2, i3);
  std::vector<char> visible;
  std::vector<std::pair<int, int>> horizon;
  for (int p = 0; p < n; ++p) {
    if (p == i0 || p == i1 || p == i2 || p == i3) continue;
    visible.assign(faces.size(), 0);
    bool any = false;
    for (size_t f = 0; f < faces.size(); ++f)
      if (faces[f].alive && dot(faces[f].n, P(p)) - faces[f].off > eps) { visible[f] = 1; any = true; }
    if (!any) continue;   // inside the current hull
    horizon.clear();
    for (size_t f = 0; f < faces.size(); ++f) {
      if (!visible[f]) continue;
      const int ed[3][2] = {{faces[f].a, faces[f].b}, {faces[f].b, faces[f].c}, {faces[f].c, faces[f].a}};
      for (int k = 0; k < 3; ++k) {
        bool shared = false;   // is the reversed edge part of another visible face?
        for (size_t g = 0; g < faces.size() && !shared; ++g) {
          if (!visible[g] || g == f) continue;
          const int eg[3][2] = {{faces[g].a, faces[g].b}, {faces[g].b, faces[g].c}, {faces[g].c, faces[g].a}};
          for (int m = 0; m < 3; ++m) shared |= eg[m][0] == ed[k][1] && eg[m][1] == ed[k][0];
        }
        if (!shared) horizon.emplace_back(ed[k][0], ed[k][1]);
      }
    }
    for (size_t f = 0; f < visible.size(); ++f) if (visible[f]) faces[f].alive = false;
    for (const auto& ed : horizon) add_face(ed.first, ed.second, p);
    // compact now and then: the visibility scans are linear in the face list
    if (faces.size() > 4096) {
      std::vector<Face> keep;
      for (const Face& f : faces) if (f.alive) keep.push_back(f);
      faces.swap(keep);
    }
  }
  out.clear();
  for (const Face& f : faces) if (f.alive) out.push_back(HullTri{f.a, f.b, f.c});
  return out.size() >= 4;
}
}  // namespace

int stomp_engine_build_sdf(void* h, const double size[3], const double origin[3], double resolution, double max_distance,
                           const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders) {
  return stomp_engine_build_sdf_points(h, size, origin, resolution, max_distance, boxes, num_boxes, cylinders, num_cylinders, nullptr, 0);
}

int stomp_engine_build_sdf_points(void* h, const double size[3], const double origin[3], double resolution, double max_distance,
                                  const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                  const double* points, int64_t num_points) {
  return stomp_engine_build_sdf_bodies(h, size, origin, resolution, max_distance, boxes, num_boxes, cylinders, num_cylinders, points,
                                       num_points, nullptr, 0);
}

int stomp_engine_build_sdf_bodies(void* h, const double size[3], const double origin[3], double resolution, double max_distance,
                                  const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                  const double* points, int64_t num_points, const stomp_body* bodies, int32_t num_bodies) {
  return stomp_engine_build_sdf_meshes(h, size, origin, resolution, max_distance, boxes, num_boxes, cylinders, num_cylinders, points,
                                       num_points, bodies, num_bodies, nullptr, 0);
}

int stomp_engine_build_sdf_meshes(void* h, const double size[3], const double origin[3], double resolution, double max_distance,
                                  const stomp_box* boxes, int32_t num_boxes, const stomp_cylinder* cylinders, int32_t num_cylinders,
                                  const double* points, int64_t num_points, const stomp_body* bodies, int32_t num_bodies,
                                  const stomp_mesh_body* meshes, int32_t num_meshes) {
  ENGINE_OR_FAIL(h);
  if (num_meshes < 0 || (num_meshes > 0 && !meshes)) return fail("bad mesh bodies");
  if (num_bodies < 0 || (num_bodies > 0 && !bodies)) return fail("bad bodies");
  if (num_points < 0 || (num_points > 0 && !points)) return fail("bad collision-map points");
  if (!size || !origin || resolution <= 0.0 || max_distance <= 0.0) return fail("bad distance field specification");
  if ((num_boxes > 0 && !boxes) || (num_cylinders > 0 && !cylinders) || num_boxes < 0 || num_cylinders < 0) return fail("bad collision objects");
  const int nx = int(size[0] / resolution), ny = int(size[1] / resolution), nz = int(size[2] / resolution);
  if (nx < 3 || ny < 3 || nz < 3) return fail("distance field too small");
  const size_t cells = size_t(nx) * ny * nz;
  if (cells >= (size_t(1) << 31)) return fail("distance field too large (cells must fit a 32-bit index)");
  const int cap = int(std::ceil(max_distance / resolution));
  if (cap < 1 || cap > 255) return fail("max_distance / resolution out of range");
  std::vector<SdfShape> shapes;
  std::vector<double> lattice;
  long long total = 0;
  auto add = [&](const double pos[3], const double quat[4], double ex, double ey, double ez, double lowx, double lowy, double lowz,
                 double radius) {
    SdfShape s;
    std::memset(&s, 0, sizeof(s));
    for (int i = 0; i < 3; ++i) s.position[i] = pos[i];
    quaternion_to_rotation(quat, s.R);
    s.radius = radius;
    append_lattice(lattice, lowx, ex, resolution, s.x_off, s.nx);
    append_lattice(lattice, lowy, ey, resolution, s.y_off, s.ny);
    append_lattice(lattice, lowz, ez, resolution, s.z_off, s.nz);
    s.first = total;
    total += (long long)s.nx * s.ny * s.nz;
    shapes.push_back(s);
  };
  for (int i = 0; i < num_boxes; ++i) {   // src/stomp_collision_space.cpp:275-291
    const stomp_box& b = boxes[i];
    add(b.position, b.orientation, b.dimensions[0], b.dimensions[1], b.dimensions[2], b.position[0] - b.dimensions[0] / 2.0,
        b.position[1] - b.dimensions[1] / 2.0, b.position[2] - b.dimensions[2] / 2.0, 0.0);
  }
  for (int i = 0; i < num_cylinders; ++i) {   // src/stomp_collision_space.cpp:238-269
    const stomp_cylinder& c = cylinders[i];
    if (!(c.radius > 0.0)) return fail("cylinder radius must be positive");
    add(c.position, c.orientation, c.radius * 2.0, c.radius * 2.0, c.height, c.position[0] - c.radius, c.position[1] - c.radius,
        c.position[2] - c.height / 2.0, c.radius);
  }
  DevBuf<unsigned char> occ, dshapes;
  DevBuf<double> dlat;
  DevBuf<uint16_t> g1, g2;
  CUDA_TRY(occ.alloc(cells));
  CUDA_TRY(g1.alloc(cells));
  CUDA_TRY(g2.alloc(cells));
  if (!shapes.empty()) {
    if (upload(e, dshapes, reinterpret_cast<const unsigned char*>(shapes.data()), shapes.size() * sizeof(SdfShape)) ||
        upload(e, dlat, lattice.data(), lattice.size()))
      return 1;
    const unsigned grid = unsigned(std::min<long long>((total + 255) / 256, 148 * 32));
    begin_launch(e);
    k_sdf_mark<<<std::max(1u, grid), 256, 0, e.ws>>>(int(shapes.size()), total, reinterpret_cast<const SdfShape*>(dshapes.p), dlat.p,
                                                        origin[0], origin[1], origin[2], resolution, nx, ny, nz, occ.p);
    if (check_launch(e, "k_sdf_mark")) return 1;
  }
  if (num_points > 0) {
    DevBuf<double> dpts;
    if (upload(e, dpts, points, size_t(num_points) * 3)) return 1;
    begin_launch(e);
    k_sdf_mark_points<<<unsigned(std::min<long long>((num_points + 255) / 256, 148 * 32)), 256, 0, e.ws>>>(
        num_points, dpts.p, origin[0], origin[1], origin[2], resolution, nx, ny, nz, occ.p);
    if (check_launch(e, "k_sdf_mark_points")) return 1;
    CUDA_TRY(cudaStreamSynchronize(e.stream));   // dpts goes out of scope
  }
  if (num_bodies > 0) {
    std::vector<SdfBody> hb;
    long long btotal = 0;
    for (int i = 0; i < num_bodies; ++i) {
      const stomp_body& sb = bodies[i];
      SdfBody b;
      std::memset(&b, 0, sizeof(b));
      double R[9];
      quaternion_to_rotation(sb.orientation, R);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) b.Rt[r * 3 + c] = R[c * 3 + r];
      double bound = 0.0;   // bodies::*::computeBoundingSphere of the scaled + padded shape
      if (sb.type == STOMP_BODY_SPHERE) {
        b.p[0] = sb.dimensions[0] * sb.scale + sb.padding;
        bound = b.p[0];
      } else if (sb.type == STOMP_BODY_BOX) {
        for (int k = 0; k < 3; ++k) b.p[k] = sb.dimensions[k] / 2.0 * sb.scale + sb.padding;
        bound = std::sqrt(b.p[0] * b.p[0] + b.p[1] * b.p[1] + b.p[2] * b.p[2]);
      } else if (sb.type == STOMP_BODY_CYLINDER) {
        b.p[0] = sb.dimensions[0] * sb.scale + sb.padding;
        b.p[1] = sb.dimensions[1] / 2.0 * sb.scale + sb.padding;
        bound = std::sqrt(b.p[0] * b.p[0] + b.p[1] * b.p[1]);
      } else {
        return fail("unknown body type (mesh bodies go through stomp_engine_build_sdf_meshes)");
      }
      if (!(bound > 0.0)) return fail("body with a non-positive size");
      b.type = sb.type;
      b.first = btotal;
      for (int k = 0; k < 3; ++k) {
        b.c[k] = sb.position[k];
        // worldToGrid(centre, centre -/+ radius): (int)((w - centre) * (1.0 / resolution)), stomp_collision_space.h:230-234
        const int gmin = int(((b.c[k] - bound) - b.c[k]) * (1.0 / resolution));
        const int gmax = int(((b.c[k] + bound) - b.c[k]) * (1.0 / resolution));
        b.gmin[k] = gmin;
        b.gn[k] = gmax - gmin + 1;
      }
      btotal += (long long)b.gn[0] * b.gn[1] * b.gn[2];
      hb.push_back(b);
    }
    DevBuf<unsigned char> dbodies;
    if (upload(e, dbodies, reinterpret_cast<const unsigned char*>(hb.data()), hb.size() * sizeof(SdfBody))) return 1;
    begin_launch(e);
    k_sdf_mark_bodies<<<unsigned(std::max<long long>(1, std::min<long long>((btotal + 255) / 256, 148 * 32))), 256, 0, e.ws>>>(
        num_bodies, btotal, reinterpret_cast<const SdfBody*>(dbodies.p), origin[0], origin[1], origin[2], resolution, nx, ny, nz, occ.p);
    if (check_launch(e, "k_sdf_mark_bodies")) return 1;
    CUDA_TRY(cudaStreamSynchronize(e.stream));   // dbodies goes out of scope
  }
  if (num_meshes > 0) {
    // bodies::ConvexMesh restated (include/stomp_b200.h): hull, centre, bounding radius, scaled + padded + posed vertices
    std::vector<SdfMesh> hm;
    std::vector<double> tris;
    std::vector<HullTri> hull;
    long long mtotal = 0;
    for (int i = 0; i < num_meshes; ++i) {
      const stomp_mesh_body& mb = meshes[i];
      if (!mb.vertices || mb.num_vertices < 4) return fail("mesh body needs at least 4 vertices");
      if (!(mb.scale > 0.0) || !(mb.padding >= 0.0)) return fail("mesh body scale must be positive and padding non-negative");
      if (!convex_hull_3d(mb.vertices, mb.num_vertices, hull)) return fail("mesh body has no volume (flat or degenerate vertex set)");
      std::vector<char> on_hull(size_t(mb.num_vertices), 0);
      for (const HullTri& t : hull) on_hull[t.a] = on_hull[t.b] = on_hull[t.c] = 1;
      double mc[3] = {0.0, 0.0, 0.0};
      int nh = 0;
      for (int v = 0; v < mb.num_vertices; ++v)      // summed in vertex order (part of the contract: the lattice hangs on these bits)
        if (on_hull[v]) { for (int k = 0; k < 3; ++k) mc[k] += mb.vertices[size_t(v) * 3 + k]; ++nh; }
      for (int k = 0; k < 3; ++k) mc[k] /= double(nh);
      double radius = 0.0;
      for (int v = 0; v < mb.num_vertices; ++v)
        if (on_hull[v]) {
          const double dx = mb.vertices[size_t(v) * 3] - mc[0], dy = mb.vertices[size_t(v) * 3 + 1] - mc[1], dz = mb.vertices[size_t(v) * 3 + 2] - mc[2];
          radius = std::max(radius, std::sqrt(dx * dx + dy * dy + dz * dz));
        }
      double R[9];
      quaternion_to_rotation(mb.orientation, R);
      auto world = [&](const double* b, double* w) {
        for (int r = 0; r < 3; ++r) w[r] = (R[r * 3] * b[0] + R[r * 3 + 1] * b[1] + R[r * 3 + 2] * b[2]) + mb.position[r];
      };
      auto scaled_world = [&](int v, double* w) {
        double d[3] = {mb.vertices[size_t(v) * 3] - mc[0], mb.vertices[size_t(v) * 3 + 1] - mc[1], mb.vertices[size_t(v) * 3 + 2] - mc[2]};
        const double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double fact = mb.scale + (l > 0.0 ? mb.padding / l : 0.0);
        const double sv[3] = {mc[0] + d[0] * fact, mc[1] + d[1] * fact, mc[2] + d[2] * fact};
        world(sv, w);
      };
      SdfMesh m;
      std::memset(&m, 0, sizeof(m));
      world(mc, m.c);
      const double bound = radius * mb.scale + mb.padding;
      if (!(bound > 0.0)) return fail("mesh body with a non-positive size");
      m.first = mtotal;
      for (int k = 0; k < 3; ++k) {
        m.gmin[k] = int(((m.c[k] - bound) - m.c[k]) * (1.0 / resolution));
        const int gmax = int(((m.c[k] + bound) - m.c[k]) * (1.0 / resolution));
        m.gn[k] = gmax - m.gmin[k] + 1;
      }
      mtotal += (long long)m.gn[0] * m.gn[1] * m.gn[2];
      m.tri_first = int(tris.size() / 9);
      m.tri_count = int(hull.size());
      for (const HullTri& t : hull) {
        double w[9];
        scaled_world(t.a, w); scaled_world(t.b, w + 3); scaled_world(t.c, w + 6);
        tris.insert(tris.end(), w, w + 9);
      }
      hm.push_back(m);
    }
    DevBuf<unsigned char> dmeshes;
    DevBuf<double> dtris;
    if (upload(e, dmeshes, reinterpret_cast<const unsigned char*>(hm.data()), hm.size() * sizeof(SdfMesh)) ||
        upload(e, dtris, tris.data(), tris.size()))
      return 1;
    begin_launch(e);
    k_sdf_mark_meshes<<<unsigned(std::max<long long>(1, std::min<long long>((mtotal + 255) / 256, 148 * 32))), 256, 0, e.ws>>>(
        num_meshes, mtotal, reinterpret_cast<const SdfMesh*>(dmeshes.p), dtris.p, origin[0], origin[1], origin[2], resolution, nx, ny, nz,
        occ.p);
    if (check_launch(e, "k_sdf_mark_meshes")) return 1;
    CUDA_TRY(cudaStreamSynchronize(e.stream));   // dmeshes / dtris go out of scope
  }
  const bool u8 = cap * cap < 256;
  CUDA_TRY(e.vox.alloc(cells * (u8 ? 1 : 2)));
  const unsigned egrid = unsigned(std::min<size_t>((cells + 255) / 256, size_t(148) * 64));
  begin_launch(e);
  k_edt_pass<0, uint16_t><<<egrid, 256, 0, e.ws>>>(nx, ny, nz, cap, occ.p, g1.p);
  if (check_launch(e, "k_edt_pass")) return 1;
  begin_launch(e);
  k_edt_pass<1, uint16_t><<<egrid, 256, 0, e.ws>>>(nx, ny, nz, cap, g1.p, g2.p);
  if (check_launch(e, "k_edt_pass")) return 1;
  begin_launch(e);
  if (u8) k_edt_pass<2, uint8_t><<<egrid, 256, 0, e.ws>>>(nx, ny, nz, cap, g2.p, reinterpret_cast<uint8_t*>(e.vox.p));
  else k_edt_pass<2, uint16_t><<<egrid, 256, 0, e.ws>>>(nx, ny, nz, cap, g2.p, reinterpret_cast<uint16_t*>(e.vox.p));
  if (check_launch(e, "k_edt_pass")) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.sdf.vox = e.vox.p;
  e.sdf.nx = nx; e.sdf.ny = ny; e.sdf.nz = nz;
  e.sdf.dtype = u8 ? STOMP_VOXEL_U8_SQ : STOMP_VOXEL_U16_SQ;
  for (int i = 0; i < 3; ++i) e.sdf.origin[i] = origin[i];
  e.sdf.res = resolution;
  e.sdf.inv_res = 1.0 / resolution;
  if (e.f32 ? upload_sqrt_table<float>(e) : upload_sqrt_table<double>(e)) return 1;
  if (build_bricks(e)) return 1;
  e.have_sdf = true;
  e.cull_dirty = true;
  ++e.config_epoch;
  return 0;
}

int stomp_engine_get_sdf(void* h, int32_t dims[3], int32_t* voxel_dtype, void* voxels, size_t bytes) {
  ENGINE_OR_FAIL(h);
  if (!e.have_sdf) return fail("no distance field has been set");
  if (dims) { dims[0] = e.sdf.nx; dims[1] = e.sdf.ny; dims[2] = e.sdf.nz; }
  if (voxel_dtype) *voxel_dtype = e.sdf.dtype;
  if (voxels) {
    if (bytes < e.vox.n) return fail("output buffer too small");
    CUDA_TRY(cudaMemcpyAsync(voxels, e.vox.p, e.vox.n, cudaMemcpyDeviceToHost, e.stream));
    CUDA_TRY(cudaStreamSynchronize(e.stream));
  }
  return 0;
}

int stomp_engine_set_constraints(void* h, const stomp_orientation_constraint* constraints, int32_t n, double constraint_cost_weight) {
  ENGINE_OR_FAIL(h);
  if (!e.have_robot) return fail("set_robot must be called before set_constraints");
  if (n < 0 || n > kMaxConstraints || (n > 0 && !constraints)) return fail("bad constraints");
  for (int i = 0; i < n; ++i)
    if (constraints[i].segment < 0 || constraints[i].segment >= int(e.robot.seg_node.size())) return fail("constraint segment out of range");
  e.constraints.assign(constraints, constraints + n);
  e.constraint_cost_weight = constraint_cost_weight;
  ++e.config_epoch;
  return e.f32 ? upload_constraints<float>(e) : upload_constraints<double>(e);
}

int stomp_engine_set_dynamics(void* h, const stomp_link_inertia* inertia, int32_t chain_root_segment, int32_t chain_tip_segment,
                              const double gravity[3], double torque_cost_weight) {
  ENGINE_OR_FAIL(h);
  if (!e.have_robot) return fail("set_robot must be called before set_dynamics");
  if (!inertia || !gravity) return fail("null argument");
  const int S = int(e.raw_segments.size());
  if (chain_root_segment < 0 || chain_root_segment >= S || chain_tip_segment < 0 || chain_tip_segment >= S)
    return fail("chain segment out of range");
  std::vector<int> path;
  for (int s = chain_tip_segment; s != chain_root_segment; s = e.raw_segments[s].parent) {
    if (s < 0) return fail("chain tip is not below chain root");
    path.push_back(s);
  }
  if (path.empty() || int(path.size()) > kMaxChain) return fail("inverse-dynamics chain is empty or longer than kMaxChain segments");
  std::vector<DevChainLink> links;
  int next_joint = 0;
  for (size_t k = path.size(); k-- > 0;) {
    const stomp_segment& g = e.raw_segments[path[k]];
    const stomp_link_inertia& in = inertia[path[k]];
    DevChainLink L;
    std::memset(&L, 0, sizeof(L));
    L.type = g.joint_type;
    L.group = g.joint_type == STOMP_JOINT_FIXED ? -1 : g.group_index;
    if (g.joint_type != STOMP_JOINT_FIXED) {
      // the reference hands the group's joint arrays to the chain solver: chain joint j must be group joint j
      if (g.group_index != next_joint) return fail("the chain's movable joints must be the group joints in order");
      ++next_joint;
    }
    for (int i = 0; i < 9; ++i) L.rot[i] = g.rot[i];
    for (int i = 0; i < 3; ++i) { L.pos[i] = g.pos[i]; L.axis[i] = g.axis[i]; L.h[i] = in.mass * in.com[i]; }
    L.m = in.mass;
    // KDL::RigidBodyInertia(m, cog, Ic): rotational inertia about the frame origin = Ic + m (|c|^2 1 - c c^T)
    const double* c = in.com;
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double full[9] = {in.inertia[0], in.inertia[3], in.inertia[4], in.inertia[3], in.inertia[1], in.inertia[5],
                            in.inertia[4], in.inertia[5], in.inertia[2]};
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) L.I[a * 3 + b] = full[a * 3 + b] + in.mass * ((a == b ? cc : 0.0) - c[a] * c[b]);
    links.push_back(L);
  }
  if (next_joint != e.D) return fail("the chain's movable joints must be the group joints in order");
  if (join_streams(e)) return 1;
  if (upload(e, e.chain, reinterpret_cast<const unsigned char*>(links.data()), links.size() * sizeof(DevChainLink))) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.chain_len = int(links.size());
  for (int k = 0; k < 3; ++k) e.gravity[k] = gravity[k];
  e.torque_weight = torque_cost_weight;
  ++e.config_epoch;
  return 0;
}

int stomp_engine_set_noise(void* h, const double* noise_stddev, const double* noise_decay) {
  ENGINE_OR_FAIL(h);
  if (!noise_stddev || !noise_decay) return fail("null argument");
  e.noise_stddev.assign(noise_stddev, noise_stddev + e.D);
  e.noise_decay.assign(noise_decay, noise_decay + e.D);
  e.pre_valid = false;   // a look-ahead pass used the old scales
  return 0;
}

int stomp_engine_set_problems(void* h, const double* start, const double* goal) {
  ENGINE_OR_FAIL(h);
  if (!start || !goal) return fail("null argument");
  std::vector<double> th(size_t(e.B) * e.D * e.N);
  for (int b = 0; b < e.B; ++b)
    for (int d = 0; d < e.D; ++d)
      sh::min_control_cost(e.pm, start[size_t(b) * e.D + d], goal[size_t(b) * e.D + d], &th[(size_t(b) * e.D + d) * e.N]);
  if (upload(e, e.theta, th.data(), th.size()) || upload(e, e.pad_start, start, size_t(e.B) * e.D) ||
      upload(e, e.pad_goal, goal, size_t(e.B) * e.D))
    return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  // PolicyImprovement::initialize / setNumRollouts state (src/policy_improvement.cpp:64-147)
  e.reused_next = false;
  e.extra_added = false;
  e.num_gen = 0;
  e.generation = 0;
  e.pre_valid = false;
  e.cand_valid = false;
  e.cur = 0;
  e.injected_pending = false;
  e.have_problems = true;
  e.steady_iterations = 0;
  return 0;
}

int stomp_engine_set_parameters(void* h, const double* theta) {
  ENGINE_OR_FAIL(h);
  if (!theta) return fail("null argument");
  if (upload(e, e.theta, theta, size_t(e.B) * e.D * e.N)) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.cand_valid = false;   // the candidate pass subtracted the old theta
  return 0;
}

int stomp_engine_get_parameters(void* h, double* theta) {
  ENGINE_OR_FAIL(h);
  if (!theta) return fail("null argument");
  CUDA_TRY(cudaMemcpyAsync(theta, e.theta.p, size_t(e.B) * e.D * e.N * 8, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return check_shard_error(e);
}

int stomp_engine_update_parameters(void* h, const double* updates) {
  ENGINE_OR_FAIL(h);
  if (!updates) return fail("null argument");
  size_t n = size_t(e.B) * e.D * e.N;
  if (upload(e, e.updates, updates, n)) return 1;
  begin_launch(e);
  k_axpy<<<unsigned((n + 255) / 256), 256, 0, e.ws>>>(n, e.updates.p, e.theta.p);
  if (check_launch(e, "k_axpy")) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.cand_valid = false;
  return 0;
}

int stomp_engine_compute_control_costs(void* h, const double* parameters, const double* noise, int32_t n, double weight,
                                       double* control_costs) {
  ENGINE_OR_FAIL(h);
  if (!parameters || !noise || !control_costs || n < 1) return fail("bad argument");
  if (!e.have_problems) return fail("set_problems must be called first (the padding values come from the policy)");
  size_t total = size_t(e.B) * n;
  if (ensure_scratch(e, total)) return 1;
  size_t cnt = total * e.D * e.N;
  CUDA_TRY(cudaMemcpyAsync(e.scratch_params.p, parameters, cnt * 8, cudaMemcpyHostToDevice, e.stream));
  CUDA_TRY(cudaMemcpyAsync(e.scratch_noise.p, noise, cnt * 8, cudaMemcpyHostToDevice, e.stream));
  GenArgs a = base_gen_args(e);
  a.R = n; a.R_gen = n;
  a.mode_generate = 0; a.mode_project = 0; a.mode_control = 1;
  a.eps_in = e.scratch_noise.p; a.params = e.scratch_params.p; a.control = e.scratch_costs.p; a.noise_projected = nullptr;
  a.control_weight = weight;
  if (launch_generate(e, a)) return 1;
  CUDA_TRY(cudaMemcpyAsync(control_costs, e.scratch_costs.p, cnt * 8, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_seed(void* h, uint64_t seed) {
  ENGINE_OR_FAIL(h);
  e.seed = seed;
  ++e.config_epoch;
  return 0;
}

int stomp_engine_inject_noise_async(void* h, const double* eps, int32_t n) {
  ENGINE_OR_FAIL(h);
  if (!eps || n < 1 || n > e.R) return fail("bad argument");
  const int buf = e.inject_write;
  e.inject_write ^= 1;
  const size_t count = size_t(e.B) * e.R * e.D * e.N;
  if (e.eps_in2[buf].n != count) CUDA_TRY(e.eps_in2[buf].alloc(count));
  // the buffer was last read by the k_generate two injections ago: do not overwrite it before that launch is done
  CUDA_TRY(cudaStreamWaitEvent(e.copy_stream, e.ev_consumed[buf], 0));
  size_t row = size_t(n) * e.D * e.N * 8;
  CUDA_TRY(cudaMemcpy2DAsync(e.eps_in2[buf].p, size_t(e.R) * e.D * e.N * 8, eps, row, row, e.B, cudaMemcpyHostToDevice, e.copy_stream));
  CUDA_TRY(cudaEventRecord(e.ev_copy_done[buf], e.copy_stream));
  e.inject_pending_buf = buf;
  e.inject_pending_n = n;
  e.injected_pending = true;
  return 0;
}

int stomp_engine_inject_noise(void* h, const double* eps, int32_t n) {
  if (stomp_engine_inject_noise_async(h, eps, n)) return 1;
  Engine& e = *E(h);
  CUDA_TRY(cudaStreamSynchronize(e.copy_stream));
  return 0;
}

int stomp_engine_sample_noise(void* h, int32_t iteration, int32_t n, double* eps_out) {
  ENGINE_OR_FAIL(h);
  if (!eps_out || n < 1) return fail("bad argument");
  size_t total = size_t(e.B) * n;
  if (ensure_scratch(e, total)) return 1;
  std::vector<double> ones(e.D, 1.0);
  if (set_noise_scale(e, ones.data())) return 1;
  GenArgs a = base_gen_args(e);
  a.R = n; a.R_gen = n;
  a.mode_generate = 1; a.mode_project = 0; a.mode_control = 0; a.injected = 0;
  a.iteration = uint32_t(iteration);
  a.noise = e.scratch_noise.p; a.params = e.scratch_params.p; a.noise_projected = nullptr;
  a.rollouts_global = std::max<int64_t>(a.rollouts_global, n);
  if (launch_generate(e, a)) return 1;
  CUDA_TRY(cudaMemcpyAsync(eps_out, e.scratch_noise.p, total * e.D * e.N * 8, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_execute(void* h, const double* parameters, int32_t n, int32_t iteration_number, double* costs,
                         int32_t* collision_free) {
  ENGINE_OR_FAIL(h);
  if (!parameters || !costs || n < 1) return fail("bad argument");
  size_t total = size_t(e.B) * n;
  if (ensure_scratch(e, total)) return 1;
  CUDA_TRY(cudaMemcpyAsync(e.scratch_params.p, parameters, total * e.D * e.N * 8, cudaMemcpyHostToDevice, e.stream));
  if (launch_cost(e, e.scratch_params.p, size_t(n) * e.D * e.N, n, e.B, iteration_number == 1, e.scratch_costs.p, size_t(n) * e.N,
                  e.scratch_flags.p, n, 0, nullptr, nullptr, e.scratch_cflags.p))
    return 1;
  CUDA_TRY(cudaMemcpyAsync(costs, e.scratch_costs.p, total * e.N * 8, cudaMemcpyDeviceToHost, e.stream));
  if (collision_free) CUDA_TRY(cudaMemcpyAsync(collision_free, e.scratch_flags.p, total * 4, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_execute_constraints_satisfied(void* h, int32_t* satisfied, size_t count) {
  ENGINE_OR_FAIL(h);
  if (!satisfied || count > e.scratch_cflags.n) return fail("bad argument");
  CUDA_TRY(cudaMemcpyAsync(satisfied, e.scratch_cflags.p, count * 4, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_execute_debug(void* h, const double* parameters, stomp_sphere_debug* debug) {
  ENGINE_OR_FAIL(h);
  if (!parameters || !debug) return fail("bad argument");
  if (ensure_scratch(e, size_t(e.B))) return 1;
  CUDA_TRY(cudaMemcpyAsync(e.scratch_params.p, parameters, size_t(e.D) * e.N * 8, cudaMemcpyHostToDevice, e.stream));
  CUDA_TRY(cudaMemsetAsync(e.debug.p, 0, e.debug.n * sizeof(stomp_sphere_debug), e.stream));
  if (launch_cost(e, e.scratch_params.p, size_t(e.D) * e.N, 1, 1, 1, e.scratch_costs.p, size_t(e.N), e.scratch_flags.p, 1, 0,
                  e.scratch_noise.p, e.debug.p))
    return 1;
  CUDA_TRY(cudaMemcpyAsync(debug, e.debug.p, size_t(e.N + 3) * e.K * sizeof(stomp_sphere_debug), cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_get_rollouts(void* h, const double* noise_stddev, double* rollouts, int32_t* num_generated) {
  ENGINE_OR_FAIL(h);
  if (!noise_stddev) return fail("null argument");
  if (!e.have_problems) return fail("set_problems must be called first");
  if (set_noise_scale(e, noise_stddev)) return 1;
  if (step_get_rollouts(e, 0, false)) return 1;
  if (gather_reused_state(e)) return 1;
  if (rollouts) {
    size_t row = size_t(e.num_gen) * e.D * e.N * 8;
    CUDA_TRY(cudaMemcpy2DAsync(rollouts, row, e.params[e.cur].p, size_t(e.R) * e.D * e.N * 8, row, e.B, cudaMemcpyDeviceToHost, e.stream));
  }
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  if (num_generated) *num_generated = e.num_gen;
  return 0;
}

int stomp_engine_set_rollout_costs(void* h, const double* costs, double control_cost_weight, double* rollout_costs_total) {
  ENGINE_OR_FAIL(h);
  if (!costs) return fail("null argument");
  if (e.num_gen < 1) return fail("get_rollouts must be called first");
  e.control_cost_weight = control_cost_weight;
  size_t row = size_t(e.num_gen) * e.N * 8;
  CUDA_TRY(cudaMemcpy2DAsync(e.state[e.cur].p, size_t(e.R) * e.N * 8, costs, row, row, e.B, cudaMemcpyHostToDevice, e.stream));
  if (step_control_costs(e)) return 1;
  if (launch_cumulative(e)) return 1;
  if (rollout_costs_total)
    CUDA_TRY(cudaMemcpy2DAsync(rollout_costs_total, size_t(e.R) * 8, e.totals.p, size_t(e.R + 1) * 8, size_t(e.R) * 8, e.B,
                               cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_improve_policy(void* h, double* updates) {
  ENGINE_OR_FAIL(h);
  if (e.desc.rollout_shard_world > 1) return fail("use stomp_engine_iterate_sharded_phase on a rollout-sharded engine");
  if (step_improve(e, 0)) return 1;
  if (updates) CUDA_TRY(cudaMemcpyAsync(updates, e.updates.p, size_t(e.B) * e.D * e.N * 8, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_add_extra_rollouts(void* h, const double* costs) {
  ENGINE_OR_FAIL(h);
  if (!costs) return fail("null argument");
  CUDA_TRY(cudaMemcpyAsync(e.extra_state.p, costs, size_t(e.B) * e.N * 8, cudaMemcpyHostToDevice, e.stream));
  if (step_extra(e, false, 0)) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_iterate(void* h, int32_t iteration_number, stomp_iter_stats* stats) {
  ENGINE_NOJOIN(h);
  if (!e.have_problems) return fail("set_problems must be called first");
  if (iterate_once(e, iteration_number)) return 1;
  return fill_stats(e, stats);
}

int stomp_engine_run(void* h, int32_t first_iteration, int32_t count, stomp_iter_stats* last_stats) {
  ENGINE_NOJOIN(h);
  if (!e.have_problems) return fail("set_problems must be called first");
  int i = 0;
  while (i < count) {
    const int it = first_iteration + i;
    if (count - i >= kGraphIters && graph_eligible(e, it)) {
      const int groups = (count - i) / kGraphIters;
      if (replay_iterations(e, it, groups)) return 1;
      i += groups * kGraphIters;
      continue;
    }
    if (iterate_once(e, it)) return 1;
    ++i;
  }
  return fill_stats(e, last_stats);
}

int stomp_engine_set_graph_mode(void* h, int32_t mode) {
  ENGINE_OR_FAIL(h);
  if (mode != 0 && mode != 1) return fail("graph mode must be 0 or 1");
  e.graph_mode = mode;
  return 0;
}

int stomp_engine_optimize(void* h, int32_t max_iterations, int32_t max_after_cf, stomp_optimize_stats* stats) {
  ENGINE_OR_FAIL(h);
  if (!e.have_problems) return fail("set_problems must be called first");
  if (max_iterations < 1) return fail("max_iterations must be >= 1");
  if (e.desc.rollout_shard_world > 1) return fail("optimize is not available on a rollout-sharded engine");
  // the reference builds a fresh PolicyImprovementLoop for every optimize() (src/stomp_optimizer.cpp:262-263): no rollouts of an
  // earlier run are reused and no extra rollout is pending.  The Philox generation counter is kept, so a second run draws new noise.
  e.reused_next = false;
  e.extra_added = false;
  e.injected_pending = false;
  e.num_gen = 0;
  e.pre_valid = false;
  e.cand_valid = false;
  std::vector<TrackState> init(e.B, TrackState{0, -1, -1, -1, 0, 0});
  if (upload(e, e.track_state, reinterpret_cast<const unsigned char*>(init.data()), init.size() * sizeof(TrackState))) return 1;
  CUDA_TRY(cudaMemsetAsync(e.num_done.p, 0, sizeof(int), e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  const bool want_log = stats && stats->costs;
  if (want_log && e.cost_log.n < size_t(max_iterations) * e.B) CUDA_TRY(e.cost_log.alloc(size_t(max_iterations) * e.B));
  if (want_log) CUDA_TRY(cudaMemsetAsync(e.cost_log.p, 0, e.cost_log.n * sizeof(double), e.stream));   // entries past a problem's last iteration read 0
  const int check_every = 16;   // host looks at the "all problems done" counter this often; results do not depend on it
  int it = 0;
  for (; it < max_iterations; ++it) {
    if (iterate_once(e, it + 1)) return 1;
    if (e.tail_dirty) e.ws = e.tail_stream;   // the bookkeeping follows the noise-less rollout on its stream
    if (e.extra_offchain) e.ws = e.cand_stream;   // ... or k_extra_total, where that runs beside the chain
    begin_launch(e);
    k_track_best<<<e.B, 128, 0, e.ws>>>(it, max_after_cf, e.D * e.N, e.R + 1, e.R, e.noiseless_sum.p, e.collision_free.p,
                                            e.constraints_ok.p, e.extra_clipped.p, reinterpret_cast<TrackState*>(e.track_state.p), e.best_cost.p,
                                            e.best_traj.p, want_log ? e.cost_log.p : nullptr, e.B, e.num_done.p);
    if (check_launch(e, "k_track_best")) return 1;
    if (e.extra_offchain) CUDA_TRY(cudaEventRecord(e.ev_extra, e.cand_stream));
    e.ws = e.stream;
    if ((it + 1) % check_every == 0 && it + 1 < max_iterations) {
      int done = 0;
      if (join_streams(e)) return 1;
      CUDA_TRY(cudaMemcpyAsync(&done, e.num_done.p, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
      CUDA_TRY(cudaStreamSynchronize(e.stream));
      if (done >= e.B) { ++it; break; }
    }
  }
  const int ran = std::min(it, max_iterations);
  if (join_streams(e)) return 1;
  std::vector<TrackState> st(e.B);
  CUDA_TRY(cudaMemcpyAsync(st.data(), e.track_state.p, st.size() * sizeof(TrackState), cudaMemcpyDeviceToHost, e.stream));
  if (stats && stats->best_cost) CUDA_TRY(cudaMemcpyAsync(stats->best_cost, e.best_cost.p, size_t(e.B) * 8, cudaMemcpyDeviceToHost, e.stream));
  if (want_log) CUDA_TRY(cudaMemcpyAsync(stats->costs, e.cost_log.p, size_t(ran) * e.B * 8, cudaMemcpyDeviceToHost, e.stream));
  // group_trajectory_ = best_group_trajectory_ (stomp_optimizer.cpp:368): the policy itself is left at the last iterate
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  if (stats)
    for (int b = 0; b < e.B; ++b) {
      if (stats->success) stats->success[b] = st[b].success_iteration >= 0;
      if (stats->success_iteration) stats->success_iteration[b] = st[b].success_iteration;
      if (stats->collision_success_iteration) stats->collision_success_iteration[b] = st[b].collision_success_iteration;
      if (stats->last_improvement_iteration) stats->last_improvement_iteration[b] = st[b].last_improvement_iteration;
      if (stats->iterations) stats->iterations[b] = st[b].iterations;
    }
  return 0;
}

int stomp_engine_synchronize(void* h) {
  ENGINE_OR_FAIL(h);
  CUDA_TRY(cudaStreamSynchronize(e.copy_stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int stomp_engine_request_results_async(void* h, double* theta, double* noiseless_cost, int32_t* noiseless_collision_free,
                                       int32_t* ticket) {
  ENGINE_NOJOIN(h);
  const int k = e.next_ticket;
  e.next_ticket ^= 1;
  const size_t BDN = size_t(e.B) * e.D * e.N;
  if (e.snap_theta[k].n != BDN) {
    CUDA_TRY(e.snap_theta[k].alloc(BDN));
    CUDA_TRY(e.snap_cost[k].alloc(size_t(e.B)));
    CUDA_TRY(e.snap_flag[k].alloc(size_t(e.B)));
  }
  // the snapshot buffers of this ticket may still be read by the previous request that used it
  if (e.ticket_used[k]) {
    CUDA_TRY(cudaStreamWaitEvent(e.stream, e.ev_results[k], 0));
    CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_results[k], 0));
  }
  // theta: after this iteration's update, before the next one (main stream order)
  if (theta) CUDA_TRY(cudaMemcpyAsync(e.snap_theta[k].p, e.theta.p, BDN * 8, cudaMemcpyDeviceToDevice, e.stream));
  CUDA_TRY(cudaEventRecord(e.ev_snap_main[k], e.stream));
  // noise-less cost / flag: after the noise-less rollout, which runs on the tail stream when iterations overlap
  cudaStream_t ts = e.tail_dirty ? e.tail_stream : e.stream;
  if (e.extra_dirty) CUDA_TRY(cudaStreamWaitEvent(ts, e.ev_extra, 0));   // k_extra_total ran beside the tail stream
  if (noiseless_cost) CUDA_TRY(cudaMemcpyAsync(e.snap_cost[k].p, e.noiseless_sum.p, size_t(e.B) * 8, cudaMemcpyDeviceToDevice, ts));
  if (noiseless_collision_free)
    CUDA_TRY(cudaMemcpy2DAsync(e.snap_flag[k].p, 4, e.collision_free.p + e.R, size_t(e.R + 1) * 4, 4, e.B, cudaMemcpyDeviceToDevice, ts));
  CUDA_TRY(cudaEventRecord(e.ev_snap_tail[k], ts));
  CUDA_TRY(cudaStreamWaitEvent(e.result_stream, e.ev_snap_main[k], 0));
  CUDA_TRY(cudaStreamWaitEvent(e.result_stream, e.ev_snap_tail[k], 0));
  if (theta) CUDA_TRY(cudaMemcpyAsync(theta, e.snap_theta[k].p, BDN * 8, cudaMemcpyDeviceToHost, e.result_stream));
  if (noiseless_cost) CUDA_TRY(cudaMemcpyAsync(noiseless_cost, e.snap_cost[k].p, size_t(e.B) * 8, cudaMemcpyDeviceToHost, e.result_stream));
  if (noiseless_collision_free)
    CUDA_TRY(cudaMemcpyAsync(noiseless_collision_free, e.snap_flag[k].p, size_t(e.B) * 4, cudaMemcpyDeviceToHost, e.result_stream));
  CUDA_TRY(cudaEventRecord(e.ev_results[k], e.result_stream));
  e.ticket_used[k] = true;
  if (ticket) *ticket = k;
  return 0;
}

int stomp_engine_wait_results(void* h, int32_t ticket) {
  ENGINE_NOJOIN(h);
  if (ticket < 0 || ticket > 1 || !e.ticket_used[ticket]) return fail("unknown result ticket");
  CUDA_TRY(cudaEventSynchronize(e.ev_results[ticket]));
  return 0;
}

int stomp_engine_last_stats(void* h, stomp_iter_stats* stats) {
  ENGINE_OR_FAIL(h);
  return fill_stats(e, stats);
}

int stomp_engine_get(void* h, int32_t field, void* out, size_t bytes) {
  ENGINE_OR_FAIL(h);
  if (!out) return fail("null argument");
  const size_t BDN = size_t(e.B) * e.D * e.N, BRDN = BDN * e.R, NN = size_t(e.N) * e.N;
  const void* src = nullptr;
  size_t need = 0;
  std::vector<double> host;
  switch (field) {
    case STOMP_FIELD_THETA: src = e.theta.p; need = BDN * 8; break;
    case STOMP_FIELD_NOISE: src = e.noise.p; need = BRDN * 8; break;
    case STOMP_FIELD_PARAMETERS: src = e.params[e.cur].p; need = BRDN * 8; break;
    case STOMP_FIELD_NOISE_PROJECTED: src = e.noise_projected.p; need = BRDN * 8; break;
    case STOMP_FIELD_STATE_COSTS: src = e.state[e.cur].p; need = size_t(e.B) * e.R * e.N * 8; break;
    case STOMP_FIELD_CONTROL_COSTS: src = e.control[e.cur].p; need = BRDN * 8; break;
    case STOMP_FIELD_CUMULATIVE_COSTS:
      if (e.cumulative_stale && launch_cumulative(e, 0, -1, false, true)) return 1;   // the iteration only needed the totals: fill the array now
      src = e.cumulative.p; need = BRDN * 8; break;
    case STOMP_FIELD_PROBABILITIES: src = e.probabilities.p; need = BRDN * 8; break;
    case STOMP_FIELD_UPDATES: src = e.updates.p; need = BDN * 8; break;
    case STOMP_FIELD_NOISELESS_COSTS: src = e.extra_state.p; need = size_t(e.B) * e.N * 8; break;
    case STOMP_FIELD_COLLISION_FREE: src = e.collision_free.p; need = size_t(e.B) * (e.R + 1) * 4; break;
    case STOMP_FIELD_CONSTRAINTS_SATISFIED: src = e.constraints_ok.p; need = size_t(e.B) * (e.R + 1) * 4; break;
    case STOMP_FIELD_ROLLOUT_TOTAL_COSTS:
      if (e.totals_stale && launch_cumulative(e, 0, -1, true)) return 1;   // nothing ranked them during the iteration: compute now
      src = e.totals.p; need = size_t(e.B) * (e.R + 1) * 8; break;
    case STOMP_FIELD_CLIPPED_PARAMETERS: src = e.clipped.p; need = BRDN * 8; break;
    case STOMP_FIELD_BEST_TRAJECTORY: src = e.best_traj.p; need = BDN * 8; break;
    case STOMP_FIELD_NOISELESS_TRAJECTORY: src = e.extra_clipped.p; need = BDN * 8; break;
    case STOMP_FIELD_INV_CONTROL_COST: host = e.pm.Rinv.a; need = NN * 8; break;
    case STOMP_FIELD_CONTROL_COST: host = e.pm.R.a; need = NN * 8; break;
    case STOMP_FIELD_QUAD_COST_INV: host = e.pm.Qinv.a; need = NN * 8; break;
    case STOMP_FIELD_NOISE_CHOLESKY: {
      sh::Dense L;
      if (!sh::cholesky_lower(e.pm.Rinv, L)) return fail("R^-1 is not positive definite");
      host = L.a; need = NN * 8; break;
    }
    case STOMP_FIELD_PROJECTION: {
      host = e.pm.Rinv.a;
      for (int i = 0; i < e.N; ++i)
        for (int p = 0; p < e.N; ++p) host[size_t(i) * e.N + p] *= e.pm.proj_scale[p];
      need = NN * 8; break;
    }
    default: return fail("unknown field");
  }
  if (bytes < need) return fail("output buffer too small");
  if (!host.empty()) { std::memcpy(out, host.data(), need); return 0; }
  if (!src) return fail("this field is only stored when keep_intermediates is set");
  CUDA_TRY(cudaMemcpyAsync(out, src, need, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return 0;
}

int64_t stomp_engine_launch_count(void* h) { return h ? E(h)->launches : -1; }
void* stomp_engine_stream(void* h) { return h ? static_cast<void*>(E(h)->stream) : nullptr; }

int stomp_engine_timer_start(void* h) {
  ENGINE_OR_FAIL(h);
  CUDA_TRY(cudaEventRecord(e.ev0, e.stream));
  return 0;
}
int stomp_engine_timer_stop(void* h, float* elapsed_ms) {
  ENGINE_OR_FAIL(h);
  CUDA_TRY(cudaEventRecord(e.ev1, e.stream));
  CUDA_TRY(cudaEventSynchronize(e.ev1));
  if (elapsed_ms) CUDA_TRY(cudaEventElapsedTime(elapsed_ms, e.ev0, e.ev1));
  return 0;
}

int stomp_engine_set_profiling(void* h, int32_t enabled) {
  ENGINE_OR_FAIL(h);
  if (join_streams(e)) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.prof_on = enabled != 0;
  e.prof_timeline = enabled == 2;
  e.prof_used = 0;
  return 0;
}

/* Writes the launches recorded since stomp_engine_set_profiling(engine, 2) as CSV: index, kernel, stream (0 main, 1 tail),
 * begin and end in microseconds after the first recorded launch began.  Event timestamps: "begin" is when the stream reached
 * the launch (its dependencies resolved), "end" when the kernel finished. */
int stomp_engine_dump_timeline(void* h, const char* path) {
  ENGINE_OR_FAIL(h);
  if (!path) return fail("null argument");
  if (join_streams(e)) return 1;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  FILE* f = std::fopen(path, "w");
  if (!f) return fail(std::string("cannot open ") + path);
  std::fprintf(f, "index,kernel,stream,begin_us,end_us\n");
  for (size_t i = 0; i < e.prof_used; ++i) {
    float t0 = 0.f, t1 = 0.f;
    if (cudaEventElapsedTime(&t0, e.prof_a[0], e.prof_a[i]) != cudaSuccess ||
        cudaEventElapsedTime(&t1, e.prof_a[0], e.prof_b[i]) != cudaSuccess) {
      std::fclose(f);
      return fail("cudaEventElapsedTime failed while writing the timeline");
    }
    std::fprintf(f, "%zu,%s,%d,%.3f,%.3f\n", i, e.prof_name[i].c_str(), e.prof_stream[i], 1e3 * t0, 1e3 * t1);
  }
  std::fclose(f);
  return 0;
}

/* Sums the event-timed durations of the launches recorded since profiling was enabled whose kernel name
 * contains `kernel_substr` ("" = all).  Resets nothing. */
int stomp_engine_get_profile(void* h, const char* kernel_substr, double* total_ms, int64_t* num_launches) {
  ENGINE_OR_FAIL(h);
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  double ms = 0.0;
  int64_t n = 0;
  for (size_t i = 0; i < e.prof_used; ++i) {
    if (kernel_substr && kernel_substr[0] && e.prof_name[i].find(kernel_substr) == std::string::npos) continue;
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, e.prof_a[i], e.prof_b[i]));
    ms += t;
    ++n;
  }
  if (total_ms) *total_ms = ms;
  if (num_launches) *num_launches = n;
  return 0;
}

int stomp_engine_shard_buffers(void* h, void** minmax_dev, void** sums_dev, size_t* bytes_each) {
  ENGINE_OR_FAIL(h);
  if (minmax_dev) *minmax_dev = e.minmax.p;
  if (sums_dev) *sums_dev = e.sums.p;
  if (bytes_each) *bytes_each = size_t(2) * e.D * e.N * 8;
  return 0;
}

/* ---- peer-memory exchange (rollout sharding without NCCL or host synchronisation between the phases) ----------------- */
static size_t xchg_bytes(const Engine& e) {
  const size_t W = size_t(e.desc.rollout_shard_world), n = size_t(2) * e.D * e.N;
  return 2 * W * n * sizeof(double) + 2 * W * sizeof(unsigned long long);
}

int stomp_engine_shard_ipc_handle(void* h, void* handle_out, size_t handle_bytes) {
  ENGINE_OR_FAIL(h);
  if (!handle_out || handle_bytes < sizeof(cudaIpcMemHandle_t)) return fail("handle buffer must hold 64 bytes");
  if (e.desc.rollout_shard_world > kMaxPeers) return fail("too many ranks for the peer exchange");
  if (e.xchg.n != xchg_bytes(e)) {
    CUDA_TRY(e.xchg.alloc(xchg_bytes(e)));   // zero-filled: flags start below every epoch
  }
  cudaIpcMemHandle_t ipc;
  CUDA_TRY(cudaIpcGetMemHandle(&ipc, e.xchg.p));
  std::memcpy(handle_out, &ipc, sizeof(ipc));
  return 0;
}

int stomp_engine_shard_open_peers(void* h, const void* handles, int32_t count) {
  ENGINE_OR_FAIL(h);
  const int W = e.desc.rollout_shard_world, rank = e.desc.rollout_shard_rank;
  if (!handles || count != W) return fail("one IPC handle per rank is required");
  if (e.xchg.n != xchg_bytes(e)) return fail("call stomp_engine_shard_ipc_handle first");
  if (e.peers_open)   // mapping again (e.g. after a peer was re-created): drop the old mappings first
    for (int r = 0; r < W; ++r)
      if (r != rank && e.peer_base[r]) { cudaIpcCloseMemHandle(e.peer_base[r]); e.peer_base[r] = nullptr; }
  e.peers_open = false;
  for (int r = 0; r < W; ++r) {
    if (r == rank) { e.peer_base[r] = e.xchg.p; continue; }
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, static_cast<const unsigned char*>(handles) + size_t(r) * sizeof(ipc), sizeof(ipc));
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    e.peer_base[r] = static_cast<unsigned char*>(p);
  }
  // A (re-)opened session starts at epoch 0 on every rank: the flag area of the own buffer still holds the epochs of the
  // previous session and must not satisfy the first waits.  Every rank clears its own flags here; the caller must barrier
  // between this call and the first exchange (a peer's first flag store must not be overtaken by the clear).
  const size_t n = size_t(2) * e.D * e.N;
  CUDA_TRY(cudaMemsetAsync(e.xchg.p + 2 * size_t(W) * n * sizeof(double), 0, 2 * size_t(W) * sizeof(unsigned long long), e.stream));
  CUDA_TRY(cudaMemsetAsync(e.xchg_err.p, 0, sizeof(int), e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  e.peers_open = true;
  e.xchg_epoch = 0;
  return 0;
}

/* One iteration of a rollout-sharded engine with both exchanges done in-kernel over peer memory: everything is enqueued on
 * the handle's stream, no host synchronisation, no NCCL.  Every rank must call it for the same iteration. */
int stomp_engine_iterate_sharded_fused(void* h, int32_t iteration_number) {
  ENGINE_NOJOIN(h);
  if (!e.have_problems) return fail("set_problems must be called first");
  if (e.B != 1) return fail("rollout sharding requires num_problems == 1");
  if (e.desc.rollout_shard_world > 1 && !e.peers_open) return fail("call stomp_engine_shard_open_peers first");
  if (e.Rre != 0) {   // with reuse the next iteration's selection needs the noise-less rollout: keep everything in order
    if (join_streams(e)) return 1;
  }
  e.ws = e.stream;
  const bool exchange = e.desc.rollout_shard_world > 1;
  if (iterate_front(e, iteration_number) || launch_shard_stats(e, true, exchange, false, 0)) return 1;
  // the previous iteration's noise-less rollout (tail stream) still reads theta: the launch that ends in the update waits for it
  if (join_streams(e)) return 1;
  if (launch_shard_stats(e, false, exchange, true, 1)) return 1;
  // the noise-less rollout of ONE problem is a ~80 us latency chain nothing in the next iteration depends on (no rollout
  // reuse when rollouts are sharded): it runs on the tail stream under the next iteration's sampling and costs
  CUDA_TRY(cudaEventRecord(e.ev_upd, e.stream));
  e.ws = e.tail_stream;
  CUDA_TRY(cudaStreamWaitEvent(e.tail_stream, e.ev_upd, 0));
  const int rc = step_extra(e, true, iteration_number);
  e.tail_dirty = true;
  e.ws = e.stream;
  return rc;
}

/* 0 when no peer exchange has timed out (blocks until the stream is idle) */
int stomp_engine_shard_status(void* h) {
  ENGINE_OR_FAIL(h);
  return check_shard_error(e);
}

/* phase 0: rollouts, costs, local {max c, max -c}  -> caller all-reduces minmax with MAX
 * phase 1: local {sum e, sum e*eps}                -> caller all-reduces sums with SUM
 * phase 2: update + noise-less rollout (replicated on every rank) */
int stomp_engine_iterate_sharded_phase(void* h, int32_t iteration_number, int32_t phase) {
  ENGINE_OR_FAIL(h);
  if (!e.have_problems) return fail("set_problems must be called first");
  if (e.B != 1) return fail("rollout sharding requires num_problems == 1");
  switch (phase) {
    case 0: if (iterate_front(e, iteration_number)) return 1; return launch_shard_stats(e, true, false, false, 0);
    case 1: return launch_shard_stats(e, false, false, false, 0);
    case 2: if (launch_finalize(e, 1)) return 1; return step_extra(e, true, iteration_number);
    default: return fail("phase must be 0, 1 or 2");
  }
}

}  // extern "C"
