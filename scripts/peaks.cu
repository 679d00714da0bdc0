// peaks.cu — micro-benchmarks of the B200 roofs the STOMP kernels are measured against (SURVEY.md section 8d obliges the
// builder to measure the L2 and fp64 peaks itself; HBM and bf16 come from the driver's MEASURED_PEAKS.json):
//   fp64_fma      dense DFMA issue, 8 independent chains per thread                 -> TFLOP/s
//   fp64_dmma     mma.sync.m8n8k4.f64 (DMMA), 4 independent accumulator tiles per warp -> TFLOP/s (the tensor-pipe fp64 roof)
//   issue         FFMA issue, 8 independent chains per thread                       -> warp instructions / s (all SMs)
//   l2_gather_u8  one random byte per lane from an L2-resident table (the distance-field access pattern of k_cost,
//                 worst case: 32 sectors per warp request)                           -> G gathers/s, GB/s of 32-byte sectors
//   l2_stream     coalesced 16-byte loads from an L2-resident buffer                 -> GB/s
// usage: peaks [out.json]      (CUDA events, best of 5, after a warm-up launch)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
// mma.sync.aligned.m8n8k4.f64 (SASS DMMA): 4 independent accumulator tiles per warp; 512 flop per instruction
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  const double av = a + threadIdx.x * 1e-9, bv = b + threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_gather(const uint8_t* __restrict__ tab, uint32_t mask, int iters, uint32_t* out) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
  for (int i = 0; i < iters; i += 4) {     // four independent gathers in flight per lane
    const uint32_t i0 = s * 1664525u + 1013904223u, i1 = i0 * 1664525u + 1013904223u, i2 = i1 * 1664525u + 1013904223u,
                   i3 = i2 * 1664525u + 1013904223u;
    acc += __ldg(tab + ((i0 >> 7) & mask)) + __ldg(tab + ((i1 >> 7) & mask)) + __ldg(tab + ((i2 >> 7) & mask)) + __ldg(tab + ((i3 >> 7) & mask));
    s = i3;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_stream(const uint4* __restrict__ buf, size_t n16, int passes, uint32_t* out) {
  uint32_t acc = 0;
  for (int p = 0; p < passes; ++p)
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x) {
      const uint4 v = __ldcg(buf + i);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float best_ms(F launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  const int ctas = sms * 8, tpb = 256;
  void* out;
  CK(cudaMalloc(&out, size_t(ctas) * tpb * 8));
  const int iters = 1 << 14;
  const float ms_d = best_ms([&] { k_dfma<<<ctas, tpb>>>(static_cast<double*>(out), iters, 1.0000001, 1e-9); });
  const double fp64_tflops = 2.0 * 8.0 * iters * double(ctas) * tpb / (ms_d * 1e-3) / 1e12;
  const int miters = 1 << 12;
  const float ms_m = best_ms([&] { k_dmma<<<ctas, tpb>>>(static_cast<double*>(out), miters, 1.0000001, 1e-9); });
  const double dmma_tflops = 512.0 * 4.0 * miters * double(ctas) * (tpb / 32) / (ms_m * 1e-3) / 1e12;
  const float ms_f = best_ms([&] { k_ffma<<<ctas, tpb>>>(static_cast<float*>(out), iters, 1.0000001f, 1e-9f); });
  const double warp_inst_per_s = 8.0 * iters * double(ctas) * tpb / 32.0 / (ms_f * 1e-3);
  // 4 MiB table: the size of the C1 / C2 distance field (3.9 MB u8), resident in the 126 MB L2
  const uint32_t tab_bytes = 4u << 20;
  uint8_t* tab;
  CK(cudaMalloc(&tab, tab_bytes));
  CK(cudaMemset(tab, 1, tab_bytes));
  const int giters = 4096;
  const float ms_g = best_ms([&] { k_gather<<<ctas, tpb>>>(tab, tab_bytes - 1, giters, static_cast<uint32_t*>(out)); });
  const double gathers_per_s = double(giters) * ctas * tpb / (ms_g * 1e-3);
  // 32 MiB buffer streamed 8 times: L2 hits after the first pass
  const size_t sbytes = size_t(32) << 20;
  uint4* sbuf;
  CK(cudaMalloc(&sbuf, sbytes));
  CK(cudaMemset(sbuf, 0, sbytes));
  const int passes = 8;
  const float ms_s = best_ms([&] { k_stream<<<ctas, tpb>>>(sbuf, sbytes / 16, passes, static_cast<uint32_t*>(out)); });
  const double l2_stream_gbs = double(sbytes) * passes / (ms_s * 1e-3) / 1e9;
  char buf[2048];
  std::snprintf(buf, sizeof(buf),
                "{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_nominal\": %.0f, "
                "\"fp64_fma_tflops\": %.3f, \"fp64_dmma_tflops\": %.3f, \"issue_warp_inst_per_s\": %.4e, \"issue_warp_inst_per_clk_per_sm\": %.3f, "
                "\"l2_gather_u8_per_s\": %.4e, \"l2_gather_sector_gbs\": %.1f, \"l2_stream_gbs\": %.1f, "
                "\"how\": \"scripts/peaks.cu: CUDA events, best of 5 after a warm-up launch; fp64: 8 DFMA chains/thread x %d CTAs x 256 thr; "
                "issue: FFMA, same shape; l2_gather: 4 independent random byte loads in flight per lane from a 4 MiB table (one 32 B sector "
                "per gather); l2_stream: 16 B coalesced loads, 32 MiB x 8 passes\"}",
                prop.name, sms, clock_khz / 1e3, fp64_tflops, dmma_tflops, warp_inst_per_s, warp_inst_per_s / (double(sms) * clock_khz * 1e3),
                gathers_per_s, gathers_per_s * 32.0 / 1e9, l2_stream_gbs, ctas);
  std::printf("%s\n", buf);
  if (argc > 1) {
    FILE* f = std::fopen(argv[1], "w");
    if (f) { std::fprintf(f, "%s\n", buf); std::fclose(f); }
  }
  return 0;
}
