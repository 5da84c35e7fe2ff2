// Micro-benchmarks of the sm_100a resources the SGFE operator leans on: DFMA issue rate, fp64 tensor (DMMA)
// rate, shared-memory load throughput (broadcast / per-lane, 64 / 128 bit), shuffles, and mixes of them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu ; run on a B200.
// Output: one line per test, rates per SM and clock (cycles from clock64 inside the kernel, one CTA per SM).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- A: DFMA, 8 independent chains per thread
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
  double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < ITERS; i++) {
    c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
    c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- B: shared-memory loads only. MODE 0: LDS.64 broadcast, 1: LDS.128 broadcast, 2: LDS.64 per-lane stride 1,
//         3: LDS.128 per-lane stride 1, 4: LDS.64 with 4 distinct addresses per warp (lane/8), 5: LDS.32 stride 1
template <int MODE>
__global__ void k_lds(double* out, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned base = smem_u32(sm) + warp * 64;
  if (MODE == 2) base += lane * 8;
  if (MODE == 3) base += lane * 16;
  if (MODE == 4) base += (lane >> 3) * 8 * 5;
  if (MODE == 5) base += lane * 4;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (MODE == 0 || MODE == 2 || MODE == 4) {
        double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + u * 512));
        acc += v;
      } else if (MODE == 5) {
        float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + u * 512));
        acc += v;
      } else {
        double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(base + u * 1024));
        acc += v; acc += w;
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- D: DMMA m8n8k4, NA independent accumulators per warp
template <int NA>
__global__ void k_dmma(double* out, long long* cyc, double a, double b) {
  double c[NA][2];
#pragma unroll
  for (int s = 0; s < NA; s++) { c[s][0] = threadIdx.x + s; c[s][1] = s; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int s = 0; s < NA; s++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[s][0]), "+d"(c[s][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int s = 0; s < NA; s++) r += c[s][0] + c[s][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- D2: DMMA m16n8k8 (sm_90+ shape): A 4 regs, B 2 regs, C 4 regs
template <int NA>
__global__ void k_dmma16(double* out, long long* cyc, double a, double b) {
  double c[NA][4];
#pragma unroll
  for (int s = 0; s < NA; s++) { c[s][0] = threadIdx.x + s; c[s][1] = s; c[s][2] = 1; c[s][3] = 2; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int s = 0; s < NA; s++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[s][0]), "+d"(c[s][1]), "+d"(c[s][2]), "+d"(c[s][3])
                   : "d"(a), "d"(b), "d"(a + 1), "d"(b + 1), "d"(a), "d"(b));
  }
  long long t1 = clock64();
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int s = 0; s < NA; s++) r += c[s][0] + c[s][1] + c[s][2] + c[s][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- E: DMMA and DFMA interleaved in the same warp (4 DMMA accumulators + 8 DFMA chains)
__global__ void k_mix(double* out, long long* cyc, double a, double b) {
  double c[4][2];
#pragma unroll
  for (int s = 0; s < 4; s++) { c[s][0] = threadIdx.x + s; c[s][1] = s; }
  double f0 = 1, f1 = 2, f2 = 3, f3 = 4, f4 = 5, f5 = 6, f6 = 7, f7 = 8;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int s = 0; s < 4; s++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[s][0]), "+d"(c[s][1]) : "d"(a), "d"(b));
    f0 = fma(f0, a, b); f1 = fma(f1, a, b); f2 = fma(f2, a, b); f3 = fma(f3, a, b);
    f4 = fma(f4, a, b); f5 = fma(f5, a, b); f6 = fma(f6, a, b); f7 = fma(f7, a, b);
  }
  long long t1 = clock64();
  __syncthreads();
  double r = f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
#pragma unroll
  for (int s = 0; s < 4; s++) r += c[s][0] + c[s][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- F: 64-bit shuffles (2 SHFL.32 each)
__global__ void k_shfl(double* out, long long* cyc) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
    v0 = __shfl_xor_sync(0xffffffffu, v0, 1); v1 = __shfl_xor_sync(0xffffffffu, v1, 2);
    v2 = __shfl_xor_sync(0xffffffffu, v2, 4); v3 = __shfl_xor_sync(0xffffffffu, v3, 8);
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- F2: 4 SHFL.32 + 2 LDS.64 (stride 1) per iteration: do shuffles take shared-memory pipe cycles?
__global__ void k_shfl_lds(double* out, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  unsigned base = smem_u32(sm) + (threadIdx.x & 31) * 8;
  int v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
    v0 = __shfl_xor_sync(0xffffffffu, v0, 1); v1 = __shfl_xor_sync(0xffffffffu, v1, 2);
    v2 = __shfl_xor_sync(0xffffffffu, v2, 4); v3 = __shfl_xor_sync(0xffffffffu, v3, 8);
    double a, b;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(base));
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b) : "r"(base + 512));
    acc += a + b;
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v0 + v1 + v2 + v3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- G: STS.64 stride 1 + LDS.64 stride 1 (exchange pattern)
__global__ void k_xchg(double* out, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  unsigned base = smem_u32(sm) + threadIdx.x * 8;
  double acc = 0, v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      asm volatile("st.shared.f64 [%0], %1;" :: "r"(base + u * 8192), "d"(v));
      double w; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(w) : "r"(base + ((u + 2) & 3) * 8192));
      acc += w;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double med_cycles(long long* d_cyc, int nb) {
  std::vector<long long> h(nb);
  CK(cudaMemcpy(h.data(), d_cyc, nb * sizeof(long long), cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  return (double)h[nb / 2];
}

int main() {
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  printf("SMs %d, clock %d kHz\n", nsm, khz);
  double* out; long long* cyc; double* xin;
  CK(cudaMalloc(&out, sizeof(double) * nsm * 1024));
  CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
  CK(cudaMalloc(&xin, sizeof(double) * 1024 * 4 * 7));
  CK(cudaMemset(xin, 0, sizeof(double) * 1024 * 4 * 7));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int smem = 32768;
#define RUN(name, per_thread_ops, unit, launch) \
  for (int rep = 0; rep < 2; rep++) { CK(cudaEventRecord(e0)); launch; CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError()); \
    if (rep) { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); double c = med_cycles(cyc, nsm); \
      printf("%-44s warps %2d: %8.2f %s/clk/SM  (%.0f cycles, %.3f ms)\n", name, nw, (double)(per_thread_ops) * nw * 32 / c, unit, c, ms); } }

  for (int nw : {4, 8, 16, 32}) {
    RUN("DFMA 8 chains", 8.0 * ITERS, "FMA", (k_dfma<<<nsm, nw * 32>>>(out, cyc, 1.0000001, 1e-9)));
  }
  for (int nw : {4, 8, 16, 32}) {
    RUN("DMMA m8n8k4 x4 acc", 4.0 * ITERS * 256 / 32, "FMA", (k_dmma<4><<<nsm, nw * 32>>>(out, cyc, 1.0000001, 1e-9)));
  }
  for (int nw : {4, 8, 16}) {
    RUN("DMMA m8n8k4 x8 acc", 8.0 * ITERS * 256 / 32, "FMA", (k_dmma<8><<<nsm, nw * 32>>>(out, cyc, 1.0000001, 1e-9)));
  }
  for (int nw : {4, 8, 16}) {
    RUN("DMMA m16n8k8 x4 acc", 4.0 * ITERS * 1024 / 32, "FMA", (k_dmma16<4><<<nsm, nw * 32>>>(out, cyc, 1.0000001, 1e-9)));
  }
  for (int nw : {4, 8, 16}) {
    RUN("mix 4 DMMA + 8 DFMA per iter", (4.0 * 8 + 8.0) * ITERS, "FMA", (k_mix<<<nsm, nw * 32>>>(out, cyc, 1.0000001, 1e-9)));
  }
  for (int nw : {8, 16, 32}) {
    RUN("LDS.64 broadcast", 8.0 * ITERS / 32, "instr", (k_lds<0><<<nsm, nw * 32, smem>>>(out, cyc)));
    RUN("LDS.128 broadcast", 8.0 * ITERS / 32, "instr", (k_lds<1><<<nsm, nw * 32, smem>>>(out, cyc)));
    RUN("LDS.64 stride 1", 8.0 * ITERS / 32, "instr", (k_lds<2><<<nsm, nw * 32, smem>>>(out, cyc)));
    RUN("LDS.128 stride 1", 8.0 * ITERS / 32, "instr", (k_lds<3><<<nsm, nw * 32, smem>>>(out, cyc)));
    RUN("LDS.64 4 addresses per warp", 8.0 * ITERS / 32, "instr", (k_lds<4><<<nsm, nw * 32, smem>>>(out, cyc)));
    RUN("LDS.32 stride 1", 8.0 * ITERS / 32, "instr", (k_lds<5><<<nsm, nw * 32, smem>>>(out, cyc)));
  }
  CK(cudaFuncSetAttribute(k_xchg, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  for (int nw : {8, 16, 32}) {
    RUN("STS.64 + LDS.64 stride 1 (pairs)", 4.0 * ITERS / 32, "pair", (k_xchg<<<nsm, nw * 32, 131072>>>(out, cyc)));
    RUN("SHFL 64-bit", 4.0 * ITERS / 32, "shfl64", (k_shfl<<<nsm, nw * 32>>>(out, cyc)));
    RUN("4 SHFL.32 + 2 LDS.64 stride 1 (iterations)", 1.0 * ITERS / 32, "iter", (k_shfl_lds<<<nsm, nw * 32, smem>>>(out, cyc)));
  }
  return 0;
}
