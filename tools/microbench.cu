/*
 * tools/microbench.cu -- FP32 pipe micro-benchmarks for the roofline denominators.
 *
 * MEASURED_PEAKS.json holds an HBM copy figure and a bf16 tensor figure, but the
 * xlating FIR is bound by the FP32 FMA pipe, so the peak it
 * is compared with must be measured too:
 *   ffma     the register pattern of the tiled kernel's inner loop (4 outputs x
 *            8 clients, complex MAC = 4 dependent-free FFMA per pair), no loads
 *   ffma2    the same math issued as packed fma.rn.f32x2 (sm_100 FFMA2)
 * Each is run at several occupancies; prints one JSON object per line with the
 * achieved FMA/clk/SM (from clock64) and TFMA/s (from CUDA events).
 * Usage: microbench [iters]
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long ull;

__device__ __forceinline__ ull pack2(float lo, float hi) {
  ull r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(ull v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ ull ffma2(ull a, ull b, ull c) {
  ull d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__global__ void k_ffma(float *out, const float *in, int iters, long long *cycles) {
  float2 x[4], t[8], acc[4][8];
  for (int i = 0; i < 4; i++) x[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 4 + i]);
  for (int c = 0; c < 8; c++) t[c] = make_float2(in[8 + c], in[16 + c]);
  for (int i = 0; i < 4; i++)
    for (int c = 0; c < 8; c++) acc[i][c] = make_float2(0.f, 0.f);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int c = 0; c < 8; c++) {
        acc[i][c].x = fmaf(x[i].x, t[c].x, acc[i][c].x);
        acc[i][c].x = fmaf(-x[i].y, t[c].y, acc[i][c].x);
        acc[i][c].y = fmaf(x[i].x, t[c].y, acc[i][c].y);
        acc[i][c].y = fmaf(x[i].y, t[c].x, acc[i][c].y);
      }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 4; i++)
    for (int c = 0; c < 8; c++) s += acc[i][c].x + acc[i][c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma2(float *out, const float *in, int iters, long long *cycles) {
  ull xa[4], xb[4], ta[8], tb[8], acc[4][8];
  for (int i = 0; i < 4; i++) {
    float xr = in[threadIdx.x + i], xi = in[threadIdx.x + 4 + i];
    xa[i] = pack2(xr, xr);
    xb[i] = pack2(-xi, xi);
  }
  for (int c = 0; c < 8; c++) {
    float tr = in[8 + c], ti = in[16 + c];
    ta[c] = pack2(tr, ti);
    tb[c] = pack2(ti, tr);
  }
  for (int i = 0; i < 4; i++)
    for (int c = 0; c < 8; c++) acc[i][c] = pack2(0.f, 0.f);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int c = 0; c < 8; c++) {
        acc[i][c] = ffma2(xa[i], ta[c], acc[i][c]);
        acc[i][c] = ffma2(xb[i], tb[c], acc[i][c]);
      }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 4; i++)
    for (int c = 0; c < 8; c++) {
      float lo, hi;
      unpack2(acc[i][c], lo, hi);
      s += lo + hi;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// dependent-chain latency of the oscillator recursion p <- p*inc (unfused complex
// multiply): MODE 0 = FMUL/FADD as written, MODE 1 = the same roundings expressed
// with FFMA only (fma(a,b,-0) == round(a*b); fma(x,1,-y) == round(x-y))
template <int MODE>
__global__ void k_chain(float *out, const float *in, int steps, long long *cycles) {
  float pr = in[threadIdx.x] + 1.0f, pi = in[threadIdx.x + 32];
  const float ir = 0.99f + in[1], ii = 0.1f + in[2];
  long long t0 = clock64();
#pragma unroll 8
  for (int k = 0; k < steps; k++) {
    float nr, ni;
    if (MODE == 0) {
      nr = __fsub_rn(__fmul_rn(pr, ir), __fmul_rn(pi, ii));
      ni = __fadd_rn(__fmul_rn(pr, ii), __fmul_rn(pi, ir));
    } else {
      const float a = __fmaf_rn(pr, ir, -0.0f), b = __fmaf_rn(pi, ii, -0.0f);
      const float c = __fmaf_rn(pr, ii, -0.0f), d = __fmaf_rn(pi, ir, -0.0f);
      nr = __fmaf_rn(a, 1.0f, -b);
      ni = __fmaf_rn(c, 1.0f, d);
    }
    pr = nr;
    pi = ni;
  }
  long long t1 = clock64();
  out[threadIdx.x] = pr + pi;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <typename K>
static void run(const char *name, K kernel, int blocks_per_sm, int threads, int iters, int sms, float *d_out,
                float *d_in, long long *d_cyc) {
  const int blocks = blocks_per_sm * sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int w = 0; w < 2; w++) kernel<<<blocks, threads>>>(d_out, d_in, iters, d_cyc);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    kernel<<<blocks, threads>>>(d_out, d_in, iters, d_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long *h = (long long *)malloc(sizeof(long long) * blocks);
  cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double mean_cyc = 0;
  for (int i = 0; i < blocks; i++) mean_cyc += (double)h[i];
  mean_cyc /= blocks;
  free(h);
  const double fma_per_thread = (double)iters * 128.0;  // 32 complex MACs x 4
  const double total = fma_per_thread * threads * (double)blocks;
  // per SM and clock: all resident blocks of an SM run concurrently for ~mean_cyc
  const double per_clk_sm = fma_per_thread * threads * blocks_per_sm / mean_cyc;
  printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"fma_per_clk_per_sm\": %.2f, \"tfma_per_s\": %.3f, "
         "\"ms\": %.4f, \"implied_sm_mhz\": %.0f}\n",
         name, blocks_per_sm * threads / 32, per_clk_sm, total / (best * 1e-3) / 1e12, best,
         mean_cyc / (best * 1e-3) / 1e6);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

int main(int argc, char **argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
    fprintf(stderr, "no CUDA device\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount;
  float *d_out, *d_in;
  long long *d_cyc;
  cudaMalloc(&d_out, sizeof(float) * 148 * 16 * 1024);
  cudaMalloc(&d_in, sizeof(float) * 2048);
  cudaMalloc(&d_cyc, sizeof(long long) * 148 * 64);
  float h_in[2048];
  for (int i = 0; i < 2048; i++) h_in[i] = 0.001f * (float)(i % 97) - 0.04f;
  cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
  const int cfg[][2] = {{1, 128}, {2, 128}, {4, 128}, {3, 128}, {2, 256}};
  for (auto &c : cfg) {
    run("ffma", k_ffma, c[0], c[1], iters, sms, d_out, d_in, d_cyc);
    run("ffma2", k_ffma2, c[0], c[1], iters, sms, d_out, d_in, d_cyc);
  }
  for (int mode = 0; mode < 2; mode++) {
    const int steps = 100000;
    if (mode == 0)
      k_chain<0><<<1, 32>>>(d_out, d_in, steps, d_cyc);
    else
      k_chain<1><<<1, 32>>>(d_out, d_in, steps, d_cyc);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("{\"bench\": \"osc_chain_%s\", \"cycles_per_step\": %.2f}\n", mode == 0 ? "fmul_fadd" : "ffma_only",
           (double)cyc / steps);
  }
  return 0;
}
