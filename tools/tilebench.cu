/*
 * tools/tilebench.cu -- register-tile shapes of the tiled FIR inner loop, data resident in
 * shared memory (no TMA, no barriers, no epilogue): which thread tile gets closest to the
 * FP32 pipe when the only competitors of the FMAs are the shared-memory loads themselves.
 *
 *   LO  output lanes per warp (the other 32/LO lane groups take different client sets)
 *   RK  outputs per thread          RC  clients per thread          PK  packed FFMA2
 *   warp tile = (LO*RK) outputs x (32/LO*RC) clients; W warps per CTA split the clients.
 *
 * Per tap a thread loads RK x-values (LDS.64) and RC taps (RC/2 LDS.128) for 4*RK*RC FMAs.
 * Prints FMA/clk/SM from CUDA-event time and the SM clock.  Usage: tilebench [passes [prod]]
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 pack2f(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

constexpr int DP = 42;   // decimation of the 48 ksps class (2-way conflicts on the x loads, as in production)
constexpr int LF = 32;   // flat taps resident per pass

template <int LO, int RK, int RC, int W, bool PK, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) k_tile(float *out, int passes) {
  constexpr int NCL = (32 / LO) * RC * W;  // clients per CTA
  constexpr int KT = LO * RK;              // outputs per CTA
  extern __shared__ __align__(128) unsigned char smem[];
  float2 *ts = reinterpret_cast<float2 *>(smem);                    // [LF][NCL]
  float2 *xs = reinterpret_cast<float2 *>(smem + LF * NCL * 8);     // (KT-1)*DP + LF
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o = lane & (LO - 1), h = lane / LO;
  const int cbase = (warp * (32 / LO) + h) * RC;
  const int xs_len = (KT - 1) * DP + LF + 8;
  for (int i = tid; i < LF * NCL; i += W * 32) ts[i] = make_float2(0.001f * (i % 13), 0.002f * (i % 7));
  for (int i = tid; i < xs_len; i += W * 32) xs[i] = make_float2(0.01f * (i % 11), 0.02f * (i % 5));
  __syncthreads();
  float2 acc[RK][RC];
  u64 RE[RK][RC / 2], IM[RK][RC / 2];
#pragma unroll
  for (int i = 0; i < RK; i++) {
#pragma unroll
    for (int c = 0; c < RC; c++) acc[i][c] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < RC / 2; q++) RE[i][q] = IM[i][q] = 0ull;
  }
  const float2 *xb[RK];
#pragma unroll
  for (int i = 0; i < RK; i++) xb[i] = xs + (o + LO * i) * DP;
  const float4 *tp = reinterpret_cast<const float4 *>(ts + cbase);
  constexpr int UN = (RK * RC >= 64) ? 4 : 8;
  for (int p = 0; p < passes; p++) {
#pragma unroll 1
    for (int f = 0; f < LF; f += UN) {
#pragma unroll
      for (int u = 0; u < UN; u++) {
        float2 x[RK];
#pragma unroll
        for (int i = 0; i < RK; i++) x[i] = xb[i][f + u];
        if constexpr (PK) {
          ulonglong2 tq[RC / 2];
#pragma unroll
          for (int q = 0; q < RC / 2; q++) tq[q] = reinterpret_cast<const ulonglong2 *>(tp)[(f + u) * (NCL / 2) + q];
#pragma unroll
          for (int i = 0; i < RK; i++) {
            const u64 XR = pack2f(x[i].x, x[i].x), XI = pack2f(x[i].y, x[i].y);
            const u64 XN = XI ^ 0x8000000080000000ull;
#pragma unroll
            for (int q = 0; q < RC / 2; q++) {
              RE[i][q] = ffma2(XR, tq[q].x, RE[i][q]);
              RE[i][q] = ffma2(XN, tq[q].y, RE[i][q]);
              IM[i][q] = ffma2(XR, tq[q].y, IM[i][q]);
              IM[i][q] = ffma2(XI, tq[q].x, IM[i][q]);
            }
          }
        } else {
          float4 tq[RC / 2];
#pragma unroll
          for (int q = 0; q < RC / 2; q++) tq[q] = tp[(f + u) * (NCL / 2) + q];
#pragma unroll
          for (int i = 0; i < RK; i++)
#pragma unroll
            for (int q = 0; q < RC / 2; q++) {
              float2 &a0 = acc[i][2 * q], &a1 = acc[i][2 * q + 1];
              a0.x = fmaf(x[i].x, tq[q].x, a0.x);
              a0.x = fmaf(-x[i].y, tq[q].y, a0.x);
              a0.y = fmaf(x[i].x, tq[q].y, a0.y);
              a0.y = fmaf(x[i].y, tq[q].x, a0.y);
              a1.x = fmaf(x[i].x, tq[q].z, a1.x);
              a1.x = fmaf(-x[i].y, tq[q].w, a1.x);
              a1.y = fmaf(x[i].x, tq[q].w, a1.y);
              a1.y = fmaf(x[i].y, tq[q].z, a1.y);
            }
        }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < RK; i++) {
#pragma unroll
    for (int c = 0; c < RC; c++) s += acc[i][c].x + acc[i][c].y;
#pragma unroll
    for (int q = 0; q < RC / 2; q++) {
      float a, b, c2, d;
      unpack2f(RE[i][q], a, b);
      unpack2f(IM[i][q], c2, d);
      s += a + b + c2 + d;
    }
  }
  out[blockIdx.x * (W * 32) + tid] = s;
}

template <int LO, int RK, int RC, int W, bool PK, int MINB>
static void run(const char *name, int per_sm, int passes, float *d_out, int sms, double mhz) {
  constexpr int NCL = (32 / LO) * RC * W, KT = LO * RK;
  const size_t smem = (size_t)LF * NCL * 8 + ((size_t)(KT - 1) * DP + LF + 8) * 8;
  auto kernel = k_tile<LO, RK, RC, W, PK, MINB>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, W * 32, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kernel);
  if (occ < per_sm) {
    printf("{\"bench\": \"%s\", \"ctas_per_sm\": %d, \"skipped\": \"occupancy %d (regs %d, smem %zu)\"}\n", name, per_sm, occ,
           fa.numRegs, smem);
    return;
  }
  const int blocks = per_sm * sms;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kernel<<<blocks, W * 32, smem>>>(d_out, 2);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("{\"bench\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(err));
    return;
  }
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    kernel<<<blocks, W * 32, smem>>>(d_out, passes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fma = 4.0 * RK * RC * LF * (double)passes * (W * 32.0) * blocks;
  const double tf = fma / (best * 1e-3) / 1e12;
  printf("{\"bench\": \"%s\", \"LO\": %d, \"RK\": %d, \"RC\": %d, \"warps\": %d, \"packed\": %d, \"ctas_per_sm\": %d, \"warps_per_sm\": %d, "
         "\"regs\": %d, \"smem\": %zu, \"tfma_per_s\": %.3f, \"fma_per_clk_per_sm_at_max_clock\": %.1f, \"ms\": %.4f}\n",
         name, LO, RK, RC, W, PK ? 1 : 0, per_sm, per_sm * W, fa.numRegs, smem, tf, tf * 1e12 / (sms * mhz * 1e6), best);
  fflush(stdout);
}

int main(int argc, char **argv) {
  int passes = argc > 1 ? atoi(argv[1]) : 400;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1e3;
  float *d_out;
  cudaMalloc(&d_out, sizeof(float) * 148 * 8 * 256);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", prop.name, sms, mhz);
  if (argc > 2) {  // `tilebench <passes> prod`: the production tile at production occupancy only (bench.py's loop ceiling)
    run<16, 4, 8, 2, false, 4>("cur_16x4x8_scalar", 4, passes, d_out, sms, mhz);
    return 0;
  }
  for (int per_sm = 5; per_sm <= 8; per_sm++) {  // more resident warps than production's 8 per SM
    run<16, 4, 8, 2, false, 8>("occ_16x4x8_scalar", per_sm, passes, d_out, sms, mhz);
    run<16, 4, 8, 2, true, 8>("occ_16x4x8_packed", per_sm, passes, d_out, sms, mhz);
  }
  // other thread tiles at production occupancy (4 CTAs x 2 warps)
  run<16, 3, 8, 2, false, 4>("t_16x3x8", 4, passes, d_out, sms, mhz);
  run<16, 5, 8, 2, false, 4>("t_16x5x8", 4, passes, d_out, sms, mhz);
  run<16, 6, 8, 2, false, 4>("t_16x6x8", 4, passes, d_out, sms, mhz);
  run<8, 6, 8, 2, false, 4>("t_8x6x8", 4, passes, d_out, sms, mhz);
  run<8, 7, 8, 2, false, 4>("t_8x7x8", 4, passes, d_out, sms, mhz);
  run<4, 7, 8, 2, false, 4>("t_4x7x8", 4, passes, d_out, sms, mhz);
  run<4, 7, 8, 8, false, 1>("t_4x7x8_8warps_long4", 1, passes, d_out, sms, mhz);
  run<16, 4, 4, 2, false, 4>("t_16x4x4", 4, passes, d_out, sms, mhz);
  run<16, 4, 12, 2, false, 4>("t_16x4x12", 4, passes, d_out, sms, mhz);
  run<16, 6, 4, 2, false, 4>("t_16x6x4", 4, passes, d_out, sms, mhz);
  run<16, 8, 4, 2, false, 4>("t_16x8x4", 4, passes, d_out, sms, mhz);
  for (int per_sm = 2; per_sm <= 4; per_sm += 2) {
    run<16, 4, 8, 2, false, 4>("cur_16x4x8_scalar", per_sm, passes, d_out, sms, mhz);
    run<16, 4, 8, 2, true, 4>("cur_16x4x8_packed", per_sm, passes, d_out, sms, mhz);
    run<8, 8, 8, 2, false, 4>("lo8_8x8_scalar", per_sm, passes, d_out, sms, mhz);
    run<8, 8, 8, 2, true, 4>("lo8_8x8_packed", per_sm, passes, d_out, sms, mhz);
    run<16, 4, 16, 2, false, 4>("lo16_4x16_scalar", per_sm, passes, d_out, sms, mhz);
    run<16, 4, 16, 2, true, 4>("lo16_4x16_packed", per_sm, passes, d_out, sms, mhz);
    run<16, 8, 8, 2, false, 4>("lo16_8x8_scalar", per_sm, passes, d_out, sms, mhz);
    run<16, 8, 8, 2, true, 4>("lo16_8x8_packed", per_sm, passes, d_out, sms, mhz);
    run<32, 4, 8, 4, false, 2>("lo32_4x8_scalar_4w", per_sm / 2, passes, d_out, sms, mhz);
    run<16, 2, 8, 2, false, 4>("lo16_2x8_scalar", per_sm, passes, d_out, sms, mhz);
  }
  return 0;
}
