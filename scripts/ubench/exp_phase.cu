// Microbenchmark (not part of the library): cycles per "exponential phase" of the attention softmax -- 128 scores per
// thread: scale-and-shift, ex2, row sum, fp16 pack -- as a function of warps per scheduler and of the instruction mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_phase exp_phase.cu ; run: ./exp_phase
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(384) k(const float *in, uint32_t *out, long long *cyc, float scale, float negm, int iters) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 4095];
  uint32_t acc = 0;
  float sum0 = 0.f, sum1 = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[64];
#pragma unroll
    for (int i = 0; i < 128; i += 2) {
      float a, b;
      if (MODE == 0) {            // full mix
        a = ex2a(fmaf(s[i], scale, negm)); b = ex2a(fmaf(s[i + 1], scale, negm));
        sum0 += a; sum1 += b;
        const __half2 h = __floats2half2_rn(a, b);
        pk[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
      } else if (MODE == 1) {     // MUFU only
        a = ex2a(s[i]); b = ex2a(s[i + 1]);
        pk[i >> 1] = __float_as_uint(a) ^ __float_as_uint(b);
      } else if (MODE == 2) {     // no MUFU: everything else
        a = fmaf(s[i], scale, negm); b = fmaf(s[i + 1], scale, negm);
        sum0 += a; sum1 += b;
        const __half2 h = __floats2half2_rn(a, b);
        pk[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
      } else if (MODE == 3) {     // FFMA + MUFU only
        a = ex2a(fmaf(s[i], scale, negm)); b = ex2a(fmaf(s[i + 1], scale, negm));
        pk[i >> 1] = __float_as_uint(a) ^ __float_as_uint(b);
      } else if (MODE == 4) {     // FFMA + MUFU + pack (no sums)
        a = ex2a(fmaf(s[i], scale, negm)); b = ex2a(fmaf(s[i + 1], scale, negm));
        const __half2 h = __floats2half2_rn(a, b);
        pk[i >> 1] = *reinterpret_cast<const uint32_t *>(&h);
      } else if (MODE == 5) {     // FFMA + MUFU + sums (no pack)
        a = ex2a(fmaf(s[i], scale, negm)); b = ex2a(fmaf(s[i + 1], scale, negm));
        sum0 += a; sum1 += b;
        pk[i >> 1] = __float_as_uint(a) ^ __float_as_uint(b);
      } else if (MODE == 6) {     // MUFU on independent inputs, no FFMA (inputs perturbed by an integer op)
        a = ex2a(__uint_as_float(__float_as_uint(s[i]) ^ it)); b = ex2a(__uint_as_float(__float_as_uint(s[i + 1]) ^ it));
        pk[i >> 1] = __float_as_uint(a) ^ __float_as_uint(b);
      } else {                    // 7: half2 exponentials: pack first, ex2.approx.f16x2, sums in half2 pairs
        const __half2 xh = __floats2half2_rn(fmaf(s[i], scale, negm), fmaf(s[i + 1], scale, negm));
        uint32_t xi = *reinterpret_cast<const uint32_t *>(&xh), yo;
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(yo) : "r"(xi));
        pk[i >> 1] = yo;
      }
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= pk[i];
    negm += 1e-3f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(sum0 + sum1);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / iters;
}

int main() {
  float *in; uint32_t *out; long long *cyc;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int mode = 0; mode < 8; ++mode)
    for (int warps = 4; warps <= 12; warps += 4) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 1) k<1><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 2) k<2><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 3) k<3><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 4) k<4><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 5) k<5><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 6) k<6><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        if (mode == 7) k<7><<<148, warps * 32>>>(in, out, cyc, 0.18f, -1.f, 200);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const char *nm[8] = {"full mix", "MUFU chain", "no MUFU", "FFMA+MUFU", "FFMA+MUFU+pack", "FFMA+MUFU+sum", "MUFU indep", "f16x2 ex2"};
      printf("mode %d (%-14s) warps per scheduler %d: %5lld cycles per 128-score phase per warp = %6.1f per warp-phase and scheduler (%s)\n",
             mode, nm[mode], warps / 4, h, (double)h / (warps / 4), cudaGetErrorString(cudaGetLastError()));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
