// Issue cost of the FP32 forms the Lagrangian kernels are made of, measured per SM sub-partition (SMSP):
//   cycles per warp-instruction = elapsed SM cycles * 4 SMSPs / (warps per SM * instructions per warp)   at saturation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_forms fma_forms.cu ; run: ./fma_forms
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long f2;
#define CHAINS 8
#define ITERS 2048

template <int FORM>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cycles, float cb, float cc) {
  f2 a[CHAINS];
  float s[CHAINS];
  f2 b, c;
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(cb + threadIdx.x * 1e-9f), "f"(cb - threadIdx.x * 2e-9f));     // true pairs (distinct halves)
  asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(cc + threadIdx.x * 1e-9f), "f"(cc * 0.5f + threadIdx.x * 1e-9f));
  f2 bb, cc2;                                                                                                // broadcast pairs (x, x)
  asm("mov.b64 %0, {%1,%1};" : "=l"(bb) : "f"(cb + threadIdx.x * 1e-9f));
  asm("mov.b64 %0, {%1,%1};" : "=l"(cc2) : "f"(cc + threadIdx.x * 1e-9f));
  float sb = cb + threadIdx.x * 1e-9f, sc = cc + threadIdx.x * 1e-9f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(1.0f + i + threadIdx.x));
    s[i] = 1.0f + i + threadIdx.x;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (FORM == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(sb), "f"(sc));                       // FFMA R,R,R
      if (FORM == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));                        // FFMA2 R,R,R
      if (FORM == 2) asm volatile("fma.rn.f32x2 %0, %0, %0, %1;" : "+l"(a[i]) : "l"(c));                                // FFMA2 R,R(same),R
      if (FORM == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));                                    // FADD2 R,R
      if (FORM == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));                                    // FMUL2 R,R
      if (FORM == 5) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(s[i]) : "f"(sb));                          // FFMA R,R,imm
      if (FORM == 6) asm volatile("{ .reg .b64 t; mov.b64 t, {%1,%1}; fma.rn.f32x2 %0, %0, t, %2; }" : "+l"(a[i]) : "f"(cb), "l"(c));   // FFMA2 R,const,R (uniform pair)
      if (FORM == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(sb));                                     // FADD R,R
      if (FORM == 8) asm volatile("fma.rn.sat.f32 %0, %0, %1, 0f3F000000;" : "+f"(s[i]) : "f"(sb));                      // FFMA.SAT R,R,imm
      if (FORM == 9) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i]));                                         // MUFU.RSQ
      if (FORM == 12) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb), "l"(cc2));                     // FFMA2 R, R.F32, R.F32 (broadcast)
      if (FORM == 13) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb), "l"(c));                       // FFMA2 R, R.F32, R
      if (FORM == 10) {                                                                                                  // FFMA2 + MUFU mix (1 : 0.17 like the kernel)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
        if (i == 0) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
      }
      if (FORM == 11) {                                                                                                  // FFMA2 + one ALU op per FFMA2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+f"(s[i]) : "f"(sb), "f"(sc));
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    acc += lo + hi + s[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int FORM>
void run(const char* name, int instr_per_chain_step) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  for (int threads : {128, 256, 512, 1024}) {
    bench<FORM><<<148, threads>>>(out, cyc, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    bench<FORM><<<148, threads>>>(out, cyc, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double warps_per_smsp = threads / 32.0 / 4.0;
    const double instr = (double)ITERS * CHAINS * instr_per_chain_step;
    printf("%-34s warps/SMSP %4.1f  cycles per warp-instruction per SMSP %.3f   (%s)\n", name, warps_per_smsp,
           avg / (instr * warps_per_smsp), cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA  R,R,R", 1);
  run<5>("FFMA  R,R,imm", 1);
  run<8>("FFMA.SAT R,R,imm", 1);
  run<7>("FADD  R,R", 1);
  run<1>("FFMA2 R,R,R", 1);
  run<2>("FFMA2 R,R(same),R", 1);
  run<12>("FFMA2 R,R.F32,R.F32 (broadcast)", 1);
  run<13>("FFMA2 R,R.F32,R", 1);
  run<6>("FFMA2 R,(c,c),R", 1);
  run<3>("FADD2 R,R", 1);
  run<4>("FMUL2 R,R", 1);
  run<9>("MUFU.RSQ", 1);
  run<10>("FFMA2 x8 + MUFU x1 (per 8 FFMA2)", 1);
  run<11>("FFMA2 + LOP3 (cycles per pair)", 1);
  return 0;
}
