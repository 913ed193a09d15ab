// Pipe-rate microbenchmark for the instructions the CWBVH node test is made of (development tool).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_pipes.cu -o gpurun_out/ubench_pipes
// Prints warp-instructions per clock per SM for each op (8 independent chains per thread, 32 warps/SM).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#include <cuda_fp16.h>

constexpr int ITER = 2048, CH = 8;

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, float fa, float fb)
{
  uint32_t x[CH];
  float f[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    x[c] = seed + threadIdx.x * 7 + c * 13;
    f[c] = (float)(threadIdx.x + c);
  }
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (OP == 0) {  // I2F.U8 (byte 1) -> feeds back as bits
        f[c] = (float)((x[c] >> 8) & 0xffu);
        x[c] = __float_as_uint(f[c]) + i;   // + IADD
      } else if (OP == 1) {  // PRMT + FADD
        f[c] = __uint_as_float(__byte_perm(x[c], 0x4B000000u, 0x7541u)) - 8388608.0f;
        x[c] = __float_as_uint(f[c]) + i;
      } else if (OP == 2) {  // IADD only (baseline for 0/1)
        x[c] = x[c] * 1 + i;
        asm volatile("" : "+r"(x[c]));
      } else if (OP == 3) {  // FFMA
        f[c] = fmaf(f[c], fa, fb);
      } else if (OP == 4) {  // FMNMX
        f[c] = fminf(f[c], __uint_as_float(x[c]));
        asm volatile("" : "+f"(f[c]));
      } else if (OP == 5) {  // LOP3
        x[c] = (x[c] & seed) ^ i;
        asm volatile("" : "+r"(x[c]));
      } else if (OP == 6) {  // PRMT
        x[c] = __byte_perm(x[c], seed, 0x7541u);
        asm volatile("" : "+r"(x[c]));
      } else if (OP == 7) {  // MUFU.RCP
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
      } else if (OP == 8) {  // FMNMX3
        f[c] = fminf(fminf(f[c], fa), __uint_as_float(x[c]));
        asm volatile("" : "+f"(f[c]));
      } else if (OP == 9) {  // I2F full 32-bit
        f[c] = (float)x[c];
        x[c] = __float_as_uint(f[c]) + i;
      } else if (OP == 10) {  // SEL
        x[c] = (i & 1) ? x[c] : seed;
        asm volatile("" : "+r"(x[c]));
      } else if (OP == 11) {  // FSETP + predicated LOP (hit mask update)
        if (f[c] <= fa) x[c] |= seed << c;
        asm volatile("" : "+r"(x[c]));
      } else if (OP == 12) {  // HADD2.F32 (half -> float) + feedback
        f[c] = __half2float(__ushort_as_half((unsigned short)(x[c] & 0xffffu)));
        x[c] = __float_as_uint(f[c]) + i;
      } else if (OP == 13) {  // bf16 -> float via shift (IMAD.U32 / SHF)
        f[c] = __uint_as_float(x[c] << 16);
        x[c] = __float_as_uint(f[c]) + i;
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += x[c] + __float_as_uint(f[c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, uint32_t* d_out, int sms, double mhz)
{
  const int blocks = sms * 4, threads = 256;  // 32 warps / SM
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<OP><<<blocks, threads>>>(d_out, 12345u, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(d_out, 12345u, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_iters = (double)blocks * threads / 32 * ITER * CH;
  const double cycles = ms * 1e-3 * mhz * 1e6;
  printf("%-28s %8.3f ms  %6.2f warp-iterations/clk/SM  (%5.2f clk per warp-iteration per SMSP)\n", name, ms,
         warp_iters / cycles / sms, cycles * sms * 4 / warp_iters);
}

int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1e3;
  printf("%s, %d SMs, %.0f MHz (nominal max; iterations = op + feedback as listed)\n", p.name, p.multiProcessorCount, mhz);
  uint32_t* d;
  cudaMalloc(&d, sizeof(uint32_t) * p.multiProcessorCount * 4 * 256);
  const int s = p.multiProcessorCount;
  run<2>("IADD (feedback only)", d, s, mhz);
  run<0>("I2F.U8 + IADD", d, s, mhz);
  run<9>("I2F.U32 + IADD", d, s, mhz);
  run<1>("PRMT + FADD + IADD", d, s, mhz);
  run<12>("HADD2.F32 + IADD", d, s, mhz);
  run<13>("SHL16 + IADD", d, s, mhz);
  run<3>("FFMA", d, s, mhz);
  run<4>("FMNMX", d, s, mhz);
  run<8>("FMNMX3", d, s, mhz);
  run<5>("LOP3 x2", d, s, mhz);
  run<6>("PRMT", d, s, mhz);
  run<10>("SEL", d, s, mhz);
  run<11>("FSETP + @p LOP3", d, s, mhz);
  run<7>("MUFU.RCP", d, s, mhz);
  return 0;
}
