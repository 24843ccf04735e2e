// Pipe-overlap microbenchmark for sm_100a: how do IMAD.WIDE (FMA-heavy), IADD3/IADD3.X/LOP3 (ALU) and DFMA (FP64)
// share a sub-partition's issue/dispatch bandwidth? Every number the accumulate kernel's design rests on is measured
// here instead of assumed. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
//
// Each thread runs 8 independent chains of the "main" instruction and, per main instruction, K "side" instructions
// on 8 other independent chains. Reported: main-instruction lanes per clock per SM, and elapsed time relative to the
// run without side instructions. If the side pipe overlapped perfectly the ratio stays 1.0 until the side pipe
// itself saturates.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

enum Main { WIDE = 0, WIDEX = 1, DFMA = 2, NONE = 3 };
enum Side { S_NONE = 0, S_IADD3X = 1, S_LOP3 = 2, S_DFMA = 3, S_WIDE = 4, S_IADD3 = 5 };

template <int MAIN, int SIDE, int K>
__global__ void __launch_bounds__(128) k(int iters, uint32_t seed, uint64_t* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = (seed * 40503u + tid * 2654435761u) | 1u;
  uint64_t c[8];
  uint32_t s[8];
  double d[8];
  const double db = 1.0 + (double)(tid & 255) * 1e-9, da = 0.999999 + (double)(tid & 63) * 1e-10;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c[i] = ((uint64_t)(tid + i) << 32) | (seed + 77u * i);
    s[i] = tid * 31u + i;
    d[i] = 1.0 + i + tid * 1e-6;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MAIN == WIDE) {
          uint32_t a = (uint32_t)c[(i + 1) & 7];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i]) : "r"(a), "r"(b));
        } else if (MAIN == WIDEX) {  // carry-chained pair: lo/hi with carry-in and carry-out (IMAD.WIDE.U32.X)
          uint32_t a = (uint32_t)c[(i + 1) & 7];
          uint32_t lo = (uint32_t)c[i], hi = (uint32_t)(c[i] >> 32);
          if (i == 0) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
          else asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(a), "r"(b));
          c[i] = ((uint64_t)hi << 32) | lo;
        } else if (MAIN == DFMA) {
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
        }
#pragma unroll
        for (int kk = 0; kk < K; kk++) {
          const int j = (i + kk) & 7;
          if (SIDE == S_IADD3X) {
            if (kk == 0) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(s[j]) : "r"(b));
            else asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(s[j]) : "r"(b));
          } else if (SIDE == S_LOP3) {
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(s[j]) : "r"(b), "r"(s[(j + 3) & 7]));
          } else if (SIDE == S_IADD3) {
            asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(s[j]) : "r"(b), "r"(s[(j + 3) & 7]));
          } else if (SIDE == S_DFMA) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[j]) : "d"(da), "d"(db));
          } else if (SIDE == S_WIDE) {
            uint32_t a = (uint32_t)c[(j + 1) & 7];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[j]) : "r"(a), "r"(b));
          }
        }
      }
    }
  }
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= c[i] ^ s[i] ^ (uint64_t)__double_as_longlong(d[i]);
  sink[tid] = acc;
}

static int sms = 0;
static double clock_ghz = 0;
static uint64_t* sink;

template <int MAIN, int SIDE, int K>
double run(const char* name, int warps_per_sm, double base_ms) {
  const int threads = 128, blocks = sms * warps_per_sm / 4;
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MAIN, SIDE, K><<<blocks, threads>>>(iters / 10, 1u, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MAIN, SIDE, K><<<blocks, threads>>>(iters, 1u, sink);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double main_ops = (double)blocks * threads * iters * 64.0 * (MAIN == NONE ? 0 : 1);
  const double side_ops = (double)blocks * threads * iters * 64.0 * K * (SIDE == S_NONE ? 0 : 1);
  const double cyc = ms * 1e-3 * clock_ghz * 1e9;
  printf("%-44s warps/SM %2d  %8.3f ms  main %6.1f lanes/clk/SM  side %6.1f lanes/clk/SM  x%.3f\n", name, warps_per_sm,
         ms, main_ops / cyc / sms, side_ops / cyc / sms, base_ms > 0 ? ms / base_ms : 1.0);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  clock_ghz = khz * 1e-6;
  printf("%s, %d SMs, %.3f GHz (max; lanes/clk assume it)\n", p.name, sms, clock_ghz);
  cudaMalloc(&sink, (size_t)sms * 16 * 128 * 8 * 8);
  for (int w : {16, 32}) {
    double t = run<WIDE, S_NONE, 0>("IMAD.WIDE alone", w, 0);
    run<WIDE, S_IADD3X, 1>("IMAD.WIDE + 1 IADD3.X (carry chain)", w, t);
    run<WIDE, S_IADD3X, 2>("IMAD.WIDE + 2 IADD3.X", w, t);
    run<WIDE, S_IADD3X, 3>("IMAD.WIDE + 3 IADD3.X", w, t);
    run<WIDE, S_LOP3, 1>("IMAD.WIDE + 1 LOP3", w, t);
    run<WIDE, S_LOP3, 2>("IMAD.WIDE + 2 LOP3", w, t);
    run<WIDE, S_IADD3, 1>("IMAD.WIDE + 1 IADD3 (3-input add)", w, t);
    run<WIDE, S_DFMA, 1>("IMAD.WIDE + 1 DFMA", w, t);
    run<WIDE, S_DFMA, 2>("IMAD.WIDE + 2 DFMA", w, t);
    double tx = run<WIDEX, S_NONE, 0>("IMAD.WIDE.X carry chain alone", w, 0);
    run<WIDEX, S_LOP3, 1>("IMAD.WIDE.X + 1 LOP3", w, tx);
    double td = run<DFMA, S_NONE, 0>("DFMA alone", w, 0);
    run<DFMA, S_LOP3, 1>("DFMA + 1 LOP3", w, td);
    run<NONE, S_IADD3X, 1>("IADD3.X alone (x1)", w, 0);
    run<NONE, S_LOP3, 1>("LOP3 alone (x1)", w, 0);
    run<NONE, S_IADD3, 1>("IADD3 alone (x1)", w, 0);
  }
  return 0;
}
