// gpu_ai_b200/csrc/microbench.cu -- integer-pipe issue-rate microbenchmarks.
//
// The playout kernels are bound by the INT32 issue rate of the SM, whose
// per-pipe rates are not documented offline.  These kernels measure them on the box so that
// bench.py can report `roofline.peak` from a measurement instead of a guess.  Each thread runs
// 8 mutually dependent-light chains of one instruction class (inline PTX so that the compiler
// cannot merge or fold them); the chip is filled with 148 x 16 warps x ... resident threads.
#include "kernels.cuh"

namespace b2p {

namespace {

constexpr int kChains = 8;
constexpr int kUnroll = 16;

template <int WHICH>
__device__ __forceinline__ void one_round(uint32_t (&x)[kChains]) {
#pragma unroll
  for (int i = 0; i < kChains; i++) {
    uint32_t &a = x[i];
    const uint32_t b = x[(i + 1) & (kChains - 1)], c = x[(i + 3) & (kChains - 1)];
    if (WHICH == 0) {
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    } else if (WHICH == 1) {
      asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));  // -> IADD3
    } else if (WHICH == 2) {
      asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));
    } else if (WHICH == 3) {
      asm volatile("popc.b32 %0, %0;" : "+r"(a));
    } else if (WHICH == 4) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    } else if (WHICH == 5) {
      if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
      else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    } else if (WHICH == 6) {
      asm volatile("brev.b32 %0, %0;" : "+r"(a));
    } else if (WHICH == 7) {
      asm volatile("bfind.u32 %0, %0;" : "+r"(a));
    } else if (WHICH == 9) {
      asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    } else if (WHICH == 10) {
      // right shift on the FMA pipe (x >> 4 == mulhi(x, 2^28)) interleaved with LOP3 on the ALU pipe
      if (i & 1) asm volatile("mul.hi.u32 %0, %0, 0x10000000;" : "+r"(a));
      else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    } else if (WHICH == 11) {
      if (i & 1) asm volatile("shr.u32 %0, %0, 4;" : "+r"(a));
      else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    } else if (WHICH == 8) {
      // 3:1 LOP3:IMAD, the mix of the playout kernel
      if ((i & 3) == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
      else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    }
  }
}

template <int WHICH>
__global__ void __launch_bounds__(256) issue_kernel(int iters, uint32_t *sink) {
  uint32_t x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; i++) x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < kUnroll; u++) one_round<WHICH>(x);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < kChains; i++) acc ^= x[i];
  if (acc == 0x12345678u) sink[0] = acc;  // keep the chains alive
}

template <int WHICH>
cudaError_t run(int iters, int grid, uint32_t *sink, cudaStream_t stream) {
  issue_kernel<WHICH><<<grid, 256, 0, stream>>>(iters, sink);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_microbench(int which, int iters, int sm_count, uint32_t *sink, cudaStream_t stream,
                              double *thread_ops) {
  const int grid = sm_count * 8;  // 8 x 256 threads = 2048 threads per SM: full occupancy
  *thread_ops = (double)grid * 256.0 * (double)iters * kUnroll * kChains;
  switch (which) {
    case 0: return run<0>(iters, grid, sink, stream);
    case 1: return run<1>(iters, grid, sink, stream);
    case 2: return run<2>(iters, grid, sink, stream);
    case 3: return run<3>(iters, grid, sink, stream);
    case 4: return run<4>(iters, grid, sink, stream);
    case 5: return run<5>(iters, grid, sink, stream);
    case 6: return run<6>(iters, grid, sink, stream);
    case 7: return run<7>(iters, grid, sink, stream);
    case 8: return run<8>(iters, grid, sink, stream);
    case 9: return run<9>(iters, grid, sink, stream);
    case 10: return run<10>(iters, grid, sink, stream);
    case 11: return run<11>(iters, grid, sink, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace b2p
