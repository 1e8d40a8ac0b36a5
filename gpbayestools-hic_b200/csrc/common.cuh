// Shared device helpers for the gpbt kernels (sm_100a, FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpbt {

constexpr int kWarp = 32;

// D(8x8) += A(8x4, row) * B(4x8, col), FP64.  On sm_100a this is the only FP64 tensor shape the
// hardware has (the m16n8k{4,8,16} PTX shapes are lowered to several DMMA.8x8x4 by ptxas).
// Fragment ownership for lane l:  A[l/4][l%4],  B[l%4][l/4],  C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 16-byte read-only load of two doubles (immutable state only: Linv, Xs, A ...)
__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__host__ __device__ __forceinline__ int64_t round_up(int64_t x, int64_t a) {
  return (x + a - 1) / a * a;
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-stream-
// serialisation attribute is scheduled while its predecessor drains; its CTAs sit at the wait until the
// predecessor grid has completed and its writes are visible.  Hides launch latency and ramp-up along
// chains of small dependent launches.  Both instructions are no-ops without the launch attribute; a
// kernel launched WITH the attribute must execute the wait before it touches anything a predecessor wrote.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }


// ---- TMA bulk copy (global -> shared) completing on an mbarrier --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace gpbt
