// Hardware probe for the 2-CTA (cta_group::2) form of tcgen05: M = 256 across a CTA pair, each CTA stages ITS 128 rows
// of A (MN-major SWIZZLE_128B_BASE32B) and ITS half of B (N/2 rows, K-major SWIZZLE_128B); the leader CTA issues the
// MMAs, tcgen05.commit multicasts completion to both CTAs, each CTA reads its own 128 TMEM lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I dnn-based_source_separation_b200/csrc -o tools/umma_unit2 tools/umma_unit2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "ctn_umma_ptx.cuh"

namespace p2 {
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit2_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
}  // namespace p2

struct Case { int N, K, b_split; const char* name; };  // b_split: 1 = each CTA holds N/2 rows of B; 0 = both hold all N rows

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) k_probe2(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                          float* __restrict__ D, Case c) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(sm + 64);
  float* sA = reinterpret_cast<float*>(sm + 1024);           // 16 KB: this CTA's 128 rows x K
  float* sB = reinterpret_cast<float*>(sm + 1024 + 16384);   // up to 32 KB
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  if (tid == 0) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::fence_mbar_init(); }
  if (warp == 0) p2::tmem_alloc2(ptx::smem_u32(tptr), 256);
  // stage A: rows rank*128 + m
  for (int i = tid; i < 128 * c.K; i += 128) {
    const int m = i % 128, k = i / 128;
    const int kg = k >> 2, r = k & 3, atom = m >> 5, c32 = (m & 31) >> 3, e = m & 7;
    sA[(kg * 2048 + atom * 512) / 4 + r * 32 + ((c32 ^ r) << 3) + e] = A[(rank * 128 + m) * c.K + k];
  }
  // stage B: this CTA's rows
  const int nrows = c.b_split ? c.N / 2 : c.N;
  const int n0 = c.b_split ? rank * (c.N / 2) : 0;
  for (int i = tid; i < nrows * c.K; i += 128) {
    const int n = i / c.K, k = i % c.K;
    sB[(n >> 3) * 256 + (n & 7) * 32 + ((((k >> 2) ^ (n & 7))) << 2) + (k & 3)] = Bm[(n0 + n) * c.K + k];
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *tptr;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = ptx::make_idesc_tf32(256, c.N, 1, 0);
    for (int kk = 0; kk < c.K / 8; ++kk) {
      const uint64_t da = ptx::make_smem_desc(ptx::smem_u32(sA) + kk * 4096, 512, 2048, 1);
      const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(sB) + kk * 32, 16, 1024, 2);
      p2::mma2_tf32(tmem, da, db, idesc, kk ? 1u : 0u);
    }
    p2::commit2_multicast(ptx::smem_u32(&bars[0]), (uint16_t)3);
  }
  ptx::mbar_wait(ptx::smem_u32(&bars[0]), 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < c.N; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)(rank * 128 + tid) * c.N + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 0) p2::tmem_dealloc2(tmem, 256);
}

static bool run_case(const Case& c) {
  const int M = 256;
  std::vector<float> A(M * c.K), B(c.N * c.K), Dref(M * c.N, 0.f), D(M * c.N, -1.f);
  for (int m = 0; m < M; ++m) for (int k = 0; k < c.K; ++k) A[m * c.K + k] = float((m + 3 * k) % 13 - 6);
  for (int n = 0; n < c.N; ++n) for (int k = 0; k < c.K; ++k) B[n * c.K + k] = float((n * 5 + k) % 7 - 3);
  for (int m = 0; m < M; ++m) for (int n = 0; n < c.N; ++n) { float s = 0; for (int k = 0; k < c.K; ++k) s += A[m * c.K + k] * B[n * c.K + k]; Dref[m * c.N + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  const size_t smem = 1024 + 1024 + 16384 + 32768;
  cudaFuncSetAttribute(k_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_probe2<<<2, 128, smem>>>(dA, dB, dD, c);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", c.name, cudaGetErrorString(e)); return false; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, bad_lo = 0; double maxerr = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs((double)D[i] - Dref[i]); if (d > 0 || d != d) { ++bad; if (i < D.size() / 2) ++bad_lo; } if (d > maxerr) maxerr = d; }
  printf("[%-40s] N=%d K=%d b_split=%d : mismatches %d/%zu (rows<128: %d) maxerr %.1f\n", c.name, c.N, c.K, c.b_split, bad, D.size(), bad_lo, maxerr);
  if (bad) {
    for (int row : {0, 1, 128, 129, 200}) {
      printf("    D[%3d][0..7] got:", row); for (int j = 0; j < 8; ++j) printf(" %6.1f", D[row * c.N + j]);
      printf("  | cols N/2..:"); for (int j = 0; j < 4; ++j) printf(" %6.1f", D[row * c.N + c.N / 2 + j]);
      printf("\n    %11s ref:", ""); for (int j = 0; j < 8; ++j) printf(" %6.1f", Dref[row * c.N + j]);
      printf("  |            "); for (int j = 0; j < 4; ++j) printf(" %6.1f", Dref[row * c.N + c.N / 2 + j]);
      printf("\n");
    }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad == 0;
}

int main() {
  std::vector<Case> cases = {{64, 8, 1, "2CTA N=64 K=8, B split N/2 per CTA"}, {64, 32, 1, "2CTA N=64 K=32, B split"},
                             {256, 32, 1, "2CTA N=256 K=32, B split"}, {64, 8, 0, "2CTA N=64 K=8, B replicated (diag)"}};
  int ok = 0;
  for (const Case& c : cases) ok += run_case(c);
  printf("passed %d / %zu\n", ok, cases.size());
  return 0;
}
